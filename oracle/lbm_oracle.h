/*
 * oracle/lbm_oracle.h -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.  The
 * product (liblbm_b200.so) never links, loads or calls anything under oracle/.
 *
 * Parity status: PINNED.  lbm_oracle.c is compiled strict-IEEE (-O2 -ffp-contract=off, no
 * fast-math) and is checked in tests/test_oracle_pins.py
 *   - bit-for-bit against oracle/_ref/lbm_ref_strict (the unmodified reference headers compiled
 *     with the same strict flags) whenever /root/reference is present, and
 *   - against the committed fixtures in tests/golden/ (generated from oracle/_ref by
 *     oracle/gen_golden.py) everywhere else.
 * The reference ships no tests or golden vectors of its own.
 *
 * The state is one x-slab [x_start, x_start+lnx) x [0, ny) of a gnx x ny channel.  With
 * x_start=0, lnx=gnx it is exactly the reference's 1-rank job, which is the only
 * self-consistent reference configuration.  Arrays use the reference's layouts
 * (include/LBMGrid.h:105-111): populations padded AoS [(gy*(lnx+2)+gx)*9+i], macroscopic
 * fields interior row-major [y*lnx+x].
 */
#ifndef LBM_ORACLE_H
#define LBM_ORACLE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    double tau;
    double inlet_velocity;
    int gnx; /* global nx */
    int ny;
    int output_frequency;
    double cylinder_x, cylinder_y, cylinder_radius; /* fractions, as LBMConfig.h:44-46 */
    int x_start; /* first global x owned by this slab */
    int lnx;     /* slab width */
} oracle_params;

typedef struct oracle_state oracle_state;

oracle_state* oracle_create(const oracle_params* p);
void oracle_destroy(oracle_state* s);

/* Grid::setup_geometry + Grid::initialise (LBMGrid.h:152-246). Returns solid cells owned. */
int oracle_initialise(oracle_state* s);

/* The phases of one Solver::run iteration (LBMSolver.h:48-76), individually callable so a
 * multi-slab test can put its own halo exchange between them. */
void oracle_collide(oracle_state* s);                             /* LBMSolver.h:84-126 */
void oracle_forces(const oracle_state* s, double* fx, double* fy); /* LBMIO.h:114-162  */
void oracle_edge_ghosts(oracle_state* s);                         /* LBMGrid.h:448-466 at domain edges (F4) */
void oracle_get_halo(const oracle_state* s, int east, double* buf9ny);   /* pack  LBMGrid.h:399-417 */
void oracle_put_halo(oracle_state* s, int east, const double* buf9ny);   /* unpack LBMGrid.h:448-466 */
void oracle_stream(oracle_state* s);                              /* LBMSolver.h:128-145 */
void oracle_boundaries(oracle_state* s);                          /* LBMSolver.h:147-265, serial order */
int oracle_check_stability(const oracle_state* s);                /* LBMGrid.h:285-317; 1 = stable */
double oracle_max_velocity(const oracle_state* s);                /* LBMGrid.h:319-344 */

/* Whole single-slab run loop (LBMSolver.h:48-76).  forces_out receives one row
 * {t, Fx, Fy, C_D, C_L} per output step (LBMIO.h:171-185), up to max_rows.  Returns the number
 * of iterations completed; *unstable_at is the failing timestep or -1.  t0 is the timestep
 * number of the first iteration (0 for a fresh run). */
int oracle_run(oracle_state* s, int t0, int nsteps, double* forces_out, int max_rows, int* n_rows, int* unstable_at);

/* Raw views (valid until destroy). */
double* oracle_f_current(oracle_state* s);
double* oracle_f_next(oracle_state* s);
double* oracle_rho(oracle_state* s);
double* oracle_ux(oracle_state* s);
double* oracle_uy(oracle_state* s);
const unsigned char* oracle_solid(const oracle_state* s); /* interior [y*lnx+x] */
size_t oracle_f_count(const oracle_state* s);             /* (lnx+2)*(ny+2)*9 */

#ifdef __cplusplus
}
#endif
#endif
