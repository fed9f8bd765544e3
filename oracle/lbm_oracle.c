/*
 * oracle/lbm_oracle.c -- plain-C restatement of the reference's per-timestep path.
 * TEST INFRASTRUCTURE ONLY (see lbm_oracle.h for who may load it and how it is pinned).
 *
 * Every function names the reference lines it follows and keeps their floating-point
 * evaluation order, so that a strict-IEEE build of this file reproduces a strict-IEEE build of
 * the reference bit for bit (tests/test_oracle_pins.py).  Compile with
 *   gcc -O2 -ffp-contract=off -fno-fast-math
 * Reference paths are relative to /root/reference.
 */
#include "lbm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define Q 9

/* include/LBMConfig.h:13-34 */
static const int CX[Q] = {0, 1, 0, -1, 0, 1, -1, -1, 1};
static const int CY[Q] = {0, 0, 1, 0, -1, 1, 1, -1, -1};
static const double W[Q] = {4.0 / 9.0,  1.0 / 9.0,  1.0 / 9.0,  1.0 / 9.0, 1.0 / 9.0,
                            1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0, 1.0 / 36.0};
static const int OPP[Q] = {0, 3, 4, 1, 2, 7, 8, 5, 6};

struct oracle_state {
    oracle_params p;
    int tnx, tny; /* padded sizes */
    int cyl_x, cyl_y, cyl_r;
    double* fc; /* f_current, padded AoS */
    double* fn; /* f_next */
    double *rho, *ux, *uy; /* interior */
    unsigned char* solid;  /* interior */
    unsigned char* solid_pad; /* padded, global mask (0 outside the global domain) */
};

static size_t fidx(const oracle_state* s, int gx, int gy, int i) { /* LBMGrid.h:105-107 */
    return ((size_t)gy * (size_t)s->tnx + (size_t)gx) * Q + (size_t)i;
}
static size_t iidx(const oracle_state* s, int x, int y) { /* LBMGrid.h:109-111 */
    return (size_t)y * (size_t)s->p.lnx + (size_t)x;
}

oracle_state* oracle_create(const oracle_params* p) {
    oracle_state* s = (oracle_state*)calloc(1, sizeof(oracle_state));
    if (!s) return NULL;
    s->p = *p;
    s->tnx = p->lnx + 2;
    s->tny = p->ny + 2;
    /* LBMConfig.h:61-65: truncating casts */
    s->cyl_x = (int)(p->cylinder_x * p->gnx);
    s->cyl_y = (int)(p->cylinder_y * p->ny);
    s->cyl_r = (int)(p->cylinder_radius * p->ny);
    const size_t nf = (size_t)s->tnx * s->tny * Q;
    const size_t ni = (size_t)p->lnx * p->ny;
    s->fc = (double*)calloc(nf, sizeof(double));
    s->fn = (double*)calloc(nf, sizeof(double));
    s->rho = (double*)malloc(ni * sizeof(double));
    s->ux = (double*)malloc(ni * sizeof(double));
    s->uy = (double*)malloc(ni * sizeof(double));
    s->solid = (unsigned char*)calloc(ni, 1);
    s->solid_pad = (unsigned char*)calloc((size_t)s->tnx * s->tny, 1);
    /* LBMGrid.h:73-76: rho=1, ux=uy=0 */
    for (size_t k = 0; k < ni; ++k) { s->rho[k] = 1.0; s->ux[k] = 0.0; s->uy[k] = 0.0; }
    return s;
}

void oracle_destroy(oracle_state* s) {
    if (!s) return;
    free(s->fc); free(s->fn); free(s->rho); free(s->ux); free(s->uy); free(s->solid); free(s->solid_pad);
    free(s);
}

/* LBMUtils.h:9-12 */
static double eq_rest(double rho, double ux, double uy) {
    const double u_sq = ux * ux + uy * uy;
    return W[0] * rho * (1.0 - 1.5 * u_sq);
}

/* LBMUtils.h:22-65, lane by lane: ((1 + 3cu) - 1.5u^2) + 4.5cu^2, times (w*rho). */
static void eq_moving(double rho, double ux, double uy, double* f_eq /* [8] -> i=1..8 */) {
    const double u_sq = ux * ux + uy * uy;
    const double term3 = 1.5 * u_sq;
    for (int i = 1; i < Q; ++i) {
        const double w = (i <= 4) ? 1.0 / 9.0 : 1.0 / 36.0;
        const double ci_u = (double)CX[i] * ux + (double)CY[i] * uy;
        const double ci_u_sq = ci_u * ci_u;
        const double term1 = 3.0 * ci_u;
        const double term2 = 4.5 * ci_u_sq;
        const double bracket = ((1.0 + term1) - term3) + term2;
        f_eq[i - 1] = (w * rho) * bracket;
    }
}

/* LBMGrid.h:152-183 (mask) then LBMGrid.h:185-246 (populations + macroscopic fields). */
int oracle_initialise(oracle_state* s) {
    const int lnx = s->p.lnx, ny = s->p.ny;
    int count = 0;
    for (int gy = 0; gy < s->tny; ++gy)
        for (int gx = 0; gx < s->tnx; ++gx) {
            const int global_x = s->p.x_start + gx - 1, global_y = gy - 1;
            unsigned char m = 0;
            if (global_x >= 0 && global_x < s->p.gnx && global_y >= 0 && global_y < ny) {
                const double dx = global_x - s->cyl_x;
                const double dy = global_y - s->cyl_y;
                const double dist_sq = dx * dx + dy * dy;
                if (dist_sq <= s->cyl_r * s->cyl_r) m = 1;
            }
            s->solid_pad[(size_t)gy * s->tnx + gx] = m;
            if (gx >= 1 && gx <= lnx && gy >= 1 && gy <= ny) {
                s->solid[iidx(s, gx - 1, gy - 1)] = m;
                count += m;
            }
        }

    const double u_in = s->p.inlet_velocity;
    double e[8];
    const double e0 = eq_rest(1.0, u_in, 0.0);
    eq_moving(1.0, u_in, 0.0, e);
    for (int gy = 0; gy < s->tny; ++gy)
        for (int gx = 0; gx < s->tnx; ++gx) {
            double* fc = &s->fc[fidx(s, gx, gy, 0)];
            double* fn = &s->fn[fidx(s, gx, gy, 0)];
            fc[0] = e0;
            for (int i = 1; i < Q; ++i) fc[i] = e[i - 1];
            for (int i = 0; i < Q; ++i) fn[i] = fc[i];
        }
    const double r0 = eq_rest(1.0, 0.0, 0.0);
    double r[8];
    eq_moving(1.0, 0.0, 0.0, r);
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < lnx; ++x) {
            const size_t k = iidx(s, x, y);
            s->rho[k] = 1.0;
            s->uy[k] = 0.0;
            if (!s->solid[k]) {
                s->ux[k] = u_in;
            } else {
                s->ux[k] = 0.0;
                double* fc = &s->fc[fidx(s, x + 1, y + 1, 0)];
                double* fn = &s->fn[fidx(s, x + 1, y + 1, 0)];
                fc[0] = r0;
                for (int i = 1; i < Q; ++i) fc[i] = r[i - 1];
                for (int i = 0; i < Q; ++i) fn[i] = fc[i];
            }
        }
    return count;
}

/* LBMSolver.h:84-126 */
void oracle_collide(oracle_state* s) {
    const double tau_inv = 1.0 / s->p.tau;
    for (int y = 0; y < s->p.ny; ++y)
        for (int x = 0; x < s->p.lnx; ++x) {
            const size_t k = iidx(s, x, y);
            if (s->solid[k]) continue;
            const double* f_curr = &s->fc[fidx(s, x + 1, y + 1, 0)];
            double* f_next = &s->fn[fidx(s, x + 1, y + 1, 0)];
            double rho_val = 0.0, ux_val = 0.0, uy_val = 0.0;
            for (int i = 0; i < Q; ++i) {
                rho_val += f_curr[i];
                ux_val += CX[i] * f_curr[i];
                uy_val += CY[i] * f_curr[i];
            }
            ux_val /= rho_val;
            uy_val /= rho_val;
            s->rho[k] = rho_val;
            s->ux[k] = ux_val;
            s->uy[k] = uy_val;
            const double u_sq = ux_val * ux_val + uy_val * uy_val;
            for (int i = 0; i < Q; ++i) {
                const double ci_u = CX[i] * ux_val + CY[i] * uy_val;
                const double f_eq_i = W[i] * rho_val * (1.0 + 3.0 * ci_u + 4.5 * ci_u * ci_u - 1.5 * u_sq);
                f_next[i] = f_curr[i] - tau_inv * (f_curr[i] - f_eq_i);
            }
        }
}

/* LBMIO.h:114-162.  Same (y, x, i) accumulation order over solid cells.  A link belongs to the
 * slab that owns its FLUID end, and the solid end may sit in a neighbouring slab (global mask):
 * unlike the reference no link is dropped at a slab face, so slab partials sum
 * to the 1-rank value.  With one slab this is the reference loop exactly. */
void oracle_forces(const oracle_state* s, double* fx, double* fy) {
    double local_fx = 0.0, local_fy = 0.0;
    const int lnx = s->p.lnx, ny = s->p.ny;
    for (int y = 0; y < ny; ++y)
        for (int x = -1; x <= lnx; ++x) {
            if (!s->solid_pad[(size_t)(y + 1) * s->tnx + (x + 1)]) continue;
            for (int i = 1; i < Q; ++i) {
                const int fluid_x = x - CX[i];
                const int fluid_y = y - CY[i];
                if (fluid_x >= 0 && fluid_x < lnx && fluid_y >= 0 && fluid_y < ny &&
                    !s->solid[iidx(s, fluid_x, fluid_y)]) {
                    const double f_i = s->fn[fidx(s, fluid_x + 1, fluid_y + 1, i)];
                    local_fx += 2.0 * CX[i] * f_i;
                    local_fy += 2.0 * CY[i] * f_i;
                }
            }
        }
    *fx = local_fx;
    *fy = local_fy;
}

/* LBMGrid.h:448-466 with the 1-rank behaviour of SURVEY.md F4: at a physical domain edge the
 * never-written, zero-initialised receive buffer is unpacked into the ghost column, rows
 * 1..ny, all nine populations, every step.  Ghost rows and corner ghosts are never touched. */
void oracle_edge_ghosts(oracle_state* s) {
    const int west_edge = (s->p.x_start == 0);
    const int east_edge = (s->p.x_start + s->p.lnx == s->p.gnx);
    for (int y = 0; y < s->p.ny; ++y) {
        if (west_edge) memset(&s->fn[fidx(s, 0, y + 1, 0)], 0, Q * sizeof(double));
        if (east_edge) memset(&s->fn[fidx(s, s->tnx - 1, y + 1, 0)], 0, Q * sizeof(double));
    }
}

/* LBMGrid.h:399-417: boundary column of f_next, rows 1..ny, nine populations per cell. */
void oracle_get_halo(const oracle_state* s, int east, double* buf) {
    const int gx = east ? s->p.lnx : 1;
    for (int y = 0; y < s->p.ny; ++y) memcpy(buf + (size_t)y * Q, &s->fn[fidx(s, gx, y + 1, 0)], Q * sizeof(double));
}

/* LBMGrid.h:448-466: into the ghost column on that side. */
void oracle_put_halo(oracle_state* s, int east, const double* buf) {
    const int gx = east ? s->tnx - 1 : 0;
    for (int y = 0; y < s->p.ny; ++y) memcpy(&s->fn[fidx(s, gx, y + 1, 0)], buf + (size_t)y * Q, Q * sizeof(double));
}

/* LBMSolver.h:128-145: every interior cell, solids included. */
void oracle_stream(oracle_state* s) {
    for (int y = 0; y < s->p.ny; ++y)
        for (int x = 0; x < s->p.lnx; ++x) {
            const int gx = x + 1, gy = y + 1;
            for (int i = 0; i < Q; ++i) s->fc[fidx(s, gx, gy, i)] = s->fn[fidx(s, gx - CX[i], gy - CY[i], i)];
        }
}

/* LBMSolver.h:147-265 in the serial (1-thread) order of SURVEY.md F5:
 * bottom wall, top wall, inlet column, outlet column, solid reversal. */
void oracle_boundaries(oracle_state* s) {
    const int lnx = s->p.lnx, ny = s->p.ny;
    const int west_edge = (s->p.x_start == 0);
    const int east_edge = (s->p.x_start + lnx == s->p.gnx);

    for (int x = 0; x < lnx; ++x) { /* :153-164 */
        if (s->solid[iidx(s, x, 0)]) continue;
        double* f = &s->fc[fidx(s, x + 1, 1, 0)];
        f[2] = f[4];
        f[5] = f[7];
        f[6] = f[8];
    }
    for (int x = 0; x < lnx; ++x) { /* :166-176 */
        if (s->solid[iidx(s, x, ny - 1)]) continue;
        double* f = &s->fc[fidx(s, x + 1, ny, 0)];
        f[4] = f[2];
        f[7] = f[5];
        f[8] = f[6];
    }
    if (west_edge) /* :179-207 */
        for (int y = 0; y < ny; ++y) {
            if (s->solid[iidx(s, 0, y)]) continue;
            double* f = &s->fc[fidx(s, 1, y + 1, 0)];
            const double u_in = s->p.inlet_velocity;
            const double v_in = 0.0;
            const double rho_bc = (f[0] + f[2] + f[4] + 2.0 * (f[3] + f[6] + f[7])) / (1.0 - u_in);
            f[1] = f[3] + (2.0 / 3.0) * rho_bc * u_in;
            f[5] = f[7] - 0.5 * (f[2] - f[4]) + (1.0 / 6.0) * rho_bc * u_in;
            f[8] = f[6] + 0.5 * (f[2] - f[4]) + (1.0 / 6.0) * rho_bc * u_in;
            s->rho[iidx(s, 0, y)] = rho_bc;
            s->ux[iidx(s, 0, y)] = u_in;
            s->uy[iidx(s, 0, y)] = v_in;
        }
    if (east_edge) /* :210-236 */
        for (int y = 0; y < ny; ++y) {
            if (s->solid[iidx(s, lnx - 1, y)]) continue;
            double* f = &s->fc[fidx(s, lnx, y + 1, 0)];
            const double rho_out = 1.0;
            const double u_out = -1.0 + (f[0] + f[2] + f[4] + 2.0 * (f[1] + f[5] + f[8])) / rho_out;
            const double v_out = 0.0;
            f[3] = f[1] - (2.0 / 3.0) * rho_out * u_out;
            f[6] = f[8] - 0.5 * (f[2] - f[4]) - (1.0 / 6.0) * rho_out * u_out;
            f[7] = f[5] + 0.5 * (f[2] - f[4]) - (1.0 / 6.0) * rho_out * u_out;
            s->rho[iidx(s, lnx - 1, y)] = rho_out;
            s->ux[iidx(s, lnx - 1, y)] = u_out;
            s->uy[iidx(s, lnx - 1, y)] = v_out;
        }
    for (int y = 0; y < ny; ++y) /* :240-263 */
        for (int x = 0; x < lnx; ++x) {
            if (!s->solid[iidx(s, x, y)]) continue;
            double* f = &s->fc[fidx(s, x + 1, y + 1, 0)];
            double f_temp[Q];
            for (int i = 0; i < Q; ++i) f_temp[i] = f[i];
            for (int i = 0; i < Q; ++i) f[i] = f_temp[OPP[i]];
            s->ux[iidx(s, x, y)] = 0.0;
            s->uy[iidx(s, x, y)] = 0.0;
        }
}

/* LBMGrid.h:285-317: every padded f_current value must be a number within [-1e5, 1e5]. */
int oracle_check_stability(const oracle_state* s) {
    const size_t n = (size_t)s->tnx * s->tny * Q;
    for (size_t k = 0; k < n; ++k) {
        const double v = s->fc[k];
        if (!(v == v)) return 0;
        if (v > 1e5 || v < -1e5) return 0;
    }
    return 1;
}

/* LBMGrid.h:319-344 */
double oracle_max_velocity(const oracle_state* s) {
    double m = 0.0;
    const size_t n = (size_t)s->p.lnx * s->p.ny;
    for (size_t k = 0; k < n; ++k) {
        const double v = s->ux[k] * s->ux[k] + s->uy[k] * s->uy[k];
        if (v > m) m = v;
    }
    return sqrt(m);
}

/* LBMSolver.h:48-76 for a single slab that is the whole domain. */
int oracle_run(oracle_state* s, int t0, int nsteps, double* forces_out, int max_rows, int* n_rows, int* unstable_at) {
    int rows = 0, done = 0;
    if (unstable_at) *unstable_at = -1;
    for (int t = t0; t < t0 + nsteps; ++t) {
        oracle_collide(s);
        if (t % s->p.output_frequency == 0) {
            double fx, fy;
            oracle_forces(s, &fx, &fy);
            /* LBMIO.h:171-185 */
            const double rho_ref = 1.0;
            const double U_ref = s->p.inlet_velocity;
            const double D_ref = 2.0 * s->cyl_r;
            const double q_ref = 0.5 * rho_ref * U_ref * U_ref * D_ref;
            const double cd = (q_ref > 1e-12) ? fx / q_ref : 0.0;
            const double cl = (q_ref > 1e-12) ? fy / q_ref : 0.0;
            if (forces_out && rows < max_rows) {
                double* r = forces_out + (size_t)rows * 5;
                r[0] = (double)t; r[1] = fx; r[2] = fy; r[3] = cd; r[4] = cl;
            }
            ++rows;
        }
        oracle_edge_ghosts(s);
        oracle_stream(s);
        oracle_boundaries(s);
        ++done;
        if (!oracle_check_stability(s)) {
            if (unstable_at) *unstable_at = t;
            break;
        }
    }
    if (n_rows) *n_rows = rows;
    return done;
}

double* oracle_f_current(oracle_state* s) { return s->fc; }
double* oracle_f_next(oracle_state* s) { return s->fn; }
double* oracle_rho(oracle_state* s) { return s->rho; }
double* oracle_ux(oracle_state* s) { return s->ux; }
double* oracle_uy(oracle_state* s) { return s->uy; }
const unsigned char* oracle_solid(const oracle_state* s) { return s->solid; }
size_t oracle_f_count(const oracle_state* s) { return (size_t)s->tnx * s->tny * Q; }
