// lbm_kernels.cuh -- launch interface of the sm_100a kernels (definitions in lbm_kernels.cu and
// lbm_aa.cu).  Host code (lbm_engine.cu) sees only these plain functions.
#pragma once
#include <cuda_runtime.h>

#include "lbm_layout.h"

namespace lbm {

// What the fix-up / observe kernels need to know about the boundaries of this slab.
struct BcArgs {
    double u_in;
    int inlet;   // Zou-He velocity inlet on interior column 0   (slab touches x = 0, not periodic)
    int outlet;  // Zou-He pressure outlet on column lnx-1        (slab touches x = gnx-1)
    int walls;   // wall reflection on rows 0 and ny-1            (not periodic in y)
    double w[9];  // populations a solid cell keeps for ever: eq(1,0,0)   
    double e[9];  // initial equilibrium eq(1,u_in,0): value of every f_current ghost cell (F4)
};

struct StepArgs {
    const double* src;  // post-collision populations of the previous iteration (or f_current when !pull)
    double* dst;        // post-collision populations of this iteration
    Layout L;
    double tau_inv;
    double Fx, Fy;   // body force (extension); 0 in the reference's channel
    int forced;      // use bgk_forced
    int* first_bad;  // device int, atomicMin'ed with bad_iter when an unstable value is seen
    int bad_iter;    // reference timestep whose check_stability these pulled values belong to
    int write;       // 0: check only, store nothing
    // Per interior column x: rows [x, y) are solid and rows [z, w) are "deep" solid (all eight
    // neighbours solid), when the solid cells of the column form one run; {0,0,0,0} otherwise.
    // Solid cells are not stored (they keep w for ever), deep ones are skipped
    // altogether.  nullptr: no table, every cell is processed and the fix-up resets the solids.
    const int4* cols;
    int col_lo, col_hi;  // only columns in [col_lo, col_hi) have a table entry worth loading
};

// BULK_TB: the temporally blocked kernels of lbm_tb.cuh (several iterations per pass over HBM; the default).
enum BulkVariant { BULK_SCALAR = 0, BULK_VEC2 = 1, BULK_TB = 2 };

// Fused pull + collide over every interior cell, no boundary logic (reference
// include/LBMSolver.h:84-145 minus the solid `continue`).  pull=false is the very first
// iteration, which collides the initial f_current in place of a pulled state.
cudaError_t launch_bulk(int variant, bool pull, const StepArgs& a, cudaStream_t s, int x_begin = 0, int x_end = -1);

// Boundary + solid fix-up (reference include/LBMSolver.h:147-265 folded into the pull): ring
// cells are re-pulled from src, get wall / Zou-He treatment in the reference's serial order, are
// checked for stability, collided and stored; solid cells are reset to w.
cudaError_t launch_fixup(bool pull, const StepArgs& a, const BcArgs& b, const int2* ring, int n_ring,
                         const int2* solids, int n_solid, cudaStream_t s);

// Both slab-edge columns in one launch (multi-slab jobs): pull, boundary rules, collide; solids = w.
cudaError_t launch_edge(bool pull, const StepArgs& a, const BcArgs& b, const unsigned char* mask, cudaStream_t s);

// The halo exchange fused into the boundary fix-up (multi-slab jobs with CUDA IPC peer access): every
// edge cell also stores the three populations that cross its slab face straight into the
// neighbouring GPU's ghost column over NVLink, and the launch synchronises with the neighbours
// through step counters in peer memory -- no NCCL call, no second stream, no event in the step.
struct P2pArgs {
    double* peer_dst_west;  // the west neighbour's destination buffer (its E ghost column receives 3,6,7) or nullptr
    double* peer_dst_east;  // the east neighbour's destination buffer (its W ghost column receives 1,5,8) or nullptr
    int* my_flags;          // [0] last step whose halo arrived from the west, [1] from the east (written by the peers)
    int* west_flag;         // -> the west neighbour's my_flags[1]
    int* east_flag;         // -> the east neighbour's my_flags[0]
    unsigned int* blocks_done;  // last-block detection
    int seq;                    // exchange number of this launch (the same on every rank)
    int* status;                // this slab's status word: non-zero = a halo wait timed out / the host gave up
    unsigned long long timeout_ns;  // bound of one wait (0: unbounded)
};
// Interior columns AND the fused edge + halo work in one launch (vectorised variant, even ny): the
// peer stores and the hand-shake run under the interior kernel.
bool bulk_p2p_supported(int variant, const StepArgs& a);
cudaError_t launch_bulk_p2p(bool pull, const StepArgs& a, const BcArgs& b, const unsigned char* mask, const P2pArgs& x,
                            cudaStream_t s);
// Fallback (odd ny / scalar variant): ring / solid fix-up of the interior columns plus both slab-edge
// columns with the fused exchange, after a bulk launch over every column.
cudaError_t launch_fixup_p2p(bool pull, const StepArgs& a, const BcArgs& b, const int2* ring, int n_ring, const int2* solids,
                             int n_solid, const unsigned char* mask, const P2pArgs& x, cudaStream_t s);
// Blocks the stream until the halos of exchange `seq` have arrived from both neighbours.
cudaError_t launch_wait_halo(const P2pArgs& x, cudaStream_t s);

// Momentum-exchange reduction (reference include/LBMIO.h:114-162) over a precomputed link list.
// tree = 0: terms added one by one in the reference's (y, x, i) order (its bits, forces.csv byte for byte);
// tree = 1: fixed parallel reduction tree (deterministic, a few us, equal to rounding).
cudaError_t launch_forces(const double* f_next, const Link* links, int n_links, double* out_fx_fy, cudaStream_t s,
                          int tree = 0);

// Ghost wrap for the periodic extensions.
cudaError_t launch_wrap(double* f, const Layout& L, int wrap_x, int wrap_y, cudaStream_t s);

// Initial state (reference include/LBMGrid.h:185-246) into both buffers.
cudaError_t launch_init(double* f0, double* f1, const Layout& L, const unsigned char* mask, const BcArgs& b,
                        int west_zero, int east_zero, int shear_wave, double u0, cudaStream_t s);

struct ObserveArgs {
    const double* cur;   // newest buffer
    const double* prev;  // the other buffer
    Layout L;
    const unsigned char* mask;  // padded, Layout indexing
    BcArgs bc;
    int cur_is_next;   // cur holds post-collision f_next (>= 1 iteration done) vs. an f_current
    int prev_is_next;  // same for prev
    int fresh;         // straight after initialise(): macroscopic fields are the exact constants
    int shear_wave;
    double u0;
};
// rho / ux / uy exactly as the reference's arrays hold them after the last iteration
// (include/LBMSolver.h:112-114 with the overrides of :203-205, :232-234, :260-261).
// Output interior row-major [y*lnx + x].
cudaError_t launch_macros(const ObserveArgs& o, double* rho, double* ux, double* uy, cudaStream_t s);
// max(ux^2+uy^2) -> *out_bits (as an order-preserving uint64 of a non-negative double).
cudaError_t launch_maxvel(const double* ux, const double* uy, long long n, unsigned long long* out_bits,
                          cudaStream_t s);
// f_current / f_next in the reference's padded AoS order.
// ... of the padded rows [row0, row0 + rows), `aos` pointing at the first of them (row0 a multiple of 32)
cudaError_t launch_export_f(const ObserveArgs& o, int which, double* aos, int row0, int rows, cudaStream_t s);
cudaError_t launch_import_f(const double* aos, double* f, const Layout& L, int row0, int rows, cudaStream_t s);
// div_pair (lbm_cell.cuh) against the IEEE division on n pseudo-random operand triples; *mismatches += differing quotients.
cudaError_t launch_selftest_div(unsigned long long seed, long long n, unsigned long long* mismatches, cudaStream_t s);
// f0[off[k]] = f1[off[k]] = val[k]: caller-written f_next values of solid cells / ghost rows (lbm_upload_f_next).
cudaError_t launch_scatter(const long long* off, const double* val, int n, double* f0, double* f1, cudaStream_t s);
// Ghost ring of one buffer back to the convention (W/E domain-edge columns 0, everything else e).
cudaError_t launch_reset_ghosts(double* f, const Layout& L, const BcArgs& b, int west_zero, int east_zero,
                                cudaStream_t s);

// ---- in-place AA variant (lbm_aa.cu) ------------------------------------------------------------
struct AaArgs {
    double* f;  // the single population buffer
    Layout L;
    double tau_inv, Fx, Fy;
    int forced;
    int* first_bad;
    int bad_iter;
    int first;      // E-step on a fresh f_current (after initialise / upload): no boundary rule, no check
    int skip_rows;  // walls: rows 0 and ny-1 are ring cells, left to the fix-up kernel
    int x_begin, x_end;  // bulk columns (inlet / outlet columns are ring cells too)
    int variant;
    const int4* cols;  // as StepArgs::cols
    int col_lo, col_hi;
};
// A slot [x][i] of a fluid cell whose upstream neighbour x - c_i is a solid or a non-periodic
// ghost: nobody pushes into it, the O-step fix-up writes the constant (0: w_i, 1: 0.0, 2: e_i).
struct AaFill {
    long long off;
    int kind, i;
};
struct AaObserve {
    const double* f;
    Layout L;
    const unsigned char* mask;
    BcArgs bc;
    int phase;        // 0: natural layout (after initialise / upload / an O-step), 1: reversed (after an E-step)
    int cur_is_next;  // at least one iteration done since initialise / upload
    int fresh;
    const double* ring_out;  // post-collision populations of the ring cells after an O-step
    int periodic_x, periodic_y;
    int open_w, open_e;  // the edge is periodic or a slab interface: populations cross it
    int west_zero, east_zero;
    int shear_wave;
    double u0;
    double tau_inv, Fx, Fy;
};
cudaError_t launch_aa_bulk(bool odd, const AaArgs& a, cudaStream_t s);
cudaError_t launch_aa_fix_even(const AaArgs& a, const BcArgs& b, const int2* ring, int n_ring, const int2* solids,
                               int n_solid, cudaStream_t s);
cudaError_t launch_aa_fix_odd(const AaArgs& a, const BcArgs& b, const int2* ring, int n_ring, const AaFill* fills,
                              int n_fill, double* ring_out, int open_x, int open_y, cudaStream_t s);
cudaError_t launch_aa_unwrap(double* f, const Layout& L, int do_x, int do_y, cudaStream_t s);
// x-slabs: the forward wrap (after an E-step) / reverse wrap (after an O-step) into the neighbouring GPUs' memory,
// then the hand-shake (publishes px.seq)
cudaError_t launch_aa_halo(const double* f, const Layout& L, int reverse, int rows_open, const P2pArgs& px, cudaStream_t s);
cudaError_t launch_aa_ghosts(double* f, const Layout& L, const BcArgs& b, int west_zero, int east_zero, cudaStream_t s);
cudaError_t launch_aa_macros(const AaObserve& o, double* rho, double* ux, double* uy, cudaStream_t s);
cudaError_t launch_aa_export(const AaObserve& o, int which, double* aos, int row0, int rows, cudaStream_t s);
cudaError_t launch_aa_check(const AaObserve& o, int* first_bad, int bad_iter, cudaStream_t s);

// The solid run of column x, or an empty one (no load at all outside [col_lo, col_hi)).
template <class Args>
__device__ __forceinline__ int4 column_run(const Args& a, int x) {
#if defined(__CUDA_ARCH__)
    if (a.cols && x >= a.col_lo && x < a.col_hi) return __ldg(a.cols + x);
#endif
    return make_int4(0, 0, 0, 0);
}

}  // namespace lbm
