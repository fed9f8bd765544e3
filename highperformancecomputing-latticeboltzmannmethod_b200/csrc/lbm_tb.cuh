// lbm_tb.cuh -- TEMPORAL BLOCKING: T lattice updates per pass over the population buffers.
//
// The A-B kernels of lbm_kernels.cu run at the HBM roofline of 144 B per update (9 fp64 loads + 9 fp64
// stores); the only way past it is to do more than one update per trip through HBM.  One thread block
// owns a strip of rows and marches along x through a chunk of columns.  At march step s
//   stage 1   updates column s           from the source buffer in HBM        (reference iteration t)
//   stage k   updates column s-(k-1)     from stage k-1's ring in SHARED MEMORY (iteration t+k-1)
//   stage T   stores column s-(T-1)      to the destination buffer in HBM
// so each population is read once and written once per T updates: 144/T B per update.  A stage's ring
// holds its last four columns (a pull needs the three columns c-1, c, c+1; four slots make one
// __syncthreads per stage and step enough).  Every stage is the complete reference iteration for its
// cells -- pull (include/LBMSolver.h:128-145), boundary rules in the reference's serial order
// (:147-236), stability check (include/LBMGrid.h:285-317), BGK collision (:84-126) -- with the quirks
// of SURVEY.md F3/F4 reproduced as VALUES of the intermediate state: solid cells are w, ghost rows are
// eq(1,u_in,0), ghost columns at the physical inlet/outlet are 0, ghost columns at a slab interface
// are computed from a halo of width T that the neighbouring GPU stored into this slab's memory.
// The per-cell arithmetic is lbm_cell.cuh, so a pass of depth T is bit-identical to T single steps
// (tests/test_tb_emulation.py runs THIS code on the host, thread for thread; tests/test_gpu_tb.py on
// the GPU).
//
// Redundant work: a strip of B threads yields B-4 rows (each stage needs one more row on either side
// than the next), a chunk of XC columns costs XC + 2(T-1) stage-1 columns: ~3 % at B = 256, XC = 128.
//
// Depth 1 is the plain fused step (pull + rules + collide in ONE launch, no fix-up kernel); it is what
// multi-slab jobs use for the single steps between passes because it stores the same wide halo.
#pragma once

#include "lbm_cell.cuh"
#include "lbm_kernels.cuh"
#include "lbm_layout.h"

namespace lbm {

enum TbEdge {
    TB_EDGE_CONST = 0,  // physical inlet / outlet: the ghost column holds 0.0
    TB_EDGE_HALO = 1,   // slab interface: ghost columns hold the neighbour's populations
    TB_EDGE_WRAP = 2    // periodic in x inside one slab: column indices wrap
};

constexpr int TB_SLOTS = 4;       // ring slots per stage
constexpr int TB_MAX_DEPTH = 3;   // == Layout::XO + 1 ghost columns are addressable

struct TbArgs {
    const double* src;
    double* dst;
    Layout L;
    double tau_inv, Fx, Fy;
    int* first_bad;
    int bad_iter;  // reference timestep whose check_stability stage 1's pulled values belong to (stage k: + k-1)
    BcArgs bc;
    const unsigned char* mask;  // padded, Layout indexing: 0 fluid, 1 solid, 2 solid with eight solid neighbours
    int mask_lo, mask_hi;       // slab columns [lo, hi) (ghost columns count) that may hold solid cells
    int west, east;             // TbEdge
    int periodic_y;
    int pull;   // depth 1 only: 0 = the first iteration after initialise / upload (collide f_current as it is)
    int write;  // 0: check / observe only, no population is stored
    int x_begin, x_end, xc;  // interior chunks tile [x_begin, x_end) in pieces of xc columns ...
    int edge_cols;           // ... after chunk 0 = [0, edge_cols) and chunk 1 = [lnx-edge_cols, lnx) when > 0
    int halo_w;              // ghost columns a slab interface keeps up to date (the deepest pass of the job)
    // rho / ux / uy exactly as the LAST stage's collision stores them (include/LBMSolver.h:112-114), in
    // the slab's native order [x*ny + y]; nullptr: not wanted
    double *m_rho, *m_ux, *m_uy;
    P2pArgs px;  // multi-slab: peer buffers and the hand-shake words (lbm_kernels.cuh)
};

// ---- host / device glue (the host side exists for the thread-for-thread emulation only) ----------
#if defined(__CUDA_ARCH__)
#define TB_SYNC() __syncthreads()
#define TB_LD(p, coherent) ((coherent) ? __ldcg(p) : __ldg(p))
#define TB_FLAG(ptr, v) atomicMin((ptr), (v))
#else
void tb_host_sync();
void tb_host_flag(int* p, int v);
#define TB_SYNC() tb_host_sync()
#define TB_LD(p, coherent) (*(p))
#define TB_FLAG(ptr, v) tb_host_flag((ptr), (v))
#endif

LBM_HD int tb_wrap(int v, int n) {
    v %= n;
    return v < 0 ? v + n : v;
}

template <int T, int B>
struct TbShape {
    static constexpr int ROFF = (T == 1) ? 0 : 2;  // thread j works on row strip*H - ROFF + j
    static constexpr int H = B - 2 * ROFF;         // rows a strip stores (a multiple of four: whole 32-byte sectors)
    static constexpr int RING_DOUBLES = (T - 1) * TB_SLOTS * Q * B;
    static_assert(T >= 1 && T <= TB_MAX_DEPTH, "depth");
    static_assert(ROFF >= T - 1, "row overlap");
};

// The chunk of columns [x0, x1) block `chunk` stores; edge = it reads ghost columns / feeds a neighbour.
LBM_HD void tb_chunk(const TbArgs& a, int chunk, int& x0, int& x1, bool& edge) {
    const int lnx = a.L.lnx;
    edge = false;
    if (a.edge_cols > 0) {
        if (chunk == 0) { x0 = 0; x1 = a.edge_cols; edge = true; return; }
        if (chunk == 1) { x0 = lnx - a.edge_cols; x1 = lnx; edge = true; return; }
        chunk -= 2;
    }
    x0 = a.x_begin + chunk * a.xc;
    x1 = x0 + a.xc < a.x_end ? x0 + a.xc : a.x_end;
}

// One thread of one block: `tid` in [0, B), rows of strip `strip`, columns of chunk `chunk`.
// `ring` is the block's shared memory (TbShape::RING_DOUBLES doubles).
template <int T, int B, bool FORCED>
LBM_HD void tb_thread(const TbArgs& a, double* ring, int tid, int strip, int chunk) {
    using S = TbShape<T, B>;
    const Layout& L = a.L;
    const int lnx = L.lnx, ny = L.ny;
    int x0, x1;
    bool edge;
    tb_chunk(a, chunk, x0, x1, edge);
    if (x0 >= x1) return;  // (uniform over the block)

    // ---- this thread's row -------------------------------------------------------------------
    const int ys = strip * S::H;
    const int y = ys - S::ROFF + tid;
    // row kinds: 0 dead, 1 a cell to compute (row yr), 2 ghost row holding eq(1,u_in,0)
    int row_kind, yr = y;
    if (a.periodic_y) {
        row_kind = (y >= -(T - 1) && y < ny + (T - 1)) ? 1 : 0;
        yr = tb_wrap(y, ny);
    } else {
        row_kind = (y >= 0 && y < ny) ? 1 : ((y == -1 || y == ny) ? 2 : 0);
    }
    // rows the pulls of row yr read: yr-1, yr, yr+1 (ghost rows -1 and ny exist in memory)
    int o_m = Layout::YO + yr - 1, o_0 = Layout::YO + yr, o_p = Layout::YO + yr + 1;
    if (a.periodic_y) {
        o_m = Layout::YO + tb_wrap(yr - 1, ny);
        o_p = Layout::YO + tb_wrap(yr + 1, ny);
    }
    const bool wall_b = a.bc.walls && yr == 0, wall_t = a.bc.walls && yr == ny - 1;
    const bool out_row = (tid >= S::ROFF && tid < S::ROFF + S::H && y >= 0 && y < ny);  // rows the last stage stores
    bool bad[T];
#pragma unroll
    for (int k = 0; k < T; ++k) bad[k] = false;

    const int s_first = x0 - (T - 1), s_last = x1 + (T - 2);
    for (int s = s_first; s <= s_last; ++s) {
#pragma unroll
        for (int k = 1; k <= T; ++k) {
            const int c = s - (k - 1);  // the column stage k works on
            const int grow = T - k;     // how far stage k reaches beyond the rows / columns the block stores
            // ---- column kind (uniform over the block) --------------------------------------------
            // 0 dead, 1 compute cell column xr, 2 ghost column at a physical edge (0.0; corners: e)
            int col_kind = 0, xr = c;
            if (c >= x0 - grow && c < x1 + grow) {
                if (c >= 0 && c < lnx) col_kind = 1;
                else {
                    const int mode = c < 0 ? a.west : a.east;
                    if (mode == TB_EDGE_WRAP) { col_kind = 1; xr = tb_wrap(c, lnx); }
                    else if (mode == TB_EDGE_HALO) col_kind = 1;  // (inside the halo by construction of the chunks)
                    else col_kind = (c == -1 || c == lnx) ? 2 : 0;
                }
            }
            const bool row_on = row_kind != 0 && tid >= S::ROFF - grow && tid < S::ROFF + S::H + grow &&
                                (k < T || out_row);
            double f[Q];
            bool have = false, solid = false;
            if (col_kind != 0 && row_on) {
                have = true;
                if (row_kind == 2) {
#pragma unroll
                    for (int i = 0; i < Q; ++i) f[i] = a.bc.e[i];
                } else if (col_kind == 2) {
#pragma unroll
                    for (int i = 0; i < Q; ++i) f[i] = 0.0;
                } else {
                    int m = 0;
                    if (xr + 1 >= a.mask_lo && xr + 1 < a.mask_hi) m = a.mask[L.at(xr + 1, yr)];
                    solid = m != 0;
                    if (m == 2) {
                        // every neighbour is solid: the pulls are w, nothing to check, nothing to compute
#pragma unroll
                        for (int i = 0; i < Q; ++i) f[i] = a.bc.w[i];
                    } else {
                        // ---- pull ----------------------------------------------------------------
                        if (k == 1) {
                            int cw = xr - 1, ce = xr + 1;
                            if (a.west == TB_EDGE_WRAP) { cw = tb_wrap(cw, lnx); ce = tb_wrap(ce, lnx); }
                            const bool pulling = (T > 1) || a.pull;
                            const long long bw = L.at(pulling ? cw + 1 : xr + 1, 0) - Layout::YO;
                            const long long b0 = L.at(xr + 1, 0) - Layout::YO;
                            const long long be = L.at(pulling ? ce + 1 : xr + 1, 0) - Layout::YO;
                            const int r_m = pulling ? o_m : o_0, r_p = pulling ? o_p : o_0;
                            const double* p = a.src;
                            const long long pl = L.plane;
                            // population i comes from column xr - c_ix, row yr - c_iy
                            f[0] = TB_LD(p + 0 * pl + b0 + o_0, edge);
                            f[1] = TB_LD(p + 1 * pl + bw + o_0, edge);
                            f[2] = TB_LD(p + 2 * pl + b0 + r_m, edge);
                            f[3] = TB_LD(p + 3 * pl + be + o_0, edge);
                            f[4] = TB_LD(p + 4 * pl + b0 + r_p, edge);
                            f[5] = TB_LD(p + 5 * pl + bw + r_m, edge);
                            f[6] = TB_LD(p + 6 * pl + be + r_m, edge);
                            f[7] = TB_LD(p + 7 * pl + be + r_p, edge);
                            f[8] = TB_LD(p + 8 * pl + bw + r_p, edge);
                        } else {
                            const double* r = ring + (size_t)(k - 2) * TB_SLOTS * Q * B;
                            const int sw = ((c - 1) & (TB_SLOTS - 1)) * Q * B, s0 = (c & (TB_SLOTS - 1)) * Q * B,
                                      se = ((c + 1) & (TB_SLOTS - 1)) * Q * B;
                            f[0] = r[s0 + 0 * B + tid];
                            f[1] = r[sw + 1 * B + tid];
                            f[2] = r[s0 + 2 * B + tid - 1];
                            f[3] = r[se + 3 * B + tid];
                            f[4] = r[s0 + 4 * B + tid + 1];
                            f[5] = r[sw + 5 * B + tid - 1];
                            f[6] = r[se + 6 * B + tid - 1];
                            f[7] = r[se + 7 * B + tid + 1];
                            f[8] = r[sw + 8 * B + tid + 1];
                        }
                        const bool rules = (T > 1) || a.pull;  // the first iteration has no boundary pass before it
                        double rho_bc = 0.0, u_out = 0.0;
                        if (!solid && rules) {
                            // the reference's serial order: bottom, top, inlet, outlet
                            if (wall_b) wall_bottom(f);
                            if (wall_t) wall_top(f);
                            if (a.bc.inlet && xr == 0) rho_bc = zou_he_inlet(f, a.bc.u_in);
                            if (a.bc.outlet && xr == lnx - 1) u_out = zou_he_outlet(f);
                        }
                        (void)rho_bc;
                        (void)u_out;
                        if (rules) {
#pragma unroll
                            for (int i = 0; i < Q; ++i) bad[k - 1] |= unstable_value(f[i]);
                        }
                        if (solid) {
#pragma unroll
                            for (int i = 0; i < Q; ++i) f[i] = a.bc.w[i];
                        } else {
                            const Moments mo = moments(f);
                            if (k == T && a.m_rho) {
                                const long long g = (long long)c * ny + y;
                                a.m_rho[g] = mo.rho;
                                a.m_ux[g] = mo.ux;
                                a.m_uy[g] = mo.uy;
                            }
                            if (FORCED)
                                bgk_forced(f, mo, a.tau_inv, a.Fx, a.Fy, f);
                            else
                                bgk(f, mo, a.tau_inv, f);
                        }
                    }
                    if (k == T && solid && a.m_rho) {  // include/LBMSolver.h:260-261; rho keeps the constructor's 1.0
                        const long long g = (long long)c * ny + y;
                        a.m_rho[g] = 1.0;
                        a.m_ux[g] = 0.0;
                        a.m_uy[g] = 0.0;
                    }
                }
            }
            if (k < T) {
                if (have) {
                    double* w = ring + (size_t)(k - 1) * TB_SLOTS * Q * B + (c & (TB_SLOTS - 1)) * Q * B + tid;
#pragma unroll
                    for (int i = 0; i < Q; ++i) w[i * B] = f[i];
                }
                TB_SYNC();
            } else if (have && a.write) {
                // ---- the last stage: HBM, and the neighbours' ghost columns over NVLink ---------------
                const long long o = L.at(c + 1, y);
                // solid cells keep w for ever and are never stored -- except by the first iteration
                // after an upload, whose destination buffer may hold anything there
                if (!solid || (T == 1 && !a.pull)) {
#pragma unroll
                    for (int i = 0; i < Q; ++i) a.dst[i * L.plane + o] = f[i];
                }
                if (edge) {
                    // Column lnx-1-d goes to the east neighbour's ghost column gx = -d.  What it needs of that column,
                    // for a pass of depth halo_w: the populations moving towards it (1,5,8) always; the ones that stay in
                    // the column (0,2,4) where it recomputes the column itself (d <= halo_w-2); the ones moving away
                    // (3,6,7) where it also recomputes the column beyond (d <= halo_w-3).
                    const int de = lnx - 1 - c, dw = c;
                    if (a.px.peer_dst_east && de < a.halo_w) {
                        double* q = a.px.peer_dst_east + L.at(-de, y);
                        q[1 * L.plane] = f[1];
                        q[5 * L.plane] = f[5];
                        q[8 * L.plane] = f[8];
                        if (de <= a.halo_w - 2) {
                            q[0 * L.plane] = f[0];
                            q[2 * L.plane] = f[2];
                            q[4 * L.plane] = f[4];
                        }
                        if (de <= a.halo_w - 3) {
                            q[3 * L.plane] = f[3];
                            q[6 * L.plane] = f[6];
                            q[7 * L.plane] = f[7];
                        }
                    }
                    if (a.px.peer_dst_west && dw < a.halo_w) {
                        double* q = a.px.peer_dst_west + L.at(lnx + 1 + dw, y);
                        q[3 * L.plane] = f[3];
                        q[6 * L.plane] = f[6];
                        q[7 * L.plane] = f[7];
                        if (dw <= a.halo_w - 2) {
                            q[0 * L.plane] = f[0];
                            q[2 * L.plane] = f[2];
                            q[4 * L.plane] = f[4];
                        }
                        if (dw <= a.halo_w - 3) {
                            q[1 * L.plane] = f[1];
                            q[5 * L.plane] = f[5];
                            q[8 * L.plane] = f[8];
                        }
                    }
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < T; ++k)
        if (bad[k]) TB_FLAG(a.first_bad, a.bad_iter + k);
}

// ---- launch interface (lbm_tb.cu) -----------------------------------------------------------
// depth 1..TB_MAX_DEPTH.  Fills in the chunking (x_begin / x_end / xc / edge_cols) itself.
cudaError_t launch_tb(int depth, TbArgs a, bool p2p, cudaStream_t s);
// Finish emitted macros: native [x*ny+y] -> the reference's interior row-major [y*lnx+x], with the inlet /
// outlet overrides of the boundary pass that follows the last collision (include/LBMSolver.h:203-205,
// 232-234), which are functions of the newest buffer alone.
cudaError_t launch_macros_finish(const ObserveArgs& o, const double* m_rho, const double* m_ux, const double* m_uy,
                                 double* rho, double* ux, double* uy, cudaStream_t s);
int tb_rows_per_block(int depth);
size_t tb_shared_bytes(int depth);

}  // namespace lbm
