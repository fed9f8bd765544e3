// oracle/prototypes/tb2_host.cpp -- CPU blueprint of TEMPORAL BLOCKING for round 2.  TEST INFRASTRUCTURE.
//
// Two lattice updates per pass over the population buffer: for an
// output tile T of TX x TY cells,
//   stage 1  computes the intermediate post-collision state G = Fn_t on T grown by one cell in every
//            direction (ghost cells: their constants; solid cells: w; everything else: pull from the
//            source buffer Fn_{t-1}, boundary rule, collide) into a tile-local array -- shared memory on
//            the GPU;
//   stage 2  computes Fn_{t+1} on T by pulling from G, boundary rule, collide, and is the only thing
//            that goes back to the buffer.
// Per-cell arithmetic is the DEVICE code itself (csrc/lbm_cell.cuh, host-compilable, no FP
// contraction), so what this file pins is the tiling / halo / boundary bookkeeping: the result must
// equal two single steps of the oracle bit for bit (tests/test_tb2_blueprint.py), for any tile shape,
// including tiles that touch walls, inlet, outlet, the corner quirks and the cylinder.
//
// Arrays are the reference's padded AoS layout [(gy*(nx+2)+gx)*9+i] (include/LBMGrid.h:105-107).
#include <algorithm>
#include <cstdint>
#include <vector>

#include "../../highperformancecomputing-latticeboltzmannmethod_b200/csrc/lbm_cell.cuh"

using namespace lbm;

namespace {

struct Ctx {
    int nx, ny;
    double tau_inv, u_in;
    const unsigned char* solid;  // interior [y*nx+x]
    double w[Q];
};

inline size_t at(const Ctx& c, int gx, int gy) { return ((size_t)gy * (c.nx + 2) + gx) * Q; }

// the reference's rules in its serial order (include/LBMSolver.h:153-236): bottom, top, inlet, outlet
inline void rules(const Ctx& c, int x, int y, double f[Q]) {
    if (y == 0) wall_bottom(f);
    if (y == c.ny - 1) wall_top(f);
    if (x == 0) zou_he_inlet(f, c.u_in);
    if (x == c.nx - 1) zou_he_outlet(f);
}

// One update of interior cell (x, y): `fetch(gx, gy, i)` returns population i of padded cell (gx, gy)
// of the previous post-collision state.  Returns true if a pulled value fails the stability test.
template <class Fetch>
inline bool update_cell(const Ctx& c, int x, int y, Fetch fetch, double out[Q]) {
    if (c.solid[(size_t)y * c.nx + x]) {
        for (int i = 0; i < Q; ++i) out[i] = c.w[i];
        return false;  // (the reference's check sees the reversed pulls here; the GPU bulk kernel does too)
    }
    double f[Q];
    for (int i = 0; i < Q; ++i) f[i] = fetch(x + 1 - cxi(i), y + 1 - cyi(i), i);
    rules(c, x, y, f);
    bool bad = false;
    for (int i = 0; i < Q; ++i) bad |= unstable_value(f[i]);
    const Moments m = moments(f);
    bgk(f, m, c.tau_inv, out);
    return bad;
}

}  // namespace

extern "C" {

// src, dst: padded AoS post-collision states Fn_{t-1} -> Fn_{t+1}.  Returns the number of cells whose
// stage-1 value was computed (interior cells of the grown tiles: the redundancy of the scheme), and
// sets *bad1 / *bad2 when a pulled value of the first / second update is unstable.
long long tb2_pass(const double* src, double* dst, const unsigned char* solid, int nx, int ny, double tau, double u_in, int TX,
                   int TY, int* bad1, int* bad2) {
    Ctx c{nx, ny, 1.0 / tau, u_in, solid, {}};
    equilibrium_init(1.0, 0.0, 0.0, c.w);
    *bad1 = *bad2 = 0;
    // ghost ring: time-invariant in a single-slab job
    for (int gy = 0; gy < ny + 2; ++gy)
        for (int gx = 0; gx < nx + 2; ++gx)
            if (gx == 0 || gx == nx + 1 || gy == 0 || gy == ny + 1)
                for (int i = 0; i < Q; ++i) dst[at(c, gx, gy) + i] = src[at(c, gx, gy) + i];
    long long stage1_cells = 0;
    std::vector<double> G;
    for (int y0 = 0; y0 < ny; y0 += TY)
        for (int x0 = 0; x0 < nx; x0 += TX) {
            const int x1 = std::min(nx, x0 + TX), y1 = std::min(ny, y0 + TY);
            // grown tile in PADDED coordinates: [x0, x1+1] x [y0, y1+1]  (= interior x0-1 .. x1, y0-1 .. y1)
            const int gw = x1 - x0 + 2, gh = y1 - y0 + 2;
            G.assign((size_t)gw * gh * Q, 0.0);
            auto g_at = [&](int gx, int gy) -> double* { return &G[((size_t)(gy - y0) * gw + (gx - x0)) * Q]; };
            for (int gy = y0; gy <= y1 + 1; ++gy)
                for (int gx = x0; gx <= x1 + 1; ++gx) {
                    double* out = g_at(gx, gy);
                    if (gx == 0 || gx == nx + 1 || gy == 0 || gy == ny + 1) {
                        for (int i = 0; i < Q; ++i) out[i] = src[at(c, gx, gy) + i];
                        continue;
                    }
                    ++stage1_cells;
                    const bool owned = gx - 1 >= x0 && gx - 1 < x1 && gy - 1 >= y0 && gy - 1 < y1;
                    const bool bad = update_cell(c, gx - 1, gy - 1, [&](int px, int py, int i) { return src[at(c, px, py) + i]; }, out);
                    if (bad && owned) *bad1 = 1;  // halo cells are checked by the tile that owns them
                }
            for (int y = y0; y < y1; ++y)
                for (int x = x0; x < x1; ++x)
                    if (update_cell(c, x, y, [&](int px, int py, int i) { return g_at(px, py)[i]; }, &dst[at(c, x + 1, y + 1)])) *bad2 = 1;
        }
    return stage1_cells;
}

}  // extern "C"
