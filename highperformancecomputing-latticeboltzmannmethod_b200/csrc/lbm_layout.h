// lbm_layout.h -- how one x-slab of populations lives in HBM.
//
// SoA, fp64, one plane per population i = 0..8.  Inside a plane the slab is stored COLUMN-major:
// y is the fastest-varying index, x the slowest, so that
//   * a halo column (fixed x, all y) of one population is one contiguous run of ny doubles and
//     goes to the neighbouring GPU without a pack kernel, and
//   * a shift in x (6 of the 9 pulls) is a whole-column offset that keeps 128-bit alignment;
//     only the +-1 shifts in y are misaligned by one element.
// The reference's one-cell ghost ring is kept (include/LBMGrid.h:63-64): columns gx = 0 and
// lnx+1, rows y = -1 and ny.  A column is padded in front so that interior row y = 0 starts on a
// 128-byte boundary, and its pitch is a multiple of 16 doubles (128 B).  XO further ghost columns on
// either side (gx = -1, -2 and lnx+2, lnx+3) hold the wider halo that a temporally blocked pass of
// depth T needs from the neighbouring slab (lbm_tb.cuh): T ghost columns per face.
//
//   element (i, gx, y)  ->  i*plane + (gx+XO)*PY + YO + y,      gx in [-XO, lnx+2+XO), y in [-1, ny]
//
// The reference's AoS order (include/LBMGrid.h:105-107) only exists at the API boundary
// (lbm_download_f / lbm_upload_f), produced by a transpose kernel.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define LBM_LAYOUT_HD __host__ __device__ __forceinline__
#else
#define LBM_LAYOUT_HD inline
#endif

namespace lbm {

struct Layout {
    int lnx;      // interior columns in this slab
    int ny;       // interior rows
    int gnx;      // global nx
    int x_start;  // global x of interior column 0
    int PY;       // column pitch, doubles
    long long plane;  // doubles per population plane = (lnx+2+2*XO)*PY

    static constexpr int YO = 16;  // interior row 0 sits at this offset inside a column
    static constexpr int XO = 2;   // extra ghost columns per side beyond the reference's one

    LBM_LAYOUT_HD long long at(int gx, int y) const { return (long long)(gx + XO) * PY + YO + y; }
    LBM_LAYOUT_HD long long cells_padded() const { return (long long)(lnx + 2 + 2 * XO) * PY; }

    static Layout make(int lnx, int ny, int gnx, int x_start) {
        Layout L;
        L.lnx = lnx;
        L.ny = ny;
        L.gnx = gnx;
        L.x_start = x_start;
        L.PY = ((YO + ny + 1 + 15) / 16) * 16;
        L.plane = (long long)(lnx + 2 + 2 * XO) * L.PY;
        return L;
    }
};

// One momentum-exchange link (reference include/LBMIO.h:133-159): the post-collision population
// `off` (absolute element offset into the SoA buffer) of a fluid cell points into a solid cell;
// it contributes 2*c_i*f to the force.
struct Link {
    long long off;
    int cx2, cy2;  // 2*c_ix, 2*c_iy
};

}  // namespace lbm
