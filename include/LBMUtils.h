// LBMUtils.h -- host-side helpers kept for source compatibility with code that includes the
// reference's include/LBMUtils.h.  The device kernels do not use this file (their arithmetic
// lives in csrc/lbm_cell.cuh); these are plain scalar C++ for callers that want an equilibrium
// value or the stability predicate on the host.
#pragma once

#include <cmath>

#include "LBMConfig.h"

namespace LBM {

// Rest-population equilibrium, reference include/LBMUtils.h:9-12.
inline double equilibrium_scalar(double rho, double ux, double uy) {
    return W_REST * rho * (1.0 - 1.5 * (ux * ux + uy * uy));
}

// The eight moving-population equilibria, f_eq[k] for direction i = k + 1, with the output convention
// and the evaluation order of the reference's AVX2 routine (include/LBMUtils.h:22-65):
// (w rho) * (((1 + 3 cu) - 1.5 u^2) + 4.5 cu^2).  The rest population is equilibrium_scalar().
inline void equilibrium_simd(double rho, double ux, double uy, double* f_eq) {
    const double usq15 = 1.5 * (ux * ux + uy * uy);
    for (int i = 1; i < Q; ++i) {
        const double cu = VELOCITIES[i][0] * ux + VELOCITIES[i][1] * uy;
        f_eq[i - 1] = (WEIGHTS[i] * rho) * (((1.0 + 3.0 * cu) - usq15) + 4.5 * (cu * cu));
    }
}

// The scalar predicate of include/LBMUtils.h:129-131.  (Grid::check_stability's vector path,
// include/LBMGrid.h:297-307, which the device kernels follow, flags v > 1e5 or v < -1e5 instead;
// the two differ only for |v| == 1e5 exactly.)
inline bool is_stable(const double value) { return std::isfinite(value) && std::abs(value) < 1e5; }

}  // namespace LBM
