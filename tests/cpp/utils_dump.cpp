// tests/cpp/utils_dump.cpp -- equilibrium_scalar / equilibrium_simd / is_stable of LBMUtils.h on a few
// inputs, printed with 17 digits.  Compiled against this repo's header and against the reference's
// (AVX2 intrinsics, -mavx2 -mfma -ffp-contract=off): identical text expected.
#include <cmath>
#include <cstdio>
#include <limits>

#include "LBMUtils.h"

int main() {
    const double cases[][3] = {{1.0, 0.01333, 0.0}, {1.0, 0.0, 0.0}, {0.97, 0.1333, -0.02}, {1.08, -0.05, 0.07}, {1.0, 0.2, 0.2}};
    for (auto& c : cases) {
        double f[8];
        LBM::equilibrium_simd(c[0], c[1], c[2], f);
        std::printf("%.17g", LBM::equilibrium_scalar(c[0], c[1], c[2]));
        for (double v : f) std::printf(" %.17g", v);
        std::printf("\n");
    }
    const double probes[] = {0.0, 1e5, -1e5, 1.0000001e5, -1.0000001e5, std::numeric_limits<double>::infinity()};
    for (double v : probes) std::printf("stable(%g) %d\n", v, (int)LBM::is_stable(v));
    return 0;
}
