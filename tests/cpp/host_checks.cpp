// tests/cpp/host_checks.cpp -- CPU-only checks of the host side of the drop-in headers (no GPU
// call is made): the exact "%.8f" formatter against snprintf, and the VTK / forces.csv writers.
//   host_checks fixed8 <count> <seed>        -> prints "ok <count>" or the first mismatch
//   host_checks vtk <nx> <ny> <seed> <t>     -> writes vtk_output/lbm_<t>.vtk from a seeded field
//                                               and dumps the field to field.bin (rho, ux, uy)
//   host_checks forces                       -> writes forces.csv rows from stdin "t fx fy"
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <sys/stat.h>

#include "LBMIO.h"

static int check_fixed8(long count, unsigned seed) {
    std::mt19937_64 rng(seed);
    char a[400], b[400];
    auto one = [&](double v) -> bool {
        const int n = LBM::textio::fixed8(v, a);
        a[n] = 0;
        std::snprintf(b, sizeof(b), "%.8f", v);
        if (std::strcmp(a, b) != 0) {
            std::printf("mismatch for %.17g (bits %016" PRIx64 "): got %s want %s\n", v, *(uint64_t*)&v, a, b);
            return false;
        }
        return true;
    };
    const double edge[] = {0.0, -0.0, 1e-9, -1e-9, 5e-9, -5e-9, 4.9999999999e-9, 0.5, 1.0, -1.0, 0.125, 0.000000005,
                           0.000000015, 0.000000025, 123456.123456785, 999999.999999995, 999999.99999999, 1e6, -1e6, 1e7,
                           3.4e15, 1e300, 999999999999999.9, 99999999.999999996, 0.999999995, 0.99999999499999, 1e14 + 0.5, 2.5e-8, 3.5e-8, 0.1, 0.2, 0.3, 358.42293907, 0.01333, 1.0 / 3.0, 2.0 / 3.0};
    for (double v : edge)
        if (!one(v)) return 1;
    std::uniform_real_distribution<double> u(-1.0, 1.0);
    for (long k = 0; k < count; ++k) {
        double v;
        switch (k % 6) {
            case 0: v = u(rng); break;                          // velocities
            case 1: v = 1.0 + 0.1 * u(rng); break;              // densities
            case 2: v = 400.0 * u(rng); break;                  // force coefficients
            case 3: v = std::ldexp(u(rng), (int)(rng() % 60) - 40); break;
            case 4: v = ((double)(int64_t)(rng() % 2000000001) - 1e9) * 0.5e-8; break;  // exact ties .5e-8 grid
            default: v = ((double)(int64_t)(rng() % 200000001) - 1e8) * 1e-8; break;    // 8-decimal values
        }
        if (!one(v)) return 1;
    }
    std::printf("ok %ld\n", count);
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 4 && !std::strcmp(argv[1], "fixed8")) return check_fixed8(std::atol(argv[2]), (unsigned)std::atoi(argv[3]));
    if (argc >= 6 && (!std::strcmp(argv[1], "vtk") || !std::strcmp(argv[1], "vtkbin"))) {
        const bool binary = !std::strcmp(argv[1], "vtkbin");
        const int nx = std::atoi(argv[2]), ny = std::atoi(argv[3]), t = std::atoi(argv[5]);
        std::mt19937_64 rng((unsigned)std::atoi(argv[4]));
        std::uniform_real_distribution<double> u(-0.2, 0.2);
        std::vector<double> rho((size_t)nx * ny), ux(rho.size()), uy(rho.size());
        for (size_t k = 0; k < rho.size(); ++k) {
            rho[k] = 1.0 + 0.05 * u(rng);
            ux[k] = u(rng);
            uy[k] = (k % 7 == 0) ? -0.0 : u(rng) * 1e-9;
        }
        mkdir("vtk_output", 0755);
        LBM::SimulationParams p;
        p.nx = nx;
        p.ny = ny;
        if (binary)
            LBM::IOManager::write_vtk_arrays_binary(ux.data(), uy.data(), rho.data(), nx, ny, t);
        else
            LBM::IOManager::write_vtk_timestep(ux, uy, rho, p, t);
        std::FILE* f = std::fopen("field.bin", "wb");
        std::fwrite(rho.data(), 8, rho.size(), f);
        std::fwrite(ux.data(), 8, ux.size(), f);
        std::fwrite(uy.data(), 8, uy.size(), f);
        std::fclose(f);
        return 0;
    }
    if (argc >= 2 && !std::strcmp(argv[1], "forces")) {
        LBM::SimulationParams p;
        LBM::IOManager io;  // opens forces.csv in cwd
        int t;
        double fx, fy;
        while (std::scanf("%d %lf %lf", &t, &fx, &fy) == 3) io.write_force_row(t, fx, fy, p);
        return 0;
    }
    std::fprintf(stderr, "usage: host_checks fixed8|vtk|forces ...\n");
    return 2;
}
