// tests/cpp/grid_api_dump.cpp -- exercises the element accessors of LBM::Grid the way external code
// written against the reference would (ghost-inclusive population indices, interior indices for the
// macroscopic fields, include/LBMGrid.h:115-129 of the reference) and dumps everything for a
// comparison with the CPU oracle.  Optionally pokes a few populations through the writable
// f_current() accessor before running (a caller-defined initial condition).
//   grid_api_dump <nx> <ny> <steps> <poke 0|1> <outdir>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "LBMConfig.h"
#include "LBMIO.h"
#include "LBMSolver.h"

static void write_doubles(const std::string& path, const std::vector<double>& v) {
    std::FILE* f = std::fopen(path.c_str(), "wb");
    std::fwrite(v.data(), sizeof(double), v.size(), f);
    std::fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    LBM::SimulationParams p;
    p.nx = std::atoi(argv[1]);
    p.ny = std::atoi(argv[2]);
    p.num_timesteps = std::atoi(argv[3]);
    p.output_frequency = 5;
    p.cylinder_radius = 0.15;
    const bool poke = std::atoi(argv[4]) != 0;
    const std::string out = argv[5];
    try {
        LBM::Solver solver(p, false);
        LBM::IOManager io;
        solver.initialise();
        LBM::Grid& g = const_cast<LBM::Grid&>(solver.get_grid());
        if (poke) {  // same edits as tests/test_gpu_cpp_driver.py applies to the oracle
            g.f_current(3, 4, 1) *= 1.25;
            g.f_current(10, 7, 6) += 0.01;
            g.f_current(p.nx, p.ny, 8) *= 0.5;
        }
        if (!solver.run(io)) return 1;
        const LBM::Grid& c = solver.get_grid();
        std::vector<double> fc((size_t)c.total_nx() * c.total_ny() * LBM::Q), fn(fc.size());
        size_t k = 0;
        for (int gy = 0; gy < c.total_ny(); ++gy)
            for (int gx = 0; gx < c.total_nx(); ++gx)
                for (int i = 0; i < LBM::Q; ++i, ++k) {
                    fc[k] = c.f_current(gx, gy, i);
                    fn[k] = c.f_next(gx, gy, i);
                }
        std::vector<double> rho((size_t)c.local_nx() * c.local_ny()), ux(rho.size()), uy(rho.size()), solid(rho.size());
        k = 0;
        for (int y = 0; y < c.local_ny(); ++y)
            for (int x = 0; x < c.local_nx(); ++x, ++k) {
                rho[k] = c.rho(x, y);
                ux[k] = c.ux(x, y);
                uy[k] = c.uy(x, y);
                solid[k] = c.is_solid(x, y) ? 1.0 : 0.0;
            }
        write_doubles(out + "/f_current.bin", fc);
        write_doubles(out + "/f_next.bin", fn);
        write_doubles(out + "/rho.bin", rho);
        write_doubles(out + "/ux.bin", ux);
        write_doubles(out + "/uy.bin", uy);
        write_doubles(out + "/solid.bin", solid);
        std::printf("getters %d %d %d %d %d %d %d %d %d %d %d %d\n", c.x_start(), c.y_start(), c.local_nx(), c.local_ny(), c.total_nx(),
                    c.total_ny(), c.global_nx(), c.global_ny(), c.mpi_rank(), c.mpi_size(), (int)c.is_left_boundary(),
                    (int)c.is_right_boundary());
        std::printf("stable %d maxvel %.17g fptr %d\n", (int)c.check_stability(), c.max_velocity(),
                    (int)(c.f_current_ptr(2, 3)[4] == c.f_current(2, 3, 4) && c.f_next_ptr(2, 3)[7] == c.f_next(2, 3, 7)));
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 3;
    }
    return 0;
}
