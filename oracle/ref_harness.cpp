/*
 * oracle/ref_harness.cpp -- drives the UNMODIFIED reference headers.  TEST INFRASTRUCTURE ONLY.
 *
 * Includes /root/reference/include/{LBMConfig,LBMGrid,LBMSolver,LBMIO}.h as they lie (never
 * copied into this repo) behind the single-rank MPI shim in oracle/shim/mpi.h, sets
 * LBM::SimulationParams fields from the command line, and runs the reference's own
 * Solver::initialise() / Solver::run() (reference include/LBMSolver.h:31,43).
 *
 * Two jobs:
 *   dump mode   (--dump DIR): run N steps, write the observable state exactly as the public const
 *               accessors expose it (include/LBMGrid.h:115-129,145): padded AoS f_current / f_next,
 *               interior rho / ux / uy, the solid mask, and forces.csv (written by the reference's
 *               own IOManager into DIR).  These pin oracle/lbm_oracle.c and feed tests/golden/.
 *   time mode   (--time): warm-up run + timed run of Solver::run, VTK off, printing
 *               one JSON line with MLUPS.  Used by bench.py as the CPU baseline ("kind":"reference").
 *
 * Built by oracle/Makefile with -fno-access-control so that time mode can change the private
 * params_.num_timesteps between the warm-up and the timed call; no reference code is altered.
 */
#include <mpi.h>  // the shim, via -Ioracle/shim

#include "LBMConfig.h"
#include "LBMGrid.h"
#include "LBMIO.h"
#include "LBMSolver.h"

#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <string>
#include <unistd.h>
#include <vector>

namespace {

void write_doubles(const std::string& path, const std::vector<double>& v) {
    std::ofstream f(path, std::ios::binary);
    f.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(double)));
}

void dump_state(const LBM::Grid& g, const std::string& dir) {
    const int tnx = g.total_nx(), tny = g.total_ny(), nx = g.local_nx(), ny = g.local_ny();
    std::vector<double> fc(static_cast<size_t>(tnx) * tny * LBM::Q), fn(fc.size());
    size_t k = 0;
    for (int gy = 0; gy < tny; ++gy)
        for (int gx = 0; gx < tnx; ++gx)
            for (int i = 0; i < LBM::Q; ++i, ++k) {
                fc[k] = g.f_current(gx, gy, i);
                fn[k] = g.f_next(gx, gy, i);
            }
    std::vector<double> rho(static_cast<size_t>(nx) * ny), ux(rho.size()), uy(rho.size());
    std::vector<uint8_t> solid(rho.size());
    k = 0;
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x, ++k) {
            rho[k] = g.rho(x, y);
            ux[k] = g.ux(x, y);
            uy[k] = g.uy(x, y);
            solid[k] = g.is_solid(x, y) ? 1 : 0;
        }
    write_doubles(dir + "/f_current.bin", fc);
    write_doubles(dir + "/f_next.bin", fn);
    write_doubles(dir + "/rho.bin", rho);
    write_doubles(dir + "/ux.bin", ux);
    write_doubles(dir + "/uy.bin", uy);
    std::ofstream f(dir + "/solid.bin", std::ios::binary);
    f.write(reinterpret_cast<const char*>(solid.data()), static_cast<std::streamsize>(solid.size()));
}

}  // namespace

int main(int argc, char** argv) {
    LBM::SimulationParams p;
    std::string dump_dir;
    bool time_mode = false;
    int warmup = 3;
    for (int a = 1; a < argc; ++a) {
        std::string k = argv[a];
        auto val = [&]() -> const char* { return (a + 1 < argc) ? argv[++a] : "0"; };
        if (k == "--nx") p.nx = std::atoi(val());
        else if (k == "--ny") p.ny = std::atoi(val());
        else if (k == "--steps") p.num_timesteps = std::atoi(val());
        else if (k == "--of") p.output_frequency = std::atoi(val());
        else if (k == "--tau") p.tau = std::atof(val());
        else if (k == "--uin") p.inlet_velocity = std::atof(val());
        else if (k == "--cx") p.cylinder_x = std::atof(val());
        else if (k == "--cy") p.cylinder_y = std::atof(val());
        else if (k == "--cr") p.cylinder_radius = std::atof(val());
        else if (k == "--dump") dump_dir = val();
        else if (k == "--time") time_mode = true;
        else if (k == "--warmup") warmup = std::atoi(val());
        else { std::fprintf(stderr, "unknown argument %s\n", k.c_str()); return 2; }
    }

    MPI_Init(&argc, &argv);
    if (!dump_dir.empty() && chdir(dump_dir.c_str()) != 0) {  // IOManager opens forces.csv in cwd (LBMIO.h:38)
        std::perror("chdir");
        return 2;
    }

    int rc = 0;
    {
        LBM::Solver solver(p, /*enable_vtk=*/false);
        LBM::IOManager io;
        solver.initialise();

        if (time_mode) {
            const int timed_steps = p.num_timesteps;
            solver.params_.num_timesteps = warmup;  // -fno-access-control
            bool ok = solver.run(io);
            solver.params_.num_timesteps = timed_steps;
            auto t0 = std::chrono::steady_clock::now();
            ok = solver.run(io) && ok;
            auto t1 = std::chrono::steady_clock::now();
            const double s = std::chrono::duration<double>(t1 - t0).count();
            const double mlups = static_cast<double>(p.nx) * p.ny * timed_steps / s / 1e6;
            std::fprintf(stderr,
                         "{\"mlups\": %.4f, \"seconds\": %.6f, \"steps\": %d, \"warmup\": %d, \"nx\": %d, \"ny\": %d, "
                         "\"threads\": %d, \"stable\": %s}\n",
                         mlups, s, timed_steps, warmup, p.nx, p.ny, omp_get_max_threads(), ok ? "true" : "false");
            rc = ok ? 0 : 1;
        } else {
            const bool ok = solver.run(io);
            if (!dump_dir.empty()) dump_state(solver.get_grid(), ".");
            rc = ok ? 0 : 1;
        }
    }  // ~IOManager closes forces.csv
    MPI_Finalize();
    return rc;
}
