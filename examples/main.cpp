// examples/main.cpp -- cylinder-flow driver on the B200 engine.
//
// Same flow as the reference's src/main.cpp:7-43 (parameters -> Solver -> IOManager ->
// initialise -> run -> write_final_results) without MPI: one process per GPU, started directly
// (1 GPU) or by `torchrun --nproc-per-node N --no-python ./lbm_solver ...` (N x-slabs).  Unlike
// the reference, whose parameters are compile-time defaults, every SimulationParams field can be
// set on the command line; with no arguments it runs the reference's default case with VTK on.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>

#include "../include/LBMConfig.h"
#include "../include/LBMIO.h"
#include "../include/LBMSolver.h"

namespace {
void usage() {
    std::cout << "lbm_solver [--nx N] [--ny N] [--steps N] [--of N] [--tau X] [--uin X] [--cx X] [--cy X] [--cr X]\n"
                 "           [--vtk 0|1] [--vtk-start N] [--sync-vtk] [--vtk-binary] [--periodic-x] [--periodic-y] [--no-cylinder]\n"
                 "           [--shear-wave] [--aa] [--fx X] [--fy X] [--no-final]\n"
                 "           [--checkpoint FILE] [--restart FILE]\n";
}
}  // namespace

int main(int argc, char* argv[]) {
    LBM::SimulationParams params;
    bool vtk = true, final_results = true;
    std::string checkpoint_out, restart_from;
    for (int a = 1; a < argc; ++a) {
        const std::string k = argv[a];
        auto val = [&]() -> const char* { return a + 1 < argc ? argv[++a] : "0"; };
        if (k == "--nx") params.nx = std::atoi(val());
        else if (k == "--ny") params.ny = std::atoi(val());
        else if (k == "--steps") params.num_timesteps = std::atoi(val());
        else if (k == "--of") params.output_frequency = std::atoi(val());
        else if (k == "--tau") params.tau = std::atof(val());
        else if (k == "--uin") params.inlet_velocity = std::atof(val());
        else if (k == "--cx") params.cylinder_x = std::atof(val());
        else if (k == "--cy") params.cylinder_y = std::atof(val());
        else if (k == "--cr") params.cylinder_radius = std::atof(val());
        else if (k == "--vtk") vtk = std::atoi(val()) != 0;
        else if (k == "--vtk-start") params.vtk_start_step = std::atoi(val());
        else if (k == "--sync-vtk") params.async_vtk = false;
        else if (k == "--vtk-binary") params.vtk_binary = true;
        else if (k == "--periodic-x") params.flags |= LBM_FLAG_PERIODIC_X;
        else if (k == "--periodic-y") params.flags |= LBM_FLAG_PERIODIC_Y;
        else if (k == "--no-cylinder") params.flags |= LBM_FLAG_NO_CYLINDER;
        else if (k == "--shear-wave") params.flags |= LBM_FLAG_SHEAR_WAVE_INIT;
        else if (k == "--aa") params.flags |= LBM_FLAG_AA;
        else if (k == "--fx") params.body_force_x = std::atof(val());
        else if (k == "--fy") params.body_force_y = std::atof(val());
        else if (k == "--no-final") final_results = false;
        else if (k == "--checkpoint") checkpoint_out = val();
        else if (k == "--restart") restart_from = val();
        else {
            usage();
            return k == "--help" || k == "-h" ? 0 : 2;
        }
    }

    try {
        LBM::Solver solver(params, vtk);
        LBM::IOManager io_manager;
        const bool root = solver.get_grid().mpi_rank() == 0;

        solver.initialise();
        if (!restart_from.empty()) solver.load_checkpoint(restart_from);
        const auto t0 = std::chrono::steady_clock::now();
        const bool success = solver.run(io_manager);
        const double seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

        if (!success) {
            if (root) std::cerr << "LBM simulation failed." << std::endl;
            return 1;
        }
        if (root)
            std::cout << "run(): " << seconds << " s, "
                      << (double)params.nx * params.ny * params.num_timesteps / seconds / 1e6 << " MLUPS (wall clock, output included)"
                      << std::endl;
        if (!checkpoint_out.empty()) solver.save_checkpoint(checkpoint_out);
        if (final_results) io_manager.write_final_results(solver.get_grid(), solver.get_params());
        if (root) std::cout << "\nSimulation completed successfully!" << std::endl;
    } catch (const std::exception& e) {
        std::cerr << "An exception occurred: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
