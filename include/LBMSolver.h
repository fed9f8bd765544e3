// LBMSolver.h -- LBM::Solver with the reference's public surface (its include/LBMSolver.h:23,
// 31, 43, 80-81) driving the B200 engine.
//
// The reference's loop body -- collision_step, record_forces, exchange_ghost_cells,
// streaming_step, apply_boundary_conditions, check_stability (:48-64) -- is ONE call per
// chunk of iterations here: lbm_run launches the fused pull+boundary+collide kernels, the NCCL
// halo exchange, the link-list force reduction on output steps and the stability flag, and hands
// back the forces rows and the first unstable timestep.  Chunks end on the reference's output
// steps, so the log lines, forces.csv rows, VTK frames and the "Simulation unstable at
// timestep t" message appear for the same t as in the reference.
#pragma once

#include <sys/stat.h>

#include <algorithm>
#include <cstdio>
#include <iomanip>
#include <iostream>
#include <memory>

#include "LBMConfig.h"
#include "LBMGrid.h"
#include "LBMIO.h"
#include "LBMUtils.h"

namespace LBM {

class Solver {
   public:
    explicit Solver(const SimulationParams& params, bool enable_vtk = false)
        : params_(params), grid_(params), enable_vtk_output_(enable_vtk) {
        if (enable_vtk && grid_.mpi_rank() == 0) mkdir("vtk_output", 0755);
    }

    void initialise() {
        if (grid_.mpi_rank() == 0) {
            std::cout << "Cylinder Flow LBM Parameters:\n";
            std::cout << "  Domain: " << params_.nx << "×" << params_.ny << "\n";
            std::cout << "  tau = " << params_.tau << ", nu = " << params_.nu() << "\n";
            std::cout << "  Inlet velocity = " << params_.inlet_velocity << "\n";
            std::cout << "  Reynolds number = " << params_.reynolds() << std::endl;
        }
        grid_.setup_geometry(params_);
        grid_.initialise(params_.inlet_velocity);
    }

    // false = the run went unstable (message on stderr), as the reference.
    bool run(IOManager& io_manager) {
        if (grid_.mpi_rank() == 0) std::cout << "Starting LBM cylinder flow simulation..." << std::endl;
        const int T = params_.num_timesteps;
        const int of = std::max(params_.output_frequency, 1);
        std::unique_ptr<FrameWriter> frames;
        if (enable_vtk_output_ && params_.async_vtk) frames = std::make_unique<FrameWriter>(grid_, params_.vtk_binary);

        int t = first_timestep_;
        while (t < T) {
            // this chunk runs iterations t .. last, where last is the next output step (or T-1)
            const int next_out = ((t + of - 1) / of) * of;
            const int last = std::min(next_out, T - 1);
            double rows[2][5];
            int n_rows = 0;
            const int unstable_at = grid_.advance(last - t + 1, &rows[0][0], 2, &n_rows);
            for (int k = 0; k < n_rows; ++k) {
                double f[2] = {rows[k][1], rows[k][2]};  // slab-local sums -> whole cylinder
                grid_.check(lbm_allreduce(grid_.handle(), f, 2, LBM_SUM));
                io_manager.write_force_row((int)rows[k][0], f[0], f[1], params_);
            }
            if (unstable_at >= 0) {
                if (grid_.mpi_rank() == 0) std::fprintf(stderr, "Simulation unstable at timestep %d\n", unstable_at);
                if (frames) frames->finish();
                return false;
            }
            if (last > 0 && last % of == 0) {
                const double max_vel = grid_.max_velocity();
                if (grid_.mpi_rank() == 0)
                    std::cout << "Timestep " << last << ": max_vel=" << std::fixed << std::setprecision(6) << max_vel
                              << std::endl;
                if (enable_vtk_output_ && last >= params_.vtk_start_step) write_vtk_frame(last, frames.get());
            }
            t = last + 1;
        }
        if (frames) frames->finish();
        return true;
    }

    // Extensions: checkpoint after run() / restart before run() (state = f_current + timestep).
    void save_checkpoint(const std::string& path) const { grid_.save_checkpoint(path); }
    void load_checkpoint(const std::string& path) { first_timestep_ = grid_.load_checkpoint(path); }

    const Grid& get_grid() const { return grid_; }
    const SimulationParams& get_params() const { return params_; }

   private:
    // rho/ux/uy of the whole channel -> vtk_output/lbm_%06d.vtk (reference :269-362).
    void write_vtk_frame(int timestep, FrameWriter* frames) {
        if (frames) {
            frames->submit(timestep);
            return;
        }
        std::vector<double> rho, ux, uy;
        if (grid_.mpi_rank() == 0) {
            const size_t n = (size_t)grid_.global_nx() * grid_.global_ny();
            rho.resize(n);
            ux.resize(n);
            uy.resize(n);
        }
        grid_.check(lbm_gather_macros(grid_.handle(), rho.data(), ux.data(), uy.data()));
        if (grid_.mpi_rank() == 0) {
            if (params_.vtk_binary)
                IOManager::write_vtk_arrays_binary(ux.data(), uy.data(), rho.data(), params_.nx, params_.ny, timestep);
            else
                IOManager::write_vtk_timestep(ux, uy, rho, params_, timestep);
        }
    }

    SimulationParams params_;
    Grid grid_;
    bool enable_vtk_output_;
    int first_timestep_ = 0;  // > 0 after load_checkpoint: run() continues at that iteration
};

}  // namespace LBM
