// LBMConfig.h -- lattice constants and run parameters of the B200 D2Q9 BGK solver.
//
// Drop-in for the reference header of the same name (its include/LBMConfig.h:9-66): identical
// names, index conventions, defaults and derived quantities, so that src/main.cpp and any code
// written against LBM::SimulationParams compiles unchanged.  Direction order (reference
// :13-25): 0 rest, 1 E, 2 N, 3 W, 4 S, 5 NE, 6 NW, 7 SW, 8 SE.
//
// Extensions (fields after vtk_start_step) default to the reference's behaviour.
#pragma once

#include <array>
#include <vector>

#include "lbm_b200.h"

namespace LBM {

constexpr int Q = 9;  // populations per cell
constexpr int D = 2;  // space dimensions

constexpr std::array<std::array<int, 2>, Q> VELOCITIES = {{
    {{0, 0}}, {{1, 0}}, {{0, 1}}, {{-1, 0}}, {{0, -1}}, {{1, 1}}, {{-1, 1}}, {{-1, -1}}, {{1, -1}}}};

constexpr double W_REST = 4.0 / 9.0, W_AXIS = 1.0 / 9.0, W_DIAG = 1.0 / 36.0;
constexpr std::array<double, Q> WEIGHTS = {W_REST, W_AXIS, W_AXIS, W_AXIS, W_AXIS, W_DIAG, W_DIAG, W_DIAG, W_DIAG};

// OPPOSITE[i] is the direction with c = -c_i
constexpr std::array<int, Q> OPPOSITE = {0, 3, 4, 1, 2, 7, 8, 5, 6};

struct SimulationParams {
    // --- the reference's fields and defaults (its include/LBMConfig.h:37-52) ---
    double tau = 0.6;                 // BGK relaxation time
    double inlet_velocity = 0.01333;  // lattice units
    int nx = 2048;
    int ny = 512;
    int num_timesteps = 120000;
    int output_frequency = 140;  // forces / log / VTK cadence
    double cylinder_x = 0.2;     // centre, fraction of nx
    double cylinder_y = 0.5;     // centre, fraction of ny
    double cylinder_radius = 0.05;  // fraction of ny
    int vtk_start_step = 0;

    // --- extensions (lbm_b200.h LBM_FLAG_*; all zero = the reference's channel) ---
    int flags = 0;
    double body_force_x = 0.0;
    double body_force_y = 0.0;
    bool async_vtk = true;   // VTK frames leave through pinned snapshots + a writer thread
    bool vtk_binary = false;  // legacy-VTK BINARY frames instead of the reference's ASCII ones

    // kinematic viscosity and Reynolds number exactly as the reference derives them (:54-58);
    // note Re uses the real-valued diameter 2*cylinder_radius*ny, C_D/C_L the integer one
    double nu() const { return (tau - 0.5) / 3.0; }
    double reynolds() const { return inlet_velocity * (2.0 * cylinder_radius * ny) / nu(); }

    // cylinder in integer cells, truncated (:61-65)
    int get_cylinder_x() const { return static_cast<int>(cylinder_x * nx); }
    int get_cylinder_y() const { return static_cast<int>(cylinder_y * ny); }
    int get_cylinder_radius_cells() const { return static_cast<int>(cylinder_radius * ny); }

    // the C-ABI view of these parameters
    lbm_params to_c() const {
        lbm_params c{};
        c.tau = tau;
        c.inlet_velocity = inlet_velocity;
        c.nx = nx;
        c.ny = ny;
        c.num_timesteps = num_timesteps;
        c.output_frequency = output_frequency;
        c.cylinder_x = cylinder_x;
        c.cylinder_y = cylinder_y;
        c.cylinder_radius = cylinder_radius;
        c.vtk_start_step = vtk_start_step;
        c.flags = flags;
        c.body_force_x = body_force_x;
        c.body_force_y = body_force_y;
        return c;
    }
};

}  // namespace LBM
