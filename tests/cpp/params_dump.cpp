// tests/cpp/params_dump.cpp -- prints the lattice constants and the derived quantities of
// LBM::SimulationParams.  Compiled once against this repo's include/LBMConfig.h and once against the
// reference's own header (-I/root/reference/include, where mounted): the outputs must be identical.
#include <cstdio>

#include "LBMConfig.h"

int main() {
    std::printf("Q %d D %d\n", LBM::Q, LBM::D);
    for (int i = 0; i < LBM::Q; ++i)
        std::printf("i %d c %d %d w %.17g opp %d\n", i, LBM::VELOCITIES[i][0], LBM::VELOCITIES[i][1], LBM::WEIGHTS[i], LBM::OPPOSITE[i]);
    LBM::SimulationParams p;
    std::printf("defaults %.17g %.17g %d %d %d %d %.17g %.17g %.17g %d\n", p.tau, p.inlet_velocity, p.nx, p.ny, p.num_timesteps,
                p.output_frequency, p.cylinder_x, p.cylinder_y, p.cylinder_radius, p.vtk_start_step);
    const double taus[] = {0.6, 0.52, 0.9095}, us[] = {0.01333, 0.1333, 0.0325520833};
    const int sizes[][2] = {{2048, 512}, {8192, 2048}, {32768, 8192}, {70, 33}, {6, 4}};
    for (double tau : taus)
        for (double u : us)
            for (auto& s : sizes) {
                p.tau = tau;
                p.inlet_velocity = u;
                p.nx = s[0];
                p.ny = s[1];
                std::printf("derived %.17g %.17g %d %d %d\n", p.nu(), p.reynolds(), p.get_cylinder_x(), p.get_cylinder_y(),
                            p.get_cylinder_radius_cells());
            }
    return 0;
}
