// LBMGrid.h -- LBM::Grid with the reference's public surface (its include/LBMGrid.h:57-150,
// 152, 185, 249, 285, 319), backed by one x-slab of the B200 engine (liblbm_b200.so).
//
// What changed underneath:
//   * the populations live on the GPU as fp64 SoA planes; the reference's padded AoS arrays
//     exist only as a host-side cache filled on demand (lbm_download_f), so the element
//     accessors f_current(gx,gy,i) / f_next / rho / ux / uy / is_solid keep their meaning and
//     index conventions (ghost-inclusive coordinates for populations, interior ones otherwise);
//   * MPI_Cart_create + 2-D decomposition (:347-392) became contiguous x-slabs, one process per
//     GPU, ranks discovered by lbm_bootstrap_env; mpi_rank()/mpi_size() report slab and count;
//   * exchange_ghost_cells() is fused into the step kernels (stores into the neighbouring GPU's memory; NCCL
//     send/recv where peer access is unavailable) and is a no-op here;
//   * the constructor's banner names the slab partition instead of the reference's MPI grid (same fields, different
//     first line).
// Errors surface as std::runtime_error carrying lbm_last_error(), which src/main.cpp:29 catches.
#pragma once

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "LBMConfig.h"
#include "lbm_b200.h"

namespace LBM {

class Grid {
   public:
    Grid(int nx, int ny) : Grid(make_params(nx, ny)) {}

    // Extension: create with the full parameter set (periodic flags must be known at creation).
    explicit Grid(const SimulationParams& params) : params_(params) {
        unsigned char id[128];
        int local = 0;
        check_create(lbm_bootstrap_env(&rank_, &world_, &local, id));
        if (world_ > 1)  // a tag every slab process of THIS launch shares and no other launch does (the NCCL id is random)
            for (int k = 0; k < 128; ++k) job_tag_ = job_tag_ * 1099511628211ULL ^ id[k];
        int ndev = 0;
        lbm_device_count(&ndev);
        const int device = ndev > 0 ? local % ndev : 0;
        lbm_params c = params_.to_c();
        check_create(world_ == 1 ? lbm_create(&c, device, &h_) : lbm_create_slab(&c, device, rank_, world_, id, &h_));
        refresh_info();
        if (rank_ == 0) {
            std::printf("B200 x-slab grid (fp64 SoA, A-B double buffer)\n");
            std::printf("  Global domain: %dx%d\n", info_.global_nx, info_.global_ny);
            std::printf("  GPU slabs: %d (%dx1 grid)\n", world_, world_);
            std::printf("  Local interior per slab: %dx%d\n", info_.local_nx, info_.local_ny);
            std::printf("  Local with ghosts: %dx%d\n", total_nx(), total_ny());
            std::printf("  Ghost layers: 1\n");
            std::printf("  Device memory per slab: %.2f MB\n", 2.0 * (double)info_.bytes_per_buffer / (1024.0 * 1024.0));
        }
    }

    ~Grid() { lbm_destroy(h_); }
    Grid(const Grid&) = delete;
    Grid& operator=(const Grid&) = delete;

    // ---- index helpers of the host-side views (reference :105-111) ----
    size_t get_f_index(int x, int y, int i) const { return f_index(x, y, i); }
    size_t get_interior_index(int x, int y) const { return m_index(x, y); }

    // ---- population accessors, ghost-inclusive coordinates (reference :115-121) ----
    const double& f_current(int x, int y, int i) const { return fetch(LBM_F_CURRENT)[f_index(x, y, i)]; }
    const double& f_next(int x, int y, int i) const { return fetch(LBM_F_NEXT)[f_index(x, y, i)]; }
    // Writable view of f_current: edits are uploaded (lbm_upload_f) before the next step, which
    // is how a caller installs its own initial condition.
    double& f_current(int x, int y, int i) {
        fetch(LBM_F_CURRENT);
        f_dirty_ = true;
        return cache_f_[LBM_F_CURRENT][f_index(x, y, i)];
    }
    // Writable view of f_next (reference :119-121).  As in the reference, a write at an iteration boundary matters only
    // in solid cells and S/N ghost rows (the next collision overwrites every fluid cell): lbm_upload_f_next.
    double& f_next(int x, int y, int i) {
        fetch(LBM_F_NEXT);
        fn_dirty_ = true;
        return cache_f_[LBM_F_NEXT][f_index(x, y, i)];
    }
    const double* f_current_ptr(int x, int y) const { return &f_current(x, y, 0); }
    const double* f_next_ptr(int x, int y) const { return &f_next(x, y, 0); }
    double* f_current_ptr(int x, int y) { return &f_current(x, y, 0); }  // (reference :116, :120: mutable pointers)
    double* f_next_ptr(int x, int y) { return &f_next(x, y, 0); }

    // ---- macroscopic fields and mask, interior coordinates (reference :124-129, :145) ----
    const double& rho(int x, int y) const { return macros()[0][m_index(x, y)]; }
    const double& ux(int x, int y) const { return macros()[1][m_index(x, y)]; }
    const double& uy(int x, int y) const { return macros()[2][m_index(x, y)]; }
    // Writable macroscopic fields (reference :124-126).  The reference's next collision overwrites them in every fluid
    // cell and its boundary pass in the inlet / outlet / solid cells, so a write is visible to the output routines until
    // the run advances -- exactly that is kept: the edit lives in the host-side view and is dropped by the next advance.
    double& rho(int x, int y) { return const_cast<std::vector<double>*>(macros())[0][m_index(x, y)]; }
    double& ux(int x, int y) { return const_cast<std::vector<double>*>(macros())[1][m_index(x, y)]; }
    double& uy(int x, int y) { return const_cast<std::vector<double>*>(macros())[2][m_index(x, y)]; }
    bool is_solid(int x, int y) const {
        if (solid_.empty()) {
            solid_.resize((size_t)local_nx() * local_ny());
            check(lbm_download_solid(h_, solid_.data()));
        }
        return solid_[m_index(x, y)] != 0;
    }

    // ---- topology getters (reference :132-150) ----
    int x_start() const { return info_.x_start; }
    int y_start() const { return info_.y_start; }
    int local_nx() const { return info_.local_nx; }
    int local_ny() const { return info_.local_ny; }
    int total_nx() const { return info_.local_nx + 2; }
    int total_ny() const { return info_.local_ny + 2; }
    int global_nx() const { return info_.global_nx; }
    int global_ny() const { return info_.global_ny; }
    int mpi_rank() const { return rank_; }
    int mpi_size() const { return world_; }
    bool is_left_boundary() const { return rank_ == 0; }
    bool is_right_boundary() const { return rank_ == world_ - 1; }
    bool is_bottom_boundary() const { return true; }  // slabs span the full height
    bool is_top_boundary() const { return true; }

    // ---- set-up (reference :152-183, :185-246) ----
    void setup_geometry(const SimulationParams& params) {
        adopt(params);
        int local_solids = 0;
        check(lbm_setup_geometry(h_, &local_solids));
        double total = local_solids;
        check(lbm_allreduce(h_, &total, 1, LBM_SUM));
        refresh_info();
        invalidate();
        solid_.clear();
        if (rank_ == 0) {
            std::printf("  Cylinder: center=(%d,%d), radius=%d cells\n", params.get_cylinder_x(), params.get_cylinder_y(),
                        params.get_cylinder_radius_cells());
            std::printf("  Solid cells: %d\n", (int)total);
        }
    }

    void initialise(double inlet_u) {
        check(lbm_initialise(h_, inlet_u));
        invalidate();
    }

    // ---- per-step services (reference :249, :285, :319) ----
    void exchange_ghost_cells() {}  // fused into lbm_step: halo columns travel by NCCL on their own stream

    bool check_stability() const {
        int ok = 1, first_bad = -1;
        check(lbm_check_stability(h_, &ok, &first_bad));  // global verdict (all-reduce MIN over slabs)
        return ok != 0;
    }

    double max_velocity() const {
        double v = 0.0;
        check(lbm_max_velocity(h_, &v));
        check(lbm_allreduce(h_, &v, 1, LBM_MAX));
        return v;
    }

    // ---- engine access for Solver / IOManager ----
    lbm_handle handle() const { return h_; }
    unsigned long long job_tag() const { return job_tag_; }
    const SimulationParams& params() const { return params_; }
    void adopt(const SimulationParams& params) {
        lbm_params c = params.to_c();
        check(lbm_set_params(h_, &c));
        params_ = params;
    }
    // Advance n iterations with Solver::run's observable behaviour; see lbm_run.
    int advance(int n, double* rows, int max_rows, int* n_rows) {
        flush_edits();
        int unstable_at = -1;
        check(lbm_run(h_, n, rows, max_rows, n_rows, &unstable_at));
        invalidate();
        return unstable_at;
    }
    // ---- checkpoint / restart (extension; the reference keeps its state in RAM only) ----
    // One file per slab: a 64-byte header and the padded AoS f_current (reference layout,
    // include/LBMGrid.h:105-107) of this slab.  Restarting from it continues the run bit for bit.
    void save_checkpoint(const std::string& path) const {
        lbm_info now;
        check(lbm_get_info(h_, &now));
        const std::vector<double>& f = fetch(LBM_F_CURRENT);
        std::FILE* fp = std::fopen(slab_file(path).c_str(), "wb");
        if (!fp) throw std::runtime_error("lbm_b200: cannot write " + slab_file(path));
        const int64_t head[8] = {0x4c424d3230304231LL, info_.global_nx, info_.global_ny, info_.local_nx, info_.x_start,
                                 now.iteration, world_, (int64_t)f.size()};
        const bool ok = std::fwrite(head, sizeof(head), 1, fp) == 1 && std::fwrite(f.data(), sizeof(double), f.size(), fp) == f.size();
        std::fclose(fp);
        if (!ok) throw std::runtime_error("lbm_b200: short write to " + slab_file(path));
    }
    // Returns the iteration the next step will execute.
    int load_checkpoint(const std::string& path) {
        std::FILE* fp = std::fopen(slab_file(path).c_str(), "rb");
        if (!fp) throw std::runtime_error("lbm_b200: cannot read " + slab_file(path));
        int64_t head[8];
        std::vector<double> f((size_t)total_nx() * total_ny() * Q);
        const bool ok = std::fread(head, sizeof(head), 1, fp) == 1 && head[0] == 0x4c424d3230304231LL &&
                        head[1] == info_.global_nx && head[2] == info_.global_ny && head[3] == info_.local_nx &&
                        head[4] == info_.x_start && head[6] == world_ && head[7] == (int64_t)f.size() &&
                        std::fread(f.data(), sizeof(double), f.size(), fp) == f.size();
        std::fclose(fp);
        if (!ok) throw std::runtime_error("lbm_b200: " + slab_file(path) + " does not match this grid / slab layout");
        check(lbm_upload_f(h_, f.data(), (int)head[5]));
        invalidate();
        return (int)head[5];
    }

    void check(int rc) const {
        if (rc != LBM_OK) throw std::runtime_error(std::string("lbm_b200: ") + lbm_last_error(h_));
    }

   private:
    static SimulationParams make_params(int nx, int ny) {
        SimulationParams p;
        p.nx = nx;
        p.ny = ny;
        return p;
    }
    static void check_create(int rc) {
        if (rc != LBM_OK) throw std::runtime_error(std::string("lbm_b200: ") + lbm_last_error(nullptr));
    }
    void refresh_info() { check(lbm_get_info(h_, &info_)); }
    std::string slab_file(const std::string& path) const {
        return world_ == 1 ? path : path + ".slab" + std::to_string(rank_) + "of" + std::to_string(world_);
    }
    size_t f_index(int x, int y, int i) const { return ((size_t)y * total_nx() + x) * Q + i; }  // reference :105-107
    size_t m_index(int x, int y) const { return (size_t)y * local_nx() + x; }                   // reference :109-111

    void invalidate() const {
        have_f_[0] = have_f_[1] = have_m_ = false;
        f_dirty_ = fn_dirty_ = false;
    }
    void flush_edits() {
        if (f_dirty_) {
            lbm_info now;
            check(lbm_get_info(h_, &now));
            check(lbm_upload_f(h_, cache_f_[LBM_F_CURRENT].data(), now.iteration));
            f_dirty_ = false;
        }
        if (fn_dirty_) {
            check(lbm_upload_f_next(h_, cache_f_[LBM_F_NEXT].data()));
            fn_dirty_ = false;
        }
    }
    const std::vector<double>& fetch(int which) const {
        if (!have_f_[which]) {
            cache_f_[which].resize((size_t)total_nx() * total_ny() * Q);
            check(lbm_download_f(h_, which, cache_f_[which].data()));
            have_f_[which] = true;
        }
        return cache_f_[which];
    }
    const std::vector<double>* macros() const {
        if (!have_m_) {
            const size_t n = (size_t)local_nx() * local_ny();
            for (auto& v : cache_m_) v.resize(n);
            check(lbm_download_macros(h_, cache_m_[0].data(), cache_m_[1].data(), cache_m_[2].data()));
            have_m_ = true;
        }
        return cache_m_;
    }

    SimulationParams params_;
    lbm_handle h_ = nullptr;
    lbm_info info_{};
    int rank_ = 0, world_ = 1;
    unsigned long long job_tag_ = 1469598103934665603ULL;
    mutable std::vector<double> cache_f_[2];
    mutable std::vector<double> cache_m_[3];
    mutable std::vector<unsigned char> solid_;
    mutable bool have_f_[2] = {false, false};
    mutable bool have_m_ = false;
    mutable bool f_dirty_ = false, fn_dirty_ = false;
};

}  // namespace LBM
