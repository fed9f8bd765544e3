// LBMIO.h -- LBM::IOManager with the reference's public surface and file formats (its
// include/LBMIO.h:35, 55, 114, 194): forces.csv, vtk_output/lbm_%06d.vtk, velocity_field.csv,
// simulation_params.csv, byte-compatible so that scripts/lift.py and
// scripts/visualise_results.py keep working.
//
// What changed underneath:
//   * the momentum-exchange sum runs on the GPU over a precomputed link list (k_forces); the
//     MPI_Reduce pair of :167-168 is one NCCL all-reduce (lbm_allreduce);
//   * fields reach the host through lbm_gather_macros (one pinned image, slabs copied with
//     strided D2H) instead of MPI_Gather + 3 x MPI_Gatherv (:225-300);
//   * VTK frames can leave asynchronously (FrameWriter): a pinned double buffer filled by
//     lbm_snapshot_begin_slot on the copy stream, formatted and written by a host thread while
//     the GPU keeps stepping;
//   * decimal formatting is done by an exact "%.8f" routine spread over the host cores (the
//     reference's iostream writer needs ~2 s for a 50 MB frame).
#pragma once

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "LBMConfig.h"
#include "LBMGrid.h"
#include "lbm_b200.h"

namespace LBM {

namespace textio {

// Writes v exactly as printf("%.8f") / std::fixed << std::setprecision(8) would (correctly
// rounded, ties to even, sign of zero kept) and returns the number of characters (at most
// FIXED8_MAX).  Fast path for |v| < 1e15: the integer part is exact, and the fraction times 1e8
// is split into a double product and its exact FMA error term, which decides the rounding of
// the 8th decimal without big-number arithmetic.
constexpr int FIXED8_MAX = 336;  // "-" + 309 digits + "." + 8 digits, rounded up

inline int fixed8(double v, char* out) {
    const double a = std::fabs(v);
    if (!(a < 1e15)) return std::snprintf(out, FIXED8_MAX, "%.8f", v);  // huge, inf, nan: the slow exact path
    const double whole = std::floor(a);
    const double fr = a - whole;  // exact, in [0, 1)
    const double hi = fr * 1e8;
    const double lo = std::fma(fr, 1e8, -hi);  // fr*1e8 == hi + lo exactly, |lo| <= ulp(hi)/2 << 0.5
    const double r = std::floor(hi);
    const double frac = hi - r;  // exact
    uint64_t ip = (uint64_t)whole;
    uint32_t fp = (uint32_t)r;
    // frac is a multiple of ulp(hi) and so is 0.5: unless frac == 0.5 the tiny lo cannot change
    // the side; at frac == 0.5 the sign of lo decides, and an exact tie goes to the even digit.
    // (frac == 0 with lo < 0 is r - eps, which still rounds to r.)
    if (frac > 0.5 || (frac == 0.5 && (lo > 0.0 || (lo == 0.0 && (fp & 1u))))) {
        if (++fp == 100000000u) {
            fp = 0;
            ++ip;
        }
    }
    char* p = out;
    if (std::signbit(v)) *p++ = '-';
    char tmp[24];
    int k = 0;
    do {
        tmp[k++] = (char)('0' + ip % 10);
        ip /= 10;
    } while (ip);
    while (k) *p++ = tmp[--k];
    *p++ = '.';
    for (int d = 7; d >= 0; --d) {
        p[d] = (char)('0' + fp % 10);
        fp /= 10;
    }
    p += 8;
    return (int)(p - out);
}

inline int integer(long long v, char* out) { return std::snprintf(out, 24, "%lld", v); }

// Formats items [0, n) with `fmt(index, char*) -> chars written` (at most max_chars_per_item
// each) on several host threads, block by block so that memory stays bounded, and writes the
// blocks to `f` in order.
inline void parallel_emit(std::FILE* f, size_t n, size_t max_chars_per_item,
                          const std::function<int(size_t, char*)>& fmt) {
    const size_t block = 1u << 20;
    unsigned hw = std::thread::hardware_concurrency();
    const unsigned nthreads = std::max(1u, std::min(hw ? hw : 1u, 16u));
    std::vector<std::vector<char>> bufs(nthreads);
    std::vector<size_t> used(nthreads);
    for (size_t b0 = 0; b0 < n; b0 += block) {
        const size_t b1 = std::min(n, b0 + block), per = (b1 - b0 + nthreads - 1) / nthreads;
        auto work = [&](unsigned t) {
            const size_t lo = std::min(b1, b0 + t * per), hi = std::min(b1, lo + per);
            std::vector<char>& buf = bufs[t];
            if (buf.size() < (hi - lo) * 48 + max_chars_per_item) buf.resize((hi - lo) * 48 + max_chars_per_item);
            size_t pos = 0;
            for (size_t k = lo; k < hi; ++k) {
                if (buf.size() - pos < max_chars_per_item) buf.resize(buf.size() * 2 + max_chars_per_item);
                pos += (size_t)fmt(k, buf.data() + pos);
            }
            used[t] = pos;
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
        for (unsigned t = 0; t < nthreads; ++t) std::fwrite(bufs[t].data(), 1, used[t], f);
    }
}

}  // namespace textio

class IOManager {
   public:
    IOManager() {
        int world = 1;
        lbm_bootstrap_env(&mpi_rank_, &world, nullptr, nullptr);
        if (mpi_rank_ == 0) {
            force_file_ = std::fopen("forces.csv", "w");
            if (force_file_)
                std::fputs("timestep,drag_force,lift_force,drag_coeff,lift_coeff\n", force_file_);
            else
                std::cerr << "ERROR: Could not open forces.csv\n";
        }
    }
    ~IOManager() {
        if (force_file_) std::fclose(force_file_);
    }
    IOManager(const IOManager&) = delete;
    IOManager& operator=(const IOManager&) = delete;

    // Momentum-exchange forces of the CURRENT post-collision state (reference :114-192): GPU link
    // reduction per slab, summed over slabs, one CSV row on rank 0.
    void record_forces(int timestep, const Grid& grid, const SimulationParams& params) {
        double f[2] = {0.0, 0.0};
        grid.check(lbm_get_forces(grid.handle(), &f[0], &f[1]));
        grid.check(lbm_allreduce(grid.handle(), f, 2, LBM_SUM));
        write_force_row(timestep, f[0], f[1], params);
    }

    // One forces.csv row from already reduced force components (Solver::run's fast path: the
    // rows come out of lbm_run).  Format and coefficients: reference :171-190.
    void write_force_row(int timestep, double fx, double fy, const SimulationParams& params) {
        if (mpi_rank_ != 0 || !force_file_) return;
        const double D_ref = 2.0 * params.get_cylinder_radius_cells();
        const double q_ref = 0.5 * 1.0 * params.inlet_velocity * params.inlet_velocity * D_ref;
        const double cd = (q_ref > 1e-12) ? fx / q_ref : 0.0;
        const double cl = (q_ref > 1e-12) ? fy / q_ref : 0.0;
        char line[4 * textio::FIXED8_MAX + 64], *p = line;
        p += textio::integer(timestep, p);
        for (double v : {fx, fy, cd, cl}) {
            *p++ = ',';
            p += textio::fixed8(v, p);
        }
        *p++ = '\n';
        std::fwrite(line, 1, (size_t)(p - line), force_file_);
        if (timestep % 10000 == 0) std::fflush(force_file_);
    }

    // ASCII legacy VTK, reference :55-111, byte for byte.
    static void write_vtk_timestep(const std::vector<double>& ux_g, const std::vector<double>& uy_g,
                                   const std::vector<double>& rho_g, const SimulationParams& p, int timestep) {
        write_vtk_arrays(ux_g.data(), uy_g.data(), rho_g.data(), p.nx, p.ny, timestep);
    }

    static void write_vtk_arrays(const double* ux, const double* uy, const double* rho, int nx, int ny, int timestep) {
        char filename[256];
        std::snprintf(filename, sizeof(filename), "vtk_output/lbm_%06d.vtk", timestep);
        std::FILE* f = std::fopen(filename, "w");
        if (!f) {
            std::cerr << "ERROR: Cannot write " << filename << "\n";
            return;
        }
        std::vector<char> big(1 << 22);
        std::setvbuf(f, big.data(), _IOFBF, big.size());
        std::fprintf(f, "# vtk DataFile Version 3.0\nLBM Flow Timestep %d\nASCII\nDATASET STRUCTURED_POINTS\n", timestep);
        std::fprintf(f, "DIMENSIONS %d %d 1\nORIGIN 0 0 0\nSPACING 1 1 1\nPOINT_DATA %d\n", nx, ny, nx * ny);
        const size_t n = (size_t)nx * ny;
        std::fputs("VECTORS velocity double\n", f);
        textio::parallel_emit(f, n, 2 * textio::FIXED8_MAX + 16, [&](size_t k, char* o) {
            char* q = o;
            q += textio::fixed8(ux[k], q);
            *q++ = ' ';
            q += textio::fixed8(uy[k], q);
            std::memcpy(q, " 0.0\n", 5);
            return (int)(q + 5 - o);
        });
        std::fputs("\nSCALARS velocity_magnitude double\nLOOKUP_TABLE default\n", f);
        textio::parallel_emit(f, n, textio::FIXED8_MAX + 16, [&](size_t k, char* o) {
            int c = textio::fixed8(std::sqrt(ux[k] * ux[k] + uy[k] * uy[k]), o);
            o[c] = '\n';
            return c + 1;
        });
        std::fputs("\nSCALARS density double\nLOOKUP_TABLE default\n", f);
        textio::parallel_emit(f, n, textio::FIXED8_MAX + 16, [&](size_t k, char* o) {
            int c = textio::fixed8(rho[k], o);
            o[c] = '\n';
            return c + 1;
        });
        std::fclose(f);
    }

    // Extension: the same frame as legacy-VTK BINARY (big-endian doubles), 42 MB instead of 51 MB at
    // 2048 x 512 and no decimal formatting; ParaView reads both.  Chosen by params.vtk_binary.
    static void write_vtk_arrays_binary(const double* ux, const double* uy, const double* rho, int nx, int ny, int timestep) {
        char filename[256];
        std::snprintf(filename, sizeof(filename), "vtk_output/lbm_%06d.vtk", timestep);
        std::FILE* f = std::fopen(filename, "wb");
        if (!f) {
            std::cerr << "ERROR: Cannot write " << filename << "\n";
            return;
        }
        std::fprintf(f, "# vtk DataFile Version 3.0\nLBM Flow Timestep %d\nBINARY\nDATASET STRUCTURED_POINTS\n", timestep);
        std::fprintf(f, "DIMENSIONS %d %d 1\nORIGIN 0 0 0\nSPACING 1 1 1\nPOINT_DATA %d\n", nx, ny, nx * ny);
        const size_t n = (size_t)nx * ny;
        auto big_endian = [](double v) {
            uint64_t b;
            std::memcpy(&b, &v, 8);
            return __builtin_bswap64(b);
        };
        std::vector<uint64_t> buf(3 * n);
        std::fputs("VECTORS velocity double\n", f);
        for (size_t k = 0; k < n; ++k) {
            buf[3 * k] = big_endian(ux[k]);
            buf[3 * k + 1] = big_endian(uy[k]);
            buf[3 * k + 2] = big_endian(0.0);
        }
        std::fwrite(buf.data(), 8, 3 * n, f);
        std::fputs("\nSCALARS velocity_magnitude double\nLOOKUP_TABLE default\n", f);
        for (size_t k = 0; k < n; ++k) buf[k] = big_endian(std::sqrt(ux[k] * ux[k] + uy[k] * uy[k]));
        std::fwrite(buf.data(), 8, n, f);
        std::fputs("\nSCALARS density double\nLOOKUP_TABLE default\n", f);
        for (size_t k = 0; k < n; ++k) buf[k] = big_endian(rho[k]);
        std::fwrite(buf.data(), 8, n, f);
        std::fputs("\n", f);
        std::fclose(f);
    }

    // velocity_field.csv, simulation_params.csv and the force-coefficient summary (reference
    // :194-219).  forces.csv is flushed first so that every row takes part in the averages (the
    // reference re-reads the file while its stream is still buffered, :367-372).
    void write_final_results(const Grid& grid, const SimulationParams& params) {
        if (mpi_rank_ == 0) std::cout << "\nGathering final results..." << std::endl;
        std::vector<double> g_rho, g_ux, g_uy;
        if (mpi_rank_ == 0) {
            const size_t n = (size_t)grid.global_nx() * grid.global_ny();
            g_rho.resize(n);
            g_ux.resize(n);
            g_uy.resize(n);
        }
        grid.check(lbm_gather_macros(grid.handle(), g_rho.data(), g_ux.data(), g_uy.data()));
        if (mpi_rank_ == 0) {
            write_velocity_field(g_ux, g_uy, g_rho, params);
            write_simulation_params(g_ux, g_uy, params);
            if (force_file_) std::fflush(force_file_);
            calculate_time_averaged_drag();
            std::cout << "Files written: velocity_field.csv, simulation_params.csv, forces.csv" << std::endl;
        }
        double token = 0.0;  // MPI_Barrier of :218
        grid.check(lbm_allreduce(grid.handle(), &token, 1, LBM_SUM));
    }

    int rank() const { return mpi_rank_; }

   private:
    static void write_velocity_field(const std::vector<double>& ux_g, const std::vector<double>& uy_g,
                                     const std::vector<double>& rho_g, const SimulationParams& p) {
        std::FILE* f = std::fopen("velocity_field.csv", "w");
        if (!f) {
            std::cerr << "ERROR: Cannot write velocity_field.csv\n";
            return;
        }
        std::vector<char> big(1 << 22);
        std::setvbuf(f, big.data(), _IOFBF, big.size());
        std::fputs("x,y,ux,uy,rho,velocity_magnitude\n", f);
        const int nx = p.nx;
        textio::parallel_emit(f, (size_t)p.nx * p.ny, 4 * textio::FIXED8_MAX + 64, [&](size_t k, char* o) {
            char* q = o;
            q += textio::integer((long long)(k % nx), q);
            *q++ = ',';
            q += textio::integer((long long)(k / nx), q);
            const double mag = std::sqrt(ux_g[k] * ux_g[k] + uy_g[k] * uy_g[k]);
            for (double v : {ux_g[k], uy_g[k], rho_g[k], mag}) {
                *q++ = ',';
                q += textio::fixed8(v, q);
            }
            *q++ = '\n';
            return (int)(q - o);
        });
        std::fclose(f);
        std::cout << "  velocity_field.csv written\n";
    }

    static void write_simulation_params(const std::vector<double>& ux_g, const std::vector<double>& uy_g,
                                        const SimulationParams& p) {
        std::ofstream file("simulation_params.csv");
        if (!file) {
            std::cerr << "ERROR: Cannot write simulation_params.csv\n";
            return;
        }
        double max_vel = 0.0, avg_vel = 0.0;  // serial, row-major: the reference's summation order (:339-346)
        const size_t n = (size_t)p.nx * p.ny;
        for (size_t k = 0; k < n; ++k) {
            const double vel = std::sqrt(ux_g[k] * ux_g[k] + uy_g[k] * uy_g[k]);
            max_vel = std::max(max_vel, vel);
            avg_vel += vel;
        }
        avg_vel /= (p.nx * p.ny);
        file << "parameter,value\n"
             << "nx," << p.nx << "\n"
             << "ny," << p.ny << "\n"
             << "tau," << std::fixed << std::setprecision(8) << p.tau << "\n"
             << "nu," << p.nu() << "\n"
             << "inlet_velocity," << p.inlet_velocity << "\n"
             << "num_timesteps," << p.num_timesteps << "\n"
             << "reynolds_number," << p.reynolds() << "\n"
             << "cylinder_x," << p.get_cylinder_x() << "\n"
             << "cylinder_y," << p.get_cylinder_y() << "\n"
             << "cylinder_radius," << p.get_cylinder_radius_cells() << "\n"
             << "max_velocity," << max_vel << "\n"
             << "avg_velocity," << avg_vel << "\n";
        std::cout << "  simulation_params.csv written\n";
    }

    // Mean and range of C_D / C_L over the rows with timestep > 1000 (reference :367-413).
    static void calculate_time_averaged_drag() {
        std::ifstream in("forces.csv");
        if (!in) {
            std::cerr << "Warning: Could not read forces.csv for averaging\n";
            return;
        }
        std::string line;
        std::getline(in, line);  // header
        double sum_cd = 0.0, sum_cl = 0.0, max_cd = -1e9, min_cd = 1e9, max_cl = -1e9, min_cl = 1e9;
        int count = 0;
        while (std::getline(in, line)) {
            int t;
            double fx, fy, cd, cl;
            if (std::sscanf(line.c_str(), "%d,%lf,%lf,%lf,%lf", &t, &fx, &fy, &cd, &cl) != 5 || t <= 1000) continue;
            sum_cd += cd;
            sum_cl += cl;
            max_cd = std::max(max_cd, cd);
            min_cd = std::min(min_cd, cd);
            max_cl = std::max(max_cl, cl);
            min_cl = std::min(min_cl, cl);
            ++count;
        }
        if (count == 0) return;
        std::cout << "\n=== Time-Averaged Force Coefficients ===\n";
        std::cout << "  Mean C_D = " << std::fixed << std::setprecision(6) << sum_cd / count << "\n";
        std::cout << "  C_D range: [" << min_cd << ", " << max_cd << "]\n";
        std::cout << "  Mean C_L = " << sum_cl / count << "\n";
        std::cout << "  C_L range: [" << min_cl << ", " << max_cl << "]\n";
        std::cout << "  (Averaged over " << count << " samples)\n";
    }

    std::FILE* force_file_ = nullptr;
    int mpi_rank_ = 0;
};

// Asynchronous VTK output (replaces the blocking gather + write of the reference's
// Solver::write_vtk_frame, include/LBMSolver.h:269-362).  Two pinned host images of rho/ux/uy;
// frame k is copied into image k%2 on the engine's copy stream while the compute stream moves
// on, and a writer thread formats it once its copy has landed.  submit() blocks only when both
// images are still being written (the ASCII format, not the GPU, is then the bottleneck).
// One image of the whole channel in POSIX shared memory, page-locked in every slab process: each GPU copies its own
// columns into it (lbm_snapshot_begin_slot2d), nobody gathers anything.  Two slots; a small header of lock-free
// counters says which frame each slab has delivered into a slot and which frames rank 0 has written.
class SharedFrameImage {
   public:
    static constexpr int kMaxSlabs = 64;
    struct Header {
        std::atomic<int64_t> magic;
        std::atomic<int64_t> attached;               // slab processes that have mapped the segment
        std::atomic<int64_t> written;                // frames rank 0 has finished writing
        std::atomic<int64_t> ready[2][kMaxSlabs];    // ready[slot][r] = k+1: slab r's part of frame k is in the slot
    };

    SharedFrameImage(const Grid& grid) : grid_(grid) {
        n_ = (size_t)grid.global_nx() * grid.global_ny();
        bytes_ = sizeof(Header) + 2 * 3 * n_ * sizeof(double);
        bytes_ = (bytes_ + 4095) / 4096 * 4096;
        char name[96];
        std::snprintf(name, sizeof(name), "/lbm_b200_frames_%016llx", grid.job_tag());
        name_ = name;
        int fd = -1;
        if (grid.mpi_rank() == 0) {
            shm_unlink(name_.c_str());
            fd = shm_open(name_.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
            if (fd < 0 || ftruncate(fd, (off_t)bytes_) != 0) throw std::runtime_error("lbm_b200: cannot create shared frame image " + name_);
        } else {
            for (int tries = 0; tries < 12000 && fd < 0; ++tries) {  // up to 60 s
                fd = shm_open(name_.c_str(), O_RDWR, 0600);
                struct stat st;
                if (fd >= 0 && (fstat(fd, &st) != 0 || (size_t)st.st_size < bytes_)) {
                    close(fd);
                    fd = -1;
                }
                if (fd < 0) usleep(5000);
            }
            if (fd < 0) throw std::runtime_error("lbm_b200: timed out waiting for shared frame image " + name_);
        }
        base_ = mmap(nullptr, bytes_, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (base_ == MAP_FAILED) throw std::runtime_error("lbm_b200: mmap of the shared frame image failed");
        hdr_ = static_cast<Header*>(base_);
        if (grid.mpi_rank() == 0) {
            hdr_->attached.store(0);
            hdr_->written.store(0);
            for (auto& slot : hdr_->ready)
                for (auto& r : slot) r.store(0);
            hdr_->magic.store(kMagic, std::memory_order_release);
        } else {
            for (int tries = 0; hdr_->magic.load(std::memory_order_acquire) != kMagic; ++tries) {
                if (tries > 12000) throw std::runtime_error("lbm_b200: shared frame image never initialised");
                usleep(5000);
            }
        }
        registered_ = lbm_host_register(base_, bytes_) == LBM_OK;  // pageable memory still works, only synchronously
        hdr_->attached.fetch_add(1);
    }
    ~SharedFrameImage() {
        if (!base_ || base_ == MAP_FAILED) return;
        if (grid_.mpi_rank() == 0) {
            for (int tries = 0; hdr_->attached.load() < grid_.mpi_size() && tries < 2000; ++tries) usleep(5000);
            shm_unlink(name_.c_str());
        }
        if (registered_) lbm_host_unregister(base_);
        munmap(base_, bytes_);
    }
    Header* header() { return hdr_; }
    double* field(int slot, int k) { return reinterpret_cast<double*>(static_cast<char*>(base_) + sizeof(Header)) + ((size_t)slot * 3 + k) * n_; }
    size_t cells() const { return n_; }

   private:
    static constexpr int64_t kMagic = 0x4c424d4652414d45LL;
    const Grid& grid_;
    std::string name_;
    void* base_ = nullptr;
    Header* hdr_ = nullptr;
    size_t n_ = 0, bytes_ = 0;
    bool registered_ = false;
};

class FrameWriter {
   public:
    explicit FrameWriter(const Grid& grid, bool binary = false) : grid_(grid), binary_(binary) {}
    ~FrameWriter() { finish(); }

    // Returns as soon as the device-side snapshot is queued: the D2H copies run on the engine's copy stream, the
    // formatting and the file write on a host thread, while the caller keeps stepping.  Multi-slab jobs: every slab
    // process calls this; its GPU fills its own columns of the shared image and rank 0's thread writes the file when
    // all parts have arrived -- no collective, no gather through rank 0's GPU.
    void submit(int timestep) {
        const int64_t k = (int64_t)submitted_;
        const int slot = (int)(k % 2);
        const bool root = grid_.mpi_rank() == 0;
        const bool multi = grid_.mpi_size() > 1;
        if (multi) {
            if (grid_.mpi_size() > SharedFrameImage::kMaxSlabs) throw std::runtime_error("lbm_b200: too many slabs for the shared frame image");
            if (!shared_) shared_ = std::make_unique<SharedFrameImage>(grid_);
            // the slot is free once rank 0 has written frame k-2 (back-pressure only when the writer is two frames behind)
            while (shared_->header()->written.load(std::memory_order_acquire) < k - 1) usleep(200);
            const size_t pitch = (size_t)grid_.global_nx();
            grid_.check(lbm_snapshot_begin_slot2d(grid_.handle(), slot, shared_->field(slot, 0) + grid_.x_start(),
                                                  shared_->field(slot, 1) + grid_.x_start(), shared_->field(slot, 2) + grid_.x_start(), pitch));
        } else {
            const size_t n = (size_t)grid_.global_nx() * grid_.global_ny();
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return !busy_[slot]; });
            }
            if (!image_[slot]) {
                void* p = nullptr;
                grid_.check(lbm_host_alloc(&p, 3 * n * sizeof(double)));
                image_[slot] = static_cast<double*>(p);
            }
            grid_.check(lbm_snapshot_begin_slot(grid_.handle(), slot, image_[slot], image_[slot] + n, image_[slot] + 2 * n));
        }
        ++submitted_;
        {
            std::lock_guard<std::mutex> lk(m_);
            busy_[slot] = true;
            jobs_.push_back({timestep, slot, k});
        }
        if (!thread_.joinable()) thread_ = std::thread([this] { loop(); });
        cv_.notify_all();
        (void)root;
    }

    // Blocks until every submitted frame is on disk (rank 0) / delivered (other slabs).
    void finish() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        cv_.notify_all();
        if (thread_.joinable()) thread_.join();
        stop_ = false;
        if (shared_) {
            // nobody unmaps before rank 0 has written the last frame out of the shared image
            while (shared_->header()->written.load(std::memory_order_acquire) < (int64_t)submitted_) usleep(500);
            shared_.reset();
        }
        for (double*& p : image_)
            if (p) {
                lbm_host_free(p);
                p = nullptr;
            }
    }

    size_t frames_written() const { return written_; }

   private:
    struct Job {
        int timestep, slot;
        int64_t index;
    };
    void loop() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return stop_ || !jobs_.empty(); });
                if (jobs_.empty()) return;
                j = jobs_.front();
                jobs_.pop_front();
            }
            lbm_snapshot_wait_slot(grid_.handle(), j.slot);  // this slab's D2H copies of the frame have landed
            const size_t n = (size_t)grid_.global_nx() * grid_.global_ny();
            const double *rho, *ux, *uy;
            bool write = true;
            if (shared_) {
                auto* hdr = shared_->header();
                hdr->ready[j.slot][grid_.mpi_rank()].store(j.index + 1, std::memory_order_release);
                write = grid_.mpi_rank() == 0;
                if (write)
                    for (int r = 0; r < grid_.mpi_size(); ++r)
                        while (hdr->ready[j.slot][r].load(std::memory_order_acquire) < j.index + 1) usleep(100);
                rho = shared_->field(j.slot, 0);
                ux = shared_->field(j.slot, 1);
                uy = shared_->field(j.slot, 2);
            } else {
                rho = image_[j.slot];
                ux = image_[j.slot] + n;
                uy = image_[j.slot] + 2 * n;
            }
            if (write) {
                if (binary_)
                    IOManager::write_vtk_arrays_binary(ux, uy, rho, grid_.global_nx(), grid_.global_ny(), j.timestep);
                else
                    IOManager::write_vtk_arrays(ux, uy, rho, grid_.global_nx(), grid_.global_ny(), j.timestep);
                if (shared_) shared_->header()->written.store(j.index + 1, std::memory_order_release);
            }
            {
                std::lock_guard<std::mutex> lk(m_);
                busy_[j.slot] = false;
                ++written_;
            }
            cv_.notify_all();
        }
    }

    const Grid& grid_;
    bool binary_ = false;
    double* image_[2] = {nullptr, nullptr};
    std::unique_ptr<SharedFrameImage> shared_;
    bool busy_[2] = {false, false};
    std::deque<Job> jobs_;
    std::mutex m_;
    std::condition_variable cv_;
    std::thread thread_;
    bool stop_ = false;
    size_t submitted_ = 0, written_ = 0;
};

}  // namespace LBM
