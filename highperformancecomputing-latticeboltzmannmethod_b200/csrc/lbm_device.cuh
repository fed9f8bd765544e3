// lbm_device.cuh -- device-side helpers shared by the kernels of lbm_kernels.cu and lbm_tb.cu: cell
// loads / stores in the SoA layout, the reference's boundary rules in its serial order, the f_current
// observer, and the peer-memory hand-shake of the fused halo exchange.
#pragma once
#include <cuda_runtime.h>

#include "lbm_cell.cuh"
#include "lbm_kernels.cuh"
#include "lbm_layout.h"

namespace lbm {

// COHERENT = true: ld.global.cg (L2, never the non-coherent path) -- for cells whose pulls reach into ghost
// columns that a neighbouring GPU stores into while this kernel is resident (the block spins on a flag
// first and loads afterwards: ld.global.nc would only be defined for data that is read-only for the whole
// kernel).  Interior cells keep the non-coherent path.
template <bool PULL, bool COHERENT = false>
__device__ __forceinline__ void load_cell(const double* __restrict__ src, const Layout& L, int gx, int y,
                                          double f[Q]) {
#pragma unroll
    for (int i = 0; i < Q; ++i) {
        const int dx = PULL ? cxi(i) : 0, dy = PULL ? cyi(i) : 0;
        const double* p = src + i * L.plane + L.at(gx - dx, y - dy);
        f[i] = COHERENT ? __ldcg(p) : __ldg(p);
    }
}

__device__ __forceinline__ void store_cell(double* __restrict__ dst, const Layout& L, int gx, int y,
                                           const double f[Q]) {
#pragma unroll
    for (int i = 0; i < Q; ++i) dst[i * L.plane + L.at(gx, y)] = f[i];
}

template <bool FORCED>
__device__ __forceinline__ void collide_cell(double f[Q], double tau_inv, double Fx, double Fy) {
    const Moments m = moments(f);
    if (FORCED)
        bgk_forced(f, m, tau_inv, Fx, Fy, f);
    else
        bgk(f, m, tau_inv, f);
}

__device__ __forceinline__ bool any_unstable(const double f[Q]) {
    bool bad = false;
#pragma unroll
    for (int i = 0; i < Q; ++i) bad |= unstable_value(f[i]);
    return bad;
}

// The reference's boundary rules in its serial order (include/LBMSolver.h:153-236):
// bottom row, top row, inlet column, outlet column.  x, y are slab-interior coordinates.
__device__ __forceinline__ void apply_bc(double f[Q], int x, int y, const Layout& L, const BcArgs& b,
                                         double& rho_bc, double& u_out) {
    if (b.walls && y == 0) wall_bottom(f);
    if (b.walls && y == L.ny - 1) wall_top(f);
    if (b.inlet && x == 0) rho_bc = zou_he_inlet(f, b.u_in);
    if (b.outlet && x == L.lnx - 1) u_out = zou_he_outlet(f);
}

// Observables.  f_current of the reference after its last iteration, for one interior cell,
// given the newest post-collision buffer: pull, then boundary rules (fluid) or reversal (solid)
// (include/LBMSolver.h:128-145, 147-265).
__device__ __forceinline__ void current_from_next(const double* __restrict__ fnext, const Layout& L,
                                                  const unsigned char* __restrict__ mask, const BcArgs& b, int x,
                                                  int y, double f[Q], double& rho_bc, double& u_out) {
    load_cell<true>(fnext, L, x + 1, y, f);
    if (mask[L.at(x + 1, y)])
        reverse(f);
    else
        apply_bc(f, x, y, L, b, rho_bc, u_out);
}

// ---- peer-memory halo hand-shake (P2pArgs in lbm_kernels.cuh) --------------------------------------
// Protocol, per launch `seq`:
//   1. wait until both neighbours have delivered exchange seq-1 (their counters in MY memory).  That
//      also means they have finished reading the ghost columns this launch is about to overwrite in
//      THEIR memory (the A-B pair alternates, so the ghost written now was read by their launch seq-1);
//   2. compute the edge cells, store them locally and push the face-crossing populations into the
//      neighbours' ghost columns (plain stores to peer memory: NVLink);
//   3. __threadfence_system(), and the last block to finish publishes `seq` in the neighbours' memory.
// No rank can run more than one exchange ahead of a neighbour, there is no cycle in the waits (launch
// seq of one GPU only waits for launch seq-1 of another), and a spinning block never keeps another
// GPU from making progress.
//
// The wait is BOUNDED: a neighbour that died, failed a launch or left the sequence would otherwise hang
// every other GPU inside a kernel for ever.  After `timeout_ns` (or as soon as the slab's status word is
// non-zero: an earlier time-out, or the host raising it) the block gives up, records LBM_HALO_TIMEOUT in
// the status word and carries on with whatever the ghost columns hold; the host finds the status word at
// its next synchronisation point and every entry point returns LBM_ERR_NCCL from then on.
constexpr int LBM_HALO_TIMEOUT = 1;

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void spin_until(const int* flag, int want, const P2pArgs& x) {
    const volatile int* f = reinterpret_cast<const volatile int*>(flag);
    const volatile int* st = reinterpret_cast<const volatile int*>(x.status);
    if (*f >= want) return;
    if (st && *st) return;
    const unsigned long long t0 = global_ns();
    unsigned int it = 0;
    while (*f < want) {
        __nanosleep(64);
        if ((++it & 255u) == 0) {
            if (st && *st) return;
            if (x.timeout_ns && global_ns() - t0 > x.timeout_ns) {
                if (x.status) atomicExch(x.status, LBM_HALO_TIMEOUT);
                return;
            }
        }
    }
}

// Steps 1 and 3 of the protocol for one block of edge work.
__device__ __forceinline__ void p2p_block_begin(const P2pArgs& x) {
    if (threadIdx.x == 0) {
        if (x.peer_dst_west) spin_until(x.my_flags + 0, x.seq - 1, x);
        if (x.peer_dst_east) spin_until(x.my_flags + 1, x.seq - 1, x);
        __threadfence_system();
    }
    __syncthreads();
}

__device__ __forceinline__ void p2p_block_end(const P2pArgs& x, unsigned int edge_blocks) {
    __threadfence_system();  // this block's peer stores first
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(x.blocks_done, 1u) == edge_blocks - 1) {
            *x.blocks_done = 0;
            __threadfence_system();
            if (x.west_flag) *reinterpret_cast<volatile int*>(x.west_flag) = x.seq;
            if (x.east_flag) *reinterpret_cast<volatile int*>(x.east_flag) = x.seq;
            __threadfence_system();
        }
    }
}

}  // namespace lbm
