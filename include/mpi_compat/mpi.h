// mpi_compat/mpi.h -- lets the reference's src/main.cpp compile UNCHANGED against these headers
// on a machine without MPI: add -Iinclude/mpi_compat.  Only the five MPI names main.cpp uses
// (src/main.cpp:5,8,30,32,37,41) exist here.  Ranks are the GPU slabs of lbm_bootstrap_env
// (one process per GPU, started by torchrun / mpirun / srun as a plain launcher); all data
// movement between slabs is NCCL inside liblbm_b200.so, none of it goes through this file.
#pragma once

#include <cstdio>

#include "lbm_b200.h"

typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0

namespace lbm_mpi_compat {
inline bool& initialised() {
    static bool flag = false;
    return flag;
}
}  // namespace lbm_mpi_compat

inline int MPI_Init(int*, char***) {
    lbm_mpi_compat::initialised() = true;
    return MPI_SUCCESS;
}
inline int MPI_Finalize() {
    lbm_mpi_compat::initialised() = false;
    return MPI_SUCCESS;
}
inline int MPI_Comm_rank(MPI_Comm, int* rank) {
    int world = 1;
    return lbm_bootstrap_env(rank, &world, nullptr, nullptr) == LBM_OK ? MPI_SUCCESS : 1;
}
inline int MPI_Comm_size(MPI_Comm, int* size) {
    int rank = 0;
    return lbm_bootstrap_env(&rank, size, nullptr, nullptr) == LBM_OK ? MPI_SUCCESS : 1;
}
namespace MPI {
inline bool Is_initialized() { return lbm_mpi_compat::initialised(); }
}  // namespace MPI
