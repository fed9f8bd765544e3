/*
 * oracle/shim/mpi.h -- single-rank stand-in for <mpi.h>.  TEST INFRASTRUCTURE ONLY.
 *
 * The reference (LGMOak/HighPerformanceComputing-LatticeBoltzmannMethod) needs an MPI
 * installation; this image has none.  This header gives the 15 MPI entry points the
 * reference calls the semantics they have in a 1-rank job, so that the reference headers
 * compile and run UNMODIFIED from /root/reference/include (see oracle/Makefile).
 *
 * Call sites covered (reference file:line):
 *   MPI_Init / MPI_Finalize / MPI::Is_initialized      src/main.cpp:8,30,37,41
 *   MPI_Comm_rank / MPI_Comm_size                      include/LBMGrid.h:58-59,353  LBMIO.h:36,230-231
 *   MPI_Cart_create / _coords / _shift                 include/LBMGrid.h:352-363
 *   MPI_Isend / MPI_Irecv / MPI_Waitall                include/LBMGrid.h:255-280
 *   MPI_Reduce / MPI_Allreduce                         include/LBMGrid.h:175,315,342  LBMIO.h:167-168
 *   MPI_Gather / MPI_Gatherv / MPI_Barrier             include/LBMSolver.h:289-337  LBMIO.h:218,237-283
 *
 * Semantics that matter for parity: with one rank every Cart_shift neighbour
 * is MPI_PROC_NULL, so Isend/Irecv are no-ops and -- crucially -- Irecv leaves the receive
 * buffer untouched.  The reference then unpacks its zero-initialised E/W receive buffers into
 * the ghost columns every step; that behaviour is part of the oracle.
 *
 * The reference also relies on <mpi.h> dragging in <cstdio>/<string> (printf at
 * LBMGrid.h:93, sscanf/getline at LBMIO.h:374-388), hence the includes below.
 */
#ifndef ORACLE_SHIM_MPI_H
#define ORACLE_SHIM_MPI_H

#include <cstdio>
#include <cstring>
#include <string>

typedef int MPI_Comm;
typedef int MPI_Request;
typedef int MPI_Op;
typedef int MPI_Datatype; /* the value IS the element size in bytes */
struct MPI_Status { int unused; };

#define MPI_COMM_WORLD 0
#define MPI_PROC_NULL (-2)
#define MPI_STATUSES_IGNORE (static_cast<MPI_Status*>(nullptr))
#define MPI_SUCCESS 0

#define MPI_BYTE 1
#define MPI_INT 4
#define MPI_DOUBLE 8

#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3

namespace oracle_shim {
inline bool& initialised_flag() {
    static bool f = false;
    return f;
}
inline void copy_elems(const void* src, void* dst, int count, MPI_Datatype size) {
    if (src != dst && count > 0) std::memcpy(dst, src, static_cast<size_t>(count) * static_cast<size_t>(size));
}
}  // namespace oracle_shim

inline int MPI_Init(int*, char***) { oracle_shim::initialised_flag() = true; return MPI_SUCCESS; }
inline int MPI_Finalize() { oracle_shim::initialised_flag() = false; return MPI_SUCCESS; }
inline int MPI_Comm_rank(MPI_Comm, int* rank) { *rank = 0; return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int* size) { *size = 1; return MPI_SUCCESS; }

inline int MPI_Cart_create(MPI_Comm comm, int, const int*, const int*, int, MPI_Comm* out) {
    *out = comm;
    return MPI_SUCCESS;
}
inline int MPI_Cart_coords(MPI_Comm, int, int ndims, int* coords) {
    for (int d = 0; d < ndims; ++d) coords[d] = 0;
    return MPI_SUCCESS;
}
inline int MPI_Cart_shift(MPI_Comm, int, int, int* src, int* dst) {
    *src = MPI_PROC_NULL;
    *dst = MPI_PROC_NULL;
    return MPI_SUCCESS;
}

/* Point-to-point with MPI_PROC_NULL completes immediately and moves nothing. */
inline int MPI_Isend(const void*, size_t, MPI_Datatype, int, int, MPI_Comm, MPI_Request* r) { *r = 0; return MPI_SUCCESS; }
inline int MPI_Irecv(void*, size_t, MPI_Datatype, int, int, MPI_Comm, MPI_Request* r) { *r = 0; return MPI_SUCCESS; }
inline int MPI_Waitall(int, MPI_Request*, MPI_Status*) { return MPI_SUCCESS; }
inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }

/* Every reduction / gather over one rank is the identity. */
inline int MPI_Reduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm) {
    oracle_shim::copy_elems(s, r, n, t);
    return MPI_SUCCESS;
}
inline int MPI_Allreduce(const void* s, void* r, int n, MPI_Datatype t, MPI_Op, MPI_Comm) {
    oracle_shim::copy_elems(s, r, n, t);
    return MPI_SUCCESS;
}
inline int MPI_Gather(const void* s, int n, MPI_Datatype t, void* r, int, MPI_Datatype, int, MPI_Comm) {
    oracle_shim::copy_elems(s, r, n, t);
    return MPI_SUCCESS;
}
inline int MPI_Gatherv(const void* s, int n, MPI_Datatype t, void* r, const int*, const int* displs, MPI_Datatype,
                       int, MPI_Comm) {
    char* base = static_cast<char*>(r);
    if (base) oracle_shim::copy_elems(s, base + static_cast<size_t>(displs ? displs[0] : 0) * t, n, t);
    return MPI_SUCCESS;
}

namespace MPI {
inline bool Is_initialized() { return oracle_shim::initialised_flag(); }
}  // namespace MPI

#endif /* ORACLE_SHIM_MPI_H */
