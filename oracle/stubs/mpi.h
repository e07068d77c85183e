/* TEST INFRASTRUCTURE ONLY (oracle/).
 *
 * Minimal single-process stand-in for <mpi.h> so that the reference's
 * code/mpi_funcs.hpp (everything inside `#ifdef USE_MPI`) can be compiled in a
 * container that has no MPI.  Only the pure, per-rank functions of that header are
 * ever CALLED by oracle/ref_driver.cpp (seg_work_sharing_arr, seg_mtx_struct,
 * localize_row_idx, collect_local_needed_heri); the communication routines merely
 * have to parse.  Every call here is a no-op that reports success.
 */
#ifndef USPMV_ORACLE_STUB_MPI_H
#define USPMV_ORACLE_STUB_MPI_H

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef long MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_REQUEST_NULL 0
#define MPI_INT 1
#define MPI_LONG 2
#define MPI_DOUBLE 3
#define MPI_FLOAT 4
#define MPI_CHAR 5
#define MPI_C_BOOL 6
#define MPI_CXX_BOOL 6
#define MPI_UNSIGNED_LONG 7
#define MPI_SHORT 8
#define MPI_BYTE 9
#define MPI_UNSIGNED 10
#define MPI_LONG_LONG 11
#define MPI_C_FLOAT_COMPLEX 12
#define MPI_C_DOUBLE_COMPLEX 13
#define MPI_UNSIGNED_CHAR 14
#define MPI_LONG_DOUBLE 15
#define MPI_UNSIGNED_SHORT 16
#define MPI_LONG_LONG_INT 11
#define MPI_UNSIGNED_LONG_LONG 17
#define MPI_INT64_T 18
#define MPI_UINT64_T 19
#define MPI_DATATYPE_NULL 0
#define MPI_SUM 1
#define MPI_MAX 2
#define MPI_MIN 3
#define MPI_IN_PLACE ((void *)1)

#ifdef __cplusplus
#define USPMV_STUB_INLINE static inline
#else
#define USPMV_STUB_INLINE static inline
#endif

USPMV_STUB_INLINE int MPI_Init(int *, char ***) { return 0; }
USPMV_STUB_INLINE int MPI_Finalize(void) { return 0; }
USPMV_STUB_INLINE int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
USPMV_STUB_INLINE int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return 0; }
USPMV_STUB_INLINE int MPI_Barrier(MPI_Comm) { return 0; }
USPMV_STUB_INLINE double MPI_Wtime(void) { return 0.0; }
USPMV_STUB_INLINE int MPI_Abort(MPI_Comm, int) { return 0; }
USPMV_STUB_INLINE int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { return 0; }
USPMV_STUB_INLINE int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { return 0; }
USPMV_STUB_INLINE int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return 0; }
USPMV_STUB_INLINE int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *) { return 0; }
USPMV_STUB_INLINE int MPI_Wait(MPI_Request *, MPI_Status *) { return 0; }
USPMV_STUB_INLINE int MPI_Waitall(int, MPI_Request *, MPI_Status *) { return 0; }
USPMV_STUB_INLINE int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
USPMV_STUB_INLINE int MPI_Gather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
USPMV_STUB_INLINE int MPI_Gatherv(const void *, int, MPI_Datatype, void *, const int *, const int *, MPI_Datatype, int, MPI_Comm) { return 0; }
USPMV_STUB_INLINE int MPI_Allgather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm) { return 0; }
USPMV_STUB_INLINE int MPI_Allreduce(const void *, void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
USPMV_STUB_INLINE int MPI_Reduce(const void *, void *, int, MPI_Datatype, int, int, MPI_Comm) { return 0; }
USPMV_STUB_INLINE MPI_Aint MPI_Aint_diff(MPI_Aint a, MPI_Aint b) { return a - b; }
USPMV_STUB_INLINE int MPI_Get_address(const void *, MPI_Aint *a) { *a = 0; return 0; }
USPMV_STUB_INLINE int MPI_Type_create_struct(int, const int *, const MPI_Aint *, const MPI_Datatype *, MPI_Datatype *) { return 0; }
USPMV_STUB_INLINE int MPI_Type_create_struct(int, int *, MPI_Aint *, MPI_Datatype *, MPI_Datatype *) { return 0; }
USPMV_STUB_INLINE int MPI_Type_vector(int, int, int, MPI_Datatype, MPI_Datatype *) { return 0; }
USPMV_STUB_INLINE int MPI_Type_contiguous(int, MPI_Datatype, MPI_Datatype *) { return 0; }
USPMV_STUB_INLINE int MPI_Type_indexed(int, const int *, const int *, MPI_Datatype, MPI_Datatype *) { return 0; }
USPMV_STUB_INLINE int MPI_Type_create_hvector(int, int, MPI_Aint, MPI_Datatype, MPI_Datatype *) { return 0; }
USPMV_STUB_INLINE int MPI_Type_create_resized(MPI_Datatype, MPI_Aint, MPI_Aint, MPI_Datatype *) { return 0; }
USPMV_STUB_INLINE int MPI_Type_commit(MPI_Datatype *) { return 0; }
USPMV_STUB_INLINE int MPI_Type_free(MPI_Datatype *) { return 0; }
USPMV_STUB_INLINE int MPI_Type_size(MPI_Datatype, int *s) { *s = 0; return 0; }

#endif
