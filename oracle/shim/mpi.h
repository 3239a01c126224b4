/* oracle/shim/mpi.h -- TEST INFRASTRUCTURE ONLY.
 * Minimal stand-in for <mpi.h> so the reference's headers compile in an image without MPI.
 * The reference includes <mpi.h> unconditionally (include/qlten/framework/hp_numeric/mpi_fun.h:21)
 * but nothing on the Contract path communicates. Every communication stub aborts. */
#ifndef QLB200_ORACLE_SHIM_MPI_H
#define QLB200_ORACLE_SHIM_MPI_H
#include <cstdio>
#include <cstdlib>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
struct MPI_Status { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; };
#define MPI_SUCCESS 0
#define MPI_MAX_ERROR_STRING 256
#define MPI_ANY_SOURCE (-1)
#define MPI_ANY_TAG (-1)
#define MPI_COMM_WORLD 0
#define MPI_THREAD_MULTIPLE 3
#define MPI_IN_PLACE ((void *)1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
enum { MPI_CHAR = 1, MPI_INT, MPI_FLOAT, MPI_DOUBLE, MPI_UNSIGNED, MPI_UNSIGNED_LONG,
       MPI_UNSIGNED_LONG_LONG, MPI_CXX_BOOL, MPI_CXX_FLOAT_COMPLEX, MPI_CXX_DOUBLE_COMPLEX };
enum { MPI_SUM = 1, MPI_MIN, MPI_MAX };
#define QLB200_MPI_ABORT(name) do { std::fprintf(stderr, "oracle shim: %s called (no MPI in this image)\n", name); std::abort(); } while (0)
inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return MPI_SUCCESS; }
inline int MPI_Comm_size(MPI_Comm, int *s) { *s = 1; return MPI_SUCCESS; }
inline int MPI_Error_string(int, char *s, int *len) { s[0] = 0; *len = 0; return MPI_SUCCESS; }
inline int MPI_Init_thread(int *, char ***, int, int *provided) { *provided = MPI_THREAD_MULTIPLE; return MPI_SUCCESS; }
inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { QLB200_MPI_ABORT("MPI_Send"); return 1; }
inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { QLB200_MPI_ABORT("MPI_Recv"); return 1; }
inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { QLB200_MPI_ABORT("MPI_Bcast"); return 1; }
inline int MPI_Gather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm) { QLB200_MPI_ABORT("MPI_Gather"); return 1; }
inline int MPI_Allgather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm) { QLB200_MPI_ABORT("MPI_Allgather"); return 1; }
inline int MPI_Allgatherv(const void *, int, MPI_Datatype, void *, const int *, const int *, MPI_Datatype, MPI_Comm) { QLB200_MPI_ABORT("MPI_Allgatherv"); return 1; }
inline int MPI_Reduce(const void *, void *, int, MPI_Datatype, MPI_Op, int, MPI_Comm) { QLB200_MPI_ABORT("MPI_Reduce"); return 1; }
inline int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm) { QLB200_MPI_ABORT("MPI_Allreduce"); return 1; }
#endif
