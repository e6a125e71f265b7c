// Single-rank stand-in for <mpi.h> (absent here).  TEST INFRASTRUCTURE of oracle/harness: the collectives the
// plugin classes call, for one process.
#pragma once
#include <string.h>
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct { int a; } MPI_Status;
typedef void *MPI_File;
typedef int MPI_Info;
typedef int MPI_Op;
#define MPI_BYTE 1
#define MPI_DOUBLE 8
#define MPI_SUCCESS 0
#define MPI_SUM 1
#define MPI_COMM_WORLD 0
static inline int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op, MPI_Comm) {
  if (sendbuf != recvbuf) memcpy(recvbuf, sendbuf, (size_t)count * (size_t)type);
  return MPI_SUCCESS;
}
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return MPI_SUCCESS; }
static inline int MPI_Barrier(MPI_Comm) { return MPI_SUCCESS; }
