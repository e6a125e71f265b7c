// Stand-in for <mpi.h> (absent here): TEST INFRASTRUCTURE for the syntax check of the plugin class.
#pragma once
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct { int a; } MPI_Status;
typedef void *MPI_File;
typedef int MPI_Info;
#define MPI_BYTE 1
#define MPI_DOUBLE 2
#define MPI_SUCCESS 0
typedef int MPI_Op;
#define MPI_SUM 1
int MPI_Allreduce(const void *sendbuf, void *recvbuf, int count, MPI_Datatype type, MPI_Op op, MPI_Comm comm);
#define MPI_COMM_WORLD 0
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);
