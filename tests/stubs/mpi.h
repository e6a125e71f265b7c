// Stand-in for <mpi.h> (absent here): TEST INFRASTRUCTURE for the syntax check of the plugin class.
#pragma once
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef struct { int a; } MPI_Status;
typedef void *MPI_File;
typedef int MPI_Info;
#define MPI_BYTE 1
#define MPI_COMM_WORLD 0
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Barrier(MPI_Comm comm);
