/* Multi-process stand-in for <mpi.h> on ONE node.  TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * This image has no MPI runtime, and BASELINE.json's north_star asks for "the reference's MPI CPU build, run on the box's own host
 * cores".  This header + mpi_mp.cpp implement exactly the MPI symbols the reference uses (Init, Finalize,
 * Comm_rank, Comm_size, Barrier, Bcast, Reduce(SUM of ints), Gather, Gatherv, Sendrecv, Pack, Unpack) for K processes started by
 * oracle/mpirun_shim.py: ranks exchange messages through buffered per-(source, destination) mailboxes in /dev/shm, so the
 * UNMODIFIED reference sources run their real multi-rank path (Communication.cpp buffer exchange, GatherMergerTrees, ...).
 * Message semantics: sends are buffered (a sender blocks only while its previous message to the same peer is unread), receives
 * block; tags are ignored (the reference's traffic per pair is strictly ordered). */
#ifndef METRO_ORACLE_MPI_SHIM_MP_H
#define METRO_ORACLE_MPI_SHIM_MP_H
#include <cstddef>
#include <cstdint>
#include <cstring>

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef struct { int unused; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_BYTE   1
#define MPI_PACKED 1
#define MPI_INT    4
#define MPI_FLOAT  5
#define MPI_SUM    0

int MPI_Init(int *, char ***);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm, int *rank);
int MPI_Comm_size(MPI_Comm, int *size);
int MPI_Barrier(MPI_Comm);
int MPI_Bcast(void *buf, int n, MPI_Datatype t, int root, MPI_Comm);
int MPI_Reduce(const void *snd, void *rcv, int n, MPI_Datatype t, MPI_Op, int root, MPI_Comm);
int MPI_Gather(const void *snd, int n, MPI_Datatype t, void *rcv, int rn, MPI_Datatype rt, int root, MPI_Comm);
int MPI_Gatherv(const void *snd, int n, MPI_Datatype t, void *rcv, const int *counts, const int *displ, MPI_Datatype rt, int root, MPI_Comm);
int MPI_Sendrecv(const void *snd, int n, MPI_Datatype t, int dest, int stag, void *rcv, int rn, MPI_Datatype rt, int source, int rtag, MPI_Comm,
		 MPI_Status *);
int MPI_Pack(const void *in, int n, MPI_Datatype t, void *out, int outsize, int *pos, MPI_Comm);
int MPI_Unpack(const void *in, int insize, int *pos, void *out, int n, MPI_Datatype t, MPI_Comm);

#endif
