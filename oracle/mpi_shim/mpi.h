/* Single-rank stand-in for <mpi.h>.  TEST INFRASTRUCTURE ONLY.
 *
 * Lets the UNMODIFIED reference sources under /root/reference/src compile and
 * run as one rank (this image has no MPI).  Covers exactly the MPI symbols the
 * reference uses: every collective degenerates to a
 * memcpy from the send buffer to the receive buffer, Pack/Unpack to a memcpy
 * with a running byte position.  Nothing here is linked into the product. */
#ifndef METRO_ORACLE_MPI_SHIM_H
#define METRO_ORACLE_MPI_SHIM_H
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

static inline size_t shim_width(MPI_Datatype t) { return t == MPI_BYTE ? 1u : 4u; }
static inline void shim_copy(void *dst, const void *src, int n, MPI_Datatype t)
{
	if (n > 0 && dst != src) memcpy(dst, src, (size_t) n * shim_width(t));
}

static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = 0; return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *size) { *size = 1; return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }

static inline int MPI_Reduce(const void *snd, void *rcv, int n, MPI_Datatype t, MPI_Op, int, MPI_Comm)
{ shim_copy(rcv, snd, n, t); return 0; }

static inline int MPI_Gather(const void *snd, int n, MPI_Datatype t, void *rcv, int, MPI_Datatype, int, MPI_Comm)
{ shim_copy(rcv, snd, n, t); return 0; }

static inline int MPI_Gatherv(const void *snd, int n, MPI_Datatype t, void *rcv, const int *, const int *displ,
			      MPI_Datatype, int, MPI_Comm)
{ shim_copy((char *) rcv + (size_t) displ[0] * shim_width(t), snd, n, t); return 0; }

static inline int MPI_Sendrecv(const void *snd, int n, MPI_Datatype t, int, int, void *rcv, int, MPI_Datatype,
			       int, int, MPI_Comm, MPI_Status *)
{ shim_copy(rcv, snd, n, t); return 0; }

static inline int MPI_Pack(const void *in, int n, MPI_Datatype t, void *out, int, int *pos, MPI_Comm)
{ size_t b = (size_t) n * shim_width(t); memcpy((char *) out + *pos, in, b); *pos += (int) b; return 0; }

static inline int MPI_Unpack(const void *in, int, int *pos, void *out, int n, MPI_Datatype t, MPI_Comm)
{ size_t b = (size_t) n * shim_width(t); memcpy(out, (const char *) in + *pos, b); *pos += (int) b; return 0; }

#endif
