// comm.cu -- the multi-GPU exchange step (one context == one rank == one GPU).
//
// Replaces the reference's MPI layer for this path (reference src/Communication.cpp:63-283
// BufferSendRecv, :289-473 GatherMergerTrees, :477-537 SyncOrphanHalos; src/Grid.cpp): instead of
// shipping buffer halos and their particle lists between spatial domains, the tuples of BOTH
// catalogues are sharded by particle-ID range, so pack / sort / join are rank-local, and only the
// reduced partial pair counts travel:
//   * one all-to-all (grouped ncclSend/ncclRecv over NVLink): every (A,B,typeA,typeB,count) record
//     goes to owner(A) for the forward trees and to owner(B) for the backward trees
//     (owner = contiguous halo-index ranges); the owner sums the partials (integer, order free);
//   * one all-reduce(max) of bestDescendant[B] (4 B per catalogue-1 halo) before the mutual-best
//     filter, one of the token flags (4 B per catalogue-0 halo) before the shift.
#include <nccl.h>

#include <chrono>
#include <stdlib.h>

#include "context.cuh"

#define NCCL_TRY(ctx, expr)                                                                       \
	do {                                                                                      \
		ncclResult_t _r = (expr);                                                         \
		if (_r != ncclSuccess) {                                                          \
			metro_set_error((ctx), "%s:%d %s -> %s", __FILE__, __LINE__, #expr, ncclGetErrorString(_r)); \
			return METRO_ERR_COMM;                                                    \
		}                                                                                 \
	} while (0)

static inline ncclComm_t comm_of(const metro_ctx *ctx) { return reinterpret_cast<ncclComm_t>(ctx->comm); }

// contiguous ownership ranges: halo h of a catalogue with H halos belongs to rank h / ceil(H / world)
static inline i64 range_chunk(i64 H, int world) { return H > 0 ? (H + world - 1) / world : 1; }

void comm_owned_range(const metro_ctx *ctx, i64 H, i64 *lo, i64 *hi)
{
	if (ctx->world <= 1) { *lo = 0; *hi = H; return; }
	const i64 c = range_chunk(H, ctx->world);
	*lo = std::min<i64>(H, c * ctx->rank);
	*hi = std::min<i64>(H, c * (ctx->rank + 1));
}

int comm_allreduce_max_i32(metro_ctx *ctx, i32 *buf, i64 n)
{
	if (ctx->world <= 1 || n <= 0) return METRO_OK;
	NCCL_TRY(ctx, ncclAllReduce(buf, buf, (size_t) n, ncclInt32, ncclMax, comm_of(ctx), ctx->stream));
	return METRO_OK;
}

int comm_allreduce_sum_u64(metro_ctx *ctx, u64 *buf, i64 n)
{
	if (ctx->world <= 1 || n <= 0) return METRO_OK;
	NCCL_TRY(ctx, ncclAllReduce(buf, buf, (size_t) n, ncclUint64, ncclSum, comm_of(ctx), ctx->stream));
	return METRO_OK;
}

// max of a status code over the ranks (no-op on one GPU); one small all-reduce + one stream synchronisation
int comm_agree_status(metro_ctx *ctx, int rc)
{
	if (ctx->world <= 1) return rc;
	cudaStream_t st = ctx->stream;
	u64 *d = ctx->counters.as<u64>() + 62;
	ctx->h_counters[62] = (u64) (rc < 0 ? 0 : rc);
	if (cudaMemcpyAsync(d, ctx->h_counters + 62, sizeof(u64), cudaMemcpyHostToDevice, st) != cudaSuccess) return rc ? rc : METRO_ERR_CUDA;
	if (ncclAllReduce(d, d, 1, ncclUint64, ncclMax, comm_of(ctx), st) != ncclSuccess) return rc ? rc : METRO_ERR_COMM;
	if (cudaMemcpyAsync(ctx->h_counters + 62, d, sizeof(u64), cudaMemcpyDeviceToHost, st) != cudaSuccess) return rc ? rc : METRO_ERR_CUDA;
	if (cudaStreamSynchronize(st) != cudaSuccess) return rc ? rc : METRO_ERR_CUDA;
	return (int) ctx->h_counters[62];
}

void comm_destroy(metro_ctx *ctx)
{
	if (ctx->comm) ncclCommDestroy(comm_of(ctx));
	ctx->comm = nullptr;
	dev_free(ctx->xsend_key); dev_free(ctx->xsend_cnt); dev_free(ctx->xrecvA_key); dev_free(ctx->xrecvA_cnt);
	dev_free(ctx->xrecvB_key); dev_free(ctx->xrecvB_cnt); dev_free(ctx->xcounts);
	dev_free(ctx->xredA_key); dev_free(ctx->xredA_cnt); dev_free(ctx->xredB_key); dev_free(ctx->xredB_cnt);
}

// ---- routing ---------------------------------------------------------------------------------
// destination slot of a record: dir * world + owner ; dir 0 = by haloA, 1 = by haloB
__global__ void route_count_kernel(const u64 *__restrict__ rec, i64 n, int halo_bits, int type_bits, i64 chunk0, i64 chunk1, int world,
				   unsigned long long *counts /* [2*world] */)
{
	extern __shared__ unsigned int sh[];
	for (int i = threadIdx.x; i < 2 * world; i += blockDim.x) sh[i] = 0;
	__syncthreads();
	const u64 hm = (1ull << halo_bits) - 1ull;
	for (i64 i = (i64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64) gridDim.x * blockDim.x) {
		const u64 hk = rec[i];
		const i64 B = (i64) ((hk >> (2 * type_bits)) & hm), A = (i64) ((hk >> (2 * type_bits + halo_bits)) & hm);
		atomicAdd(&sh[(int) (A / chunk0)], 1u);
		atomicAdd(&sh[world + (int) (B / chunk1)], 1u);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 2 * world; i += blockDim.x) if (sh[i]) atomicAdd(&counts[i], (unsigned long long) sh[i]);
}

// Block-aggregated scatter: a block counts its records per destination slot in shared memory,
// reserves one contiguous run per slot with a single global atomic each, then places its records
// (the order inside a slot is irrelevant: the owner sums the partials).
__global__ void __launch_bounds__(256) route_scatter_kernel(const u64 *__restrict__ rec, const u32 *__restrict__ cnt, i64 n, int halo_bits,
							    int type_bits, i64 chunk0, i64 chunk1, int world,
							    unsigned long long *cursor /* [2*world], preset to the slot offsets */,
							    u64 *__restrict__ out_key, u32 *__restrict__ out_cnt)
{
	extern __shared__ unsigned long long shb[];                  // [2*world] block bases, then u32 [2*world] counts
	unsigned int *shc = reinterpret_cast<unsigned int *>(shb + 2 * world);
	for (int i = threadIdx.x; i < 2 * world; i += blockDim.x) shc[i] = 0;
	__syncthreads();
	const u64 hm = (1ull << halo_bits) - 1ull;
	const i64 per = (n + gridDim.x - 1) / gridDim.x;
	const i64 lo = (i64) blockIdx.x * per, hi = lo + per < n ? lo + per : n;
	for (i64 i = lo + threadIdx.x; i < hi; i += blockDim.x) {
		const u64 hk = rec[i];
		const i64 B = (i64) ((hk >> (2 * type_bits)) & hm), A = (i64) ((hk >> (2 * type_bits + halo_bits)) & hm);
		atomicAdd(&shc[(int) (A / chunk0)], 1u);
		atomicAdd(&shc[world + (int) (B / chunk1)], 1u);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 2 * world; i += blockDim.x) {
		shb[i] = shc[i] ? atomicAdd(&cursor[i], (unsigned long long) shc[i]) : 0ull;
		shc[i] = 0;
	}
	__syncthreads();
	for (i64 i = lo + threadIdx.x; i < hi; i += blockDim.x) {
		const u64 hk = rec[i];
		const u32 c = cnt[i];
		const i64 B = (i64) ((hk >> (2 * type_bits)) & hm), A = (i64) ((hk >> (2 * type_bits + halo_bits)) & hm);
		const int sa = (int) (A / chunk0), sb = world + (int) (B / chunk1);
		const unsigned long long a = shb[sa] + atomicAdd(&shc[sa], 1u);
		out_key[a] = hk; out_cnt[a] = c;
		const unsigned long long b2 = shb[sb] + atomicAdd(&shc[sb], 1u);
		out_key[b2] = hk; out_cnt[b2] = c;
	}
}

int comm_exchange_records(metro_ctx *ctx, int halo_bits, int type_bits, const u64 *rec_key, const u32 *rec_cnt, i64 R,
			  const u64 **recA_key, const u32 **recA_cnt, i64 *RA, const u64 **recB_key, const u32 **recB_cnt, i64 *RB)
{
	if (ctx->world <= 1) {
		*recA_key = *recB_key = rec_key; *recA_cnt = *recB_cnt = rec_cnt; *RA = *RB = R;
		return METRO_OK;
	}
	cudaStream_t st = ctx->stream;
	const int W = ctx->world, me = ctx->rank;
	const i64 chunk0 = range_chunk(ctx->cat[0].n_halos, W), chunk1 = range_chunk(ctx->cat[1].n_halos, W);
	int rc;
	static const bool dbg = getenv("METRO_DEBUG") != nullptr;
	auto now = [&] { cudaStreamSynchronize(st); return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
	double t0 = dbg ? now() : 0, t1 = 0, t2 = 0, t3 = 0;
	// counts per destination slot, then the full W x 2W count matrix on every rank
	if ((rc = dev_reserve(ctx, ctx->xcounts, (size_t) (2 * W) * (W + 2) * sizeof(u64)))) return rc;
	unsigned long long *d_counts = ctx->xcounts.as<unsigned long long>();                // [2W] mine
	unsigned long long *d_cursor = d_counts + 2 * W;                                     // [2W]
	unsigned long long *d_matrix = d_cursor + 2 * W;                                     // [W][2W]
	CUDA_TRY(ctx, cudaMemsetAsync(d_counts, 0, (size_t) 2 * W * sizeof(u64), st));
	const int blocks = (int) std::min<i64>(div_up<i64>(std::max<i64>(R, 1), 256), (i64) METRO_NUM_SMS_B200 * 8);
	if (R > 0) {
		route_count_kernel<<<blocks, 256, 2 * W * sizeof(unsigned int), st>>>(rec_key, R, halo_bits, type_bits, chunk0, chunk1, W, d_counts);
		CUDA_TRY(ctx, cudaGetLastError());
		ctx->launches++;
	}
	if (dbg) t1 = now();
	NCCL_TRY(ctx, ncclAllGather(d_counts, d_matrix, (size_t) 2 * W, ncclUint64, comm_of(ctx), st));
	std::vector<u64> matrix((size_t) W * 2 * W);
	CUDA_TRY(ctx, cudaMemcpyAsync(matrix.data(), d_matrix, matrix.size() * sizeof(u64), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	if (dbg) t2 = now();
	// send layout: slots in order [A->0 .. A->W-1, B->0 .. B->W-1]
	std::vector<u64> soff(2 * W + 1, 0);
	for (int s = 0; s < 2 * W; s++) soff[s + 1] = soff[s] + matrix[(size_t) me * 2 * W + s];
	// receive layout: recvA = concat over source ranks of their slot (A -> me); recvB likewise
	std::vector<u64> roffA(W + 1, 0), roffB(W + 1, 0);
	for (int p = 0; p < W; p++) {
		roffA[p + 1] = roffA[p] + matrix[(size_t) p * 2 * W + me];
		roffB[p + 1] = roffB[p] + matrix[(size_t) p * 2 * W + W + me];
	}
	const u64 nA = roffA[W], nB = roffB[W];
	if ((rc = dev_reserve(ctx, ctx->xsend_key, (size_t) (2 * R + 1) * 8))) return rc;
	if ((rc = dev_reserve(ctx, ctx->xsend_cnt, (size_t) (2 * R + 1) * 4))) return rc;
	if ((rc = dev_reserve(ctx, ctx->xrecvA_key, (size_t) (nA + 1) * 8))) return rc;
	if ((rc = dev_reserve(ctx, ctx->xrecvA_cnt, (size_t) (nA + 1) * 4))) return rc;
	if ((rc = dev_reserve(ctx, ctx->xrecvB_key, (size_t) (nB + 1) * 8))) return rc;
	if ((rc = dev_reserve(ctx, ctx->xrecvB_cnt, (size_t) (nB + 1) * 4))) return rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(d_cursor, soff.data(), (size_t) 2 * W * sizeof(u64), cudaMemcpyHostToDevice, st));
	if (R > 0) {
		route_scatter_kernel<<<blocks, 256, (size_t) 2 * W * 12, st>>>(rec_key, rec_cnt, R, halo_bits, type_bits, chunk0, chunk1, W, d_cursor,
							      ctx->xsend_key.as<u64>(), ctx->xsend_cnt.as<u32>());
		CUDA_TRY(ctx, cudaGetLastError());
		ctx->launches++;
	}
	if (dbg) t3 = now();
	// the all-to-all: 4 messages per peer and direction, all in one NCCL group
	u64 *sk = ctx->xsend_key.as<u64>(); u32 *sc = ctx->xsend_cnt.as<u32>();
	u64 *rak = ctx->xrecvA_key.as<u64>(); u32 *rac = ctx->xrecvA_cnt.as<u32>();
	u64 *rbk = ctx->xrecvB_key.as<u64>(); u32 *rbc = ctx->xrecvB_cnt.as<u32>();
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[11], st));                    // the payload all-to-all alone: ev[11] .. ev[12]
	NCCL_TRY(ctx, ncclGroupStart());
	for (int p = 0; p < W; p++) {
		const u64 sa = matrix[(size_t) me * 2 * W + p], sb = matrix[(size_t) me * 2 * W + W + p];
		const u64 ra = matrix[(size_t) p * 2 * W + me], rb = matrix[(size_t) p * 2 * W + W + me];
		if (sa) { NCCL_TRY(ctx, ncclSend(sk + soff[p], sa, ncclUint64, p, comm_of(ctx), st)); NCCL_TRY(ctx, ncclSend(sc + soff[p], sa, ncclUint32, p, comm_of(ctx), st)); }
		if (sb) { NCCL_TRY(ctx, ncclSend(sk + soff[W + p], sb, ncclUint64, p, comm_of(ctx), st)); NCCL_TRY(ctx, ncclSend(sc + soff[W + p], sb, ncclUint32, p, comm_of(ctx), st)); }
		if (ra) { NCCL_TRY(ctx, ncclRecv(rak + roffA[p], ra, ncclUint64, p, comm_of(ctx), st)); NCCL_TRY(ctx, ncclRecv(rac + roffA[p], ra, ncclUint32, p, comm_of(ctx), st)); }
		if (rb) { NCCL_TRY(ctx, ncclRecv(rbk + roffB[p], rb, ncclUint64, p, comm_of(ctx), st)); NCCL_TRY(ctx, ncclRecv(rbc + roffB[p], rb, ncclUint32, p, comm_of(ctx), st)); }
	}
	NCCL_TRY(ctx, ncclGroupEnd());
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[12], st));
	ctx->exchange_timed = true;
	if (dbg) {
		const double t4 = now();
		fprintf(stderr, "[metro] rank %d exchange: R %lld count %.2f ms, all-gather+sync %.2f ms, reserve+scatter %.2f ms, send/recv %.2f ms (nA %llu nB %llu)\n", me,
			(long long) R, t1 - t0, t2 - t1, t3 - t2, t4 - t3, (unsigned long long) nA, (unsigned long long) nB);
	}
	*recA_key = rak; *recA_cnt = rac; *RA = (i64) nA;
	*recB_key = rbk; *recB_cnt = rbc; *RB = (i64) nB;
	ctx->exchange_bytes = (i64) ((soff[2 * W] - matrix[(size_t) me * 2 * W + me] - matrix[(size_t) me * 2 * W + W + me]) * 12);
	return METRO_OK;
}

extern "C" {

int metro_comm_unique_id(void *out128)
{
	if (!out128) return METRO_ERR_INVALID;
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	if (ncclGetUniqueId(&id) != ncclSuccess) { metro_set_error(nullptr, "ncclGetUniqueId failed"); return METRO_ERR_COMM; }
	memcpy(out128, &id, sizeof id);
	return METRO_OK;
}

int metro_comm_init(metro_ctx *ctx, int rank, int world, const void *unique_id128)
{
	if (!ctx || !unique_id128 || world < 1 || rank < 0 || rank >= world) return METRO_ERR_INVALID;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	comm_destroy(ctx);
	ctx->rank = rank; ctx->world = world;
	if (world == 1) return METRO_OK;
	ncclUniqueId id;
	memcpy(&id, unique_id128, sizeof id);
	ncclComm_t c;
	NCCL_TRY(ctx, ncclCommInitRank(&c, world, id, rank));
	ctx->comm = c;
	return METRO_OK;
}

}  // extern "C"
