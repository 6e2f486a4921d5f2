// partition.cu -- MICRO-BENCHMARKS, not part of the C ABI of include/metro_b200.h and not on the product path: the
// entry points below are called by scripts/*_bench.py only (numbers in profiles/r01_experiments.txt).
//
// (1) non-stable 256-way partition pass with fixed-capacity
// buckets, the building block of an MSD / hash-partition join.  Ranks come from one shared-memory
// atomic per key (no ballots, no look-back); bucket space is reserved with one global atomic per
// (tile, bucket).
#include "context.cuh"

#ifndef PT_THREADS
#define PT_THREADS 512
#endif
#ifndef PT_ITEMS
#define PT_ITEMS 16
#endif
#ifndef PT_MINB
#define PT_MINB 1
#endif
#define PT_TILE (PT_THREADS * PT_ITEMS)
#define PT_BUCKETS 256

__global__ void __launch_bounds__(PT_THREADS, PT_MINB) partition_kernel(const u64 *__restrict__ in, u64 *__restrict__ out, u64 n, int shift, u64 cap,
							       unsigned long long *gcount /* [256] */, u32 *overflow)
{
	extern __shared__ __align__(16) unsigned char psm[];
	u64 *s_keys = reinterpret_cast<u64 *>(psm);                         // [PT_TILE]
	i64 *s_dst = reinterpret_cast<i64 *>(s_keys + PT_TILE);             // [256]  global slot of the run minus its tile-local start
	u32 *s_cnt = reinterpret_cast<u32 *>(s_dst + PT_BUCKETS);          // [256]
	u32 *s_start = s_cnt + PT_BUCKETS;                                  // [256]
	__shared__ u32 s_warpsum[PT_BUCKETS / 32];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid < PT_BUCKETS) s_cnt[tid] = 0;
	__syncthreads();
	const u64 base = (u64) blockIdx.x * PT_TILE;
	const u64 wbase = base + (u64) warp * (PT_ITEMS * 32) + lane;
	u64 key[PT_ITEMS];
	u32 rk[PT_ITEMS];
	const bool full = base + PT_TILE <= n;
#pragma unroll
	for (int i = 0; i < PT_ITEMS; i++) {
		const u64 idx = wbase + 32 * i;
		key[i] = (full || idx < n) ? ld_stream_u64(in + idx) : 0ull;
	}
#pragma unroll
	for (int i = 0; i < PT_ITEMS; i++) {
		const u32 d = (u32) (key[i] >> shift) & (PT_BUCKETS - 1);
		rk[i] = (full || wbase + 32 * i < n) ? atomicAdd(&s_cnt[d], 1u) : 0xFFFFFFFFu;
	}
	__syncthreads();
	u32 total = 0, incl = 0;
	if (tid < PT_BUCKETS) {
		total = s_cnt[tid];
		incl = total;
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) {
			u32 v = __shfl_up_sync(0xffffffffu, incl, off);
			if (lane >= off) incl += v;
		}
		if (lane == 31) s_warpsum[warp] = incl;
	}
	__syncthreads();
	if (tid < PT_BUCKETS) {
		u32 pre = 0;
#pragma unroll
		for (int w = 0; w < PT_BUCKETS / 32; w++) pre += (w < warp) ? s_warpsum[w] : 0u;
		const u32 start = pre + incl - total;
		s_start[tid] = start;
		unsigned long long g = total ? atomicAdd(&gcount[tid], (unsigned long long) total) : 0ull;
		if (g + total > cap) { *overflow = 1; g = 0; }
		s_dst[tid] = (i64) ((u64) tid * cap + g) - (i64) start;
	}
	__syncthreads();
#pragma unroll
	for (int i = 0; i < PT_ITEMS; i++) {
		if (rk[i] == 0xFFFFFFFFu) continue;
		const u32 d = (u32) (key[i] >> shift) & (PT_BUCKETS - 1);
		s_keys[s_start[d] + rk[i]] = key[i];
	}
	__syncthreads();
	const u32 count = full ? (u32) PT_TILE : (u32) (n - base);
#pragma unroll 4
	for (u32 j = tid; j < count; j += PT_THREADS) {
		const u64 k = s_keys[j];
		const u32 d = (u32) (k >> shift) & (PT_BUCKETS - 1);
		st_stream_u64(out + s_dst[d] + (i64) j, k);
	}
}

extern "C" int metro_partition_bench(metro_ctx *ctx, const uint64_t *keys_in, uint64_t *keys_out, int64_t n, int shift, int64_t cap, int reps,
				     float *ms_out, uint64_t *counts_out /* [256] host */)
{
	if (!ctx || n <= 0) return METRO_ERR_INVALID;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	const size_t smem = (size_t) PT_TILE * 8 + PT_BUCKETS * 8 + PT_BUCKETS * 8;
	CUDA_TRY(ctx, cudaFuncSetAttribute(partition_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	unsigned long long *g = ctx->counters.as<unsigned long long>();
	int rc = dev_reserve(ctx, ctx->scratch[0], 257 * 8);
	if (rc) return rc;
	g = ctx->scratch[0].as<unsigned long long>();
	const unsigned tiles = (unsigned) div_up<i64>(n, PT_TILE);
	float best = 1e30f;
	for (int r = 0; r < reps; r++) {
		CUDA_TRY(ctx, cudaMemsetAsync(g, 0, 257 * 8, st));
		CUDA_TRY(ctx, cudaEventRecord(ctx->ev[8], st));
		partition_kernel<<<tiles, PT_THREADS, smem, st>>>(keys_in, keys_out, (u64) n, shift, (u64) cap, g, reinterpret_cast<u32 *>(g + 256));
		CUDA_TRY(ctx, cudaGetLastError());
		CUDA_TRY(ctx, cudaEventRecord(ctx->ev[9], st));
		CUDA_TRY(ctx, cudaStreamSynchronize(st));
		float ms;
		cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]);
		if (ms < best) best = ms;
	}
	if (ms_out) *ms_out = best;
	if (counts_out) CUDA_TRY(ctx, cudaMemcpy(counts_out, g, 257 * 8, cudaMemcpyDeviceToHost));
	return METRO_OK;
}

// EXPERIMENT: random 32-byte sector gather bandwidth (decides whether a global-memory hash probe can
// compete with the second partition shuffle).  Every thread reads `per_thread` pseudo-random sectors of
// a table of n_sectors x 32 B with `ilp` independent loads in flight and folds them into one word.
template <int ILP>
__global__ void __launch_bounds__(256) gather_bench_kernel(const ulonglong4 *__restrict__ table, u64 n_sectors, int per_thread, u64 *__restrict__ out)
{
	const u64 t = (u64) blockIdx.x * 256 + threadIdx.x;
	u64 acc = 0, x = t * 0x9E3779B97F4A7C15ull + 12345;
	for (int i = 0; i < per_thread; i += ILP) {
		ulonglong4 v[ILP];
#pragma unroll
		for (int q = 0; q < ILP; q++) {
			x = x * 6364136223846793005ull + 1442695040888963407ull;
			const u64 idx = __umul64hi(x, n_sectors);                   // uniform in [0, n_sectors)
			// one 32-byte sector per probe (256-bit load, sm_100)
			asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v[q].x), "=l"(v[q].y), "=l"(v[q].z), "=l"(v[q].w) : "l"(table + idx));
		}
#pragma unroll
		for (int q = 0; q < ILP; q++) acc += v[q].x ^ v[q].y ^ v[q].z ^ v[q].w;
	}
	out[t] = acc;
}

extern "C" int metro_gather_bench(metro_ctx *ctx, const void *table, uint64_t n_sectors, int64_t n_threads, int per_thread, int ilp, uint64_t *out_dev,
				  float *ms_out)
{
	if (!ctx) return METRO_ERR_INVALID;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	const unsigned blocks = (unsigned) (n_threads / 256);
	float best = 1e30f;
	for (int r = 0; r < 3; r++) {
		CUDA_TRY(ctx, cudaEventRecord(ctx->ev[8], st));
		if (ilp == 8) gather_bench_kernel<8><<<blocks, 256, 0, st>>>((const ulonglong4 *) table, n_sectors, per_thread, out_dev);
		else if (ilp == 4) gather_bench_kernel<4><<<blocks, 256, 0, st>>>((const ulonglong4 *) table, n_sectors, per_thread, out_dev);
		else gather_bench_kernel<1><<<blocks, 256, 0, st>>>((const ulonglong4 *) table, n_sectors, per_thread, out_dev);
		CUDA_TRY(ctx, cudaGetLastError());
		CUDA_TRY(ctx, cudaEventRecord(ctx->ev[9], st));
		CUDA_TRY(ctx, cudaStreamSynchronize(st));
		float ms;
		cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]);
		if (ms < best) best = ms;
	}
	*ms_out = best;
	return METRO_OK;
}

// EXPERIMENT: rate of small shared->global bulk copies (cp.async.bulk, the TMA engine), the candidate for writing
// the runs of a partition tile without going through the LSU.  Every thread issues `per_thread` copies of `bytes`
// bytes from the block's shared tile to scattered 16-byte aligned destinations.
__global__ void __launch_bounds__(256) bulk_store_bench_kernel(u64 *__restrict__ out, u64 out_words, int per_thread, int bytes)
{
	__shared__ __align__(128) u64 tile[4096];
	for (int i = threadIdx.x; i < 4096; i += 256) tile[i] = (u64) blockIdx.x * 4096 + i;
	__syncthreads();
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
	const u32 words = (u32) bytes / 8;
	const u64 t = (u64) blockIdx.x * 256 + threadIdx.x;
	u64 x = t * 0x9E3779B97F4A7C15ull + 1;
	const u32 src = (u32) __cvta_generic_to_shared(tile + ((threadIdx.x * words) & 4095 & ~1u));
	for (int i = 0; i < per_thread; i++) {
		x = x * 6364136223846793005ull + 1442695040888963407ull;
		const u64 slot = __umul64hi(x, out_words / words - 1);
		u64 *dst = out + slot * words;
		asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"((u32) bytes) : "memory");
		if ((i & 7) == 7) {
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
		}
	}
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
	asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

extern "C" int metro_bulk_store_bench(metro_ctx *ctx, void *out, uint64_t out_bytes, int blocks, int per_thread, int bytes, float *ms_out)
{
	if (!ctx || bytes % 16) return METRO_ERR_INVALID;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	float best = 1e30f;
	for (int r = 0; r < 3; r++) {
		CUDA_TRY(ctx, cudaEventRecord(ctx->ev[8], st));
		bulk_store_bench_kernel<<<blocks, 256, 0, st>>>((u64 *) out, out_bytes / 8, per_thread, bytes);
		CUDA_TRY(ctx, cudaGetLastError());
		CUDA_TRY(ctx, cudaEventRecord(ctx->ev[9], st));
		CUDA_TRY(ctx, cudaStreamSynchronize(st));
		float ms;
		cudaEventElapsedTime(&ms, ctx->ev[8], ctx->ev[9]);
		if (ms < best) best = ms;
	}
	*ms_out = best;
	return METRO_OK;
}
