// scan.cu -- exclusive prefix sum (u32) used by the selection stage for ordered compaction and
// CSR offsets.  One pass: 4096-element tiles chained by a decoupled look-back.
#include "context.cuh"

#define SC_THREADS 256
#define SC_ITEMS 16
#define SC_TILE (SC_THREADS * SC_ITEMS)

__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32 *s_warp, u32 *block_total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	u32 incl = v;
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		u32 t = __shfl_up_sync(0xffffffffu, incl, off);
		if (lane >= off) incl += t;
	}
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	u32 pre = 0, tot = 0;
#pragma unroll
	for (int w = 0; w < SC_THREADS / 32; w++) {
		u32 s = s_warp[w];
		if (w < warp) pre += s;
		tot += s;
	}
	*block_total = tot;
	return pre + incl - v;
}

// single block: scan up to SC_TILE elements and write the grand total
__global__ void __launch_bounds__(SC_THREADS) scan_top_kernel(u32 *__restrict__ data, i64 n, u64 *__restrict__ total)
{
	__shared__ u32 s_warp[SC_THREADS / 32];
	const i64 base = (i64) threadIdx.x * SC_ITEMS;
	u32 v[SC_ITEMS];
	u32 sum = 0;
#pragma unroll
	for (int i = 0; i < SC_ITEMS; i++) {
		v[i] = base + i < n ? data[base + i] : 0u;
		sum += v[i];
	}
	u32 tot;
	u32 run = block_exclusive_scan(sum, s_warp, &tot);
#pragma unroll
	for (int i = 0; i < SC_ITEMS; i++) {
		if (base + i < n) data[base + i] = run;
		run += v[i];
	}
	if (threadIdx.x == 0 && total) *total = tot;
}

// ---- single pass: decoupled look-back over 4096-element tiles ---------------------------------
// Tiles take a ticket (so a tile only ever waits for tiles that started before it), publish their sum, look back over the
// predecessors' (aggregate | inclusive prefix) words and publish their own inclusive prefix.  One launch per scan instead of three.
#define SC_AGG (1ull << 62)
#define SC_INC (2ull << 62)
#define SC_VAL ((1ull << 62) - 1ull)
__global__ void __launch_bounds__(SC_THREADS) scan_lookback_kernel(const u32 *__restrict__ in, u32 *__restrict__ out, i64 n, u64 *state /* [tiles] zeroed */,
								   u32 *ticket /* zeroed */, u64 *__restrict__ total)
{
	__shared__ u32 s_warp[SC_THREADS / 32];
	__shared__ u32 s_tile;
	__shared__ u64 s_prefix;
	if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
	__syncthreads();
	const u32 tile = s_tile;
	const i64 base = (i64) tile * SC_TILE + (i64) threadIdx.x * SC_ITEMS;
	u32 v[SC_ITEMS];
	u32 sum = 0;
#pragma unroll
	for (int i = 0; i < SC_ITEMS; i++) {
		v[i] = base + i < n ? in[base + i] : 0u;
		sum += v[i];
	}
	u32 tot;
	u32 run = block_exclusive_scan(sum, s_warp, &tot);
	// look-back by the first warp: 32 predecessors per step, newest first; a lane whose tile has published nothing yet makes the
	// warp re-read the window
	if (threadIdx.x < 32) {
		volatile u64 *vs = state;
		const int lane = threadIdx.x;
		u64 prefix = 0;
		if (tile == 0) { if (lane == 0) vs[0] = SC_INC | (u64) tot; }
		else {
			if (lane == 0) { vs[tile] = SC_AGG | (u64) tot; }
			__threadfence();
			i64 hi = (i64) tile - 1;                                     // newest tile not yet accounted for
			while (hi >= 0) {
				const i64 t = hi - lane;
				u64 w = t >= 0 ? vs[t] : SC_INC;                       // before the array: an inclusive prefix of 0
				const u32 ready = __ballot_sync(0xffffffffu, (w >> 62) != 0);
				const u32 inc = __ballot_sync(0xffffffffu, (w >> 62) == 2);
				const int stop = inc ? __ffs(inc) - 1 : 31;              // newest lane holding an inclusive prefix (else: the whole window)
				const u32 need = stop == 31 ? 0xffffffffu : ((2u << stop) - 1u);
				if ((ready & need) != need) continue;                    // a tile this side of the stop has published nothing yet: re-read
				u64 val = lane <= stop ? (w & SC_VAL) : 0;
#pragma unroll
				for (int off = 16; off > 0; off >>= 1) val += __shfl_xor_sync(0xffffffffu, val, off);
				prefix += val;
				if (inc) break;
				hi -= 32;
			}
			if (lane == 0) vs[tile] = SC_INC | (prefix + (u64) tot);
		}
		if (lane == 0) {
			s_prefix = prefix;
			if (total && (i64) (tile + 1) * SC_TILE >= n) *total = prefix + (u64) tot;
		}
	}
	__syncthreads();
	run += (u32) s_prefix;
#pragma unroll
	for (int i = 0; i < SC_ITEMS; i++) {
		if (base + i < n) out[base + i] = run;
		run += v[i];
	}
}

int scan_exclusive_u32(metro_ctx *ctx, cudaStream_t st, const u32 *in, u32 *out, i64 n, u64 *d_total)
{
	if (n <= 0) {
		if (d_total) CUDA_TRY(ctx, cudaMemsetAsync(d_total, 0, sizeof(u64), st));
		return METRO_OK;
	}
	if (n <= SC_TILE) {
		if (in != out) CUDA_TRY(ctx, cudaMemcpyAsync(out, in, (size_t) n * 4, cudaMemcpyDeviceToDevice, st));
		scan_top_kernel<<<1, SC_THREADS, 0, st>>>(out, n, d_total);
		CUDA_TRY(ctx, cudaGetLastError());
		ctx->launches++;
		return METRO_OK;
	}
	const i64 tiles = div_up<i64>(n, SC_TILE);
	int rc = dev_reserve(ctx, ctx->scan_tmp, (size_t) tiles * 8 + 64);
	if (rc) return rc;
	u64 *state = ctx->scan_tmp.as<u64>();
	CUDA_TRY(ctx, cudaMemsetAsync(state, 0, (size_t) tiles * 8 + 8, st));
	scan_lookback_kernel<<<(unsigned) tiles, SC_THREADS, 0, st>>>(in, out, n, state, reinterpret_cast<u32 *>(state + tiles), d_total);
	CUDA_TRY(ctx, cudaGetLastError());
	ctx->launches++;
	return METRO_OK;
}
