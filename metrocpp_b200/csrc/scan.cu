// scan.cu -- exclusive prefix sum (u32) used by the selection stage for ordered compaction and
// CSR offsets.  Reduce-then-scan over 4096-element tiles, recursing on the tile sums.
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

// tile-local exclusive scan; tile totals to sums[]
__global__ void __launch_bounds__(SC_THREADS) scan_tiles_kernel(const u32 *__restrict__ in, u32 *__restrict__ out,
								i64 n, u32 *__restrict__ sums)
{
	__shared__ u32 s_warp[SC_THREADS / 32];
	const i64 base = (i64) blockIdx.x * SC_TILE + (i64) threadIdx.x * SC_ITEMS;
	u32 v[SC_ITEMS];
	u32 sum = 0;
#pragma unroll
	for (int i = 0; i < SC_ITEMS; i++) {
		v[i] = base + i < n ? in[base + i] : 0u;
		sum += v[i];
	}
	u32 tot;
	u32 run = block_exclusive_scan(sum, s_warp, &tot);
#pragma unroll
	for (int i = 0; i < SC_ITEMS; i++) {
		if (base + i < n) out[base + i] = run;
		run += v[i];
	}
	if (threadIdx.x == 0) sums[blockIdx.x] = tot;
}

__global__ void scan_add_kernel(u32 *__restrict__ out, i64 n, const u32 *__restrict__ offs)
{
	const i64 i = (i64) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] += offs[i / SC_TILE];
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

static int scan_rec(metro_ctx *ctx, cudaStream_t st, const u32 *in, u32 *out, i64 n, u32 *tmp, u64 *d_total)
{
	if (n <= SC_TILE) {
		if (in != out) CUDA_TRY(ctx, cudaMemcpyAsync(out, in, (size_t) n * 4, cudaMemcpyDeviceToDevice, st));
		scan_top_kernel<<<1, SC_THREADS, 0, st>>>(out, n, d_total);
		CUDA_TRY(ctx, cudaGetLastError());
		ctx->launches++;
		return METRO_OK;
	}
	const i64 tiles = div_up<i64>(n, SC_TILE);
	scan_tiles_kernel<<<(unsigned) tiles, SC_THREADS, 0, st>>>(in, out, n, tmp);
	CUDA_TRY(ctx, cudaGetLastError());
	int rc = scan_rec(ctx, st, tmp, tmp, tiles, tmp + tiles, d_total);
	if (rc) return rc;
	scan_add_kernel<<<(unsigned) div_up<i64>(n, 256), 256, 0, st>>>(out, n, tmp);
	CUDA_TRY(ctx, cudaGetLastError());
	ctx->launches += 2;
	return METRO_OK;
}

int scan_exclusive_u32(metro_ctx *ctx, cudaStream_t st, const u32 *in, u32 *out, i64 n, u64 *d_total)
{
	if (n <= 0) {
		if (d_total) CUDA_TRY(ctx, cudaMemsetAsync(d_total, 0, sizeof(u64), st));
		return METRO_OK;
	}
	// tile sums of every level: n/4096 + n/4096^2 + ... < n/4000 + 8
	size_t need = ((size_t) n / 4000 + 64) * sizeof(u32);
	int rc = dev_reserve(ctx, ctx->scan_tmp, need);
	if (rc) return rc;
	return scan_rec(ctx, st, in, out, n, ctx->scan_tmp.as<u32>(), d_total);
}
