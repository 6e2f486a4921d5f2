// radix_sort.cu -- onesweep LSD radix sort for sm_100a.
//
// Replaces the ordering work the reference gets from std::map<uint64_t, vector<Particle>>
// (reference src/global_vars.h:77-82, filled at src/IOSettings.cpp:628-632): instead of one
// red-black-tree insert per tuple, the packed (pid|snap|halo|type) keys of both catalogues are
// sorted by pid with P = ceil(pid_bits/8) stable counting passes.
//
// Per pass ONE kernel reads every key once and writes it once (16 B/key of HBM traffic):
//   * a block takes a tile (dynamic tile id => predecessors are always running),
//   * ranks its keys per 8-bit digit (warp match + per-warp shared counters),
//   * gets the global start of each of its 256 digit runs by decoupled look-back over the
//     previous tiles' published counts (no second read of the keys, no global scan kernel),
//   * reorders the tile in shared memory so that each digit run is written with coalesced
//     stores.
// The digit histograms of ALL passes are taken in one up-front read (fused with the key packing
// in pack_hist_kernel, join.cu).
#include "context.cuh"
#include "radix_sort.cuh"

// ---- tunables (B200: 148 SMs, 227 KB smem/CTA, 64K regs/SM) ----------------------------------
// 384 threads x 16 keys = 6144-key tile (48 KB of keys in smem) -> 3 CTAs/SM by shared memory,
// 2-3 by registers; grid = number of tiles, far above 148 x 3, so the tail is < 0.1 %.
#ifndef OS_THREADS
#define OS_THREADS 384
#endif
#ifndef OS_ITEMS
#define OS_ITEMS 16
#endif
#define OS_TILE (OS_THREADS * OS_ITEMS)
#define OS_WARPS (OS_THREADS / 32)

template <typename LB> struct LbTraits;
template <> struct LbTraits<u32> {
	static constexpr u32 AGG = 1u << 30, INC = 2u << 30, FLAGS = 3u << 30, MASK = (1u << 30) - 1u;
};
template <> struct LbTraits<u64> {
	static constexpr u64 AGG = 1ull << 62, INC = 2ull << 62, FLAGS = 3ull << 62, MASK = (1ull << 62) - 1ull;
};

// ---- histogram -------------------------------------------------------------------------------
#define HIST_THREADS 512
__global__ void __launch_bounds__(HIST_THREADS) hist_kernel(const u64 *__restrict__ keys, u64 n, SortPlan plan,
							    u64 *__restrict__ ghist)
{
	__shared__ u32 sh[SORT_MAX_PASSES][SORT_RADIX];
	for (int i = threadIdx.x; i < SORT_MAX_PASSES * SORT_RADIX; i += HIST_THREADS) (&sh[0][0])[i] = 0;
	__syncthreads();
	const u64 stride = (u64) gridDim.x * HIST_THREADS;
	// 16-byte loads: two keys per thread per step
	const u64 n2 = n >> 1;
	const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(keys);
	for (u64 i = (u64) blockIdx.x * HIST_THREADS + threadIdx.x; i < n2; i += stride) {
		ulonglong2 v = __ldg(k2 + i);
		hist_accumulate(sh, v.x, plan);
		hist_accumulate(sh, v.y, plan);
	}
	if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) hist_accumulate(sh, keys[n - 1], plan);
	__syncthreads();
	for (int i = threadIdx.x; i < plan.n_pass * SORT_RADIX; i += HIST_THREADS) {
		u32 c = (&sh[0][0])[i];
		if (c) atomicAdd((unsigned long long *) &ghist[i], (unsigned long long) c);
	}
}

// counts -> exclusive offsets, one block per pass
__global__ void __launch_bounds__(SORT_RADIX) hist_scan_kernel(u64 *ghist)
{
	__shared__ u64 s[SORT_RADIX];
	u64 *h = ghist + (size_t) blockIdx.x * SORT_RADIX;
	const int t = threadIdx.x;
	u64 c = h[t];
	s[t] = c;
	__syncthreads();
	for (int off = 1; off < SORT_RADIX; off <<= 1) {
		u64 v = t >= off ? s[t - off] : 0;
		__syncthreads();
		s[t] += v;
		__syncthreads();
	}
	h[t] = s[t] - c;
}

// ---- onesweep pass ---------------------------------------------------------------------------
template <bool HAS_VALS, typename LB>
__global__ void __launch_bounds__(OS_THREADS) onesweep_kernel(const u64 *__restrict__ in, u64 *__restrict__ out,
							      const u32 *__restrict__ vin, u32 *__restrict__ vout,
							      u64 n, int shift, u32 dmask,
							      const u64 *__restrict__ goff, LB *lookback,
							      u32 *tile_counter)
{
	typedef LbTraits<LB> T;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	u64 *s_keys = reinterpret_cast<u64 *>(smem_raw);                       // [OS_TILE]
	i64 *s_dst = reinterpret_cast<i64 *>(s_keys + OS_TILE);                // [256] global - local start
	u32 *s_whist = reinterpret_cast<u32 *>(s_dst + SORT_RADIX);           // [OS_WARPS][256]
	u32 *s_binstart = s_whist + OS_WARPS * SORT_RADIX;                     // [256]
	u32 *s_vals = s_binstart + SORT_RADIX;                                 // [OS_TILE] iff HAS_VALS
	__shared__ u32 s_tile;
	__shared__ u32 s_warpsum[SORT_RADIX / 32];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) s_tile = atomicAdd(tile_counter, 1u);
	for (int i = tid; i < OS_WARPS * SORT_RADIX; i += OS_THREADS) s_whist[i] = 0;
	__syncthreads();
	const u32 tile = s_tile;
	const u64 tile_base = (u64) tile * OS_TILE;
	const u64 wbase = tile_base + (u64) warp * (OS_ITEMS * 32) + lane;
	const bool full = tile_base + OS_TILE <= n;

	// ---- load (warp-striped: item i of lane l is element wbase + 32 i), 256 B per warp request
	u64 key[OS_ITEMS];
	u32 val[HAS_VALS ? OS_ITEMS : 1];
	if (full) {
#pragma unroll
		for (int i = 0; i < OS_ITEMS; i++) key[i] = ld_stream_u64(in + wbase + 32 * i);
		if (HAS_VALS) {
#pragma unroll
			for (int i = 0; i < OS_ITEMS; i++) val[i] = __ldg(vin + wbase + 32 * i);
		}
	} else {
#pragma unroll
		for (int i = 0; i < OS_ITEMS; i++) {
			u64 idx = wbase + 32 * i;
			key[i] = idx < n ? in[idx] : ~0ull;      // padding sorts to the end of the last bin
			if (HAS_VALS) val[i] = idx < n ? vin[idx] : 0u;
		}
	}

	// ---- rank inside the warp: peers = lanes holding the same digit.  MATCH.ANY costs ~95 issue
	// cycles per warp on sm_100 (ncu r01b: >50 % of all stall samples sat on it), so the peer mask is
	// built from 8 ballots (VOTE, 2 warp-instr/clk/SM) instead; the leader bumps the warp's counter.
	u32 *wh = s_whist + warp * SORT_RADIX;
	u32 rank[OS_ITEMS];
	const u32 lt = lanemask_lt();
#pragma unroll
	for (int i = 0; i < OS_ITEMS; i++) {
		const u32 d = (u32) (key[i] >> shift) & dmask;
		u32 peers = 0xffffffffu;
#pragma unroll
		for (int b = 0; b < SORT_RADIX_BITS; b++) {
			const bool bit = (d >> b) & 1u;
			const u32 m = __ballot_sync(0xffffffffu, bit);
			peers &= bit ? m : ~m;
		}
		const int leader = __ffs(peers) - 1;
		u32 base = 0;
		if (lane == leader) { base = wh[d]; wh[d] = base + __popc(peers); }
		base = __shfl_sync(0xffffffffu, base, leader);
		rank[i] = base + __popc(peers & lt);
		__syncwarp();
	}
	__syncthreads();

	// ---- per-digit: exclusive offsets across warps, tile total, look-back
	u32 total = 0;
	if (tid < SORT_RADIX) {
		u32 run = 0;
#pragma unroll
		for (int w = 0; w < OS_WARPS; w++) {
			u32 c = s_whist[w * SORT_RADIX + tid];
			s_whist[w * SORT_RADIX + tid] = run;
			run += c;
		}
		total = run;
		// publish this tile's count as early as possible
		LB *mine = lookback + (size_t) tile * SORT_RADIX + tid;
		*reinterpret_cast<volatile LB *>(mine) = (tile == 0 ? T::INC : T::AGG) | (LB) total;
	}
	// exclusive scan of the 256 totals -> local start of each digit run
	u32 incl = total;
	if (tid < SORT_RADIX) {
#pragma unroll
		for (int off = 1; off < 32; off <<= 1) {
			u32 v = __shfl_up_sync(0xffffffffu, incl, off);
			if (lane >= off) incl += v;
		}
		if (lane == 31) s_warpsum[warp] = incl;
	}
	__syncthreads();
	u32 binstart = 0;
	if (tid < SORT_RADIX) {
		u32 pre = 0;
#pragma unroll
		for (int w = 0; w < SORT_RADIX / 32; w++) pre += (w < warp) ? s_warpsum[w] : 0u;
		binstart = pre + incl - total;
		s_binstart[tid] = binstart;
	}
	__syncthreads();

	// ---- reorder the tile in shared memory (needs only tile-local offsets) ...
#pragma unroll
	for (int i = 0; i < OS_ITEMS; i++) {
		const u32 d = (u32) (key[i] >> shift) & dmask;
		const u32 pos = s_binstart[d] + wh[d] + rank[i];
		s_keys[pos] = key[i];
		if (HAS_VALS) s_vals[pos] = val[i];
	}
	// ---- ... and only then wait for the predecessors: decoupled look-back, one digit per thread.
	// The tile's own count was published before the reorder, so the tiles behind it are not held up.
	if (tid < SORT_RADIX) {
		u64 excl = 0;
		if (tile > 0) {
			i64 t = (i64) tile - 1;
			while (true) {
				const volatile LB *p = reinterpret_cast<const volatile LB *>(lookback + (size_t) t * SORT_RADIX + tid);
				LB v;
				do { v = *p; } while ((v & T::FLAGS) == 0);
				excl += (u64) (v & T::MASK);
				if (v & T::INC) break;
				t--;
			}
			LB *mine = lookback + (size_t) tile * SORT_RADIX + tid;
			*reinterpret_cast<volatile LB *>(mine) = T::INC | (LB) (excl + total);
		}
		s_dst[tid] = (i64) (goff[tid] + excl) - (i64) binstart;
	}
	__syncthreads();

	// ---- coalesced write-out: consecutive threads -> consecutive slots of a digit run
	const u32 count = full ? (u32) OS_TILE : (u32) (n - tile_base);
#pragma unroll 4
	for (u32 j = tid; j < count; j += OS_THREADS) {
		const u64 k = s_keys[j];
		const u32 d = (u32) (k >> shift) & dmask;
		const i64 g = s_dst[d] + (i64) j;
		st_stream_u64(out + g, k);
		if (HAS_VALS) vout[g] = s_vals[j];
	}
}

static size_t onesweep_smem(bool vals)
{
	return (size_t) OS_TILE * 8 + SORT_RADIX * 8 + (size_t) OS_WARPS * SORT_RADIX * 4 + SORT_RADIX * 4 +
	       (vals ? (size_t) OS_TILE * 4 : 0);
}

// ---- host side -------------------------------------------------------------------------------
int sort_workspace_reserve(metro_ctx *ctx, SortWorkspace &ws, i64 n)
{
	if (!ws.hist) {
		CUDA_TRY(ctx, cudaMalloc(&ws.hist, SORT_MAX_PASSES * SORT_RADIX * sizeof(u64)));
		CUDA_TRY(ctx, cudaMalloc(&ws.tile_counter, SORT_MAX_PASSES * sizeof(u32)));
		CUDA_TRY(ctx, cudaFuncSetAttribute(onesweep_kernel<false, u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) onesweep_smem(false)));
		CUDA_TRY(ctx, cudaFuncSetAttribute(onesweep_kernel<false, u64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) onesweep_smem(false)));
		CUDA_TRY(ctx, cudaFuncSetAttribute(onesweep_kernel<true, u32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) onesweep_smem(true)));
		CUDA_TRY(ctx, cudaFuncSetAttribute(onesweep_kernel<true, u64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) onesweep_smem(true)));
	}
	const i64 tiles = div_up<i64>(n > 0 ? n : 1, OS_TILE);
	const size_t word = n < (1ll << 30) ? 4 : 8;
	const size_t need = (size_t) tiles * SORT_RADIX * word;
	if (need > ws.lookback_bytes) {
		if (ws.lookback) cudaFree(ws.lookback);
		ws.lookback = nullptr; ws.lookback_bytes = 0;
		CUDA_TRY(ctx, cudaMalloc(&ws.lookback, need));
		ws.lookback_bytes = need;
	}
	return METRO_OK;
}

void sort_workspace_free(SortWorkspace &ws)
{
	if (ws.hist) cudaFree(ws.hist);
	if (ws.tile_counter) cudaFree(ws.tile_counter);
	if (ws.lookback) cudaFree(ws.lookback);
	ws = SortWorkspace();
}

int sort_clear_hist(metro_ctx *ctx, cudaStream_t st, SortWorkspace &ws)
{
	CUDA_TRY(ctx, cudaMemsetAsync(ws.hist, 0, SORT_MAX_PASSES * SORT_RADIX * sizeof(u64), st));
	return METRO_OK;
}

int sort_histogram(metro_ctx *ctx, cudaStream_t st, SortWorkspace &ws, const u64 *keys, i64 n, const SortPlan &plan)
{
	if (n <= 0 || plan.n_pass == 0) return METRO_OK;
	int blocks = (int) std::min<i64>(div_up<i64>(n, 2 * HIST_THREADS), (i64) METRO_NUM_SMS_B200 * 4);
	hist_kernel<<<blocks, HIST_THREADS, 0, st>>>(keys, (u64) n, plan, ws.hist);
	CUDA_TRY(ctx, cudaGetLastError());
	return METRO_OK;
}

int sort_scatter_passes(metro_ctx *ctx, cudaStream_t st, SortWorkspace &ws, u64 *a, u64 *b, u32 *va, u32 *vb,
			i64 n, const SortPlan &plan, u64 **result, u32 **vresult, cudaEvent_t *ev, int *launches)
{
	u64 *src = a, *dst = b;
	u32 *vsrc = va, *vdst = vb;
	const bool vals = va != nullptr;
	if (n > 0 && plan.n_pass > 0) {
		hist_scan_kernel<<<plan.n_pass, SORT_RADIX, 0, st>>>(ws.hist);
		CUDA_TRY(ctx, cudaGetLastError());
		if (launches) (*launches)++;
		CUDA_TRY(ctx, cudaMemsetAsync(ws.tile_counter, 0, SORT_MAX_PASSES * sizeof(u32), st));
		const i64 tiles = div_up<i64>(n, OS_TILE);
		const bool narrow = n < (1ll << 30);
		const size_t lb_bytes = (size_t) tiles * SORT_RADIX * (narrow ? 4 : 8);
		for (int p = 0; p < plan.n_pass; p++) {
			if (ev) CUDA_TRY(ctx, cudaEventRecord(ev[p], st));
			CUDA_TRY(ctx, cudaMemsetAsync(ws.lookback, 0, lb_bytes, st));
			const u64 *goff = ws.hist + (size_t) p * SORT_RADIX;
			const size_t smem = onesweep_smem(vals);
			if (vals) {
				if (narrow) onesweep_kernel<true, u32><<<(unsigned) tiles, OS_THREADS, smem, st>>>(src, dst, vsrc, vdst, (u64) n, plan.shift[p], plan.mask[p], goff, (u32 *) ws.lookback, ws.tile_counter + p);
				else onesweep_kernel<true, u64><<<(unsigned) tiles, OS_THREADS, smem, st>>>(src, dst, vsrc, vdst, (u64) n, plan.shift[p], plan.mask[p], goff, (u64 *) ws.lookback, ws.tile_counter + p);
			} else {
				if (narrow) onesweep_kernel<false, u32><<<(unsigned) tiles, OS_THREADS, smem, st>>>(src, dst, nullptr, nullptr, (u64) n, plan.shift[p], plan.mask[p], goff, (u32 *) ws.lookback, ws.tile_counter + p);
				else onesweep_kernel<false, u64><<<(unsigned) tiles, OS_THREADS, smem, st>>>(src, dst, nullptr, nullptr, (u64) n, plan.shift[p], plan.mask[p], goff, (u64 *) ws.lookback, ws.tile_counter + p);
			}
			CUDA_TRY(ctx, cudaGetLastError());
			if (launches) (*launches)++;
			u64 *t = src; src = dst; dst = t;
			u32 *tv = vsrc; vsrc = vdst; vdst = tv;
		}
		if (ev) CUDA_TRY(ctx, cudaEventRecord(ev[plan.n_pass], st));
	}
	*result = src;
	if (vresult) *vresult = vsrc;
	return METRO_OK;
}
