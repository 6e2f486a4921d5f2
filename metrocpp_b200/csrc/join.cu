// join.cu -- key packing (+ digit histograms), the segmented run-length join over the pid-sorted
// keys, and the pair-count reduction.
//
// Replaces the hot loop of FindProgenitors (reference src/MergerTree.cpp:443-472, run twice per
// pair from src/main.cpp:199,217): for every particle ID present in both catalogues the
// reference walks std::map nodes and bumps indexCommon[haloB][type] for each (haloA, haloB)
// membership pair.  Here the same multiset of hits comes from ONE pass over the sorted keys: a
// run of equal pid holds its catalogue-0 members first, then its catalogue-1 members; every
// (member0, member1) combination is a hit, counted in a hash table keyed by
// (haloA, haloB, typeA, typeB).  One table serves both directions: the forward tree reads the
// counts by typeB, the backward tree by typeA (the type is taken from the iTwo side, :461-467).
#include "context.cuh"

#define PK_THREADS 256
#define PK_ITEMS 4

// ---- pack + histogram ------------------------------------------------------------------------
// 13 B read + 8 B written per tuple; the P digit histograms ride along (no extra pass).
__global__ void __launch_bounds__(PK_THREADS) pack_hist_kernel(const u64 *__restrict__ pid, const u32 *__restrict__ halo,
							       const u8 *__restrict__ type, u64 n, u32 snap,
							       KeyLayout L, SortPlan plan, u64 *__restrict__ keys,
							       u64 *__restrict__ ghist)
{
	__shared__ u32 sh[SORT_MAX_PASSES][SORT_RADIX];
	for (int i = threadIdx.x; i < SORT_MAX_PASSES * SORT_RADIX; i += PK_THREADS) (&sh[0][0])[i] = 0;
	__syncthreads();
	const u64 step = (u64) gridDim.x * PK_THREADS * PK_ITEMS;
	for (u64 base = (u64) blockIdx.x * PK_THREADS * PK_ITEMS; base < n; base += step) {
		u64 p[PK_ITEMS];
		u32 h[PK_ITEMS], t[PK_ITEMS];
#pragma unroll
		for (int k = 0; k < PK_ITEMS; k++) {
			const u64 i = base + (u64) k * PK_THREADS + threadIdx.x;
			if (i < n) {
				p[k] = ld_stream_u64(pid + i);
				h[k] = __ldg(halo + i);
				t[k] = type ? (u32) __ldg(type + i) : 1u;
			}
		}
#pragma unroll
		for (int k = 0; k < PK_ITEMS; k++) {
			const u64 i = base + (u64) k * PK_THREADS + threadIdx.x;
			if (i < n) {
				const u64 key = key_pack(L, p[k], snap, h[k], t[k] & L.type_mask);
				st_stream_u64(keys + i, key);
				hist_accumulate(sh, key, plan);
			}
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < plan.n_pass * SORT_RADIX; i += PK_THREADS) {
		u32 c = (&sh[0][0])[i];
		if (c) atomicAdd((unsigned long long *) &ghist[i], (unsigned long long) c);
	}
}

int stage_pack_hist(metro_ctx *ctx, const KeyLayout &L, const SortPlan &plan)
{
	cudaStream_t st = ctx->stream;
	int rc = sort_clear_hist(ctx, st, ctx->sortws);
	if (rc) return rc;
	u64 *keys = ctx->keys_a.as<u64>();
	i64 off = 0;
	for (int s = 0; s < 2; s++) {
		const DevCatalogue &c = ctx->cat[s];
		if (c.n_tuples > 0) {
			int blocks = (int) std::min<i64>(div_up<i64>(c.n_tuples, PK_THREADS * PK_ITEMS), (i64) METRO_NUM_SMS_B200 * 8);
			pack_hist_kernel<<<blocks, PK_THREADS, 0, st>>>(c.d_pid(), c.d_halo(), c.d_type(), (u64) c.n_tuples, (u32) s, L, plan,
									 keys + off, ctx->sortws.hist);
			CUDA_TRY(ctx, cudaGetLastError());
			ctx->launches++;
		}
		off += c.n_tuples;
	}
	return METRO_OK;
}

// ---- pair-count hash table -------------------------------------------------------------------
struct __align__(16) PairEntry {
	u64 key;     // (A << (hb+2tb)) | (B << 2tb) | (tA << tb) | tB ; ~0 = empty
	u32 count;   // biased: stored = true count - 1 (the table is memset to 0xFF)
	u32 pad;
};
#define PAIR_EMPTY (~0ull)
#define PAIR_MAX_PROBE 1024

__device__ __forceinline__ u32 pair_hash(u64 k, int log_cap)
{
	return (u32) ((k * 0x9E3779B97F4A7C15ull) >> (64 - log_cap));
}

__device__ __forceinline__ bool pair_add(PairEntry *tab, int log_cap, u64 hk)
{
	const u32 mask = (1u << log_cap) - 1u;
	u32 h = pair_hash(hk, log_cap);
	for (int probe = 0; probe < PAIR_MAX_PROBE; probe++) {
		PairEntry *e = tab + h;
		u64 cur = *reinterpret_cast<volatile u64 *>(&e->key);
		if (cur == PAIR_EMPTY) {
			cur = atomicCAS((unsigned long long *) &e->key, (unsigned long long) PAIR_EMPTY, (unsigned long long) hk);
			if (cur == PAIR_EMPTY) cur = hk;
		}
		if (cur == hk) {
			atomicAdd(&e->count, 1u);
			return true;
		}
		h = (h + 1) & mask;
	}
	return false;
}

#define JN_THREADS 256
// One thread per sorted key.  A catalogue-0 member scans forward through the rest of its pid run
// and emits one hit per catalogue-1 member it meets (the run holds all 0-members first).
__global__ void __launch_bounds__(JN_THREADS) join_count_kernel(const u64 *__restrict__ keys, u64 n, KeyLayout L,
								PairEntry *tab, int log_cap, u64 *counters)
{
	const u64 i = (u64) blockIdx.x * JN_THREADS + threadIdx.x;
	u32 hits = 0;
	bool overflow = false;
	if (i < n) {
		const u64 k = keys[i];
		if (key_snap(L, k) == 0) {
			const u64 pid = key_pid(L, k);
			const u64 a_part = ((u64) key_halo(L, k) << (L.halo_bits + 2 * L.type_bits)) |
					   ((u64) key_type(L, k) << L.type_bits);
			for (u64 j = i + 1; j < n; j++) {
				const u64 k2 = keys[j];
				if (key_pid(L, k2) != pid) break;
				if (key_snap(L, k2) == 0) continue;
				const u64 hk = a_part | ((u64) key_halo(L, k2) << (2 * L.type_bits)) | (u64) key_type(L, k2);
				if (!pair_add(tab, log_cap, hk)) overflow = true;
				hits++;
			}
		}
	}
	// hit statistics: one atomic per warp
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, off);
	if ((threadIdx.x & 31) == 0 && hits) atomicAdd((unsigned long long *) &counters[0], (unsigned long long) hits);
	if (overflow) counters[1] = 1;
}

__global__ void table_compact_kernel(const PairEntry *__restrict__ tab, u64 cap, u64 *__restrict__ rec_key,
				     u32 *__restrict__ rec_cnt, u64 *counters)
{
	const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
	u64 key = PAIR_EMPTY;
	u32 cnt = 0;
	if (i < cap) { key = tab[i].key; cnt = tab[i].count + 1u; }
	const bool live = key != PAIR_EMPTY;
	const u32 ballot = __ballot_sync(0xffffffffu, live);
	if (ballot == 0) return;
	const int lane = threadIdx.x & 31;
	u64 base = 0;
	if (lane == 0) base = atomicAdd((unsigned long long *) &counters[2], (unsigned long long) __popc(ballot));
	base = __shfl_sync(0xffffffffu, base, 0);
	if (live) {
		const u64 dst = base + __popc(ballot & lanemask_lt());
		rec_key[dst] = key;
		rec_cnt[dst] = cnt;
	}
}

__global__ void table_count_kernel(const PairEntry *__restrict__ tab, u64 cap, u64 *counters)
{
	const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
	const bool live = i < cap && tab[i].key != PAIR_EMPTY;
	const u32 ballot = __ballot_sync(0xffffffffu, live);
	if ((threadIdx.x & 31) == 0 && ballot) atomicAdd((unsigned long long *) &counters[3], (unsigned long long) __popc(ballot));
}

// scratch[0] = rec_key (u64[R]), scratch[1] = rec_cnt (u32[R])
int stage_join_count(metro_ctx *ctx, const KeyLayout &L, const u64 *sorted, i64 n, i64 *n_records, i64 *n_hits)
{
	cudaStream_t st = ctx->stream;
	const i64 H = ctx->cat[0].n_halos + ctx->cat[1].n_halos;
	int log_cap = 16;
	while ((1ll << log_cap) < 8 * H && log_cap < 40) log_cap++;
	u64 *counters = ctx->counters.as<u64>();
	for (int attempt = 0;; attempt++) {
		const u64 cap = 1ull << log_cap;
		int rc = dev_reserve(ctx, ctx->table, cap * sizeof(PairEntry));
		if (rc) return rc;
		CUDA_TRY(ctx, cudaMemsetAsync(ctx->table.p, 0xFF, cap * sizeof(PairEntry), st));
		CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, 8 * sizeof(u64), st));
		if (n > 0) {
			join_count_kernel<<<(unsigned) div_up<i64>(n, JN_THREADS), JN_THREADS, 0, st>>>(sorted, (u64) n, L, ctx->table.as<PairEntry>(),
													log_cap, counters);
			CUDA_TRY(ctx, cudaGetLastError());
			ctx->launches++;
		}
		table_count_kernel<<<(unsigned) div_up<u64>(cap, 256), 256, 0, st>>>(ctx->table.as<PairEntry>(), cap, counters);
		CUDA_TRY(ctx, cudaGetLastError());
		ctx->launches++;
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters, counters, 8 * sizeof(u64), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(ctx, cudaStreamSynchronize(st));
		const bool overflow = ctx->h_counters[1] != 0;
		const u64 live = ctx->h_counters[3];
		if (!overflow) {
			*n_hits = (i64) ctx->h_counters[0];
			*n_records = (i64) live;
			rc = dev_reserve(ctx, ctx->scratch[0], (size_t) (live + 1) * sizeof(u64));
			if (rc) return rc;
			rc = dev_reserve(ctx, ctx->scratch[1], (size_t) (live + 1) * sizeof(u32));
			if (rc) return rc;
			table_compact_kernel<<<(unsigned) div_up<u64>(cap, 256), 256, 0, st>>>(ctx->table.as<PairEntry>(), cap, ctx->scratch[0].as<u64>(),
											       ctx->scratch[1].as<u32>(), counters);
			CUDA_TRY(ctx, cudaGetLastError());
			ctx->launches++;
			return METRO_OK;
		}
		if (attempt >= 6 || log_cap >= 38) {
			metro_set_error(ctx, "pair table overflow at 2^%d entries", log_cap);
			return METRO_ERR_UNSUPPORTED;
		}
		log_cap += 2;       // too many distinct (A,B) pairs: grow x4 and redo the join pass
	}
}
