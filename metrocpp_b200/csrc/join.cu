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
#include <stdlib.h>

#include "context.cuh"

#define PK_THREADS 256
#define PK_ITEMS 4

// ---- pack + histogram ------------------------------------------------------------------------
// 13 B read + 8 B written per tuple; the P digit histograms ride along (no extra pass).
__global__ void __launch_bounds__(PK_THREADS) pack_hist_kernel(const u64 *__restrict__ pid, const u32 *__restrict__ halo,
							       const u8 *__restrict__ type, u64 n, u32 snap, u64 pid_base,
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
				const u64 key = key_pack(L, p[k] - pid_base, snap, h[k], t[k] & L.type_mask);
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
			pack_hist_kernel<<<blocks, PK_THREADS, 0, st>>>(c.d_pid(), c.d_halo(), c.d_type(), (u64) c.n_tuples, (u32) s, ctx->pid_base, L, plan,
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

__device__ __forceinline__ u64 pair_hash(u64 k, int log_cap)
{
	return (k * 0x9E3779B97F4A7C15ull) >> (64 - log_cap);          // log_cap may exceed 32 (the table grows x4 per retry)
}

__device__ __forceinline__ bool pair_add(PairEntry *tab, int log_cap, u64 hk)
{
	const u64 mask = (1ull << log_cap) - 1ull;
	u64 h = pair_hash(hk, log_cap);
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

// ---- per-halo direct cache ---------------------------------------------------------------------
// Almost every hit of a halo A goes to one or two partner halos B, so 4 (tag, count) ways per A
// (one 32-byte sector, H0 x 32 B: L2 resident) absorb nearly all hits; only the overflow goes to
// the general hash table above, which lives in DRAM.  tag = (B << 2tb) | (tA << tb) | tB.
struct __align__(8) CacheWay { u32 tag; u32 count; };      // count biased by -1 like PairEntry
#define CACHE_WAYS 4
#define CACHE_EMPTY 0xFFFFFFFFu

__device__ __forceinline__ bool cache_add(CacheWay *c, u32 tag)
{
#pragma unroll
	for (int w = 0; w < CACHE_WAYS; w++) {
		u32 cur = *reinterpret_cast<volatile u32 *>(&c[w].tag);
		if (cur == CACHE_EMPTY) {
			cur = atomicCAS(&c[w].tag, CACHE_EMPTY, tag);
			if (cur == CACHE_EMPTY) cur = tag;
		}
		if (cur == tag) { atomicAdd(&c[w].count, 1u); return true; }
	}
	return false;
}

#define JN_THREADS 256
#define JN_ITEMS 8
#define JN_TILE (JN_THREADS * JN_ITEMS)
#define JN_HALO 32
// Segmented run-length join.  A block stages a tile of sorted keys (+ a look-ahead halo) in shared
// memory with coalesced 16-byte loads; every catalogue-0 member then walks the rest of its pid run
// (all 0-members of a run precede its 1-members) and counts one hit per 1-member.
__global__ void __launch_bounds__(JN_THREADS) join_count_kernel(const u64 *__restrict__ keys, u64 n, KeyLayout L,
								CacheWay *cache, int use_cache, PairEntry *tab, int log_cap,
								u64 *counters)
{
	__shared__ __align__(16) u64 s[JN_TILE + JN_HALO];
	const u64 base = (u64) blockIdx.x * JN_TILE;
	const u64 avail = n - base < (u64) (JN_TILE + JN_HALO) ? n - base : (u64) (JN_TILE + JN_HALO);
	// base is a multiple of 2048 keys => 16-byte aligned
	const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(keys + base);
	for (u32 j = threadIdx.x; j < (JN_TILE + JN_HALO) / 2; j += JN_THREADS) {
		if ((u64) (2 * j + 1) < avail) {
			const ulonglong2 v = __ldg(src + j);
			s[2 * j] = v.x; s[2 * j + 1] = v.y;
		} else if ((u64) (2 * j) < avail) {
			s[2 * j] = keys[base + 2 * j];
		}
	}
	__syncthreads();
	const u32 in_tile = avail < (u64) JN_TILE ? (u32) avail : (u32) JN_TILE;
	u32 hits = 0, slow = 0;
	bool overflow = false;
	const int tsh = L.type_bits;
#pragma unroll 2
	for (int it = 0; it < JN_ITEMS; it++) {
		const u32 j = it * JN_THREADS + threadIdx.x;
		if (j >= in_tile) break;
		const u64 k = s[j];
		if (key_snap(L, k) != 0) continue;
		const u64 pid = key_pid(L, k);
		const u32 A = key_halo(L, k), tA = key_type(L, k);
		for (u64 t = j + 1; base + t < n; t++) {
			const u64 k2 = t < (u64) (JN_TILE + JN_HALO) ? s[t] : keys[base + t];
			if (key_pid(L, k2) != pid) break;
			if (key_snap(L, k2) == 0) continue;
			const u32 B = key_halo(L, k2), tB = key_type(L, k2);
			hits++;
			if (use_cache && cache_add(cache + (size_t) A * CACHE_WAYS, (B << (2 * tsh)) | (tA << tsh) | tB)) continue;
			const u64 hk = ((u64) A << (L.halo_bits + 2 * tsh)) | ((u64) B << (2 * tsh)) | ((u64) tA << tsh) | (u64) tB;
			slow++;
			if (!pair_add(tab, log_cap, hk)) overflow = true;
		}
	}
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) { hits += __shfl_xor_sync(0xffffffffu, hits, off); slow += __shfl_xor_sync(0xffffffffu, slow, off); }
	if ((threadIdx.x & 31) == 0 && hits) atomicAdd((unsigned long long *) &counters[0], (unsigned long long) hits);
	if ((threadIdx.x & 31) == 0 && slow) atomicAdd((unsigned long long *) &counters[5], (unsigned long long) slow);
	if (overflow) counters[1] = 1;
}

// live entries of the hash table (mode 0: count into counters[3]; mode 1: append records)
__global__ void table_compact_kernel(const PairEntry *__restrict__ tab, u64 cap, int mode, u64 *__restrict__ rec_key,
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
	u64 b = 0;
	if (lane == 0) b = atomicAdd((unsigned long long *) &counters[mode ? 2 : 3], (unsigned long long) __popc(ballot));
	if (!mode) return;
	b = __shfl_sync(0xffffffffu, b, 0);
	if (live) {
		const u64 dst = b + __popc(ballot & lanemask_lt());
		rec_key[dst] = key;
		rec_cnt[dst] = cnt;
	}
}

// live ways of the direct cache, same two modes; rebuilds the 64-bit hit key
__global__ void cache_compact_kernel(const CacheWay *__restrict__ cache, u64 n_ways, int mode, int halo_bits, int type_bits,
				     u64 *__restrict__ rec_key, u32 *__restrict__ rec_cnt, u64 *counters)
{
	const u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x;
	u32 tag = CACHE_EMPTY, cnt = 0;
	if (i < n_ways) { tag = cache[i].tag; cnt = cache[i].count + 1u; }
	const bool live = tag != CACHE_EMPTY;
	const u32 ballot = __ballot_sync(0xffffffffu, live);
	if (ballot == 0) return;
	const int lane = threadIdx.x & 31;
	u64 b = 0;
	if (lane == 0) b = atomicAdd((unsigned long long *) &counters[mode ? 2 : 3], (unsigned long long) __popc(ballot));
	if (!mode) return;
	b = __shfl_sync(0xffffffffu, b, 0);
	if (live) {
		const u64 dst = b + __popc(ballot & lanemask_lt());
		rec_key[dst] = ((i / CACHE_WAYS) << (halo_bits + 2 * type_bits)) | (u64) tag;
		rec_cnt[dst] = cnt;
	}
}

// ---- host side of the pair counting ------------------------------------------------------------
struct PairCountState { int log_cap; u64 cap; u64 n_ways; int use_cache; };

static int pair_counts_begin(metro_ctx *ctx, const KeyLayout &L, PairCountState &pc, int attempt)
{
	cudaStream_t st = ctx->stream;
	const i64 H0 = ctx->cat[0].n_halos, H = H0 + ctx->cat[1].n_halos;
	if (attempt == 0) {
		pc.use_cache = (H0 > 0 && L.halo_bits + 2 * L.type_bits <= 31) ? 1 : 0;   // the way tag (B, tA, tB) must fit 31 bits
		pc.n_ways = pc.use_cache ? (u64) H0 * CACHE_WAYS : 0;
		pc.log_cap = 16;
		while ((1ll << pc.log_cap) < 2 * H && pc.log_cap < 40) pc.log_cap++;
	} else {
		pc.log_cap += 2;                               // too many distinct (A,B) pairs: grow x4 and redo the pass
	}
	pc.cap = 1ull << pc.log_cap;
	int rc;
	if (pc.use_cache && (rc = dev_reserve(ctx, ctx->pair_cache, pc.n_ways * sizeof(CacheWay)))) return rc;
	if ((rc = dev_reserve(ctx, ctx->table, pc.cap * sizeof(PairEntry)))) return rc;
	CUDA_TRY(ctx, cudaMemsetAsync(ctx->table.p, 0xFF, pc.cap * sizeof(PairEntry), st));
	if (pc.use_cache) CUDA_TRY(ctx, cudaMemsetAsync(ctx->pair_cache.p, 0xFF, pc.n_ways * sizeof(CacheWay), st));
	CUDA_TRY(ctx, cudaMemsetAsync(ctx->counters.as<u64>(), 0, 8 * sizeof(u64), st));
	return METRO_OK;
}

// returns METRO_OK with *retry = true when the hash table overflowed (caller redoes the pass)
static int pair_counts_finish(metro_ctx *ctx, const KeyLayout &L, PairCountState &pc, int attempt, bool *retry, i64 *n_records, i64 *n_hits)
{
	cudaStream_t st = ctx->stream;
	u64 *counters = ctx->counters.as<u64>();
	int rc;
	*retry = false;
	table_compact_kernel<<<(unsigned) div_up<u64>(pc.cap, 256), 256, 0, st>>>(ctx->table.as<PairEntry>(), pc.cap, 0, nullptr, nullptr, counters);
	CUDA_TRY(ctx, cudaGetLastError());
	ctx->launches++;
	if (pc.use_cache) {
		cache_compact_kernel<<<(unsigned) div_up<u64>(pc.n_ways, 256), 256, 0, st>>>(ctx->pair_cache.as<CacheWay>(), pc.n_ways, 0, L.halo_bits, L.type_bits,
											      nullptr, nullptr, counters);
		CUDA_TRY(ctx, cudaGetLastError());
		ctx->launches++;
	}
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters, counters, 8 * sizeof(u64), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	if (ctx->h_counters[1] != 0) {
		if (attempt >= 8 || pc.log_cap >= 38) {
			metro_set_error(ctx, "pair table overflow at 2^%d entries", pc.log_cap);
			return METRO_ERR_UNSUPPORTED;
		}
		*retry = true;
		return METRO_OK;
	}
	const u64 live = ctx->h_counters[3];
	*n_hits = (i64) ctx->h_counters[0];
	ctx->res.info.n_table_hits = (i64) ctx->h_counters[5];
	*n_records = (i64) live;
	if ((rc = dev_reserve(ctx, ctx->scratch[0], (size_t) (live + 1) * sizeof(u64)))) return rc;
	if ((rc = dev_reserve(ctx, ctx->scratch[1], (size_t) (live + 1) * sizeof(u32)))) return rc;
	table_compact_kernel<<<(unsigned) div_up<u64>(pc.cap, 256), 256, 0, st>>>(ctx->table.as<PairEntry>(), pc.cap, 1, ctx->scratch[0].as<u64>(),
									       ctx->scratch[1].as<u32>(), counters);
	CUDA_TRY(ctx, cudaGetLastError());
	ctx->launches++;
	if (pc.use_cache) {
		cache_compact_kernel<<<(unsigned) div_up<u64>(pc.n_ways, 256), 256, 0, st>>>(ctx->pair_cache.as<CacheWay>(), pc.n_ways, 1, L.halo_bits, L.type_bits,
											      ctx->scratch[0].as<u64>(), ctx->scratch[1].as<u32>(), counters);
		CUDA_TRY(ctx, cudaGetLastError());
		ctx->launches++;
	}
	return METRO_OK;
}

// scratch[0] = rec_key (u64[R]), scratch[1] = rec_cnt (u32[R])
int stage_join_count(metro_ctx *ctx, const KeyLayout &L, const u64 *sorted, i64 n, i64 *n_records, i64 *n_hits)
{
	cudaStream_t st = ctx->stream;
	PairCountState pc;
	for (int attempt = 0;; attempt++) {
		int rc = pair_counts_begin(ctx, L, pc, attempt);
		if (rc) return rc;
		if (n > 0) {
			join_count_kernel<<<(unsigned) div_up<i64>(n, JN_TILE), JN_THREADS, 0, st>>>(sorted, (u64) n, L, ctx->pair_cache.as<CacheWay>(), pc.use_cache,
												      ctx->table.as<PairEntry>(), pc.log_cap, ctx->counters.as<u64>());
			CUDA_TRY(ctx, cudaGetLastError());
			ctx->launches++;
		}
		bool retry;
		if ((rc = pair_counts_finish(ctx, L, pc, attempt, &retry, n_records, n_hits))) return rc;
		if (!retry) return METRO_OK;
	}
}

// ================================================================================================
// Wide-key path: particle IDs so wide (hashed 64-bit IDs) that pid + snap + halo + type do not fit
// one 64-bit word.  The tuples are sorted as (key = pid, value = [snap | halo | type]) pairs with the
// key/value radix sort (8 passes at most, stable, catalogue 0 packed first), then joined by walking
// the runs of equal pid.  Same records out as the other two front ends; slower, but no input the
// reference accepts is refused.
// ================================================================================================
__global__ void __launch_bounds__(256) pack_wide_kernel(const u64 *__restrict__ pid, const u32 *__restrict__ halo, const u8 *__restrict__ type, u64 n, u32 snap,
							 u64 pid_base, int type_bits, u32 type_mask, u64 *__restrict__ key, u32 *__restrict__ val)
{
	const u64 i = (u64) blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	const u32 t = type ? (u32) type[i] : 1u;
	key[i] = pid[i] - pid_base;
	val[i] = (snap << 31) | (halo[i] << type_bits) | (t & type_mask);
}

__global__ void __launch_bounds__(256) join_count_wide_kernel(const u64 *__restrict__ key, const u32 *__restrict__ val, u64 n, KeyLayout L, CacheWay *cache,
							       int use_cache, PairEntry *tab, int log_cap, u64 *counters)
{
	const u64 i = (u64) blockIdx.x * 256 + threadIdx.x;
	u32 hits = 0, slow = 0;
	bool overflow = false;
	if (i < n && (val[i] >> 31) == 0) {
		const u64 pid = key[i];
		const int tsh = L.type_bits;
		const u32 A = (val[i] & 0x7FFFFFFFu) >> tsh, tA = val[i] & L.type_mask;
		for (u64 t = i + 1; t < n && key[t] == pid; t++) {
			const u32 v = val[t];
			if ((v >> 31) == 0) continue;
			const u32 B = (v & 0x7FFFFFFFu) >> tsh, tB = v & L.type_mask;
			hits++;
			if (use_cache && cache_add(cache + (size_t) A * CACHE_WAYS, (B << (2 * tsh)) | (tA << tsh) | tB)) continue;
			const u64 hk = ((u64) A << (L.halo_bits + 2 * tsh)) | ((u64) B << (2 * tsh)) | ((u64) tA << tsh) | (u64) tB;
			slow++;
			if (!pair_add(tab, log_cap, hk)) overflow = true;
		}
	}
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) { hits += __shfl_xor_sync(0xffffffffu, hits, off); slow += __shfl_xor_sync(0xffffffffu, slow, off); }
	if ((threadIdx.x & 31) == 0 && hits) atomicAdd((unsigned long long *) &counters[0], (unsigned long long) hits);
	if ((threadIdx.x & 31) == 0 && slow) atomicAdd((unsigned long long *) &counters[5], (unsigned long long) slow);
	if (overflow) counters[1] = 1;
}

// scratch[0] = rec_key (u64[R]), scratch[1] = rec_cnt (u32[R]); L carries halo_bits / type_bits only (pid_bits is unused)
int stage_wide_join(metro_ctx *ctx, const KeyLayout &L, int pid_bits, i64 *n_records, i64 *n_hits, int *sort_launches)
{
	cudaStream_t st = ctx->stream;
	const i64 N = ctx->cat[0].n_tuples + ctx->cat[1].n_tuples;
	int rc;
	if ((rc = dev_reserve(ctx, ctx->keys_a, (size_t) N * 8 + 16))) return rc;
	if ((rc = dev_reserve(ctx, ctx->keys_b, (size_t) N * 8 + 16))) return rc;
	if ((rc = dev_reserve(ctx, ctx->scratch[2], (size_t) N * 4 + 16))) return rc;
	if ((rc = dev_reserve(ctx, ctx->scratch[3], (size_t) N * 4 + 16))) return rc;
	if ((rc = sort_workspace_reserve(ctx, ctx->sortws, N))) return rc;
	u64 *ka = ctx->keys_a.as<u64>(), *kb = ctx->keys_b.as<u64>();
	u32 *va = ctx->scratch[2].as<u32>(), *vb = ctx->scratch[3].as<u32>();
	i64 off = 0;
	for (int s = 0; s < 2; s++) {
		const DevCatalogue &c = ctx->cat[s];
		if (c.n_tuples > 0) {
			pack_wide_kernel<<<(unsigned) div_up<i64>(c.n_tuples, 256), 256, 0, st>>>(c.d_pid(), c.d_halo(), c.d_type(), (u64) c.n_tuples, (u32) s, ctx->pid_base,
												 L.type_bits, L.type_mask, ka + off, va + off);
			CUDA_TRY(ctx, cudaGetLastError());
			ctx->launches++;
		}
		off += c.n_tuples;
	}
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[1], st));
	const SortPlan plan = make_sort_plan(0, pid_bits);
	u64 *sk = ka; u32 *sv = va;
	if (N > 0) {
		if ((rc = sort_clear_hist(ctx, st, ctx->sortws))) return rc;
		if ((rc = sort_histogram(ctx, st, ctx->sortws, ka, N, plan))) return rc;
		ctx->launches++;
		if ((rc = sort_scatter_passes(ctx, st, ctx->sortws, ka, kb, va, vb, N, plan, &sk, &sv, nullptr, &ctx->launches))) return rc;
	}
	*sort_launches = N > 0 ? plan.n_pass : 0;
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[2], st));
	PairCountState pc;
	for (int attempt = 0;; attempt++) {
		if ((rc = pair_counts_begin(ctx, L, pc, attempt))) return rc;
		if (N > 0) {
			join_count_wide_kernel<<<(unsigned) div_up<i64>(N, 256), 256, 0, st>>>(sk, sv, (u64) N, L, ctx->pair_cache.as<CacheWay>(), pc.use_cache,
											      ctx->table.as<PairEntry>(), pc.log_cap, ctx->counters.as<u64>());
			CUDA_TRY(ctx, cudaGetLastError());
			ctx->launches++;
		}
		bool retry;
		if ((rc = pair_counts_finish(ctx, L, pc, attempt, &retry, n_records, n_hits))) return rc;
		if (!retry) return METRO_OK;
	}
}
