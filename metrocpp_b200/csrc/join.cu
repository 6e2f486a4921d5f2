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

static int pair_counts_begin(metro_ctx *ctx, PairCountState &pc, int attempt)
{
	cudaStream_t st = ctx->stream;
	const i64 H0 = ctx->cat[0].n_halos, H = H0 + ctx->cat[1].n_halos;
	if (attempt == 0) {
		pc.use_cache = H0 > 0 ? 1 : 0;               // metro_match_pair guarantees halo_bits + 2 type_bits <= 31
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
		int rc = pair_counts_begin(ctx, pc, attempt);
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
// Partition join: the fast path for large pairs.
//
// The join only needs equal particle IDs to MEET, not a total order.  So instead of P stable LSD
// passes plus a run-length scan, the keys are split MSD-style on a bijective mix of the pid:
//   pass A  pack (13 B/tuple in) + 2^b1-way partition          -> 2^b1 regions of fixed capacity
//   pass B  2^b2-way partition of every region                  -> 2^(b1+b2) buckets of ~8-15 k keys
//   pass C  per bucket: shared-memory hash table of the catalogue-1 members, probed by the
//           catalogue-0 members; every match is a hit -> pair counters
// mix(pid) is uniform whatever the ID distribution, so bucket sizes are mean +- a few sigma and
// the regions have FIXED capacity: no digit histograms, no look-back, no stability requirement.
// A pass ranks a key with one shared-memory atomic (the LSD pass needs 8 ballots + a serial
// shared-memory chain) and reserves bucket space with one global atomic per (tile, bucket).
// Any overflow (adversarial duplication of one pid) raises a flag and the caller falls back to the
// LSD sort path, which has no such limit.
// ================================================================================================
#define PT_THREADS 256
#define PT_ITEMS 16
#define PT_TILE (PT_THREADS * PT_ITEMS)

// bijection on pid_bits-bit integers (odd multiplies and xor-shifts are invertible mod 2^bits)
__device__ __forceinline__ u64 mix_pid(u64 x, const PartPlan &P)
{
	const u64 mask = P.pid_bits >= 64 ? ~0ull : ((1ull << P.pid_bits) - 1ull);
	x = (x * P.mul1) & mask;
	x ^= x >> P.xs;
	x = (x * P.mul2) & mask;
	x ^= x >> P.xs;
	return x;
}

template <int NB, bool PACK>
__global__ void __launch_bounds__(PT_THREADS, 3) part_kernel(const u64 *__restrict__ in, const u32 *__restrict__ in_halo, const u8 *__restrict__ in_type,
							     u64 n, u32 snap, u64 pid_base, KeyLayout L, PartPlan P, u64 *__restrict__ out,
							     u32 *gcount_in /* level 2: counts of the level-1 regions */,
							     u32 *gcount_out, u32 *overflow)
{
	extern __shared__ __align__(16) unsigned char psm[];
	u64 *s_keys = reinterpret_cast<u64 *>(psm);                          // [PT_TILE]
	i64 *s_dst = reinterpret_cast<i64 *>(s_keys + PT_TILE);              // [NB] global slot of a run minus its tile-local start
	u32 *s_cnt = reinterpret_cast<u32 *>(s_dst + NB);                   // [NB]
	u32 *s_start = s_cnt + NB;                                           // [NB]
	__shared__ u32 s_warpsum[PT_THREADS / 32];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	constexpr int BPT = NB / PT_THREADS > 0 ? NB / PT_THREADS : 1;       // buckets per thread in the scan
	for (int i = tid; i < NB; i += PT_THREADS) s_cnt[i] = 0;

	u64 base, count;            // this tile = [base, base + count) of the input stream
	u64 region = 0;             // level 2: which level-1 region
	const u64 *src = in;
	if (PACK) {
		base = (u64) blockIdx.x * PT_TILE;
		count = n - base < (u64) PT_TILE ? n - base : (u64) PT_TILE;
	} else {
		region = blockIdx.x / P.tiles1;                              // (level-1 bucket, replica)
		const u64 t = blockIdx.x % P.tiles1;
		const u64 have = gcount_in[region];
		base = t * PT_TILE;
		if (base >= have) return;
		count = have - base < (u64) PT_TILE ? have - base : (u64) PT_TILE;
		src = in + region * P.cap1;
	}
	__syncthreads();
	const int shift = PACK ? P.shift1 : P.shift2;
	const u64 wbase = base + (u64) warp * (PT_ITEMS * 32) + lane;
	u64 key[PT_ITEMS];
	u32 rk[PT_ITEMS];
	if (PACK) {
		// all loads first (plain streaming loads the compiler may hoist), then the arithmetic
		u32 ht[PT_ITEMS];
#pragma unroll
		for (int i = 0; i < PT_ITEMS; i++) {
			const u64 idx = wbase + 32 * i;
			const bool ok = idx < base + count;
			key[i] = ok ? __ldcs(src + idx) : 0ull;
			ht[i] = ok ? __ldcs(in_halo + idx) : 0u;
			rk[i] = ok ? (in_type ? (u32) __ldcs(in_type + idx) : 1u) : 0xFFFFFFFFu;     // type for now
		}
#pragma unroll
		for (int i = 0; i < PT_ITEMS; i++) {
			const bool ok = rk[i] != 0xFFFFFFFFu;
			key[i] = key_pack(L, mix_pid(key[i] - pid_base, P), snap, ht[i], rk[i] & L.type_mask);
			rk[i] = ok ? 0u : 0xFFFFFFFFu;
		}
	} else {
#pragma unroll
		for (int i = 0; i < PT_ITEMS; i++) {
			const u64 idx = wbase + 32 * i;
			const bool ok = idx < base + count;
			key[i] = ok ? ld_stream_u64(src + idx) : 0ull;
			rk[i] = ok ? 0u : 0xFFFFFFFFu;
		}
	}
#pragma unroll
	for (int i = 0; i < PT_ITEMS; i++) {
		const u32 d = (u32) (key[i] >> shift) & (NB - 1);
		if (rk[i] == 0u) rk[i] = atomicAdd(&s_cnt[d], 1u);
	}
	__syncthreads();
	// exclusive scan of the NB tile counts (thread t owns buckets [t*BPT, (t+1)*BPT))
	u32 loc[BPT];
	u32 sum = 0;
#pragma unroll
	for (int k = 0; k < BPT; k++) { loc[k] = (tid * BPT + k) < NB ? s_cnt[tid * BPT + k] : 0u; sum += loc[k]; }
	u32 incl = sum;
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		u32 v = __shfl_up_sync(0xffffffffu, incl, off);
		if (lane >= off) incl += v;
	}
	if (lane == 31) s_warpsum[warp] = incl;
	__syncthreads();
	u32 pre = 0;
#pragma unroll
	for (int w = 0; w < PT_THREADS / 32; w++) pre += (w < warp) ? s_warpsum[w] : 0u;
	u32 run = pre + incl - sum;
#pragma unroll
	for (int k = 0; k < BPT; k++) {
		const int b = tid * BPT + k;
		if (b < NB) {
			// pass A: region (bucket b, replica); pass B: final bucket (level-1 bucket, b) - the replicas merge here
			const u64 gb = PACK ? (u64) b * P.r1 + (blockIdx.x & (P.r1 - 1)) : (region / P.r1) * NB + (u64) b;
			const u64 cap = PACK ? P.cap1 : P.cap2;
			u32 g = loc[k] ? atomicAdd(&gcount_out[gb], loc[k]) : 0u;
			if ((u64) g + loc[k] > cap) { *overflow = 1; g = 0; }
			s_start[b] = run;
			s_dst[b] = (i64) (gb * cap + g) - (i64) run;
			run += loc[k];
		}
	}
	__syncthreads();
#pragma unroll
	for (int i = 0; i < PT_ITEMS; i++) {
		if (rk[i] == 0xFFFFFFFFu) continue;
		const u32 d = (u32) (key[i] >> shift) & (NB - 1);
		s_keys[s_start[d] + rk[i]] = key[i];
	}
	__syncthreads();
#pragma unroll 4
	for (u32 j = tid; j < (u32) count; j += PT_THREADS) {
		const u64 k = s_keys[j];
		const u32 d = (u32) (k >> shift) & (NB - 1);
		st_stream_u64(out + s_dst[d] + (i64) j, k);
	}
}

// pass C: one block per final bucket
#define BJ_THREADS 512
template <int SLOTS>
__global__ void __launch_bounds__(BJ_THREADS) bucket_join_kernel(const u64 *__restrict__ keys, const u32 *__restrict__ gcount, u64 cap2, KeyLayout L,
								 CacheWay *cache, int use_cache, PairEntry *tab, int log_cap, u64 *counters,
								 u32 *overflow)
{
	extern __shared__ __align__(16) unsigned char bsm[];
	u64 *slot = reinterpret_cast<u64 *>(bsm);                            // [SLOTS] catalogue-1 members, open addressing on the mixed pid
	const u64 g = blockIdx.x;
	const u32 n = gcount[g];
	if (n == 0) return;
	if ((u64) n > cap2) return;                                          // flagged by the partition pass already
	const u64 *src = keys + g * cap2;
	for (int i = threadIdx.x; i < SLOTS; i += BJ_THREADS) slot[i] = ~0ull;
	__syncthreads();
	// build
	for (u32 j = threadIdx.x; j < n; j += BJ_THREADS) {
		const u64 k = src[j];
		if (key_snap(L, k) == 0) continue;
		u32 h = (u32) key_pid(L, k) & (SLOTS - 1);
		int probe = 0;
		for (; probe < SLOTS; probe++) {
			const u64 old = atomicCAS((unsigned long long *) &slot[h], ~0ull, (unsigned long long) k);
			if (old == ~0ull) break;
			h = (h + 1) & (SLOTS - 1);
		}
		if (probe == SLOTS) *overflow = 1;                           // more catalogue-1 members than slots
	}
	__syncthreads();
	// probe
	u32 hits = 0, slow = 0;
	bool ovf = false;
	const int tsh = L.type_bits;
	for (u32 j = threadIdx.x; j < n; j += BJ_THREADS) {
		const u64 k = src[j];
		if (key_snap(L, k) != 0) continue;
		const u64 pid = key_pid(L, k);
		const u32 A = key_halo(L, k), tA = key_type(L, k);
		u32 h = (u32) pid & (SLOTS - 1);
		for (int probe = 0; probe < SLOTS; probe++) {
			const u64 e = slot[h];
			if (e == ~0ull) break;
			if (key_pid(L, e) == pid) {
				const u32 B = key_halo(L, e), tB = key_type(L, e);
				hits++;
				if (!(use_cache && cache_add(cache + (size_t) A * CACHE_WAYS, (B << (2 * tsh)) | (tA << tsh) | tB))) {
					const u64 hk = ((u64) A << (L.halo_bits + 2 * tsh)) | ((u64) B << (2 * tsh)) | ((u64) tA << tsh) | (u64) tB;
					slow++;
					if (!pair_add(tab, log_cap, hk)) ovf = true;
				}
			}
			h = (h + 1) & (SLOTS - 1);
		}
	}
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) { hits += __shfl_xor_sync(0xffffffffu, hits, off); slow += __shfl_xor_sync(0xffffffffu, slow, off); }
	if ((threadIdx.x & 31) == 0 && hits) atomicAdd((unsigned long long *) &counters[0], (unsigned long long) hits);
	if ((threadIdx.x & 31) == 0 && slow) atomicAdd((unsigned long long *) &counters[5], (unsigned long long) slow);
	if (ovf) counters[1] = 1;
}

static size_t part_smem(int nb) { return (size_t) PT_TILE * 8 + (size_t) nb * 16; }

// Is the partition path applicable, and with which bucket bits?
bool partition_join_plan(const metro_ctx *ctx, const KeyLayout &L, i64 N, PartPlan *out)
{
	static const i64 min_n = getenv("METRO_PART_MIN_N") ? atoll(getenv("METRO_PART_MIN_N")) : (1ll << 23);
	if (N < min_n || L.pid_bits < 16) return false;                 // small pairs: the LSD path is fine
	int tb = 0;
	while ((N >> tb) > 9000 && tb < 20) tb++;                      // final buckets of ~4.5-9 k keys
	PartPlan P;
	P.b1 = tb > 18 ? tb - 9 : (tb + 1) / 2;
	P.b2 = tb - P.b1;
	if (P.b1 < 8) P.b1 = 8;
	if (P.b2 < 8) P.b2 = 8;
	if (P.b1 > 10 || P.b2 > 10 || P.b1 + P.b2 > L.pid_bits) return false;
	P.pid_bits = L.pid_bits;
	P.shift1 = L.pid_shift + L.pid_bits - P.b1;
	P.shift2 = P.shift1 - P.b2;
	P.r1 = 8;
	const double m1 = (double) N / (double) ((1ll << P.b1) * P.r1), m2 = (double) N / (double) (1ll << (P.b1 + P.b2));
	P.cap1 = (u64) (m1 + 8.0 * sqrt(m1) + 4096.0);
	P.cap1 = (P.cap1 + 1) & ~1ull;                                  // keep regions 16-byte aligned
	P.cap2 = (u64) (m2 + 9.0 * sqrt(m2) + 64.0);
	P.cap2 = (P.cap2 + 1) & ~1ull;
	if (P.cap2 > 16384) return false;
	P.tiles1 = (u32) div_up<u64>(P.cap1, PT_TILE);
	P.mul1 = 0x9E3779B97F4A7C15ull; P.mul2 = 0xD6E8FEB86659FD93ull;
	P.xs = (L.pid_bits + 1) / 2;
	(void) ctx;
	*out = P;
	return true;
}

template <bool PACK>
static int launch_part(metro_ctx *ctx, int bits, unsigned grid, const u64 *in, const u32 *ih, const u8 *it, u64 n, u32 snap, const KeyLayout &L,
		       const PartPlan &P, u64 *out, u32 *gin, u32 *gout, u32 *ovf)
{
	cudaStream_t st = ctx->stream;
	const size_t smem = part_smem(1 << bits);
#define PART_CASE(NB)                                                                                                               \
	case NB: {                                                                                                                  \
		static bool attr = false;                                                                                           \
		if (!attr) { CUDA_TRY(ctx, cudaFuncSetAttribute(part_kernel<NB, PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); attr = true; } \
		part_kernel<NB, PACK><<<grid, PT_THREADS, smem, st>>>(in, ih, it, n, snap, ctx->pid_base, L, P, out, gin, gout, ovf);   \
		break;                                                                                                              \
	}
	switch (1 << bits) {
		PART_CASE(256) PART_CASE(512) PART_CASE(1024)
	default: metro_set_error(ctx, "unsupported partition width 2^%d", bits); return METRO_ERR_UNSUPPORTED;
	}
#undef PART_CASE
	CUDA_TRY(ctx, cudaGetLastError());
	ctx->launches++;
	return METRO_OK;
}

// Whole front end of a pair on the partition path.  *fell_back = true: an overflow flag was raised
// (nothing usable was produced; the caller runs the LSD path).  Timing events: ev[1] after pass A,
// ev[2] after pass B.
int stage_partition_join(metro_ctx *ctx, const KeyLayout &L, const PartPlan &P, i64 *n_records, i64 *n_hits, bool *fell_back)
{
	cudaStream_t st = ctx->stream;
	*fell_back = false;
	const u64 nb1 = (1ull << P.b1) * P.r1, nb = 1ull << (P.b1 + P.b2);     // level-1 regions, final buckets
	int rc;
	if ((rc = dev_reserve(ctx, ctx->keys_a, (size_t) nb1 * P.cap1 * 8 + 64))) return rc;
	if ((rc = dev_reserve(ctx, ctx->keys_b, (size_t) nb * P.cap2 * 8 + 64))) return rc;
	if ((rc = dev_reserve(ctx, ctx->part_counts, (size_t) (nb1 + nb + 16) * 4))) return rc;
	u32 *g1 = ctx->part_counts.as<u32>(), *g2 = g1 + nb1, *ovf = g2 + nb;
	CUDA_TRY(ctx, cudaMemsetAsync(g1, 0, (size_t) (nb1 + nb + 16) * 4, st));
	// pass A: both catalogues (catalogue 0 first; the order inside a bucket is irrelevant)
	for (int s = 0; s < 2; s++) {
		const DevCatalogue &c = ctx->cat[s];
		if (c.n_tuples == 0) continue;
		if ((rc = launch_part<true>(ctx, P.b1, (unsigned) div_up<i64>(c.n_tuples, PT_TILE), c.d_pid(), c.d_halo(), c.d_type(), (u64) c.n_tuples, (u32) s, L, P,
					    ctx->keys_a.as<u64>(), nullptr, g1, ovf))) return rc;
	}
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[1], st));
	// pass B
	if ((rc = launch_part<false>(ctx, P.b2, (unsigned) (nb1 * P.tiles1), ctx->keys_a.as<u64>(), nullptr, nullptr, 0, 0, L, P, ctx->keys_b.as<u64>(), g1, g2,
				     ovf))) return rc;
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[2], st));
	// pass C + pair counts
	PairCountState pc;
	// the table holds the catalogue-1 members only (about half the bucket); a bucket that is mostly
	// catalogue 1 overflows the table, raises the flag and the pair falls back to the LSD path
	const int slots = P.cap2 > 12288 ? 16384 : 8192;
	for (int attempt = 0;; attempt++) {
		if ((rc = pair_counts_begin(ctx, pc, attempt))) return rc;
		if (slots == 8192) {
			static bool attr = false;
			if (!attr) { CUDA_TRY(ctx, cudaFuncSetAttribute(bucket_join_kernel<8192>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8)); attr = true; }
			bucket_join_kernel<8192><<<(unsigned) nb, BJ_THREADS, 8192 * 8, st>>>(ctx->keys_b.as<u64>(), g2, P.cap2, L, ctx->pair_cache.as<CacheWay>(), pc.use_cache,
											     ctx->table.as<PairEntry>(), pc.log_cap, ctx->counters.as<u64>(), ovf);
		} else {
			static bool attr = false;
			if (!attr) { CUDA_TRY(ctx, cudaFuncSetAttribute(bucket_join_kernel<16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8)); attr = true; }
			bucket_join_kernel<16384><<<(unsigned) nb, BJ_THREADS, 16384 * 8, st>>>(ctx->keys_b.as<u64>(), g2, P.cap2, L, ctx->pair_cache.as<CacheWay>(), pc.use_cache,
											       ctx->table.as<PairEntry>(), pc.log_cap, ctx->counters.as<u64>(), ovf);
		}
		CUDA_TRY(ctx, cudaGetLastError());
		ctx->launches++;
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters + 60, ovf, 4, cudaMemcpyDeviceToHost, st));
		bool retry;
		if ((rc = pair_counts_finish(ctx, L, pc, attempt, &retry, n_records, n_hits))) return rc;
		if (*reinterpret_cast<u32 *>(ctx->h_counters + 60) != 0) { *fell_back = true; return METRO_OK; }
		if (!retry) return METRO_OK;
	}
}
