// partjoin.cu -- partition join: the front end for large pairs.
//
// Replaces the hot loop of FindProgenitors (reference src/MergerTree.cpp:443-472, run twice per pair
// from src/main.cpp:199,217).  The join only needs equal particle IDs to MEET, not a total order, and
// the pair counts only need equal (haloA, haloB) to meet.  So the pipeline is two shuffles, each a
// streaming partition with shared-memory work in between, and NO global atomics per tuple or per hit:
//
//   pass A  pack (12-13 B/tuple in) + 2^b1-way partition on mix(pid)   -> 2^b1 x r1 regions of fixed capacity
//   pass B  2^b2-way partition of every region                          -> 2^(b1+b2) buckets of 4-8 k keys
//   pass C  per bucket: shared-memory hash table of the catalogue-1 members, probed by the catalogue-0
//           members; every match is a hit record (A, B, tA, tB), partitioned in shared memory by halo-A
//           range and appended to that range's region (8 B/hit out)
//   pass D  per halo-A range: the hits are counted in shared memory (WAYS (tag,count) slots per halo A,
//           direct mapped) and leave as (record, count) pairs - what the selection stage consumes
//
// mix(pid) is a bijection that spreads any ID distribution uniformly, so the key buckets have FIXED
// capacity (mean + 9 sigma): no digit histograms, no look-back, no stability requirement.  The hit
// regions are sized from a sampled count of the catalogue-0 tuples per halo range (every hit belongs to
// one catalogue-0 tuple; 2 hits per tuple + slack for the sampling error are provisioned).  Any overflow (adversarial
// duplication of one pid, > 2 memberships per particle on average over a whole range) raises a flag and
// the caller falls back to the LSD sort path of join.cu, which has no such limit.
#include <stdlib.h>

#include "context.cuh"

#ifndef PT_THREADS
#define PT_THREADS 256
#endif
#ifndef PT_ITEMS
#define PT_ITEMS 16
#endif
#ifndef PT_MINB
#define PT_MINB 3
#endif
#define PT_TILE (PT_THREADS * PT_ITEMS)
// Fill counters that many blocks bump with global atomics live one per 128-byte line: the L2 atomic
// unit serialises operations on the same line, not only on the same word.
#define CSTRIDE 32
#define HIST_SAMPLE 8

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
__global__ void __launch_bounds__(PT_THREADS, PT_MINB) part_kernel(const u64 *__restrict__ in, const u32 *__restrict__ in_halo, const u8 *__restrict__ in_type,
							     u64 n, u32 snap, u64 pid_base, KeyLayout L, PartPlan P, u64 *__restrict__ out,
							     u32 *gcount_in /* level 2: counts of the level-1 regions */,
							     u32 *gcount_out, u32 *overflow,
							     u32 *t0 /* pass A, catalogue 0: tuples per hit partition (may be NULL) */, int log_hp, u32 nbh)
{
	extern __shared__ __align__(16) unsigned char psm[];
	u64 *s_keys = reinterpret_cast<u64 *>(psm);                          // [PT_TILE]
	i64 *s_dst = reinterpret_cast<i64 *>(s_keys + PT_TILE);              // [NB] global slot of a run minus its tile-local start
	u32 *s_cnt = reinterpret_cast<u32 *>(s_dst + NB);                   // [NB]
	u32 *s_start = s_cnt + NB;                                           // [NB]
	u32 *s_hist = s_start + NB;                                          // [nbh] pass A, catalogue 0 only
	__shared__ u32 s_warpsum[PT_THREADS / 32];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	constexpr int BPT = NB / PT_THREADS > 0 ? NB / PT_THREADS : 1;       // buckets per thread in the scan
	for (int i = tid; i < NB; i += PT_THREADS) s_cnt[i] = 0;
	// the tuples per hit range only size the hit regions (2 hits per tuple + slack): every HIST_SAMPLE-th tile is counted and
	// scaled, which is exact to within a tile per range; the slack covers that
	const bool hist = PACK && t0 != nullptr && (blockIdx.x % HIST_SAMPLE) == 0;
	if (hist) for (u32 i = tid; i < nbh; i += PT_THREADS) s_hist[i] = 0;

	u64 base, count;            // this tile = [base, base + count) of the input stream
	u64 region = 0;             // level 2: which level-1 region
	const u64 *src = in;
	if (PACK) {
		base = (u64) blockIdx.x * PT_TILE;
		count = n - base < (u64) PT_TILE ? n - base : (u64) PT_TILE;
	} else {
		region = blockIdx.x / P.tiles1;                              // (level-1 bucket, replica)
		const u64 t = blockIdx.x % P.tiles1;
		const u64 have = gcount_in[region * CSTRIDE];
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
		if (hist) {
			// catalogue-0 tuples per hit partition: the 32 consecutive tuples of a warp load normally sit in one halo range
#pragma unroll
			for (int i = 0; i < PT_ITEMS; i++) {
				const bool ok = rk[i] != 0xFFFFFFFFu;
				const u32 d = ht[i] >> log_hp;
				const u32 d0 = __shfl_sync(0xffffffffu, d, 0);
				const u32 okm = __ballot_sync(0xffffffffu, ok);
				if (__all_sync(0xffffffffu, !ok || d == d0)) {
					if (lane == 0 && okm && d0 < nbh) atomicAdd(&s_hist[d0], (u32) __popc(okm));
				} else if (ok && d < nbh) atomicAdd(&s_hist[d], 1u);
			}
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
	// region space is reserved with one global atomic per (tile, bucket); their latency (the L2 round trip under
	// contention) is hidden behind the shared-memory reorder, which only needs the tile-local offsets
	u32 g[BPT];
#pragma unroll
	for (int k = 0; k < BPT; k++) {
		const int b = tid * BPT + k;
		g[k] = 0;
		if (b < NB) {
			// pass A: region (catalogue, bucket b, replica); pass B: final bucket (catalogue, level-1 bucket, b) - the replicas merge here
			const u64 gb = PACK ? (((u64) snap << P.b1) + (u64) b) * P.r1 + (blockIdx.x & (P.r1 - 1)) : (region / P.r1) * NB + (u64) b;
			if (loc[k]) g[k] = atomicAdd(&gcount_out[PACK ? gb * CSTRIDE : gb], loc[k]);
			s_start[b] = run;
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
#pragma unroll
	for (int k = 0; k < BPT; k++) {
		const int b = tid * BPT + k;
		if (b < NB) {
			const u64 gb = PACK ? (((u64) snap << P.b1) + (u64) b) * P.r1 + (blockIdx.x & (P.r1 - 1)) : (region / P.r1) * NB + (u64) b;
			const u64 cap = PACK ? P.cap1 : P.cap2;
			u32 g0 = g[k];
			if ((u64) g0 + loc[k] > cap) { *overflow = 1; g0 = 0; }
			s_dst[b] = (i64) (gb * cap + g0) - (i64) s_start[b];
		}
	}
	__syncthreads();
#pragma unroll 4
	for (u32 j = tid; j < (u32) count; j += PT_THREADS) {
		const u64 k = s_keys[j];
		const u32 d = (u32) (k >> shift) & (NB - 1);
		st_stream_u64(out + s_dst[d] + (i64) j, k);
	}
	if (hist)
		for (u32 i = tid; i < nbh; i += PT_THREADS) {
			const u32 c = s_hist[i];
			if (c) atomicAdd(&t0[i * CSTRIDE], c * HIST_SAMPLE);
		}
}

// ---- hit regions ------------------------------------------------------------------------------
// A hit record is the 64-bit pair key the selection stage consumes:
//   rec = (A << (halo_bits + 2 type_bits)) | tag,   tag = (B << 2 type_bits) | (tA << type_bits) | tB
// Hit partition d = A >> log_hp (HP = 2^log_hp catalogue-0 halos, the range one pass-D block counts).
struct HitPlan {
	int rec_a_shift;            // halo_bits + 2 type_bits
	int log_hp;                 // halos per hit partition (log2)
	int ways;                   // (tag,count) slots per halo in pass D
	int log_hs;                 // halos one pass-D block counts (<= log_hp): a partition is read 2^(log_hp-log_hs) times,
	                            // each block keeping the hits of its sub-range (many halos per GPU, e.g. a pid shard of a big box)
	u32 nbh;                    // hit partitions
	u32 slack;                  // fixed part of a region's capacity
};
#define HIT_MAX_PARTS 1024
#define BD_DIRECT_RB 13          // residual pid bits of a bucket for which pass C's direct-address table is 32 KB
#define HIT_EMPTY 0xFFFFFFFFu

// catalogue-0 tuples per hit partition when pass A does not read the catalogue (its buckets are resident):
// 16 bytes per lane and load, warp-uniform fast path like the one inside pass A
__global__ void __launch_bounds__(256) hit_hist_kernel(const u32 *__restrict__ halo, u64 n, int log_hp, u32 nbh, u32 *__restrict__ t0)
{
	__shared__ u32 s_hist[HIT_MAX_PARTS];
	for (u32 i = threadIdx.x; i < nbh; i += 256) s_hist[i] = 0;
	__syncthreads();
	const u32 lane = threadIdx.x & 31;
	const u64 stride = (u64) gridDim.x * 256 * 4;
	for (u64 i = ((u64) blockIdx.x * 256 + threadIdx.x) * 4; i < ((n + 127) & ~127ull); i += stride) {      // whole warps stay in the loop
		if (((i >> 12) % HIST_SAMPLE) != 0) continue;                     // every HIST_SAMPLE-th run of 4096 tuples, scaled (like pass A; warp-uniform)
		u32 d[4];
#pragma unroll
		for (int q = 0; q < 4; q++) d[q] = i + q < n ? halo[i + q] >> log_hp : 0xFFFFFFFFu;
		const u32 d0 = __shfl_sync(0xffffffffu, d[0], 0);
		const bool same = d[0] == d0 && d[1] == d0 && d[2] == d0 && d[3] == d0;
		if (__all_sync(0xffffffffu, same)) {
			if (lane == 0 && d0 < nbh) atomicAdd(&s_hist[d0], 128u * HIST_SAMPLE);
		} else {
#pragma unroll
			for (int q = 0; q < 4; q++) if (d[q] < nbh) atomicAdd(&s_hist[d[q]], (u32) HIST_SAMPLE);
		}
	}
	__syncthreads();
	for (u32 i = threadIdx.x; i < nbh; i += 256) { const u32 c = s_hist[i]; if (c) atomicAdd(&t0[i * CSTRIDE], c); }
}

// block-wide exclusive scan of one u64 per thread (1024 threads)
__device__ __forceinline__ u64 block_excl_scan_u64(u64 v, u64 *s_warp /* [32] */, u64 *total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	u64 incl = v;
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		const u64 x = __shfl_up_sync(0xffffffffu, incl, off);
		if (lane >= off) incl += x;
	}
	__syncthreads();
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	u64 pre = 0, tot = 0;
	for (int w = 0; w < (int) (blockDim.x >> 5); w++) { const u64 x = s_warp[w]; if (w < warp) pre += x; tot += x; }
	if (total) *total = tot;
	return pre + incl - v;
}

// region capacity and base per hit partition (1 block of HIT_MAX_PARTS threads)
__global__ void __launch_bounds__(HIT_MAX_PARTS) hit_plan_kernel(const u32 *__restrict__ t0, u32 nbh, u32 slack, u64 hits_cap, u64 *__restrict__ hit_base,
								  u32 *__restrict__ hit_cap, u32 *overflow)
{
	__shared__ u64 s_warp[32];
	const u32 d = threadIdx.x;
	u64 cap = 0;
	if (d < nbh) cap = (2ull * t0[d * CSTRIDE] + slack + 1ull) & ~1ull;
	if (cap > 0xFFFFFFF0ull) cap = 0xFFFFFFF0ull;                    // 32-bit fill counters; more raises the overflow flag
	const u64 base = block_excl_scan_u64(cap, s_warp, nullptr);
	// never past the allocation (the tuple counts are sampled estimates): a region that does not fit is cut and the flag raised
	if (d < nbh && base + cap > hits_cap) { cap = base < hits_cap ? (hits_cap - base) & ~1ull : 0; atomicOr(overflow, 8u); }
	if (d < nbh) { hit_base[d] = base; hit_cap[d] = (u32) cap; }
}

// ---- pass C: bucket join, hits out --------------------------------------------------------------
// Bucket g exists twice, once per catalogue (the partition passes keep the catalogues apart).  Its
// catalogue-1 members go into a bucketised shared-memory table: BJ_NB1 cells of 4 slots, filled through
// one counter per cell, so a probe is a fixed 4-way compare (no open-addressing chains, no divergence).
// A full cell (Poisson tail, ~6 % of the members) spills into BJ_NB2 secondary cells, those into a short
// list.  The catalogue-0 members stream through registers and probe; every match is a hit record staged
// in shared memory, partitioned there by halo-A range and written out in runs.
// All loops over a bucket are rolled: the kernel stays inside the 32 KB instruction cache (a fully
// unrolled version spent most of its time in instruction fetch).
#define BJ_THREADS 512
#define BJ_NB1 1536             // primary cells x 4 slots
#define BJ_NB2 256              // secondary cells x 4 slots
#define BJ_N3 32                // last-resort list
#define BJ_SH 6144              // hits staged per bucket; == 4 * BJ_NB1: the dead table is the sorted buffer
#define BJ_WCAP (BJ_SH / (BJ_THREADS / 32))   // staging slots per warp
#define BJ_KPT 10               // keys per thread and catalogue: cap2 <= BJ_THREADS * BJ_KPT
#define BJ_BK 5                 // keys per thread and load batch
#define BJ_SMEM ((size_t) BJ_NB1 * 32 + (size_t) BJ_SH * 8 + (size_t) HIT_MAX_PARTS * 16)

// cell hash of the pid bits that vary inside a bucket (the top bits of the mixed pid are the bucket number)
__device__ __forceinline__ u32 bj_hash(u64 r) { return ((u32) r ^ (u32) (r >> 32)) * 0x9E3779B1u; }

// Tail of pass C (both table variants): the hits a CTA staged in per-warp slices of `stage` are counted per hit partition,
// the partitions get their space in the global hit regions (one global atomic each), the hits are sorted by partition into
// `sorted` (the dead table) and written out in runs.  `ovf` collects overflow causes (4 hit staging full, 8 hit region full).
__device__ __forceinline__ void bj_flush_hits(const u64 *stage, u64 *sorted, i64 *s_dst, u32 *s_cnt, u32 *s_wcount, u32 *s_warpsum, u32 wbase, u32 wpos, u32 wcap,
					      const HitPlan &hp, const u64 *__restrict__ hit_base, const u32 *__restrict__ hit_cap, u32 *hit_cnt,
					      u64 *__restrict__ hits_out, u64 *counters, u32 &ovf)
{
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	u64 *t1 = sorted;
	if (lane == 0) s_wcount[warp] = wpos;
	if (wpos > wcap) { ovf |= 4u; wpos = wcap; }
	__syncthreads();                                                     // table and counters are dead from here on
	u32 nh = 0;
#pragma unroll
	for (int w = 0; w < BJ_THREADS / 32; w++) nh += s_wcount[w] < wcap ? s_wcount[w] : wcap;
	if (tid == 0 && nh) atomicAdd((unsigned long long *) &counters[0], (unsigned long long) nh);
	for (u32 i = tid; i < hp.nbh; i += BJ_THREADS) s_cnt[i] = 0;
	__syncthreads();
	// hits per partition (every warp walks its own slice)
	const int psh = hp.rec_a_shift + hp.log_hp;
	for (u32 j = lane; j < wpos; j += 32) atomicAdd(&s_cnt[(u32) (stage[wbase + j] >> psh)], 1u);
	__syncthreads();
	// exclusive scan of the partition counts (2 per thread), region space reserved with one global atomic each
	constexpr int BPT = HIT_MAX_PARTS / BJ_THREADS;
	u32 loc[BPT], sum = 0;
#pragma unroll
	for (int k = 0; k < BPT; k++) { const u32 b = tid * BPT + k; loc[k] = b < hp.nbh ? s_cnt[b] : 0u; sum += loc[k]; }
	u32 incl = sum;
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		const u32 v = __shfl_up_sync(0xffffffffu, incl, off);
		if (lane >= off) incl += v;
	}
	if (lane == 31) s_warpsum[warp] = incl;
	__syncthreads();
	u32 pre = 0;
#pragma unroll
	for (int w = 0; w < BJ_THREADS / 32; w++) pre += (w < warp) ? s_warpsum[w] : 0u;
	u32 run = pre + incl - sum;
#pragma unroll
	for (int k = 0; k < BPT; k++) {
		const u32 b = tid * BPT + k;
		if (b < hp.nbh) {
			i64 dst = INT64_MIN;
			if (loc[k]) {
				const u32 g0 = atomicAdd(&hit_cnt[b * CSTRIDE], loc[k]);
				if ((u64) g0 + loc[k] <= (u64) hit_cap[b]) dst = (i64) (hit_base[b] + g0) - (i64) run;
				else ovf |= 8u;
			}
			s_dst[b] = dst;
			s_cnt[b] = run;                                              // write cursor of the partition inside the sorted buffer
			run += loc[k];
		}
	}
	__syncthreads();
	for (u32 j = lane; j < wpos; j += 32) {
		const u64 r = stage[wbase + j];
		t1[atomicAdd(&s_cnt[(u32) (r >> psh)], 1u)] = r;
	}
	__syncthreads();
	for (u32 j = tid; j < nh; j += BJ_THREADS) {
		const u64 r = t1[j];
		const i64 dst = s_dst[(u32) (r >> psh)];
		if (dst != INT64_MIN) st_stream_u64(hits_out + dst + (i64) j, r);
	}
}


__global__ void __launch_bounds__(BJ_THREADS, 2) bucket_join_kernel(const u64 *__restrict__ keys0, const u64 *__restrict__ keys1,
								    const u32 *__restrict__ gcount0, const u32 *__restrict__ gcount1, u64 cap2,
								    KeyLayout L, u64 rmask, HitPlan hp, const u64 *__restrict__ hit_base,
								    const u32 *__restrict__ hit_cap, u32 *hit_cnt, u64 *__restrict__ hits_out,
								    u64 *counters, u32 *overflow, u32 wcap /* staging slots per warp, <= BJ_WCAP */)
{
	extern __shared__ __align__(16) unsigned char bsm[];
	u64 *t1 = reinterpret_cast<u64 *>(bsm);                              // [BJ_NB1 * 4] primary cells; later the hits sorted by partition
	u64 *stage = t1 + BJ_NB1 * 4;                                        // [BJ_SH] hits in discovery order
	unsigned char *un = reinterpret_cast<unsigned char *>(stage + BJ_SH); // 16 KB used twice:
	u64 *t2 = reinterpret_cast<u64 *>(un);                               //   build/probe: [BJ_NB2 * 4] secondary cells
	u64 *t3 = t2 + BJ_NB2 * 4;                                           //                [BJ_N3]
	u32 *cnt1 = reinterpret_cast<u32 *>(t3 + BJ_N3);                     //                [BJ_NB1]
	u32 *cnt2 = cnt1 + BJ_NB1;                                           //                [BJ_NB2]
	i64 *s_dst = reinterpret_cast<i64 *>(un);                            //   hit partition: [HIT_MAX_PARTS] global slot of a run minus its local start
	u32 *s_cnt = reinterpret_cast<u32 *>(s_dst + HIT_MAX_PARTS);        //                  [HIT_MAX_PARTS] counts, then write cursors
	static_assert((size_t) BJ_NB2 * 32 + BJ_N3 * 8 + (BJ_NB1 + BJ_NB2) * 4 <= (size_t) HIT_MAX_PARTS * 16, "union overflow");
	static_assert(BJ_SH == BJ_NB1 * 4 && BJ_KPT % BJ_BK == 0, "layout");
	__shared__ u32 s_n3, s_wcount[BJ_THREADS / 32], s_warpsum[BJ_THREADS / 32];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const u64 g = blockIdx.x;
	const u32 n0 = gcount0[g], n1 = gcount1[g];
	if (n0 == 0 || n1 == 0 || (u64) n0 > cap2 || (u64) n1 > cap2) return;       // nothing to join / overflow flagged by the partition pass already
	const u64 *src0 = keys0 + g * cap2, *src1 = keys1 + g * cap2;
	for (int i = tid; i < BJ_NB1 + BJ_NB2; i += BJ_THREADS) cnt1[i] = 0; // cnt2 follows cnt1
	if (tid == 0) s_n3 = 0;
	u32 ovf = 0;                // overflow causes: 2 table full, 4 hit staging full, 8 hit region full
	// build: the first loads fly while the counters are being cleared
#pragma unroll 1
	for (u32 j0 = 0; j0 < n1; j0 += BJ_BK * BJ_THREADS) {
		u64 key[BJ_BK];
#pragma unroll
		for (int i = 0; i < BJ_BK; i++) {
			const u32 j = j0 + i * BJ_THREADS + tid;
			key[i] = j < n1 ? ld_stream_u64(src1 + j) : 0ull;
		}
		if (j0 == 0) __syncthreads();
#pragma unroll
		for (int i = 0; i < BJ_BK; i++) {
			if (j0 + i * BJ_THREADS + tid < n1) {
				const u32 hh = bj_hash(key_pid(L, key[i]) & rmask);
				const u32 c1 = __umulhi(hh, BJ_NB1);
				const u32 p1 = atomicAdd(&cnt1[c1], 1u);
				if (p1 < 4) t1[c1 * 4 + p1] = key[i];
				else {
					const u32 c2 = (hh >> 8) & (BJ_NB2 - 1);
					const u32 p2 = atomicAdd(&cnt2[c2], 1u);
					if (p2 < 4) t2[c2 * 4 + p2] = key[i];
					else {
						const u32 p3 = atomicAdd(&s_n3, 1u);
						if (p3 < BJ_N3) t3[p3] = key[i];
						else ovf |= 2u;
					}
				}
			}
		}
	}
	__syncthreads();
	const u32 n3 = s_n3 < BJ_N3 ? s_n3 : BJ_N3;
	// probe: every match of a catalogue-0 member is one hit
	const int tb = L.type_bits;
	const u64 pmask = ~((1ull << L.pid_shift) - 1ull);
	const u64 lowmask = (1ull << L.snap_shift) - 1ull;                   // (halo, type) of a table entry
	const u32 tmask = L.type_mask;
	// (rolled, one body: the next two members of the thread are already in flight; four were no faster)
	const u32 wbase = (u32) warp * BJ_WCAP;
	u32 wpos = 0;                                                        // hits staged by this warp so far (same in all its lanes)
	u64 ka = tid < n0 ? ld_stream_u64(src0 + tid) : 0ull;
	u64 kb = tid + BJ_THREADS < n0 ? ld_stream_u64(src0 + tid + BJ_THREADS) : 0ull;
#pragma unroll 1
	for (u32 j0 = 0; j0 < n0; j0 += BJ_THREADS) {
		{
			const u64 k0 = ka;
			ka = kb;
			kb = j0 + 2 * BJ_THREADS + tid < n0 ? ld_stream_u64(src0 + j0 + 2 * BJ_THREADS + tid) : 0ull;
			u64 e0 = 0, e1 = 0;                                      // the first two matching table entries
			u32 m = 0, hh = 0;
#define BJ_CHK(e, ok)                                                                        \
			if ((ok) && (((e) ^ k0) & pmask) == 0) { if (m == 0) e0 = (e); else if (m == 1) e1 = (e); m++; }
			if (j0 + tid < n0) {
				hh = bj_hash(key_pid(L, k0) & rmask);
				const u32 c1 = __umulhi(hh, BJ_NB1);
				const u32 kc = cnt1[c1];
				const ulonglong2 ea = *reinterpret_cast<const ulonglong2 *>(t1 + c1 * 4), eb = *reinterpret_cast<const ulonglong2 *>(t1 + c1 * 4 + 2);
				BJ_CHK(ea.x, kc > 0) BJ_CHK(ea.y, kc > 1) BJ_CHK(eb.x, kc > 2) BJ_CHK(eb.y, kc > 3)
				if (kc > 4) {                                        // the cell was full: its overflow is in the secondary cell
					const u32 c2 = (hh >> 8) & (BJ_NB2 - 1);
					const u32 k2 = cnt2[c2];
					const ulonglong2 fa = *reinterpret_cast<const ulonglong2 *>(t2 + c2 * 4), fb = *reinterpret_cast<const ulonglong2 *>(t2 + c2 * 4 + 2);
					BJ_CHK(fa.x, k2 > 0) BJ_CHK(fa.y, k2 > 1) BJ_CHK(fb.x, k2 > 2) BJ_CHK(fb.y, k2 > 3)
					if (k2 > 4)
						for (u32 x = 0; x < n3; x++) { const u64 e = t3[x]; BJ_CHK(e, true) }
				}
			}
#undef BJ_CHK
			const u64 recA = ((u64) key_halo(L, k0) << hp.rec_a_shift) | ((u64) key_type(L, k0) << tb);
			auto make_rec = [&](u64 e) { const u64 low = e & lowmask; return recA | ((low >> tb) << (2 * tb)) | (low & tmask); };
			// hits go to the warp's own slice of the staging array: positions come from ballots, no shared counter
			const u32 b1 = __ballot_sync(0xffffffffu, m > 0);
			if (m > 0) { const u32 pos = wpos + __popc(b1 & lanemask_lt()); if (pos < wcap) stage[wbase + pos] = make_rec(e0); }
			wpos += __popc(b1);
			const u32 b2 = __ballot_sync(0xffffffffu, m > 1);
			if (b2) {
				if (m > 1) { const u32 pos = wpos + __popc(b2 & lanemask_lt()); if (pos < wcap) stage[wbase + pos] = make_rec(e1); }
				wpos += __popc(b2);
				u32 b3 = __ballot_sync(0xffffffffu, m > 2);
				while (b3) {                                             // three or more memberships: one lane at a time rescans (rare)
					const u32 src_lane = __ffs(b3) - 1;
					b3 &= b3 - 1;
					u32 extra = 0;
					if (lane == src_lane) {
						u32 seen = 0;
						auto rest = [&](const u64 *cell, u32 cnt) {
							for (u32 k = 0; k < cnt; k++) {
								const u64 e = cell[k];
								if (((e ^ k0) & pmask) == 0) {
									if (seen >= 2) { if (wpos + extra < wcap) stage[wbase + wpos + extra] = make_rec(e); extra++; }
									seen++;
								}
							}
						};
						const u32 c1 = __umulhi(hh, BJ_NB1), c2 = (hh >> 8) & (BJ_NB2 - 1);
						const u32 kc = cnt1[c1], k2 = cnt2[c2];
						rest(t1 + c1 * 4, kc < 4 ? kc : 4);
						if (kc > 4) {
							rest(t2 + c2 * 4, k2 < 4 ? k2 : 4);
							if (k2 > 4) rest(t3, n3);
						}
					}
					wpos += __shfl_sync(0xffffffffu, extra, src_lane);
				}
			}
		}
	}
	bj_flush_hits(stage, t1, s_dst, s_cnt, s_wcount, s_warpsum, wbase, wpos, wcap, hp, hit_base, hit_cap, hit_cnt, hits_out, counters, ovf);
	if (ovf) atomicOr(overflow, ovf);
}

// ---- mbarrier + 1-D bulk copy (TMA) helpers -----------------------------------------------------
__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 arrivals)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar)
{
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
	asm volatile("{\n"
		     ".reg .pred p;\n"
		     "MBAR_WAIT:\n"
		     "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		     "@p bra MBAR_DONE;\n"
		     "bra MBAR_WAIT;\n"
		     "MBAR_DONE:\n"
		     "}" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is signalled on `bar`
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src, u32 bytes, u64 *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)), "l"(src), "r"(bytes),
		     "r"(smem_addr(bar)) : "memory");
}

// ---- pass C, direct-address variant ------------------------------------------------------------
// mix(pid) is a bijection on pid_bits-bit integers whose top b1+b2 bits are the bucket number, so inside a bucket a particle is
// identified by the remaining rb = pid_bits - b1 - b2 bits.  Dense IDs (a full box numbers its particles 1..N^3 and about half of
// them are bound) leave rb <= 14: the catalogue-1 members of the bucket then sit in a DIRECT-ADDRESS table tab[2^rb] of 32-bit
// payloads ((halo, type) + 1; 0 = empty): no hash, no tag compare, a probe is one 4-byte shared-memory load.  Second and later
// members of a particle (it is listed under a subhalo and under its host) hang off the slot as a chain of (payload, next) nodes.
//
// Because a probe is that cheap it runs TWICE, and the hits are never staged: every thread keeps its catalogue-0 keys in registers;
// sweep 1 only counts (chain length -> one shared-memory add per KEY on the counter of its hit partition, the returned value is the
// key's rank inside the partition); the partition counts are scanned and the space in the global hit regions reserved; sweep 2 walks
// the chains again and drops every hit record at its final place in the partition-sorted buffer, which is then written out in runs.
// No warp ballots, no per-hit counting or scattering passes, five block barriers per bucket.
//
// The catalogue-1 keys of the bucket arrive through 1-D bulk copies (TMA) into the sorted buffer, which is idle until sweep 2: no
// load instructions, address arithmetic or register staging on the build side, and the copies fly while the table is cleared.
#define BD_NODES 1024           // chain nodes per bucket
#define BD_FLAG 0x80000000u     // slot / next holds a node index, not a payload
#ifndef BD_THREADS
#define BD_THREADS 512
#endif
#define BD_MAXKEYS (BJ_THREADS * BJ_KPT)                        // largest bucket (keys per catalogue) the plan produces
#define BD_KPT ((BD_MAXKEYS + BD_THREADS - 1) / BD_THREADS)     // catalogue-0 keys per thread (all held in registers)
#define BD_CHUNK (4 * BD_THREADS)                               // keys per bulk copy == keys one build sweep of the CTA consumes (4 per thread)
#define BD_MAXCHUNK ((BD_MAXKEYS + BD_CHUNK - 1) / BD_CHUNK)
#define BD_MAX_RB 14            // 64 KB table (one CTA per SM); 13 bits and less: two CTAs per SM
static size_t bd_table_bytes(int rb) { return ((size_t) 4 << rb) + (size_t) BD_NODES * 8; }
static size_t bd_smem_bytes(int rb) { return bd_table_bytes(rb) + (size_t) BJ_SH * 8 + (size_t) HIT_MAX_PARTS * 12; }

__device__ __forceinline__ void bd_insert(u32 *tab, uint2 *nodes, u32 *s_nnodes, u32 r, u32 p, u32 old, u32 &ovf)
{
	if (old != 0u) {                                                     // the particle has a member already: chain this one in front
		const u32 i = atomicAdd(s_nnodes, 1u);
		if (i < BD_NODES) nodes[i] = make_uint2(p, atomicExch(&tab[r], BD_FLAG | i));
		else ovf |= 2u;
	}
}

__global__ void __launch_bounds__(BD_THREADS, 2) bucket_join_direct_kernel(const u64 *__restrict__ keys0, const u64 *__restrict__ keys1,
									   const u32 *__restrict__ gcount0, const u32 *__restrict__ gcount1, u64 cap2,
									   KeyLayout L, int rb, HitPlan hp, const u64 *__restrict__ hit_base,
									   const u32 *__restrict__ hit_cap, u32 *hit_cnt, u64 *__restrict__ hits_out,
									   u64 *counters, u32 *overflow, u32 hcap /* hits a bucket may produce, <= BJ_SH */,
									   u32 pf_dist, u32 pf_bytes0, u32 pf_bytes1)
{
	extern __shared__ __align__(128) unsigned char dsm[];
	u32 *tab = reinterpret_cast<u32 *>(dsm);                             // [2^rb] direct-address slots
	uint2 *nodes = reinterpret_cast<uint2 *>(tab + (1u << rb));          // [BD_NODES] (payload, next)
	u64 *sorted = reinterpret_cast<u64 *>(nodes + BD_NODES);             // [BJ_SH] build: catalogue-1 keys (TMA destination); sweep 2: hits sorted by partition
	u32 *s_dst = reinterpret_cast<u32 *>(sorted + BJ_SH);                // [HIT_MAX_PARTS] global slot of a partition's run minus its local start (mod 2^32; the hit buffer holds < 2^32 records)
	u32 *s_cnt = s_dst + HIT_MAX_PARTS;                                  // [HIT_MAX_PARTS] hits per partition
	u32 *s_start = s_cnt + HIT_MAX_PARTS;                                // [HIT_MAX_PARTS] local start of a partition in the sorted buffer
	__shared__ __align__(8) u64 s_bar[BD_MAXCHUNK];
	__shared__ u32 s_nnodes, s_warpsum[BD_THREADS / 32];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const u64 g = blockIdx.x;
	const u32 n0 = gcount0[g], n1 = gcount1[g];
	if (n0 == 0 || n1 == 0 || (u64) n0 > cap2 || (u64) n1 > cap2) return;       // nothing to join / overflow flagged by the partition pass already
	const u64 *src0 = keys0 + g * cap2, *src1 = keys1 + g * cap2;
	const u32 nchunk = (n1 + BD_CHUNK - 1) / BD_CHUNK;
	if (tid == 0) {
		s_nnodes = 0;
		for (u32 c = 0; c < nchunk; c++) mbar_init(&s_bar[c], 1);
		mbar_init_fence();
		for (u32 c = 0; c < nchunk; c++) {
			const u32 have = n1 - c * BD_CHUNK < BD_CHUNK ? n1 - c * BD_CHUNK : BD_CHUNK;
			const u32 bytes = ((have + 1) & ~1u) * 8;                  // whole 16-byte units (cap2 is even: the pad key is inside the bucket)
			mbar_expect_tx(&s_bar[c], bytes);
			tma_load_1d(sorted + c * BD_CHUNK, src1 + c * BD_CHUNK, bytes, &s_bar[c]);
		}
	}
	// the buckets a later CTA of this SM will join are pulled into L2 now (blocks start in index order)
	if (tid == 32 && pf_dist && g + pf_dist < gridDim.x) {
		asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src0 + (u64) pf_dist * cap2), "r"(pf_bytes0) : "memory");
		asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src1 + (u64) pf_dist * cap2), "r"(pf_bytes1) : "memory");
	}
	// the catalogue-0 members: requested now, used after the build, kept in registers until the end
	u64 k[BD_KPT];
#pragma unroll
	for (int i = 0; i < BD_KPT; i++) {
		const u32 j = i * BD_THREADS + tid;
		k[i] = j < n0 ? ld_stream_u64(src0 + j) : 0ull;
	}
	{
		uint4 *t4 = reinterpret_cast<uint4 *>(tab);
		for (u32 i = tid; i < (1u << rb) / 4; i += BD_THREADS) t4[i] = make_uint4(0u, 0u, 0u, 0u);
		for (u32 i = tid; i < HIT_MAX_PARTS; i += BD_THREADS) s_cnt[i] = 0;
	}
	u32 ovf = 0;                // overflow causes: 2 table (chain nodes) full, 4 more hits than the sorted buffer holds, 8 hit region full
	__syncthreads();
	const u32 rmask = (1u << rb) - 1u;
	const u32 lowmask = (1u << L.snap_shift) - 1u;                       // (halo, type) of a key
	const int psft = L.pid_shift;
	// build: 4 keys per thread and sweep, all four compare-and-swaps in flight before any result is looked at
#pragma unroll 1
	for (u32 c = 0; c < nchunk; c++) {
		mbar_wait(&s_bar[c], 0);
		const u32 j = c * BD_CHUNK + 4 * tid;
		const ulonglong2 ka = *reinterpret_cast<const ulonglong2 *>(sorted + j), kb = *reinterpret_cast<const ulonglong2 *>(sorted + j + 2);
		const u64 kk[4] = {ka.x, ka.y, kb.x, kb.y};
		u32 r[4], p[4], old[4];
#pragma unroll
		for (int q = 0; q < 4; q++) {
			r[q] = (u32) (kk[q] >> psft) & rmask;
			p[q] = ((u32) kk[q] & lowmask) + 1u;
		}
#pragma unroll
		for (int q = 0; q < 4; q++) old[q] = j + q < n1 ? atomicCAS(&tab[r[q]], 0u, p[q]) : 0u;
#pragma unroll
		for (int q = 0; q < 4; q++) bd_insert(tab, nodes, &s_nnodes, r[q], p[q], old[q], ovf);
	}
	__syncthreads();                                                     // table complete; the sorted buffer is free
	const int tb = L.type_bits;
	const u32 tmask = L.type_mask;
	const int dsh = tb + hp.log_hp;                                       // key -> hit partition of its halo
	const u32 dmask = L.halo_mask >> hp.log_hp;
	// sweep 1: hits per key (its chain length) -> counter of the key's partition; the value returned is the key's rank there.
	// From here on a key is its slot value v and its low word kl = (snap, halo, type): the pid has done its job.
	u32 rank[BD_KPT], v[BD_KPT], kl[BD_KPT];
#pragma unroll
	for (int i = 0; i < BD_KPT; i++) {
		const u32 t = tab[(u32) (k[i] >> psft) & rmask];                  // (a key past the end is 0: slot 0 is read and ignored)
		v[i] = i * BD_THREADS + tid < n0 ? t : 0u;
		kl[i] = (u32) k[i];
	}
#pragma unroll
	for (int i = 0; i < BD_KPT; i++) {
		u32 len = v[i] != 0u ? 1u : 0u;
		if (v[i] & BD_FLAG) {
			u32 nx = nodes[v[i] & ~BD_FLAG].y;
			len = 2u;
			while (nx & BD_FLAG) { nx = nodes[nx & ~BD_FLAG].y; len++; }
		}
		rank[i] = 0;
		if (len) rank[i] = atomicAdd(&s_cnt[(kl[i] >> dsh) & dmask], len);
	}
	__syncthreads();
	// exclusive scan of the partition counts (2 per thread); region space reserved with one global atomic per partition
	constexpr int BPT = (HIT_MAX_PARTS + BD_THREADS - 1) / BD_THREADS;
	u32 loc[BPT], sum = 0;
#pragma unroll
	for (int q = 0; q < BPT; q++) { const u32 b = tid * BPT + q; loc[q] = b < hp.nbh ? s_cnt[b] : 0u; sum += loc[q]; }
	u32 incl = sum;
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		const u32 x = __shfl_up_sync(0xffffffffu, incl, off);
		if (lane >= off) incl += x;
	}
	if (lane == 31) s_warpsum[warp] = incl;
	__syncthreads();
	u32 pre = 0, nh = 0;
#pragma unroll
	for (int w = 0; w < BD_THREADS / 32; w++) { const u32 x = s_warpsum[w]; pre += (w < warp) ? x : 0u; nh += x; }
	if (nh > hcap) {                                                     // more hits than the buffer holds (same decision in every thread)
		if (tid == 0) atomicOr(overflow, 4u | ovf);
		return;
	}
	u32 run = pre + incl - sum;
	// the global reservations are issued here and consumed after sweep 2, which hides their latency (an L2 round trip under contention)
	u32 g0[BPT], run0[BPT];
#pragma unroll
	for (int q = 0; q < BPT; q++) {
		const u32 b = tid * BPT + q;
		g0[q] = 0; run0[q] = run;
		if (b < hp.nbh) {
			if (loc[q]) g0[q] = atomicAdd(&hit_cnt[b * CSTRIDE], loc[q]);
			s_start[b] = run;
			run += loc[q];
		}
	}
	if (tid == 0 && nh) atomicAdd((unsigned long long *) &counters[0], (unsigned long long) nh);
	__syncthreads();
	// sweep 2: every member of the slot's chain is one hit, written at its final place in the partition-sorted buffer
#pragma unroll
	for (int i = 0; i < BD_KPT; i++) {
		u32 w = v[i];
		if (w) {
			const u64 recA = ((u64) ((kl[i] >> tb) & L.halo_mask) << hp.rec_a_shift) | (u64) ((kl[i] & tmask) << tb);
			u32 pos = s_start[(kl[i] >> dsh) & dmask] + rank[i];
			do {
				u32 p;
				if (w & BD_FLAG) { const uint2 nd = nodes[w & ~BD_FLAG]; p = nd.x; w = nd.y; }
				else { p = w; w = 0u; }
				const u32 pm = p - 1u;
				sorted[pos++] = recA | (u64) (((pm >> tb) << (2 * tb)) | (pm & tmask));
			} while (w);
		}
	}
#pragma unroll
	for (int q = 0; q < BPT; q++) {
		const u32 b = tid * BPT + q;
		if (b < hp.nbh && loc[q]) {
			// a partition whose region is full keeps a valid destination only in the flag: the whole attempt is discarded by the caller
			if ((u64) g0[q] + loc[q] <= (u64) hit_cap[b]) s_dst[b] = (u32) (hit_base[b] + g0[q]) - run0[q];
			else { s_dst[b] = 0xFFFFFFFFu; ovf |= 8u; }
		}
	}
	if (__syncthreads_or(ovf & 8u)) {                                    // (block-uniform) nothing is written once a region overflowed
		if (ovf) atomicOr(overflow, ovf);
		return;
	}
	const int psh = hp.rec_a_shift + hp.log_hp;
#pragma unroll 4
	for (u32 j = tid; j < nh; j += BD_THREADS) {
		const u64 r = sorted[j];
		st_stream_u64(hits_out + (u32) (s_dst[(u32) (r >> psh)] + j), r);
	}
	if (ovf) atomicOr(overflow, ovf);
}

// ---- pass D: count the hits of one halo-A range in shared memory -------------------------------
// Work units: a partition, or a slice of a heavy one (a catalogue dominated by one halo); slices of
// the same partition emit separate records for the same pair, which the selection stage sums.
#define HR_THREADS 512
#define HR_OVF 3072             // (== the constexpr OVF of the unweighted kernel) entries of the per-block hash table for pairs beyond a halo's WAYS slots (36 KB: table + queues + 64 KB of ways = 108 KB, two blocks per SM)
#define HR_OVF_PROBES 32
#define HR_MAX_UNITS (2 * HIT_MAX_PARTS)   // sum of ceil(c_d / slice_max) <= nbh + total / slice_max <= 2 nbh (slices of at least one mean)
#define HR_MAX_SUB_LOG 4                   // a partition holds up to 2^4 sub-ranges of 2048 halos: 32 M catalogue-0 halos per GPU

__global__ void __launch_bounds__(HIT_MAX_PARTS) hit_units_kernel(const u32 *__restrict__ hit_cnt, u32 nbh, u32 slice_x2 /* slice threshold in half means */,
								   uint4 *__restrict__ units, u32 *n_units)
{
	__shared__ u64 s_warp[32];
	const u32 d = threadIdx.x;
	const u64 c = d < nbh ? hit_cnt[d * CSTRIDE] : 0;
	u64 total;
	block_excl_scan_u64(c, s_warp, &total);
	u64 slice_max = slice_x2 * (total / nbh) / 2 + 1;               // only partitions above the mean are split
	if (slice_max < 65536) slice_max = 65536;
	const u64 nsl = c ? (c + slice_max - 1) / slice_max : 0;
	// longest units first (blocks are scheduled in index order: a heavy unit started last would be the tail):
	// rank of this partition by its per-unit size, ties by index
	__shared__ u32 s_size[HIT_MAX_PARTS], s_nsl[HIT_MAX_PARTS];
	const u32 usz = nsl ? (u32) ((c + nsl - 1) / nsl) : 0u;
	s_size[d] = usz; s_nsl[d] = (u32) nsl;
	__syncthreads();
	u32 before_units = 0;
	for (u32 e = 0; e < nbh; e++) {
		const u32 z = s_size[e];
		if (z > usz || (z == usz && e < d)) before_units += s_nsl[e];
	}
	u64 tot_units;
	block_excl_scan_u64(nsl, s_warp, &tot_units);
	if (d < nbh) for (u64 s = 0; s < nsl; s++) units[before_units + s] = make_uint4(d, (u32) s, (u32) nsl, 0u);
	if (d == 0) *n_units = (u32) tot_units;
}

// THREADS: 512 with 2048 halos per block (two blocks per SM); 1024 with 4096 halos (one block per SM) in sub-range mode,
// which halves the number of times a big partition is read.
// Fast path per hit: one 16-byte load of the halo's way tags, compare, one shared-memory add.  A hit that finds no way holding its
// tag (first sight of a pair: a way has to be claimed; or all ways taken: overflow table) is NOT handled in place - 4 % of the hits
// would drag three quarters of the warp iterations through the slow code with one or two active lanes.  It is pushed on the warp's own
// small queue (positions from a ballot, no atomics); whenever 32 are queued the warp runs the slow path once with all lanes busy.
#define HR_QCAP 64              // queue entries per warp: < 32 before a push, <= 32 pushed
// WEIGHTED (multi-GPU, owner side): the input is (record, partial count) pairs from all pid shards instead of single hits, every
// record adds its count; HR_OVF_W overflow entries leave room for the weights on the queues.
#define HR_OVF_W 2560
template <int WAYS, bool SUB, int THREADS, bool WEIGHTED = false>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) hit_reduce_kernel(const u64 *__restrict__ hits, const u32 *__restrict__ weights, const u64 *__restrict__ hit_base,
								 const u32 *__restrict__ hit_cnt, const uint4 *__restrict__ units,
								 const u32 *__restrict__ n_units, HitPlan hp, u64 *__restrict__ rec_key,
								 u32 *__restrict__ rec_cnt, u64 rec_cap, u64 *counters, u32 *overflow)
{
	const int sub_log = SUB ? hp.log_hp - hp.log_hs : 0;
	const u32 sub = SUB ? blockIdx.x & ((1u << sub_log) - 1u) : 0u;
	if ((blockIdx.x >> sub_log) >= *n_units) return;
	extern __shared__ __align__(16) unsigned char hsm[];
	constexpr u32 OVF = WEIGHTED ? HR_OVF_W : 3072;                      // entries of the overflow table (== HR_OVF when unweighted)
	const u32 HP = 1u << hp.log_hs;                                      // halos counted by this block
	u32 *s_tag = reinterpret_cast<u32 *>(hsm);                           // [HP * WAYS]
	u32 *s_count = s_tag + (size_t) HP * WAYS;                           // [HP * WAYS]
	u64 *s_okey = reinterpret_cast<u64 *>(s_count + (size_t) HP * WAYS); // [OVF] pairs whose halo ran out of ways: open addressing on the record
	u32 *s_ocnt = reinterpret_cast<u32 *>(s_okey + OVF);              // [OVF]
	u64 *s_queue = reinterpret_cast<u64 *>(s_ocnt + OVF);             // [THREADS / 32][HR_QCAP]
	u32 *s_qw = reinterpret_cast<u32 *>(s_queue + (size_t) (THREADS / 32) * HR_QCAP);   // WEIGHTED: [THREADS / 32][HR_QCAP] counts of the queued records
	__shared__ u32 s_warpsum[THREADS / 32];
	__shared__ u64 s_base;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint4 u = units[blockIdx.x >> sub_log];
	const u32 d = u.x;
	const u64 cnt = hit_cnt[d * CSTRIDE];
	u64 per = (cnt + u.z - 1) / u.z;
	per = (per + 1) & ~1ull;
	const u64 lo = (u64) u.y * per, hi = lo + per < cnt ? lo + per : cnt;
	for (u32 i = tid; i < HP * WAYS; i += THREADS) { s_tag[i] = HIT_EMPTY; s_count[i] = 0; }
	for (u32 i = tid; i < OVF; i += THREADS) { s_okey[i] = ~0ull; s_ocnt[i] = 0; }
	__syncthreads();
	const u64 *src = hits + hit_base[d];
	const u32 tag_mask = (u32) ((1ull << hp.rec_a_shift) - 1ull);
	const int ash = hp.rec_a_shift;
	u32 spilled = 0;
	bool ovf = false;           // record buffer full (cause 16)
	u64 *queue = s_queue + (size_t) warp * HR_QCAP;
	u32 *queue_w = s_qw + (size_t) warp * HR_QCAP;
	u32 qn = 0;                 // entries on the warp's queue (same in all its lanes)
	const u32 lt = lanemask_lt();
	// slow path for one record per lane: claim a way, else the overflow table, else straight to the output
	auto slow = [&](u64 rec, u32 wgt) {
		const u32 a = (u32) (rec >> ash) & (HP - 1u);
		const u32 tag = (u32) rec & tag_mask;
		u32 *tg = s_tag + (size_t) a * WAYS;
		int w = -1;
#pragma unroll
		for (int x = 0; x < WAYS; x++) {
			u32 cur = *reinterpret_cast<volatile u32 *>(tg + x);
			if (cur == HIT_EMPTY) {
				cur = atomicCAS(tg + x, HIT_EMPTY, tag);
				if (cur == HIT_EMPTY) cur = tag;
			}
			if (cur == tag) { w = x; break; }
		}
		if (w >= 0) { atomicAdd(s_count + (size_t) a * WAYS + w, wgt); return; }
		u32 h = __umulhi((u32) ((rec * 0x9E3779B97F4A7C15ull) >> 32), (u32) OVF);
		int probe = 0;
		for (; probe < HR_OVF_PROBES; probe++) {
			u64 cur = *reinterpret_cast<volatile u64 *>(s_okey + h);
			if (cur == ~0ull) {
				cur = atomicCAS((unsigned long long *) (s_okey + h), ~0ull, (unsigned long long) rec);
				if (cur == ~0ull) cur = rec;
			}
			if (cur == rec) { atomicAdd(s_ocnt + h, wgt); break; }
			h = h + 1 == OVF ? 0u : h + 1;
		}
		if (probe == HR_OVF_PROBES) {
			// last resort: a (record, 1) pair straight to the output
			const u64 o = atomicAdd((unsigned long long *) &counters[2], 1ull);
			if (o < rec_cap) { rec_key[o] = rec; rec_cnt[o] = wgt; }
			else ovf = true;
		}
		spilled++;
	};
	constexpr int UNR = 8;                      // loads in flight per thread (4: 2.7 ms, 8: 2.4 ms, 16: 2.9 ms)
	const u64 span = hi > lo ? hi - lo : 0;
	const u64 iters = (span + (u64) THREADS * UNR - 1) / ((u64) THREADS * UNR);           // same for all threads: whole warps stay in the loop
	// software pipeline: the loads of batch it+1 are in flight while batch it is counted
	u64 rn[UNR];
	u32 wn[UNR];
	const u32 *wsrc = WEIGHTED ? weights + hit_base[d] : nullptr;
#pragma unroll
	for (int q = 0; q < UNR; q++) {
		const u64 j = lo + tid + (u64) q * THREADS;
		rn[q] = j < hi ? ld_stream_u64(src + j) : ~0ull;
		wn[q] = WEIGHTED && j < hi ? wsrc[j] : 1u;
	}
#pragma unroll 1
	for (u64 it = 0; it < iters; it++) {
		const u64 j1 = lo + (it + 1) * (u64) THREADS * UNR + tid;
		u64 r[UNR];
		u32 wq[UNR];
#pragma unroll
		for (int q = 0; q < UNR; q++) {
			r[q] = rn[q];
			wq[q] = wn[q];
			const u64 j = j1 + (u64) q * THREADS;
			rn[q] = j < hi ? ld_stream_u64(src + j) : ~0ull;
			wn[q] = WEIGHTED && j < hi ? wsrc[j] : 1u;
		}
#pragma unroll
		for (int q = 0; q < UNR; q++) {
			const u32 ap = (u32) (r[q] >> ash);                             // halo; its low log_hp bits index the partition
			bool valid = r[q] != ~0ull;
			if (SUB) valid = valid && ((ap >> hp.log_hs) & ((1u << sub_log) - 1u)) == sub;
			const u32 a = ap & (HP - 1u);
			const u32 tag = (u32) r[q] & tag_mask;
			int w = -1;
			if (WAYS == 4) {
				const uint4 t4 = *reinterpret_cast<const uint4 *>(s_tag + (size_t) a * 4);
				w = t4.x == tag ? 0 : t4.y == tag ? 1 : t4.z == tag ? 2 : t4.w == tag ? 3 : -1;
			} else {
				const uint2 t2 = *reinterpret_cast<const uint2 *>(s_tag + (size_t) a * 2);
				w = t2.x == tag ? 0 : t2.y == tag ? 1 : -1;
			}
			if (valid && w >= 0) atomicAdd(s_count + (size_t) a * WAYS + w, WEIGHTED ? wq[q] : 1u);
			const bool pend = valid && w < 0;
			const u32 bq = __ballot_sync(0xffffffffu, pend);
			if (bq) {
				if (pend) { queue[qn + __popc(bq & lt)] = r[q]; if (WEIGHTED) queue_w[qn + __popc(bq & lt)] = wq[q]; }
				qn += __popc(bq);
				if (qn >= 32) {
					__syncwarp();
					qn -= 32;
					const u64 rec = queue[qn + lane];
					const u32 rw = WEIGHTED ? queue_w[qn + lane] : 1u;
					__syncwarp();
					slow(rec, rw);
				}
			}
		}
	}
	__syncwarp();
	if ((u32) lane < qn) slow(queue[lane], WEIGHTED ? queue_w[lane] : 1u);
	__syncthreads();
	// emit the live ways + the overflow list
	u32 live = 0;
	for (u32 i = tid; i < HP * WAYS; i += THREADS) live += s_tag[i] != HIT_EMPTY;
	for (u32 i = tid; i < OVF; i += THREADS) live += s_okey[i] != ~0ull;
	u32 incl = live;
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		const u32 v = __shfl_up_sync(0xffffffffu, incl, off);
		if (lane >= off) incl += v;
	}
	if (lane == 31) s_warpsum[warp] = incl;
	__syncthreads();
	u32 pre = 0, tot = 0;
#pragma unroll
	for (int x = 0; x < THREADS / 32; x++) { const u32 v = s_warpsum[x]; pre += x < warp ? v : 0u; tot += v; }
	if (tid == 0) s_base = tot ? atomicAdd((unsigned long long *) &counters[2], (unsigned long long) tot) : 0ull;
	__syncthreads();
	u64 o = s_base + pre + incl - live;
	if (s_base + tot > rec_cap) ovf = true;
	else {
		for (u32 i = tid; i < HP * WAYS; i += THREADS) {
			const u32 t = s_tag[i];
			if (t != HIT_EMPTY) {
				rec_key[o] = ((u64) ((d << hp.log_hp) + (sub << hp.log_hs) + i / WAYS) << hp.rec_a_shift) | (u64) t;
				rec_cnt[o] = s_count[i];
				o++;
			}
		}
		for (u32 i = tid; i < OVF; i += THREADS) {
			const u64 k = s_okey[i];
			if (k != ~0ull) { rec_key[o] = k; rec_cnt[o] = s_ocnt[i]; o++; }
		}
	}
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) spilled += __shfl_xor_sync(0xffffffffu, spilled, off);
	if (lane == 0 && spilled) atomicAdd((unsigned long long *) &counters[5], (unsigned long long) spilled);
	if (ovf) atomicOr(overflow, 16u);
}

#define PART_MIN_BITS 5            // fewest bucket bits of one partition level (part_kernel<32>)
static size_t part_smem(int nb) { return (size_t) PT_TILE * 8 + (size_t) nb * 16 + (size_t) HIT_MAX_PARTS * 4; }

// Is the partition path applicable, and with which bucket bits?
// finer: halve the buckets `finer` times (the retry after a bucket's table or hit staging overflowed: catalogues whose
// particles have many memberships produce more hits per key than the default bucket size provides for)
// nmax_known >= 0: the tuples of the larger catalogue (the eager partition of catalogue 0 plans before catalogue 1 exists)
bool partition_join_plan(const metro_ctx *ctx, const KeyLayout &L, i64 N, PartPlan *out, i64 *nmax_cap, int finer, i64 nmax_known)
{
	const char *env_min = getenv("METRO_PART_MIN_N");                // (read per call: the tests change it inside one process)
	const i64 min_n = env_min ? atoll(env_min) : (1ll << 23);
	if (N < min_n || L.pid_bits < 16) return false;                 // small pairs: the LSD path is fine
	if (L.halo_bits + 2 * L.type_bits > 31) return false;           // the pass-D way tag (B, tA, tB) is 31 bits
	if (ctx->cat[0].n_halos > ((i64) HIT_MAX_PARTS << (11 + HR_MAX_SUB_LOG))) return false;
	int tb = 0;
	const i64 nmax = nmax_known >= 0 ? nmax_known : std::max(ctx->cat[0].n_tuples, ctx->cat[1].n_tuples);
	while ((nmax >> tb) > 3900 && tb < 20) tb++;                    // final buckets of 2-4 k keys per catalogue
	// dense IDs: one more bucket bit if that leaves BD_DIRECT_RB residual pid bits, the width at which pass C runs its direct-address
	// table at two CTAs per SM (1e9-tuple shard with 2^31 IDs: 6.5 ms instead of 8.4 with the 64 KB table of 14 residual bits).  Two
	// more bits do not pay: quarter-filled buckets cost more than the hashed table on full ones (zoom chain, 15 residual bits:
	// pass C 0.35 ms direct at +2 bits, 0.29 ms at +1 with the 64 KB table, 0.19 ms hashed at +0)
	if (L.pid_bits - tb == BD_DIRECT_RB + 1 && L.pid_bits - BD_DIRECT_RB <= 20) tb = L.pid_bits - BD_DIRECT_RB;
	const char *env_tb = getenv("METRO_TB");                         // experiments: total bucket bits
	if (env_tb && atoi(env_tb) > 0) tb = atoi(env_tb);
	tb = std::min(tb + finer, 20);
	// at least 32 buckets per level (catalogues of 10^7 tuples: 2^12 buckets of 2-4 k keys, not 2^16 of 150 - a pass-C block costs
	// about 10 ns before it has seen a key)
	tb = std::max(tb, 2 * PART_MIN_BITS);
	PartPlan P;
	P.b1 = getenv("METRO_B1") ? atoi(getenv("METRO_B1")) : (tb > 18 ? tb - 9 : (tb + 1) / 2);
	P.b2 = tb - P.b1;
	if (P.b1 < PART_MIN_BITS) P.b1 = PART_MIN_BITS;
	if (P.b2 < PART_MIN_BITS) P.b2 = PART_MIN_BITS;
	if (P.b1 > 10 || P.b2 > 10 || P.b1 + P.b2 > L.pid_bits) return false;
	P.pid_bits = L.pid_bits;
	P.shift1 = L.pid_shift + L.pid_bits - P.b1;
	P.shift2 = P.shift1 - P.b2;
	P.r1 = 8;
	// regions and buckets exist once per catalogue and are sized for the larger one, plus 1/8 so that the next
	// catalogues of a chain (and the token tuples they accumulate) fit the same plan (resident buckets, see PartResident)
	const i64 ncap = nmax + nmax / 8;
	if (nmax_cap) *nmax_cap = ncap;
	const double m1 = (double) ncap / (double) ((1ll << P.b1) * P.r1), m2 = (double) ncap / (double) (1ll << (P.b1 + P.b2));
	P.cap1 = (u64) (m1 + 8.0 * sqrt(m1) + 4096.0);
	P.cap1 = (P.cap1 + 1) & ~1ull;                                  // keep regions 16-byte aligned
	P.cap2 = (u64) (m2 + 9.0 * sqrt(m2) + 64.0);
	P.cap2 = (P.cap2 + 1) & ~1ull;
	if (P.cap2 > (u64) BJ_THREADS * BJ_KPT) return false;
	P.tiles1 = (u32) div_up<u64>(P.cap1, PT_TILE);
	P.mul1 = 0x9E3779B97F4A7C15ull; P.mul2 = 0xD6E8FEB86659FD93ull;
	P.xs = (L.pid_bits + 1) / 2;
	*out = P;
	return true;
}

// rec_halo_bits: halo field width of the hit records (a function of the halo counts only, so that every rank of a
// multi-GPU run packs its records alike; the key layout may reserve more halo bits)
static HitPlan make_hit_plan(int rec_halo_bits, int type_bits, i64 H0)
{
	HitPlan hp;
	hp.rec_a_shift = rec_halo_bits + 2 * type_bits;
	// 2048 halos x 4 ways x 8 B = 64 KB: two pass-D blocks per SM (METRO_LOG_HP=12: 4096 halos per range from the start, one
	// 1024-thread block per SM - half as many ranges for pass C to partition the hits into)
	static const int env_log = getenv("METRO_LOG_HP") && atoi(getenv("METRO_LOG_HP")) == 12 ? 12 : 11;
	auto parts = [&](int lg) { return (u64) div_up<i64>(std::max<i64>(H0, 1), 1ll << lg); };
	// catalogues with few halos (a zoom: 2e4 halos would be 10 ranges of 2048): narrower ranges, down to 16 halos, until there are at
	// least as many ranges as pass D has block slots - every SM gets work and no pair is counted in two places
	int base_log = env_log;
	while (base_log > 4 && parts(base_log) < 2 * METRO_NUM_SMS_B200) base_log--;
	hp.ways = 4; hp.log_hs = base_log;
	hp.log_hp = base_log;
	while (parts(hp.log_hp) > HIT_MAX_PARTS && hp.log_hp < 11 + HR_MAX_SUB_LOG) hp.log_hp++;
	if (hp.log_hp > hp.log_hs) hp.log_hs = 12;                       // sub-range mode: 4096 halos x 4 ways = 128 KB, one 1024-thread block per SM
	hp.nbh = (u32) parts(hp.log_hp);
	hp.slack = 2048 + 2 * HIST_SAMPLE * PT_TILE;                    // + the sampling error of the tuple counts (one tile per range end)
	return hp;
}

template <bool PACK>
static int launch_part(metro_ctx *ctx, int bits, unsigned grid, const u64 *in, const u32 *ih, const u8 *it, u64 n, u32 snap, const KeyLayout &L,
		       const PartPlan &P, u64 *out, u32 *gin, u32 *gout, u32 *ovf, u32 *t0, const HitPlan &hp)
{
	cudaStream_t st = ctx->stream;
	const size_t smem = part_smem(1 << bits);
#define PART_CASE(NB)                                                                                                               \
	case NB: {                                                                                                                  \
		/* per device, so not cached in a process-wide static: a few microseconds per launch */                         \
		CUDA_TRY(ctx, cudaFuncSetAttribute(part_kernel<NB, PACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));    \
		part_kernel<NB, PACK><<<grid, PT_THREADS, smem, st>>>(in, ih, it, n, snap, ctx->pid_base, L, P, out, gin, gout, ovf, t0, hp.log_hp, hp.nbh);   \
		break;                                                                                                              \
	}
	switch (1 << bits) {
		PART_CASE(32) PART_CASE(64) PART_CASE(128) PART_CASE(256) PART_CASE(512) PART_CASE(1024)
	default: metro_set_error(ctx, "unsupported partition width 2^%d", bits); return METRO_ERR_UNSUPPORTED;
	}
#undef PART_CASE
	CUDA_TRY(ctx, cudaGetLastError());
	ctx->launches++;
	return METRO_OK;
}

template <int WAYS, bool SUB, int THREADS>
static int launch_reduce(metro_ctx *ctx, const HitPlan &hp, const u64 *hits, const u64 *hit_base, const u32 *hit_cnt, const uint4 *units, const u32 *n_units,
			 u64 *rec_key, u32 *rec_cnt, u64 rec_cap, u32 *ovf)
{
	const size_t smem = ((size_t) WAYS << hp.log_hs) * 8 + (size_t) HR_OVF * 12 + (size_t) (THREADS / 32) * HR_QCAP * 8;
	CUDA_TRY(ctx, cudaFuncSetAttribute(hit_reduce_kernel<WAYS, SUB, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	hit_reduce_kernel<WAYS, SUB, THREADS><<<HR_MAX_UNITS << (hp.log_hp - hp.log_hs), THREADS, smem, ctx->stream>>>(hits, nullptr, hit_base, hit_cnt, units, n_units, hp, rec_key, rec_cnt, rec_cap,
										   ctx->counters.as<u64>(), ovf);
	CUDA_TRY(ctx, cudaGetLastError());
	ctx->launches++;
	return METRO_OK;
}

// ---- owner-side pre-reduce of the exchanged records (multi-GPU) -----------------------------------
// With the tuples sharded by particle ID every (A,B) pair reaches its owner once per shard: the owner of a halo range receives
// W x its pairs as partial (record, count) entries in arbitrary order.  Instead of pushing all of them through the selection stage's
// global group-by, they are summed the way pass D counts hits: partitioned by 2048-halo range of the owned halos (one streaming
// scatter), then one block per range adds the partial counts in its shared-memory (tag, count) ways.  Output: the distinct records
// of the owned range with summed counts, in the TRANSFORMED form the kernels below work in: the owned halo relative to the range
// start in the top field, and - for the records routed by haloB - haloA and haloB swapped (RecFmt in select.cu undoes both).
__device__ __forceinline__ u64 pr_transform(u64 rec, int hb, int tb, int swapped, u32 x_lo)
{
	const u64 hm = (1ull << hb) - 1ull, tm = (1ull << (2 * tb)) - 1ull;
	const u64 B = (rec >> (2 * tb)) & hm, A = (rec >> (2 * tb + hb)) & hm;
	const u64 X = (swapped ? B : A) - (u64) x_lo, Y = swapped ? A : B;
	return (X << (hb + 2 * tb)) | (Y << (2 * tb)) | (rec & tm);
}

__global__ void __launch_bounds__(256) pr_count_kernel(const u64 *__restrict__ rec, i64 n, int hb, int tb, int swapped, u32 x_lo, u32 nbh, u32 *part_cnt)
{
	__shared__ u32 sh[HIT_MAX_PARTS];
	for (u32 i = threadIdx.x; i < nbh; i += 256) sh[i] = 0;
	__syncthreads();
	const int sh_x = (swapped ? 0 : hb) + 2 * tb;
	const u64 hm = (1ull << hb) - 1ull;
	for (i64 i = (i64) blockIdx.x * 256 + threadIdx.x; i < n; i += (i64) gridDim.x * 256) {
		const u32 d = ((u32) ((rec[i] >> sh_x) & hm) - x_lo) >> 11;
		if (d < nbh) atomicAdd(&sh[d], 1u);
	}
	__syncthreads();
	for (u32 i = threadIdx.x; i < nbh; i += 256) if (sh[i]) atomicAdd(&part_cnt[i * CSTRIDE], sh[i]);
}

// bases of the partitions (exclusive scan of their counts) and the write cursors
__global__ void __launch_bounds__(HIT_MAX_PARTS) pr_base_kernel(const u32 *__restrict__ part_cnt, u32 nbh, u64 *__restrict__ base, unsigned long long *__restrict__ cursor)
{
	__shared__ u64 s_warp[32];
	const u32 d = threadIdx.x;
	const u64 c = d < nbh ? part_cnt[d * CSTRIDE] : 0;
	const u64 b = block_excl_scan_u64(c, s_warp, nullptr);
	if (d < nbh) { base[d] = b; cursor[d] = b; }
}

// block-aggregated scatter into the partitions (one global atomic per block and partition), records transformed on the way
__global__ void __launch_bounds__(256) pr_scatter_kernel(const u64 *__restrict__ rec, const u32 *__restrict__ cnt, i64 n, int hb, int tb, int swapped, u32 x_lo,
							  u32 nbh, unsigned long long *cursor, u64 *__restrict__ out_key, u32 *__restrict__ out_cnt)
{
	__shared__ unsigned long long s_base[HIT_MAX_PARTS];
	__shared__ u32 s_c[HIT_MAX_PARTS];
	for (u32 i = threadIdx.x; i < nbh; i += 256) s_c[i] = 0;
	__syncthreads();
	const int sh_x = (swapped ? 0 : hb) + 2 * tb;
	const u64 hm = (1ull << hb) - 1ull;
	const i64 per = (n + gridDim.x - 1) / gridDim.x;
	const i64 lo = (i64) blockIdx.x * per, hi = lo + per < n ? lo + per : n;
	for (i64 i = lo + threadIdx.x; i < hi; i += 256) {
		const u32 d = ((u32) ((rec[i] >> sh_x) & hm) - x_lo) >> 11;
		if (d < nbh) atomicAdd(&s_c[d], 1u);
	}
	__syncthreads();
	for (u32 i = threadIdx.x; i < nbh; i += 256) {
		s_base[i] = s_c[i] ? atomicAdd(&cursor[i], (unsigned long long) s_c[i]) : 0ull;
		s_c[i] = 0;
	}
	__syncthreads();
	for (i64 i = lo + threadIdx.x; i < hi; i += 256) {
		const u64 r = rec[i];
		const u32 d = ((u32) ((r >> sh_x) & hm) - x_lo) >> 11;
		if (d >= nbh) continue;
		const unsigned long long p = s_base[d] + atomicAdd(&s_c[d], 1u);
		out_key[p] = pr_transform(r, hb, tb, swapped, x_lo);
		out_cnt[p] = cnt[i];
	}
}

// records (key, count)[R] whose halo X (A, or B if swapped) lies in [x_lo, x_hi) -> distinct records with summed counts in
// okey / ocnt (transformed form, see above); *d_total (device) = their number.  W: 6 scratch buffers.  Returns METRO_ERR_UNSUPPORTED
// (nothing done) when the range does not fit the partition tables.
int stage_prereduce(metro_ctx *ctx, const u64 *key, const u32 *cnt, i64 R, int halo_bits, int type_bits, bool swapped, i64 x_lo, i64 x_hi, DevBuf *W,
		    DevBuf &okey, DevBuf &ocnt, u64 *d_total)
{
	cudaStream_t st = ctx->stream;
	const i64 span = x_hi - x_lo;
	if (span <= 0 || R <= 0 || halo_bits + 2 * type_bits > 31) return METRO_ERR_UNSUPPORTED;
	const u32 nbh = (u32) div_up<i64>(span, 2048);
	if (nbh > HIT_MAX_PARTS) return METRO_ERR_UNSUPPORTED;
	int rc;
	// W[0]: part_cnt u32[1024*CSTRIDE] | ovf u32 | n_units u32;  W[1]: base u64[1024] + cursor u64[1024] + out counters u64[8];  W[2]: units;
	// W[3], W[4]: partitioned key / count
	if ((rc = dev_reserve(ctx, W[0], (size_t) HIT_MAX_PARTS * CSTRIDE * 4 + 64))) return rc;
	if ((rc = dev_reserve(ctx, W[1], (size_t) HIT_MAX_PARTS * 16 + 64))) return rc;
	if ((rc = dev_reserve(ctx, W[2], (size_t) HR_MAX_UNITS * 16))) return rc;
	if ((rc = dev_reserve(ctx, W[3], (size_t) R * 8 + 64))) return rc;
	if ((rc = dev_reserve(ctx, W[4], (size_t) R * 4 + 64))) return rc;
	if ((rc = dev_reserve(ctx, okey, (size_t) R * 8 + 64))) return rc;
	if ((rc = dev_reserve(ctx, ocnt, (size_t) R * 4 + 64))) return rc;
	u32 *part_cnt = W[0].as<u32>(), *ovf = part_cnt + (size_t) HIT_MAX_PARTS * CSTRIDE, *n_units = ovf + 1;
	u64 *base = W[1].as<u64>();
	unsigned long long *cursor = reinterpret_cast<unsigned long long *>(base + HIT_MAX_PARTS);
	u64 *outc = base + 2 * HIT_MAX_PARTS;                      // [8]: the reduce kernel's counters ([2] = records out)
	CUDA_TRY(ctx, cudaMemsetAsync(W[0].p, 0, (size_t) HIT_MAX_PARTS * CSTRIDE * 4 + 64, st));
	CUDA_TRY(ctx, cudaMemsetAsync(outc, 0, 8 * sizeof(u64), st));
	const int blocks = (int) std::min<i64>(div_up<i64>(R, 256 * 8), (i64) METRO_NUM_SMS_B200 * 8);
	pr_count_kernel<<<blocks, 256, 0, st>>>(key, R, halo_bits, type_bits, swapped ? 1 : 0, (u32) x_lo, nbh, part_cnt);
	CUDA_TRY(ctx, cudaGetLastError());
	pr_base_kernel<<<1, HIT_MAX_PARTS, 0, st>>>(part_cnt, nbh, base, cursor);
	CUDA_TRY(ctx, cudaGetLastError());
	pr_scatter_kernel<<<blocks, 256, 0, st>>>(key, cnt, R, halo_bits, type_bits, swapped ? 1 : 0, (u32) x_lo, nbh, cursor, W[3].as<u64>(), W[4].as<u32>());
	CUDA_TRY(ctx, cudaGetLastError());
	hit_units_kernel<<<1, HIT_MAX_PARTS, 0, st>>>(part_cnt, nbh, 6u, reinterpret_cast<uint4 *>(W[2].p), n_units);
	CUDA_TRY(ctx, cudaGetLastError());
	HitPlan hp;
	hp.rec_a_shift = halo_bits + 2 * type_bits; hp.log_hp = 11; hp.log_hs = 11; hp.ways = 4; hp.nbh = nbh; hp.slack = 0;
	const size_t smem = ((size_t) 4 << hp.log_hs) * 8 + (size_t) HR_OVF_W * 12 + (size_t) (HR_THREADS / 32) * HR_QCAP * 12;
	CUDA_TRY(ctx, cudaFuncSetAttribute(hit_reduce_kernel<4, false, HR_THREADS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
	hit_reduce_kernel<4, false, HR_THREADS, true><<<HR_MAX_UNITS, HR_THREADS, smem, st>>>(W[3].as<u64>(), W[4].as<u32>(), base, part_cnt, reinterpret_cast<uint4 *>(W[2].p), n_units, hp,
											    okey.as<u64>(), ocnt.as<u32>(), (u64) R, outc, ovf);
	CUDA_TRY(ctx, cudaGetLastError());
	ctx->launches += 5;
	CUDA_TRY(ctx, cudaMemcpyAsync(d_total, outc + 2, sizeof(u64), cudaMemcpyDeviceToDevice, st));
	return METRO_OK;
}

// Buffers and small state of one partition join; a function of the plan and of catalogue 0 alone, so that the eager partition of
// catalogue 0 (stage_partition_eager) and the join proper lay them out alike.
struct PartState {
	u64 nbk, nb1, nb, hits_cap;
	HitPlan hp;
	u32 *g1, *hit_cnt, *t0, *g2, *hit_cap, *ovf, *n_units;
	u64 *hit_base;
	uint4 *units;
};

// grow_ok = false: return 1 instead of (re)allocating a buffer that is too small (another stream may be using it)
static int part_state_reserve(metro_ctx *ctx, const PartPlan &P, int rec_halo_bits, int type_bits, PartState *S, bool grow_ok = true)
{
	const DevCatalogue &c0 = ctx->cat[0];
	S->nbk = 1ull << (P.b1 + P.b2);                                        // buckets per catalogue
	S->nb1 = 2 * (1ull << P.b1) * P.r1; S->nb = 2 * S->nbk;                 // level-1 regions, final buckets (both catalogues)
	S->hp = make_hit_plan(rec_halo_bits, type_bits, c0.n_halos);
	// 2 hits per catalogue-0 tuple + the per-range slack; the tuple counts are sampled (every HIST_SAMPLE-th tile, scaled) by up to two
	// streams (resident part: hit_hist_kernel; the rest: pass A), each over-estimating by up to HIST_SAMPLE tiles
	S->hits_cap = 2ull * ((u64) c0.n_tuples + 2ull * HIST_SAMPLE * PT_TILE) + (u64) (S->hp.slack + 2) * S->hp.nbh;
	int rc;
	const size_t need_a = (std::max((size_t) S->nb1 * P.cap1, (size_t) S->hits_cap) + PT_TILE) * 8 + 64, need_b = ((size_t) S->nb * P.cap2 + PT_TILE) * 8 + 64;
	// small state: [g1 nb1*CSTRIDE][hit_cnt 1024*CSTRIDE][t0 1024*CSTRIDE][g2 nb][hit_cap 1024][ovf, n_units, pad..16] u32, then hit_base u64[1024],
	// units uint4[HR_MAX_UNITS]
	const size_t n32 = (size_t) S->nb1 * CSTRIDE + 2 * (size_t) HIT_MAX_PARTS * CSTRIDE + S->nb + HIT_MAX_PARTS + 16;
	const size_t off_base = (n32 * 4 + 15) & ~(size_t) 15;
	const size_t off_units = off_base + (size_t) HIT_MAX_PARTS * 8;
	const size_t need_c = off_units + (size_t) HR_MAX_UNITS * 16;
	if (!grow_ok && (!ctx->keys_a.p || ctx->keys_a.cap < need_a || !ctx->keys_b.p || ctx->keys_b.cap < need_b || !ctx->part_counts.p || ctx->part_counts.cap < need_c))
		return 1;
	// keys_a: level-1 regions, dead after pass B and reused for the hit regions
	if ((rc = dev_reserve(ctx, ctx->keys_a, need_a))) return rc;
	if ((rc = dev_reserve(ctx, ctx->keys_b, need_b))) return rc;
	if ((rc = dev_reserve(ctx, ctx->part_counts, need_c))) return rc;
	S->g1 = ctx->part_counts.as<u32>(); S->hit_cnt = S->g1 + S->nb1 * CSTRIDE; S->t0 = S->hit_cnt + (size_t) HIT_MAX_PARTS * CSTRIDE;
	S->g2 = S->t0 + (size_t) HIT_MAX_PARTS * CSTRIDE; S->hit_cap = S->g2 + S->nb;
	S->ovf = S->hit_cap + HIT_MAX_PARTS; S->n_units = S->ovf + 1;
	S->hit_base = reinterpret_cast<u64 *>(ctx->part_counts.as<unsigned char>() + off_base);
	S->units = reinterpret_cast<uint4 *>(ctx->part_counts.as<unsigned char>() + off_units);
	return METRO_OK;
}

// Whole front end of a pair on the partition path: records (scratch[0] = key u64[R], scratch[1] =
// count u32[R]; the same pair may appear more than once).  *fell_back = true: an overflow flag was
// raised (nothing usable was produced; the caller runs the LSD path).  Timing events: ev[1] after
// pass A, ev[2] after pass B, ev[10] after pass C (pass D ends at the caller's ev[3]).
// reuse0: the final buckets of catalogue 0 are resident (PartResident): only its tail (token tuples) and catalogue 1
// go through passes A and B.
int stage_partition_join(metro_ctx *ctx, const KeyLayout &L, const PartPlan &P, int rec_halo_bits, bool reuse0, i64 *n_records, i64 *n_hits, bool *fell_back,
			 u32 *overflow_mask, bool finer_plan)
{
	if (overflow_mask) *overflow_mask = 0;
	cudaStream_t st = ctx->stream;
	*fell_back = false;
	const DevCatalogue &c0 = ctx->cat[0], &c1 = ctx->cat[1];
	PartState S;
	int rc;
	if ((rc = part_state_reserve(ctx, P, rec_halo_bits, L.type_bits, &S))) return rc;
	const u64 nbk = S.nbk, nb1 = S.nb1, hits_cap = S.hits_cap;
	const HitPlan hp = S.hp;
	const u64 rec_cap = 16ull * (u64) (c0.n_halos + c1.n_halos) + (1ull << 20);
	if ((rc = dev_reserve(ctx, ctx->scratch[0], (size_t) rec_cap * sizeof(u64)))) return rc;
	if ((rc = dev_reserve(ctx, ctx->scratch[1], (size_t) rec_cap * sizeof(u32)))) return rc;
	u32 *g1 = S.g1, *hit_cnt = S.hit_cnt, *t0 = S.t0, *g2 = S.g2, *hit_cap = S.hit_cap, *ovf = S.ovf, *n_units = S.n_units;
	u64 *hit_base = S.hit_base;
	uint4 *units = S.units;
	u64 *counters = ctx->counters.as<u64>();
	PartResident &R = ctx->resident;
	const int half0 = reuse0 ? R.half0 : 0, half1 = 1 - half0;          // halves of keys_b / g2 that hold catalogue 0 / 1
	// catalogue 1 went through pass A chunk by chunk while it was uploaded (partition_eager1_*): its level-1 regions and their fill
	// counters are in place
	const bool a1 = reuse0 && R.a1 && R.a1_gen == c1.gen;
	ctx->last_a1 = a1;
	R.cat0 = false; R.cat1 = false; R.a1 = false;                       // re-established at the end
	// everything is cleared except (reuse0) the level-2 fill counters of the resident half and (a1) the level-1 counters of catalogue 1
	if (a1) {
		const size_t hw = (size_t) (nb1 / 2) * CSTRIDE;                 // level-1 counter words of one catalogue
		CUDA_TRY(ctx, cudaMemsetAsync(g1 + (size_t) half0 * hw, 0, hw * 4, st));
		CUDA_TRY(ctx, cudaMemsetAsync(hit_cnt, 0, (size_t) (g2 - hit_cnt) * 4, st));
	} else CUDA_TRY(ctx, cudaMemsetAsync(g1, 0, (size_t) (g2 - g1) * 4, st));
	CUDA_TRY(ctx, cudaMemsetAsync(g2 + (size_t) half1 * nbk, 0, (size_t) nbk * 4, st));
	if (!reuse0) CUDA_TRY(ctx, cudaMemsetAsync(g2 + (size_t) half0 * nbk, 0, (size_t) nbk * 4, st));
	CUDA_TRY(ctx, cudaMemsetAsync(hit_cap, 0, (size_t) (HIT_MAX_PARTS + 16) * 4, st));
	CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, 8 * sizeof(u64), st));
	// pass A: both catalogues (the order inside a bucket is irrelevant); the catalogue number handed to the kernel
	// only selects the half of the region / bucket arrays
	for (int s = 0; s < 2; s++) {
		const DevCatalogue &c = ctx->cat[s];
		const i64 first = (s == 0 && reuse0) ? std::min(R.tail0, c.n_tuples) : 0, cnt = c.n_tuples - first;
		if (s == 0 && reuse0 && c.n_tuples > 0) {
			// the resident part is not read: count its tuples per hit range from the halo column alone
			hit_hist_kernel<<<METRO_NUM_SMS_B200 * 8, 256, 0, st>>>(c.d_halo(), (u64) first, hp.log_hp, hp.nbh, t0);
			CUDA_TRY(ctx, cudaGetLastError());
			ctx->launches++;
		}
		if (cnt <= 0 || (s == 1 && a1)) continue;
		if ((rc = launch_part<true>(ctx, P.b1, (unsigned) div_up<i64>(cnt, PT_TILE), c.d_pid() + first, c.d_halo() + first, c.d_type() ? c.d_type() + first : nullptr,
					    (u64) cnt, (u32) (s == 0 ? half0 : half1), L, P, ctx->keys_a.as<u64>(), nullptr, g1, ovf, s == 0 ? t0 : nullptr, hp))) return rc;
	}
	// hit regions: sampled catalogue-0 tuple count per halo range (pass A / hit_hist_kernel) -> capacity and base
	hit_plan_kernel<<<1, HIT_MAX_PARTS, 0, st>>>(t0, hp.nbh, hp.slack, hits_cap, hit_base, hit_cap, ovf);
	CUDA_TRY(ctx, cudaGetLastError());
	ctx->launches++;
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[1], st));
	// pass B
	if ((rc = launch_part<false>(ctx, P.b2, (unsigned) (nb1 * P.tiles1), ctx->keys_a.as<u64>(), nullptr, nullptr, 0, 0, L, P, ctx->keys_b.as<u64>(), g1, g2,
				     ovf, nullptr, hp))) return rc;
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[2], st));
	// pass C
	{
		// test hook: a smaller staging capacity for the default bucket size forces the finer-bucket retry
		static const u32 test_wcap = getenv("METRO_TEST_WCAP") ? (u32) atoi(getenv("METRO_TEST_WCAP")) : 0u;
		static const bool no_direct = getenv("METRO_NO_DIRECT") && atoi(getenv("METRO_NO_DIRECT"));
		const u32 wcap = (test_wcap && !finer_plan) ? std::min<u32>(test_wcap, BJ_WCAP) : (u32) BJ_WCAP;
		const int rb = L.pid_bits - P.b1 - P.b2;
		const u64 *kb = ctx->keys_b.as<u64>();
		if (!no_direct && rb <= BD_MAX_RB && L.snap_shift <= 30 && hits_cap < (1ull << 32)) {
			// few residual pid bits (dense IDs): direct-address table
			const int rbk = rb < 2 ? 2 : rb;
			const size_t smem = bd_smem_bytes(rbk);
			// L2 prefetch distance in buckets (0 = off) and the bytes to pull per bucket and catalogue (mean fill + 1 sigma, 16-byte units)
			static const u32 pf_dist = getenv("METRO_C_PREFETCH") ? (u32) atoi(getenv("METRO_C_PREFETCH")) : 320u;
			u32 pf_b[2];
			for (int h = 0; h < 2; h++) {
				const double m = (double) ctx->cat[h == half0 ? 0 : 1].n_tuples / (double) nbk;
				pf_b[h] = (u32) std::min<u64>(P.cap2 * 8, (((u64) (m + sqrt(m)) + 2) & ~1ull) * 8);
			}
			// test hook: a smaller hit capacity for the default bucket size forces the finer-bucket retry
			const u32 hcap = (test_wcap && !finer_plan) ? std::min<u32>(test_wcap * 16, BJ_SH) : (u32) BJ_SH;
			CUDA_TRY(ctx, cudaFuncSetAttribute(bucket_join_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
			bucket_join_direct_kernel<<<(unsigned) nbk, BD_THREADS, smem, st>>>(kb + (size_t) half0 * nbk * P.cap2, kb + (size_t) half1 * nbk * P.cap2,
											    g2 + (size_t) half0 * nbk, g2 + (size_t) half1 * nbk, P.cap2, L, rbk, hp, hit_base,
											    hit_cap, hit_cnt, ctx->keys_a.as<u64>(), counters, ovf, hcap, pf_dist, pf_b[half0], pf_b[half1]);
		} else {
			CUDA_TRY(ctx, cudaFuncSetAttribute(bucket_join_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) BJ_SMEM));
			const u64 rmask = rb >= 64 ? ~0ull : ((1ull << rb) - 1ull);
			bucket_join_kernel<<<(unsigned) nbk, BJ_THREADS, BJ_SMEM, st>>>(kb + (size_t) half0 * nbk * P.cap2, kb + (size_t) half1 * nbk * P.cap2, g2 + (size_t) half0 * nbk,
											g2 + (size_t) half1 * nbk, P.cap2, L, rmask, hp, hit_base, hit_cap, hit_cnt,
											ctx->keys_a.as<u64>(), counters, ovf, wcap);
		}
		CUDA_TRY(ctx, cudaGetLastError());
		ctx->launches++;
	}
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[10], st));
	// pass D
	// ranges with more than 1.5 x the mean number of hits are sliced (each slice emits its pairs again; the selection stage sums them):
	// 3 x the mean: pass D 1.58 ms, 1.5 x: 1.39 ms (+0.06 ms of selection), 1 x: 1.41 ms (+0.17 ms)
	static const u32 slice_x2 = getenv("METRO_D_SLICE_X2") ? std::max(2, atoi(getenv("METRO_D_SLICE_X2"))) : 3;
	hit_units_kernel<<<1, HIT_MAX_PARTS, 0, st>>>(hit_cnt, hp.nbh, slice_x2, units, n_units);
	CUDA_TRY(ctx, cudaGetLastError());
	ctx->launches++;
	if (hp.log_hp == hp.log_hs && hp.log_hs == 12) rc = launch_reduce<4, false, 2 * HR_THREADS>(ctx, hp, ctx->keys_a.as<u64>(), hit_base, hit_cnt, units, n_units, ctx->scratch[0].as<u64>(), ctx->scratch[1].as<u32>(), rec_cap, ovf);
	else if (hp.log_hp == hp.log_hs) rc = launch_reduce<4, false, HR_THREADS>(ctx, hp, ctx->keys_a.as<u64>(), hit_base, hit_cnt, units, n_units, ctx->scratch[0].as<u64>(), ctx->scratch[1].as<u32>(), rec_cap, ovf);
	else rc = launch_reduce<4, true, 2 * HR_THREADS>(ctx, hp, ctx->keys_a.as<u64>(), hit_base, hit_cnt, units, n_units, ctx->scratch[0].as<u64>(), ctx->scratch[1].as<u32>(), rec_cap, ovf);
	if (rc) return rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters, counters, 8 * sizeof(u64), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters + 60, ovf, 4, cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	if (*reinterpret_cast<u32 *>(ctx->h_counters + 60) != 0) {
		static const bool dbg = getenv("METRO_DEBUG") != nullptr;
		if (dbg) fprintf(stderr, "[metro] partition join attempt overflowed, mask 0x%x (1 key region, 2 bucket table, 4 hit staging, 8 hit region, "
				 "16 record buffer): retry with finer buckets (2, 4) or LSD path\n", *reinterpret_cast<u32 *>(ctx->h_counters + 60));
		*fell_back = true;
		if (overflow_mask) *overflow_mask = *reinterpret_cast<u32 *>(ctx->h_counters + 60);
		return METRO_OK;
	}
	*n_hits = (i64) ctx->h_counters[0];
	*n_records = (i64) ctx->h_counters[2];
	ctx->res.info.n_table_hits = (i64) ctx->h_counters[5];
	// the buckets of catalogue 1 stay: after metro_shift they are the buckets of the next pair's catalogue 0
	R.cat1 = true; R.gen = c1.gen; R.half0 = half0; R.half1 = half1; R.P = P; R.L = L; R.pid_base = ctx->pid_base;
	return METRO_OK;
}

static bool eager_switched_off()
{
	static const bool off = (getenv("METRO_NO_EAGER") && atoi(getenv("METRO_NO_EAGER"))) || (getenv("METRO_FORCE_LSD") && atoi(getenv("METRO_FORCE_LSD"))) ||
				(getenv("METRO_NO_RESIDENT") && atoi(getenv("METRO_NO_RESIDENT")));
	return off;
}

// Eager partition of catalogue 0 (called at the end of its upload): passes A and B of its tuples run on the context's side stream
// while the caller is still sending catalogue 1 over the bus; the pair then starts with catalogue 0's buckets resident, exactly like
// the second pair of a chain.  Layout and plan are guessed from catalogue 0 alone (same ID range, at most 1/8 more tuples, at most
// twice the halos in catalogue 1); if the pair does not fit them, metro_match_pair plans afresh and this work is lost, nothing else.
int stage_partition_eager(metro_ctx *ctx, const metro_params *prm)
{
	PartResident &R = ctx->resident;
	const DevCatalogue &c0 = ctx->cat[0];
	R.cat0 = false; R.cat1 = false; R.a1 = false;
	if (eager_switched_off() || !c0.valid || c0.n_tuples <= 0 || c0.n_halos <= 0 || !ctx->eager_stream) return METRO_OK;
	u64 pid_base = c0.min_pid;
	const int pid_bits = std::max(1, ceil_log2_u64(c0.max_pid - pid_base));
	{	// the same preference as metro_match_pair: base 0, else a multiple of 2^pid_bits, when it costs no key bit
		const u64 aligned = pid_bits < 64 ? (pid_base >> pid_bits) << pid_bits : 0;
		if (std::max(1, ceil_log2_u64(c0.max_pid)) <= pid_bits) pid_base = 0;
		else if (std::max(1, ceil_log2_u64(c0.max_pid - aligned)) <= pid_bits) pid_base = aligned;
	}
	const int halo_bits = std::max(1, ceil_log2_u64((u64) (c0.n_halos - 1)));
	const int type_bits = prm->nptypes > 1 ? ceil_log2_u64((u64) prm->nptypes - 1) : 0;
	if (2 * halo_bits + 2 * type_bits > 63 || halo_bits + type_bits > 31 || pid_bits + 1 + halo_bits + type_bits > 64) return METRO_OK;
	const int hb = (halo_bits + 1 + 2 * type_bits <= 31 && pid_bits + 2 + halo_bits + type_bits <= 64) ? halo_bits + 1 : halo_bits;
	const KeyLayout L = make_key_layout(pid_bits, hb, type_bits);
	PartPlan P;
	i64 ncap = 0;
	if (!partition_join_plan(ctx, L, 2 * c0.n_tuples, &P, &ncap, ctx->part_finer, c0.n_tuples)) return METRO_OK;
	cudaStream_t keep = ctx->stream;
	ctx->stream = ctx->eager_stream;                                     // launch_part and dev_reserve work on ctx->stream
	const u64 keep_base = ctx->pid_base;
	ctx->pid_base = pid_base;
	int rc = METRO_OK;
	do {
		cudaStream_t st = ctx->stream;
		PartState S;
		if ((rc = part_state_reserve(ctx, P, halo_bits, type_bits, &S))) break;
		if (cudaMemsetAsync(S.g1, 0, (size_t) (S.hit_cap - S.g1) * 4 + (size_t) (HIT_MAX_PARTS + 16) * 4, st) != cudaSuccess ||
		    cudaEventRecord(ctx->ev_eager0, st) != cudaSuccess) { rc = METRO_ERR_CUDA; break; }         // (the upload of catalogue 1 adds to these counters)
		if ((rc = launch_part<true>(ctx, P.b1, (unsigned) div_up<i64>(c0.n_tuples, PT_TILE), c0.d_pid(), c0.d_halo(), c0.d_type(), (u64) c0.n_tuples, 0u, L, P,
					    ctx->keys_a.as<u64>(), nullptr, S.g1, S.ovf, nullptr, S.hp))) break;
		// pass B over the level-1 regions of catalogue 0 only (the first half): those of catalogue 1 may be filling up at this moment
		// (partition_eager1_chunk on the compute stream)
		if ((rc = launch_part<false>(ctx, P.b2, (unsigned) ((S.nb1 / 2) * P.tiles1), ctx->keys_a.as<u64>(), nullptr, nullptr, 0, 0, L, P, ctx->keys_b.as<u64>(), S.g1, S.g2,
					     S.ovf, nullptr, S.hp))) break;
		ctx->h_counters[61] = ~0ull;
		if (cudaMemcpyAsync(ctx->h_counters + 61, S.ovf, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaEventRecord(ctx->ev_eager, st) != cudaSuccess) {
			rc = METRO_ERR_CUDA;
			break;
		}
		ctx->eager_pending = true;
		R.cat0 = true; R.gen = c0.gen; R.half0 = 0; R.half1 = 1; R.P = P; R.L = L; R.pid_base = pid_base; R.nmax_cap = ncap; R.tail0 = c0.n_tuples;
	} while (0);
	ctx->stream = keep;
	ctx->pid_base = keep_base;
	if (rc == METRO_ERR_CUDA) metro_set_error(ctx, "eager partition of catalogue 0: %s", cudaGetErrorString(cudaGetLastError()));
	return rc;
}

// Before anything touches catalogue 0, the key buffers or the resident state: wait for the eager partition and drop its buckets if a
// region overflowed.
int partition_eager_join(metro_ctx *ctx)
{
	if (!ctx->eager_pending) return METRO_OK;
	ctx->eager_pending = false;
	CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev_eager));
	if ((u32) ctx->h_counters[61] != 0) ctx->resident.cat0 = false;
	return METRO_OK;
}


// ---- pass A of catalogue 1 under its own upload -------------------------------------------------------------------------------
// When the buckets of catalogue 0 are resident (chain step) or being built (stage_partition_eager), layout and plan of the coming
// pair are fixed, so every chunk of catalogue 1 can go through pass A as soon as it has been checked - on the compute stream, while
// the next chunk crosses the bus.  metro_match_pair then starts at pass B.  If the finished catalogue does not fit the plan after
// all (ID range, size, an overflowed region), the marker is not set and the pair partitions it again.
int partition_eager1_begin(metro_ctx *ctx, const metro_params *prm, i64 N, i64 H)
{
	PartResident &R = ctx->resident;
	const DevCatalogue &c0 = ctx->cat[0];
	ctx->e1_on = false;
	R.a1 = false;
	if (eager_switched_off() || !R.cat0 || !c0.valid || R.gen != c0.gen || N <= 0 || H <= 0 || N > R.nmax_cap) return METRO_OK;
	const int type_bits = prm->nptypes > 1 ? ceil_log2_u64((u64) prm->nptypes - 1) : 0;
	if (type_bits != R.L.type_bits || std::max(1, ceil_log2_u64((u64) (H - 1))) > R.L.halo_bits) return METRO_OK;
	PartState S;
	// (while the eager partition of catalogue 0 is running nothing may be reallocated; it reserved the same sizes)
	const int rc = part_state_reserve(ctx, R.P, R.L.halo_bits, type_bits, &S, !ctx->eager_pending);
	if (rc > 0) return METRO_OK;
	if (rc) return rc;
	cudaStream_t st = ctx->stream;
	if (ctx->eager_pending) CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_eager0, 0));      // its counters are cleared by then
	const size_t hw = (size_t) (S.nb1 / 2) * CSTRIDE;
	ctx->e1_g1 = S.g1;
	ctx->e1_ovf = S.ovf + 2;                                                           // a flag of its own, next to the pair's
	CUDA_TRY(ctx, cudaMemsetAsync(S.g1 + (size_t) R.half1 * hw, 0, hw * 4, st));
	CUDA_TRY(ctx, cudaMemsetAsync(ctx->e1_ovf, 0, 4, st));
	ctx->e1_on = true;
	return METRO_OK;
}

int partition_eager1_chunk(metro_ctx *ctx, const u64 *pid, const u32 *halo, const u8 *type, i64 n)
{
	if (!ctx->e1_on || n <= 0) return METRO_OK;
	const PartResident &R = ctx->resident;
	const u64 keep_base = ctx->pid_base;
	ctx->pid_base = R.pid_base;                                                       // (launch_part packs with ctx->pid_base)
	HitPlan hp = {};
	const int rc = launch_part<true>(ctx, R.P.b1, (unsigned) div_up<i64>(n, PT_TILE), pid, halo, type, (u64) n, (u32) R.half1, R.L, R.P, ctx->keys_a.as<u64>(), nullptr,
					 ctx->e1_g1, ctx->e1_ovf, nullptr, hp);
	ctx->pid_base = keep_base;
	return rc;
}

// before the upload's final synchronisation: bring the overflow flag along
int partition_eager1_flush(metro_ctx *ctx)
{
	if (!ctx->e1_on) return METRO_OK;
	ctx->h_counters[63] = ~0ull;
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters + 63, ctx->e1_ovf, 4, cudaMemcpyDeviceToHost, ctx->stream));
	return METRO_OK;
}

// after it: does the finished catalogue fit the plan it was partitioned with?
void partition_eager1_end(metro_ctx *ctx, const DevCatalogue &c, bool ok)
{
	if (!ctx->e1_on) return;
	ctx->e1_on = false;
	PartResident &R = ctx->resident;
	R.a1 = ok && (u32) ctx->h_counters[63] == 0 && c.n_tuples > 0 && c.min_pid >= R.pid_base && ceil_log2_u64(c.max_pid - R.pid_base) <= R.L.pid_bits;
	R.a1_gen = c.gen;
}
