// select.cu -- from reduced pair counts to merger trees, entirely on the device.
//
// Replaces, for both directions at once:
//   MergerTree::AssignMap     (reference src/MergerTree.cpp:133-169)  threshold, ascending-ID order
//   MergerTree::SortByMerit   (src/MergerTree.cpp:174-244) + utils SortIndexes (src/utils.cpp:429-449)
//   CleanTrees                (src/MergerTree.cpp:510-652)  mutual-best filter, orphan/token rule,
//                                                           second merit sort
// Work is O(M) with M = distinct (haloA, haloB) pairs (a few per halo), i.e. negligible next to
// the tuple sort; every step is a flat data-parallel kernel, a scan or a small radix sort.
#include <chrono>
#include "context.cuh"

// ---- merit: bit-exact restatement of SortByMerit's arithmetic --------------------------------
// fp32 divide of the int->float rounded sizes, then fp64 with NO fma contraction
// (ratioM*1.01 - 1.0 must be a rounded multiply followed by a subtract), two roundings to fp32.
__device__ __forceinline__ u32 merit_bits_dev(i32 nComm, i32 nPH0, i32 nPH1, i32 iM)
{
	const float q = __fdiv_rn(__int2float_rn(nPH0), __int2float_rn(nPH1));
	double ratioM = (double) q;
	if (ratioM < 1.0) ratioM = __ddiv_rn(1.0, ratioM);
	const double den = __dsub_rn(__dmul_rn(ratioM, 1.01), 1.0);
	float merit = __double2float_rn(__ddiv_rn((double) nComm, den));
	const double fac = __dadd_rn(1.0, __dmul_rn(0.00001, (double) iM));
	merit = __double2float_rn(__dmul_rn((double) merit, fac));
	return __float_as_uint(merit);
}

// monotone map fp32 bits -> u32 (ascending); -0 is folded onto +0 like operator< does
__device__ __forceinline__ u32 float_sortable(u32 b)
{
	if (b == 0x80000000u) b = 0;
	return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void merit_test_kernel(i64 n, const i32 *a, const i32 *b, const i32 *c, const i32 *d, u32 *out)
{
	const i64 i = (i64) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = merit_bits_dev(a[i], b[i], c[i], d[i]);
}

int merit_bits_device(metro_ctx *ctx, i64 n, const i32 *a, const i32 *b, const i32 *c, const i32 *d, u32 *out)
{
	merit_test_kernel<<<(unsigned) div_up<i64>(n, 256), 256, 0, ctx->stream>>>(n, a, b, c, d, out);
	CUDA_TRY(ctx, cudaGetLastError());
	return METRO_OK;
}

// ---- small kernels -----------------------------------------------------------------------------
#define TPB 256
#define GRID(n) ((unsigned) div_up<i64>((n) > 0 ? (n) : 1, TPB))
#define TID const i64 i = (i64) blockIdx.x * blockDim.x + threadIdx.x

__global__ void iota_kernel(u32 *v, i64 n) { TID; if (i < n) v[i] = (u32) i; }
__global__ void copy_u64_kernel(const u64 *a, u64 *b, i64 n) { TID; if (i < n) b[i] = a[i]; }
__global__ void invert_perm_kernel(const u32 *rank2idx, u32 *rank, i64 n) { TID; if (i < n) rank[rank2idx[i]] = (u32) i; }

struct HitLayout { int halo_bits, type_bits; };
__device__ __forceinline__ void hit_decode(const HitLayout &h, u64 hk, u32 &A, u32 &B, u32 &tA, u32 &tB)
{
	const u32 tm = (1u << h.type_bits) - 1u;
	tB = (u32) hk & tm;
	tA = (u32) (hk >> h.type_bits) & tm;
	const u64 hm = (1ull << h.halo_bits) - 1ull;
	B = (u32) ((hk >> (2 * h.type_bits)) & hm);
	A = (u32) ((hk >> (2 * h.type_bits + h.halo_bits)) & hm);
}

// Records that went through the owner-side pre-reduce (partjoin.cu: stage_prereduce) carry the owned halo relative to the start of the
// owned range in the top field and, if they were routed by haloB, haloA and haloB swapped.
struct RecFmt { int swapped; u32 off; };
__device__ __forceinline__ void rec_decode(const HitLayout &h, const RecFmt &f, u64 hk, u32 &A, u32 &B, u32 &tA, u32 &tB)
{
	u32 X, Y;
	hit_decode(h, hk, X, Y, tA, tB);
	X += f.off;
	A = f.swapped ? Y : X;
	B = f.swapped ? X : Y;
}

// record -> (group halo, order key inside the group): forward (A, idRank1[B]), backward (B, idRank0[A])
__global__ void rec_split_kernel(const u64 *rec, i64 n, HitLayout hl, RecFmt fmt, const u32 *rank_other, int backward, u32 *main, u32 *sec)
{
	TID;
	if (i >= n) return;
	u32 A, B, tA, tB;
	rec_decode(hl, fmt, rec[i], A, B, tA, tB);
	main[i] = backward ? B : A;
	sec[i] = rank_other[backward ? A : B];
}

__global__ void rank_of_kernel(const u32 *idx, const u32 *rank, i64 n_max, const u64 *nd, u32 *out)
{
	TID;
	if (i < n_max && i < (i64) *nd) out[i] = rank[idx[i]];
}

// counting sort by group: element i goes to some slot of its group's range (the order inside a group is fixed afterwards by the
// segmented sort on the order key)
__global__ void group_scatter_kernel(const u32 *main, const u32 *sec, i64 n_max, const u64 *nd, const u32 *segoff, u32 *cursor, u32 *slot_idx, u64 *slot_key)
{
	TID;
	if (i >= n_max || (nd && i >= (i64) *nd)) return;
	const u32 m = main[i];
	const u32 p = segoff[m] + atomicAdd(&cursor[m], 1u);
	slot_idx[p] = (u32) i;
	slot_key[p] = (u64) sec[i];
}

__global__ void group_perm_kernel(const u32 *sidx, const u32 *slot_idx, const u64 *slot_key, const u32 *main, i64 n_max, const u64 *nd, u32 *perm, u64 *key_sorted)
{
	TID;
	if (i >= n_max || (nd && i >= (i64) *nd)) return;
	const u32 s = sidx[i], e = slot_idx[s];
	perm[i] = e;
	key_sorted[i] = ((u64) main[e] << 32) | (u64) (u32) slot_key[s];
}

__global__ void count_group_kernel(const u32 *main, i64 n_max, const u64 *nd, u32 *cnt)
{
	TID;
	if (i < n_max && (!nd || i < (i64) *nd)) atomicAdd(&cnt[main[i]], 1u);
}

__global__ void head_flag_kernel(const u64 *key, i64 n, u32 *flag)
{
	TID;
	if (i < n) flag[i] = (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
}

// merge the (tA,tB) records of one (A,B) pair: by-type counts for both directions
__global__ void pair_merge_kernel(const u64 *rec, const u32 *rec_cnt, const u32 *order, const u32 *flag,
				  const u32 *scan, i64 n, HitLayout hl, RecFmt fmt, int T, u32 *pA, u32 *pB, i32 *pF, i32 *pK, i32 *ptot)
{
	TID;
	if (i >= n) return;
	const u32 r = order[i];
	const u32 p = scan[i] + flag[i] - 1u;
	u32 A, B, tA, tB;
	rec_decode(hl, fmt, rec[r], A, B, tA, tB);
	if (flag[i]) { pA[p] = A; pB[p] = B; }
	const i32 c = (i32) rec_cnt[r];
	atomicAdd(&pF[(i64) p * T + tB], c);      // forward tree: type on the catalogue-1 side
	atomicAdd(&pK[(i64) p * T + tA], c);      // backward tree: type on the catalogue-0 side
	atomicAdd(&ptot[p], c);
}

// Element counts that only the device knows (distinct pairs, kept pairs, clean entries) are read from device memory:
// kernels are launched over an upper bound (the record count) and stop at *nd - no host round trip in the middle of the stage.
// AssignMap threshold: keep iff sum_t nCommon > minPartCmp (strict)
__global__ void keep_flag_kernel(const i32 *ptot, i64 n_max, const u64 *nd, i32 min_part_cmp, u32 *flag)
{
	TID;
	if (i < n_max) flag[i] = (i < (i64) *nd && ptot[i] > min_part_cmp) ? 1u : 0u;
}

__global__ void pair_compact_kernel(const u32 *flag, const u32 *scan, i64 n, int T, const u32 *pA, const u32 *pB,
				    const i32 *pF, const i32 *pK, const i32 *ptot, u32 *kA, u32 *kB, i32 *kF, i32 *kK, i32 *ktot)
{
	TID;
	if (i >= n || !flag[i]) return;
	const u32 d = scan[i];
	kA[d] = pA[i]; kB[d] = pB[i]; ktot[d] = ptot[i];
	for (int t = 0; t < T; t++) { kF[(i64) d * T + t] = pF[i * T + t]; kK[(i64) d * T + t] = pK[i * T + t]; }
}

__global__ void count_by_kernel(const u32 *main, const u64 *nd, u32 *cnt) { TID; if (i < (i64) *nd) atomicAdd(&cnt[main[i]], 1u); }

__global__ void gather_dir_kernel(const u32 *perm, const u64 *nd, int T, const u32 *src_main, const u32 *src_other, const i32 *src_nc,
				  const i32 *src_tot, u32 *m_main, u32 *m_other, i32 *m_nc, i32 *m_tot)
{
	TID;
	if (i >= (i64) *nd) return;
	const u32 s = perm[i];
	m_main[i] = src_main[s]; m_other[i] = src_other[s]; m_tot[i] = src_tot[s];
	for (int t = 0; t < T; t++) m_nc[i * T + t] = src_nc[(i64) s * T + t];
}

// merit of element i (position iM inside its segment) and the (main, ~merit) sort key
__global__ void merit_key_kernel(const u32 *m_main, const u32 *m_other, const i32 *m_tot, const u32 *segoff, const u32 *segcnt,
				 const i32 *np_main, const i32 *np_other, const u64 *nd, u32 *merit, u64 *key, u32 *val)
{
	TID;
	if (i >= (i64) *nd) return;
	const u32 a = m_main[i];
	u32 mb = 0, low = 0;
	if (segcnt[a] > 1) {
		const i32 iM = (i32) ((u32) i - segoff[a]);
		const i32 o = (i32) m_other[i];
		mb = merit_bits_dev(m_tot[i], np_main[a], o >= 0 ? np_other[o] : np_main[a], iM);
		low = ~float_sortable(mb);              // descending merit
	}
	merit[i] = mb;
	key[i] = ((u64) a << 32) | low;
	val[i] = (u32) i;
}

// Segmented merit sort: the candidates of one halo (a few, rarely hundreds) are ordered by (descending
// merit, position) without touching the other segments.  Same order as the stable radix sort on
// (halo, ~merit) it replaces, at a fraction of the passes.  Three tiers: up to 8 candidates are sorted by
// 8 lanes in registers (bitonic network, four segments per warp), 9-32 by a whole warp, longer ones by
// a block (seg_sort_long_kernel); the two upper tiers work from lists filled by the first kernel.
template <int W>
__device__ __forceinline__ u64 bitonic_lanes(u64 k, u32 lane)
{
#pragma unroll
	for (u32 size = 2; size <= W; size <<= 1) {
#pragma unroll
		for (u32 stride = size >> 1; stride > 0; stride >>= 1) {
			const u64 o = __shfl_xor_sync(0xffffffffu, k, stride, W);
			const bool up = (lane & size) == 0 || size == W;        // ascending block (the last merge is ascending)
			const bool lower = (lane & stride) == 0;
			k = (up == lower) ? (o < k ? o : k) : (o > k ? o : k);
		}
	}
	return k;
}

__global__ void __launch_bounds__(256) seg_sort_small_kernel(const u64 *__restrict__ key, const u32 *__restrict__ segoff, const u32 *__restrict__ segcnt,
							      i64 lo, i64 H, u32 *__restrict__ sidx, u32 *mid_list, u32 *long_list, u32 *n_lists /* [0] mid, [1] long */)
{
	const i64 a = lo + (((i64) blockIdx.x * 256 + threadIdx.x) >> 3);                 // segments [lo, H): the ones this rank owns
	const u32 lane = threadIdx.x & 7;
	u32 L = 0, off = 0;
	if (a < H) { L = segcnt[a]; off = segoff[a]; }
	if (lane == 0) {
		if (L == 1) sidx[off] = off;
		else if (L > 32) long_list[atomicAdd(n_lists + 1, 1u)] = (u32) a;
		else if (L > 8) mid_list[atomicAdd(n_lists, 1u)] = (u32) a;
	}
	const bool mine = L >= 2 && L <= 8;
	u64 k = (mine && lane < L) ? ((key[off + lane] & 0xFFFFFFFFull) << 32) | lane : ~0ull;
	k = bitonic_lanes<8>(k, lane);                                           // every lane of the warp takes part in the shuffles
	if (mine && lane < L) sidx[off + lane] = off + (u32) (k & 0xFFFFFFFFull);
}

__global__ void __launch_bounds__(256) seg_sort_mid_kernel(const u64 *__restrict__ key, const u32 *__restrict__ segoff, const u32 *__restrict__ segcnt,
							    const u32 *__restrict__ mid_list, const u32 *__restrict__ n_lists, u32 *__restrict__ sidx)
{
	const u32 lane = threadIdx.x & 31, n = n_lists[0];
	for (u32 s = (blockIdx.x * 256 + threadIdx.x) >> 5; s < n; s += gridDim.x * 8) {       // warp-uniform
		const u32 a = mid_list[s], L = segcnt[a], off = segoff[a];
		u64 k = lane < L ? ((key[off + lane] & 0xFFFFFFFFull) << 32) | lane : ~0ull;
		k = bitonic_lanes<32>(k, lane);
		if (lane < L) sidx[off + lane] = off + (u32) (k & 0xFFFFFFFFull);
	}
}

// long segments, one block per segment: up to 4096 entries by a bitonic network on (key, position) in shared memory; beyond that
// (a halo with more than 4096 candidate records) rank by counting (keys are distinct: the position breaks ties), O(L^2)
#define SEG_LONG_SMEM 4096
__global__ void __launch_bounds__(256) seg_sort_long_kernel(const u64 *__restrict__ key, const u32 *__restrict__ segoff, const u32 *__restrict__ segcnt,
							     const u32 *__restrict__ long_list, const u32 *__restrict__ n_lists, u32 *__restrict__ sidx)
{
	__shared__ u64 s_k[SEG_LONG_SMEM];
	u32 *s_low = reinterpret_cast<u32 *>(s_k);                          // counting path: 2048 staged keys
	for (u32 s = blockIdx.x; s < n_lists[1]; s += gridDim.x) {
		const u32 a = long_list[s], L = segcnt[a], off = segoff[a];
		if (L <= SEG_LONG_SMEM) {
			u32 n = 64;
			while (n < L) n <<= 1;
			for (u32 i = threadIdx.x; i < n; i += 256) s_k[i] = i < L ? ((key[off + i] & 0xFFFFFFFFull) << 32) | i : ~0ull;
			__syncthreads();
			for (u32 size = 2; size <= n; size <<= 1) {
				for (u32 stride = size >> 1; stride > 0; stride >>= 1) {
					for (u32 t = threadIdx.x; t < n / 2; t += 256) {
						const u32 lo = 2 * t - (t & (stride - 1)), hi = lo + stride;     // partner pair of thread t
						const bool up = (lo & size) == 0;
						const u64 x = s_k[lo], y = s_k[hi];
						if ((x > y) == up) { s_k[lo] = y; s_k[hi] = x; }
					}
					__syncthreads();
				}
			}
			for (u32 i = threadIdx.x; i < L; i += 256) sidx[off + i] = off + (u32) (s_k[i] & 0xFFFFFFFFull);
			__syncthreads();
			continue;
		}
		for (u32 i0 = 0; i0 < L; i0 += 256) {
			const u32 i = i0 + threadIdx.x;
			const u32 mine = i < L ? (u32) key[off + i] : 0u;
			u32 rank = 0;
			for (u32 c0 = 0; c0 < L; c0 += 2048) {                      // candidates staged through shared memory
				const u32 cn = L - c0 < 2048 ? L - c0 : 2048;
				__syncthreads();
				for (u32 j = threadIdx.x; j < cn; j += 256) s_low[j] = (u32) key[off + c0 + j];
				__syncthreads();
				if (i < L)
					for (u32 j = 0; j < cn; j++) {
						const u32 v = s_low[j];
						rank += (v < mine) || (v == mine && c0 + j < i);
					}
			}
			if (i < L) sidx[off + rank] = off + i;
		}
		__syncthreads();
	}
}

// SortIndexes tie quirk: inside a run of equal merits every position receives the run's first
// element (lowest pre-sort index, which the stable sort put first).
__global__ void tie_quirk_kernel(const u32 *sidx, const u32 *m_main, const u32 *merit, const u32 *segoff, const u64 *nd, u32 *src)
{
	TID;
	if (i >= (i64) *nd) return;
	const u32 me = sidx[i];
	const u32 a = m_main[me];
	const u32 v = float_sortable(merit[me]);
	i64 q = i;
	const i64 lo = segoff[a];
	while (q > lo && float_sortable(merit[sidx[q - 1]]) == v) q--;
	src[i] = sidx[q];
}

__global__ void emit_raw_kernel(const u32 *src, const u64 *nd, int T, const u32 *m_main, const u32 *m_other, const i32 *m_nc, const u32 *merit,
				const u32 *segcnt, u32 *raw_main, i32 *raw_prog, i32 *raw_nc, u32 *raw_merit)
{
	TID;
	if (i >= (i64) *nd) return;
	const u32 s = src[i];
	const u32 a = m_main[s];
	raw_main[i] = a;
	raw_prog[i] = (i32) m_other[s];
	raw_merit[i] = segcnt[a] > 1 ? merit[s] : 0u;
	for (int t = 0; t < T; t++) raw_nc[i * T + t] = m_nc[(i64) s * T + t];
}

__global__ void off32_to_64_kernel(const u32 *off, i64 n, const u64 *total, i64 *out)
{
	TID;
	if (i < n) out[i] = off[i];
	if (i == n) out[n] = (i64) *total;
}

// best descendant of every catalogue-1 halo: head of its backward tree (-1 if none)
__global__ void best_desc_kernel(const u32 *off1, const u32 *cnt1, const i32 *raw_prog1, i64 H1, i32 *best)
{
	TID;
	if (i < H1) best[i] = cnt1[i] ? raw_prog1[off1[i]] : -1;
}

// mutual-best filter (CleanTrees :539-568): keep B iff ID(bestDesc[B]) == ID(A)
__global__ void clean_flag_kernel(const u32 *raw_main0, const i32 *raw_prog0, i64 n_max, const u64 *nd, const i32 *best, const u64 *id0, u32 *flag, u32 *cntc)
{
	TID;
	if (i >= n_max) return;
	if (i >= (i64) *nd) { flag[i] = 0; return; }
	const u32 a = raw_main0[i];
	const i32 d = best[raw_prog0[i]];
	const u32 f = (d >= 0 && id0[d] == id0[a]) ? 1u : 0u;
	flag[i] = f;
	if (f) atomicAdd(&cntc[a], 1u);
}

// orphan / token rule (CleanTrees :575-617)
__global__ void orphan_kernel(const u32 *cntc, const i32 *nptot0, const i32 *norph0, i64 H0, i64 a_lo, i64 a_hi, i32 min_part_halo, i32 fac,
			      i32 maxs, u32 *token, u32 *nfin, u32 *has_tree, u64 *counters)
{
	TID;
	if (i >= H0) return;
	u32 tk = 0;
	if (i >= a_lo && i < a_hi && cntc[i] == 0 && nptot0[i] > min_part_halo) {
		const i32 steps = norph0[i] + 1;
		i32 loc = 1 + nptot0[i] / fac;
		if (loc > maxs) loc = maxs;
		if (steps <= loc) tk = 1;
		else atomicAdd((unsigned long long *) &counters[4], 1ull);
	}
	token[i] = tk;
	nfin[i] = cntc[i] + tk;
	has_tree[i] = (cntc[i] + tk) > 0 ? 1u : 0u;
}

__global__ void clean_fill_kernel(const u32 *flag, const u32 *scan, const u64 *nd, int T, const u32 *raw_main0, const i32 *raw_prog0,
				  const i32 *raw_nc0, const u32 *cstart, const u32 *coff, u32 *c_main, u32 *c_other, i32 *c_nc, i32 *c_tot)
{
	TID;
	if (i >= (i64) *nd || !flag[i]) return;
	const u32 a = raw_main0[i];
	const u32 d = coff[a] + (scan[i] - cstart[a]);
	c_main[d] = a;
	c_other[d] = (u32) raw_prog0[i];
	i32 tot = 0;
	for (int t = 0; t < T; t++) { i32 v = raw_nc0[i * T + t]; c_nc[(i64) d * T + t] = v; tot += v; }
	c_tot[d] = tot;
}

__global__ void token_fill_kernel(const u32 *token, const u32 *coff, i64 H0, int T, const i32 *npart0, u32 *c_main, u32 *c_other,
				  i32 *c_nc, i32 *c_tot)
{
	TID;
	if (i >= H0 || !token[i]) return;
	const u32 d = coff[i];
	c_main[d] = (u32) i;
	c_other[d] = 0xFFFFFFFFu;                   // -1: the token copy of the main halo itself
	i32 tot = 0;
	for (int t = 0; t < T; t++) { i32 v = npart0[i * T + t]; c_nc[(i64) d * T + t] = v; tot += v; }
	c_tot[d] = tot;
}

__global__ void tree_emit_kernel(const u32 *has_tree, const u32 *tscan, const u32 *token, const u32 *coff, i64 H0, i32 *tree_main,
				 u8 *tree_orphan, i64 *tree_off)
{
	TID;
	if (i >= H0 || !has_tree[i]) return;
	const u32 t = tscan[i];
	tree_main[t] = (i32) i;
	tree_orphan[t] = (u8) token[i];
	tree_off[t] = coff[i];
}

__global__ void token_emit_kernel(const u32 *token, const u32 *kscan, i64 H0, i32 *token_halo)
{
	TID;
	if (i < H0 && token[i]) token_halo[kscan[i]] = (i32) i;
}

// tree_off[n_trees] = clean entries (both counts live on the device)
__global__ void tree_off_end_kernel(i64 *tree_off, const u64 *n_trees, const u64 *n_clean) { tree_off[*n_trees] = (i64) *n_clean; }

// ---- host orchestration ------------------------------------------------------------------------
// The stage runs WITHOUT a host round trip until its very end: counts that only the device knows (distinct pairs P, kept pairs K,
// clean entries C, trees, tokens) stay in the context's device counters; buffers are sized by upper bounds the host does know
// (records in, halos) and kernels launched over those bounds stop at the device-side count.  One copy of the counters + one stream
// synchronisation at the end give the caller its sizes.
#define TRY(x) do { int _rc = (x); if (_rc) return _rc; } while (0)
#define RES(b, bytes) TRY(dev_reserve(ctx, (b), (size_t) (bytes) + 256))
#define LAUNCHED() do { CUDA_TRY(ctx, cudaGetLastError()); ctx->launches++; } while (0)

// device counter slots (u64) of this stage inside metro_ctx::counters
enum { C_UNTRACKED = 4, C_CLEAN = 5, C_TREES = 6, C_TOKENS = 7, C_IDFLAG = 20, C_LISTS = 21, C_P_FWD = 30, C_K_FWD = 31, C_P_BWD = 32, C_K_BWD = 33, C_R_FWD = 34, C_R_BWD = 35 };

// key/value radix sort on bits [0, end_bit) using the four ping-pong buffers
static int sort_kv(metro_ctx *ctx, u64 *ka, u64 *kb, u32 *va, u32 *vb, i64 n, int end_bit, u64 **rk, u32 **rv)
{
	cudaStream_t st = ctx->stream;
	SortPlan plan = make_sort_plan(0, end_bit);
	TRY(sort_workspace_reserve(ctx, ctx->sortws, n));
	TRY(sort_clear_hist(ctx, st, ctx->sortws));
	TRY(sort_histogram(ctx, st, ctx->sortws, ka, n, plan));
	ctx->launches++;
	return sort_scatter_passes(ctx, st, ctx->sortws, ka, kb, va, vb, n, plan, rk, rv, nullptr, &ctx->launches);
}

__global__ void ids_descend_kernel(const u64 *id, i64 n, u64 *flag)
{
	TID;
	if (i + 1 < n && id[i] > id[i + 1]) *flag = 1;
}

// argsort of the halo IDs of one catalogue: rank[h] = position of h in ascending-ID order.  Cached with the catalogue
// (DevCatalogue::id_rank, valid while id_rank_gen == gen and the halo count is unchanged): a chain pays for it once per snapshot.
static int id_ranks(metro_ctx *ctx, DevCatalogue &c, DevBuf *tmp)
{
	cudaStream_t st = ctx->stream;
	const i64 H = c.n_halos;
	if (c.id_rank.p && c.id_rank_gen == c.gen && c.id_rank_n == H) return METRO_OK;
	RES(c.id_rank, H * 4);
	c.id_rank_gen = c.gen; c.id_rank_n = H;
	if (H == 0) return METRO_OK;
	{	// catalogues normally list their halos by ascending ID already: the rank is the index (saves 8 radix passes)
		u64 *flag = ctx->counters.as<u64>() + C_IDFLAG;
		CUDA_TRY(ctx, cudaMemsetAsync(flag, 0, sizeof(u64), st));
		ids_descend_kernel<<<GRID(H), TPB, 0, st>>>(c.halo_id.as<u64>(), H, flag); LAUNCHED();
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters + 8, flag, sizeof(u64), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(ctx, cudaStreamSynchronize(st));
		if (!ctx->h_counters[8]) { iota_kernel<<<GRID(H), TPB, 0, st>>>(c.id_rank.as<u32>(), H); LAUNCHED(); return METRO_OK; }
	}
	RES(tmp[0], H * 8); RES(tmp[1], H * 8); RES(tmp[2], H * 4); RES(tmp[3], H * 4);
	copy_u64_kernel<<<GRID(H), TPB, 0, st>>>(c.halo_id.as<u64>(), tmp[0].as<u64>(), H); LAUNCHED();
	iota_kernel<<<GRID(H), TPB, 0, st>>>(tmp[2].as<u32>(), H); LAUNCHED();
	u64 *rk; u32 *rv;
	TRY(sort_kv(ctx, tmp[0].as<u64>(), tmp[1].as<u64>(), tmp[2].as<u32>(), tmp[3].as<u32>(), H, 64, &rk, &rv));
	invert_perm_kernel<<<GRID(H), TPB, 0, st>>>(rv, c.id_rank.as<u32>(), H); LAUNCHED();
	return METRO_OK;
}

// Segmented sort: positions of every segment [segoff[a], +segcnt[a]) of the halos [lo, hi) ordered by (low 32 key bits, position).
static int seg_sort(metro_ctx *ctx, const u64 *key, const u32 *segoff, const u32 *segcnt, i64 H, i64 lo, i64 hi, DevBuf &mid_b, DevBuf &long_b, u32 *sidx)
{
	cudaStream_t st = ctx->stream;
	RES(mid_b, (H + 1) * 4); RES(long_b, (H + 1) * 4);
	u32 *n_lists = reinterpret_cast<u32 *>(ctx->counters.as<u64>() + C_LISTS);
	CUDA_TRY(ctx, cudaMemsetAsync(n_lists, 0, sizeof(u64), st));
	// only the halos this rank owns have candidates (multi-GPU: 1/world of them)
	seg_sort_small_kernel<<<(unsigned) div_up<i64>(std::max<i64>(hi - lo, 1) * 8, 256), 256, 0, st>>>(key, segoff, segcnt, lo, hi, sidx, mid_b.as<u32>(), long_b.as<u32>(), n_lists);
	LAUNCHED();
	seg_sort_mid_kernel<<<METRO_NUM_SMS_B200 * 4, 256, 0, st>>>(key, segoff, segcnt, mid_b.as<u32>(), n_lists, sidx);
	LAUNCHED();
	seg_sort_long_kernel<<<METRO_NUM_SMS_B200 * 2, 256, 0, st>>>(key, segoff, segcnt, long_b.as<u32>(), n_lists, sidx);
	LAUNCHED();
	return METRO_OK;
}

// Order n elements by (main ascending, sec ascending) without a radix sort: counting sort into the groups of `main` (one atomic per
// element), then the segmented sort of every group on `sec`.  Groups are a handful of entries (the partners of one halo), so this
// is one pass over the data + one pass of register sorts instead of 5-6 radix passes.  perm[j] = element at sorted position j;
// key_sorted[j] = (main << 32) | sec of that element; segcnt / segoff = the groups.  W: 5 scratch buffers.
static int group_by(metro_ctx *ctx, const u32 *main, const u32 *sec, i64 n_max, const u64 *nd, i64 H, i64 lo, i64 hi, DevBuf *W, DevBuf &segcnt_b,
		    DevBuf &segoff_b, u32 *perm, u64 *key_sorted)
{
	cudaStream_t st = ctx->stream;
	RES(segcnt_b, (H + 1) * 4); RES(segoff_b, (H + 1) * 4);
	RES(W[0], (H + 1) * 4);                                    // cursor
	RES(W[1], n_max * 4); RES(W[2], n_max * 8); RES(W[3], n_max * 4);   // slot_idx, slot_key, sidx
	CUDA_TRY(ctx, cudaMemsetAsync(segcnt_b.p, 0, (size_t) (H + 1) * 4, st));
	CUDA_TRY(ctx, cudaMemsetAsync(W[0].p, 0, (size_t) (H + 1) * 4, st));
	u32 *segcnt = segcnt_b.as<u32>(), *segoff = segoff_b.as<u32>();
	count_group_kernel<<<GRID(n_max), TPB, 0, st>>>(main, n_max, nd, segcnt); LAUNCHED();
	TRY(scan_exclusive_u32(ctx, st, segcnt, segoff, H + 1, nullptr));
	group_scatter_kernel<<<GRID(n_max), TPB, 0, st>>>(main, sec, n_max, nd, segoff, W[0].as<u32>(), W[1].as<u32>(), W[2].as<u64>()); LAUNCHED();
	TRY(seg_sort(ctx, W[2].as<u64>(), segoff, segcnt, H, lo, hi, W[4], W[5], W[3].as<u32>()));
	group_perm_kernel<<<GRID(n_max), TPB, 0, st>>>(W[3].as<u32>(), W[1].as<u32>(), W[2].as<u64>(), main, n_max, nd, perm, key_sorted); LAUNCHED();
	return METRO_OK;
}

// Per-direction list (grouped by `main`, ascending ID of `other` inside a group) -> raw tree:
// offsets per main halo, merit sort, tie quirk.  S = 8 scratch buffers.  n_max: upper bound of the list length, *nd: the length.
struct DirList { const u32 *main, *other; const i32 *nc, *tot; i64 n_max; const u64 *nd; };
static int build_tree(metro_ctx *ctx, const DirList &L, i64 H_main, i64 own_lo, i64 own_hi, int T, const i32 *np_main, const i32 *np_other,
		      DevBuf *S, DevBuf &segcnt_b, DevBuf &segoff_b, DevBuf &out_main, DevBuf &out_prog, DevBuf &out_nc, DevBuf &out_merit)
{
	cudaStream_t st = ctx->stream;
	const i64 n = L.n_max;
	RES(segcnt_b, (H_main + 1) * 4); RES(segoff_b, (H_main + 1) * 4);
	RES(out_main, n * 4); RES(out_prog, n * 4); RES(out_nc, n * T * 4); RES(out_merit, n * 4);
	CUDA_TRY(ctx, cudaMemsetAsync(segcnt_b.p, 0, (size_t) (H_main + 1) * 4, st));
	u32 *segcnt = segcnt_b.as<u32>(), *segoff = segoff_b.as<u32>();
	if (n > 0) { count_by_kernel<<<GRID(n), TPB, 0, st>>>(L.main, L.nd, segcnt); LAUNCHED(); }
	TRY(scan_exclusive_u32(ctx, st, segcnt, segoff, H_main + 1, nullptr));
	if (n == 0) return METRO_OK;
	RES(S[0], n * 4); RES(S[1], n * 8); RES(S[2], n * 8); RES(S[3], n * 4); RES(S[4], n * 4); RES(S[5], n * 4);
	u32 *merit = S[0].as<u32>();
	merit_key_kernel<<<GRID(n), TPB, 0, st>>>(L.main, L.other, L.tot, segoff, segcnt, np_main, np_other, L.nd, merit, S[1].as<u64>(), S[3].as<u32>());
	LAUNCHED();
	// order inside every segment: descending merit, ties by position (== the stable sort on (halo, ~merit))
	u32 *sidx = S[4].as<u32>();
	TRY(seg_sort(ctx, S[1].as<u64>(), segoff, segcnt, H_main, own_lo, own_hi, S[6], S[7], sidx));
	tie_quirk_kernel<<<GRID(n), TPB, 0, st>>>(sidx, L.main, merit, segoff, L.nd, S[5].as<u32>()); LAUNCHED();
	emit_raw_kernel<<<GRID(n), TPB, 0, st>>>(S[5].as<u32>(), L.nd, T, L.main, L.other, L.nc, merit, segcnt, out_main.as<u32>(),
						  out_prog.as<i32>(), out_nc.as<i32>(), out_merit.as<u32>());
	LAUNCHED();
	return METRO_OK;
}

// Records (hit key, count) -> distinct (A,B) pairs with per-type counts for both directions, thresholded (AssignMap) and compacted
// in (A, ID(B)) order (backward: (B, ID(A)) order).  Outputs in O[0..4], sized for R entries:
// kA u32, kB u32, kF i32[.*T] (counts by typeB), kK i32[.*T] (by typeA), ktot i32.  *cP / *cK (device): distinct pairs / kept pairs.
static int merge_records(metro_ctx *ctx, const metro_params *prm, HitLayout hl, RecFmt fmt, const u64 *rec_key, const u32 *rec_cnt, i64 R,
			 const u32 *rank_other, bool backward, i64 H_main, i64 own_lo, i64 own_hi, DevBuf *O, u64 *cP, u64 *cK)
{
	cudaStream_t st = ctx->stream;
	const int T = prm->nptypes;
	DevBuf *S = ctx->scratch;
	RES(O[0], R * 4); RES(O[1], R * 4); RES(O[2], R * T * 4); RES(O[3], R * T * 4); RES(O[4], R * 4);
	if (R == 0) {
		CUDA_TRY(ctx, cudaMemsetAsync(cP, 0, sizeof(u64), st));
		CUDA_TRY(ctx, cudaMemsetAsync(cK, 0, sizeof(u64), st));
		return METRO_OK;
	}
	RES(S[4], R * 8); RES(S[5], R * 4); RES(S[6], R * 4); RES(S[7], R * 4); RES(S[8], R * 4); RES(S[9], R * 4);
	rec_split_kernel<<<GRID(R), TPB, 0, st>>>(rec_key, R, hl, fmt, rank_other, backward ? 1 : 0, S[5].as<u32>(), S[6].as<u32>()); LAUNCHED();
	u64 *rk = S[4].as<u64>();
	u32 *order = S[7].as<u32>();
	TRY(group_by(ctx, S[5].as<u32>(), S[6].as<u32>(), R, nullptr, H_main, own_lo, own_hi, S + 70, S[76], S[77], order, rk));
	head_flag_kernel<<<GRID(R), TPB, 0, st>>>(rk, R, S[8].as<u32>()); LAUNCHED();
	TRY(scan_exclusive_u32(ctx, st, S[8].as<u32>(), S[9].as<u32>(), R, cP));
	// distinct pairs P <= R: everything below is sized by R and guarded by *cP
	RES(S[10], R * 4); RES(S[11], R * 4); RES(S[12], R * T * 4); RES(S[13], R * T * 4); RES(S[14], R * 4);
	CUDA_TRY(ctx, cudaMemsetAsync(S[12].p, 0, (size_t) R * T * 4, st));
	CUDA_TRY(ctx, cudaMemsetAsync(S[13].p, 0, (size_t) R * T * 4, st));
	CUDA_TRY(ctx, cudaMemsetAsync(S[14].p, 0, (size_t) R * 4, st));
	pair_merge_kernel<<<GRID(R), TPB, 0, st>>>(rec_key, rec_cnt, order, S[8].as<u32>(), S[9].as<u32>(), R, hl, fmt, T,
						    S[10].as<u32>(), S[11].as<u32>(), S[12].as<i32>(), S[13].as<i32>(), S[14].as<i32>());
	LAUNCHED();
	// AssignMap threshold, ordered compaction
	RES(S[15], R * 4); RES(S[16], R * 4);
	keep_flag_kernel<<<GRID(R), TPB, 0, st>>>(S[14].as<i32>(), R, cP, prm->min_part_cmp, S[15].as<u32>()); LAUNCHED();
	TRY(scan_exclusive_u32(ctx, st, S[15].as<u32>(), S[16].as<u32>(), R, cK));
	pair_compact_kernel<<<GRID(R), TPB, 0, st>>>(S[15].as<u32>(), S[16].as<u32>(), R, T, S[10].as<u32>(), S[11].as<u32>(), S[12].as<i32>(),
						      S[13].as<i32>(), S[14].as<i32>(), O[0].as<u32>(), O[1].as<u32>(), O[2].as<i32>(),
						      O[3].as<i32>(), O[4].as<i32>());
	LAUNCHED();
	return METRO_OK;
}

// recA: records whose haloA this rank owns (forward trees); recB: records whose haloB it owns
// (backward trees).  On one GPU both are the same array.  [a_lo, a_hi) = owned catalogue-0 halos.
int stage_select(metro_ctx *ctx, const metro_params *prm, const KeyLayout &KL, const u64 *recA_key, const u32 *recA_cnt, i64 RA,
		 const u64 *recB_key, const u32 *recB_cnt, i64 RB, i64 a_lo, i64 a_hi)
{
	cudaStream_t st = ctx->stream;
	const int T = prm->nptypes;
	DevCatalogue &c0 = ctx->cat[0], &c1 = ctx->cat[1];
	const i64 H0 = c0.n_halos, H1 = c1.n_halos;
	DevBuf *S = ctx->scratch;
	DevResult &res = ctx->res;
	u64 *cnt = ctx->counters.as<u64>();
	HitLayout hl = {KL.halo_bits, KL.type_bits};

	// METRO_PROF_SELECT=1: phase times of this stage on stderr (CUDA events; rank 0 only)
	static const bool prof = getenv("METRO_PROF_SELECT") != nullptr;
	static cudaEvent_t pev[16];
	static bool pev_ok = false;
	int n_mark = 0;
	const char *mark_name[16];
	double mark_host[16];
	auto MARK = [&](const char *name) {
		if (!prof || ctx->rank != 0 || n_mark >= 16) return;
		if (!pev_ok) { for (auto &e : pev) cudaEventCreate(&e); pev_ok = true; }
		mark_name[n_mark] = name;
		mark_host[n_mark] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
		cudaEventRecord(pev[n_mark++], st);
	};
	MARK("start");
	// halo-ID ranks (std::map iteration order of indexCommon is ascending halo ID, not file order)
	TRY(id_ranks(ctx, c0, S + 40));
	TRY(id_ranks(ctx, c1, S + 40));
	const u32 *rank0 = c0.id_rank.as<u32>(), *rank1 = c1.id_rank.as<u32>();

	i64 b_lo = 0, b_hi = H1;                                  // catalogue-1 halos whose backward trees this rank builds
	comm_owned_range(ctx, H1, &b_lo, &b_hi);
	const bool sharded = recB_key != recA_key;
	// Sharded runs: the partial counts of all pid shards are summed per 2048-halo range of the owned halos first (stage_prereduce),
	// so that the merges below see every pair once, not once per shard.
	RecFmt fmtA = {0, 0u}, fmtB = {0, 0u};
	static const bool no_prereduce = getenv("METRO_NO_PREREDUCE") != nullptr;
	if (sharded && !no_prereduce && ctx->world >= 4) {                  // (two shards: 2.79 vs 2.67 ms - the extra passes cost what the shorter lists save)
		const int rcA = stage_prereduce(ctx, recA_key, recA_cnt, RA, KL.halo_bits, KL.type_bits, false, a_lo, a_hi, S + 70, ctx->xredA_key, ctx->xredA_cnt, cnt + C_R_FWD);
		const int rcB = stage_prereduce(ctx, recB_key, recB_cnt, RB, KL.halo_bits, KL.type_bits, true, b_lo, b_hi, S + 70, ctx->xredB_key, ctx->xredB_cnt, cnt + C_R_BWD);
		if ((rcA != METRO_OK && rcA != METRO_ERR_UNSUPPORTED) || (rcB != METRO_OK && rcB != METRO_ERR_UNSUPPORTED)) return rcA != METRO_OK && rcA != METRO_ERR_UNSUPPORTED ? rcA : rcB;
		if (rcA == METRO_OK || rcB == METRO_OK) {
			CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters + C_R_FWD, cnt + C_R_FWD, 2 * sizeof(u64), cudaMemcpyDeviceToHost, st));
			CUDA_TRY(ctx, cudaStreamSynchronize(st));
		}
		if (rcA == METRO_OK) { recA_key = ctx->xredA_key.as<u64>(); recA_cnt = ctx->xredA_cnt.as<u32>(); RA = (i64) ctx->h_counters[C_R_FWD]; fmtA = {0, (u32) a_lo}; }
		if (rcB == METRO_OK) { recB_key = ctx->xredB_key.as<u64>(); recB_cnt = ctx->xredB_cnt.as<u32>(); RB = (i64) ctx->h_counters[C_R_BWD]; fmtB = {1, (u32) b_lo}; }
		MARK("pre-reduce of the exchanged records");
	}
	// 1.-2. distinct pairs above the threshold: forward set -> S[17..21] in (A, ID(B)) order
	TRY(merge_records(ctx, prm, hl, fmtA, recA_key, recA_cnt, RA, rank1, false, H0, a_lo, a_hi, S + 17, cnt + C_P_FWD, cnt + C_K_FWD));
	MARK("merge forward records");
	const i64 KB_max = sharded ? RB : RA;                    // upper bound of the backward list
	const u64 *cKB = sharded ? cnt + C_K_BWD : cnt + C_K_FWD;

	// Sharded runs: every pair arrives once per pid shard, so the record lists are `world` times longer than the pair lists they
	// merge into.  One small read-back of the kept-pair counts lets everything downstream be sized by the pairs, not the records
	// (on one GPU the two are about equal and the stage stays free of host round trips).
	i64 nA = RA, nB = KB_max;
	if (sharded) TRY(merge_records(ctx, prm, hl, fmtB, recB_key, recB_cnt, RB, rank0, true, H1, b_lo, b_hi, S + 64, cnt + C_P_BWD, cnt + C_K_BWD));
	if (ctx->world > 1) {
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters + C_P_FWD, cnt + C_P_FWD, 4 * sizeof(u64), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(ctx, cudaStreamSynchronize(st));
		nA = (i64) ctx->h_counters[C_K_FWD];
		nB = (i64) ctx->h_counters[sharded ? C_K_BWD : C_K_FWD];
	}
	MARK("merge backward records + sizes");
	// 3. forward raw tree (locMTrees[0])
	DirList fwd = {S[17].as<u32>(), S[18].as<u32>(), S[19].as<i32>(), S[21].as<i32>(), nA, cnt + C_K_FWD};
	TRY(build_tree(ctx, fwd, H0, a_lo, a_hi, T, c0.npart_tot.as<i32>(), c1.npart_tot.as<i32>(), S + 40, S[22], S[23], S[24], res.raw_prog[0],
		       res.raw_ncommon[0], res.raw_merit[0]));
	MARK("forward tree");
	// 4. backward raw tree (locMTrees[1]): the pairs ordered by (B, ID(A)), by-type counts of the 0 side
	DirList bwd;
	if (sharded) {
		// the records routed to owner(B) were merged in backward order right away: no second sort
		bwd = {S[65].as<u32>(), S[64].as<u32>(), S[67].as<i32>(), S[68].as<i32>(), nB, cKB};
	} else {
		RES(S[25], nB * 4); RES(S[26], nB * 4); RES(S[27], nB * T * 4); RES(S[28], nB * 4);
		if (nB > 0) {
			RES(S[40], nB * 4); RES(S[41], nB * 8); RES(S[42], nB * 4);
			rank_of_kernel<<<GRID(nB), TPB, 0, st>>>(S[17].as<u32>(), rank0, nB, cKB, S[40].as<u32>()); LAUNCHED();
			u32 *perm = S[42].as<u32>();
			TRY(group_by(ctx, S[18].as<u32>(), S[40].as<u32>(), nB, cKB, H1, b_lo, b_hi, S + 70, S[76], S[77], perm, S[41].as<u64>()));
			gather_dir_kernel<<<GRID(nB), TPB, 0, st>>>(perm, cKB, T, S[18].as<u32>(), S[17].as<u32>(), S[20].as<i32>(), S[21].as<i32>(),
								     S[25].as<u32>(), S[26].as<u32>(), S[27].as<i32>(), S[28].as<i32>());
			LAUNCHED();
		}
		bwd = {S[25].as<u32>(), S[26].as<u32>(), S[27].as<i32>(), S[28].as<i32>(), nB, cKB};
	}
	TRY(build_tree(ctx, bwd, H1, b_lo, b_hi, T, c1.npart_tot.as<i32>(), c0.npart_tot.as<i32>(), S + 40, S[29], S[30], S[31], res.raw_prog[1],
		       res.raw_ncommon[1], res.raw_merit[1]));
	for (int d = 0; d < 2; d++) {
		const i64 H = d ? H1 : H0;
		RES(res.raw_off[d], (H + 1) * 8);
		off32_to_64_kernel<<<GRID(H + 1), TPB, 0, st>>>((d ? S[30] : S[23]).as<u32>(), H, d ? cKB : cnt + C_K_FWD, res.raw_off[d].as<i64>());
		LAUNCHED();
	}

	MARK("backward tree");
	// 5. CleanTrees: mutual-best filter
	RES(S[32], (H1 + 1) * 4);
	best_desc_kernel<<<GRID(H1), TPB, 0, st>>>(S[30].as<u32>(), S[29].as<u32>(), res.raw_prog[1].as<i32>(), H1, S[32].as<i32>()); LAUNCHED();
	// multi-GPU: the owner of each catalogue-1 halo knows its best descendant, everybody needs it
	TRY(comm_allreduce_max_i32(ctx, S[32].as<i32>(), H1));
	MARK("best descendant + all-reduce");
	RES(S[33], nA * 4); RES(S[34], nA * 4); RES(S[35], (H0 + 1) * 4); RES(S[36], (H0 + 1) * 4);
	u32 *cntc = S[35].as<u32>(), *cstart = S[36].as<u32>();
	CUDA_TRY(ctx, cudaMemsetAsync(cntc, 0, (size_t) (H0 + 1) * 4, st));
	CUDA_TRY(ctx, cudaMemsetAsync(cnt + C_UNTRACKED, 0, sizeof(u64), st));
	if (nA > 0) {
		clean_flag_kernel<<<GRID(nA), TPB, 0, st>>>(S[24].as<u32>(), res.raw_prog[0].as<i32>(), nA, cnt + C_K_FWD, S[32].as<i32>(), c0.halo_id.as<u64>(),
							     S[33].as<u32>(), cntc);
		LAUNCHED();
		TRY(scan_exclusive_u32(ctx, st, S[33].as<u32>(), S[34].as<u32>(), nA, nullptr));
	}
	TRY(scan_exclusive_u32(ctx, st, cntc, cstart, H0 + 1, nullptr));
	// orphan / token rule
	RES(S[37], (H0 + 1) * 4); RES(S[38], (H0 + 1) * 4); RES(S[39], (H0 + 1) * 4);
	u32 *token = S[37].as<u32>(), *nfin = S[38].as<u32>(), *has_tree = S[39].as<u32>();
	CUDA_TRY(ctx, cudaMemsetAsync(token, 0, (size_t) (H0 + 1) * 4, st));
	CUDA_TRY(ctx, cudaMemsetAsync(nfin, 0, (size_t) (H0 + 1) * 4, st));
	CUDA_TRY(ctx, cudaMemsetAsync(has_tree, 0, (size_t) (H0 + 1) * 4, st));
	if (H0 > 0) {
		orphan_kernel<<<GRID(H0), TPB, 0, st>>>(cntc, c0.npart_tot.as<i32>(), c0.n_orphan_steps.as<i32>(), H0, a_lo, a_hi, prm->min_part_halo,
							 prm->fac_orphan_steps, prm->max_orphan_steps, token, nfin, has_tree, cnt);
		LAUNCHED();
	}
	DevBuf *Q = S + 48;              // second scratch bank
	RES(Q[0], (H0 + 1) * 4); RES(Q[1], (H0 + 1) * 4); RES(Q[2], (H0 + 1) * 4);
	u32 *coff = Q[0].as<u32>(), *tscan = Q[1].as<u32>(), *kscan = Q[2].as<u32>();
	TRY(scan_exclusive_u32(ctx, st, nfin, coff, H0 + 1, cnt + C_CLEAN));
	TRY(scan_exclusive_u32(ctx, st, has_tree, tscan, H0 + 1, cnt + C_TREES));
	// the token list is global (every rank appends all token halos in the shift): OR the per-rank flags
	RES(Q[11], (H0 + 1) * 4);
	u32 *gtoken = Q[11].as<u32>();
	CUDA_TRY(ctx, cudaMemcpyAsync(gtoken, token, (size_t) (H0 + 1) * 4, cudaMemcpyDeviceToDevice, st));
	TRY(comm_allreduce_max_i32(ctx, reinterpret_cast<i32 *>(gtoken), H0));
	TRY(scan_exclusive_u32(ctx, st, gtoken, kscan, H0 + 1, cnt + C_TOKENS));

	MARK("mutual best, orphan rule, token all-reduce");
	// clean list grouped by A (kept entries in first-sort order, or the single token entry): C <= K + tokens of the owned halos
	const i64 C_max = nA + (a_hi - a_lo);
	RES(Q[3], C_max * 4); RES(Q[4], C_max * 4); RES(Q[5], C_max * T * 4); RES(Q[6], C_max * 4);
	if (nA > 0) {
		clean_fill_kernel<<<GRID(nA), TPB, 0, st>>>(S[33].as<u32>(), S[34].as<u32>(), cnt + C_K_FWD, T, S[24].as<u32>(), res.raw_prog[0].as<i32>(),
							     res.raw_ncommon[0].as<i32>(), cstart, coff, Q[3].as<u32>(), Q[4].as<u32>(), Q[5].as<i32>(),
							     Q[6].as<i32>());
		LAUNCHED();
	}
	if (H0 > 0) {
		token_fill_kernel<<<GRID(H0), TPB, 0, st>>>(token, coff, H0, T, c0.npart.as<i32>(), Q[3].as<u32>(), Q[4].as<u32>(), Q[5].as<i32>(),
							     Q[6].as<i32>());
		LAUNCHED();
	}
	// second SortByMerit (:627-628) with iM = position in the kept list
	DirList cl = {Q[3].as<u32>(), Q[4].as<u32>(), Q[5].as<i32>(), Q[6].as<i32>(), C_max, cnt + C_CLEAN};
	TRY(build_tree(ctx, cl, H0, a_lo, a_hi, T, c0.npart_tot.as<i32>(), c1.npart_tot.as<i32>(), S + 40, Q[7], Q[8], Q[9], res.clean_prog,
		       res.clean_ncommon, Q[10]));
	MARK("clean list + second merit sort");
	// trees and tokens: at most one per catalogue-0 halo
	RES(res.tree_main, H0 * 4); RES(res.tree_orphan, H0); RES(res.tree_off, (H0 + 1) * 8); RES(res.token_halo, H0 * 4);
	if (H0 > 0) {
		tree_emit_kernel<<<GRID(H0), TPB, 0, st>>>(has_tree, tscan, token, coff, H0, res.tree_main.as<i32>(), res.tree_orphan.as<u8>(),
							    res.tree_off.as<i64>());
		LAUNCHED();
		token_emit_kernel<<<GRID(H0), TPB, 0, st>>>(gtoken, kscan, H0, res.token_halo.as<i32>()); LAUNCHED();
	}
	tree_off_end_kernel<<<1, 1, 0, st>>>(res.tree_off.as<i64>(), cnt + C_TREES, cnt + C_CLEAN); LAUNCHED();
	MARK("emit");
	// the one host round trip of the stage: sizes for the caller
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters, cnt, 40 * sizeof(u64), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	if (n_mark > 1) {
		fprintf(stderr, "[metro] select (RA %lld RB %lld nA %lld nB %lld H0 %lld):", (long long) RA, (long long) RB, (long long) nA, (long long) nB, (long long) H0);
		for (int i = 1; i < n_mark; i++) { float ms = 0; cudaEventElapsedTime(&ms, pev[i - 1], pev[i]); fprintf(stderr, " %s %.3f ms (host %.3f);", mark_name[i], ms, mark_host[i] - mark_host[i - 1]); }
		fprintf(stderr, "\n");
	}
	const u64 *h = ctx->h_counters;
	res.info.n_pairs = (i64) h[C_P_FWD];
	res.info.raw_count[0] = (i64) h[C_K_FWD];
	res.info.raw_count[1] = (i64) (sharded ? h[C_K_BWD] : h[C_K_FWD]);
	res.info.n_untracked = (i64) h[C_UNTRACKED];
	res.info.clean_count = (i64) h[C_CLEAN]; res.info.n_trees = (i64) h[C_TREES]; res.info.n_tokens = (i64) h[C_TOKENS];
	return METRO_OK;
}
