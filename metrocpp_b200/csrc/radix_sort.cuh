// radix_sort.cuh -- LSD radix sort of u64 keys (optional u32 values), onesweep style.
#pragma once
#include "common.cuh"

#define SORT_RADIX_BITS 8
#define SORT_RADIX 256
#define SORT_MAX_PASSES 8

struct SortPlan {
	int begin_bit, end_bit, n_pass;
	int shift[SORT_MAX_PASSES];
	u32 mask[SORT_MAX_PASSES];
};

static inline SortPlan make_sort_plan(int begin_bit, int end_bit)
{
	SortPlan p;
	p.begin_bit = begin_bit; p.end_bit = end_bit;
	int bits = end_bit - begin_bit;
	p.n_pass = bits <= 0 ? 0 : (bits + SORT_RADIX_BITS - 1) / SORT_RADIX_BITS;
	for (int i = 0; i < SORT_MAX_PASSES; i++) { p.shift[i] = 0; p.mask[i] = 0; }
	for (int i = 0; i < p.n_pass; i++) {
		int lo = begin_bit + i * SORT_RADIX_BITS;
		int nb = end_bit - lo < SORT_RADIX_BITS ? end_bit - lo : SORT_RADIX_BITS;
		p.shift[i] = lo;
		p.mask[i] = (1u << nb) - 1u;
	}
	return p;
}

// Workspace owned by the context (grow-only).
struct SortWorkspace {
	u64 *hist = nullptr;          // [SORT_MAX_PASSES][256] digit counts, then exclusive offsets
	void *lookback = nullptr;     // [tiles][256] decoupled look-back words (u32 or u64)
	size_t lookback_bytes = 0;
	u32 *tile_counter = nullptr;  // [SORT_MAX_PASSES]
};

struct metro_ctx;
int sort_workspace_reserve(metro_ctx *ctx, SortWorkspace &ws, i64 n);
void sort_workspace_free(SortWorkspace &ws);

// Zero the digit histograms (hist[] must be zero before sort_histogram / the fused pack kernel).
int sort_clear_hist(metro_ctx *ctx, cudaStream_t st, SortWorkspace &ws);
// hist[p][d] += #keys with digit d in pass p (all passes of the plan, one read of the keys).
int sort_histogram(metro_ctx *ctx, cudaStream_t st, SortWorkspace &ws, const u64 *keys, i64 n, const SortPlan &plan);
// The scatter passes.  Requires hist[] filled for `plan`.  Result lands in `a` if n_pass is even,
// else in `b` (returned through *result).  If ev != NULL, records ev[p] before pass p and
// ev[n_pass] after the last one (caller-owned events, for per-pass timing).
int sort_scatter_passes(metro_ctx *ctx, cudaStream_t st, SortWorkspace &ws, u64 *a, u64 *b, u32 *va, u32 *vb,
			i64 n, const SortPlan &plan, u64 **result, u32 **vresult, cudaEvent_t *ev, int *launches);

// block-level helper: accumulate one key into per-pass shared histograms
__device__ __forceinline__ void hist_accumulate(u32 (*sh)[SORT_RADIX], u64 key, const SortPlan &plan)
{
#pragma unroll
	for (int p = 0; p < SORT_MAX_PASSES; p++)
		if (p < plan.n_pass) atomicAdd(&sh[p][(u32) (key >> plan.shift[p]) & plan.mask[p]], 1u);
}
