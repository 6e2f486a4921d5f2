// context.cuh -- the engine context behind the C ABI (include/metro_b200.h).
#pragma once
#include <algorithm>
#include <stdarg.h>
#include <string.h>

#include <vector>

#include "../../include/metro_b200.h"
#include "common.cuh"
#include "radix_sort.cuh"

// grow-only device buffer
struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
	template <typename T> T *as() const { return reinterpret_cast<T *>(p); }
};

// One catalogue resident in HBM (SoA).  Replaces locHalos[i] + locMapParts[i] + locParts[i]
// (reference src/global_vars.h:49-82).
struct DevCatalogue {
	i64 n_halos = 0, n_tuples = 0;
	DevBuf halo_id;        // u64[H]
	DevBuf npart;          // i32[H*T]
	DevBuf npart_tot;      // i32[H]   sum over types (Halo::nAllPart)
	DevBuf n_orphan_steps; // i32[H]
	DevBuf is_token;       // u8[H]
	DevBuf pid;            // u64[N]
	DevBuf halo;           // u32[N]
	DevBuf type;           // u8[N]
	DevBuf id_rank;        // u32[H] rank of every halo in ascending-ID order (select.cu, cached while id_rank_gen == gen)
	u64 id_rank_gen = 0;
	i64 id_rank_n = -1;
	// borrowed tuple arrays (metro_catalogue_adopt_device); override pid/halo/type when set
	const u64 *b_pid = nullptr;
	const u32 *b_halo = nullptr;
	const u8 *b_type = nullptr;
	u64 max_pid = 0, min_pid = 0;
	u64 gen = 0;            // changes whenever the tuples of this slot change (metro_ctx::next_gen)
	bool valid = false;
	const u64 *d_pid() const { return b_pid ? b_pid : pid.as<u64>(); }
	const u32 *d_halo() const { return b_pid ? b_halo : halo.as<u32>(); }
	const u8 *d_type() const { return b_pid ? b_type : type.as<u8>(); }
};

// Result of the last match, resident in HBM (layout == metro_result_arrays).
struct DevResult {
	bool valid = false;
	metro_result_info info;
	DevBuf raw_off[2], raw_prog[2], raw_ncommon[2], raw_merit[2];
	DevBuf tree_main, tree_orphan, tree_off, clean_prog, clean_ncommon, token_halo;
};

struct PartPlan {
	int b1, b2;                 // bucket bits of the two levels
	int r1;                     // level-1 regions are split in r1 replicas (block index mod r1): r1 x fewer atomics per counter
	int shift1, shift2;         // key >> shift & (2^b - 1)
	u64 cap1, cap2;             // capacity of a level-1 (catalogue, bucket, replica) region / a final bucket (keys)
	u32 tiles1;                 // tiles per level-1 region (= ceil(cap1 / PT_TILE))
	int pid_bits;
	u64 mul1, mul2;
	int xs;
};

// Final key buckets that stay in HBM between the pairs of a chain: after metro_match_pair the
// buckets of catalogue 1 are kept; metro_shift makes them the buckets of the new catalogue 0 (its token
// tuples, appended behind the old catalogue-1 tuples, are still missing: tail0); the next pair then
// partitions only the new catalogue 1 and that tail, provided its keys fit the same layout and plan.
struct PartResident {
	bool cat1 = false, cat0 = false;
	u64 gen = 0;                // DevCatalogue::gen of the catalogue the buckets belong to
	int half0 = 0, half1 = 1;   // which half of keys_b (and of the level-2 fill counters) holds catalogue 0 / 1
	PartPlan P;
	KeyLayout L;
	u64 pid_base = 0;
	i64 nmax_cap = 0;           // tuples per catalogue the plan was sized for
	i64 tail0 = 0;
	bool a1 = false;            // the level-1 regions of half1 hold catalogue 1 (gen a1_gen), partitioned under its upload with P / L / pid_base
	u64 a1_gen = 0;
};

struct metro_ctx {
	int device = 0;
	cudaStream_t stream = nullptr;       // stream in use
	cudaStream_t own_stream = nullptr;   // the context's private stream
	cudaStream_t copy_stream = nullptr;  // host -> device copies of an upload (overlapped with the per-chunk checks on `stream`)
	cudaStream_t eager_stream = nullptr; // passes A and B of catalogue 0 while catalogue 1 is being uploaded (stage_partition_eager)
	cudaEvent_t ev_eager = nullptr, ev_eager0 = nullptr;
	bool eager_pending = false;
	bool last_a1 = false;                // the last partition join found catalogue 1 already through pass A
	bool e1_on = false;                  // pass A of catalogue 1 under its upload (partition_eager1_*)
	u32 *e1_g1 = nullptr, *e1_ovf = nullptr;
	cudaEvent_t ev_up[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
	char err[1024] = {0};
	int nptypes = 0;
	DevCatalogue cat[2];
	DevResult res;
	SortWorkspace sortws;
	// scratch (grow-only)
	DevBuf keys_a, keys_b;          // packed join keys, ping-pong
	DevBuf table;                   // pair-count hash table (overflow of the per-halo cache)
	DevBuf pair_cache;              // 4 (tag,count) ways per catalogue-0 halo
	DevBuf part_counts;             // partition path: bucket fill counters + overflow flag
	PartResident resident;
	int part_finer = 0;             // bucket halvings the last successful partition join needed (tried first next time)
	u64 next_gen = 1;
	DevBuf scratch[80];             // select pipeline temporaries
	DevBuf scan_tmp;                // block sums of the scan primitive
	DevBuf counters;                // small device counters (u64[64])
	u64 *h_counters = nullptr;      // pinned mirror
	cudaEvent_t ev[16];
	cudaEvent_t ev_pass[SORT_MAX_PASSES + 1];
	metro_timings timings;
	int launches = 0;
	// multi-GPU (one context per rank); comm is an ncclComm_t
	int rank = 0, world = 1;
	void *comm = nullptr;
	u64 pid_base = 0;               // subtracted from every pid when the join key is packed
	i64 exchange_bytes = 0;         // bytes this rank sent to peers in the last all-to-all
	bool exchange_timed = false;    // ev[11] .. ev[12] bracket the grouped send/recv of the last match
	DevBuf xsend_key, xsend_cnt, xrecvA_key, xrecvA_cnt, xrecvB_key, xrecvB_cnt, xcounts;
	DevBuf xredA_key, xredA_cnt, xredB_key, xredB_cnt;      // owner-side pre-reduced records (select.cu / partjoin.cu: stage_prereduce)
};

int dev_reserve(metro_ctx *ctx, DevBuf &b, size_t bytes);
int dev_reserve_keep(metro_ctx *ctx, DevBuf &b, size_t bytes, size_t keep);
void dev_free(DevBuf &b);

// ---- device primitives (scan.cu) ---------------------------------------------------------------
// out[i] = sum_{j<i} in[j]  (u32 in, u32 out; in == out allowed); *total (device u64) = sum of all
int scan_exclusive_u32(metro_ctx *ctx, cudaStream_t st, const u32 *in, u32 *out, i64 n, u64 *d_total);

// ---- pipeline stages -----------------------------------------------------------------------------
int stage_pack_hist(metro_ctx *ctx, const KeyLayout &L, const SortPlan &plan);      // join.cu

bool partition_join_plan(const metro_ctx *ctx, const KeyLayout &L, i64 N, PartPlan *out, i64 *nmax_cap, int finer, i64 nmax_known = -1);     // partjoin.cu
int stage_partition_eager(metro_ctx *ctx, const metro_params *prm);
int partition_eager_join(metro_ctx *ctx);
int partition_eager1_begin(metro_ctx *ctx, const metro_params *prm, i64 N, i64 H);
int partition_eager1_chunk(metro_ctx *ctx, const u64 *pid, const u32 *halo, const u8 *type, i64 n);
int partition_eager1_flush(metro_ctx *ctx);
void partition_eager1_end(metro_ctx *ctx, const DevCatalogue &c, bool ok);
int stage_partition_join(metro_ctx *ctx, const KeyLayout &L, const PartPlan &P, int rec_halo_bits, bool reuse0, i64 *n_records, i64 *n_hits, bool *fell_back,
			 u32 *overflow_mask, bool finer_plan);
int stage_prereduce(metro_ctx *ctx, const u64 *key, const u32 *cnt, i64 R, int halo_bits, int type_bits, bool swapped, i64 x_lo, i64 x_hi, DevBuf *W,
		    DevBuf &okey, DevBuf &ocnt, u64 *d_total);                              // partjoin.cu (multi-GPU owner side)
int stage_join_count(metro_ctx *ctx, const KeyLayout &L, const u64 *sorted, i64 n,
		     i64 *n_records, i64 *n_hits);                                     // join.cu
int stage_wide_join(metro_ctx *ctx, const KeyLayout &L, int pid_bits, i64 *n_records, i64 *n_hits, int *sort_launches);   // join.cu
int stage_select(metro_ctx *ctx, const metro_params *prm, const KeyLayout &L, const u64 *recA_key, const u32 *recA_cnt, i64 RA,
		 const u64 *recB_key, const u32 *recB_cnt, i64 RB, i64 a_lo, i64 a_hi);           // select.cu
// comm.cu (NCCL; no-ops on a single GPU)
int comm_allreduce_max_i32(metro_ctx *ctx, i32 *buf, i64 n);
int comm_allreduce_sum_u64(metro_ctx *ctx, u64 *buf, i64 n);
int comm_agree_status(metro_ctx *ctx, int rc);
int comm_exchange_records(metro_ctx *ctx, int halo_bits, int type_bits, const u64 *rec_key, const u32 *rec_cnt, i64 R,
			  const u64 **recA_key, const u32 **recA_cnt, i64 *RA, const u64 **recB_key, const u32 **recB_cnt, i64 *RB);
void comm_owned_range(const metro_ctx *ctx, i64 H, i64 *lo, i64 *hi);
void comm_destroy(metro_ctx *ctx);
int stage_shift(metro_ctx *ctx, const metro_params *prm);                            // shift.cu
int merit_bits_device(metro_ctx *ctx, i64 n, const i32 *a, const i32 *b, const i32 *c, const i32 *d, u32 *out);   // select.cu
int catalogue_scan_tuples(metro_ctx *ctx, DevCatalogue &c, int nptypes);             // shift.cu: max pid + validation
