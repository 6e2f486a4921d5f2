/* metro_b200.h -- C ABI of the B200 progenitor-matching engine (libmetro_b200.so).
 *
 * Drop-in boundary for MetroCPP's matching hot path.  The reference has no FFI:
 * its de-facto interface is four free functions over process globals
 * (reference src/MergerTree.h:85-98, src/utils.h:47, callers src/main.cpp:199-301).
 * Each entry point below names the reference call it replaces.  Plain pointers and
 * sizes only; every call returns a status code (0 = METRO_OK) and never exits the
 * process (the reference prints and exit(0)s, src/utils.cpp:74-78).
 *
 * Catalogue 0 is the LATER snapshot (descendants), catalogue 1 the EARLIER one
 * (progenitors), exactly like locHalos[0] / locHalos[1] (src/main.cpp:154-308).
 * A "tuple" is one line of an AHF _particles file: (particle ID, halo, type).
 *
 * Threading: one context per process/GPU; calls on a context are serialised by
 * the caller (the reference is single-threaded per rank too).
 * There is NO CPU fallback: every entry point fails with METRO_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef METRO_B200_H
#define METRO_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define METRO_OK               0
#define METRO_ERR_INVALID      1   /* bad argument (null pointer, index out of range, nptypes ...) */
#define METRO_ERR_CUDA         2   /* CUDA runtime failure, see metro_last_error()               */
#define METRO_ERR_UNSUPPORTED  3   /* e.g. pid bits + halo bits + type bits do not fit 64 bits   */
#define METRO_ERR_STATE        4   /* call order (match before upload, fetch before match ...)   */
#define METRO_ERR_COMM         5   /* NCCL failure                                               */

/* Which build of the reference to reproduce for token (orphan) halos:
 * GATHER = -DGATHER_TREES full-box build, token particle lists copied once
 *          (src/Communication.cpp:519-524);
 * ZOOM_DUP = -DZOOM build, lists are push_back()ed and then appended again, so a token's
 *          tuples double at every orphan step (src/MergerTree.cpp:597-603). */
#define METRO_FLAVOUR_GATHER   0
#define METRO_FLAVOUR_ZOOM_DUP 1

typedef struct metro_ctx metro_ctx;

/* Run-time settings of the path == the globals the reference reads
 * (src/global_vars.h:92-118; parsed by IOSettings::InitFromCfgFile, src/IOSettings.cpp:153-175). */
typedef struct {
	int32_t nptypes;            /* NPTYPES (compile-time in the reference), 1..8            */
	int32_t min_part_cmp;       /* minPartCmp                                               */
	int32_t min_part_halo;      /* minPartHalo                                              */
	int32_t fac_orphan_steps;   /* facOrphanSteps, must be != 0                             */
	int32_t max_orphan_steps;   /* maxOrphanSteps, 0 = unset = orphan tracking off          */
	int32_t flavour;            /* METRO_FLAVOUR_*                                          */
} metro_params;

/* One catalogue as the reference holds it after ReadHalos + ReadParticles
 * (locHalos[i], locMapParts[i]; src/IOSettings.cpp:551-811).  Host pointers. */
typedef struct {
	int64_t n_halos;
	const uint64_t *halo_id;        /* [H]  Halo::ID                                        */
	const int32_t *npart;           /* [H * nptypes]  Halo::nPart                           */
	const int32_t *n_orphan_steps;  /* [H]  Halo::nOrphanSteps (may be NULL = all 0)        */
	const uint8_t *is_token;        /* [H]  Halo::isToken      (may be NULL = all 0)        */
	int64_t n_tuples;
	const uint64_t *pid;            /* [N]  particle IDs                                    */
	const uint32_t *halo;           /* [N]  index of the owning halo in this catalogue      */
	const uint8_t *type;            /* [N]  particle type < nptypes (NULL = all 1, NOPTYPE) */
} metro_catalogue;

/* Sizes of the result of the last metro_match_pair(). */
typedef struct {
	int64_t n_halos[2];             /* halos of catalogue 0 / 1                              */
	int64_t raw_count[2];           /* candidates in the forward / backward raw trees        */
	int64_t n_trees;                /* clean trees (locCleanTrees[iStep-1].size())           */
	int64_t clean_count;            /* progenitor entries over all clean trees               */
	int64_t n_tokens;               /* orphan halos that become token halos                  */
	int64_t n_untracked;            /* orphans dropped ("untracked", src/MergerTree.cpp:614) */
	int64_t n_hits;                 /* sum over pid of mA*mB                                 */
	int64_t n_pairs;                /* distinct (A,B) halo pairs with >= 1 shared tuple      */
	int64_t n_tuples[2];
	int32_t pid_bits, halo_bits, type_bits, sort_passes;
	int64_t n_table_hits;           /* hits that missed the per-halo counter ways (hash-table path) */
	int64_t exchange_bytes;         /* multi-GPU: bytes this rank sent to its peers in the all-to-all */
	int64_t owned_lo, owned_hi;     /* catalogue-0 halo range whose clean trees this rank holds */
} metro_result_info;

/* Host destination arrays for metro_result_fetch(); any pointer may be NULL (skipped).
 * Layout mirrors the MergerTree fields (src/MergerTree.h:30-51) in CSR form. */
typedef struct {
	/* raw trees locMTrees[d]: d = 0 forward (per halo of catalogue 0, entries index catalogue 1),
	 * d = 1 backward.  Entries are in final SortByMerit order incl. the SortIndexes tie quirk. */
	int64_t *raw_off[2];            /* [H_d + 1]                                             */
	int32_t *raw_prog[2];           /* [raw_count[d]]                                        */
	int32_t *raw_ncommon[2];        /* [raw_count[d] * nptypes]   MergerTree::nCommon        */
	uint32_t *raw_merit[2];         /* [raw_count[d]] fp32 bits (0 where the tree was not sorted) */
	/* clean trees locCleanTrees[iStep-1], in catalogue-0 order */
	int32_t *tree_main;             /* [n_trees] index into catalogue 0 (mainHalo)           */
	uint8_t *tree_orphan;           /* [n_trees] MergerTree::isOrphan                        */
	int64_t *tree_off;              /* [n_trees + 1]                                         */
	int32_t *clean_prog;            /* [clean_count] catalogue-1 index; -1 = token copy of mainHalo */
	int32_t *clean_ncommon;         /* [clean_count * nptypes]                               */
	int32_t *token_halo;            /* [n_tokens] catalogue-0 indices, catalogue order       */
} metro_result_arrays;

/* Device timings (CUDA events on the context's stream) of the last metro_match_pair(), ms. */
typedef struct {
	float total, pack_hist, sort, join, select, exchange;
	int32_t sort_launches, total_launches;
	float sort_pass_ms[8];          /* LSD path: per onesweep pass; partition path: [0] = pass A (pack + partition), [1] = pass B (partition),
	                                 * [2] = pass C (bucket join, hits out), [3] = pass D (hit count) */
	int32_t path;                   /* 0 = LSD radix sort + run-length join, 1 = partition join, 2 = wide-key sort (IDs too wide for the packed key) */
	int32_t resident;               /* partition join, bit 0: the key buckets of catalogue 0 were resident - from the previous pair of the chain,
	                                 * or built while catalogue 1 was being uploaded (only catalogue 1 and the token tuples were
	                                 * partitioned); bit 1: catalogue 1 had been through pass A chunk by chunk under its own upload */
	float exchange_sendrecv;        /* multi-GPU: the grouped ncclSend/ncclRecv all-to-all alone (payload over NVLink; includes waiting
	                                 * for the slowest peer to arrive); `exchange` also covers routing counts and the scatter */
} metro_timings;

/* ---- context ------------------------------------------------------------------------- */
/* Replaces InitLocVariables + InitTrees (src/utils.cpp:43-67, src/MergerTree.cpp:503-506). */
int metro_ctx_create(int device, metro_ctx **out);
void metro_ctx_destroy(metro_ctx *ctx);
const char *metro_last_error(const metro_ctx *ctx);   /* ctx may be NULL: last create error */
/* Run all work of this context on the caller's CUDA stream (a cudaStream_t, e.g. the stream the
 * caller's own events are recorded on); NULL restores the context's private stream. */
int metro_ctx_set_stream(metro_ctx *ctx, void *cuda_stream);
const char *metro_version(void);

/* ---- catalogues ---------------------------------------------------------------------- */
/* Replaces the map build of IOSettings::ReadParticles (locMapParts[slot][pid].push_back,
 * src/IOSettings.cpp:628-632) and the locHalos[slot] fill (:775-793): copies the host arrays
 * to HBM.  slot in {0,1}.  Caller keeps ownership of the host arrays. */
int metro_catalogue_upload(metro_ctx *ctx, int slot, const metro_params *prm, const metro_catalogue *cat);
/* Same, for tuples grouped by halo the way an _particles file and locParts[slot][halo] hold them
 * (src/global_vars.h:64, filled src/IOSettings.cpp:645-668): the tuples of halo h are
 * pid[halo_offset[h] .. halo_offset[h+1]), halo_offset[0] = 0, halo_offset[n_halos] = n_tuples, non-decreasing;
 * cat->halo is ignored (may be NULL) and the per-tuple halo index is rebuilt on the device: 8 instead of
 * 12 bytes per tuple cross the bus. */
int metro_catalogue_upload_csr(metro_ctx *ctx, int slot, const metro_params *prm, const metro_catalogue *cat, const int64_t *halo_offset);
/* Same as metro_catalogue_upload_csr with the particle IDs as 32-bit offsets from pid_base (ID of tuple i = pid_base + pid32[i]):
 * the form a reader can hold them in whenever the IDs it reads span less than 2^32 (a box of up to 1600^3 particles, or the
 * particle-ID range of one rank of a larger one): 4 instead of 8 bytes per tuple cross the bus; the IDs are widened to 64 bits on
 * the device, chunk by chunk, under the copy.  cat->pid and cat->halo are ignored (may be NULL). */
int metro_catalogue_upload_csr32(metro_ctx *ctx, int slot, const metro_params *prm, const metro_catalogue *cat, const uint32_t *pid32,
				 uint64_t pid_base, const int64_t *halo_offset);
/* Same, but the tuple arrays are already DEVICE pointers (zero copy; the context borrows them
 * until the next upload/shift of that slot).  Halo arrays stay host pointers. */
int metro_catalogue_adopt_device(metro_ctx *ctx, int slot, const metro_params *prm, const metro_catalogue *cat);

/* ---- the hot path -------------------------------------------------------------------- */
/* Replaces FindProgenitors(0,1); FindProgenitors(1,0); CleanTrees(iStep)
 * (src/MergerTree.cpp:399-499, 133-244, 510-652; called from src/main.cpp:199,217,254|280). */
int metro_match_pair(metro_ctx *ctx, const metro_params *prm, metro_result_info *info);
/* Copies the result of the last match to host arrays (what WriteTree consumes,
 * src/IOSettings.cpp:1074-1138). */
int metro_result_fetch(metro_ctx *ctx, const metro_result_arrays *dst);
int metro_get_timings(const metro_ctx *ctx, metro_timings *out);

/* Replaces ShiftHalosPartsGrids (+ SyncOrphanHalos) (src/utils.cpp:245-403,
 * src/Communication.cpp:477-537): catalogue 1 becomes catalogue 0 on the device and the token
 * halos of the last match are appended with their tuples (weight per prm->flavour). */
int metro_shift(metro_ctx *ctx, const metro_params *prm);
/* Read back catalogue `slot` halo table sizes / contents (after a shift the host needs the token
 * halos it must append to locHalos[0]). */
int metro_catalogue_info(metro_ctx *ctx, int slot, int64_t *n_halos, int64_t *n_tuples);
int metro_catalogue_fetch_halos(metro_ctx *ctx, int slot, uint64_t *halo_id, int32_t *npart,
				int32_t *n_orphan_steps, uint8_t *is_token);
int metro_catalogue_fetch_tuples(metro_ctx *ctx, int slot, uint64_t *pid, uint32_t *halo, uint8_t *type);
/* Overwrite Halo::nPart of a resident catalogue ([H * nptypes], host pointer).  Needed when the
 * tuples of a catalogue are sharded over several GPUs by particle-ID range: the catalogue npart
 * entering the merit is the halo's GLOBAL size, which only the caller can sum over the shards. */
int metro_catalogue_set_npart(metro_ctx *ctx, int slot, const int32_t *npart);

/* ---- multi-GPU: one context per rank / GPU, NCCL over NVLink ---------------------------------
 * Replaces the MPI layer of the path (src/Communication.cpp, src/Grid.cpp).  The caller shards the
 * TUPLES of both catalogues by particle-ID range (any partition that keeps all memberships of a
 * particle ID on one rank) and uploads the FULL halo tables on every rank.  metro_match_pair then
 * joins locally, sends every partial (A,B,types,count) record to the ranks owning haloA / haloB
 * (contiguous halo-index ranges) with one grouped all-to-all, and all-reduces the best-descendant
 * and token arrays.  Each rank's result covers the clean trees of ITS range of catalogue-0 halos
 * (concatenate in rank order); token_halo, n_tokens are global.  metro_shift appends the token
 * tuples found in the rank's own shard.
 * rank 0 calls metro_comm_unique_id, ships the 128 bytes to the others, all call metro_comm_init. */
int metro_comm_unique_id(void *out128);
int metro_comm_init(metro_ctx *ctx, int rank, int world, const void *unique_id128);

/* ---- building blocks exposed for tests / benchmarks ------------------------------------ */
/* LSD radix sort of n u64 keys (optionally with u32 values) on bits [begin_bit, end_bit);
 * device pointers; keys_out/vals_out receive the result; *_in are clobbered. */
int metro_sort_u64(metro_ctx *ctx, uint64_t *keys_in, uint64_t *keys_out, uint32_t *vals_in,
		   uint32_t *vals_out, int64_t n, int begin_bit, int end_bit);
/* merit of MergerTree::SortByMerit computed ON THE DEVICE for n candidates (device arithmetic
 * known-answer test; src/MergerTree.cpp:181-206).  Host pointers. */
int metro_merit_bits(metro_ctx *ctx, int64_t n, const int32_t *n_comm, const int32_t *n_ph0,
		     const int32_t *n_ph1, const int32_t *i_m, uint32_t *bits_out);

/* ---- synthetic AHF-like catalogues, deterministic --------------- */
typedef struct {
	uint64_t seed;
	int64_t n_halos;        /* halos of the later catalogue                                   */
	int64_t n_tuples;       /* target tuples of the later catalogue (approximate)             */
	int32_t pid_bits;       /* particle IDs are a (cycle-walking) Feistel permutation over 2^pid_bits */
	int32_t shard, n_shards;/* keep only pids whose top bits fall in this shard (pid range)   */
} metro_synth_spec;
/* Generates BOTH catalogues of a pair directly in HBM and installs them in slots 0 and 1. */
int metro_synth_pair_device(metro_ctx *ctx, const metro_params *prm, const metro_synth_spec *spec);
/* The same pair on the host (bit-identical tuples, possibly in a different order); two calls:
 * sizes first (arrays NULL), then fill. */
int metro_synth_pair_host(const metro_params *prm, const metro_synth_spec *spec, int slot,
			  int64_t *n_halos, int64_t *n_tuples, uint64_t *halo_id, int32_t *npart,
			  uint64_t *pid, uint32_t *halo, uint8_t *type);

#ifdef __cplusplus
}
#endif
#endif
