/* metro_oracle.h -- CPU restatement of MetroCPP's progenitor-matching hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity checker for the CUDA path; it
 * may be imported / linked / executed only by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.  The product never
 * calls it and has no CPU fallback.
 *
 * Parity status: the reference ships no tests or golden vectors (SURVEY.md
 * section 4), so this restatement is PINNED AGAINST THE REFERENCE ITSELF: the
 * unmodified sources are compiled here (oracle/Makefile `ref`) and
 * tests/test_oracle_vs_reference.py + tests/golden/ require byte-identical
 * .mtree output on chains covering thresholds, merit ties, tokens (both build
 * flavours), hashed halo IDs and multi-species input.
 *
 * Data model ("tuple" = one line of an AHF _particles file = one membership):
 *   snapshot = H halos (id, nPart[nptypes], nOrphanSteps, isToken) in catalogue
 *              order + N tuples (pid, halo index, type).
 * Catalogue 0 is the LATER snapshot (descendants), catalogue 1 the EARLIER one
 * (progenitors), as in reference src/main.cpp:154-308.
 */
#ifndef METRO_ORACLE_H
#define METRO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_FLAVOUR_GATHER   0   /* -DGATHER_TREES: token tuples copied once          */
#define ORACLE_FLAVOUR_ZOOM_DUP 1   /* -DZOOM (non-GATHER): token tuples duplicated (2^k) */

typedef struct {
	int32_t nptypes;            /* NPTYPES of the build (2 with NOPTYPE, up to 6)          */
	int32_t min_part_cmp;       /* minPartCmp : keep B iff sum_t nCommon > minPartCmp      */
	int32_t min_part_halo;      /* minPartHalo: orphan rule iff nAllPart > minPartHalo     */
	int32_t fac_orphan_steps;   /* facOrphanSteps (must be != 0, as in the reference)      */
	int32_t max_orphan_steps;   /* maxOrphanSteps (0 = unset = tracking off)               */
	int32_t flavour;            /* ORACLE_FLAVOUR_*                                        */
} oracle_params;

typedef struct {
	int64_t n_halos;
	const uint64_t *halo_id;        /* [H]                       */
	const int32_t *npart;           /* [H * nptypes] row-major   */
	const int32_t *n_orphan_steps;  /* [H]                       */
	const uint8_t *is_token;        /* [H]                       */
	int64_t n_tuples;
	const uint64_t *pid;            /* [N]                       */
	const uint32_t *halo;           /* [N] index into the halos  */
	const uint8_t *type;            /* [N] < nptypes             */
} oracle_snapshot;

typedef struct {
	/* raw trees: d = 0 forward (one per halo of catalogue 0, progenitors index catalogue 1),
	 *            d = 1 backward (one per halo of catalogue 1, "progenitors" index catalogue 0);
	 * candidates are in final SortByMerit order, INCLUDING the SortIndexes tie quirk. */
	int64_t raw_count[2];
	int64_t *raw_off[2];            /* [H_d + 1]                 */
	int32_t *raw_prog[2];           /* [M_d]                     */
	int32_t *raw_ncommon[2];        /* [M_d * nptypes]           */
	uint32_t *raw_merit[2];         /* [M_d] fp32 bits; 0 for trees that were not sorted */
	/* clean trees (locCleanTrees[iStep-1]) in catalogue-0 order */
	int64_t n_trees;
	int32_t *tree_main;             /* [n_trees] index into catalogue 0 */
	uint8_t *tree_orphan;           /* [n_trees] isOrphan flag written to the .mtree */
	int64_t *tree_off;              /* [n_trees + 1]             */
	int64_t clean_count;
	int32_t *clean_prog;            /* [clean_count] catalogue-1 index, or -1 = token copy of the main halo */
	int32_t *clean_ncommon;         /* [clean_count * nptypes]   */
	/* orphan bookkeeping */
	int64_t n_tokens;
	int32_t *token_halo;            /* [n_tokens] catalogue-0 indices, detection order */
	int64_t n_untracked;
	int64_t n_hits;                 /* sum over pid of mA*mB */
	int64_t n_pairs;                /* distinct (A,B) with >= 1 shared tuple */
} oracle_result;

/* FindProgenitors(0,1) + FindProgenitors(1,0) + CleanTrees (reference
 * src/MergerTree.cpp:399-499, 133-244, 510-652).  Returns 0 on success. */
int oracle_match_pair(const oracle_params *prm, const oracle_snapshot *cat0,
		      const oracle_snapshot *cat1, oracle_result *out);
void oracle_result_free(oracle_result *r);

/* ShiftHalosPartsGrids (+SyncOrphanHalos): sizes of the next catalogue 0 ... */
void oracle_shift_sizes(const oracle_params *prm, const oracle_snapshot *cat0, const oracle_snapshot *cat1,
			const oracle_result *res, int64_t *n_halos, int64_t *n_tuples);
/* ... and its contents, into caller-allocated arrays of those sizes. */
void oracle_shift_fill(const oracle_params *prm, const oracle_snapshot *cat0, const oracle_snapshot *cat1,
		       const oracle_result *res, uint64_t *halo_id, int32_t *npart, int32_t *n_orphan_steps,
		       uint8_t *is_token, uint64_t *pid, uint32_t *halo, uint8_t *type);

/* MergerTree::SortByMerit merit of one candidate (src/MergerTree.cpp:181-206), fp32 bits. */
uint32_t oracle_merit_bits(int32_t n_comm, int32_t n_ph0, int32_t n_ph1, int32_t i_m);
/* utils SortIndexes + the inverse copy of SortByMerit (src/utils.cpp:429-449,
 * src/MergerTree.cpp:217-229): order[p] = original index placed at position p. */
void oracle_sort_by_merit_order(const uint32_t *merit_bits, int32_t n, int32_t *order);

#ifdef __cplusplus
}
#endif
#endif
