/* metro_oracle.c -- CPU restatement of MetroCPP's progenitor matching.
 * TEST INFRASTRUCTURE ONLY (see metro_oracle.h for the usage rule and the
 * parity-pinning statement).
 *
 * The reference joins the two catalogues through std::map<pid, vector<Particle>>
 * probed particle by particle (reference src/MergerTree.cpp:443-472); this file
 * computes the SAME multiset of (haloA, haloB, type) hits with a sort-merge join
 * and then follows the reference's selection rules line by line.
 */
#include "metro_oracle.h"

#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ helpers */

static void *xmalloc(size_t n) { void *p = malloc(n ? n : 1); if (!p) abort(); return p; }
static void *xcalloc(size_t n, size_t s) { void *p = calloc(n ? n : 1, s ? s : 1); if (!p) abort(); return p; }

/* Stable LSD radix sort of (u64 key, u32 value) pairs; digits shared by all keys are skipped. */
static void radix_sort_kv(uint64_t *key, uint32_t *val, int64_t n)
{
	if (n < 2) return;
	uint64_t *k2 = xmalloc((size_t) n * sizeof *k2);
	uint32_t *v2 = xmalloc((size_t) n * sizeof *v2);
	uint64_t *ka = key, *kb = k2;
	uint32_t *va = val, *vb = v2;
	for (int pass = 0; pass < 8; pass++) {
		int64_t hist[256] = {0};
		const int sh = pass * 8;
		for (int64_t i = 0; i < n; i++) hist[(ka[i] >> sh) & 255]++;
		int trivial = 0;
		for (int d = 0; d < 256; d++) if (hist[d] == n) trivial = 1;
		if (trivial) continue;
		int64_t sum = 0;
		for (int d = 0; d < 256; d++) { int64_t c = hist[d]; hist[d] = sum; sum += c; }
		for (int64_t i = 0; i < n; i++) {
			int64_t dst = hist[(ka[i] >> sh) & 255]++;
			kb[dst] = ka[i]; vb[dst] = va[i];
		}
		uint64_t *tk = ka; ka = kb; kb = tk;
		uint32_t *tv = va; va = vb; vb = tv;
	}
	if (ka != key) { memcpy(key, ka, (size_t) n * sizeof *key); memcpy(val, va, (size_t) n * sizeof *val); }
	free(k2); free(v2);
}

static int32_t npart_sum(const oracle_snapshot *s, int nptypes, int64_t h)
{
	/* Halo::nAllPart (src/Halo.cpp:63-70) and the nPH0/nPH1 sums of SortByMerit (src/MergerTree.cpp:190-194) */
	int32_t t = 0;
	for (int i = 0; i < nptypes; i++) t += s->npart[h * nptypes + i];
	return t;
}

/* ------------------------------------------------------------------ merit */

/* MergerTree::SortByMerit body (reference src/MergerTree.cpp:181-206): fp32 divide of the
 * int->float rounded sizes, fp64 reciprocal / mul / sub / div, two roundings to fp32. */
static float merit_value(int32_t nComm, int32_t nPH0, int32_t nPH1, int32_t iM)
{
	float f0 = (float) nPH0, f1 = (float) nPH1;
	float q = f0 / f1;
	double ratioM = q;
	if (ratioM < 1.0) ratioM = 1.0 / ratioM;
	double a = ratioM * 1.01;
	double b = a - 1.0;
	float merit = (float) ((double) nComm / b);
	double c = 0.00001 * (double) iM;
	double d = 1.0 + c;
	merit = (float) ((double) merit * d);
	return merit;
}

uint32_t oracle_merit_bits(int32_t n_comm, int32_t n_ph0, int32_t n_ph1, int32_t i_m)
{
	float m = merit_value(n_comm, n_ph0, n_ph1, i_m);
	uint32_t u; memcpy(&u, &m, 4);
	return u;
}

typedef struct { float v; int32_t i; } merit_item;

static int merit_cmp(const void *pa, const void *pb)
{
	const merit_item *a = pa, *b = pb;
	if (a->v < b->v) return -1;
	if (a->v > b->v) return 1;
	return (a->i > b->i) - (a->i < b->i);
}

/* utils::SortIndexes (src/utils.cpp:429-449): map<float,int>::insert keeps the FIRST index seen
 * for a value, the values are sorted ascending, idx[k] = that first index; SortByMerit then
 * writes element idx[k] to position n-1-k (src/MergerTree.cpp:217-229).  Equal merits therefore
 * duplicate the lowest-index candidate and drop the others. */
void oracle_sort_by_merit_order(const uint32_t *merit_bits, int32_t n, int32_t *order)
{
	merit_item *it = xmalloc((size_t) n * sizeof *it);
	for (int32_t i = 0; i < n; i++) { memcpy(&it[i].v, &merit_bits[i], 4); it[i].i = i; }
	qsort(it, (size_t) n, sizeof *it, merit_cmp);
	int32_t first = 0;
	for (int32_t k = 0; k < n; k++) {
		if (k > 0 && it[k].v == it[k - 1].v) { /* same key: first index of the run */ }
		else first = it[k].i;
		order[n - 1 - k] = first;
	}
	free(it);
}

/* ------------------------------------------------------------------ raw trees */

typedef struct {
	int64_t n;           /* candidates */
	int64_t *off;        /* [H+1] */
	int32_t *prog;
	int32_t *ncommon;    /* [n * nptypes] */
	uint32_t *merit;
} raw_tree;

/* Builds one direction.  hits: key = main<<32 | idRank(other), val = type on the OTHER side.
 * main_s: catalogue of the tree owners; other_s: catalogue of the candidates;
 * rank2idx: inverse of idRank for the other catalogue. */
static void build_raw(const oracle_params *prm, const oracle_snapshot *main_s, const oracle_snapshot *other_s,
		      uint64_t *hkey, uint32_t *htype, int64_t n_hits, const int32_t *rank2idx,
		      raw_tree *out, int64_t *n_pairs)
{
	const int T = prm->nptypes;
	const int64_t H = main_s->n_halos;
	radix_sort_kv(hkey, htype, n_hits);
	/* run-length: distinct (main, other) pairs in (main asc, otherID asc) order ==
	 * iteration order of map<uint64_t, vector<int>> indexCommon (src/MergerTree.cpp:138) */
	int64_t M = 0;
	for (int64_t i = 0; i < n_hits; i++) if (i == 0 || hkey[i] != hkey[i - 1]) M++;
	*n_pairs = M;
	int32_t *p_main = xmalloc((size_t) M * sizeof *p_main);
	int32_t *p_other = xmalloc((size_t) M * sizeof *p_other);
	int32_t *p_cnt = xcalloc((size_t) M * T, sizeof *p_cnt);
	int64_t m = -1;
	for (int64_t i = 0; i < n_hits; i++) {
		if (i == 0 || hkey[i] != hkey[i - 1]) {
			m++;
			p_main[m] = (int32_t) (hkey[i] >> 32);
			p_other[m] = rank2idx[(uint32_t) hkey[i]];
		}
		p_cnt[m * T + htype[i]]++;
	}
	out->off = xcalloc((size_t) H + 1, sizeof *out->off);
	out->prog = xmalloc((size_t) M * sizeof *out->prog);
	out->ncommon = xmalloc((size_t) M * T * sizeof *out->ncommon);
	out->merit = xcalloc((size_t) M, sizeof *out->merit);
	int64_t w = 0, i = 0;
	for (int64_t h = 0; h < H; h++) {
		out->off[h] = w;
		int64_t j = i;
		while (j < M && p_main[j] == h) j++;
		/* MergerTree::AssignMap (src/MergerTree.cpp:133-169): strict threshold, ascending ID order */
		int64_t start = w;
		for (int64_t k = i; k < j; k++) {
			int32_t tot = 0;
			for (int t = 0; t < T; t++) tot += p_cnt[k * T + t];
			if (tot > prm->min_part_cmp) {
				out->prog[w] = p_other[k];
				memcpy(&out->ncommon[w * T], &p_cnt[k * T], (size_t) T * sizeof(int32_t));
				w++;
			}
		}
		int32_t n = (int32_t) (w - start);
		if (n > 1) {
			/* SortByMerit, only with >= 2 progenitors (src/MergerTree.cpp:496-497) */
			int32_t nPH0 = npart_sum(main_s, T, h);
			uint32_t *mb = xmalloc((size_t) n * sizeof *mb);
			int32_t *ord = xmalloc((size_t) n * sizeof *ord);
			for (int32_t c = 0; c < n; c++) {
				int32_t nc = 0;
				for (int t = 0; t < T; t++) nc += out->ncommon[(start + c) * T + t];
				mb[c] = oracle_merit_bits(nc, nPH0, npart_sum(other_s, T, out->prog[start + c]), c);
			}
			oracle_sort_by_merit_order(mb, n, ord);
			int32_t *tp = xmalloc((size_t) n * sizeof *tp);
			int32_t *tc = xmalloc((size_t) n * T * sizeof *tc);
			for (int32_t p = 0; p < n; p++) {
				tp[p] = out->prog[start + ord[p]];
				memcpy(&tc[p * T], &out->ncommon[(start + ord[p]) * T], (size_t) T * sizeof(int32_t));
				out->merit[start + p] = mb[ord[p]];
			}
			memcpy(&out->prog[start], tp, (size_t) n * sizeof *tp);
			memcpy(&out->ncommon[start * T], tc, (size_t) n * T * sizeof *tc);
			free(mb); free(ord); free(tp); free(tc);
		}
		i = j;
	}
	out->off[H] = w;
	out->n = w;
	free(p_main); free(p_other); free(p_cnt);
}

/* argsort of the halo IDs: rank[h] = position of halo h in ascending-ID order */
static void id_ranks(const oracle_snapshot *s, uint32_t *rank, int32_t *rank2idx)
{
	int64_t H = s->n_halos;
	uint64_t *k = xmalloc((size_t) H * sizeof *k);
	uint32_t *v = xmalloc((size_t) H * sizeof *v);
	for (int64_t h = 0; h < H; h++) { k[h] = s->halo_id[h]; v[h] = (uint32_t) h; }
	radix_sort_kv(k, v, H);
	for (int64_t r = 0; r < H; r++) { rank[v[r]] = (uint32_t) r; rank2idx[r] = (int32_t) v[r]; }
	free(k); free(v);
}

int oracle_match_pair(const oracle_params *prm, const oracle_snapshot *cat0,
		      const oracle_snapshot *cat1, oracle_result *out)
{
	memset(out, 0, sizeof *out);
	const int T = prm->nptypes;
	if (T < 1 || T > 8 || prm->fac_orphan_steps == 0) return 1;
	const oracle_snapshot *cat[2] = {cat0, cat1};
	for (int c = 0; c < 2; c++)
		for (int64_t i = 0; i < cat[c]->n_tuples; i++)
			if (cat[c]->halo[i] >= (uint64_t) cat[c]->n_halos || cat[c]->type[i] >= T) return 2;

	uint32_t *rank[2]; int32_t *r2i[2];
	uint64_t *sp[2]; uint32_t *si[2];
	for (int c = 0; c < 2; c++) {
		rank[c] = xmalloc((size_t) cat[c]->n_halos * sizeof(uint32_t));
		r2i[c] = xmalloc((size_t) cat[c]->n_halos * sizeof(int32_t));
		id_ranks(cat[c], rank[c], r2i[c]);
		int64_t N = cat[c]->n_tuples;
		sp[c] = xmalloc((size_t) N * sizeof(uint64_t));
		si[c] = xmalloc((size_t) N * sizeof(uint32_t));
		memcpy(sp[c], cat[c]->pid, (size_t) N * sizeof(uint64_t));
		for (int64_t i = 0; i < N; i++) si[c][i] = (uint32_t) i;
		radix_sort_kv(sp[c], si[c], N);
	}
	/* the join: for every pid present in both catalogues, the cross product of its
	 * memberships (src/MergerTree.cpp:443-472).  Pass 0 counts, pass 1 fills. */
	uint64_t *fk = NULL, *bk = NULL; uint32_t *ft = NULL, *bt = NULL;
	int64_t n_hits = 0;
	for (int pass = 0; pass < 2; pass++) {
		int64_t i = 0, j = 0, w = 0;
		const int64_t NA = cat0->n_tuples, NB = cat1->n_tuples;
		while (i < NA && j < NB) {
			if (sp[0][i] < sp[1][j]) { i++; continue; }
			if (sp[0][i] > sp[1][j]) { j++; continue; }
			uint64_t p = sp[0][i];
			int64_t i1 = i, j1 = j;
			while (i1 < NA && sp[0][i1] == p) i1++;
			while (j1 < NB && sp[1][j1] == p) j1++;
			if (pass == 0) w += (i1 - i) * (j1 - j);
			else
				for (int64_t a = i; a < i1; a++)
					for (int64_t b = j; b < j1; b++) {
						uint32_t ta = si[0][a], tb = si[1][b];
						uint32_t hA = cat0->halo[ta], hB = cat1->halo[tb];
						fk[w] = ((uint64_t) hA << 32) | rank[1][hB]; ft[w] = cat1->type[tb];
						bk[w] = ((uint64_t) hB << 32) | rank[0][hA]; bt[w] = cat0->type[ta];
						w++;
					}
			i = i1; j = j1;
		}
		if (pass == 0) {
			n_hits = w;
			fk = xmalloc((size_t) w * 8); bk = xmalloc((size_t) w * 8);
			ft = xmalloc((size_t) w * 4); bt = xmalloc((size_t) w * 4);
		}
	}
	out->n_hits = n_hits;
	raw_tree raw[2];
	int64_t np0 = 0, np1 = 0;
	build_raw(prm, cat0, cat1, fk, ft, n_hits, r2i[1], &raw[0], &np0);
	build_raw(prm, cat1, cat0, bk, bt, n_hits, r2i[0], &raw[1], &np1);
	out->n_pairs = np0;
	free(fk); free(bk); free(ft); free(bt);
	for (int c = 0; c < 2; c++) { free(rank[c]); free(r2i[c]); free(sp[c]); free(si[c]); }
	for (int d = 0; d < 2; d++) {
		out->raw_count[d] = raw[d].n; out->raw_off[d] = raw[d].off; out->raw_prog[d] = raw[d].prog;
		out->raw_ncommon[d] = raw[d].ncommon; out->raw_merit[d] = raw[d].merit;
	}

	/* CleanTrees (src/MergerTree.cpp:510-652) */
	const int64_t H0 = cat0->n_halos;
	int64_t cap = raw[0].n + H0 + 1;
	out->tree_main = xmalloc((size_t) (H0 + 1) * sizeof(int32_t));
	out->tree_orphan = xmalloc((size_t) (H0 + 1));
	out->tree_off = xmalloc((size_t) (H0 + 2) * sizeof(int64_t));
	out->clean_prog = xmalloc((size_t) cap * sizeof(int32_t));
	out->clean_ncommon = xmalloc((size_t) cap * T * sizeof(int32_t));
	out->token_halo = xmalloc((size_t) (H0 + 1) * sizeof(int32_t));
	int64_t nt = 0, w = 0;
	for (int64_t a = 0; a < H0; a++) {
		int64_t start = w;
		for (int64_t c = raw[0].off[a]; c < raw[0].off[a + 1]; c++) {
			int32_t b = raw[0].prog[c];
			/* descID = locMTrees[1][jTree].progHalo[0].ID  (:545); the backward tree of b is never
			 * empty here because the type-summed count of a pair is the same in both directions */
			int64_t bo = raw[1].off[b];
			if (bo == raw[1].off[b + 1]) continue;
			int32_t desc = raw[1].prog[bo];
			if (cat0->halo_id[desc] == cat0->halo_id[a]) {          /* mainID == descID (:560) */
				out->clean_prog[w] = b;
				memcpy(&out->clean_ncommon[w * T], &raw[0].ncommon[c * T], (size_t) T * sizeof(int32_t));
				w++;
			}
		}
		uint8_t orphan = 0;
		int32_t nall = npart_sum(cat0, T, a);
		if (w == start && nall > prm->min_part_halo) {                  /* (:575) */
			int32_t steps = cat0->n_orphan_steps[a] + 1;
			int32_t loc = 1 + nall / prm->fac_orphan_steps;          /* (:581) int division */
			if (loc > prm->max_orphan_steps) loc = prm->max_orphan_steps;
			if (steps <= loc) {                                         /* (:588) */
				orphan = 1;
				out->clean_prog[w] = -1;
				for (int t = 0; t < T; t++) out->clean_ncommon[w * T + t] = cat0->npart[a * T + t];
				w++;
				out->token_halo[out->n_tokens++] = (int32_t) a;
			} else {
				out->n_untracked++;
			}
		}
		int32_t n = (int32_t) (w - start);
		if (n > 1) {
			/* second SortByMerit (:627-628): iM restarts from the position in the kept list */
			uint32_t *mb = xmalloc((size_t) n * sizeof *mb);
			int32_t *ord = xmalloc((size_t) n * sizeof *ord);
			for (int32_t c = 0; c < n; c++) {
				int32_t nc = 0;
				for (int t = 0; t < T; t++) nc += out->clean_ncommon[(start + c) * T + t];
				mb[c] = oracle_merit_bits(nc, nall, npart_sum(cat1, T, out->clean_prog[start + c]), c);
			}
			oracle_sort_by_merit_order(mb, n, ord);
			int32_t *tp = xmalloc((size_t) n * sizeof *tp);
			int32_t *tc = xmalloc((size_t) n * T * sizeof *tc);
			for (int32_t p = 0; p < n; p++) {
				tp[p] = out->clean_prog[start + ord[p]];
				memcpy(&tc[p * T], &out->clean_ncommon[(start + ord[p]) * T], (size_t) T * sizeof(int32_t));
			}
			memcpy(&out->clean_prog[start], tp, (size_t) n * sizeof *tp);
			memcpy(&out->clean_ncommon[start * T], tc, (size_t) n * T * sizeof *tc);
			free(mb); free(ord); free(tp); free(tc);
		}
		if (n > 0) {                                                    /* (:630-631) */
			out->tree_main[nt] = (int32_t) a;
			out->tree_orphan[nt] = orphan;
			out->tree_off[nt] = start;
			nt++;
		}
	}
	out->tree_off[nt] = w;
	out->n_trees = nt;
	out->clean_count = w;
	return 0;
}

void oracle_result_free(oracle_result *r)
{
	for (int d = 0; d < 2; d++) {
		free(r->raw_off[d]); free(r->raw_prog[d]); free(r->raw_ncommon[d]); free(r->raw_merit[d]);
	}
	free(r->tree_main); free(r->tree_orphan); free(r->tree_off);
	free(r->clean_prog); free(r->clean_ncommon); free(r->token_halo);
	memset(r, 0, sizeof *r);
}

/* ------------------------------------------------------------------ shift */

/* token tuple multiplicity: the non-GATHER flavour push_back()s the halo's lists and then
 * appends them again (src/MergerTree.cpp:597-603) => every orphan step doubles them; the
 * GATHER flavour assigns (src/Communication.cpp:519-524). */
static int token_weight(const oracle_params *prm) { return prm->flavour == ORACLE_FLAVOUR_ZOOM_DUP ? 2 : 1; }

void oracle_shift_sizes(const oracle_params *prm, const oracle_snapshot *cat0, const oracle_snapshot *cat1,
			const oracle_result *res, int64_t *n_halos, int64_t *n_tuples)
{
	int64_t *cnt = xcalloc((size_t) cat0->n_halos, sizeof *cnt);
	for (int64_t i = 0; i < cat0->n_tuples; i++) cnt[cat0->halo[i]]++;
	int64_t extra = 0;
	for (int64_t k = 0; k < res->n_tokens; k++) extra += cnt[res->token_halo[k]] * token_weight(prm);
	*n_halos = cat1->n_halos + res->n_tokens;
	*n_tuples = cat1->n_tuples + extra;
	free(cnt);
}

/* ShiftHalosPartsGrids (src/utils.cpp:245-403): catalogue 1 becomes catalogue 0 and the token
 * halos are appended after the real ones, with their particle lists under the token's index. */
void oracle_shift_fill(const oracle_params *prm, const oracle_snapshot *cat0, const oracle_snapshot *cat1,
		       const oracle_result *res, uint64_t *halo_id, int32_t *npart, int32_t *n_orphan_steps,
		       uint8_t *is_token, uint64_t *pid, uint32_t *halo, uint8_t *type)
{
	const int T = prm->nptypes;
	const int64_t H1 = cat1->n_halos;
	memcpy(halo_id, cat1->halo_id, (size_t) H1 * sizeof *halo_id);
	memcpy(npart, cat1->npart, (size_t) H1 * T * sizeof *npart);
	memcpy(n_orphan_steps, cat1->n_orphan_steps, (size_t) H1 * sizeof *n_orphan_steps);
	memcpy(is_token, cat1->is_token, (size_t) H1);
	memcpy(pid, cat1->pid, (size_t) cat1->n_tuples * sizeof *pid);
	memcpy(halo, cat1->halo, (size_t) cat1->n_tuples * sizeof *halo);
	memcpy(type, cat1->type, (size_t) cat1->n_tuples);
	int32_t *slot = xmalloc((size_t) cat0->n_halos * sizeof *slot);
	for (int64_t h = 0; h < cat0->n_halos; h++) slot[h] = -1;
	for (int64_t k = 0; k < res->n_tokens; k++) {
		int32_t a = res->token_halo[k];
		slot[a] = (int32_t) k;
		halo_id[H1 + k] = cat0->halo_id[a];                        /* same ID (:577-579) */
		memcpy(&npart[(H1 + k) * T], &cat0->npart[(int64_t) a * T], (size_t) T * sizeof *npart);
		n_orphan_steps[H1 + k] = cat0->n_orphan_steps[a] + 1;
		is_token[H1 + k] = 1;
	}
	int64_t w = cat1->n_tuples;
	const int rep = token_weight(prm);
	for (int64_t i = 0; i < cat0->n_tuples; i++) {
		int32_t k = slot[cat0->halo[i]];
		if (k < 0) continue;
		for (int r = 0; r < rep; r++) {
			pid[w] = cat0->pid[i]; halo[w] = (uint32_t) (H1 + k); type[w] = cat0->type[i];
			w++;
		}
	}
	free(slot);
}
