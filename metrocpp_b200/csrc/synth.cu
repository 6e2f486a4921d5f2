// synth.cu -- synthetic snapshot pairs generated directly in HBM (and, identically, on the host).
// See synth.h for the model.  Used by bench.py (workloads too large to build in Python) and by
// the parity tests (device generator == host generator; the CPU checker runs on the host copy).
#include <math.h>

#include <vector>

#include "context.cuh"
#include "synth.h"

struct SynOut {
	int n0, n1;
	uint32_t h0[2], h1[2];
	uint8_t t0[2], t1[2];
};

// all tuples (both catalogues) that carry particle ID v
SYN_HD void syn_tuples_of_pid(const SynthTable &t, uint64_t v, SynOut *o)
{
	o->n0 = o->n1 = 0;
	const uint64_t r = syn_feistel_inv(v, t.seed, t.pid_bits);
	if (r < t.R_tot) {
		const int64_t h = syn_find_halo(t, r);
		const uint8_t ty = syn_type_later(t, v);
		int64_t g[2] = {h, h};
		int ng = 1;
		if (t.flags[h] & SYN_F_SUB) { g[1] = t.host[h]; ng = 2; }
		for (int k = 0; k < ng; k++) {
			o->h0[o->n0] = (uint32_t) g[k]; o->t0[o->n0] = ty; o->n0++;
			uint32_t e;
			if (syn_earlier(t, v, g[k], &e)) { o->h1[o->n1] = e; o->t1[o->n1] = syn_type_earlier(t, v, ty); o->n1++; }
		}
	} else if (r < 2 * t.R_tot) {
		// v is the fresh pid standing in for later-catalogue rank r - R_tot
		const uint64_t r0 = r - t.R_tot;
		const uint64_t p0 = syn_feistel(r0, t.seed, t.pid_bits);
		const int64_t h = syn_find_halo(t, r0);
		const uint8_t ty = syn_type_earlier(t, p0, syn_type_later(t, p0));
		int64_t g[2] = {h, h};
		int ng = 1;
		if (t.flags[h] & SYN_F_SUB) { g[1] = t.host[h]; ng = 2; }
		for (int k = 0; k < ng; k++) {
			uint32_t e;
			if (!syn_earlier(t, p0, g[k], &e)) { o->h1[o->n1] = e; o->t1[o->n1] = ty; o->n1++; }
		}
	}
}

// ---- host table --------------------------------------------------------------------------------
struct HostTable {
	std::vector<uint64_t> R_off;
	std::vector<uint8_t> flags;
	std::vector<uint32_t> host, mrank;
	std::vector<int32_t> npart0;       // catalogue npart of the later halos (own + subhalo particles)
	SynthTable t;
	int64_t n_tuples0;
};

static uint64_t gcd_u64(uint64_t a, uint64_t b) { while (b) { uint64_t r = a % b; a = b; b = r; } return a; }

static int build_table(const metro_synth_spec *spec, int nptypes, HostTable &T, char *err, size_t errlen)
{
	const int64_t H = spec->n_halos;
	if (H < 1 || spec->n_tuples < H || spec->pid_bits < 8 || spec->pid_bits > 62 || spec->n_shards < 1 ||
	    spec->shard < 0 || spec->shard >= spec->n_shards || (spec->n_shards & (spec->n_shards - 1))) {
		snprintf(err, errlen, "bad synth spec");
		return METRO_ERR_INVALID;
	}
	std::vector<double> raw((size_t) H);
	T.flags.assign((size_t) H, 0); T.host.resize((size_t) H); T.mrank.assign((size_t) H, 0);
	const double cap = std::max(20.0, 1.0e-3 * (double) spec->n_tuples);
	double raw_tuples = 0;
	int64_t last_host = 0, n_merge = 0;
	for (int64_t h = 0; h < H; h++) {
		const double u = ((double) (syn_hash(spec->seed, 1, (uint64_t) h, 0) >> 11) + 1.0) * (1.0 / 9007199254740992.0);
		double s = floor(20.0 * pow(u, -1.0 / 0.9));
		if (s < 20.0) s = 20.0;
		if (s > cap) s = cap;
		raw[(size_t) h] = s;
		uint8_t f = 0;
		if (h > 0 && syn_hash(spec->seed, 2, (uint64_t) h, 0) % 100u < 15u) f |= SYN_F_SUB;
		const uint64_t q = syn_hash(spec->seed, 4, (uint64_t) h, 0) % 100u;
		if (q < 1) f |= SYN_F_ORPHAN;
		else if (q < 3) { f |= SYN_F_MERGER; T.mrank[(size_t) h] = (uint32_t) n_merge++; }
		T.flags[(size_t) h] = f;
		if (f & SYN_F_SUB) T.host[(size_t) h] = (uint32_t) last_host;
		else { T.host[(size_t) h] = (uint32_t) h; last_host = h; }
		raw_tuples += s * ((f & SYN_F_SUB) ? 2.0 : 1.0);
	}
	const double scale = (double) spec->n_tuples / raw_tuples;
	T.R_off.resize((size_t) H + 1);
	T.npart0.assign((size_t) H, 0);
	uint64_t R = 0;
	int64_t tuples = 0;
	for (int64_t h = 0; h < H; h++) {
		int64_t own = (int64_t) llround(raw[(size_t) h] * scale);
		if (own < 20) own = 20;
		T.R_off[(size_t) h] = R;
		R += (uint64_t) own;
		T.npart0[(size_t) h] += (int32_t) own;
		tuples += own;
		if (T.flags[(size_t) h] & SYN_F_SUB) { T.npart0[T.host[(size_t) h]] += (int32_t) own; tuples += own; }
	}
	T.R_off[(size_t) H] = R;
	if (2 * R > (1ull << spec->pid_bits)) {
		snprintf(err, errlen, "pid_bits %d too small: need 2 x %llu ranks", spec->pid_bits, (unsigned long long) R);
		return METRO_ERR_INVALID;
	}
	T.n_tuples0 = tuples;
	SynthTable &t = T.t;
	t.seed = spec->seed; t.H = H; t.H1 = H + n_merge; t.R_tot = R; t.pid_bits = spec->pid_bits; t.nptypes = nptypes;
	uint64_t a = (syn_hash(spec->seed, 5, 0, 0) % (uint64_t) t.H1) | 1ull;
	while (gcd_u64(a, (uint64_t) t.H1) != 1) a += 2;
	t.perm_a = a % (uint64_t) t.H1; if (t.perm_a == 0) t.perm_a = 1;
	t.perm_b = syn_hash(spec->seed, 6, 0, 0) % (uint64_t) t.H1;
	t.R_off = T.R_off.data(); t.flags = T.flags.data(); t.host = T.host.data(); t.mrank = T.mrank.data();
	return METRO_OK;
}

static void shard_range(const metro_synth_spec *spec, uint64_t *lo, uint64_t *hi)
{
	const uint64_t dom = 1ull << spec->pid_bits;
	const uint64_t per = dom / (uint64_t) spec->n_shards;
	*lo = per * (uint64_t) spec->shard;
	*hi = *lo + per;
}

// ---- device generator ----------------------------------------------------------------------------
// mode 0: count tuples; mode 1: emit (atomic append, warp aggregated) + earlier-halo histogram
__global__ void __launch_bounds__(256) synth_kernel(SynthTable t, u64 lo, u64 hi, int mode, u64 *counters, u64 *pid0, u32 *halo0, u8 *type0,
						    u64 *pid1, u32 *halo1, u8 *type1, i32 *npart1)
{
	const u64 v = lo + (u64) blockIdx.x * blockDim.x + threadIdx.x;
	SynOut o;
	o.n0 = o.n1 = 0;
	if (v < hi) syn_tuples_of_pid(t, v, &o);
	const int lane = threadIdx.x & 31;
	u32 c0 = o.n0, c1 = o.n1;
	u32 p0 = c0, p1 = c1;          // inclusive warp scans
#pragma unroll
	for (int off = 1; off < 32; off <<= 1) {
		u32 a = __shfl_up_sync(0xffffffffu, p0, off), b = __shfl_up_sync(0xffffffffu, p1, off);
		if (lane >= off) { p0 += a; p1 += b; }
	}
	const u32 w0 = __shfl_sync(0xffffffffu, p0, 31), w1 = __shfl_sync(0xffffffffu, p1, 31);
	u64 b0 = 0, b1 = 0;
	if (lane == 31) {
		if (w0) b0 = atomicAdd((unsigned long long *) &counters[0], (unsigned long long) w0);
		if (w1) b1 = atomicAdd((unsigned long long *) &counters[1], (unsigned long long) w1);
	}
	b0 = __shfl_sync(0xffffffffu, b0, 31); b1 = __shfl_sync(0xffffffffu, b1, 31);
	if (mode == 0) return;
	for (u32 k = 0; k < c0; k++) {
		const u64 d = b0 + p0 - c0 + k;
		pid0[d] = v; halo0[d] = o.h0[k];
		if (type0) type0[d] = o.t0[k];
	}
	for (u32 k = 0; k < c1; k++) {
		const u64 d = b1 + p1 - c1 + k;
		pid1[d] = v; halo1[d] = o.h1[k];
		if (type1) type1[d] = o.t1[k];
		atomicAdd(&npart1[o.h1[k]], 1);
	}
}

__global__ void synth_halo_kernel(i64 H, int T, u64 id_base, const i32 *np_tot, u64 *id, i32 *npart, i32 *orph, u8 *tok)
{
	const i64 i = (i64) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= H) return;
	id[i] = id_base + (u64) i + 1ull;
	// like IOSettings::ReadLineAHF (reference src/IOSettings.cpp:961-989): nPart[1] = npart, rest 0
	for (int t = 0; t < T; t++) npart[i * T + t] = (t == (T > 1 ? 1 : 0)) ? np_tot[i] : 0;
	orph[i] = 0; tok[i] = 0;
}

#define TRY(x) do { int _rc = (x); if (_rc) return _rc; } while (0)

extern "C" int metro_synth_pair_device(metro_ctx *ctx, const metro_params *prm, const metro_synth_spec *spec)
{
	if (!ctx || !prm || !spec) return METRO_ERR_INVALID;
	if (prm->nptypes < 1 || prm->nptypes > 8) { metro_set_error(ctx, "nptypes out of range"); return METRO_ERR_INVALID; }
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	int rc = partition_eager_join(ctx);
	if (rc) return rc;
	ctx->resident.cat0 = false; ctx->resident.cat1 = false; ctx->resident.a1 = false;
	HostTable HT;
	rc = build_table(spec, prm->nptypes, HT, ctx->err, sizeof ctx->err);
	if (rc) return rc;
	cudaStream_t st = ctx->stream;
	const int T = prm->nptypes;
	const i64 H = HT.t.H, H1 = HT.t.H1;
	DevBuf *S = ctx->scratch;
	TRY(dev_reserve(ctx, S[0], (size_t) (H + 1) * 8)); TRY(dev_reserve(ctx, S[1], (size_t) H)); TRY(dev_reserve(ctx, S[2], (size_t) H * 4));
	TRY(dev_reserve(ctx, S[3], (size_t) H * 4)); TRY(dev_reserve(ctx, S[4], (size_t) H * 4));
	CUDA_TRY(ctx, cudaMemcpyAsync(S[0].p, HT.R_off.data(), (size_t) (H + 1) * 8, cudaMemcpyHostToDevice, st));
	CUDA_TRY(ctx, cudaMemcpyAsync(S[1].p, HT.flags.data(), (size_t) H, cudaMemcpyHostToDevice, st));
	CUDA_TRY(ctx, cudaMemcpyAsync(S[2].p, HT.host.data(), (size_t) H * 4, cudaMemcpyHostToDevice, st));
	CUDA_TRY(ctx, cudaMemcpyAsync(S[3].p, HT.mrank.data(), (size_t) H * 4, cudaMemcpyHostToDevice, st));
	CUDA_TRY(ctx, cudaMemcpyAsync(S[4].p, HT.npart0.data(), (size_t) H * 4, cudaMemcpyHostToDevice, st));
	SynthTable dt = HT.t;
	dt.R_off = S[0].as<u64>(); dt.flags = S[1].as<u8>(); dt.host = S[2].as<u32>(); dt.mrank = S[3].as<u32>();
	u64 lo, hi;
	shard_range(spec, &lo, &hi);
	u64 *cnt = ctx->counters.as<u64>() + 24;
	const unsigned blocks = (unsigned) div_up<u64>(hi - lo, 256);
	CUDA_TRY(ctx, cudaMemsetAsync(cnt, 0, 2 * sizeof(u64), st));
	synth_kernel<<<blocks, 256, 0, st>>>(dt, lo, hi, 0, cnt, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
	CUDA_TRY(ctx, cudaGetLastError());
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters + 24, cnt, 2 * sizeof(u64), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	const i64 N0 = (i64) ctx->h_counters[24], N1 = (i64) ctx->h_counters[25];
	const bool typed = T > 2;
	DevCatalogue &c0 = ctx->cat[0], &c1 = ctx->cat[1];
	DevCatalogue *cs[2] = {&c0, &c1};
	const i64 Ns[2] = {N0, N1}, Hs[2] = {H, H1};
	for (int s = 0; s < 2; s++) {
		DevCatalogue &c = *cs[s];
		c.valid = false; c.b_pid = nullptr; c.b_halo = nullptr; c.b_type = nullptr;
		TRY(dev_reserve(ctx, c.pid, (size_t) Ns[s] * 8 + 8)); TRY(dev_reserve(ctx, c.halo, (size_t) Ns[s] * 4 + 8));
		if (typed) TRY(dev_reserve(ctx, c.type, (size_t) Ns[s] + 8)); else dev_free(c.type);
		TRY(dev_reserve(ctx, c.halo_id, (size_t) Hs[s] * 8 + 8)); TRY(dev_reserve(ctx, c.npart, (size_t) Hs[s] * T * 4 + 8));
		TRY(dev_reserve(ctx, c.npart_tot, (size_t) Hs[s] * 4 + 8)); TRY(dev_reserve(ctx, c.n_orphan_steps, (size_t) Hs[s] * 4 + 8));
		TRY(dev_reserve(ctx, c.is_token, (size_t) Hs[s] + 8));
		c.n_halos = Hs[s]; c.n_tuples = Ns[s];
	}
	CUDA_TRY(ctx, cudaMemcpyAsync(c0.npart_tot.p, S[4].p, (size_t) H * 4, cudaMemcpyDeviceToDevice, st));
	CUDA_TRY(ctx, cudaMemsetAsync(c1.npart_tot.p, 0, (size_t) H1 * 4, st));
	CUDA_TRY(ctx, cudaMemsetAsync(cnt, 0, 2 * sizeof(u64), st));
	synth_kernel<<<blocks, 256, 0, st>>>(dt, lo, hi, 1, cnt, c0.pid.as<u64>(), c0.halo.as<u32>(), typed ? c0.type.as<u8>() : nullptr,
					     c1.pid.as<u64>(), c1.halo.as<u32>(), typed ? c1.type.as<u8>() : nullptr, c1.npart_tot.as<i32>());
	CUDA_TRY(ctx, cudaGetLastError());
	for (int s = 0; s < 2; s++) {
		DevCatalogue &c = *cs[s];
		synth_halo_kernel<<<(unsigned) div_up<i64>(Hs[s], 256), 256, 0, st>>>(Hs[s], T, s == 0 ? 1000000000000ull : 0ull, c.npart_tot.as<i32>(),
										       c.halo_id.as<u64>(), c.npart.as<i32>(), c.n_orphan_steps.as<i32>(),
										       c.is_token.as<u8>());
		CUDA_TRY(ctx, cudaGetLastError());
	}
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	for (int s = 0; s < 2; s++) {
		TRY(catalogue_scan_tuples(ctx, *cs[s], T));
		cs[s]->gen = ctx->next_gen++;
		cs[s]->valid = true;
	}
	ctx->nptypes = T;
	ctx->res.valid = false;
	return METRO_OK;
}

// Replace the per-type particle counts of a resident catalogue (sharded runs: the catalogue npart
// of the earlier halos is the sum over all pid shards, which only the caller can form).
extern "C" int metro_catalogue_set_npart(metro_ctx *ctx, int slot, const int32_t *npart)
{
	if (!ctx || slot < 0 || slot > 1 || !npart) return METRO_ERR_INVALID;
	DevCatalogue &c = ctx->cat[slot];
	if (!c.valid) { metro_set_error(ctx, "catalogue %d is empty", slot); return METRO_ERR_STATE; }
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	const i64 H = c.n_halos;
	const int T = ctx->nptypes;
	std::vector<i32> tot((size_t) H);
	for (i64 h = 0; h < H; h++) { i32 s = 0; for (int t = 0; t < T; t++) s += npart[h * T + t]; tot[(size_t) h] = s; }
	CUDA_TRY(ctx, cudaMemcpyAsync(c.npart.p, npart, (size_t) H * T * 4, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(ctx, cudaMemcpyAsync(c.npart_tot.p, tot.data(), (size_t) H * 4, cudaMemcpyHostToDevice, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	return METRO_OK;
}

// ---- host generator (same tuples, pid-ascending order) -----------------------------------------
extern "C" int metro_synth_pair_host(const metro_params *prm, const metro_synth_spec *spec, int slot, int64_t *n_halos, int64_t *n_tuples,
				     uint64_t *halo_id, int32_t *npart, uint64_t *pid, uint32_t *halo, uint8_t *type)
{
	if (!prm || !spec || slot < 0 || slot > 1 || !n_halos || !n_tuples) return METRO_ERR_INVALID;
	HostTable HT;
	char err[256];
	int rc = build_table(spec, prm->nptypes, HT, err, sizeof err);
	if (rc) { metro_set_error(nullptr, "%s", err); return rc; }
	const int T = prm->nptypes;
	const SynthTable &t = HT.t;
	const int64_t H = slot == 0 ? t.H : t.H1;
	u64 lo, hi;
	shard_range(spec, &lo, &hi);
	const bool fill = pid != nullptr;
	std::vector<int32_t> cnt1;
	if (fill && slot == 1) cnt1.assign((size_t) H, 0);
	int64_t n = 0;
	for (u64 v = lo; v < hi; v++) {
		SynOut o;
		syn_tuples_of_pid(t, v, &o);
		const int c = slot == 0 ? o.n0 : o.n1;
		if (fill)
			for (int k = 0; k < c; k++) {
				pid[n + k] = v;
				halo[n + k] = slot == 0 ? o.h0[k] : o.h1[k];
				if (type) type[n + k] = slot == 0 ? o.t0[k] : o.t1[k];
				if (slot == 1) cnt1[o.h1[k]]++;
			}
		n += c;
	}
	*n_halos = H; *n_tuples = n;
	if (fill && halo_id && npart)
		for (int64_t h = 0; h < H; h++) {
			halo_id[h] = (slot == 0 ? 1000000000000ull : 0ull) + (u64) h + 1ull;
			const int32_t tot = slot == 0 ? HT.npart0[(size_t) h] : cnt1[(size_t) h];
			for (int k = 0; k < T; k++) npart[h * T + k] = (k == (T > 1 ? 1 : 0)) ? tot : 0;
		}
	return METRO_OK;
}
