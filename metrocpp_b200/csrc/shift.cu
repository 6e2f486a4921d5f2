// shift.cu -- catalogue maintenance on the device: tuple validation / max pid, and the
// snapshot shift with token halos.
//
// stage_shift replaces ShiftHalosPartsGrids (reference src/utils.cpp:245-403) together with the
// orphan hand-over of CleanTrees (src/MergerTree.cpp:589-603) / SyncOrphanHalos
// (src/Communication.cpp:477-537): catalogue 1 becomes catalogue 0; every token halo is appended
// after the real halos, keeps the ID and nPart of the orphan, nOrphanSteps+1, isToken, and gets
// the orphan's tuples (twice each in the ZOOM flavour, which is what doubles a token's weight
// at every orphan step in that build).
#include "context.cuh"

#define TPB 256
#define GRID(n) ((unsigned) div_up<i64>((n) > 0 ? (n) : 1, TPB))
#define TRY(x) do { int _rc = (x); if (_rc) return _rc; } while (0)
#define LAUNCHED() do { CUDA_TRY(ctx, cudaGetLastError()); ctx->launches++; } while (0)

// max pid (for the key width) and range checks of halo index / type, one read of the tuples
__global__ void tuple_scan_kernel(const u64 *__restrict__ pid, const u32 *__restrict__ halo, const u8 *__restrict__ type,
				  u64 n, u32 n_halos, u32 nptypes, u64 *out /* [0]=max pid, [1]=#invalid, [2]=min pid */)
{
	u64 mx = 0, mn = ~0ull;
	u32 bad = 0;
	for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
		const u64 p = pid[i];
		mx = p > mx ? p : mx;
		mn = p < mn ? p : mn;
		if (halo[i] >= n_halos) bad++;
		if (type && type[i] >= nptypes) bad++;
	}
#pragma unroll
	for (int off = 16; off > 0; off >>= 1) {
		const u64 o = __shfl_xor_sync(0xffffffffu, mx, off);
		mx = o > mx ? o : mx;
		const u64 o2 = __shfl_xor_sync(0xffffffffu, mn, off);
		mn = o2 < mn ? o2 : mn;
		bad += __shfl_xor_sync(0xffffffffu, bad, off);
	}
	if ((threadIdx.x & 31) == 0) {
		atomicMax((unsigned long long *) &out[0], (unsigned long long) mx);
		atomicMin((unsigned long long *) &out[2], (unsigned long long) mn);
		if (bad) atomicAdd((unsigned long long *) &out[1], (unsigned long long) bad);
	}
}

int catalogue_scan_tuples(metro_ctx *ctx, DevCatalogue &c, int nptypes)
{
	cudaStream_t st = ctx->stream;
	u64 *cnt = ctx->counters.as<u64>() + 16;      // [16],[17],[18]; [18..] is re-used by stage_shift afterwards
	CUDA_TRY(ctx, cudaMemsetAsync(cnt, 0, 2 * sizeof(u64), st));
	CUDA_TRY(ctx, cudaMemsetAsync(cnt + 2, 0xFF, sizeof(u64), st));
	if (c.n_tuples > 0) {
		int blocks = (int) std::min<i64>(div_up<i64>(c.n_tuples, TPB), (i64) METRO_NUM_SMS_B200 * 8);
		tuple_scan_kernel<<<blocks, TPB, 0, st>>>(c.d_pid(), c.d_halo(), c.d_type(), (u64) c.n_tuples, (u32) c.n_halos, (u32) nptypes, cnt);
		LAUNCHED();
	}
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters + 16, cnt, 3 * sizeof(u64), cudaMemcpyDeviceToHost, st));
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	c.max_pid = ctx->h_counters[16];
	c.min_pid = c.n_tuples > 0 ? ctx->h_counters[18] : 0;
	if (ctx->h_counters[17] != 0) {
		metro_set_error(ctx, "catalogue has %llu tuples with a halo index >= n_halos or a type >= nptypes",
				(unsigned long long) ctx->h_counters[17]);
		return METRO_ERR_INVALID;
	}
	return METRO_OK;
}

__global__ void fill_i32_kernel(i32 *p, i64 n, i32 v) { i64 i = (i64) blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

__global__ void token_slot_kernel(const i32 *token_halo, i64 nk, i32 *slot)
{
	i64 i = (i64) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < nk) slot[token_halo[i]] = (i32) i;
}

__global__ void token_tuple_flag_kernel(const u32 *halo, i64 n, const i32 *slot, u32 *flag)
{
	i64 i = (i64) blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) flag[i] = slot[halo[i]] >= 0 ? 1u : 0u;
}

__global__ void token_tuple_emit_kernel(const u64 *pid, const u32 *halo, const u8 *type, i64 n, const i32 *slot, const u32 *flag,
					const u32 *scan, int rep, u32 halo_base, u64 *opid, u32 *ohalo, u8 *otype)
{
	i64 i = (i64) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n || !flag[i]) return;
	const i64 d = (i64) scan[i] * rep;
	for (int r = 0; r < rep; r++) {
		opid[d + r] = pid[i];
		ohalo[d + r] = halo_base + (u32) slot[halo[i]];
		if (otype) otype[d + r] = type ? type[i] : (u8) 1;
	}
}

__global__ void fill_u8_kernel(u8 *p, i64 n, u8 v) { i64 i = (i64) blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

__global__ void token_halo_emit_kernel(const i32 *token_halo, i64 nk, int T, const u64 *id0, const i32 *np0, const i32 *npt0,
				       const i32 *orph0, u64 *id, i32 *np, i32 *npt, i32 *orph, u8 *tok)
{
	i64 i = (i64) blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= nk) return;
	const i32 a = token_halo[i];
	id[i] = id0[a];
	npt[i] = npt0[a];
	orph[i] = orph0[a] + 1;
	tok[i] = 1;
	for (int t = 0; t < T; t++) np[i * T + t] = np0[(i64) a * T + t];
}

int stage_shift(metro_ctx *ctx, const metro_params *prm)
{
	cudaStream_t st = ctx->stream;
	const int T = prm->nptypes;
	DevCatalogue &c0 = ctx->cat[0], &c1 = ctx->cat[1];
	DevResult &res = ctx->res;
	const i64 NK = res.info.n_tokens, H0 = c0.n_halos, H1 = c1.n_halos, N0 = c0.n_tuples, N1 = c1.n_tuples;
	const int rep = prm->flavour == METRO_FLAVOUR_ZOOM_DUP ? 2 : 1;
	DevBuf *S = ctx->scratch;
	u64 extra = 0;
	if (NK > 0 && N0 > 0) {
		TRY(dev_reserve(ctx, S[0], (size_t) (H0 + 1) * 4));
		TRY(dev_reserve(ctx, S[1], (size_t) N0 * 4));
		TRY(dev_reserve(ctx, S[2], (size_t) N0 * 4));
		fill_i32_kernel<<<GRID(H0), TPB, 0, st>>>(S[0].as<i32>(), H0, -1); LAUNCHED();
		token_slot_kernel<<<GRID(NK), TPB, 0, st>>>(res.token_halo.as<i32>(), NK, S[0].as<i32>()); LAUNCHED();
		token_tuple_flag_kernel<<<GRID(N0), TPB, 0, st>>>(c0.d_halo(), N0, S[0].as<i32>(), S[1].as<u32>()); LAUNCHED();
		u64 *cnt = ctx->counters.as<u64>() + 18;
		TRY(scan_exclusive_u32(ctx, st, S[1].as<u32>(), S[2].as<u32>(), N0, cnt));
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters + 18, cnt, sizeof(u64), cudaMemcpyDeviceToHost, st));
		CUDA_TRY(ctx, cudaStreamSynchronize(st));
		extra = ctx->h_counters[18];
	}
	const i64 NN = N1 + (i64) extra * rep, HH = H1 + NK;
	DevCatalogue n0;
	const bool move = c1.b_pid == nullptr;
	if (move) {
		// Catalogue 1 owns its buffers: they BECOME catalogue 0 - its halos and tuples stay where they are, the token halos and
		// token tuples are appended behind them (uploads leave head room for exactly this), nothing is copied, allocated or freed.
		// A chain step of the reference re-inserts every particle of snapshot 1 into a new map here (src/utils.cpp:272-278).
		TRY(dev_reserve_keep(ctx, c1.halo_id, (size_t) HH * 8 + 8, (size_t) H1 * 8)); TRY(dev_reserve_keep(ctx, c1.npart, (size_t) HH * T * 4 + 8, (size_t) H1 * T * 4));
		TRY(dev_reserve_keep(ctx, c1.npart_tot, (size_t) HH * 4 + 8, (size_t) H1 * 4));
		TRY(dev_reserve_keep(ctx, c1.n_orphan_steps, (size_t) HH * 4 + 8, (size_t) H1 * 4));
		TRY(dev_reserve_keep(ctx, c1.is_token, (size_t) HH + 8, (size_t) H1));
		TRY(dev_reserve_keep(ctx, c1.pid, (size_t) NN * 8 + 8, (size_t) N1 * 8)); TRY(dev_reserve_keep(ctx, c1.halo, (size_t) NN * 4 + 64, (size_t) N1 * 4));
		if (c1.type.p) TRY(dev_reserve_keep(ctx, c1.type, (size_t) NN + 8, (size_t) N1));
		else if (c0.d_type() && extra > 0) {
			// catalogue 1 came without a type column (all 1), the token tuples carry types: materialise the column
			TRY(dev_reserve(ctx, c1.type, (size_t) NN + 8));
			if (N1 > 0) { fill_u8_kernel<<<GRID(N1), TPB, 0, st>>>(c1.type.as<u8>(), N1, 1); LAUNCHED(); }
		}
		n0 = c1;                                                 // the buffers move with the struct
		n0.id_rank_gen = 0; n0.id_rank_n = -1;                  // (the rank buffer moves along and is refilled: the token halos changed the order)
		n0.n_halos = HH; n0.n_tuples = NN;
	} else {
		// catalogue 1 is borrowed device memory (metro_catalogue_adopt_device): the new catalogue 0 is built in fresh buffers
		n0.n_halos = HH; n0.n_tuples = NN;
		TRY(dev_reserve(ctx, n0.halo_id, (size_t) HH * 8 + 8)); TRY(dev_reserve(ctx, n0.npart, (size_t) HH * T * 4 + 8));
		TRY(dev_reserve(ctx, n0.npart_tot, (size_t) HH * 4 + 8)); TRY(dev_reserve(ctx, n0.n_orphan_steps, (size_t) HH * 4 + 8));
		TRY(dev_reserve(ctx, n0.is_token, (size_t) HH + 8));
		TRY(dev_reserve(ctx, n0.pid, (size_t) NN * 8 + 8)); TRY(dev_reserve(ctx, n0.halo, (size_t) NN * 4 + 64));
		TRY(dev_reserve(ctx, n0.type, (size_t) NN + 8));
		if (H1 > 0) {
			CUDA_TRY(ctx, cudaMemcpyAsync(n0.halo_id.p, c1.halo_id.p, (size_t) H1 * 8, cudaMemcpyDeviceToDevice, st));
			CUDA_TRY(ctx, cudaMemcpyAsync(n0.npart.p, c1.npart.p, (size_t) H1 * T * 4, cudaMemcpyDeviceToDevice, st));
			CUDA_TRY(ctx, cudaMemcpyAsync(n0.npart_tot.p, c1.npart_tot.p, (size_t) H1 * 4, cudaMemcpyDeviceToDevice, st));
			CUDA_TRY(ctx, cudaMemcpyAsync(n0.n_orphan_steps.p, c1.n_orphan_steps.p, (size_t) H1 * 4, cudaMemcpyDeviceToDevice, st));
			CUDA_TRY(ctx, cudaMemcpyAsync(n0.is_token.p, c1.is_token.p, (size_t) H1, cudaMemcpyDeviceToDevice, st));
		}
		if (N1 > 0) {
			CUDA_TRY(ctx, cudaMemcpyAsync(n0.pid.p, c1.d_pid(), (size_t) N1 * 8, cudaMemcpyDeviceToDevice, st));
			CUDA_TRY(ctx, cudaMemcpyAsync(n0.halo.p, c1.d_halo(), (size_t) N1 * 4, cudaMemcpyDeviceToDevice, st));
			if (c1.d_type()) CUDA_TRY(ctx, cudaMemcpyAsync(n0.type.p, c1.d_type(), (size_t) N1, cudaMemcpyDeviceToDevice, st));
			else { fill_u8_kernel<<<GRID(N1), TPB, 0, st>>>(n0.type.as<u8>(), N1, 1); LAUNCHED(); }
		}
	}
	n0.b_pid = nullptr; n0.b_halo = nullptr; n0.b_type = nullptr;
	if (NK > 0) {
		token_halo_emit_kernel<<<GRID(NK), TPB, 0, st>>>(res.token_halo.as<i32>(), NK, T, c0.halo_id.as<u64>(), c0.npart.as<i32>(),
								  c0.npart_tot.as<i32>(), c0.n_orphan_steps.as<i32>(), n0.halo_id.as<u64>() + H1,
								  n0.npart.as<i32>() + H1 * T, n0.npart_tot.as<i32>() + H1,
								  n0.n_orphan_steps.as<i32>() + H1, n0.is_token.as<u8>() + H1);
		LAUNCHED();
		if (extra > 0) {
			token_tuple_emit_kernel<<<GRID(N0), TPB, 0, st>>>(c0.d_pid(), c0.d_halo(), c0.d_type(), N0, S[0].as<i32>(), S[1].as<u32>(),
									   S[2].as<u32>(), rep, (u32) H1, n0.pid.as<u64>() + N1, n0.halo.as<u32>() + N1,
									   n0.type.p ? n0.type.as<u8>() + N1 : nullptr);
			LAUNCHED();
		}
	}
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	n0.max_pid = std::max(c1.max_pid, extra > 0 ? c0.max_pid : (u64) 0);
	n0.min_pid = c1.n_tuples > 0 ? (extra > 0 ? std::min(c1.min_pid, c0.min_pid) : c1.min_pid) : (extra > 0 ? c0.min_pid : 0);
	n0.valid = true;
	n0.gen = ctx->next_gen++;
	// resident key buckets of the old catalogue 1 now belong to the new catalogue 0; its tuples [0, N1) are the old
	// catalogue-1 tuples in the same order, the token tuples behind them are not in the buckets yet
	PartResident &R = ctx->resident;
	if (R.cat1 && R.gen == c1.gen) { R.cat0 = true; R.gen = n0.gen; R.half0 = R.half1; R.half1 = 1 - R.half0; R.tail0 = N1; }
	else R.cat0 = false;
	R.cat1 = false; R.a1 = false;
	if (move) {
		// the buffers of the old catalogue 0 stay with slot 1 for the next upload (no free, no allocation in a chain step)
		DevCatalogue keep;
		keep.halo_id = c0.halo_id; keep.npart = c0.npart; keep.npart_tot = c0.npart_tot; keep.n_orphan_steps = c0.n_orphan_steps; keep.is_token = c0.is_token;
		keep.pid = c0.pid; keep.halo = c0.halo; keep.type = c0.type; keep.id_rank = c0.id_rank;
		ctx->cat[0] = n0;
		ctx->cat[1] = keep;
	} else {
		// release the old catalogues (owned buffers only)
		DevBuf *old0[] = {&c0.halo_id, &c0.npart, &c0.npart_tot, &c0.n_orphan_steps, &c0.is_token, &c0.pid, &c0.halo, &c0.type, &c0.id_rank};
		for (DevBuf *b : old0) dev_free(*b);
		DevBuf *old1[] = {&c1.halo_id, &c1.npart, &c1.npart_tot, &c1.n_orphan_steps, &c1.is_token, &c1.pid, &c1.halo, &c1.type, &c1.id_rank};
		for (DevBuf *b : old1) dev_free(*b);
		ctx->cat[0] = n0;
		ctx->cat[1] = DevCatalogue();
	}
	res.valid = false;
	return METRO_OK;
}
