// capi.cu -- the extern "C" boundary (include/metro_b200.h) and the per-pair pipeline driver.
#include <chrono>
#include <stdlib.h>

#include "context.cuh"

static char g_create_error[512] = "";

void metro_set_error(metro_ctx *ctx, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	if (ctx) vsnprintf(ctx->err, sizeof ctx->err, fmt, ap);
	else vsnprintf(g_create_error, sizeof g_create_error, fmt, ap);
	va_end(ap);
}

int dev_reserve(metro_ctx *ctx, DevBuf &b, size_t bytes)
{
	if (bytes <= b.cap && b.p) return METRO_OK;
	// 1/4 head room when an existing buffer grows, and from the start for everything below 1 GB: record and pair counts
	// wander by a few entries from run to run (which slice sees a pair first) and by a few percent from pair to pair of a
	// chain; without head room that costs a cudaFree + cudaMalloc (tens of ms, and a stall of every rank of a multi-GPU
	// run in the next collective) per buffer
	if (b.p || bytes < (1ull << 30)) bytes += bytes / 4 + 4096;
	static const bool dbg = getenv("METRO_DEBUG_ALLOC") != nullptr;
	const auto t0 = std::chrono::steady_clock::now();
	const size_t old = b.cap;
	if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
	if (bytes < 256) bytes = 256;
	CUDA_TRY(ctx, cudaMalloc(&b.p, bytes));
	b.cap = bytes;
	if (dbg) fprintf(stderr, "[metro] dev_reserve %zu -> %zu bytes, %.3f ms\n", old, bytes,
			 std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
	return METRO_OK;
}

// grow a buffer whose first `keep` bytes must survive (new allocation + device copy + free, only when it is too small)
int dev_reserve_keep(metro_ctx *ctx, DevBuf &b, size_t bytes, size_t keep)
{
	if (bytes <= b.cap && b.p) return METRO_OK;
	DevBuf n;
	int rc = dev_reserve(ctx, n, bytes + bytes / 8);
	if (rc) return rc;
	if (b.p && keep) CUDA_TRY(ctx, cudaMemcpyAsync(n.p, b.p, keep, cudaMemcpyDeviceToDevice, ctx->stream));
	if (b.p) { CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(b.p); }
	b = n;
	return METRO_OK;
}

void dev_free(DevBuf &b)
{
	if (b.p) cudaFree(b.p);
	b.p = nullptr; b.cap = 0;
}

extern "C" {

const char *metro_version(void) { return "metro-b200 0.1 (sm_100a)"; }

const char *metro_last_error(const metro_ctx *ctx) { return ctx ? ctx->err : g_create_error; }

int metro_ctx_create(int device, metro_ctx **out)
{
	if (!out) return METRO_ERR_INVALID;
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) {
		metro_set_error(nullptr, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
		return METRO_ERR_CUDA;
	}
	if (device < 0 || device >= ndev) { metro_set_error(nullptr, "device %d out of range (%d visible)", device, ndev); return METRO_ERR_INVALID; }
	cudaDeviceProp prop;
	if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { metro_set_error(nullptr, "%s", cudaGetErrorString(e)); return METRO_ERR_CUDA; }
	if (prop.major != 10) {
		metro_set_error(nullptr, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
		return METRO_ERR_CUDA;
	}
	if ((e = cudaSetDevice(device)) != cudaSuccess) { metro_set_error(nullptr, "%s", cudaGetErrorString(e)); return METRO_ERR_CUDA; }
	metro_ctx *ctx = new metro_ctx();
	ctx->device = device;
	memset(&ctx->timings, 0, sizeof ctx->timings);
	memset(&ctx->res.info, 0, sizeof ctx->res.info);
	bool ok = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) == cudaSuccess;
	ctx->stream = ctx->own_stream;
	ok = ok && cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
	ok = ok && cudaStreamCreateWithFlags(&ctx->eager_stream, cudaStreamNonBlocking) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&ctx->ev_eager, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&ctx->ev_eager0, cudaEventDisableTiming) == cudaSuccess;
	for (int i = 0; ok && i < 5; i++) ok = cudaEventCreateWithFlags(&ctx->ev_up[i], cudaEventDisableTiming) == cudaSuccess;
	for (int i = 0; ok && i < 16; i++) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
	for (int i = 0; ok && i < SORT_MAX_PASSES + 1; i++) ok = cudaEventCreate(&ctx->ev_pass[i]) == cudaSuccess;
	ok = ok && cudaMallocHost(&ctx->h_counters, 64 * sizeof(u64)) == cudaSuccess;
	ok = ok && dev_reserve(ctx, ctx->counters, 64 * sizeof(u64)) == METRO_OK;
	if (!ok) {
		metro_set_error(nullptr, "context setup failed: %s", cudaGetErrorString(cudaGetLastError()));
		delete ctx;
		return METRO_ERR_CUDA;
	}
	*out = ctx;
	return METRO_OK;
}

int metro_ctx_set_stream(metro_ctx *ctx, void *cuda_stream)
{
	if (!ctx) return METRO_ERR_INVALID;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	partition_eager_join(ctx);
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->stream = cuda_stream ? (cudaStream_t) cuda_stream : ctx->own_stream;
	return METRO_OK;
}

static void free_catalogue(DevCatalogue &c)
{
	DevBuf *b[] = {&c.halo_id, &c.npart, &c.npart_tot, &c.n_orphan_steps, &c.is_token, &c.pid, &c.halo, &c.type, &c.id_rank};
	for (DevBuf *x : b) dev_free(*x);
	c = DevCatalogue();
}

void metro_ctx_destroy(metro_ctx *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	if (ctx->eager_stream) cudaStreamSynchronize(ctx->eager_stream);
	cudaStreamSynchronize(ctx->stream);
	free_catalogue(ctx->cat[0]); free_catalogue(ctx->cat[1]);
	DevResult &r = ctx->res;
	for (int d = 0; d < 2; d++) { dev_free(r.raw_off[d]); dev_free(r.raw_prog[d]); dev_free(r.raw_ncommon[d]); dev_free(r.raw_merit[d]); }
	dev_free(r.tree_main); dev_free(r.tree_orphan); dev_free(r.tree_off); dev_free(r.clean_prog); dev_free(r.clean_ncommon); dev_free(r.token_halo);
	sort_workspace_free(ctx->sortws);
	comm_destroy(ctx);
	dev_free(ctx->keys_a); dev_free(ctx->keys_b); dev_free(ctx->table); dev_free(ctx->pair_cache); dev_free(ctx->scan_tmp); dev_free(ctx->counters);
	dev_free(ctx->part_counts);
	for (DevBuf &b : ctx->scratch) dev_free(b);
	if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
	for (int i = 0; i < 16; i++) cudaEventDestroy(ctx->ev[i]);
	for (int i = 0; i < SORT_MAX_PASSES + 1; i++) cudaEventDestroy(ctx->ev_pass[i]);
	for (int i = 0; i < 5; i++) if (ctx->ev_up[i]) cudaEventDestroy(ctx->ev_up[i]);
	if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
	if (ctx->eager_stream) cudaStreamDestroy(ctx->eager_stream);
	if (ctx->ev_eager) cudaEventDestroy(ctx->ev_eager);
	if (ctx->ev_eager0) cudaEventDestroy(ctx->ev_eager0);
	cudaStreamDestroy(ctx->own_stream);
	delete ctx;
}

static int check_params(metro_ctx *ctx, const metro_params *prm)
{
	if (!prm) { metro_set_error(ctx, "params is NULL"); return METRO_ERR_INVALID; }
	if (prm->nptypes < 1 || prm->nptypes > 8) { metro_set_error(ctx, "nptypes %d out of range 1..8", prm->nptypes); return METRO_ERR_INVALID; }
	if (prm->fac_orphan_steps == 0) {
		// the reference divides by facOrphanSteps (src/MergerTree.cpp:581): unset => SIGFPE there
		metro_set_error(ctx, "facOrphanSteps must be non-zero");
		return METRO_ERR_INVALID;
	}
	if (prm->flavour != METRO_FLAVOUR_GATHER && prm->flavour != METRO_FLAVOUR_ZOOM_DUP) { metro_set_error(ctx, "unknown flavour %d", prm->flavour); return METRO_ERR_INVALID; }
	return METRO_OK;
}

static int upload_halos(metro_ctx *ctx, DevCatalogue &c, const metro_params *prm, const metro_catalogue *cat)
{
	const i64 H = cat->n_halos;
	const int T = prm->nptypes;
	cudaStream_t st = ctx->stream;
	int rc;
	const size_t Hcap = (size_t) H + (size_t) H / 16 + 4096;     // head room for the token halos a shift appends
	if ((rc = dev_reserve(ctx, c.halo_id, Hcap * 8 + 8))) return rc;
	if ((rc = dev_reserve(ctx, c.npart, Hcap * T * 4 + 8))) return rc;
	if ((rc = dev_reserve(ctx, c.npart_tot, Hcap * 4 + 8))) return rc;
	if ((rc = dev_reserve(ctx, c.n_orphan_steps, Hcap * 4 + 8))) return rc;
	if ((rc = dev_reserve(ctx, c.is_token, Hcap + 8))) return rc;
	if (H > 0) {
		// Halo::nAllPart (src/Halo.cpp:63-70): plain int sum over the types
		i32 *tot = (i32 *) malloc((size_t) H * 4);
		if (!tot) { metro_set_error(ctx, "out of host memory"); return METRO_ERR_INVALID; }
		for (i64 h = 0; h < H; h++) {
			i32 s = 0;
			for (int t = 0; t < T; t++) s += cat->npart[h * T + t];
			tot[h] = s;
		}
		cudaError_t e = cudaMemcpyAsync(c.npart_tot.p, tot, (size_t) H * 4, cudaMemcpyHostToDevice, st);
		if (e == cudaSuccess) e = cudaStreamSynchronize(st);
		free(tot);
		CUDA_TRY(ctx, e);
		CUDA_TRY(ctx, cudaMemcpyAsync(c.halo_id.p, cat->halo_id, (size_t) H * 8, cudaMemcpyHostToDevice, st));
		CUDA_TRY(ctx, cudaMemcpyAsync(c.npart.p, cat->npart, (size_t) H * T * 4, cudaMemcpyHostToDevice, st));
		if (cat->n_orphan_steps) CUDA_TRY(ctx, cudaMemcpyAsync(c.n_orphan_steps.p, cat->n_orphan_steps, (size_t) H * 4, cudaMemcpyHostToDevice, st));
		else CUDA_TRY(ctx, cudaMemsetAsync(c.n_orphan_steps.p, 0, (size_t) H * 4, st));
		if (cat->is_token) CUDA_TRY(ctx, cudaMemcpyAsync(c.is_token.p, cat->is_token, (size_t) H, cudaMemcpyHostToDevice, st));
		else CUDA_TRY(ctx, cudaMemsetAsync(c.is_token.p, 0, (size_t) H, st));
	}
	c.n_halos = H;
	return METRO_OK;
}

static int check_catalogue(metro_ctx *ctx, int slot, const metro_params *prm, const metro_catalogue *cat, bool csr = false)
{
	int rc = check_params(ctx, prm);
	if (rc) return rc;
	if (slot < 0 || slot > 1 || !cat) { metro_set_error(ctx, "bad slot / NULL catalogue"); return METRO_ERR_INVALID; }
	if (cat->n_halos < 0 || cat->n_tuples < 0 || cat->n_halos >= (1ll << 31)) { metro_set_error(ctx, "bad catalogue sizes"); return METRO_ERR_INVALID; }
	if (cat->n_halos > 0 && (!cat->halo_id || !cat->npart)) { metro_set_error(ctx, "halo_id / npart is NULL"); return METRO_ERR_INVALID; }
	if (cat->n_tuples > 0 && (!cat->pid || (!csr && !cat->halo))) { metro_set_error(ctx, "pid / halo is NULL"); return METRO_ERR_INVALID; }
	if (cat->n_tuples > 0 && cat->n_halos == 0) { metro_set_error(ctx, "tuples without halos"); return METRO_ERR_INVALID; }
	return METRO_OK;
}

// per-tuple halo index from CSR offsets: every thread owns 16 consecutive tuples, finds the halo of the
// first one by binary search (offsets are L2 resident) and walks forward from there
__global__ void __launch_bounds__(256) expand_halo_kernel(const i64 *__restrict__ off, i64 H, i64 N, u32 *__restrict__ halo)
{
	const i64 base = ((i64) blockIdx.x * 256 + threadIdx.x) * 16;
	if (base >= N) return;
	i64 lo = 0, hi = H;                                                  // last h with off[h] <= base
	while (hi - lo > 1) {
		const i64 mid = (lo + hi) >> 1;
		if (__ldg(off + mid) <= base) lo = mid; else hi = mid;
	}
	i64 h = lo, next = __ldg(off + h + 1);
	u32 v[16];
#pragma unroll
	for (int q = 0; q < 16; q++) {
		const i64 i = base + q;
		while (i >= next && h + 1 < H) { h++; next = __ldg(off + h + 1); }
		v[q] = (u32) h;
	}
	if (base + 16 <= N) {
		uint4 *dst = reinterpret_cast<uint4 *>(halo + base);
#pragma unroll
		for (int q = 0; q < 4; q++) dst[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
	} else {
		for (int q = 0; q < 16 && base + q < N; q++) halo[base + q] = v[q];
	}
}

// One chunk of an upload, already on the device: widen 32-bit ID offsets to the u64 IDs the catalogue holds (src32 != NULL), fold
// min / max ID and the range checks of halo index and type into the same read.  out: [0] = max pid, [1] = #invalid, [2] = min pid.
__global__ void __launch_bounds__(256) upload_chunk_kernel(const u32 *__restrict__ src32, u64 pid_base, u64 *__restrict__ pid, const u32 *__restrict__ halo,
							    const u8 *__restrict__ type, u64 n, u32 n_halos, u32 nptypes, u64 *out)
{
	u64 mx = 0, mn = ~0ull;
	u32 bad = 0;
	for (u64 i = (u64) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64) gridDim.x * blockDim.x) {
		u64 p;
		if (src32) { p = pid_base + (u64) src32[i]; pid[i] = p; }
		else p = pid[i];
		mx = p > mx ? p : mx;
		mn = p < mn ? p : mn;
		if (halo && halo[i] >= n_halos) bad++;
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

static const i64 UP_CHUNK = getenv("METRO_UP_CHUNK") ? std::max<i64>(1 << 16, atoll(getenv("METRO_UP_CHUNK"))) : (16ll << 20);   // tuples per upload chunk (128 MB of u64 IDs: 2-3 ms over PCIe 5)

// Tuples host -> HBM.  The copy runs in chunks on the context's copy stream; every chunk is checked (and, for 32-bit ID offsets,
// widened) on the compute stream while the next one is crossing the bus, so the upload costs the bus time and nothing else.
static int upload_tuples(metro_ctx *ctx, int slot, const metro_params *prm, const metro_catalogue *cat, const int64_t *halo_offset,
			 const uint32_t *pid32 = nullptr, u64 pid_base = 0)
{
	int rc;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	// (an upload into slot 1 does not wait for the eager partition of catalogue 0: that overlap is its purpose)
	if (slot == 0 && (rc = partition_eager_join(ctx))) return rc;
	DevCatalogue &c = ctx->cat[slot];
	c.valid = false; c.b_pid = nullptr; c.b_halo = nullptr; c.b_type = nullptr;
	if ((rc = upload_halos(ctx, c, prm, cat))) return rc;
	const i64 N = cat->n_tuples, H = cat->n_halos;
	cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
	// head room: after metro_shift this catalogue's buffers become catalogue 0 with the token tuples appended in place
	const size_t Ncap = (size_t) N + (size_t) N / 16 + 65536;
	if ((rc = dev_reserve(ctx, c.pid, Ncap * 8 + 8))) return rc;
	if ((rc = dev_reserve(ctx, c.halo, Ncap * 4 + 64))) return rc;
	if (cat->type) { if ((rc = dev_reserve(ctx, c.type, Ncap + 8))) return rc; }
	else dev_free(c.type);
	u64 *cnt = ctx->counters.as<u64>() + 16;      // [16] max pid, [17] #invalid, [18] min pid
	CUDA_TRY(ctx, cudaMemsetAsync(cnt, 0, 2 * sizeof(u64), st));
	CUDA_TRY(ctx, cudaMemsetAsync(cnt + 2, 0xFF, sizeof(u64), st));
	// catalogue 1 of a pair whose catalogue 0 is (being) partitioned: pass A of every chunk right behind its check
	ctx->e1_on = false;
	if (slot == 1 && (rc = partition_eager1_begin(ctx, prm, N, H))) return rc;
	if (N > 0) {
		if (halo_offset) {
			if (halo_offset[0] != 0 || halo_offset[H] != N) { metro_set_error(ctx, "halo_offset must run from 0 to n_tuples"); return METRO_ERR_INVALID; }
			if ((rc = dev_reserve(ctx, ctx->scratch[79], (size_t) (H + 1) * 8))) return rc;
			CUDA_TRY(ctx, cudaMemcpyAsync(ctx->scratch[79].p, halo_offset, (size_t) (H + 1) * 8, cudaMemcpyHostToDevice, st));
			expand_halo_kernel<<<(unsigned) div_up<i64>(N, 256 * 16), 256, 0, st>>>(ctx->scratch[79].as<i64>(), H, N, c.halo.as<u32>());
			CUDA_TRY(ctx, cudaGetLastError());
		}
		if (pid32 && (rc = dev_reserve(ctx, ctx->scratch[78], (size_t) 2 * UP_CHUNK * 4))) return rc;
		// the copy stream starts after everything queued on the compute stream so far (the buffers may still be in use)
		CUDA_TRY(ctx, cudaEventRecord(ctx->ev_up[0], st));
		CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ctx->ev_up[0], 0));
		int k = 0;
		for (i64 lo = 0; lo < N; lo += UP_CHUNK, k ^= 1) {
			const i64 n = std::min<i64>(UP_CHUNK, N - lo);
			u32 *stage32 = pid32 ? ctx->scratch[78].as<u32>() + (size_t) k * UP_CHUNK : nullptr;
			if (pid32) {
				if (lo >= 2 * UP_CHUNK) CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ctx->ev_up[3 + k], 0));      // staging half consumed
				CUDA_TRY(ctx, cudaMemcpyAsync(stage32, pid32 + lo, (size_t) n * 4, cudaMemcpyHostToDevice, cs));
			} else CUDA_TRY(ctx, cudaMemcpyAsync(c.pid.as<u64>() + lo, cat->pid + lo, (size_t) n * 8, cudaMemcpyHostToDevice, cs));
			if (!halo_offset) CUDA_TRY(ctx, cudaMemcpyAsync(c.halo.as<u32>() + lo, cat->halo + lo, (size_t) n * 4, cudaMemcpyHostToDevice, cs));
			if (cat->type) CUDA_TRY(ctx, cudaMemcpyAsync(c.type.as<u8>() + lo, cat->type + lo, (size_t) n, cudaMemcpyHostToDevice, cs));
			CUDA_TRY(ctx, cudaEventRecord(ctx->ev_up[1 + k], cs));
			CUDA_TRY(ctx, cudaStreamWaitEvent(st, ctx->ev_up[1 + k], 0));
			upload_chunk_kernel<<<(unsigned) std::min<i64>(div_up<i64>(n, 256 * 4), (i64) METRO_NUM_SMS_B200 * 8), 256, 0, st>>>(
				stage32, pid_base, c.pid.as<u64>() + lo, halo_offset ? nullptr : c.halo.as<u32>() + lo, cat->type ? c.type.as<u8>() + lo : nullptr,
				(u64) n, (u32) H, (u32) prm->nptypes, cnt);
			CUDA_TRY(ctx, cudaGetLastError());
			if (pid32) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_up[3 + k], st));
			if ((rc = partition_eager1_chunk(ctx, c.pid.as<u64>() + lo, c.halo.as<u32>() + lo, cat->type ? c.type.as<u8>() + lo : nullptr, n))) return rc;
		}
	}
	c.n_tuples = N;
	c.gen = ctx->next_gen++;
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters + 16, cnt, 3 * sizeof(u64), cudaMemcpyDeviceToHost, st));
	if ((rc = partition_eager1_flush(ctx))) return rc;
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	c.max_pid = ctx->h_counters[16];
	c.min_pid = N > 0 ? ctx->h_counters[18] : 0;
	partition_eager1_end(ctx, c, ctx->h_counters[17] == 0);
	if (ctx->h_counters[17] != 0) {
		metro_set_error(ctx, "catalogue has %llu tuples with a halo index >= n_halos or a type >= nptypes", (unsigned long long) ctx->h_counters[17]);
		return METRO_ERR_INVALID;
	}
	c.valid = true;
	ctx->nptypes = prm->nptypes;
	ctx->res.valid = false;
	// catalogue 0 of a pair: partition it now, under the upload of catalogue 1
	if (slot == 0) return stage_partition_eager(ctx, prm);
	return METRO_OK;
}

int metro_catalogue_upload_csr(metro_ctx *ctx, int slot, const metro_params *prm, const metro_catalogue *cat, const int64_t *halo_offset)
{
	if (!ctx) return METRO_ERR_INVALID;
	int rc = check_catalogue(ctx, slot, prm, cat, true);
	if (rc) return rc;
	if (!halo_offset) { metro_set_error(ctx, "halo_offset is NULL"); return METRO_ERR_INVALID; }
	for (i64 h = 0; h < cat->n_halos; h++)
		if (halo_offset[h + 1] < halo_offset[h]) { metro_set_error(ctx, "halo_offset decreases at halo %lld", (long long) h); return METRO_ERR_INVALID; }
	return upload_tuples(ctx, slot, prm, cat, halo_offset);
}

int metro_catalogue_upload_csr32(metro_ctx *ctx, int slot, const metro_params *prm, const metro_catalogue *cat, const uint32_t *pid32, uint64_t pid_base,
				 const int64_t *halo_offset)
{
	if (!ctx) return METRO_ERR_INVALID;
	metro_catalogue tmp;
	if (cat) { tmp = *cat; tmp.pid = reinterpret_cast<const uint64_t *>(pid32); }          // (only tested for NULL by the checks)
	int rc = check_catalogue(ctx, slot, prm, cat ? &tmp : nullptr, true);
	if (rc) return rc;
	if (!halo_offset) { metro_set_error(ctx, "halo_offset is NULL"); return METRO_ERR_INVALID; }
	if (pid_base > ~0ull - 0xFFFFFFFFull) { metro_set_error(ctx, "pid_base + 2^32 overflows"); return METRO_ERR_INVALID; }
	for (i64 h = 0; h < cat->n_halos; h++)
		if (halo_offset[h + 1] < halo_offset[h]) { metro_set_error(ctx, "halo_offset decreases at halo %lld", (long long) h); return METRO_ERR_INVALID; }
	return upload_tuples(ctx, slot, prm, cat, halo_offset, pid32, pid_base);
}

int metro_catalogue_upload(metro_ctx *ctx, int slot, const metro_params *prm, const metro_catalogue *cat)
{
	if (!ctx) return METRO_ERR_INVALID;
	int rc = check_catalogue(ctx, slot, prm, cat);
	if (rc) return rc;
	return upload_tuples(ctx, slot, prm, cat, nullptr);
}

int metro_catalogue_adopt_device(metro_ctx *ctx, int slot, const metro_params *prm, const metro_catalogue *cat)
{
	if (!ctx) return METRO_ERR_INVALID;
	int rc = check_catalogue(ctx, slot, prm, cat);
	if (rc) return rc;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	if ((rc = partition_eager_join(ctx))) return rc;
	DevCatalogue &c = ctx->cat[slot];
	c.valid = false;
	if ((rc = upload_halos(ctx, c, prm, cat))) return rc;
	dev_free(c.pid); dev_free(c.halo); dev_free(c.type);
	c.b_pid = cat->pid; c.b_halo = cat->halo; c.b_type = cat->type;
	c.n_tuples = cat->n_tuples;
	if (c.n_tuples == 0) { c.b_pid = nullptr; c.b_halo = nullptr; c.b_type = nullptr; }
	c.gen = ctx->next_gen++;
	if ((rc = catalogue_scan_tuples(ctx, c, prm->nptypes))) return rc;
	c.valid = true;
	ctx->nptypes = prm->nptypes;
	ctx->res.valid = false;
	return METRO_OK;
}

int metro_match_pair(metro_ctx *ctx, const metro_params *prm, metro_result_info *info)
{
	if (!ctx) return METRO_ERR_INVALID;
	int rc = check_params(ctx, prm);
	if (rc) return rc;
	if (!ctx->cat[0].valid || !ctx->cat[1].valid) { metro_set_error(ctx, "both catalogues must be uploaded before metro_match_pair"); return METRO_ERR_STATE; }
	if (prm->nptypes != ctx->nptypes) { metro_set_error(ctx, "nptypes differs from the uploaded catalogues"); return METRO_ERR_INVALID; }
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	if ((rc = partition_eager_join(ctx))) return rc;
	cudaStream_t st = ctx->stream;
	DevCatalogue &c0 = ctx->cat[0], &c1 = ctx->cat[1];
	const i64 N = c0.n_tuples + c1.n_tuples;
	const i64 Hmax = std::max(c0.n_halos, c1.n_halos);
	// key layout: pid | snap | halo | type must fit one u64 (see common.cuh)
	// particle IDs enter the key relative to the smallest one present (a pid-range shard of a
	// multi-GPU run spans fewer bits than the global IDs => fewer sort passes)
	u64 pid_base = ~0ull;
	if (c0.n_tuples > 0) pid_base = std::min(pid_base, c0.min_pid);
	if (c1.n_tuples > 0) pid_base = std::min(pid_base, c1.min_pid);
	if (pid_base == ~0ull) pid_base = 0;
	int pid_bits = std::max(1, ceil_log2_u64(std::max(c0.max_pid, c1.max_pid) - pid_base));
	{	// prefer a base that does not move from pair to pair of a chain (0, else a multiple of 2^pid_bits) when it costs no key bit
		const u64 maxp = std::max(c0.max_pid, c1.max_pid);
		const u64 aligned = pid_bits < 64 ? (pid_base >> pid_bits) << pid_bits : 0;
		if (std::max(1, ceil_log2_u64(maxp)) <= pid_bits) pid_base = 0;
		else if (std::max(1, ceil_log2_u64(maxp - aligned)) <= pid_bits) pid_base = aligned;
	}
	ctx->pid_base = pid_base;
	int halo_bits = std::max(1, ceil_log2_u64(Hmax > 0 ? (u64) (Hmax - 1) : 0));
	int type_bits = prm->nptypes > 1 ? ceil_log2_u64((u64) prm->nptypes - 1) : 0;
	if (2 * halo_bits + 2 * type_bits > 63 || halo_bits + type_bits > 31) {
		metro_set_error(ctx, "pair record does not fit: 2 x (halo %d + type %d bits)", halo_bits, type_bits);
		return METRO_ERR_UNSUPPORTED;
	}
	// pid | snap | halo | type in one 64-bit word if it fits; else the wide-key front end sorts (pid, payload) pairs
	const bool wide = pid_bits + 1 + halo_bits + type_bits > 64;
	const KeyLayout L = make_key_layout(wide ? 64 - 1 - halo_bits - type_bits : pid_bits, halo_bits, type_bits);
	const SortPlan plan = make_sort_plan(L.pid_shift, L.pid_shift + pid_bits);
	DevResult &res = ctx->res;
	res.valid = false;
	memset(&res.info, 0, sizeof res.info);
	res.info.n_halos[0] = c0.n_halos; res.info.n_halos[1] = c1.n_halos;
	res.info.n_tuples[0] = c0.n_tuples; res.info.n_tuples[1] = c1.n_tuples;
	res.info.pid_bits = pid_bits; res.info.halo_bits = halo_bits; res.info.type_bits = type_bits; res.info.sort_passes = plan.n_pass;
	ctx->launches = 0;
	memset(&ctx->timings, 0, sizeof ctx->timings);
	// Two front ends produce the same pair records:
	//  * partition join (large pairs): mix(pid) MSD partition x2 + per-bucket shared-memory hash join
	//  * LSD radix sort by pid + run-length join (any size; also the fallback when a bucket overflows)
	i64 R = 0, hits = 0;
	int sort_launches = 0;
	KeyLayout Lp = L;                                       // layout the partition path runs with
	auto front_end = [&]() -> int {
		static const bool force_lsd = getenv("METRO_FORCE_LSD") && atoi(getenv("METRO_FORCE_LSD"));
		static const bool no_resident = getenv("METRO_NO_RESIDENT") && atoi(getenv("METRO_NO_RESIDENT"));
		ctx->timings.path = 0;
		ctx->timings.resident = 0;
		bool part = false;
		if (!force_lsd && !wide) {
			// chain step: the buckets of catalogue 0 are resident if its tuples are the ones they were built from and
			// this pair fits the layout and plan they were built with
			PartResident &Rs = ctx->resident;
			const i64 nmax = std::max(c0.n_tuples, c1.n_tuples);
			bool reuse0 = !no_resident && Rs.cat0 && Rs.gen == c0.gen && c0.n_tuples > 0 && Rs.L.type_bits == type_bits && halo_bits <= Rs.L.halo_bits &&
				      pid_base >= Rs.pid_base && nmax <= Rs.nmax_cap &&
				      ceil_log2_u64(std::max(c0.max_pid, c1.max_pid) - Rs.pid_base) <= Rs.L.pid_bits;
			static const bool dbg = getenv("METRO_DEBUG") != nullptr;
			if (dbg && Rs.cat0 && !reuse0)
				fprintf(stderr, "[metro] resident buckets not reused: gen %d type %d halo %d (%d<=%d) base %d (%llu>=%llu) nmax %d (%lld<=%lld) pidbits %d\n",
					Rs.gen == c0.gen, Rs.L.type_bits == type_bits, halo_bits <= Rs.L.halo_bits, halo_bits, Rs.L.halo_bits, pid_base >= Rs.pid_base,
					(unsigned long long) pid_base, (unsigned long long) Rs.pid_base, nmax <= Rs.nmax_cap, (long long) nmax, (long long) Rs.nmax_cap,
					ceil_log2_u64(std::max(c0.max_pid, c1.max_pid) - Rs.pid_base) <= Rs.L.pid_bits);
			// attempt 0: resident plan; 1: fresh plan; 2, 3: fresh plan with 2x / 4x finer buckets after a per-bucket overflow
			// (table or hit staging: more memberships per particle than the default bucket size provides for)
			int finer = ctx->part_finer;
			for (int attempt = reuse0 ? 0 : 1; attempt < 4 && !part; attempt++) {
				PartPlan PP;
				bool fell_back = false;
				u32 mask = 0;
				if (attempt == 0) {
					PP = Rs.P; Lp = Rs.L; ctx->pid_base = Rs.pid_base;
				} else {
					// a fresh plan; one spare halo bit so that the token halos of the next shifts still fit the layout
					const int hb = (halo_bits + 1 + 2 * type_bits <= 31 && pid_bits + 2 + halo_bits + type_bits <= 64) ? halo_bits + 1 : halo_bits;
					Lp = make_key_layout(pid_bits, hb, type_bits);
					ctx->pid_base = pid_base;
					i64 ncap = 0;
					if (!partition_join_plan(ctx, Lp, N, &PP, &ncap, finer)) break;
					Rs.nmax_cap = ncap;
				}
				CUDA_TRY(ctx, cudaEventRecord(ctx->ev[0], st));
				if ((rc = stage_partition_join(ctx, Lp, PP, halo_bits, attempt == 0, &R, &hits, &fell_back, &mask, finer > 0))) return rc;
				if (!fell_back) { part = true; ctx->timings.path = 1; ctx->timings.resident = attempt == 0 ? (ctx->last_a1 ? 3 : 1) : 0; sort_launches = 4; if (attempt >= 1) ctx->part_finer = finer; }
				else if (attempt >= 1) {
					if ((mask & ~6u) != 0 || finer >= 2) break;      // region / record overflows are not cured by finer buckets: LSD path
					finer++;
				}
			}
			if (!part) ctx->pid_base = pid_base;
		}
		// the wide-key and LSD front ends overwrite keys_a / keys_b: whatever buckets were resident there are gone
		if (wide || !part) { ctx->resident.cat0 = false; ctx->resident.cat1 = false; ctx->resident.a1 = false; }
		if (wide) {
			CUDA_TRY(ctx, cudaEventRecord(ctx->ev[0], st));
			if ((rc = stage_wide_join(ctx, L, pid_bits, &R, &hits, &sort_launches))) return rc;
			ctx->timings.path = 2;
		} else if (!part) {
			if ((rc = dev_reserve(ctx, ctx->keys_a, (size_t) N * 8 + 16))) return rc;
			if ((rc = dev_reserve(ctx, ctx->keys_b, (size_t) N * 8 + 16))) return rc;
			if ((rc = sort_workspace_reserve(ctx, ctx->sortws, N))) return rc;
			CUDA_TRY(ctx, cudaEventRecord(ctx->ev[0], st));
			if ((rc = stage_pack_hist(ctx, L, plan))) return rc;
			CUDA_TRY(ctx, cudaEventRecord(ctx->ev[1], st));
			u64 *sorted = nullptr;
			if ((rc = sort_scatter_passes(ctx, st, ctx->sortws, ctx->keys_a.as<u64>(), ctx->keys_b.as<u64>(), nullptr, nullptr, N, plan, &sorted, nullptr,
						      ctx->ev_pass, &sort_launches))) return rc;
			ctx->launches += sort_launches;
			sort_launches = N > 0 ? plan.n_pass : 0;
			CUDA_TRY(ctx, cudaEventRecord(ctx->ev[2], st));
			if ((rc = stage_join_count(ctx, L, sorted, N, &R, &hits))) return rc;
		}
		return METRO_OK;
	};
	// Every rank must take the same exit: a rank that failed here (allocation, a table overflow on its shard) would otherwise
	// strand its peers in the collectives of the exchange.  The ranks agree on the worst status first.
	rc = front_end();
	{
		const int agreed = comm_agree_status(ctx, rc);
		if (agreed != METRO_OK) {
			if (rc == METRO_OK) metro_set_error(ctx, "another rank failed in the join stage (status %d)", agreed);
			return rc != METRO_OK ? rc : agreed;
		}
	}
	res.info.n_hits = hits;
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[3], st));
	// multi-GPU: partial pair counts -> owner(haloA) / owner(haloB) (no-op on one GPU)
	const u64 *rak, *rbk; const u32 *rac, *rbc;
	i64 RA = 0, RB = 0, a_lo = 0, a_hi = 0;
	ctx->exchange_bytes = 0;
	ctx->exchange_timed = false;
	if ((rc = comm_exchange_records(ctx, halo_bits, type_bits, ctx->scratch[0].as<u64>(), ctx->scratch[1].as<u32>(), R, &rak, &rac, &RA, &rbk,
					&rbc, &RB))) return rc;
	comm_owned_range(ctx, c0.n_halos, &a_lo, &a_hi);
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[7], st));
	if ((rc = stage_select(ctx, prm, L, rak, rac, RA, rbk, rbc, RB, a_lo, a_hi))) return rc;
	res.info.exchange_bytes = ctx->exchange_bytes;
	res.info.owned_lo = a_lo; res.info.owned_hi = a_hi;
	if (ctx->world > 1) {
		// global statistics for the caller
		u64 *d = ctx->counters.as<u64>() + 56;
		u64 h[3] = {(u64) res.info.n_hits, (u64) res.info.n_untracked, (u64) res.info.n_table_hits};
		CUDA_TRY(ctx, cudaMemcpyAsync(d, h, sizeof h, cudaMemcpyHostToDevice, st));
		if ((rc = comm_allreduce_sum_u64(ctx, d, 3))) return rc;
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_counters + 56, d, sizeof h, cudaMemcpyDeviceToHost, st));
		CUDA_TRY(ctx, cudaStreamSynchronize(st));
		res.info.n_hits = (i64) ctx->h_counters[56]; res.info.n_untracked = (i64) ctx->h_counters[57]; res.info.n_table_hits = (i64) ctx->h_counters[58];
	}
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[4], st));
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	metro_timings &tm = ctx->timings;
	cudaEventElapsedTime(&tm.total, ctx->ev[0], ctx->ev[4]);
	cudaEventElapsedTime(&tm.pack_hist, ctx->ev[0], ctx->ev[1]);
	cudaEventElapsedTime(&tm.sort, ctx->ev[1], ctx->ev[2]);
	cudaEventElapsedTime(&tm.join, ctx->ev[2], ctx->ev[3]);
	cudaEventElapsedTime(&tm.exchange, ctx->ev[3], ctx->ev[7]);
	cudaEventElapsedTime(&tm.select, ctx->ev[7], ctx->ev[4]);
	if (ctx->exchange_timed) cudaEventElapsedTime(&tm.exchange_sendrecv, ctx->ev[11], ctx->ev[12]);
	if (tm.path == 1) {
		tm.sort_pass_ms[0] = tm.pack_hist; tm.sort_pass_ms[1] = tm.sort;
		cudaEventElapsedTime(&tm.sort_pass_ms[2], ctx->ev[2], ctx->ev[10]);
		cudaEventElapsedTime(&tm.sort_pass_ms[3], ctx->ev[10], ctx->ev[3]);
	}
	else if (tm.path == 0 && N > 0) for (int p = 0; p < plan.n_pass; p++) cudaEventElapsedTime(&tm.sort_pass_ms[p], ctx->ev_pass[p], ctx->ev_pass[p + 1]);
	tm.sort_launches = sort_launches;
	tm.total_launches = ctx->launches;
	res.valid = true;
	if (info) *info = res.info;
	return METRO_OK;
}

int metro_get_timings(const metro_ctx *ctx, metro_timings *out)
{
	if (!ctx || !out) return METRO_ERR_INVALID;
	*out = ctx->timings;
	return METRO_OK;
}

static int d2h(metro_ctx *ctx, void *dst, const DevBuf &src, size_t bytes)
{
	if (!dst || bytes == 0) return METRO_OK;
	CUDA_TRY(ctx, cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
	return METRO_OK;
}

int metro_result_fetch(metro_ctx *ctx, const metro_result_arrays *dst)
{
	if (!ctx || !dst) return METRO_ERR_INVALID;
	if (!ctx->res.valid) { metro_set_error(ctx, "no result: call metro_match_pair first"); return METRO_ERR_STATE; }
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	DevResult &r = ctx->res;
	const int T = ctx->nptypes;
	int rc = 0;
	for (int d = 0; d < 2 && !rc; d++) {
		const size_t M = (size_t) r.info.raw_count[d];
		rc = d2h(ctx, dst->raw_off[d], r.raw_off[d], (size_t) (r.info.n_halos[d] + 1) * 8);
		if (!rc) rc = d2h(ctx, dst->raw_prog[d], r.raw_prog[d], M * 4);
		if (!rc) rc = d2h(ctx, dst->raw_ncommon[d], r.raw_ncommon[d], M * T * 4);
		if (!rc) rc = d2h(ctx, dst->raw_merit[d], r.raw_merit[d], M * 4);
	}
	if (!rc) rc = d2h(ctx, dst->tree_main, r.tree_main, (size_t) r.info.n_trees * 4);
	if (!rc) rc = d2h(ctx, dst->tree_orphan, r.tree_orphan, (size_t) r.info.n_trees);
	if (!rc) rc = d2h(ctx, dst->tree_off, r.tree_off, (size_t) (r.info.n_trees + 1) * 8);
	if (!rc) rc = d2h(ctx, dst->clean_prog, r.clean_prog, (size_t) r.info.clean_count * 4);
	if (!rc) rc = d2h(ctx, dst->clean_ncommon, r.clean_ncommon, (size_t) r.info.clean_count * T * 4);
	if (!rc) rc = d2h(ctx, dst->token_halo, r.token_halo, (size_t) r.info.n_tokens * 4);
	if (rc) return rc;
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	return METRO_OK;
}

int metro_shift(metro_ctx *ctx, const metro_params *prm)
{
	if (!ctx) return METRO_ERR_INVALID;
	int rc = check_params(ctx, prm);
	if (rc) return rc;
	if (!ctx->res.valid) { metro_set_error(ctx, "metro_shift needs the result of metro_match_pair (token halos)"); return METRO_ERR_STATE; }
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	if ((rc = partition_eager_join(ctx))) return rc;
	return stage_shift(ctx, prm);
}

int metro_catalogue_info(metro_ctx *ctx, int slot, int64_t *n_halos, int64_t *n_tuples)
{
	if (!ctx || slot < 0 || slot > 1) return METRO_ERR_INVALID;
	if (!ctx->cat[slot].valid) { metro_set_error(ctx, "catalogue %d is empty", slot); return METRO_ERR_STATE; }
	if (n_halos) *n_halos = ctx->cat[slot].n_halos;
	if (n_tuples) *n_tuples = ctx->cat[slot].n_tuples;
	return METRO_OK;
}

int metro_catalogue_fetch_halos(metro_ctx *ctx, int slot, uint64_t *halo_id, int32_t *npart, int32_t *n_orphan_steps, uint8_t *is_token)
{
	if (!ctx || slot < 0 || slot > 1) return METRO_ERR_INVALID;
	DevCatalogue &c = ctx->cat[slot];
	if (!c.valid) { metro_set_error(ctx, "catalogue %d is empty", slot); return METRO_ERR_STATE; }
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	const size_t H = (size_t) c.n_halos;
	int rc = d2h(ctx, halo_id, c.halo_id, H * 8);
	if (!rc) rc = d2h(ctx, npart, c.npart, H * ctx->nptypes * 4);
	if (!rc) rc = d2h(ctx, n_orphan_steps, c.n_orphan_steps, H * 4);
	if (!rc) rc = d2h(ctx, is_token, c.is_token, H);
	if (rc) return rc;
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	return METRO_OK;
}

int metro_catalogue_fetch_tuples(metro_ctx *ctx, int slot, uint64_t *pid, uint32_t *halo, uint8_t *type)
{
	if (!ctx || slot < 0 || slot > 1) return METRO_ERR_INVALID;
	DevCatalogue &c = ctx->cat[slot];
	if (!c.valid) { metro_set_error(ctx, "catalogue %d is empty", slot); return METRO_ERR_STATE; }
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	const size_t N = (size_t) c.n_tuples;
	if (N > 0) {
		if (pid) CUDA_TRY(ctx, cudaMemcpyAsync(pid, c.d_pid(), N * 8, cudaMemcpyDeviceToHost, ctx->stream));
		if (halo) CUDA_TRY(ctx, cudaMemcpyAsync(halo, c.d_halo(), N * 4, cudaMemcpyDeviceToHost, ctx->stream));
		if (type) {
			if (c.d_type()) CUDA_TRY(ctx, cudaMemcpyAsync(type, c.d_type(), N, cudaMemcpyDeviceToHost, ctx->stream));
			else memset(type, 1, N);
		}
	}
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	return METRO_OK;
}

int metro_sort_u64(metro_ctx *ctx, uint64_t *keys_in, uint64_t *keys_out, uint32_t *vals_in, uint32_t *vals_out, int64_t n,
		   int begin_bit, int end_bit)
{
	if (!ctx || n < 0 || begin_bit < 0 || end_bit > 64 || begin_bit > end_bit) return METRO_ERR_INVALID;
	if (n > 0 && (!keys_in || !keys_out)) return METRO_ERR_INVALID;
	if ((vals_in == nullptr) != (vals_out == nullptr)) return METRO_ERR_INVALID;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	cudaStream_t st = ctx->stream;
	int rc = sort_workspace_reserve(ctx, ctx->sortws, n);
	if (rc) return rc;
	SortPlan plan = make_sort_plan(begin_bit, end_bit);
	if ((rc = sort_clear_hist(ctx, st, ctx->sortws))) return rc;
	if ((rc = sort_histogram(ctx, st, ctx->sortws, keys_in, n, plan))) return rc;
	u64 *rk = nullptr; u32 *rv = nullptr;
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[5], st));
	if ((rc = sort_scatter_passes(ctx, st, ctx->sortws, keys_in, keys_out, vals_in, vals_out, n, plan, &rk, &rv, ctx->ev_pass, nullptr))) return rc;
	CUDA_TRY(ctx, cudaEventRecord(ctx->ev[6], st));
	if (n > 0 && rk != keys_out) {
		CUDA_TRY(ctx, cudaMemcpyAsync(keys_out, rk, (size_t) n * 8, cudaMemcpyDeviceToDevice, st));
		if (vals_in) CUDA_TRY(ctx, cudaMemcpyAsync(vals_out, rv, (size_t) n * 4, cudaMemcpyDeviceToDevice, st));
	}
	CUDA_TRY(ctx, cudaStreamSynchronize(st));
	memset(&ctx->timings, 0, sizeof ctx->timings);
	cudaEventElapsedTime(&ctx->timings.sort, ctx->ev[5], ctx->ev[6]);
	if (n > 0) for (int p = 0; p < plan.n_pass; p++) cudaEventElapsedTime(&ctx->timings.sort_pass_ms[p], ctx->ev_pass[p], ctx->ev_pass[p + 1]);
	ctx->timings.sort_launches = n > 0 ? plan.n_pass : 0;
	return METRO_OK;
}

int metro_merit_bits(metro_ctx *ctx, int64_t n, const int32_t *n_comm, const int32_t *n_ph0, const int32_t *n_ph1, const int32_t *i_m,
		     uint32_t *bits_out)
{
	if (!ctx || n < 0 || (n > 0 && (!n_comm || !n_ph0 || !n_ph1 || !i_m || !bits_out))) return METRO_ERR_INVALID;
	if (n == 0) return METRO_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	DevBuf *S = ctx->scratch;
	int rc = 0;
	for (int k = 0; k < 5 && !rc; k++) rc = dev_reserve(ctx, S[k], (size_t) n * 4);
	if (rc) return rc;
	const int32_t *src[4] = {n_comm, n_ph0, n_ph1, i_m};
	for (int k = 0; k < 4; k++) CUDA_TRY(ctx, cudaMemcpyAsync(S[k].p, src[k], (size_t) n * 4, cudaMemcpyHostToDevice, ctx->stream));
	if ((rc = merit_bits_device(ctx, n, S[0].as<i32>(), S[1].as<i32>(), S[2].as<i32>(), S[3].as<i32>(), S[4].as<u32>()))) return rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(bits_out, S[4].p, (size_t) n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
	return METRO_OK;
}

}  // extern "C"
