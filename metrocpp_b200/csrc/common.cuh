// common.cuh -- shared device/host helpers for libmetro_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

typedef uint64_t u64;
typedef uint32_t u32;
typedef int64_t i64;
typedef int32_t i32;
typedef uint8_t u8;

#define METRO_NUM_SMS_B200 148

// Every CUDA call inside the library goes through this: records the message in the
// context's error buffer and returns METRO_ERR_CUDA to the C-ABI caller (never exits).
#define CUDA_TRY(ctx, expr)                                                                      \
	do {                                                                                     \
		cudaError_t _e = (expr);                                                         \
		if (_e != cudaSuccess) {                                                         \
			metro_set_error((ctx), "%s:%d %s -> %s", __FILE__, __LINE__, #expr,      \
					cudaGetErrorString(_e));                                 \
			return METRO_ERR_CUDA;                                                   \
		}                                                                                \
	} while (0)

struct metro_ctx;
void metro_set_error(metro_ctx *ctx, const char *fmt, ...);

static inline int ceil_log2_u64(u64 x)
{	// number of bits needed to represent values 0..x
	int b = 0;
	while (b < 64 && (x >> b) != 0) b++;
	return b;
}

template <typename T>
static inline T div_up(T a, T b) { return (a + b - 1) / b; }

// ---- packed join key ------------------------------------------------------------------
// One u64 per tuple:  [ pid : pid_bits ][ snap : 1 ][ halo : halo_bits ][ type : type_bits ]
// (most significant first).  The LSD sort runs over the pid bit range only; it is stable and the
// catalogue-0 tuples are packed in front of the catalogue-1 tuples, so inside a run of equal
// pid all snap=0 members precede all snap=1 members.
struct KeyLayout {
	int pid_bits, halo_bits, type_bits;
	int pid_shift;      // = 1 + halo_bits + type_bits
	int snap_shift;     // = halo_bits + type_bits
	u32 halo_mask, type_mask;
};

static inline KeyLayout make_key_layout(int pid_bits, int halo_bits, int type_bits)
{
	KeyLayout L;
	L.pid_bits = pid_bits; L.halo_bits = halo_bits; L.type_bits = type_bits;
	L.snap_shift = halo_bits + type_bits;
	L.pid_shift = L.snap_shift + 1;
	L.halo_mask = halo_bits >= 32 ? 0xFFFFFFFFu : ((1u << halo_bits) - 1u);
	L.type_mask = (1u << type_bits) - 1u;
	return L;
}

__device__ __forceinline__ u64 key_pack(const KeyLayout &L, u64 pid, u32 snap, u32 halo, u32 type)
{
	return (pid << L.pid_shift) | ((u64) snap << L.snap_shift) | ((u64) halo << L.type_bits) | (u64) type;
}
__device__ __forceinline__ u64 key_pid(const KeyLayout &L, u64 k) { return k >> L.pid_shift; }
__device__ __forceinline__ u32 key_snap(const KeyLayout &L, u64 k) { return (u32) (k >> L.snap_shift) & 1u; }
__device__ __forceinline__ u32 key_halo(const KeyLayout &L, u64 k) { return (u32) (k >> L.type_bits) & L.halo_mask; }
__device__ __forceinline__ u32 key_type(const KeyLayout &L, u64 k) { return (u32) k & L.type_mask; }

// streaming loads/stores: keys are touched once per pass, keep them out of L1
__device__ __forceinline__ u64 ld_stream_u64(const u64 *p)
{
	u64 v;
	asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
	return v;
}
__device__ __forceinline__ void st_stream_u64(u64 *p, u64 v)
{
	asm volatile("st.global.L1::no_allocate.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u32 lanemask_lt()
{
	u32 m;
	asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}
