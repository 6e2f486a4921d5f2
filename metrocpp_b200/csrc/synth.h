// synth.h -- counter-based synthetic AHF-like snapshot pairs.
//
// Everything that decides a tuple is integer hashing of (seed, particle rank / pid, halo), so the
// CUDA generator and the host generator emit the same multiset of tuples.  The per-halo table
// (sizes from the power law s = max(20, 20 u^(-1/0.9)), subhalo / merger / orphan flags) is built
// once on the host in double precision and shared by both.
//
// Model.  Later catalogue ("0"): halo h owns ranks [R_h, R_{h+1}); pid(rank) = Feistel(rank).
// 15 % of the halos are subhalos: their particles are ALSO listed under the preceding host
// (AHF double membership).  Earlier catalogue ("1"), decided per later tuple (pid, g):
//   94 %  the particle is in the progenitor of g   (88 % exactly e(g), 6 % moved to e(g+d), |d|<=3)
//    6 %  unbound: the progenitor lists a fresh pid instead (rank + R_tot, never in catalogue 0)
//   2 % of the halos are 2-way mergers: 40 % of their particles come from a second progenitor;
//   1 % of the halos are orphans: the progenitor slot holds only fresh pids (no shared particle).
// Earlier halo indices are relabelled by an affine permutation e(x) = (a x + b) mod H1.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SYN_HD __host__ __device__ __forceinline__
#else
#define SYN_HD static inline
#endif

#define SYN_F_SUB 1u
#define SYN_F_ORPHAN 2u
#define SYN_F_MERGER 4u

struct SynthTable {            // pointers valid on the side (host or device) that uses them
	uint64_t seed;
	int64_t H;             // halos of the later catalogue
	int64_t H1;            // halos of the earlier catalogue = H + number of mergers
	uint64_t R_tot;        // particle ranks of the later catalogue
	int pid_bits;          // even
	uint64_t perm_a, perm_b;
	const uint64_t *R_off; // [H+1] first rank of each halo
	const uint8_t *flags;  // [H]
	const uint32_t *host;  // [H] host halo of a subhalo (else self)
	const uint32_t *mrank; // [H] running index among mergers (valid where SYN_F_MERGER)
	int nptypes;
};

SYN_HD uint64_t syn_mix(uint64_t x)
{	// splitmix64 finaliser
	x += 0x9E3779B97F4A7C15ull;
	x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
	x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
	return x ^ (x >> 31);
}
SYN_HD uint64_t syn_hash(uint64_t seed, uint64_t stream, uint64_t a, uint64_t b)
{
	return syn_mix(syn_mix(syn_mix(seed ^ (stream * 0xD6E8FEB86659FD93ull)) ^ a) ^ (b * 0xC2B2AE3D27D4EB4Full));
}

// balanced 4-round Feistel network over an even number of bits
SYN_HD uint64_t syn_feistel_even(uint64_t x, uint64_t seed, int bits)
{
	const int k = bits / 2;
	const uint64_t m = (1ull << k) - 1ull;
	uint64_t L = x >> k, R = x & m;
	for (int r = 0; r < 4; r++) {
		const uint64_t f = syn_mix(R ^ syn_mix(seed + 0x1000ull * (uint64_t) (r + 1))) & m;
		const uint64_t t = L ^ f;
		L = R; R = t;
	}
	return (L << k) | R;
}
SYN_HD uint64_t syn_feistel_even_inv(uint64_t y, uint64_t seed, int bits)
{
	const int k = bits / 2;
	const uint64_t m = (1ull << k) - 1ull;
	uint64_t L = y >> k, R = y & m;
	for (int r = 3; r >= 0; r--) {
		const uint64_t t = R;       // t = L_prev ^ f(R_prev), L = R_prev
		const uint64_t Rp = L;
		const uint64_t f = syn_mix(Rp ^ syn_mix(seed + 0x1000ull * (uint64_t) (r + 1))) & m;
		L = t ^ f; R = Rp;
	}
	return (L << k) | R;
}
// permutation of [0, 2^bits): an odd width walks the cycle of the (bits + 1)-bit network until it is back inside the domain
// (two steps on average; the restriction of a permutation to a subset by cycle walking is a permutation of the subset)
SYN_HD uint64_t syn_feistel(uint64_t x, uint64_t seed, int bits)
{
	if ((bits & 1) == 0) return syn_feistel_even(x, seed, bits);
	do { x = syn_feistel_even(x, seed, bits + 1); } while (x >> bits);
	return x;
}
SYN_HD uint64_t syn_feistel_inv(uint64_t y, uint64_t seed, int bits)
{
	if ((bits & 1) == 0) return syn_feistel_even_inv(y, seed, bits);
	do { y = syn_feistel_even_inv(y, seed, bits + 1); } while (y >> bits);
	return y;
}

SYN_HD int64_t syn_find_halo(const SynthTable &t, uint64_t rank)
{	// largest h with R_off[h] <= rank
	int64_t lo = 0, hi = t.H;
	while (hi - lo > 1) {
		const int64_t mid = (lo + hi) >> 1;
		if (t.R_off[mid] <= rank) lo = mid; else hi = mid;
	}
	return lo;
}

SYN_HD uint32_t syn_perm(const SynthTable &t, uint64_t x)
{
	return (uint32_t) ((t.perm_a * x + t.perm_b) % (uint64_t) t.H1);
}

SYN_HD uint8_t syn_type_later(const SynthTable &t, uint64_t pid)
{
	if (t.nptypes <= 2) return 1;
	const uint32_t u = (uint32_t) (syn_hash(t.seed, 7, pid, 0) % 100u);
	if (u < 25) return 0;                               // gas
	if (u < 90 || t.nptypes < 5) return 1;              // dark matter
	return 4;                                           // star
}
SYN_HD uint8_t syn_type_earlier(const SynthTable &t, uint64_t pid, uint8_t later)
{	// a fifth of the stars were still gas one snapshot earlier
	if (later == 4 && (syn_hash(t.seed, 8, pid, 0) % 5u) == 0) return 0;
	return later;
}

// Earlier-catalogue tuple derived from the later tuple (rank, pid, g).  Returns the earlier halo
// and whether the earlier halo lists the original pid (false => it lists the fresh pid instead).
SYN_HD bool syn_earlier(const SynthTable &t, uint64_t pid, int64_t g, uint32_t *halo1)
{
	const uint8_t fl = t.flags[g];
	const uint64_t u = syn_hash(t.seed, 3, pid, (uint64_t) g);
	const uint32_t dec = (uint32_t) (u % 10000u);
	int64_t src = g;
	bool same_pid = true;
	if (fl & SYN_F_ORPHAN) same_pid = false;
	else if (dec < 600) same_pid = false;
	else if (dec < 1200) {
		int64_t d = (int64_t) ((u >> 20) % 6u) - 3;      // -3..2
		if (d >= 0) d += 1;                              // -3..-1, 1..3
		src = g + d;
		if (src < 0) src = 0;
		if (src >= t.H) src = t.H - 1;
	}
	if (src == g && (fl & SYN_F_MERGER) && ((u >> 40) % 100u) < 40u) *halo1 = syn_perm(t, (uint64_t) t.H + t.mrank[g]);
	else *halo1 = syn_perm(t, (uint64_t) src);
	return same_pid;
}
