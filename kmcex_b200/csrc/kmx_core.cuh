// kmx_core.cuh -- arithmetic shared by every kernel and by the host-side known-answer entry
// points: 2-bit k-mer codec, MurmurHash64A over the ASCII expansion, exact 64-bit modulo by
// a runtime constant, bit addressing of the reference's byte arrays.
//
// Reference behaviour reproduced here (file:line relative to the reference root):
//   tools.hpp:9        HashSeeds[128]
//   tools.hpp:16-50    murmur_hash64 (MurmurHash64A, little-endian 8-byte blocks + tail)
//   tools.hpp:63-76    kmers2uint64   tools.hpp:130-139 get_complementation
//   tools.hpp:160-167  get_min_kmer   kmodel.hpp:576-588 set_bit / check_bit (MSB-first bytes)
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define KMX_HD __host__ __device__ __forceinline__
#define KMX_D __device__ __forceinline__
#else
#define KMX_HD inline
#endif

namespace kmx {

constexpr uint64_t kMurM = 0xc6a4a7935bd1e995ULL;
constexpr int kBucketLog = 18;                    // kmodel.hpp:276  bucket_size = 1 << 18
constexpr uint32_t kBucket = 1u << kBucketLog;
constexpr int kMaxArrays = 8;                     // n_bits supported by this build
constexpr int kMaxHash = 12;                      // n_hash supported by this build (status words hold 14-bit masks)
constexpr int kMaxBf = 3;                         // kmodel.hpp:50   bf_num is 1 or 3

// ---------------------------------------------------------------------------------------
// exact h % d for a runtime-constant d >= 2:  with m = floor(2^64 / d) + 1 (so m*d > 2^64 and
// m <= 2^64/d + 1), q' = mulhi(h, m) satisfies h/d < h*m/2^64 < h/d + 1, i.e. q' is q or q + 1;
// h - q'*d is then the remainder or the remainder minus d (negative as a signed word, d < 2^63):
// one sign-masked add finishes it.  (A probe position costs one of these on top of its hash.)
// ---------------------------------------------------------------------------------------
struct FastMod {
	uint64_t d;
	uint64_t magic;
};

inline FastMod make_fastmod(uint64_t d) {
	FastMod f;
	f.d = d;
	f.magic = 0;
	if (d >= 2) f.magic = (~0ULL) / d + ((~0ULL) % d == d - 1 ? 1 : 0) + 1;   // floor(2^64 / d) + 1
	return f;
}

KMX_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
	return __umul64hi(a, b);
#else
	return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

KMX_HD uint64_t fastmod(uint64_t h, const FastMod& f) {
	const uint64_t q = mulhi64(h, f.magic);
	const uint64_t r = h - q * f.d;
	return r + (f.d & (uint64_t)((int64_t)r >> 63));
}

// ---------------------------------------------------------------------------------------
// 2-bit codec.  `v` = packed k-mer, first base most significant.  `r` = the same bases with
// base j in bits [2j, 2j+1] (first base LEAST significant): this is the order the ASCII bytes
// of the string have in memory, and ~r is the reverse complement in packed form.
// ---------------------------------------------------------------------------------------
KMX_HD uint64_t mask2(int n_bases) { return n_bases >= 32 ? ~0ULL : ((1ULL << (2 * n_bases)) - 1); }

KMX_HD uint64_t reverse_bases(uint64_t v, int k) {
#ifdef __CUDA_ARCH__
	uint64_t x = __brevll(v);
	x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
#else
	uint64_t x = v;
	x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
	x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
	x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
	x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
	x = (x >> 32) | (x << 32);
#endif
	return x >> (64 - 2 * k);
}

// canonical form (tools.hpp:160-167): min(v, revcomp(v)); also returns r for the canonical one
KMX_HD uint64_t canonical(uint64_t v, int k, uint64_t* r_out) {
	uint64_t r = reverse_bases(v, k);
	uint64_t rc = (~r) & mask2(k);          // packed reverse complement
	if (v <= rc) {
		*r_out = r;
		return v;
	}
	*r_out = (~v) & mask2(k);               // reverse_bases(rc) == complement of v
	return rc;
}

// ---------------------------------------------------------------------------------------
// KMC signatures (minimisers with exclusions).  Only the random-access lookups need them: the bin of a k-mer is
// signature_map[min over its m-mers of norm(m)] (kmer_api.h:653-673, kmc_file.cpp:339-340), where norm(m) is the smaller
// of m and its reverse complement among those that are allowed, and 4^len when neither is (mmer.h:33-88).  Not allowed:
// a TTT, TGT or TT* ending, an ACA beginning, and AA anywhere except as the first two bases.  m-mers are packed like
// k-mers (first base most significant); len is 5..11.
// ---------------------------------------------------------------------------------------
KMX_HD bool mmer_allowed(uint32_t m, int len) {
	const uint32_t tail = m & 0x3Fu;
	if (tail == 0x3Fu || tail == 0x3Bu || (m & 0x3Cu) == 0x3Cu) return false;
	const uint32_t is_a = ~(m | (m >> 1)) & 0x55555555u & (uint32_t)mask2(len);   // bit 2j: base j, counted from the end, is A
	if (is_a & (is_a >> 2) & (uint32_t)mask2(len - 2)) return false;              // bases j and j + 1 are A, j <= len - 3
	return (m >> (2 * (len - 3))) != 4u;                                          // ACA in front
}

KMX_HD uint32_t mmer_norm(uint32_t m, int len) {
	const uint32_t none = 1u << (2 * len);
	const uint32_t rc = (uint32_t)((~reverse_bases((uint64_t)m, len)) & mask2(len));
	const uint32_t a = mmer_allowed(m, len) ? m : none, b = mmer_allowed(rc, len) ? rc : none;
	return a < b ? a : b;
}

KMX_HD uint32_t kmer_signature(uint64_t v, int k, int len) {
	uint32_t best = 0xFFFFFFFFu;
	for (int shift = 2 * (k - len); shift >= 0; shift -= 2) {
		const uint32_t n = mmer_norm((uint32_t)(v >> shift) & (uint32_t)mask2(len), len);
		best = n < best ? n : best;
	}
	return best;
}

// 8 bases (16 bits, base j in bits [2j,2j+1]) -> 8 ASCII bytes, byte j = "ACGT"[base j]
KMX_HD uint64_t ascii8(uint32_t x) {
#ifdef __CUDA_ARCH__
	uint32_t s = x & 0xFFFFu;
	s = (s | (s << 8)) & 0x00FF00FFu;
	s = (s | (s << 4)) & 0x0F0F0F0Fu;
	s = (s | (s << 2)) & 0x33333333u;      // nibble j = base j: a PRMT selector
	uint32_t lo = __byte_perm(0x54474341u, 0u, s & 0xFFFFu);
	uint32_t hi = __byte_perm(0x54474341u, 0u, s >> 16);
	return ((uint64_t)hi << 32) | lo;
#else
	uint64_t w = 0;
	for (int j = 0; j < 8; j++) w |= (uint64_t)(uint8_t)("ACGT"[(x >> (2 * j)) & 3]) << (8 * j);
	return w;
#endif
}

// MurmurHash64A split in its seed-independent and seed-dependent halves.  The per-block
// mixing k*=m; k^=k>>47; k*=m does not involve the seed, so it is done once per string and
// reused by every seed (up to n_bits*n_hash = 35 of them per query).
struct HashPrep {
	uint64_t w[4];     // mixed full blocks
	uint64_t tail;     // raw tail bytes (len & 7 of them)
	uint64_t h0;       // len * m
};

KMX_HD void hash_prepare(uint64_t r, int len, HashPrep& p) {
	const int nblocks = len >> 3;
	const int tb = len & 7;
#pragma unroll
	for (int b = 0; b < 4; b++) {
		if (b < nblocks) {
			uint64_t w = ascii8((uint32_t)(r >> (16 * b)));
			w *= kMurM;
			w ^= w >> 47;
			w *= kMurM;
			p.w[b] = w;
		}
	}
	p.tail = 0;
	if (tb) {
		uint64_t t = ascii8(nblocks < 4 ? (uint32_t)(r >> (16 * nblocks)) : 0u);
		p.tail = t & ((1ULL << (8 * tb)) - 1);
	}
	p.h0 = (uint64_t)len * kMurM;
}

// the same preparation from the raw bytes of a string (len <= 39): what the reference hashes when a query keeps its
// original characters (tools.hpp:160-167 returns the string itself when the forward orientation is the smaller one)
KMX_HD void hash_prepare_bytes(const uint8_t* s, int len, HashPrep& p) {
	const int nblocks = len >> 3;
	const int tb = len & 7;
#pragma unroll
	for (int b = 0; b < 4; b++) {
		if (b < nblocks) {
			uint64_t w = 0;
			for (int j = 0; j < 8; j++) w |= (uint64_t)s[8 * b + j] << (8 * j);
			w *= kMurM;
			w ^= w >> 47;
			w *= kMurM;
			p.w[b] = w;
		}
	}
	p.tail = 0;
	for (int j = 0; j < tb; j++) p.tail |= (uint64_t)s[8 * nblocks + j] << (8 * j);
	p.h0 = (uint64_t)len * kMurM;
}

KMX_HD uint64_t hash_finish(const HashPrep& p, int len, uint32_t seed) {
	const int nblocks = len >> 3;
	uint64_t h = (uint64_t)seed ^ p.h0;
#pragma unroll
	for (int b = 0; b < 4; b++) {
		if (b < nblocks) {
			h ^= p.w[b];
			h *= kMurM;
		}
	}
	if (len & 7) {
		h ^= p.tail;
		h *= kMurM;
	}
	h ^= h >> 47;
	h *= kMurM;
	h ^= h >> 47;
	return h;
}

// the (k-2)-mer the "back" filters hash = kmer.substr(1, k-2)  (kmodel.hpp:388,475,548), in r-order
KMX_HD uint64_t middle_r(uint64_t r, int k) { return (r >> 2) & mask2(k - 2); }

// bit `pos` of a reference byte array lives in byte pos>>3 under mask 0x80>>(pos&7)
// (kmodel.hpp:576-588).  Seen as little-endian 32-bit words: word pos>>5, bit (pos&31)^7.
KMX_HD uint32_t bit_mask32(uint64_t pos) { return 1u << (((uint32_t)pos & 31u) ^ 7u); }

// tools.hpp:9 -- the 128 hash seeds (a format constant: consecutive primes from 46757)
#define KMX_SEED_LIST \
	46757, 46769, 46771, 46807, 46811, 46817, 46819, 46829, 46831, 46853, 46861, 46867, \
	46877, 46889, 46901, 46919, 46933, 46957, 46993, 46997, 47017, 47041, 47051, 47057, 47059, 47087, 47093, 47111, \
	47119, 47123, 47129, 47137, 47143, 47147, 47149, 47161, 47189, 47207, 47221, 47237, 47251, 47269, 47279, 47287, \
	47293, 47297, 47303, 47309, 47317, 47339, 47351, 47353, 47363, 47381, 47387, 47389, 47407, 47417, 47419, 47431, \
	47441, 47459, 47491, 47497, 47501, 47507, 47513, 47521, 47527, 47533, 47543, 47563, 47569, 47581, 47591, 47599, \
	47609, 47623, 47629, 47639, 47653, 47657, 47659, 47681, 47699, 47701, 47711, 47713, 47717, 47737, 47741, 47743, \
	47777, 47779, 47791, 47797, 47807, 47809, 47819, 47837, 47843, 47857, 47869, 47881, 47903, 47911, 47917, 47933, \
	47939, 47947, 47951, 47963, 47969, 47977, 47981, 48017, 48023, 48029, 48049, 48073, 48079, 48091, 48109, 48119, \
	48121, 48131, 48157, 48163
#ifdef __CUDACC__
static __constant__ uint32_t c_seeds[128] = { KMX_SEED_LIST };
#endif
static const uint32_t h_seeds[128] = { KMX_SEED_LIST };

}  // namespace kmx
