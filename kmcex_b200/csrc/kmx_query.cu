// kmx_query.cu -- batched KModel::kmer_to_occ on the device.
//
// Reference path reproduced (file:line relative to the reference root):
//   kmodel.hpp:100-116  kmer_to_occ(string)          kmodel.hpp:286-323  kmer_to_bin
//   kmodel.hpp:326-342  get_candidates               kmodel.hpp:344-359  get_neighbor_kmer_bin
//   kmodel.hpp:361-371  check_all_bf                 kmodel.hpp:373-390  check_(back_)bloomfilter
//   kmodel.hpp:625-671  find_bitarray / _one         rest.hpp:223-251    KRestData::check_kmer
//   tools.hpp:160-167   get_min_kmer                 occu_bin.hpp:67-83  occ_to_bin / bin_to_mean
//
// Two kernels.  query_fast_kernel (one thread per query) answers every query whose answer
// does not need the neighbours; the others (a Bloom hit next to one array candidate, or several
// array candidates: <1 % of a typical batch, but 9x the work) are appended to a list that
// query_slow_kernel works off with 8 lanes per query, one per neighbour, so that a warp of the
// fast kernel never waits for a single lane walking 8 neighbours.
//
// All probes of a stage are issued before any of them is tested (the reference short-circuits,
// but a probe has no side effect, so the answer is the same): the Bloom/km_back probes form one
// wave of independent 4-byte loads that overlaps the rest-table lookup, the n_bits*n_hash
// coupled-array probes one wave of 8-byte loads.  The hash of a (string, seed) pair is computed
// once and reduced modulo each filter length it is used with.
#include <cuda_runtime.h>
#include "kmx_device.cuh"
#include "kmx_launch.h"

namespace kmx {

template <int K, int H, int B, int L = 0>
struct QueryCfg {
	const DevModel& m;
	// L = 1 (models beyond the L2, m.query_l2 != 0): every probe carries an L2 eviction policy chosen per structure -- km_back is
	// touched by every query and small enough to stay resident (evict-last), the coupled arrays and the Bloom filters are far
	// larger than the L2 and stream through it (evict-first).  L = 0: plain read-only loads, no policy operand.
	unsigned long long pol_back, pol_cells, pol_bloom;
	__device__ __forceinline__ explicit QueryCfg(const DevModel& mm) : m(mm) {
		if (L) {
			const unsigned long long normal = make_evict_normal_policy();
			pol_back = (mm.query_l2 & 1) ? make_evict_last_policy() : normal;
			pol_cells = (mm.query_l2 & 2) ? make_evict_first_policy() : normal;
			pol_bloom = (mm.query_l2 & 4) ? make_evict_first_policy() : normal;
		} else {
			pol_back = pol_cells = pol_bloom = 0;
		}
	}
	// (the ordinary load path instead of the read-only one makes no difference: same rate, same DRAM bytes per probe --
	// profiles/r2_f_query_ld_ab.log)
	__device__ __forceinline__ bool test_bloom(const DevFilter& f, uint64_t h) const { return L ? filter_test_hint(f, h, pol_bloom) : filter_test(f, h); }
	__device__ __forceinline__ bool test_back(const DevFilter& f, uint64_t h) const { return L ? filter_test_hint(f, h, pol_back) : filter_test(f, h); }
	__device__ __forceinline__ unsigned long long load_cell(const unsigned long long* p) const { return L ? ldg_hint64(p, pol_cells) : __ldg(p); }
	__device__ __forceinline__ int k() const { return K ? K : m.k; }
	__device__ __forceinline__ int h() const { return H ? H : m.n_hash; }
	__device__ __forceinline__ int b() const { return B ? B : m.n_bits; }
};

__host__ __device__ constexpr int kHmax(int H) { return H ? H : kMaxHash; }
__host__ __device__ constexpr int kBmax(int B) { return B ? B : kMaxArrays; }

// the seed-dependent hashes every stage shares: seeds 0..H-2 on the k-mer, 0..H-3 on the (k-2)-mer
template <int K, int H, int B>
struct QueryHashes {
	HashPrep p31;
	uint64_t h31[kHmax(H) - 1], h29[kHmax(H) - 2];
	template <int L>
	__device__ __forceinline__ void compute(const QueryCfg<K, H, B, L>& c, uint64_t r) {
		HashPrep p29;
		hash_prepare(r, c.k(), p31);
		hash_prepare(middle_r(r, c.k()), c.k() - 2, p29);
		finish(c, p29);
	}
	// from the raw characters of the string: kmer and kmer.substr(1, k - 2) (kmodel.hpp:388)
	template <int L>
	__device__ __forceinline__ void compute_bytes(const QueryCfg<K, H, B, L>& c, const uint8_t* s) {
		HashPrep p29;
		hash_prepare_bytes(s, c.k(), p31);
		hash_prepare_bytes(s + 1, c.k() - 2, p29);
		finish(c, p29);
	}
	template <int L>
	__device__ __forceinline__ void finish(const QueryCfg<K, H, B, L>& c, const HashPrep& p29) {
#pragma unroll
		for (int j = 0; j < kHmax(H) - 1; j++)
			if (j < c.h() - 1) h31[j] = hash_finish(p31, c.k(), c_seeds[j]);
#pragma unroll
		for (int j = 0; j < kHmax(H) - 2; j++)
			if (j < c.h() - 2) h29[j] = hash_finish(p29, c.k() - 2, c_seeds[j]);
	}
};

// check_all_bf (kmodel.hpp:361-371): first filter pair, in the order {0} (ci == 1) or {1,0,2},
// whose k-mer filter (n_hash-1 hashes) and back filter (n_hash-2 hashes) both hit -> i + ci.
// Probed in two stages: the first kStageA hashes of every k-mer filter, then -- only for the
// filters that passed -- the remaining probes.  A filter is about half full, so an absent k-mer
// costs ~2 + 0.25 * 9 sector loads per pair instead of 11 (the reference short-circuits too).
constexpr int kStageA = 2;

template <int K, int H, int B, int L>
__device__ __forceinline__ int check_all_bf(const QueryCfg<K, H, B, L>& c, const QueryHashes<K, H, B>& q) {
	const DevModel& m = c.m;
	const int hb = c.h() - 1, hk = c.h() - 2;
	bool hit[kMaxBf];
#pragma unroll
	for (int i = 0; i < kMaxBf; i++) {
		hit[i] = false;
		if (i < m.bf_num) {
			bool ok = true;
#pragma unroll
			for (int j = 0; j < kStageA; j++)
				if (j < hb) ok &= c.test_bloom(m.bf[i], q.h31[j]);
			hit[i] = ok;
		}
	}
	if (H != 0) {
		// stage B with one pending filter pair per lane and iteration (see probe_arrays): a pair passes stage A for
		// a quarter of the lanes, so the warp is done after max-over-lanes(pending pairs) rounds instead of bf_num
		uint32_t pending = 0, okmask = 0;
#pragma unroll
		for (int i = 0; i < kMaxBf; i++)
			if (i < m.bf_num && hit[i]) pending |= 1u << i;
		while (pending) {
			const int i = __ffs((int)pending) - 1;
			pending &= pending - 1;
			bool ok = true;
#pragma unroll
			for (int j = kStageA; j < kHmax(H) - 1; j++) ok &= c.test_bloom(m.bf[i], q.h31[j]);
#pragma unroll
			for (int j = 0; j < kHmax(H) - 2; j++) ok &= c.test_bloom(m.bf_back[i], q.h29[j]);
			okmask |= (ok ? 1u : 0u) << i;
		}
#pragma unroll
		for (int i = 0; i < kMaxBf; i++) hit[i] = hit[i] && ((okmask >> i) & 1u) != 0;
	} else {
#pragma unroll
		for (int i = 0; i < kMaxBf; i++) {
			if (i < m.bf_num && hit[i]) {
				bool ok = true;
#pragma unroll
				for (int j = kStageA; j < kHmax(H) - 1; j++)
					if (j < hb) ok &= c.test_bloom(m.bf[i], q.h31[j]);
#pragma unroll
				for (int j = 0; j < kHmax(H) - 2; j++)
					if (j < hk) ok &= c.test_bloom(m.bf_back[i], q.h29[j]);
				hit[i] = ok;
			}
		}
	}
	if (m.bf_num == 1) return hit[0] ? m.ci : 0;
	if (hit[1]) return 1 + m.ci;
	if (hit[0]) return m.ci;
	if (hit[2]) return 2 + m.ci;
	return 0;
}

template <int K, int H, int B, int L>
__device__ __forceinline__ bool check_km_back(const QueryCfg<K, H, B, L>& c, const QueryHashes<K, H, B>& q) {
	const int hk = c.h() - 2;
	bool ok = true;
#pragma unroll
	for (int j = 0; j < kHmax(H) - 2; j++)
		if (j < hk) ok &= c.test_back(c.m.km_back, q.h29[j]);
	return ok;
}

// probe every coupled array (kmodel.hpp:625-671): are all tag bits set, and which bin do the value
// bits spell (tools.hpp:54-61).  Array i, hash j uses HashSeeds[(i*H+j)%128] (kmodel.hpp:450-453);
// for i == 0 those are the seeds whose hashes are already in q.h31.  Two stages: the first
// kStageA positions of every array, then the rest only for arrays whose first tags are all set
// (a k-mer that does not live in an array passes with probability fill^2, about 0.15): an
// answer needs all n_hash tags anyway, so skipping the rest cannot change it.
template <int K, int H, int B, int L>
__device__ __forceinline__ void probe_arrays(const QueryCfg<K, H, B, L>& c, const QueryHashes<K, H, B>& q, int* bins, bool* full) {
	const DevModel& m = c.m;
	auto position = [&](int i, int j) -> uint64_t {
		const uint64_t h = (i == 0 && j < c.h() - 1) ? q.h31[j < kHmax(H) - 1 ? j : 0] : hash_finish(q.p31, c.k(), m.arr_seed[i][j]);
		return fastmod(h, m.arr_mod);
	};
	unsigned long long cellA[kBmax(B)][kStageA];
	uint32_t shA[kBmax(B)][kStageA];
#pragma unroll
	for (int i = 0; i < kBmax(B); i++) {
#pragma unroll
		for (int j = 0; j < kStageA; j++) {
			if (i < c.b() && j < c.h()) {
				const uint64_t pos = position(i, j);
				shA[i][j] = ((uint32_t)pos & 31u) ^ 7u;
				cellA[i][j] = c.load_cell(m.cells[i] + (pos >> 5));
			}
		}
	}
#pragma unroll
	for (int i = 0; i < kBmax(B); i++) {
		int bin = 0;
		bool ok = i < c.b();
#pragma unroll
		for (int j = 0; j < kStageA; j++) {
			if (i < c.b() && j < c.h()) {
				const unsigned long long x = cellA[i][j] >> shA[i][j];
				bin |= (int)((uint32_t)x & 1u) << j;          // bit 0 = value
				ok &= ((uint32_t)(x >> 32) & 1u) != 0;         // bit 32 = tag
			}
		}
		bins[i] = bin;
		full[i] = ok;
	}
	if (H != 0 && B != 0 && H <= 8) {
		// Stage B, one pending array per lane and iteration: an array passes stage A for only a few lanes of a warp
		// (its resident k-mers plus ~14 % of the others), so walking the arrays one after the other issues the same
		// ~250 instructions n_bits times with 3-4 active lanes each.  Here every lane takes ITS next pending array
		// (array index, seeds and cell pointer become per-lane values), and the warp is done after
		// max-over-lanes(pending arrays) rounds, typically 2.  The packed results go back into bins / full.
		uint32_t pending = 0;
#pragma unroll
		for (int i = 0; i < kBmax(B); i++)
			if (i < c.b() && full[i]) pending |= 1u << i;
		unsigned long long packed = 0;
		uint32_t okmask = 0;
		while (pending) {
			const int i = __ffs((int)pending) - 1;
			pending &= pending - 1;
			const unsigned long long* cells = m.cells[i];
			unsigned long long cell[kHmax(H)];
			uint32_t sh[kHmax(H)];
#pragma unroll
			for (int j = kStageA; j < kHmax(H); j++) {
				const uint64_t pos = fastmod(hash_finish(q.p31, c.k(), m.arr_seed[i][j]), m.arr_mod);
				sh[j] = ((uint32_t)pos & 31u) ^ 7u;
				cell[j] = c.load_cell(cells + (pos >> 5));
			}
			uint32_t bin = 0;
			bool ok = true;
#pragma unroll
			for (int j = kStageA; j < kHmax(H); j++) {
				const unsigned long long x = cell[j] >> sh[j];
				bin |= ((uint32_t)x & 1u) << j;
				ok &= ((uint32_t)(x >> 32) & 1u) != 0;
			}
			packed |= (unsigned long long)bin << (8 * i);
			okmask |= (ok ? 1u : 0u) << i;
		}
#pragma unroll
		for (int i = 0; i < kBmax(B); i++) {
			if (i < c.b() && full[i]) {
				bins[i] |= (int)((packed >> (8 * i)) & 0xFFu);
				full[i] = ((okmask >> i) & 1u) != 0;
			}
		}
		return;
	}
#pragma unroll
	for (int i = 0; i < kBmax(B); i++) {
		if (i < c.b() && full[i]) {
			unsigned long long cell[kHmax(H)];
			uint32_t sh[kHmax(H)];
#pragma unroll
			for (int j = kStageA; j < kHmax(H); j++) {
				if (j < c.h()) {
					const uint64_t pos = position(i, j);
					sh[j] = ((uint32_t)pos & 31u) ^ 7u;
					cell[j] = c.load_cell(m.cells[i] + (pos >> 5));
				}
			}
			int bin = bins[i];
			bool ok = true;
#pragma unroll
			for (int j = kStageA; j < kHmax(H); j++) {
				if (j < c.h()) {
					const unsigned long long x = cell[j] >> sh[j];
					bin |= (int)((uint32_t)x & 1u) << j;
					ok &= ((uint32_t)(x >> 32) & 1u) != 0;
				}
			}
			bins[i] = bin;
			full[i] = ok;
		}
	}
}

// ---- rest table: KRestData::check_kmer (rest.hpp:223-251) ----------------------------------
// The reference searches [pre_buffer[g], pre_buffer[g+1]] with an INCLUSIVE upper bound.  Its
// outcome is: the count of an entry of group g whose suffix equals the key's, if there is one
// (the group is sorted, the extra slot is only probed once everything in the group compared
// smaller); otherwise, if the key's suffix is larger than every suffix of the group and equals
// the suffix of the FIRST entry of the next group (the slot the inclusive bound lets in), that
// entry's count -- the reference's false hit; otherwise 0.  A probe past the last entry is out
// of bounds there and counts as "no match" here.  Because keys are sorted globally, "in group g
// with equal suffix" is "equal full key", found through a bucket index over the top key bits;
// the false-hit suffix of each prefix is precomputed (rest_quirk_kernel).
// Split in two so that the first wave of loads (bucket bounds, false-hit suffix) is in flight together with the
// Bloom / km_back probes of the caller; the bucket's keys (about one per bucket) are then fetched in one wave.
struct RestProbe {
	uint32_t lo, hi;
	uint64_t quirk;
};

__device__ __forceinline__ RestProbe rest_begin(const DevRest& R, uint64_t v) {
	RestProbe p;
	const uint32_t b = (uint32_t)(v >> R.fine_shift);
	p.lo = __ldg(R.fine + b);
	p.hi = __ldg(R.fine + b + 1);
	p.quirk = __ldg(R.quirk_suffix + (uint32_t)(v >> R.suffix_bits));
	return p;
}

__device__ __forceinline__ int rest_finish(const DevRest& R, uint64_t v, const RestProbe& p) {
	uint32_t lo = p.lo, hi = p.hi;
	while (hi - lo > 4) {                          // long bucket (skewed data): bisect down first
		const uint32_t mid = (lo + hi) >> 1;
		if (__ldg(R.keys + mid) <= v) lo = mid; else hi = mid;
	}
	uint64_t key[4];
#pragma unroll
	for (uint32_t e = 0; e < 4; e++) key[e] = lo + e < hi ? __ldg(R.keys + lo + e) : ~0ULL;   // packed k-mers are below 2^64 - 1
	int found = -1;
#pragma unroll
	for (uint32_t e = 0; e < 4; e++)
		if (key[e] == v) found = (int)(lo + e);
	if (found >= 0) return __ldg(R.counts + found);
	if (p.quirk == (v & R.suffix_mask)) return __ldg(R.counts + __ldg(R.quirk_index + (uint32_t)(v >> R.suffix_bits)));
	return 0;
}

__device__ __forceinline__ int rest_lookup_indexed(const DevRest& R, uint64_t v) { return rest_finish(R, v, rest_begin(R, v)); }

// get_candidates (kmodel.hpp:326-342) for one neighbour in canonical form v with its hashes q; returns -1 when it adds nothing
template <int K, int H, int B, int L = 0>
__device__ __forceinline__ int neighbour_candidate_q(const DevModel& m, uint64_t v, const QueryHashes<K, H, B>& q) {
	QueryCfg<K, H, B, L> c(m);
	int occ = rest_lookup_indexed(m.rest, v);
	if (occ > 0) return (int)__ldg(m.occ2bin + (occ > m.cs ? m.cs : occ));   // occ > cs is out of bounds in the reference
	occ = check_all_bf(c, q);
	if (occ != 0) return occ;
	if (!check_km_back(c, q)) return -1;
	// find_bitarray_one (kmodel.hpp:650-671): last fully tagged array seen, stopping at the first non-zero bin
	int bins[kBmax(B)];
	bool full[kBmax(B)];
	probe_arrays(c, q, bins, full);
	int result = -1;
#pragma unroll
	for (int i = 0; i < kBmax(B); i++)
		if (i < c.b() && full[i] && (result <= 0)) result = bins[i];
	return result;
}

template <int K, int H, int B, int L = 0>
__device__ __forceinline__ int neighbour_candidate(const DevModel& m, uint64_t nb) {
	QueryCfg<K, H, B, L> c(m);
	uint64_t r;
	const uint64_t v = canonical(nb, c.k(), &r);
	QueryHashes<K, H, B> q;
	q.compute(c, r);
	return neighbour_candidate_q<K, H, B, L>(m, v, q);
}

// neighbour number j of the canonical k-mer v (kmodel.hpp:344-359): j < 4 successors (drop the
// first base, append ACGT[j]), j >= 4 predecessors (prepend ACGT[j-4], drop the last base)
__device__ __forceinline__ uint64_t neighbour_of(uint64_t v, int j, int k) {
	return j < 4 ? (((v << 2) & mask2(k)) | (uint64_t)j) : ((v >> 2) | ((uint64_t)(j - 4) << (2 * (k - 1))));
}

// everything kmer_to_occ decides before it needs the neighbours
struct Primary {
	int occ;          // Bloom answer (check_all_bf)
	int nc;           // array candidates (bins > 0 of fully tagged arrays)
	int cands[kMaxArrays];
	int path;
	int answer;       // final occurrence when path is not 5 / 6
};

template <int K, int H, int B, int L>
__device__ __forceinline__ int bin_to_mean(const QueryCfg<K, H, B, L>& c, int bin) {
	if (bin < c.m.end1) return bin;
	return bin < (1 << c.h()) ? __ldg(c.m.bin2mean + bin) : 0;   // unordered_map::operator[] yields 0 for a missing bin
}

template <int K, int H, int B, bool RAW = false, int L = 0>
__device__ __forceinline__ void query_primary(const DevModel& m, uint64_t v, uint64_t r, Primary& P, const uint8_t* raw = nullptr) {
	QueryCfg<K, H, B, L> c(m);
	// the rest lookup's loads and the Bloom wave are independent of each other
	const RestProbe rp = rest_begin(m.rest, v);
	QueryHashes<K, H, B> q;
	if (RAW) q.compute_bytes(c, raw);
	else q.compute(c, r);
	const bool in_back = check_km_back(c, q);
	P.occ = check_all_bf(c, q);
	const int rest = rest_finish(m.rest, v, rp);
	P.nc = 0;
	if (rest != 0) {                       // kmodel.hpp:104-105
		P.path = 1;
		P.answer = rest;
		return;
	}
	if (!in_back) {                        // kmodel.hpp:109-111: Bloom answer (possibly 0) when km_back misses
		P.path = 2;
		P.answer = P.occ;
		return;
	}
	int bins[kBmax(B)];
	bool full[kBmax(B)];
	probe_arrays(c, q, bins, full);
#pragma unroll
	for (int i = 0; i < kBmax(B); i++)
		if (i < c.b() && full[i] && bins[i] > 0) P.cands[P.nc++] = bins[i];
	if (P.nc == 0) {
		P.path = 3;
		P.answer = bin_to_mean(c, P.occ);
	} else if (P.nc == 1 && P.occ == 0) {
		P.path = 4;
		P.answer = bin_to_mean(c, P.cands[0]);
	} else {
		P.path = P.nc == 1 ? 5 : 6;
		P.answer = 0;
	}
}

// 2-bit encode of ASCII k-mers (tools.hpp:63-76: bytes other than C/G/T encode as A).  A block stages the
// contiguous bytes of its 256 k-mers in shared memory with coalesced 16-byte loads (a thread reading its own
// k-mer straight from global memory would issue k byte loads at a stride of `stride` bytes across the warp).
// A string with a byte outside "ACGT" (N, lower case, ...) is also listed in `dirty`: the reference hashes the RAW
// characters of such a string when its forward orientation is the canonical one (tools.hpp:160-167), which the packed
// path cannot reproduce; query_raw_kernel answers those afterwards.
constexpr int kPackMaxStride = 64;

__device__ __forceinline__ uint64_t pack_bytes(const uint8_t* p, int k, bool* dirty) {
	uint64_t v = 0;
	bool bad = false;
	for (int i = 0; i < k; i++) {
		const uint8_t ch = p[i];
		const uint64_t code = ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 0;
		bad |= code == 0 && ch != 'A';
		v = (v << 2) | code;
	}
	*dirty = bad;
	return v;
}

__device__ __forceinline__ void note_dirty(uint32_t index, uint32_t* dirty, unsigned int* dirty_n) {
	dirty[atomicAdd(dirty_n, 1u)] = index;                   // rare: no aggregation
}

__global__ void __launch_bounds__(256) ascii_pack_kernel(const char* __restrict__ flat, size_t stride, size_t n, int k, uint64_t* __restrict__ packed,
                                                         uint32_t* __restrict__ dirty, unsigned int* __restrict__ dirty_n) {
	__shared__ uint4 s_buf[256 * kPackMaxStride / 16 + 1];
	const uint8_t* s_bytes = reinterpret_cast<const uint8_t*>(s_buf);
	for (size_t base = (size_t)blockIdx.x * 256; base < n; base += (size_t)gridDim.x * 256) {
		const size_t cnt = min((size_t)256, n - base);
		const size_t byte0 = base * stride, bytes = cnt * stride;
		const size_t head = (size_t)(reinterpret_cast<uintptr_t>(flat + byte0) & 15);   // the aligned window starts `head` bytes earlier
		const uint4* src = reinterpret_cast<const uint4*>(flat + byte0 - head);         // (inside the same allocation)
		const size_t n_full = (head + bytes) >> 4;               // whole 16-byte vectors; the tail goes byte by byte
		__syncthreads();
		for (size_t i = threadIdx.x; i < n_full; i += 256) s_buf[i] = __ldg(src + i);
		{
			uint8_t* s_w = reinterpret_cast<uint8_t*>(s_buf);
			for (size_t i = (n_full << 4) + threadIdx.x; i < head + bytes; i += 256) s_w[i] = (uint8_t)flat[byte0 - head + i];
		}
		__syncthreads();
		if (threadIdx.x < cnt) {
			bool bad;
			packed[base + threadIdx.x] = pack_bytes(s_bytes + head + threadIdx.x * stride, k, &bad);
			if (bad) note_dirty((uint32_t)(base + threadIdx.x), dirty, dirty_n);
		}
	}
}

// fallback for strides beyond the staging buffer: one thread per k-mer, straight from global memory
__global__ void ascii_pack_wide_kernel(const char* __restrict__ flat, size_t stride, size_t n, int k, uint64_t* __restrict__ packed,
                                       uint32_t* __restrict__ dirty, unsigned int* __restrict__ dirty_n) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		bool bad;
		packed[i] = pack_bytes(reinterpret_cast<const uint8_t*>(flat + i * stride), k, &bad);
		if (bad) note_dirty((uint32_t)i, dirty, dirty_n);
	}
}

// ---- queries that keep characters outside "ACGT" --------------------------------------------------
// The reference never validates a query.  get_min_kmer (tools.hpp:160-167) 2-bit encodes the string with every byte that
// is not C/G/T read as A (tools.hpp:63-76), and returns THE ORIGINAL STRING when that encoding is <= its reverse
// complement, else the decoded reverse complement.  The rest lookup re-encodes the returned string (rest.hpp:22-34,223-251),
// but the Bloom / km_back / coupled-array hashes run over its raw bytes, N and lower case included (kmodel.hpp:373-390,
// 625-646); the neighbours are built from the same characters (kmodel.hpp:344-359) and canonicalised again.  One thread per
// such query, generic geometry: the path is rare (reads with N) and only has to be right.
__device__ __forceinline__ uint64_t raw_canonical(uint8_t* s, int k) {
	bool bad;
	const uint64_t u = pack_bytes(s, k, &bad);
	const uint64_t rc = (~reverse_bases(u, k)) & mask2(k);
	if (u <= rc) return u;                                   // the string stays as it is
	for (int i = k - 1, sh = 0; i >= 0; i--, sh += 2) s[i] = (uint8_t)("ACGT"[(rc >> sh) & 3]);   // tools.hpp:90-100
	return rc;
}

__global__ void __launch_bounds__(64) query_raw_kernel(const __grid_constant__ DevModel m, const char* __restrict__ flat, size_t stride,
                                                       const uint32_t* __restrict__ dirty, const unsigned int* __restrict__ dirty_n,
                                                       int32_t* __restrict__ out) {
	QueryCfg<0, 0, 0> c(m);
	const int k = m.k;
	const unsigned int n = *dirty_n;
	for (unsigned int x = blockIdx.x * blockDim.x + threadIdx.x; x < n; x += gridDim.x * blockDim.x) {
		const uint32_t index = dirty[x];
		uint8_t s[32], t[32];
		for (int i = 0; i < k; i++) s[i] = (uint8_t)flat[(size_t)index * stride + i];
		const uint64_t v = raw_canonical(s, k);
		Primary P;
		query_primary<0, 0, 0, true>(m, v, 0, P, s);
		if (P.path < 5) {
			out[index] = P.answer;
			continue;
		}
		// kmer_to_bin's neighbour rules (kmodel.hpp:286-323) over the 8 neighbours of the (possibly raw) canonical string
		int n_cand = 0, n_low = 0, dist[kMaxArrays];
		for (int i = 0; i < kMaxArrays; i++) dist[i] = 2 << 20;
		for (int j = 0; j < 8; j++) {
			if (j < 4) {
				for (int i = 0; i + 1 < k; i++) t[i] = s[i + 1];
				t[k - 1] = (uint8_t)("ACGT"[j]);
			} else {
				for (int i = 1; i < k; i++) t[i] = s[i - 1];
				t[0] = (uint8_t)("ACGT"[j - 4]);
			}
			const uint64_t nv = raw_canonical(t, k);
			QueryHashes<0, 0, 0> q;
			q.compute_bytes(c, t);
			const int cand = neighbour_candidate_q<0, 0, 0>(m, nv, q);
			if (cand < 0) continue;
			n_cand++;
			n_low += cand < m.ci + m.bf_num ? 1 : 0;
			for (int i = 0; i < P.nc; i++) {
				int d = P.cands[i] - cand;
				d = d < 0 ? -d : d;
				dist[i] = d < dist[i] ? d : dist[i];
			}
		}
		int bin;
		if (P.nc == 1) {
			bin = (n_low >= n_cand / 2) ? P.occ : P.cands[0];
		} else if (n_cand <= 0) {
			bin = 0;
		} else {
			int min_dist = 2 << 20;
			bin = P.cands[0];
			for (int i = 0; i < P.nc; i++) {
				if (min_dist > dist[i]) {
					min_dist = dist[i];
					bin = P.cands[i];
				}
			}
		}
		out[index] = bin_to_mean(c, bin);
	}
}

#ifndef KMX_QUERY_BLOCKS
#define KMX_QUERY_BLOCKS 2
#endif
template <int K, int H, int B, int L>
__global__ void __launch_bounds__(256, KMX_QUERY_BLOCKS) query_fast_kernel(const __grid_constant__ DevModel m, const uint64_t* __restrict__ input,
                                                            size_t n, int32_t* __restrict__ out, int32_t* __restrict__ path_out,
                                                            DeferredQuery* __restrict__ defer, unsigned int* __restrict__ defer_n) {
	const int k = K ? K : m.k;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		const uint64_t raw = input[i];
		uint64_t r;
		const uint64_t v = canonical(raw & mask2(k), k, &r);
		Primary P;
		query_primary<K, H, B, false, L>(m, v, r, P);
		if (path_out) path_out[i] = P.path;
		if (P.path >= 5) {
			// warp-aggregated append
			const unsigned int active = __activemask();
			const int lane = threadIdx.x & 31;
			const int leader = __ffs(active) - 1;
			unsigned int at = 0;
			if (lane == leader) at = atomicAdd(defer_n, (unsigned int)__popc(active));
			at = __shfl_sync(active, at, leader);
			at += __popc(active & ((1u << lane) - 1u));
			defer[at].kmer = v;
			defer[at].index = (uint32_t)i;
		} else if (out) {
			out[i] = P.answer;
		}
	}
}

// kmer_to_bin's neighbour rules (kmodel.hpp:286-323), 8 lanes per deferred query
template <int K, int H, int B, int L>
__global__ void __launch_bounds__(256, 2) query_slow_kernel(const __grid_constant__ DevModel m, const DeferredQuery* __restrict__ defer,
                                                            const unsigned int* __restrict__ defer_n, int32_t* __restrict__ out) {
	QueryCfg<K, H, B, L> c(m);
	const int k = c.k();
	const unsigned int n = *defer_n;
	const unsigned int groups_per_grid = gridDim.x * (blockDim.x / 8);
	const int sub = threadIdx.x & 7;
	for (unsigned int base = 0; base < n; base += groups_per_grid) {     // uniform trip count: the shuffles below need the whole warp
		const unsigned int g = base + blockIdx.x * (blockDim.x / 8) + (threadIdx.x >> 3);
		const bool live = g < n;
		const uint64_t v = live ? defer[g].kmer : 0;
		Primary P;
		P.nc = 0;
		P.occ = 0;
		int x = -1;
		if (live) {
			query_primary<K, H, B, false, L>(m, v, reverse_bases(v, k), P);   // the same decision the fast kernel took
			x = neighbour_candidate<K, H, B, L>(m, neighbour_of(v, sub, k));
		}
		// reductions over the 8 lanes of the group
		int n_cand = x >= 0 ? 1 : 0, n_low = (x >= 0 && x < m.ci + m.bf_num) ? 1 : 0;
		int dist[kBmax(B)];
#pragma unroll
		for (int i = 0; i < kBmax(B); i++) {
			int d = 2 << 20;
			if (i < P.nc && x >= 0) {
				d = P.cands[i] - x;
				d = d < 0 ? -d : d;
			}
			dist[i] = d;
		}
#pragma unroll
		for (int off = 4; off > 0; off >>= 1) {
			n_cand += __shfl_xor_sync(0xffffffffu, n_cand, off, 8);
			n_low += __shfl_xor_sync(0xffffffffu, n_low, off, 8);
#pragma unroll
			for (int i = 0; i < kBmax(B); i++) {
				const int o = __shfl_xor_sync(0xffffffffu, dist[i], off, 8);
				dist[i] = o < dist[i] ? o : dist[i];
			}
		}
		if (live && sub == 0) {
			int bin;
			if (P.nc == 1) {                   // kmodel.hpp:293-301: Bloom hit + one array candidate: neighbour vote
				bin = (n_low >= n_cand / 2) ? P.occ : P.cands[0];
			} else if (n_cand <= 0) {          // kmodel.hpp:305-309
				bin = 0;
			} else {                           // kmodel.hpp:310-322: candidate closest to any neighbour, first wins ties
				int min_dist = 2 << 20;
				bin = P.cands[0];
#pragma unroll
				for (int i = 0; i < kBmax(B); i++) {
					if (i < P.nc && min_dist > dist[i]) {
						min_dist = dist[i];
						bin = P.cands[i];
					}
				}
			}
			out[defer[g].index] = bin_to_mean(c, bin);
		}
	}
}

// rest table side structures (built once per model, after the rest table itself)
__global__ void rest_fine_kernel(const uint64_t* __restrict__ keys, uint32_t n, int fine_shift, uint32_t n_buckets, uint32_t* __restrict__ fine) {
	for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b <= n_buckets; b += gridDim.x * blockDim.x) {
		// first entry whose bucket (key >> fine_shift) is >= b
		uint32_t lo = 0, hi = n;
		while (lo < hi) {
			const uint32_t mid = (lo + hi) >> 1;
			if ((keys[mid] >> fine_shift) < (uint64_t)b) lo = mid + 1; else hi = mid;
		}
		fine[b] = lo;
	}
}

__global__ void rest_quirk_kernel(const uint64_t* __restrict__ keys, uint64_t n, const int32_t* __restrict__ hash2index,
                                  const int32_t* __restrict__ pre_buffer, int map_size, uint64_t suffix_mask,
                                  uint64_t* __restrict__ quirk_suffix, uint32_t* __restrict__ quirk_index) {
	for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < map_size; p += gridDim.x * blockDim.x) {
		uint64_t s = ~0ULL;                    // never equals a masked suffix
		uint32_t at = 0;
		const int g = hash2index[p];
		if (g >= 0) {
			const long long lo = pre_buffer[g], hi = pre_buffer[g + 1];
			if (hi > lo && (uint64_t)hi < n) {
				const uint64_t next_first = keys[hi] & suffix_mask, last = keys[hi - 1] & suffix_mask;
				if (next_first > last) {
					s = next_first;
					at = (uint32_t)hi;
				}
			}
		}
		quirk_suffix[p] = s;
		quirk_index[p] = at;
	}
}

cudaError_t launch_rest_side_tables(const DevRest& R, int map_size, uint32_t* d_fine, uint64_t* d_quirk_suffix, uint32_t* d_quirk_index,
                                    cudaStream_t stream) {
	const uint32_t n_buckets = 1u << R.fine_bits;
	note_launch(2);
	rest_fine_kernel<<<(n_buckets + 256) / 256, 256, 0, stream>>>(R.keys, (uint32_t)R.count, R.fine_shift, n_buckets, d_fine);
	rest_quirk_kernel<<<(map_size + 255) / 256, 256, 0, stream>>>(R.keys, R.count, R.hash2index, R.pre_buffer, map_size, R.suffix_mask,
	                                                              d_quirk_suffix, d_quirk_index);
	return cudaGetLastError();
}

static int query_grid(size_t n, int sm_count) {
	size_t blocks = (n + 255) / 256;
	size_t cap = (size_t)sm_count * 16;
	return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

static cudaError_t query_packed_launches(const DevModel& m, const uint64_t* d_kmers, size_t n, int32_t* d_out, int32_t* d_path,
                                         DeferredQuery* d_defer, unsigned int* d_defer_n, int sm_count, cudaStream_t stream) {
	const int grid = query_grid(n, sm_count);
	const int slow_grid = sm_count * 2;
	if (m.k == 31 && m.n_hash == 7 && m.n_bits == 5 && m.query_l2 == 0) {
		query_fast_kernel<31, 7, 5, 0><<<grid, 256, 0, stream>>>(m, d_kmers, n, d_out, d_path, d_defer, d_defer_n);
		if (d_out) query_slow_kernel<31, 7, 5, 0><<<slow_grid, 256, 0, stream>>>(m, d_defer, d_defer_n, d_out);
	} else if (m.k == 31 && m.n_hash == 7 && m.n_bits == 5) {
		query_fast_kernel<31, 7, 5, 1><<<grid, 256, 0, stream>>>(m, d_kmers, n, d_out, d_path, d_defer, d_defer_n);
		if (d_out) query_slow_kernel<31, 7, 5, 1><<<slow_grid, 256, 0, stream>>>(m, d_defer, d_defer_n, d_out);
	} else {
		query_fast_kernel<0, 0, 0, 0><<<grid, 256, 0, stream>>>(m, d_kmers, n, d_out, d_path, d_defer, d_defer_n);
		if (d_out) query_slow_kernel<0, 0, 0, 0><<<slow_grid, 256, 0, stream>>>(m, d_defer, d_defer_n, d_out);
	}
	note_launch(d_out ? 2 : 1);
	return cudaGetLastError();
}

cudaError_t launch_query_packed(const DevModel& m, const uint64_t* d_kmers, size_t n, int32_t* d_out, int32_t* d_path,
                                DeferredQuery* d_defer, unsigned int* d_counters, int sm_count, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	cudaError_t e = cudaMemsetAsync(d_counters, 0, 2 * sizeof(unsigned int), stream);
	if (e != cudaSuccess) return e;
	return query_packed_launches(m, d_kmers, n, d_out, d_path, d_defer, d_counters, sm_count, stream);
}

// ASCII batches are 2-bit encoded into d_packed (room for n words) and take the packed path; strings with characters outside
// "ACGT" are answered again, from their raw bytes, by query_raw_kernel
cudaError_t launch_query_ascii(const DevModel& m, const char* d_flat, size_t stride, size_t n, int32_t* d_out, uint64_t* d_packed,
                               DeferredQuery* d_defer, uint32_t* d_dirty, unsigned int* d_counters, int sm_count, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	cudaError_t e = cudaMemsetAsync(d_counters, 0, 2 * sizeof(unsigned int), stream);
	if (e != cudaSuccess) return e;
	const int grid = query_grid(n, sm_count);
	if (stride <= (size_t)kPackMaxStride) ascii_pack_kernel<<<grid, 256, 0, stream>>>(d_flat, stride, n, m.k, d_packed, d_dirty, d_counters + 1);
	else ascii_pack_wide_kernel<<<grid, 256, 0, stream>>>(d_flat, stride, n, m.k, d_packed, d_dirty, d_counters + 1);
	note_launch();
	e = cudaGetLastError();
	if (e != cudaSuccess) return e;
	e = query_packed_launches(m, d_packed, n, d_out, nullptr, d_defer, d_counters, sm_count, stream);
	if (e != cudaSuccess) return e;
	query_raw_kernel<<<sm_count, 64, 0, stream>>>(m, d_flat, stride, d_dirty, d_counters + 1, d_out);
	note_launch();
	return cudaGetLastError();
}

}  // namespace kmx
