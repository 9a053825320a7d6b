// kmx_query.cu -- batched KModel::kmer_to_occ on the device (one thread per query).
//
// Reference path reproduced (file:line relative to the reference root):
//   kmodel.hpp:100-116  kmer_to_occ(string)          kmodel.hpp:286-323  kmer_to_bin
//   kmodel.hpp:326-342  get_candidates               kmodel.hpp:344-359  get_neighbor_kmer_bin
//   kmodel.hpp:361-371  check_all_bf                 kmodel.hpp:373-390  check_(back_)bloomfilter
//   kmodel.hpp:625-671  find_bitarray / _one         rest.hpp:223-251    KRestData::check_kmer
//   tools.hpp:160-167   get_min_kmer                 occu_bin.hpp:67-83  occ_to_bin / bin_to_mean
//
// All probes of a stage are issued before any of them is tested (the reference short-circuits,
// but a probe has no side effect, so the answer is the same): the Bloom/km_back probes form one
// wave of independent 4-byte loads, the n_bits*n_hash coupled-array probes one wave of 8-byte
// loads.  The hash of a (string, seed) pair is computed once and reduced modulo each filter
// length it is used with: seeds 0..n_hash-2 serve every Bloom filter and the first array.
#include <cuda_runtime.h>
#include "kmx_device.cuh"
#include "kmx_launch.h"

namespace kmx {

template <int K, int H, int B>
struct QueryCfg {
	const DevModel& m;
	__device__ __forceinline__ explicit QueryCfg(const DevModel& mm) : m(mm) {}
	__device__ __forceinline__ int k() const { return K ? K : m.k; }
	__device__ __forceinline__ int h() const { return H ? H : m.n_hash; }
	__device__ __forceinline__ int b() const { return B ? B : m.n_bits; }
};

__host__ __device__ constexpr int kHmax(int H) { return H ? H : kMaxHash; }
__host__ __device__ constexpr int kBmax(int B) { return B ? B : kMaxArrays; }

// check_all_bf (kmodel.hpp:361-371): first filter pair, in the order {0} (ci == 1) or {1,0,2},
// whose k-mer filter (n_hash-1 hashes) and back filter (n_hash-2 hashes) both hit -> i + ci
template <int K, int H, int B>
__device__ __forceinline__ int check_all_bf(const QueryCfg<K, H, B>& c, const uint64_t* h31, const uint64_t* h29) {
	const DevModel& m = c.m;
	const int hb = c.h() - 1, hk = c.h() - 2;
	bool hit[kMaxBf];
#pragma unroll
	for (int i = 0; i < kMaxBf; i++) {
		hit[i] = false;
		if (i < m.bf_num) {
			bool ok = true;
#pragma unroll
			for (int j = 0; j < kHmax(H) - 1; j++)
				if (j < hb) ok &= filter_test(m.bf[i], h31[j]);
#pragma unroll
			for (int j = 0; j < kHmax(H) - 2; j++)
				if (j < hk) ok &= filter_test(m.bf_back[i], h29[j]);
			hit[i] = ok;
		}
	}
	if (m.bf_num == 1) return hit[0] ? m.ci : 0;
	if (hit[1]) return 1 + m.ci;
	if (hit[0]) return m.ci;
	if (hit[2]) return 2 + m.ci;
	return 0;
}

template <int K, int H, int B>
__device__ __forceinline__ bool check_km_back(const QueryCfg<K, H, B>& c, const uint64_t* h29) {
	const int hk = c.h() - 2;
	bool ok = true;
#pragma unroll
	for (int j = 0; j < kHmax(H) - 2; j++)
		if (j < hk) ok &= filter_test(c.m.km_back, h29[j]);
	return ok;
}

// all seeds the coupled arrays use on the k-mer string: array i, hash j -> HashSeeds[(i*H+j)%128]
// (kmodel.hpp:450-453).  For i*H+j < 128 that is simply seed number i*H+j.
template <int K, int H, int B>
__device__ __forceinline__ uint64_t array_hash(const QueryCfg<K, H, B>& c, const HashPrep& p31, int i, int j) {
	return hash_finish(p31, c.k(), c.m.arr_seed[i][j]);
}

// decode array i: tag bits all set? and the bin spelled by the value bits (tools.hpp:54-61)
template <int K, int H, int B>
__device__ __forceinline__ void probe_arrays(const QueryCfg<K, H, B>& c, const HashPrep& p31, int* bins, bool* full) {
	const DevModel& m = c.m;
	unsigned long long cell[kBmax(B)][kHmax(H)];
	uint32_t sh[kBmax(B)][kHmax(H)];
#pragma unroll
	for (int i = 0; i < kBmax(B); i++) {
#pragma unroll
		for (int j = 0; j < kHmax(H); j++) {
			if (i < c.b() && j < c.h()) {
				uint64_t pos = fastmod(array_hash(c, p31, i, j), m.arr_mod);
				sh[i][j] = ((uint32_t)pos & 31u) ^ 7u;
				cell[i][j] = __ldg(m.cells[i] + (pos >> 5));
			}
		}
	}
#pragma unroll
	for (int i = 0; i < kBmax(B); i++) {
		int bin = 0;
		bool ok = true;
#pragma unroll
		for (int j = 0; j < kHmax(H); j++) {
			if (i < c.b() && j < c.h()) {
				uint32_t val = (uint32_t)cell[i][j], tag = (uint32_t)(cell[i][j] >> 32);
				bin |= (int)((val >> sh[i][j]) & 1u) << j;
				ok &= ((tag >> sh[i][j]) & 1u) != 0;
			}
		}
		bins[i] = bin;
		full[i] = ok;
	}
}

// get_candidates (kmodel.hpp:326-342) for one neighbour; returns -1 when it adds nothing
template <int K, int H, int B>
__device__ __noinline__ int neighbour_candidate(const DevModel& m, uint64_t nb) {
	QueryCfg<K, H, B> c(m);
	uint64_t r;
	uint64_t v = canonical(nb, c.k(), &r);
	int occ = rest_lookup(m.rest, v);
	if (occ > 0) return (int)__ldg(m.occ2bin + (occ > m.cs ? m.cs : occ));   // occ > cs is out of bounds in the reference
	HashPrep p31, p29;
	hash_prepare(r, c.k(), p31);
	hash_prepare(middle_r(r, c.k()), c.k() - 2, p29);
	uint64_t h31[kHmax(H)], h29[kHmax(H)];
#pragma unroll
	for (int j = 0; j < kHmax(H) - 1; j++)
		if (j < c.h() - 1) h31[j] = hash_finish(p31, c.k(), c_seeds[j]);
#pragma unroll
	for (int j = 0; j < kHmax(H) - 2; j++)
		if (j < c.h() - 2) h29[j] = hash_finish(p29, c.k() - 2, c_seeds[j]);
	occ = check_all_bf(c, h31, h29);
	if (occ != 0) return occ;
	if (!check_km_back(c, h29)) return -1;
	// find_bitarray_one (kmodel.hpp:650-671): last fully tagged array seen, stopping at the first non-zero bin
	int bins[kBmax(B)];
	bool full[kBmax(B)];
	probe_arrays(c, p31, bins, full);
	int result = -1;
#pragma unroll
	for (int i = 0; i < kBmax(B); i++) {
		if (i < c.b() && full[i] && (result <= 0)) result = bins[i];
	}
	return result;
}

// get_neighbor_kmer_bin (kmodel.hpp:344-359): successors (drop first base, append A,C,G,T) then
// predecessors (prepend A,C,G,T, drop last base), on the canonical string of the query
template <int K, int H, int B>
__device__ __noinline__ int neighbour_bins(const DevModel& m, uint64_t v, int* cand) {
	const int k = K ? K : m.k;
	int n = 0;
	for (int b = 0; b < 4; b++) {
		int x = neighbour_candidate<K, H, B>(m, ((v << 2) & mask2(k)) | (uint64_t)b);
		if (x >= 0) cand[n++] = x;
	}
	for (int b = 0; b < 4; b++) {
		int x = neighbour_candidate<K, H, B>(m, (v >> 2) | ((uint64_t)b << (2 * (k - 1))));
		if (x >= 0) cand[n++] = x;
	}
	return n;
}

// kmer_to_occ for one packed k-mer; *path gets the path class documented in kmx.h
template <int K, int H, int B>
__device__ __forceinline__ int query_one(const DevModel& m, uint64_t raw, int* path) {
	QueryCfg<K, H, B> c(m);
	const int k = c.k();
	uint64_t r;
	uint64_t v = canonical(raw & mask2(k), k, &r);
	int occ = rest_lookup(m.rest, v);
	if (occ != 0) {
		*path = 1;
		return occ;
	}
	HashPrep p31, p29;
	hash_prepare(r, k, p31);
	hash_prepare(middle_r(r, k), k - 2, p29);
	uint64_t h31[kHmax(H)], h29[kHmax(H)];
#pragma unroll
	for (int j = 0; j < kHmax(H) - 1; j++)
		if (j < c.h() - 1) h31[j] = hash_finish(p31, k, c_seeds[j]);
#pragma unroll
	for (int j = 0; j < kHmax(H) - 2; j++)
		if (j < c.h() - 2) h29[j] = hash_finish(p29, k - 2, c_seeds[j]);
	bool in_back = check_km_back(c, h29);
	occ = check_all_bf(c, h31, h29);
	if (!in_back) {
		*path = 2;
		return occ;          // kmodel.hpp:109-111: Bloom answer (possibly 0) when km_back misses
	}
	int bins[kBmax(B)];
	bool full[kBmax(B)];
	probe_arrays(c, p31, bins, full);
	int cands[kBmax(B)];
	int nc = 0;
#pragma unroll
	for (int i = 0; i < kBmax(B); i++)
		if (i < c.b() && full[i] && bins[i] > 0) cands[nc++] = bins[i];
	int bin;
	if (nc == 0) {
		*path = 3;
		bin = occ;
	} else if (nc == 1) {
		*path = occ ? 5 : 4;
		bin = cands[0];
		if (occ) {
			int nb[8];
			int n = neighbour_bins<K, H, B>(m, v, nb);
			int low = 0;
			for (int i = 0; i < n; i++) low += nb[i] < m.ci + m.bf_num;
			if (low >= n / 2) bin = occ;
		}
	} else {
		*path = 6;
		int nb[8];
		int n = neighbour_bins<K, H, B>(m, v, nb);
		if (n <= 0) {
			bin = 0;
		} else {
			int min_dist = 2 << 20;
			bin = cands[0];
			for (int i = 0; i < nc; i++) {
				int cur = 2 << 20;
				for (int j = 0; j < n; j++) {
					int d = cands[i] - nb[j];
					d = d < 0 ? -d : d;
					cur = d < cur ? d : cur;
				}
				if (min_dist > cur) {
					min_dist = cur;
					bin = cands[i];
				}
			}
		}
	}
	if (bin < m.end1) return bin;
	return bin < (1 << c.h()) ? __ldg(m.bin2mean + bin) : 0;   // unordered_map::operator[] yields 0 for a missing bin
}

// 2-bit encode of an ASCII k-mer (tools.hpp:63-76: bytes other than C/G/T encode as A)
__device__ __forceinline__ uint64_t encode_ascii(const char* s, int k) {
	uint64_t v = 0;
	for (int i = 0; i < k; i++) {
		char ch = s[i];
		uint64_t code = ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 0;
		v = (v << 2) | code;
	}
	return v;
}

template <int K, int H, int B>
__global__ void __launch_bounds__(256) query_packed_kernel(const __grid_constant__ DevModel m, const uint64_t* __restrict__ kmers,
                                                           size_t n, int32_t* __restrict__ out, int32_t* __restrict__ path_out) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		int path = 0;
		int occ = query_one<K, H, B>(m, kmers[i], &path);
		if (out) out[i] = occ;
		if (path_out) path_out[i] = path;
	}
}

template <int K, int H, int B>
__global__ void __launch_bounds__(256) query_ascii_kernel(const __grid_constant__ DevModel m, const char* __restrict__ flat, size_t stride,
                                                          size_t n, int32_t* __restrict__ out) {
	const int k = K ? K : m.k;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		int path = 0;
		out[i] = query_one<K, H, B>(m, encode_ascii(flat + i * stride, k), &path);
	}
}

static int query_grid(size_t n, int sm_count) {
	size_t blocks = (n + 255) / 256;
	size_t cap = (size_t)sm_count * 16;
	return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

cudaError_t launch_query_packed(const DevModel& m, const uint64_t* d_kmers, size_t n, int32_t* d_out, int32_t* d_path,
                                int sm_count, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	int grid = query_grid(n, sm_count);
	if (m.k == 31 && m.n_hash == 7 && m.n_bits == 5)
		query_packed_kernel<31, 7, 5><<<grid, 256, 0, stream>>>(m, d_kmers, n, d_out, d_path);
	else
		query_packed_kernel<0, 0, 0><<<grid, 256, 0, stream>>>(m, d_kmers, n, d_out, d_path);
	return cudaGetLastError();
}

cudaError_t launch_query_ascii(const DevModel& m, const char* d_flat, size_t stride, size_t n, int32_t* d_out, int sm_count,
                               cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	int grid = query_grid(n, sm_count);
	if (m.k == 31 && m.n_hash == 7 && m.n_bits == 5)
		query_ascii_kernel<31, 7, 5><<<grid, 256, 0, stream>>>(m, d_flat, stride, n, d_out);
	else
		query_ascii_kernel<0, 0, 0><<<grid, 256, 0, stream>>>(m, d_flat, stride, n, d_out);
	return cudaGetLastError();
}

}  // namespace kmx
