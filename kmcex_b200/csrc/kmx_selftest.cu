// kmx_selftest.cu -- device-side known-answer hook for the addressing chain every build / query kernel uses:
//   packed k-mer -> ASCII expansion -> MurmurHash64A (tools.hpp:16-50) -> exact % length (kmodel.hpp:598) ->
//   cell word pos >> 5, bit (pos & 31) ^ 7, tag in the high half (kmodel.hpp:576-588,603-618) / Bloom word + mask.
// The large shapes (NA12878: bit_array_length ~ 1.1e10 > 2^32) reach positions no small test database does; this entry
// point drives the same device functions with an arbitrary length d so that a test can check them against the oracle's
// `hash % d` and against the population count of what was set (an address truncated to 32 bits would alias).
#include <cuda_runtime.h>
#include "kmx_internal.h"

namespace kmx {

__global__ void selftest_set_kernel(const uint64_t* __restrict__ kmers, size_t n, int k, FastMod mod, const uint32_t* __restrict__ seeds, int n_seeds,
                                    unsigned long long* __restrict__ cells, uint32_t* __restrict__ filter, uint64_t* __restrict__ pos_out) {
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		HashPrep p;
		hash_prepare(reverse_bases(kmers[i], k), k, p);
		for (int j = 0; j < n_seeds; j++) {
			const uint64_t pos = fastmod(hash_finish(p, k, seeds[j]), mod);
			pos_out[i * n_seeds + j] = pos;
			const uint32_t sh = ((uint32_t)pos & 31u) ^ 7u;
			red_or64(cells + (pos >> 5), ((1ULL << 32) | (unsigned long long)(j & 1)) << sh);     // tag + value, as insert_kernel's commit
			red_or32(filter + (pos >> 5), bit_mask32(pos));                                        // as filter_set
		}
	}
}

__global__ void selftest_check_kernel(const uint64_t* __restrict__ pos, size_t n_pos, const unsigned long long* __restrict__ cells,
                                      const uint32_t* __restrict__ filter, unsigned long long* __restrict__ found) {
	unsigned long long ok = 0;
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_pos; i += (size_t)gridDim.x * blockDim.x) {
		const uint64_t p = pos[i];
		const uint32_t sh = ((uint32_t)p & 31u) ^ 7u;
		const bool tag = ((uint32_t)(cells[p >> 5] >> 32) >> sh) & 1u;
		const bool bit = (filter[p >> 5] & bit_mask32(p)) != 0;
		ok += (tag && bit) ? 1 : 0;
	}
	atomicAdd(found, ok);
}

__global__ void selftest_popc_kernel(const unsigned long long* __restrict__ cells, const uint32_t* __restrict__ filter, uint64_t n_words,
                                     unsigned long long* __restrict__ tags, unsigned long long* __restrict__ bits) {
	unsigned long long a = 0, b = 0;
	for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (uint64_t)gridDim.x * blockDim.x) {
		a += __popc((uint32_t)(cells[w] >> 32));
		b += __popc(filter[w]);
	}
	atomicAdd(tags, a);
	atomicAdd(bits, b);
}

}  // namespace kmx

using namespace kmx;

// kmers[n] (host, packed) are hashed with seeds[n_seeds] modulo d on the device; pos_out[n * n_seeds] (host) receives the
// positions; counts[0] = positions found set again through the cell + filter addressing, counts[1] = tag bits set in the whole
// array, counts[2] = filter bits set in the whole array (both must equal the number of distinct positions)
extern "C" int kmx_selftest_positions(const uint64_t* kmers, size_t n, int k, uint64_t d, const uint32_t* seeds, int n_seeds, uint64_t* pos_out,
                                      uint64_t counts[3]) {
	if (!kmers || !seeds || !pos_out || !counts || n == 0 || k < 1 || k > 32 || d < 64 || n_seeds < 1 || n_seeds > 128)
		return set_error(KMX_EARG, "kmx_selftest_positions: bad argument");
	int sm = 0;
	int rc = require_gpu(&sm);
	if (rc) return rc;
	const uint64_t n_words = (d + 31) / 32 + 1;
	uint64_t* d_k = nullptr;
	uint32_t* d_s = nullptr;
	uint64_t* d_pos = nullptr;
	unsigned long long* d_cells = nullptr;
	uint32_t* d_filter = nullptr;
	unsigned long long* d_cnt = nullptr;
	DevScope scope(nullptr);
	if ((rc = scope.alloc(&d_k, n * 8)) || (rc = scope.alloc(&d_s, (size_t)n_seeds * 4)) || (rc = scope.alloc(&d_pos, n * n_seeds * 8)) ||
	    (rc = scope.alloc(&d_cells, n_words * 8)) || (rc = scope.alloc(&d_filter, n_words * 4)) || (rc = scope.alloc(&d_cnt, 24)))
		return rc;
	CU(cudaMemcpy(d_k, kmers, n * 8, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(d_s, seeds, (size_t)n_seeds * 4, cudaMemcpyHostToDevice));
	CU(cudaMemset(d_cells, 0, n_words * 8));
	CU(cudaMemset(d_filter, 0, n_words * 4));
	CU(cudaMemset(d_cnt, 0, 24));
	selftest_set_kernel<<<sm * 4, 256>>>(d_k, n, k, make_fastmod(d), d_s, n_seeds, d_cells, d_filter, d_pos);
	selftest_check_kernel<<<sm * 4, 256>>>(d_pos, n * n_seeds, d_cells, d_filter, d_cnt);
	selftest_popc_kernel<<<sm * 8, 256>>>(d_cells, d_filter, n_words, d_cnt + 1, d_cnt + 2);
	note_launch(3);
	CU(cudaGetLastError());
	CU(cudaDeviceSynchronize());
	CU(cudaMemcpy(pos_out, d_pos, n * n_seeds * 8, cudaMemcpyDeviceToHost));
	unsigned long long h[3];
	CU(cudaMemcpy(h, d_cnt, 24, cudaMemcpyDeviceToHost));
	for (int i = 0; i < 3; i++) counts[i] = h[i];
	return KMX_OK;
}

// ---- position-sensitive 64-bit checksum of a device region: lets the ranks of a team build compare their replicas of a
// large model without writing 15 GB of files each (rank 0's files are compared with the reference's, byte for byte) ------
namespace kmx {
__global__ void checksum_kernel(const unsigned long long* __restrict__ words, uint64_t n, unsigned long long salt, unsigned long long* __restrict__ out) {
	unsigned long long acc = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		unsigned long long z = words[i] + (i + salt) * 0x9E3779B97F4A7C15ULL;
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
		acc += z ^ (z >> 31);
	}
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, d);
	if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}
}  // namespace kmx

// sums[0] Bloom filters + km_back, [1] coupled arrays (device layout), [2] rest keys, [3] rest counts + group index
extern "C" int kmx_model_checksum(kmx_model* m, uint64_t sums[4]) {
	if (!m || !sums) return set_error(KMX_EARG, "null argument");
	if (!m->built) return set_error(KMX_ESTATE, "model is not initialised");
	CU(cudaSetDevice(m->device));
	cudaStream_t s = m->x->stream;
	unsigned long long* d_out = nullptr;
	DevScope scope(s);
	int rc = scope.alloc(&d_out, 32);
	if (rc) return rc;
	CU(cudaMemsetAsync(d_out, 0, 32, s));
	const int grid = m->sm_count * 8;
	auto add = [&](const void* p, uint64_t bytes, int slot, unsigned long long salt) {
		if (!p || bytes < 8) return;
		checksum_kernel<<<grid, 256, 0, s>>>((const unsigned long long*)p, bytes / 8, salt << 40, d_out + slot);
		note_launch();
	};
	for (int i = 0; i < m->bf_num; i++) {
		add(m->d_bf[i], m->bytes[i] & ~7ULL, 0, 1 + i);
		add(m->d_bf_back[i], m->bytes[3 + i] & ~7ULL, 0, 4 + i);
	}
	add(m->d_km_back, m->bytes[7] & ~7ULL, 0, 7);
	for (int i = 0; i < m->n_bits; i++) add(m->d_cells[i], cell_words(m->bytes[6]) * 8, 1, 8 + i);
	add(m->d_rest_keys, m->rest.count * 8, 2, 16);
	add(m->d_rest_counts, (m->rest.count * 4) & ~7ULL, 3, 17);
	add(m->d_hash2index, (uint64_t)m->rest.map_size * 4, 3, 18);
	add(m->d_pre_buffer, ((uint64_t)m->rest.pre_buffer_size * 4) & ~7ULL, 3, 19);
	CU(cudaGetLastError());
	unsigned long long h[4];
	CU(cudaMemcpyAsync(h, d_out, 32, cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));
	for (int i = 0; i < 4; i++) sums[i] = h[i];
	return KMX_OK;
}
