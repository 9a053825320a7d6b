// kmx_sort.cu -- the library's own sort of (packed k-mer, count) pairs: least-significant-digit radix sort, 8-bit digits.
//
// What it is for (file:line relative to the reference root): KRestData::build = stat + sort_suffix + transform
// (rest.hpp:95-135,157-161) leaves every survivor of the greedy insert sorted by its 2-bit packed value (the groups of the
// 7-base prefix index are contiguous ranges of that order, each group sorted by suffix).  Keys are unique except for the
// stale-slot duplicates of kmodel.hpp:520-540, which are identical pairs, so any stable or unstable order of equal keys
// gives the same bytes.
//
// One pass = three kernels over chunks of 16 Ki keys:
//   radix_hist_kernel     per-chunk histogram of the pass's digit           -> hist[digit][chunk]
//   radix_scan_kernel     one block per digit: exclusive scan over the chunks, row total -> total[digit]
//   radix_scatter_kernel  a block walks its chunk in sub-tiles of 2048 keys: warp-level ranks by MATCH.ANY (stable),
//                         the sub-tile is regrouped by digit in shared memory and written out in runs (coalesced per digit)
// Traffic per pass: keys + values read twice, written once (32 bytes per pair); HBM-stream bound.
#include <cuda_runtime.h>
#include "kmx_launch.h"

namespace kmx {

constexpr int kSortThreads = 256;
constexpr int kSortPer = 8;                                  // keys per thread and sub-tile
constexpr int kSortSub = kSortThreads * kSortPer;            // 2048 keys per sub-tile
constexpr int kSortSubsPerChunk = 8;
constexpr int kSortChunk = kSortSub * kSortSubsPerChunk;     // 16384 keys per block
constexpr int kRadix = 256;

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint64_t* __restrict__ keys, size_t n, int shift, uint32_t n_chunks,
                                                                  uint32_t* __restrict__ hist) {
	__shared__ uint32_t s_h[kRadix];
	s_h[threadIdx.x] = 0;
	__syncthreads();
	const size_t base = (size_t)blockIdx.x * kSortChunk;
	const size_t end = base + kSortChunk < n ? base + kSortChunk : n;
	for (size_t i = base + threadIdx.x; i < end; i += kSortThreads) atomicAdd(&s_h[(uint32_t)(keys[i] >> shift) & (kRadix - 1)], 1u);
	__syncthreads();
	hist[(size_t)threadIdx.x * n_chunks + blockIdx.x] = s_h[threadIdx.x];
}

// block d: hist[d][0 .. n_chunks) -> exclusive scan in place, total[d] = row sum
__global__ void __launch_bounds__(kSortThreads) radix_scan_kernel(uint32_t* __restrict__ hist, uint32_t n_chunks, uint32_t* __restrict__ total) {
	__shared__ uint32_t s_warp[kSortThreads / 32];
	__shared__ uint32_t s_carry;
	uint32_t* row = hist + (size_t)blockIdx.x * n_chunks;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_carry = 0;
	__syncthreads();
	for (uint32_t b = 0; b < n_chunks; b += kSortThreads) {
		const uint32_t i = b + threadIdx.x;
		const uint32_t x = i < n_chunks ? row[i] : 0u;
		uint32_t incl = x;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
			if (lane >= d) incl += y;
		}
		if (lane == 31) s_warp[warp] = incl;
		__syncthreads();
		uint32_t before = 0, all = 0;
#pragma unroll
		for (int w = 0; w < kSortThreads / 32; w++) {
			const uint32_t v = s_warp[w];
			before += w < warp ? v : 0u;
			all += v;
		}
		const uint32_t carry = s_carry;
		if (i < n_chunks) row[i] = carry + before + incl - x;
		__syncthreads();
		if (threadIdx.x == 0) s_carry = carry + all;
		__syncthreads();
	}
	if (threadIdx.x == 0) total[blockIdx.x] = s_carry;
}

template <bool VALS>
__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                     uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, size_t n,
                                                                     int shift, uint32_t n_chunks, const uint32_t* __restrict__ hist,
                                                                     const uint32_t* __restrict__ total) {
	__shared__ uint64_t s_keys[kSortSub];
	__shared__ uint32_t s_vals[kSortSub];
	__shared__ uint32_t s_wcnt[kSortThreads / 32][kRadix];   // per warp: keys of each digit seen so far in the sub-tile
	__shared__ uint32_t s_cnt[kRadix], s_start[kRadix + 1];
	__shared__ unsigned long long s_base[kRadix];            // where the next key of each digit of this chunk goes
	__shared__ uint32_t s_scan[kSortThreads / 32];
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t lt_mask = (1u << lane) - 1u;
	{
		// digit base = keys of smaller digits (exclusive scan of the 256 row totals) + keys of this digit in earlier chunks
		const uint32_t t = total[threadIdx.x];
		uint32_t incl = t;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
			if (lane >= d) incl += y;
		}
		if (lane == 31) s_scan[warp] = incl;
		__syncthreads();
		unsigned long long before = 0;
#pragma unroll
		for (int w = 0; w < kSortThreads / 32; w++) before += w < warp ? s_scan[w] : 0u;
		s_base[threadIdx.x] = before + (incl - t) + hist[(size_t)threadIdx.x * n_chunks + blockIdx.x];
	}
	const size_t chunk_base = (size_t)blockIdx.x * kSortChunk;
	for (int sub = 0; sub < kSortSubsPerChunk; sub++) {
		const size_t sub_base = chunk_base + (size_t)sub * kSortSub;
		if (sub_base >= n) break;                               // uniform over the block
		const uint32_t n_sub = (uint32_t)(n - sub_base < (size_t)kSortSub ? n - sub_base : (size_t)kSortSub);
		__syncthreads();                                        // previous sub-tile: everybody is done with s_keys / s_wcnt / s_base
#pragma unroll
		for (int w = 0; w < kSortThreads / 32; w++) s_wcnt[w][threadIdx.x] = 0;
		__syncthreads();
		// a warp takes 256 consecutive keys, 32 at a time: the rank of a key among the keys of its digit follows key order
		uint64_t key[kSortPer];
		uint32_t val[kSortPer], dig[kSortPer], rnk[kSortPer];
#pragma unroll
		for (int j = 0; j < kSortPer; j++) {
			const uint32_t x = (uint32_t)warp * (kSortPer * 32) + (uint32_t)j * 32 + (uint32_t)lane;
			const bool valid = x < n_sub;
			key[j] = valid ? keys_in[sub_base + x] : 0;
			val[j] = (VALS && valid) ? vals_in[sub_base + x] : 0;
			dig[j] = valid ? ((uint32_t)(key[j] >> shift) & (kRadix - 1)) : (uint32_t)kRadix;   // 256 = not a key
			const uint32_t peers = __match_any_sync(0xffffffffu, dig[j]);
			const int leader = __ffs((int)peers) - 1;
			uint32_t old = 0;
			if (lane == leader && valid) {
				old = s_wcnt[warp][dig[j]];
				s_wcnt[warp][dig[j]] = old + (uint32_t)__popc(peers);
			}
			old = __shfl_sync(0xffffffffu, old, leader);
			rnk[j] = old + (uint32_t)__popc(peers & lt_mask);
			__syncwarp();
		}
		__syncthreads();
		{
			// thread d: keys of digit d in the warps before each warp; s_cnt[d] = keys of digit d in the sub-tile
			uint32_t acc = 0;
#pragma unroll
			for (int w = 0; w < kSortThreads / 32; w++) {
				const uint32_t t = s_wcnt[w][threadIdx.x];
				s_wcnt[w][threadIdx.x] = acc;
				acc += t;
			}
			s_cnt[threadIdx.x] = acc;
			uint32_t incl = acc;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				const uint32_t y = __shfl_up_sync(0xffffffffu, incl, d);
				if (lane >= d) incl += y;
			}
			if (lane == 31) s_scan[warp] = incl;
			__syncthreads();
			uint32_t before = 0;
#pragma unroll
			for (int w = 0; w < kSortThreads / 32; w++) before += w < warp ? s_scan[w] : 0u;
			s_start[threadIdx.x] = before + incl - acc;
			if (threadIdx.x == kRadix - 1) s_start[kRadix] = before + incl;
		}
		__syncthreads();
#pragma unroll
		for (int j = 0; j < kSortPer; j++) {
			if (dig[j] < (uint32_t)kRadix) {
				const uint32_t at = s_start[dig[j]] + s_wcnt[warp][dig[j]] + rnk[j];
				s_keys[at] = key[j];
				if (VALS) s_vals[at] = val[j];
			}
		}
		__syncthreads();
		for (uint32_t x = threadIdx.x; x < n_sub; x += kSortThreads) {
			const uint64_t kx = s_keys[x];
			const uint32_t d = (uint32_t)(kx >> shift) & (kRadix - 1);
			const unsigned long long dst = s_base[d] + (x - s_start[d]);
			keys_out[dst] = kx;
			if (VALS) vals_out[dst] = s_vals[x];
		}
		__syncthreads();
		s_base[threadIdx.x] += s_cnt[threadIdx.x];
	}
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static uint32_t sort_chunks(size_t n) { return (uint32_t)((n + kSortChunk - 1) / kSortChunk); }

// temp layout: [alternate keys | alternate values | hist[256][chunks] | total[256]]
size_t radix_sort_temp_bytes(size_t n) {
	if (n == 0) return 0;
	return align256(n * 8) + align256(n * 4) + align256((size_t)kRadix * sort_chunks(n) * 4) + align256(kRadix * 4);
}

// d_vals_in == nullptr: keys only (the k-mer counting stage)
cudaError_t launch_radix_sort_pairs(void* d_temp, const uint64_t* d_keys_in, uint64_t* d_keys_out, const uint32_t* d_vals_in,
                                    uint32_t* d_vals_out, size_t n, int key_bits, cudaStream_t stream) {
	if (n == 0) return cudaSuccess;
	if (n >= ((size_t)1 << 32)) return cudaErrorInvalidValue;   // 32-bit chunk histograms (the rest table's indices are int anyway, rest.hpp:66-70)
	const bool vals = d_vals_in != nullptr;
	uint8_t* t = (uint8_t*)d_temp;
	uint64_t* alt_keys = (uint64_t*)t;
	uint32_t* alt_vals = (uint32_t*)(t + align256(n * 8));
	const uint32_t n_chunks = sort_chunks(n);
	uint32_t* hist = (uint32_t*)(t + align256(n * 8) + align256(n * 4));
	uint32_t* total = (uint32_t*)((uint8_t*)hist + align256((size_t)kRadix * n_chunks * 4));
	const int n_pass = (key_bits + 7) / 8 > 0 ? (key_bits + 7) / 8 : 1;
	const uint64_t* src_k = d_keys_in;
	const uint32_t* src_v = d_vals_in;
	// the destinations alternate and the last one must be the caller's output
	bool to_out = (n_pass & 1) != 0;
	for (int p = 0; p < n_pass; p++) {
		uint64_t* dst_k = to_out ? d_keys_out : alt_keys;
		uint32_t* dst_v = to_out ? d_vals_out : alt_vals;
		const int shift = 8 * p;
		radix_hist_kernel<<<n_chunks, kSortThreads, 0, stream>>>(src_k, n, shift, n_chunks, hist);
		radix_scan_kernel<<<kRadix, kSortThreads, 0, stream>>>(hist, n_chunks, total);
		if (vals) radix_scatter_kernel<true><<<n_chunks, kSortThreads, 0, stream>>>(src_k, src_v, dst_k, dst_v, n, shift, n_chunks, hist, total);
		else radix_scatter_kernel<false><<<n_chunks, kSortThreads, 0, stream>>>(src_k, nullptr, dst_k, nullptr, n, shift, n_chunks, hist, total);
		note_launch(3);
		src_k = dst_k;
		src_v = dst_v;
		to_out = !to_out;
	}
	return cudaGetLastError();
}

}  // namespace kmx
