// kmx_microbench.cu -- random-sector microbenchmarks: the roofline denominators of this path.
// The build and the query are bound by random 32-byte-sector traffic, not
// by streaming bandwidth; these kernels measure what the chip sustains for the three access
// kinds the path uses, at a given footprint (L2-resident or HBM-resident):
//   kind 0: independent random 8-byte loads        (coupled-array probes, ld.global.nc)
//   kind 1: random 32-bit atomic OR, no return     (Bloom / km_back inserts, red.global.or.b32)
//   kind 2: random 64-bit atomic OR, no return     (coupled-array commits)
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <stdint.h>
#include "../../include/kmx.h"
#include "kmx_gridbar.cuh"

namespace {

__device__ __forceinline__ uint64_t mix(uint64_t z) {
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
	return z ^ (z >> 31);
}

// window_words != 0: the accesses of item i fall into window (i * n_windows / n_items) of the buffer, i.e. threads that
// run together touch the same window_words * 8 bytes (what binning the probes by address range would give)
template <int KIND>
__global__ void __launch_bounds__(256) random_sector_kernel(unsigned long long* buf, uint64_t n_words, uint64_t n_items, unsigned long long* sink,
                                                            uint64_t window_words) {
	constexpr int PER = 7;                       // independent accesses per item, as one array probe issues
	unsigned long long acc = 0;
	const uint64_t n_windows = window_words ? n_words / window_words : 1;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t h = mix(i * 0x9E3779B97F4A7C15ULL + 12345);
		const uint64_t wbase = window_words ? (uint64_t)(((unsigned __int128)i * n_windows) / n_items) * window_words : 0;
		unsigned long long v[PER];
#pragma unroll
		for (int j = 0; j < PER; j++) {
			h = mix(h + j);
			const uint64_t w = window_words ? wbase + h % window_words : h % n_words;
			if (KIND == 0) v[j] = __ldg(buf + w);
			else if (KIND == 1) atomicOr((unsigned int*)buf + 2 * w + (j & 1), 1u << (h >> 59));
			else atomicOr(buf + w, 1ULL << (h >> 58));
		}
		if (KIND == 0) {
#pragma unroll
			for (int j = 0; j < PER; j++) acc ^= v[j];
		}
	}
	if (KIND == 0 && acc == 0x1234567ULL) *sink = acc;   // keep the loads alive
}

}  // namespace

// footprint_bytes of device memory are touched at random; n_items * 7 accesses per launch;
// returns the average milliseconds per launch over `reps` launches (after one warm-up) in *ms_out
extern "C" int kmx_microbench_windowed(int kind, uint64_t footprint_bytes, uint64_t window_bytes, uint64_t n_items, int reps, float* ms_out);
extern "C" int kmx_microbench_random(int kind, uint64_t footprint_bytes, uint64_t n_items, int reps, float* ms_out) {
	return kmx_microbench_windowed(kind, footprint_bytes, 0, n_items, reps, ms_out);
}

extern "C" int kmx_microbench_windowed(int kind, uint64_t footprint_bytes, uint64_t window_bytes, uint64_t n_items, int reps, float* ms_out) {
	if (kind < 0 || kind > 2 || footprint_bytes < 4096 || !ms_out || reps < 1 || (window_bytes && window_bytes > footprint_bytes)) return KMX_EARG;
	const uint64_t window_words = window_bytes / 8;
	unsigned long long* buf = nullptr;
	unsigned long long* sink = nullptr;
	if (cudaMalloc(&buf, footprint_bytes) != cudaSuccess || cudaMalloc(&sink, 8) != cudaSuccess) {
		cudaGetLastError();
		cudaFree(buf);
		return KMX_ECUDA;
	}
	cudaMemset(buf, 0, footprint_bytes);
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const uint64_t n_words = footprint_bytes / 8;
	const int grid = sms * 8;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	for (int r = -1; r < reps; r++) {
		if (r == 0) cudaEventRecord(e0);
		if (kind == 0) random_sector_kernel<0><<<grid, 256>>>(buf, n_words, n_items, sink, window_words);
		else if (kind == 1) random_sector_kernel<1><<<grid, 256>>>(buf, n_words, n_items, sink, window_words);
		else random_sector_kernel<2><<<grid, 256>>>(buf, n_words, n_items, sink, window_words);
	}
	cudaEventRecord(e1);
	cudaError_t e = cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	*ms_out = ms / reps;
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(buf);
	cudaFree(sink);
	return e == cudaSuccess ? KMX_OK : KMX_ECUDA;
}

// ---- grid-barrier cost: the persistent insert kernel pays ~10 grid barriers per round ------------------------------
namespace {
template <int MODE>
__global__ void grid_barrier_kernel(int reps, unsigned int* bar, unsigned long long* sink) {
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	kmx::GridBarrier gb;
	gb.init(bar);
	grid.sync();
	unsigned long long acc = 0;
	for (int r = 0; r < reps; r++) {
		acc += threadIdx.x + r;
		if (MODE == 0) grid.sync();
		else gb.sync();
	}
	if (acc == 0x1234567ULL) *sink = acc;
}
}  // namespace

// mode 0: cooperative_groups grid.sync(); mode 1: the counter barrier of kmx_gridbar.cuh.  grid = blocks_per_sm * SMs
// blocks of `threads` threads; *us_out = microseconds per barrier
extern "C" int kmx_microbench_grid_barrier(int mode, int threads, int blocks_per_sm, int reps, float* us_out) {
	if (mode < 0 || mode > 1 || threads < 32 || threads > 1024 || blocks_per_sm < 1 || reps < 1 || !us_out) return KMX_EARG;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	unsigned int* bar = nullptr;
	unsigned long long* sink = nullptr;
	if (cudaMalloc(&bar, 256) != cudaSuccess || cudaMalloc(&sink, 8) != cudaSuccess) return KMX_ECUDA;
	cudaMemset(bar, 0, 256);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	void* args[3] = { (void*)&reps, (void*)&bar, (void*)&sink };
	const void* fn = mode == 0 ? (const void*)grid_barrier_kernel<0> : (const void*)grid_barrier_kernel<1>;
	cudaError_t e = cudaSuccess;
	for (int r = -1; r < 1 && e == cudaSuccess; r++) {
		if (r == 0) cudaEventRecord(e0);
		e = cudaLaunchCooperativeKernel(fn, dim3(blocks_per_sm * sms), dim3(threads), args, 0, 0);
	}
	cudaEventRecord(e1);
	if (e == cudaSuccess) e = cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	*us_out = 1e3f * ms / reps;
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(bar);
	cudaFree(sink);
	return e == cudaSuccess ? KMX_OK : KMX_ECUDA;
}

// ---- returning atomics on few addresses: the list appends of the insert kernel (one atomicAdd per warp and iteration) ---
namespace {
__global__ void __launch_bounds__(256) hot_atomic_kernel(unsigned int* counters, int n_counters, int per_warp, unsigned int* sink) {
	const unsigned int warp_g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	unsigned int acc = 0;
	if ((threadIdx.x & 31) == 0)
		for (int r = 0; r < per_warp; r++) acc += atomicAdd(counters + 32 * ((warp_g + r) % n_counters), 1u);   // one counter per 128-byte line
	if (acc == 0x12345u) *sink = acc;
}
}  // namespace

// every warp of a full grid (4 blocks of 256 threads per SM) issues per_warp dependent atomicAdd-with-result on one of
// n_counters addresses; *ns_out = nanoseconds per atomic (wall time of the kernel / atomics issued)
extern "C" int kmx_microbench_hot_atomic(int n_counters, int per_warp, float* ns_out) {
	if (n_counters < 1 || n_counters > 1024 || per_warp < 1 || !ns_out) return KMX_EARG;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	unsigned int* counters = nullptr;
	unsigned int* sink = nullptr;
	if (cudaMalloc(&counters, 1024 * 128) != cudaSuccess || cudaMalloc(&sink, 4) != cudaSuccess) return KMX_ECUDA;
	cudaMemset(counters, 0, 1024 * 128);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	const int grid = 4 * sms;
	hot_atomic_kernel<<<grid, 256>>>(counters, n_counters, per_warp, sink);
	cudaEventRecord(e0);
	hot_atomic_kernel<<<grid, 256>>>(counters, n_counters, per_warp, sink);
	cudaEventRecord(e1);
	cudaError_t e = cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	*ns_out = 1e6f * ms / ((float)grid * 8 * per_warp);
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(counters);
	cudaFree(sink);
	return e == cudaSuccess ? KMX_OK : KMX_ECUDA;
}

// ---- read-only streaming ceiling: what a pure HBM read kernel reaches (the denominator of the counting pass) --------------
namespace {
__global__ void __launch_bounds__(256) stream_read_kernel(const uint4* __restrict__ buf, uint64_t n_vec, unsigned long long* sink) {
	const uint64_t T = (uint64_t)gridDim.x * blockDim.x;
	uint32_t acc = 0;
	uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	for (; v + 3 * T < n_vec; v += 4 * T) {                     // four independent 16-byte loads in flight per thread
		const uint4 a = __ldg(buf + v), b = __ldg(buf + v + T), c = __ldg(buf + v + 2 * T), d = __ldg(buf + v + 3 * T);
		acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
	}
	for (; v < n_vec; v += T) {
		const uint4 a = __ldg(buf + v);
		acc ^= a.x ^ a.y ^ a.z ^ a.w;
	}
	if (acc == 0x12345677u) *sink = acc;
}
}  // namespace

// milliseconds per pass of a read-only grid-stride kernel over `bytes` of device memory (16-byte loads, 4 in flight per thread)
extern "C" int kmx_microbench_stream_read(uint64_t bytes, int blocks_per_sm, int reps, float* ms_out) {
	if (bytes < 4096 || !ms_out || reps < 1 || blocks_per_sm < 1 || blocks_per_sm > 8) return KMX_EARG;
	uint4* buf = nullptr;
	unsigned long long* sink = nullptr;
	if (cudaMalloc(&buf, bytes) != cudaSuccess || cudaMalloc(&sink, 8) != cudaSuccess) {
		cudaGetLastError();
		cudaFree(buf);
		return KMX_ECUDA;
	}
	cudaMemset(buf, 1, bytes);
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	for (int r = -1; r < reps; r++) {
		if (r == 0) cudaEventRecord(e0);
		stream_read_kernel<<<sms * blocks_per_sm, 256>>>(buf, bytes / 16, sink);
	}
	cudaEventRecord(e1);
	cudaError_t e = cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	*ms_out = ms / reps;
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(buf);
	cudaFree(sink);
	return e == cudaSuccess ? KMX_OK : KMX_ECUDA;
}

// ---- owner-routed bits over NVLink: what a remote fire-and-forget reduction costs -------------------------------------
// kind as kmx_microbench_random; the kernel runs on device `src`, the buffer lives on device `dst` (peer access).  The number
// behind the team build's choice to OR-reduce replicated filters instead of routing every bit to an address-range owner.
extern "C" int kmx_microbench_peer_random(int kind, int src, int dst, uint64_t footprint_bytes, uint64_t n_items, int reps, float* ms_out) {
	if (kind < 0 || kind > 2 || footprint_bytes < 4096 || !ms_out || reps < 1 || src == dst) return KMX_EARG;
	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || src < 0 || dst < 0 || src >= n_dev || dst >= n_dev) {
		cudaGetLastError();
		return KMX_ENOGPU;
	}
	unsigned long long* buf = nullptr;
	unsigned long long* sink = nullptr;
	cudaSetDevice(dst);
	if (cudaMalloc(&buf, footprint_bytes) != cudaSuccess) {
		cudaGetLastError();
		return KMX_ECUDA;
	}
	cudaMemset(buf, 0, footprint_bytes);
	cudaDeviceSynchronize();
	cudaSetDevice(src);
	cudaError_t pe = cudaDeviceEnablePeerAccess(dst, 0);
	if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) {
		cudaGetLastError();
		cudaSetDevice(dst);
		cudaFree(buf);
		return KMX_ECUDA;
	}
	cudaGetLastError();
	cudaMalloc(&sink, 8);
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, src);
	const uint64_t n_words = footprint_bytes / 8;
	const int grid = sms * 8;
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	for (int r = -1; r < reps; r++) {
		if (r == 0) cudaEventRecord(e0);
		if (kind == 0) random_sector_kernel<0><<<grid, 256>>>(buf, n_words, n_items, sink, 0);
		else if (kind == 1) random_sector_kernel<1><<<grid, 256>>>(buf, n_words, n_items, sink, 0);
		else random_sector_kernel<2><<<grid, 256>>>(buf, n_words, n_items, sink, 0);
	}
	cudaEventRecord(e1);
	cudaError_t e = cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	*ms_out = ms / reps;
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(sink);
	cudaSetDevice(dst);
	cudaFree(buf);
	cudaSetDevice(src);
	return e == cudaSuccess ? KMX_OK : KMX_ECUDA;
}
