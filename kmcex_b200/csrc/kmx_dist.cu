// kmx_dist.cu -- the exchanges of a team build (ONE model built by several GPUs of one node), each as ONE kernel that
// computes and moves data through peer-mapped memory (NVLink / NVSwitch) with flag barriers; no NCCL, no host round trip.
//
//   or_allreduce_kernel   cross-GPU barrier, reduce-scatter (rank r ORs slice r of every rank's copy), all-gather (rank r
//                         stores the result into slice r of every copy), then this rank pulls the coupled arrays it does not
//                         own from their owners, cross-GPU barrier
//   rest_gather_kernel    sharded rest build: this rank collects, from every owner's survivor list, the k-mers whose 7-base
//                         prefix falls in its range
//   rest_push_kernel      this rank's sorted run goes into every rank's rest table; cross-GPU barrier
//
// What they move (file:line relative to the reference root):
//   kmodel.hpp:473-506  bit_bf / bit_bf_back: the Bloom inserts are order-free ORs, so the record range of the database is
//                       split over the ranks and the partial filters are OR-ed
//   kmodel.hpp:546-550  km_back: every array owner ORs the (k-2)-mers of the items it accepted
//   kmodel.hpp:32-37    bit_array_1 / bit_array_2 of every pair: built by the owner, replicated for the query
//   rest.hpp:95-135     stat / sort_suffix / transform: survivors grouped by 7-base prefix, each group sorted
// NCCL has no OR reduction; an all-gather + local OR would move world x the bytes and need world x the memory.  Here every
// rank reads (world-1)/world of the region from its peers and writes as much.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "kmx_launch.h"

namespace cg = cooperative_groups;

namespace kmx {

// flags[p][r] = number of barriers rank r has entered, as visible on rank p
__device__ __forceinline__ bool cross_gpu_barrier(cg::grid_group& grid, const TeamLink& a, uint32_t seq) {
	__threadfence_system();
	grid.sync();
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		__threadfence_system();
		for (int p = 0; p < a.world; p++) ((volatile uint32_t*)a.flags[p])[a.rank] = seq;
		unsigned long long t0, t1;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
		for (int q = 0; q < a.world; q++) {
			while ((int)(((volatile uint32_t*)a.flags[a.rank])[q] - seq) < 0) {
				asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
				if (t1 - t0 > 20000000000ULL) {                  // 20 s: a peer died; give up instead of hanging the GPU
					*(volatile unsigned int*)a.error = 3;
					break;
				}
			}
		}
		__threadfence_system();
	}
	grid.sync();
	return *(volatile unsigned int*)a.error == 0;
}

__global__ void __launch_bounds__(256) or_allreduce_kernel(const __grid_constant__ OrReduceArgs a) {
	cg::grid_group grid = cg::this_grid();
	if (!cross_gpu_barrier(grid, a.link, a.link.seq + 1)) return;      // every rank's partial region (and owned array) is complete
	const int rank = a.link.rank, world = a.link.world;
	const unsigned long long lo = a.n_vec * (unsigned long long)rank / (unsigned long long)world;
	const unsigned long long hi = a.n_vec * (unsigned long long)(rank + 1) / (unsigned long long)world;
	const unsigned long long T = (unsigned long long)gridDim.x * blockDim.x;
	const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
	for (unsigned long long v = lo + tid; v < hi; v += T) {
		uint4 x[kMaxRanks];
#pragma unroll
		for (int p = 0; p < kMaxRanks; p++)
			if (p < world) x[p] = __ldcv(a.base[p] + v);     // all loads in flight before the first use
		uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
		for (int p = 0; p < kMaxRanks; p++) {
			if (p < world) {
				acc.x |= x[p].x; acc.y |= x[p].y; acc.z |= x[p].z; acc.w |= x[p].w;
			}
		}
#pragma unroll
		for (int p = 0; p < kMaxRanks; p++)
			if (p < world) a.base[p][v] = acc;
	}
	// the coupled arrays this rank does not own: streamed from their owners, four 16-byte loads in flight per thread
	for (int s = 0; s < a.n_pull; s++) {
		const uint4* src = a.pull[s].src;
		uint4* dst = a.pull[s].dst;
		const unsigned long long n = a.pull[s].n_vec;
		for (unsigned long long v = tid; v < n; v += 4 * T) {
			uint4 x[4];
#pragma unroll
			for (int q = 0; q < 4; q++)
				if (v + q * T < n) x[q] = __ldcv(src + v + q * T);
#pragma unroll
			for (int q = 0; q < 4; q++)
				if (v + q * T < n) dst[v + q * T] = x[q];
		}
	}
	cross_gpu_barrier(grid, a.link, a.link.seq + 2);        // every slice of every copy is written, every pull has been served
}

static cudaError_t coop_grid(const void* kernel, int sm_count, int cap_per_sm, int* blocks) {
	int per_sm = 0;
	cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0);
	if (e != cudaSuccess) return e;
	if (per_sm < 1) return cudaErrorLaunchOutOfResources;
	if (per_sm > cap_per_sm) per_sm = cap_per_sm;
	*blocks = per_sm * sm_count;
	return cudaSuccess;
}

cudaError_t launch_or_allreduce(const OrReduceArgs& a, int sm_count, cudaStream_t stream) {
	int blocks = 0;
	cudaError_t e = coop_grid((const void*)or_allreduce_kernel, sm_count, 4, &blocks);
	if (e != cudaSuccess) return e;
	void* args[1] = { (void*)&a };
	note_launch();
	return cudaLaunchCooperativeKernel((const void*)or_allreduce_kernel, dim3(blocks), dim3(256), args, 0, stream);
}

// ---- rest table, sharded by prefix range -------------------------------------------------------
__global__ void __launch_bounds__(256) prefix_hist_kernel(const uint64_t* __restrict__ keys, const unsigned long long* __restrict__ n_ptr,
                                                          int suffix_bits, uint32_t* __restrict__ hist) {
	const unsigned long long n = *n_ptr;
	for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x)
		atomicAdd(hist + (uint32_t)(keys[i] >> suffix_bits), 1u);
}

cudaError_t launch_prefix_hist(const uint64_t* d_keys, const unsigned long long* d_n, int suffix_bits, uint32_t* d_hist, int sm_count,
                               cudaStream_t stream) {
	prefix_hist_kernel<<<sm_count * 4, 256, 0, stream>>>(d_keys, d_n, suffix_bits, d_hist);
	note_launch();
	return cudaGetLastError();
}

__global__ void __launch_bounds__(256) rest_gather_kernel(const __grid_constant__ RestGatherArgs a) {
	const unsigned long long T = (unsigned long long)gridDim.x * blockDim.x;
	const int lane = threadIdx.x & 31;
	for (int o = 0; o < a.n_owners; o++) {
		const uint64_t* keys = a.kmer[o];
		const uint32_t* occ = a.occ[o];
		const unsigned long long n = a.n[o];
		const unsigned long long n_round = (n + 31) & ~31ULL;              // whole warps: the ballot below needs every lane
		for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += T) {
			const uint64_t key = i < n ? __ldcv(keys + i) : 0;
			const uint32_t pre = (uint32_t)(key >> a.suffix_bits);
			const bool mine = i < n && pre >= a.prefix_lo && pre < a.prefix_hi;
			const uint32_t ball = __ballot_sync(0xffffffffu, mine);
			if (ball == 0) continue;
			unsigned long long at = 0;
			if (lane == 0) at = atomicAdd(a.out_n, (unsigned long long)__popc(ball));
			at = __shfl_sync(0xffffffffu, at, 0) + (unsigned long long)__popc(ball & ((1u << lane) - 1u));
			if (mine && at < a.cap) {
				a.out_kmer[at] = key;
				a.out_occ[at] = __ldcv(occ + i);
			}
		}
	}
}

cudaError_t launch_rest_gather(const RestGatherArgs& a, int sm_count, cudaStream_t stream) {
	cudaError_t e = cudaMemsetAsync(a.out_n, 0, sizeof(unsigned long long), stream);
	if (e != cudaSuccess) return e;
	rest_gather_kernel<<<sm_count * 8, 256, 0, stream>>>(a);
	note_launch();
	return cudaGetLastError();
}

__global__ void __launch_bounds__(256) rest_push_kernel(const __grid_constant__ RestPushArgs a) {
	cg::grid_group grid = cg::this_grid();
	const unsigned long long T = (unsigned long long)gridDim.x * blockDim.x;
	for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += T) {
		const uint64_t key = a.keys[i];
		const int32_t cnt = (int32_t)a.counts[i];
#pragma unroll
		for (int p = 0; p < kMaxRanks; p++) {
			if (p < a.link.world) {
				a.dst_keys[p][a.offset + i] = key;
				a.dst_counts[p][a.offset + i] = cnt;
			}
		}
	}
	cross_gpu_barrier(grid, a.link, a.link.seq + 1);        // every rank's run is in every table; nobody reads a survivor list any more
}

cudaError_t launch_rest_push(const RestPushArgs& a, int sm_count, cudaStream_t stream) {
	int blocks = 0;
	cudaError_t e = coop_grid((const void*)rest_push_kernel, sm_count, 4, &blocks);
	if (e != cudaSuccess) return e;
	void* args[1] = { (void*)&a };
	note_launch();
	return cudaLaunchCooperativeKernel((const void*)rest_push_kernel, dim3(blocks), dim3(256), args, 0, stream);
}

}  // namespace kmx
