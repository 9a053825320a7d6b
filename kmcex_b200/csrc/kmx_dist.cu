// kmx_dist.cu -- bitwise-OR all-reduce of a filter region over the GPUs of one node, through
// peer-mapped memory (NVLink / NVSwitch), as ONE kernel: cross-GPU entry barrier, reduce-scatter
// (rank r ORs slice r of every rank's copy), all-gather (rank r stores the result into slice r of
// every rank's copy), cross-GPU exit barrier.
//
// What it merges (file:line relative to the reference root):
//   kmodel.hpp:473-506  bit_bf / bit_bf_back: the Bloom inserts are order-free ORs, so the record
//                       range of the database is split over the ranks and the partial filters are OR-ed
//   kmodel.hpp:546-550  km_back: every array owner ORs the (k-2)-mers of the items it accepted
// NCCL has no OR reduction; an all-gather + local OR moves world x the bytes and needs world x the
// memory.  Here every rank reads (world-1)/world of the region from its peers and writes as much.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "kmx_launch.h"

namespace cg = cooperative_groups;

namespace kmx {

// flags[p][r] = number of barriers rank r has entered, as visible on rank p
__device__ __forceinline__ bool cross_gpu_barrier(cg::grid_group& grid, const OrReduceArgs& a, uint32_t seq) {
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
	if (!cross_gpu_barrier(grid, a, a.seq + 1)) return;      // every rank's partial region is complete
	const unsigned long long lo = a.n_vec * (unsigned long long)a.rank / (unsigned long long)a.world;
	const unsigned long long hi = a.n_vec * (unsigned long long)(a.rank + 1) / (unsigned long long)a.world;
	const unsigned long long T = (unsigned long long)gridDim.x * blockDim.x;
	for (unsigned long long v = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; v < hi; v += T) {
		uint4 x[kMaxRanks];
#pragma unroll
		for (int p = 0; p < kMaxRanks; p++)
			if (p < a.world) x[p] = __ldcv(a.base[p] + v);     // all loads in flight before the first use
		uint4 acc = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
		for (int p = 0; p < kMaxRanks; p++) {
			if (p < a.world) {
				acc.x |= x[p].x; acc.y |= x[p].y; acc.z |= x[p].z; acc.w |= x[p].w;
			}
		}
#pragma unroll
		for (int p = 0; p < kMaxRanks; p++)
			if (p < a.world) a.base[p][v] = acc;
	}
	cross_gpu_barrier(grid, a, a.seq + 2);                  // every slice of every copy is written
}

cudaError_t launch_or_allreduce(const OrReduceArgs& a, int sm_count, cudaStream_t stream) {
	int per_sm = 0;
	cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, or_allreduce_kernel, 256, 0);
	if (e != cudaSuccess) return e;
	if (per_sm < 1) return cudaErrorLaunchOutOfResources;
	if (per_sm > 4) per_sm = 4;
	void* args[1] = { (void*)&a };
	return cudaLaunchCooperativeKernel((const void*)or_allreduce_kernel, dim3(per_sm * sm_count), dim3(256), args, 0, stream);
}

}  // namespace kmx
