// kmx_gridbar.cuh -- grid-wide barrier of a cooperative (co-resident) kernel on one monotonically increasing counter.
// Thread 0 of every block adds 1 and waits until the counter reaches the next multiple of the grid size; the counter is
// never reset, so the barrier state survives across launches of the same grid size... and across grid sizes as long as
// init() runs (and is followed by one cooperative_groups grid.sync()) at the start of every kernel.
#pragma once
#include <stdint.h>

namespace kmx {

struct GridBarrier {
	unsigned int* counter;
	unsigned int target;          // counter value that completes the next barrier (meaningful in thread 0 of a block)

	// every block reads the counter before anybody arrives: call, then grid.sync() once, then use sync()
	__device__ __forceinline__ void init(unsigned int* c) {
		counter = c;
		target = *(volatile unsigned int*)c + gridDim.x;
	}

	// arrive + wait without the trailing block barrier: thread 0 of the block returns when every block has arrived, the
	// other threads return at once -- the caller follows up with work of thread 0 (e.g. fetching control words into
	// shared memory) and a __syncthreads()
	__device__ __forceinline__ void arrive_wait() {
		__syncthreads();
		if (threadIdx.x == 0) {
			unsigned int seen;
			asm volatile("atom.add.release.gpu.u32 %0, [%1], 1;" : "=r"(seen) : "l"(counter) : "memory");
			seen += 1;
			while ((int)(seen - target) < 0)                     // relaxed polls: an acquire load invalidates the L1 every time round
				asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
			asm volatile("fence.acq_rel.gpu;" ::: "memory");
			target += gridDim.x;
		}
	}

	__device__ __forceinline__ void sync() {
		__syncthreads();
		if (threadIdx.x == 0) {
			unsigned int seen;
			asm volatile("atom.add.release.gpu.u32 %0, [%1], 1;" : "=r"(seen) : "l"(counter) : "memory");
			seen += 1;
			while ((int)(seen - target) < 0)                     // relaxed polls: an acquire load invalidates the L1 every time round
				asm volatile("ld.relaxed.gpu.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
			asm volatile("fence.acq_rel.gpu;" ::: "memory");
			target += gridDim.x;
		}
		__syncthreads();
	}
};

}  // namespace kmx
