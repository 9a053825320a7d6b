// kmx_build.cu -- the model build on the device: KMC listing decode, counting pass, Bloom
// inserts, the greedy coupled-array insert, the rest-table index.
//
// Reference path reproduced (file:line relative to the reference root):
//   kmc_file.cpp:428-515  CKMCFile::ReadNextKmer (record layout, LUT walk, count filter)
//   kmodel.hpp:423-434    get_km_kmer_count (pass 1)      kmodel.hpp:473-506  Bloom inserts
//   kmodel.hpp:508-527    batch formation                 kmodel.hpp:557-573  insert_with_thread (rounds)
//   kmodel.hpp:543-555    insert_array                    kmodel.hpp:590-622  insert_to_array
//   kmodel.hpp:529-540    reorder_buffer                  rest.hpp:95-135     stat / sort_suffix / transform
//
// The reference inserts the <= 2^18 items of a bucket ONE AFTER THE OTHER into one coupled
// array; the result depends on that order.  insert_kernel reproduces the sequential result in
// parallel with deterministic reservations:
//   every iteration, each undecided item (a) re-reads its n_hash cells; a conflict with the
//   committed state rejects it for good (bits are never cleared and the value under a set tag
//   never changes); (b) otherwise writes its index with atomicMin into a reservation table for
//   every position whose tag is still clear, keyed by (position, wanted value); (c) after a
//   grid barrier it is accepted iff no smaller-index undecided item wants the OPPOSITE value at
//   any of those positions.  Items accepted in one iteration cannot influence each other, the
//   smallest undecided index is always decided, and what an item sees at commit time is
//   exactly what the sequential loop would have shown it.  The reservation table is smaller
//   than the bit array (positions alias), which can only delay an acceptance, never change it.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "kmx_device.cuh"
#include "kmx_launch.h"
#include "kmx_gridbar.cuh"

// grid-wide barrier of the persistent insert kernel: 0 = cooperative_groups grid.sync(), 1 = the counter barrier of kmx_gridbar.cuh
#ifndef KMX_GRIDBAR
#define KMX_GRIDBAR 1
#endif

namespace cg = cooperative_groups;

namespace kmx {

// =========================================================================================
// block helpers (256 threads)
// =========================================================================================
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t x) {
	const int lane = threadIdx.x & 31;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
		if (lane >= d) x += y;
	}
	return x;
}

// exclusive scan of one value per thread over a block of NW warps; *total = block sum
template <int NW>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t x, uint32_t* s_warp /*[NW + 1]*/, uint32_t* total) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	uint32_t incl = warp_incl_scan(x);
	__syncthreads();                      // s_warp may still be read by a previous call
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		uint32_t w = lane < NW ? s_warp[lane] : 0;
		uint32_t wi = warp_incl_scan(w);
		if (lane < NW) s_warp[lane] = wi - w;
		if (lane == NW - 1) s_warp[NW] = wi;
	}
	__syncthreads();
	*total = s_warp[NW];
	return s_warp[warp] + incl - x;
}

__device__ __forceinline__ unsigned long long block_sum(unsigned long long x, unsigned long long* s_warp /*[8]*/) {
#pragma unroll
	for (int d = 16; d > 0; d >>= 1) x += __shfl_down_sync(0xffffffffu, x, d);
	__syncthreads();
	if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = x;
	__syncthreads();
	unsigned long long t = 0;
	if (threadIdx.x == 0)
		for (int w = 0; w < 8; w++) t += s_warp[w];
	return t;                             // valid in thread 0
}

// =========================================================================================
// KMC listing decode
// =========================================================================================
constexpr int kMaxRecBytes = 12;          // suffix <= 8 bytes (k <= 32), counter <= 4 bytes

// stage the record bytes of one tile into shared memory with 16-byte loads
__device__ __forceinline__ void stage_tile(const DevDb& db, uint64_t s0, uint32_t n_rec, uint4* s_stage) {
	const uint64_t byte0 = s0 * db.rec_bytes;                 // multiple of 16: kTile * rec_bytes is
	const uint32_t n_vec = (n_rec * db.rec_bytes + 15) >> 4;   // the device buffer is padded to 16
	const uint4* src = reinterpret_cast<const uint4*>(db.suf + byte0);
	for (uint32_t i = threadIdx.x; i < n_vec; i += blockDim.x) s_stage[i] = __ldg(src + i);
}

__device__ __forceinline__ uint32_t decode_count(const DevDb& db, const uint8_t* rec) {
	uint32_t c = 0;
	for (uint32_t b = 0; b < db.counter_bytes && b < 4; b++) c |= (uint32_t)rec[db.suffix_bytes + b] << (8 * b);
	return c;
}

// number of LUT entries <= s, minus one: the slot whose range holds record s (kmc_file.cpp:439-445)
__device__ __forceinline__ uint64_t lut_slot(const uint64_t* lut, uint64_t lo, uint64_t hi, uint64_t s) {
	// invariant: lut[lo] <= s, answer in [lo, hi]; lut[hi + 1] > s
	while (lo < hi) {
		uint64_t mid = (lo + hi + 1) >> 1;
		if (__ldg(lut + mid) <= s) lo = mid;
		else hi = mid - 1;
	}
	return lo;
}

__device__ __forceinline__ uint64_t decode_kmer(const DevDb& db, const uint8_t* rec, uint64_t slot) {
	uint64_t v = slot & db.prefix_mask;
	for (uint32_t b = 0; b < db.suffix_bytes; b++) v = (v << 8) | rec[b];
	return v;
}

// ---- bulk asynchronous copies (TMA, 1-D) with mbarrier completion: the staging of the pure streaming pass ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
	             "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	uint32_t done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(done)
		             : "r"(smem_u32(bar)), "r"(parity)
		             : "memory");
	} while (!done);
}

// pass 1 (kmodel.hpp:423-434).  LIST = true only counts listed records per tile (kmx_db_list).
// The one pure streaming kernel of the build: a block keeps kCountStages tiles in flight with cp.async.bulk (one elected
// thread issues a 16 KB bulk copy per tile, an mbarrier flips when the bytes have landed), so the HBM pipe stays full while
// the 256 threads pick the counters out of the previous tile; one __syncthreads per tile (the partial sums alternate
// between two shared buffers).
// The consume side is kept to ~12 instructions per record so that the kernel stays bandwidth-bound (at 7 TB/s an SM has
// ~700 cycles for a 16 KB tile): FAST = the record is one aligned 8-byte word (6 suffix bytes + 2 counter bytes: k = 31 with
// a 7-symbol LUT prefix, the layout of every large database) read with one LDS.64; the class counters live in registers
// across tiles and are reduced once at the end; per tile only the array-bound count (what the scan needs) is reduced, with
// one REDUX per warp.
constexpr int kCountStages = 3;
template <bool LIST, bool FAST>
__global__ void __launch_bounds__(256) count_kernel(const __grid_constant__ DevDb db, int ci, int cs, int bf_num, CountOut* out,
                                                    uint32_t* __restrict__ tile_cnt, uint64_t tile_first, uint64_t tile_end) {
	extern __shared__ __align__(128) uint8_t s_dyn[];        // kCountStages stages of stage_bytes each
	__shared__ uint64_t s_bar[kCountStages];
	__shared__ uint32_t s_sum[2][8];                          // [parity][warp]: records of the tile that go on (listed / array-bound)
	__shared__ unsigned long long s_fin[8];
	const uint32_t stage_bytes = ((uint32_t)kTile * db.rec_bytes + 127u) & ~127u;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	auto issue = [&](uint64_t tile, int stage) {              // thread 0 only
		const uint64_t s0 = tile * kTile;
		const uint32_t n_rec = (uint32_t)min((uint64_t)kTile, db.total - s0);
		const uint32_t bytes = (n_rec * db.rec_bytes + 15u) & ~15u;          // the device buffer is padded to 16
		mbar_expect_tx(&s_bar[stage], bytes);
		bulk_g2s(s_dyn + (size_t)stage * stage_bytes, db.suf + s0 * db.rec_bytes, bytes, &s_bar[stage]);
	};
	if (threadIdx.x == 0) {
		for (int q = 0; q < kCountStages; q++) mbar_init(&s_bar[q], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		for (int q = 0; q < kCountStages; q++) {
			const uint64_t tile = tile_first + blockIdx.x + (uint64_t)q * gridDim.x;
			if (tile < tile_end) issue(tile, q);
		}
	}
	__syncthreads();
	const uint32_t lo = db.min_count, hi = db.max_count, bf_end = (uint32_t)(ci + bf_num);
	uint32_t cls0 = 0, cls1 = 0, cls2 = 0, listed_all = 0, arr_all = 0, bad_all = 0;   // this thread's records, all tiles (< 2^32 per thread)
	uint32_t it = 0;
	for (uint64_t tile = tile_first + blockIdx.x; tile < tile_end; tile += gridDim.x, it++) {
		const int stage = (int)(it % kCountStages);
		const uint8_t* s_bytes = s_dyn + (size_t)stage * stage_bytes;
		const uint64_t s0 = tile * kTile;
		const uint32_t n_rec = (uint32_t)min((uint64_t)kTile, db.total - s0);
		mbar_wait(&s_bar[stage], (it / kCountStages) & 1u);
		uint32_t on = 0;                                      // records of this tile that go on: listed (LIST) / array-bound
#pragma unroll
		for (int j = 0; j < kTile / 256; j++) {
			const uint32_t t = threadIdx.x + j * 256;
			uint32_t c;
			if (FAST) c = (uint32_t)(reinterpret_cast<const unsigned long long*>(s_bytes)[t] >> 48);    // bytes 6, 7: little-endian counter
			else c = decode_count(db, s_bytes + t * db.rec_bytes);
			const bool is_listed = t < n_rec && c >= lo && c <= hi;
			if (LIST) {
				on += is_listed ? 1u : 0u;
			} else {
				const bool is_bad = is_listed && (c < (uint32_t)ci || c > (uint32_t)cs);
				const bool is_arr = is_listed && !is_bad && c >= bf_end;
				const uint32_t k = c - (uint32_t)ci;          // Bloom class when listed, not bad, not array-bound
				const bool is_bf = is_listed && !is_bad && !is_arr;
				cls0 += (is_bf && k == 0) ? 1u : 0u;
				cls1 += (is_bf && k == 1) ? 1u : 0u;
				cls2 += (is_bf && k == 2) ? 1u : 0u;
				bad_all += is_bad ? 1u : 0u;
				on += is_arr ? 1u : 0u;
			}
			listed_all += is_listed ? 1u : 0u;
		}
		arr_all += on;
		on = __reduce_add_sync(0xffffffffu, on);
		if (lane == 0) s_sum[it & 1][warp] = on;
		__syncthreads();                                     // every thread is done with the stage; the partial sums are visible
		if (threadIdx.x == 0) {
			const uint64_t next = tile + (uint64_t)kCountStages * gridDim.x;
			if (next < tile_end) issue(next, stage);
			uint32_t sum = 0;
#pragma unroll
			for (int w = 0; w < 8; w++) sum += s_sum[it & 1][w];
			tile_cnt[tile - tile_first] = sum;
		}
	}
	if (out) {
		// the six totals of the block, one after the other (once per kernel)
		const unsigned long long v[6] = { cls0, cls1, cls2, listed_all, LIST ? 0u : arr_all, bad_all };
		unsigned long long* dst[6] = { &out->class_count[0], &out->class_count[1], &out->class_count[2], &out->listed, &out->array_bound, &out->bad_count };
		for (int q = 0; q < 6; q++) {
			const unsigned long long tot = block_sum(v[q], s_fin);
			if (threadIdx.x == 0 && tot) atomicAdd(dst[q], tot);
		}
	}
}

// exclusive scan of per-tile counts (single block; the tile count is N/2048)
__global__ void __launch_bounds__(1024) tile_scan_kernel(const uint32_t* __restrict__ cnt, uint64_t n, uint64_t* __restrict__ off) {
	__shared__ unsigned long long s_warp[32];
	__shared__ unsigned long long s_carry;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_carry = 0;
	__syncthreads();
	for (uint64_t base = 0; base < n; base += 1024) {
		uint64_t i = base + threadIdx.x;
		unsigned long long x = i < n ? cnt[i] : 0, incl = x;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) {
			unsigned long long y = __shfl_up_sync(0xffffffffu, incl, d);
			if (lane >= d) incl += y;
		}
		if (lane == 31) s_warp[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			unsigned long long w = s_warp[lane], wi = w;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) {
				unsigned long long y = __shfl_up_sync(0xffffffffu, wi, d);
				if (lane >= d) wi += y;
			}
			s_warp[lane] = wi - w;
		}
		__syncthreads();
		unsigned long long carry = s_carry;
		if (i < n) off[i] = carry + s_warp[warp] + incl - x;
		__syncthreads();
		if (threadIdx.x == 1023) s_carry = carry + s_warp[warp] + incl;
		__syncthreads();
	}
	if (threadIdx.x == 0) off[n] = s_carry;
}

#ifndef KMX_ENCODE_BLOCKS
#define KMX_ENCODE_BLOCKS 3                  // A/B on B200: 4 blocks per SM (64 registers) is 10 % slower on the HC14 shape
#endif
// pass 2.  Each thread decodes 8 CONSECUTIVE records so that the compaction keeps file order.
// LIST = true: write every listed record (kmx_db_list).  LIST = false: Bloom-bound records are
// inserted into their filters (kmodel.hpp:473-477,498-506), array-bound ones go to the stream.
template <bool LIST, int K, int H>
__global__ void __launch_bounds__(256, KMX_ENCODE_BLOCKS) encode_kernel(const __grid_constant__ DevDb db, const __grid_constant__ DevModel m,
                                                     const uint64_t* __restrict__ tile_off, const __grid_constant__ ItemRoute route,
                                                     uint64_t tile_first, uint64_t tile_end) {
	__shared__ uint4 s_stage[kTile * kMaxRecBytes / 16];
	__shared__ uint32_t s_warp[9];
	__shared__ uint64_t s_slot[2];
	// Bloom-bound k-mers of the tile, compacted: a thread walks 8 consecutive records (the stream must keep file order),
	// of which some fraction is Bloom-bound -- hashing them in place would run the ~650-instruction insert 8 times per
	// warp with that fraction of the lanes; from this list every lane of the block has work (order is free for ORs)
	__shared__ uint64_t s_bloom_kmer[LIST ? 1 : kTile];
	__shared__ uint8_t s_bloom_class[LIST ? 1 : kTile];
	__shared__ unsigned int s_bloom_n;
	const uint8_t* s_bytes = reinterpret_cast<const uint8_t*>(s_stage);
	const int k = K ? K : db.k;
	const int nh = H ? H : m.n_hash;
	constexpr int PER = kTile / 256;
	for (uint64_t tile = tile_first + blockIdx.x; tile < tile_end; tile += gridDim.x) {
		const uint64_t s0 = tile * kTile;
		const uint32_t n_rec = (uint32_t)min((uint64_t)kTile, db.total - s0);
		__syncthreads();
		if (threadIdx.x == 0) s_bloom_n = 0;
		stage_tile(db, s0, n_rec, s_stage);
		if (threadIdx.x < 2) {
			uint64_t s = threadIdx.x == 0 ? s0 : s0 + n_rec - 1;
			s_slot[threadIdx.x] = lut_slot(db.lut, 0, db.lut_entries - 1, s);
		}
		__syncthreads();
		const uint64_t slot_lo = s_slot[0], slot_hi = s_slot[1];
		uint64_t kmer[PER];
		uint32_t cnt[PER];
		uint32_t keep = 0, n_keep = 0;
#pragma unroll
		for (int j = 0; j < PER; j++) {
			uint32_t t = threadIdx.x * PER + j;
			if (t < n_rec) {
				const uint8_t* rec = s_bytes + t * db.rec_bytes;
				uint32_t c = decode_count(db, rec);
				cnt[j] = c;
				if (c >= db.min_count && c <= db.max_count) {
					const bool to_stream = LIST ? true : (c >= (uint32_t)(m.ci + m.bf_num));
					kmer[j] = decode_kmer(db, rec, lut_slot(db.lut, slot_lo, slot_hi, s0 + t));
					if (to_stream) {
						keep |= 1u << j;
						n_keep++;
					} else if (!LIST && c >= (uint32_t)m.ci) {
						cg::coalesced_group g = cg::coalesced_threads();
						unsigned int at = 0;
						if (g.thread_rank() == 0) at = atomicAdd(&s_bloom_n, g.size());
						at = g.shfl(at, 0) + g.thread_rank();
						s_bloom_kmer[at] = kmer[j];
						s_bloom_class[at] = (uint8_t)(c - (uint32_t)m.ci);
					}
				}
			}
		}
		if (!LIST) {
			__syncthreads();
			const unsigned int n_bloom = s_bloom_n;
			for (unsigned int x = threadIdx.x; x < n_bloom; x += blockDim.x) {
				const int f = s_bloom_class[x];
				const uint64_t r = reverse_bases(s_bloom_kmer[x], k);
				HashPrep p;
				hash_prepare(r, k, p);
#pragma unroll
				for (int q = 0; q < (H ? H : kMaxHash) - 1; q++)
					if (q < nh - 1) filter_set(m.bf[f], hash_finish(p, k, c_seeds[q]));
				hash_prepare(middle_r(r, k), k - 2, p);
#pragma unroll
				for (int q = 0; q < (H ? H : kMaxHash) - 2; q++)
					if (q < nh - 2) filter_set(m.bf_back[f], hash_finish(p, k - 2, c_seeds[q]));
			}
		}
		uint32_t total;
		uint32_t rank = block_excl_scan<8>(n_keep, s_warp, &total);
		unsigned long long g = route.base + __ldg(tile_off + (tile - tile_first)) + rank;      // stream position (file order)
#pragma unroll
		for (int j = 0; j < PER; j++) {
			if (keep & (1u << j)) {
				// bucket B = g >> 18 -> batch B / n_bits, bucket i = B % n_bits, round-0 array i, owner i % n_active; the owner
				// keeps its buckets back to back (one GPU: the shard is the stream itself)
				int owner = 0;
				unsigned long long at = g;
				if (route.n_active > 1) route_item(g, route.n_active, route.n_bits, &owner, &at);
				route.kmer[owner][at] = kmer[j];
				route.occ[owner][at] = cnt[j];
				g++;
			}
		}
	}
}

static int stream_grid(uint64_t n_tiles, int sm_count, int per_sm) {
	uint64_t cap = (uint64_t)sm_count * per_sm;
	return (int)(n_tiles < cap ? (n_tiles ? n_tiles : 1) : cap);
}

// grid and dynamic shared memory of the counting pass: kCountStages stages per block, as many blocks per SM as fit
template <bool LIST, bool FAST>
static cudaError_t count_launch(const DevDb& db, int ci, int cs, int bf_num, CountOut* d_out, uint32_t* d_tile_cnt, uint64_t tile_first,
                                uint64_t tile_end, int sm_count, cudaStream_t stream) {
	const size_t smem = (size_t)kCountStages * (((size_t)kTile * db.rec_bytes + 127) & ~(size_t)127);
	cudaError_t e = cudaFuncSetAttribute(count_kernel<LIST, FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	if (e != cudaSuccess) return e;
	int per_sm = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, count_kernel<LIST, FAST>, 256, smem);
	if (e != cudaSuccess) return e;
	if (per_sm < 1) return cudaErrorLaunchOutOfResources;
	const int grid = stream_grid(tile_end - tile_first, sm_count, per_sm);
	count_kernel<LIST, FAST><<<grid, 256, smem, stream>>>(db, ci, cs, bf_num, d_out, d_tile_cnt, tile_first, tile_end);
	note_launch();
	return cudaGetLastError();
}

// one aligned 8-byte word per record with the counter in its top two bytes
static bool fast_records(const DevDb& db) { return db.rec_bytes == 8 && db.counter_bytes == 2 && db.suffix_bytes == 6; }

cudaError_t launch_count(const DevDb& db, int ci, int cs, int bf_num, CountOut* d_out, uint32_t* d_tile_cnt, uint64_t tile_first,
                         uint64_t tile_end, int sm_count, cudaStream_t stream) {
	if (tile_end <= tile_first) return cudaSuccess;
	if (fast_records(db)) return count_launch<false, true>(db, ci, cs, bf_num, d_out, d_tile_cnt, tile_first, tile_end, sm_count, stream);
	return count_launch<false, false>(db, ci, cs, bf_num, d_out, d_tile_cnt, tile_first, tile_end, sm_count, stream);
}

cudaError_t launch_list_count(const DevDb& db, uint32_t* d_tile_cnt, int sm_count, cudaStream_t stream) {
	if (db.total == 0) return cudaSuccess;
	const uint64_t n_tiles = (db.total + kTile - 1) / kTile;
	if (fast_records(db)) return count_launch<true, true>(db, 0, 0, 0, nullptr, d_tile_cnt, 0, n_tiles, sm_count, stream);
	return count_launch<true, false>(db, 0, 0, 0, nullptr, d_tile_cnt, 0, n_tiles, sm_count, stream);
}

cudaError_t launch_tile_scan(const uint32_t* d_tile_cnt, uint64_t n_tiles, uint64_t* d_tile_off, cudaStream_t stream) {
	tile_scan_kernel<<<1, 1024, 0, stream>>>(d_tile_cnt, n_tiles, d_tile_off);
	note_launch();
	return cudaGetLastError();
}

cudaError_t launch_encode(const DevDb& db, const DevModel& m, const uint64_t* d_tile_off, const ItemRoute& route, uint64_t tile_first,
                          uint64_t tile_end, int sm_count, cudaStream_t stream) {
	if (tile_end <= tile_first) return cudaSuccess;
	int grid = stream_grid(tile_end - tile_first, sm_count, 8);
	if (db.k == 31 && m.n_hash == 7)
		encode_kernel<false, 31, 7><<<grid, 256, 0, stream>>>(db, m, d_tile_off, route, tile_first, tile_end);
	else
		encode_kernel<false, 0, 0><<<grid, 256, 0, stream>>>(db, m, d_tile_off, route, tile_first, tile_end);
	note_launch();
	return cudaGetLastError();
}

cudaError_t launch_list(const DevDb& db, const uint64_t* d_tile_off, uint64_t* d_kmers, uint32_t* d_counts, int sm_count,
                        cudaStream_t stream) {
	if (db.total == 0) return cudaSuccess;
	uint64_t n_tiles = (db.total + kTile - 1) / kTile;
	DevModel dummy = {};
	ItemRoute route = {};
	route.kmer[0] = d_kmers;
	route.occ[0] = d_counts;
	route.n_active = 1;
	route.n_bits = 1;
	encode_kernel<true, 0, 0><<<stream_grid(n_tiles, sm_count, 8), 256, 0, stream>>>(db, dummy, d_tile_off, route, 0, n_tiles);
	note_launch();
	return cudaGetLastError();
}

// =========================================================================================
// greedy coupled-array insert (persistent, cooperative)
// =========================================================================================
constexpr uint32_t kStateShift = 30;
constexpr uint32_t kAccepted = 1u << kStateShift, kRejected = 2u << kStateShift;
constexpr uint32_t kNeedShift = 14;                          // status: state(2) | contested(14) | untagged(14)
constexpr uint32_t kMaskBits = (1u << kNeedShift) - 1;
constexpr uint32_t kEpochMax = 0x3FFFu;
#ifndef KMX_INS_THREADS
#define KMX_INS_THREADS 256
#endif
constexpr int kInsThreads = KMX_INS_THREADS;                // threads per block of the persistent insert kernel
constexpr uint32_t kIdTile = 256;                           // ids per reorder tile (one warp of the place pass takes a tile)

template <int K, int H, int B>
__global__ void __launch_bounds__(kInsThreads, 1024 / kInsThreads) insert_kernel(const __grid_constant__ DevModel m, const __grid_constant__ InsertArgs a) {
	cg::grid_group grid = cg::this_grid();
	constexpr int HM = H ? H : kMaxHash;
	constexpr int BM = B ? B : kMaxArrays;
	const int k = K ? K : m.k;
	const int nh = H ? H : m.n_hash;
	const int nb = B ? B : m.n_bits;
	const int hk = nh - 2;
	const uint32_t T = gridDim.x * blockDim.x;
	const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
	const uint32_t total_ids = (uint32_t)nb << kBucketLog;
	const uint32_t n_id_tiles = total_ids / kIdTile;
	const uint32_t slot_mask = a.resv_slots - 1;
	InsertCtl* ctl = a.ctl;
	volatile InsertCtl* vctl = a.ctl;

	uint32_t epoch = vctl->epoch;
	uint32_t seq = vctl->seq;
	const unsigned long long evict_first = make_evict_first_policy();
	const unsigned long long evict_last = make_evict_last_policy();
#if KMX_GRIDBAR
	GridBarrier gbar;
	gbar.init(&ctl->bar);
#define GSYNC() gbar.sync()
#else
#define GSYNC() grid.sync()
#endif
	// Control words every thread needs after a barrier (list lengths, error flag, survivor counts) are fetched ONCE per
	// block, by the thread that waited at the barrier, and handed out through shared memory: a word read by every warp of
	// the grid is several thousand requests to one L2 slice, microseconds per word and barrier.
	__shared__ uint32_t s_hot[12];                        // list_n[0..2], error, nfail[parity][0..7]
	auto gsync_fetch = [&](unsigned int parity) {
#if KMX_GRIDBAR
		gbar.arrive_wait();
#else
		grid.sync();
#endif
		if (threadIdx.x == 0) {
			uint4 h, n0, n1;
			asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(h.x), "=r"(h.y), "=r"(h.z), "=r"(h.w) : "l"(&ctl->list_n[0]) : "memory");
			asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(n0.x), "=r"(n0.y), "=r"(n0.z), "=r"(n0.w) : "l"(&ctl->nfail[parity][0]) : "memory");
			asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(n1.x), "=r"(n1.y), "=r"(n1.z), "=r"(n1.w) : "l"(&ctl->nfail[parity][4]) : "memory");
			s_hot[0] = h.x; s_hot[1] = h.y; s_hot[2] = h.z; s_hot[3] = h.w;
			s_hot[4] = n0.x; s_hot[5] = n0.y; s_hot[6] = n0.z; s_hot[7] = n0.w;
			s_hot[8] = n1.x; s_hot[9] = n1.y; s_hot[10] = n1.z; s_hot[11] = n1.w;
		}
		__syncthreads();
	};
	grid.sync();                                            // everybody has read ctl->epoch / seq / bar before they are rewritten

	const uint32_t per_batch_mine = buckets_per_batch(a.rank, a.n_active, nb);
	for (unsigned long long batch = a.first_batch; batch < a.first_batch + a.n_batches; batch++) {
		const unsigned long long base = batch * total_ids;
		if (base >= a.n_items) break;
		uint32_t n_cur[BM];
#pragma unroll
		for (int i = 0; i < BM; i++) {
			unsigned long long b0 = base + ((unsigned long long)i << kBucketLog);
			unsigned long long rem = (i < nb && a.n_items > b0) ? a.n_items - b0 : 0;
			n_cur[i] = (uint32_t)(rem < kBucket ? rem : kBucket);
		}
		// Stale slot 0 (kmodel.hpp:520-527 + 529-540): in the final, partial batch the trailing
		// buckets have length 0, and reorder_buffer(a, 0) returns 1 when a[0].occ != 0.  a[0] is
		// what the previous batch left there: that bucket's first survivor, which is already in
		// the rest table.  It is offered to the other arrays again, rejected again (the state only
		// grows) and pushed to the rest table a second time.  Net effect: one duplicate rest entry.
		if (tid == 0 && batch > 0) {
			for (int i = 0; i < nb; i++) {
				if (n_cur[i] == 0 && vctl->slot0_valid[i]) {
					unsigned long long at = atomicAdd(&ctl->rest_n, 1ULL);
					if (at < a.rest_cap) {
						a.rest_kmer[at] = vctl->slot0_kmer[i];
						a.rest_occ[at] = vctl->slot0_occ[i];
					} else {
						vctl->error = 2;
					}
				}
			}
		}

		for (int t = 0; t < nb; t++) {
			const bool last_round = (t == nb - 1);
			const uint64_t* src_kmer = t == 0 ? a.item_kmer : a.buf_kmer[(t - 1) & 1];
			const uint32_t* src_occ = t == 0 ? a.item_occ : a.buf_occ[(t - 1) & 1];
			// where item id = (bucket i << 18 | c) of this round lives in src_*: rounds > 0 read the ping-pong buffer, indexed by
			// id; round 0 reads the item stream, in which this owner keeps its buckets (those whose round-0 array i it owns) back
			// to back, batch after batch (one GPU: bucket i of batch b is bucket b * n_bits + i of the stream)
			uint32_t src_bucket[BM];
#pragma unroll
			for (int i = 0; i < BM; i++)
				src_bucket[i] = t == 0 ? (uint32_t)batch * per_batch_mine + (uint32_t)i / (uint32_t)a.n_active : (uint32_t)i;
			auto src_at = [&](uint32_t id) -> size_t {
				uint32_t sb = 0;
#pragma unroll
				for (int i = 0; i < BM; i++) sb = (id >> kBucketLog) == (uint32_t)i ? src_bucket[i] : sb;
				return ((size_t)sb << kBucketLog) | (id & (kBucket - 1));
			};

			// ---------------- decide every item of the round ----------------
			// ItemCtx: what every phase recomputes from the item (hashing is cheaper than keeping
			// n_hash positions per item in memory between phases)
			struct ItemCtx {
				uint64_t r;
				uint64_t pos[HM];
				uint32_t bin, c;
				int arr;
			};
			auto prepare_item = [&](uint32_t id, uint64_t v, uint32_t occ, ItemCtx& it) {
				const uint32_t i = id >> kBucketLog;
				it.c = id & (kBucket - 1);
				it.arr = (int)((i + (uint32_t)t) % (uint32_t)nb);
				it.bin = __ldg(m.occ2bin + (occ > (uint32_t)m.cs ? (uint32_t)m.cs : occ));
				it.r = reverse_bases(v, k);
				HashPrep p;
				hash_prepare(it.r, k, p);
#pragma unroll
				for (int j = 0; j < HM; j++)
					if (j < nh) it.pos[j] = fastmod(hash_finish(p, k, m.arr_seed[it.arr][j]), m.arr_mod);
			};
			auto load_item = [&](uint32_t id, ItemCtx& it) { prepare_item(id, __ldcg(src_kmer + src_at(id)), __ldcg(src_occ + src_at(id)), it); };
			// read the item's cells: conflict with the committed state? which positions are still untagged?
			auto read_cells = [&](const ItemCtx& it, bool& conflict, uint32_t& untagged) {
				unsigned long long cell[HM];
#pragma unroll
				for (int j = 0; j < HM; j++)
					if (j < nh) cell[j] = (a.stream_cells & 1) ? ld_stream64(m.cells[it.arr] + (it.pos[j] >> 5), evict_first) : __ldcg(m.cells[it.arr] + (it.pos[j] >> 5));
				conflict = false;
				untagged = 0;
#pragma unroll
				for (int j = 0; j < HM; j++) {
					if (j < nh) {
						const uint32_t sh = ((uint32_t)it.pos[j] & 31u) ^ 7u;
						const uint32_t val = ((uint32_t)cell[j] >> sh) & 1u, tag = ((uint32_t)(cell[j] >> 32) >> sh) & 1u;
						conflict |= (tag != 0) && (val != ((it.bin >> j) & 1u));
						untagged |= (tag ^ 1u) << j;
					}
				}
			};
			auto reject = [&](uint32_t id) {
				a.status[id] = kRejected;
				red_add32(a.tile_fail + (id / kIdTile), 1u);
			};
			// accept: set tag (+ value) at every position that was untagged, OR the (k-2)-mer into km_back (kmodel.hpp:546-550)
			auto commit = [&](uint32_t id, const ItemCtx& it, uint32_t untagged) {
#pragma unroll
				for (int j = 0; j < HM; j++) {
					if (j < nh && ((untagged >> j) & 1u)) {      // tagged positions already hold the wanted value
						const uint32_t sh = ((uint32_t)it.pos[j] & 31u) ^ 7u;
						const unsigned long long want = (it.bin >> j) & 1u;
						if (a.stream_cells & 2) red_or64_stream(m.cells[it.arr] + (it.pos[j] >> 5), ((1ULL << 32) | want) << sh, evict_first);
						else red_or64(m.cells[it.arr] + (it.pos[j] >> 5), ((1ULL << 32) | want) << sh);
					}
				}
				HashPrep p;
				hash_prepare(middle_r(it.r, k), k - 2, p);
#pragma unroll
				for (int j = 0; j < HM - 2; j++) {
					if (j < hk) {
						const uint64_t pos = fastmod(hash_finish(p, k - 2, c_seeds[j]), m.km_back.mod);
						if (a.stream_cells & 4) red_or32_stream(m.km_back.words + (pos >> 5), bit_mask32(pos), evict_first);
						else if (a.stream_cells & 8) red_or32_stream(m.km_back.words + (pos >> 5), bit_mask32(pos), evict_last);
						else red_or32(m.km_back.words + (pos >> 5), bit_mask32(pos));
					}
				}
				a.status[id] = kAccepted;
			};
			// (two reservation tables per array: the classic iterations use table 0 only, the merged passes alternate)
			auto reserve = [&](const ItemCtx& it, uint32_t need, uint32_t key_hi, uint32_t tab = 0) {
				uint32_t* table = a.resv + ((size_t)it.arr * 2 + tab) * 2 * a.resv_slots;
#pragma unroll
				for (int j = 0; j < HM; j++)
					if (j < nh && ((need >> j) & 1u))
						red_min32(table + 2 * ((uint32_t)it.pos[j] & slot_mask) + ((it.bin >> j) & 1u), key_hi | it.c);
			};
			auto holds_reservations = [&](const ItemCtx& it, uint32_t need, uint32_t key_hi, uint32_t tab = 0) -> bool {
				const uint32_t* table = a.resv + ((size_t)it.arr * 2 + tab) * 2 * a.resv_slots;
				const uint32_t key = key_hi | it.c;
				bool ok = true;
#pragma unroll
				for (int j = 0; j < HM; j++)
					if (j < nh && ((need >> j) & 1u))
						ok &= __ldcg(table + 2 * ((uint32_t)it.pos[j] & slot_mask) + (((it.bin >> j) & 1u) ^ 1u)) >= key;
				return ok;
			};
			auto append = [&](uint32_t* list, unsigned int* counter, uint32_t id) {
				cg::coalesced_group g = cg::coalesced_threads();
				uint32_t at = 0;
				if (g.thread_rank() == 0) at = atomicAdd(counter, g.size());
				at = g.shfl(at, 0);
				list[at + g.thread_rank()] = id;
			};
			// With several GPUs, array x lives on rank x % n_active: this rank works on the buckets whose
			// array of this round it owns (all of them on a single GPU).
			const unsigned int par = (unsigned int)((batch * (unsigned long long)nb + (unsigned long long)t) & 1ULL);
			uint32_t n_work[BM];
#pragma unroll
			for (int i = 0; i < BM; i++)
				n_work[i] = (i < nb && ((i + t) % nb) % a.n_active == a.rank) ? n_cur[i] : 0;
			// dense walk over this rank's items of the round: x in [0, n_round) -> id
			uint32_t n_round = 0;
#pragma unroll
			for (int i = 0; i < BM; i++) n_round += n_work[i];
			auto dense_to_id = [&](uint32_t x) -> uint32_t {
				uint32_t i = 0;
#pragma unroll
				for (int q = 0; q < BM - 1; q++) {
					if (x >= n_work[q] && i == (uint32_t)q) {
						x -= n_work[q];
						i++;
					}
				}
				return (i << kBucketLog) | x;
			};
			// claim bitmap of this round: sized to the round (about 64 bits per claim), cleared at the end
			uint32_t n_arr_max = 0;                             // items offered to one array this round
#pragma unroll
			for (int i = 0; i < BM; i++) n_arr_max = n_work[i] > n_arr_max ? n_work[i] : n_arr_max;
			uint32_t claim_log2 = 15;
			while (claim_log2 < a.claim_log2 && (1u << claim_log2) < 64u * (uint32_t)nh * (n_arr_max + 1)) claim_log2++;
			const uint32_t claim_mask = (1u << claim_log2) - 1;
			const size_t claim_stride = (size_t)1 << (a.claim_log2 - 5);     // words per (array, want) bitmap

			if (epoch + 2 >= kEpochMax) {                      // reservation keys can get no smaller: start the epochs over
				for (size_t x = tid; x < (size_t)nb * 4 * a.resv_slots; x += T) a.resv[x] = 0xFFFFFFFFu;
				epoch = 0;
				GSYNC();
			}
			long long tick = clock64();
			if (tid == 0) {                                     // nobody touches the lists during phase 0
				vctl->list_n[0] = 0;
				vctl->list_n[1] = 0;
				vctl->list_n[2] = 0;
			}
			// ---- first iteration, phase 0: reject on the committed state, or claim (position, wanted value) ----
			// (the next item's k-mer and count are fetched while the current one is hashed and probed)
			uint32_t id_n = tid < n_round ? dense_to_id(tid) : 0;
			uint64_t v_n = tid < n_round ? __ldcg(src_kmer + src_at(id_n)) : 0;
			uint32_t occ_n = tid < n_round ? __ldcg(src_occ + src_at(id_n)) : 0;
			// what this thread learns about its first item (x = tid) stays in registers for phase 1, whose first
			// iteration is the same item: rounds of at most one item per thread skip two dependent load waves there
			const uint32_t id_first = id_n;
			const uint64_t v_first = v_n;
			const uint32_t occ_first = occ_n;
			uint32_t st_first = 0;
			for (uint32_t x = tid; x < n_round; x += T) {
				const uint32_t id = id_n;
				const uint64_t v_c = v_n;
				const uint32_t occ_c = occ_n;
				if (x + T < n_round) {
					id_n = dense_to_id(x + T);
					v_n = __ldcg(src_kmer + src_at(id_n));
					occ_n = __ldcg(src_occ + src_at(id_n));
				}
				ItemCtx it;
				prepare_item(id, v_c, occ_c, it);
				bool conflict = false;
				uint32_t untagged = (1u << nh) - 1u;
				// claim_first (models that do not fit the L2): claim every position without looking at the cells, so that
				// phase 1 reads a cell and commits to it back to back while its sector is still in the L2 -- one
				// DRAM round trip per sector instead of a read here and a read-modify-write there.  Claims on
				// tagged positions and claims of items that turn out rejected are harmless: a claim can only
				// send somebody through the reservation path.
				if (!a.claim_first) read_cells(it, conflict, untagged);
				if (conflict) {
					reject(id);
					if (x == tid) st_first = kRejected;
				} else {
					uint32_t* cl = a.claim + (size_t)it.arr * 2 * claim_stride;
#pragma unroll
					for (int j = 0; j < HM; j++) {
						if (j < nh && ((untagged >> j) & 1u)) {
							const uint32_t bit = (uint32_t)it.pos[j] & claim_mask;
							red_or32(cl + ((it.bin >> j) & 1u) * claim_stride + (bit >> 5), 1u << (bit & 31u));
						}
					}
					if (!a.claim_first) {
						a.status[id] = untagged;
						if (x == tid) st_first = untagged;
					}
				}
			}
			// (diagnostic: slots 2 / 4 hold the part of phases 0 / 1 block 0 spent in its own loop, the rest is barrier wait)
			if (tid == 0 && a.merged && (a.phase_round < 0 || a.phase_round == t)) vctl->phase_cycles[2] += (unsigned long long)(clock64() - tick);
			GSYNC();
			if (tid == 0 && (a.phase_round < 0 || a.phase_round == t)) {
				long long now = clock64();
				vctl->phase_cycles[0] += (unsigned long long)(now - tick);
				tick = now;
			}
			// ---- phase 1: an item nobody contests (no live item wants the opposite value at any of its
			// untagged positions) interacts with nobody and commits at once; the others reserve ----
			uint32_t key_hi = (kEpochMax - epoch) << kBucketLog;
			// (iteration 0 comes from phase 0's registers, the following ones are fetched one step ahead)
			id_n = id_first;
			v_n = v_first;
			occ_n = occ_first;
			uint32_t st_n = st_first;
			for (uint32_t x = tid; x < n_round; x += T) {
				const uint32_t id = id_n;
				const uint64_t v_c = v_n;
				const uint32_t occ_c = occ_n;
				uint32_t untagged = st_n;
				if (x + T < n_round) {
					id_n = dense_to_id(x + T);
					st_n = a.claim_first ? 0u : __ldcg(a.status + id_n);
					v_n = __ldcg(src_kmer + src_at(id_n));
					occ_n = __ldcg(src_occ + src_at(id_n));
				}
				if (untagged >> kStateShift) continue;
				ItemCtx it;
				prepare_item(id, v_c, occ_c, it);
				if (a.claim_first) {
					bool conflict;
					read_cells(it, conflict, untagged);
					if (conflict) {
						reject(id);
						continue;
					}
				}
				const uint32_t* cl = a.claim + (size_t)it.arr * 2 * claim_stride;
				uint32_t need = 0;
#pragma unroll
				for (int j = 0; j < HM; j++) {
					if (j < nh && ((untagged >> j) & 1u)) {
						const uint32_t bit = (uint32_t)it.pos[j] & claim_mask;
						const uint32_t w = __ldcg(cl + (((it.bin >> j) & 1u) ^ 1u) * claim_stride + (bit >> 5));
						need |= ((w >> (bit & 31u)) & 1u) << j;
					}
				}
				if (need == 0) {
					commit(id, it, untagged);
				} else {
					reserve(it, need, key_hi);
					a.status[id] = untagged | (need << kNeedShift);
					append(a.list[1], &ctl->list_n[1], id);
				}
			}
			if (tid == 0 && a.merged && (a.phase_round < 0 || a.phase_round == t)) vctl->phase_cycles[4] += (unsigned long long)(clock64() - tick);
			gsync_fetch(par);
			if (tid == 0 && (a.phase_round < 0 || a.phase_round == t)) {
				long long now = clock64();
				vctl->phase_cycles[1] += (unsigned long long)(now - tick);
				tick = now;
			}
			uint32_t iter = 1;
			if (a.merged) {
				// ---- merged passes: reserve for the next pass while committing for this one -------------------------
				// A contested item that holds its reservations of the PREVIOUS pass commits; one that does not reserves
				// in the OTHER table for the next pass, in the same sweep -- one barrier per iteration instead of two and
				// one hashing of the item instead of two.  Two items that both commit in a pass never want opposite values
				// at a shared untagged position (both would have it in their contested set and only the smaller index
				// holds it), so commits and the cell reads of the same pass may interleave freely; an item that loses
				// re-reads its cells at the start of the next pass and is rejected there if the winner conflicts with it.
				// A reservation left behind by an item that gets rejected can only delay somebody by one pass.
				uint32_t n_list = s_hot[1];
				int cur = 1;
				uint32_t tab = 0;
				bool fresh = true;                                // status words are current: no need to re-read the cells
				long long pass_tick = clock64();
				while (n_list != 0) {
					if (epoch + 2 >= kEpochMax) {
						// keys can get no smaller: clear the tables and let every undecided item reserve again before
						// anybody commits (the first half of a classic iteration)
						for (size_t x = tid; x < (size_t)nb * 4 * a.resv_slots; x += T) a.resv[x] = 0xFFFFFFFFu;
						epoch = 0;
						GSYNC();
						key_hi = (kEpochMax - epoch) << kBucketLog;
						for (uint32_t x = tid; x < n_list; x += T) {
							const uint32_t id = __ldcg(a.list[cur] + x);
							const uint32_t st = a.status[id];
							if (st >> kStateShift) continue;
							ItemCtx it;
							load_item(id, it);
							bool conflict;
							uint32_t untagged;
							read_cells(it, conflict, untagged);
							if (conflict) {
								reject(id);
								continue;
							}
							const uint32_t need = (st >> kNeedShift) & untagged;
							a.status[id] = untagged | (need << kNeedShift);
							if (need) reserve(it, need, key_hi, tab);
						}
						GSYNC();
						fresh = true;
					}
					const uint32_t key_prev = key_hi;             // what the undecided items reserved with, in table `tab`
					epoch++;
					key_hi = (kEpochMax - epoch) << kBucketLog;   // what this pass reserves with, in table `tab ^ 1`
					// three lists in rotation: this pass reads `cur`, appends to `nxt` and clears the counter of the third,
					// which the pass before last read (every block has fetched that length two barriers ago) and the
					// next pass appends to
					const int nxt = cur == 2 ? 0 : cur + 1, spare = nxt == 2 ? 0 : nxt + 1;
					const uint32_t* list_cur = a.list[cur];
					if (tid == 0) vctl->list_n[spare] = 0;
					for (uint32_t x = tid; x < n_list; x += T) {
						const uint32_t id = __ldcg(list_cur + x);
						const uint32_t st = a.status[id];
						if (st >> kStateShift) continue;          // rejected while the epochs were restarted
						ItemCtx it;
						load_item(id, it);
						uint32_t untagged = st & kMaskBits, need = (st >> kNeedShift) & kMaskBits;
						if (!fresh) {
							bool conflict;
							uint32_t now_untagged;
							read_cells(it, conflict, now_untagged);
							if (conflict) {
								reject(id);
								continue;
							}
							need &= now_untagged;                 // contested positions that are still open
							untagged = now_untagged;
						}
						if (need == 0 || holds_reservations(it, need, key_prev, tab)) {
							commit(id, it, untagged);
						} else {
							reserve(it, need, key_hi, tab ^ 1u);
							a.status[id] = untagged | (need << kNeedShift);
							append(a.list[nxt], &ctl->list_n[nxt], id);
						}
					}
					if (tid == 0 && (a.phase_round < 0 || a.phase_round == t)) {
						vctl->phase_cycles[10] += (unsigned long long)(clock64() - pass_tick);
						vctl->phase_cycles[11] += 1;
					}
					gsync_fetch(par);
					pass_tick = clock64();
					n_list = s_hot[nxt];
					cur = nxt;
					tab ^= 1u;
					fresh = false;
					iter++;
					if (iter >= a.max_iterations) {
						if (tid == 0) vctl->error = 1;
						return;                                     // uniform over the grid
					}
				}
			} else {
				// ---- phase 2: contested items that hold all their reservations commit, the rest go to the list ----
				const uint32_t n_contested = s_hot[1];
				for (uint32_t x = tid; x < n_contested; x += T) {
					const uint32_t id = __ldcg(a.list[1] + x);
					const uint32_t st = a.status[id];
					ItemCtx it;
					load_item(id, it);
					if (holds_reservations(it, st >> kNeedShift, key_hi)) commit(id, it, st & kMaskBits);
					else append(a.list[0], &ctl->list_n[0], id);
				}
				gsync_fetch(par);
				uint32_t n_list = s_hot[0];
				iter = 1;
				int cur = 0;
				epoch++;
				if (tid == 0 && (a.phase_round < 0 || a.phase_round == t)) {
					long long now = clock64();
					vctl->phase_cycles[2] += (unsigned long long)(now - tick);
					tick = now;
				}
				// ---- later iterations walk the list of still undecided (contested) items ----
				while (n_list != 0) {
					if (epoch + 1 >= kEpochMax) {                   // keys can get no smaller: start over
						for (size_t x = tid; x < (size_t)nb * 4 * a.resv_slots; x += T) a.resv[x] = 0xFFFFFFFFu;
						epoch = 0;
						GSYNC();
					}
					key_hi = (kEpochMax - epoch) << kBucketLog;
					const uint32_t* list_cur = a.list[cur];
					if (tid == 0) vctl->list_n[cur ^ 1] = 0;
					for (uint32_t x = tid; x < n_list; x += T) {
						const uint32_t id = __ldcg(list_cur + x);
						const uint32_t st = a.status[id];
						ItemCtx it;
						load_item(id, it);
						bool conflict;
						uint32_t untagged;
						read_cells(it, conflict, untagged);
						if (conflict) {
							reject(id);
							continue;
						}
						const uint32_t need = (st >> kNeedShift) & untagged;     // contested positions that are still open
						if (need == 0) {
							commit(id, it, untagged);
						} else {
							reserve(it, need, key_hi);
							a.status[id] = untagged | (need << kNeedShift);
						}
					}
					GSYNC();
					for (uint32_t x = tid; x < n_list; x += T) {
						const uint32_t id = __ldcg(list_cur + x);
						const uint32_t st = a.status[id];
						if (st >> kStateShift) continue;
						ItemCtx it;
						load_item(id, it);
						if (holds_reservations(it, st >> kNeedShift, key_hi)) commit(id, it, st & kMaskBits);
						else append(a.list[cur ^ 1], &ctl->list_n[cur ^ 1], id);
					}
					gsync_fetch(par);
					n_list = s_hot[cur ^ 1];
					cur ^= 1;
					iter++;
					epoch++;
					if (iter >= a.max_iterations) {
						if (tid == 0) vctl->error = 1;
						return;                                     // uniform over the grid
					}
				}
			}
			if (tid == 0 && (a.phase_round < 0 || a.phase_round == t)) {
				long long now = clock64();
				vctl->phase_cycles[3] += (unsigned long long)(now - tick);
				tick = now;
			}

			// ---------------- reorder_buffer (kmodel.hpp:529-540) in closed form ----------------
			// F = number of rejected items.  Rejected items below F stay where they are; the
			// accepted slots below F ("holes", ascending) are filled by the rejected items at or
			// above F taken in DESCENDING index order.
			//
			// Place pass, one WARP per tile of 256 ids (lane l takes ids l, l + 32, ...: every access is coalesced).  The
			// warp sums the per-tile reject counters of its bucket (<= 1024 words, L2-resident) into the number of rejects
			// before its tile and the bucket total F -- there is no separate scan phase.  In the last round every survivor
			// goes to the rest list, whose order does not matter (it is sorted afterwards): a tile takes its range with one
			// atomicAdd and the round needs no move pass; only buffer slot 0 (kmodel.hpp:520-540) is tracked.
			constexpr uint32_t kTilesPerBucket = kBucket / kIdTile;
			static_assert(kIdTile == 256 && kTilesPerBucket == 1024, "the place pass gives a tile of 256 ids to a warp");
			if (tid == 0) {                                     // empty buckets of this rank: nobody visits their tiles
				for (int i = 0; i < nb; i++) {
					if (((i + t) % nb) % a.n_active == a.rank && n_work[i] == 0) {
						for (int p = 0; p < a.n_active; p++) ((volatile InsertCtl*)a.peer_ctl[p])->nfail[par][i] = 0;
						if (last_round) vctl->slot0_valid[i] = 0u;
					}
				}
			}
			{
				const uint32_t lane = threadIdx.x & 31u;
				const uint32_t warp_g = tid >> 5, n_warps = T >> 5;
				for (uint32_t tile = warp_g; tile < n_id_tiles; tile += n_warps) {
					const uint32_t id_base = tile * kIdTile;
					const uint32_t i = id_base >> kBucketLog, c_base = id_base & (kBucket - 1);
					const uint32_t nw = n_work[i];
					if (c_base >= nw) continue;                    // uniform over the warp: empty tile
					uint32_t st[8];
#pragma unroll
					for (int q = 0; q < 8; q++) st[q] = __ldcg(a.status + id_base + q * 32u + lane);
					const uint32_t tb = c_base / kIdTile, tiles_used = (nw + kIdTile - 1) / kIdTile;
					const uint32_t* tf = a.tile_fail + i * kTilesPerBucket;
					uint32_t before = 0, F = 0;
					// (a lane sums up to 32 counters: eight independent loads in flight at a time instead of one L2 round trip each)
					for (uint32_t q0 = 0; q0 * 32u < tiles_used; q0 += 8u) {
						uint32_t v[8];
#pragma unroll
						for (uint32_t u = 0; u < 8u; u++) {
							const uint32_t k = lane + 32u * (q0 + u);
							v[u] = k < tiles_used ? __ldcg(tf + k) : 0u;
						}
#pragma unroll
						for (uint32_t u = 0; u < 8u; u++) {
							const uint32_t k = lane + 32u * (q0 + u);
							F += v[u];
							before += k < tb ? v[u] : 0u;
						}
					}
#pragma unroll
					for (int d = 16; d > 0; d >>= 1) {
						F += __shfl_xor_sync(0xffffffffu, F, d);
						before += __shfl_xor_sync(0xffffffffu, before, d);
					}
					if (tb == 0 && lane == 0) {
						for (int p = 0; p < a.n_active; p++) ((volatile InsertCtl*)a.peer_ctl[p])->nfail[par][i] = F;   // every rank tracks every bucket
						if (last_round) vctl->slot0_valid[i] = F ? 1u : 0u;
					}
					unsigned long long rest_at = 0;
					bool first_accepted = false;
					if (last_round) {
						const uint32_t mine_cnt = __ldcg(tf + tb);
						if (lane == 0 && mine_cnt) {
							rest_at = atomicAdd(&ctl->rest_n, (unsigned long long)mine_cnt);
							if (rest_at + mine_cnt > a.rest_cap) vctl->error = 2;
						}
						rest_at = __shfl_sync(0xffffffffu, rest_at, 0);
						if (rest_at + mine_cnt > a.rest_cap) continue;          // uniform over the warp; the grid stops after the barrier
						first_accepted = (__ldcg(a.status + (i << kBucketLog)) >> kStateShift) != 2u;
					}
					uint32_t excl = before;
#pragma unroll
					for (int q = 0; q < 8; q++) {
						const uint32_t c = c_base + q * 32u + lane, id = id_base + q * 32u + lane;
						const bool valid = c < nw;
						const bool failed = valid && (st[q] >> kStateShift) == 2u;
						const uint32_t ball = __ballot_sync(0xffffffffu, failed);
						const uint32_t mine = excl + __popc(ball & ((1u << lane) - 1u));
						excl += __popc(ball);
						if (!valid) continue;
						if (failed) {
							if (last_round) {
								const uint64_t v = __ldcg(src_kmer + src_at(id));
								const uint32_t occ = __ldcg(src_occ + src_at(id));
								a.rest_kmer[rest_at + (mine - before)] = v;
								a.rest_occ[rest_at + (mine - before)] = occ;
								// what reorder_buffer leaves in slot 0: item 0 if it was rejected, else the last rejected item
								if (c == 0 || (first_accepted && mine == F - 1u)) {
									vctl->slot0_kmer[i] = v;
									vctl->slot0_occ[i] = occ;
								}
							} else if (c < F) {
								const int nx = ((int)(i + (uint32_t)t + 1u) % nb) % a.n_active;   // owner of this bucket's next array
								a.peer_buf_kmer[t & 1][nx][id] = __ldcg(src_kmer + src_at(id));
								a.peer_buf_occ[t & 1][nx][id] = __ldcg(src_occ + src_at(id));
							} else {
								a.excl_rank[id] = mine;
							}
						} else if (!last_round && c < F) {
							a.holepos[(i << kBucketLog) + (c - mine)] = c;
						}
					}
				}
			}
			if (tid == 0 && (a.phase_round < 0 || a.phase_round == t)) vctl->phase_cycles[8] += (unsigned long long)(clock64() - tick);
			gsync_fetch(par);
			if (s_hot[3]) return;                               // uniform over the grid (survivor list overflow)
			if (tid == 0 && (a.phase_round < 0 || a.phase_round == t)) {
				long long now = clock64();
				vctl->phase_cycles[5] += (unsigned long long)(now - tick);
				tick = now;
			}
			uint32_t n_next[BM];
#pragma unroll
			for (int i = 0; i < BM; i++) {
				const bool mine = i < nb && ((i + t) % nb) % a.n_active == a.rank;
				n_next[i] = mine ? s_hot[4 + i] : 0;                     // the other buckets' counts arrive with the round barrier
			}
			if (!last_round) {
				for (uint32_t x = tid; x < n_round; x += T) {
					const uint32_t id = dense_to_id(x);
					const uint32_t i = id >> kBucketLog, c = id & (kBucket - 1);
					const uint32_t F = n_next[i];
					if (c < F) continue;
					// status, rank and the item itself are independent loads (one L2 round trip); only the hole position depends on the rank
					const uint32_t st = __ldcg(a.status + id);
					const uint32_t er = __ldcg(a.excl_rank + id);
					const uint64_t v = __ldcg(src_kmer + src_at(id));
					const uint32_t occ = __ldcg(src_occ + src_at(id));
					if ((st >> kStateShift) != 2u) continue;
					const uint32_t to = __ldcg(a.holepos + (i << kBucketLog) + (F - er - 1));
					const int nx = ((int)(i + (uint32_t)t + 1u) % nb) % a.n_active;
					a.peer_buf_kmer[t & 1][nx][(i << kBucketLog) + to] = v;
					a.peer_buf_occ[t & 1][nx][(i << kBucketLog) + to] = occ;
				}
			}
			for (uint32_t x = tid; x < n_id_tiles; x += T) a.tile_fail[x] = 0;
			{   // clear the part of the claim bitmaps this round used (only the arrays this rank worked on), 16 bytes per store;
				// a bitmap that spans the whole array (claim bit = position, small models) is only as long as the array
				const unsigned long long span = (1ULL << claim_log2) < m.arr_mod.d ? (1ULL << claim_log2) : m.arr_mod.d;
				const uint32_t vecs = (uint32_t)((span + 127) >> 7);
				uint32_t act = 0, n_act = 0;                     // 4 bits per active array
#pragma unroll
				for (int i = 0; i < BM; i++) {
					if (n_work[i] != 0) {
						act |= (uint32_t)((i + t) % nb) << (4 * n_act);
						n_act++;
					}
				}
				const uint32_t total = n_act * 2 * vecs;
				for (uint32_t x = tid; x < total; x += T) {
					const uint32_t q = x / vecs;
					const uint32_t arr = (act >> (4 * (q >> 1))) & 15u;
					reinterpret_cast<uint4*>(a.claim + ((size_t)arr * 2 + (q & 1u)) * claim_stride)[x % vecs] = make_uint4(0u, 0u, 0u, 0u);
				}
			}
			if (tid == 0) {
				unsigned long long att = 0, fail = 0;
				for (int i = 0; i < nb; i++) {
					att += n_work[i];
					fail += n_next[i];
				}
				vctl->attempts += att;
				vctl->accepted += att - fail;
				vctl->iterations += iter;
			}
			if (tid == 0 && (a.phase_round < 0 || a.phase_round == t)) vctl->phase_cycles[9] += (unsigned long long)(clock64() - tick);
			if (a.n_active > 1) __threadfence_system();          // survivors written into a peer's buffers
			if (a.n_active > 1) GSYNC();
			else gsync_fetch(par);
			if (tid == 0 && (a.phase_round < 0 || a.phase_round == t)) vctl->phase_cycles[6] += (unsigned long long)(clock64() - tick);
			if (a.n_active > 1) {
				// round barrier across the GPUs: publish "round seq done" on every rank, wait for all of them
				seq++;
				if (tid == 0) {
					const long long wait0 = clock64();
					__threadfence_system();
					for (int p = 0; p < a.n_active; p++) ((volatile uint32_t*)a.peer_flags[p])[a.rank] = seq;
					unsigned long long t0, t1;
					asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
					for (int q = 0; q < a.n_active; q++) {
						while ((int)(((volatile uint32_t*)a.peer_flags[a.rank])[q] - seq) < 0) {
							asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
							if (t1 - t0 > 20000000000ULL) {          // 20 s: a peer died; give up instead of hanging the GPU
								vctl->error = 3;
								break;
							}
						}
					}
					__threadfence_system();
					vctl->phase_cycles[7] += (unsigned long long)(clock64() - wait0);
				}
				gsync_fetch(par);
			}
			if (s_hot[3]) return;                               // uniform over the grid
#pragma unroll
			for (int i = 0; i < BM; i++) n_cur[i] = i < nb ? s_hot[4 + i] : 0;
		}
	}
	if (tid == 0) {
		vctl->epoch = epoch;
		vctl->seq = seq;
	}
}

#undef GSYNC

cudaError_t insert_grid_size(int* blocks_out, int sm_count) {
	int per_sm = 0;
	cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, insert_kernel<31, 7, 5>, kInsThreads, 0);
	if (e != cudaSuccess) return e;
	int per_sm_g = 0;
	e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_g, insert_kernel<0, 0, 0>, kInsThreads, 0);
	if (e != cudaSuccess) return e;
	if (per_sm_g < per_sm) per_sm = per_sm_g;
	if (per_sm < 1) return cudaErrorLaunchOutOfResources;
	*blocks_out = per_sm * sm_count;
	return cudaSuccess;
}

cudaError_t launch_insert(const DevModel& m, const InsertArgs& a, int grid_blocks, cudaStream_t stream) {
	if (grid_blocks < m.n_bits) return cudaErrorInvalidConfiguration;
	void* args[2] = { (void*)&m, (void*)&a };
	note_launch();
	if (m.k == 31 && m.n_hash == 7 && m.n_bits == 5)
		return cudaLaunchCooperativeKernel((const void*)insert_kernel<31, 7, 5>, dim3(grid_blocks), dim3(kInsThreads), args, 0, stream);
	return cudaLaunchCooperativeKernel((const void*)insert_kernel<0, 0, 0>, dim3(grid_blocks), dim3(kInsThreads), args, 0, stream);
}

// =========================================================================================
// rest table (rest.hpp:95-135): survivors sorted by packed value (kmx_sort.cu), group index per prefix
// =========================================================================================
__global__ void rest_first_kernel(const uint64_t* __restrict__ keys, uint64_t n, int suffix_bits, int32_t* __restrict__ first) {
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t pre = (uint32_t)(keys[i] >> suffix_bits);
		if (i == 0 || (uint32_t)(keys[i - 1] >> suffix_bits) != pre) first[pre] = (int32_t)i;
	}
}

// stat() + the index half of transform() (rest.hpp:95-105,115-126): dense ids for the
// non-empty prefixes in ascending order, pre_buffer = first entry of each group, then n
__global__ void __launch_bounds__(1024) rest_index_kernel(const int32_t* __restrict__ first, int map_size, uint64_t n,
                                                          int32_t* __restrict__ hash2index, int32_t* __restrict__ pre_buffer,
                                                          int32_t* __restrict__ groups) {
	__shared__ uint32_t s_warp[32];
	__shared__ uint32_t s_carry;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (threadIdx.x == 0) s_carry = 0;
	__syncthreads();
	for (int base = 0; base < map_size; base += 1024) {
		int p = base + threadIdx.x;
		int32_t f = p < map_size ? first[p] : -1;
		uint32_t x = f >= 0 ? 1u : 0u;
		uint32_t incl = warp_incl_scan(x);
		if (lane == 31) s_warp[warp] = incl;
		__syncthreads();
		if (warp == 0) {
			uint32_t w = s_warp[lane];
			uint32_t wi = warp_incl_scan(w);
			s_warp[lane] = wi - w;
		}
		__syncthreads();
		uint32_t carry = s_carry;
		uint32_t g = carry + s_warp[warp] + incl - x;
		if (p < map_size) {
			hash2index[p] = x ? (int32_t)g : -1;
			if (x) pre_buffer[g] = f;
		}
		__syncthreads();
		if (threadIdx.x == 1023) s_carry = g + x;
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		pre_buffer[s_carry] = (int32_t)n;
		*groups = (int32_t)s_carry;
	}
}

cudaError_t launch_rest_index(const uint64_t* d_keys, uint64_t n, int suffix_bits, int map_size, int32_t* d_first,
                              int32_t* d_hash2index, int32_t* d_pre_buffer, int32_t* d_groups, cudaStream_t stream) {
	cudaError_t e = cudaMemsetAsync(d_first, 0xFF, (size_t)map_size * sizeof(int32_t), stream);
	if (e != cudaSuccess) return e;
	if (n) {
		uint64_t blocks = (n + 255) / 256;
		rest_first_kernel<<<(int)(blocks < 4096 ? blocks : 4096), 256, 0, stream>>>(d_keys, n, suffix_bits, d_first);
		note_launch();
	}
	rest_index_kernel<<<1, 1024, 0, stream>>>(d_first, map_size, n, d_hash2index, d_pre_buffer, d_groups);
	note_launch();
	return cudaGetLastError();
}

// =========================================================================================
// coupled arrays: on-disk (bit_array_1, bit_array_2 separate) <-> device (interleaved cells)
// =========================================================================================
__global__ void split_cells_kernel(const unsigned long long* __restrict__ cells, uint64_t n, uint32_t* __restrict__ val,
                                   uint32_t* __restrict__ tag) {
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		unsigned long long c = cells[i];
		val[i] = (uint32_t)c;
		tag[i] = (uint32_t)(c >> 32);
	}
}
__global__ void merge_cells_kernel(const uint32_t* __restrict__ val, const uint32_t* __restrict__ tag, uint64_t n,
                                   unsigned long long* __restrict__ cells) {
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		cells[i] = ((unsigned long long)tag[i] << 32) | val[i];
}

cudaError_t launch_split_cells(const unsigned long long* d_cells, uint64_t n_words, uint32_t* d_val, uint32_t* d_tag, cudaStream_t stream) {
	if (n_words == 0) return cudaSuccess;
	uint64_t blocks = (n_words + 255) / 256;
	split_cells_kernel<<<(int)(blocks < 65535 ? blocks : 65535), 256, 0, stream>>>(d_cells, n_words, d_val, d_tag);
	note_launch();
	return cudaGetLastError();
}
cudaError_t launch_merge_cells(const uint32_t* d_val, const uint32_t* d_tag, uint64_t n_words, unsigned long long* d_cells, cudaStream_t stream) {
	if (n_words == 0) return cudaSuccess;
	uint64_t blocks = (n_words + 255) / 256;
	merge_cells_kernel<<<(int)(blocks < 65535 ? blocks : 65535), 256, 0, stream>>>(d_val, d_tag, n_words, d_cells);
	note_launch();
	return cudaGetLastError();
}

}  // namespace kmx
