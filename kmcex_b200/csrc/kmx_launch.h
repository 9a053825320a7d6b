// kmx_launch.h -- host-callable launchers of the kernels in kmx_query.cu / kmx_build.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "kmx_device.cuh"

namespace kmx {

// every kernel launch of the library is counted (bench.py reports the number inside its timed region)
void note_launch(int n = 1);

// ---- query (kmx_query.cu) ------------------------------------------------------------------
struct DeferredQuery {               // a query whose answer needs its 8 neighbours
	uint64_t kmer;                   // canonical form
	uint32_t index;                  // position in the batch
	uint32_t pad;
};
// d_defer needs room for n entries (worst case: every query is deferred); d_defer_n is one counter
// d_counters: two words, [0] deferred queries, [1] dirty ASCII queries
cudaError_t launch_query_packed(const DevModel& m, const uint64_t* d_kmers, size_t n, int32_t* d_out, int32_t* d_path,
                                DeferredQuery* d_defer, unsigned int* d_counters, int sm_count, cudaStream_t stream);
// ASCII batches: strings with a byte outside "ACGT" (N, lower case ...) are listed in d_dirty (room for n indices) and answered
// by a kernel that hashes the raw bytes the way the reference does (tools.hpp:160-167)
cudaError_t launch_query_ascii(const DevModel& m, const char* d_flat, size_t stride, size_t n, int32_t* d_out, uint64_t* d_packed,
                               DeferredQuery* d_defer, uint32_t* d_dirty, unsigned int* d_counters, int sm_count, cudaStream_t stream);
// bucket index + false-hit table of the rest table (R.keys/hash2index/pre_buffer/fine_bits set)
cudaError_t launch_rest_side_tables(const DevRest& R, int map_size, uint32_t* d_fine, uint64_t* d_quirk_suffix, uint32_t* d_quirk_index,
                                    cudaStream_t stream);

// ---- build (kmx_build.cu) ------------------------------------------------------------------
constexpr int kTile = 2048;          // records decoded per block step (256 threads x 8)

struct CountOut {                    // device-resident result of the counting pass
	unsigned long long class_count[kMaxBf];   // kmer_counts[c - ci]          kmodel.hpp:423-434
	unsigned long long listed;                // records passing the [min_count,max_count] filter
	unsigned long long array_bound;           // listed records with count >= ci + bf_num
	unsigned long long bad_count;             // listed records with count < ci or > cs (reference: out of bounds)
};

// pass 1 over the tiles [tile_first, tile_end): histogram of the low-count classes + array-bound records per tile
// (d_tile_cnt[tile - tile_first])
cudaError_t launch_count(const DevDb& db, int ci, int cs, int bf_num, CountOut* d_out, uint32_t* d_tile_cnt, uint64_t tile_first,
                         uint64_t tile_end, int sm_count, cudaStream_t stream);
// exclusive scan of the per-tile counts into 64-bit offsets
cudaError_t launch_tile_scan(const uint32_t* d_tile_cnt, uint64_t n_tiles, uint64_t* d_tile_off, cudaStream_t stream);
constexpr int kMaxRanks = 8;

// Where the array-bound k-mers go.  The stream position g of an item (file order, kmodel.hpp:508-518) fixes its
// bucket B = g >> 18, i.e. batch B / n_bits and bucket i = B % n_bits, whose round-0 array is i.  In a team build array i
// lives on rank i % n_active: the decoding rank stores the item straight into that owner's shard of the stream
// (NVLink peer store), where the owner keeps its buckets back to back.  One GPU: n_active = 1, the shard is the stream.
struct ItemRoute {
	uint64_t* kmer[kMaxRanks];        // item shard of every array owner as mapped on this device
	uint32_t* occ[kMaxRanks];
	unsigned long long base;          // stream position of the first item of this launch's tile range
	int n_active, n_bits;
};
__host__ __device__ inline uint32_t buckets_per_batch(int owner, int n_active, int n_bits) {
	return (uint32_t)((n_bits - owner + n_active - 1) / n_active);     // #{i < n_bits : i % n_active == owner}
}
// stream position g -> the rank that owns the item's round-0 array and the item's index in that owner's shard, where the
// owner keeps its buckets back to back, batch after batch
__host__ __device__ inline void route_item(unsigned long long g, int n_active, int n_bits, int* owner, unsigned long long* at) {
	const uint32_t B = (uint32_t)(g >> kBucketLog), c = (uint32_t)g & (kBucket - 1);
	const uint32_t batch = B / (uint32_t)n_bits, i = B - batch * (uint32_t)n_bits;
	*owner = (int)(i % (uint32_t)n_active);
	*at = ((unsigned long long)(batch * buckets_per_batch(*owner, n_active, n_bits) + i / (uint32_t)n_active) << kBucketLog) | c;
}

// pass 2 over the tiles [tile_first, tile_end): decode again; Bloom-bound k-mers are OR-ed into the filters of `m`, array-bound
// k-mers are written (file order preserved) to the item stream through `route`; d_tile_off[tile - tile_first] = items before
// the tile within the range
cudaError_t launch_encode(const DevDb& db, const DevModel& m, const uint64_t* d_tile_off, const ItemRoute& route, uint64_t tile_first,
                          uint64_t tile_end, int sm_count, cudaStream_t stream);
// plain listing (kmx_db_list): every listed record, file order, compacted
cudaError_t launch_list(const DevDb& db, const uint64_t* d_tile_off, uint64_t* d_kmers, uint32_t* d_counts, int sm_count,
                        cudaStream_t stream);
cudaError_t launch_list_count(const DevDb& db, uint32_t* d_tile_cnt, int sm_count, cudaStream_t stream);

// greedy coupled-array insert: persistent cooperative kernel over batches [first, first+count)

struct InsertCtl {                    // lives in device memory, survives across launches
	unsigned long long rest_n;        // survivors appended to the rest list so far
	unsigned long long attempts, accepted, iterations;
	unsigned int list_n[3];           // entries of the undecided-item lists (the classic iterations use two, the merged passes rotate three)
	unsigned int error;               // non-zero: iteration cap hit / rest overflow
	unsigned int nfail[2][kMaxArrays];   // survivors of each bucket after a round, by round parity (peers write it too)
	unsigned int seq;                 // rounds completed: the cross-GPU barrier counts with it
	unsigned int bar;                 // arrivals at the grid barrier (kmx_gridbar.cuh), never reset
	unsigned int epoch;               // reservation epoch (keys of newer epochs are smaller)
	unsigned int pad0;
	unsigned long long slot0_kmer[kMaxArrays];   // buffer slot 0 of each bucket after the last full batch
	unsigned int slot0_occ[kMaxArrays];
	unsigned int slot0_valid[kMaxArrays];
	unsigned long long phase_cycles[12];  // SM cycles seen by thread 0 of block 0: [0] phase 0, [1] phase 1, [3] contested passes, [5] place, [6] move,
	                                      // [7] cross-GPU wait; the part of it spent in its own loop (the rest is barrier wait): [2] phase 0, [4] phase 1,
	                                      // [8] place, [9] move, [10] contested passes; [11] contested passes run
};

struct InsertArgs {
	const uint64_t* item_kmer;        // array-bound stream, file order (team build: this owner's buckets back to back)
	const uint32_t* item_occ;
	unsigned long long n_items;       // items of the whole stream (all owners)
	uint64_t* buf_kmer[2];            // [n_bits * kBucket] ping-pong survivor buffers
	uint32_t* buf_occ[2];
	uint32_t* status;                 // [n_bits * kBucket] state<<30 | reserve mask
	uint32_t* excl_rank;              // [n_bits * kBucket]
	uint32_t* holepos;                // [n_bits * kBucket]
	uint32_t* tile_fail;              // [n_bits * kBucket / 256]
	uint32_t* list[3];                // [n_bits * kBucket] ids still undecided (two ping-pong, three rotating for the merged passes)
	uint32_t* resv;                   // [n_bits][2 tables][2 * resv_slots]
	uint32_t resv_slots;              // power of two
	uint32_t* claim;                  // [n_bits][2][2^claim_log2 / 32] claim bitmaps (position, wanted value)
	uint32_t claim_log2;              // bits per bitmap (upper bound; a round uses a prefix sized to its items)
	int claim_first;                  // 1: phase 0 claims blindly, phase 1 reads and commits back to back (HBM-resident arrays)
	int merged;                       // 1: contested items are resolved by merged reserve/commit passes (one barrier per iteration)
	int stream_cells;                 // the arrays are far larger than the L2: L2 evict-first policy for 1 = cell loads, 2 = cell reductions, 4 = km_back reductions; 8 = evict-last for km_back reductions
	uint64_t* rest_kmer;              // survivors of all batches
	uint32_t* rest_occ;
	unsigned long long rest_cap;
	InsertCtl* ctl;
	// array-owner decomposition over GPUs (SURVEY.md 8e option A): array a lives on active rank a % n_active;
	// the peer_* pointers are this device's views of every active rank's exchange buffers (index = rank)
	int rank, n_active;
	uint64_t* peer_buf_kmer[2][kMaxRanks];
	uint32_t* peer_buf_occ[2][kMaxRanks];
	InsertCtl* peer_ctl[kMaxRanks];
	uint32_t* peer_flags[kMaxRanks];  // peer_flags[p][r]: rounds rank r has completed, as visible on rank p
	unsigned long long first_batch, n_batches;
	unsigned int max_iterations;      // safety cap per round
	int phase_round;                  // diagnostic: phase_cycles count only round t == phase_round of every batch (-1: all rounds)
};

// ---- exchanges between the GPUs of a team through peer-mapped memory (kmx_dist.cu) ----------------
struct TeamLink {                     // who is who, and the flag barrier the exchange kernels share
	int rank, world;
	uint32_t* flags[kMaxRanks];       // flags[p][r]: barriers rank r has entered, as visible on rank p
	uint32_t seq;                     // barriers completed before this launch
	unsigned int* error;              // local word, set to 3 when a peer does not show up
};
struct PullSeg {
	const uint4* src;                 // on the owner
	uint4* dst;                       // here
	unsigned long long n_vec;
};
// bitwise-OR all-reduce of a filter region over the ranks' copies, then (optionally) this rank copies the segments it does
// not own from their owners: cross-GPU barrier, reduce-scatter, all-gather, pulls, cross-GPU barrier (link.seq + 2 afterwards)
struct OrReduceArgs {
	TeamLink link;
	unsigned long long n_vec;         // 16-byte vectors in the region
	uint4* base[kMaxRanks];           // every rank's copy of the region as mapped on this device (index = rank)
	int n_pull;
	PullSeg pull[kMaxArrays];
};
cudaError_t launch_or_allreduce(const OrReduceArgs& a, int sm_count, cudaStream_t stream);

// histogram of the survivors' rest-table prefixes (rest.hpp:95-105: one group per prefix); *d_n is read on the device
cudaError_t launch_prefix_hist(const uint64_t* d_keys, const unsigned long long* d_n, int suffix_bits, uint32_t* d_hist, int sm_count,
                               cudaStream_t stream);
// sharded rest build: the survivors of every owner whose prefix falls in [prefix_lo, prefix_hi) -> this rank's local list
struct RestGatherArgs {
	int n_owners;
	const uint64_t* kmer[kMaxRanks];
	const uint32_t* occ[kMaxRanks];
	unsigned long long n[kMaxRanks];
	uint32_t prefix_lo, prefix_hi;
	int suffix_bits;
	uint64_t* out_kmer;
	uint32_t* out_occ;
	unsigned long long* out_n;        // zeroed by the launcher
	unsigned long long cap;
};
cudaError_t launch_rest_gather(const RestGatherArgs& a, int sm_count, cudaStream_t stream);
// this rank's sorted run -> every rank's rest table at `offset`; ends with a cross-GPU barrier (link.seq + 1 afterwards)
struct RestPushArgs {
	TeamLink link;
	const uint64_t* keys;
	const uint32_t* counts;
	unsigned long long n, offset;
	uint64_t* dst_keys[kMaxRanks];
	int32_t* dst_counts[kMaxRanks];
};
cudaError_t launch_rest_push(const RestPushArgs& a, int sm_count, cudaStream_t stream);

cudaError_t insert_grid_size(int* blocks_out, int sm_count);
cudaError_t launch_insert(const DevModel& m, const InsertArgs& a, int grid_blocks, cudaStream_t stream);

// rest table: survivors -> sorted keys + group index (rest.hpp:95-135).  The sort is the library's own LSD radix sort
// (kmx_sort.cu): 8-bit digits, per-chunk histograms, stable block-level scatter.
size_t radix_sort_temp_bytes(size_t n);
cudaError_t launch_radix_sort_pairs(void* d_temp, const uint64_t* d_keys_in, uint64_t* d_keys_out, const uint32_t* d_vals_in,
                                    uint32_t* d_vals_out, size_t n, int key_bits, cudaStream_t stream);
cudaError_t launch_rest_index(const uint64_t* d_keys, uint64_t n, int suffix_bits, int map_size, int32_t* d_first,
                              int32_t* d_hash2index, int32_t* d_pre_buffer, int32_t* d_groups, cudaStream_t stream);

// (de)interleave of the coupled arrays between the on-disk and the device layout
cudaError_t launch_split_cells(const unsigned long long* d_cells, uint64_t n_words, uint32_t* d_val, uint32_t* d_tag, cudaStream_t stream);
cudaError_t launch_merge_cells(const uint32_t* d_val, const uint32_t* d_tag, uint64_t n_words, unsigned long long* d_cells, cudaStream_t stream);

}  // namespace kmx
