// kmx_internal.h -- host-side objects and helpers shared by the translation units of libkmx.so
// (kmx_host.cu: model lifetime, KMC database, single-GPU build, save/load, query pipelines;
//  kmx_team.cu: ONE model built by several GPUs).  Nothing here is part of the C ABI (include/kmx.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <chrono>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/kmx.h"
#include "kmx_device.cuh"
#include "kmx_launch.h"

namespace kmx {

// ---- errors -------------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);        // stores the message for kmx_last_error(), returns code
const char* last_error();
int last_error_code();                                 // code of the last failure on this thread

#define CU(call)                                                                                         \
	do {                                                                                                 \
		cudaError_t e__ = (call);                                                                        \
		if (e__ != cudaSuccess)                                                                          \
			return kmx::set_error(KMX_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
	} while (0)

// KMX_TRACE=1: host-side timeline of a build on stderr (milliseconds since the call started)
bool trace_on();
#define TRACE(t0, what)                                                                                             \
	do {                                                                                                           \
		if (kmx::trace_on())                                                                                       \
			fprintf(stderr, "[kmx] %8.3f ms  %s\n",                                                                \
			        1e3 * std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - (t0)).count(), what); \
	} while (0)

// ---- device memory --------------------------------------------------------------------------
// Stream-ordered allocations from the device's default memory pool, which is told to keep freed blocks.
int dev_alloc_impl(void** p, size_t bytes, cudaStream_t s);
template <class T>
static inline int dev_alloc(T** p, size_t bytes, cudaStream_t s) { return dev_alloc_impl((void**)p, bytes, s); }
void dev_free(void* p, cudaStream_t s);
#define DA(ptr, bytes, stream)                                   \
	do {                                                         \
		int rc__ = kmx::dev_alloc(ptr, bytes, stream);           \
		if (rc__) return rc__;                                   \
	} while (0)

// frees a list of stream-ordered allocations when it goes out of scope (early returns of a multi-allocation function)
struct DevScope {
	cudaStream_t s;
	std::vector<void*> ptrs;
	explicit DevScope(cudaStream_t st) : s(st) {}
	template <class T>
	int alloc(T** p, size_t bytes) {
		int rc = dev_alloc(p, bytes, s);
		if (!rc) ptrs.push_back((void*)*p);
		return rc;
	}
	void keep(void* p) {                       // ownership moves elsewhere
		for (auto& q : ptrs)
			if (q == p) q = nullptr;
	}
	~DevScope() {
		for (void* p : ptrs) dev_free(p, s);
	}
};

// cudaMalloc blocks that other GPUs / processes map (exchange slabs); cached by size for the life of the process
int slab_acquire(void** out, size_t bytes, int device);
void slab_release(void* ptr, size_t bytes, int device);

// the device objects are created on: per thread when the thread called kmx_set_device, else the process-wide one
int current_device();
void set_thread_device(int ordinal);
int require_gpu(int* sm_count);

// ---- per-device execution context -----------------------------------------------------------
// Streams, events, pinned scratch and the query staging buffers.  Creating these costs milliseconds, so contexts are
// pooled for the life of the process: a model borrows one at its first device use and returns it when destroyed.
struct DevCtx {
	int device = 0;
	cudaStream_t stream = nullptr, stream2 = nullptr;
	cudaStream_t reader[16] = {};
	cudaEvent_t reader_ev[16][2] = {};
	cudaEvent_t ev_build[7] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };   // [6]: team build, start of the rest stage
	struct Pinned {                   // small device->host results, pinned so the copies are truly asynchronous
		CountOut count;
		InsertCtl ctl;
		int32_t groups;
		unsigned long long rest_n[kMaxRanks];
	}* h_pinned = nullptr;
	// pinned staging for host-pointer queries (two slots)
	void* h_in[2] = { nullptr, nullptr };
	int32_t* h_out[2] = { nullptr, nullptr };
	void* d_in[2] = { nullptr, nullptr };
	int32_t* d_out[2] = { nullptr, nullptr };
	uint64_t* d_pack[2] = { nullptr, nullptr };
	DeferredQuery* d_defer[2] = { nullptr, nullptr };
	uint32_t* d_dirty[2] = { nullptr, nullptr };
	unsigned int* d_defer_n[2] = { nullptr, nullptr };      // [0] deferred queries, [1] dirty (non-ACGT) queries
	size_t stage_bytes = 0, stage_items = 0;
	cudaEvent_t ev_done[2] = { nullptr, nullptr };
	std::mutex query_mu;              // host-pointer queries share the staging slots: one batch at a time per model
};
int ctx_acquire(DevCtx** out);
void ctx_release(DevCtx* c);

}  // namespace kmx

// ---- the opaque objects of the C ABI ---------------------------------------------------------
struct kmx_db {
	kmx_db_info_t info;
	std::vector<uint64_t> lut;        // lut_entries + 1 (guard = total + 1, kmc_file.cpp:223)
	int fd = -1;                      // .kmc_suf, kept open until the records are on the device
	uint64_t rec_lo = 0, rec_hi = 0;  // records resident on the device: [rec_lo, rec_hi)
	size_t suf_alloc = 0;
	uint8_t* d_suf = nullptr;         // first byte = record rec_lo
	uint64_t* d_lut = nullptr;
	int device = 0, sm_count = 0;
	float ms_upload = 0;
	// random access (kmx_ra.cu): the signature map is read from .kmc_pre at the first CheckKmer, not at open
	std::string pre_name;
	uint64_t sig_offset = 0;          // byte offset of the signature map in .kmc_pre
	bool both_strands = true;         // kmc_file.cpp:208-209
	uint32_t orig_min_count = 0, orig_max_count = 0;   // the header's counter range (SetMinCount / SetMaxCount narrow info.min/max_count)
	std::mutex ra_mu;                 // guards the lazy creation of d_sigmap
	uint32_t* d_sigmap = nullptr;     // [4^signature_len + 1] bin of every signature
};

namespace kmx {

struct RestHost {
	int32_t k = 0, pre_len = 0, map_size = 0, pre_buffer_size = 0;
	uint64_t suff_bin_size = 0, count = 0;
};

struct TeamState;                     // kmx_team.cu

// what a build carries from one stage to the next
struct BuildState {
	std::chrono::high_resolution_clock::time_point wall0;
	uint64_t n_items = 0, n_batches = 0, rest_cap = 0;
	uint64_t* d_item_kmer = nullptr;
	uint32_t* d_item_occ = nullptr;
	InsertArgs a = {};
	int grid = 0;
	float ms_upload = 0;
	TeamState* team = nullptr;        // non-null while a team build is in flight
};

}  // namespace kmx

struct kmx_model {
	int ci = 1, cs = 1023, n_hash = 7, n_bits = 5, bf_num = 1, k = 0;
	int device = 0, sm_count = 0;
	bool built = false;
	uint64_t total_kmers = 0, km_kmers = 0, kmer_counts[3] = { 0, 0, 0 };
	uint64_t bytes[8] = { 0 };        // see kmx_host_sizes
	std::vector<int32_t> occ2bin, bin2mean;
	// device
	uint32_t* d_bf[3] = { nullptr, nullptr, nullptr };
	uint32_t* d_bf_back[3] = { nullptr, nullptr, nullptr };
	uint32_t* d_km_back = nullptr;
	unsigned long long* d_cells[kmx::kMaxArrays] = { nullptr };
	uint16_t* d_occ2bin = nullptr;
	int32_t* d_bin2mean = nullptr;
	int32_t* d_hash2index = nullptr;
	int32_t* d_pre_buffer = nullptr;
	uint64_t* d_rest_keys = nullptr;
	int32_t* d_rest_counts = nullptr;
	uint32_t* d_fine = nullptr;
	uint64_t* d_quirk_suffix = nullptr;
	uint32_t* d_quirk_index = nullptr;
	int fine_bits = 8;
	// team build: filters + coupled arrays live in one peer-mapped slab, the rest keys/counts in another (owned by the model)
	void* mslab = nullptr;
	size_t mslab_bytes = 0;
	void* rslab = nullptr;
	size_t rslab_bytes = 0;
	kmx::RestHost rest;
	kmx::DevModel dm;
	kmx_info_t info;
	kmx::DevCtx* x = nullptr;         // borrowed execution context (streams, events, staging)
	kmx::BuildState bs;
	std::vector<uint16_t> occ2bin16;
	// single-process multi-GPU (kmx_set_devices / KMX_GPUS): replicas of this model on the other devices
	std::vector<kmx_model*> replicas;
};

namespace kmx {

inline size_t pad8(uint64_t bytes) { return (size_t)((bytes + 7) & ~7ULL) + 8; }
inline uint64_t cell_words(uint64_t km_byte_size) { return (km_byte_size + 3) / 4; }
inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

void model_sizes(const uint64_t kmer_counts[3], int bf_num, uint64_t km_kmers, int n_hash, uint64_t bytes[8]);
int rest_prefix_len(int k);
int model_attach_device(kmx_model* m);
int check_model_sizes(kmx_model* m);                  // the reference's failure corners on degenerate sizes
void fill_dev_model(kmx_model* m);
void fill_info(kmx_model* m);
void free_model_device(kmx_model* m);
void build_state_free(kmx_model* m);
DevDb dev_db(const kmx_db* db);
int db_upload_range(kmx_db* db, uint64_t rec_lo, uint64_t rec_hi, int reader_threads);
int build_rest_side_tables(kmx_model* m);
int build_stage_insert_setup(kmx_model* m, int rank, int n_active, bool team);
int build_stage_insert_run(kmx_model* m);
void team_state_free(kmx_model* m);
int team_build_in_process(kmx_model* m, const char* db_base, const std::vector<int>& devices);
const std::vector<int>& team_devices();               // kmx_set_devices / KMX_GPUS

}  // namespace kmx
