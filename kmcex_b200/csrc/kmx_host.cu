// kmx_host.cu -- host side of libkmx.so: KMC header parse, model object, build orchestration,
// save/load in the reference's on-disk layout, query pipelines, and the C ABI of include/kmx.h.
//
// Reference interfaces mirrored (file:line relative to the reference root):
//   kmodel.hpp:45-55,674-696  get_model / KModel ctor      kmodel.hpp:57-86    KModel::init
//   kmodel.hpp:173-235        save / load                  kmodel.hpp:402-456  size formulas
//   occu_bin.hpp:27-83        OccuBin                      rest.hpp:163-221    rest.bin I/O
//   kmc_file.cpp:66-99,132-171,177-235  OpenForListing / header + LUT parse
#include <cuda_runtime.h>
#include <errno.h>
#include <fcntl.h>
#include <unistd.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <algorithm>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "../../include/kmx.h"
#include "kmx_device.cuh"
#include "kmx_launch.h"

using namespace kmx;

// =========================================================================================
// errors
// =========================================================================================
static thread_local char g_err[512] = "";
static int g_device = 0;

// KMX_TRACE=1: host-side timeline of a build on stderr (milliseconds since the call started)
static bool trace_on() {
	static int on = -1;
	if (on < 0) on = getenv("KMX_TRACE") ? 1 : 0;
	return on == 1;
}
#define TRACE(t0, what)                                                                                             \
	do {                                                                                                           \
		if (trace_on())                                                                                            \
			fprintf(stderr, "[kmx] %8.3f ms  %s\n",                                                                \
			        1e3 * std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - (t0)).count(), what); \
	} while (0)

static int fail(int code, const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}

namespace kmx {
// error reporting for the other translation units of the library
int set_error(int code, const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}
}  // namespace kmx

#define CU(call)                                                                                         \
	do {                                                                                                 \
		cudaError_t e__ = (call);                                                                        \
		if (e__ != cudaSuccess) return fail(KMX_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
	} while (0)

// Stream-ordered allocations from the device's default memory pool, which is told to keep
// freed blocks: a rebuild (or the next query staging) reuses them instead of paying
// cudaMalloc/cudaFree (hundreds of microseconds each, and cudaFree synchronises the device).
static int dev_alloc_impl(void** p, size_t bytes, cudaStream_t s) {
	static bool pool_ready[64] = { false };
	int dev = 0;
	CU(cudaGetDevice(&dev));
	if (dev < 64 && !pool_ready[dev]) {
		cudaMemPool_t pool;
		CU(cudaDeviceGetDefaultMemPool(&pool, dev));
		uint64_t keep = ~0ULL;
		CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
		pool_ready[dev] = true;
	}
	CU(cudaMallocAsync(p, bytes ? bytes : 8, s));
	return KMX_OK;
}
template <class T>
static int dev_alloc(T** p, size_t bytes, cudaStream_t s) { return dev_alloc_impl((void**)p, bytes, s); }
static void dev_free(void* p, cudaStream_t s) {
	if (p) cudaFreeAsync(p, s);
}
#define DA(ptr, bytes, stream)                                   \
	do {                                                         \
		int rc__ = dev_alloc(ptr, bytes, stream);                \
		if (rc__) return rc__;                                   \
	} while (0)

static int require_gpu(int* sm_count) {
	// cudaGetDeviceProperties costs milliseconds: ask once per device
	static int cached_sm[64] = { 0 };
	static int cached_n = -1;
	if (cached_n < 0) {
		int n = 0;
		cudaError_t e = cudaGetDeviceCount(&n);
		if (e != cudaSuccess || n <= 0) {
			cudaGetLastError();
			return fail(KMX_ENOGPU, "no usable CUDA device (libkmx has no CPU path): %s", e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
		}
		cached_n = n;
	}
	if (g_device >= cached_n || g_device >= 64) return fail(KMX_ENOGPU, "device %d requested, %d present", g_device, cached_n);
	CU(cudaSetDevice(g_device));
	if (cached_sm[g_device] == 0) {
		cudaDeviceProp prop;
		CU(cudaGetDeviceProperties(&prop, g_device));
		if (prop.major < 10) return fail(KMX_ENOGPU, "device %d is sm_%d%d; this library is built for sm_100a only", g_device, prop.major, prop.minor);
		cached_sm[g_device] = prop.multiProcessorCount;
	}
	if (sm_count) *sm_count = cached_sm[g_device];
	return KMX_OK;
}

// =========================================================================================
// OccuBin (occu_bin.hpp:27-83) as two lookup tables
// =========================================================================================
// Three zones over the occurrence axis: [0, E1) one bin per value; then 2^(H-1) bins of width
// 3 whose mean is first+1; then 2^(H-2) bins of width W = (max_counter - zone3_start) / 2^(H-2)
// whose mean is (2*first + W)/2; what is left shares the last bin.  bin -> mean keeps the FIRST
// mean registered for a bin (unordered_map::insert), missing bins read as 0.
static int occubin_tables(int max_counter, int n_hash, std::vector<int32_t>& occ2bin, std::vector<int32_t>& bin2mean) {
	if (n_hash < 3 || n_hash > kMaxHash || max_counter < 1) return KMX_EARG;
	const int e3 = 1 << n_hash, e1 = e3 / 4, e2 = e1 + e3 / 2;
	const int z3 = e1 + 3 * (e3 / 2);
	if (z3 > max_counter) return KMX_ERANGE;       // the reference writes past occ_bin_meta here
	const int w3 = (max_counter - z3) / (e3 / 4);
	const int z4 = z3 + w3 * (e3 / 4);
	occ2bin.assign(max_counter, 0);
	bin2mean.assign(e3, 0);
	std::vector<char> seen(e3, 0);
	for (int b = 0; b < e1; b++) bin2mean[b] = b;
	for (int occ = 0; occ < max_counter; occ++) {
		int bin, mean;
		if (occ < e1) {
			occ2bin[occ] = occ;
			continue;
		} else if (occ < z3) {
			int idx = (occ - e1) / 3;
			bin = e1 + idx;
			mean = e1 + 3 * idx + 1;
		} else if (occ < z4) {
			int idx = (occ - z3) / w3;
			bin = e2 + idx;
			mean = (2 * (z3 + w3 * idx) + w3) / 2;
		} else {
			bin = e3 - 1;
			mean = (2 * z4 - w3) / 2;
		}
		occ2bin[occ] = bin;
		if (!seen[bin]) {
			seen[bin] = 1;
			bin2mean[bin] = mean;
		}
	}
	return KMX_OK;
}

// kmodel.hpp:402-456.  The Bloom size is evaluated in double exactly as the reference writes it.
static void model_sizes(const uint64_t kmer_counts[3], int bf_num, uint64_t km_kmers, int n_hash, uint64_t bytes[8]) {
	const int hb = n_hash - 1, hk = n_hash - 2;
	for (int i = 0; i < 8; i++) bytes[i] = 0;
	for (int i = 0; i < bf_num; i++) {
		bytes[i] = (uint64_t)(kmer_counts[i] / 5.5 * hb);
		bytes[3 + i] = (kmer_counts[i] >> 3) * (uint64_t)hk;
	}
	bytes[6] = (km_kmers >> 4) * (uint64_t)n_hash;
	bytes[7] = (km_kmers >> 4) * (uint64_t)hk;
}

static int rest_prefix_len(int k) {              // rest.hpp:78-83
	for (int i = 7; i >= 3; i--)
		if ((k - i) % 4 == 0) return i;
	return 3;
}

// =========================================================================================
// objects
// =========================================================================================
// Per-device execution context: streams, events, pinned scratch and the query staging buffers.
// Creating these costs milliseconds (cudaHostAlloc, cudaStreamCreate), so contexts are pooled
// for the life of the process: a model borrows one at its first device use and returns it when
// destroyed; a database upload borrows one for the duration of the copy.
struct DevCtx {
	int device = 0;
	cudaStream_t stream = nullptr, stream2 = nullptr;
	cudaStream_t reader[16] = {};
	cudaEvent_t reader_ev[16][2] = {};
	cudaEvent_t ev_build[6] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
	struct Pinned {                   // small device->host results, pinned so the copies are truly asynchronous
		CountOut count;
		InsertCtl ctl;
		int32_t groups;
	}* h_pinned = nullptr;
	// pinned staging for host-pointer queries (two slots)
	void* h_in[2] = { nullptr, nullptr };
	int32_t* h_out[2] = { nullptr, nullptr };
	void* d_in[2] = { nullptr, nullptr };
	int32_t* d_out[2] = { nullptr, nullptr };
	uint64_t* d_pack[2] = { nullptr, nullptr };
	DeferredQuery* d_defer[2] = { nullptr, nullptr };
	unsigned int* d_defer_n[2] = { nullptr, nullptr };
	size_t stage_bytes = 0, stage_items = 0;
	cudaEvent_t ev_done[2] = { nullptr, nullptr };
	std::mutex query_mu;              // host-pointer queries share the staging slots: one batch at a time per model
};

namespace {
std::mutex g_ctx_mu;
std::vector<DevCtx*> g_ctx_free;
}

static int ctx_acquire(DevCtx** out) {
	{
		std::lock_guard<std::mutex> lock(g_ctx_mu);
		for (size_t i = 0; i < g_ctx_free.size(); i++) {
			if (g_ctx_free[i]->device == g_device) {
				*out = g_ctx_free[i];
				g_ctx_free.erase(g_ctx_free.begin() + i);
				return KMX_OK;
			}
		}
	}
	DevCtx* c = new DevCtx();
	c->device = g_device;
	CU(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	CU(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
	for (auto& r : c->reader) CU(cudaStreamCreateWithFlags(&r, cudaStreamNonBlocking));
	for (auto& r : c->reader_ev)
		for (auto& e : r) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	for (auto& e : c->ev_build) CU(cudaEventCreate(&e));
	for (auto& e : c->ev_done) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
	CU(cudaHostAlloc((void**)&c->h_pinned, sizeof(DevCtx::Pinned), cudaHostAllocDefault));
	*out = c;
	return KMX_OK;
}

static void ctx_release(DevCtx* c) {
	if (!c) return;
	cudaStreamSynchronize(c->stream);
	cudaStreamSynchronize(c->stream2);
	std::lock_guard<std::mutex> lock(g_ctx_mu);
	g_ctx_free.push_back(c);
}

struct kmx_db {
	kmx_db_info_t info;
	std::vector<uint64_t> lut;        // lut_entries + 1 (guard = total + 1, kmc_file.cpp:223)
	int fd = -1;                      // .kmc_suf, kept open until the records are on the device
	size_t suf_alloc = 0;
	uint8_t* d_suf = nullptr;
	uint64_t* d_lut = nullptr;
	int device = 0, sm_count = 0;
	float ms_upload = 0;
};

struct RestHost {
	int32_t k = 0, pre_len = 0, map_size = 0, pre_buffer_size = 0;
	uint64_t suff_bin_size = 0, count = 0;
};

// what a build carries from one stage to the next (kmx_init_from_db runs the stages back to back)
struct BuildState {
	std::chrono::high_resolution_clock::time_point wall0;
	uint64_t n_items = 0, n_batches = 0, rest_cap = 0;
	uint64_t* d_item_kmer = nullptr;
	uint32_t* d_item_occ = nullptr;
	InsertArgs a = {};
	int grid = 0;
	bool worst_case = true;
	float ms_upload = 0;
	// multi-GPU: one cudaMalloc slab with everything peers write into, and the peers' slabs as mapped here
	void* slab = nullptr;
	size_t slab_bytes = 0, off[6] = { 0 };
	uint32_t* flags = nullptr;
	void* peer_slab[kMaxRanks] = { nullptr };
	// multi-GPU: the Bloom filters and km_back live in a second cudaMalloc slab that every rank maps, so that the
	// partial filters can be OR-ed through peer memory (kmx_dist.cu).  Layout: [Bloom filters | km_back | flags | error]
	void* fslab = nullptr;
	size_t fslab_bytes = 0, f_bloom_bytes = 0, f_kmback_off = 0, f_kmback_bytes = 0, f_flags_off = 0;
	void* peer_fslab[kMaxRanks] = { nullptr };
	uint32_t fseq = 0;
	int world = 1, dist_rank = 0;
};

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
	unsigned long long* d_cells[kMaxArrays] = { nullptr };
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
	RestHost rest;
	DevModel dm;
	kmx_info_t info;
	DevCtx* x = nullptr;              // borrowed execution context (streams, events, staging)
	BuildState bs;
	std::vector<uint16_t> occ2bin16;
};

static size_t pad8(uint64_t bytes) { return (size_t)((bytes + 7) & ~7ULL) + 8; }
static uint64_t cell_words(uint64_t km_byte_size) { return (km_byte_size + 3) / 4; }

static void free_model_device(kmx_model* m) {
	cudaStream_t s = m->x->stream;
	for (int i = 0; i < 3; i++) {
		dev_free(m->d_bf[i], s);
		dev_free(m->d_bf_back[i], s);
		m->d_bf[i] = m->d_bf_back[i] = nullptr;
	}
	dev_free(m->d_km_back, s);
	m->d_km_back = nullptr;
	for (int i = 0; i < kMaxArrays; i++) {
		dev_free(m->d_cells[i], s);
		m->d_cells[i] = nullptr;
	}
	dev_free(m->d_hash2index, s);
	dev_free(m->d_pre_buffer, s);
	dev_free(m->d_rest_keys, s);
	dev_free(m->d_rest_counts, s);
	dev_free(m->d_fine, s);
	dev_free(m->d_quirk_suffix, s);
	dev_free(m->d_quirk_index, s);
	m->d_fine = nullptr;
	m->d_quirk_suffix = nullptr;
	m->d_quirk_index = nullptr;
	m->d_hash2index = m->d_pre_buffer = nullptr;
	m->d_rest_keys = nullptr;
	m->d_rest_counts = nullptr;
}

static int slab_acquire(void** out, size_t bytes, int device);

// allocate + zero every filter of the model from m->kmer_counts / m->km_kmers (kmodel.hpp:402-456);
// in_slab: Bloom filters and km_back are carved from one exportable cudaMalloc block (multi-GPU build)
static int alloc_filters(kmx_model* m, bool in_slab = false) {
	model_sizes(m->kmer_counts, m->bf_num, m->km_kmers, m->n_hash, m->bytes);
	for (int i = 0; i < m->bf_num; i++) {
		if (m->bytes[i] == 0 || m->bytes[3 + i] == 0)
			return fail(KMX_ERANGE, "count class %d holds %llu k-mers: the reference aborts on a zero-length Bloom filter (kmodel.hpp:413-417)",
			            m->ci + i, (unsigned long long)m->kmer_counts[i]);
	}
	if (m->bytes[6] == 0 || m->bytes[7] == 0)
		return fail(KMX_ERANGE, "%llu k-mers for the coupled arrays: the reference aborts on zero-length arrays (kmodel.hpp:443-447)",
		            (unsigned long long)m->km_kmers);
	if (in_slab) {
		BuildState& b = m->bs;
		auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
		size_t off = 0, o_bf[3] = { 0, 0, 0 }, o_bb[3] = { 0, 0, 0 };
		for (int i = 0; i < m->bf_num; i++) {
			o_bf[i] = off; off += up(pad8(m->bytes[i]));
			o_bb[i] = off; off += up(pad8(m->bytes[3 + i]));
		}
		b.f_bloom_bytes = off;
		b.f_kmback_off = off;
		b.f_kmback_bytes = up(pad8(m->bytes[7]));
		off += b.f_kmback_bytes;
		b.f_flags_off = off;
		b.fslab_bytes = off + 256;                           // kMaxRanks barrier counters + the error word
		int rc = slab_acquire(&b.fslab, b.fslab_bytes, m->device);
		if (rc) return rc;
		CU(cudaMemsetAsync(b.fslab, 0, b.fslab_bytes, m->x->stream));
		uint8_t* base = (uint8_t*)b.fslab;
		for (int i = 0; i < m->bf_num; i++) {
			m->d_bf[i] = (uint32_t*)(base + o_bf[i]);
			m->d_bf_back[i] = (uint32_t*)(base + o_bb[i]);
		}
		m->d_km_back = (uint32_t*)(base + b.f_kmback_off);
	} else {
		for (int i = 0; i < m->bf_num; i++) {
			DA(&m->d_bf[i], pad8(m->bytes[i]), m->x->stream);
			CU(cudaMemsetAsync(m->d_bf[i], 0, pad8(m->bytes[i]), m->x->stream));
			DA(&m->d_bf_back[i], pad8(m->bytes[3 + i]), m->x->stream);
			CU(cudaMemsetAsync(m->d_bf_back[i], 0, pad8(m->bytes[3 + i]), m->x->stream));
		}
		DA(&m->d_km_back, pad8(m->bytes[7]), m->x->stream);
		CU(cudaMemsetAsync(m->d_km_back, 0, pad8(m->bytes[7]), m->x->stream));
	}
	const uint64_t words = cell_words(m->bytes[6]);
	for (int i = 0; i < m->n_bits; i++) {
		DA(&m->d_cells[i], (words + 1) * 8, m->x->stream);
		CU(cudaMemsetAsync(m->d_cells[i], 0, (words + 1) * 8, m->x->stream));
	}
	return KMX_OK;
}

static void fill_dev_model(kmx_model* m) {
	DevModel& d = m->dm;
	memset(&d, 0, sizeof(d));
	d.k = m->k;
	d.n_hash = m->n_hash;
	d.n_bits = m->n_bits;
	d.bf_num = m->bf_num;
	d.ci = m->ci;
	d.cs = m->cs;
	d.hb = m->n_hash - 1;
	d.hk = m->n_hash - 2;
	d.end1 = (1 << m->n_hash) / 4;
	for (int i = 0; i < m->bf_num; i++) {
		d.bf[i].words = m->d_bf[i];
		d.bf[i].mod = make_fastmod(m->bytes[i] * 8);
		d.bf_back[i].words = m->d_bf_back[i];
		d.bf_back[i].mod = make_fastmod(m->bytes[3 + i] * 8);
	}
	d.km_back.words = m->d_km_back;
	d.km_back.mod = make_fastmod(m->bytes[7] * 8);
	d.arr_mod = make_fastmod(m->bytes[6] * 8);
	for (int i = 0; i < m->n_bits; i++) {
		d.cells[i] = m->d_cells[i];
		for (int j = 0; j < m->n_hash; j++) d.arr_seed[i][j] = h_seeds[(i * m->n_hash + j) % 128];   // kmodel.hpp:450-453
	}
	d.occ2bin = m->d_occ2bin;
	d.bin2mean = m->d_bin2mean;
	d.rest.hash2index = m->d_hash2index;
	d.rest.pre_buffer = m->d_pre_buffer;
	d.rest.keys = m->d_rest_keys;
	d.rest.counts = m->d_rest_counts;
	d.rest.count = m->rest.count;
	d.rest.k = m->rest.k;
	d.rest.suffix_bits = 2 * (m->rest.k - m->rest.pre_len);
	d.rest.suffix_mask = mask2(m->rest.k - m->rest.pre_len);
	d.rest.fine = m->d_fine;
	d.rest.quirk_suffix = m->d_quirk_suffix;
	d.rest.quirk_index = m->d_quirk_index;
	d.rest.fine_bits = m->fine_bits;
	d.rest.fine_shift = 2 * m->rest.k - m->fine_bits;
}

static void fill_info(kmx_model* m) {
	kmx_info_t& f = m->info;
	f.ci = m->ci; f.cs = m->cs; f.n_hash = m->n_hash; f.n_bits = m->n_bits; f.bf_num = m->bf_num; f.k = m->k;
	f.total_kmers = m->total_kmers;
	f.km_kmers = m->km_kmers;
	f.bf_kmers = 0;
	f.bf_bytes = 0;
	for (int i = 0; i < 3; i++) {
		f.kmer_counts[i] = m->kmer_counts[i];
		if (i < m->bf_num) {
			f.bf_kmers += m->kmer_counts[i];
			f.bf_bytes += m->bytes[i] + m->bytes[3 + i];
		}
	}
	f.km_bytes = 2ULL * m->n_bits * m->bytes[6];
	f.km_back_bytes = m->bytes[7];
	f.rest_kmers = m->rest.count;
	f.rest_bytes = m->rest.suff_bin_size + 4ULL * m->rest.count + 4ULL * m->rest.pre_buffer_size + 4ULL * m->rest.map_size;
}

// =========================================================================================
// process-wide
// =========================================================================================
extern "C" const char* kmx_last_error(void) { return g_err; }
extern "C" const char* kmx_version(void) { return "kmx 0.1 (sm_100a)"; }

extern "C" int kmx_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

extern "C" int kmx_set_device(int ordinal) {
	int n = kmx_device_count();
	if (ordinal < 0 || ordinal >= n) return fail(KMX_ENOGPU, "device %d requested, %d present", ordinal, n);
	g_device = ordinal;
	CU(cudaSetDevice(ordinal));
	return KMX_OK;
}

// =========================================================================================
// model lifetime
// =========================================================================================
extern "C" kmx_model* kmx_create(int ci, int cs, int n_hash, int n_bits) {
	if (ci < 1 || cs < ci || cs > 65535 || n_hash < 3 || n_hash > kMaxHash || n_bits < 1 || n_bits > kMaxArrays) {
		fail(KMX_EARG, "unsupported parameters ci=%d cs=%d n_hash=%d n_bits=%d (need 1<=ci<=cs<=65535, 3<=n_hash<=%d, 1<=n_bits<=%d)", ci, cs,
		     n_hash, n_bits, kMaxHash, kMaxArrays);
		return nullptr;
	}
	kmx_model* m = new kmx_model();
	m->ci = ci; m->cs = cs; m->n_hash = n_hash; m->n_bits = n_bits;
	m->bf_num = ci == 1 ? 1 : 3;                         // kmodel.hpp:50
	memset(&m->info, 0, sizeof(m->info));
	memset(&m->dm, 0, sizeof(m->dm));
	int rc = occubin_tables(cs + 1, n_hash, m->occ2bin, m->bin2mean);
	if (rc != KMX_OK) {
		fail(rc, "OccuBin(%d, %d): the reference indexes past its table for cs + 1 < %d (occu_bin.hpp:38-44)", cs + 1, n_hash,
		     (1 << n_hash) / 4 + 3 * (1 << n_hash) / 2);
		delete m;
		return nullptr;
	}
	fill_info(m);
	return m;
}

static void build_state_free(kmx_model* m);

static int model_attach_device(kmx_model* m) {
	if (m->x) return KMX_OK;
	int rc = require_gpu(&m->sm_count);
	if (rc) return rc;
	m->device = g_device;
	if ((rc = ctx_acquire(&m->x))) return rc;
	m->occ2bin16.resize(m->occ2bin.size());
	for (size_t i = 0; i < m->occ2bin16.size(); i++) m->occ2bin16[i] = (uint16_t)m->occ2bin[i];
	DA(&m->d_occ2bin, m->occ2bin16.size() * 2, m->x->stream);
	DA(&m->d_bin2mean, m->bin2mean.size() * 4, m->x->stream);
	// both sources live as long as the model; kernels that read the tables run on the same stream
	CU(cudaMemcpyAsync(m->d_occ2bin, m->occ2bin16.data(), m->occ2bin16.size() * 2, cudaMemcpyHostToDevice, m->x->stream));
	CU(cudaMemcpyAsync(m->d_bin2mean, m->bin2mean.data(), m->bin2mean.size() * 4, cudaMemcpyHostToDevice, m->x->stream));
	return KMX_OK;
}

extern "C" void kmx_destroy(kmx_model* m) {
	if (!m) return;
	if (m->x) {
		cudaSetDevice(m->device);
		cudaStreamSynchronize(m->x->stream);
		cudaStreamSynchronize(m->x->stream2);
		build_state_free(m);
		free_model_device(m);
		dev_free(m->d_occ2bin, m->x->stream);
		dev_free(m->d_bin2mean, m->x->stream);
		ctx_release(m->x);
		m->x = nullptr;
	}
	delete m;
}

extern "C" void kmx_info(const kmx_model* m, kmx_info_t* info) {
	if (m && info) *info = m->info;
}

extern "C" int kmx_model_sync(kmx_model* m) {
	if (!m || !m->x) return fail(KMX_ESTATE, "model has no device state");
	CU(cudaStreamSynchronize(m->x->stream));
	return KMX_OK;
}

// =========================================================================================
// KMC database (listing subset)
// =========================================================================================
static bool read_exact(FILE* f, void* dst, size_t n) { return n == 0 || fread(dst, 1, n, f) == n; }

// kmc_file.cpp:66-99,132-171,177-235: both files carry a 4-byte marker at either end; the header
// sits at the end of .kmc_pre.  Only the header and the LUT are read here; the record area of
// .kmc_suf goes straight to the device in kmx_db_upload.
extern "C" kmx_db* kmx_db_open(const char* db_base) {
	if (!db_base) {
		fail(KMX_EARG, "null database name");
		return nullptr;
	}
	std::string pre_name = std::string(db_base) + ".kmc_pre", suf_name = std::string(db_base) + ".kmc_suf";
	FILE* fp = fopen(pre_name.c_str(), "rb");
	if (!fp) {
		fail(KMX_EIO, "can't open the kmer_data_base %s (%s)", db_base, strerror(errno));
		return nullptr;
	}
	fseeko(fp, 0, SEEK_END);
	const uint64_t pre_size = (uint64_t)ftello(fp);
	std::vector<uint8_t> pre(pre_size);
	rewind(fp);
	bool ok = pre_size >= 32 && read_exact(fp, pre.data(), pre_size);
	fclose(fp);
	if (!ok || memcmp(pre.data(), "KMCP", 4) != 0 || memcmp(pre.data() + pre_size - 4, "KMCP", 4) != 0) {
		fail(KMX_EFORMAT, "%s is not a KMC prefix file", pre_name.c_str());
		return nullptr;
	}
	kmx_db* db = new kmx_db();
	memset(&db->info, 0, sizeof(db->info));
	kmx_db_info_t& h = db->info;
	memcpy(&h.kmc_version, &pre[pre_size - 12], 4);       // kmc_file.cpp:180-184
	if (h.kmc_version != 0x200) {
		fail(KMX_EFORMAT, "%s: KMC database version 0x%x is not supported (only the KMC 2/3 layout 0x200)", pre_name.c_str(), h.kmc_version);
		delete db;
		return nullptr;
	}
	const uint32_t header_offset = pre[pre_size - 8];     // kmc_file.cpp:190-193: one byte
	if ((uint64_t)header_offset + 8 > pre_size || header_offset < 37) {
		fail(KMX_EFORMAT, "%s: bad header offset %u", pre_name.c_str(), header_offset);
		delete db;
		return nullptr;
	}
	const uint8_t* p = &pre[pre_size - 8 - header_offset];   // kmc_file.cpp:197-209
	memcpy(&h.k, p, 4);
	memcpy(&h.mode, p + 4, 4);
	memcpy(&h.counter_size, p + 8, 4);
	memcpy(&h.lut_prefix_length, p + 12, 4);
	memcpy(&h.signature_len, p + 16, 4);
	memcpy(&h.min_count, p + 20, 4);
	memcpy(&h.max_count, p + 24, 4);
	memcpy(&h.total_kmers, p + 28, 8);
	const uint64_t body = pre_size - 12;                  // two markers and the header_offset word removed
	const uint64_t sig_bytes = ((1ULL << (2 * (h.signature_len & 31))) + 1) * 4;
	if (h.signature_len > 11 || sig_bytes + header_offset + 8 > body) {
		fail(KMX_EFORMAT, "%s: signature map does not fit the file", pre_name.c_str());
		delete db;
		return nullptr;
	}
	const uint64_t lut_bytes = body - (sig_bytes + header_offset + 8);   // kmc_file.cpp:212
	h.lut_entries = lut_bytes / 8;
	if (h.k < 3 || h.k > 32 || h.lut_prefix_length > h.k || (h.k - h.lut_prefix_length) % 4 != 0 || h.lut_entries == 0 ||
	    h.counter_size > 4) {
		fail(KMX_EFORMAT, "%s: unsupported geometry k=%u lut_prefix_length=%u counter_size=%u (k <= 32 only: the reference packs k-mers in 64 bits, tools.hpp:63-76)",
		     pre_name.c_str(), h.k, h.lut_prefix_length, h.counter_size);
		delete db;
		return nullptr;
	}
	if (h.mode != 0) {
		fail(KMX_EFORMAT, "%s: mode %u (quality-weighted counters) is not supported", pre_name.c_str(), h.mode);
		delete db;
		return nullptr;
	}
	db->lut.resize(h.lut_entries + 1);
	memcpy(db->lut.data(), &pre[4], (h.lut_entries + 1) * 8);
	db->lut[h.lut_entries] = h.total_kmers + 1;           // kmc_file.cpp:223
	h.record_bytes = (h.k - h.lut_prefix_length) / 4 + h.counter_size;   // kmc_file.cpp:230-232
	h.suffix_bytes = (uint64_t)h.record_bytes * h.total_kmers;

	int fd = open(suf_name.c_str(), O_RDONLY);
	if (fd < 0) {
		fail(KMX_EIO, "can't open the kmer_data_base %s (%s)", db_base, strerror(errno));
		delete db;
		return nullptr;
	}
	struct stat st;
	char m0[4] = { 0 }, m1[4] = { 0 };
	ok = fstat(fd, &st) == 0 && st.st_size >= 8 && pread(fd, m0, 4, 0) == 4 && pread(fd, m1, 4, st.st_size - 4) == 4 &&
	     memcmp(m0, "KMCS", 4) == 0 && memcmp(m1, "KMCS", 4) == 0 && (uint64_t)st.st_size - 8 >= h.suffix_bytes;
	if (!ok) {
		close(fd);
		fail(KMX_EFORMAT, "%s is not a KMC suffix file holding %llu records", suf_name.c_str(), (unsigned long long)h.total_kmers);
		delete db;
		return nullptr;
	}
	db->fd = fd;
	db->suf_alloc = (size_t)((h.suffix_bytes + 15) & ~15ULL) + 16;
	db->device = g_device;
	return db;
}

extern "C" void kmx_db_info(const kmx_db* db, kmx_db_info_t* info) {
	if (db && info) {
		*info = db->info;
		info->on_device = db->d_suf != nullptr;
	}
}

// Process-wide pinned bounce buffers for file -> device streaming (allocated once, kept).
namespace {
constexpr size_t kChunk = 2u << 20;                 // small chunks: the first PCIe transfer starts after 2 MiB of page-cache copy
constexpr int kReaders = 16, kSlotsPerReader = 2;
struct Bounce {
	std::mutex mu;
	uint8_t* buf[kReaders][kSlotsPerReader] = {};
	bool ready = false;
	int acquire() {
		if (ready) return KMX_OK;
		for (auto& r : buf)
			for (auto& b : r)
				if (cudaHostAlloc((void**)&b, kChunk, cudaHostAllocPortable) != cudaSuccess) return fail(KMX_ECUDA, "pinned bounce buffer allocation failed");
		ready = true;
		return KMX_OK;
	}
};
Bounce g_bounce;
}  // namespace

// .kmc_suf record area -> HBM: reader threads pread() 2 MiB chunks into pinned bounce buffers and push each with
// its own stream, so page-cache copies and PCIe transfers overlap.  The page-cache copy (5-6 GB/s per thread) is the
// slow half: one thread per host core, up to kReaders.
extern "C" int kmx_db_upload(kmx_db* db) {
	if (!db) return fail(KMX_EARG, "null database");
	if (db->d_suf) return KMX_OK;
	int rc = require_gpu(&db->sm_count);
	if (rc) return rc;
	db->device = g_device;
	auto t0 = std::chrono::high_resolution_clock::now();
	DevCtx* x = nullptr;
	if ((rc = ctx_acquire(&x))) return rc;
	struct Release {
		DevCtx* x;
		~Release() { ctx_release(x); }
	} release{ x };
	std::lock_guard<std::mutex> lock(g_bounce.mu);
	if ((rc = g_bounce.acquire())) return rc;
	DA(&db->d_suf, db->suf_alloc, x->stream);
	DA(&db->d_lut, db->lut.size() * 8, x->stream);
	CU(cudaMemcpyAsync(db->d_lut, db->lut.data(), db->lut.size() * 8, cudaMemcpyHostToDevice, x->stream));
	const uint64_t bytes = db->info.suffix_bytes;
	CU(cudaMemsetAsync(db->d_suf + (bytes & ~15ULL), 0, db->suf_alloc - (bytes & ~15ULL), x->stream));
	CU(cudaStreamSynchronize(x->stream));                // the allocation is usable from the reader streams now
	TRACE(t0, "upload: buffers ready");
	const uint64_t n_chunks = (bytes + kChunk - 1) / kChunk;
	int want_thr = std::max(1, std::min<int>(kReaders, (int)std::thread::hardware_concurrency()));
	if (const char* e = getenv("KMX_READERS")) want_thr = std::max(1, std::min(kReaders, atoi(e)));
	const int n_thr = (int)std::min<uint64_t>(want_thr, n_chunks);
	std::vector<int> status(kReaders, KMX_OK);
	auto work = [&](int t) {
		cudaSetDevice(db->device);
		cudaStream_t st = x->reader[t];
		int slot = 0;
		for (uint64_t c = t; c < n_chunks && status[t] == KMX_OK; c += n_thr, slot ^= 1) {
			const uint64_t off = c * kChunk, len = std::min<uint64_t>(kChunk, bytes - off);
			cudaEventSynchronize(x->reader_ev[t][slot]);       // the previous copy out of this slot is done
			uint8_t* b = g_bounce.buf[t][slot];
			uint64_t got = 0;
			while (got < len) {
				ssize_t r = pread(db->fd, b + got, len - got, (off_t)(4 + off + got));
				if (r <= 0) { status[t] = KMX_EIO; break; }
				got += (uint64_t)r;
			}
			if (status[t] != KMX_OK) break;
			if (cudaMemcpyAsync(db->d_suf + off, b, len, cudaMemcpyHostToDevice, st) != cudaSuccess) { status[t] = KMX_ECUDA; break; }
			cudaEventRecord(x->reader_ev[t][slot], st);
		}
		if (cudaStreamSynchronize(st) != cudaSuccess && status[t] == KMX_OK) status[t] = KMX_ECUDA;
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < n_thr; t++) pool.emplace_back(work, t);
	if (n_thr > 0) work(0);
	for (auto& th : pool) th.join();
	for (int t = 0; t < n_thr; t++)
		if (status[t] != KMX_OK) return fail(status[t], status[t] == KMX_EIO ? "short read on the .kmc_suf file" : "host-to-device copy of the database failed");
	db->ms_upload = (float)(1e3 * std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count());
	TRACE(t0, "upload: done");
	return KMX_OK;
}

extern "C" void kmx_db_close(kmx_db* db) {
	if (!db) return;
	if (db->d_suf || db->d_lut) {
		cudaSetDevice(db->device);
		cudaDeviceSynchronize();
		dev_free(db->d_suf, nullptr);
		dev_free(db->d_lut, nullptr);
	}
	if (db->fd >= 0) close(db->fd);
	delete db;
}

static DevDb dev_db(const kmx_db* db) {
	DevDb d;
	memset(&d, 0, sizeof(d));
	d.suf = db->d_suf;
	d.lut = db->d_lut;
	d.lut_entries = db->info.lut_entries;
	d.total = db->info.total_kmers;
	d.prefix_mask = (1ULL << (2 * db->info.lut_prefix_length)) - 1;
	d.suffix_bytes = (db->info.k - db->info.lut_prefix_length) / 4;
	d.counter_bytes = db->info.counter_size;
	d.rec_bytes = db->info.record_bytes;
	d.min_count = db->info.min_count;
	d.max_count = db->info.max_count;
	d.k = (int)db->info.k;
	return d;
}

extern "C" int kmx_db_list(kmx_db* db, uint64_t* kmers, uint32_t* counts, uint64_t* n_out) {
	if (!db || !kmers || !counts || !n_out) return fail(KMX_EARG, "null argument");
	int rc = kmx_db_upload(db);
	if (rc) return rc;
	CU(cudaSetDevice(db->device));
	const uint64_t total = db->info.total_kmers;
	*n_out = 0;
	if (total == 0) return KMX_OK;
	const uint64_t n_tiles = (total + kTile - 1) / kTile;
	uint32_t* d_cnt = nullptr;
	uint64_t* d_off = nullptr;
	uint64_t* d_k = nullptr;
	uint32_t* d_c = nullptr;
	DevCtx* x = nullptr;
	if ((rc = ctx_acquire(&x))) return rc;
	struct Release {
		DevCtx* x;
		~Release() { ctx_release(x); }
	} release{ x };
	cudaStream_t s = x->stream;
	DA(&d_cnt, n_tiles * 4, s);
	DA(&d_off, (n_tiles + 1) * 8, s);
	DA(&d_k, total * 8, s);
	DA(&d_c, total * 4, s);
	DevDb d = dev_db(db);
	CU(launch_list_count(d, d_cnt, db->sm_count, s));
	CU(launch_tile_scan(d_cnt, n_tiles, d_off, s));
	CU(launch_list(d, d_off, d_k, d_c, db->sm_count, s));
	uint64_t listed = 0;
	CU(cudaMemcpyAsync(&listed, d_off + n_tiles, 8, cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));
	CU(cudaMemcpyAsync(kmers, d_k, listed * 8, cudaMemcpyDeviceToHost, s));
	CU(cudaMemcpyAsync(counts, d_c, listed * 4, cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));
	dev_free(d_cnt, s); dev_free(d_off, s); dev_free(d_k, s); dev_free(d_c, s);
	CU(cudaStreamSynchronize(s));
	*n_out = listed;
	return KMX_OK;
}

// =========================================================================================
// build: KModel::init (kmodel.hpp:57-86), in stages so that the multi-GPU path can interleave
// its exchanges:  encode (count + Bloom + item stream)  ->  insert  ->  rest table
// =========================================================================================
// bucket index over the top key bits (about one entry per bucket) + the per-prefix false-hit table
// that reproduces the inclusive upper bound of KRestData::check_kmer (rest.hpp:233-237)
static int build_rest_side_tables(kmx_model* m) {
	const RestHost& r = m->rest;
	cudaStream_t s = m->x->stream;
	int bits = 8;
	while (bits < 26 && bits < 2 * r.k && (1ULL << bits) < r.count) bits++;
	m->fine_bits = bits;
	DA(&m->d_fine, ((size_t)(1u << bits) + 1) * 4, s);
	DA(&m->d_quirk_suffix, (size_t)r.map_size * 8, s);
	DA(&m->d_quirk_index, (size_t)r.map_size * 4, s);
	fill_dev_model(m);
	CU(launch_rest_side_tables(m->dm.rest, r.map_size, m->d_fine, m->d_quirk_suffix, m->d_quirk_index, s));
	return KMX_OK;
}

static int build_rest_table(kmx_model* m, const uint64_t* d_surv_kmer, const uint32_t* d_surv_occ, uint64_t n, int32_t* h_groups) {
	RestHost& r = m->rest;
	cudaStream_t s = m->x->stream;
	r.k = m->k;
	r.pre_len = rest_prefix_len(m->k);                    // rest.hpp:140-149
	r.map_size = 1 << (2 * r.pre_len);
	r.count = n;
	r.suff_bin_size = n * (uint64_t)((m->k - r.pre_len) / 4);
	if (n > 0x7FFFFFFFULL) return fail(KMX_ERANGE, "%llu rest entries overflow the reference's int indices (rest.hpp:66-70)", (unsigned long long)n);
	DA(&m->d_hash2index, (size_t)r.map_size * 4, s);
	DA(&m->d_pre_buffer, ((size_t)r.map_size + 1) * 4, s);
	DA(&m->d_rest_keys, (n + 1) * 8, s);
	DA(&m->d_rest_counts, (n + 1) * 4, s);
	int32_t* d_first = nullptr;
	int32_t* d_groups = nullptr;
	void* d_temp = nullptr;
	size_t temp_bytes = 0;
	DA(&d_first, (size_t)r.map_size * 4, s);
	DA(&d_groups, 4, s);
	CU(rest_sort_bytes(n, &temp_bytes));
	if (temp_bytes) DA(&d_temp, temp_bytes, s);
	CU(launch_rest_sort(d_temp, temp_bytes, d_surv_kmer, m->d_rest_keys, d_surv_occ, m->d_rest_counts, n, 2 * m->k, s));
	CU(launch_rest_index(m->d_rest_keys, n, 2 * (m->k - r.pre_len), r.map_size, d_first, m->d_hash2index, m->d_pre_buffer, d_groups, s));
	CU(cudaMemcpyAsync(h_groups, d_groups, 4, cudaMemcpyDeviceToHost, s));
	dev_free(d_first, s); dev_free(d_groups, s); dev_free(d_temp, s);
	return KMX_OK;
}

// Exchange slabs and peer mappings are kept for the life of the process: cudaMalloc / cudaFree and
// cudaIpcOpenMemHandle cost milliseconds each, a rebuild with the same geometry reuses them.
namespace {
struct CachedSlab {
	void* ptr;
	size_t bytes;
	int device;
};
std::mutex g_slab_mu;
std::vector<CachedSlab> g_slab_free;
std::map<std::string, void*> g_ipc_open;             // 64 handle bytes -> mapping in this process
}  // namespace

static int slab_acquire(void** out, size_t bytes, int device) {
	{
		std::lock_guard<std::mutex> lock(g_slab_mu);
		for (size_t q = 0; q < g_slab_free.size(); q++) {
			if (g_slab_free[q].device == device && g_slab_free[q].bytes == bytes) {
				*out = g_slab_free[q].ptr;
				g_slab_free.erase(g_slab_free.begin() + q);
				return KMX_OK;
			}
		}
	}
	CU(cudaMalloc(out, bytes));
	return KMX_OK;
}

static void slab_release(void* ptr, size_t bytes, int device) {
	std::lock_guard<std::mutex> lock(g_slab_mu);
	g_slab_free.push_back(CachedSlab{ ptr, bytes, device });
}

// the filters move from the exchange slab into allocations of their own (the slab goes back to the cache)
static int filters_leave_slab(kmx_model* m) {
	BuildState& b = m->bs;
	if (!b.fslab) return KMX_OK;
	cudaStream_t s = m->x->stream;
	auto move = [&](uint32_t** p, uint64_t bytes) -> int {
		uint32_t* fresh = nullptr;
		DA(&fresh, pad8(bytes), s);
		CU(cudaMemcpyAsync(fresh, *p, pad8(bytes), cudaMemcpyDeviceToDevice, s));
		*p = fresh;
		return KMX_OK;
	};
	for (int i = 0; i < m->bf_num; i++) {
		int rc = move(&m->d_bf[i], m->bytes[i]);
		if (!rc) rc = move(&m->d_bf_back[i], m->bytes[3 + i]);
		if (rc) return rc;
	}
	int rc = move(&m->d_km_back, m->bytes[7]);
	if (rc) return rc;
	CU(cudaStreamSynchronize(s));
	slab_release(b.fslab, b.fslab_bytes, m->device);
	b.fslab = nullptr;
	fill_dev_model(m);
	return KMX_OK;
}

static void build_state_free(kmx_model* m) {
	BuildState& b = m->bs;
	if (!m->x) return;
	cudaStream_t s = m->x->stream;
	InsertArgs& a = b.a;
	if (b.fslab) {                                        // a build that failed half-way: the filters still point into the slab
		cudaStreamSynchronize(s);
		for (int i = 0; i < 3; i++) m->d_bf[i] = m->d_bf_back[i] = nullptr;
		m->d_km_back = nullptr;
		slab_release(b.fslab, b.fslab_bytes, m->device);
		b.fslab = nullptr;
	}
	dev_free(b.d_item_kmer, s);
	dev_free(b.d_item_occ, s);
	if (!b.slab) {
		for (int q = 0; q < 2; q++) {
			dev_free(a.buf_kmer[q], s);
			dev_free(a.buf_occ[q], s);
		}
		dev_free(a.ctl, s);
	}
	dev_free(a.status, s); dev_free(a.excl_rank, s); dev_free(a.holepos, s); dev_free(a.list[0], s); dev_free(a.list[1], s); dev_free(a.list[2], s);
	dev_free(a.tile_fail, s); dev_free(a.resv, s); dev_free(a.claim, s);
	dev_free(a.rest_kmer, s); dev_free(a.rest_occ, s);
	if (b.slab) {
		cudaStreamSynchronize(s);
		slab_release(b.slab, b.slab_bytes, m->device);
	}
	b = BuildState();
}

// stage 1: pass 1 (class histogram, kmodel.hpp:423-434), filter allocation (kmodel.hpp:402-456),
// pass 2 (Bloom inserts + array-bound stream, kmodel.hpp:68-74)
// dist_world > 1: this rank inserts the Bloom-bound records of its share of the tiles only (the partial filters are OR-ed
// afterwards, kmx_dist_merge) and writes the item stream only if stream_items (ranks that own a coupled array)
static int build_stage_encode(kmx_model* m, kmx_db* db, int dist_rank = 0, int dist_world = 1, bool stream_items = true) {
	BuildState& b = m->bs;
	b.wall0 = std::chrono::high_resolution_clock::now();
	b.world = dist_world;
	b.dist_rank = dist_rank;
	int rc = model_attach_device(m);
	if (rc) return rc;
	TRACE(b.wall0, "device attached");
	CU(cudaSetDevice(m->device));
	rc = kmx_db_upload(db);
	if (rc) return rc;
	TRACE(b.wall0, "database on device");
	if (db->device != m->device) return fail(KMX_EARG, "database is on device %d, model on device %d", db->device, m->device);
	m->k = (int)db->info.k;
	m->total_kmers = db->info.total_kmers;
	if (m->k < 3) return fail(KMX_ERANGE, "k=%d: the (k-2)-mer filters need k >= 3", m->k);
	const uint64_t total = m->total_kmers;
	const uint64_t n_tiles = (total + kTile - 1) / kTile;
	DevDb d = dev_db(db);
	cudaStream_t s = m->x->stream;
	cudaEvent_t* ev = m->x->ev_build;
	b.ms_upload = db->ms_upload;

	uint32_t* d_tile_cnt = nullptr;
	uint64_t* d_tile_off = nullptr;
	CountOut* d_count = nullptr;
	DA(&d_tile_cnt, (n_tiles + 1) * 4, s);
	DA(&d_tile_off, (n_tiles + 1) * 8, s);
	DA(&d_count, sizeof(CountOut), s);
	CU(cudaEventRecord(ev[0], s));
	CU(cudaMemsetAsync(d_count, 0, sizeof(CountOut), s));
	CU(launch_count(d, m->ci, m->cs, m->bf_num, d_count, d_tile_cnt, m->sm_count, s));
	CU(launch_tile_scan(d_tile_cnt, n_tiles, d_tile_off, s));
	CountOut& cnt = m->x->h_pinned->count;
	CU(cudaMemcpyAsync(&cnt, d_count, sizeof(cnt), cudaMemcpyDeviceToHost, s));
	CU(cudaEventRecord(ev[1], s));
	TRACE(b.wall0, "count pass queued");
	CU(cudaStreamSynchronize(s));                        // sync 1 of 3: the sizes depend on the counts
	TRACE(b.wall0, "count pass done");
	dev_free(d_tile_cnt, s);
	dev_free(d_count, s);
	if (cnt.bad_count) {
		dev_free(d_tile_off, s);
		return fail(KMX_ERANGE, "%llu records have a count below ci=%d or above cs=%d: the reference indexes out of bounds there (kmodel.hpp:427, occu_bin.hpp:70)",
		            (unsigned long long)cnt.bad_count, m->ci, m->cs);
	}
	uint64_t bf_kmers = 0;
	for (int i = 0; i < m->bf_num; i++) {
		m->kmer_counts[i] = cnt.class_count[i];
		bf_kmers += cnt.class_count[i];
	}
	m->km_kmers = total - bf_kmers;                       // kmodel.hpp:433: header total, not the listed count
	rc = alloc_filters(m, dist_world > 1);
	if (rc) {
		dev_free(d_tile_off, s);
		return rc;
	}
	m->rest.k = m->k;
	m->rest.pre_len = rest_prefix_len(m->k);
	fill_dev_model(m);
	if (dist_world > 1) CU(cudaStreamSynchronize(s));     // the filter slab is zeroed before its handle leaves this process

	b.n_items = cnt.array_bound;
	if (stream_items) {
		DA(&b.d_item_kmer, (b.n_items + 1) * 8, s);
		DA(&b.d_item_occ, (b.n_items + 1) * 4, s);
	}
	const uint64_t tile_lo = n_tiles * (uint64_t)dist_rank / (uint64_t)dist_world, tile_hi = n_tiles * (uint64_t)(dist_rank + 1) / (uint64_t)dist_world;
	CU(launch_encode(d, m->dm, d_tile_off, b.d_item_kmer, b.d_item_occ, tile_lo, tile_hi, stream_items, m->sm_count, s));
	dev_free(d_tile_off, s);
	CU(cudaEventRecord(ev[2], s));
	return KMX_OK;
}

// stage 2a: buffers of the greedy insert (kmodel.hpp:508-573).  shared = true puts the buffers
// other GPUs write into (survivor ping-pong, control block, barrier flags) in one cudaMalloc
// slab that can be exported with cudaIpcGetMemHandle.
static int build_stage_insert_setup(kmx_model* m, int rank, int n_active, bool shared) {
	BuildState& b = m->bs;
	cudaStream_t s = m->x->stream;
	InsertArgs& a = b.a;
	memset(&a, 0, sizeof(a));
	const uint64_t batch_items = (uint64_t)m->n_bits << kBucketLog;
	b.n_batches = (b.n_items + batch_items - 1) / batch_items;
	a.rank = rank;
	a.n_active = n_active;
	a.item_kmer = b.d_item_kmer;
	a.item_occ = b.d_item_occ;
	a.n_items = b.n_items;
	if (shared) {
		auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
		const size_t o_k0 = 0, o_k1 = o_k0 + up(batch_items * 8), o_o0 = o_k1 + up(batch_items * 8), o_o1 = o_o0 + up(batch_items * 4),
		             o_ctl = o_o1 + up(batch_items * 4), o_flags = o_ctl + up(sizeof(InsertCtl));
		b.slab_bytes = o_flags + up(kMaxRanks * 4);
		{
			int rc = slab_acquire(&b.slab, b.slab_bytes, m->device);
			if (rc) return rc;
		}
		CU(cudaMemsetAsync((uint8_t*)b.slab + o_ctl, 0, b.slab_bytes - o_ctl, s));   // control block + barrier flags start at zero
		b.off[0] = o_k0; b.off[1] = o_k1; b.off[2] = o_o0; b.off[3] = o_o1; b.off[4] = o_ctl; b.off[5] = o_flags;
		uint8_t* base = (uint8_t*)b.slab;
		a.buf_kmer[0] = (uint64_t*)(base + o_k0);
		a.buf_kmer[1] = (uint64_t*)(base + o_k1);
		a.buf_occ[0] = (uint32_t*)(base + o_o0);
		a.buf_occ[1] = (uint32_t*)(base + o_o1);
		a.ctl = (InsertCtl*)(base + o_ctl);
		b.flags = (uint32_t*)(base + o_flags);
		b.peer_slab[rank] = b.slab;
	} else {
		for (int q = 0; q < 2; q++) {
			DA(&a.buf_kmer[q], batch_items * 8, s);
			DA(&a.buf_occ[q], batch_items * 4, s);
		}
		DA(&a.ctl, sizeof(InsertCtl), s);
		CU(cudaMemsetAsync(a.ctl, 0, sizeof(InsertCtl), s));
	}
	if (const char* e = getenv("KMX_TEST_EPOCH_START")) {    // lets a small test cross the epoch wrap-around
		const unsigned int start = (unsigned int)atoi(e);
		CU(cudaMemcpyAsync(&a.ctl->epoch, &start, sizeof(start), cudaMemcpyHostToDevice, s));
		CU(cudaStreamSynchronize(s));
	}
	for (int q = 0; q < 2; q++) {
		a.peer_buf_kmer[q][rank] = a.buf_kmer[q];
		a.peer_buf_occ[q][rank] = a.buf_occ[q];
	}
	a.peer_ctl[rank] = a.ctl;
	a.peer_flags[rank] = b.flags;
	DA(&a.status, batch_items * 4, s);
	DA(&a.excl_rank, batch_items * 4, s);
	DA(&a.holepos, batch_items * 4, s);
	DA(&a.list[0], batch_items * 4, s);
	DA(&a.list[1], batch_items * 4, s);
	DA(&a.list[2], batch_items * 4, s);
	DA(&a.tile_fail, batch_items / 256 * 4, s);      // one counter per reorder tile (a tile is >= 256 ids)
	CU(cudaMemsetAsync(a.tile_fail, 0, batch_items / 256 * 4, s));
	a.resv_slots = 1u << 20;
	if (const char* e = getenv("KMX_RESV_LOG2")) {
		int v = atoi(e);
		if (v >= 10 && v <= 26) a.resv_slots = 1u << v;
	}
	DA(&a.resv, (size_t)m->n_bits * 4 * a.resv_slots * 4, s);
	CU(cudaMemsetAsync(a.resv, 0xFF, (size_t)m->n_bits * 4 * a.resv_slots * 4, s));
	// contested items: merged reserve/commit passes (hc14 shape: insert 163 -> 155 ms, RS shape: 3.12 -> 3.08 ms); the
	// multi-GPU build keeps the classic two-barrier iterations it was validated with on 8 GPUs
	a.merged = n_active == 1 ? 1 : 0;
	if (const char* e = getenv("KMX_MERGED_PASSES")) a.merged = atoi(e) ? 1 : 0;
	a.claim_log2 = 25;
	if (const char* e = getenv("KMX_CLAIM_LOG2")) {
		int v = atoi(e);
		if (v >= 15 && v <= 30) a.claim_log2 = (uint32_t)v;
	}
	DA(&a.claim, (size_t)m->n_bits * 2 * ((size_t)1 << (a.claim_log2 - 5)) * 4, s);
	CU(cudaMemsetAsync(a.claim, 0, (size_t)m->n_bits * 2 * ((size_t)1 << (a.claim_log2 - 5)) * 4, s));
	// claim_first merges the cell read with the commit (see insert_kernel phase 0).  Measured on B200 it is a wash
	// for HBM-resident models (hc14 shape: 187 ms against 186 ms) and slower for L2-resident ones, so it stays off.
	a.claim_first = 0;
	if (const char* e = getenv("KMX_CLAIM_FIRST")) a.claim_first = atoi(e) ? 1 : 0;
	// arrays + km_back well beyond the L2 (126 MB): their random sectors should not wash the insert's hot structures out of it
	// (measured on the hc14 shape: evict-first cell loads + reductions 182 -> 168 ms; evict-first on km_back as well: 171 ms)
	a.stream_cells = (2ULL * m->n_bits * m->bytes[6] + m->bytes[7]) > (192ULL << 20) ? 3 : 0;
	if (const char* e = getenv("KMX_STREAM_CELLS")) a.stream_cells = atoi(e) & 15;
	a.max_iterations = kBucket + 64;
	a.phase_round = -1;
	if (const char* e = getenv("KMX_PHASE_ROUND")) a.phase_round = atoi(e);
	CU(insert_grid_size(&b.grid, m->sm_count));
	// Survivor list: sized for the worst case (nothing accepted) while that is cheap, which lets
	// all launches queue without a host round trip; beyond that it grows between launches.
	b.worst_case = b.n_items <= (1ULL << 28) && !getenv("KMX_TEST_GROW_REST");   // the env var lets a small test take the growing path
	b.rest_cap = b.worst_case ? b.n_items + m->n_bits : 0;
	if (b.worst_case) {
		DA(&a.rest_kmer, b.rest_cap * 8, s);
		DA(&a.rest_occ, b.rest_cap * 4, s);
	}
	return KMX_OK;
}

// stage 2b: the launches (64 batches each); returns with the control block on the host
static int build_stage_insert_run(kmx_model* m) {
	BuildState& b = m->bs;
	cudaStream_t s = m->x->stream;
	InsertArgs& a = b.a;
	InsertCtl& ctl = m->x->h_pinned->ctl;
	memset(&ctl, 0, sizeof(ctl));
	const uint64_t batch_items = (uint64_t)m->n_bits << kBucketLog;
	const bool participates = a.rank < a.n_active;
	CU(cudaEventRecord(m->x->ev_build[5], s));           // the multi-GPU path spends host time between encode and here
	if (b.n_items > 0 && participates) {
		const uint64_t chunk = 64;                         // batches per launch
		for (uint64_t b0 = 0; b0 < b.n_batches; b0 += chunk) {
			const uint64_t nb = std::min<uint64_t>(chunk, b.n_batches - b0);
			const uint64_t need = ctl.rest_n + nb * batch_items + m->n_bits;
			if (!b.worst_case && need > b.rest_cap) {
				uint64_t new_cap = std::max<uint64_t>(need, b.rest_cap * 2);
				uint64_t* nk = nullptr;
				uint32_t* no = nullptr;
				DA(&nk, new_cap * 8, s);
				DA(&no, new_cap * 4, s);
				if (ctl.rest_n) {
					CU(cudaMemcpyAsync(nk, a.rest_kmer, ctl.rest_n * 8, cudaMemcpyDeviceToDevice, s));
					CU(cudaMemcpyAsync(no, a.rest_occ, ctl.rest_n * 4, cudaMemcpyDeviceToDevice, s));
				}
				dev_free(a.rest_kmer, s);
				dev_free(a.rest_occ, s);
				a.rest_kmer = nk;
				a.rest_occ = no;
				b.rest_cap = new_cap;
			}
			a.rest_cap = b.rest_cap;
			a.first_batch = b0;
			a.n_batches = nb;
			CU(launch_insert(m->dm, a, b.grid, s));
			if (!b.worst_case) {
				CU(cudaMemcpyAsync(&ctl, a.ctl, sizeof(ctl), cudaMemcpyDeviceToHost, s));
				CU(cudaStreamSynchronize(s));
				if (ctl.error) return fail(KMX_ECUDA, "insert kernel stopped with error %u (1: iteration cap, 2: survivor list overflow, 3: peer GPU timed out)", ctl.error);
			}
		}
		CU(cudaMemcpyAsync(&ctl, a.ctl, sizeof(ctl), cudaMemcpyDeviceToHost, s));
	}
	CU(cudaEventRecord(m->x->ev_build[3], s));
	TRACE(b.wall0, "encode + insert queued");
	CU(cudaStreamSynchronize(s));                        // sync 2 of 3: the sort needs the survivor count
	TRACE(b.wall0, "insert done");
	if (ctl.error) return fail(KMX_ECUDA, "insert kernel stopped with error %u (1: iteration cap, 2: survivor list overflow, 3: peer GPU timed out)", ctl.error);
	return KMX_OK;
}

// stage 3: rest table (rest.hpp:157-161) from the survivors at d_rest_* (n of them), statistics, clean-up
static int build_stage_finish(kmx_model* m, const uint64_t* d_rest_kmer, const uint32_t* d_rest_occ, uint64_t rest_n) {
	BuildState& b = m->bs;
	cudaStream_t s = m->x->stream;
	cudaEvent_t* ev = m->x->ev_build;
	const InsertCtl ctl = m->x->h_pinned->ctl;
	int32_t& groups = m->x->h_pinned->groups;
	int rc = build_rest_table(m, d_rest_kmer, d_rest_occ, rest_n, &groups);
	if (rc) return rc;
	rc = build_rest_side_tables(m);
	if (rc) return rc;
	CU(cudaEventRecord(ev[4], s));
	TRACE(b.wall0, "rest table queued");
	CU(cudaStreamSynchronize(s));                        // sync 3 of 3
	TRACE(b.wall0, "rest table done");
	m->rest.pre_buffer_size = groups + 1;                 // rest.hpp:119: new int[++pre_buffer_size]
	fill_dev_model(m);
	m->built = true;
	kmx_info_t& f = m->info;
	fill_info(m);
	f.insert_attempts = ctl.attempts;
	f.insert_accepted = ctl.accepted;
	f.insert_iterations = ctl.iterations;
	f.batches = b.n_batches;
	for (int i = 0; i < 8; i++) f.insert_phase_cycles[i] = ctl.phase_cycles[i];
	f.ms_upload = b.ms_upload;
	CU(cudaEventElapsedTime(&f.ms_count, ev[0], ev[1]));
	CU(cudaEventElapsedTime(&f.ms_encode, ev[1], ev[2]));
	CU(cudaEventElapsedTime(&f.ms_insert, ev[5], ev[3]));
	CU(cudaEventElapsedTime(&f.ms_rest, ev[3], ev[4]));
	CU(cudaEventElapsedTime(&f.ms_total_device, ev[0], ev[4]));
	f.build_time_cost = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - b.wall0).count();
	build_state_free(m);
	return KMX_OK;
}

extern "C" int kmx_init_from_db(kmx_model* m, kmx_db* db) {
	if (!m || !db) return fail(KMX_EARG, "null argument");
	if (m->built) return fail(KMX_ESTATE, "model already initialised (KModel::init is one-shot)");
	int rc = build_stage_encode(m, db);
	if (!rc) rc = build_stage_insert_setup(m, 0, 1, false);
	if (!rc) rc = build_stage_insert_run(m);
	if (!rc) rc = build_stage_finish(m, m->bs.a.rest_kmer, m->bs.a.rest_occ, m->x->h_pinned->ctl.rest_n);
	if (rc) build_state_free(m);
	return rc;
}

extern "C" int kmx_init_from_kmc(kmx_model* m, const char* db_base) {
	if (!m) return fail(KMX_EARG, "null model");
	int rc = require_gpu(nullptr);
	if (rc) return rc;
	kmx_db* db = kmx_db_open(db_base);
	if (!db) return KMX_EIO;
	rc = kmx_init_from_db(m, db);
	kmx_db_close(db);
	return rc;
}

// ---- multi-GPU build, array-owner decomposition (option A) ----------------------
// Every rank decodes the database and fills the Bloom filters (replicated: order-free and cheap);
// the coupled arrays are split by ownership -- array a lives on rank a % n_active -- and the
// persistent insert kernels of the ranks hand each bucket's survivors to the next owner through
// peer-mapped buffers, with a flag barrier per round.  The caller (kmcex_b200/distributed.py)
// moves IPC handles and, afterwards, the finished pieces with torch.distributed.
extern "C" int kmx_dist_prepare(kmx_model* m, kmx_db* db, int rank, int n_active, int world, void* ipc_handles_out) {
	if (!m || !db || !ipc_handles_out) return fail(KMX_EARG, "null argument");
	if (m->built) return fail(KMX_ESTATE, "model already initialised (KModel::init is one-shot)");
	if (n_active < 1 || n_active > kMaxRanks || n_active > m->n_bits || rank < 0) return fail(KMX_EARG, "n_active must be in 1..min(%d, n_bits)", kMaxRanks);
	if (world < n_active || world > kMaxRanks || rank >= world) return fail(KMX_EARG, "world must be in n_active..%d and rank below it", kMaxRanks);
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles travel as 64 bytes");
	int rc = build_stage_encode(m, db, rank, world, rank < n_active);
	if (!rc) rc = build_stage_insert_setup(m, rank, n_active, true);
	if (!rc) {
		CU(cudaStreamSynchronize(m->x->stream));         // the slabs are zeroed before anybody maps them
		CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)ipc_handles_out, m->bs.slab));
		memset((uint8_t*)ipc_handles_out + 64, 0, 64);
		if (m->bs.fslab) CU(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)((uint8_t*)ipc_handles_out + 64), m->bs.fslab));
	}
	if (rc) build_state_free(m);
	return rc;
}

static int ipc_map(const void* handle64, void** out) {
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, 64);
	std::string key((const char*)&h, 64);
	std::lock_guard<std::mutex> lock(g_slab_mu);
	auto it = g_ipc_open.find(key);
	if (it == g_ipc_open.end()) {
		void* mapped = nullptr;
		CU(cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
		it = g_ipc_open.emplace(key, mapped).first;
	}
	*out = it->second;
	return KMX_OK;
}

extern "C" int kmx_dist_connect(kmx_model* m, const void* handles) {
	if (!m || !handles || !m->bs.slab) return fail(KMX_ESTATE, "kmx_dist_prepare first");
	BuildState& b = m->bs;
	InsertArgs& a = b.a;
	const uint8_t* hs = (const uint8_t*)handles;
	if (a.rank < a.n_active) {                             // an idle rank takes no part in the survivor exchange
		for (int p = 0; p < a.n_active; p++) {
			if (p != a.rank) {
				int rc = ipc_map(hs + 128 * p, &b.peer_slab[p]);
				if (rc) return rc;
			}
			uint8_t* base = (uint8_t*)b.peer_slab[p];
			a.peer_buf_kmer[0][p] = (uint64_t*)(base + b.off[0]);
			a.peer_buf_kmer[1][p] = (uint64_t*)(base + b.off[1]);
			a.peer_buf_occ[0][p] = (uint32_t*)(base + b.off[2]);
			a.peer_buf_occ[1][p] = (uint32_t*)(base + b.off[3]);
			a.peer_ctl[p] = (InsertCtl*)(base + b.off[4]);
			a.peer_flags[p] = (uint32_t*)(base + b.off[5]);
		}
	}
	if (b.fslab) {                                         // every rank takes part in the filter merges
		for (int p = 0; p < b.world; p++) {
			if (p == b.dist_rank) b.peer_fslab[p] = b.fslab;
			else {
				int rc = ipc_map(hs + 128 * p + 64, &b.peer_fslab[p]);
				if (rc) return rc;
			}
		}
	}
	return KMX_OK;
}

// OR all-reduce over the ranks of the Bloom filters (which = 0, after kmx_dist_prepare) or of km_back (which = 1, after
// kmx_dist_insert): one kernel per rank, exchanges through peer memory, asynchronous on the model's stream
extern "C" int kmx_dist_merge(kmx_model* m, int which) {
	if (!m || !m->bs.fslab || !m->bs.peer_fslab[0]) return fail(KMX_ESTATE, "kmx_dist_prepare / kmx_dist_connect first");
	BuildState& b = m->bs;
	OrReduceArgs r;
	memset(&r, 0, sizeof(r));
	r.rank = b.dist_rank;
	r.world = b.world;
	const size_t off = which == 0 ? 0 : b.f_kmback_off, bytes = which == 0 ? b.f_bloom_bytes : b.f_kmback_bytes;
	r.n_vec = bytes / 16;
	for (int p = 0; p < b.world; p++) {
		r.base[p] = (uint4*)((uint8_t*)b.peer_fslab[p] + off);
		r.flags[p] = (uint32_t*)((uint8_t*)b.peer_fslab[p] + b.f_flags_off);
	}
	r.seq = b.fseq;
	r.error = (unsigned int*)((uint8_t*)b.fslab + b.f_flags_off + 128);
	b.fseq += 2;
	CU(cudaSetDevice(m->device));
	CU(launch_or_allreduce(r, m->sm_count, m->x->stream));
	return KMX_OK;
}

extern "C" int kmx_dist_insert(kmx_model* m) {
	if (!m || !m->bs.slab) return fail(KMX_ESTATE, "kmx_dist_prepare first");
	return build_stage_insert_run(m);
}

extern "C" int kmx_dist_buffers(kmx_model* m, kmx_dist_buffers_t* out) {
	if (!m || !out || !m->bs.slab) return fail(KMX_ESTATE, "kmx_dist_prepare first");
	memset(out, 0, sizeof(*out));
	out->n_bits = m->n_bits;
	out->cell_bytes = (cell_words(m->bytes[6]) + 1) * 8;
	for (int i = 0; i < m->n_bits; i++) out->cells[i] = m->d_cells[i];
	out->km_back = m->d_km_back;
	out->km_back_bytes = pad8(m->bytes[7]);
	out->rest_kmer = m->bs.a.rest_kmer;
	out->rest_occ = m->bs.a.rest_occ;
	out->rest_n = m->x->h_pinned->ctl.rest_n;
	out->insert_attempts = m->x->h_pinned->ctl.attempts;
	out->insert_accepted = m->x->h_pinned->ctl.accepted;
	return KMX_OK;
}

extern "C" int kmx_dist_finish(kmx_model* m, const uint64_t* d_rest_kmer, const uint32_t* d_rest_occ, uint64_t rest_n,
                               uint64_t attempts, uint64_t accepted) {
	if (!m || !m->bs.slab) return fail(KMX_ESTATE, "kmx_dist_prepare first");
	if (rest_n && (!d_rest_kmer || !d_rest_occ)) return fail(KMX_EARG, "null survivor list");
	m->x->h_pinned->ctl.attempts = attempts;
	m->x->h_pinned->ctl.accepted = accepted;
	CU(cudaDeviceSynchronize());                         // the caller's collectives ran on its own streams
	if (m->bs.fslab) {
		unsigned int err = 0;
		CU(cudaMemcpy(&err, (uint8_t*)m->bs.fslab + m->bs.f_flags_off + 128, 4, cudaMemcpyDeviceToHost));
		if (err) {
			build_state_free(m);
			return fail(KMX_ECUDA, "filter merge stopped: a peer GPU did not reach the barrier within 20 s");
		}
	}
	int rc = filters_leave_slab(m);
	if (!rc) rc = build_stage_finish(m, d_rest_kmer, d_rest_occ, rest_n);
	if (rc) build_state_free(m);
	return rc;
}

// =========================================================================================
// save / load (kmodel.hpp:173-235, rest.hpp:163-221)
// =========================================================================================
static int write_device(FILE* f, const void* d_ptr, uint64_t bytes, std::vector<uint8_t>& tmp) {
	if (bytes == 0) return KMX_OK;
	const size_t padded = (size_t)((bytes + 7) & ~7ULL);
	if (tmp.size() < padded) tmp.resize(padded);
	CU(cudaMemcpy(tmp.data(), d_ptr, padded, cudaMemcpyDeviceToHost));
	if (fwrite(tmp.data(), 1, bytes, f) != bytes) return fail(KMX_EIO, "short write (%s)", strerror(errno));
	return KMX_OK;
}

extern "C" int kmx_save(kmx_model* m, const char* dir) {
	if (!m || !dir) return fail(KMX_EARG, "null argument");
	if (!m->built) return fail(KMX_ESTATE, "model is not initialised");
	CU(cudaSetDevice(m->device));
	CU(cudaStreamSynchronize(m->x->stream));
	std::string base(dir);
	FILE* fh = fopen((base + "/header").c_str(), "w");
	if (!fh) return fail(KMX_EIO, "cannot write %s/header (%s); the directory must exist (README.md:77)", dir, strerror(errno));
	fprintf(fh, "number_hash %d\nnumber_bit %d\nci %d\ncs %d\n", m->n_hash, m->n_bits, m->ci, m->cs);   // kmodel.hpp:175-180
	fclose(fh);
	FILE* f = fopen((base + "/km.bin").c_str(), "wb");
	if (!f) return fail(KMX_EIO, "cannot write %s/km.bin (%s)", dir, strerror(errno));
	std::vector<uint8_t> tmp;
	int rc = KMX_OK;
	fwrite(&m->km_kmers, 8, 1, f);
	for (int i = 0; i < m->bf_num; i++) fwrite(&m->kmer_counts[i], 8, 1, f);
	for (int i = 0; i < m->bf_num && !rc; i++) {
		rc = write_device(f, m->d_bf[i], m->bytes[i], tmp);
		if (!rc) rc = write_device(f, m->d_bf_back[i], m->bytes[3 + i], tmp);
	}
	if (!rc) rc = write_device(f, m->d_km_back, m->bytes[7], tmp);
	const uint64_t words = cell_words(m->bytes[6]);
	uint32_t* d_val = nullptr;
	uint32_t* d_tag = nullptr;
	if (!rc) {
		if (cudaMalloc(&d_val, (words + 2) * 4) != cudaSuccess || cudaMalloc(&d_tag, (words + 2) * 4) != cudaSuccess)
			rc = fail(KMX_ECUDA, "out of device memory while saving");
	}
	for (int i = 0; i < m->n_bits && !rc; i++) {
		cudaError_t e = launch_split_cells(m->d_cells[i], words, d_val, d_tag, m->x->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(m->x->stream);
		if (e != cudaSuccess) rc = fail(KMX_ECUDA, "split kernel: %s", cudaGetErrorString(e));
		if (!rc) rc = write_device(f, d_val, m->bytes[6], tmp);      // bit_array_1
		if (!rc) rc = write_device(f, d_tag, m->bytes[6], tmp);      // bit_array_2
	}
	cudaFree(d_val);
	cudaFree(d_tag);
	if (fclose(f) != 0 && !rc) rc = fail(KMX_EIO, "close %s/km.bin (%s)", dir, strerror(errno));
	if (rc) return rc;

	// rest.bin (rest.hpp:197-221)
	const RestHost& r = m->rest;
	std::vector<uint64_t> keys(r.count);
	std::vector<int32_t> counts(r.count), h2i(r.map_size), pre(r.pre_buffer_size);
	if (r.count) {
		CU(cudaMemcpy(keys.data(), m->d_rest_keys, r.count * 8, cudaMemcpyDeviceToHost));
		CU(cudaMemcpy(counts.data(), m->d_rest_counts, r.count * 4, cudaMemcpyDeviceToHost));
	}
	CU(cudaMemcpy(h2i.data(), m->d_hash2index, (size_t)r.map_size * 4, cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(pre.data(), m->d_pre_buffer, (size_t)r.pre_buffer_size * 4, cudaMemcpyDeviceToHost));
	const int sg = (r.k - r.pre_len) / 4;
	std::vector<uint8_t> suffix((size_t)r.suff_bin_size);
	for (uint64_t i = 0; i < r.count; i++)
		for (int b = 0; b < sg; b++) suffix[i * sg + b] = (uint8_t)(keys[i] >> (8 * (sg - 1 - b)));
	FILE* fr = fopen((base + "/rest.bin").c_str(), "wb");
	if (!fr) return fail(KMX_EIO, "cannot write %s/rest.bin (%s)", dir, strerror(errno));
	int32_t hdr[4] = { r.k, r.pre_len, r.map_size, r.pre_buffer_size };
	fwrite(hdr, 4, 4, fr);
	fwrite(&r.suff_bin_size, 8, 1, fr);
	fwrite(&r.count, 8, 1, fr);
	fwrite(h2i.data(), 4, h2i.size(), fr);
	fwrite(pre.data(), 4, pre.size(), fr);
	fwrite(suffix.data(), 1, suffix.size(), fr);
	fwrite(counts.data(), 4, counts.size(), fr);
	if (fclose(fr) != 0) return fail(KMX_EIO, "close %s/rest.bin (%s)", dir, strerror(errno));
	return KMX_OK;
}

// host -> device on `stream`, complete on return.  (A plain cudaMemcpy from pageable memory may
// return while the DMA is still in flight, and the model's streams are non-blocking: kernels
// launched on them would not wait for it.)
static int h2d_sync(void* d_ptr, const void* h_ptr, size_t bytes, cudaStream_t stream) {
	if (bytes == 0) return KMX_OK;
	CU(cudaMemcpyAsync(d_ptr, h_ptr, bytes, cudaMemcpyHostToDevice, stream));
	CU(cudaStreamSynchronize(stream));
	return KMX_OK;
}

static int read_to_device(FILE* f, void* d_ptr, uint64_t bytes, std::vector<uint8_t>& tmp, cudaStream_t stream) {
	const size_t padded = (size_t)((bytes + 7) & ~7ULL);
	if (tmp.size() < padded) tmp.resize(padded);
	if (bytes && fread(tmp.data(), 1, bytes, f) != bytes) return fail(KMX_EFORMAT, "km.bin is shorter than its header implies");
	memset(tmp.data() + bytes, 0, padded - bytes);
	return h2d_sync(d_ptr, tmp.data(), padded, stream);
}

static int load_into(kmx_model* m, const std::string& base) {
	int rc = model_attach_device(m);
	if (rc) return rc;
	CU(cudaSetDevice(m->device));
	FILE* f = fopen((base + "/km.bin").c_str(), "rb");
	if (!f) return fail(KMX_EIO, "cannot open %s/km.bin (%s)", base.c_str(), strerror(errno));
	bool ok = fread(&m->km_kmers, 8, 1, f) == 1;
	for (int i = 0; i < m->bf_num; i++) ok = ok && fread(&m->kmer_counts[i], 8, 1, f) == 1;
	if (!ok) {
		fclose(f);
		return fail(KMX_EFORMAT, "%s/km.bin: truncated header", base.c_str());
	}
	rc = alloc_filters(m);
	if (!rc && cudaStreamSynchronize(m->x->stream) != cudaSuccess) rc = fail(KMX_ECUDA, "zero-fill of the filters failed");   // the copies below run on the default stream
	std::vector<uint8_t> tmp;
	for (int i = 0; i < m->bf_num && !rc; i++) {
		rc = read_to_device(f, m->d_bf[i], m->bytes[i], tmp, m->x->stream);
		if (!rc) rc = read_to_device(f, m->d_bf_back[i], m->bytes[3 + i], tmp, m->x->stream);
	}
	if (!rc) rc = read_to_device(f, m->d_km_back, m->bytes[7], tmp, m->x->stream);
	const uint64_t words = cell_words(m->bytes[6]);
	uint32_t* d_val = nullptr;
	uint32_t* d_tag = nullptr;
	if (!rc && (cudaMalloc(&d_val, (words + 2) * 4) != cudaSuccess || cudaMalloc(&d_tag, (words + 2) * 4) != cudaSuccess))
		rc = fail(KMX_ECUDA, "out of device memory while loading");
	for (int i = 0; i < m->n_bits && !rc; i++) {
		rc = read_to_device(f, d_val, m->bytes[6], tmp, m->x->stream);
		if (!rc) rc = read_to_device(f, d_tag, m->bytes[6], tmp, m->x->stream);
		if (!rc) {
			cudaError_t e = launch_merge_cells(d_val, d_tag, words, m->d_cells[i], m->x->stream);
			if (e == cudaSuccess) e = cudaStreamSynchronize(m->x->stream);
			if (e != cudaSuccess) rc = fail(KMX_ECUDA, "merge kernel: %s", cudaGetErrorString(e));
		}
	}
	cudaFree(d_val);
	cudaFree(d_tag);
	fclose(f);
	if (rc) return rc;

	// rest.bin (rest.hpp:163-195)
	FILE* fr = fopen((base + "/rest.bin").c_str(), "rb");
	if (!fr) return fail(KMX_EIO, "cannot open %s/rest.bin (%s)", base.c_str(), strerror(errno));
	RestHost& r = m->rest;
	int32_t hdr[4];
	ok = fread(hdr, 4, 4, fr) == 4 && fread(&r.suff_bin_size, 8, 1, fr) == 1 && fread(&r.count, 8, 1, fr) == 1;
	if (ok) {
		r.k = hdr[0]; r.pre_len = hdr[1]; r.map_size = hdr[2]; r.pre_buffer_size = hdr[3];
		ok = r.k >= 3 && r.k <= 32 && r.pre_len >= 1 && r.pre_len <= r.k && r.map_size == (1 << (2 * r.pre_len)) && r.pre_buffer_size >= 1 &&
		     r.pre_buffer_size <= r.map_size + 1 && r.suff_bin_size == r.count * (uint64_t)((r.k - r.pre_len) / 4) && (r.k - r.pre_len) % 4 == 0;
	}
	if (!ok) {
		fclose(fr);
		return fail(KMX_EFORMAT, "%s/rest.bin: bad header", base.c_str());
	}
	const int sg = (r.k - r.pre_len) / 4;
	std::vector<int32_t> h2i(r.map_size), pre(r.pre_buffer_size), counts(r.count);
	std::vector<uint8_t> suffix((size_t)r.suff_bin_size);
	ok = read_exact(fr, h2i.data(), h2i.size() * 4) && read_exact(fr, pre.data(), pre.size() * 4) &&
	     read_exact(fr, suffix.data(), suffix.size()) && read_exact(fr, counts.data(), counts.size() * 4);
	fclose(fr);
	if (!ok) return fail(KMX_EFORMAT, "%s/rest.bin: truncated", base.c_str());
	// rebuild the full keys: entry e of group g carries the prefix p with hash2index[p] == g
	std::vector<uint64_t> keys(r.count);
	for (int p = 0; p < r.map_size; p++) {
		int g = h2i[p];
		if (g < 0) continue;
		if (g + 1 >= r.pre_buffer_size) return fail(KMX_EFORMAT, "%s/rest.bin: group index out of range", base.c_str());
		for (int64_t e = pre[g]; e < pre[g + 1]; e++) {
			if (e < 0 || (uint64_t)e >= r.count) return fail(KMX_EFORMAT, "%s/rest.bin: entry index out of range", base.c_str());
			uint64_t s = 0;
			for (int b = 0; b < sg; b++) s = (s << 8) | suffix[(size_t)e * sg + b];
			keys[e] = ((uint64_t)p << (8 * sg)) | s;
		}
	}
	m->k = r.k;
	DA(&m->d_hash2index, (size_t)r.map_size * 4, m->x->stream);
	DA(&m->d_pre_buffer, ((size_t)r.pre_buffer_size + 1) * 4, m->x->stream);
	DA(&m->d_rest_keys, (r.count + 1) * 8, m->x->stream);
	DA(&m->d_rest_counts, (r.count + 1) * 4, m->x->stream);
	if ((rc = h2d_sync(m->d_hash2index, h2i.data(), h2i.size() * 4, m->x->stream))) return rc;
	if ((rc = h2d_sync(m->d_pre_buffer, pre.data(), pre.size() * 4, m->x->stream))) return rc;
	if ((rc = h2d_sync(m->d_rest_keys, keys.data(), r.count * 8, m->x->stream))) return rc;
	if ((rc = h2d_sync(m->d_rest_counts, counts.data(), r.count * 4, m->x->stream))) return rc;
	if ((rc = build_rest_side_tables(m))) return rc;
	CU(cudaStreamSynchronize(m->x->stream));
	fill_dev_model(m);
	fill_info(m);
	m->built = true;
	return KMX_OK;
}

extern "C" kmx_model* kmx_load(const char* dir) {
	if (!dir) {
		fail(KMX_EARG, "null directory");
		return nullptr;
	}
	std::string base(dir);
	FILE* fh = fopen((base + "/header").c_str(), "r");
	if (!fh) {
		fail(KMX_EIO, "load_model: cant't open the header of the model ! (%s/header)", dir);
		return nullptr;
	}
	char key[128];
	int v[4] = { 0, 0, 0, 0 };
	bool ok = true;
	for (int i = 0; i < 4; i++) ok = ok && fscanf(fh, "%127s %d", key, &v[i]) == 2;   // kmodel.hpp:686-691: positional
	fclose(fh);
	if (!ok) {
		fail(KMX_EFORMAT, "%s/header: expected four 'name value' lines", dir);
		return nullptr;
	}
	if (require_gpu(nullptr)) return nullptr;
	kmx_model* m = kmx_create(v[2], v[3], v[0], v[1]);
	if (!m) return nullptr;
	if (load_into(m, base) != KMX_OK) {
		kmx_destroy(m);
		return nullptr;
	}
	return m;
}

// =========================================================================================
// retrieval (kmodel.hpp:90-116)
// =========================================================================================
// device-resident batches: in steps of 16 Mi queries so that the deferred-query list (sized for the
// worst case, every query deferred) stays at 256 MiB; the list comes from the stream-ordered pool
template <bool ASCII>
static int query_device(kmx_model* m, const void* d_in, size_t stride, size_t n, int32_t* d_out, void* stream) {
	if (!m || (n && (!d_in || !d_out))) return fail(KMX_EARG, "null argument");
	if (!m->built) return fail(KMX_ESTATE, "model is not initialised");
	if (ASCII && stride < (size_t)m->k) return fail(KMX_EARG, "stride %zu shorter than k=%d", stride, m->k);
	if (n > 0x7FFFFFFFULL) return fail(KMX_ERANGE, "batch of %zu: the reference's loop index is an int (kmodel.hpp:91)", n);
	if (n == 0) return KMX_OK;
	cudaStream_t s = stream ? (cudaStream_t)stream : m->x->stream;
	const size_t step = 1u << 24;
	DeferredQuery* d_defer = nullptr;
	unsigned int* d_defer_n = nullptr;
	uint64_t* d_packed = nullptr;
	DA(&d_defer, std::min(n, step) * sizeof(DeferredQuery), s);
	DA(&d_defer_n, sizeof(unsigned int), s);
	if (ASCII) DA(&d_packed, std::min(n, step) * 8, s);
	for (size_t off = 0; off < n; off += step) {
		const size_t cnt = std::min(step, n - off);
		if (ASCII) CU(launch_query_ascii(m->dm, (const char*)d_in + off * stride, stride, cnt, d_out + off, d_packed, d_defer, d_defer_n, m->sm_count, s));
		else CU(launch_query_packed(m->dm, (const uint64_t*)d_in + off, cnt, d_out + off, nullptr, d_defer, d_defer_n, m->sm_count, s));
	}
	dev_free(d_defer, s);
	dev_free(d_defer_n, s);
	dev_free(d_packed, s);
	return KMX_OK;
}

extern "C" int kmx_query_packed_device(kmx_model* m, const uint64_t* d_kmers, size_t n, int32_t* d_out, void* stream) {
	return query_device<false>(m, d_kmers, 8, n, d_out, stream);
}

extern "C" int kmx_query_ascii_device(kmx_model* m, const char* d_flat, size_t stride, size_t n, int32_t* d_out, void* stream) {
	return query_device<true>(m, d_flat, stride, n, d_out, stream);
}

static int ensure_staging(kmx_model* m, size_t item_bytes, bool need_h_in, bool need_h_out) {
	DevCtx* x = m->x;
	const size_t items = 1u << 22;                         // queries per pipeline step
	const size_t bytes = items * item_bytes;
	x->stage_items = items;
	bool fresh = false;
	for (int s = 0; s < 2; s++) {
		if (x->stage_bytes < bytes) {
			dev_free(x->d_in[s], x->stream);
			x->d_in[s] = nullptr;
			DA(&x->d_in[s], bytes, x->stream);
			if (x->h_in[s]) {
				cudaFreeHost(x->h_in[s]);
				x->h_in[s] = nullptr;
			}
			fresh = true;
		}
		if (need_h_in && !x->h_in[s]) CU(cudaMallocHost(&x->h_in[s], std::max(bytes, x->stage_bytes)));
		if (need_h_out && !x->h_out[s]) CU(cudaMallocHost((void**)&x->h_out[s], items * 4));
		if (!x->d_out[s]) {
			DA(&x->d_out[s], items * 4, x->stream);
			DA(&x->d_pack[s], items * 8, x->stream);
			DA(&x->d_defer[s], items * sizeof(DeferredQuery), x->stream);
			DA(&x->d_defer_n[s], sizeof(unsigned int), x->stream);
			fresh = true;
		}
	}
	if (x->stage_bytes < bytes) x->stage_bytes = bytes;
	if (fresh) CU(cudaStreamSynchronize(x->stream));      // the staging buffers are used from both streams
	return KMX_OK;
}

static bool is_pinned(const void* p) {
	cudaPointerAttributes at;
	if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
		cudaGetLastError();
		return false;
	}
	return at.type == cudaMemoryTypeHost;
}

// host buffers -> device in steps of 4 Mi queries, two slots on two streams so that the copy of
// step i+1 overlaps the kernel and the read-back of step i; pinned caller buffers skip staging
static int query_host(kmx_model* m, const void* in, size_t item_bytes, size_t stride, size_t n, int32_t* out, int32_t* path, bool ascii) {
	if (!m || (n && (!in || (!out && !path)))) return fail(KMX_EARG, "null argument");
	if (!m->built) return fail(KMX_ESTATE, "model is not initialised");
	if (n > 0x7FFFFFFFULL) return fail(KMX_ERANGE, "batch of %zu: the reference's loop index is an int (kmodel.hpp:91)", n);
	if (n == 0) return KMX_OK;
	std::lock_guard<std::mutex> lock(m->x->query_mu);       // kmer_to_occ may be called from several threads (kmodel.hpp:90-98 is read-only)
	CU(cudaSetDevice(m->device));
	int32_t* res = path ? path : out;
	const bool in_pinned = is_pinned(in), out_pinned = is_pinned(res);
	int rc = ensure_staging(m, item_bytes, !in_pinned, !out_pinned);
	if (rc) return rc;
	cudaStream_t st[2] = { m->x->stream, m->x->stream2 };
	size_t done_off[2] = { 0, 0 }, done_n[2] = { 0, 0 };
	const uint8_t* src = (const uint8_t*)in;
	int slot = 0;
	for (size_t off = 0; off < n; off += m->x->stage_items, slot ^= 1) {
		const size_t cnt = std::min(m->x->stage_items, n - off);
		if (done_n[slot]) {                                // slot busy with an earlier step: drain it
			CU(cudaEventSynchronize(m->x->ev_done[slot]));
			if (!out_pinned) memcpy(res + done_off[slot], m->x->h_out[slot], done_n[slot] * 4);
			done_n[slot] = 0;
		}
		const void* h_src = src + off * item_bytes;
		if (!in_pinned) {
			memcpy(m->x->h_in[slot], h_src, cnt * item_bytes);
			h_src = m->x->h_in[slot];
		}
		CU(cudaMemcpyAsync(m->x->d_in[slot], h_src, cnt * item_bytes, cudaMemcpyHostToDevice, st[slot]));
		if (ascii) CU(launch_query_ascii(m->dm, (const char*)m->x->d_in[slot], stride, cnt, m->x->d_out[slot], m->x->d_pack[slot], m->x->d_defer[slot], m->x->d_defer_n[slot], m->sm_count, st[slot]));
		else CU(launch_query_packed(m->dm, (const uint64_t*)m->x->d_in[slot], cnt, path ? nullptr : m->x->d_out[slot], path ? m->x->d_out[slot] : nullptr,
		                            m->x->d_defer[slot], m->x->d_defer_n[slot], m->sm_count, st[slot]));
		CU(cudaMemcpyAsync(out_pinned ? (void*)(res + off) : (void*)m->x->h_out[slot], m->x->d_out[slot], cnt * 4, cudaMemcpyDeviceToHost, st[slot]));
		CU(cudaEventRecord(m->x->ev_done[slot], st[slot]));
		done_off[slot] = off;
		done_n[slot] = cnt;
	}
	for (int s = 0; s < 2; s++) {
		if (done_n[s]) {
			CU(cudaEventSynchronize(m->x->ev_done[s]));
			if (!out_pinned) memcpy(res + done_off[s], m->x->h_out[s], done_n[s] * 4);
		}
	}
	return KMX_OK;
}

extern "C" int kmx_query_packed(kmx_model* m, const uint64_t* kmers, size_t n, int32_t* out) {
	return query_host(m, kmers, 8, 8, n, out, nullptr, false);
}

extern "C" int kmx_query_path_packed(kmx_model* m, const uint64_t* kmers, size_t n, int32_t* path) {
	return query_host(m, kmers, 8, 8, n, nullptr, path, false);
}

extern "C" int kmx_query_ascii(kmx_model* m, const char* flat, size_t stride, size_t n, int32_t* out) {
	if (m && stride < (size_t)m->k) return fail(KMX_EARG, "stride %zu shorter than k=%d", stride, m->k);
	return query_host(m, flat, stride, stride, n, out, nullptr, true);
}

// =========================================================================================
// host-side known-answer entry points (no GPU needed)
// =========================================================================================
extern "C" uint64_t kmx_host_murmur64(const void* key, int len, uint32_t seed) {
	// MurmurHash64A on raw bytes, via the same block/tail split the kernels use
	const uint8_t* p = (const uint8_t*)key;
	uint64_t h = (uint64_t)seed ^ ((uint64_t)len * kMurM);
	const int nblocks = len / 8;
	for (int b = 0; b < nblocks; b++) {
		uint64_t w;
		memcpy(&w, p + 8 * b, 8);
		w *= kMurM; w ^= w >> 47; w *= kMurM;
		h ^= w; h *= kMurM;
	}
	if (len & 7) {
		uint64_t t = 0;
		memcpy(&t, p + 8 * nblocks, len & 7);
		h ^= t; h *= kMurM;
	}
	h ^= h >> 47; h *= kMurM; h ^= h >> 47;
	return h;
}

extern "C" uint64_t kmx_host_hash_packed(uint64_t kmer, int len, uint32_t seed) {
	if (len < 1 || len > 32) return 0;
	HashPrep p;
	hash_prepare(reverse_bases(kmer & mask2(len), len), len, p);
	return hash_finish(p, len, seed);
}

extern "C" uint64_t kmx_host_canonical(uint64_t kmer, int k) {
	uint64_t r;
	return canonical(kmer & mask2(k), k, &r);
}

extern "C" uint32_t kmx_host_seed(int i) { return h_seeds[i & 127]; }

extern "C" int kmx_host_occubin(int max_counter, int n_hash, int32_t* occ2bin, int32_t* bin2mean) {
	std::vector<int32_t> a, b;
	int rc = occubin_tables(max_counter, n_hash, a, b);
	if (rc) return rc;
	memcpy(occ2bin, a.data(), a.size() * 4);
	memcpy(bin2mean, b.data(), b.size() * 4);
	return KMX_OK;
}

extern "C" void kmx_host_sizes(const uint64_t kmer_counts[3], int bf_num, uint64_t km_kmers, int n_hash, uint64_t bytes[8]) {
	model_sizes(kmer_counts, bf_num, km_kmers, n_hash, bytes);
}

extern "C" uint64_t kmx_host_fastmod(uint64_t h, uint64_t d) { return d >= 2 ? fastmod(h, make_fastmod(d)) : 0; }   // filter lengths are >= 8 bits

// the closed form insert_kernel uses for reorder_buffer (kmodel.hpp:529-540)
extern "C" int kmx_host_reorder(const uint8_t* failed, int n, int32_t* perm) {
	int F = 0;
	for (int i = 0; i < n; i++) F += failed[i] ? 1 : 0;
	std::vector<int32_t> hole(F > 0 ? F : 1);
	int excl = 0;
	for (int i = 0; i < n; i++) {
		if (failed[i]) {
			if (i < F) perm[i] = i;
			excl++;
		} else if (i < F) {
			hole[i - excl] = i;
		}
	}
	excl = 0;
	for (int i = 0; i < n; i++) {
		if (failed[i]) {
			if (i >= F) perm[hole[F - excl - 1]] = i;
			excl++;
		}
	}
	return F;
}
