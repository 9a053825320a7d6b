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
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "kmx_internal.h"

using namespace kmx;

// =========================================================================================
// errors, device selection, launch counter
// =========================================================================================
static thread_local char g_err[512] = "";
static thread_local int g_err_code = 0;
static int g_device = 0;                       // process-wide default (kmx_set_device)
static thread_local int t_device = -1;         // per-thread override (threads of an in-process team build)
static std::atomic<unsigned long long> g_launches{ 0 };
constexpr size_t kL2FetchDefault = 0;         // 0: leave the device's setting alone (set from the A/B, KMX_L2_FETCH)

namespace kmx {

int set_error(int code, const char* fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	g_err_code = code;
	return code;
}
const char* last_error() { return g_err; }
int last_error_code() { return g_err_code; }

bool trace_on() {
	static int on = -1;
	if (on < 0) on = getenv("KMX_TRACE") ? 1 : 0;
	return on == 1;
}

void note_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }

int current_device() { return t_device >= 0 ? t_device : g_device; }
void set_thread_device(int ordinal) { t_device = ordinal; }

// Stream-ordered allocations from the device's default memory pool, which is told to keep
// freed blocks: a rebuild (or the next query staging) reuses them instead of paying
// cudaMalloc/cudaFree (hundreds of microseconds each, and cudaFree synchronises the device).
int dev_alloc_impl(void** p, size_t bytes, cudaStream_t s) {
	static std::atomic<bool> pool_ready[64];
	int dev = 0;
	CU(cudaGetDevice(&dev));
	if (dev < 64 && !pool_ready[dev].load()) {
		cudaMemPool_t pool;
		CU(cudaDeviceGetDefaultMemPool(&pool, dev));
		uint64_t keep = ~0ULL;
		CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
		pool_ready[dev].store(true);
	}
	if (trace_on() && bytes >= (256ULL << 20)) {
		// KMX_TRACE: large requests that the pool could not serve from what it holds show up as milliseconds of host time here
		const auto t0 = std::chrono::high_resolution_clock::now();
		CU(cudaMallocAsync(p, bytes, s));
		const double ms = 1e3 * std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count();
		cudaMemPool_t pool;
		uint64_t reserved = 0, used = 0;
		if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
			cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved);
			cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used);
		}
		fprintf(stderr, "[kmx]   alloc %8.1f MB at %p took %.3f ms (pool: %.1f MB reserved, %.1f MB in use)\n", bytes / 1048576.0, *p, ms, reserved / 1048576.0,
		        used / 1048576.0);
		return KMX_OK;
	}
	CU(cudaMallocAsync(p, bytes ? bytes : 8, s));
	return KMX_OK;
}
void dev_free(void* p, cudaStream_t s) {
	if (p) cudaFreeAsync(p, s);
}

int require_gpu(int* sm_count) {
	// cudaGetDeviceProperties costs milliseconds: ask once per device
	static std::mutex mu;
	static int cached_sm[64] = { 0 };
	static int cached_n = -1;
	std::lock_guard<std::mutex> lock(mu);
	if (cached_n < 0) {
		int n = 0;
		cudaError_t e = cudaGetDeviceCount(&n);
		if (e != cudaSuccess || n <= 0) {
			cudaGetLastError();
			return set_error(KMX_ENOGPU, "no usable CUDA device (libkmx has no CPU path): %s", e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
		}
		cached_n = n;
	}
	const int dev = current_device();
	if (dev >= cached_n || dev >= 64) return set_error(KMX_ENOGPU, "device %d requested, %d present", dev, cached_n);
	CU(cudaSetDevice(dev));
	if (cached_sm[dev] == 0) {
		cudaDeviceProp prop;
		CU(cudaGetDeviceProperties(&prop, dev));
		if (prop.major < 10) return set_error(KMX_ENOGPU, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
		cached_sm[dev] = prop.multiProcessorCount;
		// The path is random 8-byte probes of arrays far larger than the L2, and the L2 fetches at least 64 bytes from DRAM per
		// missing 32-byte sector (ncu on the HC14 shape: 60-95 DRAM bytes read per probe).  KMX_L2_FETCH=32|64|128 passes the
		// cudaLimitMaxL2FetchGranularity hint; on B200 it changes nothing (profiles/r2_e_l2_fetch_ab.log), so it is left alone.
		size_t fetch = kL2FetchDefault;
		if (const char* e = getenv("KMX_L2_FETCH")) fetch = (size_t)atoi(e);
		if (fetch == 32 || fetch == 64 || fetch == 128) {
			if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, fetch) != cudaSuccess) cudaGetLastError();
		}
	}
	if (sm_count) *sm_count = cached_sm[dev];
	return KMX_OK;
}

}  // namespace kmx

#define fail kmx::set_error

constexpr int kQueryL2Default = 7;            // A/B on the HC14 shape (profiles/r2_d_query_l2_ab.log): 2.55 -> 2.79 G queries/s

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
void kmx::model_sizes(const uint64_t kmer_counts[3], int bf_num, uint64_t km_kmers, int n_hash, uint64_t bytes[8]) {
	const int hb = n_hash - 1, hk = n_hash - 2;
	for (int i = 0; i < 8; i++) bytes[i] = 0;
	for (int i = 0; i < bf_num; i++) {
		bytes[i] = (uint64_t)(kmer_counts[i] / 5.5 * hb);
		bytes[3 + i] = (kmer_counts[i] >> 3) * (uint64_t)hk;
	}
	bytes[6] = (km_kmers >> 4) * (uint64_t)n_hash;
	bytes[7] = (km_kmers >> 4) * (uint64_t)hk;
}

int kmx::rest_prefix_len(int k) {              // rest.hpp:78-83
	for (int i = 7; i >= 3; i--)
		if ((k - i) % 4 == 0) return i;
	return 3;
}

// =========================================================================================
// execution contexts (kmx_internal.h: DevCtx), pooled per device
// =========================================================================================
namespace {
std::mutex g_ctx_mu;
std::vector<DevCtx*> g_ctx_free;
}

int kmx::ctx_acquire(DevCtx** out) {
	{
		std::lock_guard<std::mutex> lock(g_ctx_mu);
		for (size_t i = 0; i < g_ctx_free.size(); i++) {
			if (g_ctx_free[i]->device == current_device()) {
				*out = g_ctx_free[i];
				g_ctx_free.erase(g_ctx_free.begin() + i);
				return KMX_OK;
			}
		}
	}
	DevCtx* c = new DevCtx();
	c->device = current_device();
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

void kmx::ctx_release(DevCtx* c) {
	if (!c) return;
	cudaStreamSynchronize(c->stream);
	cudaStreamSynchronize(c->stream2);
	std::lock_guard<std::mutex> lock(g_ctx_mu);
	g_ctx_free.push_back(c);
}

void kmx::free_model_device(kmx_model* m) {
	cudaStream_t s = m->x->stream;
	if (m->mslab) {
		// team build: filters and coupled arrays are carved from one peer-mapped slab
		cudaStreamSynchronize(s);
		slab_release(m->mslab, m->mslab_bytes, m->device);
		m->mslab = nullptr;
		for (int i = 0; i < 3; i++) m->d_bf[i] = m->d_bf_back[i] = nullptr;
		m->d_km_back = nullptr;
		for (int i = 0; i < kMaxArrays; i++) m->d_cells[i] = nullptr;
	}
	if (m->rslab) {
		cudaStreamSynchronize(s);
		slab_release(m->rslab, m->rslab_bytes, m->device);
		m->rslab = nullptr;
		m->d_rest_keys = nullptr;
		m->d_rest_counts = nullptr;
	}
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

// sizes from m->kmer_counts / m->km_kmers (kmodel.hpp:402-456) and the reference's failure corners on degenerate ones
int kmx::check_model_sizes(kmx_model* m) {
	model_sizes(m->kmer_counts, m->bf_num, m->km_kmers, m->n_hash, m->bytes);
	for (int i = 0; i < m->bf_num; i++) {
		if (m->bytes[i] == 0 || m->bytes[3 + i] == 0)
			return fail(KMX_ERANGE, "count class %d holds %llu k-mers, which makes a Bloom filter of length 0: the reference (g++ >= 5) dies with "
			            "std::bad_array_new_length at kmodel.hpp:413-417 (`new uint8_t[0]{ 0 }`) and would divide by zero at kmodel.hpp:378",
			            m->ci + i, (unsigned long long)m->kmer_counts[i]);
	}
	if (m->bytes[6] == 0 || m->bytes[7] == 0)
		return fail(KMX_ERANGE, "%llu k-mers for the coupled arrays make arrays of length 0: the reference dies with std::bad_array_new_length at kmodel.hpp:441-447",
		            (unsigned long long)m->km_kmers);
	return KMX_OK;
}

// allocate + zero every filter of the model (single GPU: stream-ordered pool allocations)
static int alloc_filters(kmx_model* m) {
	int rc = check_model_sizes(m);
	if (rc) return rc;
	for (int i = 0; i < m->bf_num; i++) {
		DA(&m->d_bf[i], pad8(m->bytes[i]), m->x->stream);
		CU(cudaMemsetAsync(m->d_bf[i], 0, pad8(m->bytes[i]), m->x->stream));
		DA(&m->d_bf_back[i], pad8(m->bytes[3 + i]), m->x->stream);
		CU(cudaMemsetAsync(m->d_bf_back[i], 0, pad8(m->bytes[3 + i]), m->x->stream));
	}
	DA(&m->d_km_back, pad8(m->bytes[7]), m->x->stream);
	CU(cudaMemsetAsync(m->d_km_back, 0, pad8(m->bytes[7]), m->x->stream));
	const uint64_t words = cell_words(m->bytes[6]);
	for (int i = 0; i < m->n_bits; i++) {
		DA(&m->d_cells[i], (words + 1) * 8, m->x->stream);
		CU(cudaMemsetAsync(m->d_cells[i], 0, (words + 1) * 8, m->x->stream));
	}
	return KMX_OK;
}

void kmx::fill_dev_model(kmx_model* m) {
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
	// query probes of a model beyond the L2: km_back stays resident, arrays and Bloom filters stream
	d.query_l2 = (2ULL * m->n_bits * m->bytes[6] + m->bytes[7] + m->bytes[0] + m->bytes[1] + m->bytes[2]) > (192ULL << 20) ? kQueryL2Default : 0;
	if (const char* e = getenv("KMX_QUERY_L2")) d.query_l2 = atoi(e) & 7;
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

void kmx::fill_info(kmx_model* m) {
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
extern "C" unsigned long long kmx_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" const char* kmx_version(void) { return "kmx 0.2 (sm_100a)"; }

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
	t_device = -1;
	CU(cudaSetDevice(ordinal));
	return require_gpu(nullptr);                          // architecture check + per-device settings, once
}

// the GPUs one KModel::init is spread over inside this process (kmx_set_devices; default from the environment: KMX_GPUS=N)
namespace {
std::mutex g_team_mu;
std::vector<int> g_team_devices;
bool g_team_set = false;
}  // namespace

const std::vector<int>& kmx::team_devices() {
	std::lock_guard<std::mutex> lock(g_team_mu);
	if (!g_team_set) {
		g_team_set = true;
		if (const char* e = getenv("KMX_GPUS")) {
			const int want = atoi(e), have = kmx_device_count();
			for (int d = 0; d < want && d < have && d < kMaxRanks; d++) g_team_devices.push_back(d);
			if (g_team_devices.size() < 2) g_team_devices.clear();
		}
	}
	return g_team_devices;
}

extern "C" int kmx_set_devices(const int* ordinals, int n) {
	if (n < 0 || n > kMaxRanks || (n > 0 && !ordinals)) return fail(KMX_EARG, "kmx_set_devices: 0..%d device ordinals", kMaxRanks);
	const int have = kmx_device_count();
	std::vector<int> devs;
	for (int i = 0; i < n; i++) {
		if (ordinals[i] < 0 || ordinals[i] >= have) return fail(KMX_ENOGPU, "device %d requested, %d present", ordinals[i], have);
		for (int d : devs)
			if (d == ordinals[i]) return fail(KMX_EARG, "kmx_set_devices: device %d listed twice", d);
		devs.push_back(ordinals[i]);
	}
	std::lock_guard<std::mutex> lock(g_team_mu);
	g_team_set = true;
	g_team_devices = devs.size() >= 2 ? devs : std::vector<int>();
	if (devs.size() == 1) g_device = devs[0];
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

int kmx::model_attach_device(kmx_model* m) {
	if (m->x) return KMX_OK;
	int rc = require_gpu(&m->sm_count);
	if (rc) return rc;
	m->device = current_device();
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
	for (kmx_model* r : m->replicas) kmx_destroy(r);
	m->replicas.clear();
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
	// only the tail (header, at most 255 + 8 bytes, and the markers) is parsed on the host; the LUT goes straight into its
	// vector and the signature map, which the listing does not use (kmc_file.cpp:449), is never read
	const int pfd = open(pre_name.c_str(), O_RDONLY);
	if (pfd < 0) {
		fail(KMX_EIO, "can't open the kmer_data_base %s (%s)", db_base, strerror(errno));
		return nullptr;
	}
	struct stat pst;
	uint64_t pre_size = 0;
	const uint64_t kTail = 512;                          // header_offset is one byte: header + offset word + marker <= 267 bytes
	std::vector<uint8_t> tail;
	char head4[4] = { 0, 0, 0, 0 };
	bool ok = fstat(pfd, &pst) == 0 && (pre_size = (uint64_t)pst.st_size) >= 32;
	if (ok) {
		tail.resize((size_t)std::min<uint64_t>(kTail, pre_size));
		ok = pread(pfd, tail.data(), tail.size(), (off_t)(pre_size - tail.size())) == (ssize_t)tail.size() && pread(pfd, head4, 4, 0) == 4;
	}
	if (!ok || memcmp(head4, "KMCP", 4) != 0 || memcmp(tail.data() + tail.size() - 4, "KMCP", 4) != 0) {
		close(pfd);
		fail(KMX_EFORMAT, "%s is not a KMC prefix file", pre_name.c_str());
		return nullptr;
	}
	// `pre` = a view of the file's last bytes with the offsets of the whole file: pre[i] is valid for i >= pre_size - tail.size()
	struct TailView {
		const uint8_t* base;
		uint64_t first;
		const uint8_t& operator[](uint64_t i) const { return base[i - first]; }
	} pre{ tail.data(), pre_size - tail.size() };
	struct CloseFd {
		int fd;
		~CloseFd() { close(fd); }
	} close_pre{ pfd };
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
	db->orig_min_count = h.min_count;
	db->orig_max_count = h.max_count;
	db->both_strands = p[36] == 0;                        // kmc_file.cpp:208-209: the byte says "forward strand only"
	h.both_strands = db->both_strands ? 1 : 0;
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
	{
		uint8_t* dst = (uint8_t*)db->lut.data();
		uint64_t want = h.lut_entries * 8, got = 0;      // the guard word is overwritten below, as the reference does
		while (got < want) {
			const ssize_t r = pread(pfd, dst + got, want - got, (off_t)(4 + got));
			if (r <= 0) break;
			got += (uint64_t)r;
		}
		if (got != want) {
			fail(KMX_EIO, "short read on %s", pre_name.c_str());
			delete db;
			return nullptr;
		}
	}
	db->lut[h.lut_entries] = h.total_kmers + 1;           // kmc_file.cpp:223
	db->pre_name = pre_name;
	db->sig_offset = 4 + (h.lut_entries + 1) * 8;         // kmc_file.cpp:224-226: the map follows the LUT and its guard word
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
	db->device = current_device();
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

// .kmc_suf record area (records [rec_lo, rec_hi)) -> HBM: reader threads pread() 2 MiB chunks into pinned bounce buffers and
// push each with its own stream, so page-cache copies and PCIe transfers overlap.  The page-cache copy (5-6 GB/s per
// thread) is the slow half: one thread per host core, up to kReaders (a team build in N processes asks for cores / N each).
int kmx::db_upload_range(kmx_db* db, uint64_t rec_lo, uint64_t rec_hi, int reader_threads) {
	if (!db) return fail(KMX_EARG, "null database");
	if (rec_hi > db->info.total_kmers) rec_hi = db->info.total_kmers;
	if (rec_lo > rec_hi) rec_lo = rec_hi;
	if (db->d_suf && db->rec_lo <= rec_lo && db->rec_hi >= rec_hi) return KMX_OK;     // already resident
	if (db->d_suf) return fail(KMX_ESTATE, "database holds records [%llu, %llu) on the device, [%llu, %llu) requested", (unsigned long long)db->rec_lo,
	                           (unsigned long long)db->rec_hi, (unsigned long long)rec_lo, (unsigned long long)rec_hi);
	int rc = require_gpu(&db->sm_count);
	if (rc) return rc;
	db->device = current_device();
	auto t0 = std::chrono::high_resolution_clock::now();
	DevCtx* x = nullptr;
	if ((rc = ctx_acquire(&x))) return rc;
	struct Release {
		DevCtx* x;
		~Release() { ctx_release(x); }
	} release{ x };
	std::lock_guard<std::mutex> lock(g_bounce.mu);
	if ((rc = g_bounce.acquire())) return rc;
	const uint64_t rec_bytes = db->info.record_bytes;
	const uint64_t byte_lo = rec_lo * rec_bytes, bytes = (rec_hi - rec_lo) * rec_bytes;
	db->suf_alloc = (size_t)((bytes + 15) & ~15ULL) + 16;
	DA(&db->d_suf, db->suf_alloc, x->stream);
	DA(&db->d_lut, db->lut.size() * 8, x->stream);
	CU(cudaMemcpyAsync(db->d_lut, db->lut.data(), db->lut.size() * 8, cudaMemcpyHostToDevice, x->stream));
	CU(cudaMemsetAsync(db->d_suf + (bytes & ~15ULL), 0, db->suf_alloc - (bytes & ~15ULL), x->stream));
	CU(cudaStreamSynchronize(x->stream));                // the allocation is usable from the reader streams now
	TRACE(t0, "upload: buffers ready");
	const uint64_t n_chunks = (bytes + kChunk - 1) / kChunk;
	// (copying out of a shared mapping of the file instead of pread() is slower here: 28.7 against 24.5 ms for half of the HC14
	// records per rank, profiles/r2_j_bench_team_hc14_n2_mmap*.log)
	int want_thr = std::max(1, std::min<int>(kReaders, reader_threads > 0 ? reader_threads : (int)std::thread::hardware_concurrency()));
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
				ssize_t r = pread(db->fd, b + got, len - got, (off_t)(4 + byte_lo + off + got));
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
	for (int t = 0; t < n_thr; t++) {
		if (status[t] != KMX_OK) {
			dev_free(db->d_suf, x->stream);
			dev_free(db->d_lut, x->stream);
			db->d_suf = nullptr;
			db->d_lut = nullptr;
			return fail(status[t], status[t] == KMX_EIO ? "short read on the .kmc_suf file" : "host-to-device copy of the database failed");
		}
	}
	db->rec_lo = rec_lo;
	db->rec_hi = rec_hi;
	db->ms_upload = (float)(1e3 * std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - t0).count());
	TRACE(t0, "upload: done");
	return KMX_OK;
}

extern "C" int kmx_db_upload(kmx_db* db) {
	if (!db) return fail(KMX_EARG, "null database");
	return db_upload_range(db, 0, db->info.total_kmers, 0);
}

// the share of a team build's rank: whole decode tiles, [n_tiles * rank / world, n_tiles * (rank + 1) / world)
extern "C" int kmx_db_upload_share(kmx_db* db, int rank, int world) {
	if (!db || world < 1 || rank < 0 || rank >= world) return fail(KMX_EARG, "kmx_db_upload_share: bad rank / world");
	const uint64_t n_tiles = (db->info.total_kmers + kTile - 1) / kTile;
	const uint64_t lo = n_tiles * (uint64_t)rank / (uint64_t)world, hi = n_tiles * (uint64_t)(rank + 1) / (uint64_t)world;
	const int cores = (int)std::thread::hardware_concurrency();
	return db_upload_range(db, lo * kTile, hi * kTile, std::max(2, cores / world));
}

extern "C" void kmx_db_close(kmx_db* db) {
	if (!db) return;
	if (db->d_suf || db->d_lut) {
		cudaSetDevice(db->device);
		cudaDeviceSynchronize();
		dev_free(db->d_suf, nullptr);
		dev_free(db->d_lut, nullptr);
		dev_free(db->d_sigmap, nullptr);
	}
	if (db->fd >= 0) close(db->fd);
	delete db;
}

DevDb kmx::dev_db(const kmx_db* db) {
	DevDb d;
	memset(&d, 0, sizeof(d));
	d.suf = db->d_suf - db->rec_lo * db->info.record_bytes;      // indexed by the global record number; only [rec_lo, rec_hi) is touched
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
	DevScope scope(s);                                    // frees on every return path
	if ((rc = scope.alloc(&d_cnt, n_tiles * 4))) return rc;
	if ((rc = scope.alloc(&d_off, (n_tiles + 1) * 8))) return rc;
	if ((rc = scope.alloc(&d_k, total * 8))) return rc;
	if ((rc = scope.alloc(&d_c, total * 4))) return rc;
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
	*n_out = listed;
	return KMX_OK;
}

// =========================================================================================
// build: KModel::init (kmodel.hpp:57-86), in stages (kmx_team.cu interleaves its exchanges with them):
//   encode (count + Bloom + item stream)  ->  insert  ->  rest table
// =========================================================================================
// bucket index over the top key bits (about one entry per bucket) + the per-prefix false-hit table
// that reproduces the inclusive upper bound of KRestData::check_kmer (rest.hpp:233-237)
int kmx::build_rest_side_tables(kmx_model* m) {
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

// scratch / scratch_bytes: dead device memory the sort may use instead of allocating its temporary buffers (may be null / 0)
static int build_rest_table(kmx_model* m, const uint64_t* d_surv_kmer, const uint32_t* d_surv_occ, uint64_t n, int32_t* h_groups, void* scratch, size_t scratch_bytes) {
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
	DevScope scope(s);
	int rc;
	if ((rc = scope.alloc(&d_first, (size_t)r.map_size * 4))) return rc;
	if ((rc = scope.alloc(&d_groups, 4))) return rc;
	const size_t temp_bytes = radix_sort_temp_bytes(n);
	if (temp_bytes && temp_bytes <= scratch_bytes) d_temp = scratch;
	else if (temp_bytes && (rc = scope.alloc(&d_temp, temp_bytes))) return rc;
	CU(launch_radix_sort_pairs(d_temp, d_surv_kmer, m->d_rest_keys, d_surv_occ, (uint32_t*)m->d_rest_counts, n, 2 * m->k, s));
	CU(launch_rest_index(m->d_rest_keys, n, 2 * (m->k - r.pre_len), r.map_size, d_first, m->d_hash2index, m->d_pre_buffer, d_groups, s));
	CU(cudaMemcpyAsync(h_groups, d_groups, 4, cudaMemcpyDeviceToHost, s));
	return KMX_OK;
}

// cudaMalloc blocks other GPUs map are kept for the life of the process: cudaMalloc / cudaFree and cudaIpcOpenMemHandle cost
// milliseconds each, a rebuild with the same geometry reuses them (and the mappings its peers hold stay valid).
namespace {
struct CachedSlab {
	void* ptr;
	size_t bytes;
	int device;
};
std::mutex g_slab_mu;
std::vector<CachedSlab> g_slab_free;
}  // namespace

int kmx::slab_acquire(void** out, size_t bytes, int device) {
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
	if (cudaMalloc(out, bytes) == cudaSuccess) return KMX_OK;
	cudaGetLastError();
	{
		// out of memory: give the cached slabs of other sizes on this device back and try once more (a peer process may still
		// hold a mapping of them from an earlier build of another geometry; it no longer touches it)
		std::lock_guard<std::mutex> lock(g_slab_mu);
		for (size_t q = 0; q < g_slab_free.size();) {
			if (g_slab_free[q].device == device) {
				cudaFree(g_slab_free[q].ptr);
				g_slab_free.erase(g_slab_free.begin() + q);
			} else {
				q++;
			}
		}
	}
	CU(cudaMalloc(out, bytes));
	return KMX_OK;
}

void kmx::slab_release(void* ptr, size_t bytes, int device) {
	std::lock_guard<std::mutex> lock(g_slab_mu);
	g_slab_free.push_back(CachedSlab{ ptr, bytes, device });
}

void kmx::build_state_free(kmx_model* m) {
	BuildState& b = m->bs;
	if (!m->x) return;
	cudaStream_t s = m->x->stream;
	InsertArgs& a = b.a;
	const bool team = b.team != nullptr;
	if (team) team_state_free(m);                         // item shard, ping-pong buffers, control block, survivor list live in its slab
	if (!team) {
		if (a.rest_kmer != b.d_item_kmer) {               // n_bits == 1: a list of its own (see build_stage_insert_setup)
			dev_free(a.rest_kmer, s);
			dev_free(a.rest_occ, s);
		}
		dev_free(b.d_item_kmer, s);
		dev_free(b.d_item_occ, s);
		for (int q = 0; q < 2; q++) {
			dev_free(a.buf_kmer[q], s);
			dev_free(a.buf_occ[q], s);
		}
		dev_free(a.ctl, s);
	}
	dev_free(a.status, s); dev_free(a.excl_rank, s); dev_free(a.holepos, s); dev_free(a.list[0], s); dev_free(a.list[1], s); dev_free(a.list[2], s);
	dev_free(a.tile_fail, s); dev_free(a.resv, s); dev_free(a.claim, s);
	b = BuildState();
}

// stage 1: pass 1 (class histogram, kmodel.hpp:423-434), filter allocation (kmodel.hpp:402-456),
// pass 2 (Bloom inserts + array-bound stream, kmodel.hpp:68-74)
static int build_stage_encode(kmx_model* m, kmx_db* db) {
	BuildState& b = m->bs;
	b.wall0 = std::chrono::high_resolution_clock::now();
	int rc = model_attach_device(m);
	if (rc) return rc;
	TRACE(b.wall0, "device attached");
	CU(cudaSetDevice(m->device));
	rc = kmx_db_upload(db);
	if (rc) return rc;
	TRACE(b.wall0, "database on device");
	if (db->device != m->device) return fail(KMX_EARG, "database is on device %d, model on device %d", db->device, m->device);
	if (db->rec_lo != 0 || db->rec_hi != db->info.total_kmers) return fail(KMX_ESTATE, "only a share of the database is on the device (team build); KModel::init on one GPU needs all of it");
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
	DevScope scope(s);
	if ((rc = scope.alloc(&d_tile_cnt, (n_tiles + 1) * 4))) return rc;
	if ((rc = scope.alloc(&d_tile_off, (n_tiles + 1) * 8))) return rc;
	if ((rc = scope.alloc(&d_count, sizeof(CountOut)))) return rc;
	CU(cudaEventRecord(ev[0], s));
	CU(cudaMemsetAsync(d_count, 0, sizeof(CountOut), s));
	CU(launch_count(d, m->ci, m->cs, m->bf_num, d_count, d_tile_cnt, 0, n_tiles, m->sm_count, s));
	CU(launch_tile_scan(d_tile_cnt, n_tiles, d_tile_off, s));
	CountOut& cnt = m->x->h_pinned->count;
	CU(cudaMemcpyAsync(&cnt, d_count, sizeof(cnt), cudaMemcpyDeviceToHost, s));
	CU(cudaEventRecord(ev[1], s));
	TRACE(b.wall0, "count pass queued");
	CU(cudaStreamSynchronize(s));                        // sync 1 of 3: the sizes depend on the counts
	TRACE(b.wall0, "count pass done");
	if (cnt.bad_count)
		return fail(KMX_ERANGE, "%llu records have a count below ci=%d or above cs=%d: the reference indexes out of bounds there (kmodel.hpp:427, occu_bin.hpp:70)",
		            (unsigned long long)cnt.bad_count, m->ci, m->cs);
	uint64_t bf_kmers = 0;
	for (int i = 0; i < m->bf_num; i++) {
		m->kmer_counts[i] = cnt.class_count[i];
		bf_kmers += cnt.class_count[i];
	}
	m->km_kmers = total - bf_kmers;                       // kmodel.hpp:433: header total, not the listed count
	if ((rc = alloc_filters(m))) return rc;
	m->rest.k = m->k;
	m->rest.pre_len = rest_prefix_len(m->k);
	fill_dev_model(m);

	b.n_items = cnt.array_bound;
	// + n_bits: the stream doubles as the survivor list (build_stage_insert_setup), which can hold that many stale-slot duplicates
	DA(&b.d_item_kmer, (b.n_items + m->n_bits + 1) * 8, s);
	DA(&b.d_item_occ, (b.n_items + m->n_bits + 1) * 4, s);
	ItemRoute route;
	memset(&route, 0, sizeof(route));
	route.kmer[0] = b.d_item_kmer;
	route.occ[0] = b.d_item_occ;
	route.base = 0;
	route.n_active = 1;
	route.n_bits = m->n_bits;
	CU(launch_encode(d, m->dm, d_tile_off, route, 0, n_tiles, m->sm_count, s));
	CU(cudaEventRecord(ev[2], s));
	return KMX_OK;
}

// stage 2a: buffers of the greedy insert (kmodel.hpp:508-573).  team = true: the buffers other GPUs write into (item shard,
// survivor ping-pong, control block, barrier flags, survivor list) were carved from the team's exchange slab by kmx_team.cu
// and are already in m->bs.a; only the private scratch is allocated here.
int kmx::build_stage_insert_setup(kmx_model* m, int rank, int n_active, bool team) {
	BuildState& b = m->bs;
	cudaStream_t s = m->x->stream;
	InsertArgs& a = b.a;
	if (!team) memset(&a, 0, sizeof(a));
	const uint64_t batch_items = (uint64_t)m->n_bits << kBucketLog;
	b.n_batches = (b.n_items + batch_items - 1) / batch_items;
	a.rank = rank;
	a.n_active = n_active;
	a.n_items = b.n_items;
	if (team && rank >= n_active) return KMX_OK;          // this rank owns no coupled array: it launches no insert kernel
	if (!team) {
		a.item_kmer = b.d_item_kmer;
		a.item_occ = b.d_item_occ;
		for (int q = 0; q < 2; q++) {
			DA(&a.buf_kmer[q], batch_items * 8, s);
			DA(&a.buf_occ[q], batch_items * 4, s);
		}
		DA(&a.ctl, sizeof(InsertCtl), s);
		CU(cudaMemsetAsync(a.ctl, 0, sizeof(InsertCtl), s));
		for (int q = 0; q < 2; q++) {
			a.peer_buf_kmer[q][rank] = a.buf_kmer[q];
			a.peer_buf_occ[q][rank] = a.buf_occ[q];
		}
		a.peer_ctl[rank] = a.ctl;
		a.peer_flags[rank] = nullptr;
	}
	if (const char* e = getenv("KMX_TEST_EPOCH_START")) {    // lets a small test cross the epoch wrap-around
		const unsigned int start = (unsigned int)atoi(e);
		CU(cudaMemcpyAsync(&a.ctl->epoch, &start, sizeof(start), cudaMemcpyHostToDevice, s));
		CU(cudaStreamSynchronize(s));
	}
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
	// contested items: merged reserve/commit passes (hc14 shape: insert 163 -> 155 ms, RS shape: 3.12 -> 3.08 ms)
	a.merged = 1;
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
	// what THIS GPU probes (its share of the coupled arrays + its copy of km_back) well beyond the L2 (126 MB): those random
	// sectors should not wash the insert's hot structures out of it (measured on the hc14 shape, one GPU: evict-first cell
	// loads + reductions 182 -> 168 ms; evict-first on km_back as well: 171 ms)
	const uint64_t my_arrays = n_active > 1 ? (uint64_t)buckets_per_batch(rank < n_active ? rank : 0, n_active, m->n_bits) : (uint64_t)m->n_bits;
	a.stream_cells = (2ULL * my_arrays * m->bytes[6] + m->bytes[7]) > (192ULL << 20) ? 3 : 0;
	if (const char* e = getenv("KMX_STREAM_CELLS")) a.stream_cells = atoi(e) & 15;
	a.max_iterations = kBucket + 64;
	a.phase_round = -1;
	if (const char* e = getenv("KMX_PHASE_ROUND")) a.phase_round = atoi(e);
	CU(insert_grid_size(&b.grid, m->sm_count));
	if (!team) {
		// Survivor list.  Batch b appends its survivors after its last round, when every item of the batch has long been read
		// from the stream (round 0 reads the stream, later rounds the ping-pong buffers), and at most as many items survive as
		// were read: survivors of the batches up to b never reach the items of batch b + 1.  So the list lives IN the item
		// stream, from its start -- no second allocation of the worst-case size (NA12878 shape: 36 GB), no list that grows
		// between launches with a host round trip each, and every launch of a build can be queued at once.  The stale-slot
		// duplicates of the last, partial batch (kmodel.hpp:520-540; fewer than n_bits) are written before that batch is
		// read: at least n_bits items of batch 0 were accepted (the first item of every bucket meets an empty array), so they
		// stay below its first item too.  With n_bits == 1 the last round IS round 0 and reads the stream while survivors are
		// appended: that geometry gets a list of its own.
		b.rest_cap = b.n_items + m->n_bits;
		if (m->n_bits >= 2) {
			a.rest_kmer = b.d_item_kmer;
			a.rest_occ = b.d_item_occ;
		} else {
			DA(&a.rest_kmer, b.rest_cap * 8, s);
			DA(&a.rest_occ, b.rest_cap * 4, s);
		}
	}
	return KMX_OK;
}

// stage 2b: the launches (64 batches each); returns with the control block on the host
int kmx::build_stage_insert_run(kmx_model* m) {
	BuildState& b = m->bs;
	cudaStream_t s = m->x->stream;
	InsertArgs& a = b.a;
	InsertCtl& ctl = m->x->h_pinned->ctl;
	memset(&ctl, 0, sizeof(ctl));
	const bool participates = a.rank < a.n_active;
	CU(cudaEventRecord(m->x->ev_build[5], s));
	if (b.n_items > 0 && participates) {
		const uint64_t chunk = 64;                         // batches per launch
		a.rest_cap = b.rest_cap;
		for (uint64_t b0 = 0; b0 < b.n_batches; b0 += chunk) {
			a.first_batch = b0;
			a.n_batches = std::min<uint64_t>(chunk, b.n_batches - b0);
			CU(launch_insert(m->dm, a, b.grid, s));
		}
		CU(cudaMemcpyAsync(&ctl, a.ctl, sizeof(ctl), cudaMemcpyDeviceToHost, s));
	}
	CU(cudaEventRecord(m->x->ev_build[3], s));
	return KMX_OK;
}

// stage 3: rest table (rest.hpp:157-161) from the survivors at d_rest_* (n of them), statistics, clean-up
static int build_stage_finish(kmx_model* m, const uint64_t* d_rest_kmer, const uint32_t* d_rest_occ, uint64_t rest_n) {
	BuildState& b = m->bs;
	cudaStream_t s = m->x->stream;
	cudaEvent_t* ev = m->x->ev_build;
	const InsertCtl ctl = m->x->h_pinned->ctl;
	int32_t& groups = m->x->h_pinned->groups;
	// the insert's scratch is dead: back to the pool first, the rest table's allocations come out of it.  The survivors sit at
	// the head of the item stream (build_stage_insert_setup); its tail is dead too and serves as the sort's temporary storage
	// (NA12878 shape: 1.9 GB that would otherwise be a pool allocation between two timed events)
	for (int q = 0; q < 2; q++) {
		dev_free(b.a.buf_kmer[q], s);
		dev_free(b.a.buf_occ[q], s);
		b.a.buf_kmer[q] = nullptr;
		b.a.buf_occ[q] = nullptr;
	}
	dev_free(b.a.resv, s);
	dev_free(b.a.claim, s);
	b.a.resv = nullptr;
	b.a.claim = nullptr;
	void* scratch = nullptr;
	size_t scratch_bytes = 0;
	if (b.d_item_kmer && d_rest_kmer == b.d_item_kmer) {
		const size_t stream_bytes = (size_t)(b.n_items + m->n_bits + 1) * 8, head = up256((size_t)rest_n * 8);
		if (stream_bytes > head) {
			scratch = (uint8_t*)b.d_item_kmer + head;
			scratch_bytes = stream_bytes - head;
		}
	} else if (b.d_item_kmer) {                           // n_bits == 1: the list is a buffer of its own, the whole stream is dead
		scratch = b.d_item_kmer;
		scratch_bytes = (size_t)(b.n_items + m->n_bits + 1) * 8;
	}
	int rc = build_rest_table(m, d_rest_kmer, d_rest_occ, rest_n, &groups, scratch, scratch_bytes);
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
	for (int i = 0; i < 12; i++) f.insert_phase_cycles[i] = ctl.phase_cycles[i];
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
	if (!rc) {
		TRACE(m->bs.wall0, "encode + insert queued");
		if (cudaStreamSynchronize(m->x->stream) != cudaSuccess) rc = fail(KMX_ECUDA, "the build stream failed: %s", cudaGetErrorString(cudaGetLastError()));   // sync 2 of 3: the sort needs the survivor count
		else if (m->x->h_pinned->ctl.error)
			rc = fail(KMX_ECUDA, "insert kernel stopped with error %u (1: iteration cap, 2: survivor list overflow, 3: peer GPU timed out)", m->x->h_pinned->ctl.error);
		TRACE(m->bs.wall0, "insert done");
	}
	if (!rc) rc = build_stage_finish(m, m->bs.a.rest_kmer, m->bs.a.rest_occ, m->x->h_pinned->ctl.rest_n);
	if (rc) build_state_free(m);
	return rc;
}

// KModel::init(db_file).  With more than one device selected (kmx_set_devices, or KMX_GPUS=N in the environment) the build is
// spread over them inside this process (kmx_team.cu): one host thread per GPU, peer access instead of IPC, the model ends
// up replicated and kmer_to_occ batches are sharded over the replicas.
extern "C" int kmx_init_from_kmc(kmx_model* m, const char* db_base) {
	if (!m) return fail(KMX_EARG, "null model");
	int rc = require_gpu(nullptr);
	if (rc) return rc;
	const std::vector<int> devs = team_devices();
	if (devs.size() >= 2) return team_build_in_process(m, db_base, devs);
	kmx_db* db = kmx_db_open(db_base);
	if (!db) return last_error_code() ? last_error_code() : KMX_EIO;     // unreadable file vs. not a KMC database: kmx_db_open's own code
	rc = kmx_init_from_db(m, db);
	kmx_db_close(db);
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
	const bool header_ok = fprintf(fh, "number_hash %d\nnumber_bit %d\nci %d\ncs %d\n", m->n_hash, m->n_bits, m->ci, m->cs) > 0;   // kmodel.hpp:175-180
	if (fclose(fh) != 0 || !header_ok) return fail(KMX_EIO, "cannot write %s/header (%s)", dir, strerror(errno));
	FILE* f = fopen((base + "/km.bin").c_str(), "wb");
	if (!f) return fail(KMX_EIO, "cannot write %s/km.bin (%s)", dir, strerror(errno));
	std::vector<uint8_t> tmp;
	int rc = KMX_OK;
	bool head_ok = fwrite(&m->km_kmers, 8, 1, f) == 1;
	for (int i = 0; i < m->bf_num; i++) head_ok = head_ok && fwrite(&m->kmer_counts[i], 8, 1, f) == 1;
	if (!head_ok) rc = fail(KMX_EIO, "short write on %s/km.bin (%s)", dir, strerror(errno));
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
	if (!rc && (fflush(f) != 0 || ferror(f))) rc = fail(KMX_EIO, "write error on %s/km.bin (%s)", dir, strerror(errno));
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
	bool ok = fwrite(hdr, 4, 4, fr) == 4 && fwrite(&r.suff_bin_size, 8, 1, fr) == 1 && fwrite(&r.count, 8, 1, fr) == 1 &&
	          fwrite(h2i.data(), 4, h2i.size(), fr) == h2i.size() && fwrite(pre.data(), 4, pre.size(), fr) == pre.size() &&
	          fwrite(suffix.data(), 1, suffix.size(), fr) == suffix.size() && fwrite(counts.data(), 4, counts.size(), fr) == counts.size() &&
	          fflush(fr) == 0 && !ferror(fr);
	if (!ok) {
		const int err = errno;
		fclose(fr);
		return fail(KMX_EIO, "short write on %s/rest.bin (%s)", dir, strerror(err));
	}
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
	unsigned int* d_counters = nullptr;
	uint64_t* d_packed = nullptr;
	uint32_t* d_dirty = nullptr;
	DevScope scope(s);                                    // the scratch goes back to the pool on every return path
	int rc;
	if ((rc = scope.alloc(&d_defer, std::min(n, step) * sizeof(DeferredQuery)))) return rc;
	if ((rc = scope.alloc(&d_counters, 2 * sizeof(unsigned int)))) return rc;
	if (ASCII && (rc = scope.alloc(&d_packed, std::min(n, step) * 8))) return rc;
	if (ASCII && (rc = scope.alloc(&d_dirty, std::min(n, step) * 4))) return rc;
	for (size_t off = 0; off < n; off += step) {
		const size_t cnt = std::min(step, n - off);
		if (ASCII) CU(launch_query_ascii(m->dm, (const char*)d_in + off * stride, stride, cnt, d_out + off, d_packed, d_defer, d_dirty, d_counters, m->sm_count, s));
		else CU(launch_query_packed(m->dm, (const uint64_t*)d_in + off, cnt, d_out + off, nullptr, d_defer, d_counters, m->sm_count, s));
	}
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
			DA(&x->d_dirty[s], items * 4, x->stream);
			DA(&x->d_defer_n[s], 2 * sizeof(unsigned int), x->stream);
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
static int query_host_one(kmx_model* m, const void* in, size_t item_bytes, size_t stride, size_t n, int32_t* out, int32_t* path, bool ascii) {
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
		if (ascii) CU(launch_query_ascii(m->dm, (const char*)m->x->d_in[slot], stride, cnt, m->x->d_out[slot], m->x->d_pack[slot], m->x->d_defer[slot], m->x->d_dirty[slot], m->x->d_defer_n[slot], m->sm_count, st[slot]));
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

// A model built over several GPUs inside this process (kmx_set_devices / KMX_GPUS) is replicated on each of them: the batch
// is split into contiguous shards, one host thread per replica (SURVEY.md 8e "Query": no per-query communication)
static int query_host(kmx_model* m, const void* in, size_t item_bytes, size_t stride, size_t n, int32_t* out, int32_t* path, bool ascii) {
	if (!m || (n && (!in || (!out && !path)))) return fail(KMX_EARG, "null argument");
	if (!m->built) return fail(KMX_ESTATE, "model is not initialised");
	if (n > 0x7FFFFFFFULL) return fail(KMX_ERANGE, "batch of %zu: the reference's loop index is an int (kmodel.hpp:91)", n);
	const size_t n_rep = 1 + m->replicas.size();
	if (n_rep == 1 || n < (size_t)1 << 16) return query_host_one(m, in, item_bytes, stride, n, out, path, ascii);
	std::vector<int> rcs(n_rep, KMX_OK);
	std::vector<std::string> msgs(n_rep);
	std::vector<std::thread> pool;
	auto part = [&](size_t r) {
		kmx_model* rep = r == 0 ? m : m->replicas[r - 1];
		const size_t lo = n * r / n_rep, hi = n * (r + 1) / n_rep;
		rcs[r] = query_host_one(rep, (const uint8_t*)in + lo * item_bytes, item_bytes, stride, hi - lo, out ? out + lo : nullptr, path ? path + lo : nullptr, ascii);
		if (rcs[r]) msgs[r] = last_error();
	};
	for (size_t r = 1; r < n_rep; r++) pool.emplace_back(part, r);
	part(0);
	for (auto& t : pool) t.join();
	for (size_t r = 0; r < n_rep; r++)
		if (rcs[r]) return fail(rcs[r], "replica %zu: %s", r, msgs[r].c_str());
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
