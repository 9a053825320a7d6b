// kmx_count.cu -- FASTQ -> KMC database on the GPU (row N3 of SURVEY.md section 8f).
//
// The reference does not count k-mers itself: it shells out to the external `kmc` binary
// (main.cpp:136-140: kmc -k -t -ci -cs <input> <output> <tmp>), which is not part of the
// reference tree.  This file provides that stage so that the command line works end to end:
// canonical k-mers (min of forward / reverse complement, as KMC without -b) of every read, windows
// containing a non-ACGT base skipped, counts saturated at cs, k-mers seen fewer than ci times
// dropped, written in the KMC 2/3 on-disk layout the listing reader parses
// (kmc_file.cpp:177-235, 428-515): one bin, records sorted by k-mer.  (A real KMC run spreads the
// records over minimiser bins, so its listing order -- and therefore the greedy array contents --
// differs; the k-mer set and the counts are the same.)
//
// Pipeline: file bytes -> HBM; newline index (count / scan / scatter); one thread per read rolls
// the forward and reverse-complement words over the sequence line and emits one u64 per window
// (sentinel for invalid windows); radix sort + run-length encode (CUB); filter/clamp; records
// and prefix LUT are packed on the device and written by the host.
#include <cuda_runtime.h>
#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <string>
#include <vector>
#include "../../include/kmx.h"
#include "kmx_core.cuh"

namespace kmx {
int set_error(int code, const char* fmt, ...);   // kmx_host.cu
}
using namespace kmx;

namespace {

constexpr int kTextTile = 4096;

__global__ void newline_count_kernel(const uint8_t* __restrict__ text, uint64_t n, uint32_t* __restrict__ tile_cnt) {
	const uint64_t tiles = (n + kTextTile - 1) / kTextTile;
	for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
		const uint64_t lo = tile * kTextTile, hi = min(n, lo + kTextTile);
		uint32_t c = 0;
		for (uint64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) c += text[i] == '\n';
		c = __reduce_add_sync(0xffffffffu, c);
		__shared__ uint32_t s[8];
		if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
		__syncthreads();
		if (threadIdx.x == 0) {
			uint32_t t = 0;
			for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s[w];
			tile_cnt[tile] = t;
		}
		__syncthreads();
	}
}

// line_start[j] = offset of the first byte of line j (line 0 starts at 0); one warp per tile keeps order
__global__ void newline_scatter_kernel(const uint8_t* __restrict__ text, uint64_t n, const uint64_t* __restrict__ tile_off,
                                       uint64_t* __restrict__ line_start) {
	const uint64_t tiles = (n + kTextTile - 1) / kTextTile;
	const int lane = threadIdx.x & 31;
	const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
	for (uint64_t tile = warp; tile < tiles; tile += n_warps) {
		const uint64_t lo = tile * kTextTile, hi = min(n, lo + kTextTile);
		uint64_t at = tile_off[tile] + 1;                  // +1: line 0 is implicit
		for (uint64_t base = lo; base < hi; base += 32) {
			const uint64_t i = base + lane;
			const bool nl = i < hi && text[i] == '\n';
			const uint32_t m = __ballot_sync(0xffffffffu, nl);
			if (nl) line_start[at + __popc(m & ((1u << lane) - 1u))] = i + 1;
			at += __popc(m);
		}
	}
}

// windows per read (sequence line = line 4r+1 of a 4-line FASTQ record)
__global__ void window_count_kernel(const uint8_t* __restrict__ text, const uint64_t* __restrict__ line_start, uint64_t n_reads, int k,
                                    uint64_t* __restrict__ windows) {
	for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t a = line_start[4 * r + 1];
		uint64_t b = line_start[4 * r + 2] - 1;           // the '\n'
		if (b > a && text[b - 1] == '\r') b--;
		const uint64_t len = b - a;
		windows[r] = len >= (uint64_t)k ? len - k + 1 : 0;
	}
}

constexpr uint64_t kInvalid = ~0ULL;

__global__ void extract_kernel(const uint8_t* __restrict__ text, const uint64_t* __restrict__ line_start, const uint64_t* __restrict__ win_off,
                               uint64_t n_reads, int k, uint64_t* __restrict__ out) {
	const uint64_t mask = mask2(k);
	for (uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t a = line_start[4 * r + 1];
		uint64_t b = line_start[4 * r + 2] - 1;
		if (b > a && text[b - 1] == '\r') b--;
		if (b - a < (uint64_t)k) continue;
		uint64_t o = win_off[r];
		uint64_t fwd = 0, rc = 0;
		int run = 0;
		for (uint64_t i = a; i < b; i++) {
			const uint8_t ch = text[i];
			int c = ch == 'A' || ch == 'a' ? 0 : ch == 'C' || ch == 'c' ? 1 : ch == 'G' || ch == 'g' ? 2 : ch == 'T' || ch == 't' ? 3 : -1;
			if (c < 0) {
				run = 0;
				fwd = rc = 0;
			} else {
				run++;
				fwd = ((fwd << 2) | (uint64_t)c) & mask;
				rc = (rc >> 2) | ((uint64_t)(3 - c) << (2 * (k - 1)));
			}
			if (i - a + 1 >= (uint64_t)k) out[o++] = run >= k ? (fwd < rc ? fwd : rc) : kInvalid;
		}
	}
}

struct KeepCount {
	uint32_t ci, cx;
	__host__ __device__ bool operator()(const uint32_t& c) const { return c >= ci && c <= cx; }
};

__global__ void flag_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, uint64_t n, uint32_t ci, uint32_t cx,
                            uint8_t* __restrict__ flags) {
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		flags[i] = keys[i] != kInvalid && counts[i] >= ci && counts[i] <= cx;
}

// record bytes (big-endian suffix + little-endian counter, kmc_file.cpp:447-494) of every kept k-mer
__global__ void pack_records_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ counts, uint64_t n, int suffix_bytes,
                                    int counter_bytes, uint32_t cs, uint8_t* __restrict__ rec) {
	const int rb = suffix_bytes + counter_bytes;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t v = keys[i];
		uint32_t c = counts[i];
		if (c > cs) c = cs;                                // -cs: maximal value of a counter
		uint8_t* p = rec + i * rb;
		for (int b = 0; b < suffix_bytes; b++) p[b] = (uint8_t)(v >> (8 * (suffix_bytes - 1 - b)));
		for (int b = 0; b < counter_bytes; b++) p[suffix_bytes + b] = (uint8_t)(c >> (8 * b));
	}
}

// lut[p] = index of the first record whose prefix (key >> 8*suffix_bytes) is >= p
__global__ void lut_kernel(const uint64_t* __restrict__ keys, uint64_t n, int shift, uint64_t n_prefix, uint64_t* __restrict__ lut) {
	for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p <= n_prefix; p += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t lo = 0, hi = n;
		while (lo < hi) {
			const uint64_t mid = (lo + hi) >> 1;
			if ((keys[mid] >> shift) < p) lo = mid + 1; else hi = mid;
		}
		lut[p] = lo;
	}
}

struct Scratch {
	std::vector<void*> ptrs;
	~Scratch() {
		for (void* p : ptrs) cudaFree(p);
	}
	template <class T>
	cudaError_t alloc(T** p, size_t bytes) {
		cudaError_t e = cudaMalloc((void**)p, bytes ? bytes : 8);
		if (e == cudaSuccess) ptrs.push_back(*p);
		return e;
	}
};

bool write_all(FILE* f, const void* p, size_t n) { return n == 0 || fwrite(p, 1, n, f) == n; }

}  // namespace

#define CUC(call)                                                                                                  \
	do {                                                                                                           \
		cudaError_t e__ = (call);                                                                                  \
		if (e__ != cudaSuccess) return set_error(KMX_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
	} while (0)

extern "C" int kmx_count_fastq(const char* const* fastq_paths, int n_files, int k, int ci, int cs, const char* out_base, kmx_count_info_t* info) {
	if (!fastq_paths || n_files < 1 || !out_base) return set_error(KMX_EARG, "null argument");
	if (k < 3 || k > 32 || ci < 1 || cs < ci || cs > 65535) return set_error(KMX_EARG, "unsupported k=%d ci=%d cs=%d (3 <= k <= 32, 1 <= ci <= cs <= 65535)", k, ci, cs);
	if (kmx_device_count() < 1) return set_error(KMX_ENOGPU, "no usable CUDA device (libkmx has no CPU path)");
	// ---- read the files (plain-text 4-line FASTQ) ----
	std::vector<uint8_t> text;
	for (int f = 0; f < n_files; f++) {
		FILE* fp = fopen(fastq_paths[f], "rb");
		if (!fp) return set_error(KMX_EIO, "cannot open %s (%s)", fastq_paths[f], strerror(errno));
		unsigned char magic[2] = { 0, 0 };
		size_t got = fread(magic, 1, 2, fp);
		if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
			fclose(fp);
			return set_error(KMX_EFORMAT, "%s is gzip-compressed; decompress it first (only plain-text FASTQ is read)", fastq_paths[f]);
		}
		fseeko(fp, 0, SEEK_END);
		const uint64_t sz = (uint64_t)ftello(fp);
		rewind(fp);
		const size_t old = text.size();
		text.resize(old + sz + 1);
		if (sz && fread(text.data() + old, 1, sz, fp) != sz) {
			fclose(fp);
			return set_error(KMX_EIO, "short read on %s", fastq_paths[f]);
		}
		fclose(fp);
		if (sz && text[old + sz - 1] != '\n') text[old + sz] = '\n';      // every file ends with a newline
		else text.resize(old + sz);
	}
	const uint64_t n_bytes = text.size();
	int dev = 0, sms = 148;
	CUC(cudaGetDevice(&dev));
	CUC(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
	Scratch S;
	uint8_t* d_text = nullptr;
	CUC(S.alloc(&d_text, n_bytes + 16));
	CUC(cudaMemcpy(d_text, text.data(), n_bytes, cudaMemcpyHostToDevice));
	std::vector<uint8_t>().swap(text);
	// ---- newline index ----
	const uint64_t tiles = (n_bytes + kTextTile - 1) / kTextTile;
	uint32_t* d_tile_cnt = nullptr;
	uint64_t* d_tile_off = nullptr;
	CUC(S.alloc(&d_tile_cnt, (tiles + 1) * 4));
	CUC(S.alloc(&d_tile_off, (tiles + 1) * 8));
	const int grid = sms * 8;
	uint64_t n_lines = 0;
	if (tiles) {
		newline_count_kernel<<<grid, 256>>>(d_text, n_bytes, d_tile_cnt);
		void* d_tmp = nullptr;
		size_t tmp_bytes = 0;
		CUC(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_tile_cnt, d_tile_off, (int64_t)tiles + 1));
		CUC(S.alloc(&d_tmp, tmp_bytes));
		CUC(cudaMemset(d_tile_cnt + tiles, 0, 4));
		CUC(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_tile_cnt, d_tile_off, (int64_t)tiles + 1));
		CUC(cudaMemcpy(&n_lines, d_tile_off + tiles, 8, cudaMemcpyDeviceToHost));
	}
	const uint64_t n_reads = n_lines / 4;
	uint64_t* d_line = nullptr;
	CUC(S.alloc(&d_line, (n_lines + 2) * 8));
	CUC(cudaMemset(d_line, 0, 8));
	if (tiles) newline_scatter_kernel<<<grid, 256>>>(d_text, n_bytes, d_tile_off, d_line);
	// ---- windows ----
	uint64_t* d_win = nullptr;
	uint64_t* d_win_off = nullptr;
	CUC(S.alloc(&d_win, (n_reads + 1) * 8));
	CUC(S.alloc(&d_win_off, (n_reads + 1) * 8));
	uint64_t n_windows = 0;
	if (n_reads) {
		window_count_kernel<<<grid, 256>>>(d_text, d_line, n_reads, k, d_win);
		CUC(cudaMemset(d_win + n_reads, 0, 8));
		void* d_tmp = nullptr;
		size_t tmp_bytes = 0;
		CUC(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, d_win, d_win_off, (int64_t)n_reads + 1));
		CUC(S.alloc(&d_tmp, tmp_bytes));
		CUC(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, d_win, d_win_off, (int64_t)n_reads + 1));
		CUC(cudaMemcpy(&n_windows, d_win_off + n_reads, 8, cudaMemcpyDeviceToHost));
	}
	uint64_t* d_keys = nullptr;
	uint64_t* d_sorted = nullptr;
	CUC(S.alloc(&d_keys, (n_windows + 1) * 8));
	CUC(S.alloc(&d_sorted, (n_windows + 1) * 8));
	if (n_windows) extract_kernel<<<grid, 256>>>(d_text, d_line, d_win_off, n_reads, k, d_keys);
	// ---- sort + run-length encode ----
	uint64_t* d_unique = nullptr;
	uint32_t* d_counts = nullptr;
	uint64_t* d_runs = nullptr;
	CUC(S.alloc(&d_unique, (n_windows + 1) * 8));
	CUC(S.alloc(&d_counts, (n_windows + 1) * 4));
	CUC(S.alloc(&d_runs, 8));
	uint64_t n_unique = 0;
	if (n_windows) {
		void* d_tmp = nullptr;
		size_t tmp_bytes = 0;
		CUC(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, d_keys, d_sorted, (int64_t)n_windows, 0, 64));
		CUC(S.alloc(&d_tmp, tmp_bytes));
		CUC(cub::DeviceRadixSort::SortKeys(d_tmp, tmp_bytes, d_keys, d_sorted, (int64_t)n_windows, 0, 64));
		void* d_tmp2 = nullptr;
		size_t tmp2 = 0;
		CUC(cub::DeviceRunLengthEncode::Encode(nullptr, tmp2, d_sorted, d_unique, d_counts, d_runs, (int64_t)n_windows));
		CUC(S.alloc(&d_tmp2, tmp2));
		CUC(cub::DeviceRunLengthEncode::Encode(d_tmp2, tmp2, d_sorted, d_unique, d_counts, d_runs, (int64_t)n_windows));
		CUC(cudaMemcpy(&n_unique, d_runs, 8, cudaMemcpyDeviceToHost));
	}
	// ---- filter (ci <= count), keep order ----
	uint8_t* d_flags = nullptr;
	uint64_t* d_kept = nullptr;
	uint32_t* d_kept_cnt = nullptr;
	uint64_t* d_nkept = nullptr;
	CUC(S.alloc(&d_flags, n_unique + 1));
	CUC(S.alloc(&d_kept, (n_unique + 1) * 8));
	CUC(S.alloc(&d_kept_cnt, (n_unique + 1) * 4));
	CUC(S.alloc(&d_nkept, 8));
	uint64_t n_kept = 0;
	if (n_unique) {
		flag_kernel<<<grid, 256>>>(d_unique, d_counts, n_unique, (uint32_t)ci, 0xFFFFFFFFu, d_flags);
		void* d_tmp = nullptr;
		size_t tmp_bytes = 0;
		CUC(cub::DeviceSelect::Flagged(nullptr, tmp_bytes, d_unique, d_flags, d_kept, d_nkept, (int64_t)n_unique));
		CUC(S.alloc(&d_tmp, tmp_bytes));
		CUC(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, d_unique, d_flags, d_kept, d_nkept, (int64_t)n_unique));
		CUC(cub::DeviceSelect::Flagged(d_tmp, tmp_bytes, d_counts, d_flags, d_kept_cnt, d_nkept, (int64_t)n_unique));
		CUC(cudaMemcpy(&n_kept, d_nkept, 8, cudaMemcpyDeviceToHost));
	}
	// ---- KMC 2/3 files ----
	int lut = k % 4 == 0 ? 4 : k % 4;                       // (k - lut) % 4 == 0
	if (lut < 3) lut += 4;
	if (n_kept > (1ULL << 24) && lut + 4 <= k && lut < 7) lut += 4;
	const int suffix_bytes = (k - lut) / 4;
	const int counter_bytes = cs < 256 ? 1 : 2;
	const int rb = suffix_bytes + counter_bytes;
	const uint64_t n_prefix = 1ULL << (2 * lut);
	uint8_t* d_rec = nullptr;
	uint64_t* d_lut = nullptr;
	CUC(S.alloc(&d_rec, n_kept * rb + 16));
	CUC(S.alloc(&d_lut, (n_prefix + 1) * 8));
	if (n_kept) pack_records_kernel<<<grid, 256>>>(d_kept, d_kept_cnt, n_kept, suffix_bytes, counter_bytes, (uint32_t)cs, d_rec);
	lut_kernel<<<(int)((n_prefix + 256) / 256), 256>>>(d_kept, n_kept, 8 * suffix_bytes, n_prefix, d_lut);
	CUC(cudaDeviceSynchronize());
	std::vector<uint8_t> rec((size_t)n_kept * rb);
	std::vector<uint64_t> lut_h(n_prefix + 1);
	if (n_kept) CUC(cudaMemcpy(rec.data(), d_rec, rec.size(), cudaMemcpyDeviceToHost));
	CUC(cudaMemcpy(lut_h.data(), d_lut, lut_h.size() * 8, cudaMemcpyDeviceToHost));
	const std::string base(out_base);
	FILE* fs = fopen((base + ".kmc_suf").c_str(), "wb");
	if (!fs) return set_error(KMX_EIO, "cannot write %s.kmc_suf (%s)", out_base, strerror(errno));
	bool ok = write_all(fs, "KMCS", 4) && write_all(fs, rec.data(), rec.size()) && write_all(fs, "KMCS", 4);
	ok = (fclose(fs) == 0) && ok;
	FILE* fp = fopen((base + ".kmc_pre").c_str(), "wb");
	if (!fp) return set_error(KMX_EIO, "cannot write %s.kmc_pre (%s)", out_base, strerror(errno));
	const uint32_t sig_len = 7;
	std::vector<uint32_t> sigmap((1u << (2 * sig_len)) + 1, 0);
	// header: kmer_length, mode, counter_size, lut_prefix_length, signature_len, min_count, max_count (u32 each),
	// total_kmers (u64), both-strands byte, padding, kmc_version = 0x200 as the last word (kmc_file.cpp:180-209)
	uint8_t hdr[68];
	memset(hdr, 0, sizeof(hdr));
	const uint32_t w[7] = { (uint32_t)k, 0u, (uint32_t)counter_bytes, (uint32_t)lut, sig_len, (uint32_t)ci, (uint32_t)cs };
	memcpy(hdr, w, 28);
	memcpy(hdr + 28, &n_kept, 8);
	const uint32_t version = 0x200, hdr_len = sizeof(hdr);
	memcpy(hdr + 64, &version, 4);
	ok = ok && write_all(fp, "KMCP", 4) && write_all(fp, lut_h.data(), lut_h.size() * 8) && write_all(fp, sigmap.data(), sigmap.size() * 4) &&
	     write_all(fp, hdr, sizeof(hdr)) && write_all(fp, &hdr_len, 4) && write_all(fp, "KMCP", 4);
	ok = (fclose(fp) == 0) && ok;
	if (!ok) return set_error(KMX_EIO, "short write on %s.kmc_pre/.kmc_suf", out_base);
	if (info) {
		info->n_reads = n_reads;
		info->n_windows = n_windows;
		info->n_unique = n_unique;
		info->n_kept = n_kept;
		info->lut_prefix_length = (uint32_t)lut;
		info->counter_size = (uint32_t)counter_bytes;
	}
	return KMX_OK;
}
