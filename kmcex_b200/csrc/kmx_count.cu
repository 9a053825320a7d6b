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
// (sentinel for invalid windows); the library's own radix sort (kmx_sort.cu, keys only); run heads and the -ci filter
// are two stream compactions (count per tile / scan / write); records and prefix LUT are packed on the device and
// written by the host.  Input files may be gzip-compressed (zlib on the host).  No library kernels.
#include <cuda_runtime.h>
#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <zlib.h>
#include <string>
#include <vector>
#include "../../include/kmx.h"
#include "kmx_core.cuh"
#include "kmx_launch.h"

namespace kmx {
int set_error(int code, const char* fmt, ...);   // kmx_host.cu
}
using namespace kmx;

namespace {

constexpr int kTextTile = 4096;

__global__ void newline_count_kernel(const uint8_t* __restrict__ text, uint64_t n, uint64_t* __restrict__ tile_cnt) {
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

// ---- exclusive scan of n 64-bit values (n + 1 outputs: out[n] = total): block sums, one-block scan of the sums, block scans ----
constexpr int kScanTile = 2048;

__device__ __forceinline__ unsigned long long block_excl_scan_u64(unsigned long long x, unsigned long long* s_warp /*[9]*/, unsigned long long* total) {
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned long long incl = x;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) {
		const unsigned long long y = __shfl_up_sync(0xffffffffu, incl, d);
		if (lane >= d) incl += y;
	}
	__syncthreads();
	if (lane == 31) s_warp[warp] = incl;
	__syncthreads();
	unsigned long long before = 0, all = 0;
#pragma unroll
	for (int w = 0; w < 8; w++) {
		const unsigned long long v = s_warp[w];
		before += w < warp ? v : 0;
		all += v;
	}
	*total = all;
	return before + incl - x;
}

__global__ void __launch_bounds__(256) scan_sums_kernel(const uint64_t* __restrict__ in, uint64_t n, uint64_t* __restrict__ sums) {
	__shared__ unsigned long long s_warp[9];
	const uint64_t base = (uint64_t)blockIdx.x * kScanTile;
	unsigned long long x = 0;
	for (int j = 0; j < kScanTile / 256; j++) {
		const uint64_t i = base + (uint64_t)j * 256 + threadIdx.x;
		x += i < n ? in[i] : 0;
	}
	unsigned long long total;
	block_excl_scan_u64(x, s_warp, &total);
	if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) scan_of_sums_kernel(uint64_t* __restrict__ sums, uint64_t n_tiles, uint64_t* __restrict__ grand_total) {
	__shared__ unsigned long long s_warp[9];
	unsigned long long carry = 0;
	for (uint64_t b = 0; b < n_tiles; b += 256) {
		const uint64_t i = b + threadIdx.x;
		const unsigned long long x = i < n_tiles ? sums[i] : 0;
		unsigned long long total;
		const unsigned long long ex = block_excl_scan_u64(x, s_warp, &total);
		if (i < n_tiles) sums[i] = carry + ex;
		carry += total;
		__syncthreads();
	}
	if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void __launch_bounds__(256) scan_write_kernel(const uint64_t* __restrict__ in, uint64_t n, const uint64_t* __restrict__ sums,
                                                         uint64_t* __restrict__ out) {
	__shared__ unsigned long long s_warp[9];
	const uint64_t base = (uint64_t)blockIdx.x * kScanTile;
	// thread t takes 8 consecutive elements
	unsigned long long v[kScanTile / 256], x = 0;
#pragma unroll
	for (int j = 0; j < kScanTile / 256; j++) {
		const uint64_t i = base + (uint64_t)threadIdx.x * (kScanTile / 256) + j;
		v[j] = i < n ? in[i] : 0;
		x += v[j];
	}
	unsigned long long total;
	unsigned long long ex = block_excl_scan_u64(x, s_warp, &total) + sums[blockIdx.x];
#pragma unroll
	for (int j = 0; j < kScanTile / 256; j++) {
		const uint64_t i = base + (uint64_t)threadIdx.x * (kScanTile / 256) + j;
		if (i < n) out[i] = ex;
		ex += v[j];
	}
}

// out[0 .. n) = exclusive prefix sums of in, out[n] = total (also copied to *total_h); d_sums needs n / 2048 + 2 words
cudaError_t exclusive_scan_u64(const uint64_t* d_in, uint64_t n, uint64_t* d_out, uint64_t* d_sums, uint64_t* total_h) {
	const uint64_t n_tiles = (n + kScanTile - 1) / kScanTile;
	if (n_tiles) {
		scan_sums_kernel<<<(unsigned)n_tiles, 256>>>(d_in, n, d_sums);
		scan_of_sums_kernel<<<1, 256>>>(d_sums, n_tiles, d_out + n);
		scan_write_kernel<<<(unsigned)n_tiles, 256>>>(d_in, n, d_sums, d_out);
		note_launch(3);
	} else {
		cudaError_t e = cudaMemset(d_out, 0, 8);
		if (e != cudaSuccess) return e;
	}
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return e;
	return cudaMemcpy(total_h, d_out + n, 8, cudaMemcpyDeviceToHost);
}

// ---- stream compaction in two passes around a scan: flags[i] (0 / 1 as u64) -> positions --------------------------------
// run heads of the sorted window list: head[i] = keys[i] != keys[i - 1]
__global__ void head_flag_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint64_t* __restrict__ flag) {
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}
// run j starts at head_at[j]; its length is the distance to the next head (head_at[n_runs] = n)
__global__ void head_write_kernel(const uint64_t* __restrict__ flag, const uint64_t* __restrict__ pos, uint64_t n, uint64_t n_runs,
                                  uint64_t* __restrict__ head_at) {
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		if (flag[i]) head_at[pos[i]] = i;
	if (blockIdx.x == 0 && threadIdx.x == 0) head_at[n_runs] = n;
}
// keep[j] = run j is a k-mer (not the invalid-window sentinel) seen at least ci times
__global__ void keep_flag_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ head_at, uint64_t n_runs, uint32_t ci,
                                 uint64_t* __restrict__ flag) {
	for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_runs; j += (uint64_t)gridDim.x * blockDim.x)
		flag[j] = (keys[head_at[j]] != kInvalid && head_at[j + 1] - head_at[j] >= (uint64_t)ci) ? 1 : 0;
}
__global__ void keep_write_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ head_at, const uint64_t* __restrict__ flag,
                                  const uint64_t* __restrict__ pos, uint64_t n_runs, uint64_t* __restrict__ kept, uint32_t* __restrict__ kept_cnt) {
	for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_runs; j += (uint64_t)gridDim.x * blockDim.x) {
		if (flag[j]) {
			const uint64_t len = head_at[j + 1] - head_at[j];
			kept[pos[j]] = keys[head_at[j]];
			kept_cnt[pos[j]] = len > 0xFFFFFFFFULL ? 0xFFFFFFFFu : (uint32_t)len;
		}
	}
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
	// ---- read the files: 4-line FASTQ, plain text or gzip (zlib reads both transparently) ----
	std::vector<uint8_t> text;
	for (int f = 0; f < n_files; f++) {
		gzFile gz = gzopen(fastq_paths[f], "rb");
		if (!gz) return set_error(KMX_EIO, "cannot open %s (%s)", fastq_paths[f], strerror(errno));
		gzbuffer(gz, 1u << 20);
		const size_t old = text.size();
		size_t at = old;
		for (;;) {
			if (text.size() < at + (8u << 20)) text.resize(at + (64u << 20));
			const int got = gzread(gz, text.data() + at, 8u << 20);
			if (got < 0) {
				int errnum = 0;
				const char* msg = gzerror(gz, &errnum);
				std::string m = msg ? msg : "read error";
				gzclose(gz);
				return set_error(KMX_EIO, "%s: %s", fastq_paths[f], m.c_str());
			}
			if (got == 0) break;
			at += (size_t)got;
		}
		gzclose(gz);
		text.resize(at + 1);
		if (at > old && text[at - 1] != '\n') text[at] = '\n';           // every file ends with a newline
		else text.resize(at);
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
	uint64_t* d_tile_cnt = nullptr;
	uint64_t* d_tile_off = nullptr;
	uint64_t* d_sums = nullptr;
	CUC(S.alloc(&d_tile_cnt, (tiles + 1) * 8));
	CUC(S.alloc(&d_tile_off, (tiles + 1) * 8));
	CUC(S.alloc(&d_sums, (tiles / kScanTile + 2) * 8));
	const int grid = sms * 8;
	uint64_t n_lines = 0;
	if (tiles) {
		newline_count_kernel<<<grid, 256>>>(d_text, n_bytes, d_tile_cnt);
		note_launch();
		CUC(exclusive_scan_u64(d_tile_cnt, tiles, d_tile_off, d_sums, &n_lines));
	}
	const uint64_t n_reads = n_lines / 4;
	uint64_t* d_line = nullptr;
	CUC(S.alloc(&d_line, (n_lines + 2) * 8));
	CUC(cudaMemset(d_line, 0, 8));
	if (tiles) {
		newline_scatter_kernel<<<grid, 256>>>(d_text, n_bytes, d_tile_off, d_line);
		note_launch();
	}
	// ---- windows ----
	uint64_t* d_win = nullptr;
	uint64_t* d_win_off = nullptr;
	uint64_t* d_sums2 = nullptr;
	CUC(S.alloc(&d_win, (n_reads + 1) * 8));
	CUC(S.alloc(&d_win_off, (n_reads + 1) * 8));
	CUC(S.alloc(&d_sums2, (n_reads / kScanTile + 2) * 8));
	uint64_t n_windows = 0;
	if (n_reads) {
		window_count_kernel<<<grid, 256>>>(d_text, d_line, n_reads, k, d_win);
		note_launch();
		CUC(exclusive_scan_u64(d_win, n_reads, d_win_off, d_sums2, &n_windows));
	}
	if (n_windows >= (1ULL << 32)) return set_error(KMX_ERANGE, "%llu k-mer windows: this single-pass counter holds every window on the device (< 2^32)", (unsigned long long)n_windows);
	uint64_t* d_keys = nullptr;
	uint64_t* d_sorted = nullptr;
	CUC(S.alloc(&d_keys, (n_windows + 1) * 8));
	CUC(S.alloc(&d_sorted, (n_windows + 1) * 8));
	if (n_windows) {
		extract_kernel<<<grid, 256>>>(d_text, d_line, d_win_off, n_reads, k, d_keys);
		note_launch();
	}
	// ---- sort; run heads; runs that are k-mers seen at least ci times (order kept) ----
	uint64_t n_unique = 0, n_kept = 0;
	uint64_t* d_kept = nullptr;
	uint32_t* d_kept_cnt = nullptr;
	if (n_windows) {
		void* d_tmp = nullptr;
		CUC(S.alloc(&d_tmp, radix_sort_temp_bytes(n_windows)));
		CUC(launch_radix_sort_pairs(d_tmp, d_keys, d_sorted, nullptr, nullptr, n_windows, 64, nullptr));
		uint64_t* d_flag = d_keys;                              // the unsorted windows are not needed any more
		uint64_t* d_pos = nullptr;
		uint64_t* d_sums3 = nullptr;
		CUC(S.alloc(&d_pos, (n_windows + 1) * 8));
		CUC(S.alloc(&d_sums3, (n_windows / kScanTile + 2) * 8));
		head_flag_kernel<<<grid, 256>>>(d_sorted, n_windows, d_flag);
		note_launch();
		CUC(exclusive_scan_u64(d_flag, n_windows, d_pos, d_sums3, &n_unique));
		uint64_t* d_head = nullptr;
		CUC(S.alloc(&d_head, (n_unique + 1) * 8));
		head_write_kernel<<<grid, 256>>>(d_flag, d_pos, n_windows, n_unique, d_head);
		keep_flag_kernel<<<grid, 256>>>(d_sorted, d_head, n_unique, (uint32_t)ci, d_flag);
		note_launch(2);
		CUC(exclusive_scan_u64(d_flag, n_unique, d_pos, d_sums3, &n_kept));
		CUC(S.alloc(&d_kept, (n_kept + 1) * 8));
		CUC(S.alloc(&d_kept_cnt, (n_kept + 1) * 4));
		keep_write_kernel<<<grid, 256>>>(d_sorted, d_head, d_flag, d_pos, n_unique, d_kept, d_kept_cnt);
		note_launch();
		CUC(cudaGetLastError());
		// n_unique counts k-mers: the sentinel run of the invalid windows is not one
		uint64_t last = 0;
		CUC(cudaMemcpy(&last, d_sorted + n_windows - 1, 8, cudaMemcpyDeviceToHost));
		if (last == kInvalid) n_unique--;
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
	note_launch(2);
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
