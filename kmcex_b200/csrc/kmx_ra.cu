// kmx_ra.cu -- random access into the device-resident KMC database.
//
// Reference interfaces mirrored (file:line relative to the reference root):
//   kmc_file.cpp:27-58      CKMCFile::OpenForRA            (whole .kmc_suf + LUT + signature map in memory: here, in HBM)
//   kmc_file.cpp:320-356    CKMCFile::CheckKmer            (signature -> bin -> LUT slot -> binary search)
//   kmc_file.cpp:1358-1436  CKMCFile::BinarySearch         (suffix bytes compared most significant first; counter range check)
//   kmc_file.cpp:879-897,1130-1352  GetCountersForRead     (per window: 0 when it holds a non-ACGT byte, else the canonical /
//                                                           forward k-mer's counter from the bin of its signature)
//   kmer_api.h:653-673, mmer.h:33-88  signatures           (kmx_core.cuh: kmer_signature)
//
// kmcEx itself never calls these; they are the exact-count ground truth next to kmer_to_occ.  One thread per lookup: a
// lookup is ~25 m-mer normalisations in registers and a binary search of ~10 dependent 8-byte loads, i.e. bound by the
// latency of random sector reads like the query kernels.
#include <fcntl.h>
#include <unistd.h>
#include "kmx_internal.h"
#include "kmx_core.cuh"

using namespace kmx;
#define fail kmx::set_error

namespace {

struct RaDb {
	DevDb db;
	const uint32_t* sigmap;
	uint64_t n_bins, slots_per_bin;
	int sig_len, suffix_bases;
	bool both_strands;
};

__device__ __forceinline__ uint64_t ra_suffix(const RaDb& r, uint64_t rec) {
	const uint8_t* p = r.db.suf + rec * r.db.rec_bytes;
	uint64_t s = 0;
	for (uint32_t b = 0; b < r.db.suffix_bytes; b++) s = (s << 8) | __ldg(p + b);
	return s;
}

// CheckKmer for one packed k-mer: its counter, or 0
__device__ __forceinline__ uint32_t ra_lookup(const RaDb& r, uint64_t v) {
	const uint64_t prefix = r.suffix_bases >= 32 ? 0 : v >> (2 * r.suffix_bases);
	const uint64_t want = r.suffix_bases >= 32 ? v : (v & mask2(r.suffix_bases));
	const uint32_t bin = __ldg(r.sigmap + kmer_signature(v, r.db.k, r.sig_len));
	if (bin >= r.n_bins) return 0;                         // a bin the LUT does not have (the reference would index past it)
	const uint64_t slot = (uint64_t)bin * r.slots_per_bin + prefix;
	uint64_t lo = __ldg(r.db.lut + slot), hi = __ldg(r.db.lut + slot + 1);   // records [lo, hi) of this (bin, prefix)
	if (hi > r.db.total) hi = r.db.total;                  // the guard word is total + 1 (kmc_file.cpp:223)
	while (lo < hi) {
		const uint64_t mid = (lo + hi) >> 1;
		const uint64_t s = ra_suffix(r, mid);
		if (s == want) {
			const uint8_t* p = r.db.suf + mid * r.db.rec_bytes + r.db.suffix_bytes;
			uint32_t c = 0;
			for (uint32_t b = 0; b < r.db.counter_bytes && b < 4; b++) c |= (uint32_t)__ldg(p + b) << (8 * b);
			return (c >= r.db.min_count && c <= r.db.max_count) ? c : 0;      // kmc_file.cpp:1426-1434
		}
		if (s < want) lo = mid + 1;
		else hi = mid;
	}
	return 0;
}

__global__ void __launch_bounds__(256) ra_check_kernel(const __grid_constant__ RaDb r, const uint64_t* __restrict__ kmers, int64_t n,
                                                        uint32_t* __restrict__ counts) {
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
		counts[i] = ra_lookup(r, __ldg(kmers + i) & mask2(r.db.k));
}

// one thread per counter: counter c belongs to the read whose counter range [coff[q], coff[q + 1]) holds c
__global__ void __launch_bounds__(256) ra_reads_kernel(const __grid_constant__ RaDb r, const char* __restrict__ bases, const int64_t* __restrict__ off,
                                                        const int64_t* __restrict__ coff, int64_t n_reads, int64_t n_counters,
                                                        uint32_t* __restrict__ counters) {
	const int k = r.db.k;
	for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_counters; c += (int64_t)gridDim.x * blockDim.x) {
		int64_t lo = 0, hi = n_reads - 1;                  // last read with coff[read] <= c (reads without counters repeat an offset)
		while (lo < hi) {
			const int64_t mid = (lo + hi + 1) >> 1;
			if (__ldg(coff + mid) <= c) lo = mid;
			else hi = mid - 1;
		}
		const char* w = bases + __ldg(off + lo) + (c - __ldg(coff + lo));
		uint64_t v = 0;
		bool valid = true;
		for (int j = 0; j < k; j++) {
			const unsigned char ch = (unsigned char)__ldg(w + j) & 0xDFu;     // upper case: CKmerAPI::num_codes takes both (kmer_api.h:270-273)
			const uint32_t code = ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 4u;
			valid &= code < 4u;
			v = (v << 2) | (code & 3u);
		}
		uint32_t out = 0;
		if (valid) {
			if (r.both_strands) {
				const uint64_t rc = (~reverse_bases(v, k)) & mask2(k);
				v = v < rc ? v : rc;                       // kmc_file.cpp:1261-1264
			}
			out = ra_lookup(r, v);
		}
		counters[c] = out;
	}
}

// signature map -> device at the first lookup; the whole record area must be resident
int ra_prepare(kmx_db* db, RaDb* r, cudaStream_t s) {
	int rc = kmx_db_upload(db);
	if (rc) return rc;
	const kmx_db_info_t& h = db->info;
	if (db->rec_lo != 0 || db->rec_hi != h.total_kmers) return fail(KMX_ESTATE, "only a share of the database is on the device (team build); random access needs all of it");
	if (h.signature_len < 5 || h.signature_len > 11 || h.signature_len > h.k)
		return fail(KMX_EFORMAT, "signature length %u: KMC signatures are 5..11 bases (mmer.h:25-31)", h.signature_len);
	const uint64_t slots = 1ULL << (2 * h.lut_prefix_length);
	if (h.lut_entries % slots != 0) return fail(KMX_EFORMAT, "%llu LUT entries are not a whole number of bins of 4^%u slots", (unsigned long long)h.lut_entries, h.lut_prefix_length);
	CU(cudaSetDevice(db->device));
	std::lock_guard<std::mutex> lock(db->ra_mu);
	if (!db->d_sigmap) {
		const size_t n_sig = ((size_t)1 << (2 * h.signature_len)) + 1;
		std::vector<uint32_t> map(n_sig);
		const int fd = open(db->pre_name.c_str(), O_RDONLY);
		if (fd < 0) return fail(KMX_EIO, "can't reopen %s for its signature map", db->pre_name.c_str());
		size_t got = 0;
		while (got < n_sig * 4) {
			const ssize_t n = pread(fd, (uint8_t*)map.data() + got, n_sig * 4 - got, (off_t)(db->sig_offset + got));
			if (n <= 0) break;
			got += (size_t)n;
		}
		close(fd);
		if (got != n_sig * 4) return fail(KMX_EIO, "short read on the signature map of %s", db->pre_name.c_str());
		uint32_t* d = nullptr;
		DA(&d, n_sig * 4, s);
		CU(cudaMemcpyAsync(d, map.data(), n_sig * 4, cudaMemcpyHostToDevice, s));
		CU(cudaStreamSynchronize(s));                      // `map` is pageable and goes out of scope
		db->d_sigmap = d;
	}
	memset(r, 0, sizeof(*r));
	r->db = dev_db(db);
	r->sigmap = db->d_sigmap;
	r->slots_per_bin = slots;
	r->n_bins = h.lut_entries / slots;
	r->sig_len = (int)h.signature_len;
	r->suffix_bases = (int)(h.k - h.lut_prefix_length);
	r->both_strands = db->both_strands;
	return KMX_OK;
}

int ra_grid(int64_t n, int sm_count) {
	const int64_t blocks = (n + 255) / 256, cap = (int64_t)sm_count * 8;
	return (int)(blocks < 1 ? 1 : blocks < cap ? blocks : cap);
}

struct CtxLease {
	DevCtx* x = nullptr;
	~CtxLease() {
		if (x) ctx_release(x);
	}
};

}  // namespace

extern "C" uint32_t kmx_host_signature(uint64_t kmer, int k, int signature_len) {
	if (k < 1 || k > 32 || signature_len < 5 || signature_len > 11 || signature_len > k) return 0xFFFFFFFFu;
	return kmer_signature(kmer & mask2(k), k, signature_len);
}

// CKMCFile::SetMinCount / SetMaxCount / ResetMinMaxCounts (kmc_file.cpp:670-734): the range may only be narrowed within the
// header's; it applies to the listing filter and to the random-access range check alike, as in the reference
extern "C" int kmx_db_set_count_range(kmx_db* db, uint32_t min_count, uint32_t max_count) {
	if (!db) return fail(KMX_EARG, "null argument");
	if (min_count < db->orig_min_count || max_count > db->orig_max_count || min_count > max_count)
		return fail(KMX_ERANGE, "counter range [%u, %u] is not inside the database's [%u, %u]", min_count, max_count, db->orig_min_count, db->orig_max_count);
	db->info.min_count = min_count;
	db->info.max_count = max_count;
	return KMX_OK;
}

extern "C" int kmx_db_reset_count_range(kmx_db* db) {
	if (!db) return fail(KMX_EARG, "null argument");
	db->info.min_count = db->orig_min_count;
	db->info.max_count = db->orig_max_count;
	return KMX_OK;
}

extern "C" int kmx_db_check_kmers(kmx_db* db, const uint64_t* kmers, int64_t n, uint32_t* counts) {
	if (!db || n < 0 || (n > 0 && (!kmers || !counts))) return fail(KMX_EARG, "null argument");
	int sm = 0;
	int rc = require_gpu(&sm);
	if (rc) return rc;
	if (n == 0) return KMX_OK;
	CU(cudaSetDevice(db->device));
	CtxLease lease;
	if ((rc = ctx_acquire(&lease.x))) return rc;
	cudaStream_t s = lease.x->stream;
	RaDb r;
	if ((rc = ra_prepare(db, &r, s))) return rc;
	DevScope scope(s);
	uint64_t* d_k = nullptr;
	uint32_t* d_c = nullptr;
	if ((rc = scope.alloc(&d_k, (size_t)n * 8))) return rc;
	if ((rc = scope.alloc(&d_c, (size_t)n * 4))) return rc;
	CU(cudaMemcpyAsync(d_k, kmers, (size_t)n * 8, cudaMemcpyHostToDevice, s));
	ra_check_kernel<<<ra_grid(n, db->sm_count ? db->sm_count : sm), 256, 0, s>>>(r, d_k, n, d_c);
	note_launch();
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(counts, d_c, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));
	return KMX_OK;
}

extern "C" int kmx_db_counters_for_reads(kmx_db* db, const char* bases, const int64_t* offsets, int64_t n_reads, uint32_t* counters, int64_t* n_counters_out) {
	if (n_counters_out) *n_counters_out = 0;
	if (!db || n_reads < 0 || (n_reads > 0 && (!bases || !offsets))) return fail(KMX_EARG, "null argument");
	int sm = 0;
	int rc = require_gpu(&sm);
	if (rc) return rc;
	if (n_reads == 0) return KMX_OK;
	const int64_t k = (int64_t)db->info.k;
	std::vector<int64_t> coff((size_t)n_reads + 1);
	coff[0] = 0;
	for (int64_t q = 0; q < n_reads; q++) {
		const int64_t len = offsets[q + 1] - offsets[q];
		if (len < 0) return fail(KMX_EARG, "read offsets must not decrease (read %lld)", (long long)q);
		coff[(size_t)q + 1] = coff[(size_t)q] + (len >= k ? len - k + 1 : 0);      // kmc_file.cpp:884-888: a read shorter than k has no counters
	}
	const int64_t n_counters = coff[(size_t)n_reads], n_bases = offsets[n_reads] - offsets[0];
	if (n_counters_out) *n_counters_out = n_counters;
	if (n_counters == 0) return KMX_OK;
	if (!counters) return fail(KMX_EARG, "null argument");
	CU(cudaSetDevice(db->device));
	CtxLease lease;
	if ((rc = ctx_acquire(&lease.x))) return rc;
	cudaStream_t s = lease.x->stream;
	RaDb r;
	if ((rc = ra_prepare(db, &r, s))) return rc;
	DevScope scope(s);
	char* d_b = nullptr;
	int64_t* d_off = nullptr;
	int64_t* d_coff = nullptr;
	uint32_t* d_c = nullptr;
	if ((rc = scope.alloc(&d_b, (size_t)n_bases))) return rc;
	if ((rc = scope.alloc(&d_off, ((size_t)n_reads + 1) * 8))) return rc;
	if ((rc = scope.alloc(&d_coff, ((size_t)n_reads + 1) * 8))) return rc;
	if ((rc = scope.alloc(&d_c, (size_t)n_counters * 4))) return rc;
	std::vector<int64_t> rel((size_t)n_reads + 1);         // offsets relative to the first byte that is copied
	for (int64_t q = 0; q <= n_reads; q++) rel[(size_t)q] = offsets[q] - offsets[0];
	CU(cudaMemcpyAsync(d_b, bases + offsets[0], (size_t)n_bases, cudaMemcpyHostToDevice, s));
	CU(cudaMemcpyAsync(d_off, rel.data(), ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, s));
	CU(cudaMemcpyAsync(d_coff, coff.data(), ((size_t)n_reads + 1) * 8, cudaMemcpyHostToDevice, s));
	ra_reads_kernel<<<ra_grid(n_counters, db->sm_count ? db->sm_count : sm), 256, 0, s>>>(r, d_b, d_off, d_coff, n_reads, n_counters, d_c);
	note_launch();
	CU(cudaGetLastError());
	CU(cudaMemcpyAsync(counters, d_c, (size_t)n_counters * 4, cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));                          // also covers the pageable host vectors above
	return KMX_OK;
}
