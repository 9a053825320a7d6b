// kmc_ra.hpp -- the random-access part of the reference's kmc_api (kmc_file.h: CKMCFile, kmer_api.h: CKmerAPI) on top of
// libkmx.so: same class and method names and argument meaning for OpenForRA / CheckKmer / IsKmer / GetCountersForRead and the
// small accessors around them (kmc_file.cpp:27-58,300-397,670-760,879-897), k <= 32.  The database lives in GPU memory; every
// call is one kernel launch, so the batch forms at the end (additive) are what a GPU user wants.  SURVEY.md 8f row N4.
//
//     CKMCFile db;                                   // kmc_file.h:32
//     db.OpenForRA("reads.res");                     // kmc_file.cpp:27
//     CKmerAPI kmer(db.KmerLength());  kmer.from_string("ACGT...");
//     uint32 c;  if (db.CheckKmer(kmer, c)) ...      // kmc_file.cpp:320
//     std::vector<uint32> counters;  db.GetCountersForRead(read, counters);   // kmc_file.cpp:879
#pragma once
#ifndef KMX_KMC_RA_HPP
#define KMX_KMC_RA_HPP

#include <stdint.h>
#include <string.h>
#include <string>
#include <vector>
#include "kmx.h"

typedef uint32_t uint32;
typedef uint64_t uint64;

// the k-mer holder CheckKmer takes (kmer_api.h): 2-bit packed, first base most significant, k <= 32
class CKmerAPI {
public:
	explicit CKmerAPI(uint32 length = 0) : k_(length), v_(0) {}
	// kmer_api.h:502-510: false (and no change) when a character is not one of ACGTacgt
	bool from_string(const std::string& s) {
		uint64_t v = 0;
		for (size_t i = 0; i < s.size(); i++) {
			const int c = code(s[i]);
			if (c < 0) return false;
			v = (v << 2) | (uint64_t)c;
		}
		if (s.size() > 32) return false;
		k_ = (uint32)s.size();
		v_ = v;
		return true;
	}
	std::string to_string() const {                   // kmer_api.h: to_string
		std::string s(k_, 'A');
		for (uint32 i = 0; i < k_; i++) s[i] = "ACGT"[(v_ >> (2 * (k_ - 1 - i))) & 3];
		return s;
	}
	bool reverse() {                                  // kmer_api.h:515: reverse complement in place
		uint64_t r = 0, t = v_;
		for (uint32 i = 0; i < k_; i++) {
			r = (r << 2) | (3 - (t & 3));
			t >>= 2;
		}
		v_ = r;
		return true;
	}
	bool operator<(const CKmerAPI& o) const { return v_ < o.v_; }
	bool operator==(const CKmerAPI& o) const { return k_ == o.k_ && v_ == o.v_; }
	uint32 get_signature(uint32 sig_len) const { return kmx_host_signature(v_, (int)k_, (int)sig_len); }   // kmer_api.h:653-673
	uint64_t packed() const { return v_; }
	uint32 length() const { return k_; }
	void from_packed(uint64_t v, uint32 k) {
		v_ = k >= 32 ? v : (v & ((1ULL << (2 * k)) - 1));
		k_ = k;
	}

private:
	static int code(char ch) {                        // CKmerAPI::num_codes, kmer_api.h:268-273
		switch (ch) {
			case 'A': case 'a': return 0;
			case 'C': case 'c': return 1;
			case 'G': case 'g': return 2;
			case 'T': case 't': return 3;
			default: return -1;
		}
	}
	uint32 k_;
	uint64_t v_;
};

class CKMCFile {
public:
	CKMCFile() : db_(nullptr) {}
	~CKMCFile() { Close(); }
	CKMCFile(const CKMCFile&) = delete;
	CKMCFile& operator=(const CKMCFile&) = delete;

	// kmc_file.cpp:27-58: header, LUT and the whole record area are loaded -- into GPU memory here
	bool OpenForRA(const std::string& file_name) {
		if (db_) return false;
		db_ = kmx_db_open(file_name.c_str());
		if (!db_) return false;
		if (kmx_db_upload(db_) != KMX_OK) {
			Close();
			return false;
		}
		return true;
	}
	bool Close() {                                    // kmc_file.cpp:617-640
		if (!db_) return false;
		kmx_db_close(db_);
		db_ = nullptr;
		return true;
	}

	// kmc_file.cpp:320-356 / 364-397: true when the k-mer is stored and its counter lies inside [min_count, max_count]
	bool CheckKmer(CKmerAPI& kmer, uint32& count) {
		if (!db_ || kmer.length() != KmerLength()) return false;
		const uint64_t v = kmer.packed();
		uint32_t c = 0;
		if (kmx_db_check_kmers(db_, &v, 1, &c) != KMX_OK || c == 0) return false;
		count = c;
		return true;
	}
	bool CheckKmer(CKmerAPI& kmer, uint64& count) {
		uint32 c = 0;
		if (!CheckKmer(kmer, c)) return false;
		count = c;
		return true;
	}
	bool CheckKmer(CKmerAPI& kmer, float& count) {     // kmc_file.cpp:300-312 (mode 0: the counter as a float)
		uint32 c = 0;
		if (!CheckKmer(kmer, c)) return false;
		count = (float)c;
		return true;
	}
	bool IsKmer(CKmerAPI& kmer) {                     // kmc_file.cpp:750-757
		uint32 c;
		return CheckKmer(kmer, c);
	}

	// kmc_file.cpp:879-897: one counter per k-mer of the read, 0 for k-mers holding a character other than ACGTacgt;
	// false (and no counters) when the read is shorter than k
	bool GetCountersForRead(const std::string& read, std::vector<uint32>& counters) {
		if (!db_) return false;
		const uint32 k = KmerLength();
		if (read.size() < k) {
			counters.clear();
			return false;
		}
		counters.assign(read.size() - k + 1, 0);
		const int64_t offsets[2] = { 0, (int64_t)read.size() };
		return kmx_db_counters_for_reads(db_, read.data(), offsets, 1, counters.data(), nullptr) == KMX_OK;
	}
	bool GetCountersForRead(const std::string& read, std::vector<float>& counters) {   // kmc_file.cpp:904-927
		std::vector<uint32> c;
		if (!GetCountersForRead(read, c)) return false;
		counters.assign(c.begin(), c.end());
		return true;
	}

	// kmc_file.cpp:670-734
	bool SetMinCount(uint32 x) { return db_ && kmx_db_set_count_range(db_, x, info().max_count) == KMX_OK; }
	bool SetMaxCount(uint32 x) { return db_ && kmx_db_set_count_range(db_, info().min_count, x) == KMX_OK; }
	void ResetMinMaxCounts() {
		if (db_) kmx_db_reset_count_range(db_);
	}
	uint32 GetMinCount() { return info().min_count; }
	uint64 GetMaxCount() { return info().max_count; }
	bool GetBothStrands() { return info().both_strands != 0; }
	uint64 KmerCount() { return info().total_kmers; }
	uint32 KmerLength() { return info().k; }
	// kmc_file.cpp:843-860
	bool Info(uint32& _kmer_length, uint32& _mode, uint32& _counter_size, uint32& _lut_prefix_length, uint32& _signature_len, uint32& _min_count,
	          uint64& _max_count, uint64& _total_kmers) {
		if (!db_) return false;
		const kmx_db_info_t i = info();
		_kmer_length = i.k;
		_mode = i.mode;
		_counter_size = i.counter_size;
		_lut_prefix_length = i.lut_prefix_length;
		_signature_len = i.signature_len;
		_min_count = i.min_count;
		_max_count = i.max_count;
		_total_kmers = i.total_kmers;
		return true;
	}

	// ---- additive: batches, one kernel launch each ----
	bool CheckKmers(const std::vector<uint64_t>& packed_kmers, std::vector<uint32>& counts) {
		if (!db_) return false;
		counts.assign(packed_kmers.size(), 0);
		return kmx_db_check_kmers(db_, packed_kmers.data(), (int64_t)packed_kmers.size(), counts.data()) == KMX_OK;
	}
	bool GetCountersForReads(const std::vector<std::string>& reads, std::vector<std::vector<uint32> >& counters) {
		if (!db_) return false;
		const uint32 k = KmerLength();
		std::string flat;
		std::vector<int64_t> offsets(reads.size() + 1, 0);
		size_t total = 0;
		for (size_t i = 0; i < reads.size(); i++) {
			flat += reads[i];
			offsets[i + 1] = (int64_t)flat.size();
			total += reads[i].size() >= k ? reads[i].size() - k + 1 : 0;
		}
		std::vector<uint32> all(total ? total : 1);
		if (kmx_db_counters_for_reads(db_, flat.data(), offsets.data(), (int64_t)reads.size(), all.data(), nullptr) != KMX_OK) return false;
		counters.assign(reads.size(), std::vector<uint32>());
		size_t at = 0;
		for (size_t i = 0; i < reads.size(); i++) {
			const size_t n = reads[i].size() >= k ? reads[i].size() - k + 1 : 0;
			counters[i].assign(all.begin() + at, all.begin() + at + n);
			at += n;
		}
		return true;
	}

	kmx_db* handle() { return db_; }

private:
	kmx_db_info_t info() {
		kmx_db_info_t i;
		memset(&i, 0, sizeof(i));
		if (db_) kmx_db_info(db_, &i);
		return i;
	}
	kmx_db* db_;
};

#endif
