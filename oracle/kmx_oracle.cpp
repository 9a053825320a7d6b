// oracle/kmx_oracle.cpp -- TEST INFRASTRUCTURE ONLY. NOT PRODUCT CODE.
//
// A sequential CPU restatement of the kmcEx model build + kmer_to_occ path (reference:
// lzhLab/kmcEx, files cited per function as file:line relative to the reference root).
// It exists so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg can
// check the CUDA path bit for bit. The product (kmcex_b200/libkmx.so) never links, loads
// or calls this file; it has no CPU fallback.
//
// Parity pin: the reference ships no tests or golden vectors, so this
// restatement is pinned against the UNMODIFIED reference compiled from /root/reference
// (oracle/_ref/ref_driver, built by oracle/Makefile): tests/test_oracle.py
// compares header/km.bin/rest.bin and query outputs byte for byte when the binary is
// present, and tests/golden/ holds digests + KATs generated from it by
// tests/golden/make_golden.py.
//
// The restatement works on 2-bit packed k-mers (first base in the most significant bits,
// A=0 C=1 G=2 T=3, the value CKmerAPI holds for k<=32) and expands them to ASCII only to
// hash, because the reference hashes the ASCII string.

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include <unordered_map>

namespace {

// tools.hpp:9 -- the 128 hash seeds (data, reproduced as the format constant it is).
const uint32_t kSeeds[128] = { 46757, 46769, 46771, 46807, 46811, 46817, 46819, 46829, 46831, 46853, 46861, 46867,
	46877, 46889, 46901, 46919, 46933, 46957, 46993, 46997, 47017, 47041, 47051, 47057, 47059, 47087, 47093, 47111,
	47119, 47123, 47129, 47137, 47143, 47147, 47149, 47161, 47189, 47207, 47221, 47237, 47251, 47269, 47279, 47287,
	47293, 47297, 47303, 47309, 47317, 47339, 47351, 47353, 47363, 47381, 47387, 47389, 47407, 47417, 47419, 47431,
	47441, 47459, 47491, 47497, 47501, 47507, 47513, 47521, 47527, 47533, 47543, 47563, 47569, 47581, 47591, 47599,
	47609, 47623, 47629, 47639, 47653, 47657, 47659, 47681, 47699, 47701, 47711, 47713, 47717, 47737, 47741, 47743,
	47777, 47779, 47791, 47797, 47807, 47809, 47819, 47837, 47843, 47857, 47869, 47881, 47903, 47911, 47917, 47933,
	47939, 47947, 47951, 47963, 47969, 47977, 47981, 48017, 48023, 48029, 48049, 48073, 48079, 48091, 48109, 48119,
	48121, 48131, 48157, 48163 };

const uint32_t kBucket = 1u << 18;  // kmodel.hpp:276 bucket_size (defines batch boundaries)

// tools.hpp:16-50 -- MurmurHash64A (Austin Appleby), little-endian 8-byte blocks, tail, avalanche.
uint64_t murmur64a(const uint8_t* p, int len, uint32_t seed) {
	const uint64_t M = 0xc6a4a7935bd1e995ULL;
	uint64_t h = (uint64_t)seed ^ ((uint64_t)(int64_t)len * M);
	int nblocks = len / 8;
	for (int b = 0; b < nblocks; b++) {
		uint64_t w;
		memcpy(&w, p + 8 * b, 8);
		w *= M; w ^= w >> 47; w *= M;
		h ^= w; h *= M;
	}
	int tail = len & 7;
	if (tail) {
		uint64_t t = 0;
		for (int i = tail - 1; i >= 0; i--) t = (t << 8) | p[8 * nblocks + i];
		h ^= t; h *= M;
	}
	h ^= h >> 47; h *= M; h ^= h >> 47;
	return h;
}

// tools.hpp:90-100 -- packed -> ASCII ("ACGT"[code], first base = top bits).
void to_ascii(uint64_t v, int len, uint8_t* out) {
	for (int i = len - 1; i >= 0; i--) { out[i] = "ACGT"[v & 3]; v >>= 2; }
}

// tools.hpp:130-139 -- reverse complement of a len-base packed value.
uint64_t revcomp(uint64_t v, int len) {
	uint64_t r = 0;
	for (int i = 0; i < len; i++) { r = (r << 2) | (3 - (v & 3)); v >>= 2; }
	return r;
}

// tools.hpp:160-167 -- canonical form = min(fwd, revcomp) by packed value.
uint64_t canonical(uint64_t v, int k) {
	uint64_t r = revcomp(v, k);
	return v <= r ? v : r;
}

// the (k-2)-mer the "back" filters hash: kmer.substr(1, k-2)   (kmodel.hpp:388, 475, 548)
uint64_t middle(uint64_t v, int k) {
	return (v >> 2) & ((k - 2 >= 32) ? ~0ULL : ((1ULL << (2 * (k - 2))) - 1));
}

uint64_t hash_packed(uint64_t v, int len, uint32_t seed) {
	uint8_t s[40];
	to_ascii(v, len, s);
	return murmur64a(s, len, seed);
}

// kmodel.hpp:576-588 -- bit `pos` lives in byte pos>>3, mask 0x80 >> (pos&7).
inline void set_bit(std::vector<uint8_t>& a, uint64_t pos) { a[pos >> 3] |= (uint8_t)(0x80u >> (pos & 7)); }
inline int get_bit(const std::vector<uint8_t>& a, uint64_t pos) { return (a[pos >> 3] >> (7 - (pos & 7))) & 1; }

// occu_bin.hpp:27-83 -- count quantiser; occ2bin has max_counter entries, bin2mean 1<<n_hash.
struct OccuBinO {
	int max_counter = 0, n_hash = 0, end1 = 0;
	std::vector<int> occ2bin, bin2mean;
	// returns false where the reference would write out of bounds (occu_bin.hpp:38-44)
	bool init(int max_counter_, int n_hash_) {
		max_counter = max_counter_; n_hash = n_hash_;
		int end3 = 1 << n_hash;
		end1 = end3 / 4;
		int end2 = end1 + end3 / 2;
		int zone2_bins = end3 / 2, zone2_cap = 3;
		if (end1 + zone2_bins * zone2_cap > max_counter) return false;
		std::vector<int> mean(max_counter, -1);
		occ2bin.assign(max_counter, -1);
		int start = end1;
		for (int i = 0; i < zone2_bins; i++) {            // occu_bin.hpp:38-44
			for (int j = 0; j < zone2_cap; j++) { mean[start + j] = start + 1; occ2bin[start + j] = end1 + i; }
			start += zone2_cap;
		}
		int zone3_bins = end3 / 4;
		int zone3_cap = (max_counter - start) / zone3_bins;   // occu_bin.hpp:47
		for (int i = 0; i < zone3_bins; i++) {            // occu_bin.hpp:48-54
			for (int j = 0; j < zone3_cap; j++) { mean[start + j] = (2 * start + zone3_cap) / 2; occ2bin[start + j] = end2 + i; }
			start += zone3_cap;
		}
		for (int i = start; i < max_counter; i++) {       // occu_bin.hpp:56-59
			mean[i] = (2 * start - zone3_cap) / 2; occ2bin[i] = end3 - 1;
		}
		for (int i = 0; i < end1 && i < max_counter; i++) occ2bin[i] = i;   // occu_bin.hpp:68-69
		bin2mean.assign(end3, 0);                         // missing bins read as 0 (unordered_map operator[])
		std::vector<char> seen(end3, 0);
		for (int i = 0; i < end1; i++) bin2mean[i] = i;   // occu_bin.hpp:80-81
		for (int i = end1; i < max_counter; i++) {        // occu_bin.hpp:61-63: insert() keeps the FIRST mean per bin
			int b = occ2bin[i];
			if (!seen[b]) { seen[b] = 1; bin2mean[b] = mean[i]; }
		}
		return true;
	}
};

// rest.hpp:78-83
int rest_prefix_len(int k) {
	for (int i = 7; i >= 3; i--) if ((k - i) % 4 == 0) return i;
	return 3;
}

struct RestO {
	int k = 0, pre_len = 0, map_size = 0, pre_buffer_size = 0, suff_group = 0;
	uint64_t suff_bin_size = 0, suffix_bin_count = 0;
	std::vector<int32_t> hash2index, pre_buffer, count_bin;
	std::vector<uint8_t> suffix_bin;

	// rest.hpp:95-135 (stat + sort_suffix + transform): net effect = all entries ordered by
	// (prefix, suffix bytes) = by packed value; pre_buffer is the cumulative count per
	// non-empty prefix starting at 0 (g++ evaluates rest.hpp:125 that way).
	void build(int k_, std::vector<std::pair<uint64_t, int32_t>>& items) {
		k = k_; pre_len = rest_prefix_len(k); map_size = 1 << (2 * pre_len);
		suff_group = (k - pre_len) / 4;
		std::sort(items.begin(), items.end());
		suffix_bin_count = items.size();
		suff_bin_size = suffix_bin_count * suff_group;
		hash2index.assign(map_size, -1);
		pre_buffer.clear(); pre_buffer.push_back(0);
		suffix_bin.resize(suff_bin_size);
		count_bin.resize(suffix_bin_count);
		int sbits = 2 * (k - pre_len);
		int groups = 0;
		for (size_t i = 0; i < items.size(); i++) {
			uint64_t v = items[i].first;
			uint32_t pre = (uint32_t)(v >> sbits);
			if (hash2index[pre] < 0) { hash2index[pre] = groups++; pre_buffer.push_back(pre_buffer.back()); }
			pre_buffer.back() += 1;
			for (int b = 0; b < suff_group; b++) suffix_bin[i * suff_group + b] = (uint8_t)(v >> (8 * (suff_group - 1 - b)));
			count_bin[i] = items[i].second;
		}
		pre_buffer_size = groups + 1;
	}

	// rest.hpp:197-221
	bool save(const std::string& path) const {
		FILE* f = fopen(path.c_str(), "wb");
		if (!f) return false;
		int32_t hdr[4] = { k, pre_len, map_size, pre_buffer_size };
		fwrite(hdr, 4, 4, f);
		fwrite(&suff_bin_size, 8, 1, f);
		fwrite(&suffix_bin_count, 8, 1, f);
		fwrite(hash2index.data(), 4, map_size, f);
		fwrite(pre_buffer.data(), 4, pre_buffer_size, f);
		fwrite(suffix_bin.data(), 1, suff_bin_size, f);
		fwrite(count_bin.data(), 4, suffix_bin_count, f);
		fclose(f);
		return true;
	}

	// rest.hpp:163-195
	bool load(const std::string& path) {
		FILE* f = fopen(path.c_str(), "rb");
		if (!f) return false;
		int32_t hdr[4];
		if (fread(hdr, 4, 4, f) != 4) { fclose(f); return false; }
		k = hdr[0]; pre_len = hdr[1]; map_size = hdr[2]; pre_buffer_size = hdr[3];
		bool ok = fread(&suff_bin_size, 8, 1, f) == 1 && fread(&suffix_bin_count, 8, 1, f) == 1;
		suff_group = (k - pre_len) / 4;
		hash2index.resize(map_size); pre_buffer.resize(pre_buffer_size);
		suffix_bin.resize(suff_bin_size); count_bin.resize(suffix_bin_count);
		ok = ok && fread(hash2index.data(), 4, map_size, f) == (size_t)map_size;
		ok = ok && fread(pre_buffer.data(), 4, pre_buffer_size, f) == (size_t)pre_buffer_size;
		ok = ok && fread(suffix_bin.data(), 1, suff_bin_size, f) == suff_bin_size;
		ok = ok && fread(count_bin.data(), 4, suffix_bin_count, f) == suffix_bin_count;
		fclose(f);
		return ok;
	}

	// rest.hpp:223-251 -- binary search with the INCLUSIVE upper bound pre_buffer[g+1]; the
	// probe at index suffix_bin_count (past the last group) is out of bounds in the
	// reference and is treated as "no match" here.
	int check(uint64_t v) const {
		int sbits = 2 * (k - pre_len);
		uint32_t pre = (uint32_t)(v >> sbits);
		int g = hash2index[pre];
		if (g < 0) return 0;
		uint8_t key[16];
		for (int b = 0; b < suff_group; b++) key[b] = (uint8_t)(v >> (8 * (suff_group - 1 - b)));
		int64_t low = pre_buffer[g], high = pre_buffer[g + 1];
		while (low <= high) {
			int64_t mid = (low + high) / 2;
			if ((uint64_t)mid >= suffix_bin_count) return 0;
			int c = memcmp(key, &suffix_bin[mid * suff_group], suff_group);
			if (c < 0) high = mid - 1;
			else if (c > 0) low = mid + 1;
			else return count_bin[mid];
		}
		return 0;
	}
};

struct ModelO {
	int ci = 1, cs = 1023, n_hash = 7, n_bits = 5, bf_num = 1, k = 0;
	int hb = 6, hk = 5;                                    // kmodel.hpp:52-54
	uint64_t total = 0, n_km = 0, kmer_counts[3] = { 0, 0, 0 };
	uint64_t byte_bf[3] = { 0 }, byte_bf_back[3] = { 0 }, byte_km_back = 0, km_byte_size = 0;
	std::vector<uint8_t> bf[3], bf_back[3], km_back;
	std::vector<std::vector<uint8_t>> val, tag;            // bit_array_1 / bit_array_2 per coupled pair
	std::vector<std::vector<uint32_t>> seeds;
	OccuBinO ob;
	RestO rest;

	bool configure(int ci_, int cs_, int nh, int nb) {     // kmodel.hpp:45-55, 674-677
		ci = ci_; cs = cs_; n_hash = nh; n_bits = nb;
		bf_num = ci == 1 ? 1 : 3;
		hb = nh - 1; hk = nh - 2;
		return ob.init(cs + 1, nh);
	}
	// kmodel.hpp:402-420 (the double expression at :411 is evaluated exactly as written)
	bool size_filters() {
		for (int i = 0; i < bf_num; i++) {
			byte_bf[i] = (uint64_t)(kmer_counts[i] / 5.5 * hb);
			byte_bf_back[i] = (kmer_counts[i] >> 3) * hk;
			if (byte_bf[i] == 0 || byte_bf_back[i] == 0) return false;   // reference: new uint8_t[0]{0} throws
			bf[i].assign(byte_bf[i], 0);
			bf_back[i].assign(byte_bf_back[i], 0);
		}
		return true;
	}
	// kmodel.hpp:436-456
	bool size_arrays() {
		km_byte_size = (n_km >> 4) * n_hash;
		byte_km_back = (n_km >> 4) * hk;
		if (km_byte_size == 0 || byte_km_back == 0) return false;
		km_back.assign(byte_km_back, 0);
		val.assign(n_bits, std::vector<uint8_t>(km_byte_size, 0));
		tag.assign(n_bits, std::vector<uint8_t>(km_byte_size, 0));
		seeds.assign(n_bits, std::vector<uint32_t>(n_hash));
		for (int i = 0; i < n_bits; i++) for (int j = 0; j < n_hash; j++) seeds[i][j] = kSeeds[(i * n_hash + j) % 128];
		return true;
	}

	// kmodel.hpp:498-506 insert_bloomfilter / 373-383 check_bloomfilter on a packed string
	void filter_insert(std::vector<uint8_t>& f, uint64_t v, int len, int nh) {
		uint64_t bits = (uint64_t)f.size() * 8;
		for (int j = 0; j < nh; j++) set_bit(f, hash_packed(v, len, kSeeds[j]) % bits);
	}
	bool filter_check(const std::vector<uint8_t>& f, uint64_t v, int len, int nh) const {
		uint64_t bits = (uint64_t)f.size() * 8;
		for (int j = 0; j < nh; j++) if (!get_bit(f, hash_packed(v, len, kSeeds[j]) % bits)) return false;
		return true;
	}
	// kmodel.hpp:590-622 -- check against the PRE-item state only, then set.
	bool array_insert(uint64_t v, uint32_t bin, int a) {
		uint64_t bits = km_byte_size * 8;
		uint64_t pos[64]; int want[64];
		for (int j = 0; j < n_hash; j++) { want[j] = (bin >> j) & 1; pos[j] = hash_packed(v, k, seeds[a][j]) % bits; }
		for (int j = 0; j < n_hash; j++) if (get_bit(tag[a], pos[j]) && get_bit(val[a], pos[j]) != want[j]) return false;
		for (int j = 0; j < n_hash; j++) { if (want[j]) set_bit(val[a], pos[j]); set_bit(tag[a], pos[j]); }
		return true;
	}
	// kmodel.hpp:361-371
	int check_all_bf(uint64_t v) const {
		static const int order3[3] = { 1, 0, 2 };          // kmodel.hpp:246
		for (int j = 0; j < bf_num; j++) {
			int i = ci == 1 ? j : order3[j];
			bool a = filter_check(bf[i], v, k, hb);
			bool b = filter_check(bf_back[i], middle(v, k), k - 2, hk);
			if (a && b) return i + ci;
		}
		return 0;
	}
	// kmodel.hpp:625-646 (all arrays, bins > 0) and 650-671 (first non-zero; -1 if no array matched)
	int decode_array(uint64_t v, int a, bool* all_tags) const {
		uint64_t bits = km_byte_size * 8;
		int bin = 0; bool ok = true;
		for (int j = 0; j < n_hash; j++) {
			uint64_t p = hash_packed(v, k, seeds[a][j]) % bits;
			bin |= get_bit(val[a], p) << j;
			if (!get_bit(tag[a], p)) ok = false;
		}
		*all_tags = ok;
		return bin;
	}
	std::vector<int> find_bitarray(uint64_t v) const {
		std::vector<int> out;
		for (int a = 0; a < n_bits; a++) { bool ok; int bin = decode_array(v, a, &ok); if (ok && bin > 0) out.push_back(bin); }
		return out;
	}
	int find_bitarray_one(uint64_t v) const {
		int result = -1;
		for (int a = 0; a < n_bits; a++) { bool ok; int bin = decode_array(v, a, &ok); if (ok) { result = bin; if (bin != 0) break; } }
		return result;
	}
	// kmodel.hpp:326-342
	void candidates_of(uint64_t nb, std::vector<int>& c) const {
		uint64_t v = canonical(nb, k);
		int r = rest.check(v);
		if (r > 0) { c.push_back(ob.occ2bin[r]); return; }
		int occ = check_all_bf(v);
		if (occ != 0) { c.push_back(occ); return; }
		if (filter_check(km_back, middle(v, k), k - 2, hk)) { int b = find_bitarray_one(v); if (b > -1) c.push_back(b); }
	}
	// kmodel.hpp:344-359 -- 4 successors (drop first base, append A,C,G,T) then 4 predecessors
	std::vector<int> neighbour_bins(uint64_t v) const {
		std::vector<int> c;
		uint64_t mask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1);
		for (uint64_t b = 0; b < 4; b++) candidates_of(((v << 2) & mask) | b, c);
		for (uint64_t b = 0; b < 4; b++) candidates_of((v >> 2) | (b << (2 * (k - 1))), c);
		return c;
	}
	// kmodel.hpp:286-323
	int kmer_to_bin(uint64_t v, int occ) const {
		std::vector<int> bins = find_bitarray(v);
		if (bins.empty()) return occ;
		if (bins.size() == 1) {
			if (occ) {
				std::vector<int> c = neighbour_bins(v);
				size_t low = 0;
				for (int x : c) if (x < ci + bf_num) low++;
				if (low >= c.size() / 2) return occ;
			}
			return bins[0];
		}
		std::vector<int> c = neighbour_bins(v);
		if (c.empty()) return 0;
		int best = bins[0], best_d = 2 << 20;
		for (int b : bins) {
			int d = 2 << 20;
			for (int x : c) d = std::min(d, abs(b - x));
			if (best_d > d) { best_d = d; best = b; }
		}
		return best;
	}
	// kmodel.hpp:100-116
	int kmer_to_occ(uint64_t raw) const {
		uint64_t v = canonical(raw, k);
		int occ = rest.check(v);
		if (occ != 0) return occ;
		bool in_back = filter_check(km_back, middle(v, k), k - 2, hk);
		occ = check_all_bf(v);
		if (occ != 0 && !in_back) return occ;
		if (!in_back) return 0;
		int bin = kmer_to_bin(v, occ);
		return ob.bin2mean[bin];
	}

	// ---- the same retrieval on STRINGS, as the reference runs it (kmodel.hpp:100-116): a query is never validated.
	// kmers2uint64 (tools.hpp:63-76) reads every byte that is not C/G/T as A; get_min_kmer (tools.hpp:160-167) returns the
	// ORIGINAL string when that encoding is <= its reverse complement, else the decoded reverse complement; the rest lookup
	// re-encodes what it is given (rest.hpp:22-34,223-251), the filters hash its raw bytes (kmodel.hpp:373-390,625-671).
	static uint64_t pack_str(const std::string& s) {
		uint64_t v = 0;
		for (char ch : s) { v <<= 2; v |= ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : 0; }
		return v;
	}
	static std::string min_kmer_str(const std::string& s) {
		int len = (int)s.size();
		uint64_t u = pack_str(s), rc = revcomp(u, len);
		if (u <= rc) return s;
		std::string out(len, 'A');
		to_ascii(rc, len, (uint8_t*)&out[0]);
		return out;
	}
	bool filter_check_str(const std::vector<uint8_t>& f, const std::string& s, int nh) const {
		uint64_t bits = (uint64_t)f.size() * 8;
		for (int j = 0; j < nh; j++) if (!get_bit(f, murmur64a((const uint8_t*)s.data(), (int)s.size(), kSeeds[j]) % bits)) return false;
		return true;
	}
	int check_all_bf_str(const std::string& s) const {
		static const int order3[3] = { 1, 0, 2 };
		for (int j = 0; j < bf_num; j++) {
			int i = ci == 1 ? j : order3[j];
			bool a = filter_check_str(bf[i], s, hb);
			bool b = filter_check_str(bf_back[i], s.substr(1, s.size() - 2), hk);
			if (a && b) return i + ci;
		}
		return 0;
	}
	int decode_array_str(const std::string& s, int a, bool* all_tags) const {
		uint64_t bits = km_byte_size * 8;
		int bin = 0; bool ok = true;
		for (int j = 0; j < n_hash; j++) {
			uint64_t p = murmur64a((const uint8_t*)s.data(), (int)s.size(), seeds[a][j]) % bits;
			bin |= get_bit(val[a], p) << j;
			if (!get_bit(tag[a], p)) ok = false;
		}
		*all_tags = ok;
		return bin;
	}
	void candidates_of_str(const std::string& nb, std::vector<int>& c) const {
		std::string s = min_kmer_str(nb);
		int r = rest.check(pack_str(s));
		if (r > 0) { c.push_back(ob.occ2bin[r]); return; }
		int occ = check_all_bf_str(s);
		if (occ != 0) { c.push_back(occ); return; }
		if (filter_check_str(km_back, s.substr(1, s.size() - 2), hk)) {
			int result = -1;
			for (int a = 0; a < n_bits; a++) { bool ok; int bin = decode_array_str(s, a, &ok); if (ok) { result = bin; if (bin != 0) break; } }
			if (result > -1) c.push_back(result);
		}
	}
	std::vector<int> neighbour_bins_str(const std::string& s) const {
		std::vector<int> c;
		const char* base = "ACGT";
		for (int b = 0; b < 4; b++) candidates_of_str(s.substr(1) + base[b], c);
		for (int b = 0; b < 4; b++) candidates_of_str(base[b] + s.substr(0, s.size() - 1), c);
		return c;
	}
	int kmer_to_occ_str(const std::string& query) const {
		if ((int)query.size() != k) return 0;               // (the reference probes the filters with the odd length; out of scope)
		std::string s = min_kmer_str(query);
		int occ = rest.check(pack_str(s));
		if (occ != 0) return occ;
		bool in_back = filter_check_str(km_back, s.substr(1, s.size() - 2), hk);
		occ = check_all_bf_str(s);
		if (occ != 0 && !in_back) return occ;
		if (!in_back) return 0;
		std::vector<int> bins;
		for (int a = 0; a < n_bits; a++) { bool ok; int bin = decode_array_str(s, a, &ok); if (ok && bin > 0) bins.push_back(bin); }
		int bin;
		if (bins.empty()) bin = occ;
		else if (bins.size() == 1) {
			bin = bins[0];
			if (occ) {
				std::vector<int> c = neighbour_bins_str(s);
				size_t low = 0;
				for (int x : c) if (x < ci + bf_num) low++;
				if (low >= c.size() / 2) bin = occ;
			}
		} else {
			std::vector<int> c = neighbour_bins_str(s);
			if (c.empty()) bin = 0;
			else {
				int best = bins[0], best_d = 2 << 20;
				for (int b : bins) {
					int d = 2 << 20;
					for (int x : c) d = std::min(d, abs(b - x));
					if (best_d > d) { best_d = d; best = b; }
				}
				bin = best;
			}
		}
		return ob.bin2mean[bin];
	}

	// kmodel.hpp:173-206
	bool save(const std::string& dir) const {
		FILE* h = fopen((dir + "/header").c_str(), "w");
		if (!h) return false;
		fprintf(h, "number_hash %d\nnumber_bit %d\nci %d\ncs %d\n", n_hash, n_bits, ci, cs);
		fclose(h);
		FILE* f = fopen((dir + "/km.bin").c_str(), "wb");
		if (!f) return false;
		fwrite(&n_km, 8, 1, f);
		for (int i = 0; i < bf_num; i++) fwrite(&kmer_counts[i], 8, 1, f);
		for (int i = 0; i < bf_num; i++) { fwrite(bf[i].data(), 1, byte_bf[i], f); fwrite(bf_back[i].data(), 1, byte_bf_back[i], f); }
		fwrite(km_back.data(), 1, byte_km_back, f);
		for (int i = 0; i < n_bits; i++) { fwrite(val[i].data(), 1, km_byte_size, f); fwrite(tag[i].data(), 1, km_byte_size, f); }
		fclose(f);
		return rest.save(dir + "/rest.bin");
	}
	// kmodel.hpp:680-696 + 209-235
	bool load(const std::string& dir) {
		FILE* h = fopen((dir + "/header").c_str(), "r");
		if (!h) return false;
		char key[64]; int nh, nb, ci_, cs_;
		bool ok = fscanf(h, "%63s %d", key, &nh) == 2 && fscanf(h, "%63s %d", key, &nb) == 2 &&
			fscanf(h, "%63s %d", key, &ci_) == 2 && fscanf(h, "%63s %d", key, &cs_) == 2;
		fclose(h);
		if (!ok || !configure(ci_, cs_, nh, nb)) return false;
		FILE* f = fopen((dir + "/km.bin").c_str(), "rb");
		if (!f) return false;
		ok = fread(&n_km, 8, 1, f) == 1;
		for (int i = 0; i < bf_num; i++) ok = ok && fread(&kmer_counts[i], 8, 1, f) == 1;
		ok = ok && size_filters();
		for (int i = 0; ok && i < bf_num; i++) {
			ok = fread(bf[i].data(), 1, byte_bf[i], f) == byte_bf[i] && fread(bf_back[i].data(), 1, byte_bf_back[i], f) == byte_bf_back[i];
		}
		ok = ok && size_arrays();
		ok = ok && fread(km_back.data(), 1, byte_km_back, f) == byte_km_back;
		for (int i = 0; ok && i < n_bits; i++) {
			ok = fread(val[i].data(), 1, km_byte_size, f) == km_byte_size && fread(tag[i].data(), 1, km_byte_size, f) == km_byte_size;
		}
		fclose(f);
		if (!ok || !rest.load(dir + "/rest.bin")) return false;
		k = rest.k;
		return true;
	}
};

// ---- KMC database listing (kmc_file.cpp:66-99, 132-171, 177-235, 428-515; KMC2/3 layout, version 0x200) ----
struct KmcDbO {
	uint32_t k = 0, mode = 0, counter_size = 0, lut_prefix_length = 0, signature_len = 0, min_count = 0, max_count = 0;
	uint64_t total = 0;
	std::vector<uint64_t> lut;      // n entries + guard
	std::vector<uint8_t> suf;       // record bytes (markers stripped)
	std::vector<uint32_t> signature_map;   // 4^signature_len + 1 bin numbers (kmc_file.cpp:211,224-226)
	bool both_strands = true;       // kmc_file.cpp:208-209
	uint32_t sufix_size = 0, rec_size = 0;

	static bool slurp(const std::string& path, const char* marker, std::vector<uint8_t>& out) {
		FILE* f = fopen(path.c_str(), "rb");
		if (!f) return false;
		fseek(f, 0, SEEK_END);
		long sz = ftell(f);
		fseek(f, 0, SEEK_SET);
		out.resize(sz);
		bool ok = sz >= 8 && fread(out.data(), 1, sz, f) == (size_t)sz;
		fclose(f);
		return ok && memcmp(out.data(), marker, 4) == 0 && memcmp(out.data() + sz - 4, marker, 4) == 0;
	}
	bool open(const std::string& base) {
		std::vector<uint8_t> pre;
		if (!slurp(base + ".kmc_pre", "KMCP", pre)) return false;
		size_t sz = pre.size();
		uint32_t version;
		memcpy(&version, &pre[sz - 12], 4);                      // kmc_file.cpp:180-184
		if (version != 0x200) return false;
		uint32_t header_offset = pre[sz - 8];                    // kmc_file.cpp:190-193 (one byte)
		const uint8_t* h = &pre[sz - 8 - header_offset];         // kmc_file.cpp:197
		memcpy(&k, h, 4); memcpy(&mode, h + 4, 4); memcpy(&counter_size, h + 8, 4);
		memcpy(&lut_prefix_length, h + 12, 4); memcpy(&signature_len, h + 16, 4);
		memcpy(&min_count, h + 20, 4); memcpy(&max_count, h + 24, 4); memcpy(&total, h + 28, 8);
		uint64_t body = sz - 12;                                  // without 2 markers and the header_offset word
		uint64_t sig_bytes = ((1ULL << (2 * signature_len)) + 1) * 4;
		uint64_t lut_bytes = body - (sig_bytes + header_offset + 8);   // kmc_file.cpp:212
		uint64_t n = lut_bytes / 8;
		lut.resize(n + 1);
		memcpy(lut.data(), &pre[4], (n + 1) * 8);
		lut[n] = total + 1;                                       // kmc_file.cpp:223
		both_strands = h[36] == 0;                                // kmc_file.cpp:208-209 (the byte means "one strand only")
		signature_map.resize(sig_bytes / 4);
		memcpy(signature_map.data(), &pre[4 + (n + 1) * 8], sig_bytes);   // kmc_file.cpp:224-226
		sufix_size = (k - lut_prefix_length) / 4;                 // kmc_file.cpp:230-232
		rec_size = sufix_size + counter_size;
		std::vector<uint8_t> raw;
		if (!slurp(base + ".kmc_suf", "KMCS", raw)) return false;
		suf.assign(raw.begin() + 4, raw.end() - 4);
		return k >= 1 && k <= 32 && mode == 0 && (uint64_t)rec_size * total <= suf.size();
	}
	// kmc_file.cpp:428-515 -- records in file order; LUT slot advances while record index == lut[slot+1]
	template <class F> void for_each(F fn) const {
		uint64_t slot = 0, mask = (1ULL << (2 * lut_prefix_length)) - 1;
		for (uint64_t s = 0; s < total; s++) {
			if (s == lut[slot + 1]) { slot++; while (lut[slot] == lut[slot + 1]) slot++; }
			const uint8_t* r = &suf[s * rec_size];
			uint64_t v = slot & mask;
			for (uint32_t b = 0; b < sufix_size; b++) v = (v << 8) | r[b];
			uint32_t c = 0;
			for (uint32_t b = 0; b < counter_size && b < 4; b++) c |= (uint32_t)r[sufix_size + b] << (8 * b);
			if (c < min_count || c > max_count) continue;
			fn(v, c);
		}
	}
};

// ---- random access: CKMCFile::CheckKmer / GetCountersForRead ----------------------------------
// mmer.h:33-58 -- which m-mers may be signatures (mmer packed like a k-mer, first base most significant)
bool mmer_is_allowed(uint32_t mmer, uint32_t len) {
	if ((mmer & 0x3f) == 0x3f) return false;                  // ...TTT
	if ((mmer & 0x3f) == 0x3b) return false;                  // ...TGT
	if ((mmer & 0x3c) == 0x3c) return false;                  // ...TT?
	for (uint32_t j = 0; j + 3 < len; j++) {                  // AA anywhere behind the third base
		if ((mmer & 0xf) == 0) return false;
		mmer >>= 2;
	}
	if (mmer == 0) return false;                              // AAA...
	if (mmer == 0x04) return false;                           // ACA...
	if ((mmer & 0xf) == 0) return false;                      // ?AA...
	return true;
}

// mmer.h:63-88 -- norm[]: the smaller of an m-mer and its reverse complement among the allowed ones, 4^len if neither is
uint32_t mmer_norm(uint32_t mmer, uint32_t len) {
	const uint32_t special = 1u << (2 * len);
	const uint32_t rev = (uint32_t)revcomp(mmer, (int)len);
	const uint32_t a = mmer_is_allowed(mmer, len) ? mmer : special, b = mmer_is_allowed(rev, len) ? rev : special;
	return std::min(a, b);
}

// kmer_api.h:653-673 -- CKmerAPI::get_signature: the minimum norm over the k-mer's m-mers
uint32_t kmer_signature(uint64_t v, int k, int len) {
	uint32_t best = 0xFFFFFFFFu;
	for (int i = 0; i + len <= k; i++) {
		const uint32_t mmer = (uint32_t)(v >> (2 * (k - len - i))) & ((1u << (2 * len)) - 1);
		best = std::min(best, mmer_norm(mmer, (uint32_t)len));
	}
	return best;
}

// kmc_file.cpp:320-356 (CheckKmer) + :1358-1436 (BinarySearch, inclusive bounds, suffix bytes compared most significant
// first, counter range check).  The guard word makes the last slot end one record past the file; that probe reads beyond
// the buffer in the reference and is treated as "no record" here.
uint32_t db_check_kmer(const KmcDbO& db, uint64_t v) {
	const uint32_t suffix_bases = db.k - db.lut_prefix_length;
	const uint64_t prefix = suffix_bases >= 32 ? 0 : v >> (2 * suffix_bases);
	const uint64_t single_lut = 1ULL << (2 * db.lut_prefix_length);
	if (db.signature_len < 5 || db.signature_len > 11) return 0;
	const uint64_t bin_start = (uint64_t)db.signature_map[kmer_signature(v, (int)db.k, (int)db.signature_len)] * single_lut;
	if (bin_start + prefix + 1 >= db.lut.size()) return 0;
	int64_t index_start = (int64_t)db.lut[bin_start + prefix], index_stop = (int64_t)db.lut[bin_start + prefix + 1] - 1;
	if ((uint64_t)index_start >= db.total) return 0;
	while (index_start <= index_stop) {
		const int64_t mid = (index_start + index_stop) / 2;
		if ((uint64_t)mid >= db.total) { index_stop = mid - 1; continue; }
		const uint8_t* rec = &db.suf[(uint64_t)mid * db.rec_size];
		int cmp = 0;                                              // sign of (record suffix - pattern suffix)
		for (uint32_t a = 0; a < db.sufix_size && cmp == 0; a++) {
			const uint32_t pattern = (uint32_t)(v >> (8 * (db.sufix_size - 1 - a))) & 0xff;
			cmp = (int)rec[a] - (int)pattern;
		}
		if (cmp == 0) {
			uint32_t c = 0;
			for (uint32_t b = 0; b < db.counter_size && b < 4; b++) c |= (uint32_t)rec[db.sufix_size + b] << (8 * b);
			return (c >= db.min_count && c <= db.max_count) ? c : 0;
		}
		if (cmp < 0) index_start = mid + 1;
		else index_stop = mid - 1;
	}
	return 0;
}

// kmc_file.cpp:879-897, 1130-1352 -- GetCountersForRead (KMC2 layout).  The reference walks super-k-mers (maximal runs of
// windows that share a signature); per window that is: a byte outside ACGTacgt -> 0, else the counter of the window's
// k-mer (the smaller of it and its reverse complement when the database holds both strands) in the bin of its signature.
int64_t db_counters_for_read(const KmcDbO& db, const char* read, int64_t len, uint32_t* counters) {
	if (len < (int64_t)db.k) return 0;                            // kmc_file.cpp:884-888
	const int k = (int)db.k;
	for (int64_t i = 0; i + k <= len; i++) {
		uint64_t v = 0;
		bool valid = true;
		for (int j = 0; j < k; j++) {
			int code;
			switch (read[i + j]) {                                // CKmerAPI::num_codes, kmer_api.h:268-273
				case 'A': case 'a': code = 0; break;
				case 'C': case 'c': code = 1; break;
				case 'G': case 'g': code = 2; break;
				case 'T': case 't': code = 3; break;
				default: code = -1;
			}
			if (code < 0) { valid = false; break; }
			v = (v << 2) | (uint64_t)code;
		}
		if (!valid) { counters[i] = 0; continue; }
		if (db.both_strands) {
			const uint64_t rc = revcomp(v, k);
			if (!(v < rc)) v = rc;                                // kmc_file.cpp:1261-1264
		}
		counters[i] = db_check_kmer(db, v);
	}
	return len - k + 1;
}

struct Item { uint64_t kmer; uint32_t occ; };

// kmodel.hpp:529-540, literally (two-pointer compaction; returns the new length)
int reorder_literal(Item* a, int n) {
	int il = 0, ir = n - 1;
	while (il < ir) {
		while (il < ir && !a[ir].occ) ir--;
		while (il < ir && a[il].occ) il++;
		if (il < ir) { a[il] = a[ir]; a[ir].occ = 0; }
	}
	return a[il].occ ? il + 1 : 0;
}

// kmodel.hpp:57-86 with its helpers 423-434, 479-496, 508-527, 543-573
int build_model(ModelO& m, const KmcDbO& db, int64_t* stats) {
	m.k = db.k;
	m.total = db.total;
	for (int i = 0; i < 3; i++) m.kmer_counts[i] = 0;
	uint32_t lim = m.ci + m.bf_num;
	bool below_ci = false;
	db.for_each([&](uint64_t, uint32_t c) {
		if (c < (uint32_t)m.ci || c > (uint32_t)m.cs) below_ci = true;
		else if (c < lim) m.kmer_counts[c - m.ci]++;
	});
	if (below_ci) return 3;           // reference indexes kmer_counts[c-ci] / occ_bin_meta[c] out of bounds
	if (!m.size_filters()) return 4;
	uint64_t n_bf = 0;
	for (int i = 0; i < m.bf_num; i++) n_bf += m.kmer_counts[i];
	m.n_km = m.total - n_bf;
	if (!m.size_arrays()) return 5;

	const int B = m.n_bits;
	std::vector<std::vector<Item>> buf(B, std::vector<Item>(kBucket, Item{ 0, 0 }));   // zero pages, as a fresh new[] is in practice
	std::vector<int> buf_n(B, (int)kBucket);
	std::vector<std::pair<uint64_t, int32_t>> rest_items;
	uint32_t fill = 0;
	int64_t attempts = 0, accepted = 0;

	auto run_batch = [&]() {          // kmodel.hpp:557-573
		for (int t = 0; t < B; t++) {
			for (int i = 0; i < B; i++) {     // buckets are independent within a round (distinct arrays)
				int a = (i + t) % B;
				Item* items = buf[i].data();
				for (int c = 0; c < buf_n[i]; c++) {       // kmodel.hpp:543-555
					attempts++;
					uint32_t bin = (uint32_t)m.ob.occ2bin[items[c].occ];
					if (m.array_insert(items[c].kmer, bin, a)) {
						accepted++;
						m.filter_insert(m.km_back, middle(items[c].kmer, m.k), m.k - 2, m.hk);
						items[c].occ = 0;
					}
				}
				buf_n[i] = reorder_literal(items, buf_n[i]);
			}
		}
		for (int i = 0; i < B; i++) {
			for (int j = 0; j < buf_n[i]; j++) rest_items.push_back({ buf[i][j].kmer, (int32_t)buf[i][j].occ });
			buf_n[i] = kBucket;
		}
	};

	db.for_each([&](uint64_t v, uint32_t c) {
		if (c < lim) {                // kmodel.hpp:473-477
			m.filter_insert(m.bf[c - m.ci], v, m.k, m.hb);
			m.filter_insert(m.bf_back[c - m.ci], middle(v, m.k), m.k - 2, m.hk);
		} else {                      // kmodel.hpp:508-518
			buf[fill / kBucket][fill % kBucket] = Item{ v, c };
			if (++fill >= kBucket * (uint32_t)B) { run_batch(); fill = 0; }
		}
	});
	if (fill > 0) {                   // kmodel.hpp:520-527 (fill == 0 is undefined behaviour in the reference; skipped)
		int row = (fill - 1) / kBucket, col = (fill - 1) % kBucket;
		buf_n[row] = col + 1;
		for (int i = row + 1; i < B; i++) buf_n[i] = 0;   // stale slot 0 of these buckets can resurface via reorder_literal(a, 0)
		run_batch();
	}
	m.rest.build(m.k, rest_items);
	if (stats) { stats[0] = attempts; stats[1] = accepted; stats[2] = (int64_t)rest_items.size(); }
	return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// C interface for ctypes (tests / smoke / bench cpu_baseline only)
// ---------------------------------------------------------------------------------------------
extern "C" {

uint64_t kmxo_murmur64(const void* key, int len, uint32_t seed) { return murmur64a((const uint8_t*)key, len, seed); }
uint32_t kmxo_seed(int i) { return kSeeds[i & 127]; }
uint64_t kmxo_canonical(uint64_t v, int k) { return canonical(v, k); }
uint64_t kmxo_hash_packed(uint64_t v, int len, uint32_t seed) { return hash_packed(v, len, seed); }

// fills occ2bin[max_counter], bin2mean[1<<n_hash]; returns 0 on success
int kmxo_occubin(int max_counter, int n_hash, int32_t* occ2bin, int32_t* bin2mean) {
	OccuBinO ob;
	if (!ob.init(max_counter, n_hash)) return 1;
	for (int i = 0; i < max_counter; i++) occ2bin[i] = ob.occ2bin[i];
	for (int i = 0; i < (1 << n_hash); i++) bin2mean[i] = ob.bin2mean[i];
	return 0;
}

// literal reorder on a flag array: failed[i] != 0 means the item stays; writes the source index of every
// output slot into perm[0..ret) and returns the new length
int kmxo_reorder(const uint8_t* failed, int n, int32_t* perm) {
	std::vector<Item> a(n > 0 ? n : 1, Item{ 0, 0 });
	for (int i = 0; i < n; i++) { a[i].kmer = (uint64_t)i; a[i].occ = failed[i] ? 1 : 0; }
	int r = reorder_literal(a.data(), n);
	for (int i = 0; i < r; i++) perm[i] = (int32_t)a[i].kmer;
	return r;
}

// listing: returns number of records listed, or -1; caller supplies capacity-checked buffers (may be NULL to size)
int64_t kmxo_list(const char* db_base, uint64_t* kmers, uint32_t* counts, int64_t cap, int32_t* k_out, uint64_t* total_out) {
	KmcDbO db;
	if (!db.open(db_base)) return -1;
	int64_t n = 0;
	db.for_each([&](uint64_t v, uint32_t c) { if (kmers && n < cap) { kmers[n] = v; counts[n] = c; } n++; });
	if (k_out) *k_out = (int32_t)db.k;
	if (total_out) *total_out = db.total;
	return n;
}

uint32_t kmxo_signature(uint64_t kmer, int k, int len) { return kmer_signature(kmer, k, len); }

// CheckKmer per packed k-mer; returns n or -1
int64_t kmxo_check_kmers(const char* db_base, const uint64_t* kmers, int64_t n, uint32_t* counts) {
	KmcDbO db;
	if (!db.open(db_base)) return -1;
	for (int64_t i = 0; i < n; i++) counts[i] = db_check_kmer(db, kmers[i]);
	return n;
}

// GetCountersForRead for reads stored back to back (read r = bytes offsets[r] .. offsets[r+1]); returns the counters written or -1
int64_t kmxo_counters_for_reads(const char* db_base, const char* bases, const int64_t* offsets, int64_t n_reads, uint32_t* counters) {
	KmcDbO db;
	if (!db.open(db_base)) return -1;
	int64_t at = 0;
	for (int64_t r = 0; r < n_reads; r++) at += db_counters_for_read(db, bases + offsets[r], offsets[r + 1] - offsets[r], counters + at);
	return at;
}

// build from a KMC db and save to out_dir; stats[3] = {attempts, accepted, rest}
int kmxo_build(const char* db_base, int ci, int cs, int nh, int nb, const char* out_dir, int64_t* stats) {
	KmcDbO db;
	if (!db.open(db_base)) return 1;
	ModelO m;
	if (!m.configure(ci, cs, nh, nb)) return 2;
	int rc = build_model(m, db, stats);
	if (rc) return rc;
	return m.save(out_dir) ? 0 : 6;
}

void* kmxo_load(const char* dir) {
	ModelO* m = new ModelO();
	if (!m->load(dir)) { delete m; return nullptr; }
	return m;
}
void kmxo_free(void* h) { delete (ModelO*)h; }
int kmxo_k(void* h) { return ((ModelO*)h)->k; }

void kmxo_query_packed(void* h, const uint64_t* kmers, int64_t n, int32_t* out) {
	const ModelO* m = (const ModelO*)h;
	for (int64_t i = 0; i < n; i++) out[i] = m->kmer_to_occ(kmers[i]);
}

// n strings of k characters at flat + i * stride, answered the way the reference answers strings (N / lower case included)
void kmxo_query_ascii(void* h, const char* flat, int64_t stride, int64_t n, int32_t* out) {
	const ModelO* m = (const ModelO*)h;
	for (int64_t i = 0; i < n; i++) out[i] = m->kmer_to_occ_str(std::string(flat + i * stride, (size_t)m->k));
}

// path classification for tests: 1 rest, 2 bf-only/absent, 3 array no candidate, 4 single, 5 single+vote, 6 multi
void kmxo_query_path(void* h, const uint64_t* kmers, int64_t n, int32_t* path) {
	const ModelO* m = (const ModelO*)h;
	for (int64_t i = 0; i < n; i++) {
		uint64_t v = canonical(kmers[i], m->k);
		if (m->rest.check(v)) { path[i] = 1; continue; }
		if (!m->filter_check(m->km_back, middle(v, m->k), m->k - 2, m->hk)) { path[i] = 2; continue; }
		int occ = m->check_all_bf(v);
		size_t nb = m->find_bitarray(v).size();
		path[i] = nb == 0 ? 3 : (nb == 1 ? (occ ? 5 : 4) : 6);
	}
}

}  // extern "C"
