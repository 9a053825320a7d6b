// kmx_device.cuh -- device-side view of a model and of a KMC database, shared by the build
// and query kernels.  Layout in HBM ("Data layout"):
//   * every Bloom filter is the reference's byte array, addressed as little-endian u32 words
//   * every coupled array pair (bit_array_1 = value, bit_array_2 = tag, kmodel.hpp:32-37) is
//     ONE array of u64 cells: cell w = (tag_word[w] << 32) | value_word[w], so a probe reads
//     both bits of a position with one 8-byte load and an insert sets both with one 64-bit
//     atomic OR; save() de-interleaves back into the on-disk order
//   * the rest table keeps the sorted full 2-bit k-mers (u64) + counts, plus the reference's
//     hash2index / pre_buffer group index (rest.hpp:95-135)
#pragma once
#include "kmx_core.cuh"

namespace kmx {

struct DevFilter {
	uint32_t* words;
	FastMod mod;          // mod.d = length in bits
};

struct DevRest {
	const int32_t* hash2index;       // [map_size]  dense group id or -1
	const int32_t* pre_buffer;       // [groups+1]  first entry of each group, cumulative
	const uint64_t* keys;            // [count]     full packed k-mers, ascending
	const int32_t* counts;           // [count]
	uint64_t count;
	uint64_t suffix_mask;            // low 2*(k-pre_len) bits
	int suffix_bits;
	int k;
};

struct DevModel {
	int k, n_hash, n_bits, bf_num, ci, cs, hb, hk, end1;
	DevFilter bf[kMaxBf], bf_back[kMaxBf], km_back;
	unsigned long long* cells[kMaxArrays];
	FastMod arr_mod;                 // bit_array_length of every pair (kmodel.hpp:445)
	uint32_t arr_seed[kMaxArrays][kMaxHash];   // kmodel.hpp:450-453
	const uint16_t* occ2bin;         // [cs+1]      occu_bin.hpp:67-77
	const int32_t* bin2mean;         // [1<<n_hash] occu_bin.hpp:79-83
	DevRest rest;
};

struct DevDb {
	const uint8_t* suf;              // record bytes, markers stripped, 16-byte aligned
	const uint64_t* lut;             // [lut_entries + 1], guard = total + 1
	uint64_t lut_entries;
	uint64_t total;
	uint64_t prefix_mask;            // 4^lut_prefix_length - 1
	uint32_t suffix_bytes, counter_bytes, rec_bytes;
	uint32_t min_count, max_count;
	int k;
};

#ifdef __CUDACC__

// ---- bit probes -------------------------------------------------------------------------
KMX_D bool filter_test(const DevFilter& f, uint64_t h) {
	uint64_t pos = fastmod(h, f.mod);
	return (__ldg(f.words + (pos >> 5)) & bit_mask32(pos)) != 0;
}
KMX_D void filter_set(const DevFilter& f, uint64_t h) {
	uint64_t pos = fastmod(h, f.mod);
	atomicOr(f.words + (pos >> 5), bit_mask32(pos));
}

// ---- rest table: KRestData::check_kmer (rest.hpp:223-251) ----------------------------------
// Binary search over [pre_buffer[g], pre_buffer[g+1]] with the reference's INCLUSIVE upper
// bound: the probe can land on the first entry of the next group, where only suffix bytes are
// compared, which yields the reference's false hits.  A probe at index == count is out of
// bounds in the reference and counts as "no match".
KMX_D int rest_lookup(const DevRest& R, uint64_t v) {
	if (R.count == 0 && R.hash2index == nullptr) return 0;
	uint32_t pre = (uint32_t)(v >> R.suffix_bits);
	int g = __ldg(R.hash2index + pre);
	if (g < 0) return 0;
	uint64_t key = v & R.suffix_mask;
	long long low = __ldg(R.pre_buffer + g), high = __ldg(R.pre_buffer + g + 1);
	while (low <= high) {
		long long mid = (low + high) >> 1;
		if ((uint64_t)mid >= R.count) return 0;
		uint64_t s = __ldg(R.keys + mid) & R.suffix_mask;
		if (key < s) high = mid - 1;
		else if (key > s) low = mid + 1;
		else return __ldg(R.counts + mid);
	}
	return 0;
}

#endif  // __CUDACC__

}  // namespace kmx
