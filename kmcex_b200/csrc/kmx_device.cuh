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
	const int32_t* hash2index;       // [map_size]  dense group id or -1          (on-disk form)
	const int32_t* pre_buffer;       // [groups+1]  first entry of each group      (on-disk form)
	const uint64_t* keys;            // [count]     full packed k-mers, ascending
	const int32_t* counts;           // [count]
	const uint32_t* fine;            // [2^fine_bits + 1] first entry of each bucket of the top fine_bits key bits
	const uint64_t* quirk_suffix;    // [map_size]  the one suffix that false-hits for this prefix, or ~0
	const uint32_t* quirk_index;     // [map_size]  entry whose count that false hit returns
	uint64_t count;
	uint64_t suffix_mask;            // low 2*(k-pre_len) bits
	int suffix_bits;
	int fine_bits, fine_shift;       // bucket = key >> fine_shift
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
	int query_l2;                    // L2 policy of the query probes for models beyond the L2: 1 km_back evict-last, 2 arrays evict-first, 4 Bloom evict-first
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

// ---- fire-and-forget reductions -----------------------------------------------------------
// atomicOr / atomicAdd / atomicMin with the result unused still compile to ATOMG (a returning atomic) inside the
// cooperative insert kernel; the scattered bit traffic of this path wants RED: no response packet, no scoreboard slot.
KMX_D void red_or32(uint32_t* p, uint32_t v) { asm volatile("red.global.or.b32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(v) : "memory"); }
KMX_D void red_or64(unsigned long long* p, unsigned long long v) { asm volatile("red.global.or.b64 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "l"(v) : "memory"); }
KMX_D void red_add32(uint32_t* p, uint32_t v) { asm volatile("red.global.add.u32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(v) : "memory"); }
KMX_D void red_min32(uint32_t* p, uint32_t v) { asm volatile("red.global.min.u32 [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(v) : "memory"); }

// ---- random probes of arrays far larger than the L2 -----------------------------------------
// A probe sector of an HBM-resident array is not touched again before hundreds of megabytes of other sectors have
// passed through the L2: let it be the first thing the L2 drops, so that the small hot structures of the insert
// (claim bitmaps, status words, lists) stay resident instead of being washed out by the probe stream.
KMX_D unsigned long long make_evict_first_policy() {
	unsigned long long pol;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
KMX_D unsigned long long make_evict_last_policy() {
	unsigned long long pol;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
KMX_D unsigned long long make_evict_normal_policy() {
	unsigned long long pol;
	asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
	return pol;
}
KMX_D unsigned long long ld_stream64(const unsigned long long* p, unsigned long long pol) {
	unsigned long long v;
	asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(__cvta_generic_to_global(p)), "l"(pol));
	return v;
}
KMX_D void red_or64_stream(unsigned long long* p, unsigned long long v, unsigned long long pol) {
	asm volatile("red.global.or.L2::cache_hint.b64 [%0], %1, %2;" ::"l"(__cvta_generic_to_global(p)), "l"(v), "l"(pol) : "memory");
}
KMX_D void red_or32_stream(uint32_t* p, uint32_t v, unsigned long long pol) {
	asm volatile("red.global.or.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(__cvta_generic_to_global(p)), "r"(v), "l"(pol) : "memory");
}

// read-only probes with an L2 eviction policy (query kernels; models that do not fit the L2)
KMX_D uint32_t ldg_hint32(const uint32_t* p, unsigned long long pol) {
	uint32_t v;
	asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(__cvta_generic_to_global(p)), "l"(pol));
	return v;
}
KMX_D unsigned long long ldg_hint64(const unsigned long long* p, unsigned long long pol) {
	unsigned long long v;
	asm volatile("ld.global.nc.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(__cvta_generic_to_global(p)), "l"(pol));
	return v;
}

// ---- bit probes -------------------------------------------------------------------------
KMX_D bool filter_test(const DevFilter& f, uint64_t h) {
	uint64_t pos = fastmod(h, f.mod);
	return (__ldg(f.words + (pos >> 5)) & bit_mask32(pos)) != 0;
}
KMX_D bool filter_test_hint(const DevFilter& f, uint64_t h, unsigned long long pol) {
	uint64_t pos = fastmod(h, f.mod);
	return (ldg_hint32(f.words + (pos >> 5), pol) & bit_mask32(pos)) != 0;
}
KMX_D void filter_set(const DevFilter& f, uint64_t h) {
	uint64_t pos = fastmod(h, f.mod);
	red_or32(f.words + (pos >> 5), bit_mask32(pos));
}

#endif  // __CUDACC__

}  // namespace kmx
