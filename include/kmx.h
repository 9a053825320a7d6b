/* kmx.h -- C ABI of libkmx.so: the B200 (sm_100a) build + retrieval path of a kmcEx model.
 *
 * The reference (lzhLab/kmcEx) has no plugin ABI: its boundary is the header-only C++ class
 * KModel in kmodel.hpp.  include/kmodel.hpp in this repository is a drop-in replacement of
 * that header whose methods forward to the entry points below; the entry points are what a
 * cgo / JNI / ctypes binding of the same path would bind.  Every function cites the reference
 * interface it replaces (file:line relative to the reference root).
 *
 * Conventions
 *   - plain pointers and sizes only; HOST pointers unless the name ends in _device
 *   - int-returning functions return 0 on success and a non-zero KMX_E* code on failure;
 *     kmx_last_error() gives the message (the reference prints a message and calls exit(1),
 *     kmodel.hpp:394-397,682-685 -- the C++ shim keeps that behaviour on top of these codes)
 *   - there is NO CPU fallback: every compute entry point fails with KMX_ENOGPU when no
 *     sm_100-class CUDA device is usable
 *   - k-mers in packed form are 2 bits per base, first base in the most significant bits,
 *     A=0 C=1 G=2 T=3, right-aligned in a uint64_t (the value CKmerAPI holds for k <= 32,
 *     kmc_api/kmer_api.h:283-305)
 */
#ifndef KMX_H
#define KMX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMX_OK        0
#define KMX_EARG      1   /* bad argument / unsupported parameter combination               */
#define KMX_EIO       2   /* file cannot be opened / short read / short write               */
#define KMX_EFORMAT   3   /* not a KMC database or not a kmcEx model directory              */
#define KMX_ENOGPU    4   /* no usable CUDA device (there is no CPU path)                   */
#define KMX_ECUDA     5   /* a CUDA call or kernel failed                                   */
#define KMX_ERANGE    6   /* input outside what the reference defines (its UB corners)      */
#define KMX_ESTATE    7   /* call order: model not built / already built                    */

typedef struct kmx_model kmx_model;   /* opaque: a KModel (kmodel.hpp:39)                   */
typedef struct kmx_db kmx_db;         /* opaque: a KMC database opened for listing           */

/* what KModel::show_header_info / show_kmodel_info print (kmodel.hpp:118-169), plus timings */
typedef struct kmx_info_t {
	int32_t ci, cs, n_hash, n_bits, bf_num, k;
	uint64_t total_kmers;        /* CKMCFile::KmerCount()                    kmc_file.cpp:763 */
	uint64_t bf_kmers;           /* k-mers kept in the per-count Bloom filters               */
	uint64_t km_kmers;           /* k-mers offered to the coupled bit arrays                 */
	uint64_t rest_kmers;         /* entries of the rest table                rest.hpp:252    */
	uint64_t kmer_counts[3];
	uint64_t bf_bytes;           /* sum byte_bf + byte_bf_back                               */
	uint64_t km_bytes;           /* 2 * n_bits * km_byte_size                                */
	uint64_t km_back_bytes;
	uint64_t rest_bytes;         /* KRestData::get_all_byte_size             rest.hpp:256    */
	/* build statistics (0 for a loaded model) */
	uint64_t insert_attempts;    /* calls of insert_to_array                 kmodel.hpp:590  */
	uint64_t insert_accepted;
	uint64_t insert_iterations;  /* reservation iterations summed over all rounds            */
	uint64_t batches;
	uint64_t insert_phase_cycles[8];  /* SM cycles per phase of the insert kernel (diagnostic): reserve/commit of
	                                      the first iteration, of later iterations, tile scan, place, move */
	/* device times of the last build, milliseconds (CUDA events on the build stream) */
	float ms_upload, ms_count, ms_encode, ms_insert, ms_rest, ms_total_device;
	double build_time_cost;      /* host wall seconds of init, as the reference reports it   */
} kmx_info_t;

typedef struct kmx_db_info_t {
	uint32_t k, mode, counter_size, lut_prefix_length, signature_len, min_count, max_count, kmc_version;
	uint64_t total_kmers;
	uint64_t lut_entries;        /* LUT slots (bins * 4^lut_prefix_length), guard excluded   */
	uint64_t suffix_bytes;       /* record bytes of .kmc_suf (markers excluded)              */
	uint32_t record_bytes;
	int32_t on_device;
} kmx_db_info_t;

/* ---- process-wide ---------------------------------------------------------------------- */
const char* kmx_last_error(void);                 /* thread-local message of the last failure */
int kmx_device_count(void);                       /* usable CUDA devices (0 = none)           */
int kmx_set_device(int ordinal);                  /* device used by objects created afterwards */
const char* kmx_version(void);

/* ---- model lifetime: get_model(ci,cs,num_hash,num_bit) kmodel.hpp:674; get_model(dir) :680 */
kmx_model* kmx_create(int ci, int cs, int n_hash, int n_bits);
kmx_model* kmx_load(const char* dir);             /* header + km.bin + rest.bin -> device      */
void kmx_destroy(kmx_model* m);

/* ---- build: KModel::init(db_file) kmodel.hpp:57-86 ----------------------------------------
 * db_base is the KMC base name; ".kmc_pre"/".kmc_suf" are appended (kmc_file.cpp:77,86).     */
int kmx_init_from_kmc(kmx_model* m, const char* db_base);
/* the same build from an already opened (and possibly already device-resident) database     */
int kmx_init_from_db(kmx_model* m, kmx_db* db);

/* ---- persistence: KModel::save kmodel.hpp:173-206, KRestData::save_file rest.hpp:197-221 --
 * dir must exist (README.md:77).                                                             */
int kmx_save(kmx_model* m, const char* dir);

/* ---- retrieval: KModel::kmer_to_occ kmodel.hpp:90-116 -------------------------------------
 * n k-mers of length k (the model's k) as ASCII, record i at flat + i*stride (stride >= k);
 * upper-case ACGT is the supported alphabet.  out[i] = occurrence estimate.                  */
int kmx_query_ascii(kmx_model* m, const char* flat, size_t stride, size_t n, int32_t* out);
int kmx_query_packed(kmx_model* m, const uint64_t* kmers, size_t n, int32_t* out);
/* device-resident variants: pointers are CUDA device pointers on the model's device,
 * stream is a cudaStream_t (NULL = the model's own stream); asynchronous w.r.t. the host     */
int kmx_query_packed_device(kmx_model* m, const uint64_t* d_kmers, size_t n, int32_t* d_out, void* stream);
int kmx_query_ascii_device(kmx_model* m, const char* d_flat, size_t stride, size_t n, int32_t* d_out, void* stream);
/* per-query path class (test/diagnostic): 1 rest, 2 not in km_back, 3 no array candidate,
 * 4 one candidate, 5 one candidate + Bloom hit (neighbour vote), 6 several candidates        */
int kmx_query_path_packed(kmx_model* m, const uint64_t* kmers, size_t n, int32_t* path);

void kmx_info(const kmx_model* m, kmx_info_t* info);
int kmx_model_sync(kmx_model* m);                 /* wait for the model's stream               */

/* ---- KMC database listing: CKMCFile::OpenForListing / ReadNextKmer kmc_file.cpp:66-99,428-515 */
kmx_db* kmx_db_open(const char* db_base);         /* parse .kmc_pre (header + LUT), check .kmc_suf; records stay on disk */
int kmx_db_upload(kmx_db* db);                    /* LUT + records -> device through pinned bounce buffers (idempotent) */
void kmx_db_info(const kmx_db* db, kmx_db_info_t* info);
/* GPU listing: all records in file order, count filter of ReadNextKmer applied; returns the
 * number listed through *n_out; kmers/counts need room for total_kmers entries (host)        */
int kmx_db_list(kmx_db* db, uint64_t* kmers, uint32_t* counts, uint64_t* n_out);
void kmx_db_close(kmx_db* db);

/* ---- host-side pieces of the path, exposed for known-answer tests (no GPU needed) --------- */
uint64_t kmx_host_murmur64(const void* key, int len, uint32_t seed);   /* tools.hpp:16-50    */
uint64_t kmx_host_hash_packed(uint64_t kmer, int len, uint32_t seed);  /* same, on the ASCII expansion */
uint64_t kmx_host_canonical(uint64_t kmer, int k);                     /* tools.hpp:160-167  */
uint32_t kmx_host_seed(int i);                                         /* tools.hpp:9        */
int kmx_host_occubin(int max_counter, int n_hash, int32_t* occ2bin, int32_t* bin2mean); /* occu_bin.hpp:27-83 */
/* filter sizes as the reference computes them (kmodel.hpp:402-456); bytes[0..2]=byte_bf,
 * bytes[3..5]=byte_bf_back, bytes[6]=km_byte_size, bytes[7]=byte_km_back                     */
void kmx_host_sizes(const uint64_t kmer_counts[3], int bf_num, uint64_t km_kmers, int n_hash, uint64_t bytes[8]);
uint64_t kmx_host_fastmod(uint64_t h, uint64_t d);                     /* the device's exact h % d */
/* survivor permutation of reorder_buffer (kmodel.hpp:529-540) in closed form: failed[i] != 0
 * keeps item i; perm[j] = source index of output slot j; returns the new length            */
int kmx_host_reorder(const uint8_t* failed, int n, int32_t* perm);

/* ---- k-mer counting: the stage the reference delegates to the external `kmc` binary (main.cpp:136-140)
 * Plain-text 4-line FASTQ files -> <out_base>.kmc_pre/.kmc_suf (KMC 2/3 layout, one bin): canonical k-mers, windows
 * with a non-ACGT base skipped, k-mers seen fewer than ci times dropped, counters saturated at cs.            */
typedef struct kmx_count_info_t {
	uint64_t n_reads, n_windows, n_unique, n_kept;
	uint32_t lut_prefix_length, counter_size;
} kmx_count_info_t;
int kmx_count_fastq(const char* const* fastq_paths, int n_files, int k, int ci, int cs, const char* out_base, kmx_count_info_t* info);

/* ---- multi-GPU build: array-owner decomposition (option A) -------------
 * One process per GPU, `world` of them.  Every rank calls prepare: it runs the counting pass, inserts the
 * Bloom-bound records of ITS share of the database (record range rank/world) into its copy of the filters,
 * writes the item stream if it owns a coupled array (rank < n_active), allocates its exchange buffers and
 * returns two 64-byte CUDA IPC handles (survivor exchange slab, filter slab).  The ranks exchange the handles
 * (any out-of-band channel) and call connect with the handles of ranks 0..world-1 concatenated (128 bytes
 * each).  merge(0) ORs the partial Bloom filters over the ranks through peer memory (one kernel: reduce-scatter
 * + all-gather with flag barriers, no NCCL); insert: rank r < n_active owns the coupled arrays a with
 * a % n_active == r, the persistent kernels pass each bucket's survivors to the next owner through peer memory
 * with a flag barrier per round; merge(1) ORs km_back.  buffers() exposes the device pointers the caller's
 * collectives complete (owned arrays are broadcast, survivor lists are concatenated); finish() builds the rest
 * table.  Results are identical to kmx_init_from_db for every world / n_active.                       */
typedef struct kmx_dist_buffers_t {
	int32_t n_bits;
	uint64_t cell_bytes;              /* bytes of one coupled array in the device layout          */
	void* cells[8];                   /* device pointers, array a valid on its owner after insert */
	void* km_back;
	uint64_t km_back_bytes;
	void* rest_kmer;                  /* this rank's survivors: u64 k-mers ...                    */
	void* rest_occ;                   /* ... and u32 counts                                       */
	uint64_t rest_n;
	uint64_t insert_attempts, insert_accepted;
} kmx_dist_buffers_t;
int kmx_dist_prepare(kmx_model* m, kmx_db* db, int rank, int n_active, int world, void* ipc_handles_out /* 128 bytes */);
int kmx_dist_connect(kmx_model* m, const void* handles /* world * 128 bytes */);
int kmx_dist_merge(kmx_model* m, int which /* 0: Bloom filters, 1: km_back */);
int kmx_dist_insert(kmx_model* m);
int kmx_dist_buffers(kmx_model* m, kmx_dist_buffers_t* out);
int kmx_dist_finish(kmx_model* m, const uint64_t* d_rest_kmer, const uint32_t* d_rest_occ, uint64_t rest_n,
                    uint64_t attempts, uint64_t accepted);

/* ---- roofline denominators: random 32-byte-sector throughput of the device (diagnostic) ------
 * kind 0: random 8-byte loads, 1: random 32-bit atomic OR, 2: random 64-bit atomic OR; 7 accesses
 * per item over `footprint_bytes` of device memory; *ms_out = milliseconds per launch             */
int kmx_microbench_random(int kind, uint64_t footprint_bytes, uint64_t n_items, int reps, float* ms_out);
/* the same with locality: the accesses of items that run together fall into one window of window_bytes (0 = none) */
/* microseconds per grid-wide barrier of a co-resident kernel: mode 0 cooperative_groups grid.sync(), 1 the library's counter barrier */
int kmx_microbench_grid_barrier(int mode, int threads, int blocks_per_sm, int reps, float* us_out);
/* nanoseconds per returning atomicAdd when every warp of a full grid hammers n_counters addresses (list appends) */
int kmx_microbench_hot_atomic(int n_counters, int per_warp, float* ns_out);
int kmx_microbench_windowed(int kind, uint64_t footprint_bytes, uint64_t window_bytes, uint64_t n_items, int reps, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* KMX_H */
