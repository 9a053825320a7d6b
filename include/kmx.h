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
	uint64_t insert_phase_cycles[12]; /* SM cycles per phase of the insert kernel as seen by one thread (diagnostic): [0] phase 0,
	                                     [1] phase 1, [3] contested passes, [5] place, [6] move, [7] cross-GPU wait; the part spent
	                                     in the thread's own loop, the rest being barrier wait: [2] phase 0, [4] phase 1, [8] place,
	                                     [9] move, [10] contested passes; [11] contested passes run */
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
	uint32_t both_strands;       /* 1: the database holds canonical k-mers (kmc_file.cpp:208-209) */
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
 * out[i] = occurrence estimate.                                                              */
int kmx_query_ascii(kmx_model* m, const char* flat, size_t stride, size_t n, int32_t* out);
int kmx_query_packed(kmx_model* m, const uint64_t* kmers, size_t n, int32_t* out);
/* ASCII queries are not validated by the reference (tools.hpp:63-76,160-167): a byte that is not C/G/T encodes as A for
 * the canonical-form decision and the rest lookup, but when the string's forward orientation is the canonical one the
 * filters are probed with hashes of its RAW bytes (N, lower case included).  kmx_query_ascii* reproduce exactly that.
 * device-resident variants: pointers are CUDA device pointers on the model's device,
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

/* ---- KMC random access on the device-resident database: CKMCFile::OpenForRA / CheckKmer
 * kmc_file.cpp:27-58,320-356,1358-1436 and GetCountersForRead kmc_file.cpp:879-897,1130-1352.  Not used by kmcEx itself; it is
 * the exact-count ground truth for accuracy reports on kmer_to_occ (tools/accuracy_report.py).
 * kmx_db_check_kmers: one CheckKmer per packed k-mer (first base most significant; NOT canonicalised, like CheckKmer): the bin
 * comes from the k-mer's signature and the database's signature map, the LUT slot from its first lut_prefix_length bases, the
 * record from a binary search over the suffixes; counts[i] = the counter, or 0 when the k-mer is absent or its counter lies
 * outside [min_count, max_count].  All pointers are host pointers.
 * kmx_db_counters_for_reads: GetCountersForRead for n_reads reads stored back to back in `bases` (read r = bytes
 * offsets[r] .. offsets[r+1]); a read of length L yields max(0, L - k + 1) counters, all reads' counters back to back in
 * `counters` (room for the sum); a window with a byte other than ACGTacgt counts 0; canonical k-mers are looked up when the
 * database holds both strands.  *n_counters_out (may be NULL) = counters written.                                          */
int kmx_db_check_kmers(kmx_db* db, const uint64_t* kmers, int64_t n, uint32_t* counts);
/* CKMCFile::SetMinCount / SetMaxCount / ResetMinMaxCounts kmc_file.cpp:670-734: narrow the counter range (inside the header's)
 * that the listing and the random-access calls apply; kmx_db_info reports the current range                                */
int kmx_db_set_count_range(kmx_db* db, uint32_t min_count, uint32_t max_count);
int kmx_db_reset_count_range(kmx_db* db);
int kmx_db_counters_for_reads(kmx_db* db, const char* bases, const int64_t* offsets, int64_t n_reads, uint32_t* counters, int64_t* n_counters_out);

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
uint32_t kmx_host_signature(uint64_t kmer, int k, int signature_len);  /* CKmerAPI::get_signature, kmer_api.h:653-673 (len 5..11) */
/* survivor permutation of reorder_buffer (kmodel.hpp:529-540) in closed form: failed[i] != 0
 * keeps item i; perm[j] = source index of output slot j; returns the new length            */
int kmx_host_reorder(const uint8_t* failed, int n, int32_t* perm);

/* ---- k-mer counting: the stage the reference delegates to the external `kmc` binary (main.cpp:136-140)
 * 4-line FASTQ files (plain text or gzip) -> <out_base>.kmc_pre/.kmc_suf (KMC 2/3 layout, one bin): canonical k-mers, windows
 * with a non-ACGT base skipped, k-mers seen fewer than ci times dropped, counters saturated at cs.            */
typedef struct kmx_count_info_t {
	uint64_t n_reads, n_windows, n_unique, n_kept;
	uint32_t lut_prefix_length, counter_size;
} kmx_count_info_t;
int kmx_count_fastq(const char* const* fastq_paths, int n_files, int k, int ci, int cs, const char* out_base, kmx_count_info_t* info);

/* ---- multi-GPU build: ONE model built by the GPUs of one node ----------------
 * KModel::init (kmodel.hpp:57-86) over `world` GPUs; the result is byte-identical to kmx_init_from_kmc on one GPU for every
 * world size.  Record range, Bloom inserts and the rest-table sort are sharded over all ranks; the coupled arrays are split
 * by ownership (array a on rank a % min(world, n_bits) -- the reference's own decomposition, kmodel.hpp:561-565); array-bound
 * k-mers go from the decoding rank straight into the owner's memory, survivors from owner to owner, the finished pieces to
 * every rank -- all through peer-mapped device memory (NVLink), no NCCL.  Every rank ends with the complete model.
 *
 * (a) inside ONE process: kmx_set_devices({d0, d1, ...}) (or KMX_GPUS=N in the environment) before kmx_init_from_kmc -- one
 *     host thread per GPU, plain peer access.  The model handle then also shards kmx_query_* batches over its replicas.
 * (b) one process per GPU: every rank runs the kmx_team_steps() steps in lock step; after each step the ranks all-gather
 *     their kmx_team_blob_bytes()-byte blob (any channel: MPI, a process-group all-gather, a pipe) and pass the `world` blobs, in rank
 *     order, to the next step.  The blobs carry counts, CUDA IPC handles and return codes; bulk data never touches them.
 *     A rank runs every step even after a failure (the codes travel in the blobs and fail the next step on all ranks). */
int kmx_set_devices(const int* ordinals, int n);      /* n >= 2: team of these GPUs for kmx_init_from_kmc; n <= 1: single GPU */
int kmx_db_upload_share(kmx_db* db, int rank, int world);   /* only this rank's tile range of the records -> device */
int kmx_team_steps(void);
int kmx_team_blob_bytes(void);
int kmx_team_step(kmx_model* m, kmx_db* db, int rank, int world, int step, const void* blobs_in, void* blob_out);

/* host-side pieces of the team build, for tests: where stream item g goes (owner rank, index in the owner's shard), and the
 * prefix ranges of the sharded rest build (cut[world + 1], cut_off[world + 1]) */
void kmx_host_route(uint64_t g, int n_active, int n_bits, int32_t* owner, uint64_t* index);
int kmx_host_prefix_cuts(const uint32_t* hist, int map_size, int world, uint32_t* cut, uint64_t* cut_off);

/* device-side known-answer hook: the hash -> exact modulo -> word / bit addressing chain of the build and query kernels with an
 * arbitrary array length d (the NA12878 shape has d ~ 1.1e10 > 2^32).  kmers[n] are hashed with seeds[n_seeds] modulo d on
 * the device and set as tag + value in a cell array and as bits in a Bloom-style filter of d bits; pos_out[n * n_seeds] gets
 * the positions, counts[0] = positions found set again, counts[1] / counts[2] = tag / filter bits set in the whole arrays
 * (an address truncated anywhere would make them differ from the number of distinct positions). */
int kmx_selftest_positions(const uint64_t* kmers, size_t n, int k, uint64_t d, const uint32_t* seeds, int n_seeds, uint64_t* pos_out,
                           uint64_t counts[3]);

/* position-sensitive 64-bit checksums of the model's device arrays: sums[0] Bloom filters + km_back, [1] coupled arrays, [2] rest
 * keys, [3] rest counts + group index.  Equal on every replica of a team build (bench.py compares them across the ranks when
 * the model is too large to save on every rank; rank 0's files are compared with the reference's byte for byte). */
int kmx_model_checksum(kmx_model* m, uint64_t sums[4]);

/* kernels launched by this process's libkmx so far (bench.py reports the launches inside its timed region) */
unsigned long long kmx_launch_count(void);

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
/* the same random accesses issued by device `src` into a buffer on device `dst` over NVLink (peer access): what routing every
 * Bloom bit to an address-range owner would cost, against OR-reducing replicated filters */
int kmx_microbench_peer_random(int kind, int src, int dst, uint64_t footprint_bytes, uint64_t n_items, int reps, float* ms_out);
/* milliseconds per pass of a read-only streaming kernel over `bytes` of device memory: the ceiling of the counting pass */
int kmx_microbench_stream_read(uint64_t bytes, int blocks_per_sm, int reps, float* ms_out);

#ifdef __cplusplus
}
#endif
#endif /* KMX_H */
