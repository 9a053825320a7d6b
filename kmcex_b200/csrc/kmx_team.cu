// kmx_team.cu -- KModel::init (kmodel.hpp:57-86) spread over the GPUs of one node: ONE model built by `world` ranks, one GPU
// each, either one process per GPU (the caller moves four 256-byte blobs per rank between the steps with any all-gather it
// has -- kmcex_b200/distributed.py uses torch.distributed) or one host thread per GPU inside this process
// (kmx_set_devices / KMX_GPUS, used by KModel::init of include/kmodel.hpp).  All data moves through peer-mapped device memory
// (NVLink / NVSwitch): CUDA IPC handles between processes, plain peer access inside one.  No NCCL on the data path.
//
// What is sharded, and the reference code each part reproduces:
//   decode            rank r uploads, counts and decodes only the records of ITS tile range (kmc_file.cpp:428-515); the class
//                     counts are summed and the item offsets chained through the step-0 blobs (file order defines batch and
//                     bucket membership, kmodel.hpp:508-518)
//   Bloom filters     every rank inserts the Bloom-bound records of its range into its own copy, the copies are OR-ed by
//                     or_allreduce_kernel (kmodel.hpp:473-506: order-free)
//   coupled arrays    array a lives on rank a % n_active, n_active = min(world, n_bits) -- the reference's own decomposition
//                     (kmodel.hpp:561-565).  The decoding rank stores an array-bound k-mer straight into the item shard of the
//                     owner of its round-0 array (the exchange of k-mer batches over NVLink); the owners' persistent insert
//                     kernels hand each bucket's survivors to the next owner through peer memory with a flag barrier per round
//   km_back           OR-ed like the Bloom filters (kmodel.hpp:546-550); the same kernel then pulls the arrays a rank does
//                     not own from their owners, so that every rank ends with the complete model (replicated for the query)
//   rest table        the survivors are split by 7-base prefix range (balanced on the global prefix histogram): rank r gathers
//                     its range from every owner's list, sorts it and pushes the run into every rank's table (rest.hpp:95-135)
// The result is byte-identical to kmx_init_from_db for every world size.
#include <unistd.h>
#include <algorithm>
#include <condition_variable>
#include <thread>
#include "kmx_internal.h"

using namespace kmx;

#define fail kmx::set_error

namespace kmx {

constexpr int kTeamBlob = 256;
constexpr int kTeamSteps = 4;

struct XLayout {                      // exchange slab of one rank: what its peers write into or read from
	size_t item_k = 0, item_o = 0, buf_k[2] = { 0, 0 }, buf_o[2] = { 0, 0 }, ctl = 0, flags = 0, hist = 0, rest_k = 0, rest_o = 0, bytes = 0;
	uint64_t shard_items = 0, rest_cap = 0;
};

struct MLayout {                      // model slab (the same on every rank): Bloom filters | km_back | coupled arrays | link flags
	size_t bf[3] = { 0, 0, 0 }, bb[3] = { 0, 0, 0 }, bloom_bytes = 0, kmback = 0, kmback_bytes = 0, cells[kMaxArrays] = { 0 }, cell_bytes = 0,
	       flags = 0, bytes = 0;
};

struct TeamState {
	int rank = 0, world = 1, n_active = 1;
	kmx_db* db = nullptr;
	uint64_t n_tiles = 0, tile_lo = 0, tile_hi = 0;
	uint64_t* d_tile_off = nullptr;
	CountOut local = {};
	uint64_t item_base = 0;
	float ms_count = 0;
	MLayout ml;
	XLayout xl[kMaxRanks];
	void* xslab = nullptr;
	void* peer_m[kMaxRanks] = { nullptr };
	void* peer_x[kMaxRanks] = { nullptr };
	void* peer_r[kMaxRanks] = { nullptr };
	uint32_t seq = 0;                 // cross-GPU barriers of the exchange kernels completed so far
	// rest table
	std::vector<uint32_t> hist;       // global prefix histogram
	uint64_t owner_rest_n[kMaxRanks] = { 0 };
	uint64_t rest_total = 0, my_off = 0, my_n = 0;
	uint32_t prefix_lo = 0, prefix_hi = 0;
	uint64_t* d_gk = nullptr;         // this rank's prefix range: gathered, then sorted
	uint32_t* d_go = nullptr;
	uint64_t* d_sk = nullptr;
	uint32_t* d_so = nullptr;
	unsigned long long* d_gn = nullptr;
	void* d_sort_temp = nullptr;
	InsertCtl ctl = {};               // this rank's insert statistics (zero on ranks that own no array)
};

struct TeamBlob0 {
	int32_t rc, pid;
	uint64_t total_kmers;
	CountOut cnt;
};
struct TeamBlob1 {
	int32_t rc, pid, device, pad;
	cudaIpcMemHandle_t hm, hx;
	uint64_t ptr_m, ptr_x;
};
struct TeamBlob2 {
	int32_t rc, pid, device, pad;
	cudaIpcMemHandle_t hr;
	uint64_t ptr_r;
	uint64_t attempts, accepted, iterations;
	uint64_t phase_cycles[12];
	float ms_encode, ms_insert;
};
static_assert(sizeof(TeamBlob0) <= kTeamBlob && sizeof(TeamBlob1) <= kTeamBlob && sizeof(TeamBlob2) <= kTeamBlob, "a step's blob is 256 bytes");

static XLayout x_layout(const kmx_model* m, uint64_t n_batches, int rank, int n_active) {
	XLayout L;
	const uint64_t batch_items = (uint64_t)m->n_bits << kBucketLog;
	size_t off = 0;
	auto take = [&](size_t bytes) {
		const size_t at = off;
		off += up256(bytes);
		return at;
	};
	if (rank < n_active) {
		L.shard_items = n_batches * buckets_per_batch(rank, n_active, m->n_bits) * (uint64_t)kBucket;
		L.rest_cap = L.shard_items + (uint64_t)m->n_bits;    // nothing accepted + the stale-slot duplicates (kmodel.hpp:520-540)
		L.item_k = take((L.shard_items + 1) * 8);
		L.item_o = take((L.shard_items + 1) * 4);
		for (int q = 0; q < 2; q++) L.buf_k[q] = take(batch_items * 8);
		for (int q = 0; q < 2; q++) L.buf_o[q] = take(batch_items * 4);
		L.rest_k = take(L.rest_cap * 8);
		L.rest_o = take(L.rest_cap * 4);
	}
	L.ctl = take(sizeof(InsertCtl));
	L.flags = take(kMaxRanks * 4);
	L.hist = take(((size_t)1 << (2 * rest_prefix_len(m->k))) * 4);
	L.bytes = off;
	return L;
}

static MLayout m_layout(const kmx_model* m) {
	MLayout L;
	size_t off = 0;
	auto take = [&](size_t bytes) {
		const size_t at = off;
		off += up256(bytes);
		return at;
	};
	for (int i = 0; i < m->bf_num; i++) {
		L.bf[i] = take(pad8(m->bytes[i]));
		L.bb[i] = take(pad8(m->bytes[3 + i]));
	}
	L.bloom_bytes = off;
	L.kmback = take(pad8(m->bytes[7]));
	L.kmback_bytes = off - L.kmback;
	L.cell_bytes = up256((cell_words(m->bytes[6]) + 1) * 8);
	for (int i = 0; i < m->n_bits; i++) L.cells[i] = take(L.cell_bytes);
	L.flags = take(256);                                  // kMaxRanks barrier counters + the error word at +128
	L.bytes = off;
	return L;
}

void team_state_free(kmx_model* m) {
	TeamState* t = m->bs.team;
	if (!t) return;
	cudaStream_t s = m->x->stream;
	cudaStreamSynchronize(s);
	dev_free(t->d_tile_off, s);
	dev_free(t->d_gk, s); dev_free(t->d_go, s); dev_free(t->d_sk, s); dev_free(t->d_so, s); dev_free(t->d_gn, s); dev_free(t->d_sort_temp, s);
	if (t->xslab) slab_release(t->xslab, t->xl[t->rank].bytes, m->device);
	delete t;
	m->bs.team = nullptr;
}

// a peer's slab as seen from this device: the raw pointer inside one process (peer access), an IPC mapping otherwise
namespace {
std::mutex g_map_mu;
std::vector<std::pair<std::string, void*>> g_ipc_open;    // 64 handle bytes -> mapping in this process
}

static int map_peer(int my_device, int32_t peer_pid, int peer_device, const cudaIpcMemHandle_t& h, uint64_t raw, void** out) {
	if (peer_pid == (int32_t)getpid()) {
		if (peer_device != my_device) {
			int can = 0;
			CU(cudaDeviceCanAccessPeer(&can, my_device, peer_device));
			if (!can) return fail(KMX_ECUDA, "device %d cannot access device %d's memory (no peer path)", my_device, peer_device);
			cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(KMX_ECUDA, "cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(e));
			cudaGetLastError();
		}
		*out = (void*)(uintptr_t)raw;
		return KMX_OK;
	}
	std::string key((const char*)&h, sizeof(h));
	std::lock_guard<std::mutex> lock(g_map_mu);
	for (auto& kv : g_ipc_open) {
		if (kv.first == key) {
			*out = kv.second;
			return KMX_OK;
		}
	}
	void* mapped = nullptr;
	CU(cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess));
	g_ipc_open.emplace_back(key, mapped);
	*out = mapped;
	return KMX_OK;
}

// sharded rest build: prefix ranges with about total / world entries each.  Rank r takes the prefixes [cut[r], cut[r+1]),
// whose entries start at cut_off[r] in the sorted table (every rank computes the same cuts from the same histogram).
static void prefix_cuts(const uint32_t* hist, int map_size, int world, uint32_t* cut, uint64_t* cut_off) {
	uint64_t total = 0;
	for (int p = 0; p < map_size; p++) total += hist[p];
	uint64_t acc = 0;
	int r = 0;
	cut[0] = 0;
	cut_off[0] = 0;
	for (int p = 0; p < map_size; p++) {
		while (r + 1 < world && acc >= total * (uint64_t)(r + 1) / (uint64_t)world) {
			r++;
			cut[r] = (uint32_t)p;
			cut_off[r] = acc;
		}
		acc += hist[p];
	}
	while (r + 1 < world) {
		r++;
		cut[r] = (uint32_t)map_size;
		cut_off[r] = acc;
	}
	cut[world] = (uint32_t)map_size;
	cut_off[world] = acc;
}

static TeamLink team_link(kmx_model* m, TeamState* t) {
	TeamLink L;
	memset(&L, 0, sizeof(L));
	L.rank = t->rank;
	L.world = t->world;
	for (int p = 0; p < t->world; p++) L.flags[p] = (uint32_t*)((uint8_t*)t->peer_m[p] + t->ml.flags);
	L.seq = t->seq;
	L.error = (unsigned int*)((uint8_t*)m->mslab + t->ml.flags + 128);
	return L;
}

// ---- step 0: this rank's share of the records -> device, counting pass over it -------------------------------
static int team_step0(kmx_model* m, kmx_db* db, int rank, int world, TeamBlob0* out) {
	if (m->built || m->bs.team) return fail(KMX_ESTATE, "model already initialised (KModel::init is one-shot)");
	BuildState& b = m->bs;
	b.wall0 = std::chrono::high_resolution_clock::now();
	int rc = model_attach_device(m);
	if (rc) return rc;
	CU(cudaSetDevice(m->device));
	TeamState* t = new TeamState();
	b.team = t;
	t->rank = rank;
	t->world = world;
	t->n_active = std::min(world, m->n_bits);
	t->db = db;
	m->k = (int)db->info.k;
	m->total_kmers = db->info.total_kmers;
	if (m->k < 3) return fail(KMX_ERANGE, "k=%d: the (k-2)-mer filters need k >= 3", m->k);
	t->n_tiles = (m->total_kmers + kTile - 1) / kTile;
	t->tile_lo = t->n_tiles * (uint64_t)rank / (uint64_t)world;
	t->tile_hi = t->n_tiles * (uint64_t)(rank + 1) / (uint64_t)world;
	const int cores = (int)std::thread::hardware_concurrency();
	rc = db_upload_range(db, t->tile_lo * kTile, t->tile_hi * kTile, std::max(2, cores / world));
	if (rc) return rc;
	if (db->device != m->device) return fail(KMX_EARG, "database is on device %d, model on device %d", db->device, m->device);
	b.ms_upload = db->ms_upload;
	TRACE(b.wall0, "team: share of the database on device");
	cudaStream_t s = m->x->stream;
	cudaEvent_t* ev = m->x->ev_build;
	const uint64_t my_tiles = t->tile_hi - t->tile_lo;
	uint32_t* d_tile_cnt = nullptr;
	CountOut* d_count = nullptr;
	DevScope scope(s);
	if ((rc = scope.alloc(&d_tile_cnt, (my_tiles + 1) * 4))) return rc;
	if ((rc = scope.alloc(&d_count, sizeof(CountOut)))) return rc;
	DA(&t->d_tile_off, (my_tiles + 1) * 8, s);
	DevDb d = dev_db(db);
	CU(cudaEventRecord(ev[0], s));
	CU(cudaMemsetAsync(d_count, 0, sizeof(CountOut), s));
	CU(launch_count(d, m->ci, m->cs, m->bf_num, d_count, d_tile_cnt, t->tile_lo, t->tile_hi, m->sm_count, s));
	CU(launch_tile_scan(d_tile_cnt, my_tiles, t->d_tile_off, s));
	CountOut& cnt = m->x->h_pinned->count;
	CU(cudaMemcpyAsync(&cnt, d_count, sizeof(cnt), cudaMemcpyDeviceToHost, s));
	CU(cudaEventRecord(ev[1], s));
	CU(cudaStreamSynchronize(s));
	CU(cudaEventElapsedTime(&t->ms_count, ev[0], ev[1]));
	t->local = cnt;
	out->total_kmers = m->total_kmers;
	out->cnt = cnt;
	TRACE(b.wall0, "team: counted");
	return KMX_OK;
}

// ---- step 1: global counts -> sizes; the two slabs peers map ---------------------------------------------------
static int team_step1(kmx_model* m, const TeamBlob0* in, size_t stride, TeamBlob1* out) {
	TeamState* t = m->bs.team;
	if (!t) return fail(KMX_ESTATE, "team build: step 0 first");
	BuildState& b = m->bs;
	cudaStream_t s = m->x->stream;
	CU(cudaSetDevice(m->device));
	auto blob = [&](int p) { return (const TeamBlob0*)((const uint8_t*)in + (size_t)p * stride); };
	uint64_t cls[3] = { 0, 0, 0 }, bad = 0, items = 0;
	t->item_base = 0;
	for (int p = 0; p < t->world; p++) {
		const TeamBlob0* q = blob(p);
		if (q->total_kmers != m->total_kmers) return fail(KMX_EARG, "team build: rank %d opened another database (%llu records, %llu here)", p,
		                                                  (unsigned long long)q->total_kmers, (unsigned long long)m->total_kmers);
		for (int i = 0; i < 3; i++) cls[i] += q->cnt.class_count[i];
		bad += q->cnt.bad_count;
		if (p < t->rank) t->item_base += q->cnt.array_bound;
		items += q->cnt.array_bound;
	}
	if (bad)
		return fail(KMX_ERANGE, "%llu records have a count below ci=%d or above cs=%d: the reference indexes out of bounds there (kmodel.hpp:427, occu_bin.hpp:70)",
		            (unsigned long long)bad, m->ci, m->cs);
	uint64_t bf_kmers = 0;
	for (int i = 0; i < m->bf_num; i++) {
		m->kmer_counts[i] = cls[i];
		bf_kmers += cls[i];
	}
	m->km_kmers = m->total_kmers - bf_kmers;               // kmodel.hpp:433: header total, not the listed count
	int rc = check_model_sizes(m);
	if (rc) return rc;
	m->rest.k = m->k;
	m->rest.pre_len = rest_prefix_len(m->k);
	b.n_items = items;
	const uint64_t batch_items = (uint64_t)m->n_bits << kBucketLog;
	b.n_batches = (items + batch_items - 1) / batch_items;
	t->ml = m_layout(m);
	for (int p = 0; p < t->world; p++) t->xl[p] = x_layout(m, b.n_batches, p, t->n_active);
	const XLayout& X = t->xl[t->rank];
	m->mslab_bytes = t->ml.bytes;
	if ((rc = slab_acquire(&m->mslab, m->mslab_bytes, m->device))) return rc;
	if ((rc = slab_acquire(&t->xslab, X.bytes, m->device))) return rc;
	CU(cudaMemsetAsync(m->mslab, 0, m->mslab_bytes, s));   // filters, km_back, arrays and the barrier flags start at zero
	CU(cudaMemsetAsync((uint8_t*)t->xslab + X.ctl, 0, X.bytes - X.ctl, s));   // control block, round flags, prefix histogram
	uint8_t* mb = (uint8_t*)m->mslab;
	for (int i = 0; i < m->bf_num; i++) {
		m->d_bf[i] = (uint32_t*)(mb + t->ml.bf[i]);
		m->d_bf_back[i] = (uint32_t*)(mb + t->ml.bb[i]);
	}
	m->d_km_back = (uint32_t*)(mb + t->ml.kmback);
	for (int i = 0; i < m->n_bits; i++) m->d_cells[i] = (unsigned long long*)(mb + t->ml.cells[i]);
	fill_dev_model(m);
	CU(cudaStreamSynchronize(s));                         // zeroed before anybody maps the slabs
	out->device = m->device;
	out->ptr_m = (uint64_t)(uintptr_t)m->mslab;
	out->ptr_x = (uint64_t)(uintptr_t)t->xslab;
	CU(cudaIpcGetMemHandle(&out->hm, m->mslab));
	CU(cudaIpcGetMemHandle(&out->hx, t->xslab));
	TRACE(b.wall0, "team: slabs ready");
	return KMX_OK;
}

// ---- step 2: encode, Bloom merge, insert, km_back merge + array replication; sizes of the rest table -------------
static int team_step2(kmx_model* m, const TeamBlob1* in, size_t stride, TeamBlob2* out) {
	TeamState* t = m->bs.team;
	if (!t || !m->mslab) return fail(KMX_ESTATE, "team build: step 1 first");
	BuildState& b = m->bs;
	cudaStream_t s = m->x->stream;
	cudaEvent_t* ev = m->x->ev_build;
	CU(cudaSetDevice(m->device));
	int rc;
	for (int p = 0; p < t->world; p++) {
		const TeamBlob1* q = (const TeamBlob1*)((const uint8_t*)in + (size_t)p * stride);
		if (p == t->rank) {
			t->peer_m[p] = m->mslab;
			t->peer_x[p] = t->xslab;
			continue;
		}
		if ((rc = map_peer(m->device, q->pid, q->device, q->hm, q->ptr_m, &t->peer_m[p]))) return rc;
		if ((rc = map_peer(m->device, q->pid, q->device, q->hx, q->ptr_x, &t->peer_x[p]))) return rc;
	}
	const XLayout& X = t->xl[t->rank];
	const bool owner = t->rank < t->n_active;
	// the insert's arguments: exchange buffers in the slabs, private scratch from the pool
	InsertArgs& a = b.a;
	memset(&a, 0, sizeof(a));
	uint8_t* xb = (uint8_t*)t->xslab;
	a.ctl = (InsertCtl*)(xb + X.ctl);
	if (owner) {
		a.item_kmer = (const uint64_t*)(xb + X.item_k);
		a.item_occ = (const uint32_t*)(xb + X.item_o);
		for (int q = 0; q < 2; q++) {
			a.buf_kmer[q] = (uint64_t*)(xb + X.buf_k[q]);
			a.buf_occ[q] = (uint32_t*)(xb + X.buf_o[q]);
		}
		a.rest_kmer = (uint64_t*)(xb + X.rest_k);
		a.rest_occ = (uint32_t*)(xb + X.rest_o);
		for (int p = 0; p < t->n_active; p++) {
			uint8_t* pb = (uint8_t*)t->peer_x[p];
			const XLayout& P = t->xl[p];
			for (int q = 0; q < 2; q++) {
				a.peer_buf_kmer[q][p] = (uint64_t*)(pb + P.buf_k[q]);
				a.peer_buf_occ[q][p] = (uint32_t*)(pb + P.buf_o[q]);
			}
			a.peer_ctl[p] = (InsertCtl*)(pb + P.ctl);
			a.peer_flags[p] = (uint32_t*)(pb + P.flags);
		}
	}
	b.rest_cap = X.rest_cap;
	if ((rc = build_stage_insert_setup(m, t->rank, t->n_active, true))) return rc;

	// decode this rank's tiles: Bloom-bound records into its copy of the filters, array-bound ones into the owners' shards
	ItemRoute route;
	memset(&route, 0, sizeof(route));
	for (int o = 0; o < t->n_active; o++) {
		route.kmer[o] = (uint64_t*)((uint8_t*)t->peer_x[o] + t->xl[o].item_k);
		route.occ[o] = (uint32_t*)((uint8_t*)t->peer_x[o] + t->xl[o].item_o);
	}
	route.base = t->item_base;
	route.n_active = t->n_active;
	route.n_bits = m->n_bits;
	CU(cudaEventRecord(ev[1], s));
	CU(launch_encode(dev_db(t->db), m->dm, t->d_tile_off, route, t->tile_lo, t->tile_hi, m->sm_count, s));
	{
		OrReduceArgs r;
		memset(&r, 0, sizeof(r));
		r.link = team_link(m, t);
		r.n_vec = t->ml.bloom_bytes / 16;
		for (int p = 0; p < t->world; p++) r.base[p] = (uint4*)t->peer_m[p];
		CU(launch_or_allreduce(r, m->sm_count, s));           // its entry barrier: every rank's items are in their owners' shards
		t->seq += 2;
	}
	CU(cudaEventRecord(ev[2], s));
	if ((rc = build_stage_insert_run(m))) return rc;
	if (owner) CU(launch_prefix_hist(a.rest_kmer, &a.ctl->rest_n, 2 * (m->k - m->rest.pre_len), (uint32_t*)(xb + X.hist), m->sm_count, s));
	{
		OrReduceArgs r;
		memset(&r, 0, sizeof(r));
		r.link = team_link(m, t);
		r.n_vec = t->ml.kmback_bytes / 16;
		for (int p = 0; p < t->world; p++) r.base[p] = (uint4*)((uint8_t*)t->peer_m[p] + t->ml.kmback);
		for (int i = 0; i < m->n_bits; i++) {
			const int o = i % t->n_active;
			if (o == t->rank) continue;
			PullSeg& seg = r.pull[r.n_pull++];
			seg.src = (const uint4*)((uint8_t*)t->peer_m[o] + t->ml.cells[i]);
			seg.dst = (uint4*)((uint8_t*)m->mslab + t->ml.cells[i]);
			seg.n_vec = t->ml.cell_bytes / 16;
		}
		CU(launch_or_allreduce(r, m->sm_count, s));           // its entry barrier: every owner's insert is complete
		t->seq += 2;
	}
	CU(cudaEventRecord(ev[6], s));                        // everything after this is the rest table
	TRACE(b.wall0, "team: build queued");
	CU(cudaStreamSynchronize(s));
	TRACE(b.wall0, "team: arrays built and replicated");
	if (owner) t->ctl = m->x->h_pinned->ctl;
	unsigned int link_err = 0;
	CU(cudaMemcpy(&link_err, (uint8_t*)m->mslab + t->ml.flags + 128, 4, cudaMemcpyDeviceToHost));
	if (link_err) return fail(KMX_ECUDA, "team build: a peer GPU did not reach a barrier within 20 s");
	// every owner's survivor count and prefix histogram (read through the peer mappings)
	const int map_size = 1 << (2 * m->rest.pre_len);
	t->hist.assign(map_size, 0);
	std::vector<uint32_t> part(map_size);
	t->rest_total = 0;
	for (int o = 0; o < t->n_active; o++) {
		InsertCtl pc;
		CU(cudaMemcpy(&pc, (uint8_t*)t->peer_x[o] + t->xl[o].ctl, sizeof(pc), cudaMemcpyDeviceToHost));
		if (pc.error) return fail(KMX_ECUDA, "insert kernel of rank %d stopped with error %u (1: iteration cap, 2: survivor list overflow, 3: peer GPU timed out)", o, pc.error);
		t->owner_rest_n[o] = pc.rest_n;
		t->rest_total += pc.rest_n;
		CU(cudaMemcpy(part.data(), (uint8_t*)t->peer_x[o] + t->xl[o].hist, (size_t)map_size * 4, cudaMemcpyDeviceToHost));
		for (int p = 0; p < map_size; p++) t->hist[p] += part[p];
	}
	if (t->rest_total > 0x7FFFFFFFULL) return fail(KMX_ERANGE, "%llu rest entries overflow the reference's int indices (rest.hpp:66-70)", (unsigned long long)t->rest_total);
	{
		uint32_t cut[kMaxRanks + 1];
		uint64_t cut_off[kMaxRanks + 1];
		prefix_cuts(t->hist.data(), map_size, t->world, cut, cut_off);
		t->prefix_lo = cut[t->rank];
		t->prefix_hi = cut[t->rank + 1];
		t->my_off = cut_off[t->rank];
		t->my_n = cut_off[t->rank + 1] - cut_off[t->rank];
	}
	// the rest table every rank ends with (peers push their runs into it) + this rank's sort buffers
	const uint64_t n = t->rest_total;
	const size_t keys_bytes = up256((n + 1) * 8);
	m->rslab_bytes = keys_bytes + up256((n + 1) * 4);
	if ((rc = slab_acquire(&m->rslab, m->rslab_bytes, m->device))) return rc;
	m->d_rest_keys = (uint64_t*)m->rslab;
	m->d_rest_counts = (int32_t*)((uint8_t*)m->rslab + keys_bytes);
	DA(&t->d_gk, (t->my_n + 1) * 8, s);
	DA(&t->d_go, (t->my_n + 1) * 4, s);
	DA(&t->d_sk, (t->my_n + 1) * 8, s);
	DA(&t->d_so, (t->my_n + 1) * 4, s);
	DA(&t->d_gn, 8, s);
	const size_t temp = radix_sort_temp_bytes(t->my_n);
	if (temp) DA(&t->d_sort_temp, temp, s);
	CU(cudaStreamSynchronize(s));
	out->device = m->device;
	out->ptr_r = (uint64_t)(uintptr_t)m->rslab;
	CU(cudaIpcGetMemHandle(&out->hr, m->rslab));
	out->attempts = t->ctl.attempts;
	out->accepted = t->ctl.accepted;
	out->iterations = t->ctl.iterations;
	for (int i = 0; i < 12; i++) out->phase_cycles[i] = t->ctl.phase_cycles[i];
	CU(cudaEventElapsedTime(&out->ms_encode, ev[1], ev[2]));
	CU(cudaEventElapsedTime(&out->ms_insert, ev[5], ev[3]));
	return KMX_OK;
}

// ---- step 3: sharded rest table; the model is complete on every rank -----------------------------------------------
static int team_step3(kmx_model* m, const TeamBlob2* in, size_t stride) {
	TeamState* t = m->bs.team;
	if (!t || !m->rslab) return fail(KMX_ESTATE, "team build: step 2 first");
	BuildState& b = m->bs;
	cudaStream_t s = m->x->stream;
	cudaEvent_t* ev = m->x->ev_build;
	CU(cudaSetDevice(m->device));
	int rc;
	uint64_t attempts = 0, accepted = 0, iterations = 0, cycles[12] = { 0 };
	float ms_encode = 0, ms_insert = 0;
	for (int p = 0; p < t->world; p++) {
		const TeamBlob2* q = (const TeamBlob2*)((const uint8_t*)in + (size_t)p * stride);
		attempts += q->attempts;
		accepted += q->accepted;
		iterations = std::max<uint64_t>(iterations, q->iterations);
		for (int i = 0; i < 12; i++) cycles[i] = std::max<uint64_t>(cycles[i], q->phase_cycles[i]);
		ms_encode = std::max(ms_encode, q->ms_encode);
		ms_insert = std::max(ms_insert, q->ms_insert);
		if (p == t->rank) t->peer_r[p] = m->rslab;
		else if ((rc = map_peer(m->device, q->pid, q->device, q->hr, q->ptr_r, &t->peer_r[p]))) return rc;
	}
	RestHost& r = m->rest;
	r.k = m->k;
	r.pre_len = rest_prefix_len(m->k);
	r.map_size = 1 << (2 * r.pre_len);
	r.count = t->rest_total;
	r.suff_bin_size = r.count * (uint64_t)((m->k - r.pre_len) / 4);
	const int suffix_bits = 2 * (m->k - r.pre_len);
	// this rank's prefix range: gather from every owner's list, sort, push into every rank's table
	RestGatherArgs g;
	memset(&g, 0, sizeof(g));
	g.n_owners = t->n_active;
	for (int o = 0; o < t->n_active; o++) {
		g.kmer[o] = (const uint64_t*)((uint8_t*)t->peer_x[o] + t->xl[o].rest_k);
		g.occ[o] = (const uint32_t*)((uint8_t*)t->peer_x[o] + t->xl[o].rest_o);
		g.n[o] = t->owner_rest_n[o];
	}
	g.prefix_lo = t->prefix_lo;
	g.prefix_hi = t->prefix_hi;
	g.suffix_bits = suffix_bits;
	g.out_kmer = t->d_gk;
	g.out_occ = t->d_go;
	g.out_n = t->d_gn;
	g.cap = t->my_n;
	CU(launch_rest_gather(g, m->sm_count, s));
	unsigned long long gathered = 0;
	CU(cudaMemcpyAsync(&gathered, t->d_gn, 8, cudaMemcpyDeviceToHost, s));
	CU(launch_radix_sort_pairs(t->d_sort_temp, t->d_gk, t->d_sk, t->d_go, t->d_so, t->my_n, 2 * m->k, s));
	RestPushArgs pa;
	memset(&pa, 0, sizeof(pa));
	pa.link = team_link(m, t);
	pa.keys = t->d_sk;
	pa.counts = t->d_so;
	pa.n = t->my_n;
	pa.offset = t->my_off;
	const size_t keys_bytes = up256((t->rest_total + 1) * 8);
	for (int p = 0; p < t->world; p++) {
		pa.dst_keys[p] = (uint64_t*)t->peer_r[p];
		pa.dst_counts[p] = (int32_t*)((uint8_t*)t->peer_r[p] + keys_bytes);
	}
	CU(launch_rest_push(pa, m->sm_count, s));
	t->seq += 1;
	// group index (rest.hpp:95-105,115-126) straight from the global prefix histogram: dense ids of the non-empty prefixes,
	// first entry of each group, then the entry count
	std::vector<int32_t> h2i(r.map_size), pre;
	pre.reserve((size_t)r.map_size + 1);
	uint64_t acc = 0;
	for (int p = 0; p < r.map_size; p++) {
		if (t->hist[p]) {
			h2i[p] = (int32_t)pre.size();
			pre.push_back((int32_t)acc);
			acc += t->hist[p];
		} else {
			h2i[p] = -1;
		}
	}
	pre.push_back((int32_t)acc);
	r.pre_buffer_size = (int32_t)pre.size();                // rest.hpp:119: groups + 1
	DA(&m->d_hash2index, (size_t)r.map_size * 4, s);
	DA(&m->d_pre_buffer, ((size_t)r.map_size + 1) * 4, s);
	CU(cudaMemcpyAsync(m->d_hash2index, h2i.data(), h2i.size() * 4, cudaMemcpyHostToDevice, s));
	CU(cudaMemcpyAsync(m->d_pre_buffer, pre.data(), pre.size() * 4, cudaMemcpyHostToDevice, s));
	if ((rc = build_rest_side_tables(m))) return rc;
	CU(cudaEventRecord(ev[4], s));
	CU(cudaStreamSynchronize(s));                         // (h2i / pre are read by the copies above until here)
	TRACE(b.wall0, "team: rest table done");
	if (gathered != t->my_n)
		return fail(KMX_ECUDA, "team build: gathered %llu survivors for prefixes [%u, %u), the histogram says %llu", gathered, t->prefix_lo, t->prefix_hi,
		            (unsigned long long)t->my_n);
	unsigned int link_err = 0;
	CU(cudaMemcpy(&link_err, (uint8_t*)m->mslab + t->ml.flags + 128, 4, cudaMemcpyDeviceToHost));
	if (link_err) return fail(KMX_ECUDA, "team build: a peer GPU did not reach a barrier within 20 s");
	fill_dev_model(m);
	m->built = true;
	kmx_info_t& f = m->info;
	fill_info(m);
	f.insert_attempts = attempts;
	f.insert_accepted = accepted;
	f.insert_iterations = iterations;
	f.batches = b.n_batches;
	for (int i = 0; i < 12; i++) f.insert_phase_cycles[i] = cycles[i];
	f.ms_upload = b.ms_upload;
	f.ms_count = t->ms_count;
	f.ms_encode = ms_encode;
	f.ms_insert = ms_insert;
	float since_encode = 0;
	CU(cudaEventElapsedTime(&since_encode, ev[1], ev[4]));
	CU(cudaEventElapsedTime(&f.ms_rest, ev[6], ev[4]));     // includes the host side of the step-2 / step-3 exchange
	f.ms_total_device = t->ms_count + since_encode;
	f.build_time_cost = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - b.wall0).count();
	build_state_free(m);
	return KMX_OK;
}

}  // namespace kmx

// host-side pieces of the team build, exposed for tests (no GPU needed)
extern "C" void kmx_host_route(uint64_t g, int n_active, int n_bits, int32_t* owner, uint64_t* index) {
	int o = 0;
	unsigned long long at = g;
	if (n_active > 1) route_item(g, n_active, n_bits, &o, &at);
	*owner = o;
	*index = at;
}
extern "C" int kmx_host_prefix_cuts(const uint32_t* hist, int map_size, int world, uint32_t* cut, uint64_t* cut_off) {
	if (!hist || !cut || !cut_off || map_size < 1 || world < 1 || world > kMaxRanks) return fail(KMX_EARG, "kmx_host_prefix_cuts: bad argument");
	prefix_cuts(hist, map_size, world, cut, cut_off);
	return KMX_OK;
}

extern "C" int kmx_team_steps(void) { return kTeamSteps; }
extern "C" int kmx_team_blob_bytes(void) { return kTeamBlob; }

// One step of a team build on this rank.  blobs_in: the `world` blobs (256 bytes each, rank order) the ranks produced in the
// previous step (NULL for step 0); blob_out: this rank's 256 bytes for the next step.  Every rank must run every step even
// when its own previous step failed: the first word of a blob is the step's return code, and a step whose input holds a
// non-zero code fails on every rank without touching the device, so that nobody waits at a barrier for a rank that gave up.
extern "C" int kmx_team_step(kmx_model* m, kmx_db* db, int rank, int world, int step, const void* blobs_in, void* blob_out) {
	if (!m || !blob_out || step < 0 || step >= kTeamSteps || world < 1 || world > kMaxRanks || rank < 0 || rank >= world || (step == 0 && !db) ||
	    (step > 0 && !blobs_in)) {
		if (blob_out) {
			memset(blob_out, 0, kTeamBlob);
			*(int32_t*)blob_out = KMX_EARG;
		}
		return fail(KMX_EARG, "kmx_team_step: bad argument (step %d of %d, rank %d of %d)", step, kTeamSteps, rank, world);
	}
	memset(blob_out, 0, kTeamBlob);
	int rc = KMX_OK;
	if (step > 0) {
		for (int p = 0; p < world && !rc; p++) {
			const int32_t prc = *(const int32_t*)((const uint8_t*)blobs_in + (size_t)p * kTeamBlob);
			if (prc) rc = p == rank ? prc : fail(prc, "team build: rank %d failed in step %d (code %d)", p, step - 1, prc);
		}
	}
	if (!rc) {
		switch (step) {
		case 0: rc = team_step0(m, db, rank, world, (TeamBlob0*)blob_out); break;
		case 1: rc = team_step1(m, (const TeamBlob0*)blobs_in, kTeamBlob, (TeamBlob1*)blob_out); break;
		case 2: rc = team_step2(m, (const TeamBlob1*)blobs_in, kTeamBlob, (TeamBlob2*)blob_out); break;
		default: rc = team_step3(m, (const TeamBlob2*)blobs_in, kTeamBlob); break;
		}
	}
	((int32_t*)blob_out)[0] = rc;
	((int32_t*)blob_out)[1] = (int32_t)getpid();
	if (rc && m->x) {
		cudaSetDevice(m->device);
		build_state_free(m);
	}
	return rc;
}

// ---- the same build with one host thread per GPU inside this process ------------------------------------------------
namespace {
struct HostBarrier {
	std::mutex mu;
	std::condition_variable cv;
	int n, waiting = 0;
	unsigned long long phase = 0;
	explicit HostBarrier(int count) : n(count) {}
	void wait() {
		std::unique_lock<std::mutex> lock(mu);
		const unsigned long long my = phase;
		if (++waiting == n) {
			waiting = 0;
			phase++;
			cv.notify_all();
		} else {
			cv.wait(lock, [&] { return phase != my; });
		}
	}
};
}  // namespace

int kmx::team_build_in_process(kmx_model* m, const char* db_base, const std::vector<int>& devices) {
	const int world = (int)devices.size();
	if (m->built) return fail(KMX_ESTATE, "model already initialised (KModel::init is one-shot)");
	if (m->x && m->device != devices[0]) return fail(KMX_ESTATE, "the model already lives on device %d, the team starts on device %d", m->device, devices[0]);
	std::vector<kmx_model*> member(world, nullptr);
	member[0] = m;
	for (int r = 1; r < world; r++) {
		member[r] = kmx_create(m->ci, m->cs, m->n_hash, m->n_bits);
		if (!member[r]) {
			for (int q = 1; q < r; q++) kmx_destroy(member[q]);
			return last_error_code();
		}
	}
	std::vector<uint8_t> board[2] = { std::vector<uint8_t>((size_t)world * kTeamBlob), std::vector<uint8_t>((size_t)world * kTeamBlob) };
	std::vector<int> rcs(world, KMX_OK);
	std::vector<std::string> msgs(world);
	HostBarrier bar(world);
	auto run = [&](int r) {
		set_thread_device(devices[r]);
		kmx_db* db = kmx_db_open(db_base);
		int open_rc = db ? KMX_OK : (last_error_code() ? last_error_code() : KMX_EIO);
		if (open_rc) msgs[r] = last_error();
		for (int step = 0; step < kTeamSteps; step++) {
			uint8_t* mine = board[step & 1].data() + (size_t)r * kTeamBlob;
			int rc;
			if (step == 0 && open_rc) {
				memset(mine, 0, kTeamBlob);
				*(int32_t*)mine = open_rc;
				rc = open_rc;
			} else {
				rc = kmx_team_step(member[r], db, r, world, step, step ? board[(step - 1) & 1].data() : nullptr, mine);
				if (rc && msgs[r].empty()) msgs[r] = last_error();
			}
			if (rc && !rcs[r]) rcs[r] = rc;
			bar.wait();                                         // every rank's blob of this step is on the board
		}
		if (db) kmx_db_close(db);
		set_thread_device(-1);
	};
	std::vector<std::thread> pool;
	for (int r = 1; r < world; r++) pool.emplace_back(run, r);
	run(0);
	for (auto& th : pool) th.join();
	// the last step's codes travel in no blob exchange: look at every rank's own
	for (int r = 0; r < world; r++) {
		if (rcs[r]) {
			for (int q = 1; q < world; q++) kmx_destroy(member[q]);
			return fail(rcs[r], "%s", msgs[r].empty() ? "team build failed" : msgs[r].c_str());
		}
	}
	m->replicas.assign(member.begin() + 1, member.end());
	return KMX_OK;
}
