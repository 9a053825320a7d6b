// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
//
// A thin command-line driver around the UNMODIFIED reference (lzhLab/kmcEx), compiled
// from the sources where they lie under /root/reference (see oracle/Makefile). It is
// the pin for the CPU restatement in oracle/kmx_oracle.cpp and the "reference" CPU
// baseline for bench.py. Nothing in the product path links or executes it.
//
// The reference API it drives:
//   get_model(ci,cs,nh,nb)      kmodel.hpp:674     KModel::init      kmodel.hpp:57
//   get_model(dir)              kmodel.hpp:680     KModel::save      kmodel.hpp:173
//   KModel::kmer_to_occ(vector) kmodel.hpp:90      CKMCFile listing  kmc_file.cpp:66,428
//
// Sub-commands
//   build  <db_base> <out_dir> <ci> <cs> <nh> <nb>     init + save, prints timings (JSON)
//   query  <model_dir> <packed_u64.bin> <k> <out_i32.bin> <threads>
//   query_ascii <model_dir> <strings.bin> <k> <out_i32.bin> <threads>    n strings of k raw bytes each (N, lower case allowed)
//   list   <db_base> <out.bin>         (u64 kmer, u32 count) records in listing order
//   check  <db_base> <packed_u64.bin> <out_u32.bin>     CKMCFile::OpenForRA + CheckKmer per k-mer (0 when it says false)
//   reads  <db_base> <reads.txt> <out_u32.bin>          CKMCFile::GetCountersForRead per line, counters back to back
//   kat    <out.txt>                   known-answer values for hash / canonical / OccuBin / KMC signatures
#include "kmodel.hpp"
#include <chrono>

static double now_s() {
	return chrono::duration<double>(chrono::high_resolution_clock::now().time_since_epoch()).count();
}

static int cmd_build(int argc, char** argv) {
	if (argc < 8) return 2;
	string db = argv[2], out = argv[3];
	int ci = atoi(argv[4]), cs = atoi(argv[5]), nh = atoi(argv[6]), nb = atoi(argv[7]);
	double t0 = now_s();
	KModel* m = get_model(ci, cs, nh, nb);
	m->init(db);
	double t1 = now_s();
	m->show_kmodel_info();
	m->save(out);
	double t2 = now_s();
	printf("{\"init_s\": %.6f, \"save_s\": %.6f}\n", t1 - t0, t2 - t1);
	return 0;
}

static int cmd_query(int argc, char** argv) {
	if (argc < 7) return 2;
	string dir = argv[2];
	int k = atoi(argv[4]);
	int threads = atoi(argv[6]);
	FILE* f = fopen(argv[3], "rb");
	if (!f) { printf("cannot open %s\n", argv[3]); return 1; }
	fseek(f, 0, SEEK_END);
	size_t n = ftell(f) / 8;
	fseek(f, 0, SEEK_SET);
	vector<uint64_t> packed(n);
	if (n && fread(packed.data(), 8, n, f) != n) return 1;
	fclose(f);
	vector<string> kmers(n);
	for (size_t i = 0; i < n; i++) kmers[i] = Tools::uint64_to_string(packed[i], k);
	double t0 = now_s();
	KModel* m = get_model(dir);
	double t1 = now_s();
	vector<int> occ = m->kmer_to_occ(kmers, threads);
	double t2 = now_s();
	FILE* fo = fopen(argv[5], "wb");
	if (n) fwrite(occ.data(), 4, n, fo);
	fclose(fo);
	printf("{\"load_s\": %.6f, \"query_s\": %.6f, \"n\": %zu, \"threads\": %d}\n", t1 - t0, t2 - t1, n, threads);
	return 0;
}

static int cmd_query_ascii(int argc, char** argv) {
	if (argc < 7) return 2;
	string dir = argv[2];
	int k = atoi(argv[4]);
	int threads = atoi(argv[6]);
	FILE* f = fopen(argv[3], "rb");
	if (!f) { printf("cannot open %s\n", argv[3]); return 1; }
	fseek(f, 0, SEEK_END);
	size_t n = ftell(f) / k;
	fseek(f, 0, SEEK_SET);
	vector<char> flat(n * k);
	if (n && fread(flat.data(), k, n, f) != n) return 1;
	fclose(f);
	vector<string> kmers(n);
	for (size_t i = 0; i < n; i++) kmers[i] = string(flat.data() + i * k, k);
	KModel* m = get_model(dir);
	vector<int> occ = m->kmer_to_occ(kmers, threads);
	FILE* fo = fopen(argv[5], "wb");
	if (n) fwrite(occ.data(), 4, n, fo);
	fclose(fo);
	printf("{\"n\": %zu}\n", n);
	return 0;
}

static int cmd_list(int argc, char** argv) {
	if (argc < 4) return 2;
	CKMCFile db;
	if (!db.OpenForListing(argv[2])) { printf("cannot open db %s\n", argv[2]); return 1; }
	CKmerAPI kmer(db.KmerLength());
	uint32 c;
	FILE* fo = fopen(argv[3], "wb");
	uint64_t n = 0;
	while (db.ReadNextKmer(kmer, c)) {
		uint64_t v = Tools::kmers2uint64(kmer.to_string());
		fwrite(&v, 8, 1, fo);
		fwrite(&c, 4, 1, fo);
		n++;
	}
	fclose(fo);
	printf("{\"listed\": %llu, \"total\": %llu, \"k\": %u}\n", (unsigned long long)n,
		(unsigned long long)db.KmerCount(), db.KmerLength());
	return 0;
}

// random access: kmc_file.cpp:27-58 (OpenForRA), :320-356 (CheckKmer), :879-897 (GetCountersForRead)
static int cmd_check(int argc, char** argv) {
	if (argc < 5) return 2;
	CKMCFile db;
	if (!db.OpenForRA(argv[2])) { printf("cannot open db %s\n", argv[2]); return 1; }
	FILE* f = fopen(argv[3], "rb");
	if (!f) { printf("cannot open %s\n", argv[3]); return 1; }
	fseek(f, 0, SEEK_END);
	size_t n = ftell(f) / 8;
	fseek(f, 0, SEEK_SET);
	vector<uint64_t> packed(n);
	if (n && fread(packed.data(), 8, n, f) != n) return 1;
	fclose(f);
	const int k = (int)db.KmerLength();
	vector<uint32_t> out(n, 0);
	CKmerAPI kmer(k);
	size_t hits = 0;
	for (size_t i = 0; i < n; i++) {
		kmer.from_string(Tools::uint64_to_string(packed[i], k));
		uint32 c = 0;
		if (db.CheckKmer(kmer, c)) { out[i] = c; hits++; }
	}
	FILE* fo = fopen(argv[4], "wb");
	if (n) fwrite(out.data(), 4, n, fo);
	fclose(fo);
	printf("{\"n\": %zu, \"hits\": %zu}\n", n, hits);
	return 0;
}

static int cmd_reads(int argc, char** argv) {
	if (argc < 5) return 2;
	CKMCFile db;
	if (!db.OpenForRA(argv[2])) { printf("cannot open db %s\n", argv[2]); return 1; }
	ifstream in(argv[3]);
	FILE* fo = fopen(argv[4], "wb");
	string line;
	size_t n_reads = 0, n_counters = 0;
	while (getline(in, line)) {
		vector<uint32_t> counters;
		if (db.GetCountersForRead(line, counters) && !counters.empty()) {
			fwrite(counters.data(), 4, counters.size(), fo);
			n_counters += counters.size();
		}
		n_reads++;
	}
	fclose(fo);
	printf("{\"reads\": %zu, \"counters\": %zu}\n", n_reads, n_counters);
	return 0;
}

static int cmd_kat(int argc, char** argv) {
	if (argc < 3) return 2;
	FILE* fo = fopen(argv[2], "w");
	// murmur over a deterministic family of strings, lengths 1..32, several seeds
	const char* alpha = "ACGT";
	uint64_t x = 88172645463325252ULL;
	for (int len = 1; len <= 32; len++) {
		for (int rep = 0; rep < 4; rep++) {
			string s(len, 'A');
			for (int i = 0; i < len; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; s[i] = alpha[x & 3]; }
			for (int si = 0; si < 128; si += 37) {
				fprintf(fo, "murmur %s %u %llu\n", s.c_str(), HashSeeds[si],
					(unsigned long long)Tools::murmur_hash64(s.c_str(), len, HashSeeds[si]));
			}
			fprintf(fo, "minkmer %s %s\n", s.c_str(), Tools::get_min_kmer(s).c_str());
			for (int sl = 5; sl <= 11; sl += 2) {             // CKmerAPI::get_signature, kmer_api.h:653-673
				if (sl > len) break;
				CKmerAPI ka(len);
				ka.from_string(s);
				fprintf(fo, "signature %s %d %u\n", s.c_str(), sl, ka.get_signature(sl));
			}
		}
	}
	{
		// signatures of strings rich in the excluded patterns (AA inside, TT* / TGT endings, ACA beginning)
		const char* pat[] = { "AAAAAAAAAAAAAAAAAAAAAAAAAAAAAAA", "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTT", "ACAACAACAACAACAACAACAACAACAACAA", "TGTTGTTGTTGTTGTTGTTGTTGTTGTTGTT",
		                      "AACCAACCAACCAACCAACCAACCAACCAAC", "ACGTTTACGTTTACGTTTACGTTTACGTTTA", "CAAGCAAGCAAGCAAGCAAGCAAGCAAGCAA", "GATTACAGATTACAGATTACAGATTACAGAT" };
		for (const char* p : pat) {
			string s(p);
			for (int sl = 5; sl <= 11; sl++) {
				CKmerAPI ka((uint32)s.size());
				ka.from_string(s);
				fprintf(fo, "signature %s %d %u\n", s.c_str(), sl, ka.get_signature(sl));
			}
		}
	}
	int cfgs[4][2] = { {1024, 7}, {256, 7}, {1024, 6}, {65536, 8} };
	for (auto& c : cfgs) {
		OccuBin ob(c[0], c[1]);
		for (int occ = 0; occ < c[0]; occ++) {
			int bin = ob.occ_to_bin(occ);
			fprintf(fo, "occubin %d %d %d %d %u\n", c[0], c[1], occ, bin, ob.bin_to_mean((uint32_t)bin));
		}
	}
	fclose(fo);
	return 0;
}

int main(int argc, char** argv) {
	if (argc < 2) { printf("usage: ref_driver build|query|query_ascii|list|check|reads|kat ...\n"); return 2; }
	string c = argv[1];
	if (c == "build") return cmd_build(argc, argv);
	if (c == "query") return cmd_query(argc, argv);
	if (c == "query_ascii") return cmd_query_ascii(argc, argv);
	if (c == "list") return cmd_list(argc, argv);
	if (c == "check") return cmd_check(argc, argv);
	if (c == "reads") return cmd_reads(argc, argv);
	if (c == "kat") return cmd_kat(argc, argv);
	return 2;
}
