// A user program written against the random-access part of the reference's kmc_api (CKMCFile / CKmerAPI), compiled against
// include/kmc_ra.hpp + libkmx.so by tests/test_gpu_cpp_api.py.
//   ra_program <db_base> <kmers.txt> <reads.txt> <out.txt>
// out.txt: one line per k-mer ("1 <count>" / "0 0"), then one line per read (its counters, or "-" when GetCountersForRead
// says false), then the same k-mers again after SetMinCount(3), then the batch forms.
#include <fstream>
#include <iostream>
#include "kmc_ra.hpp"

int main(int argc, char** argv) {
	if (argc < 5) return 2;
	CKMCFile db;
	if (!db.OpenForRA(argv[1])) {
		std::cout << "cannot open " << argv[1] << ": " << kmx_last_error() << std::endl;
		return 1;
	}
	std::ifstream kin(argv[2]), rin(argv[3]);
	std::ofstream out(argv[4]);
	std::vector<std::string> kmers, reads;
	std::string line;
	while (std::getline(kin, line)) kmers.push_back(line);
	while (std::getline(rin, line)) reads.push_back(line);
	out << "k " << db.KmerLength() << " total " << db.KmerCount() << " min " << db.GetMinCount() << " max " << db.GetMaxCount() << " both " << db.GetBothStrands() << "\n";
	CKmerAPI kmer(db.KmerLength());
	for (auto& s : kmers) {
		uint32 c = 0;
		const bool ok = kmer.from_string(s) && db.CheckKmer(kmer, c);
		out << (ok ? 1 : 0) << " " << (ok ? c : 0) << "\n";
	}
	for (auto& r : reads) {
		std::vector<uint32> counters;
		if (!db.GetCountersForRead(r, counters)) {
			out << "-\n";
			continue;
		}
		for (size_t i = 0; i < counters.size(); i++) out << (i ? " " : "") << counters[i];
		out << "\n";
	}
	const bool set_ok = db.SetMinCount(3);                 // separate statements: operands of one << chain have no fixed order before C++17
	const uint32 now_min = db.GetMinCount();
	const bool too_low = db.SetMinCount(0);
	out << "setmin " << set_ok << " " << now_min << " toolow " << too_low << "\n";
	for (auto& s : kmers) {
		uint32 c = 0;
		const bool ok = kmer.from_string(s) && db.CheckKmer(kmer, c);
		out << (ok ? c : 0) << "\n";
	}
	db.ResetMinMaxCounts();
	std::vector<uint64_t> packed;
	for (auto& s : kmers) {
		kmer.from_string(s);
		packed.push_back(kmer.packed());
	}
	std::vector<uint32> counts;
	std::vector<std::vector<uint32> > per_read;
	if (!db.CheckKmers(packed, counts) || !db.GetCountersForReads(reads, per_read)) return 3;
	for (size_t i = 0; i < counts.size(); i++) out << counts[i] << "\n";
	for (auto& c : per_read) {
		out << c.size();
		for (auto x : c) out << " " << x;
		out << "\n";
	}
	db.Close();
	return 0;
}
