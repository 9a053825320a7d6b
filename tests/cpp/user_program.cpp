// The reference's documented usage (README.md:64-93), compiled against include/kmodel.hpp of this
// repository and linked to libkmx.so.  argv: <kmc_db_base> <model_dir> <queries.txt> <out.txt> <ci>
#include <fstream>
#include "kmodel.hpp"

int main(int argc, char** argv) {
	if (argc < 6) return 2;
	std::string kmc_database = argv[1], model_dir = argv[2];
	int n_hash = 7, n_bit = 5, ci = atoi(argv[5]), cs = 1023;
	// 1) create a model, and save it to a disk
	KModel* kmodel = get_model(ci, cs, n_hash, n_bit);
	kmodel->init_KModel(kmc_database);
	kmodel->show_kmodel_info();
	kmodel->save_model(model_dir);
	// 2) load a model from a disk, and retrieve the occurrence of kmers
	KModel* loaded = get_model(model_dir);
	std::ifstream fin(argv[3]);
	std::vector<std::string> kmer_v;
	std::string s;
	while (fin >> s) kmer_v.push_back(s);
	std::vector<int> out = loaded->kmer_to_occ(kmer_v);
	std::ofstream fout(argv[4]);
	for (size_t i = 0; i < out.size(); i++) fout << out[i] << "\n";
	fout << "single " << loaded->kmer_to_occ(kmer_v[0]) << "\n";
	return 0;
}
