// kmcex_cli.cpp -- the `kmcEx` command line on top of include/kmodel.hpp (row N2 of SURVEY.md 8f).
// Same positional arguments, options, defaults and stdout lines as the reference driver
// (main.cpp:16-150).  The counting stage is the external `kmc` binary (main.cpp:136-140), which
// is not part of this repository: it is invoked when ./kmc_api/kmc exists; otherwise an existing
// <output_file_name>.kmc_pre/.kmc_suf database is used as it is, and if there is none the FASTQ
// input is counted on the GPU (kmx_count_fastq).
//
//   g++ -std=c++11 -O2 -Iinclude tools/kmcex_cli.cpp -Lkmcex_b200 -lkmx -Wl,-rpath,$PWD/kmcex_b200 -o kmcEx
#include <sys/stat.h>
#include <cstdio>
#include <cstring>
#include <fstream>
#include "kmodel.hpp"

struct Options {
	int k = 31, num_hash = 7, num_bit = 5, ci = 1, cs = 1023, t = 4;      // main.cpp:16-27
	std::string input, output, workdir = "/tmp";
};

static void usage() {
	std::cout << "kmcEx (B200 build path) [options] <input_file_name> <output_file_name> <working_directory>\n"
	             "  -k<len> k-mer length (31)   -t<value> threads for the kmc stage (4)\n"
	             "  -ci<value> minimum count (1)   -cs<value> counter ceiling (1023)\n"
	             "  -nh<value> number of hash functions (7)   -nb<value> number of coupled bit arrays (5)\n"
	             "  input_file_name may be a FASTQ file or @list, as for the reference; when ./kmc_api/kmc is absent\n"
	             "  the database <output_file_name>.kmc_pre/.kmc_suf must already exist\n";
}

static bool parse(int argc, char** argv, Options& o) {
	if (argc < 4) return false;
	int i = 1;
	for (; i < argc; ++i) {
		const char* a = argv[i];
		if (a[0] != '-') break;
		if (!strncmp(a, "-nh", 3)) o.num_hash = atoi(a + 3);
		else if (!strncmp(a, "-nb", 3)) o.num_bit = atoi(a + 3);
		else if (!strncmp(a, "-ci", 3)) o.ci = atoi(a + 3);
		else if (!strncmp(a, "-cs", 3)) o.cs = atoi(a + 3);
		else if (!strncmp(a, "-t", 2)) o.t = atoi(a + 2);
		else if (!strncmp(a, "-k", 2)) o.k = atoi(a + 2);
	}
	if (argc - i < 3) return false;
	o.input = argv[argc - 3];
	o.output = argv[argc - 2];
	o.workdir = argv[argc - 1];
	return !o.input.empty() && !o.output.empty() && !o.workdir.empty();
}

int main(int argc, char** argv) {
	Options o;
	if (!parse(argc, argv, o)) {
		usage();
		return 255;
	}
	struct stat st;
	if (stat("./kmc_api/kmc", &st) == 0) {
		char cmd[2048];
		snprintf(cmd, sizeof(cmd), "./kmc_api/kmc -k%d -t%d -ci%d -cs%d %s %s %s", o.k, o.t, o.ci, o.cs, o.input.c_str(), o.output.c_str(),
		         o.workdir.c_str());
		std::cout << cmd << std::endl;
		if (system(cmd) != 0) std::cout << "kmc returned a non-zero status" << std::endl;
		std::cout << std::endl;
	} else if (stat((o.output + ".kmc_pre").c_str(), &st) == 0) {
		std::cout << "./kmc_api/kmc not found: using the existing database " << o.output << std::endl;
	} else {
		// the counting stage on the GPU (kmx_count_fastq): plain-text FASTQ, or @file listing one path per line
		std::vector<std::string> files;
		if (!o.input.empty() && o.input[0] == '@') {
			std::ifstream lst(o.input.substr(1));
			std::string line;
			while (std::getline(lst, line))
				if (!line.empty()) files.push_back(line);
		} else {
			files.push_back(o.input);
		}
		std::vector<const char*> paths;
		for (auto& f : files)
			if (stat(f.c_str(), &st) == 0) paths.push_back(f.c_str());
		if (paths.size() != files.size() || paths.empty()) {
			// like a failed kmc run in the reference: carry on, KModel::init reports the missing database
			std::cout << "input file(s) not found, no k-mer database produced" << std::endl;
			paths.clear();
		}
		std::cout << "./kmc_api/kmc not found: counting " << files.size() << " FASTQ file(s) on the GPU" << std::endl;
		kmx_count_info_t ci = {};
		if (!paths.empty() && kmx_count_fastq(paths.data(), (int)paths.size(), o.k, o.ci, o.cs, o.output.c_str(), &ci) != KMX_OK) {
			std::cout << kmx_last_error() << std::endl;
			return 1;
		}
		if (!paths.empty()) std::cout << "   reads " << ci.n_reads << ", k-mers " << ci.n_windows << ", unique " << ci.n_unique << ", kept " << ci.n_kept << std::endl;
	}
	KModel* kmodel = get_model(o.ci, o.cs, o.num_hash, o.num_bit);
	kmodel->init(o.output);
	kmodel->show_kmodel_info();
	size_t slash = o.output.find_last_of('/');
	std::string save_dir = o.workdir + "/" + (slash == std::string::npos ? o.output : o.output.substr(slash + 1));
	if (system(("mkdir -p " + save_dir).c_str()) != 0) return 1;        // main.cpp:148
	kmodel->save(save_dir);
	return 0;
}
