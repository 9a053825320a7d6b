// kmcex_cli.cpp -- the `kmcEx` command line on top of include/kmodel.hpp (row N2 of SURVEY.md 8f).
// Same positional arguments, options, defaults and stdout lines as the reference driver
// (main.cpp:16-150).  The counting stage is the external `kmc` binary (main.cpp:136-140), which
// is not part of this repository: it is invoked when ./kmc_api/kmc exists; otherwise an existing
// <output_file_name>.kmc_pre/.kmc_suf database is used as it is, and if there is none the FASTQ
// input is counted on the GPU (kmx_count_fastq).
//
//   g++ -std=c++11 -O2 -Iinclude tools/kmcex_cli.cpp -Lkmcex_b200 -lkmx -Wl,-rpath,$PWD/kmcex_b200 -o kmcEx
#include <sys/stat.h>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <fstream>
#include "kmodel.hpp"

struct Options {
	int k = 31, num_hash = 7, num_bit = 5, ci = 1, cs = 1023, t = 4;      // main.cpp:16-27
	std::string input, output, workdir = "/tmp";
};

// the reference's read_me() (main.cpp:30-55): same sections, options, defaults and examples -- the command line is part of the
// interface; the note at the end is what differs here
static void usage() {
	const char* bar = "----------------------------------------------------------------------";
	std::cout << bar << "\n           kmcEx: counted k-mer encoding & decoding                   \n" << bar << "\n";
	std::cout << "VERSION: 1.5 (B200 build path, libkmx " << kmx_version() << ")\nDATE   : Nov 2nd, 2019\n" << bar << "\n\n";
	std::cout << "1. USAGE\n"
	             "     kmcEx [options] <input_file_name> <output_file_name> <working_directory>\n"
	             "     kmcEx [options] <@input_file_names> <output_file_name> <working_directory>\n"
	             "2. OPTIONS\n"
	             "     1) REQUIRED\n"
	             "        input_file_name    - single file in FASTQ format (gziped or not)\n"
	             "        @input_file_names  - file name with list of input files in FASTQ format (gziped or not)\n"
	             "        working_directory  - save temporary files\n"
	             "     2) OPTIONAL\n"
	             "        -k<len>            - k-mer length (default: 31)\n"
	             "        -t<value>          - total number of threads (default: 4)\n"
	             "        -ci<value>         - exclude k-mers occurring less than <value> times (default: 1)\n"
	             "        -cs<value>         - maximal value of a counter (default: 1023)\n"
	             "        -nh<value>         - number of hash (default: 7)\n"
	             "        -nb<value>         - number of bit array (default: 5)\n"
	             "3. EXAMPLES\n"
	             "     kmcEx -k31 -nh7 -nb5  rs.fastq rs.res /tmp\n"
	             "     kmcEx -k31 -nh7 -nb5  @rs.lst rs.res /tmp\n\n"
	             "   (the k-mer counting stage is ./kmc_api/kmc when that binary exists; otherwise an existing\n"
	             "    <output_file_name>.kmc_pre/.kmc_suf database is used, and if there is none plain-text FASTQ input\n"
	             "    is counted on the GPU; KMX_GPUS=<n> spreads the model build over n GPUs)\n\n";
}

// a path as ONE shell word: the kmc command line goes through system() as in main.cpp:136-140
static std::string shell_quote(const std::string& s) {
	std::string q = "'";
	for (char c : s) {
		if (c == '\'') q += "'\\''";
		else q += c;
	}
	return q + "'";
}

// mkdir -p without a shell (main.cpp:148 uses system("mkdir -p ..."))
static bool make_dirs(const std::string& path) {
	std::string cur;
	for (size_t i = 0; i <= path.size(); i++) {
		if (i == path.size() || path[i] == '/') {
			if (!cur.empty() && cur != "/" && mkdir(cur.c_str(), 0777) != 0 && errno != EEXIST) return false;
		}
		if (i < path.size()) cur += path[i];
	}
	struct stat st;
	return stat(path.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
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
		char opts[256];
		snprintf(opts, sizeof(opts), "./kmc_api/kmc -k%d -t%d -ci%d -cs%d ", o.k, o.t, o.ci, o.cs);
		const std::string cmd = std::string(opts) + shell_quote(o.input) + " " + shell_quote(o.output) + " " + shell_quote(o.workdir);
		std::cout << cmd << std::endl;
		if (system(cmd.c_str()) != 0) std::cout << "kmc returned a non-zero status" << std::endl;
		std::cout << std::endl;
	} else if (stat((o.output + ".kmc_pre").c_str(), &st) == 0) {
		std::cout << "./kmc_api/kmc not found: using the existing database " << o.output << std::endl;
		// a database left by an earlier run may have been counted with other options: say so instead of silently building from it
		if (kmx_db* db = kmx_db_open(o.output.c_str())) {
			kmx_db_info_t di;
			kmx_db_info(db, &di);
			kmx_db_close(db);
			if ((int)di.k != o.k || (int)di.min_count != o.ci || (int)di.max_count > o.cs)
				std::cout << "   WARNING: that database holds k=" << di.k << " -ci" << di.min_count << " -cs" << di.max_count << ", the command line asks for k=" << o.k
				          << " -ci" << o.ci << " -cs" << o.cs << std::endl;
		}
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
	if (!make_dirs(save_dir)) {                                          // main.cpp:148
		std::cout << "cannot create " << save_dir << std::endl;
		return 1;
	}
	kmodel->save(save_dir);
	return 0;
}
