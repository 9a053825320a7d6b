// kmodel.hpp -- drop-in replacement of the reference's kmodel.hpp for the model build and
// retrieval path.  Same free functions, class and method names, argument meaning and error
// behaviour as lzhLab/kmcEx (kmodel.hpp:39-235, 674-696; README.md:64-93), but every method
// forwards to the C ABI of libkmx.so (include/kmx.h), whose kernels run on a B200.
//
//     #include "kmodel.hpp"
//     KModel* kmodel = get_model(ci, cs, n_hash, n_bit);   // kmodel.hpp:674
//     kmodel->init_KModel(kmc_database);                   // README.md:76  (= init, kmodel.hpp:57)
//     kmodel->save_model(model_dir);                       // README.md:78  (= save, kmodel.hpp:173)
//     KModel* q = get_model(model_dir);                    // kmodel.hpp:680
//     int occ = q->kmer_to_occ(kmer);                      // kmodel.hpp:100
//     vector<int> out = q->kmer_to_occ(kmer_v);            // kmodel.hpp:90
//
// Build:  g++ -std=c++11 -Iinclude user.cpp -Lkmcex_b200 -lkmx -Wl,-rpath,<dir of libkmx.so>
// Unlike the reference header this one may be included from several translation units.
#pragma once
#ifndef KMODEL_H
#define KMODEL_H

#include <stdint.h>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <vector>
#include "kmx.h"

class KModel {
public:
	explicit KModel(kmx_model* handle) : h_(handle) {}

	// kmodel.hpp:57-86.  db_file is the KMC base name (".kmc_pre"/".kmc_suf" are appended).
	void init(std::string db_file) {
		if (kmx_init_from_kmc(h_, db_file.c_str()) != KMX_OK) die();
		show_header_info();                                  // the reference prints it from inside init (kmodel.hpp:66)
	}
	void init_KModel(std::string db_file) { init(db_file); }    // README.md:76

	// kmodel.hpp:173-206.  save_dir must exist (README.md:77).
	void save(std::string save_dir) {
		if (kmx_save(h_, save_dir.c_str()) != KMX_OK) die();
	}
	void save_model(std::string save_dir) { save(save_dir); }   // README.md:78

	// kmodel.hpp:209-235
	void load(std::string save_dir) {
		kmx_model* fresh = kmx_load(save_dir.c_str());
		if (!fresh) die();
		kmx_destroy(h_);
		h_ = fresh;
	}

	// kmodel.hpp:90-98.  The batch is flattened into one buffer (t_num host threads) and answered
	// by one pipelined GPU call; strings whose length is not the model's k answer 0.
	std::vector<int> kmer_to_occ(std::vector<std::string> kmer_v, int t_num = 4) {
		const size_t n = kmer_v.size();
		std::vector<int> occ_v(n);
		if (n == 0) return occ_v;
		kmx_info_t info;
		kmx_info(h_, &info);
		const size_t k = (size_t)info.k;
		std::vector<char> flat(n * k);
		std::vector<char> bad(n, 0);
		const size_t n_thr = n < 65536 ? 1 : (size_t)(t_num < 1 ? 1 : (t_num > 64 ? 64 : t_num));
		auto pack = [&](size_t lo, size_t hi) {
			for (size_t i = lo; i < hi; i++) {
				if (kmer_v[i].size() == k) memcpy(&flat[i * k], kmer_v[i].data(), k);
				else { memset(&flat[i * k], 'A', k); bad[i] = 1; }
			}
		};
		if (n_thr == 1) {
			pack(0, n);
		} else {
			std::vector<std::thread> pool;
			for (size_t t = 0; t < n_thr; t++) pool.emplace_back(pack, n * t / n_thr, n * (t + 1) / n_thr);
			for (auto& th : pool) th.join();
		}
		static_assert(sizeof(int) == sizeof(int32_t), "int is 32 bits");
		if (kmx_query_ascii(h_, flat.data(), k, n, reinterpret_cast<int32_t*>(occ_v.data())) != KMX_OK) die();
		for (size_t i = 0; i < n; i++)
			if (bad[i]) occ_v[i] = 0;
		return occ_v;
	}

	// kmodel.hpp:100-116 (r_occ is unused by the reference as well)
	int kmer_to_occ(std::string kmer, uint32_t r_occ = 0) {
		(void)r_occ;
		kmx_info_t info;
		kmx_info(h_, &info);
		if ((int)kmer.size() != info.k) return 0;
		int32_t occ = 0;
		if (kmx_query_ascii(h_, kmer.data(), kmer.size(), 1, &occ) != KMX_OK) die();
		return occ;
	}

	// additive: packed 2-bit k-mers in, no string marshalling
	std::vector<int> kmer_to_occ_packed(const std::vector<uint64_t>& kmers) {
		std::vector<int> occ_v(kmers.size());
		if (kmx_query_packed(h_, kmers.data(), kmers.size(), reinterpret_cast<int32_t*>(occ_v.data())) != KMX_OK) die();
		return occ_v;
	}

	// kmodel.hpp:118-125
	void show_header_info() {
		kmx_info_t i;
		kmx_info(h_, &i);
		std::cout << "KMCEX:" << std::endl;
		std::cout << "   kmodel number hash                 :     " << i.n_hash << std::endl;
		std::cout << "   kmodel bit array                   :     " << i.n_bits << std::endl;
		std::cout << "   total kmercount                    :     " << i.total_kmers << std::endl;
		std::cout << "   kmercount in blommfilter           :     " << i.bf_kmers << std::endl;
		std::cout << "   kmercount in kmodel                :     " << i.km_kmers << std::endl;
	}

	// kmodel.hpp:127-146
	void show_kmodel_info() {
		kmx_info_t i;
		kmx_info(h_, &i);
		const uint64_t mb = 1024 * 1024;
		const uint64_t total = i.bf_bytes + i.km_bytes + i.rest_bytes + i.km_back_bytes;
		std::cout << "   kmercount hash map                 :     " << i.rest_kmers << std::endl;
		std::cout << "   memory bloomfilter                 :     " << i.bf_bytes / mb << "MB" << std::endl;
		std::cout << "   memory bit array                   :     " << i.km_bytes / mb << "MB" << std::endl;
		std::cout << "   memory rest map                    :     " << i.rest_bytes / mb << "MB" << std::endl;
		std::cout << "   total memory                       :     " << total / mb << "MB" << std::endl;
		std::cout << "   build time cost                    :     " << i.build_time_cost << std::endl;
	}

	kmx_model* handle() { return h_; }

private:
	kmx_model* h_;

	// the reference's error convention: a message on stdout, then exit(1) (kmodel.hpp:394-397,682-685)
	static void die() {
		std::cout << kmx_last_error() << std::endl;
		exit(1);
	}
};

// kmodel.hpp:674-677
inline KModel* get_model(int ci = 1, int cs = 1023, int num_hash = 7, int num_bit = 5) {
	kmx_model* h = kmx_create(ci, cs, num_hash, num_bit);
	if (!h) {
		std::cout << kmx_last_error() << std::endl;
		exit(1);
	}
	return new KModel(h);
}

// kmodel.hpp:680-696
inline KModel* get_model(std::string save_dir) {
	kmx_model* h = kmx_load(save_dir.c_str());
	if (!h) {
		std::cout << kmx_last_error() << std::endl;
		exit(1);
	}
	return new KModel(h);
}

#endif
