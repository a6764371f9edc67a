// Main_KSS_List_b200.cpp -- batch driver: the loop body of the reference's Main_KSS_List.cpp:133-166
// (load pair, KSSICP_init, KSSICP_Registration, PCR_QM, record the time) for a list of pairs, but handing the
// whole list to ONE kss_register_batch call so the GPU sees every pair at once.  The reference file itself is
// an empty translation unit as shipped (its body is inside a comment and its list is empty), so it cannot be
// "driven unchanged"; this is the equivalent driver.
//
// usage: Main_KSS_List_b200 [--gpus N] [--hyp-shard] <list.txt> [step=8] [iter=1000] [out_dir]
//   list.txt: one "<source> <target>" per line; .ply (ASCII) or the count-prefixed .xyz/.wlop/.gird text format;
//   out_dir: write <k>Align.xyz per pair like Main_KSS_List.cpp:138-141
//   --gpus N    : one context per GPU, host threads inside kss_register_batch_multi, pairs in contiguous blocks
//   --hyp-shard : instead, EVERY GPU takes every pair and the rotation hypotheses are sharded (NCCL inside the library)
#include <chrono>
#include <fstream>
#include <iostream>
#include <sstream>
#include <thread>

#include "KSS_ICP.hpp"
#include "registrationMeasure.hpp"
#include "xyzIO.hpp"

static std::vector<std::vector<double>> Load_PLY(const std::string& f) {
	if (!kss_has_suffix(f, ".ply")) return Load_XYZ(f);          // .xyz / .txt / .wlop / .gird

	CPLYLoader l;
	std::vector<char> p(f.begin(), f.end()); p.push_back(0);
	l.LoadModel(p.data());
	return l.points;
}

int main(int argc, char** argv) {
	int gpus = 1; bool hyp_shard = false;
	while (argc > 1 && argv[1][0] == '-' && argv[1][1] == '-') {
		if (std::string(argv[1]) == "--gpus" && argc > 2) { gpus = std::atoi(argv[2]); argv += 2; argc -= 2; }
		else if (std::string(argv[1]) == "--hyp-shard") { hyp_shard = true; ++argv; --argc; }
		else break;
	}
	if (argc < 2 || gpus < 1) { std::cerr << "usage: Main_KSS_List_b200 [--gpus N] [--hyp-shard] <list.txt> [step] [iter] [out_dir]\n"; return 2; }
	const double step = argc > 2 ? std::atof(argv[2]) : 8.0;
	const int iter = argc > 3 ? std::atoi(argv[3]) : 1000;
	std::ifstream lf(argv[1]);
	std::string line;
	std::vector<std::pair<std::string, std::string>> names;
	while (std::getline(lf, line)) {
		std::istringstream ls(line);
		std::string a, b;
		if (ls >> a >> b) names.push_back({a, b});
	}
	// text parsing is the slow part once the GPU does the rest: files are read by a few host threads
	std::vector<kss_host::Cloud> S0(names.size()), T0(names.size());
	{
		const int nt = (int)std::min<size_t>(std::max(1u, std::thread::hardware_concurrency()), std::max<size_t>(1, names.size()));
		std::vector<std::thread> th;
		for (int w = 0; w < nt; ++w)
			th.emplace_back([&, w]() { for (size_t i = w; i < names.size(); i += nt) { S0[i] = Load_PLY(names[i].first); T0[i] = Load_PLY(names[i].second); } });
		for (auto& t : th) t.join();
	}
	std::vector<kss_host::Cloud> S, T;
	for (size_t i = 0; i < names.size(); ++i) {
		if (S0[i].empty() || T0[i].empty()) { std::cerr << "skip " << names[i].first << " " << names[i].second << "\n"; continue; }
		S.push_back(std::move(S0[i])); T.push_back(std::move(T0[i]));
	}
	const int P = (int)S.size();
	if (P == 0) { std::cerr << "no pairs\n"; return 1; }
	// raw clouds: sim_s = sim_t = NULL makes the library run the pNumber rule and both AIVS simplifications
	// (KSS_ICP.hpp:53-82) on the device, in the same batch
	size_t cS = 0, cT = 0;
	for (int p = 0; p < P; ++p) { cS = std::max(cS, S[p].size()); cT = std::max(cT, T[p].size()); }
	std::vector<double> bS(P * cS * 3), bT(P * cT * 3);
	std::vector<int> nS(P), nT(P);
	auto put = [](std::vector<double>& dst, size_t cap, int p, const kss_host::Cloud& c) {
		for (size_t i = 0; i < c.size(); ++i) for (int a = 0; a < 3; ++a) dst[((size_t)p * cap + i) * 3 + a] = c[i][a];
	};
	for (int p = 0; p < P; ++p) { put(bS, cS, p, S[p]); put(bT, cT, p, T[p]); nS[p] = (int)S[p].size(); nT[p] = (int)T[p].size(); }
	kss_batch b;
	kss_batch_default(&b);
	b.n_pairs = P; b.cap_S = (int)cS; b.cap_T = (int)cT;
	b.full_s = bS.data(); b.full_t = bT.data();
	b.cnt_S = nS.data(); b.cnt_T = nT.data();
	b.step = step; b.icp.max_iterations = iter;
	std::vector<kss_pair_result> res(P);
	const std::string out_dir = argc > 4 ? argv[4] : "";
	std::vector<double> aligned(out_dir.empty() ? 0 : (size_t)P * cS * 3);
	double* pa = out_dir.empty() ? nullptr : aligned.data();
	std::vector<kss_ctx*> ctxs(gpus, nullptr);
	if (gpus > 1) {
		for (int g = 0; g < gpus; ++g)
			if (kss_ctx_create(g, &ctxs[g]) != KSS_OK) { std::cerr << "kss_ctx_create(" << g << ") failed: " << gpus << " GPUs are needed\n"; return 1; }
		if (hyp_shard && kss_ctx_nccl_init_all(ctxs.data(), gpus) != KSS_OK) { std::cerr << "NCCL: " << kss_last_error(ctxs[0]) << "\n"; return 1; }
	}
	auto t0 = std::chrono::steady_clock::now();
	if (gpus == 1) {
		if (!kss_host::ok(kss_register_batch(kss_host::ctx(), &b, res.data(), pa), "kss_register_batch")) return 1;
	} else if (!hyp_shard) {
		const int rc = kss_register_batch_multi(ctxs.data(), gpus, &b, res.data(), pa);
		if (rc != KSS_OK) { std::cerr << "kss_register_batch_multi failed (" << rc << "): " << kss_last_error(ctxs[0]) << "\n"; return 1; }
	} else {
		// every rank takes the whole batch; rank 0's buffers are the output, the others' identical results are dropped
		std::vector<int> rc(gpus, 0);
		std::vector<std::vector<kss_pair_result>> rr(gpus, std::vector<kss_pair_result>(P));
		std::vector<std::thread> th;
		for (int g = 0; g < gpus; ++g)
			th.emplace_back([&, g]() { rc[g] = kss_register_batch_hyp_sharded(ctxs[g], &b, g ? rr[g].data() : res.data(), g ? nullptr : pa); });
		for (auto& t : th) t.join();
		for (int g = 0; g < gpus; ++g) if (rc[g] != KSS_OK) { std::cerr << "rank " << g << " failed (" << rc[g] << "): " << kss_last_error(ctxs[g]) << "\n"; return 1; }
	}
	const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	for (auto c : ctxs) if (c) kss_ctx_destroy(c);
	for (int p = 0; p < P && !out_dir.empty(); ++p) {                     // Main_KSS_List.cpp:138-141: the aligned source per pair
		kss_host::Cloud c(S[p].size(), std::vector<double>(3));
		for (size_t i = 0; i < c.size(); ++i) for (int a = 0; a < 3; ++a) c[i][a] = aligned[((size_t)p * cS + i) * 3 + a];
		Save_XYZ(c, out_dir + "/" + std::to_string(p) + "Align.xyz");
	}
	for (int p = 0; p < P; ++p)
		std::cout << "pair " << p << " MSE: " << res[p].mse << " RMSE: " << res[p].rmse << " MAE: " << res[p].mae
		          << " fitness: " << res[p].final_fitness << " hypotheses: " << res[p].n_minima << " winner: " << res[p].winner << "\n";
	std::cout << P << " registrations in " << sec << " s (" << P / sec << " registrations/s, host buffers in and out, " << gpus
	          << (gpus > 1 ? (hyp_shard ? " GPUs, hypotheses sharded" : " GPUs, pairs sharded") : " GPU") << ")\n";
	return 0;
}
