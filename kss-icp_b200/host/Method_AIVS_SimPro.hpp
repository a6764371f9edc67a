// Method_AIVS_SimPro.hpp -- B200 drop-in for AIVS_Simplification_Pro (Method_AIVS_SimPro.hpp:44-154 of the
// reference), restricted to the two calls KSS-ICP makes: AIVS_Pro_init(br, name) and AIVS_simplification(n)
// (KSS_ICP.hpp:72-82).  The box grid, the coloured farthest point sampling and the greedy trim all run on the GPU
// behind kss_aivs_simplify (include/kss_icp_b200.h, kss-icp_b200/csrc/kss_aivs.cu); the kept points are the same
// input points in the same order as the reference's serial loop produces (tests/test_gpu_aivs.py).
#pragma once
#include <string>
#include <vector>

#include "kss_host.hpp"
#include "pointPipeline.hpp"

class AIVS_Simplification_Pro {
	std::vector<std::vector<double>> cloud;
public:
	std::vector<int> keptIndex;          // positions of the kept points in the input cloud (extra, for callers that want them)
	void AIVS_Pro_init(BallRegion br, std::string name) { (void)name; cloud = br.pointCloudData; }
	std::vector<std::vector<double>> AIVS_simplification(int pointNum) {
		const int N = (int)cloud.size();
		keptIndex.clear();
		if (N < 1 || pointNum < 1) return std::vector<std::vector<double>>();
		std::vector<double> in = kss_host::pack(cloud), out((size_t)N * 3);
		std::vector<int32_t> idx((size_t)N);
		int m = 0;
		if (!kss_host::ok(kss_aivs_simplify(kss_host::ctx(), in.data(), N, pointNum, out.data(), N, &m, idx.data()),
		                  "kss_aivs_simplify"))
			return std::vector<std::vector<double>>();
		out.resize((size_t)m * 3);
		keptIndex.assign(idx.begin(), idx.begin() + m);
		return kss_host::unpack(out);
	}
};
