// Method_AIVS_SimPro.hpp -- STAND-IN for AIVS_Simplification_Pro (Method_AIVS_SimPro.hpp:44-154).
// Returns exactly min(n, |cloud|) points by a deterministic, order-preserving stride decimation.  It is NOT
// the reference's voxel-coloured farthest-point sampling (that component sits before the hot path; SURVEY.md
// 8 f1, DESIGN.md "out of scope"): use the reference's header in its own tree to get its point selection.
#pragma once
#include <string>
#include <vector>

#include "pointPipeline.hpp"

class AIVS_Simplification_Pro {
	std::vector<std::vector<double>> cloud;
public:
	void AIVS_Pro_init(BallRegion br, std::string name) { (void)name; cloud = br.pointCloudData; }
	std::vector<std::vector<double>> AIVS_simplification(int pointNum) {
		const size_t N = cloud.size();
		if (pointNum <= 0 || (size_t)pointNum >= N) return cloud;
		std::vector<std::vector<double>> out;
		out.reserve(pointNum);
		for (size_t k = 0; k < (size_t)pointNum; ++k) out.push_back(cloud[(k * N) / (size_t)pointNum]);
		return out;
	}
};
