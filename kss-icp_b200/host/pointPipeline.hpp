// pointPipeline.hpp -- STAND-IN for the reference's pre-processing front end (pointPipeline.hpp:88-101 +
// ballRegionCompute.hpp).  The voxel-grid / AIVS simplification runs BEFORE the registration hot path and is
// outside this round's scope; this header only keeps the call surface KSS_ICP.hpp uses
// (`pointPipeline::pointPipeline_init_point_withoutUniform(cloud)` and the `br` member handed to AIVS).
// In the reference's own tree, keep its pointPipeline.hpp instead.
#pragma once
#include <vector>

#include "PlyLoad.h"

struct BallRegion {
	std::vector<std::vector<double>> pointCloudData;   // same member name as ballRegionCompute.hpp
};

class pointPipeline {
public:
	BallRegion br;
	void pointPipeline_init_point_withoutUniform(std::vector<std::vector<double>> pointCloud) {
		br.pointCloudData = pointCloud;
	}
};
