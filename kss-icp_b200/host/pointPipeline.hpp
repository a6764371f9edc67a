// pointPipeline.hpp -- B200 drop-in for the part of the reference's front end KSS-ICP uses:
// pointPipeline::pointPipeline_init_point_withoutUniform(cloud) (pointPipeline.hpp:88-101) and the `br` member it
// hands to AIVS_Pro_init.  The border scan and the BallRegion box grid (ballRegionCompute.hpp) are rebuilt on the GPU
// inside kss_aivs_simplify, so `br` only carries the cloud.  Loading, normal estimation and the other pointPipeline
// entry points are outside the registration path (out of scope).
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
