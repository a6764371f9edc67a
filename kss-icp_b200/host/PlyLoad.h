// PlyLoad.h -- minimal ASCII PLY vertex reader with the reference's CPLYLoader surface that the mains use
// (PlyLoad.h:21-42: LoadModel(char*), public `points`).  Own code; like the reference it reads x y z as
// float and widens to double (PlyLoad.cpp:93-101).  Unlike the reference it does not need an
// `element face` header line.  I/O is outside the GPU path.
#pragma once
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

class CPLYLoader {
public:
	std::vector<std::vector<double>> points;
	int LoadModel(char* filename) {
		points.clear();
		std::ifstream in(filename);
		if (!in) { std::printf("File can't be opened: %s\n", filename); return -1; }
		std::string line;
		long nv = -1;
		bool header_done = false;
		while (std::getline(in, line)) {
			if (line.compare(0, 14, "element vertex") == 0) nv = std::atol(line.c_str() + 14);
			if (line.compare(0, 10, "end_header") == 0) { header_done = true; break; }
		}
		if (!header_done || nv < 0) return -1;
		for (long i = 0; i < nv && std::getline(in, line); ++i) {
			float x, y, z;
			if (std::sscanf(line.c_str(), "%f %f %f", &x, &y, &z) != 3) { --i; continue; }
			std::vector<double> p(3);
			p[0] = x; p[1] = y; p[2] = z;
			points.push_back(p);
		}
		return 0;
	}
};
