// xyzIO.hpp -- the count-prefixed text format the reference writes and ships:
//   save_PointCloud           Main_KSS_ICP.cpp:49-59   ("<n>\n" then "x y z\n" per point, default ostream precision,
//                                                       a trailing blank line, file opened in APPEND mode)
//   TransferPC_SavePC         transferPC.hpp:150-180   (the .wlop / .gird resamples under data/registration/)
// Own code.  Load_XYZ also accepts files without the count line (plain "x y z" rows, the .xyz/.txt flavour
// LoadPointCloud.hpp:45-49 reads) and takes the LAST block of an appended file when asked to.
#pragma once
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

inline bool kss_has_suffix(const std::string& s, const char* suf) {
	const std::string t(suf);
	return s.size() >= t.size() && s.compare(s.size() - t.size(), t.size(), t) == 0;
}

// returns the points of the first block (or of the whole file when it has no count line); empty on failure
inline std::vector<std::vector<double>> Load_XYZ(const std::string& path) {
	std::vector<std::vector<double>> pts;
	std::ifstream in(path);
	if (!in) return pts;
	std::string line;
	long want = -1;
	bool first = true;
	while (std::getline(in, line)) {
		std::istringstream ls(line);
		double a, b, c;
		if (first) {
			first = false;
			std::istringstream l1(line);
			double n; std::string rest;
			if ((l1 >> n) && !(l1 >> rest)) { want = (long)n; continue; }      // a single number: the count line
		}
		if (!(ls >> a >> b >> c)) { if (want >= 0 && (long)pts.size() >= want) break; continue; }
		std::vector<double> p(3);
		p[0] = a; p[1] = b; p[2] = c;
		pts.push_back(p);
		if (want >= 0 && (long)pts.size() >= want) break;
	}
	return pts;
}

// the reference's save_PointCloud, byte for byte (append mode included)
inline void Save_XYZ(const std::vector<std::vector<double>>& pointCloud, const std::string& path) {
	std::ofstream fout(path, std::ios::app);
	fout << pointCloud.size() << std::endl;
	for (size_t i = 0; i < pointCloud.size(); i++)
		fout << pointCloud[i][0] << " " << pointCloud[i][1] << " " << pointCloud[i][2] << std::endl;
	fout << std::endl;
	fout.close();
}
