// KSS_ICP_cli.cpp -- the command line of the released binary (EXE/Readme.txt: "KSS-ICP.exe PointSource.ply
// PointTarget.ply", "The result is output into a .xyz format file"), on top of the drop-in classes of this directory:
// the same calls as Main_KSS_ICP.cpp:79-88 (KSSICP_init(S, T, 8), KSSICP_Registration(1000), PCR_QM), clouds from
// ASCII .ply or the count-prefixed text formats, the aligned source written with the reference's save_PointCloud.
//
// usage: KSS_ICP_cli <source> <target> [result.xyz = Registration.xyz] [step = 8] [iter = 1000]
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

#include "KSS_ICP.hpp"
#include "registrationMeasure.hpp"
#include "xyzIO.hpp"

static std::vector<std::vector<double>> load_cloud(const std::string& f) {
	if (!kss_has_suffix(f, ".ply")) return Load_XYZ(f);
	CPLYLoader l;
	std::vector<char> p(f.begin(), f.end()); p.push_back(0);
	l.LoadModel(p.data());
	return l.points;
}

int main(int argc, char** argv) {
	if (argc < 3) { std::cerr << "usage: " << argv[0] << " <source> <target> [result.xyz] [step] [iter]\n"; return 2; }
	const std::string out = argc > 3 ? argv[3] : "Registration.xyz";
	const double step = argc > 4 ? std::atof(argv[4]) : 8.0;
	const int iter = argc > 5 ? std::atoi(argv[5]) : 1000;
	std::vector<std::vector<double>> pointSource = load_cloud(argv[1]), pointTarget = load_cloud(argv[2]);
	if (pointSource.empty() || pointTarget.empty()) { std::cerr << "could not read the clouds\n"; return 1; }
	KSSICP ki;
	ki.KSSICP_init(pointSource, pointTarget, step);
	ki.KSSICP_Registration(iter);
	std::vector<std::vector<double>> pointAlign = ki.pointAlign;
	PCR_QM pq;
	pq.PCR_QM_init(pointAlign, pointTarget);              // runs PCR_QM_Start (registrationMeasure.hpp:19-23)
	std::vector<double> m = pq.PCR_QM_ReturnResult();
	std::cout << "Registration Measure" << ":" << "MSE: " << m[0] << " RMSE: " << m[1] << " MAE: " << m[2] << std::endl;
	std::remove(out.c_str());                           // save_PointCloud appends; the CLI starts from a fresh file
	Save_XYZ(pointAlign, out);
	return 0;
}
