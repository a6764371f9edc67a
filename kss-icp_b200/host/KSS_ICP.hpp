// KSS_ICP.hpp -- B200 drop-in for the reference header of the same name (PS_AIS_Simplification/KSS_ICP.hpp:38-393).
// Same class KSSICP, same public methods and the public member pointAlign; every pcl::IterativeClosestPoint
// run, the similarity applies and the final 4x4 apply go through the C ABI (include/kss_icp_b200.h).
// The simplification step before the path (pointPipeline + AIVS_Simplification_Pro, KSS_ICP.hpp:71-81) goes through
// the headers of the same names in this directory, which forward to kss_aivs_simplify: the same points are kept.
#pragma once
#include <cfloat>
#include <fstream>
#include <iostream>

#include "pointPipeline.hpp"
#include "initRegistrationKSS.hpp"
#include "transferPC.hpp"
#include "Method_AIVS_SimPro.hpp"

using namespace std;

class KSSICP {

private:

	vector<vector<double>> pointSource;
	vector<vector<double>> pointTarget;
	int pNumber;
	double accurateG;

public:

	vector<vector<double>> pointAlign;

public:

	// KSS_ICP.hpp:53-67
	void KSSICP_init(vector<vector<double>> ps, vector<vector<double>> pt, double accurate) {
		accurateG = accurate;
		pointSource = ps;
		pointTarget = pt;
		if (pointSource.size() > pointTarget.size()) {
			pNumber = pointTarget.size();
		}
		else {
			pNumber = pointSource.size();
		}
		pNumber = pNumber / 2;
		if (pNumber > 2000) {
			pNumber = 2000;
		}
	}

	// KSS_ICP.hpp:69-131.  After the two simplifications the whole flow (MiddleAlign, sweep, judge ICP,
	// hypothesis ICPs, selection, similarity + 4x4 on the full source) is ONE kss_register call.
	void KSSICP_Registration(int iter) {

		pointPipeline ppt;
		ppt.pointPipeline_init_point_withoutUniform(pointTarget);
		AIVS_Simplification_Pro asp;
		asp.AIVS_Pro_init(ppt.br, "target");
		vector<vector<double>> pointCloudT = asp.AIVS_simplification(pNumber);

		pointPipeline pps;
		pps.pointPipeline_init_point_withoutUniform(pointSource);
		AIVS_Simplification_Pro asps;
		asps.AIVS_Pro_init(pps.br, "source");
		vector<vector<double>> pointCloudS = asps.AIVS_simplification(pNumber);

		std::vector<double> ss = kss_host::pack(pointCloudS), st = kss_host::pack(pointCloudT);
		std::vector<double> fs = kss_host::pack(pointSource), ft = kss_host::pack(pointTarget), pa(fs.size());
		kss_pair_result res;
		cout << "initRegistration start." << endl;
		if (!kss_host::ok(kss_register(kss_host::ctx(), ss.data(), (int)pointCloudS.size(), st.data(), (int)pointCloudT.size(),
			fs.data(), (int)pointSource.size(), ft.data(), (int)pointTarget.size(), accurateG, iter, &res, pa.data()), "kss_register"))
			return;
		lastResult = res;
		const int G = res.G;
		cout << "i:" << res.best_h / (G * G) << " j:" << (res.best_h / G) % G << " k:" << res.best_h % G << " (grid indices)" << endl;
		if (res.branch_multi) cout << "kernel" << res.winner << ":" << res.final_fitness << " (best of " << res.n_minima << ")" << endl;
		std::cout << "has converged: " << res.final_converged << std::endl;
		std::cout << "score: " << res.final_fitness << std::endl;
		for (int r = 0; r < 4; r++) std::cout << res.T[4 * r] << " " << res.T[4 * r + 1] << " " << res.T[4 * r + 2] << " " << res.T[4 * r + 3] << std::endl;

		// KSS_ICP.hpp:127-129: pointSource becomes the similarity-aligned full source
		double acc[64], lst[64];
		kss_sweep_angles(accurateG, acc, lst, 64);
		const double* tab = res.use_list ? lst : acc;
		const double ang[3] = { tab[res.used_h / (G * G)], tab[(res.used_h / G) % G], tab[res.used_h % G] };
		std::vector<double> sim(fs.size());
		if (kss_host::ok(kss_apply_similarity(kss_host::ctx(), fs.data(), (int)pointSource.size(), res.align, ang, sim.data()), "kss_apply_similarity"))
			pointSource = kss_host::unpack(sim);
		pointAlign = kss_host::unpack(pa);
	}

	// KSS_ICP.hpp:133-183: full-resolution ICP on the member clouds, pointAlign = PCL's transformed output (float)
	double shapeRegistration_ICP(int iter) {
		float T[16]; double fit = DBL_MAX;
		if (!run_icp(iter, pointSource, pointTarget, T, &fit)) return fit;
		print_icp(fit, T);
		pointAlign.clear();
		for (size_t i = 0; i < pointSource.size(); i++) {     // output = final * input in float
			const float x = (float)pointSource[i][0], y = (float)pointSource[i][1], z = (float)pointSource[i][2];
			vector<double> ppi;
			for (int r = 0; r < 3; r++) ppi.push_back((double)(((T[4 * r] * x + T[4 * r + 1] * y) + T[4 * r + 2] * z) + T[4 * r + 3]));
			pointAlign.push_back(ppi);
		}
		return fit;
	}

	// KSS_ICP.hpp:185-233: ICP on (ps, pt); the final 4x4 is applied to the MEMBER pointSource (B8)
	double shapeRegistration_ICP(int iter, vector<vector<double>> ps, vector<vector<double>> pt) {
		float T[16]; double fit = DBL_MAX;
		if (!run_icp(iter, ps, pt, T, &fit)) return fit;
		print_icp(fit, T);
		pointAlign.clear();
		if (!pointSource.empty()) {
			std::vector<double> in = kss_host::pack(pointSource), out(in.size());
			if (kss_host::ok(kss_apply_transform(kss_host::ctx(), T, in.data(), (int)pointSource.size(), out.data()), "kss_apply_transform"))
				pointAlign = kss_host::unpack(out);
		}
		return fit;
	}

	// KSS_ICP.hpp:236-274 (parameter Q is unused there too, B6)
	double shapeRegistration_ICP_AngleList(int iter, double Q, vector<vector<double>> ps, vector<vector<double>> pt) {
		(void)Q;
		float T[16]; double fit = DBL_MAX;
		run_icp(iter, ps, pt, T, &fit);
		return fit;
	}

	// KSS_ICP.hpp:276-321
	vector<vector<double>> shapeRegistration_ICP_AngleListV(int iter, double Q, vector<vector<double>> ps, vector<vector<double>> pt) {
		(void)Q;
		float T[16]; double fit = DBL_MAX;
		vector<vector<double>> registrationT;
		if (!run_icp(iter, ps, pt, T, &fit)) return registrationT;
		for (size_t i = 0; i < ps.size(); i++) {
			const float x = (float)ps[i][0], y = (float)ps[i][1], z = (float)ps[i][2];
			vector<double> ppi;
			for (int r = 0; r < 3; r++) ppi.push_back((double)(((T[4 * r] * x + T[4 * r + 1] * y) + T[4 * r + 2] * z) + T[4 * r + 3]));
			registrationT.push_back(ppi);
		}
		return registrationT;
	}

	// KSS_ICP.hpp:323-356
	double shapeRegistration_ICP_Judge(int iter, vector<vector<double>> ps, vector<vector<double>> pt) {
		float T[16]; double fit = DBL_MAX;
		run_icp(iter, ps, pt, T, &fit);
		return fit;
	}

	vector<vector<double>> IntrinsicICP_pointSource() {
		return pointSource;
	}
	vector<vector<double>> IntrinsicICP_pointTarget() {
		return pointTarget;
	}

	kss_pair_result lastResult = kss_pair_result();   // extra: everything the last KSSICP_Registration computed

private:

	// one pcl::IterativeClosestPoint run with the reference's fixed parameters (KSS_ICP.hpp:156-159)
	bool run_icp(int iter, const vector<vector<double>>& ps, const vector<vector<double>>& pt, float T[16], double* fit) {
		*fit = DBL_MAX;                                       // getFitnessScore() of an empty/failed run
		if (ps.empty() || pt.empty()) return false;
		std::vector<double> s = kss_host::pack(ps), t = kss_host::pack(pt);
		kss_icp_params prm;
		kss_icp_params_default(&prm);
		prm.max_iterations = iter;
		int iters = 0;
		lastConverged = 0;
		return kss_host::ok(kss_icp(kss_host::ctx(), s.data(), (int)ps.size(), t.data(), (int)pt.size(), &prm, T, fit, &iters, &lastConverged, nullptr), "kss_icp");
	}
	int lastConverged = 0;                                    // icp.hasConverged() of the last run_icp

	void print_icp(double fit, const float T[16]) {
		std::cout << "has converged: " << lastConverged << std::endl;
		std::cout << "score: " << fit << std::endl;
		for (int r = 0; r < 4; r++) std::cout << T[4 * r] << " " << T[4 * r + 1] << " " << T[4 * r + 2] << " " << T[4 * r + 3] << std::endl;
	}

	void save_PointCloud(vector<vector<double>> pointCloud, string Path) {
		ofstream fout(Path, ios::app);
		fout << pointCloud.size() << endl;
		for (size_t i = 0; i < pointCloud.size(); i++) {
			fout << pointCloud[i][0] << " " << pointCloud[i][1] << " " << pointCloud[i][2] << endl;
		}
		fout << endl;
		fout.close();
	}
};
