// initRegistrationKSS.hpp -- B200 drop-in for the reference header of the same name
// (PS_AIS_Simplification/initRegistrationKSS.hpp:28-140): same class, public methods, parameter types and
// public data members; the bodies forward to the C ABI.  The public `kdtree` member (a PCL type no caller
// touches) is replaced by an opaque handle of the same name.
#pragma once
#include <algorithm>
#include <ctime>
#include <iostream>
#include <math.h>

#include "kss_host.hpp"

using namespace std;

class initRegistration_KSS {

private:

	double step;
	vector<vector<vector<double>>> value;
	int irange, jrange, krange;
	int r = 2; //kernel radius (initRegistrationKSS.hpp:35), fixed inside the device minima test

public:

	double x_middle_S;//target middle point
	double y_middle_S;
	double z_middle_S;
	double x_middle;//source to target middle vector
	double y_middle;
	double z_middle;
	double scale;//source transfer scale
	vector<double> angle;
	vector<vector<double>> angleList;
	vector<vector<double>> pointSource;
	vector<vector<double>> pointTarget;
	struct { void* opaque = nullptr; } kdtree;   // was pcl::KdTreeFLANN<pcl::PointXYZ>; the NN index lives on the GPU
	vector<double> rotationRecord;

public:

	// initRegistrationKSS.hpp:54-73
	void initRegistration_init(vector<vector<double>> pointinput, vector<vector<double>> pointinput2, double accurate) {
		step = accurate;
		cout << "initRegistration start." << endl;
		pointSource = pointinput;
		pointTarget = pointinput2;
		cout << "initRegistration middle align." << endl;
		initRegistration_MiddleAlign();
		cout << "initRegistration rotation." << endl;
		clock_t start = clock();
		initRegistration_Rotation();
		clock_t end = clock();
		cout << "alignment time cost:" << (double)(end - start) / CLOCKS_PER_SEC << "s" << endl;
	}

	// initRegistrationKSS.hpp:75-91
	vector<vector<double>> initRegistration_Rotation(vector<vector<double>> sourceOri) {
		return initRegistration_Rotation_Angle(sourceOri, angle);
	}

	// initRegistrationKSS.hpp:93-109
	vector<vector<double>> initRegistration_Rotation_Angle(vector<vector<double>> sourceOri, vector<double> angle_T) {
		if (sourceOri.empty()) return sourceOri;
		std::vector<double> in = kss_host::pack(sourceOri), out(in.size());
		const double a7[7] = { x_middle_S, y_middle_S, z_middle_S, x_middle, y_middle, z_middle, scale };
		const double ang[3] = { angle_T[0], angle_T[1], angle_T[2] };
		if (!kss_host::ok(kss_apply_similarity(kss_host::ctx(), in.data(), (int)sourceOri.size(), a7, ang, out.data()), "kss_apply_similarity"))
			return sourceOri;
		return kss_host::unpack(out);
	}

	// initRegistrationKSS.hpp:111-140 (never called by the reference; keeps its x_middle_S-on-all-axes quirk, B16)
	vector<vector<double>> initRegistration_Rotation_Axis(vector<vector<double>> sourceOri, int axis, double angleV) {
		if (axis < 1 || axis > 3) {
			cout << "error! illegal rotation" << endl;
			return sourceOri;
		}
		const double a7[7] = { 0, 0, 0, -x_middle_S, -x_middle_S, -x_middle_S, 1.0 };
		double ang[3] = { 0, 0, 0 };
		ang[axis - 1] = angleV;
		std::vector<double> in = kss_host::pack(sourceOri), out(in.size());
		if (sourceOri.empty() || !kss_host::ok(kss_apply_similarity(kss_host::ctx(), in.data(), (int)sourceOri.size(), a7, ang, out.data()), "kss_apply_similarity"))
			return sourceOri;
		for (size_t i = 0; i < out.size(); ++i) out[i] = out[i] + x_middle_S;
		return kss_host::unpack(out);
	}

	// score grid of the last sweep, value[i][j][k] (private in the reference; exposed read-only for tests)
	const vector<vector<vector<double>>>& initRegistration_Value() const { return value; }

private:

	// initRegistrationKSS.hpp:144-220
	void initRegistration_MiddleAlign() {
		std::vector<double> s = kss_host::pack(pointSource), t = kss_host::pack(pointTarget), al(s.size());
		double o7[7] = { 0, 0, 0, 0, 0, 0, 1 };
		if (!kss_host::ok(kss_middle_align(kss_host::ctx(), s.data(), (int)pointSource.size(), t.data(), (int)pointTarget.size(), o7, al.data()), "kss_middle_align"))
			return;
		x_middle_S = o7[0]; y_middle_S = o7[1]; z_middle_S = o7[2];
		x_middle = o7[3]; y_middle = o7[4]; z_middle = o7[5];
		scale = o7[6];
		pointSource = kss_host::unpack(al);
	}

	// initRegistrationKSS.hpp:222-296 (sweep, first-strict argmin, local minima -> angleList)
	void initRegistration_Rotation() {
		std::vector<double> s = kss_host::pack(pointSource), t = kss_host::pack(pointTarget);
		double acc[64], lst[64];
		const int G = kss_sweep_angles(step, acc, lst, 64);
		std::vector<double> val((size_t)G * G * G);
		std::vector<int> minima((size_t)G * G * G * 3);
		double best[3] = { 0, 0, 0 };
		int bi[3] = { 0, 0, 0 }, nmin = 0, Gout = 0;
		if (!kss_host::ok(kss_rotation_sweep(kss_host::ctx(), s.data(), (int)pointSource.size(), t.data(), (int)pointTarget.size(), step,
			KSS_SCORE_AVE, val.data(), &Gout, best, bi, minima.data(), &nmin), "kss_rotation_sweep"))
			return;
		value.assign(G, vector<vector<double>>(G, vector<double>(G)));
		for (int i = 0; i < G; i++) for (int j = 0; j < G; j++) for (int k = 0; k < G; k++) value[i][j][k] = val[((size_t)i * G + j) * G + k];
		irange = jrange = krange = G;
		for (int l = 0; l < nmin; l++) {
			vector<double> angleijk;
			angleijk.push_back((double)minima[3 * l] * 6.3 / (double)step);
			angleijk.push_back((double)minima[3 * l + 1] * 6.3 / (double)step);
			angleijk.push_back((double)minima[3 * l + 2] * 6.3 / (double)step);
			angleList.push_back(angleijk);
		}
		angle.push_back(best[0]);
		angle.push_back(best[1]);
		angle.push_back(best[2]);
		cout << "i:" << best[0] << "j:" << best[1] << "k:" << best[2] << endl;
		cout << endl;
	}
};
