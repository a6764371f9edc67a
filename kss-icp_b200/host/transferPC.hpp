// transferPC.hpp -- the reference's test-pair generator (transferPC.hpp:33-182, class TransferPC): two resamplings of one
// model, the second ("gird") rotated about a coordinate axis / scaled about its centroid / translated, both written in
// the count-prefixed text format (.wlop / .gird, append mode).  This is how data/registration/* and
// data/registration_scale/* were made (transfer.txt holds the axis and angle per model).
//
// Same class, method names, argument meaning and arithmetic as the reference.  The one difference is the resampling
// itself: the reference calls CGAL (WLOP to 8000 points, grid simplification with cell radius / 1.5;
// transferPC.hpp:134-141, Method_CGAL.hpp), which is not available here.  TransferPC_init therefore takes the model
// from a point file and resamples it with two OWN stand-ins -- an evenly strided subset (for WLOP) and a
// one-point-per-voxel grid simplification (for Grid) -- and TransferPC_init_points takes two resamplings made elsewhere.
#pragma once
#include <cmath>
#include <fstream>
#include <iostream>
#include <map>
#include <string>
#include <vector>

#include "xyzIO.hpp"

class TransferPC {

public:

	std::string filewlop;
	std::string filegird;

private:

	std::vector<std::vector<double>> pointResampleWlop;
	std::vector<std::vector<double>> pointResampleGird;

public:

	// transferPC.hpp:52-64, with the stand-in resamplers (see the header comment); wlopNumber as in :137
	void TransferPC_init(std::string filePath, int wlopNumber = 8000, double gridCell = 0.0) {
		std::vector<std::vector<double>> pts = Load_XYZ(filePath);
		const size_t index = filePath.find_last_of(".");
		filewlop = filePath.substr(0, index) + ".wlop";
		filegird = filePath.substr(0, index) + ".gird";
		TransferPC_Resample(pts, wlopNumber, gridCell);
	}

	void TransferPC_init_points(const std::vector<std::vector<double>>& wlop, const std::vector<std::vector<double>>& gird,
		std::string fileStem = "") {
		pointResampleWlop = wlop;
		pointResampleGird = gird;
		filewlop = fileStem.empty() ? "" : fileStem + ".wlop";
		filegird = fileStem.empty() ? "" : fileStem + ".gird";
	}

	// transferPC.hpp:66-98 -- cord 1: x, 2: y, 3 (anything else): z; the same axis convention as
	// initRegistration_Transfer (initRegistrationKSS.hpp:365-404); cos / sin evaluated per point as there
	void TransferPC_Transfer(int cord, double angle) {
		for (size_t i = 0; i < pointResampleGird.size(); i++) {
			std::vector<double>& p = pointResampleGird[i];
			double xi, yi, zi;
			if (cord == 1) {
				xi = p[0];
				yi = p[1] * cos(angle) - p[2] * sin(angle);
				zi = p[1] * sin(angle) + p[2] * cos(angle);
			}
			else if (cord == 2) {
				xi = p[2] * sin(angle) + p[0] * cos(angle);
				yi = p[1];
				zi = p[2] * cos(angle) - p[0] * sin(angle);
			}
			else {
				xi = p[0] * cos(angle) - p[1] * sin(angle);
				yi = p[0] * sin(angle) + p[1] * cos(angle);
				zi = p[2];
			}
			p[0] = xi; p[1] = yi; p[2] = zi;
		}
	}

	// transferPC.hpp:100-121 -- scale about the centroid (serial sums in index order)
	void TransferPC_Scale(double rate) {
		double x_sum = 0, y_sum = 0, z_sum = 0;
		for (size_t i = 0; i < pointResampleGird.size(); i++) {
			x_sum = x_sum + pointResampleGird[i][0];
			y_sum = y_sum + pointResampleGird[i][1];
			z_sum = z_sum + pointResampleGird[i][2];
		}
		x_sum = x_sum / pointResampleGird.size();
		y_sum = y_sum / pointResampleGird.size();
		z_sum = z_sum / pointResampleGird.size();
		for (size_t i = 0; i < pointResampleGird.size(); i++) {
			pointResampleGird[i][0] = (pointResampleGird[i][0] - x_sum) * rate + x_sum;
			pointResampleGird[i][1] = (pointResampleGird[i][1] - y_sum) * rate + y_sum;
			pointResampleGird[i][2] = (pointResampleGird[i][2] - z_sum) * rate + z_sum;
		}
	}

	// transferPC.hpp:123-130 -- the same offset on all three coordinates
	void TransferPC_Translate(double dis) {
		for (size_t i = 0; i < pointResampleGird.size(); i++) {
			pointResampleGird[i][0] = pointResampleGird[i][0] + dis;
			pointResampleGird[i][1] = pointResampleGird[i][1] + dis;
			pointResampleGird[i][2] = pointResampleGird[i][2] + dis;
		}
	}

	// transferPC.hpp:132-141 -- {wlop, gird}; writes both files first (append mode, like the reference)
	std::vector<std::vector<std::vector<double>>> TransferPC_ReturnPoints() {
		TransferPC_SavePC();
		std::vector<std::vector<std::vector<double>>> result;
		result.push_back(pointResampleWlop);
		result.push_back(pointResampleGird);
		return result;
	}

private:

	// stand-ins for simplification_Method_CGAL_WLOP(8000) and _Grid(radius / 1.5) (transferPC.hpp:144-151)
	void TransferPC_Resample(const std::vector<std::vector<double>>& pts, int wlopNumber, double gridCell) {
		pointResampleWlop.clear(); pointResampleGird.clear();
		if (pts.empty()) return;
		const size_t n = pts.size(), m = wlopNumber > 0 && (size_t)wlopNumber < n ? (size_t)wlopNumber : n;
		for (size_t k = 0; k < m; ++k) pointResampleWlop.push_back(pts[(k * n) / m]);        // evenly strided subset
		double lo[3] = { pts[0][0], pts[0][1], pts[0][2] }, hi[3] = { pts[0][0], pts[0][1], pts[0][2] };
		for (size_t i = 0; i < n; ++i) for (int a = 0; a < 3; ++a) { lo[a] = std::min(lo[a], pts[i][a]); hi[a] = std::max(hi[a], pts[i][a]); }
		if (!(gridCell > 0.0)) gridCell = std::max(hi[0] - lo[0], std::max(hi[1] - lo[1], hi[2] - lo[2])) / 64.0;
		std::map<long long, size_t> seen;                                                  // first point of every voxel, in input order
		for (size_t i = 0; i < n; ++i) {
			const long long kx = (long long)((pts[i][0] - lo[0]) / gridCell), ky = (long long)((pts[i][1] - lo[1]) / gridCell),
				kz = (long long)((pts[i][2] - lo[2]) / gridCell);
			const long long key = (kz * 2097152ll + ky) * 2097152ll + kx;
			if (seen.insert(std::make_pair(key, i)).second) pointResampleGird.push_back(pts[i]);
		}
	}

	// transferPC.hpp:153-180
	void TransferPC_SavePC() {
		if (filewlop.size() <= 2) std::cout << "normal file name is empty!" << std::endl;
		else Save_XYZ(pointResampleWlop, filewlop);
		if (filegird.size() <= 2) std::cout << "normal file name is empty!" << std::endl;
		else Save_XYZ(pointResampleGird, filegird);
	}

};
