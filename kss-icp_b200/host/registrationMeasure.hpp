// registrationMeasure.hpp -- B200 drop-in for PCR_QM (PS_AIS_Simplification/registrationMeasure.hpp:21-99)
#pragma once
#include <iostream>

#include "kss_host.hpp"

using namespace std;

class PCR_QM {

private:

	vector<vector<double>> a;
	vector<vector<double>> t;
	vector<double> MSERA;

public:

	void PCR_QM_init(vector<vector<double>> alignV, vector<vector<double>> templateV) {
		a = alignV;
		t = templateV;
		PCR_QM_Start();
	}

	vector<double> PCR_QM_ReturnResult() {
		return MSERA;
	}

private:

	void PCR_QM_Start() {
		std::vector<double> fa = kss_host::pack(a), ft = kss_host::pack(t);
		double m[3] = { 0, 0, 0 };
		if (!a.empty() && !t.empty())
			kss_host::ok(kss_nn_metrics(kss_host::ctx(), fa.data(), (int)a.size(), ft.data(), (int)t.size(), m), "kss_nn_metrics");
		std::cout << "Result:" << endl;
		std::cout << "MSE:  " << m[0] << endl;
		std::cout << "RMSE: " << m[1] << endl;
		std::cout << "MAE:  " << m[2] << endl;
		MSERA.clear();
		MSERA.push_back(m[0]);
		MSERA.push_back(m[1]);
		MSERA.push_back(m[2]);
	}
};
