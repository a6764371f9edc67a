// kss_large.cu -- placeholder until the hierarchical large-cloud path lands.
#include "kss_large.h"
namespace kss {
int large_nn_device(cudaStream_t, long long*, const double*, int, const double*, int, int*, float*, const DevAlloc&) { return KSS_ERR_UNSUPPORTED; }
int large_metrics_device(cudaStream_t, long long*, const double*, const int*, int, const double*, const int*, int, double*, const DevAlloc&) { return KSS_ERR_UNSUPPORTED; }
int large_icp_host(cudaStream_t, long long*, const double*, int, const double*, int, const kss_icp_params*, float*, double*, int*, int*, const DevAlloc&) { return KSS_ERR_UNSUPPORTED; }
}
