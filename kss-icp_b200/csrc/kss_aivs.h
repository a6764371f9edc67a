// kss_aivs.h -- batched AIVS simplification on the device (kss_aivs.cu), the step in front of the registration path
// (KSSICP_Registration, KSS_ICP.hpp:53-84: pNumber = min(|S|, |T|) / 2 capped at 2000, then one
// AIVS_simplification(pNumber) per cloud).
#pragma once
#include <cuda_runtime.h>
#include "kss_large.h"   // DevAlloc

namespace kss {

// clouds [P][cap][3] doubles (d_cnt null: every cloud has cap points); the target count is d_point_num[p] (or
// point_num_all when null).  Outputs: d_out [P][out_cap][3], d_out_cnt [P], optional d_out_idx [P][out_cap] (positions
// in the input cloud).  *d_bad (an int the caller cleared) becomes 1 for a degenerate cloud (zero extent or too many
// boxes), 2 when more than 16384 samples would have to be trimmed, 3 when out_cap is too small.
int aivs_simplify_device(cudaStream_t st, long long* launches, int P, const double* d_pts, const int* d_cnt, int cap,
                         const int* d_point_num, int point_num_all, double* d_out, int out_cap, int* d_out_cnt,
                         int* d_out_idx, int* d_bad, const DevAlloc& alloc, const char* tag);

// pn[p] = min(|S_p|, |T_p|) / 2, at most 2000 (KSS_ICP.hpp:53-67)
int aivs_pnumber_device(cudaStream_t st, long long* launches, int P, const int* cnt_S, int cap_S, const int* cnt_T,
                        int cap_T, int* pn);

}  // namespace kss
