// kss_large.h -- clouds beyond one CTA's shared memory (> 2048 points): block-grid NN with TMA-staged
// target tiles (box pyramid as exact fallback), canonical-order reductions and the full-resolution ICP loop
// (shapeRegistration_ICP(int iter), KSS_ICP.hpp:133-183; PCR_QM at full resolution).
// Nothing here synchronises with the host except large_icp_run's convergence poll and large_icp_result.
#pragma once
#include <cuda_runtime.h>
#include <functional>
#include "../../include/kss_icp_b200.h"

namespace kss {

// named grow-only device allocation supplied by the context
typedef std::function<int(const char* name, size_t bytes, void** out)> DevAlloc;

// exact 1-NN of n_q double queries (narrowed RN) against n_t targets; device pointers
int large_nn_device(cudaStream_t st, long long* launches, const double* d_q, int n_q, const double* d_t, int n_t,
                    int* d_idx, float* d_d2, const DevAlloc& alloc);

// PCR_QM (registrationMeasure.hpp:47-98) for one pair of n_q / n_t points (device pointers, host counts)
int large_metrics_device(cudaStream_t st, long long* launches, const double* d_q, int n_q, const double* d_t, int n_t,
                         double* d_out3, const DevAlloc& alloc);

// a prepared full-resolution ICP run (device buffers owned by the context's allocator)
struct LargeIcp {
    alignas(8) unsigned char pyramid[160];
    alignas(8) unsigned char grid[72];
    int n_s = 0, n_t = 0, nchunks = 0;
    void *inp = nullptr, *cur = nullptr, *tg = nullptr, *tg2 = nullptr, *cert = nullptr, *state = nullptr;   // inp: sorted input; cur / tg: by original index
    int *partK = nullptr, *perm = nullptr, *inv = nullptr, *worklist = nullptr;
    unsigned char *flagS = nullptr, *dirty = nullptr;
    int* h_unres = nullptr;             // pinned host word (mapped): left-over count of an earlier iteration
    float *d2 = nullptr, *partA = nullptr, *partB = nullptr, *l2f = nullptr;     // partials [quantity][S], level-2 results [9][S2]
    double *partD = nullptr, *l2d = nullptr, *out3 = nullptr;
    int* l2k = nullptr; unsigned* grpcnt = nullptr;
    int S = 0, S2 = 0;                  // padded chunk / 256-chunk-group counts
    // optional stage marks (CUDA-event timing by the context): mark(user, KSS_STAGE_*, begin?1:0)
    void (*mark)(void* user, int stage, int begin) = nullptr;
    void* mark_user = nullptr;
};
// `prefix` names the run's buffers: a prepared run that outlives the call (kss_icp_large_begin / _iterate / _end) gets
// its own, so that other large-path calls on the same context cannot regrow them underneath it
int large_icp_prepare(cudaStream_t st, long long* launches, const double* d_s, int n_s, const double* d_t, int n_t,
                      const DevAlloc& alloc, LargeIcp* run, const char* prefix);
int large_icp_iterations(cudaStream_t st, long long* launches, LargeIcp* run, const kss_icp_params* prm, int count);
int large_icp_run(cudaStream_t st, long long* launches, LargeIcp* run, const kss_icp_params* prm, int poll);
int large_icp_result(cudaStream_t st, LargeIcp* run, float T[16], double* fitness, int* iters, int* converged);

// one full-resolution PCL-ICP run from host clouds
int large_icp_host(cudaStream_t st, long long* launches, const double* src, int n_s, const double* tgt, int n_t,
                   const kss_icp_params* prm, float T[16], double* fitness, int* iters, int* converged,
                   const DevAlloc& alloc);

}  // namespace kss
