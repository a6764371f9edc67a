// kss_kernels.h -- internal launch interface between the C ABI (kss_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/kss_icp_b200.h"

namespace kss {

constexpr int SMALL_MAX = KSS_SMALL_MAX;     // max points per cloud on the small (smem-resident) path
constexpr int TILE = 32;                     // target points per culling tile
constexpr int MAX_TILES = SMALL_MAX / TILE;  // 64 -> two tiles per lane

typedef kss_pair_result PairOut;

// candidate grid buffers (kss_cg.cuh): per pair 8 floats of geometry, CG_HDR_TOTAL header words,
// CG_ARENA u16 list entries, an arena cursor and an ok flag
struct CgBuffers {
    float* geom; unsigned long long* hdr; unsigned short* arena; unsigned* cursor; int* ok;
    unsigned short* wl; unsigned* wl_cnt; uint4* wl2; unsigned* wl2_cnt;      // finest-level cells to refine (kss_cg.cuh), per pair CG_WL_CAP entries
};
size_t cg_hdr_words_per_pair();
size_t cg_arena_entries_per_pair();
size_t cg_worklist_entries_per_pair();
size_t cg_worklist2_entries_per_pair();

struct IcpArgs {
    // source: original (double) simplified cloud, similarity applied in-kernel (mode 0/1) or used as is (mode 2)
    const double* src_f64; const int* cnt_s; int cap_s;
    const unsigned short* s_perm;
    const float4* t_sorted; const float* t_box; const unsigned short* t_inv; const int* cnt_t; int cap_t; int cap_tpad;
    const double* align8;
    int mode;                 // 0 judge, 1 hypothesis slot, 2 explicit input
    int runs_per_pair;        // 1 + hypothesis slots
    int hpad, G;
    const int* best_h; const int* minima; const int* n_minima;
    const double* trig_accum; const double* trig_list;
    double judge_thr;         // mode 1 runs only if judge fitness > thr (thr < 0: always)
    // candidate grid of the pair (null: tile search)
    const float* cg_geom; const unsigned long long* cg_hdr; const unsigned short* cg_arena; const int* cg_ok;
    // PCL parameters
    int max_iter; double max_dist_sqr, rot_thr, trans_thr, mse_rel, mse_abs;
    // per-run outputs, run = pair * runs_per_pair + (mode==1 ? 1 + slot : 0)
    float* run_T; double* run_fit; int* run_iters; int* run_conv;
    // hypothesis slots that serve several hypotheses (more local minima than slots): run_hyp = 2 * l + qualifies of the
    // run the slot ended up holding, run_tot = {sum of iterations, number of runs} of the slot (both may be null)
    int* run_hyp; int* run_tot;
    // hypothesis sharding over ranks (KSS_ICP.hpp:102-118 is a serial loop over independent runs): the CTA of slot s runs
    // l = s * world + rank, then strides by slots * world; hyp_fit [P][2][hpad] (null: off) receives fitness and iteration
    // count of EVERY run at its l -- the vector one all-reduce(MIN) completes on every rank
    int hyp_rank, hyp_world; double* hyp_fit;
    // optional trace
    unsigned long long* phase_cycles;      // diagnostics (tools/): per-phase clock64 sums of thread 0, or null
    int trace_cap; int32_t* trace_idx; float* trace_T; double* trace_mse; float* trace_src;
};

size_t icp_smem_bytes(int cap_s, int cap_t, int cap_tpad);

cudaError_t launch_sort_target(cudaStream_t st, int P, const double* pts, const int* cnt, int cap,
                               float4* t_sorted, float* t_box, unsigned short* t_inv, int cap_pad);
cudaError_t launch_sort_source(cudaStream_t st, int P, const double* pts, const int* cnt, int cap,
                               unsigned short* perm);
cudaError_t launch_middle_align(cudaStream_t st, int P, const double* sim_s, const int* cnt_s, int cap_s,
                                const double* sim_t, const int* cnt_t, int cap_t, double* align8, double* s_al);
cudaError_t launch_sweep(cudaStream_t st, int P, const double* s_al, const int* cnt_s, int cap_s,
                         const unsigned short* s_perm, const float4* t_sorted, const float* t_box,
                         const int* cnt_t, int cap_tpad, const double* trig_accum, int G, int score_mode,
                         float* rbuf, int hpad, const CgBuffers* cg, int ij_lo, int ij_hi);
cudaError_t launch_cg_build(cudaStream_t st, int P, int geom_mode, const double* a, const int* cnt_a, int cap_a,
                            const double* b, const int* cnt_b, int cap_b, const float4* t_sorted,
                            const unsigned short* t_inv, int cap_t, const int* cnt_t,
                            int cap_tpad, const CgBuffers& cg, int* launches);
cudaError_t launch_sweep_finalize(cudaStream_t st, int P, const float* rbuf, const int* cnt_s, int cap_s, int hpad,
                                  int G, int score_mode, double* value, int* best_h, int* minima, int* n_minima,
                                  int h_lo, int h_hi, int phases);
cudaError_t launch_icp(cudaStream_t st, int P, int slots, const IcpArgs& a);
cudaError_t launch_select(cudaStream_t st, int P, int runs_per_pair, int hpad, int G, double judge_thr,
                          const double* align8, const double* run_fit, const int* run_iters, const float* run_T,
                          const int* run_conv, const int* run_hyp, const int* run_tot,
                          const int* best_h, const int* minima, const int* n_minima, PairOut* out);
cudaError_t launch_select_sharded(cudaStream_t st, int P, int runs_per_pair, int hpad, int G, double judge_thr, const double* align8,
                                  const double* run_fit, const int* run_iters, const double* hyp_fit, const int* best_h,
                                  const int* minima, const int* n_minima, PairOut* out, int* win_minima, int* n_win);
cudaError_t launch_finish_sharded(cudaStream_t st, int P, int runs_per_pair, const double* run_fit, const int* run_iters,
                                  const int* run_conv, const float* run_T, const double* run2_fit, const int* run2_iters,
                                  const int* run2_conv, const float* run2_T, PairOut* out);
cudaError_t launch_final_apply(cudaStream_t st, int P, const double* full_s, const int* cnt_S, int cap_S,
                               const double* align8, const PairOut* out, const double* trig_accum,
                               const double* trig_list, int G, double* point_align);
cudaError_t launch_apply_similarity(cudaStream_t st, const double* pts, int n, const double* a7, const double* cs, double* out);
cudaError_t launch_apply_transform(cudaStream_t st, const double* pts, int n, const float* T, double* out);
// mode 0: idx/d2 out ; mode 1: PCR_QM sums -> out3[p*out3_stride + {0,1,2}]
cudaError_t launch_nn_small(cudaStream_t st, int P, int mode, const double* q, const int* cnt_q, int cap_q,
                            const unsigned short* q_perm, const float4* t_sorted, const float* t_box,
                            const int* cnt_t, int cap_tpad, int* idx, float* d2, double* out3, int out3_stride);

}  // namespace kss
