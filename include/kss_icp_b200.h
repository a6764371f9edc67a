/*
 * kss_icp_b200.h -- C ABI of the B200-native KSS-ICP registration hot path.
 *
 * The reference (vvvwo/KSS-ICP) has no FFI boundary: the path is reached through two
 * header-only C++ classes (KSS_ICP.hpp, initRegistrationKSS.hpp) plus PCR_QM
 * (registrationMeasure.hpp) that call PCL 1.8.1 directly.  Each entry point below
 * replaces the body of one of those methods; the same-named C++ classes in
 * kss-icp_b200/host/ forward to them.
 *
 * Conventions
 *   - clouds are packed double[n][3] (the flattened vector<vector<double>> of the reference);
 *     narrowing to float happens on the device with round-to-nearest exactly where the
 *     reference narrows (KSS_ICP.hpp:328-333, initRegistrationKSS.hpp:231-235, 440-442).
 *   - every function returns 0 on success, a negative KSS_ERR_* otherwise;
 *     kss_last_error(ctx) gives the message.  Nothing throws, nothing falls back to the CPU.
 *   - the caller owns all host buffers; the library owns device memory inside ctx.
 *   - one host thread per ctx; one ctx per GPU; different ctx may run concurrently.
 *   - "_device" variants take device pointers on ctx's device and enqueue on ctx's stream.
 */
#ifndef KSS_ICP_B200_H
#define KSS_ICP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KSS_OK               0
#define KSS_ERR_ARG         -1   /* bad argument (null pointer, n < 1, ...)              */
#define KSS_ERR_CUDA        -2   /* a CUDA runtime call failed (message in last_error)  */
#define KSS_ERR_UNSUPPORTED -3   /* size / mode outside what this build implements      */
#define KSS_ERR_NCCL        -4
#define KSS_ERR_NOMEM       -5

#define KSS_SMALL_MAX 2048       /* clouds up to this size run on the shared-memory path */

/* sweep score modes: live one is AVE (initRegistrationKSS.hpp:255); the dead siblings
 * _Error (:406-428) and _Error_Diff (:452-479) stay selectable */
#define KSS_SCORE_AVE  0
#define KSS_SCORE_MAX  1
#define KSS_SCORE_DIFF 2
/* NOT a mode of the released sources: a reading of the authors' closed CUDA build (EXE/KSS-ICP-VCG-Cuda.exe, symbols
 * voxelCudaBlock / pt_VoxelJudge): the target becomes an NV^3 voxel-occupancy grid (NV = sqrt(n_t / 16), 4..32) over the cube +-1.1 max|coordinate| and
 * a hypothesis scores the share of source points that land in an empty voxel -- no nearest-neighbour search at all.
 * Only kss_rotation_sweep takes it; the registration entry points always use KSS_SCORE_AVE like KSSICP_Registration. */
#define KSS_SCORE_VOXEL 3

typedef struct kss_ctx kss_ctx;

/* ICP parameters as set at KSS_ICP.hpp:156-159 (defaults of kss_icp_params_default). */
typedef struct kss_icp_params {
    int    max_iterations;        /* setMaximumIterations        (1000, Main_KSS_ICP.cpp:81) */
    double max_corr_dist;         /* setMaxCorrespondenceDistance (1)     */
    double transformation_eps;    /* setTransformationEpsilon     (1e-10) */
    double fitness_eps;           /* setEuclideanFitnessEpsilon   (0.001) */
} kss_icp_params;

/* optional per-iteration trace of kss_icp (tests): any pointer may be NULL */
typedef struct kss_icp_trace {
    int      cap_iters;
    int32_t* corr_idx;            /* [cap_iters][n_s] target index per source point, -1 = rejected */
    float*   T_k;                 /* [cap_iters][16] */
    double*  mse;                 /* [cap_iters]     */
    float*   src_k;               /* [cap_iters][n_s][3] source cloud before the iteration */
} kss_icp_trace;

/* result of one registration (KSSICP_Registration + PCR_QM) */
typedef struct kss_pair_result {
    double align[8];              /* x_middle_S,y,z ; x_middle,y,z ; scale ; pad (initRegistrationKSS.hpp:38-44) */
    double judge_fitness;         /* E_d_init (KSS_ICP.hpp:93)  */
    double final_fitness;         /* resultFitness (KSS_ICP.hpp:130) */
    double mse, rmse, mae;        /* PCR_QM (registrationMeasure.hpp:83-96) */
    float  T[16];                 /* icp.getFinalTransformation(), row-major */
    int    G;                     /* angles per axis (9 at step 8) */
    int    best_h;                /* flattened (i*G+j)*G+k index of the sweep winner */
    int    n_minima;              /* |angleList| */
    int    branch_multi;          /* 1 iff judge_fitness > 0.0005 (KSS_ICP.hpp:99) */
    int    winner;                /* angleList index chosen (KSS_ICP.hpp:113-116), -1 otherwise */
    int    used_h;                /* flattened index of the angles finally applied */
    int    use_list;              /* 1: used angles are index*6.3/step, 0: accumulated loop values */
    int    judge_iters, final_iters, total_icp_iters, n_icp_runs;
    int    overflow;              /* informational: the pair had more local minima than hypothesis slots, so some
                                     slots ran several hypotheses one after the other (same results either way) */
    int    final_converged;       /* icp.hasConverged() of the final run (KSS_ICP.hpp:217): 0 iff it stopped with
                                     fewer than 3 correspondences */
    int    reserved_[2];
} kss_pair_result;

/* batch of registrations: arrays are [n_pairs][cap][3] doubles, counts optional (NULL = cap) */
typedef struct kss_batch {
    int n_pairs;
    int cap_s, cap_t;             /* simplified clouds (<= KSS_SMALL_MAX)          */
    int cap_S, cap_T;             /* full-resolution clouds                        */
    const double* sim_s;          /* AIVS-simplified source  (pointCloudS, KSS_ICP.hpp:81) */
    const double* sim_t;          /* AIVS-simplified target  (pointCloudT, KSS_ICP.hpp:75) */
    const double* full_s;         /* pointSource */
    const double* full_t;         /* pointTarget */
    const int* cnt_s; const int* cnt_t; const int* cnt_S; const int* cnt_T;
    double step;                  /* `accurate` of KSSICP_init (8, Main_KSS_ICP.cpp:80) */
    kss_icp_params icp;
    double judge_threshold;       /* 0.0005 (KSS_ICP.hpp:99) */
} kss_batch;

/* ---- context ------------------------------------------------------------------------ */
int  kss_ctx_create(int device, kss_ctx** out);
int  kss_ctx_create_on_stream(int device, void* cuda_stream, kss_ctx** out);
void kss_ctx_destroy(kss_ctx* ctx);
const char* kss_last_error(kss_ctx* ctx);
int  kss_ctx_synchronize(kss_ctx* ctx);
void kss_icp_params_default(kss_icp_params* p);
void kss_batch_default(kss_batch* b);
/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
long long kss_ctx_launch_count(kss_ctx* ctx);
/* hypothesis ICP runs per pair that get their own CTA in the batched launch (default 32, 1..729); pairs with more
 * local minima (KSS_ICP.hpp:102-118) are handled inside the same launch: a slot then runs several hypotheses in
 * turn.  Results do not depend on the value. */
int kss_ctx_set_hyp_slots(kss_ctx* ctx, int slots);

/* per-stage device timing of the batched pipeline: CUDA events recorded on ctx's stream around
 * each stage (bench.py derives the per-kernel roofline from these; off by default) */
#define KSS_STAGE_PREP            0   /* MiddleAlign + Morton sorts        */
#define KSS_STAGE_SWEEP           1   /* sweep_kernel                       */
#define KSS_STAGE_SWEEP_FINALIZE  2
#define KSS_STAGE_ICP_JUDGE       3
#define KSS_STAGE_ICP_HYP         4
#define KSS_STAGE_SELECT_APPLY    5
#define KSS_STAGE_METRICS         6
#define KSS_STAGE_LARGE_BUILD     7   /* large path: Morton bucket sort + box pyramid */
#define KSS_STAGE_LARGE_NN        8   /* large path: hierarchical NN kernel           */
#define KSS_STAGE_LARGE_REDUCE    9   /* large path: canonical reductions + SVD       */
#define KSS_STAGE_CG_BUILD       10   /* candidate grid build (per pair, once)         */
#define KSS_STAGE_AIVS           11   /* AIVS simplification of raw clouds             */
#define KSS_STAGE_LARGE_TRACK    12   /* large path: the streaming kernel of an ICP iteration alone (inside LARGE_NN) */
#define KSS_STAGE_COUNT          13
int kss_ctx_set_timing(kss_ctx* ctx, int enable);      /* also resets the accumulators */
int kss_ctx_stage_ms(kss_ctx* ctx, int stage, double* ms, long long* calls);

/* ---- single-object entry points ------------------------------------------------------ */

/* replaces initRegistration_KSS::initRegistration_MiddleAlign (initRegistrationKSS.hpp:144-220).
 * out7 = {x_middle_S,y,z, x_middle,y,z, scale}; src_aligned [n_s][3] optional. */
int kss_middle_align(kss_ctx* ctx, const double* src, int n_s, const double* tgt, int n_t,
                     double out7[7], double* src_aligned);

/* the angle grid of initRegistrationKSS.hpp:245 (host arithmetic only): returns G */
int kss_sweep_angles(double step, double* accum, double* list, int cap);

/* replaces initRegistration_KSS::initRegistration_Rotation() (initRegistrationKSS.hpp:222-296,
 * 430-450, 481-522).  src_aligned is the cloud MiddleAlign produced.  value [G^3]; best_angle =
 * accumulated loop values; minima [<=G^3][3] grid indices in loop order (angleList = idx*6.3/step). */
int kss_rotation_sweep(kss_ctx* ctx, const double* src_aligned, int n_s, const double* tgt, int n_t,
                       double step, int score_mode, double* value, int* G_out,
                       double best_angle[3], int best_index[3], int* minima, int* n_minima);

/* replaces initRegistration_Rotation(src) / _Rotation_Angle(src, angle) (initRegistrationKSS.hpp:75-109) */
int kss_apply_similarity(kss_ctx* ctx, const double* pts, int n, const double align7[7],
                         const double angles[3], double* out);

/* replaces one pcl::IterativeClosestPoint run as wrapped by shapeRegistration_ICP_Judge /
 * _AngleList / _ICP (KSS_ICP.hpp:185-274, 323-356): T row-major 4x4, fitness = getFitnessScore(). */
int kss_icp(kss_ctx* ctx, const double* src, int n_s, const double* tgt, int n_t,
            const kss_icp_params* prm, float T[16], double* fitness, int* iters, int* converged,
            kss_icp_trace* trace);

/* the same ICP on clouds of any size, driven in steps (full-resolution overload
 * shapeRegistration_ICP(int iter), KSS_ICP.hpp:133-183; bench.py times `iterate`):
 * begin = H2D + Morton bucket sort + box pyramid; iterate = enqueue `count` iterations
 * (NN, reductions, SVD, transform; no-ops once converged; fitness_eps < 0 disables the
 * convergence tests for steady-state timing); end = getFitnessScore pass + results. */
int kss_icp_large_begin(kss_ctx* ctx, const double* src, int n_s, const double* tgt, int n_t);
int kss_icp_large_iterate(kss_ctx* ctx, const kss_icp_params* prm, int count);
int kss_icp_large_end(kss_ctx* ctx, const kss_icp_params* prm, float T[16], double* fitness, int* iters, int* converged);

/* replaces the final apply loop at KSS_ICP.hpp:222-230 */
int kss_apply_transform(kss_ctx* ctx, const float T[16], const double* pts, int n, double* out);

/* replaces PCR_QM::PCR_QM_Start (registrationMeasure.hpp:47-98): out3 = {MSE, RMSE, MAE} */
int kss_nn_metrics(kss_ctx* ctx, const double* a, int n_a, const double* t, int n_t, double out3[3]);

/* exact 1-NN (replaces pcl::KdTreeFLANN::nearestKSearch, K=1): idx into t, squared distance (float).
 * Tie rule: lowest target index among fp32-equal distances. */
int kss_nn_search(kss_ctx* ctx, const double* q, int n_q, const double* t, int n_t,
                  int32_t* idx, float* d2);

/* ---- AIVS simplification (the step in front of the path) --------------- */

/* replaces pointPipeline_init_point_withoutUniform + AIVS_Pro_init + AIVS_simplification(pointNum)
 * (pointPipeline.hpp:62-103, Method_AIVS_SimPro.hpp:66-123, called at KSS_ICP.hpp:72-82): out [out_cap][3]
 * receives the simplified cloud (out_cap >= n always suffices; point_num does unless the trim step ends
 * early, which is reported as KSS_ERR_UNSUPPORTED), *out_n its size, out_idx (optional) the
 * position of every kept point in `pts`.  n >= 1; clouds with zero extent are KSS_ERR_UNSUPPORTED. */
int kss_aivs_simplify(kss_ctx* ctx, const double* pts, int n, int point_num, double* out, int out_cap,
                      int* out_n, int32_t* out_idx);

/* many clouds at once: pts [n_clouds][cap][3], cnt (optional) their sizes, point_num_each (optional,
 * else point_num for all); out [n_clouds][out_cap][3], out_cnt [n_clouds], out_idx optional */
int kss_aivs_simplify_batch(kss_ctx* ctx, int n_clouds, const double* pts, const int* cnt, int cap,
                            const int* point_num_each, int point_num, double* out, int out_cap,
                            int* out_cnt, int32_t* out_idx);
/* the same on device pointers, enqueued on ctx's stream; kss_aivs_status() afterwards synchronizes
 * and reports a cloud the simplification could not handle */
int kss_aivs_simplify_batch_device(kss_ctx* ctx, int n_clouds, const double* d_pts, const int* d_cnt, int cap,
                                   const int* d_point_num, int point_num, double* d_out, int out_cap,
                                   int* d_out_cnt, int32_t* d_out_idx);
int kss_aivs_status(kss_ctx* ctx);

/* ---- batched registration (KSSICP_Registration + PCR_QM) -------------------------------- */
/* kss_batch.sim_s == sim_t == NULL selects the whole of KSSICP_Registration (KSS_ICP.hpp:70-130): the
 * library computes pNumber = min(|S|,|T|)/2 (<= 2000) per pair and simplifies both clouds with AIVS on the
 * device; cap_s / cap_t / cnt_s / cnt_t are ignored.  Otherwise the given simplified clouds are used. */

/* host buffers in, host results out (H2D/D2H inside). point_align [n_pairs][cap_S][3] optional.
 * The clouds are copied from wherever the pointers point; from page-locked memory (cudaHostAlloc / cudaHostRegister)
 * the copies run at the link's rate and overlap the kernels, from pageable memory the driver stages them (a 1M-point
 * pair: 9 ms instead of 5). */
int kss_register_batch(kss_ctx* ctx, const kss_batch* b, kss_pair_result* results, double* point_align);

/* device buffers in (b->* are device pointers), results/point_align are device pointers;
 * everything is enqueued on ctx's stream; call kss_ctx_synchronize before reading. */
int kss_register_batch_device(kss_ctx* ctx, const kss_batch* b, kss_pair_result* d_results,
                              double* d_point_align);

/* ---- multi-GPU (one ctx per GPU; a single pair / ICP run is never split) ----------------- */

/* registration pairs are independent (Main_KSS_List.cpp:132-167 is a plain loop): contiguous blocks of the batch on
 * n_ctx contexts (one per GPU, this process), one host thread each, host buffers, no data-path collective */
int kss_register_batch_multi(kss_ctx** ctxs, int n_ctx, const kss_batch* b, kss_pair_result* results, double* point_align);

/* NCCL plumbing for hypothesis sharding.  One process per GPU: rank 0 calls kss_nccl_get_unique_id, hands the 128 bytes to
 * the other ranks (torch.distributed / MPI / a file), every rank calls kss_ctx_nccl_init.  One process, several GPUs:
 * kss_ctx_nccl_init_all on the array of contexts.  libnccl.so.2 is resolved at run time. */
int kss_nccl_get_unique_id(void* id128);
int kss_ctx_nccl_init(kss_ctx* ctx, const void* id128, int rank, int world);
int kss_ctx_nccl_init_all(kss_ctx** ctxs, int n);

/* KSSICP_Registration + PCR_QM with the rotation hypotheses of every pair sharded over the ranks of ctx's communicator
 * (every rank passes the SAME batch and receives the SAME results): the 729-hypothesis sweep is slabbed over the ranks and
 * completed by one ncclAllGather of the score grid (initRegistrationKSS.hpp:245-268); the ICP runs over angleList
 * (KSS_ICP.hpp:102-118) are dealt round-robin and their fp64 fitness vector is completed by ONE ncclAllReduce(min); every
 * rank then applies the reference's selection rule (KSS_ICP.hpp:113) and repeats the winner's run (KSS_ICP.hpp:130).
 * Host buffers; the collectives run on ctx's stream.  world == 1 (no kss_ctx_nccl_init) is the unsharded path. */
int kss_register_batch_hyp_sharded(kss_ctx* ctx, const kss_batch* b, kss_pair_result* results, double* point_align);

/* one pair, the exact flow of KSSICP_Registration (KSS_ICP.hpp:86-130); sim_s = sim_t = NULL: raw clouds,
 * simplified by the library first (KSS_ICP.hpp:53-82), n_s / n_t ignored */
int kss_register(kss_ctx* ctx, const double* sim_s, int n_s, const double* sim_t, int n_t,
                 const double* full_s, int N_s, const double* full_t, int N_t,
                 double step, int max_iter, kss_pair_result* result, double* point_align);

/* diagnostics for tools/: copies a range of a named internal device buffer to the host (names are internal) */
int kss_debug_read(kss_ctx* ctx, const char* name, size_t offset, size_t bytes, void* dst);

#ifdef __cplusplus
}
#endif
#endif
