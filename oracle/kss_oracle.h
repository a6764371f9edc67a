/*
 * kss_oracle.h -- CPU restatement of the KSS-ICP registration hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under kss-icp_b200/ may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the CPU arm.
 *
 * PARITY UNPINNED: the reference (vvvwo/KSS-ICP) ships no tests or golden
 * vectors, and its ICP / kd-tree / SVD live in un-vendored PCL 1.8.1 + FLANN +
 * Eigen3 (PS_AIS_Simplification.vcxproj:80,90,126), none of which exists in
 * this environment.  This file restates
 *   - KSS_ICP.hpp:53-131,185-233,323-356      (orchestration, constants)
 *   - initRegistrationKSS.hpp:75-109,144-296,365-404,430-522 (align, sweep, minima)
 *   - registrationMeasure.hpp:47-98           (MSE / RMSE / MAE)
 *   - the published PCL 1.8.1 algorithms reached from those call sites
 *     (IterativeClosestPoint::computeTransformation, CorrespondenceEstimation,
 *      pcl::umeyama, DefaultConvergenceCriteria, Registration::getFitnessScore,
 *      KdTreeFLANN/L2_Simple exact 1-NN) as described in SURVEY.md Appendix A.
 * The only external anchors are the reference's shipped data pairs
 * (data/registration/<Model>.{wlop,gird} + transfer.txt known rotations).
 *
 * All floating-point expressions are written with one rounding per operator and
 * the file must be compiled with -ffp-contract=off (no FMA contraction).
 */
#ifndef KSS_ORACLE_H
#define KSS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* NN search implementation inside the oracle (results are identical; the
 * kd-tree exists so that the CPU baseline is not handicapped by brute force). */
enum { OKSS_NN_BRUTE = 0, OKSS_NN_KDTREE = 1 };

/* Summation order of the float reductions inside the PCL-ICP restatement.
 * SERIAL   = plain index-order loops (what scalar Eigen / PCL loops would do).
 * CANON256 = 32-lane strided partials + xor-butterfly per 256-element chunk,
 *            applied recursively (what a vectorised Eigen redux looks like, and
 *            the order contract of the CUDA path; see DESIGN.md "reduction
 *            order contract").  Eigen's real order is not recoverable
 *            (SURVEY.md A.4), so both are equally faithful. */
enum { OKSS_SUM_SERIAL = 0, OKSS_SUM_CANON256 = 1 };

/* Sweep score (initRegistrationKSS.hpp:406-479): Ave is the live one. */
enum { OKSS_SCORE_AVE = 0, OKSS_SCORE_MAX = 1, OKSS_SCORE_DIFF = 2 };

typedef struct okss_icp_params {
    int    max_iterations;       /* KSS_ICP.hpp:159  (1000 from Main_KSS_ICP.cpp:81) */
    double max_corr_dist;        /* KSS_ICP.hpp:156  = 1       */
    double transformation_eps;   /* KSS_ICP.hpp:157  = 1e-10   */
    double fitness_eps;          /* KSS_ICP.hpp:158  = 0.001   */
    int    sum_order;            /* OKSS_SUM_*                 */
    int    nn_method;            /* OKSS_NN_*                  */
} okss_icp_params;

/* optional per-iteration trace (any pointer may be NULL) */
typedef struct okss_icp_trace {
    int      cap_iters;      /* rows available in the arrays below            */
    int32_t* corr_idx;       /* [cap_iters][n_s] target index, -1 if rejected */
    float*   T_k;            /* [cap_iters][16] row-major per-iteration T     */
    double*  mse;            /* [cap_iters] mean squared corr. distance       */
    float*   src_k;          /* [cap_iters][n_s][3] source cloud BEFORE iter  */
} okss_icp_trace;

typedef struct okss_pair_result {
    double align[7];         /* x_middle_S,y_,z_, x_middle,y_,z_, scale       */
    double best_angle[3];    /* accumulated loop values (initRegistrationKSS.hpp:259-261) */
    int    best_index[3];
    int    G;                /* angles per axis                               */
    int    n_minima;         /* |angleList|                                   */
    int    branch_multi;     /* 1 iff judge fitness > 0.0005 (KSS_ICP.hpp:99) */
    int    winner;           /* angleList index chosen, -1 in the else branch */
    double used_angle[3];    /* angles finally applied to the full source     */
    double judge_fitness;
    double final_fitness;
    int    judge_iters;
    int    final_iters;
    int    total_icp_iters;  /* summed over all ICP runs of the pair          */
    int    n_icp_runs;
    float  T[16];            /* final ICP transformation, row-major           */
    double mse, rmse, mae;   /* PCR_QM on pointAlign vs full target           */
} okss_pair_result;

/* exact 1-NN, float, FLANN L2_Simple expression ((dx*dx)+dy*dy)+dz*dz.
 * Tie rule of this oracle: lowest target index among fp32-equal distances
 * (FLANN's own tie order depends on its tree and is not reproducible). */
void okss_nn(const float* q, int nq, const float* t, int nt, int method,
             int32_t* idx, float* d2);

/* initRegistration_MiddleAlign (initRegistrationKSS.hpp:144-220) */
void okss_middle_align(const double* src, int n_s, const double* tgt, int n_t,
                       double out7[7], double* src_aligned /* [n_s][3] or NULL */);

/* the angle grid of initRegistrationKSS.hpp:245 ; returns G, fills accum[G]
 * with the accumulated loop values and list[G] with index*6.3/step. */
int okss_sweep_angles(double step, double* accum, double* list, int cap);

/* initRegistration_Rotation() (initRegistrationKSS.hpp:222-296, 481-522).
 * value: [G*G*G]; minima_idx: [G*G*G][3] in loop order; returns G. */
int okss_sweep(const double* src_aligned, int n_s, const double* tgt, int n_t,
               double step, int score_mode, int nn_method,
               double* value, double best_angle[3], int best_index[3],
               int* minima_idx, int* n_minima);

/* one hypothesis score (initRegistration_Error_Ave etc.), also returns the
 * float queries actually searched and their NN (for teacher-forced tests). */
double okss_sweep_score(const double* src_aligned, int n_s, const double* tgt, int n_t,
                        const double angles[3], int score_mode, int nn_method,
                        float* queries /* [n_s][3] or NULL */,
                        int32_t* idx /* or NULL */, float* d2 /* or NULL */);

/* initRegistration_Rotation(src) / _Rotation_Angle (initRegistrationKSS.hpp:75-109) */
void okss_apply_similarity(const double* pts, int n, const double align7[7],
                           const double angles[3], double* out);

/* one PCL-1.8.1 ICP run as wrapped by KSS_ICP.hpp:323-356.
 * Returns fitness (getFitnessScore). T is row-major 4x4 float. */
double okss_icp(const double* src, int n_s, const double* tgt, int n_t,
                const okss_icp_params* p, float T[16], int* iters, int* converged,
                okss_icp_trace* trace /* or NULL */);

/* KSS_ICP.hpp:222-230: float coefficients promoted to double */
void okss_apply_transform(const float T[16], const double* pts, int n, double* out);

/* registrationMeasure.hpp:47-98 -> out = {MSE, RMSE, MAE} */
void okss_nn_metrics(const double* a, int n_a, const double* t, int n_t,
                     int nn_method, double out3[3]);

/* umeyama rigid fit on explicit correspondences (float), for unit tests */
void okss_umeyama(const float* src, const float* dst, int n, int sum_order, float T[16]);

/* 3x3 SVD used by okss_umeyama (two-sided Jacobi, float): A = U diag(s) V^T */
void okss_svd3(const float A[9], float U[9], float s[3], float V[9]);

/* KSSICP_Registration after simplification (KSS_ICP.hpp:86-130) + PCR_QM.
 * sim_* are the simplified clouds (AIVS output), full_* the originals.
 * point_align: [N_s][3] or NULL. */
void okss_register(const double* sim_s, int n_s, const double* sim_t, int n_t,
                   const double* full_s, int N_s, const double* full_t, int N_t,
                   double step, int max_iter, int sum_order, int nn_method,
                   okss_pair_result* res, double* point_align);

/* batch of equally-shaped pairs, std::thread pool over pairs (threads<=0: all cores).
 * Arrays are concatenated per pair. Returns the number of threads used.
 * sim_s == sim_t == NULL: every pair is simplified first (pNumber rule + okss_aivs_simplify), i.e. the whole of
 * KSSICP_init + KSSICP_Registration (KSS_ICP.hpp:53-130). */
int okss_register_batch(int n_pairs,
                        const double* sim_s, int n_s, const double* sim_t, int n_t,
                        const double* full_s, int N_s, const double* full_t, int N_t,
                        double step, int max_iter, int sum_order, int nn_method,
                        int threads, okss_pair_result* res);

/* full-resolution ICP iteration pieces for the large-cloud parity tests:
 * one iteration on float clouds: returns n_corr, fills T_k, mse. */
int okss_icp_iteration(const float* src, int n_s, const float* tgt, int n_t,
                       double max_corr_dist, int sum_order, int nn_method,
                       int32_t* idx, float* d2, float T_k[16], double* mse,
                       float* src_out /* transformed, or NULL */);

/* AIVS simplification + BallRegion voxel grid (the step before the hot path):
 * pointPipeline_init_point_withoutUniform + AIVS_Pro_init + AIVS_simplification(pointNum)
 * (KSS_ICP.hpp:71-81).  out [<= n][3], out_idx original indices (either may be NULL); returns the count,
 * which can be below pointNum (B9). */
int okss_aivs_simplify(const double* pts, int n, int pointNum, double* out, int32_t* out_idx);

/* canonical sums exposed for tests */
float  okss_canon_sum_f32(const float* v, int n);
double okss_canon_sum_f64(const double* v, int n);

int okss_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
