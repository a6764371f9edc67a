// kss_small.cu -- sm_100a kernels for clouds that fit one CTA's shared memory (<= 2048 points):
// the sizes the reference actually runs its sweep and ICP on (pNumber <= 2000, KSS_ICP.hpp:57-66).
//
//   sort_cloud_kernel      Morton order + 32-point tiles + tile boxes of a target; Morton perm of a source
//   middle_align_kernel    initRegistration_MiddleAlign            (initRegistrationKSS.hpp:144-220)
//   sweep_kernel           initRegistration_Rotation() inner loops (initRegistrationKSS.hpp:245-256, 430-450)
//   sweep_finalize_kernel  serial score sums, argmin, local minima  (initRegistrationKSS.hpp:258-289, 481-522)
//   icp_small_kernel       one whole PCL-1.8.1 ICP run per CTA      (KSS_ICP.hpp:323-356 + SURVEY.md A.2-A.7)
//   select_kernel          hypothesis choice                        (KSS_ICP.hpp:99-125)
//   final_apply_kernel     similarity + final 4x4 on the full cloud (KSS_ICP.hpp:119-124, 222-230)
//   nn_small_kernel        bare exact 1-NN (kss_nn_search) and, with serial sums per cloud, PCR_QM (registrationMeasure.hpp:47-98)
//   cg_* kernels           the per-pair candidate grid (kss_cg.cuh) the sweep and the ICP runs query
#include "kss_device.cuh"
#include "kss_cg.cuh"
#include "kss_kernels.h"

namespace kss {

// =============================================================== sort_cloud_kernel
// One CTA per cloud.  mode 0: target -> t_sorted (float4 Morton order, padded), t_box, t_inv
//                     mode 1: source -> perm only (u16 original index per Morton position)
// pts_f64 may be the original double cloud (narrowed here with RN, like KSS_ICP.hpp:328-333).
__global__ void __launch_bounds__(256)
sort_cloud_kernel(const double* __restrict__ pts, const int* __restrict__ cnt, int cap, int mode,
                  float4* __restrict__ t_sorted, float* __restrict__ t_box, unsigned short* __restrict__ t_inv,
                  int cap_pad, unsigned short* __restrict__ perm) {
    extern __shared__ unsigned char smem_raw[];
    unsigned* keys = reinterpret_cast<unsigned*>(smem_raw);            // [2048]
    float* cx = reinterpret_cast<float*>(keys + SMALL_MAX);            // [2048] x3
    float* cy = cx + SMALL_MAX;
    float* cz = cy + SMALL_MAX;
    __shared__ unsigned bb[6];
    __shared__ float lo[3], inv[3];

    const int p = blockIdx.x;
    const int n = cnt ? cnt[p] : cap;
    const double* src = pts + (size_t)p * cap * 3;
    if (threadIdx.x < 3) { bb[threadIdx.x] = 0xffffffffu; bb[3 + threadIdx.x] = 0u; }
    __syncthreads();
    unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        float x = (float)src[3 * i], y = (float)src[3 * i + 1], z = (float)src[3 * i + 2];
        cx[i] = x; cy[i] = y; cz[i] = z;
        unsigned ox = f2ord(x), oy = f2ord(y), oz = f2ord(z);
        mn[0] = min(mn[0], ox); mn[1] = min(mn[1], oy); mn[2] = min(mn[2], oz);
        mx[0] = max(mx[0], ox); mx[1] = max(mx[1], oy); mx[2] = max(mx[2], oz);
    }
    for (int a = 0; a < 3; ++a) {
        unsigned m0 = __reduce_min_sync(KSS_FULL, mn[a]);
        unsigned m1 = __reduce_max_sync(KSS_FULL, mx[a]);
        if ((threadIdx.x & 31) == 0) { atomicMin(&bb[a], m0); atomicMax(&bb[3 + a], m1); }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float l = ord2f(bb[threadIdx.x]), h = ord2f(bb[3 + threadIdx.x]);
        float e = h - l;
        lo[threadIdx.x] = l;
        inv[threadIdx.x] = e > 0.0f ? 127.999f / e : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SMALL_MAX; i += blockDim.x)
        keys[i] = i < n ? ((morton21(cx[i], cy[i], cz[i], lo, inv) << 11) | (unsigned)i) : 0xffffffffu;
    bitonic_sort_smem<SMALL_MAX>(keys);

    if (mode == 1) {
        unsigned short* out = perm + (size_t)p * cap;
        for (int j = threadIdx.x; j < n; j += blockDim.x) out[j] = (unsigned short)(keys[j] & 2047u);
        return;
    }
    const int npad = (n + TILE - 1) / TILE * TILE;
    float4* ts = t_sorted + (size_t)p * cap_pad;
    unsigned short* ti = t_inv + (size_t)p * cap;
    for (int j = threadIdx.x; j < npad; j += blockDim.x) {
        float4 v;
        if (j < n) {
            int o = keys[j] & 2047u;
            v = make_float4(cx[o], cy[o], cz[o], __int_as_float(o));
            ti[o] = (unsigned short)j;
        } else {
            v = make_float4(PAD_COORD, PAD_COORD, PAD_COORD, __int_as_float(0x7fffffff));
        }
        ts[j] = v;
    }
    // tile boxes over real points only: one warp per tile
    float* tb = t_box + (size_t)p * 6 * MAX_TILES;
    const int ntiles = npad / TILE;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int t = warp; t < ntiles; t += nwarps) {
        int j = t * TILE + lane;
        int o = keys[min(j, n - 1)] & 2047u;  // pads replicate the last real point
        float x = cx[o], y = cy[o], z = cz[o];
        float v0 = warp_min_f(x), v1 = warp_min_f(y), v2 = warp_min_f(z);
        float v3 = warp_max_f(x), v4 = warp_max_f(y), v5 = warp_max_f(z);
        if (lane == 0) {
            tb[0 * MAX_TILES + t] = v0; tb[1 * MAX_TILES + t] = v1; tb[2 * MAX_TILES + t] = v2;
            tb[3 * MAX_TILES + t] = v3; tb[4 * MAX_TILES + t] = v4; tb[5 * MAX_TILES + t] = v5;
        }
    }
}

// =============================================================== middle_align_kernel
// One CTA per pair.  Every sum of the reference is a serial index-order double loop
// (initRegistrationKSS.hpp:150-207); here each such chain runs on one thread over
// shared-memory data so the 7 outputs are bit-identical to the serial definition.
__global__ void __launch_bounds__(256)
middle_align_kernel(const double* __restrict__ sim_s, const int* __restrict__ cnt_s, int cap_s,
                    const double* __restrict__ sim_t, const int* __restrict__ cnt_t, int cap_t,
                    double* __restrict__ align8, double* __restrict__ s_al) {
    extern __shared__ unsigned char smem_raw[];
    double* buf = reinterpret_cast<double*>(smem_raw);   // [2048*3]
    double* len = buf + 3 * SMALL_MAX;                   // [2048]
    __shared__ double mean[2][3];
    __shared__ double avg[2];

    const int p = blockIdx.x;
    for (int which = 0; which < 2; ++which) {            // 0: source, 1: target
        const int n = which == 0 ? (cnt_s ? cnt_s[p] : cap_s) : (cnt_t ? cnt_t[p] : cap_t);
        const double* src = which == 0 ? sim_s + (size_t)p * cap_s * 3 : sim_t + (size_t)p * cap_t * 3;
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) buf[i] = src[i];
        __syncthreads();
        if (threadIdx.x < 3) {
            double s = 0.0;
            const double* b = buf + threadIdx.x;
#pragma unroll 8
            for (int i = 0; i < n; ++i) s = __dadd_rn(s, b[3 * i]);
            mean[which][threadIdx.x] = __ddiv_rn(s, (double)n);
        }
        __syncthreads();
        const double mx = mean[which][0], my = mean[which][1], mz = mean[which][2];
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            double xl = __dsub_rn(buf[3 * i], mx), yl = __dsub_rn(buf[3 * i + 1], my), zl = __dsub_rn(buf[3 * i + 2], mz);
            double q = __dadd_rn(__dadd_rn(__dmul_rn(xl, xl), __dmul_rn(yl, yl)), __dmul_rn(zl, zl));
            len[i] = __dsqrt_rn(q);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
#pragma unroll 8
            for (int i = 0; i < n; ++i) s = __dadd_rn(s, len[i]);
            avg[which] = __ddiv_rn(s, (double)n);
        }
    }
    __syncthreads();
    // align8 = {x_middle_S, y_, z_, x_middle, y_, z_, scale, pad}
    double* a8 = align8 + (size_t)p * 8;
    const double scale = __ddiv_rn(avg[1], avg[0]);
    if (threadIdx.x < 3) {
        a8[threadIdx.x] = mean[1][threadIdx.x];
        a8[3 + threadIdx.x] = __dsub_rn(mean[1][threadIdx.x], mean[0][threadIdx.x]);
    }
    if (threadIdx.x == 3) { a8[6] = scale; a8[7] = 0.0; }
    if (s_al) {
        const int n = cnt_s ? cnt_s[p] : cap_s;
        const double* src = sim_s + (size_t)p * cap_s * 3;
        double* dst = s_al + (size_t)p * cap_s * 3;
        for (int i = threadIdx.x; i < 3 * n; i += blockDim.x) {
            int a = i % 3;
            dst[i] = align_coord(src[i], mean[1][a], __dsub_rn(mean[1][a], mean[0][a]), scale);
        }
    }
}


// =============================================================== cg_geom_kernel
// mode 0: queries live in the ball |q| <= R about the origin, R = max |p| over the (aligned) source
//         (every sweep hypothesis is a rotation about the origin)
// mode 1: explicit ICP input: cube around the bounding box of source and target
__global__ void __launch_bounds__(256)
cg_geom_kernel(int mode, const double* __restrict__ a, const int* __restrict__ cnt_a, int cap_a,
               const double* __restrict__ b, const int* __restrict__ cnt_b, int cap_b,
               float* __restrict__ geom, unsigned* __restrict__ cursor, int* __restrict__ ok, unsigned* __restrict__ wl_cnt,
               unsigned* __restrict__ wl2_cnt) {
    __shared__ unsigned bb[6];
    __shared__ unsigned long long r2max;
    const int p = blockIdx.x;
    if (threadIdx.x < 3) { bb[threadIdx.x] = 0xffffffffu; bb[3 + threadIdx.x] = 0u; }
    if (threadIdx.x == 0) r2max = 0ull;
    __syncthreads();
    double r2 = 0.0;
    unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    for (int which = 0; which < (mode == 1 ? 2 : 1); ++which) {
        const double* pts = which == 0 ? a + (size_t)p * cap_a * 3 : b + (size_t)p * cap_b * 3;
        const int n = which == 0 ? cnt_a[p] : cnt_b[p];
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
            r2 = fmax(r2, x * x + y * y + z * z);
            const unsigned ox = f2ord((float)x), oy = f2ord((float)y), oz = f2ord((float)z);
            mn[0] = min(mn[0], ox); mn[1] = min(mn[1], oy); mn[2] = min(mn[2], oz);
            mx[0] = max(mx[0], ox); mx[1] = max(mx[1], oy); mx[2] = max(mx[2], oz);
        }
    }
    atomicMax(&r2max, (unsigned long long)__double_as_longlong(r2));       // r2 >= 0: bits order as integers
    for (int k = 0; k < 3; ++k) { atomicMin(&bb[k], mn[k]); atomicMax(&bb[3 + k], mx[k]); }
    __syncthreads();
    if (threadIdx.x == 0) {
        float* g = geom + (size_t)p * 8;
        if (mode == 0) {
            const float R = fmaxf((float)(sqrt(__longlong_as_double((long long)r2max)) * 1.0001), 1e-12f);
            g[0] = g[1] = g[2] = 0.0f; g[3] = R * 1.1f; g[4] = R * 1.1f;   // 10% slack for ICP excursions
        } else {
            float half = 0.0f;
            for (int k = 0; k < 3; ++k) {
                const float l = ord2f(bb[k]), hgh = ord2f(bb[3 + k]);
                g[k] = 0.5f * (l + hgh);
                half = fmaxf(half, 0.5f * (hgh - l));
            }
            g[3] = fmaxf(half * 1.1f, 1e-12f);
            g[4] = __int_as_float(0x7f800000);
        }
        g[5] = g[6] = g[7] = 0.0f;
        cursor[p] = 0u;
        ok[p] = 1;
        wl_cnt[2 * p] = wl_cnt[2 * p + 1] = 0u;
        wl2_cnt[2 * p] = wl2_cnt[2 * p + 1] = 0u;
    }
}

// =============================================================== cg_level0_kernel
// Level 0 of the candidate grid: 64 cells, each filtering ALL targets.  One warp per cell, lanes stride over the
// targets in original-index order; the same passes as cg_level_kernel (four nearest to the centre, sphere rule +
// dominance), the kept ones compacted in order with ballots.  Lists of this level always go to the arena.
__global__ void __launch_bounds__(256)
cg_level0_kernel(const float4* __restrict__ t_sorted, const unsigned short* __restrict__ t_inv_all, int cap_t,
                 const int* __restrict__ cnt_t, int cap_tpad, const float* __restrict__ geom,
                 cg_hdr_t* __restrict__ hdr_all, unsigned short* __restrict__ arena_all, unsigned* __restrict__ cursor,
                 int* __restrict__ ok) {
    const int p = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int ng = CG_NG0;
    if (cell >= ng * ng * ng) return;
    const int n_t = cnt_t[p];
    const float4* __restrict__ tgt = t_sorted + (size_t)p * cap_tpad;
    const unsigned short* __restrict__ tinv = t_inv_all + (size_t)p * cap_t;
    cg_hdr_t* hdr = hdr_all + (size_t)p * CG_HDR_TOTAL;
    unsigned short* arena = arena_all + (size_t)p * CG_ARENA;
    const float* gm = geom + (size_t)p * 8;
    const float R = gm[3], ballR = gm[4];
    const float h = 2.0f * R / (float)ng;
    const float rho = h * 0.8660254f * 1.002f;
    const int ix = cell & 3, iy = (cell >> 2) & 3, iz = cell >> 4;
    const float ox = -R + ((float)ix + 0.5f) * h, oy = -R + ((float)iy + 0.5f) * h, oz = -R + ((float)iz + 0.5f) * h;
    const float cx = gm[0] + ox, cy = gm[1] + oy, cz = gm[2] + oz;
    const float rr = ballR + rho;
    if (!(ox * ox + oy * oy + oz * oz <= rr * rr)) { if (lane == 0) hdr[cell] = 0ull; return; }
    // pass A: per-lane four nearest, then four rounds of warp arg-min
    const float INF = __int_as_float(0x7f800000);
    float e0 = INF, e1 = INF, e2 = INF, e3 = INF;
    int i0 = -1, i1 = -1, i2 = -1, i3 = -1;
    for (int j = lane; j < n_t; j += 32) {
        const int id = (int)tinv[j];
        const float4 q = __ldg(tgt + id);
        const float d = d2_rn(cx, cy, cz, q.x, q.y, q.z);
        {   // sorted insert into (e0 <= e1 <= e2 <= e3) with selects only
            const bool p0 = d < e0, p1 = d < e1, p2 = d < e2, p3 = d < e3;
            e3 = p2 ? e2 : (p3 ? d : e3); i3 = p2 ? i2 : (p3 ? id : i3);
            e2 = p1 ? e1 : (p2 ? d : e2); i2 = p1 ? i1 : (p2 ? id : i2);
            e1 = p0 ? e0 : (p1 ? d : e1); i1 = p0 ? i0 : (p1 ? id : i1);
            e0 = p0 ? d : e0; i0 = p0 ? id : i0;
        }
    }
    float4 cp[4]; float cd[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const unsigned m = __reduce_min_sync(KSS_FULL, __float_as_uint(e0));          // d2 >= 0: bits order as integers
        const unsigned bal = __ballot_sync(KSS_FULL, __float_as_uint(e0) == m);
        const int src = __ffs(bal) - 1;
        const int id = __shfl_sync(KSS_FULL, i0, src);
        cd[c] = (id >= 0) ? __uint_as_float(m) : INF;
        cp[c] = __ldg(tgt + (id >= 0 ? id : 0));
        if (lane == src) { e0 = e1; i0 = i1; e1 = e2; i1 = i2; e2 = e3; i2 = i3; e3 = INF; i3 = -1; }
    }
    const float thr = (sqrtf(cd[0]) + 2.0f * rho) * 1.0001f;
    const float thr2 = thr * thr;
    const float hh = h * 1.002f * 1.0001f;
    const float rr4 = 4.0f * rho * rho;
    // dominated by competitor c  <=>  (d - cd) - hh * s > 1e-5 * (d + cd + rr4), rearranged so that the per-competitor
    // part ca[c] = cd (1 + 1e-5) + 1e-5 * rr4 is computed once
    float ca[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) ca[c] = cd[c] * (1.0f + 1e-5f) + 1e-5f * rr4;
    // branch-free: a warp pays for its slowest lane anyway, and early exits only added divergence (-7 % build time)
    auto keep_test = [&](const float4& q, float d) -> bool {
        const float dl = d * (1.0f - 1e-5f);
        bool dom = false;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float s = fabsf(q.x - cp[c].x) + fabsf(q.y - cp[c].y) + fabsf(q.z - cp[c].z);
            dom |= (dl - ca[c]) - hh * s > 0.0f;
        }
        return d <= thr2 && !dom;
    };
    // pass B: count; pass C: ordered write
    int k = 0;
    for (int j0 = 0; j0 < n_t; j0 += 32) {
        const int j = j0 + lane;
        bool keep = false;
        if (j < n_t) { const float4 q = __ldg(tgt + (int)tinv[j]); keep = keep_test(q, d2_rn(cx, cy, cz, q.x, q.y, q.z)); }
        k += __popc(__ballot_sync(KSS_FULL, keep));
    }
    if (k == 0) { if (lane == 0) hdr[cell] = 0ull; return; }
    if (k > CG_EXT_MAX) { if (lane == 0) hdr[cell] = (cg_hdr_t)CG_TAG_HUGE << 60; return; }
    const int tot4 = (k + 3) & ~3;
    unsigned off = 0;
    if (lane == 0) off = atomicAdd(&cursor[p], (unsigned)tot4);
    off = __shfl_sync(KSS_FULL, off, 0);
    if ((size_t)off + (size_t)tot4 > CG_ARENA) { if (lane == 0) { atomicExch(&ok[p], 0); hdr[cell] = 0ull; } return; }
    unsigned short* out = arena + off;
    int w = 0, first = 0;
    for (int j0 = 0; j0 < n_t; j0 += 32) {
        const int j = j0 + lane;
        bool keep = false; int id = 0;
        if (j < n_t) { id = (int)tinv[j]; const float4 q = __ldg(tgt + id); keep = keep_test(q, d2_rn(cx, cy, cz, q.x, q.y, q.z)); }
        const unsigned bal = __ballot_sync(KSS_FULL, keep);
        if (keep) out[w + __popc(bal & ((1u << lane) - 1u))] = (unsigned short)id;
        if (w == 0 && bal) first = __shfl_sync(KSS_FULL, id, __ffs(bal) - 1);
        w += __popc(bal);
    }
    if (lane < tot4 - k) out[k + lane] = (unsigned short)first;
    if (lane == 0) hdr[cell] = ((cg_hdr_t)CG_TAG_EXT << 60) | ((cg_hdr_t)(unsigned)k << 32) | (cg_hdr_t)(off >> 2);
}

// =============================================================== cg_level_kernel
// Builds one level of the candidate grid (kss_cg.cuh).  One thread per child cell; the 8 children
// of a parent sit in adjacent lanes, so parent-list and target reads are broadcasts (4 distinct
// addresses per warp).  Pass A: the four targets nearest to the centre (the first is p_c, d_c^2).  Pass B: keep p iff
// d2 <= ((d_c + 2 rho)(1+1e-4))^2 and none of the four dominates p over the whole cell; remembered as a bit mask
// (lists <= 64) so pass C only re-reads indices.  Lists of up to CG_INLINE_MAX entries go into the 64-bit header,
// longer ones into the pair's arena (one atomicAdd per warp on its cursor).
// SPARSE 1: one thread per octant of every worklisted 32^3 cell (64^3 resolution); SPARSE 2: the same once more for the
// worklisted octants (128^3 resolution), whose headers live in the arena.
// Two-ended work lists: cells with short lists are pushed from the front, cells with long lists from the back, so that
// the warps of the refinement kernels hold parents of similar list length (the 8 siblings of a parent loop over its
// list; a warp holds 4 parents).  cnt[2p] = front count, cnt[2p + 1] = back count.
constexpr int CG_WL_SPLIT = 9;                // lists with >= this many entries go to the back
__device__ __forceinline__ int cg_wl_push(unsigned* cnt2, int cap, int k) {
    const bool back = k >= CG_WL_SPLIT;
    const unsigned slot = atomicAdd(&cnt2[back ? 1 : 0], 1u);
    const unsigned lim = back ? (unsigned)(cap * 3 / 8) : (unsigned)(cap * 5 / 8);
    if (slot >= lim) return -1;
    return back ? cap - 1 - (int)slot : (int)slot;
}
__device__ __forceinline__ int cg_wl_count(const unsigned* cnt2, int cap) {
    return (int)min(cnt2[0], (unsigned)(cap * 5 / 8)) + (int)min(cnt2[1], (unsigned)(cap * 3 / 8));
}
__device__ __forceinline__ int cg_wl_index(const unsigned* cnt2, int cap, int t) {     // t-th item -> position in the array
    const int nf = (int)min(cnt2[0], (unsigned)(cap * 5 / 8));
    return t < nf ? t : cap - 1 - (t - nf);
}
#ifndef CG_NCOMP
#define CG_NCOMP 4
#endif
#ifndef CG_REFINE_MIN2
#define CG_REFINE_MIN2 CG_REFINE_MIN       // octants (64^3) with at least this many candidates are cut once more
#endif
template <int SPARSE>
__device__ __forceinline__ void cg_build_cell(int tid, int ncells, int level, const float4* __restrict__ t_sorted,
                                              const unsigned short* __restrict__ t_inv_all, int cap_t,
                                              const int* __restrict__ cnt_t, int cap_tpad,
                                              const float* __restrict__ geom, cg_hdr_t* __restrict__ hdr_all,
                                              unsigned short* __restrict__ arena_all, unsigned* __restrict__ cursor, int* __restrict__ ok,
                                              unsigned short* __restrict__ wl_all, unsigned* __restrict__ wl_cnt,
                                              uint4* __restrict__ wl2_all, unsigned* __restrict__ wl2_cnt) {
    const int p = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int ng = SPARSE == 2 ? 4 * CG_NG : (SPARSE == 1 ? 2 * CG_NG : cg_ng(level));
    const int png = ng >> 1;
    const bool live = tid < ncells;
    const int n_t = cnt_t[p];
    const float4* __restrict__ tgt = t_sorted + (size_t)p * cap_tpad;
    // level 0 enumerates the targets in ORIGINAL index order (t_inv: original index -> Morton position) and every
    // level keeps its parent's order, so all lists are ascending in original index: a query can resolve equal
    // distances by "first one wins" (kss_cg.cuh)
    const unsigned short* __restrict__ tinv = t_inv_all + (size_t)p * cap_t;
    cg_hdr_t* hdr = hdr_all + (size_t)p * CG_HDR_TOTAL + cg_hdr_base(SPARSE ? CG_LEVELS - 1 : level);
    unsigned short* arena = arena_all + (size_t)p * CG_ARENA;
    const float* gm = geom + (size_t)p * 8;
    const float R = gm[3], ballR = gm[4];
    const float h = 2.0f * R / (float)ng;
    const float rho = h * 0.8660254f * 1.002f;
    const int child = tid & 7;
    int parent = tid >> 3;

    // the parent's list: inline in its header, in the arena, or "all targets" (level 0 / over-long parent)
    int ix = 0, iy = 0, iz = 0, m_p = 0;
    cg_hdr_t ph = 0ull;
    const unsigned short* plist = nullptr;
    bool p_inline = false;
    if (live) {
        bool from_parent = true;
        if (SPARSE == 2) {
            const uint4 e = wl2_all[(size_t)p * CG_WL2_CAP + cg_wl_index(wl2_cnt + 2 * p, CG_WL2_CAP, parent)];   // an octant of a refined cell
            parent = (int)e.x;                                               // arena offset of its header
            ix = 2 * (int)(e.y & 0xffffu) + (child & 1); iy = 2 * (int)(e.y >> 16) + ((child >> 1) & 1); iz = 2 * (int)e.z + (child >> 2);
            ph = *reinterpret_cast<const cg_hdr_t*>(arena + parent);
        } else if (SPARSE == 1) {
            parent = wl_all[(size_t)p * CG_WL_CAP + cg_wl_index(wl_cnt + 2 * p, CG_WL_CAP, parent)];   // a cell of the finest dense level
            const int px = parent & (CG_NG - 1), py = (parent / CG_NG) & (CG_NG - 1), pz = parent / (CG_NG * CG_NG);
            ix = 2 * px + (child & 1); iy = 2 * py + ((child >> 1) & 1); iz = 2 * pz + (child >> 2);
            ph = hdr[parent];
        } else if (level == 0) {
            ix = tid & 3; iy = (tid >> 2) & 3; iz = tid >> 4;
            m_p = n_t; from_parent = false;
        } else {
            const int lp = level + 1, pm = png - 1;                      // png = 2 << level is a power of two
            const int px = parent & pm, py = (parent >> lp) & pm, pz = parent >> (2 * lp);
            ix = 2 * px + (child & 1); iy = 2 * py + ((child >> 1) & 1); iz = 2 * pz + (child >> 2);
            ph = (hdr_all + (size_t)p * CG_HDR_TOTAL + cg_hdr_base(level - 1))[parent];
        }
        if (from_parent) {
            const unsigned tag = cg_tag(ph);
            if (tag >= 1u && tag <= (unsigned)CG_INLINE_MAX) { m_p = (int)tag; p_inline = true; }
            else if (tag == CG_TAG_EXT) { m_p = (int)cg_ext_count(ph); plist = arena + cg_offset(ph); }
            else if (tag == CG_TAG_HUGE) m_p = n_t;                        // over-long parent: all points
            else m_p = 0;                                                  // empty parent (outside the query ball)
        }
    }
    // (a parent with an inline list is never iterated: dense levels copy its header, sparse levels only refine
    // external lists)
    const unsigned short* __restrict__ csrc = plist ? plist : tinv;      // level 0 / over-long parent: all targets, original order
    auto cand = [&](int j) -> int { return (int)csrc[j]; };
    const int cell = ix + ng * (iy + ng * iz);
    // A parent list that already fits a header costs a query the same five branch-free evaluations however much a
    // child would shorten it: the children of such a cell just copy its header.  (cand(child) is a subset of
    // cand(parent), so the parent's list is valid everywhere inside it.)
    // (The lane still takes part in the warp-wide allocation scan below, with nothing to allocate.)
    const bool copy = !SPARSE && p_inline;
    if (p_inline) m_p = 0;
    const float ox = -R + ((float)ix + 0.5f) * h, oy = -R + ((float)iy + 0.5f) * h, oz = -R + ((float)iz + 0.5f) * h;
    const float cx = gm[0] + ox, cy = gm[1] + oy, cz = gm[2] + oz;
    const float rr = ballR + rho;
    const bool inball = ox * ox + oy * oy + oz * oz <= rr * rr;
    if (!inball) m_p = 0;                                                  // outside the query ball: empty list

    // ---- pass A: the four targets nearest to the centre; ties do not matter here (any target is a valid competitor)
    float e0 = __int_as_float(0x7f800000), e1 = e0, e2 = e0, e3 = e0;
    int i0 = -1, i1 = -1, i2 = -1, i3 = -1;
    for (int j = 0; j < m_p; ++j) {
        const int id = cand(j);
        const float4 q = __ldg(tgt + id);
        const float d = d2_rn(cx, cy, cz, q.x, q.y, q.z);
        {   // sorted insert into (e0 <= e1 <= e2 <= e3) with selects only
            const bool p0 = d < e0, p1 = d < e1, p2 = d < e2, p3 = d < e3;
            e3 = p2 ? e2 : (p3 ? d : e3); i3 = p2 ? i2 : (p3 ? id : i3);
            e2 = p1 ? e1 : (p2 ? d : e2); i2 = p1 ? i1 : (p2 ? id : i2);
            e1 = p0 ? e0 : (p1 ? d : e1); i1 = p0 ? i0 : (p1 ? id : i1);
            e0 = p0 ? d : e0; i0 = p0 ? id : i0;
        }
    }
    const float mn = e0;
    const float thr = (sqrtf(mn) + 2.0f * rho) * 1.0001f;
    const float thr2 = thr * thr;
    // dominance test against a competitor p': g(x) = |x-p|^2 - |x-p'|^2 is linear in x, so its minimum over the
    // cell is g(c) - h * (|dx|+|dy|+|dz|), d = p - p'.  If that is > 0 (with a margin far above fp32 rounding of
    // any query's two distances) p' beats p everywhere in the cell: p can never be a nearest neighbour of a query
    // in this cell, nor tie with one.  Competitors: the (up to) four targets nearest to the centre.
    const int ci[4] = {i0, i1, i2, i3};
    const float ce[4] = {e0, e1, e2, e3};
    float4 cp[4]; float cd[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const bool have = ci[c] >= 0;
        cp[c] = __ldg(tgt + (have ? ci[c] : 0));
        cd[c] = have ? ce[c] : __int_as_float(0x7f800000);                 // +inf: never dominates
    }
    const float hh = h * 1.002f * 1.0001f;
    const float rr4 = 4.0f * rho * rho;
    // dominated by competitor c  <=>  (d - cd) - hh * s > 1e-5 * (d + cd + rr4), rearranged so that the per-competitor
    // part ca[c] = cd (1 + 1e-5) + 1e-5 * rr4 is computed once
    float ca[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) ca[c] = cd[c] * (1.0f + 1e-5f) + 1e-5f * rr4;
    // branch-free: a warp pays for its slowest lane anyway, and early exits only added divergence (-7 % build time)
    auto keep_test = [&](const float4& q, float d) -> bool {
        const float dl = d * (1.0f - 1e-5f);
        bool dom = false;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float s = fabsf(q.x - cp[c].x) + fabsf(q.y - cp[c].y) + fabsf(q.z - cp[c].z);
            dom |= (dl - ca[c]) - hh * s > 0.0f;
        }
        return d <= thr2 && !dom;
    };
    // ---- pass B: count (and remember) the candidates; the first CG_INLINE_MAX of them are packed on the way
    unsigned long long mask = 0ull;
    cg_hdr_t packed = 0ull;
    int k = 0, first = 0;
    for (int j = 0; j < m_p; ++j) {
        const int id = cand(j);
        const float4 q = __ldg(tgt + id);
        if (keep_test(q, d2_rn(cx, cy, cz, q.x, q.y, q.z))) {
            if (j < 64) mask |= 1ull << j;
            if (k == 0) first = id;
            if (k < CG_INLINE_MAX) packed |= (cg_hdr_t)(unsigned)id << (11 * k);
            ++k;
        }
    }
    const bool huge = k > CG_EXT_MAX;
    const bool ext = k > CG_INLINE_MAX && !huge;
    const int tot4 = ext ? ((k + 3) & ~3) : 0;
    // warp exclusive scan of the arena space needed, one atomicAdd per warp; a refined cell also needs 8 child headers
    // (64 bytes = 32 entries), reserved by its octant 0
    const int need = tot4 + ((SPARSE && live && child == 0) ? 32 : 0);
    int incl = need;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(KSS_FULL, incl, o); if (lane >= o) incl += y; }
    const int wtot = __shfl_sync(KSS_FULL, incl, 31);
    unsigned base = 0;
    if (lane == 0 && wtot > 0) base = atomicAdd(&cursor[p], (unsigned)wtot);
    base = __shfl_sync(KSS_FULL, base, 0);
    unsigned off = base + (unsigned)(incl - need);
    const unsigned blk = __shfl_sync(KSS_FULL, off, lane & ~7);           // the child-header block of this thread's parent
    if ((size_t)base + (size_t)wtot > CG_ARENA) {
        if (SPARSE) return;                                               // no room: the cells keep their unrefined lists
        if (lane == 0) atomicExch(&ok[p], 0);                             // arena exhausted: pair falls back
        if (live) hdr[cell] = 0ull;
        return;
    }
    if (!live) return;
    if (SPARSE && child == 0) off += 32u;
    cg_hdr_t me;
    if (huge) me = (cg_hdr_t)CG_TAG_HUGE << 60;
    else if (ext) me = ((cg_hdr_t)CG_TAG_EXT << 60) | ((cg_hdr_t)(unsigned)k << 32) | (cg_hdr_t)(off >> 2);
    else me = ((cg_hdr_t)(unsigned)k << 60) | packed;                     // k = 0: empty
    if (copy) me = ph;
    if (SPARSE) {
        reinterpret_cast<cg_hdr_t*>(arena + blk)[child] = me;
        if (child == 0) {
            const cg_hdr_t mark = ((cg_hdr_t)CG_TAG_REFINED << 60) | (cg_hdr_t)(blk >> 2);
            if (SPARSE == 2) *reinterpret_cast<cg_hdr_t*>(arena + parent) = mark; else hdr[parent] = mark;
        }
        if (SPARSE == 1 && wl2_all != nullptr && !huge && k >= CG_REFINE_MIN2) {
            const int slot = cg_wl_push(wl2_cnt + 2 * p, CG_WL2_CAP, k);
            if (slot >= 0)
                wl2_all[(size_t)p * CG_WL2_CAP + slot] = make_uint4(blk + 4u * (unsigned)child, (unsigned)ix | ((unsigned)iy << 16), (unsigned)iz, 0u);
        }
    } else {
        hdr[cell] = me;
        if (level == CG_LEVELS - 1 && wl_all != nullptr && !huge && k >= CG_REFINE_MIN) {
            const int slot = cg_wl_push(wl_cnt + 2 * p, CG_WL_CAP, k);
            if (slot >= 0) wl_all[(size_t)p * CG_WL_CAP + slot] = (unsigned short)cell;
        }
    }
    if (!ext) return;
    // ---- pass C: write the external list, padded to a multiple of 4 with a valid candidate
    unsigned short* out = arena + off;
    int w = 0;
    if (m_p <= 64) {
        while (mask) {
            const int j = __ffsll((long long)mask) - 1;
            mask &= mask - 1ull;
            out[w++] = (unsigned short)cand(j);
        }
    } else {
        for (int j = 0; j < m_p; ++j) {
            const int id = cand(j);
            const float4 q = __ldg(tgt + id);
            if (keep_test(q, d2_rn(cx, cy, cz, q.x, q.y, q.z))) out[w++] = (unsigned short)id;
        }
    }
    for (; w < tot4; ++w) out[w] = (unsigned short)first;
}

// =============================================================== cg_level_coop_kernel
// Dense levels 1..3 with EIGHT LANES PER CELL: the parent lists of these levels are long (tens to hundreds of entries)
// and, with one thread per cell, a warp mixed cells outside the query ball, short and long lists (11 of 32 lanes
// active in ncu).  Here a warp holds four sibling cells -- same parent, same list, same trip count -- and the eight
// lanes of a cell stride over the list.  Same rules and the same (ascending original index) output order as
// cg_build_cell: kept candidates are remembered as per-lane bit masks and written with ballot-ordered positions.
// Collectives inside the per-cell loops use the cell's own 8-lane mask (trip counts differ between the four cells).
__global__ void __launch_bounds__(256)
cg_level_coop_kernel(int level, const float4* __restrict__ t_sorted, const unsigned short* __restrict__ t_inv_all, int cap_t,
                     const int* __restrict__ cnt_t, int cap_tpad,
                     const float* __restrict__ geom, cg_hdr_t* __restrict__ hdr_all,
                     unsigned short* __restrict__ arena_all, unsigned* __restrict__ cursor, int* __restrict__ ok,
                     unsigned short* __restrict__ wl_all, unsigned* __restrict__ wl_cnt) {
    const int p = blockIdx.y;
    const int lane = threadIdx.x & 31, sub = lane & 7, gsh = lane & 24;
    const unsigned gmask = 0xffu << gsh;
    const int ng = cg_ng(level), png = ng >> 1;
    const int ncells = ng * ng * ng;
    const int tid = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;     // = parent * 8 + child
    if (tid - (lane >> 3) >= ncells) return;                           // whole warp out of range (ncells is a multiple of 4)
    const bool live = tid < ncells;
    const int n_t = cnt_t[p];
    const float4* __restrict__ tgt = t_sorted + (size_t)p * cap_tpad;
    const unsigned short* __restrict__ tinv = t_inv_all + (size_t)p * cap_t;
    cg_hdr_t* hdr = hdr_all + (size_t)p * CG_HDR_TOTAL + cg_hdr_base(level);
    unsigned short* arena = arena_all + (size_t)p * CG_ARENA;
    const float* gm = geom + (size_t)p * 8;
    const float R = gm[3], ballR = gm[4];
    const float h = 2.0f * R / (float)ng;
    const float rho = h * 0.8660254f * 1.002f;
    const int child = tid & 7, parent = tid >> 3;
    const int lp = level + 1, pm = png - 1;                            // png = 2 << level is a power of two
    const int px = parent & pm, py = (parent >> lp) & pm, pz = parent >> (2 * lp);
    const int ix = 2 * px + (child & 1), iy = 2 * py + ((child >> 1) & 1), iz = 2 * pz + (child >> 2);
    const int cell = ix + ng * (iy + ng * iz);
    cg_hdr_t ph = 0ull;
    int m_p = 0;
    const unsigned short* plist = nullptr;
    bool p_inline = false;
    if (live) {
        ph = (hdr_all + (size_t)p * CG_HDR_TOTAL + cg_hdr_base(level - 1))[parent];
        const unsigned tag = cg_tag(ph);
        if (tag >= 1u && tag <= (unsigned)CG_INLINE_MAX) { m_p = (int)tag; p_inline = true; }
        else if (tag == CG_TAG_EXT) { m_p = (int)cg_ext_count(ph); plist = arena + cg_offset(ph); }
        else if (tag == CG_TAG_HUGE) m_p = n_t;
    }
    // (a parent with an inline list is never iterated: dense levels copy its header, sparse levels only refine
    // external lists)
    const unsigned short* __restrict__ csrc = plist ? plist : tinv;      // level 0 / over-long parent: all targets, original order
    auto cand = [&](int j) -> int { return (int)csrc[j]; };
    const bool copy = p_inline;                                        // a list that fits a header is not refined (see cg_build_cell)
    if (copy) m_p = 0;
    const float ox = -R + ((float)ix + 0.5f) * h, oy = -R + ((float)iy + 0.5f) * h, oz = -R + ((float)iz + 0.5f) * h;
    const float cx = gm[0] + ox, cy = gm[1] + oy, cz = gm[2] + oz;
    const float rr = ballR + rho;
    if (!(ox * ox + oy * oy + oz * oz <= rr * rr)) m_p = 0;             // outside the query ball: empty list
    const float INF = __int_as_float(0x7f800000);

    // ---- pass A: per-lane four nearest to the centre, merged over the cell's eight lanes
    float e0 = INF, e1 = INF, e2 = INF, e3 = INF;
    int i0 = -1, i1 = -1, i2 = -1, i3 = -1;
    for (int j = sub; j < m_p; j += 8) {
        const int id = cand(j);
        const float4 q = __ldg(tgt + id);
        const float d = d2_rn(cx, cy, cz, q.x, q.y, q.z);
        {   // sorted insert into (e0 <= e1 <= e2 <= e3) with selects only
            const bool p0 = d < e0, p1 = d < e1, p2 = d < e2, p3 = d < e3;
            e3 = p2 ? e2 : (p3 ? d : e3); i3 = p2 ? i2 : (p3 ? id : i3);
            e2 = p1 ? e1 : (p2 ? d : e2); i2 = p1 ? i1 : (p2 ? id : i2);
            e1 = p0 ? e0 : (p1 ? d : e1); i1 = p0 ? i0 : (p1 ? id : i1);
            e0 = p0 ? d : e0; i0 = p0 ? id : i0;
        }
    }
    float4 cp[4]; float cd[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        float m = e0;
        m = fminf(m, __shfl_xor_sync(gmask, m, 1));
        m = fminf(m, __shfl_xor_sync(gmask, m, 2));
        m = fminf(m, __shfl_xor_sync(gmask, m, 4));
        const unsigned bal = __ballot_sync(gmask, e0 == m);
        const int src = __ffs(bal) - 1;                                 // bal != 0: the minimum is somebody's e0
        const int id = __shfl_sync(gmask, i0, src);
        cd[c] = id >= 0 ? m : INF;                                      // +inf: never dominates
        cp[c] = __ldg(tgt + (id >= 0 ? id : 0));
        if (lane == src) { e0 = e1; i0 = i1; e1 = e2; i1 = i2; e2 = e3; i2 = i3; e3 = INF; i3 = -1; }
    }
    const float thr = (sqrtf(cd[0]) + 2.0f * rho) * 1.0001f;
    const float thr2 = thr * thr;
    const float hh = h * 1.002f * 1.0001f;
    const float rr4 = 4.0f * rho * rho;
    // dominated by competitor c  <=>  (d - cd) - hh * s > 1e-5 * (d + cd + rr4), rearranged so that the per-competitor
    // part ca[c] = cd (1 + 1e-5) + 1e-5 * rr4 is computed once
    float ca[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) ca[c] = cd[c] * (1.0f + 1e-5f) + 1e-5f * rr4;
    // branch-free: a warp pays for its slowest lane anyway, and early exits only added divergence (-7 % build time)
    auto keep_test = [&](const float4& q, float d) -> bool {
        const float dl = d * (1.0f - 1e-5f);
        bool dom = false;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float s = fabsf(q.x - cp[c].x) + fabsf(q.y - cp[c].y) + fabsf(q.z - cp[c].z);
            dom |= (dl - ca[c]) - hh * s > 0.0f;
        }
        return d <= thr2 && !dom;
    };
    // ---- pass B: this lane's kept candidates as a bit mask (bit jj <-> list position jj * 8 + sub; lists beyond
    //      512 entries re-run the test in pass C), the cell's count by ballots
    unsigned long long mask = 0ull;
    int k = 0;
    for (int j0 = 0, jj = 0; j0 < m_p; j0 += 8, ++jj) {
        const int j = j0 + sub;
        bool keep = false;
        if (j < m_p) { const float4 q = __ldg(tgt + cand(j)); keep = keep_test(q, d2_rn(cx, cy, cz, q.x, q.y, q.z)); }
        if (keep && jj < 64) mask |= 1ull << jj;
        k += __popc(__ballot_sync(gmask, keep));
    }
    const bool huge = k > CG_EXT_MAX;
    const bool ext = k > CG_INLINE_MAX && !huge;
    const int tot4 = ext ? ((k + 3) & ~3) : 0;
    // arena space: the cells' leaders take part in a warp scan, one atomicAdd per warp
    const int need = (sub == 0 && live) ? tot4 : 0;
    int incl = need;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(KSS_FULL, incl, o); if (lane >= o) incl += y; }
    const int wtot = __shfl_sync(KSS_FULL, incl, 31);
    unsigned base = 0;
    if (lane == 0 && wtot > 0) base = atomicAdd(&cursor[p], (unsigned)wtot);
    base = __shfl_sync(KSS_FULL, base, 0);
    const unsigned off = __shfl_sync(KSS_FULL, base + (unsigned)(incl - need), gsh);      // the leader's offset
    if ((size_t)base + (size_t)wtot > CG_ARENA) {
        if (lane == 0) atomicExch(&ok[p], 0);                             // arena exhausted: pair falls back
        if (live && sub == 0) hdr[cell] = 0ull;
        return;
    }
    // ---- pass C: ordered positions from ballots; inline lists are OR-ed together over the lanes
    cg_hdr_t packed = 0ull;
    int first = 0;
    if (!huge && k > 0) {
        unsigned short* out = arena + off;
        int w = 0;
        for (int j0 = 0, jj = 0; j0 < m_p; j0 += 8, ++jj) {
            const int j = j0 + sub;
            bool keep = false; int id = 0;
            if (j < m_p) {
                id = cand(j);
                if (jj < 64) keep = (mask >> jj) & 1ull;
                else { const float4 q = __ldg(tgt + id); keep = keep_test(q, d2_rn(cx, cy, cz, q.x, q.y, q.z)); }
            }
            const unsigned bal = __ballot_sync(gmask, keep) >> gsh;
            const int pos = w + __popc(bal & ((1u << sub) - 1u));
            if (keep) {
                if (ext) out[pos] = (unsigned short)id;
                else packed |= (cg_hdr_t)(unsigned)id << (11 * pos);
            }
            if (w == 0 && bal) first = __shfl_sync(gmask, id, gsh + __ffs(bal) - 1);
            w += __popc(bal);
        }
        if (ext) { if (sub < tot4 - k) out[k + sub] = (unsigned short)first; }
        else {
            packed |= __shfl_xor_sync(gmask, packed, 1);
            packed |= __shfl_xor_sync(gmask, packed, 2);
            packed |= __shfl_xor_sync(gmask, packed, 4);
        }
    }
    if (!live || sub != 0) return;
    cg_hdr_t me;
    if (huge) me = (cg_hdr_t)CG_TAG_HUGE << 60;
    else if (ext) me = ((cg_hdr_t)CG_TAG_EXT << 60) | ((cg_hdr_t)(unsigned)k << 32) | (cg_hdr_t)(off >> 2);
    else me = ((cg_hdr_t)(unsigned)k << 60) | packed;                     // k = 0: empty
    if (copy) me = ph;
    hdr[cell] = me;
    if (level == CG_LEVELS - 1 && wl_all != nullptr && !huge && !copy && k >= CG_REFINE_MIN) {
        const int slot = cg_wl_push(wl_cnt + 2 * p, CG_WL_CAP, k);
        if (slot >= 0) wl_all[(size_t)p * CG_WL_CAP + slot] = (unsigned short)cell;
    }
}

// dense levels: one pass over ng^3 cells (the grid covers them exactly); sparse levels: a fixed, small grid per pair
// loops over the worklist (its length is only known on the device)
#ifndef CG_MIN_CTAS
#define CG_MIN_CTAS 4
#endif
template <int SPARSE>
__global__ void __launch_bounds__(256, CG_MIN_CTAS)
cg_level_kernel(int level, const float4* __restrict__ t_sorted, const unsigned short* __restrict__ t_inv_all, int cap_t,
                const int* __restrict__ cnt_t, int cap_tpad,
                const float* __restrict__ geom, cg_hdr_t* __restrict__ hdr_all,
                unsigned short* __restrict__ arena_all, unsigned* __restrict__ cursor, int* __restrict__ ok,
                unsigned short* __restrict__ wl_all, unsigned* __restrict__ wl_cnt,
                uint4* __restrict__ wl2_all, unsigned* __restrict__ wl2_cnt) {
    const int p = blockIdx.y;
    const int ng = cg_ng(level);
    const int ncells = SPARSE == 2 ? 8 * cg_wl_count(wl2_cnt + 2 * p, CG_WL2_CAP)
                     : (SPARSE == 1 ? 8 * cg_wl_count(wl_cnt + 2 * p, CG_WL_CAP) : ng * ng * ng);
    const int lane = threadIdx.x & 31;
    for (int tid = blockIdx.x * blockDim.x + threadIdx.x; tid - lane < ncells; tid += gridDim.x * blockDim.x)
        cg_build_cell<SPARSE>(tid, ncells, level, t_sorted, t_inv_all, cap_t, cnt_t, cap_tpad, geom, hdr_all, arena_all, cursor, ok,
                              wl_all, wl_cnt, wl2_all, wl2_cnt);
}

// =============================================================== sweep_kernel
// grid (G*G, P): CTA (i,j) of pair p applies Rx(i), Ry(j) once per point and loops Rz(k),
// searching the exact NN of every rotated point (narrowed to float) in the Morton-tiled
// target.  rbuf[p][orig][h] receives sqrtf(d2) as a float (widened exactly later) (score modes AVE/DIFF) or (double)d2
// (MAX); the serial sums are taken by sweep_finalize_kernel.
#ifndef SWEEP_U
#define SWEEP_U 3
#endif
#ifndef SWEEP_MIN_CTAS
#define SWEEP_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(256, SWEEP_MIN_CTAS)
sweep_kernel(const double* __restrict__ s_al, const int* __restrict__ cnt_s, int cap_s,
             const unsigned short* __restrict__ s_perm,
             const float4* __restrict__ t_sorted, const float* __restrict__ t_box,
             const int* __restrict__ cnt_t, int cap_tpad,
             const double* __restrict__ trig_accum /* [G][2] cos,sin */, int G, int score_mode,
             float* __restrict__ rbuf, int hpad, int ij_lo /* first (i, j) pair of this launch: hypothesis slabs */,
             const float* __restrict__ cg_geom, const cg_hdr_t* __restrict__ cg_hdr,
             const unsigned short* __restrict__ cg_arena, const int* __restrict__ cg_ok) {
    extern __shared__ unsigned char smem_raw[];
    float4* tgt = reinterpret_cast<float4*>(smem_raw);
    const int p = blockIdx.y;
    const int n_t = cnt_t[p];
    const int npad = (n_t + TILE - 1) / TILE * TILE;
    float* box = reinterpret_cast<float*>(tgt + npad);
    const float4* gts = t_sorted + (size_t)p * cap_tpad;
    for (int j = threadIdx.x; j < npad; j += blockDim.x) tgt[j] = gts[j];
    const float* gtb = t_box + (size_t)p * 6 * MAX_TILES;
    for (int j = threadIdx.x; j < 6 * MAX_TILES; j += blockDim.x) box[j] = gtb[j];
    __syncthreads();
    TileView tv{tgt, box, npad / TILE};
    const bool use_cg = cg_hdr != nullptr && cg_ok[p] != 0;
    CgView cg{};
    if (use_cg) cg = cg_view(cg_geom, cg_hdr, cg_arena, cg_ok, p);

    const int gi = (blockIdx.x + ij_lo) / G, gj = (blockIdx.x + ij_lo) % G;
    const double ci = trig_accum[2 * gi], si = trig_accum[2 * gi + 1];
    const double cj = trig_accum[2 * gj], sj = trig_accum[2 * gj + 1];
    const int n_s = cnt_s[p];
    const unsigned short* perm = s_perm + (size_t)p * cap_s;
    const double* sa = s_al + (size_t)p * cap_s * 3;
    float* rb = rbuf + (size_t)p * cap_s * hpad;
    const int hbase = (gi * G + gj) * G;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

    if (score_mode == KSS_SCORE_VOXEL) {
        // voxel-occupancy score (see KSS_SCORE_VOXEL in the header): 1 for a source point in an empty voxel, else 0
        constexpr int NVMAX = 32;
        const int NV = min(NVMAX, max(4, (int)floor(sqrt((double)n_t / 16.0))));        // ~16 target points per occupied face voxel
        __shared__ unsigned s_occ[NVMAX * NVMAX * NVMAX / 32];
        __shared__ unsigned s_max;
        for (int j = threadIdx.x; j < NVMAX * NVMAX * NVMAX / 32; j += blockDim.x) s_occ[j] = 0u;
        if (threadIdx.x == 0) s_max = 0u;
        __syncthreads();
        float m = 0.0f;
        for (int j = threadIdx.x; j < n_t; j += blockDim.x) m = fmaxf(m, fmaxf(fabsf(tgt[j].x), fmaxf(fabsf(tgt[j].y), fabsf(tgt[j].z))));
        atomicMax(&s_max, __float_as_uint(m));                     // (non-negative floats order as their bits)
        __syncthreads();
        const float E = __fmul_rn(1.1f, __uint_as_float(s_max));
        const float inv = __fdiv_rn((float)NV, __fmul_rn(2.0f, E));
        auto vox = [&](float v) { return (int)floorf(__fmul_rn(__fadd_rn(v, E), inv)); };
        if (E > 0.0f)
            for (int j = threadIdx.x; j < n_t; j += blockDim.x) {
                const int b = (vox(tgt[j].z) * NV + vox(tgt[j].y)) * NV + vox(tgt[j].x);
                atomicOr(&s_occ[b >> 5], 1u << (b & 31));
            }
        __syncthreads();
        for (int o = threadIdx.x; o < n_s; o += blockDim.x) {
            double x = sa[3 * o], y = sa[3 * o + 1], z = sa[3 * o + 2];
            rot_x(ci, si, y, z);
            rot_y(cj, sj, x, z);
            for (int k = 0; k < G; ++k) {
                double xx = x, yy = y;
                rot_z(trig_accum[2 * k], trig_accum[2 * k + 1], xx, yy);
                const int ix = vox((float)xx), iy = vox((float)yy), iz = vox((float)z);
                bool hit = false;
                if (E > 0.0f && ix >= 0 && ix < NV && iy >= 0 && iy < NV && iz >= 0 && iz < NV) {
                    const int b = (iz * NV + iy) * NV + ix;
                    hit = (s_occ[b >> 5] >> (b & 31)) & 1u;
                }
                rb[(size_t)o * hpad + hbase + k] = hit ? 0.0f : 1.0f;
            }
        }
        return;
    }

    for (int base = warp * 32; base < n_s; base += nwarps * 32) {
        const int jpos = base + lane;
        const bool valid = jpos < n_s;
        const int o = perm[valid ? jpos : n_s - 1];
        double x = sa[3 * o], y = sa[3 * o + 1], z = sa[3 * o + 2];
        rot_x(ci, si, y, z);
        rot_y(cj, sj, x, z);
        float* ro = rb + (size_t)o * hpad + hbase;
        if (use_cg) {
            // three z-rotations per batch: their grid look-ups overlap (cg_query_batch)
            constexpr int U = SWEEP_U;
            for (int k0 = 0; k0 < G; k0 += U) {
                float qx[U], qy[U], qz[U];
                unsigned long long key[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int k = min(k0 + u, G - 1);
                    const double ck = trig_accum[2 * k], sk = trig_accum[2 * k + 1];
                    double xx = x, yy = y;
                    rot_z(ck, sk, xx, yy);
                    qx[u] = (float)xx; qy[u] = (float)yy; qz[u] = (float)z;   // :440-442 narrowing
                }
                cg_query_batch<U, false>(cg, tgt, n_t, qx, qy, qz, key);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float d2 = __uint_as_float((unsigned)(key[u] >> 32));
                    // float sqrt (:444); the reference widens it to double before summing: sweep_finalize_kernel does
                    const float r = score_mode == 1 ? d2 : __fsqrt_rn(d2);
                    if (valid && k0 + u < G) ro[k0 + u] = r;
                }
            }
        } else {
            for (int k = 0; k < G; ++k) {
                const double ck = trig_accum[2 * k], sk = trig_accum[2 * k + 1];
                double xx = x, yy = y;
                rot_z(ck, sk, xx, yy);
                const float qx = (float)xx, qy = (float)yy, qz = (float)z;   // :440-442 narrowing
                const unsigned long long key = warp_nn<false>(tv, qx, qy, qz);
                const float d2 = __uint_as_float((unsigned)(key >> 32));
                const float r = score_mode == 1 ? d2 : __fsqrt_rn(d2);
                if (valid) ro[k] = r;
            }
        }
    }
}

// =============================================================== sweep_finalize_kernel
// One CTA per pair.  Thread h sums rbuf[p][0..n_s)[h] serially in index order (the
// reference's `distanceSum = distanceSum + distance_i`), then argmin (first strict <,
// errorT = 9999) and the clamped 5x5x5 local-minimum test in (i,j,k) loop order.
__global__ void __launch_bounds__(1024)
sweep_finalize_kernel(const float* __restrict__ rbuf, const int* __restrict__ cnt_s, int cap_s, int hpad,
                      int G, int score_mode, int h_lo, int h_hi /* hypotheses this launch sums */,
                      int phases /* 1: the sums, 2: argmin + local minima over the full grid in `value`, 3: both */,
                      double* __restrict__ value /* [P][hpad] */, int* __restrict__ best_h /* [P] */,
                      int* __restrict__ minima /* [P][hpad] */, int* __restrict__ n_minima /* [P] */) {
    extern __shared__ unsigned char smem_raw[];
    double* val = reinterpret_cast<double*>(smem_raw);                 // [H]
    unsigned char* flag = reinterpret_cast<unsigned char*>(val + G * G * G);
    __shared__ unsigned long long bestkey;
    const int p = blockIdx.x;
    const int H = G * G * G;
    const int n = cnt_s[p];
    const float* rb = rbuf + (size_t)p * cap_s * hpad;
    if (threadIdx.x == 0) bestkey = 0xffffffffffffffffull;
    for (int h = h_lo + threadIdx.x; (phases & 1) && h < h_hi; h += blockDim.x) {
        double sum = 0.0, dmax = -9999.0;
        // 16 independent loads in flight per thread, then the strictly ordered adds
        int i = 0;
        for (; i + 16 <= n; i += 16) {
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = __ldg(rb + (size_t)(i + u) * hpad + h);
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const double r = (double)v[u];
                if (score_mode == 1) { if (r > dmax) dmax = r; }
                else { sum = __dadd_rn(sum, r); if (dmax < r) dmax = r; }
            }
        }
        for (; i < n; ++i) {
            const double r = (double)rb[(size_t)i * hpad + h];
            if (score_mode == 1) { if (r > dmax) dmax = r; }
            else { sum = __dadd_rn(sum, r); if (dmax < r) dmax = r; }
        }
        double v;
        if (score_mode == 1) v = dmax;
        else if (score_mode == 2) v = __dsub_rn(dmax, __ddiv_rn(sum, (double)n));
        else v = __ddiv_rn(sum, (double)n);
        val[h] = v;
        value[(size_t)p * hpad + h] = v;
    }
    if (!(phases & 2)) return;
    if (phases == 2)                         // the grid was summed elsewhere (hypothesis slabs gathered from all ranks)
        for (int h = threadIdx.x; h < H; h += blockDim.x) val[h] = value[(size_t)p * hpad + h];
    __syncthreads();
    // first strict minimum below 9999 in loop order == smallest (value, h) pair
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        const double v = val[h];
        if (v < 9999.0) {
            // order-preserving map of a double to uint64
            unsigned long long u = (unsigned long long)__double_as_longlong(v);
            u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
            // 64-bit value key cannot also hold h: reduce value first, then index
            atomicMin(&bestkey, u);
        }
    }
    __syncthreads();
    __shared__ int besth;
    if (threadIdx.x == 0) besth = 0x7fffffff;
    __syncthreads();
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        const double v = val[h];
        if (v < 9999.0) {
            unsigned long long u = (unsigned long long)__double_as_longlong(v);
            u = (u >> 63) ? ~u : (u | 0x8000000000000000ull);
            if (u == bestkey) atomicMin(&besth, h);
        }
    }
    // local minima (initRegistration_kernel): reject iff centre > some neighbour
    for (int h = threadIdx.x; h < H; h += blockDim.x) {
        const int i = h / (G * G), j = (h / G) % G, k = h % G;
        const double c = val[h];
        bool ok = true;
        for (int ii = max(0, i - 2); ii <= min(G - 1, i + 2) && ok; ++ii)
            for (int jj = max(0, j - 2); jj <= min(G - 1, j + 2) && ok; ++jj)
                for (int kk = max(0, k - 2); kk <= min(G - 1, k + 2); ++kk)
                    if (c > val[(ii * G + jj) * G + kk]) { ok = false; break; }
        flag[h] = ok ? 1 : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) best_h[p] = besth == 0x7fffffff ? 0 : besth;
    if (threadIdx.x < 32) {                                  // ordered compaction by one warp
        int count = 0;
        int* out = minima + (size_t)p * hpad;
        for (int base = 0; base < H; base += 32) {
            const int h = base + threadIdx.x;
            const bool f = h < H && flag[h];
            const unsigned bal = __ballot_sync(KSS_FULL, f);
            if (f) out[count + __popc(bal & ((1u << threadIdx.x) - 1u))] = h;
            count += __popc(bal);
        }
        if (threadIdx.x == 0) n_minima[p] = count;
    }
}

// =============================================================== icp_small_kernel
// One CTA = one complete PCL-1.8.1 ICP run on a <=2048 x <=2048 pair:
// correspondences (exact NN, reject d2 > max^2), umeyama (two-pass float, CANON256 sums),
// in-place float transform, final = T_k * final, DefaultConvergenceCriteria, and the
// getFitnessScore pass.  No host round trip inside a run.
//
// mode 0: judge run      -> angles = accumulated loop values of best_h      (KSS_ICP.hpp:92-93)
// mode 1: hypothesis l   -> angles = index*6.3/step of minima[l], only if judge fitness > thr
// mode 2: explicit input -> src_f64 is used as is (kss_icp API, KSS_ICP.hpp:323-356)
#ifndef ICP_U
#define ICP_U 1
#endif
#ifndef ICP_MIN_CTAS
#define ICP_MIN_CTAS 4
#endif
// diagnostics (-DKSS_ICP_PHASE_TIMING): thread 0 adds the cycles since its previous tick to phase_cycles[ph]
#ifdef KSS_ICP_PHASE_TIMING
#define ICP_TICK(ph)                                                                                   \
    do {                                                                                               \
        if (a.phase_cycles != nullptr && threadIdx.x == 0) {                                           \
            const long long now_ = clock64();                                                          \
            atomicAdd(&a.phase_cycles[ph], (unsigned long long)(now_ - tick_));                        \
            tick_ = now_;                                                                              \
        }                                                                                              \
    } while (0)
#else
#define ICP_TICK(ph) do { } while (0)
#endif
// TRACE: the per-iteration records of kss_icp's trace argument (tests); compiled out of the batched launches
template <bool TRACE>
__global__ void __launch_bounds__(256, ICP_MIN_CTAS)
icp_small_kernel(IcpArgs a) {
#ifdef KSS_ICP_PHASE_TIMING
    long long tick_ = a.phase_cycles != nullptr ? clock64() : 0;
#endif
    extern __shared__ unsigned char smem_raw[];
    const int p = blockIdx.y;
    const int slot = blockIdx.x;
    const int n_s = a.cnt_s ? a.cnt_s[p] : a.cap_s;
    const int n_t = a.cnt_t[p];
    // mode 3 = one launch for everything: slot 0 is the judge run, slots >= 1 the hypothesis runs (not gated on
    // the judge's fitness -- the two are independent; select_kernel ignores them when the judge passes)
    const int mode = a.mode == 3 ? (slot == 0 ? 0 : 1) : a.mode;
    const int hslot = a.mode == 3 ? slot - 1 : slot;
    const int run = p * a.runs_per_pair + (mode == 1 ? 1 + hslot : 0);
    // A pair with more local minima than hypothesis slots: the CTA of slot s runs the hypotheses s, s + slots,
    // s + 2 slots, ... one after the other and keeps the lexicographic minimum (fitness, l) among those with
    // 0 <= fitness < 9999 -- exactly what the serial scan `ri < Q && ri >= 0`, Q = 9999 (KSS_ICP.hpp:100-116) keeps.
    int hs = hslot;                                   // hypothesis (angleList index) of the current run

    __shared__ double ang_c[3], ang_s[3];             // cos / sin of the run's start rotation (shared: register cap)
    if (threadIdx.x < 3) { ang_c[threadIdx.x] = 1.0; ang_s[threadIdx.x] = 0.0; }
    __syncthreads();
    if (mode == 1) {
        const bool active = a.mode == 3 || a.judge_thr < 0.0 || a.run_fit[p * a.runs_per_pair] > a.judge_thr;
        hs = hslot * max(a.hyp_world, 1) + a.hyp_rank; // hypothesis-sharded: this rank runs l = rank (mod world)
        if (!active || hs >= a.n_minima[p] || hslot >= a.runs_per_pair - 1) return;
    } else if (mode == 0) {
        if (threadIdx.x < 3) {
            const int h = a.best_h[p], k = threadIdx.x;
            const int id = k == 0 ? h / (a.G * a.G) : k == 1 ? (h / a.G) % a.G : h % a.G;
            ang_c[k] = a.trig_accum[2 * id]; ang_s[k] = a.trig_accum[2 * id + 1];
        }
    }

    const int npad = (n_t + TILE - 1) / TILE * TILE;
    float4* tgt = reinterpret_cast<float4*>(smem_raw);
    float* box = reinterpret_cast<float*>(tgt + npad);
    float* cur = box + 6 * MAX_TILES;                 // SoA by ORIGINAL source index: x[n_s] y[n_s] z[n_s]
    float* d2s = cur + 3 * n_s;                       // [n_s]
    unsigned short* perm = reinterpret_cast<unsigned short*>(d2s + n_s);   // [n_s] Morton pos -> original
    unsigned short* mpos = perm + n_s;                // [n_s] matched target position, 0xffff = rejected
    unsigned short* tinv = mpos + n_s;                // [n_t] original target index -> Morton position
    __shared__ float red[16];
    __shared__ double redd;
    __shared__ float partF[8 * 16];
    __shared__ double partD[8];
    __shared__ float Tk[16], fin[16];
    __shared__ int kept, done;
    __shared__ double prev_mse;

    {
        const float4* gts = a.t_sorted + (size_t)p * a.cap_tpad;
        for (int j = threadIdx.x; j < npad; j += blockDim.x) tgt[j] = gts[j];
        const float* gtb = a.t_box + (size_t)p * 6 * MAX_TILES;
        for (int j = threadIdx.x; j < 6 * MAX_TILES; j += blockDim.x) box[j] = gtb[j];
        const unsigned short* gp = a.s_perm + (size_t)p * a.cap_s;
        for (int j = threadIdx.x; j < n_s; j += blockDim.x) perm[j] = gp[j];
        const unsigned short* gi = a.t_inv + (size_t)p * a.cap_t;
        for (int j = threadIdx.x; j < n_t; j += blockDim.x) tinv[j] = gi[j];
    }
    __syncthreads();
        ICP_TICK(0);
    TileView tv{tgt, box, npad / TILE};
    const bool use_cg = a.cg_hdr != nullptr && a.cg_ok[p] != 0;
    CgView cg{};
    if (use_cg) cg = cg_view(a.cg_geom, a.cg_hdr, a.cg_arena, a.cg_ok, p);

    const double* src = a.src_f64 + (size_t)p * a.cap_s * 3;
    const double* a8 = a.align8 ? a.align8 + (size_t)p * 8 : nullptr;
    auto input_point = [&](int o, float& x, float& y, float& z) {
        double dx = src[3 * o], dy = src[3 * o + 1], dz = src[3 * o + 2];
        if (mode != 2) {                                              // initRegistrationKSS.hpp:75-109
            dx = align_coord(dx, a8[0], a8[3], a8[6]);
            dy = align_coord(dy, a8[1], a8[4], a8[6]);
            dz = align_coord(dz, a8[2], a8[5], a8[6]);
            rot_x(ang_c[0], ang_s[0], dy, dz);
            rot_y(ang_c[1], ang_s[1], dx, dz);
            rot_z(ang_c[2], ang_s[2], dx, dy);
        }
        x = (float)dx; y = (float)dy; z = (float)dz;                  // KSS_ICP.hpp:328-333
    };

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // thread 0's bookkeeping over the runs of this CTA lives in shared memory (the kernel sits at its register cap)
    __shared__ double best_fit;                       // best qualifying run so far
    __shared__ int have_best, stored, tot_iters, tot_runs;
    if (threadIdx.x == 0) { best_fit = 0.0; have_best = 0; stored = 0; tot_iters = 0; tot_runs = 0; }
  for (;;) {                                          // ---- runs of this CTA (one, unless the pair overflows the slots)
    __syncthreads();                                  // the previous run's readers of fin / d2s / the angles are done
    if (mode == 1 && threadIdx.x < 3) {
        const int h = a.minima[(size_t)p * a.hpad + hs], k = threadIdx.x;
        const int id = k == 0 ? h / (a.G * a.G) : k == 1 ? (h / a.G) % a.G : h % a.G;
        ang_c[k] = a.trig_list[2 * id]; ang_s[k] = a.trig_list[2 * id + 1];
    }
    if (threadIdx.x < 16) fin[threadIdx.x] = (threadIdx.x % 5 == 0) ? 1.0f : 0.0f;
    if (threadIdx.x == 0) { kept = 0; done = 0; prev_mse = DBL_MAX; }
    __syncthreads();
    int iters = 0;
    int converged = 0;

    for (;;) {
        // ---- (1) lazily apply the previous T_k, then correspondences for Morton-consecutive queries
        int my_kept = 0;
        auto record = [&](bool valid, int o, float x, float y, float z, unsigned long long key) {
            const float d2 = __uint_as_float((unsigned)(key >> 32));
            const unsigned orig = (unsigned)(key & 0xffffffffu);
            const bool keep = !((double)d2 > a.max_dist_sqr);        // A.3: skip iff d2 > max_dist^2
            if (valid) {
                d2s[o] = d2;
                mpos[o] = keep ? tinv[orig] : (unsigned short)0xffff;
                my_kept += keep ? 1 : 0;
                if (TRACE && a.trace_idx && iters < a.trace_cap)
                    a.trace_idx[((size_t)run * a.trace_cap + iters) * a.cap_s + o] = keep ? (int)orig : -1;
                if (TRACE && a.trace_src && iters < a.trace_cap) {
                    float* ts = a.trace_src + (((size_t)run * a.trace_cap + iters) * a.cap_s + o) * 3;
                    ts[0] = x; ts[1] = y; ts[2] = z;
                }
            }
        };
        if (use_cg) {
            // four queries per thread at once: their grid look-ups overlap (cg_query_batch)
            constexpr int U = ICP_U;
            for (int base0 = warp * 32; base0 < n_s; base0 += U * nwarps * 32) {
                float x[U], y[U], z[U]; int o[U]; bool valid[U];
                unsigned long long key[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int jpos = base0 + u * nwarps * 32 + lane;
                    valid[u] = jpos < n_s;
                    o[u] = perm[valid[u] ? jpos : n_s - 1];
                    if (iters == 0) input_point(o[u], x[u], y[u], z[u]);
                    else { xform_point(Tk, cur[o[u]], cur[n_s + o[u]], cur[2 * n_s + o[u]], x[u], y[u], z[u]); }
                }
                __syncwarp();
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (valid[u]) { cur[o[u]] = x[u]; cur[n_s + o[u]] = y[u]; cur[2 * n_s + o[u]] = z[u]; }
                cg_query_batch<U, true>(cg, tgt, n_t, x, y, z, key);
#pragma unroll
                for (int u = 0; u < U; ++u) record(valid[u], o[u], x[u], y[u], z[u], key[u]);
            }
        } else {
            for (int base = warp * 32; base < n_s; base += nwarps * 32) {
                const int jpos = base + lane;
                const bool valid = jpos < n_s;
                const int o = perm[valid ? jpos : n_s - 1];
                float x, y, z;
                if (iters == 0) input_point(o, x, y, z);
                else { xform_point(Tk, cur[o], cur[n_s + o], cur[2 * n_s + o], x, y, z); }
                __syncwarp();
                if (valid) { cur[o] = x; cur[n_s + o] = y; cur[2 * n_s + o] = z; }
                record(valid, o, x, y, z, warp_nn<true>(tv, x, y, z));
            }
        }
        my_kept = __reduce_add_sync(KSS_FULL, my_kept);
        if (lane == 0 && my_kept) atomicAdd(&kept, my_kept);
        __syncthreads();
        ICP_TICK(1);
        const int cnt = kept;
        if (cnt < 3) { converged = 0; break; }                       // min_number_correspondences_
        // ---- (2) pass A, level 1: one warp per 256-slot chunk sums kept source xyz, matched target
        //      xyz (float) and d2 (double) in the CANON256 order (lane-strided, then butterfly)
        const int nc = (n_s + 255) >> 8;
        for (int c = warp; c < nc; c += nwarps) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, t0 = 0.f, t1 = 0.f, t2 = 0.f;
            double dd = 0.0;
            const int hi = min(n_s, (c + 1) << 8);
            for (int i = (c << 8) + lane; i < hi; i += 32) {
                const unsigned m = mpos[i];
                if (m == 0xffff) continue;
                const float4 t = tgt[m];
                s0 = add_(s0, cur[i]); s1 = add_(s1, cur[n_s + i]); s2 = add_(s2, cur[2 * n_s + i]);
                t0 = add_(t0, t.x); t1 = add_(t1, t.y); t2 = add_(t2, t.z);
                dd = __dadd_rn(dd, (double)d2s[i]);
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                s0 = add_(s0, __shfl_xor_sync(KSS_FULL, s0, off)); s1 = add_(s1, __shfl_xor_sync(KSS_FULL, s1, off));
                s2 = add_(s2, __shfl_xor_sync(KSS_FULL, s2, off)); t0 = add_(t0, __shfl_xor_sync(KSS_FULL, t0, off));
                t1 = add_(t1, __shfl_xor_sync(KSS_FULL, t1, off)); t2 = add_(t2, __shfl_xor_sync(KSS_FULL, t2, off));
                dd = __dadd_rn(dd, __shfl_xor_sync(KSS_FULL, dd, off));
            }
            if (lane == 0) {
                float* pf = partF + c * 16;
                pf[0] = s0; pf[1] = s1; pf[2] = s2; pf[3] = t0; pf[4] = t1; pf[5] = t2;
                partD[c] = dd;
            }
        }
        __syncthreads();
        ICP_TICK(2);
        // level 2 (every warp redundantly, identical arithmetic): chunk results by the same rule
        auto lvl2f = [&](int q) -> float {
            if (nc == 1) return partF[q];
            float v = lane < nc ? add_(0.0f, partF[lane * 16 + q]) : 0.0f;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) v = add_(v, __shfl_xor_sync(KSS_FULL, v, off));
            return v;
        };
        const float one_over_n = div_(1.0f, (float)cnt);
        const float sm0 = mul_(lvl2f(0), one_over_n), sm1 = mul_(lvl2f(1), one_over_n), sm2 = mul_(lvl2f(2), one_over_n);
        const float dm0 = mul_(lvl2f(3), one_over_n), dm1 = mul_(lvl2f(4), one_over_n), dm2 = mul_(lvl2f(5), one_over_n);
        double dsum;
        if (nc == 1) dsum = partD[0];
        else {
            dsum = lane < nc ? __dadd_rn(0.0, partD[lane]) : 0.0;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) dsum = __dadd_rn(dsum, __shfl_xor_sync(KSS_FULL, dsum, off));
        }
        __syncthreads();                                   // partF is reused by pass B
        ICP_TICK(3);
        // ---- (3) pass B, level 1: sigma(a,b) partials = sum (d_a - dmean_a) * (s_b - smean_b)
        for (int c = warp; c < nc; c += nwarps) {
            float v[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            const int hi = min(n_s, (c + 1) << 8);
            for (int i = (c << 8) + lane; i < hi; i += 32) {
                const unsigned m = mpos[i];
                if (m == 0xffff) continue;
                const float4 t = tgt[m];
                const float sx = sub_(cur[i], sm0), sy = sub_(cur[n_s + i], sm1), sz = sub_(cur[2 * n_s + i], sm2);
                const float dx = sub_(t.x, dm0), dy = sub_(t.y, dm1), dz = sub_(t.z, dm2);
                v[0] = add_(v[0], mul_(dx, sx)); v[1] = add_(v[1], mul_(dx, sy)); v[2] = add_(v[2], mul_(dx, sz));
                v[3] = add_(v[3], mul_(dy, sx)); v[4] = add_(v[4], mul_(dy, sy)); v[5] = add_(v[5], mul_(dy, sz));
                v[6] = add_(v[6], mul_(dz, sx)); v[7] = add_(v[7], mul_(dz, sy)); v[8] = add_(v[8], mul_(dz, sz));
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
                for (int q = 0; q < 9; ++q) v[q] = add_(v[q], __shfl_xor_sync(KSS_FULL, v[q], off));
            if (lane == 0)
#pragma unroll
                for (int q = 0; q < 9; ++q) partF[c * 16 + q] = v[q];
        }
        __syncthreads();
        ICP_TICK(4);
        if (warp == 0) {
            float sg[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) sg[q] = mul_(one_over_n, lvl2f(q));
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 9; ++q) red[7 + q] = sg[q];
                redd = dsum;
            }
        }
        __syncwarp();
        ICP_TICK(11);
        // ---- (4) one thread: SVD/Kabsch, accumulate, convergence
        if (threadIdx.x == 0) {
            float sigma[9], smean[3] = {sm0, sm1, sm2}, dmean[3] = {dm0, dm1, dm2}, T[16];
#pragma unroll
            for (int i = 0; i < 9; ++i) sigma[i] = red[7 + i];
            umeyama_finish(sigma, smean, dmean, T);
            ICP_TICK(8);
            float F[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) F[i] = fin[i];
            mat4_mul(T, F, F);
            ICP_TICK(9);
#pragma unroll
            for (int i = 0; i < 16; ++i) { Tk[i] = T[i]; fin[i] = F[i]; }
            const double mse = __ddiv_rn(redd, (double)cnt);
            if (TRACE && a.trace_T && iters < a.trace_cap)
                for (int i = 0; i < 16; ++i) a.trace_T[((size_t)run * a.trace_cap + iters) * 16 + i] = T[i];
            if (TRACE && a.trace_mse && iters < a.trace_cap) a.trace_mse[(size_t)run * a.trace_cap + iters] = mse;
            const int it = iters + 1;
            int dn = 0;
            if (it >= a.max_iter) dn = 1;                                            // A.6 (1)
            else {
                const double cos_angle = 0.5 * (double)sub_(add_(add_(T[0], T[5]), T[10]), 1.0f);
                const double tr2 = (double)add_(add_(mul_(T[3], T[3]), mul_(T[7], T[7])), mul_(T[11], T[11]));
                if (cos_angle >= a.rot_thr && tr2 <= a.trans_thr) dn = 1;           // A.6 (2)
                else if (fabs(__dsub_rn(mse, prev_mse)) < a.mse_abs) dn = 1;        // A.6 (3) absolute
                else if (__ddiv_rn(fabs(__dsub_rn(mse, prev_mse)), prev_mse) < a.mse_rel) dn = 1;  // relative
                else prev_mse = mse;
            }
            ICP_TICK(10);
            done = dn; kept = 0;
        }
        __syncthreads();
        ICP_TICK(5);
        ++iters;
        if (done) { converged = 1; break; }
    }

    // ---- getFitnessScore: final * ORIGINAL input (one rounding), NN, mean of d2 in double (A.7)
    __syncthreads();
        ICP_TICK(6);
    for (int base = warp * 32; base < n_s; base += nwarps * 32) {
        const int jpos = base + lane;
        const bool valid = jpos < n_s;
        const int o = perm[valid ? jpos : n_s - 1];
        float x, y, z, fx, fy, fz;
        input_point(o, x, y, z);
        xform_point(fin, x, y, z, fx, fy, fz);
        const unsigned long long key = use_cg ? cg_query<false>(cg, tgt, n_t, fx, fy, fz) : warp_nn<false>(tv, fx, fy, fz);
        if (valid) d2s[o] = __uint_as_float((unsigned)(key >> 32));
    }
    __syncthreads();
        ICP_TICK(7);
    if (warp == 0) {
        double s = canon_sum_warp_f64(n_s, [&](int i, double& v) { v = (double)d2s[i]; return true; });
        int take = 0;
        if (lane == 0) {
            const double fit = n_s > 0 ? __ddiv_rn(s, (double)n_s) : DBL_MAX;
            const int qual = (fit >= 0.0 && fit < 9999.0) ? 1 : 0;
            tot_iters += iters; ++tot_runs;
            // the first run is always stored (it is what the reference uses when nothing qualifies: angleIndex = 0);
            // later runs (increasing l) replace it only by a strictly smaller qualifying fitness
            if (a.hyp_fit) {                          // the full fitness / iteration vectors of a hypothesis-sharded pair
                a.hyp_fit[(size_t)p * 2 * a.hpad + hs] = fit;
                a.hyp_fit[(size_t)p * 2 * a.hpad + a.hpad + hs] = (double)iters;
            }
            take = !stored || (qual && (!have_best || fit < best_fit));
            if (take) {
                a.run_fit[run] = fit;
                a.run_iters[run] = iters;
                a.run_conv[run] = converged;
                if (a.run_hyp) a.run_hyp[run] = hs * 2 + qual;
                stored = 1;
                if (qual) { have_best = 1; best_fit = fit; }
            }
        }
        take = __shfl_sync(KSS_FULL, take, 0);
        if (take && lane < 16) a.run_T[(size_t)run * 16 + lane] = fin[lane];
    }
    if (mode != 1) break;
    hs += (a.runs_per_pair - 1) * max(a.hyp_world, 1);
    if (hs >= a.n_minima[p]) break;
  }
    if (threadIdx.x == 0 && a.run_tot) { a.run_tot[2 * run] = tot_iters; a.run_tot[2 * run + 1] = tot_runs; }
}

// =============================================================== select_kernel
// KSS_ICP.hpp:99-125: multi-hypothesis branch iff judge fitness > 0.0005; winner = first
// strict minimum with ri >= 0 starting from Q = 9999; else the sweep winner.
__global__ void select_kernel(int P, int runs_per_pair, int hpad, int G, double judge_thr,
                              const double* __restrict__ align8,
                              const double* __restrict__ run_fit, const int* __restrict__ run_iters,
                              const float* __restrict__ run_T, const int* __restrict__ run_conv,
                              const int* __restrict__ run_hyp, const int* __restrict__ run_tot,
                              const int* __restrict__ best_h, const int* __restrict__ minima,
                              const int* __restrict__ n_minima,
                              PairOut* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    PairOut& o = out[p];
    for (int i = 0; i < 8; ++i) o.align[i] = align8[(size_t)p * 8 + i];
    const double E = run_fit[p * runs_per_pair];
    o.judge_fitness = E;
    o.judge_iters = run_iters[p * runs_per_pair];
    o.n_minima = n_minima[p];
    o.best_h = best_h[p];
    o.G = G;
    int total = o.judge_iters, nruns = 1;
    int use_run = 0, used_h = best_h[p], winner = -1, multi = 0, overflow = 0;
    if (E > judge_thr) {
        multi = 1;
        // slot s holds the lexicographic minimum (fitness, l) of the hypotheses l = s, s + slots, ... that pass
        // `ri >= 0 && ri < 9999` (or, when none does, its first run); the minimum over the slots is the l the serial
        // scan `if (ri < Q && ri >= 0) { Q = ri; angleIndex = l; }` ends with (first strict minimum in l order)
        double Q = 9999.0;
        int angleIndex = 0, slot_of = 0, have = 0;
        int L = n_minima[p];
        int S = L;
        if (L > runs_per_pair - 1) { overflow = 1; S = runs_per_pair - 1; }
        for (int sidx = 0; sidx < S; ++sidx) {
            const int rr = p * runs_per_pair + 1 + sidx;
            const double ri = run_fit[rr];
            const int hv = run_hyp[rr], l = hv >> 1;
            total += run_tot[2 * rr]; nruns += run_tot[2 * rr + 1];
            if ((hv & 1) && (!have || ri < Q || (ri == Q && l < angleIndex))) { Q = ri; angleIndex = l; slot_of = sidx; have = 1; }
        }
        winner = angleIndex;
        use_run = 1 + slot_of;
        used_h = minima[(size_t)p * hpad + angleIndex];
    }
    // the final ICP (KSS_ICP.hpp:130) repeats the winner's run on identical input: reuse it
    o.branch_multi = multi; o.winner = winner; o.used_h = used_h; o.use_list = multi;
    o.final_fitness = run_fit[p * runs_per_pair + use_run];
    o.final_iters = run_iters[p * runs_per_pair + use_run];
    o.final_converged = run_conv[p * runs_per_pair + use_run];
    o.reserved_[0] = 0; o.reserved_[1] = 0;
    o.total_icp_iters = total + o.final_iters;
    o.n_icp_runs = nruns + 1;
    o.overflow = overflow;
    for (int i = 0; i < 16; ++i) o.T[i] = run_T[((size_t)p * runs_per_pair + use_run) * 16 + i];
}

// =============================================================== hypothesis-sharded selection
// The fitness vector is complete on every rank after ONE all-reduce(MIN) (entries a rank did not run were +inf):
// every rank applies KSS_ICP.hpp:100-116 to it -- Q = 9999, first l with ri < Q && ri >= 0 -- and queues the winner's
// start for the final ICP (KSS_ICP.hpp:130), which every rank repeats (deterministic, so identical).
// hyp_fit [P][2][hpad]: fitness | iteration count (as double) per hypothesis l.
__global__ void select_sharded_kernel(int P, int runs_per_pair, int hpad, int G, double judge_thr,
                                      const double* __restrict__ align8, const double* __restrict__ run_fit,
                                      const int* __restrict__ run_iters, const double* __restrict__ hyp_fit,
                                      const int* __restrict__ best_h, const int* __restrict__ minima,
                                      const int* __restrict__ n_minima, PairOut* __restrict__ out,
                                      int* __restrict__ win_minima /* [P][hpad], entry 0 */, int* __restrict__ n_win /* [P] */) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    PairOut& o = out[p];
    for (int i = 0; i < 8; ++i) o.align[i] = align8[(size_t)p * 8 + i];
    const double E = run_fit[p * runs_per_pair];
    o.judge_fitness = E;
    o.judge_iters = run_iters[p * runs_per_pair];
    o.n_minima = n_minima[p];
    o.best_h = best_h[p];
    o.G = G;
    int total = o.judge_iters, nruns = 1, used_h = best_h[p], winner = -1, multi = 0;
    if (E > judge_thr) {
        multi = 1;
        double Q = 9999.0;
        int angleIndex = 0;
        const int L = n_minima[p];
        const double* fit = hyp_fit + (size_t)p * 2 * hpad;
        for (int l = 0; l < L; ++l) {
            const double ri = fit[l];
            total += (int)fit[hpad + l]; ++nruns;
            if (ri < Q && ri >= 0.0) { Q = ri; angleIndex = l; }
        }
        winner = angleIndex;
        used_h = minima[(size_t)p * hpad + angleIndex];
    }
    o.branch_multi = multi; o.winner = winner; o.used_h = used_h; o.use_list = multi;
    o.total_icp_iters = total; o.n_icp_runs = nruns + 1;          // (+ the final run, whose iterations are added below)
    o.overflow = 0; o.reserved_[0] = 0; o.reserved_[1] = 0;
    win_minima[(size_t)p * hpad] = used_h;
    n_win[p] = multi;
}
// after the final run: run2_* has two slots per pair, slot 1 = the winner's repeated run (multi pairs only)
__global__ void finish_sharded_kernel(int P, int runs_per_pair, const double* __restrict__ run_fit, const int* __restrict__ run_iters,
                                      const int* __restrict__ run_conv, const float* __restrict__ run_T,
                                      const double* __restrict__ run2_fit, const int* __restrict__ run2_iters,
                                      const int* __restrict__ run2_conv, const float* __restrict__ run2_T, PairOut* __restrict__ out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    PairOut& o = out[p];
    const bool m = o.branch_multi != 0;
    const size_t r1 = (size_t)p * runs_per_pair, r2 = (size_t)p * 2 + 1;
    o.final_fitness = m ? run2_fit[r2] : run_fit[r1];
    o.final_iters = m ? run2_iters[r2] : run_iters[r1];
    o.final_converged = m ? run2_conv[r2] : run_conv[r1];
    o.total_icp_iters += o.final_iters;
    for (int i = 0; i < 16; ++i) o.T[i] = m ? run2_T[r2 * 16 + i] : run_T[r1 * 16 + i];
}
cudaError_t launch_select_sharded(cudaStream_t st, int P, int runs_per_pair, int hpad, int G, double judge_thr, const double* align8,
                                  const double* run_fit, const int* run_iters, const double* hyp_fit, const int* best_h,
                                  const int* minima, const int* n_minima, PairOut* out, int* win_minima, int* n_win) {
    select_sharded_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, runs_per_pair, hpad, G, judge_thr, align8, run_fit, run_iters, hyp_fit,
                                                          best_h, minima, n_minima, out, win_minima, n_win);
    return cudaGetLastError();
}
cudaError_t launch_finish_sharded(cudaStream_t st, int P, int runs_per_pair, const double* run_fit, const int* run_iters,
                                  const int* run_conv, const float* run_T, const double* run2_fit, const int* run2_iters,
                                  const int* run2_conv, const float* run2_T, PairOut* out) {
    finish_sharded_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, runs_per_pair, run_fit, run_iters, run_conv, run_T, run2_fit, run2_iters,
                                                          run2_conv, run2_T, out);
    return cudaGetLastError();
}

// =============================================================== final_apply_kernel
// pointAlign = rt * Rotation_Angle(pointSource) : similarity in double (initRegistrationKSS.hpp:75-109),
// then float 3x4 coefficients promoted to double, left-to-right sums (KSS_ICP.hpp:224-230).
__global__ void __launch_bounds__(256)
final_apply_kernel(const double* __restrict__ full_s, const int* __restrict__ cnt_S, int cap_S,
                   const double* __restrict__ align8, const PairOut* __restrict__ out,
                   const double* __restrict__ trig_accum, const double* __restrict__ trig_list, int G,
                   double* __restrict__ point_align) {
    const int p = blockIdx.y;
    const int n = cnt_S ? cnt_S[p] : cap_S;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const PairOut& o = out[p];
    const double* trig = o.use_list ? trig_list : trig_accum;
    const int h = o.used_h;
    const int ix[3] = {h / (G * G), (h / G) % G, h % G};
    const double* a8 = align8 + (size_t)p * 8;
    const double* src = full_s + ((size_t)p * cap_S + i) * 3;
    double x = align_coord(src[0], a8[0], a8[3], a8[6]);
    double y = align_coord(src[1], a8[1], a8[4], a8[6]);
    double z = align_coord(src[2], a8[2], a8[5], a8[6]);
    rot_x(trig[2 * ix[0]], trig[2 * ix[0] + 1], y, z);
    rot_y(trig[2 * ix[1]], trig[2 * ix[1] + 1], x, z);
    rot_z(trig[2 * ix[2]], trig[2 * ix[2] + 1], x, y);
    double* dst = point_align + ((size_t)p * cap_S + i) * 3;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double v = __dmul_rn((double)o.T[4 * r + 0], x);
        v = __dadd_rn(v, __dmul_rn((double)o.T[4 * r + 1], y));
        v = __dadd_rn(v, __dmul_rn((double)o.T[4 * r + 2], z));
        v = __dadd_rn(v, (double)o.T[4 * r + 3]);
        dst[r] = v;
    }
}

// =============================================================== generic point kernels (any n)
__global__ void __launch_bounds__(256)
apply_similarity_kernel(const double* __restrict__ pts, int n, const double* __restrict__ a7,
                        const double* __restrict__ cs /* c0,s0,c1,s1,c2,s2 */, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x = align_coord(pts[3 * i], a7[0], a7[3], a7[6]);
    double y = align_coord(pts[3 * i + 1], a7[1], a7[4], a7[6]);
    double z = align_coord(pts[3 * i + 2], a7[2], a7[5], a7[6]);
    rot_x(cs[0], cs[1], y, z);
    rot_y(cs[2], cs[3], x, z);
    rot_z(cs[4], cs[5], x, y);
    out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = z;
}

__global__ void __launch_bounds__(256)
apply_transform_kernel(const double* __restrict__ pts, int n, const float* __restrict__ T, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double v = __dmul_rn((double)T[4 * r + 0], x);
        v = __dadd_rn(v, __dmul_rn((double)T[4 * r + 1], y));
        v = __dadd_rn(v, __dmul_rn((double)T[4 * r + 2], z));
        v = __dadd_rn(v, (double)T[4 * r + 3]);
        out[3 * i + r] = v;
    }
}

// =============================================================== nn / metrics on the small path
// One CTA per pair: exact NN of every query (double, narrowed with RN) against the
// Morton-tiled target.  MODE 0: write idx/d2.  MODE 1: PCR_QM serial double sums
// (registrationMeasure.hpp:66-88) -> out3 = {MSE, RMSE, MAE}.
template <int MODE>
__global__ void __launch_bounds__(256)
nn_small_kernel(const double* __restrict__ q, const int* __restrict__ cnt_q, int cap_q,
                const unsigned short* __restrict__ q_perm,
                const float4* __restrict__ t_sorted, const float* __restrict__ t_box,
                const int* __restrict__ cnt_t, int cap_tpad,
                int* __restrict__ idx, float* __restrict__ d2out, double* __restrict__ out3, int out3_stride) {
    extern __shared__ unsigned char smem_raw[];
    const int p = blockIdx.x;
    const int n_q = cnt_q ? cnt_q[p] : cap_q;
    const int n_t = cnt_t[p];
    const int npad = (n_t + TILE - 1) / TILE * TILE;
    float4* tgt = reinterpret_cast<float4*>(smem_raw);
    float* box = reinterpret_cast<float*>(tgt + npad);
    float* d2s = box + 6 * MAX_TILES;                  // [n_q]      (MODE 1)
    const float4* gts = t_sorted + (size_t)p * cap_tpad;
    for (int j = threadIdx.x; j < npad; j += blockDim.x) tgt[j] = gts[j];
    const float* gtb = t_box + (size_t)p * 6 * MAX_TILES;
    for (int j = threadIdx.x; j < 6 * MAX_TILES; j += blockDim.x) box[j] = gtb[j];
    __syncthreads();
    TileView tv{tgt, box, npad / TILE};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const double* qs = q + (size_t)p * cap_q * 3;
    const unsigned short* perm = q_perm + (size_t)p * cap_q;
    for (int base = warp * 32; base < n_q; base += nwarps * 32) {
        const int jpos = base + lane;
        const bool valid = jpos < n_q;
        const int o = perm[valid ? jpos : n_q - 1];
        const float x = (float)qs[3 * o], y = (float)qs[3 * o + 1], z = (float)qs[3 * o + 2];
        const unsigned long long key = warp_nn<MODE == 0>(tv, x, y, z);
        const float d2 = __uint_as_float((unsigned)(key >> 32));
        if (valid) {
            if (MODE == 0) { idx[(size_t)p * cap_q + o] = (int)(key & 0xffffffffu); d2out[(size_t)p * cap_q + o] = d2; }
            else d2s[o] = d2;
        }
    }
    if (MODE == 1) {
        __syncthreads();
        __shared__ double sums[2];
        // the square roots by everybody (into the target tile, which is no longer needed), so that the two serial chains
        // of the reference's loop (registrationMeasure.hpp:83-88) are additions only
        double* sq = reinterpret_cast<double*>(tgt);
        const bool par = (size_t)n_q * sizeof(double) <= (size_t)npad * sizeof(float4);
        if (par) for (int i = threadIdx.x; i < n_q; i += blockDim.x) sq[i] = __dsqrt_rn((double)d2s[i]);
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
#pragma unroll 8
            for (int i = 0; i < n_q; ++i) s = __dadd_rn(s, (double)d2s[i]);
            sums[0] = s;
        }
        if (threadIdx.x == 32) {
            double s = 0.0;
            if (par) {
#pragma unroll 8
                for (int i = 0; i < n_q; ++i) s = __dadd_rn(s, sq[i]);
            } else {
#pragma unroll 4
                for (int i = 0; i < n_q; ++i) s = __dadd_rn(s, __dsqrt_rn((double)d2s[i]));
            }
            sums[1] = s;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const double mse = __ddiv_rn(sums[0], (double)n_q);
            out3[(size_t)p * out3_stride + 0] = mse;
            out3[(size_t)p * out3_stride + 1] = __dsqrt_rn(mse);
            out3[(size_t)p * out3_stride + 2] = __ddiv_rn(sums[1], (double)n_q);
        }
    }
}

// =============================================================== host-side launchers
static inline size_t sort_smem() { return SMALL_MAX * sizeof(unsigned) + 3 * SMALL_MAX * sizeof(float); }

cudaError_t launch_sort_target(cudaStream_t st, int P, const double* pts, const int* cnt, int cap,
                               float4* t_sorted, float* t_box, unsigned short* t_inv, int cap_pad) {
    sort_cloud_kernel<<<P, 256, sort_smem(), st>>>(pts, cnt, cap, 0, t_sorted, t_box, t_inv, cap_pad, nullptr);
    return cudaGetLastError();
}
cudaError_t launch_sort_source(cudaStream_t st, int P, const double* pts, const int* cnt, int cap,
                               unsigned short* perm) {
    sort_cloud_kernel<<<P, 256, sort_smem(), st>>>(pts, cnt, cap, 1, nullptr, nullptr, nullptr, 0, perm);
    return cudaGetLastError();
}
cudaError_t launch_middle_align(cudaStream_t st, int P, const double* sim_s, const int* cnt_s, int cap_s,
                                const double* sim_t, const int* cnt_t, int cap_t, double* align8, double* s_al) {
    const size_t smem = 4 * SMALL_MAX * sizeof(double);
    // the opt-in is per device: set on every launch (one ctx per GPU, several ctxs per process)
    cudaFuncSetAttribute(middle_align_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    middle_align_kernel<<<P, 256, smem, st>>>(sim_s, cnt_s, cap_s, sim_t, cnt_t, cap_t, align8, s_al);
    return cudaGetLastError();
}
cudaError_t launch_sweep(cudaStream_t st, int P, const double* s_al, const int* cnt_s, int cap_s,
                         const unsigned short* s_perm, const float4* t_sorted, const float* t_box,
                         const int* cnt_t, int cap_tpad, const double* trig_accum, int G, int score_mode,
                         float* rbuf, int hpad, const CgBuffers* cg, int ij_lo, int ij_hi) {
    const size_t smem = (size_t)cap_tpad * sizeof(float4) + 6 * MAX_TILES * sizeof(float);
    cudaFuncSetAttribute(sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (ij_hi <= ij_lo) return cudaSuccess;
    sweep_kernel<<<dim3(ij_hi - ij_lo, P), 256, smem, st>>>(s_al, cnt_s, cap_s, s_perm, t_sorted, t_box, cnt_t, cap_tpad,
                                                   trig_accum, G, score_mode, rbuf, hpad, ij_lo,
                                                   cg ? cg->geom : nullptr, cg ? cg->hdr : nullptr,
                                                   cg ? cg->arena : nullptr, cg ? cg->ok : nullptr);
    return cudaGetLastError();
}
// candidate grid: geometry, then one launch per level (a level reads the previous one)
cudaError_t launch_cg_build(cudaStream_t st, int P, int geom_mode, const double* a, const int* cnt_a, int cap_a,
                            const double* b, const int* cnt_b, int cap_b, const float4* t_sorted,
                            const unsigned short* t_inv, int cap_t, const int* cnt_t,
                            int cap_tpad, const CgBuffers& cg, int* launches) {
    cg_geom_kernel<<<P, 256, 0, st>>>(geom_mode, a, cnt_a, cap_a, b, cnt_b, cap_b, cg.geom, cg.cursor, cg.ok, cg.wl_cnt, cg.wl2_cnt);
    int n = 1;
    const char* e = getenv("KSS_CG_NO_REFINE");
    const bool refine = !(e && e[0] == '1');
    cg_level0_kernel<<<dim3((CG_NG0 * CG_NG0 * CG_NG0 + 7) / 8, P), 256, 0, st>>>(t_sorted, t_inv, cap_t, cnt_t, cap_tpad, cg.geom, cg.hdr,
                                                                             cg.arena, cg.cursor, cg.ok);
    ++n;
    for (int l = 1; l < CG_LEVELS; ++l) {
        const int ncells = cg_ng(l) * cg_ng(l) * cg_ng(l);
        int coop_levels = 0x6;                                      // levels 1 and 2: long parent lists
        const char* ec = getenv("KSS_CG_COOP_LEVELS");
        if (ec) coop_levels = atoi(ec);
        if (!((coop_levels >> l) & 1))
            cg_level_kernel<0><<<dim3((ncells + 255) / 256, P), 256, 0, st>>>(l, t_sorted, t_inv, cap_t, cnt_t, cap_tpad, cg.geom, cg.hdr,
                                                                         cg.arena, cg.cursor, cg.ok, refine ? cg.wl : nullptr, cg.wl_cnt,
                                                                         nullptr, cg.wl2_cnt);
        else
            cg_level_coop_kernel<<<dim3((ncells * 8 + 255) / 256, P), 256, 0, st>>>(l, t_sorted, t_inv, cap_t, cnt_t, cap_tpad, cg.geom,
                                                                               cg.hdr, cg.arena, cg.cursor, cg.ok,
                                                                               refine ? cg.wl : nullptr, cg.wl_cnt);
        ++n;
    }
    if (refine) {
        cg_level_kernel<1><<<dim3(48, P), 256, 0, st>>>(CG_LEVELS, t_sorted, t_inv, cap_t, cnt_t, cap_tpad, cg.geom,
                                                                            cg.hdr, cg.arena, cg.cursor, cg.ok, cg.wl, cg.wl_cnt,
                                                                            cg.wl2, cg.wl2_cnt);
        cg_level_kernel<2><<<dim3(8, P), 256, 0, st>>>(CG_LEVELS + 1, t_sorted, t_inv, cap_t, cnt_t, cap_tpad,
                                                                             cg.geom, cg.hdr, cg.arena, cg.cursor, cg.ok, cg.wl,
                                                                             cg.wl_cnt, cg.wl2, cg.wl2_cnt);
        n += 2;
    }
    if (launches) *launches = n;
    return cudaGetLastError();
}
size_t cg_hdr_words_per_pair() { return CG_HDR_TOTAL; }
size_t cg_arena_entries_per_pair() { return CG_ARENA; }
size_t cg_worklist_entries_per_pair() { return CG_WL_CAP; }
size_t cg_worklist2_entries_per_pair() { return CG_WL2_CAP; }

cudaError_t launch_sweep_finalize(cudaStream_t st, int P, const float* rbuf, const int* cnt_s, int cap_s, int hpad,
                                  int G, int score_mode, double* value, int* best_h, int* minima, int* n_minima,
                                  int h_lo, int h_hi, int phases) {
    const int H = G * G * G;
    const size_t smem = (size_t)H * sizeof(double) + H;
    int threads = H < 1024 ? (H + 31) / 32 * 32 : 1024;
    sweep_finalize_kernel<<<P, threads, smem, st>>>(rbuf, cnt_s, cap_s, hpad, G, score_mode, h_lo, h_hi, phases, value, best_h,
                                                    minima, n_minima);
    return cudaGetLastError();
}
size_t icp_smem_bytes(int cap_s, int cap_t, int cap_tpad) {
    return (size_t)cap_tpad * sizeof(float4) + 6 * MAX_TILES * sizeof(float) + (size_t)cap_s * 4 * sizeof(float) +
           (size_t)cap_s * 2 * sizeof(unsigned short) + (size_t)cap_t * sizeof(unsigned short) + 16;
}
cudaError_t launch_icp(cudaStream_t st, int P, int slots, const IcpArgs& a) {
    const size_t smem = icp_smem_bytes(a.cap_s, a.cap_t, a.cap_tpad);
    const bool trace = a.trace_cap > 0 && (a.trace_idx || a.trace_T || a.trace_mse || a.trace_src);
    if (trace) {
        cudaFuncSetAttribute(icp_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        icp_small_kernel<true><<<dim3(slots, P), 256, smem, st>>>(a);
    } else {
        cudaFuncSetAttribute(icp_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        icp_small_kernel<false><<<dim3(slots, P), 256, smem, st>>>(a);
    }
    return cudaGetLastError();
}
cudaError_t launch_select(cudaStream_t st, int P, int runs_per_pair, int hpad, int G, double judge_thr,
                          const double* align8, const double* run_fit, const int* run_iters, const float* run_T,
                          const int* run_conv, const int* run_hyp, const int* run_tot, const int* best_h,
                          const int* minima, const int* n_minima, PairOut* out) {
    select_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, runs_per_pair, hpad, G, judge_thr, align8, run_fit, run_iters, run_T,
                                                  run_conv, run_hyp, run_tot, best_h, minima, n_minima, out);
    return cudaGetLastError();
}
cudaError_t launch_final_apply(cudaStream_t st, int P, const double* full_s, const int* cnt_S, int cap_S,
                               const double* align8, const PairOut* out, const double* trig_accum,
                               const double* trig_list, int G, double* point_align) {
    final_apply_kernel<<<dim3((cap_S + 255) / 256, P), 256, 0, st>>>(full_s, cnt_S, cap_S, align8, out, trig_accum,
                                                                    trig_list, G, point_align);
    return cudaGetLastError();
}
cudaError_t launch_apply_similarity(cudaStream_t st, const double* pts, int n, const double* a7, const double* cs, double* out) {
    apply_similarity_kernel<<<(n + 255) / 256, 256, 0, st>>>(pts, n, a7, cs, out);
    return cudaGetLastError();
}
cudaError_t launch_apply_transform(cudaStream_t st, const double* pts, int n, const float* T, double* out) {
    apply_transform_kernel<<<(n + 255) / 256, 256, 0, st>>>(pts, n, T, out);
    return cudaGetLastError();
}
cudaError_t launch_nn_small(cudaStream_t st, int P, int mode, const double* q, const int* cnt_q, int cap_q,
                            const unsigned short* q_perm, const float4* t_sorted, const float* t_box,
                            const int* cnt_t, int cap_tpad, int* idx, float* d2, double* out3, int out3_stride) {
    const size_t smem = (size_t)cap_tpad * sizeof(float4) + 6 * MAX_TILES * sizeof(float) + (size_t)cap_q * sizeof(float);
    if (mode == 0) {
        cudaFuncSetAttribute(nn_small_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        nn_small_kernel<0><<<P, 256, smem, st>>>(q, cnt_q, cap_q, q_perm, t_sorted, t_box, cnt_t, cap_tpad, idx, d2, out3, out3_stride);
    } else {
        cudaFuncSetAttribute(nn_small_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        nn_small_kernel<1><<<P, 256, smem, st>>>(q, cnt_q, cap_q, q_perm, t_sorted, t_box, cnt_t, cap_tpad, idx, d2, out3, out3_stride);
    }
    return cudaGetLastError();
}

}  // namespace kss
