// kss_large.cu -- clouds beyond one CTA's shared memory (> 2048 points).
//
// Replaces pcl::KdTreeFLANN + the PCL ICP loop at full resolution (shapeRegistration_ICP(int),
// KSS_ICP.hpp:133-183; PCR_QM, registrationMeasure.hpp:47-98) with
//   build : Morton-cell bucket sort of the target (histogram + scan + atomic scatter) and a 32-ary
//           box pyramid over 32-point tiles (order never affects results: the NN is exact and ties
//           go to the lowest ORIGINAL index)
//   nn    : warp-cooperative best-first descent of the pyramid, one query per lane,
//           32 Morton-consecutive queries per warp; tiles are staged through shared memory and
//           scanned with broadcast LDS.128
//   reduce: CANON256 sums by ORIGINAL source index (one warp per 256-element chunk, last CTA
//           finishes the upper levels), umeyama/SVD on one thread, device-side convergence
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <vector>

#include "kss_device.cuh"
#include "kss_large.h"

namespace kss {

constexpr int LG_MAX_LEVELS = 4;     // 32^4 tiles * 32 points = 33.5 M points
constexpr int LG_WARPS = 8;          // warps per CTA in the NN kernel

struct Pyramid {
    const float4* tp;                // Morton-cell ordered targets {x,y,z,bits(orig)}, padded to x32
    int nlev;                        // number of box levels (level 0 = tiles)
    int cnt[LG_MAX_LEVELS];          // boxes per level
    int pad[LG_MAX_LEVELS];          // allocated boxes per level (multiple of 32)
    const float* box[LG_MAX_LEVELS]; // SoA [6][pad[l]]
};

// ------------------------------------------------------------------ build kernels
__global__ void lg_init_bbox_kernel(unsigned* bb) {
    if (threadIdx.x < 3) { bb[threadIdx.x] = 0xffffffffu; bb[3 + threadIdx.x] = 0u; }
}

// double[n][3] -> float4 {x,y,z,bits(i)} with RN narrowing (KSS_ICP.hpp:137-152), plus bounding box
__global__ void __launch_bounds__(256)
lg_convert_bbox_kernel(const double* __restrict__ pts, int n, float4* __restrict__ out, unsigned* __restrict__ bb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    if (i < n) {
        const float x = (float)pts[3 * (size_t)i], y = (float)pts[3 * (size_t)i + 1], z = (float)pts[3 * (size_t)i + 2];
        out[i] = make_float4(x, y, z, __int_as_float(i));
        mn[0] = mx[0] = f2ord(x); mn[1] = mx[1] = f2ord(y); mn[2] = mx[2] = f2ord(z);
    }
    for (int a = 0; a < 3; ++a) {
        const unsigned m0 = __reduce_min_sync(KSS_FULL, mn[a]);
        const unsigned m1 = __reduce_max_sync(KSS_FULL, mx[a]);
        if ((threadIdx.x & 31) == 0 && m0 <= m1) { atomicMin(&bb[a], m0); atomicMax(&bb[3 + a], m1); }
    }
}

__device__ __forceinline__ unsigned lg_cell(float4 p, const unsigned* bb, int bits) {
    const float cells = (float)(1 << bits) - 0.001f;
    unsigned code = 0;
    const float v[3] = {p.x, p.y, p.z};
    unsigned q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float lo = ord2f(bb[a]), hi = ord2f(bb[3 + a]);
        const float e = hi - lo;
        const float s = e > 0.0f ? cells / e : 0.0f;
        q[a] = (unsigned)fminf(fmaxf((v[a] - lo) * s, 0.0f), cells);
    }
    for (int b = 0; b < bits; ++b)
        code |= (((q[0] >> b) & 1u) << (3 * b)) | (((q[1] >> b) & 1u) << (3 * b + 1)) | (((q[2] >> b) & 1u) << (3 * b + 2));
    return code;
}

__global__ void __launch_bounds__(256)
lg_hist_kernel(const float4* __restrict__ p4, int n, const unsigned* __restrict__ bb, int bits,
               unsigned* __restrict__ cellid, unsigned* __restrict__ hist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned c = lg_cell(p4[i], bb, bits);
    cellid[i] = c;
    atomicAdd(&hist[c], 1u);
}

// exclusive scan of m counters: (1) per-1024 block scan + block totals, (2) scan of totals, (3) add
__global__ void __launch_bounds__(1024)
lg_scan1_kernel(unsigned* __restrict__ data, int m, unsigned* __restrict__ totals) {
    __shared__ unsigned ws[32];
    const int i = blockIdx.x * 1024 + threadIdx.x;
    const unsigned v = i < m ? data[i] : 0u;
    unsigned x = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(KSS_FULL, x, o); if (lane >= o) x += y; }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned w = ws[lane];
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(KSS_FULL, w, o); if (lane >= o) w += y; }
        ws[lane] = w;
    }
    __syncthreads();
    const unsigned incl = x + (warp ? ws[warp - 1] : 0u);
    if (i < m) data[i] = incl - v;
    if (threadIdx.x == 1023) totals[blockIdx.x] = incl;
}
__global__ void __launch_bounds__(1024)
lg_scan2_kernel(unsigned* __restrict__ totals, int nb) {          // nb <= 4096 block totals, one CTA
    __shared__ unsigned ws[32];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < nb ? totals[i] : 0u;
        unsigned x = v;
        for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(KSS_FULL, x, o); if (lane >= o) x += y; }
        if (lane == 31) ws[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned w = ws[lane];
            for (int o = 1; o < 32; o <<= 1) { unsigned y = __shfl_up_sync(KSS_FULL, w, o); if (lane >= o) w += y; }
            ws[lane] = w;
        }
        __syncthreads();
        const unsigned incl = x + (warp ? ws[warp - 1] : 0u) + carry;
        if (i < nb) totals[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(1024)
lg_scan3_kernel(unsigned* __restrict__ data, int m, const unsigned* __restrict__ totals) {
    const int i = blockIdx.x * 1024 + threadIdx.x;
    if (i < m) data[i] += totals[blockIdx.x];
}

// mode 0: scatter target points ; mode 1: scatter original indices (query order)
__global__ void __launch_bounds__(256)
lg_scatter_kernel(const float4* __restrict__ p4, int n, const unsigned* __restrict__ cellid,
                  unsigned* __restrict__ offs, float4* __restrict__ tp, int* __restrict__ perm) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned pos = atomicAdd(&offs[cellid[i]], 1u);
    if (tp) tp[pos] = p4[i];
    if (perm) perm[pos] = i;
}

__global__ void lg_pad_kernel(float4* tp, int n, int npad) {
    const int i = n + blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npad) tp[i] = make_float4(PAD_COORD, PAD_COORD, PAD_COORD, __int_as_float(0x7fffffff));
}

// level 0: one warp per tile, pads replicate the tile's first point (every tile has >= 1 real point)
__global__ void __launch_bounds__(256)
lg_tile_box_kernel(const float4* __restrict__ tp, int n, int ntiles, int pad, float* __restrict__ box) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= ntiles) return;
    const int j = t * TILE + lane;
    const float4 p = tp[j < n ? j : t * TILE];
    const float v0 = warp_min_f(p.x), v1 = warp_min_f(p.y), v2 = warp_min_f(p.z);
    const float v3 = warp_max_f(p.x), v4 = warp_max_f(p.y), v5 = warp_max_f(p.z);
    if (lane == 0) {
        box[0 * pad + t] = v0; box[1 * pad + t] = v1; box[2 * pad + t] = v2;
        box[3 * pad + t] = v3; box[4 * pad + t] = v4; box[5 * pad + t] = v5;
    }
}
// level l from level l-1: one warp per parent, lane = child
__global__ void __launch_bounds__(256)
lg_up_box_kernel(const float* __restrict__ cbox, int ccnt, int cpad, int pcnt, int ppad, float* __restrict__ pbox) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (t >= pcnt) return;
    const int c = min(t * 32 + lane, ccnt - 1);
    const float v0 = warp_min_f(cbox[0 * cpad + c]), v1 = warp_min_f(cbox[1 * cpad + c]), v2 = warp_min_f(cbox[2 * cpad + c]);
    const float v3 = warp_max_f(cbox[3 * cpad + c]), v4 = warp_max_f(cbox[4 * cpad + c]), v5 = warp_max_f(cbox[5 * cpad + c]);
    if (lane == 0) {
        pbox[0 * ppad + t] = v0; pbox[1 * ppad + t] = v1; pbox[2 * ppad + t] = v2;
        pbox[3 * ppad + t] = v3; pbox[4 * ppad + t] = v4; pbox[5 * ppad + t] = v5;
    }
}

// ------------------------------------------------------------------ warp-cooperative pyramid NN
struct NNState {
    float qx, qy, qz;
    float lx, ly, lz, hx, hy, hz;    // warp query box
    float best;
    unsigned long long bestkey;
    float4* slot;                    // this warp's 32-point staging tile in shared memory
};

template <bool IDX>
__device__ __forceinline__ void lg_scan_tile(const Pyramid& py, int t, NNState& s) {
    const int lane = threadIdx.x & 31;
    {   // per-lane point-to-box bound: skip if no lane can improve
        const float* b = py.box[0]; const int pd = py.pad[0];
        const float gx = gap(s.qx, s.qx, b[0 * pd + t], b[3 * pd + t]);
        const float gy = gap(s.qy, s.qy, b[1 * pd + t], b[4 * pd + t]);
        const float gz = gap(s.qz, s.qz, b[2 * pd + t], b[5 * pd + t]);
        const float lbp = fmaf(gz, gz, fmaf(gy, gy, gx * gx));
        if (__ballot_sync(KSS_FULL, lbp <= s.best * CULL_SLACK) == 0u) return;
    }
    __syncwarp();
    s.slot[lane] = py.tp[(size_t)t * TILE + lane];          // one coalesced 512 B load per tile
    __syncwarp();
#pragma unroll 8
    for (int j = 0; j < TILE; ++j) {
        const float4 p = s.slot[j];
        const float d = d2_rn(s.qx, s.qy, s.qz, p.x, p.y, p.z);
        if (IDX) {
            const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_uint(p.w);
            s.bestkey = key < s.bestkey ? key : s.bestkey;
        } else {
            s.best = fminf(s.best, d);
        }
    }
    if (IDX) s.best = __uint_as_float((unsigned)(s.bestkey >> 32));
}

// visit the children (level L-1 boxes, or tiles when L == 1 ... ) of node `node` at level L
template <int L, bool IDX>
__device__ __forceinline__ void lg_descend(const Pyramid& py, int node, NNState& s) {
    if constexpr (L == 0) {
        lg_scan_tile<IDX>(py, node, s);
    } else {
        const int lane = threadIdx.x & 31;
        const int c = node * 32 + lane;
        const int ccnt = py.cnt[L - 1], pd = py.pad[L - 1];
        const float* b = py.box[L - 1];
        unsigned lb = 0xffffffffu;
        if (c < ccnt) {
            const float gx = gap(s.lx, s.hx, b[0 * pd + c], b[3 * pd + c]);
            const float gy = gap(s.ly, s.hy, b[1 * pd + c], b[4 * pd + c]);
            const float gz = gap(s.lz, s.hz, b[2 * pd + c], b[5 * pd + c]);
            lb = __float_as_uint(fmaf(gz, gz, fmaf(gy, gy, gx * gx)));
        }
        for (;;) {
            const unsigned m = __reduce_min_sync(KSS_FULL, lb);
            if (m == 0xffffffffu) break;
            const float B = __uint_as_float(__reduce_max_sync(KSS_FULL, __float_as_uint(s.best)));
            if (__uint_as_float(m) > B * CULL_SLACK) break;
            const int src = __ffs(__ballot_sync(KSS_FULL, lb == m)) - 1;
            if (lane == src) lb = 0xffffffffu;
            lg_descend<L - 1, IDX>(py, node * 32 + src, s);
        }
    }
}

template <bool IDX>
__device__ __forceinline__ unsigned long long lg_warp_nn(const Pyramid& py, float qx, float qy, float qz, float4* slot) {
    NNState s;
    s.qx = qx; s.qy = qy; s.qz = qz;
    s.lx = warp_min_f(qx); s.ly = warp_min_f(qy); s.lz = warp_min_f(qz);
    s.hx = warp_max_f(qx); s.hy = warp_max_f(qy); s.hz = warp_max_f(qz);
    s.best = __int_as_float(0x7f800000);
    s.bestkey = 0xffffffffffffffffull;
    s.slot = slot;
    switch (py.nlev) {      // virtual root above the top level
        case 1: lg_descend<1, IDX>(py, 0, s); break;
        case 2: lg_descend<2, IDX>(py, 0, s); break;
        case 3: lg_descend<3, IDX>(py, 0, s); break;
        default: lg_descend<4, IDX>(py, 0, s); break;
    }
    if (!IDX) s.bestkey = (unsigned long long)__float_as_uint(s.best) << 32;
    return s.bestkey;
}

// ICP state kept on the device between kernels of one run
struct LgState {
    float Tk[16], fin[16];
    float smean[3], dmean[3];
    float one_over_n;
    double prev_mse, mse, fitness;
    int iters, done, converged, kept, apply_T;
    unsigned ticketA, ticketB, ticketF;
};

// NN kernel.  MODE 0: plain queries from q4 (float4 by original index) -> idx/d2 by original index
//             MODE 1: ICP iteration: lazily apply st->Tk to cur (in place), correspondences with rejection
//             MODE 2: fitness pass: query = st->fin * inp (original input), d2 only
template <int MODE>
__global__ void __launch_bounds__(LG_WARPS * 32)
lg_nn_kernel(Pyramid py, const int* __restrict__ perm, int n_q, float4* __restrict__ cur,
             const float4* __restrict__ inp, int* __restrict__ idx, float* __restrict__ d2out,
             LgState* __restrict__ st, double max_dist_sqr) {
    __shared__ float4 slots[LG_WARPS][TILE];
    __shared__ float T[16];
    if (MODE != 0) {
        if (st->done && MODE == 1) return;
        if (threadIdx.x < 16) T[threadIdx.x] = MODE == 1 ? st->Tk[threadIdx.x] : st->fin[threadIdx.x];
        __syncthreads();
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base = (blockIdx.x * LG_WARPS + warp) * 32;
    if (base >= n_q) return;
    const int j = min(base + lane, n_q - 1);
    const bool valid = base + lane < n_q;
    const int o = perm[j];
    float x, y, z;
    if (MODE == 2) {
        const float4 p = inp[o];
        xform_point(T, p.x, p.y, p.z, x, y, z);
    } else {
        const float4 p = cur[o];
        x = p.x; y = p.y; z = p.z;
        if (MODE == 1 && st->apply_T) {
            xform_point(T, p.x, p.y, p.z, x, y, z);     // transformCloud of the previous iteration (A.5)
            __syncwarp();
            if (valid) cur[o] = make_float4(x, y, z, p.w);
        }
    }
    const unsigned long long key = (MODE == 2) ? lg_warp_nn<false>(py, x, y, z, slots[warp])
                                               : lg_warp_nn<true>(py, x, y, z, slots[warp]);
    if (!valid) return;
    const float d2 = __uint_as_float((unsigned)(key >> 32));
    d2out[o] = d2;
    if (MODE == 0) idx[o] = (int)(key & 0xffffffffu);
    if (MODE == 1) idx[o] = ((double)d2 > max_dist_sqr) ? -1 : (int)(key & 0xffffffffu);   // A.3
}

// ------------------------------------------------------------------ canonical reductions, large n
// level 1: one warp per 256-element chunk of ORIGINAL indices (lane-strided partials + butterfly)
// upper levels: the last CTA to finish reduces the chunk partials by the same rule (<= 2 more levels)
template <int NQ>
__device__ __forceinline__ void lg_finish_f32(const float* __restrict__ part, int nchunks, float* out /* [NQ] smem */) {
    // called by one full CTA (256 threads = 8 warps); part is [nchunks][NQ]
    __shared__ float lvl2[256 * 16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int n2 = (nchunks + 255) >> 8;             // level-2 chunks (<= 256 supported -> n <= 16.7 M)
    for (int q = 0; q < NQ; ++q) {
        for (int c = warp; c < n2; c += nwarps) {
            float p = 0.0f;
            const int hi = min(nchunks, (c + 1) << 8);
            for (int i = (c << 8) + lane; i < hi; i += 32) p = __fadd_rn(p, part[(size_t)i * NQ + q]);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) p = __fadd_rn(p, __shfl_xor_sync(KSS_FULL, p, off));
            if (lane == 0) lvl2[c * 16 + q] = p;
        }
    }
    __syncthreads();
    if (warp == 0) {
        for (int q = 0; q < NQ; ++q) {
            float r;
            if (n2 == 1) r = lvl2[q];
            else {
                float p = 0.0f;
                for (int i = lane; i < n2; i += 32) p = __fadd_rn(p, lvl2[i * 16 + q]);
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) p = __fadd_rn(p, __shfl_xor_sync(KSS_FULL, p, off));
                r = p;
            }
            if (lane == 0) out[q] = r;
        }
    }
    __syncthreads();
}
__device__ __forceinline__ double lg_finish_f64(const double* __restrict__ part, int nchunks) {
    __shared__ double lvl2d[256];
    __shared__ double res;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int n2 = (nchunks + 255) >> 8;
    for (int c = warp; c < n2; c += nwarps) {
        double p = 0.0;
        const int hi = min(nchunks, (c + 1) << 8);
        for (int i = (c << 8) + lane; i < hi; i += 32) p = __dadd_rn(p, part[i]);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) p = __dadd_rn(p, __shfl_xor_sync(KSS_FULL, p, off));
        if (lane == 0) lvl2d[c] = p;
    }
    __syncthreads();
    if (warp == 0) {
        double r;
        if (n2 == 1) r = lvl2d[0];
        else {
            double p = 0.0;
            for (int i = lane; i < n2; i += 32) p = __dadd_rn(p, lvl2d[i]);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) p = __dadd_rn(p, __shfl_xor_sync(KSS_FULL, p, off));
            r = p;
        }
        if (lane == 0) res = r;
    }
    __syncthreads();
    return res;
}

__device__ __forceinline__ bool lg_last_block(unsigned* ticket) {
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
        if (last) *ticket = 0u;
    }
    __syncthreads();
    if (last) __threadfence();
    return last;
}

// pass A: per chunk sums of kept source xyz, matched target xyz (float), d2 (double), kept count
__global__ void __launch_bounds__(256)
lg_passA_kernel(const float4* __restrict__ cur, const float4* __restrict__ t_orig, const int* __restrict__ idx,
                const float* __restrict__ d2, int n, int nchunks, float* __restrict__ partA /* [nchunks][6] */,
                double* __restrict__ partD, int* __restrict__ partK, LgState* __restrict__ st) {
    if (st->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 8 + warp;
    if (c < nchunks) {
        float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        double dsum = 0.0;
        int k = 0;
        const int hi = min(n, (c + 1) << 8);
        for (int i = (c << 8) + lane; i < hi; i += 32) {
            const int m = idx[i];
            if (m < 0) continue;
            const float4 s = cur[i];
            const float4 t = t_orig[m];
            a[0] = __fadd_rn(a[0], s.x); a[1] = __fadd_rn(a[1], s.y); a[2] = __fadd_rn(a[2], s.z);
            a[3] = __fadd_rn(a[3], t.x); a[4] = __fadd_rn(a[4], t.y); a[5] = __fadd_rn(a[5], t.z);
            dsum = __dadd_rn(dsum, (double)d2[i]);
            ++k;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
            for (int q = 0; q < 6; ++q) a[q] = __fadd_rn(a[q], __shfl_xor_sync(KSS_FULL, a[q], off));
            dsum = __dadd_rn(dsum, __shfl_xor_sync(KSS_FULL, dsum, off));
        }
        k = __reduce_add_sync(KSS_FULL, k);
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 6; ++q) partA[(size_t)c * 6 + q] = a[q];
            partD[c] = dsum; partK[c] = k;
        }
    }
    if (!lg_last_block(&st->ticketA)) return;
    // ---- last CTA: upper levels, means, mse
    __shared__ float tot[16];
    __shared__ int kept_s;
    if (nchunks == 1) { if (threadIdx.x < 6) tot[threadIdx.x] = partA[threadIdx.x]; __syncthreads(); }
    else lg_finish_f32<6>(partA, nchunks, tot);
    const double dtot = nchunks == 1 ? partD[0] : lg_finish_f64(partD, nchunks);
    if (threadIdx.x == 0) kept_s = 0;
    __syncthreads();
    int k = 0;
    for (int i = threadIdx.x; i < nchunks; i += blockDim.x) k += partK[i];
    k = __reduce_add_sync(KSS_FULL, k);
    if ((threadIdx.x & 31) == 0 && k) atomicAdd(&kept_s, k);
    __syncthreads();
    if (threadIdx.x == 0) {
        const int cnt = kept_s;
        st->kept = cnt;
        if (cnt >= 3) {
            const float oon = div_(1.0f, (float)cnt);
            st->one_over_n = oon;
            for (int a2 = 0; a2 < 3; ++a2) { st->smean[a2] = mul_(tot[a2], oon); st->dmean[a2] = mul_(tot[3 + a2], oon); }
            st->mse = __ddiv_rn(dtot, (double)cnt);
        }
    }
}

// pass B: sigma partials, then (last CTA) umeyama + accumulate + convergence
__global__ void __launch_bounds__(256)
lg_passB_kernel(const float4* __restrict__ cur, const float4* __restrict__ t_orig, const int* __restrict__ idx,
                int n, int nchunks, float* __restrict__ partB /* [nchunks][9] */, LgState* __restrict__ st,
                int max_iter, double rot_thr, double trans_thr, double mse_rel, double mse_abs) {
    if (st->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 8 + warp;
    const bool enough = st->kept >= 3;
    if (c < nchunks && enough) {
        const float sm0 = st->smean[0], sm1 = st->smean[1], sm2 = st->smean[2];
        const float dm0 = st->dmean[0], dm1 = st->dmean[1], dm2 = st->dmean[2];
        float a[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const int hi = min(n, (c + 1) << 8);
        for (int i = (c << 8) + lane; i < hi; i += 32) {
            const int m = idx[i];
            if (m < 0) continue;
            const float4 s = cur[i];
            const float4 t = t_orig[m];
            const float sx = sub_(s.x, sm0), sy = sub_(s.y, sm1), sz = sub_(s.z, sm2);
            const float dx = sub_(t.x, dm0), dy = sub_(t.y, dm1), dz = sub_(t.z, dm2);
            a[0] = add_(a[0], mul_(dx, sx)); a[1] = add_(a[1], mul_(dx, sy)); a[2] = add_(a[2], mul_(dx, sz));
            a[3] = add_(a[3], mul_(dy, sx)); a[4] = add_(a[4], mul_(dy, sy)); a[5] = add_(a[5], mul_(dy, sz));
            a[6] = add_(a[6], mul_(dz, sx)); a[7] = add_(a[7], mul_(dz, sy)); a[8] = add_(a[8], mul_(dz, sz));
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
            for (int q = 0; q < 9; ++q) a[q] = __fadd_rn(a[q], __shfl_xor_sync(KSS_FULL, a[q], off));
        if (lane == 0)
#pragma unroll
            for (int q = 0; q < 9; ++q) partB[(size_t)c * 9 + q] = a[q];
    }
    if (!lg_last_block(&st->ticketB)) return;
    __shared__ float tot[16];
    if (enough) {
        if (nchunks == 1) { if (threadIdx.x < 9) tot[threadIdx.x] = partB[threadIdx.x]; __syncthreads(); }
        else lg_finish_f32<9>(partB, nchunks, tot);
    }
    if (threadIdx.x == 0) {
        if (!enough) { st->done = 1; st->converged = 0; return; }      // min_number_correspondences_
        float sigma[9], T[16], F[16], sm[3], dm[3];
        for (int i = 0; i < 9; ++i) sigma[i] = mul_(st->one_over_n, tot[i]);
        for (int i = 0; i < 3; ++i) { sm[i] = st->smean[i]; dm[i] = st->dmean[i]; }
        umeyama_finish(sigma, sm, dm, T);
        for (int i = 0; i < 16; ++i) F[i] = st->fin[i];
        mat4_mul(T, F, F);
        for (int i = 0; i < 16; ++i) { st->Tk[i] = T[i]; st->fin[i] = F[i]; }
        st->apply_T = 1;
        const int it = st->iters + 1;
        st->iters = it;
        const double mse = st->mse, prev = st->prev_mse;
        int dn = 0;
        if (it >= max_iter) dn = 1;
        else {
            const double cos_angle = 0.5 * (double)sub_(add_(add_(T[0], T[5]), T[10]), 1.0f);
            const double tr2 = (double)add_(add_(mul_(T[3], T[3]), mul_(T[7], T[7])), mul_(T[11], T[11]));
            if (cos_angle >= rot_thr && tr2 <= trans_thr) dn = 1;
            else if (fabs(__dsub_rn(mse, prev)) < mse_abs) dn = 1;
            else if (__ddiv_rn(fabs(__dsub_rn(mse, prev)), prev) < mse_rel) dn = 1;
            else st->prev_mse = mse;
        }
        if (dn) { st->done = 1; st->converged = 1; }
    }
}

// double sums of d2 (and sqrt d2) by original index: fitness (A.7) / PCR_QM (large clouds)
__global__ void __launch_bounds__(256)
lg_passF_kernel(const float* __restrict__ d2, int n, int nchunks, double* __restrict__ partD, double* __restrict__ partS,
                unsigned* __restrict__ ticket, double* __restrict__ out /* {sum d2 / n, sqrt(.), sum sqrt / n} */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 8 + warp;
    if (c < nchunks) {
        double a = 0.0, b = 0.0;
        const int hi = min(n, (c + 1) << 8);
        for (int i = (c << 8) + lane; i < hi; i += 32) {
            const double v = (double)d2[i];
            a = __dadd_rn(a, v);
            if (partS) b = __dadd_rn(b, __dsqrt_rn(v));
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            a = __dadd_rn(a, __shfl_xor_sync(KSS_FULL, a, off));
            b = __dadd_rn(b, __shfl_xor_sync(KSS_FULL, b, off));
        }
        if (lane == 0) { partD[c] = a; if (partS) partS[c] = b; }
    }
    if (!lg_last_block(ticket)) return;
    const double s1 = nchunks == 1 ? partD[0] : lg_finish_f64(partD, nchunks);
    double s2 = 0.0;
    if (partS) s2 = nchunks == 1 ? partS[0] : lg_finish_f64(partS, nchunks);
    if (threadIdx.x == 0) {
        const double mse = __ddiv_rn(s1, (double)n);
        out[0] = mse;
        out[1] = __dsqrt_rn(mse);
        out[2] = partS ? __ddiv_rn(s2, (double)n) : 0.0;
    }
}

__global__ void lg_state_init_kernel(LgState* st) {
    if (threadIdx.x < 16) { st->Tk[threadIdx.x] = st->fin[threadIdx.x] = (threadIdx.x % 5 == 0) ? 1.0f : 0.0f; }
    if (threadIdx.x == 0) {
        st->prev_mse = DBL_MAX; st->mse = 0; st->fitness = 0; st->iters = 0; st->done = 0; st->converged = 0;
        st->kept = 0; st->apply_T = 0; st->ticketA = st->ticketB = st->ticketF = 0u; st->one_over_n = 0.f;
    }
}

// ------------------------------------------------------------------ host orchestration
namespace {

struct Ctx {
    cudaStream_t st; long long* launches; const DevAlloc& alloc; int err = 0;
    template <class T> T* get(const char* name, size_t count) {
        void* p = nullptr;
        int r = alloc(name, count * sizeof(T), &p);
        if (r != KSS_OK) { err = r; return nullptr; }
        return reinterpret_cast<T*>(p);
    }
    bool ok() { if (cudaGetLastError() != cudaSuccess) err = KSS_ERR_CUDA; return err == 0; }
    void launched(int k = 1) { *launches += k; }
};

inline int cell_bits(int n) {
    int b = (int)std::ceil(std::log2((double)std::max(n, 2)) / 3.0) + 1;
    return std::min(7, std::max(2, b));
}

// Morton-cell bucket order of a cloud.  target: tp (padded) + t_orig ; source: perm only.
int order_cloud(Ctx& c, const char* tag, const double* d_pts, int n, float4* p4, float4* tp, int npad, int* perm) {
    char nm[64];
    const int bits = cell_bits(n);
    const int m = 1 << (3 * bits);
    snprintf(nm, sizeof(nm), "%s_bb", tag);   unsigned* bb = c.get<unsigned>(nm, 8);
    snprintf(nm, sizeof(nm), "%s_cell", tag); unsigned* cellid = c.get<unsigned>(nm, n);
    snprintf(nm, sizeof(nm), "%s_hist", tag); unsigned* hist = c.get<unsigned>(nm, m);
    snprintf(nm, sizeof(nm), "%s_tot", tag);  unsigned* tot = c.get<unsigned>(nm, 4096);
    if (c.err) return c.err;
    const int nb = (m + 1023) / 1024;
    cudaMemsetAsync(hist, 0, sizeof(unsigned) * m, c.st);
    lg_init_bbox_kernel<<<1, 32, 0, c.st>>>(bb);
    lg_convert_bbox_kernel<<<(n + 255) / 256, 256, 0, c.st>>>(d_pts, n, p4, bb);
    lg_hist_kernel<<<(n + 255) / 256, 256, 0, c.st>>>(p4, n, bb, bits, cellid, hist);
    lg_scan1_kernel<<<nb, 1024, 0, c.st>>>(hist, m, tot);
    lg_scan2_kernel<<<1, 1024, 0, c.st>>>(tot, nb);
    lg_scan3_kernel<<<nb, 1024, 0, c.st>>>(hist, m, tot);
    lg_scatter_kernel<<<(n + 255) / 256, 256, 0, c.st>>>(p4, n, cellid, hist, tp, perm);
    c.launched(7);
    if (tp && npad > n) { lg_pad_kernel<<<1, 32, 0, c.st>>>(tp, n, npad); c.launched(); }
    return c.ok() ? KSS_OK : c.err;
}

int build_pyramid(Ctx& c, const double* d_t, int n_t, Pyramid* py, float4** t_orig_out) {
    const int npad = (n_t + 31) / 32 * 32;
    float4* t_orig = c.get<float4>("lg_t_orig", n_t);
    float4* tp = c.get<float4>("lg_tp", npad);
    if (c.err) return c.err;
    int r = order_cloud(c, "lg_t", d_t, n_t, t_orig, tp, npad, nullptr);
    if (r) return r;
    py->tp = tp;
    int cnt = npad / 32, lev = 0;
    float* boxes[LG_MAX_LEVELS];
    for (;;) {
        if (lev >= LG_MAX_LEVELS) return KSS_ERR_UNSUPPORTED;
        char nm[32]; snprintf(nm, sizeof(nm), "lg_box%d", lev);
        const int pad = (cnt + 31) / 32 * 32;
        boxes[lev] = c.get<float>(nm, (size_t)6 * pad);
        if (c.err) return c.err;
        py->cnt[lev] = cnt; py->pad[lev] = pad; py->box[lev] = boxes[lev];
        if (lev == 0) lg_tile_box_kernel<<<(cnt * 32 + 255) / 256, 256, 0, c.st>>>(tp, n_t, cnt, pad, boxes[0]);
        else lg_up_box_kernel<<<(cnt * 32 + 255) / 256, 256, 0, c.st>>>(boxes[lev - 1], py->cnt[lev - 1], py->pad[lev - 1], cnt, pad, boxes[lev]);
        c.launched();
        ++lev;
        if (cnt <= 32) break;
        cnt = (cnt + 31) / 32;
    }
    py->nlev = lev;
    for (int l = lev; l < LG_MAX_LEVELS; ++l) { py->cnt[l] = 0; py->pad[l] = 0; py->box[l] = nullptr; }
    if (t_orig_out) *t_orig_out = t_orig;
    return c.ok() ? KSS_OK : c.err;
}

inline int nn_grid(int n_q) { return (n_q + LG_WARPS * 32 - 1) / (LG_WARPS * 32); }

}  // namespace

int large_nn_device(cudaStream_t st, long long* launches, const double* d_q, int n_q, const double* d_t, int n_t,
                    int* d_idx, float* d_d2, const DevAlloc& alloc) {
    Ctx c{st, launches, alloc};
    Pyramid py;
    int r = build_pyramid(c, d_t, n_t, &py, nullptr);
    if (r) return r;
    float4* q4 = c.get<float4>("lg_q4", n_q);
    int* perm = c.get<int>("lg_perm", n_q);
    if (c.err) return c.err;
    r = order_cloud(c, "lg_q", d_q, n_q, q4, nullptr, 0, perm);
    if (r) return r;
    lg_nn_kernel<0><<<nn_grid(n_q), LG_WARPS * 32, 0, st>>>(py, perm, n_q, q4, nullptr, d_idx, d_d2, nullptr, 0.0);
    c.launched();
    return c.ok() ? KSS_OK : c.err;
}

int large_metrics_device(cudaStream_t st, long long* launches, const double* d_q, const int* d_cnt_q, int cap_q,
                         const double* d_t, const int* d_cnt_t, int cap_t, double* d_out3, const DevAlloc& alloc) {
    // ragged large clouds would need the counts on the host: the batch API passes full capacity here
    (void)d_cnt_q; (void)d_cnt_t;
    Ctx c{st, launches, alloc};
    const int n_q = cap_q, n_t = cap_t;
    int* idx = c.get<int>("lg_idx", n_q);
    float* d2 = c.get<float>("lg_d2", n_q);
    if (c.err) return c.err;
    int r = large_nn_device(st, launches, d_q, n_q, d_t, n_t, idx, d2, alloc);
    if (r) return r;
    const int nchunks = (n_q + 255) / 256;
    double* pd = c.get<double>("lg_partD", nchunks);
    double* ps = c.get<double>("lg_partS", nchunks);
    unsigned* tk = c.get<unsigned>("lg_ticket", 4);
    if (c.err) return c.err;
    cudaMemsetAsync(tk, 0, 16, st);
    lg_passF_kernel<<<(nchunks + 7) / 8, 256, 0, st>>>(d2, n_q, nchunks, pd, ps, tk, d_out3);
    c.launched();
    return c.ok() ? KSS_OK : c.err;
}

int large_icp_host(cudaStream_t st, long long* launches, const double* src, int n_s, const double* tgt, int n_t,
                   const kss_icp_params* prm, float T[16], double* fitness, int* iters, int* converged,
                   const DevAlloc& alloc) {
    Ctx c{st, launches, alloc};
    double* d_s = c.get<double>("lg_in_s", (size_t)n_s * 3);
    double* d_t = c.get<double>("lg_in_t", (size_t)n_t * 3);
    if (c.err) return c.err;
    cudaMemcpyAsync(d_s, src, sizeof(double) * 3 * (size_t)n_s, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_t, tgt, sizeof(double) * 3 * (size_t)n_t, cudaMemcpyHostToDevice, st);
    LargeIcp run;
    int r = large_icp_prepare(st, launches, d_s, n_s, d_t, n_t, alloc, &run);
    if (r) return r;
    r = large_icp_run(st, launches, &run, prm, 0);
    if (r) return r;
    return large_icp_result(st, &run, T, fitness, iters, converged);
}

// ---- reusable pieces (bench.py times large_icp_iterations on a prepared run)
int large_icp_prepare(cudaStream_t st, long long* launches, const double* d_s, int n_s, const double* d_t, int n_t,
                      const DevAlloc& alloc, LargeIcp* run) {
    Ctx c{st, launches, alloc};
    Pyramid py;
    float4* t_orig = nullptr;
    int r = build_pyramid(c, d_t, n_t, &py, &t_orig);
    if (r) return r;
    float4* inp = c.get<float4>("lg_inp", n_s);
    float4* cur = c.get<float4>("lg_cur", n_s);
    int* perm = c.get<int>("lg_perm", n_s);
    int* idx = c.get<int>("lg_idx", n_s);
    float* d2 = c.get<float>("lg_d2", n_s);
    const int nchunks = (n_s + 255) / 256;
    float* pA = c.get<float>("lg_partA", (size_t)nchunks * 6);
    float* pB = c.get<float>("lg_partB", (size_t)nchunks * 9);
    double* pD = c.get<double>("lg_partD", nchunks);
    int* pK = c.get<int>("lg_partK", nchunks);
    LgState* state = c.get<LgState>("lg_state", 1);
    double* out3 = c.get<double>("lg_out3", 4);
    if (c.err) return c.err;
    r = order_cloud(c, "lg_q", d_s, n_s, inp, nullptr, 0, perm);
    if (r) return r;
    cudaMemcpyAsync(cur, inp, sizeof(float4) * (size_t)n_s, cudaMemcpyDeviceToDevice, st);
    lg_state_init_kernel<<<1, 32, 0, st>>>(state);
    c.launched();
    static_assert(sizeof(Pyramid) <= sizeof(run->pyramid), "LargeIcp::pyramid too small");
    memcpy(run->pyramid, &py, sizeof(py));
    run->n_s = n_s; run->n_t = n_t; run->nchunks = nchunks;
    run->t_orig = t_orig; run->inp = inp; run->cur = cur; run->perm = perm; run->idx = idx; run->d2 = d2;
    run->partA = pA; run->partB = pB; run->partD = pD; run->partK = pK; run->state = state; run->out3 = out3;
    return c.ok() ? KSS_OK : c.err;
}

// enqueue `count` ICP iterations (each = NN + pass A + pass B); kernels are no-ops once the run is done
int large_icp_iterations(cudaStream_t st, long long* launches, LargeIcp* run, const kss_icp_params* prm, int count) {
    Pyramid py; memcpy(&py, run->pyramid, sizeof(py));
    LgState* state = (LgState*)run->state;
    const double max2 = prm->max_corr_dist * prm->max_corr_dist;
    const int n = run->n_s, nch = run->nchunks;
    const double mse_abs = prm->fitness_eps < 0.0 ? -1.0 : 1e-12;   // fitness_eps < 0: never converge (steady-state timing)
    for (int k = 0; k < count; ++k) {
        if (run->mark) run->mark(run->mark_user, KSS_STAGE_LARGE_NN, 1);
        lg_nn_kernel<1><<<nn_grid(n), LG_WARPS * 32, 0, st>>>(py, run->perm, n, (float4*)run->cur, nullptr, run->idx,
                                                              run->d2, state, max2);
        if (run->mark) { run->mark(run->mark_user, KSS_STAGE_LARGE_NN, 0); run->mark(run->mark_user, KSS_STAGE_LARGE_REDUCE, 1); }
        lg_passA_kernel<<<(nch + 7) / 8, 256, 0, st>>>((const float4*)run->cur, (const float4*)run->t_orig, run->idx,
                                                       run->d2, n, nch, run->partA, run->partD, run->partK, state);
        lg_passB_kernel<<<(nch + 7) / 8, 256, 0, st>>>((const float4*)run->cur, (const float4*)run->t_orig, run->idx, n,
                                                       nch, run->partB, state, prm->max_iterations,
                                                       1.0 - prm->transformation_eps, prm->transformation_eps,
                                                       prm->fitness_eps, mse_abs);
        if (run->mark) run->mark(run->mark_user, KSS_STAGE_LARGE_REDUCE, 0);
        *launches += 3;
    }
    return cudaGetLastError() == cudaSuccess ? KSS_OK : KSS_ERR_CUDA;
}

int large_icp_run(cudaStream_t st, long long* launches, LargeIcp* run, const kss_icp_params* prm, int poll) {
    if (poll <= 0) poll = 4;
    LgState* state = (LgState*)run->state;
    int done = 0, issued = 0;
    while (!done && issued < prm->max_iterations) {
        const int k = std::min(poll, prm->max_iterations - issued);
        int r = large_icp_iterations(st, launches, run, prm, k);
        if (r) return r;
        issued += k;
        if (cudaMemcpyAsync(&done, &state->done, sizeof(int), cudaMemcpyDeviceToHost, st) != cudaSuccess) return KSS_ERR_CUDA;
        if (cudaStreamSynchronize(st) != cudaSuccess) return KSS_ERR_CUDA;
    }
    // getFitnessScore: final * ORIGINAL input, NN, mean d2 in double (A.7)
    Pyramid py; memcpy(&py, run->pyramid, sizeof(py));
    lg_nn_kernel<2><<<nn_grid(run->n_s), LG_WARPS * 32, 0, st>>>(py, run->perm, run->n_s, nullptr, (const float4*)run->inp,
                                                               nullptr, run->d2, state, 0.0);
    lg_passF_kernel<<<(run->nchunks + 7) / 8, 256, 0, st>>>(run->d2, run->n_s, run->nchunks, run->partD, nullptr,
                                                            &state->ticketF, run->out3);
    *launches += 2;
    return cudaGetLastError() == cudaSuccess ? KSS_OK : KSS_ERR_CUDA;
}

int large_icp_result(cudaStream_t st, LargeIcp* run, float T[16], double* fitness, int* iters, int* converged) {
    LgState h;
    double o3[3];
    if (cudaMemcpyAsync(&h, run->state, sizeof(LgState), cudaMemcpyDeviceToHost, st) != cudaSuccess) return KSS_ERR_CUDA;
    if (cudaMemcpyAsync(o3, run->out3, sizeof(o3), cudaMemcpyDeviceToHost, st) != cudaSuccess) return KSS_ERR_CUDA;
    if (cudaStreamSynchronize(st) != cudaSuccess) return KSS_ERR_CUDA;
    if (T) for (int i = 0; i < 16; ++i) T[i] = h.fin[i];
    if (fitness) *fitness = o3[0];
    if (iters) *iters = h.iters;
    if (converged) *converged = h.converged;
    return KSS_OK;
}

}  // namespace kss
