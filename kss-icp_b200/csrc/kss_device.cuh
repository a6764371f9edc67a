// kss_device.cuh -- device-side building blocks of the KSS-ICP B200 path.
//
// Arithmetic contract (DESIGN.md "numerics"): every parity-relevant expression is
// written with explicit round-to-nearest intrinsics (__fmul_rn/__fadd_rn/__dmul_rn/
// __dadd_rn) so that no FMA contraction can change a result, independent of -fmad.
// fmaf() is used only in conservative lower bounds that never reach an output.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include "kss_kernels.h"

#define KSS_FULL 0xffffffffu

namespace kss {

constexpr float CULL_SLACK = 1.00001f;       // lower bounds must exceed best*slack before a tile is skipped
constexpr float PAD_COORD = 1.0e18f;         // coordinate of padding targets (d2 ~ 3e36, finite)

// ---------------------------------------------------------------- ordering helpers
__device__ __forceinline__ unsigned f2ord(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ float warp_min_f(float v) { return ord2f(__reduce_min_sync(KSS_FULL, f2ord(v))); }
__device__ __forceinline__ float warp_max_f(float v) { return ord2f(__reduce_max_sync(KSS_FULL, f2ord(v))); }

// ---------------------------------------------------------------- exact distance
// FLANN L2_Simple<float>: ((dx*dx) + dy*dy) + dz*dz, one rounding per operator
// (replaces pcl::KdTreeFLANN distance; initRegistrationKSS.hpp:443, registrationMeasure.hpp:79)
__device__ __forceinline__ float d2_rn(float qx, float qy, float qz, float tx, float ty, float tz) {
    float dx = __fsub_rn(qx, tx), dy = __fsub_rn(qy, ty), dz = __fsub_rn(qz, tz);
    float r = __fmul_rn(dx, dx);
    r = __fadd_rn(r, __fmul_rn(dy, dy));
    r = __fadd_rn(r, __fmul_rn(dz, dz));
    return r;
}

// ---------------------------------------------------------------- target tiles in shared memory
// tgt : float4 {x, y, z, bits(original index)} in Morton order, padded to a multiple of 32
// box : SoA [6][MAX_TILES] = minx,miny,minz,maxx,maxy,maxz of each 32-point tile
struct TileView {
    const float4* tgt;
    const float* box;
    int ntiles;
};

__device__ __forceinline__ float gap(float lo, float hi, float blo, float bhi) {
    // distance between intervals [lo,hi] and [blo,bhi] (0 if they overlap)
    return fmaxf(0.0f, fmaxf(blo - hi, lo - bhi));
}

__device__ __forceinline__ float boxbox_lb(const float* box, int t, float lx, float ly, float lz,
                                           float hx, float hy, float hz) {
    float gx = gap(lx, hx, box[0 * MAX_TILES + t], box[3 * MAX_TILES + t]);
    float gy = gap(ly, hy, box[1 * MAX_TILES + t], box[4 * MAX_TILES + t]);
    float gz = gap(lz, hz, box[2 * MAX_TILES + t], box[5 * MAX_TILES + t]);
    return fmaf(gz, gz, fmaf(gy, gy, gx * gx));
}

// Exact 1-NN of one query per lane against all tiles, warp-synchronous best-first
// traversal with conservative culling.  All 32 lanes must be active and should carry
// spatially coherent queries (Morton-consecutive points).  Result: bit-exact minimum of
// d2_rn over ALL target points; WITH_INDEX additionally returns the lowest original
// index among fp32-equal minima (the oracle's tie rule) packed as (d2bits<<32)|orig.
template <bool WITH_INDEX>
__device__ __forceinline__ unsigned long long warp_nn(const TileView& tv, float qx, float qy, float qz) {
    const int lane = threadIdx.x & 31;
    // warp query box
    const float lx = warp_min_f(qx), ly = warp_min_f(qy), lz = warp_min_f(qz);
    const float hx = warp_max_f(qx), hy = warp_max_f(qy), hz = warp_max_f(qz);
    // each lane owns tiles lane and lane+32 ; lower bounds kept as uint (>=0 floats order as uints)
    unsigned lbA = 0xffffffffu, lbB = 0xffffffffu;
    if (lane < tv.ntiles) lbA = __float_as_uint(boxbox_lb(tv.box, lane, lx, ly, lz, hx, hy, hz));
    if (lane + 32 < tv.ntiles) lbB = __float_as_uint(boxbox_lb(tv.box, lane + 32, lx, ly, lz, hx, hy, hz));

    float best = __int_as_float(0x7f800000);  // +inf
    unsigned long long bestkey = 0xffffffffffffffffull;

    for (;;) {
        const unsigned c = min(lbA, lbB);
        const unsigned m = __reduce_min_sync(KSS_FULL, c);
        if (m == 0xffffffffu) break;
        const float B = __uint_as_float(__reduce_max_sync(KSS_FULL, __float_as_uint(best)));
        if (__uint_as_float(m) > B * CULL_SLACK) break;
        const unsigned bal = __ballot_sync(KSS_FULL, c == m);
        const int src = __ffs(bal) - 1;
        int t = (lbA == m) ? lane : lane + 32;
        t = __shfl_sync(KSS_FULL, t, src);
        if (lane == src) { if (lbA == m) lbA = 0xffffffffu; else lbB = 0xffffffffu; }
        // per-lane point-to-box bound: skip the tile if it cannot improve any lane
        {
            float gx = gap(qx, qx, tv.box[0 * MAX_TILES + t], tv.box[3 * MAX_TILES + t]);
            float gy = gap(qy, qy, tv.box[1 * MAX_TILES + t], tv.box[4 * MAX_TILES + t]);
            float gz = gap(qz, qz, tv.box[2 * MAX_TILES + t], tv.box[5 * MAX_TILES + t]);
            float lbp = fmaf(gz, gz, fmaf(gy, gy, gx * gx));
            if (__ballot_sync(KSS_FULL, lbp <= best * CULL_SLACK) == 0u) continue;
        }
        const float4* tp = tv.tgt + t * TILE;
#pragma unroll 8
        for (int j = 0; j < TILE; ++j) {
            const float4 p = tp[j];  // warp-uniform address: one broadcast LDS.128 per 32 distances
            const float d = d2_rn(qx, qy, qz, p.x, p.y, p.z);
            if (WITH_INDEX) {
                const unsigned long long key =
                    ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_uint(p.w);
                bestkey = key < bestkey ? key : bestkey;
            } else {
                best = fminf(best, d);
            }
        }
        if (WITH_INDEX) best = __uint_as_float((unsigned)(bestkey >> 32));
    }
    if (!WITH_INDEX) bestkey = (unsigned long long)__float_as_uint(best) << 32;
    return bestkey;
}

// ---------------------------------------------------------------- canonical reductions (CANON256)
// Order contract (DESIGN.md "reduction order contract", also implemented by the CPU checker): per 256-element
// chunk, lane l accumulates elements i = l (mod 32) in increasing i starting from +0, then an
// xor butterfly 16,8,4,2,1; chunk results are reduced by the same rule.  One warp runs one
// reduction; n <= 8192 here (two levels).  get(i, v) returns false for masked slots.
template <class Get>
__device__ __forceinline__ float canon_sum_warp_f32(int n, Get get) {
    const int lane = threadIdx.x & 31;
    const int nc = (n + 255) >> 8;
    float acc = 0.0f;
    float p = 0.0f;
    for (int c = 0; c < nc || c == 0; ++c) {
        p = 0.0f;
        const int hi = min(n, (c + 1) << 8);
        for (int i = (c << 8) + lane; i < hi; i += 32) {
            float v;
            if (get(i, v)) p = __fadd_rn(p, v);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) p = __fadd_rn(p, __shfl_xor_sync(KSS_FULL, p, off));
        if (lane == c) acc = __fadd_rn(0.0f, p);
        if (nc <= 1) return p;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(KSS_FULL, acc, off));
    return acc;
}

template <class Get>
__device__ __forceinline__ double canon_sum_warp_f64(int n, Get get) {
    const int lane = threadIdx.x & 31;
    const int nc = (n + 255) >> 8;
    double acc = 0.0;
    double p = 0.0;
    for (int c = 0; c < nc || c == 0; ++c) {
        p = 0.0;
        const int hi = min(n, (c + 1) << 8);
        for (int i = (c << 8) + lane; i < hi; i += 32) {
            double v;
            if (get(i, v)) p = __dadd_rn(p, v);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) p = __dadd_rn(p, __shfl_xor_sync(KSS_FULL, p, off));
        if (lane == c) acc = __dadd_rn(0.0, p);
        if (nc <= 1) return p;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(KSS_FULL, acc, off));
    return acc;
}

// ---------------------------------------------------------------- 3x3 SVD / Kabsch (one thread per problem)
// Two-sided Jacobi SVD of a real 3x3 in float, the small-fixed-size algorithm behind
// Eigen::JacobiSVD which pcl::umeyama calls.  Operation order is the
// contract with the oracle's svd3(); all ops are single-rounded.
struct Rot2 { float c, s; };

__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float sqrt_(float a) { return __fsqrt_rn(a); }

// All indices below are compile-time constants so that W, U, V stay in registers: the SVD runs on ONE thread between
// two barriers of the ICP loop, and local-memory round trips in its dependency chain were a third of the loop time.
template <int P, int Q>
__device__ __forceinline__ void rot_left(float (&W)[9], Rot2 j) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float x = W[3 * P + k], y = W[3 * Q + k];
        W[3 * P + k] = add_(mul_(j.c, x), mul_(j.s, y));
        W[3 * Q + k] = add_(mul_(-j.s, x), mul_(j.c, y));
    }
}
template <int P, int Q>
__device__ __forceinline__ void rot_right(float (&W)[9], Rot2 j) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float x = W[3 * k + P], y = W[3 * k + Q];
        W[3 * k + P] = sub_(mul_(j.c, x), mul_(j.s, y));
        W[3 * k + Q] = add_(mul_(j.s, x), mul_(j.c, y));
    }
}
__device__ __forceinline__ Rot2 sym_jacobi(float x, float y, float z) {
    Rot2 r;
    float deno = mul_(2.0f, fabsf(y));
    if (deno < FLT_MIN) { r.c = 1.0f; r.s = 0.0f; return r; }
    float tau = div_(sub_(x, z), deno);
    float w = sqrt_(add_(mul_(tau, tau), 1.0f));
    float t = tau > 0.0f ? div_(1.0f, add_(tau, w)) : div_(1.0f, sub_(tau, w));
    float sign_t = t > 0.0f ? 1.0f : -1.0f;
    float n = div_(1.0f, sqrt_(add_(mul_(t, t), 1.0f)));
    r.s = mul_(mul_(mul_(-sign_t, div_(y, fabsf(y))), fabsf(t)), n);
    r.c = n;
    return r;
}

// one two-sided Jacobi step on the (P,Q) block; returns false when the block is already diagonal enough
template <int P, int Q>
__device__ __forceinline__ bool svd3_step(float (&W)[9], float (&U)[9], float (&V)[9], float& maxDiag) {
    const float precision = 2.0f * FLT_EPSILON;
    const float tiny = FLT_MIN;
    const float thr = fmaxf(tiny, mul_(precision, maxDiag));
    if (!(fabsf(W[3 * P + Q]) > thr || fabsf(W[3 * Q + P]) > thr)) return false;
    const float m00 = W[3 * P + P], m01 = W[3 * P + Q], m10 = W[3 * Q + P], m11 = W[3 * Q + Q];
    Rot2 r1;
    const float t = add_(m00, m11), d = sub_(m10, m01);
    if (fabsf(d) < tiny) { r1.s = 0.0f; r1.c = 1.0f; }
    else {
        const float u = div_(t, d);
        const float tmp = sqrt_(add_(1.0f, mul_(u, u)));
        r1.s = div_(1.0f, tmp); r1.c = div_(u, tmp);
    }
    const float n00 = add_(mul_(r1.c, m00), mul_(r1.s, m10));
    const float n01 = add_(mul_(r1.c, m01), mul_(r1.s, m11));
    const float n11 = add_(mul_(-r1.s, m01), mul_(r1.c, m11));
    const Rot2 jr = sym_jacobi(n00, n01, n11);
    Rot2 jl;
    jl.c = sub_(mul_(r1.c, jr.c), mul_(r1.s, -jr.s));
    jl.s = add_(mul_(r1.c, -jr.s), mul_(r1.s, jr.c));
    rot_left<P, Q>(W, jl);
    Rot2 jlt; jlt.c = jl.c; jlt.s = -jl.s;
    rot_right<P, Q>(U, jlt);
    rot_right<P, Q>(W, jr);
    rot_right<P, Q>(V, jr);
    maxDiag = fmaxf(maxDiag, fmaxf(fabsf(W[3 * P + P]), fabsf(W[3 * Q + Q])));
    return true;
}

template <int I, int J>
__device__ __forceinline__ void svd3_swap(float (&s)[3], float (&U)[9], float (&V)[9]) {
    const float ts = s[I]; s[I] = s[J]; s[J] = ts;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float tu = U[3 * k + I]; U[3 * k + I] = U[3 * k + J]; U[3 * k + J] = tu;
        const float tv = V[3 * k + I]; V[3 * k + I] = V[3 * k + J]; V[3 * k + J] = tv;
    }
}

__device__ __forceinline__ void svd3(const float (&A)[9], float (&U)[9], float (&s)[3], float (&V)[9]) {
    float scale = 0.0f;
#pragma unroll
    for (int i = 0; i < 9; ++i) scale = fmaxf(scale, fabsf(A[i]));
    if (!(scale > 0.0f)) scale = 1.0f;
    float W[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) W[i] = div_(A[i], scale);
#pragma unroll
    for (int i = 0; i < 9; ++i) { U[i] = (i % 4 == 0) ? 1.0f : 0.0f; V[i] = U[i]; }
    float maxDiag = fmaxf(fabsf(W[0]), fmaxf(fabsf(W[4]), fabsf(W[8])));
    bool finished = false;
    int guard = 0;
    while (!finished && guard++ < 64) {
        // sweep order (p,q) = (1,0), (2,0), (2,1)
        const bool a = svd3_step<1, 0>(W, U, V, maxDiag);
        const bool b = svd3_step<2, 0>(W, U, V, maxDiag);
        const bool c = svd3_step<2, 1>(W, U, V, maxDiag);
        finished = !(a || b || c);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a = W[4 * i];
        s[i] = fabsf(a);
        if (a < 0.0f) {
#pragma unroll
            for (int k = 0; k < 3; ++k) U[3 * k + i] = -U[3 * k + i];
        }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) s[i] = mul_(s[i], scale);
    // selection sort, descending: position of the first strict maximum of the tail; stop at a zero maximum
    {
        const int pos = (s[2] > (s[1] > s[0] ? s[1] : s[0])) ? 2 : (s[1] > s[0] ? 1 : 0);
        const float m = pos == 2 ? s[2] : (pos == 1 ? s[1] : s[0]);
        if (m == 0.0f) return;
        if (pos == 1) svd3_swap<0, 1>(s, U, V);
        else if (pos == 2) svd3_swap<0, 2>(s, U, V);
    }
    {
        const bool two = s[2] > s[1];
        if ((two ? s[2] : s[1]) == 0.0f) return;
        if (two) svd3_swap<1, 2>(s, U, V);
    }
}

__device__ __forceinline__ float det3(const float (&m)[9]) {
    float h0 = mul_(m[0], sub_(mul_(m[4], m[8]), mul_(m[5], m[7])));
    float h1 = mul_(m[1], sub_(mul_(m[3], m[8]), mul_(m[5], m[6])));
    float h2 = mul_(m[2], sub_(mul_(m[3], m[7]), mul_(m[4], m[6])));
    return add_(sub_(h0, h1), h2);
}

// pcl::umeyama tail: sigma, means -> 4x4 row-major rigid transform
__device__ __forceinline__ void umeyama_finish(const float (&sigma)[9], const float (&smean)[3], const float (&dmean)[3], float (&T)[16]) {
    float U[9], sv[3], V[9];
    svd3(sigma, U, sv, V);
    float S[3] = {1.0f, 1.0f, 1.0f};
    if (det3(sigma) < 0.0f) S[2] = -1.0f;
    int rank = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i)
        if (!(fabsf(sv[i]) <= mul_(fabsf(sv[0]), 1e-5f))) ++rank;
    if (rank == 2) {
        if (mul_(det3(U), det3(V)) > 0.0f) { S[0] = S[1] = S[2] = 1.0f; }
        else { S[2] = -1.0f; }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) T[i] = 0.0f;
    T[15] = 1.0f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            float r = mul_(mul_(U[3 * a + 0], S[0]), V[3 * b + 0]);
            r = add_(r, mul_(mul_(U[3 * a + 1], S[1]), V[3 * b + 1]));
            r = add_(r, mul_(mul_(U[3 * a + 2], S[2]), V[3 * b + 2]));
            T[4 * a + b] = r;
        }
        float rs = mul_(T[4 * a + 0], smean[0]);
        rs = add_(rs, mul_(T[4 * a + 1], smean[1]));
        rs = add_(rs, mul_(T[4 * a + 2], smean[2]));
        T[4 * a + 3] = sub_(dmean[a], rs);
    }
}

// Matrix4f * (x,y,z,1) : ((m0*x + m1*y) + m2*z) + m3   
__device__ __forceinline__ void xform_point(const float* T, float x, float y, float z,
                                            float& ox, float& oy, float& oz) {
    ox = add_(add_(add_(mul_(T[0], x), mul_(T[1], y)), mul_(T[2], z)), T[3]);
    oy = add_(add_(add_(mul_(T[4], x), mul_(T[5], y)), mul_(T[6], z)), T[7]);
    oz = add_(add_(add_(mul_(T[8], x), mul_(T[9], y)), mul_(T[10], z)), T[11]);
}

__device__ __forceinline__ void mat4_mul(const float (&A)[16], const float (&B)[16], float (&C)[16]) {
    float R[16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float r = mul_(A[4 * i + 0], B[0 + j]);
            r = add_(r, mul_(A[4 * i + 1], B[4 + j]));
            r = add_(r, mul_(A[4 * i + 2], B[8 + j]));
            r = add_(r, mul_(A[4 * i + 3], B[12 + j]));
            R[4 * i + j] = r;
        }
#pragma unroll
    for (int i = 0; i < 16; ++i) C[i] = R[i];
}

// ---------------------------------------------------------------- fp64 similarity pieces
// initRegistration_Transfer (initRegistrationKSS.hpp:365-404) with host-supplied cos/sin
__device__ __forceinline__ void rot_x(double c, double s, double& y, double& z) {
    double ny = __dsub_rn(__dmul_rn(y, c), __dmul_rn(z, s));
    double nz = __dadd_rn(__dmul_rn(y, s), __dmul_rn(z, c));
    y = ny; z = nz;
}
__device__ __forceinline__ void rot_y(double c, double s, double& x, double& z) {
    double nx = __dadd_rn(__dmul_rn(z, s), __dmul_rn(x, c));
    double nz = __dsub_rn(__dmul_rn(z, c), __dmul_rn(x, s));
    x = nx; z = nz;
}
__device__ __forceinline__ void rot_z(double c, double s, double& x, double& y) {
    double nx = __dsub_rn(__dmul_rn(x, c), __dmul_rn(y, s));
    double ny = __dadd_rn(__dmul_rn(x, s), __dmul_rn(y, c));
    x = nx; y = ny;
}
// initRegistrationKSS.hpp:77-84 : p + middle ; c + (p - c) * scale
__device__ __forceinline__ double align_coord(double v, double mid_s, double mid, double scale) {
    double w = __dadd_rn(v, mid);
    return __dadd_rn(mid_s, __dmul_rn(__dsub_rn(w, mid_s), scale));
}

// Morton helpers (ordering only; never affects results)
__device__ __forceinline__ unsigned spread7(unsigned v) {  // 7 bits -> every third bit
    unsigned r = 0;
#pragma unroll
    for (int b = 0; b < 7; ++b) r |= ((v >> b) & 1u) << (3 * b);
    return r;
}
__device__ __forceinline__ unsigned morton21(float x, float y, float z, const float* lo, const float* inv) {
    unsigned ix = (unsigned)fminf(fmaxf((x - lo[0]) * inv[0], 0.0f), 127.0f);
    unsigned iy = (unsigned)fminf(fmaxf((y - lo[1]) * inv[1], 0.0f), 127.0f);
    unsigned iz = (unsigned)fminf(fmaxf((z - lo[2]) * inv[2], 0.0f), 127.0f);
    return spread7(ix) | (spread7(iy) << 1) | (spread7(iz) << 2);
}

// in-place bitonic sort of NP (power of two) uint keys in shared memory, ascending
template <int NP>
__device__ __forceinline__ void bitonic_sort_smem(unsigned* keys) {
    for (int k = 2; k <= NP; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < NP; i += blockDim.x) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned a = keys[i], b = keys[ixj];
                    bool up = ((i & k) == 0);
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
        }
    __syncthreads();
}

}  // namespace kss
