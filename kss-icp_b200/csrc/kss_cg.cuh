// kss_cg.cuh -- the "candidate grid": an exact per-pair NN accelerator for small targets.
//
// One target cloud (<= 2048 points) serves ~1e6 queries per registration (729 hypotheses x n_s
// sweep queries + every ICP iteration), all inside the ball |q| <= R around the world origin
// (the sweep rotates about the origin, initRegistrationKSS.hpp:365-404).  So the cube [-R,R]^3
// is cut into NG^3 cells and every cell stores the list of ALL target points that can be the
// nearest neighbour of ANY query in that cell:
//       cand(C) = { p : |c - p| <= d_c + 2*rho },  c = centre, rho = half diagonal,
//                                                   d_c = distance from c to its nearest target
// (for q in C with NN p*:  |q-p*| <= |q-p_c| <= d_c + rho  =>  |c-p*| <= d_c + 2 rho).
// On top of that a candidate is dropped when the centre's nearest target p_c is strictly closer than
// it for EVERY point of the cell (a linear function of the position, so one evaluation at the worst
// corner decides; see cg_level_kernel).  For cells far from the surface this removes almost all of
// the spherical-shell candidates: finest-level lists hold ~2-3 entries.
// A query then evaluates only its cell's list with the exact FLANN distance expression, so the
// result (minimum d2, lowest original index among equal minima) is bit-identical to a full
// search.  Lists are built coarse-to-fine (4^3 -> 64^3): cand(child) is a subset of cand(parent).
// All thresholds carry a 1e-4 relative slack, far above fp32 rounding of the bound arithmetic.
#pragma once
#include "kss_device.cuh"

namespace kss {

constexpr int CG_LEVELS = 4;                  // 4, 8, 16, 32 cells per axis
constexpr int CG_NG0 = 4;
constexpr int CG_NG = CG_NG0 << (CG_LEVELS - 1);
constexpr int CG_CNT_BITS = 11;               // header = (offset/4 << 11) | count ; count 2047 = "search everything"
constexpr unsigned CG_CNT_MASK = (1u << CG_CNT_BITS) - 1u;
constexpr size_t CG_ARENA = (size_t)3 << 19;  // 1.5 M u16 entries per pair (all levels); overflow -> tile search
// Sparse fifth level: a finest-level cell whose list has >= CG_REFINE_MIN entries is cut into its 8 octants, each with
// its own (shorter) list.  Its header then carries count CG_REFINED and the offset of 8 child headers in the arena.
constexpr unsigned CG_REFINED = CG_CNT_MASK - 1u;
constexpr int CG_WL_CAP = 12288;              // refined cells per pair at most (the rest simply keep their list)
#ifndef CG_REFINE_MIN
#define CG_REFINE_MIN 7
#endif
__host__ __device__ constexpr int cg_ng(int l) { return CG_NG0 << l; }
__host__ __device__ constexpr size_t cg_hdr_base(int l) {     // start of level l in the per-pair header array
    size_t b = 0;
    for (int i = 0; i < l; ++i) b += (size_t)cg_ng(i) * cg_ng(i) * cg_ng(i);
    return b;
}
constexpr size_t CG_HDR_TOTAL = cg_hdr_base(CG_LEVELS);

struct CgView {
    const unsigned* hdr;            // finest level headers [NG^3], linear (ix + NG*(iy + NG*iz))
    const unsigned short* list;     // this pair's arena
    float lox, loy, loz, inv_h;     // cell = floor((q - lo) * inv_h) per axis
    int ok;                         // 0: arena overflowed, use the tile search
};

// per-pair cube: geom[0..2] = centre, [3] = half size, [4] = ball radius around the centre that
// bounds all queries (+inf: whole cube), [5..7] unused
__device__ __forceinline__ CgView cg_view(const float* __restrict__ geom, const unsigned* __restrict__ hdr_all,
                                          const unsigned short* __restrict__ arena_all, const int* __restrict__ ok, int p);


// exact 1-NN through the candidate grid; tgt = Morton-ordered float4 {x,y,z,bits(orig)} in shared memory.
// IDX: key = (d2 bits << 32) | original index (lowest original index among equal d2), else d2 bits << 32.
template <bool IDX>
__device__ __forceinline__ unsigned long long cg_query(const CgView& g, const float4* __restrict__ tgt, int n_t,
                                                       float qx, float qy, float qz) {
    const float fx = (qx - g.lox) * g.inv_h, fy = (qy - g.loy) * g.inv_h, fz = (qz - g.loz) * g.inv_h;
    const float ngf = (float)CG_NG;
    unsigned cnt = CG_CNT_MASK, off = 0;
    if (fx >= 0.0f && fy >= 0.0f && fz >= 0.0f && fx < ngf && fy < ngf && fz < ngf) {
        const int cell = (int)fx + CG_NG * ((int)fy + CG_NG * (int)fz);
        unsigned h = g.hdr[cell];
        cnt = h & CG_CNT_MASK;
        off = (h >> CG_CNT_BITS) << 2;
        if (cnt == CG_REFINED) {
            const int oct = ((fx - (float)(int)fx) >= 0.5f ? 1 : 0) | ((fy - (float)(int)fy) >= 0.5f ? 2 : 0) |
                            ((fz - (float)(int)fz) >= 0.5f ? 4 : 0);
            h = __ldg(reinterpret_cast<const unsigned*>(g.list + off) + oct);
            cnt = h & CG_CNT_MASK;
            off = (h >> CG_CNT_BITS) << 2;
        }
    }
    float best = __int_as_float(0x7f800000);
    unsigned bi = 0u;
    unsigned long long bestkey = 0xffffffffffffffffull;
    if (cnt == 0u || cnt == CG_CNT_MASK) {
        // outside the cube, an empty header or an over-long list: plain scan of every target (rare)
        for (int j = 0; j < n_t; ++j) {
            const float4 p = tgt[j];
            const float d = d2_rn(qx, qy, qz, p.x, p.y, p.z);
            if (IDX) {
                const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_uint(p.w);
                bestkey = key < bestkey ? key : bestkey;
            } else best = fminf(best, d);
        }
    } else {
        const uint2* lp = reinterpret_cast<const uint2*>(g.list + off);      // lists are padded to x4 entries
        const unsigned n4 = (cnt + 3u) >> 2;
        for (unsigned j = 0; j < n4; ++j) {
            const uint2 w = __ldg(lp + j);
            const unsigned id[4] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float4 p = tgt[id[k]];
                const float d = d2_rn(qx, qy, qz, p.x, p.y, p.z);
                // lists are ascending in original index (cg_level_kernel): the first of equal distances wins
                if (IDX) { if (d < best) { best = d; bi = id[k]; } }
                else best = fminf(best, d);
            }
        }
        if (IDX) bestkey = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned)__float_as_uint(tgt[bi].w);
    }
    if (!IDX) bestkey = (unsigned long long)__float_as_uint(best) << 32;
    return bestkey;
}

// U independent queries of one thread at once: the three dependent global reads of a query (cell header, octant
// header of a refined cell, first list word) are issued for all U queries before any is consumed, so their
// latencies overlap instead of adding up.  Same arithmetic and result as cg_query for every query.
template <int U, bool IDX>
__device__ __forceinline__ void cg_query_batch(const CgView& g, const float4* __restrict__ tgt, int n_t,
                                               const float (&qx)[U], const float (&qy)[U], const float (&qz)[U],
                                               unsigned long long (&out)[U]) {
    const float ngf = (float)CG_NG;
    unsigned h[U]; int oct[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const float fx = (qx[u] - g.lox) * g.inv_h, fy = (qy[u] - g.loy) * g.inv_h, fz = (qz[u] - g.loz) * g.inv_h;
        h[u] = CG_CNT_MASK;                                   // outside the cube: scan everything
        oct[u] = ((fx - (float)(int)fx) >= 0.5f ? 1 : 0) | ((fy - (float)(int)fy) >= 0.5f ? 2 : 0) |
                 ((fz - (float)(int)fz) >= 0.5f ? 4 : 0);
        if (fx >= 0.0f && fy >= 0.0f && fz >= 0.0f && fx < ngf && fy < ngf && fz < ngf)
            h[u] = g.hdr[(int)fx + CG_NG * ((int)fy + CG_NG * (int)fz)];
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if ((h[u] & CG_CNT_MASK) == CG_REFINED)
            h[u] = __ldg(reinterpret_cast<const unsigned*>(g.list + ((h[u] >> CG_CNT_BITS) << 2)) + oct[u]);
    uint2 w0[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const unsigned cnt = h[u] & CG_CNT_MASK;
        w0[u] = make_uint2(0u, 0u);
        if (cnt != 0u && cnt != CG_CNT_MASK) w0[u] = __ldg(reinterpret_cast<const uint2*>(g.list + ((h[u] >> CG_CNT_BITS) << 2)));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const unsigned cnt = h[u] & CG_CNT_MASK, off = (h[u] >> CG_CNT_BITS) << 2;
        float best = __int_as_float(0x7f800000);
        unsigned bi = 0u;
        unsigned long long bestkey = 0xffffffffffffffffull;
        if (cnt == 0u || cnt == CG_CNT_MASK) {
            for (int j = 0; j < n_t; ++j) {
                const float4 p = tgt[j];
                const float d = d2_rn(qx[u], qy[u], qz[u], p.x, p.y, p.z);
                if (IDX) {
                    const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_uint(p.w);
                    bestkey = key < bestkey ? key : bestkey;
                } else best = fminf(best, d);
            }
        } else {
            const uint2* lp = reinterpret_cast<const uint2*>(g.list + off);
            const unsigned n4 = (cnt + 3u) >> 2;
            uint2 w = w0[u];
            for (unsigned j = 0; j < n4; ++j) {
                const unsigned id[4] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16};
                if (j + 1 < n4) w = __ldg(lp + j + 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 p = tgt[id[k]];
                    const float d = d2_rn(qx[u], qy[u], qz[u], p.x, p.y, p.z);
                    if (IDX) { if (d < best) { best = d; bi = id[k]; } }      // ascending original index: first wins
                    else best = fminf(best, d);
                }
            }
            if (IDX) bestkey = ((unsigned long long)__float_as_uint(best) << 32) | (unsigned)__float_as_uint(tgt[bi].w);
        }
        if (!IDX) bestkey = (unsigned long long)__float_as_uint(best) << 32;
        out[u] = bestkey;
    }
}

__device__ __forceinline__ CgView cg_view(const float* __restrict__ geom, const unsigned* __restrict__ hdr_all,
                                          const unsigned short* __restrict__ arena_all, const int* __restrict__ ok, int p) {
    CgView g;
    const float* gm = geom + (size_t)p * 8;
    g.hdr = hdr_all + (size_t)p * CG_HDR_TOTAL + cg_hdr_base(CG_LEVELS - 1);
    g.list = arena_all + (size_t)p * CG_ARENA;
    g.lox = gm[0] - gm[3]; g.loy = gm[1] - gm[3]; g.loz = gm[2] - gm[3];
    g.inv_h = (float)CG_NG / (2.0f * gm[3]);
    g.ok = ok[p];
    return g;
}

}  // namespace kss
