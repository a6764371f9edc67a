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
constexpr size_t CG_ARENA = (size_t)3 << 19;  // 1.5 M u16 entries per pair (all levels); overflow -> tile search
// 64-bit cell header.  The sweep is bound by the sectors its gathers pull through L2 and by the instructions its
// divergent list walks cost, so a list of up to five candidates (11-bit positions in the Morton-ordered target array)
// lives INSIDE the header: one 8-byte read, five branch-free evaluations.
//   tag = bits 63..60:  0 empty (never built: scan everything)   1..5 inline list, id_k = bits [11k+10 : 11k]
//                      13 over-long list: scan everything        14 external list: count = bits 59..32,
//                      15 refined: bits 31..0 = offset/4 of        offset/4 (u16 arena entries) = bits 31..0,
//                         8 child headers (one per octant)          padded to a multiple of 4 entries
typedef unsigned long long cg_hdr_t;
constexpr unsigned CG_TAG_HUGE = 13u, CG_TAG_EXT = 14u, CG_TAG_REFINED = 15u;
constexpr int CG_INLINE_MAX = 5;
constexpr int CG_EXT_MAX = 2046;              // longer lists are stored as "scan everything"
// Two sparse levels below the dense 32^3 grid: a cell whose list has >= CG_REFINE_MIN entries is cut into its 8 octants
// (64^3 resolution), and an octant that still has that many once more (128^3).  External lists are then rare: a lane
// that walks one stalls its whole warp.
constexpr int CG_WL_CAP = 12288;              // refined 32^3 cells per pair at most (the rest simply keep their list)
constexpr int CG_WL2_CAP = 4096;              // refined octants per pair at most
#ifndef CG_REFINE_MIN
#define CG_REFINE_MIN 6
#endif
__host__ __device__ constexpr int cg_ng(int l) { return CG_NG0 << l; }
__host__ __device__ constexpr size_t cg_hdr_base(int l) {     // start of level l in the per-pair header array
    size_t b = 0;
    for (int i = 0; i < l; ++i) b += (size_t)cg_ng(i) * cg_ng(i) * cg_ng(i);
    return b;
}
constexpr size_t CG_HDR_TOTAL = cg_hdr_base(CG_LEVELS);

__host__ __device__ __forceinline__ unsigned cg_tag(cg_hdr_t h) { return (unsigned)(h >> 60); }
__host__ __device__ __forceinline__ unsigned cg_inline_id(cg_hdr_t h, int k) { return (unsigned)(h >> (11 * k)) & 0x7ffu; }
__host__ __device__ __forceinline__ unsigned cg_ext_count(cg_hdr_t h) { return (unsigned)(h >> 32) & 0x0fffffffu; }
__host__ __device__ __forceinline__ unsigned cg_offset(cg_hdr_t h) { return (unsigned)(h & 0xffffffffu) << 2; }   // u16 entries

struct CgView {
    const cg_hdr_t* hdr;            // finest level headers [NG^3], linear (ix + NG*(iy + NG*iz))
    const unsigned short* list;     // this pair's arena
    float lox, loy, loz, inv_h;     // cell = floor((q - lo) * inv_h) per axis
    int ok;                         // 0: arena overflowed, use the tile search
};

// per-pair cube: geom[0..2] = centre, [3] = half size, [4] = ball radius around the centre that
// bounds all queries (+inf: whole cube), [5..7] unused
__device__ __forceinline__ CgView cg_view(const float* __restrict__ geom, const cg_hdr_t* __restrict__ hdr_all,
                                          const unsigned short* __restrict__ arena_all, const int* __restrict__ ok, int p);

// one candidate: exact FLANN distance; lists are ascending in original index (cg_level_kernel), so the first of equal
// distances wins and `bi` (a position in tgt) identifies it
template <bool IDX>
__device__ __forceinline__ void cg_eval(const float4* __restrict__ tgt, unsigned id, float qx, float qy, float qz, float& best, unsigned& bi) {
    const float4 p = tgt[id];
    const float d = d2_rn(qx, qy, qz, p.x, p.y, p.z);
    if (IDX) { if (d < best) { best = d; bi = id; } }
    else best = fminf(best, d);
}

// the list of one (already octant-resolved) header
template <bool IDX>
__device__ __forceinline__ unsigned long long cg_scan(const CgView& g, const float4* __restrict__ tgt, int n_t, cg_hdr_t h,
                                                      float qx, float qy, float qz) {
    const unsigned tag = cg_tag(h);
    float best = __int_as_float(0x7f800000);
    unsigned bi = 0u;
    if (tag >= 1u && tag <= (unsigned)CG_INLINE_MAX) {
        // branch-free: slots past the count re-evaluate entry 0 (same distance, never strictly smaller)
        const unsigned id0 = cg_inline_id(h, 0);
        cg_eval<IDX>(tgt, id0, qx, qy, qz, best, bi);
#pragma unroll
        for (int k = 1; k < CG_INLINE_MAX; ++k)                          // constant shifts, one select per slot
            cg_eval<IDX>(tgt, (unsigned)k < tag ? cg_inline_id(h, k) : id0, qx, qy, qz, best, bi);
    } else if (tag == CG_TAG_EXT) {
        const uint2* lp = reinterpret_cast<const uint2*>(g.list + cg_offset(h));
        const unsigned n4 = (cg_ext_count(h) + 3u) >> 2;
        uint2 w = __ldg(lp);
        for (unsigned j = 0; j < n4; ++j) {
            const unsigned id[4] = {w.x & 0xffffu, w.x >> 16, w.y & 0xffffu, w.y >> 16};
            if (j + 1 < n4) w = __ldg(lp + j + 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) cg_eval<IDX>(tgt, id[k], qx, qy, qz, best, bi);
        }
    } else {
        // outside the cube, a cell that was never built or an over-long list: plain scan of every target (rare);
        // Morton order, so equal distances are resolved explicitly by the original index
        unsigned long long bestkey = 0xffffffffffffffffull;
        for (int j = 0; j < n_t; ++j) {
            const float4 p = tgt[j];
            const float d = d2_rn(qx, qy, qz, p.x, p.y, p.z);
            if (IDX) {
                const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_uint(p.w);
                bestkey = key < bestkey ? key : bestkey;
            } else best = fminf(best, d);
        }
        if (IDX) return bestkey;
    }
    if (IDX) return ((unsigned long long)__float_as_uint(best) << 32) | (unsigned)__float_as_uint(tgt[bi].w);
    return (unsigned long long)__float_as_uint(best) << 32;
}

// exact 1-NN through the candidate grid; tgt = Morton-ordered float4 {x,y,z,bits(orig)} in shared memory.
// IDX: key = (d2 bits << 32) | original index (lowest original index among equal d2), else d2 bits << 32.
// U independent queries of one thread at once: the dependent global reads of a query (cell header, then up to two
// octant headers of refined cells) are issued for all U queries before any is consumed.
template <int U, bool IDX>
__device__ __forceinline__ void cg_query_batch(const CgView& g, const float4* __restrict__ tgt, int n_t,
                                               const float (&qx)[U], const float (&qy)[U], const float (&qz)[U],
                                               unsigned long long (&out)[U]) {
    const float ngf = (float)CG_NG;
    cg_hdr_t h[U]; int oct1[U], oct2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const float fx = (qx[u] - g.lox) * g.inv_h, fy = (qy[u] - g.loy) * g.inv_h, fz = (qz[u] - g.loz) * g.inv_h;
        h[u] = 0ull;                                          // outside the cube: scan everything
        // quarter-cell coordinates (the multiplication by 4 is exact): bit 1 = octant, bit 0 = octant of the octant
        const int ax = (int)(fx * 4.0f), ay = (int)(fy * 4.0f), az = (int)(fz * 4.0f);
        oct1[u] = ((ax >> 1) & 1) | (ay & 2) | ((az & 2) << 1);
        oct2[u] = (ax & 1) | ((ay & 1) << 1) | ((az & 1) << 2);
        if (fx >= 0.0f && fy >= 0.0f && fz >= 0.0f && fx < ngf && fy < ngf && fz < ngf)
            h[u] = g.hdr[(ax >> 2) + CG_NG * ((ay >> 2) + CG_NG * (az >> 2))];
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (cg_tag(h[u]) == CG_TAG_REFINED) h[u] = __ldg(reinterpret_cast<const cg_hdr_t*>(g.list + cg_offset(h[u])) + oct1[u]);
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (cg_tag(h[u]) == CG_TAG_REFINED) h[u] = __ldg(reinterpret_cast<const cg_hdr_t*>(g.list + cg_offset(h[u])) + oct2[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) out[u] = cg_scan<IDX>(g, tgt, n_t, h[u], qx[u], qy[u], qz[u]);
}

template <bool IDX>
__device__ __forceinline__ unsigned long long cg_query(const CgView& g, const float4* __restrict__ tgt, int n_t,
                                                       float qx, float qy, float qz) {
    const float x[1] = {qx}, y[1] = {qy}, z[1] = {qz};
    unsigned long long out[1];
    cg_query_batch<1, IDX>(g, tgt, n_t, x, y, z, out);
    return out[0];
}

__device__ __forceinline__ CgView cg_view(const float* __restrict__ geom, const cg_hdr_t* __restrict__ hdr_all,
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
