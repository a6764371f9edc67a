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
#include <cstdlib>
#include <cstring>
#include <ctime>
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

// ------------------------------------------------------------------ large candidate grid (LCG)
// The idea of kss_cg.cuh carried to clouds of any size: a DENSE grid over the target's bounding box
// (power-of-two cells per axis, anisotropic, ~32 cells per target point, <= 2^26) whose every cell lists
// ALL targets that can be the nearest neighbour of ANY query inside it (sphere rule + dominance rule,
// same rigorous margins).  Built coarse-to-fine from a single root cell; a cell is handled by a group
// of GS threads (256 / 32 / 1) chosen per level from the mean parent-list length.  Queries then cost
// one header read plus a handful of candidate distances per THREAD (no warp cooperation); queries
// outside the box fall back to the box pyramid for their whole warp.
struct LcgGeom {
    float lo[3], h[3], inv_h[3];     // finest level: cell = floor((q - lo) * inv_h)
    int bits[3];                     // finest level: 2^bits cells per axis
    int levels;                      // level 0 = root ... levels-1 = finest
};
struct LcgView {
    const unsigned long long* hdr;   // finest level headers: far flag | count << 40 | offset
    const float4* arena4;            // finest-level lists: the candidate points themselves, .w = bits(original index)
    const unsigned* arena;           // u32 lists of unrefined far cells inherited from coarser levels
    const float4* tgt;               // targets by original index (for the u32 lists)
    LcgGeom g;
    int ok;
};
__host__ __device__ inline int lcg_bits(const LcgGeom& g, int level, int a) {
    const int L = g.levels - 1;
    const int b = g.bits[a] - (L - level);
    return b > 0 ? b : 0;
}

struct LcgLevelArgs {
    int level;
    int bits[3], pbits[3];           // this level / parent level bits per axis
    float lo[3], h[3];               // this level: cell sizes
    const float4* tgt; int n_t;
    const unsigned long long* hdr_prev; unsigned long long* hdr_cur;
    unsigned* arena; unsigned long long* cursor; unsigned long long cap;   // scratch region of this level: cursor[0] < cap
    unsigned long long cap_persist;                                       // persistent region: cursor[1] < cap_persist
    int from_all;                                                         // first built level: the parent list is every target
    const unsigned* worklist; unsigned nwork;                             // GS=1 pass 2: cells to build (compacted)
    unsigned* worklist_out; unsigned* worklist_long; unsigned* nwork_out; // GS=1 pass 1: cells that do not inherit (short / long parent lists)
    int persist_all;                                                      // finest level: everything is persistent ...
    float4* arena4;                                                       // ... and stored as float4 {x,y,z,bits(orig)} here (offsets in float4 units)
    int* ok;
};

// header: bit 63 = "far" (centre more than 3 half-diagonals from every target; a far cell with a short list is
// not refined, its descendants inherit the header), bit 62 = list is u32 indices (else float4 points),
// bits 61:36 = count, bits 35:0 = offset (multiple of 4) into the respective arena.
// Lists are padded to a multiple of 4 entries with a repeated valid candidate.
constexpr unsigned long long LCG_FAR = 1ull << 63;
constexpr unsigned long long LCG_IDX = 1ull << 62;     // list holds u32 indices in `arena` (all intermediate levels)
__host__ __device__ inline int lcg_count(unsigned long long h) { return (int)((h >> 36) & 0x3ffffffull); }
__host__ __device__ inline unsigned long long lcg_offset(unsigned long long h) { return h & 0xfffffffffull; }
__host__ __device__ inline unsigned long long lcg_make(bool far, bool idx, unsigned long long cnt, unsigned long long off) {
    return (far ? LCG_FAR : 0ull) | (idx ? LCG_IDX : 0ull) | (cnt << 36) | off;
}

template <int GS>
__global__ void __launch_bounds__(256)
lcg_level_kernel(LcgLevelArgs a) {
    __shared__ unsigned long long s_key[8];
    __shared__ unsigned s_cnt[8];
    __shared__ unsigned long long s_base;
    const int lane = threadIdx.x & 31;
    const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long cell = gtid / GS;
    const int r = (int)(gtid % GS);
    const long long ncells = a.worklist ? (long long)a.nwork : 1ll << (a.bits[0] + a.bits[1] + a.bits[2]);
    if (GS == 1) { if (gtid - lane >= ncells) return; }
    else if (cell >= ncells) return;                                  // group-uniform (GS is 32 or 256 = block)
    const bool live = cell < ncells;
    if (a.worklist) cell = live ? (long long)a.worklist[cell] : 0;    // compacted pass: the cells that need building
    const int ix = (int)(cell & ((1ll << a.bits[0]) - 1));
    const int iy = (int)((cell >> a.bits[0]) & ((1ll << a.bits[1]) - 1));
    const int iz = (int)(cell >> (a.bits[0] + a.bits[1]));
    int m_p = 0;
    unsigned long long ph = 0ull;
    const unsigned* plist = nullptr;
    if (live) {
        if (a.from_all) m_p = a.n_t;
        else {
            const long long parent = (long long)(ix >> (a.bits[0] - a.pbits[0])) +
                                     ((long long)(iy >> (a.bits[1] - a.pbits[1])) << a.pbits[0]) +
                                     ((long long)(iz >> (a.bits[2] - a.pbits[2])) << (a.pbits[0] + a.pbits[1]));
            ph = a.hdr_prev[parent];
            m_p = lcg_count(ph);
            plist = a.arena + lcg_offset(ph);
        }
    }
    // far parents whose list is already short are not refined: the child inherits the header (the parent's list
    // is a valid superset for every query inside the child).  This is empty space away from the surface.
    // (finest level only: level L-1 stays in its scratch region, which nothing overwrites afterwards)
    const bool inherit = live && a.persist_all && !a.from_all && (ph & LCG_FAR) && m_p <= 128;
    if (GS == 1 && a.worklist_out) {
        // pass 1 of the finest level: write inherited headers, compact everything else into the work list so
        // that pass 2 runs with dense warps (almost all cells of the finest level inherit)
        if (inherit) a.hdr_cur[cell] = ph;
        const bool want = live && !inherit;
        const bool lng = want && m_p > 48;                            // long parent lists get a warp per cell
        const unsigned ns = __ballot_sync(KSS_FULL, want && !lng), nl = __ballot_sync(KSS_FULL, lng);
        unsigned bs = 0, bl = 0;
        if (lane == 0) { if (ns) bs = atomicAdd(a.nwork_out, (unsigned)__popc(ns)); if (nl) bl = atomicAdd(a.nwork_out + 1, (unsigned)__popc(nl)); }
        bs = __shfl_sync(KSS_FULL, bs, 0); bl = __shfl_sync(KSS_FULL, bl, 0);
        const unsigned below = (1u << lane) - 1u;
        if (want && !lng) a.worklist_out[bs + __popc(ns & below)] = (unsigned)cell;
        if (lng) a.worklist_long[bl + __popc(nl & below)] = (unsigned)cell;
        return;
    }
    if (inherit) m_p = 0;
    const float cx = a.lo[0] + ((float)ix + 0.5f) * a.h[0], cy = a.lo[1] + ((float)iy + 0.5f) * a.h[1],
                cz = a.lo[2] + ((float)iz + 0.5f) * a.h[2];
    const float rho = 0.5f * sqrtf(a.h[0] * a.h[0] + a.h[1] * a.h[1] + a.h[2] * a.h[2]) * 1.002f;
    auto cand = [&](int j) -> unsigned { return plist ? plist[j] : (unsigned)j; };

    // ---- pass A: nearest target of the centre (GS = 1: the four nearest, used as dominance competitors)
    unsigned long long key = 0xffffffffffffffffull, k1 = ~0ull, k2 = ~0ull, k3 = ~0ull;
    for (int j = r; j < m_p; j += GS) {
        const unsigned id = cand(j);
        const float4 q = __ldg(a.tgt + id);
        const unsigned long long kk = ((unsigned long long)__float_as_uint(d2_rn(cx, cy, cz, q.x, q.y, q.z)) << 32) | id;
        if (GS <= 32) {
            if (kk < k3) {
                if (kk < key) { k3 = k2; k2 = k1; k1 = key; key = kk; }
                else if (kk < k1) { k3 = k2; k2 = k1; k1 = kk; }
                else if (kk < k2) { k3 = k2; k2 = kk; }
                else k3 = kk;
            }
        } else key = kk < key ? kk : key;
    }
    if (GS == 32) {
        // merge the lanes' sorted top-4 lists into the warp's top-4: four rounds of "global minimum, pop it at its owner"
        unsigned long long g[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            unsigned long long m = key;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) { const unsigned long long o = __shfl_xor_sync(KSS_FULL, m, off); m = o < m ? o : m; }
            g[t] = m;
            if (key == m && m != ~0ull) { key = k1; k1 = k2; k2 = k3; k3 = ~0ull; }   // keys are unique (they carry the index)
        }
        key = g[0]; k1 = g[1]; k2 = g[2]; k3 = g[3];
    }
    if (GS == 256) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) { const unsigned long long o = __shfl_xor_sync(KSS_FULL, key, off); key = o < key ? o : key; }
        if (lane == 0) s_key[threadIdx.x >> 5] = key;
        __syncthreads();
        key = s_key[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) key = s_key[w] < key ? s_key[w] : key;
    }
    const float mn = __uint_as_float((unsigned)(key >> 32));
    const unsigned amin = m_p > 0 ? (unsigned)(key & 0xffffffffu) : 0u;
    const float thr = (sqrtf(mn) + 2.0f * rho) * 1.0001f, thr2 = thr * thr;
    const float sx = a.h[0] * 1.002f * 1.0001f, sy = a.h[1] * 1.002f * 1.0001f, sz = a.h[2] * 1.002f * 1.0001f;
    const float rr4 = 4.0f * rho * rho;
    constexpr int NC = GS <= 32 ? 4 : 1;
    const unsigned long long ck[4] = {key, k1, k2, k3};
    float4 cp[NC]; float cd[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        const bool have = m_p > 0 && ck[c] != ~0ull;
        cp[c] = __ldg(a.tgt + (have ? (unsigned)(ck[c] & 0xffffffffu) : 0u));
        cd[c] = have ? __uint_as_float((unsigned)(ck[c] >> 32)) : __int_as_float(0x7f800000);   // +inf: never dominates
    }
    auto keep_test = [&](const float4& q, float d) -> bool {       // sphere rule and dominance rule (see kss_cg.cuh)
        if (!(d <= thr2)) return false;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const float s = sx * fabsf(q.x - cp[c].x) + sy * fabsf(q.y - cp[c].y) + sz * fabsf(q.z - cp[c].z);
            if ((d - cd[c]) - s > 1e-5f * (d + cd[c] + rr4)) return false;
        }
        return true;
    };
    // ---- pass B: count
    unsigned long long mask = 0ull;
    unsigned k = 0;
    for (int j = r, jj = 0; j < m_p; j += GS, ++jj) {
        const float4 q = __ldg(a.tgt + cand(j));
        if (keep_test(q, d2_rn(cx, cy, cz, q.x, q.y, q.z))) { if (jj < 64) mask |= 1ull << jj; ++k; }
    }
    // ---- allocate (lists padded to x4).  Lists of far cells are inherited by descendants of any depth, so they
    //      go to the persistent region (cursor[1]); all other lists of an intermediate level go to this level's
    //      ping-pong scratch region (cursor[0]).  One atomicAdd per warp and region.
    const bool far = m_p > 0 && sqrtf(mn) > 3.0f * rho;
    unsigned long long base = 0;
    unsigned pre = 0, ktot = k;
    bool persist;
    if (GS == 1) {
        persist = a.persist_all != 0;
        const unsigned mine = (k + 3u) & ~3u;
        unsigned inc0 = persist ? 0u : mine, inc1 = persist ? mine : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y0 = __shfl_up_sync(KSS_FULL, inc0, o), y1 = __shfl_up_sync(KSS_FULL, inc1, o);
            if (lane >= o) { inc0 += y0; inc1 += y1; }
        }
        const unsigned t0 = __shfl_sync(KSS_FULL, inc0, 31), t1 = __shfl_sync(KSS_FULL, inc1, 31);
        unsigned long long b0 = 0, b1 = 0;
        if (lane == 0) { if (t0) b0 = atomicAdd(a.cursor, (unsigned long long)t0); if (t1) b1 = atomicAdd(a.cursor + 1, (unsigned long long)t1); }
        b0 = __shfl_sync(KSS_FULL, b0, 0); b1 = __shfl_sync(KSS_FULL, b1, 0);
        if (b0 + t0 > a.cap || b1 + t1 > a.cap_persist) { atomicExch(a.ok, 0); if (live) a.hdr_cur[cell] = 0ull; return; }
        base = persist ? b1 : b0;
        pre = (persist ? inc1 : inc0) - mine;
    } else {
        unsigned incl = k;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(KSS_FULL, incl, o); if (lane >= o) incl += y; }
        unsigned tot = __shfl_sync(KSS_FULL, incl, 31);
        pre = incl - k;
        if (GS == 256) {
            if (lane == 31) s_cnt[threadIdx.x >> 5] = incl;
            __syncthreads();
            unsigned before = 0; tot = 0;
            for (int w = 0; w < 8; ++w) { if (w < (int)(threadIdx.x >> 5)) before += s_cnt[w]; tot += s_cnt[w]; }
            pre += before;
            persist = a.persist_all != 0;
            if (threadIdx.x == 0) s_base = tot ? atomicAdd(a.cursor + (persist ? 1 : 0), (unsigned long long)((tot + 3u) & ~3u)) : 0ull;
            __syncthreads();
            base = s_base;
        } else {
            persist = a.persist_all != 0;
            if (lane == 0 && tot) base = atomicAdd(a.cursor + (persist ? 1 : 0), (unsigned long long)((tot + 3u) & ~3u));
            base = __shfl_sync(KSS_FULL, base, 0);
        }
        ktot = tot;
        if (base + ((tot + 3u) & ~3u) > (persist ? a.cap_persist : a.cap)) { atomicExch(a.ok, 0); if (live && r == 0) a.hdr_cur[cell] = 0ull; return; }
    }
    if (!live) return;
    if (inherit) { a.hdr_cur[cell] = ph; return; }
    if (r == 0) a.hdr_cur[cell] = lcg_make(far, a.arena4 == nullptr, ktot, GS == 1 ? base + pre : base);
    // ---- pass C: write (intermediate levels: u32 indices; finest level: the points themselves, so that a
    //      query's candidates are contiguous 16-byte records instead of one 32-byte sector per gather)
    unsigned* out = a.arena + base + pre;
    float4* out4 = a.arena4 ? a.arena4 + base + pre : nullptr;
    unsigned w = 0;
    auto emit = [&](unsigned id) {
        if (out4) { float4 q = __ldg(a.tgt + id); q.w = __uint_as_float(id); out4[w] = q; } else out[w] = id;
        ++w;
    };
    if ((m_p + GS - 1) / GS <= 64) {
        while (mask) {
            const int jj = __ffsll((long long)mask) - 1;
            mask &= mask - 1ull;
            emit(cand(r + GS * jj));
        }
    } else {
        for (int j = r; j < m_p; j += GS) {
            const unsigned id = cand(j);
            const float4 q = __ldg(a.tgt + id);
            if (keep_test(q, d2_rn(cx, cy, cz, q.x, q.y, q.z))) emit(id);
        }
    }
    // padding to x4 (the centre's nearest target is always a member of the list)
    if (GS == 1) { for (const unsigned e = (k + 3u) & ~3u; w < e;) emit(amin); }
    else if (r == 0) {
        for (unsigned t = ktot; t < ((ktot + 3u) & ~3u); ++t) {
            if (a.arena4) { float4 q = __ldg(a.tgt + amin); q.w = __uint_as_float(amin); a.arena4[base + t] = q; }
            else a.arena[base + t] = amin;
        }
    }
}

// exact 1-NN through the large candidate grid (per thread); returns 0 on success, 1 if the query is outside the
// box, 2 for an empty header, 3 for an unrefined far cell with a long list (the caller then uses the box pyramid)
template <bool IDX>
__device__ __forceinline__ int lcg_query(const LcgView& v, float qx, float qy, float qz,
                                          unsigned long long& keyout) {
    const float fx = (qx - v.g.lo[0]) * v.g.inv_h[0], fy = (qy - v.g.lo[1]) * v.g.inv_h[1], fz = (qz - v.g.lo[2]) * v.g.inv_h[2];
    const float nx = (float)(1 << v.g.bits[0]), ny = (float)(1 << v.g.bits[1]), nz = (float)(1 << v.g.bits[2]);
    if (!(fx >= 0.0f && fy >= 0.0f && fz >= 0.0f && fx < nx && fy < ny && fz < nz)) return 1;
    const size_t cell = (size_t)(int)fx + ((size_t)(int)fy << v.g.bits[0]) + ((size_t)(int)fz << (v.g.bits[0] + v.g.bits[1]));
    const unsigned long long h = __ldg(v.hdr + cell);
    const int cnt = lcg_count(h);
    if (cnt == 0) return 2;
    if (cnt > 1024) return 3;
    float best = __int_as_float(0x7f800000);
    unsigned long long bestkey = 0xffffffffffffffffull;
    if (h & LCG_IDX) {           // inherited list of an unrefined far cell: u32 indices (never on the steady-state path)
        const unsigned* ip = v.arena + lcg_offset(h);
        for (int j = 0; j < cnt; ++j) {
            const unsigned id = __ldg(ip + j);
            const float4 p = __ldg(v.tgt + id);
            const float d = d2_rn(qx, qy, qz, p.x, p.y, p.z);
            if (IDX) { const unsigned long long k0 = ((unsigned long long)__float_as_uint(d) << 32) | id; bestkey = k0 < bestkey ? k0 : bestkey; }
            else best = fminf(best, d);
        }
        keyout = IDX ? bestkey : ((unsigned long long)__float_as_uint(best) << 32);
        return 0;
    }
    const float4* lp = v.arena4 + lcg_offset(h);
    const int n4 = (cnt + 3) >> 2;
    for (int j = 0; j < n4; ++j) {
        const float4 p0 = __ldg(lp + 4 * j), p1 = __ldg(lp + 4 * j + 1), p2 = __ldg(lp + 4 * j + 2), p3 = __ldg(lp + 4 * j + 3);
        const float d0 = d2_rn(qx, qy, qz, p0.x, p0.y, p0.z), d1 = d2_rn(qx, qy, qz, p1.x, p1.y, p1.z);
        const float d2 = d2_rn(qx, qy, qz, p2.x, p2.y, p2.z), d3 = d2_rn(qx, qy, qz, p3.x, p3.y, p3.z);
        if (IDX) {                                                       // .w = original index
            const unsigned long long k0 = ((unsigned long long)__float_as_uint(d0) << 32) | __float_as_uint(p0.w);
            const unsigned long long k1 = ((unsigned long long)__float_as_uint(d1) << 32) | __float_as_uint(p1.w);
            const unsigned long long k2 = ((unsigned long long)__float_as_uint(d2) << 32) | __float_as_uint(p2.w);
            const unsigned long long k3 = ((unsigned long long)__float_as_uint(d3) << 32) | __float_as_uint(p3.w);
            const unsigned long long m01 = k0 < k1 ? k0 : k1, m23 = k2 < k3 ? k2 : k3;
            const unsigned long long m = m01 < m23 ? m01 : m23;
            bestkey = m < bestkey ? m : bestkey;
        } else best = fminf(best, fminf(fminf(d0, d1), fminf(d2, d3)));
    }
    keyout = IDX ? bestkey : ((unsigned long long)__float_as_uint(best) << 32);
    return 0;
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
    unsigned miss[5];            // queries not served by the candidate grid, by reason (diagnostics)
};

// NN kernel.  MODE 0: plain queries from q4 (float4 by original index) -> idx/d2 by original index
//             MODE 1: ICP iteration: lazily apply st->Tk to cur (in place), correspondences with rejection
//             MODE 2: fitness pass: query = st->fin * inp (original input), d2 only
template <int MODE>
__global__ void __launch_bounds__(LG_WARPS * 32)
lg_nn_kernel(Pyramid py, LcgView lcg, const float4* __restrict__ t_orig, const int* __restrict__ perm, int n_q,
             float4* __restrict__ cur, const float4* __restrict__ inp, int* __restrict__ idx, float* __restrict__ d2out,
             int2* __restrict__ corr /* MODE 1: {target index or -1, d2 bits} in one 8-byte record */,
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
    unsigned long long key = 0ull;
    int code = 4;
    if (lcg.ok) code = (MODE == 2) ? lcg_query<false>(lcg, x, y, z, key) : lcg_query<true>(lcg, x, y, z, key);
    const bool hit = code == 0;
    if (MODE == 1 && !hit && valid) atomicAdd(&st->miss[code], 1u);
    if (__any_sync(KSS_FULL, !hit)) {            // some query left the grid's box: exact pyramid search for the warp
        const unsigned long long k2 = (MODE == 2) ? lg_warp_nn<false>(py, x, y, z, slots[warp])
                                                  : lg_warp_nn<true>(py, x, y, z, slots[warp]);
        if (!hit) key = k2;
    }
    if (!valid) return;
    const float d2 = __uint_as_float((unsigned)(key >> 32));
    if (MODE == 1) corr[o] = make_int2(((double)d2 > max_dist_sqr) ? -1 : (int)(key & 0xffffffffu), __float_as_int(d2));   // A.3
    else d2out[o] = d2;
    if (MODE == 0) idx[o] = (int)(key & 0xffffffffu);
}

// ------------------------------------------------------------------ canonical reductions, large n
// level 1: one warp per 256-element chunk of ORIGINAL indices (lane-strided partials + butterfly)
// upper levels: the last CTA to finish reduces the chunk partials by the same rule (<= 2 more levels)
template <int NQ>
__device__ __forceinline__ void lg_finish_f32(const float* __restrict__ part, int nchunks, float* out /* [NQ] smem */) {
    // called by one full CTA (256 threads = 8 warps); part is [nchunks][NQ].  Level 2: one warp per 256 partials,
    // all NQ quantities at once, the lane's 8 rows loaded up front (the accumulation order is unchanged).
    __shared__ float lvl2[256 * 16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int n2 = (nchunks + 255) >> 8;             // level-2 chunks (<= 256 supported -> n <= 16.7 M)
    for (int c = warp; c < n2; c += nwarps) {
        float v[8][NQ];
        bool have[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = (c << 8) + lane + 32 * u;
            have[u] = i < nchunks;
#pragma unroll
            for (int q = 0; q < NQ; ++q) v[u][q] = have[u] ? part[(size_t)i * NQ + q] : 0.0f;
        }
        float p[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) p[q] = 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (have[u])
#pragma unroll
                for (int q = 0; q < NQ; ++q) p[q] = __fadd_rn(p[q], v[u][q]);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
            for (int q = 0; q < NQ; ++q) p[q] = __fadd_rn(p[q], __shfl_xor_sync(KSS_FULL, p[q], off));
        if (lane == 0)
#pragma unroll
            for (int q = 0; q < NQ; ++q) lvl2[c * 16 + q] = p[q];
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
        double v[8];
        bool have[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = (c << 8) + lane + 32 * u; have[u] = i < nchunks; v[u] = have[u] ? part[i] : 0.0; }
        double p = 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (have[u]) p = __dadd_rn(p, v[u]);
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
lg_passA_kernel(const float4* __restrict__ cur, const float4* __restrict__ t_orig, const int2* __restrict__ corr,
                int n, int nchunks, float* __restrict__ partA /* [nchunks][6] */,
                double* __restrict__ partD, int* __restrict__ partK, LgState* __restrict__ st) {
    if (st->done) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 8 + warp;
    if (c < nchunks) {
        float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        double dsum = 0.0;
        int k = 0;
        const int i0 = (c << 8) + lane;
        // all loads of the lane's 8 slots are issued before the (ordered) accumulation
        int m[8]; float4 sv[8], tv[8]; float dv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + 32 * u;
            m[u] = -1;
            if (i < n) { const int2 cr = corr[i]; m[u] = cr.x; dv[u] = __int_as_float(cr.y); }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + 32 * u;
            if (m[u] >= 0) { sv[u] = cur[i]; tv[u] = __ldg(t_orig + m[u]); }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (m[u] < 0) continue;
            a[0] = __fadd_rn(a[0], sv[u].x); a[1] = __fadd_rn(a[1], sv[u].y); a[2] = __fadd_rn(a[2], sv[u].z);
            a[3] = __fadd_rn(a[3], tv[u].x); a[4] = __fadd_rn(a[4], tv[u].y); a[5] = __fadd_rn(a[5], tv[u].z);
            dsum = __dadd_rn(dsum, (double)dv[u]);
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
lg_passB_kernel(const float4* __restrict__ cur, const float4* __restrict__ t_orig, const int2* __restrict__ corr,
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
        const int i0 = (c << 8) + lane;
        int m[8]; float4 sv[8], tv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + 32 * u; m[u] = i < n ? corr[i].x : -1; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + 32 * u;
            if (m[u] >= 0) { sv[u] = cur[i]; tv[u] = __ldg(t_orig + m[u]); }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (m[u] < 0) continue;
            const float4 s = sv[u], t = tv[u];
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
#pragma unroll
        for (int i = 0; i < 9; ++i) sigma[i] = mul_(st->one_over_n, tot[i]);
#pragma unroll
        for (int i = 0; i < 3; ++i) { sm[i] = st->smean[i]; dm[i] = st->dmean[i]; }
        umeyama_finish(sigma, sm, dm, T);
#pragma unroll
        for (int i = 0; i < 16; ++i) F[i] = st->fin[i];
        mat4_mul(T, F, F);
#pragma unroll
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
        for (int i = 0; i < 5; ++i) st->miss[i] = 0u;
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

bool lcg_enabled() { const char* e = getenv("KSS_NO_LCG"); return !(e && e[0] == '1'); }

// large candidate grid over the target's bounding box (device bbox from order_cloud's "<tag>_bb")
int build_lcg(Ctx& c, const float4* t_orig, int n_t, LcgView* view) {
    memset(view, 0, sizeof(*view));
    if (!lcg_enabled()) return KSS_OK;
    unsigned* bb = c.get<unsigned>("lg_t_bb", 8);
    if (c.err) return c.err;
    unsigned hb[6];
    if (cudaMemcpyAsync(hb, bb, sizeof(hb), cudaMemcpyDeviceToHost, c.st) != cudaSuccess) return KSS_ERR_CUDA;
    if (cudaStreamSynchronize(c.st) != cudaSuccess) return KSS_ERR_CUDA;
    auto o2f = [](unsigned u) { unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u; float f; memcpy(&f, &v, 4); return f; };
    float lo[3], ext[3], emax = 0.f;
    for (int a = 0; a < 3; ++a) { lo[a] = o2f(hb[a]); ext[a] = o2f(hb[3 + a]) - lo[a]; emax = std::max(emax, ext[a]); }
    if (!(emax > 0.f)) return KSS_OK;                                   // degenerate cloud: pyramid only
    const float margin = 0.02f * emax + 1e-6f;
    double vol = 1.0;
    for (int a = 0; a < 3; ++a) { lo[a] -= margin; ext[a] += 2.f * margin; vol *= std::max(ext[a], 1e-3f * emax); }
    const double cells_target = std::min(67108864.0, std::max(4096.0, 32.0 * (double)n_t));
    const double h = std::cbrt(vol / cells_target);
    LcgGeom g;
    int L = 0;
    for (int a = 0; a < 3; ++a) {
        int b = (int)std::lround(std::log2(std::max(ext[a], 1e-3f * emax) / h));
        b = std::min(10, std::max(0, b));
        g.bits[a] = b; L = std::max(L, b);
        g.lo[a] = lo[a]; g.h[a] = ext[a] / (float)(1 << b); g.inv_h[a] = (float)(1 << b) / ext[a];
    }
    g.levels = L + 1;
    size_t hdr_total = 0;
    std::vector<size_t> hdr_off(g.levels);
    for (int l = 0; l < g.levels; ++l) {
        hdr_off[l] = hdr_total;
        hdr_total += (size_t)1 << (lcg_bits(g, l, 0) + lcg_bits(g, l, 1) + lcg_bits(g, l, 2));
    }
    const size_t finest = (size_t)1 << (g.bits[0] + g.bits[1] + g.bits[2]);
    // list storage: two ping-pong scratch regions for the intermediate levels (a level only reads its
    // parent level) and one region for the finest level, which is the only one queries use
    const unsigned long long cap_s = 112ull * (unsigned long long)n_t + 65536ull;
    const unsigned long long cap_f = 4ull * finest + 32ull * (unsigned long long)n_t + 65536ull;
    unsigned long long* hdr = c.get<unsigned long long>("lcg_hdr", hdr_total);
    // u32 arena: [scratch A | scratch B | persistent far lists of intermediate levels]; float4 arena: finest level
    unsigned* arena = c.get<unsigned>("lcg_arena", 2 * cap_s);
    float4* arena4 = c.get<float4>("lcg_arena4", cap_f);
    unsigned long long* cursor = c.get<unsigned long long>("lcg_cursor", 2);
    int* ok = c.get<int>("lcg_ok", 2);
    if (c.err) { c.err = 0; return KSS_OK; }                             // not enough memory: pyramid only
    const int one = 1;
    cudaMemcpyAsync(ok, &one, sizeof(int), cudaMemcpyHostToDevice, c.st);
    unsigned long long prev_cursor = 0;
    double mean_parent = (double)n_t;
    // the hierarchy starts at the last level with <= 256 cells, built directly from all targets (one block per cell)
    int l_start = 0;
    for (int l = 0; l < g.levels; ++l)
        if ((lcg_bits(g, l, 0) + lcg_bits(g, l, 1) + lcg_bits(g, l, 2)) <= 8) l_start = l;
    unsigned* worklist = nullptr; unsigned* nwork = nullptr;
    for (int l = l_start; l < g.levels; ++l) {
        LcgLevelArgs a;
        memset(&a, 0, sizeof(a));
        a.level = l;
        for (int k = 0; k < 3; ++k) {
            a.bits[k] = lcg_bits(g, l, k); a.pbits[k] = l ? lcg_bits(g, l - 1, k) : 0;
            a.lo[k] = g.lo[k]; a.h[k] = ext[k] / (float)(1 << a.bits[k]);
        }
        a.tgt = t_orig; a.n_t = n_t;
        a.hdr_prev = l ? hdr + hdr_off[l - 1] : nullptr; a.hdr_cur = hdr + hdr_off[l];
        a.from_all = l == l_start ? 1 : 0;
        const bool last = l == g.levels - 1;
        const unsigned long long region = (unsigned long long)(l & 1) * cap_s;
        a.arena = arena; a.cursor = cursor; a.cap = region + cap_s; a.persist_all = last ? 1 : 0;
        a.cap_persist = cap_f; a.arena4 = last ? arena4 : nullptr;
        a.ok = ok;
        const unsigned long long init2[2] = {region, 0ull};                 // scratch cursor, float4-arena cursor
        cudaMemcpyAsync(cursor, init2, sizeof(init2), cudaMemcpyHostToDevice, c.st);
        if (cudaStreamSynchronize(c.st) != cudaSuccess) return KSS_ERR_CUDA;
        prev_cursor = region;
        const long long ncells = 1ll << (a.bits[0] + a.bits[1] + a.bits[2]);
        const int gs_used = (a.from_all || (mean_parent > 1024.0 && ncells <= (1 << 20))) ? 256 : (mean_parent > 40.0 ? 32 : 1);
        if (gs_used == 256) lcg_level_kernel<256><<<(unsigned)ncells, 256, 0, c.st>>>(a);
        else if (gs_used == 32) lcg_level_kernel<32><<<(unsigned)((ncells * 32 + 255) / 256), 256, 0, c.st>>>(a);
        else if (!last) lcg_level_kernel<1><<<(unsigned)((ncells + 255) / 256), 256, 0, c.st>>>(a);
        else {
            // finest level, two passes: (1) inherit or enqueue, (2) build the enqueued cells with dense warps
            if (!worklist) { worklist = c.get<unsigned>("lcg_work", (size_t)ncells * 2); nwork = c.get<unsigned>("lcg_nwork", 2); }
            if (c.err) { c.err = 0; return KSS_OK; }
            cudaMemsetAsync(nwork, 0, 8, c.st);
            LcgLevelArgs p1 = a; p1.worklist_out = worklist; p1.worklist_long = worklist + ncells; p1.nwork_out = nwork;
            lcg_level_kernel<1><<<(unsigned)((ncells + 255) / 256), 256, 0, c.st>>>(p1);
            unsigned hn[2] = {0, 0};
            if (cudaMemcpyAsync(hn, nwork, sizeof(hn), cudaMemcpyDeviceToHost, c.st) != cudaSuccess) return KSS_ERR_CUDA;
            if (cudaStreamSynchronize(c.st) != cudaSuccess) return KSS_ERR_CUDA;
            if (getenv("KSS_LCG_VERBOSE"))
                fprintf(stderr, "[lcg] finest level: %u short + %u long of %lld cells are built, the rest inherit\n", hn[0], hn[1], ncells);
            if (hn[0]) {
                LcgLevelArgs p2 = a; p2.worklist = worklist; p2.nwork = hn[0];
                lcg_level_kernel<1><<<(hn[0] + 255) / 256, 256, 0, c.st>>>(p2);
            }
            if (hn[1]) {
                LcgLevelArgs p3 = a; p3.worklist = worklist + ncells; p3.nwork = hn[1];
                lcg_level_kernel<32><<<(unsigned)(((unsigned long long)hn[1] * 32 + 255) / 256), 256, 0, c.st>>>(p3);
            }
            c.launched(2);
        }
        c.launched();
        unsigned long long cur2[2] = {0, 0};
        if (cudaMemcpyAsync(cur2, cursor, sizeof(cur2), cudaMemcpyDeviceToHost, c.st) != cudaSuccess) return KSS_ERR_CUDA;
        if (cudaStreamSynchronize(c.st) != cudaSuccess) return KSS_ERR_CUDA;
        const unsigned long long cur = last ? cur2[1] + region : cur2[0];
        mean_parent = (double)(cur - prev_cursor) / (double)ncells;
        if (getenv("KSS_LCG_VERBOSE")) {
            static double t_prev = 0; struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
            const double t_now = ts.tv_sec + 1e-9 * ts.tv_nsec;
            fprintf(stderr, "[lcg] level %d GS %d bits %d,%d,%d cells %lld entries %llu mean %.2f  (+%.3f ms)\n", l, gs_used, a.bits[0], a.bits[1],
                    a.bits[2], ncells, cur - prev_cursor, mean_parent, l > l_start ? (t_now - t_prev) * 1e3 : 0.0);
            t_prev = t_now;
        }
        prev_cursor = cur;
    }
    int hok = 0;
    if (cudaMemcpyAsync(&hok, ok, sizeof(int), cudaMemcpyDeviceToHost, c.st) != cudaSuccess) return KSS_ERR_CUDA;
    if (cudaStreamSynchronize(c.st) != cudaSuccess) return KSS_ERR_CUDA;
    view->hdr = hdr + hdr_off[g.levels - 1];
    view->arena4 = arena4;
    view->arena = arena;
    view->tgt = t_orig;
    view->g = g;
    view->ok = hok;
    if (getenv("KSS_LCG_VERBOSE"))
        fprintf(stderr, "[lcg] n_t=%d bits=%d,%d,%d levels=%d cells=%zu caps %llu/%llu finest mean list %.2f ok=%d\n", n_t,
                g.bits[0], g.bits[1], g.bits[2], g.levels, finest, cap_s, cap_f, mean_parent, hok);
    return c.ok() ? KSS_OK : c.err;
}

inline int nn_grid(int n_q) { return (n_q + LG_WARPS * 32 - 1) / (LG_WARPS * 32); }

}  // namespace

int large_nn_device(cudaStream_t st, long long* launches, const double* d_q, int n_q, const double* d_t, int n_t,
                    int* d_idx, float* d_d2, const DevAlloc& alloc) {
    Ctx c{st, launches, alloc};
    Pyramid py;
    float4* t_orig = nullptr;
    int r = build_pyramid(c, d_t, n_t, &py, &t_orig);
    if (r) return r;
    LcgView lcg;
    r = build_lcg(c, t_orig, n_t, &lcg);
    if (r) return r;
    float4* q4 = c.get<float4>("lg_q4", n_q);
    int* perm = c.get<int>("lg_perm", n_q);
    if (c.err) return c.err;
    r = order_cloud(c, "lg_q", d_q, n_q, q4, nullptr, 0, perm);
    if (r) return r;
    lg_nn_kernel<0><<<nn_grid(n_q), LG_WARPS * 32, 0, st>>>(py, lcg, t_orig, perm, n_q, q4, nullptr, d_idx, d_d2, nullptr, nullptr, 0.0);
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
    LcgView lcg;
    r = build_lcg(c, t_orig, n_t, &lcg);
    if (r) return r;
    static_assert(sizeof(LcgView) <= sizeof(run->lcg), "LargeIcp::lcg too small");
    memcpy(run->lcg, &lcg, sizeof(lcg));
    float4* inp = c.get<float4>("lg_inp", n_s);
    float4* cur = c.get<float4>("lg_cur", n_s);
    int* perm = c.get<int>("lg_perm", n_s);
    int* idx = c.get<int>("lg_corr", (size_t)n_s * 2);      // int2 {index, d2 bits} per source point
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
    LcgView lcg; memcpy(&lcg, run->lcg, sizeof(lcg));
    LgState* state = (LgState*)run->state;
    const double max2 = prm->max_corr_dist * prm->max_corr_dist;
    const int n = run->n_s, nch = run->nchunks;
    const double mse_abs = prm->fitness_eps < 0.0 ? -1.0 : 1e-12;   // fitness_eps < 0: never converge (steady-state timing)
    for (int k = 0; k < count; ++k) {
        if (run->mark) run->mark(run->mark_user, KSS_STAGE_LARGE_NN, 1);
        lg_nn_kernel<1><<<nn_grid(n), LG_WARPS * 32, 0, st>>>(py, lcg, (const float4*)run->t_orig, run->perm, n, (float4*)run->cur, nullptr, nullptr,
                                                              nullptr, (int2*)run->idx, state, max2);
        if (run->mark) { run->mark(run->mark_user, KSS_STAGE_LARGE_NN, 0); run->mark(run->mark_user, KSS_STAGE_LARGE_REDUCE, 1); }
        lg_passA_kernel<<<(nch + 7) / 8, 256, 0, st>>>((const float4*)run->cur, (const float4*)run->t_orig, (const int2*)run->idx,
                                                       n, nch, run->partA, run->partD, run->partK, state);
        lg_passB_kernel<<<(nch + 7) / 8, 256, 0, st>>>((const float4*)run->cur, (const float4*)run->t_orig, (const int2*)run->idx, n,
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
    LcgView lcg; memcpy(&lcg, run->lcg, sizeof(lcg));
    lg_nn_kernel<2><<<nn_grid(run->n_s), LG_WARPS * 32, 0, st>>>(py, lcg, (const float4*)run->t_orig, run->perm, run->n_s, nullptr, (const float4*)run->inp,
                                                               nullptr, run->d2, nullptr, state, 0.0);
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
    if (getenv("KSS_LCG_VERBOSE"))
        fprintf(stderr, "[lcg] grid misses over the run: outside box %u, empty %u, far/long %u, grid off %u (iterations %d)\n", h.miss[1], h.miss[2], h.miss[3], h.miss[4], h.iters);
    if (T) for (int i = 0; i < 16; ++i) T[i] = h.fin[i];
    if (fitness) *fitness = o3[0];
    if (iters) *iters = h.iters;
    if (converged) *converged = h.converged;
    return KSS_OK;
}

}  // namespace kss
