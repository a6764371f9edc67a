// kss_large.cu -- clouds beyond one CTA's shared memory (> 2048 points).
//
// Replaces pcl::KdTreeFLANN + the PCL ICP loop at full resolution (shapeRegistration_ICP(int),
// KSS_ICP.hpp:133-183; PCR_QM, registrationMeasure.hpp:47-98) with
//   build : (all on the device, no host synchronisation) bounding box, a 256-sample nearest-neighbour probe that
//           sets the cell size h to ~2 point spacings, a uniform grid of fine cells grouped into 4x4x4 BLOCKS,
//           targets counting-sorted by (Morton block, fine cell), one 64-entry offset table per occupied block,
//           and a 32-ary box pyramid over 32-point tiles of the sorted array (exact fallback for far queries)
//   nn    : one CTA = 512 consecutive queries of the Morton-sorted source.  The CTA finds the blocks its queries can
//           touch, stages their points and offset tables in shared memory with cp.async.bulk (TMA 1-D bulk copies
//           completing on an mbarrier), then every thread searches 2x2x2 fine cells around its query (4x4x4 if the
//           first stage cannot prove exactness) from shared memory.  A search is accepted only when the best distance
//           is below the distance to the unsearched region, so the result is the exact minimum over ALL targets
//           (ties: lowest ORIGINAL index); anything else falls back to the box pyramid (seeded with the distance to
//           the previous match when there is one).
//   icp   : per source point, by ORIGINAL index, {position, certificate} and {match}: the certificate is a lower bound
//           of the distance to every target other than the match.  lg_track_kernel streams the state once per
//           iteration (transform, certificate test, re-certification from the 7-neighbour list of the match, pass A of
//           the canonical sums for chunks without open points); lg_refine_kernel / lg_nn_kernel<1> / lg_left_kernel
//           search what stays open; see the comment above LG_CERT_UP.
//   reduce: CANON256 sums by ORIGINAL source index (one warp per 256-element chunk, SoA partials, the upper levels by
//           whichever warp completes a group of 256 chunks), umeyama/SVD on one thread, device-side convergence
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "kss_device.cuh"
#include "kss_large.h"

#ifndef KSS_SCAN_UNROLL
#define KSS_SCAN_UNROLL 4
#endif

namespace kss {

constexpr int SCAN_UNROLL = KSS_SCAN_UNROLL;
#ifndef KSS_LG_CUBE3
#define KSS_LG_CUBE3 3.0f
#endif
constexpr float cube3 = KSS_LG_CUBE3;  // a warp scans the cube q +- bound cell by cell up to this many cells of reach; beyond, the (seeded) box pyramid
#ifndef KSS_LG_STAGED_DIV
#define KSS_LG_STAGED_DIV 32u
#endif
constexpr unsigned LG_STAGED_DIV = KSS_LG_STAGED_DIV;   // more than n / this many points left open by the refine kernel: the staged kernel (512 per CTA) instead of one warp each
constexpr int LG_MAX_LEVELS = 4;     // 32^4 tiles * 32 points = 33.5 M points
constexpr int LG_SAMPLES = 256;      // nearest-neighbour probe of the build

struct Pyramid {
    const float4* tp;                // (Morton block, fine cell) ordered targets {x,y,z,bits(orig)}, padded to x32
    int nlev;                        // number of box levels (level 0 = tiles)
    int cnt[LG_MAX_LEVELS];          // boxes per level
    int pad[LG_MAX_LEVELS];          // allocated boxes per level (multiple of 32)
    const float* box[LG_MAX_LEVELS]; // SoA [6][pad[l]]
};

// geometry of the block grid; written by lg_geom_kernel, read by everything after it
struct LgGeom {
    float lo[3];                     // bounding-box corner
    float h, inv_h;                  // fine cell size (cubic)
    int nf[3];                       // fine cells per axis
    int nb[3];                       // blocks per axis (4 fine cells each)
    int bbits[3];                    // ceil(log2(nb))
    int nbp;                         // 1 << (sum of bbits): entries of the Morton-compact block table
    int pad_[2];
};
constexpr unsigned LG_EMPTY = 0xffffffffu;

struct LgGridView {
    const LgGeom* geom;
    const unsigned* blk_rank;        // [nbp] block code -> rank among occupied blocks, or LG_EMPTY
    const unsigned* fine_start;      // [n_occ * 64 + 68] absolute start offset of every fine cell (ends with n_t)
    const float4* tp;                // sorted targets
    const float4* t_orig;            // targets by original index
    const unsigned* lut;             // [3][1024] per-axis bit spreads of the Morton-compact block code: code = x | y | z
    const float4* knn;               // neighbour lists by ORIGINAL target index (ICP runs only), see lg_knn_kernel
    int n_t;
};
constexpr int LG_LUT = 1024;         // blocks per axis are at most ~1001 (lg_geom_kernel)

// ------------------------------------------------------------------ build kernels
__global__ void lg_init_bbox_kernel(unsigned* bb, float* sample_d2) {
    if (threadIdx.x < 3) { bb[threadIdx.x] = 0xffffffffu; bb[3 + threadIdx.x] = 0u; }
    if (sample_d2) for (int i = threadIdx.x; i < LG_SAMPLES; i += blockDim.x) sample_d2[i] = __int_as_float(0x7f800000);
}

// double[n][3] -> float4 {x,y,z,bits(i)} with RN narrowing (KSS_ICP.hpp:137-152), plus bounding box
__global__ void __launch_bounds__(256)
lg_convert_bbox_kernel(const double* __restrict__ pts, int n, float4* __restrict__ out, unsigned* __restrict__ bb) {
    // grid-stride, one set of atomics per CTA (six atomics per WARP on the same six words cost 0.12 ms at 1M points)
    unsigned mn[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, mx[3] = {0u, 0u, 0u};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float x = (float)pts[3 * (size_t)i], y = (float)pts[3 * (size_t)i + 1], z = (float)pts[3 * (size_t)i + 2];
        out[i] = make_float4(x, y, z, __int_as_float(i));
        const unsigned ox = f2ord(x), oy = f2ord(y), oz = f2ord(z);
        mn[0] = min(mn[0], ox); mx[0] = max(mx[0], ox);
        mn[1] = min(mn[1], oy); mx[1] = max(mx[1], oy);
        mn[2] = min(mn[2], oz); mx[2] = max(mx[2], oz);
    }
    __shared__ unsigned s_mn[3][8], s_mx[3][8];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const unsigned m0 = __reduce_min_sync(KSS_FULL, mn[a]);
        const unsigned m1 = __reduce_max_sync(KSS_FULL, mx[a]);
        if ((threadIdx.x & 31) == 0) { s_mn[a][threadIdx.x >> 5] = m0; s_mx[a][threadIdx.x >> 5] = m1; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned m0 = 0xffffffffu, m1 = 0u;
        for (int w = 0; w < 8; ++w) { m0 = min(m0, s_mn[threadIdx.x][w]); m1 = max(m1, s_mx[threadIdx.x][w]); }
        if (m0 <= m1) { atomicMin(&bb[threadIdx.x], m0); atomicMax(&bb[3 + threadIdx.x], m1); }
    }
}

// ---- target grid
// nearest-neighbour probe: LG_SAMPLES evenly strided targets against all targets (brute force, tiles of 1024 targets in
// shared memory, thread = sample); sample_d2[s] = smallest POSITIVE squared distance (coincident points do not count)
__global__ void __launch_bounds__(LG_SAMPLES)
lg_probe_kernel(const float4* __restrict__ p4, int n, float* __restrict__ sample_d2) {
    __shared__ float4 tile[1024];
    const int s = threadIdx.x;
    const int si = (int)(((long long)s * n) / LG_SAMPLES);
    const float4 q = p4[si < n ? si : n - 1];
    const int base = blockIdx.x * 1024;
    for (int j = threadIdx.x; j < 1024; j += LG_SAMPLES) {
        const int i = base + j;
        tile[j] = i < n ? p4[i] : make_float4(PAD_COORD, PAD_COORD, PAD_COORD, 0.0f);
    }
    __syncthreads();
    float best = __int_as_float(0x7f800000);
#pragma unroll 8
    for (int j = 0; j < 1024; ++j) {
        const float4 p = tile[j];
        const float d = d2_rn(q.x, q.y, q.z, p.x, p.y, p.z);
        best = (d > 0.0f && d < best) ? d : best;
    }
    if (best < __int_as_float(0x7f800000)) atomicMin(reinterpret_cast<unsigned*>(sample_d2) + s, __float_as_uint(best));   // positive floats order as uints
}

__host__ __device__ inline int lg_ceil_log2(int v) { int b = 0; while ((1 << b) < v) ++b; return b; }

// h = hc * 2 * median positive NN distance of the probe (2 * E[d_nn] is the point spacing of a uniformly sampled
// surface; hc = 2 measured best on the 1M-point scan: ~4 points per occupied fine cell), enlarged until the Morton-compact block table fits max_bits bits
__global__ void __launch_bounds__(LG_SAMPLES)
lg_geom_kernel(const unsigned* __restrict__ bb, const float* __restrict__ sample_d2, int n, float hc, int max_bits,
               LgGeom* __restrict__ g) {
    __shared__ unsigned keys[LG_SAMPLES];
    keys[threadIdx.x] = __float_as_uint(sample_d2[threadIdx.x]);        // +inf (no positive neighbour) sorts last
    bitonic_sort_smem<LG_SAMPLES>(keys);
    if (threadIdx.x != 0) return;
    int valid = 0;
    while (valid < LG_SAMPLES && keys[valid] < 0x7f800000u) ++valid;
    float lo[3], ext[3], emax = 0.0f;
    for (int a = 0; a < 3; ++a) { lo[a] = ord2f(bb[a]); ext[a] = ord2f(bb[3 + a]) - lo[a]; emax = fmaxf(emax, ext[a]); }
    float h = emax > 0.0f ? emax * 0.015625f : 1.0f;
    if (valid > 0) h = hc * 2.0f * sqrtf(__uint_as_float(keys[valid >> 1]));
    if (!(h > 0.0f) || !(h < 3.0e38f)) h = emax > 0.0f ? emax : 1.0f;
    if (emax > 0.0f) h = fmaxf(h, emax * (1.0f / 4000.0f));             // at most ~4000 fine cells (1000 blocks) per axis
    int nf[3], nb[3], bbits[3];
    for (int it = 0; it < 200; ++it) {
        int tot = 0;
        for (int a = 0; a < 3; ++a) {
            nf[a] = (int)(ext[a] / h) + 1;
            nb[a] = (nf[a] + 3) >> 2;
            bbits[a] = lg_ceil_log2(nb[a]);
            tot += bbits[a];
        }
        if (tot <= max_bits) break;
        h *= 1.26f;
    }
    for (int a = 0; a < 3; ++a) { g->lo[a] = lo[a]; g->nf[a] = nf[a]; g->nb[a] = nb[a]; g->bbits[a] = bbits[a]; }
    g->h = h; g->inv_h = 1.0f / h;
    g->nbp = 1 << (bbits[0] + bbits[1] + bbits[2]);
    g->pad_[0] = n; g->pad_[1] = 0;
}

// fine cell of a coordinate (the ONE definition used for binning; queries add slack for its rounding)
__device__ __forceinline__ int lg_fine(float v, float lo, float inv_h, int nf) {
    const float f = (v - lo) * inv_h;
    int c = (int)fminf(fmaxf(f, 0.0f), 2.0e9f);
    return c < nf ? c : nf - 1;
}
// Morton-compact block code: bit k of an axis is interleaved only while k < bbits[axis]
__device__ __forceinline__ unsigned lg_block_code(int bx, int by, int bz, const int* bbits) {
    unsigned code = 0; int sh = 0;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        if (k < bbits[0]) { code |= (unsigned)((bx >> k) & 1) << sh; ++sh; }
        if (k < bbits[1]) { code |= (unsigned)((by >> k) & 1) << sh; ++sh; }
        if (k < bbits[2]) { code |= (unsigned)((bz >> k) & 1) << sh; ++sh; }
    }
    return code;
}

// per-axis tables so that the block code is three loads and two ORs in the search kernels
__global__ void __launch_bounds__(256)
lg_lut_kernel(const LgGeom* __restrict__ gp, unsigned* __restrict__ lut) {
    const LgGeom g = *gp;
    for (int i = threadIdx.x; i < 3 * LG_LUT; i += blockDim.x) {
        const int a = i / LG_LUT, v = i % LG_LUT;
        lut[i] = lg_block_code(a == 0 ? v : 0, a == 1 ? v : 0, a == 2 ? v : 0, g.bbits);
    }
}
__device__ __forceinline__ unsigned lg_code(const unsigned* __restrict__ lut, int bx, int by, int bz) {
    return __ldg(lut + bx) | __ldg(lut + LG_LUT + by) | __ldg(lut + 2 * LG_LUT + bz);
}

__global__ void __launch_bounds__(256)
lg_bin_kernel(const float4* __restrict__ p4, int n, const LgGeom* __restrict__ gp, unsigned* __restrict__ pkey,
              unsigned* __restrict__ blk_cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const LgGeom g = *gp;
    const float4 p = p4[i];
    const int ix = lg_fine(p.x, g.lo[0], g.inv_h, g.nf[0]), iy = lg_fine(p.y, g.lo[1], g.inv_h, g.nf[1]),
              iz = lg_fine(p.z, g.lo[2], g.inv_h, g.nf[2]);
    const unsigned b = lg_block_code(ix >> 2, iy >> 2, iz >> 2, g.bbits);
    const unsigned f = (unsigned)(((iz & 3) << 4) | ((iy & 3) << 2) | (ix & 3));
    pkey[i] = (b << 6) | f;
    atomicAdd(&blk_cnt[b], 1u);
}

// one 64-bit scan gives both the rank among occupied blocks (high word) and the point base (low word)
__global__ void __launch_bounds__(1024)
lg_bscan1_kernel(const unsigned* __restrict__ blk_cnt, const LgGeom* __restrict__ gp, unsigned long long* __restrict__ scan,
                 unsigned long long* __restrict__ totals) {
    __shared__ unsigned long long ws[32];
    const int m = gp->nbp;
    if (blockIdx.x * 1024 >= m) return;
    const int i = blockIdx.x * 1024 + threadIdx.x;
    const unsigned c = i < m ? blk_cnt[i] : 0u;
    const unsigned long long v = c ? ((1ull << 32) | c) : 0ull;
    unsigned long long x = v;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(KSS_FULL, x, o); if (lane >= o) x += y; }
    if (lane == 31) ws[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = ws[lane];
        for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(KSS_FULL, w, o); if (lane >= o) w += y; }
        ws[lane] = w;
    }
    __syncthreads();
    const unsigned long long incl = x + (warp ? ws[warp - 1] : 0ull);
    if (i < m) scan[i] = incl - v;
    if (threadIdx.x == 1023) totals[blockIdx.x] = incl;
}
__global__ void __launch_bounds__(1024)
lg_bscan2_kernel(unsigned long long* __restrict__ totals, const LgGeom* __restrict__ gp, int* __restrict__ n_occ) {
    __shared__ unsigned long long ws[32];
    __shared__ unsigned long long carry;
    const int nb = (gp->nbp + 1023) >> 10;
    if (threadIdx.x == 0) carry = 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned long long v = i < nb ? totals[i] : 0ull;
        unsigned long long x = v;
        for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(KSS_FULL, x, o); if (lane >= o) x += y; }
        if (lane == 31) ws[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = ws[lane];
            for (int o = 1; o < 32; o <<= 1) { unsigned long long y = __shfl_up_sync(KSS_FULL, w, o); if (lane >= o) w += y; }
            ws[lane] = w;
        }
        __syncthreads();
        const unsigned long long incl = x + (warp ? ws[warp - 1] : 0ull) + carry;
        if (i < nb) totals[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_occ = (int)(carry >> 32);
}
__global__ void __launch_bounds__(1024)
lg_bscan3_kernel(const unsigned* __restrict__ blk_cnt, const unsigned long long* __restrict__ scan,
                 const unsigned long long* __restrict__ totals, const LgGeom* __restrict__ gp,
                 unsigned* __restrict__ blk_rank, unsigned* __restrict__ blk_base) {
    const int m = gp->nbp;
    const int i = blockIdx.x * 1024 + threadIdx.x;
    if (i >= m) return;
    const unsigned c = blk_cnt[i];
    if (!c) { blk_rank[i] = LG_EMPTY; return; }
    const unsigned long long e = scan[i] + totals[blockIdx.x];
    const unsigned r = (unsigned)(e >> 32);
    blk_rank[i] = r;
    blk_base[r] = (unsigned)(e & 0xffffffffull);
}

// zero the counters of the occupied blocks (their number is only known on the device): grid-stride
__global__ void __launch_bounds__(256)
lg_zero_fine_kernel(const int* __restrict__ n_occ, uint4* __restrict__ fine_cnt4) {
    const long long total = (long long)*n_occ * 16;           // uint4 per block row = 64 counters / 4
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x)
        fine_cnt4[i] = make_uint4(0u, 0u, 0u, 0u);
}
__global__ void __launch_bounds__(256)
lg_fine_hist_kernel(const unsigned* __restrict__ pkey, int n, const unsigned* __restrict__ blk_rank, unsigned* __restrict__ fine_cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned k = pkey[i];
    atomicAdd(&fine_cnt[(size_t)blk_rank[k >> 6] * 64 + (k & 63u)], 1u);
}
// one warp per occupied block: offsets of its 64 fine cells (absolute), counters reset to serve as scatter cursors
__global__ void __launch_bounds__(256)
lg_fine_scan_kernel(const int* __restrict__ n_occ, const unsigned* __restrict__ blk_base, unsigned* __restrict__ fine_cnt,
                    unsigned* __restrict__ fine_start, int n) {
    const int lane = threadIdx.x & 31;
    const int nocc = *n_occ;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = wid; r < nocc; r += nw) {
        uint2* c2 = reinterpret_cast<uint2*>(fine_cnt + (size_t)r * 64) + lane;
        const uint2 c = *c2;
        unsigned s = c.x + c.y, x = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(KSS_FULL, x, o); if (lane >= o) x += y; }
        const unsigned st = blk_base[r] + (x - s);
        reinterpret_cast<uint2*>(fine_start + (size_t)r * 64)[lane] = make_uint2(st, st + c.x);
        *c2 = make_uint2(0u, 0u);
    }
    if (wid == 0)                                                   // sentinel + padding for the 272-byte table copies
        for (int j = lane; j < 68; j += 32) fine_start[(size_t)nocc * 64 + j] = (unsigned)n;
}
__global__ void __launch_bounds__(256)
lg_grid_scatter_kernel(const float4* __restrict__ p4, int n, const unsigned* __restrict__ pkey, const unsigned* __restrict__ blk_rank,
                       const unsigned* __restrict__ fine_start, unsigned* __restrict__ fine_cnt, float4* __restrict__ tp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned k = pkey[i];
    const size_t c = (size_t)blk_rank[k >> 6] * 64 + (k & 63u);
    const unsigned pos = fine_start[c] + atomicAdd(&fine_cnt[c], 1u);
    tp[pos] = p4[i];
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

// ------------------------------------------------------------------ warp-cooperative pyramid NN (exact fallback)
struct NNState {
    float qx, qy, qz;
    float lx, ly, lz, hx, hy, hz;    // warp query box
    float best;
    unsigned long long bestkey;
    float4* slot;                    // this warp's 32-point staging tile in shared memory
};

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
        const unsigned long long key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_uint(p.w);
        s.bestkey = key < s.bestkey ? key : s.bestkey;
    }
    s.best = __uint_as_float((unsigned)(s.bestkey >> 32));
}

// visit the children (level L-1 boxes, or the tile when L == 0) of node `node` at level L
template <int L>
__device__ __forceinline__ void lg_descend(const Pyramid& py, int node, NNState& s) {
    if constexpr (L == 0) {
        lg_scan_tile(py, node, s);
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
            lg_descend<L - 1>(py, node * 32 + src, s);
        }
    }
}

// all 32 lanes carry a query (idle lanes repeat a neighbour's); returns (d2 bits << 32) | original index
// seed_d2: an upper bound of the lane's answer (the distance to ANY target, e.g. last iteration's match; +inf: none).
// Boxes and tiles are skipped only when strictly farther (with slack), so the seeding target itself is met again and
// returned with its index if nothing beats it.
__device__ __forceinline__ unsigned long long lg_warp_nn(const Pyramid& py, float qx, float qy, float qz, float4* slot,
                                                         float seed_d2 = __builtin_inff()) {
    NNState s;
    s.qx = qx; s.qy = qy; s.qz = qz;
    s.lx = warp_min_f(qx); s.ly = warp_min_f(qy); s.lz = warp_min_f(qz);
    s.hx = warp_max_f(qx); s.hy = warp_max_f(qy); s.hz = warp_max_f(qz);
    s.best = seed_d2;
    s.bestkey = ((unsigned long long)__float_as_uint(seed_d2) << 32) | 0xffffffffull;
    s.slot = slot;
    switch (py.nlev) {      // virtual root above the top level
        case 1: lg_descend<1>(py, 0, s); break;
        case 2: lg_descend<2>(py, 0, s); break;
        case 3: lg_descend<3>(py, 0, s); break;
        default: lg_descend<4>(py, 0, s); break;
    }
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
    unsigned miss[5];            // diagnostics: [0] certificates re-established from the neighbour lists, [1] queries sent to the search kernels, [2] pyramid fallbacks, [3] passes on global tables, [4] CTAs split in two passes
    unsigned n_unres;            // queries of this iteration the refine kernel left to the general kernels
    unsigned n_fail;             // queries of this iteration whose certificate did not hold (lg_track_kernel)
};

// Programmatic dependent launch: the kernels of an iteration are launched so that one may be scheduled while its
// predecessor drains (launch latency and independent loads overlap the predecessor's serial tail); nothing the
// predecessor writes may be read, and nothing written, before this returns.  A no-op in an ordinary launch.
__device__ __forceinline__ void lg_wait_prior() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------ staged grid search
constexpr int NN_THREADS = 256;
constexpr int NN_QPT = 2;                       // queries per thread
constexpr int NN_QPC = NN_THREADS * NN_QPT;     // queries per CTA
constexpr int NN_MAXBOX = 2048;                 // blocks in the bounding box of a pass's queries (bitmap of touched blocks)
constexpr int NN_MAXREG = 128;                  // blocks actually touched (the region)
#ifndef KSS_NN_MAXOCC
#define KSS_NN_MAXOCC 48
#define KSS_NN_PTS_CAP 2048
#define KSS_NN_CTAS 3
#endif
constexpr int NN_MAXOCC = KSS_NN_MAXOCC;        // occupied blocks staged
constexpr int NN_PTS_CAP = KSS_NN_PTS_CAP;      // staged points (40 KB)
__device__ unsigned lg_dbg[8];
__device__ unsigned lg_dbg2[4];                 // diagnostics: third stage {no bound, cube too large, failed, resolved}
__device__ unsigned long long lg_cnt[4];        // -DKSS_LG_COUNT: first-stage candidates {evaluated, lane slots, queries}                  // diagnostics: why passes could not stage (KSS_LG_VERBOSE)
constexpr int NN_TAB_STRIDE = 68;               // u32 per staged table: 65 used, copied as 272 bytes
constexpr int NN_RNG_STRIDE = 9;                // per-thread list of candidate ranges: 8 + sentinel (odd stride: no bank conflicts)
constexpr size_t NN_SMEM = (size_t)NN_PTS_CAP * sizeof(float4) + (size_t)NN_MAXOCC * NN_TAB_STRIDE * sizeof(unsigned) +
                           (size_t)NN_QPC * sizeof(float4) + (size_t)NN_THREADS * NN_RNG_STRIDE * sizeof(unsigned);
static_assert(NN_PTS_CAP < 65536 && NN_MAXOCC < 128, "ranges are packed as two u16, slots as int8");
static_assert(NN_MAXREG % 32 == 0 && 8 * TILE * sizeof(float4) <= NN_PTS_CAP * sizeof(float4), "pyramid tiles reuse the point area");

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "KSS_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra KSS_MBAR_DONE;\n"
        "bra KSS_MBAR_WAIT;\n"
        "KSS_MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on the mbarrier (sizes / addresses multiples of 16)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// block accessors: table of 65 absolute offsets + a pointer such that pts[absolute offset] is the candidate
struct SmemAcc {
    const float4* pts; const unsigned* tab; const signed char* slot; const int* delta;
    int bx0, by0, bz0, rx, ry;
    __device__ __forceinline__ bool block(int bx, int by, int bz, const unsigned*& t, const float4*& p) const {
        const int s = slot[((bz - bz0) * ry + (by - by0)) * rx + (bx - bx0)];
        if (s < 0) return false;
        t = tab + s * NN_TAB_STRIDE;
        p = pts + delta[s];
        return true;
    }
};
struct GlobAcc {
    const unsigned* blk_rank; const unsigned* fine_start; const float4* tp; const unsigned* lut;
    __device__ __forceinline__ bool block(int bx, int by, int bz, const unsigned*& t, const float4*& p) const {
        const unsigned r = __ldg(blk_rank + lg_code(lut, bx, by, bz));
        if (r == LG_EMPTY) return false;
        t = fine_start + (size_t)r * 64;
        p = tp;
        return true;
    }
};

struct Best { float d; unsigned idx; };          // smallest d2_rn so far, lowest original index among equals

// all targets of the fine cells [x0..x1] x [y0..y1] x [z0..z1] (inside the grid)
template <class Acc>
__device__ __forceinline__ void lg_scan_cells(const Acc& acc, int x0, int x1, int y0, int y1, int z0, int z1,
                                              float qx, float qy, float qz, Best& b) {
    for (int iz = z0; iz <= z1; ++iz)
        for (int iy = y0; iy <= y1; ++iy) {
            const int f = ((iz & 3) << 4) | ((iy & 3) << 2);
            for (int ix = x0; ix <= x1;) {
                const int bx = ix >> 2;
                const int xe = min(x1, (bx << 2) | 3);            // cells of one block along x are contiguous
                const unsigned* t; const float4* p;
                if (acc.block(bx, iy >> 2, iz >> 2, t, p)) {
                    const unsigned s = t[f + (ix & 3)], e = t[f + (xe & 3) + 1];
                    // d2 >= 0: float bits order as integers, so (d2 bits, index) compares as ONE 64-bit key -- smallest
                    // distance, lowest original index among equals
                    unsigned long long bk = ((unsigned long long)__float_as_uint(b.d) << 32) | b.idx;
                    // (lanes of a warp hold segments of different lengths: a deep unroll makes every lane pay the remainder
                    // blocks of the longest -- 4 measured best: 0.112 ms vs 0.125 ms with 8 for the 1M-point search)
#pragma unroll SCAN_UNROLL
                    for (unsigned j = s; j < e; ++j) {
                        const float4 c = p[j];
                        const unsigned long long ck = ((unsigned long long)__float_as_uint(d2_rn(qx, qy, qz, c.x, c.y, c.z)) << 32) | __float_as_uint(c.w);
                        bk = ck < bk ? ck : bk;
                    }
                    b.d = __uint_as_float((unsigned)(bk >> 32)); b.idx = (unsigned)bk;
                }
                ix = xe + 1;
            }
        }
}

// the query's position in cell units: nearest cell corner c and the distance m to it (max over the axes)
struct QCell { int cx, cy, cz; float dx, dy, dz, m; bool inside; };      // d* = signed offset from the corner, in cells
__device__ __forceinline__ QCell lg_qcell(const LgGeom& g, float x, float y, float z) {
    QCell q;
    const float fx = (x - g.lo[0]) * g.inv_h, fy = (y - g.lo[1]) * g.inv_h, fz = (z - g.lo[2]) * g.inv_h;
    // more than 8 cells outside the grid nothing can be found by either stage: clamping there only keeps the ints sane
    const float cxf = floorf(fminf(fmaxf(fx, -8.0f), (float)g.nf[0] + 8.0f) + 0.5f);
    const float cyf = floorf(fminf(fmaxf(fy, -8.0f), (float)g.nf[1] + 8.0f) + 0.5f);
    const float czf = floorf(fminf(fmaxf(fz, -8.0f), (float)g.nf[2] + 8.0f) + 0.5f);
    q.cx = (int)cxf; q.cy = (int)cyf; q.cz = (int)czf;
    q.dx = fx - cxf; q.dy = fy - cyf; q.dz = fz - czf;
    q.m = fmaxf(fabsf(q.dx), fmaxf(fabsf(q.dy), fabsf(q.dz)));
    // the 4x4x4 neighbourhood [c-2, c+1] meets the grid?  (m > 0.5 only for clamped, far-away queries)
    q.inside = q.cx + 1 >= 0 && q.cx - 2 < g.nf[0] && q.cy + 1 >= 0 && q.cy - 2 < g.nf[1] && q.cz + 1 >= 0 && q.cz - 2 < g.nf[2] &&
               q.m <= 0.75f;
    return q;
}
// the cells one query reads: offsets from the nearest corner, each in -2 .. 1, plus "bounded" (exact without proof)
__device__ __forceinline__ int lg_box_pack(int x0, int x1, int y0, int y1, int z0, int z1, bool bounded) {
    return (x0 + 2) | ((x1 + 2) << 2) | ((y0 + 2) << 4) | ((y1 + 2) << 6) | ((z0 + 2) << 8) | ((z1 + 2) << 10) | (bounded ? 1 << 12 : 0);
}
__device__ __forceinline__ int lg_box_get(int b, int i) { return ((b >> (2 * i)) & 3) - 2; }
__device__ __forceinline__ bool lg_box_bounded(int b) { return (b >> 12) & 1; }

// Every target outside the cells [c - reach, c + reach - 1]^3 is farther than (reach - m) cells from the query, up to the
// rounding of the binning (< 1e-4 cells, covered by the 0.002 slack), so a best distance below that bound is the exact
// nearest neighbour over ALL targets.
__device__ __forceinline__ bool lg_proven(const LgGeom& g, const QCell& q, int reach, float best) {
    const float r = g.h * ((float)reach - 0.002f - q.m);
    return best <= r * r * 0.9999f;
}

// NN kernel.  MODE 0: plain queries (sorted, w = original index) -> idx/d2 by original index
//             MODE 1: ICP iteration: the queries lg_track_kernel / lg_refine_kernel left flagged (flagS, by sorted
//                     position; state by ORIGINAL index through perm): match -> tg[i], no certificate (cur[i].w = 0)
//             MODE 2: fitness pass: query = st->fin * input, d2 by original index
// One CTA = NN_QPC consecutive queries.  A pass = region of blocks around the pass's queries -> TMA bulk copies of the
// blocks' points and offset tables into shared memory -> 2x2x2 search per thread -> warp-cooperative 4x4x4 search of the
// few queries the first stage could not prove.  A CTA whose region does not fit runs two passes over its halves; a pass
// that still does not fit reads the grid from global memory.  What neither stage proves goes to the box pyramid.
template <int MODE>
__global__ void __launch_bounds__(NN_THREADS, KSS_NN_CTAS)
lg_nn_kernel(Pyramid py, LgGridView gv, int n_q, float4* __restrict__ cur_s /* MODE 1: by original index */,
             const float4* __restrict__ inp_s, int* __restrict__ idx_out, float* __restrict__ d2out,
             float4* __restrict__ tg /* MODE 1: matches by original index */, const int* __restrict__ perm,
             unsigned char* __restrict__ flagS /* MODE 1: 1 = still to do */, LgState* __restrict__ st) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    float4* s_pts = reinterpret_cast<float4*>(dyn_smem);
    unsigned* s_tab = reinterpret_cast<unsigned*>(s_pts + NN_PTS_CAP);
    float4* s_lq = reinterpret_cast<float4*>(s_tab + NN_MAXOCC * NN_TAB_STRIDE);      // queries the first stage left over
    unsigned* s_rng = reinterpret_cast<unsigned*>(s_lq + NN_QPC);                     // per-thread candidate ranges of the first stage
    __shared__ __align__(8) unsigned long long s_bar;
    __shared__ LgGeom g;
    __shared__ float T[16];
    __shared__ int s_lo[3], s_hi[3];
    __shared__ signed char s_slot[NN_MAXBOX];    // box-local block index -> staged slot or -1 (touched blocks only)
    __shared__ unsigned s_mark[NN_MAXBOX / 32];  // touched blocks of the box
    __shared__ short s_list[NN_MAXREG];          // the touched blocks, compacted
    __shared__ int s_delta[NN_MAXOCC];
    __shared__ unsigned short s_pyr[2 * NN_QPC]; // s_lq entries past the second stage | past the third (pyramid)
    __shared__ float s_lub[NN_QPC];              // their bound (the previous match's distance), +inf if none
    __shared__ int s_path;                       // 0 staged, 1 global tables, 2 nothing in reach
    __shared__ unsigned s_p0[NN_MAXREG], s_cnt[NN_MAXREG], s_rank[NN_MAXREG];
    __shared__ int s_nleft, s_npyr, s_n2, s_nglob, s_nreg;

    if (MODE == 1) lg_wait_prior();
    if (MODE == 1 && st->done) return;
    if (MODE == 1 && st->n_unres <= (unsigned)n_q / LG_STAGED_DIV) return;       // few left: lg_left_kernel takes them one warp each
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < (int)(sizeof(LgGeom) / 4)) reinterpret_cast<int*>(&g)[tid] = reinterpret_cast<const int*>(gv.geom)[tid];
    if (MODE != 0 && tid < 16) T[tid] = MODE == 1 ? st->Tk[tid] : st->fin[tid];
    if (tid == 0) { s_nleft = 0; s_npyr = 0; s_n2 = 0; s_nglob = 0; mbar_init(&s_bar, 1); }
    __syncthreads();

    // ---- this thread's queries (coalesced), the pending transform, their position in the grid
    const int base = blockIdx.x * NN_QPC;
    float qx[NN_QPT], qy[NN_QPT], qz[NN_QPT]; unsigned qo[NN_QPT]; bool valid[NN_QPT]; QCell qc[NN_QPT];
    int box[NN_QPT];                             // cells to read, relative to the nearest corner (lg_box_pack)
    Best ub[NN_QPT];                             // MODE 1: the previous iteration's match is a target like any other --
                                                 // its distance bounds the search before a single cell is read
    unsigned long long key[NN_QPT];
    int left_at[NN_QPT];
#pragma unroll
    for (int k = 0; k < NN_QPT; ++k) {
        const int pos = base + k * NN_THREADS + tid;
        valid[k] = pos < n_q && (MODE != 1 || flagS[pos]);
        qx[k] = qy[k] = qz[k] = 0.0f; qo[k] = 0u;
        qc[k].inside = false; qc[k].cx = qc[k].cy = qc[k].cz = 0; qc[k].m = 1.0f;
        key[k] = 0xffffffffffffffffull; left_at[k] = -1;
        ub[k].d = __int_as_float(0x7f800000); ub[k].idx = 0xffffffffu; box[k] = 0;
        if (valid[k]) {
            const int oi = MODE == 1 ? perm[pos] : 0;
            const float4 tprev = MODE == 1 ? tg[oi] : make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
            const int pv = __float_as_int(tprev.w);                          // last iteration's match (-1: none yet)
            const float4 p = MODE == 2 ? inp_s[pos] : MODE == 1 ? cur_s[oi] : cur_s[pos];
            qo[k] = MODE == 1 ? (unsigned)oi : __float_as_uint(p.w);
            if (MODE == 2) xform_point(T, p.x, p.y, p.z, qx[k], qy[k], qz[k]);      // final * input (A.7)
            else { qx[k] = p.x; qy[k] = p.y; qz[k] = p.z; }                   // (MODE 1: lg_track_kernel applied T_k already)
            qc[k] = lg_qcell(g, qx[k], qy[k], qz[k]);
            if (pv >= 0) { ub[k].d = d2_rn(qx[k], qy[k], qz[k], tprev.x, tprev.y, tprev.z); ub[k].idx = (unsigned)pv; }
            // a bound beyond the reach of the cell-by-cell stages: nothing here can prove the answer, the query goes
            // straight to the seeded pyramid (and does not stretch the region this CTA stages)
            if (MODE == 1 && pv >= 0 && sqrtf(ub[k].d) * g.inv_h > cube3) qc[k].inside = false;
            // the cells a search has to read.  With a bound: the cells meeting the cube q +- sqrt(bound) -- every target that
            // can beat or tie the bound lies inside (0.003 cells of slack for the rounding of the binning), so scanning them
            // is exact without any further proof.  Without (or when the cube leaves the 4x4x4 cells around the nearest corner,
            // which is all a pass stages): the 2x2x2 cells around the corner, to be proven by lg_proven.
            box[k] = lg_box_pack(-1, 0, -1, 0, -1, 0, false);
            if (pv >= 0 && qc[k].inside) {
                const float rc = sqrtf(ub[k].d) * g.inv_h * 1.00001f + 0.003f;
                const float fx = (float)qc[k].cx + qc[k].dx, fy = (float)qc[k].cy + qc[k].dy, fz = (float)qc[k].cz + qc[k].dz;
                const int x0 = (int)floorf(fx - rc) - qc[k].cx, x1 = (int)floorf(fx + rc) - qc[k].cx;
                const int y0 = (int)floorf(fy - rc) - qc[k].cy, y1 = (int)floorf(fy + rc) - qc[k].cy;
                const int z0 = (int)floorf(fz - rc) - qc[k].cz, z1 = (int)floorf(fz + rc) - qc[k].cz;
                // at most 4 rows of at most 4 cells: at most 8 ranges (a row crosses at most one block boundary)
                if (x0 >= -2 && x1 <= 1 && y0 >= -2 && y1 <= 1 && z0 >= -2 && z1 <= 1 && (y1 - y0 + 1) * (z1 - z0 + 1) <= 4)
                    box[k] = lg_box_pack(x0, x1, y0, y1, z0, z1, true);
            }
        }
    }

    int npass = 1;
    unsigned phase = 0u;
    for (int pass = 0; pass < npass; ++pass) {
        // ---- (1) region of blocks around the pass's queries
        if (tid < 3) { s_lo[tid] = INT_MAX; s_hi[tid] = INT_MIN; }
        if (tid < NN_MAXBOX / 32) s_mark[tid] = 0u;
        __syncthreads();
        int lo3[3] = {INT_MAX, INT_MAX, INT_MAX}, hi3[3] = {INT_MIN, INT_MIN, INT_MIN};
        // the cells query k reads along axis a (0 = low end, 1 = high end), clamped to the grid; unbounded queries may
        // need the second stage: 4x4x4 cells
        auto cell_end = [&](int k, int a, int hi) {
            const int c = a == 0 ? qc[k].cx : a == 1 ? qc[k].cy : qc[k].cz;
            const int o = lg_box_bounded(box[k]) ? lg_box_get(box[k], 2 * a + hi) : (hi ? 1 : -2);
            return hi ? min(c + o, g.nf[a] - 1) : max(c + o, 0);
        };
#pragma unroll
        for (int k = 0; k < NN_QPT; ++k) {
            if (!(valid[k] && qc[k].inside && (npass == 1 || k == pass))) continue;
#pragma unroll
            for (int a = 0; a < 3; ++a) { lo3[a] = min(lo3[a], cell_end(k, a, 0)); hi3[a] = max(hi3[a], cell_end(k, a, 1)); }
        }
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int l = __reduce_min_sync(KSS_FULL, lo3[a]), h = __reduce_max_sync(KSS_FULL, hi3[a]);
            if (lane == 0 && l <= h) { atomicMin(&s_lo[a], l); atomicMax(&s_hi[a], h); }
        }
        __syncthreads();
        const bool any = s_lo[0] <= s_hi[0];
        const int bx0 = any ? s_lo[0] >> 2 : 0, by0 = any ? s_lo[1] >> 2 : 0, bz0 = any ? s_lo[2] >> 2 : 0;
        const int rx = any ? (s_hi[0] >> 2) - bx0 + 1 : 0, ry = any ? (s_hi[1] >> 2) - by0 + 1 : 0, rz = any ? (s_hi[2] >> 2) - bz0 + 1 : 0;
        const long long nbox_ll = (long long)rx * ry * rz;
        const bool boxed = any && nbox_ll <= NN_MAXBOX;
        // the blocks the pass's queries can touch: bitmap over the box, then a compact list -- a Z-order range of queries
        // is not a box, its union of blocks is what gets staged
        if (boxed) {
#pragma unroll
            for (int k = 0; k < NN_QPT; ++k) {
                if (!(valid[k] && qc[k].inside && (npass == 1 || k == pass))) continue;
                const int xb = cell_end(k, 0, 1) >> 2, yb = cell_end(k, 1, 1) >> 2, zb = cell_end(k, 2, 1) >> 2;
                for (int bz = cell_end(k, 2, 0) >> 2; bz <= zb; ++bz)
                    for (int by = cell_end(k, 1, 0) >> 2; by <= yb; ++by)
                        for (int bx = cell_end(k, 0, 0) >> 2; bx <= xb; ++bx) {
                            const int li = ((bz - bz0) * ry + (by - by0)) * rx + (bx - bx0);
                            if (!((s_mark[li >> 5] >> (li & 31)) & 1u)) atomicOr(&s_mark[li >> 5], 1u << (li & 31));
                        }
            }
        }
        __syncthreads();
        if (warp == 0) {
            static_assert(NN_MAXBOX == 2048, "two bitmap words per lane");
            unsigned w0 = s_mark[2 * lane], w1 = s_mark[2 * lane + 1];
            const unsigned c = (unsigned)(__popc(w0) + __popc(w1));
            unsigned x = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(KSS_FULL, x, o); if (lane >= o) x += y; }
            const unsigned total = __shfl_sync(KSS_FULL, x, 31);
            unsigned e = x - c;
            const int nr0 = (boxed && total <= (unsigned)NN_MAXREG) ? (int)total : 0;
            if (nr0) {
                while (w0) { const int b = __ffs(w0) - 1; w0 &= w0 - 1u; s_list[e++] = (short)(64 * lane + b); }
                while (w1) { const int b = __ffs(w1) - 1; w1 &= w1 - 1u; s_list[e++] = (short)(64 * lane + 32 + b); }
            }
            if (lane == 0) s_nreg = nr0;
        }
        __syncthreads();
        const int nreg = s_nreg;
        if (tid < nreg) {                                        // ranks and point ranges of the touched blocks, one thread each
            const int li = s_list[tid];
            const unsigned r = __ldg(gv.blk_rank + lg_code(gv.lut, bx0 + li % rx, by0 + (li / rx) % ry, bz0 + li / (rx * ry)));
            unsigned q0 = 0u, cnt = 0u;
            if (r != LG_EMPTY) { q0 = __ldg(gv.fine_start + (size_t)r * 64); cnt = __ldg(gv.fine_start + (size_t)r * 64 + 64) - q0; }
            s_rank[tid] = r; s_p0[tid] = q0; s_cnt[tid] = cnt;
        }
        __syncthreads();
        if (warp == 0) {
            // slots and shared-memory offsets (prefix sums, EPL consecutive entries per lane), then the TMA bulk copies
            constexpr int EPL = NN_MAXREG / 32;
            unsigned cn[EPL], p0[EPL], rk[EPL], ps = 0u, os = 0u;
#pragma unroll
            for (int j = 0; j < EPL; ++j) {
                const int en = lane * EPL + j;
                cn[j] = en < nreg ? s_cnt[en] : 0u; p0[j] = en < nreg ? s_p0[en] : 0u; rk[j] = en < nreg ? s_rank[en] : LG_EMPTY;
                ps += cn[j]; os += cn[j] ? 1u : 0u;
            }
            unsigned px = ps, ox = os;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned yp = __shfl_up_sync(KSS_FULL, px, o), yo = __shfl_up_sync(KSS_FULL, ox, o);
                if (lane >= o) { px += yp; ox += yo; }
            }
            const unsigned tp_total = __shfl_sync(KSS_FULL, px, 31), nocc = __shfl_sync(KSS_FULL, ox, 31);
            int path = 0;
            if (!any) path = 2;
            else if (nreg == 0 || tp_total > (unsigned)NN_PTS_CAP || nocc > (unsigned)NN_MAXOCC) {
                path = 1;
                if (lane == 0 && npass > 1) {
                    atomicAdd(&lg_dbg[!boxed ? 0 : nreg == 0 ? 1 : tp_total > (unsigned)NN_PTS_CAP ? 2 : 3], 1u);
                    atomicAdd(&lg_dbg[4], tp_total); atomicAdd(&lg_dbg[5], nocc); atomicAdd(&lg_dbg[6], (unsigned)nreg); atomicAdd(&lg_dbg[7], (unsigned)min(nbox_ll, 100000ll));
                }
            }
            if (path == 0 && tp_total > 0u) {
                if (lane == 0) mbar_arrive_expect_tx(&s_bar, tp_total * 16u + nocc * (unsigned)(NN_TAB_STRIDE * 4));
                __syncwarp();
                unsigned poff = px - ps, slot = ox - os;
#pragma unroll
                for (int j = 0; j < EPL; ++j) {
                    const int en = lane * EPL + j;
                    if (en < nreg) s_slot[s_list[en]] = cn[j] ? (signed char)slot : (signed char)-1;
                    if (cn[j]) {
                        s_delta[slot] = (int)poff - (int)p0[j];
                        // TMA bulk copies: the block's points (contiguous in the sorted array) and its offset table
                        bulk_g2s(s_pts + poff, gv.tp + p0[j], cn[j] * 16u, &s_bar);
                        bulk_g2s(s_tab + slot * NN_TAB_STRIDE, gv.fine_start + (size_t)rk[j] * 64, NN_TAB_STRIDE * 4, &s_bar);
                        poff += cn[j]; ++slot;
                    }
                }
            } else if (path == 0) path = 2;                     // the region holds no target at all
            if (lane == 0) s_path = path;
        }
        __syncthreads();
        const int path = s_path;
        if (path == 1 && npass == 1 && NN_QPT > 1) { npass = NN_QPT; pass = -1; continue; }     // retry in halves (CTA-uniform)
        if (path == 0) { mbar_wait(&s_bar, phase & 1u); ++phase; }
        if (path == 1 && tid == 0) ++s_nglob;

        // ---- (2) first stage, one thread per query: the 2x2x2 cells around the nearest cell corner.  Cells farther from
        //      the query than the bound it already holds are skipped (their targets are strictly farther); the ranges of
        //      the rest go to a per-thread list and ONE flat loop walks them, so that a warp iterates max-over-lanes of the
        //      candidate TOTALS instead of the sum over cells of per-cell maxima
        SmemAcc sacc{s_pts, s_tab, s_slot, s_delta, bx0, by0, bz0, rx, ry};
        GlobAcc gacc{gv.blk_rank, gv.fine_start, gv.tp, gv.lut};
        unsigned* rng = s_rng + tid * NN_RNG_STRIDE;
#pragma unroll
        for (int k = 0; k < NN_QPT; ++k) {
            const bool mine = valid[k] && (npass == 1 || k == pass);
            const bool search = mine && qc[k].inside && path != 2;
            const bool staged4 = search && !lg_box_bounded(box[k]);                 // its 4x4x4 cells are part of the region
            Best b = ub[k];
            if (path == 0) {
                unsigned total = 0u; int nr = 0;
                bool overflow = false;                                               // more ranges than the list holds: not proven
                if (search) {
                    const float h2 = g.h * g.h;
                    const float fx = (float)qc[k].cx + qc[k].dx, fy = (float)qc[k].cy + qc[k].dy, fz = (float)qc[k].cz + qc[k].dz;
                    const int bxl = qc[k].cx + lg_box_get(box[k], 0), bxh = qc[k].cx + lg_box_get(box[k], 1);
                    const int byl = max(qc[k].cy + lg_box_get(box[k], 2), 0), byh = min(qc[k].cy + lg_box_get(box[k], 3), g.nf[1] - 1);
                    const int bzl = max(qc[k].cz + lg_box_get(box[k], 4), 0), bzh = min(qc[k].cz + lg_box_get(box[k], 5), g.nf[2] - 1);
                    // gap of the query to a cell along one axis, in cells, 0.002 short (rounding of the binning): a lower
                    // bound of the distance to everything in the cell
                    auto gap1 = [](float f, int i) { return fmaxf(fmaxf((float)i - f, f - (float)(i + 1)) - 0.002f, 0.0f); };
                    for (int iz = bzl; iz <= bzh; ++iz) {
                        const float gz = gap1(fz, iz);
                        for (int iy = byl; iy <= byh; ++iy) {
                            const float gy = gap1(fy, iy);
                            const float lb = (gy * gy + gz * gz) * h2;
                            if (lb > b.d) continue;                                      // everything in this row is strictly farther
                            int xa = bxl, xb = bxh;
                            while (xa < xb && lb + gap1(fx, xa) * gap1(fx, xa) * h2 > b.d) ++xa;
                            while (xb > xa && lb + gap1(fx, xb) * gap1(fx, xb) * h2 > b.d) --xb;
                            xa = max(xa, 0); xb = min(xb, g.nf[0] - 1);
                            const int f = ((iz & 3) << 4) | ((iy & 3) << 2);
                            for (int ix = xa; ix <= xb;) {
                                const int bx = ix >> 2, xe = min(xb, (bx << 2) | 3);
                                const int sl = s_slot[(((iz >> 2) - bz0) * ry + ((iy >> 2) - by0)) * rx + (bx - bx0)];
                                if (sl >= 0) {
                                    const unsigned* t = s_tab + sl * NN_TAB_STRIDE;
                                    const unsigned st0 = t[f + (ix & 3)], en = t[f + (xe & 3) + 1];
                                    if (en > st0 && nr < NN_RNG_STRIDE - 1) { rng[nr++] = (st0 + (unsigned)s_delta[sl]) | ((en + (unsigned)s_delta[sl]) << 16); total += en - st0; }
                                    else if (en > st0) overflow = true;
                                }
                                ix = xe + 1;
                            }
                        }
                    }
                }
                rng[nr] = 0u;
                const unsigned maxtot = __reduce_max_sync(KSS_FULL, total);
#ifdef KSS_LG_COUNT
                { const unsigned sum = __reduce_add_sync(KSS_FULL, total); if (lane == 0) { atomicAdd(&lg_cnt[0], (unsigned long long)sum); atomicAdd(&lg_cnt[1], (unsigned long long)maxtot * 32u); atomicAdd(&lg_cnt[2], 32ull); } }
#endif
                unsigned cur = rng[0], j = cur & 0xffffu, e = cur >> 16; int r = 0;
                unsigned long long bk = ((unsigned long long)__float_as_uint(b.d) << 32) | b.idx;
                for (unsigned it = 0; it < maxtot; ++it) {
                    if (it < total) {
                        if (j == e) { ++r; cur = rng[r]; j = cur & 0xffffu; e = cur >> 16; }
                        const float4 c = s_pts[j]; ++j;
                        const unsigned long long ck = ((unsigned long long)__float_as_uint(d2_rn(qx[k], qy[k], qz[k], c.x, c.y, c.z)) << 32) | __float_as_uint(c.w);
                        bk = ck < bk ? ck : bk;                                         // d2 >= 0: float bits order as integers; ties -> lowest index
                    }
                }
                b.d = __uint_as_float((unsigned)(bk >> 32)); b.idx = (unsigned)bk;
                if (overflow) box[k] &= ~(1 << 12), b.d = __int_as_float(0x7f800000);   // (the second stage starts afresh)
            } else if (search) {
                const int x0 = max(qc[k].cx - 1, 0), x1 = min(qc[k].cx, g.nf[0] - 1);
                const int y0 = max(qc[k].cy - 1, 0), y1 = min(qc[k].cy, g.nf[1] - 1);
                const int z0 = max(qc[k].cz - 1, 0), z1 = min(qc[k].cz, g.nf[2] - 1);
                lg_scan_cells(gacc, x0, x1, y0, y1, z0, z1, qx[k], qy[k], qz[k], b);
            }
            if (!mine) continue;
            const bool ok = search && ((path == 0 && lg_box_bounded(box[k])) || lg_proven(g, qc[k], 1, b.d));
            if (ok) key[k] = ((unsigned long long)__float_as_uint(b.d) << 32) | b.idx;
            else {
                left_at[k] = atomicAdd(&s_nleft, 1);
                // (an unbounded query with nothing at all in its 2x2x2 cells is far from the target: the 4x4x4 stage would
                //  most likely come back empty-handed too -- to the pyramid)
                const bool try4 = staged4 && (b.d < __int_as_float(0x7f800000) || ub[k].d < __int_as_float(0x7f800000));
                s_lq[left_at[k]] = make_float4(qx[k], qy[k], qz[k], __int_as_float(try4 ? 1 : 0));
                s_lub[left_at[k]] = ub[k].d;
            }
        }
        __syncthreads();
        // ---- (3) second stage, one WARP per left-over query of this pass: lane = one (y, z) row of half the 4x4x4 cells
        const int nleft = s_nleft;
        const int first = (npass == 1 || pass == 0) ? 0 : s_n2;                   // entries queued by this pass
        for (int li = first + warp; li < nleft; li += NN_THREADS / 32) {
            const float4 q = s_lq[li];
            bool ok = false;
            unsigned long long kk = 0xffffffffffffffffull;
            if (__float_as_int(q.w) == 1) {
                const QCell c = lg_qcell(g, q.x, q.y, q.z);
                const int iy = c.cy - 2 + (lane & 3), iz = c.cz - 2 + ((lane >> 2) & 3);
                const int xa = c.cx - 2 + 2 * (lane >> 4);
                Best b; b.d = __int_as_float(0x7f800000); b.idx = 0xffffffffu;
                if (iy >= 0 && iy < g.nf[1] && iz >= 0 && iz < g.nf[2]) {
                    const int x0 = max(xa, 0), x1 = min(xa + 1, g.nf[0] - 1);
                    if (path == 0) lg_scan_cells(sacc, x0, x1, iy, iy, iz, iz, q.x, q.y, q.z, b);
                    else lg_scan_cells(gacc, x0, x1, iy, iy, iz, iz, q.x, q.y, q.z, b);
                }
                kk = ((unsigned long long)__float_as_uint(b.d) << 32) | b.idx;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) { const unsigned long long o = __shfl_xor_sync(KSS_FULL, kk, off); kk = o < kk ? o : kk; }
                ok = lg_proven(g, c, 2, __uint_as_float((unsigned)(kk >> 32)));
            }
            if (lane == 0) {
                if (ok) { s_lq[li].x = __uint_as_float((unsigned)(kk >> 32)); s_lq[li].y = __uint_as_float((unsigned)kk); s_lq[li].w = __int_as_float(2); }
                else s_pyr[atomicAdd(&s_npyr, 1)] = (unsigned short)li;
            }
        }
        __syncthreads();
        if (tid == 0) s_n2 = nleft;
        // (the next pass's first barrier orders this write and the reuse of the staging area)
    }

    // ---- (4) left-overs that carry a bound (the previous match): one WARP per query reads, from the grid in global memory,
    //      the cells meeting the cube q +- sqrt(bound), lane = (y, z) row -- exact without proof, like the bounded first
    //      stage, for isolated points whose neighbour is many cells away.  What has no bound, or a cube of more than
    //      17 x 17 rows, is left to the box pyramid.
    {
        const int n3 = s_npyr;
        __syncthreads();
        if (tid == 0) s_npyr = 0;
        __syncthreads();
        GlobAcc gacc{gv.blk_rank, gv.fine_start, gv.tp, gv.lut};
        for (int e3 = warp; e3 < n3; e3 += NN_THREADS / 32) {
            const int li = s_pyr[e3];
            const float4 q = s_lq[li];
            const float ubd = s_lub[li];
            const float rc = sqrtf(ubd) * g.inv_h * 1.00001f + 0.003f;
            bool done3 = false;
            if (rc <= cube3) {
                const float fx = (q.x - g.lo[0]) * g.inv_h, fy = (q.y - g.lo[1]) * g.inv_h, fz = (q.z - g.lo[2]) * g.inv_h;
                // (a query with a bound this small is within 8 cells of a target, hence of the grid: no overflow)
                const int xl = max((int)floorf(fx - rc), 0), xh = min((int)floorf(fx + rc), g.nf[0] - 1);
                const int yl = max((int)floorf(fy - rc), 0), yh = min((int)floorf(fy + rc), g.nf[1] - 1);
                const int zl = max((int)floorf(fz - rc), 0), zh = min((int)floorf(fz + rc), g.nf[2] - 1);
                const int ny = yh - yl + 1, nrows = ny * (zh - zl + 1);
                auto gap1 = [](float f, int i) { return fmaxf(fmaxf((float)i - f, f - (float)(i + 1)) - 0.002f, 0.0f); };
                const float h2 = g.h * g.h;
                Best b; b.d = __int_as_float(0x7f800000); b.idx = 0xffffffffu;
                for (int rw = lane; rw < nrows; rw += 32) {
                    const int iy = yl + rw % ny, iz = zl + rw / ny;
                    const float gy = gap1(fy, iy), gz = gap1(fz, iz);
                    const float lb = (gy * gy + gz * gz) * h2;
                    if (lb > ubd) continue;
                    int xa = xl, xb = xh;
                    while (xa < xb && lb + gap1(fx, xa) * gap1(fx, xa) * h2 > ubd) ++xa;
                    while (xb > xa && lb + gap1(fx, xb) * gap1(fx, xb) * h2 > ubd) --xb;
                    lg_scan_cells(gacc, xa, xb, iy, iy, iz, iz, q.x, q.y, q.z, b);
                }
                unsigned long long kk = ((unsigned long long)__float_as_uint(b.d) << 32) | b.idx;
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) { const unsigned long long o = __shfl_xor_sync(KSS_FULL, kk, off); kk = o < kk ? o : kk; }
                done3 = nrows > 0 && (unsigned)(kk >> 32) <= __float_as_uint(ubd);       // (the previous match itself is in the cube)
                if (done3 && lane == 0) { s_lq[li].x = __uint_as_float((unsigned)(kk >> 32)); s_lq[li].y = __uint_as_float((unsigned)kk); s_lq[li].w = __int_as_float(2); }
            }
            if (!done3 && lane == 0) s_pyr[NN_QPC + atomicAdd(&s_npyr, 1)] = (unsigned short)li;
            if (lane == 0) atomicAdd(&lg_dbg2[done3 ? 3 : !(ubd < 3.0e38f) ? 0 : rc > cube3 ? 1 : 2], 1u);
        }
        __syncthreads();
    }
    // ---- (5) whatever is left: the box pyramid, 32 queries per warp
    const int npyr = s_npyr;
    if (npyr > 0) {
        float4* slots = s_pts + warp * TILE;                     // the staged points are no longer needed
        for (int b0 = warp * 32; b0 < npyr; b0 += NN_THREADS) {
            const int li = s_pyr[NN_QPC + min(b0 + lane, npyr - 1)];      // idle lanes repeat the last query
            const float4 q = s_lq[li];
            const unsigned long long kk = lg_warp_nn(py, q.x, q.y, q.z, slots, s_lub[li]);
            __syncwarp();
            if (b0 + lane < npyr) { s_lq[li].x = __uint_as_float((unsigned)(kk >> 32)); s_lq[li].y = __uint_as_float((unsigned)kk); s_lq[li].w = __int_as_float(2); }
        }
    }
    __syncthreads();
    if (MODE == 1 && tid == 0) {
        if (npyr) atomicAdd(&st->miss[2], (unsigned)npyr);
        if (s_nglob) atomicAdd(&st->miss[3], (unsigned)s_nglob);
        if (npass > 1) atomicAdd(&st->miss[4], 1u);
    }

    // ---- (6) results by ORIGINAL index
#pragma unroll
    for (int k = 0; k < NN_QPT; ++k) {
        if (!valid[k]) continue;
        if (left_at[k] >= 0) {
            const float4 r = s_lq[left_at[k]];
            key[k] = ((unsigned long long)__float_as_uint(r.x) << 32) | __float_as_uint(r.y);
        }
        const float d2 = __uint_as_float((unsigned)(key[k] >> 32));
        const unsigned ti = (unsigned)(key[k] & 0xffffffffu);
        if (MODE == 1) {
            const float4 t = __ldg(gv.t_orig + ti);
            tg[qo[k]] = make_float4(t.x, t.y, t.z, __int_as_float((int)ti));
            cur_s[qo[k]] = make_float4(qx[k], qy[k], qz[k], 0.0f);                          // no certificate from this kernel
            flagS[base + k * NN_THREADS + tid] = 0;
        } else {
            d2out[qo[k]] = d2;
            if (MODE == 0) idx_out[qo[k]] = (int)ti;
        }
    }
}

// ---- certificates.  Per source point (by ORIGINAL index) the run keeps cur[i] = {current position, lb} and
// tg[i] = {matched target, its index}: lb is a strict lower bound of the distance from the current position to every
// target OTHER than the match.  lg_track_kernel moves the point by T_k, lowers lb by the length of the move and keeps the
// match whenever its distance is still below lb (then it is the unique nearest neighbour: nothing to search, the pass
// streams).  The search kernels below establish lb = min(distance to the second-nearest target scanned, distance to the
// unscanned region).  Both bounds are kept 1e-5 (relative) apart from the values they are compared with, four orders of
// magnitude above the fp32 rounding of d2_rn, so a kept match is also the search's answer bit for bit (no tie possible).
constexpr float LG_CERT_UP = 1.00001f, LG_CERT_DOWN = 0.99999f;

// trackers of a scan.  Best2: best (d2 bits << 32 | original index) and the second-smallest d2 among the scanned targets;
// Best3: the two best keys and the third-smallest d2 (for points whose two nearest targets are about equally far: the
// certificate then names both and bounds the third)
constexpr unsigned long long LG_KEY_NONE = (0x7f800000ull << 32) | 0xffffffffull;       // d2 = +inf
struct Best2 {
    unsigned long long key; float second;
    __device__ __forceinline__ void init() { key = 0xffffffffffffffffull; second = __int_as_float(0x7f800000); }
    __device__ __forceinline__ void update(unsigned long long ck) {
        const unsigned long long hi = ck < key ? key : ck;              // the loser of (candidate, best) bounds the second
        key = ck < key ? ck : key;
        second = fminf(second, __uint_as_float((unsigned)(hi >> 32)));
    }
    __device__ __forceinline__ float thr() const { return second; }
};
struct Best3 {
    unsigned long long k1, k2; float third;
    __device__ __forceinline__ void init() { k1 = k2 = LG_KEY_NONE; third = __int_as_float(0x7f800000); }
    __device__ __forceinline__ void update(unsigned long long ck) {
        const unsigned long long h1 = ck < k1 ? k1 : ck;
        k1 = ck < k1 ? ck : k1;
        const unsigned long long h2 = h1 < k2 ? k2 : h1;
        k2 = h1 < k2 ? h1 : k2;
        third = fminf(third, __uint_as_float((unsigned)(h2 >> 32)));
    }
    __device__ __forceinline__ float thr() const { return third; }
};

template <class Acc, class Trk>
__device__ __forceinline__ void lg_scan_cells2(const Acc& acc, int x0, int x1, int iy, int iz, float qx, float qy, float qz, Trk& b) {
    const int f = ((iz & 3) << 4) | ((iy & 3) << 2);
    for (int ix = x0; ix <= x1;) {
        const int bx = ix >> 2;
        const int xe = min(x1, (bx << 2) | 3);                        // cells of one block along x are contiguous
        const unsigned* t; const float4* p;
        if (acc.block(bx, iy >> 2, iz >> 2, t, p)) {
            const unsigned s = t[f + (ix & 3)], e = t[f + (xe & 3) + 1];
#pragma unroll SCAN_UNROLL
            for (unsigned j = s; j < e; ++j) {
                const float4 c = p[j];
                b.update(((unsigned long long)__float_as_uint(d2_rn(qx, qy, qz, c.x, c.y, c.z)) << 32) | __float_as_uint(c.w));
            }
        }
        ix = xe + 1;
    }
}

// the cells meeting the cube q +- rho, rows and row ends farther than min(second so far, rho) skipped: exact best, and
// every target that is not the best is at least min(sqrt(second), rho) away (gaps are 0.002 cells short, the cube 0.003
// cells wide of the rounding of the binning).  rows_per_thread: lanes of a warp share one query when nl > 1.
template <class Acc, class Trk>
__device__ __forceinline__ void lg_cube_search(const Acc& acc, const LgGeom& g, float x, float y, float z, float rho, int lane, int nl, Trk& b) {
    const float rc = rho * g.inv_h * 1.00001f + 0.003f;
    const float fx = (x - g.lo[0]) * g.inv_h, fy = (y - g.lo[1]) * g.inv_h, fz = (z - g.lo[2]) * g.inv_h;
    const int xl = max((int)floorf(fx - rc), 0), xh = min((int)floorf(fx + rc), g.nf[0] - 1);
    const int yl = max((int)floorf(fy - rc), 0), yh = min((int)floorf(fy + rc), g.nf[1] - 1);
    const int zl = max((int)floorf(fz - rc), 0), zh = min((int)floorf(fz + rc), g.nf[2] - 1);
    const int ny = yh - yl + 1, nrows = ny * (zh - zl + 1);
    auto gap1 = [](float f, int i) { return fmaxf(fmaxf((float)i - f, f - (float)(i + 1)) - 0.002f, 0.0f); };
    const float h2 = g.h * g.h, rho2 = rho * rho;
    // the query's own row first: it usually holds the two nearest targets, whose distances then prune the other rows
    const int r0 = (min(max((int)floorf(fz), zl), zh) - zl) * ny + (min(max((int)floorf(fy), yl), yh) - yl);
    for (int k = lane; k < nrows; k += nl) {
        const int rw = k == 0 ? r0 : (k <= r0 ? k - 1 : k);
        const int iy = yl + rw % ny, iz = zl + rw / ny;
        const float gy = gap1(fy, iy), gz = gap1(fz, iz);
        const float lb = (gy * gy + gz * gz) * h2;
        const float thr = fminf(b.thr(), rho2);
        if (lb > thr) continue;                                      // everything in this row is strictly farther
        int xa = xl, xb = xh;
        while (xa < xb && lb + gap1(fx, xa) * gap1(fx, xa) * h2 > thr) ++xa;
        while (xb > xa && lb + gap1(fx, xb) * gap1(fx, xb) * h2 > thr) --xb;
        lg_scan_cells2(acc, xa, xb, iy, iz, x, y, z, b);
    }
}

// ---- neighbour lists of the target (built once per pair): knn[p * LG_KSLOTS] = {R, count, -, -}, then LG_KNN targets
// {x, y, z, bits(index)} (pads far away) such that EVERY target closer to p than R is listed.  A query q matched to p
// then has every target within R - |q p| of itself in {p} + list(p): one 128-byte line replaces a walk over grid cells.
#ifndef KSS_LG_KSLOTS
#define KSS_LG_KSLOTS 8
#endif
constexpr int LG_KSLOTS = KSS_LG_KSLOTS;         // float4 slots per target (8 = one 128-byte line)
constexpr int LG_KNN = LG_KSLOTS - 1;

__global__ void __launch_bounds__(128)
lg_knn_kernel(LgGridView gv, float4* __restrict__ knn) {
    __shared__ LgGeom g;
    if (threadIdx.x < (int)(sizeof(LgGeom) / 4)) reinterpret_cast<int*>(&g)[threadIdx.x] = reinterpret_cast<const int*>(gv.geom)[threadIdx.x];
    __syncthreads();
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= gv.n_t) return;
    const float4 p = gv.tp[j];                                    // sorted order: neighbouring threads share cells
    const unsigned self = __float_as_uint(p.w);
    const int cx = lg_fine(p.x, g.lo[0], g.inv_h, g.nf[0]), cy = lg_fine(p.y, g.lo[1], g.inv_h, g.nf[1]), cz = lg_fine(p.z, g.lo[2], g.inv_h, g.nf[2]);
    const float fx = (p.x - g.lo[0]) * g.inv_h, fy = (p.y - g.lo[1]) * g.inv_h, fz = (p.z - g.lo[2]) * g.inv_h;
    GlobAcc acc{gv.blk_rank, gv.fine_start, gv.tp, gv.lut};
    float dd[LG_KNN]; float4 ee[LG_KNN];
    int cnt = 0; float maxd = 0.0f; int maxj = 0; float cover = 0.0f;
    for (int r = 1; r <= 2; ++r) {
        cnt = 0; maxd = 0.0f; maxj = 0;
#pragma unroll
        for (int k = 0; k < LG_KNN; ++k) { dd[k] = __int_as_float(0x7f800000); ee[k] = make_float4(PAD_COORD, PAD_COORD, PAD_COORD, __int_as_float(0x7fffffff)); }
        const int x0 = max(cx - r, 0), x1 = min(cx + r, g.nf[0] - 1);
        const int y0 = max(cy - r, 0), y1 = min(cy + r, g.nf[1] - 1);
        const int z0 = max(cz - r, 0), z1 = min(cz + r, g.nf[2] - 1);
        for (int iz = z0; iz <= z1; ++iz)
            for (int iy = y0; iy <= y1; ++iy) {
                const int f = ((iz & 3) << 4) | ((iy & 3) << 2);
                for (int ix = x0; ix <= x1;) {
                    const int bx = ix >> 2, xe = min(x1, (bx << 2) | 3);
                    const unsigned* t; const float4* pts;
                    if (acc.block(bx, iy >> 2, iz >> 2, t, pts)) {
                        const unsigned s0 = t[f + (ix & 3)], e0 = t[f + (xe & 3) + 1];
                        for (unsigned q = s0; q < e0; ++q) {
                            const float4 c = pts[q];
                            if (__float_as_uint(c.w) == self) continue;
                            const float d = d2_rn(p.x, p.y, p.z, c.x, c.y, c.z);
                            if (cnt < LG_KNN) {                      // fill
#pragma unroll
                                for (int k = 0; k < LG_KNN; ++k) if (k == cnt) { dd[k] = d; ee[k] = c; }
                                ++cnt;
                                if (cnt == LG_KNN) {
                                    maxd = dd[0]; maxj = 0;
#pragma unroll
                                    for (int k = 1; k < LG_KNN; ++k) if (dd[k] > maxd) { maxd = dd[k]; maxj = k; }
                                }
                            } else if (d < maxd) {                   // replace the farthest
#pragma unroll
                                for (int k = 0; k < LG_KNN; ++k) if (k == maxj) { dd[k] = d; ee[k] = c; }
                                maxd = dd[0]; maxj = 0;
#pragma unroll
                                for (int k = 1; k < LG_KNN; ++k) if (dd[k] > maxd) { maxd = dd[k]; maxj = k; }
                            }
                        }
                    }
                    ix = xe + 1;
                }
            }
        // every target outside the scanned cells is at least `cover` away (no targets beyond the grid; the binning is
        // exact to < 1e-4 cells, 0.002 allowed)
        float cv = __int_as_float(0x7f800000);
        if (cx - r > 0) cv = fminf(cv, fx - (float)(cx - r));
        if (cx + r < g.nf[0] - 1) cv = fminf(cv, (float)(cx + r + 1) - fx);
        if (cy - r > 0) cv = fminf(cv, fy - (float)(cy - r));
        if (cy + r < g.nf[1] - 1) cv = fminf(cv, (float)(cy + r + 1) - fy);
        if (cz - r > 0) cv = fminf(cv, fz - (float)(cz - r));
        if (cz + r < g.nf[2] - 1) cv = fminf(cv, (float)(cz + r + 1) - fz);
        cover = fmaxf(cv - 0.002f, 0.0f) * g.h;
        if (cnt == LG_KNN) break;
    }
    // full list: everything closer than its farthest entry is listed (equal distances may be missing: strict bound)
    float R = cover;
    if (cnt == LG_KNN) R = fminf(R, sqrtf(maxd));
    R = fminf(R, 1.0e30f) * LG_CERT_DOWN;
    float4* out = knn + (size_t)self * LG_KSLOTS;
    out[0] = make_float4(R, (float)cnt, 0.0f, 0.0f);
#pragma unroll
    for (int k = 0; k < LG_KNN; ++k) out[1 + k] = ee[k];
}


// One WARP, one query (lane = row of cells): exact best over the cube q +- rho, written with the strongest certificate
// the scan supports -- the runner-up by name when it is inside the cube (cur.w = -lb3, tg2 = runner-up: lg_track_kernel
// then compares the two every iteration and needs a search only when the THIRD target comes close).  False if the cube
// holds nothing (the caller falls back).
__device__ __forceinline__ bool lg_warp_search_store(const GlobAcc& gacc, const LgGeom& g, const float4* __restrict__ t_orig, const float4& q,
                                                     const float4& t0, float rho, int lane, int i, float4* __restrict__ cur,
                                                     float4* __restrict__ tg, float4* __restrict__ tg2, float4* __restrict__ cert) {
    Best3 b; b.init();
    lg_cube_search(gacc, g, q.x, q.y, q.z, rho, lane, 32, b);
    unsigned long long K1 = b.k1;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) { const unsigned long long o = __shfl_xor_sync(KSS_FULL, K1, off); K1 = o < K1 ? o : K1; }
    if (K1 == LG_KEY_NONE) return false;
    unsigned long long K2 = b.k1 == K1 ? b.k2 : b.k1;                 // (the lanes scan disjoint rows: K1 lives in one lane)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) { const unsigned long long o = __shfl_xor_sync(KSS_FULL, K2, off); K2 = o < K2 ? o : K2; }
    float rest = b.k1 == K1 ? (b.k2 == K2 ? b.third : __uint_as_float((unsigned)(b.k2 >> 32)))
               : b.k1 == K2 ? __uint_as_float((unsigned)(b.k2 >> 32)) : __uint_as_float((unsigned)(b.k1 >> 32));
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) rest = fminf(rest, __shfl_xor_sync(KSS_FULL, rest, off));
    if (lane == 0) {
        const float rho2 = rho * rho;
        const unsigned ti = (unsigned)K1;
        const float4 t = ti == (unsigned)__float_as_int(t0.w) ? t0 : __ldg(t_orig + ti);
        tg[i] = make_float4(t.x, t.y, t.z, __int_as_float((int)ti));
        const float d2nd = __uint_as_float((unsigned)(K2 >> 32));
        if (d2nd < rho2) {                                            // the runner-up is known by name
            const unsigned t2i = (unsigned)K2;
            const float4 t2 = __ldg(t_orig + t2i);
            tg2[i] = make_float4(t2.x, t2.y, t2.z, __int_as_float((int)t2i));
            cur[i] = cert[i] = make_float4(q.x, q.y, q.z, -(sqrtf(fminf(rest, rho2)) * LG_CERT_DOWN));
        } else cur[i] = cert[i] = make_float4(q.x, q.y, q.z, sqrtf(rho2) * LG_CERT_DOWN);
    }
    return true;
}

// The flagged queries (by sorted position, so that a warp's searches share cache lines), one thread each: a query that
// carries a match holds a bound before a single cell is read.  It scans the cube q +- (bound + margin) from the grid in
// global memory (L2-resident: points 16 B, tables, ranks) when that is at most 3 x 3 rows of cells -- exact without
// proof, and the margin is what the certificate lives on.  The margin is dropped while T_k still moves the point
// farther than a certificate could survive.  Everything else is counted and queued for the general kernels.
// the flagged queries of a segment, G lanes each (G = 1: one thread per query)
template <int G>
__device__ __forceinline__ void lg_refine_queries(const GlobAcc& gacc, const LgGeom& g, const LgGridView& gv, const float* T, float m, bool walk,
                                                  const int* s_list, int nl, float4* __restrict__ cur, float4* __restrict__ tg, float4* __restrict__ cert,
                                                  const int* __restrict__ perm, unsigned char* __restrict__ flagS, int* __restrict__ worklist,
                                                  LgState* __restrict__ st) {
    const int tid = threadIdx.x, lane = tid & 31, sub = tid % G;
    constexpr int QP = 256 / G;                  // queries per pass of the CTA
    const unsigned gmask = G == 1 ? 0u : (G == 32 ? KSS_FULL : ((1u << G) - 1u) << (lane & ~(G - 1)));
    for (int e0 = 0; e0 < nl; e0 += QP) {
        const int e = e0 + tid / G;
        const bool valid = e < nl;
        bool unres = false;
        int pos = 0;
        if (valid) {
            pos = s_list[e];
            const int i = perm[pos];
            const float4 q = cur[i];
            float4 t0 = tg[i];
            const int t_first = __float_as_int(t0.w);
            unres = true;
            if (t_first >= 0) {
                float sd = sqrtf(d2_rn(q.x, q.y, q.z, t0.x, t0.y, t0.z));
                // a match that T_k left far behind (beyond the radius its neighbour list covers) bounds the search badly:
                // walk downhill over the neighbour lists first -- ANY target is a valid bound, and a better one makes
                // the cube (or, for the general kernels, the seeded pyramid descent) small  (every lane of the group walks)
                if (walk) {
                    float d0 = sd * sd;
                    for (int step = 0; step < 64; ++step) {
                        const float4* L = gv.knn + (size_t)__float_as_int(t0.w) * LG_KSLOTS;
                        if (!(sd > __ldg(L).x)) break;
                        float4 nb = t0; float dn = d0;
#pragma unroll
                        for (int k = 1; k < LG_KSLOTS; ++k) {
                            const float4 c = __ldg(L + k);
                            const float d = d2_rn(q.x, q.y, q.z, c.x, c.y, c.z);
                            if (d < dn) { dn = d; nb = c; }
                        }
                        if (!(dn < d0)) break;
                        t0 = nb; d0 = dn; sd = sqrtf(dn);
                    }
                    if (sub == 0 && __float_as_int(t0.w) != t_first) tg[i] = t0;          // (the general kernels seed from it)
                }
                float nx, ny, nz;
                xform_point(T, q.x, q.y, q.z, nx, ny, nz);                // how far T_k would move the point once more
                const float mv = sqrtf((nx - q.x) * (nx - q.x) + (ny - q.y) * (ny - q.y) + (nz - q.z) * (nz - q.z));
                const float rho = sd * LG_CERT_UP + (2.0f * mv < m ? m : 0.0f);
                if (rho * g.inv_h * 1.00001f + 0.003f <= 2.6f) {          // (false for NaN too)
                    Best2 b; b.init();
                    lg_cube_search(gacc, g, q.x, q.y, q.z, rho, sub, G, b);
                    if (G > 1) {                                          // the lanes scanned disjoint rows
#pragma unroll
                        for (int off = G / 2; off >= 1; off >>= 1) {
                            const unsigned long long ok = __shfl_xor_sync(gmask, b.key, off);
                            const float os = __shfl_xor_sync(gmask, b.second, off);
                            b.update(ok);
                            b.second = fminf(b.second, os);
                        }
                    }
                    if (b.key != 0xffffffffffffffffull) {                 // (the match itself lies in the cube)
                        unres = false;
                        if (sub == 0) {
                            const unsigned ti = (unsigned)b.key;
                            const float4 t = ti == (unsigned)__float_as_int(t0.w) ? t0 : __ldg(gv.t_orig + ti);
                            tg[i] = make_float4(t.x, t.y, t.z, __int_as_float((int)ti));
                            cur[i] = cert[i] = make_float4(q.x, q.y, q.z, sqrtf(fminf(b.second, rho * rho)) * LG_CERT_DOWN);
                            flagS[pos] = 0;
                        }
                    }
                }
            }
        }
        const bool mine = valid && unres && sub == 0;
        const unsigned bal = __ballot_sync(KSS_FULL, mine);
        if (bal) {
            unsigned at = 0u;
            if (lane == 0) at = atomicAdd(&st->n_unres, (unsigned)__popc(bal));
            at = __shfl_sync(KSS_FULL, at, 0);
            if (mine) worklist[at + __popc(bal & ((1u << lane) - 1u))] = pos;
        }
    }
}

constexpr int RF_SEG = 2048;                     // sorted positions per CTA of lg_refine_kernel
__global__ void __launch_bounds__(256)
lg_refine_kernel(LgGridView gv, int n_q, float4* __restrict__ cur, float4* __restrict__ tg, float4* __restrict__ tg2, float4* __restrict__ cert, const int* __restrict__ perm,
                 unsigned char* __restrict__ flagS, int* __restrict__ worklist, LgState* __restrict__ st, float margin_cells, int gsel /* lanes per query (1 or 4); 0 = by the count */) {
    __shared__ LgGeom g;
    __shared__ float T[16];
    __shared__ int s_list[RF_SEG];               // flagged positions of the segment, compacted: every lane of a warp searches
    __shared__ int s_n;
    lg_wait_prior();
    if (st->done || st->n_fail == 0u) return;
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < (int)(sizeof(LgGeom) / 4)) reinterpret_cast<int*>(&g)[tid] = reinterpret_cast<const int*>(gv.geom)[tid];
    if (tid < 16) T[tid] = st->Tk[tid];
    if (tid == 0) s_n = 0;
    __syncthreads();
    {   // 8 flags per thread (one 8-byte load), ballot-free compaction: order within the segment does not matter
        const int p0 = blockIdx.x * RF_SEG + tid * 8;
        unsigned long long f8 = 0ull;
        if (p0 + 8 <= n_q) f8 = *reinterpret_cast<const unsigned long long*>(flagS + p0);
        else for (int j = 0; j < 8; ++j) if (p0 + j < n_q && flagS[p0 + j]) f8 |= 1ull << (8 * j);
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) cnt += (int)((f8 >> (8 * j)) & 1ull);
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(KSS_FULL, incl, o); if (lane >= o) incl += y; }
        int wbase = 0;
        if (lane == 31 && incl) wbase = atomicAdd(&s_n, incl);
        wbase = __shfl_sync(KSS_FULL, wbase, 31);
        int at = wbase + incl - cnt;
#pragma unroll
        for (int j = 0; j < 8; ++j) if ((f8 >> (8 * j)) & 1ull) s_list[at++] = p0 + j;
    }
    __syncthreads();
    const int nl = s_n;
    const float m = margin_cells * g.h;
    const bool walk = gv.knn != nullptr;
    GlobAcc gacc{gv.blk_rank, gv.fine_start, gv.tp, gv.lut};
    if (nl <= 8) {
        // a handful (the steady state: points beyond the target's rim, whose second-nearest target is as far as the
        // nearest, keep failing): one WARP per query, lane = row of cells, so that nobody waits for a long serial search
        const int w = tid >> 5;
        if (w >= nl) return;
        const int pos = s_list[w];
        const int i = perm[pos];
        const float4 q = cur[i], t0 = tg[i];
        bool unres = true;
        if (__float_as_int(t0.w) >= 0) {
            const float rho = sqrtf(d2_rn(q.x, q.y, q.z, t0.x, t0.y, t0.z)) * 1.001f + m;
            if (rho * g.inv_h * 1.00001f + 0.003f <= cube3 && lg_warp_search_store(gacc, g, gv.t_orig, q, t0, rho, lane, i, cur, tg, tg2, cert)) {
                unres = false;
                if (lane == 0) flagS[pos] = 0;
#ifdef KSS_LG_DEBUG
                if (lane == 0 && st->iters > 400 && st->iters < 404) {
                    const float4 c2 = cur[i], n1 = tg[i], n2 = tg2[i];
                    const float4 L0 = __ldg(gv.knn + (size_t)__float_as_int(t0.w) * LG_KSLOTS);
                    printf("it %d i %d: sd %.6g (cells %.3f) old match %d -> %d, new cert %.6g (gap %.3g), runner-up %d at %.6g, list R %.6g, rho %.6g\n", st->iters, i,
                           sqrtf(d2_rn(q.x, q.y, q.z, t0.x, t0.y, t0.z)), sqrtf(d2_rn(q.x, q.y, q.z, t0.x, t0.y, t0.z)) * g.inv_h, __float_as_int(t0.w), __float_as_int(n1.w), c2.w,
                           fabsf(c2.w) - sqrtf(d2_rn(q.x, q.y, q.z, n1.x, n1.y, n1.z)), c2.w < 0 ? __float_as_int(n2.w) : -1,
                           c2.w < 0 ? sqrtf(d2_rn(q.x, q.y, q.z, n2.x, n2.y, n2.z)) : 0.f, L0.x, rho);
                }
#endif
            }
        }
        if (unres && lane == 0) worklist[atomicAdd(&st->n_unres, 1u)] = pos;
        return;
    }
    // Up to 128 flagged points in the segment (the iterations around PCL's convergence): four lanes per query, lane = every
    // fourth row of cells of the cube -- one thread each leaves most of the CTA idle and every search one serial chain of
    // dependent loads.  Beyond that one thread per query is the cheaper order: the query's own row comes first and its two
    // distances prune most other rows, which lanes that start elsewhere scan for nothing (measured: 2, 4, 8 lanes; 32 .. 512).
    if (gsel == 0) gsel = nl <= 128 ? 4 : 1;
    if (gsel >= 4) lg_refine_queries<4>(gacc, g, gv, T, m, walk, s_list, nl, cur, tg, cert, perm, flagS, worklist, st);
    else lg_refine_queries<1>(gacc, g, gv, T, m, walk, s_list, nl, cur, tg, cert, perm, flagS, worklist, st);
}

// The few queries lg_refine_kernel could not finish, one WARP each (persistent grid over the work list): a query with a
// match reads the cells meeting the cube q +- bound (lane = row, up to 17 x 17 rows); the rest goes to the pyramid.
__global__ void __launch_bounds__(256)
lg_left_kernel(Pyramid py, LgGridView gv, int n_q, float4* __restrict__ cur, float4* __restrict__ tg, float4* __restrict__ tg2, float4* __restrict__ cert, const int* __restrict__ perm,
               unsigned char* __restrict__ flagS, const int* __restrict__ worklist, LgState* __restrict__ st, int staged_launched) {
    __shared__ LgGeom g;
    __shared__ float4 slots[8][TILE];
    lg_wait_prior();
    if (st->done) return;
    const unsigned nu = st->n_unres;
    if (nu == 0u || (staged_launched && nu > (unsigned)n_q / LG_STAGED_DIV)) return;          // many: lg_nn_kernel<1> staged them
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < (int)(sizeof(LgGeom) / 4)) reinterpret_cast<int*>(&g)[tid] = reinterpret_cast<const int*>(gv.geom)[tid];
    __syncthreads();
    GlobAcc gacc{gv.blk_rank, gv.fine_start, gv.tp, gv.lut};
    const unsigned nw = gridDim.x * (blockDim.x >> 5);
    unsigned n_pyr = 0u;
    for (unsigned e = blockIdx.x * (blockDim.x >> 5) + warp; e < nu; e += nw) {
        const int pos = worklist[e];
        const int i = perm[pos];
        const float4 q = cur[i], t0 = tg[i];
        bool found = false;
        if (__float_as_int(t0.w) >= 0) {
            // (0.1 % beyond the match: far from the target the second-nearest is about as far as the nearest)
            const float rho = sqrtf(d2_rn(q.x, q.y, q.z, t0.x, t0.y, t0.z)) * 1.001f + 0.25f * g.h;
            if (rho * g.inv_h * 1.00001f + 0.003f <= cube3) found = lg_warp_search_store(gacc, g, gv.t_orig, q, t0, rho, lane, i, cur, tg, tg2, cert);
        }
        if (!found) {
            const float seed = __float_as_int(t0.w) >= 0 ? d2_rn(q.x, q.y, q.z, t0.x, t0.y, t0.z) : __int_as_float(0x7f800000);
            const unsigned long long kk = lg_warp_nn(py, q.x, q.y, q.z, slots[warp], seed); ++n_pyr;       // every lane carries the same query
            if (lane == 0) {
                const unsigned ti = (unsigned)kk;
                const float4 t = __ldg(gv.t_orig + ti);
                tg[i] = make_float4(t.x, t.y, t.z, __int_as_float((int)ti));
                cur[i] = make_float4(q.x, q.y, q.z, 0.0f);
            }
        }
        if (lane == 0) flagS[pos] = 0;
    }
    if (lane == 0 && n_pyr) atomicAdd(&st->miss[2], n_pyr);
}

// ------------------------------------------------------------------ canonical reductions, large n
// level 1: one warp per 256-element chunk of ORIGINAL indices (lane-strided partials + butterfly)
// upper levels (<= 2 more, same rule): for the one-off double sums (fitness, PCR_QM: lg_passF_kernel) the last CTA to
// finish reduces the chunk partials (lg_finish_f64); the ICP passes use the tree further down (LgTree)
__device__ __forceinline__ double lg_finish_f64(const double* __restrict__ part, int nchunks) {
    __shared__ double lvl2d[256];
    __shared__ double res;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int n2 = (nchunks + 255) >> 8;
    for (int c = warp; c < n2; c += nwarps) {
        double v[8];
        bool have[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = (c << 8) + lane + 32 * u; have[u] = i < nchunks; v[u] = have[u] ? __ldcg(part + i) : 0.0; }
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

#ifdef KSS_LG_TIMELINE
__device__ unsigned long long lg_tl[16];
__device__ __forceinline__ unsigned long long lg_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define LG_TL_MIN(i) do { if (threadIdx.x == 0) atomicMin(&lg_tl[i], lg_now()); } while (0)
#define LG_TL_SET(i) do { if (threadIdx.x == 0) lg_tl[i] = lg_now(); } while (0)
#else
#define LG_TL_MIN(i) do {} while (0)
#define LG_TL_SET(i) do {} while (0)
#endif
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

// ---- the two reduction passes of an iteration (pcl::umeyama demeans before it multiplies: the means first, then sigma)
// A lane's 8 points of a 256-chunk of ORIGINAL indices, as the reductions want them: sv = {source xyz, d2},
// tv = {target xyz, bits(index, or -1 when rejected: d2 > max_dist^2, A.3)} -- two coalesced streams (cur, tg)
__device__ __forceinline__ void lg_pair(const float4& c, const float4& t, double max2, float4& sv, float4& tv) {
    const float d2 = d2_rn(c.x, c.y, c.z, t.x, t.y, t.z);
    const int ti = __float_as_int(t.w);
    sv = make_float4(c.x, c.y, c.z, d2);
    tv = make_float4(t.x, t.y, t.z, __int_as_float((ti < 0 || (double)d2 > max2) ? -1 : ti));
}
// pass A of one 256-element chunk (one warp): sums of kept source xyz, matched target xyz (float), d2 (double), kept count.
// A lane adds ITS points in increasing index (the canonical order): lg_accum_A can be called for the first four and then the
// last four; lg_store_A is the butterfly + the partials.
template <int N>
__device__ __forceinline__ void lg_accum_A(const float4 (&sv)[N], const float4 (&tv)[N], float (&a)[6], double& dsum, int& k) {
#pragma unroll
    for (int u = 0; u < N; ++u) {
        if (__float_as_int(tv[u].w) < 0) continue;
        a[0] = __fadd_rn(a[0], sv[u].x); a[1] = __fadd_rn(a[1], sv[u].y); a[2] = __fadd_rn(a[2], sv[u].z);
        a[3] = __fadd_rn(a[3], tv[u].x); a[4] = __fadd_rn(a[4], tv[u].y); a[5] = __fadd_rn(a[5], tv[u].z);
        dsum = __dadd_rn(dsum, (double)sv[u].w);
        ++k;
    }
}
__device__ __forceinline__ void lg_store_A(float (&a)[6], double dsum, int k, int c, int S, float* __restrict__ partA,
                                           double* __restrict__ partD, int* __restrict__ partK) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
#pragma unroll
        for (int q = 0; q < 6; ++q) a[q] = __fadd_rn(a[q], __shfl_xor_sync(KSS_FULL, a[q], off));
        dsum = __dadd_rn(dsum, __shfl_xor_sync(KSS_FULL, dsum, off));
    }
    k = __reduce_add_sync(KSS_FULL, k);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int q = 0; q < 6; ++q) partA[(size_t)q * S + c] = a[q];
        partD[c] = dsum; partK[c] = k;
    }
}
__device__ __forceinline__ void lg_chunk_A(const float4 (&sv)[8], const float4 (&tv)[8], int c, int S, float* __restrict__ partA,
                                           double* __restrict__ partD, int* __restrict__ partK) {
    float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    double dsum = 0.0;
    int k = 0;
    lg_accum_A<8>(sv, tv, a, dsum, k);
    lg_store_A(a, dsum, k, c, S, partA, partD, partK);
}
// ---- upper levels of the canonical sums, spread over the grid: partials are SoA [quantity][S] (S = padded chunk
// count).  The CTA that completes a group of 256 chunks (per-group counters) reduces the group with ONE warp (level 2,
// lane-strided + butterfly as everywhere); the warp that completes the last group does level 3 and the finish.  Nothing
// waits: the serial tail of a pass is one group's level 2 + level 3 + the scalar finish.
struct LgTree {
    unsigned* grpcnt;                // [S2] chunks of each group that have arrived (reset by the last warp)
    float* l2f;                      // [9][S2] level-2 results
    double* l2d;                     // [S2]
    int* l2k;                        // [S2]
    int S, S2;                       // padded chunk / group counts
};
template <int NQ>
__device__ __forceinline__ void lg_group_f32(const float* __restrict__ part, int S, int nchunks, int g, float* __restrict__ l2, int S2) {
    const int lane = threadIdx.x & 31;
    float v[NQ][8];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = (g << 8) + lane + 32 * u;
            v[q][u] = i < nchunks ? __ldcg(part + (size_t)q * S + i) : 0.0f;          // (L2: written by other CTAs)
        }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        float p = 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u) if ((g << 8) + lane + 32 * u < nchunks) p = __fadd_rn(p, v[q][u]);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) p = __fadd_rn(p, __shfl_xor_sync(KSS_FULL, p, off));
        if (nchunks == 1) p = __shfl_sync(KSS_FULL, v[q][0], 0);                       // a single chunk is its own total
        if (lane == 0) l2[(size_t)q * S2 + g] = p;
    }
}
__device__ __forceinline__ void lg_group_f64(const double* __restrict__ part, int nchunks, int g, double* __restrict__ l2) {
    const int lane = threadIdx.x & 31;
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { const int i = (g << 8) + lane + 32 * u; v[u] = i < nchunks ? __ldcg(part + i) : 0.0; }
    double p = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) if ((g << 8) + lane + 32 * u < nchunks) p = __dadd_rn(p, v[u]);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) p = __dadd_rn(p, __shfl_xor_sync(KSS_FULL, p, off));
    if (nchunks == 1) p = __shfl_sync(KSS_FULL, v[0], 0);
    if (lane == 0) l2[g] = p;
}
// level 3 (one warp): every lane ends up with the totals.  All loads are issued before the first add (n2 <= 256: up to
// 8 values per lane and quantity), the adds keep the canonical order (lane-strided, then the butterfly).
template <int NQ>
__device__ __forceinline__ void lg_top_f32(const float* __restrict__ l2, int S2, int n2, float (&tot)[NQ]) {
    const int lane = threadIdx.x & 31;
    if (n2 <= 32) {
        float v[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) v[q] = lane < n2 ? __ldcg(l2 + (size_t)q * S2 + lane) : 0.0f;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            if (n2 == 1) { tot[q] = __shfl_sync(KSS_FULL, v[q], 0); continue; }
            float p = lane < n2 ? __fadd_rn(0.0f, v[q]) : 0.0f;
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) p = __fadd_rn(p, __shfl_xor_sync(KSS_FULL, p, off));
            tot[q] = p;
        }
        return;
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = lane + 32 * u < n2 ? __ldcg(l2 + (size_t)q * S2 + lane + 32 * u) : 0.0f;
        float p = 0.0f;
#pragma unroll
        for (int u = 0; u < 8; ++u) if (lane + 32 * u < n2) p = __fadd_rn(p, v[u]);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) p = __fadd_rn(p, __shfl_xor_sync(KSS_FULL, p, off));
        tot[q] = p;
    }
}
__device__ __forceinline__ double lg_top_f64(const double* __restrict__ l2, int n2) {
    const int lane = threadIdx.x & 31;
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = lane + 32 * u < n2 ? __ldcg(l2 + lane + 32 * u) : 0.0;
    if (n2 == 1) return __shfl_sync(KSS_FULL, v[0], 0);
    double p = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) if (lane + 32 * u < n2) p = __dadd_rn(p, v[u]);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) p = __dadd_rn(p, __shfl_xor_sync(KSS_FULL, p, off));
    return p;
}
// CTA b holds the chunks 8b .. 8b+7 (all of group b >> 5) and brings `mine` of them (thread 0's value counts; 0: no
// arrival).  Returns the group to reduce in warp 0 of the CTA that completed it, -1 everywhere else.
__device__ __forceinline__ int lg_group_arrive(unsigned* __restrict__ grpcnt, int nchunks, unsigned mine) {
    __shared__ int s_g;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const int g = blockIdx.x >> 5;
        const unsigned all = (unsigned)min(256, nchunks - (g << 8));
        s_g = (mine && atomicAdd(&grpcnt[g], mine) + mine == all) ? g : -1;
    }
    __syncthreads();
    const int g = s_g;
    if (g < 0 || threadIdx.x >= 32) return -1;
    __threadfence();
    return g;
}
// after the group's level 2 is written: true in the warp that completed the last group (counters reset for the next pass)
__device__ __forceinline__ bool lg_groups_done(unsigned* __restrict__ grpcnt, unsigned* __restrict__ ticket, int n2) {
    const int lane = threadIdx.x & 31;
    __threadfence();
    __syncwarp();
    int last = 0;
    if (lane == 0) last = atomicAdd(ticket, 1u) == (unsigned)(n2 - 1);
    last = __shfl_sync(KSS_FULL, last, 0);
    if (!last) return false;
    __threadfence();
    for (int i = lane; i < n2; i += 32) grpcnt[i] = 0u;
    if (lane == 0) *ticket = 0u;
    return true;
}

__device__ __forceinline__ void lg_group_A(const float* __restrict__ partA, const double* __restrict__ partD, const int* __restrict__ partK,
                                           const LgTree& tr, int nchunks, int g) {
    lg_group_f32<6>(partA, tr.S, nchunks, g, tr.l2f, tr.S2);
    lg_group_f64(partD, nchunks, g, tr.l2d);
    int k = 0;                                   // kept counts: integer, any order
    for (int i = (g << 8) + (threadIdx.x & 31); i < min(nchunks, (g + 1) << 8); i += 32) k += __ldcg(partK + i);
    k = __reduce_add_sync(KSS_FULL, k);
    if ((threadIdx.x & 31) == 0) tr.l2k[g] = k;
}
// one warp: level 3 of pass A, the means, the mse
__device__ __forceinline__ void lg_finish_A(int nchunks, const LgTree& tr, LgState* __restrict__ st, volatile int* __restrict__ h_unres) {
    const int n2 = (nchunks + 255) >> 8;
    float tot[6];
    int k = 0;
    for (int i = threadIdx.x & 31; i < n2; i += 32) k += __ldcg(tr.l2k + i);
    const unsigned nu = st->n_unres, nf = st->n_fail, m1 = st->miss[1];      // (loads before the first store to *st)
    const double dtot = lg_top_f64(tr.l2d, n2);
    lg_top_f32<6>(tr.l2f, tr.S2, n2, tot);
    k = __reduce_add_sync(KSS_FULL, k);
    if ((threadIdx.x & 31) == 0) {
        if (h_unres) *h_unres = (int)min(nu, 0x7fffffffu);           // the host reads it, stale, to decide which general kernel to enqueue
        st->miss[1] = m1 + nf;
        st->n_unres = 0u; st->n_fail = 0u;                           // (all search kernels of this iteration are done)
        const int cnt = k;
        st->kept = cnt;
        if (cnt >= 3) {
            const float oon = div_(1.0f, (float)cnt);
            st->one_over_n = oon;
            for (int a2 = 0; a2 < 3; ++a2) { st->smean[a2] = mul_(tot[a2], oon); st->dmean[a2] = mul_(tot[3 + a2], oon); }
            st->mse = __ddiv_rn(dtot, (double)cnt);
        }
    }
}
// pass B of one chunk: sigma(a, b) partials = sum (d_a - dmean_a) * (s_b - smean_b), in two halves of four points per lane
// (fewer live registers: every CTA of the pass is resident at once); the first half may be loaded before the means are known
__device__ __forceinline__ void lg_load_half(const float4* __restrict__ cur, const float4* __restrict__ tg, int n, int c, int half,
                                             float4 (&cv)[4], float4 (&gv)[4]) {
    const int i0 = (c << 8) + (threadIdx.x & 31) + 128 * half;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = i0 + 32 * u;
        gv[u].w = __int_as_float(-1);
        if (i < n) { cv[u] = cur[i]; gv[u] = tg[i]; }
    }
}
__device__ __forceinline__ void lg_accum_B(const float4 (&cv)[4], const float4 (&gv)[4], double max2, float sm0, float sm1, float sm2,
                                           float dm0, float dm1, float dm2, float (&a)[9]) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int ti = __float_as_int(gv[u].w);
        if (ti < 0) continue;                                       // (beyond n, or no match)
        const float4 s = cv[u], t = gv[u];
        if ((double)d2_rn(s.x, s.y, s.z, t.x, t.y, t.z) > max2) continue;       // rejected (A.3), as lg_pair decides
        const float sx = sub_(s.x, sm0), sy = sub_(s.y, sm1), sz = sub_(s.z, sm2);
        const float dx = sub_(t.x, dm0), dy = sub_(t.y, dm1), dz = sub_(t.z, dm2);
        a[0] = add_(a[0], mul_(dx, sx)); a[1] = add_(a[1], mul_(dx, sy)); a[2] = add_(a[2], mul_(dx, sz));
        a[3] = add_(a[3], mul_(dy, sx)); a[4] = add_(a[4], mul_(dy, sy)); a[5] = add_(a[5], mul_(dy, sz));
        a[6] = add_(a[6], mul_(dz, sx)); a[7] = add_(a[7], mul_(dz, sy)); a[8] = add_(a[8], mul_(dz, sz));
    }
}
// one warp: level 3 of pass B, then umeyama + accumulate + convergence on one thread
__device__ __forceinline__ void lg_finish_B(int nchunks, const LgTree& tr, LgState* __restrict__ st, bool enough,
                                            int max_iter, double rot_thr, double trans_thr, double mse_rel, double mse_abs) {
    const int n2 = (nchunks + 255) >> 8;
    float tot[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) tot[q] = 0.0f;
    // everything the scalar finish reads from *st, before anything is stored to it (one round trip to L2)
    const int lane = threadIdx.x & 31;
    const float stv = lane < 16 ? st->fin[lane] : lane < 19 ? st->smean[lane - 16] : lane < 22 ? st->dmean[lane - 19] : lane == 22 ? st->one_over_n : 0.0f;
    const int it0 = st->iters;
    const double mse = st->mse, prev = st->prev_mse;
    if (enough) lg_top_f32<9>(tr.l2f, tr.S2, n2, tot);
    float F[16], sm[3], dm[3];
#pragma unroll
    for (int i = 0; i < 16; ++i) F[i] = __shfl_sync(KSS_FULL, stv, i);
#pragma unroll
    for (int i = 0; i < 3; ++i) { sm[i] = __shfl_sync(KSS_FULL, stv, 16 + i); dm[i] = __shfl_sync(KSS_FULL, stv, 19 + i); }
    const float oon = __shfl_sync(KSS_FULL, stv, 22);
    if (threadIdx.x == 0) {
        if (!enough) { st->done = 1; st->converged = 0; return; }      // min_number_correspondences_
        float sigma[9], T[16];
#pragma unroll
        for (int i = 0; i < 9; ++i) sigma[i] = mul_(oon, tot[i]);
        umeyama_finish(sigma, sm, dm, T);
        mat4_mul(T, F, F);
#pragma unroll
        for (int i = 0; i < 16; ++i) { st->Tk[i] = T[i]; st->fin[i] = F[i]; }
        st->apply_T = 1;
        const int it = it0 + 1;
        st->iters = it;
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

// The streaming pass of an iteration, by ORIGINAL index (one warp per 256-chunk, 8 points per lane -- the order of the
// canonical sums): transformCloud of the previous iteration (A.5) and the certificate test.  Failed certificates are
// queued per warp and re-established 32 at a time from the neighbour list of the match (every target within
// R(p) - |q p| of q is in {p} + list(p): best and second-best are exact whenever the best is that close).  Chunks in
// which every match was settled here get pass A right away.  Points that stay open are flagged at their SORTED
// position for the search kernels; their chunks are marked dirty for lg_passA_kernel.
constexpr int TK_QCAP = 64;                      // queue entries per warp (a round of 32 is processed whenever it holds >= 32)

// one round: lane e < cnt takes queue entry e.  Returns 1 if the entry stays open.
__device__ __forceinline__ unsigned lg_list_round(const float4* __restrict__ qa, const float4* __restrict__ qb, int cnt, int lane,
                                                  const float4* __restrict__ knn, float4* __restrict__ cur, float4* __restrict__ tg,
                                                  float4* __restrict__ cert, unsigned& nlist) {
    if (lane >= cnt) return 0u;
    const float4 q = qa[lane];                   // {x, y, z, bits(i)}
    const float4 p = qb[lane];                   // the match {x, y, z, bits(index)}
    const int i = __float_as_int(q.w);
    const float4* L = knn + (size_t)__float_as_int(p.w) * LG_KSLOTS;
    float4 e[LG_KSLOTS];
#pragma unroll
    for (int k = 0; k < LG_KSLOTS; ++k) e[k] = __ldg(L + k);
    unsigned long long bk = ((unsigned long long)__float_as_uint(d2_rn(q.x, q.y, q.z, p.x, p.y, p.z)) << 32) | __float_as_uint(p.w);
    const float rho = e[0].x - sqrtf(__uint_as_float((unsigned)(bk >> 32))) * LG_CERT_UP;      // everything within rho of q is in {p} + list(p)
    float second = __int_as_float(0x7f800000);
    int kb = 0;
#pragma unroll
    for (int k = 1; k < LG_KSLOTS; ++k) {
        const unsigned long long ck = ((unsigned long long)__float_as_uint(d2_rn(q.x, q.y, q.z, e[k].x, e[k].y, e[k].z)) << 32) | __float_as_uint(e[k].w);
        const bool better = ck < bk;
        const unsigned long long hi = better ? bk : ck;
        second = fminf(second, __uint_as_float((unsigned)(hi >> 32)));
        bk = better ? ck : bk;
        kb = better ? k : kb;
    }
    const float db = sqrtf(__uint_as_float((unsigned)(bk >> 32)));
    if (db * LG_CERT_UP < rho) {                 // (false for NaN and for rho <= 0)
        cur[i] = cert[i] = make_float4(q.x, q.y, q.z, fminf(sqrtf(second), rho) * LG_CERT_DOWN);
        if (kb) {
            float4 t = e[1];
#pragma unroll
            for (int k = 2; k < LG_KSLOTS; ++k) if (k == kb) t = e[k];
            tg[i] = t;
        }
        ++nlist;
        return 0u;
    }
    return 1u;                                   // stays open: the caller decides who searches it
}

__device__ __forceinline__ float lg_sqrt_approx(float x) {      // relative error 2^-22, far inside the certificate margins
    float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r;
}
__global__ void __launch_bounds__(256, 2)
lg_track_kernel(int n, int nchunks, float4* __restrict__ cur, float4* __restrict__ tg, float4* __restrict__ tg2, float4* __restrict__ cert, const int* __restrict__ inv,
                const float4* __restrict__ knn, unsigned char* __restrict__ flagS, unsigned char* __restrict__ dirty,
                float* __restrict__ partA, double* __restrict__ partD, int* __restrict__ partK, LgTree tr,
                LgState* __restrict__ st, double max2, volatile int* __restrict__ h_unres) {
    __shared__ float T[16];
    __shared__ float4 s_qa[8][TK_QCAP], s_qb[8][TK_QCAP];
    __shared__ unsigned s_clean;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 8 + warp;
    const int i0 = (c << 8) + lane;
    float4 sv[8], tv[8], t2v[8];                     // {position, d2}, {match, index}: what the sums need; runner-ups
    if (c < nchunks) {                               // (pass B only reads them: before the wait)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + 32 * u;
            if (i < n) { sv[u] = cur[i]; tv[u] = tg[i]; }
        }
        // runner-ups of the two-candidate certificates (8 % of the points at a fixed point, so nearly every warp has one in
        // every round of 32): all loads in flight at once -- fetched inside the loop they were eight serial round trips --
        // and before the wait too (pass B does not write them either)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + 32 * u;
            t2v[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            if (i < n && sv[u].w < 0.0f) t2v[u] = tg2[i];
        }
    }
#ifdef KSS_LG_TIMELINE
    const unsigned long long k0 = lg_now();
#endif
    lg_wait_prior();
    if (st->done) return;
    if (threadIdx.x < 16) T[threadIdx.x] = st->Tk[threadIdx.x];
    if (threadIdx.x == 0) s_clean = 0u;
    __syncthreads();
#ifdef KSS_LG_TIMELINE
    const unsigned long long k1 = lg_now();
#endif
    const bool applyT = st->apply_T != 0;
    if (c < nchunks) {
        float4* qa = s_qa[warp]; float4* qb = s_qb[warp];
        // entries a list round left open go to the search kernels
        auto leftover = [&](unsigned open, int) -> unsigned {
            if (open) { const float4 q = qa[lane]; const int i = __float_as_int(q.w); cur[i] = make_float4(q.x, q.y, q.z, 0.0f); flagS[inv[i]] = 1; }
            return open;
        };
        unsigned nfail = 0u, nlist = 0u, redo = 0u;  // redo: bit u = this lane's point u went through the queue
        int qn = 0;                                  // queue fill (warp-uniform)
#ifdef KSS_LG_TIMELINE
        const unsigned long long ka = lg_now();
        if (__float_as_int(sv[0].w) == 0x7fc12345 || __float_as_int(tv[7].w) == 0x12345678) atomicAdd(&lg_tl[14], 1ull);   // (first use of the loaded data)
        const unsigned long long kb = lg_now();
#endif
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + 32 * u;
            bool push = false, open = false;
            if (i < n) {
                const float4 cv = sv[u];
                float x = cv.x, y = cv.y, z = cv.z, lb = fabsf(cv.w);
                const bool two = cv.w < 0.0f;                               // the certificate names the runner-up (tg2)
                if (applyT) {
                    xform_point(T, cv.x, cv.y, cv.z, x, y, z);
                    const float mx = x - cv.x, my = y - cv.y, mz = z - cv.z;
                    lb = (lb - lg_sqrt_approx(mx * mx + my * my + mz * mz) * LG_CERT_UP) * LG_CERT_DOWN;
                }
                float d2 = d2_rn(x, y, z, tv[u].x, tv[u].y, tv[u].z);
                if (two) {                                                  // (rare: points about equally far from two targets)
                    const float4 t2 = t2v[u];
                    const float d2b = d2_rn(x, y, z, t2.x, t2.y, t2.z);
                    const unsigned long long ka = ((unsigned long long)__float_as_uint(d2) << 32) | __float_as_uint(tv[u].w);
                    const unsigned long long kb = ((unsigned long long)__float_as_uint(d2b) << 32) | __float_as_uint(t2.w);
                    if (kb < ka) { tg2[i] = tv[u]; tg[i] = t2; tv[u] = t2; d2 = d2b; }
                }
                sv[u] = make_float4(x, y, z, d2);
                const bool has = __float_as_int(tv[u].w) >= 0;
                // sqrt(d2) * UP < lb, squared (lb > 0)
                bool ok = has && lb > 0.0f && d2 * (LG_CERT_UP * LG_CERT_UP) < lb * lb;
                if (!ok && has && cv.w != 0.0f) {
                    // lb lost the length of the whole PATH since the certificate was established; what counts is the
                    // DISPLACEMENT from where it was established (at a fixed point the cloud jitters by an fp32 ulp per
                    // iteration and goes nowhere): second chance from the recorded position, and a fresh start for lb
                    const float4 c0 = cert[i];
                    const float ex = x - c0.x, ey = y - c0.y, ez = z - c0.z;
                    lb = (fabsf(c0.w) - sqrtf(ex * ex + ey * ey + ez * ez) * LG_CERT_UP) * LG_CERT_DOWN;
                    ok = lb > 0.0f && d2 * (LG_CERT_UP * LG_CERT_UP) < lb * lb;
                }
                if (ok) cur[i] = make_float4(x, y, z, two ? -lb : lb);
                else if (has) push = true;
                else { open = true; cur[i] = make_float4(x, y, z, 0.0f); flagS[inv[i]] = 1; }
            }
            if (open) ++nfail;
            const unsigned pb = __ballot_sync(KSS_FULL, push);
            if (pb) {                                                     // (warp-uniform)
                if (push) {
                    const int at = qn + __popc(pb & ((1u << lane) - 1u));
                    qa[at] = make_float4(sv[u].x, sv[u].y, sv[u].z, __int_as_float(i));
                    qb[at] = tv[u];
                    redo |= 1u << u;
                }
                qn += __popc(pb);
                __syncwarp();
                if (qn >= 32) {
                    nfail += leftover(lg_list_round(qa, qb, 32, lane, knn, cur, tg, cert, nlist), 32);
                    __syncwarp();
                    if (lane + 32 < qn) { const float4 a = qa[lane + 32], b2 = qb[lane + 32]; __syncwarp(); qa[lane] = a; qb[lane] = b2; }
                    else __syncwarp();
                    qn -= 32;
                    __syncwarp();
                }
            }
        }
        if (qn > 0) nfail += leftover(lg_list_round(qa, qb, qn, lane, knn, cur, tg, cert, nlist), qn);
#ifdef KSS_LG_TIMELINE
        const unsigned long long kc = lg_now();
        if (threadIdx.x == 0 && st->iters > 400) { atomicAdd(&lg_tl[1], kb - ka); atomicAdd(&lg_tl[2], kc - kb); }
#endif
        nfail = __reduce_add_sync(KSS_FULL, nfail);
        nlist = __reduce_add_sync(KSS_FULL, nlist);
        if (lane == 0) {
            dirty[c] = nfail ? 1 : 0;
            if (nfail) atomicAdd(&st->n_fail, nfail);
            if (nlist) atomicAdd(&st->miss[0], nlist);
            if (!nfail) atomicAdd(&s_clean, 1u);
        }
        __syncwarp();                                // (without it lane 0 and the other 31 run the whole of pass A one after the other)
        if (!nfail) {
            // pass A of the chunk: matches that went through the queue may have changed
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int i = i0 + 32 * u;
                if (i >= n) { tv[u].w = __int_as_float(-1); continue; }
                if ((redo >> u) & 1u) tv[u] = __ldcg(tg + i);
                lg_pair(sv[u], tv[u], max2, sv[u], tv[u]);
            }
            lg_chunk_A(sv, tv, c, tr.S, partA, partD, partK);
        }
#ifdef KSS_LG_TIMELINE
        if (threadIdx.x == 0 && st->iters > 400) atomicAdd(&lg_tl[3], lg_now() - kc);
#endif
    }
    // the clean chunks arrive at their group; a group (and, with nothing open anywhere, the whole of pass A) that
    // completes here is reduced here
#ifdef KSS_LG_TIMELINE
    const unsigned long long k2 = lg_now();
#endif
    __syncthreads();
    const int grp = lg_group_arrive(tr.grpcnt, nchunks, s_clean);
#ifdef KSS_LG_TIMELINE
    if (threadIdx.x == 0 && st->iters > 400) {
        const unsigned long long k3 = lg_now();
        atomicAdd(&lg_tl[8], k1 - k0); atomicAdd(&lg_tl[9], k2 - k1); atomicAdd(&lg_tl[10], k3 - k2); atomicAdd(&lg_tl[11], 1ull);
    }
#endif
    if (grp < 0) return;
    lg_group_A(partA, partD, partK, tr, nchunks, grp);
    if (!lg_groups_done(tr.grpcnt, &st->ticketA, (nchunks + 255) >> 8)) return;
    lg_finish_A(nchunks, tr, st, h_unres);
}

// pass A of the dirty chunks (their searches are done by now), then the upper levels and the means
__global__ void __launch_bounds__(256, 4)
lg_passA_kernel(const float4* __restrict__ cur, const float4* __restrict__ tg, const unsigned char* __restrict__ dirty, double max2,
                int n, int nchunks, float* __restrict__ partA /* [6][S] */, double* __restrict__ partD, int* __restrict__ partK,
                LgTree tr, LgState* __restrict__ st,
                volatile int* __restrict__ h_unres /* pinned host word: how many queries the refine kernel left over */) {
    lg_wait_prior();
    if (st->done) return;
    __shared__ unsigned s_mine;
    if (threadIdx.x == 0) s_mine = 0u;
    __syncthreads();
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c < nchunks && dirty[c]) {               // (the clean chunks arrived in lg_track_kernel)
        float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        double dsum = 0.0; int k = 0;
#pragma unroll
        for (int half = 0; half < 2; ++half) {   // two halves of four points per lane: every CTA of the pass is resident at once
            float4 cv[4], gv[4], sv[4], tv[4];
            lg_load_half(cur, tg, n, c, half, cv, gv);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                tv[u].w = __int_as_float(-1);
                if (__float_as_int(gv[u].w) >= 0) lg_pair(cv[u], gv[u], max2, sv[u], tv[u]);
            }
            lg_accum_A<4>(sv, tv, a, dsum, k);
        }
        lg_store_A(a, dsum, k, c, tr.S, partA, partD, partK);
        if ((threadIdx.x & 31) == 0) atomicAdd(&s_mine, 1u);
    }
    __syncthreads();
    const int g = lg_group_arrive(tr.grpcnt, nchunks, s_mine);
    if (g < 0) return;
    lg_group_A(partA, partD, partK, tr, nchunks, g);
    if (!lg_groups_done(tr.grpcnt, &st->ticketA, (nchunks + 255) >> 8)) return;
    lg_finish_A(nchunks, tr, st, h_unres);
}
__global__ void __launch_bounds__(256, 4)
lg_passB_kernel(const float4* __restrict__ cur, const float4* __restrict__ tg, double max2, int n, int nchunks,
                float* __restrict__ partB /* [9][S] */, LgTree tr,
                LgState* __restrict__ st, int max_iter, double rot_thr, double trans_thr, double mse_rel, double mse_abs) {
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5);
    float4 cv[4], gv[4];
    if (c < nchunks) lg_load_half(cur, tg, n, c, 0, cv, gv);          // (pass A does not write them: before the wait)
    lg_wait_prior();
    if (st->done) return;
    const bool enough = st->kept >= 3;
#ifdef KSS_LG_TIMELINE
    const unsigned long long tl0 = lg_now();
    unsigned long long tl1 = tl0, tl2 = tl0;
#endif
    if (c < nchunks && enough) {
        const float sm0 = st->smean[0], sm1 = st->smean[1], sm2 = st->smean[2], dm0 = st->dmean[0], dm1 = st->dmean[1], dm2 = st->dmean[2];
        float a[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        lg_accum_B(cv, gv, max2, sm0, sm1, sm2, dm0, dm1, dm2, a);
        lg_load_half(cur, tg, n, c, 1, cv, gv);
        lg_accum_B(cv, gv, max2, sm0, sm1, sm2, dm0, dm1, dm2, a);
#ifdef KSS_LG_TIMELINE
        tl1 = lg_now();
#endif
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1)
#pragma unroll
            for (int q = 0; q < 9; ++q) a[q] = __fadd_rn(a[q], __shfl_xor_sync(KSS_FULL, a[q], off));
        if ((threadIdx.x & 31) == 0)
#pragma unroll
            for (int q = 0; q < 9; ++q) partB[(size_t)q * tr.S + c] = a[q];
#ifdef KSS_LG_TIMELINE
        tl2 = lg_now();
#endif
    }
    const int g = lg_group_arrive(tr.grpcnt, nchunks, (unsigned)min(8, nchunks - (int)blockIdx.x * 8));
#ifdef KSS_LG_TIMELINE
    if (threadIdx.x == 0) {
        const unsigned long long tl3 = lg_now();
        (void)tl3;
    }
#endif
    if (g < 0) return;
    if (enough) lg_group_f32<9>(partB, tr.S, nchunks, g, tr.l2f, tr.S2);
    if (!lg_groups_done(tr.grpcnt, &st->ticketB, (nchunks + 255) >> 8)) return;
    lg_finish_B(nchunks, tr, st, enough, max_iter, rot_thr, trans_thr, mse_rel, mse_abs);
}

// state of a run from the sorted source: positions by original index without a certificate, no matches, the two maps
__global__ void __launch_bounds__(256)
lg_run_init_kernel(const float4* __restrict__ inp_s, int n, float4* __restrict__ cur, float4* __restrict__ tg,
                   int* __restrict__ perm, int* __restrict__ inv, unsigned char* __restrict__ flagS) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n) return;
    const float4 p = inp_s[pos];
    const int i = __float_as_int(p.w);
    perm[pos] = i; inv[i] = pos; flagS[pos] = 0;
    cur[i] = make_float4(p.x, p.y, p.z, 0.0f);
    tg[i] = make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1));
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
        st->n_unres = 0u; st->n_fail = 0u;
    }
}

// ------------------------------------------------------------------ host orchestration
namespace {

struct Ctx {
    cudaStream_t st; long long* launches; const DevAlloc& alloc; const char* pfx; int err = 0;
    template <class T> T* get(const char* name, size_t count) {
        void* p = nullptr;
        char nm[96]; snprintf(nm, sizeof(nm), "%s%s", pfx, name);
        int r = alloc(nm, count * sizeof(T), &p);
        if (r != KSS_OK) { err = r; return nullptr; }
        return reinterpret_cast<T*>(p);
    }
    bool ok() { if (cudaGetLastError() != cudaSuccess) err = KSS_ERR_CUDA; return err == 0; }
    void launched(int k = 1) { *launches += k; }
};

float grid_hc() {                       // cell size in units of the estimated point spacing (A/B switch)
    const char* e = getenv("KSS_LG_HC");
    const float v = e ? (float)atof(e) : 0.0f;
    return v > 0.0f ? v : 2.0f;
}

// counting sort of a cloud by (Morton block, fine cell) of the grid `geom`: scratch is shared between the target and
// the query cloud (stream order), only blk_rank / fine_start / n_occ are per cloud
struct GridScratch {
    size_t nbmax; int max_bits;
    unsigned *pkey, *blk_cnt, *blk_base, *fine_cnt;
    unsigned long long *blk_scan, *blk_tot;
};
int grid_scratch(Ctx& c, int n_max, int max_bits, GridScratch* s) {
    s->max_bits = max_bits; s->nbmax = (size_t)1 << max_bits;
    const size_t occ_max = std::min((size_t)n_max, s->nbmax);
    s->pkey = c.get<unsigned>("lg_pkey", n_max);
    s->blk_cnt = c.get<unsigned>("lg_blk_cnt", s->nbmax);
    s->blk_scan = c.get<unsigned long long>("lg_blk_scan", s->nbmax);
    s->blk_tot = c.get<unsigned long long>("lg_blk_tot", s->nbmax / 1024 + 1);
    s->blk_base = c.get<unsigned>("lg_blk_base", occ_max + 1);
    s->fine_cnt = c.get<unsigned>("lg_fine_cnt", occ_max * 64);
    return c.err;
}
void grid_sort(Ctx& c, const GridScratch& s, const LgGeom* geom, const float4* p4, int n, unsigned* blk_rank, unsigned* fine_start,
               int* n_occ, float4* sorted) {
    const int gp = (n + 255) / 256;
    const int nsb = (int)(s.nbmax / 1024);
    cudaMemsetAsync(s.blk_cnt, 0, sizeof(unsigned) * s.nbmax, c.st);
    lg_bin_kernel<<<gp, 256, 0, c.st>>>(p4, n, geom, s.pkey, s.blk_cnt);
    lg_bscan1_kernel<<<nsb, 1024, 0, c.st>>>(s.blk_cnt, geom, s.blk_scan, s.blk_tot);
    lg_bscan2_kernel<<<1, 1024, 0, c.st>>>(s.blk_tot, geom, n_occ);
    lg_bscan3_kernel<<<nsb, 1024, 0, c.st>>>(s.blk_cnt, s.blk_scan, s.blk_tot, geom, blk_rank, s.blk_base);
    lg_zero_fine_kernel<<<148 * 4, 256, 0, c.st>>>(n_occ, reinterpret_cast<uint4*>(s.fine_cnt));
    lg_fine_hist_kernel<<<gp, 256, 0, c.st>>>(s.pkey, n, blk_rank, s.fine_cnt);
    lg_fine_scan_kernel<<<148 * 4, 256, 0, c.st>>>(n_occ, s.blk_base, s.fine_cnt, fine_start, n);
    lg_grid_scatter_kernel<<<gp, 256, 0, c.st>>>(p4, n, s.pkey, blk_rank, fine_start, s.fine_cnt, sorted);
    c.launched(8);
}
inline int grid_max_bits(int n_t) { return std::min(22, std::max(12, lg_ceil_log2(std::max(1, n_t)) + 2)); }

// a query cloud in the order of the TARGET's grid (clamped into it): a CTA's 512 consecutive queries then touch a
// handful of blocks.  p4 = by original index, sorted: w = original index.  Order never affects results.
int order_queries(Ctx& c, const GridScratch& s, const LgGridView& gv, const double* d_pts, int n, float4* p4, float4* sorted) {
    const size_t occ_max = std::min((size_t)n, s.nbmax);
    unsigned* bb = c.get<unsigned>("lg_q_bb", 8);
    unsigned* blk_rank = c.get<unsigned>("lg_q_blk_rank", s.nbmax);
    unsigned* fine_start = c.get<unsigned>("lg_q_fine_start", occ_max * 64 + 68);
    int* n_occ = c.get<int>("lg_q_nocc", 1);
    if (c.err) return c.err;
    lg_init_bbox_kernel<<<1, 32, 0, c.st>>>(bb, nullptr);
    lg_convert_bbox_kernel<<<std::min((n + 255) / 256, 148 * 8), 256, 0, c.st>>>(d_pts, n, p4, bb);
    c.launched(2);
    grid_sort(c, s, gv.geom, p4, n, blk_rank, fine_start, n_occ, sorted);
    return c.ok() ? KSS_OK : c.err;
}

// block grid + pyramid of the target, everything enqueued, no host synchronisation
int build_target(Ctx& c, const double* d_t, int n_t, int n_q_max, Pyramid* py, LgGridView* gv, GridScratch* scr) {
    const int npad = (n_t + 31) / 32 * 32;
    const int max_bits = grid_max_bits(n_t);
    if (grid_scratch(c, std::max(n_t, n_q_max), max_bits, scr)) return c.err;
    const size_t occ_max = std::min((size_t)n_t, scr->nbmax);
    float4* t_orig = c.get<float4>("lg_t_orig", n_t);
    float4* tp = c.get<float4>("lg_tp", npad);
    unsigned* bb = c.get<unsigned>("lg_t_bb", 8);
    float* sample = c.get<float>("lg_sample", LG_SAMPLES);
    LgGeom* geom = c.get<LgGeom>("lg_geom", 1);
    unsigned* blk_rank = c.get<unsigned>("lg_blk_rank", scr->nbmax);
    unsigned* fine_start = c.get<unsigned>("lg_fine_start", occ_max * 64 + 68);
    int* n_occ = c.get<int>("lg_t_nocc", 1);
    unsigned* lut = c.get<unsigned>("lg_lut", 3 * LG_LUT);
    if (c.err) return c.err;
    lg_init_bbox_kernel<<<1, 256, 0, c.st>>>(bb, sample);
    lg_convert_bbox_kernel<<<std::min((n_t + 255) / 256, 148 * 8), 256, 0, c.st>>>(d_t, n_t, t_orig, bb);
    lg_probe_kernel<<<(n_t + 1023) / 1024, LG_SAMPLES, 0, c.st>>>(t_orig, n_t, sample);
    lg_geom_kernel<<<1, LG_SAMPLES, 0, c.st>>>(bb, sample, n_t, grid_hc(), max_bits, geom);
    lg_lut_kernel<<<1, 256, 0, c.st>>>(geom, lut);
    c.launched(5);
    grid_sort(c, *scr, geom, t_orig, n_t, blk_rank, fine_start, n_occ, tp);
    if (npad > n_t) { lg_pad_kernel<<<1, 32, 0, c.st>>>(tp, n_t, npad); c.launched(); }
    // box pyramid over 32-point tiles of the sorted array
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
    gv->geom = geom; gv->blk_rank = blk_rank; gv->fine_start = fine_start; gv->tp = tp; gv->t_orig = t_orig; gv->lut = lut; gv->knn = nullptr; gv->n_t = n_t;
    if (getenv("KSS_LG_VERBOSE")) {
        LgGeom h; int no = 0;
        cudaMemcpyAsync(&h, geom, sizeof(h), cudaMemcpyDeviceToHost, c.st);
        cudaMemcpyAsync(&no, n_occ, sizeof(int), cudaMemcpyDeviceToHost, c.st);
        cudaStreamSynchronize(c.st);
        fprintf(stderr, "[lg] n_t=%d h=%g fine %d,%d,%d blocks %d,%d,%d (table %d) occupied %d -> %.1f points / block\n", n_t, h.h,
                h.nf[0], h.nf[1], h.nf[2], h.nb[0], h.nb[1], h.nb[2], h.nbp, no, no ? (double)n_t / no : 0.0);
    }
    return c.ok() ? KSS_OK : c.err;
}

inline int nn_grid(int n_q) { return (n_q + NN_QPC - 1) / NN_QPC; }

// launch with programmatic stream serialisation (see lg_wait_prior)
template <class... KArgs, class... Args>
void launch_pdl(bool pdl, void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)block); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

template <int MODE>
void launch_nn(cudaStream_t st, const Pyramid& py, const LgGridView& gv, int n_q, float4* cur_s, const float4* inp_s, int* idx,
               float* d2, float4* tg, const int* perm, unsigned char* flagS, LgState* state) {
    cudaFuncSetAttribute(lg_nn_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NN_SMEM);
    lg_nn_kernel<MODE><<<nn_grid(n_q), NN_THREADS, NN_SMEM, st>>>(py, gv, n_q, cur_s, inp_s, idx, d2, tg, perm, flagS, state);
}

}  // namespace

int large_nn_device(cudaStream_t st, long long* launches, const double* d_q, int n_q, const double* d_t, int n_t,
                    int* d_idx, float* d_d2, const DevAlloc& alloc) {
    Ctx c{st, launches, alloc, ""};
    Pyramid py; LgGridView gv; GridScratch scr;
    int r = build_target(c, d_t, n_t, n_q, &py, &gv, &scr);
    if (r) return r;
    float4* q4 = c.get<float4>("lg_q4", n_q);
    float4* qs = c.get<float4>("lg_qs", n_q);
    if (c.err) return c.err;
    r = order_queries(c, scr, gv, d_q, n_q, q4, qs);
    if (r) return r;
    launch_nn<0>(st, py, gv, n_q, qs, nullptr, d_idx, d_d2, nullptr, nullptr, nullptr, nullptr);
    c.launched();
    return c.ok() ? KSS_OK : c.err;
}

int large_metrics_device(cudaStream_t st, long long* launches, const double* d_q, int n_q, const double* d_t, int n_t,
                         double* d_out3, const DevAlloc& alloc) {
    Ctx c{st, launches, alloc, ""};
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
    Ctx c{st, launches, alloc, ""};
    double* d_s = c.get<double>("lg_in_s", (size_t)n_s * 3);
    double* d_t = c.get<double>("lg_in_t", (size_t)n_t * 3);
    if (c.err) return c.err;
    cudaMemcpyAsync(d_s, src, sizeof(double) * 3 * (size_t)n_s, cudaMemcpyHostToDevice, st);
    cudaMemcpyAsync(d_t, tgt, sizeof(double) * 3 * (size_t)n_t, cudaMemcpyHostToDevice, st);
    LargeIcp run;
    int r = large_icp_prepare(st, launches, d_s, n_s, d_t, n_t, alloc, &run, "");
    if (r) return r;
    r = large_icp_run(st, launches, &run, prm, 0);
    if (!r) r = large_icp_result(st, &run, T, fitness, iters, converged);
    if (run.h_unres) { cudaStreamSynchronize(st); cudaFreeHost(run.h_unres); }
    return r;
}

// ---- reusable pieces (bench.py times large_icp_iterations on a prepared run)
int large_icp_prepare(cudaStream_t st, long long* launches, const double* d_s, int n_s, const double* d_t, int n_t,
                      const DevAlloc& alloc, LargeIcp* run, const char* prefix) {
    Ctx c{st, launches, alloc, prefix ? prefix : ""};
    Pyramid py; LgGridView gv; GridScratch scr;
    int r = build_target(c, d_t, n_t, n_s, &py, &gv, &scr);
    if (r) return r;
    static_assert(sizeof(LgGridView) <= sizeof(run->grid), "LargeIcp::grid too small");
    memcpy(run->grid, &gv, sizeof(gv));
    float4* q4 = c.get<float4>("lg_q4", n_s);                 // by original index (scratch of the sort)
    float4* inp = c.get<float4>("lg_inp", n_s);               // sorted input (fitness pass), w = original index
    float4* cur = c.get<float4>("lg_cur", n_s);               // by original index: position (transformed in place every iteration), certificate
    float4* tg = c.get<float4>("lg_tg", n_s);                 // by original index: matched target, its index (-1: none yet)
    float4* cert = c.get<float4>("lg_cert", n_s);             // by original index: where and with which bound the certificate was established
    float4* tg2 = c.get<float4>("lg_tg2", n_s);               // by original index: the runner-up, where the certificate names it (cur.w < 0)
    float* d2 = c.get<float>("lg_d2", n_s);
    int* perm = c.get<int>("lg_perm", n_s);                   // sorted position -> original index
    int* inv = c.get<int>("lg_inv", n_s);                     // original index -> sorted position
    unsigned char* flagS = c.get<unsigned char>("lg_flagS", n_s);     // by sorted position: to be searched
    int* worklist = c.get<int>("lg_worklist", n_s);
    const int nchunks = (n_s + 255) / 256;
    unsigned char* dirty = c.get<unsigned char>("lg_dirty", nchunks);
    float4* knn = c.get<float4>("lg_knn", (size_t)n_t * LG_KSLOTS);  // neighbour lists of the target, by original index
    const int S = (nchunks + 31) / 32 * 32, n2 = (nchunks + 255) / 256, S2 = (n2 + 31) / 32 * 32;
    float* pA = c.get<float>("lg_partA", (size_t)S * 6);
    float* pB = c.get<float>("lg_partB", (size_t)S * 9);
    double* pD = c.get<double>("lg_partD", S);
    int* pK = c.get<int>("lg_partK", S);
    float* l2f = c.get<float>("lg_l2f", (size_t)S2 * 9);
    double* l2d = c.get<double>("lg_l2d", S2);
    int* l2k = c.get<int>("lg_l2k", S2);
    unsigned* grpcnt = c.get<unsigned>("lg_grpcnt", S2);
    LgState* state = c.get<LgState>("lg_state", 1);
    double* out3 = c.get<double>("lg_out3", 4);
    if (c.err) return c.err;
    r = order_queries(c, scr, gv, d_s, n_s, q4, inp);
    if (r) return r;
    if (n2 > 256) return KSS_ERR_UNSUPPORTED;                     // three reduction levels: n <= 16.7 M source points
#ifdef KSS_LG_TIMELINE
    { unsigned long long z[16]; for (int i = 0; i < 16; ++i) z[i] = 0ull; cudaMemcpyToSymbol(lg_tl, z, sizeof(z)); }
#endif
    cudaMemsetAsync(grpcnt, 0, sizeof(unsigned) * S2, st);
    lg_run_init_kernel<<<(n_s + 255) / 256, 256, 0, st>>>(inp, n_s, cur, tg, perm, inv, flagS);
    lg_knn_kernel<<<(n_t + 127) / 128, 128, 0, st>>>(gv, knn);
    gv.knn = knn;
    memcpy(run->grid, &gv, sizeof(gv));
    c.launched(2);
    if (!run->h_unres && cudaHostAlloc((void**)&run->h_unres, sizeof(int), cudaHostAllocMapped) != cudaSuccess) { cudaGetLastError(); run->h_unres = nullptr; }
    if (run->h_unres) *run->h_unres = 0x7fffffff;             // nothing known yet: the first iterations stage
    lg_state_init_kernel<<<1, 32, 0, st>>>(state);
    c.launched();
    static_assert(sizeof(Pyramid) <= sizeof(run->pyramid), "LargeIcp::pyramid too small");
    memcpy(run->pyramid, &py, sizeof(py));
    run->n_s = n_s; run->n_t = n_t; run->nchunks = nchunks;
    run->inp = inp; run->cur = cur; run->tg = tg; run->tg2 = tg2; run->cert = cert; run->d2 = d2; run->perm = perm; run->inv = inv; run->flagS = flagS;
    run->dirty = dirty; run->worklist = worklist;
    run->partA = pA; run->partB = pB; run->partD = pD; run->partK = pK; run->state = state; run->out3 = out3;
    run->l2f = l2f; run->l2d = l2d; run->l2k = l2k; run->grpcnt = grpcnt; run->S = S; run->S2 = S2;
    return c.ok() ? KSS_OK : c.err;
}

// enqueue `count` ICP iterations (each = NN + pass A + pass B); kernels are no-ops once the run is done
int large_icp_iterations(cudaStream_t st, long long* launches, LargeIcp* run, const kss_icp_params* prm, int count) {
    Pyramid py; memcpy(&py, run->pyramid, sizeof(py));
    LgGridView gv; memcpy(&gv, run->grid, sizeof(gv));
    LgState* state = (LgState*)run->state;
    const double max2 = prm->max_corr_dist * prm->max_corr_dist;
    const int n = run->n_s, nch = run->nchunks;
    const double mse_abs = prm->fitness_eps < 0.0 ? -1.0 : 1e-12;   // fitness_eps < 0: never converge (steady-state timing)
    static const float margin = [] { const char* e = getenv("KSS_LG_MARGIN"); return e ? (float)atof(e) : 0.25f; }();  // certificate margin, in cells
    float4 *cur = (float4*)run->cur, *tg = (float4*)run->tg;
    const LgTree tr{run->grpcnt, run->l2f, run->l2d, run->l2k, run->S, run->S2};
    static const bool pdl = !getenv("KSS_LG_NO_PDL");
    for (int k = 0; k < count; ++k) {
        if (run->mark) run->mark(run->mark_user, KSS_STAGE_LARGE_NN, 1);
        // correspondences: the streaming pass keeps every match whose certificate holds; the refine kernel searches the
        // rest around their last match; the general kernels take what it leaves (staged CTAs when many, one warp per
        // query when few).  (The host knows the left-over count of an EARLIER iteration from a pinned word passA writes:
        // it only decides whether the staged kernel is worth a launch; the one-warp-per-query kernel takes any count.)
        const int known = run->h_unres ? *(volatile int*)run->h_unres : 0x7fffffff;
        const bool staged = known > n / (int)LG_STAGED_DIV;
        static const int refine_g = [] { const char* e = getenv("KSS_LG_REFINE_G"); return e ? atoi(e) : 0; }();    // A/B: lanes per query of lg_refine_kernel (1 or 4; default: by the count)
        if (run->mark) run->mark(run->mark_user, KSS_STAGE_LARGE_TRACK, 1);
        launch_pdl(pdl, lg_track_kernel, (nch + 7) / 8, 256, 0, st, n, nch, cur, tg, (float4*)run->tg2, (float4*)run->cert, run->inv, gv.knn, run->flagS, run->dirty, run->partA,
                   run->partD, run->partK, tr, state, max2, (volatile int*)run->h_unres);
        if (run->mark) run->mark(run->mark_user, KSS_STAGE_LARGE_TRACK, 0);
        launch_pdl(pdl, lg_refine_kernel, (n + RF_SEG - 1) / RF_SEG, 256, 0, st, gv, n, cur, tg, (float4*)run->tg2, (float4*)run->cert, run->perm, run->flagS, run->worklist, state, margin, refine_g);
        if (staged) {
            cudaFuncSetAttribute(lg_nn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NN_SMEM);
            launch_pdl(pdl, lg_nn_kernel<1>, nn_grid(n), NN_THREADS, NN_SMEM, st, py, gv, n, cur, (const float4*)nullptr, (int*)nullptr,
                       (float*)nullptr, tg, run->perm, run->flagS, state);
        }
        launch_pdl(pdl, lg_left_kernel, 148 * 4, 256, 0, st, py, gv, n, cur, tg, (float4*)run->tg2, (float4*)run->cert, run->perm, run->flagS, run->worklist, state, staged ? 1 : 0);
        if (run->mark) { run->mark(run->mark_user, KSS_STAGE_LARGE_NN, 0); run->mark(run->mark_user, KSS_STAGE_LARGE_REDUCE, 1); }
        launch_pdl(pdl, lg_passA_kernel, (nch + 7) / 8, 256, 0, st, (const float4*)cur, (const float4*)tg, run->dirty, max2, n, nch, run->partA,
                   run->partD, run->partK, tr, state, (volatile int*)run->h_unres);
        launch_pdl(pdl, lg_passB_kernel, (nch + 7) / 8, 256, 0, st, (const float4*)cur, (const float4*)tg, max2, n, nch, run->partB, tr, state,
                   prm->max_iterations, 1.0 - prm->transformation_eps, prm->transformation_eps, prm->fitness_eps, mse_abs);
        if (run->mark) run->mark(run->mark_user, KSS_STAGE_LARGE_REDUCE, 0);
        *launches += staged ? 6 : 5;
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
    LgGridView gv; memcpy(&gv, run->grid, sizeof(gv));
    launch_nn<2>(st, py, gv, run->n_s, nullptr, (const float4*)run->inp, nullptr, run->d2, nullptr, nullptr, nullptr, state);
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
    if (getenv("KSS_LG_VERBOSE")) {
        unsigned d[8]; cudaMemcpyFromSymbol(d, lg_dbg, sizeof(d));
        unsigned d2[4]; cudaMemcpyFromSymbol(d2, lg_dbg2, sizeof(d2));
        fprintf(stderr, "[lg] third stage: no bound %u, cube too large %u, failed %u, resolved %u\n", d2[0], d2[1], d2[2], d2[3]);
        fprintf(stderr, "[lg] unstaged half-passes by reason: box %u, touched blocks %u, points %u, occupied %u; sums: points %u occupied %u touched %u box %u\n",
                d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]);
    }
#ifdef KSS_LG_TIMELINE
    { unsigned long long t[16]; cudaMemcpyFromSymbol(t, lg_tl, sizeof(t));
      fprintf(stderr, "[lg] passB timeline (first CTA start of ANY launch is the min; use a 1-iteration run): first CTA -> last group %.2f us, level 3 %.2f us, SVD etc. %.2f us\n",
              (t[1] - t[0]) * 1e-3, (t[2] - t[1]) * 1e-3, (t[3] - t[2]) * 1e-3);
      fprintf(stderr, "[lg] track per CTA after iteration 400 (thread 0): start->wait+T %.2f us, chunk work %.2f us [wait for loads %.2f, loop %.2f, sums %.2f], arrive %.2f us (%llu CTA-launches)\n",
              t[8] * 1e-3 / t[11], t[9] * 1e-3 / t[11], (t[1] + 1) * 1e-3 / t[11], (t[2] + 1) * 1e-3 / t[11], (t[3] + 1) * 1e-3 / t[11], t[10] * 1e-3 / t[11], t[11]);
      unsigned long long z[16]; for (int i = 0; i < 16; ++i) z[i] = ~0ull; cudaMemcpyToSymbol(lg_tl, z, sizeof(z)); }
#endif
#ifdef KSS_LG_COUNT
    { unsigned long long c4[4]; cudaMemcpyFromSymbol(c4, lg_cnt, sizeof(c4));
      fprintf(stderr, "[lg] first stage: %.2f candidates per query, %.2f lane slots per query (over %llu query slots)\n", (double)c4[0] / c4[2], (double)c4[1] / c4[2], c4[2]); }
#endif
    if (getenv("KSS_LG_VERBOSE"))
        fprintf(stderr, "[lg] over the run: %u second-stage searches, %u pyramid fallbacks, %u CTAs on global tables (iterations %d)\n",
                h.miss[1], h.miss[2], h.miss[3], h.iters);
    if (T) for (int i = 0; i < 16; ++i) T[i] = h.fin[i];
    if (fitness) *fitness = o3[0];
    if (iters) *iters = h.iters;
    if (converged) *converged = h.converged;
    return KSS_OK;
}

}  // namespace kss
