// kss_aivs.cu -- the simplification that feeds the registration path, batched on the GPU:
//   pointPipeline_Border                      pointPipeline.hpp:105-158
//   BallRegion_EstimateBoxScale / _AchieveXYZ ballRegionCompute.hpp:1194-1214, 690-758
//   BallRegion_ReturnBoxCenter_Center         ballRegionCompute.hpp:1150-1172   (keeps its missing y wrap, B10)
//   BallRegion_BoxInput                       ballRegionCompute.hpp:632-688     (1-based boxes)
//   BallRegion_ReturnNeiborBox_Box            ballRegionCompute.hpp:975-1040    (keeps its missing x wrap, B10)
//   AIVS_initBoxIndexNumber                   Method_AIVS_SimPro.hpp:587-643    (8 colours = 8 launches)
//   AIVS_BoxSimplification_Points             Method_AIVS_SimPro.hpp:776-794
//   AIVS_Voroni_OpenMP_KNN                    Method_AIVS_SimPro.hpp:222-376    (per-box farthest point sampling)
//   AIVS_AccurateCut_Optimization             Method_AIVS_SimPro.hpp:848-957    (greedy closest-pair trim)
// Layout: clouds [P][cap][3] doubles with per-cloud counts; per cloud a 1-based box grid of at most `bcap`
// boxes; points grouped by box in ascending original index (the order the reference's push_back produces).
// Same-colour boxes never see each other's points, so one thread per box and one launch per colour reproduces
// the reference's sequential-by-colour semantics exactly.  The kd-tree searches of the reference are exact and
// only feed minima, so plain scans give identical values; float d2 on float-narrowed coordinates, float sqrt.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "kss_device.cuh"
#include "kss_aivs.h"

namespace kss {

struct AivsGrid {            // per cloud
    double mn[3];
    double unit;
    int nx, ny, nz, nbox;    // nbox = nx*ny*nz + 1 (box 0 is never used)
    int n, point_num;
    int n_samples, pad;
};

__device__ __forceinline__ int aivs_box_scale(int n) {
    if (n < 10000) return 10;
    if (n < 50000) return 20;
    if (n < 100000) return 30;
    if (n < 500000) return 40;
    if (n < 1000000) return 50;
    // (int)pow(n / 8.0, 1.0 / 3.0) with glibc/MSVC pow: 1.0/3.0 is below one third, so a perfect cube comes out one
    // ulp low and truncates to b - 1 (49 at n = 1e6); everything else is floor(cbrt)
    const double x = (double)n / 8.0;
    long long b = (long long)cbrt(x);
    while ((double)(b + 1) * (double)(b + 1) * (double)(b + 1) <= x) ++b;
    while ((double)b * (double)b * (double)b > x) --b;
    if ((double)b * (double)b * (double)b == x) --b;
    return (int)b;
}

__device__ __forceinline__ void aivs_center(const AivsGrid& g, int box, double c[3]) {
    int z_num = box / (g.nx * g.ny) + 1;
    const int leveZ = box % (g.nx * g.ny);
    int y_num = leveZ / g.nx + 1;
    int x_num = leveZ % g.nx;
    if (x_num == 0) { x_num = g.nx; y_num = y_num - 1; }
    c[0] = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(g.mn[0], __dmul_rn((double)(x_num - 1), g.unit)), g.mn[0]), __dmul_rn((double)x_num, g.unit)), 2.0);
    c[1] = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(g.mn[1], __dmul_rn((double)(y_num - 1), g.unit)), g.mn[1]), __dmul_rn((double)y_num, g.unit)), 2.0);
    c[2] = __ddiv_rn(__dadd_rn(__dadd_rn(__dadd_rn(g.mn[2], __dmul_rn((double)(z_num - 1), g.unit)), g.mn[2]), __dmul_rn((double)z_num, g.unit)), 2.0);
}

// ---------------------------------------------------------------- 1. border (many CTAs per cloud), grid geometry
__device__ __forceinline__ unsigned long long aivs_ord(double v) {          // monotone map double -> u64
    unsigned long long a = (unsigned long long)__double_as_longlong(v);
    return (a >> 63) ? ~a : (a | 0x8000000000000000ull);
}
__device__ __forceinline__ double aivs_unord(unsigned long long a) {
    a = (a >> 63) ? (a & 0x7fffffffffffffffull) : ~a;
    return __longlong_as_double((long long)a);
}
// ext[p][0..2] = min keys (initialised to all ones), ext[p][3..5] = max keys (initialised to zero)
__global__ void __launch_bounds__(256)
aivs_extent_kernel(const double* __restrict__ pts, const int* __restrict__ cnt, int cap, unsigned long long* __restrict__ ext) {
    const int p = blockIdx.y;
    const int n = cnt ? cnt[p] : cap;
    const double* P = pts + (size_t)p * cap * 3;
    double lmin[3] = {INFINITY, INFINITY, INFINITY}, lmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int d = 0; d < 3; ++d) { const double v = P[3 * (size_t)i + d]; lmin[d] = fmin(lmin[d], v); lmax[d] = fmax(lmax[d], v); }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        unsigned long long a = aivs_ord(lmin[d]), b = aivs_ord(lmax[d]);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const unsigned long long oa = __shfl_xor_sync(KSS_FULL, a, off), ob = __shfl_xor_sync(KSS_FULL, b, off);
            a = oa < a ? oa : a; b = ob > b ? ob : b;
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(&ext[(size_t)p * 6 + d], a); atomicMax(&ext[(size_t)p * 6 + 3 + d], b); }
    }
}
__global__ void __launch_bounds__(128)
aivs_grid_kernel(const unsigned long long* __restrict__ ext, const int* __restrict__ cnt, int cap, const int* __restrict__ point_num,
                 int point_num_all, int bcap, int P, AivsGrid* __restrict__ grids, int* __restrict__ bad) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int n = cnt ? cnt[p] : cap;
    AivsGrid g;
    g.n = n; g.point_num = point_num ? point_num[p] : point_num_all;
    const int boxNum = aivs_box_scale(n);
    double dis[3];
    for (int d = 0; d < 3; ++d) {
        const double vmin = aivs_unord(ext[(size_t)p * 6 + d]), vmax = aivs_unord(ext[(size_t)p * 6 + 3 + d]);
        g.mn[d] = vmin; dis[d] = fabs(__dsub_rn(vmax, vmin));
    }
    double large = dis[0];
    if (large < dis[1]) large = dis[1];
    if (large < dis[2]) large = dis[2];
    g.unit = __ddiv_rn(large, (double)boxNum);
    int num[3];
    for (int d = 0; d < 3; ++d) {
        const double nd = __ddiv_rn(dis[d], g.unit);
        num[d] = (int)nd;
        if (nd > (double)num[d]) num[d]++;
    }
    g.nx = num[0]; g.ny = num[1]; g.nz = num[2];
    const long long nb = (long long)g.nx * g.ny * g.nz + 1;
    g.nbox = (int)nb; g.n_samples = 0; g.pad = 0;
    if (!(large > 0.0) || nb > bcap || g.nx < 1 || g.ny < 1 || g.nz < 1) { g.nbox = 0; atomicExch(bad, 1); }   // degenerate cloud
    grids[p] = g;
}

__device__ __forceinline__ int aivs_box_of(const AivsGrid& g, const double* q) {
    int id[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double v = __ddiv_rn(__dsub_rn(q[d], g.mn[d]), g.unit);
        id[d] = (int)v;
        if ((double)id[d] < v || id[d] == 0) id[d]++;
    }
    return id[0] + g.nx * (id[1] - 1) + (g.nx * g.ny * (id[2] - 1));
}

// ---------------------------------------------------------------- 2. box membership: count, scan, fill, sort
__global__ void __launch_bounds__(256)
aivs_count_kernel(const double* __restrict__ pts, int cap, const AivsGrid* __restrict__ grids, int bcap,
                  int* __restrict__ box_of, int* __restrict__ box_cnt) {
    const int p = blockIdx.y;
    const AivsGrid g = grids[p];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n || g.nbox == 0) return;
    int b = aivs_box_of(g, pts + ((size_t)p * cap + i) * 3);
    if (b < 0 || b >= g.nbox) b = -1;                       // the reference only prints "Hello!" here (UB); dropped
    box_of[(size_t)p * cap + i] = b;
    if (b >= 0) atomicAdd(&box_cnt[(size_t)p * bcap + b], 1);
}

// exclusive scan of the box counts of one cloud (one CTA per cloud; nbox <= ~125k)
__global__ void __launch_bounds__(1024)
aivs_scan_kernel(const AivsGrid* __restrict__ grids, int bcap, const int* __restrict__ box_cnt, int* __restrict__ box_start,
                 int* __restrict__ cursor) {
    __shared__ int ws[32];
    __shared__ int carry;
    const int p = blockIdx.x;
    const int nb = grids[p].nbox;
    const int* c = box_cnt + (size_t)p * bcap;
    int* s = box_start + (size_t)p * (bcap + 1);
    int* cu = cursor ? cursor + (size_t)p * bcap : nullptr;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? c[i] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(KSS_FULL, x, o); if (lane >= o) x += y; }
        if (lane == 31) ws[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int w = ws[lane];
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(KSS_FULL, w, o); if (lane >= o) w += y; }
            ws[lane] = w;
        }
        __syncthreads();
        const int incl = x + (warp ? ws[warp - 1] : 0) + carry;
        if (i < nb) { s[i] = incl - v; if (cursor) cu[i] = incl - v; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) s[nb] = carry;
}

__global__ void __launch_bounds__(256)
aivs_fill_kernel(int cap, const AivsGrid* __restrict__ grids, int bcap, const int* __restrict__ box_of,
                 int* __restrict__ cursor, int* __restrict__ members) {
    const int p = blockIdx.y;
    const AivsGrid g = grids[p];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n || g.nbox == 0) return;
    const int b = box_of[(size_t)p * cap + i];
    if (b < 0) return;
    members[(size_t)p * cap + atomicAdd(&cursor[(size_t)p * bcap + b], 1)] = i;
}

// per box (one WARP): the centre-nearest member and the quota.  The reference keeps a box's members in ascending
// original index (push_back order) and breaks every tie by position; the member lists here are in the arbitrary order
// of the atomic fill, and every tie is broken by the ORIGINAL INDEX instead -- the same choice, without a sort.
__global__ void __launch_bounds__(256)
aivs_box_kernel(const double* __restrict__ pts, int cap, const AivsGrid* __restrict__ grids, int bcap,
                const int* __restrict__ box_start, const int* __restrict__ members, int* __restrict__ center_pos,
                int* __restrict__ quota) {
    const int p = blockIdx.y;
    const AivsGrid g = grids[p];
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= g.nbox) return;
    const int* st = box_start + (size_t)p * (bcap + 1);
    const int s = st[b], m = st[b + 1] - s;
    const int* mem = members + (size_t)p * cap + s;
    double c[3];
    aivs_center(g, b, c);
    double best = 9999.0; int bi = -1, bidx = 0x7fffffff;
    const double* P = pts + (size_t)p * cap * 3;
    for (int i = lane; i < m; i += 32) {                       // BoxInput: first strict minimum of the distance to the centre
        const int pi = mem[i];
        const double* q = P + 3 * (size_t)pi;
        const double dx = __dsub_rn(c[0], q[0]), dy = __dsub_rn(c[1], q[1]), dz = __dsub_rn(c[2], q[2]);
        const double dm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        if (best > dm || (bi >= 0 && best == dm && pi < bidx)) { best = dm; bi = i; bidx = pi; }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {                  // smaller distance, the lower original index among equals
        const double ob = __shfl_xor_sync(KSS_FULL, best, off);
        const int oi = __shfl_xor_sync(KSS_FULL, bi, off), ox = __shfl_xor_sync(KSS_FULL, bidx, off);
        if (oi >= 0 && (bi < 0 || ob < best || (ob == best && ox < bidx))) { best = ob; bi = oi; bidx = ox; }
    }
    if (lane == 0) {
        center_pos[(size_t)p * bcap + b] = bi;
        const double rate = __ddiv_rn((double)g.point_num, (double)g.n);
        const double sb = __dmul_rn((double)m, rate);
        const int t = (int)sb;
        quota[(size_t)p * bcap + b] = (__dsub_rn(sb, (double)t) > 0.2) ? t + 1 : t;
    }
}

// ---------------------------------------------------------------- 3. farthest point sampling, one colour per launch
__device__ __forceinline__ int aivs_colour(int i, int j, int k) {
    const bool io = i & 1, jo = j & 1, ko = k & 1;
    if (io && jo && ko) return 0;
    if (!io && jo && ko) return 1;
    if (!io && !jo && ko) return 2;
    if (io && !jo && ko) return 3;
    if (io && jo && !ko) return 4;
    if (!io && jo && !ko) return 5;
    if (!io && !jo && !ko) return 6;
    return 7;
}

// one box of AIVS_Voroni_OpenMP_KNN, run by one WARP: lane l owns the members t = l (mod 32) of the box (their running
// minimum distances md[t] live in global memory); seeds are found 32 candidates at a time and applied one after the other
// (a minimum does not care about the order); the arg-max keeps the reference's first strict maximum
__device__ void aivs_fps_box(int p, int b, const AivsGrid& g, const double* __restrict__ pts, int cap, int bcap,
                             const int* __restrict__ box_start, const int* __restrict__ members,
                             const int* __restrict__ center_pos, const int* __restrict__ quota,
                             unsigned char* __restrict__ selected /* labelG == 0 */, double* __restrict__ mind,
                             int* __restrict__ sel, int* __restrict__ sel_cnt) {
    const int lane = threadIdx.x & 31;
    const int* st = box_start + (size_t)p * (bcap + 1);
    const int s = st[b], m = st[b + 1] - s;
    const int simNum = quota[(size_t)p * bcap + b];
    if (m == 0 || simNum == 0) return;
    const int* mem = members + (size_t)p * cap;
    const double* P = pts + (size_t)p * cap * 3;
    unsigned char* lab = selected + (size_t)p * cap;
    double* md = mind + (size_t)p * cap + s;
    int* out = sel + (size_t)p * cap + s;
    double pc[3];
    aivs_center(g, b, pc);
    const double radius = __ddiv_rn(__dmul_rn(g.unit, 3.0), 4.0);
    const double cl[3] = {__dsub_rn(pc[0], radius), __dsub_rn(pc[1], radius), __dsub_rn(pc[2], radius)};
    const double ch[3] = {__dadd_rn(pc[0], radius), __dadd_rn(pc[1], radius), __dadd_rn(pc[2], radius)};
    // neighbour boxes through the reference's decode (no x wrap, B10)
    const int z_num = b / (g.nx * g.ny) + 1;
    const int leveZ = b % (g.nx * g.ny);
    const int y_num = leveZ / g.nx + 1;
    const int x_num = leveZ % g.nx;
    int xs[3], ys[3], zs[3], nxs = 0, nys = 0, nzs = 0;
    if (x_num > 1) xs[nxs++] = x_num - 1;
    xs[nxs++] = x_num;
    if (x_num < g.nx) xs[nxs++] = x_num + 1;
    if (y_num > 1) ys[nys++] = y_num - 1;
    ys[nys++] = y_num;
    if (y_num < g.ny) ys[nys++] = y_num + 1;
    if (z_num > 1) zs[nzs++] = z_num - 1;
    zs[nzs++] = z_num;
    if (z_num < g.nz) zs[nzs++] = z_num + 1;
    auto relax = [&](float sx, float sy, float sz, bool assign) {     // md[t] = min(md[t], |member t - s|) over the lane's members
        for (int t = lane; t < m; t += 32) {
            const double* w = P + 3 * (size_t)mem[s + t];
            const double d = (double)__fsqrt_rn(d2_rn((float)w[0], (float)w[1], (float)w[2], sx, sy, sz));
            if (assign || d < md[t]) md[t] = d;
        }
    };
    // initial distances: to the nearest already-selected point of the neighbour boxes inside the seed cube,
    // or (if there is none) to the box's centre-nearest member, which then becomes the first sample.  The selected
    // points of a box are its sample list so far (sel / sel_cnt): lane k looks at neighbour box k.
    for (int t = lane; t < m; t += 32) md[t] = INFINITY;
    bool any_seed = false;
    {
        int nb = -1, ncnt = 0;
        if (lane < 27) {
            const int a = lane % 3, c = (lane / 3) % 3, e = lane / 9;
            if (a < nxs && c < nys && e < nzs && !(xs[a] == x_num && ys[c] == y_num && zs[e] == z_num)) {
                const int q = xs[a] + (ys[c] - 1) * g.nx + (zs[e] - 1) * g.nx * g.ny;
                if (q < g.nbox && q >= 0) { nb = q; ncnt = sel_cnt[(size_t)p * bcap + q]; }
            }
        }
        const int rounds = __reduce_max_sync(KSS_FULL, ncnt);
        for (int l = 0; l < rounds; ++l) {
            bool seed = false;
            float sx = 0.f, sy = 0.f, sz = 0.f;
            if (l < ncnt) {
                const int pt = sel[(size_t)p * cap + st[nb] + l];
                const double* q = P + 3 * (size_t)pt;
                seed = q[0] <= ch[0] && q[0] >= cl[0] && q[1] <= ch[1] && q[1] >= cl[1] && q[2] <= ch[2] && q[2] >= cl[2];
                sx = (float)q[0]; sy = (float)q[1]; sz = (float)q[2];
            }
            unsigned bal = __ballot_sync(KSS_FULL, seed);
            any_seed |= bal != 0u;
            while (bal) {
                const int src = __ffs(bal) - 1; bal &= bal - 1u;
                relax(__shfl_sync(KSS_FULL, sx, src), __shfl_sync(KSS_FULL, sy, src), __shfl_sync(KSS_FULL, sz, src), false);
            }
        }
    }
    for (int t = lane; t < m; t += 32) if (md[t] == INFINITY) md[t] = 9999.0;
    int sampled = 0;
    if (!any_seed) {
        const int ci = center_pos[(size_t)p * bcap + b];
        if (ci >= 0 && ci < m) {
            const double* q = P + 3 * (size_t)mem[s + ci];
            relax((float)q[0], (float)q[1], (float)q[2], true);
            __syncwarp();
            if (lane == 0) { md[ci] = 0.0; out[sampled] = mem[s + ci]; lab[mem[s + ci]] = 1; }
            ++sampled;
            __syncwarp();
        }
    }
    while (sampled < simNum) {
        int pick = -1, pidx = 0x7fffffff; double mx = 0.0;
        for (int t = lane; t < m; t += 32) {                   // the first strict maximum = largest value, lowest original index
            const double v = md[t];
            if (v > mx || (pick >= 0 && v == mx && mem[s + t] < pidx)) { pick = t; mx = v; pidx = mem[s + t]; }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const double om = __shfl_xor_sync(KSS_FULL, mx, off);
            const int op = __shfl_xor_sync(KSS_FULL, pick, off), ox = __shfl_xor_sync(KSS_FULL, pidx, off);
            if (op >= 0 && (pick < 0 || om > mx || (om == mx && ox < pidx))) { mx = om; pick = op; pidx = ox; }
        }
        if (pick == -1) break;
        const int pp = mem[s + pick];
        if (lane == 0) { md[pick] = 0.0; lab[pp] = 1; out[sampled] = pp; }
        ++sampled;
        __syncwarp();
        const double* q = P + 3 * (size_t)pp;
        relax((float)q[0], (float)q[1], (float)q[2], false);
        __syncwarp();
    }
    if (lane == 0) sel_cnt[(size_t)p * bcap + b] = sampled;
}

// Same-colour boxes never see each other's points -- except the boxes whose index is a multiple of nx * ny: the
// reference's centre decode misplaces their centre (B10), so their seed cube can reach same-colour boxes, including
// each other.  In the reference's loop order they are the last boxes of their colour, in ascending index.  So a colour
// is two phases: 2c = all other boxes, one warp each; 2c + 1 = the misplaced ones, one after the other (one warp).
__global__ void __launch_bounds__(256)
aivs_fps_kernel(int phase, const double* __restrict__ pts, int cap, const AivsGrid* __restrict__ grids, int bcap,
                const int* __restrict__ box_start, const int* __restrict__ members, const int* __restrict__ center_pos,
                const int* __restrict__ quota, unsigned char* __restrict__ selected, double* __restrict__ mind,
                int* __restrict__ sel, int* __restrict__ sel_cnt) {
    const int p = blockIdx.y;
    const AivsGrid g = grids[p];
    const int colour = phase >> 1;
    if (phase & 1) {
        if (blockIdx.x != 0 || threadIdx.x >= 32) return;
        for (int k = 1; k <= g.nz; ++k) {
            const int b = g.nx * g.ny * k;
            if (b >= g.nbox || aivs_colour(g.nx, g.ny, k) != colour) continue;
            aivs_fps_box(p, b, g, pts, cap, bcap, box_start, members, center_pos, quota, selected, mind, sel, sel_cnt);
            __threadfence();                                   // (the next misplaced box may read this one's labels)
            __syncwarp();
        }
        return;
    }
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b < 1 || b >= g.nbox || (b % (g.nx * g.ny)) == 0) return;
    // true (i,j,k) of the box, as the colouring loop enumerates them (Method_AIVS_SimPro.hpp:598-602)
    const int b0 = b - 1;
    if (aivs_colour(b0 % g.nx + 1, (b0 / g.nx) % g.ny + 1, b0 / (g.nx * g.ny) + 1) != colour) return;
    aivs_fps_box(p, b, g, pts, cap, bcap, box_start, members, center_pos, quota, selected, mind, sel, sel_cnt);
}

// The even phases with the box staged in shared memory: the members of a warp's box (coordinates narrowed as the reference
// narrows them, original indices, running minimum distances -- all of them floats: a distance is a widened float, 9999 and
// 0 are floats too) are read once; every relax pass and every arg-max then runs over shared memory.  (aivs_fps_box walks
// its members through global gathers and global minimum distances 10-30 times per box: a 1M-point cloud spent 0.13-0.33 ms
// per colour there, one dependent chain per warp.)  The arg-max is the reference's first strict maximum = largest value,
// lowest original index; the order in which seeds are applied does not matter to a minimum.  Boxes with more members than
// fit take aivs_fps_box.
constexpr int FPS_WCAP = 512;                // members staged per box by a warp (20 bytes each)
constexpr int FPS_MCAP = 4096;               // ... by a CTA (aivs_fps_big_kernel)
constexpr int FPS_BIGCAP = 1024;             // boxes per cloud and colour that can be handed to aivs_fps_big_kernel
constexpr size_t FPS_SMEM = (size_t)8 * (FPS_WCAP * 20 + 27 * 3 * 4);
__global__ void __launch_bounds__(256)
aivs_fps_smem_kernel(int colour, const double* __restrict__ pts, int cap, const AivsGrid* __restrict__ grids, int bcap,
                     const int* __restrict__ box_start, const int* __restrict__ members, const int* __restrict__ center_pos,
                     const int* __restrict__ quota, unsigned char* __restrict__ selected, double* __restrict__ mind,
                     int* __restrict__ sel, int* __restrict__ sel_cnt, int* __restrict__ biglist, int* __restrict__ bigcnt) {
    extern __shared__ unsigned char fps_raw[];
    const int p = blockIdx.y;
    const AivsGrid g = grids[p];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * 8 + warp;
    if (b < 1 || b >= g.nbox || (b % (g.nx * g.ny)) == 0) return;
    const int b0 = b - 1;
    if (aivs_colour(b0 % g.nx + 1, (b0 / g.nx) % g.ny + 1, b0 / (g.nx * g.ny) + 1) != colour) return;
    const int* st = box_start + (size_t)p * (bcap + 1);
    const int s = st[b], m = st[b + 1] - s;
    const int simNum = quota[(size_t)p * bcap + b];
    if (m == 0 || simNum == 0) return;
    if (m > FPS_WCAP) {
        // a big box would be the long pole of the launch on one warp: aivs_fps_big_kernel gives it a CTA (same colour, so
        // nobody here reads its samples)
        int slot = FPS_BIGCAP;                                   // (no list: small clouds, where eight more launches cost more than the pole)
        if (biglist && lane == 0) slot = atomicAdd(&bigcnt[p * 8 + colour], 1);
        slot = __shfl_sync(KSS_FULL, slot, 0);
        if (slot < FPS_BIGCAP) { if (lane == 0) biglist[((size_t)p * 8 + colour) * FPS_BIGCAP + slot] = b; }
        else aivs_fps_box(p, b, g, pts, cap, bcap, box_start, members, center_pos, quota, selected, mind, sel, sel_cnt);
        return;
    }
    float* X = reinterpret_cast<float*>(fps_raw + (size_t)warp * (FPS_WCAP * 20 + 27 * 3 * 4));
    float* Y = X + FPS_WCAP;
    float* Z = Y + FPS_WCAP;
    float* MD = Z + FPS_WCAP;
    int* IDX = reinterpret_cast<int*>(MD + FPS_WCAP);
    float* sd = reinterpret_cast<float*>(IDX + FPS_WCAP);        // [27][3] seeds of one round
    const int* mem = members + (size_t)p * cap;
    const double* P = pts + (size_t)p * cap * 3;
    unsigned char* lab = selected + (size_t)p * cap;
    int* out = sel + (size_t)p * cap + s;
    for (int t = lane; t < m; t += 32) {
        const int i = mem[s + t];
        const double* w = P + 3 * (size_t)i;
        X[t] = (float)w[0]; Y[t] = (float)w[1]; Z[t] = (float)w[2];
        IDX[t] = i;
        MD[t] = INFINITY;
    }
    // seeds: the already-selected points of the neighbour boxes inside the seed cube (lane k looks at neighbour box k)
    double pc[3];
    aivs_center(g, b, pc);
    const double radius = __ddiv_rn(__dmul_rn(g.unit, 3.0), 4.0);
    const double cl[3] = {__dsub_rn(pc[0], radius), __dsub_rn(pc[1], radius), __dsub_rn(pc[2], radius)};
    const double ch[3] = {__dadd_rn(pc[0], radius), __dadd_rn(pc[1], radius), __dadd_rn(pc[2], radius)};
    const int z_num = b / (g.nx * g.ny) + 1;
    const int leveZ = b % (g.nx * g.ny);
    const int y_num = leveZ / g.nx + 1;
    const int x_num = leveZ % g.nx;
    int xs[3], ys[3], zs[3], nxs = 0, nys = 0, nzs = 0;
    if (x_num > 1) xs[nxs++] = x_num - 1;
    xs[nxs++] = x_num;
    if (x_num < g.nx) xs[nxs++] = x_num + 1;
    if (y_num > 1) ys[nys++] = y_num - 1;
    ys[nys++] = y_num;
    if (y_num < g.ny) ys[nys++] = y_num + 1;
    if (z_num > 1) zs[nzs++] = z_num - 1;
    zs[nzs++] = z_num;
    if (z_num < g.nz) zs[nzs++] = z_num + 1;
    int nb = -1, ncnt = 0;
    if (lane < 27) {
        const int a = lane % 3, c = (lane / 3) % 3, e = lane / 9;
        if (a < nxs && c < nys && e < nzs && !(xs[a] == x_num && ys[c] == y_num && zs[e] == z_num)) {
            const int q = xs[a] + (ys[c] - 1) * g.nx + (zs[e] - 1) * g.nx * g.ny;
            if (q < g.nbox && q >= 0) { nb = q; ncnt = sel_cnt[(size_t)p * bcap + q]; }
        }
    }
    const int rounds = __reduce_max_sync(KSS_FULL, ncnt);
    bool any_seed = false;
    __syncwarp();
    for (int l = 0; l < rounds; ++l) {
        bool seed = false;
        float sx = 0.f, sy = 0.f, sz = 0.f;
        if (l < ncnt) {
            const int pt = sel[(size_t)p * cap + st[nb] + l];
            const double* q = P + 3 * (size_t)pt;
            seed = q[0] <= ch[0] && q[0] >= cl[0] && q[1] <= ch[1] && q[1] >= cl[1] && q[2] <= ch[2] && q[2] >= cl[2];
            sx = (float)q[0]; sy = (float)q[1]; sz = (float)q[2];
        }
        const unsigned bal = __ballot_sync(KSS_FULL, seed);
        if (!bal) continue;
        any_seed = true;
        if (seed) { const int k = __popc(bal & ((1u << lane) - 1u)); sd[3 * k] = sx; sd[3 * k + 1] = sy; sd[3 * k + 2] = sz; }
        __syncwarp();
        const int ns = __popc(bal);
        for (int t = lane; t < m; t += 32) {
            float md = MD[t];
            const float x = X[t], y = Y[t], z = Z[t];
            for (int k = 0; k < ns; ++k) {
                const float d = __fsqrt_rn(d2_rn(x, y, z, sd[3 * k], sd[3 * k + 1], sd[3 * k + 2]));
                if (d < md) md = d;
            }
            MD[t] = md;
        }
        __syncwarp();
    }
    for (int t = lane; t < m; t += 32) if (MD[t] == INFINITY) MD[t] = 9999.0f;
    __syncwarp();
    int sampled = 0;
    if (!any_seed) {
        const int ci = center_pos[(size_t)p * bcap + b];
        if (ci >= 0 && ci < m) {
            const float sx = X[ci], sy = Y[ci], sz = Z[ci];
            for (int t = lane; t < m; t += 32) MD[t] = __fsqrt_rn(d2_rn(X[t], Y[t], Z[t], sx, sy, sz));
            __syncwarp();
            if (lane == 0) { MD[ci] = 0.0f; out[0] = IDX[ci]; lab[IDX[ci]] = 1; }
            sampled = 1;
            __syncwarp();
        }
    }
    while (sampled < simNum) {
        int pick = -1, pidx = 0x7fffffff; float mx = 0.0f;
        for (int t = lane; t < m; t += 32) {                   // the first strict maximum = largest value, lowest original index
            const float v = MD[t];
            if (v > mx || (pick >= 0 && v == mx && IDX[t] < pidx)) { pick = t; mx = v; pidx = IDX[t]; }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const float om = __shfl_xor_sync(KSS_FULL, mx, off);
            const int op = __shfl_xor_sync(KSS_FULL, pick, off), ox = __shfl_xor_sync(KSS_FULL, pidx, off);
            if (op >= 0 && (pick < 0 || om > mx || (om == mx && ox < pidx))) { mx = om; pick = op; pidx = ox; }
        }
        if (pick == -1) break;
        if (lane == 0) { MD[pick] = 0.0f; lab[pidx] = 1; out[sampled] = pidx; }
        ++sampled;
        __syncwarp();
        const float sx = X[pick], sy = Y[pick], sz = Z[pick];
        for (int t = lane; t < m; t += 32) {
            const float d = __fsqrt_rn(d2_rn(X[t], Y[t], Z[t], sx, sy, sz));
            if (d < MD[t]) MD[t] = d;
        }
        __syncwarp();
    }
    if (lane == 0) sel_cnt[(size_t)p * bcap + b] = sampled;
}

// The boxes aivs_fps_smem_kernel set aside (more than FPS_WCAP members), one CTA each: the same algorithm with 256 threads
// over the staged members, the arg-max reduced over the CTA with the same rule.
__global__ void __launch_bounds__(256)
aivs_fps_big_kernel(int colour, const double* __restrict__ pts, int cap, const AivsGrid* __restrict__ grids, int bcap,
                    const int* __restrict__ box_start, const int* __restrict__ members, const int* __restrict__ center_pos,
                    const int* __restrict__ quota, unsigned char* __restrict__ selected, double* __restrict__ mind,
                    int* __restrict__ sel, int* __restrict__ sel_cnt, const int* __restrict__ biglist, const int* __restrict__ bigcnt) {
    extern __shared__ unsigned char fps_raw[];
    __shared__ float sd[27][3];
    __shared__ int s_nseed, s_any, s_rounds, s_pick;
    __shared__ float r_v[8];
    __shared__ int r_t[8], r_i[8];
    const int p = blockIdx.y;
    const AivsGrid g = grids[p];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nbig = min(bigcnt[p * 8 + colour], FPS_BIGCAP);
    float* X = reinterpret_cast<float*>(fps_raw);
    float* Y = X + FPS_MCAP;
    float* Z = Y + FPS_MCAP;
    float* MD = Z + FPS_MCAP;
    int* IDX = reinterpret_cast<int*>(MD + FPS_MCAP);
    const int* st = box_start + (size_t)p * (bcap + 1);
    const int* mem = members + (size_t)p * cap;
    const double* P = pts + (size_t)p * cap * 3;
    unsigned char* lab = selected + (size_t)p * cap;
    for (int e = blockIdx.x; e < nbig; e += gridDim.x) {
        __syncthreads();                                         // (the previous box of this CTA is done with the shared arrays)
        const int b = biglist[((size_t)p * 8 + colour) * FPS_BIGCAP + e];
        const int s = st[b], m = st[b + 1] - s;
        const int simNum = quota[(size_t)p * bcap + b];
        if (m > FPS_MCAP) {
            if (tid < 32) aivs_fps_box(p, b, g, pts, cap, bcap, box_start, members, center_pos, quota, selected, mind, sel, sel_cnt);
            continue;
        }
        int* out = sel + (size_t)p * cap + s;
        for (int t = tid; t < m; t += 256) {
            const int i = mem[s + t];
            const double* w = P + 3 * (size_t)i;
            X[t] = (float)w[0]; Y[t] = (float)w[1]; Z[t] = (float)w[2];
            IDX[t] = i;
            MD[t] = INFINITY;
        }
        // seeds: warp 0 finds them, 27 neighbour boxes at a time
        int nb = -1, ncnt = 0;
        double cl[3] = {0.0, 0.0, 0.0}, ch[3] = {0.0, 0.0, 0.0};
        if (warp == 0) {
            double pc[3];
            aivs_center(g, b, pc);
            const double radius = __ddiv_rn(__dmul_rn(g.unit, 3.0), 4.0);
            for (int a = 0; a < 3; ++a) { cl[a] = __dsub_rn(pc[a], radius); ch[a] = __dadd_rn(pc[a], radius); }
            const int z_num = b / (g.nx * g.ny) + 1;
            const int leveZ = b % (g.nx * g.ny);
            const int y_num = leveZ / g.nx + 1;
            const int x_num = leveZ % g.nx;
            int xs[3], ys[3], zs[3], nxs = 0, nys = 0, nzs = 0;
            if (x_num > 1) xs[nxs++] = x_num - 1;
            xs[nxs++] = x_num;
            if (x_num < g.nx) xs[nxs++] = x_num + 1;
            if (y_num > 1) ys[nys++] = y_num - 1;
            ys[nys++] = y_num;
            if (y_num < g.ny) ys[nys++] = y_num + 1;
            if (z_num > 1) zs[nzs++] = z_num - 1;
            zs[nzs++] = z_num;
            if (z_num < g.nz) zs[nzs++] = z_num + 1;
            if (lane < 27) {
                const int a = lane % 3, c = (lane / 3) % 3, e2 = lane / 9;
                if (a < nxs && c < nys && e2 < nzs && !(xs[a] == x_num && ys[c] == y_num && zs[e2] == z_num)) {
                    const int q = xs[a] + (ys[c] - 1) * g.nx + (zs[e2] - 1) * g.nx * g.ny;
                    if (q < g.nbox && q >= 0) { nb = q; ncnt = sel_cnt[(size_t)p * bcap + q]; }
                }
            }
            const int rounds = __reduce_max_sync(KSS_FULL, ncnt);
            if (lane == 0) { s_rounds = rounds; s_any = 0; }
        }
        __syncthreads();
        const int rounds = s_rounds;
        for (int l = 0; l < rounds; ++l) {
            if (warp == 0) {
                bool seed = false;
                float sx = 0.f, sy = 0.f, sz = 0.f;
                if (l < ncnt) {
                    const int pt = sel[(size_t)p * cap + st[nb] + l];
                    const double* q = P + 3 * (size_t)pt;
                    seed = q[0] <= ch[0] && q[0] >= cl[0] && q[1] <= ch[1] && q[1] >= cl[1] && q[2] <= ch[2] && q[2] >= cl[2];
                    sx = (float)q[0]; sy = (float)q[1]; sz = (float)q[2];
                }
                const unsigned bal = __ballot_sync(KSS_FULL, seed);
                if (seed) { const int k = __popc(bal & ((1u << lane) - 1u)); sd[k][0] = sx; sd[k][1] = sy; sd[k][2] = sz; }
                if (lane == 0) { s_nseed = __popc(bal); if (bal) s_any = 1; }
            }
            __syncthreads();
            const int ns = s_nseed;
            if (ns)
                for (int t = tid; t < m; t += 256) {
                    float md = MD[t];
                    const float x = X[t], y = Y[t], z = Z[t];
                    for (int k = 0; k < ns; ++k) {
                        const float d = __fsqrt_rn(d2_rn(x, y, z, sd[k][0], sd[k][1], sd[k][2]));
                        if (d < md) md = d;
                    }
                    MD[t] = md;
                }
            __syncthreads();
        }
        for (int t = tid; t < m; t += 256) if (MD[t] == INFINITY) MD[t] = 9999.0f;
        __syncthreads();
        int sampled = 0;
        if (!s_any) {
            const int ci = center_pos[(size_t)p * bcap + b];
            if (ci >= 0 && ci < m) {
                const float sx = X[ci], sy = Y[ci], sz = Z[ci];
                for (int t = tid; t < m; t += 256) MD[t] = __fsqrt_rn(d2_rn(X[t], Y[t], Z[t], sx, sy, sz));
                __syncthreads();
                if (tid == 0) { MD[ci] = 0.0f; out[0] = IDX[ci]; lab[IDX[ci]] = 1; }
                sampled = 1;
                __syncthreads();
            }
        }
        while (sampled < simNum) {
            int pick = -1, pidx = 0x7fffffff; float mx = 0.0f;
            for (int t = tid; t < m; t += 256) {
                const float v = MD[t];
                if (v > mx || (pick >= 0 && v == mx && IDX[t] < pidx)) { pick = t; mx = v; pidx = IDX[t]; }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                const float om = __shfl_xor_sync(KSS_FULL, mx, off);
                const int op = __shfl_xor_sync(KSS_FULL, pick, off), ox = __shfl_xor_sync(KSS_FULL, pidx, off);
                if (op >= 0 && (pick < 0 || om > mx || (om == mx && ox < pidx))) { mx = om; pick = op; pidx = ox; }
            }
            if (lane == 0) { r_v[warp] = mx; r_t[warp] = pick; r_i[warp] = pidx; }
            __syncthreads();
            if (tid == 0) {
                for (int w = 1; w < 8; ++w) {
                    const float om = r_v[w]; const int op = r_t[w], ox = r_i[w];
                    if (op >= 0 && (pick < 0 || om > mx || (om == mx && ox < pidx))) { mx = om; pick = op; pidx = ox; }
                }
                s_pick = pick;
                if (pick >= 0) { MD[pick] = 0.0f; lab[pidx] = 1; out[sampled] = pidx; }
            }
            __syncthreads();
            pick = s_pick;
            if (pick == -1) break;
            ++sampled;
            const float sx = X[pick], sy = Y[pick], sz = Z[pick];
            for (int t = tid; t < m; t += 256) {
                const float d = __fsqrt_rn(d2_rn(X[t], Y[t], Z[t], sx, sy, sz));
                if (d < MD[t]) MD[t] = d;
            }
            __syncthreads();
        }
        if (tid == 0) sel_cnt[(size_t)p * bcap + b] = sampled;
    }
}

// ---------------------------------------------------------------- 4. samples in box order, K = 3 lists, greedy trim
constexpr int AIVS_MAX_SAMPLES = 16384;      // trim works on at most this many samples per cloud (14-bit ids in the sort key)

__global__ void __launch_bounds__(256)
aivs_gather_kernel(int cap, const AivsGrid* __restrict__ grids, int bcap, const int* __restrict__ box_start,
                   const int* __restrict__ sel, const int* __restrict__ sel_cnt, const int* __restrict__ sel_start,
                   int* __restrict__ sample) {
    const int p = blockIdx.y;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= grids[p].nbox) return;
    const int c = sel_cnt[(size_t)p * bcap + b];
    if (c == 0) return;
    const int src = box_start[(size_t)p * (bcap + 1) + b], dst = sel_start[(size_t)p * (bcap + 1) + b];
    for (int r = 0; r < c; ++r) sample[(size_t)p * cap + dst + r] = sel[(size_t)p * cap + src + r];
}

// K = 3 nearest samples of every sample, ascending (d2, position): entry [0] is the sample itself (or an exact
// duplicate in front of it), [1] the partner of the trim step, [2] decides which of the two is deleted
__global__ void __launch_bounds__(256)
aivs_k3_kernel(const double* __restrict__ pts, int cap, const AivsGrid* __restrict__ grids, int bcap,
               const int* __restrict__ sel_start, const int* __restrict__ sample,
               unsigned long long* __restrict__ key1, float* __restrict__ dis2) {
    __shared__ float tx[256], ty[256], tz[256];
    const int p = blockIdx.y;
    const AivsGrid g = grids[p];
    if (g.nbox == 0) return;
    const int S = sel_start[(size_t)p * (bcap + 1) + g.nbox];
    if (S <= g.point_num || S < 3 || S > AIVS_MAX_SAMPLES || (int)(blockIdx.x * blockDim.x) >= S) return;
    const double* P = pts + (size_t)p * cap * 3;
    const int* smp = sample + (size_t)p * cap;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float x = 0.f, y = 0.f, z = 0.f;
    if (i < S) { const double* q = P + 3 * (size_t)smp[i]; x = (float)q[0]; y = (float)q[1]; z = (float)q[2]; }
    unsigned long long k0 = ~0ull, k1 = ~0ull, k2 = ~0ull;
    for (int base = 0; base < S; base += 256) {
        __syncthreads();
        if (base + (int)threadIdx.x < S) {
            const double* q = P + 3 * (size_t)smp[base + threadIdx.x];
            tx[threadIdx.x] = (float)q[0]; ty[threadIdx.x] = (float)q[1]; tz[threadIdx.x] = (float)q[2];
        }
        __syncthreads();
        const int m = min(256, S - base);
        for (int j = 0; j < m; ++j) {
            const unsigned long long key = ((unsigned long long)__float_as_uint(d2_rn(x, y, z, tx[j], ty[j], tz[j])) << 32) | (unsigned)(base + j);
            if (key < k2) {
                if (key < k0) { k2 = k1; k1 = k0; k0 = key; }
                else if (key < k1) { k2 = k1; k1 = key; }
                else k2 = key;
            }
        }
    }
    if (i < S) {
        const float d1 = __fsqrt_rn(__uint_as_float((unsigned)(k1 >> 32)));
        key1[(size_t)p * cap + i] = ((unsigned long long)__float_as_uint(d1) << 32) | ((unsigned long long)i << 16) | (k1 & 0xffffull);
        dis2[(size_t)p * cap + i] = __fsqrt_rn(__uint_as_float((unsigned)(k2 >> 32)));
    }
}

// The reference's loop "take the first strict minimum of dis[1] among samples whose pair is still alive, delete one of
// the two, repeat" never updates the lists, and a pair that is dead stays dead: so one walk over the samples sorted by
// (dis[1], position) visits exactly the same pairs in the same order.
__global__ void __launch_bounds__(512)
aivs_cut_kernel(const double* __restrict__ pts, int cap, AivsGrid* __restrict__ grids, int bcap,
                const int* __restrict__ sel_start, const int* __restrict__ sample,
                const unsigned long long* __restrict__ key1, const float* __restrict__ dis2, int np2,
                double* __restrict__ out, int out_cap, int* __restrict__ out_cnt, int* __restrict__ out_idx,
                int* __restrict__ bad) {
    extern __shared__ unsigned long long smem_keys[];                  // [np2] keys, [np2] float dis2, [np2] keep
    const int p = blockIdx.x;
    const AivsGrid g = grids[p];
    const int S = g.nbox ? sel_start[(size_t)p * (bcap + 1) + g.nbox] : 0;
    const int* smp = sample + (size_t)p * cap;
    const double* P = pts + (size_t)p * cap * 3;
    float* d2s = reinterpret_cast<float*>(smem_keys + np2);
    unsigned char* keep = reinterpret_cast<unsigned char*>(d2s + np2);
    int dT = S - g.point_num;
    if (dT > 0 && S > AIVS_MAX_SAMPLES) { if (threadIdx.x == 0) atomicExch(bad, 2); dT = 0; }
    const bool trim = dT > 0 && S >= 3;
    if (trim) {
        int np = 1; while (np < S) np <<= 1;
        for (int i = threadIdx.x; i < np; i += blockDim.x) {
            smem_keys[i] = i < S ? key1[(size_t)p * cap + i] : ~0ull;
            if (i < S) { d2s[i] = dis2[(size_t)p * cap + i]; keep[i] = 1; }
        }
        for (int k = 2; k <= np; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                __syncthreads();
                for (int i = threadIdx.x; i < np; i += blockDim.x) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const unsigned long long a = smem_keys[i], b = smem_keys[ixj];
                        if ((a > b) == ((i & k) == 0)) { smem_keys[i] = b; smem_keys[ixj] = a; }
                    }
                }
            }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int pos = 0; pos < S && dT > 0; ++pos) {
                const unsigned long long key = smem_keys[pos];
                if (!((double)__uint_as_float((unsigned)(key >> 32)) < 9999.0)) break;
                const int b1 = (int)((key >> 16) & 0xffffu), b2 = (int)(key & 0xffffu);
                if (!keep[b1] || !keep[b2]) continue;
                keep[((double)d2s[b1] > (double)d2s[b2]) ? b2 : b1] = 0;
                --dT;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {                                            // ordered output of the kept samples
        int count = 0;
        for (int base = 0; base < S; base += 32) {
            const int i = base + threadIdx.x;
            const bool f = i < S && (!trim || keep[i]);
            const unsigned bal = __ballot_sync(KSS_FULL, f);
            if (f) {
                const int pos = count + __popc(bal & ((1u << threadIdx.x) - 1u));
                if (pos < out_cap) {
                    const double* q = P + 3 * (size_t)smp[i];
                    double* o = out + ((size_t)p * out_cap + pos) * 3;
                    o[0] = q[0]; o[1] = q[1]; o[2] = q[2];
                    if (out_idx) out_idx[(size_t)p * out_cap + pos] = smp[i];
                }
            }
            count += __popc(bal);
        }
        if (threadIdx.x == 0) {
            out_cnt[p] = min(count, out_cap);
            if (count > out_cap) atomicExch(bad, 3);
            grids[p].n_samples = S;
        }
    }
}

// ================================================================ clouds of at most 2048 points: one CTA per cloud
// The whole simplification in one launch, everything in shared memory: float copies of the points, the box
// membership as a counting sort by box followed by per-box insertion sorts (ascending index), farthest point sampling with GL lanes per box (same
// colour boxes in parallel, a barrier between colours), the K = 3 lists and the sorted greedy trim.  Same arithmetic
// and the same selection rules as the general kernels above (the min-distance array is kept in float: every value
// is a float square root or 9999, so comparisons are unchanged).
constexpr int AS_THREADS = 256;
constexpr int AS_SB = 1336;                   // >= 11^3 + 2 boxes (10 per axis, 11 when the division rounds up)
constexpr int AS_MAX = 2048;
#ifndef AS_GL
#define AS_GL 16                             // lanes per box in the sampling phases
#endif

__device__ __forceinline__ int as_block_excl_scan(int* a, int L, int* wsum) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int per = (L + AS_THREADS - 1) / AS_THREADS;
    const int lo = min(L, tid * per), hi = min(L, lo + per);
    int sum = 0;
    for (int i = lo; i < hi; ++i) sum += a[i];
    int x = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(KSS_FULL, x, o); if (lane >= o) x += y; }
    if (lane == 31) wsum[warp] = x;
    __syncthreads();
    int base = x - sum, total = 0;
#pragma unroll
    for (int w = 0; w < AS_THREADS / 32; ++w) { const int v = wsum[w]; if (w < warp) base += v; total += v; }
    for (int i = lo; i < hi; ++i) { const int v = a[i]; a[i] = base; base += v; }
    __syncthreads();
    return total;
}

template <typename K>
__device__ __forceinline__ void as_bitonic(K* keys, int np) {
    for (int k = 2; k <= np; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int i = threadIdx.x; i < np; i += AS_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const K a = keys[i], b = keys[ixj];
                    if ((a > b) == ((i & k) == 0)) { keys[i] = b; keys[ixj] = a; }
                }
            }
        }
    __syncthreads();
}

struct AsLayout {               // byte offsets into dynamic shared memory
    int xf, yf, zf, sample;
    int mkey, start, centre, quota, clist, md, sel, lab, selc;       // while the boxes are sampled
    int key64, d2s, keep, sbox;                                      // for the trim (aliases everything above but selc)
    int total;
};
__host__ __device__ inline AsLayout as_layout(int cap) {
    AsLayout L;
    const int capA = (cap + 7) & ~7;
    int np = 1; while (np < cap) np <<= 1;
    int o = 0;
    L.xf = o; o += 4 * capA; L.yf = o; o += 4 * capA; L.zf = o; o += 4 * capA;
    L.sample = o; o += 2 * capA;
    o = (o + 15) & ~15;
    const int a_base = o;
    L.mkey = o; o += 4 * np; L.start = o; o += 4 * AS_SB; L.centre = o; o += 2 * AS_SB; L.quota = o; o += 2 * AS_SB;
    L.clist = o; o += 2 * AS_SB;
    L.md = o; o += 4 * capA; L.sel = o; o += 2 * capA; L.lab = o; o += capA;
    L.selc = o; o += 4 * AS_SB;
    int end = o;
    o = a_base;
    L.key64 = o; o += 8 * np; L.d2s = o; o += 4 * capA; L.keep = o; o += capA;
    // sample -> box table: written while `sel` is read and `selc` is live, read while key64/d2s/keep are written:
    // inside the dead `md` block when it lies behind the trim arrays, else appended
    L.sbox = (o > L.md ? o : L.md);
    L.sbox = (L.sbox + 7) & ~7;
    if (L.sbox + 2 * capA > L.sel) { L.sbox = (end + 7) & ~7; end = L.sbox + 2 * capA; }
    if (o > L.selc) { /* trim arrays would reach selc: move selc behind them */ L.selc = (end + 7) & ~7; end = L.selc + 4 * AS_SB; }
    L.total = end + 16;
    return L;
}

template <int GL>
__global__ void __launch_bounds__(AS_THREADS, 3)
aivs_small_kernel(const double* __restrict__ pts, const int* __restrict__ cnt, int cap, const int* __restrict__ point_num,
                  int point_num_all, double* __restrict__ out, int out_cap, int* __restrict__ out_cnt,
                  int* __restrict__ out_idx, int* __restrict__ bad) {
    extern __shared__ __align__(16) unsigned char as_smem[];
    __shared__ AivsGrid g_s;
    __shared__ double red[6][AS_THREADS / 32];
    __shared__ int wsum[AS_THREADS / 32];
    __shared__ int ccount[16], coff[17], ccur[16];   // colour * 2 + (misplaced-centre box ? 1 : 0), see aivs_fps_kernel
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = min(cnt ? cnt[p] : cap, cap);
    const int pn = point_num ? point_num[p] : point_num_all;
    const double* P = pts + (size_t)p * cap * 3;
    const AsLayout L = as_layout(cap);
    float* xf = reinterpret_cast<float*>(as_smem + L.xf);
    float* yf = reinterpret_cast<float*>(as_smem + L.yf);
    float* zf = reinterpret_cast<float*>(as_smem + L.zf);
    unsigned short* sample = reinterpret_cast<unsigned short*>(as_smem + L.sample);
    unsigned* mkey = reinterpret_cast<unsigned*>(as_smem + L.mkey);
    float* md = reinterpret_cast<float*>(as_smem + L.md);
    int* start = reinterpret_cast<int*>(as_smem + L.start);
    int* selc = reinterpret_cast<int*>(as_smem + L.selc);
    unsigned short* sel = reinterpret_cast<unsigned short*>(as_smem + L.sel);
    unsigned short* centre = reinterpret_cast<unsigned short*>(as_smem + L.centre);
    unsigned short* sbox = reinterpret_cast<unsigned short*>(as_smem + L.sbox);
    unsigned short* quota = reinterpret_cast<unsigned short*>(as_smem + L.quota);
    unsigned short* clist = reinterpret_cast<unsigned short*>(as_smem + L.clist);
    unsigned char* lab = as_smem + L.lab;
    unsigned long long* key64 = reinterpret_cast<unsigned long long*>(as_smem + L.key64);
    float* d2s = reinterpret_cast<float*>(as_smem + L.d2s);
    unsigned char* keep = as_smem + L.keep;
    if (n < 1) { if (tid == 0) { out_cnt[p] = 0; atomicExch(bad, 1); } return; }

    // ---- border (pointPipeline_Border) + float copies
    {
        double lmin[3] = {INFINITY, INFINITY, INFINITY}, lmax[3] = {-INFINITY, -INFINITY, -INFINITY};
        for (int i = tid; i < n; i += AS_THREADS) {
            const double q0 = P[3 * i], q1 = P[3 * i + 1], q2 = P[3 * i + 2];
            xf[i] = (float)q0; yf[i] = (float)q1; zf[i] = (float)q2;
            lmin[0] = fmin(lmin[0], q0); lmax[0] = fmax(lmax[0], q0);
            lmin[1] = fmin(lmin[1], q1); lmax[1] = fmax(lmax[1], q1);
            lmin[2] = fmin(lmin[2], q2); lmax[2] = fmax(lmax[2], q2);
        }
#pragma unroll
        for (int d = 0; d < 3; ++d)
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                lmin[d] = fmin(lmin[d], __shfl_xor_sync(KSS_FULL, lmin[d], o));
                lmax[d] = fmax(lmax[d], __shfl_xor_sync(KSS_FULL, lmax[d], o));
            }
        if (lane == 0) for (int d = 0; d < 3; ++d) { red[d][warp] = lmin[d]; red[3 + d][warp] = lmax[d]; }
        for (int i = tid; i < AS_SB; i += AS_THREADS) { start[i] = 0; selc[i] = 0; }
        for (int i = tid; i < n; i += AS_THREADS) lab[i] = 0;
        if (tid < 16) { ccount[tid] = 0; ccur[tid] = 0; }
        __syncthreads();
        if (tid == 0) {
            AivsGrid g;
            double dis[3];
            for (int d = 0; d < 3; ++d) {
                double a = red[d][0], b = red[3 + d][0];
                for (int w = 1; w < AS_THREADS / 32; ++w) { a = fmin(a, red[d][w]); b = fmax(b, red[3 + d][w]); }
                g.mn[d] = a; dis[d] = fabs(__dsub_rn(b, a));
            }
            g.n = n; g.point_num = pn;
            const int boxNum = aivs_box_scale(n);
            double large = dis[0];
            if (large < dis[1]) large = dis[1];
            if (large < dis[2]) large = dis[2];
            g.unit = __ddiv_rn(large, (double)boxNum);
            int num[3];
            for (int d = 0; d < 3; ++d) {
                const double nd = __ddiv_rn(dis[d], g.unit);
                num[d] = (int)nd;
                if (nd > (double)num[d]) num[d]++;
            }
            g.nx = num[0]; g.ny = num[1]; g.nz = num[2];
            const long long nb = (long long)g.nx * g.ny * g.nz + 1;
            g.nbox = (int)nb; g.n_samples = 0; g.pad = 0;
            if (!(large > 0.0) || nb > AS_SB - 2 || g.nx < 1 || g.ny < 1 || g.nz < 1) { g.nbox = 0; atomicExch(bad, 1); out_cnt[p] = 0; }
            g_s = g;
        }
        __syncthreads();
    }
    const AivsGrid g = g_s;
    if (g.nbox == 0) return;
    const int nbox = g.nbox;
    int NP = 1; while (NP < n) NP <<= 1;

    // ---- BallRegion_BoxInput: counting sort by box (counts -> scan -> scatter), then every box's short list is put
    //      into ascending index order (the order push_back gives) by one thread; mkey[pos] = point index
    for (int i = tid; i < n; i += AS_THREADS) {
        int b = aivs_box_of(g, P + 3 * (size_t)i);
        if (b < 0 || b >= nbox) b = nbox; else atomicAdd(&start[b], 1);
        sel[i] = (unsigned short)b;                               // sel is free until the sampling starts
    }
    __syncthreads();
    as_block_excl_scan(start, nbox + 1, wsum);
    for (int i = tid; i < n; i += AS_THREADS) {
        const int b = sel[i];
        if (b < nbox) mkey[start[b] + atomicAdd(&selc[b], 1)] = (unsigned)i;      // selc: zeroed above, zeroed again below
    }
    __syncthreads();
    for (int b = 1 + tid; b < nbox; b += AS_THREADS) {
        const int s0 = start[b], m = start[b + 1] - s0;
        selc[b] = 0;
        for (int i = 1; i < m; ++i) {                              // insertion sort (lists hold a few points)
            const unsigned v = mkey[s0 + i];
            int j = i - 1;
            while (j >= 0 && mkey[s0 + j] > v) { mkey[s0 + j + 1] = mkey[s0 + j]; --j; }
            mkey[s0 + j + 1] = v;
        }
    }
    __syncthreads();

    // ---- per box: centre-nearest member, quota, colour census
    for (int b = 1 + tid; b < nbox; b += AS_THREADS) {
        const int s = start[b], m = start[b + 1] - s;
        if (m == 0) { quota[b] = 0; continue; }
        double c[3];
        aivs_center(g, b, c);
        double best = 9999.0; int bi = -1;
        for (int i = 0; i < m; ++i) {
            const double* q = P + 3 * (size_t)(mkey[s + i] & 2047u);
            const double dx = __dsub_rn(c[0], q[0]), dy = __dsub_rn(c[1], q[1]), dz = __dsub_rn(c[2], q[2]);
            const double dm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            if (best > dm) { best = dm; bi = i; }
        }
        centre[b] = bi < 0 ? 0xffffu : (unsigned short)bi;
        const double rate = __ddiv_rn((double)pn, (double)n);
        const double sb = __dmul_rn((double)m, rate);
        const int t = (int)sb;
        const int qv = (__dsub_rn(sb, (double)t) > 0.2) ? t + 1 : t;
        quota[b] = (unsigned short)min(qv, 65535);
        if (qv > 0) {
            const int b0 = b - 1;
            atomicAdd(&ccount[aivs_colour(b0 % g.nx + 1, (b0 / g.nx) % g.ny + 1, b0 / (g.nx * g.ny) + 1) * 2 +
                              ((b % (g.nx * g.ny)) == 0 ? 1 : 0)], 1);
        }
    }
    __syncthreads();
    if (tid == 0) { int o = 0; for (int c = 0; c < 16; ++c) { coff[c] = o; o += ccount[c]; } coff[16] = o; }
    __syncthreads();
    for (int b = 1 + tid; b < nbox; b += AS_THREADS) {
        if (start[b + 1] == start[b] || quota[b] == 0) continue;
        const int b0 = b - 1;
        const int c = aivs_colour(b0 % g.nx + 1, (b0 / g.nx) % g.ny + 1, b0 / (g.nx * g.ny) + 1) * 2 +
                      ((b % (g.nx * g.ny)) == 0 ? 1 : 0);
        clist[coff[c] + atomicAdd(&ccur[c], 1)] = (unsigned short)b;
    }
    __syncthreads();

    // ---- AIVS_Voroni_OpenMP_KNN: colour by colour, GL lanes per box
    {
        const int grp = tid / GL, gl = tid % GL, ngrp = AS_THREADS / GL;
        const unsigned gmask = GL == 32 ? 0xffffffffu : (((1u << GL) - 1u) << ((lane / GL) * GL));
        const double radius = __ddiv_rn(__dmul_rn(g.unit, 3.0), 4.0);
        for (int col = 0; col < 16; ++col) {
            // even phase: the colour's boxes in parallel; odd phase: its misplaced-centre boxes (index multiple of
            // nx * ny, see aivs_fps_kernel) one after the other in ascending index, by group 0
            const bool serial = col & 1;
            if (ccount[col] == 0) continue;                              // nothing of this kind: skip the barrier too (uniform)
            const int nb_col = serial ? g.nz : ccount[col];
            for (int gi = serial ? (grp == 0 ? 0 : nb_col) : grp; gi < nb_col; gi += serial ? 1 : ngrp) {
                int b;
                if (serial) {
                    b = g.nx * g.ny * (gi + 1);
                    if (b >= nbox || aivs_colour(g.nx, g.ny, gi + 1) != (col >> 1) || start[b + 1] == start[b] || quota[b] == 0) continue;
                } else {
                    b = clist[coff[col] + gi];
                }
                const int s = start[b], m = start[b + 1] - s, simNum = quota[b];
                double pc[3];
                aivs_center(g, b, pc);
                const double lo0 = __dsub_rn(pc[0], radius), hi0 = __dadd_rn(pc[0], radius);
                const double lo1 = __dsub_rn(pc[1], radius), hi1 = __dadd_rn(pc[1], radius);
                const double lo2 = __dsub_rn(pc[2], radius), hi2 = __dadd_rn(pc[2], radius);
                const float lf0 = (float)lo0, hf0 = (float)hi0, lf1 = (float)lo1, hf1 = (float)hi1, lf2 = (float)lo2, hf2 = (float)hi2;
                const int z_num = b / (g.nx * g.ny) + 1;
                const int leveZ = b % (g.nx * g.ny);
                const int y_num = leveZ / g.nx + 1;
                const int x_num = leveZ % g.nx;                              // no wrap fix (B10)
                for (int t = gl; t < m; t += GL) md[s + t] = INFINITY;
                bool any_seed = false;
                for (int xi = x_num - 1; xi <= x_num + 1; ++xi) {
                    if ((xi < x_num && !(x_num > 1)) || (xi > x_num && !(x_num < g.nx))) continue;
                    for (int yj = y_num - 1; yj <= y_num + 1; ++yj) {
                        if ((yj < y_num && !(y_num > 1)) || (yj > y_num && !(y_num < g.ny))) continue;
                        for (int zk = z_num - 1; zk <= z_num + 1; ++zk) {
                            if ((zk < z_num && !(z_num > 1)) || (zk > z_num && !(z_num < g.nz))) continue;
                            if (xi == x_num && yj == y_num && zk == z_num) continue;
                            const int nb = xi + (yj - 1) * g.nx + (zk - 1) * g.nx * g.ny;
                            if (nb >= nbox || nb < 0) continue;
                            const int ns = start[nb], nm = start[nb + 1] - ns;
                            for (int l0 = 0; l0 < nm; l0 += GL) {
                                const int l = l0 + gl;
                                bool ok = false; int pt = 0;
                                if (l < nm) {
                                    pt = (int)(mkey[ns + l] & 2047u);
                                    if (lab[pt]) {
                                        // rounding to float is monotonic: a float coordinate strictly inside (outside)
                                        // the float-rounded bounds decides the double comparison; only equality needs it
                                        const float a0 = xf[pt], a1 = yf[pt], a2 = zf[pt];
                                        if (a0 > hf0 || a0 < lf0 || a1 > hf1 || a1 < lf1 || a2 > hf2 || a2 < lf2) ok = false;
                                        else if (a0 < hf0 && a0 > lf0 && a1 < hf1 && a1 > lf1 && a2 < hf2 && a2 > lf2) ok = true;
                                        else {
                                            const double* q = P + 3 * (size_t)pt;
                                            const double q0 = q[0], q1 = q[1], q2 = q[2];
                                            ok = q0 <= hi0 && q0 >= lo0 && q1 <= hi1 && q1 >= lo1 && q2 <= hi2 && q2 >= lo2;
                                        }
                                    }
                                }
                                unsigned bal = __ballot_sync(gmask, ok) & gmask;
                                while (bal) {
                                    const int src = __ffs(bal) - 1; bal &= bal - 1;
                                    const int spt = __shfl_sync(gmask, pt, src);
                                    const float sx = xf[spt], sy = yf[spt], sz = zf[spt];
                                    for (int t = gl; t < m; t += GL) {
                                        const int w = (int)(mkey[s + t] & 2047u);
                                        const float d = __fsqrt_rn(d2_rn(xf[w], yf[w], zf[w], sx, sy, sz));
                                        if (d < md[s + t]) md[s + t] = d;
                                    }
                                    any_seed = true;
                                }
                            }
                        }
                    }
                }
                for (int t = gl; t < m; t += GL) if (md[s + t] == INFINITY) md[s + t] = 9999.0f;
                int sampled = 0;
                if (!any_seed) {
                    const int ci = centre[b];
                    if (ci != 0xffff && ci < m) {
                        const int cpt = (int)(mkey[s + ci] & 2047u);
                        const float sx = xf[cpt], sy = yf[cpt], sz = zf[cpt];
                        for (int t = gl; t < m; t += GL) {
                            const int w = (int)(mkey[s + t] & 2047u);
                            md[s + t] = t == ci ? 0.0f : __fsqrt_rn(d2_rn(xf[w], yf[w], zf[w], sx, sy, sz));
                        }
                        if (gl == 0) { sel[s] = (unsigned short)cpt; lab[cpt] = 1; }
                        sampled = 1;
                    }
                }
                while (sampled < simNum) {
                    float bv = 0.0f; int bt = -1;
                    for (int t = gl; t < m; t += GL) { const float v = md[s + t]; if (v > bv) { bv = v; bt = t; } }
#pragma unroll
                    for (int o = GL / 2; o >= 1; o >>= 1) {
                        const float ov = __shfl_xor_sync(gmask, bv, o);
                        const int ot = __shfl_xor_sync(gmask, bt, o);
                        if (ov > bv || (ov == bv && ot < bt)) { bv = ov; bt = ot; }
                    }
                    if (bt < 0) break;
                    const int ppt = (int)(mkey[s + bt] & 2047u);
                    if (gl == 0) { lab[ppt] = 1; sel[s + sampled] = (unsigned short)ppt; }
                    ++sampled;
                    const float sx = xf[ppt], sy = yf[ppt], sz = zf[ppt];
                    for (int t = gl; t < m; t += GL) {
                        const int w = (int)(mkey[s + t] & 2047u);
                        const float d = __fsqrt_rn(d2_rn(sx, sy, sz, xf[w], yf[w], zf[w]));
                        if (d < md[s + t]) md[s + t] = d;
                    }
                }
                if (gl == 0) selc[b] = sampled;
                __syncwarp(gmask);
            }
            __syncthreads();
        }
    }

    // ---- samples in box order
    const int S = as_block_excl_scan(selc, nbox + 1, wsum);
    for (int b = 1 + tid; b < nbox; b += AS_THREADS) {
        const int dst = selc[b], c = selc[b + 1] - dst, src = start[b];
        for (int r = 0; r < c; ++r) { sample[dst + r] = sel[src + r]; sbox[dst + r] = (unsigned short)b; }
    }
    __syncthreads();

    // ---- AIVS_AccurateCut_Optimization
    int dT = S - pn;
    const bool trim = dT > 0 && S >= 3;
    if (trim) {
        int NP2 = 1; while (NP2 < S) NP2 <<= 1;
        // K = 3 nearest samples in ascending (d2, position).  Candidates are visited in ascending position, so a later
        // one replaces an entry only when strictly nearer.  First the samples of the 27 boxes around the sample's own
        // box (ascending box index = ascending position); a sample outside that block is at least one box width away
        // along some axis, so when the third distance found is below that (minus the float narrowing of the
        // coordinates, `lim2`), the three are the global answer; otherwise all samples are scanned.
        float lim2 = 0.0f;
        {
            double amax = 0.0;
            for (int d = 0; d < 3; ++d) amax = fmax(amax, fmax(fabs(g.mn[d]), fabs(g.mn[d] + g.unit * (double)(d == 0 ? g.nx : d == 1 ? g.ny : g.nz))));
            const double lim = g.unit * 0.999 - amax * 1.0e-6;
            if (lim > 0.0) lim2 = (float)(lim * lim * 0.999);
        }
        for (int i0 = 0; i0 < S; i0 += AS_THREADS) {
            const int i = i0 + tid;
            if (i >= S) break;
            const int me = sample[i];
            const float x = xf[me], y = yf[me], z = zf[me];
            float e0 = INFINITY, e1 = INFINITY, e2 = INFINITY; int j1 = 0, j0 = 0;
            const int b0 = (int)sbox[i] - 1;
            const int bi = b0 % g.nx + 1, bj = (b0 / g.nx) % g.ny + 1, bk = b0 / (g.nx * g.ny) + 1;
            for (int nk = max(1, bk - 1); nk <= min(g.nz, bk + 1); ++nk)
                for (int nj = max(1, bj - 1); nj <= min(g.ny, bj + 1); ++nj) {
                    const int ilo = max(1, bi - 1), ihi = min(g.nx, bi + 1);
                    const int row = g.nx * (nj - 1) + g.nx * g.ny * (nk - 1);
                    const int jlo = selc[row + ilo], jhi = selc[row + ihi + 1];        // consecutive boxes = one run of samples
                    for (int j = jlo; j < jhi; ++j) {
                        const int o = sample[j];
                        const float d = d2_rn(x, y, z, xf[o], yf[o], zf[o]);
                        {   // sorted insert with selects only (branches here diverge on almost every candidate)
                            const bool p0 = d < e0, p1 = d < e1, p2 = d < e2;
                            e2 = p1 ? e1 : (p2 ? d : e2);
                            j1 = p0 ? j0 : (p1 ? j : j1);
                            e1 = p0 ? e0 : (p1 ? d : e1);
                            j0 = p0 ? j : j0;
                            e0 = p0 ? d : e0;
                        }
                    }
                }
            if (!(e2 < lim2)) {
                e0 = INFINITY; e1 = INFINITY; e2 = INFINITY; j1 = 0; j0 = 0;
                for (int j = 0; j < S; ++j) {
                    const int o = sample[j];
                    const float d = d2_rn(x, y, z, xf[o], yf[o], zf[o]);
                    if (d < e2) {
                        if (d < e0) { e2 = e1; e1 = e0; j1 = j0; e0 = d; j0 = j; }
                        else if (d < e1) { e2 = e1; e1 = d; j1 = j; }
                        else e2 = d;
                    }
                }
            }
            const float d1 = __fsqrt_rn(e1);
            key64[i] = ((unsigned long long)__float_as_uint(d1) << 32) | ((unsigned long long)i << 16) | (unsigned long long)j1;
            d2s[i] = __fsqrt_rn(e2);
            keep[i] = 1;
        }
        for (int i = S + tid; i < NP2; i += AS_THREADS) key64[i] = ~0ull;
        as_bitonic(key64, NP2);
        if (tid == 0) {
            for (int pos = 0; pos < S && dT > 0; ++pos) {
                const unsigned long long key = key64[pos];
                if (!((double)__uint_as_float((unsigned)(key >> 32)) < 9999.0)) break;
                const int b1 = (int)((key >> 16) & 0xffffu), b2 = (int)(key & 0xffffu);
                if (!keep[b1] || !keep[b2]) continue;
                keep[((double)d2s[b1] > (double)d2s[b2]) ? b2 : b1] = 0;
                --dT;
            }
        }
        __syncthreads();
    }

    // ---- ordered output
    {
        const int per = (S + AS_THREADS - 1) / AS_THREADS;
        const int lo = min(S, tid * per), hi = min(S, lo + per);
        int c = 0;
        for (int i = lo; i < hi; ++i) c += (!trim || keep[i]) ? 1 : 0;
        int x = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(KSS_FULL, x, o); if (lane >= o) x += y; }
        if (lane == 31) wsum[warp] = x;
        __syncthreads();
        int pos = x - c, total = 0;
        for (int w = 0; w < AS_THREADS / 32; ++w) { if (w < warp) pos += wsum[w]; total += wsum[w]; }
        for (int i = lo; i < hi; ++i)
            if (!trim || keep[i]) {
                if (pos < out_cap) {
                    const int me = sample[i];
                    const double* q = P + 3 * (size_t)me;
                    double* o = out + ((size_t)p * out_cap + pos) * 3;
                    o[0] = q[0]; o[1] = q[1]; o[2] = q[2];
                    if (out_idx) out_idx[(size_t)p * out_cap + pos] = me;
                }
                ++pos;
            }
        if (tid == 0) { out_cnt[p] = min(total, out_cap); if (total > out_cap) atomicExch(bad, 3); }
    }
}

// pNumber = min(|S|, |T|) / 2 capped at 2000 (KSS_ICP.hpp:53-67)
__global__ void aivs_pnumber_kernel(int P, const int* cnt_S, int cap_S, const int* cnt_T, int cap_T, int* pn) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int a = cnt_S ? cnt_S[p] : cap_S, b = cnt_T ? cnt_T[p] : cap_T;
    int v = (a > b ? b : a) / 2;
    if (v > 2000) v = 2000;
    pn[p] = v;
}

// ---------------------------------------------------------------- host
int aivs_box_cap(int cap) {
    int bn;
    if (cap < 10000) bn = 10; else if (cap < 50000) bn = 20; else if (cap < 100000) bn = 30;
    else if (cap < 500000) bn = 40; else if (cap < 1000000) bn = 50; else bn = (int)std::cbrt((double)cap / 8.0) + 1;
    return (bn + 1) * (bn + 1) * (bn + 1) + 2;
}

int aivs_simplify_device(cudaStream_t st, long long* launches, int P, const double* d_pts, const int* d_cnt, int cap,
                         const int* d_point_num, int point_num_all, double* d_out, int out_cap, int* d_out_cnt,
                         int* d_out_idx, int* d_bad, const DevAlloc& alloc, const char* tag) {
    if (cap <= AS_MAX && !getenv("KSS_AIVS_GENERAL")) {
        const AsLayout L = as_layout(cap);
        const char* e = getenv("KSS_AIVS_GL");
        const int gl = e ? atoi(e) : AS_GL;
        auto kern = gl == 4 ? aivs_small_kernel<4> : gl == 8 ? aivs_small_kernel<8> : gl == 32 ? aivs_small_kernel<32> : aivs_small_kernel<16>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total) != cudaSuccess) return KSS_ERR_CUDA;
        kern<<<P, AS_THREADS, L.total, st>>>(d_pts, d_cnt, cap, d_point_num, point_num_all, d_out, out_cap, d_out_cnt, d_out_idx,
                                             d_bad);
        *launches += 1;
        return cudaGetLastError() == cudaSuccess ? KSS_OK : KSS_ERR_CUDA;
    }
    const int bcap = aivs_box_cap(cap);
    auto get = [&](const char* name, size_t bytes, void** out) {
        char nm[64]; snprintf(nm, sizeof(nm), "aivs_%s_%s", tag, name);
        return alloc(nm, bytes, out);
    };
    AivsGrid* grids; int *box_of, *box_cnt, *box_start, *cursor, *members, *center_pos, *quota, *sel, *sel_cnt, *sel_start, *sample;
    unsigned char* selected; double* mind; unsigned long long *key1, *ext; float* dis2;
    int *biglist, *bigcnt;
    int r = 0;
    r |= get("biglist", sizeof(int) * (size_t)P * 8 * FPS_BIGCAP, (void**)&biglist);
    r |= get("bigcnt", sizeof(int) * (size_t)P * 8, (void**)&bigcnt);
    r |= get("ext", sizeof(unsigned long long) * (size_t)P * 6, (void**)&ext);
    r |= get("grids", sizeof(AivsGrid) * (size_t)P, (void**)&grids);
    r |= get("box_of", sizeof(int) * (size_t)P * cap, (void**)&box_of);
    r |= get("box_cnt", sizeof(int) * (size_t)P * bcap, (void**)&box_cnt);
    r |= get("box_start", sizeof(int) * (size_t)P * (bcap + 1), (void**)&box_start);
    r |= get("cursor", sizeof(int) * (size_t)P * bcap, (void**)&cursor);
    r |= get("members", sizeof(int) * (size_t)P * cap, (void**)&members);
    r |= get("center", sizeof(int) * (size_t)P * bcap, (void**)&center_pos);
    r |= get("quota", sizeof(int) * (size_t)P * bcap, (void**)&quota);
    r |= get("sel", sizeof(int) * (size_t)P * cap, (void**)&sel);
    r |= get("sel_cnt", sizeof(int) * (size_t)P * bcap, (void**)&sel_cnt);
    r |= get("sel_start", sizeof(int) * (size_t)P * (bcap + 1), (void**)&sel_start);
    r |= get("sample", sizeof(int) * (size_t)P * cap, (void**)&sample);
    r |= get("selected", (size_t)P * cap, (void**)&selected);
    r |= get("mind", sizeof(double) * (size_t)P * cap, (void**)&mind);
    const int smax = std::min(cap, AIVS_MAX_SAMPLES);
    r |= get("key1", sizeof(unsigned long long) * (size_t)P * cap, (void**)&key1);
    r |= get("dis2", sizeof(float) * (size_t)P * cap, (void**)&dis2);
    if (r) return KSS_ERR_NOMEM;
    cudaMemsetAsync(box_cnt, 0, sizeof(int) * (size_t)P * bcap, st);
    cudaMemsetAsync(sel_cnt, 0, sizeof(int) * (size_t)P * bcap, st);
    cudaMemsetAsync(selected, 0, (size_t)P * cap, st);
    cudaMemsetAsync(bigcnt, 0, sizeof(int) * (size_t)P * 8, st);
    cudaMemset2DAsync(ext, 48, 0xff, 24, P, st);                      // min keys: all ones
    cudaMemset2DAsync(ext + 3, 48, 0x00, 24, P, st);                  // max keys: zero
    aivs_extent_kernel<<<dim3(std::max(1, std::min((cap + 2047) / 2048, 1184 / std::max(P, 1) + 1)), P), 256, 0, st>>>(d_pts, d_cnt, cap, ext);
    aivs_grid_kernel<<<(P + 127) / 128, 128, 0, st>>>(ext, d_cnt, cap, d_point_num, point_num_all, bcap, P, grids, d_bad);
    const dim3 gp((cap + 255) / 256, P), gb((bcap + 255) / 256, P), gw((bcap + 7) / 8, P);
    aivs_count_kernel<<<gp, 256, 0, st>>>(d_pts, cap, grids, bcap, box_of, box_cnt);
    aivs_scan_kernel<<<P, 1024, 0, st>>>(grids, bcap, box_cnt, box_start, cursor);
    aivs_fill_kernel<<<gp, 256, 0, st>>>(cap, grids, bcap, box_of, cursor, members);
    aivs_box_kernel<<<gw, 256, 0, st>>>(d_pts, cap, grids, bcap, box_start, members, center_pos, quota);
    static const bool fps_warp = getenv("KSS_AIVS_FPS_WARP") != nullptr;          // A/B: the even phases through global memory too (aivs_fps_box)
    if (!fps_warp && (cudaFuncSetAttribute(aivs_fps_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FPS_SMEM) != cudaSuccess ||
                      cudaFuncSetAttribute(aivs_fps_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FPS_MCAP * 20) != cudaSuccess)) return KSS_ERR_CUDA;
    const bool use_big = cap >= 250000;                                            // (boxes beyond FPS_WCAP members are the rule there)
    for (int c = 0; c < 16; ++c)
        if (!(c & 1) && !fps_warp) {
            aivs_fps_smem_kernel<<<gw, 256, FPS_SMEM, st>>>(c >> 1, d_pts, cap, grids, bcap, box_start, members, center_pos,
                                                            quota, selected, mind, sel, sel_cnt, use_big ? biglist : nullptr, bigcnt);
            if (use_big) aivs_fps_big_kernel<<<dim3(std::max(1, 296 / std::max(P, 1)), P), 256, FPS_MCAP * 20, st>>>(c >> 1, d_pts, cap, grids, bcap, box_start, members, center_pos,
                                                                                   quota, selected, mind, sel, sel_cnt, biglist, bigcnt);
        } else
            aivs_fps_kernel<<<(c & 1) ? dim3(1, P) : gw, (c & 1) ? 32 : 256, 0, st>>>(c, d_pts, cap, grids, bcap, box_start, members, center_pos,
                                                                                   quota, selected, mind, sel, sel_cnt);
    aivs_scan_kernel<<<P, 1024, 0, st>>>(grids, bcap, sel_cnt, sel_start, nullptr);
    aivs_gather_kernel<<<gb, 256, 0, st>>>(cap, grids, bcap, box_start, sel, sel_cnt, sel_start, sample);
    aivs_k3_kernel<<<dim3((smax + 255) / 256, P), 256, 0, st>>>(d_pts, cap, grids, bcap, sel_start, sample, key1, dis2);
    int np2 = 1; while (np2 < smax) np2 <<= 1;
    const size_t smem = (size_t)np2 * (8 + 4 + 1);
    if (cudaFuncSetAttribute(aivs_cut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return KSS_ERR_CUDA;
    aivs_cut_kernel<<<P, 512, smem, st>>>(d_pts, cap, grids, bcap, sel_start, sample, key1, dis2, np2, d_out, out_cap, d_out_cnt,
                                          d_out_idx, d_bad);
    *launches += (fps_warp || !use_big) ? 26 : 34;
    return cudaGetLastError() == cudaSuccess ? KSS_OK : KSS_ERR_CUDA;
}

int aivs_pnumber_device(cudaStream_t st, long long* launches, int P, const int* cnt_S, int cap_S, const int* cnt_T, int cap_T, int* pn) {
    aivs_pnumber_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, cnt_S, cap_S, cnt_T, cap_T, pn);
    *launches += 1;
    return cudaGetLastError() == cudaSuccess ? KSS_OK : KSS_ERR_CUDA;
}

}  // namespace kss
