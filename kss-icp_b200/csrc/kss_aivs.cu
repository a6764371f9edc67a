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

// ---------------------------------------------------------------- 1. border, grid geometry (one CTA per cloud)
__global__ void __launch_bounds__(256)
aivs_grid_kernel(const double* __restrict__ pts, const int* __restrict__ cnt, int cap, const int* __restrict__ point_num,
                 int point_num_all, int bcap, AivsGrid* __restrict__ grids, int* __restrict__ bad) {
    __shared__ unsigned long long kmin[3], kmax[3];      // (ordered double bits, index) packed: value high bits not enough ->
    __shared__ double vmin[3], vmax[3];
    __shared__ int imin[3], imax[3];
    const int p = blockIdx.x;
    const int n = cnt ? cnt[p] : cap;
    const double* P = pts + (size_t)p * cap * 3;
    // two steps: extreme values, then the lowest index attaining them (= first strict extreme in a forward scan)
    double lmin[3] = {INFINITY, INFINITY, INFINITY}, lmax[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        for (int d = 0; d < 3; ++d) { const double v = P[3 * (size_t)i + d]; lmin[d] = fmin(lmin[d], v); lmax[d] = fmax(lmax[d], v); }
    if (threadIdx.x < 3) { kmin[threadIdx.x] = 0xffffffffffffffffull; kmax[threadIdx.x] = 0ull; imin[threadIdx.x] = 0x7fffffff; imax[threadIdx.x] = 0x7fffffff; }
    __syncthreads();
    for (int d = 0; d < 3; ++d) {
        unsigned long long a = (unsigned long long)__double_as_longlong(lmin[d]); a = (a >> 63) ? ~a : (a | 0x8000000000000000ull);
        unsigned long long b = (unsigned long long)__double_as_longlong(lmax[d]); b = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
        atomicMin(&kmin[d], a); atomicMax(&kmax[d], b);
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long a = kmin[threadIdx.x]; a = (a >> 63) ? (a & 0x7fffffffffffffffull) : ~a;
        unsigned long long b = kmax[threadIdx.x]; b = (b >> 63) ? (b & 0x7fffffffffffffffull) : ~b;
        vmin[threadIdx.x] = __longlong_as_double((long long)a); vmax[threadIdx.x] = __longlong_as_double((long long)b);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
        for (int d = 0; d < 3; ++d) {
            const double v = P[3 * (size_t)i + d];
            if (v == vmin[d]) atomicMin(&imin[d], i);
            if (v == vmax[d]) atomicMin(&imax[d], i);
        }
    __syncthreads();
    if (threadIdx.x == 0) {
        AivsGrid g;
        g.n = n; g.point_num = point_num ? point_num[p] : point_num_all;
        const int boxNum = aivs_box_scale(n);
        double dis[3];
        for (int d = 0; d < 3; ++d) { g.mn[d] = vmin[d]; dis[d] = fabs(__dsub_rn(vmax[d], vmin[d])); }
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

// per box: ascending original index (the reference's push_back order), the centre-nearest member, the quota
__global__ void __launch_bounds__(256)
aivs_box_kernel(const double* __restrict__ pts, int cap, const AivsGrid* __restrict__ grids, int bcap,
                const int* __restrict__ box_start, int* __restrict__ members, int* __restrict__ center_pos,
                int* __restrict__ quota) {
    const int p = blockIdx.y;
    const AivsGrid g = grids[p];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= g.nbox) return;
    const int* st = box_start + (size_t)p * (bcap + 1);
    const int s = st[b], m = st[b + 1] - s;
    int* mem = members + (size_t)p * cap + s;
    for (int i = 1; i < m; ++i) {                              // insertion sort (lists are short)
        const int v = mem[i];
        int j = i - 1;
        while (j >= 0 && mem[j] > v) { mem[j + 1] = mem[j]; --j; }
        mem[j + 1] = v;
    }
    double c[3];
    aivs_center(g, b, c);
    double best = 9999.0; int bi = -1;
    const double* P = pts + (size_t)p * cap * 3;
    for (int i = 0; i < m; ++i) {                              // BoxInput: first strict minimum of the distance to the centre
        const double* q = P + 3 * (size_t)mem[i];
        const double dx = __dsub_rn(c[0], q[0]), dy = __dsub_rn(c[1], q[1]), dz = __dsub_rn(c[2], q[2]);
        const double dm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
        if (best > dm) { best = dm; bi = i; }
    }
    center_pos[(size_t)p * bcap + b] = bi;
    const double rate = __ddiv_rn((double)g.point_num, (double)g.n);
    const double sb = __dmul_rn((double)m, rate);
    const int t = (int)sb;
    quota[(size_t)p * bcap + b] = (__dsub_rn(sb, (double)t) > 0.2) ? t + 1 : t;
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

__global__ void __launch_bounds__(128)
aivs_fps_kernel(int colour, const double* __restrict__ pts, int cap, const AivsGrid* __restrict__ grids, int bcap,
                const int* __restrict__ box_start, const int* __restrict__ members, const int* __restrict__ center_pos,
                const int* __restrict__ quota, unsigned char* __restrict__ selected /* labelG == 0 */,
                double* __restrict__ mind, int* __restrict__ sel, int* __restrict__ sel_cnt) {
    const int p = blockIdx.y;
    const AivsGrid g = grids[p];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < 1 || b >= g.nbox) return;
    // true (i,j,k) of the box, as the colouring loop enumerates them (Method_AIVS_SimPro.hpp:598-602)
    const int b0 = b - 1;
    const int i = b0 % g.nx + 1, j = (b0 / g.nx) % g.ny + 1, k = b0 / (g.nx * g.ny) + 1;
    if (aivs_colour(i, j, k) != colour) return;
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
    // initial distances: to the nearest already-selected point of the neighbour boxes inside the seed cube,
    // or (if there is none) to the box's centre-nearest member, which then becomes the first sample
    for (int t = 0; t < m; ++t) md[t] = INFINITY;
    bool any_seed = false;
    for (int a = 0; a < nxs; ++a) for (int c = 0; c < nys; ++c) for (int e = 0; e < nzs; ++e) {
        if (xs[a] == x_num && ys[c] == y_num && zs[e] == z_num) continue;
        const int nb = xs[a] + (ys[c] - 1) * g.nx + (zs[e] - 1) * g.nx * g.ny;
        if (nb >= g.nbox || nb < 0) continue;
        const int ns = st[nb], nm = st[nb + 1] - ns;
        for (int l = 0; l < nm; ++l) {
            const int pt = mem[ns + l];
            if (!lab[pt]) continue;
            const double* q = P + 3 * (size_t)pt;
            if (!(q[0] <= __dadd_rn(pc[0], radius) && q[0] >= __dsub_rn(pc[0], radius) && q[1] <= __dadd_rn(pc[1], radius) &&
                  q[1] >= __dsub_rn(pc[1], radius) && q[2] <= __dadd_rn(pc[2], radius) && q[2] >= __dsub_rn(pc[2], radius))) continue;
            any_seed = true;
            const float sx = (float)q[0], sy = (float)q[1], sz = (float)q[2];
            for (int t = 0; t < m; ++t) {
                const double* w = P + 3 * (size_t)mem[s + t];
                const double d = (double)__fsqrt_rn(d2_rn((float)w[0], (float)w[1], (float)w[2], sx, sy, sz));
                if (d < md[t]) md[t] = d;
            }
        }
    }
    for (int t = 0; t < m; ++t) if (md[t] == INFINITY) md[t] = 9999.0;
    int sampled = 0;
    if (!any_seed) {
        const int ci = center_pos[(size_t)p * bcap + b];
        if (ci >= 0 && ci < m) {
            const double* q = P + 3 * (size_t)mem[s + ci];
            const float sx = (float)q[0], sy = (float)q[1], sz = (float)q[2];
            for (int t = 0; t < m; ++t) {
                const double* w = P + 3 * (size_t)mem[s + t];
                md[t] = (double)__fsqrt_rn(d2_rn((float)w[0], (float)w[1], (float)w[2], sx, sy, sz));
            }
            md[ci] = 0.0; out[sampled++] = mem[s + ci]; lab[mem[s + ci]] = 1;
        }
    }
    while (sampled < simNum) {
        int pick = -1; double mx = 0.0;
        for (int t = 0; t < m; ++t) if (md[t] > mx) { pick = t; mx = md[t]; }     // first strict maximum
        if (pick == -1) break;
        md[pick] = 0.0; lab[mem[s + pick]] = 1; out[sampled++] = mem[s + pick];
        const double* q = P + 3 * (size_t)mem[s + pick];
        const float sx = (float)q[0], sy = (float)q[1], sz = (float)q[2];
        for (int t = 0; t < m; ++t) {
            const double* w = P + 3 * (size_t)mem[s + t];
            const double d = (double)__fsqrt_rn(d2_rn(sx, sy, sz, (float)w[0], (float)w[1], (float)w[2]));
            if (d < md[t]) md[t] = d;
        }
    }
    sel_cnt[(size_t)p * bcap + b] = sampled;
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
    const int bcap = aivs_box_cap(cap);
    auto get = [&](const char* name, size_t bytes, void** out) {
        char nm[64]; snprintf(nm, sizeof(nm), "aivs_%s_%s", tag, name);
        return alloc(nm, bytes, out);
    };
    AivsGrid* grids; int *box_of, *box_cnt, *box_start, *cursor, *members, *center_pos, *quota, *sel, *sel_cnt, *sel_start, *sample;
    unsigned char* selected; double* mind; unsigned long long* key1; float* dis2;
    int r = 0;
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
    aivs_grid_kernel<<<P, 256, 0, st>>>(d_pts, d_cnt, cap, d_point_num, point_num_all, bcap, grids, d_bad);
    const dim3 gp((cap + 255) / 256, P), gb((bcap + 255) / 256, P);
    aivs_count_kernel<<<gp, 256, 0, st>>>(d_pts, cap, grids, bcap, box_of, box_cnt);
    aivs_scan_kernel<<<P, 1024, 0, st>>>(grids, bcap, box_cnt, box_start, cursor);
    aivs_fill_kernel<<<gp, 256, 0, st>>>(cap, grids, bcap, box_of, cursor, members);
    aivs_box_kernel<<<gb, 256, 0, st>>>(d_pts, cap, grids, bcap, box_start, members, center_pos, quota);
    for (int c = 0; c < 8; ++c)
        aivs_fps_kernel<<<dim3((bcap + 127) / 128, P), 128, 0, st>>>(c, d_pts, cap, grids, bcap, box_start, members, center_pos,
                                                                  quota, selected, mind, sel, sel_cnt);
    aivs_scan_kernel<<<P, 1024, 0, st>>>(grids, bcap, sel_cnt, sel_start, nullptr);
    aivs_gather_kernel<<<gb, 256, 0, st>>>(cap, grids, bcap, box_start, sel, sel_cnt, sel_start, sample);
    aivs_k3_kernel<<<dim3((smax + 255) / 256, P), 256, 0, st>>>(d_pts, cap, grids, bcap, sel_start, sample, key1, dis2);
    int np2 = 1; while (np2 < smax) np2 <<= 1;
    const size_t smem = (size_t)np2 * (8 + 4 + 1);
    static size_t smem_set = 0;
    if (smem > smem_set) {
        if (cudaFuncSetAttribute(aivs_cut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return KSS_ERR_CUDA;
        smem_set = smem;
    }
    aivs_cut_kernel<<<P, 512, smem, st>>>(d_pts, cap, grids, bcap, sel_start, sample, key1, dis2, np2, d_out, out_cap, d_out_cnt,
                                          d_out_idx, d_bad);
    *launches += 17;
    return cudaGetLastError() == cudaSuccess ? KSS_OK : KSS_ERR_CUDA;
}

int aivs_pnumber_device(cudaStream_t st, long long* launches, int P, const int* cnt_S, int cap_S, const int* cnt_T, int cap_T, int* pn) {
    aivs_pnumber_kernel<<<(P + 127) / 128, 128, 0, st>>>(P, cnt_S, cap_S, cnt_T, cap_T, pn);
    *launches += 1;
    return cudaGetLastError() == cudaSuccess ? KSS_OK : KSS_ERR_CUDA;
}

}  // namespace kss
