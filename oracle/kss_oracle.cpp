/*
 * kss_oracle.cpp -- CPU restatement of the KSS-ICP registration hot path.
 * TEST INFRASTRUCTURE ONLY; PARITY UNPINNED except the nearest-neighbour stage, which is checked against real FLANN
 * code (see kss_oracle.h header comment).
 *
 * Build:  g++ -O2 -std=c++17 -ffp-contract=off -pthread -fPIC -shared
 * (-ffp-contract=off is REQUIRED: every expression below is meant to round once
 *  per operator, like MSVC /fp:precise which the reference project uses.)
 *
 * Reference citations are file:line under /root/reference/PS_AIS_Simplification/.
 */
#include "kss_oracle.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#include <atomic>
#include <thread>

namespace {

/* ------------------------------------------------------------------ NN --- */

/* FLANN L2_Simple<float>: result = 0; result += diff*diff per dim */
inline float dist2f(const float* a, const float* b) {
    float dx = a[0] - b[0];
    float dy = a[1] - b[1];
    float dz = a[2] - b[2];
    float r = dx * dx;
    r = r + dy * dy;
    r = r + dz * dz;
    return r;
}

/* lexicographic (d2, index) strict improvement: lowest index among equal d2 */
inline bool better(float d, int i, float bd, int bi) {
    return d < bd || (d == bd && i < bi);
}

struct KdTree {
    struct Node { int lo, hi, dim, left, right; float split; };
    const float* pts = nullptr;
    int n = 0;
    std::vector<int> order;
    std::vector<Node> nodes;
    static constexpr int LEAF = 12;

    void build(const float* p, int count) {
        pts = p; n = count;
        order.resize(n);
        for (int i = 0; i < n; ++i) order[i] = i;
        nodes.clear();
        nodes.reserve(2 * (n / LEAF + 2));
        if (n > 0) build_rec(0, n);
    }
    int build_rec(int lo, int hi) {
        int id = (int)nodes.size();
        nodes.push_back(Node{lo, hi, -1, -1, -1, 0.f});
        if (hi - lo <= LEAF) return id;
        float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int k = lo; k < hi; ++k)
            for (int d = 0; d < 3; ++d) {
                float v = pts[3 * order[k] + d];
                mn[d] = std::min(mn[d], v); mx[d] = std::max(mx[d], v);
            }
        int dim = 0;
        if (mx[1] - mn[1] > mx[dim] - mn[dim]) dim = 1;
        if (mx[2] - mn[2] > mx[dim] - mn[dim]) dim = 2;
        if (!(mx[dim] - mn[dim] > 0.f)) return id;              /* all identical: keep as leaf */
        int mid = (lo + hi) / 2;
        std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi,
                         [&](int a, int b) { return pts[3 * a + dim] < pts[3 * b + dim]; });
        float split = pts[3 * order[mid] + dim];
        int l = build_rec(lo, mid);
        int r = build_rec(mid, hi);
        nodes[id].dim = dim; nodes[id].split = split; nodes[id].left = l; nodes[id].right = r;
        return id;
    }
    void search_rec(int id, const float* q, float& bd, int& bi) const {
        const Node& nd = nodes[id];
        if (nd.dim < 0) {
            for (int k = nd.lo; k < nd.hi; ++k) {
                int i = order[k];
                float d = dist2f(q, pts + 3 * i);
                if (better(d, i, bd, bi)) { bd = d; bi = i; }
            }
            return;
        }
        float diff = q[nd.dim] - nd.split;
        int nearc = diff < 0.f ? nd.left : nd.right;
        int farc  = diff < 0.f ? nd.right : nd.left;
        search_rec(nearc, q, bd, bi);
        /* every far-side point p has |q_d - p_d| >= |diff|, and fl() is monotone, so
         * its computed d2 >= fl(diff*diff); equal must still be visited (index tie rule) */
        float plane = diff * diff;
        if (!(plane > bd)) search_rec(farc, q, bd, bi);
    }
    void query(const float* q, int& bi, float& bd) const {
        bd = std::numeric_limits<float>::infinity(); bi = -1;
        if (n > 0) search_rec(0, q, bd, bi);
    }
};

struct NNIndex {
    int method; const float* t; int nt; KdTree tree;
    NNIndex(const float* tgt, int n, int m) : method(m), t(tgt), nt(n) {
        if (method == OKSS_NN_KDTREE) tree.build(tgt, n);
    }
    inline void query(const float* q, int& bi, float& bd) const {
        if (method == OKSS_NN_KDTREE) { tree.query(q, bi, bd); return; }
        bd = std::numeric_limits<float>::infinity(); bi = -1;
        for (int j = 0; j < nt; ++j) {
            float d = dist2f(q, t + 3 * j);
            if (d < bd) { bd = d; bi = j; }      /* first strict minimum = lowest index */
        }
    }
};

/* --------------------------------------------------------------- sums --- */

template <class T, class Get>
T sum_serial(int n, Get get) {
    T s = (T)0;
    for (int i = 0; i < n; ++i) { T v; if (get(i, v)) s = s + v; }
    return s;
}

template <class T>
T butterfly32(T p[32]) {
    for (int off = 16; off >= 1; off >>= 1) {
        T q[32];
        for (int l = 0; l < 32; ++l) q[l] = p[l] + p[l ^ off];
        for (int l = 0; l < 32; ++l) p[l] = q[l];
    }
    return p[0];
}

/* CANON256: per 256-element chunk, 32 strided partials (index order inside a
 * lane) + xor butterfly 16,8,4,2,1; chunk results reduced by the same rule. */
template <class T>
T sum_canon_array(const T* v, int n) {
    if (n <= 256) {
        T p[32];
        for (int l = 0; l < 32; ++l) p[l] = (T)0;
        for (int i = 0; i < n; ++i) p[i & 31] = p[i & 31] + v[i];
        return butterfly32(p);
    }
    int nc = (n + 255) / 256;
    std::vector<T> part(nc);
    for (int c = 0; c < nc; ++c) {
        T p[32];
        for (int l = 0; l < 32; ++l) p[l] = (T)0;
        int hi = std::min(n, (c + 1) * 256);
        for (int i = c * 256; i < hi; ++i) p[i & 31] = p[i & 31] + v[i];
        part[c] = butterfly32(p);
    }
    return sum_canon_array<T>(part.data(), nc);
}

template <class T, class Get>
T sum_canon(int n, Get get) {
    int nc = (n + 255) / 256;
    if (nc < 1) nc = 1;
    std::vector<T> part(nc);
    for (int c = 0; c < nc; ++c) {
        T p[32];
        for (int l = 0; l < 32; ++l) p[l] = (T)0;
        int hi = std::min(n, (c + 1) * 256);
        for (int i = c * 256; i < hi; ++i) { T v; if (get(i, v)) p[i & 31] = p[i & 31] + v; }
        part[c] = butterfly32(p);
    }
    if (nc == 1) return part[0];
    return sum_canon_array<T>(part.data(), nc);
}

template <class T, class Get>
T sum_by(int order, int n, Get get) {
    return order == OKSS_SUM_CANON256 ? sum_canon<T>(n, get) : sum_serial<T>(n, get);
}

/* ---------------------------------------------------------------- SVD --- */

/* 2x2 plane rotation J = [[c, s], [-s, c]] */
struct Rot { float c, s; };

/* rows p,q of W <- J * rows ; x' = c*x + s*y ; y' = -s*x + c*y */
inline void rot_left(float W[9], int p, int q, Rot j) {
    for (int k = 0; k < 3; ++k) {
        float x = W[3 * p + k], y = W[3 * q + k];
        W[3 * p + k] = j.c * x + j.s * y;
        W[3 * q + k] = -j.s * x + j.c * y;
    }
}
/* cols p,q of W <- cols * J ; x' = c*x - s*y ; y' = s*x + c*y */
inline void rot_right(float W[9], int p, int q, Rot j) {
    for (int k = 0; k < 3; ++k) {
        float x = W[3 * k + p], y = W[3 * k + q];
        W[3 * k + p] = j.c * x - j.s * y;
        W[3 * k + q] = j.s * x + j.c * y;
    }
}

/* Jacobi rotation that diagonalises the symmetric 2x2 [[x, y], [y, z]] */
inline Rot sym_jacobi(float x, float y, float z) {
    Rot r;
    float deno = 2.0f * std::fabs(y);
    if (deno < FLT_MIN) { r.c = 1.0f; r.s = 0.0f; return r; }
    float tau = (x - z) / deno;
    float w = std::sqrt(tau * tau + 1.0f);
    float t = tau > 0.0f ? 1.0f / (tau + w) : 1.0f / (tau - w);
    float sign_t = t > 0.0f ? 1.0f : -1.0f;
    float n = 1.0f / std::sqrt(t * t + 1.0f);
    r.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
    r.c = n;
    return r;
}

/* two-sided Jacobi SVD of a real 3x3 (the algorithm Eigen::JacobiSVD uses for
 * small fixed sizes).  A row-major; A = U diag(s) V^T, s sorted
 * descending, s >= 0. */
void svd3(const float A[9], float U[9], float s[3], float V[9]) {
    float scale = 0.0f;
    for (int i = 0; i < 9; ++i) scale = std::max(scale, std::fabs(A[i]));
    if (!(scale > 0.0f)) scale = 1.0f;
    float W[9];
    for (int i = 0; i < 9; ++i) W[i] = A[i] / scale;
    for (int i = 0; i < 9; ++i) { U[i] = (i % 4 == 0) ? 1.0f : 0.0f; V[i] = U[i]; }
    const float precision = 2.0f * FLT_EPSILON;
    const float tiny = FLT_MIN;
    float maxDiag = std::max(std::fabs(W[0]), std::max(std::fabs(W[4]), std::fabs(W[8])));
    bool finished = false;
    int guard = 0;
    while (!finished && guard++ < 64) {
        finished = true;
        for (int p = 1; p < 3; ++p)
            for (int q = 0; q < p; ++q) {
                float thr = std::max(tiny, precision * maxDiag);
                if (std::fabs(W[3 * p + q]) > thr || std::fabs(W[3 * q + p]) > thr) {
                    finished = false;
                    /* 2x2 block m = [[Wpp, Wpq], [Wqp, Wqq]] : first make it symmetric */
                    float m00 = W[3 * p + p], m01 = W[3 * p + q], m10 = W[3 * q + p], m11 = W[3 * q + q];
                    Rot r1;
                    float t = m00 + m11, d = m10 - m01;
                    if (std::fabs(d) < tiny) { r1.s = 0.0f; r1.c = 1.0f; }
                    else {
                        float u = t / d;
                        float tmp = std::sqrt(1.0f + u * u);
                        r1.s = 1.0f / tmp; r1.c = u / tmp;
                    }
                    /* m <- r1 * m (rows) */
                    float n00 = r1.c * m00 + r1.s * m10;
                    float n01 = r1.c * m01 + r1.s * m11;
                    float n11 = -r1.s * m01 + r1.c * m11;
                    Rot jr = sym_jacobi(n00, n01, n11);
                    /* j_left = r1 * jr^T ; jr^T = (c, -s) */
                    Rot jl;
                    jl.c = r1.c * jr.c - r1.s * (-jr.s);
                    jl.s = r1.c * (-jr.s) + r1.s * jr.c;
                    rot_left(W, p, q, jl);
                    Rot jlt; jlt.c = jl.c; jlt.s = -jl.s;
                    rot_right(U, p, q, jlt);
                    rot_right(W, p, q, jr);
                    rot_right(V, p, q, jr);
                    maxDiag = std::max(maxDiag, std::max(std::fabs(W[3 * p + p]), std::fabs(W[3 * q + q])));
                }
            }
    }
    for (int i = 0; i < 3; ++i) {
        float a = W[4 * i];
        s[i] = std::fabs(a);
        if (a < 0.0f) for (int k = 0; k < 3; ++k) U[3 * k + i] = -U[3 * k + i];
    }
    for (int i = 0; i < 3; ++i) s[i] = s[i] * scale;
    for (int i = 0; i < 3; ++i) {                      /* selection sort, descending */
        int pos = i;
        for (int k = i + 1; k < 3; ++k) if (s[k] > s[pos]) pos = k;
        if (s[pos] == 0.0f) break;
        if (pos != i) {
            std::swap(s[i], s[pos]);
            for (int k = 0; k < 3; ++k) { std::swap(U[3 * k + i], U[3 * k + pos]); std::swap(V[3 * k + i], V[3 * k + pos]); }
        }
    }
}

inline float det3(const float m[9]) {
    /* Eigen's 3x3 determinant: sum of m(0,a) * (m(1,b)*m(2,c) - m(1,c)*m(2,b)) */
    float h0 = m[0] * (m[4] * m[8] - m[5] * m[7]);
    float h1 = m[1] * (m[3] * m[8] - m[5] * m[6]);
    float h2 = m[2] * (m[3] * m[7] - m[4] * m[6]);
    return h0 - h1 + h2;
}

/* pcl::umeyama(src, dst, with_scaling=false) in float.
 * get(k, s[3], d[3]) -> false if slot k carries no correspondence. */
template <class GetPair>
void umeyama(int nslots, int n_corr, int order, GetPair get, float T[16]) {
    const float one_over_n = 1.0f / (float)n_corr;
    float smean[3], dmean[3];
    for (int a = 0; a < 3; ++a) {
        float ss = sum_by<float>(order, nslots, [&](int i, float& v) { float s[3], d[3]; if (!get(i, s, d)) return false; v = s[a]; return true; });
        float ds = sum_by<float>(order, nslots, [&](int i, float& v) { float s[3], d[3]; if (!get(i, s, d)) return false; v = d[a]; return true; });
        smean[a] = ss * one_over_n;
        dmean[a] = ds * one_over_n;
    }
    float sigma[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            float acc = sum_by<float>(order, nslots, [&](int i, float& v) {
                float s[3], d[3]; if (!get(i, s, d)) return false;
                float dd = d[a] - dmean[a];
                float sd = s[b] - smean[b];
                v = dd * sd; return true; });
            sigma[3 * a + b] = one_over_n * acc;
        }
    float U[9], sv[3], V[9];
    svd3(sigma, U, sv, V);
    float S[3] = {1.0f, 1.0f, 1.0f};
    if (det3(sigma) < 0.0f) S[2] = -1.0f;
    int rank = 0;
    for (int i = 0; i < 3; ++i)
        if (!(std::fabs(sv[i]) <= std::fabs(sv[0]) * 1e-5f)) ++rank;   /* !isMuchSmallerThan */
    if (rank == 2) {
        if (det3(U) * det3(V) > 0.0f) { S[0] = S[1] = S[2] = 1.0f; }
        else { S[2] = -1.0f; }
    }
    float R[9];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) {
            float r = (U[3 * a + 0] * S[0]) * V[3 * b + 0];
            r = r + (U[3 * a + 1] * S[1]) * V[3 * b + 1];
            r = r + (U[3 * a + 2] * S[2]) * V[3 * b + 2];
            R[3 * a + b] = r;
        }
    for (int i = 0; i < 16; ++i) T[i] = 0.0f;
    T[15] = 1.0f;
    for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) T[4 * a + b] = R[3 * a + b];
        float rs = R[3 * a + 0] * smean[0];
        rs = rs + R[3 * a + 1] * smean[1];
        rs = rs + R[3 * a + 2] * smean[2];
        T[4 * a + 3] = dmean[a] - rs;
    }
}

/* Matrix4f * (x,y,z,1): column-axpy order ((m0*x + m1*y) + m2*z) + m3 */
inline void xform_point(const float T[16], const float p[3], float o[3]) {
    for (int a = 0; a < 3; ++a) {
        float r = T[4 * a + 0] * p[0];
        r = r + T[4 * a + 1] * p[1];
        r = r + T[4 * a + 2] * p[2];
        r = r + T[4 * a + 3];
        o[a] = r;
    }
}

inline void mat4_mul(const float A[16], const float B[16], float C[16]) {
    float R[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float r = A[4 * i + 0] * B[0 + j];
            r = r + A[4 * i + 1] * B[4 + j];
            r = r + A[4 * i + 2] * B[8 + j];
            r = r + A[4 * i + 3] * B[12 + j];
            R[4 * i + j] = r;
        }
    std::memcpy(C, R, sizeof(R));
}

/* initRegistration_Transfer (initRegistrationKSS.hpp:365-404), one point */
inline void rotate_axis(int cord, double c, double s, double p[3]) {
    double x = p[0], y = p[1], z = p[2];
    if (cord == 1)      { p[0] = x;             p[1] = y * c - z * s; p[2] = y * s + z * c; }
    else if (cord == 2) { p[0] = z * s + x * c; p[1] = y;             p[2] = z * c - x * s; }
    else                { p[0] = x * c - y * s; p[1] = x * s + y * c; p[2] = z; }
}

/* one ICP iteration's estimation step; returns n_corr (or <3 and leaves T untouched) */
int icp_estimate(const float* cur, int n_s, const float* tgt, const NNIndex& nn,
                 double max_dist_sqr, int order, int32_t* idx, float* d2,
                 float T_k[16], double* mse) {
    int cnt = 0;
    for (int i = 0; i < n_s; ++i) {
        int bi; float bd;
        nn.query(cur + 3 * i, bi, bd);
        d2[i] = bd;
        if ((double)bd > max_dist_sqr) { idx[i] = -1; }          /* A.3: skip if d2 > max^2 */
        else { idx[i] = bi; ++cnt; }
    }
    if (cnt < 3) return cnt;
    umeyama(n_s, cnt, order,
            [&](int i, float s[3], float d[3]) {
                if (idx[i] < 0) return false;
                for (int a = 0; a < 3; ++a) { s[a] = cur[3 * i + a]; d[a] = tgt[3 * idx[i] + a]; }
                return true; },
            T_k);
    /* calculateMSE: double sum of the float squared distances / count (A.6) */
    double sum = sum_by<double>(order, n_s, [&](int i, double& v) { if (idx[i] < 0) return false; v = (double)d2[i]; return true; });
    *mse = sum / (double)cnt;
    return cnt;
}

}  // namespace


/* =========================================================== AIVS simplification
 * Restates, with its quirks, the pre-processing that feeds the hot path:
 *   pointPipeline_Border                      pointPipeline.hpp:105-158
 *   BallRegion_EstimateBoxScale / _AchieveXYZ ballRegionCompute.hpp:1194-1214, 690-758
 *   BallRegion_ReturnBoxCenter_Center         ballRegionCompute.hpp:1150-1172   (forgets the y wrap, B10)
 *   BallRegion_BoxInput                       ballRegionCompute.hpp:632-688     (1-based boxes)
 *   BallRegion_ReturnNeiborBox_Box            ballRegionCompute.hpp:975-1040    (forgets the x wrap, B10)
 *   AIVS_initBoxIndexNumber                   Method_AIVS_SimPro.hpp:587-643    (8 colours)
 *   AIVS_BoxSimplification_Points             Method_AIVS_SimPro.hpp:776-794    (quota, +1 if frac > 0.2)
 *   AIVS_Voroni_OpenMP_KNN                    Method_AIVS_SimPro.hpp:222-376    (per-box farthest point sampling)
 *   AIVS_AccurateCut_Optimization             Method_AIVS_SimPro.hpp:848-957    (greedy closest-pair trim, stale lists)
 * kd-tree searches of the reference are exact, so they are replaced by plain scans; their only tie-dependent
 * use is the K=3 list of the trim step, ordered here by (d2, index).  Distances are float d2 on float-narrowed
 * coordinates, sqrt in float (float argument -> float overload), as in the reference. */
namespace {

struct Aivs {
    int n; const double* P;
    int nx, ny, nz; double unit; double mn[3];
    std::vector<std::vector<int>> boxes;
    std::vector<double> creal;       /* [boxes][3] */
    std::vector<int> center;         /* position inside the box list of the point nearest the (quirky) centre */
    int nbox() const { return (int)boxes.size(); }
};

int aivs_box_scale(int n) {
    if (n < 10000) return 10;
    if (n < 50000) return 20;
    if (n < 100000) return 30;
    if (n < 500000) return 40;
    if (n < 1000000) return 50;
    return (int)std::pow((double)n / 8.0, 1.0 / 3.0);
}

void aivs_center(const Aivs& a, int boxIndex, double c[3]) {
    int z_num = boxIndex / (a.nx * a.ny) + 1;
    int leveZ = boxIndex % (a.nx * a.ny);
    int y_num = leveZ / a.nx + 1;
    int x_num = leveZ % a.nx;
    if (x_num == 0) { x_num = a.nx; y_num = y_num - 1; }
    c[0] = (a.mn[0] + (x_num - 1) * a.unit + a.mn[0] + x_num * a.unit) / 2;
    c[1] = (a.mn[1] + (y_num - 1) * a.unit + a.mn[1] + y_num * a.unit) / 2;
    c[2] = (a.mn[2] + (z_num - 1) * a.unit + a.mn[2] + z_num * a.unit) / 2;
}

std::vector<int> aivs_neighbours(const Aivs& a, int boxIndex) {
    int z_num = boxIndex / (a.nx * a.ny) + 1;
    int leveZ = boxIndex % (a.nx * a.ny);
    int y_num = leveZ / a.nx + 1;
    int x_num = leveZ % a.nx;                                     /* no wrap fix here (B10) */
    std::vector<int> xs, ys, zs, out;
    if (x_num > 1) xs.push_back(x_num - 1);
    xs.push_back(x_num);
    if (x_num < a.nx) xs.push_back(x_num + 1);
    if (y_num > 1) ys.push_back(y_num - 1);
    ys.push_back(y_num);
    if (y_num < a.ny) ys.push_back(y_num + 1);
    if (z_num > 1) zs.push_back(z_num - 1);
    zs.push_back(z_num);
    if (z_num < a.nz) zs.push_back(z_num + 1);
    for (int xi : xs) for (int yj : ys) for (int zk : zs) {
        if (xi == x_num && yj == y_num && zk == z_num) continue;
        int idx = xi + (yj - 1) * a.nx + (zk - 1) * a.nx * a.ny;
        if (idx < a.nbox()) out.push_back(idx);
    }
    return out;
}

}  // namespace

extern "C" int okss_aivs_simplify(const double* pts, int n, int pointNum, double* out, int32_t* out_idx) {
    if (n < 1) return 0;
    Aivs a; a.n = n; a.P = pts;
    /* pointPipeline_Border: first index of each strict extreme */
    int imin[3] = {0, 0, 0}, imax[3] = {0, 0, 0};
    double vmin[3] = {pts[0], pts[1], pts[2]}, vmax[3] = {pts[0], pts[1], pts[2]};
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) {
            double v = pts[3 * i + d];
            if (v < vmin[d]) { vmin[d] = v; imin[d] = i; }
            if (v > vmax[d]) { vmax[d] = v; imax[d] = i; }
        }
    const int boxNum = aivs_box_scale(n);
    for (int d = 0; d < 3; ++d) a.mn[d] = pts[3 * imin[d] + d];
    double dis[3];
    for (int d = 0; d < 3; ++d) dis[d] = std::fabs(pts[3 * imax[d] + d] - a.mn[d]);
    double large = dis[0];
    if (large < dis[1]) large = dis[1];
    if (large < dis[2]) large = dis[2];
    a.unit = large / double(boxNum);
    int num[3];
    for (int d = 0; d < 3; ++d) {
        double nd = dis[d] / a.unit;
        num[d] = (int)nd;
        if (nd > (double)num[d]) num[d]++;
    }
    a.nx = num[0]; a.ny = num[1]; a.nz = num[2];
    a.boxes.assign((size_t)a.nx * a.ny * a.nz + 1, std::vector<int>());
    a.creal.resize(3 * a.boxes.size());
    for (int b = 0; b < a.nbox(); ++b) aivs_center(a, b, &a.creal[3 * b]);
    /* BallRegion_BoxInput */
    std::vector<double> cmin(a.boxes.size(), 9999);
    a.center.assign(a.boxes.size(), -1);
    for (int i = 0; i < n; ++i) {
        int id[3];
        for (int d = 0; d < 3; ++d) {
            double v = (pts[3 * i + d] - a.mn[d]) / a.unit;
            id[d] = (int)v;
            if (id[d] < v || id[d] == 0) id[d]++;
        }
        int b = id[0] + a.nx * (id[1] - 1) + (a.nx * a.ny * (id[2] - 1));
        if (b < 0 || b >= a.nbox()) continue;                         /* the reference only prints "Hello!" (UB) */
        const double* c = &a.creal[3 * b];
        double dm = std::sqrt((c[0] - pts[3 * i]) * (c[0] - pts[3 * i]) + (c[1] - pts[3 * i + 1]) * (c[1] - pts[3 * i + 1]) +
                              (c[2] - pts[3 * i + 2]) * (c[2] - pts[3 * i + 2]));
        a.boxes[b].push_back(i);
        if (cmin[b] > dm) { cmin[b] = dm; a.center[b] = (int)a.boxes[b].size() - 1; }
    }
    /* AIVS_initBoxIndexNumber: colours in x-outer, y, z-inner loop order */
    std::vector<int> colour[8];
    for (int i = 1; i <= a.nx; ++i)
        for (int j = 1; j <= a.ny; ++j)
            for (int k = 1; k <= a.nz; ++k) {
                int b = i + a.nx * (j - 1) + (a.nx * a.ny * (k - 1));
                if (a.boxes[b].empty()) continue;
                int c;
                if (i % 2 == 1 && j % 2 == 1 && k % 2 == 1) c = 0;
                else if (i % 2 == 0 && j % 2 == 1 && k % 2 == 1) c = 1;
                else if (i % 2 == 0 && j % 2 == 0 && k % 2 == 1) c = 2;
                else if (i % 2 == 1 && j % 2 == 0 && k % 2 == 1) c = 3;
                else if (i % 2 == 1 && j % 2 == 1 && k % 2 == 0) c = 4;
                else if (i % 2 == 0 && j % 2 == 1 && k % 2 == 0) c = 5;
                else if (i % 2 == 0 && j % 2 == 0 && k % 2 == 0) c = 6;
                else c = 7;
                colour[c].push_back(b);
            }
    /* quotas */
    const double rate = (double)pointNum / (double)n;
    std::vector<int> quota(a.boxes.size(), 0);
    for (int b = 0; b < a.nbox(); ++b) {
        double sb = (double)a.boxes[b].size() * rate;
        int t = (int)sb;
        quota[b] = (sb - t > 0.2) ? t + 1 : t;
    }
    std::vector<int> labelG(n, 1);
    std::vector<std::vector<int>> simiT(a.boxes.size());
    for (int b = 0; b < a.nbox(); ++b) simiT[b].assign(a.boxes[b].size(), -1);
    const double radius = a.unit * 3.0 / 4.0;
    std::vector<float> pf;
    for (int c = 0; c < 8; ++c)
        for (int b : colour[c]) {
            const int simNum = quota[b];
            if (simNum == 0) continue;
            const double* pc = &a.creal[3 * b];
            std::vector<int> pointTemp = a.boxes[b];
            std::vector<int> labelTemp(pointTemp.size(), 1);
            bool addJ = true;
            for (int nb : aivs_neighbours(a, b))
                for (int pt : a.boxes[nb]) {
                    const double* q = pts + 3 * pt;
                    if (q[0] <= pc[0] + radius && q[0] >= pc[0] - radius && q[1] <= pc[1] + radius && q[1] >= pc[1] - radius &&
                        q[2] <= pc[2] + radius && q[2] >= pc[2] - radius && labelG[pt] == 0) {
                        pointTemp.push_back(pt); labelTemp.push_back(2); addJ = false;
                    }
                }
            if (addJ && a.center[b] >= 0 && a.center[b] < (int)pointTemp.size()) labelTemp[a.center[b]] = 0;
            const int m = (int)pointTemp.size();
            pf.resize(3 * (size_t)m);
            for (int k = 0; k < m; ++k) for (int d = 0; d < 3; ++d) pf[3 * k + d] = (float)pts[3 * pointTemp[k] + d];
            int sampled = 0;
            std::vector<double> mind(m, -1);
            for (int k = 0; k < m; ++k) {
                if (labelTemp[k] == 0) {
                    mind[k] = 0; simiT[b][sampled] = pointTemp[k]; labelG[pointTemp[k]] = 0; sampled++;
                } else if (labelTemp[k] == 2) {
                    mind[k] = 0;
                } else {
                    double minTemp = 9999;
                    float best = std::numeric_limits<float>::infinity();
                    for (int l = 0; l < m; ++l)
                        if (labelTemp[l] == 0 || labelTemp[l] == 2) best = std::min(best, dist2f(&pf[3 * k], &pf[3 * l]));
                    if (best != std::numeric_limits<float>::infinity()) minTemp = (double)std::sqrt(best);
                    mind[k] = minTemp;
                }
            }
            while (sampled < simNum) {
                int sel = -1; double mx = 0;
                for (int k = 0; k < m; ++k) if (labelTemp[k] == 1 && mind[k] > mx) { sel = k; mx = mind[k]; }
                if (sel == -1) break;
                mind[sel] = 0; labelG[pointTemp[sel]] = 0; simiT[b][sampled] = pointTemp[sel]; sampled++;
                for (int k = 0; k < m; ++k)
                    if (labelTemp[k] == 1) {
                        double d = (double)std::sqrt(dist2f(&pf[3 * sel], &pf[3 * k]));
                        if (d < mind[k]) mind[k] = d;
                    }
            }
        }
    /* AIVS_AccurateCut_Optimization */
    std::vector<int> sample;
    for (int b = 0; b < a.nbox(); ++b)
        for (int v : simiT[b]) { if (v == -1) break; sample.push_back(v); }
    const int S = (int)sample.size();
    int dTiff = S - pointNum;
    std::vector<char> keep(S, 1);
    if (dTiff > 0 && S >= 3) {
        std::vector<float> sf(3 * (size_t)S);
        for (int i = 0; i < S; ++i) for (int d = 0; d < 3; ++d) sf[3 * i + d] = (float)pts[3 * sample[i] + d];
        std::vector<int> nidx(3 * (size_t)S);
        std::vector<float> ndis(3 * (size_t)S);
        for (int i = 0; i < S; ++i) {                                   /* K = 3, ascending (d2, index) */
            float bd[3] = {INFINITY, INFINITY, INFINITY}; int bi[3] = {-1, -1, -1};
            for (int j = 0; j < S; ++j) {
                float d = dist2f(&sf[3 * i], &sf[3 * j]);
                int jj = j;
                for (int r = 0; r < 3; ++r)
                    if (better(d, jj, bd[r], bi[r] < 0 ? INT32_MAX : bi[r])) { std::swap(d, bd[r]); std::swap(jj, bi[r]); }
            }
            for (int r = 0; r < 3; ++r) { nidx[3 * i + r] = bi[r]; ndis[3 * i + r] = std::sqrt(bd[r]); }
        }
        while (dTiff > 0) {
            double mn = 9999; int b1 = -1, b2 = -1;
            for (int i = 0; i < S; ++i) {
                int b2t = nidx[3 * i + 1];
                double dt = ndis[3 * i + 1];
                if (dt < mn && keep[i] && keep[b2t]) { mn = dt; b1 = i; b2 = b2t; }
            }
            if (mn == 9999 || b1 == -1 || b2 == -1) break;
            int del = b1;
            if ((double)ndis[3 * b1 + 2] > (double)ndis[3 * b2 + 2]) del = b2;
            keep[del] = 0; dTiff--;
        }
    }
    int m = 0;
    for (int i = 0; i < S; ++i)
        if (keep[i]) {
            if (out) for (int d = 0; d < 3; ++d) out[3 * m + d] = pts[3 * sample[i] + d];
            if (out_idx) out_idx[m] = sample[i];
            ++m;
        }
    return m;
}

/* ===================================================================== API */

extern "C" {

int okss_max_threads(void) {
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

float okss_canon_sum_f32(const float* v, int n) {
    return sum_canon<float>(n, [&](int i, float& o) { o = v[i]; return true; });
}
double okss_canon_sum_f64(const double* v, int n) {
    return sum_canon<double>(n, [&](int i, double& o) { o = v[i]; return true; });
}

void okss_nn(const float* q, int nq, const float* t, int nt, int method, int32_t* idx, float* d2) {
    NNIndex nn(t, nt, method);
    for (int i = 0; i < nq; ++i) { int bi; float bd; nn.query(q + 3 * i, bi, bd); idx[i] = bi; d2[i] = bd; }
}

void okss_svd3(const float A[9], float U[9], float s[3], float V[9]) { svd3(A, U, s, V); }

void okss_umeyama(const float* src, const float* dst, int n, int sum_order, float T[16]) {
    umeyama(n, n, sum_order,
            [&](int i, float s[3], float d[3]) { for (int a = 0; a < 3; ++a) { s[a] = src[3 * i + a]; d[a] = dst[3 * i + a]; } return true; },
            T);
}

/* initRegistrationKSS.hpp:144-220 ; every sum is a serial index-order double loop */
void okss_middle_align(const double* src, int n_s, const double* tgt, int n_t,
                       double out7[7], double* src_aligned) {
    double xs = 0, ys = 0, zs = 0, avg = 0;
    for (int i = 0; i < n_s; ++i) { xs = xs + src[3 * i]; ys = ys + src[3 * i + 1]; zs = zs + src[3 * i + 2]; }
    xs = xs / n_s; ys = ys / n_s; zs = zs / n_s;
    for (int i = 0; i < n_s; ++i) {
        double xl = src[3 * i] - xs, yl = src[3 * i + 1] - ys, zl = src[3 * i + 2] - zs;
        double len = std::sqrt(xl * xl + yl * yl + zl * zl);
        avg = avg + len;
    }
    avg = avg / n_s;
    double xt = 0, yt = 0, zt = 0, avg2 = 0;
    for (int i = 0; i < n_t; ++i) { xt = xt + tgt[3 * i]; yt = yt + tgt[3 * i + 1]; zt = zt + tgt[3 * i + 2]; }
    xt = xt / n_t; yt = yt / n_t; zt = zt / n_t;
    for (int i = 0; i < n_t; ++i) {
        double xl = tgt[3 * i] - xt, yl = tgt[3 * i + 1] - yt, zl = tgt[3 * i + 2] - zt;
        double len = std::sqrt(xl * xl + yl * yl + zl * zl);
        avg2 = avg2 + len;
    }
    avg2 = avg2 / n_t;
    out7[0] = xt; out7[1] = yt; out7[2] = zt;
    out7[3] = xt - xs; out7[4] = yt - ys; out7[5] = zt - zs;
    out7[6] = avg2 / avg;
    if (src_aligned) {
        for (int i = 0; i < n_s; ++i)
            for (int a = 0; a < 3; ++a) {
                double v = src[3 * i + a] + out7[3 + a];
                src_aligned[3 * i + a] = out7[a] + (v - out7[a]) * out7[6];
            }
    }
}

/* initRegistrationKSS.hpp:245 : for (double a = 0; a < 6.3; a = a + 6.3 / step) */
int okss_sweep_angles(double step, double* accum, double* list, int cap) {
    int g = 0;
    for (double a = 0; a < 6.3; a = a + 6.3 / step) {
        if (g < cap) {
            if (accum) accum[g] = a;
            if (list) list[g] = (double)g * 6.3 / (double)step;   /* :282-284 */
        }
        ++g;
    }
    return g;
}

/* KSS_SCORE_VOXEL (a reading of the authors' closed CUDA build, not of the released sources): the target as an NV^3
 * occupancy grid over the cube +-1.1 max|coordinate|; a hypothesis scores the share of source points in empty voxels */
struct VoxelGrid {
    int NV;                                     /* voxels per axis: ~16 target points per occupied face voxel, 4..32 */
    std::vector<unsigned char> occ;
    float E, inv;
    VoxelGrid(const float* t, int n_t) {
        NV = (int)std::floor(std::sqrt((double)n_t / 16.0));
        NV = std::min(32, std::max(4, NV));
        occ.assign((size_t)NV * NV * NV, 0);
        float m = 0.0f;
        for (int i = 0; i < 3 * n_t; ++i) m = std::max(m, std::fabs(t[i]));
        E = 1.1f * m;
        inv = (float)NV / (2.0f * E);
        if (E > 0.0f)
            for (int i = 0; i < n_t; ++i) occ[((size_t)vox(t[3 * i + 2]) * NV + vox(t[3 * i + 1])) * NV + vox(t[3 * i])] = 1;
    }
    int vox(float v) const { float f = v + E; f = f * inv; return (int)std::floor(f); }
    bool hit(const float q[3]) const {
        if (!(E > 0.0f)) return false;
        const int ix = vox(q[0]), iy = vox(q[1]), iz = vox(q[2]);
        if (ix < 0 || ix >= NV || iy < 0 || iy >= NV || iz < 0 || iz >= NV) return false;
        return occ[((size_t)iz * NV + iy) * NV + ix] != 0;
    }
};
static double score_voxel(const std::vector<double>& pts, int n_s, const VoxelGrid& vg) {
    double sum = 0;
    for (int i = 0; i < n_s; ++i) {
        float q[3] = {(float)pts[3 * i], (float)pts[3 * i + 1], (float)pts[3 * i + 2]};
        sum = sum + (vg.hit(q) ? 0.0 : 1.0);
    }
    return sum / n_s;
}

static double score_cloud(const std::vector<double>& pts, int n_s, const NNIndex& nn, int score_mode,
                          float* queries, int32_t* idx, float* d2out) {
    /* initRegistrationKSS.hpp:406-479 : K=2 search, neighbour [0] only */
    double sum = 0, dmax = -9999;
    for (int i = 0; i < n_s; ++i) {
        float q[3] = {(float)pts[3 * i], (float)pts[3 * i + 1], (float)pts[3 * i + 2]};
        int bi; float bd;
        nn.query(q, bi, bd);
        if (queries) { queries[3 * i] = q[0]; queries[3 * i + 1] = q[1]; queries[3 * i + 2] = q[2]; }
        if (idx) idx[i] = bi;
        if (d2out) d2out[i] = bd;
        if (score_mode == OKSS_SCORE_MAX) {
            double di = bd;
            if (di > dmax) dmax = di;
        } else {
            /* `sqrt(pointNKNSquaredDistance[0])` has a float argument: with <math.h> + `using namespace std`
             * overload resolution picks float sqrt(float) (MSVC and GCC alike); the result is widened afterwards
             * (initRegistrationKSS.hpp:444, 468) */
            double di = (double)std::sqrt(bd);
            sum = sum + di;
            if (dmax < di) dmax = di;
        }
    }
    if (score_mode == OKSS_SCORE_MAX) return dmax;
    if (score_mode == OKSS_SCORE_DIFF) return dmax - sum / n_s;
    return sum / n_s;
}

double okss_sweep_score(const double* src_aligned, int n_s, const double* tgt, int n_t,
                        const double angles[3], int score_mode, int nn_method,
                        float* queries, int32_t* idx, float* d2) {
    std::vector<float> tf(3 * (size_t)n_t);
    for (int i = 0; i < 3 * n_t; ++i) tf[i] = (float)tgt[i];
    NNIndex nn(tf.data(), n_t, nn_method);
    std::vector<double> p(src_aligned, src_aligned + 3 * (size_t)n_s);
    for (int ax = 0; ax < 3; ++ax) {
        double c = std::cos(angles[ax]), s = std::sin(angles[ax]);
        for (int i = 0; i < n_s; ++i) rotate_axis(ax + 1, c, s, &p[3 * i]);
    }
    return score_cloud(p, n_s, nn, score_mode, queries, idx, d2);
}

int okss_sweep(const double* src_aligned, int n_s, const double* tgt, int n_t,
               double step, int score_mode, int nn_method,
               double* value, double best_angle[3], int best_index[3],
               int* minima_idx, int* n_minima) {
    std::vector<float> tf(3 * (size_t)n_t);
    for (int i = 0; i < 3 * n_t; ++i) tf[i] = (float)tgt[i];          /* :231-235 narrowing */
    NNIndex nn(tf.data(), n_t, nn_method);
    VoxelGrid vg(tf.data(), score_mode == OKSS_SCORE_VOXEL ? n_t : 0);
    std::vector<double> ang(1024);
    int G = okss_sweep_angles(step, ang.data(), nullptr, 1024);
    double errorT = 9999;
    best_angle[0] = best_angle[1] = best_angle[2] = 0;
    best_index[0] = best_index[1] = best_index[2] = 0;
    std::vector<double> px(3 * (size_t)n_s), pxy(3 * (size_t)n_s), pxyz(3 * (size_t)n_s);
    for (int i = 0; i < G; ++i) {
        double ci = std::cos(ang[i]), si = std::sin(ang[i]);
        px.assign(src_aligned, src_aligned + 3 * (size_t)n_s);
        for (int k = 0; k < n_s; ++k) rotate_axis(1, ci, si, &px[3 * k]);
        for (int j = 0; j < G; ++j) {
            double cj = std::cos(ang[j]), sj = std::sin(ang[j]);
            pxy = px;
            for (int k = 0; k < n_s; ++k) rotate_axis(2, cj, sj, &pxy[3 * k]);
            for (int kk = 0; kk < G; ++kk) {
                double ck = std::cos(ang[kk]), sk = std::sin(ang[kk]);
                pxyz = pxy;
                for (int k = 0; k < n_s; ++k) rotate_axis(3, ck, sk, &pxyz[3 * k]);
                double e = score_mode == OKSS_SCORE_VOXEL ? score_voxel(pxyz, n_s, vg)
                                                          : score_cloud(pxyz, n_s, nn, score_mode, nullptr, nullptr, nullptr);
                value[((size_t)i * G + j) * G + kk] = e;
                if (e < errorT) {                                       /* :258 first strict min */
                    errorT = e;
                    best_angle[0] = ang[i]; best_angle[1] = ang[j]; best_angle[2] = ang[kk];
                    best_index[0] = i; best_index[1] = j; best_index[2] = kk;
                }
            }
        }
    }
    /* initRegistration_kernel (:481-522): clamped, non-periodic +-2 window, plateaus count */
    const int r = 2;
    int nm = 0;
    for (int i = 0; i < G; ++i)
        for (int j = 0; j < G; ++j)
            for (int k = 0; k < G; ++k) {
                double c = value[((size_t)i * G + j) * G + k];
                bool ok = true;
                for (int ii = std::max(0, i - r); ii <= std::min(G - 1, i + r) && ok; ++ii)
                    for (int jj = std::max(0, j - r); jj <= std::min(G - 1, j + r) && ok; ++jj)
                        for (int kk = std::max(0, k - r); kk <= std::min(G - 1, k + r); ++kk)
                            if (c > value[((size_t)ii * G + jj) * G + kk]) { ok = false; break; }
                if (ok) {
                    if (minima_idx) { minima_idx[3 * nm] = i; minima_idx[3 * nm + 1] = j; minima_idx[3 * nm + 2] = k; }
                    ++nm;
                }
            }
    if (n_minima) *n_minima = nm;
    return G;
}

/* initRegistrationKSS.hpp:75-109 */
void okss_apply_similarity(const double* pts, int n, const double al[7], const double angles[3], double* out) {
    double c[3], s[3];
    for (int a = 0; a < 3; ++a) { c[a] = std::cos(angles[a]); s[a] = std::sin(angles[a]); }
    for (int i = 0; i < n; ++i) {
        double p[3];
        for (int a = 0; a < 3; ++a) {
            double v = pts[3 * i + a] + al[3 + a];
            p[a] = al[a] + (v - al[a]) * al[6];
        }
        rotate_axis(1, c[0], s[0], p);
        rotate_axis(2, c[1], s[1], p);
        rotate_axis(3, c[2], s[2], p);
        out[3 * i] = p[0]; out[3 * i + 1] = p[1]; out[3 * i + 2] = p[2];
    }
}

int okss_icp_iteration(const float* src, int n_s, const float* tgt, int n_t,
                       double max_corr_dist, int sum_order, int nn_method,
                       int32_t* idx, float* d2, float T_k[16], double* mse, float* src_out) {
    NNIndex nn(tgt, n_t, nn_method);
    std::vector<int32_t> li; std::vector<float> ld;
    if (!idx) { li.resize(n_s); idx = li.data(); }
    if (!d2) { ld.resize(n_s); d2 = ld.data(); }
    double m = 0;
    for (int i = 0; i < 16; ++i) T_k[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    int cnt = icp_estimate(src, n_s, tgt, nn, max_corr_dist * max_corr_dist, sum_order, idx, d2, T_k, &m);
    if (mse) *mse = m;
    if (src_out && cnt >= 3)
        for (int i = 0; i < n_s; ++i) xform_point(T_k, src + 3 * i, src_out + 3 * i);
    return cnt;
}

/* PCL 1.8.1 IterativeClosestPoint::computeTransformation + getFitnessScore
 *  as wrapped by KSS_ICP.hpp:323-356 */
double okss_icp(const double* src, int n_s, const double* tgt, int n_t,
                const okss_icp_params* p, float T[16], int* iters_out, int* converged_out,
                okss_icp_trace* tr) {
    std::vector<float> input(3 * (size_t)n_s), target(3 * (size_t)n_t);
    for (int i = 0; i < 3 * n_s; ++i) input[i] = (float)src[i];      /* KSS_ICP.hpp:328-333 */
    for (int i = 0; i < 3 * n_t; ++i) target[i] = (float)tgt[i];
    NNIndex nn(target.data(), n_t, p->nn_method);
    std::vector<float> cur(input);
    std::vector<int32_t> idx(n_s);
    std::vector<float> d2(n_s);
    float fin[16];
    for (int i = 0; i < 16; ++i) fin[i] = (i % 5 == 0) ? 1.0f : 0.0f;
    const double max_dist_sqr = p->max_corr_dist * p->max_corr_dist;
    const double rot_thr = 1.0 - p->transformation_eps;              /* A.6 */
    const double trans_thr = p->transformation_eps;
    const double mse_rel = p->fitness_eps, mse_abs = 1e-12;
    double prev_mse = std::numeric_limits<double>::max();
    int iters = 0; bool converged = false;
    do {
        float Tk[16]; double mse = 0;
        if (tr && tr->src_k && iters < tr->cap_iters)
            std::memcpy(tr->src_k + (size_t)iters * 3 * n_s, cur.data(), sizeof(float) * 3 * n_s);
        int cnt = icp_estimate(cur.data(), n_s, target.data(), nn, max_dist_sqr, p->sum_order,
                               idx.data(), d2.data(), Tk, &mse);
        if (cnt < 3) { converged = false; break; }                   /* min_number_correspondences_ */
        if (tr && iters < tr->cap_iters) {
            if (tr->corr_idx) std::memcpy(tr->corr_idx + (size_t)iters * n_s, idx.data(), sizeof(int32_t) * n_s);
            if (tr->T_k) std::memcpy(tr->T_k + (size_t)iters * 16, Tk, sizeof(Tk));
            if (tr->mse) tr->mse[iters] = mse;
        }
        for (int i = 0; i < n_s; ++i) {                              /* transformCloud, in place */
            float o[3]; xform_point(Tk, &cur[3 * i], o);
            cur[3 * i] = o[0]; cur[3 * i + 1] = o[1]; cur[3 * i + 2] = o[2];
        }
        mat4_mul(Tk, fin, fin);                                      /* final = T_k * final */
        ++iters;
        /* DefaultConvergenceCriteria<float>::hasConverged */
        if (iters >= p->max_iterations) { converged = true; break; }
        double cos_angle = 0.5 * (Tk[0] + Tk[5] + Tk[10] - 1);       /* float sum, then double */
        double translation_sqr = Tk[3] * Tk[3] + Tk[7] * Tk[7] + Tk[11] * Tk[11];
        if (cos_angle >= rot_thr && translation_sqr <= trans_thr) { converged = true; break; }
        if (std::fabs(mse - prev_mse) < mse_abs) { converged = true; break; }
        if (std::fabs(mse - prev_mse) / prev_mse < mse_rel) { converged = true; break; }
        prev_mse = mse;
    } while (!converged);
    std::memcpy(T, fin, sizeof(fin));
    if (iters_out) *iters_out = iters;
    if (converged_out) *converged_out = converged ? 1 : 0;
    /* getFitnessScore(): final * ORIGINAL input, one rounding (A.7) */
    double fit = 0; int nr = 0;
    double fsum = sum_by<double>(p->sum_order, n_s, [&](int i, double& v) {
        float o[3]; xform_point(fin, &input[3 * i], o);
        int bi; float bd; nn.query(o, bi, bd);
        v = (double)bd; return true; });
    nr = n_s; fit = fsum;
    if (nr > 0) return fit / nr;
    return std::numeric_limits<double>::max();
}

/* KSS_ICP.hpp:222-230 : rt(i,j) float promoted to double, left-to-right sums */
void okss_apply_transform(const float T[16], const double* pts, int n, double* out) {
    for (int i = 0; i < n; ++i) {
        double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
        for (int a = 0; a < 3; ++a) {
            double r = (double)T[4 * a + 0] * x;
            r = r + (double)T[4 * a + 1] * y;
            r = r + (double)T[4 * a + 2] * z;
            r = r + (double)T[4 * a + 3];
            out[3 * i + a] = r;
        }
    }
}

/* registrationMeasure.hpp:47-98 */
void okss_nn_metrics(const double* a, int n_a, const double* t, int n_t, int nn_method, double out3[3]) {
    std::vector<float> tf(3 * (size_t)n_t);
    for (int i = 0; i < 3 * n_t; ++i) tf[i] = (float)t[i];
    NNIndex nn(tf.data(), n_t, nn_method);
    double mse = 0, mae = 0;
    for (int i = 0; i < n_a; ++i) {
        float q[3] = {(float)a[3 * i], (float)a[3 * i + 1], (float)a[3 * i + 2]};
        int bi; float bd; nn.query(q, bi, bd);
        double m = bd;
        mse = mse + m;
        mae = mae + std::sqrt(m);
    }
    mse = mse / (double)n_a;
    mae = mae / (double)n_a;
    out3[0] = mse; out3[1] = std::sqrt(mse); out3[2] = mae;
}

/* KSSICP_Registration after the AIVS simplification (KSS_ICP.hpp:86-130) + PCR_QM */
void okss_register(const double* sim_s, int n_s, const double* sim_t, int n_t,
                   const double* full_s, int N_s, const double* full_t, int N_t,
                   double step, int max_iter, int sum_order, int nn_method,
                   okss_pair_result* res, double* point_align) {
    std::memset(res, 0, sizeof(*res));
    okss_icp_params ip{max_iter, 1.0, 1e-10, 0.001, sum_order, nn_method};
    std::vector<double> aligned(3 * (size_t)n_s);
    okss_middle_align(sim_s, n_s, sim_t, n_t, res->align, aligned.data());
    std::vector<double> accum(1024), list(1024);
    int G = okss_sweep_angles(step, accum.data(), list.data(), 1024);
    std::vector<double> value((size_t)G * G * G);
    std::vector<int> minima(3 * (size_t)G * G * G);
    int nmin = 0;
    okss_sweep(aligned.data(), n_s, sim_t, n_t, step, OKSS_SCORE_AVE, nn_method,
               value.data(), res->best_angle, res->best_index, minima.data(), &nmin);
    res->G = G; res->n_minima = nmin;
    std::vector<double> sss(3 * (size_t)n_s);
    float T[16]; int it = 0, cv = 0;
    okss_apply_similarity(sim_s, n_s, res->align, res->best_angle, sss.data());   /* :92 */
    double E = okss_icp(sss.data(), n_s, sim_t, n_t, &ip, T, &it, &cv, nullptr); /* :93 judge */
    res->judge_fitness = E; res->judge_iters = it; res->total_icp_iters = it; res->n_icp_runs = 1;
    double used[3] = {res->best_angle[0], res->best_angle[1], res->best_angle[2]};
    res->winner = -1;
    if (E > 0.0005) {                                                             /* :99 */
        res->branch_multi = 1;
        double Q = 9999; int angleIndex = 0;
        for (int l = 0; l < nmin; ++l) {
            double ang[3] = {list[minima[3 * l]], list[minima[3 * l + 1]], list[minima[3 * l + 2]]};
            okss_apply_similarity(sim_s, n_s, res->align, ang, sss.data());       /* :103 */
            double ri = okss_icp(sss.data(), n_s, sim_t, n_t, &ip, T, &it, &cv, nullptr);
            res->total_icp_iters += it; res->n_icp_runs += 1;
            if (ri < Q && ri >= 0) { Q = ri; angleIndex = l; }                    /* :113 */
        }
        res->winner = angleIndex;
        for (int a = 0; a < 3; ++a) used[a] = list[minima[3 * angleIndex + a]];
    }
    for (int a = 0; a < 3; ++a) res->used_angle[a] = used[a];
    okss_apply_similarity(sim_s, n_s, res->align, used, sss.data());              /* :119 / :123 */
    std::vector<double> fullA(3 * (size_t)N_s);
    okss_apply_similarity(full_s, N_s, res->align, used, fullA.data());           /* :120 / :124 */
    res->final_fitness = okss_icp(sss.data(), n_s, sim_t, n_t, &ip, res->T, &it, &cv, nullptr); /* :130 */
    res->final_iters = it; res->total_icp_iters += it; res->n_icp_runs += 1;
    std::vector<double> pa(3 * (size_t)N_s);
    okss_apply_transform(res->T, fullA.data(), N_s, pa.data());                   /* :224-230 */
    double m3[3];
    okss_nn_metrics(pa.data(), N_s, full_t, N_t, nn_method, m3);
    res->mse = m3[0]; res->rmse = m3[1]; res->mae = m3[2];
    if (point_align) std::memcpy(point_align, pa.data(), sizeof(double) * 3 * (size_t)N_s);
}

int okss_register_batch(int n_pairs,
                        const double* sim_s, int n_s, const double* sim_t, int n_t,
                        const double* full_s, int N_s, const double* full_t, int N_t,
                        double step, int max_iter, int sum_order, int nn_method,
                        int threads, okss_pair_result* res) {
    int used = threads > 0 ? threads : okss_max_threads();
    if (used > n_pairs) used = n_pairs > 0 ? n_pairs : 1;
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            int p = next.fetch_add(1);
            if (p >= n_pairs) break;
            const double* fs = full_s + (size_t)p * 3 * N_s;
            const double* ft = full_t + (size_t)p * 3 * N_t;
            if (sim_s && sim_t) {
                okss_register(sim_s + (size_t)p * 3 * n_s, n_s, sim_t + (size_t)p * 3 * n_t, n_t, fs, N_s, ft, N_t,
                              step, max_iter, sum_order, nn_method, res + p, nullptr);
            } else {
                /* the whole of KSSICP_init + KSSICP_Registration (KSS_ICP.hpp:53-130): pNumber rule, then AIVS of
                 * the target and of the source */
                int pn = (N_s > N_t ? N_t : N_s) / 2;
                if (pn > 2000) pn = 2000;
                std::vector<double> st((size_t)3 * N_t), ss((size_t)3 * N_s);
                const int mt = okss_aivs_simplify(ft, N_t, pn, st.data(), nullptr);
                const int ms = okss_aivs_simplify(fs, N_s, pn, ss.data(), nullptr);
                okss_register(ss.data(), ms, st.data(), mt, fs, N_s, ft, N_t, step, max_iter, sum_order, nn_method,
                              res + p, nullptr);
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < used; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
    return used;
}

}  // extern "C"
