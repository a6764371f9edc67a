// kss_api.cu -- the C ABI (include/kss_icp_b200.h): context, device memory, host<->device
// copies and the kernel pipelines.  No computation of the path happens on the host; the only
// host arithmetic is the angle grid and its libm cos/sin tables, which the reference also
// evaluates on the CPU (initRegistrationKSS.hpp:245, 371-397).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>            // types only: the library is resolved at run time (kss_nccl below), never linked

#include "kss_kernels.h"
#include "kss_large.h"
#include "kss_aivs.h"

using namespace kss;

struct kss_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    long long launches = 0;
    // angle grid cache
    double step = -1.0;
    int G = 0;
    std::vector<double> accum, list;
    double* d_trig_accum = nullptr;   // [G][2] cos, sin of the accumulated loop values
    double* d_trig_list = nullptr;    // [G][2] cos, sin of index*6.3/step
    // grow-only named device buffers
    struct Buf { void* p = nullptr; size_t cap = 0; };
    std::map<std::string, Buf> bufs;
    size_t ws_budget = (size_t)48 << 30;
    // hypothesis sharding (kss_ctx_nccl_init): this ctx is rank `rank` of `world`, one ctx per GPU
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    int hyp_slots = 32;               // hypothesis CTAs per pair in the batched ICP launch (kss_ctx_set_hyp_slots)
    int* aivs_bad = nullptr;          // device flag written by the last raw-cloud batch (kss_aivs.h)
    // batch lanes: chunks of a batch alternate between a few internal streams, so that one chunk's copies and the
    // thin tail of its ICP launch overlap the next chunk's kernels; every lane has its own set of named buffers
    static constexpr int MAX_LANES = 4;
    cudaStream_t lane_stream[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t lane_done[MAX_LANES] = {nullptr, nullptr, nullptr, nullptr};
    // full-resolution PCR_QM of a batch of large pairs: the pairs of a chunk go round-robin over a few side streams (each
    // with its own large-path buffers), so that the many small launches of one pair overlap those of its neighbours
    static constexpr int MQ = 4;
    cudaStream_t mq_stream[MQ] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t mq_done[MQ] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t mq_fork = nullptr;
    cudaEvent_t fork_ev = nullptr;
    std::string buf_suffix;           // appended to buffer names while a lane is active
    // optional per-stage CUDA-event timing (bench.py roofline): events on the launching stream
    bool timing = false;
    struct Span { int stage; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> ev_pool;
    LargeIcp large_run;
    bool large_ready = false;
    cudaEvent_t open_ev[KSS_STAGE_COUNT] = {nullptr};
    double stage_ms[KSS_STAGE_COUNT] = {0};
    long long stage_calls[KSS_STAGE_COUNT] = {0};
};

namespace {

// NCCL entry points, resolved from the copy already loaded in the process (torch's) or from libnccl.so.2
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi* kss_nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.ok ? &api : nullptr;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return nullptr;
#define KSS_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name))
    KSS_SYM(GetUniqueId, "ncclGetUniqueId"); KSS_SYM(CommInitRank, "ncclCommInitRank"); KSS_SYM(CommInitAll, "ncclCommInitAll");
    KSS_SYM(CommDestroy, "ncclCommDestroy"); KSS_SYM(AllGather, "ncclAllGather"); KSS_SYM(AllReduce, "ncclAllReduce");
    KSS_SYM(GroupStart, "ncclGroupStart"); KSS_SYM(GroupEnd, "ncclGroupEnd"); KSS_SYM(GetErrorString, "ncclGetErrorString");
#undef KSS_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllGather && api.AllReduce &&
             api.GroupStart && api.GroupEnd && api.GetErrorString;
    return api.ok ? &api : nullptr;
}

int fail(kss_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            char b_[512];                                                                     \
            snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                     __FILE__, __LINE__);                                                     \
            return fail(ctx, KSS_ERR_CUDA, b_);                                               \
        }                                                                                     \
    } while (0)

#define NC(call)                                                                              \
    do {                                                                                      \
        ncclResult_t e_ = (call);                                                             \
        if (e_ != ncclSuccess) {                                                              \
            char b_[512];                                                                     \
            snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #call, kss_nccl()->GetErrorString(e_), \
                     __FILE__, __LINE__);                                                     \
            return fail(ctx, KSS_ERR_NCCL, b_);                                               \
        }                                                                                     \
    } while (0)

#define KL(call)        \
    do {                \
        CU(call);       \
        ctx->launches++;\
    } while (0)

cudaEvent_t ev_get(kss_ctx* c) {
    cudaEvent_t e;
    if (!c->ev_pool.empty()) { e = c->ev_pool.back(); c->ev_pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}
struct StageTimer {
    kss_ctx* c; int stage; cudaEvent_t a = nullptr;
    StageTimer(kss_ctx* ctx, int st) : c(ctx), stage(st) {
        if (c->timing) { a = ev_get(c); cudaEventRecord(a, c->stream); }
    }
    ~StageTimer() {
        if (a) { cudaEvent_t b = ev_get(c); cudaEventRecord(b, c->stream); c->spans.push_back({stage, a, b}); }
    }
};
void collect_spans(kss_ctx* c) {
    for (auto& sp : c->spans) {
        float ms = 0.f;
        if (cudaEventSynchronize(sp.b) == cudaSuccess && cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
            c->stage_ms[sp.stage] += ms; c->stage_calls[sp.stage] += 1;
        }
        c->ev_pool.push_back(sp.a); c->ev_pool.push_back(sp.b);
    }
    c->spans.clear();
}

template <class T>
int dev_buf(kss_ctx* ctx, const char* name, size_t count, T** out) {
    kss_ctx::Buf& b = ctx->bufs[ctx->buf_suffix.empty() ? std::string(name) : std::string(name) + ctx->buf_suffix];
    size_t bytes = count * sizeof(T);
    if (bytes == 0) bytes = 16;
    if (b.cap < bytes) {
        if (b.p) { cudaDeviceSynchronize(); cudaFree(b.p); b.p = nullptr; b.cap = 0; }
        size_t want = bytes + bytes / 8;
        cudaError_t e = cudaMalloc(&b.p, want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&b.p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            char m[256];
            snprintf(m, sizeof(m), "cudaMalloc(%zu bytes) for '%s' failed: %s", bytes, name, cudaGetErrorString(e));
            return fail(ctx, KSS_ERR_NOMEM, m);
        }
        b.cap = want;
    }
    *out = reinterpret_cast<T*>(b.p);
    return KSS_OK;
}
#define BUF(name, count, ptr)                              \
    do {                                                   \
        int r_ = dev_buf(ctx, name, (size_t)(count), ptr); \
        if (r_ != KSS_OK) return r_;                       \
    } while (0)

__global__ void fill_f64_kernel(double* p, size_t n, double v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
// hypothesis slabs of the score grid: value [P][hpad] <-> vg [world][P][slab] (equal counts for ncclAllGather)
__global__ void slab_pack_kernel(const double* value, int P, int hpad, int h_lo, int h_hi, int slab, double* vg_mine) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (i < slab) vg_mine[(size_t)p * slab + i] = h_lo + i < h_hi ? value[(size_t)p * hpad + h_lo + i] : 0.0;
}
__global__ void slab_unpack_kernel(const double* vg, int P, int hpad, int H, int slab, int world, double* value) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x, p = blockIdx.y;
    if (h < H) { const int r = h / slab; value[(size_t)p * hpad + h] = vg[((size_t)r * P + p) * slab + (h - r * slab)]; }
}
__global__ void fill_int_kernel(int* p, int n, int v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

int angle_grid(double step, std::vector<double>& accum, std::vector<double>& list) {
    accum.clear(); list.clear();
    if (!(step > 0.0) || 6.3 / step < 1e-3) return 0;
    int g = 0;
    for (double a = 0; a < 6.3; a = a + 6.3 / step) {      // initRegistrationKSS.hpp:245
        accum.push_back(a);
        list.push_back((double)g * 6.3 / (double)step);    // initRegistrationKSS.hpp:282-284
        ++g;
    }
    return g;
}

int ensure_trig(kss_ctx* ctx, double step) {
    if (ctx->step == step && ctx->G > 0) return KSS_OK;
    int G = angle_grid(step, ctx->accum, ctx->list);
    if (G < 1 || G > 16) return fail(ctx, KSS_ERR_UNSUPPORTED, "step gives an angle grid outside 1..16 per axis");
    std::vector<double> ta(2 * G), tl(2 * G);
    for (int g = 0; g < G; ++g) {
        ta[2 * g] = std::cos(ctx->accum[g]); ta[2 * g + 1] = std::sin(ctx->accum[g]);
        tl[2 * g] = std::cos(ctx->list[g]);  tl[2 * g + 1] = std::sin(ctx->list[g]);
    }
    BUF("trig_accum", 2 * 16, &ctx->d_trig_accum);
    BUF("trig_list", 2 * 16, &ctx->d_trig_list);
    CU(cudaMemcpyAsync(ctx->d_trig_accum, ta.data(), sizeof(double) * 2 * G, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_trig_list, tl.data(), sizeof(double) * 2 * G, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));     // ta/tl are stack-owned
    ctx->step = step; ctx->G = G;
    return KSS_OK;
}

int counts_or_fill(kss_ctx* ctx, const char* name, const int* d_cnt, int P, int cap, const int** out) {
    if (d_cnt) { *out = d_cnt; return KSS_OK; }
    int* p;
    BUF(name, P, &p);
    fill_int_kernel<<<(P + 255) / 256, 256, 0, ctx->stream>>>(p, P, cap);
    KL(cudaGetLastError());
    *out = p;
    return KSS_OK;
}

inline int pad32(int n) { return (n + 31) / 32 * 32; }

// KSS_NO_CG=1 disables the candidate grid (tile search everywhere): A/B switch for tests and profiling
bool cg_enabled() { const char* e = getenv("KSS_NO_CG"); return !(e && e[0] == '1'); }

int cg_buffers(kss_ctx* ctx, int P, CgBuffers* cg) {
    BUF("cg_geom", (size_t)P * 8, &cg->geom);
    BUF("cg_hdr", (size_t)P * cg_hdr_words_per_pair(), &cg->hdr);
    BUF("cg_arena", (size_t)P * cg_arena_entries_per_pair(), &cg->arena);
    BUF("cg_cursor", (size_t)P, &cg->cursor);
    BUF("cg_ok", (size_t)P, &cg->ok);
    BUF("cg_wl", (size_t)P * cg_worklist_entries_per_pair(), &cg->wl);
    BUF("cg_wl_cnt", (size_t)P * 2, &cg->wl_cnt);
    BUF("cg_wl2", (size_t)P * cg_worklist2_entries_per_pair(), &cg->wl2);
    BUF("cg_wl2_cnt", (size_t)P * 2, &cg->wl2_cnt);
    return KSS_OK;
}

void icp_fill(IcpArgs& a, const kss_icp_params& prm) {
    a.max_iter = prm.max_iterations;
    a.max_dist_sqr = prm.max_corr_dist * prm.max_corr_dist;
    a.rot_thr = 1.0 - prm.transformation_eps;
    a.trans_thr = prm.transformation_eps;
    a.mse_rel = prm.fitness_eps;
    a.mse_abs = 1e-12;
}

// The whole KSSICP_Registration flow for pairs [0, P) whose inputs are device resident.
// All kernels go to ctx->stream, no host synchronisation.
int pipeline_device(kss_ctx* ctx, int P, const kss_batch& b, const double* sim_s, const double* sim_t,
                    const double* full_s, const double* full_t, const int* cnt_s, const int* cnt_t,
                    const int* cnt_S, const int* cnt_T, const int* h_cnt_S, const int* h_cnt_T, int slots,
                    kss_pair_result* d_out, double* d_point_align, bool shard) {
    cudaStream_t st = ctx->stream;
    const int G = ctx->G, H = G * G * G, hpad = H;
    const int cap_s = b.cap_s, cap_t = b.cap_t, cap_S = b.cap_S, cap_T = b.cap_T;
    const int cap_tpad = pad32(cap_t), cap_Tpad = pad32(cap_T);
    const int R = 1 + slots;
    double *align8, *s_al, *value, *run_fit; float* rbuf;
    unsigned short *s_perm, *t_inv, *S_perm, *T_inv;
    float4 *t_sorted, *T_sorted;
    float *t_box, *T_box, *run_T;
    int *best_h, *minima, *n_minima, *run_iters, *run_conv, *run_hyp, *run_tot;
    BUF("align8", (size_t)P * 8, &align8);
    BUF("s_al", (size_t)P * cap_s * 3, &s_al);
    BUF("s_perm", (size_t)P * cap_s, &s_perm);
    BUF("t_sorted", (size_t)P * cap_tpad, &t_sorted);
    BUF("t_box", (size_t)P * 6 * MAX_TILES, &t_box);
    BUF("t_inv", (size_t)P * cap_t, &t_inv);
    BUF("rbuf", (size_t)P * cap_s * hpad, &rbuf);
    BUF("value", (size_t)P * hpad, &value);
    BUF("best_h", (size_t)P, &best_h);
    BUF("minima", (size_t)P * hpad, &minima);
    BUF("n_minima", (size_t)P, &n_minima);
    BUF("run_T", (size_t)P * R * 16, &run_T);
    BUF("run_fit", (size_t)P * R, &run_fit);
    BUF("run_iters", (size_t)P * R, &run_iters);
    BUF("run_conv", (size_t)P * R, &run_conv);
    BUF("run_hyp", (size_t)P * R, &run_hyp);
    BUF("run_tot", (size_t)P * R * 2, &run_tot);

    {
        StageTimer tm(ctx, KSS_STAGE_PREP);
        KL(launch_middle_align(st, P, sim_s, cnt_s, cap_s, sim_t, cnt_t, cap_t, align8, s_al));
        KL(launch_sort_target(st, P, sim_t, cnt_t, cap_t, t_sorted, t_box, t_inv, cap_tpad));
        KL(launch_sort_source(st, P, sim_s, cnt_s, cap_s, s_perm));
    }
    CgBuffers cgb{}; const CgBuffers* cg = nullptr;
    if (cg_enabled()) {
        StageTimer tm(ctx, KSS_STAGE_CG_BUILD);
        int r = cg_buffers(ctx, P, &cgb); if (r) return r;
        int nl = 0;
        CU(launch_cg_build(st, P, 0, s_al, cnt_s, cap_s, nullptr, nullptr, 0, t_sorted, t_inv, cap_t, cnt_t, cap_tpad, cgb, &nl));
        ctx->launches += nl;
        cg = &cgb;
    }
    const bool sharded = shard && ctx->comm && ctx->world > 1;
    if (!sharded) {
        {
            StageTimer tm(ctx, KSS_STAGE_SWEEP);
            KL(launch_sweep(st, P, s_al, cnt_s, cap_s, s_perm, t_sorted, t_box, cnt_t, cap_tpad, ctx->d_trig_accum, G,
                            KSS_SCORE_AVE, rbuf, hpad, cg, 0, G * G));
        }
        {
            StageTimer tm(ctx, KSS_STAGE_SWEEP_FINALIZE);
            KL(launch_sweep_finalize(st, P, rbuf, cnt_s, cap_s, hpad, G, KSS_SCORE_AVE, value, best_h, minima, n_minima, 0, H, 3));
        }
    } else {
        // rotation hypotheses slabbed over the ranks (initRegistrationKSS.hpp:245-268 scores every (i, j, k) independently):
        // this rank sweeps the (i, j) pairs [ij_lo, ij_hi), sums its slab, ONE all-gather completes the grid everywhere,
        // then argmin + local minima (initRegistration_kernel needs the whole grid) run identically on every rank
        const int per = (G * G + ctx->world - 1) / ctx->world, slab = per * G;
        const int ij_lo = std::min(G * G, ctx->rank * per), ij_hi = std::min(G * G, ij_lo + per);
        double* vg;
        BUF("hyp_vg", (size_t)ctx->world * P * slab, &vg);
        {
            StageTimer tm(ctx, KSS_STAGE_SWEEP);
            KL(launch_sweep(st, P, s_al, cnt_s, cap_s, s_perm, t_sorted, t_box, cnt_t, cap_tpad, ctx->d_trig_accum, G,
                            KSS_SCORE_AVE, rbuf, hpad, cg, ij_lo, ij_hi));
        }
        StageTimer tm(ctx, KSS_STAGE_SWEEP_FINALIZE);
        KL(launch_sweep_finalize(st, P, rbuf, cnt_s, cap_s, hpad, G, KSS_SCORE_AVE, value, best_h, minima, n_minima, ij_lo * G, ij_hi * G, 1));
        double* mine = vg + (size_t)ctx->rank * P * slab;
        slab_pack_kernel<<<dim3((slab + 127) / 128, P), 128, 0, st>>>(value, P, hpad, ij_lo * G, ij_hi * G, slab, mine);
        KL(cudaGetLastError());
        NC(kss_nccl()->AllGather(mine, vg, (size_t)P * slab, ncclDouble, ctx->comm, st));
        slab_unpack_kernel<<<dim3((H + 127) / 128, P), 128, 0, st>>>(vg, P, hpad, H, slab, ctx->world, value);
        KL(cudaGetLastError());
        KL(launch_sweep_finalize(st, P, rbuf, cnt_s, cap_s, hpad, G, KSS_SCORE_AVE, value, best_h, minima, n_minima, 0, H, 2));
    }

    IcpArgs a{};
    a.src_f64 = sim_s; a.cnt_s = cnt_s; a.cap_s = cap_s; a.s_perm = s_perm;
    a.t_sorted = t_sorted; a.t_box = t_box; a.t_inv = t_inv; a.cnt_t = cnt_t; a.cap_t = cap_t; a.cap_tpad = cap_tpad;
    a.align8 = align8; a.runs_per_pair = R; a.hpad = hpad; a.G = G;
    a.best_h = best_h; a.minima = minima; a.n_minima = n_minima;
    a.trig_accum = ctx->d_trig_accum; a.trig_list = ctx->d_trig_list;
    a.judge_thr = b.judge_threshold;
    if (getenv("KSS_ICP_PHASES")) {       // diagnostics for tools/: 8 per-phase cycle counters, read with kss_debug_read
        unsigned long long* pc; BUF("icp_phase", 16, &pc);
        CU(cudaMemsetAsync(pc, 0, 128, st));
        a.phase_cycles = pc;
    }
    if (cg) { a.cg_geom = cg->geom; a.cg_hdr = cg->hdr; a.cg_arena = cg->arena; a.cg_ok = cg->ok; }
    icp_fill(a, b.icp);
    a.run_T = run_T; a.run_fit = run_fit; a.run_iters = run_iters; a.run_conv = run_conv;
    a.run_hyp = run_hyp; a.run_tot = run_tot; a.hyp_rank = 0; a.hyp_world = 1;
    a.mode = 3;                                                // judge (KSS_ICP.hpp:93) + hypothesis runs (:102-118), one launch
    double* pa = d_point_align;
    if (!pa) BUF("point_align", (size_t)P * cap_S * 3, &pa);
    if (!sharded) {
        {
            StageTimer tm(ctx, KSS_STAGE_ICP_HYP);
            KL(launch_icp(st, P, 1 + slots, a));
        }
        StageTimer tm(ctx, KSS_STAGE_SELECT_APPLY);
        KL(launch_select(st, P, R, hpad, G, b.judge_threshold, align8, run_fit, run_iters, run_T, run_conv, run_hyp,
                         run_tot, best_h, minima, n_minima, d_out));
        KL(launch_final_apply(st, P, full_s, cnt_S, cap_S, align8, d_out, ctx->d_trig_accum, ctx->d_trig_list, G, pa));
    } else {
        // ICP runs over angleList dealt round-robin to the ranks (KSS_ICP.hpp:102-118 is a serial loop over independent
        // runs); the kernel writes every run's fitness (and iteration count) at its l into a vector pre-filled with +inf,
        // and ONE all-reduce(MIN, double) on this ctx's stream completes it on every rank
        double* hyp_fit; double* run2_fit; float* run2_T; int *run2_iters, *run2_conv, *win_minima, *n_win;
        BUF("hyp_fit", (size_t)P * 2 * hpad, &hyp_fit);
        BUF("run2_T", (size_t)P * 2 * 16, &run2_T); BUF("run2_fit", (size_t)P * 2, &run2_fit);
        BUF("run2_iters", (size_t)P * 2, &run2_iters); BUF("run2_conv", (size_t)P * 2, &run2_conv);
        BUF("win_minima", (size_t)P * hpad, &win_minima); BUF("n_win", (size_t)P, &n_win);
        const size_t nfit = (size_t)P * 2 * hpad;
        fill_f64_kernel<<<(unsigned)((nfit + 255) / 256), 256, 0, st>>>(hyp_fit, nfit, (double)INFINITY);
        KL(cudaGetLastError());
        a.hyp_rank = ctx->rank; a.hyp_world = ctx->world; a.hyp_fit = hyp_fit;
        {
            StageTimer tm(ctx, KSS_STAGE_ICP_HYP);
            KL(launch_icp(st, P, 1 + slots, a));
            NC(kss_nccl()->AllReduce(hyp_fit, hyp_fit, nfit, ncclDouble, ncclMin, ctx->comm, st));
        }
        StageTimer tm(ctx, KSS_STAGE_SELECT_APPLY);
        KL(launch_select_sharded(st, P, R, hpad, G, b.judge_threshold, align8, run_fit, run_iters, hyp_fit, best_h, minima, n_minima,
                                 d_out, win_minima, n_win));
        // KSS_ICP.hpp:130: the final ICP repeats the winner's run -- on every rank, from the same start, hence identical
        IcpArgs f = a;
        f.mode = 1; f.runs_per_pair = 2; f.judge_thr = -1.0; f.minima = win_minima; f.n_minima = n_win;
        f.hyp_rank = 0; f.hyp_world = 1; f.hyp_fit = nullptr; f.run_hyp = nullptr; f.run_tot = nullptr;
        f.run_T = run2_T; f.run_fit = run2_fit; f.run_iters = run2_iters; f.run_conv = run2_conv;
        KL(launch_icp(st, P, 1, f));
        KL(launch_finish_sharded(st, P, R, run_fit, run_iters, run_conv, run_T, run2_fit, run2_iters, run2_conv, run2_T, d_out));
        KL(launch_final_apply(st, P, full_s, cnt_S, cap_S, align8, d_out, ctx->d_trig_accum, ctx->d_trig_list, G, pa));
    }
    StageTimer tm_metrics(ctx, KSS_STAGE_METRICS);
    if (cap_T <= SMALL_MAX && cap_S <= SMALL_MAX) {
        BUF("S_perm", (size_t)P * cap_S, &S_perm);
        BUF("T_sorted", (size_t)P * cap_Tpad, &T_sorted);
        BUF("T_box", (size_t)P * 6 * MAX_TILES, &T_box);
        BUF("T_inv", (size_t)P * cap_T, &T_inv);
        KL(launch_sort_target(st, P, full_t, cnt_T, cap_T, T_sorted, T_box, T_inv, cap_Tpad));
        KL(launch_sort_source(st, P, full_s, cnt_S, cap_S, S_perm));
        KL(launch_nn_small(st, P, 1, pa, cnt_S, cap_S, S_perm, T_sorted, T_box, cnt_T, cap_Tpad, nullptr, nullptr,
                           &d_out[0].mse, (int)(sizeof(kss_pair_result) / sizeof(double))));
    } else {
        // full-resolution clouds beyond the shared-memory path: block-grid NN per pair (every launch is enqueued without a
        // host synchronisation; h_cnt_* are the per-pair sizes on the host, null = capacity).  One pair: on this stream.
        // Several: round-robin over MQ side streams with their own buffers, forked from and joined to this stream.
        const int nq = P > 1 ? std::min(P, (int)kss_ctx::MQ) : 0;
        if (nq) {
            if (!ctx->mq_fork) CU(cudaEventCreateWithFlags(&ctx->mq_fork, cudaEventDisableTiming));
            for (int k = 0; k < nq; ++k) {
                if (!ctx->mq_stream[k]) CU(cudaStreamCreateWithFlags(&ctx->mq_stream[k], cudaStreamNonBlocking));
                if (!ctx->mq_done[k]) CU(cudaEventCreateWithFlags(&ctx->mq_done[k], cudaEventDisableTiming));
            }
            CU(cudaEventRecord(ctx->mq_fork, st));
            for (int k = 0; k < nq; ++k) CU(cudaStreamWaitEvent(ctx->mq_stream[k], ctx->mq_fork, 0));
        }
        for (int p = 0; p < P; ++p) {
            const int k = nq ? p % nq : 0;
            char pfx[16]; snprintf(pfx, sizeof(pfx), nq ? "pq%d:" : "", k);
            int r = large_metrics_device(nq ? ctx->mq_stream[k] : st, &ctx->launches, pa + (size_t)p * cap_S * 3, h_cnt_S ? h_cnt_S[p] : cap_S,
                                         full_t + (size_t)p * cap_T * 3, h_cnt_T ? h_cnt_T[p] : cap_T, &d_out[p].mse,
                                         [&](const char* name, size_t bytes, void** out) {
                                             unsigned char* q; int rr = dev_buf(ctx, (std::string(pfx) + name).c_str(), bytes, &q); *out = q; return rr; });
            if (r != KSS_OK) return fail(ctx, r, "large-cloud metrics failed");
        }
        for (int k = 0; k < nq; ++k) {
            CU(cudaEventRecord(ctx->mq_done[k], ctx->mq_stream[k]));
            CU(cudaStreamWaitEvent(st, ctx->mq_done[k], 0));
        }
    }
    return KSS_OK;
}

size_t per_pair_ws_bytes(const kss_batch& b, int H, int slots) {
    size_t v = (cg_enabled() ? cg_hdr_words_per_pair() * 8 + cg_arena_entries_per_pair() * 2 + cg_worklist_entries_per_pair() * 2 + cg_worklist2_entries_per_pair() * 16 + 64 : 0) +
               (size_t)b.cap_s * H * 4 + (size_t)b.cap_s * 30 + (size_t)pad32(b.cap_t) * 18 + (size_t)H * 12 +
               (size_t)(1 + slots) * 96 + (size_t)b.cap_S * 26 + (size_t)pad32(b.cap_T) * 18 + 8192;
    return v;
}

int check_batch(kss_ctx* ctx, const kss_batch* b) {
    if (!ctx) return KSS_ERR_ARG;
    if (!b || b->n_pairs < 1 || !b->full_s || !b->full_t || (!b->sim_s) != (!b->sim_t))
        return fail(ctx, KSS_ERR_ARG, "kss_batch: null pointer or n_pairs < 1");
    if (b->cap_S < 1 || b->cap_T < 1) return fail(ctx, KSS_ERR_ARG, "kss_batch: empty cloud");
    if (!b->sim_s) {                                   // raw clouds: the library runs AIVS itself
        if (b->cap_S < 4 || b->cap_T < 4) return fail(ctx, KSS_ERR_ARG, "kss_batch: raw clouds need at least 4 points");
        return KSS_OK;
    }
    if (b->cap_s < 1 || b->cap_t < 1) return fail(ctx, KSS_ERR_ARG, "kss_batch: empty cloud");
    if (b->cap_s > SMALL_MAX || b->cap_t > SMALL_MAX)
        return fail(ctx, KSS_ERR_UNSUPPORTED, "simplified clouds must have <= 2048 points (reference caps pNumber at 2000)");
    return KSS_OK;
}

}  // namespace

// ====================================================================== C ABI
extern "C" {

void kss_icp_params_default(kss_icp_params* p) {
    p->max_iterations = 1000; p->max_corr_dist = 1.0; p->transformation_eps = 1e-10; p->fitness_eps = 0.001;
}
void kss_batch_default(kss_batch* b) {
    std::memset(b, 0, sizeof(*b));
    b->step = 8.0; b->judge_threshold = 0.0005;
    kss_icp_params_default(&b->icp);
}

int kss_ctx_create_on_stream(int device, void* cuda_stream, kss_ctx** out) {
    if (!out) return KSS_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) { cudaGetLastError(); return KSS_ERR_CUDA; }
    if (device < 0 || device >= ndev) return KSS_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return KSS_ERR_CUDA;
    kss_ctx* c = new kss_ctx();
    c->device = device;
    if (cuda_stream) { c->stream = (cudaStream_t)cuda_stream; c->own_stream = false; }
    else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return KSS_ERR_CUDA; }
        c->own_stream = true;
    }
    const char* ws = getenv("KSS_WS_BYTES");
    if (ws) { long long v = atoll(ws); if (v > (1ll << 26)) c->ws_budget = (size_t)v; }
    *out = c;
    return KSS_OK;
}
int kss_ctx_create(int device, kss_ctx** out) { return kss_ctx_create_on_stream(device, nullptr, out); }

void kss_ctx_destroy(kss_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    collect_spans(ctx);
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    for (auto& kv : ctx->bufs) if (kv.second.p) cudaFree(kv.second.p);
    if (ctx->large_run.h_unres) cudaFreeHost(ctx->large_run.h_unres);
    for (int l = 0; l < kss_ctx::MAX_LANES; ++l) {
        if (ctx->lane_stream[l]) cudaStreamDestroy(ctx->lane_stream[l]);
        if (ctx->lane_done[l]) cudaEventDestroy(ctx->lane_done[l]);
    }
    if (ctx->fork_ev) cudaEventDestroy(ctx->fork_ev);
    for (int k = 0; k < kss_ctx::MQ; ++k) {
        if (ctx->mq_stream[k]) cudaStreamDestroy(ctx->mq_stream[k]);
        if (ctx->mq_done[k]) cudaEventDestroy(ctx->mq_done[k]);
    }
    if (ctx->mq_fork) cudaEventDestroy(ctx->mq_fork);
    if (ctx->comm && kss_nccl()) kss_nccl()->CommDestroy(ctx->comm);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}
const char* kss_last_error(kss_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
long long kss_ctx_launch_count(kss_ctx* ctx) { return ctx ? ctx->launches : 0; }
int kss_ctx_set_hyp_slots(kss_ctx* ctx, int slots) {
    if (!ctx) return KSS_ERR_ARG;
    if (slots < 1 || slots > 729) return fail(ctx, KSS_ERR_ARG, "kss_ctx_set_hyp_slots: 1..729");
    ctx->hyp_slots = slots;
    return KSS_OK;
}
int kss_ctx_synchronize(kss_ctx* ctx) {
    if (!ctx) return KSS_ERR_ARG;
    CU(cudaStreamSynchronize(ctx->stream));
    return KSS_OK;
}

int kss_ctx_set_timing(kss_ctx* ctx, int enable) {
    if (!ctx) return KSS_ERR_ARG;
    collect_spans(ctx);
    ctx->timing = enable != 0;
    for (int i = 0; i < KSS_STAGE_COUNT; ++i) { ctx->stage_ms[i] = 0; ctx->stage_calls[i] = 0; }
    return KSS_OK;
}
int kss_ctx_stage_ms(kss_ctx* ctx, int stage, double* ms, long long* calls) {
    if (!ctx || stage < 0 || stage >= KSS_STAGE_COUNT) return KSS_ERR_ARG;
    cudaStreamSynchronize(ctx->stream);
    collect_spans(ctx);
    if (ms) *ms = ctx->stage_ms[stage];
    if (calls) *calls = ctx->stage_calls[stage];
    return KSS_OK;
}

int kss_icp_large_begin(kss_ctx* ctx, const double* src, int n_s, const double* tgt, int n_t) {
    if (!ctx) return KSS_ERR_ARG;
    if (!src || !tgt || n_s < 1 || n_t < 1) return fail(ctx, KSS_ERR_ARG, "kss_icp_large_begin: bad argument");
    CU(cudaSetDevice(ctx->device));
    double *d_s, *d_t;
    ctx->large_ready = false;
    BUF("run:lg_in_s", (size_t)n_s * 3, &d_s); BUF("run:lg_in_t", (size_t)n_t * 3, &d_t);
    CU(cudaMemcpyAsync(d_s, src, sizeof(double) * 3 * (size_t)n_s, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_t, tgt, sizeof(double) * 3 * (size_t)n_t, cudaMemcpyHostToDevice, ctx->stream));
    auto alloc = [&](const char* name, size_t bytes, void** out) {
        unsigned char* q; int rr = dev_buf(ctx, name, bytes, &q); *out = q; return rr; };
    int r;
    {
        StageTimer tm(ctx, KSS_STAGE_LARGE_BUILD);
        r = large_icp_prepare(ctx->stream, &ctx->launches, d_s, n_s, d_t, n_t, alloc, &ctx->large_run, "run:");
    }
    if (r != KSS_OK) return fail(ctx, r, "kss_icp_large_begin: build failed");
    ctx->large_run.mark_user = ctx;
    ctx->large_run.mark = [](void* u, int stage, int begin) {
        kss_ctx* c = (kss_ctx*)u;
        if (!c->timing) return;
        if (begin) { c->open_ev[stage] = ev_get(c); cudaEventRecord(c->open_ev[stage], c->stream); }
        else if (c->open_ev[stage]) {
            cudaEvent_t b = ev_get(c); cudaEventRecord(b, c->stream);
            c->spans.push_back({stage, c->open_ev[stage], b}); c->open_ev[stage] = nullptr;
        }
    };
    ctx->large_ready = true;
    return KSS_OK;
}
int kss_icp_large_iterate(kss_ctx* ctx, const kss_icp_params* prm, int count) {
    if (!ctx || !prm) return KSS_ERR_ARG;
    if (!ctx->large_ready) return fail(ctx, KSS_ERR_ARG, "kss_icp_large_iterate: call kss_icp_large_begin first");
    int r = large_icp_iterations(ctx->stream, &ctx->launches, &ctx->large_run, prm, count);
    if (r != KSS_OK) return fail(ctx, r, "kss_icp_large_iterate failed");
    return KSS_OK;
}
int kss_icp_large_end(kss_ctx* ctx, const kss_icp_params* prm, float T[16], double* fitness, int* iters, int* converged) {
    if (!ctx || !prm) return KSS_ERR_ARG;
    if (!ctx->large_ready) return fail(ctx, KSS_ERR_ARG, "kss_icp_large_end: call kss_icp_large_begin first");
    kss_icp_params stop = *prm;
    stop.max_iterations = 0;                       // no further iterations: fitness pass only
    int r = large_icp_run(ctx->stream, &ctx->launches, &ctx->large_run, &stop, 1);
    if (r == KSS_OK) r = large_icp_result(ctx->stream, &ctx->large_run, T, fitness, iters, converged);
    if (r != KSS_OK) return fail(ctx, r, "kss_icp_large_end failed");
    return KSS_OK;
}

int kss_sweep_angles(double step, double* accum, double* list, int cap) {
    std::vector<double> a, l;
    int g = angle_grid(step, a, l);
    for (int i = 0; i < g && i < cap; ++i) { if (accum) accum[i] = a[i]; if (list) list[i] = l[i]; }
    return g;
}

int kss_middle_align(kss_ctx* ctx, const double* src, int n_s, const double* tgt, int n_t, double out7[7],
                     double* src_aligned) {
    if (!ctx) return KSS_ERR_ARG;
    if (!src || !tgt || !out7 || n_s < 1 || n_t < 1) return fail(ctx, KSS_ERR_ARG, "kss_middle_align: bad argument");
    if (n_s > SMALL_MAX || n_t > SMALL_MAX) return fail(ctx, KSS_ERR_UNSUPPORTED, "kss_middle_align: > 2048 points");
    CU(cudaSetDevice(ctx->device));
    double *d_s, *d_t, *d_a8, *d_al; const int *c_s, *c_t;
    BUF("one_s", (size_t)n_s * 3, &d_s); BUF("one_t", (size_t)n_t * 3, &d_t);
    BUF("align8", 8, &d_a8); BUF("s_al", (size_t)n_s * 3, &d_al);
    CU(cudaMemcpyAsync(d_s, src, sizeof(double) * 3 * n_s, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_t, tgt, sizeof(double) * 3 * n_t, cudaMemcpyHostToDevice, ctx->stream));
    int r = counts_or_fill(ctx, "cnt_s", nullptr, 1, n_s, &c_s); if (r) return r;
    r = counts_or_fill(ctx, "cnt_t", nullptr, 1, n_t, &c_t); if (r) return r;
    KL(launch_middle_align(ctx->stream, 1, d_s, c_s, n_s, d_t, c_t, n_t, d_a8, d_al));
    double a8[8];
    CU(cudaMemcpyAsync(a8, d_a8, sizeof(a8), cudaMemcpyDeviceToHost, ctx->stream));
    if (src_aligned) CU(cudaMemcpyAsync(src_aligned, d_al, sizeof(double) * 3 * n_s, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 7; ++i) out7[i] = a8[i];
    return KSS_OK;
}

int kss_rotation_sweep(kss_ctx* ctx, const double* src_aligned, int n_s, const double* tgt, int n_t, double step,
                       int score_mode, double* value, int* G_out, double best_angle[3], int best_index[3],
                       int* minima, int* n_minima) {
    if (!ctx) return KSS_ERR_ARG;
    if (!src_aligned || !tgt || n_s < 1 || n_t < 1) return fail(ctx, KSS_ERR_ARG, "kss_rotation_sweep: bad argument");
    if (n_s > SMALL_MAX || n_t > SMALL_MAX) return fail(ctx, KSS_ERR_UNSUPPORTED, "kss_rotation_sweep: > 2048 points");
    if (score_mode < 0 || score_mode > KSS_SCORE_VOXEL) return fail(ctx, KSS_ERR_ARG, "kss_rotation_sweep: score_mode");
    CU(cudaSetDevice(ctx->device));
    int r = ensure_trig(ctx, step); if (r) return r;
    const int G = ctx->G, H = G * G * G;
    double *d_s, *d_t, *d_val; float* rbuf; const int *c_s, *c_t;
    unsigned short *s_perm, *t_inv; float4* t_sorted; float* t_box; int *d_best, *d_min, *d_nmin;
    const int tpad = pad32(n_t);
    BUF("one_s", (size_t)n_s * 3, &d_s); BUF("one_t", (size_t)n_t * 3, &d_t);
    BUF("s_perm", n_s, &s_perm); BUF("t_sorted", tpad, &t_sorted); BUF("t_box", 6 * MAX_TILES, &t_box); BUF("t_inv", n_t, &t_inv);
    BUF("rbuf", (size_t)n_s * H, &rbuf); BUF("value", H, &d_val); BUF("best_h", 1, &d_best);
    BUF("minima", H, &d_min); BUF("n_minima", 1, &d_nmin);
    CU(cudaMemcpyAsync(d_s, src_aligned, sizeof(double) * 3 * n_s, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_t, tgt, sizeof(double) * 3 * n_t, cudaMemcpyHostToDevice, ctx->stream));
    r = counts_or_fill(ctx, "cnt_s", nullptr, 1, n_s, &c_s); if (r) return r;
    r = counts_or_fill(ctx, "cnt_t", nullptr, 1, n_t, &c_t); if (r) return r;
    KL(launch_sort_target(ctx->stream, 1, d_t, c_t, n_t, t_sorted, t_box, t_inv, tpad));
    KL(launch_sort_source(ctx->stream, 1, d_s, c_s, n_s, s_perm));
    CgBuffers cgb{}; const CgBuffers* cg = nullptr;
    if (cg_enabled()) {
        r = cg_buffers(ctx, 1, &cgb); if (r) return r;
        int nl = 0;
        CU(launch_cg_build(ctx->stream, 1, 0, d_s, c_s, n_s, nullptr, nullptr, 0, t_sorted, t_inv, n_t, c_t, tpad, cgb, &nl));
        ctx->launches += nl;
        cg = &cgb;
    }
    KL(launch_sweep(ctx->stream, 1, d_s, c_s, n_s, s_perm, t_sorted, t_box, c_t, tpad, ctx->d_trig_accum, G, score_mode, rbuf, H, cg, 0, G * G));
    KL(launch_sweep_finalize(ctx->stream, 1, rbuf, c_s, n_s, H, G, score_mode, d_val, d_best, d_min, d_nmin, 0, H, 3));
    std::vector<int> hmin(H);
    int hbest = 0, nmin = 0;
    if (value) CU(cudaMemcpyAsync(value, d_val, sizeof(double) * H, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&hbest, d_best, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&nmin, d_nmin, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(hmin.data(), d_min, sizeof(int) * H, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (G_out) *G_out = G;
    const int bi[3] = {hbest / (G * G), (hbest / G) % G, hbest % G};
    for (int k = 0; k < 3; ++k) {
        if (best_index) best_index[k] = bi[k];
        if (best_angle) best_angle[k] = ctx->accum[bi[k]];
    }
    if (n_minima) *n_minima = nmin;
    if (minima)
        for (int l = 0; l < nmin; ++l) {
            minima[3 * l] = hmin[l] / (G * G); minima[3 * l + 1] = (hmin[l] / G) % G; minima[3 * l + 2] = hmin[l] % G;
        }
    return KSS_OK;
}

int kss_apply_similarity(kss_ctx* ctx, const double* pts, int n, const double align7[7], const double angles[3],
                         double* out) {
    if (!ctx) return KSS_ERR_ARG;
    if (!pts || !align7 || !angles || !out || n < 1) return fail(ctx, KSS_ERR_ARG, "kss_apply_similarity: bad argument");
    CU(cudaSetDevice(ctx->device));
    double *d_p, *d_o, *d_par;
    BUF("pts_in", (size_t)n * 3, &d_p); BUF("pts_out", (size_t)n * 3, &d_o); BUF("sim_par", 16, &d_par);
    double par[16] = {0};
    for (int i = 0; i < 7; ++i) par[i] = align7[i];
    for (int k = 0; k < 3; ++k) { par[8 + 2 * k] = std::cos(angles[k]); par[9 + 2 * k] = std::sin(angles[k]); }
    CU(cudaMemcpyAsync(d_p, pts, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_par, par, sizeof(par), cudaMemcpyHostToDevice, ctx->stream));
    KL(launch_apply_similarity(ctx->stream, d_p, n, d_par, d_par + 8, d_o));
    CU(cudaMemcpyAsync(out, d_o, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KSS_OK;
}

int kss_apply_transform(kss_ctx* ctx, const float T[16], const double* pts, int n, double* out) {
    if (!ctx) return KSS_ERR_ARG;
    if (!T || !pts || !out || n < 1) return fail(ctx, KSS_ERR_ARG, "kss_apply_transform: bad argument");
    CU(cudaSetDevice(ctx->device));
    double *d_p, *d_o; float* d_T;
    BUF("pts_in", (size_t)n * 3, &d_p); BUF("pts_out", (size_t)n * 3, &d_o); BUF("T16", 16, &d_T);
    CU(cudaMemcpyAsync(d_p, pts, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_T, T, sizeof(float) * 16, cudaMemcpyHostToDevice, ctx->stream));
    KL(launch_apply_transform(ctx->stream, d_p, n, d_T, d_o));
    CU(cudaMemcpyAsync(out, d_o, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KSS_OK;
}

int kss_icp(kss_ctx* ctx, const double* src, int n_s, const double* tgt, int n_t, const kss_icp_params* prm,
            float T[16], double* fitness, int* iters, int* converged, kss_icp_trace* trace) {
    if (!ctx) return KSS_ERR_ARG;
    if (!src || !tgt || n_s < 1 || n_t < 1) return fail(ctx, KSS_ERR_ARG, "kss_icp: bad argument");
    kss_icp_params dp; kss_icp_params_default(&dp);
    if (!prm) prm = &dp;
    CU(cudaSetDevice(ctx->device));
    auto alloc = [&](const char* name, size_t bytes, void** out) {
        unsigned char* q; int rr = dev_buf(ctx, name, bytes, &q); *out = q; return rr; };
    if (n_s > SMALL_MAX || n_t > SMALL_MAX) {
        // full-resolution overload shapeRegistration_ICP(int iter) (KSS_ICP.hpp:133-183)
        if (trace) return fail(ctx, KSS_ERR_UNSUPPORTED, "kss_icp: trace is only available on the <= 2048-point path");
        int r = large_icp_host(ctx->stream, &ctx->launches, src, n_s, tgt, n_t, prm, T, fitness, iters, converged, alloc);
        if (r != KSS_OK) return fail(ctx, r, "kss_icp (large path) failed");
        return KSS_OK;
    }
    double *d_s, *d_t, *run_fit; const int *c_s, *c_t;
    unsigned short *s_perm, *t_inv; float4* t_sorted; float *t_box, *run_T; int *run_iters, *run_conv;
    const int tpad = pad32(n_t);
    BUF("one_s", (size_t)n_s * 3, &d_s); BUF("one_t", (size_t)n_t * 3, &d_t);
    BUF("s_perm", n_s, &s_perm); BUF("t_sorted", tpad, &t_sorted); BUF("t_box", 6 * MAX_TILES, &t_box); BUF("t_inv", n_t, &t_inv);
    BUF("run_T", 16, &run_T); BUF("run_fit", 1, &run_fit); BUF("run_iters", 1, &run_iters); BUF("run_conv", 1, &run_conv);
    CU(cudaMemcpyAsync(d_s, src, sizeof(double) * 3 * n_s, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_t, tgt, sizeof(double) * 3 * n_t, cudaMemcpyHostToDevice, ctx->stream));
    int r = counts_or_fill(ctx, "cnt_s", nullptr, 1, n_s, &c_s); if (r) return r;
    r = counts_or_fill(ctx, "cnt_t", nullptr, 1, n_t, &c_t); if (r) return r;
    KL(launch_sort_target(ctx->stream, 1, d_t, c_t, n_t, t_sorted, t_box, t_inv, tpad));
    KL(launch_sort_source(ctx->stream, 1, d_s, c_s, n_s, s_perm));
    IcpArgs a{};
    a.src_f64 = d_s; a.cnt_s = c_s; a.cap_s = n_s; a.s_perm = s_perm;
    a.t_sorted = t_sorted; a.t_box = t_box; a.t_inv = t_inv; a.cnt_t = c_t; a.cap_t = n_t; a.cap_tpad = tpad;
    a.mode = 2; a.runs_per_pair = 1; a.judge_thr = -1.0; a.hyp_world = 1;
    CgBuffers cgb{};
    if (cg_enabled()) {
        r = cg_buffers(ctx, 1, &cgb); if (r) return r;
        int nl = 0;
        CU(launch_cg_build(ctx->stream, 1, 1, d_s, c_s, n_s, d_t, c_t, n_t, t_sorted, t_inv, n_t, c_t, tpad, cgb, &nl));
        ctx->launches += nl;
        a.cg_geom = cgb.geom; a.cg_hdr = cgb.hdr; a.cg_arena = cgb.arena; a.cg_ok = cgb.ok;
    }
    icp_fill(a, *prm);
    a.run_T = run_T; a.run_fit = run_fit; a.run_iters = run_iters; a.run_conv = run_conv;
    int cap = 0;
    if (trace && trace->cap_iters > 0) {
        cap = trace->cap_iters; a.trace_cap = cap;
        if (trace->corr_idx) { BUF("trace_idx", (size_t)cap * n_s, &a.trace_idx); CU(cudaMemsetAsync(a.trace_idx, 0xff, sizeof(int) * (size_t)cap * n_s, ctx->stream)); }
        if (trace->T_k) { BUF("trace_T", (size_t)cap * 16, &a.trace_T); CU(cudaMemsetAsync(a.trace_T, 0, sizeof(float) * (size_t)cap * 16, ctx->stream)); }
        if (trace->mse) { BUF("trace_mse", (size_t)cap, &a.trace_mse); CU(cudaMemsetAsync(a.trace_mse, 0, sizeof(double) * (size_t)cap, ctx->stream)); }
        if (trace->src_k) { BUF("trace_src", (size_t)cap * n_s * 3, &a.trace_src); CU(cudaMemsetAsync(a.trace_src, 0, sizeof(float) * (size_t)cap * n_s * 3, ctx->stream)); }
    }
    KL(launch_icp(ctx->stream, 1, 1, a));
    float hT[16]; double hf = 0; int hi = 0, hc = 0;
    CU(cudaMemcpyAsync(hT, run_T, sizeof(hT), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&hf, run_fit, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&hi, run_iters, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&hc, run_conv, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    if (cap > 0) {
        if (trace->corr_idx) CU(cudaMemcpyAsync(trace->corr_idx, a.trace_idx, sizeof(int) * (size_t)cap * n_s, cudaMemcpyDeviceToHost, ctx->stream));
        if (trace->T_k) CU(cudaMemcpyAsync(trace->T_k, a.trace_T, sizeof(float) * (size_t)cap * 16, cudaMemcpyDeviceToHost, ctx->stream));
        if (trace->mse) CU(cudaMemcpyAsync(trace->mse, a.trace_mse, sizeof(double) * (size_t)cap, cudaMemcpyDeviceToHost, ctx->stream));
        if (trace->src_k) CU(cudaMemcpyAsync(trace->src_k, a.trace_src, sizeof(float) * (size_t)cap * n_s * 3, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    if (T) std::memcpy(T, hT, sizeof(hT));
    if (fitness) *fitness = hf;
    if (iters) *iters = hi;
    if (converged) *converged = hc;
    return KSS_OK;
}

static int nn_common(kss_ctx* ctx, const double* q, int n_q, const double* t, int n_t, int mode, int32_t* idx,
                     float* d2, double* out3) {
    CU(cudaSetDevice(ctx->device));
    auto alloc = [&](const char* name, size_t bytes, void** out) {
        unsigned char* p; int rr = dev_buf(ctx, name, bytes, &p); *out = p; return rr; };
    double *d_q, *d_t;
    BUF("one_s", (size_t)n_q * 3, &d_q); BUF("one_t", (size_t)n_t * 3, &d_t);
    CU(cudaMemcpyAsync(d_q, q, sizeof(double) * 3 * n_q, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_t, t, sizeof(double) * 3 * n_t, cudaMemcpyHostToDevice, ctx->stream));
    const int *c_q, *c_t;
    int r = counts_or_fill(ctx, "cnt_s", nullptr, 1, n_q, &c_q); if (r) return r;
    r = counts_or_fill(ctx, "cnt_t", nullptr, 1, n_t, &c_t); if (r) return r;
    int* d_idx = nullptr; float* d_d2 = nullptr; double* d_o3 = nullptr;
    if (mode == 0) { BUF("nn_idx", n_q, &d_idx); BUF("nn_d2", n_q, &d_d2); }
    else BUF("nn_out3", 3, &d_o3);
    if (n_q <= SMALL_MAX && n_t <= SMALL_MAX) {
        unsigned short *q_perm, *t_inv; float4* t_sorted; float* t_box;
        const int tpad = pad32(n_t);
        BUF("s_perm", n_q, &q_perm); BUF("t_sorted", tpad, &t_sorted); BUF("t_box", 6 * MAX_TILES, &t_box); BUF("t_inv", n_t, &t_inv);
        KL(launch_sort_target(ctx->stream, 1, d_t, c_t, n_t, t_sorted, t_box, t_inv, tpad));
        KL(launch_sort_source(ctx->stream, 1, d_q, c_q, n_q, q_perm));
        KL(launch_nn_small(ctx->stream, 1, mode, d_q, c_q, n_q, q_perm, t_sorted, t_box, c_t, tpad, d_idx, d_d2, d_o3, 3));
    } else if (mode == 0) {
        r = large_nn_device(ctx->stream, &ctx->launches, d_q, n_q, d_t, n_t, d_idx, d_d2, alloc);
        if (r != KSS_OK) return fail(ctx, r, "large-cloud NN failed");
    } else {
        r = large_metrics_device(ctx->stream, &ctx->launches, d_q, n_q, d_t, n_t, d_o3, alloc);
        if (r != KSS_OK) return fail(ctx, r, "large-cloud metrics failed");
    }
    if (mode == 0) {
        if (idx) CU(cudaMemcpyAsync(idx, d_idx, sizeof(int) * n_q, cudaMemcpyDeviceToHost, ctx->stream));
        if (d2) CU(cudaMemcpyAsync(d2, d_d2, sizeof(float) * n_q, cudaMemcpyDeviceToHost, ctx->stream));
    } else {
        CU(cudaMemcpyAsync(out3, d_o3, sizeof(double) * 3, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return KSS_OK;
}

int kss_nn_search(kss_ctx* ctx, const double* q, int n_q, const double* t, int n_t, int32_t* idx, float* d2) {
    if (!ctx) return KSS_ERR_ARG;
    if (!q || !t || n_q < 1 || n_t < 1) return fail(ctx, KSS_ERR_ARG, "kss_nn_search: bad argument");
    return nn_common(ctx, q, n_q, t, n_t, 0, idx, d2, nullptr);
}

int kss_nn_metrics(kss_ctx* ctx, const double* a, int n_a, const double* t, int n_t, double out3[3]) {
    if (!ctx) return KSS_ERR_ARG;
    if (!a || !t || !out3 || n_a < 1 || n_t < 1) return fail(ctx, KSS_ERR_ARG, "kss_nn_metrics: bad argument");
    return nn_common(ctx, a, n_a, t, n_t, 1, nullptr, nullptr, out3);
}

namespace {

int aivs_status(kss_ctx* ctx, int bad) {
    if (bad == 1) return fail(ctx, KSS_ERR_UNSUPPORTED, "AIVS: a cloud has zero extent or needs more boxes than its point count allows");
    if (bad == 2) return fail(ctx, KSS_ERR_UNSUPPORTED, "AIVS: more than 16384 samples before the trim step");
    if (bad == 3) return fail(ctx, KSS_ERR_UNSUPPORTED, "AIVS: output capacity too small");
    return KSS_OK;
}

int ensure_lanes(kss_ctx* ctx, int n) {
    for (int l = 0; l < n; ++l) {
        if (!ctx->lane_stream[l]) CU(cudaStreamCreateWithFlags(&ctx->lane_stream[l], cudaStreamNonBlocking));
        if (!ctx->lane_done[l]) CU(cudaEventCreateWithFlags(&ctx->lane_done[l], cudaEventDisableTiming));
    }
    if (!ctx->fork_ev) CU(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
    return KSS_OK;
}

// The batch in chunks; `host` = b's clouds/counts and results/point_align are host memory (copied per chunk on the
// chunk's lane), else device memory.  Everything is ordered after the work already on ctx->stream, and ctx->stream
// waits for all lanes at the end.
int batch_core(kss_ctx* ctx, const kss_batch* b_in, bool host, kss_pair_result* results, double* point_align, bool shard = false) {
    kss_batch bb = *b_in;
    const bool raw = !bb.sim_s;
    if (raw) {   // pNumber (KSS_ICP.hpp:53-67) plus room for a trim step that stops early on its stale neighbour lists
        const int pn = std::min(2000, std::min(bb.cap_S, bb.cap_T) / 2);
        bb.cap_s = std::min(std::min(pn + 48, SMALL_MAX), bb.cap_S);
        bb.cap_t = std::min(std::min(pn + 48, SMALL_MAX), bb.cap_T);
    }
    int r = ensure_trig(ctx, bb.step); if (r) return r;
    const int H = ctx->G * ctx->G * ctx->G;
    int slots = ctx->hyp_slots;
    const char* es = getenv("KSS_HYP_SLOTS");                        // A/B switch for tests
    if (es && atoi(es) > 0) slots = atoi(es);
    if (slots > H) slots = H;
    int lanes = 2;
    const char* el = getenv("KSS_LANES");
    if (el && atoi(el) > 0) lanes = std::min(atoi(el), (int)kss_ctx::MAX_LANES);
    const size_t per = per_pair_ws_bytes(bb, H, slots) + (raw ? (size_t)(bb.cap_S + bb.cap_T) * 64 + 65536 : 0) +
                       (host ? (size_t)(bb.cap_S + bb.cap_T + bb.cap_s + bb.cap_t) * 24 : 0);
    const int NP = bb.n_pairs;
    int chunk = (int)std::min<size_t>((size_t)NP, std::max<size_t>(1, ctx->ws_budget / (per * (size_t)lanes)));
    // two chunks per lane when the batch is big enough for every chunk to give each SM a pair; equal chunk sizes
    if (lanes > 1) {
        // measured on B200 (2,468 / 1,234 / 617 / 309 pairs): two chunks hide one chunk's thin ICP tail behind the other's
        // kernels (+8 % at 309-617 pairs, neutral at 2,468).  Host buffers: the second chunk's copies overlap the first
        // chunk's kernels either way; 2,468 pairs end to end 78.4 ms with 2 chunks, 79.8 with 3, 79.6 with 4 (round 2 kernels)
        int want = NP >= 296 ? 2 : 1;
        const char* em = getenv("KSS_CHUNKS");
        if (em && atoi(em) > 0) want = atoi(em);
        int nchunks = std::max((NP + chunk - 1) / chunk, want);
        if (nchunks < 1) nchunks = 1;
        chunk = (NP + nchunks - 1) / nchunks;
    }
    if (shard) {     // collectives must be issued in the same order on every rank: one lane, chunks in sequence
        lanes = 1;
        chunk = (int)std::min<size_t>((size_t)NP, std::max<size_t>(1, ctx->ws_budget / per));
    }
    if (chunk >= NP) lanes = 1;
    r = ensure_lanes(ctx, lanes); if (r) return r;
    cudaStream_t main_st = ctx->stream;
    int* d_bad = nullptr;
    if (raw) { BUF("aivs_bad", 1, &d_bad); CU(cudaMemsetAsync(d_bad, 0, sizeof(int), main_st)); ctx->aivs_bad = d_bad; }
    CU(cudaEventRecord(ctx->fork_ev, main_st));
    for (int l = 0; l < lanes; ++l) CU(cudaStreamWaitEvent(ctx->lane_stream[l], ctx->fork_ev, 0));
    auto alloc = [&](const char* name, size_t bytes, void** out) {
        unsigned char* q; int rr = dev_buf(ctx, name, bytes, &q); *out = q; return rr; };
    int rc = KSS_OK, ci = 0;
    // host buffers, two chunks: nothing runs until the first chunk's clouds are on the device, so the first chunk is the
    // smaller one (KSS_CHUNK_SKEW = its share).  Measured on the 2,468-pair batch, end to end: 0.40 -> 76.8 ms, while 0.35,
    // 0.45 and 0.50 all give 79.0 ms (three runs each): the gain is as much the two lanes falling out of step (one lane's
    // thin ICP tail under the other's sweep) as the shorter first copy, and it is specific to this split
    int first = chunk;
    if (host && lanes == 2 && chunk < NP && 2 * chunk >= NP) {
        static const double skew = [] { const char* e = getenv("KSS_CHUNK_SKEW"); const double v = e ? atof(e) : 0.0; return v > 0.0 && v < 1.0 ? v : 0.4; }();
        first = std::max(1, (int)(NP * skew + 0.5));
        chunk = std::max(first, NP - first);
    }
    for (int p0 = 0, P = 0; p0 < NP && rc == KSS_OK; p0 += P, ++ci) {
        P = std::min(ci == 0 ? first : chunk, NP - p0);
        const int lane = ci % lanes;
        cudaStream_t st = ctx->lane_stream[lane];
        ctx->stream = st;
        ctx->buf_suffix = "#" + std::to_string(lane);
        rc = [&]() -> int {
            const double *full_s, *full_t, *sim_s = nullptr, *sim_t = nullptr;
            const int *c_s = nullptr, *c_t = nullptr, *c_S = nullptr, *c_T = nullptr;
            kss_pair_result* d_res = results ? results + p0 : nullptr;
            double* d_pa = point_align ? point_align + (size_t)p0 * bb.cap_S * 3 : nullptr;
            if (host) {
                double *d_fs, *d_ft;
                BUF("in_full_s", (size_t)P * bb.cap_S * 3, &d_fs); BUF("in_full_t", (size_t)P * bb.cap_T * 3, &d_ft);
                CU(cudaMemcpyAsync(d_ft, bb.full_t + (size_t)p0 * bb.cap_T * 3, sizeof(double) * 3 * (size_t)P * bb.cap_T, cudaMemcpyHostToDevice, st));
                CU(cudaMemcpyAsync(d_fs, bb.full_s + (size_t)p0 * bb.cap_S * 3, sizeof(double) * 3 * (size_t)P * bb.cap_S, cudaMemcpyHostToDevice, st));
                full_s = d_fs; full_t = d_ft;
                if (!raw) {
                    double *d_ss, *d_st;
                    BUF("in_sim_s", (size_t)P * bb.cap_s * 3, &d_ss); BUF("in_sim_t", (size_t)P * bb.cap_t * 3, &d_st);
                    CU(cudaMemcpyAsync(d_ss, bb.sim_s + (size_t)p0 * bb.cap_s * 3, sizeof(double) * 3 * (size_t)P * bb.cap_s, cudaMemcpyHostToDevice, st));
                    CU(cudaMemcpyAsync(d_st, bb.sim_t + (size_t)p0 * bb.cap_t * 3, sizeof(double) * 3 * (size_t)P * bb.cap_t, cudaMemcpyHostToDevice, st));
                    sim_s = d_ss; sim_t = d_st;
                }
                auto up = [&](const char* name, const int* h, const int** out) -> int {
                    if (!h) return KSS_OK;
                    int* d; BUF(name, P, &d);
                    CU(cudaMemcpyAsync(d, h + p0, sizeof(int) * P, cudaMemcpyHostToDevice, st));
                    *out = d; return KSS_OK;
                };
                int q;
                if (!raw) { q = up("in_cnt_s", bb.cnt_s, &c_s); if (q) return q; q = up("in_cnt_t", bb.cnt_t, &c_t); if (q) return q; }
                q = up("in_cnt_S", bb.cnt_S, &c_S); if (q) return q;
                q = up("in_cnt_T", bb.cnt_T, &c_T); if (q) return q;
                BUF("out_res", (size_t)P, &d_res);
                if (point_align) BUF("out_pa", (size_t)P * bb.cap_S * 3, &d_pa);
            } else {
                full_s = bb.full_s + (size_t)p0 * bb.cap_S * 3; full_t = bb.full_t + (size_t)p0 * bb.cap_T * 3;
                if (!raw) { sim_s = bb.sim_s + (size_t)p0 * bb.cap_s * 3; sim_t = bb.sim_t + (size_t)p0 * bb.cap_t * 3; }
                if (!raw && bb.cnt_s) c_s = bb.cnt_s + p0;
                if (!raw && bb.cnt_t) c_t = bb.cnt_t + p0;
                if (bb.cnt_S) c_S = bb.cnt_S + p0;
                if (bb.cnt_T) c_T = bb.cnt_T + p0;
            }
            int q = counts_or_fill(ctx, "cnt_S", c_S, P, bb.cap_S, &c_S); if (q) return q;
            q = counts_or_fill(ctx, "cnt_T", c_T, P, bb.cap_T, &c_T); if (q) return q;
            if (raw) {
                // KSSICP_Registration's first half (KSS_ICP.hpp:72-84): AIVS_simplification(pNumber) of target and source
                StageTimer tm(ctx, KSS_STAGE_AIVS);
                double *d_ss, *d_st; int *d_pn, *d_cs, *d_ct;
                BUF("aivs_sim_s", (size_t)P * bb.cap_s * 3, &d_ss); BUF("aivs_sim_t", (size_t)P * bb.cap_t * 3, &d_st);
                BUF("aivs_pn", P, &d_pn); BUF("aivs_cnt_s", P, &d_cs); BUF("aivs_cnt_t", P, &d_ct);
                q = aivs_pnumber_device(st, &ctx->launches, P, c_S, bb.cap_S, c_T, bb.cap_T, d_pn);
                // Clouds beyond one CTA (the any-size kernels: ~30 small launches per batch of clouds, mostly chains of latencies):
                // the sources on a side stream next to the targets (separate scratch, tags "s" / "t"; one pair of 10k points
                // 3.5 -> 3.1 ms, 64 of them: AIVS 6.3 -> 5.3 ms).  Batches of small clouds are throughput-bound, one stream.
                const bool side = (bb.cap_S > SMALL_MAX || bb.cap_T > SMALL_MAX) && !getenv("KSS_AIVS_ONE_STREAM");
                cudaStream_t st_s = st;
                if (side) {
                    if (!ctx->mq_fork) CU(cudaEventCreateWithFlags(&ctx->mq_fork, cudaEventDisableTiming));
                    if (!ctx->mq_stream[0]) CU(cudaStreamCreateWithFlags(&ctx->mq_stream[0], cudaStreamNonBlocking));
                    if (!ctx->mq_done[0]) CU(cudaEventCreateWithFlags(&ctx->mq_done[0], cudaEventDisableTiming));
                    CU(cudaEventRecord(ctx->mq_fork, st));
                    CU(cudaStreamWaitEvent(ctx->mq_stream[0], ctx->mq_fork, 0));
                    st_s = ctx->mq_stream[0];
                }
                if (!q) q = aivs_simplify_device(st, &ctx->launches, P, full_t, c_T, bb.cap_T, d_pn, 0, d_st, bb.cap_t, d_ct, nullptr, d_bad, alloc, "t");
                if (!q) q = aivs_simplify_device(st_s, &ctx->launches, P, full_s, c_S, bb.cap_S, d_pn, 0, d_ss, bb.cap_s, d_cs, nullptr, d_bad, alloc, "s");
                if (side) {
                    CU(cudaEventRecord(ctx->mq_done[0], st_s));
                    CU(cudaStreamWaitEvent(st, ctx->mq_done[0], 0));
                }
                if (q) return fail(ctx, q, "AIVS simplification failed to launch");
                sim_s = d_ss; sim_t = d_st; c_s = d_cs; c_t = d_ct;
            } else {
                q = counts_or_fill(ctx, "cnt_s", c_s, P, bb.cap_s, &c_s); if (q) return q;
                q = counts_or_fill(ctx, "cnt_t", c_t, P, bb.cap_t, &c_t); if (q) return q;
            }
            // sizes of ragged full-resolution clouds are needed on the host when they take the large path
            const int *hS = nullptr, *hT = nullptr;
            std::vector<int> hostS, hostT;
            if (bb.cap_S > SMALL_MAX || bb.cap_T > SMALL_MAX) {
                if (host) { hS = bb.cnt_S ? bb.cnt_S + p0 : nullptr; hT = bb.cnt_T ? bb.cnt_T + p0 : nullptr; }
                else {
                    if (bb.cnt_S) { hostS.resize(P); CU(cudaMemcpyAsync(hostS.data(), bb.cnt_S + p0, sizeof(int) * P, cudaMemcpyDeviceToHost, st)); hS = hostS.data(); }
                    if (bb.cnt_T) { hostT.resize(P); CU(cudaMemcpyAsync(hostT.data(), bb.cnt_T + p0, sizeof(int) * P, cudaMemcpyDeviceToHost, st)); hT = hostT.data(); }
                    if (hS || hT) CU(cudaStreamSynchronize(st));
                }
                for (int i = 0; i < P; ++i)
                    if ((hS && (hS[i] < 1 || hS[i] > bb.cap_S)) || (hT && (hT[i] < 1 || hT[i] > bb.cap_T)))
                        return fail(ctx, KSS_ERR_ARG, "kss_batch: a full-resolution count is outside 1..capacity");
            }
            q = pipeline_device(ctx, P, bb, sim_s, sim_t, full_s, full_t, c_s, c_t, c_S, c_T, hS, hT, slots, d_res, d_pa, shard);
            if (q) return q;
            if (host) {
                CU(cudaMemcpyAsync(results + p0, d_res, sizeof(kss_pair_result) * (size_t)P, cudaMemcpyDeviceToHost, st));
                if (point_align)
                    CU(cudaMemcpyAsync(point_align + (size_t)p0 * bb.cap_S * 3, d_pa, sizeof(double) * 3 * (size_t)P * bb.cap_S, cudaMemcpyDeviceToHost, st));
            }
            return KSS_OK;
        }();
    }
    ctx->stream = main_st;
    ctx->buf_suffix.clear();
    for (int l = 0; l < lanes; ++l) {
        if (cudaEventRecord(ctx->lane_done[l], ctx->lane_stream[l]) != cudaSuccess ||
            cudaStreamWaitEvent(main_st, ctx->lane_done[l], 0) != cudaSuccess) { if (rc == KSS_OK) rc = fail(ctx, KSS_ERR_CUDA, "lane join failed"); }
    }
    return rc;
}

}  // namespace

int kss_register_batch_device(kss_ctx* ctx, const kss_batch* b, kss_pair_result* d_results, double* d_point_align) {
    int r = check_batch(ctx, b); if (r) return r;
    if (!d_results) return fail(ctx, KSS_ERR_ARG, "kss_register_batch_device: null results");
    CU(cudaSetDevice(ctx->device));
    return batch_core(ctx, b, false, d_results, d_point_align);
}

int kss_register_batch(kss_ctx* ctx, const kss_batch* b, kss_pair_result* results, double* point_align) {
    int r = check_batch(ctx, b); if (r) return r;
    if (!results) return fail(ctx, KSS_ERR_ARG, "kss_register_batch: null results");
    CU(cudaSetDevice(ctx->device));
    const int P = b->n_pairs;
    const bool raw = !b->sim_s;
    r = batch_core(ctx, b, true, results, point_align); if (r) return r;
    cudaStream_t st = ctx->stream;
    int bad = 0;
    if (raw) CU(cudaMemcpyAsync(&bad, ctx->aivs_bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (bad) return aivs_status(ctx, bad);
    return KSS_OK;
}

// ---- multi-GPU -----------------------------------------------------------------------
int kss_nccl_get_unique_id(void* id128) {
    if (!id128) return KSS_ERR_ARG;
    NcclApi* n = kss_nccl();
    if (!n) return KSS_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (n->GetUniqueId(&id) != ncclSuccess) return KSS_ERR_NCCL;
    std::memcpy(id128, &id, sizeof(id));
    return KSS_OK;
}

int kss_ctx_nccl_init(kss_ctx* ctx, const void* id128, int rank, int world) {
    if (!ctx) return KSS_ERR_ARG;
    if (!id128 || world < 1 || rank < 0 || rank >= world) return fail(ctx, KSS_ERR_ARG, "kss_ctx_nccl_init: bad argument");
    if (!kss_nccl()) return fail(ctx, KSS_ERR_NCCL, "libnccl.so.2 could not be loaded");
    CU(cudaSetDevice(ctx->device));
    if (ctx->comm) { kss_nccl()->CommDestroy(ctx->comm); ctx->comm = nullptr; }
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof(id));
    NC(kss_nccl()->CommInitRank(&ctx->comm, world, id, rank));
    ctx->rank = rank; ctx->world = world;
    return KSS_OK;
}

int kss_ctx_nccl_init_all(kss_ctx** ctxs, int n) {
    if (!ctxs || n < 1) return KSS_ERR_ARG;
    kss_ctx* ctx = ctxs[0];
    if (!ctx) return KSS_ERR_ARG;
    if (!kss_nccl()) return fail(ctx, KSS_ERR_NCCL, "libnccl.so.2 could not be loaded");
    std::vector<int> devs(n);
    std::vector<ncclComm_t> comms(n);
    for (int i = 0; i < n; ++i) { if (!ctxs[i]) return fail(ctx, KSS_ERR_ARG, "kss_ctx_nccl_init_all: null ctx"); devs[i] = ctxs[i]->device; }
    NC(kss_nccl()->CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) {
        if (ctxs[i]->comm) kss_nccl()->CommDestroy(ctxs[i]->comm);
        ctxs[i]->comm = comms[i]; ctxs[i]->rank = i; ctxs[i]->world = n;
    }
    return KSS_OK;
}

int kss_register_batch_hyp_sharded(kss_ctx* ctx, const kss_batch* b, kss_pair_result* results, double* point_align) {
    int r = check_batch(ctx, b); if (r) return r;
    if (!results) return fail(ctx, KSS_ERR_ARG, "kss_register_batch_hyp_sharded: null results");
    if (ctx->world > 1 && !ctx->comm) return fail(ctx, KSS_ERR_NCCL, "kss_register_batch_hyp_sharded: call kss_ctx_nccl_init first");
    CU(cudaSetDevice(ctx->device));
    const bool raw = !b->sim_s;
    r = batch_core(ctx, b, true, results, point_align, true); if (r) return r;
    int bad = 0;
    if (raw) CU(cudaMemcpyAsync(&bad, ctx->aivs_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (bad) return aivs_status(ctx, bad);
    return KSS_OK;
}

int kss_register_batch_multi(kss_ctx** ctxs, int n_ctx, const kss_batch* b, kss_pair_result* results, double* point_align) {
    if (!ctxs || n_ctx < 1 || !ctxs[0]) return KSS_ERR_ARG;
    int r = check_batch(ctxs[0], b); if (r) return r;
    if (!results) return fail(ctxs[0], KSS_ERR_ARG, "kss_register_batch_multi: null results");
    // pairs are independent (Main_KSS_List.cpp:132-167 is a plain loop): contiguous blocks, one host thread per context, no
    // data-path collective
    const int NP = b->n_pairs, per = (NP + n_ctx - 1) / n_ctx;
    std::vector<int> rc(n_ctx, KSS_OK);
    std::vector<std::thread> th;
    for (int i = 0; i < n_ctx; ++i) {
        const int p0 = std::min(NP, i * per), p1 = std::min(NP, p0 + per);
        if (p1 <= p0) continue;
        th.emplace_back([=, &rc]() {
            kss_batch one = *b;
            one.n_pairs = p1 - p0;
            one.full_s = b->full_s + (size_t)p0 * b->cap_S * 3; one.full_t = b->full_t + (size_t)p0 * b->cap_T * 3;
            if (b->sim_s) { one.sim_s = b->sim_s + (size_t)p0 * b->cap_s * 3; one.sim_t = b->sim_t + (size_t)p0 * b->cap_t * 3; }
            one.cnt_s = b->cnt_s ? b->cnt_s + p0 : nullptr; one.cnt_t = b->cnt_t ? b->cnt_t + p0 : nullptr;
            one.cnt_S = b->cnt_S ? b->cnt_S + p0 : nullptr; one.cnt_T = b->cnt_T ? b->cnt_T + p0 : nullptr;
            rc[i] = kss_register_batch(ctxs[i], &one, results + p0, point_align ? point_align + (size_t)p0 * b->cap_S * 3 : nullptr);
        });
    }
    for (auto& t : th) t.join();
    for (int i = 0; i < n_ctx; ++i)
        if (rc[i] != KSS_OK) { if (i) ctxs[0]->err = "context " + std::to_string(i) + ": " + ctxs[i]->err; return rc[i]; }
    return KSS_OK;
}

int kss_aivs_simplify_batch_device(kss_ctx* ctx, int n_clouds, const double* d_pts, const int* d_cnt, int cap,
                                   const int* d_point_num, int point_num, double* d_out, int out_cap, int* d_out_cnt,
                                   int32_t* d_out_idx) {
    if (!ctx) return KSS_ERR_ARG;
    if (n_clouds < 1 || !d_pts || cap < 1 || !d_out || !d_out_cnt || out_cap < 1 || (!d_point_num && point_num < 1))
        return fail(ctx, KSS_ERR_ARG, "kss_aivs_simplify_batch_device: bad argument");
    CU(cudaSetDevice(ctx->device));
    auto alloc = [&](const char* name, size_t bytes, void** out) {
        unsigned char* q; int rr = dev_buf(ctx, name, bytes, &q); *out = q; return rr; };
    int* d_bad;
    BUF("aivs_bad", 1, &d_bad);
    CU(cudaMemsetAsync(d_bad, 0, sizeof(int), ctx->stream));
    ctx->aivs_bad = d_bad;
    StageTimer tm(ctx, KSS_STAGE_AIVS);
    int r = aivs_simplify_device(ctx->stream, &ctx->launches, n_clouds, d_pts, d_cnt, cap, d_point_num, point_num, d_out,
                                 out_cap, d_out_cnt, d_out_idx, d_bad, alloc, "u");
    if (r) return fail(ctx, r, "AIVS simplification failed to launch");
    return KSS_OK;
}

// diagnostics: copy part of a named internal device buffer to the host (tools/ only; names are not a stable interface)
int kss_debug_read(kss_ctx* ctx, const char* name, size_t offset, size_t bytes, void* dst) {
    if (!ctx || !name || !dst) return KSS_ERR_ARG;
    auto it = ctx->bufs.find(name);
    if (it == ctx->bufs.end()) it = ctx->bufs.find(std::string(name) + "#0");       // batch buffers live per lane
    if (it == ctx->bufs.end() || !it->second.p || offset + bytes > it->second.cap) return fail(ctx, KSS_ERR_ARG, "kss_debug_read: no such buffer / range");
    CU(cudaMemcpyAsync(dst, (const unsigned char*)it->second.p + offset, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return KSS_OK;
}

int kss_aivs_status(kss_ctx* ctx) {
    if (!ctx) return KSS_ERR_ARG;
    if (!ctx->aivs_bad) return KSS_OK;
    int bad = 0;
    CU(cudaMemcpyAsync(&bad, ctx->aivs_bad, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return aivs_status(ctx, bad);
}

int kss_aivs_simplify_batch(kss_ctx* ctx, int n_clouds, const double* pts, const int* cnt, int cap, const int* point_num_each,
                            int point_num, double* out, int out_cap, int* out_cnt, int32_t* out_idx) {
    if (!ctx) return KSS_ERR_ARG;
    if (n_clouds < 1 || !pts || cap < 1 || !out || !out_cnt || out_cap < 1 || (!point_num_each && point_num < 1))
        return fail(ctx, KSS_ERR_ARG, "kss_aivs_simplify_batch: bad argument");
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t P = (size_t)n_clouds;
    double *d_pts, *d_out; int *d_cnt = nullptr, *d_pn = nullptr, *d_ocnt, *d_oidx = nullptr;
    BUF("aivs_in", P * cap * 3, &d_pts); BUF("aivs_out", P * out_cap * 3, &d_out); BUF("aivs_ocnt", P, &d_ocnt);
    if (out_idx) BUF("aivs_oidx", P * out_cap, &d_oidx);
    CU(cudaMemcpyAsync(d_pts, pts, sizeof(double) * 3 * P * cap, cudaMemcpyHostToDevice, st));
    if (cnt) { BUF("aivs_icnt", P, &d_cnt); CU(cudaMemcpyAsync(d_cnt, cnt, sizeof(int) * P, cudaMemcpyHostToDevice, st)); }
    if (point_num_each) { BUF("aivs_ipn", P, &d_pn); CU(cudaMemcpyAsync(d_pn, point_num_each, sizeof(int) * P, cudaMemcpyHostToDevice, st)); }
    int r = kss_aivs_simplify_batch_device(ctx, n_clouds, d_pts, d_cnt, cap, d_pn, point_num, d_out, out_cap, d_ocnt, d_oidx);
    if (r) return r;
    CU(cudaMemcpyAsync(out, d_out, sizeof(double) * 3 * P * out_cap, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(out_cnt, d_ocnt, sizeof(int) * P, cudaMemcpyDeviceToHost, st));
    if (out_idx) CU(cudaMemcpyAsync(out_idx, d_oidx, sizeof(int) * P * out_cap, cudaMemcpyDeviceToHost, st));
    return kss_aivs_status(ctx);
}

int kss_aivs_simplify(kss_ctx* ctx, const double* pts, int n, int point_num, double* out, int out_cap, int* out_n,
                      int32_t* out_idx) {
    if (!ctx) return KSS_ERR_ARG;
    if (!out_n) return fail(ctx, KSS_ERR_ARG, "kss_aivs_simplify: null out_n");
    return kss_aivs_simplify_batch(ctx, 1, pts, nullptr, n, nullptr, point_num, out, out_cap, out_n, out_idx);
}

int kss_register(kss_ctx* ctx, const double* sim_s, int n_s, const double* sim_t, int n_t, const double* full_s,
                 int N_s, const double* full_t, int N_t, double step, int max_iter, kss_pair_result* result,
                 double* point_align) {
    kss_batch b; kss_batch_default(&b);
    b.n_pairs = 1; b.cap_s = n_s; b.cap_t = n_t; b.cap_S = N_s; b.cap_T = N_t;
    b.sim_s = sim_s; b.sim_t = sim_t; b.full_s = full_s; b.full_t = full_t;
    b.step = step; b.icp.max_iterations = max_iter;
    return kss_register_batch(ctx, &b, result, point_align);
}

}  // extern "C"
