// bayesn_capi.cu - extern "C" boundary of libbayesn_b200.so (see include/bayesn_b200.h).
#include <algorithm>
#include <type_traits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bayesn_kernels.cuh"
#include "bayesn_train.cuh"
#include "bayesn_train_nuts.cuh"
#include "bayesn_tiles.cuh"

using namespace bayesn;

struct bayesn_ctx {
    int device = 0;
    std::string err;
    bool have_model = false, have_data = false;
    int LP = 0, TP = 0, VPL = 0;
    ModelDev M{};
    DataDev D{};
    std::vector<void*> owned;   // device allocations made by set_model
    long long launches = 0;
    TrainDev T{};               // training buffers (allocated on first use for a given (chains, N))
    std::vector<void*> train_owned;
    int train_C = 0, train_N = 0;
    TrainNutsDev TN{};          // training sampler state
    XchgDev X{};                // peer-memory exchange of the training sums (world 1 = off)
    std::vector<void*> tn_owned;
    bool have_tn = false;
    int train_kw = 1;           // warps per (SN, chain) unit of the last training SN launch
    bool fuse_ctl = false;      // training pass: second reduction stage (+ peer exchange) inside the control kernel
};

namespace {

int fail(bayesn_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    return code;
}

#define CU_CHECK(ctx, expr)                                                                       \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return fail(ctx, -2, std::string(#expr) + ": " + cudaGetErrorString(_e));             \
    } while (0)

template <typename T>
int upload(bayesn_ctx* ctx, const std::vector<T>& h, const T** out) {
    void* d = nullptr;
    CU_CHECK(ctx, cudaMalloc(&d, h.size() * sizeof(T)));
    ctx->owned.push_back(d);
    CU_CHECK(ctx, cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = static_cast<const T*>(d);
    return 0;
}

void free_train(bayesn_ctx* ctx) {
    for (void* p : ctx->train_owned) cudaFree(p);
    ctx->train_owned.clear();
    ctx->train_C = ctx->train_N = 0;
    for (void* p : ctx->tn_owned) cudaFree(p);
    ctx->tn_owned.clear();
    ctx->have_tn = false;
}

void free_owned(bayesn_ctx* ctx) {
    for (void* p : ctx->owned) cudaFree(p);
    ctx->owned.clear();
    free_train(ctx);
}

template <typename T>
int train_alloc(bayesn_ctx* ctx, T** out, size_t n) {
    void* d = nullptr;
    CU_CHECK(ctx, cudaMalloc(&d, n * sizeof(T)));
    CU_CHECK(ctx, cudaMemset(d, 0, n * sizeof(T)));
    ctx->train_owned.push_back(d);
    *out = static_cast<T*>(d);
    return 0;
}

int train_prepare(bayesn_ctx* ctx, int C) {
    if (ctx->train_C == C && ctx->train_N == ctx->D.N) return 0;
    free_train(ctx);
    TrainDev& T = ctx->T;
    const int ne = ctx->M.n_eps, L = ctx->M.L, Tn = ctx->M.T;
    T = TrainDev{};
    T.C = C;
    T.n_corr = ne * (ne - 1) / 2;
    T.G = 2 * L * Tn + ne + T.n_corr + 3;
    T.Dl = ne + 3;
    T.LPTP = ctx->LP * ctx->TP;
    T.PR = train_part_stride(T.LPTP);
    T.R = train_red_len(T.LPTP, ne);
    int rc = 0;
    rc |= train_alloc(ctx, &T.W0, (size_t)C * T.LPTP);
    rc |= train_alloc(ctx, &T.W1, (size_t)C * T.LPTP);
    rc |= train_alloc(ctx, &T.Ls, (size_t)C * ne * ne);
    rc |= train_alloc(ctx, &T.LsT, (size_t)C * ne * ne);
    rc |= train_alloc(ctx, &T.ad, (size_t)C * JROWS + OVERRUN);
    rc |= train_alloc(ctx, &T.dad, (size_t)C * JROWS + OVERRUN);
    rc |= train_alloc(ctx, &T.sc, (size_t)C * 8);
    rc |= train_alloc(ctx, &T.sig, (size_t)C * ne);
    rc |= train_alloc(ctx, &T.Lo, (size_t)C * ne * ne);
    rc |= train_alloc(ctx, &T.tm, (size_t)C * ne * ne);
    rc |= train_alloc(ctx, &T.cum, (size_t)C * ne * ne);
    rc |= train_alloc(ctx, &T.clog, (size_t)C * ne * ne);
    rc |= train_alloc(ctx, &T.scr, (size_t)C * ne * ne);
    rc |= train_alloc(ctx, &T.lpg, (size_t)C);
    rc |= train_alloc(ctx, &T.part, (size_t)ctx->D.N * C * T.PR);
    rc |= train_alloc(ctx, &T.rpart, (size_t)((ctx->D.N + TRAIN_RCHUNK - 1) / TRAIN_RCHUNK) * C * T.R);
    if (rc) return rc;
    ctx->train_C = C;
    ctx->train_N = ctx->D.N;
    return 0;
}

struct Combo { int LP, TP, VPL; };
const Combo kCombos[] = {{6, 6, 1}, {9, 6, 2}, {11, 6, 2}, {10, 9, 3}, {12, 10, 4}};

// capacity (number of distinct SNe) of a CTA of upc consecutive (SN, chain) units (upc = WPB / warps per unit)
int ntile_for(int C, int upc = WPB) { return (C % upc == 0) ? 1 : ((upc % C == 0) ? upc / C : (upc + C - 1) / C + 1); }

int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// the model descriptor of one launch in latency mode: kw warps per (SN, chain) unit, where the teammates'
// exchange slots are (bayesn_device.cuh: ModelDev::kw)
ModelDev team_model(const ModelDev& M0, int kw, const SmemPlan& P) {
    ModelDev M = M0;
    M.kw = kw; M.kwl = ilog2(kw); M.xstride = P.warpStride; M.xoff = P.wX - P.wSc;
    return M;
}

template <int LP, int TP, int VPL, bool RVFREE, bool STAGE>
int launch_logp(bayesn_ctx* ctx, int C, int kw, const float* q, float* logp, float* grad, cudaStream_t st) {
    const int upc = WPB / kw;
    const long long units = (long long)ctx->D.N * C;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    // Big batches (more than four waves of CTAs) run the persistent form: tables staged once per CTA, every warp strides
    // over the units with its own slot of SN data; BAYESN_LOGP_PERSIST=0|1 overrides (measurements, tests).  Small
    // batches and latency-mode launches keep one CTA per eight units.
    bool persist = kw == 1 && units > 4LL * 2 * sms * WPB;
    if (const char* e = getenv("BAYESN_LOGP_PERSIST")) persist = kw == 1 && atoi(e) != 0;
    const int ntile = persist ? WPB : ntile_for(C, upc);
    // one evaluation per warp: staging L_Sigma in shared memory (12 KB per CTA) costs more than the L1 walks it saves;
    // the persistent form and latency-mode launches take the sampler's shared-memory path
    const bool lsmem = kw > 1 || persist;
    const SmemPlan P = make_plan(LP, TP, ctx->M.n_eps, ctx->D.n_b, ctx->D.N_obs, ctx->D.wt_stride, ntile, RVFREE, STAGE, false, kw, lsmem);
    const size_t bytes = (size_t)P.totalFloats * 4;
    auto kern = lsmem ? logp_grad_kernel<LP, TP, VPL, RVFREE, STAGE, true> : logp_grad_kernel<LP, TP, VPL, RVFREE, STAGE, false>;
    if (bytes > 227 * 1024) return fail(ctx, -3, "dataset needs more shared memory per CTA than sm_100a has");
    CU_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    unsigned grid = (unsigned)((units + upc - 1) / upc);
    if (persist) {
        int occ = 0;
        CU_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WPB * 32, bytes));
        if (occ < 1) return fail(ctx, -3, "persistent log p kernel does not fit");
        grid = (unsigned)std::min<long long>(grid, (long long)occ * sms);
    }
    kern<<<grid, WPB * 32, bytes, st>>>(team_model(ctx->M, kw, P), ctx->D, C, ntile, q, logp, grad, persist ? 1 : 0);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}

template <int LP, int TP, int VPL, bool RVFREE, bool STAGE>
int launch_nuts(bayesn_ctx* ctx, const NutsDev& cfg, int kw, unsigned long long* queue_slot, const float* q0, float* samples,
                float* stats, float* info, float* work, cudaStream_t st) {
    const bool use_queue = queue_slot != nullptr;
    const int C = cfg.num_chains;
    const int upc = WPB / kw;
    const int ntile = use_queue ? upc : ntile_for(C, upc);      // queue: one slot of SN data per team
    const SmemPlan P = make_plan(LP, TP, ctx->M.n_eps, ctx->D.n_b, ctx->D.N_obs, ctx->D.wt_stride, ntile, RVFREE, STAGE, false, kw, true);
    const size_t bytes = (size_t)P.totalFloats * 4;
    auto kern = nuts_kernel<LP, TP, VPL, RVFREE, STAGE>;
    if (bytes > 227 * 1024) return fail(ctx, -3, "dataset needs more shared memory per CTA than sm_100a has");
    CU_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    const long long units = (long long)ctx->D.N * C;
    unsigned grid = (unsigned)((units + upc - 1) / upc);
    unsigned long long* queue = nullptr;
    if (use_queue) {
        int occ = 0, sms = 0;
        CU_CHECK(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
        CU_CHECK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, WPB * 32, bytes));
        if (occ < 1) return fail(ctx, -3, "work-queue schedule does not fit");
        CU_CHECK(ctx, cudaMemsetAsync(queue_slot, 0, sizeof(unsigned long long), st));
        queue = queue_slot;
        grid = (unsigned)std::min<long long>(grid, (long long)occ * sms);      // one resident wave
    }
    kern<<<grid, WPB * 32, bytes, st>>>(team_model(ctx->M, kw, P), ctx->D, cfg, ntile, q0, samples, stats, info, work, queue);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}

// Schedule of the fit sampler.  The work queue (one resident wave of CTAs whose warps fetch (SN, chain)
// units) needs a slot of SN data per warp instead of per CTA; it is used when that costs no occupancy
// against the static schedule and there is more than one wave of work.  When the staged slots do not fit
// (C5: 60 observations, 4.9 KB tiles) the static schedule stays: reading the tiles through L1 instead
// was measured slower than the imbalance it removes (87 vs 93 SNe/s, profiles/README.md).
// BAYESN_NUTS_SCHED=static|queue|queue-l1 overrides (measurements, tests).
void choose_nuts_schedule(bayesn_ctx* ctx, int C, int kw, bool* use_queue, bool* stage) {
    const bool rvfree = ctx->M.rv_mode != BAYESN_RV_FIXED;
    const int upc = WPB / kw;
    auto occ = [&](int ntile, bool st) {
        const SmemPlan P = make_plan(ctx->LP, ctx->TP, ctx->M.n_eps, ctx->D.n_b, ctx->D.N_obs, ctx->D.wt_stride, ntile, rvfree, st, false, kw, true);
        const long long bytes = (long long)P.totalFloats * 4;
        if (bytes > 227 * 1024) return 0;
        return (int)std::min<long long>(16 / WPB, (228LL * 1024) / (bytes + 1024));    // 128 registers x 16 warps fill the SM
    };
    const bool stage_static = C >= 2;       // one chain per SN: nothing to share, the tile is read through L1
    const int occ_static = std::max(occ(ntile_for(C, upc), stage_static), 1);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const long long ctas = ((long long)ctx->D.N * C + upc - 1) / upc;
    *use_queue = false;
    *stage = stage_static;
    const char* e = getenv("BAYESN_NUTS_SCHED");
    if (e && !strcmp(e, "static")) return;
    if (e && !strcmp(e, "queue-l1")) { *use_queue = occ(upc, false) > 0; if (*use_queue) *stage = false; return; }
    if (e && !strcmp(e, "queue")) { *use_queue = occ(upc, stage_static) > 0; return; }
    *use_queue = ctas > (long long)occ_static * sms && occ(upc, stage_static) >= occ_static;
}

// compile-time dispatch over the supported (knots, parameter-vector) shapes
template <typename F>
int dispatch(bayesn_ctx* ctx, bool stage, F&& f) {
    const bool rvfree = ctx->M.rv_mode != BAYESN_RV_FIXED;
#define BAYESN_CASE(LPv, TPv, VPLv)                                                                   \
    if (ctx->LP == LPv && ctx->TP == TPv && ctx->VPL == VPLv) {                                       \
        if (rvfree) return stage ? f.template operator()<LPv, TPv, VPLv, true, true>()               \
                                 : f.template operator()<LPv, TPv, VPLv, true, false>();             \
        return stage ? f.template operator()<LPv, TPv, VPLv, false, true>()                           \
                     : f.template operator()<LPv, TPv, VPLv, false, false>();                         \
    }
    BAYESN_CASE(6, 6, 1)
    BAYESN_CASE(9, 6, 2)
    BAYESN_CASE(11, 6, 2)
    BAYESN_CASE(10, 9, 3)
    BAYESN_CASE(12, 10, 4)
#undef BAYESN_CASE
    return fail(ctx, -4, "no kernel instantiation for this model shape (compiled: L<=6,T<=6,D<=32 | L<=9,T<=6,D<=64 | "
                         "L<=11,T<=6,D<=64 | L<=10,T<=9,D<=96 | L<=12,T<=10,D<=128)");
}

template <int LP, int TP, int VPL, bool RVFREE, bool STAGE>
int launch_train_sn(bayesn_ctx* ctx, const float* ql, float* grad_l, cudaStream_t st) {
    if (RVFREE) return fail(ctx, -1, "training needs a fixed-RV model context (rv_mode = BAYESN_RV_FIXED)");
    const int C = ctx->T.C;
    const int ntile = ntile_for(C);
    const SmemPlan P = make_plan(LP, TP, ctx->M.n_eps, ctx->D.n_b, ctx->D.N_obs, ctx->D.wt_stride, ntile, false, STAGE, true);
    const size_t bytes = (size_t)P.totalFloats * 4;
    auto kern = train_sn_kernel<LP, TP, (LP * TP - 2 * TP + 3 + 31) / 32, STAGE>;
    if (bytes > 227 * 1024) return fail(ctx, -3, "dataset needs more shared memory per CTA than sm_100a has");
    CU_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    const long long units = (long long)ctx->D.N * C;
    const unsigned grid = (unsigned)((units + WPB - 1) / WPB);
    kern<<<grid, WPB * 32, bytes, st>>>(ctx->M, ctx->D, ctx->T, ntile, ql, grad_l);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}

// two-stage deterministic reduction of the per-SN rows (bayesn_train.cuh T3)
int launch_train_reduce(bayesn_ctx* ctx, const float* lat, size_t chain_stride, size_t sn_stride, double* red,
                        cudaStream_t st, bool final_stage = true) {
    const TrainDev& T = ctx->T;
    const int nchunk = (ctx->D.N + TRAIN_RCHUNK - 1) / TRAIN_RCHUNK;
    const size_t smem = (size_t)TRAIN_RCHUNK * (T.PR + ctx->M.n_eps + 2) * sizeof(float);
    CU_CHECK(ctx, cudaFuncSetAttribute(train_reduce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    train_reduce_kernel<<<dim3(nchunk, T.C, (T.R + 255) / 256), 256, smem, st>>>(ctx->M, T, ctx->D.N, ctx->LP, ctx->TP, lat, chain_stride,
                                                              sn_stride, T.rpart);
    if (final_stage) train_reduce_final_kernel<<<dim3((T.R + 255) / 256, T.C), 256, 0, st>>>(T, nchunk, T.rpart, red, ctx->X);
    ctx->launches += final_stage ? 2 : 1;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}

struct TrainSnCall {
    bayesn_ctx* ctx; const float* ql; float* grad_l; cudaStream_t st;
    template <int LP, int TP, int VPL, bool RV, bool ST>
    int operator()() { return launch_train_sn<LP, TP, VPL, RV, ST>(ctx, ql, grad_l, st); }
};

template <int LP, int TP, int VPL, bool RVFREE, bool STAGE>
int launch_train_nuts_sn(bayesn_ctx* ctx, cudaStream_t st) {
    if (RVFREE) return fail(ctx, -1, "training needs a fixed-RV model context (rv_mode = BAYESN_RV_FIXED)");
    const int C = ctx->T.C;
    const long long units = (long long)ctx->D.N * C;
    // latency mode: when this rank's (SN, chain) units leave warp slots idle (SNe sharded over GPUs), kw warps
    // share the observation pairs of each unit's one evaluation per pass; BAYESN_TRAIN_KW overrides (tests)
    int kw = 1, sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    for (int k = WPB; k >= 2; k >>= 1)
        if (units * k <= 16LL * sms) { kw = k; break; }
    if (const char* e = getenv("BAYESN_TRAIN_KW")) {
        const int v = atoi(e);
        if (v >= 1 && v <= WPB && !(v & (v - 1))) kw = v;
    }
    ctx->train_kw = kw;
    const int upc = WPB / kw;
    const int ntile = ntile_for(C, upc);
    const SmemPlan P = make_plan(LP, TP, ctx->M.n_eps, ctx->D.n_b, ctx->D.N_obs, ctx->D.wt_stride, ntile, false, STAGE, true, kw);
    const size_t bytes = (size_t)P.totalFloats * 4;
    constexpr int TVPL = (LP * TP - 2 * TP + 3 + 31) / 32;
    auto kern = train_nuts_sn_kernel<LP, TP, TVPL, STAGE>;
    if (bytes > 227 * 1024) return fail(ctx, -3, "dataset needs more shared memory per CTA than sm_100a has");
    if (ctx->TN.VPL != TVPL) return fail(ctx, -1, "internal: latent vector width mismatch");
    CU_CHECK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    const unsigned grid = (unsigned)((units + upc - 1) / upc);
    kern<<<grid, WPB * 32, bytes, st>>>(team_model(ctx->M, kw, P), ctx->D, ctx->T, ctx->TN, ntile);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}

struct TrainNutsSnCall {
    bayesn_ctx* ctx; cudaStream_t st;
    template <int LP, int TP, int VPL, bool RV, bool ST>
    int operator()() { return launch_train_nuts_sn<LP, TP, VPL, RV, ST>(ctx, st); }
};

struct LogpCall {
    bayesn_ctx* ctx; int C; int kw; const float* q; float* logp; float* grad; cudaStream_t st;
    template <int LP, int TP, int VPL, bool RV, bool ST>
    int operator()() { return launch_logp<LP, TP, VPL, RV, ST>(ctx, C, kw, q, logp, grad, st); }
};
struct NutsCall {
    bayesn_ctx* ctx; NutsDev cfg; int kw; unsigned long long* queue_slot; const float* q0; float* samples; float* stats; float* info; float* work;
    cudaStream_t st;
    template <int LP, int TP, int VPL, bool RV, bool ST>
    int operator()() { return launch_nuts<LP, TP, VPL, RV, ST>(ctx, cfg, kw, queue_slot, q0, samples, stats, info, work, st); }
};

// warps per (SN, chain) unit: 0 / 1 = throughput layout, else a power of two that divides the CTA
int team_width(bayesn_ctx* ctx, int w, int* kw) {
    if (w <= 1) { *kw = 1; return 0; }
    if ((w & (w - 1)) || w > WPB) return fail(ctx, -1, "warps_per_chain must be 1, 2, 4 or 8");
    *kw = w;
    return 0;
}

// numpyro.infer.hmc_util.build_adaptation_schedule (Stan windows)
std::vector<int> adaptation_window_ends(int num_steps) {
    std::vector<int> ends;
    if (num_steps <= 0) return ends;
    if (num_steps < 20) { ends.push_back(num_steps - 1); return ends; }
    int start_buffer = 75, end_buffer = 50, init_window = 25;
    if (start_buffer + end_buffer + init_window > num_steps) {
        start_buffer = (int)(0.15 * num_steps);
        end_buffer = (int)(0.1 * num_steps);
        init_window = num_steps - start_buffer - end_buffer;
    }
    ends.push_back(start_buffer - 1);
    const int end_window_start = num_steps - end_buffer;
    int next_size = init_window, next_start = start_buffer;
    while (next_start < end_window_start) {
        int cur_start = next_start, cur_size = next_size;
        if (3 * cur_size <= end_window_start - cur_start) next_size = 2 * cur_size;
        else cur_size = end_window_start - cur_start;
        next_start = cur_start + cur_size;
        ends.push_back(next_start - 1);
    }
    ends.push_back(num_steps - 1);
    return ends;
}

}  // namespace

extern "C" {

int bayesn_abi_version(void) { return 2; }

int bayesn_ctx_create(int device, bayesn_ctx** out) {
    if (!out) return -1;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || device < 0 || device >= n) return -2;
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    bayesn_ctx* c = new bayesn_ctx();
    c->device = device;
    *out = c;
    return 0;
}

void bayesn_ctx_destroy(bayesn_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    free_owned(ctx);
    delete ctx;
}

const char* bayesn_last_error(bayesn_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

long long bayesn_launch_count(bayesn_ctx* ctx) { return ctx ? ctx->launches : 0; }

int bayesn_param_dim(bayesn_ctx* ctx) { return (ctx && ctx->have_model) ? ctx->M.D : -1; }

int bayesn_train_warps_per_unit(bayesn_ctx* ctx) { return ctx ? ctx->train_kw : -1; }

int bayesn_train_fuse_reduction(bayesn_ctx* ctx, int on) {
    if (!ctx) return -1;
    ctx->fuse_ctl = on != 0;
    return 0;
}

int bayesn_set_model(bayesn_ctx* ctx, const bayesn_model_desc* m) {
    if (!ctx || !m) return -1;
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    const int L = m->L, T = m->T;
    if (L < 3 || L > BAYESN_MAX_L || T < 2 || T > BAYESN_MAX_T) return fail(ctx, -1, "unsupported knot counts");
    const int ne = (L - 2) * T;
    const int n_sc = (m->rv_mode == BAYESN_RV_FIXED) ? 4 : 5;
    const int D = ne + n_sc;
    ctx->LP = 0;
    for (const Combo& c : kCombos)
        if (c.LP >= L && c.TP >= T && 32 * c.VPL >= D) { ctx->LP = c.LP; ctx->TP = c.TP; ctx->VPL = c.VPL; break; }
    if (!ctx->LP)
        return fail(ctx, -4, "model shape not supported: " + std::to_string(L) + " wavelength x " + std::to_string(T) +
                                 " time knots (compiled: L<=6,T<=6 | L<=9,T<=6 | L<=11,T<=6 | L<=10,T<=9 | L<=12,T<=10)");
    free_owned(ctx);
    ctx->have_model = false;
    const int LP = ctx->LP, TP = ctx->TP;
    const int RS = jl_row_stride(LP);
    std::vector<float> Jl((size_t)JROWS * RS, 0.f), hs((size_t)(NPH + 1) * BP + OVERRUN + 32, 0.f), Mf((size_t)9 * BP, 0.f), ad(JROWS, 0.f);
    std::vector<float> W0((size_t)LP * TP, 0.f), W1((size_t)LP * TP, 0.f), KD((size_t)TP * TP, 0.f), tau(TP, 0.f);
    const int lss = ls_row_stride(LP, TP);       // >= ne; compile-time row stride of the evaluator
    std::vector<float> Ls((size_t)ne * lss, 0.f), LsT((size_t)ne * lss, 0.f);
    for (int b = 0; b < NB; ++b) {
        for (int l = 0; l < L; ++l) Jl[(size_t)b * RS + l] = m->J_l[(size_t)b * L + l];
        for (int i = 0; i < NPH; ++i) hs[(size_t)i * BP + b] = m->hsiao[(size_t)b * NPH + i];
        for (int j = 0; j < 9; ++j) Mf[(size_t)j * BP + b] = m->M_fitz[(size_t)b * 9 + j];
        if (m->a_dust) ad[b] = m->a_dust[b];
    }
    for (int l = 0; l < L; ++l)
        for (int t = 0; t < T; ++t) {
            W0[(size_t)l * TP + t] = m->W0[(size_t)l * T + t];
            W1[(size_t)l * TP + t] = m->W1[(size_t)l * T + t];
        }
    for (int a = 0; a < T; ++a) {
        tau[a] = m->tau_knots[a];
        for (int b = 0; b < T; ++b) KD[(size_t)a * TP + b] = m->KD_t[(size_t)a * T + b];
    }
    bool lower = true;
    for (int i = 0; i < ne; ++i)
        for (int j = 0; j < ne; ++j) {
            Ls[(size_t)i * lss + j] = m->L_Sigma[(size_t)i * ne + j];
            LsT[(size_t)j * lss + i] = m->L_Sigma[(size_t)i * ne + j];
            if (j > i && m->L_Sigma[(size_t)i * ne + j] != 0.f) lower = false;
        }
    ModelDev& M = ctx->M;
    M = ModelDev{};
    M.kw = 1; M.kwl = 0; M.xstride = 0; M.xoff = 0;
    M.L = L; M.T = T; M.n_eps = ne; M.D = D; M.n_sc = n_sc; M.rv_mode = m->rv_mode;
    M.ls_lower = lower ? 1 : 0;
    M.fix_tmax = m->fix_tmax; M.fix_theta = m->fix_theta; M.fix_AV = m->fix_AV;
    M.theta_val = m->theta_val; M.AV_val = m->AV_val;
    M.inv_tauA = 1.f / m->tauA;
    M.inv_sd2 = 1.f / (m->Ds_prior_sd * m->Ds_prior_sd);
    M.RV = m->RV; M.mu_R = m->mu_R; M.sigma_R = m->sigma_R;
    M.phi_alpha = (m->rv_mode == BAYESN_RV_POP)
                      ? (float)(0.5 * erfc(-((double)m->trunc_R - m->mu_R) / m->sigma_R / sqrt(2.0))) : 0.f;
    int rc = 0;
    rc |= upload(ctx, Jl, &M.Jl);   rc |= upload(ctx, hs, &M.hs);   rc |= upload(ctx, Mf, &M.Mf);
    rc |= upload(ctx, ad, &M.ad);   rc |= upload(ctx, W0, &M.W0);   rc |= upload(ctx, W1, &M.W1);
    rc |= upload(ctx, KD, &M.KD);   rc |= upload(ctx, tau, &M.tau); rc |= upload(ctx, Ls, &M.Ls);
    rc |= upload(ctx, LsT, &M.LsT);
    if (rc) return rc;
    ctx->have_model = true;
    return 0;
}

int bayesn_bind_dataset(bayesn_ctx* ctx, const bayesn_dataset_desc* d) {
    if (!ctx || !d) return -1;
    if (d->N <= 0 || d->N_obs <= 0 || d->n_b <= 0) return fail(ctx, -1, "empty dataset");
    if (d->wt_stride <= 0 || (d->wt_stride % 4) || ((uintptr_t)d->obs4 & 15) || ((uintptr_t)d->weights & 15) ||
        ((uintptr_t)d->support & 15))
        return fail(ctx, -1, "dataset arrays must be 16-byte aligned and wt_stride a positive multiple of 4");
    DataDev& D = ctx->D;
    D.N = d->N; D.N_obs = d->N_obs; D.n_b = d->n_b; D.wt_stride = d->wt_stride;
    D.obs4 = reinterpret_cast<const float4*>(d->obs4);
    D.n_valid = d->n_valid;
    D.wt = d->weights;
    D.sup = reinterpret_cast<const int4*>(d->support);
    D.muhat = d->muhat;
    D.lconst = d->lconst;
    D.muerr2 = d->muhat_err2;
    ctx->have_data = true;
    return 0;
}

int bayesn_logp_grad_fit_team(bayesn_ctx* ctx, int n_chains, int warps_per_chain, const float* q, float* logp,
                              float* grad, void* stream) {
    if (!ctx) return -1;
    if (!ctx->have_model || !ctx->have_data) return fail(ctx, -1, "set_model and bind_dataset first");
    if (n_chains < 1) return fail(ctx, -1, "n_chains must be >= 1");
    int kw = 1;
    if (int rc = team_width(ctx, warps_per_chain, &kw)) return rc;
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    LogpCall call{ctx, n_chains, kw, q, logp, grad, (cudaStream_t)stream};
    return dispatch(ctx, /*stage=*/n_chains >= 2, call);
}

int bayesn_logp_grad_fit(bayesn_ctx* ctx, int n_chains, const float* q, float* logp, float* grad, void* stream) {
    return bayesn_logp_grad_fit_team(ctx, n_chains, 1, q, logp, grad, stream);
}

static size_t nuts_vector_bytes(bayesn_ctx* ctx, const bayesn_nuts_cfg* cfg) {
    // every warp of a team keeps its own copy of the tree vectors (latency mode)
    const size_t kw = cfg->warps_per_chain > 1 ? (size_t)cfg->warps_per_chain : 1;
    const size_t b = (size_t)ctx->D.N * cfg->num_chains * kw * nuts_num_vectors(cfg->max_tree_depth) * 32 * ctx->VPL * sizeof(float);
    return (b + 255) & ~(size_t)255;
}

size_t bayesn_nuts_workspace_bytes(bayesn_ctx* ctx, const bayesn_nuts_cfg* cfg) {
    if (!ctx || !cfg || !ctx->have_model || !ctx->have_data) return 0;
    // tree vectors of every (SN, chain) + 256 bytes for the work-queue counter of the launch (kept in the
    // caller's workspace, not in the context, so that launches on different streams never share one)
    return nuts_vector_bytes(ctx, cfg) + 256;
}

int bayesn_nuts_fit(bayesn_ctx* ctx, const bayesn_nuts_cfg* cfg, const float* q_init, float* samples, float* stats,
                    float* chain_info, void* work, size_t work_bytes, void* stream) {
    if (!ctx || !cfg) return -1;
    if (!ctx->have_model || !ctx->have_data) return fail(ctx, -1, "set_model and bind_dataset first");
    if (cfg->num_chains < 1 || cfg->num_samples < 1 || cfg->num_warmup < 0 || cfg->max_tree_depth < 1 ||
        cfg->max_tree_depth > 12)
        return fail(ctx, -1, "bad NUTS configuration");
    int kw = 1;
    if (int rc = team_width(ctx, cfg->warps_per_chain, &kw)) return rc;
    if (work_bytes < bayesn_nuts_workspace_bytes(ctx, cfg)) return fail(ctx, -1, "workspace too small");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    NutsDev nd{};
    nd.num_warmup = cfg->num_warmup; nd.num_samples = cfg->num_samples; nd.num_chains = cfg->num_chains;
    nd.max_depth = cfg->max_tree_depth; nd.target_accept = cfg->target_accept; nd.init_step = cfg->init_step_size;
    nd.adapt_step = cfg->adapt_step_size; nd.adapt_mass = cfg->adapt_mass_matrix;
    nd.regularize = cfg->regularize_mass_matrix;
    nd.seed_lo = (unsigned)(cfg->seed & 0xffffffffu); nd.seed_hi = (unsigned)(cfg->seed >> 32);
    nd.unit_offset = (unsigned)cfg->unit_offset;
    std::vector<int> ends = adaptation_window_ends(cfg->num_warmup);
    if (ends.size() > 24) return fail(ctx, -1, "too many adaptation windows");
    nd.n_windows = (int)ends.size();
    for (size_t i = 0; i < 24; ++i) nd.win_end[i] = i < ends.size() ? ends[i] : 0x7fffffff;
    bool use_queue = false, stage = false;
    choose_nuts_schedule(ctx, cfg->num_chains, kw, &use_queue, &stage);
    unsigned long long* queue_slot =
        use_queue ? reinterpret_cast<unsigned long long*>(static_cast<char*>(work) + nuts_vector_bytes(ctx, cfg)) : nullptr;
    NutsCall call{ctx, nd, kw, queue_slot, q_init, samples, stats, chain_info, (float*)work, (cudaStream_t)stream};
    return dispatch(ctx, stage, call);
}

int bayesn_init_to_median(bayesn_ctx* ctx, int n_chains, uint64_t seed, uint64_t unit_offset, float* q, void* stream) {
    if (!ctx) return -1;
    if (!ctx->have_model || !ctx->have_data) return fail(ctx, -1, "set_model and bind_dataset first");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    InitArgs A{};
    A.N = ctx->D.N; A.C = n_chains; A.D = ctx->M.D; A.n_eps = ctx->M.n_eps; A.rv_mode = ctx->M.rv_mode;
    A.tauA = 1.f / ctx->M.inv_tauA; A.sdD = 1.f / sqrtf(ctx->M.inv_sd2); A.muhat = ctx->D.muhat;
    A.seed_lo = (unsigned)(seed & 0xffffffffu); A.seed_hi = (unsigned)(seed >> 32); A.q = q;
    A.unit_offset = (unsigned)unit_offset;
    const long long total = (long long)A.N * A.C * A.D;
    init_median_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(A);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}

int bayesn_debug_poison_smem(bayesn_ctx* ctx, void* stream) {
    if (!ctx) return -1;
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    int sms = 148;
    CU_CHECK(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    const int bytes = 110 * 1024;                     // two CTAs per SM cover ~220 KB of the 228 KB
    CU_CHECK(ctx, cudaFuncSetAttribute(poison_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    poison_smem_kernel<<<2 * sms, 256, bytes, (cudaStream_t)stream>>>(bytes / 4);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}

int bayesn_ffma_peak(bayesn_ctx* ctx, double* tflops_out) {
    if (!ctx || !tflops_out) return -1;
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CU_CHECK(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    float* d = nullptr;
    CU_CHECK(ctx, cudaMalloc(&d, (size_t)blocks * threads * sizeof(float)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        ffma_peak_kernel<<<blocks, threads>>>(d, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    CU_CHECK(ctx, cudaGetLastError());
    *tflops_out = best;
    return 0;
}

int bayesn_flux_batch(bayesn_ctx* ctx, int N, int N_obs, int n_b, int mag, float M0, const float* theta,
                      const float* AV, const float* W0, const float* W1, const float* eps, const float* Ds,
                      const float* RV, const int32_t* band_indices, const uint8_t* mask, const float* J_t,
                      const float* hsiao_interp, const float* weights, const double* band_log10_scale, float* out,
                      void* stream) {
    if (!ctx) return -1;
    if (!ctx->have_model) return fail(ctx, -1, "set_model first");
    if (N <= 0 || N_obs <= 0) return 0;
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<double> bs(n_b);
    for (int k = 0; k < n_b; ++k) bs[k] = band_log10_scale[k] * 3.321928094887362;   // -> log2
    double* dbs = nullptr;
    CU_CHECK(ctx, cudaMallocAsync((void**)&dbs, n_b * sizeof(double), st));
    CU_CHECK(ctx, cudaMemcpyAsync(dbs, bs.data(), n_b * sizeof(double), cudaMemcpyHostToDevice, st));
    CU_CHECK(ctx, cudaStreamSynchronize(st));   // bs is a host temporary
    FluxArgs A{};
    A.N = N; A.N_obs = N_obs; A.n_b = n_b; A.L = ctx->M.L; A.T = ctx->M.T; A.mag = mag; A.M0 = M0;
    A.theta = theta; A.AV = AV; A.W0 = W0; A.W1 = W1; A.eps = eps; A.Ds = Ds; A.RV = RV;
    A.band = band_indices; A.mask = mask; A.J_t = J_t; A.hsi = hsiao_interp; A.weights = weights;
    A.bscale = dbs; A.Jl = ctx->M.Jl; A.hs = ctx->M.hs; A.Mf = ctx->M.Mf; A.LPs = jl_row_stride(ctx->LP); A.out = out;
    const long long items = (long long)N * N_obs;
    flux_kernel<false><<<(unsigned)((items + 3) / 4), 128, 0, st>>>(A);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    CU_CHECK(ctx, cudaFreeAsync(dbs, st));
    return 0;
}


int bayesn_spectra_batch(bayesn_ctx* ctx, int N, int N_obs, const float* theta, const float* AV, const float* W0,
                         const float* W1, const float* eps, const float* RV, const float* J_t, const float* hsiao_interp,
                         float* out, void* stream) {
    if (!ctx) return -1;
    if (!ctx->have_model) return fail(ctx, -1, "set_model first");
    if (N <= 0 || N_obs <= 0) return 0;
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    FluxArgs A{};
    A.N = N; A.N_obs = N_obs; A.L = ctx->M.L; A.T = ctx->M.T;
    A.theta = theta; A.AV = AV; A.W0 = W0; A.W1 = W1; A.eps = eps; A.RV = RV; A.J_t = J_t; A.hsi = hsiao_interp;
    A.Jl = ctx->M.Jl; A.hs = ctx->M.hs; A.Mf = ctx->M.Mf; A.LPs = jl_row_stride(ctx->LP); A.out = out;
    const long long items = (long long)N * N_obs;
    flux_kernel<true><<<(unsigned)((items + 3) / 4), 128, 0, (cudaStream_t)stream>>>(A);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}

int bayesn_chain_flux(bayesn_ctx* ctx, int N, int S, int n_t, int max_bands, int n_b, int mag, int rv_per_draw,
                      const float* t, const float* theta, const float* AV, const float* tmax, const float* dD,
                      const float* RV, const float* eps, const int32_t* bands, const float* weights,
                      const int32_t* support, int wt_stride, float* out, void* stream) {
    if (!ctx) return -1;
    if (!ctx->have_model) return fail(ctx, -1, "set_model first");
    if (N <= 0 || S <= 0 || n_t <= 0 || max_bands <= 0) return 0;
    if (rv_per_draw && !RV) return fail(ctx, -1, "rv_per_draw needs RV");
    if (wt_stride <= 0 || (wt_stride % 4) || ((uintptr_t)support & 15))
        return fail(ctx, -1, "support must be 16-byte aligned and wt_stride a positive multiple of 4");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    ChainFluxArgs A{};
    A.N = N; A.S = S; A.n_t = n_t; A.max_bands = max_bands; A.n_b = n_b; A.L = ctx->M.L; A.T = ctx->M.T;
    A.n_eps = ctx->M.n_eps; A.LP = ctx->LP; A.TP = ctx->TP; A.RS = jl_row_stride(ctx->LP); A.mag = mag; A.rvfree = rv_per_draw;
    A.t = t; A.theta = theta; A.AV = AV; A.tmax = tmax; A.dD = dD; A.RV = RV; A.eps = eps; A.bands = bands;
    A.wt = weights; A.sup = reinterpret_cast<const int4*>(support); A.wt_stride = wt_stride;
    A.Jl = ctx->M.Jl; A.hs = ctx->M.hs; A.Mf = ctx->M.Mf; A.ad = ctx->M.ad; A.W0 = ctx->M.W0; A.W1 = ctx->M.W1;
    A.KD = ctx->M.KD; A.tau = ctx->M.tau; A.out = out;
    const long long items = (long long)N * S * n_t;
    const unsigned grid = (unsigned)((items + 3) / 4);
    cudaStream_t st = (cudaStream_t)stream;
    if (ctx->TP == 6) chain_flux_kernel<6><<<grid, 128, 0, st>>>(A);
    else if (ctx->TP == 9) chain_flux_kernel<9><<<grid, 128, 0, st>>>(A);
    else if (ctx->TP == 10) chain_flux_kernel<10><<<grid, 128, 0, st>>>(A);
    else return fail(ctx, -4, "no kernel instantiation for this model shape");
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}

int bayesn_train_dims(bayesn_ctx* ctx, int* G, int* Dl, int* R) {
    if (!ctx || !ctx->have_model) return -1;
    const int ne = ctx->M.n_eps, L = ctx->M.L, T = ctx->M.T;
    if (G) *G = 2 * L * T + ne + ne * (ne - 1) / 2 + 3;
    if (Dl) *Dl = ne + 3;
    if (R) *R = train_red_len(ctx->LP * ctx->TP, ne);
    return 0;
}

int bayesn_train_begin(bayesn_ctx* ctx, int n_chains, const float* qg, const float* ql, float* grad_l, double* red,
                       void* stream) {
    if (!ctx) return -1;
    if (!ctx->have_model || !ctx->have_data) return fail(ctx, -1, "set_model and bind_dataset first");
    if (ctx->M.rv_mode != BAYESN_RV_FIXED) return fail(ctx, -1, "training needs rv_mode = BAYESN_RV_FIXED");
    if (!ctx->D.muerr2) return fail(ctx, -1, "training data set needs muhat_err2");
    if (n_chains < 1) return fail(ctx, -1, "n_chains must be >= 1");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    if (int rc = train_prepare(ctx, n_chains)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t msm = (size_t)ctx->M.n_eps * ctx->M.n_eps * sizeof(float);
    train_globals_kernel<<<n_chains, 128, msm, st>>>(ctx->M, ctx->T, ctx->LP, ctx->TP, qg);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    TrainSnCall call{ctx, ql, grad_l, st};
    if (int rc = dispatch(ctx, /*stage=*/n_chains >= 2, call)) return rc;
    if (int rc = launch_train_reduce(ctx, ql, (size_t)ctx->T.Dl, (size_t)n_chains * ctx->T.Dl, red, st)) return rc;
    return 0;
}

int bayesn_train_finish(bayesn_ctx* ctx, int n_chains, const float* qg, double* red, double* logp,
                        float* grad_g, void* stream) {
    if (!ctx) return -1;
    if (!ctx->have_model || ctx->train_C != n_chains) return fail(ctx, -1, "bayesn_train_begin first");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    train_finish_kernel<<<n_chains, 128, (size_t)ctx->M.n_eps * ctx->M.n_eps * sizeof(float), (cudaStream_t)stream>>>(ctx->M, ctx->T, ctx->LP, ctx->TP, qg, red, logp,
                                                                     grad_g, ctx->X);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}


int bayesn_train_nuts_init(bayesn_ctx* ctx, const bayesn_nuts_cfg* cfg, int sn_offset, const float* qg0,
                           const float* ql0, float* samples_g, float* samples_l, float* stats, void* stream) {
    if (!ctx || !cfg) return -1;
    if (!ctx->have_model || !ctx->have_data) return fail(ctx, -1, "set_model and bind_dataset first");
    if (ctx->M.rv_mode != BAYESN_RV_FIXED) return fail(ctx, -1, "training needs rv_mode = BAYESN_RV_FIXED");
    if (!ctx->D.muerr2) return fail(ctx, -1, "training data set needs muhat_err2");
    if (cfg->num_chains < 1 || cfg->num_samples < 1 || cfg->num_warmup < 0 || cfg->max_tree_depth < 1 ||
        cfg->max_tree_depth > 12)
        return fail(ctx, -1, "bad NUTS configuration");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    const int C = cfg->num_chains;
    if (int rc = train_prepare(ctx, C)) return rc;
    for (void* p : ctx->tn_owned) cudaFree(p);
    ctx->tn_owned.clear();
    ctx->have_tn = false;
    TrainNutsDev& TN = ctx->TN;
    TN = TrainNutsDev{};
    NutsDev& nd = TN.cfg;
    nd.num_warmup = cfg->num_warmup; nd.num_samples = cfg->num_samples; nd.num_chains = C;
    nd.max_depth = cfg->max_tree_depth; nd.target_accept = cfg->target_accept; nd.init_step = cfg->init_step_size;
    nd.adapt_step = cfg->adapt_step_size; nd.adapt_mass = cfg->adapt_mass_matrix;
    nd.regularize = cfg->regularize_mass_matrix;
    nd.seed_lo = (unsigned)(cfg->seed & 0xffffffffu); nd.seed_hi = (unsigned)(cfg->seed >> 32);
    std::vector<int> ends = adaptation_window_ends(cfg->num_warmup);
    if (ends.size() > 24) return fail(ctx, -1, "too many adaptation windows");
    nd.n_windows = (int)ends.size();
    for (size_t i = 0; i < 24; ++i) nd.win_end[i] = i < ends.size() ? ends[i] : 0x7fffffff;
    TN.VPL = (ctx->LP * ctx->TP - 2 * ctx->TP + 3 + 31) / 32;
    TN.sn_offset = sn_offset;
    TN.samples_g = samples_g; TN.samples_l = samples_l; TN.stats = stats;
    const int nvec = train_nuts_num_vectors(nd.max_depth);
    auto alloc = [&](auto** out, size_t n) -> int {
        void* d = nullptr;
        cudaError_t e = cudaMalloc(&d, n * sizeof(**out));
        if (e != cudaSuccess) return fail(ctx, -2, std::string("cudaMalloc: ") + cudaGetErrorString(e));
        ctx->tn_owned.push_back(d);
        *out = static_cast<std::remove_reference_t<decltype(**out)>*>(d);
        return 0;
    };
    int rc = 0;
    rc |= alloc(&TN.ctl, (size_t)C);
    rc |= alloc(&TN.work_l, (size_t)ctx->D.N * C * nvec * 32 * TN.VPL);
    rc |= alloc(&TN.work_g, (size_t)C * nvec * ctx->T.G);
    rc |= alloc(&TN.gtmp, (size_t)C * ctx->T.G);
    if (rc) return rc;
    train_nuts_init_kernel<<<C, 256, (size_t)ctx->M.n_eps * ctx->M.n_eps * sizeof(float), (cudaStream_t)stream>>>(ctx->M, ctx->T, TN, ctx->D, ctx->LP, ctx->TP, qg0, ql0);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    ctx->have_tn = true;
    return 0;
}

int bayesn_train_nuts_sn(bayesn_ctx* ctx, double* red, void* stream) {
    if (!ctx) return -1;
    if (!ctx->have_tn) return fail(ctx, -1, "bayesn_train_nuts_init first");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    TrainNutsSnCall call{ctx, st};
    if (int rc = dispatch(ctx, /*stage=*/ctx->T.C >= 2, call)) return rc;
    const size_t vec = (size_t)32 * ctx->TN.VPL, nvec = train_nuts_num_vectors(ctx->TN.cfg.max_depth);
    // fused mode: the control kernel adds the chunk partials (and exchanges them) itself
    if (int rc = launch_train_reduce(ctx, ctx->TN.work_l + TV_CZ * vec, nvec * vec, (size_t)ctx->T.C * nvec * vec, red, st,
                                     !ctx->fuse_ctl))
        return rc;
    return 0;
}

int bayesn_train_nuts_ctl(bayesn_ctx* ctx, double* red, void* stream) {
    if (!ctx) return -1;
    if (!ctx->have_tn) return fail(ctx, -1, "bayesn_train_nuts_init first");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    const size_t csm = (size_t)ctx->T.R * sizeof(double) +
                       ((size_t)ctx->M.n_eps * ctx->M.n_eps + ctx->T.G) * sizeof(float);
    CU_CHECK(ctx, cudaFuncSetAttribute(train_nuts_ctl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csm));
    train_nuts_ctl_kernel<<<ctx->T.C, 256, csm, (cudaStream_t)stream>>>(ctx->M, ctx->T, ctx->TN, ctx->LP, ctx->TP, red, ctx->X,
                                                                        ctx->fuse_ctl ? 1 : 0, (ctx->D.N + TRAIN_RCHUNK - 1) / TRAIN_RCHUNK);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}

int bayesn_train_nuts_status(bayesn_ctx* ctx, float* out, void* stream) {
    if (!ctx || !out) return -1;
    if (!ctx->have_tn) return fail(ctx, -1, "bayesn_train_nuts_init first");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    const int C = ctx->T.C;
    if (ctx->X.world > 1) {
        unsigned xerr = 0;
        CU_CHECK(ctx, cudaMemcpyAsync(&xerr, ctx->X.err, sizeof(unsigned), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
        CU_CHECK(ctx, cudaStreamSynchronize((cudaStream_t)stream));
        if (xerr) return fail(ctx, -5, "peer-memory exchange timed out waiting for a rank: the training chains were stopped");
    }
    std::vector<TrainCtl> h(C);
    CU_CHECK(ctx, cudaMemcpyAsync(h.data(), ctx->TN.ctl, C * sizeof(TrainCtl), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU_CHECK(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    int done = 0;
    for (int c = 0; c < C; ++c) {
        float* o = out + (size_t)c * 8;
        o[0] = (float)h[c].phase; o[1] = (float)h[c].it; o[2] = h[c].step; o[3] = h[c].mean_acc;
        o[4] = (float)h[c].total_steps; o[5] = (float)h[c].n_div; o[6] = (float)h[c].pe; o[7] = (float)h[c].depth;
        done += h[c].phase == 2;
    }
    return done == C ? 1 : 0;
}


static int tiles_dev(bayesn_ctx* ctx, const bayesn_tiles_desc* d, TilesDev& T) {
    if (!ctx || !d) return -1;
    if (d->N <= 0 || d->n_b <= 0 || d->n_b > 32 || d->n_bw < 300 * 51 + 2) return fail(ctx, -1, "bad tile description");
    T.N = d->N; T.n_b = d->n_b; T.n_bw = d->n_bw;
    T.master = d->master; T.z = d->z; T.ebv = d->ebv; T.fold = d->fold; T.bscale = d->bscale; T.used = d->used;
    T.model_wave = d->model_wave; T.band_spacing = d->band_spacing; T.RV_MW = d->RV_MW;
    for (int i = 0; i < 9; ++i) { T.xk[i] = d->f99_xk[i]; T.yk[i] = d->f99_yk[i]; T.mk[i] = d->f99_mk[i]; }
    return 0;
}

int bayesn_tiles_support(bayesn_ctx* ctx, const bayesn_tiles_desc* d, int32_t* support, int32_t* total_width,
                         void* stream) {
    TilesDev T{};
    if (int rc = tiles_dev(ctx, d, T)) return rc;
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const long long units = (long long)T.N * T.n_b;
    tiles_support_kernel<<<(unsigned)((units + 3) / 4), 128, 0, st>>>(T, reinterpret_cast<int4*>(support));
    tiles_offsets_kernel<<<(T.N + 255) / 256, 256, 0, st>>>(T.N, T.n_b, reinterpret_cast<int4*>(support), total_width);
    ctx->launches += 2;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}

int bayesn_tiles_fill(bayesn_ctx* ctx, const bayesn_tiles_desc* d, const int32_t* support, int wt_stride,
                      float* weights, void* stream) {
    TilesDev T{};
    if (int rc = tiles_dev(ctx, d, T)) return rc;
    if (wt_stride <= 0 || wt_stride % 4) return fail(ctx, -1, "wt_stride must be a positive multiple of 4");
    CU_CHECK(ctx, cudaSetDevice(ctx->device));
    const long long units = (long long)T.N * T.n_b;
    tiles_fill_kernel<<<(unsigned)((units + 3) / 4), 128, 0, (cudaStream_t)stream>>>(
        T, reinterpret_cast<const int4*>(support), wt_stride, weights);
    ctx->launches++;
    CU_CHECK(ctx, cudaGetLastError());
    return 0;
}


int bayesn_train_set_exchange(bayesn_ctx* ctx, int world, int rank, const uint64_t* peer_bufs, const uint64_t* peer_flags,
                              void* scratch) {
    if (!ctx) return -1;
    if (world < 1 || world > 8 || rank < 0 || rank >= world) return fail(ctx, -1, "exchange supports 1..8 ranks");
    XchgDev X{};
    X.world = world; X.rank = rank;
    if (world > 1) {
        if (!peer_bufs || !peer_flags || !scratch) return fail(ctx, -1, "exchange needs peer buffers, flags and scratch");
        for (int p = 0; p < world; ++p) {
            X.peer_buf[p] = reinterpret_cast<double*>(peer_bufs[p]);
            X.peer_flag[p] = reinterpret_cast<unsigned*>(peer_flags[p]);
        }
        X.my_buf = X.peer_buf[rank];
        X.my_flag = X.peer_flag[rank];
        unsigned* sc = static_cast<unsigned*>(scratch);      // [4] zero-initialised device words: seq, done, err
        X.seq = sc; X.done = sc + 1; X.err = sc + 2;
    }
    ctx->X = X;
    return 0;
}

#ifdef BAYESN_PHASE_TIMING
int bayesn_debug_phase(unsigned long long* out16, int reset) {
    if (out16 && cudaMemcpyFromSymbol(out16, bayesn::g_phase, sizeof(unsigned long long) * 16) != cudaSuccess) return -1;
    if (reset) {
        unsigned long long z[16] = {};
        if (cudaMemcpyToSymbol(bayesn::g_phase, z, sizeof(z)) != cudaSuccess) return -1;
    }
    return 0;
}
#endif

#ifdef CTL_TIMING
int bayesn_debug_ctl_times(long long* out) { return cudaMemcpyFromSymbol(out, bayesn::g_ctl_t, sizeof(long long) * 16) == cudaSuccess ? 0 : -1; }
#endif
}  // extern "C"
