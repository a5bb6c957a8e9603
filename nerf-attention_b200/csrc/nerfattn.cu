// libnerfattn.so -- C ABI + host orchestration (see include/nerfattn.h).
//
// nerfattn_fit_batched groups the job list by shape (N, D, H, L); all fits of a
// group run through the same grouped kernels, one epoch of all groups is
// captured once as a CUDA graph (groups on parallel branches) and replayed
// `epochs` times with a device-side epoch counter -- no host sync, no per-epoch
// launch storm (reference: ~45 launches + 1 sync per epoch, siren.py:98-105).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "common.cuh"
#include "siren_fp32.cuh"
#include "siren_tc.cuh"
#include "siren_chain.cuh"
#include "siren_dw.cuh"
#include "siren_resident.cuh"
#include "decode.cuh"
#include "synth.cuh"

namespace na {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static bool env_flag(const char* name) {
    const char* v = getenv(name);
    return v && v[0] && strcmp(v, "0") != 0;
}
static bool chain_enabled() { return !env_flag("NERFATTN_NO_CHAIN"); }
// narrow one-hidden-layer fits (tiny, small) train in the fit-resident kernel (siren_resident.cuh); NERFATTN_NO_RESIDENT=1
// sends them through the row-tile chain like every other shape (the tests compare the two)
static bool resident_enabled() { return chain_enabled() && !env_flag("NERFATTN_NO_RESIDENT"); }
// widest hidden layer that takes the fit-resident kernel (NERFATTN_RESIDENT_MAX_H: experiments, profiles/shard_ab.py)
static int resident_max_h() { const char* e = getenv("NERFATTN_RESIDENT_MAX_H"); return e ? atoi(e) : 128; }
// NERFATTN_RESIDENT=1 sends every eligible group to the fit-resident kernel, whatever else is in the call (the tests
// use it: their calls are too small for the planner below to pick it)
static bool resident_forced() { const char* e = getenv("NERFATTN_RESIDENT"); return e && atoi(e) == 1; }

// The fit-resident kernel is the cheapest way to train a narrow fit in SM-time (one SM per fit, no activation traffic),
// but not in latency: a fit is a serial chain of row tiles, ~12.1 us (`small`) / ~7.5 us (`tiny`) per tile and epoch
// on a B200, however few fits there are -- while the same fits take ~2 us each per epoch through the chain + dW
// kernels (40 us at least), spread over all SMs.  So it pays exactly when its CTAs can hide beside enough other work:
// measured on the shards of the 280-fit sweep (profiles/shard_ab.py), 280 / 140 / 70 fits per GPU: all narrow fits
// resident is fastest; 35 fits per GPU (8 GPUs): `small` resident is the critical path (0.288 ms per epoch against
// 0.244 with only `tiny` resident).  Rule: resident iff the group's epoch on its SMs takes at most 3/4 of the epoch of
// everything that is not resident-eligible (estimated from its FLOPs at the sweep's measured rate); a call of narrow
// fits only takes the kernel that finishes first.
static bool resident_pays(int H, int mtiles, int nf_group, double rest_seconds_per_epoch) {
    const double t_res = mtiles * (H <= 64 ? 7.5e-6 : 12.1e-6);
    if (rest_seconds_per_epoch > 0) return t_res <= 0.75 * rest_seconds_per_epoch;
    const double t_chain = std::max(40e-6, nf_group * (H <= 64 ? 1.8e-6 : 2.2e-6) * (mtiles / 16.0));
    return t_res < t_chain;
}

// ------------------------------------------------------------------ planning
struct Group {
    int N, D, H, L, nf;
    LayerMap lm;
    int mtiles;            // row tiles of 128 (also the granularity of bias-gradient partials)
    int nsplit, ksplit;    // split-K of the dW GEMMs (fp32 path)
    std::vector<int> fit_idx;
    FitRec* d_recs;
    // activations / cos / dZ: [nf][N][H]; fp32 path uses float, tensor path bf16
    void* act[kMaxHidden + 1];
    void* cosb[kMaxHidden + 1];
    void* dz[2];
    void* dy;              // [nf][N][D]  (also y of the final evaluation: always fp32-sized)
    float* yeval;          // [nf][N][D] fp32
    float* evalact[2];     // fp32 ping-pong activations for the final fp32 forward (tensor mode)
    float* gradpart;       // [nsplit][nf][P]
    float* colpart; size_t colpart_layer_off[kMaxLayers];
    float* xpart;          // [nf][mtiles][H]
    float* losspart; int losspart_per_fit;
    __nv_bfloat16* wbf16;  // [nf][P] bf16 mirror of the weights (tensor path)
    tc::GroupMaps* maps;   // TMA descriptors (unfused tensor path), host-side
    // fused row-tile chain (siren_chain.cuh): forward + loss + dX chain in one kernel; cosb[l] then
    // holds dz_l (the dW operand) and cos_l lives in the per-CTA scratch
    bool use_chain;
    bool use_resident;     // siren_resident.cuh: one persistent CTA per fit, one launch for all epochs (no per-epoch kernels)
    __nv_bfloat16* chain_scratch;
    float* psc;            // omega-prescaled W0 / sine-layer biases [nf][(L+2)H]
    chain::ChainMaps cmaps;
    dw::DwMaps dmaps;
    // layer-0 gradient operand (chain::xop_kernel): one table per distinct position vector of the group
    std::vector<const float*> pos_tabs; std::vector<int> posid;    // posid[k]: table of the group's k-th fit
    const float** d_pos_tabs; __nv_bfloat16* xop;
    // every group counts its own epochs (the Adam kernel that ends a group's epoch increments the counter), so
    // groups never wait for one another inside a multi-epoch graph
    int* d_epoch; unsigned int* d_done;
};

// A shape group larger than this is split into several groups ("units") with their own buffers: the two-lane
// schedule (record_epochs) overlaps the HBM-bound dW + Adam kernel of one unit with the chain kernel of the next,
// which needs more than a few units per epoch.  NERFATTN_UNIT_FITS overrides (0 = never split).
static int unit_fits_cap() {
    const char* e = getenv("NERFATTN_UNIT_FITS");
    const int v = e ? atoi(e) : 0;
    return v > 0 ? v : (1 << 30);
}

struct Plan {
    std::vector<Group> groups;
    // unique target tensors
    std::vector<const float*> uniq_ptr;
    std::vector<int> uniq_N, uniq_D, uniq_prenorm, uniq_first_fit;
    std::vector<size_t> uniq_tnorm_off, uniq_stat_off;
    const float** d_uniq_ptr; int* d_uniq_prenorm;
    float* tnorm; float* ustat_mean; float* ustat_std;
    float* d_step_size; float* d_bc2;
    size_t bytes;
};

static int validate(const na_fit_t* fits, int nfits, int precision) {
    if (!fits || nfits <= 0) { set_error("fits must be a non-empty array"); return NA_ERR_INVALID; }
    if (precision != NA_PREC_FP32 && precision != NA_PREC_BF16) {
        set_error("precision %d not implemented (0 = fp32, 2 = bf16)", precision);
        return NA_ERR_UNSUPPORTED;
    }
    for (int i = 0; i < nfits; ++i) {
        const na_fit_t& f = fits[i];
        if (f.N < 2 || f.D < 1 || f.H < 1 || f.L < 0) { set_error("fit %d: bad shape", i); return NA_ERR_INVALID; }
        if (f.L > kMaxHidden) { set_error("fit %d: hidden_layers %d > %d", i, f.L, kMaxHidden); return NA_ERR_UNSUPPORTED; }
        if (f.H % 8 || f.D % 4) { set_error("fit %d: H must be a multiple of 8 and D a multiple of 4", i); return NA_ERR_UNSUPPORTED; }
        if (precision == NA_PREC_BF16 && !tc::shape_supported(f.N, f.D, f.H, f.L) &&
            !(chain_enabled() && chain::shape_supported(f.N, f.D, f.H, f.L)) &&
            !(resident_enabled() && res::shape_supported(f.N, f.D, f.H, f.L))) {
            set_error("fit %d: bf16 path needs H in {64,128,256,512}, D in {64,128,256} and hidden_layers >= 1 "
                      "(any N), or N %% 128 == 0 without hidden layers (got N=%d D=%d H=%d L=%d)", i, f.N, f.D, f.H, f.L);
            return NA_ERR_UNSUPPORTED;
        }
        if (!f.positions || !f.targets || !f.params) { set_error("fit %d: null pointer", i); return NA_ERR_INVALID; }
    }
    return NA_OK;
}

// Lay the workspace out.  With ws == nullptr this only computes sizes.
static void make_plan(const na_fit_t* fits, int nfits, int epochs, int precision, void* ws, Plan& plan) {
    Arena ar(ws);
    std::map<std::tuple<int, int, int, int>, int> gid;                 // shape -> the group currently being filled
    const int cap = (precision == NA_PREC_BF16) ? unit_fits_cap() : (1 << 30);
    for (int i = 0; i < nfits; ++i) {
        auto key = std::make_tuple(fits[i].N, fits[i].D, fits[i].H, fits[i].L);
        auto it = gid.find(key);
        if (it == gid.end() || (int)plan.groups[it->second].fit_idx.size() >= cap) {
            Group g{};
            g.N = fits[i].N; g.D = fits[i].D; g.H = fits[i].H; g.L = fits[i].L;
            g.lm = make_layer_map(g.H, g.L, g.D);
            gid[key] = (int)plan.groups.size();
            it = gid.find(key);
            plan.groups.push_back(g);
        }
        plan.groups[it->second].fit_idx.push_back(i);
    }
    // unique targets
    std::map<std::tuple<const float*, int, int, int, const float*>, int> uid;
    std::vector<int> fit_uniq(nfits);
    size_t tnorm_total = 0, stat_total = 0;
    for (int i = 0; i < nfits; ++i) {
        const int pre = (fits[i].flags & NA_FIT_TARGETS_PRENORMALISED) ? 1 : 0;
        auto key = std::make_tuple(fits[i].targets, fits[i].N, fits[i].D, pre, pre ? fits[i].mean : nullptr);
        auto it = uid.find(key);
        if (it == uid.end()) {
            it = uid.emplace(key, (int)plan.uniq_ptr.size()).first;
            plan.uniq_ptr.push_back(fits[i].targets);
            plan.uniq_N.push_back(fits[i].N); plan.uniq_D.push_back(fits[i].D);
            plan.uniq_prenorm.push_back(pre); plan.uniq_first_fit.push_back(i);
            plan.uniq_tnorm_off.push_back(tnorm_total); plan.uniq_stat_off.push_back(stat_total);
            tnorm_total += align_up((size_t)fits[i].N * fits[i].D, 64);
            stat_total += align_up((size_t)fits[i].D, 64);
        }
        fit_uniq[i] = it->second;
    }
    const int nu = (int)plan.uniq_ptr.size();
    plan.d_uniq_ptr = ar.take<const float*>(nu);
    plan.d_uniq_prenorm = ar.take<int>(nu);
    plan.tnorm = ar.take<float>(tnorm_total);
    plan.ustat_mean = ar.take<float>(stat_total);
    plan.ustat_std = ar.take<float>(stat_total);
    plan.d_step_size = ar.take<float>(std::max(epochs, 1));
    plan.d_bc2 = ar.take<float>(std::max(epochs, 1));

    const bool bf = precision == NA_PREC_BF16;
    const size_t esz = bf ? 2 : 4;
    auto resident_eligible = [&](const Group& g) {
        return bf && resident_enabled() && g.H <= resident_max_h() && res::shape_supported(g.N, g.D, g.H, g.L);
    };
    double rest_seconds = 0;                                // one epoch of everything that cannot be fit-resident
    for (const Group& g : plan.groups)
        if (!resident_eligible(g))
            rest_seconds += (double)g.fit_idx.size() * (6.0 * g.N * ((double)g.L * g.H * g.H + (double)g.H * g.D) + 4.0 * g.N * g.H) / 400e12;
    for (Group& g : plan.groups) {
        g.nf = (int)g.fit_idx.size();
        g.mtiles = ceil_div(g.N, 128);
        g.d_recs = ar.take<FitRec>(g.nf);
        g.use_resident = resident_eligible(g) && (resident_forced() || resident_pays(g.H, g.mtiles, g.nf, rest_seconds));
        g.use_chain = !g.use_resident && bf && chain_enabled() && chain::shape_supported(g.N, g.D, g.H, g.L);
        g.d_epoch = ar.take<int>(64);
        g.d_done = ar.take<unsigned int>(64);
        const size_t nh = (size_t)g.nf * g.N * g.H, nd = (size_t)g.nf * g.N * g.D;
        if (g.use_resident) {                              // everything between two Adam steps lives on the fit's SM
            g.yeval = ar.take<float>(nd);
            g.evalact[0] = ar.take<float>(nh); g.evalact[1] = ar.take<float>(nh);
            g.nsplit = 1; g.ksplit = g.N;
            continue;
        }
        for (int l = 0; l <= g.L; ++l) {
            g.act[l] = ar.take<char>(nh * esz);
            if (l > 0 || !g.use_chain) g.cosb[l] = ar.take<char>(nh * esz);          // chain: dz_0 never leaves the SM
        }
        if (!g.use_chain) { g.dz[0] = ar.take<char>(nh * esz); g.dz[1] = ar.take<char>(nh * esz); }
        g.dy = ar.take<char>(nd * esz);
        g.yeval = ar.take<float>(nd);
        if (bf) { g.evalact[0] = ar.take<float>(nh); g.evalact[1] = ar.take<float>(nh); }
        // split-K of dW so that the launch fills the GPU: ~2 waves of CTAs on the fp32 path (partials summed by Adam);
        // the tensor paths contract all rows of a fit in one tile
        if (bf) {
            g.nsplit = 1;
            g.ksplit = g.N;
        } else {
            int out_tiles = 0;
            for (int l = 1; l <= g.L + 1; ++l)
                out_tiles = std::max(out_tiles, ceil_div(g.lm.out_dim[l], 128) * ceil_div(g.lm.in_dim[l], 128));
            int want = std::max(1, (2 * 148 * 2) / std::max(1, g.nf * out_tiles));
            int max_split = std::max(1, g.N / 256);
            g.nsplit = std::min(want, max_split);
            g.ksplit = (int)align_up((size_t)ceil_div(g.N, g.nsplit), 16);
            g.nsplit = ceil_div(g.N, g.ksplit);
        }
        g.gradpart = g.use_chain ? nullptr : ar.take<float>((size_t)g.nsplit * g.nf * g.lm.P);   // chain: dW goes straight into Adam
        size_t coff = 0;
        for (int l = 0; l <= g.L + 1; ++l) {
            g.colpart_layer_off[l] = coff;
            coff += (size_t)g.nf * g.mtiles * g.lm.out_dim[l];
        }
        g.colpart = ar.take<float>(coff);
        g.xpart = ar.take<float>((size_t)g.nf * g.mtiles * g.H);
        g.losspart_per_fit = g.use_chain ? chain::loss_partials_per_fit(g.N, g.H)
                             : bf ? tc::loss_partials_per_fit(g.N, g.D) : g.mtiles * ceil_div(g.D, 128);
        g.losspart = ar.take<float>((size_t)g.nf * g.losspart_per_fit);
        g.wbf16 = bf ? ar.take<__nv_bfloat16>((size_t)g.nf * g.lm.P) : nullptr;
        g.maps = nullptr;
        g.posid.assign(g.nf, 0);
        for (int k = 0; k < g.nf; ++k) {
            const float* pos = fits[g.fit_idx[k]].positions;
            size_t t = 0;
            while (t < g.pos_tabs.size() && g.pos_tabs[t] != pos) ++t;
            if (t == g.pos_tabs.size()) g.pos_tabs.push_back(pos);
            g.posid[k] = (int)t;
        }
        g.d_pos_tabs = g.use_chain ? ar.take<const float*>(g.pos_tabs.size()) : nullptr;
        g.xop = g.use_chain ? ar.take<__nv_bfloat16>(chain::xop_elems(g.mtiles, (int)g.pos_tabs.size())) : nullptr;
        g.chain_scratch = g.use_chain ? ar.take<__nv_bfloat16>(chain::scratch_elems(g.H, g.L)) : nullptr;
        g.psc = g.use_chain ? ar.take<float>((size_t)g.nf * chain::psc_floats(g.H, g.L)) : nullptr;
    }
    plan.bytes = ar.bytes();
    // stash per-fit unique ids in uniq_first_fit's tail: callers use fit_uniq via closure
    plan.uniq_first_fit.insert(plan.uniq_first_fit.end(), fit_uniq.begin(), fit_uniq.end());
}

// ------------------------------------------------------------------ fp32 launches
template <int MODE, bool AK, bool BK_>
static void launch_sgemm(const f32::GemmArgs& a, int nf, int splits, cudaStream_t s) {
    dim3 grid(ceil_div(a.N, f32::BN), ceil_div(a.M, f32::BM) * splits, nf);
    f32::sgemm_kernel<MODE, AK, BK_><<<grid, f32::NTHREADS, 0, s>>>(a);
}

static void base_args(f32::GemmArgs& a, const Group& g) {
    memset(&a, 0, sizeof(a));
    a.recs = g.d_recs;
    a.a_param_off = -1; a.b_param_off = -1;
    a.loss_scale = 2.0f / ((float)g.N * (float)g.D);
}

// forward through the sine layers (fp32).  actbuf(l)/cosbuf(l) give the outputs of layer l.
// `cosb` may be nullptr (evaluation / decode do not need the derivative factor); layers 1..last.
static void fp32_forward_hidden(const Group& g, float* const* act, float* const* cosb, int last, cudaStream_t s) {
    const size_t nh = (size_t)g.N * g.H;
    {
        const size_t total8 = nh / 8;
        dim3 grid((unsigned)ceil_div(total8, (size_t)256), g.nf);
        f32::layer0_kernel<float><<<grid, 256, 0, s>>>(g.d_recs, g.N, g.H, act[0], cosb ? cosb[0] : nullptr, nh);
    }
    for (int l = 1; l <= last; ++l) {
        f32::GemmArgs a; base_args(a, g);
        a.M = g.N; a.N = g.H; a.K = g.H;
        a.A = act[l - 1]; a.a_fit = nh; a.lda = g.H;
        a.b_param_off = g.lm.w_off[l]; a.ldb = g.H;
        a.bias_off = g.lm.b_off[l];
        a.out0 = act[l]; a.out0_fit = nh;
        a.out1 = cosb ? cosb[l] : nullptr; a.out1_fit = nh;
        launch_sgemm<f32::kFwdSine, true, true>(a, g.nf, 1, s);
    }
}

static void fp32_output_layer(const Group& g, const float* actL, bool eval, cudaStream_t s,
                              float* const* yout = nullptr, int denorm = 0) {
    const size_t nh = (size_t)g.N * g.H, nd = (size_t)g.N * g.D;
    f32::GemmArgs a; base_args(a, g);
    a.M = g.N; a.N = g.D; a.K = g.H;
    a.A = actL; a.a_fit = nh; a.lda = g.H;
    a.b_param_off = g.lm.w_off[g.L + 1]; a.ldb = g.H;
    a.bias_off = g.lm.b_off[g.L + 1];
    if (eval) {
        a.out0 = yout ? nullptr : g.yeval; a.out0_fit = nd;
        a.yout = yout; a.denorm = denorm;
        launch_sgemm<f32::kFwdEval, true, true>(a, g.nf, 1, s);
    } else {
        a.out0 = (float*)g.dy; a.out0_fit = nd;
        a.colpart = g.colpart + g.colpart_layer_off[g.L + 1]; a.colpart_fit = (size_t)g.mtiles * g.D;
        a.losspart = g.losspart; a.losspart_per_fit = g.losspart_per_fit;
        launch_sgemm<f32::kFwdOut, true, true>(a, g.nf, 1, s);
    }
}

static void launch_adam(const Group& g, const Plan& plan, double beta1, double beta2, double eps, cudaStream_t s,
                        bool layer0_only = false) {
    f32::AdamArgs a{};
    a.recs = g.d_recs; a.lm = g.lm;
    a.et.epoch = g.d_epoch; a.et.step_size = plan.d_step_size; a.et.bc2_sqrt = plan.d_bc2;
    a.epoch_rw = g.d_epoch; a.done = g.d_done;            // the last block of this launch ends the group's epoch
    a.gradpart = g.gradpart; a.grad_split_stride = (size_t)g.nf * g.lm.P; a.grad_fit = g.lm.P; a.nsplit = g.nsplit;
    a.colpart = g.colpart;
    for (int l = 0; l < kMaxLayers; ++l) a.colpart_layer_off[l] = g.colpart_layer_off[l];
    for (int l = 0; l < kMaxLayers; ++l) a.col_mt[l] = (g.wbf16 && l > 0) ? g.nsplit : g.mtiles;
    a.xpart = g.xpart;
    a.losspart = g.losspart; a.losspart_per_fit = g.losspart_per_fit;
    a.loss_inv_count = 1.0f / ((float)g.N * (float)g.D);
    a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps;
    a.wbf16 = g.wbf16; a.wbf16_fit = g.lm.P;
    a.psc = g.psc; a.psc_fit = g.psc ? chain::psc_floats(g.H, g.L) : 0; a.H = g.H; a.L = g.L;
    // 64-thread blocks (3 K registers, no shared memory): small enough to be co-scheduled on SMs whose
    // registers and shared memory are almost entirely held by another group's chain CTA, so this
    // HBM-bound update overlaps that group's issue-bound kernel instead of waiting for free SMs
    a.p_end = layer0_only ? g.lm.w_off[1] : g.lm.P;
    const int adam_threads = env_flag("NERFATTN_ADAM256") ? 256 : 64;
    dim3 grid(ceil_div(a.p_end, adam_threads * 4), g.nf);
    f32::adam_kernel<<<grid, adam_threads, 0, s>>>(a);
}

// one training epoch of one group, fp32 SIMT path
static void fp32_epoch(const Group& g, const Plan& plan, double b1, double b2, double eps, cudaStream_t s) {
    const size_t nh = (size_t)g.N * g.H, nd = (size_t)g.N * g.D;
    float* act[kMaxHidden + 1]; float* cosb[kMaxHidden + 1];
    for (int l = 0; l <= g.L; ++l) { act[l] = (float*)g.act[l]; cosb[l] = (float*)g.cosb[l]; }
    fp32_forward_hidden(g, act, cosb, g.L, s);
    fp32_output_layer(g, act[g.L], false, s);

    // backward.  dz_cur holds dL/dz of layer `l`; for l = L+1 that is dY.
    const float* dz_cur = (const float*)g.dy;
    int cur_width = g.D;
    for (int l = g.L + 1; l >= 1; --l) {
        // dW_l = dz_l^T . act_{l-1}      [out_l x N] . [N x H]
        {
            f32::GemmArgs a; base_args(a, g);
            a.M = g.lm.out_dim[l]; a.N = g.H; a.K = g.N;
            a.A = dz_cur; a.a_fit = (size_t)g.N * cur_width; a.lda = cur_width;
            a.B = act[l - 1]; a.b_fit = nh; a.ldb = g.H;
            a.gradpart = g.gradpart; a.grad_split_stride = (size_t)g.nf * g.lm.P; a.grad_fit = g.lm.P;
            a.grad_off = g.lm.w_off[l]; a.ksplit = g.ksplit;
            launch_sgemm<f32::kDw, false, false>(a, g.nf, g.nsplit, s);
        }
        // dz_{l-1} = (dz_l . W_l) * omega * cos_{l-1}
        {
            f32::GemmArgs a; base_args(a, g);
            a.M = g.N; a.N = g.H; a.K = g.lm.out_dim[l];
            a.A = dz_cur; a.a_fit = (size_t)g.N * cur_width; a.lda = cur_width;
            a.b_param_off = g.lm.w_off[l]; a.ldb = g.H;
            float* dst = (float*)g.dz[(l - 1) & 1];
            a.out0 = dst; a.out0_fit = nh;
            a.cprev = cosb[l - 1]; a.cprev_fit = nh;
            a.colpart = g.colpart + g.colpart_layer_off[l - 1]; a.colpart_fit = (size_t)g.mtiles * g.H;
            if (l == 1) { a.xpart = g.xpart; a.xpart_fit = (size_t)g.mtiles * g.H; }
            launch_sgemm<f32::kDx, true, false>(a, g.nf, 1, s);
            dz_cur = dst; cur_width = g.H;
        }
    }
    (void)nd;
    launch_adam(g, plan, b1, b2, eps, s);
}

// One training epoch of one group on the fused chain path (BF16) is three launches:
//   chain_part   the row-tile chain (forward, loss, dX, layer-0 gradient partials)          -- FP32/SFU-issue-bound
//   update_part  the grouped dW + Adam kernel, then Adam of the 2H layer-0 parameters (it also writes losses[e]
//                and ends the group's epoch)                                                -- HBM-bound
// max_ctas caps the persistent grids (0 = all SMs): the two-lane schedule runs the two halves of different groups
// side by side on disjoint sets of SMs.
static int chain_part(const Group& g, int max_ctas, cudaStream_t s, bool pack = false) {
    if (!(chain::phase_mask() & 1)) return NA_OK;
    return chain::train_step(g.N, g.D, g.H, g.L, g.nf, g.lm, g.d_recs, g.cmaps, g.chain_scratch, g.losspart,
                             g.losspart_per_fit, g.mtiles, g.psc, g.xpart, g.colpart + g.colpart_layer_off[0], max_ctas, s, pack);
}
static int update_part(const Group& g, const Plan& plan, double b1, double b2, double eps, int max_ctas, cudaStream_t s) {
    const int phases = chain::phase_mask();
    if (phases & 2) {
        dw::DwArgs da{};
        dw::fill_args(da, g.N, g.D, g.H, g.L, g.lm);
        da.epoch = g.d_epoch; da.step_size = plan.d_step_size; da.bc2 = plan.d_bc2;
        da.beta1 = (float)b1; da.beta2 = (float)b2; da.eps = (float)eps;
        da.nf = g.nf; da.recs = g.d_recs;
        da.wbf16 = g.wbf16; da.wbf16_fit = g.lm.P;
        da.psc = g.psc; da.psc_fit = chain::psc_floats(g.H, g.L);
        int rc = dw::launch(g.dmaps, da, max_ctas, s);
        if (rc) return rc;
    }
    launch_adam(g, plan, b1, b2, eps, s, /*layer0_only=*/true);      // always: it ends the epoch (profiling masks included)
    return NA_OK;
}

// final evaluation: fp32 forward with the trained weights + metrics (siren.py:119-125).
// Always fp32 SIMT, also in the BF16 mode: the numbers must be what torch's model(positions)
// gives for the returned weights.
static void final_eval(const Group& g, int precision, cudaStream_t s) {
    float* act[kMaxHidden + 1];
    for (int l = 0; l <= g.L; ++l)
        act[l] = (precision == NA_PREC_FP32) ? (float*)g.act[l] : g.evalact[l & 1];
    fp32_forward_hidden(g, act, nullptr, g.L, s);
    fp32_output_layer(g, act[g.L], true, s);
    dim3 grid(ceil_div(g.N, 8), g.nf);
    f32::row_metrics_kernel<<<grid, 256, 0, s>>>(g.d_recs, g.yeval, (size_t)g.N * g.D, g.N, g.D);
    f32::fit_scalars_kernel<<<g.nf, 256, 0, s>>>(g.d_recs, g.N);
}

// Executable graphs may still be running when nerfattn_fit_batched returns (the library never
// synchronises).  They are parked here with an event and destroyed by a later call once done.
struct Parked { cudaGraphExec_t exec; cudaGraph_t graph; cudaEvent_t done; };
static std::mutex g_park_mu;
static std::vector<Parked> g_parked;
static void park_graph(cudaGraphExec_t exec, cudaGraph_t graph, cudaStream_t stream) {
    Parked p{exec, graph, nullptr};
    cudaEventCreateWithFlags(&p.done, cudaEventDisableTiming);
    cudaEventRecord(p.done, stream);
    std::lock_guard<std::mutex> lk(g_park_mu);
    g_parked.push_back(p);
}
static void reap_graphs() {
    std::lock_guard<std::mutex> lk(g_park_mu);
    for (size_t i = 0; i < g_parked.size();) {
        if (cudaEventQuery(g_parked[i].done) == cudaSuccess) {
            cudaGraphExecDestroy(g_parked[i].exec);
            cudaGraphDestroy(g_parked[i].graph);
            cudaEventDestroy(g_parked[i].done);
            g_parked[i] = g_parked.back();
            g_parked.pop_back();
        } else ++i;
    }
    cudaGetLastError();   // cudaEventQuery's cudaErrorNotReady is not an error of ours
}


// Streams and events of a graph capture (one capture stream, one side stream + join event per shape group, one
// fork event).  Creating and destroying them per call costs more than the capture itself, so they are pooled per
// device; a Lease hands one kit to a call and returns it when the call leaves, on every path.
struct CaptureKit {
    int device = -1;
    cudaStream_t cap = nullptr;
    std::vector<cudaStream_t> aux;                // one per fit-resident group: their kernels run beside the epoch graphs
    std::vector<cudaEvent_t> aux_join;
    cudaEvent_t aux_fork = nullptr;
    std::vector<cudaStream_t> streams;            // lanes / per-group branches
    std::vector<cudaEvent_t> events;              // fork, joins, cross-lane dependencies
    bool grow(size_t nstreams, size_t nevents, size_t naux) {
        if (!cap && cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) != cudaSuccess) return false;
        if (!aux_fork && cudaEventCreateWithFlags(&aux_fork, cudaEventDisableTiming) != cudaSuccess) return false;
        while (aux.size() < naux) {
            cudaStream_t st = nullptr; cudaEvent_t ev = nullptr;
            if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return false;
            aux.push_back(st);
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return false;
            aux_join.push_back(ev);
        }
        while (streams.size() < nstreams) {
            cudaStream_t st = nullptr;
            if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return false;
            streams.push_back(st);
        }
        while (events.size() < nevents) {
            cudaEvent_t ev = nullptr;
            if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) return false;
            events.push_back(ev);
        }
        return true;
    }
};
static std::mutex g_kit_mu;
static std::vector<CaptureKit*> g_kits;          // idle kits of every device
struct KitLease {
    CaptureKit* kit = nullptr;
    KitLease(size_t nstreams, size_t nevents, size_t naux) {
        const int dev = tc::current_device();
        {
            std::lock_guard<std::mutex> lk(g_kit_mu);
            for (size_t i = 0; i < g_kits.size(); ++i)
                if (g_kits[i]->device == dev) { kit = g_kits[i]; g_kits[i] = g_kits.back(); g_kits.pop_back(); break; }
        }
        if (!kit) { kit = new CaptureKit(); kit->device = dev; }
        if (!kit->grow(nstreams, nevents, naux)) {                  // leave what exists in the pool; the caller reports the CUDA error
            std::lock_guard<std::mutex> lk(g_kit_mu);
            g_kits.push_back(kit);
            kit = nullptr;
        }
    }
    ~KitLease() {
        if (!kit) return;
        std::lock_guard<std::mutex> lk(g_kit_mu);
        g_kits.push_back(kit);
    }
    KitLease(const KitLease&) = delete;
    KitLease& operator=(const KitLease&) = delete;
};
// A captured graph and its executable: destroyed on scope exit unless handed to park_graph()
struct GraphHold {
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    ~GraphHold() { if (exec) cudaGraphExecDestroy(exec); if (graph) cudaGraphDestroy(graph); }
    void release() { graph = nullptr; exec = nullptr; }
};

}  // namespace na

// =========================================================================== C ABI
using namespace na;

extern "C" int nerfattn_abi_version(void) { return NERFATTN_ABI_VERSION; }
extern "C" const char* nerfattn_last_error(void) { return g_err; }

extern "C" size_t nerfattn_param_count(int32_t H, int32_t L, int32_t D) {
    return (size_t)2 * H + (size_t)L * ((size_t)H * H + H) + (size_t)H * D + D;
}

extern "C" int nerfattn_fit_workspace_bytes(const na_fit_t* fits, int32_t nfits, int32_t precision, size_t* bytes) {
    if (!bytes) { set_error("bytes is null"); return NA_ERR_INVALID; }
    int rc = validate(fits, nfits, precision);
    if (rc) return rc;
    Plan plan{};
    make_plan(fits, nfits, 1 << 20, precision, nullptr, plan);   // tables sized for up to 1M epochs
    *bytes = plan.bytes;
    return NA_OK;
}

extern "C" long long nerfattn_fit_launch_count(const na_fit_t* fits, int32_t nfits, int32_t epochs, int32_t precision) {
    if (validate(fits, nfits, precision)) return -1;
    Plan plan{};
    make_plan(fits, nfits, 1, precision, nullptr, plan);
    long long setup = (long long)plan.uniq_ptr.size() /* <= one norm launch per tensor */ + plan.groups.size();
    long long per_epoch = 0, fin = 0;
    for (const Group& g : plan.groups) {
        // layer0 + L fwd + out + (L+1) x (dW, dX) + Adam; the tensor path adds the layer-0 gradient kernel
        if (g.use_resident) setup += 1;                           // one launch for all epochs
        else if (g.use_chain) per_epoch += 3;                     // chain, dW + Adam, Adam of layer 0
        else per_epoch += 1 + g.L + 1 + 2 * (g.L + 1) + 1 + (precision == NA_PREC_BF16 ? 1 : 0);
        fin += 1 + g.L + 1 + 2;
        if (precision == NA_PREC_BF16 && !g.use_resident) setup += 1;   // bf16 weight mirror
    }
    return setup + per_epoch * epochs + fin;
}

extern "C" int nerfattn_fit_batched(const na_fit_t* fits, int32_t nfits, int32_t epochs, const double* lr_table,
                                    double beta1, double beta2, double eps, int32_t first_step,
                                    int32_t precision, void* workspace, size_t workspace_bytes,
                                    na_stream_t stream_) {
    return nerfattn_fit_batched_ex(fits, nfits, epochs, lr_table, beta1, beta2, eps, first_step, precision, 0, nullptr,
                                   workspace, workspace_bytes, stream_);
}

extern "C" int nerfattn_fit_batched_ex(const na_fit_t* fits, int32_t nfits, int32_t epochs, const double* lr_table,
                                       double beta1, double beta2, double eps, int32_t first_step,
                                       int32_t precision, int32_t log_every, float* progress, void* workspace,
                                       size_t workspace_bytes, na_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    reap_graphs();
    int rc = validate(fits, nfits, precision);
    if (rc) return rc;
    if (epochs < 0 || epochs > (1 << 20) || (epochs > 0 && !lr_table)) { set_error("bad epochs / lr_table"); return NA_ERR_INVALID; }
    for (int i = 0; i < nfits; ++i) {
        const na_fit_t& f = fits[i];
        if (!f.mean || !f.std || !f.adam_m || !f.adam_v || !f.cos_sims || !f.per_pos_mse || !f.scalars ||
            (epochs > 0 && !f.losses)) { set_error("fit %d: null output pointer", i); return NA_ERR_INVALID; }
    }
    if (!workspace || ((uintptr_t)workspace & 255)) { set_error("workspace must be a 256-byte aligned device pointer"); return NA_ERR_WORKSPACE; }
    Plan plan{};
    make_plan(fits, nfits, 1 << 20, precision, workspace, plan);
    if (plan.bytes > workspace_bytes) {
        set_error("workspace too small: need %zu bytes, got %zu", plan.bytes, workspace_bytes);
        return NA_ERR_WORKSPACE;
    }
    const int nu = (int)plan.uniq_ptr.size();
    const int* fit_uniq = plan.uniq_first_fit.data() + nu;

    // ---- host tables -> device
    NA_CUDA_OK(cudaMemcpyAsync(plan.d_uniq_ptr, plan.uniq_ptr.data(), nu * sizeof(float*), cudaMemcpyHostToDevice, stream));
    NA_CUDA_OK(cudaMemcpyAsync(plan.d_uniq_prenorm, plan.uniq_prenorm.data(), nu * sizeof(int), cudaMemcpyHostToDevice, stream));
    for (Group& g : plan.groups) {
        NA_CUDA_OK(cudaMemsetAsync(g.d_epoch, 0, sizeof(int), stream));
        NA_CUDA_OK(cudaMemsetAsync(g.d_done, 0, sizeof(unsigned int), stream));
    }
    if (epochs > 0) {
        std::vector<float> ss(epochs), bc2(epochs);
        for (int e = 0; e < epochs; ++e) {          // torch/optim/adam.py: python-float (double) math
            const double t = (double)(first_step + e + 1);
            ss[e] = (float)(lr_table[e] / (1.0 - std::pow(beta1, t)));
            bc2[e] = (float)std::sqrt(1.0 - std::pow(beta2, t));
        }
        NA_CUDA_OK(cudaMemcpyAsync(plan.d_step_size, ss.data(), epochs * sizeof(float), cudaMemcpyHostToDevice, stream));
        NA_CUDA_OK(cudaMemcpyAsync(plan.d_bc2, bc2.data(), epochs * sizeof(float), cudaMemcpyHostToDevice, stream));
    }
    for (Group& g : plan.groups) {
        std::vector<FitRec> recs(g.nf);
        for (int k = 0; k < g.nf; ++k) {
            const na_fit_t& f = fits[g.fit_idx[k]];
            const int u = fit_uniq[g.fit_idx[k]];
            FitRec& r = recs[k];
            r.pos = f.positions; r.traw = f.targets;
            r.tnorm = plan.tnorm + plan.uniq_tnorm_off[u];
            r.mean = plan.ustat_mean + plan.uniq_stat_off[u];
            r.stdv = plan.ustat_std + plan.uniq_stat_off[u];
            r.params = f.params; r.m = f.adam_m; r.v = f.adam_v; r.losses = f.losses;
            r.cos = f.cos_sims; r.ppmse = f.per_pos_mse; r.scalars = f.scalars;
            r.mean_out = f.mean; r.std_out = f.std; r.omega = f.omega0; r.uniq = u; r.fit_index = g.fit_idx[k];
            r.prenorm = (f.flags & NA_FIT_TARGETS_PRENORMALISED) ? 1 : 0;
            r.posid = g.posid.empty() ? 0 : g.posid[k];
        }
        NA_CUDA_OK(cudaMemcpyAsync(g.d_recs, recs.data(), g.nf * sizeof(FitRec), cudaMemcpyHostToDevice, stream));
    }
    // pre-normalised targets carry their statistics in: copy them into the per-tensor slots
    for (int u = 0; u < nu; ++u) {
        if (!plan.uniq_prenorm[u]) continue;
        const na_fit_t& f = fits[plan.uniq_first_fit[u]];
        NA_CUDA_OK(cudaMemcpyAsync(plan.ustat_mean + plan.uniq_stat_off[u], f.mean, f.D * sizeof(float), cudaMemcpyDeviceToDevice, stream));
        NA_CUDA_OK(cudaMemcpyAsync(plan.ustat_std + plan.uniq_stat_off[u], f.std, f.D * sizeof(float), cudaMemcpyDeviceToDevice, stream));
    }
    // ---- normalisation: one launch per unique tensor shape class (usually one)
    for (int u = 0; u < nu;) {
        int v = u;
        while (v < nu && plan.uniq_N[v] == plan.uniq_N[u] && plan.uniq_D[v] == plan.uniq_D[u] &&
               plan.uniq_tnorm_off[v] - plan.uniq_tnorm_off[u] == (size_t)(v - u) * align_up((size_t)plan.uniq_N[u] * plan.uniq_D[u], 64))
            ++v;
        f32::NormArgs a{};
        a.traw = plan.d_uniq_ptr + u;
        a.tnorm = plan.tnorm + plan.uniq_tnorm_off[u];
        a.tnorm_stride = align_up((size_t)plan.uniq_N[u] * plan.uniq_D[u], 64);
        a.mean = plan.ustat_mean + plan.uniq_stat_off[u];
        a.stdv = plan.ustat_std + plan.uniq_stat_off[u];
        a.prenorm = plan.d_uniq_prenorm + u;
        a.N = plan.uniq_N[u]; a.D = plan.uniq_D[u];
        // stats stride must equal align_up(D,64) per tensor: norm_kernel indexes mean[u*D+d], so
        // launch per tensor when D is not a multiple of 64
        if (a.D % 64 == 0) {
            dim3 grid(ceil_div(a.D, 32), v - u);
            f32::norm_kernel<<<grid, dim3(32, 32), 0, stream>>>(a);
        } else {
            for (int w = u; w < v; ++w) {
                f32::NormArgs b = a;
                b.traw = plan.d_uniq_ptr + w;
                b.tnorm = plan.tnorm + plan.uniq_tnorm_off[w];
                b.mean = plan.ustat_mean + plan.uniq_stat_off[w];
                b.stdv = plan.ustat_std + plan.uniq_stat_off[w];
                b.prenorm = plan.d_uniq_prenorm + w;
                dim3 grid(ceil_div(b.D, 32), 1);
                f32::norm_kernel<<<grid, dim3(32, 32), 0, stream>>>(b);
            }
        }
        NA_LAUNCH_OK("norm_kernel");
        u = v;
    }
    for (Group& g : plan.groups) {
        f32::scatter_stats_kernel<<<g.nf, 128, 0, stream>>>(g.d_recs, g.D);
        NA_LAUNCH_OK("scatter_stats_kernel");
    }

    // ---- tensor path set-up: TMA descriptors + initial bf16 weight mirror
    std::vector<tc::GroupMaps> maps(plan.groups.size());
    if (precision == NA_PREC_BF16) {
        if ((rc = tc::configure_all())) return rc;
        if ((rc = chain::configure_all())) return rc;
        if ((rc = dw::configure_all())) return rc;
        if ((rc = res::configure_all())) return rc;
        for (size_t gi = 0; gi < plan.groups.size(); ++gi) {
            Group& g = plan.groups[gi];
            if (g.use_resident) continue;
            if (g.use_chain) {
                if ((rc = chain::build_maps(g.N, g.D, g.H, g.L, g.nf, g.lm, g.wbf16, g.act, g.cosb, g.dy, g.xop,
                                            g.mtiles * (int)g.pos_tabs.size(), g.cmaps))) return rc;
                NA_CUDA_OK(cudaMemcpyAsync(g.d_pos_tabs, g.pos_tabs.data(), g.pos_tabs.size() * sizeof(float*),
                                           cudaMemcpyHostToDevice, stream));
                chain::xop_kernel<<<dim3(g.mtiles, (unsigned)g.pos_tabs.size()), tc::BM, 0, stream>>>(g.d_pos_tabs, g.N, g.mtiles, g.xop);
                NA_LAUNCH_OK("xop_kernel");
                if ((rc = dw::build_maps(g.N, g.D, g.H, g.L, g.nf, g.act, g.cosb, g.dy, g.dmaps))) return rc;
            } else {
                g.maps = &maps[gi];
                if ((rc = tc::build_group_maps(g.N, g.D, g.H, g.L, g.nf, g.lm, g.act, g.cosb, g.dz, g.dy, g.wbf16, *g.maps)))
                    return rc;
            }
            tc::mirror_weights(g.d_recs, g.lm, g.nf, g.wbf16, stream);
            if (g.psc) chain::scale_params(g.d_recs, g.nf, g.H, g.L, g.psc, stream);
            NA_LAUNCH_OK("mirror_weights");
        }
    }

    // ---- epoch loop
    // One epoch of one group, all of it on stream s (eager mode, fp32, unfused tensor path, single-group calls).
    // cap: most CTAs a persistent kernel of this epoch may use (0 = all SMs) -- while fit-resident kernels hold an SM
    // each, a 148-CTA persistent grid would run as two or three ragged waves on what is left
    // several shape groups in one call compete for the SMs: launches that cannot fill both tile slots of every CTA are packed
    // onto half the CTAs (chain::launch_h); NERFATTN_NO_PACK=1 keeps one tile per CTA
    const bool contended = plan.groups.size() >= 2 && !env_flag("NERFATTN_NO_PACK");
    auto group_epoch = [&](const Group& g, cudaStream_t s, int cap) -> int {
        if (precision == NA_PREC_FP32) { fp32_epoch(g, plan, beta1, beta2, eps, s); return NA_OK; }
        if (g.use_chain) {
            int r2 = chain_part(g, cap, s, contended);
            return r2 ? r2 : update_part(g, plan, beta1, beta2, eps, cap, s);
        }
        int r2 = tc::epoch(g.N, g.D, g.H, g.L, g.nf, g.lm, g.d_recs, *g.maps, g.act, g.cosb, g.dz, g.dy, g.gradpart,
                           g.colpart, g.colpart_layer_off, g.xpart, g.losspart, g.losspart_per_fit, g.mtiles, s);
        if (r2) return r2;
        launch_adam(g, plan, beta1, beta2, eps, s);
        return NA_OK;
    };
    // groups with per-epoch kernels (eg) and fit-resident groups (rg: one launch covers a whole stretch of epochs)
    std::vector<const Group*> eg, rg;
    for (const Group& g : plan.groups) (g.use_resident ? rg : eg).push_back(&g);
    const size_t ng = eg.size();
    // `streams`: one per fit-resident group (their kernels use an SM per fit, so different groups run side by side), or
    // null: all on `s`
    auto launch_resident = [&](int e0, int count, cudaStream_t s, const cudaStream_t* streams) -> int {
        if (!(chain::phase_mask() & 16)) return NA_OK;
        for (size_t ri = 0; ri < rg.size(); ++ri) {
            const Group& g = *rg[ri];
            if (streams) s = streams[ri];
            res::ResArgs a{};
            a.N = g.N; a.mtiles = g.mtiles; a.recs = g.d_recs;
            for (int l = 0; l < 3; ++l) { a.w_off[l] = g.lm.w_off[l]; a.b_off[l] = g.lm.b_off[l]; }
            a.e_begin = e0; a.e_count = count;
            a.step_size = plan.d_step_size; a.bc2 = plan.d_bc2;
            a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.eps = (float)eps;
            a.loss_scale = 2.0f / ((float)g.N * (float)g.D); a.loss_inv_count = 1.0f / ((float)g.N * (float)g.D);
            a.sincos_mode = chain::sincos_mode();
            int r2 = res::launch(g.H, g.D, a, g.nf, s);
            if (r2) return r2;
        }
        return NA_OK;
    };
    // Two-lane schedule of the chain path (all groups on it, at least two of them): the chain kernels of all groups run
    // back to back on the "compute lane" with grid_c persistent CTAs, each group's dW + Adam follows on the "memory
    // lane" with the remaining SMs while the next group's chain kernel runs -- the chain is FP32/SFU-issue-bound with
    // HBM to spare, dW + Adam is HBM-bound with the SMs idle, and neither can share an SM with the other (shared memory).
    // Every group counts its own epochs, so inside a multi-epoch graph a group's next chain kernel only waits for that
    // group's own Adam.  NERFATTN_LANES = CTAs of the compute lane; default 0 = off (one branch per group, full grids):
    // measured on B200 (profiles/README.md, round 2) the schedule LOSES -- 1.75 ms per epoch at 108 + 40 CTAs against
    // 1.67 ms without it -- because the dW + Adam kernel is bound by the load latency of each SM (128 KB of operand
    // stages + the Adam stream in flight per SM), so its time grows as 148 / grid_m (0.54 -> 1.29 ms on 40 SMs) and
    // the memory lane becomes the bottleneck; the SM-time of the two halves is conserved.
    bool all_chain = precision == NA_PREC_BF16;
    for (const Group* g : eg) all_chain = all_chain && g->use_chain;
    int grid_c = 0;
    { const char* e = getenv("NERFATTN_LANES"); if (e) grid_c = atoi(e); }
    grid_c &= ~1;                                           // CTA pairs
    const int sms = tc::num_sms();
    const bool lanes = all_chain && ng >= 2 && grid_c >= 2 && grid_c <= sms - 8;
    const int grid_m = sms - grid_c;

    // Enqueue `count` epochs of every group behind whatever is on `main`; with a kit (stream capture) the groups fan out.
    auto record_epochs = [&](int count, cudaStream_t main, CaptureKit* kit, int cap) -> int {
        int r2 = NA_OK;
        if (!kit || ng == 1) {
            for (int e = 0; e < count && !r2; ++e)
                for (size_t gi = 0; gi < ng && !r2; ++gi) r2 = group_epoch(*eg[gi], main, cap);
        } else if (lanes) {
            cudaStream_t lc = kit->streams[0], lm = kit->streams[1];
            cudaEvent_t* ev = kit->events.data();            // [0] fork, [1] [2] joins, [3 + 2g] chain done, [4 + 2g] Adam done
            cudaEventRecord(ev[0], main);
            cudaStreamWaitEvent(lc, ev[0], 0);
            cudaStreamWaitEvent(lm, ev[0], 0);
            for (int e = 0; e < count && !r2; ++e)
                for (size_t gi = 0; gi < ng && !r2; ++gi) {
                    const Group& g = *eg[gi];
                    if (e > 0) cudaStreamWaitEvent(lc, ev[4 + 2 * gi], 0);       // this group's previous epoch has ended
                    if ((r2 = chain_part(g, grid_c, lc))) break;
                    cudaEventRecord(ev[3 + 2 * gi], lc);
                    cudaStreamWaitEvent(lm, ev[3 + 2 * gi], 0);
                    if ((r2 = update_part(g, plan, beta1, beta2, eps, grid_m, lm))) break;
                    cudaEventRecord(ev[4 + 2 * gi], lm);
                }
            cudaEventRecord(ev[1], lc); cudaStreamWaitEvent(main, ev[1], 0);
            cudaEventRecord(ev[2], lm); cudaStreamWaitEvent(main, ev[2], 0);
        } else {
            cudaEvent_t* ev = kit->events.data();            // [0] fork, [1 + g] joins
            cudaEventRecord(ev[0], main);
            for (size_t gi = 0; gi < ng; ++gi) {
                cudaStream_t sg = kit->streams[gi];
                cudaStreamWaitEvent(sg, ev[0], 0);
                for (int e = 0; e < count && !r2; ++e) r2 = group_epoch(*eg[gi], sg, cap);
                cudaEventRecord(ev[1 + gi], sg);
                cudaStreamWaitEvent(main, ev[1 + gi], 0);
            }
        }
        if (r2) return r2;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("epoch launch failed: %s", cudaGetErrorString(e)); return NA_ERR_CUDA; }
        return NA_OK;
    };

    // progress metrics (siren.py:107-115): an fp32 evaluation with the weights epoch e starts from
    auto log_progress = [&](int e) -> int {
        if (log_every <= 0 || !progress || (e + 1) % log_every) return NA_OK;
        float* dst = progress + (size_t)((e + 1) / log_every - 1) * nfits * 2;
        for (Group& g : plan.groups) {
            final_eval(g, precision, stream);
            f32::progress_copy_kernel<<<ceil_div(g.nf, 128), 128, 0, stream>>>(g.d_recs, g.nf, dst);
            NA_LAUNCH_OK("progress metrics");
        }
        return NA_OK;
    };
    const bool logging = log_every > 0 && progress;

    // epochs until the next progress evaluation (or the end): the stretch a fit-resident launch covers
    auto stretch = [&](int e) { return logging ? std::min(epochs - e, log_every - ((e + 1) % log_every)) : epochs - e; };
    if (epochs > 0) {
        if (!env_flag("NERFATTN_NO_GRAPH")) {
            // Graphs of up to `glen` epochs (NERFATTN_GRAPH_EPOCHS), replayed; a shorter one covers remainders and the
            // stretches between progress evaluations.  Replays are stream-ordered, so the lanes drain once per replay.
            int glen = 20;
            { const char* e = getenv("NERFATTN_GRAPH_EPOCHS"); if (e && atoi(e) > 0) glen = atoi(e); }
            KitLease lease(lanes ? 2 : std::max<size_t>(ng, 1), lanes ? 3 + 2 * ng : 1 + ng, rg.size());
            if (!lease.kit) { set_error("cannot create capture streams / events: %s", cudaGetErrorString(cudaGetLastError())); return NA_ERR_CUDA; }
            std::map<std::pair<int, int>, GraphHold> graphs;  // by (length, CTA cap); destroyed on every early return
            auto graph_of = [&](int len, int cap, cudaGraphExec_t* out) -> int {
                GraphHold& gh = graphs[std::make_pair(len, cap)];
                if (!gh.exec) {
                    NA_CUDA_OK(cudaStreamBeginCapture(lease.kit->cap, cudaStreamCaptureModeThreadLocal));
                    const int r2 = record_epochs(len, lease.kit->cap, lease.kit, cap);
                    cudaError_t ce = cudaStreamEndCapture(lease.kit->cap, &gh.graph);   // always ends the capture, also after an error
                    if (r2) return r2;
                    if (ce != cudaSuccess) { set_error("graph capture failed: %s", cudaGetErrorString(ce)); return NA_ERR_CUDA; }
                    ce = cudaGraphInstantiate(&gh.exec, gh.graph, 0);
                    if (ce != cudaSuccess) { set_error("graph instantiate failed: %s", cudaGetErrorString(ce)); return NA_ERR_CUDA; }
                }
                *out = gh.exec;
                return NA_OK;
            };
            int res_until = 0;                               // the fit-resident launches cover epochs [0, res_until)
            bool res_pending = false;
            // (Replaying the first epochs from graphs whose persistent grids are capped to the SMs the fit-resident CTAs leave
            // free was measured: no difference, 175.4-176.0 k fit-epochs/s for caps over 0-60 epochs -- CTAs that find
            // their SM taken simply start when one frees up.)
            for (int e = 0; e < epochs;) {
                if (res_pending && e >= res_until) {
                    for (size_t ri = 0; ri < rg.size(); ++ri) NA_CUDA_OK(cudaStreamWaitEvent(stream, lease.kit->aux_join[ri], 0));
                    res_pending = false;
                }
                if ((rc = log_progress(e))) return rc;
                if (!rg.empty() && e >= res_until) {
                    // the fit-resident groups run this whole stretch in one launch per group, beside the epoch graphs of the
                    // other groups (their CTAs take an SM each for the stretch; the graphs' persistent kernels get the rest)
                    const int len = stretch(e);
                    NA_CUDA_OK(cudaEventRecord(lease.kit->aux_fork, stream));
                    for (size_t ri = 0; ri < rg.size(); ++ri) NA_CUDA_OK(cudaStreamWaitEvent(lease.kit->aux[ri], lease.kit->aux_fork, 0));
                    if ((rc = launch_resident(e, len, stream, lease.kit->aux.data()))) return rc;
                    for (size_t ri = 0; ri < rg.size(); ++ri) NA_CUDA_OK(cudaEventRecord(lease.kit->aux_join[ri], lease.kit->aux[ri]));
                    res_pending = true;
                    res_until = e + len;
                }
                if (!ng) { e = res_until; continue; }
                int len = std::max(1, std::min(glen, stretch(e)));
                cudaGraphExec_t exec = nullptr;
                if ((rc = graph_of(len, 0, &exec))) return rc;
                cudaError_t ce = cudaGraphLaunch(exec, stream);
                if (ce != cudaSuccess) { set_error("graph launch failed: %s", cudaGetErrorString(ce)); return NA_ERR_CUDA; }
                e += len;
            }
            if (res_pending)
                for (size_t ri = 0; ri < rg.size(); ++ri) NA_CUDA_OK(cudaStreamWaitEvent(stream, lease.kit->aux_join[ri], 0));
            for (auto& kv : graphs) {                        // still running: destroyed by a later call
                if (kv.second.exec) park_graph(kv.second.exec, kv.second.graph, stream);
                kv.second.release();
            }
        } else {
            for (int e = 0; e < epochs;) {
                if ((rc = log_progress(e))) return rc;
                const int len = stretch(e);
                if ((rc = launch_resident(e, len, stream, nullptr))) return rc;
                for (int k = 0; k < len; ++k)
                    if (ng && (rc = record_epochs(1, stream, nullptr, 0))) return rc;
                e += len;
            }
        }
    }
    // ---- final metrics
    for (Group& g : plan.groups) {
        final_eval(g, precision, stream);
        NA_LAUNCH_OK("final_eval");
    }
    return NA_OK;
}

// =========================================================================== forward / decode
namespace na {

// common set-up of the inference entries: one shape for all models
struct InferPlan {
    Group g;
    float** d_out;        // device array of per-model output pointers
    float* u; float* c0;  // decode: folded query
    float* psc;           // decode, chain path: omega-prescaled W0 / biases
    float* dotpart; int nparts;
    size_t bytes;
};

static int infer_validate(const na_fit_t* m, int n, bool need_stats) {
    if (!m || n <= 0) { set_error("models must be a non-empty array"); return NA_ERR_INVALID; }
    for (int i = 0; i < n; ++i) {
        if (m[i].N != m[0].N || m[i].D != m[0].D || m[i].H != m[0].H || m[i].L != m[0].L) {
            set_error("model %d: all models of one call must share (N, D, H, L)", i); return NA_ERR_UNSUPPORTED;
        }
        if (m[i].N < 1 || m[i].H % 8 || m[i].D % 4 || m[i].L < 0 || m[i].L > kMaxHidden) { set_error("model %d: unsupported shape", i); return NA_ERR_UNSUPPORTED; }
        if (!m[i].positions || !m[i].params || (need_stats && (!m[i].mean || !m[i].std))) { set_error("model %d: null pointer", i); return NA_ERR_INVALID; }
    }
    return NA_OK;
}

static void infer_plan(const na_fit_t* m, int n, int precision, bool decode, void* ws, InferPlan& p) {
    Arena ar(ws);
    Group& g = p.g;
    g = Group{};
    g.N = m[0].N; g.D = m[0].D; g.H = m[0].H; g.L = m[0].L; g.nf = n;
    g.lm = make_layer_map(g.H, g.L, g.D);
    g.mtiles = ceil_div(g.N, 128);
    g.d_recs = ar.take<FitRec>(n);
    p.d_out = ar.take<float*>(n);
    const bool bf = precision == NA_PREC_BF16;
    const size_t nh = (size_t)n * g.N * g.H;
    g.act[0] = ar.take<char>(nh * (bf ? 2 : 4));
    g.act[1] = ar.take<char>(nh * (bf ? 2 : 4));
    g.wbf16 = bf ? ar.take<__nv_bfloat16>((size_t)n * g.lm.P) : nullptr;
    p.u = p.c0 = p.dotpart = p.psc = nullptr; p.nparts = 0;
    if (decode) {
        p.u = ar.take<float>((size_t)n * g.H);
        p.c0 = ar.take<float>(n);
        const bool dchain = bf && chain_enabled() && chain::shape_supported(g.N, g.D, g.H, g.L);
        p.nparts = dchain ? chain::decode_parts(g.H) : bf ? (g.H / tc::hidden_bn(g.H)) * 2 : ceil_div(g.H, f32::BN);
        p.dotpart = ar.take<float>((size_t)n * p.nparts * g.N);
        p.psc = dchain ? ar.take<float>((size_t)n * chain::psc_floats(g.H, g.L)) : nullptr;
    }
    p.bytes = ar.bytes();
}

static int infer_upload(const na_fit_t* m, int n, const InferPlan& p, float* const* outs, cudaStream_t stream) {
    std::vector<FitRec> recs(n);
    for (int i = 0; i < n; ++i) {
        FitRec& r = recs[i];
        memset(&r, 0, sizeof(r));
        r.pos = m[i].positions; r.params = m[i].params; r.omega = m[i].omega0;
        r.mean = m[i].mean; r.stdv = m[i].std;
    }
    NA_CUDA_OK(cudaMemcpyAsync(p.g.d_recs, recs.data(), n * sizeof(FitRec), cudaMemcpyHostToDevice, stream));
    NA_CUDA_OK(cudaMemcpyAsync(p.d_out, outs, n * sizeof(float*), cudaMemcpyHostToDevice, stream));
    return NA_OK;
}

}  // namespace na

extern "C" int nerfattn_forward_workspace_bytes(const na_fit_t* models, int32_t n, size_t* bytes) {
    if (!bytes) { set_error("bytes is null"); return NA_ERR_INVALID; }
    int rc = infer_validate(models, n, false);
    if (rc) return rc;
    InferPlan p; infer_plan(models, n, NA_PREC_FP32, false, nullptr, p);
    *bytes = p.bytes;
    return NA_OK;
}

extern "C" int nerfattn_siren_forward(const na_fit_t* models, int32_t n, int32_t denormalise, float* const* out,
                                      void* workspace, size_t workspace_bytes, na_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = infer_validate(models, n, denormalise != 0);
    if (rc) return rc;
    if (!out) { set_error("out is null"); return NA_ERR_INVALID; }
    if (!workspace || ((uintptr_t)workspace & 255)) { set_error("workspace must be a 256-byte aligned device pointer"); return NA_ERR_WORKSPACE; }
    InferPlan p; infer_plan(models, n, NA_PREC_FP32, false, workspace, p);
    if (p.bytes > workspace_bytes) { set_error("workspace too small: need %zu bytes, got %zu", p.bytes, workspace_bytes); return NA_ERR_WORKSPACE; }
    if ((rc = infer_upload(models, n, p, out, stream))) return rc;
    const Group& g = p.g;
    float* act[kMaxHidden + 1];
    for (int l = 0; l <= g.L; ++l) act[l] = (float*)g.act[l & 1];
    fp32_forward_hidden(g, act, nullptr, g.L, stream);
    fp32_output_layer(g, act[g.L], true, stream, p.d_out, denormalise);
    NA_LAUNCH_OK("siren_forward");
    return NA_OK;
}

extern "C" int nerfattn_decode_workspace_bytes(const na_fit_t* models, int32_t n, int32_t precision, size_t* bytes) {
    if (!bytes) { set_error("bytes is null"); return NA_ERR_INVALID; }
    int rc = infer_validate(models, n, true);
    if (rc) return rc;
    if (precision != NA_PREC_FP32 && precision != NA_PREC_BF16) { set_error("precision %d not implemented", precision); return NA_ERR_UNSUPPORTED; }
    InferPlan p; infer_plan(models, n, precision, true, nullptr, p);
    *bytes = p.bytes;
    return NA_OK;
}

extern "C" int nerfattn_decode_qk(const na_fit_t* models, int32_t n, const void* q_fp16, float* const* scores,
                                  int32_t precision, int32_t reuse_setup, void* workspace, size_t workspace_bytes,
                                  na_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = infer_validate(models, n, true);
    if (rc) return rc;
    if (precision != NA_PREC_FP32 && precision != NA_PREC_BF16) { set_error("precision %d not implemented", precision); return NA_ERR_UNSUPPORTED; }
    if (!q_fp16 || !scores) { set_error("null q / scores"); return NA_ERR_INVALID; }
    if (!workspace || ((uintptr_t)workspace & 255)) { set_error("workspace must be a 256-byte aligned device pointer"); return NA_ERR_WORKSPACE; }
    const bool bf = precision == NA_PREC_BF16;
    if (bf && !tc::shape_supported(models[0].N, 128, models[0].H, models[0].L) &&
        !(chain_enabled() && chain::shape_supported(models[0].N, 128, models[0].H, models[0].L))) {
        set_error("bf16 decode needs H in {64,128,256,512} and hidden_layers >= 1 (any N), or N %% 128 == 0"); return NA_ERR_UNSUPPORTED;
    }
    InferPlan p; infer_plan(models, n, precision, true, workspace, p);
    if (p.bytes > workspace_bytes) { set_error("workspace too small: need %zu bytes, got %zu", p.bytes, workspace_bytes); return NA_ERR_WORKSPACE; }
    const Group& g = p.g;
    const size_t nh = (size_t)g.N * g.H;
    if (bf && (rc = tc::configure_all())) return rc;
    if (!reuse_setup) {
        if ((rc = infer_upload(models, n, p, scores, stream))) return rc;
        if (bf) { tc::mirror_weights(g.d_recs, g.lm, n, g.wbf16, stream); NA_LAUNCH_OK("mirror_weights"); }
        if (p.psc) { chain::scale_params(g.d_recs, n, g.H, g.L, p.psc, stream); NA_LAUNCH_OK("scale_params"); }
    }
    dec::decode_prep_kernel<<<n, 256, 0, stream>>>(g.d_recs, (const __half*)q_fp16, g.D, g.H, g.lm.w_off[g.L + 1],
                                                   g.lm.b_off[g.L + 1], p.u, p.c0);
    NA_LAUNCH_OK("decode_prep_kernel");
    if (g.L == 0) {
        dec::l0dot_kernel<<<dim3(ceil_div(g.N, 8), n), 256, 0, stream>>>(g.d_recs, g.N, g.H, p.u, p.c0, p.d_out);
        NA_LAUNCH_OK("l0dot_kernel");
        return NA_OK;
    }
    if (!bf) {
        float* act[kMaxHidden + 1];
        for (int l = 0; l <= g.L; ++l) act[l] = (float*)g.act[l & 1];
        fp32_forward_hidden(g, act, nullptr, g.L - 1, stream);
        f32::GemmArgs a; base_args(a, g);
        a.M = g.N; a.N = g.H; a.K = g.H;
        a.A = act[g.L - 1]; a.a_fit = nh; a.lda = g.H;
        a.b_param_off = g.lm.w_off[g.L]; a.ldb = g.H;
        a.bias_off = g.lm.b_off[g.L];
        a.dotvec = p.u; a.dotvec_fit = g.H;
        a.dotpart = p.dotpart; a.dotpart_fit = (size_t)p.nparts * g.N;
        launch_sgemm<f32::kFwdDot, true, true>(a, n, 1, stream);
    } else if (chain_enabled() && chain::shape_supported(g.N, g.D, g.H, g.L)) {
        // fused forward chain: layer 0, the hidden layers and the u . sin(.) reduction in one kernel
        if ((rc = chain::configure_all())) return rc;
        chain::ChainMaps cm;
        if ((rc = chain::build_fwd_maps(g.N, g.H, g.L, n, g.lm, g.wbf16, cm))) return rc;
        if ((rc = chain::launch_decode(g.N, g.D, g.H, g.L, n, g.lm, g.d_recs, cm, p.psc, p.u, p.dotpart, stream))) return rc;
    } else {
        const int bn = tc::hidden_bn(g.H);
        {
            const size_t total8 = nh / 8;
            dim3 grid((unsigned)ceil_div(total8, (size_t)256), n);
            f32::layer0_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(g.d_recs, g.N, g.H, (__nv_bfloat16*)g.act[0], nullptr, nh);
        }
        tc::TcArgs base{};
        base.nb = n; base.recs = g.d_recs;
        for (int l = 1; l <= g.L; ++l) {
            tc::GemmMaps maps;
            if ((rc = tc::make_operand_map(&maps.a, g.act[(l - 1) & 1], g.N, g.H, n, nh, false, tc::BM))) return rc;
            if ((rc = tc::make_operand_map(&maps.b, g.wbf16 + g.lm.w_off[l], g.H, g.H, n, g.lm.P, false, bn))) return rc;
            tc::TcArgs a = base;
            a.M = g.N; a.N = g.H; a.K = g.H; a.bias_off = g.lm.b_off[l];
            if (l < g.L) {
                a.out0 = (__nv_bfloat16*)g.act[l & 1]; a.out0_fit = nh; a.out1 = nullptr;
                if ((rc = tc::launch_bn<tc::kFwdSine, false, false>(bn, maps, a, stream))) return rc;
            } else {
                a.dotvec = p.u; a.dotvec_fit = g.H;
                a.dotpart = p.dotpart; a.dotpart_fit = (size_t)p.nparts * g.N;
                if ((rc = tc::launch_bn<tc::kFwdDot, false, false>(bn, maps, a, stream))) return rc;
            }
        }
    }
    dec::decode_finish_kernel<<<dim3(ceil_div(g.N, 256), n), 256, 0, stream>>>(p.dotpart, p.nparts, g.N, p.c0, p.d_out);
    NA_LAUNCH_OK("decode");
    return NA_OK;
}

extern "C" int nerfattn_kvread_qk(const void* k_fp16, const void* q_fp16, float* scores, int32_t n, int32_t N,
                                  int32_t D, na_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!k_fp16 || !q_fp16 || !scores || n <= 0 || N <= 0) { set_error("bad argument"); return NA_ERR_INVALID; }
    if (D != 64 && D != 128 && D != 256) { set_error("kvread: D must be 64, 128 or 256"); return NA_ERR_UNSUPPORTED; }
    const long long rows = (long long)n * N;
    const int G = D / 16;                                        // lanes per row, 32 B each
    const long long groups_needed = (rows + 3) / 4;             // 4 rows per lane group and batch
    long long blocks = (groups_needed * G + 255) / 256;
    // persistent grid: as many CTAs as are resident at once (registers: two batches of 4 x 32 B in flight per thread)
    int occ = 0;
    if (G == 4) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dec::kvread_qk_kernel<4>, 256, 0);
    else if (G == 8) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dec::kvread_qk_kernel<8>, 256, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, dec::kvread_qk_kernel<16>, 256, 0);
    const long long cap = (long long)tc::num_sms() * std::max(occ, 1);
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    const uint32_t* K = (const uint32_t*)k_fp16; const uint32_t* q = (const uint32_t*)q_fp16;
    if (G == 4) dec::kvread_qk_kernel<4><<<(unsigned)blocks, 256, 0, stream>>>(K, q, scores, rows, N);
    else if (G == 8) dec::kvread_qk_kernel<8><<<(unsigned)blocks, 256, 0, stream>>>(K, q, scores, rows, N);
    else dec::kvread_qk_kernel<16><<<(unsigned)blocks, 256, 0, stream>>>(K, q, scores, rows, N);
    NA_LAUNCH_OK("kvread_qk_kernel");
    return NA_OK;
}

// =========================================================================== attention decode (softmax, P.V)
namespace na {
struct PvPlan {
    Group g;
    float* partial;        // kvread / fp32: [n][chunks][D];  bf16 chain: [n][N/32][H]
    int chunks;
    float* vfull;          // fp32 mode: reconstructed values [n][N][D]
    float** d_out;         // fp32 mode: per-model output pointers into vfull
    bool chain;
    float* psc;
    size_t bytes;
};
static void pv_plan(const na_fit_t* m, int n, int N, int D, int precision, void* ws, PvPlan& p) {
    Arena ar(ws);
    p = PvPlan{};
    p.chunks = ceil_div(N, dec::kPvChunk);
    if (!m) {                                           // KV-read baseline
        p.partial = ar.take<float>((size_t)n * p.chunks * D);
        p.bytes = ar.bytes();
        return;
    }
    Group& g = p.g;
    g.N = m[0].N; g.D = m[0].D; g.H = m[0].H; g.L = m[0].L; g.nf = n;
    g.lm = make_layer_map(g.H, g.L, g.D);
    g.mtiles = ceil_div(g.N, 128);
    p.chunks = ceil_div(g.N, dec::kPvChunk);
    g.d_recs = ar.take<FitRec>(n);
    p.chain = precision == NA_PREC_BF16 && chain_enabled() && chain::shape_supported(g.N, g.D, g.H, g.L);
    if (p.chain) {
        g.wbf16 = ar.take<__nv_bfloat16>((size_t)n * g.lm.P);
        p.partial = ar.take<float>((size_t)n * (g.mtiles * 4) * g.H);
        p.psc = ar.take<float>((size_t)n * chain::psc_floats(g.H, g.L));
    } else {
        const size_t nh = (size_t)n * g.N * g.H;
        g.act[0] = ar.take<char>(nh * 4);
        g.act[1] = ar.take<char>(nh * 4);
        p.vfull = ar.take<float>((size_t)n * g.N * g.D);
        p.d_out = ar.take<float*>(n);
        p.partial = ar.take<float>((size_t)n * p.chunks * g.D);
    }
    p.bytes = ar.bytes();
}
}  // namespace na

extern "C" int nerfattn_softmax(float* scores, int32_t n, int32_t N, float scale, na_stream_t stream_) {
    if (!scores || n <= 0 || N <= 0) { set_error("bad argument"); return NA_ERR_INVALID; }
    dec::softmax_kernel<<<n, 1024, 0, (cudaStream_t)stream_>>>(scores, N, scale);
    NA_LAUNCH_OK("softmax_kernel");
    return NA_OK;
}

extern "C" int nerfattn_pv_workspace_bytes(const na_fit_t* models, int32_t n, int32_t N, int32_t D, int32_t precision,
                                           size_t* bytes) {
    if (!bytes || n <= 0) { set_error("bad argument"); return NA_ERR_INVALID; }
    if (models) {
        int rc = infer_validate(models, n, true);
        if (rc) return rc;
        N = models[0].N; D = models[0].D;
    }
    if (N <= 0 || D <= 0 || D % 8) { set_error("pv: D must be a positive multiple of 8"); return NA_ERR_UNSUPPORTED; }
    PvPlan p; pv_plan(models, n, N, D, precision, nullptr, p);
    *bytes = p.bytes;
    return NA_OK;
}

extern "C" int nerfattn_kvread_pv(const void* v_fp16, const float* p, float* out, int32_t n, int32_t N, int32_t D,
                                  void* workspace, size_t workspace_bytes, na_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!v_fp16 || !p || !out || n <= 0 || N <= 0) { set_error("bad argument"); return NA_ERR_INVALID; }
    if (D <= 0 || D % 8 || 256 % (D / 8)) { set_error("pv: D must be 8, 16, .., 2048 with D/8 dividing 256"); return NA_ERR_UNSUPPORTED; }
    if (!workspace || ((uintptr_t)workspace & 255)) { set_error("workspace must be a 256-byte aligned device pointer"); return NA_ERR_WORKSPACE; }
    PvPlan pl; pv_plan(nullptr, n, N, D, 0, workspace, pl);
    if (pl.bytes > workspace_bytes) { set_error("workspace too small: need %zu bytes, got %zu", pl.bytes, workspace_bytes); return NA_ERR_WORKSPACE; }
    const size_t smem = (size_t)(256 / (D / 8)) * D * sizeof(float);
    dec::pv_kernel<__half><<<dim3(pl.chunks, n), 256, smem, stream>>>((const __half*)v_fp16, p, pl.partial, N, D, pl.chunks);
    dec::pv_finish_kernel<<<n, 128, 0, stream>>>(pl.partial, pl.chunks, D, out);
    NA_LAUNCH_OK("kvread_pv");
    return NA_OK;
}

extern "C" int nerfattn_decode_pv(const na_fit_t* models, int32_t n, const float* p, float* out, int32_t precision,
                                  void* workspace, size_t workspace_bytes, na_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = infer_validate(models, n, true);
    if (rc) return rc;
    if (precision != NA_PREC_FP32 && precision != NA_PREC_BF16) { set_error("precision %d not implemented", precision); return NA_ERR_UNSUPPORTED; }
    if (!p || !out) { set_error("null p / out"); return NA_ERR_INVALID; }
    if (!workspace || ((uintptr_t)workspace & 255)) { set_error("workspace must be a 256-byte aligned device pointer"); return NA_ERR_WORKSPACE; }
    const int D = models[0].D;
    if (D % 8 || 256 % (D / 8)) { set_error("pv: D/8 must divide 256"); return NA_ERR_UNSUPPORTED; }
    PvPlan pl; pv_plan(models, n, 0, 0, precision, workspace, pl);
    if (pl.bytes > workspace_bytes) { set_error("workspace too small: need %zu bytes, got %zu", pl.bytes, workspace_bytes); return NA_ERR_WORKSPACE; }
    const Group& g = pl.g;
    {   // model table
        std::vector<FitRec> recs(n);
        std::vector<float*> outs(n);
        for (int i = 0; i < n; ++i) {
            FitRec& r = recs[i];
            memset(&r, 0, sizeof(r));
            r.pos = models[i].positions; r.params = models[i].params; r.omega = models[i].omega0;
            r.mean = models[i].mean; r.stdv = models[i].std;
            if (!pl.chain) outs[i] = pl.vfull + (size_t)i * g.N * g.D;
        }
        NA_CUDA_OK(cudaMemcpyAsync(g.d_recs, recs.data(), n * sizeof(FitRec), cudaMemcpyHostToDevice, stream));
        if (!pl.chain) NA_CUDA_OK(cudaMemcpyAsync(pl.d_out, outs.data(), n * sizeof(float*), cudaMemcpyHostToDevice, stream));
    }
    if (pl.chain) {
        // V never exists: the forward chain reduces p_t * h_L(t) over the positions, the output layer acts on the sum
        if ((rc = tc::configure_all()) || (rc = chain::configure_all())) return rc;
        tc::mirror_weights(g.d_recs, g.lm, n, g.wbf16, stream);
        chain::scale_params(g.d_recs, n, g.H, g.L, pl.psc, stream);
        NA_LAUNCH_OK("mirror_weights");
        chain::ChainMaps cm;
        if ((rc = chain::build_fwd_maps(g.N, g.H, g.L, n, g.lm, g.wbf16, cm))) return rc;
        if ((rc = chain::launch_decode(g.N, g.D, g.H, g.L, n, g.lm, g.d_recs, cm, pl.psc, nullptr, nullptr, stream, p, pl.partial))) return rc;
        dec::attn_finish_kernel<<<n, 256, g.H * sizeof(float), stream>>>(g.d_recs, pl.partial, g.mtiles * 4, g.H, g.D,
                                                                         g.lm.w_off[g.L + 1], g.lm.b_off[g.L + 1], out);
        NA_LAUNCH_OK("attn_finish_kernel");
        return NA_OK;
    }
    // fp32 (and shapes outside the tensor path): reconstruct V in fp32, then the same P.V kernel as the baseline
    float* act[kMaxHidden + 1];
    for (int l = 0; l <= g.L; ++l) act[l] = (float*)g.act[l & 1];
    fp32_forward_hidden(g, act, nullptr, g.L, stream);
    fp32_output_layer(g, act[g.L], true, stream, pl.d_out, 1);
    const size_t smem = (size_t)(256 / (D / 8)) * D * sizeof(float);
    dec::pv_kernel<float><<<dim3(pl.chunks, n), 256, smem, stream>>>(pl.vfull, p, pl.partial, g.N, D, pl.chunks);
    dec::pv_finish_kernel<<<n, 128, 0, stream>>>(pl.partial, pl.chunks, D, out);
    NA_LAUNCH_OK("decode_pv");
    return NA_OK;
}

extern "C" int nerfattn_debug_gemm_bf16(const void* a_bf16, const void* b_bf16, float* c, int32_t M, int32_t N,
                                        int32_t K, int32_t batch, int32_t a_mn, int32_t b_mn, na_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!a_bf16 || !b_bf16 || !c || M <= 0 || batch <= 0) { set_error("bad argument"); return NA_ERR_INVALID; }
    if (N % 64 || K % 64 || N <= 0 || K <= 0 || (a_mn && M % 8)) { set_error("debug gemm: N and K must be multiples of 64"); return NA_ERR_UNSUPPORTED; }
    int rc = tc::configure_all();
    if (rc) return rc;
    const int bn = (N % 256 == 0) ? 256 : (N % 128 == 0) ? 128 : 64;
    tc::GemmMaps maps;
    if (a_mn) rc = tc::make_operand_map(&maps.a, a_bf16, K, M, batch, (size_t)K * M, true, 0);
    else rc = tc::make_operand_map(&maps.a, a_bf16, M, K, batch, (size_t)M * K, false, tc::BM);
    if (rc) return rc;
    if (b_mn) rc = tc::make_operand_map(&maps.b, b_bf16, K, N, batch, (size_t)K * N, true, 0);
    else rc = tc::make_operand_map(&maps.b, b_bf16, N, K, batch, (size_t)N * K, false, bn);
    if (rc) return rc;
    tc::TcArgs a{};
    a.M = M; a.N = N; a.K = K; a.nb = batch;
    a.fout = c; a.fout_fit = (size_t)M * N; a.fout_off = 0; a.ldf = N;
    if (!a_mn && !b_mn) return tc::launch_bn<tc::kRaw, false, false>(bn, maps, a, stream);
    if (!a_mn && b_mn) return tc::launch_bn<tc::kRaw, false, true>(bn, maps, a, stream);
    if (a_mn && !b_mn) return tc::launch_bn<tc::kRaw, true, false>(bn, maps, a, stream);
    return tc::launch_bn<tc::kRaw, true, true>(bn, maps, a, stream);
}

namespace na {
__global__ void debug_sincos_kernel(const float* x, float* s, float* c, long long n, int mode) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a[1] = {x[i]}, sv[1], cv[1];
    if (mode == 1) sincos_group_mufu(a, sv, cv); else sincos_group(a, sv, cv);
    s[i] = sv[0]; c[i] = cv[0];
}
}  // namespace na

// ---------------------------------------------------------------------------
// synthetic KV generator (synth.cuh; reference nerf_attention/extract.py:182-259)
namespace na {
struct SynthPlan { synth::Stream* d_streams; float** d_keys; float** d_values; float* staging; size_t bytes; };
static void synth_plan(int nstreams, int N, int D, void* ws, SynthPlan& p) {
    Arena ar(ws);
    p.d_streams = ar.take<synth::Stream>(nstreams);
    p.d_keys = ar.take<float*>(nstreams);
    p.d_values = ar.take<float*>(nstreams);
    p.staging = ar.take<float>((size_t)nstreams * 2 * N * D);
    p.bytes = ar.bytes();
}
}  // namespace na

extern "C" int nerfattn_synth_workspace_bytes(int32_t nstreams, int32_t N, int32_t D, size_t* bytes) {
    if (!bytes || nstreams <= 0 || N <= 0 || D <= 0) { set_error("bad argument"); return NA_ERR_INVALID; }
    SynthPlan p; synth_plan(nstreams, N, D, nullptr, p);
    *bytes = p.bytes;
    return NA_OK;
}

extern "C" int nerfattn_synth_kv(const na_synth_stream_t* streams, int32_t nstreams, int32_t N, int32_t D,
                                 const float* positions, void* workspace, size_t workspace_bytes, na_stream_t stream_) {
    if (!streams || !positions || nstreams <= 0 || N <= 0 || D <= 0) { set_error("bad argument"); return NA_ERR_INVALID; }
    cudaStream_t stream = (cudaStream_t)stream_;
    SynthPlan p; synth_plan(nstreams, N, D, workspace, p);
    if (!workspace || workspace_bytes < p.bytes) { set_error("workspace too small: need %zu bytes", p.bytes); return NA_ERR_WORKSPACE; }
    std::vector<synth::Stream> hs(nstreams);
    std::vector<float*> hk(nstreams), hv(nstreams);
    for (int i = 0; i < nstreams; ++i) {
        const na_synth_stream_t& st = streams[i];
        if (!st.keys || !st.values) { set_error("stream %d: null output", i); return NA_ERR_INVALID; }
        if (st.n_spikes < 0 || st.n_spikes > synth::kMaxSpikes || st.max_width < 2) {
            set_error("stream %d: n_spikes must be in [0, %d] and max_width >= 2", i, synth::kMaxSpikes);
            return NA_ERR_UNSUPPORTED;
        }
        hs[i].seed = st.seed; hs[i].n_spikes = st.n_spikes; hs[i].max_width = st.max_width;
        hs[i].keys_t = p.staging + (size_t)(2 * i) * N * D;
        hs[i].values_t = p.staging + (size_t)(2 * i + 1) * N * D;
        hk[i] = st.keys; hv[i] = st.values;
    }
    NA_CUDA_OK(cudaMemcpyAsync(p.d_streams, hs.data(), nstreams * sizeof(synth::Stream), cudaMemcpyHostToDevice, stream));
    NA_CUDA_OK(cudaMemcpyAsync(p.d_keys, hk.data(), nstreams * sizeof(float*), cudaMemcpyHostToDevice, stream));
    NA_CUDA_OK(cudaMemcpyAsync(p.d_values, hv.data(), nstreams * sizeof(float*), cudaMemcpyHostToDevice, stream));
    synth::synth_kv_kernel<<<nstreams, synth::kThreads, 0, stream>>>(p.d_streams, positions, N, D);
    NA_LAUNCH_OK("synth_kv_kernel");
    synth::TransposeArgs ta{p.d_streams, p.d_keys, p.d_values, N, D};
    synth::transpose_kernel<<<dim3(ceil_div(N, 32), ceil_div(D, 32), 2 * nstreams), dim3(32, 8), 0, stream>>>(ta);
    NA_LAUNCH_OK("synth transpose_kernel");
    return NA_OK;
}

extern "C" int nerfattn_debug_sincos(const float* x, float* s, float* c, int64_t n, int32_t mode, na_stream_t stream_) {
    if (!x || !s || !c || n <= 0 || (mode != 0 && mode != 1)) { set_error("bad argument"); return NA_ERR_INVALID; }
    debug_sincos_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream_>>>(x, s, c, (long long)n, mode);
    NA_LAUNCH_OK("debug_sincos_kernel");
    return NA_OK;
}
