// Fit-resident training kernel for the narrow SIRENs of the sweep (`tiny` H = 64 and `small` H = 128, one hidden
// layer; reference types.py:93-94): ONE persistent CTA per fit, ONE launch for all epochs.
//
// The row-tile chain (siren_chain.cuh) + grouped dW / Adam kernel (siren_dw.cuh) stream every activation of an epoch
// through L2 / HBM between the two kernels and take at least three launches per group and epoch; for these two
// architectures that machinery ran at 0.05 / 0.11 of the tensor roofline (profiles/README.md, round 1): 640 tiles on
// 296 tile slots, five-step chains, per-step weight streams.  Here everything a fit needs between two Adam steps stays
// on the SM that owns the fit:
//
//   shared memory   the bf16 weights W1 [H x H], Wf [D x H] in the 128B-swizzled layout tcgen05 reads -- K-major for
//                   the forward layers and, the same bytes, MN-major for the backward ones; the activations of the
//                   current 128-row tile (h0, h1, dY, cos1 -> dz1 in place, dz0 over h0) as A operands; the layer-0
//                   gradient operand of the tile; the omega-prescaled layer-0 weights and biases
//   tensor memory   the step accumulator (128 columns) and, across all row tiles of the epoch, the gradient
//                   accumulators dWf, dW1 (operands MN-major straight from the activation buffers), dbf, db1 (ones
//                   products) and dW0 / db0 (positions operand): 176 + 2H of the 512 columns
//   global memory   per epoch only the fit's normalised targets (N x D fp32, the one stream from HBM), and the fp32
//                   master weights and Adam moments once in the Adam phase (3 x 4P B read + written; 399 KB for
//                   `small` does not fit next to the tile buffers, so the optimiser state is the one thing that is not
//                   SM-resident; it is 22 MB for all 80 such fits of the sweep and stays in L2)
//
// Warp roles: warp 1 issues every MMA, warps 4..19 are the epilogue (row = TMEM lane, four column groups) and the
// Adam phase; two mbarriers (operand written / accumulator complete) alternate strictly, one tile in flight.
// Numerics are those of the chain path (bf16 operands, fp32 accumulation, fp32 layer 0, torch-order Adam): the
// weight gradients accumulate over the rows in the same order as dw::dw_adam_kernel, i.e. bit for bit.
#pragma once

#include "siren_chain.cuh"

namespace na {
namespace res {

using namespace tc;
using chain::make_idesc_m;
using chain::st_shared_128;
using chain::tmem_ld16;
using chain::XOP_BYTES;
using chain::XOP_N;

constexpr int NCTRL = 4, NEPI = 16, NTHREADS = (NCTRL + NEPI) * 32;
constexpr int CHUNK = BM * 128;                    // one 64-column chunk of an activation buffer: 128 rows x 128 B

inline bool shape_supported(int N, int D, int H, int L) {
    return N >= 1 && L == 1 && (H == 64 || H == 128) && (D == 64 || D == 128);
}

template <int H, int D> struct Cfg {
    static constexpr int W1_BYTES = H * H * 2, WF_BYTES = D * H * 2;
    static constexpr int AH_BYTES = BM * H * 2, AD_BYTES = BM * D * 2;
    static constexpr int OFF_W1 = 0;
    static constexpr int OFF_WF = OFF_W1 + W1_BYTES;
    static constexpr int OFF_H0 = OFF_WF + WF_BYTES;       // h0, later dz0
    static constexpr int OFF_H1 = OFF_H0 + AH_BYTES;
    static constexpr int OFF_DY = OFF_H1 + AH_BYTES;
    static constexpr int OFF_C1 = OFF_DY + AD_BYTES;       // cos1, later dz1 (in place)
    // H = 64: the dW1 MMA reads dz1^T as an M = 128 operand, i.e. one 16 KB chunk past the 64 real columns (rows 64..127 of
    // the product are never read back); the pad keeps that read inside the allocation
    static constexpr int OFF_XOP = OFF_C1 + (H == 64 ? 2 * AH_BYTES : AH_BYTES);
    static constexpr int OFF_ONES = OFF_XOP + XOP_BYTES;
    static constexpr int OFF_VEC = OFF_ONES + ONES_BYTES;  // fp32: w0s[H] b0s[H] b1s[H] bfs[D] red[NEPI]
    static constexpr int VEC_FLOATS = 3 * H + D + NEPI;
    static constexpr int OFF_BAR = (OFF_VEC + VEC_FLOATS * 4 + 63) / 64 * 64;
    static constexpr int SMEM = OFF_BAR + 64 + 1024;       // + alignment slack
    // tensor memory columns
    static constexpr int T_ACC = 0;                         // step accumulator, max(H, D) <= 128 columns
    static constexpr int T_DWF = 128, T_DBF = T_DWF + H, T_DW1 = T_DBF + 16, T_DB1 = T_DW1 + H, T_L0 = T_DB1 + 16;
    static_assert(T_L0 + 16 <= 512, "tensor memory");
    static_assert(SMEM <= 227 * 1024, "shared memory");
};

struct ResArgs {
    int N, mtiles;
    const FitRec* recs;                   // one CTA per record
    int w_off[3], b_off[3];               // layer 0, hidden layer, output layer inside the packed parameter vector
    int e_begin, e_count;                 // epochs [e_begin, e_begin + e_count) of the tables / of losses[]
    const float* step_size; const float* bc2;
    float beta1, beta2, eps;
    float loss_scale, loss_inv_count;     // 2 / (N D), 1 / (N D)
    int sincos_mode;
};

// 16 bf16 (two 16-byte units) of row r, columns [col, col + 16) of a [rows x 64k] operand stored as 64-column chunks
// of `chunk_bytes` (rows x 128 B, 128B swizzle) -- activations (chunk = 16 KB) and weights (chunk = rows x 128 B) alike
__device__ __forceinline__ void store16(uint32_t base, uint32_t chunk_bytes, int r, int col, const uint32_t (&pk)[8]) {
    const uint32_t rowaddr = base + (uint32_t)(col >> 6) * chunk_bytes + (uint32_t)r * 128;
    const int u0 = (col & 63) >> 3;
    st_shared_128(rowaddr + (uint32_t)((u0 ^ (r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
    st_shared_128(rowaddr + (uint32_t)(((u0 + 1) ^ (r & 7)) << 4), pk[4], pk[5], pk[6], pk[7]);
}
__device__ __forceinline__ void load16(uint32_t base, uint32_t chunk_bytes, int r, int col, uint32_t (&pk)[8]) {
    const uint32_t rowaddr = base + (uint32_t)(col >> 6) * chunk_bytes + (uint32_t)r * 128;
    const int u0 = (col & 63) >> 3;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(pk[0]), "=r"(pk[1]), "=r"(pk[2]), "=r"(pk[3])
                 : "r"(rowaddr + (uint32_t)((u0 ^ (r & 7)) << 4)) : "memory");
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(pk[4]), "=r"(pk[5]), "=r"(pk[6]), "=r"(pk[7])
                 : "r"(rowaddr + (uint32_t)(((u0 + 1) ^ (r & 7)) << 4)) : "memory");
}
// Adam for the fit-resident kernel: the same update as f32::adam_update with the two IEEE divisions and the IEEE square
// root replaced by the SFU approximations (sqrt.approx, rcp.approx: <= 2 ulp each) and 1 / bc2 folded into a multiply.
// The exact sequences are ~100 issue slots per parameter and were 19 % of this kernel's stall samples; the deviation
// (~1e-7 relative per step) is three orders of magnitude below the bf16 rounding of the gradient itself, and the fp32
// parity mode never runs this kernel.
__device__ __forceinline__ void adam_fast(float g, float& m, float& v, float& w, float ob1, float beta2, float ob2,
                                          float eps, float inv_bc2, float nss) {
    m = fmaf(ob1, g - m, m);
    v = fmaf(v, beta2, (ob2 * g) * g);
    float sq;
    asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(v));
    float rc;
    asm("rcp.approx.f32 %0, %1;" : "=f"(rc) : "f"(fmaf(sq, inv_bc2, eps)));
    w = fmaf(nss * m, rc, w);
}
__device__ __forceinline__ void bar_all() { asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory"); }

template <int H, int D>
__global__ void __launch_bounds__(NTHREADS, 1)
resident_kernel(const ResArgs g) {
    using C = Cfg<H, D>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t s_w1 = smem_u32(smem + C::OFF_W1), s_wf = smem_u32(smem + C::OFF_WF);
    const uint32_t s_h0 = smem_u32(smem + C::OFF_H0), s_h1 = smem_u32(smem + C::OFF_H1);
    const uint32_t s_dy = smem_u32(smem + C::OFF_DY), s_c1 = smem_u32(smem + C::OFF_C1);
    const uint32_t s_xop = smem_u32(smem + C::OFF_XOP), s_ones = smem_u32(smem + C::OFF_ONES);
    float* vec = reinterpret_cast<float*>(smem + C::OFF_VEC);
    float* w0s = vec; float* b0s = vec + H; float* b1s = vec + 2 * H; float* bfs = vec + 3 * H; float* red = vec + 3 * H + D;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* act_ready = bars; uint64_t* acc_full = bars + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const FitRec& rec = g.recs[blockIdx.x];
    const float omega = rec.omega;
    const int mtiles = g.mtiles;

    // ---------------------------------------------------------------- set-up
    if (threadIdx.x == 0) {
        mbar_init(act_ready, NEPI); mbar_init(acc_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += NTHREADS) reinterpret_cast<uint32_t*>(smem + C::OFF_ONES)[i] = 0x3F803F80u;
    // bf16 copies of the master weights in the MMA layout: 8 consecutive k per 16-byte unit
    auto stage_weights = [&](const float* w, int rows, uint32_t base) {
        for (int i = threadIdx.x; i < rows * (H / 8); i += NTHREADS) {
            const int n = i / (H / 8), k0 = (i - n * (H / 8)) * 8;
            const float4 a = *reinterpret_cast<const float4*>(w + (size_t)n * H + k0);
            const float4 b = *reinterpret_cast<const float4*>(w + (size_t)n * H + k0 + 4);
            st_shared_128(base + (uint32_t)(k0 >> 6) * (uint32_t)(rows * 128) + (uint32_t)n * 128 + (uint32_t)((((k0 & 63) >> 3) ^ (n & 7)) << 4),
                          pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
        }
    };
    stage_weights(rec.params + g.w_off[1], H, s_w1);
    stage_weights(rec.params + g.w_off[2], D, s_wf);
    for (int j = threadIdx.x; j < H; j += NTHREADS) {
        w0s[j] = omega * rec.params[g.w_off[0] + j];
        b0s[j] = omega * rec.params[g.b_off[0] + j];
        b1s[j] = omega * rec.params[g.b_off[1] + j];
    }
    for (int j = threadIdx.x; j < D; j += NTHREADS) bfs[j] = rec.params[g.b_off[2] + j];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < NCTRL) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        if (warp == 1) {
            // ===================================================== MMA issuer
            uint32_t rdy = 0;
            const uint32_t i_fh = make_idesc(H, false, false), i_fd = make_idesc(D, false, false);   // forward: K-major B
            const uint32_t i_bh = make_idesc(H, false, true);                                         // backward: MN-major B
            const uint32_t i_dw = make_idesc(H, true, true), i_one = make_idesc(16, true, false), i_x = make_idesc(XOP_N, true, false);
            auto wait_ready = [&] { mbar_wait(act_ready, rdy); rdy ^= 1u; tc_fence_after(); };
            // descriptors: the start-address field (bytes >> 4) is the only part that changes inside a step
            auto kmaj = [](uint32_t base) { return make_desc(base, 0, 1024); };                       // K-major, chunk-advanced by hand
            const uint64_t a_h0 = kmaj(s_h0), a_h1 = kmaj(s_h1), a_dy = kmaj(s_dy), a_c1 = kmaj(s_c1);
            const uint64_t b_w1 = kmaj(s_w1), b_wf = kmaj(s_wf), b_one = kmaj(s_ones), b_x = kmaj(s_xop);
            const uint64_t bt_wf = make_desc(s_wf, D * 128, 1024), bt_w1 = make_desc(s_w1, H * 128, 1024);   // MN-major weights (backward)
            const uint64_t t_dy = make_desc(s_dy, CHUNK, 1024), t_h1 = make_desc(s_h1, CHUNK, 1024);       // MN-major activations (dW)
            const uint64_t t_c1 = make_desc(s_c1, CHUNK, 1024), t_h0 = make_desc(s_h0, CHUNK, 1024);
            auto koff = [](int k, int chunk_bytes) { return (uint64_t)(((k >> 2) * chunk_bytes + (k & 3) * 32) >> 4); };   // K-major: 16 columns = 32 B
            constexpr uint64_t ROWS16 = (UMMA_K * 128) >> 4;                                               // MN-major: 16 rows of 128 B
            for (int e = 0; e < g.e_count; ++e) {
                for (int t = 0; t < mtiles; ++t) {
                    const uint32_t accum = t > 0 ? 1u : 0u;             // gradient accumulators: fresh at the first tile of an epoch
                    // step 1: z1 = h0 W1^T
                    wait_ready();
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < H / UMMA_K; ++k)
                            tc_mma_bf16(tmem_base + C::T_ACC, a_h0 + koff(k, CHUNK), b_w1 + koff(k, H * 128), i_fh, k > 0 ? 1u : 0u);
                        tc_commit(acc_full);
                    }
                    __syncwarp();
                    // step 2: y = h1 Wf^T
                    wait_ready();
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < H / UMMA_K; ++k)
                            tc_mma_bf16(tmem_base + C::T_ACC, a_h1 + koff(k, CHUNK), b_wf + koff(k, D * 128), i_fd, k > 0 ? 1u : 0u);
                        tc_commit(acc_full);
                    }
                    __syncwarp();
                    // step 3: dh1 = dY Wf -- committed at once: the epilogue turns it into dz1 (in place over cos1, which no
                    // MMA of this step reads) while dWf += dY^T h1 and dbf += dY^T 1 run behind it; the tensor pipe is in
                    // order, so they are complete before anything of step 4 starts
                    wait_ready();
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < D / UMMA_K; ++k)
                            tc_mma_bf16(tmem_base + C::T_ACC, a_dy + koff(k, CHUNK), bt_wf + k * ROWS16, i_bh, k > 0 ? 1u : 0u);
                        tc_commit(acc_full);
#pragma unroll
                        for (int k = 0; k < BM / UMMA_K; ++k) {
                            tc_mma_bf16(tmem_base + C::T_DWF, t_dy + k * ROWS16, t_h1 + k * ROWS16, i_dw, (accum | (uint32_t)(k > 0)));
                            tc_mma_bf16(tmem_base + C::T_DBF, t_dy + k * ROWS16, b_one + (uint64_t)((k & 3) * 2), i_one, (accum | (uint32_t)(k > 0)));
                        }
                    }
                    __syncwarp();
                    // step 4: dW1 += dz1^T h0, db1 += dz1^T 1, dh0 = dz1 W1: one commit for all three -- the epilogue
                    // overwrites h0 with dz0, so it may only start once dW1 has read h0
                    wait_ready();
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < BM / UMMA_K; ++k) {
                            tc_mma_bf16(tmem_base + C::T_DW1, t_c1 + k * ROWS16, t_h0 + k * ROWS16, i_dw, (accum | (uint32_t)(k > 0)));
                            tc_mma_bf16(tmem_base + C::T_DB1, t_c1 + k * ROWS16, b_one + (uint64_t)((k & 3) * 2), i_one, (accum | (uint32_t)(k > 0)));
                        }
#pragma unroll
                        for (int k = 0; k < H / UMMA_K; ++k)
                            tc_mma_bf16(tmem_base + C::T_ACC, a_c1 + koff(k, CHUNK), bt_w1 + k * ROWS16, i_bh, k > 0 ? 1u : 0u);
                        tc_commit(acc_full);
                    }
                    __syncwarp();
                    // step 5: layer-0 gradient: {sum dz0, sum dz0 x} += dz0^T {1, x_hi, x_mid, x_lo}
                    wait_ready();
                    if (lane == 0) {
#pragma unroll
                        for (int k = 0; k < BM / UMMA_K; ++k)
                            tc_mma_bf16(tmem_base + C::T_L0, t_h0 + k * ROWS16, b_x + koff(k, XOP_BYTES / 2), i_x, (accum | (uint32_t)(k > 0)));
                        tc_commit(acc_full);
                    }
                    __syncwarp();
                }
                bar_all();            // the Adam phase has rewritten the weight buffers (and fenced them for the async proxy)
            }
        } else {
            for (int e = 0; e < g.e_count; ++e) bar_all();
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ===================================================== epilogue warps + Adam phase
        const int ei = warp - NCTRL;
        const int q = warp & 3, cg = ei >> 2;
        const int r = q * 32 + lane;                                 // row inside the tile = TMEM lane
        constexpr int CW = H / 4, NU = CW / 16;                      // hidden columns per thread, 16-column units
        constexpr int OW = D / 4, NUO = OW / 16;                     // output columns per thread
        const int col0 = cg * CW, ocol0 = cg * OW;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool mufu_hidden = (g.sincos_mode & 1) != 0, mufu_l0 = (g.sincos_mode & 2) != 0;
        uint32_t full = 0;
        auto wait_acc = [&] { mbar_wait(acc_full, full); full ^= 1u; tc_fence_after(); };
        auto hand_over = [&] {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(act_ready);
        };
        auto sincos16 = [&](const float (&arg)[16], float (&sn)[16], float (&cs)[16], bool mufu) {
#pragma unroll
            for (int gi = 0; gi < 2; ++gi) {
                float a8[8], s8[8], c8[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) a8[j] = arg[gi * 8 + j];
                if (mufu) chain::sincos8<true>(a8, s8, c8); else chain::sincos8<false>(a8, s8, c8);
#pragma unroll
                for (int j = 0; j < 8; ++j) { sn[gi * 8 + j] = s8[j]; cs[gi * 8 + j] = c8[j]; }
            }
        };
        for (int e = 0; e < g.e_count; ++e) {
            const int ee = g.e_begin + e;
            float sq = 0.f;
            for (int t = 0; t < mtiles; ++t) {
                const int row = t * BM + r;
                const bool row_ok = row < g.N;
                const int row_c = row_ok ? row : g.N - 1;
                const float x = __ldg(rec.pos + row_c);
                // targets of this thread (OW fp32): pulled into L2 now, loaded right before the OUT step waits for its MMA
                const float* tn = rec.tnorm + (size_t)row_c * D + ocol0;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(tn));
                // ---------------- E0: layer 0 (fp32, siren.py:33-34 with in_features = 1).  The math runs while the previous
                // tile's layer-0 gradient MMA still reads dz0 out of this buffer; only the stores wait for it.
                uint32_t so0[NU][8];
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    float arg[16], sn[16], cs[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) arg[j] = fmaf(x, w0s[col0 + u * 16 + j], b0s[col0 + u * 16 + j]);
                    sincos16(arg, sn, cs, mufu_l0);
#pragma unroll
                    for (int j = 0; j < 16; j += 2) so0[u][j / 2] = pack_bf16(sn[j], sn[j + 1]);
                }
                if (t > 0) wait_acc();
                if (cg == 0) {                                       // the tile's positions as the B operand {1, x_hi, x_mid, x_lo}
                    const float xv = row_ok ? x : 0.f;
                    const __nv_bfloat16 hi = __float2bfloat16_rn(xv);
                    const float r1 = xv - __bfloat162float(hi);
                    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
                    const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
                    const __nv_bfloat16 vals[4] = {__float2bfloat16_rn(row_ok ? 1.f : 0.f), hi, mid, lo};
                    __nv_bfloat16* xo = reinterpret_cast<__nv_bfloat16*>(smem + C::OFF_XOP);
#pragma unroll
                    for (int j = 0; j < XOP_N; ++j) {
                        const int off = (r >> 6) * (XOP_BYTES / 2) + j * 128 + ((((r & 63) >> 3) ^ (j & 7)) << 4) + (r & 7) * 2;
                        xo[off / 2] = j < 4 ? vals[j] : __float2bfloat16_rn(0.f);
                    }
                }
#pragma unroll
                for (int u = 0; u < NU; ++u) store16(s_h0, CHUNK, r, col0 + u * 16, so0[u]);
                hand_over();
                // ---------------- S1: h1 = sin(w z1 + w b1), cos1 parked in shared memory
                wait_acc();
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    uint32_t v[16];
                    tmem_ld16(t_lane + C::T_ACC + col0 + u * 16, v);
                    tmem_ld_wait();
                    float arg[16], sn[16], cs[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) arg[j] = fmaf(__uint_as_float(v[j]), omega, b1s[col0 + u * 16 + j]);
                    sincos16(arg, sn, cs, mufu_hidden);
                    uint32_t so[8], co[8];
#pragma unroll
                    for (int j = 0; j < 16; j += 2) { so[j / 2] = pack_bf16(sn[j], sn[j + 1]); co[j / 2] = pack_bf16(cs[j], cs[j + 1]); }
                    store16(s_h1, CHUNK, r, col0 + u * 16, so);
                    store16(s_c1, CHUNK, r, col0 + u * 16, co);
                }
                hand_over();
                // ---------------- OUT: dY = 2 (y - t) / (N D), loss (siren.py:101)
                uint32_t tg[OW];
#pragma unroll
                for (int j = 0; j < OW; j += 8) chain::ld_global_nc_na_256(tn + j, &tg[j]);
                wait_acc();
                {
                    const float rmask = row_ok ? 1.f : 0.f;
#pragma unroll
                    for (int u = 0; u < NUO; ++u) {
                        uint32_t v[16];
                        tmem_ld16(t_lane + C::T_ACC + ocol0 + u * 16, v);
                        tmem_ld_wait();
                        uint32_t dout[8];
#pragma unroll
                        for (int j = 0; j < 16; j += 2) {
                            const float e0 = rmask * ((__uint_as_float(v[j]) + bfs[ocol0 + u * 16 + j]) - __uint_as_float(tg[u * 16 + j]));
                            const float e1 = rmask * ((__uint_as_float(v[j + 1]) + bfs[ocol0 + u * 16 + j + 1]) - __uint_as_float(tg[u * 16 + j + 1]));
                            sq = fmaf(e0, e0, sq); sq = fmaf(e1, e1, sq);
                            dout[j / 2] = pack_bf16(e0 * g.loss_scale, e1 * g.loss_scale);
                        }
                        store16(s_dy, CHUNK, r, ocol0 + u * 16, dout);
                    }
                }
                hand_over();
                // ---------------- DXF: dz1 = (dY Wf) * w cos1, in place over cos1 (read while the MMA runs)
                uint32_t cc1[NU][8];
#pragma unroll
                for (int u = 0; u < NU; ++u) load16(s_c1, CHUNK, r, col0 + u * 16, cc1[u]);
                wait_acc();
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    uint32_t v[16], dout[8];
                    uint32_t (&cc)[8] = cc1[u];
                    tmem_ld16(t_lane + C::T_ACC + col0 + u * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float c0, c1;
                        unpack_bf16(cc[j], c0, c1);
                        dout[j] = pack_bf16(__uint_as_float(v[2 * j]) * (omega * c0), __uint_as_float(v[2 * j + 1]) * (omega * c1));
                    }
                    store16(s_c1, CHUNK, r, col0 + u * 16, dout);
                }
                hand_over();
                // ---------------- DX1: dz0 = (dz1 W1) * w cos0, cos0 recomputed as in E0 (while the MMAs run); over h0 (dW1 has read it)
                float wc0[NU][16];
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    float arg[16], sn[16], cs[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) arg[j] = fmaf(x, w0s[col0 + u * 16 + j], b0s[col0 + u * 16 + j]);
                    sincos16(arg, sn, cs, mufu_l0);
#pragma unroll
                    for (int j = 0; j < 16; ++j) wc0[u][j] = omega * cs[j];
                }
                wait_acc();
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    uint32_t v[16], dout[8];
                    tmem_ld16(t_lane + C::T_ACC + col0 + u * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dout[j] = pack_bf16(__uint_as_float(v[2 * j]) * wc0[u][2 * j], __uint_as_float(v[2 * j + 1]) * wc0[u][2 * j + 1]);
                    store16(s_h0, CHUNK, r, col0 + u * 16, dout);
                }
                hand_over();
            }
            wait_acc();                                              // the last tile's layer-0 gradient MMA: every accumulator is complete

            // ---------------- Adam (torch _single_tensor_adam order) straight from the TMEM accumulators
            const float bc2 = 1.0f / g.bc2[ee], nss = -g.step_size[ee];      // bc2: the reciprocal (adam_fast)
            const float ob1 = 1.0f - g.beta1, ob2 = 1.0f - g.beta2;
            float* pw = rec.params; float* pm = rec.m; float* pv = rec.v;
            // weights of a [rows x H] layer: this thread owns row r, columns [col0, col0 + CW)
            auto adam_rows = [&](int t_col, int rows, int w_off, uint32_t s_w) {
                if (r >= rows) return;
#pragma unroll
                for (int u = 0; u < NU; ++u) {
                    uint32_t v[16];
                    tmem_ld16(t_lane + t_col + col0 + u * 16, v);
                    tmem_ld_wait();
                    const size_t o = (size_t)w_off + (size_t)r * H + col0 + u * 16;
                    uint32_t nb[8];
                    float4 w4[4], m4[4], v4[4];                      // 12 loads in flight before the first store (no aliasing stalls)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        w4[j] = *reinterpret_cast<const float4*>(pw + o + 4 * j);
                        m4[j] = *reinterpret_cast<const float4*>(pm + o + 4 * j);
                        v4[j] = *reinterpret_cast<const float4*>(pv + o + 4 * j);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        adam_fast(__uint_as_float(v[4 * j]), m4[j].x, v4[j].x, w4[j].x, ob1, g.beta2, ob2, g.eps, bc2, nss);
                        adam_fast(__uint_as_float(v[4 * j + 1]), m4[j].y, v4[j].y, w4[j].y, ob1, g.beta2, ob2, g.eps, bc2, nss);
                        adam_fast(__uint_as_float(v[4 * j + 2]), m4[j].z, v4[j].z, w4[j].z, ob1, g.beta2, ob2, g.eps, bc2, nss);
                        adam_fast(__uint_as_float(v[4 * j + 3]), m4[j].w, v4[j].w, w4[j].w, ob1, g.beta2, ob2, g.eps, bc2, nss);
                        *reinterpret_cast<float4*>(pm + o + 4 * j) = m4[j];
                        *reinterpret_cast<float4*>(pv + o + 4 * j) = v4[j];
                        *reinterpret_cast<float4*>(pw + o + 4 * j) = w4[j];
                        nb[2 * j] = pack_bf16(w4[j].x, w4[j].y); nb[2 * j + 1] = pack_bf16(w4[j].z, w4[j].w);
                    }
                    store16(s_w, (uint32_t)(rows * 128), r, col0 + u * 16, nb);
                }
            };
            adam_rows(C::T_DWF, D, g.w_off[2], s_wf);
            adam_rows(C::T_DW1, H, g.w_off[1], s_w1);
            // vectors: one parameter per lane (TMEM lane = feature), spread over the column groups
            auto adam_one = [&](float gr, size_t pi) {
                float mm = pm[pi], vv = pv[pi], ww = pw[pi];
                adam_fast(gr, mm, vv, ww, ob1, g.beta2, ob2, g.eps, bc2, nss);
                pm[pi] = mm; pv[pi] = vv; pw[pi] = ww;
                return ww;
            };
            if (cg == 0 && r < D) {
                const uint32_t gv = tmem_ld1(t_lane + C::T_DBF); tmem_ld_wait();
                bfs[r] = adam_one(__uint_as_float(gv), (size_t)g.b_off[2] + r);
            }
            if (cg == 1 && r < H) {
                const uint32_t gv = tmem_ld1(t_lane + C::T_DB1); tmem_ld_wait();
                b1s[r] = omega * adam_one(__uint_as_float(gv), (size_t)g.b_off[1] + r);
            }
            if (cg == 2 && r < H) {
                uint32_t v[16];
                tmem_ld16(t_lane + C::T_L0, v); tmem_ld_wait();
                const float gw = (__uint_as_float(v[1]) + __uint_as_float(v[2])) + __uint_as_float(v[3]);
                w0s[r] = omega * adam_one(gw, (size_t)g.w_off[0] + r);
                b0s[r] = omega * adam_one(__uint_as_float(v[0]), (size_t)g.b_off[0] + r);
            }
            // loss of this epoch (siren.py:105): fixed-order sum of the 16 warp partials
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
            if (lane == 0) red[ei] = sq;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // new weights -> visible to the tensor core
            tc_fence_before();
            bar_all();
            if (ei == 0 && lane == 0) {
                float s = 0.f;
                for (int i = 0; i < NEPI; ++i) s += red[i];
                rec.losses[ee] = s * g.loss_inv_count;
            }
            // red[] is rewritten only after the next epoch's bar_all, by which time thread 0 has read it: the next
            // write happens after a full epoch of tiles, each with its own barriers
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// ------------------------------------------------------------------ host
template <int H, int D>
inline cudaError_t launch_hd(const ResArgs& a, int nf, cudaStream_t s) {
    resident_kernel<H, D><<<nf, NTHREADS, Cfg<H, D>::SMEM, s>>>(a);
    return cudaGetLastError();
}
inline int launch(int H, int D, const ResArgs& a, int nf, cudaStream_t s) {
    cudaError_t e = (H == 64 && D == 64) ? launch_hd<64, 64>(a, nf, s) : (H == 64) ? launch_hd<64, 128>(a, nf, s)
                    : (D == 64) ? launch_hd<128, 64>(a, nf, s) : launch_hd<128, 128>(a, nf, s);
    if (e != cudaSuccess) { set_error("resident_kernel launch failed: %s", cudaGetErrorString(e)); return NA_ERR_CUDA; }
    return NA_OK;
}
inline int configure_all() {
    static std::once_flag once_dev[kMaxDevices];
    static cudaError_t err_dev[kMaxDevices] = {};
    const int dev = current_device();
    cudaError_t& err = err_dev[dev];
    std::call_once(once_dev[dev], [&err] {
        auto acc = [&](cudaError_t e) { if (e != cudaSuccess && err == cudaSuccess) err = e; };
        acc(cudaFuncSetAttribute(resident_kernel<64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64, 64>::SMEM));
        acc(cudaFuncSetAttribute(resident_kernel<64, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64, 128>::SMEM));
        acc(cudaFuncSetAttribute(resident_kernel<128, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128, 64>::SMEM));
        acc(cudaFuncSetAttribute(resident_kernel<128, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128, 128>::SMEM));
    });
    if (err != cudaSuccess) { set_error("cudaFuncSetAttribute(resident smem) failed: %s", cudaGetErrorString(err)); return NA_ERR_CUDA; }
    return NA_OK;
}

}  // namespace res
}  // namespace na
