// Weight gradients + Adam of the BF16 chain path in ONE grouped kernel (sm_100a).
//
//   dW_l = dz_l^T h_{l-1}   (tcgen05, both operands MN-major straight from the chain kernel's bf16 stores,
//   db_l = dz_l^T 1          contraction over the N rows of a fit, accumulators in TMEM)
//   W_l, b_l  <-  torch _single_tensor_adam(W_l, dW_l, m, v)            (reference siren.py:102-103)
//
// for every layer l = 1..L+1 of every fit of a shape group: one launch per group and epoch instead of
// L+1 GEMM launches + a gradient round trip through HBM + a separate Adam pass.  The gradient tile never
// leaves the SM: the epilogue warps pull it from TMEM into a per-warp shared-memory staging tile
// (row = lane), then walk the tile ROW-WISE so that the fp32 master weights and both Adam moments are
// read and written in 256-byte contiguous segments (the first attempt updated them from the TMEM layout
// directly -- one row per lane, 32 B per lane at row stride -- and was slower than a separate coalesced
// Adam kernel).  Two accumulator stages: the Adam stream of tile i (HBM-bound, 26 B per parameter)
// overlaps the MMAs of tile i+1 (L2-bound operand fetch).  Fixed K order, no split-K, no atomics:
// bitwise deterministic and bitwise equal to "dW GEMM, then adam_kernel".
#pragma once

#include "siren_tc.cuh"

namespace na {
namespace dw {

using namespace tc;

constexpr int NEPI = 8;                           // epilogue warps 2..9
constexpr int NTHREADS = 64 + NEPI * 32;          // 320

template <int BN> struct Cfg {
    static constexpr int A_STAGE = BM * BK * 2;                  // 16 KB: 128 (M) x 64 (K) bf16, two 64 x 64 boxes
    static constexpr int B_STAGE = BN * BK * 2;
    static constexpr int STAGE = A_STAGE + B_STAGE;
    static constexpr int CW = BN / 2;                            // gradient columns per epilogue warp
    static constexpr int STG_LD = CW + 4;                        // floats per staged row (+4: conflict-free 128-bit rows)
    static constexpr int STAGING = NEPI * 32 * STG_LD * 4;       // bytes
    static constexpr int STAGES = (BN == 128) ? 4 : 6;
    static constexpr int ACC_STRIDE = BN + 32;                   // + the bias-gradient columns (ones product)
    static constexpr int TMEM_COLS = (2 * ACC_STRIDE <= 256) ? 256 : 512;
    static constexpr int SMEM = STAGES * STAGE + ONES_BYTES + STAGING + 256 + 1024;
};

struct DwArgs {
    int N, H, L, nf;                      // rows (= K of the contraction), hidden width, hidden layers, fits of this launch
    const FitRec* recs;
    int w_off[kMaxLayers], b_off[kMaxLayers], out_dim[kMaxLayers];
    int tile_start[kMaxLayers + 1];       // [l] first tile of layer l inside one fit, l = 1..L+1; [L+2] = tiles_per_fit
    int tiles_per_fit, n_tiles;           // n_tiles = H / BN
    const int* epoch; const float* step_size; const float* bc2;   // device epoch counter + per-epoch tables
    float beta1, beta2, eps;
    __nv_bfloat16* wbf16; size_t wbf16_fit;       // bf16 mirror the MMAs of the next epoch read
    float* psc; size_t psc_fit;                   // omega-prescaled sine-layer biases the chain kernel reads
};
// a[l]: dz_l (l <= L) / dY (l = L+1) as MN-major A operand, b[l]: h_{l-1} as MN-major B operand; 64 x 64 boxes
struct DwMaps { CUtensorMap a[kMaxHidden + 2]; CUtensorMap b[kMaxHidden + 2]; };

template <int BN>
__global__ void __launch_bounds__(NTHREADS, 1)
dw_adam_kernel(const __grid_constant__ DwMaps maps, const DwArgs g) {
    using C = Cfg<BN>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * C::A_STAGE;
    uint8_t* smem_ones = smem + STAGES * C::STAGE;
    float* staging = reinterpret_cast<float*>(smem_ones + ONES_BYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(staging) + C::STAGING);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = ceil_div(g.N, BK);                 // a ragged last block is zero-filled by TMA
    const int total_tiles = g.nf * g.tiles_per_fit;

    struct Tile { int f, l, mt, nt, M; };
    auto decode = [&](int tile) {
        Tile t;
        t.f = tile / g.tiles_per_fit;
        const int r = tile - t.f * g.tiles_per_fit;
        t.l = 1;
#pragma unroll 1
        for (int l = 2; l <= g.L + 1; ++l) if (r >= g.tile_start[l]) t.l = l;
        const int w = r - g.tile_start[t.l];
        t.mt = w / g.n_tiles; t.nt = w - t.mt * g.n_tiles;
        t.M = g.out_dim[t.l];
        return t;
    };

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], NEPI); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < ONES_BYTES / 4; i += NTHREADS)            // 16 x 64 bf16 ones
        reinterpret_cast<uint32_t*>(smem_ones)[i] = 0x3F803F80u;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)C::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================================================== TMA producer
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const Tile t = decode(tile);
                const int a_boxes = min(2, ceil_div(t.M - t.mt * BM, 64));
                const uint32_t tx_bytes = a_boxes * 8192 + C::B_STAGE;
                const CUtensorMap* ma = &maps.a[t.l];
                const CUtensorMap* mb = &maps.b[t.l];
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], tx_bytes);
                    uint8_t* sa = smem_a + stage * C::A_STAGE;
                    uint8_t* sb = smem_b + stage * C::B_STAGE;
                    for (int i = 0; i < a_boxes; ++i) tma_load_3d(sa + i * 8192, ma, &full[stage], t.mt * BM + i * 64, kb * BK, t.f);
#pragma unroll
                    for (int i = 0; i < BN / 64; ++i) tma_load_3d(sb + i * 8192, mb, &full[stage], t.nt * BN + i * 64, kb * BK, t.f);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        constexpr uint32_t idesc = make_idesc(BN, true, true);
        constexpr uint32_t idesc_ones = make_idesc(16, true, false);
        constexpr uint32_t KADV = (UMMA_K * 128) >> 4;              // MN-major: 16 K rows x 128 B per UMMA_K
        int stage = 0; uint32_t phase = 0; int iter = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
            const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
            mbar_wait(&tmem_empty[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * C::ACC_STRIDE;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t adesc0 = make_desc(smem_u32(smem_a + stage * C::A_STAGE), 8192, 1024);
                    const uint64_t bdesc0 = make_desc(smem_u32(smem_b + stage * C::B_STAGE), 8192, 1024);
                    const uint64_t odesc0 = make_desc(smem_u32(smem_ones), 0, 1024);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;
                        tc_mma_bf16(d_tmem, adesc0 + (uint64_t)(k * KADV), bdesc0 + (uint64_t)(k * KADV), idesc, acc);
                        tc_mma_bf16(d_tmem + BN, adesc0 + (uint64_t)(k * KADV), odesc0 + (uint64_t)(k * 2), idesc_ones, acc);
                    }
                    tc_commit(&empty[stage]);
                    if (kb == num_kb - 1) tc_commit(&tmem_full[as]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================================================== epilogue warps: TMEM -> staging -> row-wise Adam
        constexpr int CW = C::CW, LD = C::STG_LD;
        constexpr int LPR = CW / 4;                   // lanes per row in the row-wise pass (one float4 each)
        constexpr int RPI = 32 / LPR;                 // rows per warp instruction
        constexpr int UNR = 4;                        // row groups in flight: 3 * UNR 128-bit loads per lane
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;             // which half of the BN columns
        float* stg = staging + (size_t)(warp - 2) * 32 * LD;
        const int sub = lane / LPR, cl = (lane % LPR) * 4;
        const float ob1 = 1.0f - g.beta1, ob2 = 1.0f - g.beta2;
        int iter = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
            const Tile t = decode(tile);
            const int as = iter & 1; const uint32_t aphase = (iter >> 1) & 1;
            const FitRec* rec = &g.recs[t.f];
            const int e = *g.epoch;
            const float bc2 = g.bc2[e], nss = -g.step_size[e];
            const int row0 = t.mt * BM + q * 32;                   // first output row (= out feature) of this warp
            const int rows_ok = min(32, t.M - row0);               // <= 0: nothing to do (M = 64 tiles)
            float* pw = rec->params; float* pm = rec->m; float* pv = rec->v;

            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + as * C::ACC_STRIDE + ((uint32_t)(q * 32) << 16);
#pragma unroll
            for (int c = 0; c < CW / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(t_row + half * CW + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<uint4*>(stg + lane * LD + c * 32 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            uint32_t dbv = 0;
            const bool do_bias = half == 0 && t.nt == 0;
            if (do_bias) { dbv = tmem_ld1(t_row + BN); tmem_ld_wait(); }   // every column of the ones product equals db[row]
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);           // the MMAs of tile i+2 may overwrite the accumulator

            if (do_bias && lane < rows_ok) {
                const size_t pi = (size_t)g.b_off[t.l] + row0 + lane;
                float mm = pm[pi], vv = pv[pi], ww = pw[pi];
                f32::adam_update(__uint_as_float(dbv), mm, vv, ww, ob1, g.beta2, ob2, g.eps, bc2, nss);
                pm[pi] = mm; pv[pi] = vv; pw[pi] = ww;
                if (g.psc && t.l <= g.L) g.psc[(size_t)t.f * g.psc_fit + (size_t)(t.l + 1) * g.H + row0 + lane] = rec->omega * ww;
            }
            if (rows_ok > 0) {
                const size_t col = (size_t)t.nt * BN + half * CW + cl;
                const size_t base = (size_t)g.w_off[t.l] + (size_t)row0 * g.H + col;
                __nv_bfloat16* pb = g.wbf16 + (size_t)t.f * g.wbf16_fit;
#pragma unroll 1
                for (int i0 = 0; i0 < 32; i0 += RPI * UNR) {
                    float4 w4[UNR], m4[UNR], v4[UNR], g4[UNR];
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        const int r = i0 + u * RPI + sub;
                        if (r < rows_ok) {
                            const size_t o = base + (size_t)r * g.H;
                            w4[u] = *reinterpret_cast<const float4*>(pw + o);
                            m4[u] = *reinterpret_cast<const float4*>(pm + o);
                            v4[u] = *reinterpret_cast<const float4*>(pv + o);
                            g4[u] = *reinterpret_cast<const float4*>(stg + r * LD + cl);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < UNR; ++u) {
                        const int r = i0 + u * RPI + sub;
                        if (r < rows_ok) {
                            const size_t o = base + (size_t)r * g.H;
                            f32::adam_update(g4[u].x, m4[u].x, v4[u].x, w4[u].x, ob1, g.beta2, ob2, g.eps, bc2, nss);
                            f32::adam_update(g4[u].y, m4[u].y, v4[u].y, w4[u].y, ob1, g.beta2, ob2, g.eps, bc2, nss);
                            f32::adam_update(g4[u].z, m4[u].z, v4[u].z, w4[u].z, ob1, g.beta2, ob2, g.eps, bc2, nss);
                            f32::adam_update(g4[u].w, m4[u].w, v4[u].w, w4[u].w, ob1, g.beta2, ob2, g.eps, bc2, nss);
                            *reinterpret_cast<float4*>(pm + o) = m4[u];
                            *reinterpret_cast<float4*>(pv + o) = v4[u];
                            *reinterpret_cast<float4*>(pw + o) = w4[u];
                            *reinterpret_cast<uint2*>(pb + o) = make_uint2(pack_bf16(w4[u].x, w4[u].y), pack_bf16(w4[u].z, w4[u].w));
                        }
                    }
                }
            }
            __syncwarp();                                          // staging is rewritten by the next tile
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ host
inline int bn_for(int H) { return H >= 128 ? 128 : 64; }

inline int build_maps(int N, int D, int H, int L, int nf, void* const* act, void* const* dzs, void* dy, DwMaps& m) {
    int rc;
    for (int l = 1; l <= L + 1; ++l) {
        const int width = (l == L + 1) ? D : H;
        const void* dzl = (l == L + 1) ? dy : dzs[l];
        if ((rc = make_operand_map(&m.a[l], dzl, N, width, nf, (size_t)N * width, true, 0))) return rc;
        if ((rc = make_operand_map(&m.b[l], act[l - 1], N, H, nf, (size_t)N * H, true, 0))) return rc;
    }
    return NA_OK;
}

inline void fill_args(DwArgs& a, int N, int D, int H, int L, const LayerMap& lm) {
    a.N = N; a.H = H; a.L = L;
    const int bn = bn_for(H);
    a.n_tiles = H / bn;
    int t = 0;
    for (int l = 0; l <= L + 1; ++l) { a.w_off[l] = lm.w_off[l]; a.b_off[l] = lm.b_off[l]; a.out_dim[l] = lm.out_dim[l]; }
    a.tile_start[0] = 0;
    for (int l = 1; l <= L + 1; ++l) { a.tile_start[l] = t; t += ceil_div(lm.out_dim[l], BM) * a.n_tiles; }
    a.tile_start[L + 2] = t;
    a.tiles_per_fit = t;
    (void)D;
}

inline int launch(const DwMaps& maps, const DwArgs& a, int max_ctas, cudaStream_t s) {
    const int tiles = a.nf * a.tiles_per_fit;
    const int grid = std::min(tiles, (max_ctas > 0) ? std::min(max_ctas, num_sms()) : num_sms());
    if (bn_for(a.H) == 128) dw_adam_kernel<128><<<grid, NTHREADS, Cfg<128>::SMEM, s>>>(maps, a);
    else dw_adam_kernel<64><<<grid, NTHREADS, Cfg<64>::SMEM, s>>>(maps, a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("dw_adam_kernel launch failed: %s", cudaGetErrorString(e)); return NA_ERR_CUDA; }
    return NA_OK;
}

inline int configure_all() {
    static std::once_flag once_dev[kMaxDevices];
    static cudaError_t err_dev[kMaxDevices] = {};
    const int dev = current_device();
    cudaError_t& err = err_dev[dev];
    std::call_once(once_dev[dev], [&err] {
        err = cudaFuncSetAttribute(dw_adam_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<128>::SMEM);
        if (err == cudaSuccess)
            err = cudaFuncSetAttribute(dw_adam_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<64>::SMEM);
    });
    if (err != cudaSuccess) { set_error("cudaFuncSetAttribute(dw smem) failed: %s", cudaGetErrorString(err)); return NA_ERR_CUDA; }
    return NA_OK;
}

}  // namespace dw
}  // namespace na
