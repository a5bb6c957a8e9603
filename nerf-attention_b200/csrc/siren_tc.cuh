// BF16 tensor-core path of the SIREN fit for sm_100a: TMA-staged operands in
// 128B-swizzled shared memory, tcgen05.mma (cta_group::1, M=128) accumulating in
// TMEM, warp-specialised persistent CTAs (1 TMA warp, 1 MMA warp, 8 epilogue
// warps) with a double-buffered accumulator so that the sine/cosine epilogue of
// tile i overlaps the MMAs of tile i+1.
//
// One kernel template serves the four GEMM shapes of a training step:
//   kFwdSine  act_l   = sin(w (act_{l-1} W_l^T + b_l)), cos_l        A K-major,  B K-major
//   kFwdOut   dY      = 2 (act_L Wf^T + bf - t) / (N D), loss        A K-major,  B K-major
//   kDx       dz_{l-1}= (dz_l W_l) * w cos_{l-1}                     A K-major,  B MN-major
//   kDw       dW_l    = dz_l^T act_{l-1},  db_l = dz_l^T 1           A MN-major, B MN-major
// (reference call sites: siren.py:33-34,100-102).  Layer 0
// (K = 1) and its gradient, the loss reduction and Adam stay fp32 SIMT.
#pragma once

#include <cuda.h>

#include <stdlib.h>

#include <algorithm>
#include <mutex>

#include "common.cuh"
#include "siren_fp32.cuh"

namespace na {
namespace tc {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;                    // warps 2..9
constexpr int NTHREADS = 64 + NUM_EPI_WARPS * 32;   // 320
constexpr int A_STAGE_BYTES = BM * BK * 2;          // 16 KB
constexpr int ONES_BYTES = 16 * BK * 2;             // 2 KB, B operand of the bias-gradient MMA

constexpr int kDefaultOcc2Mask = 0;     // tuned on B200, see profiles/README.md
enum Mode { kRaw = 0, kFwdSine = 1, kFwdOut = 2, kDx = 3, kDw = 4, kFwdDot = 5 };

// OCC = CTAs per SM the kernel is sized for.  OCC 1: deep smem ring, double-buffered accumulator,
// 16-wide sincos groups.  OCC 2: <= 113 KB smem and <= 256 TMEM columns per CTA so that two CTAs
// (possibly of two different kernels: a sincos-bound forward and an HBM-bound backward of another
// group) share an SM and fill each other's stalls.
template <int BN, int OCC> struct Cfg {
    static constexpr int B_STAGE_BYTES = BN * BK * 2;
    static constexpr int STAGES = (OCC == 1) ? ((BN == 256) ? 4 : 6) : ((BN == 256) ? 2 : (BN == 128) ? 3 : 4);
    static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
};
// barriers (2*STAGES + 4) * 8 B + tmem slot, rounded up; + 1 KB slack for the 1024 B alignment
template <int BN, int OCC> constexpr int smem_bytes() {
    return Cfg<BN, OCC>::STAGES * Cfg<BN, OCC>::STAGE_BYTES + ONES_BYTES + 256 + 1024;
}

struct TcArgs {
    int M, N, K;            // per-fit GEMM extents
    int nb, m_tiles, n_tiles;
    const FitRec* recs;
    int bias_off;                                         // fwd: params offset of the layer bias
    __nv_bfloat16* out0; size_t out0_fit;                 // act_l | dY | dz_{l-1}
    __nv_bfloat16* out1; size_t out1_fit;                 // cos_l
    const __nv_bfloat16* cprev; size_t cprev_fit;         // kDx: cos_{l-1}
    float* fout; size_t fout_fit; int fout_off; int ldf;  // kRaw: C; kDw: gradpart + w_off
    float* biasgrad; size_t biasgrad_fit;                 // kDw: db_l [nb][ksplits][M]
    int ksplits; size_t fout_split;                       // kDw: split-K (partials [ksplits][...], summed by Adam)
    float* losspart; int losspart_per_fit;                // kFwdOut
    float loss_scale;
    const float* dotvec; size_t dotvec_fit;               // kFwdDot: u [N] per fit (fp32)
    float* dotpart; size_t dotpart_fit;                   // kFwdDot: [n_tiles*2][M] partial row sums
};

inline bool shape_supported(int N, int D, int H, int L) {
    (void)L;
    return N % 128 == 0 && (H == 64 || H == 128 || H == 256 || H == 512) && (D == 64 || D == 128 || D == 256);
}
inline int out_bn(int D) { return D >= 256 ? 256 : (D >= 128 ? 128 : 64); }
inline int loss_partials_per_fit(int N, int D) {
    return (N / BM) * ceil_div(D, out_bn(D)) * NUM_EPI_WARPS;
}

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trap (launch failure), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {
            printf("nerfattn: mbarrier timeout block %d thread %d\n", (int)blockIdx.x, (int)threadIdx.x);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
    return v;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

// shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version field = 1
// (cute/arch/mma_sm100_desc.hpp: start[0,14) LBO[16,30) SBO[32,46) version[46,48) layout[61,64))
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor kind::f16: D=F32, A=B=BF16, M=128
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// 256-bit global accesses (LDG/STG.E.ENL2.256 on sm_100): one full 32 B sector per lane, so the
// row-per-thread epilogue never issues partial-sector writes and needs half the LSU instructions
__device__ __forceinline__ void st_global_256(void* p, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void ld_global_nc_256(const void* p, uint32_t* v) {
    asm volatile("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
}

__device__ __forceinline__ void ld_global_256(const void* p, uint32_t* v) {      // coherent: data this kernel also writes
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack_bf16(uint32_t u, float& a, float& b) {
    a = __uint_as_float(u << 16);
    b = __uint_as_float(u & 0xFFFF0000u);
}

// ------------------------------------------------------------------ the kernel
template <int MODE, bool A_MN, bool B_MN, int BN, int OCC>
__global__ void __launch_bounds__(NTHREADS, OCC)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const TcArgs g) {
    using C = Cfg<BN, OCC>;
    constexpr int STAGES = C::STAGES;
    constexpr int ACC_STRIDE = (MODE == kDw) ? BN + 32 : BN;       // TMEM columns per accumulator stage
    constexpr int TMEM_BUDGET = 512 / OCC;
    constexpr int NACC = (2 * ACC_STRIDE <= TMEM_BUDGET) ? 2 : 1;  // accumulator stages
    constexpr int TMEM_NEED = NACC * ACC_STRIDE;
    constexpr int TMEM_COLS = (TMEM_NEED <= 32) ? 32 : (TMEM_NEED <= 64) ? 64 :
                              (TMEM_NEED <= 128) ? 128 : (TMEM_NEED <= 256) ? 256 : 512;
    static_assert(TMEM_COLS <= TMEM_BUDGET, "accumulators do not fit the TMEM share of this CTA");
    static_assert(BN == 64 || BN == 128 || BN == 256, "BN");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
    uint8_t* smem_ones = smem + STAGES * C::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_ones + ONES_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + STAGES;
    uint64_t* tmem_full = bars + 2 * STAGES;
    uint64_t* tmem_empty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // split-K (kDw of small groups: one output tile per fit would leave most SMs idle): tile index
    // = (fit, m tile, n tile, K slice), partial results go to separate slots and are summed by Adam
    const int ksplits = (MODE == kDw && g.ksplits > 1) ? g.ksplits : 1;
    const int num_kb = ceil_div(g.K, BK) / ksplits;       // kDw: K = N rows, a ragged last block is zero-filled by TMA
    const int total_tiles = g.nb * g.m_tiles * g.n_tiles * ksplits;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], NUM_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tma_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tma_b) : "memory");
    }
    if (MODE == kDw) {
        // 16 x 64 BF16 ones (any layout of all-ones is the same matrix)
        for (int i = threadIdx.x; i < ONES_BYTES / 4; i += NTHREADS)
            reinterpret_cast<uint32_t*>(smem_ones)[i] = 0x3F803F80u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)TMEM_COLS) : "memory");
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
                const int ks = tile % ksplits, t2 = tile / ksplits;
                const int nt = t2 % g.n_tiles, mt = (t2 / g.n_tiles) % g.m_tiles, b = t2 / (g.n_tiles * g.m_tiles);
                const int a_boxes = A_MN ? min(2, ceil_div(g.M - mt * BM, 64)) : 1;
                const uint32_t tx_bytes = (A_MN ? a_boxes * 8192 : A_STAGE_BYTES) + C::B_STAGE_BYTES;
                for (int kb = ks * num_kb; kb < (ks + 1) * num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], tx_bytes);
                    uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
                    uint8_t* sb = smem_b + stage * C::B_STAGE_BYTES;
                    if (!A_MN) tma_load_3d(sa, &tma_a, &full[stage], kb * BK, mt * BM, b);
                    else
                        for (int i = 0; i < a_boxes; ++i) tma_load_3d(sa + i * 8192, &tma_a, &full[stage], mt * BM + i * 64, kb * BK, b);
                    if (!B_MN) tma_load_3d(sb, &tma_b, &full[stage], kb * BK, nt * BN, b);
                    else
#pragma unroll
                        for (int i = 0; i < BN / 64; ++i) tma_load_3d(sb + i * 8192, &tma_b, &full[stage], nt * BN + i * 64, kb * BK, b);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        constexpr uint32_t idesc = make_idesc(BN, A_MN, B_MN);
        constexpr uint32_t idesc_ones = make_idesc(16, A_MN, false);
        constexpr uint32_t A_LBO = A_MN ? 8192 : 0, B_LBO = B_MN ? 8192 : 0;
        constexpr uint32_t A_KADV = A_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;   // encoded (>>4) start-address step per UMMA_K
        constexpr uint32_t B_KADV = B_MN ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
        int stage = 0; uint32_t phase = 0; int iter = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
            const int as = iter % NACC; const uint32_t aphase = (iter / NACC) & 1;
            mbar_wait(&tmem_empty[as], aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
            for (int kb = 0; kb < num_kb; ++kb) {
                mbar_wait(&full[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t adesc0 = make_desc(smem_u32(smem_a + stage * A_STAGE_BYTES), A_LBO, 1024);
                    const uint64_t bdesc0 = make_desc(smem_u32(smem_b + stage * C::B_STAGE_BYTES), B_LBO, 1024);
                    const uint64_t odesc0 = make_desc(smem_u32(smem_ones), 0, 1024);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint32_t acc = (kb > 0 || k > 0) ? 1u : 0u;     // kb counts within this tile's K slice
                        tc_mma_bf16(d_tmem, adesc0 + (uint64_t)(k * A_KADV), bdesc0 + (uint64_t)(k * B_KADV), idesc, acc);
                        if (MODE == kDw)
                            tc_mma_bf16(d_tmem + BN, adesc0 + (uint64_t)(k * A_KADV), odesc0 + (uint64_t)(k * 2), idesc_ones, acc);
                    }
                    tc_commit(&empty[stage]);                       // frees the smem stage when the MMAs retire
                    if (kb == num_kb - 1) tc_commit(&tmem_full[as]);  // accumulator ready for the epilogue
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================================================== epilogue warps
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int half = (warp - 2) >> 2;             // which half of the BN columns
        constexpr int CHUNKS = BN / 64;               // 32-column chunks per warp
        constexpr bool kHeavy = (MODE == kFwdSine || MODE == kFwdDot);   // big inlined bodies: keep the chunk loop rolled
        int iter = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++iter) {
            const int ks = tile % ksplits, t2 = tile / ksplits;
            const int nt = t2 % g.n_tiles, mt = (t2 / g.n_tiles) % g.m_tiles, b = t2 / (g.n_tiles * g.m_tiles);
            const int as = iter % NACC; const uint32_t aphase = (iter / NACC) & 1;
            const int row = mt * BM + q * 32 + lane;
            const bool row_ok = row < g.M;
            const int col_base = nt * BN + half * (BN / 2);
            float sq = 0.f;

            // everything the epilogue needs from global memory is requested *before* waiting for the
            // accumulator, so the latency hides under the MMAs of this tile
            float omega = 0.f;
            const float* bias = nullptr;
            const float* tn = nullptr;
            if (MODE == kFwdSine || MODE == kDx || MODE == kFwdDot || MODE == kFwdOut) {
                const FitRec* rec = &g.recs[b];
                omega = rec->omega;
                if (MODE != kDx) bias = rec->params + g.bias_off + col_base;
                if (MODE == kFwdOut) tn = rec->tnorm + (size_t)row * g.N + col_base;
            }
            uint32_t pre[(MODE == kDx) ? CHUNKS * 16 : 1];
            if (MODE == kDx && row_ok) {
                const __nv_bfloat16* cp = g.cprev + (size_t)b * g.cprev_fit + (size_t)row * g.N + col_base;
#pragma unroll
                for (int j = 0; j < CHUNKS * 2; ++j) ld_global_nc_256(cp + j * 16, &pre[j * 8]);
            }

            constexpr int BW = kHeavy ? ((OCC == 1) ? 16 : 8) : 1;
            float bcur[BW];
            if (kHeavy) {
#pragma unroll
                for (int j = 0; j < BW; j += 4) {
                    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + j));
                    bcur[j] = bb.x; bcur[j + 1] = bb.y; bcur[j + 2] = bb.z; bcur[j + 3] = bb.w;
                }
            }

            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + as * ACC_STRIDE + ((uint32_t)(q * 32) << 16) + half * (BN / 2);

            auto process = [&](const uint32_t (&v)[32], int c) {
                const int col = col_base + c * 32;
                if (MODE == kRaw || MODE == kDw) {
                    if (row_ok) {
                        float* dst = g.fout + (size_t)ks * g.fout_split + (size_t)b * g.fout_fit + g.fout_off + (size_t)row * g.ldf + col;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) st_global_256(dst + j, &v[j]);
                    }
                } else if (MODE == kFwdSine || MODE == kFwdDot) {
                    uint32_t so[16], co[16];
                    constexpr int GW = (OCC == 1) ? 16 : 8;      // independent sincos chains per group
#pragma unroll
                    for (int gi = 0; gi < 32 / GW; ++gi) {
                        float arg[GW], sn[GW], cs[GW];
#pragma unroll
                        for (int j = 0; j < GW; ++j) arg[j] = omega * (__uint_as_float(v[gi * GW + j]) + bcur[j]);
                        // bias of the *next* group is requested now: a whole group of math hides the L1 latency
                        const int gidx = c * (32 / GW) + gi + 1;
                        const int next = (gidx < CHUNKS * (32 / GW)) ? gidx * GW : 0;
#pragma unroll
                        for (int j = 0; j < GW; j += 4) {
                            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + next + j));
                            bcur[j] = bb.x; bcur[j + 1] = bb.y; bcur[j + 2] = bb.z; bcur[j + 3] = bb.w;
                        }
                        sincos_group(arg, sn, cs);
                        if (MODE == kFwdDot) {
                            const float* u = g.dotvec + (size_t)b * g.dotvec_fit + col + gi * GW;
#pragma unroll
                            for (int j = 0; j < GW; j += 4) {
                                const float4 uu = __ldg(reinterpret_cast<const float4*>(u + j));
                                sq = fmaf(uu.x, sn[j], sq); sq = fmaf(uu.y, sn[j + 1], sq);
                                sq = fmaf(uu.z, sn[j + 2], sq); sq = fmaf(uu.w, sn[j + 3], sq);
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < GW; j += 2) {
                                so[(gi * GW + j) / 2] = pack_bf16(sn[j], sn[j + 1]);
                                co[(gi * GW + j) / 2] = pack_bf16(cs[j], cs[j + 1]);
                            }
                        }
                    }
                    if (MODE == kFwdSine && row_ok) {
                        const size_t o = (size_t)row * g.N + col;
                        __nv_bfloat16* d0 = g.out0 + (size_t)b * g.out0_fit + o;
                        st_global_256(d0, &so[0]); st_global_256(d0 + 16, &so[8]);
                        if (g.out1) {
                            __nv_bfloat16* d1 = g.out1 + (size_t)b * g.out1_fit + o;
                            st_global_256(d1, &co[0]); st_global_256(d1 + 16, &co[8]);
                        }
                    }
                } else if (MODE == kFwdOut) {
                    if (row_ok) {
                        uint32_t tt[32], dout[16];
#pragma unroll
                        for (int j = 0; j < 32; j += 8) ld_global_nc_256(tn + c * 32 + j, &tt[j]);
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c * 32 + j));
                            const float e0 = (__uint_as_float(v[j + 0]) + bb.x) - __uint_as_float(tt[j + 0]);
                            const float e1 = (__uint_as_float(v[j + 1]) + bb.y) - __uint_as_float(tt[j + 1]);
                            const float e2 = (__uint_as_float(v[j + 2]) + bb.z) - __uint_as_float(tt[j + 2]);
                            const float e3 = (__uint_as_float(v[j + 3]) + bb.w) - __uint_as_float(tt[j + 3]);
                            sq = fmaf(e0, e0, sq); sq = fmaf(e1, e1, sq); sq = fmaf(e2, e2, sq); sq = fmaf(e3, e3, sq);
                            dout[j / 2] = pack_bf16(e0 * g.loss_scale, e1 * g.loss_scale);
                            dout[j / 2 + 1] = pack_bf16(e2 * g.loss_scale, e3 * g.loss_scale);
                        }
                        __nv_bfloat16* d0 = g.out0 + (size_t)b * g.out0_fit + (size_t)row * g.N + col;
                        st_global_256(d0, &dout[0]); st_global_256(d0 + 16, &dout[8]);
                    }
                } else if (MODE == kDx) {
                    if (row_ok) {
                        uint32_t dout[16];
#pragma unroll
                        for (int t = 0; t < 16; ++t) {
                            float c0, c1;
                            unpack_bf16(pre[c * 16 + t], c0, c1);
                            dout[t] = pack_bf16(__uint_as_float(v[2 * t]) * (omega * c0),
                                                __uint_as_float(v[2 * t + 1]) * (omega * c1));
                        }
                        __nv_bfloat16* d0 = g.out0 + (size_t)b * g.out0_fit + (size_t)row * g.N + col;
                        st_global_256(d0, &dout[0]); st_global_256(d0 + 16, &dout[8]);
                    }
                }
            };

            if constexpr (OCC == 1 && CHUNKS > 1) {
                // software-pipelined TMEM reads: the load of chunk c+1 is in flight while chunk c is processed
                uint32_t va[32], vb[32];
                tmem_ld32(t_row, va);
                if constexpr (kHeavy) {
#pragma unroll 1
                    for (int c = 0; c < CHUNKS; c += 2) {
                        tmem_ld_wait();
                        tmem_ld32(t_row + (c + 1) * 32, vb);
                        process(va, c);
                        tmem_ld_wait();
                        if (c + 2 < CHUNKS) tmem_ld32(t_row + (c + 2) * 32, va);
                        process(vb, c + 1);
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < CHUNKS; c += 2) {
                        tmem_ld_wait();
                        tmem_ld32(t_row + (c + 1) * 32, vb);
                        process(va, c);
                        tmem_ld_wait();
                        if (c + 2 < CHUNKS) tmem_ld32(t_row + (c + 2) * 32, va);
                        process(vb, c + 1);
                    }
                }
            } else {
                // two CTAs per SM: the co-resident CTA hides the TMEM latency, keep registers low
                uint32_t va[32];
                if constexpr (kHeavy) {
#pragma unroll 1
                    for (int c = 0; c < CHUNKS; ++c) { tmem_ld32(t_row + c * 32, va); tmem_ld_wait(); process(va, c); }
                } else {
#pragma unroll
                    for (int c = 0; c < CHUNKS; ++c) { tmem_ld32(t_row + c * 32, va); tmem_ld_wait(); process(va, c); }
                }
            }
            if (MODE == kDw && half == 0 && nt == 0) {
                const uint32_t dbv = tmem_ld1(tmem_base + as * ACC_STRIDE + ((uint32_t)(q * 32) << 16) + BN);   // every column of the ones-product equals db[row]
                tmem_ld_wait();
                if (row_ok) g.biasgrad[(size_t)b * g.biasgrad_fit + (size_t)ks * g.M + row] = __uint_as_float(dbv);
            }
            if (MODE == kFwdDot && row_ok)     // this thread's row, this warp's half of the tile's columns
                g.dotpart[(size_t)b * g.dotpart_fit + (size_t)(nt * 2 + half) * g.M + row] = sq;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
            if (MODE == kFwdOut) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                if (lane == 0)
                    g.losspart[(size_t)b * g.losspart_per_fit + (size_t)(mt * g.n_tiles + nt) * NUM_EPI_WARPS + (warp - 2)] = sq;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ SIMT helpers of the tensor path
// bf16 mirror of the packed parameters (the MMAs read weights from it)
__global__ void mirror_kernel(const FitRec* recs, int P, __nv_bfloat16* wbf16) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) wbf16[(size_t)blockIdx.y * P + p] = __float2bfloat16_rn(recs[blockIdx.y].params[p]);
}
inline void mirror_weights(const FitRec* recs, const LayerMap& lm, int nf, __nv_bfloat16* wbf16, cudaStream_t s) {
    dim3 grid(ceil_div(lm.P, 256), nf);
    mirror_kernel<<<grid, 256, 0, s>>>(recs, lm.P, wbf16);
}

// gradient of layer 0 (fp32, positions never quantised): per 128-row tile
//   xpart[f][t][j] = sum_r dz0[r][j] * x[r]      colpart0[f][t][j] = sum_r dz0[r][j]
// HBM-bound (reads dz0 once): 16-byte loads, 8 columns per thread, the 128 rows of the tile split over
// the thread groups of the block and combined through shared memory in a fixed order (deterministic).
__global__ void __launch_bounds__(256)
layer0_grad_kernel(const FitRec* recs, const __nv_bfloat16* dz0, size_t dz_fit, int N, int H, int mtiles,
                   float* xpart, float* colpart0) {
    __shared__ float red[2][8][8 * 32 + 8];               // [sum kind][row slice][columns of this pass]
    const int f = blockIdx.y, t = blockIdx.x;
    const float* pos = recs[f].pos;
    const int r0 = t * 128, r1 = min(N, r0 + 128);
    const int cu = threadIdx.x & 31, rs = threadIdx.x >> 5;          // column unit (8 columns), row slice
    for (int c0 = 0; c0 < H; c0 += 256) {                             // 256 columns per pass
        const int col = c0 + cu * 8;
        float sx[8], s1[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) { sx[k] = 0.f; s1[k] = 0.f; }
        if (col < H) {
            const __nv_bfloat16* src = dz0 + (size_t)f * dz_fit + col;
#pragma unroll 4
            for (int r = r0 + rs; r < r1; r += 8) {
                const uint4 u = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * H));
                const float x = __ldg(pos + r);
                const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float a, b;
                    unpack_bf16(w[k], a, b);
                    sx[2 * k] = fmaf(a, x, sx[2 * k]); sx[2 * k + 1] = fmaf(b, x, sx[2 * k + 1]);
                    s1[2 * k] += a; s1[2 * k + 1] += b;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) { red[0][rs][cu * 8 + k] = sx[k]; red[1][rs][cu * 8 + k] = s1[k]; }
        __syncthreads();
        {
            const int j = threadIdx.x;                                // one column per thread
            if (c0 + j < H) {
                float ax = 0.f, a1 = 0.f;
#pragma unroll
                for (int s = 0; s < 8; ++s) { ax += red[0][s][j]; a1 += red[1][s][j]; }
                xpart[((size_t)f * mtiles + t) * H + c0 + j] = ax;
                colpart0[((size_t)f * mtiles + t) * H + c0 + j] = a1;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ host: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 3-D map over a bf16 array [batch][outer][inner] (inner contiguous), 128B swizzle.
inline int make_map(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t batch,
                    uint64_t outer_stride_elems, uint64_t batch_stride_elems, uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled not available from the driver"); return NA_ERR_CUDA; }
    cuuint64_t dims[3] = {inner, outer, batch};
    cuuint64_t strides[2] = {outer_stride_elems * 2, batch_stride_elems * 2};
    cuuint32_t box[3] = {box_inner, box_outer, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu batch=%llu box=%ux%u", (int)r,
                  (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)batch, box_inner, box_outer);
        return NA_ERR_CUDA;
    }
    return NA_OK;
}

// Operand stored [batch][rows][cols] row-major.
//   K-major use  (cols = K):  box {64, box_rows}
//   MN-major use (rows = K):  box {64, 64}
inline int make_operand_map(CUtensorMap* map, const void* base, int rows, int cols, int batch, size_t batch_stride,
                            bool mn_major, int box_rows) {
    return make_map(map, base, cols, rows, batch, cols, batch_stride, 64, mn_major ? 64 : box_rows);
}

struct GemmMaps { CUtensorMap a, b; };
struct GroupMaps {
    GemmMaps fwd[kMaxHidden + 1];    // [l] for l = 1..L
    GemmMaps out;
    GemmMaps dx[kMaxHidden + 2];     // [l] for l = 1..L+1
    GemmMaps dw[kMaxHidden + 2];     // [l] for l = 1..L+1
};

inline int hidden_bn(int H) { return H >= 256 ? 256 : H; }
inline int dw_bn(int H) { return H >= 128 ? 128 : 64; }

// `dz_per_layer` (chain mode): dz_l has its own array per layer instead of the two ping-pong buffers
inline int build_group_maps(int N, int D, int H, int L, int nf, const LayerMap& lm, void* const* act, void* const* cosb,
                            void* const* dz, void* dy, __nv_bfloat16* wbf16, GroupMaps& m,
                            void* const* dz_per_layer = nullptr) {
    (void)cosb;
    const size_t nh = (size_t)N * H, nd = (size_t)N * D;
    int rc;
    for (int l = 1; l <= L; ++l) {
        if ((rc = make_operand_map(&m.fwd[l].a, act[l - 1], N, H, nf, nh, false, BM))) return rc;
        if ((rc = make_operand_map(&m.fwd[l].b, wbf16 + lm.w_off[l], H, H, nf, lm.P, false, hidden_bn(H)))) return rc;
    }
    if ((rc = make_operand_map(&m.out.a, act[L], N, H, nf, nh, false, BM))) return rc;
    if ((rc = make_operand_map(&m.out.b, wbf16 + lm.w_off[L + 1], D, H, nf, lm.P, false, out_bn(D)))) return rc;
    for (int l = L + 1; l >= 1; --l) {
        const int width = lm.out_dim[l];
        const void* dzl = (l == L + 1) ? dy : (dz_per_layer ? dz_per_layer[l] : dz[l & 1]);
        const size_t dz_fit = (l == L + 1) ? nd : nh;
        // dx: A = dz_l [N x width] K-major, B = W_l [width x H] MN-major
        if ((rc = make_operand_map(&m.dx[l].a, dzl, N, width, nf, dz_fit, false, BM))) return rc;
        if ((rc = make_operand_map(&m.dx[l].b, wbf16 + lm.w_off[l], width, H, nf, lm.P, true, 0))) return rc;
        // dw: A = dz_l [N x width] MN-major (M = width), B = act_{l-1} [N x H] MN-major
        if ((rc = make_operand_map(&m.dw[l].a, dzl, N, width, nf, dz_fit, true, 0))) return rc;
        if ((rc = make_operand_map(&m.dw[l].b, act[l - 1], N, H, nf, nh, true, 0))) return rc;
    }
    return NA_OK;
}

// ------------------------------------------------------------------ host: launches
// Kernel attributes (the dynamic shared-memory opt-in) and the SM count belong to a device, not to the process:
// both are cached per device ordinal, so a second GPU used from the same process is configured on first use.
constexpr int kMaxDevices = 64;
inline int current_device() { int dev = 0; cudaGetDevice(&dev); return (dev >= 0 && dev < kMaxDevices) ? dev : 0; }
inline int num_sms() {
    static int n[kMaxDevices] = {};
    const int dev = current_device();
    if (!n[dev]) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        n[dev] = v > 0 ? v : 148;
    }
    return n[dev];
}

// bit m of the mask selects the 2-CTA/SM variant for Mode m (NERFATTN_OCC2 overrides the default)
inline unsigned occ2_mask() {
    static int mask = -1;
    if (mask < 0) {
        const char* e = getenv("NERFATTN_OCC2");
        mask = e ? (int)strtol(e, nullptr, 0) : kDefaultOcc2Mask;
    }
    return (unsigned)mask;
}

template <int MODE, bool A_MN, bool B_MN, int BN, int OCC>
inline int launch_occ(const GemmMaps& maps, TcArgs a, cudaStream_t s) {
    constexpr int smem = smem_bytes<BN, OCC>();
    a.m_tiles = ceil_div(a.M, BM);
    a.n_tiles = a.N / BN;
    const int tiles = a.nb * a.m_tiles * a.n_tiles * ((MODE == kDw && a.ksplits > 1) ? a.ksplits : 1);
    const int grid = std::min(tiles, num_sms() * OCC);
    tc_gemm_kernel<MODE, A_MN, B_MN, BN, OCC><<<grid, NTHREADS, smem, s>>>(maps.a, maps.b, a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("tc_gemm launch failed: %s", cudaGetErrorString(e)); return NA_ERR_CUDA; }
    return NA_OK;
}
template <int MODE, bool A_MN, bool B_MN, int BN>
inline int launch_one(const GemmMaps& maps, const TcArgs& a, cudaStream_t s) {
    if ((occ2_mask() >> MODE) & 1u) return launch_occ<MODE, A_MN, B_MN, BN, 2>(maps, a, s);
    return launch_occ<MODE, A_MN, B_MN, BN, 1>(maps, a, s);
}

template <int MODE, bool A_MN, bool B_MN>
inline int launch_bn(int bn, const GemmMaps& maps, const TcArgs& a, cudaStream_t s) {
    switch (bn) {
        case 64: return launch_one<MODE, A_MN, B_MN, 64>(maps, a, s);
        case 128: return launch_one<MODE, A_MN, B_MN, 128>(maps, a, s);
        case 256:
            if constexpr (MODE != kDw) return launch_one<MODE, A_MN, B_MN, 256>(maps, a, s);
        default: break;
    }
    set_error("tc_gemm: unsupported BN %d", bn);
    return NA_ERR_UNSUPPORTED;
}

// One training epoch of one group on the tensor path (everything except Adam).
inline int epoch(int N, int D, int H, int L, int nf, const LayerMap& lm, const FitRec* recs, const GroupMaps& m,
                 void* const* act, void* const* cosb, void* const* dz, void* dy, float* gradpart, float* colpart,
                 const size_t* colpart_layer_off, float* xpart, float* losspart, int losspart_per_fit, int mtiles,
                 cudaStream_t s) {
    const size_t nh = (size_t)N * H, nd = (size_t)N * D;
    int rc;
    {
        const size_t total8 = nh / 8;
        dim3 grid((unsigned)ceil_div(total8, (size_t)256), nf);
        f32::layer0_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(recs, N, H, (__nv_bfloat16*)act[0], (__nv_bfloat16*)cosb[0], nh);
    }
    TcArgs base{};
    base.nb = nf; base.recs = recs; base.loss_scale = 2.0f / ((float)N * (float)D);
    for (int l = 1; l <= L; ++l) {
        TcArgs a = base;
        a.M = N; a.N = H; a.K = H; a.bias_off = lm.b_off[l];
        a.out0 = (__nv_bfloat16*)act[l]; a.out0_fit = nh;
        a.out1 = (__nv_bfloat16*)cosb[l]; a.out1_fit = nh;
        if ((rc = launch_bn<kFwdSine, false, false>(hidden_bn(H), m.fwd[l], a, s))) return rc;
    }
    {
        TcArgs a = base;
        a.M = N; a.N = D; a.K = H; a.bias_off = lm.b_off[L + 1];
        a.out0 = (__nv_bfloat16*)dy; a.out0_fit = nd;
        a.losspart = losspart; a.losspart_per_fit = losspart_per_fit;
        if ((rc = launch_bn<kFwdOut, false, false>(out_bn(D), m.out, a, s))) return rc;
    }
    for (int l = L + 1; l >= 1; --l) {
        const int width = lm.out_dim[l];
        {   // dW_l, db_l
            TcArgs a = base;
            a.M = width; a.N = H; a.K = N;
            a.fout = gradpart; a.fout_fit = lm.P; a.fout_off = lm.w_off[l]; a.ldf = H;
            a.biasgrad = colpart + colpart_layer_off[l]; a.biasgrad_fit = width;
            if ((rc = launch_bn<kDw, true, true>(dw_bn(H), m.dw[l], a, s))) return rc;
        }
        {   // dz_{l-1}
            TcArgs a = base;
            a.M = N; a.N = H; a.K = width;
            a.out0 = (__nv_bfloat16*)dz[(l - 1) & 1]; a.out0_fit = nh;
            a.cprev = (const __nv_bfloat16*)cosb[l - 1]; a.cprev_fit = nh;
            if ((rc = launch_bn<kDx, false, true>(hidden_bn(H), m.dx[l], a, s))) return rc;
        }
    }
    layer0_grad_kernel<<<dim3(mtiles, nf), 256, 0, s>>>(recs, (const __nv_bfloat16*)dz[0], nh, N, H, mtiles, xpart,
                                                        colpart + colpart_layer_off[0]);
    return NA_OK;
}

// cudaFuncSetAttribute for every instantiation, once, outside any stream capture.
template <int MODE, bool A_MN, bool B_MN, int BN>
inline cudaError_t configure_one() {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<MODE, A_MN, B_MN, BN, 1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<BN, 1>());
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(tc_gemm_kernel<MODE, A_MN, B_MN, BN, 2>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<BN, 2>());
}
inline int configure_all() {
    static std::once_flag once_dev[kMaxDevices];
    static cudaError_t err_dev[kMaxDevices] = {};
    const int dev = current_device();
    cudaError_t& err = err_dev[dev];
    std::call_once(once_dev[dev], [&err] {
        auto acc = [&](cudaError_t e) { if (e != cudaSuccess && err == cudaSuccess) err = e; };
#define NA_CFG3(MODE, A, B) acc(configure_one<MODE, A, B, 64>()); acc(configure_one<MODE, A, B, 128>()); acc(configure_one<MODE, A, B, 256>());
        NA_CFG3(kRaw, false, false) NA_CFG3(kRaw, false, true) NA_CFG3(kRaw, true, false) NA_CFG3(kRaw, true, true)
        NA_CFG3(kFwdSine, false, false) NA_CFG3(kFwdOut, false, false) NA_CFG3(kDx, false, true)
        NA_CFG3(kFwdDot, false, false)
#undef NA_CFG3
        acc(configure_one<kDw, true, true, 64>()); acc(configure_one<kDw, true, true, 128>());
    });
    if (err != cudaSuccess) { set_error("cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(err)); return NA_ERR_CUDA; }
    return NA_OK;
}

}  // namespace tc
}  // namespace na
