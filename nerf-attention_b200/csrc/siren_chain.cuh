// Fused row-tile chain of the BF16 tensor path (sm_100a).
//
// The unfused path (siren_tc.cuh) runs one grouped GEMM kernel per layer and direction and
// moves every activation through HBM between them: ~25 N*H*2 B per fit-epoch for `medium`,
// which made the backward kernels HBM-bound while the sine epilogues were FP32-issue-bound,
// one after the other.  Everything except dW is row-local, so here ONE persistent CTA takes a
// 128-row tile of one fit through the whole chain
//
//   E0      h_0 = sin(w (x W0^T + b0))                          SIMT, positions never quantised
//   S_l     z_l = h_{l-1} W_l^T  (tcgen05, A = h_{l-1} in smem)  -> h_l = sin(w (z_l + b_l))     l = 1..L
//   OUT     y   = h_L Wf^T + bf                                  -> dY = 2 (y - t)/(N D), loss
//   DXF     dh  = dY Wf          (A = dY in smem)                -> dz_L = dh * w cos_L
//   DX_l    dh  = dz_l W_l       (A = dz_l in smem)              -> dz_{l-1} = dh * w cos_{l-1}  l = L..1
//
// (reference: siren.py:33-34,60-61,100-102).  The epilogue warps write each result straight
// into shared memory in the 128B-swizzled K-major layout the next MMA reads as its A operand;
// only the weights stream in (TMA, 3-stage ring).  The same buffer is what the dW GEMMs need
// as h_l / dz_l, so one elected thread TMA-stores it to global (bf16) -- no LSU traffic for it.
// dW contracts over all rows of a fit and stays a separate kernel.  cos_l goes to a small
// per-CTA scratch that the same thread reads back in the backward half.
//
// Roles (640 threads): warp 0 TMA producer, warp 1 MMA issuer, warps 4..19 epilogue.  For
// H <= 256 two tiles ("slots") are in flight per CTA and all 16 epilogue warps take them in turn, step by
// step: while they run one slot's epilogue the tensor core runs the other slot's MMAs and the store warp
// writes its operand buffer out, and with 4 epilogue warps per scheduler the TMEM / shared / L2 latencies
// of one warp hide under the others.  (Earlier versions: 8 warps alternating over both slots issued 28 % of
// cycles; one set of 8 warps per slot left each set waiting a third of the time for its own MMA.)
// TMEM: 2 x 256 accumulator columns.  H = 512 needs the whole TMEM and 128 KB of shared memory
// for one tile: one slot, all 16 warps on it (a quarter of the columns each).
#pragma once

#include "siren_tc.cuh"

// profiling-only debug addressing (libnerfattn_prof.so, -DNA_PROFILING); constant 0 in the release library
#ifdef NA_PROFILING
#define NA_DBG(g) ((g).dbg)
#else
#define NA_DBG(g) 0
#endif

namespace na {
namespace chain {

using namespace tc;

constexpr int NCTRL = 4;                          // warpgroup 0: warp 0 = TMA, warp 1 = MMA, warps 2-3 idle
constexpr int NEPI = 16;                          // epilogue warps 4..19
constexpr int NTHREADS = (NCTRL + NEPI) * 32;     // 640
constexpr int CHUNK_BYTES = BM * 64 * 2;          // one K-chunk of an A operand: 128 rows x 64 bf16
constexpr int STAGE_BYTES = 256 * 64 * 2;         // one K-chunk of a weight operand: <= 256 x 64 bf16
constexpr int BAR_BYTES = 512;                    // mbarriers (<= 32) + the TMEM base slot
// What a tile needs from its FitRec, resolved once per round by one lane of every epilogue warp and kept in shared memory
// (one record per warp and slot): held in registers across the step loop these values were spilled to local memory, and
// with 228 KB of the SM's array given to shared memory the reloads came from L2 (profiles/README.md, round 2).
struct alignas(16) SlotRec { int fit, mt; float omega; int pad0; const float* pos; const float* tnorm; const float* obias; const float* pad1; };
constexpr int TAIL_BYTES = BAR_BYTES + 16 * 2 * (int)sizeof(SlotRec);
constexpr int SMEM_BUDGET = 227 * 1024 - TAIL_BYTES;   // dynamic shared memory per CTA for the operand buffers and the weight ring

// NS = tiles in flight per CTA (2 needs 2 x max(H, 256) accumulator columns and operand buffers)
template <int H, int NS> struct Cfg {
    static_assert(NS == 1 || (NS == 2 && H <= 256), "two slots need H <= 256");
    static constexpr int BN = (H >= 256) ? 256 : H;              // N of one hidden-layer MMA
    static constexpr int NPARTS = H / BN;
    static constexpr int NSLOT = NS;
    static constexpr int EPW = NEPI;                              // all epilogue warps work on one slot at a time (alternating)
    static constexpr int CG = EPW / 4;                            // column groups per slot
    static constexpr int CW = H / CG;                             // columns per thread in a hidden step
    static constexpr int NU = CW / 16;                            // 16-column units per thread
    static constexpr int ACT_CHUNKS = ((H > 256) ? H : 256) / 64;   // dY (D <= 256) aliases the buffer
    static constexpr int ACT_BYTES = ACT_CHUNKS * CHUNK_BYTES;
    static constexpr int ACC_COLS = (H > 256) ? 512 : 256;
    static constexpr int STAGES = ((SMEM_BUDGET - NSLOT * ACT_BYTES) / STAGE_BYTES < 6) ? (SMEM_BUDGET - NSLOT * ACT_BYTES) / STAGE_BYTES : 6;
    static constexpr int SMEM = NSLOT * ACT_BYTES + STAGES * STAGE_BYTES + TAIL_BYTES;   // + barriers and the per-warp slot records
    static_assert(STAGES >= 3, "weight ring");
    static_assert(NU >= 1 && CW % 16 == 0, "column split");
};

struct ChainArgs {
    int N, D, L, nf, mtiles;
    const FitRec* recs;
    int w_off[kMaxLayers], b_off[kMaxLayers];
    __nv_bfloat16* scratch;                 // cos_l of the tiles in flight: [grid][NSLOT][L+1][H/16][128][16] (slot 0 of L+1 unused: cos_0 is recomputed)
    float* losspart; int losspart_per_fit; float loss_scale;
    const float* dotvec; float* dotpart;    // forward-only (decode): u [nf][H] fp32, partial scores [nf][CG][N]
    const float* pvec; float* pvpart;       // forward-only (decode, values): p [nf][N] fp32, partial sums [nf][N/32][H]
    const float* psc; int psc_fit;          // per fit w*W0[H], w*b0[H], w*b_1[H] .. w*b_L[H] (scale_params_kernel; Adam keeps it current)
    float* xpart; float* colpart0;          // training: layer-0 gradient partials per row tile [nf][mtiles][H]: sum_r dz0[r][j] x[r], sum_r dz0[r][j]
    int pair_tiles;                         // tile_of(): adjacent tiles in the two slots of a CTA (set by launch_h)
    int dbg;                                // NERFATTN_CHAIN_DBG (libnerfattn_prof.so only; 0 in the release library)
    int sincos_mode;                        // bit 0: hidden layers, bit 1: layer 0 use the MUFU-core sincos (common.cuh)
};
// wk / wmn: weights of layers 1..L+1 as K-major (forward) and MN-major (backward) B operands (loads);
// hout[l] / zout[l] / yout: h_l, dz_l [nf][N][H] and dY [nf][N][D] as 64 x 128 boxes (stores)
// xop: the B operand of the layer-0 gradient MMA, [position table][row tile][16][128] bf16 (xop_kernel)
struct ChainMaps {
    CUtensorMap wk[kMaxHidden + 2]; CUtensorMap wmn[kMaxHidden + 2];
    CUtensorMap hout[kMaxHidden + 1]; CUtensorMap zout[kMaxHidden + 1]; CUtensorMap yout;
    CUtensorMap xop;
};

// Any sequence length: the last row tile may be ragged (rows >= N are masked out of the loss, produce dY = 0 and
// hence no gradient, and are clipped by the TMA stores / zero-filled by the dW kernel's TMA loads).
inline bool shape_supported(int N, int D, int H, int L) {
    return N >= 1 && (H == 64 || H == 128 || H == 256 || H == 512) && (D == 64 || D == 128 || D == 256) && L >= 1;
}
inline int slots_for(int H);
inline int loss_partials_per_fit(int N, int /*H*/) { return ceil_div(N, BM) * NEPI; }   // one per epilogue warp of the tile

__device__ __forceinline__ void st_shared_128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// L2 residency control: the cos scratch is re-read by the same thread a few microseconds later
// and then overwritten in place by the next tile (keep it: evict-last, and keep it out of the
// small L1 that holds the layer-0 weights); targets are streamed past L1.
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_global_256_hint(void* p, const uint32_t* v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8}, %9;"
                 ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "l"(pol) : "memory");
}
__device__ __forceinline__ void ld_global_256_hint(const void* p, uint32_t* v, uint64_t pol) {   // coherent (same-kernel data)
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8], %9;"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p), "l"(pol) : "memory");
}
__device__ __forceinline__ void ld_global_nc_na_256(const void* p, uint32_t* v) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
                 ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "l"(pol) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// cta_group::2: one MMA spans the CTA pair (M = 256: 128 rows of A and D in each CTA, B split in halves between
// the two shared memories), issued by the leader CTA only
__device__ __forceinline__ void tc_mma_bf16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// TMA load of the peer CTA's half of a weight stage: lands in the issuing CTA's shared memory, completes the
// transaction count of the LEADER's barrier (the MMA that consumes both halves is issued there)
__device__ __forceinline__ void tma_load_3d_2cta(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(const void* local, uint32_t cta_rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local)), "r"(cta_rank));
    return ra;
}
__device__ __forceinline__ void tc_commit_2cta(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* local_bar, uint32_t cta_rank) {     // arrive on `cta_rank`'s copy
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_bar)), "r"(cta_rank));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__host__ __device__ constexpr uint32_t make_idesc_m(int m, int n, bool b_mn, bool a_mn = false) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}
constexpr int XOP_N = 16;                         // rows of the layer-0 gradient operand: ones, x_hi, x_mid, x_lo, 12 x zero
constexpr int XOP_BYTES = XOP_N * BM * 2;         // one row tile of it: [16][128] bf16 = two 64-K chunks of 2 KB
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"((uint64_t)map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {      // bytes: multiple of 16
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void set_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// 16 consecutive bf16 columns [col, col+16) of tile row r -> A-operand buffer (K-major, SWIZZLE_128B:
// 64-column chunks of 128 rows x 128 B, 16-byte units XOR-ed with row % 8 -- the TMA layout)
__device__ __forceinline__ void act_store16(uint32_t act_u32, int r, int col, const uint32_t (&pk)[8]) {
    const uint32_t rowaddr = act_u32 + (uint32_t)(col >> 6) * CHUNK_BYTES + (uint32_t)r * 128;
    const int u0 = (col & 63) >> 3;
    st_shared_128(rowaddr + (uint32_t)((u0 ^ (r & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
    st_shared_128(rowaddr + (uint32_t)(((u0 + 1) ^ (r & 7)) << 4), pk[4], pk[5], pk[6], pk[7]);
}

// what MMA step s (1-based; step 0 is the SIMT layer 0) contracts
struct Step { int layer; int mn; int kch; int n; int nparts; };
template <int H>
__device__ __forceinline__ Step step_info(int s, int L, int D) {
    using C = Cfg<H, 1>;
    Step st;
    if (s <= L) { st.layer = s; st.mn = 0; st.kch = H / 64; st.n = C::BN; st.nparts = C::NPARTS; }
    else if (s == L + 1) { st.layer = L + 1; st.mn = 0; st.kch = H / 64; st.n = D; st.nparts = 1; }
    else if (s == L + 2) { st.layer = L + 1; st.mn = 1; st.kch = D / 64; st.n = C::BN; st.nparts = C::NPARTS; }
    else { st.layer = 2 * L + 3 - s; st.mn = 1; st.kch = H / 64; st.n = C::BN; st.nparts = C::NPARTS; }
    return st;
}

#ifndef NA_RECOMPUTE_COS0
#define NA_RECOMPUTE_COS0 1
#endif
#ifdef NA_EXP_TIMING                              // profiles/: per-step cycle counters of two epilogue warps, printed per launch
#define NA_CHAIN_TIMING 1
#endif
template <bool MUFU>
__device__ __forceinline__ void sincos8(const float (&x)[8], float (&s)[8], float (&c)[8]) {
    if (MUFU) sincos_group_mufu(x, s, c); else sincos_group(x, s, c);
}

// FWD = false: the training chain above.  FWD = true: forward only, for the fused decode
// (evaluate.py:173-242 times this reconstruction): E0, S_1..S_L, and instead of materialising K the
// last sine epilogue reduces u . sin(.) per position, u = Wf^T (q * std) (decode.cuh) -- no cos, no
// global stores except one partial score per position and column group.
// CL = CTAs per cluster (1 or 2).  With CL = 2 the pair (2c, 2c+1) takes adjacent row tiles of one fit, each
// CTA issues half of every weight stage's 64 x 64 boxes as a multicast to both, so the L2 -> SM weight
// traffic and the time to fill a stage halve; nothing else is shared (each CTA issues its own MMAs).
// MODE 0: training chain; 1: forward only, u . sin(.) per position (decode logits); 2: forward only,
// p_t * sin(.) summed over the positions (decode, values)
template <int H, int NS, int MODE, int CL>
__global__ void __launch_bounds__(NTHREADS, 1)
chain_kernel(const __grid_constant__ ChainMaps maps, const ChainArgs g) {
    constexpr bool FWD = MODE != 0;
    using C = Cfg<H, NS>;
    constexpr int NSLOT = C::NSLOT;
    // CTA pair: each CTA holds half of every weight stage -> twice as many stages of half the size in the same ring
    constexpr int STAGES = (CL == 2) ? 2 * C::STAGES : C::STAGES;
    constexpr int STAGE_SZ = (CL == 2) ? STAGE_BYTES / 2 : STAGE_BYTES;
    const uint32_t crank = (CL == 2) ? cluster_ctarank() : 0u;
    // tile of (round, slot): clusters stride over tile pairs, CTAs of a cluster take consecutive tiles
    const int cl_stride = (int)gridDim.x / CL;
    // g.pair_tiles (two slots, enough tiles to fill both everywhere): the two slots of a CTA take ADJACENT tiles, i.e. the
    // same fit -- its layer-0 parameters, biases and weight prefetches are then shared by both slots through L1 / L2
    // (-1.6 % kernel time); otherwise tiles are dealt round-robin so that a small launch spreads over all SMs first
    auto tile_of = [&](int round, int slot) {
        return (CL == 1 && NSLOT == 2 && g.pair_tiles) ? 2 * ((int)blockIdx.x + round * (int)gridDim.x) + slot
                                                       : CL * ((int)blockIdx.x / CL + (round * NSLOT + slot) * cl_stride) + (int)crank;
    };
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* smem_ring = smem + NSLOT * C::ACT_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_ring + STAGES * STAGE_SZ);
    uint64_t* full = bars;                       // [STAGES]  weights landed
    uint64_t* empty = bars + STAGES;             // [STAGES]  MMAs that read the stage retired
    uint64_t* acc_full = bars + 2 * STAGES;      // [2]       accumulator of the slot complete
    uint64_t* act_ready = bars + 2 * STAGES + 2; // [2]       A operand of the slot written, accumulator drained
    uint64_t* buf_free = bars + 2 * STAGES + 4;  // [2]       the TMA store of the slot's operand buffer has read it
    uint64_t* mma_ready = bars + 2 * STAGES + 6; // [2]       CL = 2, leader: the operand buffers of both CTAs are written
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = g.L, D = g.D;
    const int total_tiles = g.nf * g.mtiles;
    const int nsteps = FWD ? L + 1 : 2 * L + 3;  // step 0 (layer 0) + the MMA steps

    if (threadIdx.x == 0) {
        if (smem_u32(smem) & 1023u) { printf("nerfattn: chain smem base not 1024-aligned\n"); __trap(); }
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&acc_full[i], 1); mbar_init(&act_ready[i], C::EPW); mbar_init(&buf_free[i], 1);
            mbar_init(&mma_ready[i], 2 * C::EPW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CL == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();              // the peer's barriers exist before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // register reallocation (sm_90+): the two control warps need few registers, the epilogue warps want more
    // than the 96 a 640-thread CTA gets by default
    if (warp < NCTRL) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
        // ===================================================== TMA producer: weights only
        // The ring holds 96 KB, one hidden step reads 128 KB, and the 16 CTAs working on a fit ask for
        // the same chunk at the same moment: without help every step waits one DRAM latency for
        // its last chunk.  So the weights of step s+1 are pulled into L2 while step s is loaded.
        if (elect_one()) {
            auto prefetch_step = [&](int s, int fit) {
                const Step st = step_info<H>(s, L, D);
                const CUtensorMap* map = st.mn ? &maps.wmn[st.layer] : &maps.wk[st.layer];
                for (int np = 0; np < st.nparts; ++np)
                    for (int kc = 0; kc < st.kch; ++kc) {
                        if (CL == 2) {
                            const int nb = st.n / 128, b0 = (int)crank * nb;
                            for (int j = 0; j < nb; ++j) {
                                if (!st.mn) tma_prefetch_3d(map, kc * 64, np * 256 + (b0 + j) * 64, fit);
                                else tma_prefetch_3d(map, np * 256 + (b0 + j) * 64, kc * 64, fit);
                            }
                        } else if (!st.mn) tma_prefetch_3d(map, kc * 64, np * 256, fit);
                        else
                            for (int i = 0; i < st.n / 64; ++i) tma_prefetch_3d(map, np * 256 + i * 64, kc * 64, fit);
                    }
            };
            int stage = 0; uint32_t phase = 0;
            // the 128 x D fp32 target rows of a tile are contiguous: pull them into L2 when the tile starts
            // (the OUT epilogue, half a tile later, otherwise walks them at DRAM latency)
            auto prefetch_targets = [&](int tile) {
                if (FWD) return;
                const int fit = tile / g.mtiles, mt = tile - fit * g.mtiles;
                l2_prefetch_bulk(g.recs[fit].tnorm + (size_t)mt * BM * D, (uint32_t)(BM * D * 4));
            };
            for (int slot = 0; slot < NSLOT; ++slot) {
                const int tile = tile_of(0, slot);
                if (tile < total_tiles) { prefetch_step(1, tile / g.mtiles); prefetch_targets(tile); }
            }
            for (int round = 0;; ++round) {
                if (tile_of(round, 0) >= total_tiles) break;
                for (int s = 1; s <= (FWD ? nsteps - 1 : nsteps); ++s) {
                    if (s == nsteps) {
                        // training: the B operand of the layer-0 gradient MMA (positions of this row tile, siren_chain.cuh
                        // "layer-0 gradient"): one 4 KB stage per tile, two 64-K chunks of [16][64] bf16
                        for (int slot = 0; slot < NSLOT; ++slot) {
                            const int tile = tile_of(round, slot);
                            if (tile >= total_tiles) break;
                            const int fit = tile / g.mtiles, mt = tile - fit * g.mtiles;
                            const int xrow = g.recs[fit].posid * g.mtiles + mt;
                            mbar_wait(&empty[stage], phase ^ 1);
                            if (CL == 1 || crank == 0) mbar_expect_tx(&full[stage], (uint32_t)(CL * XOP_BYTES));
                            uint8_t* sb = smem_ring + stage * STAGE_SZ;
                            const uint32_t lbar = (CL == 2) ? mapa_u32(&full[stage], 0) : 0u;
                            for (int c = 0; c < 2; ++c) {
                                if (CL == 1 || crank == 0) tma_load_3d(sb + c * (XOP_BYTES / 2), &maps.xop, &full[stage], c * 64, 0, xrow);
                                else tma_load_3d_2cta(sb + c * (XOP_BYTES / 2), &maps.xop, lbar, c * 64, 0, xrow);
                            }
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                        continue;
                    }
                    const Step st = step_info<H>(s, L, D);
                    const CUtensorMap* map = st.mn ? &maps.wmn[st.layer] : &maps.wk[st.layer];
                    const uint32_t tx = (uint32_t)st.n * 128u;           // both halves land on the leader's barrier when CL == 2
                    for (int slot = 0; slot < NSLOT; ++slot) {
                        const int tile = tile_of(round, slot);
                        if (tile >= total_tiles) break;
                        const int fit = tile / g.mtiles;
                        if (s + 1 < nsteps) prefetch_step(s + 1, fit);
                        else {
                            const int ntile = tile_of(round + 1, slot);
                            if (ntile < total_tiles) {
                                if (ntile / g.mtiles != fit) prefetch_step(1, ntile / g.mtiles);
                                prefetch_targets(ntile);
                            }
                        }
                        for (int np = 0; np < st.nparts; ++np)
                            for (int kc = 0; kc < st.kch; ++kc) {
                                mbar_wait(&empty[stage], phase ^ 1);
                                if (CL == 1 || crank == 0) mbar_expect_tx(&full[stage], tx);
                                uint8_t* sb = smem_ring + stage * STAGE_SZ;
                                if (CL == 2) {
                                    // this CTA's half of B (n/2 rows K-major, or n/2 columns MN-major) as 64 x 64 boxes
                                    const int nb = st.n / 128, b0 = (int)crank * nb;
                                    const uint32_t lbar = mapa_u32(&full[stage], 0);
                                    for (int j = 0; j < nb; ++j) {
                                        const int c0 = st.mn ? np * 256 + (b0 + j) * 64 : kc * 64;
                                        const int c1 = st.mn ? kc * 64 : np * 256 + (b0 + j) * 64;
                                        if (crank == 0) tma_load_3d(sb + j * 8192, map, &full[stage], c0, c1, fit);
                                        else tma_load_3d_2cta(sb + j * 8192, map, lbar, c0, c1, fit);
                                    }
                                } else if (!st.mn) tma_load_3d(sb, map, &full[stage], kc * 64, np * 256, fit);
                                else
                                    for (int i = 0; i < st.n / 64; ++i)
                                        tma_load_3d(sb + i * 8192, map, &full[stage], np * 256 + i * 64, kc * 64, fit);
                                if (++stage == STAGES) { stage = 0; phase ^= 1; }
                            }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================== MMA issuer
        // Step s contracts what the epilogue wrote in step s-1 (warp 2 stores the same buffer to global
        // meanwhile).  Step `nsteps` has no MMA: it only acknowledges the phase (see below).
        int stage = 0; uint32_t phase = 0;
        uint32_t rdy_phase = 0;                       // bit `slot` = parity of act_ready[slot] / mma_ready[slot]
        for (int round = 0;; ++round) {
            if (tile_of(round, 0) >= total_tiles) break;
            for (int s = 1; s <= (FWD ? nsteps - 1 : nsteps); ++s) {
                const Step st = step_info<H>(s < nsteps ? s : 1, L, D);
                const uint32_t idesc = (CL == 2) ? make_idesc_m(256, st.n, st.mn != 0) : make_idesc(st.n, false, st.mn != 0);
                const uint32_t b_lbo = st.mn ? 8192u : 0u;
                const uint32_t b_kadv = st.mn ? (UMMA_K * 128) >> 4 : (UMMA_K * 2) >> 4;
                for (int slot = 0; slot < NSLOT; ++slot) {
                    const int tile = tile_of(round, slot);
                    if (tile >= total_tiles) break;
                    if (CL == 2 && crank != 0) continue;         // peer CTA: the leader issues the pair's MMAs
                    mbar_wait(CL == 2 ? &mma_ready[slot] : &act_ready[slot], (rdy_phase >> slot) & 1u);
                    rdy_phase ^= 1u << slot;
                    tc_fence_after();
                    uint8_t* const act = smem + slot * C::ACT_BYTES;
                    if (s == nsteps) {
                        // Layer-0 gradient of this tile: the buffer holds dz_0 [128 rows x H] bf16.  Read as an MN-major A
                        // operand (M = 128 columns of dz_0 per MMA, K = the 128 rows) and contracted with the [16 x 128]
                        // operand {1, x_hi, x_mid, x_lo} of the tile's positions it gives sum_r dz0[r][j] and
                        // sum_r dz0[r][j] x[r] (x exact: three bf16 pieces) in 16 accumulator columns per 128 columns of
                        // dz_0 -- dz_0 never goes to HBM and there is no separate gradient kernel.  The commit on
                        // acc_full also acknowledges the phase: the next tile's E0 waits for it (a fast slot must not
                        // complete act_ready twice before this warp looked at it: parity waits alias after two phases),
                        // drains these columns, and only then rewrites the buffer.
                        constexpr int NH = (H >= 128) ? H / 128 : 1;
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        if (lane == 0) {
                            const uint32_t act_u32 = smem_u32(act);
                            const uint32_t xb = smem_u32(smem_ring + stage * STAGE_SZ);
                            const uint32_t xidesc = (CL == 2) ? make_idesc_m(256, 2 * XOP_N, false, true) : make_idesc(XOP_N, true, false);
#pragma unroll
                            for (int mh = 0; mh < NH; ++mh) {
                                const uint32_t d_tmem = tmem_base + slot * C::ACC_COLS + mh * (CL * XOP_N);
#pragma unroll
                                for (int k = 0; k < BM / UMMA_K; ++k) {
                                    const uint64_t adesc = make_desc(act_u32 + mh * 2 * CHUNK_BYTES + k * (UMMA_K * 128), CHUNK_BYTES, 1024);
                                    const uint64_t bdesc = make_desc(xb + (k >> 2) * (XOP_BYTES / 2) + (k & 3) * (UMMA_K * 2), 0, 1024);
                                    if (CL == 2) tc_mma_bf16_2cta(d_tmem, adesc, bdesc, xidesc, k > 0 ? 1u : 0u);
                                    else tc_mma_bf16(d_tmem, adesc, bdesc, xidesc, k > 0 ? 1u : 0u);
                                }
                            }
                            if (CL == 2) { tc_commit_2cta(&empty[stage], 3); tc_commit_2cta(&acc_full[slot], 3); }
                            else { tc_commit(&empty[stage]); tc_commit(&acc_full[slot]); }
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    const uint32_t act_u32 = smem_u32(act);
                    for (int np = 0; np < st.nparts; ++np) {
                        const uint32_t d_tmem = tmem_base + slot * C::ACC_COLS + np * 256;
                        for (int kc = 0; kc < st.kch; ++kc) {
                            mbar_wait(&full[stage], phase);
                            tc_fence_after();
                            if (lane == 0) {
                                const uint64_t adesc0 = make_desc(act_u32 + kc * CHUNK_BYTES, 0, 1024);
                                const uint64_t bdesc0 = make_desc(smem_u32(smem_ring + stage * STAGE_SZ), b_lbo, 1024);
                                const bool last = np == st.nparts - 1 && kc == st.kch - 1;
                                if (CL == 2) {
#pragma unroll
                                    for (int k = 0; k < 64 / UMMA_K; ++k)
                                        tc_mma_bf16_2cta(d_tmem, adesc0 + (uint64_t)(k * 2), bdesc0 + (uint64_t)(k * b_kadv), idesc,
                                                         (kc > 0 || k > 0) ? 1u : 0u);
                                    tc_commit_2cta(&empty[stage], 3);
                                    if (last) tc_commit_2cta(&acc_full[slot], 3);
                                } else {
#pragma unroll
                                    for (int k = 0; k < 64 / UMMA_K; ++k)
                                        tc_mma_bf16(d_tmem, adesc0 + (uint64_t)(k * 2), bdesc0 + (uint64_t)(k * b_kadv), idesc,
                                                    (kc > 0 || k > 0) ? 1u : 0u);
                                    tc_commit(&empty[stage]);
                                    if (last) tc_commit(&acc_full[slot]);
                                }
                            }
                            __syncwarp();
                            if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 2 && !FWD) {
        // ===================================================== store warp: every finished operand buffer is also a dW
        // operand (h_l / dY / dz_l): TMA-store it to global while the tensor core contracts it, and tell the
        // epilogue when the store has read the buffer (its own warp, so that this wait never delays an MMA)
        uint32_t rdy_phase = 0;
        const uint64_t pol_stream = policy_evict_first();   // read by a later kernel: do not displace the cos scratch
        for (int round = 0;; ++round) {
            if (tile_of(round, 0) >= total_tiles) break;
            for (int s = 1; s <= nsteps; ++s) {
                const int ps = s - 1;                                // the step whose output is stored
                const CUtensorMap* omap = (ps <= L) ? &maps.hout[ps] : (ps == L + 1) ? &maps.yout : &maps.zout[2 * L + 2 - ps];
                const int ochunks = (ps == L + 1) ? D / 64 : H / 64;
                for (int slot = 0; slot < NSLOT; ++slot) {
                    const int tile = tile_of(round, slot);
                    if (tile >= total_tiles) break;
                    const int fit = tile / g.mtiles, mt = tile - fit * g.mtiles;
                    mbar_wait(&act_ready[slot], (rdy_phase >> slot) & 1u);
                    rdy_phase ^= 1u << slot;
                    if (lane == 0) {
                        if (ps < nsteps - 1 && !(NA_DBG(g) & 1)) {       // dz_0 (the last step) is consumed on the SM: no store
                            const uint8_t* act = smem + slot * C::ACT_BYTES;
                            for (int kc = 0; kc < ochunks; ++kc)
                                tma_store_3d(omap, act + kc * CHUNK_BYTES, kc * 64, ((NA_DBG(g) & 4) ? (int)(blockIdx.x % g.mtiles) : mt) * BM, (NA_DBG(g) & 4) ? 0 : fit, pol_stream);
                            tma_store_commit();
                            tma_store_wait_read();
                        }
                        mbar_arrive(&buf_free[slot]);
                    }
                    __syncwarp();
                }
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // stores complete before the CTA retires
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        // ===================================================== epilogue warps
        // All 16 warps take the two slots in turn, step by step: (s, slot 0), (s, slot 1), (s+1, slot 0), ..  While they
        // are busy with one slot the tensor core contracts what they just wrote for the other one, so an MMA (and the
        // TMA store of its operand) is hidden behind half a step of epilogue work instead of stalling half the warps.
        const int ei = warp - NCTRL;
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int cg = ei >> 2;                       // column group inside the slot
        const int r = q * 32 + lane;                  // row inside the tile
        constexpr int NU = C::NU;
        constexpr int PFD = (NU < 4) ? NU : 4;        // cos units in flight in the backward epilogue
        const int col0 = cg * C::CW;
        const uint64_t pol_keep = policy_evict_last();
        const bool mufu_hidden = (g.sincos_mode & 1) != 0, mufu_l0 = (g.sincos_mode & 2) != 0;
        // cos of layer 0 is recomputed in the last backward step from the position and the prescaled W0 / b0 (one FFMA,
        // a reduction and one MUFU per element) instead of being parked in the scratch: a third of the scratch traffic
        // (and its DRAM spill) for the price of SFU work moved from the longest step into the shortest
        constexpr bool recompute_cos0 = !FWD && NA_RECOMPUTE_COS0;
        // One slot and H <= 256: the accumulator needs 256 of the 512 TMEM columns; the other 256 park cos_l of the hidden
        // layers (two bf16 per 32-bit cell, H / 2 columns per layer, the thread's own lane and columns) instead of the
        // global scratch -- written with tcgen05.st in the sine epilogue, read back with tcgen05.ld next to the
        // accumulator in the backward half: no L2 / DRAM round trip at all.  Layers that do not fit (deep: the third
        // hidden layer at H = 256) keep using the scratch.
        constexpr bool COSTM = !FWD && NS == 1 && H <= 256;
        constexpr int COSTM_LAYERS = COSTM ? 256 / (H / 2) : 0;
        const uint32_t t_cos0 = tmem_base + 256 + ((uint32_t)(q * 32) << 16) + col0 / 2;     // layer l at + (l - 1) * H / 2
        // cos scratch of one layer: [H/16 units][128 rows][16 columns], so that the 32 lanes of a warp (consecutive
        // rows, 32 B each) touch 1 KB of contiguous memory per access instead of 32 lines at row stride
        constexpr int SCR_U = BM * 16;                // elements between consecutive 16-column units
        // Parity of the next wait on acc_full[slot] (and, training, on buf_free[slot]: always waited for in the same places):
        // a slot waits once per step with an MMA behind it, so the parity follows from the round and the step -- no
        // per-slot phase registers to carry (and spill) through the loop
        auto wait_parity = [&](int round_, int s_) -> uint32_t {
            return FWD ? (uint32_t)(round_ * (nsteps - 1) + s_ - 1) & 1u : (uint32_t)(round_ * nsteps + s_ - 1) & 1u;
        };
        SlotRec* const wrec = reinterpret_cast<SlotRec*>(reinterpret_cast<uint8_t*>(bars) + BAR_BYTES) + ei * 2;
        // Layer-0 gradient partials of the tile a slot has just finished: the MMA warp left sum_r dz0[r][j] {1, x_hi, x_mid,
        // x_lo}[r] in 16 accumulator columns per 128 columns of dz_0 (lane = column j); the column group cg drains the
        // cg-th block.  Per-tile partials, summed over the row tiles in a fixed order by adam_kernel: deterministic.
        constexpr int L0_NH = (H >= 128) ? H / 128 : 1;
        auto drain_l0grad = [&](int slot, int pfit, int pmt) {
            if (cg < L0_NH) {
                uint32_t v[16];
                tmem_ld16(tmem_base + slot * C::ACC_COLS + ((uint32_t)(q * 32) << 16) + cg * (CL * XOP_N) + (int)crank * XOP_N, v);
                tmem_ld_wait();
                const int j = cg * 128 + q * 32 + lane;
                if (j < H) {
                    const size_t o = ((size_t)pfit * g.mtiles + pmt) * H + j;
                    g.colpart0[o] = __uint_as_float(v[0]);
                    g.xpart[o] = (__uint_as_float(v[1]) + __uint_as_float(v[2])) + __uint_as_float(v[3]);
                }
            }
        };
#ifdef NA_CHAIN_TIMING
        long long t_acc = 0, t_kind[4] = {0, 0, 0, 0}, t_begin = clock64(), tq = 0, ts = 0, t_acc0 = 0;
#define NA_T0() tq = clock64()
#define NA_T1() t_acc += clock64() - tq
#else
#define NA_T0()
#define NA_T1()
#endif
        int round = 0;
        for (;; ++round) {
            // the tiles of this round (per slot), resolved once: the step bodies below run 2 x nsteps times per round
            const int tile0 = tile_of(round, 0), tile1 = (NSLOT > 1) ? tile_of(round, 1) : total_tiles;
            if (tile0 >= total_tiles) break;
            const bool two = tile1 < total_tiles;
            __syncwarp();                                      // every lane is done with the previous round's records
            if (lane < NSLOT && (lane == 0 || two)) {
                const int tile = lane ? tile1 : tile0;
                SlotRec sr;
                sr.fit = tile / g.mtiles; sr.mt = tile - sr.fit * g.mtiles;
                const FitRec& fr = g.recs[sr.fit];
                sr.omega = fr.omega; sr.pad0 = 0;
                sr.pos = fr.pos; sr.tnorm = fr.tnorm; sr.obias = fr.params + g.b_off[L + 1]; sr.pad1 = nullptr;
                wrec[lane] = sr;
            }
            __syncwarp();
            for (int s = 0; s < nsteps; ++s) {
#pragma unroll 1
            for (int slot = 0; slot < NSLOT; ++slot) {
                if (slot == 1 && !two) break;
                const SlotRec* const rec = &wrec[slot];
                const int fit = rec->fit, mt = rec->mt;
                const int row = mt * BM + r;
                const bool row_ok = row < g.N;                 // ragged last tile
                const int row_c = row_ok ? row : g.N - 1;      // a valid row to read inputs from
                const float omega = rec->omega;
                const uint32_t act_u32 = smem_u32(smem + slot * C::ACT_BYTES);
                const uint32_t t_row = tmem_base + slot * C::ACC_COLS + ((uint32_t)(q * 32) << 16) + col0;
                __nv_bfloat16* const scr = g.scratch + ((size_t)(blockIdx.x * NSLOT + slot) * (L + 1)) * (BM * H)
                                           + (size_t)(col0 / 16) * SCR_U + r * 16;
#ifdef NA_CHAIN_TIMING
                ts = clock64(); t_acc0 = t_acc;
#endif
                if (s == 0) {
                    // ---------------- layer 0: outer product + sine, fp32 (siren.py:33-34 with in_features = 1)
                    // The position and the first omega-prescaled weights / biases (kept current by Adam: one FFMA per argument)
                    // are requested before the waits below: they come from L2 and would otherwise stall all four warps of a
                    // scheduler at the top of every tile
                    const float x = __ldg(rec->pos + row_c);
                    const float* w0 = g.psc + (size_t)fit * g.psc_fit + col0;
                    const float* b0 = g.psc + (size_t)fit * g.psc_fit + H + col0;
                    float4 wn[2], bn[2];                         // weights / biases of the next 8 columns
                    wn[0] = __ldg(reinterpret_cast<const float4*>(w0)); wn[1] = __ldg(reinterpret_cast<const float4*>(w0) + 1);
                    bn[0] = __ldg(reinterpret_cast<const float4*>(b0)); bn[1] = __ldg(reinterpret_cast<const float4*>(b0) + 1);
                    // the buffer is free once the MMA warp has stored the previous tile's dz_0
                    if (!FWD && round > 0) {
                        NA_T0();
                        const uint32_t par = wait_parity(round, s); mbar_wait(&acc_full[slot], par);       // the layer-0 gradient MMA of the previous tile is complete
                        mbar_wait(&buf_free[slot], par);     // the store warp has seen the last step
                        NA_T1();
                        tc_fence_after();
                        const int ptile = tile_of(round - 1, slot), pfit = ptile / g.mtiles;
                        drain_l0grad(slot, pfit, ptile - pfit * g.mtiles);
                    }
#pragma unroll 1
                    for (int u = 0; u < NU; ++u) {
                        uint32_t so[8], co[8];
#pragma unroll
                        for (int gi = 0; gi < 2; ++gi) {
                            float arg[8], sn[8], cs[8];
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                arg[4 * j] = fmaf(x, wn[j].x, bn[j].x); arg[4 * j + 1] = fmaf(x, wn[j].y, bn[j].y);
                                arg[4 * j + 2] = fmaf(x, wn[j].z, bn[j].z); arg[4 * j + 3] = fmaf(x, wn[j].w, bn[j].w);
                            }
                            const int nxt = (u * 2 + gi + 1 < NU * 2) ? (u * 2 + gi + 1) * 8 : 0;
                            wn[0] = __ldg(reinterpret_cast<const float4*>(w0 + nxt)); wn[1] = __ldg(reinterpret_cast<const float4*>(w0 + nxt) + 1);
                            bn[0] = __ldg(reinterpret_cast<const float4*>(b0 + nxt)); bn[1] = __ldg(reinterpret_cast<const float4*>(b0 + nxt) + 1);
                            if (mufu_l0) sincos8<true>(arg, sn, cs); else sincos8<false>(arg, sn, cs);
#pragma unroll
                            for (int j = 0; j < 8; j += 2) {
                                so[(gi * 8 + j) / 2] = pack_bf16(sn[j], sn[j + 1]);
                                co[(gi * 8 + j) / 2] = pack_bf16(cs[j], cs[j + 1]);
                            }
                        }
                        act_store16(act_u32, r, col0 + u * 16, so);
                        if (!FWD && !recompute_cos0) st_global_256_hint(scr + u * SCR_U, co, pol_keep);
                    }
                } else if (s <= L) {
                    // ---------------- hidden sine layer s
                    const float* bsrc = g.psc + (size_t)fit * g.psc_fit + (size_t)(s + 1) * H + col0;
                    float4 bn[4];                                // bias of the next 16 columns (L1-resident)
#pragma unroll
                    for (int j = 0; j < 4; ++j) bn[j] = __ldg(reinterpret_cast<const float4*>(bsrc) + j);
                    NA_T0(); const uint32_t par = wait_parity(round, s); mbar_wait(&acc_full[slot], par);
                    if (!FWD) { mbar_wait(&buf_free[slot], par); }
                    NA_T1();
                    tc_fence_after();
                    __nv_bfloat16* const cdst = scr + (size_t)s * (BM * H);
                    uint32_t v[16];
                    float dot = 0.f;
                    const float prow = (MODE == 2 && s == L && row_ok) ? __ldg(g.pvec + (size_t)fit * g.N + row) : 0.f;
                    tmem_ld16(t_row, v);
#pragma unroll 1
                    for (int u = 0; u < NU; ++u) {
                        float arg[16];
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            arg[4 * j] = fmaf(__uint_as_float(v[4 * j]), omega, bn[j].x);
                            arg[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), omega, bn[j].y);
                            arg[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), omega, bn[j].z);
                            arg[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), omega, bn[j].w);
                        }
                        if (u + 1 < NU) {                        // in flight during the sincos below
                            tmem_ld16(t_row + (u + 1) * 16, v);
#pragma unroll
                            for (int j = 0; j < 4; ++j) bn[j] = __ldg(reinterpret_cast<const float4*>(bsrc + (u + 1) * 16) + j);
                        }
                        uint32_t so[8], co[8];
                        float wv[16];
#pragma unroll
                        for (int gi = 0; gi < 2; ++gi) {
                            float a8[8], sn[8], cs[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) a8[j] = arg[gi * 8 + j];
                            if (mufu_hidden) sincos8<true>(a8, sn, cs); else sincos8<false>(a8, sn, cs);
                            if (MODE == 2 && s == L) {     // decode, values: p_t * sin(.) summed over rows below
#pragma unroll
                                for (int j = 0; j < 8; ++j) wv[gi * 8 + j] = prow * sn[j];
                            } else if (MODE == 1 && s == L) {          // decode, keys: u . sin(.) of this row, fp32
                                const float* uv = g.dotvec + (size_t)fit * H + col0 + u * 16 + gi * 8;
                                const float4 u0 = __ldg(reinterpret_cast<const float4*>(uv)), u1 = __ldg(reinterpret_cast<const float4*>(uv) + 1);
                                dot = fmaf(u0.x, sn[0], dot); dot = fmaf(u0.y, sn[1], dot); dot = fmaf(u0.z, sn[2], dot); dot = fmaf(u0.w, sn[3], dot);
                                dot = fmaf(u1.x, sn[4], dot); dot = fmaf(u1.y, sn[5], dot); dot = fmaf(u1.z, sn[6], dot); dot = fmaf(u1.w, sn[7], dot);
                            }
#pragma unroll
                            for (int j = 0; j < 8; j += 2) {
                                so[(gi * 8 + j) / 2] = pack_bf16(sn[j], sn[j + 1]);
                                co[(gi * 8 + j) / 2] = pack_bf16(cs[j], cs[j + 1]);
                            }
                        }
                        if (!(FWD && s == L)) {
                            act_store16(act_u32, r, col0 + u * 16, so);
#ifndef NA_EXP_NOSCRATCH
                            if (COSTM && s <= COSTM_LAYERS) tmem_st8(t_cos0 + (s - 1) * (H / 2) + u * 8, co);
                            else if (!FWD) st_global_256_hint(cdst + u * SCR_U, co, pol_keep);
#endif
                        } else if (MODE == 2) {
                            // sum the 16 columns over the 32 rows of this warp: transpose-reduce, 16 shuffles
                            float w8[8], w4[4], w2[2];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float send = (lane & 16) ? wv[j] : wv[j + 8], keep = (lane & 16) ? wv[j + 8] : wv[j];
                                w8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float send = (lane & 8) ? w8[j] : w8[j + 4], keep = (lane & 8) ? w8[j + 4] : w8[j];
                                w4[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                            }
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const float send = (lane & 4) ? w4[j] : w4[j + 2], keep = (lane & 4) ? w4[j + 2] : w4[j];
                                w2[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                            }
                            const float send = (lane & 2) ? w2[0] : w2[1], keep = (lane & 2) ? w2[1] : w2[0];
                            float w1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                            w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
                            // lanes 2c and 2c+1 hold column c = (lane >> 1) & 15 of this unit
                            if (!(lane & 1))
                                g.pvpart[((size_t)fit * (g.mtiles * 4) + (size_t)mt * 4 + q) * H + col0 + u * 16 + ((lane >> 1) & 15)] = w1;
                        }
                    }
                    if (MODE == 1 && s == L && row_ok) g.dotpart[((size_t)fit * C::CG + cg) * g.N + row] = dot;
                } else if (s == L + 1) {
                    // ---------------- output layer: dY = 2 (y - t) / (N D), loss partial (siren.py:101)
                    const int ow = D / C::CG;                    // output columns of this thread
                    const int ocol0 = cg * ow;
                    const float* bsrc = rec->obias + ocol0;
                    const float* tn = rec->tnorm + (size_t)row_c * D + ocol0;
                    const float rmask = row_ok ? 1.f : 0.f;      // rows past the sequence: no loss, dY = 0, no gradient
                    const int nuo = ow / 16;
                    uint32_t ta[16], tb[16];                     // targets of this unit and the next: two units in flight
                    ld_global_nc_na_256(tn, &ta[0]); ld_global_nc_na_256(tn + 8, &ta[8]);
                    if (nuo > 1) { ld_global_nc_na_256(tn + 16, &tb[0]); ld_global_nc_na_256(tn + 24, &tb[8]); }
                    NA_T0(); const uint32_t par = wait_parity(round, s); mbar_wait(&acc_full[slot], par);
                    if (!FWD) { mbar_wait(&buf_free[slot], par); }
                    NA_T1();
                    tc_fence_after();
                    const uint32_t t_out = tmem_base + slot * C::ACC_COLS + ((uint32_t)(q * 32) << 16) + ocol0;
                    float sq = 0.f;
                    auto out_unit = [&](int u, uint32_t (&tt)[16]) {
                        uint32_t v[16];
                        tmem_ld16(t_out + u * 16, v);
                        float4 b4[4];                            // the unit's output bias: in flight together with the TMEM load
#pragma unroll
                        for (int j = 0; j < 4; ++j) b4[j] = __ldg(reinterpret_cast<const float4*>(bsrc + u * 16) + j);
                        tmem_ld_wait();
                        uint32_t dout[8];
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 bb = b4[j / 4];
                            const float e0 = rmask * ((__uint_as_float(v[j + 0]) + bb.x) - __uint_as_float(tt[j + 0]));
                            const float e1 = rmask * ((__uint_as_float(v[j + 1]) + bb.y) - __uint_as_float(tt[j + 1]));
                            const float e2 = rmask * ((__uint_as_float(v[j + 2]) + bb.z) - __uint_as_float(tt[j + 2]));
                            const float e3 = rmask * ((__uint_as_float(v[j + 3]) + bb.w) - __uint_as_float(tt[j + 3]));
                            sq = fmaf(e0, e0, sq); sq = fmaf(e1, e1, sq); sq = fmaf(e2, e2, sq); sq = fmaf(e3, e3, sq);
                            dout[j / 2] = pack_bf16(e0 * g.loss_scale, e1 * g.loss_scale);
                            dout[j / 2 + 1] = pack_bf16(e2 * g.loss_scale, e3 * g.loss_scale);
                        }
                        if (u + 2 < nuo) { ld_global_nc_na_256(tn + (u + 2) * 16, &tt[0]); ld_global_nc_na_256(tn + (u + 2) * 16 + 8, &tt[8]); }
                        act_store16(act_u32, r, ocol0 + u * 16, dout);
                    };
#pragma unroll 1
                    for (int u = 0; u < nuo; u += 2) {
                        out_unit(u, ta);
                        if (u + 1 < nuo) out_unit(u + 1, tb);
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
                    if (lane == 0) g.losspart[(size_t)fit * g.losspart_per_fit + mt * C::EPW + ei] = sq;
                } else {
                    // ---------------- backward: dz_{lp} = (dz_{lp+1} W_{lp+1}) * w cos_{lp}
                    const int lp = 2 * L + 2 - s;                // L, L-1, .., 0
                    if (lp == 0 && recompute_cos0) {
                        const float x = __ldg(rec->pos + row_c);
                        const float* w0 = g.psc + (size_t)fit * g.psc_fit + col0;
                        const float* b0 = g.psc + (size_t)fit * g.psc_fit + H + col0;
                        // weights / biases of the next 8 columns, requested one group ahead as in layer 0: with 228 KB of the
                        // SM's array given to shared memory these 32-byte loads come from L2 (~300 cycles), and a load issued
                        // right before its use stalled all four warps of a scheduler (4.5 % of the kernel's stall samples)
                        float4 wn[2], bn[2];
                        wn[0] = __ldg(reinterpret_cast<const float4*>(w0)); wn[1] = __ldg(reinterpret_cast<const float4*>(w0) + 1);
                        bn[0] = __ldg(reinterpret_cast<const float4*>(b0)); bn[1] = __ldg(reinterpret_cast<const float4*>(b0) + 1);
                        NA_T0(); const uint32_t par = wait_parity(round, s); mbar_wait(&acc_full[slot], par);
                        mbar_wait(&buf_free[slot], par);
                        NA_T1();
                        tc_fence_after();
                        uint32_t v[16];
                        tmem_ld16(t_row, v);
#pragma unroll 1
                        for (int u = 0; u < NU; ++u) {
                            float cs[16];
#pragma unroll
                            for (int gi = 0; gi < 2; ++gi) {         // the same arguments and sin/cos routine as layer 0 (s == 0)
                                float arg[8], sn[8], c8[8];
#pragma unroll
                                for (int j = 0; j < 2; ++j) {
                                    arg[4 * j] = fmaf(x, wn[j].x, bn[j].x); arg[4 * j + 1] = fmaf(x, wn[j].y, bn[j].y);
                                    arg[4 * j + 2] = fmaf(x, wn[j].z, bn[j].z); arg[4 * j + 3] = fmaf(x, wn[j].w, bn[j].w);
                                }
                                const int nxt = (u * 2 + gi + 1 < NU * 2) ? (u * 2 + gi + 1) * 8 : 0;
                                wn[0] = __ldg(reinterpret_cast<const float4*>(w0 + nxt)); wn[1] = __ldg(reinterpret_cast<const float4*>(w0 + nxt) + 1);
                                bn[0] = __ldg(reinterpret_cast<const float4*>(b0 + nxt)); bn[1] = __ldg(reinterpret_cast<const float4*>(b0 + nxt) + 1);
                                if (mufu_l0) sincos8<true>(arg, sn, c8); else sincos8<false>(arg, sn, c8);
#pragma unroll
                                for (int j = 0; j < 8; ++j) cs[gi * 8 + j] = c8[j];
                            }
                            tmem_ld_wait();
                            uint32_t dout[8];
#pragma unroll
                            for (int t = 0; t < 8; ++t)
                                dout[t] = pack_bf16(__uint_as_float(v[2 * t]) * (omega * cs[2 * t]),
                                                    __uint_as_float(v[2 * t + 1]) * (omega * cs[2 * t + 1]));
                            if (u + 1 < NU) tmem_ld16(t_row + (u + 1) * 16, v);
                            act_store16(act_u32, r, col0 + u * 16, dout);
                        }
                    } else if (COSTM && lp <= COSTM_LAYERS) {
                        const uint32_t t_c = t_cos0 + (lp - 1) * (H / 2);
                        NA_T0(); const uint32_t par = wait_parity(round, s); mbar_wait(&acc_full[slot], par);
                        mbar_wait(&buf_free[slot], par);
                        NA_T1();
                        tc_fence_after();
                        uint32_t va[16], vb[16], ca[8], cb[8];
                        tmem_ld16(t_row, va); tmem_ld8(t_c, ca);
#pragma unroll
                        for (int u = 0; u < NU; ++u) {
                            uint32_t (&v)[16] = (u & 1) ? vb : va;
                            uint32_t (&c)[8] = (u & 1) ? cb : ca;
                            tmem_ld_wait();
                            if (u + 1 < NU) {
                                tmem_ld16(t_row + (u + 1) * 16, (u & 1) ? va : vb);
                                tmem_ld8(t_c + (u + 1) * 8, (u & 1) ? ca : cb);
                            }
                            uint32_t dout[8];
#pragma unroll
                            for (int t = 0; t < 8; ++t) {
                                float c0, c1;
                                unpack_bf16(c[t], c0, c1);
                                dout[t] = pack_bf16(__uint_as_float(v[2 * t]) * (omega * c0),
                                                    __uint_as_float(v[2 * t + 1]) * (omega * c1));
                            }
                            act_store16(act_u32, r, col0 + u * 16, dout);
                        }
                    } else {
                    const __nv_bfloat16* csrc = scr + (size_t)lp * (BM * H);
                    uint32_t cc[PFD][8];
#ifdef NA_EXP_NOSCRATCH
#define NA_LD_COS(dst, src) do { for (int z_ = 0; z_ < 8; ++z_) (dst)[z_] = 0x3f003f00u + (uint32_t)lane; } while (0)
#else
#define NA_LD_COS(dst, src) ld_global_256_hint(src, dst, pol_keep)
#endif
#pragma unroll
                    for (int p = 0; p < PFD; ++p) NA_LD_COS(cc[p], csrc + p * SCR_U);
                    NA_T0(); const uint32_t par = wait_parity(round, s); mbar_wait(&acc_full[slot], par);
                    if (!FWD) { mbar_wait(&buf_free[slot], par); }
                    NA_T1();
                    tc_fence_after();
                    uint32_t va[16], vb[16];
                    tmem_ld16(t_row, va);
#pragma unroll
                    for (int u = 0; u < NU; ++u) {
                        uint32_t (&v)[16] = (u & 1) ? vb : va;
                        tmem_ld_wait();
                        if (u + 1 < NU) tmem_ld16(t_row + (u + 1) * 16, (u & 1) ? va : vb);
                        uint32_t dout[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            float c0, c1;
                            unpack_bf16(cc[u % PFD][t], c0, c1);
                            dout[t] = pack_bf16(__uint_as_float(v[2 * t]) * (omega * c0),
                                                __uint_as_float(v[2 * t + 1]) * (omega * c1));
                        }
                        if (u + PFD < NU) NA_LD_COS(cc[u % PFD], csrc + (u + PFD) * SCR_U);
                        act_store16(act_u32, r, col0 + u * 16, dout);
                    }
                    }
                }
                // operand buffer complete: hand it to the MMA warp (stores it to global, then contracts it)
#ifdef NA_CHAIN_TIMING
                t_kind[s == 0 ? 0 : s <= L ? 1 : s == L + 1 ? 2 : 3] += (clock64() - ts) - (t_acc - t_acc0);
#endif
                if (FWD && s == nsteps - 1) { tc_fence_before(); continue; }     // nothing consumes the last forward step
                if (COSTM && s >= 1 && s <= L && s <= COSTM_LAYERS) tmem_st_wait();
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy smem writes -> async proxy
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&act_ready[slot]);
                    if (CL == 2) mbar_arrive_cluster(&mma_ready[slot], 0);     // the pair's MMA is issued by the leader CTA
                }
            }
            }
        }
        if (!FWD) {
            // the last tile of each slot: its layer-0 gradient MMA is the last thing the tensor core does for this CTA
#pragma unroll 1
            for (int slot = 0; slot < NSLOT; ++slot) {
                const int ptile = (round > 0) ? tile_of(round - 1, slot) : total_tiles;
                if (ptile >= total_tiles) continue;
                const int pfit = ptile / g.mtiles;
                mbar_wait(&acc_full[slot], wait_parity(round, 0));        // as the next tile's layer 0 would have waited
                tc_fence_after();
                drain_l0grad(slot, pfit, ptile - pfit * g.mtiles);
            }
        }
#ifdef NA_CHAIN_TIMING
        if (lane == 0 && (ei == 0 || ei == 5 || ei == 15) && (blockIdx.x == 0 || blockIdx.x == 77))
            printf("chain timing cta %d warp %d: total %lld wait_acc %lld E0 %lld sine %lld out %lld dx %lld\n", (int)blockIdx.x,
                   ei, clock64() - t_begin, t_acc, t_kind[0], t_kind[1], t_kind[2], t_kind[3]);
#endif
    }

    tc_fence_before();
    __syncthreads();
    if (CL == 2) cluster_sync_all();              // the peer may still multicast into / arrive on this CTA's shared memory
    tc_fence_after();
    if (warp == 1) {
        if (CL == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------ host
// NERFATTN_CHAIN_SLOTS=1 forces one tile in flight per CTA (16 epilogue warps on it, deeper weight ring)
inline int slots_for(int H) {
    const char* e = getenv("NERFATTN_CHAIN_SLOTS");
    return (H <= 256 && !(e && atoi(e) == 1)) ? 2 : 1;
}
// NERFATTN_CLUSTER=1 enables the CTA-pair variant (cta_group::2 MMAs, each CTA loads half of every weight stage); it
// needs an even number of row tiles per fit so that a pair never straddles two fits
inline bool use_cluster(int N, int H = 256, int D = 128) {
    if (H < 128 || D < 128) return false;             // each CTA of the pair holds >= 64 rows / columns of every B operand
    // default: pair mode for H = 512 (one slot, 512 KB of weights per step: halving the per-SM weight stream gains
    // ~9 %); for H <= 256 the two-slot kernel is as fast without the pair's lock-step (measured, profiles/README.md)
    const char* e = getenv("NERFATTN_CLUSTER");
    const bool on = e ? atoi(e) != 0 : H >= 512;
    const char* sl = getenv("NERFATTN_CHAIN_SLOTS");
    return on && !(sl && atoi(sl) == 1) && ceil_div(N, BM) % 2 == 0 && num_sms() % 2 == 0;
}
inline size_t scratch_elems(int H, int L) {
    return (size_t)num_sms() * 2 * (L + 1) * BM * H;
}

// B operand of the layer-0 gradient MMA for every row tile of one position vector: xop[tile][j][r], r = row inside the
// tile, j = 0: 1, j = 1..3: the three bf16 pieces of x[r] (x_hi + x_mid + x_lo == x to fp32 precision, so the product
// with the bf16 dz_0 is exact), j >= 4: 0.  Rows past the sequence are all-zero.  grid (mtiles, tables), block 128.
__global__ void xop_kernel(const float* const* pos_tab, int N, int mtiles, __nv_bfloat16* xop) {
    const int row = blockIdx.x * BM + threadIdx.x;
    const bool ok = row < N;
    const float x = ok ? pos_tab[blockIdx.y][row] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(hi);
    const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
    const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
    __nv_bfloat16* dst = xop + ((size_t)blockIdx.y * mtiles + blockIdx.x) * (XOP_N * BM) + threadIdx.x;
    dst[0] = __float2bfloat16_rn(ok ? 1.f : 0.f);
    dst[BM] = hi; dst[2 * BM] = mid; dst[3 * BM] = lo;
    for (int j = 4; j < XOP_N; ++j) dst[j * BM] = __float2bfloat16_rn(0.f);
}
inline size_t xop_elems(int mtiles, int ntab) { return (size_t)ntab * mtiles * XOP_N * BM; }

inline int build_maps(int N, int D, int H, int L, int nf, const LayerMap& lm, __nv_bfloat16* wbf16,
                      void* const* act, void* const* dzs, void* dy, const __nv_bfloat16* xop, int xop_tiles, ChainMaps& m) {
    int rc;
    // [tile][16][128]: box {64 K, 16 rows}, two boxes per tile
    if ((rc = make_map(&m.xop, xop, BM, XOP_N, xop_tiles, BM, (uint64_t)XOP_N * BM, 64, XOP_N))) return rc;
    for (int l = 0; l <= L; ++l) {
        if ((rc = make_operand_map(&m.hout[l], act[l], N, H, nf, (size_t)N * H, false, BM))) return rc;
        // dz_0 is consumed on the SM (layer-0 gradient MMA): never stored, no map
        if (l > 0 && (rc = make_operand_map(&m.zout[l], dzs[l], N, H, nf, (size_t)N * H, false, BM))) return rc;
    }
    if ((rc = make_operand_map(&m.yout, dy, N, D, nf, (size_t)N * D, false, BM))) return rc;
    for (int l = 1; l <= L + 1; ++l) {
        const int rows = lm.out_dim[l];
        const int box = use_cluster(N, H, D) ? 64 : (l == L + 1) ? D : (H >= 256 ? 256 : H);
        if ((rc = make_operand_map(&m.wk[l], wbf16 + lm.w_off[l], rows, H, nf, lm.P, false, box))) return rc;
        if ((rc = make_operand_map(&m.wmn[l], wbf16 + lm.w_off[l], rows, H, nf, lm.P, true, 0))) return rc;
    }
    return NA_OK;
}

template <int H, int NS, int MODE, int CL>
inline cudaError_t launch_one(int grid, const ChainMaps& maps, const ChainArgs& a, cudaStream_t s) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = Cfg<H, NS>::SMEM; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, chain_kernel<H, NS, MODE, CL>, maps, a);
}
template <int H, int NS>
inline cudaError_t launch_mode(int mode, int grid, const ChainMaps& maps, const ChainArgs& a, cudaStream_t s) {
    return mode == 0 ? launch_one<H, NS, 0, 1>(grid, maps, a, s)
         : mode == 1 ? launch_one<H, NS, 1, 1>(grid, maps, a, s) : launch_one<H, NS, 2, 1>(grid, maps, a, s);
}
// mode: 0 training, 1 decode logits, 2 decode values.  The 2-CTA cluster variant exists for training only.
template <int H>
inline int launch_h(const ChainMaps& maps, const ChainArgs& a_in, int mode, cudaStream_t s, int max_ctas = 0, bool pack = false) {
    ChainArgs a = a_in;
    const int tiles = a.nf * a.mtiles;
    const bool cl = mode == 0 && use_cluster(a.N, H, a.D);
    int grid = std::min(tiles, (max_ctas > 0) ? std::min(max_ctas, num_sms()) : num_sms());
    if (cl) grid &= ~1;
    // A launch that cannot fill both slots of every CTA: with other kernels competing for the SMs (`pack`: several shape
    // groups in one call) two tiles on half the CTAs cost barely more time than one tile each -- the slots alternate -- and
    // half the SM-time (the 35-fit shards of the 8-GPU sweep: 0.245 -> 0.231 ms per epoch); alone, one tile per CTA is the
    // lower latency and stays
    if (pack && mode == 0 && !cl && H <= 256 && slots_for(H) == 2 && tiles < 2 * grid) grid = ceil_div(tiles, 2);
    a.pair_tiles = (tiles >= 2 * grid && !getenv("NERFATTN_NO_PAIR_TILES")) ? 1 : 0;
    cudaError_t e;
    constexpr int NSD = (H <= 256) ? 2 : 1;                  // default slots
    if (cl && slots_for(H) == NSD) e = launch_one<H, NSD, 0, 2>(grid, maps, a, s);
    else if constexpr (H <= 256) {
        if (slots_for(H) == 2) e = launch_mode<H, 2>(mode, grid, maps, a, s);
        else e = launch_mode<H, 1>(mode, grid, maps, a, s);
    } else e = launch_mode<H, 1>(mode, grid, maps, a, s);
    if (e != cudaSuccess) { set_error("chain_kernel launch failed: %s", cudaGetErrorString(e)); return NA_ERR_CUDA; }
    return NA_OK;
}
inline int launch(int H, const ChainMaps& maps, const ChainArgs& a, int mode, cudaStream_t s, int max_ctas = 0, bool pack = false) {
    switch (H) {
        case 64: return launch_h<64>(maps, a, mode, s, max_ctas, pack);
        case 128: return launch_h<128>(maps, a, mode, s, max_ctas, pack);
        case 256: return launch_h<256>(maps, a, mode, s, max_ctas, pack);
        case 512: return launch_h<512>(maps, a, mode, s, max_ctas, pack);
        default: set_error("chain: unsupported H %d", H); return NA_ERR_UNSUPPORTED;
    }
}

// NERFATTN_SINCOS: 0 = polynomial everywhere, 1 = MUFU core in the hidden layers only, 3 = in layer 0
// too (default: every result is rounded to bf16 at once; the reduction is exact for |x| <= 8192)
inline int sincos_mode() {
    const char* e = getenv("NERFATTN_SINCOS");
    return e ? (int)strtol(e, nullptr, 0) & 3 : 3;
}
// Profiling switches exist only in libnerfattn_prof.so (-DNA_PROFILING, `make libnerfattn_prof.so`); the release
// library always launches every kernel and has no debug addressing.
// NERFATTN_PHASE (results are meaningless): 1 = launch only the chain kernels of an epoch, 2 = only the dW GEMMs
// (+ their fused Adam), 16 = only the fit-resident kernels; 8 = none; 0 / 7 / unset = everything.  bench.py loads the
// profiling build beside the release one to time the dominant kernel alone, live, with CUDA events.
#ifdef NA_PROFILING
inline int phase_mask() {                          // + 16 = the fit-resident kernels (siren_resident.cuh); 7 / unset = everything
    const char* e = getenv("NERFATTN_PHASE");
    const int m = e ? atoi(e) : 0;
    return (m && m != 7) ? m : 23;
}
inline int chain_dbg() { const char* e = getenv("NERFATTN_CHAIN_DBG"); return e ? atoi(e) : 0; }
#else
constexpr int phase_mask() { return 23; }
constexpr int chain_dbg() { return 0; }
#endif

// One training epoch of one group: the chain, then the dW GEMMs (contraction over all rows of a fit)
// and the layer-0 gradient; Adam follows in the caller.
// omega-prescaled layer-0 weights and all sine-layer biases (one FFMA per sine argument): [n][(L + 2) * H]
__global__ void scale_params_kernel(const FitRec* recs, int H, int L, const int* /*unused*/, float* psc) {
    const FitRec& rec = recs[blockIdx.x];
    float* dst = psc + (size_t)blockIdx.x * (L + 2) * H;
    const int per = H * H + H;                                   // one hidden layer in the packed vector
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        dst[j] = rec.omega * rec.params[j];                      // W0
        dst[H + j] = rec.omega * rec.params[H + j];              // b0
        for (int l = 1; l <= L; ++l) dst[(l + 1) * H + j] = rec.omega * rec.params[2 * H + (l - 1) * per + H * H + j];
    }
}
inline size_t psc_floats(int H, int L) { return (size_t)(L + 2) * H; }
inline void scale_params(const FitRec* recs, int n, int H, int L, float* psc, cudaStream_t s) {
    scale_params_kernel<<<n, 256, 0, s>>>(recs, H, L, nullptr, psc);
}

// One training step of one group (or sub-batch of a group): forward, loss and the dX chain of every 128-row tile.
// The weight gradients and Adam follow in dw::dw_adam_kernel (siren_dw.cuh), the layer-0 parameters in adam_kernel.
inline int train_step(int N, int D, int H, int L, int nf, const LayerMap& lm, const FitRec* recs, const ChainMaps& cm,
                      __nv_bfloat16* scratch, float* losspart, int losspart_per_fit, int mtiles, const float* psc,
                      float* xpart, float* colpart0, int max_ctas, cudaStream_t s, bool pack = false) {
    ChainArgs a{};
    a.N = N; a.D = D; a.L = L; a.nf = nf; a.mtiles = mtiles; a.recs = recs;
    for (int l = 0; l <= L + 1; ++l) { a.w_off[l] = lm.w_off[l]; a.b_off[l] = lm.b_off[l]; }
    a.scratch = scratch;
    a.losspart = losspart; a.losspart_per_fit = losspart_per_fit;
    a.loss_scale = 2.0f / ((float)N * (float)D);
    a.sincos_mode = sincos_mode();
    a.psc = psc; a.psc_fit = (int)psc_floats(H, L);
    a.xpart = xpart; a.colpart0 = colpart0;
    a.dbg = chain_dbg();
    return launch(H, cm, a, 0, s, max_ctas, pack);
}

// Forward-only chain for the fused decode: partial scores [n][CG][N] (decode_finish sums them).
inline int decode_parts(int /*H*/) { return NEPI / 4; }      // column groups (Cfg::CG)
inline int build_fwd_maps(int /*N*/, int H, int L, int nf, const LayerMap& lm, __nv_bfloat16* wbf16, ChainMaps& m) {
    int rc;
    for (int l = 1; l <= L; ++l)
        if ((rc = make_operand_map(&m.wk[l], wbf16 + lm.w_off[l], H, H, nf, lm.P, false, H >= 256 ? 256 : H))) return rc;
    return NA_OK;
}
inline int launch_decode(int N, int D, int H, int L, int nf, const LayerMap& lm, const FitRec* recs, const ChainMaps& cm,
                         const float* psc, const float* u, float* dotpart, cudaStream_t s, const float* pvec = nullptr,
                         float* pvpart = nullptr) {
    ChainArgs a{};
    a.N = N; a.D = D; a.L = L; a.nf = nf; a.mtiles = ceil_div(N, BM); a.recs = recs;
    for (int l = 0; l <= L + 1; ++l) { a.w_off[l] = lm.w_off[l]; a.b_off[l] = lm.b_off[l]; }
    a.dotvec = u; a.dotpart = dotpart; a.pvec = pvec; a.pvpart = pvpart;
    a.psc = psc; a.psc_fit = (int)psc_floats(H, L);
    a.sincos_mode = sincos_mode();
    return launch(H, cm, a, pvec ? 2 : 1, s);
}

inline int configure_all() {
    static std::once_flag once_dev[kMaxDevices];
    static cudaError_t err_dev[kMaxDevices] = {};
    const int dev = current_device();
    cudaError_t& err = err_dev[dev];
    std::call_once(once_dev[dev], [&err] {
        auto acc = [&](cudaError_t e) { if (e != cudaSuccess && err == cudaSuccess) err = e; };
#define NA_CHAIN_CFG1(HH, NS, MODE, CL) \
        acc(cudaFuncSetAttribute(chain_kernel<HH, NS, MODE, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<HH, NS>::SMEM));
#define NA_CHAIN_CFG(HH, NS) NA_CHAIN_CFG1(HH, NS, 0, 1) NA_CHAIN_CFG1(HH, NS, 1, 1) NA_CHAIN_CFG1(HH, NS, 2, 1)
        NA_CHAIN_CFG(64, 2) NA_CHAIN_CFG(128, 2) NA_CHAIN_CFG(256, 2)
        NA_CHAIN_CFG(64, 1) NA_CHAIN_CFG(128, 1) NA_CHAIN_CFG(256, 1) NA_CHAIN_CFG(512, 1)
        NA_CHAIN_CFG1(64, 2, 0, 2) NA_CHAIN_CFG1(128, 2, 0, 2) NA_CHAIN_CFG1(256, 2, 0, 2) NA_CHAIN_CFG1(512, 1, 0, 2)
#undef NA_CHAIN_CFG1
#undef NA_CHAIN_CFG
    });
    if (err != cudaSuccess) { set_error("cudaFuncSetAttribute(chain smem) failed: %s", cudaGetErrorString(err)); return NA_ERR_CUDA; }
    return NA_OK;
}

}  // namespace chain
}  // namespace na
