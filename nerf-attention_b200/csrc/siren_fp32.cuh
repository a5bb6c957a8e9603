// FP32 SIMT kernels of the SIREN fit: the parity mode (NA_PREC_FP32) and the
// pieces both precisions share (normalisation, layer 0, Adam, final metrics).
//
// Reference semantics (paths relative to the reference tree):
//   normalisation      nerf_attention/siren.py:85-87
//   sine layer         nerf_attention/siren.py:33-34     sin(omega_0 * (x W^T + b))
//   loss               nerf_attention/siren.py:101       mean((y - t_norm)^2)
//   Adam               torch/optim/adam.py  _single_tensor_adam (non-capturable)
//   final metrics      nerf_attention/siren.py:119-125, 137-139
#pragma once

#include "common.cuh"

namespace na {
namespace f32 {

constexpr int BM = 128, BN = 128, BK = 16, NTHREADS = 256, PADM = 4;

enum GemmMode { kFwdSine = 0, kFwdOut = 1, kDx = 2, kDw = 3, kFwdEval = 4, kFwdDot = 5 };

// One launch = the same GEMM for every fit of a group.
struct GemmArgs {
    const FitRec* recs;
    int M, N, K;             // per-fit GEMM extents
    // operand A
    const float* A; size_t a_fit; int a_param_off; int lda;   // a_param_off >= 0: A = rec.params + off
    // operand B
    const float* B; size_t b_fit; int b_param_off; int ldb;
    // epilogue
    int bias_off;            // params offset of the bias added to the accumulator (fwd modes)
    float* out0; size_t out0_fit;        // fwd-sine: activation; fwd-out: dY; dx: dZ_prev; eval: y
    float* out1; size_t out1_fit;        // fwd-sine: cos
    const float* cprev; size_t cprev_fit;  // dx: cos of the previous layer
    float* colpart; size_t colpart_fit;  // [mtiles][N] column sums (bias grads)
    float* xpart; size_t xpart_fit;      // dx into layer 0: [mtiles][N] sum_m g*x[m]   (dW0)
    float* losspart; int losspart_per_fit;   // fwd-out: [mtiles*ntiles]
    float* gradpart; size_t grad_split_stride; size_t grad_fit; int grad_off;  // dw: [S][nf][P]
    int ksplit;              // dw: K rows per split
    float loss_scale;        // 2/(N*D)
    int denorm;              // eval: write y*std+mean instead of y
    float* const* yout;      // eval: per-fit output pointers (used when out0 == nullptr)
    const float* dotvec; size_t dotvec_fit;      // dot: u [N] per fit
    float* dotpart; size_t dotpart_fit;          // dot: [ntiles][M] partial row sums
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// A_KCONT: A(m,k) at A[m*lda + k]   else A(m,k) at A[k*lda + m]
// B_KCONT: B(k,n) at B[n*ldb + k]   else B(k,n) at B[k*ldb + n]
template <int MODE, bool A_KCONT, bool B_KCONT>
__global__ void __launch_bounds__(NTHREADS, 2)
sgemm_kernel(const GemmArgs g) {
    __shared__ __align__(16) float As[2][BK][BM + PADM];
    __shared__ __align__(16) float Bs[2][BK][BN + PADM];

    const int f = blockIdx.z;
    const FitRec& rec = g.recs[f];
    const int tid = threadIdx.x;
    const int n0 = blockIdx.x * BN;
    int mtile = blockIdx.y, split = 0;
    int kbeg = 0, kend = g.K;
    if (MODE == kDw) {
        const int mtiles = ceil_div(g.M, BM);
        split = blockIdx.y / mtiles;
        mtile = blockIdx.y % mtiles;
        kbeg = split * g.ksplit;
        kend = min(g.K, kbeg + g.ksplit);
    }
    const int m0 = mtile * BM;

    const float* A = (g.a_param_off >= 0) ? rec.params + g.a_param_off : g.A + (size_t)f * g.a_fit;
    const float* B = (g.b_param_off >= 0) ? rec.params + g.b_param_off : g.B + (size_t)f * g.b_fit;

    float4 ra[2], rb[2];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * NTHREADS;
            if (A_KCONT) {
                const int m = idx >> 2, kq = (idx & 3) * 4;
                ra[i] = (m0 + m < g.M && k0 + kq < kend)
                            ? ldg4(A + (size_t)(m0 + m) * g.lda + k0 + kq) : make_float4(0, 0, 0, 0);
            } else {
                const int k = idx >> 5, mq = (idx & 31) * 4;
                ra[i] = (k0 + k < kend && m0 + mq < g.M)
                            ? ldg4(A + (size_t)(k0 + k) * g.lda + m0 + mq) : make_float4(0, 0, 0, 0);
            }
            if (B_KCONT) {
                const int n = idx >> 2, kq = (idx & 3) * 4;
                rb[i] = (n0 + n < g.N && k0 + kq < kend)
                            ? ldg4(B + (size_t)(n0 + n) * g.ldb + k0 + kq) : make_float4(0, 0, 0, 0);
            } else {
                const int k = idx >> 5, nq = (idx & 31) * 4;
                rb[i] = (k0 + k < kend && n0 + nq < g.N)
                            ? ldg4(B + (size_t)(k0 + k) * g.ldb + n0 + nq) : make_float4(0, 0, 0, 0);
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * NTHREADS;
            if (A_KCONT) {
                const int m = idx >> 2, kq = (idx & 3) * 4;
                As[buf][kq + 0][m] = ra[i].x; As[buf][kq + 1][m] = ra[i].y;
                As[buf][kq + 2][m] = ra[i].z; As[buf][kq + 3][m] = ra[i].w;
            } else {
                const int k = idx >> 5, mq = (idx & 31) * 4;
                *reinterpret_cast<float4*>(&As[buf][k][mq]) = ra[i];
            }
            if (B_KCONT) {
                const int n = idx >> 2, kq = (idx & 3) * 4;
                Bs[buf][kq + 0][n] = rb[i].x; Bs[buf][kq + 1][n] = rb[i].y;
                Bs[buf][kq + 2][n] = rb[i].z; Bs[buf][kq + 3][n] = rb[i].w;
            } else {
                const int k = idx >> 5, nq = (idx & 31) * 4;
                *reinterpret_cast<float4*>(&Bs[buf][k][nq]) = rb[i];
            }
        }
    };

    const int tx = tid & 15, ty = tid >> 4;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    int buf = 0;
    load_tiles(kbeg);
    store_tiles(0);
    __syncthreads();
    for (int k0 = kbeg; k0 < kend; k0 += BK) {
        const bool more = k0 + BK < kend;
        if (more) load_tiles(k0 + BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (more) store_tiles(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

    // ---------------------------------------------------------------- epilogue
    // thread owns rows m0 + {ty*4+i, 64+ty*4+i}, cols n0 + {tx*4+j, 64+tx*4+j}
    float colsum[8], xsum[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { colsum[j] = 0.f; xsum[j] = 0.f; }
    float sq = 0.f;
    float rowdot[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) rowdot[i] = 0.f;

#pragma unroll
    for (int ih = 0; ih < 2; ++ih) {
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
            const int i = ih * 4 + ii;
            const int m = m0 + ih * 64 + ty * 4 + ii;
            if (m >= g.M) continue;
            float xm = 0.f;
            if (MODE == kDx && g.xpart) xm = __ldg(rec.pos + m);
#pragma unroll
            for (int jh = 0; jh < 2; ++jh) {
                const int n = n0 + jh * 64 + tx * 4;
                if (n >= g.N) continue;
                float r[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
                if (MODE == kFwdSine) {
                    const float4 bb = ldg4(rec.params + g.bias_off + n);
                    const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
                    float s[4], c[4], arg[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) arg[j] = rec.omega * (r[j] + bv[j]);
                    sincos_group(arg, s, c);
                    const size_t o = (size_t)m * g.N + n;
                    *reinterpret_cast<float4*>(g.out0 + (size_t)f * g.out0_fit + o) = make_float4(s[0], s[1], s[2], s[3]);
                    if (g.out1)
                        *reinterpret_cast<float4*>(g.out1 + (size_t)f * g.out1_fit + o) = make_float4(c[0], c[1], c[2], c[3]);
                } else if (MODE == kFwdDot) {
                    const float4 bb = ldg4(rec.params + g.bias_off + n);
                    const float4 uu = ldg4(g.dotvec + (size_t)f * g.dotvec_fit + n);
                    const float bv[4] = {bb.x, bb.y, bb.z, bb.w}, uv[4] = {uu.x, uu.y, uu.z, uu.w};
                    float s[4], c[4], arg[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) arg[j] = rec.omega * (r[j] + bv[j]);
                    sincos_group(arg, s, c);
#pragma unroll
                    for (int j = 0; j < 4; ++j) rowdot[i] = fmaf(uv[j], s[j], rowdot[i]);
                } else if (MODE == kFwdOut || MODE == kFwdEval) {
                    const float4 bb = ldg4(rec.params + g.bias_off + n);
                    const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
                    const size_t o = (size_t)m * g.N + n;
                    if (MODE == kFwdEval) {
                        float y[4] = {r[0] + bv[0], r[1] + bv[1], r[2] + bv[2], r[3] + bv[3]};
                        if (g.denorm) {
                            const float4 sd = ldg4(rec.stdv + n), mu = ldg4(rec.mean + n);
                            y[0] = fmaf(y[0], sd.x, mu.x); y[1] = fmaf(y[1], sd.y, mu.y);
                            y[2] = fmaf(y[2], sd.z, mu.z); y[3] = fmaf(y[3], sd.w, mu.w);
                        }
                        float* dst = g.out0 ? g.out0 + (size_t)f * g.out0_fit : g.yout[f];
                        *reinterpret_cast<float4*>(dst + o) = make_float4(y[0], y[1], y[2], y[3]);
                    } else {
                        const float4 tt = ldg4(rec.tnorm + o);
                        const float tv[4] = {tt.x, tt.y, tt.z, tt.w};
                        float d[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float diff = (r[j] + bv[j]) - tv[j];
                            sq = fmaf(diff, diff, sq);
                            d[j] = diff * g.loss_scale;
                            colsum[jh * 4 + j] += d[j];
                        }
                        *reinterpret_cast<float4*>(g.out0 + (size_t)f * g.out0_fit + o) = make_float4(d[0], d[1], d[2], d[3]);
                    }
                } else if (MODE == kDx) {
                    const size_t o = (size_t)m * g.N + n;
                    const float4 cc = ldg4(g.cprev + (size_t)f * g.cprev_fit + o);
                    const float cv[4] = {cc.x, cc.y, cc.z, cc.w};
                    float d[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        d[j] = r[j] * (rec.omega * cv[j]);
                        colsum[jh * 4 + j] += d[j];
                        xsum[jh * 4 + j] = fmaf(d[j], xm, xsum[jh * 4 + j]);
                    }
                    *reinterpret_cast<float4*>(g.out0 + (size_t)f * g.out0_fit + o) = make_float4(d[0], d[1], d[2], d[3]);
                } else {  // kDw
                    float* dst = g.gradpart + (size_t)split * g.grad_split_stride + (size_t)f * g.grad_fit +
                                 g.grad_off + (size_t)m * g.N + n;
                    *reinterpret_cast<float4*>(dst) = make_float4(r[0], r[1], r[2], r[3]);
                }
            }
        }
    }

    if (MODE == kFwdDot) {
        // row sums over this tile's 128 columns: the 16 tx lanes of a half-warp hold one row group
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = rowdot[i];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            const int m = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
            if (tx == 0 && m < g.M) g.dotpart[(size_t)f * g.dotpart_fit + (size_t)blockIdx.x * g.M + m] = v;
        }
    }
    if (MODE == kFwdOut || MODE == kDx) {
        // deterministic column sums over the tile's rows: 16 row-groups -> smem -> 128 threads
        float (*red)[BN] = reinterpret_cast<float (*)[BN]>(&As[0][0][0]);   // 16 x 128 floats = 8 KB
        __syncthreads();
        auto reduce_cols = [&](const float* vals, float* dst_base) {
#pragma unroll
            for (int jh = 0; jh < 2; ++jh)
#pragma unroll
                for (int j = 0; j < 4; ++j) red[ty][jh * 64 + tx * 4 + j] = vals[jh * 4 + j];
            __syncthreads();
            if (tid < BN && n0 + tid < g.N) {
                float s = 0.f;
#pragma unroll
                for (int r = 0; r < 16; ++r) s += red[r][tid];
                dst_base[(size_t)mtile * g.N + n0 + tid] = s;
            }
            __syncthreads();
        };
        reduce_cols(colsum, g.colpart + (size_t)f * g.colpart_fit);
        if (MODE == kDx && g.xpart) reduce_cols(xsum, g.xpart + (size_t)f * g.xpart_fit);
    }
    if (MODE == kFwdOut) {
        // block sum of squared errors -> one partial per tile, summed in fixed order later
        __shared__ float wsum[NTHREADS / 32];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        if ((tid & 31) == 0) wsum[tid >> 5] = sq;
        __syncthreads();
        if (tid == 0) {
            float s = 0.f;
            for (int w = 0; w < NTHREADS / 32; ++w) s += wsum[w];
            g.losspart[(size_t)f * g.losspart_per_fit + mtile * gridDim.x + blockIdx.x] = s;
        }
    }
}

// ---------------------------------------------------------------------------
// per-dimension mean / unbiased std / normalised copy of one unique target tensor
// grid (ceil(D/32), nuniq), block (32, 32)
struct NormArgs {
    const float* const* traw;   // [nuniq] device pointers
    float* tnorm; size_t tnorm_stride;   // [nuniq][N*D]
    float* mean; float* stdv;   // [nuniq][D]
    const int* prenorm;         // [nuniq] 1: targets already normalised, mean/std given
    int N, D;
};

__global__ void __launch_bounds__(1024) norm_kernel(const NormArgs a) {
    __shared__ float red[32][33];
    const int u = blockIdx.y;
    const int d = blockIdx.x * 32 + threadIdx.x;
    const float* t = a.traw[u];
    float* tn = a.tnorm + (size_t)u * a.tnorm_stride;
    const bool live = d < a.D;
    float mean, stdv;
    if (a.prenorm[u]) {
        mean = live ? a.mean[(size_t)u * a.D + d] : 0.f;
        stdv = live ? a.stdv[(size_t)u * a.D + d] : 1.f;
        for (int n = threadIdx.y; n < a.N; n += 32)
            if (live) tn[(size_t)n * a.D + d] = t[(size_t)n * a.D + d];
        return;
    }
    float s = 0.f;
    for (int n = threadIdx.y; n < a.N; n += 32) if (live) s += t[(size_t)n * a.D + d];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    s = 0.f;
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    mean = s / (float)a.N;
    __syncthreads();
    float q = 0.f;
    for (int n = threadIdx.y; n < a.N; n += 32)
        if (live) { const float e = t[(size_t)n * a.D + d] - mean; q = fmaf(e, e, q); }
    red[threadIdx.y][threadIdx.x] = q;
    __syncthreads();
    q = 0.f;
    for (int r = 0; r < 32; ++r) q += red[r][threadIdx.x];
    stdv = fmaxf(sqrtf(q / (float)max(a.N - 1, 1)), 1e-3f);       // unbiased, clamp(min=1e-3)
    if (live && threadIdx.y == 0) {
        a.mean[(size_t)u * a.D + d] = mean;
        a.stdv[(size_t)u * a.D + d] = stdv;
    }
    for (int n = threadIdx.y; n < a.N; n += 32)
        if (live) tn[(size_t)n * a.D + d] = (t[(size_t)n * a.D + d] - mean) / stdv;
}

// copy the per-tensor statistics to every fit's own mean/std output
__global__ void scatter_stats_kernel(const FitRec* recs, int D) {
    const FitRec& r = recs[blockIdx.x];
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        r.mean_out[d] = r.mean[d];
        r.std_out[d] = r.stdv[d];
    }
}

// ---------------------------------------------------------------------------
// layer 0: a0 = sin(w*(x*W0+b0)), c0 = cos(.)  -- an outer product, never a GEMM.
// OutT = float (fp32 mode) or __nv_bfloat16 (tensor mode feeds the MMA from it).
template <typename OutT>
__global__ void __launch_bounds__(256)
layer0_kernel(const FitRec* recs, int N, int H, OutT* act, OutT* cosb, size_t fit_stride) {
    const FitRec& rec = recs[blockIdx.y];
    const size_t i8 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // 8 consecutive features of one row
    const size_t total8 = (size_t)N * H / 8;
    if (i8 >= total8) return;
    const int n = (int)(i8 * 8 / H), j = (int)(i8 * 8 % H);
    const float x = __ldg(rec.pos + n);
    const float4 w0 = ldg4(rec.params + j), w1 = ldg4(rec.params + j + 4);
    const float4 b0 = ldg4(rec.params + H + j), b1 = ldg4(rec.params + H + j + 4);
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float arg[8], s[8], c[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) arg[k] = rec.omega * fmaf(x, wv[k], bv[k]);
    sincos_group(arg, s, c);
    const size_t o = (size_t)blockIdx.y * fit_stride + i8 * 8;
    if constexpr (sizeof(OutT) == 4) {
        *reinterpret_cast<float4*>(act + o) = make_float4(s[0], s[1], s[2], s[3]);
        *reinterpret_cast<float4*>(act + o + 4) = make_float4(s[4], s[5], s[6], s[7]);
        if (cosb) {
            *reinterpret_cast<float4*>(cosb + o) = make_float4(c[0], c[1], c[2], c[3]);
            *reinterpret_cast<float4*>(cosb + o + 4) = make_float4(c[4], c[5], c[6], c[7]);
        }
    } else {
        auto pk = [](float a, float b) { __nv_bfloat162 v = __floats2bfloat162_rn(a, b); return *reinterpret_cast<uint32_t*>(&v); };
        *reinterpret_cast<uint4*>(act + o) = make_uint4(pk(s[0], s[1]), pk(s[2], s[3]), pk(s[4], s[5]), pk(s[6], s[7]));
        if (cosb) *reinterpret_cast<uint4*>(cosb + o) = make_uint4(pk(c[0], c[1]), pk(c[2], c[3]), pk(c[4], c[5]), pk(c[6], c[7]));
    }
}

// ---------------------------------------------------------------------------
// Adam over the packed parameter vector of every fit of a group.
// grid (ceil(P/256), nf)
struct AdamArgs {
    const FitRec* recs;
    LayerMap lm;
    EpochTables et;
    // weight gradients of layers 1..L+1: [nsplit][nf][P] fp32 (indexed by param offset)
    const float* gradpart; size_t grad_split_stride; size_t grad_fit; int nsplit;
    // bias gradients: per layer [nf][mtiles][out_dim]; layer-0 weight gradient in xpart
    const float* colpart; size_t colpart_layer_off[kMaxLayers]; int col_mt[kMaxLayers];
    const float* xpart;
    // loss partials of this epoch
    const float* losspart; int losspart_per_fit; float loss_inv_count;
    float beta1, beta2, eps;
    // BF16 copies of the hidden/output weights for the tensor path (nullptr in fp32 mode)
    __nv_bfloat16* wbf16; size_t wbf16_fit;
    float* psc; size_t psc_fit; int H, L;   // chain mode: omega-prescaled copies of W0 / the sine-layer biases to refresh
    int p_end;              // parameters [0, p_end) are updated here (the rest in the fused dW + Adam kernel, siren_dw.cuh)
    int* epoch_rw; unsigned int* done;   // the last block to finish increments the group's epoch counter (done: arrival count)
};

// One parameter of torch _single_tensor_adam (non-capturable branch), in torch's operation order and with IEEE
// roundings: lerp_, mul_ + addcmul_, sqrt / bias_correction2_sqrt + eps, addcdiv_(value = -step_size).
// ob1 = 1 - beta1, ob2 = 1 - beta2, nss = -step_size.  Shared by adam_kernel and the fused dW epilogue (siren_dw.cuh).
__device__ __forceinline__ void adam_update(float g, float& m, float& v, float& w, float ob1, float beta2, float ob2,
                                            float eps, float bc2, float nss) {
    m = __fadd_rn(m, __fmul_rn(ob1, __fsub_rn(g, m)));
    v = __fadd_rn(__fmul_rn(v, beta2), __fmul_rn(__fmul_rn(ob2, g), g));
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), bc2), eps);
    w = __fadd_rn(w, __fdiv_rn(__fmul_rn(nss, m), denom));
}

__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs a) {
    const int f = blockIdx.y;
    const FitRec& rec = a.recs[f];
    const int e = *a.et.epoch;
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) * 4;      // 4 consecutive parameters (all regions are 4-aligned)

    if (blockIdx.x == 0 && threadIdx.x == 0) {          // siren.py:105  losses.append(loss.item())
        float s = 0.f;
        for (int i = 0; i < a.losspart_per_fit; ++i) s += a.losspart[(size_t)f * a.losspart_per_fit + i];
        rec.losses[e] = s * a.loss_inv_count;
    }
    if (p < a.p_end) {
        // which layer / weight-or-bias does p belong to
        int layer = 0;
#pragma unroll
        for (int i = 1; i < kMaxLayers; ++i)
            if (i < a.lm.nlayers && p >= a.lm.w_off[i]) layer = i;
        const bool is_bias = p >= a.lm.b_off[layer];
        const int width = a.lm.out_dim[layer];

        float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (is_bias || layer == 0) {
            const int j = is_bias ? p - a.lm.b_off[layer] : p;
            const int mt = is_bias ? a.col_mt[layer] : a.col_mt[0];
            const float* src = (is_bias ? a.colpart + a.colpart_layer_off[layer] : a.xpart) + (size_t)f * mt * width + j;
            for (int t = 0; t < mt; ++t) {
                const float4 v = ldg4(src + (size_t)t * width);
                g4.x += v.x; g4.y += v.y; g4.z += v.z; g4.w += v.w;
            }
        } else {
            const float* src = a.gradpart + (size_t)f * a.grad_fit + p;
            for (int s = 0; s < a.nsplit; ++s) {
                const float4 v = ldg4(src + (size_t)s * a.grad_split_stride);
                g4.x += v.x; g4.y += v.y; g4.z += v.z; g4.w += v.w;
            }
        }

        // torch _single_tensor_adam: lerp, mul+addcmul, sqrt/bc2_sqrt + eps, addcdiv(value=-step_size)
        const float4 m4 = *reinterpret_cast<const float4*>(rec.m + p), v4 = *reinterpret_cast<const float4*>(rec.v + p);
        const float4 w4 = *reinterpret_cast<const float4*>(rec.params + p);
        const float gs[4] = {g4.x, g4.y, g4.z, g4.w};
        float m[4] = {m4.x, m4.y, m4.z, m4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
        const float bc2 = a.et.bc2_sqrt[e], nss = -a.et.step_size[e];
#pragma unroll
        for (int i = 0; i < 4; ++i) adam_update(gs[i], m[i], v[i], w[i], 1.0f - a.beta1, a.beta2, 1.0f - a.beta2, a.eps, bc2, nss);
        *reinterpret_cast<float4*>(rec.m + p) = make_float4(m[0], m[1], m[2], m[3]);
        *reinterpret_cast<float4*>(rec.v + p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(rec.params + p) = make_float4(w[0], w[1], w[2], w[3]);
        if (a.psc && (layer == 0 || (is_bias && layer <= a.L))) {       // what the chain kernel's sine arguments read
            float* dst = a.psc + (size_t)f * a.psc_fit + (layer == 0 ? p : (layer + 1) * a.H + (p - a.lm.b_off[layer]));
            *reinterpret_cast<float4*>(dst) = make_float4(rec.omega * w[0], rec.omega * w[1], rec.omega * w[2], rec.omega * w[3]);
        }
        if (a.wbf16 && layer >= 1 && !is_bias) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(w[0], w[1]), hi = __floats2bfloat162_rn(w[2], w[3]);
            *reinterpret_cast<uint2*>(a.wbf16 + (size_t)f * a.wbf16_fit + p) =
                make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
        }
    }
    // End of this group's epoch: every thread of the launch has read `e` by the time its block arrives here, so the
    // last block to arrive may advance the counter (and re-arm the arrival count for the next launch).
    if (a.done) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            const unsigned int nblocks = gridDim.x * gridDim.y;
            if (atomicAdd(a.done, 1u) == nblocks - 1) { *a.done = 0u; *a.epoch_rw = e + 1; }
        }
    }
}

// progress metrics of one logging point (siren.py:107-115): scalars[0] = RealMSE, scalars[1] = mean CosSim
__global__ void progress_copy_kernel(const FitRec* recs, int nf, float* progress) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nf) return;
    progress[2 * recs[k].fit_index] = recs[k].scalars[0];
    progress[2 * recs[k].fit_index + 1] = recs[k].scalars[1];
}


// ---------------------------------------------------------------------------
// final metrics (siren.py:119-125): one warp per row, then one block per fit.
// y = pred_norm [nf][N][D]
__global__ void __launch_bounds__(256)
row_metrics_kernel(const FitRec* recs, const float* y, size_t y_fit, int N, int D) {
    const FitRec& rec = recs[blockIdx.y];
    const int row = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= N) return;
    const float* yr = y + (size_t)blockIdx.y * y_fit + (size_t)row * D;
    const float* tr = rec.traw + (size_t)row * D;
    float pp = 0.f, tt = 0.f, se = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float pr = fmaf(yr[d], rec.stdv[d], rec.mean[d]);   // pred_norm * std + mean
        const float t = rec.prenorm ? fmaf(tr[d], rec.stdv[d], rec.mean[d]) : tr[d];   // pre-normalised targets: back to real units
        pp = fmaf(pr, pr, pp); tt = fmaf(t, t, tt);
        const float e = pr - t; se = fmaf(e, e, se);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        pp += __shfl_xor_sync(0xffffffffu, pp, o);
        tt += __shfl_xor_sync(0xffffffffu, tt, o);
        se += __shfl_xor_sync(0xffffffffu, se, o);
    }
    // torch cosine_similarity: sum((x/max(|x|,eps)) * (y/max(|y|,eps))), eps = 1e-8
    const float np = fmaxf(sqrtf(pp), 1e-8f), nt = fmaxf(sqrtf(tt), 1e-8f);
    float dot = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float pr = fmaf(yr[d], rec.stdv[d], rec.mean[d]);
        const float t = rec.prenorm ? fmaf(tr[d], rec.stdv[d], rec.mean[d]) : tr[d];
        dot = fmaf(pr / np, t / nt, dot);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    if (lane == 0) { rec.cos[row] = dot; rec.ppmse[row] = se / (float)D; }
}

// scalars: final_mse, cos_mean, cos_min, cos_std (unbiased)
__global__ void __launch_bounds__(256) fit_scalars_kernel(const FitRec* recs, int N) {
    const FitRec& rec = recs[blockIdx.x];
    __shared__ double sh[3][256];
    __shared__ float shmin[256];
    double s = 0, s_mse = 0;
    float mn = 3.4e38f;
    for (int i = threadIdx.x; i < N; i += 256) {
        const float c = rec.cos[i];
        s += c; s_mse += rec.ppmse[i]; mn = fminf(mn, c);
    }
    sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = s_mse; shmin[threadIdx.x] = mn;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            sh[0][threadIdx.x] += sh[0][threadIdx.x + o];
            sh[1][threadIdx.x] += sh[1][threadIdx.x + o];
            shmin[threadIdx.x] = fminf(shmin[threadIdx.x], shmin[threadIdx.x + o]);
        }
        __syncthreads();
    }
    const double mean = sh[0][0] / N;
    const double mse = sh[1][0] / N;
    const float cmin = shmin[0];
    __syncthreads();
    double q = 0;
    for (int i = threadIdx.x; i < N; i += 256) { const double e = rec.cos[i] - mean; q += e * e; }
    sh[2][threadIdx.x] = q;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) sh[2][threadIdx.x] += sh[2][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        rec.scalars[0] = (float)mse;
        rec.scalars[1] = (float)mean;
        rec.scalars[2] = cmin;
        rec.scalars[3] = (float)sqrt(sh[2][0] / (double)max(N - 1, 1));
        rec.scalars[4] = rec.scalars[5] = rec.scalars[6] = rec.scalars[7] = 0.f;
    }
}

// de-normalise a forward output in place: y = y*std + mean (nerfattn_siren_forward)
__global__ void denorm_kernel(float* y, const float* mean, const float* stdv, int N, int D) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)N * D) return;
    const int d = (int)(i % D);
    y[i] = fmaf(y[i], stdv[d], mean[d]);
}

}  // namespace f32
}  // namespace na
