// Decode-side kernels: attention logits q.K for one new token, either from the
// KV cache in HBM (the baseline) or from the SIREN that replaces it.
//
// Reference: evaluate.py:196-200 times the unfused SIREN forward and compares it
// with bytes/bandwidth constants (evaluate.py:210-213).  Here both sides are real
// kernels.  The SIREN side never materialises K:
//     q.(std*(Wf h + bf) + mean) = (Wf^T (q*std)).h + q.(std*bf + mean) = u.h + c0
// so the output layer (2 N H D flop, N D writes) collapses into one H-vector u per
// head and the last sine layer's epilogue reduces u.sin(.) per position.
#pragma once

#include "common.cuh"

namespace na {
namespace dec {

// ---------------------------------------------------------------------------
// Baseline: scores[i][t] = q[i] . K[i][t],  K fp16 [n][N][D] streamed once from HBM.
// G = D/16 lanes share a row: one 256-bit load (a full 32 B sector) per lane per row, streaming (no-allocate) loads.
// Persistent grid (a multiple of the SM count), grid-stride over batches of UNROLL consecutive rows per lane group,
// software-pipelined: the loads of batch i+1 are in flight while batch i is reduced, so the memory system never
// drains between iterations (the first version issued 8 rows, waited, computed, and only then asked for more:
// 4.05 TB/s at 134 MB per launch; profiles/README.md).  One 32-bit division per batch finds the head.
template <int G>
__global__ void __launch_bounds__(256)
kvread_qk_kernel(const uint32_t* __restrict__ K, const uint32_t* __restrict__ q, float* __restrict__ scores,
                 long long rows_total, int N) {
    constexpr int UNROLL = 4;
    const int lane_in_group = threadIdx.x % G;
    const long long group = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long stride = (long long)gridDim.x * blockDim.x / G * UNROLL;
    auto load = [&](long long r0, uint32_t (&kv)[UNROLL][8]) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long r = r0 + u;
            if (r < rows_total) {
                asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                             : "=r"(kv[u][0]), "=r"(kv[u][1]), "=r"(kv[u][2]), "=r"(kv[u][3]), "=r"(kv[u][4]),
                               "=r"(kv[u][5]), "=r"(kv[u][6]), "=r"(kv[u][7])
                             : "l"(K + (r * G + lane_in_group) * 8));
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) kv[u][j] = 0u;
            }
        }
    };
    long long head_cached = -1;
    uint32_t qv[8];
    auto reduce = [&](long long r0, const uint32_t (&kv)[UNROLL][8]) {
        long long head = r0 / N;
        int rem = (int)(r0 - head * N);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long long r = r0 + u;
            if (rem == N) { rem = 0; ++head; }
            ++rem;
            if (head != head_cached && r < rows_total) {
                const uint4 a = __ldg(reinterpret_cast<const uint4*>(q + (head * G + lane_in_group) * 8));
                const uint4 b = __ldg(reinterpret_cast<const uint4*>(q + (head * G + lane_in_group) * 8 + 4));
                qv[0] = a.x; qv[1] = a.y; qv[2] = a.z; qv[3] = a.w; qv[4] = b.x; qv[5] = b.y; qv[6] = b.z; qv[7] = b.w;
                head_cached = head;
            }
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 kf = __half22float2(*reinterpret_cast<const __half2*>(&kv[u][j]));
                const float2 qf = __half22float2(*reinterpret_cast<const __half2*>(&qv[j]));
                acc = fmaf(kf.x, qf.x, acc);
                acc = fmaf(kf.y, qf.y, acc);
            }
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o, G);
            if (lane_in_group == 0 && r < rows_total) scores[r] = acc;
        }
    };
    uint32_t ka[UNROLL][8], kb[UNROLL][8];
    long long r0 = group * UNROLL;
    if (r0 < rows_total) load(r0, ka);
    while (r0 < rows_total) {                                   // two batches per trip: ka / kb alternate as the prefetch target
        const long long r1 = r0 + stride;
        if (r1 < rows_total) load(r1, kb);
        reduce(r0, ka);
        if (r1 >= rows_total) break;
        const long long r2 = r1 + stride;
        if (r2 < rows_total) load(r2, ka);
        reduce(r1, kb);
        r0 = r2;
    }
}

// ---------------------------------------------------------------------------
// u[i][j] = sum_d q[d] std[d] Wf[d][j],  c0[i] = sum_d q[d] (std[d] bf[d] + mean[d])
__global__ void __launch_bounds__(256)
decode_prep_kernel(const FitRec* recs, const __half* q, int D, int H, int wf_off, int bf_off, float* u, float* c0) {
    const FitRec& rec = recs[blockIdx.x];
    const __half* qi = q + (size_t)blockIdx.x * D;
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        float s = 0.f;
        for (int d = 0; d < D; ++d)
            s = fmaf(__half2float(qi[d]) * rec.stdv[d], rec.params[wf_off + (size_t)d * H + j], s);
        u[(size_t)blockIdx.x * H + j] = s;
    }
    if (threadIdx.x < 32) {
        float s = 0.f;
        for (int d = threadIdx.x; d < D; d += 32)
            s = fmaf(__half2float(qi[d]), fmaf(rec.stdv[d], rec.params[bf_off + d], rec.mean[d]), s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) c0[blockIdx.x] = s;
    }
}

// scores[i][t] = c0[i] + sum_p part[i][p][t]
__global__ void decode_finish_kernel(const float* part, int nparts, int N, const float* c0, float* const* scores) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N) return;
    const float* p = part + (size_t)blockIdx.y * nparts * N + t;
    float s = c0[blockIdx.y];
    for (int k = 0; k < nparts; ++k) s += p[(size_t)k * N];
    scores[blockIdx.y][t] = s;
}

// hidden_layers == 0: the whole network is layer 0, so fuse it with the dot product.
__global__ void __launch_bounds__(256)
l0dot_kernel(const FitRec* recs, int N, int H, const float* u, const float* c0, float* const* scores) {
    const FitRec& rec = recs[blockIdx.y];
    const int row = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= N) return;
    const float x = rec.pos[row];
    float s = 0.f;
    for (int j = lane; j < H; j += 32)
        s = fmaf(u[(size_t)blockIdx.y * H + j], sinf(rec.omega * fmaf(x, rec.params[j], rec.params[H + j])), s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) scores[blockIdx.y][row] = s + c0[blockIdx.y];
}

// ---------------------------------------------------------------------------
// Attention over the cached positions for one new token (the decode integration the
// reference describes, README.md:3-8, but never builds).  scores -> softmax -> sum_t p_t V_t.

// p[i][:] = softmax(scale * scores[i][:]) in place; one block per head.
__global__ void __launch_bounds__(1024) softmax_kernel(float* s, int N, float scale) {
    __shared__ float red[32];
    float* row = s + (size_t)blockIdx.x * N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float m = -3.0e38f;
    for (int t = threadIdx.x; t < N; t += blockDim.x) m = fmaxf(m, row[t] * scale);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) red[warp] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float l = 0.f;
    for (int t = threadIdx.x; t < N; t += blockDim.x) { const float e = expf(row[t] * scale - m); row[t] = e; l += e; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
    if (lane == 0) red[warp] = l;
    __syncthreads();
    l = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) l += red[w];             // fixed order: deterministic
    const float inv = 1.0f / l;
    for (int t = threadIdx.x; t < N; t += blockDim.x) row[t] *= inv;
}

// partial[i][c][d] = sum_{t in chunk c} p[i][t] V[i][t][d]; V streamed once (fp16 KV cache, or the fp32
// reconstruction of the fp32 SIREN path).  256 rows per chunk, D/8 lanes per row, 16/32-byte loads.
constexpr int kPvChunk = 256;
template <typename T>
__global__ void __launch_bounds__(256) pv_kernel(const T* __restrict__ V, const float* __restrict__ p, float* __restrict__ partial,
                                                 int N, int D, int chunks) {
    extern __shared__ float pv_red[];                    // [256 / LG][D]
    const int LG = D / 8, rows_per_pass = 256 / LG;
    const int lg = threadIdx.x % LG, rp = threadIdx.x / LG;
    const int i = blockIdx.y, c = blockIdx.x;
    const int t0 = c * kPvChunk, t1 = min(N, t0 + kPvChunk);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int t = t0 + rp; t < t1; t += rows_per_pass) {
        const float w = __ldg(p + (size_t)i * N + t);
        const T* src = V + ((size_t)i * N + t) * D + lg * 8;
        if constexpr (sizeof(T) == 2) {
            uint32_t v[4];
            asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "l"(src));
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&v[k]));
                acc[2 * k] = fmaf(w, f.x, acc[2 * k]); acc[2 * k + 1] = fmaf(w, f.y, acc[2 * k + 1]);
            }
        } else {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
            acc[0] = fmaf(w, a.x, acc[0]); acc[1] = fmaf(w, a.y, acc[1]); acc[2] = fmaf(w, a.z, acc[2]); acc[3] = fmaf(w, a.w, acc[3]);
            acc[4] = fmaf(w, b.x, acc[4]); acc[5] = fmaf(w, b.y, acc[5]); acc[6] = fmaf(w, b.z, acc[6]); acc[7] = fmaf(w, b.w, acc[7]);
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) pv_red[rp * D + lg * 8 + k] = acc[k];
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float s = 0.f;
        for (int q = 0; q < rows_per_pass; ++q) s += pv_red[q * D + d];
        partial[((size_t)i * chunks + c) * D + d] = s;
    }
}
__global__ void pv_finish_kernel(const float* partial, int chunks, int D, float* out) {
    const int i = blockIdx.x;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float s = 0.f;
        for (int c = 0; c < chunks; ++c) s += partial[((size_t)i * chunks + c) * D + d];
        out[(size_t)i * D + d] = s;
    }
}

// SIREN values, tensor path: g = sum of the chain kernel's per-warp partial sums of p_t * h_L(t) (fixed order),
// then out = std * (Wf g + bf) + mean  (sum_t p_t = 1): V is never materialised.
__global__ void __launch_bounds__(256)
attn_finish_kernel(const FitRec* recs, const float* pvpart, int nparts, int H, int D, int wf_off, int bf_off, float* out) {
    extern __shared__ float gsum[];                      // [H]
    const FitRec& rec = recs[blockIdx.x];
    const float* part = pvpart + (size_t)blockIdx.x * nparts * H;
    for (int j = threadIdx.x; j < H; j += blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < nparts; ++k) s += part[(size_t)k * H + j];
        gsum[j] = s;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float* w = rec.params + wf_off + (size_t)d * H;
        float s = 0.f;
        for (int j = 0; j < H; ++j) s = fmaf(w[j], gsum[j], s);
        out[(size_t)blockIdx.x * D + d] = fmaf(rec.stdv[d], s + rec.params[bf_off + d], rec.mean[d]);
    }
}

}  // namespace dec
}  // namespace na
