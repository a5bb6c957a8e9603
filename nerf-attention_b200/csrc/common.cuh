// Shared definitions for libnerfattn (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/nerfattn.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libnerfattn targets sm_100a (B200) only"
#endif

namespace na {

constexpr int kMaxHidden = 8;          // L <= kMaxHidden
constexpr int kMaxLayers = kMaxHidden + 2;

// Device-visible record of one fit (built on the host, uploaded once per call).
struct FitRec {
    const float* pos;       // [N]
    const float* traw;      // [N,D] raw targets
    const float* tnorm;     // [N,D] normalised targets (workspace, shared by fits of one tensor)
    const float* mean;      // [D]  (workspace copy, per unique tensor)
    const float* stdv;      // [D]
    float* params;          // [P]
    float* m;               // [P]
    float* v;               // [P]
    float* losses;          // [epochs]
    float* cos;             // [N]
    float* ppmse;           // [N]
    float* scalars;         // [8]
    float* mean_out;        // caller's [D]
    float* std_out;         // caller's [D]
    float omega;
    int prenorm;            // NA_FIT_TARGETS_PRENORMALISED: traw holds (t - mean) / std
    int posid;              // index of this fit's position vector among the distinct ones of its group
    int uniq;               // index of the unique target tensor
    int fit_index;          // position in the caller's job list
};

// Offsets of each layer inside the packed parameter vector.
// layer 0 = first sine layer (in=1), 1..L hidden sine layers, L+1 = output layer.
struct LayerMap {
    int nlayers;                 // L + 2
    int w_off[kMaxLayers];
    int b_off[kMaxLayers];
    int in_dim[kMaxLayers];
    int out_dim[kMaxLayers];
    int P;
};

inline LayerMap make_layer_map(int H, int L, int D) {
    LayerMap lm{};
    lm.nlayers = L + 2;
    int off = 0;
    for (int i = 0; i < L + 2; ++i) {
        int in = (i == 0) ? 1 : H;
        int out = (i == L + 1) ? D : H;
        lm.in_dim[i] = in;
        lm.out_dim[i] = out;
        lm.w_off[i] = off; off += in * out;
        lm.b_off[i] = off; off += out;
    }
    lm.P = off;
    return lm;
}

// Tables that change per epoch, read through the device-side epoch counter so
// that one captured CUDA graph can be replayed for every epoch.
struct EpochTables {
    const int* epoch;        // device counter, incremented by k_tick at the end of each epoch
    const float* step_size;  // [epochs]  lr_t / (1 - beta1^t)            (torch adam.py step_size)
    const float* bc2_sqrt;   // [epochs]  sqrt(1 - beta2^t)
};

void set_error(const char* fmt, ...);

#define NA_CUDA_OK(expr)                                                              \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            na::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                          __FILE__, __LINE__);                                        \
            return NA_ERR_CUDA;                                                       \
        }                                                                             \
    } while (0)

#define NA_LAUNCH_OK(what)                                                            \
    do {                                                                              \
        cudaError_t _e = cudaGetLastError();                                          \
        if (_e != cudaSuccess) {                                                      \
            na::set_error("launch of %s failed: %s (%s:%d)", what,                    \
                          cudaGetErrorString(_e), __FILE__, __LINE__);                \
            return NA_ERR_CUDA;                                                       \
        }                                                                             \
    } while (0)

// sin and cos of x to fp32 accuracy (<= 1.5 ulp measured for |x| <= 1000, oracle/README), with
// no branch: 3-constant Cody-Waite reduction by pi/2 + the minimax polynomials on [-pi/4, pi/4]
// + quadrant fix-up with integer ops.  omega_0 up to 60 puts first-layer arguments near 120, far
// outside MUFU.SIN's accurate range, and libdevice sincosf() carries a Payne-Hanek slow path whose
// branch stops the compiler from interleaving independent evaluations -- the sine epilogue is
// latency-bound without that interleaving.  Callers handle |x| > kSincosFastLimit separately.
constexpr float kSincosFastLimit = 8192.0f;
__device__ __forceinline__ void fast_sincos(float x, float& s, float& c) {
    float kf = fmaf(x, 0.636619772f, 12582912.0f);          // round(x * 2/pi) in the low mantissa bits
    const int q = __float_as_int(kf);
    kf -= 12582912.0f;
    float r = fmaf(kf, -1.57079601e+00f, x);
    r = fmaf(kf, -3.13916473e-07f, r);
    r = fmaf(kf, -5.39030253e-15f, r);
    const float r2 = r * r;
    float sp = fmaf(r2, -1.95152959e-4f, 8.33216087e-3f);
    sp = fmaf(sp, r2, -1.66666546e-1f);
    const float sn = fmaf(sp * r2, r, r);
    float cp = fmaf(r2, 2.44331571e-5f, -1.38873163e-3f);
    cp = fmaf(cp, r2, 4.16666457e-2f);
    cp = fmaf(cp, r2, -0.5f);
    const float cs = fmaf(cp, r2, 1.0f);
    const bool odd = q & 1;
    const float a = odd ? cs : sn, b = odd ? sn : cs;
    s = __int_as_float(__float_as_int(a) ^ ((q & 2) << 30));
    c = __int_as_float(__float_as_int(b) ^ (((q + 1) & 2) << 30));
}
// n evaluations, interleaved by the compiler (the loop is branch-free); rare huge arguments fall
// back to libdevice for the whole group.
__device__ __noinline__ float2 slow_sincos(float x) {      // out of line: keeps the hot loop small
    float2 r;
    sincosf(x, &r.x, &r.y);
    return r;
}
template <int N>
__device__ __forceinline__ void sincos_group(const float (&x)[N], float (&s)[N], float (&c)[N]) {
    float big = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) { fast_sincos(x[i], s[i], c[i]); big = fmaxf(big, fabsf(x[i])); }
    if (big > kSincosFastLimit) {
#pragma unroll
        for (int i = 0; i < N; ++i) { const float2 r = slow_sincos(x[i]); s[i] = r.x; c[i] = r.y; }   // static indices only
    }
}

// Tensor-path variant for values that are rounded to bf16 right away: the same exact-in-fp32
// Cody-Waite reduction (here by 2*pi, so no quadrant fix-up is needed), then the SFU evaluates
// sin/cos of the reduced argument in [-pi, pi], where MUFU.SIN/COS have their documented
// 2^-21.4 absolute error.  Total |error| <= 5e-7 (tests/test_gpu_tc.py::test_sincos_accuracy),
// four orders of magnitude below the bf16 rounding (2^-9 relative) applied to the result; what
// rules out bare __sinf at omega_0 = 60 is its range reduction, which this does not use.
// 8 issue slots per sin/cos pair instead of 24: the sine epilogue is FP32-issue-bound.
__device__ __forceinline__ void mufu_sincos(float x, float& s, float& c) {
    float kf = fmaf(x, 0.159154943f, 12582912.0f);          // round(x / 2pi)
    kf -= 12582912.0f;
    float r = fmaf(kf, -6.28125f, x);                       // 2pi = 6.28125 (8 bits: k*c1 exact) + 1.93530717e-3
    r = fmaf(kf, -1.93530717e-3f, r);
    s = __sinf(r);
    c = __cosf(r);
}
template <int N>
__device__ __forceinline__ void sincos_group_mufu(const float (&x)[N], float (&s)[N], float (&c)[N]) {
    float big = 0.f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        mufu_sincos(x[i], s[i], c[i]);
#ifndef NA_EXP_NOBIG                       // sensitivity experiments (profiles/README.md): never defined in a product build
        big = fmaxf(big, fabsf(x[i]));
#endif
#ifdef NA_EXP_HALFMUFU
        c[i] = s[i];
#endif
    }
    if (big > kSincosFastLimit) {
#pragma unroll
        for (int i = 0; i < N; ++i) { const float2 r = slow_sincos(x[i]); s[i] = r.x; c[i] = r.y; }
    }
}

template <typename T> __host__ __device__ inline T ceil_div(T a, T b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace; with base == nullptr it only sizes.
struct Arena {
    char* base;
    size_t off = 0;
    explicit Arena(void* b) : base(static_cast<char*>(b)) {}
    template <typename T> T* take(size_t count) {
        off = align_up(off, 256);
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
    size_t bytes() const { return align_up(off, 256); }
};

}  // namespace na
