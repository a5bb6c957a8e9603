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
    int uniq;               // index of the unique target tensor
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
