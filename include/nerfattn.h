/*
 * nerfattn.h -- C ABI of libnerfattn.so, the B200 (sm_100a) replacement for the
 * SIREN fit / reconstruction hot path of ruskaruma/nerf-attention.
 *
 * The reference has no FFI: its seam is the Python API.  Every
 * entry point below therefore names the reference *Python* code it replaces
 * (paths relative to the reference tree).  The Python drop-in
 * (nerf-attention_b200/nerf_attention/_native.py) binds exactly these symbols
 * with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain C types only; every pointer inside na_fit_t is a DEVICE pointer
 *     owned by the caller; arrays of na_fit_t themselves live on the HOST
 *   - all work is enqueued on the caller's stream and is asynchronous with
 *     respect to the host; the library never synchronises the device
 *   - return value 0 = success, negative = NA_ERR_*; nerfattn_last_error()
 *     returns a thread-local message for the last failing call
 *   - no global mutable state besides cached kernel attributes and the
 *     thread-local error string: safe from several host threads on distinct
 *     streams and distinct workspaces
 */
#ifndef NERFATTN_H_
#define NERFATTN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NERFATTN_ABI_VERSION 8

#if defined(__GNUC__)
#define NA_API __attribute__((visibility("default")))
#else
#define NA_API
#endif

typedef struct CUstream_st* na_stream_t;      /* == cudaStream_t */

enum {
    NA_OK = 0,
    NA_ERR_INVALID = -1,       /* bad argument (null pointer, non-positive size, ...)   */
    NA_ERR_UNSUPPORTED = -2,   /* shape / precision outside what the kernels implement */
    NA_ERR_WORKSPACE = -3,     /* workspace missing or too small                         */
    NA_ERR_CUDA = -4           /* a CUDA runtime / driver call failed                    */
};

/*
 * Two precisions.  BASELINE.json's north_star allows "TF32/BF16" for the tensor mode; only BF16 is built:
 * with fp32 accumulation, an fp32 layer 0 and fp32 master weights it stays within 1.6e-4 of the reference's
 * final CosSim at the benched configuration (gate 5e-3; tests/test_gpu_full_length.py), at twice the tensor
 * rate and half the operand bytes a kind::tf32 variant would have.  Code 1 is unassigned (NA_ERR_UNSUPPORTED).
 */
enum {
    NA_PREC_FP32 = 0,          /* SIMT fp32 FMA everywhere: the parity mode (1e-5 / 1e-3) */
    NA_PREC_BF16 = 2           /* tcgen05 BF16 x BF16 -> FP32 for the H->H / H->D layers;
                                  layer 0, loss, Adam and master weights stay fp32        */
};

enum {
    NA_FIT_TARGETS_PRENORMALISED = 1   /* targets already (t-mean)/std; mean/std are inputs (device, [D]); the
                                          final metrics are computed against targets * std + mean */
};

/*
 * One fit = one (layer, head, key|value, architecture) job of the sweep.
 * Replaces the per-call state of reference fit_siren (nerf_attention/siren.py:80-93).
 *
 * params layout (fp32, nn.Linear [out,in] row-major, i.e. state_dict order,
 * nerf_attention/siren.py:43-58):
 *     W0[H,1] b0[H] { Wl[H,H] bl[H] } x L  Wf[D,H] bf[D]
 * nerfattn_param_count() gives the total P.
 */
typedef struct na_fit {
    int32_t N;                 /* seq_len, rows of the KV tensor                 */
    int32_t D;                 /* head_dim, SIREN out_features                   */
    int32_t H;                 /* SIRENConfig.hidden_features                    */
    int32_t L;                 /* SIRENConfig.hidden_layers (H->H sine layers)   */
    float   omega0;            /* SIRENConfig.omega_0                            */
    int32_t flags;             /* NA_FIT_*                                        */
    const float* positions;    /* [N]   as produced by torch.linspace(0,1,N)     */
    const float* targets;      /* [N,D] raw KV tensor, row-major fp32            */
    float* mean;               /* [D]   out: per-dim mean            (siren.py:85) */
    float* std;                /* [D]   out: unbiased std, >= 1e-3   (siren.py:86) */
    float* params;             /* [P]   in: initial weights, out: trained weights */
    float* adam_m;             /* [P]   in/out, caller zero-fills before 1st call */
    float* adam_v;             /* [P]   in/out, caller zero-fills before 1st call */
    float* losses;             /* [epochs] out: normalised MSE per epoch (siren.py:105) */
    float* cos_sims;           /* [N]   out: per-position CosSim    (siren.py:124) */
    float* per_pos_mse;        /* [N]   out: per-position MSE       (siren.py:125) */
    float* scalars;            /* [8]   out: final_mse, cos_mean, cos_min, cos_std(unbiased),
                                          then 4 reserved                        */
} na_fit_t;

NA_API int         nerfattn_abi_version(void);
NA_API const char* nerfattn_last_error(void);

/* P = 2H + L(H*H+H) + (H*D + D); replaces SIREN.count_parameters (siren.py:63-64). */
NA_API size_t nerfattn_param_count(int32_t H, int32_t L, int32_t D);

/* Scratch bytes nerfattn_fit_batched needs for this job list and precision. */
NA_API int nerfattn_fit_workspace_bytes(const na_fit_t* fits, int32_t nfits, int32_t precision,
                                 size_t* bytes);

/*
 * Train every fit of the list for `epochs` full-batch Adam steps and write the
 * final metrics.  Replaces the serial loop nest of fit_kv_cache
 * (nerf_attention/fit.py:54-76) around fit_siren's normalisation, epoch loop
 * and final evaluation (nerf_attention/siren.py:85-87, 98-105, 119-125), and
 * torch.optim.Adam + CosineAnnealingLR underneath them.
 *
 *   lr_table   HOST array [epochs], float64: the learning rate optimizer.step()
 *              uses at each epoch (the caller builds it from the real scheduler)
 *   beta1/2, eps  Adam hyper-parameters (torch defaults 0.9, 0.999, 1e-8)
 *   first_step 0-based count of Adam steps already applied to adam_m/adam_v
 *              (0 for a fresh fit); step numbers continue from there
 *   workspace  DEVICE scratch of >= nerfattn_fit_workspace_bytes() bytes,
 *              256-byte aligned; contents undefined afterwards
 */
NA_API int nerfattn_fit_batched(const na_fit_t* fits, int32_t nfits, int32_t epochs,
                         const double* lr_table, double beta1, double beta2, double eps,
                         int32_t first_step, int32_t precision,
                         void* workspace, size_t workspace_bytes, na_stream_t stream);

/*
 * Same, plus the reference's progress metrics (nerf_attention/siren.py:107-115): whenever
 * (epoch + 1) % log_every == 0 the de-normalised MSE and the mean CosSim of the prediction made
 * with the weights that epoch starts from are written to
 *   progress[((epoch + 1) / log_every - 1) * nfits * 2 + fit * 2 + {0: RealMSE, 1: CosSim}]
 * (DEVICE float array of (epochs / log_every) * nfits * 2 values; fp32 evaluation in every mode).
 * log_every <= 0 or progress == NULL: no progress metrics (== nerfattn_fit_batched).
 */
NA_API int nerfattn_fit_batched_ex(const na_fit_t* fits, int32_t nfits, int32_t epochs,
                            const double* lr_table, double beta1, double beta2, double eps,
                            int32_t first_step, int32_t precision, int32_t log_every, float* progress,
                            void* workspace, size_t workspace_bytes, na_stream_t stream);

/* Number of kernel launches nerfattn_fit_batched enqueues for this job list (graph replays
 * counted once per epoch): set-up + epochs * per-epoch + final metrics. */
NA_API long long nerfattn_fit_launch_count(const na_fit_t* fits, int32_t nfits, int32_t epochs,
                                           int32_t precision);

/*
 * Full-sequence reconstruction out[i][N,D] = SIREN_i(positions) (optionally
 * * std + mean).  Replaces model(positions) in profile_latency
 * (nerf_attention/evaluate.py:196-200) and the reconstruction of
 * plot_per_position_error (evaluate.py:148-152).  fp32 SIMT, no scratch.
 * Only N, D, H, L, omega0, positions, params (and mean/std when denormalise
 * != 0) of each na_fit_t are read.  `out` is a HOST array of n DEVICE pointers.
 */
NA_API int nerfattn_siren_forward(const na_fit_t* models, int32_t n, int32_t denormalise,
                           float* const* out, void* workspace, size_t workspace_bytes,
                           na_stream_t stream);
NA_API int nerfattn_forward_workspace_bytes(const na_fit_t* models, int32_t n, size_t* bytes);

/*
 * Decode-step attention logits from the SIREN instead of the KV cache:
 *     scores[i][t] = q[i] . (SIREN_i(pos_t) * std_i + mean_i),  t < N
 * with the output layer folded into the query (DESIGN.md "decode"), so K is never
 * materialised.  New functionality: the reference only times the unfused forward
 * (evaluate.py:196-200) against a theoretical HBM read (evaluate.py:210-213).
 *   q_fp16       DEVICE fp16 [n, D], the new token's query per head
 *   scores       HOST array of n DEVICE pointers to fp32 [N]
 *   precision    NA_PREC_FP32 (SIMT) or NA_PREC_BF16 (tcgen05 hidden layers)
 *   reuse_setup  0: upload model table / bf16 weight mirror into the workspace, then run;
 *                1: the workspace still holds the set-up of an earlier call with the same
 *                   models and the same score pointers (only q changed): run only -- this is
 *                   the per-token cost; `scores` is not re-read
 * All n models must share (N, D, H, L).  Reads N, D, H, L, omega0, positions, params,
 * mean, std of each na_fit_t.
 */
NA_API int nerfattn_decode_workspace_bytes(const na_fit_t* key_models, int32_t n, int32_t precision,
                                    size_t* bytes);
NA_API int nerfattn_decode_qk(const na_fit_t* key_models, int32_t n, const void* q_fp16,
                       float* const* scores, int32_t precision, int32_t reuse_setup,
                       void* workspace, size_t workspace_bytes, na_stream_t stream);

/*
 * The baseline the SIREN decode is compared with: stream fp16 keys K[n,N,D]
 * from HBM and compute scores[i*N+t] = q[i] . K[i][t] (fp32 accumulate).
 * Replaces the constants raw_bytes/272e9 and raw_bytes/3350e9 of
 * evaluate.py:210-211 with a measured, bandwidth-saturating read.
 */
NA_API int nerfattn_kvread_qk(const void* k_fp16, const void* q_fp16, float* scores,
                       int32_t n, int32_t N, int32_t D, na_stream_t stream);

/*
 * Single-query attention over the cached positions (the decode integration the reference describes
 * in README.md:3-8 and never builds):
 *   scores = nerfattn_decode_qk / nerfattn_kvread_qk,  p = softmax(scale * scores),  out = sum_t p_t V_t.
 *
 * nerfattn_softmax     p[i][:] = softmax(scale * scores[i][:]) in place, scores DEVICE fp32 [n][N]
 * nerfattn_kvread_pv   baseline: out[i] = sum_t p[i][t] V[i][t][:], V DEVICE fp16 [n][N][D] streamed from HBM
 * nerfattn_decode_pv   the same with V_i(t) = SIREN_i(position t) * std_i + mean_i (value models); in the
 *                      BF16 mode V is never materialised: the fused forward kernel reduces p_t * h_L(t)
 *                      over the positions and the output layer is applied once to the H-vector
 * out: DEVICE fp32 [n][D].  Workspaces: nerfattn_pv_workspace_bytes (models == NULL: the kvread variant).
 */
NA_API int nerfattn_softmax(float* scores, int32_t n, int32_t N, float scale, na_stream_t stream);
NA_API int nerfattn_pv_workspace_bytes(const na_fit_t* value_models, int32_t n, int32_t N, int32_t D,
                                int32_t precision, size_t* bytes);
NA_API int nerfattn_kvread_pv(const void* v_fp16, const float* p, float* out, int32_t n, int32_t N, int32_t D,
                       void* workspace, size_t workspace_bytes, na_stream_t stream);
NA_API int nerfattn_decode_pv(const na_fit_t* value_models, int32_t n, const float* p, float* out,
                       int32_t precision, void* workspace, size_t workspace_bytes, na_stream_t stream);

/*
 * Synthetic KV-cache generator.
 * Replaces the per-(layer, head) body of extract_kv_cache_synthetic, nerf_attention/extract.py:199-237:
 * one numpy legacy RandomState(seed) stream per entry (seed = layer * num_kv_heads + head, :207), consumed in
 * the reference's order; n_spikes = int(3 * sharpness) and max_width = max(2, int(5 / sharpness)) with
 * sharpness = 1 + 2 * layer / max(num_layers - 1, 1) (:204,:222-224) are computed by the caller.
 * positions: DEVICE fp32 [N] = torch.linspace(0, 1, N) (:197).  keys / values: DEVICE fp32 [N][D] out
 * (one head of layer_XX.pt's [kv_heads, N, D] tensors).  `streams` itself is a HOST array.
 * Results agree with the CPU generator to float32 rounding of the smooth terms (libm / numpy sin, cos, log and
 * exp differ from CUDA's in the last place); the random stream itself is reproduced exactly.
 */
typedef struct na_synth_stream {
    uint32_t seed;
    int32_t n_spikes, max_width;
    float* keys; float* values;
} na_synth_stream_t;
NA_API int nerfattn_synth_workspace_bytes(int32_t nstreams, int32_t N, int32_t D, size_t* bytes);
NA_API int nerfattn_synth_kv(const na_synth_stream_t* streams, int32_t nstreams, int32_t N, int32_t D,
                      const float* positions, void* workspace, size_t workspace_bytes, na_stream_t stream);

/*
 * Diagnostic: C[M,N] (fp32) = A x B with BF16 operands through the same
 * tcgen05/TMA tile pipeline the BF16 fit uses.  a_mn_major / b_mn_major select
 * the operand storage:  A is [M,K] row-major (0) or [K,M] row-major (1);
 * B is [N,K] row-major (0) or [K,N] row-major (1).  Used by the tests to pin
 * the shared-memory descriptors against a plain matmul.
 */
NA_API int nerfattn_debug_gemm_bf16(const void* a_bf16, const void* b_bf16, float* c,
                             int32_t M, int32_t N, int32_t K, int32_t batch,
                             int32_t a_mn_major, int32_t b_mn_major, na_stream_t stream);

/*
 * Diagnostic: s[i], c[i] = sin, cos of x[i] (device fp32 arrays of n values) through the
 * device functions the epilogues use.  mode 0: branch-free polynomial (fp32 path, layer 0),
 * mode 1: exact Cody-Waite reduction + SFU core (hidden layers of the BF16 path, whose results
 * are rounded to bf16).  Replaces torch.sin (siren.py:34); the tests bound both against float64.
 */
NA_API int nerfattn_debug_sincos(const float* x, float* s, float* c, int64_t n, int32_t mode,
                          na_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* NERFATTN_H_ */
