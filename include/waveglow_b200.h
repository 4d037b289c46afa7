/* waveglow_b200.h -- C ABI of the B200-native WaveGlow flow engine (libwaveglow_b200.so).
 *
 * The reference (stefantaubert/waveglow) has no FFI layer: its hot path is the Python method
 * `WaveGlow.infer(spect, sigma)` (src/waveglow/model.py:223-274) and, for the training direction,
 * `WaveGlow.forward((spect, audio))` (model.py:178-221).  This header is the boundary a
 * maintainer binds (ctypes) to replace the body of those two methods.
 *
 * Conventions
 *  - every function returns 0 on success or a negative wg_status; nothing throws across the ABI;
 *    wg_last_error() returns a human readable message for the last failure on that engine;
 *  - all tensor memory is owned by the caller (the torch allocator); the engine owns only its
 *    packed weight copies;
 *  - wg_infer / wg_forward are stream-ordered and asynchronous on `stream` (a cudaStream_t);
 *  - one engine per device; an engine is not re-entrant;
 *  - there is NO CPU fallback: without a CUDA device of compute capability 10.x wg_create fails.
 */
#ifndef WAVEGLOW_B200_H_
#define WAVEGLOW_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wg_engine wg_engine;

typedef enum wg_status {
  WG_OK = 0,
  WG_ERR_INVALID_ARGUMENT = -1,
  WG_ERR_UNSUPPORTED = -2,   /* hparams / device the kernels are not built for */
  WG_ERR_CUDA = -3,          /* a CUDA runtime/driver call failed (see wg_last_error) */
  WG_ERR_WORKSPACE = -4,     /* workspace too small / misaligned */
  WG_ERR_WEIGHTS = -5        /* missing / mis-shaped weight, or weights not loaded */
} wg_status;

/* Arithmetic mode of the WN contractions.
 *  WG_PRECISION_FP32: CUDA-core fp32 FMA everywhere (correctness mode, max-abs <= 1e-5 vs reference).
 *  WG_PRECISION_FP16: tcgen05 kind::f16 tensor cores, fp16 operands, fp32 accumulation in TMEM
 *                     (throughput mode; SNR >= 40 dB, max-abs <= 2e-3 vs reference). */
typedef enum wg_precision {
  WG_PRECISION_FP32 = 0,
  WG_PRECISION_FP16 = 1
} wg_precision;

/* Mirrors ModelHParams, src/waveglow/hparams.py:20-31 (same names, same defaults). */
typedef struct wg_config {
  int32_t n_mel_channels;   /* 80   */
  int32_t n_flows;          /* 12   */
  int32_t n_group;          /* 8    */
  int32_t n_early_every;    /* 4    */
  int32_t n_early_size;     /* 2    */
  int32_t n_layers;         /* 8    */
  int32_t n_channels;       /* 256  */
  int32_t kernel_size;      /* 3    */
  int32_t segment_length;   /* 16000 (unused by the flow itself) */
  int32_t precision;        /* wg_precision */
  int32_t device;           /* CUDA device ordinal */
} wg_config;

/* One FOLDED (weight-norm removed, model.py:276-297) fp32 tensor in the reference's layout,
 * named by its state_dict key after remove_weightnorm, e.g. "WN.3.in_layers.2.weight" (512,256,3),
 * "WN.3.cond_layer.bias" (4096), "convinv.7.conv.weight" (6,6,1), "upsample.weight" (80,80,1024).
 * `data` may be a host or a device pointer (UVA); the engine copies what it needs. */
typedef struct wg_weight_desc {
  const char* name;
  const float* data;
  int32_t ndim;
  int64_t shape[4];
} wg_weight_desc;

/* Replaces: WaveGlow.__init__ (model.py:142-176).  Validates hparams, selects the device. */
int wg_create(const wg_config* cfg, wg_engine** out);

/* Replaces: load_state_dict + WaveGlow.remove_weightnorm + the lazily cached W_inverse
 * (train.py:48-55, model.py:276-297, model.py:49-58).  Packs the engine's private copies
 * (K-major fp16 / fp32 layouts, folded biases, skip->end fold, inverse 1x1 matrices, log-dets).
 * May be called again after the caller's weights changed. */
int wg_load_weights(wg_engine* e, const wg_weight_desc* descs, int32_t n_descs);

/* Packed-weight cache (the reference rebuilds and re-folds the whole module tree on every start,
 * synthesizer.py:32-52).  wg_save_packed writes the engine's packed arrays (K-major 16-bit weights, folded biases, the
 * skip->end fold, composed start weights, 1x1 inverses and log-dets, and the fp32 conditioning weights from which the
 * phase-composed V is rebuilt on the device) to `path`; wg_load_packed restores them instead of wg_load_weights.
 * `fingerprint` identifies the source weights (e.g. checkpoint size and mtime); a cache written for other weights, another
 * wg_config, another library version or another kernel selection is rejected with WG_ERR_WEIGHTS and the caller packs
 * from the state_dict as usual. */
int wg_save_packed(wg_engine* e, const char* path, uint64_t fingerprint);
int wg_load_packed(wg_engine* e, const char* path, uint64_t fingerprint);

/* Bytes of scratch wg_infer / wg_forward need for a batch of `batch` mels of `frames` frames. */
size_t wg_workspace_bytes(const wg_engine* e, int32_t batch, int32_t frames);

/* Replaces: the body of WaveGlow.infer (model.py:223-274) after the noise draw.
 *  mel      [batch, n_mel_channels, frames] fp32, device
 *  z        n_z device pointers to the pre-drawn N(0,1) tensors in the order the reference draws
 *           them (model.py:239-243, 266-270): z[0] = [batch, n_remaining_channels, T'], then one
 *           [batch, n_early_size, T'] per early output (k descending); T' = frames*256/n_group
 *  sigma    scales every z (model.py:244, 271); 0.0 is valid (Denoiser, denoiser.py:46)
 *  audio    [batch, frames*256] fp32, device, written
 *  workspace/ws_bytes  >= wg_workspace_bytes(e, batch, frames), 1024-byte aligned
 *  stream   cudaStream_t */
int wg_infer(wg_engine* e, const float* mel, int32_t batch, int32_t frames,
             const float* const* z, int32_t n_z, float sigma, float* audio,
             void* workspace, size_t ws_bytes, void* stream);

/* Replaces: WaveGlow.forward (model.py:178-221) + the reductions WaveGlowLoss needs (train.py:31-45).
 *  mel      [batch, n_mel_channels, frames] fp32;  audio_in [batch, n_samples] fp32,
 *           n_samples % n_group == 0 and n_samples <= frames*256 + 768 (front-aligned trim, model.py:186-189)
 *  z_out    [batch, n_group, T'] fp32 (reference layout, channel-major), T' = n_samples / n_group
 *  log_s_out   optional (may be NULL): n_flows device pointers, each [batch, n_half_k, T'] fp32
 *  log_s_sum   [n_flows] fp32 device: sum over all elements of log_s of flow k
 *  log_det_w   [n_flows] fp32 device: batch * T' * logdet(W_k)   (model.py:63)
 *  z_sq_sum    optional (may be NULL): [1] fp32 device: sum over all elements of z^2 (train.py:41), reduced by the kernel
 *              that writes z_out, so WaveGlowLoss needs no pass over z or log_s */
int wg_forward(wg_engine* e, const float* mel, int32_t batch, int32_t frames,
               const float* audio_in, int32_t n_samples, float* z_out,
               float* const* log_s_out, float* log_s_sum, float* log_det_w, float* z_sq_sum,
               void* workspace, size_t ws_bytes, void* stream);

/* Replaces: the per-call arithmetic of Denoiser.forward between its two STFTs (denoiser.py:51-57 with stft.py:155-165:
 * magnitude, atan2 phase, bias subtraction, clamp, cos / sin recombination) by the equivalent in-place real scaling
 *   X <- X * max(|X| - strength * bias[f], 0) / |X|       (0 where |X| == 0, like the reference)
 *  spec     interleaved (re, im) fp32, device: element (b, f, t) at complex index b*stride_b + f*stride_f + t*stride_t
 *           (strides in complex elements: the [B, n_freq, n_frames] view torch.stft returns is not contiguous)
 *  bias     [n_freq] fp32, device: the model's bias magnitude spectrum (denoiser.py:47-49)
 * Stateless (no engine handle): `device` is the CUDA ordinal the pointers live on; stream-ordered on `stream`. */
int wg_denoise_spectrum(float* spec, const float* bias, float strength, int32_t batch, int32_t n_freq, int32_t n_frames,
                        int64_t stride_b, int64_t stride_f, int64_t stride_t, int32_t device, void* stream);

/* Number of kernel launches the last wg_infer / wg_forward call enqueued (for bench.py's gpu_launches). */
int wg_last_launch_count(const wg_engine* e);

/* CUDA-event time (ms, on the call's stream) spent in the WN layer kernels of the last wg_infer /
 * wg_forward call and how many layer launches that covers; blocks until those kernels finished.
 * bench.py divides the kernels' algorithmic FLOPs by this for the live roofline figure. */
int wg_profile_layers(wg_engine* e, float* total_ms, int32_t* n_launches);

/* The per-flow timing events behind wg_profile_layers are recorded by default; switch them off (0) while the caller captures
 * wg_infer into a CUDA graph (cudaStreamBeginCapture on `stream`), on again (1) afterwards. */
int wg_set_profiling(wg_engine* e, int32_t enabled);

/* Diagnostic GEMM D[M,N] fp32 = A[M,K] . B[N,K]^T on 16-bit device operands (fmt 0 = fp16, 1 = bf16)
 * through the same TMA / tcgen05 building blocks as the layer kernel (tests only). */
int wg_debug_gemm(wg_engine* e, const void* a, const void* b, float* d, int32_t m, int32_t n, int32_t k,
                  int32_t fmt, void* stream);

/* Same through tcgen05.mma.cta_group::2 on CTA pairs (M % 256 == 0): validates the 2-CTA mechanisms in isolation.
 * ks = K elements per operand stage: 64 (128-byte rows, SWIZZLE_128B) or 32 (64-byte rows, SWIZZLE_64B) -- the two
 * operand-ring formats of the layer kernel. */
int wg_debug_gemm2(wg_engine* e, const void* a, const void* b, float* d, int32_t m, int32_t n, int32_t k,
                   int32_t fmt, int32_t ks, void* stream);

/* Test helper: n_ctas CTAs that each hold one SM (200 KB of shared memory, 64 threads) for `cycles` clock ticks on `stream`.
 * tests/ use it to take SMs away from a chained layer launch: claimed work items must still complete (no co-residency assumption). */
int wg_debug_occupy(wg_engine* e, int32_t n_ctas, int64_t cycles, void* stream);

/* Diagnostics: when set (device pointer to [n_ctas][16] int64), every fused layer launch writes clock64
 * stamps of its pipeline phases per CTA (slot meaning: tools/timeline.py).  NULL disables (default). */
int wg_debug_timeline(wg_engine* e, long long* device_buffer);

/* Test/diagnostic hooks (used by tests/ only).
 * wg_debug_stop_after: make the next wg_infer return after flow `flow` (counting down from
 * n_flows-1), WN layer `layer` (-1 = after the start conv, n_layers = after the whole flow incl. the
 * coupling/1x1 step).  flow = -1 disables.  The intermediates then sit in the workspace:
 * wg_debug_view returns the device pointer/byte size of a named buffer for (batch, frames):
 *   "cond"  [B,T',n_mel*n_group]   operand dtype     "x"   [B,T',n_channels] operand dtype (current stream)
 *   "e"     [B,T',8] fp32 skip-end accumulator        "state" = the audio buffer itself [B,T',n_group] fp32 */
int wg_debug_stop_after(wg_engine* e, int32_t flow, int32_t layer);
int wg_debug_view(const wg_engine* e, const char* name, int32_t batch, int32_t frames,
                  void* workspace, void** ptr, size_t* bytes);

const char* wg_last_error(const wg_engine* e);
const char* wg_version(void);
void wg_destroy(wg_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* WAVEGLOW_B200_H_ */
