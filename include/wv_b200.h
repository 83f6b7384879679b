/*
 * wv_b200.h - C ABI of the B200 (sm_100a) WaveVerify embed / detect / locate hot path.
 *
 * The reference (pujariaditya/WaveVerify) is pure Python/PyTorch: it has no FFI of its own.
 * The entry points below are what a binding for THIS path has to expose so that the
 * reference's Python call sites can be served by hand-written CUDA:
 *
 *   wv_generator_forward  <- model/generator.py:360-423  Generator.forward  (+ the add at
 *                            model/watermarking.py:440)
 *   wv_generator_encode / wv_generator_decode
 *                         <- model/generator.py:290-358  Generator.encode / .decode
 *   wv_detector_forward   <- model/detector.py:366-391   Detector.forward, fused with the bit
 *                            decode of waveverify/core.py:577-586 + waveverify/utils.py:385-401
 *                            and the masked variant of scripts/evaluate.py:471-494
 *   wv_locator_forward    <- model/locator.py:268-299    Locator.forward, fused with the mask
 *                            threshold of model/watermarking.py:717,797 and the sigmoid of
 *                            waveverify/core.py:632
 *   wv_metrics_accumulate <- scripts/evaluate.py:498-505 (BER) and :636-656 (MIoU) as six exact
 *                            integer counters (the only quantity all-reduced across GPUs)
 *
 * Conventions: plain pointers and sizes only, no torch types.  Every function returns 0 on
 * success or a negative code; wv_last_error() gives the message (thread local).  All data
 * pointers passed to *_forward are DEVICE pointers owned by the caller; nullable outputs are
 * skipped.  Work is enqueued on `stream` (a cudaStream_t passed as void*), no implicit sync.
 * A wv_net owns its weights, TMA descriptors and workspace; it is bound to one device and is
 * not safe for concurrent use from several threads.  There is no CPU fallback.
 */
#ifndef WV_B200_H
#define WV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WV_KIND_GENERATOR 0
#define WV_KIND_DETECTOR 1
#define WV_KIND_LOCATOR 2

#define WV_OK 0
#define WV_ERR_INVALID -1
#define WV_ERR_CUDA -2
#define WV_ERR_MISSING_WEIGHT -3
#define WV_ERR_UNSUPPORTED -4

typedef struct wv_net wv_net;

/* Topology: the supported subset of the reference constructor kwargs (conf/base.yml:5-112). */
typedef struct wv_net_config {
  int kind;            /* WV_KIND_* */
  int sample_rate;
  int dimension;       /* latent channels */
  int channels_enc;    /* n_filters of the encoder */
  int channels_dec;    /* n_filters of the decoder (generator only) */
  int n_fft_base;
  int n_residual_enc;
  int n_residual_dec;
  int n_strides;
  int strides[8];      /* as given to the reference (the encoder walks them reversed) */
  float res_scale_enc;
  float res_scale_dec;
  int nbits;           /* detector: 16, locator: 1 */
  int output_dim;      /* head hidden width (32) */
  int msg_dimension;
  int embedding_dim;
  int embedding_layers;
  int freq_bands;
  int precise;         /* 1: fp32-accurate arithmetic (split-fp16 tensor-core operands, fp32 epilogues) for the nets whose
                          outputs are thresholded: the locator mask (model/watermarking.py:717) and the detector re-check of
                          near-threshold bits (waveverify/core.py:577-586).  Detector / locator only. */
} wv_net_config;

/* One folded fp32 host tensor, keyed by the reference state_dict name with the weight-norm /
 * weight-standardisation parametrisation collapsed to `<prefix>.weight`. */
typedef struct wv_tensor {
  const char* name;
  const float* data;   /* host pointer, contiguous */
  int ndim;
  int64_t shape[4];
} wv_tensor;

int wv_version(void);
const char* wv_last_error(void);

int wv_net_create(const wv_net_config* cfg, const wv_tensor* tensors, int n_tensors, int device,
                  wv_net** out);
int wv_net_destroy(wv_net* net);
/* Build (or fetch) the launch plan for [B, 1, T] inputs and size the workspace. */
int wv_net_reserve(wv_net* net, int B, int T);
size_t wv_net_workspace_bytes(const wv_net* net);
/* Number of kernel launches one forward at the reserved shape enqueues. */
int wv_net_launches(const wv_net* net, int B, int T);
/* Max clips processed per internal sub-batch (0 = whole batch). */
int wv_net_set_chunk(wv_net* net, int max_clip_samples);

/* x [B,T] fp32, msg [B,msg_dimension] fp32 (0/1).  wm_out [B,T] = watermark residual,
 * y_out [B,T] = x + wm, latent_out [B,dimension,ceil(T/hop)] fp32.  Any output may be NULL. */
int wv_generator_forward(wv_net* net, const float* x, const float* msg, int B, int T,
                         float* wm_out, float* y_out, float* latent_out, void* stream);
int wv_generator_encode(wv_net* net, const float* x, const float* msg, int B, int T,
                        float* latent_out, void* stream);
/* z [B,dimension,F] fp32 -> wav_out [B, F*hop] fp32 */
int wv_generator_decode(wv_net* net, const float* z, int B, int F, float* wav_out, void* stream);

/* y [B,T] fp32.  logits [B,nbits,T] fp32 raw; bits [B,nbits] u8 = (avg >= 0.5);
 * avg [B,nbits] fp32 = (masked) time-mean of sigmoid(logit); conf [B] = mean over bits;
 * valid [B,nbits] u8 = bit has >=1 unmasked sample; presence [B,T] u8 mask or NULL. */
int wv_detector_forward(wv_net* net, const float* y, int B, int T, float* logits, uint8_t* bits,
                        float* avg, float* conf, uint8_t* valid, const uint8_t* presence,
                        void* stream);
/* Exact bit decisions without a host round trip (waveverify/core.py:577-586, scripts/evaluate.py:471-494): `net` is the
 * PRECISE detector net; bits / avg / conf / valid (/ logits) hold the fp16 fast path's results for y and are overwritten
 * for every clip that has a (valid) bit with |avg - 0.5| < tau (tau_short for clips / masks shorter than short_samples).
 * Selection, the re-evaluation of the selected clips `slots` at a time and the write-back run inside one CUDA graph with a
 * device-side WHILE node; counters (device int[2], nullable) += {clips re-evaluated, passes}.  Nothing runs when no clip
 * is near the threshold. */
int wv_detector_refine(wv_net* net, const float* y, int B, int T, float* logits, uint8_t* bits, float* avg,
                       float* conf, uint8_t* valid, const uint8_t* presence, float tau, float tau_short,
                       int short_samples, int slots, int* counters, void* stream);
/* y [B,T] fp32.  logits [B,T] fp32 raw, mask [B,T] u8 = (logit > 0.5), probs = sigmoid(logit). */
int wv_locator_forward(wv_net* net, const float* y, int B, int T, float* logits, uint8_t* mask,
                       float* probs, void* stream);

/* counters[6] (device int64): += {bit_errors, valid_bits, I_fg, U_fg, I_bg, U_bg}. */
int wv_metrics_accumulate(const uint8_t* bits, const uint8_t* valid, const uint8_t* msg_bits,
                          int B, int nbits, const uint8_t* pred_mask, const uint8_t* gt_mask,
                          long long n_mask, long long* counters, void* stream);

/* ---- validation path: device pointers, fp32 [B, T] ------------------
 * wv_augment_gather replaces LocalizationAugmentation.forward (utils/localization_augmentation.py:
 * 212-325) and SequenceAugmentation.forward (utils/seq_augmentation.py:100-277) with one gather pass.
 *   seg_op [B, n_seg] u8 (NULL = no localization step): 0 keep, 1 revert to original, 2 zeros,
 *   3 original of clip seg_src[b, s]; segment s covers samples [s*seg_len, (s+1)*seg_len).
 *   seq_kind: 0 identity, 1 reverse, 2 circular shift by seq_a, 3 shuffle of n_perm segments of
 *   seq_c samples (seq_perm device int[n_perm]; T_out = n_perm*seq_c), 4 swap chunks [seq_a,+seq_c)
 *   and [seq_b,+seq_c).  T_out = T except for kind 3.
 *   gt_in (nullable): presence of the input (all ones when NULL).  Each out_* is nullable. */
int wv_augment_gather(const float* original, const float* watermarked, const float* gt_in, int B, int T,
                      const uint8_t* seg_op, const int* seg_src, int seg_len, int n_seg, int seq_kind,
                      int seq_a, int seq_b, int seq_c, const int* seq_perm, int n_perm, int T_out,
                      float* out_wm, float* out_orig, float* out_gt, void* stream);
/* utils/effect_augmentation.py pointwise effects over n samples: 0 identity (:1364), 1 amplitude_scaling
 * (p0 = scale, :2000), 2 quantization (p0 = 2^(bit_depth-1) - 1, :1090-1111), 3 additive noise
 * x + noise*p0 with the caller's N(0,1) draw (:2105, :2338), 4 the same with an in-kernel Philox draw. */
int wv_effect_pointwise(int effect, const float* in, long long n, float p0, const float* noise,
                        unsigned long long seed, float* out, void* stream);
/* sample_suppression (:2061-2103): audio[b, idx[b, j]] = 0 and mask[b, idx[b, j]] = 0 (mask nullable), in place. */
int wv_effect_suppress(float* audio, float* mask, const long long* idx, int B, int T, int k, void* stream);
/* median_filter (:1246-1312, scipy.signal.medfilt: zero-padded ends), odd k <= 31. */
int wv_effect_median(const float* in, int B, int T, int k, float* out, void* stream);

/* julius low / high / band-pass as utils/effect_augmentation.py:1684-1871 calls them: odd-length FIR (taps on the
 * device, designed by the host), replicate padding, out = subtract ? in - fir(in) : fir(in). */
int wv_effect_fir(const float* in, const float* taps, int n_taps, int B, int T, int subtract, float* out, void* stream);

/* torchaudio.transforms.Resample(orig, new) as the `resample` effect uses it (utils/effect_augmentation.py:1451-1502) and
 * the resampling half of `speed` (:1381-1449).  orig / nw = the two rates divided by their gcd; taps [nw][2*width + orig]
 * (device, designed by the host); T_mid = ceil(T*nw/orig) = resampled length.  lerp = 1 additionally stretches the
 * resampled signal to T_out samples by linear interpolation (align_corners = False), as `speed` does (:187-215). */
int wv_effect_resample(const float* in, const float* taps, int B, int T, int orig, int nw, int width, int T_mid, int T_out,
                       int lerp, float* out, void* stream);

/* ---- fp16 range check -------------------------------------------------------------------------
 * The fast nets store activations as fp16: every fp32 -> fp16 conversion saturates at +-65504 and the packed-half2 epilogue
 * arithmetic can overflow to inf, both silently.  With the check enabled every launch's fp16 outputs are scanned by an extra
 * kernel (slow: run it once after loading a checkpoint, on representative full-scale audio).  wv_net_range_read synchronises,
 * returns the counts accumulated since the last read (values AT the saturation bound, non-finite values, largest finite |v|)
 * and resets them. */
int wv_net_set_range_check(wv_net* net, int enable);
int wv_net_range_read(wv_net* net, unsigned long long* saturated, unsigned long long* nonfinite, float* max_abs);

/* ---- profiling / debugging (used by bench.py and the tests) ------------------------------- */
/* When enabled, every launch of a forward is bracketed by CUDA events on the caller's stream. */
int wv_net_set_profile(wv_net* net, int enable);
/* Per-launch device time (ms), algorithmic FLOPs and bytes, and kernel class of the last profiled
 * forward (one sub-batch).  Returns the number of launches written, or a negative code. */
int wv_net_profile_read(wv_net* net, int max_ops, float* ms, double* flops, double* bytes, int* cls);
const char* wv_net_profile_tag(wv_net* net, int i);
/* Run up to the launch tagged `tag` and copy its output (0 = raw, 1 = activated) to dst. */
int wv_debug_tap(wv_net* net, const float* x, const float* msg, int B, int T, const char* tag,
                 int which, void* dst, size_t dst_bytes, size_t* written);

/* ---- single-kernel entry points (unit tests / micro-benchmarks; device pointers) -------- */
/* out = epilogue(A[M,K] * W[N,K]^T): fp16 in, fp32 accumulate on tcgen05; see DESIGN.md. */
int wv_op_gemm(const void* A, int lda, const void* W, int ldw, int M, int N, int K,
               const float* bias, const void* residual, void* out_raw, void* out_act,
               float act_scale, int a_is_fp16, void* stream);
/* A [B,T,K] fp16 -> 1x1 conv (W [N,K]) -> causal depthwise k=5 (dw_w5n [5][N], bias) fused in the
 * GEMM epilogue (+ residual [B,T,N]) -> out_raw / out_act = ELU(v*act_scale), each [B,T,N] fp16. */
int wv_op_gemm_dw5(const void* A, const void* W, int B, int T, int N, int K, const float* dw_w5n,
                   const float* bias, const void* residual, void* out_raw, void* out_act,
                   float act_scale, void* stream);
/* One fused SEANet residual block (modules/seanet.py:245-281), C <= 128, C % 32 == 0:
 * x' = dw5(W2 ELU(dw5(W1 A) + b1)) + b2 + X ; out_raw = x', out_act = ELU(x'*act_scale), with A = ELU(X*pre_scale) as the
 * producing launch stores it.  X, A, out_* are [B,T,C] fp16; W1, W2 [C,C] fp16; taps [5][C], biases [C] fp32 (RS folded
 * into dw2). */
int wv_op_resblock(const void* X, const void* A, const void* W1, const float* dw1_w5c, const float* dw1_b,
                   const void* W2, const float* dw2_w5c, const float* dw2_b, int B, int T, int C,
                   float pre_scale, void* out_raw, void* out_act, float act_scale, void* stream);
int wv_op_dw5(const void* in, const float* w5c, const float* bias, const void* residual,
              void* out_raw, void* out_act, float act_scale, int B, int T, int C, void* stream);
int wv_op_down(const void* in, const float* wkc, const float* bias, const float* film,
               int film_stride, int bands, void* out_raw, void* out_act, float act_scale, int B,
               int Tin, int C, int r, void* stream);
int wv_op_up(const void* in, const float* wkc, void* out, int B, int Tin, int C, int r,
             void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WV_B200_H */
