// Validation-path kernels: the temporal augmentations and the cheap
// signal effects that sit between the Generator and the Detector / Locator in
// model/watermarking.py:443-525 (_forward_valid, _apply_augmentations) and :757-806
// (_evaluate_single_effect).  The reference runs them on the host (numpy / torch-CPU, a GPU -> CPU ->
// GPU bounce per effect); here every tensor stays in HBM.  All of them are memory-bound, one pass:
// fp32 [B, T] in, fp32 out, 4-byte or 16-byte coalesced accesses, grid-stride loops.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace wv {

// ---------------------------------------------------------------------------------------------
// LocalizationAugmentation.forward (utils/localization_augmentation.py:212-325) and
// SequenceAugmentation.forward (utils/seq_augmentation.py:100-277) as ONE gather pass.
//
// Localization: per clip, a table of segments (seg_len samples each); op 0 = unchanged, 1 = revert to
// the original (:126-149), 2 = zeros (:151-175), 3 = the original of clip seg_src (:177-210).
// Sequence: one index map for the whole batch: 0 identity, 1 reverse (torch.flip), 2 circular shift
// (torch.roll by a), 3 shuffle of n_perm segments of `c` samples (unfold + index + view: the tail
// past n_perm * c is dropped, so T_out = n_perm * c), 4 swap of the chunks [a, a+c) and [b, b+c).
// The sequence map is applied to the OUTPUT of the localization step, i.e. the gather composes
// t_out -> t_src -> segment of t_src.
struct SeqMap {
  int kind, a, b, c;
  const int* perm;   // kind 3: device array [n_perm]
  int n_perm;
};

__device__ __forceinline__ int seq_source(const SeqMap& m, int t, int T) {
  switch (m.kind) {
    case 1: return T - 1 - t;
    case 2: { int s = t - m.a; return s < 0 ? s + T : s; }                 // out[(i + a) % T] = in[i]
    case 3: { const int q = t / m.c; return __ldg(m.perm + q) * m.c + (t - q * m.c); }
    case 4:
      if (t >= m.a && t < m.a + m.c) return m.b + (t - m.a);
      if (t >= m.b && t < m.b + m.c) return m.a + (t - m.b);
      return t;
    default: return t;
  }
}

// One block handles AUG_TILE consecutive output samples of one clip; a thread owns AUG_PER_THREAD of
// them, 256 apart (every access of a warp is a contiguous - or, for `reverse`, reversed - 128-byte
// run).  All index arithmetic is 32-bit; the loads of a thread's samples are issued before any store.
constexpr int AUG_PER_THREAD = 8;
constexpr int AUG_TILE = 256 * AUG_PER_THREAD;

__global__ void __launch_bounds__(256)
augment_gather_kernel(const float* __restrict__ original, const float* __restrict__ watermarked,
                      const float* __restrict__ gt_in, int B, int T, const uint8_t* __restrict__ seg_op,
                      const int* __restrict__ seg_src, int seg_len, int n_seg, SeqMap map, int T_out,
                      float* __restrict__ out_wm, float* __restrict__ out_orig, float* __restrict__ out_gt) {
  const int tiles = (T_out + AUG_TILE - 1) / AUG_TILE;
  const long long work = static_cast<long long>(B) * tiles;
  for (long long w = blockIdx.x; w < work; w += gridDim.x) {
    const int b = static_cast<int>(w / tiles);
    const int t0 = static_cast<int>(w - static_cast<long long>(b) * tiles) * AUG_TILE + threadIdx.x;
    const float* wrow = watermarked + static_cast<long long>(b) * T;
    const float* grow = gt_in != nullptr ? gt_in + static_cast<long long>(b) * T : nullptr;
    const uint8_t* ops = seg_op != nullptr ? seg_op + static_cast<long long>(b) * n_seg : nullptr;
    const int* srcs = seg_op != nullptr ? seg_src + static_cast<long long>(b) * n_seg : nullptr;
    // pass 1: source index, segment operation and source row per sample (ALU + two tiny table reads)
    int ts[AUG_PER_THREAD], op[AUG_PER_THREAD];
    const float* orow[AUG_PER_THREAD];
#pragma unroll
    for (int j = 0; j < AUG_PER_THREAD; ++j) {
      const int t = t0 + 256 * j;
      ts[j] = t < T_out ? seq_source(map, t, T) : 0;
      op[j] = t < T_out ? 0 : 2;                                   // past the end: no loads
      int src = b;
      if (ops != nullptr && t < T_out) {
        const unsigned s = static_cast<unsigned>(ts[j]) / static_cast<unsigned>(seg_len);
        op[j] = ops[s];
        if (op[j] == 3) src = srcs[s];
      }
      orow[j] = original + static_cast<long long>(src) * T;
    }
    // pass 2: all loads of the thread in flight together (predicated, no divergent control flow)
    float wm[AUG_PER_THREAD], og[AUG_PER_THREAD], gt[AUG_PER_THREAD];
    const bool need_orig = out_orig != nullptr;
#pragma unroll
    for (int j = 0; j < AUG_PER_THREAD; ++j) {
      const bool keep = op[j] == 0, zero = op[j] == 2;
      const float o = (!zero && (need_orig || !keep)) ? __ldg(orow[j] + ts[j]) : 0.f;
      const float w_ = keep ? __ldg(wrow + ts[j]) : o;
      const float g_ = keep ? (grow != nullptr ? __ldg(grow + ts[j]) : 1.f) : 0.f;
      wm[j] = w_; og[j] = o; gt[j] = g_;
    }
    const long long obase = static_cast<long long>(b) * T_out;
#pragma unroll
    for (int j = 0; j < AUG_PER_THREAD; ++j) {
      const int t = t0 + 256 * j;
      if (t >= T_out) continue;
      if (out_wm != nullptr) __stcs(out_wm + obase + t, wm[j]);
      if (out_orig != nullptr) __stcs(out_orig + obase + t, og[j]);
      if (out_gt != nullptr) __stcs(out_gt + obase + t, gt[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Pointwise effects (utils/effect_augmentation.py): 0 identity (:1364), 1 amplitude_scaling
// (x * scale, :2000-2028), 2 quantization (round-half-even(x * m) / m, m = 2^(bits-1) - 1,
// :1090-1111), 3 additive Gaussian noise with the N(0,1) draw supplied by the caller
// (x + noise * std, :2105-2133 random_noise / :2338-2368 white_noise: bit-exact against torch when
// the same draw is used), 4 the same with the draw generated in the kernel (Philox4x32-10 +
// Box-Muller, keyed by (seed, element index)).
// Explicitly rounded intrinsics: the reference rounds after every operation (no FMA contraction).
__device__ __forceinline__ uint2 mulhilo32(uint32_t a, uint32_t b) {
  const unsigned long long p = static_cast<unsigned long long>(a) * b;
  return make_uint2(static_cast<uint32_t>(p), static_cast<uint32_t>(p >> 32));   // (lo, hi)
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint2 p0 = mulhilo32(0xD2511F53u, ctr.x), p1 = mulhilo32(0xCD9E8D57u, ctr.z);
    ctr = make_uint4(p1.y ^ ctr.y ^ key.x, p1.x, p0.y ^ ctr.w ^ key.y, p0.x);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

__device__ __forceinline__ float4 gauss4(unsigned long long seed, unsigned long long quad) {
  const uint4 r = philox4x32_10(make_uint4(static_cast<uint32_t>(quad), static_cast<uint32_t>(quad >> 32), 0u, 0u),
                                make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)));
  const float k = 2.3283064365386963e-10f;   // 2^-32
  const float u0 = (static_cast<float>(r.x) + 0.5f) * k, u1 = static_cast<float>(r.y) * k;
  const float u2 = (static_cast<float>(r.z) + 0.5f) * k, u3 = static_cast<float>(r.w) * k;
  const float m0 = sqrtf(-2.f * __logf(fminf(u0, 0.99999994f))), m1 = sqrtf(-2.f * __logf(fminf(u2, 0.99999994f)));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  return make_float4(m0 * c0, m0 * s0, m1 * c1, m1 * s1);
}

__device__ __forceinline__ float effect_one(int effect, float x, float p0, float nz) {
  switch (effect) {
    case 1: return __fmul_rn(x, p0);
    case 2: return __fdiv_rn(rintf(__fmul_rn(x, p0)), p0);
    case 3:
    case 4: return __fadd_rn(x, __fmul_rn(nz, p0));
    default: return x;
  }
}

__global__ void __launch_bounds__(256)
effect_pointwise_kernel(int effect, const float* __restrict__ in, long long n, float p0,
                        const float* __restrict__ noise, unsigned long long seed, float* __restrict__ out) {
  const long long n4 = n >> 2;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long start = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  for (long long q = start; q < n4; q += stride) {
    const float4 x = __ldcs(reinterpret_cast<const float4*>(in) + q);
    float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    if (effect == 3) z = __ldcs(reinterpret_cast<const float4*>(noise) + q);
    else if (effect == 4) z = gauss4(seed, static_cast<unsigned long long>(q));
    float4 y;
    y.x = effect_one(effect, x.x, p0, z.x);
    y.y = effect_one(effect, x.y, p0, z.y);
    y.z = effect_one(effect, x.z, p0, z.z);
    y.w = effect_one(effect, x.w, p0, z.w);
    reinterpret_cast<float4*>(out)[q] = y;
  }
  for (long long i = (n4 << 2) + start; i < n; i += stride) {   // tail (< 4 elements)
    float z = 0.f;
    if (effect == 3) z = noise[i];
    else if (effect == 4) {
      const float4 g = gauss4(seed, static_cast<unsigned long long>(n4));
      const int j = static_cast<int>(i - (n4 << 2));
      z = j == 0 ? g.x : j == 1 ? g.y : g.z;
    }
    out[i] = effect_one(effect, in[i], p0, z);
  }
}

// sample_suppression (utils/effect_augmentation.py:2061-2103): audio[b, idx] = 0 and mask[b, idx] = 0
// for the k indices drawn per clip (the draw - torch.randperm - stays on the host side).
__global__ void __launch_bounds__(256)
effect_suppress_kernel(float* __restrict__ audio, float* __restrict__ mask, const long long* __restrict__ idx,
                       int B, int T, int k) {
  const long long total = static_cast<long long>(B) * k;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / k);
    const long long t = idx[i];
    if (t < 0 || t >= T) continue;
    audio[static_cast<long long>(b) * T + t] = 0.f;
    if (mask != nullptr) mask[static_cast<long long>(b) * T + t] = 0.f;
  }
}

// median_filter (utils/effect_augmentation.py:1246-1312 -> scipy.signal.medfilt): odd window K, the
// signal is zero-padded at both ends; exact (selection only, no arithmetic).  A block stages its
// samples + halo in shared memory; each thread insertion-sorts its window in registers.
constexpr int MEDIAN_MAX_K = 31;
constexpr int MEDIAN_THREADS = 256;
constexpr int MEDIAN_PER_THREAD = 8;
constexpr int MEDIAN_TILE = MEDIAN_THREADS * MEDIAN_PER_THREAD;

template <int K>
__global__ void __launch_bounds__(MEDIAN_THREADS)
effect_median_kernel(const float* __restrict__ in, int B, int T, float* __restrict__ out) {
  constexpr int H = K / 2;
  __shared__ float sh[MEDIAN_TILE + 2 * (MEDIAN_MAX_K / 2)];
  const int tiles = (T + MEDIAN_TILE - 1) / MEDIAN_TILE;
  for (long long tile = blockIdx.x; tile < static_cast<long long>(B) * tiles; tile += gridDim.x) {
    const int b = static_cast<int>(tile / tiles);
    const int t0 = static_cast<int>(tile - static_cast<long long>(b) * tiles) * MEDIAN_TILE;
    const float* row = in + static_cast<long long>(b) * T;
    __syncthreads();
    for (int i = threadIdx.x; i < MEDIAN_TILE + 2 * H; i += MEDIAN_THREADS) {
      const int t = t0 - H + i;
      sh[i] = (t >= 0 && t < T) ? __ldg(row + t) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < MEDIAN_PER_THREAD; ++u) {
      const int l = threadIdx.x + u * MEDIAN_THREADS;
      const int t = t0 + l;
      if (t >= T) break;
      float w[K];
#pragma unroll
      for (int j = 0; j < K; ++j) w[j] = sh[l + j];
      // partial selection sort up to the median position
#pragma unroll
      for (int i = 0; i <= H; ++i) {
#pragma unroll
        for (int j = i + 1; j < K; ++j) {
          const float lo = fminf(w[i], w[j]), hi = fmaxf(w[i], w[j]);
          w[i] = lo; w[j] = hi;
        }
      }
      __stcs(out + static_cast<long long>(b) * T + t, w[H]);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// FIR effects: julius.lowpass_filter / highpass_filter / bandpass_filter as the reference calls them
// (utils/effect_augmentation.py:1684-1871).  julius (third party, unpinned in requirements.txt, absent
// from the reference tree) builds windowed-sinc low-pass filters of 2*half+1 taps, pads the signal with
// `half` replicated samples on both sides and cross-correlates; high-pass = x - lowpass(x), band-pass =
// lowpass_high(x) - lowpass_low(x) with a shared length (taps pre-subtracted on the host).
//   y[t] = sum_k h[k] * x[clamp(t + k - half, 0, T-1)];   out = subtract ? x[t] - y[t] : y[t]
// One block = FIR_TILE outputs of one clip; samples + halo and the taps live in shared memory; fp32 FMAs.
constexpr int FIR_THREADS = 256;
constexpr int FIR_PER_THREAD = 4;
constexpr int FIR_TILE = FIR_THREADS * FIR_PER_THREAD;
constexpr int FIR_MAX_TAPS = 2049;

// A thread owns FIR_PER_THREAD = 4 CONSECUTIVE outputs and slides an 8-sample register window over its inputs: per four
// taps one 16-byte shared-memory read of samples and one (broadcast) of taps feed 16 FMAs.
__global__ void __launch_bounds__(FIR_THREADS)
effect_fir_kernel(const float* __restrict__ in, const float* __restrict__ taps, int n_taps, int B, int T,
                  int subtract, float* __restrict__ out) {
  extern __shared__ __align__(16) float fir_sh[];   // [taps padded to 4] then [FIR_TILE + taps_pad] samples
  const int taps_pad = (n_taps + 3) & ~3;
  float* hs = fir_sh;
  float* xs = fir_sh + taps_pad;
  const int half = n_taps >> 1;
  for (int i = threadIdx.x; i < taps_pad; i += FIR_THREADS) hs[i] = i < n_taps ? __ldg(taps + i) : 0.f;
  const int tiles = (T + FIR_TILE - 1) / FIR_TILE;
  for (long long tile = blockIdx.x; tile < static_cast<long long>(B) * tiles; tile += gridDim.x) {
    const int b = static_cast<int>(tile / tiles);
    const int t0 = static_cast<int>(tile - static_cast<long long>(b) * tiles) * FIR_TILE;
    const float* row = in + static_cast<long long>(b) * T;
    __syncthreads();
    for (int i = threadIdx.x; i < FIR_TILE + taps_pad; i += FIR_THREADS) {
      const int t = min(max(t0 - half + i, 0), T - 1);   // replicate padding
      xs[i] = __ldg(row + t);
    }
    __syncthreads();
    const float4* xv = reinterpret_cast<const float4*>(xs) + threadIdx.x;   // samples 4*tid .. of the tile
    const float4* hv = reinterpret_cast<const float4*>(hs);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float4 lo = xv[0];
    for (int j = 0; j < taps_pad / 4; ++j) {
      const float4 hi = xv[j + 1];
      const float4 h = hv[j];
      const float w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
      const float hh[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = fmaf(hh[kk], w[kk + u], acc[u]);
      lo = hi;
    }
    const int l = threadIdx.x * 4;
    float y[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) y[u] = subtract ? xs[l + u + half] - acc[u] : acc[u];
    float* op = out + static_cast<long long>(b) * T + t0 + l;
    if ((T & 3) == 0 && t0 + l + 3 < T) {
      __stcs(reinterpret_cast<float4*>(op), make_float4(y[0], y[1], y[2], y[3]));   // rows are 16-byte aligned when T % 4 == 0
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (t0 + l + u < T) __stcs(op + u, y[u]);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Polyphase windowed-sinc resampler = torchaudio.transforms.Resample as the reference's `resample` effect calls it
// (utils/effect_augmentation.py:1451-1502: down to new_sample_rate and back), and the resampling half of `speed`
// (:1381-1449: SoX speed + rate, then utils/effect_augmentation.py:187-215 stretches the result back to the input
// length by linear interpolation).  With orig / new reduced by their gcd and ntaps = 2*width + orig:
//   y[b, q*new + p] = sum_k h[p][k] * x[b, q*orig + k - width]          (x zero outside [0, T))
// lerp != 0: out[j] = linear interpolation of y (length T_mid) at PyTorch's align_corners=False source index
// (src = (j + 0.5) * T_mid / T_out - 0.5, clamped at 0), i.e. the two steps of `speed` in one pass.
__global__ void __launch_bounds__(256)
effect_resample_kernel(const float* __restrict__ in, const float* __restrict__ taps, int B, int T, int orig, int nw,
                       int width, int ntaps, int T_mid, int T_out, int lerp, float* __restrict__ out) {
  const long long total = static_cast<long long>(B) * T_out;
  const float scale = static_cast<float>(T_mid) / static_cast<float>(T_out);
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(idx % T_out);
    const int b = static_cast<int>(idx / T_out);
    const float* xp = in + static_cast<long long>(b) * T;
    auto fir = [&](int i) {
      const int p = i % nw, q = i / nw;
      const float* h = taps + static_cast<long long>(p) * ntaps;
      const int t0 = q * orig - width;
      float acc = 0.f;
      for (int k = 0; k < ntaps; ++k) {
        const int t = t0 + k;
        if (t >= 0 && t < T) acc = fmaf(__ldg(h + k), __ldg(xp + t), acc);
      }
      return acc;
    };
    float v;
    if (!lerp) {
      v = fir(j);
    } else {
      float src = (static_cast<float>(j) + 0.5f) * scale - 0.5f;
      if (src < 0.f) src = 0.f;
      const int i0 = static_cast<int>(src);
      const int i1 = i0 + (i0 < T_mid - 1 ? 1 : 0);
      const float l1 = src - static_cast<float>(i0), l0 = 1.f - l1;
      v = l0 * fir(i0) + l1 * fir(i1);
    }
    out[idx] = v;
  }
}

}  // namespace wv
