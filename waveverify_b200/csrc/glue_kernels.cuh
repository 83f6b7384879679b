// Memory-bound kernels around the GEMMs.  Activations are channels-last fp16 [B, T, C]; every
// thread owns 8 consecutive channels (one 128-bit vector) so a warp reads/writes whole rows
// coalesced.  Depthwise weights are stored tap-major fp32 [k][C] for the same reason.
#pragma once
#include "ptx_sm100.cuh"

namespace wv {

struct F8 {
  float v[8];
};
// (elu1 comes from ptx_sm100.cuh)
__device__ __forceinline__ F8 ld_act8(const act_t* p) {   // activations: L2-coherent load
  const uint4 u = __ldcg(reinterpret_cast<const uint4*>(p));
  F8 r;
  unpack_act2(u.x, r.v[0], r.v[1]);
  unpack_act2(u.y, r.v[2], r.v[3]);
  unpack_act2(u.z, r.v[4], r.v[5]);
  unpack_act2(u.w, r.v[6], r.v[7]);
  return r;
}
__device__ __forceinline__ void st_act8(act_t* p, const F8& r) {
  *reinterpret_cast<uint4*>(p) =
      make_uint4(pack_act2(r.v[0], r.v[1]), pack_act2(r.v[2], r.v[3]),
                 pack_act2(r.v[4], r.v[5]), pack_act2(r.v[6], r.v[7]));
}
__device__ __forceinline__ F8 ld_f32x8(const float* p) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  F8 r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}

// ---------------------------------------------------------------------------------------
// Causal depthwise conv k=5 (SConv1d groups=C, modules/conv.py:715-763; left pad 4 zeros):
//   v[t,c] = bias[c] + sum_j w[j][c] * in[t-4+j, c]  (+ residual[t,c])
//   out_raw = v ; out_act = ELU(v * act_scale)            (each optional)
// The resblock tail scale RS*res_scale_param (modules/seanet.py:271-277) is folded into w/bias.
constexpr int DW_TT = 8;  // consecutive time steps per thread (sliding window in registers)

__global__ void __launch_bounds__(256)
dw5_kernel(const act_t* __restrict__ in, const float* __restrict__ w,
           const float* __restrict__ bias, const act_t* __restrict__ residual,
           act_t* __restrict__ out_raw, act_t* __restrict__ out_act,
           float act_scale, int B, int T, int C) {
  pdl_wait();
  pdl_launch_dependents();
  const int C8 = C >> 3;
  const int runs = (T + DW_TT - 1) / DW_TT;
  const long long total = static_cast<long long>(B) * runs * C8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(idx % C8);
    const long long rr = idx / C8;
    const int run = static_cast<int>(rr % runs);
    const int b = static_cast<int>(rr / runs);
    const int c = cg * 8;
    const int t0 = run * DW_TT;
    F8 wt[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) wt[j] = ld_f32x8(w + j * C + c);
    F8 bs;
    if (bias != nullptr) bs = ld_f32x8(bias + c);
    else {
#pragma unroll
      for (int i = 0; i < 8; ++i) bs.v[i] = 0.f;
    }
    const act_t* ip = in + (static_cast<long long>(b) * T) * C + c;
    F8 win[5];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 - 4 + j;
      if (t >= 0) win[j + 1] = ld_act8(ip + static_cast<long long>(t) * C);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) win[j + 1].v[i] = 0.f;
      }
    }
#pragma unroll
    for (int s = 0; s < DW_TT; ++s) {
      const int t = t0 + s;
      if (t >= T) break;
#pragma unroll
      for (int j = 0; j < 4; ++j) win[j] = win[j + 1];
      win[4] = ld_act8(ip + static_cast<long long>(t) * C);
      F8 o = bs;
#pragma unroll
      for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = fmaf(wt[j].v[i], win[j].v[i], o.v[i]);
      const long long off = (static_cast<long long>(b) * T + t) * C + c;
      if (residual != nullptr) {
        const F8 rs = ld_act8(residual + off);
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] += rs.v[i];
      }
      if (out_raw != nullptr) st_act8(out_raw + off, o);
      if (out_act != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o.v[i] = elu1(o.v[i] * act_scale);
        st_act8(out_act + off, o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Causal strided depthwise down-conv k=2r, s=r (modules/seanet.py:759-770; left pad r zeros,
// right zero padding up to a whole window, modules/conv.py:160-203) + FiLM (seanet.py:928-966):
//   v[i,c] = bias[c] + sum_{j<2r} w[j][c] * in[i*r - r + j, c];  v = v*gamma[b,band] + beta
template <int R>
__global__ void __launch_bounds__(256)
down_kernel(const act_t* __restrict__ in, const float* __restrict__ w,
            const float* __restrict__ bias, const float* __restrict__ film, int film_stride,
            int bands, act_t* __restrict__ out_raw, act_t* __restrict__ out_act,
            float act_scale, int B, int Tin, int Tout, int C) {
  pdl_wait();
  pdl_launch_dependents();
  const int C8 = C >> 3;
  const long long total = static_cast<long long>(B) * Tout * C8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(idx % C8);
    const long long rr = idx / C8;
    const int i = static_cast<int>(rr % Tout);
    const int b = static_cast<int>(rr / Tout);
    const int c = cg * 8;
    const act_t* ip = in + (static_cast<long long>(b) * Tin) * C + c;
    const int tb = i * R - R;
    uint4 xv[2 * R];                       // all 2R window rows in flight before any math
#pragma unroll
    for (int j = 0; j < 2 * R; ++j) {
      const int t = tb + j;
      xv[j] = (t >= 0 && t < Tin) ? __ldcg(reinterpret_cast<const uint4*>(ip + static_cast<long long>(t) * C))
                                  : make_uint4(0, 0, 0, 0);
    }
    F8 o;
    if (bias != nullptr) o = ld_f32x8(bias + c);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < 2 * R; ++j) {
      const F8 wj = ld_f32x8(w + j * C + c);
      float x[8];
      unpack_act2(xv[j].x, x[0], x[1]);
      unpack_act2(xv[j].y, x[2], x[3]);
      unpack_act2(xv[j].z, x[4], x[5]);
      unpack_act2(xv[j].w, x[6], x[7]);
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = fmaf(wj.v[k], x[k], o.v[k]);
    }
    if (film != nullptr) {
      const int band = c / (C / bands);
      const float gm = __ldcg(film + static_cast<long long>(b) * film_stride + band * 2);
      const float bt = __ldcg(film + static_cast<long long>(b) * film_stride + band * 2 + 1);
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = fmaf(o.v[k], gm, bt);
    }
    const long long off = (static_cast<long long>(b) * Tout + i) * C + c;
    if (out_raw != nullptr) st_act8(out_raw + off, o);
    if (out_act != nullptr) {
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = elu1(o.v[k] * act_scale);
      st_act8(out_act + off, o);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Causal depthwise transposed conv k=2r, s=r, right-trim r (modules/conv.py:838-874):
//   out[i*r + j, c] = a[i,c]*w[j][c] + a[i-1,c]*w[j+r][c],  0 <= j < r,  a[-1] = 0
__global__ void __launch_bounds__(256)
up_kernel(const act_t* __restrict__ in, const float* __restrict__ w,
          act_t* __restrict__ out, int B, int Tin, int C, int r) {
  pdl_wait();
  pdl_launch_dependents();
  const int C8 = C >> 3;
  const long long total = static_cast<long long>(B) * Tin * C8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(idx % C8);
    const long long rr = idx / C8;
    const int i = static_cast<int>(rr % Tin);
    const int b = static_cast<int>(rr / Tin);
    const int c = cg * 8;
    const act_t* ip = in + (static_cast<long long>(b) * Tin + i) * C + c;
    const F8 a0 = ld_act8(ip);
    F8 a1;
    if (i > 0) a1 = ld_act8(ip - C);
    else {
#pragma unroll
      for (int k = 0; k < 8; ++k) a1.v[k] = 0.f;
    }
    act_t* op = out + (static_cast<long long>(b) * Tin * r + static_cast<long long>(i) * r) * C + c;
    for (int j = 0; j < r; ++j) {
      const F8 w0 = ld_f32x8(w + j * C + c);
      const F8 w1 = ld_f32x8(w + (j + r) * C + c);
      F8 o;
#pragma unroll
      for (int k = 0; k < 8; ++k) o.v[k] = fmaf(a0.v[k], w0.v[k], a1.v[k] * w1.v[k]);
      st_act8(op + static_cast<long long>(j) * C, o);
    }
  }
}

// ---------------------------------------------------------------------------------------
// conv_pre: 1 -> C, k=5 causal, 1/wav_std folded into w (modules/seanet.py:657-664):
//   v[t,c] = bias[c] + sum_j w[j][c] * x[t-4+j];  out_raw = v, out_act = ELU(v*act_scale)
constexpr int PRE_TT = 8;   // consecutive time steps per thread (taps loaded once)
__global__ void __launch_bounds__(256)
conv_pre_kernel(const float* __restrict__ x, const float* __restrict__ w,
                const float* __restrict__ bias, act_t* __restrict__ out_raw,
                act_t* __restrict__ out_act, float act_scale, int B, int T, int C) {
  pdl_wait();
  pdl_launch_dependents();
  const int C8 = C >> 3;
  const int runs = (T + PRE_TT - 1) / PRE_TT;
  const long long total = static_cast<long long>(B) * runs * C8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(idx % C8);
    const long long rr = idx / C8;
    const int t0 = static_cast<int>(rr % runs) * PRE_TT;
    const int b = static_cast<int>(rr / runs);
    const int c = cg * 8;
    const float* xp = x + static_cast<long long>(b) * T;
    float xs[PRE_TT + 4];
#pragma unroll
    for (int j = 0; j < PRE_TT + 4; ++j) {
      const int tt = t0 - 4 + j;
      xs[j] = (tt >= 0 && tt < T) ? __ldcg(xp + tt) : 0.f;
    }
    F8 wt[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) wt[j] = ld_f32x8(w + j * C + c);
    const F8 bs = ld_f32x8(bias + c);
#pragma unroll
    for (int s = 0; s < PRE_TT; ++s) {
      const int t = t0 + s;
      if (t >= T) break;
      F8 o = bs;
#pragma unroll
      for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = fmaf(wt[j].v[k], xs[s + j], o.v[k]);
      const long long off = (static_cast<long long>(b) * T + t) * C + c;
      if (out_raw != nullptr) st_act8(out_raw + off, o);
      if (out_act != nullptr) {
#pragma unroll
        for (int k = 0; k < 8; ++k) o.v[k] = elu1(o.v[k] * act_scale);
        st_act8(out_act + off, o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Decoder tail (modules/seanet.py:1177-1202) fused with the trim (model/generator.py:410) and the
// watermark add (model/watermarking.py:440):  in = ELU'd activations [B, Tp, C] (Tp >= T)
//   s[t] = b + sum_j sum_c w[j][c] * in[t-4+j, c];  wm = tanh(s)   (wav_std folded into w, b)
//   wm_out[b,t] = wm ; y_out[b,t] = x[b,t] + wm        for t < T
// One block per 128-step time tile; the tile (+4 halo rows) is staged in shared memory with a
// padded row pitch so that the 16-byte reads of consecutive rows hit distinct bank groups.
constexpr int CL_TILE = 128;
__global__ void __launch_bounds__(CL_TILE)
conv_last_kernel(const act_t* __restrict__ in, const float* __restrict__ w, float bias,
                 const float* __restrict__ x, float* __restrict__ wm_out,
                 float* __restrict__ y_out, int B, int Tp, int T, int C) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ uint8_t cl_smem[];
  const int pitch = C * 2 + 16;  // bytes
  float* ws = reinterpret_cast<float*>(cl_smem);                        // [5][C]
  uint8_t* tile = cl_smem + ((5 * C * 4 + 15) & ~15);                    // [(CL_TILE+4)][pitch]
  const int tiles = (T + CL_TILE - 1) / CL_TILE;
  const int b = blockIdx.x / tiles;
  const int t0 = (blockIdx.x % tiles) * CL_TILE;
  for (int i = threadIdx.x; i < 5 * C; i += blockDim.x) ws[i] = w[i];
  const int C8 = C >> 3;
  for (int i = threadIdx.x; i < (CL_TILE + 4) * C8; i += blockDim.x) {
    const int row = i / C8, cg = i % C8;
    const int t = t0 - 4 + row;
    uint4 u = make_uint4(0, 0, 0, 0);
    if (t >= 0 && t < Tp)
      u = __ldcg(reinterpret_cast<const uint4*>(in + (static_cast<long long>(b) * Tp + t) * C) + cg);
    *reinterpret_cast<uint4*>(tile + row * pitch + cg * 16) = u;
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= T) return;
  float acc = bias;
  for (int j = 0; j < 5; ++j) {
    const uint8_t* rp = tile + (threadIdx.x + j) * pitch;
    for (int cg = 0; cg < C8; ++cg) {
      const uint4 u = *reinterpret_cast<const uint4*>(rp + cg * 16);
      float a[8];
      unpack_act2(u.x, a[0], a[1]);
      unpack_act2(u.y, a[2], a[3]);
      unpack_act2(u.z, a[4], a[5]);
      unpack_act2(u.w, a[6], a[7]);
      const float* wp = ws + j * C + cg * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc = fmaf(a[k], wp[k], acc);
    }
  }
  const float wm = tanhf(acc);
  const long long o = static_cast<long long>(b) * T + t;
  if (wm_out != nullptr) wm_out[o] = wm;
  if (y_out != nullptr) y_out[o] = __ldcg(x + o) + wm;
}

// ---------------------------------------------------------------------------------------
// Waveform staging for the conv-as-DFT STFTs (modules/conv.py:1036-1068): fp16 copies of x*scale
// per clip with `lead` leading zeros (the causal n_fft-1 padding of the largest scale) and a zero
// tail up to the row pitch.  Eight copies per clip, copy s shifted left by s samples: a frame
// sequence with hop h < 8 is then `8/h` interleaved views with a 16-byte row stride (the TMA stride
// unit), phase r reading copy r*h - no frame matrix is materialised.   out[b][s][p] = staged[p + s]
constexpr int WAV_COPIES = 8;
__global__ void __launch_bounds__(256)
wav_stage_kernel(const float* __restrict__ x, __half* __restrict__ out, float scale, int B, int T,
                 int lead, int pitch) {
  pdl_wait();
  pdl_launch_dependents();
  const int p8n = pitch >> 3;                                  // pitch is a multiple of 8: one 16-byte store per thread
  const long long total = static_cast<long long>(B) * WAV_COPIES * p8n;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(idx % p8n) * 8;
    const long long bs = idx / p8n;
    const int sft = static_cast<int>(bs % WAV_COPIES);
    const int b = static_cast<int>(bs / WAV_COPIES);
    const int t0 = p + sft - lead;
    const float* xp = x + static_cast<long long>(b) * T;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (t0 + k >= 0 && t0 + k < T) ? __ldcg(xp + t0 + k) * scale : 0.f;
    *reinterpret_cast<uint4*>(out + bs * pitch + p) =
        make_uint4(pack_act2(v[0], v[1]), pack_act2(v[2], v[3]), pack_act2(v[4], v[5]), pack_act2(v[6], v[7]));
  }
}

// Frame matrix for hops that TMA cannot stride (hop*2 B not a multiple of 16):
//   frames[b*F + f, n] = wav16[b, base + f*hop + n],  n < n_fft
__global__ void __launch_bounds__(256)
frames_kernel(const __half* __restrict__ wav16, __half* __restrict__ frames, int B, int F, int hop,
              int n_fft, int base, int pitch) {
  pdl_wait();
  pdl_launch_dependents();
  const int N8 = n_fft >> 3;
  const long long total = static_cast<long long>(B) * F * N8;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n8 = static_cast<int>(idx % N8);
    const long long rr = idx / N8;
    const int f = static_cast<int>(rr % F);
    const int b = static_cast<int>(rr / F);
    const __half* sp = wav16 + static_cast<long long>(b) * pitch + base + f * hop + n8 * 8;
    __half h[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) h[k] = __ldcg(sp + k);
    *reinterpret_cast<uint4*>(frames + (rr * n_fft) + n8 * 8) = *reinterpret_cast<uint4*>(h);
  }
}

// ---------------------------------------------------------------------------------------
// Precise-mode glue (see gemm_sm100.cuh "PRECISE MODE"): fp32 raw streams, split-fp16 GEMM operands.
// conv_pre: raw fp32 [B,T,C] + activated split [B,T,2C]
__global__ void __launch_bounds__(256)
conv_pre_pm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ out_raw, act_t* __restrict__ out_act, float act_scale, int B, int T, int C) {
  pdl_wait();
  pdl_launch_dependents();
  const int C4 = C >> 2;
  const int runs = (T + PRE_TT - 1) / PRE_TT;
  const long long total = static_cast<long long>(B) * runs * C4;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(idx % C4);
    const long long rr = idx / C4;
    const int t0 = static_cast<int>(rr % runs) * PRE_TT;
    const int b = static_cast<int>(rr / runs);
    const int c = cg * 4;
    const float* xp = x + static_cast<long long>(b) * T;
    float xs[PRE_TT + 4];
#pragma unroll
    for (int j = 0; j < PRE_TT + 4; ++j) {
      const int tt = t0 - 4 + j;
      xs[j] = (tt >= 0 && tt < T) ? __ldcg(xp + tt) : 0.f;
    }
    float4 wt[5];
#pragma unroll
    for (int j = 0; j < 5; ++j) wt[j] = __ldg(reinterpret_cast<const float4*>(w + j * C + c));
    const float4 bs = __ldg(reinterpret_cast<const float4*>(bias + c));
#pragma unroll
    for (int s = 0; s < PRE_TT; ++s) {
      const int t = t0 + s;
      if (t >= T) break;
      float4 o = bs;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        o.x = fmaf(wt[j].x, xs[s + j], o.x); o.y = fmaf(wt[j].y, xs[s + j], o.y);
        o.z = fmaf(wt[j].z, xs[s + j], o.z); o.w = fmaf(wt[j].w, xs[s + j], o.w);
      }
      const long long row = static_cast<long long>(b) * T + t;
      if (out_raw != nullptr) *reinterpret_cast<float4*>(out_raw + row * C + c) = o;
      if (out_act != nullptr)
        st_split4(out_act + row * 2 * C + c, C, elu_precise(o.x * act_scale), elu_precise(o.y * act_scale),
                  elu_precise(o.z * act_scale), elu_precise(o.w * act_scale));
    }
  }
}

// causal depthwise k=5 on a split tensor [B,T,2C] -> split [B,T,2C] (conv_post of the encoder tail; no activation)
__global__ void __launch_bounds__(256)
dw5_pm_kernel(const act_t* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
              act_t* __restrict__ out, int B, int T, int C) {
  pdl_wait();
  pdl_launch_dependents();
  const int C4 = C >> 2;
  const long long total = static_cast<long long>(B) * T * C4;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cg = static_cast<int>(idx % C4);
    const long long rr = idx / C4;
    const int t = static_cast<int>(rr % T);
    const int b = static_cast<int>(rr / T);
    const int c = cg * 4;
    float4 o = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int tt = t - 4 + j;
      if (tt < 0) continue;
      const act_t* ip = in + (static_cast<long long>(b) * T + tt) * 2 * C + c;
      const uint2 h = __ldcg(reinterpret_cast<const uint2*>(ip));
      const uint2 l = __ldcg(reinterpret_cast<const uint2*>(ip + C));
      float h0, h1, h2, h3, l0, l1, l2, l3;
      unpack_act2(h.x, h0, h1); unpack_act2(h.y, h2, h3);
      unpack_act2(l.x, l0, l1); unpack_act2(l.y, l2, l3);
      const float4 wj = __ldg(reinterpret_cast<const float4*>(w + j * C + c));
      o.x = fmaf(wj.x, h0 + l0, o.x); o.y = fmaf(wj.y, h1 + l1, o.y);
      o.z = fmaf(wj.z, h2 + l2, o.z); o.w = fmaf(wj.w, h3 + l3, o.w);
    }
    st_split4(out + (static_cast<long long>(b) * T + t) * 2 * C + c, C, o.x, o.y, o.z, o.w);
  }
}

// waveform staging as split pairs: out_hi / out_lo [B][8 copies][pitch] (cf. wav_stage_kernel)
__global__ void __launch_bounds__(256)
wav_stage_pm_kernel(const float* __restrict__ x, __half* __restrict__ out_hi, __half* __restrict__ out_lo, float scale,
                    int B, int T, int lead, int pitch) {
  pdl_wait();
  pdl_launch_dependents();
  const int p8n = pitch >> 3;
  const long long total = static_cast<long long>(B) * WAV_COPIES * p8n;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(idx % p8n) * 8;
    const long long bs = idx / p8n;
    const int sft = static_cast<int>(bs % WAV_COPIES);
    const int b = static_cast<int>(bs / WAV_COPIES);
    const int t0 = p + sft - lead;
    const float* xp = x + static_cast<long long>(b) * T;
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = (t0 + k >= 0 && t0 + k < T) ? __ldcg(xp + t0 + k) * scale : 0.f;
    uint32_t h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) split2(v[2 * k], v[2 * k + 1], h[k], l[k]);
    *reinterpret_cast<uint4*>(out_hi + bs * pitch + p) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(out_lo + bs * pitch + p) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// ---------------------------------------------------------------------------------------
// Message MLP + FiLM scalars (modules/seanet.py:830-839, 518-550): one block of E threads per
// clip.  e = ReLU(L3 ReLU(L2 (L1 m + b1) + b2) + b3);  film[b, s, band, {gamma,beta}].
struct FilmArgs {
  const float* w[4];   // L1 [E,msg], then up to 3 [E,E]
  const float* b[4];
  int n_hidden;        // number of (Linear, ReLU) pairs after L1
  const float* gw;     // [S*bands, E] gamma weights
  const float* gb;     // [S*bands]
  const float* bw;     // [S*bands, E] beta weights
  const float* bb;     // [S*bands]
  int msg_dim, E, n_film;
};
__global__ void film_kernel(const float* __restrict__ msg, float* __restrict__ film, FilmArgs a) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float fs[];
  float* e0 = fs;
  float* e1 = fs + a.E;
  const int b = blockIdx.x, i = threadIdx.x;
  if (i < a.E) {
    float s = a.b[0][i];
    for (int k = 0; k < a.msg_dim; ++k) s = fmaf(a.w[0][i * a.msg_dim + k], __ldcg(msg + b * a.msg_dim + k), s);
    e0[i] = s;
  }
  __syncthreads();
  for (int l = 1; l <= a.n_hidden; ++l) {
    if (i < a.E) {
      float s = a.b[l][i];
      for (int k = 0; k < a.E; ++k) s = fmaf(a.w[l][i * a.E + k], e0[k], s);
      e1[i] = fmaxf(s, 0.f);
    }
    __syncthreads();
    float* t = e0; e0 = e1; e1 = t;
  }
  for (int f = i; f < a.n_film * 2; f += blockDim.x) {
    const int q = f >> 1;
    const float* wv = (f & 1) ? a.bw + q * a.E : a.gw + q * a.E;
    float s = (f & 1) ? a.bb[q] : a.gb[q];
    for (int k = 0; k < a.E; ++k) s = fmaf(wv[k], e0[k], s);
    film[static_cast<long long>(b) * a.n_film * 2 + f] = s;
  }
}

// ---------------------------------------------------------------------------------------
// Bit decode finish (waveverify/core.py:577-586, waveverify/utils.py:385-401, masked variant
// scripts/evaluate.py:471-494): one warp per (clip, bit) sums the per-frame partial sigmoid sums
// in a fixed order (deterministic), divides by T (or mask count + 1e-8) and thresholds >= 0.5.
__global__ void bits_finish_kernel(const float* __restrict__ partial, int F, int tiles_n,
                                   int tiles_per_bit, const uint8_t* __restrict__ presence, int T,
                                   int nbits, uint8_t* __restrict__ bits, float* __restrict__ avg,
                                   uint8_t* __restrict__ valid) {
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x, o = blockIdx.y, lane = threadIdx.x;
  float s = 0.f;
  const int n = F * tiles_per_bit;
  for (int i = lane; i < n; i += 32) {
    const int f = i / tiles_per_bit, tl = i % tiles_per_bit;
    s += __ldcg(partial + (static_cast<long long>(b) * F + f) * tiles_n + o * tiles_per_bit + tl);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  float cnt = static_cast<float>(T);
  bool ok = true;
  if (presence != nullptr) {
    int c = 0;
    for (int t = lane; t < T; t += 32) c += __ldcg(presence + static_cast<long long>(b) * T + t) ? 1 : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    ok = c > 0;
    cnt = static_cast<float>(c) + 1e-8f;
  }
  if (lane == 0) {
    const float a = s / cnt;
    if (avg) avg[b * nbits + o] = a;
    if (bits) bits[b * nbits + o] = a >= 0.5f ? 1 : 0;
    if (valid) valid[b * nbits + o] = ok ? 1 : 0;
  }
}
__global__ void conf_kernel(const float* __restrict__ avg, float* __restrict__ conf, int B,
                            int nbits) {
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float s = 0.f;
  for (int o = 0; o < nbits; ++o) s += __ldcg(avg + b * nbits + o);
  conf[b] = s / nbits;
}

// ---------------------------------------------------------------------------------------
// BER / MIoU counters (scripts/evaluate.py:498-505, 636-656): exact int64 sums, one atomic per
// block per counter.  counters += {bit_errors, valid_bits, I_fg, U_fg, I_bg, U_bg}.
__global__ void __launch_bounds__(256)
metrics_kernel(const uint8_t* __restrict__ bits, const uint8_t* __restrict__ valid,
               const uint8_t* __restrict__ msg, long long n_bits,
               const uint8_t* __restrict__ pred, const uint8_t* __restrict__ gt, long long n_mask,
               unsigned long long* __restrict__ counters) {
  unsigned int c[6] = {0, 0, 0, 0, 0, 0};
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const long long start = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (bits != nullptr)
    for (long long i = start; i < n_bits; i += stride) {
      const bool v = valid ? valid[i] != 0 : true;
      c[0] += (v && ((bits[i] != 0) != (msg[i] != 0))) ? 1 : 0;
      c[1] += v ? 1 : 0;
    }
  if (pred != nullptr)
    for (long long i = start; i < n_mask; i += stride) {
      const bool p = pred[i] != 0, q = gt[i] != 0;
      c[2] += (p && q) ? 1 : 0;
      c[3] += (p || q) ? 1 : 0;
      c[4] += (!p && !q) ? 1 : 0;
      c[5] += (!p || !q) ? 1 : 0;
    }
  __shared__ unsigned int sh[6];
  if (threadIdx.x < 6) sh[threadIdx.x] = 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    unsigned int v = c[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh[k], v);
  }
  __syncthreads();
  if (threadIdx.x < 6 && sh[threadIdx.x])
    atomicAdd(&counters[threadIdx.x], static_cast<unsigned long long>(sh[threadIdx.x]));
}

// fp32 [B, C, F] -> fp16 [B, F, C] (Generator.decode entry, model/generator.py:334)
__global__ void latent_in_kernel(const float* __restrict__ z, act_t* __restrict__ out,
                                 int B, int C, int F) {
  pdl_wait();
  pdl_launch_dependents();
  const long long total = static_cast<long long>(B) * F * C;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    const long long rr = idx / C;
    const int f = static_cast<int>(rr % F);
    const int b = static_cast<int>(rr / F);
    out[idx] = __float2half_rn(__ldcg(z + (static_cast<long long>(b) * C + c) * F + f));
  }
}

}  // namespace wv

// ---------------------------------------------------------------------------------------------
// Detector refine (wv_detector_refine): the decoded bits of the fp16 fast path are re-evaluated by the
// fp32-accurate net for every clip with a bit whose (masked) mean lies within tau of the 0.5 threshold
// (waveverify/core.py:577-586, scripts/evaluate.py:471-494).  Everything runs on the device inside one CUDA
// graph: a selection kernel compacts the near clips and sets the condition of a WHILE node whose body
// (gather K clips -> precise plan -> scatter) runs ceil(count / K) times; no host round trip.
namespace wv {

struct RefineIo {              // per-call pointers (device block written by refine_params_kernel)
  const float* y;              // [B, T]
  const uint8_t* presence;     // [B, T] or null
  float* logits;               // [B, nbits, T] or null
  uint8_t* bits;               // [B, nbits]
  float* avg;                  // [B, nbits]
  float* conf;                 // [B] or null
  uint8_t* valid;              // [B, nbits] or null
  int* counters;               // device int[2]: += {clips re-evaluated, passes of the precise net}, or null
  float tau, tau_short;
  int short_samples;
};
struct RefineState { int count, it; };

__global__ void refine_params_kernel(RefineIo* dst, RefineIo v) { *dst = v; }

// unmasked samples per clip (masked decode only): the threshold band depends on how many samples the mean averages
__global__ void __launch_bounds__(256)
refine_neff_kernel(const RefineIo* __restrict__ io, int T, int* __restrict__ neff) {
  const uint8_t* p = io->presence + static_cast<long long>(blockIdx.x) * T;
  int c = 0;
  for (int t = threadIdx.x; t < T; t += 256) c += __ldcg(p + t) ? 1 : 0;
  __shared__ int s[8];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int i = 0; i < 8; ++i) tot += s[i];
    neff[blockIdx.x] = tot;
  }
}

// one warp: ordered list of the clips that need the precise net, their count, the loop condition
__global__ void __launch_bounds__(32)
refine_select_kernel(cudaGraphConditionalHandle loop, const RefineIo* __restrict__ io, int B, int nbits, int T,
                     const int* __restrict__ neff /* null: unmasked */, int* __restrict__ list, RefineState* __restrict__ st) {
  const int lane = threadIdx.x;
  const float tau = io->tau, tau_s = io->tau_short;
  const int short_n = io->short_samples;
  const float* avg = io->avg;
  const uint8_t* valid = io->valid;
  int count = 0;
  for (int b0 = 0; b0 < B; b0 += 32) {
    const int b = b0 + lane;
    bool near = false;
    if (b < B) {
      const int n = neff != nullptr ? neff[b] : T;
      const float tb = n >= short_n ? tau : tau_s;
      for (int i = 0; i < nbits; ++i) {
        const bool ok = valid == nullptr || neff == nullptr || valid[b * nbits + i] != 0;
        near = near || (ok && fabsf(avg[b * nbits + i] - 0.5f) < tb);
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, near);
    if (near) list[count + __popc(m & ((1u << lane) - 1u))] = b;
    count += __popc(m);
  }
  if (lane == 0) {
    st->count = count;
    st->it = 0;
    if (io->counters != nullptr && count > 0) atomicAdd(io->counters, count);
    cudaGraphSetConditional(loop, count > 0 ? 1u : 0u);
  }
}

// slot k of pass `it` <- clip list[it*K + k] (slots past the end repeat the last listed clip: same values, never scattered)
__global__ void __launch_bounds__(256)
refine_gather_kernel(const RefineIo* __restrict__ io, const int* __restrict__ list, const RefineState* __restrict__ st,
                     int K, int T, float* __restrict__ ysub, uint8_t* __restrict__ psub) {
  const int k = blockIdx.y;
  const int j = min(st->it * K + k, st->count - 1);
  const long long src = static_cast<long long>(list[j]) * T, dst = static_cast<long long>(k) * T;
  const float* y = io->y;
  const uint8_t* p = io->presence;
  for (int t = blockIdx.x * 256 + threadIdx.x; t < T; t += gridDim.x * 256) {
    ysub[dst + t] = __ldcg(y + src + t);
    if (psub != nullptr) psub[dst + t] = p != nullptr ? __ldcg(p + src + t) : 1;
  }
}

__global__ void __launch_bounds__(256)
refine_scatter_kernel(const RefineIo* __restrict__ io, const int* __restrict__ list, const RefineState* __restrict__ st,
                      int K, int T, int nbits, const uint8_t* __restrict__ bits_s, const float* __restrict__ avg_s,
                      const float* __restrict__ conf_s, const uint8_t* __restrict__ valid_s, const float* __restrict__ logits_s) {
  const int k = blockIdx.y;
  const int j = st->it * K + k;
  if (j >= st->count) return;
  const int b = list[j];
  if (blockIdx.x == 0 && threadIdx.x < nbits) {
    const int i = threadIdx.x;
    io->bits[b * nbits + i] = bits_s[k * nbits + i];
    io->avg[b * nbits + i] = avg_s[k * nbits + i];
    if (io->valid != nullptr) io->valid[b * nbits + i] = valid_s[k * nbits + i];
    if (i == 0 && io->conf != nullptr) io->conf[b] = conf_s[k];
  }
  if (io->logits != nullptr && logits_s != nullptr) {
    const long long n = static_cast<long long>(nbits) * T;
    float* dst = io->logits + static_cast<long long>(b) * n;
    const float* src = logits_s + static_cast<long long>(k) * n;
    for (long long t = blockIdx.x * 256 + threadIdx.x; t < n; t += static_cast<long long>(gridDim.x) * 256) dst[t] = src[t];
  }
}

__global__ void refine_advance_kernel(cudaGraphConditionalHandle loop, const RefineIo* __restrict__ io, RefineState* st, int K) {
  const int it = st->it + 1;
  st->it = it;
  if (io->counters != nullptr) atomicAdd(io->counters + 1, 1);
  cudaGraphSetConditional(loop, it * K < st->count ? 1u : 0u);
}


// ---------------------------------------------------------------------------------------------
// Range check of a stored fp16 tensor (wv_net_set_range_check): every fp32 -> fp16 conversion of the path saturates at
// +-65504 (cvt.rn.satfinite) and the packed half2 epilogue arithmetic can overflow to inf, both silently.  With the check
// enabled every launch's fp16 outputs are scanned: stats[0] += values at the saturation bound, stats[1] += non-finite
// values, stats[2] = max |v| (fp16 bits; order-preserving for non-negative halves).
__global__ void __launch_bounds__(256)
range_check_kernel(const uint4* __restrict__ p, long long n16, unsigned long long* __restrict__ stats) {
  unsigned sat = 0, bad = 0, mx = 0;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n16; i += static_cast<long long>(gridDim.x) * 256) {
    const uint4 u = __ldcg(p + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const unsigned a = (w[k] >> (16 * hh)) & 0x7fffu;
        sat += a == 0x7bffu;
        bad += a >= 0x7c00u;
        mx = max(mx, a < 0x7c00u ? a : 0u);
      }
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    sat += __shfl_xor_sync(0xffffffffu, sat, d);
    bad += __shfl_xor_sync(0xffffffffu, bad, d);
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  }
  if ((threadIdx.x & 31) == 0) {
    if (sat) atomicAdd(stats, static_cast<unsigned long long>(sat));
    if (bad) atomicAdd(stats + 1, static_cast<unsigned long long>(bad));
    atomicMax(stats + 2, static_cast<unsigned long long>(mx));
  }
}

}  // namespace wv
