// Pointwise-conv / STFT / head GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T  (16-bit in, fp32
// accumulate in TMEM), persistent + warp specialised:
//   warp 0     TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B, 4-stage mbarrier ring)
//   warp 1     MMA issuer     (tcgen05.mma cta_group::1, M=128, N=block_n, K=16 per instr)
//   warp 2     TMEM allocator (512 columns = 2 accumulator stages x 256)
//   warps 4-7  epilogue       (tcgen05.ld 32x32b -> registers -> fused epilogue -> global)
// Activations are channels-last [clip, time, channel] so "time" is the MMA M dimension and the
// channel contraction is K-major for both operands.  A is addressed through a 3-D tensor map
// (k, row-in-clip, clip) so the same kernel serves the flattened [B*T, C] activations
// (n_clips = 1) and the strided, overlapping STFT frame view of the padded waveform.
//
// Epilogues (template EPI):
//   STD     v = acc + bias[n] + residual[m,n];  out_raw = bf16(v);  out_act = bf16(ELU(v*s))
//   L2NORM  v = acc + bias;  v *= scale / max(||v||_2 over N, 1e-12)      (modules/seanet.py:288)
//   STFT    columns are (re,im) pairs: y = (0.5*ln(max(re^2+im^2, c)) - mu) / sigma
//           (modules/conv.py:1076 + modules/seanet.py:482-494); pair 0 carries the two purely
//           real bins k=0 and k=N/2
//   HEAD    ConvTranspose1d(k=s=hop) folded with the 1x1 last_layer (model/detector.py:304-310):
//           column n = o*hop + j -> logit[b, o, f*hop + j]; fused mask / sigmoid / partial
//           sigmoid sums for the bit decode
#pragma once
#include "ptx_sm100.cuh"

namespace wv {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int STAGES = 4;
constexpr int MAX_BN = 256;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int B_STAGE_BYTES = MAX_BN * BK * 2;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = 512;
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align*/ + 256;

enum { EPI_STD = 0, EPI_L2NORM = 1, EPI_STFT = 2, EPI_HEAD = 3 };

struct GemmArgs {
  int rows_per_clip;  // rows of A per clip (flat: total M)
  int n_clips;
  int N, K;
  int block_n;
  uint32_t idesc;
  // STD / L2NORM
  const float* bias;
  const __nv_bfloat16* residual;
  __nv_bfloat16* out_raw;
  __nv_bfloat16* out_act;
  float act_scale;
  int ldo;
  float l2_scale;
  float* out_f32_t;  // [n_clips', N, F] fp32 (latent for the API), nullable
  int f32_F;         // frames per clip for out_f32_t (flat A: clip = m / f32_F)
  // STFT
  float log_offset, inv_sigma, clamp_sq;
  int n_half;
  // HEAD
  float* logits;
  uint8_t* mask_out;
  float* probs;
  float* partial;  // [M, N / block_n]
  const uint8_t* presence;
  int hop, T, n_out, head_F;
};

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + STAGES * A_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES));
  uint64_t* full = bars;                      // [STAGES]
  uint64_t* empty = bars + STAGES;            // [STAGES]
  uint64_t* acc_full = bars + 2 * STAGES;     // [ACC_STAGES]
  uint64_t* acc_empty = acc_full + ACC_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + ACC_STAGES);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m_per_clip = (g.rows_per_clip + BM - 1) / BM;
  const int tiles_m = tiles_m_per_clip * g.n_clips;
  const int tiles_n = g.N / g.block_n;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (g.K + BK - 1) / BK;
  const uint32_t stage_bytes = static_cast<uint32_t>((BM + g.block_n) * BK * 2);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&acc_empty[i], 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / tiles_n, nt = tile % tiles_n;
        const int clip = mt / tiles_m_per_clip;
        const int r0 = (mt % tiles_m_per_clip) * BM;
        const int n0 = nt * g.block_n;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], stage_bytes);
          tma_load_3d(smemA + stage * A_STAGE_BYTES, &tmA, &full[stage], kb * BK, r0, clip);
          tma_load_2d(smemB + stage * B_STAGE_BYTES, &tmB, &full[stage], kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t as_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      mbar_wait(&acc_empty[as], as_phase ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * MAX_BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smemA + stage * A_STAGE_BYTES));
          const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smemB + stage * B_STAGE_BYTES));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 B per K=16 step inside the 128 B swizzle row -> +2 in the (addr>>4) field
            umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, g.idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (kb == num_kb - 1) umma_commit(&acc_full[as]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++as == ACC_STAGES) { as = 0; as_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue
    const int q = warp - 4;  // TMEM lane quarter: this warp may touch lanes [32q, 32q+32)
    int as = 0;
    uint32_t as_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int mt = tile / tiles_n, nt = tile % tiles_n;
      const int clip = mt / tiles_m_per_clip;
      const int r = (mt % tiles_m_per_clip) * BM + q * 32 + lane;
      const bool row_ok = r < g.rows_per_clip;
      const long long m = static_cast<long long>(clip) * g.rows_per_clip + r;
      const int n0 = nt * g.block_n;
      const int chunks = g.block_n / 32;
      mbar_wait(&acc_full[as], as_phase);
      tc_fence_after();
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * MAX_BN);
      uint32_t v[32];

      if constexpr (EPI == EPI_STD) {
        for (int c = 0; c < chunks; ++c) {
          const int n = n0 + c * 32;
          uint4 rres[4];
          if (g.residual != nullptr && row_ok) {
            const uint4* rp = reinterpret_cast<const uint4*>(g.residual + m * g.ldo + n);
#pragma unroll
            for (int i = 0; i < 4; ++i) rres[i] = __ldg(rp + i);
          }
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
            if (g.bias != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] += __ldg(g.bias + n + i);
            }
            if (g.residual != nullptr) {
              const uint32_t* rw = reinterpret_cast<const uint32_t*>(rres);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float a, b;
                unpack_bf16x2(rw[i], a, b);
                f[2 * i] += a;
                f[2 * i + 1] += b;
              }
            }
            if (g.out_raw != nullptr) {
              uint4* op = reinterpret_cast<uint4*>(g.out_raw + m * g.ldo + n);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                op[i] = make_uint4(pack_bf16x2(f[8 * i], f[8 * i + 1]),
                                   pack_bf16x2(f[8 * i + 2], f[8 * i + 3]),
                                   pack_bf16x2(f[8 * i + 4], f[8 * i + 5]),
                                   pack_bf16x2(f[8 * i + 6], f[8 * i + 7]));
            }
            if (g.out_act != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = elu1(f[i] * g.act_scale);
              uint4* op = reinterpret_cast<uint4*>(g.out_act + m * g.ldo + n);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                op[i] = make_uint4(pack_bf16x2(f[8 * i], f[8 * i + 1]),
                                   pack_bf16x2(f[8 * i + 2], f[8 * i + 3]),
                                   pack_bf16x2(f[8 * i + 4], f[8 * i + 5]),
                                   pack_bf16x2(f[8 * i + 6], f[8 * i + 7]));
            }
          }
        }
      } else if constexpr (EPI == EPI_L2NORM) {
        float ss = 0.f;
        for (int c = 0; c < chunks; ++c) {
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x = __uint_as_float(v[i]) + (g.bias ? __ldg(g.bias + n0 + c * 32 + i) : 0.f);
            ss += x * x;
          }
        }
        const float sc = g.l2_scale / fmaxf(sqrtf(ss), 1e-12f);
        for (int c = 0; c < chunks; ++c) {
          const int n = n0 + c * 32;
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i)
              f[i] = (__uint_as_float(v[i]) + (g.bias ? __ldg(g.bias + n + i) : 0.f)) * sc;
            uint4* op = reinterpret_cast<uint4*>(g.out_raw + m * g.ldo + n);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              op[i] = make_uint4(pack_bf16x2(f[8 * i], f[8 * i + 1]),
                                 pack_bf16x2(f[8 * i + 2], f[8 * i + 3]),
                                 pack_bf16x2(f[8 * i + 4], f[8 * i + 5]),
                                 pack_bf16x2(f[8 * i + 6], f[8 * i + 7]));
            if (g.out_f32_t != nullptr) {
              const long long b = m / g.f32_F, fr = m % g.f32_F;
              float* tp = g.out_f32_t + (b * g.N + n) * g.f32_F + fr;
#pragma unroll
              for (int i = 0; i < 32; ++i) tp[static_cast<long long>(i) * g.f32_F] = f[i];
            }
          }
        }
      } else if constexpr (EPI == EPI_STFT) {
        for (int c = 0; c < chunks; ++c) {
          const int p0 = (n0 + c * 32) >> 1;
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            float y[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float re = __uint_as_float(v[2 * i]), im = __uint_as_float(v[2 * i + 1]);
              y[i] = (0.5f * logf(fmaxf(re * re + im * im, g.clamp_sq)) - g.log_offset) * g.inv_sigma;
            }
            __nv_bfloat16* yp = g.out_raw + m * g.ldo;
            if (p0 == 0) {
              const float re0 = __uint_as_float(v[0]), ren = __uint_as_float(v[1]);
              y[0] = (0.5f * logf(fmaxf(re0 * re0, g.clamp_sq)) - g.log_offset) * g.inv_sigma;
              const float yn = (0.5f * logf(fmaxf(ren * ren, g.clamp_sq)) - g.log_offset) * g.inv_sigma;
              yp[g.n_half] = __float2bfloat16_rn(yn);
            }
            uint4* op = reinterpret_cast<uint4*>(yp + p0);
#pragma unroll
            for (int i = 0; i < 2; ++i)
              op[i] = make_uint4(pack_bf16x2(y[8 * i], y[8 * i + 1]),
                                 pack_bf16x2(y[8 * i + 2], y[8 * i + 3]),
                                 pack_bf16x2(y[8 * i + 4], y[8 * i + 5]),
                                 pack_bf16x2(y[8 * i + 6], y[8 * i + 7]));
          }
        }
      } else {  // EPI_HEAD
        const long long b = m / g.head_F;
        const int fr = static_cast<int>(m % g.head_F);
        const int o = n0 / g.hop;
        const int j0 = n0 % g.hop;
        const float bo = g.bias ? __ldg(g.bias + o) : 0.f;
        float psum = 0.f;
        for (int c = 0; c < chunks; ++c) {
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          if (row_ok) {
            const int t0 = fr * g.hop + j0 + c * 32;
            const long long base = b * g.T + t0;
            float* lp = g.logits ? g.logits + (b * g.n_out + o) * static_cast<long long>(g.T) + t0
                                 : nullptr;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (t0 + i < g.T) {
                const float l = __uint_as_float(v[i]) + bo;
                const float p = 1.f / (1.f + __expf(-l));
                if (lp) lp[i] = l;
                if (g.mask_out) g.mask_out[base + i] = l > 0.5f ? 1 : 0;
                if (g.probs) g.probs[base + i] = p;
                if (g.partial) psum += g.presence ? (g.presence[base + i] ? p : 0.f) : p;
              }
            }
          }
        }
        if (g.partial != nullptr && row_ok) g.partial[m * tiles_n + nt] = psum;
      }

      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      if (++as == ACC_STAGES) { as = 0; as_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace wv
