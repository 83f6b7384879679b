// Fused SEANet residual block for sm_100a (modules/seanet.py:245-281):
//   a  = ELU(x * pre_scale)                      (the activated stream the producing launch already wrote: TMA operand)
//   h  = ELU(dw5(W1 a) + b1)                     (stays in shared memory as the second GEMM's operand)
//   x' = RS * (dw5(W2 h) + b2) + x               (RS folded into the taps / bias; x = the raw stream)
//   out_raw = x' ;  out_act = ELU(x' * act_scale)        (each optional)
// HBM traffic: read a and x once, write the requested outputs once: 4C bytes per element and resblock
// instead of the 6C of the two-launch form (h1: read a, write h; out: read h and x, write x' and a').
// (Round 1 activated x in shared memory instead of reading a - 2C bytes - but that extra pass over the
// tile made the math warps, not HBM, the bound: 290 us against 236 us for the two launches at C = 96.)
//
// One persistent 640-thread CTA per SM; channels-last fp16 [clip, time, C] with C <= 128.  A tile is
// 128 time rows of one clip starting 8 rows before its first output row (two causal k=5 halos):
// 120 output rows per tile; rows before the clip start are TMA zero fill.  Two tiles (slots a, b)
// are in flight so that the tensor core / drain work of one overlaps the epilogue math of the other:
//
//   warp 0       TMA producer: W1, W2 once (resident), then one x tile per tile into a ring of nx
//                buffers (num_kb k-blocks of 128 rows x 64 channels, SWIZZLE_128B)
//   warp 1       MMA issuer:   MMA1(a) MMA1(b) MMA2(a) MMA2(b) per pair; accumulators in TMEM
//                (one 256-column slot per tile of the pair, reused by both GEMMs of the tile)
//   warps 4-7    drain:        TMEM -> saturating fp16 -> staging tile of the slot
//   warps 8-19   math:         M1: taps + bias + ELU on the staged GEMM1 tile, written back into the
//                              operand buffer in the SWIZZLE_128B layout (rows before the clip start
//                              as zeros = the causal padding of the second depthwise conv);
//                              M2: taps + bias + residual + ELU -> global; the residual rows of a
//                              thread's first two units are requested before the hand-off barrier
#pragma once
#include "gemm_sm100.cuh"

namespace wv {

#ifdef WV_TIMELINE
#define RB_DBG(slot, pair) do { if (g.dbg != nullptr && blockIdx.x == 0 && (pair) < 24) g.dbg[(pair) * 32 + (slot)] = clock64(); } while (0)
#else
#define RB_DBG(slot, pair) do { } while (0)
#endif

constexpr int RB_HALO = 8;
constexpr int RB_ROWS_OUT = BM - RB_HALO;   // 120
constexpr int RB_MAX_NX = 4;
constexpr int RB_BAR_BYTES = 512;
constexpr int RB_SLOT_COLS = 256;
constexpr int BAR_RB_FULL = 6;              // named barriers 6,7: staging tile of slot s written
constexpr int BAR_RB_EMPTY = 8;             // 8,9: staging tile of slot s free for the next pair
// Measured at C = 96 (64 x 16 000 rows): 12 math warps / 640 threads 280 us, 24 math warps / 1024
// threads at 64 registers 313 us (spills), the two-launch form 265 us.  The kernel executes ~2x the
// instructions of the two launches it replaces (three ELU passes, the in-place activation of the
// padded tile, halo rows twice) and ends up issue-bound; see DESIGN.md.
constexpr int RB_MATH_WARPS = 12;
constexpr int RB_MATH_THREADS = RB_MATH_WARPS * 32;
constexpr int RB_EPI_THREADS = (P1_WARPS + RB_MATH_WARPS) * 32;
constexpr int RB_THREADS = 128 + RB_EPI_THREADS;

struct ResblockArgs {
  int C, num_kb;
  int T, n_clips;
  int tiles_m_per_clip, num_tiles;
  uint32_t magic_m;
  int nx;                       // x tile buffers (2..4)
  uint32_t idesc;
  float pre_scale;
  const float* dw1_w;           // [5][C]
  const float* dw1_b;           // [C]
  const float* dw2_w;           // [5][C], RS * res_scale_param folded
  const float* dw2_b;           // [C], same
  const act_t* x;               // [n_clips, T, C]
  act_t* out_raw;
  act_t* out_act;
  float act_scale;
  long long* dbg;               // -DWV_TIMELINE builds: per-pair clock probes of CTA 0 (scripts/timeline_rb.py)
};

__host__ __device__ inline int rb_xtile_bytes(int num_kb) { return num_kb * A_STAGE_BYTES; }
__host__ __device__ inline int rb_w_bytes(int C, int num_kb) { return num_kb * C * BK * 2; }
__host__ __device__ inline int rb_taps_bytes(int C) { return 2 * 6 * C * 2; }   // [conv][5 taps + bias][C] fp16
__host__ inline int rb_smem_bytes(int C, int num_kb, int nx) {
  return 1024 + nx * rb_xtile_bytes(num_kb) + 2 * rb_w_bytes(C, num_kb) + RB_BAR_BYTES +
         2 * BM * staged_pitch_bytes(C) + rb_taps_bytes(C);
}
__host__ inline int rb_pick_nx(int C, int num_kb) {
  for (int nx = RB_MAX_NX; nx >= 2; --nx)
    if (rb_smem_bytes(C, num_kb, nx) <= GEMM_SMEM_LIMIT) return nx;
  return 0;
}

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts_u2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

// M1 unit: 4 output rows x 4 channels of h = ELU(dw5(S) + b1); tile rows jo..jo+3 (jo = ro + 4), written
// into the operand buffer (SWIZZLE_128B, K-major).  Rows j < zero_rows are written as zeros.
__device__ __forceinline__ void rb_unit_m1(uint32_t srow, int pitch, uint32_t xa_u32, int jo, int c, int zero_rows,
                                           const __half2 (&wt)[5][2], const __half2 (&bs)[2]) {
  __half2 x[8][2];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint2 u = lds_u2(srow + j * pitch);
    x[j][0] = as_h2(u.x);
    x[j][1] = as_h2(u.y);
  }
  const uint32_t kb_off = static_cast<uint32_t>(c >> 6) * A_STAGE_BYTES;
  const int cc = c & 63;
  const uint32_t chunk = static_cast<uint32_t>(cc >> 3), half_off = static_cast<uint32_t>((cc >> 2) & 1) * 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half2 a0 = bs[0], a1 = bs[1];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      a0 = __hfma2(wt[j][0], x[i + j][0], a0);
      a1 = __hfma2(wt[j][1], x[i + j][1], a1);
    }
    a0 = elu_h2(a0);
    a1 = elu_h2(a1);
    const int row = jo + i;
    if (row < zero_rows) { a0 = h2_from(0.f, 0.f); a1 = a0; }
    const uint32_t addr = xa_u32 + kb_off + static_cast<uint32_t>(row) * 128u + ((chunk ^ (static_cast<uint32_t>(row) & 7u)) << 4) + half_off;
    sts_u2(addr, as_u32(a0), as_u32(a1));
  }
}

// M2 over one tile.  A thread owns up to three 4-row units (row groups grp0, grp0 + gstride, grp0 + 2 gstride); the
// residual rows of the first two arrive in r0 / r1 (requested before the hand-off barrier), those of a third unit are
// requested while the first is being computed.
__device__ __forceinline__ void rb_load_res(const ResblockArgs& g, size_t base, int oo, int rows_left, uint2 (&r)[4]) {
  const char* rp = reinterpret_cast<const char*>(g.x + base + static_cast<size_t>(oo) * g.C);
  const size_t row_bytes = static_cast<size_t>(g.C) * 2;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    r[i] = oo + i < rows_left ? __ldcg(reinterpret_cast<const uint2*>(rp + i * row_bytes)) : make_uint2(0u, 0u);
}
template <bool RAW, bool ACT>
__device__ __forceinline__ void rb_m2_unit(const ResblockArgs& g, const GemmArgs& gf, uint32_t tile_u32, int pitch, size_t base,
                                           int oo, int rows_left, const __half2 (&wt)[5][2], const __half2 (&bs)[2],
                                           const uint2 (&rres)[4]) {
  const size_t row_bytes = static_cast<size_t>(g.C) * 2;
  const size_t off = base + static_cast<size_t>(oo) * g.C;
  const uint32_t srow = tile_u32 + static_cast<uint32_t>(oo + 4) * pitch;   // staged rows oo+4 .. oo+11
  if (oo + 4 <= rows_left)
    staged_unit<5, 4, true, RAW, ACT, true, true>(gf, srow, pitch, off, row_bytes, 4, wt, bs, g.act_scale, rres);
  else
    staged_unit<5, 4, true, RAW, ACT, false, true>(gf, srow, pitch, off, row_bytes, rows_left - oo, wt, bs, g.act_scale, rres);
}
template <bool RAW, bool ACT>
__device__ __forceinline__ void rb_m2_tile(const ResblockArgs& g, const GemmArgs& gf, uint32_t tile_u32, int pitch,
                                           int grp0, int gstride, bool active, size_t base, int rows_left,
                                           const __half2 (&wt)[5][2], const __half2 (&bs)[2], uint2 (&r0)[4], uint2 (&r1)[4]) {
  if (!active) return;
  constexpr int NG = RB_ROWS_OUT / 4;
  const int o0 = grp0 * 4, o1 = (grp0 + gstride) * 4, o2 = (grp0 + 2 * gstride) * 4;
  const bool h0 = grp0 < NG && o0 < rows_left, h1 = grp0 + gstride < NG && o1 < rows_left, h2 = grp0 + 2 * gstride < NG && o2 < rows_left;
  if (!h0) return;
  rb_m2_unit<RAW, ACT>(g, gf, tile_u32, pitch, base, o0, rows_left, wt, bs, r0);
  if (h2) rb_load_res(g, base, o2, rows_left, r0);
  if (h1) rb_m2_unit<RAW, ACT>(g, gf, tile_u32, pitch, base, o1, rows_left, wt, bs, r1);
  if (h2) rb_m2_unit<RAW, ACT>(g, gf, tile_u32, pitch, base, o2, rows_left, wt, bs, r0);
}

__global__ void __launch_bounds__(RB_THREADS, 1)
resblock_sm100_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                      const __grid_constant__ CUtensorMap tmW2, const ResblockArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int xt_bytes = rb_xtile_bytes(g.num_kb);
  const int w_bytes = rb_w_bytes(g.C, g.num_kb);
  const int wkb_bytes = g.C * BK * 2;
  uint8_t* xa = smem;                                  // [nx][num_kb][128][64] fp16
  uint8_t* w1 = xa + g.nx * xt_bytes;                  // [num_kb][C][64]
  uint8_t* w2 = w1 + w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w2 + w_bytes);
  uint64_t* x_full = bars;                             // [RB_MAX_NX]
  uint64_t* x_empty = x_full + RB_MAX_NX;              // [RB_MAX_NX]
  uint64_t* acc_full = x_empty + RB_MAX_NX;            // [2]
  uint64_t* tm_empty = acc_full + 2;                   // [2]
  uint64_t* t0_done = tm_empty + 2;                    // [2]
  uint64_t* m1_done = t0_done + 2;                     // [2]
  uint64_t* w_full = m1_done + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  uint8_t* stage_tiles = reinterpret_cast<uint8_t*>(bars) + RB_BAR_BYTES;   // [2][128][pitch]
  const int pitch = staged_pitch_bytes(g.C);
  uint8_t* taps_tbl = stage_tiles + 2 * BM * pitch;    // [2][6][C] fp16: depthwise taps + bias of both convs

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int my_tiles = (g.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < RB_MAX_NX; ++i) {
      mbar_init(&x_full[i], 1);
      mbar_init(&x_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&acc_full[i], 1);
      mbar_init(&tm_empty[i], P1_WARPS);
      mbar_init(&t0_done[i], RB_MATH_WARPS);
      mbar_init(&m1_done[i], RB_MATH_WARPS);
    }
    mbar_init(w_full, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 2 * 6 * g.C; i += blockDim.x) {   // the depthwise taps live in shared memory: a phase
    const int conv = i / (6 * g.C), r = (i / g.C) % 6, ch = i % g.C;  // loads its 6 half2 pairs when it starts
    const float* wsrc = conv ? g.dw2_w : g.dw1_w;
    const float* bsrc = conv ? g.dw2_b : g.dw1_b;
    reinterpret_cast<__half*>(taps_tbl)[i] = __float2half_rn(r < 5 ? __ldg(wsrc + r * g.C + ch) : __ldg(bsrc + ch));
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_launch_dependents();

  auto tile_rc = [&](int it, int& clip, int& mi) {    // it-th tile of this CTA -> (clip, m tile)
    const uint32_t tile = blockIdx.x + static_cast<uint32_t>(it) * gridDim.x;
    if (g.tiles_m_per_clip == 1) { clip = static_cast<int>(tile); mi = 0; }
    else fast_divmod(tile, static_cast<uint32_t>(g.tiles_m_per_clip), g.magic_m, clip, mi);
  };

  if (warp < 4) reg_dealloc<REGS_LIGHT>();           // same split as the GEMM: TMA / MMA warpgroup 40, drain 96, math 112
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(w_full, static_cast<uint32_t>(2 * w_bytes));
      for (int kb = 0; kb < g.num_kb; ++kb) {
        tma_load_2d(w1 + kb * wkb_bytes, &tmW1, w_full, kb * BK, 0);
        tma_load_2d(w2 + kb * wkb_bytes, &tmW2, w_full, kb * BK, 0);
      }
      for (int it = 0; it < my_tiles; ++it) {
        const int buf = it % g.nx;
        const uint32_t ph = static_cast<uint32_t>(it / g.nx) & 1u;
        int clip, mi;
        tile_rc(it, clip, mi);
        mbar_wait(&x_empty[buf], ph ^ 1);
        mbar_arrive_expect_tx(&x_full[buf], static_cast<uint32_t>(xt_bytes));
        for (int kb = 0; kb < g.num_kb; ++kb)
          tma_load_3d(xa + buf * xt_bytes + kb * A_STAGE_BYTES, &tmX, &x_full[buf], kb * BK, mi * RB_ROWS_OUT - RB_HALO, clip);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      mbar_wait(w_full, 0);
      auto issue = [&](uint8_t* a_tile, uint8_t* w_tile, uint32_t tmem_d) {
        for (int kb = 0; kb < g.num_kb; ++kb) {
          const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(a_tile + kb * A_STAGE_BYTES));
          const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(w_tile + kb * wkb_bytes));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, g.idesc, (kb | k) ? 1u : 0u);
        }
      };
      for (int it = 0, pair = 0; it < my_tiles; it += 2, ++pair) {
        const int nb = my_tiles - it < 2 ? my_tiles - it : 2;
        const uint32_t pph = static_cast<uint32_t>(pair) & 1u;
        for (int s = 0; s < nb; ++s) {
          mbar_wait(&tm_empty[s], pph ^ 1);          // TMEM slot drained (second GEMM of the previous pair)
          mbar_wait(&x_full[(it + s) % g.nx], static_cast<uint32_t>((it + s) / g.nx) & 1u);   // activated tile landed
          tc_fence_after();
          RB_DBG(0 + s, pair);
          issue(xa + ((it + s) % g.nx) * xt_bytes, w1, tmem_base + static_cast<uint32_t>(s * RB_SLOT_COLS));
          umma_commit(&acc_full[s]);
        }
        for (int s = 0; s < nb; ++s) {
          mbar_wait(&m1_done[s], pph);               // h written into the x buffer (operand layout)
          tc_fence_after();
          RB_DBG(2 + s, pair);
          issue(xa + ((it + s) % g.nx) * xt_bytes, w2, tmem_base + static_cast<uint32_t>(s * RB_SLOT_COLS));
          umma_commit(&acc_full[s]);
          umma_commit(&x_empty[(it + s) % g.nx]);    // the x buffer may be refilled once GEMM2 has read it
        }
      }
    }
  } else if (warp >= 4 && warp < 4 + P1_WARPS) {
    reg_dealloc<REGS_DRAIN>();                       // (the math warps' 120 registers come out of this warpgroup's share)
    // ------------------------------------------------------------ drain warps: TMEM -> fp16 -> staging[slot]
    const int q = warp - 4;
    const int chunks = g.C / 32;
    const uint32_t stage_u32 = smem_u32(stage_tiles);
    uint32_t acc_phase[2] = {0u, 0u};
    uint32_t v[32];
    auto drain = [&](int s) {
      mbar_wait(&acc_full[s], acc_phase[s]);
      acc_phase[s] ^= 1u;
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(s * RB_SLOT_COLS);
      const uint32_t rowp = stage_u32 + s * (BM * pitch) + (q * 32 + lane) * pitch;
      for (int c = 0; c < chunks; ++c) {
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 4; ++i)
          sts_u4(rowp + c * 64 + i * 16,
                 pack_act2(__uint_as_float(v[8 * i]), __uint_as_float(v[8 * i + 1])),
                 pack_act2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3])),
                 pack_act2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5])),
                 pack_act2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7])));
      }
      tc_fence_before();
      __syncwarp();
    };
    for (int it = 0, pair = 0; it < my_tiles; it += 2, ++pair) {
      const int nb = my_tiles - it < 2 ? my_tiles - it : 2;
      for (int s = 0; s < nb; ++s) {                 // GEMM1 tiles
        if (pair > 0) named_bar_sync(BAR_RB_EMPTY + s, RB_EPI_THREADS);   // M2 of the previous pair left staging[s]
        if (q == 0 && lane == 0) RB_DBG(4 + s, pair);
        drain(s);
        if (q == 0 && lane == 0) RB_DBG(6 + s, pair);
        named_bar_arrive(BAR_RB_FULL + s, RB_EPI_THREADS);
      }
      for (int s = 0; s < nb; ++s) {                 // GEMM2 tiles (staging[s] was released by M1: GEMM2 waited for it)
        if (q == 0 && lane == 0) RB_DBG(8 + s, pair);
        drain(s);
        if (q == 0 && lane == 0) RB_DBG(10 + s, pair);
        if (lane == 0) mbar_arrive(&tm_empty[s]);
        named_bar_arrive(BAR_RB_FULL + s, RB_EPI_THREADS);
      }
    }
  } else if (warp >= 4 + P1_WARPS) {
    // ------------------------------------------------------------ math warps
    reg_alloc<REGS_MATH>();
    const int et = threadIdx.x - (128 + P1_WARPS * 32);
    const int cgs = g.C >> 2;
    const int gstride = RB_MATH_THREADS / cgs;
    const int cg = et % cgs, grp0 = et / cgs;
    const bool active = grp0 < gstride;
    const int c = cg * 4;
    const uint32_t stage_u32 = smem_u32(stage_tiles) + cg * 8;
    const uint32_t xa_u32 = smem_u32(xa);
    const uint32_t taps_u32 = smem_u32(taps_tbl) + c * 2;
    auto load_taps = [&](int conv, __half2 (&wt)[5][2], __half2 (&bs)[2]) {
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const uint2 u = lds_u2(taps_u32 + (conv * 6 + j) * g.C * 2);
        wt[j][0] = as_h2(u.x); wt[j][1] = as_h2(u.y);
      }
      const uint2 u = lds_u2(taps_u32 + (conv * 6 + 5) * g.C * 2);
      bs[0] = as_h2(u.x); bs[1] = as_h2(u.y);
    };
    GemmArgs gf = {};                                // the output pointers staged_unit reads
    gf.residual = g.x; gf.out_raw = g.out_raw; gf.out_act = g.out_act;
    for (int it = 0, pair = 0; it < my_tiles; it += 2, ++pair) {
      const int nb = my_tiles - it < 2 ? my_tiles - it : 2;
      // ---- M1: h = ELU(dw5(S1) + b1) -> operand layout in the x buffer
      for (int s = 0; s < nb; ++s) {
        int clip, mi;
        tile_rc(it + s, clip, mi);
        const int zero_rows = RB_HALO - mi * RB_ROWS_OUT;   // tile rows before the clip start (8 for mi = 0, else <= 0)
        named_bar_sync(BAR_RB_FULL + s, RB_EPI_THREADS);
        if (et == 0) RB_DBG(18 + s, pair);
        if (active) {
          __half2 wt1[5][2], bs1[2];
          load_taps(0, wt1, bs1);
          const uint32_t tile_u32 = stage_u32 + s * (BM * pitch);
          const uint32_t xbuf = xa_u32 + ((it + s) % g.nx) * xt_bytes;
          for (int grp = grp0; grp < (BM - 4) / 4; grp += gstride)
            rb_unit_m1(tile_u32 + grp * 4 * pitch, pitch, xbuf, grp * 4 + 4, c, zero_rows, wt1, bs1);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&m1_done[s]);
        if (et == 0) RB_DBG(20 + s, pair);
      }
      // ---- M2: x' = dw5(S2) + b2 + x -> global
      for (int s = 0; s < nb; ++s) {
        int clip, mi;
        tile_rc(it + s, clip, mi);
        const int r_base = mi * RB_ROWS_OUT;
        const int rows_left = min(g.T - r_base, RB_ROWS_OUT);
        const size_t base = (static_cast<size_t>(clip) * g.T + r_base) * g.C + c;
        uint2 r0[4], r1[4];
        if (active) {                                // residual rows do not depend on the staged tile: in flight across the wait
          if (grp0 * 4 < rows_left) rb_load_res(g, base, grp0 * 4, rows_left, r0);
          if (grp0 + gstride < RB_ROWS_OUT / 4 && (grp0 + gstride) * 4 < rows_left) rb_load_res(g, base, (grp0 + gstride) * 4, rows_left, r1);
        }
        named_bar_sync(BAR_RB_FULL + s, RB_EPI_THREADS);
        if (et == 0) RB_DBG(22 + s, pair);
        const uint32_t tile_u32 = stage_u32 + s * (BM * pitch);
        __half2 wt2[5][2], bs2[2];
        if (active) load_taps(1, wt2, bs2);
        if (g.out_raw != nullptr && g.out_act != nullptr)
          rb_m2_tile<true, true>(g, gf, tile_u32, pitch, grp0, gstride, active, base, rows_left, wt2, bs2, r0, r1);
        else if (g.out_raw != nullptr)
          rb_m2_tile<true, false>(g, gf, tile_u32, pitch, grp0, gstride, active, base, rows_left, wt2, bs2, r0, r1);
        else
          rb_m2_tile<false, true>(g, gf, tile_u32, pitch, grp0, gstride, active, base, rows_left, wt2, bs2, r0, r1);
        __syncwarp();
        if (et == 0) RB_DBG(24 + s, pair);
        if (it + 2 + s < my_tiles) named_bar_arrive(BAR_RB_EMPTY + s, RB_EPI_THREADS);   // staging[s] free for the next pair
      }
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
