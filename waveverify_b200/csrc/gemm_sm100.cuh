// Pointwise-conv / STFT / head GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T  (16-bit in, fp32
// accumulate in TMEM), persistent + warp specialised (640 threads, 1 CTA per SM):
//   warp 0      TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B, mbarrier ring of 3-6 stages)
//   warp 1      MMA issuer     (tcgen05.mma cta_group::1, M=128, N=block_n, K=16 per instr)
//   warp 2      TMEM allocator (512 columns = 2 accumulator stages x 256)
//   warps 4-19  epilogue       (tcgen05.ld 32x32b -> registers -> ... -> global)
// Activations are channels-last [clip, time, channel] so "time" is the MMA M dimension and the
// channel contraction is K-major for both operands.  A is addressed through a 3-D tensor map
// (k, row-in-clip, clip): flattened [B*T, C] activations use n_clips = 1; per-clip tiles (with
// negative / overlapping row origins) serve the causal depthwise halo and the strided,
// overlapping STFT frame view of the padded waveform.
//
// Epilogues (template EPI):
//   STAGED  warps 4-7 ("drain") round the fp32 accumulator tile to fp16 (11-bit mantissa,
//           saturating) into one of two padded shared-memory tiles and release the TMEM stage;
//           warps 8-19 ("math", 112 registers via setmaxnreg) walk the staged tile in a coalesced
//           (4 rows x 4 channels) layout.  Taps, bias, residual add and ELU all run on packed half2
//           pairs (two channels per instruction; fp16 activations are loaded / stored as they are):
//             v = bias[c] + sum_{j<taps} w[j][c] * S[r-taps+1+j][c]      taps = 1 or 5
//             v += residual[m,c];  out_raw = v;  out_act = ELU(v*s)        (all packed half2)
//           taps = 5 fuses the causal depthwise conv that follows every resblock 1x1
//           (modules/seanet.py:85-109): tiles overlap by 4 rows (128 rows in, 124 out).
//           The math loop is compiled per (taps, residual, raw, act) combination so the hot loop
//           carries no runtime feature tests.
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

// Per-tile clock probes of CTA 0 (scripts/timeline.py); compiled in only with -DWV_TIMELINE.
#ifdef WV_TIMELINE
#define WV_DBG_MODE(m) ((g.dbg_mode & (m)) != 0)   // bit 1: no global stores, 2: no units, 4: no drain
#define WV_DBG(slot, it) do { if (g.dbg != nullptr && blockIdx.x == 0 && (it) < 48) g.dbg[(it) * 40 + (slot)] = clock64(); } while (0)
#else
#define WV_DBG_MODE(m) false
#define WV_DBG(slot, it) do { } while (0)
#endif

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int MAX_STAGES = 8;
constexpr int MAX_BN = 256;
constexpr int A_STAGE_BYTES = BM * BK * 2;
constexpr int ACC_STAGES = 2;               // accumulator stages of MAX_BN columns; pair mode: 4 stages of 128 columns
constexpr int MAX_ACC_STAGES = 4;
constexpr int TMEM_COLS = 512;
constexpr int EPI_WARPS = 16;               // L2NORM / STFT / HEAD epilogues (640-thread CTA)
constexpr int EPI_SPLIT = EPI_WARPS / 4;    // warps sharing one TMEM lane quarter split the columns
constexpr int GEMM_THREADS = 128 + EPI_WARPS * 32;
// STAGED: 4 drain + 12 math warps.  (A 1024-thread CTA with 24 math warps at 64 registers was
// measured: the math phase shortens but the drain slows by as much; no gain.)
constexpr int P1_WARPS = 4;                 // TMEM -> smem drain warps (one per lane quarter)
constexpr int P2_WARPS = 12;                // smem -> math -> global warps
constexpr int P2_THREADS = P2_WARPS * 32;
constexpr int EPI_THREADS = (P1_WARPS + P2_WARPS) * 32;   // participants of the drain <-> math barriers
constexpr int STAGED_THREADS = 128 + EPI_THREADS;
template <int EPI> constexpr int gemm_threads() { return (EPI == 0 || EPI == 4) ? STAGED_THREADS : GEMM_THREADS; }
constexpr int STAGE_BUFS = 2;               // staging tiles (drain of tile i+1 overlaps math of tile i)
constexpr int STAGED_MAX_BN = 128;
constexpr int BAR_TAPS = 1;                // named barrier ids: 1 = down-conv tap staging,
constexpr int BAR_ST_FULL = 2;             // 2,3 = staging tile written (drain arrive, math sync),
constexpr int BAR_ST_EMPTY = 4;            // 4,5 = staging tile consumed (math arrive, drain sync)
constexpr int GEMM_SMEM_LIMIT = 227 * 1024;
constexpr int GEMM_BAR_BYTES = 512;
constexpr int RES_BAR_INDEX = 32;           // mbarriers [32, 36) of the barrier block: residual tile full[2] / empty[2]
constexpr int DOWN_W_BYTES = 2 * 16 * STAGED_MAX_BN * 2;   // staged down-conv taps [half][2r <= 16][block_n] fp16 (CTA-pair mode: both halves of the wide n tile)

// Epilogues 4..6 are the PRECISE variants (fp32-accurate path for the thresholded outputs, see the
// "precise mode" block below): split-fp16 operands (hi + lo) on the tensor cores, fp32 epilogue math.
enum { EPI_STAGED = 0, EPI_L2NORM = 1, EPI_STFT = 2, EPI_HEAD = 3, EPI_STAGED_PM = 4, EPI_L2NORM_PM = 5, EPI_STFT_PM = 6 };
constexpr int STAGED_PM_MAX_BN = 64;                       // fp32 staging tiles: 2 x 128 x (64*4+16) = 68 KB
constexpr int DOWN_W_BYTES_PM = 16 * STAGED_PM_MAX_BN * 4;  // staged down-conv taps [2r <= 16][block_n] fp32

struct GemmArgs {
  int rows_per_clip;  // rows of A per clip (flat: total M)
  int rows_per_clip_out;   // output rows per clip (== rows_per_clip unless the epilogue downsamples)
  int n_clips;
  int N, K;
  int block_n;
  int stages;         // smem ring depth (host computed from block_n)
  int stage_bufs;     // STAGED: staging tiles (2; 1 for long-K layers, whose ring gets the space instead)
  int pair;           // 1: two M tiles (same n tile) share every W k-block: one ring stage = 2 A tiles + 1 W tile,
                      //    four 128-column accumulator stages (long-K layers are bound by L2 -> SM operand traffic)
  int cg2;            // 1: CTA-pair mode (cluster of 2, tcgen05 cta_group::2): one MMA of M = 256 covers the M tiles of both CTAs
                      //    and 2 * block_n columns; each CTA loads its own A tile and HALF of the W k-block (block_n rows), so a
                      //    128 x 128 x 64 block of MACs costs 16 KB of L2 -> SM operand traffic instead of 32 KB (24 KB in pair
                      //    mode).  The 2 * block_n accumulator columns are drained as two block_n-wide tiles.
  uint32_t idesc2;    //    instruction descriptor of that MMA (M = 256, N = 2 * block_n)
  uint32_t magic_n2;  //    floor(2^32 / (tiles_n / 2))
  int acc_stages, acc_cols;
  int kb_split;       // > 0: k-blocks >= kb_split re-read the A rows shifted by one (row r-1) at k - kb_split*64:
                      //      [a[i] | a[i-1]] contraction of the fused transposed-conv + 1x1 (decoder upsample)
  int a2_split;       // > 0: k-blocks >= a2_split are read from a SECOND A tensor (tmR) at k - a2_split*64, same rows
  int a3_split;       // > 0: k-blocks >= a3_split are read from the FIRST A tensor again at k - a3_split*64 (precise mode:
                      //      [hi | lo | hi] x [Wh | Wh | Wl] contraction of the split-fp16 STFT)
  int dual;           // 1 (with a2_split): the tile's columns are [acc1 | acc2], block_n/2 channels each: acc1 = the 1x1 conv
                      //    that feeds the depthwise taps, acc2 = a second 1x1 conv over the second A tensor that is added
                      //    un-tapped (last encoder resblock + spectrogram branch in one launch, modules/seanet.py:936-943)
  int math_groups;    // STAGED: 2 = the twelve math warps form two groups of six, one per staging tile, so that the per-tile
                      //   serial part of a group (tile coordinates, barrier hand-off, last partial pass) overlaps the other group's math
  int a_prefetch;     // > 0: the producer asks L2 for the A rows of the tile this many tiles ahead (cp.async.bulk.prefetch.tensor):
                      //   DRAM -> L2 runs further ahead than the shared-memory ring can hold
  int res_early2;     // 1: the residual rows of a thread's second unit are requested before the hand-off barrier as well
  int res_tma;        // 1: the residual tile of every output tile is TMA-loaded (tmR, unswizzled [block_n x 128] box) into one of two
                      //    shared-memory buffers by warp 3, one tile ahead; the math warps read it from there instead of issuing
                      //    per-thread ld.global (W resident: the ring streams A only and can spare the 2 x 128 x block_n x 2 bytes)
  int unit_rows;      // STAGED math units: rows per unit (4, or 6 for tile widths whose 4-row groups leave the second pass mostly idle)
  int a_evict_first;  // 1: A operand loads carry the L2 evict_first hint (streamed once)
  int reverse;        // 1: walk the tiles from the last one down (L2 reuse across consecutive launches)
  int resident_b;     // 1: this CTA's W tile (all k-blocks) stays in shared memory; the ring carries A only
  uint32_t idesc;
  // tile -> (m tile, n tile, clip) without integer division (host computed)
  int tiles_n, tiles_m_per_clip, num_tiles;
  uint32_t magic_n, magic_m;   // floor(2^32 / d)
  int tile_stride;    // A rows between consecutive M tiles (128, 124 with the dw5 halo, outs*r for down)
  int tile_halo;      // rows loaded before the first row a tile produces output for
  // STAGED
  int down_r;         // > 0: fused strided depthwise down-conv k=2r, s=r (+ FiLM); taps unused
  const float* film;  // [clips, film_stride] (gamma, beta) pairs per band, nullable
  int film_stride, film_bands;
  int taps;           // 1, or 5 = fused causal depthwise conv over time
  const float* dw_w;  // [taps][N] fp32 (taps == 5)
  const float* bias;  // [N] (added after the depthwise taps), nullable
  const act_t* residual;
  act_t* out_raw;
  act_t* out_act;
  float act_scale;
  int ldo;
  // precise mode (EPI_*_PM): raw streams are fp32 [rows, ldo]; 16-bit tensors that feed a GEMM are SPLIT fp16
  // [rows, ldo_act] = [hi (lo_off columns) | lo]: v = hi + lo to ~22 bits (hi = rn16(v), lo = rn16(v - hi))
  const float* residual32;
  float* out_raw32;
  int ldo_act, lo_off;
  // STAGED, pre mode (pre_w != nullptr, residual == nullptr): the residual stream of the encoder's first resblock is the
  // output of conv_pre (1 -> C, k = 5 causal, modules/seanet.py:657-664); the math warps recompute it from five waveform
  // samples (same fp32 FMA order and fp16 rounding as conv_pre_kernel: bit-identical) instead of reading C channels
  const float* pre_w;   // [5][C] fp32, 1/wav_std folded
  const float* pre_b;   // [C]
  const float* pre_x;   // [clips, pre_T] fp32 waveform (set at run time)
  int pre_T;
  // STAGED, last_mode: the decoder's output conv C -> 1, k = 5 (modules/seanet.py:1177-1202) as a GEMM with one
  // column per tap (P[t, j] = w_j . a[t]) and out[t] = tanh(b + sum_j P[t-4+j, j]); fp32 staging; fused trim + watermark add
  int last_mode, last_T;
  float last_bias;
  const float* last_x;
  float* last_wm;
  float* last_y;
  // L2NORM
  float l2_scale;
  float* out_f32_t;  // [clips, N, F] fp32 (latent for the API), nullable
  int f32_F;         // frames per clip for out_f32_t (flat A: clip = m / f32_F)
  // STFT
  float log_offset, inv_sigma, clamp_sq;
  int n_half;
  int epi_groups;      // 2: narrow tiles (<= 64 columns): the 16 epilogue warps form two groups of 8, one per accumulator
                       //    stage, so that two tiles are in the epilogue at once (it is latency-bound per tile)
  int phases;          // > 1: frames with hop < 8 samples read as `phases` interleaved 16-byte-strided views of
  int frames_per_clip; //      shifted waveform copies (4-D tensor map); "clip" = clip * phases + phase, row q = frame q*phases + phase
  // HEAD
  float* logits;
  uint8_t* mask_out;
  float* probs;
  float* partial;  // [M, EPI_SPLIT * N / block_n]
  const uint8_t* presence;
  int hop, T, n_out, head_F;
  long long* dbg;   // timeline probe (profiling builds of scripts/microbench only), nullptr otherwise
  int dbg_mode;     // -DWV_TIMELINE builds only: 1 = math warps skip global stores, 2 = skip the units, 3 = skip drain conversion+stores
};

__host__ __device__ inline int staged_pitch_bytes(int block_n, bool pm = false) { return block_n * (pm ? 4 : 2) + 16; }
// Shared-memory plan.  With resident_b the CTA's whole W tile (num_kb k-blocks) is loaded once and the
// ring stages hold A k-blocks only (16 KB each): the ring then covers 4-5 tiles of loads in flight
// instead of 3, which is what bounds the small-K layers (TMA issue -> data is ~8 000 cycles under load).
__host__ inline int gemm_fixed_smem(int block_n, bool staged, int stage_bufs = STAGE_BUFS, bool pm = false) {
  return 1024 + GEMM_BAR_BYTES + (staged ? (pm ? DOWN_W_BYTES_PM : DOWN_W_BYTES) + stage_bufs * BM * staged_pitch_bytes(block_n, pm) : 0);
}
__host__ inline bool gemm_resident_b(int block_n, int num_kb, bool staged, bool nt_fixed, bool pm = false) {
  const int w_bytes = num_kb * block_n * BK * 2;
  return nt_fixed && w_bytes <= 64 * 1024 &&
         (GEMM_SMEM_LIMIT - gemm_fixed_smem(block_n, staged, STAGE_BUFS, pm) - w_bytes) / A_STAGE_BYTES >= 4;
}
__host__ inline int gemm_stage_count(int block_n, bool staged, int num_kb = 0, bool resident = false, int stage_bufs = STAGE_BUFS,
                                     bool pair = false, bool pm = false) {
  const int avail = GEMM_SMEM_LIMIT - gemm_fixed_smem(block_n, staged, stage_bufs, pm);
  const int a_bytes = pair ? 2 * A_STAGE_BYTES : A_STAGE_BYTES;
  int s = resident ? (avail - num_kb * block_n * BK * 2) / a_bytes : avail / (a_bytes + block_n * BK * 2);
  return s > MAX_STAGES ? MAX_STAGES : s;
}
__host__ inline int gemm_smem_bytes(int block_n, bool staged, int num_kb = 0, bool resident = false, int stage_bufs = STAGE_BUFS,
                                    bool pair = false, bool pm = false) {
  const int s = gemm_stage_count(block_n, staged, num_kb, resident, stage_bufs, pair, pm);
  const int a_bytes = pair ? 2 * A_STAGE_BYTES : A_STAGE_BYTES;
  return gemm_fixed_smem(block_n, staged, stage_bufs, pm) +
         (resident ? s * a_bytes + num_kb * block_n * BK * 2 : s * (a_bytes + block_n * BK * 2));
}

template <int N>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
// 640 threads x 96 registers at launch = 61440 (only registers released by this CTA's own warps
// can be re-acquired): the TMA / MMA warpgroup drops to 40, the drain warpgroup drops to 80, the
// twelve math warps grow to 120:  128*40 + 128*80 + 384*120 = 61440.
constexpr int REGS_LIGHT = 40;
constexpr int REGS_DRAIN = 80;     // the drain warps hold one 32-register TMEM chunk: what they give back goes to the math warps
constexpr int REGS_MATH = 120;
static_assert(128 * REGS_LIGHT + 128 * REGS_DRAIN + P2_THREADS * REGS_MATH <= STAGED_THREADS * 96,
              "setmaxnreg budget exceeds the CTA's launch allocation");

// q = x / d, r = x % d with a host-computed magic = floor(2^32 / d): one mul.hi + one fix-up.
__device__ __forceinline__ void fast_divmod(uint32_t x, uint32_t d, uint32_t magic, int& q, int& r) {
  uint32_t qq = __umulhi(x, magic);
  uint32_t rr = x - qq * d;
  if (rr >= d) { ++qq; rr -= d; }
  q = static_cast<int>(qq);
  r = static_cast<int>(rr);
}
struct TileCoord {
  int clip, mi, nt;   // clip, m tile inside the clip, n tile
};
__device__ __forceinline__ TileCoord tile_coord(const GemmArgs& g, int tile) {
  TileCoord t;
  int mt;
  if (g.tiles_n == 1) { mt = tile; t.nt = 0; }
  else fast_divmod(static_cast<uint32_t>(tile), static_cast<uint32_t>(g.tiles_n), g.magic_n, mt, t.nt);
  if (g.n_clips == 1) { t.clip = 0; t.mi = mt; }
  else fast_divmod(static_cast<uint32_t>(mt), static_cast<uint32_t>(g.tiles_m_per_clip), g.magic_m, t.clip, t.mi);
  return t;
}

// The k-th tile of this CTA (physical tile index, -1 if that slot is empty, done when the CTA has no more).
// Unpaired: l = blockIdx.x + k*gridDim.x.  Pair mode: the CTA walks UNITS u = blockIdx.x + (k>>1)*gridDim.x of
// two M tiles with the same n tile, u -> (mp, nt), tile = (2*mp + (k&1))*tiles_n + nt; with an odd number of
// M tiles the second half of the last units does not exist.  With g.reverse the order is walked from the end,
// so that a kernel starts on the rows the previous kernel of the plan wrote last (still in L2).
template <bool CG2>
__device__ __forceinline__ int seq_tile(const GemmArgs& g, int k, bool& done) {
  if constexpr (CG2) {
    // CTA-pair mode: cluster c = blockIdx.x / 2 walks UNITS u = c + (k>>1) * clusters = (M tile pair mp, wide n tile nw);
    // the CTA of rank r owns M tile 2*mp + r (the last one twice when the count is odd: both CTAs then write the same values)
    // and visits the two block_n-wide halves of the wide n tile in turn: every unit is exactly two tiles of this CTA.
    const int tiles_m = g.tiles_m_per_clip * g.n_clips;
    const int ntw = g.tiles_n >> 1;
    const int num_units = ((tiles_m + 1) >> 1) * ntw;
    int u = static_cast<int>(blockIdx.x >> 1) + (k >> 1) * static_cast<int>(gridDim.x >> 1);
    done = u >= num_units;
    if (done) return -1;
    if (g.reverse) u = num_units - 1 - u;
    int mp, nw;
    if (ntw == 1) { mp = u; nw = 0; }
    else fast_divmod(static_cast<uint32_t>(u), static_cast<uint32_t>(ntw), g.magic_n2, mp, nw);
    const int mt = min(2 * mp + static_cast<int>(blockIdx.x & 1), tiles_m - 1);
    return mt * g.tiles_n + 2 * nw + (k & 1);
  }
  if (!g.pair) {
    const int l = static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x);
    done = l >= g.num_tiles;
    return done ? -1 : (g.reverse ? g.num_tiles - 1 - l : l);
  }
  const int tiles_m = g.tiles_m_per_clip * g.n_clips;
  const int num_units = ((tiles_m + 1) >> 1) * g.tiles_n;
  int u = static_cast<int>(blockIdx.x) + (k >> 1) * static_cast<int>(gridDim.x);
  done = u >= num_units;
  if (done) return -1;
  if (g.reverse) u = num_units - 1 - u;
  int mp, nt;
  if (g.tiles_n == 1) { mp = u; nt = 0; }
  else fast_divmod(static_cast<uint32_t>(u), static_cast<uint32_t>(g.tiles_n), g.magic_n, mp, nt);
  const int mt = 2 * mp + (k & 1);
  return mt < tiles_m ? mt * g.tiles_n + nt : -1;
}
template <bool CG2>
__device__ __forceinline__ int cta_tile_count(const GemmArgs& g) {
  int n = 0;
  for (int k = 0;; ++k) {
    bool done;
    const int l = seq_tile<CG2>(g, k, done);
    if (done) break;
    n += l >= 0;
  }
  return n;
}
template <bool CG2>
struct TileWalker {
  int k, tile, clip, mi, nt;   // tile = physical index, num_tiles when the CTA is done
  __device__ __forceinline__ TileWalker(const GemmArgs& g) : k(-1), tile(0), clip(0), mi(0), nt(0) { next(g); }
  __device__ __forceinline__ void next(const GemmArgs& g) {
    for (;;) {
      ++k;
      bool done;
      const int l = seq_tile<CG2>(g, k, done);
      if (done) { tile = g.num_tiles; return; }
      if (l < 0) continue;
      tile = l;
      const TileCoord t = tile_coord(g, l);
      clip = t.clip; mi = t.mi; nt = t.nt;
      return;
    }
  }
};

// ---------------------------------------------------------------------------------------------
// explicit shared-state accesses (32-bit addresses; generic pointers cost 64-bit address math)
__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
  uint2 r;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr));
  return r;
}
__device__ __forceinline__ void sts_u4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// One math unit: R output rows x 4 channels.  nrow < R only at the end of a tile / clip (FULL=false):
// rows past nrow are neither read (they may lie beyond the staging tile) nor written.
template <int TAPS, int R, bool RES, bool RAW, bool ACT, bool FULL, bool SCALE, bool DUAL = false>
__device__ __forceinline__ void staged_unit(const GemmArgs& g, uint32_t srow /*smem addr of tile row ro*/,
                                            int pitch, size_t off, size_t row_bytes, int nrow,
                                            const __half2 (&wt)[TAPS][2], const __half2 (&bs)[2], float s_act,
                                            const uint2 (&rres)[R], uint32_t dual_off = 0) {
  constexpr int HALO = TAPS - 1;
  __half2 oh[R][2];
  if constexpr (TAPS > 1) {
    __half2 x[R + HALO][2];                                  // tile rows ro .. ro+R+HALO-1
#pragma unroll
    for (int j = 0; j < R + HALO; ++j) {
      const uint2 u = (FULL || j < nrow + HALO) ? lds_u2(srow + j * pitch) : make_uint2(0u, 0u);
      x[j][0] = as_h2(u.x);
      x[j][1] = as_h2(u.y);
    }
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        __half2 a = bs[p];
#pragma unroll
        for (int j = 0; j < TAPS; ++j) a = __hfma2(wt[j][p], x[i + j][p], a);
        oh[i][p] = a;
      }
  } else {
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const uint2 u = (FULL || i < nrow) ? lds_u2(srow + i * pitch) : make_uint2(0u, 0u);
      oh[i][0] = __hadd2(as_h2(u.x), bs[0]);
      oh[i][1] = __hadd2(as_h2(u.y), bs[1]);
    }
  }
  if constexpr (DUAL) {   // second accumulator (un-tapped): same output rows = tile rows ro+HALO .. , columns + block_n/2
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const uint2 u = (FULL || i < nrow) ? lds_u2(srow + (TAPS - 1 + i) * pitch + dual_off) : make_uint2(0u, 0u);
      oh[i][0] = __hadd2(oh[i][0], as_h2(u.x));
      oh[i][1] = __hadd2(oh[i][1], as_h2(u.y));
    }
  }
  if constexpr (RES) {
#pragma unroll
    for (int i = 0; i < R; ++i) {
      oh[i][0] = __hadd2(oh[i][0], as_h2(rres[i].x));
      oh[i][1] = __hadd2(oh[i][1], as_h2(rres[i].y));
    }
  }
  if (WV_DBG_MODE(1)) nrow = -1;
  if constexpr (RAW) {
    char* op = reinterpret_cast<char*>(g.out_raw + off);
#pragma unroll
    for (int i = 0; i < R; ++i)
      if ((FULL && !WV_DBG_MODE(1)) || i < nrow)
        *reinterpret_cast<uint2*>(op + i * row_bytes) = make_uint2(as_u32(oh[i][0]), as_u32(oh[i][1]));
  }
  if constexpr (ACT) {
    char* op = reinterpret_cast<char*>(g.out_act + off);
    const __half2 s2 = h2_from(s_act, s_act);
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const __half2 a0 = elu_h2(SCALE ? __hmul2(oh[i][0], s2) : oh[i][0]);
      const __half2 a1 = elu_h2(SCALE ? __hmul2(oh[i][1], s2) : oh[i][1]);
      if ((FULL && !WV_DBG_MODE(1)) || i < nrow)
        *reinterpret_cast<uint2*>(op + i * row_bytes) = make_uint2(as_u32(a0), as_u32(a1));
    }
  }
}

// STAGED math warps.  thread = (4-channel group, R-row group): 8-byte smem reads / global accesses
// keep a warp on contiguous row segments; a thread walks its row groups in passes of 384 threads.
template <int TAPS, int R, bool RES, bool RAW, bool ACT, bool SCALE, bool DUAL = false, bool PRE = false, bool CG2 = false, bool RT = false>
__device__ __forceinline__ void staged_math_loop(const GemmArgs& g, const uint8_t* stage_tiles, int lane) {
  constexpr int HALO = TAPS - 1;
  constexpr int ROWS_OUT = BM - HALO;
  constexpr int N_GROUPS = (ROWS_OUT + R - 1) / R;
  const int pitch = staged_pitch_bytes(g.block_n);
  const int et_all = threadIdx.x - (128 + P1_WARPS * 32);
  const int gthreads = g.math_groups == 2 ? P2_THREADS / 2 : P2_THREADS;   // threads that share a tile
  const int my_grp = et_all >= gthreads ? 1 : 0;                 // group g owns the CTA's tiles g, g+2, .. and staging tile g
  const int et = et_all - my_grp * gthreads;
  const int bar_threads = P1_WARPS * 32 + gthreads;              // drain warps + this group
  const int bn = DUAL ? g.block_n >> 1 : g.block_n;              // output channels per tile
  const int n_ch = DUAL ? g.N >> 1 : g.N;                        // output channels of the layer
  const int cgs = bn >> 2;                                       // 4-channel groups per row
  const int gstride = gthreads / cgs;                            // row groups per pass
  const int cg = et % cgs, grp0 = et / cgs;
  const bool active = grp0 < gstride;
  const size_t row_bytes = static_cast<size_t>(g.ldo) * 2;
  const float s_act = g.act_scale;
  const uint32_t stage_u32 = smem_u32(stage_tiles) + cg * 8;
  int cached_nt = -1;
  __half2 wt[TAPS][2], bs[2];
  const bool two = g.math_groups == 2;
  // residual tiles staged by TMA (g.res_tma): buffers behind the staging tiles, barriers in the CTA's barrier block
  constexpr bool rt = RES && !PRE && RT;   // compile-time: the two residual paths do not share a loop body (registers)
  const int res_tile_bytes = BM * g.block_n * 2;
  const uint32_t res_u32 = ((smem_u32(stage_tiles) + g.stage_bufs * BM * pitch + 127u) & ~127u) + cg * 8;
  uint64_t* res_bars = reinterpret_cast<uint64_t*>(const_cast<uint8_t*>(stage_tiles) - DOWN_W_BYTES - GEMM_BAR_BYTES) + RES_BAR_INDEX;
  int sb = two ? my_grp : 0;
  const int tiles_cta = cta_tile_count<CG2>(g);
  int tiles_left = two ? (tiles_cta - my_grp + 1) >> 1 : tiles_cta;   // tiles this group still has to process
  const int keep = two ? 1 : g.stage_bufs;                         // no later drain waits for the last `keep` tiles
  int dbg_it = 0, idx = 0;
  for (TileWalker<CG2> tc(g); tc.tile < g.num_tiles; tc.next(g), ++idx) {
    if (two && (idx & 1) != my_grp) continue;
    const int r_base = tc.mi * ROWS_OUT;                           // first OUTPUT row of the tile
    const int c = tc.nt * bn + cg * 4;
    if (active && tc.nt != cached_nt) {                            // per-CTA constant when N fits one tile
      cached_nt = tc.nt;
      if (g.bias != nullptr) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.bias + c));
        bs[0] = h2_from(b0.x, b0.y); bs[1] = h2_from(b0.z, b0.w);
      } else {
        bs[0] = bs[1] = h2_from(0.f, 0.f);
      }
      if constexpr (TAPS > 1) {
#pragma unroll
        for (int j = 0; j < TAPS; ++j) {
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(g.dw_w + j * n_ch + c));
          wt[j][0] = h2_from(w0.x, w0.y); wt[j][1] = h2_from(w0.z, w0.w);
        }
      }
    }
    const size_t base = (static_cast<size_t>(tc.clip) * g.rows_per_clip + r_base) * g.ldo + c;
    const int rows_left = min(g.rows_per_clip - r_base, ROWS_OUT);   // valid output rows of this tile
    // The residual rows do not depend on the staged tile: those of the first unit are requested
    // before waiting for the drain warps, those of a
    // second unit before the first unit's math, so their latency overlaps the wait / the math.
    uint2 rres[R], rnext[R];
    const int rbuf = idx & 1;                                      // two groups: == my_grp
    auto load_res = [&](int ro, uint2 (&r)[R]) {
      if constexpr (rt) {                                          // rows past the clip end are TMA zero fill (and never stored)
        const uint32_t a = res_u32 + rbuf * res_tile_bytes + ro * (g.block_n * 2);
#pragma unroll
        for (int i = 0; i < R; ++i) r[i] = ro + i < BM ? lds_u2(a + i * (g.block_n * 2)) : make_uint2(0u, 0u);
        return;
      }
      if constexpr (PRE) {   // v[t,c] = b[c] + sum_j w[j][c] * x[t-4+j], rounded to fp16 like conv_pre_kernel stores it
        const float* xp = g.pre_x + static_cast<size_t>(tc.clip) * g.pre_T;
        const int t0 = r_base + ro - 4;
        float xs[R + 4];
#pragma unroll
        for (int k = 0; k < R + 4; ++k) {
          const int tt = t0 + k;
          xs[k] = (tt >= 0 && tt < g.pre_T) ? __ldg(xp + tt) : 0.f;
        }
        // taps and bias come from L1 every unit (24 registers would not fit next to the unit's working set)
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.pre_b + c));
        float o[R][4];
#pragma unroll
        for (int i = 0; i < R; ++i) { o[i][0] = b0.x; o[i][1] = b0.y; o[i][2] = b0.z; o[i][3] = b0.w; }
#pragma unroll
        for (int j = 0; j < 5; ++j) {   // per output the FMA order is j = 0..4, as in conv_pre_kernel
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(g.pre_w + j * n_ch + c));
#pragma unroll
          for (int i = 0; i < R; ++i) {
            o[i][0] = fmaf(w0.x, xs[i + j], o[i][0]);
            o[i][1] = fmaf(w0.y, xs[i + j], o[i][1]);
            o[i][2] = fmaf(w0.z, xs[i + j], o[i][2]);
            o[i][3] = fmaf(w0.w, xs[i + j], o[i][3]);
          }
        }
#pragma unroll
        for (int i = 0; i < R; ++i)
          r[i] = ro + i < rows_left ? make_uint2(pack_act2(o[i][0], o[i][1]), pack_act2(o[i][2], o[i][3])) : make_uint2(0u, 0u);
      } else {
        const char* rp = reinterpret_cast<const char*>(g.residual + base + static_cast<size_t>(ro) * g.ldo);
#pragma unroll
        for (int i = 0; i < R; ++i)
          r[i] = ro + i < rows_left ? __ldcg(reinterpret_cast<const uint2*>(rp + i * row_bytes)) : make_uint2(0u, 0u);
      }
    };
    int grp = grp0;
    bool have = active && grp < N_GROUPS && grp * R < rows_left;
    bool early2 = false;                                           // second unit's residual already requested
    if constexpr (RES) {
      if (have && !rt) load_res(grp * R, rres);
      if (!PRE && g.res_early2 && !rt) {                           // both units of the tile in flight across the barrier wait
        const int gn0 = grp + gstride;
        if (have && gn0 < N_GROUPS && gn0 * R < rows_left) { load_res(gn0 * R, rnext); early2 = true; }
      }
    }
    if (et_all == 0) WV_DBG(7, dbg_it);                           // warp 0 reaches the hand-off barrier
    named_bar_sync(BAR_ST_FULL + sb, bar_threads);                 // drain warps staged tile sb
    if constexpr (RES) {
      if constexpr (rt) {
        mbar_wait(&res_bars[rbuf], static_cast<uint32_t>(idx >> 1) & 1u);   // the tile's residual rows landed
        if (have) load_res(grp * R, rres);
      }
    }
    if (et_all == 0) WV_DBG(5, dbg_it);
    if (lane == 0) WV_DBG(24 + (et_all >> 5), dbg_it);   // per math warp: start
    const uint32_t tile_u32 = stage_u32 + sb * (BM * pitch);
    if (WV_DBG_MODE(2)) have = false;
    while (have) {
      const int ro = grp * R;                                      // tile-relative output row
      const int gn = grp + gstride;
      const bool have_next = gn < N_GROUPS && gn * R < rows_left;
      if constexpr (RES) {
        if (have_next && !early2) load_res(gn * R, rnext);
        early2 = false;
      }
      const size_t off = base + static_cast<size_t>(ro) * g.ldo;
      const uint32_t srow = tile_u32 + ro * pitch;
      if (ro + R <= rows_left)
        staged_unit<TAPS, R, RES, RAW, ACT, true, SCALE, DUAL>(g, srow, pitch, off, row_bytes, R, wt, bs, s_act, rres, bn * 2);
      else
        staged_unit<TAPS, R, RES, RAW, ACT, false, SCALE, DUAL>(g, srow, pitch, off, row_bytes, rows_left - ro, wt, bs, s_act, rres, bn * 2);
      if constexpr (RES) {
#pragma unroll
        for (int i = 0; i < R; ++i) rres[i] = rnext[i];
      }
      grp = gn;
      have = have_next;
    }
    __syncwarp();
    if constexpr (RES) {
      if (rt && lane == 0) mbar_arrive(&res_bars[2 + rbuf]);       // this warp is done with the residual buffer
    }
    if (lane == 0) WV_DBG(12 + (et_all >> 5), dbg_it);   // per math warp: end
    if (--tiles_left >= keep) named_bar_arrive(BAR_ST_EMPTY + sb, bar_threads);   // tile sb may be refilled
    if (et_all == 0) WV_DBG(6, dbg_it);              // warp 0: after releasing the staging tile
    ++dbg_it;
    if (!two && ++sb == g.stage_bufs) sb = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// STAGED math warps, strided variant: the causal depthwise down-conv k=2R, s=R that follows the
// encoder's channel-doubling 1x1 (modules/seanet.py:745-771) + FiLM (seanet.py:928-966):
//   v[i,c] = bias[c] + sum_{j<2R} w[j][c] * S[i*R - R + j][c];  v = v*gamma + beta
// A tile holds input rows [mi*OUTS*R - R, +128) and produces OUTS = 128/R - 1 output rows.
// thread = (4-channel group, output row); taps are staged in shared memory (2R*4 floats per thread
// would not fit in registers for R = 8).
template <int R, bool CG2>
__device__ __forceinline__ void staged_down_loop(const GemmArgs& g, const uint8_t* stage_tiles, uint8_t* down_w, int lane) {
  constexpr int OUTS = BM / R - 1;
  const int pitch = staged_pitch_bytes(g.block_n);
  const int et = threadIdx.x - (128 + P1_WARPS * 32);
  const int cgs = g.block_n >> 2;
  const int gstride = P2_THREADS / cgs;                          // output rows per pass
  const int cg = et % cgs, row0 = et / cgs;
  const bool active = row0 < gstride;
  const bool raw = g.out_raw != nullptr, act = g.out_act != nullptr;
  const float s_act = g.act_scale;
  const uint32_t stage_u32 = smem_u32(stage_tiles) + cg * 8;
  const uint32_t w_u32 = smem_u32(down_w) + cg * 8;              // taps staged as fp16 [2R][block_n]
  const int band_w = g.film != nullptr ? g.N / g.film_bands : 1;
  int cached_nt = -1, cached_half = -1;
  __half2 bs2[2];
  const __half2 s2 = h2_from(s_act, s_act);
  int sb = 0;
  int tiles_left = cta_tile_count<CG2>(g);
  for (TileWalker<CG2> tc(g); tc.tile < g.num_tiles; tc.next(g)) {
    const int c = tc.nt * g.block_n + cg * 4;
    // taps are staged per n tile; CTA-pair mode alternates between the two halves of a wide n tile, so both halves are
    // staged together (keyed by the wide tile) and the bias is re-read (L1) when the half changes
    const int tap_key = CG2 ? (tc.nt >> 1) : tc.nt;
    const int half = CG2 ? (tc.nt & 1) : 0;
    if (tap_key != cached_nt) {
      cached_nt = tap_key;
      named_bar_sync(BAR_TAPS, P2_THREADS);                           // previous taps no longer read
      const int n_half = CG2 ? 2 : 1;
      const int nt0 = CG2 ? (tc.nt & ~1) : tc.nt;
      for (int i = et; i < n_half * 2 * R * cgs; i += P2_THREADS) {
        const int hh = i / (2 * R * cgs), ii = i % (2 * R * cgs);
        const int j = ii / cgs, q = ii % cgs;
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(g.dw_w + j * g.N + (nt0 + hh) * g.block_n + q * 4));
        const __half2 h0 = h2_from(w4.x, w4.y), h1 = h2_from(w4.z, w4.w);
        *reinterpret_cast<uint2*>(down_w + ((hh * 2 * R + j) * g.block_n + q * 4) * 2) =
            make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      }
      named_bar_sync(BAR_TAPS, P2_THREADS);
      cached_half = -1;
    }
    if (half != cached_half) {
      cached_half = half;
      if (g.bias != nullptr) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(g.bias + c));
        bs2[0] = h2_from(b0.x, b0.y); bs2[1] = h2_from(b0.z, b0.w);
      } else {
        bs2[0] = bs2[1] = h2_from(0.f, 0.f);
      }
    }
    const uint32_t w_half_u32 = w_u32 + half * (2 * R * g.block_n * 2);
    __half2 gm2 = h2_from(1.f, 1.f), bt2 = h2_from(0.f, 0.f);
    if (g.film != nullptr) {
      const float* fp = g.film + static_cast<size_t>(tc.clip) * g.film_stride + (c / band_w) * 2;
      const float gm = __ldcg(fp), bt = __ldcg(fp + 1);
      gm2 = h2_from(gm, gm);
      bt2 = h2_from(bt, bt);
    }
    named_bar_sync(BAR_ST_FULL + sb, EPI_THREADS);
    if (active) {
      const uint32_t tile_u32 = stage_u32 + sb * (BM * pitch);
      const int i_base = tc.mi * OUTS;                              // first output row of the tile
      const int outs_left = g.rows_per_clip_out - i_base;
      const size_t base = (static_cast<size_t>(tc.clip) * g.rows_per_clip_out + i_base) * g.ldo + c;
      for (int lo = row0; lo < OUTS && lo < outs_left; lo += gstride) {
        // two independent HFMA2 chains per channel pair (even / odd taps) keep the fp16 partial
        // sums short
        __half2 a0[2] = {h2_from(0.f, 0.f), h2_from(0.f, 0.f)};
        __half2 a1[2] = {h2_from(0.f, 0.f), h2_from(0.f, 0.f)};
        const uint32_t srow = tile_u32 + lo * R * pitch;
#pragma unroll
        for (int j = 0; j < 2 * R; ++j) {
          const uint2 u = lds_u2(srow + j * pitch);
          const uint2 w = lds_u2(w_half_u32 + j * g.block_n * 2);
          if (j & 1) {
            a1[0] = __hfma2(as_h2(w.x), as_h2(u.x), a1[0]);
            a1[1] = __hfma2(as_h2(w.y), as_h2(u.y), a1[1]);
          } else {
            a0[0] = __hfma2(as_h2(w.x), as_h2(u.x), a0[0]);
            a0[1] = __hfma2(as_h2(w.y), as_h2(u.y), a0[1]);
          }
        }
        __half2 o0 = __hadd2(__hadd2(a0[0], a1[0]), bs2[0]);
        __half2 o1 = __hadd2(__hadd2(a0[1], a1[1]), bs2[1]);
        o0 = __hfma2(o0, gm2, bt2);                                // FiLM (identity when absent)
        o1 = __hfma2(o1, gm2, bt2);
        const size_t off = base + static_cast<size_t>(lo) * g.ldo;
        if (raw) *reinterpret_cast<uint2*>(g.out_raw + off) = make_uint2(as_u32(o0), as_u32(o1));
        if (act)
          *reinterpret_cast<uint2*>(g.out_act + off) =
              make_uint2(as_u32(elu_h2(__hmul2(o0, s2))), as_u32(elu_h2(__hmul2(o1, s2))));
      }
    }
    __syncwarp();
    if (--tiles_left >= g.stage_bufs) named_bar_arrive(BAR_ST_EMPTY + sb, EPI_THREADS);
    if (++sb == g.stage_bufs) sb = 0;
  }
}

// last_mode math: one thread per output sample of the tile (124 of the 384 math threads).
template <bool CG2>
__device__ __forceinline__ void staged_last_loop(const GemmArgs& g, const uint8_t* stage_tiles) {
  constexpr int ROWS_OUT = BM - 4;
  const int pitch = staged_pitch_bytes(g.block_n);
  const int et = threadIdx.x - (128 + P1_WARPS * 32);
  const uint32_t stage_u32 = smem_u32(stage_tiles);
  int sb = 0;
  int tiles_left = cta_tile_count<CG2>(g);
  for (TileWalker<CG2> tc(g); tc.tile < g.num_tiles; tc.next(g)) {
    named_bar_sync(BAR_ST_FULL + sb, EPI_THREADS);
    const int t = tc.mi * ROWS_OUT + et;
    if (et < ROWS_OUT && t < g.last_T) {
      const uint32_t row = stage_u32 + sb * (BM * pitch) + et * pitch;
      float acc = g.last_bias;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(row + j * pitch + j * 4));
        acc += v;
      }
      const float wm = tanhf(acc);
      const long long o = static_cast<long long>(tc.clip) * g.last_T + t;
      if (g.last_wm != nullptr) g.last_wm[o] = wm;
      if (g.last_y != nullptr) g.last_y[o] = __ldcg(g.last_x + o) + wm;
    }
    __syncwarp();
    if (--tiles_left >= g.stage_bufs) named_bar_arrive(BAR_ST_EMPTY + sb, EPI_THREADS);
    if (++sb == g.stage_bufs) sb = 0;
  }
}

template <int TAPS, int R, bool CG2>
__device__ __forceinline__ void staged_math_dispatch(const GemmArgs& g, const uint8_t* stage_tiles, int lane) {
  const bool res = g.residual != nullptr, raw = g.out_raw != nullptr, act = g.out_act != nullptr;
  const bool scale = act && g.act_scale != 1.f;
#define WV_MATH(RS, W, A, S) staged_math_loop<TAPS, R, RS, W, A, S, false, false, CG2>(g, stage_tiles, lane)
#define WV_MATH_RT(W, A, S) staged_math_loop<TAPS, R, true, W, A, S, false, false, CG2, true>(g, stage_tiles, lane)
  if (res && g.res_tma) {   // residual tile staged in shared memory by warp 3
    if (raw && act) { if (scale) WV_MATH_RT(true, true, true); else WV_MATH_RT(true, true, false); }
    else if (raw) WV_MATH_RT(true, false, false);
    else { if (scale) WV_MATH_RT(false, true, true); else WV_MATH_RT(false, true, false); }
  } else if (res) {
    if (raw && act) { if (scale) WV_MATH(true, true, true, true); else WV_MATH(true, true, true, false); }
    else if (raw) WV_MATH(true, true, false, false);
    else { if (scale) WV_MATH(true, false, true, true); else WV_MATH(true, false, true, false); }
  } else {
    if (raw && act) { if (scale) WV_MATH(false, true, true, true); else WV_MATH(false, true, true, false); }
    else if (raw) WV_MATH(false, true, false, false);
    else { if (scale) WV_MATH(false, false, true, true); else WV_MATH(false, false, true, false); }
  }
#undef WV_MATH
#undef WV_MATH_RT
}
template <int TAPS, bool CG2>
__device__ __forceinline__ void staged_math_rows(const GemmArgs& g, const uint8_t* stage_tiles, int lane) {
  if constexpr (TAPS == 5) {
    if (g.pre_w != nullptr) {   // first encoder resblock: residual recomputed from the waveform (conv_pre)
      const bool sc = g.act_scale != 1.f;
#define WV_PRE(RR, RAWO, SC, DU) staged_math_loop<5, RR, true, RAWO, true, SC, DU, true, CG2>(g, stage_tiles, lane)
      if (g.dual) {             // single-resblock stages (Locator): the same launch also carries the spectrogram 1x1
        if (g.unit_rows == 6) { if (sc) WV_PRE(6, false, true, true); else WV_PRE(6, false, false, true); }
        else { if (sc) WV_PRE(4, false, true, true); else WV_PRE(4, false, false, true); }
      } else {
        if (g.unit_rows == 6) { if (sc) WV_PRE(6, true, true, false); else WV_PRE(6, true, false, false); }
        else { if (sc) WV_PRE(4, true, true, false); else WV_PRE(4, true, false, false); }
      }
#undef WV_PRE
      return;
    }
    if (g.dual) {   // last encoder resblock + spectrogram 1x1: residual in, activated output only
      if (g.unit_rows == 6) {
        if (g.act_scale != 1.f) staged_math_loop<5, 6, true, false, true, true, true, false, CG2>(g, stage_tiles, lane);
        else staged_math_loop<5, 6, true, false, true, false, true, false, CG2>(g, stage_tiles, lane);
      } else {
        if (g.act_scale != 1.f) staged_math_loop<5, 4, true, false, true, true, true, false, CG2>(g, stage_tiles, lane);
        else staged_math_loop<5, 4, true, false, true, false, true, false, CG2>(g, stage_tiles, lane);
      }
      return;
    }
  }
  // 6- and 8-row units were measured slower at 96 columns (16 row groups per pass: 4-row units already fill two
  // passes); at 64 columns (24 row groups per pass) 4-row units leave the second pass 7/24 occupied
  if (g.unit_rows == 6) staged_math_dispatch<TAPS, 6, CG2>(g, stage_tiles, lane);
  else staged_math_dispatch<TAPS, 4, CG2>(g, stage_tiles, lane);
}

// ---------------------------------------------------------------------------------------------
// PRECISE MODE (EPI_STAGED_PM / EPI_L2NORM_PM / EPI_STFT_PM).  The thresholded outputs of the path
// (locator mask = logit > 0.5, model/watermarking.py:717; decoded bits = mean sigmoid >= 0.5,
// waveverify/core.py:577-586) must equal the fp32 reference's, so the nets that produce them can run
// with fp32-accurate arithmetic on the same tensor-core pipeline:
//   * every 16-bit GEMM operand is a SPLIT pair  v = hi + lo  (hi = rn16(v), lo = rn16(v - hi): ~22 bits);
//     the contraction [hi | lo | hi] x [Wh | Wh | Wl] = (hi+lo)*Wh + hi*Wl drops only lo*Wl (2^-22 relative);
//     products of fp16 operands are exact in the fp32 accumulator;
//   * the STAGED epilogue stages the accumulator tile as fp32 and runs taps / bias / residual / ELU in fp32;
//     raw residual streams are stored fp32, activated streams as split pairs for the next GEMM.
// Same tile geometry, ring, TMEM staging and barriers as the fp16 path; the loops below are written for
// accuracy, not for issue rate (the Locator is 2.6 % of the path's FLOPs).
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}

// STAGED_PM math warps: thread = (4-channel group, 3-row group) over the fp32 staging tile (42 / 43 row groups per
// tile: one pass of the 384 math threads at 32 columns, two at 64).
//   v = bias[c] + sum_j w[j][c] * S[r-taps+1+j][c]  (+ residual32);  out_raw32 = v;  out_act = split(ELU(v*s))
template <int TAPS>
__device__ __forceinline__ void pm_math_loop(const GemmArgs& g, const uint8_t* stage_tiles) {
  constexpr int HALO = TAPS - 1;
  constexpr int ROWS_OUT = BM - HALO;
  constexpr int R = 3;
  const int pitch = staged_pitch_bytes(g.block_n, true);
  const int et = threadIdx.x - (128 + P1_WARPS * 32);
  const int cgs = g.block_n >> 2;
  const int gstride = P2_THREADS / cgs;
  const int cg = et % cgs, grp0 = et / cgs;
  const bool active = grp0 < gstride;
  const float s_act = g.act_scale;
  const bool has_res = g.residual32 != nullptr, has_raw = g.out_raw32 != nullptr, has_act = g.out_act != nullptr;
  const uint32_t stage_u32 = smem_u32(stage_tiles) + cg * 16;
  int cached_nt = -1;
  float4 wt[TAPS], bs = make_float4(0.f, 0.f, 0.f, 0.f);
  int sb = 0;
  int tiles_left = cta_tile_count<false>(g);
  for (TileWalker<false> tc(g); tc.tile < g.num_tiles; tc.next(g)) {
    const int r_base = tc.mi * ROWS_OUT;
    const int c = tc.nt * g.block_n + cg * 4;
    if (active && tc.nt != cached_nt) {
      cached_nt = tc.nt;
      bs = g.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(g.bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if constexpr (TAPS > 1) {
#pragma unroll
        for (int j = 0; j < TAPS; ++j) wt[j] = __ldg(reinterpret_cast<const float4*>(g.dw_w + j * g.N + c));
      }
    }
    const int rows_left = min(g.rows_per_clip - r_base, ROWS_OUT);
    const size_t row0 = static_cast<size_t>(tc.clip) * g.rows_per_clip + r_base;
    // the residual rows of the first unit do not depend on the staged tile: requested before the hand-off barrier
    float4 rr[R];
    int ro = grp0 * R;
    auto load_res = [&](int ro_) {
#pragma unroll
      for (int i = 0; i < R; ++i)
        rr[i] = (has_res && ro_ + i < rows_left) ? __ldcg(reinterpret_cast<const float4*>(g.residual32 + (row0 + ro_ + i) * g.ldo + c))
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    if (active && ro < rows_left) load_res(ro);
    named_bar_sync(BAR_ST_FULL + sb, EPI_THREADS);
    if (active) {
      const uint32_t tile_u32 = stage_u32 + sb * (BM * pitch);
      for (; ro < rows_left; ro += gstride * R) {
        float4 x[R + HALO];
#pragma unroll
        for (int j = 0; j < R + HALO; ++j) x[j] = lds_f4(tile_u32 + min(ro + j, BM - 1) * pitch);   // rows past 127 feed no valid output
        float4 v[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
          if constexpr (TAPS > 1) {
            v[i] = bs;
#pragma unroll
            for (int j = 0; j < TAPS; ++j) {
              v[i].x = fmaf(wt[j].x, x[i + j].x, v[i].x); v[i].y = fmaf(wt[j].y, x[i + j].y, v[i].y);
              v[i].z = fmaf(wt[j].z, x[i + j].z, v[i].z); v[i].w = fmaf(wt[j].w, x[i + j].w, v[i].w);
            }
          } else {
            v[i] = make_float4(x[i].x + bs.x, x[i].y + bs.y, x[i].z + bs.z, x[i].w + bs.w);
          }
          v[i].x += rr[i].x; v[i].y += rr[i].y; v[i].z += rr[i].z; v[i].w += rr[i].w;
        }
        const int ro_next = ro + gstride * R;
        const int ro_cur = ro;
        if (has_raw) {
#pragma unroll
          for (int i = 0; i < R; ++i)
            if (ro_cur + i < rows_left) *reinterpret_cast<float4*>(g.out_raw32 + (row0 + ro_cur + i) * g.ldo + c) = v[i];
        }
        if (ro_next < rows_left) load_res(ro_next);          // next unit's residual in flight during the ELUs
        if (has_act) {
#pragma unroll
          for (int i = 0; i < R; ++i)
            if (ro_cur + i < rows_left)
              st_split4(g.out_act + (row0 + ro_cur + i) * g.ldo_act + c, g.lo_off, elu_precise(v[i].x * s_act), elu_precise(v[i].y * s_act),
                        elu_precise(v[i].z * s_act), elu_precise(v[i].w * s_act));
        }
      }
    }
    __syncwarp();
    if (--tiles_left >= g.stage_bufs) named_bar_arrive(BAR_ST_EMPTY + sb, EPI_THREADS);
    if (++sb == g.stage_bufs) sb = 0;
  }
}

// STAGED_PM strided variant: causal depthwise down-conv k=2R, s=R (+ FiLM) in fp32 (cf. staged_down_loop)
template <int R>
__device__ __forceinline__ void pm_down_loop(const GemmArgs& g, const uint8_t* stage_tiles, uint8_t* down_w) {
  constexpr int OUTS = BM / R - 1;
  const int pitch = staged_pitch_bytes(g.block_n, true);
  const int et = threadIdx.x - (128 + P1_WARPS * 32);
  const int cgs = g.block_n >> 2;
  const int gstride = P2_THREADS / cgs;
  const int cg = et % cgs, row0 = et / cgs;
  const bool active = row0 < gstride;
  const float s_act = g.act_scale;
  const uint32_t stage_u32 = smem_u32(stage_tiles) + cg * 16;
  const uint32_t w_u32 = smem_u32(down_w) + cg * 16;             // taps staged as fp32 [2R][block_n]
  const int band_w = g.film != nullptr ? g.N / g.film_bands : 1;
  int cached_nt = -1;
  float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
  int sb = 0;
  int tiles_left = cta_tile_count<false>(g);
  for (TileWalker<false> tc(g); tc.tile < g.num_tiles; tc.next(g)) {
    const int c = tc.nt * g.block_n + cg * 4;
    if (tc.nt != cached_nt) {
      cached_nt = tc.nt;
      named_bar_sync(BAR_TAPS, P2_THREADS);
      for (int i = et; i < 2 * R * cgs; i += P2_THREADS) {
        const int j = i / cgs, q = i % cgs;
        *reinterpret_cast<float4*>(down_w + (j * g.block_n + q * 4) * 4) =
            __ldg(reinterpret_cast<const float4*>(g.dw_w + j * g.N + tc.nt * g.block_n + q * 4));
      }
      named_bar_sync(BAR_TAPS, P2_THREADS);
      bs = g.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(g.bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float gm = 1.f, bt = 0.f;
    if (g.film != nullptr) {
      const float* fp = g.film + static_cast<size_t>(tc.clip) * g.film_stride + (c / band_w) * 2;
      gm = __ldcg(fp); bt = __ldcg(fp + 1);
    }
    named_bar_sync(BAR_ST_FULL + sb, EPI_THREADS);
    if (active) {
      const uint32_t tile_u32 = stage_u32 + sb * (BM * pitch);
      const int i_base = tc.mi * OUTS;
      const int outs_left = g.rows_per_clip_out - i_base;
      const size_t rowb = static_cast<size_t>(tc.clip) * g.rows_per_clip_out + i_base;
      for (int lo = row0; lo < OUTS && lo < outs_left; lo += gstride) {
        float4 v = bs;
        const uint32_t srow = tile_u32 + lo * R * pitch;
#pragma unroll
        for (int j = 0; j < 2 * R; ++j) {
          const float4 u = lds_f4(srow + j * pitch);
          const float4 w = lds_f4(w_u32 + j * g.block_n * 4);
          v.x = fmaf(w.x, u.x, v.x); v.y = fmaf(w.y, u.y, v.y); v.z = fmaf(w.z, u.z, v.z); v.w = fmaf(w.w, u.w, v.w);
        }
        if (g.film != nullptr) { v.x = fmaf(v.x, gm, bt); v.y = fmaf(v.y, gm, bt); v.z = fmaf(v.z, gm, bt); v.w = fmaf(v.w, gm, bt); }
        const size_t row = rowb + lo;
        if (g.out_raw32 != nullptr) *reinterpret_cast<float4*>(g.out_raw32 + row * g.ldo + c) = v;
        if (g.out_act != nullptr)
          st_split4(g.out_act + row * g.ldo_act + c, g.lo_off, elu_precise(v.x * s_act), elu_precise(v.y * s_act),
                    elu_precise(v.z * s_act), elu_precise(v.w * s_act));
      }
    }
    __syncwarp();
    if (--tiles_left >= g.stage_bufs) named_bar_arrive(BAR_ST_EMPTY + sb, EPI_THREADS);
    if (++sb == g.stage_bufs) sb = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// CG2 (STAGED only): the CTA-pair instantiation.  It is a separate kernel because a kernel that contains cta_group::2
// instructions is marked as such in its ELF attributes and can only be launched as clusters of two.
template <int EPI, bool CG2 = false>
__global__ void __launch_bounds__(gemm_threads<EPI>(), 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ CUtensorMap tmR, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  const int b_stage_bytes = g.block_n * BK * 2;
  const int a_stage_bytes = g.pair ? 2 * A_STAGE_BYTES : A_STAGE_BYTES;
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + g.stages * a_stage_bytes;
  const int num_kb = (g.K + BK - 1) / BK;
  uint8_t* after = smemB + (g.resident_b ? num_kb : g.stages) * b_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(after);
  uint64_t* full = bars;                          // [MAX_STAGES]
  uint64_t* empty = bars + MAX_STAGES;            // [MAX_STAGES]
  uint64_t* acc_full = bars + 2 * MAX_STAGES;     // [MAX_ACC_STAGES]
  uint64_t* acc_empty = acc_full + MAX_ACC_STAGES;
  uint64_t* w_full = acc_empty + MAX_ACC_STAGES;  // resident W tile landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);
  uint8_t* down_w = after + GEMM_BAR_BYTES;       // [2r][block_n] fp32 (STAGED down-conv only)
  constexpr bool STG = EPI == EPI_STAGED || EPI == EPI_STAGED_PM;
  constexpr bool PM = EPI >= EPI_STAGED_PM;
  constexpr bool IS_STFT = EPI == EPI_STFT || EPI == EPI_STFT_PM;
  constexpr bool IS_L2N = EPI == EPI_L2NORM || EPI == EPI_L2NORM_PM;
  uint8_t* stage_tiles = down_w + (PM ? DOWN_W_BYTES_PM : DOWN_W_BYTES);   // [STAGE_BUFS][BM][pitch] (STAGED only)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile geometry (host computed): overlapping per-clip tiles carry the depthwise halo
  const int halo = g.tile_halo;
  const int rows_out = g.tile_stride;
  const uint32_t stage_bytes = static_cast<uint32_t>(A_STAGE_BYTES + (g.resident_b ? 0 : b_stage_bytes));
  const int acc_stages = g.acc_stages, acc_cols = g.acc_cols;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (g.a2_split > 0) tma_prefetch_desc(&tmR);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < g.stages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    mbar_init(w_full, 1);
    if (STG && g.res_tma) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bars[RES_BAR_INDEX + i], 1);                                                    // full: the producer's expect_tx
        mbar_init(&bars[RES_BAR_INDEX + 2 + i], g.math_groups == 2 ? P2_WARPS / 2 : P2_WARPS);     // empty: one arrive per math warp
      }
    }
    for (int i = 0; i < MAX_ACC_STAGES; ++i) {
      mbar_init(&acc_full[i], 1);
      // one arrive per draining warp (CTA-pair mode: the leader's barrier also collects the peer's drain warps)
      mbar_init(&acc_empty[i], STG ? (CG2 ? 2 * P1_WARPS : P1_WARPS) : (IS_STFT && g.epi_groups > 1 ? EPI_WARPS / g.epi_groups : EPI_WARPS));
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (CG2) tmem_alloc_cg2(tmem_slot, TMEM_COLS);
    else tmem_alloc(tmem_slot, TMEM_COLS);
  }
  tc_fence_before();
  if constexpr (CG2) cluster_sync_all();   // the peer's barriers are initialised before anything signals them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above overlaps the previous kernel's tail (programmatic dependent launch)
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if constexpr (STG) reg_dealloc<REGS_LIGHT>();
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int dbg_it = 0;
      const uint64_t pol = l2_policy_evict_first();
      if (g.resident_b) {   // the n tile of a CTA is fixed (grid is a multiple of tiles_n): load W once
        bool d0;
        const int n_fixed = tile_coord(g, seq_tile<CG2>(g, 0, d0)).nt * g.block_n;
        mbar_arrive_expect_tx(w_full, static_cast<uint32_t>(num_kb * b_stage_bytes));
        for (int kb = 0; kb < num_kb; ++kb) tma_load_2d(smemB + kb * b_stage_bytes, &tmB, w_full, kb * BK, n_fixed);
      }
      if constexpr (CG2) {
        // CTA pair: this CTA's A tile and ITS HALF (block_n rows) of the 2 * block_n wide W k-block; the bytes of both CTAs
        // complete on the leader's full barrier, which the leader alone arms
        const bool leader = (blockIdx.x & 1) == 0;
        const uint32_t pair_bytes = 2u * static_cast<uint32_t>(A_STAGE_BYTES + b_stage_bytes);
        for (int ku = 0;; ++ku) {
          bool done;
          const int l0 = seq_tile<CG2>(g, 2 * ku, done);
          if (done) break;
          const TileCoord t0 = tile_coord(g, l0);
          const int r0 = t0.mi * rows_out - halo;
          const int nrow0 = (t0.nt + static_cast<int>(blockIdx.x & 1)) * g.block_n;   // t0.nt is even: first half of the wide tile
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            if (leader) mbar_arrive_expect_tx(&full[stage], pair_bytes);
            const bool sh = g.kb_split > 0 && kb >= g.kb_split;
            tma_load_3d_cg2(smemA + stage * A_STAGE_BYTES, &tmA, &full[stage], (sh ? kb - g.kb_split : kb) * BK, r0 - (sh ? 1 : 0), t0.clip);
            tma_load_2d_cg2(smemB + stage * b_stage_bytes, &tmB, &full[stage], kb * BK, nrow0);
            if (++stage == g.stages) { stage = 0; phase ^= 1; }
          }
        }
      } else if (g.pair) {
        // one ring stage = the k-block of TWO M tiles (same n tile) + the W k-block they share
        for (int ku = 0;; ++ku) {
          bool done, d1;
          const int l0 = seq_tile<CG2>(g, 2 * ku, done);
          if (done) break;
          const int l1 = seq_tile<CG2>(g, 2 * ku + 1, d1);
          const TileCoord t0 = tile_coord(g, l0);
          const TileCoord t1 = l1 >= 0 ? tile_coord(g, l1) : t0;
          const int ra = t0.mi * rows_out - halo, rb = t1.mi * rows_out - halo;
          const int n0 = t0.nt * g.block_n;
          const uint32_t bytes = static_cast<uint32_t>((l1 >= 0 ? 2 : 1) * A_STAGE_BYTES + b_stage_bytes);
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_arrive_expect_tx(&full[stage], bytes);
            uint8_t* a0 = smemA + stage * a_stage_bytes;
            const bool sh = g.kb_split > 0 && kb >= g.kb_split;      // [a[i] | a[i-1]] contraction (see kb_split)
            const int kc = (sh ? kb - g.kb_split : kb) * BK;
            tma_load_3d(a0, &tmA, &full[stage], kc, ra - (sh ? 1 : 0), t0.clip);
            if (l1 >= 0) tma_load_3d(a0 + A_STAGE_BYTES, &tmA, &full[stage], kc, rb - (sh ? 1 : 0), t1.clip);
            tma_load_2d(smemB + stage * b_stage_bytes, &tmB, &full[stage], kb * BK, n0);
            if (++stage == g.stages) { stage = 0; phase ^= 1; }
          }
        }
      } else
      {
      TileWalker<CG2> pf(g);
      for (int i = 0; i < g.a_prefetch && pf.tile < g.num_tiles; ++i) pf.next(g);
      for (TileWalker<CG2> tc(g); tc.tile < g.num_tiles; tc.next(g), ++dbg_it) {
        if (g.a_prefetch > 0 && pf.tile < g.num_tiles) {
          const int pr0 = pf.mi * rows_out - halo;
          if (pf.nt == 0 && pr0 >= 0 && g.phases <= 1) {
            const int kbs = g.a2_split > 0 ? g.a2_split : (g.kb_split > 0 ? g.kb_split : num_kb);
            for (int kb = 0; kb < kbs; ++kb) tma_prefetch_l2_3d(&tmA, kb * BK, pr0, pf.clip);
            if (g.a2_split > 0)
              for (int kb = g.a2_split; kb < num_kb; ++kb) tma_prefetch_l2_3d(&tmR, (kb - g.a2_split) * BK, pr0, pf.clip);
          }
          pf.next(g);
        }
        const int r0 = tc.mi * rows_out - halo;   // may be negative: zero fill
        const int n0 = tc.nt * g.block_n;
        // (An L2 prefetch of the tile's residual rows from here - cp.async.bulk.prefetch.tensor through
        // tmR - was measured: +30 % DRAM reads on the residual kernels for no gain in time.)
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (kb == 0) WV_DBG(0, dbg_it);
          mbar_arrive_expect_tx(&full[stage], stage_bytes);
          // k-block -> (A tensor, column): [tmA | tmR from a2_split | tmA again from a3_split]
          const CUtensorMap* mp = &tmA;
          int kk = kb;
          if (g.a3_split > 0 && kb >= g.a3_split) kk = kb - g.a3_split;
          else if (g.a2_split > 0 && kb >= g.a2_split) { mp = &tmR; kk = kb - g.a2_split; }
          if (g.phases > 1) tma_load_4d(smemA + stage * A_STAGE_BYTES, mp, &full[stage], kk * BK, r0, tc.clip % g.phases, tc.clip / g.phases);
          else if (kk != kb) tma_load_3d(smemA + stage * A_STAGE_BYTES, mp, &full[stage], kk * BK, r0, tc.clip);
          else if (g.kb_split > 0 && kb >= g.kb_split) tma_load_3d(smemA + stage * A_STAGE_BYTES, &tmA, &full[stage], (kb - g.kb_split) * BK, r0 - 1, tc.clip);
          else if (g.a_evict_first) tma_load_3d_hint(smemA + stage * A_STAGE_BYTES, &tmA, &full[stage], kb * BK, r0, tc.clip, pol);
          else tma_load_3d(smemA + stage * A_STAGE_BYTES, &tmA, &full[stage], kb * BK, r0, tc.clip);
          if (!g.resident_b) tma_load_2d(smemB + stage * b_stage_bytes, &tmB, &full[stage], kb * BK, n0);
          if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
      }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if constexpr (STG) reg_dealloc<REGS_LIGHT>();
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t as_phase = 0;
      int dbg_it = 0;
      if (g.resident_b) mbar_wait(w_full, 0);
      if constexpr (CG2) {
        if ((blockIdx.x & 1) == 0) {                 // the even CTA issues for the pair; commits arrive in both CTAs
          for (int ku = 0;; ++ku) {
            bool done;
            seq_tile<CG2>(g, 2 * ku, done);
            if (done) break;
            mbar_wait(&acc_empty[as], as_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * acc_cols);
            for (int kb = 0; kb < num_kb; ++kb) {
              mbar_wait(&full[stage], phase);
              tc_fence_after();
              const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smemA + stage * A_STAGE_BYTES));
              const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smemB + stage * b_stage_bytes));
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_f16_cg2(tmem_d, adesc + 2 * k, bdesc + 2 * k, g.idesc2, (kb | k) ? 1u : 0u);
              umma_commit_cg2(&empty[stage]);
              if (kb == num_kb - 1) umma_commit_cg2(&acc_full[as]);
              if (++stage == g.stages) { stage = 0; phase ^= 1; }
            }
            if (++as == acc_stages) { as = 0; as_phase ^= 1; }
          }
        }
      } else if (g.pair) {
        int c = 0;                                   // tiles issued so far: accumulator stage c % acc_stages
        for (int ku = 0;; ++ku) {
          bool done, d1;
          seq_tile<CG2>(g, 2 * ku, done);
          if (done) break;
          const int nv = seq_tile<CG2>(g, 2 * ku + 1, d1) >= 0 ? 2 : 1;
          const int s0 = c % acc_stages, s1 = (c + 1) % acc_stages;
          mbar_wait(&acc_empty[s0], ((c / acc_stages) & 1) ^ 1);
          if (nv == 2) mbar_wait(&acc_empty[s1], (((c + 1) / acc_stages) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d0 = tmem_base + static_cast<uint32_t>(s0 * acc_cols), d1t = tmem_base + static_cast<uint32_t>(s1 * acc_cols);
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint64_t adesc0 = make_sw128_kmajor_desc(smem_u32(smemA + stage * a_stage_bytes));
            const uint64_t adesc1 = make_sw128_kmajor_desc(smem_u32(smemA + stage * a_stage_bytes + A_STAGE_BYTES));
            const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smemB + stage * b_stage_bytes));
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_f16(d0, adesc0 + 2 * k, bdesc + 2 * k, g.idesc, (kb | k) ? 1u : 0u);
            if (nv == 2) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) umma_f16(d1t, adesc1 + 2 * k, bdesc + 2 * k, g.idesc, (kb | k) ? 1u : 0u);
            }
            umma_commit(&empty[stage]);
            if (kb == num_kb - 1) {
              umma_commit(&acc_full[s0]);
              if (nv == 2) umma_commit(&acc_full[s1]);
            }
            if (++stage == g.stages) { stage = 0; phase ^= 1; }
          }
          c += nv;
        }
      } else
      for (int k_ = 0;; ++k_, ++dbg_it) {
        bool done;
        const int l_ = seq_tile<CG2>(g, k_, done);
        if (done) break;
        if (l_ < 0) continue;
        mbar_wait(&acc_empty[as], as_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(as * acc_cols);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full[stage], phase);
          if (kb == num_kb - 1) WV_DBG(1, dbg_it);
          tc_fence_after();
          const uint64_t adesc = make_sw128_kmajor_desc(smem_u32(smemA + stage * A_STAGE_BYTES));
          const uint64_t bdesc = make_sw128_kmajor_desc(smem_u32(smemB + (g.resident_b ? kb : stage) * b_stage_bytes));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 B per K=16 step inside the 128 B swizzle row -> +2 in the (addr>>4) field
            umma_f16(tmem_d, adesc + 2 * k, bdesc + 2 * k, g.idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (kb == num_kb - 1) { umma_commit(&acc_full[as]); WV_DBG(2, dbg_it); }
          if (++stage == g.stages) { stage = 0; phase ^= 1; }
        }
        if (++as == acc_stages) { as = 0; as_phase ^= 1; }
      }
    }
  } else if (STG && warp < 4) {
    reg_dealloc<REGS_LIGHT>();   // warps 2 (TMEM allocator) and 3: the whole warpgroup must take part
    if constexpr (!PM && !CG2) {
      if (warp == 3 && lane == 0 && g.res_tma) {
        // residual producer: the [128 rows x block_n] residual tile of every tile of this CTA, one tile ahead of the math warps
        const int pitch_ = staged_pitch_bytes(g.block_n);
        uint8_t* res_tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(stage_tiles + g.stage_bufs * BM * pitch_) + 127) & ~static_cast<uintptr_t>(127));
        const uint32_t bytes = static_cast<uint32_t>(BM * g.block_n * 2);
        int i = 0;
        for (TileWalker<CG2> tc(g); tc.tile < g.num_tiles; tc.next(g), ++i) {
          const int buf = i & 1;
          mbar_wait(&bars[RES_BAR_INDEX + 2 + buf], (static_cast<uint32_t>(i >> 1) & 1u) ^ 1u);
          mbar_arrive_expect_tx(&bars[RES_BAR_INDEX + buf], bytes);
          tma_load_3d(res_tiles + buf * bytes, &tmR, &bars[RES_BAR_INDEX + buf], tc.nt * g.block_n, tc.mi * rows_out, tc.clip);
        }
      }
    }
  } else if (STG && warp >= 4) {
    const int pitch = staged_pitch_bytes(g.block_n, PM);
    const int chunks = g.block_n / 32;
    if (warp < 4 + P1_WARPS) {
      // ---------------------------------------------------------- drain warps: TMEM -> fp16 -> smem
      reg_dealloc<REGS_DRAIN>();
      const int q = warp - 4;   // == warp % 4: TMEM lane quarter this warp may touch
      const int drain_bar_threads = P1_WARPS * 32 + (g.math_groups == 2 ? P2_THREADS / 2 : P2_THREADS);
      const uint32_t stage_u32 = smem_u32(stage_tiles);
      int as = 0, sb = 0, it = 0;
      uint32_t as_phase = 0;
      uint32_t v[32];
      for (int k_ = 0;; ++k_) {
        bool done;
        const int l_ = seq_tile<CG2>(g, k_, done);
        if (done) break;
        if (l_ < 0) continue;
        if (it >= g.stage_bufs) named_bar_sync(BAR_ST_EMPTY + sb, drain_bar_threads);   // math warps left tile sb
        const int sub = CG2 ? (it & 1) : 0;        // CTA-pair mode: the accumulator holds two block_n-wide tiles
        if (sub == 0) mbar_wait(&acc_full[as], as_phase);
        if (q == 0 && lane == 0) WV_DBG(3, it);
        if (lane == 0) WV_DBG(36 + q, it);           // per drain warp: start
        tc_fence_after();
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * acc_cols + sub * g.block_n);
        const uint32_t rowp = stage_u32 + sb * (BM * pitch) + (q * 32 + lane) * pitch;
        for (int c = 0; c < (WV_DBG_MODE(4) ? 0 : chunks); ++c) {
          tmem_ld32(taddr + c * 32, v);
          tmem_ld_wait();
          if constexpr (PM) {  // precise mode: the accumulator tile is staged as fp32
#pragma unroll
            for (int i = 0; i < 8; ++i) sts_u4(rowp + c * 128 + i * 16, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            continue;
          }
          if (g.last_mode) {   // output conv: the five tap columns stay fp32 (32 bytes per row)
            sts_u4(rowp, v[0], v[1], v[2], v[3]);
            sts_u4(rowp + 16, v[4], v[5], v[6], v[7]);
            continue;
          }
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
        if constexpr (!CG2) {
          if (lane == 0) mbar_arrive(&acc_empty[as]);   // TMEM stage free: the MMAs of tile i+2 may start
        } else if (sub == 1) {
          if (lane == 0) mbar_arrive_leader(&acc_empty[as]);   // both halves drained: tell the issuing CTA
        }
        if (q == 0 && lane == 0) WV_DBG(4, it);
        if (lane == 0) WV_DBG(8 + q, it);            // per drain warp: end
        named_bar_arrive(BAR_ST_FULL + sb, drain_bar_threads);   // release: this warp's 32 rows are staged
        if (!CG2 || sub == 1) { if (++as == acc_stages) { as = 0; as_phase ^= 1; } }
        if (++sb == g.stage_bufs) sb = 0;
        ++it;
      }
    } else {
      // ---------------------------------------------------------- math warps: smem -> epilogue -> global
      reg_alloc<REGS_MATH>();
      if constexpr (PM) {
        if (g.down_r > 0) {
          switch (g.down_r) {
            case 2: pm_down_loop<2>(g, stage_tiles, down_w); break;
            case 4: pm_down_loop<4>(g, stage_tiles, down_w); break;
            case 5: pm_down_loop<5>(g, stage_tiles, down_w); break;
            default: pm_down_loop<8>(g, stage_tiles, down_w); break;
          }
        } else if (g.taps == 5) pm_math_loop<5>(g, stage_tiles);
        else pm_math_loop<1>(g, stage_tiles);
      } else
      if (g.last_mode) {
        staged_last_loop<CG2>(g, stage_tiles);
      } else if (g.down_r > 0) {
        switch (g.down_r) {
          case 2: staged_down_loop<2, CG2>(g, stage_tiles, down_w, lane); break;
          case 4: staged_down_loop<4, CG2>(g, stage_tiles, down_w, lane); break;
          case 5: staged_down_loop<5, CG2>(g, stage_tiles, down_w, lane); break;
          default: staged_down_loop<8, CG2>(g, stage_tiles, down_w, lane); break;
        }
      } else if (g.taps == 5) staged_math_rows<5, CG2>(g, stage_tiles, lane);
      else staged_math_rows<1, CG2>(g, stage_tiles, lane);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (16 warps)
    const int e = warp - 4;
    const int q = e & 3;    // TMEM lane quarter: warp (w % 4) may touch lanes [32q, 32q+32)
    // STFT: the 16 epilogue warps may form 2 or 4 groups (of 8 / 4 warps: every group covers the four TMEM lane quarters);
    // group g owns accumulator stage g = every epi_groups-th tile, so that several tiles are in the epilogue at once
    const int n_groups = IS_STFT ? g.epi_groups : 1;
    const bool two_groups = n_groups > 1;
    const int gw = EPI_WARPS / n_groups;                 // warps per group
    const int h = two_groups ? (e % gw) >> 2 : e >> 2;   // column split: this warp takes chunks c with c % split == h
    const int split = two_groups ? gw >> 2 : EPI_SPLIT;
    const int grp = e / gw;
    int as = 0;
    uint32_t as_phase = 0;
    for (int k_ = 0;; ++k_) {
      bool done;
      const int l_ = seq_tile<CG2>(g, k_, done);
      if (done) break;
      if (l_ < 0) continue;
      if (two_groups && as != grp) {                     // another group's tile (acc_stages == epi_groups: stage = tile index mod groups)
        if (++as == acc_stages) { as = 0; as_phase ^= 1; }
        continue;
      }
      const TileCoord tc = tile_coord(g, l_);
      const int nt = tc.nt, clip = tc.clip;
      const int r_base = tc.mi * rows_out;   // first OUTPUT row of the tile
      const int n0 = nt * g.block_n;
      const int chunks = g.block_n / 32;
      mbar_wait(&acc_full[as], as_phase);
      tc_fence_after();
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(as * acc_cols);
      uint32_t v[32];

      {
        const int r = r_base + q * 32 + lane;
        bool row_ok = r < g.rows_per_clip;
        long long m = static_cast<long long>(clip) * g.rows_per_clip + r;
        if (IS_STFT && g.phases > 1) {      // row q of (clip, phase) is frame q*phases + phase of the clip
          const int f = r * g.phases + clip % g.phases;
          row_ok = row_ok && f < g.frames_per_clip;
          m = static_cast<long long>(clip / g.phases) * g.frames_per_clip + f;
        }
        if constexpr (IS_L2N) {
          if (h == 0) {
            float ss = 0.f;
            for (int c = 0; c < chunks; ++c) {
              tmem_ld32(taddr + c * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const float x = __uint_as_float(v[i]) + (g.bias ? __ldg(g.bias + n0 + c * 32 + i) : 0.f);
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
                if constexpr (PM) {   // split latent [hi | lo], row pitch ldo, lo at + lo_off
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    st_split4(g.out_raw + m * g.ldo + n + 4 * i, g.lo_off, f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                } else {
                uint4* op = reinterpret_cast<uint4*>(g.out_raw + m * g.ldo + n);
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  op[i] = make_uint4(pack_act2(f[8 * i], f[8 * i + 1]),
                                     pack_act2(f[8 * i + 2], f[8 * i + 3]),
                                     pack_act2(f[8 * i + 4], f[8 * i + 5]),
                                     pack_act2(f[8 * i + 6], f[8 * i + 7]));
                }
                if (g.out_f32_t != nullptr) {
                  const long long b = m / g.f32_F, fr = m % g.f32_F;
                  float* tp = g.out_f32_t + (b * g.N + n) * g.f32_F + fr;
#pragma unroll
                  for (int i = 0; i < 32; ++i) tp[static_cast<long long>(i) * g.f32_F] = f[i];
                }
              }
            }
          }
        } else if constexpr (EPI == EPI_STFT_PM) {
          // log-magnitude -> split pairs [hi | lo].  __logf = MUFU.LG2 * ln 2: absolute error ~1e-7 on ln, i.e. 2e-8 on the
          // normalised value (an fp32 rounding of an O(1) quantity is 6e-8); the libdevice logf cost ~20 instructions per bin
          // and made this epilogue the bound of the precise STFT launches
          for (int c = h; c < chunks; c += split) {
            const int p0 = (n0 + c * 32) >> 1;
            tmem_ld32(taddr + c * 32, v);
            tmem_ld_wait();
            if (row_ok) {
              float y[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float re = __uint_as_float(v[2 * i]), im = __uint_as_float(v[2 * i + 1]);
                y[i] = (0.5f * __logf(fmaxf(re * re + im * im, g.clamp_sq)) - g.log_offset) * g.inv_sigma;
              }
              act_t* yp = g.out_raw + m * g.ldo;
              if (p0 == 0) {
                const float re0 = __uint_as_float(v[0]), ren = __uint_as_float(v[1]);
                y[0] = (0.5f * __logf(fmaxf(re0 * re0, g.clamp_sq)) - g.log_offset) * g.inv_sigma;
                const float yn = (0.5f * __logf(fmaxf(ren * ren, g.clamp_sq)) - g.log_offset) * g.inv_sigma;
                uint32_t hn, ln;
                split2(yn, 0.f, hn, ln);   // bin N/2 + 7 zero pad columns (row half-width = n_half + 8)
                *reinterpret_cast<uint4*>(yp + g.n_half) = make_uint4(hn, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(yp + g.lo_off + g.n_half) = make_uint4(ln, 0u, 0u, 0u);
              }
#pragma unroll
              for (int i = 0; i < 4; ++i)
                st_split4(yp + p0 + 4 * i, g.lo_off, y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
            }
          }
        } else if constexpr (EPI == EPI_STFT) {
          for (int c = h; c < chunks; c += split) {
            const int p0 = (n0 + c * 32) >> 1;
            tmem_ld32(taddr + c * 32, v);
            tmem_ld_wait();
            if (row_ok) {
              float y[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const float re = __uint_as_float(v[2 * i]), im = __uint_as_float(v[2 * i + 1]);
                y[i] = (0.5f * __logf(fmaxf(re * re + im * im, g.clamp_sq)) - g.log_offset) * g.inv_sigma;
              }
              act_t* yp = g.out_raw + m * g.ldo;
              if (p0 == 0) {
                const float re0 = __uint_as_float(v[0]), ren = __uint_as_float(v[1]);
                y[0] = (0.5f * __logf(fmaxf(re0 * re0, g.clamp_sq)) - g.log_offset) * g.inv_sigma;
                const float yn = (0.5f * __logf(fmaxf(ren * ren, g.clamp_sq)) - g.log_offset) * g.inv_sigma;
                // bin N/2 plus the zero pad columns up to the row pitch (ldo = n_half + 8 or + 16)
                *reinterpret_cast<uint4*>(yp + g.n_half) = make_uint4(pack_act2(yn, 0.f), 0u, 0u, 0u);
                if (g.ldo - g.n_half > 8) *reinterpret_cast<uint4*>(yp + g.n_half + 8) = make_uint4(0u, 0u, 0u, 0u);
              }
              uint4* op = reinterpret_cast<uint4*>(yp + p0);
#pragma unroll
              for (int i = 0; i < 2; ++i)
                op[i] = make_uint4(pack_act2(y[8 * i], y[8 * i + 1]),
                                   pack_act2(y[8 * i + 2], y[8 * i + 3]),
                                   pack_act2(y[8 * i + 4], y[8 * i + 5]),
                                   pack_act2(y[8 * i + 6], y[8 * i + 7]));
            }
          }
        } else {  // EPI_HEAD
          const long long b = m / g.head_F;
          const int fr = static_cast<int>(m % g.head_F);
          const int o = n0 / g.hop;
          const int j0 = n0 % g.hop;
          const float bo = g.bias ? __ldg(g.bias + o) : 0.f;
          float psum = 0.f;
          for (int c = h; c < chunks; c += EPI_SPLIT) {
            tmem_ld32(taddr + c * 32, v);
            tmem_ld_wait();
            if (row_ok) {
              const int t0 = fr * g.hop + j0 + c * 32;
              const long long base = b * g.T + t0;
              float* lp = g.logits ? g.logits + (b * g.n_out + o) * static_cast<long long>(g.T) + t0
                                   : nullptr;
              if (t0 + 32 <= g.T && ((g.T | t0) & 15) == 0) {
                // whole 32-sample run inside the clip and every row 16-byte aligned: the thread's
                // 128 B of logits / probabilities leave as eight 16-byte stores, its 32 mask bytes as two
                float l[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) l[i] = __uint_as_float(v[i]) + bo;
                if (lp) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) reinterpret_cast<float4*>(lp)[i] = make_float4(l[4 * i], l[4 * i + 1], l[4 * i + 2], l[4 * i + 3]);
                }
                if (g.mask_out) {
                  uint32_t mw[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    mw[i] = (l[4 * i] > 0.5f ? 1u : 0u) | (l[4 * i + 1] > 0.5f ? 0x100u : 0u) | (l[4 * i + 2] > 0.5f ? 0x10000u : 0u) |
                            (l[4 * i + 3] > 0.5f ? 0x1000000u : 0u);
                  uint4* mp = reinterpret_cast<uint4*>(g.mask_out + base);
                  mp[0] = make_uint4(mw[0], mw[1], mw[2], mw[3]);
                  mp[1] = make_uint4(mw[4], mw[5], mw[6], mw[7]);
                }
                if (g.probs || g.partial) {
                  uint4 pw[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
                  if (g.partial && g.presence) { pw[0] = __ldcg(reinterpret_cast<const uint4*>(g.presence + base)); pw[1] = __ldcg(reinterpret_cast<const uint4*>(g.presence + base) + 1); }
                  const uint32_t pwv[8] = {pw[0].x, pw[0].y, pw[0].z, pw[0].w, pw[1].x, pw[1].y, pw[1].z, pw[1].w};
#pragma unroll
                  for (int i = 0; i < 32; ++i) {
                    const float p = __fdividef(1.f, 1.f + __expf(-l[i]));   // MUFU.RCP: 2 ulp, far inside the 3e-4 bit margin
                    l[i] = p;
                    if (g.partial) psum += g.presence ? (((pwv[i >> 2] >> ((i & 3) * 8)) & 0xffu) ? p : 0.f) : p;
                  }
                  if (g.probs) {
                    float4* pp = reinterpret_cast<float4*>(g.probs + base);
#pragma unroll
                    for (int i = 0; i < 8; ++i) pp[i] = make_float4(l[4 * i], l[4 * i + 1], l[4 * i + 2], l[4 * i + 3]);
                  }
                }
              } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                if (t0 + i < g.T) {
                  const float l = __uint_as_float(v[i]) + bo;
                  const float p = __fdividef(1.f, 1.f + __expf(-l));   // MUFU.RCP: 2 ulp, far inside the 3e-4 bit margin
                  if (lp) lp[i] = l;
                  if (g.mask_out) g.mask_out[base + i] = l > 0.5f ? 1 : 0;
                  if (g.probs) g.probs[base + i] = p;
                  if (g.partial) psum += g.presence ? (__ldcg(g.presence + base + i) ? p : 0.f) : p;
                }
              }
              }
            }
          }
          if (g.partial != nullptr && row_ok) g.partial[(m * g.tiles_n + nt) * EPI_SPLIT + h] = psum;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc_empty[as]);
      }
      if (++as == acc_stages) { as = 0; as_phase ^= 1; }
    }
  }

  tc_fence_before();
  if constexpr (CG2) cluster_sync_all();   // the leader's MMAs read the peer's shared memory and both signal each other's barriers
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (CG2) tmem_dealloc_cg2(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace wv
