// Host side of the C ABI (include/wv_b200.h): weight folding/packing, TMA descriptors, the
// per-(B,T) launch plan with an arena-planned workspace, and the forward entry points.
// sm_100a only; every compute step is a kernel from gemm_sm100.cuh / glue_kernels.cuh.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../include/wv_b200.h"
#include "gemm_sm100.cuh"
#include "glue_kernels.cuh"
#include "validation_kernels.cuh"
#include "resblock_sm100.cuh"

namespace {

using namespace wv;
typedef act_t h16;   // fp16 activations / weights (ptx_sm100.cuh)

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
template <typename... Args>
int failf(int code, const char* fmt, Args... args) {
  char b[512];
  snprintf(b, sizeof(b), fmt, args...);
  return fail(code, std::string(b));
}
struct WvError {
  int code;
  std::string msg;
};
#define WV_THROW(code, ...)                           \
  do {                                                \
    char _b[512];                                     \
    snprintf(_b, sizeof(_b), __VA_ARGS__);            \
    throw WvError{code, std::string(_b)};             \
  } while (0)
#define CK(expr)                                                                      \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess)                                                            \
      WV_THROW(WV_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
               __LINE__);                                                             \
  } while (0)

constexpr float WAV_STD = 0.1122080159f;                                   // seanet.py:631
const float SPEC_MEANS[5] = {-4.554f, -4.315f, -4.021f, -3.726f, -3.477f};  // seanet.py:632
const float SPEC_STDS[5] = {2.830f, 2.837f, 2.817f, 2.796f, 2.871f};        // seanet.py:633
constexpr float WAV_FP16_SCALE = 64.f;  // keeps quiet audio out of fp16 subnormals (exact 2^6)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
  if (q != cudaDriverEntryPointSuccess || !p)
    WV_THROW(WV_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  return fn;
}

// 16-bit tensor map over (k, row, clip): element strides in ELEMENTS for row / clip.
CUtensorMap make_tmap(const void* base, int rank, uint64_t dim_k, uint64_t dim_row, uint64_t dim_clip,
                      uint64_t row_stride, uint64_t clip_stride, int box_k, int box_row, bool /*fp16*/,
                      bool swizzle = true) {
  CUtensorMap m;
  cuuint64_t dims[3] = {dim_k, dim_row, dim_clip};
  cuuint64_t strides[2] = {row_stride * 2, clip_stride * 2};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_k), static_cast<cuuint32_t>(box_row), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (rank == 3 && (strides[1] & 15)))
    WV_THROW(WV_ERR_INVALID, "TMA alignment violated (base %p, strides %llu %llu)", base,
             (unsigned long long)strides[0], (unsigned long long)strides[1]);
  CUresult r = get_encode_fn()(
      &m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank,   // every 16-bit tensor of the path is fp16
      const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) WV_THROW(WV_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return m;
}

// fp16 4-D map (k, row, phase, clip) for the phased STFT frame view; strides in ELEMENTS.
CUtensorMap make_tmap4(const void* base, uint64_t dim_k, uint64_t dim_row, uint64_t dim_phase, uint64_t dim_clip,
                       uint64_t row_stride, uint64_t phase_stride, uint64_t clip_stride, int box_k, int box_row) {
  CUtensorMap m;
  cuuint64_t dims[4] = {dim_k, dim_row, dim_phase, dim_clip};
  cuuint64_t strides[3] = {row_stride * 2, phase_stride * 2, clip_stride * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_k), static_cast<cuuint32_t>(box_row), 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (strides[0] & 15) || (strides[1] & 15) || (strides[2] & 15))
    WV_THROW(WV_ERR_INVALID, "TMA alignment violated (4-D frame view)");
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) WV_THROW(WV_ERR_CUDA, "cuTensorMapEncodeTiled (4-D) failed with %d", (int)r);
  return m;
}

bool g_cg2_bn96 = true;     // see pick_block_n (WV_CG2_BN96=0: keep 128-wide tiles for N = 384)
int pick_block_n(int N, int must_divide = 0, int max_bn = MAX_BN, bool pairable = false) {
  static const int cands[] = {256, 192, 160, 128, 96, 64, 32};
  // STAGED layers whose 128-wide n tiles come in an odd number (N = 384): 96-wide tiles pair up for the CTA-pair mode.
  // Only for the compute-bound uses (resblock first halves, channel-halving 1x1): the residual launches are DRAM-bound
  // and measured 14 % slower with 96-wide tiles (profiles/r02_cta_pairs.md)
  if (pairable && max_bn == 128 && must_divide == 0 && N >= 384 && N % 128 == 0 && (N / 128) % 2 == 1 && N % 192 == 0 && g_cg2_bn96) return 96;
  for (int c : cands)
    if (c <= max_bn && N % c == 0 && (must_divide == 0 || must_divide % c == 0)) return c;
  WV_THROW(WV_ERR_UNSUPPORTED, "no tile width for N=%d", N);
}

int g_num_sms = 0;
long long g_graph_max_samples = 128 * 16000;  // calls of up to this many samples run as one CUDA graph (WV_GRAPH_MAX_SAMPLES, 0 = off);
                                               // measured r02 on the 64 x 1 s step: +1.5 % (launch gaps), I/O staging copies included
int g_ldy_align = 8;        // log-spectrogram row pitch in elements (WV_LDY_ALIGN: 8 = 16 B, 16 = 32 B = one DRAM sector per chunk)
unsigned g_rows6_mask = (1u << (64 / 32)) | (1u << (128 / 32));   // STAGED tiles of these widths (bit = channels / 32) use 6-row math units:
                            // 4-row groups leave their last pass mostly idle there (WV_ROWS6_BN = comma list of widths, 0 = off)
int g_math_groups = 4;      // STAGED math warps as two groups of six, one per staging tile: 1 = never, 2 = always, 3 = launches without a
                            // residual, 4 = per launch class as measured (WV_MATH_GROUPS)
int g_a_prefetch = 0;       // STAGED / STFT producers prefetch the A rows of the tile this many tiles ahead into L2 (WV_A_PREFETCH, 0 = off)
bool g_pre_fuse = false;    // the first encoder resblock recomputes its residual (= conv_pre output) from the waveform instead of reading it
                            // (WV_PRE_FUSE=1; bit-identical; measured r01h: conv_pre 68 -> 61 us, r0.out 113 -> 121 us, +-0 overall: off)
int g_res1_kb = 0;          // largest W tile (KB) kept resident next to a single staging tile (WV_RES1_KB; 0 = off)
bool g_res_tma = true;      // residual tiles of the resblock second halves by TMA into shared memory (WV_RES_TMA=0: per-thread ld.global;
                            // see gemm_sm100.cuh res_tma; measured r02: dec.u2 / u3 second halves -10..-13 %)
int g_res_tma_min_stages = 4;   // fewest A ring stages left next to the residual buffers (WV_RES_TMA_MIN_STAGES)
bool g_res_early2 = true;   // residual rows of the second unit of a tile requested before the drain hand-off too (WV_RES_EARLY2=0: after it)
int g_spec_fuse_maxc = 128; // encoder stages up to this width run the last resblock's second half and the spectrogram 1x1 as ONE launch (WV_SPEC_FUSE_MAXC, 0 = off)
bool g_last_gemm = true;    // decoder output conv (C -> 1, k = 5) on the tensor cores (WV_LAST_GEMM=0: CUDA-core kernel)
int g_epi_groups = 2;       // STFT epilogue warp groups, one accumulator stage each (WV_EPI_GROUPS: 0/1 = one group of 16 warps, 2 = two groups for
                            // tiles of <= 64 columns, 4 = additionally four groups for tiles of <= 128 columns)
int g_cg2_min_kb = 4;       // STAGED layers with >= this many k-blocks, streamed W and an even number of n tiles run as CTA pairs
                            // (cluster of 2, tcgen05 cta_group::2, 256-row MMAs; WV_CG2_MIN_KB, 0 = off)
int g_max_clusters = 0;     // co-resident CTA pairs of the STAGED kernel (cudaOccupancyMaxActiveClusters)
int g_pair_min_kb = 4;      // STAGED layers with >= this many k-blocks and streamed W run two M tiles per W k-block (WV_PAIR_MIN_KB, 0 = off)
int g_one_buf_kb = 0;       // STAGED layers with >= this many k-blocks and streamed W use one staging tile (WV_ONE_BUF_KB, 0 = off)
int g_up_fuse_maxc = 384;   // decoder stages up to this input width run upsample + 1x1 as one GEMM (WV_UP_FUSE_MAXC, 0 = off)
bool g_phased_stft = true;  // hop < 8 STFTs read frames through phased strided TMA views (WV_PHASED_STFT=0: frame matrix)
bool g_evict_first = true;  // A operand TMA loads carry an L2 evict_first hint (WV_EVICT_FIRST=0 disables)
bool g_serpentine = true;   // consecutive GEMM launches walk their tiles in opposite directions (WV_SERPENTINE=0 disables)
int g_rb_maxc = 0;     // widest resblock that runs as ONE fused kernel (resblock_sm100.cuh, WV_RB_MAXC=96 enables it);
                       // measured r01: correct, but bound by its three ELU passes (MUFU) - 280 vs 265 us at C=96
bool g_pdl = false;  // programmatic dependent launch for every plan kernel (WV_PDL=1 enables; measured
                     // 1-2 % slower on the 64-clip batch, where launch gaps are already hidden)
thread_local bool g_pm = false;   // weights are being built for a PRECISE net (split-fp16 operands, see gemm_sm100.cuh)
// Per-device initialisation (kernel attributes belong to the device's context); guarded for concurrent first use.
void init_device_once() {
  static std::mutex mu;
  static std::set<int> inited;
  std::lock_guard<std::mutex> lock(mu);
  int dev = 0;
  CK(cudaGetDevice(&dev));
  if (inited.count(dev)) return;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    WV_THROW(WV_ERR_UNSUPPORTED, "this library targets sm_100a (B200); device is sm_%d%d",
             prop.major, prop.minor);
  g_num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("WV_PDL")) g_pdl = atoi(e) != 0;   // A/B switch for the profiling scripts
  CK(cudaFuncSetAttribute(gemm_sm100_kernel<EPI_STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_LIMIT));
  CK(cudaFuncSetAttribute(gemm_sm100_kernel<EPI_STAGED, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_LIMIT));
  CK(cudaFuncSetAttribute(gemm_sm100_kernel<EPI_L2NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_LIMIT));
  CK(cudaFuncSetAttribute(gemm_sm100_kernel<EPI_STFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_LIMIT));
  CK(cudaFuncSetAttribute(gemm_sm100_kernel<EPI_HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_LIMIT));
  CK(cudaFuncSetAttribute(gemm_sm100_kernel<EPI_STAGED_PM>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_LIMIT));
  CK(cudaFuncSetAttribute(gemm_sm100_kernel<EPI_L2NORM_PM>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_LIMIT));
  CK(cudaFuncSetAttribute(gemm_sm100_kernel<EPI_STFT_PM>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_LIMIT));
  CK(cudaFuncSetAttribute(resblock_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_LIMIT));
  if (const char* e = getenv("WV_GRAPH_MAX_SAMPLES")) g_graph_max_samples = atoll(e);
  if (const char* e = getenv("WV_LDY_ALIGN")) g_ldy_align = atoi(e);
  if (const char* e = getenv("WV_SPEC_FUSE_MAXC")) g_spec_fuse_maxc = atoi(e);
  if (const char* e = getenv("WV_RES_EARLY2")) g_res_early2 = atoi(e) != 0;
  if (const char* e = getenv("WV_RES_TMA")) g_res_tma = atoi(e) != 0;
  if (const char* e = getenv("WV_RES_TMA_MIN_STAGES")) g_res_tma_min_stages = std::max(2, atoi(e));
  if (const char* e = getenv("WV_RES1_KB")) g_res1_kb = atoi(e);
  if (const char* e = getenv("WV_PRE_FUSE")) g_pre_fuse = atoi(e) != 0;
  if (const char* e = getenv("WV_A_PREFETCH")) g_a_prefetch = atoi(e);
  if (const char* e = getenv("WV_MATH_GROUPS")) g_math_groups = atoi(e);
  if (const char* e = getenv("WV_ROWS6_BN")) {
    g_rows6_mask = 0;
    for (const char* q = e; *q;) {
      const int v = atoi(q);
      if (v >= 32 && v <= 256) g_rows6_mask |= 1u << (v / 32);
      while (*q && *q != ',') ++q;
      if (*q == ',') ++q;
    }
  }
  if (const char* e = getenv("WV_LAST_GEMM")) g_last_gemm = atoi(e) != 0;
  if (const char* e = getenv("WV_EPI_GROUPS")) g_epi_groups = atoi(e);
  if (const char* e = getenv("WV_PAIR_MIN_KB")) g_pair_min_kb = atoi(e);
  if (const char* e = getenv("WV_CG2_MIN_KB")) g_cg2_min_kb = atoi(e);
  if (const char* e = getenv("WV_CG2_BN96")) g_cg2_bn96 = atoi(e) != 0;
  {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * g_num_sms); cfg.blockDim = dim3(STAGED_THREADS); cfg.dynamicSmemBytes = GEMM_SMEM_LIMIT;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, gemm_sm100_kernel<EPI_STAGED, true>, &cfg) == cudaSuccess) g_max_clusters = nc;
    else { cudaGetLastError(); g_max_clusters = 0; }
    if (getenv("WV_VERBOSE")) fprintf(stderr, "[wv] %d SMs, %d co-resident CTA pairs\n", g_num_sms, g_max_clusters);
  }
  if (const char* e = getenv("WV_ONE_BUF_KB")) g_one_buf_kb = atoi(e);
  if (const char* e = getenv("WV_UP_FUSE_MAXC")) g_up_fuse_maxc = atoi(e);
  if (const char* e = getenv("WV_PHASED_STFT")) g_phased_stft = atoi(e) != 0;
  if (const char* e = getenv("WV_EVICT_FIRST")) g_evict_first = atoi(e) != 0;
  if (const char* e = getenv("WV_SERPENTINE")) g_serpentine = atoi(e) != 0;
  if (const char* e = getenv("WV_RB_MAXC")) g_rb_maxc = atoi(e);   // 0 disables the fused resblock kernel
  CK(cudaFuncSetAttribute(conv_last_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  inited.insert(dev);
}

// ------------------------------------------------------------------------------------------
// weights
struct HostTensor {
  const float* data;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

struct DevMem {
  std::vector<void*> ptrs;
  ~DevMem() {
    for (void* p : ptrs) cudaFree(p);
  }
  void* alloc(size_t bytes) {
    void* p = nullptr;
    CK(cudaMalloc(&p, std::max<size_t>(bytes, 16)));
    ptrs.push_back(p);
    return p;
  }
  template <typename T>
  T* upload(const std::vector<T>& h) {
    T* d = static_cast<T*>(alloc(h.size() * sizeof(T)));
    CK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
  }
};

struct GemmW {          // B operand of a GEMM: 16-bit [N, ldw] K-major, zero padded
  void* w = nullptr;
  float* bias = nullptr;
  int N = 0, K = 0, ldw = 0, block_n = 0;
  bool fp16 = true;
  CUtensorMap tm;
  // precise nets: K = the split contraction length; a2 / a3 = first k-block of the 2nd / 3rd operand segment
  int a2 = 0, a3 = 0, K_alg = 0;
};
struct DwW {            // depthwise taps fp32 [k][C]
  float* w = nullptr;
  float* bias = nullptr;
  int k = 0, C = 0;
};
struct ResW {
  GemmW pw1, pw2;
  DwW dw1, dw2;          // dw2 carries RS * res_scale_param
  float pre_scale;
};
struct SpecW {
  GemmW dft;             // fp16 [n_fft, n_fft] (re,im) interleaved rows
  GemmW layer;           // h16 [C, n_fft/2+1] * RS * scale_param
  int n_fft, hop;
  float mean, stdv;
};
struct EncStageW {
  std::vector<ResW> res;
  SpecW spec;
  GemmW dual;            // [W2 | 0 ; 0 | Wspec] per n tile: last resblock's second 1x1 and the spectrogram 1x1 in one GEMM
  int dual_split_kb = 0; // k-blocks of the first operand (h1); 0 = not built
  int dual_ldy = 0;      // log-spectrogram row pitch the weights were laid out for
  GemmW down_pw;
  DwW down_dw;
  int r, C;
};
struct EncoderW {
  DwW conv_pre;          // [5][C0], 1/wav_std folded
  std::vector<EncStageW> stages;
  SpecW spec_post;
  DwW post_dw;
  GemmW post_pw;
  int C0, C_last, dim, hop, nfft_max;
  bool has_film = false;
  FilmArgs film;
  int n_film = 0, bands = 0;
};
struct DecStageW {
  DwW up;                // [2r][C]
  GemmW halve;           // [C/2, C] + bias
  GemmW up_halve;        // fused: [r*C/2, 2C] = halve o transposed depthwise conv (valid when .w != nullptr)
  std::vector<ResW> res;
  int r, C;
};
struct DecoderW {
  GemmW pw0;
  DwW dw0;
  std::vector<DecStageW> stages;
  float* last_w = nullptr;   // [5][C]
  float last_b = 0.f;
  GemmW last_gemm;           // the same conv as a GEMM: [32, C] fp16, row j < 5 = tap j (wav_std folded), rest zero
  int C_last;
  float stage_scale;
};
struct HeadW {
  GemmW w;               // [n_out*hop, dim] pre-multiplied, bias = combined [n_out]
  int n_out, hop;
};

struct Weights {
  std::map<std::string, HostTensor> host;
  DevMem dev;
  const HostTensor& get(const std::string& name) const {
    auto it = host.find(name);
    if (it == host.end()) WV_THROW(WV_ERR_MISSING_WEIGHT, "missing tensor '%s'", name.c_str());
    return it->second;
  }
  const HostTensor* find(const std::string& name) const {
    auto it = host.find(name);
    return it == host.end() ? nullptr : &it->second;
  }
  float scalar_or(const std::string& name, float dflt) const {
    auto* t = find(name);
    return t ? t->data[0] : dflt;
  }
};

uint16_t f2h(float f) {
  __half h = __float2half_rn(f);
  uint16_t u;
  memcpy(&u, &h, 2);
  return u;
}

// Precise nets: the A operand is a split tensor [rows, 2*Ka] = [hi | lo] (or, stft: separate hi / lo frame views of
// K columns each).  Weight rows  [Wh | Wh] over the k-blocks of [hi | lo], then [Wl] over the hi k-blocks again:
//   (hi + lo) Wh + hi Wl = v W - lo Wl   (the dropped term is 2^-22 relative).   Wh = rn16(w), Wl = rn16(w - Wh).
GemmW make_gemm_w_pm(Weights& W, const std::vector<float>& rows, int N, int K, int Ka, int block_n,
                     const float* bias, int n_bias, bool stft) {
  GemmW g;
  const int n1 = stft ? 2 * ceil_div(K, BK) : ceil_div(2 * Ka, BK), n2 = ceil_div(stft ? K : Ka, BK);
  const int off2 = stft ? ceil_div(K, BK) * BK : Ka;      // column of the second Wh copy
  g.N = N; g.K = (n1 + n2) * BK; g.ldw = g.K; g.block_n = block_n; g.fp16 = true; g.K_alg = K;
  g.a2 = stft ? n1 / 2 : 0;
  g.a3 = n1;
  std::vector<uint16_t> h(static_cast<size_t>(N) * g.ldw, 0);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const float w = rows[static_cast<size_t>(n) * K + k];
      if (!(std::fabs(w) <= 65504.f)) WV_THROW(WV_ERR_UNSUPPORTED, "weight %g does not fit fp16", w);
      const __half wh = __float2half_rn(w);
      const __half wl = __float2half_rn(w - __half2float(wh));
      uint16_t uh, ul;
      memcpy(&uh, &wh, 2); memcpy(&ul, &wl, 2);
      uint16_t* r = h.data() + static_cast<size_t>(n) * g.ldw;
      r[k] = uh; r[off2 + k] = uh; r[n1 * BK + k] = ul;
    }
  g.w = W.dev.upload(h);
  if (bias) g.bias = W.dev.upload(std::vector<float>(bias, bias + n_bias));
  g.tm = make_tmap(g.w, 2, g.ldw, N, 1, g.ldw, 0, BK, block_n, true);
  return g;
}

// rows[N][K] fp32 (already scaled) -> device 16-bit [N, ldw] + tensor map
GemmW make_gemm_w(Weights& W, const std::vector<float>& rows, int N, int K, bool fp16, int block_n,
                  const float* bias, int n_bias) {
  if (g_pm) return make_gemm_w_pm(W, rows, N, K, static_cast<int>(round_up(K, 8)), block_n, bias, n_bias, false);
  GemmW g;
  fp16 = true;   // activations and weights are fp16 throughout (bf16 packing is no longer used)
  g.N = N; g.K = K; g.fp16 = fp16;
  for (float v : rows)
    if (!(std::fabs(v) <= 65504.f)) WV_THROW(WV_ERR_UNSUPPORTED, "weight %g does not fit fp16", v);
  g.ldw = static_cast<int>(round_up(K, 8));
  g.block_n = block_n;
  std::vector<uint16_t> h(static_cast<size_t>(N) * g.ldw, 0);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k)
      h[static_cast<size_t>(n) * g.ldw + k] =
          f2h(rows[static_cast<size_t>(n) * K + k]);
  g.w = W.dev.upload(h);
  if (bias) g.bias = W.dev.upload(std::vector<float>(bias, bias + n_bias));
  g.tm = make_tmap(g.w, 2, g.ldw, N, 1, g.ldw, 0, BK, block_n, fp16);   // padded K (zeros)
  return g;
}

GemmW pointwise(Weights& W, const std::string& p, float scale, bool with_bias, bool pairable = false) {
  const HostTensor& t = W.get(p + ".weight");
  if (t.shape.size() != 3 || t.shape[2] != 1)
    WV_THROW(WV_ERR_INVALID, "'%s.weight' is not a 1x1 conv", p.c_str());
  const int N = static_cast<int>(t.shape[0]), K = static_cast<int>(t.shape[1]);
  std::vector<float> rows(t.data, t.data + static_cast<size_t>(N) * K);
  for (auto& v : rows) v *= scale;
  const HostTensor* b = with_bias ? W.find(p + ".bias") : nullptr;
  std::vector<float> bb;
  if (b) {
    bb.assign(b->data, b->data + N);
    for (auto& v : bb) v *= scale;
  }
  // pointwise convs run through the STAGED epilogue (two staging tiles): tile width <= 128
  return make_gemm_w(W, rows, N, K, false, pick_block_n(N, 0, g_pm ? STAGED_PM_MAX_BN : STAGED_MAX_BN, pairable && !g_pm), b ? bb.data() : nullptr, N);
}

// depthwise [C,1,k] -> [k][C]; transposed conv weights have the same memory shape
DwW depthwise(Weights& W, const std::string& p, float scale, bool with_bias, float bias_scale = -1.f) {
  if (bias_scale < 0.f) bias_scale = scale;
  const HostTensor& t = W.get(p + ".weight");
  if (t.shape.size() != 3 || t.shape[1] != 1)
    WV_THROW(WV_ERR_INVALID, "'%s.weight' is not depthwise", p.c_str());
  DwW d;
  d.C = static_cast<int>(t.shape[0]);
  d.k = static_cast<int>(t.shape[2]);
  std::vector<float> h(static_cast<size_t>(d.k) * d.C);
  for (int c = 0; c < d.C; ++c)
    for (int j = 0; j < d.k; ++j) h[static_cast<size_t>(j) * d.C + c] = t.data[c * d.k + j] * scale;
  d.w = W.dev.upload(h);
  const HostTensor* b = with_bias ? W.find(p + ".bias") : nullptr;
  if (b) {
    std::vector<float> bb(b->data, b->data + d.C);
    for (auto& v : bb) v *= bias_scale;
    d.bias = W.dev.upload(bb);
  }
  return d;
}

ResW resblock_w(Weights& W, const std::string& p, int idx, float rs) {
  ResW r;
  r.pre_scale = 1.f / std::sqrt(1.f + idx * rs * rs);                     // seanet.py:183
  r.pw1 = pointwise(W, p + ".block.1.conv.conv", 1.f, false, true);
  r.dw1 = depthwise(W, p + ".block.2.conv.conv", 1.f, true);
  r.pw2 = pointwise(W, p + ".block.4.conv.conv", 1.f, false);
  const float tail = rs * W.scalar_or(p + ".res_scale_param", 1.f);       // seanet.py:271-277
  r.dw2 = depthwise(W, p + ".block.5.conv.conv", tail, true);
  return r;
}

SpecW spec_w(Weights& W, const std::string& p, int n_fft, int hop, float mean, float stdv, float rs) {
  SpecW s;
  s.n_fft = n_fft; s.hop = hop; s.mean = mean; s.stdv = stdv;
  const HostTensor& d = W.get(p + ".spec.weight");                        // [(n_fft+2), 1, n_fft]
  if (d.shape.size() != 3 || d.shape[0] != n_fft + 2 || d.shape[2] != n_fft)
    WV_THROW(WV_ERR_INVALID, "'%s.spec.weight' has an unexpected shape", p.c_str());
  const int half = n_fft / 2;
  auto row = [&](int r) { return d.data + static_cast<size_t>(r) * n_fft; };
  // imaginary rows of bins 0 and n_fft/2 must vanish: they share a column pair (see gemm epilogue)
  for (int n = 0; n < n_fft; ++n)
    if (std::fabs(row(half + 1)[n]) > 1e-3f || std::fabs(row(half + 1 + half)[n]) > 1e-3f)
      WV_THROW(WV_ERR_UNSUPPORTED, "'%s.spec.weight' is not a real-input DFT basis", p.c_str());
  std::vector<float> rows(static_cast<size_t>(n_fft) * n_fft);
  for (int pr = 0; pr < half; ++pr) {
    const float* re = pr == 0 ? row(0) : row(pr);
    const float* im = pr == 0 ? row(half) : row(half + 1 + pr);          // pair 0 = (bin 0, bin N/2)
    std::copy(re, re + n_fft, rows.begin() + static_cast<size_t>(2 * pr) * n_fft);
    std::copy(im, im + n_fft, rows.begin() + static_cast<size_t>(2 * pr + 1) * n_fft);
  }
  s.dft = g_pm ? make_gemm_w_pm(W, rows, n_fft, n_fft, n_fft, pick_block_n(n_fft), nullptr, 0, true)
               : make_gemm_w(W, rows, n_fft, n_fft, true, pick_block_n(n_fft), nullptr, 0);
  const float sc = rs * W.scalar_or(p + ".scale_param", 1.f);            // seanet.py:499-505
  s.layer = pointwise(W, p + ".layer.conv.conv", sc, false);
  return s;
}

// Last resblock of an encoder stage + spectrogram branch as one GEMM (modules/seanet.py:936-943):
//   x'' = RS*(dw5(W2 h1) + b2) + x + RS*scale*(Ws y)
// Operand rows of n tile nt (bn = block_n/2 output channels): rows [0, bn) = [W2 | 0], rows [bn, 2bn) = [0 | Ws];
// the contraction is [h1 (K1 padded to 64) | y (ldy padded to 64)], so accumulator columns [0, bn) hold
// W2 h1 (input of the depthwise taps) and columns [bn, 2bn) hold Ws y (added un-tapped).
void build_dual_w(Weights& W, EncStageW& st, const std::string& res_p, const std::string& spec_p, float rs, int C) {
  const HostTensor& w2 = W.get(res_p + ".block.4.conv.conv.weight");
  const HostTensor& ws = W.get(spec_p + ".layer.conv.conv.weight");
  const int K2 = st.spec.n_fft / 2 + 1;
  if (w2.shape.size() != 3 || w2.shape[0] != C || w2.shape[1] != C || ws.shape.size() != 3 || ws.shape[0] != C || ws.shape[1] != K2) return;
  const float sc = rs * W.scalar_or(spec_p + ".scale_param", 1.f);
  const int ldy = static_cast<int>(round_up(K2, g_ldy_align));
  const int k1p = static_cast<int>(round_up(C, BK)), k2p = static_cast<int>(round_up(ldy, BK));
  const int Kc = k1p + k2p;
  const int block_n = std::min(2 * C, STAGED_MAX_BN), bn = block_n / 2;
  if (block_n % 32 != 0 || C % bn != 0) return;
  std::vector<float> rows(static_cast<size_t>(2) * C * Kc, 0.f);
  for (int n = 0; n < C; ++n) {
    const int nt = n / bn, j = n % bn;
    float* r1 = rows.data() + (static_cast<size_t>(nt) * block_n + j) * Kc;
    float* r2 = rows.data() + (static_cast<size_t>(nt) * block_n + bn + j) * Kc;
    for (int k = 0; k < C; ++k) r1[k] = w2.data[static_cast<size_t>(n) * C + k];
    for (int k = 0; k < K2; ++k) r2[k1p + k] = ws.data[static_cast<size_t>(n) * K2 + k] * sc;
  }
  st.dual = make_gemm_w(W, rows, 2 * C, Kc, true, block_n, nullptr, 0);
  st.dual_split_kb = k1p / BK;
  st.dual_ldy = ldy;
}

// ------------------------------------------------------------------------------------------
// plan
struct Arena {
  struct Blk { size_t off, size; };
  std::vector<Blk> free_list;
  size_t top = 0, high = 0;
  size_t alloc(size_t bytes) {
    bytes = round_up(std::max<size_t>(bytes, 256), 256);
    for (size_t i = 0; i < free_list.size(); ++i)
      if (free_list[i].size >= bytes) {
        size_t off = free_list[i].off;
        free_list[i].off += bytes;
        free_list[i].size -= bytes;
        if (free_list[i].size == 0) free_list.erase(free_list.begin() + i);
        return off;
      }
    if (!free_list.empty() && free_list.back().off + free_list.back().size == top) {
      size_t off = free_list.back().off;          // grow the trailing free block
      top = off + bytes;
      free_list.pop_back();
      high = std::max(high, top);
      return off;
    }
    size_t off = top;
    top += bytes;
    high = std::max(high, top);
    return off;
  }
  void release(size_t off, size_t bytes) {
    bytes = round_up(std::max<size_t>(bytes, 256), 256);
    free_list.push_back({off, bytes});
    std::sort(free_list.begin(), free_list.end(), [](const Blk& a, const Blk& b) { return a.off < b.off; });
    std::vector<Blk> m;
    for (auto& b : free_list) {
      if (!m.empty() && m.back().off + m.back().size == b.off) m.back().size += b.size;
      else m.push_back(b);
    }
    free_list.swap(m);
  }
};

struct Buf {
  size_t off = 0, bytes = 0;
  bool valid = false;
};

struct IoPtrs {
  const float* x = nullptr;       // [B,T]
  const float* msg = nullptr;
  float* wm = nullptr;
  float* y = nullptr;
  float* latent = nullptr;
  const float* z_in = nullptr;
  float* logits = nullptr;
  uint8_t* bits = nullptr;
  float* avg = nullptr;
  float* conf = nullptr;
  uint8_t* valid = nullptr;
  const uint8_t* presence = nullptr;
  uint8_t* mask = nullptr;
  float* probs = nullptr;
};

enum OpType { OP_RESBLOCK = 20, OP_GEMM = 0, OP_DW5, OP_DOWN, OP_UP, OP_CONV_PRE, OP_CONV_LAST, OP_WAV_STAGE, OP_FRAMES,
              OP_FILM, OP_BITS, OP_CONF, OP_LATENT_IN, OP_CONV_PRE_PM = 30, OP_DW5_PM, OP_WAV_STAGE_PM };

struct Op {
  OpType type;
  // GEMM
  int epi = 0;
  CUtensorMap tmA, tmB, tmR;   // tmR: residual tile view for the L2 prefetch (else a copy of tmB)
  GemmArgs g;
  ResblockArgs rb;   // OP_RESBLOCK (tmA = x tiles, tmB = W1, tmR = W2)
  int grid = 0;
  // generic
  const void* in = nullptr;
  const void* res = nullptr;
  void* out0 = nullptr;
  void* out1 = nullptr;
  const float* w = nullptr;
  const float* bias = nullptr;
  const float* film = nullptr;
  float fa = 0.f, fb = 0.f;
  int i[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  FilmArgs fargs;
  // bookkeeping for taps / profiling
  std::string tag;
  double flops = 0, bytes = 0;      // algorithmic work of this launch
  size_t out_bytes[2] = {0, 0};
};

// A small (launch-bound) plan is captured once per I/O signature into a CUDA graph whose kernels
// read and write library-owned I/O buffers; a call then is: copy inputs in, one graph launch, copy
// outputs out (single-clip latency 2.0 ms -> see DESIGN.md).
struct PlanGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  uint8_t* mem = nullptr;
  IoPtrs io;                         // the library-owned buffers inside `mem`
  size_t bytes[14] = {0};
};
struct Plan {
  int B = 0, T = 0;
  std::vector<Op> ops;
  std::vector<cudaEvent_t> events;   // profiling: ops.size() + 1 events
  size_t ws_bytes = 0;
  Buf latent;          // h16 [B*F, dim]
  int F = 0;
  std::map<uint32_t, PlanGraph> graphs;   // keyed by the set of non-null I/O pointers
  std::map<uint32_t, int> graph_seen;     // calls per I/O signature: a graph is captured on the second one
  unsigned long long stamp = 0;           // last use (LRU eviction of cached plans)
  ~Plan() {
    for (auto ev : events) cudaEventDestroy(ev);
    for (auto& kv : graphs) {
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
      if (kv.second.graph) cudaGraphDestroy(kv.second.graph);
      if (kv.second.mem) cudaFree(kv.second.mem);
    }
  }
};

// Device-side re-evaluation of near-threshold detector clips (wv_detector_refine): one CUDA graph per shape with a WHILE
// node around {gather K clips, precise plan, scatter}; see glue_kernels.cuh.
struct RefineGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  uint8_t* mem = nullptr;
  RefineIo* io = nullptr;
  ~RefineGraph() {
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    if (mem) cudaFree(mem);
  }
};

struct PlanCtx {
  Arena arena;
  uint8_t* base = nullptr;   // nullptr => sizing pass (no tensor maps encoded)
  std::vector<Op>* ops = nullptr;
  int B = 0, T = 0;
  std::string next_tag;
  void push(Op& op) {
    if (op.type == OP_GEMM && g_serpentine) op.g.reverse = static_cast<int>(ops->size() & 1);   // alternate directions
    if (op.type == OP_GEMM) op.g.a_evict_first = g_evict_first ? 1 : 0;
    op.tag = next_tag;
    next_tag.clear();
    ops->push_back(op);
  }
  PlanCtx& tag(const std::string& t) {
    next_tag = t;
    return *this;
  }
  Buf alloc(size_t bytes) {
    Buf b;
    b.bytes = bytes;
    b.off = arena.alloc(bytes);
    b.valid = true;
    return b;
  }
  void release(Buf& b) {
    if (b.valid) arena.release(b.off, b.bytes);
    b.valid = false;
  }
  template <typename T>
  T* ptr(const Buf& b) const { return reinterpret_cast<T*>(base ? base + b.off : nullptr); }
  bool dry() const { return base == nullptr; }
};

int elem_grid(long long total_threads) {
  long long blocks = (total_threads + 255) / 256;
  const long long cap = static_cast<long long>(g_num_sms) * 16;
  return static_cast<int>(std::max<long long>(1, std::min(blocks, cap)));
}

// generic GEMM op: A given either as a flat [M, lda] 16-bit buffer or as a prebuilt 3-D view
void add_gemm(PlanCtx& c, int epi, const GemmW& w, const void* A, int lda, long long M, int K,
              GemmArgs g, const CUtensorMap* custom_tmA = nullptr, int rows_per_clip = 0,
              int n_clips = 1) {
  Op op;
  op.type = OP_GEMM;
  op.epi = epi;
  g.N = w.N;
  g.K = K;
  g.block_n = w.block_n;
  g.idesc = make_idesc_f16(BM, w.block_n, w.fp16);
  const bool staged = epi == EPI_STAGED || epi == EPI_STAGED_PM;
  const bool pm = epi == EPI_STAGED_PM;                      // fp32 staging tiles (smem plan)
  if (w.a3 > 0) { g.a2_split = w.a2; g.a3_split = w.a3; }    // split-fp16 operand segments (precise nets)
  g.unit_rows = (staged && ((g_rows6_mask >> ((g.dual ? w.block_n / 2 : w.block_n) / 32)) & 1u) && !g.last_mode && g.down_r == 0) ? 6 : 4;
  g.epi_groups = 1;
  const bool is_stft = epi == EPI_STFT || epi == EPI_STFT_PM;
  if (is_stft && g_epi_groups >= 2 && w.block_n <= 64) g.epi_groups = 2;
  if (is_stft && g_epi_groups >= 4 && w.block_n <= 128) g.epi_groups = 4;
  if (staged && g.taps != 1 && g.taps != 5) WV_THROW(WV_ERR_INVALID, "taps must be 1 or 5");
  const int num_kb = ceil_div(K, BK);
  const int tiles_n_ = w.N / w.block_n;
  const bool nt_fixed = tiles_n_ <= g_num_sms;   // the grid is rounded down to a multiple of tiles_n below
  bool resident = gemm_resident_b(w.block_n, num_kb, staged, nt_fixed, pm);
  // W tiles of 64..g_res1_kb KB (C = 512 layers: 128 KB) fit next to ONE staging tile: the ring then streams A only,
  // which halves the L2 -> SM operand traffic of those layers (they are bound by it); the price is that the drain of
  // a tile waits for the math of the previous one
  bool res1 = false;
  if (!resident && staged && !pm && nt_fixed && g_res1_kb > 0) {
    const int w_bytes = num_kb * w.block_n * BK * 2;
    res1 = w_bytes <= g_res1_kb * 1024 && (GEMM_SMEM_LIMIT - gemm_fixed_smem(w.block_n, true, 1) - w_bytes) / A_STAGE_BYTES >= 3;
    resident = res1;
  }
  g.resident_b = resident ? 1 : 0;
  // long-K layers (>= g_one_buf_kb k-blocks per tile, W streamed) could run with one staging tile and a deeper
  // operand ring; measured: no gain (the deep stages are bound by L2 -> SM operand traffic, not ring depth): off
  g.a_prefetch = staged ? g_a_prefetch : 0;
  g.res_early2 = g_res_early2 ? 1 : 0;
  g.stage_bufs = (res1 || (staged && !resident && g_one_buf_kb > 0 && num_kb >= g_one_buf_kb)) ? 1 : STAGE_BUFS;
  {
    // measured per launch class (profiles/r01h_math_groups.md): two groups win for the 1x1 + dw5 launches without a
    // residual (resblock first halves, -5..-12 %) and for the 64-column tiles / the spec-fused launch with one;
    // they lose 2-4 % on the residual launches with 96 / 128-column tiles and on the plain 1x1 launches
    const int bn_eff = g.dual ? w.block_n / 2 : w.block_n;
    const bool heur = g.residual == nullptr ? g.taps == 5 : (bn_eff == 64 || g.dual);
    const bool want = g_math_groups == 2 || (g_math_groups == 3 && g.residual == nullptr) || (g_math_groups == 4 && heur);
    g.math_groups = (staged && !pm && want && g.stage_bufs == 2 && g.down_r == 0 && !g.last_mode) ? 2 : 1;
  }
  // pair mode (two M tiles per W k-block) for the long-K layers whose W tile does not stay resident
  const bool pair = staged && !pm && !resident && g.down_r == 0 && g.a2_split == 0 && g.a3_split == 0 && w.block_n <= 128 && g_pair_min_kb > 0 && num_kb >= g_pair_min_kb;
  g.pair = pair ? 1 : 0;
  // CTA-pair mode (cta_group::2) for the same class of layers when the n tiles pair up: half the operand bytes per MAC
  const bool cg2 = epi == EPI_STAGED && !resident && !g.dual && !g.last_mode && g.pre_w == nullptr && g.a2_split == 0 &&
                   g.a3_split == 0 && g_cg2_min_kb > 0 && num_kb >= g_cg2_min_kb && g_max_clusters > 0 && tiles_n_ % 2 == 0 &&
                   2 * w.block_n <= MAX_BN && g.stage_bufs == STAGE_BUFS;
  g.cg2 = 0;
  if (cg2) { g.pair = 0; }
  g.acc_stages = (g.pair || g.epi_groups == 4) ? MAX_ACC_STAGES : ACC_STAGES;   // four epilogue groups: one 128-column stage each
  g.acc_cols = (g.pair || g.epi_groups == 4) ? TMEM_COLS / MAX_ACC_STAGES : MAX_BN;
  g.stages = gemm_stage_count(w.block_n, staged, num_kb, resident, g.stage_bufs, g.pair != 0, pm);
  if (g.stages < 2) WV_THROW(WV_ERR_UNSUPPORTED, "not enough shared memory for block_n=%d", w.block_n);
  op.i[7] = gemm_smem_bytes(w.block_n, staged, num_kb, resident, g.stage_bufs, g.pair != 0, pm);
  if (custom_tmA) {
    g.rows_per_clip = rows_per_clip;
    g.n_clips = n_clips;
    op.tmA = *custom_tmA;
  } else {
    g.rows_per_clip = static_cast<int>(M);
    g.n_clips = 1;
    // split operands: the tensor is lda wide (reads past it must zero-fill), the contraction K revisits its k-blocks
    if (!c.dry()) op.tmA = make_tmap(A, 3, w.a3 > 0 ? lda : K, M, 1, lda, static_cast<uint64_t>(lda) * M, BK, BM, w.fp16);
  }
  op.tmB = w.tm;
  op.tmR = w.tm;
  if (staged && g.residual != nullptr && g.a2_split == 0 && !c.dry())   // residual shares the output's [clip, row, ldo] layout
    op.tmR = make_tmap(g.residual, 3, g.ldo, g.rows_per_clip, g.n_clips, g.ldo,
                       static_cast<uint64_t>(g.ldo) * g.rows_per_clip, w.block_n, BM, false, false);
  op.g = g;
  op.out0 = pm ? static_cast<void*>(g.out_raw32) : static_cast<void*>(g.out_raw);
  op.out1 = g.out_act;
  int rows_out = BM - (staged ? g.taps - 1 : 0);
  g.tile_halo = staged ? g.taps - 1 : 0;
  g.rows_per_clip_out = g.rows_per_clip;
  if (staged && g.down_r > 0) {
    const int outs = BM / g.down_r - 1;
    rows_out = outs * g.down_r;
    g.tile_halo = g.down_r;
    g.rows_per_clip_out = ceil_div(g.rows_per_clip, g.down_r);
  }
  g.tile_stride = rows_out;
  g.tiles_n = w.N / w.block_n;
  g.tiles_m_per_clip = (staged && g.down_r > 0) ? ceil_div(g.rows_per_clip_out, BM / g.down_r - 1)
                                                 : ceil_div(g.rows_per_clip, rows_out);
  g.magic_n = static_cast<uint32_t>((1ull << 32) / static_cast<uint64_t>(g.tiles_n));
  g.magic_m = static_cast<uint32_t>((1ull << 32) / static_cast<uint64_t>(g.tiles_m_per_clip));
  if (g.tiles_n == 1) g.magic_n = 0;            // 2^32 / 1 does not fit; the kernel special-cases 1
  if (g.tiles_m_per_clip == 1) g.magic_m = 0xFFFFFFFFu;   // x * (2^32-1) >> 32 = x - 1, fixed up in-kernel
  const long long tiles_ll = static_cast<long long>(g.tiles_m_per_clip) * g.n_clips * g.tiles_n;
  if (tiles_ll >= (1ll << 31)) WV_THROW(WV_ERR_UNSUPPORTED, "too many tiles (%lld)", tiles_ll);
  const int tiles = static_cast<int>(tiles_ll);
  g.num_tiles = tiles;
  op.g = g;
  op.grid = std::min(tiles, g_num_sms);
  if (cg2) {
    const long long units = (static_cast<long long>(g.tiles_m_per_clip) * g.n_clips + 1) / 2 * (g.tiles_n / 2);
    if (units >= 2LL * g_max_clusters) {   // at least two units per CTA pair, else the single-CTA path spreads the tiles better
      g.cg2 = 1;
      g.idesc2 = make_idesc_f16(2 * BM, 2 * w.block_n, w.fp16);
      const int ntw = g.tiles_n / 2;
      g.magic_n2 = ntw == 1 ? 0u : static_cast<uint32_t>((1ull << 32) / static_cast<uint64_t>(ntw));
      op.grid = 2 * static_cast<int>(std::min<long long>(units, g_max_clusters));
      op.g = g;
    } else if (pair) {
      g.pair = 1;   // fall back to the in-CTA pair mode (re-derive its layout)
      g.acc_stages = MAX_ACC_STAGES; g.acc_cols = TMEM_COLS / MAX_ACC_STAGES;
      g.stages = gemm_stage_count(w.block_n, staged, num_kb, false, g.stage_bufs, true, pm);
      op.i[7] = gemm_smem_bytes(w.block_n, staged, num_kb, false, g.stage_bufs, true, pm);
      op.g = g;
    }
  }
  if (g.pair) {
    const long long units = (static_cast<long long>(g.tiles_m_per_clip) * g.n_clips + 1) / 2 * g.tiles_n;
    if (units < 4LL * g_num_sms) {   // too few units: wave quantisation costs more than the shared W loads save
      g.pair = 0;
      g.acc_stages = ACC_STAGES; g.acc_cols = MAX_BN;
      g.stages = gemm_stage_count(w.block_n, staged, num_kb, false, g.stage_bufs, false, pm);
      op.i[7] = gemm_smem_bytes(w.block_n, staged, num_kb, false, g.stage_bufs, false, pm);
      op.g = g;
    } else {
      op.grid = static_cast<int>(std::min<long long>(units, g_num_sms));
    }
  }
  if (g.resident_b && op.grid > g.tiles_n) op.grid -= op.grid % g.tiles_n;   // every CTA keeps one n tile
  if (g.resident_b && op.grid < g.tiles_n) {   // fewer tiles than n tiles: stream W through the ring
    g.resident_b = 0;
    g.stages = gemm_stage_count(w.block_n, staged, num_kb, false, g.stage_bufs, false, pm);
    op.i[7] = gemm_smem_bytes(w.block_n, staged, num_kb, false, g.stage_bufs, false, pm);
    op.g = g;
  }
  // residual tile by TMA: W resident (the ring streams A only), plain single-CTA launches with a stored residual
  if (g_res_tma && epi == EPI_STAGED && g.residual != nullptr && g.resident_b && !g.cg2 && !g.pair && !g.dual && g.a2_split == 0 &&
      g.down_r == 0 && !g.last_mode && g.pre_w == nullptr && g.ldo == w.N) {
    const int res_bytes = 2 * BM * w.block_n * 2 + 128;
    int st = g.stages, bytes = op.i[7];
    while (st > g_res_tma_min_stages && bytes + res_bytes > GEMM_SMEM_LIMIT) { --st; bytes -= A_STAGE_BYTES; }
    if (bytes + res_bytes <= GEMM_SMEM_LIMIT) {
      g.res_tma = 1; g.stages = st;
      op.i[7] = bytes + res_bytes;
      op.g = g;
    }
  }
  {
    const double Mt = static_cast<double>(g.rows_per_clip) * g.n_clips;
    op.flops = 2.0 * Mt * w.N * (w.K_alg > 0 ? w.K_alg : K);   // algorithmic (the split contraction issues 3-4x the MMAs)
    double outs = 0;
    if (is_stft) outs = Mt * (w.N / 2 + 1) * 2;
    else if (epi == EPI_HEAD) outs = 0;   // logits / mask bytes are caller dependent, added at run time
    else outs = Mt * w.N * 2.0 * ((g.out_raw ? 1 : 0) + (g.out_act ? 1 : 0));
    // the strided STFT frame view re-reads each sample n_fft/hop times from L2; algorithmic = once
    const double a_bytes = custom_tmA ? Mt * 2.0 * std::min<double>(K, 64) : Mt * K * 2.0;
    op.bytes = a_bytes + outs + (g.residual ? Mt * w.N * 2.0 : 0.0) + static_cast<double>(w.N) * K * 2.0;
    const double Mo = static_cast<double>(g.rows_per_clip_out) * g.n_clips;
    op.out_bytes[0] = g.out_raw ? static_cast<size_t>(Mo) * g.ldo * 2 : 0;
    op.out_bytes[1] = g.out_act ? static_cast<size_t>(Mo) * g.ldo * 2 : 0;
    if (pm) {
      op.out_bytes[0] = g.out_raw32 ? static_cast<size_t>(Mo) * g.ldo * 4 : 0;
      op.out_bytes[1] = g.out_act ? static_cast<size_t>(Mo) * g.ldo_act * 2 : 0;
      op.bytes = Mt * (w.K_alg > 0 ? w.K_alg : K) * 4.0 + Mo * w.N * 4.0 * ((g.out_raw32 ? 1 : 0) + (g.out_act ? 1 : 0) + (g.residual32 ? 1 : 0)) +
                 static_cast<double>(w.N) * K * 2.0;
    }
  }
  c.push(op);
}

GemmArgs std_args(const float* bias, const h16* res, h16* out_raw, h16* out_act, float act_scale, int ldo) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.bias = bias; g.residual = res; g.out_raw = out_raw; g.out_act = out_act;
  g.act_scale = act_scale; g.ldo = ldo;
  g.taps = 1;
  return g;
}

void add_dw5(PlanCtx& c, const DwW& w, const h16* in, const h16* res, h16* out_raw, h16* out_act,
             float act_scale, int T, int C) {
  Op op;
  op.type = OP_DW5;
  op.in = in; op.res = res; op.out0 = out_raw; op.out1 = out_act;
  op.w = w.w; op.bias = w.bias; op.fa = act_scale;
  op.i[0] = c.B; op.i[1] = T; op.i[2] = C;
  op.grid = elem_grid(static_cast<long long>(c.B) * ceil_div(T, DW_TT) * (C / 8));
  const double n = static_cast<double>(c.B) * T * C;
  op.flops = 10.0 * n;
  op.bytes = 2.0 * n * (1 + (res ? 1 : 0) + (out_raw ? 1 : 0) + (out_act ? 1 : 0));
  op.out_bytes[0] = out_raw ? static_cast<size_t>(n) * 2 : 0;
  op.out_bytes[1] = out_act ? static_cast<size_t>(n) * 2 : 0;
  c.push(op);
}


// 1x1 conv (tcgen05 GEMM) with the following causal depthwise k=5 conv fused into its epilogue:
// per-clip overlapping tiles (128 rows in, 124 out).  A is [B, T, K] channels-last.
void add_gemm_dw(PlanCtx& c, const GemmW& pw, const DwW& dw, const h16* A, int T, int K, const h16* res,
                 h16* out_raw, h16* out_act, float act_scale, const DwW* pre = nullptr) {
  if (dw.k != 5 || dw.C != pw.N) WV_THROW(WV_ERR_INVALID, "fused depthwise conv must be k=5 over the GEMM's N");
  GemmArgs g = std_args(dw.bias, res, out_raw, out_act, act_scale, pw.N);
  g.taps = 5;
  g.dw_w = dw.w;
  if (pre != nullptr) {   // residual = conv_pre(waveform), recomputed in the epilogue (io.x at run time)
    if (res != nullptr || (!c.dry() && (!out_raw || !out_act)) || pre->k != 5 || pre->C != pw.N || !pre->bias)
      WV_THROW(WV_ERR_INVALID, "pre-conv residual needs k=5 taps over N, both outputs and no stored residual");
    g.pre_w = pre->w; g.pre_b = pre->bias; g.pre_T = T;
  }
  CUtensorMap tm;
  if (!c.dry()) tm = make_tmap(A, 3, K, T, c.B, K, static_cast<uint64_t>(K) * T, BK, BM, false);
  add_gemm(c, EPI_STAGED, pw, nullptr, 0, 0, K, g, &tm, T, c.B);
  Op& op = c.ops->back();
  const double n = static_cast<double>(c.B) * T;
  op.flops += 10.0 * n * pw.N;
  op.bytes = n * 2.0 * (K + pw.N * ((res ? 1 : 0) + (out_raw ? 1 : 0) + (out_act ? 1 : 0))) + static_cast<double>(pw.N) * K * 2.0 +
             (pre ? n * 4.0 : 0.0);
}

// Encoder downsample (modules/seanet.py:745-771): 1x1 conv C -> 2C (tcgen05 GEMM) with the strided
// depthwise conv k=2r, s=r, its bias and the FiLM affine fused into the epilogue.  A = [B, T, K].
void add_gemm_down(PlanCtx& c, const GemmW& pw, const DwW& dw, int r, const h16* A, int T, int K, const float* film,
                   int film_stride, int bands, h16* out_raw, h16* out_act, float act_scale) {
  if (dw.k != 2 * r || dw.C != pw.N) WV_THROW(WV_ERR_INVALID, "fused down-conv must be k=2r over the GEMM's N");
  if (r != 2 && r != 4 && r != 5 && r != 8) WV_THROW(WV_ERR_UNSUPPORTED, "fused down-conv stride %d (2, 4, 5, 8)", r);
  GemmArgs g = std_args(dw.bias, nullptr, out_raw, out_act, act_scale, pw.N);
  g.down_r = r;
  g.dw_w = dw.w;
  g.film = film; g.film_stride = film_stride; g.film_bands = bands;
  CUtensorMap tm;
  if (!c.dry()) tm = make_tmap(A, 3, K, T, c.B, K, static_cast<uint64_t>(K) * T, BK, BM, false);
  add_gemm(c, EPI_STAGED, pw, nullptr, 0, 0, K, g, &tm, T, c.B);
  Op& op = c.ops->back();
  const double n_in = static_cast<double>(c.B) * T, n_out = static_cast<double>(c.B) * ceil_div(T, r);
  op.flops += 2.0 * 2 * r * n_out * pw.N;
  op.bytes = n_in * K * 2.0 + n_out * pw.N * 2.0 * ((out_raw ? 1 : 0) + (out_act ? 1 : 0)) + static_cast<double>(pw.N) * K * 2.0;
}

bool resblock_fusable(const ResW& r, int C) {
  return g_rb_maxc > 0 && C <= g_rb_maxc && C <= 128 && C % 32 == 0 && r.pw1.N == C && r.pw1.K == C && r.pw2.N == C &&
         r.pw2.K == C && r.pw1.block_n == C && r.pw2.block_n == C && r.dw1.bias && r.dw2.bias &&
         rb_pick_nx(C, ceil_div(C, BK)) >= 2;
}

// One residual block as ONE launch (resblock_sm100.cuh): reads the activated stream A (GEMM operand) and the raw stream X
// (residual): 4C bytes per element instead of the 6C of the two-launch form.
void add_resblock_fused(PlanCtx& c, const ResW& r, const h16* X, const h16* A, int T, int C, h16* out_raw, h16* out_act,
                        float act_scale, const std::string& name) {
  Op op;
  op.type = OP_RESBLOCK;
  ResblockArgs& g = op.rb;
  memset(&g, 0, sizeof(g));
  g.C = C; g.num_kb = ceil_div(C, BK);
  g.T = T; g.n_clips = c.B;
  g.tiles_m_per_clip = ceil_div(T, RB_ROWS_OUT);
  const long long tiles = static_cast<long long>(g.tiles_m_per_clip) * c.B;
  if (tiles >= (1ll << 31)) WV_THROW(WV_ERR_UNSUPPORTED, "too many tiles (%lld)", tiles);
  g.num_tiles = static_cast<int>(tiles);
  g.magic_m = g.tiles_m_per_clip == 1 ? 0xFFFFFFFFu : static_cast<uint32_t>((1ull << 32) / static_cast<uint64_t>(g.tiles_m_per_clip));
  g.nx = rb_pick_nx(C, g.num_kb);
  g.idesc = make_idesc_f16(BM, C, true);
  g.pre_scale = r.pre_scale;
  g.dw1_w = r.dw1.w; g.dw1_b = r.dw1.bias; g.dw2_w = r.dw2.w; g.dw2_b = r.dw2.bias;
  g.x = X; g.out_raw = out_raw; g.out_act = out_act; g.act_scale = act_scale;
  if (!c.dry()) op.tmA = make_tmap(A, 3, C, T, c.B, C, static_cast<uint64_t>(C) * T, BK, BM, true);
  op.tmB = r.pw1.tm;
  op.tmR = r.pw2.tm;
  op.grid = std::min(g.num_tiles, g_num_sms);
  op.i[7] = rb_smem_bytes(C, g.num_kb, g.nx);
  op.out0 = out_raw; op.out1 = out_act;
  const double n = static_cast<double>(c.B) * T;
  op.flops = n * C * (4.0 * C + 20.0);
  op.bytes = n * C * 2.0 * (2 + (out_raw ? 1 : 0) + (out_act ? 1 : 0)) + 4.0 * C * C;
  op.out_bytes[0] = out_raw ? static_cast<size_t>(n) * C * 2 : 0;
  op.out_bytes[1] = out_act ? static_cast<size_t>(n) * C * 2 : 0;
  c.tag(name + ".fused");
  c.push(op);
}

// One residual block (modules/seanet.py:245-281).  X raw (residual), A = ELU(X*pre_scale).
// Produces Xn (raw, if need_raw) and An = ELU(Xn*next_act_scale) (if need_act); frees X and A.
void plan_resblock(PlanCtx& c, const ResW& r, Buf& X, Buf& A, int T, int C, bool need_raw, bool need_act,
                   float next_act_scale, Buf& Xn, Buf& An, const std::string& name, const DwW* pre = nullptr) {
  const long long M = static_cast<long long>(c.B) * T;
  const size_t bytes = static_cast<size_t>(M) * C * 2;
  if (resblock_fusable(r, C) && pre == nullptr && A.valid && X.valid) {
    Xn = Buf(); An = Buf();
    if (need_raw) Xn = c.alloc(bytes);
    if (need_act) An = c.alloc(bytes);
    add_resblock_fused(c, r, c.ptr<h16>(X), c.ptr<h16>(A), T, C, need_raw ? c.ptr<h16>(Xn) : nullptr,
                       need_act ? c.ptr<h16>(An) : nullptr, next_act_scale, name);
    c.release(A);
    c.release(X);
    return;
  }
  // half 1: A2 = ELU(dw5(W1 * A) + b1)
  Buf A2 = c.alloc(bytes);
  c.tag(name + ".h1");
  add_gemm_dw(c, r.pw1, r.dw1, c.ptr<h16>(A), T, C, nullptr, nullptr, c.ptr<h16>(A2), 1.f);
  c.release(A);
  // half 2: Xn = RS*(dw5(W2 * A2) + b2) + X ; An = ELU(Xn * next_scale)   (RS folded into dw2)
  Xn = Buf(); An = Buf();
  if (need_raw) Xn = c.alloc(bytes);
  if (need_act) An = c.alloc(bytes);
  c.tag(name + ".out");
  add_gemm_dw(c, r.pw2, r.dw2, c.ptr<h16>(A2), T, C, pre ? nullptr : c.ptr<h16>(X), need_raw ? c.ptr<h16>(Xn) : nullptr,
              need_act ? c.ptr<h16>(An) : nullptr, next_act_scale, pre);
  c.release(A2);
  if (X.valid) c.release(X);
}

struct Net;
void plan_spec(PlanCtx& c, const SpecW& s, const Buf& wav16, int pitch, int lead, int F, Buf& X,
               int C, float act_scale, Buf& Aout, const std::string& name);

}  // namespace

// ------------------------------------------------------------------------------------------
struct wv_net {
  wv_net_config cfg;
  int device = 0;
  Weights W;
  EncoderW enc;
  DecoderW dec;
  HeadW head;
  std::map<std::pair<int, int>, std::unique_ptr<Plan>> plans;       // full forward
  std::map<std::pair<int, int>, std::unique_ptr<Plan>> dec_plans;   // decode-only (B, F)
  uint8_t* ws = nullptr;
  size_t ws_bytes = 0;
  long long chunk_samples = 0;
  bool profile = false;
  const Plan* last_plan = nullptr;   // plan whose events hold the last profiled run
  cudaStream_t cap_stream = nullptr; // capture stream for the small-batch CUDA graphs
  unsigned long long clock = 0;      // LRU clock of the plan caches
  std::map<std::vector<int>, std::unique_ptr<RefineGraph>> refine;   // keyed by (B, T, K, logits, presence)
  bool check_range = false;               // scan every launch's fp16 outputs for saturated / non-finite values (wv_net_set_range_check)
  unsigned long long* range_stats = nullptr;   // device: {saturated, non-finite, max |v| as fp16 bits}
  ~wv_net() {
    if (range_stats) cudaFree(range_stats);
    refine.clear();
    if (cap_stream) cudaStreamDestroy(cap_stream);
    plans.clear();
    dec_plans.clear();
    if (ws) cudaFree(ws);
  }
};

namespace {

void build_encoder_w(wv_net& n) {
  Weights& W = n.W;
  const wv_net_config& cf = n.cfg;
  EncoderW& e = n.enc;
  const std::string p = "encoder";
  const float rs = cf.res_scale_enc;
  e.C0 = cf.channels_enc;
  e.dim = cf.dimension;
  // x/wav_std feeds the conv (seanet.py:658): fold into the taps, NOT into the bias
  e.conv_pre = depthwise(W, p + ".conv_pre.1.conv.conv", 1.f / WAV_STD, true, 1.f);
  if (e.conv_pre.k != 5) WV_THROW(WV_ERR_UNSUPPORTED, "conv_pre kernel size %d (only 5)", e.conv_pre.k);
  if (!e.conv_pre.bias) WV_THROW(WV_ERR_MISSING_WEIGHT, "conv_pre bias missing (bias=True required)");
  int C = e.C0, nfft = cf.n_fft_base, hop = 1;
  const int S = cf.n_strides;
  if (S + 1 > 5) WV_THROW(WV_ERR_UNSUPPORTED, "more than 4 strides");
  for (int s = 0; s < S; ++s) {
    EncStageW st;
    st.r = cf.strides[S - 1 - s];                                              // seanet.py:646
    st.C = C;
    for (int j = 1; j <= cf.n_residual_enc; ++j)                                // idx=j, seanet.py:684
      st.res.push_back(resblock_w(W, p + ".blocks." + std::to_string(s) + "." + std::to_string(j - 1), j, rs));
    st.spec = spec_w(W, p + ".spec_blocks." + std::to_string(s), nfft, hop, SPEC_MEANS[s], SPEC_STDS[s], rs);
    if (cf.n_residual_enc >= 1 && C <= g_spec_fuse_maxc && !g_pm)
      build_dual_w(W, st, p + ".blocks." + std::to_string(s) + "." + std::to_string(cf.n_residual_enc - 1),
                   p + ".spec_blocks." + std::to_string(s), rs, C);
    st.down_pw = pointwise(W, p + ".downsample." + std::to_string(s) + ".2.conv.conv", 1.f, false);
    st.down_dw = depthwise(W, p + ".downsample." + std::to_string(s) + ".3.conv.conv", 1.f, true);
    if (st.down_dw.k != 2 * st.r) WV_THROW(WV_ERR_INVALID, "downsample kernel != 2*stride");
    e.stages.push_back(st);
    C *= 2; nfft *= 2; hop *= st.r;
  }
  e.C_last = C; e.hop = hop; e.nfft_max = nfft;
  e.spec_post = spec_w(W, p + ".spec_post", nfft, hop, SPEC_MEANS[4], SPEC_STDS[4], rs);
  e.post_dw = depthwise(W, p + ".conv_post.1.conv.conv", 1.f, false);
  e.post_pw = pointwise(W, p + ".conv_post.2.conv.conv", 1.f, true);
  if (e.post_pw.N > MAX_BN) WV_THROW(WV_ERR_UNSUPPORTED, "latent dimension > 256");
  e.post_pw.block_n = e.post_pw.N;   // L2 norm needs the whole row in one tile
  e.post_pw.tm = make_tmap(e.post_pw.w, 2, e.post_pw.K, e.post_pw.N, 1, e.post_pw.ldw, 0, BK, e.post_pw.N, false);
  if (cf.kind == WV_KIND_GENERATOR) {
    e.has_film = true;
    FilmArgs& f = e.film;
    memset(&f, 0, sizeof(f));
    f.msg_dim = cf.msg_dimension; f.E = cf.embedding_dim; f.n_hidden = cf.embedding_layers;
    if (f.n_hidden > 3) WV_THROW(WV_ERR_UNSUPPORTED, "embedding_layers > 3");
    auto up = [&](const std::string& name, int64_t n_expect) {
      const HostTensor& t = W.get(name);
      if (t.numel() != n_expect) WV_THROW(WV_ERR_INVALID, "'%s' has %lld elements, expected %lld", name.c_str(), (long long)t.numel(), (long long)n_expect);
      return W.dev.upload(std::vector<float>(t.data, t.data + t.numel()));
    };
    f.w[0] = up(p + ".msg_embedding.0.weight", (int64_t)f.E * f.msg_dim);
    f.b[0] = up(p + ".msg_embedding.0.bias", f.E);
    for (int l = 0; l < f.n_hidden; ++l) {
      f.w[l + 1] = up(p + ".msg_embedding." + std::to_string(1 + 2 * l) + ".weight", (int64_t)f.E * f.E);
      f.b[l + 1] = up(p + ".msg_embedding." + std::to_string(1 + 2 * l) + ".bias", f.E);
    }
    e.bands = cf.freq_bands;
    e.n_film = S * e.bands;
    std::vector<float> gw, gb, bw, bb;
    for (int s = 0; s < S; ++s)
      for (int b = 0; b < e.bands; ++b) {
        const std::string q = p + ".film_layers." + std::to_string(s) + "." + std::to_string(b);
        const HostTensor& a = W.get(q + ".gamma_layer.weight");
        const HostTensor& c2 = W.get(q + ".beta_layer.weight");
        gw.insert(gw.end(), a.data, a.data + f.E);
        bw.insert(bw.end(), c2.data, c2.data + f.E);
        gb.push_back(W.get(q + ".gamma_layer.bias").data[0]);
        bb.push_back(W.get(q + ".beta_layer.bias").data[0]);
      }
    f.gw = W.dev.upload(gw); f.gb = W.dev.upload(gb);
    f.bw = W.dev.upload(bw); f.bb = W.dev.upload(bb);
    f.n_film = e.n_film;
    for (auto& st : e.stages)
      if ((2 * st.C) % e.bands != 0 || ((2 * st.C) / e.bands) % 8 != 0)
        WV_THROW(WV_ERR_INVALID, "Number of channels (%d) must be divisible by freq_bands (%d)", 2 * st.C, e.bands);
  }
}

void build_decoder_w(wv_net& n) {
  Weights& W = n.W;
  const wv_net_config& cf = n.cfg;
  DecoderW& d = n.dec;
  const std::string p = "decoder.model";
  const float rs = cf.res_scale_dec;
  int i = 0;
  d.pw0 = pointwise(W, p + "." + std::to_string(i++) + ".conv.conv", 1.f, false);
  d.dw0 = depthwise(W, p + "." + std::to_string(i++) + ".conv.conv", 1.f, true);
  int C = d.pw0.N;
  d.stage_scale = 1.f / std::sqrt(1.f + cf.n_residual_dec * rs * rs);            // seanet.py:1104
  for (int s = 0; s < cf.n_strides; ++s) {
    DecStageW st;
    st.r = cf.strides[s];
    st.C = C;
    i += 2;
    st.up = depthwise(W, p + "." + std::to_string(i++) + ".convtr.convtr", 1.f, false);
    if (st.up.k != 2 * st.r) WV_THROW(WV_ERR_INVALID, "upsample kernel != 2*stride");
    const std::string up_name = p + "." + std::to_string(i - 1) + ".convtr.convtr";
    const std::string hv_name = p + "." + std::to_string(i) + ".conv.conv";
    st.halve = pointwise(W, p + "." + std::to_string(i++) + ".conv.conv", 1.f, true, true);
    if (g_up_fuse_maxc > 0 && C <= g_up_fuse_maxc && C % BK == 0) {
      // out[r*i + j, n] = b[n] + sum_c Wh[n,c] (a[i,c] w[c,j] + a[i-1,c] w[c,j+r])   (modules/conv.py:838-874 followed by
      // the 1x1): one GEMM over the LOW-rate rows with K = [a[i] | a[i-1]] and N = (j, n); its [Ts, r*C/2]
      // output IS the [Ts*r, C/2] high-rate tensor.  The upsampled tensor never exists.
      const HostTensor& uw = W.get(up_name + ".weight");      // [C, 1, 2r]
      const HostTensor& hw = W.get(hv_name + ".weight");      // [C/2, C, 1]
      const HostTensor* hb = W.find(hv_name + ".bias");
      const int r = st.r, Ch = C / 2;
      std::vector<float> rows(static_cast<size_t>(r) * Ch * 2 * C), bias(static_cast<size_t>(r) * Ch, 0.f);
      for (int j = 0; j < r; ++j)
        for (int n = 0; n < Ch; ++n) {
          float* dst = rows.data() + (static_cast<size_t>(j) * Ch + n) * 2 * C;
          for (int cc = 0; cc < C; ++cc) {
            dst[cc] = hw.data[static_cast<size_t>(n) * C + cc] * uw.data[static_cast<size_t>(cc) * 2 * r + j];
            dst[C + cc] = hw.data[static_cast<size_t>(n) * C + cc] * uw.data[static_cast<size_t>(cc) * 2 * r + j + r];
          }
          if (hb) bias[static_cast<size_t>(j) * Ch + n] = hb->data[n];
        }
      st.up_halve = make_gemm_w(W, rows, r * Ch, 2 * C, true, pick_block_n(r * Ch, 0, STAGED_MAX_BN, true), bias.data(), r * Ch);
    }
    for (int j = 0; j < cf.n_residual_dec; ++j)
      st.res.push_back(resblock_w(W, p + "." + std::to_string(i++), j, rs));     // idx=j, seanet.py:1159
    d.stages.push_back(st);
    C /= 2;
  }
  i += 2;
  const HostTensor& lw = W.get(p + "." + std::to_string(i) + ".conv.conv.weight");   // [1, C, 5]
  if (lw.shape.size() != 3 || lw.shape[0] != 1 || lw.shape[1] != C || lw.shape[2] != 5)
    WV_THROW(WV_ERR_INVALID, "decoder output conv has an unexpected shape");
  std::vector<float> h(5 * C);
  for (int c = 0; c < C; ++c)
    for (int j = 0; j < 5; ++j) h[j * C + c] = lw.data[c * 5 + j] * WAV_STD;        // seanet.py:1193
  d.last_w = W.dev.upload(h);
  {
    std::vector<float> rows(static_cast<size_t>(32) * C, 0.f);
    for (int j = 0; j < 5; ++j)
      for (int c = 0; c < C; ++c) rows[static_cast<size_t>(j) * C + c] = h[j * C + c];
    d.last_gemm = make_gemm_w(W, rows, 32, C, true, 32, nullptr, 0);
  }
  const HostTensor* lb = W.find(p + "." + std::to_string(i) + ".conv.conv.bias");
  d.last_b = lb ? lb->data[0] * WAV_STD : 0.f;
  d.C_last = C;
}

void build_head_w(wv_net& n) {
  Weights& W = n.W;
  const HostTensor& rc = W.get("reverse_convolution.weight");      // [dim, OD, hop]
  const HostTensor& rb = W.get("reverse_convolution.bias");        // [OD]
  const HostTensor& ll = W.get("last_layer.weight");               // [O, OD, 1]
  const HostTensor& lb = W.get("last_layer.bias");                 // [O]
  const int dim = static_cast<int>(rc.shape[0]), OD = static_cast<int>(rc.shape[1]), hop = static_cast<int>(rc.shape[2]);
  const int O = static_cast<int>(ll.shape[0]);
  if (hop != n.enc.hop || dim != n.cfg.dimension || ll.shape[1] != OD)
    WV_THROW(WV_ERR_INVALID, "head shapes inconsistent with the encoder");
  // logits[o, f*hop+j] = b_ll[o] + sum_c W_ll[o,c] (b_rc[c] + sum_z W_rc[z,c,j] z[z,f])
  // (model/detector.py:304-310): both maps are linear with nothing in between -> pre-multiply.
  std::vector<float> rows(static_cast<size_t>(O) * hop * dim);
  for (int o = 0; o < O; ++o)
    for (int j = 0; j < hop; ++j)
      for (int z = 0; z < dim; ++z) {
        double s = 0;
        for (int c = 0; c < OD; ++c)
          s += static_cast<double>(ll.data[o * OD + c]) * rc.data[(static_cast<size_t>(z) * OD + c) * hop + j];
        rows[(static_cast<size_t>(o) * hop + j) * dim + z] = static_cast<float>(s);
      }
  std::vector<float> bias(O);
  for (int o = 0; o < O; ++o) {
    double s = lb.data[o];
    for (int c = 0; c < OD; ++c) s += static_cast<double>(ll.data[o * OD + c]) * rb.data[c];
    bias[o] = static_cast<float>(s);
  }
  n.head.n_out = O; n.head.hop = hop;
  n.head.w = make_gemm_w(W, rows, O * hop, dim, false, pick_block_n(O * hop, hop), bias.data(), O);
}

// STFT branch + 1x1 + residual add (modules/seanet.py:463-507): Aout = ELU((X + spec) * act_scale)
// STFT -> log-magnitude -> normalise (modules/conv.py:1036-1080, seanet.py:463-497): Y [B*F, ldy] fp16
Buf plan_spec_stft(PlanCtx& c, const SpecW& s, const Buf& wav16, int pitch, int lead, int F, const std::string& name, int& ldy_out) {
  const long long M = static_cast<long long>(c.B) * F;
  const int K2 = s.n_fft / 2 + 1;
  // row pitch of the log-spectrogram: next multiple of 8 elements (16 B, the TMA stride unit); pad
  // columns are written as zeros.  (A 64-byte-aligned pitch was measured: the 1x1 gains 10 % but the
  // STFT epilogue loses more to the extra pad stores.)
  const int ldy = static_cast<int>(round_up(K2, g_ldy_align));
  Buf Y = c.alloc(static_cast<size_t>(M) * ldy * 2);
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.out_raw = c.ptr<h16>(Y);
  g.ldo = ldy;
  g.log_offset = s.mean + std::log(WAV_FP16_SCALE);
  g.inv_sigma = 1.f / s.stdv;
  g.clamp_sq = (1e-5f * WAV_FP16_SCALE) * (1e-5f * WAV_FP16_SCALE);
  g.n_half = s.n_fft / 2;
  const int base = lead - (s.n_fft - 1);       // first sample of frame 0 inside the staged clip
  const uint64_t clip_stride = static_cast<uint64_t>(WAV_COPIES) * pitch;   // 8 shifted copies per clip
  if ((s.hop * 2) % 16 == 0) {
    CUtensorMap tm;
    if (!c.dry())
      tm = make_tmap(c.ptr<__half>(wav16) + base, 3, s.n_fft, F, c.B, s.hop, clip_stride, BK, BM, true);
    c.tag(name + ".stft");
    add_gemm(c, EPI_STFT, s.dft, nullptr, 0, 0, s.n_fft, g, &tm, F, c.B);
  } else if (WAV_COPIES % s.hop == 0 && g_phased_stft) {
    // hop < 8 samples: frame f = P*q + r starts at sample 8q + r*hop -> phase r = rows of copy r*hop
    // with a 16-byte stride.  One GEMM over B*P "virtual clips" of ceil(F/P) rows.
    const int P = WAV_COPIES / s.hop;
    const int Fq = ceil_div(F, P);
    g.phases = P;
    g.frames_per_clip = F;
    CUtensorMap tm;
    if (!c.dry())
      tm = make_tmap4(c.ptr<__half>(wav16) + base, s.n_fft, Fq, P, c.B, 8, static_cast<uint64_t>(s.hop) * pitch, clip_stride, BK, BM);
    c.tag(name + ".stft");
    add_gemm(c, EPI_STFT, s.dft, nullptr, 0, 0, s.n_fft, g, &tm, Fq, c.B * P);
    Op& op = c.ops->back();
    op.flops = 2.0 * static_cast<double>(M) * s.dft.N * s.n_fft;
    op.bytes = static_cast<double>(M) * (2.0 * 8 + (s.dft.N / 2 + 1) * 2.0) + static_cast<double>(s.dft.N) * s.n_fft * 2.0;
  } else {
    Buf FR = c.alloc(static_cast<size_t>(M) * s.n_fft * 2);
    Op op;
    op.type = OP_FRAMES;
    op.in = c.ptr<__half>(wav16); op.out0 = c.ptr<__half>(FR);
    op.i[0] = c.B; op.i[1] = F; op.i[2] = s.hop; op.i[3] = s.n_fft; op.i[4] = base; op.i[5] = static_cast<int>(clip_stride);
    op.grid = elem_grid(M * (s.n_fft / 8));
    op.bytes = static_cast<double>(M) * s.n_fft * 2.0 + static_cast<double>(c.B) * pitch * 2.0;
    op.out_bytes[0] = static_cast<size_t>(M) * s.n_fft * 2;
    c.tag(name + ".frames");
    c.push(op);
    c.tag(name + ".stft");
    add_gemm(c, EPI_STFT, s.dft, c.ptr<__half>(FR), s.n_fft, M, s.n_fft, g);
    c.release(FR);
  }
  ldy_out = ldy;
  return Y;
}

void plan_spec(PlanCtx& c, const SpecW& s, const Buf& wav16, int pitch, int lead, int F, Buf& X,
               int C, float act_scale, Buf& Aout, const std::string& name) {
  const long long M = static_cast<long long>(c.B) * F;
  int ldy = 0;
  Buf Y = plan_spec_stft(c, s, wav16, pitch, lead, F, name, ldy);
  Aout = c.alloc(static_cast<size_t>(M) * C * 2);
  c.tag(name + ".out");
  // contraction over the padded width ldy (pad columns of Y and of the weights are zero): a TMA
  // inner extent of 33/65/.. elements is not a multiple of 16 B and loads slowly
  add_gemm(c, EPI_STAGED, s.layer, c.ptr<h16>(Y), ldy, M, ldy,
           std_args(nullptr, c.ptr<h16>(X), nullptr, c.ptr<h16>(Aout), act_scale, C));
  c.release(Y);
  c.release(X);
}

// SEANetEncoder.forward (modules/seanet.py:883-976).  Leaves the latent h16 [B*F, dim].
void plan_encoder(PlanCtx& c, wv_net& n, Plan& plan) {
  const EncoderW& e = n.enc;
  const int B = c.B, T = c.T;
  const int lead = e.nfft_max - 1;
  const int pitch = static_cast<int>(round_up(lead + T, 8));
  Buf wav16 = c.alloc(static_cast<size_t>(B) * WAV_COPIES * pitch * 2 + 256);   // + slack: the last frame view may run a few samples past the last copy
  {
    Op op;
    op.type = OP_WAV_STAGE;
    op.out0 = c.ptr<__half>(wav16);
    op.fa = WAV_FP16_SCALE;
    op.i[0] = B; op.i[1] = T; op.i[2] = lead; op.i[3] = pitch;
    op.grid = elem_grid(static_cast<long long>(B) * WAV_COPIES * (pitch / 8));
    op.bytes = static_cast<double>(B) * (T * 4.0 + WAV_COPIES * pitch * 2.0);
    op.out_bytes[0] = static_cast<size_t>(B) * WAV_COPIES * pitch * 2;
    c.tag("enc.wav16");
    c.push(op);
  }
  Buf film;
  if (e.has_film) {
    film = c.alloc(static_cast<size_t>(B) * e.n_film * 2 * 4);
    Op op;
    op.type = OP_FILM;
    op.out0 = c.ptr<float>(film);
    op.fargs = e.film;
    op.i[0] = B;
    op.out_bytes[0] = static_cast<size_t>(B) * e.n_film * 2 * 4;
    c.tag("enc.film");
    c.push(op);
  }
  int Ts = T, C = e.C0;
  // the first resblock can recompute its residual (this conv's raw output) from the waveform: X is then never stored
  const bool pre_mode = g_pre_fuse && !e.stages.empty() && !e.stages[0].res.empty() && !resblock_fusable(e.stages[0].res[0], C) &&
                        (e.stages[0].res.size() >= 2 || e.stages[0].dual_split_kb > 0);
  Buf X = pre_mode ? Buf() : c.alloc(static_cast<size_t>(B) * Ts * C * 2), A = c.alloc(static_cast<size_t>(B) * Ts * C * 2);
  {
    Op op;
    op.type = OP_CONV_PRE;
    op.out0 = pre_mode ? nullptr : c.ptr<h16>(X); op.out1 = c.ptr<h16>(A);
    op.w = e.conv_pre.w; op.bias = e.conv_pre.bias;
    op.fa = e.stages[0].res[0].pre_scale;
    op.i[0] = B; op.i[1] = Ts; op.i[2] = C;
    op.grid = elem_grid(static_cast<long long>(B) * ceil_div(Ts, PRE_TT) * (C / 8));
    op.flops = 10.0 * B * Ts * C;
    op.bytes = static_cast<double>(B) * Ts * (4.0 + (pre_mode ? 2.0 : 4.0) * C);
    op.out_bytes[0] = pre_mode ? 0 : static_cast<size_t>(B) * Ts * C * 2;
    op.out_bytes[1] = static_cast<size_t>(B) * Ts * C * 2;
    c.tag("enc.pre");
    c.push(op);
  }
  const float down_scale = 1.f / std::sqrt(1.f + n.cfg.n_residual_enc * n.cfg.res_scale_enc * n.cfg.res_scale_enc);
  const int S = static_cast<int>(e.stages.size());
  for (int s = 0; s < S; ++s) {
    const EncStageW& st = e.stages[s];
    const int nres = static_cast<int>(st.res.size());
    const bool fuse_spec = st.dual_split_kb > 0 && nres >= 1 && !resblock_fusable(st.res[nres - 1], C);
    for (int j = 0; j < nres; ++j) {
      const bool last = j == nres - 1;
      const std::string rname = "enc.s" + std::to_string(s) + ".r" + std::to_string(j);
      if (last && fuse_spec) {
        // first half as usual; the second half also contracts the log-spectrogram with the spec 1x1
        // (second accumulator) and writes only A = ELU((x' + spec) * scale): x' never reaches HBM
        const ResW& r = st.res[j];
        const size_t bytes = static_cast<size_t>(B) * Ts * C * 2;
        Buf H = c.alloc(bytes);
        c.tag(rname + ".h1");
        add_gemm_dw(c, r.pw1, r.dw1, c.ptr<h16>(A), Ts, C, nullptr, nullptr, c.ptr<h16>(H), 1.f);
        c.release(A);
        int ldy = 0;
        Buf Y = plan_spec_stft(c, st.spec, wav16, pitch, lead, Ts, "enc.s" + std::to_string(s) + ".spec", ldy);
        if (ldy != st.dual_ldy) WV_THROW(WV_ERR_INVALID, "log-spectrogram pitch %d != %d", ldy, st.dual_ldy);
        A = c.alloc(bytes);
        const bool pre_here = pre_mode && s == 0 && j == 0;
        GemmArgs g = std_args(r.dw2.bias, pre_here ? nullptr : c.ptr<h16>(X), nullptr, c.ptr<h16>(A), down_scale, C);
        g.taps = 5;
        g.dw_w = r.dw2.w;
        g.dual = 1;
        g.a2_split = st.dual_split_kb;
        if (pre_here) { g.pre_w = e.conv_pre.w; g.pre_b = e.conv_pre.bias; g.pre_T = Ts; }
        CUtensorMap tm;
        if (!c.dry()) tm = make_tmap(c.ptr<h16>(H), 3, C, Ts, B, C, static_cast<uint64_t>(C) * Ts, BK, BM, false);
        c.tag(rname + ".out+spec");
        add_gemm(c, EPI_STAGED, st.dual, nullptr, 0, 0, st.dual.K, g, &tm, Ts, B);
        Op& op = c.ops->back();
        if (!c.dry()) op.tmR = make_tmap(c.ptr<h16>(Y), 3, ldy, Ts, B, ldy, static_cast<uint64_t>(ldy) * Ts, BK, BM, false);
        const double n = static_cast<double>(B) * Ts;
        op.flops = 2.0 * n * C * (C + ldy) + 10.0 * n * C;
        op.bytes = n * 2.0 * (C + ldy + (pre_here ? 0 : C) + C) + (pre_here ? n * 4.0 : 0.0) + static_cast<double>(C) * (C + ldy) * 2.0;   // h1, y, x in; A out
        op.out_bytes[0] = 0;
        op.out_bytes[1] = bytes;
        c.release(H);
        c.release(Y);
        if (X.valid) c.release(X);
        break;
      }
      Buf Xn, An;
      plan_resblock(c, st.res[j], X, A, Ts, C, true, !last, last ? 1.f : st.res[j + 1].pre_scale, Xn, An, rname,
                    (pre_mode && s == 0 && j == 0) ? &e.conv_pre : nullptr);
      X = Xn; A = An;
    }
    if (!fuse_spec)
      plan_spec(c, st.spec, wav16, pitch, lead, Ts, X, C, down_scale, A, "enc.s" + std::to_string(s) + ".spec");   // A = ELU((x+spec)*scale)
    const int To = ceil_div(Ts, st.r);
    const bool last_stage = s == S - 1;
    Buf Ain = A;
    X = c.alloc(static_cast<size_t>(B) * To * 2 * C * 2);
    if (!last_stage) A = c.alloc(static_cast<size_t>(B) * To * 2 * C * 2); else A = Buf();
    c.tag("enc.s" + std::to_string(s) + ".down");
    add_gemm_down(c, st.down_pw, st.down_dw, st.r, c.ptr<h16>(Ain), Ts, C,
                  e.has_film ? c.ptr<float>(film) + static_cast<size_t>(s) * e.bands * 2 : nullptr, e.n_film * 2, e.bands,
                  c.ptr<h16>(X), last_stage ? nullptr : c.ptr<h16>(A),
                  last_stage ? 1.f : e.stages[s + 1].res[0].pre_scale);
    c.release(Ain);
    Ts = To; C *= 2;
  }
  plan.F = Ts;
  plan_spec(c, e.spec_post, wav16, pitch, lead, Ts, X, C, 1.f, A, "enc.post.spec");         // ELU(x + spec)
  c.release(wav16);
  if (film.valid) c.release(film);
  const long long M = static_cast<long long>(B) * Ts;
  Buf D = c.alloc(static_cast<size_t>(M) * C * 2);
  c.tag("enc.post.dw");
  add_dw5(c, e.post_dw, c.ptr<h16>(A), nullptr, c.ptr<h16>(D), nullptr, 1.f, Ts, C);
  c.release(A);
  plan.latent = c.alloc(static_cast<size_t>(M) * e.dim * 2);
  GemmArgs g = std_args(e.post_pw.bias, nullptr, c.ptr<h16>(plan.latent), nullptr, 1.f, e.dim);
  g.l2_scale = std::sqrt(static_cast<float>(e.dim));                         // seanet.py:299
  g.f32_F = Ts;
  c.tag("enc.latent");
  add_gemm(c, EPI_L2NORM, e.post_pw, c.ptr<h16>(D), C, M, C, g);
  c.release(D);
}

// SEANetDecoder.forward (modules/seanet.py:1212-1227) + trim + watermark add.
void plan_decoder(PlanCtx& c, wv_net& n, const Buf& Z, int F, int T_out) {
  const DecoderW& d = n.dec;
  const int B = c.B;
  int Ts = F, C = d.pw0.N;
  long long M = static_cast<long long>(B) * Ts;
  Buf A = c.alloc(static_cast<size_t>(M) * C * 2);
  c.tag("dec.in");
  add_gemm_dw(c, d.pw0, d.dw0, c.ptr<h16>(Z), Ts, d.pw0.K, nullptr, nullptr, c.ptr<h16>(A), 1.f);   // ELU of stage 0
  for (size_t s = 0; s < d.stages.size(); ++s) {
    const DecStageW& st = d.stages[s];
    const int To = Ts * st.r;
    if (st.up_halve.w != nullptr) {   // fused transposed conv + 1x1: one GEMM over the low-rate rows
      const int Ch = C / 2, Nf = st.r * Ch;
      Buf X = c.alloc(static_cast<size_t>(B) * To * Ch * 2);
      Buf An = c.alloc(static_cast<size_t>(B) * To * Ch * 2);
      GemmArgs g = std_args(st.up_halve.bias, nullptr, c.ptr<h16>(X), c.ptr<h16>(An),
                            st.res.empty() ? d.stage_scale : st.res[0].pre_scale, Nf);
      g.kb_split = C / BK;
      CUtensorMap tm;
      if (!c.dry()) tm = make_tmap(c.ptr<h16>(A), 3, C, Ts, B, C, static_cast<uint64_t>(C) * Ts, BK, BM, true);
      c.tag("dec.u" + std::to_string(s) + ".uphalve");
      add_gemm(c, EPI_STAGED, st.up_halve, nullptr, 0, 0, 2 * C, g, &tm, Ts, B);
      Op& op = c.ops->back();
      op.bytes = static_cast<double>(B) * Ts * C * 2.0 + 2.0 * B * To * Ch * 2.0 + static_cast<double>(Nf) * 2 * C * 2.0;
      c.release(A);
      A = An;
      Ts = To;
      M = static_cast<long long>(B) * Ts;
      C = Ch;
      const int nres = static_cast<int>(st.res.size());
      for (int j = 0; j < nres; ++j) {
        const bool last = j == nres - 1;
        Buf Xn, An2;
        plan_resblock(c, st.res[j], X, A, Ts, C, !last, true, last ? d.stage_scale : st.res[j + 1].pre_scale, Xn, An2,
                      "dec.u" + std::to_string(s) + ".r" + std::to_string(j));
        X = Xn; A = An2;
      }
      if (nres == 0) c.release(X);
      continue;
    }
    Buf U = c.alloc(static_cast<size_t>(B) * To * C * 2);
    {
      Op op;
      op.type = OP_UP;
      op.in = c.ptr<h16>(A); op.out0 = c.ptr<h16>(U); op.w = st.up.w;
      op.i[0] = B; op.i[1] = Ts; op.i[2] = C; op.i[3] = st.r;
      op.grid = elem_grid(static_cast<long long>(B) * Ts * (C / 8));
      op.flops = 4.0 * B * To * C;
      op.bytes = 2.0 * B * (static_cast<double>(Ts) + To) * C;
      op.out_bytes[0] = static_cast<size_t>(B) * To * C * 2;
      c.tag("dec.u" + std::to_string(s) + ".up");
      c.push(op);
    }
    c.release(A);
    Ts = To;
    M = static_cast<long long>(B) * Ts;
    const int Ch = C / 2;
    Buf X = c.alloc(static_cast<size_t>(M) * Ch * 2);
    A = c.alloc(static_cast<size_t>(M) * Ch * 2);
    c.tag("dec.u" + std::to_string(s) + ".halve");
    add_gemm(c, EPI_STAGED, st.halve, c.ptr<h16>(U), C, M, C,
             std_args(st.halve.bias, nullptr, c.ptr<h16>(X), c.ptr<h16>(A), st.res.empty() ? d.stage_scale : st.res[0].pre_scale, Ch));
    c.release(U);
    C = Ch;
    const int nres = static_cast<int>(st.res.size());
    for (int j = 0; j < nres; ++j) {
      const bool last = j == nres - 1;
      Buf Xn, An;
      plan_resblock(c, st.res[j], X, A, Ts, C, !last, true, last ? d.stage_scale : st.res[j + 1].pre_scale, Xn, An,
                    "dec.u" + std::to_string(s) + ".r" + std::to_string(j));
      X = Xn; A = An;
    }
    if (nres == 0) c.release(X);
  }
  if (g_last_gemm) {
    // output conv C -> 1, k = 5 as a tcgen05 GEMM with one column per tap; the epilogue sums the five
    // shifted columns, applies tanh, trims to T_out and adds the watermark to x (io pointers at run time)
    GemmArgs g = std_args(nullptr, nullptr, nullptr, nullptr, 1.f, 32);
    g.taps = 5;                // tile geometry: 4-row causal halo, 124 outputs per tile
    g.last_mode = 1;
    g.last_T = T_out;
    g.last_bias = d.last_b;
    CUtensorMap tm;
    if (!c.dry()) tm = make_tmap(c.ptr<h16>(A), 3, C, Ts, B, C, static_cast<uint64_t>(C) * Ts, BK, BM, true);
    c.tag("dec.last");
    add_gemm(c, EPI_STAGED, d.last_gemm, nullptr, 0, 0, C, g, &tm, Ts, B);
    Op& op = c.ops->back();
    op.flops = 2.0 * 5 * C * B * T_out;
    op.bytes = static_cast<double>(B) * T_out * (2.0 * C + 12.0);
  } else {
    Op op;
    op.type = OP_CONV_LAST;
    op.in = c.ptr<h16>(A); op.w = d.last_w; op.fa = d.last_b;
    op.i[0] = B; op.i[1] = Ts; op.i[2] = T_out; op.i[3] = C;
    op.grid = B * ceil_div(T_out, CL_TILE);
    op.flops = 2.0 * 5 * C * B * T_out;
    op.bytes = static_cast<double>(B) * T_out * (2.0 * C + 12.0);
    c.tag("dec.last");
    c.push(op);
  }
  c.release(A);
}

void plan_head(PlanCtx& c, wv_net& n, const Buf& Z, int F) {
  const HeadW& h = n.head;
  const long long M = static_cast<long long>(c.B) * F;
  const int tiles_n = h.w.N / h.w.block_n;
  Buf partial = c.alloc(static_cast<size_t>(M) * tiles_n * EPI_SPLIT * 4);   // one slot per column split
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.bias = h.w.bias;
  g.partial = n.cfg.kind == WV_KIND_DETECTOR ? c.ptr<float>(partial) : nullptr;
  g.hop = h.hop; g.T = c.T; g.n_out = h.n_out; g.head_F = F;
  c.tag("head.gemm");
  // precise nets: the latent is a split pair [M, 2*dim]
  add_gemm(c, EPI_HEAD, h.w, c.ptr<h16>(Z), n.cfg.precise ? 2 * n.enc.dim : h.w.K, M, h.w.K, g);
  if (n.cfg.kind == WV_KIND_DETECTOR) {   // bit decode exists for the detector only
    Op op;
    op.type = OP_BITS;
    op.in = c.ptr<float>(partial);
    op.i[0] = c.B; op.i[1] = F; op.i[2] = EPI_SPLIT * tiles_n; op.i[3] = EPI_SPLIT * (h.hop / h.w.block_n); op.i[4] = c.T; op.i[5] = h.n_out;
    c.tag("head.bits");
    c.push(op);
    Op op2;
    op2.type = OP_CONF;
    op2.i[0] = c.B; op2.i[1] = h.n_out;
    c.tag("head.conf");
    c.push(op2);
  }
  c.release(partial);
}

// ------------------------------------------------------------------------------------------
// PRECISE nets (cfg.precise): the same SEANet encoder + head with fp32-accurate arithmetic (gemm_sm100.cuh,
// "PRECISE MODE").  Raw streams X are fp32 [B, T, C]; every tensor that feeds a GEMM is a split-fp16 pair
// [B, T, 2C] = [hi | lo].  No launch fusion beyond the GEMM epilogues (1x1 + dw5, 1x1 + down-conv).
GemmArgs pm_args(const float* bias, const float* res32, float* raw32, h16* act, float act_scale, int C) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.bias = bias; g.residual32 = res32; g.out_raw32 = raw32; g.out_act = act;
  g.act_scale = act_scale; g.ldo = C; g.ldo_act = 2 * C; g.lo_off = C;
  g.taps = 1;
  return g;
}

// 1x1 conv over a split tensor A [B, T, 2K] + causal depthwise k=5 (+ fp32 residual) -> raw32 [B,T,N] / act split [B,T,2N]
void add_gemm_dw_pm(PlanCtx& c, const GemmW& pw, const DwW& dw, const h16* A, int T, int Kc, const float* res32,
                    float* raw32, h16* act, float act_scale) {
  if (dw.k != 5 || dw.C != pw.N) WV_THROW(WV_ERR_INVALID, "fused depthwise conv must be k=5 over the GEMM's N");
  GemmArgs g = pm_args(dw.bias, res32, raw32, act, act_scale, pw.N);
  g.taps = 5;
  g.dw_w = dw.w;
  CUtensorMap tm;
  if (!c.dry()) tm = make_tmap(A, 3, 2 * Kc, T, c.B, 2 * Kc, static_cast<uint64_t>(2 * Kc) * T, BK, BM, false);
  add_gemm(c, EPI_STAGED_PM, pw, nullptr, 0, 0, pw.K, g, &tm, T, c.B);
}

void add_gemm_down_pm(PlanCtx& c, const GemmW& pw, const DwW& dw, int r, const h16* A, int T, int Kc, float* raw32,
                      h16* act, float act_scale) {
  if (dw.k != 2 * r || dw.C != pw.N) WV_THROW(WV_ERR_INVALID, "fused down-conv must be k=2r over the GEMM's N");
  if (r != 2 && r != 4 && r != 5 && r != 8) WV_THROW(WV_ERR_UNSUPPORTED, "fused down-conv stride %d (2, 4, 5, 8)", r);
  GemmArgs g = pm_args(dw.bias, nullptr, raw32, act, act_scale, pw.N);
  g.down_r = r;
  g.dw_w = dw.w;
  CUtensorMap tm;
  if (!c.dry()) tm = make_tmap(A, 3, 2 * Kc, T, c.B, 2 * Kc, static_cast<uint64_t>(2 * Kc) * T, BK, BM, false);
  add_gemm(c, EPI_STAGED_PM, pw, nullptr, 0, 0, pw.K, g, &tm, T, c.B);
}

// STFT -> log-magnitude -> normalise as split pairs Y [B*F, 2*ldy]; frames of the hi / lo waveform copies
Buf plan_spec_stft_pm(PlanCtx& c, const SpecW& s, const Buf& wavh, const Buf& wavl, int pitch, int lead, int F,
                      const std::string& name, int& ldy_out) {
  const long long M = static_cast<long long>(c.B) * F;
  const int ldy = s.n_fft / 2 + 8;
  Buf Y = c.alloc(static_cast<size_t>(M) * 2 * ldy * 2);
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.out_raw = c.ptr<h16>(Y);
  g.ldo = 2 * ldy;
  g.lo_off = ldy;
  g.log_offset = s.mean + std::log(WAV_FP16_SCALE);
  g.inv_sigma = 1.f / s.stdv;
  g.clamp_sq = (1e-5f * WAV_FP16_SCALE) * (1e-5f * WAV_FP16_SCALE);
  g.n_half = s.n_fft / 2;
  const int base = lead - (s.n_fft - 1);
  const uint64_t clip_stride = static_cast<uint64_t>(WAV_COPIES) * pitch;
  CUtensorMap tmh, tml;
  c.tag(name + ".stft");
  if ((s.hop * 2) % 16 == 0) {
    if (!c.dry()) {
      tmh = make_tmap(c.ptr<__half>(wavh) + base, 3, s.n_fft, F, c.B, s.hop, clip_stride, BK, BM, true);
      tml = make_tmap(c.ptr<__half>(wavl) + base, 3, s.n_fft, F, c.B, s.hop, clip_stride, BK, BM, true);
    }
    add_gemm(c, EPI_STFT_PM, s.dft, nullptr, 0, 0, s.dft.K, g, &tmh, F, c.B);
  } else if (WAV_COPIES % s.hop == 0) {
    const int P = WAV_COPIES / s.hop;
    const int Fq = ceil_div(F, P);
    g.phases = P;
    g.frames_per_clip = F;
    if (!c.dry()) {
      tmh = make_tmap4(c.ptr<__half>(wavh) + base, s.n_fft, Fq, P, c.B, 8, static_cast<uint64_t>(s.hop) * pitch, clip_stride, BK, BM);
      tml = make_tmap4(c.ptr<__half>(wavl) + base, s.n_fft, Fq, P, c.B, 8, static_cast<uint64_t>(s.hop) * pitch, clip_stride, BK, BM);
    }
    add_gemm(c, EPI_STFT_PM, s.dft, nullptr, 0, 0, s.dft.K, g, &tmh, Fq, c.B * P);
    c.ops->back().flops = 2.0 * static_cast<double>(M) * s.dft.N * s.n_fft;
  } else {
    WV_THROW(WV_ERR_UNSUPPORTED, "precise STFT with hop %d", s.hop);
  }
  Op& op = c.ops->back();
  if (!c.dry()) op.tmR = tml;
  op.bytes = static_cast<double>(M) * (4.0 * 8 + (s.dft.N / 2 + 1) * 4.0) + static_cast<double>(s.dft.N) * s.dft.K * 2.0;
  op.out_bytes[0] = static_cast<size_t>(M) * 2 * ldy * 2;
  ldy_out = ldy;
  return Y;
}

// x += spec branch, then the activated split stream: Aout = split(ELU((X + Ws y) * act_scale)); releases X
void plan_spec_pm(PlanCtx& c, const SpecW& s, const Buf& wavh, const Buf& wavl, int pitch, int lead, int F, Buf& X,
                  int C, float act_scale, Buf& Aout, const std::string& name) {
  const long long M = static_cast<long long>(c.B) * F;
  int ldy = 0;
  Buf Y = plan_spec_stft_pm(c, s, wavh, wavl, pitch, lead, F, name, ldy);
  Aout = c.alloc(static_cast<size_t>(M) * 2 * C * 2);
  c.tag(name + ".out");
  add_gemm(c, EPI_STAGED_PM, s.layer, c.ptr<h16>(Y), 2 * ldy, M, s.layer.K,
           pm_args(nullptr, c.ptr<float>(X), nullptr, c.ptr<h16>(Aout), act_scale, C));
  c.release(Y);
  c.release(X);
}

void plan_encoder_pm(PlanCtx& c, wv_net& n, Plan& plan) {
  const EncoderW& e = n.enc;
  const int B = c.B, T = c.T;
  const int lead = e.nfft_max - 1;
  const int pitch = static_cast<int>(round_up(lead + T, 8));
  const size_t wav_bytes = static_cast<size_t>(B) * WAV_COPIES * pitch * 2 + 256;
  Buf wavh = c.alloc(wav_bytes), wavl = c.alloc(wav_bytes);
  {
    Op op;
    op.type = OP_WAV_STAGE_PM;
    op.out0 = c.ptr<__half>(wavh); op.out1 = c.ptr<__half>(wavl);
    op.fa = WAV_FP16_SCALE;
    op.i[0] = B; op.i[1] = T; op.i[2] = lead; op.i[3] = pitch;
    op.grid = elem_grid(static_cast<long long>(B) * WAV_COPIES * (pitch / 8));
    op.bytes = static_cast<double>(B) * (T * 4.0 + 2.0 * WAV_COPIES * pitch * 2.0);
    op.out_bytes[0] = op.out_bytes[1] = static_cast<size_t>(B) * WAV_COPIES * pitch * 2;
    c.tag("enc.wav16");
    c.push(op);
  }
  int Ts = T, C = e.C0;
  auto xbytes = [&](int t, int ch) { return static_cast<size_t>(B) * t * ch * 4; };   // fp32 raw == split pair bytes
  Buf X = c.alloc(xbytes(Ts, C)), A = c.alloc(xbytes(Ts, C));
  {
    Op op;
    op.type = OP_CONV_PRE_PM;
    op.out0 = c.ptr<float>(X); op.out1 = c.ptr<h16>(A);
    op.w = e.conv_pre.w; op.bias = e.conv_pre.bias;
    op.fa = e.stages[0].res[0].pre_scale;
    op.i[0] = B; op.i[1] = Ts; op.i[2] = C;
    op.grid = elem_grid(static_cast<long long>(B) * ceil_div(Ts, PRE_TT) * (C / 4));
    op.flops = 10.0 * B * Ts * C;
    op.bytes = static_cast<double>(B) * Ts * (4.0 + 8.0 * C);
    op.out_bytes[0] = op.out_bytes[1] = xbytes(Ts, C);
    c.tag("enc.pre");
    c.push(op);
  }
  const float down_scale = 1.f / std::sqrt(1.f + n.cfg.n_residual_enc * n.cfg.res_scale_enc * n.cfg.res_scale_enc);
  const int S = static_cast<int>(e.stages.size());
  for (int s = 0; s < S; ++s) {
    const EncStageW& st = e.stages[s];
    const int nres = static_cast<int>(st.res.size());
    for (int j = 0; j < nres; ++j) {
      const ResW& r = st.res[j];
      const bool last = j == nres - 1;
      const std::string rname = "enc.s" + std::to_string(s) + ".r" + std::to_string(j);
      Buf H = c.alloc(xbytes(Ts, C));
      c.tag(rname + ".h1");
      add_gemm_dw_pm(c, r.pw1, r.dw1, c.ptr<h16>(A), Ts, C, nullptr, nullptr, c.ptr<h16>(H), 1.f);
      c.release(A);
      Buf Xn = c.alloc(xbytes(Ts, C));
      Buf An = last ? Buf() : c.alloc(xbytes(Ts, C));
      c.tag(rname + ".out");
      add_gemm_dw_pm(c, r.pw2, r.dw2, c.ptr<h16>(H), Ts, C, c.ptr<float>(X), c.ptr<float>(Xn),
                     last ? nullptr : c.ptr<h16>(An), last ? 1.f : st.res[j + 1].pre_scale);
      c.release(H);
      c.release(X);
      X = Xn; A = An;
    }
    plan_spec_pm(c, st.spec, wavh, wavl, pitch, lead, Ts, X, C, down_scale, A, "enc.s" + std::to_string(s) + ".spec");
    const int To = ceil_div(Ts, st.r);
    const bool last_stage = s == S - 1;
    Buf Ain = A;
    X = c.alloc(xbytes(To, 2 * C));
    A = last_stage ? Buf() : c.alloc(xbytes(To, 2 * C));
    c.tag("enc.s" + std::to_string(s) + ".down");
    add_gemm_down_pm(c, st.down_pw, st.down_dw, st.r, c.ptr<h16>(Ain), Ts, C, c.ptr<float>(X),
                     last_stage ? nullptr : c.ptr<h16>(A), last_stage ? 1.f : e.stages[s + 1].res[0].pre_scale);
    c.release(Ain);
    Ts = To; C *= 2;
  }
  plan.F = Ts;
  plan_spec_pm(c, e.spec_post, wavh, wavl, pitch, lead, Ts, X, C, 1.f, A, "enc.post.spec");
  c.release(wavh);
  c.release(wavl);
  const long long M = static_cast<long long>(B) * Ts;
  Buf D = c.alloc(static_cast<size_t>(M) * 2 * C * 2);
  {
    Op op;
    op.type = OP_DW5_PM;
    op.in = c.ptr<h16>(A); op.out0 = c.ptr<h16>(D);
    op.w = e.post_dw.w; op.bias = e.post_dw.bias;
    op.i[0] = B; op.i[1] = Ts; op.i[2] = C;
    op.grid = elem_grid(M * (C / 4));
    op.flops = 10.0 * M * C;
    op.bytes = 8.0 * M * C;
    op.out_bytes[0] = static_cast<size_t>(M) * 2 * C * 2;
    c.tag("enc.post.dw");
    c.push(op);
  }
  c.release(A);
  plan.latent = c.alloc(static_cast<size_t>(M) * 2 * e.dim * 2);
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.bias = e.post_pw.bias;
  g.out_raw = c.ptr<h16>(plan.latent);
  g.ldo = 2 * e.dim; g.lo_off = e.dim;
  g.l2_scale = std::sqrt(static_cast<float>(e.dim));
  g.f32_F = Ts;
  c.tag("enc.latent");
  add_gemm(c, EPI_L2NORM_PM, e.post_pw, c.ptr<h16>(D), 2 * C, M, e.post_pw.K, g);
  c.release(D);
}

void build_plan_pass(wv_net& n, Plan& plan, uint8_t* base) {
  plan.ops.clear();
  PlanCtx c;
  c.base = base; c.ops = &plan.ops; c.B = plan.B; c.T = plan.T;
  if (n.cfg.precise) plan_encoder_pm(c, n, plan);
  else plan_encoder(c, n, plan);
  if (n.cfg.kind == WV_KIND_GENERATOR) plan_decoder(c, n, plan.latent, plan.F, plan.T);
  else plan_head(c, n, plan.latent, plan.F);
  plan.ws_bytes = c.arena.high;
}

void ensure_ws(wv_net& n, size_t bytes) {
  if (bytes <= n.ws_bytes) return;
  n.last_plan = nullptr;
  CK(cudaDeviceSynchronize());   // launches of cached plans may still be in flight on the old workspace
  // the workspace moves: every cached plan holds pointers/tensor maps into the old one
  n.plans.clear();
  n.dec_plans.clear();
  n.refine.clear();
  if (n.ws) CK(cudaFree(n.ws));
  n.ws = nullptr; n.ws_bytes = 0;
  CK(cudaMalloc(reinterpret_cast<void**>(&n.ws), bytes));
  n.ws_bytes = bytes;
}

// The file API and the request batcher feed arbitrary clip lengths: every distinct (B, T) costs a plan (launch list
// with encoded tensor maps) and, for small calls, CUDA graphs with their own I/O staging.  The caches are bounded:
// least-recently-used plans beyond MAX_CACHED_PLANS are dropped (the two plans of one sub-batched call are always the
// most recent ones).
constexpr size_t MAX_CACHED_PLANS = 24;
template <typename Map>
void evict_plans(wv_net& n, Map& plans) {
  while (plans.size() > MAX_CACHED_PLANS) {
    auto victim = plans.begin();
    for (auto it = plans.begin(); it != plans.end(); ++it)
      if (it->second->stamp < victim->second->stamp) victim = it;
    if (n.last_plan == victim->second.get()) n.last_plan = nullptr;
    CK(cudaDeviceSynchronize());   // its graphs / staging buffers may still be in flight
    plans.erase(victim);
  }
}

Plan& get_plan(wv_net& n, int B, int T) {
  auto key = std::make_pair(B, T);
  auto it = n.plans.find(key);
  if (it != n.plans.end()) { it->second->stamp = ++n.clock; return *it->second; }
  evict_plans(n, n.plans);
  auto plan = std::make_unique<Plan>();
  plan->stamp = ++n.clock;
  plan->B = B; plan->T = T;
  build_plan_pass(n, *plan, nullptr);
  ensure_ws(n, plan->ws_bytes);
  build_plan_pass(n, *plan, n.ws);
  Plan& ref = *plan;
  n.plans[key] = std::move(plan);
  return ref;
}

Plan& get_dec_plan(wv_net& n, int B, int F) {
  auto key = std::make_pair(B, F);
  auto it = n.dec_plans.find(key);
  if (it != n.dec_plans.end()) { it->second->stamp = ++n.clock; return *it->second; }
  evict_plans(n, n.dec_plans);
  auto build = [&](Plan& plan, uint8_t* base) {
    plan.ops.clear();
    PlanCtx c;
    c.base = base; c.ops = &plan.ops; c.B = B; c.T = F * n.enc.hop;
    plan.latent = c.alloc(static_cast<size_t>(B) * F * n.enc.dim * 2);
    Op op;
    op.type = OP_LATENT_IN;
    op.out0 = c.ptr<h16>(plan.latent);
    op.i[0] = B; op.i[1] = n.enc.dim; op.i[2] = F;
    op.grid = elem_grid(static_cast<long long>(B) * F * n.enc.dim);
    c.push(op);
    plan_decoder(c, n, plan.latent, F, F * n.enc.hop);
    plan.ws_bytes = c.arena.high;
  };
  auto plan = std::make_unique<Plan>();
  plan->stamp = ++n.clock;
  plan->B = B; plan->T = F * n.enc.hop; plan->F = F;
  build(*plan, nullptr);
  ensure_ws(n, plan->ws_bytes);
  build(*plan, n.ws);
  Plan& ref = *plan;
  n.dec_plans[key] = std::move(plan);
  return ref;
}

// Launch with programmatic stream serialization (PDL): the kernel's prologue may overlap the tail of
// the previous kernel in the stream; every kernel calls griddepcontrol.wait before touching data.
thread_local int g_launch_cluster = 1;   // cluster size of the next launch_k call (2 = CTA pairs; reset after the launch)
template <typename... KArgs, typename... Args>
void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (g_pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (g_launch_cluster > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = static_cast<unsigned>(g_launch_cluster); at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
    g_launch_cluster = 1;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  CK(cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...));
}

void launch_down(int grid, cudaStream_t st, const h16* in, const float* w, const float* bias, const float* film,
                 int film_stride, int bands, h16* out_raw, h16* out_act, float act_scale, int B, int Tin, int Tout,
                 int C, int r) {
  switch (r) {
#define WV_DOWN_CASE(R)                                                                                      \
  case R:                                                                                                    \
    launch_k(down_kernel<R>, grid, 256, 0, st, in, w, bias, film, film_stride, bands, out_raw, out_act,      \
             act_scale, B, Tin, Tout, C);                                                                    \
    break;
    WV_DOWN_CASE(2) WV_DOWN_CASE(3) WV_DOWN_CASE(4) WV_DOWN_CASE(5) WV_DOWN_CASE(6) WV_DOWN_CASE(8)
#undef WV_DOWN_CASE
    default: WV_THROW(WV_ERR_UNSUPPORTED, "stride %d is not supported (2, 3, 4, 5, 6, 8)", r);
  }
}

void launch_gemm(const Op& op, const GemmArgs& g, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(op.i[7]);
  switch (op.epi) {
    case EPI_STAGED:
      if (g.cg2) {
        g_launch_cluster = 2;
        launch_k(gemm_sm100_kernel<EPI_STAGED, true>, op.grid, STAGED_THREADS, smem, st, op.tmA, op.tmB, op.tmR, g);
      } else {
        launch_k(gemm_sm100_kernel<EPI_STAGED>, op.grid, STAGED_THREADS, smem, st, op.tmA, op.tmB, op.tmR, g);
      }
      break;
    case EPI_L2NORM: launch_k(gemm_sm100_kernel<EPI_L2NORM>, op.grid, GEMM_THREADS, smem, st, op.tmA, op.tmB, op.tmR, g); break;
    case EPI_STFT: launch_k(gemm_sm100_kernel<EPI_STFT>, op.grid, GEMM_THREADS, smem, st, op.tmA, op.tmB, op.tmR, g); break;
    case EPI_STAGED_PM: launch_k(gemm_sm100_kernel<EPI_STAGED_PM>, op.grid, STAGED_THREADS, smem, st, op.tmA, op.tmB, op.tmR, g); break;
    case EPI_L2NORM_PM: launch_k(gemm_sm100_kernel<EPI_L2NORM_PM>, op.grid, GEMM_THREADS, smem, st, op.tmA, op.tmB, op.tmR, g); break;
    case EPI_STFT_PM: launch_k(gemm_sm100_kernel<EPI_STFT_PM>, op.grid, GEMM_THREADS, smem, st, op.tmA, op.tmB, op.tmR, g); break;
    default: launch_k(gemm_sm100_kernel<EPI_HEAD>, op.grid, GEMM_THREADS, smem, st, op.tmA, op.tmB, op.tmR, g); break;
  }
}

void run_plan(wv_net& n, Plan& plan, const IoPtrs& io, cudaStream_t st, int stop_after = -1) {
  const bool prof = n.profile;
  if (prof) {
    while (plan.events.size() < plan.ops.size() + 1) {
      cudaEvent_t ev;
      CK(cudaEventCreate(&ev));
      plan.events.push_back(ev);
    }
    CK(cudaEventRecord(plan.events[0], st));
    n.last_plan = &plan;
  }
  int op_index = -1;
  for (const Op& op : plan.ops) {
    ++op_index;
    switch (op.type) {
      case OP_GEMM: {
        GemmArgs g = op.g;
        if (op.epi == EPI_L2NORM || op.epi == EPI_L2NORM_PM) g.out_f32_t = io.latent;
        if (g.last_mode) { g.last_x = io.x; g.last_wm = io.wm; g.last_y = io.y; }
        if (g.pre_w != nullptr) g.pre_x = io.x;
        if (op.epi == EPI_HEAD) {
          g.logits = io.logits; g.mask_out = io.mask; g.probs = io.probs; g.presence = io.presence;
          if (!(io.bits || io.avg || io.conf || io.valid)) g.partial = nullptr;
        }
        launch_gemm(op, g, st);
        break;
      }
      case OP_RESBLOCK:
        launch_k(resblock_sm100_kernel, op.grid, RB_THREADS, static_cast<size_t>(op.i[7]), st, op.tmA, op.tmB, op.tmR, op.rb);
        break;
      case OP_DW5:
        launch_k(dw5_kernel, op.grid, 256, 0, st, static_cast<const h16*>(op.in), op.w, op.bias, static_cast<const h16*>(op.res),
                 static_cast<h16*>(op.out0), static_cast<h16*>(op.out1), op.fa, op.i[0], op.i[1], op.i[2]);
        break;
      case OP_DOWN:
        launch_down(op.grid, st, static_cast<const h16*>(op.in), op.w, op.bias, op.film, op.i[5], op.i[6],
                    static_cast<h16*>(op.out0), static_cast<h16*>(op.out1), op.fa, op.i[0], op.i[1], op.i[2], op.i[3], op.i[4]);
        break;
      case OP_UP:
        launch_k(up_kernel, op.grid, 256, 0, st, static_cast<const h16*>(op.in), op.w, static_cast<h16*>(op.out0), op.i[0], op.i[1], op.i[2], op.i[3]);
        break;
      case OP_CONV_PRE:
        launch_k(conv_pre_kernel, op.grid, 256, 0, st, io.x, op.w, op.bias, static_cast<h16*>(op.out0), static_cast<h16*>(op.out1), op.fa, op.i[0], op.i[1], op.i[2]);
        break;
      case OP_CONV_LAST: {
        const int C = op.i[3];
        const size_t smem = ((5 * C * 4 + 15) & ~15) + static_cast<size_t>(CL_TILE + 4) * (C * 2 + 16);
        launch_k(conv_last_kernel, op.grid, CL_TILE, smem, st, static_cast<const h16*>(op.in), op.w, op.fa, io.x, io.wm, io.y, op.i[0], op.i[1], op.i[2], C);
        break;
      }
      case OP_WAV_STAGE:
        launch_k(wav_stage_kernel, op.grid, 256, 0, st, io.x, static_cast<__half*>(op.out0), op.fa, op.i[0], op.i[1], op.i[2], op.i[3]);
        break;
      case OP_FRAMES:
        launch_k(frames_kernel, op.grid, 256, 0, st, static_cast<const __half*>(op.in), static_cast<__half*>(op.out0), op.i[0], op.i[1], op.i[2], op.i[3], op.i[4], op.i[5]);
        break;
      case OP_FILM:
        launch_k(film_kernel, op.i[0], std::max(64, op.fargs.E), 2 * op.fargs.E * sizeof(float), st, io.msg, static_cast<float*>(op.out0), op.fargs);
        break;
      case OP_BITS:
        if (io.bits || io.avg || io.conf || io.valid) {
          // avg is needed for conf: use caller's buffer or skip conf when absent
          launch_k(bits_finish_kernel, dim3(op.i[0], op.i[5]), 32, 0, st, static_cast<const float*>(op.in), op.i[1], op.i[2], op.i[3], io.presence, op.i[4], op.i[5], io.bits, io.avg, io.valid);
        }
        break;
      case OP_CONF:
        if (io.conf && io.avg) launch_k(conf_kernel, ceil_div(op.i[0], 128), 128, 0, st, io.avg, io.conf, op.i[0], op.i[1]);
        break;
      case OP_CONV_PRE_PM:
        launch_k(conv_pre_pm_kernel, op.grid, 256, 0, st, io.x, op.w, op.bias, static_cast<float*>(op.out0), static_cast<h16*>(op.out1), op.fa, op.i[0], op.i[1], op.i[2]);
        break;
      case OP_DW5_PM:
        launch_k(dw5_pm_kernel, op.grid, 256, 0, st, static_cast<const h16*>(op.in), op.w, op.bias, static_cast<h16*>(op.out0), op.i[0], op.i[1], op.i[2]);
        break;
      case OP_WAV_STAGE_PM:
        launch_k(wav_stage_pm_kernel, op.grid, 256, 0, st, io.x, static_cast<__half*>(op.out0), static_cast<__half*>(op.out1), op.fa, op.i[0], op.i[1], op.i[2], op.i[3]);
        break;
      case OP_LATENT_IN:
        launch_k(latent_in_kernel, op.grid, 256, 0, st, io.z_in, static_cast<h16*>(op.out0), op.i[0], op.i[1], op.i[2]);
        break;
    }
    if (n.check_range && n.range_stats != nullptr) {
      // fp16 outputs only: the precise nets store their raw streams as fp32 (out0 of their STAGED / conv_pre launches)
      const bool out0_f32 = (op.type == OP_GEMM && op.epi == EPI_STAGED_PM) || op.type == OP_CONV_PRE_PM || op.type == OP_FILM;
      const void* outs[2] = {out0_f32 ? nullptr : op.out0, op.out1};
      for (int w = 0; w < 2; ++w)
        if (outs[w] != nullptr && op.out_bytes[w] >= 16) {
          const long long n16 = static_cast<long long>(op.out_bytes[w] / 16);
          range_check_kernel<<<elem_grid(n16), 256, 0, st>>>(static_cast<const uint4*>(outs[w]), n16, n.range_stats);
          if (getenv("WV_RANGE_VERBOSE")) {   // per-launch report (diagnostics): running totals after this output
            unsigned long long h[4];
            CK(cudaStreamSynchronize(st));
            CK(cudaMemcpy(h, n.range_stats, sizeof(h), cudaMemcpyDeviceToHost));
            const uint16_t bits = static_cast<uint16_t>(h[2]);
            __half hv;
            memcpy(&hv, &bits, 2);
            fprintf(stderr, "[wv range] %-24s out%d: saturated %llu non-finite %llu max|v| %g (running)\n", op.tag.c_str(), w, h[0], h[1], __half2float(hv));
          }
        }
    }
    if (prof) CK(cudaEventRecord(plan.events[op_index + 1], st));
    if (op_index == stop_after) break;
  }
  CK(cudaGetLastError());
}

int sub_batch(const wv_net& n, int B, int T) {
  if (n.chunk_samples <= 0) return B;
  long long bc = n.chunk_samples / std::max(1, T);
  return static_cast<int>(std::max<long long>(1, std::min<long long>(B, bc)));
}

template <typename Fn>
int guarded(Fn&& fn) {
  try {
    fn();
    return WV_OK;
  } catch (const WvError& e) {
    return fail(e.code, e.msg);
  } catch (const std::exception& e) {
    return fail(WV_ERR_INVALID, e.what());
  }
}

struct DeviceGuard {
  int prev = 0;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    want = dev;
  }
  ~DeviceGuard() {
    if (prev != want) cudaSetDevice(prev);
  }
  int want;
};

}  // namespace

// ==========================================================================================
extern "C" {

int wv_version(void) { return 1; }
const char* wv_last_error(void) { return g_err.c_str(); }

int wv_net_create(const wv_net_config* cfg, const wv_tensor* tensors, int n_tensors, int device, wv_net** out) {
  if (!cfg || !tensors || !out) return fail(WV_ERR_INVALID, "null argument");
  *out = nullptr;
  return guarded([&] {
    DeviceGuard dg(device);
    init_device_once();
    std::unique_ptr<wv_net> n(new wv_net());
    n->cfg = *cfg;
    n->device = device;
    if (cfg->n_strides < 1 || cfg->n_strides > 4) WV_THROW(WV_ERR_UNSUPPORTED, "n_strides must be 1..4");
    if (cfg->n_residual_enc < 1) WV_THROW(WV_ERR_UNSUPPORTED, "n_residual_enc must be >= 1");
    if (cfg->channels_enc % 32 != 0) WV_THROW(WV_ERR_UNSUPPORTED, "channels_enc must be a multiple of 32");
    if (cfg->precise && cfg->kind == WV_KIND_GENERATOR) WV_THROW(WV_ERR_UNSUPPORTED, "precise mode covers the detector / locator (thresholded outputs) only");
    struct PmScope { bool prev; explicit PmScope(bool v) : prev(g_pm) { g_pm = v; } ~PmScope() { g_pm = prev; } } pm_scope(cfg->precise != 0);
    for (int i = 0; i < n_tensors; ++i) {
      HostTensor t;
      t.data = tensors[i].data;
      t.shape.assign(tensors[i].shape, tensors[i].shape + tensors[i].ndim);
      n->W.host[tensors[i].name] = t;
    }
    build_encoder_w(*n);
    if (cfg->kind == WV_KIND_GENERATOR) {
      if (cfg->channels_dec % 32 != 0) WV_THROW(WV_ERR_UNSUPPORTED, "channels_dec must be a multiple of 32");
      build_decoder_w(*n);
    } else {
      build_head_w(*n);
    }
    n->W.host.clear();   // host pointers are only valid during this call
    CK(cudaDeviceSynchronize());
    *out = n.release();
  });
}

int wv_net_destroy(wv_net* net) {
  if (!net) return WV_OK;
  DeviceGuard dg(net->device);   // restores the caller's current device (this runs from Python finalisers)
  cudaDeviceSynchronize();
  delete net;
  return WV_OK;
}

int wv_net_reserve(wv_net* net, int B, int T) {
  if (!net || B < 1 || T < 1) return fail(WV_ERR_INVALID, "bad arguments to wv_net_reserve");
  return guarded([&] {
    DeviceGuard dg(net->device);
    const int bc = sub_batch(*net, B, T);
    // size for the largest sub-batch first so that the workspace never moves afterwards
    get_plan(*net, bc, T);
    if (B % bc) get_plan(*net, B % bc, T);
    get_plan(*net, bc, T);
  });
}

size_t wv_net_workspace_bytes(const wv_net* net) { return net ? net->ws_bytes : 0; }

int wv_net_launches(const wv_net* net, int B, int T) {
  if (!net) return 0;
  wv_net* n = const_cast<wv_net*>(net);
  const int bc = sub_batch(*n, B, T);
  int total = 0;
  for (int b0 = 0; b0 < B; b0 += bc) {
    auto it = n->plans.find(std::make_pair(std::min(bc, B - b0), T));
    if (it == n->plans.end()) return -1;
    total += static_cast<int>(it->second->ops.size());
  }
  return total;
}

int wv_net_set_chunk(wv_net* net, int max_clip_samples) {
  if (!net) return fail(WV_ERR_INVALID, "null net");
  net->chunk_samples = max_clip_samples;
  return WV_OK;
}

static uint32_t io_key(const IoPtrs& io) {
  const void* user[13] = {io.x, io.msg, io.presence, io.wm, io.y, io.latent, io.logits, io.bits, io.avg, io.conf, io.valid, io.mask, io.probs};
  uint32_t key = 0;
  for (int i = 0; i < 13; ++i) if (user[i]) key |= 1u << i;
  return key;
}

// Run `plan` through a CUDA graph (captured on first use for this set of requested outputs).
static void run_plan_graph(wv_net& n, Plan& plan, const IoPtrs& io, cudaStream_t st) {
  const int B = plan.B, T = plan.T, nb = n.cfg.nbits;
  const size_t BT = static_cast<size_t>(B) * T;
  // field order: x msg presence | wm y latent logits bits avg conf valid mask probs
  const void* user[13] = {io.x, io.msg, io.presence, io.wm, io.y, io.latent, io.logits, io.bits, io.avg, io.conf, io.valid, io.mask, io.probs};
  const size_t sizes[13] = {BT * 4, static_cast<size_t>(B) * n.cfg.msg_dimension * 4, BT, BT * 4, BT * 4,
                            static_cast<size_t>(B) * n.enc.dim * plan.F * 4, BT * nb * 4, static_cast<size_t>(B) * nb,
                            static_cast<size_t>(B) * nb * 4, static_cast<size_t>(B) * 4, static_cast<size_t>(B) * nb, BT, BT * 4};
  uint32_t key = 0;
  for (int i = 0; i < 13; ++i) if (user[i]) key |= 1u << i;
  PlanGraph& pg = plan.graphs[key];
  if (pg.exec == nullptr) {
    size_t total = 0, off[13];
    for (int i = 0; i < 13; ++i) { off[i] = total; if (user[i]) total += round_up(sizes[i], 256); }
    CK(cudaMalloc(reinterpret_cast<void**>(&pg.mem), std::max<size_t>(total, 256)));
    auto at = [&](int i) -> uint8_t* { return user[i] ? pg.mem + off[i] : nullptr; };
    pg.io.x = reinterpret_cast<const float*>(at(0)); pg.io.msg = reinterpret_cast<const float*>(at(1));
    pg.io.presence = at(2); pg.io.wm = reinterpret_cast<float*>(at(3)); pg.io.y = reinterpret_cast<float*>(at(4));
    pg.io.latent = reinterpret_cast<float*>(at(5)); pg.io.logits = reinterpret_cast<float*>(at(6)); pg.io.bits = at(7);
    pg.io.avg = reinterpret_cast<float*>(at(8)); pg.io.conf = reinterpret_cast<float*>(at(9)); pg.io.valid = at(10);
    pg.io.mask = at(11); pg.io.probs = reinterpret_cast<float*>(at(12));
    for (int i = 0; i < 13; ++i) pg.bytes[i] = user[i] ? sizes[i] : 0;
    if (!n.cap_stream) CK(cudaStreamCreateWithFlags(&n.cap_stream, cudaStreamNonBlocking));
    CK(cudaStreamBeginCapture(n.cap_stream, cudaStreamCaptureModeThreadLocal));
    try {
      run_plan(n, plan, pg.io, n.cap_stream);
    } catch (...) {
      cudaGraph_t dead = nullptr;
      cudaStreamEndCapture(n.cap_stream, &dead);
      if (dead) cudaGraphDestroy(dead);
      cudaFree(pg.mem);
      plan.graphs.erase(key);
      throw;
    }
    CK(cudaStreamEndCapture(n.cap_stream, &pg.graph));
    CK(cudaGraphInstantiate(&pg.exec, pg.graph, 0));
  }
  const void* internal[13] = {pg.io.x, pg.io.msg, pg.io.presence, pg.io.wm, pg.io.y, pg.io.latent, pg.io.logits, pg.io.bits,
                              pg.io.avg, pg.io.conf, pg.io.valid, pg.io.mask, pg.io.probs};
  for (int i = 0; i < 3; ++i)
    if (user[i]) CK(cudaMemcpyAsync(const_cast<void*>(internal[i]), user[i], pg.bytes[i], cudaMemcpyDeviceToDevice, st));
  CK(cudaGraphLaunch(pg.exec, st));
  for (int i = 3; i < 13; ++i)
    if (user[i]) CK(cudaMemcpyAsync(const_cast<void*>(user[i]), internal[i], pg.bytes[i], cudaMemcpyDeviceToDevice, st));
}

static int forward_common(wv_net* net, int B, int T, const IoPtrs& io0, cudaStream_t st) {
  return guarded([&] {
    DeviceGuard dg(net->device);
    if (B < 1 || T < 1) WV_THROW(WV_ERR_INVALID, "empty batch or clip (B=%d, T=%d)", B, T);
    const int bc = sub_batch(*net, B, T);
    get_plan(*net, bc, T);
    if (B % bc) get_plan(*net, B % bc, T);
    const int nb = net->cfg.nbits;
    const int F = ceil_div(T, net->enc.hop);
    for (int b0 = 0; b0 < B; b0 += bc) {
      const int bn = std::min(bc, B - b0);
      Plan& plan = get_plan(*net, bn, T);
      IoPtrs io = io0;
      const size_t so = static_cast<size_t>(b0) * T;
      if (io.x) io.x += so;
      if (io.msg) io.msg += static_cast<size_t>(b0) * net->cfg.msg_dimension;
      if (io.wm) io.wm += so;
      if (io.y) io.y += so;
      if (io.latent) io.latent += static_cast<size_t>(b0) * net->enc.dim * F;
      if (io.logits) io.logits += so * nb;
      if (io.bits) io.bits += static_cast<size_t>(b0) * nb;
      if (io.avg) io.avg += static_cast<size_t>(b0) * nb;
      if (io.conf) io.conf += b0;
      if (io.valid) io.valid += static_cast<size_t>(b0) * nb;
      if (io.presence) io.presence += so;
      if (io.mask) io.mask += so;
      if (io.probs) io.probs += so;
      // small (launch-bound) calls replay a CUDA graph; it is captured the second time a shape + output set is seen,
      // so one-off clip lengths (file API) do not pay capture + instantiation + staging memory for a single use
      bool graph = bn == B && !net->profile && !net->check_range && g_graph_max_samples > 0 && static_cast<long long>(B) * T <= g_graph_max_samples;
      if (graph) {
        const uint32_t key = io_key(io);
        graph = plan.graphs.count(key) > 0 || ++plan.graph_seen[key] >= 2;
      }
      if (graph) run_plan_graph(*net, plan, io, st);
      else run_plan(*net, plan, io, st);
    }
  });
}

int wv_generator_forward(wv_net* net, const float* x, const float* msg, int B, int T, float* wm_out,
                         float* y_out, float* latent_out, void* stream) {
  if (!net || net->cfg.kind != WV_KIND_GENERATOR) return fail(WV_ERR_INVALID, "not a generator net");
  if (!x || !msg) return fail(WV_ERR_INVALID, "x and msg are required");
  IoPtrs io;
  io.x = x; io.msg = msg; io.wm = wm_out; io.y = y_out; io.latent = latent_out;
  return forward_common(net, B, T, io, static_cast<cudaStream_t>(stream));
}

int wv_generator_encode(wv_net* net, const float* x, const float* msg, int B, int T, float* latent_out, void* stream) {
  // The decoder tail is cheap to skip only with a separate plan; encode() is an API-completeness
  // entry (model/generator.py:290), so it runs the full plan and discards the waveform.
  return wv_generator_forward(net, x, msg, B, T, nullptr, nullptr, latent_out, stream);
}

int wv_generator_decode(wv_net* net, const float* z, int B, int F, float* wav_out, void* stream) {
  if (!net || net->cfg.kind != WV_KIND_GENERATOR) return fail(WV_ERR_INVALID, "not a generator net");
  if (!z || !wav_out || B < 1 || F < 1) return fail(WV_ERR_INVALID, "bad arguments to wv_generator_decode");
  return guarded([&] {
    DeviceGuard dg(net->device);
    Plan& plan = get_dec_plan(*net, B, F);
    IoPtrs io;
    io.z_in = z; io.wm = wav_out;
    run_plan(*net, plan, io, static_cast<cudaStream_t>(stream));
  });
}

int wv_detector_forward(wv_net* net, const float* y, int B, int T, float* logits, uint8_t* bits, float* avg,
                        float* conf, uint8_t* valid, const uint8_t* presence, void* stream) {
  if (!net || net->cfg.kind != WV_KIND_DETECTOR) return fail(WV_ERR_INVALID, "not a detector net");
  if (!y) return fail(WV_ERR_INVALID, "y is required");
  if (conf && !avg) return fail(WV_ERR_INVALID, "conf requires avg");
  IoPtrs io;
  io.x = y; io.logits = logits; io.bits = bits; io.avg = avg; io.conf = conf; io.valid = valid; io.presence = presence;
  return forward_common(net, B, T, io, static_cast<cudaStream_t>(stream));
}

int wv_locator_forward(wv_net* net, const float* y, int B, int T, float* logits, uint8_t* mask, float* probs, void* stream) {
  if (!net || net->cfg.kind != WV_KIND_LOCATOR) return fail(WV_ERR_INVALID, "not a locator net");
  if (!y) return fail(WV_ERR_INVALID, "y is required");
  IoPtrs io;
  io.x = y; io.logits = logits; io.mask = mask; io.probs = probs;
  return forward_common(net, B, T, io, static_cast<cudaStream_t>(stream));
}


namespace {
RefineGraph& get_refine(wv_net& n, int B, int T, int K, bool want_logits, bool masked) {
  const std::vector<int> key = {B, T, K, want_logits ? 1 : 0, masked ? 1 : 0};
  get_plan(n, K, T);                       // may move the workspace, which drops every cached refine graph
  auto it = n.refine.find(key);
  if (it != n.refine.end()) return *it->second;
  if (n.refine.size() >= MAX_CACHED_PLANS) { CK(cudaDeviceSynchronize()); n.refine.clear(); }
  Plan& plan = get_plan(n, K, T);
  auto rg = std::make_unique<RefineGraph>();
  const int nb = n.cfg.nbits;
  const size_t KT = static_cast<size_t>(K) * T;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += round_up(std::max<size_t>(bytes, 16), 256); return o; };
  const size_t o_io = take(sizeof(RefineIo)), o_st = take(sizeof(RefineState)), o_list = take(static_cast<size_t>(B) * 4),
               o_neff = take(static_cast<size_t>(B) * 4), o_y = take(KT * 4), o_p = take(masked ? KT : 0),
               o_bits = take(static_cast<size_t>(K) * nb), o_avg = take(static_cast<size_t>(K) * nb * 4), o_conf = take(static_cast<size_t>(K) * 4),
               o_valid = take(static_cast<size_t>(K) * nb), o_lg = take(want_logits ? KT * nb * 4 : 0);
  CK(cudaMalloc(reinterpret_cast<void**>(&rg->mem), off));
  CK(cudaMemset(rg->mem, 0, off));
  uint8_t* m = rg->mem;
  rg->io = reinterpret_cast<RefineIo*>(m + o_io);
  RefineIo* io_d = rg->io;
  RefineState* st_d = reinterpret_cast<RefineState*>(m + o_st);
  int* list_d = reinterpret_cast<int*>(m + o_list);
  int* neff_d = masked ? reinterpret_cast<int*>(m + o_neff) : nullptr;
  float* ysub = reinterpret_cast<float*>(m + o_y);
  uint8_t* psub = masked ? m + o_p : nullptr;
  IoPtrs sub;
  sub.x = ysub; sub.presence = psub; sub.bits = m + o_bits; sub.avg = reinterpret_cast<float*>(m + o_avg);
  sub.conf = reinterpret_cast<float*>(m + o_conf); sub.valid = m + o_valid;
  sub.logits = want_logits ? reinterpret_cast<float*>(m + o_lg) : nullptr;

  CK(cudaGraphCreate(&rg->graph, 0));
  cudaGraphConditionalHandle loop;
  CK(cudaGraphConditionalHandleCreate(&loop, rg->graph, 0, cudaGraphCondAssignDefault));
  std::vector<cudaGraphNode_t> deps;
  int Bv = B, Tv = T, nbv = nb;
  if (masked) {
    cudaKernelNodeParams kp = {};
    void* args[3] = {&io_d, &Tv, &neff_d};
    kp.func = reinterpret_cast<void*>(refine_neff_kernel); kp.gridDim = dim3(B); kp.blockDim = dim3(256); kp.kernelParams = args;
    cudaGraphNode_t node;
    CK(cudaGraphAddKernelNode(&node, rg->graph, nullptr, 0, &kp));
    deps.push_back(node);
  }
  {
    cudaKernelNodeParams kp = {};
    void* args[8] = {&loop, &io_d, &Bv, &nbv, &Tv, &neff_d, &list_d, &st_d};
    kp.func = reinterpret_cast<void*>(refine_select_kernel); kp.gridDim = dim3(1); kp.blockDim = dim3(32); kp.kernelParams = args;
    cudaGraphNode_t node;
    CK(cudaGraphAddKernelNode(&node, rg->graph, deps.data(), deps.size(), &kp));
    deps.assign(1, node);
  }
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = loop;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t cond;
  CK(cudaGraphAddNode(&cond, rg->graph, deps.data(), deps.size(), &np));
  cudaGraph_t body = np.conditional.phGraph_out[0];
  if (!n.cap_stream) CK(cudaStreamCreateWithFlags(&n.cap_stream, cudaStreamNonBlocking));
  const bool prof = n.profile;
  n.profile = false;
  CK(cudaStreamBeginCaptureToGraph(n.cap_stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal));
  try {
    const dim3 gg(std::max(1, std::min(64, ceil_div(T, 1024))), K);
    refine_gather_kernel<<<gg, 256, 0, n.cap_stream>>>(io_d, list_d, st_d, K, T, ysub, psub);
    run_plan(n, plan, sub, n.cap_stream);
    const dim3 gs(want_logits ? std::max(1, std::min(128, ceil_div(T * nb, 2048))) : 1, K);
    refine_scatter_kernel<<<gs, 256, 0, n.cap_stream>>>(io_d, list_d, st_d, K, T, nb, sub.bits, sub.avg, sub.conf, sub.valid, sub.logits);
    refine_advance_kernel<<<1, 1, 0, n.cap_stream>>>(loop, io_d, st_d, K);
    CK(cudaGetLastError());
  } catch (...) {
    cudaGraph_t dead = nullptr;
    cudaStreamEndCapture(n.cap_stream, &dead);
    n.profile = prof;
    throw;
  }
  n.profile = prof;
  cudaGraph_t same = nullptr;
  CK(cudaStreamEndCapture(n.cap_stream, &same));
  CK(cudaGraphInstantiate(&rg->exec, rg->graph, 0));
  RefineGraph& ref = *rg;
  n.refine[key] = std::move(rg);
  return ref;
}
}  // namespace

int wv_detector_refine(wv_net* net, const float* y, int B, int T, float* logits, uint8_t* bits, float* avg, float* conf,
                       uint8_t* valid, const uint8_t* presence, float tau, float tau_short, int short_samples, int slots,
                       int* counters, void* stream) {
  if (!net || net->cfg.kind != WV_KIND_DETECTOR || !net->cfg.precise) return fail(WV_ERR_INVALID, "refine needs a precise detector net");
  if (!y || !bits || !avg) return fail(WV_ERR_INVALID, "y, bits and avg are required");
  if (presence && !valid) return fail(WV_ERR_INVALID, "masked refine needs valid");
  if (B < 1 || T < 1 || slots < 1) return fail(WV_ERR_INVALID, "bad refine shape");
  return guarded([&] {
    DeviceGuard dg(net->device);
    int K = std::min(slots, B);
    if (net->chunk_samples > 0) K = static_cast<int>(std::max<long long>(1, std::min<long long>(K, net->chunk_samples / std::max(1, T))));
    RefineGraph& rg = get_refine(*net, B, T, K, logits != nullptr, presence != nullptr);
    RefineIo v;
    v.y = y; v.presence = presence; v.logits = logits; v.bits = bits; v.avg = avg; v.conf = conf; v.valid = valid;
    v.counters = counters; v.tau = tau; v.tau_short = tau_short; v.short_samples = short_samples;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    refine_params_kernel<<<1, 1, 0, st>>>(rg.io, v);
    CK(cudaGetLastError());
    CK(cudaGraphLaunch(rg.exec, st));
  });
}

int wv_metrics_accumulate(const uint8_t* bits, const uint8_t* valid, const uint8_t* msg_bits, int B, int nbits,
                          const uint8_t* pred_mask, const uint8_t* gt_mask, long long n_mask, long long* counters, void* stream) {
  if (!counters) return fail(WV_ERR_INVALID, "counters is required");
  if (bits && !msg_bits) return fail(WV_ERR_INVALID, "msg_bits is required with bits");
  if (pred_mask && !gt_mask) return fail(WV_ERR_INVALID, "gt_mask is required with pred_mask");
  return guarded([&] {
    init_device_once();
    const long long nb = static_cast<long long>(B) * nbits;
    const long long work = std::max(bits ? nb : 0, pred_mask ? n_mask : 0);
    if (work <= 0) return;
    metrics_kernel<<<elem_grid(work), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        bits, valid, msg_bits, nb, pred_mask, gt_mask, n_mask, reinterpret_cast<unsigned long long*>(counters));
    CK(cudaGetLastError());
  });
}

// ---- validation path: temporal augmentations + cheap effects (validation_kernels.cuh) ---------
int wv_augment_gather(const float* original, const float* watermarked, const float* gt_in, int B, int T,
                      const uint8_t* seg_op, const int* seg_src, int seg_len, int n_seg, int seq_kind, int seq_a,
                      int seq_b, int seq_c, const int* seq_perm, int n_perm, int T_out, float* out_wm,
                      float* out_orig, float* out_gt, void* stream) {
  if (!watermarked) return fail(WV_ERR_INVALID, "watermarked is required");
  if (B < 0 || T <= 0) return failf(WV_ERR_INVALID, "bad shape B=%d T=%d", B, T);
  if ((seg_op != nullptr || out_orig != nullptr) && !original) return fail(WV_ERR_INVALID, "original is required");
  if (seg_op != nullptr && (seg_len <= 0 || static_cast<long long>(n_seg) * seg_len < T))
    return failf(WV_ERR_INVALID, "segment table [%d x %d] does not cover T=%d", n_seg, seg_len, T);
  if (seg_op != nullptr && !seg_src) return fail(WV_ERR_INVALID, "seg_src is required with seg_op");
  if (seq_kind < 0 || seq_kind > 4) return failf(WV_ERR_INVALID, "unknown sequence map %d", seq_kind);
  int want_T = T;
  if (seq_kind == 2 && (seq_a < 0 || seq_a >= T)) return failf(WV_ERR_INVALID, "shift %d outside [0, %d)", seq_a, T);
  if (seq_kind == 3) {
    if (!seq_perm || seq_c <= 0 || n_perm <= 0 || static_cast<long long>(n_perm) * seq_c > T)
      return failf(WV_ERR_INVALID, "bad shuffle: %d segments of %d samples for T=%d", n_perm, seq_c, T);
    want_T = n_perm * seq_c;
  }
  if (seq_kind == 4) {
    const int lo = std::min(seq_a, seq_b), hi = std::max(seq_a, seq_b);
    if (seq_c <= 0 || lo < 0 || hi + seq_c > T || hi - lo < seq_c)
      return failf(WV_ERR_INVALID, "bad chunk swap [%d,+%d) <-> [%d,+%d) for T=%d", seq_a, seq_c, seq_b, seq_c, T);
  }
  if (T_out != want_T) return failf(WV_ERR_INVALID, "T_out=%d, expected %d", T_out, want_T);
  return guarded([&] {
    init_device_once();
    if (B == 0) return;
    SeqMap m{seq_kind, seq_a, seq_b, seq_c, seq_perm, n_perm};
    const long long aug_blocks = static_cast<long long>(B) * ceil_div(T_out, AUG_TILE);
    augment_gather_kernel<<<static_cast<int>(std::min<long long>(aug_blocks, static_cast<long long>(g_num_sms) * 32)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        original, watermarked, gt_in, B, T, seg_op, seg_src, seg_len, n_seg, m, T_out, out_wm, out_orig, out_gt);
    CK(cudaGetLastError());
  });
}

int wv_effect_pointwise(int effect, const float* in, long long n, float p0, const float* noise,
                        unsigned long long seed, float* out, void* stream) {
  if (effect < 0 || effect > 4) return failf(WV_ERR_INVALID, "unknown pointwise effect %d", effect);
  if (!in || !out || n < 0) return fail(WV_ERR_INVALID, "in / out are required");
  if (effect == 3 && !noise) return fail(WV_ERR_INVALID, "effect 3 needs the noise draw");
  if (effect == 2 && !(p0 >= 1.f)) return failf(WV_ERR_INVALID, "quantization scale %g < 1", p0);
  if ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(noise)) & 15)
    return fail(WV_ERR_INVALID, "buffers must be 16-byte aligned");
  return guarded([&] {
    init_device_once();
    if (n == 0) return;
    effect_pointwise_kernel<<<elem_grid((n + 3) / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(effect, in, n, p0, noise, seed, out);
    CK(cudaGetLastError());
  });
}

int wv_effect_suppress(float* audio, float* mask, const long long* idx, int B, int T, int k, void* stream) {
  if (!audio || (k > 0 && !idx) || B < 0 || T <= 0 || k < 0) return fail(WV_ERR_INVALID, "bad suppress arguments");
  return guarded([&] {
    init_device_once();
    if (B == 0 || k == 0) return;
    effect_suppress_kernel<<<elem_grid(static_cast<long long>(B) * k), 256, 0, static_cast<cudaStream_t>(stream)>>>(audio, mask, idx, B, T, k);
    CK(cudaGetLastError());
  });
}

int wv_effect_median(const float* in, int B, int T, int k, float* out, void* stream) {
  if (!in || !out || B < 0 || T <= 0) return fail(WV_ERR_INVALID, "bad median arguments");
  if (k < 1 || k > MEDIAN_MAX_K || k % 2 == 0) return failf(WV_ERR_INVALID, "median window %d: odd, 1..%d", k, MEDIAN_MAX_K);
  return guarded([&] {
    init_device_once();
    if (B == 0) return;
    const long long tiles = static_cast<long long>(B) * ceil_div(T, MEDIAN_TILE);
    const int grid = static_cast<int>(std::min<long long>(tiles, static_cast<long long>(g_num_sms) * 16));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    switch (k) {
#define WV_MED(K) case K: effect_median_kernel<K><<<grid, MEDIAN_THREADS, 0, st>>>(in, B, T, out); break;
      WV_MED(1) WV_MED(3) WV_MED(5) WV_MED(7) WV_MED(9) WV_MED(11) WV_MED(13) WV_MED(15) WV_MED(17) WV_MED(19)
      WV_MED(21) WV_MED(23) WV_MED(25) WV_MED(27) WV_MED(29) WV_MED(31)
#undef WV_MED
    }
    CK(cudaGetLastError());
  });
}

int wv_effect_fir(const float* in, const float* taps, int n_taps, int B, int T, int subtract, float* out, void* stream) {
  if (!in || !out || !taps || B < 0 || T <= 0) return fail(WV_ERR_INVALID, "bad FIR arguments");
  if (n_taps < 1 || n_taps > FIR_MAX_TAPS || n_taps % 2 == 0) return failf(WV_ERR_INVALID, "FIR length %d: odd, 1..%d", n_taps, FIR_MAX_TAPS);
  return guarded([&] {
    init_device_once();
    if (B == 0) return;
    const size_t smem = static_cast<size_t>(2 * ((n_taps + 3) & ~3) + FIR_TILE) * sizeof(float);
    static bool attr = false;
    if (!attr) {
      CK(cudaFuncSetAttribute(effect_fir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              static_cast<int>((2 * ((FIR_MAX_TAPS + 3) & ~3) + FIR_TILE) * sizeof(float))));
      attr = true;
    }
    const long long tiles = static_cast<long long>(B) * ceil_div(T, FIR_TILE);
    const int grid = static_cast<int>(std::min<long long>(tiles, static_cast<long long>(g_num_sms) * 8));
    effect_fir_kernel<<<grid, FIR_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(in, taps, n_taps, B, T, subtract, out);
    CK(cudaGetLastError());
  });
}

int wv_effect_resample(const float* in, const float* taps, int B, int T, int orig, int nw, int width, int T_mid, int T_out,
                       int lerp, float* out, void* stream) {
  if (!in || !out || !taps || B < 0 || T <= 0) return fail(WV_ERR_INVALID, "bad resample arguments");
  if (orig < 1 || nw < 1 || width < 0 || T_mid < 1 || T_out < 1) return fail(WV_ERR_INVALID, "bad resample geometry");
  if (!lerp && T_out != T_mid) return fail(WV_ERR_INVALID, "T_out must equal T_mid without the interpolation step");
  if (static_cast<long long>(T_mid) > (static_cast<long long>(T) * nw + orig - 1) / orig)
    return failf(WV_ERR_INVALID, "T_mid=%d exceeds ceil(T*new/orig)", T_mid);
  return guarded([&] {
    init_device_once();
    if (B == 0) return;
    effect_resample_kernel<<<elem_grid(static_cast<long long>(B) * T_out), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        in, taps, B, T, orig, nw, width, 2 * width + orig, T_mid, T_out, lerp, out);
    CK(cudaGetLastError());
  });
}

// ---- fp16 range check ------------------------------------------------------------------------
int wv_net_set_range_check(wv_net* net, int enable) {
  if (!net) return fail(WV_ERR_INVALID, "null net");
  return guarded([&] {
    DeviceGuard dg(net->device);
    if (enable && !net->range_stats) {
      CK(cudaMalloc(reinterpret_cast<void**>(&net->range_stats), 4 * sizeof(unsigned long long)));
      CK(cudaMemset(net->range_stats, 0, 4 * sizeof(unsigned long long)));
    }
    net->check_range = enable != 0;
  });
}

int wv_net_range_read(wv_net* net, unsigned long long* saturated, unsigned long long* nonfinite, float* max_abs) {
  if (!net || !net->range_stats) return fail(WV_ERR_INVALID, "range check was never enabled on this net");
  return guarded([&] {
    DeviceGuard dg(net->device);
    unsigned long long h[4];
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, net->range_stats, sizeof(h), cudaMemcpyDeviceToHost));
    CK(cudaMemset(net->range_stats, 0, sizeof(h)));
    if (saturated) *saturated = h[0];
    if (nonfinite) *nonfinite = h[1];
    if (max_abs) {
      const uint16_t bits = static_cast<uint16_t>(h[2]);
      __half hv;
      memcpy(&hv, &bits, 2);
      *max_abs = __half2float(hv);
    }
  });
}

// ---- profiling / debugging -----------------------------------------------------------------
int wv_net_set_profile(wv_net* net, int enable) {
  if (!net) return fail(WV_ERR_INVALID, "null net");
  net->profile = enable != 0;
  return WV_OK;
}

// After a profiled forward (single sub-batch): per-op device time (ms), algorithmic flops/bytes,
// op class (OpType, GEMMs: 100 + epilogue id).  Returns the number of ops, or a negative code.
int wv_net_profile_read(wv_net* net, int max_ops, float* ms, double* flops, double* bytes, int* cls) {
  if (!net || !net->last_plan) return fail(WV_ERR_INVALID, "no profiled run");
  int count = 0;
  int rc = guarded([&] {
    const Plan& p = *net->last_plan;
    CK(cudaEventSynchronize(p.events[p.ops.size()]));
    count = static_cast<int>(std::min<size_t>(p.ops.size(), static_cast<size_t>(max_ops)));
    for (int i = 0; i < count; ++i) {
      float t = 0.f;
      CK(cudaEventElapsedTime(&t, p.events[i], p.events[i + 1]));
      ms[i] = t;
      flops[i] = p.ops[i].flops;
      bytes[i] = p.ops[i].bytes;
      cls[i] = p.ops[i].type == OP_GEMM ? 100 + p.ops[i].epi : static_cast<int>(p.ops[i].type);   // OP_RESBLOCK = 20
    }
  });
  return rc == WV_OK ? count : rc;
}

const char* wv_net_profile_tag(wv_net* net, int i) {
  if (!net || !net->last_plan || i < 0 || i >= static_cast<int>(net->last_plan->ops.size())) return "";
  return net->last_plan->ops[i].tag.c_str();
}

// Run the plan up to (and including) the op tagged `tag` and copy its output buffer `which`
// (0 = raw, 1 = activated) to dst (device or host pointer).  Test / debug only.
int wv_debug_tap(wv_net* net, const float* x, const float* msg, int B, int T, const char* tag, int which,
                 void* dst, size_t dst_bytes, size_t* written) {
  if (!net || !x || !tag || !dst) return fail(WV_ERR_INVALID, "bad arguments to wv_debug_tap");
  return guarded([&] {
    DeviceGuard dg(net->device);
    const long long saved = net->chunk_samples;
    net->chunk_samples = 0;
    Plan& plan = get_plan(*net, B, T);
    net->chunk_samples = saved;
    int idx = -1;
    for (size_t i = 0; i < plan.ops.size(); ++i)
      if (plan.ops[i].tag == tag) idx = static_cast<int>(i);
    if (idx < 0) WV_THROW(WV_ERR_INVALID, "no op tagged '%s'", tag);
    const Op& op = plan.ops[idx];
    const void* src = which ? op.out1 : op.out0;
    const size_t nbytes = std::min(dst_bytes, op.out_bytes[which ? 1 : 0]);
    if (!src || nbytes == 0) WV_THROW(WV_ERR_INVALID, "op '%s' has no output %d", tag, which);
    IoPtrs io;
    io.x = x; io.msg = msg;
    run_plan(*net, plan, io, nullptr, idx);
    CK(cudaStreamSynchronize(nullptr));
    CK(cudaMemcpy(dst, src, nbytes, cudaMemcpyDefault));
    if (written) *written = nbytes;
  });
}

// ---- single-kernel entry points --------------------------------------------------------
int wv_op_gemm(const void* A, int lda, const void* Wt, int ldw, int M, int N, int K, const float* bias,
               const void* residual, void* out_raw, void* out_act, float act_scale, int a_is_fp16, void* stream) {
  return guarded([&] {
    init_device_once();
    GemmW w;
    w.w = const_cast<void*>(Wt); w.N = N; w.K = K; w.ldw = ldw; w.fp16 = true; (void)a_is_fp16;
    w.block_n = pick_block_n(N, 0, STAGED_MAX_BN);
    w.tm = make_tmap(Wt, 2, K, N, 1, ldw, 0, BK, w.block_n, w.fp16);
    std::vector<Op> ops;
    PlanCtx c;
    c.base = reinterpret_cast<uint8_t*>(16);   // non-null: encode tensor maps
    c.ops = &ops;
    add_gemm(c, EPI_STAGED, w, A, lda, M, K,
             std_args(bias, static_cast<const h16*>(residual), static_cast<h16*>(out_raw), static_cast<h16*>(out_act), act_scale, N));
    launch_gemm(ops[0], ops[0].g, static_cast<cudaStream_t>(stream));
    CK(cudaGetLastError());
  });
}

int wv_op_gemm_dw5(const void* A, const void* Wt, int B, int T, int N, int K, const float* dw_w5n, const float* bias,
                   const void* residual, void* out_raw, void* out_act, float act_scale, void* stream) {
  return guarded([&] {
    init_device_once();
    GemmW w;
    w.w = const_cast<void*>(Wt); w.N = N; w.K = K; w.ldw = K; w.fp16 = true;
    w.block_n = pick_block_n(N, 0, STAGED_MAX_BN);
    w.tm = make_tmap(Wt, 2, K, N, 1, K, 0, BK, w.block_n, false);
    DwW dw;
    dw.w = const_cast<float*>(dw_w5n); dw.bias = const_cast<float*>(bias); dw.k = 5; dw.C = N;
    std::vector<Op> ops;
    PlanCtx c;
    c.base = reinterpret_cast<uint8_t*>(16);
    c.ops = &ops;
    c.B = B;
    add_gemm_dw(c, w, dw, static_cast<const h16*>(A), T, K, static_cast<const h16*>(residual),
                static_cast<h16*>(out_raw), static_cast<h16*>(out_act), act_scale);
    if (getenv("WV_TIMELINE") && atoi(getenv("WV_TIMELINE")) >= 10) {   // ablation timing: mode = value - 10, no probes
      ops[0].g.dbg_mode = atoi(getenv("WV_TIMELINE")) - 10;
      launch_gemm(ops[0], ops[0].g, static_cast<cudaStream_t>(stream));
      CK(cudaGetLastError());
      return;
    }
    if (getenv("WV_TIMELINE")) {   // per-tile clock probes of CTA 0 (scripts/timeline.py)
      long long* d = nullptr;
      CK(cudaMalloc(&d, 48 * 40 * sizeof(long long)));
      CK(cudaMemset(d, 0, 48 * 40 * sizeof(long long)));
      ops[0].g.dbg = d;
      if (getenv("WV_TIMELINE_MODE")) ops[0].g.dbg_mode = atoi(getenv("WV_TIMELINE_MODE"));   // probes + ablation bits
      launch_gemm(ops[0], ops[0].g, static_cast<cudaStream_t>(stream));
      CK(cudaDeviceSynchronize());
      std::vector<long long> h(48 * 40);
      CK(cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      CK(cudaFree(d));
      const long long t0 = h[0];
      printf("tile: tma_issue  data_in  mma_commit  drain_start drain_end  math_start math_end   (cycles from first TMA)\n");
      for (int i = 0; i < 40; ++i) {
        printf("%3d:", i);
        for (int k = 0; k < 8; ++k) printf(" %9lld", h[i * 40 + k] ? h[i * 40 + k] - t0 : -1);   // column 8 = warp 0 at the hand-off barrier
        printf(" | drain start");
        for (int k = 36; k < 40; ++k) printf(" %6lld", h[i * 40 + k] - t0);
        printf(" end");
        for (int k = 8; k < 12; ++k) printf(" %6lld", h[i * 40 + k] - t0);
        printf(" | math start");
        for (int k = 24; k < 36; ++k) printf(" %6lld", h[i * 40 + k] - t0);
        printf(" end");
        for (int k = 12; k < 24; ++k) printf(" %6lld", h[i * 40 + k] - t0);
        printf("\n");
      }
      return;
    }
    launch_gemm(ops[0], ops[0].g, static_cast<cudaStream_t>(stream));
    CK(cudaGetLastError());
  });
}

int wv_op_resblock(const void* X, const void* A, const void* W1, const float* dw1_w5c, const float* dw1_b, const void* W2,
                   const float* dw2_w5c, const float* dw2_b, int B, int T, int C, float pre_scale, void* out_raw,
                   void* out_act, float act_scale, void* stream) {
  if (!X || !A) return fail(WV_ERR_INVALID, "X (raw) and A = ELU(X * pre_scale) are required");
  return guarded([&] {
    init_device_once();
    ResW r;
    for (GemmW* w : {&r.pw1, &r.pw2}) {
      w->N = C; w->K = C; w->ldw = C; w->fp16 = true; w->block_n = C;
    }
    r.pw1.w = const_cast<void*>(W1); r.pw2.w = const_cast<void*>(W2);
    r.pw1.tm = make_tmap(W1, 2, C, C, 1, C, 0, BK, C, true);
    r.pw2.tm = make_tmap(W2, 2, C, C, 1, C, 0, BK, C, true);
    r.dw1.w = const_cast<float*>(dw1_w5c); r.dw1.bias = const_cast<float*>(dw1_b); r.dw1.k = 5; r.dw1.C = C;
    r.dw2.w = const_cast<float*>(dw2_w5c); r.dw2.bias = const_cast<float*>(dw2_b); r.dw2.k = 5; r.dw2.C = C;
    r.pre_scale = pre_scale;
    if (C > 128 || C % 32 != 0 || rb_pick_nx(C, ceil_div(C, BK)) < 2)
      WV_THROW(WV_ERR_UNSUPPORTED, "fused resblock needs C %% 32 == 0 and C <= 128 (got %d)", C);
    std::vector<Op> ops;
    PlanCtx c;
    c.base = reinterpret_cast<uint8_t*>(16);
    c.ops = &ops;
    c.B = B;
    add_resblock_fused(c, r, static_cast<const h16*>(X), static_cast<const h16*>(A), T, C, static_cast<h16*>(out_raw), static_cast<h16*>(out_act), act_scale, "op");
    long long* dbg = nullptr;
    if (getenv("WV_TIMELINE_RB")) {
      CK(cudaMalloc(&dbg, 24 * 32 * sizeof(long long)));
      CK(cudaMemset(dbg, 0, 24 * 32 * sizeof(long long)));
      ops[0].rb.dbg = dbg;
    }
    launch_k(resblock_sm100_kernel, ops[0].grid, RB_THREADS, static_cast<size_t>(ops[0].i[7]), static_cast<cudaStream_t>(stream),
             ops[0].tmA, ops[0].tmB, ops[0].tmR, ops[0].rb);
    CK(cudaGetLastError());
    if (dbg) {
      CK(cudaDeviceSynchronize());
      std::vector<long long> h(24 * 32);
      CK(cudaMemcpy(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
      CK(cudaFree(dbg));
      long long t0 = h[12];
      printf("pair: mma1a mma1b mma2a mma2b | dr1 start a b end a b | dr2 start a b end a b | T0 wait a b  got a b  done a b | M1 start a b done a b | M2 start a b done a b\n");
      for (int p = 0; p < 14; ++p) {
        printf("%2d:", p);
        for (int k = 0; k < 26; ++k) printf(" %6lld", h[p * 32 + k] ? h[p * 32 + k] - t0 : -1);
        printf("\n");
      }
    }
  });
}

int wv_op_dw5(const void* in, const float* w5c, const float* bias, const void* residual, void* out_raw, void* out_act,
              float act_scale, int B, int T, int C, void* stream) {
  return guarded([&] {
    init_device_once();
    dw5_kernel<<<elem_grid(static_cast<long long>(B) * ceil_div(T, DW_TT) * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const h16*>(in), w5c, bias, static_cast<const h16*>(residual), static_cast<h16*>(out_raw),
        static_cast<h16*>(out_act), act_scale, B, T, C);
    CK(cudaGetLastError());
  });
}

int wv_op_down(const void* in, const float* wkc, const float* bias, const float* film, int film_stride, int bands,
               void* out_raw, void* out_act, float act_scale, int B, int Tin, int C, int r, void* stream) {
  return guarded([&] {
    init_device_once();
    const int To = ceil_div(Tin, r);
    launch_down(elem_grid(static_cast<long long>(B) * To * (C / 8)), static_cast<cudaStream_t>(stream),
                static_cast<const h16*>(in), wkc, bias, film, film_stride, bands, static_cast<h16*>(out_raw),
                static_cast<h16*>(out_act), act_scale, B, Tin, To, C, r);
    CK(cudaGetLastError());
  });
}

int wv_op_up(const void* in, const float* wkc, void* out, int B, int Tin, int C, int r, void* stream) {
  return guarded([&] {
    init_device_once();
    up_kernel<<<elem_grid(static_cast<long long>(B) * Tin * (C / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const h16*>(in), wkc, static_cast<h16*>(out), B, Tin, C, r);
    CK(cudaGetLastError());
  });
}

}  // extern "C"
