// Persistent warp-specialised bf16 GEMM for sm_100a:
//   TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory ring
//   -> tcgen05.mma (single issuing thread, fp32 accumulators in TMEM, double buffered)
//   -> tcgen05.ld epilogue (bias / q-scale / QuickGELU / dQuickGELU / residual) -> global.
// One CTA per SM, 12 warps: warp0 = TMA producer, warp1 = MMA issuer, warp2 = TMEM allocator,
// warps 4..11 = epilogue (two warps per TMEM lane quarter, each taking half of the columns).
//
// Replaces (see include/xpretrain_b200.h) every nn.Linear forward/backward on
// the CLIP-ViP hot path: CLIP_ViP.py:341-343,379,393-395,1141-1145 and the
// patch-embedding conv :178 (as an im2col GEMM).
#include "../../include/xpretrain_b200.h"
#include "common.h"
#include "ptx.cuh"
#include <cstdlib>

namespace xp {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 384;  // 4 control warps + 8 epilogue warps
constexpr int EPI_THREADS = 256;

struct GemmDev {
  void* c;
  const float* bias;
  const __nv_bfloat16* residual;
  __nv_bfloat16* aux;
  int M, N, K;
  long long ldc, ldr, ld_aux;
  long long c_group, c_group_stride, r_group, r_group_stride;
  int act;
  int splits;
  int scale_cols;
  float alpha, col_scale;
  int wide;                 // C / aux / residual rows are 32-byte aligned (ld % 16 == 0)
  uint32_t mn_lbo, mn_sbo;  // MN-major descriptor strides (bytes): 64-element atom stride, 8-k-row group stride
  int dbg;                  // profiling only (XP_GEMM_DEBUG): 1 no stores at all, 2 no wait for the staging boxes (racy), 4 stage but
                            // do not issue the TMA stores, 8 activation = identity, 16 no epilogue input loads, 32 no aux store
  int tma_c, tma_aux;       // pair kernel: C / the aux output leave through shared-memory staging + TMA stores
};

template <int BN>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int TMEM_COLS = 2 * BN;
  // ring + 1 KiB alignment slack + barriers
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256 + 2 * BN * 4;
};

__device__ __forceinline__ float act_gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float act_gelu_erf_grad(float x) {
  float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
  float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// 16 consecutive bf16 (32 bytes) of one row: one 256-bit access when the row segment is 32-byte aligned,
// else two 128-bit accesses (the second only if those 8 columns exist).
__device__ __forceinline__ void ld_bf16x16(const __nv_bfloat16* ptr, bool wide, bool second, uint32_t (&w)[8]) {
  if (wide) {
    asm volatile("ld.global.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(ptr));
  } else {
    const uint4 a = *reinterpret_cast<const uint4*>(ptr);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    w[4] = w[5] = w[6] = w[7] = 0u;
    if (second) {
      const uint4 b = *reinterpret_cast<const uint4*>(ptr + 8);
      w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    }
  }
}
__device__ __forceinline__ void st_bf16x16(__nv_bfloat16* ptr, bool wide, bool second, const float (&v)[16]) {
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = pack_bf16(v[2 * i], v[2 * i + 1]);
  if (wide) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                 "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                 : "memory");
  } else {
    *reinterpret_cast<uint4*>(ptr) = make_uint4(w[0], w[1], w[2], w[3]);
    if (second) *reinterpret_cast<uint4*>(ptr + 8) = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

template <int BN, int A_MN, int B_MN, int OUT, int ACT>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmDev p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_smem = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + 256);  // [2][BN]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (p.M + BM - 1) / BM;
  const int num_n = (p.N + BN - 1) / BN;
  const int num_mn = num_m * num_n;
  const int total = num_mn * p.splits;
  const int kb_total = (p.K + BK - 1) / BK;
  const int kb_per = (kb_total + p.splits - 1) / p.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int split = tile / num_mn;
        const int mn = tile - split * num_mn;
        const int m_blk = mn / num_n;
        const int n_blk = mn - m_blk * num_n;
        const int k0 = split * kb_per;
        const int k1 = min(kb_total, k0 + kb_per);
        for (int kb = k0; kb < k1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + Cfg::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              tma_load_2d(sA + i * (BK * 128), &tmA, &full_bar[stage], m_blk * BM + i * 64, kb * BK);
          } else {
            tma_load_2d(sA, &tmA, &full_bar[stage], kb * BK, m_blk * BM);
          }
          if (B_MN) {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)
              tma_load_2d(sB + i * (BK * 128), &tmB, &full_bar[stage], n_blk * BN + i * 64, kb * BK);
          } else {
            tma_load_2d(sB, &tmB, &full_bar[stage], kb * BK, n_blk * BN);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // -------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN, B_MN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int split = tile / num_mn;
        const int k0 = split * kb_per;
        const int k1 = min(kb_total, k0 + kb_per);
        if (k0 >= k1) continue;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = k0; kb < k1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sA = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t sB = sA + Cfg::A_BYTES;
#pragma unroll
          for (int j = 0; j < BK / UMMA_K; ++j) {
            // K-major: 16 elements = 32 B further along the swizzled row; 8-row groups 1024 B apart.
            // MN-major: 16 k-rows = 2048 B further; 64-element MN atoms BK*128 B apart.
            const uint64_t adesc = A_MN ? make_smem_desc_sw128(sA + j * (UMMA_K * 128), p.mn_lbo, p.mn_sbo)
                                        : make_smem_desc_sw128(sA + j * (UMMA_K * 2), 16, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(sB + j * (UMMA_K * 128), p.mn_lbo, p.mn_sbo)
                                        : make_smem_desc_sw128(sB + j * (UMMA_K * 2), 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (kb > k0 || j > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tmem_full[acc]);  // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------- epilogue
    // 8 warps: warp (4+e) owns TMEM lane quarter e&3 and column half e>>2 of every accumulator.
    const int ew = warp - 4;
    const int quarter = ew & 3;  // == warp % 4, the lane quarter this warp may read
    const int half = ew >> 2;
    const int etid = threadIdx.x - 128;  // 0..255
    constexpr int CHUNKS = BN / 64;      // 32-column chunks per warp per tile
    // kernel parameters used per element live in registers (the asm "memory" clobbers would otherwise force
    // ptxas to re-read them from the constant bank inside the loops)
    const bool wide = p.wide != 0;       // every bf16 row segment is 32-byte aligned -> 256-bit accesses
    constexpr int act = ACT;             // compile-time: the unused activation branches are not even generated
    const int N = p.N, scale_cols = p.scale_cols;
    const float alpha = p.alpha, col_scale = p.col_scale;
    __nv_bfloat16* const aux_p = p.aux;
    const bool need_aux_in = (act == XP_ACT_DQUICK_GELU || act == XP_ACT_DGELU_ERF);
    // the one extra bf16 INPUT the epilogue streams: the saved pre-activation (dGELU) or the residual
    const __nv_bfloat16* const xin_p = need_aux_in ? p.aux : p.residual;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
      const int split = tile / num_mn;
      const int mn = tile - split * num_mn;
      const int m_blk = mn / num_n;
      const int n_blk = mn - m_blk * num_n;
      const int k0 = split * kb_per;
      const int k1 = min(kb_total, k0 + kb_per);
      if (k0 >= k1) continue;
      // stage this tile's bias slice in shared memory (double buffered by accumulator stage)
      float* sbias = bias_smem + acc * BN;
      for (int i = etid; i < BN; i += EPI_THREADS) {
        const int n = n_blk * BN + i;
        sbias[i] = (p.bias != nullptr && n < N && split == 0) ? p.bias[n] : 0.f;
      }
      const int row = m_blk * BM + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      constexpr bool kTmaEpi = false;                    // TMA-store epilogue: 2-CTA kernel only (names below are unused here)
      const uint32_t stg = 0;
      const CUtensorMap& tmC = tmA;
      const CUtensorMap& tmX = tmA;
      const int m_tile0 = 0;
      uint32_t st_pairs = 0;
      constexpr bool kBiasDirect = false;
      const float* const bias_t = nullptr;
      (void)bias_t;
      constexpr bool xin_tma = false, xin_live = false;  // TMA-loaded epilogue input: 2-CTA dGELU kernels only
      uint64_t* const xbar = nullptr;
      uint32_t xph = 0;
      const int num_clusters = 0;
      auto xin_issue = [](int, int) {};
      (void)stg; (void)tmC; (void)tmX; (void)m_tile0; (void)st_pairs; (void)xbar; (void)xph; (void)num_clusters; (void)xin_issue;
#include "gemm_epilogue.inc"
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, int A_MN, int B_MN, int OUT, int ACT>
static int launch_gemm(const XpGemm* g, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDev& dev, int grid,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kern = gemm_kernel<BN, A_MN, B_MN, OUT, ACT>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    XP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, dev);
  XP_CHECK_LAUNCH("gemm_kernel");
  return 0;
}

template <int BN, int OUT, int ACT>
static int dispatch_layout(const XpGemm* g, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDev& dev,
                           int grid, cudaStream_t stream) {
  if (g->a_layout == 0 && g->b_layout == 0) return launch_gemm<BN, 0, 0, OUT, ACT>(g, tmA, tmB, dev, grid, stream);
  if (g->a_layout == 0 && g->b_layout == 1) return launch_gemm<BN, 0, 1, OUT, ACT>(g, tmA, tmB, dev, grid, stream);
  if (g->a_layout == 1 && g->b_layout == 1) return launch_gemm<BN, 1, 1, OUT, ACT>(g, tmA, tmB, dev, grid, stream);
  if (g->a_layout == 1 && g->b_layout == 0) return launch_gemm<BN, 1, 0, OUT, ACT>(g, tmA, tmB, dev, grid, stream);
  return fail("xp_gemm: a_layout/b_layout must be 0 or 1");
}

// The activation epilogues exist for bf16 outputs only (forward activations / their gradients).
template <int BN>
static int dispatch_act_bf16(const XpGemm* g, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDev& dev, int grid,
                             cudaStream_t stream) {
  switch (g->act) {
    case XP_ACT_NONE: return dispatch_layout<BN, XP_OUT_BF16, XP_ACT_NONE>(g, tmA, tmB, dev, grid, stream);
    case XP_ACT_QUICK_GELU: return dispatch_layout<BN, XP_OUT_BF16, XP_ACT_QUICK_GELU>(g, tmA, tmB, dev, grid, stream);
    case XP_ACT_DQUICK_GELU: return dispatch_layout<BN, XP_OUT_BF16, XP_ACT_DQUICK_GELU>(g, tmA, tmB, dev, grid, stream);
    case XP_ACT_GELU_ERF: return dispatch_layout<BN, XP_OUT_BF16, XP_ACT_GELU_ERF>(g, tmA, tmB, dev, grid, stream);
    case XP_ACT_DGELU_ERF: return dispatch_layout<BN, XP_OUT_BF16, XP_ACT_DGELU_ERF>(g, tmA, tmB, dev, grid, stream);
  }
  return fail("xp_gemm: bad act");
}

template <int BN>
static int dispatch_out(const XpGemm* g, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDev& dev, int grid,
                        cudaStream_t stream) {
  switch (g->out) {
    case XP_OUT_BF16: return dispatch_act_bf16<BN>(g, tmA, tmB, dev, grid, stream);
    case XP_OUT_F32: return dispatch_layout<BN, XP_OUT_F32, XP_ACT_NONE>(g, tmA, tmB, dev, grid, stream);
    case XP_OUT_F32_ATOMIC: return dispatch_layout<BN, XP_OUT_F32_ATOMIC, XP_ACT_NONE>(g, tmA, tmB, dev, grid, stream);
  }
  return fail("xp_gemm: bad out mode");
}

#include "gemm_pair.inc"

// tensor maps of the TMA-store epilogue (set by xp_gemm before dispatch_pair; copies of tmA when unused)
static thread_local CUtensorMap t_tmC, t_tmX;

template <int A_MN, int B_MN, int OUT, int ACT>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDev& dev, int clusters,
                       cudaStream_t stream) {
  auto kern = gemm_pair_kernel<A_MN, B_MN, OUT, ACT>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    XP_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_BYTES));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = PAIR_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  XP_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, t_tmC, t_tmX, dev));
  XP_CHECK_LAUNCH("gemm_pair_kernel");
  return 0;
}

template <int OUT, int ACT>
static int dispatch_pair_layout(const XpGemm* g, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDev& dev,
                                int clusters, cudaStream_t stream) {
  if (g->a_layout == 0 && g->b_layout == 0) return launch_pair<0, 0, OUT, ACT>(tmA, tmB, dev, clusters, stream);
  if (g->a_layout == 0 && g->b_layout == 1) return launch_pair<0, 1, OUT, ACT>(tmA, tmB, dev, clusters, stream);
  if (g->a_layout == 1 && g->b_layout == 1) return launch_pair<1, 1, OUT, ACT>(tmA, tmB, dev, clusters, stream);
  if (g->a_layout == 1 && g->b_layout == 0) return launch_pair<1, 0, OUT, ACT>(tmA, tmB, dev, clusters, stream);
  return fail("xp_gemm: a_layout/b_layout must be 0 or 1");
}

static int dispatch_pair(const XpGemm* g, const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmDev& dev,
                         int clusters, cudaStream_t stream) {
  switch (g->out) {
    case XP_OUT_BF16:
      switch (g->act) {
        case XP_ACT_NONE: return dispatch_pair_layout<XP_OUT_BF16, XP_ACT_NONE>(g, tmA, tmB, dev, clusters, stream);
        case XP_ACT_QUICK_GELU: return dispatch_pair_layout<XP_OUT_BF16, XP_ACT_QUICK_GELU>(g, tmA, tmB, dev, clusters, stream);
        case XP_ACT_DQUICK_GELU: return dispatch_pair_layout<XP_OUT_BF16, XP_ACT_DQUICK_GELU>(g, tmA, tmB, dev, clusters, stream);
        case XP_ACT_GELU_ERF: return dispatch_pair_layout<XP_OUT_BF16, XP_ACT_GELU_ERF>(g, tmA, tmB, dev, clusters, stream);
        case XP_ACT_DGELU_ERF: return dispatch_pair_layout<XP_OUT_BF16, XP_ACT_DGELU_ERF>(g, tmA, tmB, dev, clusters, stream);
      }
      return fail("xp_gemm: bad act");
    case XP_OUT_F32: return dispatch_pair_layout<XP_OUT_F32, XP_ACT_NONE>(g, tmA, tmB, dev, clusters, stream);
    case XP_OUT_F32_ATOMIC: return dispatch_pair_layout<XP_OUT_F32_ATOMIC, XP_ACT_NONE>(g, tmA, tmB, dev, clusters, stream);
  }
  return fail("xp_gemm: bad out mode");
}

static int g_dbg_mn_lbo = 0, g_dbg_mn_sbo = 0;
}  // namespace xp

// Debug hook for tools/gemm_selftest (not part of the public ABI).
extern "C" void xp_debug_gemm_mn_desc(int lbo, int sbo) {
  xp::g_dbg_mn_lbo = lbo;
  xp::g_dbg_mn_sbo = sbo;
}

extern "C" int xp_gemm(const XpGemm* g, void* stream_v) {
  using namespace xp;
  if (!g) return fail("xp_gemm: null args");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  if (g->M <= 0 || g->N <= 0 || g->K <= 0) return fail("xp_gemm: M, N, K must be positive");
  if (g->N % 8 != 0) return fail("xp_gemm: N must be a multiple of 8");
  if (!g->a || !g->b || !g->c) return fail("xp_gemm: null operand");
  XP_ENTER(g->a);
  const int splits = g->splits <= 0 ? 1 : g->splits;
  if (splits > 1 && g->out != XP_OUT_F32_ATOMIC) return fail("xp_gemm: split-K requires XP_OUT_F32_ATOMIC");
  if ((g->act == XP_ACT_DQUICK_GELU || g->act == XP_ACT_DGELU_ERF) && !g->aux)
    return fail("xp_gemm: dGELU epilogue needs aux (the forward pre-activation)");
  if (g->act != XP_ACT_NONE && g->out != XP_OUT_BF16) return fail("xp_gemm: activation epilogues need a bf16 output");
  if ((g->act == XP_ACT_DQUICK_GELU || g->act == XP_ACT_DGELU_ERF) && g->residual)
    return fail("xp_gemm: a dGELU epilogue cannot be combined with a residual add");
  const int elem_c = g->out == XP_OUT_BF16 ? 2 : 4;
  if ((reinterpret_cast<uintptr_t>(g->c) & 15) || (g->ldc * elem_c) % 16)
    return fail("xp_gemm: C must be 16-byte aligned with a 16-byte multiple row pitch");
  if (g->bias && (reinterpret_cast<uintptr_t>(g->bias) & 15)) return fail("xp_gemm: bias must be 16-byte aligned");
  if (g->residual && ((reinterpret_cast<uintptr_t>(g->residual) & 15) || (g->ldr % 8)))
    return fail("xp_gemm: residual must be 16-byte aligned, ldr % 8 == 0");
  if (g->aux && ((reinterpret_cast<uintptr_t>(g->aux) & 15) || (g->ld_aux % 8)))
    return fail("xp_gemm: aux must be 16-byte aligned, ld_aux % 8 == 0");

  const int nsm = sm_count();
  const int num_m = static_cast<int>((g->M + BM - 1) / BM);
  int bn = g->block_n;
  if (bn == 0) {
    const long long tiles256 = static_cast<long long>(num_m) * ((g->N + 255) / 256) * splits;
    bn = (g->N >= 256 && tiles256 >= nsm) ? 256 : 128;
  }
  if (bn != 128 && bn != 256) return fail("xp_gemm: block_n must be 0, 128 or 256");
  // 2-CTA pairs (256 x 256 tiles, UMMA M = 256) whenever the problem is big enough; cta_pair: 0 auto, 1 never, 2 force
  bool pair = g->cta_pair == 2 || (g->cta_pair == 0 && g->block_n != 128 && g->N >= 256 && g->M >= 256);
  if (g->cta_pair < 0 || g->cta_pair > 2) return fail("xp_gemm: cta_pair must be 0, 1 or 2");
  if (pair) bn = 256;
  const long long total = pair ? static_cast<long long>((g->M + 2 * BM - 1) / (2 * BM)) * ((g->N + 255) / 256) * splits
                               : static_cast<long long>(num_m) * ((g->N + bn - 1) / bn) * splits;
  int grid = g->max_ctas > 0 ? g->max_ctas : nsm;
  if (pair) grid /= 2;   // clusters
  if (grid < 1) grid = 1;
  if (grid > total) grid = static_cast<int>(total);

  CUtensorMap tmA, tmB;
  int rc;
  if (g->a_layout == 0)
    rc = make_tmap_bf16_2d(&tmA, g->a, g->K, g->M, g->lda, BK, BM);
  else
    rc = make_tmap_bf16_2d(&tmA, g->a, g->M, g->K, g->lda, 64, BK);
  if (rc) return rc;
  if (g->b_layout == 0)
    rc = make_tmap_bf16_2d(&tmB, g->b, g->K, g->N, g->ldb, BK, pair ? 128 : bn);   // a pair CTA stages half of B
  else
    rc = make_tmap_bf16_2d(&tmB, g->b, g->N, g->K, g->ldb, 64, BK);
  if (rc) return rc;

  GemmDev dev;
  dev.c = g->c;
  dev.bias = g->bias;
  dev.residual = static_cast<const __nv_bfloat16*>(g->residual);
  dev.aux = static_cast<__nv_bfloat16*>(g->aux);
  dev.M = static_cast<int>(g->M);
  dev.N = static_cast<int>(g->N);
  dev.K = static_cast<int>(g->K);
  dev.ldc = g->ldc;
  dev.ldr = g->ldr;
  dev.ld_aux = g->ld_aux;
  dev.c_group = g->c_group;
  dev.c_group_stride = g->c_group_stride;
  dev.r_group = g->r_group;
  dev.r_group_stride = g->r_group_stride;
  dev.act = g->act;
  dev.splits = splits;
  dev.scale_cols = g->scale_cols;
  dev.alpha = g->alpha;
  dev.col_scale = g->col_scale;
  {
    auto ok32 = [](const void* ptr, long long ld) { return ptr == nullptr || ((reinterpret_cast<uintptr_t>(ptr) & 31) == 0 && ld % 16 == 0); };
    dev.wide = (g->out != XP_OUT_BF16 || ok32(g->c, g->ldc)) && ok32(g->aux, g->ld_aux) && ok32(g->residual, g->ldr) &&
               (g->c_group_stride % 16 == 0) && (g->r_group_stride % 16 == 0);
  }
  static const int gemm_dbg = [] { const char* e = getenv("XP_GEMM_DEBUG"); return e ? atoi(e) : 0; }();
  dev.dbg = gemm_dbg;
  dev.mn_lbo = g_dbg_mn_lbo ? g_dbg_mn_lbo : BK * 128;
  dev.mn_sbo = g_dbg_mn_sbo ? g_dbg_mn_sbo : 1024;

  dev.tma_c = dev.tma_aux = 0;
  t_tmC = tmA;
  t_tmX = tmA;
  static const bool tma_epi_off = getenv("XP_GEMM_NO_TMA_STORE") != nullptr;
  if (pair && g->out == XP_OUT_BF16 && g->c_group == 0 && !tma_epi_off) {
    // epilogue through shared memory + TMA stores: boxes of 64 columns x 32 rows (one epilogue warp's rows), 128B swizzle
    if (make_tmap_bf16_2d(&t_tmC, g->c, g->N, g->M, g->ldc, 64, 32)) return -1;
    dev.tma_c = 1;
    static const bool tma_aux_off = getenv("XP_GEMM_NO_TMA_AUX") != nullptr;
    // aux by TMA: stored by the GELU epilogues, loaded (next tile's boxes, ahead of time) by the dGELU epilogues
    if (g->aux && g->act != XP_ACT_NONE && splits == 1 && !tma_aux_off) {
      if (make_tmap_bf16_2d(&t_tmX, g->aux, g->N, g->M, g->ld_aux, 64, 32)) return -1;
      dev.tma_aux = 1;
    }
  }
  if (pair) return dispatch_pair(g, tmA, tmB, dev, grid, stream);
  return bn == 256 ? dispatch_out<256>(g, tmA, tmB, dev, grid, stream)
                   : dispatch_out<128>(g, tmA, tmB, dev, grid, stream);
}
