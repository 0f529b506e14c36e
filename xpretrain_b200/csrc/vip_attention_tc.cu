// tcgen05 / TMEM implementation of the video-proxy (ViP) attention of CLIP-ViP (CLIPAttention.forward2,
// CLIP_ViP.py:332-381) — forward.  One CTA per (batch, head, frame), 128 threads = one thread per TMEM lane
// (query row), two CTAs per SM (256 TMEM columns each).
//
//   stage   q/k/v rows {0..M-1} U {M+t*L ..} of the fused qkv buffer -> shared memory in the UMMA
//           SWIZZLE_128B K-major layout (cp.async; the same bytes TMA would have produced)
//   S       = Q'[128 x 64] . K'[208 x 64]^T      tcgen05.mma (SS), fp32 in TMEM columns [0, 208)
//   softmax thread-per-row on TMEM (tcgen05.ld), P written back as packed bf16 over the dead S columns
//           (tcgen05.st) — FlashAttention-4 style S/P aliasing, no shared-memory round trip
//   O       = P[128 x 208] . V'[208 x 64]        tcgen05.mma with the A operand read from TMEM (TS), V' MN-major
//   epilogue tcgen05.ld O, normalise, 128-byte row stores; the M global-query rows emit per-frame partials
//           (max, sum, unnormalised O) merged by vip_attn_fwd_combine (vip_attention.cu).
// q arrives pre-scaled by head_dim**-0.5 from the QKV GEMM epilogue (CLIP_ViP.py:341).
#include "../../include/xpretrain_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace xp {

constexpr int TC_HD = 64;
constexpr int TC_KEYS = 208;      // padded keys per CTA (M + L <= 208), multiple of 16
constexpr int TC_QROWS = 256;     // two 128-row query tiles
constexpr int TC_THREADS = 128;
constexpr float TC_LOG2E = 1.4426950408889634f;

struct TcDims {
  int B, H, T, L, M, C;
  long long S, ld_qkv, ld_o;
};

__device__ __forceinline__ uint32_t sw128(uint32_t base, int row, int chunk) {
  return base + row * 128 + ((chunk ^ (row & 7)) << 4);
}
__device__ __forceinline__ long long tc_token(const TcDims& d, int b, int t, int i) {
  return static_cast<long long>(b) * d.S + (i < d.M ? i : d.M + static_cast<long long>(t) * d.L + (i - d.M));
}

// A operand from TMEM (packed bf16 pairs, lane = row), B from shared memory.
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// tcgen05.wait::ld tied to the destination registers (their first use cannot be scheduled above the wait)
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tc_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tc_cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

// part: [B, H, T, M, 66] fp32 = {max, sum, unnormalised out[64]} of the global queries over this frame's keys.
__global__ void __launch_bounds__(TC_THREADS, 2)
vip_attn_fwd_tc_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                       float* __restrict__ part, const TcDims d) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sQ = (raw + 1023u) & ~1023u;
  const uint32_t sK = sQ + TC_QROWS * 128;
  const uint32_t sV = sK + TC_KEYS * 128;
  uint8_t* tail = smem_raw + (sV + TC_KEYS * 128 - raw);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(tail);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 8);

  const int t = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int nq = d.M + d.L;

  if (tid == 0) {
    mbar_init(mbar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  // ---- stage Q' (256 rows), K', V' (208 rows); rows past M+L are zero
  for (int idx = tid; idx < (TC_QROWS + 2 * TC_KEYS) * 8; idx += TC_THREADS) {
    int mat, row;
    const int r8 = idx >> 3, chunk = idx & 7;
    if (r8 < TC_QROWS) { mat = 0; row = r8; }
    else if (r8 < TC_QROWS + TC_KEYS) { mat = 1; row = r8 - TC_QROWS; }
    else { mat = 2; row = r8 - TC_QROWS - TC_KEYS; }
    const uint32_t dst = sw128(mat == 0 ? sQ : (mat == 1 ? sK : sV), row, chunk);
    if (row < nq)
      tc_cp_async16(dst, qkv + tc_token(d, b, t, row) * d.ld_qkv + static_cast<long long>(mat) * d.C + h * TC_HD + chunk * 8);
    else
      asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst), "r"(0) : "memory");
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

  constexpr uint32_t idesc_s = make_idesc_bf16(128, TC_KEYS, 0, 0);   // S = Q' K'^T, both K-major
  constexpr uint32_t idesc_o = make_idesc_bf16(128, TC_HD, 0, 1);     // O = P V', P from TMEM, V' MN-major
  const int ntiles = (nq + 127) / 128;
  uint32_t phase = 0;

  for (int tile = 0; tile < ntiles; ++tile) {
    if (tid == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma_bf16(tmem_base, make_smem_desc_sw128(sQ + tile * (128 * 128) + ks * 32, 16, 1024),
                  make_smem_desc_sw128(sK + ks * 32, 16, 1024), idesc_s, ks > 0 ? 1u : 0u);
      umma_commit(mbar);
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    tc_fence_after();

    const int row = tile * 128 + tid;
    const bool grow = row < d.M;                 // a global (cls / proxy) query row
    const bool mask_gk = grow && (t != 0);       // its global keys are counted by frame 0 only
    // ---- pass 1: row maximum (keys >= nq are padding)
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < TC_KEYS / 16; ++c) {
      uint32_t r[16];
      tmem_ld16(t_lane + c * 16, r);
      tmem_ld_wait16(r);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int key = c * 16 + i;
        const bool dead = key >= nq || (mask_gk && key < d.M);
        mx = fmaxf(mx, dead ? -INFINITY : __uint_as_float(r[i]));
      }
    }
    const float mb = mx * TC_LOG2E;
    // ---- pass 2: P = exp(S - max), row sum; packed bf16 P overwrites the S columns already consumed
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < TC_KEYS / 16; ++c) {
      uint32_t r[16], pk[8];
      tmem_ld16(t_lane + c * 16, r);
      tmem_ld_wait16(r);
      float pv[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int key = c * 16 + i;
        const bool dead = key >= nq || (mask_gk && key < d.M);
        pv[i] = dead ? 0.f : tc_exp2(fmaf(__uint_as_float(r[i]), TC_LOG2E, -mb));
        sum += pv[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(pv[2 * i], pv[2 * i + 1]);
      tmem_st8(t_lane + c * 8, pk);   // columns [8c, 8c+8) <= columns already read ([0, 16c+16))
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < TC_KEYS / 16; ++ks)
        umma_bf16_ts(tmem_base + 128, tmem_base + ks * 8, make_smem_desc_sw128(sV + ks * 2048, TC_KEYS * 128, 1024),
                     idesc_o, ks > 0 ? 1u : 0u);
      umma_commit(mbar);
    }
    mbar_wait(mbar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue
    uint32_t o[2][32];
    tmem_ld32(t_lane + 128, o[0]);
    tmem_ld32(t_lane + 160, o[1]);
    tmem_ld_wait(o[0]);
    tmem_ld_wait(o[1]);
    if (row < nq) {
      if (grow) {
        float* p = part + (((static_cast<long long>(b) * d.H + h) * d.T + t) * d.M + row) * 66;
        p[0] = mx;
        p[1] = sum;
#pragma unroll
        for (int i = 0; i < 64; ++i) p[2 + i] = __uint_as_float(o[i >> 5][i & 31]);
      } else {
        const float inv = 1.f / sum;
        const long long tok = tc_token(d, b, t, row);
        __nv_bfloat16* dst = out + tok * d.ld_o + h * TC_HD;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4 v;
          const uint32_t* s = &o[q >> 2][(q & 3) * 8];
          v.x = pack_bf16(__uint_as_float(s[0]) * inv, __uint_as_float(s[1]) * inv);
          v.y = pack_bf16(__uint_as_float(s[2]) * inv, __uint_as_float(s[3]) * inv);
          v.z = pack_bf16(__uint_as_float(s[4]) * inv, __uint_as_float(s[5]) * inv);
          v.w = pack_bf16(__uint_as_float(s[6]) * inv, __uint_as_float(s[7]) * inv);
          *reinterpret_cast<uint4*>(dst + q * 8) = v;
        }
        lse[(static_cast<long long>(b) * d.H + h) * d.S + (tok - static_cast<long long>(b) * d.S)] = mx + logf(sum);
      }
    }
    tc_fence_before();
    __syncthreads();   // every thread is done with this tile's TMEM before the next S overwrites it
    tc_fence_after();
  }
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace xp

using namespace xp;

extern "C" int xp_vip_attention_fwd_tc_partial(const void* qkv, void* out, float* lse, float* workspace, int32_t B,
                                               int32_t H, int32_t T, int32_t L, int32_t M, int32_t C, void* stream) {
  XP_ENTER(qkv);
  if (C != H * TC_HD) return fail("vip_attention: head_dim must be 64 (C == 64*H)");
  if (M + L > TC_KEYS) return fail("vip_attention: M + L must be <= 208");
  if (M < 1 || M > 8) return fail("vip_attention: 1 <= M <= 8 global tokens");
  TcDims d;
  d.B = B; d.H = H; d.T = T; d.L = L; d.M = M; d.C = C;
  d.S = static_cast<long long>(M) + static_cast<long long>(T) * L;
  d.ld_qkv = 3LL * C;
  d.ld_o = C;
  const int smem = (TC_QROWS + 2 * TC_KEYS) * 128 + 1024 + 64;
  static bool attr = false;
  if (!attr) {
    XP_CHECK_CUDA(cudaFuncSetAttribute(vip_attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  vip_attn_fwd_tc_kernel<<<dim3(T, H, B), TC_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), lse, workspace, d);
  XP_CHECK_LAUNCH("vip_attn_fwd_tc_kernel");
  return 0;
}
