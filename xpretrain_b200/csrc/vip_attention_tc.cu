// tcgen05 / TMEM implementation of the video-proxy (ViP) attention of CLIP-ViP (CLIPAttention.forward2,
// CLIP_ViP.py:332-381) — forward.  Persistent, warp-specialised, one CTA per SM looping over (batch, head, frame)
// problems:
//
//   warps 0,2,3   producers: cp.async-stage the problem's q / k / v rows into shared memory in the UMMA
//                 SWIZZLE_128B layout (double buffered, so problem n+1 loads while problem n computes)
//   warp 1        MMA issuer (one thread): S = Q' K'^T  (tcgen05.mma SS, fp32 in TMEM), then O = P V' with the A
//                 operand read from TMEM (tcgen05.mma TS) and V' MN-major from shared memory
//   warps 4-7     softmax warpgroup of query tile 0 (rows 0..127), one thread per TMEM lane
//   warps 8-11    softmax warpgroup of query tile 1 (rows 128..255)
//
// Row layout per problem (frame t of video b, head h):
//   queries  sQ rows [0, L) = the frame's patch tokens, rows [200, 200+M) = the M global tokens (cls + proxies);
//            everything else is zero (written once) — row 200 keeps the global rows on a 1024-byte swizzle atom
//   keys     sK rows [0, L) frame keys (N = 208 MMA), sKg rows [0, M) global keys (N = 16 MMA); same for values
//   S tile   TMEM columns [0,208) frame keys | [208,224) global keys; P (packed bf16 pairs) overwrites columns
//            [0,112) of the dead S (FlashAttention-4 style aliasing); O accumulates in columns [128,192)
// Frame queries see [global ; own frame] keys (CLIP_ViP.py:352-363).  The global queries' softmax over all frames
// (:366-375) is assembled from per-frame partials (max, sum, unnormalised O) by vip_attn_fwd_combine; their global
// keys are counted by frame 0 only.  q arrives pre-scaled by head_dim**-0.5 (CLIP_ViP.py:341).
#include "../../include/xpretrain_b200.h"
#include "common.h"
#include "ptx.cuh"

namespace xp {

constexpr int TC_HD = 64;
constexpr int TC_FK = 208;        // frame-key columns (padded), multiple of 16
constexpr int TC_GK = 16;         // global-key columns (padded)
constexpr int TC_GROW = 200;      // smem / tile row of the first global query
constexpr int TC_QROWS = 256;
constexpr int TC_THREADS = 384;
constexpr int TC_BUF_BYTES = (TC_QROWS + 2 * TC_FK + 2 * TC_GK) * 128;   // 90112
constexpr float TC_LOG2E = 1.4426950408889634f;

struct TcDims {
  int B, H, T, L, M, C;
  long long S, ld_qkv, ld_o;
};

__device__ __forceinline__ uint32_t sw128(uint32_t base, int row, int chunk) {
  return base + row * 128 + ((chunk ^ (row & 7)) << 4);
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
__global__ void __launch_bounds__(TC_THREADS, 1)
vip_attn_fwd_tc_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                       float* __restrict__ part, const TcDims d) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t buf0 = (raw + 1023u) & ~1023u;
  uint8_t* tail = smem_raw + (buf0 + 2 * TC_BUF_BYTES - raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);      // [2] smem buffer filled
  uint64_t* empty = full + 2;                              // [2] smem buffer consumed
  uint64_t* s_ready = empty + 2;                           // [2] S of tile j in TMEM
  uint64_t* p_ready = s_ready + 2;                         // [2] P of tile j in TMEM
  uint64_t* o_ready = p_ready + 2;                         // [2] O of tile j in TMEM
  uint64_t* t_free = o_ready + 2;                          // [2] TMEM region of tile j drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_free + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = d.B * d.H * d.T;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 3);
      mbar_init(&empty[i], 1);
      mbar_init(&s_ready[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&o_ready[i], 1);
      mbar_init(&t_free[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // zero both buffers once: the producer only ever writes rows [0,L) / [200,200+M) / [0,M)
  for (int i = tid; i < 2 * TC_BUF_BYTES / 16; i += TC_THREADS)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(buf0 + i * 16), "r"(0) : "memory");
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 || warp == 2 || warp == 3) {
    // ------------------------------------------- producers: warp 0 -> Q rows, warp 2 -> K rows, warp 3 -> V rows
    const int mat = warp == 0 ? 0 : warp - 1;
    const int chunk = lane & 7, r0 = lane >> 3;     // 4 rows x 8 sixteen-byte chunks per warp instruction
    int n = 0;
    for (int prob = blockIdx.x; prob < total; prob += gridDim.x, ++n) {
      const int s = n & 1;
      const int t = prob % d.T, h = (prob / d.T) % d.H, b = prob / (d.T * d.H);
      mbar_wait(&empty[s], ((n >> 1) & 1) ^ 1);
      const uint32_t sQ = buf0 + s * TC_BUF_BYTES, sK = sQ + TC_QROWS * 128, sV = sK + TC_FK * 128;
      const uint32_t sKg = sV + TC_FK * 128, sVg = sKg + TC_GK * 128;
      const uint32_t dstF = mat == 0 ? sQ : (mat == 1 ? sK : sV);
      const uint32_t dstG = mat == 0 ? sQ + TC_GROW * 128 : (mat == 1 ? sKg : sVg);   // (TC_GROW & 7) == 0
      const __nv_bfloat16* gsrc = qkv + static_cast<long long>(b) * d.S * d.ld_qkv + mat * d.C + h * TC_HD + chunk * 8;
      const __nv_bfloat16* fsrc = gsrc + (d.M + static_cast<long long>(t) * d.L) * d.ld_qkv;
      for (int row = r0; row < d.L; row += 4) tc_cp_async16(sw128(dstF, row, chunk), fsrc + static_cast<long long>(row) * d.ld_qkv);
      for (int row = r0; row < d.M; row += 4) tc_cp_async16(sw128(dstG, row, chunk), gsrc + static_cast<long long>(row) * d.ld_qkv);
      asm volatile("cp.async.wait_all;" ::: "memory");
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_sf = make_idesc_bf16(128, TC_FK, 0, 0);
      constexpr uint32_t idesc_sg = make_idesc_bf16(128, TC_GK, 0, 0);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, TC_HD, 0, 1);
      int n = 0;
      for (int prob = blockIdx.x; prob < total; prob += gridDim.x, ++n) {
        const int s = n & 1;
        const uint32_t pp = n & 1;
        const uint32_t sQ = buf0 + s * TC_BUF_BYTES, sK = sQ + TC_QROWS * 128, sV = sK + TC_FK * 128;
        const uint32_t sKg = sV + TC_FK * 128, sVg = sKg + TC_GK * 128;
        mbar_wait(&full[s], (n >> 1) & 1);
        fence_proxy_async_smem();
        for (int j = 0; j < 2; ++j) {
          mbar_wait(&t_free[j], pp ^ 1);
          tc_fence_after();
          const uint32_t tS = tmem_base + j * 256;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t adesc = make_smem_desc_sw128(sQ + j * (128 * 128) + ks * 32, 16, 1024);
            umma_bf16(tS, adesc, make_smem_desc_sw128(sK + ks * 32, 16, 1024), idesc_sf, ks > 0 ? 1u : 0u);
            umma_bf16(tS + TC_FK, adesc, make_smem_desc_sw128(sKg + ks * 32, 16, 1024), idesc_sg, ks > 0 ? 1u : 0u);
          }
          umma_commit(&s_ready[j]);
        }
        for (int j = 0; j < 2; ++j) {
          mbar_wait(&p_ready[j], pp);
          tc_fence_after();
          const uint32_t tP = tmem_base + j * 256, tO = tP + 128;
#pragma unroll
          for (int ks = 0; ks < TC_FK / 16; ++ks)
            umma_bf16_ts(tO, tP + ks * 8, make_smem_desc_sw128(sV + ks * 2048, TC_FK * 128, 1024), idesc_o,
                         ks > 0 ? 1u : 0u);
          umma_bf16_ts(tO, tP + TC_FK / 2, make_smem_desc_sw128(sVg, TC_GK * 128, 1024), idesc_o, 1u);
          umma_commit(&o_ready[j]);
        }
        umma_commit(&empty[s]);   // every MMA that reads smem buffer s has been issued
      }
    }
  } else if (warp >= 4) {
    // --------------------------------------------------- softmax warpgroups (tile j = 0 / 1)
    const int j = (warp - 4) >> 2;
    const int wq = warp & 3;                       // TMEM lane quarter
    const int trow = wq * 32 + lane;               // row within the tile == TMEM lane
    const int row = j * 128 + trow;                // row within sQ
    const bool is_frame = row < d.L;
    const bool is_glob = row >= TC_GROW && row < TC_GROW + d.M;
    const uint32_t t_lane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + j * 256;
    // a warp whose 32 rows hold neither frame nor global queries only takes part in the barrier protocol
    const bool warp_active = (j * 128 + wq * 32 < d.L) || (j * 128 + wq * 32 + 31 >= TC_GROW && j * 128 + wq * 32 < TC_GROW + d.M);
    int n = 0;
    for (int prob = blockIdx.x; prob < total; prob += gridDim.x, ++n) {
      const uint32_t pp = n & 1;
      const int t = prob % d.T, h = (prob / d.T) % d.H, b = prob / (d.T * d.H);
      const bool mask_gk = is_glob && (t != 0);   // a global query counts the global keys in frame 0 only
      mbar_wait(&s_ready[j], pp);
      tc_fence_after();
      // ---- pass 1: row maximum (software-pipelined TMEM loads; chunks with only live keys skip the masks)
      float mx = -INFINITY;
      uint32_t r[2][16];
      if (warp_active) {
        tmem_ld16(t_lane, r[0]);
#pragma unroll
        for (int c = 0; c < (TC_FK + TC_GK) / 16; ++c) {
          tmem_ld_wait16(r[c & 1]);
          if (c + 1 < (TC_FK + TC_GK) / 16) tmem_ld16(t_lane + (c + 1) * 16, r[(c + 1) & 1]);
          if (c * 16 + 16 <= d.L) {
#pragma unroll
            for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(r[c & 1][i]));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = c * 16 + i;
              const bool dead = col < TC_FK ? (col >= d.L) : ((col - TC_FK) >= d.M || mask_gk);
              mx = fmaxf(mx, dead ? -INFINITY : __uint_as_float(r[c & 1][i]));
            }
          }
        }
      }
      const float mb = mx * TC_LOG2E;
      // ---- pass 2: P = exp(S - max), row sum; packed bf16 P overwrites S columns already consumed
      float sum = 0.f;
      if (warp_active) {
        tmem_ld16(t_lane, r[0]);
#pragma unroll
        for (int c = 0; c < (TC_FK + TC_GK) / 16; ++c) {
          tmem_ld_wait16(r[c & 1]);
          if (c + 1 < (TC_FK + TC_GK) / 16) tmem_ld16(t_lane + (c + 1) * 16, r[(c + 1) & 1]);
          uint32_t pk[8];
          float pv[16];
          if (c * 16 + 16 <= d.L) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              pv[i] = tc_exp2(fmaf(__uint_as_float(r[c & 1][i]), TC_LOG2E, -mb));
              sum += pv[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int col = c * 16 + i;
              const bool dead = col < TC_FK ? (col >= d.L) : ((col - TC_FK) >= d.M || mask_gk);
              pv[i] = dead ? 0.f : tc_exp2(fmaf(__uint_as_float(r[c & 1][i]), TC_LOG2E, -mb));
              sum += pv[i];
            }
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(pv[2 * i], pv[2 * i + 1]);
          // chunk c+1 (columns [16c+16, 16c+32)) is already in flight; P goes to columns [8c, 8c+8) < 16c+16
          tmem_st8(t_lane + c * 8, pk);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[j]);
      // ---- epilogue
      mbar_wait(&o_ready[j], pp);
      tc_fence_after();
      const long long srow = d.M + static_cast<long long>(t) * d.L + row;      // position in the sequence
      float* gp = part + (((static_cast<long long>(b) * d.H + h) * d.T + t) * d.M + (row - TC_GROW)) * 66;
      __nv_bfloat16* dst = out + (static_cast<long long>(b) * d.S + srow) * d.ld_o + h * TC_HD;
      const float inv = 1.f / sum;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t o[32];
        if (warp_active) {
          tmem_ld32(t_lane + 128 + half * 32, o);
          tmem_ld_wait(o);
        }
        if (half == 1) {
          tc_fence_before();
          mbar_arrive(&t_free[j]);    // registers hold the last of O: the next S may overwrite this TMEM region
        }
        if (is_glob) {
#pragma unroll
          for (int i = 0; i < 32; ++i) gp[2 + half * 32 + i] = __uint_as_float(o[i]);
        } else if (is_frame) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 v;
            v.x = pack_bf16(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
            v.y = pack_bf16(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
            v.z = pack_bf16(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
            v.w = pack_bf16(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + half * 32 + q * 8) = v;
          }
        }
      }
      if (is_glob) {
        gp[0] = mx;
        gp[1] = sum;
      } else if (is_frame) {
        lse[(static_cast<long long>(b) * d.H + h) * d.S + srow] = mx + logf(sum);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace xp

using namespace xp;

extern "C" int xp_vip_attention_fwd_tc_partial(const void* qkv, void* out, float* lse, float* workspace, int32_t B,
                                               int32_t H, int32_t T, int32_t L, int32_t M, int32_t C, void* stream) {
  XP_ENTER(qkv);
  if (C != H * TC_HD) return fail("vip_attention: head_dim must be 64 (C == 64*H)");
  if (L < 1 || L > 196) return fail("vip_attention: 1 <= L <= 196 patch tokens per frame");
  if (M < 1 || M > 8) return fail("vip_attention: 1 <= M <= 8 global tokens");
  TcDims d;
  d.B = B; d.H = H; d.T = T; d.L = L; d.M = M; d.C = C;
  d.S = static_cast<long long>(M) + static_cast<long long>(T) * L;
  d.ld_qkv = 3LL * C;
  d.ld_o = C;
  const int smem = 2 * TC_BUF_BYTES + 1024 + 128;
  static bool attr = false;
  if (!attr) {
    XP_CHECK_CUDA(cudaFuncSetAttribute(vip_attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const long long total = static_cast<long long>(B) * H * T;
  const int grid = static_cast<int>(total < sm_count() ? total : sm_count());
  vip_attn_fwd_tc_kernel<<<grid, TC_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), lse, workspace, d);
  XP_CHECK_LAUNCH("vip_attn_fwd_tc_kernel");
  return 0;
}
