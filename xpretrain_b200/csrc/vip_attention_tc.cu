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

// =====================================================================================================
// Backward.  Persistent, one CTA per SM, 14 warps:
//   warps 0,2,3,12  producers (cp.async): Q' / dO' rows (same row layout as forward), K' (+global), V' (+global),
//                   and lse / delta of the query rows
//   warp 1          MMA issuer (one thread)
//   warps 4-11      two math warpgroups: thread = query row of the current 128-row q tile; WG A owns key
//                   columns [0,64) of the step, WG B columns [64,128); in the epilogues WG A drains dV and dQ_0,
//                   WG B dK and dQ_1
// A problem = 2 key tiles x 2 query tiles = 4 steps of [128 q x 128 keys]:
//   S  = Q_j K_i^T, dP = dO_j V_i^T                (tcgen05.mma SS -> TMEM columns [0,128), [128,256))
//   P  = exp(S - lse), dS = P * (dP - delta)       (threads; bf16 P and dS written to shared memory, SWIZZLE_128B)
//   dV_i += P^T dO_j, dK_i += dS^T Q_j             (A = P / dS read MN-major from shared memory; TMEM [256,384))
//   dQ_j += dS K_i                                 (A = dS read K-major; TMEM [384,512))
// Key tile 0 = frame keys [0,128); key tile 1 = frame keys [128,208) in columns [0,80) + the M global keys in
// columns [80,96).  Query tile 0 = frame rows [0,128); tile 1 = frame rows [128,196) + global rows at 200..
// delta_i = sum_d dO_id O_id is precomputed (vip_attn_delta_kernel).  The gradients of the M global rows are
// emitted as per-frame fp32 partials and reduced by vip_attn_bwd_combine (vip_attention.cu).
constexpr int TB_THREADS = 448;
constexpr int TB_SQ = 0;                                   // [256][128 B]
constexpr int TB_SDO = TB_SQ + 256 * 128;                  // [256][128 B]
constexpr int TB_SK = TB_SDO + 256 * 128;                  // [208][128 B]
constexpr int TB_SKG = TB_SK + TC_FK * 128;                // [16][128 B]
constexpr int TB_SV = TB_SKG + TC_GK * 128;
constexpr int TB_SVG = TB_SV + TC_FK * 128;
constexpr int TB_SP = TB_SVG + TC_GK * 128;                // [2 atoms][128 rows][128 B]
constexpr int TB_SDS = TB_SP + 2 * 128 * 128;
constexpr int TB_STAT = TB_SDS + 2 * 128 * 128;            // lse*log2e [256] f32, delta [256] f32
constexpr int TB_BAR = TB_STAT + 2 * 256 * 4;
constexpr int TB_SMEM = TB_BAR + 128;

__global__ void __launch_bounds__(256)
vip_attn_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                      float* __restrict__ delta, long long rows, long long S, int C, int H) {
  // one warp per token row; 8 lanes x 8 elements cover one 64-wide head
  const long long r = static_cast<long long>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const long long b = r / S, s = r - b * S;
  for (int c0 = lane * 8; c0 < C; c0 += 256) {
    const uint4 o = *reinterpret_cast<const uint4*>(out + r * C + c0);
    const uint4 g = *reinterpret_cast<const uint4*>(dout + r * C + c0);
    const uint32_t ow[4] = {o.x, o.y, o.z, o.w}, gw[4] = {g.x, g.y, g.z, g.w};
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) dot += bf16_lo(ow[i]) * bf16_lo(gw[i]) + bf16_hi(ow[i]) * bf16_hi(gw[i]);
    dot += __shfl_xor_sync(0xffffffffu, dot, 1);
    dot += __shfl_xor_sync(0xffffffffu, dot, 2);
    dot += __shfl_xor_sync(0xffffffffu, dot, 4);
    if ((lane & 7) == 0) delta[(b * H + c0 / TC_HD) * S + s] = dot;
  }
}

__global__ void __launch_bounds__(TB_THREADS, 1)
vip_attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_qkv_f, const __grid_constant__ CUtensorMap tm_qkv_g,
                       const __grid_constant__ CUtensorMap tm_do_f, const __grid_constant__ CUtensorMap tm_do_g,
                       const float* __restrict__ lse, const float* __restrict__ delta,
                       __nv_bfloat16* __restrict__ dqkv, float* __restrict__ gpart, const TcDims d, float q_scale) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  float* s_lse = reinterpret_cast<float*>(gbase + TB_STAT);
  float* s_del = s_lse + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(gbase + TB_BAR);
  uint64_t* full = bars + 0;        // operands of problem n staged (4 producer warps)
  uint64_t* smem_free = bars + 1;   // every MMA of problem n retired
  uint64_t* sd_ready = bars + 2;    // S, dP of a step in TMEM
  uint64_t* pds_ready = bars + 3;   // P, dS of a step in shared memory (256 threads)
  uint64_t* kv_ready = bars + 4;    // dK_i, dV_i complete
  uint64_t* kv_free = bars + 5;     // ... and drained (256 threads)
  uint64_t* q_ready = bars + 6;     // dQ_0, dQ_1 complete
  uint64_t* q_free = bars + 7;      // ... and drained (256 threads)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = d.B * d.H * d.T;

  if (tid == 0) {
    mbar_init(full, 3);              // TMA expect_tx arrival + the two statistics warps
    mbar_init(smem_free, 1);
    mbar_init(sd_ready, 1);
    mbar_init(pds_ready, 256);
    mbar_init(kv_ready, 1);
    mbar_init(kv_free, 256);
    mbar_init(q_ready, 1);
    mbar_init(q_free, 256);
    fence_barrier_init();
  }
  if (warp == 13) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = tid; i < TB_STAT / 16; i += TB_THREADS)   // zero all operand tiles once (padding rows stay zero)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + i * 16), "r"(0) : "memory");
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320,
                 tdQ = tmem_base + 384;

  if (warp == 0 || warp == 12) {
    // ----------------------------------------------------------------------------------- producers
    // warp 0 lane 0: eight TMA box loads per problem — the L frame rows and the M global rows of the Q', K', V' head
    // slices of qkv and of dO — land in the 128B-swizzled tiles the UMMA descriptors read (padding rows stay zero from
    // the initial clear); warps 0 / 12 also stage lse * log2(e) / delta of the query rows.
    const int role = warp == 0 ? 0 : 1;
    int n = 0;
    for (int prob = blockIdx.x; prob < total; prob += gridDim.x, ++n) {
      const int t = prob % d.T, h = (prob / d.T) % d.H, b = prob / (d.T * d.H);
      mbar_wait(smem_free, (n & 1) ^ 1);
      const long long tok_g = static_cast<long long>(b) * d.S, tok_f = tok_g + d.M + static_cast<long long>(t) * d.L;
      if (role == 0 && lane == 0) {
        mbar_arrive_expect_tx(full, 4u * static_cast<uint32_t>(d.L + d.M) * 128u);
        const int rf = static_cast<int>(tok_f), rg = static_cast<int>(tok_g);
        const int cq = h * TC_HD, ck = d.C + h * TC_HD, cv = 2 * d.C + h * TC_HD;
        tma_load_2d(gbase + TB_SQ, &tm_qkv_f, full, cq, rf);
        tma_load_2d(gbase + TB_SQ + TC_GROW * 128, &tm_qkv_g, full, cq, rg);
        tma_load_2d(gbase + TB_SDO, &tm_do_f, full, cq, rf);
        tma_load_2d(gbase + TB_SDO + TC_GROW * 128, &tm_do_g, full, cq, rg);
        tma_load_2d(gbase + TB_SK, &tm_qkv_f, full, ck, rf);
        tma_load_2d(gbase + TB_SKG, &tm_qkv_g, full, ck, rg);
        tma_load_2d(gbase + TB_SV, &tm_qkv_f, full, cv, rf);
        tma_load_2d(gbase + TB_SVG, &tm_qkv_g, full, cv, rg);
      }
      {   // per-row statistics of the query rows: lse * log2(e) (+inf on padding rows) / delta
        const float* src = (role == 0 ? lse : delta) + (static_cast<long long>(b) * d.H + h) * d.S;
        float* dst = role == 0 ? s_lse : s_del;
        for (int row = lane; row < 256; row += 32) {
          float v = role == 0 ? INFINITY : 0.f;
          if (row < d.L) v = src[d.M + static_cast<long long>(t) * d.L + row] * (role == 0 ? TC_LOG2E : 1.f);
          else if (row >= TC_GROW && row < TC_GROW + d.M) v = src[row - TC_GROW] * (role == 0 ? TC_LOG2E : 1.f);
          dst[row] = v;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(full);
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t id_s128 = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t id_s80 = make_idesc_bf16(128, 80, 0, 0);
      constexpr uint32_t id_s16 = make_idesc_bf16(128, 16, 0, 0);
      constexpr uint32_t id_kv = make_idesc_bf16(128, TC_HD, 1, 1);   // A = P / dS MN-major, B = dO / Q MN-major
      constexpr uint32_t id_q = make_idesc_bf16(128, TC_HD, 0, 1);    // A = dS K-major, B = K MN-major
      const uint32_t sQ = base + TB_SQ, sdO = base + TB_SDO, sK = base + TB_SK, sKg = base + TB_SKG, sV = base + TB_SV,
                     sVg = base + TB_SVG, sP = base + TB_SP, sdS = base + TB_SDS;
      int n = 0;
      uint32_t step = 0, kvt = 0;
      for (int prob = blockIdx.x; prob < total; prob += gridDim.x, ++n) {
        mbar_wait(full, n & 1);
        fence_proxy_async_smem();
        for (int i = 0; i < 2; ++i) {
          for (int j = 0; j < 2; ++j, ++step) {
            // ---- S = Q_j K_i^T, dP = dO_j V_i^T
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t aq = make_smem_desc_sw128(sQ + j * 16384 + ks * 32, 16, 1024);
              const uint64_t ao = make_smem_desc_sw128(sdO + j * 16384 + ks * 32, 16, 1024);
              const uint32_t acc = ks > 0 ? 1u : 0u;
              if (i == 0) {
                umma_bf16(tS, aq, make_smem_desc_sw128(sK + ks * 32, 16, 1024), id_s128, acc);
                umma_bf16(tdP, ao, make_smem_desc_sw128(sV + ks * 32, 16, 1024), id_s128, acc);
              } else {
                umma_bf16(tS, aq, make_smem_desc_sw128(sK + 128 * 128 + ks * 32, 16, 1024), id_s80, acc);
                umma_bf16(tS + 80, aq, make_smem_desc_sw128(sKg + ks * 32, 16, 1024), id_s16, acc);
                umma_bf16(tdP, ao, make_smem_desc_sw128(sV + 128 * 128 + ks * 32, 16, 1024), id_s80, acc);
                umma_bf16(tdP + 80, ao, make_smem_desc_sw128(sVg + ks * 32, 16, 1024), id_s16, acc);
              }
            }
            umma_commit(sd_ready);
            // ---- wait for P, dS of this step, then the three gradient products
            mbar_wait(pds_ready, step & 1);
            fence_proxy_async_smem();
            tc_fence_after();
            if (j == 0) {
              mbar_wait(kv_free, (kvt & 1) ^ 1);   // dK_i / dV_i accumulators drained by the previous key tile
              tc_fence_after();
            }
            if (i == 0 && j == 0) {
              mbar_wait(q_free, (n & 1) ^ 1);
              tc_fence_after();
            }
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {   // K = the 128 query rows of tile j
              const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
              umma_bf16(tdV, make_smem_desc_sw128(sP + ks * 2048, 16384, 1024),
                        make_smem_desc_sw128(sdO + j * 16384 + ks * 2048, 16384, 1024), id_kv, acc);
              umma_bf16(tdK, make_smem_desc_sw128(sdS + ks * 2048, 16384, 1024),
                        make_smem_desc_sw128(sQ + j * 16384 + ks * 2048, 16384, 1024), id_kv, acc);
            }
            const uint32_t tq = tdQ + j * 64;
            if (i == 0) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)   // K = frame keys [0,128): atom ks/4, 32 B per k-step inside it
                umma_bf16(tq, make_smem_desc_sw128(sdS + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                          make_smem_desc_sw128(sK + ks * 2048, 16384, 1024), id_q, ks > 0 ? 1u : 0u);
            } else {
#pragma unroll
              for (int ks = 0; ks < 5; ++ks)   // frame keys [128,208)
                umma_bf16(tq, make_smem_desc_sw128(sdS + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024),
                          make_smem_desc_sw128(sK + (128 + ks * 16) * 128, 16384, 1024), id_q, 1u);
              umma_bf16(tq, make_smem_desc_sw128(sdS + 16384 + 32, 16, 1024),   // columns [80,96): the global keys
                        make_smem_desc_sw128(sKg, 16384, 1024), id_q, 1u);
            }
            if (j == 1) {
              umma_commit(kv_ready);
              ++kvt;
            }
          }
        }
        umma_commit(q_ready);
        umma_commit(smem_free);
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------------------------------------------------- math warpgroups
    const int wg = (warp - 4) >> 2;               // 0: key columns [0,64) / dV / dQ_0;  1: [64,128) / dK / dQ_1
    const int wq = warp & 3;
    const int trow = wq * 32 + lane;               // TMEM lane == row of the current tile
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t sP = base + TB_SP + wg * 16384, sdS = base + TB_SDS + wg * 16384;
    int n = 0;
    uint32_t step = 0, kvt = 0;
    for (int prob = blockIdx.x; prob < total; prob += gridDim.x, ++n) {
      const int t = prob % d.T, h = (prob / d.T) % d.H, b = prob / (d.T * d.H);
      const long long tok_g = static_cast<long long>(b) * d.S, tok_f = tok_g + d.M + static_cast<long long>(t) * d.L;
      float* gp = gpart + ((static_cast<long long>(b) * d.H + h) * d.T + t) * d.M * 3 * TC_HD;
      for (int i = 0; i < 2; ++i) {
        for (int j = 0; j < 2; ++j, ++step) {
          const int row = j * 128 + trow;
          const bool q_glob = row >= TC_GROW && row < TC_GROW + d.M;
          mbar_wait(sd_ready, step & 1);
          tc_fence_after();
          const float l2 = s_lse[row], dl = s_del[row];   // +inf lse on padding rows -> P = 0
#pragma unroll
          for (int c2 = 0; c2 < 2; ++c2) {                 // 32 key columns per TMEM load pair
            uint32_t rs32[32], rp32[32];
            tmem_ld32(tS + lane_off + wg * 64 + c2 * 32, rs32);
            tmem_ld32(tdP + lane_off + wg * 64 + c2 * 32, rp32);
            tmem_ld_wait(rs32);
            tmem_ld_wait(rp32);
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {                 // 16 key columns = two 16-byte chunks of P and of dS
            const int c = c2 * 2 + ch;
            const int col0 = wg * 64 + c * 16;
            const uint32_t* rs = rs32 + ch * 16;
            const uint32_t* rp = rp32 + ch * 16;
            uint32_t pk[8], dk[8];
#pragma unroll
            for (int e = 0; e < 16; e += 2) {
              float pv[2], dv[2];
#pragma unroll
              for (int u = 0; u < 2; ++u) {
                const int col = col0 + e + u;
                bool live;
                if (i == 0) live = col < d.L;
                else live = col < 80 ? (128 + col < d.L) : (col < 96 && (col - 80) < d.M && !(q_glob && t != 0));
                // dead columns may hold stale TMEM bits (even NaN): select, never multiply by them
                const float p = live ? tc_exp2(fmaf(__uint_as_float(rs[e + u]), TC_LOG2E, -l2)) : 0.f;
                pv[u] = p;
                dv[u] = live ? p * (__uint_as_float(rp[e + u]) - dl) : 0.f;
              }
              pk[e >> 1] = pack_bf16(pv[0], pv[1]);
              dk[e >> 1] = pack_bf16(dv[0], dv[1]);
            }
            // row `trow`, 16-byte chunks 2c and 2c+1 of this warpgroup's 64-column atom
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sw128(sP, trow, 2 * c)), "r"(pk[0]), "r"(pk[1]),
                         "r"(pk[2]), "r"(pk[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sw128(sP, trow, 2 * c + 1)), "r"(pk[4]),
                         "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sw128(sdS, trow, 2 * c)), "r"(dk[0]), "r"(dk[1]),
                         "r"(dk[2]), "r"(dk[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sw128(sdS, trow, 2 * c + 1)), "r"(dk[4]),
                         "r"(dk[5]), "r"(dk[6]), "r"(dk[7]) : "memory");
          }
          }
          fence_proxy_async_smem();
          tc_fence_before();
          mbar_arrive(pds_ready);
          if (j == 1) {
            // ---- dV_i (WG A) / dK_i (WG B): thread = key row of key tile i
            mbar_wait(kv_ready, kvt & 1);
            ++kvt;
            tc_fence_after();
            int key = -1, gk = -1;
            if (i == 0) { if (trow < d.L) key = trow; }
            else if (trow < 80) { if (128 + trow < d.L) key = 128 + trow; }
            else if (trow < 80 + d.M) gk = trow - 80;
            const uint32_t tsrc = (wg == 0 ? tdV : tdK) + lane_off;
            __nv_bfloat16* drow = dqkv + (tok_f + key) * d.ld_qkv + (wg == 0 ? 2 : 1) * d.C + h * TC_HD;
            float* grow = gp + static_cast<long long>(gk) * 3 * TC_HD + (wg == 0 ? 2 : 1) * TC_HD;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t o[32];
              tmem_ld32(tsrc + half * 32, o);
              tmem_ld_wait(o);
              if (half == 1) {
                tc_fence_before();
                mbar_arrive(kv_free);
              }
              if (key >= 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  uint4 v;
                  v.x = pack_bf16(__uint_as_float(o[q * 8 + 0]), __uint_as_float(o[q * 8 + 1]));
                  v.y = pack_bf16(__uint_as_float(o[q * 8 + 2]), __uint_as_float(o[q * 8 + 3]));
                  v.z = pack_bf16(__uint_as_float(o[q * 8 + 4]), __uint_as_float(o[q * 8 + 5]));
                  v.w = pack_bf16(__uint_as_float(o[q * 8 + 6]), __uint_as_float(o[q * 8 + 7]));
                  *reinterpret_cast<uint4*>(drow + half * 32 + q * 8) = v;
                }
              } else if (gk >= 0) {
#pragma unroll
                for (int e = 0; e < 32; ++e) grow[half * 32 + e] = __uint_as_float(o[e]);
              }
            }
          }
        }
      }
      // ---- dQ_0 (WG A) / dQ_1 (WG B): thread = query row
      {
        mbar_wait(q_ready, n & 1);
        tc_fence_after();
        const int row = wg * 128 + trow;
        const bool is_frame = row < d.L;
        const int gq = (row >= TC_GROW && row < TC_GROW + d.M) ? row - TC_GROW : -1;
        __nv_bfloat16* drow = dqkv + (tok_f + row) * d.ld_qkv + h * TC_HD;
        float* grow = gp + static_cast<long long>(gq) * 3 * TC_HD;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t o[32];
          tmem_ld32(tdQ + wg * 64 + lane_off + half * 32, o);
          tmem_ld_wait(o);
          if (half == 1) {
            tc_fence_before();
            mbar_arrive(q_free);
          }
          if (is_frame) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              uint4 v;
              v.x = pack_bf16(__uint_as_float(o[q * 8 + 0]) * q_scale, __uint_as_float(o[q * 8 + 1]) * q_scale);
              v.y = pack_bf16(__uint_as_float(o[q * 8 + 2]) * q_scale, __uint_as_float(o[q * 8 + 3]) * q_scale);
              v.z = pack_bf16(__uint_as_float(o[q * 8 + 4]) * q_scale, __uint_as_float(o[q * 8 + 5]) * q_scale);
              v.w = pack_bf16(__uint_as_float(o[q * 8 + 6]) * q_scale, __uint_as_float(o[q * 8 + 7]) * q_scale);
              *reinterpret_cast<uint4*>(drow + half * 32 + q * 8) = v;
            }
          } else if (gq >= 0) {
#pragma unroll
            for (int e = 0; e < 32; ++e) grow[half * 32 + e] = __uint_as_float(o[e]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 13) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


extern "C" int xp_vip_attention_bwd_tc_partial(const void* qkv, const void* out, const void* dout, const float* lse,
                                               void* dqkv, float* workspace, float* delta, int32_t B, int32_t H,
                                               int32_t T, int32_t L, int32_t M, int32_t C, float q_scale, void* stream) {
  using namespace xp;
  XP_ENTER(qkv);
  if (C != H * TC_HD) return fail("vip_attention: head_dim must be 64 (C == 64*H)");
  if (L < 1 || L > 196) return fail("vip_attention: 1 <= L <= 196 patch tokens per frame");
  if (M < 1 || M > 8) return fail("vip_attention: 1 <= M <= 8 global tokens");
  TcDims d;
  d.B = B; d.H = H; d.T = T; d.L = L; d.M = M; d.C = C;
  d.S = static_cast<long long>(M) + static_cast<long long>(T) * L;
  d.ld_qkv = 3LL * C;
  d.ld_o = C;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long rows = static_cast<long long>(B) * d.S;
  vip_attn_delta_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, st>>>(
      static_cast<const __nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(dout), delta, rows, d.S, C, H);
  XP_CHECK_LAUNCH("vip_attn_delta_kernel");
  const int smem = TB_SMEM + 1024;
  static bool attr = false;
  if (!attr) {
    XP_CHECK_CUDA(cudaFuncSetAttribute(vip_attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const long long total = static_cast<long long>(B) * H * T;
  const int grid = static_cast<int>(total < sm_count() ? total : sm_count());
  // TMA boxes: 64 columns (one head slice) x L frame rows / x M global rows of the token-major buffers
  CUtensorMap tm_qkv_f, tm_qkv_g, tm_do_f, tm_do_g;
  if (make_tmap_bf16_2d(&tm_qkv_f, qkv, 3ULL * C, rows, d.ld_qkv, TC_HD, L)) return -1;
  if (make_tmap_bf16_2d(&tm_qkv_g, qkv, 3ULL * C, rows, d.ld_qkv, TC_HD, M)) return -1;
  if (make_tmap_bf16_2d(&tm_do_f, dout, C, rows, d.ld_o, TC_HD, L)) return -1;
  if (make_tmap_bf16_2d(&tm_do_g, dout, C, rows, d.ld_o, TC_HD, M)) return -1;
  vip_attn_bwd_tc_kernel<<<grid, TB_THREADS, smem, st>>>(
      tm_qkv_f, tm_qkv_g, tm_do_f, tm_do_g, lse, delta, static_cast<__nv_bfloat16*>(dqkv), workspace, d, q_scale);
  XP_CHECK_LAUNCH("vip_attn_bwd_tc_kernel");
  return 0;
}

