// tcgen05 / TMEM implementation of the video-proxy (ViP) attention of CLIP-ViP (CLIPAttention.forward2,
// CLIP_ViP.py:332-381) — forward.  Persistent, warp-specialised, one CTA per SM looping over (batch, head, frame)
// problems:
//
//   warp 0        TMA producer (one thread): six cp.async.bulk.tensor box loads stage the problem's q / k / v rows in shared
//                 memory in the UMMA SWIZZLE_128B layout (double buffered: problem n+1 loads while problem n computes)
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
#include <cstdlib>
#include <cstdio>

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

// part: [B, H, T, M, 66] fp32 = {max, sum, unnormalised out[64]} of the global queries over this frame's keys.
struct TfMaps {
  CUtensorMap qkv_f, qkv_g;     // boxes of 64 columns x {L frame rows, M global rows} over qkv [rows, 3C]
};

__global__ void __launch_bounds__(TC_THREADS, 1)
vip_attn_fwd_tc_kernel(const __grid_constant__ TfMaps tm, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
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
      mbar_init(&full[i], 1);
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

  if (warp == 0) {
    // ------------------------------------------- TMA producer (one thread): six box loads per problem — the L frame rows and
    // the M global rows of the Q', K', V' head slices — land in the 128B-swizzled tiles the UMMA descriptors read
    if (lane == 0) {
      uint8_t* gbuf0 = smem_raw + (buf0 - raw);
      int n = 0;
      for (int prob = blockIdx.x; prob < total; prob += gridDim.x, ++n) {
        const int s = n & 1;
        const int t = prob % d.T, h = (prob / d.T) % d.H, b = prob / (d.T * d.H);
        mbar_wait(&empty[s], ((n >> 1) & 1) ^ 1);
        uint8_t* sQ = gbuf0 + s * TC_BUF_BYTES;
        uint8_t* sK = sQ + TC_QROWS * 128;
        uint8_t* sV = sK + TC_FK * 128;
        uint8_t* sKg = sV + TC_FK * 128;
        uint8_t* sVg = sKg + TC_GK * 128;
        const int rg = static_cast<int>(static_cast<long long>(b) * d.S), rf = rg + d.M + t * d.L;
        const int cq = h * TC_HD, ck = d.C + h * TC_HD, cv = 2 * d.C + h * TC_HD;
        mbar_arrive_expect_tx(&full[s], 3u * static_cast<uint32_t>(d.L + d.M) * 128u);
        tma_load_2d(sQ, &tm.qkv_f, &full[s], cq, rf);
        tma_load_2d(sK, &tm.qkv_f, &full[s], ck, rf);
        tma_load_2d(sQ + TC_GROW * 128, &tm.qkv_g, &full[s], cq, rg);      // (TC_GROW & 7) == 0: on a swizzle atom
        tma_load_2d(sKg, &tm.qkv_g, &full[s], ck, rg);
        tma_load_2d(sV, &tm.qkv_f, &full[s], cv, rf);
        tma_load_2d(sVg, &tm.qkv_g, &full[s], cv, rg);
      }
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
    const uint32_t stage = buf0 + 2 * TC_BUF_BYTES + 1024 + static_cast<uint32_t>(warp - 4) * 4096u;   // output staging, 4 KB per warp
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
        }
        if (warp_active) {      // bf16 rows into this warp's 4 KB staging slot (dead rows: finite garbage, never stored)
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t v0 = pack_bf16(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
            const uint32_t v1 = pack_bf16(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
            const uint32_t v2 = pack_bf16(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
            const uint32_t v3 = pack_bf16(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sw128(stage, lane, half * 4 + q)), "r"(v0), "r"(v1),
                         "r"(v2), "r"(v3) : "memory");
          }
        }
      }
      // coalesced output: 8 lanes cover one 128-byte row, an instruction writes 4 full lines (a lane-per-row store would touch
      // 32 lines per instruction and serialise in the LSU)
      if (warp_active) {
        __syncwarp();
        const int c = lane & 7;
        __nv_bfloat16* obase = out + (static_cast<long long>(b) * d.S + d.M + static_cast<long long>(t) * d.L) * d.ld_o + h * TC_HD + c * 8;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + (lane >> 3);
          const int frow = j * 128 + wq * 32 + rr;              // row within sQ; frame rows are [0, L)
          uint4 v;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sw128(stage, rr, c)) : "memory");
          if (frow < d.L) *reinterpret_cast<uint4*>(obase + static_cast<long long>(frow) * d.ld_o) = v;
        }
        __syncwarp();
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
  const int smem = 2 * TC_BUF_BYTES + 1024 + 1024 + 8 * 4096;     // two operand buffers, barriers, alignment slack, output staging
  static bool attr = false;
  if (!attr) {
    XP_CHECK_CUDA(cudaFuncSetAttribute(vip_attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  const long long total = static_cast<long long>(B) * H * T;
  const int grid = static_cast<int>(total < sm_count() ? total : sm_count());
  const long long rows = static_cast<long long>(B) * d.S;
  if (rows >= (1LL << 31)) return fail("vip_attention: more than 2^31 token rows");
  TfMaps tm;
  if (make_tmap_bf16_2d(&tm.qkv_f, qkv, 3ULL * C, rows, d.ld_qkv, TC_HD, L)) return -1;
  if (make_tmap_bf16_2d(&tm.qkv_g, qkv, 3ULL * C, rows, d.ld_qkv, TC_HD, M)) return -1;
  vip_attn_fwd_tc_kernel<<<grid, TC_THREADS, smem, static_cast<cudaStream_t>(stream)>>>(
      tm, static_cast<__nv_bfloat16*>(out), lse, workspace, d);
  XP_CHECK_LAUNCH("vip_attn_fwd_tc_kernel");
  return 0;
}

// =====================================================================================================
// Backward (round 2: software-pipelined).  Persistent, one CTA per SM, 18 warps:
//   warp 0        TMA producer (one thread): per item, five independently released operand groups —
//                 K' (+global keys; DOUBLE-buffered across items), Q'/dO' rows [0,128) ("q0"), Q'/dO' rows [128,L) + the
//                 M global rows ("q1"), V' keys [0,128) ("v0"), V' keys [128,L) + global ("v1") — so that the next item's
//                 first step is already staged while the current item's last steps run
//   warp 1        MMA issuer A (one thread): S, dP, dV; also allocates / frees TMEM
//   warp 18       MMA issuer B (one thread): dK, dQ — the N = 64, K = 16 gradient MMAs take ~32 tensor cycles each but ~45
//                 cycles to issue from one thread (descriptor moves into uniform registers), so the issue is split in two
//   warps 2-17    four math warpgroups; thread = query row (TMEM lane = 32 (warp % 4) + lane) of the current 128-row query
//                 tile, warpgroup g = (warp - 2) / 4 owns key columns [32g, 32g+32) of the step
// An item = (batch b, head h, frame t) = 2 key tiles x 2 query tiles = 4 steps of [128 q x 128 keys] (s = 2 i + j):
//   S  = Q_j K_i^T, dP = dO_j V_i^T                (tcgen05.mma SS -> TMEM columns [0,128), [128,256))
//   phase 1: P = exp(S - lse) kept as packed bf16 in REGISTERS -> the S columns are released at once, so S of step
//            n+1 is computed while phase 2 of step n runs
//   phase 2: dS = P * (dP - delta); P and dS (bf16) written to shared memory, SWIZZLE_128B
//   G:  dV_i += P^T dO_j, dK_i += dS^T Q_j (A MN-major from smem; TMEM [256,384)); dQ_j += dS K_i (TMEM [384,512))
// Issue order per step n: S_n | dP_n | G_(n-1): the tensor pipe works on dP_n and the three gradient products of the
// previous step while the math warps are in phase 1 of step n, and on S_(n+1) during their phase 2.  Accumulators are
// drained (dV_i/dK_i after the key tile's second step, dQ after the item) by the math warps inside the NEXT step, between
// its two phases, when the products they wait for have retired anyway.
// Key tile 0 = frame keys [0,128); key tile 1 = frame keys [128,208) in columns [0,80) + the M global keys in columns
// [80,96).  Query tile 0 = frame rows [0,128); tile 1 = frame rows [128,196) + global rows at 200..
// delta_i = sum_d dO_id O_id is precomputed (vip_attn_delta_kernel).  The gradients of the M global rows are emitted as
// per-frame fp32 partials and reduced by vip_attn_bwd_combine (vip_attention.cu).
constexpr int TB_MATH_WARPS = 16;
constexpr int TB_THREADS = (3 + TB_MATH_WARPS) * 32;       // 608: producer, two MMA issuers, 16 math warps
constexpr int TB_KBUF = (TC_FK + TC_GK) * 128;             // one K' or V' buffer: [208 frame + 16 global rows][128 B]
constexpr int TB_SQ = 0;                                   // [256][128 B]
constexpr int TB_SDO = TB_SQ + 256 * 128;                  // [256][128 B]
constexpr int TB_SK = TB_SDO + 256 * 128;                  // 2 x TB_KBUF
constexpr int TB_SV = TB_SK + 2 * TB_KBUF;
constexpr int TB_SP = TB_SV + TB_KBUF;                     // [2 atoms][128 rows][128 B]
constexpr int TB_SDS = TB_SP + 2 * 128 * 128;
constexpr int TB_BAR = TB_SDS + 2 * 128 * 128;
constexpr int TB_SMEM = TB_BAR + 256;
static_assert(TB_KBUF % 1024 == 0 && TB_SMEM + 1024 <= 232448, "shared-memory plan");

enum TbBar { Q0_FULL = 0, Q1_FULL, V0_FULL, V1_FULL, K_FULL, K_FULL1, Q0_FREE, Q1_FREE, V0_FREE, V1_FREE, K_FREE, K_FREE1,
             S_READY, S_FREE, DP_READY, PDS_READY, G_DONE, TB_NBAR };

struct TbMaps {
  CUtensorMap qkv_a, qkv_b, qkv_l, qkv_g;   // boxes of 64 columns x {min(L,128), L-128, L, M} rows over qkv [rows, 3C]
  CUtensorMap do_a, do_b, do_g;             // same row boxes over dO [rows, C]
};

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

// live-key bit mask of the 32 columns [32 wg, 32 wg + 32) of key tile i (bit e = column 32 wg + e)
__device__ __forceinline__ uint32_t tb_col_mask(int i, int wg, int L, int M, bool glob_keys_masked) {
  const int c0 = wg * 32;
  if (i == 0) {
    const int n = L - c0;
    return n >= 32 ? 0xffffffffu : (n <= 0 ? 0u : ((1u << n) - 1u));
  }
  const int nf = min(L - 128, 80) - c0;                      // live frame keys from column c0 on
  uint32_t m = nf >= 32 ? 0xffffffffu : (nf <= 0 ? 0u : ((1u << nf) - 1u));
  if (!glob_keys_masked) {
    const int g0 = 80 - c0;                                  // bit of the first global key
    if (g0 >= 0 && g0 < 32) m |= ((1u << M) - 1u) << g0;     // M <= 8 and 80 % 32 == 16: never crosses the word
  }
  return m;
}

__global__ void __launch_bounds__(TB_THREADS, 1)
vip_attn_bwd_tc_kernel(const __grid_constant__ TbMaps tm, const float* __restrict__ lse, const float* __restrict__ delta,
                       __nv_bfloat16* __restrict__ dqkv, float* __restrict__ gpart, const TcDims d, float q_scale,
                       const int dbg, long long* __restrict__ trace) {
  // dbg (profiling only, XP_ATTN_BWD_DEBUG): bit 0 = issue no MMAs (commits only), bit 1 = math warps run the barrier
  // protocol without their loads / exps / stores — isolates the tensor-pipe time from the math time; bit 2 = CTA 0 records
  // clock64() time stamps of its first 24 steps (issuer: 4 per step, math warp 2: 5 per step) into `trace`
#define TB_TRACE(role, k)                                                                              \
  do {                                                                                                 \
    if ((dbg & 4) && blockIdx.x == 0 && step < 24) trace[(role) * 24 * 8 + step * 8 + (k)] = clock64(); \
  } while (0)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - raw);
  uint64_t* bar = reinterpret_cast<uint64_t*>(gbase + TB_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + TB_NBAR);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = d.B * d.H * d.T;
  const int L1 = d.L < 128 ? d.L : 128, L2 = d.L - L1;

  if (tid == 0) {
    for (int i = 0; i < TB_NBAR; ++i)
      mbar_init(&bar[i], (i == S_FREE || i == PDS_READY) ? TB_MATH_WARPS
                         : (i == G_DONE || i == Q0_FREE || i == Q1_FREE || i == K_FREE || i == K_FREE1) ? 2 : 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  for (int i = tid; i < TB_BAR / 16; i += TB_THREADS)   // zero all operand tiles once (padding rows stay zero)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + i * 16), "r"(0) : "memory");
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 320,
                 tdQ = tmem_base + 384;

  if (warp == 0) {
    // ----------------------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int n = 0;
      for (int prob = blockIdx.x; prob < total; prob += gridDim.x, ++n) {
        const int t = prob % d.T, h = (prob / d.T) % d.H, b = prob / (d.T * d.H);
        const int kb = n & 1;
        const uint32_t pi = n & 1, pk = (n >> 1) & 1;
        const int rg = static_cast<int>(static_cast<long long>(b) * d.S);
        const int rf = rg + d.M + t * d.L;
        const int cq = h * TC_HD, ck = d.C + h * TC_HD, cv = 2 * d.C + h * TC_HD;
        uint8_t* sK = gbase + TB_SK + kb * TB_KBUF;
        mbar_wait(&bar[K_FREE + kb], pk ^ 1);
        mbar_arrive_expect_tx(&bar[K_FULL + kb], static_cast<uint32_t>(d.L + d.M) * 128u);
        tma_load_2d(sK, &tm.qkv_l, &bar[K_FULL + kb], ck, rf);
        tma_load_2d(sK + TC_FK * 128, &tm.qkv_g, &bar[K_FULL + kb], ck, rg);
        mbar_wait(&bar[Q0_FREE], pi ^ 1);
        mbar_arrive_expect_tx(&bar[Q0_FULL], 2u * static_cast<uint32_t>(L1) * 128u);
        tma_load_2d(gbase + TB_SQ, &tm.qkv_a, &bar[Q0_FULL], cq, rf);
        tma_load_2d(gbase + TB_SDO, &tm.do_a, &bar[Q0_FULL], cq, rf);
        mbar_wait(&bar[V0_FREE], pi ^ 1);
        mbar_arrive_expect_tx(&bar[V0_FULL], static_cast<uint32_t>(L1) * 128u);
        tma_load_2d(gbase + TB_SV, &tm.qkv_a, &bar[V0_FULL], cv, rf);
        mbar_wait(&bar[Q1_FREE], pi ^ 1);
        mbar_arrive_expect_tx(&bar[Q1_FULL], 2u * static_cast<uint32_t>(L2 + d.M) * 128u);
        if (L2 > 0) {
          tma_load_2d(gbase + TB_SQ + 128 * 128, &tm.qkv_b, &bar[Q1_FULL], cq, rf + 128);
          tma_load_2d(gbase + TB_SDO + 128 * 128, &tm.do_b, &bar[Q1_FULL], cq, rf + 128);
        }
        tma_load_2d(gbase + TB_SQ + TC_GROW * 128, &tm.qkv_g, &bar[Q1_FULL], cq, rg);
        tma_load_2d(gbase + TB_SDO + TC_GROW * 128, &tm.do_g, &bar[Q1_FULL], cq, rg);
        mbar_wait(&bar[V1_FREE], pi ^ 1);
        mbar_arrive_expect_tx(&bar[V1_FULL], static_cast<uint32_t>(L2 + d.M) * 128u);
        if (L2 > 0) tma_load_2d(gbase + TB_SV + 128 * 128, &tm.qkv_b, &bar[V1_FULL], cv, rf + 128);
        tma_load_2d(gbase + TB_SV + TC_FK * 128, &tm.qkv_g, &bar[V1_FULL], cv, rg);
      }
    }
  } else if (warp == 1 || warp == 18) {
    // ---------------------------------------------------------------------------------- MMA issuers A (warp 1) / B (warp 18)
    if (lane == 0) {
      constexpr uint32_t id_s128 = make_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t id_s80 = make_idesc_bf16(128, 80, 0, 0);
      constexpr uint32_t id_s16 = make_idesc_bf16(128, 16, 0, 0);
      constexpr uint32_t id_kv = make_idesc_bf16(128, TC_HD, 1, 1);   // A = P / dS MN-major, B = dO / Q MN-major
      constexpr uint32_t id_q = make_idesc_bf16(128, TC_HD, 0, 1);    // A = dS K-major, B = K MN-major
      // One thread issues ~60 MMAs per step: descriptors are built ONCE (a descriptor of the same layout at +x bytes is the
      // base descriptor + (x >> 4): the 14-bit start-address field never overflows inside the 227 KB of shared memory), so
      // an MMA costs one 64-bit add per operand instead of a shift / mask / or chain.
      auto at = [](uint64_t desc, uint32_t byte_off) { return desc + (byte_off >> 4); };
      const uint64_t q_k = make_smem_desc_sw128(base + TB_SQ, 16, 1024), o_k = make_smem_desc_sw128(base + TB_SDO, 16, 1024);
      const uint64_t q_m = make_smem_desc_sw128(base + TB_SQ, 16384, 1024), o_m = make_smem_desc_sw128(base + TB_SDO, 16384, 1024);
      const uint64_t p_m = make_smem_desc_sw128(base + TB_SP, 16384, 1024), s_m = make_smem_desc_sw128(base + TB_SDS, 16384, 1024);
      const uint64_t s_k = make_smem_desc_sw128(base + TB_SDS, 16, 1024);
      const uint64_t v_k = make_smem_desc_sw128(base + TB_SV, 16, 1024);
      const uint64_t k_k0 = make_smem_desc_sw128(base + TB_SK, 16, 1024), k_m0 = make_smem_desc_sw128(base + TB_SK, 16384, 1024);
      // the three gradient products of step (i, j) with K' in buffer kb
      auto issue_grads = [&](int i, int j, int kb) {
        if (dbg & 1) return;
        const uint64_t k_m = at(k_m0, kb * TB_KBUF);
        const uint64_t ob = at(o_m, j * 16384), qb = at(q_m, j * 16384);
        if (warp == 1) {                   // issuer A: dV_i += P^T dO_j
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)   // K = the 128 query rows of tile j
            umma_bf16(tdV, at(p_m, ks * 2048), at(ob, ks * 2048), id_kv, (j > 0 || ks > 0) ? 1u : 0u);
          return;
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)     // issuer B: dK_i += dS^T Q_j, then dQ_j += dS K_i
          umma_bf16(tdK, at(s_m, ks * 2048), at(qb, ks * 2048), id_kv, (j > 0 || ks > 0) ? 1u : 0u);
        const uint32_t tq = tdQ + j * 64;
        if (i == 0) {
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)   // K = frame keys [0,128): atom ks/4, 32 B per k-step inside it
            umma_bf16(tq, at(s_k, (ks >> 2) * 16384 + (ks & 3) * 32), at(k_m, ks * 2048), id_q, ks > 0 ? 1u : 0u);
        } else {
#pragma unroll
          for (int ks = 0; ks < 5; ++ks)   // frame keys [128,208)
            umma_bf16(tq, at(s_k, (ks >> 2) * 16384 + (ks & 3) * 32), at(k_m, (128 + ks * 16) * 128), id_q, 1u);
          umma_bf16(tq, at(s_k, 16384 + 32), at(k_m, TC_FK * 128), id_q, 1u);   // columns [80,96): the global keys
        }
      };
      const bool is_a = warp == 1;
      int n = 0;
      uint32_t step = 0;
      for (int prob = blockIdx.x; prob < total; prob += gridDim.x, ++n) {
        const int kb = n & 1;
        const uint32_t pi = n & 1, pk = (n >> 1) & 1;
        const uint64_t k_k = at(k_k0, kb * TB_KBUF);
#pragma unroll
        for (int s = 0; s < 4; ++s, ++step) {
          const int i = s >> 1, j = s & 1;
          // ---- S_n = Q_j K_i^T                                                              (issuer A)
          if (is_a) {
            if (s == 0) {
              mbar_wait(&bar[Q0_FULL], pi);
              mbar_wait(&bar[K_FULL + kb], pk);
            } else if (s == 1) {
              mbar_wait(&bar[Q1_FULL], pi);
            }
            if (step > 0) mbar_wait(&bar[S_FREE], (step - 1) & 1);    // phase 1 of the previous step has read its S
            tc_fence_after();
            TB_TRACE(0, 0);
            if (!(dbg & 1)) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t aq = at(q_k, j * 16384 + ks * 32);
                const uint32_t acc = ks > 0 ? 1u : 0u;
                if (i == 0) {
                  umma_bf16(tS, aq, at(k_k, ks * 32), id_s128, acc);
                } else {
                  umma_bf16(tS, aq, at(k_k, 128 * 128 + ks * 32), id_s80, acc);
                  umma_bf16(tS + 80, aq, at(k_k, TC_FK * 128 + ks * 32), id_s16, acc);
                }
              }
            }
            umma_commit(&bar[S_READY]);
            TB_TRACE(0, 1);
            if (s == 0) mbar_wait(&bar[V0_FULL], pi);
            else if (s == 2) mbar_wait(&bar[V1_FULL], pi);
          }
          if (step > 0) {
            mbar_wait(&bar[PDS_READY], (step - 1) & 1);               // dP of the previous step consumed; its P, dS staged
            fence_proxy_async_smem();
          }
          tc_fence_after();
          // ---- dP_n = dO_j V_i^T                                                             (issuer A)
          if (is_a) {
            TB_TRACE(0, 2);
            if (!(dbg & 1)) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t ao = at(o_k, j * 16384 + ks * 32);
                const uint32_t acc = ks > 0 ? 1u : 0u;
                if (i == 0) {
                  umma_bf16(tdP, ao, at(v_k, ks * 32), id_s128, acc);
                } else {
                  umma_bf16(tdP, ao, at(v_k, 128 * 128 + ks * 32), id_s80, acc);
                  umma_bf16(tdP + 80, ao, at(v_k, TC_FK * 128 + ks * 32), id_s16, acc);
                }
              }
            }
            umma_commit(&bar[DP_READY]);
            if (s == 1) umma_commit(&bar[V0_FREE]);
            else if (s == 3) umma_commit(&bar[V1_FREE]);
          }
          // ---- gradient products of the previous step: dV (A) | dK, dQ (B); both commit on G_DONE and on the operand frees
          if (step > 0) {
            const int ps = (s + 3) & 3, pkb = s == 0 ? (kb ^ 1) : kb;
            issue_grads(ps >> 1, ps & 1, pkb);
            umma_commit(&bar[G_DONE]);
            if (ps == 2) umma_commit(&bar[Q0_FREE]);
            else if (ps == 3) {
              umma_commit(&bar[Q1_FREE]);
              umma_commit(&bar[K_FREE + pkb]);
            }
          }
          if (is_a) TB_TRACE(0, 3);
        }
      }
      if (step > 0) {   // the last step's gradient products
        mbar_wait(&bar[PDS_READY], (step - 1) & 1);
        fence_proxy_async_smem();
        tc_fence_after();
        issue_grads(1, 1, (n - 1) & 1);
        umma_commit(&bar[G_DONE]);
      }
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------------------- math warpgroups
    const int wg = (warp - 2) >> 2;               // key columns [32 wg, 32 wg + 32) of a step
    const int wq = warp & 3;                       // TMEM lane quarter
    const int trow = wq * 32 + lane;               // TMEM lane == row of the current query tile
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const uint32_t sPw = base + TB_SP + (wg >> 1) * 16384, sdSw = base + TB_SDS + (wg >> 1) * 16384;
    const int chunk0 = (wg & 1) * 4;               // first 16-byte chunk of this warpgroup inside its 64-column atom
    // rows of the two query tiles this thread owns, and whether its warp has any live row there
    bool warp_live[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r0 = j * 128 + wq * 32;
      warp_live[j] = (r0 < d.L) || (r0 + 31 >= TC_GROW && r0 < TC_GROW + d.M);
    }
    auto load_stats = [&](int prob, float (&l2)[2], float (&dl)[2]) {
      const int t = prob % d.T, h = (prob / d.T) % d.H, b = prob / (d.T * d.H);
      const long long sb = (static_cast<long long>(b) * d.H + h) * d.S;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int row = j * 128 + trow;
        long long idx = -1;
        if (row < d.L) idx = sb + d.M + static_cast<long long>(t) * d.L + row;
        else if (row >= TC_GROW && row < TC_GROW + d.M) idx = sb + (row - TC_GROW);
        l2[j] = idx >= 0 ? lse[idx] * TC_LOG2E : INFINITY;     // +inf on padding rows -> P = 0
        dl[j] = idx >= 0 ? delta[idx] : 0.f;
      }
    };
    uint32_t step = 0;
    // Accumulator drain.  One warp moves 32 rows x 64 columns of ONE accumulator (acc 0: dV_i, 1: dK_i, 2: dQ_0, 3: dQ_1; rows =
    // its TMEM lane quarter): TMEM -> registers -> bf16 -> its 4 KB slot of the P / dS tiles (free between G_DONE of the previous
    // step and this step's phase 2b) -> global stores in which 8 lanes cover one 128-byte row segment.  A lane-per-row store
    // touches 32 different lines per instruction and cost ~5 k cycles per drain (profiles/r02_attn_bwd_trace.md); this way an
    // instruction writes 4 full lines.  The M global rows (fp32 per-frame partials) are written by their owning lanes.
    auto drain = [&](int acc, int i, int prob) {
      const int t = prob % d.T, h = (prob / d.T) % d.H, b = prob / (d.T * d.H);
      const long long tok_g = static_cast<long long>(b) * d.S, tok_f = tok_g + d.M + static_cast<long long>(t) * d.L;
      // accumulator row ar (0..127) -> sequence row of the frame (or -1), and global-token index (or -1)
      auto frame_row = [&](int ar) {
        if (acc >= 2) { const int row = (acc - 2) * 128 + ar; return row < d.L ? row : -1; }
        if (i == 0) return ar < d.L ? ar : -1;
        return (ar < 80 && 128 + ar < d.L) ? 128 + ar : -1;
      };
      auto glob_row = [&](int ar) {
        if (acc >= 2) { const int row = (acc - 2) * 128 + ar; return (row >= TC_GROW && row < TC_GROW + d.M) ? row - TC_GROW : -1; }
        return (i == 1 && ar >= 80 && ar < 80 + d.M) ? ar - 80 : -1;
      };
      const int ar0 = wq * 32;
      // this warp's 32 rows hold a live frame row iff the first one is live (live rows are a prefix), or a global row
      const bool any = frame_row(ar0) >= 0 ||
                       (acc >= 2 ? (acc == 3 && ar0 + 31 >= TC_GROW - 128 && ar0 < TC_GROW - 128 + d.M)
                                 : (i == 1 && ar0 + 31 >= 80 && ar0 < 80 + d.M));
      if (!any) return;
      const uint32_t tsrc = (acc == 0 ? tdV : acc == 1 ? tdK : tdQ + (acc - 2) * 64) + lane_off;
      const int sect = acc == 0 ? 2 : (acc == 1 ? 1 : 0);            // [q | k | v] section of dqkv / of the partials
      const float scale = acc >= 2 ? q_scale : 1.f;
      const uint32_t slot = base + TB_SP + static_cast<uint32_t>(warp - 2) * 4096u;
      const int my_g = glob_row(ar0 + lane);
      float* grow = gpart + (((static_cast<long long>(b) * d.H + h) * d.T + t) * d.M + my_g) * 3 * TC_HD + sect * TC_HD;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t o[32];
        tmem_ld32(tsrc + half * 32, o);
        tmem_ld_wait(o);
        if (my_g >= 0) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(grow + half * 32 + q * 4) = make_float4(__uint_as_float(o[q * 4]), __uint_as_float(o[q * 4 + 1]),
                                                                                __uint_as_float(o[q * 4 + 2]), __uint_as_float(o[q * 4 + 3]));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t v0 = pack_bf16(__uint_as_float(o[q * 8 + 0]) * scale, __uint_as_float(o[q * 8 + 1]) * scale);
          const uint32_t v1 = pack_bf16(__uint_as_float(o[q * 8 + 2]) * scale, __uint_as_float(o[q * 8 + 3]) * scale);
          const uint32_t v2 = pack_bf16(__uint_as_float(o[q * 8 + 4]) * scale, __uint_as_float(o[q * 8 + 5]) * scale);
          const uint32_t v3 = pack_bf16(__uint_as_float(o[q * 8 + 6]) * scale, __uint_as_float(o[q * 8 + 7]) * scale);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sw128(slot, lane, half * 4 + q)), "r"(v0), "r"(v1),
                       "r"(v2), "r"(v3) : "memory");
        }
      }
      __syncwarp();
      if (warp == 2 && lane == 0) TB_TRACE(1, 6);
      const int c = lane & 7;
      __nv_bfloat16* gcol = dqkv + sect * d.C + h * TC_HD + c * 8;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + (lane >> 3);                      // row inside this warp's 32
        const int fr = frame_row(ar0 + rr);
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(sw128(slot, rr, c)) : "memory");
        if (fr >= 0) *reinterpret_cast<uint4*>(gcol + (tok_f + fr) * d.ld_qkv) = v;
      }
      if (warp == 2 && lane == 0) TB_TRACE(1, 7);
    };
    // 512 math threads: the staging slots alias the P / dS tiles, so nobody may start phase 2b before every drain has been read back
    auto drain_sync = [] { asm volatile("bar.sync 1, 512;" ::: "memory"); };

    float l2n[2], dln[2];
    if (static_cast<int>(blockIdx.x) < total) load_stats(blockIdx.x, l2n, dln);
    int prev_prob = -1;
    for (int prob = blockIdx.x; prob < total; prob += gridDim.x) {
      const int t = prob % d.T;
      const float l2[2] = {l2n[0], l2n[1]}, dl[2] = {dln[0], dln[1]};
      if (prob + static_cast<int>(gridDim.x) < total) load_stats(prob + gridDim.x, l2n, dln);   // prefetch the next item's rows
      for (int s = 0; s < 4; ++s, ++step) {
        const int i = s >> 1, j = s & 1;
        const int row = j * 128 + trow;
        const bool q_glob = row >= TC_GROW && row < TC_GROW + d.M;
        const bool cols_on = i == 0 ? (wg * 32 < d.L) : (wg < 3);
        const bool active = (j ? warp_live[1] : warp_live[0]) && cols_on && !(dbg & 2);
        const uint32_t mask = tb_col_mask(i, wg, d.L, d.M, q_glob && t != 0);
        // ---- phase 1: P = exp(S - lse) -> packed bf16 in registers
        uint32_t ppk[16];
        mbar_wait(&bar[S_READY], step & 1);
        tc_fence_after();
        if (warp == 2 && lane == 0) TB_TRACE(1, 0);
        if (active) {
          uint32_t r[2][16];
          tmem_ld16(tS + lane_off + wg * 32, r[0]);
          tmem_ld16(tS + lane_off + wg * 32 + 16, r[1]);
          tmem_ld_wait16(r[0]);
          tmem_ld_wait16(r[1]);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar[S_FREE]);     // S is in registers: the tensor pipe may overwrite it
          const float nl = -(j ? l2[1] : l2[0]);
          if (mask == 0xffffffffu) {
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
              for (int e = 0; e < 16; e += 2)
                ppk[c * 8 + (e >> 1)] = pack_bf16(tc_exp2(fmaf(__uint_as_float(r[c][e]), TC_LOG2E, nl)),
                                                  tc_exp2(fmaf(__uint_as_float(r[c][e + 1]), TC_LOG2E, nl)));
          } else {
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
              for (int e = 0; e < 16; e += 2) {
                const float p0 = (mask >> (c * 16 + e)) & 1u ? tc_exp2(fmaf(__uint_as_float(r[c][e]), TC_LOG2E, nl)) : 0.f;
                const float p1 = (mask >> (c * 16 + e + 1)) & 1u ? tc_exp2(fmaf(__uint_as_float(r[c][e + 1]), TC_LOG2E, nl)) : 0.f;
                ppk[c * 8 + (e >> 1)] = pack_bf16(p0, p1);
              }
          }
        } else {
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar[S_FREE]);
        }
        // ---- phase 2a: dS = P * (dP - delta) into registers.  dP of this step was issued before the previous step's gradient
        //      products, so it is ready long before they retire: everything but the shared-memory stores stays off the
        //      G_DONE -> PDS_READY critical chain
        uint32_t dkp[16];
        if (warp == 2 && lane == 0) TB_TRACE(1, 1);
        mbar_wait(&bar[DP_READY], step & 1);
        tc_fence_after();
        if (warp == 2 && lane == 0) TB_TRACE(1, 2);
        if (active) {
          uint32_t r[2][16];
          tmem_ld16(tdP + lane_off + wg * 32, r[0]);
          tmem_ld16(tdP + lane_off + wg * 32 + 16, r[1]);
          tmem_ld_wait16(r[0]);
          tmem_ld_wait16(r[1]);
          const float dlj = j ? dl[1] : dl[0];
#pragma unroll
          for (int c = 0; c < 2; ++c)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const uint32_t pw = ppk[c * 8 + e];
              dkp[c * 8 + e] = pack_bf16(bf16_lo(pw) * (__uint_as_float(r[c][2 * e]) - dlj),
                                         bf16_hi(pw) * (__uint_as_float(r[c][2 * e + 1]) - dlj));
            }
        }
        // ---- the previous step's gradient products have retired: its P / dS tiles are free, its accumulators final
        if (warp == 2 && lane == 0) TB_TRACE(1, 3);
        if (step > 0) {
          mbar_wait(&bar[G_DONE], (step - 1) & 1);
          tc_fence_after();
          if (warp == 2 && lane == 0) TB_TRACE(1, 4);
          if (s == 2) {                         // key tile 0 of this item is complete: dV_0 / dK_0
            if (wg < 2) drain(wg, 0, prob);
            drain_sync();
          } else if (s == 0) {                  // the previous item is complete: dV_1 / dK_1 / dQ_0 / dQ_1, one per warpgroup
            drain(wg, 1, prev_prob);
            drain_sync();
          }
        }
        // ---- phase 2b: stage P and dS for this step's gradient products
        if (active) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int ch = chunk0 + c * 2;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sw128(sPw, trow, ch)), "r"(ppk[c * 8 + 0]),
                         "r"(ppk[c * 8 + 1]), "r"(ppk[c * 8 + 2]), "r"(ppk[c * 8 + 3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sw128(sPw, trow, ch + 1)), "r"(ppk[c * 8 + 4]),
                         "r"(ppk[c * 8 + 5]), "r"(ppk[c * 8 + 6]), "r"(ppk[c * 8 + 7]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sw128(sdSw, trow, ch)), "r"(dkp[c * 8 + 0]),
                         "r"(dkp[c * 8 + 1]), "r"(dkp[c * 8 + 2]), "r"(dkp[c * 8 + 3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sw128(sdSw, trow, ch + 1)), "r"(dkp[c * 8 + 4]),
                         "r"(dkp[c * 8 + 5]), "r"(dkp[c * 8 + 6]), "r"(dkp[c * 8 + 7]) : "memory");
          }
          fence_proxy_async_smem();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[PDS_READY]);
        if (warp == 2 && lane == 0) TB_TRACE(1, 5);
      }
      prev_prob = prob;
    }
    if (step > 0) {   // accumulators of the last item
      mbar_wait(&bar[G_DONE], (step - 1) & 1);
      tc_fence_after();
      drain(wg, 1, prev_prob);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
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
  if (rows >= (1LL << 31)) return fail("vip_attention: more than 2^31 token rows");
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
  // TMA boxes: 64 columns (one head slice) x {first 128, remaining, all L, M global} rows of the token-major buffers
  const int L1 = L < 128 ? L : 128, L2 = L - L1;
  TbMaps tm;
  if (make_tmap_bf16_2d(&tm.qkv_a, qkv, 3ULL * C, rows, d.ld_qkv, TC_HD, L1)) return -1;
  if (make_tmap_bf16_2d(&tm.qkv_b, qkv, 3ULL * C, rows, d.ld_qkv, TC_HD, L2 > 0 ? L2 : 1)) return -1;
  if (make_tmap_bf16_2d(&tm.qkv_l, qkv, 3ULL * C, rows, d.ld_qkv, TC_HD, L)) return -1;
  if (make_tmap_bf16_2d(&tm.qkv_g, qkv, 3ULL * C, rows, d.ld_qkv, TC_HD, M)) return -1;
  if (make_tmap_bf16_2d(&tm.do_a, dout, C, rows, d.ld_o, TC_HD, L1)) return -1;
  if (make_tmap_bf16_2d(&tm.do_b, dout, C, rows, d.ld_o, TC_HD, L2 > 0 ? L2 : 1)) return -1;
  if (make_tmap_bf16_2d(&tm.do_g, dout, C, rows, d.ld_o, TC_HD, M)) return -1;
  static const int dbg = [] { const char* e = getenv("XP_ATTN_BWD_DEBUG"); return e ? atoi(e) : 0; }();
  static long long* trace = nullptr;
  if ((dbg & 4) && trace == nullptr) XP_CHECK_CUDA(cudaMalloc(&trace, 2 * 24 * 8 * sizeof(long long)));
  if (dbg & 4) XP_CHECK_CUDA(cudaMemsetAsync(trace, 0, 2 * 24 * 8 * sizeof(long long), st));
  vip_attn_bwd_tc_kernel<<<grid, TB_THREADS, smem, st>>>(tm, lse, delta, static_cast<__nv_bfloat16*>(dqkv), workspace, d,
                                                         q_scale, dbg, trace);
  XP_CHECK_LAUNCH("vip_attn_bwd_tc_kernel");
  if (dbg & 4) {   // profiling only: print CTA 0's step timeline of this launch (cycles relative to the first stamp)
    static int printed = 0;
    long long h[2 * 24 * 8];
    XP_CHECK_CUDA(cudaStreamSynchronize(st));
    XP_CHECK_CUDA(cudaMemcpy(h, trace, sizeof(h), cudaMemcpyDeviceToHost));
    if (printed++ == 3) {
      const long long t0 = h[0];
      fprintf(stderr, "attn_bwd trace (cycles from t0): issuer [S_FREE ok, S issued, PDS ok, dP+G issued] | math warp 2 [S_READY ok, P1 done, DP ok, dS done, G_DONE ok, PDS arrive, drain staged, drain stored]\n");
      for (int sidx = 0; sidx < 24; ++sidx) {
        fprintf(stderr, "step %2d: I", sidx);
        for (int k = 0; k < 4; ++k) fprintf(stderr, " %7lld", h[sidx * 8 + k] ? h[sidx * 8 + k] - t0 : -1);
        fprintf(stderr, " | M");
        for (int k = 0; k < 8; ++k) fprintf(stderr, " %7lld", h[24 * 8 + sidx * 8 + k] ? h[24 * 8 + sidx * 8 + k] - t0 : -1);
        fprintf(stderr, "\n");
      }
    }
  }
  return 0;
}

