// Video-proxy (ViP) attention of CLIP-ViP, forward and backward, one CTA per (batch, head, frame).
//
// Reference: CLIPAttention.forward2, CLIP_ViP.py:332-381.  Patch queries of frame t attend to
// [M global keys ; L keys of frame t] (:352-363); the M global queries (cls + video proxies) attend to
// all M + T*L keys (:366-375).  The reference materialises the per-frame K/V with repeat+cat; here a CTA
// stages rows {0..M-1} U {M+t*L .. M+(t+1)*L-1} of the fused qkv buffer once in shared memory and treats
// the global queries as extra query rows: their softmax over all frames is assembled from per-frame
// partials (max, sum, unnormalised output) by a small combine kernel.  q arrives pre-scaled by
// head_dim**-0.5 from the QKV GEMM epilogue (CLIP_ViP.py:341).
//
// Tensor-core path: warp-level mma.sync.m16n8k16 (bf16 -> fp32) with ldmatrix-fed fragments; 200 keys fit
// one CTA so the softmax is a two-block online pass.  (A tcgen05/TMEM version is the planned upgrade; the
// attention is 4 % of the block's FLOPs, the GEMMs around it are tcgen05.)
#include "../../include/xpretrain_b200.h"
#include "common.h"
#include "ptx.cuh"
#include "mma_frag.cuh"

namespace xp {

constexpr int ROWS = 208;       // padded query/key rows per CTA (13 m16 tiles); M + L <= ROWS
constexpr int NTILE = ROWS / 16;
constexpr int ATT_WARPS = 7;
constexpr int ATT_THREADS = ATT_WARPS * 32;

struct AttnDims {
  int B, H, T, L, M;
  long long S;       // M + T*L
  long long ld_qkv;  // 3*C
  long long ld_o;    // C
  int C;
};


__device__ __forceinline__ long long token_row(const AttnDims& d, int b, int t, int i) {
  return static_cast<long long>(b) * d.S + (i < d.M ? i : d.M + static_cast<long long>(t) * d.L + (i - d.M));
}

// Stage the q/k/v rows of (b, h, t) with cp.async; rows >= M+L are zero.
__device__ __forceinline__ void stage_qkv(const __nv_bfloat16* __restrict__ qkv, const AttnDims& d, int b, int h, int t,
                                          uint32_t sQ, uint32_t sK, uint32_t sV) {
  const int nq = d.M + d.L;
  for (int idx = threadIdx.x; idx < 3 * ROWS * 8; idx += ATT_THREADS) {
    const int mat = idx / (ROWS * 8);
    const int rem = idx - mat * (ROWS * 8);
    const int row = rem >> 3, chunk = rem & 7;
    const uint32_t dst = tile_addr(mat == 0 ? sQ : (mat == 1 ? sK : sV), row, chunk);
    if (row < nq) {
      cp_async16(dst, qkv + token_row(d, b, t, row) * d.ld_qkv + static_cast<long long>(mat) * d.C + h * HD + chunk * 8);
    } else {
      st_shared_zero16(dst);
    }
  }
}

// ======================================================================== forward
// grid (T, H, B).  out: [B*S, C] bf16 (frame rows); lse: [B, H, S] fp32 (frame rows);
// part: [B, H, T, M, 66] fp32 = {max, sum, unnormalised out[64]} of the global queries over this frame's keys.
template <int NT>  // n8 tiles in this key block
__device__ __forceinline__ void fwd_key_block(uint32_t sK, uint32_t sV, int key0, int lane, const uint32_t (&qa)[4][4],
                                              float (&o)[8][4], float (&m_run)[2], float (&l_run)[2], int nq, int M,
                                              bool mask_global_keys, int row_lo) {
  float s[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
  for (int np = 0; np < NT / 2; ++np) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t b[4];
      load_b_nk(sK, key0 + np * 16, ks, lane, b);
      mma_bf16(s[2 * np], qa[ks], b[0], b[1]);
      mma_bf16(s[2 * np + 1], qa[ks], b[2], b[3]);
    }
  }
  // masks: padded keys; for global QUERY rows, the global keys count only once (frame 0)
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int i = 0; i < NT; ++i) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int key = key0 + i * 8 + (lane & 3) * 2 + (e & 1);
      const int row = row_lo + (e >> 1) * 8;
      const bool dead = key >= nq || (mask_global_keys && row < M && key < M);
      if (dead) s[i][e] = -INFINITY;
      mx[e >> 1] = fmaxf(mx[e >> 1], s[i][e]);
    }
  }
  float corr[2], m_new[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    m_new[r] = fmaxf(m_run[r], mx[r]);
    corr[r] = (m_new[r] == -INFINITY) ? 1.f : fast_exp2((m_run[r] - m_new[r]) * LOG2E);
    l_run[r] *= corr[r];
    m_run[r] = m_new[r];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    o[i][0] *= corr[0]; o[i][1] *= corr[0];
    o[i][2] *= corr[1]; o[i][3] *= corr[1];
  }
  const float mb[2] = {m_new[0] == -INFINITY ? 0.f : m_new[0] * LOG2E, m_new[1] == -INFINITY ? 0.f : m_new[1] * LOG2E};
#pragma unroll
  for (int i = 0; i < NT; ++i) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float pv = fast_exp2(fmaf(s[i][e], LOG2E, -mb[e >> 1]));  // exp2(-inf) = 0 for masked entries
      s[i][e] = pv;
      l_run[e >> 1] += pv;
    }
  }
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
    pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
    pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int dp = 0; dp < 4; ++dp) {
      uint32_t b[4];
      load_b_kn(sV, key0 + kk * 16, dp, lane, b);
      mma_bf16(o[2 * dp], pa, b[0], b[1]);
      mma_bf16(o[2 * dp + 1], pa, b[2], b[3]);
    }
  }
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
vip_attn_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                    float* __restrict__ part, const AttnDims d) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sQ = (raw + 127u) & ~127u;
  const uint32_t sK = sQ + ROWS * 128, sV = sK + ROWS * 128;
  uint8_t* sQ_ptr = smem_raw + (sQ - raw);
  const int t = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = d.M + d.L;

  stage_qkv(qkv, d, b, h, t, sQ, sK, sV);
  cp_async_wait_all();
  __syncthreads();

  for (int mt = warp; mt < NTILE; mt += ATT_WARPS) {
    if (mt * 16 >= nq) break;
    uint32_t qa[4][4];
    load_a_frags(sQ, mt * 16, lane, qa);
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    const int row_lo = mt * 16 + (lane >> 2);
    fwd_key_block<14>(sK, sV, 0, lane, qa, o, m_run, l_run, nq, d.M, t != 0, row_lo);
    fwd_key_block<12>(sK, sV, 112, lane, qa, o, m_run, l_run, nq, d.M, t != 0, row_lo);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
      l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    }
    // ---- global-query rows: write the per-frame partial (fp32, unnormalised)
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = row_lo + r * 8;
      if (row < d.M) {
        float* p = part + (((static_cast<long long>(b) * d.H + h) * d.T + t) * d.M + row) * 66;
        if ((lane & 3) == 0) {
          p[0] = m_run[r];
          p[1] = l_run[r];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          p[2 + i * 8 + (lane & 3) * 2] = o[i][2 * r];
          p[2 + i * 8 + (lane & 3) * 2 + 1] = o[i][2 * r + 1];
        }
      }
    }
    // ---- frame rows: normalise, stage through this tile's (now dead) Q rows, store 128-byte rows
    const float inv[2] = {l_run[0] > 0.f ? 1.f / l_run[0] : 0.f, l_run[1] > 0.f ? 1.f / l_run[1] : 0.f};
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = mt * 16 + (lane >> 2) + r * 8;
        const uint32_t v = pack_bf16(o[i][2 * r] * inv[r], o[i][2 * r + 1] * inv[r]);
        *reinterpret_cast<uint32_t*>(sQ_ptr + (tile_addr(sQ, row, i) - sQ) + (lane & 3) * 4) = v;
      }
    }
    __syncwarp();
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int idx = lane + it * 32;
      const int row = mt * 16 + (idx >> 3), chunk = idx & 7;
      if (row >= d.M && row < nq) {
        const uint4 v = *reinterpret_cast<const uint4*>(sQ_ptr + (tile_addr(sQ, row, chunk) - sQ));
        *reinterpret_cast<uint4*>(out + token_row(d, b, t, row) * d.ld_o + h * HD + chunk * 8) = v;
      }
    }
    if ((lane & 3) == 0) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = row_lo + r * 8;
        if (row >= d.M && row < nq)
          lse[(static_cast<long long>(b) * d.H + h) * d.S + (token_row(d, b, t, row) - static_cast<long long>(b) * d.S)] =
              m_run[r] + logf(l_run[r]);
      }
    }
  }
}

// Merge the per-frame partials of the M global queries: grid (H, B), block 64 threads (one per head-dim column).
__global__ void __launch_bounds__(64)
vip_attn_fwd_combine_kernel(const float* __restrict__ part, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                            const AttnDims d) {
  const int h = blockIdx.x, b = blockIdx.y, c = threadIdx.x;
  for (int m = 0; m < d.M; ++m) {
    const float* p0 = part + ((static_cast<long long>(b) * d.H + h) * d.T * d.M + m) * 66;
    float mx = -INFINITY;
    for (int t = 0; t < d.T; ++t) mx = fmaxf(mx, p0[static_cast<long long>(t) * d.M * 66]);
    float l = 0.f, acc = 0.f;
    for (int t = 0; t < d.T; ++t) {
      const float* p = p0 + static_cast<long long>(t) * d.M * 66;
      const float w = __expf(p[0] - mx);
      l += p[1] * w;
      acc += p[2 + c] * w;
    }
    out[(static_cast<long long>(b) * d.S + m) * d.ld_o + h * HD + c] = __float2bfloat16(acc / l);
    if (c == 0) lse[(static_cast<long long>(b) * d.H + h) * d.S + m] = mx + logf(l);
  }
}

// ======================================================================= backward
// dqkv: [B*S, 3C] bf16 (frame rows written here; the M global rows by the combine kernel);
// gpart: [B, H, T, M, 3, 64] fp32 partial dq/dk/dv of the global rows from this frame.
struct BwdSmem {
  uint32_t sQ, sK, sV, sdO;
  float* lse;
  float* delta;
};

__device__ __forceinline__ void p_and_ds(float s, float dp, float lse_l2, float delta, bool valid, float& p, float& ds) {
  p = valid ? fast_exp2(fmaf(s, LOG2E, -lse_l2)) : 0.f;
  ds = p * (dp - delta);
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
vip_attn_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                    const __nv_bfloat16* __restrict__ dout, const float* __restrict__ lse,
                    __nv_bfloat16* __restrict__ dqkv, float* __restrict__ gpart, const AttnDims d, float q_scale) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t sQ = (raw + 127u) & ~127u;
  const uint32_t sK = sQ + ROWS * 128, sV = sK + ROWS * 128, sdO = sV + ROWS * 128;
  float* s_lse = reinterpret_cast<float*>(smem_raw + (sdO + ROWS * 128 - raw));  // lse * log2(e), +inf on padding
  float* s_delta = s_lse + ROWS;
  const int t = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = d.M + d.L;
  const bool mask_gg = (t != 0);  // (global query, global key) pairs belong to frame 0 only

  stage_qkv(qkv, d, b, h, t, sQ, sK, sV);
  // dO rows + delta_i = sum_d dO[i,d] * O[i,d]  (8 lanes per row, 16 bytes each)
  for (int base = 0; base < ROWS; base += ATT_THREADS / 8) {
    const int row = base + (threadIdx.x >> 3), chunk = threadIdx.x & 7;
    if (row < ROWS) {
      float dot = 0.f;
      uint4 g = make_uint4(0, 0, 0, 0);
      if (row < nq) {
        const long long off = token_row(d, b, t, row) * d.ld_o + h * HD + chunk * 8;
        g = *reinterpret_cast<const uint4*>(dout + off);
        const uint4 o = *reinterpret_cast<const uint4*>(out + off);
        const uint32_t gw[4] = {g.x, g.y, g.z, g.w}, ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) dot += bf16_lo(gw[i]) * bf16_lo(ow[i]) + bf16_hi(gw[i]) * bf16_hi(ow[i]);
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tile_addr(sdO, row, chunk)), "r"(g.x), "r"(g.y),
                   "r"(g.z), "r"(g.w)
                   : "memory");
      dot += __shfl_xor_sync(0xffffffffu, dot, 1);
      dot += __shfl_xor_sync(0xffffffffu, dot, 2);
      dot += __shfl_xor_sync(0xffffffffu, dot, 4);
      if (chunk == 0) {
        s_delta[row] = dot;
        s_lse[row] = row < nq ? lse[(static_cast<long long>(b) * d.H + h) * d.S +
                                    (token_row(d, b, t, row) - static_cast<long long>(b) * d.S)] * LOG2E
                              : INFINITY;
      }
    }
  }
  cp_async_wait_all();
  __syncthreads();

  float* gp = gpart + ((static_cast<long long>(b) * d.H + h) * d.T + t) * d.M * 3 * HD;

  // ------------------------------------------------ pass A: key-stationary -> dK, dV
  for (int kt = warp; kt < NTILE; kt += ATT_WARPS) {
    if (kt * 16 >= nq) break;
    uint32_t ka[4][4], va[4][4];
    load_a_frags(sK, kt * 16, lane, ka);
    load_a_frags(sV, kt * 16, lane, va);
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    const int key_lo = kt * 16 + (lane >> 2);
#pragma unroll 1
    for (int qb = 0; qb < NTILE; ++qb) {
      if (qb * 16 >= nq) break;
      float st[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dpt[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bq[4], bo[4];
        load_b_nk(sQ, qb * 16, ks, lane, bq);
        load_b_nk(sdO, qb * 16, ks, lane, bo);
        mma_bf16(st[0], ka[ks], bq[0], bq[1]);
        mma_bf16(st[1], ka[ks], bq[2], bq[3]);
        mma_bf16(dpt[0], va[ks], bo[0], bo[1]);
        mma_bf16(dpt[1], va[ks], bo[2], bo[3]);
      }
      float pt[2][4], dst[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int q = qb * 16 + i * 8 + (lane & 3) * 2 + (e & 1);
          const int key = key_lo + (e >> 1) * 8;
          const bool valid = q < nq && key < nq && !(mask_gg && q < d.M && key < d.M);
          p_and_ds(st[i][e], dpt[i][e], s_lse[q], s_delta[q], valid, pt[i][e], dst[i][e]);
        }
      }
      uint32_t pa[4], da[4];
      pa[0] = pack_bf16(pt[0][0], pt[0][1]); pa[1] = pack_bf16(pt[0][2], pt[0][3]);
      pa[2] = pack_bf16(pt[1][0], pt[1][1]); pa[3] = pack_bf16(pt[1][2], pt[1][3]);
      da[0] = pack_bf16(dst[0][0], dst[0][1]); da[1] = pack_bf16(dst[0][2], dst[0][3]);
      da[2] = pack_bf16(dst[1][0], dst[1][1]); da[3] = pack_bf16(dst[1][2], dst[1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t bo[4], bq[4];
        load_b_kn(sdO, qb * 16, dp, lane, bo);
        load_b_kn(sQ, qb * 16, dp, lane, bq);
        mma_bf16(dv[2 * dp], pa, bo[0], bo[1]);
        mma_bf16(dv[2 * dp + 1], pa, bo[2], bo[3]);
        mma_bf16(dk[2 * dp], da, bq[0], bq[1]);
        mma_bf16(dk[2 * dp + 1], da, bq[2], bq[3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int key = key_lo + r * 8;
      if (key >= nq) continue;
      if (key >= d.M) {
        __nv_bfloat16* row = dqkv + token_row(d, b, t, key) * d.ld_qkv + h * HD + (lane & 3) * 2;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          *reinterpret_cast<uint32_t*>(row + d.C + i * 8) = pack_bf16(dk[i][2 * r], dk[i][2 * r + 1]);
          *reinterpret_cast<uint32_t*>(row + 2 * d.C + i * 8) = pack_bf16(dv[i][2 * r], dv[i][2 * r + 1]);
        }
      } else {
        float* g = gp + static_cast<long long>(key) * 3 * HD + (lane & 3) * 2;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          g[HD + i * 8] = dk[i][2 * r]; g[HD + i * 8 + 1] = dk[i][2 * r + 1];
          g[2 * HD + i * 8] = dv[i][2 * r]; g[2 * HD + i * 8 + 1] = dv[i][2 * r + 1];
        }
      }
    }
  }

  // ---------------------------------------------------- pass B: query-stationary -> dQ
  for (int qt = warp; qt < NTILE; qt += ATT_WARPS) {
    if (qt * 16 >= nq) break;
    uint32_t qa[4][4], oa[4][4];
    load_a_frags(sQ, qt * 16, lane, qa);
    load_a_frags(sdO, qt * 16, lane, oa);
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    const int q_lo = qt * 16 + (lane >> 2);
    const float lse_r[2] = {s_lse[q_lo], s_lse[q_lo + 8]}, del_r[2] = {s_delta[q_lo], s_delta[q_lo + 8]};
#pragma unroll 1
    for (int kb = 0; kb < NTILE; ++kb) {
      if (kb * 16 >= nq) break;
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp_[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t bk[4], bv[4];
        load_b_nk(sK, kb * 16, ks, lane, bk);
        load_b_nk(sV, kb * 16, ks, lane, bv);
        mma_bf16(s[0], qa[ks], bk[0], bk[1]);
        mma_bf16(s[1], qa[ks], bk[2], bk[3]);
        mma_bf16(dp_[0], oa[ks], bv[0], bv[1]);
        mma_bf16(dp_[1], oa[ks], bv[2], bv[3]);
      }
      float ds[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = kb * 16 + i * 8 + (lane & 3) * 2 + (e & 1);
          const int q = q_lo + (e >> 1) * 8;
          const bool valid = q < nq && key < nq && !(mask_gg && q < d.M && key < d.M);
          float p;
          p_and_ds(s[i][e], dp_[i][e], lse_r[e >> 1], del_r[e >> 1], valid, p, ds[i][e]);
        }
      }
      uint32_t da[4];
      da[0] = pack_bf16(ds[0][0], ds[0][1]); da[1] = pack_bf16(ds[0][2], ds[0][3]);
      da[2] = pack_bf16(ds[1][0], ds[1][1]); da[3] = pack_bf16(ds[1][2], ds[1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t bk[4];
        load_b_kn(sK, kb * 16, dp, lane, bk);
        mma_bf16(dq[2 * dp], da, bk[0], bk[1]);
        mma_bf16(dq[2 * dp + 1], da, bk[2], bk[3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int q = q_lo + r * 8;
      if (q >= nq) continue;
      if (q >= d.M) {
        __nv_bfloat16* row = dqkv + token_row(d, b, t, q) * d.ld_qkv + h * HD + (lane & 3) * 2;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<uint32_t*>(row + i * 8) = pack_bf16(dq[i][2 * r] * q_scale, dq[i][2 * r + 1] * q_scale);
      } else {
        float* g = gp + static_cast<long long>(q) * 3 * HD + (lane & 3) * 2;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          g[i * 8] = dq[i][2 * r];
          g[i * 8 + 1] = dq[i][2 * r + 1];
        }
      }
    }
  }
}

// Sum the per-frame partial gradients of the M global rows: grid (H, B), block 192 = 3 x 64.
__global__ void __launch_bounds__(192)
vip_attn_bwd_combine_kernel(const float* __restrict__ gpart, __nv_bfloat16* __restrict__ dqkv, const AttnDims d,
                            float q_scale) {
  const int h = blockIdx.x, b = blockIdx.y;
  const int which = threadIdx.x / HD, c = threadIdx.x % HD;
  for (int m = 0; m < d.M; ++m) {
    float acc = 0.f;
    for (int t = 0; t < d.T; ++t)
      acc += gpart[((((static_cast<long long>(b) * d.H + h) * d.T + t) * d.M + m) * 3 + which) * HD + c];
    if (which == 0) acc *= q_scale;
    dqkv[(static_cast<long long>(b) * d.S + m) * d.ld_qkv + static_cast<long long>(which) * d.C + h * HD + c] =
        __float2bfloat16(acc);
  }
}

static int make_dims(AttnDims& d, int B, int H, int T, int L, int M, int C) {
  if (C != H * HD) return fail("vip_attention: head_dim must be 64 (C == 64*H)");
  if (M + L > ROWS) return fail("vip_attention: M + L must be <= 208");
  if (M < 1 || M > 8) return fail("vip_attention: 1 <= M <= 8 global tokens");
  d.B = B; d.H = H; d.T = T; d.L = L; d.M = M;
  d.S = static_cast<long long>(M) + static_cast<long long>(T) * L;
  d.C = C;
  d.ld_qkv = 3LL * C;
  d.ld_o = C;
  return 0;
}

}  // namespace xp

using namespace xp;

extern "C" int64_t xp_vip_attention_workspace_bytes(int32_t B, int32_t H, int32_t T, int32_t M) {
  // forward partials (66 floats) and backward partials (192 floats) per (b, h, t, m); sized for the larger
  return static_cast<int64_t>(B) * H * T * M * 3 * HD * sizeof(float);
}

extern "C" int xp_vip_attention_fwd(const void* qkv, void* out, float* lse, float* workspace, int32_t B, int32_t H,
                                    int32_t T, int32_t L, int32_t M, int32_t C, void* stream) {
  XP_ENTER(qkv);
  AttnDims d;
  if (make_dims(d, B, H, T, L, M, C)) return -1;
  const int smem = 3 * ROWS * 128 + 128;
  static bool attr = false;
  if (!attr) {
    XP_CHECK_CUDA(cudaFuncSetAttribute(vip_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  vip_attn_fwd_kernel<<<dim3(T, H, B), ATT_THREADS, smem, st>>>(static_cast<const __nv_bfloat16*>(qkv),
                                                               static_cast<__nv_bfloat16*>(out), lse, workspace, d);
  XP_CHECK_LAUNCH("vip_attn_fwd_kernel");
  vip_attn_fwd_combine_kernel<<<dim3(H, B), 64, 0, st>>>(workspace, static_cast<__nv_bfloat16*>(out), lse, d);
  XP_CHECK_LAUNCH("vip_attn_fwd_combine_kernel");
  return 0;
}

extern "C" int xp_vip_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                    float* workspace, int32_t B, int32_t H, int32_t T, int32_t L, int32_t M, int32_t C,
                                    float q_scale, void* stream) {
  XP_ENTER(qkv);
  AttnDims d;
  if (make_dims(d, B, H, T, L, M, C)) return -1;
  const int smem = 4 * ROWS * 128 + 2 * ROWS * 4 + 128;
  static bool attr = false;
  if (!attr) {
    XP_CHECK_CUDA(cudaFuncSetAttribute(vip_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  vip_attn_bwd_kernel<<<dim3(T, H, B), ATT_THREADS, smem, st>>>(
      static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(out),
      static_cast<const __nv_bfloat16*>(dout), lse, static_cast<__nv_bfloat16*>(dqkv), workspace, d, q_scale);
  XP_CHECK_LAUNCH("vip_attn_bwd_kernel");
  vip_attn_bwd_combine_kernel<<<dim3(H, B), 192, 0, st>>>(workspace, static_cast<__nv_bfloat16*>(dqkv), d, q_scale);
  XP_CHECK_LAUNCH("vip_attn_bwd_combine_kernel");
  return 0;
}

// tcgen05 forward (vip_attention_tc.cu) + the shared combine kernel for the global-query rows.
extern "C" int xp_vip_attention_fwd_tc_partial(const void* qkv, void* out, float* lse, float* workspace, int32_t B,
                                               int32_t H, int32_t T, int32_t L, int32_t M, int32_t C, void* stream);
extern "C" int xp_vip_attention_fwd_tc(const void* qkv, void* out, float* lse, float* workspace, int32_t B, int32_t H,
                                       int32_t T, int32_t L, int32_t M, int32_t C, void* stream) {
  XP_ENTER(qkv);
  AttnDims d;
  if (make_dims(d, B, H, T, L, M, C)) return -1;
  if (xp_vip_attention_fwd_tc_partial(qkv, out, lse, workspace, B, H, T, L, M, C, stream)) return -1;
  vip_attn_fwd_combine_kernel<<<dim3(H, B), 64, 0, static_cast<cudaStream_t>(stream)>>>(
      workspace, static_cast<__nv_bfloat16*>(out), lse, d);
  XP_CHECK_LAUNCH("vip_attn_fwd_combine_kernel");
  return 0;
}

// tcgen05 backward (vip_attention_tc.cu) + the shared combine kernel for the M global rows.
extern "C" int xp_vip_attention_bwd_tc_partial(const void* qkv, const void* out, const void* dout, const float* lse,
                                               void* dqkv, float* workspace, float* delta, int32_t B, int32_t H,
                                               int32_t T, int32_t L, int32_t M, int32_t C, float q_scale, void* stream);
extern "C" int xp_vip_attention_bwd_tc(const void* qkv, const void* out, const void* dout, const float* lse, void* dqkv,
                                       float* workspace, float* delta, int32_t B, int32_t H, int32_t T, int32_t L,
                                       int32_t M, int32_t C, float q_scale, void* stream) {
  XP_ENTER(qkv);
  AttnDims d;
  if (make_dims(d, B, H, T, L, M, C)) return -1;
  if (xp_vip_attention_bwd_tc_partial(qkv, out, dout, lse, dqkv, workspace, delta, B, H, T, L, M, C, q_scale, stream))
    return -1;
  vip_attn_bwd_combine_kernel<<<dim3(H, B), 192, 0, static_cast<cudaStream_t>(stream)>>>(
      workspace, static_cast<__nv_bfloat16*>(dqkv), d, q_scale);
  XP_CHECK_LAUNCH("vip_attn_bwd_combine_kernel");
  return 0;
}
