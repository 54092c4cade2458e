// Tensor-core self-attention for LONG sequences and head dims 48 / 64 (sm_100a): the attention of the PLM head, which
// the reference runs ACROSS the N news of a call (text.py:96: batch-first states handed to a batch_first=False
// nn.MultiheadAttention -> sequence length S = N, e.g. 400 clicked news, head dim 768 / 16 = 48), and any other
// MHSA + additive block whose sequence does not fit the register / tile-resident kernels.  Replaces the fp32 SIMT
// streaming kernels (attn_fwd_kernel / attn_bwd_kernel) on that path.
//
// Flash-style tiling, no S x S matrix in memory:
//   forward   CTA = (batch item, head, block of 128 queries): 8 warps x 16 query rows; the 64-key blocks of K and V are
//             staged one after the other; scores, online softmax (running max / sum per row) and the PV product stay in
//             registers.
//   backward  delta kernel (D_t = dO_t . O_t per row and head), then
//             dQ kernel   CTA = 128 queries, walks the key blocks:   P = exp(s - lse), dS = P (dO V^T - D), dQ += dS K
//             dKV kernel  CTA = 128 keys, walks the query blocks with the transposed tiles: dV += P^T dO, dK += dS^T Q
// Operands are staged from the fp32 qkv / dO rows (any (seq_stride, batch_stride) geometry) into shared memory as bf16
// hi / lo planes and fetched with ldmatrix; every contraction is mma.sync.m16n8k16 with the three-pass hi/lo scheme
// (single pass in the bf16 configuration), fp32 accumulation -- the same arithmetic as the other attention kernels.
#pragma once
#include "nrl_tfm.cuh"

namespace nrl {

template <int DH>
struct FlashCfg {
  static constexpr int PITCH = DH + 8;   // bf16 per staged row: 16-byte row starts fall into distinct bank groups
  static constexpr int ROWB = PITCH * 2;
  static constexpr int KS = DH / 16;     // k-steps of a product over the head dim
  static constexpr int DT = DH / 8;      // 8-wide n-tiles of the head dim
  static constexpr int QB = 128, KB = 64;
};
template <int ROWB>
__device__ __forceinline__ uint32_t fl_a_addr(uint32_t base, int row0, int k0, int lane) {
  const int mi = lane >> 3;
  return base + (uint32_t)((row0 + (lane & 7) + 8 * (mi & 1)) * ROWB + (k0 + 8 * (mi >> 1)) * 2);
}
template <int ROWB>
__device__ __forceinline__ uint32_t fl_bt_addr(uint32_t base, int n0, int k0, int lane) {
  const int mi = lane >> 3;
  return base + (uint32_t)((n0 + (lane & 7) + 8 * (mi >> 1)) * ROWB + (k0 + 8 * (mi & 1)) * 2);
}
template <int ROWB>
__device__ __forceinline__ uint32_t fl_bn_addr(uint32_t base, int k0, int n0, int lane) {
  const int mi = lane >> 3;
  return base + (uint32_t)((k0 + (lane & 7) + 8 * (mi & 1)) * ROWB + (n0 + 8 * (mi >> 1)) * 2);
}
// stage ROWS sequence rows s0 .. of one (batch item, column offset) slice as hi / lo planes; rows >= S are zero.  256
// threads; every thread issues all its 16-byte loads before the first conversion (one trip to L2 / HBM per matrix).
template <int DH, int ROWS>
__device__ __forceinline__ void fl_stage(uint8_t* sm, uint32_t o_hi, uint32_t o_lo, const float* __restrict__ src,
                                         long long ld, long long seq_stride, long long row_base, int s0, int S, float mul) {
  constexpr int C4 = DH / 4, ROWB = FlashCfg<DH>::ROWB, TOTAL = ROWS * C4, ITER = (TOTAL + 255) / 256;
  float4 v[ITER];
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int i = threadIdx.x + it * 256, r = i / C4, c = (i % C4) * 4;
    v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < TOTAL && s0 + r < S)
      v[it] = __ldg(reinterpret_cast<const float4*>(src + ((long long)(s0 + r) * seq_stride + row_base) * ld + c));
  }
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int i = threadIdx.x + it * 256, r = i / C4, c = (i % C4) * 4;
    if (i < TOTAL) {
      uint32_t h0, l0, h1, l1;
      split_pack2(v[it].x * mul, v[it].y * mul, h0, l0);
      split_pack2(v[it].z * mul, v[it].w * mul, h1, l1);
      const uint32_t off = (uint32_t)(r * ROWB + c * 2);
      *reinterpret_cast<uint2*>(sm + o_hi + off) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(sm + o_lo + off) = make_uint2(l0, l1);
    }
  }
}
// optional key-padding mask and attention-probability dropout (the PLM transformer at more than 128 tokens): per block,
// 128 key-mask bytes and a 128 x 64 (64 x 128 in the dK/dV kernel) keep-bit tile = 256 words, drawn by the CTA's 256
// threads with the element numbering of tfm_drop_bits (SP = S rounded up to 32), so that every kernel and
// nrl_tfm_attn_dropout_mask see the same bits
constexpr int FLASH_AUX_BYTES = 128 + 256 * 4;
__device__ __forceinline__ void fl_drop_tile(uint32_t* bits, int words_per_row, unsigned long long item, int SP, int t0,
                                             int u0, unsigned long long seed, uint32_t site, uint32_t thr) {
  const int r = threadIdx.x / words_per_row, w = threadIdx.x % words_per_row;
  bits[threadIdx.x] = drop_keep_bits32(seed, site, (item * (unsigned long long)SP + (unsigned long long)(t0 + r)) * SP +
                                                       (unsigned long long)(u0 + 32 * w), thr);
}

template <int DH>
__host__ __device__ constexpr int flash_fwd_smem() { return (2 * 128 + 4 * 64) * FlashCfg<DH>::ROWB + 16 + FLASH_AUX_BYTES; }
template <int DH>
__host__ __device__ constexpr int flash_bwd_smem() { return (4 * 128 + 4 * 64) * FlashCfg<DH>::ROWB + 2 * 128 * 4 + 16 + FLASH_AUX_BYTES; }

// ---------------------------------------------------------------- forward
template <int DH>
__global__ void __launch_bounds__(256)
attn_fwd_flash_kernel(const float* __restrict__ qkv, int E, int ldq, int heads, int S, long long seq_stride, int NB,
                      long long batch_stride, float scale, __nv_bfloat16* __restrict__ o_hi,
                      __nv_bfloat16* __restrict__ o_lo, int ep, float* __restrict__ lse, int three_i,
                      const unsigned char* __restrict__ kmask = nullptr, int drop_on = 0, uint32_t thr = 0,
                      float dscale = 1.f, unsigned long long seed = 0, uint32_t site = 0) {
  using C = FlashCfg<DH>;
  constexpr int ROWB = C::ROWB;
  extern __shared__ __align__(16) uint8_t fsm[];
  const bool three = three_i != 0;
  const int nqb = (S + C::QB - 1) / C::QB;
  const int qb = blockIdx.x % nqb, h = (blockIdx.x / nqb) % heads, b = blockIdx.x / (nqb * heads);
  const long long row_base = (long long)b * batch_stride;
  const uint32_t oQh = 0, oQl = C::QB * ROWB, oKh = 2 * C::QB * ROWB, oKl = oKh + C::KB * ROWB, oVh = oKl + C::KB * ROWB,
                 oVl = oVh + C::KB * ROWB;
  unsigned char* km = fsm + oVl + C::KB * ROWB;
  uint32_t* bits = reinterpret_cast<uint32_t*>(km + 128);
  const int SP = (S + 31) & ~31;
  const unsigned long long item = (unsigned long long)b * heads + h;
  const int q0 = qb * C::QB;
  fl_stage<DH, C::QB>(fsm, oQh, oQl, qkv + h * DH, ldq, seq_stride, row_base, q0, S, scale * TFM_LOG2E);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int w0 = 16 * warp;
  const bool active = q0 + w0 < S;
  const uint32_t sb = smem_u32(fsm);
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float o[C::DT][4];
#pragma unroll
  for (int j = 0; j < C::DT; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  for (int k0 = 0; k0 < S; k0 += C::KB) {
    __syncthreads();  // the previous block's K / V are no longer read (first pass: Q staged)
    fl_stage<DH, C::KB>(fsm, oKh, oKl, qkv + E + h * DH, ldq, seq_stride, row_base, k0, S, 1.f);
    fl_stage<DH, C::KB>(fsm, oVh, oVl, qkv + 2 * E + h * DH, ldq, seq_stride, row_base, k0, S, 1.f);
    if (threadIdx.x < C::KB) km[threadIdx.x] = (k0 + threadIdx.x < S && (!kmask || kmask[(long long)b * S + k0 + threadIdx.x])) ? 1 : 0;
    if (drop_on) fl_drop_tile(bits, 2, item, SP, q0, k0, seed, site, thr);
    __syncthreads();
    if (!active) continue;
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < C::KS; ++kk) {
      uint32_t ah[4], al[4] = {0u, 0u, 0u, 0u};
      ldsm_x4(ah, fl_a_addr<ROWB>(sb + oQh, w0, 16 * kk, lane));
      if (three) ldsm_x4(al, fl_a_addr<ROWB>(sb + oQl, w0, 16 * kk, lane));
#pragma unroll
      for (int j2 = 0; j2 < 4; ++j2) {
        uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
        ldsm_x4(bh, fl_bt_addr<ROWB>(sb + oKh, 16 * j2, 16 * kk, lane));
        if (three) ldsm_x4(bl, fl_bt_addr<ROWB>(sb + oKl, 16 * j2, 16 * kk, lane));
        tfm_mma3(s[2 * j2], ah, al, bh[0], bh[1], bl[0], bl[1], three);
        tfm_mma3(s[2 * j2 + 1], ah, al, bh[2], bh[3], bl[2], bl[3], three);
      }
    }
    // online softmax (log2 domain); keys >= S and masked keys take no part
    float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int u = 8 * j + 2 * tg;
      if (!km[u]) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
      if (!km[u + 1]) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      bm0 = fmaxf(bm0, fmaxf(s[j][0], s[j][1]));
      bm1 = fmaxf(bm1, fmaxf(s[j][2], s[j][3]));
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1)); bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1)); bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    float n0 = fmaxf(m0, bm0), n1 = fmaxf(m1, bm1);
    if (n0 == -INFINITY) n0 = 0.f;  // nothing but masked keys so far: every p below is exp2(-inf) = 0
    if (n1 == -INFINITY) n1 = 0.f;
    const float c0 = ex2_approx(m0 - n0), c1 = ex2_approx(m1 - n1);
    m0 = n0; m1 = n1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = ex2_approx(s[j][0] - n0); s[j][1] = ex2_approx(s[j][1] - n0);
      s[j][2] = ex2_approx(s[j][2] - n1); s[j][3] = ex2_approx(s[j][3] - n1);
      ps0 += s[j][0] + s[j][1]; ps1 += s[j][2] + s[j][3];
    }
    ps0 += __shfl_xor_sync(0xffffffffu, ps0, 1); ps0 += __shfl_xor_sync(0xffffffffu, ps0, 2);
    ps1 += __shfl_xor_sync(0xffffffffu, ps1, 1); ps1 += __shfl_xor_sync(0xffffffffu, ps1, 2);
    l0 = l0 * c0 + ps0; l1 = l1 * c1 + ps1;
#pragma unroll
    for (int j = 0; j < C::DT; ++j) { o[j][0] *= c0; o[j][1] *= c0; o[j][2] *= c1; o[j][3] *= c1; }
    if (drop_on) {  // dropped probabilities leave the PV product, not the normaliser
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int sh = 8 * (j & 3) + 2 * tg;
        const uint32_t b0 = bits[(w0 + g) * 2 + (j >> 2)] >> sh, b1 = bits[(w0 + g + 8) * 2 + (j >> 2)] >> sh;
        if (!(b0 & 1u)) s[j][0] = 0.f;
        if (!(b0 & 2u)) s[j][1] = 0.f;
        if (!(b1 & 1u)) s[j][2] = 0.f;
        if (!(b1 & 2u)) s[j][3] = 0.f;
      }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {  // 64 keys = 4 k-steps
      uint32_t ah[4], al[4];
      split_pack2(s[2 * kk][0], s[2 * kk][1], ah[0], al[0]);
      split_pack2(s[2 * kk][2], s[2 * kk][3], ah[1], al[1]);
      split_pack2(s[2 * kk + 1][0], s[2 * kk + 1][1], ah[2], al[2]);
      split_pack2(s[2 * kk + 1][2], s[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
      for (int jd2 = 0; jd2 < C::DT / 2; ++jd2) {
        uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
        ldsm_x4_t(bh, fl_bn_addr<ROWB>(sb + oVh, 16 * kk, 16 * jd2, lane));
        if (three) ldsm_x4_t(bl, fl_bn_addr<ROWB>(sb + oVl, 16 * kk, 16 * jd2, lane));
        tfm_mma3(o[2 * jd2], ah, al, bh[0], bh[1], bl[0], bl[1], three);
        tfm_mma3(o[2 * jd2 + 1], ah, al, bh[2], bh[3], bl[2], bl[3], three);
      }
    }
  }
  if (!active) return;
  const float keep_scale = drop_on ? dscale : 1.f;
  const float i0 = l0 > 0.f ? keep_scale / l0 : 0.f, i1 = l1 > 0.f ? keep_scale / l1 : 0.f;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int t = q0 + w0 + g + 8 * half;
    if (t >= S) continue;
    const long long grow = (long long)t * seq_stride + row_base;
    const float inv = half ? i1 : i0;
#pragma unroll
    for (int j = 0; j < C::DT; ++j) {
      uint32_t hh, ll;
      split_pack2(o[j][2 * half] * inv, o[j][2 * half + 1] * inv, hh, ll);
      const long long off = grow * ep + h * DH + 8 * j + 2 * tg;
      *reinterpret_cast<uint32_t*>(o_hi + off) = hh;
      if (o_lo) *reinterpret_cast<uint32_t*>(o_lo + off) = ll;
    }
    if (tg == 0) lse[grow * heads + h] = (half ? l1 : l0) > 0.f ? ((half ? m1 : m0) + log2f(half ? l1 : l0)) * TFM_LN2 : 0.f;
    if (h == 0)
      for (int c = E + tg; c < ep; c += 4) {
        o_hi[grow * ep + c] = __float2bfloat16_rn(c == E ? 1.f : 0.f);
        if (o_lo) o_lo[grow * ep + c] = __float2bfloat16_rn(0.f);
      }
  }
}

// ---------------------------------------------------------------- backward
// delta[row, h] = dO[row, h*DH .. ] . O[row, h*DH ..]  (O from its hi / lo planes).  One warp per row.
__global__ void attn_delta_kernel(const float* __restrict__ d_o, long long ld_do, const __nv_bfloat16* __restrict__ o_hi,
                                  const __nv_bfloat16* __restrict__ o_lo, int ep, long long rows, int heads, int DH,
                                  float* __restrict__ delta) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < rows; r += nwarps) {
    for (int h = 0; h < heads; ++h) {
      float acc = 0.f;
      for (int d = lane; d < DH; d += 32) {
        const int c = h * DH + d;
        float ov = __bfloat162float(o_hi[r * ep + c]);
        if (o_lo) ov += __bfloat162float(o_lo[r * ep + c]);
        acc += d_o[r * ld_do + c] * ov;
      }
      acc = warp_sum(acc);
      if (lane == 0) delta[r * heads + h] = acc;
    }
  }
}

// per-row scalars of a block of sequence rows -> shared memory: lse * log2(e) (+inf beyond S: exp -> 0) and delta
__device__ __forceinline__ void fl_stage_rowstats(float* lse2, float* dd, const float* __restrict__ lse,
                                                  const float* __restrict__ delta, int heads, int h, long long seq_stride,
                                                  long long row_base, int s0, int rows, int S) {
  for (int r = threadIdx.x; r < rows; r += blockDim.x) {
    const long long grow = (long long)(s0 + r) * seq_stride + row_base;
    lse2[r] = s0 + r < S ? lse[grow * heads + h] * TFM_LOG2E : INFINITY;
    dd[r] = s0 + r < S ? delta[grow * heads + h] : 0.f;
  }
}

template <int DH>
__global__ void __launch_bounds__(256)
attn_bwd_flash_dq_kernel(const float* __restrict__ qkv, const float* __restrict__ d_o, long long ld_do,
                         const float* __restrict__ lse, const float* __restrict__ delta, int E, int ldq, int heads, int S,
                         long long seq_stride, int NB, long long batch_stride, float scale,
                         __nv_bfloat16* __restrict__ g_hi, __nv_bfloat16* __restrict__ g_lo, int p3, int three_i,
                         const unsigned char* __restrict__ kmask = nullptr, int drop_on = 0, uint32_t thr = 0,
                         float dscale = 1.f, unsigned long long seed = 0, uint32_t site = 0) {
  using C = FlashCfg<DH>;
  constexpr int ROWB = C::ROWB;
  extern __shared__ __align__(16) uint8_t fsm[];
  const bool three = three_i != 0;
  const int nqb = (S + C::QB - 1) / C::QB;
  const int qb = blockIdx.x % nqb, h = (blockIdx.x / nqb) % heads, b = blockIdx.x / (nqb * heads);
  const long long row_base = (long long)b * batch_stride;
  const uint32_t oQh = 0, oQl = C::QB * ROWB, oGh = 2 * C::QB * ROWB, oGl = 3 * C::QB * ROWB, oKh = 4 * C::QB * ROWB,
                 oKl = oKh + C::KB * ROWB, oVh = oKl + C::KB * ROWB, oVl = oVh + C::KB * ROWB;
  float* lse2 = reinterpret_cast<float*>(fsm + oVl + C::KB * ROWB);
  float* dd = lse2 + C::QB;
  unsigned char* km = reinterpret_cast<unsigned char*>(dd + C::QB);
  uint32_t* bits = reinterpret_cast<uint32_t*>(km + 128);
  const int SP = (S + 31) & ~31;
  const unsigned long long item = (unsigned long long)b * heads + h;
  const float c_keep = drop_on ? dscale : 1.f;
  const int q0 = qb * C::QB;
  fl_stage<DH, C::QB>(fsm, oQh, oQl, qkv + h * DH, ldq, seq_stride, row_base, q0, S, scale * TFM_LOG2E);
  fl_stage<DH, C::QB>(fsm, oGh, oGl, d_o + h * DH, ld_do, seq_stride, row_base, q0, S, 1.f);
  fl_stage_rowstats(lse2, dd, lse, delta, heads, h, seq_stride, row_base, q0, C::QB, S);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int w0 = 16 * warp;
  const bool active = q0 + w0 < S;
  const uint32_t sb = smem_u32(fsm);
  float dq[C::DT][4];
#pragma unroll
  for (int j = 0; j < C::DT; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
  for (int k0 = 0; k0 < S; k0 += C::KB) {
    __syncthreads();
    fl_stage<DH, C::KB>(fsm, oKh, oKl, qkv + E + h * DH, ldq, seq_stride, row_base, k0, S, 1.f);
    fl_stage<DH, C::KB>(fsm, oVh, oVl, qkv + 2 * E + h * DH, ldq, seq_stride, row_base, k0, S, 1.f);
    if (threadIdx.x < C::KB) km[threadIdx.x] = (k0 + threadIdx.x < S && (!kmask || kmask[(long long)b * S + k0 + threadIdx.x])) ? 1 : 0;
    if (drop_on) fl_drop_tile(bits, 2, item, SP, q0, k0, seed, site, thr);
    __syncthreads();
    if (!active) continue;
    const float ls0 = lse2[w0 + g], ls1 = lse2[w0 + g + 8], d0 = dd[w0 + g], d1 = dd[w0 + g + 8];
    float s[8][4], dp[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < C::KS; ++kk) {
      uint32_t qh[4], ql[4] = {0u, 0u, 0u, 0u}, gh[4], gl[4] = {0u, 0u, 0u, 0u};
      ldsm_x4(qh, fl_a_addr<ROWB>(sb + oQh, w0, 16 * kk, lane));
      ldsm_x4(gh, fl_a_addr<ROWB>(sb + oGh, w0, 16 * kk, lane));
      if (three) {
        ldsm_x4(ql, fl_a_addr<ROWB>(sb + oQl, w0, 16 * kk, lane));
        ldsm_x4(gl, fl_a_addr<ROWB>(sb + oGl, w0, 16 * kk, lane));
      }
#pragma unroll
      for (int j2 = 0; j2 < 4; ++j2) {
        uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
        ldsm_x4(bh, fl_bt_addr<ROWB>(sb + oKh, 16 * j2, 16 * kk, lane));
        if (three) ldsm_x4(bl, fl_bt_addr<ROWB>(sb + oKl, 16 * j2, 16 * kk, lane));
        tfm_mma3(s[2 * j2], qh, ql, bh[0], bh[1], bl[0], bl[1], three);
        tfm_mma3(s[2 * j2 + 1], qh, ql, bh[2], bh[3], bl[2], bl[3], three);
        ldsm_x4(bh, fl_bt_addr<ROWB>(sb + oVh, 16 * j2, 16 * kk, lane));
        if (three) ldsm_x4(bl, fl_bt_addr<ROWB>(sb + oVl, 16 * j2, 16 * kk, lane));
        tfm_mma3(dp[2 * j2], gh, gl, bh[0], bh[1], bl[0], bl[1], three);
        tfm_mma3(dp[2 * j2 + 1], gh, gl, bh[2], bh[3], bl[2], bl[3], three);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int u = 8 * j + 2 * tg;
      const bool v0 = km[u] != 0, v1 = km[u + 1] != 0;
      float k00 = c_keep, k01 = c_keep, k10 = c_keep, k11 = c_keep;
      if (drop_on) {
        const int sh = 8 * (j & 3) + 2 * tg;
        const uint32_t b0 = bits[(w0 + g) * 2 + (j >> 2)] >> sh, b1 = bits[(w0 + g + 8) * 2 + (j >> 2)] >> sh;
        if (!(b0 & 1u)) k00 = 0.f;
        if (!(b0 & 2u)) k01 = 0.f;
        if (!(b1 & 1u)) k10 = 0.f;
        if (!(b1 & 2u)) k11 = 0.f;
      }
      const float p00 = v0 ? ex2_approx(s[j][0] - ls0) : 0.f, p01 = v1 ? ex2_approx(s[j][1] - ls0) : 0.f;
      const float p10 = v0 ? ex2_approx(s[j][2] - ls1) : 0.f, p11 = v1 ? ex2_approx(s[j][3] - ls1) : 0.f;
      s[j][0] = p00 * (k00 * dp[j][0] - d0); s[j][1] = p01 * (k01 * dp[j][1] - d0);
      s[j][2] = p10 * (k10 * dp[j][2] - d1); s[j][3] = p11 * (k11 * dp[j][3] - d1);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ah[4], al[4];
      split_pack2(s[2 * kk][0], s[2 * kk][1], ah[0], al[0]);
      split_pack2(s[2 * kk][2], s[2 * kk][3], ah[1], al[1]);
      split_pack2(s[2 * kk + 1][0], s[2 * kk + 1][1], ah[2], al[2]);
      split_pack2(s[2 * kk + 1][2], s[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
      for (int jd2 = 0; jd2 < C::DT / 2; ++jd2) {
        uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
        ldsm_x4_t(bh, fl_bn_addr<ROWB>(sb + oKh, 16 * kk, 16 * jd2, lane));
        if (three) ldsm_x4_t(bl, fl_bn_addr<ROWB>(sb + oKl, 16 * kk, 16 * jd2, lane));
        tfm_mma3(dq[2 * jd2], ah, al, bh[0], bh[1], bl[0], bl[1], three);
        tfm_mma3(dq[2 * jd2 + 1], ah, al, bh[2], bh[3], bl[2], bl[3], three);
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int t = q0 + w0 + g + 8 * half;
    if (t >= S) continue;
    const long long rowoff = ((long long)t * seq_stride + row_base) * p3;
#pragma unroll
    for (int j = 0; j < C::DT; ++j) {
      uint32_t hh, ll;
      split_pack2(dq[j][2 * half] * scale, dq[j][2 * half + 1] * scale, hh, ll);
      const long long off = rowoff + h * DH + 8 * j + 2 * tg;
      *reinterpret_cast<uint32_t*>(g_hi + off) = hh;
      if (g_lo) *reinterpret_cast<uint32_t*>(g_lo + off) = ll;
    }
    if (h == 0)
      for (int c = 3 * E + tg; c < p3; c += 4) {
        g_hi[rowoff + c] = __float2bfloat16_rn(0.f);
        if (g_lo) g_lo[rowoff + c] = __float2bfloat16_rn(0.f);
      }
  }
}

template <int DH>
__global__ void __launch_bounds__(256)
attn_bwd_flash_dkv_kernel(const float* __restrict__ qkv, const float* __restrict__ d_o, long long ld_do,
                          const float* __restrict__ lse, const float* __restrict__ delta, int E, int ldq, int heads, int S,
                          long long seq_stride, int NB, long long batch_stride, float scale,
                          __nv_bfloat16* __restrict__ g_hi, __nv_bfloat16* __restrict__ g_lo, int p3, int three_i,
                          const unsigned char* __restrict__ kmask = nullptr, int drop_on = 0, uint32_t thr = 0,
                          float dscale = 1.f, unsigned long long seed = 0, uint32_t site = 0) {
  using C = FlashCfg<DH>;
  constexpr int ROWB = C::ROWB;
  extern __shared__ __align__(16) uint8_t fsm[];
  const bool three = three_i != 0;
  const int nkb = (S + C::QB - 1) / C::QB;  // key blocks of 128 rows (the CTA's own), query blocks of 64
  const int kb = blockIdx.x % nkb, h = (blockIdx.x / nkb) % heads, b = blockIdx.x / (nkb * heads);
  const long long row_base = (long long)b * batch_stride;
  const uint32_t oKh = 0, oKl = C::QB * ROWB, oVh = 2 * C::QB * ROWB, oVl = 3 * C::QB * ROWB, oQh = 4 * C::QB * ROWB,
                 oQl = oQh + C::KB * ROWB, oGh = oQl + C::KB * ROWB, oGl = oGh + C::KB * ROWB;
  float* lse2 = reinterpret_cast<float*>(fsm + oGl + C::KB * ROWB);
  float* dd = lse2 + C::QB;
  uint32_t* bits = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(dd + C::QB) + 128);
  const int SP = (S + 31) & ~31;
  const unsigned long long item = (unsigned long long)b * heads + h;
  const float c_keep = drop_on ? dscale : 1.f;
  const int k0 = kb * C::QB;
  fl_stage<DH, C::QB>(fsm, oKh, oKl, qkv + E + h * DH, ldq, seq_stride, row_base, k0, S, 1.f);
  fl_stage<DH, C::QB>(fsm, oVh, oVl, qkv + 2 * E + h * DH, ldq, seq_stride, row_base, k0, S, 1.f);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int w0 = 16 * warp;
  const bool active = k0 + w0 < S;
  const int u0 = k0 + w0 + g, u1 = u0 + 8;
  const bool kv0 = u0 < S && (!kmask || kmask[(long long)b * S + u0]), kv1 = u1 < S && (!kmask || kmask[(long long)b * S + u1]);
  const uint32_t sb = smem_u32(fsm);
  float dk[C::DT][4], dv[C::DT][4];
#pragma unroll
  for (int j = 0; j < C::DT; ++j) {
    dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
    dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
  }
  for (int q0 = 0; q0 < S; q0 += C::KB) {
    __syncthreads();
    fl_stage<DH, C::KB>(fsm, oQh, oQl, qkv + h * DH, ldq, seq_stride, row_base, q0, S, scale * TFM_LOG2E);
    fl_stage<DH, C::KB>(fsm, oGh, oGl, d_o + h * DH, ld_do, seq_stride, row_base, q0, S, 1.f);
    fl_stage_rowstats(lse2, dd, lse, delta, heads, h, seq_stride, row_base, q0, C::KB, S);
    if (drop_on) fl_drop_tile(bits, 4, item, SP, q0, k0, seed, site, thr);  // [64 queries][128 keys]
    __syncthreads();
    if (!active) continue;
    float st[8][4], dpt[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
      dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < C::KS; ++kk) {
      uint32_t kh[4], kl[4] = {0u, 0u, 0u, 0u}, vh[4], vl[4] = {0u, 0u, 0u, 0u};
      ldsm_x4(kh, fl_a_addr<ROWB>(sb + oKh, w0, 16 * kk, lane));
      ldsm_x4(vh, fl_a_addr<ROWB>(sb + oVh, w0, 16 * kk, lane));
      if (three) {
        ldsm_x4(kl, fl_a_addr<ROWB>(sb + oKl, w0, 16 * kk, lane));
        ldsm_x4(vl, fl_a_addr<ROWB>(sb + oVl, w0, 16 * kk, lane));
      }
#pragma unroll
      for (int j2 = 0; j2 < 4; ++j2) {
        uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
        ldsm_x4(bh, fl_bt_addr<ROWB>(sb + oQh, 16 * j2, 16 * kk, lane));
        if (three) ldsm_x4(bl, fl_bt_addr<ROWB>(sb + oQl, 16 * j2, 16 * kk, lane));
        tfm_mma3(st[2 * j2], kh, kl, bh[0], bh[1], bl[0], bl[1], three);
        tfm_mma3(st[2 * j2 + 1], kh, kl, bh[2], bh[3], bl[2], bl[3], three);
        ldsm_x4(bh, fl_bt_addr<ROWB>(sb + oGh, 16 * j2, 16 * kk, lane));
        if (three) ldsm_x4(bl, fl_bt_addr<ROWB>(sb + oGl, 16 * j2, 16 * kk, lane));
        tfm_mma3(dpt[2 * j2], vh, vl, bh[0], bh[1], bl[0], bl[1], three);
        tfm_mma3(dpt[2 * j2 + 1], vh, vl, bh[2], bh[3], bl[2], bl[3], three);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int t = 8 * j + 2 * tg + e;  // query of the block (lse2 = +inf beyond S: P = 0)
        const float ls = lse2[t], dt = dd[t];
        float kp0 = c_keep, kp1 = c_keep;
        if (drop_on) {
          const uint32_t wd = bits[t * 4 + ((w0 + g) >> 5)];  // keys w0 + g and + 8 share a word
          if (!((wd >> ((w0 + g) & 31)) & 1u)) kp0 = 0.f;
          if (!((wd >> ((w0 + g + 8) & 31)) & 1u)) kp1 = 0.f;
        }
        const float p0 = kv0 ? ex2_approx(st[j][e] - ls) : 0.f, p1 = kv1 ? ex2_approx(st[j][2 + e] - ls) : 0.f;
        st[j][e] = p0 * (kp0 * dpt[j][e] - dt);
        st[j][2 + e] = p1 * (kp1 * dpt[j][2 + e] - dt);
        dpt[j][e] = p0 * kp0;
        dpt[j][2 + e] = p1 * kp1;
      }
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t sh_[4], sl_[4], ph[4], pl[4];
      split_pack2(st[2 * kk][0], st[2 * kk][1], sh_[0], sl_[0]);
      split_pack2(st[2 * kk][2], st[2 * kk][3], sh_[1], sl_[1]);
      split_pack2(st[2 * kk + 1][0], st[2 * kk + 1][1], sh_[2], sl_[2]);
      split_pack2(st[2 * kk + 1][2], st[2 * kk + 1][3], sh_[3], sl_[3]);
      split_pack2(dpt[2 * kk][0], dpt[2 * kk][1], ph[0], pl[0]);
      split_pack2(dpt[2 * kk][2], dpt[2 * kk][3], ph[1], pl[1]);
      split_pack2(dpt[2 * kk + 1][0], dpt[2 * kk + 1][1], ph[2], pl[2]);
      split_pack2(dpt[2 * kk + 1][2], dpt[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
      for (int jd2 = 0; jd2 < C::DT / 2; ++jd2) {
        uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
        ldsm_x4_t(bh, fl_bn_addr<ROWB>(sb + oQh, 16 * kk, 16 * jd2, lane));
        if (three) ldsm_x4_t(bl, fl_bn_addr<ROWB>(sb + oQl, 16 * kk, 16 * jd2, lane));
        tfm_mma3(dk[2 * jd2], sh_, sl_, bh[0], bh[1], bl[0], bl[1], three);
        tfm_mma3(dk[2 * jd2 + 1], sh_, sl_, bh[2], bh[3], bl[2], bl[3], three);
        ldsm_x4_t(bh, fl_bn_addr<ROWB>(sb + oGh, 16 * kk, 16 * jd2, lane));
        if (three) ldsm_x4_t(bl, fl_bn_addr<ROWB>(sb + oGl, 16 * kk, 16 * jd2, lane));
        tfm_mma3(dv[2 * jd2], ph, pl, bh[0], bh[1], bl[0], bl[1], three);
        tfm_mma3(dv[2 * jd2 + 1], ph, pl, bh[2], bh[3], bl[2], bl[3], three);
      }
    }
  }
  if (!active) return;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int u = k0 + w0 + g + 8 * half;
    if (u >= S) continue;
    const long long rowoff = ((long long)u * seq_stride + row_base) * p3;
#pragma unroll
    for (int j = 0; j < C::DT; ++j) {
      uint32_t hh, ll;
      const long long off = rowoff + h * DH + 8 * j + 2 * tg;
      split_pack2(dk[j][2 * half] * TFM_LN2, dk[j][2 * half + 1] * TFM_LN2, hh, ll);  // staged Q carries log2(e) / sqrt(d_h)
      *reinterpret_cast<uint32_t*>(g_hi + off + E) = hh;
      if (g_lo) *reinterpret_cast<uint32_t*>(g_lo + off + E) = ll;
      split_pack2(dv[j][2 * half], dv[j][2 * half + 1], hh, ll);
      *reinterpret_cast<uint32_t*>(g_hi + off + 2 * E) = hh;
      if (g_lo) *reinterpret_cast<uint32_t*>(g_lo + off + 2 * E) = ll;
    }
  }
}

}  // namespace nrl
