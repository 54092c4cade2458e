// Backward of the short-sequence attention (S <= 32: the 30-token titles of the NRMS news encoder) with the operands
// staged ONCE per (title, head) in shared memory as bf16 hi / lo planes and fetched with ldmatrix -- the layout of the
// transformer attention kernels (nrl_tfm.cuh) at head dims that are not a multiple of 16 (d_h = 20 is padded to 32
// columns of zeros in shared memory; the products over the padding cost MMAs, not memory traffic).
//
// Why a second design beside attn_bwd_mma_kernel (one warp per (title, head), the whole problem in registers,
// transposes by movmatrix): that kernel is instruction-bound at 12 warps per SM -- every fragment is split into hi / lo
// in registers each time it is (re)loaded, 18 % of its instructions are address arithmetic for 80-byte row segments,
// and 168 registers cap the occupancy.  Here the split happens once while staging (all of a thread's 16-byte loads in
// flight together), ldmatrix delivers both orientations, two warps share a head (16 query rows / 16 key rows each) and
// a CTA of HG heads keeps 18 warps per SM resident.
//   phase A (rows = queries)  P = exp(s - lse), dP = dO V^T, D = rowsum(P dP), dS = P (dP - D), dQ = dS K / sqrt(d_h)
//   phase B (rows = keys)     the transposed tiles K Q^T, V dO^T:  dV = P^T dO,  dK = dS^T Q / sqrt(d_h)
#pragma once
#include "nrl_tfm.cuh"

namespace nrl {

template <int DH, int NK32 = 1>  // sequence padded to SK = 32 * NK32 rows (S <= 32: titles; S <= 64: the NRMS user encoder)
struct TitleCfg {
  static constexpr int SK = 32 * NK32;
  static constexpr int DHP = (DH + 15) / 16 * 16;  // staged columns (zeros beyond DH)
  static constexpr int PITCH = DHP + 8;
  static constexpr int ROWB = PITCH * 2;
  static constexpr int KS = DHP / 16;
  static constexpr int DT = (DH + 7) / 8;          // 8-wide output tiles that hold real columns
  static constexpr int NT = SK / 8;                // 8-wide tiles along the sequence
  static constexpr int PLANE = SK * ROWB;
  static constexpr int HEAD_BYTES = 8 * PLANE + 2 * SK * 4;  // Q K V dO (hi, lo) + lse2[SK] + dd[SK]
  static constexpr int HEAD_THREADS = 32 * (SK / 16);        // one warp per 16 rows
};
template <int ROWB>
__device__ __forceinline__ uint32_t tt_a_addr(uint32_t base, int row0, int k0, int lane) {
  const int mi = lane >> 3;
  return base + (uint32_t)((row0 + (lane & 7) + 8 * (mi & 1)) * ROWB + (k0 + 8 * (mi >> 1)) * 2);
}
template <int ROWB>
__device__ __forceinline__ uint32_t tt_bt_addr(uint32_t base, int n0, int k0, int lane) {
  const int mi = lane >> 3;
  return base + (uint32_t)((n0 + (lane & 7) + 8 * (mi >> 1)) * ROWB + (k0 + 8 * (mi & 1)) * 2);
}
template <int ROWB>
__device__ __forceinline__ uint32_t tt_bn_addr(uint32_t base, int k0, int n0, int lane) {
  const int mi = lane >> 3;
  return base + (uint32_t)((k0 + (lane & 7) + 8 * (mi & 1)) * ROWB + (n0 + 8 * (mi >> 1)) * 2);
}

template <int DH, int HG, int NK32>
__global__ void __launch_bounds__(64 * NK32 * HG, NK32 == 1 ? 3 : 2)
attn_bwd_ldsm_kernel(const float* __restrict__ qkv, const float* __restrict__ d_o, long long ld_do,
                     const float* __restrict__ lse, int E, int ldq, int heads, int S, long long seq_stride, int NB,
                     long long batch_stride, float scale, __nv_bfloat16* __restrict__ g_hi,
                     __nv_bfloat16* __restrict__ g_lo, int p3, int three_i) {
  pdl_launch_dependents();  // the next launch (a PDL-launched GEMM) may begin its prologue while this grid runs
  using C = TitleCfg<DH, NK32>;
  constexpr int ROWB = C::ROWB, C4 = C::DHP / 4, R4 = DH / 4, SK = C::SK, NT = C::NT, HT = C::HEAD_THREADS;
  static_assert(DH % 4 == 0, "16-byte row segments");
  extern __shared__ __align__(16) uint8_t ttsm[];
  const bool three = three_i != 0;
  const int groups = (heads + HG - 1) / HG;
  const int b = blockIdx.x / groups, hl = threadIdx.x / HT, h = (blockIdx.x % groups) * HG + hl;
  const bool head_ok = h < heads;
  const int t2 = threadIdx.x % HT;  // thread within the head's warps
  uint8_t* hs = ttsm + hl * C::HEAD_BYTES;
  float* lse2 = reinterpret_cast<float*>(hs + 8 * C::PLANE);
  float* dd = lse2 + SK;
  const long long row_base = (long long)b * batch_stride;
  // ---- stage Q (pre-scaled), K, V, dO: 4 matrices x 32 rows x C4 chunks, two batches of loads in flight ----
  if (head_ok) {
    constexpr int ITEMS = 4 * SK * C4, ITER = ITEMS / HT, HALF = ITER / 2;
    static_assert(ITER * HT == ITEMS && HALF * 2 == ITER, "staging split");
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 v[HALF];
#pragma unroll
      for (int it = 0; it < HALF; ++it) {
        const int i = t2 + (half * HALF + it) * HT, m = i / (SK * C4), r = (i / C4) % SK, c = i % C4;
        v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < S && c < R4) {
          const long long grow = (long long)r * seq_stride + row_base;
          const float* src = m < 3 ? qkv + grow * ldq + m * E + h * DH : d_o + grow * ld_do + h * DH;
          v[it] = __ldg(reinterpret_cast<const float4*>(src) + c);
        }
      }
#pragma unroll
      for (int it = 0; it < HALF; ++it) {
        const int i = t2 + (half * HALF + it) * HT, m = i / (SK * C4), r = (i / C4) % SK, c = i % C4;
        const float mul = m == 0 ? scale * TFM_LOG2E : 1.f;
        uint32_t h0, l0, h1, l1;
        split_pack2(v[it].x * mul, v[it].y * mul, h0, l0);
        split_pack2(v[it].z * mul, v[it].w * mul, h1, l1);
        const uint32_t off = (uint32_t)(2 * m * C::PLANE + r * ROWB + c * 8);
        *reinterpret_cast<uint2*>(hs + off) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(hs + off + C::PLANE) = make_uint2(l0, l1);
      }
    }
    if (t2 < SK) {
      lse2[t2] = t2 < S ? lse[((long long)t2 * seq_stride + row_base) * heads + h] * TFM_LOG2E : INFINITY;
      dd[t2] = 0.f;  // rows of a warp that has nothing to do (S <= 16) are still read by phase B
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int w0 = 16 * (t2 >> 5);
  const bool active = head_ok && w0 < S;
  const uint32_t sb = smem_u32(hs);
  const uint32_t oQh = 0, oQl = C::PLANE, oKh = 2 * C::PLANE, oKl = 3 * C::PLANE, oVh = 4 * C::PLANE, oVl = 5 * C::PLANE,
                 oGh = 6 * C::PLANE, oGl = 7 * C::PLANE;
  const int r0 = w0 + g, r1 = r0 + 8;
  // =========================== phase A: D and dQ (rows = queries) ===========================
  if (active) {
    float s[NT][4], dp[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      dp[j][0] = dp[j][1] = dp[j][2] = dp[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < C::KS; ++kk) {
      uint32_t qh[4], ql[4] = {0u, 0u, 0u, 0u}, gh[4], gl[4] = {0u, 0u, 0u, 0u};
      ldsm_x4(qh, tt_a_addr<ROWB>(sb + oQh, w0, 16 * kk, lane));
      ldsm_x4(gh, tt_a_addr<ROWB>(sb + oGh, w0, 16 * kk, lane));
      if (three) {
        ldsm_x4(ql, tt_a_addr<ROWB>(sb + oQl, w0, 16 * kk, lane));
        ldsm_x4(gl, tt_a_addr<ROWB>(sb + oGl, w0, 16 * kk, lane));
      }
#pragma unroll
      for (int j2 = 0; j2 < NT / 2; ++j2) {
        uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
        ldsm_x4(bh, tt_bt_addr<ROWB>(sb + oKh, 16 * j2, 16 * kk, lane));
        if (three) ldsm_x4(bl, tt_bt_addr<ROWB>(sb + oKl, 16 * j2, 16 * kk, lane));
        tfm_mma3(s[2 * j2], qh, ql, bh[0], bh[1], bl[0], bl[1], three);
        tfm_mma3(s[2 * j2 + 1], qh, ql, bh[2], bh[3], bl[2], bl[3], three);
        ldsm_x4(bh, tt_bt_addr<ROWB>(sb + oVh, 16 * j2, 16 * kk, lane));
        if (three) ldsm_x4(bl, tt_bt_addr<ROWB>(sb + oVl, 16 * j2, 16 * kk, lane));
        tfm_mma3(dp[2 * j2], gh, gl, bh[0], bh[1], bl[0], bl[1], three);
        tfm_mma3(dp[2 * j2 + 1], gh, gl, bh[2], bh[3], bl[2], bl[3], three);
      }
    }
    const float ls0 = lse2[r0], ls1 = lse2[r1];
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int u = 8 * j + 2 * tg;
      s[j][0] = u < S ? ex2_approx(s[j][0] - ls0) : 0.f; s[j][1] = u + 1 < S ? ex2_approx(s[j][1] - ls0) : 0.f;
      s[j][2] = u < S ? ex2_approx(s[j][2] - ls1) : 0.f; s[j][3] = u + 1 < S ? ex2_approx(s[j][3] - ls1) : 0.f;
      d0 += s[j][0] * dp[j][0] + s[j][1] * dp[j][1];
      d1 += s[j][2] * dp[j][2] + s[j][3] * dp[j][3];
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    if (tg == 0) { dd[r0] = d0; dd[r1] = d1; }
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      s[j][0] *= dp[j][0] - d0; s[j][1] *= dp[j][1] - d0;
      s[j][2] *= dp[j][2] - d1; s[j][3] *= dp[j][3] - d1;
    }
    float dq[C::DT][4];
#pragma unroll
    for (int j = 0; j < C::DT; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {  // all keys
      uint32_t ah[4], al[4];
      split_pack2(s[2 * kk][0], s[2 * kk][1], ah[0], al[0]);
      split_pack2(s[2 * kk][2], s[2 * kk][3], ah[1], al[1]);
      split_pack2(s[2 * kk + 1][0], s[2 * kk + 1][1], ah[2], al[2]);
      split_pack2(s[2 * kk + 1][2], s[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
      for (int jd2 = 0; jd2 < (C::DT + 1) / 2; ++jd2) {
        uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
        ldsm_x4_t(bh, tt_bn_addr<ROWB>(sb + oKh, 16 * kk, 16 * jd2, lane));
        if (three) ldsm_x4_t(bl, tt_bn_addr<ROWB>(sb + oKl, 16 * kk, 16 * jd2, lane));
        tfm_mma3(dq[2 * jd2], ah, al, bh[0], bh[1], bl[0], bl[1], three);
        if (2 * jd2 + 1 < C::DT) tfm_mma3(dq[2 * jd2 + 1], ah, al, bh[2], bh[3], bl[2], bl[3], three);
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = half ? r1 : r0;
      if (r >= S) continue;
      const long long rowoff = ((long long)r * seq_stride + row_base) * p3;
#pragma unroll
      for (int j = 0; j < C::DT; ++j) {
        const int c = 8 * j + 2 * tg;
        if (c >= DH) continue;
        uint32_t hh, ll;
        split_pack2(dq[j][2 * half] * scale, dq[j][2 * half + 1] * scale, hh, ll);
        const long long off = rowoff + h * DH + c;
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
  __syncthreads();  // D of all queries of the head is in shared memory
  // =========================== phase B: dK, dV (rows = keys) ===========================
  if (active) {
    float st[NT][4], dpt[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
      dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < C::KS; ++kk) {
      uint32_t kh[4], kl[4] = {0u, 0u, 0u, 0u}, vh[4], vl[4] = {0u, 0u, 0u, 0u};
      ldsm_x4(kh, tt_a_addr<ROWB>(sb + oKh, w0, 16 * kk, lane));
      ldsm_x4(vh, tt_a_addr<ROWB>(sb + oVh, w0, 16 * kk, lane));
      if (three) {
        ldsm_x4(kl, tt_a_addr<ROWB>(sb + oKl, w0, 16 * kk, lane));
        ldsm_x4(vl, tt_a_addr<ROWB>(sb + oVl, w0, 16 * kk, lane));
      }
#pragma unroll
      for (int j2 = 0; j2 < NT / 2; ++j2) {
        uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
        ldsm_x4(bh, tt_bt_addr<ROWB>(sb + oQh, 16 * j2, 16 * kk, lane));
        if (three) ldsm_x4(bl, tt_bt_addr<ROWB>(sb + oQl, 16 * j2, 16 * kk, lane));
        tfm_mma3(st[2 * j2], kh, kl, bh[0], bh[1], bl[0], bl[1], three);
        tfm_mma3(st[2 * j2 + 1], kh, kl, bh[2], bh[3], bl[2], bl[3], three);
        ldsm_x4(bh, tt_bt_addr<ROWB>(sb + oGh, 16 * j2, 16 * kk, lane));
        if (three) ldsm_x4(bl, tt_bt_addr<ROWB>(sb + oGl, 16 * j2, 16 * kk, lane));
        tfm_mma3(dpt[2 * j2], vh, vl, bh[0], bh[1], bl[0], bl[1], three);
        tfm_mma3(dpt[2 * j2 + 1], vh, vl, bh[2], bh[3], bl[2], bl[3], three);
      }
    }
    const bool kv0 = r0 < S, kv1 = r1 < S;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int t = 8 * j + 2 * tg + e;  // query (lse2 = +inf beyond S: P = 0)
        const float ls = lse2[t], dt = dd[t];
        const float p0 = kv0 ? ex2_approx(st[j][e] - ls) : 0.f, p1 = kv1 ? ex2_approx(st[j][2 + e] - ls) : 0.f;
        st[j][e] = p0 * (dpt[j][e] - dt);
        st[j][2 + e] = p1 * (dpt[j][2 + e] - dt);
        dpt[j][e] = p0;
        dpt[j][2 + e] = p1;
      }
    }
    float dk[C::DT][4], dv[C::DT][4];
#pragma unroll
    for (int j = 0; j < C::DT; ++j) {
      dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
      dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < NT / 2; ++kk) {  // all queries
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
      for (int jd2 = 0; jd2 < (C::DT + 1) / 2; ++jd2) {
        uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
        ldsm_x4_t(bh, tt_bn_addr<ROWB>(sb + oQh, 16 * kk, 16 * jd2, lane));
        if (three) ldsm_x4_t(bl, tt_bn_addr<ROWB>(sb + oQl, 16 * kk, 16 * jd2, lane));
        tfm_mma3(dk[2 * jd2], sh_, sl_, bh[0], bh[1], bl[0], bl[1], three);
        if (2 * jd2 + 1 < C::DT) tfm_mma3(dk[2 * jd2 + 1], sh_, sl_, bh[2], bh[3], bl[2], bl[3], three);
        ldsm_x4_t(bh, tt_bn_addr<ROWB>(sb + oGh, 16 * kk, 16 * jd2, lane));
        if (three) ldsm_x4_t(bl, tt_bn_addr<ROWB>(sb + oGl, 16 * kk, 16 * jd2, lane));
        tfm_mma3(dv[2 * jd2], ph, pl, bh[0], bh[1], bl[0], bl[1], three);
        if (2 * jd2 + 1 < C::DT) tfm_mma3(dv[2 * jd2 + 1], ph, pl, bh[2], bh[3], bl[2], bl[3], three);
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int u = half ? r1 : r0;
      if (u >= S) continue;
      const long long rowoff = ((long long)u * seq_stride + row_base) * p3;
#pragma unroll
      for (int j = 0; j < C::DT; ++j) {
        const int c = 8 * j + 2 * tg;
        if (c >= DH) continue;
        uint32_t hh, ll;
        const long long off = rowoff + h * DH + c;
        split_pack2(dk[j][2 * half] * TFM_LN2, dk[j][2 * half + 1] * TFM_LN2, hh, ll);  // staged Q carries log2(e) / sqrt(d_h)
        *reinterpret_cast<uint32_t*>(g_hi + off + E) = hh;
        if (g_lo) *reinterpret_cast<uint32_t*>(g_lo + off + E) = ll;
        split_pack2(dv[j][2 * half], dv[j][2 * half + 1], hh, ll);
        *reinterpret_cast<uint32_t*>(g_hi + off + 2 * E) = hh;
        if (g_lo) *reinterpret_cast<uint32_t*>(g_lo + off + 2 * E) = ll;
      }
    }
  }
}

}  // namespace nrl
