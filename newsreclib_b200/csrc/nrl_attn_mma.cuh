// Tensor-core self-attention for short sequences (S <= 32: the 30-token titles), sm_100a.
//
// One warp owns one (batch item, head): the whole problem -- Q, K, V (and dO) slices of
// [S <= 32][DH <= 32] -- lives in the warp's registers as bf16 hi/lo fragment blocks, and every
// contraction (S = QK^T, O = PV; backward: dP = dO V^T, dQ = dS K, dK = dS^T Q, dV = P^T dO)
// is issued as warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with the same three-pass
// hi/lo scheme as the GEMMs (lo*hi + hi*lo + hi*hi), i.e. fp32-equivalent results.  A per-head
// problem is 30 x 30 x 20: far below the tcgen05 tile (M >= 64, operands through shared-memory
// descriptors, accumulator in TMEM), which would waste >= 75 % of the issued MMA work on
// block-diagonal padding and need TMEM round trips for the softmax; the warp-level MMA is the
// tensor instruction whose shape fits, and keeps softmax / dS in the accumulator registers.
//
// Fragment bookkeeping.  A [32 x 32] operand is held as 4 x 4 "blocks" of 8 x 8 elements, one
// 32-bit register per block per plane, in the canonical row layout
//     lane (g = lane >> 2, tg = lane & 3) holds  M[8 rb + g][8 cb + 2 tg, 8 cb + 2 tg + 1]
// which is simultaneously (a) a quarter of the m16n8k16 A fragment, (b) half of the B fragment
// of the TRANSPOSED operand (B[k][n] = M[n][k]), and (c) the accumulator layout.  Operands that
// are needed with the other orientation (V in PV, K in dS K, Q / dO in the transposed products,
// P^T / dS^T) are produced with movmatrix.trans on those registers: nothing is re-read from
// memory and nothing goes through shared memory.
#pragma once
#include "nrl_kernels.cuh"

namespace nrl {

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2,
                                          uint32_t a3, uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
      "{%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// same with C = 0: the first product of an accumulator tile needs no zeroed registers
__device__ __forceinline__ void mma_16816_z(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                            uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, "
      "{%8, %9}, {%10, %10, %10, %10};"
      : "=f"(c[0]), "=f"(c[1]), "=f"(c[2]), "=f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.f));
}
template <bool Z>
__device__ __forceinline__ void mma_16816_t(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                            uint32_t b0, uint32_t b1) {
  if (Z) mma_16816_z(c, a0, a1, a2, a3, b0, b1);
  else mma_16816(c, a0, a1, a2, a3, b0, b1);
}
__device__ __forceinline__ uint32_t movm_t(uint32_t x) {
  uint32_t y;
  asm("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
// (x, y) -> packed bf16x2 hi (x in the low half) and the bf16x2 of the residuals
__device__ __forceinline__ void split_pack2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// A [32 x 32] operand as 4 x 4 blocks, hi and lo planes.
struct Blk {
  uint32_t h[4][4], l[4][4];
};

// Load rows [0, S) x cols [0, DH) of a row-major fp32 slice (row r at base + r * row_stride
// floats), scaled, into row-layout blocks; everything outside is zero.
template <int DH>
__device__ __forceinline__ void load_blocks(Blk& m, const float* __restrict__ base, long long row_stride,
                                            int S, float mul, int g, int tg) {
#pragma unroll
  for (int rb = 0; rb < 4; ++rb) {
    const int r = 8 * rb + g;
    const bool rok = r < S;
    const float* rp = base + (long long)(rok ? r : 0) * row_stride;
#pragma unroll
    for (int cb = 0; cb < 4; ++cb) {
      if (8 * cb >= DH) {  // compile-time: no such columns
        m.h[rb][cb] = 0u; m.l[rb][cb] = 0u;
        continue;
      }
      const int c = 8 * cb + 2 * tg;
      float2 v = make_float2(0.f, 0.f);
      if (rok && c < DH) v = __ldg(reinterpret_cast<const float2*>(rp + c));
      split_pack2(v.x * mul, v.y * mul, m.h[rb][cb], m.l[rb][cb]);
    }
  }
}

// The three passes of a k-step are issued pass-major (all (i, j) tiles of one pass, then the next
// pass): consecutive HMMAs then hit different accumulators instead of forming a dependent chain of
// three on the same tile.
// c[i][j] += sum over k-steps  A(i, kk) * B(j, kk)   with A blocks a[2i + ..][2kk + ..] (row layout,
// M x K) and B given as the row-layout blocks of the [N x K] operand b[j][2kk + ..] ("NT" product).
// INIT: the accumulator tiles c[i][j], j < NT, are (re)initialised by the first product (no zeroing needed; the
// hi*hi pass goes first then, so that the same instruction initialises in both precisions).
template <int KSTEPS, int NT, bool INIT = false>
__device__ __forceinline__ void mma_nt(float (&c)[2][4][4], const Blk& a, const Blk& b, bool three) {
#pragma unroll
  for (int kk = 0; kk < KSTEPS; ++kk) {
    if (INIT && kk == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
          mma_16816_z(c[i][j], a.h[2 * i][0], a.h[2 * i + 1][0], a.h[2 * i][1], a.h[2 * i + 1][1], b.h[j][0], b.h[j][1]);
    }
    if (three) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
          mma_16816(c[i][j], a.l[2 * i][2 * kk], a.l[2 * i + 1][2 * kk], a.l[2 * i][2 * kk + 1],
                    a.l[2 * i + 1][2 * kk + 1], b.h[j][2 * kk], b.h[j][2 * kk + 1]);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
          mma_16816(c[i][j], a.h[2 * i][2 * kk], a.h[2 * i + 1][2 * kk], a.h[2 * i][2 * kk + 1],
                    a.h[2 * i + 1][2 * kk + 1], b.l[j][2 * kk], b.l[j][2 * kk + 1]);
    }
    if (!(INIT && kk == 0)) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
          mma_16816(c[i][j], a.h[2 * i][2 * kk], a.h[2 * i + 1][2 * kk], a.h[2 * i][2 * kk + 1],
                    a.h[2 * i + 1][2 * kk + 1], b.h[j][2 * kk], b.h[j][2 * kk + 1]);
    }
  }
}
// c[i][j] += A(i, kk) * B(kk, j) with B given as row-layout blocks of the [K x N] operand
// b[2kk + ..][j] ("NN" product: the B fragments are the movmatrix transposes of those blocks).
template <int NT, bool INIT = false>
__device__ __forceinline__ void mma_nn(float (&c)[2][4][4], const Blk& a, const Blk& b, bool three) {
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      bh[j][0] = movm_t(b.h[2 * kk][j]); bh[j][1] = movm_t(b.h[2 * kk + 1][j]);
      bl[j][0] = three ? movm_t(b.l[2 * kk][j]) : 0u; bl[j][1] = three ? movm_t(b.l[2 * kk + 1][j]) : 0u;
    }
    if (INIT && kk == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
          mma_16816_z(c[i][j], a.h[2 * i][0], a.h[2 * i + 1][0], a.h[2 * i][1], a.h[2 * i + 1][1], bh[j][0], bh[j][1]);
    }
    if (three) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
          mma_16816(c[i][j], a.l[2 * i][2 * kk], a.l[2 * i + 1][2 * kk], a.l[2 * i][2 * kk + 1],
                    a.l[2 * i + 1][2 * kk + 1], bh[j][0], bh[j][1]);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
          mma_16816(c[i][j], a.h[2 * i][2 * kk], a.h[2 * i + 1][2 * kk], a.h[2 * i][2 * kk + 1],
                    a.h[2 * i + 1][2 * kk + 1], bl[j][0], bl[j][1]);
    }
    if (!(INIT && kk == 0)) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
          mma_16816(c[i][j], a.h[2 * i][2 * kk], a.h[2 * i + 1][2 * kk], a.h[2 * i][2 * kk + 1],
                    a.h[2 * i + 1][2 * kk + 1], bh[j][0], bh[j][1]);
    }
  }
}
// c[i][j] += sum_t X[t][16 i + ..] * Y[t][8 j + ..]  ("TN" product, reduction over the ROWS of both
// operands): A fragments are movmatrix transposes of X's blocks, B fragments those of Y's blocks,
// formed just in time so that no transposed copy of X stays live.
template <int NT, bool INIT = false>
__device__ __forceinline__ void mma_tn(float (&c)[2][4][4], const Blk& x, const Blk& y, bool three) {
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      bh[j][0] = movm_t(y.h[2 * kk][j]); bh[j][1] = movm_t(y.h[2 * kk + 1][j]);
      bl[j][0] = three ? movm_t(y.l[2 * kk][j]) : 0u; bl[j][1] = three ? movm_t(y.l[2 * kk + 1][j]) : 0u;
    }
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      ah[i][0] = movm_t(x.h[2 * kk][2 * i]); ah[i][1] = movm_t(x.h[2 * kk][2 * i + 1]);
      ah[i][2] = movm_t(x.h[2 * kk + 1][2 * i]); ah[i][3] = movm_t(x.h[2 * kk + 1][2 * i + 1]);
#pragma unroll
      for (int r = 0; r < 4; ++r) al[i][r] = 0u;
      if (three) {
        al[i][0] = movm_t(x.l[2 * kk][2 * i]); al[i][1] = movm_t(x.l[2 * kk][2 * i + 1]);
        al[i][2] = movm_t(x.l[2 * kk + 1][2 * i]); al[i][3] = movm_t(x.l[2 * kk + 1][2 * i + 1]);
      }
    }
    if (INIT && kk == 0) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) mma_16816_z(c[i][j], ah[i][0], ah[i][1], ah[i][2], ah[i][3], bh[j][0], bh[j][1]);
    }
    if (three) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) mma_16816(c[i][j], al[i][0], al[i][1], al[i][2], al[i][3], bh[j][0], bh[j][1]);
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) mma_16816(c[i][j], ah[i][0], ah[i][1], ah[i][2], ah[i][3], bl[j][0], bl[j][1]);
    }
    if (!(INIT && kk == 0)) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j) mma_16816(c[i][j], ah[i][0], ah[i][1], ah[i][2], ah[i][3], bh[j][0], bh[j][1]);
    }
  }
}
// Accumulator tile set c[2][4][4] (rows 16 i + g (+8), cols 8 j + 2 tg (+1)) -> row-layout blocks.
__device__ __forceinline__ void acc_to_blocks(Blk& m, const float (&c)[2][4][4]) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      split_pack2(c[i][j][0], c[i][j][1], m.h[2 * i][j], m.l[2 * i][j]);
      split_pack2(c[i][j][2], c[i][j][3], m.h[2 * i + 1][j], m.l[2 * i + 1][j]);
    }
}
__device__ __forceinline__ void transpose_blocks(Blk& t, const Blk& m, bool three) {
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      t.h[c][r] = movm_t(m.h[r][c]);
      t.l[c][r] = three ? movm_t(m.l[r][c]) : 0u;
    }
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
  return v;
}
// Store accumulator tiles c[2][NT][4] (cols < DH, rows < S) as split planes at dst(row) + col.
template <int DH, int NT>
__device__ __forceinline__ void store_acc_split(const float (&c)[2][4][4], float mul0, float mul1,
                                                float mul2, float mul3, __nv_bfloat16* __restrict__ hi,
                                                __nv_bfloat16* __restrict__ lo, long long pitch,
                                                const long long (&grow)[4], int S, int col0, int g, int tg) {
  // mul[rb] scales rows 8 rb + g
  const float mul[4] = {mul0, mul1, mul2, mul3};
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int rb = 2 * i + half;
      if (8 * rb + g >= S) continue;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int col = 8 * j + 2 * tg;
        if (col >= DH) continue;
        uint32_t h, l;
        split_pack2(c[i][j][2 * half] * mul[rb], c[i][j][2 * half + 1] * mul[rb], h, l);
        const long long off = grow[rb] * pitch + col0 + col;
        *reinterpret_cast<uint32_t*>(hi + off) = h;
        if (lo) *reinterpret_cast<uint32_t*>(lo + off) = l;
      }
    }
}

// ---- operand sources ------------------------------------------------------------------------
// Global: matrix m of the item lives at base[m] + r * stride[m] (fp32 rows); rows >= S read as 0.
template <int DH, bool RELOAD = false>
struct GmemSrc {
  static constexpr bool kReload = RELOAD;  // re-read operand fragments (L1 / L2 hits) instead of keeping them live
  const float* base[4];
  long long stride[4];
  int S;
  __device__ __forceinline__ void load(Blk& m, int which, float mul, int g, int tg) const {
    load_blocks<DH>(m, base[which], stride[which], S, mul, g, tg);
  }
  __device__ __forceinline__ float lse2(const float* lse, long long grow, int heads, int h, int /*r*/) const {
    return __ldg(lse + grow * heads + h) * NRL_LOG2E;
  }
};
// ---- one (batch item, head) problem, forward ----------------------------------------------------
struct NoRelease {
  __device__ __forceinline__ void operator()() const {}
};
// `release()` is called once, right after the last operand fragment has been read from `src` (the staged kernels
// hand the shared-memory stage back to the TMA producer there).
template <int DH, class Src, class Release = NoRelease>
__device__ __forceinline__ void attn_fwd_item(const Src& src, int E, int heads, int S, long long seq_stride,
                                              long long batch_stride, float scale, int b, int h,
                                              __nv_bfloat16* __restrict__ o_hi, __nv_bfloat16* __restrict__ o_lo,
                                              int ep, float* __restrict__ lse, int lane, Release release = Release()) {
  constexpr int KS = (DH + 15) / 16;  // k-steps over the head dim
  constexpr int ND = (DH + 7) / 8;    // 8-column blocks of the head dim
  const int g = lane >> 2, tg = lane & 3;
  const bool three = o_lo != nullptr;
  Blk q, k;
  src.load(q, 0, scale * NRL_LOG2E, g, tg);
  src.load(k, 1, 1.f, g, tg);
  float s[2][4][4];
  mma_nt<KS, 4, true>(s, q, k, three);

  // softmax over the key axis (columns); rows 8 rb + g, rb = 2 i + half
  float inv_l[4], row_lse[4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = 8 * j + 2 * tg + c;
          if (col >= S) s[i][j][2 * half + c] = -INFINITY;
          m = fmaxf(m, s[i][j][2 * half + c]);
        }
      m = quad_max(m);
      float l = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float p = ex2_approx(s[i][j][2 * half + c] - m);
          s[i][j][2 * half + c] = p;
          l += p;
        }
      l = quad_sum(l);
      inv_l[2 * i + half] = 1.f / l;
      row_lse[2 * i + half] = m * NRL_LN2 + logf(l);
    }
  Blk p;
  acc_to_blocks(p, s);
  Blk v;
  src.load(v, 2, 1.f, g, tg);
  release();
  float o[2][4][4];
  mma_nn<ND, true>(o, p, v, three);

  long long grow[4];
#pragma unroll
  for (int rb = 0; rb < 4; ++rb) grow[rb] = (long long)(8 * rb + g) * seq_stride + (long long)b * batch_stride;
  store_acc_split<DH, ND>(o, inv_l[0], inv_l[1], inv_l[2], inv_l[3], o_hi, o_lo, ep, grow, S, h * DH, g, tg);
  if (tg == 0) {
#pragma unroll
    for (int rb = 0; rb < 4; ++rb)
      if (8 * rb + g < S) lse[grow[rb] * heads + h] = row_lse[rb];
  }
  if (h == 0) {  // pad columns of the plane rows: ones column at E, zeros after
    const int npad = ep - E;
    for (int i = lane; i < S * npad; i += 32) {
      const int srow = i / npad, c = E + i % npad;
      const long long gr = (long long)srow * seq_stride + (long long)b * batch_stride;
      o_hi[gr * ep + c] = __float2bfloat16_rn(c == E ? 1.f : 0.f);
      if (o_lo) o_lo[gr * ep + c] = __float2bfloat16_rn(0.f);
    }
  }
}

// ---- one (batch item, head) problem, backward ---------------------------------------------------
// P is recomputed from Q, K and the saved log-sum-exp; D = rowsum(P * dP).  Writes dQ | dK | dV as
// split planes [2][R][p3].
template <int DH, class Src, class Release = NoRelease>
__device__ __forceinline__ void attn_bwd_item(const Src& src, const float* __restrict__ lse, int E, int heads,
                                              int S, long long seq_stride, long long batch_stride, float scale,
                                              int b, int h, __nv_bfloat16* __restrict__ g_hi,
                                              __nv_bfloat16* __restrict__ g_lo, int p3, int lane,
                                              Release release = Release()) {
  constexpr int KS = (DH + 15) / 16;
  constexpr int ND = (DH + 7) / 8;
  const int g = lane >> 2, tg = lane & 3;
  const bool three = g_lo != nullptr;
  long long grow[4];
#pragma unroll
  for (int rb = 0; rb < 4; ++rb) grow[rb] = (long long)(8 * rb + g) * seq_stride + (long long)b * batch_stride;

  // Src::kReload (operands staged in shared memory): fragments are re-read where they are needed again instead
  // of being kept in registers across the whole item (four operand blocks = 128 registers), and the stage is
  // released at the end; otherwise (operands in global memory) everything is loaded once, up front.
  constexpr bool kReload = Src::kReload;
  Blk q, k, v, go;
  src.load(q, 0, scale * NRL_LOG2E, g, tg);  // Qs = Q * scale * log2(e)
  src.load(k, 1, 1.f, g, tg);
  float lse2[4];
#pragma unroll
  for (int rb = 0; rb < 4; ++rb) lse2[rb] = (8 * rb + g < S) ? src.lse2(lse, grow[rb], heads, h, 8 * rb + g) : 0.f;
  float s[2][4][4], dp[2][4][4];
  if (!kReload) {
    src.load(v, 2, 1.f, g, tg);
    src.load(go, 3, 1.f, g, tg);
    release();
  }
  mma_nt<KS, 4, true>(s, q, k, three);    // S (log2 domain)  [t][u]
  if (kReload) {
    src.load(v, 2, 1.f, g, tg);
    src.load(go, 3, 1.f, g, tg);
  }
  mma_nt<KS, 4, true>(dp, go, v, three);  // dP = dO V^T      [t][u]

  // P = 2^(S - lse2), D_t = sum_u P dP, dS = P (dP - D)   (natural-domain gradient of the scaled scores)
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int rb = 2 * i + half;
      const bool rok = 8 * rb + g < S;
      float dd = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = 8 * j + 2 * tg + c;
          const float p = (rok && col < S) ? ex2_approx(s[i][j][2 * half + c] - lse2[rb]) : 0.f;
          s[i][j][2 * half + c] = p;
          dd += p * dp[i][j][2 * half + c];
        }
      dd = quad_sum(dd);
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c)
          dp[i][j][2 * half + c] = s[i][j][2 * half + c] * (dp[i][j][2 * half + c] - dd);
    }
  Blk pb, ds;
  acc_to_blocks(pb, s);
  acc_to_blocks(ds, dp);

  float acc[2][4][4];
  // dQ = scale * dS K
  if (kReload) src.load(k, 1, 1.f, g, tg);
  mma_nn<ND, true>(acc, ds, k, three);
  store_acc_split<DH, ND>(acc, scale, scale, scale, scale, g_hi, g_lo, p3, grow, S, h * DH, g, tg);
  // dK = scale * dS^T Q = ln2 * dS^T Qs ;  dV = P^T dO
  if (kReload) src.load(q, 0, scale * NRL_LOG2E, g, tg);
  mma_tn<ND, true>(acc, ds, q, three);
  store_acc_split<DH, ND>(acc, NRL_LN2, NRL_LN2, NRL_LN2, NRL_LN2, g_hi, g_lo, p3, grow, S, E + h * DH, g, tg);
  if (kReload) {
    src.load(go, 3, 1.f, g, tg);
    release();
  }
  mma_tn<ND, true>(acc, pb, go, three);
  store_acc_split<DH, ND>(acc, 1.f, 1.f, 1.f, 1.f, g_hi, g_lo, p3, grow, S, 2 * E + h * DH, g, tg);

  if (h == 0 && p3 > 3 * E) {
    const int npad = p3 - 3 * E;
    for (int i = lane; i < S * npad; i += 32) {
      const int srow = i / npad, c = 3 * E + i % npad;
      const long long gr = (long long)srow * seq_stride + (long long)b * batch_stride;
      g_hi[gr * p3 + c] = __float2bfloat16_rn(0.f);
      if (g_lo) g_lo[gr * p3 + c] = __float2bfloat16_rn(0.f);
    }
  }
}

// ---- direct kernels: one warp = one item, operands straight from global memory -------------------
template <int DH>
__global__ void __launch_bounds__(128)
attn_fwd_mma_kernel(const float* __restrict__ qkv, int E, int ldq, int heads, int S, long long seq_stride,
                    int NB, long long batch_stride, float scale, __nv_bfloat16* __restrict__ o_hi,
                    __nv_bfloat16* __restrict__ o_lo, int ep, float* __restrict__ lse) {
  const int lane = threadIdx.x & 31;
  const long long item = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (item >= (long long)NB * heads) return;
  const int b = (int)(item / heads), h = (int)(item % heads);
  const float* base = qkv + (long long)b * batch_stride * ldq + h * DH;
  GmemSrc<DH> src;
  src.S = S;
#pragma unroll
  for (int m = 0; m < 3; ++m) { src.base[m] = base + m * E; src.stride[m] = seq_stride * ldq; }
  attn_fwd_item<DH>(src, E, heads, S, seq_stride, batch_stride, scale, b, h, o_hi, o_lo, ep, lse, lane);
}

// RELOAD: operand fragments are re-read from global memory (L1 / L2 hits) where they are needed again instead of being
// held in registers across the item; MINB: CTAs per SM the register budget is cut for.
template <int DH, bool RELOAD = false, int MINB = 3>
__global__ void __launch_bounds__(128, MINB)
attn_bwd_mma_kernel(const float* __restrict__ qkv, const float* __restrict__ d_o, long long ld_do,
                    const float* __restrict__ lse, int E, int ldq, int heads, int S, long long seq_stride,
                    int NB, long long batch_stride, float scale, __nv_bfloat16* __restrict__ g_hi,
                    __nv_bfloat16* __restrict__ g_lo, int p3) {
  const int lane = threadIdx.x & 31;
  const long long item = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (item >= (long long)NB * heads) return;
  const int b = (int)(item / heads), h = (int)(item % heads);
  const float* base = qkv + (long long)b * batch_stride * ldq + h * DH;
  GmemSrc<DH, RELOAD> src;
  src.S = S;
#pragma unroll
  for (int m = 0; m < 3; ++m) { src.base[m] = base + m * E; src.stride[m] = seq_stride * ldq; }
  src.base[3] = d_o + (long long)b * batch_stride * ld_do + h * DH;
  src.stride[3] = seq_stride * ld_do;
  attn_bwd_item<DH>(src, lse, E, heads, S, seq_stride, batch_stride, scale, b, h, g_hi, g_lo, p3, lane);
}

// ---- staged kernels: operands arrive in shared memory through TMA ------------------------------------
// The direct kernels above read four 30 x 20 fp32 slices per warp as 80-byte row segments and wait for them
// (ncu: 40 % of the stall samples are long-scoreboard, 18 % of the instructions are address arithmetic).  Here
// a CTA owns one (batch item, group of HG heads) at a time: ONE elected thread asks the TMA unit for the
// [S rows] x [HG * DH columns] boxes of Q, K, V (and dO) -- whole 400-byte row segments, full DRAM bursts --
// of the item AFTER the next while the HG warps (one head each) compute the current one from shared memory;
// two stages, one mbarrier each.  The geometry (which rows form a sequence) lives in the tensor map, so the
// same kernel serves the title encoder (sequence = the tokens of a title) and any other (seq_stride,
// batch_stride) layout with S <= 32.
//   box = [S][pitch] fp32 with pitch = HG * DH rounded up so that pitch % 32 == 8: the 64-bit fragment loads
//   of a half-warp (rows g = 0..3, columns 2 tg) then hit 32 distinct banks.  The extra columns are the next
//   head group's (or zero fill past the tensor edge) and are never used.
struct SmemSrc {
  static constexpr bool kReload = true;
  const float* tile[4];  // shared-memory boxes of this stage: Q, K, V, dO
  int pitch, col0, S;    // col0 = (head within the group) * DH
  template <int DH>
  __device__ __forceinline__ void load_t(Blk& m, int which, float mul, int g, int tg) const {
    const float* base = tile[which] + col0;
#pragma unroll
    for (int rb = 0; rb < 4; ++rb) {
      const int r = 8 * rb + g;
      const bool rok = r < S;
      const float* rp = base + (rok ? r : 0) * pitch;
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        if (8 * cb >= DH) {
          m.h[rb][cb] = 0u; m.l[rb][cb] = 0u;
          continue;
        }
        const int c = 8 * cb + 2 * tg;
        float2 v = make_float2(0.f, 0.f);
        if (rok && c < DH) v = *reinterpret_cast<const float2*>(rp + c);
        split_pack2(v.x * mul, v.y * mul, m.h[rb][cb], m.l[rb][cb]);
      }
    }
  }
};
template <int DH>
struct SmemSrcT : SmemSrc {
  __device__ __forceinline__ void load(Blk& m, int which, float mul, int g, int tg) const {
    this->template load_t<DH>(m, which, mul, g, tg);
  }
  __device__ __forceinline__ float lse2(const float* lse, long long grow, int heads, int h, int /*r*/) const {
    return __ldg(lse + grow * heads + h) * NRL_LOG2E;
  }
};

constexpr int ATTN_TMA_MAX_HG = 5;
__host__ __device__ constexpr int attn_tma_pitch(int cols) {  // smallest p >= cols with p % 32 == 8
  return cols + ((8 - cols % 32) + 32) % 32;
}
__host__ __device__ constexpr int attn_tma_tile_bytes(int S, int pitch) {  // TMA destinations are 128-byte aligned
  return (S * pitch * 4 + 127) / 128 * 128;
}

// NT = number of operand tensors per item (3 forward: Q K V; 4 backward: + dO).
// Items: b-major, head group fastest; item -> (b = item / groups, grp = item % groups).
template <int NT>
struct AttnStage {
  uint32_t smem_base, full_bar;  // shared-space addresses
  int tile_bytes;
  __device__ __forceinline__ uint32_t tile(int stage, int t) const {
    return smem_base + (uint32_t)((stage * NT + t) * tile_bytes);
  }
  // issued by ONE thread
  __device__ __forceinline__ void issue(int stage, const CUtensorMap* tmQKV, const CUtensorMap* tmDO, long long item,
                                        int groups, int hg, int DH, int E) const {
    const int b = (int)(item / groups), grp = (int)(item % groups);
    const uint32_t bar = full_bar + 8u * (uint32_t)stage;
    mbar_expect_tx(bar, (uint32_t)(NT * tile_bytes_payload));
#pragma unroll
    for (int t = 0; t < 3; ++t) tma_load_3d(tile(stage, t), tmQKV, bar, t * E + grp * hg * DH, 0, b);
    if (NT == 4) tma_load_3d(tile(stage, 3), tmDO, bar, grp * hg * DH, 0, b);
  }
  int tile_bytes_payload;  // S * pitch * 4 (what one box transfers)
};

template <int DH>
__global__ void __launch_bounds__(32 * ATTN_TMA_MAX_HG, 3)
attn_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmQKV, int E, int heads, int hg, int S, long long seq_stride,
                    int NB, long long batch_stride, float scale, __nv_bfloat16* __restrict__ o_hi,
                    __nv_bfloat16* __restrict__ o_lo, int ep, float* __restrict__ lse) {
  extern __shared__ uint8_t attn_smem_raw[];
  const uint32_t base = (smem_u32(attn_smem_raw) + 127u) & ~127u;
  const int pitch = attn_tma_pitch(hg * DH);
  AttnStage<3> st;
  st.tile_bytes = attn_tma_tile_bytes(S, pitch);
  st.tile_bytes_payload = S * pitch * 4;
  st.full_bar = base;
  st.smem_base = base + 128u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int groups = (heads + hg - 1) / hg;
  const long long items = (long long)NB * groups;
  if (threadIdx.x == 0) {
    mbar_init(st.full_bar, 1);
    mbar_init(st.full_bar + 8u, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQKV);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if ((long long)blockIdx.x < items) st.issue(0, &tmQKV, nullptr, blockIdx.x, groups, hg, DH, E);
    if ((long long)blockIdx.x + gridDim.x < items) st.issue(1, &tmQKV, nullptr, (long long)blockIdx.x + gridDim.x, groups, hg, DH, E);
  }
  const uint8_t* gen = attn_smem_raw + (st.smem_base - smem_u32(attn_smem_raw));
  int i = 0;
  for (long long item = blockIdx.x; item < items; item += gridDim.x, ++i) {
    const int stage = i & 1;
    mbar_wait(st.full_bar + 8u * (uint32_t)stage, (uint32_t)((i >> 1) & 1));
    const int b = (int)(item / groups), grp = (int)(item % groups);
    const int h = grp * hg + warp;
    SmemSrcT<DH> src;
#pragma unroll
    for (int t = 0; t < 3; ++t) src.tile[t] = reinterpret_cast<const float*>(gen + (size_t)(stage * 3 + t) * st.tile_bytes);
    src.tile[3] = nullptr;
    src.pitch = pitch; src.col0 = warp * DH; src.S = S;
    const long long nxt = item + 2ll * gridDim.x;
    auto release = [&]() {
      __syncthreads();  // every warp has its fragments in registers: the stage may be overwritten
      if (threadIdx.x == 0 && nxt < items) st.issue(stage, &tmQKV, nullptr, nxt, groups, hg, DH, E);
    };
    if (warp < hg && h < heads) {
      attn_fwd_item<DH>(src, E, heads, S, seq_stride, batch_stride, scale, b, h, o_hi, o_lo, ep, lse, lane, release);
    } else {
      release();
    }
  }
}

template <int DH>
__global__ void __launch_bounds__(32 * ATTN_TMA_MAX_HG, 2)
attn_bwd_tma_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                    const float* __restrict__ lse, int E, int heads, int hg, int S, long long seq_stride, int NB,
                    long long batch_stride, float scale, __nv_bfloat16* __restrict__ g_hi,
                    __nv_bfloat16* __restrict__ g_lo, int p3) {
  extern __shared__ uint8_t attn_smem_raw[];
  const uint32_t base = (smem_u32(attn_smem_raw) + 127u) & ~127u;
  const int pitch = attn_tma_pitch(hg * DH);
  AttnStage<4> st;
  st.tile_bytes = attn_tma_tile_bytes(S, pitch);
  st.tile_bytes_payload = S * pitch * 4;
  st.full_bar = base;
  st.smem_base = base + 128u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int groups = (heads + hg - 1) / hg;
  const long long items = (long long)NB * groups;
  if (threadIdx.x == 0) {
    mbar_init(st.full_bar, 1);
    mbar_init(st.full_bar + 8u, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if ((long long)blockIdx.x < items) st.issue(0, &tmQKV, &tmDO, blockIdx.x, groups, hg, DH, E);
    if ((long long)blockIdx.x + gridDim.x < items) st.issue(1, &tmQKV, &tmDO, (long long)blockIdx.x + gridDim.x, groups, hg, DH, E);
  }
  const uint8_t* gen = attn_smem_raw + (st.smem_base - smem_u32(attn_smem_raw));
  int i = 0;
  for (long long item = blockIdx.x; item < items; item += gridDim.x, ++i) {
    const int stage = i & 1;
    mbar_wait(st.full_bar + 8u * (uint32_t)stage, (uint32_t)((i >> 1) & 1));
    const int b = (int)(item / groups), grp = (int)(item % groups);
    const int h = grp * hg + warp;
    SmemSrcT<DH> src;
#pragma unroll
    for (int t = 0; t < 4; ++t) src.tile[t] = reinterpret_cast<const float*>(gen + (size_t)(stage * 4 + t) * st.tile_bytes);
    src.pitch = pitch; src.col0 = warp * DH; src.S = S;
    const long long nxt = item + 2ll * gridDim.x;
    auto release = [&]() {
      __syncthreads();
      if (threadIdx.x == 0 && nxt < items) st.issue(stage, &tmQKV, &tmDO, nxt, groups, hg, DH, E);
    };
    if (warp < hg && h < heads) {
      attn_bwd_item<DH>(src, lse, E, heads, S, seq_stride, batch_stride, scale, b, h, g_hi, g_lo, p3, lane, release);
    } else {
      release();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// 32 < S <= 64 (the NRMS user encoder: the reference attends across the B = 64 impressions of the
// batch at every history position): the same register-fragment scheme on 32 x 32 blocks.
//   forward : one warp per (item, 32-query block); both key blocks' scores live in registers.
//   backward: four warps per item -- two own a query block (dQ, summed over the key blocks), two own
//             a key block (dK, dV, summed over the query blocks).  D_t = sum_d dO[t, d] O[t, d] comes
//             from the saved O planes, so every (query block, key block) pair is independent.
// ------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(128)
attn_fwd_mma64_kernel(const float* __restrict__ qkv, int E, int ldq, int heads, int S, long long seq_stride,
                      int NB, long long batch_stride, float scale, __nv_bfloat16* __restrict__ o_hi,
                      __nv_bfloat16* __restrict__ o_lo, int ep, float* __restrict__ lse) {
  pdl_launch_dependents();  // the next launch (a PDL-launched GEMM) may begin its prologue while this grid runs
  constexpr int KS = (DH + 15) / 16, ND = (DH + 7) / 8;
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long item = wid >> 1;
  const int qb = (int)(wid & 1);
  if (item >= (long long)NB * heads) return;
  const int b = (int)(item / heads), h = (int)(item % heads);
  const bool three = o_lo != nullptr;
  const float* base = qkv + (long long)b * batch_stride * ldq + h * DH;
  const long long rstride = seq_stride * ldq;
  const int Sq = min(32, S - 32 * qb);
  if (Sq <= 0) return;
  Blk q;
  load_blocks<DH>(q, base + 32ll * qb * rstride, rstride, Sq, scale * NRL_LOG2E, g, tg);
  float s[2][2][4][4];  // [key block][i][j][c]
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) s[kb][i][j][c] = 0.f;
    Blk k;
    load_blocks<DH>(k, base + E + 32ll * kb * rstride, rstride, S - 32 * kb, 1.f, g, tg);
    mma_nt<KS, 4>(s[kb], q, k, three);
  }
  float inv_l[4], row_lse[4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float m = -INFINITY;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const int col = 32 * kb + 8 * j + 2 * tg + c;
            if (col >= S) s[kb][i][j][2 * half + c] = -INFINITY;
            m = fmaxf(m, s[kb][i][j][2 * half + c]);
          }
      m = quad_max(m);
      float l = 0.f;
#pragma unroll
      for (int kb = 0; kb < 2; ++kb)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const float p = ex2_approx(s[kb][i][j][2 * half + c] - m);
            s[kb][i][j][2 * half + c] = p;
            l += p;
          }
      l = quad_sum(l);
      inv_l[2 * i + half] = 1.f / l;
      row_lse[2 * i + half] = m * NRL_LN2 + logf(l);
    }
  float o[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) o[i][j][c] = 0.f;
#pragma unroll
  for (int kb = 0; kb < 2; ++kb) {
    Blk p, v;
    acc_to_blocks(p, s[kb]);
    load_blocks<DH>(v, base + 2 * E + 32ll * kb * rstride, rstride, S - 32 * kb, 1.f, g, tg);
    mma_nn<ND>(o, p, v, three);
  }
  long long grow[4];
#pragma unroll
  for (int rb = 0; rb < 4; ++rb)
    grow[rb] = (long long)(32 * qb + 8 * rb + g) * seq_stride + (long long)b * batch_stride;
  store_acc_split<DH, ND>(o, inv_l[0], inv_l[1], inv_l[2], inv_l[3], o_hi, o_lo, ep, grow, Sq, h * DH, g, tg);
  if (tg == 0) {
#pragma unroll
    for (int rb = 0; rb < 4; ++rb)
      if (8 * rb + g < Sq) lse[grow[rb] * heads + h] = row_lse[rb];
  }
  if (h == 0) {
    const int npad = ep - E;
    for (int i = lane; i < Sq * npad; i += 32) {
      const int srow = 32 * qb + i / npad, c = E + i % npad;
      const long long gr = (long long)srow * seq_stride + (long long)b * batch_stride;
      o_hi[gr * ep + c] = __float2bfloat16_rn(c == E ? 1.f : 0.f);
      if (o_lo) o_lo[gr * ep + c] = __float2bfloat16_rn(0.f);
    }
  }
}

// D_t = sum_d dO[t, d] * O[t, d] for the 32 query rows starting at row0 (rows 8 rb + g of the block),
// O = hi + lo of the saved planes.  Result replicated over the quad.
template <int DH>
__device__ __forceinline__ void attn_row_D(float (&D)[4], const float* __restrict__ d_o, long long ld_do,
                                           const __nv_bfloat16* __restrict__ o_hi,
                                           const __nv_bfloat16* __restrict__ o_lo, int ep, int h, int row0, int Sq,
                                           long long seq_stride, long long batch_stride, int b, int g, int tg) {
#pragma unroll
  for (int rb = 0; rb < 4; ++rb) {
    float acc = 0.f;
    if (8 * rb + g < Sq) {
      const long long grow = (long long)(row0 + 8 * rb + g) * seq_stride + (long long)b * batch_stride;
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        const int c = 8 * cb + 2 * tg;
        if (8 * cb < DH && c < DH) {
          const float2 dv = __ldg(reinterpret_cast<const float2*>(d_o + grow * ld_do + h * DH + c));
          const __nv_bfloat162 oh = *reinterpret_cast<const __nv_bfloat162*>(o_hi + grow * ep + h * DH + c);
          float2 ov = __bfloat1622float2(oh);
          if (o_lo) {
            const float2 ol = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(o_lo + grow * ep + h * DH + c));
            ov.x += ol.x; ov.y += ol.y;
          }
          acc += dv.x * ov.x + dv.y * ov.y;
        }
      }
    }
    D[rb] = quad_sum(acc);
  }
}

// P and dS of one (query block, key block) pair: s, dp in -> p (in s), ds (in dp).
__device__ __forceinline__ void attn_p_ds(float (&s)[2][4][4], float (&dp)[2][4][4], const float (&lse2)[4],
                                          const float (&D)[4], int Sq, int Sk, int g, int tg) {
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int rb = 2 * i + half;
      const bool rok = 8 * rb + g < Sq;
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int col = 8 * j + 2 * tg + c;
          const float p = (rok && col < Sk) ? ex2_approx(s[i][j][2 * half + c] - lse2[rb]) : 0.f;
          s[i][j][2 * half + c] = p;
          dp[i][j][2 * half + c] = p * (dp[i][j][2 * half + c] - D[rb]);
        }
    }
}

template <int DH>
__global__ void __launch_bounds__(128, 2)
attn_bwd_mma64_kernel(const float* __restrict__ qkv, const float* __restrict__ d_o, long long ld_do,
                      const __nv_bfloat16* __restrict__ o_hi, const __nv_bfloat16* __restrict__ o_lo, int ep,
                      const float* __restrict__ lse, int E, int ldq, int heads, int S, long long seq_stride,
                      int NB, long long batch_stride, float scale, __nv_bfloat16* __restrict__ g_hi,
                      __nv_bfloat16* __restrict__ g_lo, int p3) {
  constexpr int KS = (DH + 15) / 16, ND = (DH + 7) / 8;
  const int lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3, role = threadIdx.x >> 5;
  const long long item = blockIdx.x;  // one CTA (4 warps) per (batch item, head)
  if (item >= (long long)NB * heads) return;
  const int b = (int)(item / heads), h = (int)(item % heads);
  const bool three = g_lo != nullptr;
  const float* base = qkv + (long long)b * batch_stride * ldq + h * DH;
  const float* dbase = d_o + (long long)b * batch_stride * ld_do + h * DH;
  const long long rstride = seq_stride * ldq, dstride = seq_stride * ld_do;
  const int mine = role & 1;          // the block this warp owns (query block for roles 0/1, key block for 2/3)
  const int Sm = min(32, S - 32 * mine);
  if (Sm <= 0) return;
  long long grow[4];
#pragma unroll
  for (int rb = 0; rb < 4; ++rb)
    grow[rb] = (long long)(32 * mine + 8 * rb + g) * seq_stride + (long long)b * batch_stride;
  float acc[2][4][4], acc2[2][4][4];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 4; ++c) { acc[i][j][c] = 0.f; acc2[i][j][c] = 0.f; }

  if (role < 2) {
    // ---- query block `mine`: dQ = scale * sum_kb dS[mine][kb] K[kb] ----
    Blk q, go;
    load_blocks<DH>(q, base + 32ll * mine * rstride, rstride, Sm, scale * NRL_LOG2E, g, tg);
    load_blocks<DH>(go, dbase + 32ll * mine * dstride, dstride, Sm, 1.f, g, tg);
    float lse2[4], D[4];
#pragma unroll
    for (int rb = 0; rb < 4; ++rb) lse2[rb] = (8 * rb + g < Sm) ? __ldg(lse + grow[rb] * heads + h) * NRL_LOG2E : 0.f;
    attn_row_D<DH>(D, d_o, ld_do, o_hi, o_lo, ep, h, 32 * mine, Sm, seq_stride, batch_stride, b, g, tg);
#pragma unroll
    for (int kb = 0; kb < 2; ++kb) {
      const int Sk = min(32, S - 32 * kb);
      if (Sk <= 0) continue;
      Blk k, v;
      load_blocks<DH>(k, base + E + 32ll * kb * rstride, rstride, Sk, 1.f, g, tg);
      load_blocks<DH>(v, base + 2 * E + 32ll * kb * rstride, rstride, Sk, 1.f, g, tg);
      float s[2][4][4], dp[2][4][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c) { s[i][j][c] = 0.f; dp[i][j][c] = 0.f; }
      mma_nt<KS, 4>(s, q, k, three);
      mma_nt<KS, 4>(dp, go, v, three);
      attn_p_ds(s, dp, lse2, D, Sm, Sk, g, tg);
      Blk ds;
      acc_to_blocks(ds, dp);
      mma_nn<ND>(acc, ds, k, three);
    }
    store_acc_split<DH, ND>(acc, scale, scale, scale, scale, g_hi, g_lo, p3, grow, Sm, h * DH, g, tg);
    if (h == 0 && p3 > 3 * E) {
      const int npad = p3 - 3 * E;
      for (int i = lane; i < Sm * npad; i += 32) {
        const int srow = 32 * mine + i / npad, c = 3 * E + i % npad;
        const long long gr = (long long)srow * seq_stride + (long long)b * batch_stride;
        g_hi[gr * p3 + c] = __float2bfloat16_rn(0.f);
        if (g_lo) g_lo[gr * p3 + c] = __float2bfloat16_rn(0.f);
      }
    }
  } else {
    // ---- key block `mine`: dK = ln2 * sum_qb dS[qb][mine]^T Qs[qb],  dV = sum_qb P[qb][mine]^T dO[qb] ----
    Blk k, v;
    load_blocks<DH>(k, base + E + 32ll * mine * rstride, rstride, Sm, 1.f, g, tg);
    load_blocks<DH>(v, base + 2 * E + 32ll * mine * rstride, rstride, Sm, 1.f, g, tg);
#pragma unroll
    for (int qb = 0; qb < 2; ++qb) {
      const int Sq = min(32, S - 32 * qb);
      if (Sq <= 0) continue;
      Blk q, go;
      load_blocks<DH>(q, base + 32ll * qb * rstride, rstride, Sq, scale * NRL_LOG2E, g, tg);
      load_blocks<DH>(go, dbase + 32ll * qb * dstride, dstride, Sq, 1.f, g, tg);
      float lse2[4], D[4];
#pragma unroll
      for (int rb = 0; rb < 4; ++rb) {
        const long long gr = (long long)(32 * qb + 8 * rb + g) * seq_stride + (long long)b * batch_stride;
        lse2[rb] = (8 * rb + g < Sq) ? __ldg(lse + gr * heads + h) * NRL_LOG2E : 0.f;
      }
      attn_row_D<DH>(D, d_o, ld_do, o_hi, o_lo, ep, h, 32 * qb, Sq, seq_stride, batch_stride, b, g, tg);
      float s[2][4][4], dp[2][4][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < 4; ++c) { s[i][j][c] = 0.f; dp[i][j][c] = 0.f; }
      mma_nt<KS, 4>(s, q, k, three);
      mma_nt<KS, 4>(dp, go, v, three);
      attn_p_ds(s, dp, lse2, D, Sq, Sm, g, tg);
      Blk pb, ds;
      acc_to_blocks(pb, s);
      acc_to_blocks(ds, dp);
      mma_tn<ND>(acc, ds, q, three);    // dK
      mma_tn<ND>(acc2, pb, go, three);  // dV
    }
    store_acc_split<DH, ND>(acc, NRL_LN2, NRL_LN2, NRL_LN2, NRL_LN2, g_hi, g_lo, p3, grow, Sm, E + h * DH, g, tg);
    store_acc_split<DH, ND>(acc2, 1.f, 1.f, 1.f, 1.f, g_hi, g_lo, p3, grow, Sm, 2 * E + h * DH, g, tg);
  }
}

}  // namespace nrl
