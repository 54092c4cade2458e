// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[m, n] = sum over segments s, sum_k  A_{pa[s]}[m, k] * B_{pb[s]}[n, k]          (fp32 in TMEM)
//
// A and B are "split-plane" bf16 matrices (plane 0 = hi, plane 1 = lo, see nrl_ptx.cuh).  With
// one plane this is a plain bf16 GEMM; with two planes every k-step issues three MMAs
// (lo*hi, hi*lo, hi*hi) which reproduce an fp32 GEMM to ~2e-5 relative, which is what the 1e-4
// logit-parity bar of the reference needs.  A pipeline stage holds ALL planes of one k-block
// (A_hi, A_lo, B_hi, B_lo), each loaded once: on B200 this kernel is bound by the L2 -> SM fill
// bandwidth (~6.3 KB/clk chip-wide), so three MMAs per four tile loads is 1.5x the intensity of
// re-streaming the operands once per pass.
//
// Two operand layouts:
//   mn_major = 0  ("NT"):  A[M, K] and B[N, K] row-major, K contiguous (forward / dgrad GEMMs).
//   mn_major = 1  ("TN"):  A stored [K, M], B stored [K, N] row-major, i.e. the reduction runs
//                          over ROWS of both (weight-gradient GEMMs dW = dOut^T * In), with
//                          split-K across CTAs and a TMA reduce-add (fp32) epilogue.
//
// Roles per CTA (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..9 = epilogue (two per TMEM lane quarter, alternate 32-column chunks).  N-stage smem ring (TMA <-> MMA), 2-stage TMEM accumulator ring
// (MMA <-> epilogue) so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Epilogue: each warp owns 32 accumulator rows (its TMEM lane quarter) and walks the tile in
// 32-column chunks: tcgen05.ld -> fused element-wise work with thread = row (bias rides on a
// ones column of A; dropout keep-bits from the counter RNG; tanh + query dot; rank-1 addend)
// -> swizzled smem staging box -> ONE TMA store per sink (fp32 box, or a [2 planes] bf16 hi/lo
// box, or a TMA reduce-add for split-K weight gradients).  TMA clips the box against the tensor
// extents, so ragged edges need no per-element bounds checks and every global write is a full
// bulk transaction.
//
// Tiles: BM = 128 rows; the N extent is cut into n-tiles of BN (<= 256) columns and the MMA of
// the last n-tile only issues the N it needs (any multiple of 16).  Work units are handed out
// round-robin: one unit = all n-tiles of one m-block (the A block stays hot in L2 and every
// unit costs the same), or single tiles for small / split-K problems.
#pragma once
#include "nrl_ptx.cuh"

namespace nrl {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // bf16 elements per k-block = one 128-byte swizzle span
constexpr int GEMM_MAX_STAGES = 6;
constexpr int GEMM_A_BYTES = GEMM_BM * 128;  // 16 KB
constexpr int GEMM_EPI_WARPS = 8;  // two per TMEM lane quarter: they take alternate 32-column chunks
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_BAR_BYTES = 256 + 1024;  // mbarriers + TMEM pointer, then the additive query vector (<= 256 floats)
constexpr int GEMM_SMEM_LIMIT = 227 * 1024;
constexpr int GEMM_TMEM_COLS = 512;

struct GemmEpi {
  // x = acc ; x += add_w[row] * add_vec[row / add_L, col] ; relu ; dropout ; positive-mask ;
  // tanh + query dot ; sinks
  const float* add_w;   const float* add_vec; long long ld_addvec; int add_L;
  int relu;             // x = max(x, 0)  (CNN / category-encoder forward)
  int gelu;             // the PLANE sink receives gelu(x) = 0.5 x (1 + erf(x / sqrt 2)) (HF "gelu"), the fp32 sink keeps
                        // the pre-activation x (what the backward pass needs): transformer feed-forward
  const float* add_mat; long long ld_addmat;  // x += add_mat[row, col] after dropout (residual connections)
  const float* gelu_pre; long long ld_gelu;   // x *= gelu'(gelu_pre[row, col])  (feed-forward backward)
  const float* pos_mask; long long ld_pos;  // x = pos_mask[row, col] > 0 ? x : 0  (ReLU backward)
  const uint32_t* drop_words; int drop_mw; float drop_scale;  // keep-bit words [row][drop_mw], or null
  uint32_t* pos_words; int pos_mw;  // out: bit i of word [row][col / 32] = (x > 0) after relu + dropout.  May alias
                                    // drop_words (each word is read and then written by the one thread that owns the
                                    // chunk): the backward pass then applies ReLU' and dropout' from ONE word per chunk
  const float* qvec;    float* score;         // x = tanh(x); score[row] = sum_col x * qvec[col]
  int f32_sink;         // 0 none, 1 TMA store of x to tmOut, 2 TMA reduce-add of x into tmOut
  int f32_cols;         // column extent of tmOut
  int sp_sink;          // 1: bf16 hi(/lo) planes of x to tmSp
  int sp_cols;          // column extent of tmSp (>= N: pad columns are written too)
  int sp_two;           // the lo plane exists
  int ones_col;         // plane column that is set to 1.0 (bias column of the next GEMM), or -1
  float* gb;  int gb_col;  // accumulator column gb_col is atomically added to gb[row] (bias grads)
};

struct GemmParams {
  int M, N, K;  // K (reduction extent) is a multiple of 16
  int BN;       // n-tile width (multiple of 32: epilogue boxes never straddle n-tiles)
  int n_extent; // columns the epilogue must produce (max over sinks, >= N)
  int mn_major;
  int planes;   // 1 = bf16 (hi only), 2 = hi + lo (three MMAs per k-step)
  int k_splits;
  int stages;
  int epi_buf_bytes;   // staging bytes per epilogue warp per buffer (4096 or 8192)
  int epi_bufs;        // staging buffers per epilogue warp (1 or 2): TMA stores in flight per warp
  int tiles_per_unit;  // n_tiles (unit = m-block) or 1
  int pair;            // 1: nrl_gemm_tc2_kernel (CTA pairs, cta_group::2)
  int fuse_n;          // pair + MN-major: both n-tiles accumulate in ONE k-loop (A is streamed once)
  int astat;           // nrl_gemm_tc2a_kernel: A panel of a unit resident in shared memory
  int direct;          // fp32 sink written straight from registers (16x256b TMEM loads, 8-byte global stores): plain or
                       // dropout epilogues with an fp32 sink only
  long long ld_out;    // row pitch (floats) of the fp32 sink for the direct path
  float* out;          // its base
  int dbg;             // timing experiments only (NRL_EPI_DEBUG): low bits 1 no TMA stores, 2 no staging either, 3 no TMEM loads; bit 8: release (not relaxed) tmem-empty arrive
  GemmEpi epi;
};

// Exact (erf) GELU of torch / HF "gelu" and its derivative.  Phi(x) = 0.5 (1 + erf(x / sqrt 2)) from the rational
// approximation erf(z) = 1 - (a1 t + ... + a5 t^5) exp(-z^2), t = 1 / (1 + p z)  (Abramowitz & Stegun 7.1.26, absolute
// error <= 1.5e-7, i.e. fp32 rounding level on Phi) -- its exp(-z^2) = exp(-x^2 / 2) is also the Gaussian of the
// derivative, so gelu'(x) = Phi(x) + x phi(x) costs one exponential in all.  ~16 instructions against ~45 for
// erff + expf: these run once per output element in the feed-forward GEMM epilogues, which are as long as their main loops.
__device__ __forceinline__ void gelu_parts(float x, float& Phi, float& gauss) {
  const float z = fabsf(x) * 0.70710678118654752f;
  gauss = __expf(-z * z);
  const float t = __fdividef(1.f, fmaf(0.3275911f, z, 1.f));
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float erf_abs = fmaf(-poly * t, gauss, 1.f);
  Phi = fmaf(0.5f, copysignf(erf_abs, x), 0.5f);
}
__device__ __forceinline__ float gelu_fwd(float x) {
  float Phi, g;
  gelu_parts(x, Phi, g);
  return x * Phi;
}
__device__ __forceinline__ float gelu_grad(float x) {
  float Phi, g;
  gelu_parts(x, Phi, g);
  return fmaf(x * 0.3989422804014327f, g, Phi);
}

struct GemmTile {
  int m0, n0, n_cur, kb0, kb1;
};

__device__ __forceinline__ GemmTile gemm_tile(const GemmParams& p, int tile, int m_tiles,
                                              int n_tiles, int kb_total) {
  GemmTile t;
  int m_blk, n_blk, ks;
  if (p.tiles_per_unit > 1) {  // unit = m-block, n-tiles consecutive
    n_blk = tile % n_tiles;
    m_blk = tile / n_tiles;
    ks = 0;
  } else {  // k-split major, then m-block, n-block fastest: the CTAs of one wave share their A rows
            // (the n-tiles of one (m-block, k-range) run side by side) and their B rows (all m-blocks
            // of a k-range) while those are still in L2 -- each operand leaves DRAM once
    const int per_ks = m_tiles * n_tiles;
    ks = tile / per_ks;
    const int r = tile % per_ks;
    m_blk = r / n_tiles;
    n_blk = r % n_tiles;
  }
  const int kb_per = (kb_total + p.k_splits - 1) / p.k_splits;
  t.m0 = m_blk * GEMM_BM;
  t.n0 = n_blk * p.BN;
  t.n_cur = min(p.BN, (p.n_extent - t.n0 + 15) & ~15);
  t.kb0 = ks * kb_per;
  t.kb1 = min(kb_total, t.kb0 + kb_per);
  return t;
}

// keep-bits of 32 consecutive elements starting at flat index e0 (bit i = element e0 + i).
// Element e uses Philox group e >> 3, 16-bit slot e & 7 (nrl_ptx.cuh); 32 elements span four or
// five groups depending on the alignment of e0.
__device__ __forceinline__ uint32_t philox_keep8(const Philox4& r, uint32_t thr) {
  // bit s = (slot s of the group is kept), slots in element order: x.lo, x.hi, y.lo, y.hi, ...
  uint32_t b = 0;
  b |= ((r.x & 0xFFFFu) >= thr) ? 1u : 0u;   b |= ((r.x >> 16) >= thr) ? 2u : 0u;
  b |= ((r.y & 0xFFFFu) >= thr) ? 4u : 0u;   b |= ((r.y >> 16) >= thr) ? 8u : 0u;
  b |= ((r.z & 0xFFFFu) >= thr) ? 16u : 0u;  b |= ((r.z >> 16) >= thr) ? 32u : 0u;
  b |= ((r.w & 0xFFFFu) >= thr) ? 64u : 0u;  b |= ((r.w >> 16) >= thr) ? 128u : 0u;
  return b;
}
__device__ __forceinline__ uint32_t drop_keep_bits32(unsigned long long seed, uint32_t site,
                                                     unsigned long long e0, uint32_t thr) {
  const unsigned long long g0 = e0 >> 3;
  const int sh = (int)(e0 & 7ull);
  unsigned long long bits = 0;  // keep bits of groups g0 .. g0 + 4 (40 slots), slot 0 of g0 at bit 0
#pragma unroll
  for (int k = 0; k < 4; ++k)
    bits |= (unsigned long long)philox_keep8(philox4x32(seed, g0 + k, site), thr) << (8 * k);
  if (sh) bits |= (unsigned long long)philox_keep8(philox4x32(seed, g0 + 4, site), thr) << 32;
  return (uint32_t)(bits >> sh);
}

// Epilogue of one accumulator tile for one epilogue warp (its 32-row TMEM lane quarter): walks
// the tile in 32-column chunks, applies the fused element-wise work and issues the TMA stores.
__device__ __forceinline__ void gemm_epilogue_tile(const GemmParams& p, const CUtensorMap* tmOutP,
                                                   const CUtensorMap* tmSpP, const GemmTile& t,
                                                   uint32_t t_row, int quarter, int lane,
                                                   uint32_t my_stage, uint32_t sp_off,
                                                   uint32_t& chunk_ctr, bool one, int half, uint32_t qv_smem) {
    const GemmEpi& e = p.epi;
    const CUtensorMap& tmOut = *tmOutP;
    const CUtensorMap& tmSp = *tmSpP;
    if (p.direct) {
      // ---- register-direct fp32 sink: no shared-memory staging, no TMA store (the staging buffer's turn-around --
      // waiting for the TMA unit to have read it, behind the producer's loads -- was 16 % of the epilogue warps' time,
      // and its reads and writes share the shared-memory port with the MMA operands)
      const int g = lane >> 2, tq = lane & 3;
      const int row0 = t.m0 + quarter * 32;
      if (row0 >= p.M) return;
      for (int c = 32 * half; c < t.n_cur; c += 32 * (GEMM_EPI_WARPS / 4)) {
        const int col_base = t.n0 + c;
        uint32_t r[2][16];
        tmem_ld_16x256b_x4(t_row + (uint32_t)c, r[0]);
        tmem_ld_16x256b_x4(t_row + (16u << 16) + (uint32_t)c, r[1]);
        tmem_wait_ld();
#pragma unroll
        for (int hb = 0; hb < 2; ++hb)
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            const int row = row0 + 16 * hb + 8 * h8 + g;
            if (row >= p.M) continue;
            uint32_t bits = 0xffffffffu;
            if (e.drop_words) bits = __ldg(e.drop_words + (long long)row * e.drop_mw + (col_base >> 5));
            float* orow = p.out + (long long)row * p.ld_out + col_base + 2 * tq;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (col_base + 8 * j + 2 * tq >= e.f32_cols) continue;
              float v0 = __uint_as_float(r[hb][4 * j + 2 * h8]), v1 = __uint_as_float(r[hb][4 * j + 2 * h8 + 1]);
              if (e.drop_words) {
                const uint32_t b2 = bits >> (8 * j + 2 * tq);
                v0 = (b2 & 1u) ? v0 * e.drop_scale : 0.f;
                v1 = (b2 & 2u) ? v1 * e.drop_scale : 0.f;
              }
              *reinterpret_cast<float2*>(orow + 8 * j) = make_float2(v0, v1);
            }
          }
      }
      return;
    }
    const int row_base = t.m0 + quarter * 32;
    const int row = row_base + lane;
    const bool row_ok = row < p.M;
    float score_acc = 0.f;
    float addw = 0.f;
    const float* addv = nullptr;
    if (e.add_w && row_ok) {
      addw = __ldg(e.add_w + row);
      addv = e.add_vec + (long long)(row / e.add_L) * e.ld_addvec;
    }
    for (int c = 32 * half; c < t.n_cur; c += 32 * (GEMM_EPI_WARPS / 4)) {
      const int col_base = t.n0 + c;
      float v[32];
      if ((p.dbg & 7) >= 3) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
      } else if (c + 16 < t.n_cur) {
        tmem_ld32(t_row + (uint32_t)c, v);
      } else {
        tmem_ld16(t_row + (uint32_t)c, v);
#pragma unroll
        for (int i = 16; i < 32; ++i) v[i] = 0.f;
      }
      if (row_base < p.M) {  // warp-uniform: rows past M produce nothing
        if (addv) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (col_base + 4 * q < p.N) {
              const float4 a4 = __ldg(reinterpret_cast<const float4*>(addv + col_base) + q);
              v[4 * q] += addw * a4.x; v[4 * q + 1] += addw * a4.y;
              v[4 * q + 2] += addw * a4.z; v[4 * q + 3] += addw * a4.w;
            }
          }
        }
        if (e.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (e.drop_words) {  // col_base % 32 == 0: one word holds this chunk's keep-bits
          const uint32_t* wp = e.drop_words + (long long)row * e.drop_mw + (col_base >> 5);
          // a word this launch also writes (pos_words aliasing drop_words) must not go through the read-only path
          const uint32_t bits = row_ok ? (e.pos_words ? *wp : __ldg(wp)) : 0u;
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = ((bits >> i) & 1u) ? v[i] * e.drop_scale : 0.f;
        }
        if (e.pos_words && row_ok && col_base < p.N) {  // chunks of pad columns own no word
          uint32_t pos = 0u;
#pragma unroll
          for (int i = 0; i < 32; ++i) pos |= (col_base + i < p.N && v[i] > 0.f) ? (1u << i) : 0u;
          e.pos_words[(long long)row * e.pos_mw + (col_base >> 5)] = pos;
        }
        if (e.add_mat && row_ok) {
          const float* am = e.add_mat + (long long)row * e.ld_addmat + col_base;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (col_base + 4 * q < p.N) {
              const float4 a4 = __ldg(reinterpret_cast<const float4*>(am) + q);
              v[4 * q] += a4.x; v[4 * q + 1] += a4.y; v[4 * q + 2] += a4.z; v[4 * q + 3] += a4.w;
            }
          }
        }
        if (e.gelu_pre && row_ok) {
          const float* gp = e.gelu_pre + (long long)row * e.ld_gelu + col_base;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (col_base + 4 * q < p.N) {
              const float4 u4 = __ldg(reinterpret_cast<const float4*>(gp) + q);
              v[4 * q] *= gelu_grad(u4.x); v[4 * q + 1] *= gelu_grad(u4.y);
              v[4 * q + 2] *= gelu_grad(u4.z); v[4 * q + 3] *= gelu_grad(u4.w);
            }
          }
        }
        if (e.pos_mask && row_ok) {
          const float* pm = e.pos_mask + (long long)row * e.ld_pos + col_base;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            if (col_base + 4 * q < p.N) {
              const float4 m4 = __ldg(reinterpret_cast<const float4*>(pm) + q);
              if (!(m4.x > 0.f)) v[4 * q] = 0.f;
              if (!(m4.y > 0.f)) v[4 * q + 1] = 0.f;
              if (!(m4.z > 0.f)) v[4 * q + 2] = 0.f;
              if (!(m4.w > 0.f)) v[4 * q + 3] = 0.f;
            }
          }
        }
        if (e.qvec) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float a = 0.f;
            if (col_base + i < p.N) {
              a = tanh_fast(v[i]);
              score_acc += a * __ldg(e.qvec + col_base + i);
            }
            v[i] = a;
          }
        }
        if (e.gb && row_ok && e.gb_col >= col_base && e.gb_col < col_base + 32) {
          float g = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col_base + i == e.gb_col) g = v[i];
          atomicAdd(e.gb + row, g);
        }
        if ((p.dbg & 7) >= 2) {  // keep the values alive without staging them
          float acc = 0.f;
#pragma unroll
          for (int i = 0; i < 32; ++i) acc += v[i];
          if (acc == 1.2345e30f) score_acc += acc;
          continue;
        }
        // ---- staging + TMA store ----
        const uint32_t buf = my_stage + (chunk_ctr & (uint32_t)(p.epi_bufs - 1)) * (uint32_t)p.epi_buf_bytes;
        ++chunk_ctr;
        if (one && (p.dbg & 7) == 0) {  // the store that last read this buffer is done
          if (p.epi_bufs == 1) bulk_wait_read<0>(); else bulk_wait_read<1>();
        }
        __syncwarp();
        if (e.f32_sink) {  // [32 rows][32 fp32], 128-byte rows, SWIZZLE_128B
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t dst = buf + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(v[4 * q]),
                         "f"(v[4 * q + 1]), "f"(v[4 * q + 2]), "f"(v[4 * q + 3])
                         : "memory");
          }
        }
        if (e.sp_sink) {  // [planes][32 rows][32 bf16], 64-byte rows, SWIZZLE_64B
          if (e.gelu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = gelu_fwd(v[i]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int i0 = 8 * q + 2 * u;
              const int c0 = col_base + i0;
              const float x0 = c0 < p.N ? v[i0] : (c0 == e.ones_col ? 1.f : 0.f);
              const float x1 = c0 + 1 < p.N ? v[i0 + 1] : (c0 + 1 == e.ones_col ? 1.f : 0.f);
              __nv_bfloat16 h0, l0, h1, l1;
              split_bf16(x0, h0, l0);
              split_bf16(x1, h1, l1);
              hw[u] = pack_bf16x2(h0, h1);
              lw[u] = pack_bf16x2(l0, l1);
            }
            const uint32_t dst =
                buf + sp_off + (uint32_t)lane * 64u + (uint32_t)((q ^ ((lane >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(hw[0]),
                         "r"(hw[1]), "r"(hw[2]), "r"(hw[3])
                         : "memory");
            if (e.sp_two)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 2048u),
                           "r"(lw[0]), "r"(lw[1]), "r"(lw[2]), "r"(lw[3])
                           : "memory");
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (one && (p.dbg & 7) == 0) {
          if (e.f32_sink == 1 && col_base < e.f32_cols) tma_store_2d(&tmOut, buf, col_base, row_base);
          else if (e.f32_sink == 2 && col_base < e.f32_cols) tma_reduce_add_2d(&tmOut, buf, col_base, row_base);
          if (e.sp_sink && col_base < e.sp_cols) tma_store_3d(&tmSp, buf + sp_off, col_base, row_base, 0);
          bulk_commit();
        }
      }
    }
    // two warps own a row (alternate chunks): each adds its partial into the zeroed score buffer
    // (exactly two addends onto 0: the result does not depend on the order)
    if (e.score && row_ok) atomicAdd(e.score + row, score_acc);
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
nrl_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmSp,
                   const GemmParams p) {
  pdl_launch_dependents();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_boxes = (uint32_t)(p.BN + 63) / 64u;  // mn_major: 64-column boxes of 8 KB
  const uint32_t b_bytes = p.mn_major ? b_boxes * 8192u : (uint32_t)p.BN * 128u;
  const uint32_t stage_bytes = (uint32_t)p.planes * (GEMM_A_BYTES + b_bytes);
  const uint32_t epi_base = smem_base + (uint32_t)p.stages * stage_bytes;
  const uint32_t bar_base = epi_base + (uint32_t)GEMM_EPI_WARPS * (uint32_t)p.epi_bufs * (uint32_t)p.epi_buf_bytes;
  // barrier layout (8 B each): full[S], empty[S], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (GEMM_MAX_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_MAX_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_MAX_STAGES + 2 + s); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * GEMM_MAX_STAGES + 4);

  // warp-uniform warp index and ONE elected lane per warp: inside `if (one)` the compiler knows a
  // single thread is active, so TMA / tcgen05 operands go straight to uniform registers (with
  // `lane == 0` it emits a vote loop around every UTMALDG / UTCHMMA / UTMASTG)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const bool one = elect_one();

  const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int n_tiles = (p.n_extent + p.BN - 1) / p.BN;
  const int kb_total = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int num_tiles = m_tiles * n_tiles * p.k_splits;
  const int tpu = p.tiles_per_unit;
  const int num_units = num_tiles / tpu;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), GEMM_EPI_WARPS);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.epi.f32_sink) tma_prefetch_desc(&tmOut);
    if (p.epi.sp_sink) tma_prefetch_desc(&tmSp);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, GEMM_TMEM_COLS);
    tmem_relinquish();
  }
  // everything above touches only this CTA's shared memory, TMEM and the kernel parameters: under a programmatic
  // dependent launch it overlaps the tail of the previous kernel; from here on global memory is read
  pdl_wait();
  if (p.epi.qvec)
    for (int i = threadIdx.x; i < p.N && i < 256; i += blockDim.x)
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(bar_base + 256u + 4u * (uint32_t)i), "f"(__ldg(p.epi.qvec + i)) : "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (one) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes = stage_bytes;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        for (int j = 0; j < tpu; ++j) {
          const GemmTile t = gemm_tile(p, unit * tpu + j, m_tiles, n_tiles, kb_total);
          if (t.kb0 >= t.kb1) continue;
          for (int kb = t.kb0; kb < t.kb1; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), tx_bytes);
            const uint32_t a_dst = smem_base + stage * stage_bytes;
            const uint32_t b_dst = a_dst + (uint32_t)p.planes * GEMM_A_BYTES;
            for (int pl = 0; pl < p.planes; ++pl) {
              if (!p.mn_major) {
                tma_load_3d(a_dst + pl * GEMM_A_BYTES, &tmA, full_bar(stage), kb * GEMM_BK, t.m0, pl);
                tma_load_3d(b_dst + pl * b_bytes, &tmB, full_bar(stage), kb * GEMM_BK, t.n0, pl);
              } else {
                for (int q = 0; q < 2; ++q)
                  tma_load_3d(a_dst + pl * GEMM_A_BYTES + q * 8192, &tmA, full_bar(stage), t.m0 + q * 64,
                              kb * GEMM_BK, pl);
                for (int q = 0; q < (int)b_boxes; ++q)
                  tma_load_3d(b_dst + pl * b_bytes + q * 8192, &tmB, full_bar(stage), t.n0 + q * 64,
                              kb * GEMM_BK, pl);
              }
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (one) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        for (int j = 0; j < tpu; ++j) {
          const GemmTile t = gemm_tile(p, unit * tpu + j, m_tiles, n_tiles, kb_total);
          if (t.kb0 >= t.kb1) continue;
          const uint32_t idesc = umma_idesc_bf16(GEMM_BM, t.n_cur, p.mn_major, p.mn_major);
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
          uint32_t accumulate = 0;
          for (int kb = t.kb0; kb < t.kb1; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_src = smem_base + stage * stage_bytes;
            const uint32_t b_src = a_src + (uint32_t)p.planes * GEMM_A_BYTES;
            const int nks = min(GEMM_BK / 16, (p.K - kb * GEMM_BK + 15) / 16);
            for (int k = 0; k < nks; ++k) {
              const uint32_t koff = p.mn_major ? k * 2048 : k * 32;
              const uint32_t lbo = p.mn_major ? 8192 : 16;
              const uint64_t a_hi = umma_desc_sw128(a_src + koff, lbo, 1024);
              const uint64_t b_hi = umma_desc_sw128(b_src + koff, lbo, 1024);
              if (p.planes == 2) {
                const uint64_t a_lo = umma_desc_sw128(a_src + GEMM_A_BYTES + koff, lbo, 1024);
                const uint64_t b_lo = umma_desc_sw128(b_src + b_bytes + koff, lbo, 1024);
                umma_bf16(d_tmem, a_lo, b_hi, idesc, accumulate);
                umma_bf16(d_tmem, a_hi, b_lo, idesc, 1);
                umma_bf16(d_tmem, a_hi, b_hi, idesc, 1);
              } else {
                umma_bf16(d_tmem, a_hi, b_hi, idesc, accumulate);
              }
              accumulate = 1;
            }
            umma_commit(empty_bar(stage));
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
          umma_commit(tfull_bar(acc));
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const GemmEpi& e = p.epi;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may touch
    const uint32_t my_stage = epi_base + (uint32_t)(warp - 2) * (uint32_t)p.epi_bufs * (uint32_t)p.epi_buf_bytes;
    const uint32_t sp_off = e.f32_sink ? 4096u : 0u;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t chunk_ctr = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      for (int j = 0; j < tpu; ++j) {
        const GemmTile t = gemm_tile(p, unit * tpu + j, m_tiles, n_tiles, kb_total);
        if (t.kb0 >= t.kb1) continue;
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(quarter * 32) << 16);
        gemm_epilogue_tile(p, &tmOut, &tmSp, t, t_row, quarter, lane, my_stage, sp_off, chunk_ctr, one, (warp - 2) >> 2, bar_base + 256u);
        tc_fence_before();
        __syncwarp();
        // relaxed arrive: the TMEM reads are complete (tcgen05.wait::ld) and fenced; a release arrive costs a MEMBAR that
        // also waits for this lane's TMA stores in flight (ncu: 19 % of the epilogue warps' time)
        if (one) { if (p.dbg & 8) mbar_arrive(tempty_bar(acc)); else mbar_arrive_relaxed(tempty_bar(acc)); }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
    if (one) bulk_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, GEMM_TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2) of the NT GEMM for the big row-streaming GEMMs.
// Two CTAs of a cluster (one TPC) own 256 consecutive rows: each loads ITS 128 rows of A and
// HALF of the B tile (BN / 2 weight rows) per k-block, the leader's MMA thread issues M = 256
// tcgen05.mma.cta_group::2 instructions that read both CTAs' shared memory and write each CTA's
// half of the accumulator into that CTA's TMEM.  Per SM this halves the B bytes pulled from L2 per
// flop -- the 1-CTA kernel is bound by the L2 -> SM fill rate, not by the tensor pipe -- and the
// smaller stage (A 32 KB + B/2) leaves room for a 3-deep ring.
//   barriers (same offsets in both CTAs): full[s]  -- used in the LEADER only: 1 arrival (leader
//   producer, expect_tx of both CTAs' bytes) + the TMA bytes of both CTAs; empty[s] / tmem_full[a]
//   -- one multicast tcgen05.commit arrival in each CTA; tmem_empty[a] -- leader only, 8 arrivals
//   (4 epilogue warps of each CTA, the peer's through mapa).
// Work unit = (k-split, 256-row block) with all its n-tiles.  Both operand layouts: K-major "NT" (row
// streaming GEMMs) and MN-major "TN" with split-K (weight gradients whose M fills the row pairs).
// ------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
nrl_gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmSp,
                    const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t half_n = (uint32_t)p.BN / 2u;  // B rows (NT) / columns (TN) held by each CTA
  // NT: half_n K-major rows of 128 B; TN: 64-column boxes of 8 KB ([64 k-rows][64 columns])
  // fused n-tiles (weight gradients with N = BN + a narrow second tile): boxes of tile 0 then tile 1
  const int n1_cur = p.fuse_n ? ((p.n_extent - p.BN + 15) & ~15) : 0;  // width of the second tile
  const uint32_t nb0 = (half_n + 63u) / 64u, nb1 = p.fuse_n ? ((uint32_t)n1_cur / 2u + 63u) / 64u : 0u;
  const uint32_t b_bytes = p.mn_major ? (nb0 + nb1) * 8192u : half_n * 128u;
  const uint32_t stage_bytes = (uint32_t)p.planes * (GEMM_A_BYTES + b_bytes);
  const uint32_t epi_base = smem_base + (uint32_t)p.stages * stage_bytes;
  const uint32_t bar_base = epi_base + (uint32_t)GEMM_EPI_WARPS * (uint32_t)p.epi_bufs * (uint32_t)p.epi_buf_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (GEMM_MAX_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_MAX_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_MAX_STAGES + 2 + s); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * GEMM_MAX_STAGES + 4);

  // warp-uniform warp index and ONE elected lane per warp (see nrl_gemm_tc_kernel)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const bool one = elect_one();
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  const int m_pairs = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const int n_tiles = p.fuse_n ? 1 : (p.n_extent + p.BN - 1) / p.BN;  // passes over K per unit
  const int kb_total = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int kb_per = (kb_total + p.k_splits - 1) / p.k_splits;
  // unit = (k-split, row pair) with all its n-tiles, or -- tiles_per_unit == 1, NT GEMMs only -- ONE (row pair, n-tile):
  // when the row pairs do not fill a whole number of waves (150 pairs on 74 clusters = 3 waves of 12-tile units, the
  // last with 2 clusters busy) single tiles hand the remainder out evenly (1800 tiles = 25 rounds instead of 36)
  const bool fine = p.tiles_per_unit == 1 && !p.fuse_n && p.k_splits == 1 && !p.mn_major;
  const int num_units = fine ? m_pairs * n_tiles : m_pairs * p.k_splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * GEMM_EPI_WARPS);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.epi.f32_sink) tma_prefetch_desc(&tmOut);
    if (p.epi.sp_sink) tma_prefetch_desc(&tmSp);
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr_addr, GEMM_TMEM_COLS);
    tmem_relinquish_pair();
  }
  if (p.epi.qvec)
    for (int i = threadIdx.x; i < p.N && i < 256; i += blockDim.x)
      asm volatile("st.shared.f32 [%0], %1;" ::"r"(bar_base + 256u + 4u * (uint32_t)i), "f"(__ldg(p.epi.qvec + i)) : "memory");
  tc_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before anybody signals them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  auto tile_of = [&](int unit, int j) {
    GemmTile t;
    const int ks = unit / m_pairs, mp = unit - ks * m_pairs;
    t.m0 = mp * 2 * GEMM_BM + (int)rank * GEMM_BM;  // this CTA's 128 rows
    t.n0 = j * p.BN;
    t.n_cur = min(p.BN, (p.n_extent - t.n0 + 15) & ~15);
    t.kb0 = ks * kb_per;
    t.kb1 = min(kb_total, t.kb0 + kb_per);
    return t;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (one) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
        const int ub = fine ? unit / n_tiles : unit, j0 = fine ? unit - ub * n_tiles : 0, j1 = fine ? j0 + 1 : n_tiles;
        for (int j = j0; j < j1; ++j) {
          const GemmTile t = tile_of(ub, j);
          const int b0 = t.n0 + (int)rank * (t.n_cur / 2);  // my half of the n_cur weight rows / columns
          const int nb = (t.n_cur / 2 + 63) / 64;           // TN: 64-column boxes of my half
          const uint32_t tx = p.mn_major ? (uint32_t)p.planes * (GEMM_A_BYTES + ((uint32_t)nb + nb1) * 8192u) : stage_bytes;
          for (int kb = t.kb0; kb < t.kb1; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
            if (leader) mbar_expect_tx(full_bar(stage), 2u * tx);
            const uint32_t a_dst = smem_base + stage * stage_bytes;
            const uint32_t b_dst = a_dst + (uint32_t)p.planes * GEMM_A_BYTES;
            for (int pl = 0; pl < p.planes; ++pl) {
              if (!p.mn_major) {
                tma_load_3d_pair(a_dst + pl * GEMM_A_BYTES, &tmA, lead_full, kb * GEMM_BK, t.m0, pl);
                tma_load_3d_pair(b_dst + pl * b_bytes, &tmB, lead_full, kb * GEMM_BK, b0, pl);
              } else {
                for (int q = 0; q < 2; ++q)
                  tma_load_3d_pair(a_dst + pl * GEMM_A_BYTES + q * 8192, &tmA, lead_full, t.m0 + q * 64,
                                   kb * GEMM_BK, pl);
                for (int q = 0; q < nb; ++q)
                  tma_load_3d_pair(b_dst + pl * b_bytes + q * 8192, &tmB, lead_full, b0 + q * 64, kb * GEMM_BK, pl);
                for (int q = 0; q < (int)nb1; ++q)  // my half of the fused second n-tile
                  tma_load_3d_pair(b_dst + pl * b_bytes + (nb + q) * 8192, &tmB, lead_full,
                                   p.BN + (int)rank * (n1_cur / 2) + q * 64, kb * GEMM_BK, pl);
              }
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (one && leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
        const int ub = fine ? unit / n_tiles : unit, j0 = fine ? unit - ub * n_tiles : 0, j1 = fine ? j0 + 1 : n_tiles;
        for (int j = j0; j < j1; ++j) {
          const GemmTile t = tile_of(ub, j);
          const uint32_t idesc = umma_idesc_bf16(2 * GEMM_BM, t.n_cur, p.mn_major, p.mn_major);
          const uint32_t idesc1 = umma_idesc_bf16(2 * GEMM_BM, p.fuse_n ? n1_cur : 16, p.mn_major, p.mn_major);
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
          uint32_t accumulate = 0;
          for (int kb = t.kb0; kb < t.kb1; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_src = smem_base + stage * stage_bytes;
            const uint32_t b_src = a_src + (uint32_t)p.planes * GEMM_A_BYTES;
            const int nks = min(GEMM_BK / 16, (p.K - kb * GEMM_BK + 15) / 16);
            for (int k = 0; k < nks; ++k) {
              const uint32_t koff = p.mn_major ? k * 2048 : k * 32;
              const uint32_t lbo = p.mn_major ? 8192 : 16;
              const uint64_t a_hi = umma_desc_sw128(a_src + koff, lbo, 1024);
              const uint64_t b_hi = umma_desc_sw128(b_src + koff, lbo, 1024);
              if (p.planes == 2) {
                const uint64_t a_lo = umma_desc_sw128(a_src + GEMM_A_BYTES + koff, lbo, 1024);
                const uint64_t b_lo = umma_desc_sw128(b_src + b_bytes + koff, lbo, 1024);
                umma_bf16_pair(d_tmem, a_lo, b_hi, idesc, accumulate);
                umma_bf16_pair(d_tmem, a_hi, b_lo, idesc, 1);
                umma_bf16_pair(d_tmem, a_hi, b_hi, idesc, 1);
                if (p.fuse_n) {  // second n-tile from the same A k-block, accumulator columns 256..
                  const uint64_t c_hi = umma_desc_sw128(b_src + nb0 * 8192u + koff, lbo, 1024);
                  const uint64_t c_lo = umma_desc_sw128(b_src + b_bytes + nb0 * 8192u + koff, lbo, 1024);
                  umma_bf16_pair(d_tmem + 256u, a_lo, c_hi, idesc1, accumulate);
                  umma_bf16_pair(d_tmem + 256u, a_hi, c_lo, idesc1, 1);
                  umma_bf16_pair(d_tmem + 256u, a_hi, c_hi, idesc1, 1);
                }
              } else {
                umma_bf16_pair(d_tmem, a_hi, b_hi, idesc, accumulate);
                if (p.fuse_n) {
                  const uint64_t c_hi = umma_desc_sw128(b_src + nb0 * 8192u + koff, lbo, 1024);
                  umma_bf16_pair(d_tmem + 256u, a_hi, c_hi, idesc1, accumulate);
                }
              }
              accumulate = 1;
            }
            umma_commit_pair(empty_bar(stage));  // frees the slot in both CTAs
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
          umma_commit_pair(tfull_bar(acc));  // accumulator halves ready in both CTAs
          if (p.fuse_n) {
            acc_phase ^= 1u;  // one accumulator set (buffer 0), a new phase per unit
          } else {
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1u;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps (2..9), both CTAs =====================
    const GemmEpi& e = p.epi;
    const int quarter = warp & 3;
    const uint32_t my_stage = epi_base + (uint32_t)(warp - 2) * (uint32_t)p.epi_bufs * (uint32_t)p.epi_buf_bytes;
    const uint32_t sp_off = e.f32_sink ? 4096u : 0u;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t chunk_ctr = 0;
    for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
      const int ub = fine ? unit / n_tiles : unit, j0 = fine ? unit - ub * n_tiles : 0, j1 = fine ? j0 + 1 : n_tiles;
      for (int j = j0; j < j1; ++j) {
        const GemmTile t = tile_of(ub, j);
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(quarter * 32) << 16);
        gemm_epilogue_tile(p, &tmOut, &tmSp, t, t_row, quarter, lane, my_stage, sp_off, chunk_ctr, one, (warp - 2) >> 2, bar_base + 256u);
        if (p.fuse_n) {  // the second n-tile's accumulator sits 256 columns further
          GemmTile t1 = t;
          t1.n0 = p.BN;
          t1.n_cur = n1_cur;
          gemm_epilogue_tile(p, &tmOut, &tmSp, t1, t_row + 256u, quarter, lane, my_stage, sp_off, chunk_ctr, one,
                             (warp - 2) >> 2, bar_base + 256u);
        }
        tc_fence_before();
        __syncwarp();
        if (one) {
          if (p.dbg & 8) {  // A/B: release arrive (MEMBAR in front)
            if (leader) mbar_arrive(tempty_bar(acc));
            else mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
          } else {
            if (leader) mbar_arrive_relaxed(tempty_bar(acc));
            else mbar_arrive_cluster_relaxed(mapa_shared(tempty_bar(acc), 0));
          }
        }
        if (p.fuse_n) {
          acc_phase ^= 1u;
        } else {
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      }
    }
    if (one) bulk_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, GEMM_TMEM_COLS);
  }
}


// ------------------------------------------------------------------------------------------
// A-stationary CTA-pair NT GEMM ("tc2a").  The pair kernel above re-streams the A panel of a 256-row unit once per
// n-tile (in-projection: 4 x) and its main loop is bound by the L2 -> SM fill rate: 62 KB per k-block and SM against
// 1440 clocks of MMA is 43 B / clk / SM = the ~6.3 KB / clk the chip's L2 delivers, so the tensor pipe idles 40 % of
// the time (ncu, profiles/r02_ncu_gemm.md).  Here the A panel of a unit -- all k-blocks, both planes, 160 KB at
// K = 304 -- stays in shared memory while the n-tiles of the unit stream only their B halves through a small ring:
// fill traffic of the in-projection drops from 2.3 MB to 1.4 MB per unit.  The room comes from the register-direct
// epilogue (no staging buffers).  A slot kb of the NEXT unit is reloaded as soon as the LAST n-tile of the current unit
// has consumed it (per-k-block full / empty barriers), so the pipeline never drains between units.
//   barriers: a_full[kb] / b_full[s] (leader only, expect_tx of both CTAs' bytes), a_empty[kb] / b_empty[s] / tmem_full[a]
//   (multicast tcgen05.commit, one arrival per CTA), tmem_empty[a] (leader only, 16 arrivals).
// Requirements (host checks): K-major operands, k_splits == 1, fp32 sink with a plain or dropout epilogue (direct path),
// ceil(K / 64) <= GEMM_A_SLOTS.
// ------------------------------------------------------------------------------------------
constexpr int GEMM_A_SLOTS = 8;
constexpr int GEMM_B_STAGES_MAX = 4;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
nrl_gemm_tc2a_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmSp,
                     const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int kb_total = (p.K + GEMM_BK - 1) / GEMM_BK;
  const uint32_t half_n = (uint32_t)p.BN / 2u;
  const uint32_t a_slot_bytes = (uint32_t)p.planes * GEMM_A_BYTES;
  const uint32_t b_bytes = half_n * 128u;                         // one plane of this CTA's half of a B k-block
  const uint32_t b_stage_bytes = (uint32_t)p.planes * b_bytes;
  const uint32_t b_base = smem_base + (uint32_t)kb_total * a_slot_bytes;
  const uint32_t bar_base = b_base + (uint32_t)p.stages * b_stage_bytes;
  auto a_full = [&](int kb) { return bar_base + 8u * kb; };
  auto a_empty = [&](int kb) { return bar_base + 8u * (GEMM_A_SLOTS + kb); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * GEMM_A_SLOTS + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * GEMM_A_SLOTS + GEMM_B_STAGES_MAX + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_A_SLOTS + 2 * GEMM_B_STAGES_MAX + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_A_SLOTS + 2 * GEMM_B_STAGES_MAX + 2 + s); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * GEMM_A_SLOTS + 2 * GEMM_B_STAGES_MAX + 4);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const bool one = elect_one();
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int m_pairs = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  const int n_tiles = (p.n_extent + p.BN - 1) / p.BN;
  const int num_units = m_pairs;

  if (threadIdx.x == 0) {
    for (int kb = 0; kb < kb_total; ++kb) {
      mbar_init(a_full(kb), 1);
      mbar_init(a_empty(kb), 1);
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 2 * GEMM_EPI_WARPS);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc_pair(tmem_ptr_addr, GEMM_TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  auto tile_of = [&](int unit, int j) {
    GemmTile t;
    t.m0 = unit * 2 * GEMM_BM + (int)rank * GEMM_BM;
    t.n0 = j * p.BN;
    t.n_cur = min(p.BN, (p.n_extent - t.n0 + 15) & ~15);
    t.kb0 = 0;
    t.kb1 = kb_total;
    return t;
  };

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (one) {
      int bs = 0;
      uint32_t bphase = 0, uiter = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters, ++uiter) {
        const uint32_t aph = uiter & 1u;
        for (int j = 0; j < n_tiles; ++j) {
          const GemmTile t = tile_of(unit, j);
          const int b0 = t.n0 + (int)rank * (t.n_cur / 2);
          for (int kb = 0; kb < kb_total; ++kb) {
            if (j == 0) {  // this unit's A k-block: its slot is free once the previous unit's last n-tile has used it
              mbar_wait(a_empty(kb), aph ^ 1u);
              const uint32_t lead = mapa_shared(a_full(kb), 0);
              if (leader) mbar_expect_tx(a_full(kb), 2u * a_slot_bytes);
              for (int pl = 0; pl < p.planes; ++pl)
                tma_load_3d_pair(smem_base + (uint32_t)kb * a_slot_bytes + (uint32_t)pl * GEMM_A_BYTES, &tmA, lead,
                                 kb * GEMM_BK, t.m0, pl);
            }
            mbar_wait(b_empty(bs), bphase ^ 1u);
            const uint32_t leadb = mapa_shared(b_full(bs), 0);
            if (leader) mbar_expect_tx(b_full(bs), 2u * b_stage_bytes);
            for (int pl = 0; pl < p.planes; ++pl)
              tma_load_3d_pair(b_base + (uint32_t)bs * b_stage_bytes + (uint32_t)pl * b_bytes, &tmB, leadb, kb * GEMM_BK, b0, pl);
            if (++bs == p.stages) { bs = 0; bphase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (one && leader) {
      int bs = 0, acc = 0;
      uint32_t bphase = 0, acc_phase = 0, uiter = 0;
      for (int unit = cluster_id; unit < num_units; unit += num_clusters, ++uiter) {
        const uint32_t aph = uiter & 1u;
        for (int j = 0; j < n_tiles; ++j) {
          const GemmTile t = tile_of(unit, j);
          const uint32_t idesc = umma_idesc_bf16(2 * GEMM_BM, t.n_cur, 0, 0);
          mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
          uint32_t accumulate = 0;
          for (int kb = 0; kb < kb_total; ++kb) {
            if (j == 0) mbar_wait(a_full(kb), aph);
            mbar_wait(b_full(bs), bphase);
            tc_fence_after();
            const uint32_t a_src = smem_base + (uint32_t)kb * a_slot_bytes;
            const uint32_t b_src = b_base + (uint32_t)bs * b_stage_bytes;
            const int nks = min(GEMM_BK / 16, (p.K - kb * GEMM_BK + 15) / 16);
            for (int k = 0; k < nks; ++k) {
              const uint32_t koff = k * 32;
              const uint64_t a_hi = umma_desc_sw128(a_src + koff, 16, 1024);
              const uint64_t b_hi = umma_desc_sw128(b_src + koff, 16, 1024);
              if (p.planes == 2) {
                const uint64_t a_lo = umma_desc_sw128(a_src + GEMM_A_BYTES + koff, 16, 1024);
                const uint64_t b_lo = umma_desc_sw128(b_src + b_bytes + koff, 16, 1024);
                umma_bf16_pair(d_tmem, a_lo, b_hi, idesc, accumulate);
                umma_bf16_pair(d_tmem, a_hi, b_lo, idesc, 1);
                umma_bf16_pair(d_tmem, a_hi, b_hi, idesc, 1);
              } else {
                umma_bf16_pair(d_tmem, a_hi, b_hi, idesc, accumulate);
              }
              accumulate = 1;
            }
            umma_commit_pair(b_empty(bs));
            if (j == n_tiles - 1) umma_commit_pair(a_empty(kb));  // the unit is done with this A k-block
            if (++bs == p.stages) { bs = 0; bphase ^= 1u; }
          }
          umma_commit_pair(tfull_bar(acc));
          acc ^= 1;
          if (acc == 0) acc_phase ^= 1u;
        }
      }
    }
  } else {
    // ===================== epilogue warps (2..9), both CTAs: register-direct fp32 sink =====================
    const int quarter = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0, chunk_ctr = 0;
    for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
      for (int j = 0; j < n_tiles; ++j) {
        const GemmTile t = tile_of(unit, j);
        mbar_wait(tfull_bar(acc), acc_phase);
        tc_fence_after();
        const uint32_t t_row = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(quarter * 32) << 16);
        gemm_epilogue_tile(p, &tmOut, &tmSp, t, t_row, quarter, lane, 0u, 0u, chunk_ctr, one, (warp - 2) >> 2, 0u);
        tc_fence_before();
        __syncwarp();
        if (one) {
          if (leader) mbar_arrive_relaxed(tempty_bar(acc));
          else mbar_arrive_cluster_relaxed(mapa_shared(tempty_bar(acc), 0));
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, GEMM_TMEM_COLS);
  }
}

}  // namespace nrl
