// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[m, n] = sum over segments s, sum_k  A_{pa[s]}[m, k] * B_{pb[s]}[n, k]          (fp32 in TMEM)
//
// A and B are "split-plane" bf16 matrices (plane 0 = hi, plane 1 = lo, see nrl_ptx.cuh).  One
// segment (hi*hi) is a plain bf16 GEMM; three segments (lo*hi, hi*lo, hi*hi) reproduce an fp32
// GEMM to ~2e-5 relative, which is what the 1e-4 logit-parity bar of the reference needs.
//
// Two operand layouts:
//   mn_major = 0  ("NT"):  A[M, K] and B[N, K] row-major, K contiguous (forward / dgrad GEMMs).
//   mn_major = 1  ("TN"):  A stored [K, M], B stored [K, N] row-major, i.e. the reduction runs
//                          over ROWS of both (weight-gradient GEMMs dW = dOut^T * In), with
//                          split-K across CTAs and an atomic fp32 epilogue.
//
// Roles per CTA (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> fused epilogue -> global).  4-stage smem ring
// (TMA <-> MMA), 2-stage TMEM accumulator ring (MMA <-> epilogue) so the epilogue of tile i
// overlaps the MMAs of tile i+1.  Grid = min(#tiles, #SMs), static round-robin tile schedule.
#pragma once
#include "nrl_ptx.cuh"

namespace nrl {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // bf16 elements per k-block = one 128-byte swizzle span
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_A_BYTES = GEMM_BM * 128;  // 16 KB
constexpr int GEMM_B_BYTES = 256 * 128;      // 32 KB (BN <= 256)
constexpr int GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_B_BYTES;
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int GEMM_TMEM_COLS = 512;

struct GemmEpi {
  // v = acc (+ addend) ; optional dropout ; then any of the sinks below.
  const float* addend;  long long ld_add;
  float* out;           long long ld_out;  int out_cols;   // fp32 row-major sink, cols < out_cols
  __nv_bfloat16* hi;    __nv_bfloat16* lo; long long ld_sp; int sp_cols; int ones_col;
  float drop_scale;     uint32_t drop_thr; uint32_t drop_site; unsigned long long seed; int drop_ld;
  int   use_dropout;
  // additive-attention score fusion (needs a single n-tile): a = tanh(v); score[m] = sum_n a*q[n]
  const float* qvec;    float* tanh_out;   long long ld_tanh; float* score;
  // atomic weight-gradient sink: col < gw_cols -> gw[m*ld_gw + col]; col == gw_cols -> gb[m]
  float* gw;            long long ld_gw;   int gw_cols;      float* gb;
};

struct GemmParams {
  int M, N, K;  // K (reduction extent) is a multiple of 16
  int BN;
  int mn_major;
  int num_segs;
  int seg_a[3];
  int seg_b[3];
  int k_splits;
  GemmEpi epi;
};

__device__ __forceinline__ void gemm_tile_coords(const GemmParams& p, int tile, int m_tiles,
                                                 int n_tiles, int kb_total, int& m0, int& n0,
                                                 int& kb0, int& kb1) {
  int n_blk = tile % n_tiles;
  int t = tile / n_tiles;
  int m_blk = t % m_tiles;
  int ks = t / m_tiles;
  int kb_per = (kb_total + p.k_splits - 1) / p.k_splits;
  m0 = m_blk * GEMM_BM;
  n0 = n_blk * p.BN;
  kb0 = ks * kb_per;
  kb1 = min(kb_total, kb0 + kb_per);
}

__device__ __forceinline__ void gemm_epilogue_chunk(const GemmEpi& e, int row, int col0, int N,
                                                    bool row_ok, float* v, float& score_acc) {
  // v[16] = accumulator columns col0 .. col0+15 of this thread's row
  if (!row_ok) return;
  if (e.addend) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (col0 + i < N) v[i] += __ldg(e.addend + (long long)row * e.ld_add + col0 + i);
  }
  if (e.use_dropout) {
    // col0 is a multiple of 16 and drop_ld a multiple of 8 -> two aligned groups of 8
    unsigned long long base = (unsigned long long)row * (unsigned)e.drop_ld + (unsigned)col0;
    if ((base & 7ull) == 0) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        Philox4 r = philox4x32_10(e.seed, (base >> 3) + g, e.drop_site);
#pragma unroll
        for (int i = 0; i < 8; ++i)
          v[g * 8 + i] = philox_u16(r, i) >= e.drop_thr ? v[g * 8 + i] * e.drop_scale : 0.f;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        v[i] = drop_keep(e.seed, e.drop_site, base + i, e.drop_thr) ? v[i] * e.drop_scale : 0.f;
    }
  }
  if (e.qvec) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float a = (col0 + i < N) ? tanhf(v[i]) : 0.f;
      v[i] = a;
      if (col0 + i < N) score_acc += a * __ldg(e.qvec + col0 + i);
    }
    if (e.tanh_out) {
      float* dst = e.tanh_out + (long long)row * e.ld_tanh + col0;
      if (col0 + 16 <= N && (e.ld_tanh & 3) == 0) {
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (col0 + i < N) dst[i] = v[i];
      }
    }
  }
  if (e.out) {
    float* dst = e.out + (long long)row * e.ld_out + col0;
    if (col0 + 16 <= e.out_cols && (e.ld_out & 3) == 0) {
#pragma unroll
      for (int i = 0; i < 16; i += 4)
        *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (col0 + i < e.out_cols) dst[i] = v[i];
    }
  }
  if (e.hi) {
    if (col0 < e.sp_cols) {
      uint32_t h[8], l[8];
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        float a = (col0 + i < N) ? v[i] : (col0 + i == e.ones_col ? 1.f : 0.f);
        float b = (col0 + i + 1 < N) ? v[i + 1] : (col0 + i + 1 == e.ones_col ? 1.f : 0.f);
        __nv_bfloat16 ah, al, bh, bl;
        split_bf16(a, ah, al);
        split_bf16(b, bh, bl);
        h[i >> 1] = pack_bf16x2(ah, bh);
        l[i >> 1] = pack_bf16x2(al, bl);
      }
      // sp_cols and ld_sp are multiples of 8 -> 16-byte aligned half-chunks
      long long off = (long long)row * e.ld_sp + col0;
      if (col0 + 8 <= e.sp_cols) {
        *reinterpret_cast<uint4*>(e.hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
        if (e.lo) *reinterpret_cast<uint4*>(e.lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
      }
      if (col0 + 16 <= e.sp_cols) {
        *reinterpret_cast<uint4*>(e.hi + off + 8) = make_uint4(h[4], h[5], h[6], h[7]);
        if (e.lo) *reinterpret_cast<uint4*>(e.lo + off + 8) = make_uint4(l[4], l[5], l[6], l[7]);
      }
    }
  }
  if (e.gw) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      int c = col0 + i;
      if (c < e.gw_cols) atomicAdd(e.gw + (long long)row * e.ld_gw + c, v[i]);
      else if (c == e.gw_cols && e.gb) atomicAdd(e.gb + row, v[i]);
    }
  }
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
nrl_gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + GEMM_STAGES * GEMM_STAGE_BYTES;
  // barrier layout (8 B each): full[S], empty[S], tmem_full[2], tmem_empty[2], then tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (GEMM_STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_STAGES + 2 + s); };
  const uint32_t tmem_ptr_addr = bar_base + 8u * (2 * GEMM_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM;
  const int n_extent = (p.epi.hi && p.epi.sp_cols > p.N) ? p.epi.sp_cols : p.N;
  const int n_tiles = (n_extent + p.BN - 1) / p.BN;
  const int kb_total = (p.K + GEMM_BK - 1) / GEMM_BK;
  const int num_tiles = m_tiles * n_tiles * p.k_splits;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GEMM_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), 4);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_addr, GEMM_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_addr));

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t tx_bytes =
          p.mn_major ? (uint32_t)(2 + p.BN / 64) * 8192u : (uint32_t)(GEMM_BM + p.BN) * 128u;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m0, n0, kb0, kb1;
        gemm_tile_coords(p, tile, m_tiles, n_tiles, kb_total, m0, n0, kb0, kb1);
        if (kb0 >= kb1) continue;
        for (int s = 0; s < p.num_segs; ++s) {
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1u);
            mbar_expect_tx(full_bar(stage), tx_bytes);
            const uint32_t a_dst = smem_base + stage * GEMM_STAGE_BYTES;
            const uint32_t b_dst = a_dst + GEMM_A_BYTES;
            if (!p.mn_major) {
              tma_load_3d(a_dst, &tmA, full_bar(stage), kb * GEMM_BK, m0, p.seg_a[s]);
              tma_load_3d(b_dst, &tmB, full_bar(stage), kb * GEMM_BK, n0, p.seg_b[s]);
            } else {
              for (int j = 0; j < 2; ++j)
                tma_load_3d(a_dst + j * 8192, &tmA, full_bar(stage), m0 + j * 64, kb * GEMM_BK,
                            p.seg_a[s]);
              for (int j = 0; j < p.BN / 64; ++j)
                tma_load_3d(b_dst + j * 8192, &tmB, full_bar(stage), n0 + j * 64, kb * GEMM_BK,
                            p.seg_b[s]);
            }
            if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint32_t idesc = umma_idesc_bf16(GEMM_BM, p.BN, p.mn_major, p.mn_major);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        int m0, n0, kb0, kb1;
        gemm_tile_coords(p, tile, m_tiles, n_tiles, kb_total, m0, n0, kb0, kb1);
        if (kb0 >= kb1) continue;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
        uint32_t accumulate = 0;
        for (int s = 0; s < p.num_segs; ++s) {
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            const uint32_t a_src = smem_base + stage * GEMM_STAGE_BYTES;
            const uint32_t b_src = a_src + GEMM_A_BYTES;
            const int nks = min(GEMM_BK / 16, (p.K - kb * GEMM_BK + 15) / 16);
            for (int k = 0; k < nks; ++k) {
              uint64_t ad, bd;
              if (!p.mn_major) {
                ad = umma_desc_sw128(a_src + k * 32, 16, 1024);
                bd = umma_desc_sw128(b_src + k * 32, 16, 1024);
              } else {
                ad = umma_desc_sw128(a_src + k * 2048, 8192, 1024);
                bd = umma_desc_sw128(b_src + k * 2048, 8192, 1024);
              }
              umma_bf16(d_tmem, ad, bd, idesc, accumulate);
              accumulate = 1;
            }
            umma_commit(empty_bar(stage));
            if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1u; }
          }
        }
        umma_commit(tfull_bar(acc));
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // ===================== epilogue warps (2..5) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may touch
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      int m0, n0, kb0, kb1;
      gemm_tile_coords(p, tile, m_tiles, n_tiles, kb_total, m0, n0, kb0, kb1);
      if (kb0 >= kb1) continue;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t t_row = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(quarter * 32) << 16);
      float score_acc = 0.f;
      // columns that any sink of this tile can consume
      int col_end = min(p.BN, max(max(p.N, p.epi.hi ? p.epi.sp_cols : 0), p.epi.gw ? p.epi.gw_cols + 1 : 0) - n0);
      for (int c = 0; c < col_end; c += 16) {
        float v[16];
        tmem_ld16(t_row + (uint32_t)c, v);
        gemm_epilogue_chunk(p.epi, row, n0 + c, p.N, row_ok, v, score_acc);
      }
      if (p.epi.score && row_ok) p.epi.score[row] = score_acc;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, GEMM_TMEM_COLS);
  }
}

}  // namespace nrl
