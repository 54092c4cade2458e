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
constexpr int GEMM_EPI_STAGE_BYTES = 4 * (32 * 33 * 4 + 32 * 4);  // per epilogue warp: [32][33] fp32 + 32 keep words
constexpr int GEMM_SMEM_BYTES =
    GEMM_STAGES * GEMM_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + GEMM_EPI_STAGE_BYTES;
constexpr int GEMM_TMEM_COLS = 512;

struct GemmEpi {
  // x = acc (+ addend) ; optional dropout ; then any of the sinks below.
  const float* addend;  long long ld_add;             // x += addend[row, col]
  const float* add_w;   const float* add_vec; long long ld_addvec; int add_L;
                                                      // x += add_w[row] * add_vec[row / add_L, col]
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

// keep-bits of 32 consecutive elements starting at flat index e0 (bit i = element e0 + i)
__device__ __forceinline__ uint32_t drop_keep_bits32(unsigned long long seed, uint32_t site,
                                                     unsigned long long e0, uint32_t thr) {
  uint32_t bits = 0;
  unsigned long long g = e0 >> 3;
  int slot = (int)(e0 & 7ull);
  Philox4 r = philox4x32_10(seed, g, site);
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    if (philox_u16(r, slot) >= thr) bits |= (1u << i);
    if (++slot == 8 && i != 31) {
      slot = 0;
      r = philox4x32_10(seed, ++g, site);
    }
  }
  return bits;
}

// Epilogue of one 32-row x (<=32)-column block of the accumulator, executed by one warp.
// Phase 1 (thread = row, straight out of TMEM): tanh / query-dot and the dropout keep-bits;
// the block is then transposed through a padded smem stage so that in phase 2 (lane = column)
// every global load / store / atomic of the warp covers one contiguous row segment.
__device__ __forceinline__ void gemm_epilogue_block(const GemmEpi& e, int M, int N, int row_base,
                                                    int col_base, int ncol, float* v,
                                                    float* stage /*[32][33]*/, uint32_t* kbits /*[32]*/,
                                                    float& score_acc) {
  const int lane = threadIdx.x & 31;
  const int my_row = row_base + lane;
  if (e.qvec) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int col = col_base + i;
      float a = 0.f;
      if (i < ncol && col < N) {
        a = tanhf(v[i]);
        score_acc += a * __ldg(e.qvec + col);
      }
      v[i] = a;
    }
  }
  if (e.use_dropout) {
    kbits[lane] = drop_keep_bits32(e.seed, e.drop_site,
                                   (unsigned long long)my_row * (unsigned)e.drop_ld + (unsigned)col_base,
                                   e.drop_thr);
  }
#pragma unroll
  for (int i = 0; i < 32; ++i) stage[lane * 33 + i] = v[i];
  __syncwarp();
  const int col = col_base + lane;
  const bool lane_ok = lane < ncol;
  const int rows = min(32, M - row_base);
  for (int r = 0; r < rows; ++r) {
    const long long row = row_base + r;
    float x = stage[r * 33 + lane];
    if (e.addend && lane_ok && col < N) x += __ldg(e.addend + row * e.ld_add + col);
    if (e.add_w && lane_ok && col < N)
      x += __ldg(e.add_w + row) * __ldg(e.add_vec + (row / e.add_L) * e.ld_addvec + col);
    if (e.use_dropout) x = ((kbits[r] >> lane) & 1u) ? x * e.drop_scale : 0.f;
    if (!lane_ok) continue;
    if (e.tanh_out && col < N) e.tanh_out[row * e.ld_tanh + col] = x;
    if (e.out && col < e.out_cols) e.out[row * e.ld_out + col] = x;
    if (e.hi && col < e.sp_cols) {
      const float val = col < N ? x : (col == e.ones_col ? 1.f : 0.f);
      __nv_bfloat16 h, l;
      split_bf16(val, h, l);
      e.hi[row * e.ld_sp + col] = h;
      if (e.lo) e.lo[row * e.ld_sp + col] = l;
    }
    if (e.gw) {
      if (col < e.gw_cols) atomicAdd(e.gw + row * e.ld_gw + col, x);
      else if (col == e.gw_cols && e.gb) atomicAdd(e.gb + row, x);
    }
  }
  __syncwarp();
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
  float* epi_stage = reinterpret_cast<float*>(smem_raw + (smem_base - smem_u32(smem_raw)) +
                                              GEMM_STAGES * GEMM_STAGE_BYTES + 256);

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
      const int row_base = m0 + quarter * 32;
      const uint32_t t_row = tmem_base + (uint32_t)acc * 256u + ((uint32_t)(quarter * 32) << 16);
      float score_acc = 0.f;
      // columns that any sink of this tile can consume (multiple of 16)
      const int col_end = min(p.BN, ((max(max(p.N, p.epi.hi ? p.epi.sp_cols : 0), p.epi.gw ? p.epi.gw_cols + 1 : 0) - n0) + 15) & ~15);
      float* stage = epi_stage + (warp - 2) * (32 * 33 + 32);
      uint32_t* kbits = reinterpret_cast<uint32_t*>(stage + 32 * 33);
      for (int c = 0; c < col_end; c += 32) {
        const int ncol = min(32, col_end - c);
        float v[32];
        tmem_ld16(t_row + (uint32_t)c, v);
        if (ncol > 16) tmem_ld16(t_row + (uint32_t)c + 16u, v + 16);
        else {
#pragma unroll
          for (int i = 16; i < 32; ++i) v[i] = 0.f;
        }
        if (row_base < p.M)
          gemm_epilogue_block(p.epi, p.M, p.N, row_base, n0 + c, ncol, v, stage, kbits, score_acc);
      }
      if (p.epi.score && row_base + lane < p.M) p.epi.score[row_base + lane] = score_acc;
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
