// the transformer inside the PLM news encoder (HF RobertaModel / BertModel layer stack,
// reference text.py:67-73,92) on sm_100a.  Everything that is not a tcgen05 GEMM (nrl_gemm.cuh) lives here:
//   * weight packing into the GEMM's split-plane operands (q / k / v rows stacked into one in-projection),
//   * position ids + key-padding bytes from input_ids / attention_mask,
//   * embeddings (word + position + token-type -> LayerNorm -> dropout) forward and backward,
//   * LayerNorm forward (fp32 row + bf16 hi/lo planes of the next GEMM) and backward,
//   * key-padding-masked multi-head attention for S <= 128, head dim 64, on warp-level tensor-core MMAs
//     (mma.sync.m16n8k16 bf16, hi/lo three-pass = fp32-equivalent), operands staged in shared memory as bf16 hi/lo
//     planes and fetched with ldmatrix, attention-probability dropout from a per-(title, head) Philox bit matrix.
// Post-LN layer (RobertaLayer.forward):  h1 = LN(x + drop(attn(x) W_ao + b)),  h2 = LN(h1 + drop(gelu(h1 W_i + b) W_o + b)).
#pragma once
#include "nrl_attn_mma.cuh"

namespace nrl {

// ------------------------------------------------------------------------------------
// weight packing: W [n_out, k_in] (+ bias) -> forward operand wf [planes][n_total][kp] rows n_off.. (bias at column
// k_in) and transposed operand wt [planes][k_in][np] columns n_off.. (data-gradient GEMMs)
// ------------------------------------------------------------------------------------
struct TfmPackJob {
  const float* W; const float* bias;
  int n_out, k_in, kp, np, n_off, n_total;
  __nv_bfloat16 *wf, *wt;
};
struct TfmPackJobs {
  TfmPackJob j[6];
};
// blockIdx.y = job, blockIdx.z = 0: forward operand (coalesced both ways), 1: transposed operand through a 32 x 33
// shared-memory tile (W is read along k, wt is written along n: both coalesced)
__global__ void __launch_bounds__(256) tfm_pack_kernel(const TfmPackJobs jobs, int two_planes) {
  const TfmPackJob& q = jobs.j[blockIdx.y];
  if (!q.W) return;
  const long long plane_f = (long long)q.n_total * q.kp, plane_t = (long long)q.k_in * q.np;
  if (blockIdx.z == 0) {  // one output row per CTA pass: no per-element index divisions
    for (int n = blockIdx.x; n < q.n_out; n += gridDim.x) {
      const float* src = q.W + (long long)n * q.k_in;
      const long long row = (long long)(q.n_off + n) * q.kp;
      for (int k = threadIdx.x; k < q.kp; k += blockDim.x) {
        const float x = k < q.k_in ? src[k] : (k == q.k_in && q.bias ? q.bias[n] : 0.f);
        __nv_bfloat16 h, l;
        split_bf16(x, h, l);
        q.wf[row + k] = h;
        if (two_planes) q.wf[plane_f + row + k] = l;
      }
    }
    return;
  }
  __shared__ float tile[32][33];
  const bool last = q.n_off + q.n_out == q.n_total;
  const int tcols = last ? q.np - q.n_off : q.n_out;  // the last job also zeroes the pad columns of wt
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int kt = (q.k_in + 31) / 32, nt = (tcols + 31) / 32;
  for (int t = blockIdx.x; t < kt * nt; t += gridDim.x) {
    const int k0 = (t % kt) * 32, n0 = (t / kt) * 32;
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int n = n0 + r, k = k0 + tx;
      tile[r][tx] = (n < q.n_out && k < q.k_in) ? q.W[(long long)n * q.k_in + k] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
      const int k = k0 + r, n = n0 + tx;
      if (k < q.k_in && n < tcols) {
        __nv_bfloat16 h, l;
        split_bf16(tile[tx][r], h, l);
        const long long off = (long long)k * q.np + q.n_off + n;
        q.wt[off] = h;
        if (two_planes) q.wt[plane_t + off] = l;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// position ids (RobertaEmbeddings.create_position_ids_from_input_ids: cumsum(ids != pad) * (ids != pad) + pad;
// BERT-style absolute positions when pad_idx < 0: pos = t) and key-padding bytes from attention_mask.
// One warp per title.
// ------------------------------------------------------------------------------------
__global__ void tfm_prepare_kernel(const long long* __restrict__ ids, const long long* __restrict__ attn_mask,
                                   int N, int T, int pad_idx, int* __restrict__ pos, unsigned char* __restrict__ kmask) {
  const int lane = threadIdx.x & 31;
  const int warp0 = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  const int nwarps = (int)(((long long)gridDim.x * blockDim.x) >> 5);
  for (int n = warp0; n < N; n += nwarps) {
    int run = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
      const int t = t0 + lane;
      const int nz = (t < T && pad_idx >= 0 && ids[(long long)n * T + t] != pad_idx) ? 1 : 0;
      int inc = nz;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      if (t < T) {
        pos[(long long)n * T + t] = pad_idx >= 0 ? (run + inc) * nz + pad_idx : t;
        kmask[(long long)n * T + t] = (!attn_mask || attn_mask[(long long)n * T + t] != 0) ? 1 : 0;
      }
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
}

// ------------------------------------------------------------------------------------
// LayerNorm rows.  One warp per row, the row in registers as float4 chunks (D % 4 == 0, D <= 1024); torch
// semantics: biased variance, eps inside the square root.
// ------------------------------------------------------------------------------------
constexpr int TFM_LN_CHUNKS = 8;  // float4 chunks per lane: D <= 32 * 4 * 8

struct LnRow {
  float4 v[TFM_LN_CHUNKS];
};
__device__ __forceinline__ void ln_stats(const LnRow& r, int D, int lane, float& mean, float& rstd, float eps) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < TFM_LN_CHUNKS; ++i)
    if (lane * 4 + 128 * i < D) s += (r.v[i].x + r.v[i].y) + (r.v[i].z + r.v[i].w);
  mean = warp_sum(s) / (float)D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < TFM_LN_CHUNKS; ++i)
    if (lane * 4 + 128 * i < D) {
      const float a = r.v[i].x - mean, b = r.v[i].y - mean, c = r.v[i].z - mean, d = r.v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  rstd = rsqrtf(warp_sum(q) / (float)D + eps);
}
// y (and / or planes with the ones column at D) of one normalised row; keep-bit words optional (dropout AFTER the norm)
__device__ __forceinline__ void ln_emit(const LnRow& r, int D, int dp, int lane, float mean, float rstd,
                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                        const uint32_t* __restrict__ words, float dscale, float* __restrict__ y,
                                        __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
#pragma unroll
  for (int i = 0; i < TFM_LN_CHUNKS; ++i) {
    const int c = lane * 4 + 128 * i;
    if (c < D) {
      const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c));
      float o[4] = {(r.v[i].x - mean) * rstd * g4.x + b4.x, (r.v[i].y - mean) * rstd * g4.y + b4.y,
                    (r.v[i].z - mean) * rstd * g4.z + b4.z, (r.v[i].w - mean) * rstd * g4.w + b4.w};
      if (words) {
        const uint32_t bits = __ldg(words + (c >> 5)) >> (c & 31);
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] = ((bits >> k) & 1u) ? o[k] * dscale : 0.f;
      }
      if (y) *reinterpret_cast<float4*>(y + c) = make_float4(o[0], o[1], o[2], o[3]);
      if (hi) {
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) split_bf16(o[k], h[k], l[k]);
        *reinterpret_cast<uint2*>(hi + c) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
        if (lo) *reinterpret_cast<uint2*>(lo + c) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
      }
    }
  }
  if (hi)
    for (int c = D + lane; c < dp; c += 32) {
      hi[c] = __float2bfloat16_rn(c == D ? 1.f : 0.f);
      if (lo) lo[c] = __float2bfloat16_rn(0.f);
    }
}

// y = LayerNorm(s) * gamma + beta: fp32 rows and / or bf16 hi / lo planes [R][dp] (ones column at D)
__global__ void __launch_bounds__(256)
tfm_ln_fwd_kernel(const float* __restrict__ s, long long R, int D, int dp, const float* __restrict__ gamma,
                  const float* __restrict__ beta, float eps, float* __restrict__ y, __nv_bfloat16* __restrict__ hi,
                  __nv_bfloat16* __restrict__ lo) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < R; r += nwarps) {
    LnRow row;
#pragma unroll
    for (int i = 0; i < TFM_LN_CHUNKS; ++i) {
      const int c = lane * 4 + 128 * i;
      row.v[i] = c < D ? *reinterpret_cast<const float4*>(s + r * D + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float mean, rstd;
    ln_stats(row, D, lane, mean, rstd, eps);
    ln_emit(row, D, dp, lane, mean, rstd, gamma, beta, nullptr, 1.f, y ? y + r * D : nullptr,
            hi ? hi + r * dp : nullptr, lo ? lo + r * dp : nullptr);
  }
}

// Backward of y = LN(s): ds = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dy * gamma.
//   ds_out  [R][D] fp32 (nullable): the gradient w.r.t. s -- the residual branch takes it as is;
//   dt planes [R][dtp] (nullable): ds with the keep-bit words of the dropout that FOLLOWED the producing GEMM applied
//           (the A operand of that GEMM's data- and weight-gradient products);
//   dgamma / dbeta (PG: the layer is trainable) accumulated per CTA, then one atomic per column.
// 128-thread CTAs; the frozen-layer instance carries no accumulators (half the registers): 16 warps per SM against 12.
template <bool PG>
__global__ void __launch_bounds__(128, PG ? 3 : 4)
tfm_ln_bwd_kernel(const float* __restrict__ s, const float* __restrict__ dy, long long R, int D,
                  const float* __restrict__ gamma, float eps, float* __restrict__ ds_out,
                  __nv_bfloat16* __restrict__ dt_hi, __nv_bfloat16* __restrict__ dt_lo, int dtp,
                  const uint32_t* __restrict__ words, int mw, float dscale, float* __restrict__ dgamma,
                  float* __restrict__ dbeta) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float4 ag[TFM_LN_CHUNKS], ab[TFM_LN_CHUNKS];
#pragma unroll
  for (int i = 0; i < TFM_LN_CHUNKS; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = warp0; r < R; r += nwarps) {
    LnRow row;
    float4 g[TFM_LN_CHUNKS];
#pragma unroll
    for (int i = 0; i < TFM_LN_CHUNKS; ++i) {
      const int c = lane * 4 + 128 * i;
      row.v[i] = c < D ? *reinterpret_cast<const float4*>(s + r * D + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      g[i] = c < D ? *reinterpret_cast<const float4*>(dy + r * D + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // ONE pass, four sums reduced together (their shuffle chains interleave: one reduction latency instead of four
    // dependent ones).  The row is shifted by its first element so that sum (x - k)^2 - (sum (x - k))^2 / D does not
    // cancel when |mean| >> std;  sum g xhat = rstd (sum g x' - mean' sum g) with the same shifted values.
    const float shift = __shfl_sync(0xffffffffu, row.v[0].x, 0);
    float sx = 0.f, sxx = 0.f, sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < TFM_LN_CHUNKS; ++i) {
      const int c = lane * 4 + 128 * i;
      if (c < D) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(gamma + c));
        float4& x = row.v[i];
        x.x -= shift; x.y -= shift; x.z -= shift; x.w -= shift;
        sx += (x.x + x.y) + (x.z + x.w);
        sxx += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
        const float4 gw = make_float4(g[i].x * w4.x, g[i].y * w4.y, g[i].z * w4.z, g[i].w * w4.w);
        sg += (gw.x + gw.y) + (gw.z + gw.w);
        sgx += (gw.x * x.x + gw.y * x.y) + (gw.z * x.z + gw.w * x.w);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      sx += __shfl_xor_sync(0xffffffffu, sx, o);
      sxx += __shfl_xor_sync(0xffffffffu, sxx, o);
      sg += __shfl_xor_sync(0xffffffffu, sg, o);
      sgx += __shfl_xor_sync(0xffffffffu, sgx, o);
    }
    const float inv_d = 1.f / (float)D;
    const float mean = sx * inv_d;  // of the shifted row
    const float rstd = rsqrtf(fmaxf(sxx * inv_d - mean * mean, 0.f) + eps);
    const float mg = sg * inv_d, mgx = rstd * (sgx - mean * sg) * inv_d;
#pragma unroll
    for (int i = 0; i < TFM_LN_CHUNKS; ++i) {
      const int c = lane * 4 + 128 * i;
      if (c < D) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(gamma + c));
        float4& x = row.v[i];
        x.x = (x.x - mean) * rstd; x.y = (x.y - mean) * rstd; x.z = (x.z - mean) * rstd; x.w = (x.w - mean) * rstd;
        if (PG) {
          ag[i].x += g[i].x * x.x; ag[i].y += g[i].y * x.y; ag[i].z += g[i].z * x.z; ag[i].w += g[i].w * x.w;
          ab[i].x += g[i].x; ab[i].y += g[i].y; ab[i].z += g[i].z; ab[i].w += g[i].w;
        }
        g[i].x *= w4.x; g[i].y *= w4.y; g[i].z *= w4.z; g[i].w *= w4.w;
      }
    }
#pragma unroll
    for (int i = 0; i < TFM_LN_CHUNKS; ++i) {
      const int c = lane * 4 + 128 * i;
      if (c < D) {
        const float4& x = row.v[i];
        float o[4] = {rstd * (g[i].x - mg - x.x * mgx), rstd * (g[i].y - mg - x.y * mgx),
                      rstd * (g[i].z - mg - x.z * mgx), rstd * (g[i].w - mg - x.w * mgx)};
        if (ds_out) *reinterpret_cast<float4*>(ds_out + r * D + c) = make_float4(o[0], o[1], o[2], o[3]);
        if (dt_hi) {
          if (words) {
            const uint32_t bits = __ldg(words + r * mw + (c >> 5)) >> (c & 31);
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = ((bits >> k) & 1u) ? o[k] * dscale : 0.f;
          }
          __nv_bfloat16 h[4], l[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) split_bf16(o[k], h[k], l[k]);
          *reinterpret_cast<uint2*>(dt_hi + r * dtp + c) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
          if (dt_lo)
            *reinterpret_cast<uint2*>(dt_lo + r * dtp + c) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
        }
      }
    }
    if (dt_hi)
      for (int c = D + lane; c < dtp; c += 32) {
        dt_hi[r * dtp + c] = __float2bfloat16_rn(0.f);
        if (dt_lo) dt_lo[r * dtp + c] = __float2bfloat16_rn(0.f);
      }
  }
  if (PG) {  // warps -> shared-memory accumulators -> ONE global atomic per column per CTA (the launch is capped at
                 // two CTAs per SM: thousands of warps adding into the same 2 D addresses serialise in L2)
    extern __shared__ float ln_acc[];  // [2][D]
    for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) ln_acc[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < TFM_LN_CHUNKS; ++i) {
      const int c = lane * 4 + 128 * i;
      if (c < D) {
        atomicAdd(ln_acc + c, ag[i].x); atomicAdd(ln_acc + c + 1, ag[i].y);
        atomicAdd(ln_acc + c + 2, ag[i].z); atomicAdd(ln_acc + c + 3, ag[i].w);
        atomicAdd(ln_acc + D + c, ab[i].x); atomicAdd(ln_acc + D + c + 1, ab[i].y);
        atomicAdd(ln_acc + D + c + 2, ab[i].z); atomicAdd(ln_acc + D + c + 3, ab[i].w);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
      atomicAdd(dgamma + i, ln_acc[i]);
      atomicAdd(dbeta + i, ln_acc[D + i]);
    }
  }
}

// ------------------------------------------------------------------------------------
// embeddings (RobertaEmbeddings.forward): e = word[id] + position[pos] + token_type[0]; x0 = drop(LN(e)).
// Writes the fp32 rows (residual of layer 0) and the planes (its in-projection operand).  One warp per token.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ bool tfm_embed_row(LnRow& row, long long r, int lane, const long long* __restrict__ ids,
                                              const int* __restrict__ pos, const float* __restrict__ word, long long V,
                                              const float* __restrict__ pe, int P, const float* __restrict__ type0,
                                              int D, long long& id, int& p) {
  id = ids[r];
  p = pos[r];
  bool ok = true;
  if (id < 0 || id >= V || p < 0 || p >= P) {  // nn.Embedding raises here: flagged, row 0 read instead
    if (lane == 0) dev_error(DEV_ERR_TOKEN_ID);
    id = 0; p = 0; ok = false;
  }
#pragma unroll
  for (int i = 0; i < TFM_LN_CHUNKS; ++i) {
    const int c = lane * 4 + 128 * i;
    if (c < D) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(word + id * D + c));
      const float4 b = __ldg(reinterpret_cast<const float4*>(pe + (long long)p * D + c));
      const float4 t = __ldg(reinterpret_cast<const float4*>(type0 + c));
      // HF order: (inputs_embeds + token_type_embeddings) + position_embeddings
      row.v[i] = make_float4((a.x + t.x) + b.x, (a.y + t.y) + b.y, (a.z + t.z) + b.z, (a.w + t.w) + b.w);
    } else {
      row.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  return ok;
}
__global__ void __launch_bounds__(256)
tfm_embed_fwd_kernel(const long long* __restrict__ ids, const int* __restrict__ pos, long long R,
                     const float* __restrict__ word, long long V, const float* __restrict__ pe, int P,
                     const float* __restrict__ type0, int D, int dp, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, const uint32_t* __restrict__ words, int mw,
                     float dscale, float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < R; r += nwarps) {
    LnRow row;
    long long id; int p;
    tfm_embed_row(row, r, lane, ids, pos, word, V, pe, P, type0, D, id, p);
    float mean, rstd;
    ln_stats(row, D, lane, mean, rstd, eps);
    ln_emit(row, D, dp, lane, mean, rstd, gamma, beta, words ? words + r * mw : nullptr, dscale, x + r * D,
            hi + r * dp, lo ? lo + r * dp : nullptr);
  }
}
// backward: g = dx0 (keep-bits applied) -> LN backward -> de; d_word[id] += de (padding_idx row skipped),
// d_pos[pos] += de (RoBERTa: its padding_idx row skipped; BERT: pos_pad_idx = -1, every row), d_type[0] += sum de,
// dgamma / dbeta.
__global__ void __launch_bounds__(256)
tfm_embed_bwd_kernel(const long long* __restrict__ ids, const int* __restrict__ pos, long long R,
                     const float* __restrict__ word, long long V, const float* __restrict__ pe, int P,
                     const float* __restrict__ type0, int D, const float* __restrict__ gamma, float eps,
                     const uint32_t* __restrict__ words, int mw, float dscale, int pad_idx, int pos_pad_idx,
                     const float* __restrict__ dx, float* __restrict__ d_word, float* __restrict__ d_pos,
                     float* __restrict__ d_type, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  float4 ag[TFM_LN_CHUNKS], ab[TFM_LN_CHUNKS], at[TFM_LN_CHUNKS];
#pragma unroll
  for (int i = 0; i < TFM_LN_CHUNKS; ++i) ag[i] = ab[i] = at[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long r = warp0; r < R; r += nwarps) {
    LnRow row;
    long long id; int p;
    const bool ok = tfm_embed_row(row, r, lane, ids, pos, word, V, pe, P, type0, D, id, p);
    float mean, rstd;
    ln_stats(row, D, lane, mean, rstd, eps);
    float4 g[TFM_LN_CHUNKS];
    float sg = 0.f, sgx = 0.f;
#pragma unroll
    for (int i = 0; i < TFM_LN_CHUNKS; ++i) {
      const int c = lane * 4 + 128 * i;
      g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < D) {
        g[i] = *reinterpret_cast<const float4*>(dx + r * D + c);
        if (words) {
          const uint32_t bits = __ldg(words + r * mw + (c >> 5)) >> (c & 31);
          g[i].x = (bits & 1u) ? g[i].x * dscale : 0.f; g[i].y = (bits & 2u) ? g[i].y * dscale : 0.f;
          g[i].z = (bits & 4u) ? g[i].z * dscale : 0.f; g[i].w = (bits & 8u) ? g[i].w * dscale : 0.f;
        }
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(gamma + c));
        float4& x = row.v[i];
        x.x = (x.x - mean) * rstd; x.y = (x.y - mean) * rstd; x.z = (x.z - mean) * rstd; x.w = (x.w - mean) * rstd;
        ag[i].x += g[i].x * x.x; ag[i].y += g[i].y * x.y; ag[i].z += g[i].z * x.z; ag[i].w += g[i].w * x.w;
        ab[i].x += g[i].x; ab[i].y += g[i].y; ab[i].z += g[i].z; ab[i].w += g[i].w;
        g[i].x *= w4.x; g[i].y *= w4.y; g[i].z *= w4.z; g[i].w *= w4.w;
        sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        sgx += (g[i].x * x.x + g[i].y * x.y) + (g[i].z * x.z + g[i].w * x.w);
      }
    }
    const float mg = warp_sum(sg) / (float)D, mgx = warp_sum(sgx) / (float)D;
#pragma unroll
    for (int i = 0; i < TFM_LN_CHUNKS; ++i) {
      const int c = lane * 4 + 128 * i;
      if (c < D) {
        const float4& x = row.v[i];
        const float4 de = make_float4(rstd * (g[i].x - mg - x.x * mgx), rstd * (g[i].y - mg - x.y * mgx),
                                      rstd * (g[i].z - mg - x.z * mgx), rstd * (g[i].w - mg - x.w * mgx));
        at[i].x += de.x; at[i].y += de.y; at[i].z += de.z; at[i].w += de.w;
        if (ok && id != pad_idx) atomicAdd(reinterpret_cast<float4*>(d_word + id * D + c), de);
        if (ok && p != pos_pad_idx) atomicAdd(reinterpret_cast<float4*>(d_pos + (long long)p * D + c), de);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < TFM_LN_CHUNKS; ++i) {
    const int c = lane * 4 + 128 * i;
    if (c < D) {
      atomicAdd(reinterpret_cast<float4*>(dgamma + c), ag[i]);
      atomicAdd(reinterpret_cast<float4*>(dbeta + c), ab[i]);
      atomicAdd(reinterpret_cast<float4*>(d_type + c), at[i]);
    }
  }
}

// ------------------------------------------------------------------------------------
// attention (RobertaSelfAttention, eager math): per (title n, head h)
//     P = softmax(q k^T / sqrt(d_h) + key-padding mask);  ctx = dropout(P) v
// One CTA per (n, h); Q (pre-scaled by log2(e) / sqrt(d_h)), K, V (and dO) of the head are staged in shared memory as
// bf16 hi / lo planes [SK][72] (144-byte pitch: the eight 16-byte rows of an ldmatrix tile fall into distinct bank
// groups); warp w owns query rows 16 w .. 16 w + 15 (forward, backward phase A) or key rows (backward phase B).
// Every contraction is mma.sync.m16n8k16 (bf16 in, fp32 accumulate), three passes (lo*hi + hi*lo + hi*hi) unless the
// library runs in single-pass bf16.  Masked keys (attention_mask == 0) and the padding up to SK take no part in any
// softmax; a title with NO valid key yields ctx = 0 (HF would return the mean of v: never happens with <s> ... </s>).
// ------------------------------------------------------------------------------------
constexpr int TFM_DH = 64;
constexpr int TFM_PITCH = 72;                    // bf16 elements per staged row
constexpr int TFM_PLANE_ROW_BYTES = TFM_PITCH * 2;
constexpr float TFM_LOG2E = 1.4426950408889634f, TFM_LN2 = 0.6931471805599453f;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// A fragment (16 rows x 16 k) of a staged row-major matrix: rows row0.., columns k0..
__device__ __forceinline__ uint32_t tfm_a_addr(uint32_t base, int row0, int k0, int lane) {
  const int mi = lane >> 3;
  return base + (uint32_t)((row0 + (lane & 7) + 8 * (mi & 1)) * TFM_PLANE_ROW_BYTES + (k0 + 8 * (mi >> 1)) * 2);
}
// B fragments of the product with the TRANSPOSE of a staged matrix M [n][k] (scores = A M^T): n-tiles n0, n0 + 8 at k0
//   r[0], r[1] = (b0, b1) of n-tile n0;  r[2], r[3] = those of n-tile n0 + 8
__device__ __forceinline__ uint32_t tfm_bt_addr(uint32_t base, int n0, int k0, int lane) {
  const int mi = lane >> 3;
  return base + (uint32_t)((n0 + (lane & 7) + 8 * (mi >> 1)) * TFM_PLANE_ROW_BYTES + (k0 + 8 * (mi & 1)) * 2);
}
// B fragments of the product with a staged matrix M [k][n] itself (ctx = P M), loaded with .trans: k rows k0.. (16),
// n-tiles n0, n0 + 8:  r[0], r[1] = (b0, b1) of n-tile n0;  r[2], r[3] = those of n-tile n0 + 8
__device__ __forceinline__ uint32_t tfm_bn_addr(uint32_t base, int k0, int n0, int lane) {
  const int mi = lane >> 3;
  return base + (uint32_t)((k0 + (lane & 7) + 8 * (mi & 1)) * TFM_PLANE_ROW_BYTES + (n0 + 8 * (mi >> 1)) * 2);
}
// c += A B with hi / lo planes of both operands (three passes) or hi only
__device__ __forceinline__ void tfm_mma3(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                         uint32_t bh1, uint32_t bl0, uint32_t bl1, bool three) {
  if (three) {
    mma_16816(c, al[0], al[1], al[2], al[3], bh0, bh1);
    mma_16816(c, ah[0], ah[1], ah[2], ah[3], bl0, bl1);
  }
  mma_16816(c, ah[0], ah[1], ah[2], ah[3], bh0, bh1);
}

// stage rows [0, S) x 64 columns of a row-major fp32 slice (pitch ld) as hi / lo planes, rows [S, SK) zero.
// NT = threads of the CTA (compile time): every thread issues ALL its 16-byte loads before the first conversion, so a
// matrix costs one trip to L2 / HBM instead of one per loop iteration.
template <int SK, int NT>
__device__ __forceinline__ void tfm_stage(uint8_t* smem, uint32_t plane_hi, uint32_t plane_lo, const float* __restrict__ src,
                                          long long ld, int S, float mul) {
  constexpr int ITER = SK * 16 / NT;
  static_assert(ITER * NT == SK * 16, "thread count must divide the tile");
  float4 v[ITER];
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int i = threadIdx.x + it * NT, t = i >> 4, c = (i & 15) * 4;
    v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < S) v[it] = __ldg(reinterpret_cast<const float4*>(src + (long long)t * ld + c));
  }
#pragma unroll
  for (int it = 0; it < ITER; ++it) {
    const int i = threadIdx.x + it * NT, t = i >> 4, c = (i & 15) * 4;
    uint32_t h0, l0, h1, l1;
    split_pack2(v[it].x * mul, v[it].y * mul, h0, l0);
    split_pack2(v[it].z * mul, v[it].w * mul, h1, l1);
    const uint32_t off = (uint32_t)(t * TFM_PLANE_ROW_BYTES + c * 2);
    *reinterpret_cast<uint2*>(smem + plane_hi + off) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(smem + plane_lo + off) = make_uint2(l0, l1);
  }
}
// dropout keep-bit matrix of one (title, head): bits[t][w] bit i = probability (query t, key 32 w + i) is kept
__device__ __forceinline__ void tfm_drop_bits(uint32_t* bits, int SK, unsigned long long seed, uint32_t site,
                                              unsigned long long item, uint32_t thr) {
  const int wpr = SK / 32;
  for (int i = threadIdx.x; i < SK * wpr; i += blockDim.x) {
    const int t = i / wpr, w = i % wpr;
    bits[i] = drop_keep_bits32(seed, site, (item * (unsigned long long)SK + (unsigned long long)t) * SK + 32ull * w, thr);
  }
}
__host__ __device__ inline int tfm_attn_fwd_smem(int SK) { return 6 * SK * TFM_PLANE_ROW_BYTES + SK + SK * (SK / 32) * 4 + 16; }
__host__ __device__ inline int tfm_attn_bwd_smem(int SK) {
  return 8 * SK * TFM_PLANE_ROW_BYTES + 2 * SK * 4 + SK + SK * (SK / 32) * 4 + 16;
}

// CTAs are launched with 64 * NK32 threads (two warps per 32 padded rows: warps beyond ceil(T / 16) only help staging);
// the minimum-CTA hints keep 12 warps per SM resident for every size but the largest
template <int NK32>  // padded key count SK = 32 * NK32 (<= 128)
__global__ void __launch_bounds__(64 * NK32, NK32 == 1 ? 6 : NK32 == 2 ? 3 : NK32 == 3 ? 2 : 1)
tfm_attn_fwd_kernel(const float* __restrict__ qkv, int ldq, int D, int H, int T, const unsigned char* __restrict__ kmask,
                    float scale, __nv_bfloat16* __restrict__ o_hi, __nv_bfloat16* __restrict__ o_lo, int dp,
                    float* __restrict__ lse, int three_i, int drop_on, uint32_t thr, float dscale,
                    unsigned long long seed, uint32_t site) {
  constexpr int SK = 32 * NK32, NT = SK / 8, NTHR = 64 * NK32;
  extern __shared__ __align__(16) uint8_t tsm[];
  const bool three = three_i != 0;
  const int n = blockIdx.x / H, h = blockIdx.x % H;
  const int S = T;
  const uint32_t plane = (uint32_t)SK * TFM_PLANE_ROW_BYTES;
  const uint32_t oQh = 0, oQl = plane, oKh = 2 * plane, oKl = 3 * plane, oVh = 4 * plane, oVl = 5 * plane;
  unsigned char* km = tsm + 6 * plane;
  uint32_t* bits = reinterpret_cast<uint32_t*>(tsm + ((6 * plane + SK + 15) & ~15u));
  const float* base = qkv + (long long)n * T * ldq + h * TFM_DH;
  tfm_stage<SK, NTHR>(tsm, oQh, oQl, base, ldq, S, scale * TFM_LOG2E);
  tfm_stage<SK, NTHR>(tsm, oKh, oKl, base + D, ldq, S, 1.f);
  tfm_stage<SK, NTHR>(tsm, oVh, oVl, base + 2 * D, ldq, S, 1.f);
  for (int u = threadIdx.x; u < SK; u += blockDim.x) km[u] = (u < S && kmask[(long long)n * T + u]) ? 1 : 0;
  if (drop_on) tfm_drop_bits(bits, SK, seed, site, (unsigned long long)blockIdx.x, thr);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int q0 = 16 * warp;
  if (q0 >= S) return;
  const uint32_t sb = smem_u32(tsm);
  // ---- scores (log2 domain) ----
  float s[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint32_t ah[4], al[4] = {0u, 0u, 0u, 0u};
    ldsm_x4(ah, tfm_a_addr(sb + oQh, q0, 16 * kk, lane));
    if (three) ldsm_x4(al, tfm_a_addr(sb + oQl, q0, 16 * kk, lane));
#pragma unroll
    for (int j2 = 0; j2 < NT / 2; ++j2) {
      uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
      ldsm_x4(bh, tfm_bt_addr(sb + oKh, 16 * j2, 16 * kk, lane));
      if (three) ldsm_x4(bl, tfm_bt_addr(sb + oKl, 16 * j2, 16 * kk, lane));
      tfm_mma3(s[2 * j2], ah, al, bh[0], bh[1], bl[0], bl[1], three);
      tfm_mma3(s[2 * j2 + 1], ah, al, bh[2], bh[3], bl[2], bl[3], three);
    }
  }
  // ---- masked softmax over the keys: rows r0 = q0 + g (values [0], [1]) and r1 = r0 + 8 ([2], [3]) ----
  uint32_t valid = 0;  // bit 2 j + e: key 8 j + 2 tg + e takes part
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int u = 8 * j + 2 * tg;
    valid |= (km[u] ? 1u : 0u) << (2 * j);
    valid |= (km[u + 1] ? 1u : 0u) << (2 * j + 1);
  }
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    if ((valid >> (2 * j)) & 1u) { m0 = fmaxf(m0, s[j][0]); m1 = fmaxf(m1, s[j][2]); }
    if ((valid >> (2 * j + 1)) & 1u) { m0 = fmaxf(m0, s[j][1]); m1 = fmaxf(m1, s[j][3]); }
  }
  m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
  m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
  if (m0 == -INFINITY) m0 = 0.f;
  if (m1 == -INFINITY) m1 = 0.f;
  float l0 = 0.f, l1 = 0.f;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const bool v0 = (valid >> (2 * j)) & 1u, v1 = (valid >> (2 * j + 1)) & 1u;
    s[j][0] = v0 ? ex2_approx(s[j][0] - m0) : 0.f; s[j][1] = v1 ? ex2_approx(s[j][1] - m0) : 0.f;
    s[j][2] = v0 ? ex2_approx(s[j][2] - m1) : 0.f; s[j][3] = v1 ? ex2_approx(s[j][3] - m1) : 0.f;
    l0 += s[j][0] + s[j][1]; l1 += s[j][2] + s[j][3];
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const int r0 = q0 + g, r1 = r0 + 8;
  float i0 = l0 > 0.f ? 1.f / l0 : 0.f, i1 = l1 > 0.f ? 1.f / l1 : 0.f;
  if (tg == 0) {
    if (r0 < S) lse[((long long)n * T + r0) * H + h] = l0 > 0.f ? (m0 + log2f(l0)) * TFM_LN2 : 0.f;
    if (r1 < S) lse[((long long)n * T + r1) * H + h] = l1 > 0.f ? (m1 + log2f(l1)) * TFM_LN2 : 0.f;
  }
  if (drop_on) { i0 *= dscale; i1 *= dscale; }
  // ---- ctx = dropout(P) V ----
  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < NT / 2; ++kk) {
    float p[2][4];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = 2 * kk + e;
      p[e][0] = s[j][0] * i0; p[e][1] = s[j][1] * i0; p[e][2] = s[j][2] * i1; p[e][3] = s[j][3] * i1;
      if (drop_on) {
        const int sh = 8 * (j & 3) + 2 * tg;
        const uint32_t b0 = bits[r0 * (SK / 32) + (j >> 2)] >> sh, b1 = bits[r1 * (SK / 32) + (j >> 2)] >> sh;
        if (!(b0 & 1u)) p[e][0] = 0.f;
        if (!(b0 & 2u)) p[e][1] = 0.f;
        if (!(b1 & 1u)) p[e][2] = 0.f;
        if (!(b1 & 2u)) p[e][3] = 0.f;
      }
    }
    uint32_t ah[4], al[4];
    split_pack2(p[0][0], p[0][1], ah[0], al[0]);
    split_pack2(p[0][2], p[0][3], ah[1], al[1]);
    split_pack2(p[1][0], p[1][1], ah[2], al[2]);
    split_pack2(p[1][2], p[1][3], ah[3], al[3]);
#pragma unroll
    for (int jd2 = 0; jd2 < 4; ++jd2) {
      uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
      ldsm_x4_t(bh, tfm_bn_addr(sb + oVh, 16 * kk, 16 * jd2, lane));
      if (three) ldsm_x4_t(bl, tfm_bn_addr(sb + oVl, 16 * kk, 16 * jd2, lane));
      tfm_mma3(o[2 * jd2], ah, al, bh[0], bh[1], bl[0], bl[1], three);
      tfm_mma3(o[2 * jd2 + 1], ah, al, bh[2], bh[3], bl[2], bl[3], three);
    }
  }
  // ---- planes of the context rows (columns h * 64 ..; head 0 also writes the ones / pad columns) ----
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = half ? r1 : r0;
    if (r >= S) continue;
    const long long rowoff = ((long long)n * T + r) * dp;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t hh, ll;
      split_pack2(o[j][2 * half], o[j][2 * half + 1], hh, ll);
      const long long off = rowoff + h * TFM_DH + 8 * j + 2 * tg;
      *reinterpret_cast<uint32_t*>(o_hi + off) = hh;
      if (o_lo) *reinterpret_cast<uint32_t*>(o_lo + off) = ll;
    }
    if (h == 0)
      for (int c = D + tg; c < dp; c += 4) {
        o_hi[rowoff + c] = __float2bfloat16_rn(c == D ? 1.f : 0.f);
        if (o_lo) o_lo[rowoff + c] = __float2bfloat16_rn(0.f);
      }
  }
}

// Backward: dqkv planes [R][p3] (dQ | dK | dV column sections of width D) from qkv, dO (fp32 [R][D]), the saved
// context planes (D_t = dO_t . ctx_t) and lse.
//   phase A (warp = 16 queries): P = exp(s - lse), dP = M c (dO V^T), dS = P (dP - D_t), dQ = dS K / sqrt(d_h)
//   phase B (warp = 16 keys):    the same tiles transposed (K Q^T, V dO^T), dV = (M c P)^T dO, dK = dS^T Q / sqrt(d_h)
// Both phases walk the other axis in blocks of 64 so that a warp never holds more than two 16 x 64 fp32 tiles.
template <int NK32>
__global__ void __launch_bounds__(64 * NK32, NK32 == 1 ? 6 : NK32 == 2 ? 3 : NK32 == 3 ? 2 : 1)
tfm_attn_bwd_kernel(const float* __restrict__ qkv, int ldq, int D, int H, int T, const unsigned char* __restrict__ kmask,
                    float scale, const float* __restrict__ d_o, const __nv_bfloat16* __restrict__ o_hi,
                    const __nv_bfloat16* __restrict__ o_lo, int dp, const float* __restrict__ lse,
                    __nv_bfloat16* __restrict__ g_hi, __nv_bfloat16* __restrict__ g_lo, int p3, int three_i, int drop_on,
                    uint32_t thr, float dscale, unsigned long long seed, uint32_t site) {
  constexpr int SK = 32 * NK32;
  constexpr int BW = (NK32 & 1) ? 32 : 64;  // keys (phase A) / queries (phase B) per block
  constexpr int BT = BW / 8;                // 8-wide tiles per block
  constexpr int NTHR = 64 * NK32;
  extern __shared__ __align__(16) uint8_t tsm[];
  const bool three = three_i != 0;
  const int n = blockIdx.x / H, h = blockIdx.x % H;
  const int S = T;
  const uint32_t plane = (uint32_t)SK * TFM_PLANE_ROW_BYTES;
  const uint32_t oQh = 0, oQl = plane, oKh = 2 * plane, oKl = 3 * plane, oVh = 4 * plane, oVl = 5 * plane,
                 oGh = 6 * plane, oGl = 7 * plane;
  float* lse2 = reinterpret_cast<float*>(tsm + 8 * plane);  // [SK] lse * log2(e); +inf beyond S
  float* dd = lse2 + SK;                                    // [SK] D_t
  unsigned char* km = reinterpret_cast<unsigned char*>(dd + SK);
  uint32_t* bits = reinterpret_cast<uint32_t*>(tsm + ((8 * plane + 2 * SK * 4 + SK + 15) & ~15u));
  const float* base = qkv + (long long)n * T * ldq + h * TFM_DH;
  tfm_stage<SK, NTHR>(tsm, oQh, oQl, base, ldq, S, scale * TFM_LOG2E);
  tfm_stage<SK, NTHR>(tsm, oKh, oKl, base + D, ldq, S, 1.f);
  tfm_stage<SK, NTHR>(tsm, oVh, oVl, base + 2 * D, ldq, S, 1.f);
  {  // dO with D_t = dO_t . ctx_t on the way (16 consecutive lanes own a row); all loads issued before the first use
    constexpr int ITER = SK * 16 / NTHR;
    float4 v[ITER];
    uint2 oh[ITER], ol[ITER];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int i = threadIdx.x + it * NTHR, t = i >> 4, c = (i & 15) * 4;
      v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      oh[it] = ol[it] = make_uint2(0u, 0u);
      if (t < S) {
        const long long row = (long long)n * T + t;
        v[it] = __ldg(reinterpret_cast<const float4*>(d_o + row * D + h * TFM_DH + c));
        oh[it] = *reinterpret_cast<const uint2*>(o_hi + row * dp + h * TFM_DH + c);
        if (o_lo) ol[it] = *reinterpret_cast<const uint2*>(o_lo + row * dp + h * TFM_DH + c);
      }
    }
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
      const int i = threadIdx.x + it * NTHR, t = i >> 4, c = (i & 15) * 4;
      float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&oh[it].x));
      float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&oh[it].y));
      const float2 a2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ol[it].x));
      const float2 b2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ol[it].y));
      a.x += a2.x; a.y += a2.y; b.x += b2.x; b.y += b2.y;
      float part = (v[it].x * a.x + v[it].y * a.y) + (v[it].z * b.x + v[it].w * b.y);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if ((i & 15) == 0) dd[t] = part;
      uint32_t h0, l0, h1, l1;
      split_pack2(v[it].x, v[it].y, h0, l0);
      split_pack2(v[it].z, v[it].w, h1, l1);
      const uint32_t off = (uint32_t)(t * TFM_PLANE_ROW_BYTES + c * 2);
      *reinterpret_cast<uint2*>(tsm + oGh + off) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(tsm + oGl + off) = make_uint2(l0, l1);
    }
  }
  for (int u = threadIdx.x; u < SK; u += blockDim.x) {
    km[u] = (u < S && kmask[(long long)n * T + u]) ? 1 : 0;
    lse2[u] = u < S ? lse[((long long)n * T + u) * H + h] * TFM_LOG2E : INFINITY;
  }
  if (drop_on) tfm_drop_bits(bits, SK, seed, site, (unsigned long long)blockIdx.x, thr);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
  const int w0 = 16 * warp;  // first query (phase A) / key (phase B) row of this warp
  if (w0 >= S) return;
  const uint32_t sb = smem_u32(tsm);
  const int r0 = w0 + g, r1 = r0 + 8;
  const float c_keep = drop_on ? dscale : 1.f;
  // =========================== phase A: dQ ===========================
  {
    float dq[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) dq[j][0] = dq[j][1] = dq[j][2] = dq[j][3] = 0.f;
    const float ls0 = lse2[r0], ls1 = lse2[r1], d0 = dd[r0], d1 = dd[r1];
    for (int kb = 0; kb < SK / BW; ++kb) {  // BW keys at a time
      if (BW * kb >= S) break;
      float s[BT][4], dpv[BT][4];
#pragma unroll
      for (int j = 0; j < BT; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        dpv[j][0] = dpv[j][1] = dpv[j][2] = dpv[j][3] = 0.f;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t qh[4], ql[4] = {0u, 0u, 0u, 0u}, gh[4], gl[4] = {0u, 0u, 0u, 0u};
        ldsm_x4(qh, tfm_a_addr(sb + oQh, w0, 16 * kk, lane));
        ldsm_x4(gh, tfm_a_addr(sb + oGh, w0, 16 * kk, lane));
        if (three) {
          ldsm_x4(ql, tfm_a_addr(sb + oQl, w0, 16 * kk, lane));
          ldsm_x4(gl, tfm_a_addr(sb + oGl, w0, 16 * kk, lane));
        }
#pragma unroll
        for (int j2 = 0; j2 < BT / 2; ++j2) {
          uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
          ldsm_x4(bh, tfm_bt_addr(sb + oKh, BW * kb + 16 * j2, 16 * kk, lane));
          if (three) ldsm_x4(bl, tfm_bt_addr(sb + oKl, BW * kb + 16 * j2, 16 * kk, lane));
          tfm_mma3(s[2 * j2], qh, ql, bh[0], bh[1], bl[0], bl[1], three);
          tfm_mma3(s[2 * j2 + 1], qh, ql, bh[2], bh[3], bl[2], bl[3], three);
          ldsm_x4(bh, tfm_bt_addr(sb + oVh, BW * kb + 16 * j2, 16 * kk, lane));
          if (three) ldsm_x4(bl, tfm_bt_addr(sb + oVl, BW * kb + 16 * j2, 16 * kk, lane));
          tfm_mma3(dpv[2 * j2], gh, gl, bh[0], bh[1], bl[0], bl[1], three);
          tfm_mma3(dpv[2 * j2 + 1], gh, gl, bh[2], bh[3], bl[2], bl[3], three);
        }
      }
      // dS (natural-log domain) in place of s
#pragma unroll
      for (int j = 0; j < BT; ++j) {
        const int u = BW * kb + 8 * j + 2 * tg;
        const bool v0 = km[u] != 0, v1 = km[u + 1] != 0;
        float k00 = c_keep, k01 = c_keep, k10 = c_keep, k11 = c_keep;
        if (drop_on) {
          const int jj = BT * kb + j, sh = 8 * (jj & 3) + 2 * tg;
          const uint32_t b0 = bits[r0 * (SK / 32) + (jj >> 2)] >> sh, b1 = bits[r1 * (SK / 32) + (jj >> 2)] >> sh;
          if (!(b0 & 1u)) k00 = 0.f;
          if (!(b0 & 2u)) k01 = 0.f;
          if (!(b1 & 1u)) k10 = 0.f;
          if (!(b1 & 2u)) k11 = 0.f;
        }
        const float p00 = v0 ? ex2_approx(s[j][0] - ls0) : 0.f, p01 = v1 ? ex2_approx(s[j][1] - ls0) : 0.f;
        const float p10 = v0 ? ex2_approx(s[j][2] - ls1) : 0.f, p11 = v1 ? ex2_approx(s[j][3] - ls1) : 0.f;
        s[j][0] = p00 * (k00 * dpv[j][0] - d0); s[j][1] = p01 * (k01 * dpv[j][1] - d0);
        s[j][2] = p10 * (k10 * dpv[j][2] - d1); s[j][3] = p11 * (k11 * dpv[j][3] - d1);
      }
      // dQ += dS K  (reduction over the keys of the block)
#pragma unroll
      for (int kk = 0; kk < BT / 2; ++kk) {
        uint32_t ah[4], al[4];
        split_pack2(s[2 * kk][0], s[2 * kk][1], ah[0], al[0]);
        split_pack2(s[2 * kk][2], s[2 * kk][3], ah[1], al[1]);
        split_pack2(s[2 * kk + 1][0], s[2 * kk + 1][1], ah[2], al[2]);
        split_pack2(s[2 * kk + 1][2], s[2 * kk + 1][3], ah[3], al[3]);
#pragma unroll
        for (int jd2 = 0; jd2 < 4; ++jd2) {
          uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
          ldsm_x4_t(bh, tfm_bn_addr(sb + oKh, BW * kb + 16 * kk, 16 * jd2, lane));
          if (three) ldsm_x4_t(bl, tfm_bn_addr(sb + oKl, BW * kb + 16 * kk, 16 * jd2, lane));
          tfm_mma3(dq[2 * jd2], ah, al, bh[0], bh[1], bl[0], bl[1], three);
          tfm_mma3(dq[2 * jd2 + 1], ah, al, bh[2], bh[3], bl[2], bl[3], three);
        }
      }
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = half ? r1 : r0;
      if (r >= S) continue;
      const long long rowoff = ((long long)n * T + r) * p3;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t hh, ll;
        split_pack2(dq[j][2 * half] * scale, dq[j][2 * half + 1] * scale, hh, ll);
        const long long off = rowoff + h * TFM_DH + 8 * j + 2 * tg;
        *reinterpret_cast<uint32_t*>(g_hi + off) = hh;
        if (g_lo) *reinterpret_cast<uint32_t*>(g_lo + off) = ll;
      }
      if (h == 0)
        for (int c = 3 * D + tg; c < p3; c += 4) {
          g_hi[rowoff + c] = __float2bfloat16_rn(0.f);
          if (g_lo) g_lo[rowoff + c] = __float2bfloat16_rn(0.f);
        }
    }
  }
  // =========================== phase B: dK, dV (rows = keys) ===========================
  {
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dk[j][0] = dk[j][1] = dk[j][2] = dk[j][3] = 0.f;
      dv[j][0] = dv[j][1] = dv[j][2] = dv[j][3] = 0.f;
    }
    const bool kv0 = km[r0] != 0, kv1 = km[r1] != 0;
    for (int qb = 0; qb < SK / BW; ++qb) {  // BW queries at a time
      if (BW * qb >= S) break;
      float st[BT][4], dpt[BT][4];
#pragma unroll
      for (int j = 0; j < BT; ++j) {
        st[j][0] = st[j][1] = st[j][2] = st[j][3] = 0.f;
        dpt[j][0] = dpt[j][1] = dpt[j][2] = dpt[j][3] = 0.f;
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t kh[4], kl[4] = {0u, 0u, 0u, 0u}, vh[4], vl[4] = {0u, 0u, 0u, 0u};
        ldsm_x4(kh, tfm_a_addr(sb + oKh, w0, 16 * kk, lane));
        ldsm_x4(vh, tfm_a_addr(sb + oVh, w0, 16 * kk, lane));
        if (three) {
          ldsm_x4(kl, tfm_a_addr(sb + oKl, w0, 16 * kk, lane));
          ldsm_x4(vl, tfm_a_addr(sb + oVl, w0, 16 * kk, lane));
        }
#pragma unroll
        for (int j2 = 0; j2 < BT / 2; ++j2) {
          uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
          ldsm_x4(bh, tfm_bt_addr(sb + oQh, BW * qb + 16 * j2, 16 * kk, lane));
          if (three) ldsm_x4(bl, tfm_bt_addr(sb + oQl, BW * qb + 16 * j2, 16 * kk, lane));
          tfm_mma3(st[2 * j2], kh, kl, bh[0], bh[1], bl[0], bl[1], three);
          tfm_mma3(st[2 * j2 + 1], kh, kl, bh[2], bh[3], bl[2], bl[3], three);
          ldsm_x4(bh, tfm_bt_addr(sb + oGh, BW * qb + 16 * j2, 16 * kk, lane));
          if (three) ldsm_x4(bl, tfm_bt_addr(sb + oGl, BW * qb + 16 * j2, 16 * kk, lane));
          tfm_mma3(dpt[2 * j2], vh, vl, bh[0], bh[1], bl[0], bl[1], three);
          tfm_mma3(dpt[2 * j2 + 1], vh, vl, bh[2], bh[3], bl[2], bl[3], three);
        }
      }
      // st -> dS^T, dpt -> (M c P)^T   (columns = queries t, rows = keys r0 / r1)
#pragma unroll
      for (int j = 0; j < BT; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int t = BW * qb + 8 * j + 2 * tg + e;
          const float ls = lse2[t], dt = dd[t];
          float k0 = c_keep, k1 = c_keep;
          if (drop_on) {
            const uint32_t wd = bits[t * (SK / 32) + (r0 >> 5)];
            if (!((wd >> (r0 & 31)) & 1u)) k0 = 0.f;
            if (!((wd >> (r1 & 31)) & 1u)) k1 = 0.f;
          }
          const float p0 = kv0 ? ex2_approx(st[j][e] - ls) : 0.f, p1 = kv1 ? ex2_approx(st[j][2 + e] - ls) : 0.f;
          st[j][e] = p0 * (k0 * dpt[j][e] - dt);
          st[j][2 + e] = p1 * (k1 * dpt[j][2 + e] - dt);
          dpt[j][e] = p0 * k0;
          dpt[j][2 + e] = p1 * k1;
        }
      }
#pragma unroll
      for (int kk = 0; kk < BT / 2; ++kk) {  // reduction over the queries of the block
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
        for (int jd2 = 0; jd2 < 4; ++jd2) {
          uint32_t bh[4], bl[4] = {0u, 0u, 0u, 0u};
          ldsm_x4_t(bh, tfm_bn_addr(sb + oQh, BW * qb + 16 * kk, 16 * jd2, lane));
          if (three) ldsm_x4_t(bl, tfm_bn_addr(sb + oQl, BW * qb + 16 * kk, 16 * jd2, lane));
          tfm_mma3(dk[2 * jd2], sh_, sl_, bh[0], bh[1], bl[0], bl[1], three);
          tfm_mma3(dk[2 * jd2 + 1], sh_, sl_, bh[2], bh[3], bl[2], bl[3], three);
          ldsm_x4_t(bh, tfm_bn_addr(sb + oGh, BW * qb + 16 * kk, 16 * jd2, lane));
          if (three) ldsm_x4_t(bl, tfm_bn_addr(sb + oGl, BW * qb + 16 * kk, 16 * jd2, lane));
          tfm_mma3(dv[2 * jd2], ph, pl, bh[0], bh[1], bl[0], bl[1], three);
          tfm_mma3(dv[2 * jd2 + 1], ph, pl, bh[2], bh[3], bl[2], bl[3], three);
        }
      }
    }
    // the staged Q carries log2(e) / sqrt(d_h): dK = dS^T Q / sqrt(d_h) = (dS^T Q_staged) * ln 2
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int r = half ? r1 : r0;
      if (r >= S) continue;
      const long long rowoff = ((long long)n * T + r) * p3;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint32_t hh, ll;
        const long long off = rowoff + h * TFM_DH + 8 * j + 2 * tg;
        split_pack2(dk[j][2 * half] * TFM_LN2, dk[j][2 * half + 1] * TFM_LN2, hh, ll);
        *reinterpret_cast<uint32_t*>(g_hi + off + D) = hh;
        if (g_lo) *reinterpret_cast<uint32_t*>(g_lo + off + D) = ll;
        split_pack2(dv[j][2 * half], dv[j][2 * half + 1], hh, ll);
        *reinterpret_cast<uint32_t*>(g_hi + off + 2 * D) = hh;
        if (g_lo) *reinterpret_cast<uint32_t*>(g_lo + off + 2 * D) = ll;
      }
    }
  }
}

// keep-flags [T][T] of the attention-probability dropout of one (title, head): test helper
__global__ void tfm_attn_mask_kernel(unsigned char* keep, int T, int SK, unsigned long long item, unsigned long long seed,
                                     uint32_t thr) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < T * T; i += gridDim.x * blockDim.x) {
    const int t = i / T, u = i % T;
    const unsigned long long e = (item * (unsigned long long)SK + (unsigned long long)t) * SK + (unsigned long long)u;
    keep[i] = drop_keep(seed, 2u, e, thr) ? 1 : 0;
  }
}

}  // namespace nrl
