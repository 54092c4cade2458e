// Memory-bound / small-contraction kernels of the NRMS hot path (sm_100a, fp32 SIMT):
// weight packing, embedding gather (+dropout, bf16 hi/lo split), per-head self-attention
// forward/backward (any sequence axis: title tokens, or the reference's batch-axis quirk),
// additive-attention pooling forward/backward, ragged<->dense, scorer, soft-target CE,
// embedding-gradient scatter, Adam.
#pragma once
#include <cfloat>

#include "nrl_ptx.cuh"

namespace nrl {

// Sticky device-side input-validation flag (what nn.Embedding's index check is for the reference: an id
// outside the table raises there).  Kernels that index with caller-provided ids / segment ids record the FIRST
// violation here and neutralise the access (id 0 / skipped row) instead of touching memory out of bounds; the host
// reads and clears it with nrl_device_status().
enum { DEV_ERR_TOKEN_ID = 1, DEV_ERR_SEGMENT_ID = 2, DEV_ERR_SEGMENT_LEN = 3, DEV_ERR_ROW_INDEX = 4 };
__device__ unsigned int g_dev_error = 0;
__device__ __forceinline__ void dev_error(unsigned int code) { atomicCAS(&g_dev_error, 0u, code); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum, result broadcast to every thread; `red` holds >= 33 floats
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? red[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// ------------------------------------------------------------------------------------
// Weight packing.  W [n_out, k_in] fp32 (+ bias [n_out]) ->
//   wf[plane][n_out][kp]   forward operand, K-major; column k_in holds the bias (the activation
//                          planes carry a constant 1.0 there), other pad columns 0
//   wt[plane][k_in][np]    transposed operand for the data-gradient GEMM, pad columns 0
// ------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                   int n_out, int k_in, int kp, int np, __nv_bfloat16* wf,
                                   __nv_bfloat16* wt, int two_planes) {
  const long long nf = (long long)n_out * kp, nt = (long long)k_in * np;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nf + nt;
       i += (long long)gridDim.x * blockDim.x) {
    float x;
    __nv_bfloat16* dst;
    long long plane_stride, off;
    if (i < nf) {
      int n = (int)(i / kp), k = (int)(i % kp);
      x = k < k_in ? W[(long long)n * k_in + k] : (k == k_in && bias ? bias[n] : 0.f);
      dst = wf; plane_stride = nf; off = i;
    } else {
      long long j = i - nf;
      int k = (int)(j / np), n = (int)(j % np);
      x = n < n_out ? W[(long long)n * k_in + k] : 0.f;
      dst = wt; plane_stride = nt; off = j;
    }
    __nv_bfloat16 h, l;
    split_bf16(x, h, l);
    dst[off] = h;
    if (two_planes) dst[plane_stride + off] = l;
  }
}

// All weight matrices of one or two MHSA + additive blocks in ONE launch (blockIdx.y = job): the step is a chain of
// ~45 dependent launches, and six 4-microsecond packing kernels are 1 % of it.
struct PackJob {
  const float* W; const float* bias;
  int n_out, k_in, kp, np;
  __nv_bfloat16 *wf, *wt;
};
struct PackJobs {
  PackJob j[6];
};
__global__ void pack_weights_multi_kernel(const PackJobs jobs, int two_planes) {
  const PackJob& q = jobs.j[blockIdx.y];
  const long long nf = (long long)q.n_out * q.kp, nt = (long long)q.k_in * q.np;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nf + nt;
       i += (long long)gridDim.x * blockDim.x) {
    float x;
    __nv_bfloat16* dst;
    long long plane_stride, off;
    if (i < nf) {
      int n = (int)(i / q.kp), k = (int)(i % q.kp);
      x = k < q.k_in ? q.W[(long long)n * q.k_in + k] : (k == q.k_in && q.bias ? q.bias[n] : 0.f);
      dst = q.wf; plane_stride = nf; off = i;
    } else {
      long long j = i - nf;
      int k = (int)(j / q.np), n = (int)(j % q.np);
      x = n < q.n_out ? q.W[(long long)n * q.k_in + k] : 0.f;
      dst = q.wt; plane_stride = nt; off = j;
    }
    __nv_bfloat16 h, l;
    split_bf16(x, h, l);
    dst[off] = h;
    if (two_planes) dst[plane_stride + off] = l;
  }
}

// ------------------------------------------------------------------------------------
// a2/a3: embedding gather (+ dropout site 0) -> split planes [2][R][ep], ones column at E.
// One warp per token row; rows of the table are 16-byte aligned when E % 4 == 0.
// ------------------------------------------------------------------------------------
__global__ void gather_split_kernel(const long long* __restrict__ ids, long long R,
                                    const float* __restrict__ table, long long V1, int E, int ep,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                    float* __restrict__ x_f32, const uint32_t* __restrict__ drop_words,
                                    int drop_mw, float drop_scale) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < R; r += nwarps) {
    long long id = ids ? ids[r] : r;  // ids == nullptr: dense rows (identity gather)
    if (id < 0 || id >= V1) {           // nn.Embedding raises here; row 0 is read instead and the flag is set
      if (lane == 0) dev_error(DEV_ERR_TOKEN_ID);
      id = 0;
    }
    const float* src = table + id * E;
    for (int c = lane * 4; c < ep; c += 128) {
      float v[4];
      if (c + 4 <= E && (E & 3) == 0) {
        float4 t = __ldg(reinterpret_cast<const float4*>(src + c));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = (c + i < E) ? __ldg(src + c + i) : 0.f;
      }
      if (drop_words && c < E) {  // c % 4 == 0: the four keep-bits sit in one word
        const uint32_t bits = __ldg(drop_words + r * drop_mw + (c >> 5)) >> (c & 31);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = ((bits >> i) & 1u) ? v[i] * drop_scale : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (c + i >= E) v[i] = (c + i == E) ? 1.f : 0.f;
      if (x_f32) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (c + i < E) x_f32[r * E + c + i] = v[i];
      }
      __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_bf16(v[i], h[i], l[i]);
      const long long off = r * ep + c;  // ep % 8 == 0 and c % 4 == 0 -> 8-byte aligned
      if (hi)
        *reinterpret_cast<uint2*>(hi + off) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
      if (lo)
        *reinterpret_cast<uint2*>(lo + off) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
    }
  }
}

// ------------------------------------------------------------------------------------
// a4 core: per-(batch item, head) softmax(q k^T) v over an arbitrary sequence axis.
//   row(seq s, batch b) = s * seq_stride + b * batch_stride        (rows of QKV [R, 3E] fp32)
//   title encoder : S = L,  seq_stride = 1,    NB = #news, batch_stride = L
//   user encoder  : S = B,  seq_stride = Hmax, NB = Hmax,  batch_stride = 1   (reference quirk)
// One warp per (b, head, 32-query chunk); keys streamed through per-warp smem in chunks of 32
// with an online softmax.  Writes O as split planes (+ ones column / zero pad) and the row
// log-sum-exp for the backward pass.
// ------------------------------------------------------------------------------------
// warps per CTA of the streaming attention kernels (static smem = 2 * NW * 32 * DH floats <= 48 KB)
__host__ __device__ constexpr int attn_stream_warps(int dh) { return dh <= 32 ? 4 : 2; }

template <int DH>
__global__ void __launch_bounds__(32 * attn_stream_warps(DH))
attn_fwd_kernel(const float* __restrict__ qkv, int E, int ldq, int heads, int S, long long seq_stride,
                int NB, long long batch_stride, float scale, __nv_bfloat16* __restrict__ o_hi,
                __nv_bfloat16* __restrict__ o_lo, int ep, float* __restrict__ lse) {
  __shared__ __align__(16) float sk[attn_stream_warps(DH)][32 * DH];
  __shared__ __align__(16) float sv[attn_stream_warps(DH)][32 * DH];
  constexpr int NW = attn_stream_warps(DH);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qchunks = (S + 31) / 32;
  const long long items = (long long)NB * heads * qchunks;
  const int ld = ldq;
  for (long long it = blockIdx.x * (long long)NW + warp; it < items; it += gridDim.x * (long long)NW) {
    const int qc = (int)(it % qchunks);
    const int h = (int)((it / qchunks) % heads);
    const int b = (int)(it / ((long long)qchunks * heads));
    const int t = qc * 32 + lane;
    const bool q_ok = t < S;
    const long long qrow = (long long)(q_ok ? t : 0) * seq_stride + (long long)b * batch_stride;
    float q[DH], o[DH];
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      float4 t4 = __ldg(reinterpret_cast<const float4*>(qkv + qrow * ld + h * DH + d));
      q[d] = t4.x * scale; q[d + 1] = t4.y * scale; q[d + 2] = t4.z * scale; q[d + 3] = t4.w * scale;
      o[d] = o[d + 1] = o[d + 2] = o[d + 3] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < S; k0 += 32) {
      const int nk = min(32, S - k0);
      __syncwarp();
      for (int i = lane; i < nk * (DH / 4); i += 32) {
        const int u = i / (DH / 4), d4 = i % (DH / 4);
        const long long krow = (long long)(k0 + u) * seq_stride + (long long)b * batch_stride;
        reinterpret_cast<float4*>(sk[warp])[u * (DH / 4) + d4] =
            __ldg(reinterpret_cast<const float4*>(qkv + krow * ld + E + h * DH) + d4);
        reinterpret_cast<float4*>(sv[warp])[u * (DH / 4) + d4] =
            __ldg(reinterpret_cast<const float4*>(qkv + krow * ld + 2 * E + h * DH) + d4);
      }
      __syncwarp();
      float s[32];
      float cmax = -INFINITY;
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        float acc = 0.f;
        if (u < nk) {
#pragma unroll
          for (int d = 0; d < DH; d += 4) {
            float4 k4 = reinterpret_cast<const float4*>(sk[warp])[u * (DH / 4) + d / 4];
            acc += q[d] * k4.x + q[d + 1] * k4.y + q[d + 2] * k4.z + q[d + 3] * k4.w;
          }
          cmax = fmaxf(cmax, acc);
        }
        s[u] = acc;
      }
      const float m_new = fmaxf(m, cmax);
      const float corr = __expf(m - m_new);  // m = -inf on the first chunk -> 0
      l *= corr;
#pragma unroll
      for (int d = 0; d < DH; ++d) o[d] *= corr;
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        if (u < nk) {
          const float pu = expf(s[u] - m_new);
          l += pu;
#pragma unroll
          for (int d = 0; d < DH; d += 4) {
            float4 v4 = reinterpret_cast<const float4*>(sv[warp])[u * (DH / 4) + d / 4];
            o[d] += pu * v4.x; o[d + 1] += pu * v4.y; o[d + 2] += pu * v4.z; o[d + 3] += pu * v4.w;
          }
        }
      }
      m = m_new;
    }
    if (q_ok) {
      const float inv = 1.f / l;
      lse[qrow * heads + h] = m + logf(l);
      const long long off = qrow * ep + h * DH;  // DH % 4 == 0 -> 8-byte aligned
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        __nv_bfloat16 hh[4], ll[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_bf16(o[d + i] * inv, hh[i], ll[i]);
        *reinterpret_cast<uint2*>(o_hi + off + d) = make_uint2(pack_bf16x2(hh[0], hh[1]), pack_bf16x2(hh[2], hh[3]));
        if (o_lo)
          *reinterpret_cast<uint2*>(o_lo + off + d) = make_uint2(pack_bf16x2(ll[0], ll[1]), pack_bf16x2(ll[2], ll[3]));
      }
      if (h == 0) {  // pad columns: ones column at E, zeros after
        for (int c = E; c < ep; ++c) {
          o_hi[qrow * ep + c] = __float2bfloat16_rn(c == E ? 1.f : 0.f);
          if (o_lo) o_lo[qrow * ep + c] = __float2bfloat16_rn(0.f);
        }
      }
    }
  }
}

// Backward of the attention core.  One warp per (b, head): phase A (lane = query) gives dQ,
// phase B (lane = key) gives dK / dV; P is recomputed from Q, K and the saved log-sum-exp,
// D = rowsum(dO * O) from the saved O planes.  Writes dQKV as split planes [2][R][p3].
template <int DH>
__global__ void __launch_bounds__(32 * attn_stream_warps(DH))
attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ d_o, long long ld_do,
                const __nv_bfloat16* __restrict__ o_hi, const __nv_bfloat16* __restrict__ o_lo,
                int ep, const float* __restrict__ lse, int E, int ldq, int heads, int S,
                long long seq_stride, int NB, long long batch_stride, float scale,
                __nv_bfloat16* __restrict__ g_hi, __nv_bfloat16* __restrict__ g_lo, int p3) {
  // per-warp staging: "x" rows (k/v in phase A, q/dO in phase B), plus lse / D per row
  __shared__ __align__(16) float sa[attn_stream_warps(DH)][32 * DH];
  __shared__ __align__(16) float sb[attn_stream_warps(DH)][32 * DH];
  __shared__ float s_lse[attn_stream_warps(DH)][32];
  __shared__ float s_dd[attn_stream_warps(DH)][32];
  constexpr int NW = attn_stream_warps(DH);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long items = (long long)NB * heads;
  const int ld = ldq;
  const int chunks = (S + 31) / 32;
  auto store_split = [&](long long row, int col, const float* v) {
    const long long off = row * p3 + col;
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      __nv_bfloat16 hh[4], ll[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_bf16(v[d + i], hh[i], ll[i]);
      *reinterpret_cast<uint2*>(g_hi + off + d) = make_uint2(pack_bf16x2(hh[0], hh[1]), pack_bf16x2(hh[2], hh[3]));
      if (g_lo)
        *reinterpret_cast<uint2*>(g_lo + off + d) = make_uint2(pack_bf16x2(ll[0], ll[1]), pack_bf16x2(ll[2], ll[3]));
    }
  };
  for (long long it = blockIdx.x * (long long)NW + warp; it < items; it += gridDim.x * (long long)NW) {
    const int h = (int)(it % heads);
    const int b = (int)(it / heads);
    // ---------------- phase A: lane = query, loop over key chunks -> dQ ----------------
    for (int qc = 0; qc < chunks; ++qc) {
      const int t = qc * 32 + lane;
      const bool ok = t < S;
      const long long row = (long long)(ok ? t : 0) * seq_stride + (long long)b * batch_stride;
      float q[DH], go[DH], dq[DH];
      float dd = 0.f;
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        float4 t4 = __ldg(reinterpret_cast<const float4*>(qkv + row * ld + h * DH + d));
        q[d] = t4.x * scale; q[d + 1] = t4.y * scale; q[d + 2] = t4.z * scale; q[d + 3] = t4.w * scale;
        float4 g4 = __ldg(reinterpret_cast<const float4*>(d_o + row * ld_do + h * DH + d));
        go[d] = g4.x; go[d + 1] = g4.y; go[d + 2] = g4.z; go[d + 3] = g4.w;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float ov = __bfloat162float(o_hi[row * ep + h * DH + d + i]);
          if (o_lo) ov += __bfloat162float(o_lo[row * ep + h * DH + d + i]);
          dd += go[d + i] * ov;
          dq[d + i] = 0.f;
        }
      }
      const float my_lse = lse[row * heads + h];
      for (int k0 = 0; k0 < S; k0 += 32) {
        const int nk = min(32, S - k0);
        __syncwarp();
        for (int i = lane; i < nk * (DH / 4); i += 32) {
          const int u = i / (DH / 4), d4 = i % (DH / 4);
          const long long krow = (long long)(k0 + u) * seq_stride + (long long)b * batch_stride;
          reinterpret_cast<float4*>(sa[warp])[u * (DH / 4) + d4] =
              __ldg(reinterpret_cast<const float4*>(qkv + krow * ld + E + h * DH) + d4);
          reinterpret_cast<float4*>(sb[warp])[u * (DH / 4) + d4] =
              __ldg(reinterpret_cast<const float4*>(qkv + krow * ld + 2 * E + h * DH) + d4);
        }
        __syncwarp();
        for (int u = 0; u < nk; ++u) {
          float sc = 0.f, dp = 0.f;
#pragma unroll
          for (int d = 0; d < DH; d += 4) {
            float4 k4 = reinterpret_cast<const float4*>(sa[warp])[u * (DH / 4) + d / 4];
            float4 v4 = reinterpret_cast<const float4*>(sb[warp])[u * (DH / 4) + d / 4];
            sc += q[d] * k4.x + q[d + 1] * k4.y + q[d + 2] * k4.z + q[d + 3] * k4.w;
            dp += go[d] * v4.x + go[d + 1] * v4.y + go[d + 2] * v4.z + go[d + 3] * v4.w;
          }
          const float ds = expf(sc - my_lse) * (dp - dd);
#pragma unroll
          for (int d = 0; d < DH; d += 4) {
            float4 k4 = reinterpret_cast<const float4*>(sa[warp])[u * (DH / 4) + d / 4];
            dq[d] += ds * k4.x; dq[d + 1] += ds * k4.y; dq[d + 2] += ds * k4.z; dq[d + 3] += ds * k4.w;
          }
        }
      }
      if (ok) {
#pragma unroll
        for (int d = 0; d < DH; ++d) dq[d] *= scale;
        store_split(row, h * DH, dq);
        if (h == 0) {
          for (int c = 3 * E; c < p3; ++c) {
            g_hi[row * p3 + c] = __float2bfloat16_rn(0.f);
            if (g_lo) g_lo[row * p3 + c] = __float2bfloat16_rn(0.f);
          }
        }
      }
    }
    // ---------------- phase B: lane = key, loop over query chunks -> dK, dV ----------------
    for (int kc = 0; kc < chunks; ++kc) {
      const int u = kc * 32 + lane;
      const bool ok = u < S;
      const long long row = (long long)(ok ? u : 0) * seq_stride + (long long)b * batch_stride;
      float kk[DH], vv[DH], dk[DH], dv[DH];
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        float4 k4 = __ldg(reinterpret_cast<const float4*>(qkv + row * ld + E + h * DH + d));
        float4 v4 = __ldg(reinterpret_cast<const float4*>(qkv + row * ld + 2 * E + h * DH + d));
        kk[d] = k4.x; kk[d + 1] = k4.y; kk[d + 2] = k4.z; kk[d + 3] = k4.w;
        vv[d] = v4.x; vv[d + 1] = v4.y; vv[d + 2] = v4.z; vv[d + 3] = v4.w;
        dk[d] = dk[d + 1] = dk[d + 2] = dk[d + 3] = 0.f;
        dv[d] = dv[d + 1] = dv[d + 2] = dv[d + 3] = 0.f;
      }
      for (int q0 = 0; q0 < S; q0 += 32) {
        const int nq = min(32, S - q0);
        __syncwarp();
        for (int i = lane; i < nq * (DH / 4); i += 32) {
          const int t = i / (DH / 4), d4 = i % (DH / 4);
          const long long qrow = (long long)(q0 + t) * seq_stride + (long long)b * batch_stride;
          float4 q4 = __ldg(reinterpret_cast<const float4*>(qkv + qrow * ld + h * DH) + d4);
          q4.x *= scale; q4.y *= scale; q4.z *= scale; q4.w *= scale;
          reinterpret_cast<float4*>(sa[warp])[t * (DH / 4) + d4] = q4;
          reinterpret_cast<float4*>(sb[warp])[t * (DH / 4) + d4] =
              __ldg(reinterpret_cast<const float4*>(d_o + qrow * ld_do + h * DH) + d4);
        }
        if (lane < nq) {
          const long long qrow = (long long)(q0 + lane) * seq_stride + (long long)b * batch_stride;
          float dd = 0.f;
          for (int d = 0; d < DH; ++d) {
            float ov = __bfloat162float(o_hi[qrow * ep + h * DH + d]);
            if (o_lo) ov += __bfloat162float(o_lo[qrow * ep + h * DH + d]);
            dd += d_o[qrow * ld_do + h * DH + d] * ov;
          }
          s_dd[warp][lane] = dd;
          s_lse[warp][lane] = lse[qrow * heads + h];
        }
        __syncwarp();
        for (int t = 0; t < nq; ++t) {
          float sc = 0.f, dp = 0.f;
#pragma unroll
          for (int d = 0; d < DH; d += 4) {
            float4 q4 = reinterpret_cast<const float4*>(sa[warp])[t * (DH / 4) + d / 4];
            float4 g4 = reinterpret_cast<const float4*>(sb[warp])[t * (DH / 4) + d / 4];
            sc += q4.x * kk[d] + q4.y * kk[d + 1] + q4.z * kk[d + 2] + q4.w * kk[d + 3];
            dp += g4.x * vv[d] + g4.y * vv[d + 1] + g4.z * vv[d + 2] + g4.w * vv[d + 3];
          }
          const float pr = expf(sc - s_lse[warp][t]);
          const float ds = pr * (dp - s_dd[warp][t]);
#pragma unroll
          for (int d = 0; d < DH; d += 4) {
            float4 q4 = reinterpret_cast<const float4*>(sa[warp])[t * (DH / 4) + d / 4];
            float4 g4 = reinterpret_cast<const float4*>(sb[warp])[t * (DH / 4) + d / 4];
            dk[d] += ds * q4.x; dk[d + 1] += ds * q4.y; dk[d + 2] += ds * q4.z; dk[d + 3] += ds * q4.w;
            dv[d] += pr * g4.x; dv[d + 1] += pr * g4.y; dv[d + 2] += pr * g4.z; dv[d + 3] += pr * g4.w;
          }
        }
      }
      if (ok) {
        store_split(row, E + h * DH, dk);
        store_split(row, 2 * E + h * DH, dv);
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// Tile-resident attention (the fast path).  One CTA per batch item b and head group: the
// Q|K|V slices of `hp` heads for all S sequence rows are staged in shared memory with fully
// coalesced row-segment loads (cp.async-free: 16-byte __ldg -> st.shared), every warp then
// owns (head, 32-query chunk) items with lane = query and walks the keys in smem (broadcast
// reads), and the result tile is written back with coalesced row segments.  Chosen whenever
// one head of the sequence fits the smem budget; otherwise the streaming kernels above run.
// smem row layout: [ q(hp*DH) | k(hp*DH) | v(hp*DH) ]  (+ [ dO | dQ ] in the backward kernel)
// ------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(256, DH <= 32 ? 2 : 1)
attn_fwd_tile_kernel(const float* __restrict__ qkv, int E, int ldq, int heads, int S, long long seq_stride,
                     int NB, long long batch_stride, float scale, int hp,
                     __nv_bfloat16* __restrict__ o_hi, __nv_bfloat16* __restrict__ o_lo, int ep,
                     float* __restrict__ lse) {
  extern __shared__ __align__(16) float tile[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int passes = (heads + hp - 1) / hp;
  const int ld = ldq;
  const int qchunks = (S + 31) / 32;
  for (long long work = blockIdx.x; work < (long long)NB * passes; work += gridDim.x) {
    const int b = (int)(work / passes), pass = (int)(work % passes);
    const int h0 = pass * hp, nh = min(hp, heads - h0);
    const int W = nh * DH;       // floats per q / k / v segment
    const int P = 3 * W;         // smem row pitch
    const int segv = W / 4;      // float4 per segment
    __syncthreads();
    {  // stage the tile with four independent 16-byte loads in flight per thread (the loop is
       // latency-bound otherwise: one DRAM round trip per iteration)
      const int total = S * 3 * segv;
      for (int i0 = threadIdx.x; i0 < total; i0 += 4 * blockDim.x) {
        float4 v[4];
        int dst[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * blockDim.x;
          dst[u] = -1;
          if (i < total) {
            const int srow = i / (3 * segv), rem = i % (3 * segv), seg = rem / segv, c4 = rem % segv;
            const long long grow = (long long)srow * seq_stride + (long long)b * batch_stride;
            v[u] = __ldg(reinterpret_cast<const float4*>(qkv + grow * ld + seg * E + h0 * DH) + c4);
            dst[u] = srow * P + seg * W + 4 * c4;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (dst[u] >= 0) *reinterpret_cast<float4*>(tile + dst[u]) = v[u];
      }
    }
    __syncthreads();
    for (int it = warp; it < nh * qchunks; it += nwarps) {
      const int hl = it / qchunks, qc = it % qchunks;
      const int t = qc * 32 + lane;
      const bool ok = t < S;
      const float* qrow = tile + (long long)(ok ? t : 0) * P + hl * DH;
      float q[DH], o[DH];
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        float4 t4 = *reinterpret_cast<const float4*>(qrow + d);
        q[d] = t4.x * scale; q[d + 1] = t4.y * scale; q[d + 2] = t4.z * scale; q[d + 3] = t4.w * scale;
        o[d] = o[d + 1] = o[d + 2] = o[d + 3] = 0.f;
      }
      float m = -INFINITY;
      for (int u = 0; u < S; ++u) {
        const float* kr = tile + (long long)u * P + W + hl * DH;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;  // four independent FMA chains
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          float4 k4 = *reinterpret_cast<const float4*>(kr + d);
          a0 += q[d] * k4.x; a1 += q[d + 1] * k4.y; a2 += q[d + 2] * k4.z; a3 += q[d + 3] * k4.w;
        }
        const float acc = (a0 + a1) + (a2 + a3);
        m = fmaxf(m, acc);
      }
      float l = 0.f;
      for (int u = 0; u < S; ++u) {
        const float* kr = tile + (long long)u * P + W + hl * DH;
        const float* vr = kr + W;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;  // four independent FMA chains
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          float4 k4 = *reinterpret_cast<const float4*>(kr + d);
          a0 += q[d] * k4.x; a1 += q[d + 1] * k4.y; a2 += q[d + 2] * k4.z; a3 += q[d + 3] * k4.w;
        }
        const float acc = (a0 + a1) + (a2 + a3);
        const float pu = expf(acc - m);
        l += pu;
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          float4 v4 = *reinterpret_cast<const float4*>(vr + d);
          o[d] += pu * v4.x; o[d + 1] += pu * v4.y; o[d + 2] += pu * v4.z; o[d + 3] += pu * v4.w;
        }
      }
      __syncwarp();
      if (ok) {
        const float inv = 1.f / l;
        // the q slice of (t, head) is read by this lane only: reuse it for the output row
        float* orow = tile + (long long)t * P + hl * DH;
#pragma unroll
        for (int d = 0; d < DH; ++d) orow[d] = o[d] * inv;
        const long long grow = (long long)t * seq_stride + (long long)b * batch_stride;
        lse[grow * heads + h0 + hl] = m + logf(l);
      }
    }
    __syncthreads();
    // coalesced write-back of the [S, W] output tile as split planes
    for (int i = threadIdx.x; i < S * (W / 2); i += blockDim.x) {
      const int srow = i / (W / 2), c = (i % (W / 2)) * 2;
      const long long grow = (long long)srow * seq_stride + (long long)b * batch_stride;
      const float2 val = *reinterpret_cast<const float2*>(tile + (long long)srow * P + c);
      __nv_bfloat16 ah, al, bh, bl;
      split_bf16(val.x, ah, al);
      split_bf16(val.y, bh, bl);
      const long long off = grow * ep + h0 * DH + c;
      *reinterpret_cast<uint32_t*>(o_hi + off) = pack_bf16x2(ah, bh);
      if (o_lo) *reinterpret_cast<uint32_t*>(o_lo + off) = pack_bf16x2(al, bl);
    }
    if (pass == 0) {
      for (int i = threadIdx.x; i < S * (ep - E); i += blockDim.x) {
        const int srow = i / (ep - E), c = E + i % (ep - E);
        const long long grow = (long long)srow * seq_stride + (long long)b * batch_stride;
        o_hi[grow * ep + c] = __float2bfloat16_rn(c == E ? 1.f : 0.f);
        if (o_lo) o_lo[grow * ep + c] = __float2bfloat16_rn(0.f);
      }
    }
  }
}

template <int DH>
__global__ void __launch_bounds__(256, DH <= 32 ? 2 : 1)
attn_bwd_tile_kernel(const float* __restrict__ qkv, const float* __restrict__ d_o, long long ld_do,
                     const float* __restrict__ lse, int E, int ldq, int heads, int S, long long seq_stride,
                     int NB, long long batch_stride, float scale, int hp,
                     __nv_bfloat16* __restrict__ g_hi, __nv_bfloat16* __restrict__ g_lo, int p3) {
  extern __shared__ __align__(16) float tile[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int passes = (heads + hp - 1) / hp;
  const int ld = ldq;
  const int chunks = (S + 31) / 32;
  for (long long work = blockIdx.x; work < (long long)NB * passes; work += gridDim.x) {
    const int b = (int)(work / passes), pass = (int)(work % passes);
    const int h0 = pass * hp, nh = min(hp, heads - h0);
    const int W = nh * DH;
    const int P = 5 * W;  // [q | k | v | dO | dQ]
    const int segv = W / 4;
    float* s_lse = tile + (long long)S * P;   // [S][nh]
    float* s_dd = s_lse + S * nh;             // [S][nh]
    __syncthreads();
    {  // four independent 16-byte loads in flight per thread
      const int total = S * 4 * segv;
      for (int i0 = threadIdx.x; i0 < total; i0 += 4 * blockDim.x) {
        float4 v[4];
        int dst[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * blockDim.x;
          dst[u] = -1;
          if (i < total) {
            const int srow = i / (4 * segv), rem = i % (4 * segv), seg = rem / segv, c4 = rem % segv;
            const long long grow = (long long)srow * seq_stride + (long long)b * batch_stride;
            const float4* src = seg < 3 ? reinterpret_cast<const float4*>(qkv + grow * ld + seg * E + h0 * DH)
                                        : reinterpret_cast<const float4*>(d_o + grow * ld_do + h0 * DH);
            v[u] = __ldg(src + c4);
            dst[u] = srow * P + seg * W + 4 * c4;
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (dst[u] >= 0) *reinterpret_cast<float4*>(tile + dst[u]) = v[u];
      }
    }
    for (int i = threadIdx.x; i < S * nh; i += blockDim.x) {
      const int srow = i / nh, hl = i % nh;
      const long long grow = (long long)srow * seq_stride + (long long)b * batch_stride;
      s_lse[i] = lse[grow * heads + h0 + hl];
    }
    __syncthreads();
    // ---- phase A: lane = query.  sweep 1: D_t = sum_u p dp;  sweep 2: dq_t ----
    for (int it = warp; it < nh * chunks; it += nwarps) {
      const int hl = it / chunks, qc = it % chunks;
      const int t = qc * 32 + lane;
      const bool ok = t < S;
      const float* base = tile + (long long)(ok ? t : 0) * P + hl * DH;
      float q[DH], go[DH], dq[DH];
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        float4 t4 = *reinterpret_cast<const float4*>(base + d);
        float4 g4 = *reinterpret_cast<const float4*>(base + 3 * W + d);
        q[d] = t4.x * scale; q[d + 1] = t4.y * scale; q[d + 2] = t4.z * scale; q[d + 3] = t4.w * scale;
        go[d] = g4.x; go[d + 1] = g4.y; go[d + 2] = g4.z; go[d + 3] = g4.w;
        dq[d] = dq[d + 1] = dq[d + 2] = dq[d + 3] = 0.f;
      }
      const float my_lse = s_lse[(ok ? t : 0) * nh + hl];
      float dd = 0.f;
      for (int u = 0; u < S; ++u) {
        const float* kr = tile + (long long)u * P + W + hl * DH;
        float s0 = 0.f, s1 = 0.f, p0 = 0.f, p1 = 0.f;  // independent FMA chains
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          float4 k4 = *reinterpret_cast<const float4*>(kr + d);
          float4 v4 = *reinterpret_cast<const float4*>(kr + W + d);
          s0 += q[d] * k4.x + q[d + 1] * k4.y; s1 += q[d + 2] * k4.z + q[d + 3] * k4.w;
          p0 += go[d] * v4.x + go[d + 1] * v4.y; p1 += go[d + 2] * v4.z + go[d + 3] * v4.w;
        }
        const float sc = s0 + s1, dp = p0 + p1;
        dd += expf(sc - my_lse) * dp;
      }
      for (int u = 0; u < S; ++u) {
        const float* kr = tile + (long long)u * P + W + hl * DH;
        float s0 = 0.f, s1 = 0.f, p0 = 0.f, p1 = 0.f;  // independent FMA chains
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          float4 k4 = *reinterpret_cast<const float4*>(kr + d);
          float4 v4 = *reinterpret_cast<const float4*>(kr + W + d);
          s0 += q[d] * k4.x + q[d + 1] * k4.y; s1 += q[d + 2] * k4.z + q[d + 3] * k4.w;
          p0 += go[d] * v4.x + go[d + 1] * v4.y; p1 += go[d + 2] * v4.z + go[d + 3] * v4.w;
        }
        const float sc = s0 + s1, dp = p0 + p1;
        const float ds = expf(sc - my_lse) * (dp - dd);
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          float4 k4 = *reinterpret_cast<const float4*>(kr + d);
          dq[d] += ds * k4.x; dq[d + 1] += ds * k4.y; dq[d + 2] += ds * k4.z; dq[d + 3] += ds * k4.w;
        }
      }
      if (ok) {
        float* dst = tile + (long long)t * P + 4 * W + hl * DH;
#pragma unroll
        for (int d = 0; d < DH; ++d) dst[d] = dq[d] * scale;
        s_dd[t * nh + hl] = dd;
      }
    }
    __syncthreads();
    // ---- phase B: lane = key -> dK, dV (written in place over the K / V slices) ----
    for (int it = warp; it < nh * chunks; it += nwarps) {
      const int hl = it / chunks, kc = it % chunks;
      const int u = kc * 32 + lane;
      const bool ok = u < S;
      float* krow = tile + (long long)(ok ? u : 0) * P + W + hl * DH;
      float kk[DH], vv[DH], dk[DH], dv[DH];
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        float4 k4 = *reinterpret_cast<const float4*>(krow + d);
        float4 v4 = *reinterpret_cast<const float4*>(krow + W + d);
        kk[d] = k4.x; kk[d + 1] = k4.y; kk[d + 2] = k4.z; kk[d + 3] = k4.w;
        vv[d] = v4.x; vv[d + 1] = v4.y; vv[d + 2] = v4.z; vv[d + 3] = v4.w;
        dk[d] = dk[d + 1] = dk[d + 2] = dk[d + 3] = 0.f;
        dv[d] = dv[d + 1] = dv[d + 2] = dv[d + 3] = 0.f;
      }
      for (int t = 0; t < S; ++t) {
        const float* qr = tile + (long long)t * P + hl * DH;
        float s0 = 0.f, s1 = 0.f, p0 = 0.f, p1 = 0.f;
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          float4 q4 = *reinterpret_cast<const float4*>(qr + d);
          float4 g4 = *reinterpret_cast<const float4*>(qr + 3 * W + d);
          s0 += q4.x * kk[d] + q4.y * kk[d + 1]; s1 += q4.z * kk[d + 2] + q4.w * kk[d + 3];
          p0 += g4.x * vv[d] + g4.y * vv[d + 1]; p1 += g4.z * vv[d + 2] + g4.w * vv[d + 3];
        }
        const float sc = s0 + s1, dp = p0 + p1;
        const float pr = expf(sc * scale - s_lse[t * nh + hl]);
        const float ds = pr * (dp - s_dd[t * nh + hl]) * scale;
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          float4 q4 = *reinterpret_cast<const float4*>(qr + d);
          float4 g4 = *reinterpret_cast<const float4*>(qr + 3 * W + d);
          dk[d] += ds * q4.x; dk[d + 1] += ds * q4.y; dk[d + 2] += ds * q4.z; dk[d + 3] += ds * q4.w;
          dv[d] += pr * g4.x; dv[d + 1] += pr * g4.y; dv[d + 2] += pr * g4.z; dv[d + 3] += pr * g4.w;
        }
      }
      __syncwarp();
      if (ok) {
#pragma unroll
        for (int d = 0; d < DH; ++d) { krow[d] = dk[d]; krow[W + d] = dv[d]; }
      }
    }
    __syncthreads();
    // ---- coalesced write-back: dQ | dK | dV row segments -> split planes ----
    for (int i = threadIdx.x; i < S * 3 * (W / 2); i += blockDim.x) {
      const int srow = i / (3 * (W / 2)), rem = i % (3 * (W / 2)), seg = rem / (W / 2), c = (rem % (W / 2)) * 2;
      const long long grow = (long long)srow * seq_stride + (long long)b * batch_stride;
      const int ssrc = seg == 0 ? 4 : seg;  // dQ lives in segment 4, dK / dV replaced segments 1 / 2
      const float2 val = *reinterpret_cast<const float2*>(tile + (long long)srow * P + ssrc * W + c);
      __nv_bfloat16 ah, al, bh, bl;
      split_bf16(val.x, ah, al);
      split_bf16(val.y, bh, bl);
      const long long off = grow * p3 + seg * E + h0 * DH + c;
      *reinterpret_cast<uint32_t*>(g_hi + off) = pack_bf16x2(ah, bh);
      if (g_lo) *reinterpret_cast<uint32_t*>(g_lo + off) = pack_bf16x2(al, bl);
    }
    if (pass == 0 && p3 > 3 * E) {
      for (int i = threadIdx.x; i < S * (p3 - 3 * E); i += blockDim.x) {
        const int srow = i / (p3 - 3 * E), c = 3 * E + i % (p3 - 3 * E);
        const long long grow = (long long)srow * seq_stride + (long long)b * batch_stride;
        g_hi[grow * p3 + c] = __float2bfloat16_rn(0.f);
        if (g_lo) g_lo[grow * p3 + c] = __float2bfloat16_rn(0.f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// Register-resident attention for short sequences (S <= 32: the 30-token titles).  One CTA =
// one batch item x one group of `hg` heads, one warp per head, lane = query (phase A) / key
// (phase B).  The whole score row of a query lives in registers (fully unrolled, 32
// independent FMA chains), K / V rows are broadcast reads from the smem tile, whose row pitch
// is an odd number of 16-byte words so that row-strided float4 accesses are conflict free.
// ------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int attn_pitch(int floats) {  // floats % 4 == 0
  while (((floats >> 2) & 1) == 0) floats += 4;
  return floats;
}

// packed fp32 pairs (Blackwell FFMA2: two FMAs per issued instruction)
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
constexpr float NRL_LOG2E = 1.4426950408889634f;
constexpr float NRL_LN2 = 0.6931471805599453f;

// Cooperative tile copy helper: thread -> (row sub-index, 16-byte column) computed once per work
// item; rows are then walked with a fixed stride (no div/mod in the loop).
struct TileMap {
  int r0, rstep, col;  // col in float4 units within the row's [nseg * segv] vector; r0 < 0: idle
};
__device__ __forceinline__ TileMap tile_map(int per_row) {
  TileMap m;
  const int rpp = blockDim.x / per_row;
  if (rpp >= 1) {
    m.rstep = rpp;
    m.r0 = (int)threadIdx.x < rpp * per_row ? (int)threadIdx.x / per_row : -1;
    m.col = (int)threadIdx.x % per_row;
  } else {
    m.rstep = 0; m.r0 = 0; m.col = 0;
  }
  return m;
}

template <int DH>
__global__ void __launch_bounds__(160, 4)
attn_fwd_s32_kernel(const float* __restrict__ qkv, int E, int ldq, int heads, int S, long long seq_stride,
                    int NB, long long batch_stride, float scale, __nv_bfloat16* __restrict__ o_hi,
                    __nv_bfloat16* __restrict__ o_lo, int ep, float* __restrict__ lse) {
  extern __shared__ __align__(16) float tile[];  // [S][P]: q | k | v of the head group
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hg = blockDim.x >> 5;
  const int groups = (heads + hg - 1) / hg;
  const int ld = ldq;
  const float qscale = scale * NRL_LOG2E;  // scores live in the log2 domain: p = 2^(s - m)
  for (long long work = blockIdx.x; work < (long long)NB * groups; work += gridDim.x) {
    const int b = (int)(work / groups), grp = (int)(work % groups);
    const int h0 = grp * hg, nh = min(hg, heads - h0);
    const int W = nh * DH, P = attn_pitch(3 * W), segv = W / 4;
    __syncthreads();
    {
      const TileMap tm = tile_map(3 * segv);
      if (tm.rstep > 0) {
        if (tm.r0 >= 0) {
          const int seg = tm.col / segv, c4 = tm.col - seg * segv;
          const float* src0 = qkv + seg * E + h0 * DH + 4 * c4;
          float* dst0 = tile + seg * W + 4 * c4;
          for (int srow = tm.r0; srow < S; srow += tm.rstep) {
            const long long grow = (long long)srow * seq_stride + (long long)b * batch_stride;
            *reinterpret_cast<float4*>(dst0 + srow * P) = __ldg(reinterpret_cast<const float4*>(src0 + grow * ld));
          }
        }
      } else {
        for (int i = threadIdx.x; i < S * 3 * segv; i += blockDim.x) {
          const int srow = i / (3 * segv), rem = i - srow * 3 * segv, seg = rem / segv, c4 = rem - seg * segv;
          const long long grow = (long long)srow * seq_stride + (long long)b * batch_stride;
          reinterpret_cast<float4*>(tile + srow * P + seg * W)[c4] =
              __ldg(reinterpret_cast<const float4*>(qkv + grow * ld + seg * E + h0 * DH) + c4);
        }
      }
    }
    __syncthreads();
    if (warp < nh) {
      const bool ok = lane < S;
      float* qrow = tile + (ok ? lane : 0) * P + warp * DH;
      float2 q2[DH / 2];
#pragma unroll
      for (int d = 0; d < DH; d += 4) {
        const float4 t4 = *reinterpret_cast<const float4*>(qrow + d);
        q2[d / 2] = make_float2(t4.x * qscale, t4.y * qscale);
        q2[d / 2 + 1] = make_float2(t4.z * qscale, t4.w * qscale);
      }
      float sc[32];
      float m = -INFINITY;
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        sc[u] = -INFINITY;
        if (u < S) {
          const float* kr = tile + u * P + W + warp * DH;
          float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
          for (int d = 0; d < DH; d += 4) {
            const float4 k4 = *reinterpret_cast<const float4*>(kr + d);
            a0 = ffma2(q2[d / 2], make_float2(k4.x, k4.y), a0);
            a1 = ffma2(q2[d / 2 + 1], make_float2(k4.z, k4.w), a1);
          }
          sc[u] = (a0.x + a0.y) + (a1.x + a1.y);
          m = fmaxf(m, sc[u]);
        }
      }
      float l = 0.f;
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        const float pu = (u < S) ? ex2_approx(sc[u] - m) : 0.f;
        sc[u] = pu;
        l += pu;
      }
      float2 o2[DH / 2];
#pragma unroll
      for (int d = 0; d < DH / 2; ++d) o2[d] = make_float2(0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 32; ++u) {
        if (u < S) {
          const float* vr = tile + u * P + 2 * W + warp * DH;
          const float2 p2 = make_float2(sc[u], sc[u]);
#pragma unroll
          for (int d = 0; d < DH; d += 4) {
            const float4 v4 = *reinterpret_cast<const float4*>(vr + d);
            o2[d / 2] = ffma2(p2, make_float2(v4.x, v4.y), o2[d / 2]);
            o2[d / 2 + 1] = ffma2(p2, make_float2(v4.z, v4.w), o2[d / 2 + 1]);
          }
        }
      }
      if (ok) {  // the q slice of (row, head) is read by this lane only: it becomes the output row
        const float inv = 1.f / l;
#pragma unroll
        for (int d = 0; d < DH; d += 4)
          *reinterpret_cast<float4*>(qrow + d) =
              make_float4(o2[d / 2].x * inv, o2[d / 2].y * inv, o2[d / 2 + 1].x * inv, o2[d / 2 + 1].y * inv);
        const long long grow = (long long)lane * seq_stride + (long long)b * batch_stride;
        lse[grow * heads + h0 + warp] = m * NRL_LN2 + logf(l);  // natural-log LSE of the scaled scores
      }
    }
    __syncthreads();
    // write-back of the [S, W] output tile as split planes, 4 columns (8 bytes) per thread
    {
      const TileMap tm = tile_map(segv);
      const int rstep = tm.rstep > 0 ? tm.rstep : 1;
      if (tm.rstep > 0 ? tm.r0 >= 0 : true) {
        for (int idx = (tm.rstep > 0 ? tm.r0 : (int)threadIdx.x); idx < (tm.rstep > 0 ? S : S * segv);
             idx += (tm.rstep > 0 ? rstep : (int)blockDim.x)) {
          const int srow = tm.rstep > 0 ? idx : idx / segv;
          const int c = (tm.rstep > 0 ? tm.col : idx - srow * segv) * 4;
          const long long grow = (long long)srow * seq_stride + (long long)b * batch_stride;
          const float4 val = *reinterpret_cast<const float4*>(tile + srow * P + c);
          __nv_bfloat16 h[4], l4[4];
          split_bf16(val.x, h[0], l4[0]); split_bf16(val.y, h[1], l4[1]);
          split_bf16(val.z, h[2], l4[2]); split_bf16(val.w, h[3], l4[3]);
          const long long off = grow * ep + h0 * DH + c;
          *reinterpret_cast<uint2*>(o_hi + off) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
          if (o_lo) *reinterpret_cast<uint2*>(o_lo + off) = make_uint2(pack_bf16x2(l4[0], l4[1]), pack_bf16x2(l4[2], l4[3]));
        }
      }
    }
    if (grp == 0) {
      for (int i = threadIdx.x; i < S * (ep - E); i += blockDim.x) {
        const int srow = i / (ep - E), c = E + i % (ep - E);
        const long long grow = (long long)srow * seq_stride + (long long)b * batch_stride;
        o_hi[grow * ep + c] = __float2bfloat16_rn(c == E ? 1.f : 0.f);
        if (o_lo) o_lo[grow * ep + c] = __float2bfloat16_rn(0.f);
      }
    }
  }
}

// Backward, S <= 32.  Phase A (lane = query t): p[u], dS[u] for all keys in registers, dq.
// dS and P are then transposed through a per-warp [32][33] smem buffer so that phase B
// (lane = key u) gets dK = dS^T (scale q) and dV = P^T dO without recomputing any product.
// smem row: q | k | v | dO of the head group; dq / dk / dv replace q / k / v before write-back.
template <int DH>
__global__ void __launch_bounds__(160, 3)
attn_bwd_s32_kernel(const float* __restrict__ qkv, const float* __restrict__ d_o, long long ld_do,
                    const float* __restrict__ lse, int E, int ldq, int heads, int S, long long seq_stride,
                    int NB, long long batch_stride, float scale, __nv_bfloat16* __restrict__ g_hi,
                    __nv_bfloat16* __restrict__ g_lo, int p3) {
  extern __shared__ __align__(16) float tile[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hg = blockDim.x >> 5;
  const int groups = (heads + hg - 1) / hg;
  const int ld = ldq;
  const int Pmax = attn_pitch(4 * hg * DH);
  const float qscale = scale * NRL_LOG2E;
  float* tbuf = tile + S * Pmax + warp * (32 * 33);
  for (long long work = blockIdx.x; work < (long long)NB * groups; work += gridDim.x) {
    const int b = (int)(work / groups), grp = (int)(work % groups);
    const int h0 = grp * hg, nh = min(hg, heads - h0);
    const int W = nh * DH, P = attn_pitch(4 * W), segv = W / 4;
    __syncthreads();
    {
      const TileMap tm = tile_map(4 * segv);
      if (tm.rstep > 0) {
        if (tm.r0 >= 0) {
          const int seg = tm.col / segv, c4 = tm.col - seg * segv;
          const float* src0 = seg < 3 ? qkv + seg * E + h0 * DH + 4 * c4 : d_o + h0 * DH + 4 * c4;
          const long long sld = seg < 3 ? (long long)ld : ld_do;
          float* dst0 = tile + seg * W + 4 * c4;
          for (int srow = tm.r0; srow < S; srow += tm.rstep) {
            const long long grow = (long long)srow * seq_stride + (long long)b * batch_stride;
            *reinterpret_cast<float4*>(dst0 + srow * P) = __ldg(reinterpret_cast<const float4*>(src0 + grow * sld));
          }
        }
      } else {
        for (int i = threadIdx.x; i < S * 4 * segv; i += blockDim.x) {
          const int srow = i / (4 * segv), rem = i - srow * 4 * segv, seg = rem / segv, c4 = rem - seg * segv;
          const long long grow = (long long)srow * seq_stride + (long long)b * batch_stride;
          const float4* src = seg < 3 ? reinterpret_cast<const float4*>(qkv + grow * ld + seg * E + h0 * DH)
                                      : reinterpret_cast<const float4*>(d_o + grow * ld_do + h0 * DH);
          reinterpret_cast<float4*>(tile + srow * P + seg * W)[c4] = __ldg(src + c4);
        }
      }
    }
    __syncthreads();
    if (warp < nh) {
      const bool ok = lane < S;
      float* base = tile + (ok ? lane : 0) * P + warp * DH;
      const long long grow_l = (long long)(ok ? lane : 0) * seq_stride + (long long)b * batch_stride;
      const float my_lse2 = lse[grow_l * heads + h0 + warp] * NRL_LOG2E;
      float pr[32], ds[32];
      float2 dq2[DH / 2];
      {
        float2 q2[DH / 2], go2[DH / 2];
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          const float4 t4 = *reinterpret_cast<const float4*>(base + d);
          const float4 g4 = *reinterpret_cast<const float4*>(base + 3 * W + d);
          q2[d / 2] = make_float2(t4.x * qscale, t4.y * qscale);
          q2[d / 2 + 1] = make_float2(t4.z * qscale, t4.w * qscale);
          go2[d / 2] = make_float2(g4.x, g4.y);
          go2[d / 2 + 1] = make_float2(g4.z, g4.w);
          dq2[d / 2] = make_float2(0.f, 0.f);
          dq2[d / 2 + 1] = make_float2(0.f, 0.f);
        }
        float dd = 0.f;
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          pr[u] = 0.f; ds[u] = 0.f;
          if (u < S) {
            const float* kr = tile + u * P + W + warp * DH;
            float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
            float2 p0 = make_float2(0.f, 0.f), p1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int d = 0; d < DH; d += 4) {
              const float4 k4 = *reinterpret_cast<const float4*>(kr + d);
              const float4 v4 = *reinterpret_cast<const float4*>(kr + W + d);
              s0 = ffma2(q2[d / 2], make_float2(k4.x, k4.y), s0);
              s1 = ffma2(q2[d / 2 + 1], make_float2(k4.z, k4.w), s1);
              p0 = ffma2(go2[d / 2], make_float2(v4.x, v4.y), p0);
              p1 = ffma2(go2[d / 2 + 1], make_float2(v4.z, v4.w), p1);
            }
            const float pu = ok ? ex2_approx((s0.x + s0.y) + (s1.x + s1.y) - my_lse2) : 0.f;
            pr[u] = pu;
            ds[u] = (p0.x + p0.y) + (p1.x + p1.y);
            dd += pu * ds[u];
          }
        }
#pragma unroll
        for (int u = 0; u < 32; ++u) {
          ds[u] = pr[u] * (ds[u] - dd);
          if (u < S) {
            const float* kr = tile + u * P + W + warp * DH;
            const float2 d2 = make_float2(ds[u], ds[u]);
#pragma unroll
            for (int d = 0; d < DH; d += 4) {
              const float4 k4 = *reinterpret_cast<const float4*>(kr + d);
              dq2[d / 2] = ffma2(d2, make_float2(k4.x, k4.y), dq2[d / 2]);
              dq2[d / 2 + 1] = ffma2(d2, make_float2(k4.z, k4.w), dq2[d / 2 + 1]);
            }
          }
        }
      }
      // ---- transpose dS, then P: lane t holds row t, wants column `lane` ----
#pragma unroll
      for (int u = 0; u < 32; ++u) tbuf[lane * 33 + u] = ds[u];
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 32; ++t) ds[t] = tbuf[t * 33 + lane];
      __syncwarp();
#pragma unroll
      for (int u = 0; u < 32; ++u) tbuf[lane * 33 + u] = pr[u];
      __syncwarp();
#pragma unroll
      for (int t = 0; t < 32; ++t) pr[t] = tbuf[t * 33 + lane];
      // ---- phase B: lane = key.  dk = scale * sum_t dS[t][u] q_t ; dv = sum_t P[t][u] dO_t ----
      float2 dk2[DH / 2], dv2[DH / 2];
#pragma unroll
      for (int d = 0; d < DH / 2; ++d) { dk2[d] = make_float2(0.f, 0.f); dv2[d] = make_float2(0.f, 0.f); }
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        if (t < S) {
          const float* qr = tile + t * P + warp * DH;
          const float2 d2 = make_float2(ds[t], ds[t]), p2 = make_float2(pr[t], pr[t]);
#pragma unroll
          for (int d = 0; d < DH; d += 4) {
            const float4 q4 = *reinterpret_cast<const float4*>(qr + d);
            const float4 g4 = *reinterpret_cast<const float4*>(qr + 3 * W + d);
            dk2[d / 2] = ffma2(d2, make_float2(q4.x, q4.y), dk2[d / 2]);
            dk2[d / 2 + 1] = ffma2(d2, make_float2(q4.z, q4.w), dk2[d / 2 + 1]);
            dv2[d / 2] = ffma2(p2, make_float2(g4.x, g4.y), dv2[d / 2]);
            dv2[d / 2 + 1] = ffma2(p2, make_float2(g4.z, g4.w), dv2[d / 2 + 1]);
          }
        }
      }
      __syncwarp();  // every lane is done reading this head's q / k / v / dO slices
      if (ok) {
#pragma unroll
        for (int d = 0; d < DH; d += 4) {
          *reinterpret_cast<float4*>(base + d) = make_float4(dq2[d / 2].x * scale, dq2[d / 2].y * scale,
                                                             dq2[d / 2 + 1].x * scale, dq2[d / 2 + 1].y * scale);
          *reinterpret_cast<float4*>(base + W + d) = make_float4(dk2[d / 2].x * scale, dk2[d / 2].y * scale,
                                                                 dk2[d / 2 + 1].x * scale, dk2[d / 2 + 1].y * scale);
          *reinterpret_cast<float4*>(base + 2 * W + d) =
              make_float4(dv2[d / 2].x, dv2[d / 2].y, dv2[d / 2 + 1].x, dv2[d / 2 + 1].y);
        }
      }
    }
    __syncthreads();
    // ---- write-back: dQ | dK | dV row segments -> split planes, 4 columns per thread ----
    {
      const TileMap tm = tile_map(3 * segv);
      auto put = [&](int srow, int seg, int c) {
        const long long grow = (long long)srow * seq_stride + (long long)b * batch_stride;
        const float4 val = *reinterpret_cast<const float4*>(tile + srow * P + seg * W + c);
        __nv_bfloat16 h[4], l4[4];
        split_bf16(val.x, h[0], l4[0]); split_bf16(val.y, h[1], l4[1]);
        split_bf16(val.z, h[2], l4[2]); split_bf16(val.w, h[3], l4[3]);
        const long long off = grow * p3 + seg * E + h0 * DH + c;
        *reinterpret_cast<uint2*>(g_hi + off) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
        if (g_lo) *reinterpret_cast<uint2*>(g_lo + off) = make_uint2(pack_bf16x2(l4[0], l4[1]), pack_bf16x2(l4[2], l4[3]));
      };
      if (tm.rstep > 0) {
        if (tm.r0 >= 0) {
          const int seg = tm.col / segv, c = (tm.col - seg * segv) * 4;
          for (int srow = tm.r0; srow < S; srow += tm.rstep) put(srow, seg, c);
        }
      } else {
        for (int i = threadIdx.x; i < S * 3 * segv; i += blockDim.x) {
          const int srow = i / (3 * segv), rem = i - srow * 3 * segv, seg = rem / segv;
          put(srow, seg, (rem - seg * segv) * 4);
        }
      }
    }
    if (grp == 0 && p3 > 3 * E) {
      for (int i = threadIdx.x; i < S * (p3 - 3 * E); i += blockDim.x) {
        const int srow = i / (p3 - 3 * E), c = 3 * E + i % (p3 - 3 * E);
        const long long grow = (long long)srow * seq_stride + (long long)b * batch_stride;
        g_hi[grow * p3 + c] = __float2bfloat16_rn(0.f);
        if (g_lo) g_lo[grow * p3 + c] = __float2bfloat16_rn(0.f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// a5: additive pooling.  Group g owns rows g*L .. g*L+L-1.  score[r] = tanh(xW+b).q comes
// from the GEMM epilogue; here: w = softmax_L(score), out[g] = sum_t w_t * Y[row_t].
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
pool_fwd_kernel(const float* __restrict__ score, const float* __restrict__ Y, int E, int L,
                long long G, float* __restrict__ w_out, float* __restrict__ out) {
  pdl_launch_dependents();  // the next launch (a PDL-launched GEMM) may begin its prologue while this grid runs
  extern __shared__ float sw[];  // L weights
  __shared__ float red[33];
  // blockIdx.y = block of 4 * blockDim columns (every CTA recomputes the L softmax weights: cheap)
  const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  const bool vec4 = (E & 3) == 0;
  for (long long g = blockIdx.x; g < G; g += gridDim.x) {
    const long long r0 = g * L;
    if (L <= 32) {  // one warp, no block-wide reductions
      if (threadIdx.x < 32) {
        const int t = threadIdx.x;
        const float sc = t < L ? score[r0 + t] : -INFINITY;
        const float mx = warp_max(sc);
        const float e = t < L ? expf(sc - mx) : 0.f;
        const float sum = warp_sum(e);
        if (t < L) {
          const float w = e * (1.f / sum);
          sw[t] = w;
          if (blockIdx.y == 0) w_out[r0 + t] = w;
        }
      }
    } else {
      float mx = -INFINITY;
      for (int t = threadIdx.x; t < L; t += blockDim.x) mx = fmaxf(mx, score[r0 + t]);
      mx = block_max(mx, red);
      float sum = 0.f;
      for (int t = threadIdx.x; t < L; t += blockDim.x) {
        float e = expf(score[r0 + t] - mx);
        sw[t] = e;
        sum += e;
      }
      sum = block_sum(sum, red);
      const float inv = 1.f / sum;
      for (int t = threadIdx.x; t < L; t += blockDim.x) {
        float w = sw[t] * inv;
        sw[t] = w;
        if (blockIdx.y == 0) w_out[r0 + t] = w;
      }
    }
    __syncthreads();
    if (vec4) {  // four columns per thread, eight independent 16-byte loads in flight
      if (c0 < E) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* yc = Y + r0 * E + c0;
        int t = 0;
        for (; t + 8 <= L; t += 8) {
          float4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(yc + (long long)(t + u) * E));
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float wt = sw[t + u];
            acc.x += wt * v[u].x; acc.y += wt * v[u].y; acc.z += wt * v[u].z; acc.w += wt * v[u].w;
          }
        }
        for (; t < L; ++t) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(yc + (long long)t * E));
          const float wt = sw[t];
          acc.x += wt * v.x; acc.y += wt * v.y; acc.z += wt * v.z; acc.w += wt * v.w;
        }
        *reinterpret_cast<float4*>(out + g * E + c0) = acc;
      }
    } else {
      for (int c = c0; c < min(E, c0 + 4); ++c) {
        float acc = 0.f;
        for (int t = 0; t < L; ++t) acc += sw[t] * __ldg(Y + (r0 + t) * E + c);
        out[g * E + c] = acc;
      }
    }
    __syncthreads();
  }
}

// The same pooling with the [L][E] tile of a group staged by the TMA unit: the rows of a group are contiguous, so ONE
// cp.async.bulk (1-D bulk copy, completion on an mbarrier) brings the whole 36 KB tile into shared memory; persistent
// CTAs, two stages: thread 0 asks for the tile of the CTA's NEXT group before the current one is reduced, warp 0 computes
// the L <= 64 softmax weights meanwhile, then every thread reduces four columns out of shared memory (16-byte reads,
// conflict-free) -- no global-load instructions, no address arithmetic, the copy engine keeps the HBM requests in flight.
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__global__ void __launch_bounds__(128)
pool_fwd_tma_kernel(const float* __restrict__ score, const float* __restrict__ Y, int E, int L, long long G,
                    float* __restrict__ w_out, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char psm[];
  const uint32_t tile_bytes = (uint32_t)L * E * 4u;             // multiple of 16 (E % 4 == 0)
  const uint32_t stage_bytes = (tile_bytes + 127u) & ~127u;
  float* sw = reinterpret_cast<float*>(psm + 2 * stage_bytes);  // [64]
  const uint32_t bar0 = smem_u32(psm + 2 * stage_bytes + 256);
  const uint32_t base = smem_u32(psm);
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0 && (long long)blockIdx.x < G) {
    mbar_expect_tx(bar0, tile_bytes);
    bulk_load_1d(base, Y + (long long)blockIdx.x * L * E, tile_bytes, bar0);
  }
  int it = 0;
  for (long long g = blockIdx.x; g < G; g += gridDim.x, ++it) {
    const int st = it & 1;
    const long long gn = g + gridDim.x;
    if (threadIdx.x == 0 && gn < G) {  // stage st ^ 1 was released by the barrier that ended the previous iteration
      mbar_expect_tx(bar0 + 8 * (st ^ 1), tile_bytes);
      bulk_load_1d(base + (st ^ 1) * stage_bytes, Y + gn * L * E, tile_bytes, bar0 + 8 * (st ^ 1));
    }
    const long long r0 = g * L;
    if (threadIdx.x < 32) {  // softmax over the L <= 64 scores of the group
      const int t = threadIdx.x;
      const float s0 = t < L ? score[r0 + t] : -INFINITY, s1 = t + 32 < L ? score[r0 + t + 32] : -INFINITY;
      const float mx = warp_max(fmaxf(s0, s1));
      const float e0 = t < L ? expf(s0 - mx) : 0.f, e1 = t + 32 < L ? expf(s1 - mx) : 0.f;
      const float inv = 1.f / warp_sum(e0 + e1);
      if (t < L) { sw[t] = e0 * inv; w_out[r0 + t] = e0 * inv; }
      if (t + 32 < L) { sw[t + 32] = e1 * inv; w_out[r0 + t + 32] = e1 * inv; }
    }
    __syncthreads();
    mbar_wait(bar0 + 8 * st, (uint32_t)(it >> 1) & 1u);
    const float* tile = reinterpret_cast<const float*>(psm + st * stage_bytes);
    for (int c = threadIdx.x * 4; c < E; c += blockDim.x * 4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 5
      for (int t = 0; t < L; ++t) {
        const float4 v = *reinterpret_cast<const float4*>(tile + t * E + c);
        const float wt = sw[t];
        acc.x += wt * v.x; acc.y += wt * v.y; acc.z += wt * v.z; acc.w += wt * v.w;
      }
      *reinterpret_cast<float4*>(out + g * E + c) = acc;
    }
    __syncthreads();  // the tile and sw are free again
  }
}
__host__ __device__ inline size_t pool_fwd_tma_smem(int E, int L) {
  return 2 * (((size_t)L * E * 4 + 127) & ~(size_t)127) + 256 + 64;
}

// Backward of additive pooling (through softmax, the q-dot and tanh):
//   dY1[r]   = w_r * dOut[g]                                  (fp32, later += dApre * W_add)
//   dApre[r] = ds_r * q * (1 - A_r^2),  ds_r = w_r (dOut.Y_r - sum_u w_u dOut.Y_u)   (split planes)
//   dq      += sum_r ds_r * A_r                                (atomics, once per block)
//   db      += sum_r dApre[r]   (fp32: the terms cancel almost exactly because sum_t ds_t = 0
//                                per group, so this sum is NOT routed through the bf16 split)
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ Y,
                const float* __restrict__ w, const float* __restrict__ A,
                const float* __restrict__ qvec, int E, int Q, int qp, int L, long long G,
                float* __restrict__ dY1, __nv_bfloat16* __restrict__ da_hi,
                __nv_bfloat16* __restrict__ da_lo, float* __restrict__ dq_accum,
                float* __restrict__ db_accum) {
  pdl_launch_dependents();  // the next launch (a PDL-launched GEMM) may begin its prologue while this grid runs
  extern __shared__ float sm[];  // [L] ds
  float* s_ds = sm;
  __shared__ float red[33];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // this thread owns columns j = tid, tid + blockDim, ... of the [*, qp] tanh matrix
  float dq_acc[2] = {0.f, 0.f}, db_acc[2] = {0.f, 0.f};  // qp <= 2 * blockDim (Q <= 256... host checks)
  // 16-byte path: rows of Y / dOut / A and the plane rows start on 16 / 8-byte boundaries
  const bool vec = !(E & 3) && !(Q & 3) && !(qp & 3) && !dY1 &&
                   !((reinterpret_cast<uintptr_t>(Y) | reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(A) |
                      reinterpret_cast<uintptr_t>(qvec)) & 15) &&
                   !((reinterpret_cast<uintptr_t>(da_hi) | reinterpret_cast<uintptr_t>(da_lo)) & 7);
  const int chunks = qp / 4, rgs = vec ? max(1, (int)blockDim.x / chunks) : 1;
  float vq[4] = {0.f, 0.f, 0.f, 0.f}, vb[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long g = blockIdx.x; g < G; g += gridDim.x) {
    const long long r0 = g * L;
    // dw_t = dOut . Y_t  (one warp per row, coalesced over E; four rows in flight per warp; 16-byte loads when the
    // rows allow it: a quarter of the load instructions -- the kernel was issue-bound on 4-byte accesses)
    for (int t = warp; t < L; t += 4 * nw) {
      const float* yr[4];
      bool ok[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ok[u] = t + u * nw < L;
        yr[u] = Y + (r0 + (ok[u] ? t + u * nw : t)) * E;
      }
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      if (vec) {
        for (int c = lane; c < E / 4; c += 32) {
          const float4 dv = __ldg(reinterpret_cast<const float4*>(d_out + g * E) + c);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4 y4 = __ldg(reinterpret_cast<const float4*>(yr[u]) + c);
            acc[u] += (dv.x * y4.x + dv.y * y4.y) + (dv.z * y4.z + dv.w * y4.w);
          }
        }
      } else {
#pragma unroll 2
        for (int c = lane; c < E; c += 32) {
          const float dv = __ldg(d_out + g * E + c);
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] += dv * __ldg(yr[u] + c);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float a = warp_sum(acc[u]);
        if (lane == 0 && ok[u]) s_ds[t + u * nw] = a;
      }
    }
    __syncthreads();
    float part = 0.f;
    for (int t = threadIdx.x; t < L; t += blockDim.x) part += w[r0 + t] * s_ds[t];
    const float dbar = block_sum(part, red);
    for (int t = threadIdx.x; t < L; t += blockDim.x) s_ds[t] = w[r0 + t] * (s_ds[t] - dbar);
    __syncthreads();
    if (dY1) {  // otherwise the data-gradient GEMM adds w_r * dOut[g] in its epilogue
      for (int t = 0; t < L; ++t)
        for (int c = threadIdx.x; c < E; c += blockDim.x)
          dY1[(r0 + t) * E + c] = w[r0 + t] * d_out[g * E + c];
    }
    // single pass over A: dApre (split planes), and the dq / db partial sums of this thread's columns
    if (vec) {
      // thread = (4-column chunk, row group): 16-byte loads of A, 8-byte plane stores, four rows in flight
      if ((int)threadIdx.x < rgs * chunks) {
        const int ch = threadIdx.x % chunks, rg = threadIdx.x / chunks, j0 = 4 * ch;
        const bool in = j0 < Q;  // Q % 4 == 0: a chunk is all data or all padding
        const float4 q4 = in ? __ldg(reinterpret_cast<const float4*>(qvec + j0)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float qv[4] = {q4.x, q4.y, q4.z, q4.w};
        for (int t0 = rg; t0 < L; t0 += 4 * rgs) {
          float4 a4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * rgs;
            a4[u] = (in && t < L) ? __ldg(reinterpret_cast<const float4*>(A + (r0 + t) * Q + j0)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int t = t0 + u * rgs;
            if (t < L) {
              const float sd = s_ds[t];
              const float av[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w};
              __nv_bfloat16 hh[4], ll[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const float v = sd * qv[k] * (1.f - av[k] * av[k]);
                split_bf16(v, hh[k], ll[k]);
                vq[k] += sd * av[k];
                vb[k] -= sd * qv[k] * (av[k] * av[k]);  // exact-cancellation form of db, see the scalar path below
              }
              const long long off = (r0 + t) * qp + j0;
              *reinterpret_cast<uint2*>(da_hi + off) = make_uint2(pack_bf16x2(hh[0], hh[1]), pack_bf16x2(hh[2], hh[3]));
              if (da_lo)
                *reinterpret_cast<uint2*>(da_lo + off) = make_uint2(pack_bf16x2(ll[0], ll[1]), pack_bf16x2(ll[2], ll[3]));
            }
          }
        }
      }
    } else {
      int slot = 0;
      for (int j = threadIdx.x; j < qp; j += blockDim.x, ++slot) {
        const bool in = j < Q;
        const float qj = in ? qvec[j] : 0.f;
        float aq = 0.f, ab = 0.f;
        for (int t0 = 0; t0 < L; t0 += 10) {
          float a[10];
#pragma unroll
          for (int u = 0; u < 10; ++u) a[u] = (in && t0 + u < L) ? __ldg(A + (r0 + t0 + u) * Q + j) : 0.f;
#pragma unroll
          for (int u = 0; u < 10; ++u) {
            if (t0 + u < L) {
              const float sd = s_ds[t0 + u];
              const float v = sd * qj * (1.f - a[u] * a[u]);
              __nv_bfloat16 hh, ll;
              split_bf16(v, hh, ll);
              da_hi[(r0 + t0 + u) * qp + j] = hh;
              if (da_lo) da_lo[(r0 + t0 + u) * qp + j] = ll;
              aq += sd * a[u];
              // db_j = sum_r ds_r q_j (1 - a_rj^2) = -q_j sum_r ds_r a_rj^2: sum_t ds_t = 0 within every softmax
              // group EXACTLY (ds_t = w_t (dw_t - sum_u w_u dw_u), sum_t w_t = 1), so the "1" part only contributes
              // its own fp32 rounding noise -- which dominates this gradient when the rows of a group resemble each
              // other (a_rj^2 nearly constant over r: the result is then a second cancellation on top of the first)
              ab -= sd * qj * (a[u] * a[u]);
            }
          }
        }
        dq_acc[slot] += aq;
        db_acc[slot] += ab;
      }
    }
    __syncthreads();
  }
  if (vec) {  // row groups -> shared memory -> ONE global atomic per column per CTA (as many as the scalar path issues)
    __shared__ float s_acc[2][256];
    for (int j = threadIdx.x; j < 2 * 256; j += blockDim.x) (&s_acc[0][0])[j] = 0.f;
    __syncthreads();
    if ((int)threadIdx.x < rgs * chunks) {
      const int j0 = 4 * ((int)threadIdx.x % chunks);
      if (j0 < Q) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          atomicAdd(&s_acc[0][j0 + k], vq[k]);
          atomicAdd(&s_acc[1][j0 + k], vb[k]);
        }
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < Q; j += blockDim.x) {
      atomicAdd(dq_accum + j, s_acc[0][j]);
      atomicAdd(db_accum + j, s_acc[1][j]);
    }
    return;
  }
  int slot = 0;
  for (int j = threadIdx.x; j < Q; j += blockDim.x, ++slot) {
    atomicAdd(dq_accum + j, dq_acc[slot]);
    atomicAdd(db_accum + j, db_acc[slot]);
  }
}

// ------------------------------------------------------------------------------------
// a8: ragged -> dense (to_dense_batch) and back.  off[B+1] are CSR offsets of the sorted
// segment ids.  Dense rows beyond a segment's length are zero (their ones column stays 1:
// padded history rows still receive the in-projection bias, as in the reference).
// ------------------------------------------------------------------------------------
// max_count > 0: a segment longer than the dense width the caller announced (Hmax / Cmax), or ids outside
// [0, B) (off[B] != n, seg[0] < 0), set the device error flag -- the dense kernels clip such rows.
__global__ void segment_offsets_kernel(const long long* __restrict__ seg, long long n, int B,
                                       int* __restrict__ off, int max_count) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b > B) return;
  long long lo = 0, hi = n;  // first index with seg[i] >= b
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if (seg[mid] < b) lo = mid + 1; else hi = mid;
  }
  off[b] = (int)lo;
  if (max_count > 0) {
    if (b == B && lo != n) dev_error(DEV_ERR_SEGMENT_ID);
    if (b == 0 && (lo != 0 || (n > 0 && seg[0] < 0))) dev_error(DEV_ERR_SEGMENT_ID);
    if (b < B) {
      long long l2 = lo, h2 = n;  // first index with seg[i] >= b + 1
      while (l2 < h2) {
        long long mid = (l2 + h2) >> 1;
        if (seg[mid] < b + 1) l2 = mid + 1; else h2 = mid;
      }
      if (l2 - lo > max_count) dev_error(DEV_ERR_SEGMENT_LEN);
    }
  }
}

// history and candidate offsets of one batch in one launch (blockIdx.y = 0 / 1)
__device__ __forceinline__ void segment_offsets_body(const long long* __restrict__ seg, long long n, int B,
                                                     int* __restrict__ off, int max_count, int b) {
  if (b > B) return;
  long long lo = 0, hi = n;
  while (lo < hi) {
    long long mid = (lo + hi) >> 1;
    if (seg[mid] < b) lo = mid + 1; else hi = mid;
  }
  off[b] = (int)lo;
  if (max_count > 0) {
    if (b == B && lo != n) dev_error(DEV_ERR_SEGMENT_ID);
    if (b == 0 && (lo != 0 || (n > 0 && seg[0] < 0))) dev_error(DEV_ERR_SEGMENT_ID);
    if (b < B) {
      long long l2 = lo, h2 = n;
      while (l2 < h2) {
        long long mid = (l2 + h2) >> 1;
        if (seg[mid] < b + 1) l2 = mid + 1; else h2 = mid;
      }
      if (l2 - lo > max_count) dev_error(DEV_ERR_SEGMENT_LEN);
    }
  }
}
__global__ void segment_offsets2_kernel(const long long* __restrict__ seg0, long long n0, int* __restrict__ off0,
                                        int max0, const long long* __restrict__ seg1, long long n1,
                                        int* __restrict__ off1, int max1, int B) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (blockIdx.y == 0) segment_offsets_body(seg0, n0, B, off0, max0, b);
  else segment_offsets_body(seg1, n1, B, off1, max1, b);
}

__global__ void dense_scatter_kernel(const float* __restrict__ x, const int* __restrict__ off,
                                     int B, int M, int E, int ep, float* __restrict__ dense,
                                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  pdl_launch_dependents();  // the next launch (a PDL-launched GEMM) may begin its prologue while this grid runs
  for (long long r = blockIdx.x; r < (long long)B * M; r += gridDim.x) {
    const int b = (int)(r / M), j = (int)(r % M);
    // off == nullptr: identity (every dense row present, M == 1 per "segment")
    const int cnt = off ? off[b + 1] - off[b] : M;
    const float* src = j < cnt ? x + (off ? (long long)(off[b] + j) : r) * E : nullptr;
    for (int c = threadIdx.x; c < ep; c += blockDim.x) {
      float v = c < E ? (src ? src[c] : 0.f) : (c == E ? 1.f : 0.f);
      if (dense && c < E) dense[r * E + c] = v;
      if (hi) {
        __nv_bfloat16 hh, ll;
        split_bf16(v, hh, ll);
        hi[r * ep + c] = hh;
        if (lo) lo[r * ep + c] = ll;
      }
    }
  }
}

__global__ void dense_gather_kernel(const float* __restrict__ d_dense, const int* __restrict__ off,
                                    int B, int M, int E, float* __restrict__ dx) {
  pdl_launch_dependents();  // the next launch (a PDL-launched GEMM) may begin its prologue while this grid runs
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    const int cnt = min(off[b + 1] - off[b], M);  // longer segments were flagged by segment_offsets_kernel
    for (int j = blockIdx.x; j < cnt; j += gridDim.x)
      for (int c = threadIdx.x; c < E; c += blockDim.x)
        dx[(long long)(off[b] + j) * E + c] = d_dense[((long long)b * M + j) * E + c];
  }
}

// a10: late fusion, user = sum_j dense[b, j, :] / count_b  (nrms_module.py:243-248)
__global__ void late_fusion_fwd_kernel(const float* __restrict__ x, const int* __restrict__ off,
                                       int B, int E, float* __restrict__ user) {
  const int b = blockIdx.x;
  const int cnt = off[b + 1] - off[b];
  for (int c = threadIdx.x; c < E; c += blockDim.x) {
    float acc = 0.f;
    for (int j = 0; j < cnt; ++j) acc += x[(long long)(off[b] + j) * E + c];
    user[(long long)b * E + c] = acc / (float)cnt;
  }
}
__global__ void late_fusion_bwd_kernel(const float* __restrict__ d_user, const int* __restrict__ off,
                                       int B, int E, float* __restrict__ dx) {
  const int b = blockIdx.x;
  const int cnt = off[b + 1] - off[b];
  for (int j = 0; j < cnt; ++j)
    for (int c = threadIdx.x; c < E; c += blockDim.x)
      dx[(long long)(off[b] + j) * E + c] = d_user[(long long)b * E + c] / (float)cnt;
}

// ------------------------------------------------------------------------------------
// a11: scorer.  scores[b, c] = user_b . cand[off[b]+c]  (exactly 0.0 in padded slots).
// One warp per (b, c).
// ------------------------------------------------------------------------------------
__global__ void score_fwd_kernel(const float* __restrict__ user, const float* __restrict__ cand,
                                 const int* __restrict__ off, int B, int C, int E,
                                 float* __restrict__ scores) {
  const int lane = threadIdx.x & 31;
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (wid >= (long long)B * C) return;
  const int b = (int)(wid / C), c = (int)(wid % C);
  const int cnt = off[b + 1] - off[b];
  float acc = 0.f;
  if (c < cnt) {
    const float* u = user + (long long)b * E;
    const float* n = cand + (long long)(off[b] + c) * E;
    for (int i = lane; i < E; i += 32) acc += u[i] * n[i];
    acc = warp_sum(acc);
  }
  if (lane == 0) scores[wid] = acc;
}
// d_user[b] = sum_c ds[b,c] cand[b,c];  d_cand[off[b]+c] = ds[b,c] user[b].  Block per b.
__global__ void score_bwd_kernel(const float* __restrict__ d_scores, const float* __restrict__ user,
                                 const float* __restrict__ cand, const int* __restrict__ off, int B,
                                 int C, int E, float* __restrict__ d_user, float* __restrict__ d_cand) {
  const int b = blockIdx.x;
  const int cnt = min(off[b + 1] - off[b], C);
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    const float u = user[(long long)b * E + i];
    float acc = 0.f;
    for (int c = 0; c < cnt; ++c) {
      const float ds = d_scores[(long long)b * C + c];
      acc += ds * cand[(long long)(off[b] + c) * E + i];
      d_cand[(long long)(off[b] + c) * E + i] = ds * u;
    }
    d_user[(long long)b * E + i] = acc;
  }
}

// ------------------------------------------------------------------------------------
// a12: CrossEntropyLoss with float targets over the dense [B, C] score matrix, mean over B.
// Labels arrive ragged ([N_c], same offsets as the candidates); padded slots have target 0
// and score 0 and take part in the softmax.  One warp per row.
// ------------------------------------------------------------------------------------
__global__ void ce_fwd_kernel(const float* __restrict__ scores, const float* __restrict__ labels,
                              const int* __restrict__ off, int B, int C, float* __restrict__ loss_rows,
                              float* __restrict__ loss_mean, float* __restrict__ y_dense) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  const int cnt = off[b + 1] - off[b];
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, scores[(long long)b * C + c]);
  mx = warp_max(mx);
  float se = 0.f;
  for (int c = lane; c < C; c += 32) se += expf(scores[(long long)b * C + c] - mx);
  se = warp_sum(se);
  const float lz = mx + logf(se);
  float acc = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float y = c < cnt ? labels[off[b] + c] : 0.f;
    if (y_dense) y_dense[(long long)b * C + c] = y;
    acc += y * (lz - scores[(long long)b * C + c]);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    if (loss_rows) loss_rows[b] = acc;
    if (loss_mean) atomicAdd(loss_mean, acc / (float)B);
  }
}
// d s[b,c] = g/B * (softmax(s)[b,c] * sum_c' y[b,c'] - y[b,c])
__global__ void ce_bwd_kernel(const float* __restrict__ scores, const float* __restrict__ labels,
                              const int* __restrict__ off, int B, int C,
                              const float* __restrict__ g_loss, float g_scale,
                              float* __restrict__ d_scores) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  const int cnt = off[b + 1] - off[b];
  const float g = (g_loss ? g_loss[0] : 1.f) * g_scale / (float)B;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, scores[(long long)b * C + c]);
  mx = warp_max(mx);
  float se = 0.f, sy = 0.f;
  for (int c = lane; c < C; c += 32) {
    se += expf(scores[(long long)b * C + c] - mx);
    sy += c < cnt ? labels[off[b] + c] : 0.f;
  }
  se = warp_sum(se);
  sy = warp_sum(sy);
  for (int c = lane; c < C; c += 32) {
    const float y = c < cnt ? labels[off[b] + c] : 0.f;
    const float pr = expf(scores[(long long)b * C + c] - mx) / se;
    d_scores[(long long)b * C + c] = g * (pr * sy - y);
  }
}

// ------------------------------------------------------------------------------------
// f4: per-impression retrieval metrics of the epoch-end hooks (nrms_module.py:182-191,380-396: torchmetrics
// RetrievalMRR and RetrievalNormalizedDCG(top_k) grouped by `indexes` = the impression of each candidate).
// One CTA per impression over the ragged epoch outputs (scores / labels concatenated per impression, 64-bit offsets):
//   rank_i  = 1 + #{j : s_j > s_i or (s_j == s_i and j < i)}        (descending, stable -- an O(c^2) count in shared
//   lrank_i = the same over the labels                                 memory instead of a sort: c <= a few hundred)
//   mrr     = 1 / min{rank_i : y_i > 0}                   (0 without a positive: empty_target_action = "neg")
//   ndcg@k  = sum_{rank_i <= k} y_i / log2(rank_i + 1)  /  sum_{lrank_i <= k} y_i / log2(lrank_i + 1)   (0 when the ideal is 0)
// out [B][1 + nk]; ranks [N] (optional) receives rank_i for callers that need the top-k membership.
// Impressions longer than RM_SMEM candidates are ranked straight from global memory.
// ------------------------------------------------------------------------------------
constexpr int RM_SMEM = 2048;
constexpr int RM_MAXK = 4;
struct RankKs {
  int n;
  int k[RM_MAXK];
};
__global__ void __launch_bounds__(128)
rank_metrics_kernel(const float* __restrict__ scores, const float* __restrict__ labels,
                    const long long* __restrict__ off, int B, RankKs ks, float* __restrict__ out,
                    int* __restrict__ ranks) {
  __shared__ float s_s[RM_SMEM], s_y[RM_SMEM];
  __shared__ float s_red[4][2 * RM_MAXK + 1];
  const int b = blockIdx.x;
  const long long o = off[b];
  const int c = (int)(off[b + 1] - o);
  const float* gs = scores + o;
  const float* gy = labels + o;
  const bool in_smem = c <= RM_SMEM;
  if (in_smem) {
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
      s_s[i] = gs[i];
      s_y[i] = gy[i];
    }
    __syncthreads();
  }
  const float* ps = in_smem ? s_s : gs;
  const float* py = in_smem ? s_y : gy;
  float dcg[RM_MAXK], idcg[RM_MAXK];
#pragma unroll
  for (int q = 0; q < RM_MAXK; ++q) dcg[q] = idcg[q] = 0.f;
  float best = 0.f;  // 1 / (smallest rank of a positive) seen by this thread
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    const float si = ps[i], yi = py[i];
    int r = 1, lr = 1;
    for (int j = 0; j < c; ++j) {
      const float sj = ps[j], yj = py[j];
      r += (sj > si) || (sj == si && j < i);
      lr += (yj > yi) || (yj == yi && j < i);
    }
    if (ranks) ranks[o + i] = r;
    if (yi > 0.f) best = fmaxf(best, 1.f / (float)r);
#pragma unroll
    for (int q = 0; q < RM_MAXK; ++q)
      if (q < ks.n) {
        if (r <= ks.k[q]) dcg[q] += yi / log2f((float)r + 1.f);
        if (lr <= ks.k[q]) idcg[q] += yi / log2f((float)lr + 1.f);
      }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  best = warp_max(best);
#pragma unroll
  for (int q = 0; q < RM_MAXK; ++q) {
    dcg[q] = warp_sum(dcg[q]);
    idcg[q] = warp_sum(idcg[q]);
  }
  if (lane == 0) {
    s_red[warp][0] = best;
#pragma unroll
    for (int q = 0; q < RM_MAXK; ++q) {
      s_red[warp][1 + 2 * q] = dcg[q];
      s_red[warp][2 + 2 * q] = idcg[q];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float* ob = out + (long long)b * (1 + ks.n);
    ob[0] = fmaxf(fmaxf(s_red[0][0], s_red[1][0]), fmaxf(s_red[2][0], s_red[3][0]));
    for (int q = 0; q < ks.n; ++q) {
      const float d_ = s_red[0][1 + 2 * q] + s_red[1][1 + 2 * q] + s_red[2][1 + 2 * q] + s_red[3][1 + 2 * q];
      const float i_ = s_red[0][2 + 2 * q] + s_red[1][2 + 2 * q] + s_red[2][2 + 2 * q] + s_red[3][2 + 2 * q];
      ob[1 + q] = i_ > 0.f ? d_ / i_ : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------
// f4: supervised-contrastive loss over the dense score matrix (nrms_module.py:290-316 + components/losses.py:12-40 on
// pytorch-metric-learning's SupConLoss; temperature is the constructor default 0.1, abstract_recommender.py:117-120).
// Row b: positives = real slots with a non-zero label, negatives = real slots with label 0, padded slots in neither.
//   x = s / T - max_c(s / T)   (maximum over ALL C slots, padded zeros included; detached)
//   loss_b = -sum_pos (x - logsumexp_kept x) / (npos + FLT_MIN)
// The batch loss is the mean over the rows with loss_b > 0 (AvgNonZeroReducer); zero when every index list of the
// batch has at most one element, or there is no positive / no negative at all.  ONE CTA (the batch reduction is a
// count of rows): warp w takes rows w, w + nwarps, ...   stats = {loss, rows counted, 1 if the batch loss is live}.
// With `ce_loss` given the value written to `out_loss` is the dual loss (1 - coef) * CE + coef * SupCon
// (nrms_module.py:326-328).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void supcon_row(const float* __restrict__ s, const float* __restrict__ y, int cnt, int C,
                                           float inv_t, int lane, float& mx, float& lse, float& sum_pos, int& npos) {
  mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, s[c] * inv_t);
  mx = warp_max(mx);
  float m2 = -INFINITY;
  for (int c = lane; c < cnt; c += 32) m2 = fmaxf(m2, s[c] * inv_t - mx);
  m2 = warp_max(m2);
  float se = 0.f;
  sum_pos = 0.f;
  npos = 0;
  for (int c = lane; c < cnt; c += 32) {
    const float x = s[c] * inv_t - mx;
    se += expf(x - m2);
    if (y[c] != 0.f) {
      sum_pos += x;
      ++npos;
    }
  }
  se = warp_sum(se);
  sum_pos = warp_sum(sum_pos);
  npos = (int)warp_sum((float)npos);
  lse = cnt > 0 ? m2 + logf(se) : 0.f;
}

__global__ void __launch_bounds__(1024)
supcon_fwd_kernel(const float* __restrict__ scores, const float* __restrict__ labels, const int* __restrict__ off,
                  int B, int C, float inv_t, const float* __restrict__ ce_loss, float coef,
                  float* __restrict__ row_loss, float* __restrict__ out_loss, float* __restrict__ stats) {
  __shared__ float s_sum[32], s_rows[32], s_pos[32], s_neg[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float sum = 0.f, rows = 0.f, tot_pos = 0.f, tot_neg = 0.f;
  for (int b = warp; b < B; b += nw) {
    const int cnt = min(off[b + 1] - off[b], C);
    float mx, lse, sp;
    int np;
    supcon_row(scores + (long long)b * C, labels + off[b], cnt, C, inv_t, lane, mx, lse, sp, np);
    // sum_pos (x - lse) = sum_pos x - npos * lse
    const float l = -(sp - (float)np * lse) / ((float)np + FLT_MIN);
    if (lane == 0) row_loss[b] = l;
    if (l > 0.f) {
      sum += l;
      rows += 1.f;
    }
    tot_pos += (float)np;
    tot_neg += (float)(cnt - np);
  }
  if (lane == 0) {
    s_sum[warp] = sum;
    s_rows[warp] = rows;
    s_pos[warp] = tot_pos;
    s_neg[warp] = tot_neg;
  }
  __syncthreads();
  if (warp == 0) {
    sum = warp_sum(lane < nw ? s_sum[lane] : 0.f);
    rows = warp_sum(lane < nw ? s_rows[lane] : 0.f);
    tot_pos = warp_sum(lane < nw ? s_pos[lane] : 0.f);
    tot_neg = warp_sum(lane < nw ? s_neg[lane] : 0.f);
    if (lane == 0) {
      const bool live = !(tot_pos <= 1.f && tot_neg <= 1.f) && tot_pos > 0.f && tot_neg > 0.f && rows > 0.f;
      const float scl = live ? sum / rows : 0.f;
      stats[0] = scl;
      stats[1] = rows;
      stats[2] = live ? 1.f : 0.f;
      out_loss[0] = ce_loss ? (1.f - coef) * ce_loss[0] + coef * scl : scl;
    }
  }
}
// d s[b,c] = g / (rows * T) * (softmax_kept(x)[c] * npos / (npos + FLT_MIN) - [c positive] / (npos + FLT_MIN)) for the real
// slots of the rows that were counted; everything else 0.  `accumulate` adds to d_scores (dual loss after ce_bwd).
__global__ void supcon_bwd_kernel(const float* __restrict__ scores, const float* __restrict__ labels,
                                  const int* __restrict__ off, int B, int C, float inv_t,
                                  const float* __restrict__ row_loss, const float* __restrict__ stats,
                                  const float* __restrict__ g_loss, float g_scale, int accumulate,
                                  float* __restrict__ d_scores) {
  const int lane = threadIdx.x & 31;
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  float* d = d_scores + (long long)b * C;
  const bool counted = stats[2] != 0.f && row_loss[b] > 0.f;
  if (!counted) {
    if (!accumulate)
      for (int c = lane; c < C; c += 32) d[c] = 0.f;
    return;
  }
  const int cnt = min(off[b + 1] - off[b], C);
  const float* s = scores + (long long)b * C;
  const float* y = labels + off[b];
  float mx, lse, sp;
  int np;
  supcon_row(s, y, cnt, C, inv_t, lane, mx, lse, sp, np);
  const float g = (g_loss ? g_loss[0] : 1.f) * g_scale * inv_t / stats[1];
  const float inv_np = 1.f / ((float)np + FLT_MIN);
  for (int c = lane; c < C; c += 32) {
    float v = 0.f;
    if (c < cnt) {
      const float pr = expf(s[c] * inv_t - mx - lse);
      v = g * (pr * (float)np * inv_np - (y[c] != 0.f ? inv_np : 0.f));
    }
    d[c] = accumulate ? d[c] + v : v;
  }
}

// a11 + a12 (+ their backward) in ONE launch for the fused step: CTA b scores impression b's candidates, takes the
// soft-target CE of the padded row and, when d_scores is given, goes straight on to d_scores, d_user and d_cand --
// the four kernels above are each a few microseconds of launch latency on a dependent chain.  Same arithmetic, same
// order of operations per row as score_fwd / ce_fwd / ce_bwd / score_bwd.  Dynamic smem: 2 * C floats.
__global__ void __launch_bounds__(128)
score_loss_kernel(const float* __restrict__ user, const float* __restrict__ cand, const float* __restrict__ labels,
                  const int* __restrict__ off, int B, int C, int E, float* __restrict__ scores,
                  float* __restrict__ loss_mean, const float* __restrict__ g_loss, float* __restrict__ d_scores,
                  float* __restrict__ d_user, float* __restrict__ d_cand) {
  pdl_launch_dependents();  // the next launch (a PDL-launched GEMM) may begin its prologue while this grid runs
  extern __shared__ float sl_smem[];
  float* s_sc = sl_smem;       // [C] scores of this row
  float* s_ds = sl_smem + C;   // [C] d loss / d scores
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int cnt = min(off[b + 1] - off[b], C);
  const float* u = user + (long long)b * E;
  for (int c = warp; c < C; c += nw) {
    float acc = 0.f;
    if (c < cnt) {
      const float* n = cand + (long long)(off[b] + c) * E;
      for (int i = lane; i < E; i += 32) acc += u[i] * n[i];
      acc = warp_sum(acc);
    }
    if (lane == 0) {
      s_sc[c] = acc;
      scores[(long long)b * C + c] = acc;
    }
  }
  __syncthreads();
  if (!loss_mean && !d_scores) return;
  if (warp == 0) {
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, s_sc[c]);
    mx = warp_max(mx);
    float se = 0.f, sy = 0.f;
    for (int c = lane; c < C; c += 32) {
      se += expf(s_sc[c] - mx);
      sy += c < cnt ? labels[off[b] + c] : 0.f;
    }
    se = warp_sum(se);
    sy = warp_sum(sy);
    const float lz = mx + logf(se);
    float acc = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float y = c < cnt ? labels[off[b] + c] : 0.f;
      acc += y * (lz - s_sc[c]);
      if (d_scores) {
        const float pr = expf(s_sc[c] - mx) / se;
        float ds = (pr * sy - y) / (float)B;
        if (g_loss) ds *= g_loss[0];
        s_ds[c] = ds;
        d_scores[(long long)b * C + c] = ds;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0 && loss_mean) atomicAdd(loss_mean, acc / (float)B);
  }
  if (!d_scores) return;
  __syncthreads();
  for (int i = threadIdx.x; i < E; i += blockDim.x) {
    const float ui = u[i];
    float acc = 0.f;
    for (int c = 0; c < cnt; ++c) {
      const float ds = s_ds[c];
      acc += ds * cand[(long long)(off[b] + c) * E + i];
      d_cand[(long long)(off[b] + c) * E + i] = ds * ui;
    }
    d_user[(long long)b * E + i] = acc;
  }
}

// ------------------------------------------------------------------------------------
// Embedding gradient: d_table[ids[r]] += dX[r]  (dense [V1, E] table gradient, row 0 = the
// padding_idx row is skipped so its gradient stays exactly zero, text.py:215-217).
// One warp per token row, 128-bit vector atomics.
// ------------------------------------------------------------------------------------
__global__ void emb_grad_kernel(const long long* __restrict__ ids, long long R, long long V1,
                                const float* __restrict__ dX, int E, float* __restrict__ d_table) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = warp0; r < R; r += nwarps) {
    const long long id = ids[r];
    if (id == 0) continue;
    if (id < 0 || id >= V1) {  // never write outside the table: flag and skip
      if (lane == 0) dev_error(DEV_ERR_TOKEN_ID);
      continue;
    }
    const float* src = dX + r * E;
    float* dst = d_table + id * E;
    if ((E & 3) == 0) {
      for (int c = lane * 4; c < E; c += 128) {
        float4 v = *reinterpret_cast<const float4*>(src + c);
        atomicAdd(reinterpret_cast<float4*>(dst + c), v);
      }
    } else {
      for (int c = lane; c < E; c += 32) atomicAdd(dst + c, src[c]);
    }
  }
}

// torch.optim.Adam (no weight decay / amsgrad), dense over n elements; 128-bit accesses over the
// 16-byte-aligned body (the flat parameter buffers are), scalar tail.
// zero_grad != 0: the gradient element is overwritten with 0 after it has been consumed (optimizer.zero_grad()
// folded into the step: the buffers accumulate, so they must be clean before the next backward pass).
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr, float b1, float b2, float omb1,
                            float omb2, float eps, float bc1, float sqrt_bc2, float g_scale, int zero_grad) {
  const float lr_bc1 = lr / bc1;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) |
                     reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = tid; i < n4; i += nth) {
    float4 p4 = reinterpret_cast<float4*>(p)[i];
    const float4 g4 = reinterpret_cast<const float4*>(g)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
    {  // sqrt(v) / sqrt_bc2 is kept as a division (torch's formula), not a reciprocal multiply
      float* pp = &p4.x; const float* gg = &g4.x; float* mm = &m4.x; float* vv = &v4.x;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float gi = gg[k] * g_scale;
        const float mi = b1 * mm[k] + omb1 * gi;   // omb = fp32(1 - beta) formed in double on the host, as torch does
        const float vi = b2 * vv[k] + omb2 * gi * gi;
        mm[k] = mi;
        vv[k] = vi;
        const float denom = sqrtf(vi) / sqrt_bc2 + eps;
        pp[k] -= lr_bc1 * (mi / denom);
      }
    }
    reinterpret_cast<float4*>(p)[i] = p4;
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
    // after every load of the iteration has been issued: a store in between serialises the loads behind it
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long i = 4 * n4 + tid; i < n; i += nth) {
    const float gi = g[i] * g_scale;
    if (zero_grad) g[i] = 0.f;
    const float mi = b1 * m[i] + omb1 * gi;
    const float vi = b2 * v[i] + omb2 * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrt_bc2 + eps;
    p[i] -= lr_bc1 * (mi / denom);
  }
}

// Keep-bit words of the two dropout sites of one encoder pass, drawn ONCE per step by all SMs
// (the Philox rounds are a long dependent chain: far too slow for the four epilogue warps of
// the GEMM) and read back by the gather kernel, the GEMM epilogues and the backward pass:
// words[site][r * mw + w], bit i = element (r, 32 w + i) of site `site` is kept.
// row0: index of this buffer's first row in the numbering the draws follow (a pass split over two buffers keeps the bits
// of the unsplit pass).
__global__ void dropout_words_kernel(unsigned long long seed, uint32_t thr, long long R, int E, int mw,
                                     uint32_t* __restrict__ w0, uint32_t* __restrict__ w1, long long row0 = 0) {
  const long long per_site = R * mw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < 2 * per_site;
       i += (long long)gridDim.x * blockDim.x) {
    const uint32_t site = i >= per_site ? 1u : 0u;
    const long long j = i - site * per_site;
    const long long r = j / mw;
    const int w = (int)(j - r * mw);
    const uint32_t bits = drop_keep_bits32(seed, site, (unsigned long long)(r + row0) * (unsigned)E + 32u * (unsigned)w, thr);
    (site ? w1 : w0)[j] = bits;
  }
}

__global__ void dropout_mask_kernel(unsigned char* __restrict__ keep, long long n,
                                    unsigned long long seed, uint32_t site, uint32_t thr) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    keep[i] = drop_keep(seed, site, (unsigned long long)i, thr) ? 1 : 0;
}

// Row gather out[i, :] = table[idx[i], :] over rows of `words` W-byte words (W = 8 for token-id /
// label rows, 4 for fp32 rows): the device-side collate (DatasetCollate._tokenize_df,
// rec_dataset.py:189-285, over a news table already resident in HBM) and the news-vector cache
// lookup of the evaluation path.  Consecutive threads copy consecutive words: coalesced both ways.
template <typename Wd>
__global__ void gather_rows_kernel(const Wd* __restrict__ table, long long n_rows, int words,
                                   const long long* __restrict__ idx, long long n, Wd* __restrict__ out) {
  const long long total = n * words;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / words;
    const int w = (int)(i - r * words);
    long long src = idx[r];
    if (src < 0 || src >= n_rows) {  // flagged; row 0 is copied instead of reading out of bounds
      if (w == 0) dev_error(DEV_ERR_ROW_INDEX);
      src = 0;
    }
    out[i] = __ldg(table + src * words + w);
  }
}

// out[i] = hi[i] + lo[i]  (test helper: read back a split-plane GEMM sink)
__global__ void planes_to_f32_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                     long long n, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = __bfloat162float(hi[i]) + (lo ? __bfloat162float(lo[i]) : 0.f);
}

// acc[i] += x[i]
__global__ void add_inplace_kernel(float* __restrict__ acc, const float* __restrict__ x, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    acc[i] += x[i];
}

}  // namespace nrl
