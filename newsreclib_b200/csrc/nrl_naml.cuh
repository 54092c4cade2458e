// NAML-specific memory-bound kernels (sm_100a, fp32 SIMT): the CNN text encoder's
// gather + im2col (Conv2d(1, F, (w, E), padding=((w-1)/2, 0)) becomes ONE K = w*E tensor-core
// GEMM, reference encoders/news/text.py:155-170), its transpose (col2im fused with the
// embedding-gradient scatter), and the ReLU backward of the category encoder
// (encoders/news/category.py:73-82).
#pragma once
#include "nrl_kernels.cuh"

namespace nrl {

// a13: token ids [N][L] -> split planes A[2][R = N*L][kp], row (n, t) =
//   [ x(n, t-pad) | x(n, t-pad+1) | ... | x(n, t-pad+w-1) | 1 | 0... ]      (x = dropout0(table[id]))
// with x(n, t') = 0 outside [0, L) (the conv's zero padding along the token axis).
// One warp per output row; the keep-bit words belong to the SOURCE token row.
__global__ void gather_im2col_kernel(const long long* __restrict__ ids, long long N, int L,
                                     const float* __restrict__ table, long long V1, int E, int win, int kp,
                                     __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                     const uint32_t* __restrict__ drop_words, int drop_mw,
                                     float drop_scale) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long R = N * L;
  const int pad = (win - 1) / 2;
  for (long long r = warp0; r < R; r += nwarps) {
    const long long n = r / L;
    const int t = (int)(r - n * L);
    __nv_bfloat16* hrow = hi + r * kp;
    __nv_bfloat16* lrow = lo ? lo + r * kp : nullptr;
    for (int j = 0; j < win; ++j) {
      const int ts = t + j - pad;
      const bool in = ts >= 0 && ts < L;
      const long long rs = n * L + ts;
      long long id = in ? ids[rs] : 0;
      if (id < 0 || id >= V1) {  // flagged; row 0 is read instead (nn.Embedding raises here)
        if (lane == 0) dev_error(DEV_ERR_TOKEN_ID);
        id = 0;
      }
      const float* src = in ? table + id * E : nullptr;
      for (int c = lane * 4; c < E; c += 128) {  // E % 4 == 0 (host checks)
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (in) {
          const float4 t4 = __ldg(reinterpret_cast<const float4*>(src + c));
          v[0] = t4.x; v[1] = t4.y; v[2] = t4.z; v[3] = t4.w;
          if (drop_words) {
            const uint32_t bits = __ldg(drop_words + rs * drop_mw + (c >> 5)) >> (c & 31);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[i] = ((bits >> i) & 1u) ? v[i] * drop_scale : 0.f;
          }
        }
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_bf16(v[i], h[i], l[i]);
        const int off = j * E + c;  // 8-byte aligned: E % 4 == 0, kp % 8 == 0
        *reinterpret_cast<uint2*>(hrow + off) = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
        if (lrow)
          *reinterpret_cast<uint2*>(lrow + off) = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
      }
    }
    for (int c = win * E + lane; c < kp; c += 32) {
      hrow[c] = __float2bfloat16_rn(c == win * E ? 1.f : 0.f);
      if (lrow) lrow[c] = __float2bfloat16_rn(0.f);
    }
  }
}

// Transpose of the above fused with the embedding-gradient scatter:
//   dx(n, t') = dropout0'( sum_j dA[(n, t' - j + pad)][j*E : (j+1)*E] ),   d_table[id(n, t')] += dx
// (row 0 of the table = padding_idx is skipped, text.py:151-153).  One warp per source token.
__global__ void col2im_emb_grad_kernel(const long long* __restrict__ ids, long long N, int L, long long V1,
                                       const float* __restrict__ dA, long long ld_da, int E, int win,
                                       const uint32_t* __restrict__ drop_words, int drop_mw,
                                       float drop_scale, float* __restrict__ d_table) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const long long R = N * L;
  const int pad = (win - 1) / 2;
  for (long long rs = warp0; rs < R; rs += nwarps) {
    const long long id = ids[rs];
    if (id == 0) continue;
    if (id < 0 || id >= V1) {  // never write outside the table: flag and skip
      if (lane == 0) dev_error(DEV_ERR_TOKEN_ID);
      continue;
    }
    const long long n = rs / L;
    const int ts = (int)(rs - n * L);
    float* dst = d_table + id * E;
    for (int c = lane * 4; c < E; c += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = 0; j < win; ++j) {
        const int t = ts - j + pad;
        if (t < 0 || t >= L) continue;
        const float4 g = __ldg(reinterpret_cast<const float4*>(dA + (n * L + t) * ld_da + j * E + c));
        acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
      }
      if (drop_words) {
        const uint32_t bits = __ldg(drop_words + rs * drop_mw + (c >> 5)) >> (c & 31);
        acc.x = (bits & 1u) ? acc.x * drop_scale : 0.f;
        acc.y = (bits & 2u) ? acc.y * drop_scale : 0.f;
        acc.z = (bits & 4u) ? acc.z * drop_scale : 0.f;
        acc.w = (bits & 8u) ? acc.w * drop_scale : 0.f;
      }
      atomicAdd(reinterpret_cast<float4*>(dst + c), acc);
    }
  }
}

// d_pre = d_out * (out > 0) -> split planes [2][n][dp] (pad columns zero): ReLU backward of
// LinearEncoder (category.py:79-80) feeding its weight- and data-gradient GEMMs.
__global__ void relu_bwd_split_kernel(const float* __restrict__ d_out, const float* __restrict__ out,
                                      long long n, int D, int dp, __nv_bfloat16* __restrict__ hi,
                                      __nv_bfloat16* __restrict__ lo) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n * dp;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / dp;
    const int c = (int)(i - r * dp);
    float v = 0.f;
    if (c < D && out[r * D + c] > 0.f) v = d_out[r * D + c];
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
  }
}

// Keep-bit words of ONE dropout site over [R][width] elements (element index r*width + c keys
// the counter RNG, same convention as dropout_words_kernel).
__global__ void dropout_site_words_kernel(unsigned long long seed, uint32_t site, uint32_t thr,
                                          long long R, int width, int mw, uint32_t* __restrict__ words) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < R * mw;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / mw;
    const int w = (int)(i - r * mw);
    words[i] = drop_keep_bits32(seed, site, (unsigned long long)r * (unsigned)width + 32u * (unsigned)w, thr);
  }
}

}  // namespace nrl
