// Microbenchmark (not part of the library): NVLink peer-memory bandwidth seen by ONE GPU's SMs for the access patterns
// of the fused gradient-exchange kernel (csrc/nrl_exchange.cuh), single process, two GPUs with peer access enabled.
//
// Why: the exchange kernel moves ~240 GB/s per direction (profiles/r01g_exchange.md) against NVLink 5's 900 GB/s.
// This program separates the candidates: the load flavour (ld.relaxed.sys as used today, ld.global.cg, plain ld.global),
// the number of 16-byte loads in flight per thread, the grid (CTAs per SM), and read-only / write-only / read+write.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o p2p_bw experiments/p2p_bw.cu && ./p2p_bw      (needs 2 GPUs)
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      printf("%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__);                \
      exit(1);                                                                                  \
    }                                                                                           \
  } while (0)

enum { LD_SYS = 0, LD_CG = 1, LD_PLAIN = 2 };

template <int FLAVOUR>
__device__ __forceinline__ float4 ld16(const float4* p) {
  float4 v;
  if (FLAVOUR == LD_SYS)
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  else if (FLAVOUR == LD_CG)
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  else
    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// mode bit 0: read `src` (peer), bit 1: write `dst` (peer).  U independent 16-byte accesses per thread and iteration.
template <int FLAVOUR, int U>
__global__ void __launch_bounds__(256) p2p_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                                                  float4* __restrict__ sink, long long n4, int mode) {
  const long long nth = (long long)gridDim.x * blockDim.x;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < n4; i0 += U * nth) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * nth;
      v[u] = make_float4(1.f, 2.f, 3.f, 4.f);
      if ((mode & 1) && i < n4) v[u] = ld16<FLAVOUR>(src + i);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i = i0 + u * nth;
      if ((mode & 2) && i < n4) dst[i] = v[u];
      acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
    }
  }
  if (acc.x == -1.f) sink[0] = acc;  // keeps the loads alive
}

template <int FLAVOUR, int U>
static float run(const float4* src, float4* dst, float4* sink, long long n4, int mode, int ctas) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  p2p_kernel<FLAVOUR, U><<<ctas, 256>>>(src, dst, sink, n4, mode);  // warm-up
  CK(cudaEventRecord(e0));
  for (int r = 0; r < 5; ++r) p2p_kernel<FLAVOUR, U><<<ctas, 256>>>(src, dst, sink, n4, mode);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  return ms / 5.f;
}

int main() {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 2) { printf("needs 2 GPUs (found %d)\n", ndev); return 0; }
  int can = 0;
  CK(cudaDeviceCanAccessPeer(&can, 0, 1));
  if (!can) { printf("GPU 0 cannot access GPU 1\n"); return 1; }
  const long long bytes = 64ll << 20, n4 = bytes / 16;  // 64 MB: what one rank pulls at W = 4
  float4 *peer_src, *peer_dst, *sink;
  CK(cudaSetDevice(1));
  CK(cudaMalloc(&peer_src, bytes));
  CK(cudaMalloc(&peer_dst, bytes));
  CK(cudaMemset(peer_src, 0, bytes));
  CK(cudaDeviceSynchronize());
  CK(cudaSetDevice(0));
  CK(cudaDeviceEnablePeerAccess(1, 0));
  CK(cudaMalloc(&sink, 64));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("%s x2, %d SMs, 64 MB per direction; GB/s per direction\n", prop.name, sms);
  const char* mode_name[4] = {"", "read ", "write", "rd+wr"};
  for (int mode = 1; mode <= 3; ++mode) {
    printf("\n%s   CTAs/SM |  sys U=4   sys U=8  sys U=16 |   cg U=4    cg U=8   cg U=16 | plain U=8\n", mode_name[mode]);
    for (int per_sm : {1, 2, 4, 8}) {
      const int ctas = per_sm * sms;
      auto gbs = [&](float ms) { return bytes / (ms * 1e-3) / 1e9; };
      printf("        %6d | %8.0f  %8.0f  %8.0f | %8.0f  %8.0f  %8.0f | %8.0f\n", per_sm,
             gbs(run<LD_SYS, 4>(peer_src, peer_dst, sink, n4, mode, ctas)), gbs(run<LD_SYS, 8>(peer_src, peer_dst, sink, n4, mode, ctas)),
             gbs(run<LD_SYS, 16>(peer_src, peer_dst, sink, n4, mode, ctas)), gbs(run<LD_CG, 4>(peer_src, peer_dst, sink, n4, mode, ctas)),
             gbs(run<LD_CG, 8>(peer_src, peer_dst, sink, n4, mode, ctas)), gbs(run<LD_CG, 16>(peer_src, peer_dst, sink, n4, mode, ctas)),
             gbs(run<LD_PLAIN, 8>(peer_src, peer_dst, sink, n4, mode, ctas)));
    }
  }
  printf("\n(the exchange kernel today: sys loads, 8 in flight per thread, 4 CTAs/SM, reads and writes in one loop)\n");
  return 0;
}
