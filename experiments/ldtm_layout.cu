// Microbenchmark (not part of the library): which accumulator elements does each thread receive from
// tcgen05.ld.16x256b (and .32x32b, for reference)?  TMEM is filled through tcgen05.st.32x32b (thread t writes lane t,
// one column per register) with value = lane * 1000 + column, then read back with the shape under test.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ldtm experiments/ldtm_layout.cu && /tmp/ldtm
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__global__ void k(uint32_t* out) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tptr + ((uint32_t)(warp * 32) << 16);
  // fill: lane L (= 32 * warp + lane), columns 0..31
  for (int c = 0; c < 32; ++c) {
    const uint32_t v = (uint32_t)((32 * warp + lane) * 1000 + c);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(base + (uint32_t)c), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t r[8];
  // 16x256b.x2: lanes [0,16) of this warp's quarter, 16 columns
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(base)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) out[(threadIdx.x * 2 + 0) * 8 + i] = r[i];
  // upper 16 lanes of the quarter
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(base + (16u << 16))
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) out[(threadIdx.x * 2 + 1) * 8 + i] = r[i];
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tptr) : "memory");
}

int main() {
  uint32_t* d;
  cudaMalloc(&d, 128 * 2 * 8 * 4);
  k<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  static uint32_t h[128 * 2 * 8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("tcgen05.ld.16x256b.x2: thread -> (lane*1000 + column) of registers r0..r7; second line = taddr lane + 16\n");
  for (int t : {0, 1, 2, 3, 4, 5, 8, 31, 32, 33, 64, 127}) {
    for (int half = 0; half < 2; ++half) {
      printf("t%3d%s:", t, half ? "+16" : "   ");
      for (int i = 0; i < 8; ++i) printf(" %6u", h[(t * 2 + half) * 8 + i]);
      printf("\n");
    }
  }
  return 0;
}
