// Microbenchmark (not part of the library): issue rate and latency of the warp-level tensor instruction the short-sequence
// attention kernels are built on (mma.sync.m16n8k16 bf16 -> fp32, csrc/nrl_attn_mma.cuh) and of movmatrix, on sm_100a.
//
// Why: attn_bwd_mma_kernel issues 204 HMMAs + 72 movmatrix per (news, head) problem and runs at 0.13 instructions per
// cycle and warp (ncu, profiles/r01f_ncu_memory_kernels.md).  Whether that is the HMMA pipe (then: fewer passes / fewer
// products), dependent-chain latency (then: more independent accumulators in flight) or memory decides what to change
// (DESIGN.md section 10, item 2).  This program prints, for ILP = 1, 2, 4, 8 independent accumulators per warp and 1..16
// warps per SM, the measured cycles per HMMA per SM sub-partition.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_rate experiments/hmma_rate.cu && ./hmma_rate
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movm_t(uint32_t x) {
  uint32_t y;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

template <int ILP>
__global__ void hmma_kernel(int iters, float* out, long long* cycles) {
  float c[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  const uint32_t a0 = 0x3f803f80u + threadIdx.x, a1 = a0 ^ 0x00010001u, a2 = a0 + 3, a3 = a0 + 5, b0 = a0 + 7, b1 = a0 + 11;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) mma_16816(c[i], a0, a1, a2, a3, b0, b1);
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void movm_kernel(int iters, uint32_t* out, long long* cycles) {
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 2654435761u + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = movm_t(x[i]);
  }
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ILP>
static double run_hmma(int warps_per_sm, int sms, int iters, float* out, long long* cyc) {
  hmma_kernel<ILP><<<sms, 32 * warps_per_sm>>>(iters, out, cyc);
  cudaDeviceSynchronize();
  std::vector<long long> h(sms);
  cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (long long v : h) mean += (double)v;
  mean /= sms;
  // HMMAs issued per SM sub-partition: warps_per_sm / 4 warps (rounded up) each iters * ILP
  const double per_smsp = (double)((warps_per_sm + 3) / 4) * iters * ILP;
  return mean / per_smsp;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount, iters = 4096;
  float* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  printf("%s, %d SMs; cycles per mma.sync.m16n8k16 (bf16 -> f32) per SM sub-partition\n", prop.name, sms);
  printf("warps/SM   ILP=1    ILP=2    ILP=4    ILP=8\n");
  for (int w : {1, 4, 8, 12, 16, 32}) {
    run_hmma<1>(w, sms, 64, out, cyc);  // warm-up
    printf("%8d %7.2f  %7.2f  %7.2f  %7.2f\n", w, run_hmma<1>(w, sms, iters, out, cyc), run_hmma<2>(w, sms, iters, out, cyc),
           run_hmma<4>(w, sms, iters, out, cyc), run_hmma<8>(w, sms, iters, out, cyc));
  }
  printf("(ILP=1, 1 warp = dependent-chain latency; the floor of a row = issue interval of the pipe.\n"
         " One m16n8k16 = 4096 flop: an interval of t cycles = %d SMs x 4 x 4096 / t flop per clock chip-wide.)\n", sms);
  for (int w : {1, 4, 16}) {
    movm_kernel<<<sms, 32 * w>>>(iters, reinterpret_cast<uint32_t*>(out), cyc);
    cudaDeviceSynchronize();
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (long long v : h) mean += (double)v;
    printf("movmatrix.trans, %2d warps/SM: %.2f cycles per instruction per sub-partition (8 independent per warp)\n", w,
           mean / sms / ((double)((w + 3) / 4) * iters * 8));
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : 1;
}
