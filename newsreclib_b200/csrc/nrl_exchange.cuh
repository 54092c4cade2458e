// The path's one exchange step fused with the optimizer step, over NVLink / NVSwitch peer memory.
//
// Data-parallel training of the two-tower path (Lightning DDP for the reference,
// configs/trainer/ddp.yaml) ends every step with: sum the gradients over the ranks, then
// torch.optim.Adam on every rank (configs/model/nrms.yaml:49-52).  As library calls that is an
// all-reduce (each gradient byte crosses the links twice and is re-read from HBM by a dense Adam on
// EVERY rank).  Here it is ONE kernel per rank on peer-mapped buffers:
//
//   rank r owns elements [r * per, (r + 1) * per) of the flat parameter / gradient buffers
//   1. "ready" barrier: every rank tells every peer that its backward pass has finished
//      (st.release.sys of the step epoch into the peer's flag block), and waits for all peers
//   2. for the OWNED slice only: g = sum over ranks of grads[rank][i] (peer loads over NVLink, summed
//      in rank order so the result does not depend on timing), Adam on the local moments, and the
//      new parameter value is stored into EVERY rank's parameter buffer (peer stores)
//   3. "done" barrier: the last CTA of a rank tells every peer that this rank has consumed their
//      gradients and finished writing their parameters, and waits for the same from all peers;
//      the kernel's end is therefore the point where the local replica is complete and the local
//      gradient buffer may be overwritten by the next step
//
// = reduce-scatter + sharded Adam + all-gather without intermediate buffers: a gradient byte crosses
// the links once as a gradient and once as a parameter, and Adam touches n / world elements per rank.
// All replicas receive the SAME bits (one owner computes each element), which an all-reduce followed
// by per-rank Adam only guarantees if the collective is deterministic.
//
// Flag block (u64 words, lives in the owner's peer-mapped allocation, zeroed once):
//   [0 .. 16)   ready[src]   written by rank src: epoch of the last step whose gradients are final
//   [16 .. 32)  done[src]    written by rank src: epoch of the last step it has finished
//   [32]        error        0, or the first failure (1 = ready wait timed out, 2 = done wait timed out)
//   [33]        cta counter  local: CTAs of the running kernel that have finished their slice
// Waits poll with ld.acquire.sys and give up after `timeout_ns` (a peer that died must not hang the
// GPU); the host reads the error word with nrl_exchange_status.  Once the error word is set, later
// launches on this rank return immediately.
#pragma once
#include "nrl_kernels.cuh"

namespace nrl {

constexpr int XCHG_MAX_RANKS = 16;
constexpr int XCHG_READY = 0, XCHG_DONE = 16, XCHG_ERR = 32, XCHG_CTAS = 33, XCHG_FLAG_WORDS = 64;

struct PeerSet {
  int world, rank;
  float* params[XCHG_MAX_RANKS];
  const float* grads[XCHG_MAX_RANKS];
  unsigned long long* flags[XCHG_MAX_RANKS];
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// peer gradient load: system-coherent (never served from a stale line), 16 bytes
__device__ __forceinline__ float4 ld_sys_f4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// spin until *p >= epoch; false on timeout
__device__ __forceinline__ bool wait_flag(const unsigned long long* p, unsigned long long epoch,
                                          unsigned long long timeout_ns) {
  if (ld_acquire_sys_u64(p) >= epoch) return true;
  const unsigned long long t0 = globaltimer_ns();
  while (ld_acquire_sys_u64(p) < epoch) {
    if (globaltimer_ns() - t0 > timeout_ns) return false;
    __nanosleep(64);
  }
  return true;
}

// torch.optim.Adam on one element (same expression as adam_kernel: the division by sqrt_bc2 is kept)
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float b1, float b2,
                                          float eps, float lr_bc1, float sqrt_bc2, float g_scale) {
  const float gi = g * g_scale;
  const float mi = b1 * m + (1.f - b1) * gi;
  const float vi = b2 * v + (1.f - b2) * gi * gi;
  m = mi;
  v = vi;
  const float denom = sqrtf(vi) / sqrt_bc2 + eps;
  p -= lr_bc1 * (mi / denom);
}

// W = world size when it is one of the built sizes (all peer loads of an element are then in flight
// together, in registers), 0 = any world size <= 16 (loads issued in rank order, summed as they arrive).
// n4 = elements / 4; m, v are indexed like the parameters (only the owned slice is touched).
template <int W>
__global__ void __launch_bounds__(256)
exchange_adam_kernel(PeerSet ps, float* __restrict__ m, float* __restrict__ v, long long n4,
                     unsigned long long epoch, unsigned long long timeout_ns, float lr, float b1,
                     float b2, float eps, float bc1, float sqrt_bc2, float g_scale) {
  const int world = W ? W : ps.world, rank = ps.rank;
  unsigned long long* my_flags = ps.flags[rank];
  __shared__ int s_flag;

  // a barrier that timed out once (a peer died) poisons the block: later launches return at once instead of
  // waiting out the timeout again; the host sees the error word through nrl_exchange_status
  if (threadIdx.x == 0) s_flag = ld_acquire_sys_u64(my_flags + XCHG_ERR) == 0ull ? 1 : 0;
  __syncthreads();
  if (!s_flag) return;
  __syncthreads();  // s_flag (= 1) is reused below as "all peers ready"

  // ---- 1. ready barrier
  if (threadIdx.x < world && threadIdx.x != rank) {
    if (blockIdx.x == 0) {
      __threadfence_system();
      st_release_sys_u64(ps.flags[threadIdx.x] + XCHG_READY + rank, epoch);
    }
    if (!wait_flag(my_flags + XCHG_READY + threadIdx.x, epoch, timeout_ns)) {
      atomicCAS(my_flags + XCHG_ERR, 0ull, 1ull);
      s_flag = 0;
    }
  }
  __syncthreads();
  const bool ready = s_flag != 0;

  // ---- 2. owned slice: reduce over ranks, Adam, broadcast.  U elements of 16 bytes per thread and
  // iteration so that ~8 peer loads per thread are in flight whatever the world size.
  if (ready) {
    constexpr int U = W >= 8 ? 1 : W >= 4 ? 2 : W >= 2 ? 4 : 1;
    const long long per = (n4 + world - 1) / world;
    const long long lo = (long long)rank * per;
    const long long hi = lo + per < n4 ? lo + per : n4;
    const float lr_bc1 = lr / bc1;
    const long long nth = (long long)gridDim.x * blockDim.x;
    float4* p_loc = reinterpret_cast<float4*>(ps.params[rank]);
    for (long long i0 = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < hi; i0 += U * nth) {
      float4 g4[U];
      if (W) {
        float4 gv[U][W ? W : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long i = i0 + u * nth;
#pragma unroll
          for (int r = 0; r < W; ++r) {
            if (i < hi)
              gv[u][r] = r == rank ? __ldg(reinterpret_cast<const float4*>(ps.grads[r]) + i)
                                   : ld_sys_f4(ps.grads[r] + 4 * i);
            else
              gv[u][r] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          g4[u] = gv[u][0];
#pragma unroll
          for (int r = 1; r < W; ++r) {
            g4[u].x += gv[u][r].x; g4[u].y += gv[u][r].y; g4[u].z += gv[u][r].z; g4[u].w += gv[u][r].w;
          }
        }
      } else {
        g4[0] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {
          const float4 t = r == rank ? __ldg(reinterpret_cast<const float4*>(ps.grads[r]) + i0)
                                     : ld_sys_f4(ps.grads[r] + 4 * i0);
          if (r == 0) g4[0] = t;
          else { g4[0].x += t.x; g4[0].y += t.y; g4[0].z += t.z; g4[0].w += t.w; }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + u * nth;
        if (i >= hi) break;
        float4 p4 = p_loc[i];
        float4 m4 = reinterpret_cast<float4*>(m)[i], v4 = reinterpret_cast<float4*>(v)[i];
        adam_elem(p4.x, g4[u].x, m4.x, v4.x, b1, b2, eps, lr_bc1, sqrt_bc2, g_scale);
        adam_elem(p4.y, g4[u].y, m4.y, v4.y, b1, b2, eps, lr_bc1, sqrt_bc2, g_scale);
        adam_elem(p4.z, g4[u].z, m4.z, v4.z, b1, b2, eps, lr_bc1, sqrt_bc2, g_scale);
        adam_elem(p4.w, g4[u].w, m4.w, v4.w, b1, b2, eps, lr_bc1, sqrt_bc2, g_scale);
        reinterpret_cast<float4*>(m)[i] = m4;
        reinterpret_cast<float4*>(v)[i] = v4;
#pragma unroll
        for (int r = 0; r < (W ? W : XCHG_MAX_RANKS); ++r)
          if (r < world) reinterpret_cast<float4*>(ps.params[r])[i] = p4;
      }
    }
  }

  // ---- 3. done barrier: the last CTA of this rank signals and waits
  __threadfence_system();  // this thread's peer stores are visible before anything that follows
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long prev = atomicAdd(my_flags + XCHG_CTAS, 1ull);
    s_flag = prev == gridDim.x - 1 ? 1 : 0;
  }
  __syncthreads();
  if (s_flag) {
    __threadfence_system();  // acquire side of the counter: every CTA's stores precede the signal
    if (threadIdx.x == 0) my_flags[XCHG_CTAS] = 0ull;  // next launch starts from zero
    if (threadIdx.x < world && threadIdx.x != rank) {
      st_release_sys_u64(ps.flags[threadIdx.x] + XCHG_DONE + rank, epoch);
      if (!wait_flag(my_flags + XCHG_DONE + threadIdx.x, epoch, timeout_ns))
        atomicCAS(my_flags + XCHG_ERR, 0ull, 2ull);
    }
  }
}

}  // namespace nrl
