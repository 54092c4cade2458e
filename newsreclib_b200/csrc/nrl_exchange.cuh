// The path's one exchange step fused with the optimizer step, over NVLink / NVSwitch peer memory.
//
// Data-parallel training of the two-tower path (Lightning DDP for the reference,
// configs/trainer/ddp.yaml) ends every step with: sum the gradients over the ranks, then
// torch.optim.Adam on every rank (configs/model/nrms.yaml:49-52).  As library calls that is an
// all-reduce (each gradient byte crosses the links twice and is re-read from HBM by a dense Adam on
// EVERY rank).  Here it is ONE kernel per rank on peer-mapped buffers:
//
//   rank r owns elements [r * per, (r + 1) * per) of the flat parameter / gradient buffers
//   0. (sparse region only) every rank scans its OWN gradient of the embedding table -- the first
//      `sparse_rows` rows of `row_f4` 16-byte granules; a step touches ~15 % of them, the rest of the dense
//      [V+1, E] gradient is zeros -- and writes one bit per row ("this row is non-zero here") into slot [rank]
//      of every rank's bitmap area (peer stores, 9 KB per peer at V = 70 000)
//   1. "ready" barrier: the last CTA of a rank to finish step 0 tells every peer that this rank's gradients
//      (and bitmap) are final (st.release.sys of the step epoch into the peer's flag block), waits for all
//      peers, and publishes ONE decision for the whole grid (go / timed out) in local memory: either every
//      CTA updates its part of the slice or none does
//   2. for the OWNED slice only: g = sum over ranks of grads[rank][i] -- peer loads over NVLink, skipped for
//      rows whose bit says "all zero on that rank", summed in rank order so the result does not depend on
//      timing (adding the skipped zeros would not change a bit) -- Adam on the local moments, the new
//      parameter value stored into EVERY rank's parameter buffer (peer stores), and zeros stored over every
//      gradient element that was read (optimizer.zero_grad() for the next step, on every rank, by the owner)
//   3. "done" barrier: the last CTA of a rank tells every peer that this rank has consumed (and cleared)
//      their gradients and finished writing their parameters, and waits for the same from all peers;
//      the kernel's end is therefore the point where the local replica is complete and the local
//      gradient buffer is zero and may be accumulated into by the next step
//
// = reduce-scatter + sharded Adam + all-gather without intermediate buffers: a NON-ZERO gradient byte crosses
// the links once as a gradient, every parameter byte once as a parameter, and Adam touches n / world elements
// per rank.  All replicas receive the SAME bits (one owner computes each element), which an all-reduce followed
// by per-rank Adam only guarantees if the collective is deterministic.
//
// Flag block (u64 words, lives in the owner's peer-mapped allocation, zeroed once):
//   [0 .. 16)   ready[src]   written by rank src: epoch of the last step whose gradients are final
//   [16 .. 32)  done[src]    written by rank src: epoch of the last step it has finished
//   [32]        error        0, or the first failure (1 = ready wait timed out, 2 = done wait timed out)
//   [33]        cta counter  local: CTAs of the running kernel that have finished their slice
//   [34]        scan counter local: CTAs of the running kernel that have finished step 0
//   [35]        go           local: epoch of the last step whose ready barrier completed
//   [40 .. 45)  timeline     local: %globaltimer of the last launch (start, scan done, go, slice done, done barrier)
// Waits poll with ld.acquire.sys and give up after `timeout_ns` (a peer that died must not hang the
// GPU); the host reads the error word with nrl_exchange_status.  Once the error word is set, later
// launches on this rank return immediately.
#pragma once
#include "nrl_kernels.cuh"

namespace nrl {

constexpr int XCHG_MAX_RANKS = 16;
constexpr int XCHG_READY = 0, XCHG_DONE = 16, XCHG_ERR = 32, XCHG_CTAS = 33, XCHG_SCAN = 34, XCHG_GO = 35,
              XCHG_T0 = 40,  // [40..45): %globaltimer (ns) of the last launch: first CTA in, row scan finished, go, slice
                             // finished by the last CTA, done barrier passed -- a timeline the host can read
              XCHG_FLAG_WORDS = 64;

struct PeerSet {
  int world, rank;
  float* params[XCHG_MAX_RANKS];
  float* grads[XCHG_MAX_RANKS];
  unsigned long long* flags[XCHG_MAX_RANKS];
  unsigned int* bitmaps[XCHG_MAX_RANKS];  // [world][bm_words] per rank (slot s = rank s's rows), or null
};
struct SparseCfg {
  long long rows;   // rows of the sparse region (the embedding table), 0 = everything dense
  int row_f4;       // 16-byte granules per row (E / 4)
  int bm_words;     // (rows + 31) / 32
  int zero_grads;   // store zeros over every gradient element that has been consumed
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
__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, float b1, float b2, float omb1,
                                          float omb2, float eps, float lr_bc1, float sqrt_bc2, float g_scale) {
  const float gi = g * g_scale;
  const float mi = b1 * m + omb1 * gi;
  const float vi = b2 * v + omb2 * gi * gi;
  m = mi;
  v = vi;
  const float denom = sqrtf(vi) / sqrt_bc2 + eps;
  p -= lr_bc1 * (mi / denom);
}

__device__ __forceinline__ unsigned int ld_sys_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ bool f4_nonzero(const float4& g) {
  return (__float_as_uint(g.x) | __float_as_uint(g.y) | __float_as_uint(g.z) | __float_as_uint(g.w)) << 1 != 0u;
}

// W = world size when it is one of the built sizes (all peer loads of an element are then in flight
// together, in registers), 0 = any world size <= 16 (loads issued in rank order, summed as they arrive).
// n4 = elements / 4; m, v are indexed like the parameters (only the owned slice is touched).
template <int W>
__global__ void __launch_bounds__(256)
exchange_adam_kernel(PeerSet ps, float* __restrict__ m, float* __restrict__ v, long long n4,
                     unsigned long long epoch, unsigned long long timeout_ns, float lr, float b1,
                     float b2, float omb1, float omb2, float eps, float bc1, float sqrt_bc2, float g_scale,
                     SparseCfg sp) {
  const int world = W ? W : ps.world, rank = ps.rank;
  unsigned long long* my_flags = ps.flags[rank];
  __shared__ int s_flag;
  const bool sparse = sp.rows > 0 && world > 1;
  const long long sparse_n4 = sparse ? sp.rows * sp.row_f4 : 0;

  // a barrier that timed out once (a peer died) poisons the block: later launches return at once instead of
  // waiting out the timeout again; the host sees the error word through nrl_exchange_status
  if (threadIdx.x == 0) s_flag = ld_acquire_sys_u64(my_flags + XCHG_ERR) == 0ull ? 1 : 0;
  __syncthreads();
  if (!s_flag) return;
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) my_flags[XCHG_T0] = globaltimer_ns();

  // ---- 0. which rows of MY table gradient are non-zero: 16 rows (half a bitmap word) per warp, eight rows' loads in
  // flight per lane, published to every rank as one 16-bit store each
  if (sparse) {
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float4* g_loc = reinterpret_cast<const float4*>(ps.grads[rank]);
    const long long halves = 2ll * sp.bm_words;
    for (long long hw = gw; hw < halves; hw += nwarps) {
      unsigned int bits = 0;
#pragma unroll
      for (int r0 = 0; r0 < 16; r0 += 8) {
        bool any[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          any[u] = false;
          const long long row = hw * 16 + r0 + u;
          if (row < sp.rows)
            for (int c = lane; c < sp.row_f4; c += 32) any[u] |= f4_nonzero(g_loc[row * sp.row_f4 + c]);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (__any_sync(0xffffffffu, any[u])) bits |= 1u << (r0 + u);
      }
      if (lane < world)  // little-endian: half `hw & 1` of word `hw >> 1`
        reinterpret_cast<unsigned short*>(ps.bitmaps[lane] + (long long)rank * sp.bm_words)[hw] = (unsigned short)bits;
    }
  }

  // ---- 1. ready barrier; ONE decision per grid
  __threadfence_system();  // this thread's bitmap stores are visible system-wide before the counter / flags
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long prev = atomicAdd(my_flags + XCHG_SCAN, 1ull);
    s_flag = prev == gridDim.x - 1 ? 1 : 0;
  }
  __syncthreads();
  if (s_flag) {  // last CTA through step 0: every CTA's bitmap words precede the signal
    __threadfence_system();
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
      my_flags[XCHG_SCAN] = 0ull;
      my_flags[XCHG_T0 + 1] = globaltimer_ns();
      s_ok = 1;
    }
    __syncthreads();
    if (threadIdx.x < world && threadIdx.x != rank) {
      st_release_sys_u64(ps.flags[threadIdx.x] + XCHG_READY + rank, epoch);
      if (!wait_flag(my_flags + XCHG_READY + threadIdx.x, epoch, timeout_ns)) s_ok = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      my_flags[XCHG_T0 + 2] = globaltimer_ns();
      if (s_ok) st_release_sys_u64(my_flags + XCHG_GO, epoch);  // local word: the other CTAs poll it
      else atomicCAS(my_flags + XCHG_ERR, 0ull, 1ull);
    }
  }
  if (threadIdx.x == 0) {
    int go = 0;
    for (;;) {
      if (ld_acquire_sys_u64(my_flags + XCHG_GO) >= epoch) { go = 1; break; }
      if (ld_acquire_sys_u64(my_flags + XCHG_ERR) != 0ull) break;
      __nanosleep(32);
    }
    s_flag = go;
  }
  __syncthreads();
  if (!s_flag) return;  // the whole grid gives up together: no parameter of this slice has been touched

  // ---- 2. owned slice: reduce over ranks, Adam, broadcast, clear.  U elements of 16 bytes per thread and
  // iteration so that ~8 peer loads per thread are in flight whatever the world size.
  {
    constexpr int U = W >= 8 ? 1 : W >= 4 ? 2 : W >= 2 ? 4 : 1;
    constexpr int WR = W ? W : 1;
    const long long per = (n4 + world - 1) / world;
    const long long lo = (long long)rank * per;
    const long long hi = lo + per < n4 ? lo + per : n4;
    const float lr_bc1 = lr / bc1;
    const long long nth = (long long)gridDim.x * blockDim.x;
    const unsigned int* bm = sparse ? ps.bitmaps[rank] : nullptr;  // local copy of every rank's row bits
    float4* p_loc = reinterpret_cast<float4*>(ps.params[rank]);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long i0 = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i0 < hi; i0 += U * nth) {
      float4 g4[U];
      if (W) {
        float4 gv[U][WR];
        bool need[U][WR];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const long long i = i0 + u * nth;
          const bool in = i < hi;
          const bool sp_i = in && i < sparse_n4;
          const long long row = sp_i ? i / sp.row_f4 : 0;
#pragma unroll
          for (int r = 0; r < W; ++r) {
            need[u][r] = in && (!sp_i || r == rank ||
                                ((ld_sys_u32(bm + (long long)r * sp.bm_words + (row >> 5)) >> (row & 31)) & 1u));
            gv[u][r] = zero4;
          }
#pragma unroll
          for (int r = 0; r < W; ++r)
            if (need[u][r])
              gv[u][r] = r == rank ? reinterpret_cast<const float4*>(ps.grads[r])[i] : ld_sys_f4(ps.grads[r] + 4 * i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          g4[u] = gv[u][0];
#pragma unroll
          for (int r = 1; r < W; ++r) {
            g4[u].x += gv[u][r].x; g4[u].y += gv[u][r].y; g4[u].z += gv[u][r].z; g4[u].w += gv[u][r].w;
          }
          if (sp.zero_grads) {
            const long long i = i0 + u * nth;
#pragma unroll
            for (int r = 0; r < W; ++r)
              if (need[u][r]) reinterpret_cast<float4*>(ps.grads[r])[i] = zero4;
          }
        }
      } else {
        g4[0] = zero4;
        const bool sp_i = i0 < sparse_n4;
        const long long row = sp_i ? i0 / sp.row_f4 : 0;
        for (int r = 0; r < world; ++r) {
          const bool need = !sp_i || r == rank ||
                            ((ld_sys_u32(bm + (long long)r * sp.bm_words + (row >> 5)) >> (row & 31)) & 1u);
          float4 t = zero4;
          if (need) {
            t = r == rank ? reinterpret_cast<const float4*>(ps.grads[r])[i0] : ld_sys_f4(ps.grads[r] + 4 * i0);
            if (sp.zero_grads) reinterpret_cast<float4*>(ps.grads[r])[i0] = zero4;
          }
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
        adam_elem(p4.x, g4[u].x, m4.x, v4.x, b1, b2, omb1, omb2, eps, lr_bc1, sqrt_bc2, g_scale);
        adam_elem(p4.y, g4[u].y, m4.y, v4.y, b1, b2, omb1, omb2, eps, lr_bc1, sqrt_bc2, g_scale);
        adam_elem(p4.z, g4[u].z, m4.z, v4.z, b1, b2, omb1, omb2, eps, lr_bc1, sqrt_bc2, g_scale);
        adam_elem(p4.w, g4[u].w, m4.w, v4.w, b1, b2, omb1, omb2, eps, lr_bc1, sqrt_bc2, g_scale);
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
    if (threadIdx.x == 0) {
      my_flags[XCHG_CTAS] = 0ull;  // next launch starts from zero
      my_flags[XCHG_T0 + 3] = globaltimer_ns();
    }
    if (threadIdx.x < world && threadIdx.x != rank) {
      st_release_sys_u64(ps.flags[threadIdx.x] + XCHG_DONE + rank, epoch);
      if (!wait_flag(my_flags + XCHG_DONE + threadIdx.x, epoch, timeout_ns))
        atomicCAS(my_flags + XCHG_ERR, 0ull, 2ull);
    }
    __syncthreads();
    if (threadIdx.x == 0) my_flags[XCHG_T0 + 4] = globaltimer_ns();
  }
}

}  // namespace nrl
