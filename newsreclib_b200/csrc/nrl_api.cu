// C-ABI entry points (include/nrl.h) and the host-side orchestration of the NRMS hot path.
// Everything below the ABI is CUDA for sm_100a; there is no CPU or library fallback.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cudaTypedefs.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/nrl.h"
#include "nrl_gemm.cuh"
#include "nrl_kernels.cuh"
#include "nrl_attn_mma.cuh"
#include "nrl_naml.cuh"
#include "nrl_exchange.cuh"
#include "nrl_tfm.cuh"
#include "nrl_attn_flash.cuh"
#include "nrl_attn_title.cuh"

using namespace nrl;
typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------------------
// error / bookkeeping
// ----------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      return fail(NRL_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                  __FILE__, __LINE__);                                                     \
  } while (0)
// Optional per-launch timing (bench.py's roofline leg): when enabled, one CUDA event is
// recorded on the profiled stream after every launch; on an in-order stream the duration of
// launch i is end_i - end_{i-1}.
struct Profiler {
  bool on = false;
  cudaStream_t stream = nullptr;
  std::vector<cudaEvent_t> pool;
  std::vector<const char*> names;
  size_t used = 0;
};
static Profiler g_prof;
static void prof_mark(const char* name) {
  if (!g_prof.on) return;
  if (g_prof.used == g_prof.pool.size()) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    g_prof.pool.push_back(e);
  }
  cudaEventRecord(g_prof.pool[g_prof.used++], g_prof.stream);
  g_prof.names.push_back(name);
}
#define LAUNCH_CHECK(name)                                                                 \
  do {                                                                                     \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                    \
    cudaError_t _e = cudaPeekAtLastError();                                                \
    if (_e != cudaSuccess)                                                                 \
      return fail(NRL_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e));  \
    prof_mark(name);                                                                       \
  } while (0)
#define TRY(expr)                 \
  do {                            \
    int _s = (expr);              \
    if (_s != NRL_OK) return _s;  \
  } while (0)

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline int grid_for(long long work, int per_block, int cap) {
  long long g = (work + per_block - 1) / per_block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

struct Device {
  int sm_count = 0;
  bool ok = false;
  PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
};
static Device g_dev;
static std::mutex g_dev_mu;

static int device_init() {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (g_dev.ok) return NRL_OK;
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10)
    return fail(NRL_ERR_UNSUPPORTED, "newsreclib_b200 needs an sm_100 GPU, found sm_%d%d",
                prop.major, prop.minor);
  g_dev.sm_count = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  if (!fn || qres != cudaDriverEntryPointSuccess)
    return fail(NRL_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
  g_dev.encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  CUDA_TRY(cudaFuncSetAttribute(nrl_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                GEMM_SMEM_LIMIT));
  CUDA_TRY(cudaFuncSetAttribute(nrl_gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                GEMM_SMEM_LIMIT));
  CUDA_TRY(cudaFuncSetAttribute(nrl_gemm_tc2a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                GEMM_SMEM_LIMIT));
  g_dev.ok = true;
  return NRL_OK;
}
static int attn_attrs_init();

// ----------------------------------------------------------------------------------------
// workspace carving (same sequence in ws_bytes / fwd / bwd -> same offsets)
// ----------------------------------------------------------------------------------------
struct Bump {
  char* base;
  size_t off = 0;
  explicit Bump(void* b) : base(static_cast<char*>(b)) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 1023) & ~size_t(1023);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct Dims {
  int E, H, Q, DH, Ep, Qp, P3, MW, LDQ;
};
static int make_dims(nrl_dims d, Dims& o) {
  if (d.embed_dim <= 0 || d.num_heads <= 0 || d.query_dim <= 0 || d.embed_dim % d.num_heads)
    return fail(NRL_ERR_INVALID_ARG, "bad dims E=%d heads=%d Q=%d", d.embed_dim, d.num_heads,
                d.query_dim);
  o.E = d.embed_dim; o.H = d.num_heads; o.Q = d.query_dim; o.DH = o.E / o.H;
  if (o.DH != 16 && o.DH != 20 && o.DH != 32 && o.DH != 48 && o.DH != 64)
    return fail(NRL_ERR_UNSUPPORTED, "head dim %d not built (16, 20, 32, 48, 64)", o.DH);
  if (o.Q > 256) return fail(NRL_ERR_UNSUPPORTED, "query_dim %d > 256", o.Q);
  o.Ep = round_up(o.E + 1, 16);
  o.Qp = round_up(o.Q, 16);
  o.P3 = round_up(3 * o.E, 16);
  o.MW = (o.E + 31) / 32;
  // fp32 qkv rows start on 128-byte lines: every 32-column TMA store row is one full line
  o.LDQ = round_up(3 * o.E, 32);
  return NRL_OK;
}

struct BlockWs {
  bf16 *win_f, *win_t, *wout_f, *wout_t, *wadd_f, *wadd_t;
  bf16* x;     // [2][R][Ep]   input planes (ones column at E)
  float* qkv;  // [R][3E]
  float* lse;  // [R][H]
  float* delta;  // [R][H]      dO . O per row and head (backward of the long-sequence attention kernels)
  bf16* o;     // [2][R][Ep]
  float* y;    // [R][E]       MHSA output (after dropout site 1)
  bf16* yp;    // [2][R][Ep]
  float* a;    // [R][Q]       tanh(yW+b)
  float *s, *w;  // [R]
  float* dy1;  // [R][E]
  bf16* dap;   // [2][R][Qp]
  bf16* dyp;   // [2][R][Ep]
  float* d_o;  // [R][E]
  bf16* dqkv;  // [2][R][P3]
  float* dx;   // [R][E]
  uint32_t *mask0, *mask1;  // [R][MW] dropout keep-bit words of the two sites
};
static void carve_block(Bump& b, long long R, const Dims& d, BlockWs& w) {
  w.win_f = b.take<bf16>(2ull * 3 * d.E * d.Ep);
  w.win_t = b.take<bf16>(2ull * d.E * d.P3);
  w.wout_f = b.take<bf16>(2ull * d.E * d.Ep);
  w.wout_t = b.take<bf16>(2ull * d.E * d.Ep);
  w.wadd_f = b.take<bf16>(2ull * d.Q * d.Ep);
  w.wadd_t = b.take<bf16>(2ull * d.E * d.Qp);
  w.x = b.take<bf16>(2ull * R * d.Ep);
  w.qkv = b.take<float>((size_t)R * d.LDQ);
  w.lse = b.take<float>((size_t)R * d.H);
  w.delta = b.take<float>((size_t)R * d.H);
  w.o = b.take<bf16>(2ull * R * d.Ep);
  w.y = b.take<float>((size_t)R * d.E);
  w.yp = b.take<bf16>(2ull * R * d.Ep);
  w.a = b.take<float>((size_t)R * d.Q);
  w.s = b.take<float>((size_t)R);
  w.w = b.take<float>((size_t)R);
  w.dy1 = b.take<float>((size_t)R * d.E);
  w.dap = b.take<bf16>(2ull * R * d.Qp);
  w.dyp = b.take<bf16>(2ull * R * d.Ep);
  w.d_o = b.take<float>((size_t)R * d.E);
  w.dqkv = b.take<bf16>(2ull * R * d.P3);
  w.dx = b.take<float>((size_t)R * d.E);
  w.mask0 = b.take<uint32_t>((size_t)R * d.MW);
  w.mask1 = b.take<uint32_t>((size_t)R * d.MW);
}

// ----------------------------------------------------------------------------------------
// tensor-core GEMM launcher
// ----------------------------------------------------------------------------------------
struct Ctx {
  cudaStream_t stream;
  int precision;
  bool two_planes() const { return precision == NRL_PREC_BF16X3; }
};

static int make_tmap(CUtensorMap* m, const bf16* base, unsigned long long inner,
                     unsigned long long rows, unsigned long long pitch, unsigned box_inner,
                     unsigned box_rows) {
  cuuint64_t gdim[3] = {inner, rows, 2};
  cuuint64_t gstr[2] = {pitch * sizeof(bf16), rows * pitch * sizeof(bf16)};
  cuuint32_t box[3] = {box_inner, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_dev.encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(base), gdim,
                            gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(NRL_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): inner=%llu rows=%llu pitch=%llu",
                (int)r, inner, rows, pitch);
  return NRL_OK;
}

// fp32 sink [rows, cols] with row pitch ld (floats): 32 x 32 boxes, SWIZZLE_128B staging
static int make_tmap_f32(CUtensorMap* m, const float* base, unsigned long long cols,
                         unsigned long long rows, unsigned long long ld) {
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld & 3))
    return fail(NRL_ERR_INVALID_ARG, "fp32 GEMM sink must be 16-byte aligned with a pitch that is a multiple of 4 floats");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_dev.encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(NRL_ERR_CUDA, "cuTensorMapEncodeTiled(f32 sink) failed (%d): cols=%llu rows=%llu ld=%llu",
                (int)r, cols, rows, ld);
  return NRL_OK;
}
// fp32 operand boxes of the staged attention kernels: element (col, s, b) lives at
// base + (s * seq_stride + b * batch_stride) * ld + col; box = [1][S][box_cols], no swizzle.
static int make_tmap_attn(CUtensorMap* m, const float* base, unsigned long long cols, unsigned long long ld,
                          unsigned long long S, long long seq_stride, unsigned long long NB, long long batch_stride,
                          unsigned box_cols) {
  cuuint64_t gdim[3] = {cols, S, NB};
  cuuint64_t gstr[2] = {(cuuint64_t)seq_stride * ld * sizeof(float), (cuuint64_t)batch_stride * ld * sizeof(float)};
  cuuint32_t box[3] = {box_cols, (cuuint32_t)S, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_dev.encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(NRL_ERR_CUDA, "cuTensorMapEncodeTiled(attention operand) failed (%d): cols=%llu ld=%llu S=%llu NB=%llu",
                (int)r, cols, ld, S, NB);
  return NRL_OK;
}

// split-plane bf16 sink [planes][rows][pitch]: 32 x 32 x planes boxes, SWIZZLE_64B staging
static int make_tmap_planes(CUtensorMap* m, const bf16* base, unsigned long long cols,
                            unsigned long long rows, unsigned long long pitch, int planes) {
  cuuint64_t gdim[3] = {cols, rows, (cuuint64_t)planes};
  cuuint64_t gstr[2] = {pitch * sizeof(bf16), rows * pitch * sizeof(bf16)};
  cuuint32_t box[3] = {32, 32, (cuuint32_t)planes};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_dev.encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(base), gdim, gstr,
                            box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(NRL_ERR_CUDA, "cuTensorMapEncodeTiled(plane sink) failed (%d): cols=%llu rows=%llu pitch=%llu",
                (int)r, cols, rows, pitch);
  return NRL_OK;
}

// Host-side description of where a GEMM's result goes (turned into tensor maps + GemmEpi).
struct Sinks {
  float* f32 = nullptr; long long ld_f32 = 0; int f32_cols = 0; bool reduce = false;  // fp32 [M, f32_cols]
  bf16* sp = nullptr; long long ld_sp = 0; int sp_cols = 0; int ones_col = -1;       // planes [2][M, sp_cols]
};

static int launch_gemm(const Ctx& c, GemmParams& p, const CUtensorMap& ta, const CUtensorMap& tb,
                       const Sinks& sk, const char* name) {
  CUtensorMap tout, tsp;
  memset(&tout, 0, sizeof(tout));
  memset(&tsp, 0, sizeof(tsp));
  p.epi.f32_sink = 0; p.epi.sp_sink = 0;
  if (sk.f32) {
    TRY(make_tmap_f32(&tout, sk.f32, sk.f32_cols, p.M, sk.ld_f32));
    p.epi.f32_sink = sk.reduce ? 2 : 1;
    p.epi.f32_cols = sk.f32_cols;
  }
  if (sk.sp) {
    TRY(make_tmap_planes(&tsp, sk.sp, sk.sp_cols, p.M, sk.ld_sp, c.two_planes() ? 2 : 1));
    p.epi.sp_sink = 1;
    p.epi.sp_cols = sk.sp_cols;
    p.epi.sp_two = c.two_planes() ? 1 : 0;
    p.epi.ones_col = sk.ones_col;
  }
  p.epi_buf_bytes = (sk.f32 && sk.sp) ? 8192 : 4096;
  p.epi_bufs = 1;
  // register-direct fp32 sink (NRL_EPI_DIRECT=1).  Measured on B200 (profiles/r02_gemm_epilogue.md): the same time as
  // the TMA-store epilogue (in-proj 0.181 ms either way) -- what the epilogue costs is the output's trip through L2, which
  // these GEMMs already saturate with operand fills, not the way the bytes leave the SM.  Kept as an option: it needs no
  // staging buffers, which an operand-stationary main loop could use.
  static const bool direct_on = [] { const char* e = getenv("NRL_EPI_DIRECT"); return e && e[0] == '1'; }();
  const GemmEpi& ee = p.epi;
  p.direct = (direct_on && sk.f32 && !sk.reduce && !sk.sp && !ee.add_w && !ee.relu && !ee.pos_mask && !ee.pos_words && !ee.qvec && !ee.gb && !ee.gelu &&
              !ee.add_mat && !ee.gelu_pre &&
              !(sk.f32_cols & 1) && !(sk.ld_f32 & 1) && !(reinterpret_cast<uintptr_t>(sk.f32) & 7)) ? 1 : 0;
  p.out = sk.f32; p.ld_out = sk.ld_f32;
  static const int epi_dbg = [] { const char* e = getenv("NRL_EPI_DEBUG"); return e ? atoi(e) : 0; }();
  p.dbg = epi_dbg;  // timing experiments only: results are wrong when != 0
  if (p.astat) {  // A-stationary pair kernel: the unit's A panel resident, B halves through a ring, register-direct epilogue
    const int kb_total = (p.K + GEMM_BK - 1) / GEMM_BK;
    const int a_bytes = kb_total * p.planes * GEMM_A_BYTES;
    const int b_stage = p.planes * (p.BN / 2) * 128;
    int bst = (GEMM_SMEM_LIMIT - 1024 - GEMM_BAR_BYTES - a_bytes) / b_stage;
    if (bst > GEMM_B_STAGES_MAX) bst = GEMM_B_STAGES_MAX;
    if (bst < 2) return fail(NRL_ERR_UNSUPPORTED, "A-stationary GEMM does not fit shared memory");
    p.direct = 1;
    p.stages = bst;
    const int smem3 = 1024 + a_bytes + bst * b_stage + GEMM_BAR_BYTES;
    const int m_pairs = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
    const int clusters = m_pairs < g_dev.sm_count / 2 ? m_pairs : g_dev.sm_count / 2;
    nrl_gemm_tc2a_kernel<<<2 * clusters, GEMM_THREADS, smem3, c.stream>>>(ta, tb, tout, tsp, p);
    LAUNCH_CHECK(name);
    return NRL_OK;
  }
  if (p.pair) {  // CTA pairs (cta_group::2): each CTA stages its 128 rows of A and half of the B tile
    int half_b = p.mn_major ? (p.BN / 2 + 63) / 64 * 8192 : p.BN / 2 * 128;
    if (p.fuse_n) half_b += (((p.n_extent - p.BN + 15) & ~15) / 2 + 63) / 64 * 8192;  // boxes of the second n-tile
    const int stage_bytes2 = p.planes * (GEMM_A_BYTES + half_b);
    int stages2 = (GEMM_SMEM_LIMIT - 1024 - GEMM_BAR_BYTES - GEMM_EPI_WARPS * p.epi_bufs * p.epi_buf_bytes) / stage_bytes2;
    if (stages2 > GEMM_MAX_STAGES) stages2 = GEMM_MAX_STAGES;
    if (stages2 < 2) return fail(NRL_ERR_UNSUPPORTED, "pair GEMM tile does not fit shared memory");
    p.stages = stages2;
    p.tiles_per_unit = (p.n_extent + p.BN - 1) / p.BN;
    const int smem2 = 1024 + stages2 * stage_bytes2 + GEMM_EPI_WARPS * p.epi_bufs * p.epi_buf_bytes + GEMM_BAR_BYTES;
    const int m_pairs = (p.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
    int units = m_pairs * p.k_splits;
    {  // NT GEMMs: hand out single tiles instead of whole row pairs when that saves >= 4 % of the tile rounds (the row
       // pairs of a wave share their B tiles either way; NRL_GEMM_FINE=0: always whole row pairs)
      static const bool fine_on = [] { const char* e = getenv("NRL_GEMM_FINE"); return !(e && e[0] == '0'); }();
      const int nt = p.tiles_per_unit, cl = g_dev.sm_count / 2;
      const long long rounds_unit = (long long)((m_pairs + cl - 1) / cl) * nt;
      const long long rounds_tile = ((long long)m_pairs * nt + cl - 1) / cl;
      if (fine_on && !p.mn_major && !p.fuse_n && p.k_splits == 1 && nt > 1 && rounds_tile * 100 <= rounds_unit * 96) {
        p.tiles_per_unit = 1;
        units = m_pairs * nt;
      }
    }
    const int clusters = units < g_dev.sm_count / 2 ? units : g_dev.sm_count / 2;
    nrl_gemm_tc2_kernel<<<2 * clusters, GEMM_THREADS, smem2, c.stream>>>(ta, tb, tout, tsp, p);
    LAUNCH_CHECK(name);
    return NRL_OK;
  }
  const int stage_bytes = p.planes * (GEMM_A_BYTES + (p.mn_major ? (p.BN + 63) / 64 * 8192 : p.BN * 128));
  int stages = (GEMM_SMEM_LIMIT - 1024 - GEMM_BAR_BYTES - GEMM_EPI_WARPS * p.epi_bufs * p.epi_buf_bytes) / stage_bytes;
  if (stages > GEMM_MAX_STAGES) stages = GEMM_MAX_STAGES;
  if (stages < 2) return fail(NRL_ERR_UNSUPPORTED, "GEMM tile does not fit shared memory");
  p.stages = stages;
  const int smem = 1024 + stages * stage_bytes + GEMM_EPI_WARPS * p.epi_bufs * p.epi_buf_bytes + GEMM_BAR_BYTES;
  const int m_tiles = (p.M + GEMM_BM - 1) / GEMM_BM, n_tiles = (p.n_extent + p.BN - 1) / p.BN;
  const int tiles = m_tiles * n_tiles * p.k_splits;
  // unit = whole m-block when there are enough m-blocks to fill the machine twice over
  p.tiles_per_unit = (p.k_splits == 1 && n_tiles > 1 && m_tiles >= 2 * g_dev.sm_count) ? n_tiles : 1;
  const int units = tiles / p.tiles_per_unit;
  const int grid = units < g_dev.sm_count ? units : g_dev.sm_count;
  // programmatic dependent launch: the kernel's prologue (barrier init, TMEM allocation, tensor-map prefetch) and its
  // launch latency overlap the tail of whatever precedes it in the stream; it waits (griddepcontrol.wait) before its
  // first global access.  Meant for the chains of small dependent launches (the user block).  Measured on B200
  // (profiles/r02pdl_{1,0}.json, two runs each): 2.061 / 2.062 ms per step with it, 2.064 / 2.074 without -- inside the
  // run-to-run spread, so it stays OFF by default (NRL_PDL=1 enables it; the GPU suite passes either way).
  static const bool pdl_on = [] { const char* e = getenv("NRL_PDL"); return e && e[0] == '1'; }();
  if (pdl_on && !p.mn_major) {  // the NT GEMMs of the main stream; the weight-gradient (TN) GEMMs follow event waits on the side stream
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = c.stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, nrl_gemm_tc_kernel, ta, tb, tout, tsp, p));
  } else {
    nrl_gemm_tc_kernel<<<grid, GEMM_THREADS, smem, c.stream>>>(ta, tb, tout, tsp, p);
  }
  LAUNCH_CHECK(name);
  return NRL_OK;
}

static void set_segs(const Ctx& c, GemmParams& p) { p.planes = c.two_planes() ? 2 : 1; }

// balanced n-tiles: as few tiles as possible (<= 256 columns each), all the same width, and
// narrow enough that at least two pipeline stages fit beside the epilogue staging buffers
static int balanced_bn(int n_extent, int planes, bool both_sinks, int mn_major, bool single_tile = false,
                       bool pair = false, bool small = false) {
  const int avail = GEMM_SMEM_LIMIT - 1024 - GEMM_BAR_BYTES - GEMM_EPI_WARPS * 1 * (both_sinks ? 8192 : 4096);
  static const int bn_max = [] { const char* e = getenv("NRL_GEMM_BN_MAX"); return e ? atoi(e) : 256; }();
  // small = a launch that cannot fill the machine with 256-wide tiles (the user block's 3 200 rows): it is bound by the
  // latency of ONE tile, and a 64-wide tile keeps all of K = 304 in flight (4-5 stages instead of 2) and spreads the work
  // over more SMs.  Measured (experiments/small_gemm_latency.py, profiles/r02y_small_gemm_latency.txt): 3 200 x 300 x 304
  // 15.4 -> 11.3 us, 3 200 x 200 x 304 17.4 -> 11.2 us, 3 200 x 900 x 304 19.5 -> 17.4 us; large launches unchanged.
  const int cap = single_tile ? 256 : (small && bn_max > 64 ? 64 : bn_max);
  for (int nt = (n_extent + cap - 1) / cap;; ++nt) {
    // multiple of 32: the epilogue stores 32-column boxes, which must not straddle two n-tiles
    // (the overhang of the LAST tile lies outside the tensor and is clipped by TMA)
    const int bn = round_up((n_extent + nt - 1) / nt, 32);
    const int stage = planes * (GEMM_A_BYTES + (mn_major ? (bn + 63) / 64 * 8192 : (pair ? bn / 2 : bn) * 128));
    if (avail / stage >= 2 || bn <= 32) return bn;
  }
}

// K-major ("NT") GEMM: D[M,N] = A[M,K] * B[N,K]^T.  a/b point at plane 0; plane 1 follows.
static int gemm_nt(const Ctx& c, const bf16* A, long long M, int a_pitch, const bf16* B, int N,
                   int b_pitch, int K, const GemmEpi& epi, const Sinks& sk, const char* name) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = (int)M; p.N = N; p.K = K;
  p.n_extent = (sk.sp && sk.sp_cols > N) ? sk.sp_cols : N;
  p.mn_major = 0;
  set_segs(c, p);
  // big row-streaming GEMMs run on CTA pairs (NRL_GEMM_PAIR=0 keeps the 1-CTA kernel: A/B runs)
  static const bool pair_on = [] { const char* e = getenv("NRL_GEMM_PAIR"); return !(e && e[0] == '0'); }();
  p.pair = (pair_on && (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM) >= g_dev.sm_count / 2) ? 1 : 0;
  static const bool small_on = [] { const char* e = getenv("NRL_GEMM_SMALL_BN"); return !(e && e[0] == '0'); }();
  const bool small = small_on && !p.pair &&
                     (M + GEMM_BM - 1) / GEMM_BM * ((p.n_extent + 255) / 256) < (long long)g_dev.sm_count;
  p.BN = balanced_bn(p.n_extent, p.planes, sk.f32 && sk.sp, 0, epi.score != nullptr, p.pair != 0, small);
  // A-stationary pair kernel (nrl_gemm_tc2a_kernel) for plain / dropout fp32-sink GEMMs whose A panel fits: the n-tile is
  // narrowed until THREE B stages fit beside the panel (measured: 240-wide tiles with a 2-deep ring 0.183 ms, 160-wide
  // with a 3-deep ring 0.175 ms for the in-projection, against 0.189 ms for the streaming kernel; NRL_GEMM_ASTAT=0: off)
  static const bool astat_on = [] { const char* e = getenv("NRL_GEMM_ASTAT"); return !(e && e[0] == '0'); }();
  p.astat = 0;
  if (p.pair && astat_on && sk.f32 && !sk.reduce && !sk.sp && !epi.add_w && !epi.relu && !epi.pos_mask && !epi.qvec &&
      !epi.gb && !epi.gelu && !epi.add_mat && !epi.gelu_pre && !(sk.f32_cols & 1) && !(sk.ld_f32 & 1) && !(reinterpret_cast<uintptr_t>(sk.f32) & 7)) {
    const int kb_total = (K + GEMM_BK - 1) / GEMM_BK;
    const int avail = GEMM_SMEM_LIMIT - 1024 - GEMM_BAR_BYTES - kb_total * p.planes * GEMM_A_BYTES;
    int bn = avail > 0 ? (avail / (3 * p.planes * 64)) & ~31 : 0;
    if (bn > 256) bn = 256;
    if (kb_total <= GEMM_A_SLOTS && bn >= 64 && p.n_extent > bn) {
      const int nt = (p.n_extent + bn - 1) / bn;
      p.BN = round_up((p.n_extent + nt - 1) / nt, 32);
      p.astat = 1;
    }
  }
  if (epi.score && p.BN < N) return fail(NRL_ERR_UNSUPPORTED, "score fusion needs a single n-tile (N <= 256)");
  if (epi.score) CUDA_TRY(cudaMemsetAsync(epi.score, 0, (size_t)M * sizeof(float), c.stream));
  p.k_splits = 1;
  p.epi = epi;
  CUtensorMap ta, tb;
  TRY(make_tmap(&ta, A, K, M, a_pitch, GEMM_BK, GEMM_BM));
  TRY(make_tmap(&tb, B, K, N, b_pitch, GEMM_BK, p.pair ? p.BN / 2 : p.BN));
  return launch_gemm(c, p, ta, tb, sk, name);
}

// MN-major ("TN") weight-gradient GEMM: G[M,N] += sum_r A[r, m] * B[r, n], r < R (split-K,
// partial tiles are TMA-reduced into G).  Accumulator column gb_col (the ones column of B)
// goes to the bias gradient gb[m].
static int gemm_tn(const Ctx& c, const bf16* A, int M, int a_pitch, const bf16* B, int N,
                   int b_pitch, long long R, float* gw, long long ld_gw, int gw_cols, float* gb,
                   const char* name) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = (int)R;
  p.n_extent = N;
  p.mn_major = 1;
  set_segs(c, p);
  const int kb_total = (int)((R + GEMM_BK - 1) / GEMM_BK);
  // CTA pairs when the 256-row pairs waste no more MMA rows than the 128-row tiles would (M = 900, 200,
  // 400 ...; not M = 300) -- doubles the operand reuse per byte pulled from L2
  static const bool pair_on = [] { const char* e = getenv("NRL_GEMM_PAIR"); return !(e && e[0] == '0'); }();
  const int m_tiles = (M + GEMM_BM - 1) / GEMM_BM, m_pairs = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
  p.pair = (pair_on && 2 * m_pairs == m_tiles && kb_total >= g_dev.sm_count) ? 1 : 0;
  int ks;
  if (p.pair) {
    p.BN = N > 128 ? 256 : 128;  // each CTA's half is whole 64-column boxes (the last tile may be narrower)
    // N = BN + a narrow rest (304 = 256 + 48): both n-tiles accumulate in one k-loop (two TMEM regions), so the
    // A operand -- the big one, [R, M] -- is streamed from DRAM once instead of once per n-tile
    p.fuse_n = (N > p.BN && N <= 2 * p.BN && p.BN == 256) ? 1 : 0;
    ks = (g_dev.sm_count / 2) / m_pairs;
  } else {
    p.BN = balanced_bn(N, p.planes, false, 1);
    ks = g_dev.sm_count / m_tiles;  // every n-block's tiles fill the machine once
  }
  if (ks < 1) ks = 1;
  if (ks > kb_total) ks = kb_total;
  // no empty splits: shrink until every split owns at least one k-block
  while (ks > 1 && (long long)(ks - 1) * ((kb_total + ks - 1) / ks) >= kb_total) --ks;
  p.k_splits = ks;
  memset(&p.epi, 0, sizeof(p.epi));
  p.epi.ones_col = -1;
  p.epi.gb = gb; p.epi.gb_col = gw_cols;
  Sinks sk;
  sk.f32 = gw; sk.ld_f32 = ld_gw; sk.f32_cols = gw_cols; sk.reduce = true;
  CUtensorMap ta, tb;
  TRY(make_tmap(&ta, A, a_pitch, R, a_pitch, 64, 64));
  TRY(make_tmap(&tb, B, b_pitch, R, b_pitch, 64, 64));
  return launch_gemm(c, p, ta, tb, sk, name);
}

static GemmEpi epi_none() {
  GemmEpi e;
  memset(&e, 0, sizeof(e));
  e.ones_col = -1;
  return e;
}

// ----------------------------------------------------------------------------------------
// small launch helpers
// ----------------------------------------------------------------------------------------
static void pack_jobs(const Dims& d, const nrl_block_params* p, BlockWs& w, PackJob* j) {
  j[0] = PackJob{p->in_proj_weight, p->in_proj_bias, 3 * d.E, d.E, d.Ep, d.P3, w.win_f, w.win_t};
  j[1] = PackJob{p->out_proj_weight, p->out_proj_bias, d.E, d.E, d.Ep, d.Ep, w.wout_f, w.wout_t};
  j[2] = PackJob{p->add_weight, p->add_bias, d.Q, d.E, d.Ep, d.Qp, w.wadd_f, w.wadd_t};
}
// bf16 hi / lo operand copies of the weights of one block (p2 == nullptr) or two blocks, ONE launch
static int pack_weights(const Ctx& c, const Dims& d, const nrl_block_params* p, BlockWs& w,
                        const nrl_block_params* p2 = nullptr, BlockWs* w2 = nullptr) {
  PackJobs jobs;
  memset(&jobs, 0, sizeof(jobs));
  pack_jobs(d, p, w, jobs.j);
  if (p2) pack_jobs(d, p2, *w2, jobs.j + 3);
  const long long biggest = 3ll * d.E * d.Ep + (long long)d.E * d.P3;
  pack_weights_multi_kernel<<<dim3((unsigned)grid_for(biggest, 256, 1024), p2 ? 6 : 3), 256, 0, c.stream>>>(
      jobs, c.two_planes() ? 1 : 0);
  LAUNCH_CHECK("pack_weights");
  return NRL_OK;
}

struct AttnGeom {
  int S; long long seq_stride; int NB; long long batch_stride;
};

constexpr int ATTN_BWD_DEFAULT_VARIANT = 1;
constexpr int ATTN_FWD_SMEM_BUDGET = 110 * 1024;
constexpr int ATTN_BWD_SMEM_BUDGET = 100 * 1024;

constexpr int ATTN_S32_SMEM_BUDGET = 96 * 1024;
constexpr int ATTN_TMA_SMEM_MAX = 112 * 1024;  // two stages of four [32][168] fp32 boxes + barriers

// Which S <= 32 kernels take their operands through TMA-staged shared memory.  Measured on B200 at the bench size
// (profiles/r02_attention.md): forward 0.170 ms staged vs 0.174 direct; backward 0.444 staged vs 0.422 direct (its
// 168 registers allow only 2 CTAs of 5 warps beside the two 50 KB stages).  Default: forward staged, backward direct.
// NRL_ATTN_TMA=0 none, =1 both (A/B runs).
static bool attn_use_tma(bool backward) {
  static const int v = [] { const char* e = getenv("NRL_ATTN_TMA"); return e ? atoi(e) : -1; }();
  return v < 0 ? !backward : v != 0;
}
// shared memory of the staged kernels (nt = 3 forward, 4 backward), 0 if the shape does not qualify
static int attn_tma_smem(int nt, int hg, int DH, int S, int E, int ld_other) {
  const int pitch = attn_tma_pitch(hg * DH);
  if (pitch > 256 || S > 32 || (E & 3) || (ld_other & 3) || ((hg * DH) & 3)) return 0;
  const int bytes = 256 + 2 * nt * attn_tma_tile_bytes(S, pitch);
  return bytes <= ATTN_TMA_SMEM_MAX ? bytes : 0;
}

// NRL_ATTN_SIMT=1 routes S <= 32 attention to the older fp32 SIMT kernels (profiling A/B only)
static bool attn_force_simt() {
  static const bool v = [] { const char* e = getenv("NRL_ATTN_SIMT"); return e && e[0] == '1'; }();
  return v;
}

static bool attn_flash_on() {
  static const bool v = [] { const char* e = getenv("NRL_ATTN_FLASH"); return !(e && e[0] == '0'); }();
  return v;
}

template <int DH>
static int attn_set_attrs() {
  CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tile_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                ATTN_FWD_SMEM_BUDGET));
  CUDA_TRY(cudaFuncSetAttribute(attn_bwd_tile_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                ATTN_BWD_SMEM_BUDGET));
  if constexpr (DH == 48 || DH == 64) {
    CUDA_TRY(cudaFuncSetAttribute(attn_fwd_flash_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  flash_fwd_smem<DH>()));
    CUDA_TRY(cudaFuncSetAttribute(attn_bwd_flash_dq_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  flash_bwd_smem<DH>()));
    CUDA_TRY(cudaFuncSetAttribute(attn_bwd_flash_dkv_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  flash_bwd_smem<DH>()));
  }
  if constexpr (DH <= 32) {
    CUDA_TRY(cudaFuncSetAttribute(attn_bwd_ldsm_kernel<DH, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  3 * TitleCfg<DH>::HEAD_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(attn_bwd_ldsm_kernel<DH, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  TitleCfg<DH, 2>::HEAD_BYTES));
    CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tma_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ATTN_TMA_SMEM_MAX));
    CUDA_TRY(cudaFuncSetAttribute(attn_bwd_tma_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ATTN_TMA_SMEM_MAX));
    CUDA_TRY(cudaFuncSetAttribute(attn_fwd_s32_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ATTN_S32_SMEM_BUDGET));
    CUDA_TRY(cudaFuncSetAttribute(attn_bwd_s32_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  ATTN_S32_SMEM_BUDGET));
  }
  return NRL_OK;
}

// heads per CTA of the S <= 32 kernels: the largest divisor of H that is <= 5 (else 4);
// the kernels are compiled for at most 160 threads
static int attn_head_group(int H) {
  for (int hg = 5; hg >= 2; --hg)
    if (H % hg == 0) return hg;
  return H < 4 ? H : 4;
}

template <int DH>
static void launch_attn_fwd(const Ctx& c, const Dims& d, const AttnGeom& g, const BlockWs& w, long long R) {
  bf16* lo = c.two_planes() ? w.o + R * d.Ep : nullptr;
  if constexpr (DH <= 32) {
  if (g.S <= 32 && !attn_force_simt() && attn_use_tma(false)) {  // staged: operands through TMA into shared memory
    const int hg = attn_head_group(d.H), groups = (d.H + hg - 1) / hg;
    const int smem = attn_tma_smem(3, hg, DH, g.S, d.E, d.LDQ);
    CUtensorMap tq;
    if (smem && make_tmap_attn(&tq, w.qkv, d.LDQ, d.LDQ, g.S, g.seq_stride, g.NB, g.batch_stride,
                               attn_tma_pitch(hg * DH)) == NRL_OK) {
      const long long items = (long long)g.NB * groups;
      attn_fwd_tma_kernel<DH><<<grid_for(items, 1, 3 * g_dev.sm_count), 32 * hg, smem, c.stream>>>(
          tq, d.E, d.H, hg, g.S, g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH), w.o, lo, d.Ep, w.lse);
      return;
    }
  }
  if (g.S <= 32 && !attn_force_simt()) {  // warp-level tensor-core path (title tokens)
    const long long items = (long long)g.NB * d.H;
    attn_fwd_mma_kernel<DH><<<(unsigned)((items + 3) / 4), 128, 0, c.stream>>>(
        w.qkv, d.E, d.LDQ, d.H, g.S, g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH), w.o, lo, d.Ep, w.lse);
    return;
  }
  if (g.S <= 64 && !attn_force_simt()) {  // two 32-key blocks per warp (the NRMS user encoder, S = B = 64)
    const long long warps = 2ll * g.NB * d.H;
    attn_fwd_mma64_kernel<DH><<<(unsigned)((warps + 3) / 4), 128, 0, c.stream>>>(
        w.qkv, d.E, d.LDQ, d.H, g.S, g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH), w.o, lo, d.Ep, w.lse);
    return;
  }
  if (g.S <= 32) {  // register-resident SIMT path (NRL_ATTN_SIMT=1: A/B comparison only)
    const int hg = attn_head_group(d.H), groups = (d.H + hg - 1) / hg;
    const size_t smem = (size_t)g.S * attn_pitch(3 * hg * DH) * sizeof(float);
    if (smem <= (size_t)ATTN_S32_SMEM_BUDGET) {
      attn_fwd_s32_kernel<DH><<<grid_for((long long)g.NB * groups, 1, 16 * g_dev.sm_count), 32 * hg, smem, c.stream>>>(
          w.qkv, d.E, d.LDQ, d.H, g.S, g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH), w.o, lo, d.Ep, w.lse);
      return;
    }
  }
  }
  if constexpr (DH == 48 || DH == 64) {
    // long sequences / wide heads (the PLM head: attention across the N news of the call): flash-style tensor-core kernel
    // (NRL_ATTN_FLASH=0: the fp32 SIMT kernels below, A/B runs)
    if (g.S > 32 && attn_flash_on()) {
      const int nqb = (g.S + 127) / 128;
      attn_fwd_flash_kernel<DH><<<(unsigned)((long long)g.NB * d.H * nqb), 256, flash_fwd_smem<DH>(), c.stream>>>(
          w.qkv, d.E, d.LDQ, d.H, g.S, g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH), w.o, lo, d.Ep, w.lse,
          c.two_planes() ? 1 : 0);
      return;
    }
  }
  const size_t per_head = (size_t)g.S * 3 * DH * sizeof(float);
  int hp = (int)(ATTN_FWD_SMEM_BUDGET / per_head);
  if (hp > d.H) hp = d.H;
  {  // one (head, 32-query chunk) item per warp (8 warps): more, shorter CTAs fill the machine
    const int per_cta = 8 / ((g.S + 31) / 32);
    if (hp > per_cta && per_cta >= 1) hp = per_cta;
  }
  if (hp >= 1) {  // tile-resident fast path
    const int passes = (d.H + hp - 1) / hp;
    attn_fwd_tile_kernel<DH><<<grid_for((long long)g.NB * passes, 1, 8 * g_dev.sm_count), 256, hp * per_head, c.stream>>>(
        w.qkv, d.E, d.LDQ, d.H, g.S, g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH), hp, w.o, lo, d.Ep, w.lse);
    return;
  }
  const long long items = (long long)g.NB * d.H * ((g.S + 31) / 32);
  attn_fwd_kernel<DH><<<grid_for(items, attn_stream_warps(DH), 1 << 20), 32 * attn_stream_warps(DH), 0, c.stream>>>(
      w.qkv, d.E, d.LDQ, d.H, g.S, g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH), w.o, lo, d.Ep, w.lse);
}
template <int DH>
static void launch_attn_bwd(const Ctx& c, const Dims& d, const AttnGeom& g, const BlockWs& w, long long R) {
  bf16* lo = c.two_planes() ? w.dqkv + R * d.P3 : nullptr;
  if constexpr (DH <= 32) {
  if (g.S <= 32 && !attn_force_simt() && attn_use_tma(true)) {
    const int hg = attn_head_group(d.H), groups = (d.H + hg - 1) / hg;
    const int smem = attn_tma_smem(4, hg, DH, g.S, d.E, d.E);
    CUtensorMap tq, tdo;
    if (smem &&
        make_tmap_attn(&tq, w.qkv, d.LDQ, d.LDQ, g.S, g.seq_stride, g.NB, g.batch_stride, attn_tma_pitch(hg * DH)) == NRL_OK &&
        make_tmap_attn(&tdo, w.d_o, d.E, d.E, g.S, g.seq_stride, g.NB, g.batch_stride, attn_tma_pitch(hg * DH)) == NRL_OK) {
      const long long items = (long long)g.NB * groups;
      attn_bwd_tma_kernel<DH><<<grid_for(items, 1, 2 * g_dev.sm_count), 32 * hg, smem, c.stream>>>(
          tq, tdo, w.lse, d.E, d.H, hg, g.S, g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH), w.dqkv, lo, d.P3);
      return;
    }
  }
  if (g.S <= 32 && !attn_force_simt()) {
    const long long items = (long long)g.NB * d.H;
    // NRL_ATTN_BWD_VARIANT (A/B runs): 0 = fragments held in registers (168 registers, 3 CTAs / SM), 1 = fragments
    // re-read when needed again, 2 = re-read + register budget of 4 CTAs / SM, 3 = operands staged once per head as bf16
    // hi / lo planes in shared memory + ldmatrix (nrl_attn_title.cuh)
    static const int variant = [] { const char* e = getenv("NRL_ATTN_BWD_VARIANT"); return e ? atoi(e) : ATTN_BWD_DEFAULT_VARIANT; }();
    if (variant == 3 && (DH % 4) == 0 && !(d.E & 3) && !(d.LDQ & 3)) {
      constexpr int HG = 3;
      const int groups = (d.H + HG - 1) / HG;
      attn_bwd_ldsm_kernel<DH, HG, 1><<<(unsigned)((long long)g.NB * groups), 64 * HG, HG * TitleCfg<DH>::HEAD_BYTES, c.stream>>>(
          w.qkv, w.d_o, d.E, w.lse, d.E, d.LDQ, d.H, g.S, g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH), w.dqkv, lo,
          d.P3, c.two_planes() ? 1 : 0);
      return;
    }
#define NRL_ATTN_BWD_LAUNCH(RELOAD, MINB)                                                                            \
    attn_bwd_mma_kernel<DH, RELOAD, MINB><<<(unsigned)((items + 3) / 4), 128, 0, c.stream>>>(                        \
        w.qkv, w.d_o, d.E, w.lse, d.E, d.LDQ, d.H, g.S, g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH), w.dqkv, lo, \
        d.P3)
    if (variant == 0) NRL_ATTN_BWD_LAUNCH(false, 3);
    else if (variant == 2) NRL_ATTN_BWD_LAUNCH(true, 4);
    else NRL_ATTN_BWD_LAUNCH(true, 3);
#undef NRL_ATTN_BWD_LAUNCH
    return;
  }
  // 32 < S <= 64 (the NRMS user encoder: S = B = 64 along the batch axis): operands staged once per (item, head) as
  // bf16 hi / lo planes + ldmatrix, four warps of 16 rows (nrl_attn_title.cuh).  NRL_ATTN_BWD64=0: the register /
  // movmatrix kernel below (250 registers, 61 us for 750 problems against ~15 us here)
  static const bool ldsm64 = [] { const char* e = getenv("NRL_ATTN_BWD64"); return !(e && e[0] == '0'); }();
  if (g.S <= 64 && !attn_force_simt() && ldsm64 && (DH % 4) == 0 && !(d.E & 3) && !(d.LDQ & 3)) {
    attn_bwd_ldsm_kernel<DH, 1, 2><<<(unsigned)((long long)g.NB * d.H), 128, TitleCfg<DH, 2>::HEAD_BYTES, c.stream>>>(
        w.qkv, w.d_o, d.E, w.lse, d.E, d.LDQ, d.H, g.S, g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH), w.dqkv, lo,
        d.P3, c.two_planes() ? 1 : 0);
    return;
  }
  if (g.S <= 64 && !attn_force_simt()) {
    attn_bwd_mma64_kernel<DH><<<(unsigned)((long long)g.NB * d.H), 128, 0, c.stream>>>(
        w.qkv, w.d_o, d.E, w.o, c.two_planes() ? w.o + R * d.Ep : nullptr, d.Ep, w.lse, d.E, d.LDQ, d.H, g.S,
        g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH), w.dqkv, lo, d.P3);
    return;
  }
  if (g.S <= 32) {
    const int hg = attn_head_group(d.H), groups = (d.H + hg - 1) / hg;
    const size_t smem = ((size_t)g.S * attn_pitch(4 * hg * DH) + (size_t)hg * 32 * 33) * sizeof(float);
    if (smem <= (size_t)ATTN_S32_SMEM_BUDGET) {
      attn_bwd_s32_kernel<DH><<<grid_for((long long)g.NB * groups, 1, 16 * g_dev.sm_count), 32 * hg, smem, c.stream>>>(
          w.qkv, w.d_o, d.E, w.lse, d.E, d.LDQ, d.H, g.S, g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH),
          w.dqkv, lo, d.P3);
      return;
    }
  }
  }
  if constexpr (DH == 48 || DH == 64) {
    if (g.S > 32 && attn_flash_on()) {
      const int nqb = (g.S + 127) / 128;
      const unsigned grid = (unsigned)((long long)g.NB * d.H * nqb);
      attn_delta_kernel<<<grid_for(R, 8, 8 * g_dev.sm_count), 256, 0, c.stream>>>(
          w.d_o, d.E, w.o, c.two_planes() ? w.o + R * d.Ep : nullptr, d.Ep, R, d.H, DH, w.delta);
      attn_bwd_flash_dq_kernel<DH><<<grid, 256, flash_bwd_smem<DH>(), c.stream>>>(
          w.qkv, w.d_o, d.E, w.lse, w.delta, d.E, d.LDQ, d.H, g.S, g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH),
          w.dqkv, lo, d.P3, c.two_planes() ? 1 : 0);
      attn_bwd_flash_dkv_kernel<DH><<<grid, 256, flash_bwd_smem<DH>(), c.stream>>>(
          w.qkv, w.d_o, d.E, w.lse, w.delta, d.E, d.LDQ, d.H, g.S, g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH),
          w.dqkv, lo, d.P3, c.two_planes() ? 1 : 0);
      return;
    }
  }
  const size_t per_head = (size_t)g.S * (5 * DH + 2) * sizeof(float);
  int hp = (int)(ATTN_BWD_SMEM_BUDGET / per_head);
  if (hp > d.H) hp = d.H;
  {
    const int per_cta = 8 / ((g.S + 31) / 32);
    if (hp > per_cta && per_cta >= 1) hp = per_cta;
  }
  if (hp >= 1) {
    const int passes = (d.H + hp - 1) / hp;
    attn_bwd_tile_kernel<DH><<<grid_for((long long)g.NB * passes, 1, 8 * g_dev.sm_count), 256, hp * per_head, c.stream>>>(
        w.qkv, w.d_o, d.E, w.lse, d.E, d.LDQ, d.H, g.S, g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH), hp,
        w.dqkv, lo, d.P3);
    return;
  }
  const long long items = (long long)g.NB * d.H;
  attn_bwd_kernel<DH><<<grid_for(items, attn_stream_warps(DH), 1 << 20), 32 * attn_stream_warps(DH), 0, c.stream>>>(
      w.qkv, w.d_o, d.E, w.o, c.two_planes() ? w.o + R * d.Ep : nullptr, d.Ep, w.lse, d.E, d.LDQ, d.H, g.S,
      g.seq_stride, g.NB, g.batch_stride, sqrtf(1.0f / DH), w.dqkv, lo, d.P3);
}

struct DropCfg {
  int on; float scale; uint32_t thr; unsigned long long seed;
};
static void epi_dropout(GemmEpi& e, const DropCfg& dc, const uint32_t* words, int mw) {
  if (!dc.on) return;
  e.drop_words = words; e.drop_mw = mw; e.drop_scale = dc.scale;
}
static DropCfg make_drop(float p, int training, unsigned long long seed) {
  DropCfg dc;
  dc.on = (training && p > 0.f) ? 1 : 0;
  dc.scale = 1.0f / (1.0f - p);
  dc.thr = drop_threshold(p);
  dc.seed = seed;
  return dc;
}

static int attn_attrs_init() {
  static bool done = false;
  if (done) return NRL_OK;
  TRY(attn_set_attrs<16>());
  TRY(attn_set_attrs<20>());
  TRY(attn_set_attrs<32>());
  TRY(attn_set_attrs<48>());
  TRY(attn_set_attrs<64>());
  done = true;
  return NRL_OK;
}

// additive pooling forward: the TMA-staged kernel when the [L][E] tile of a group can be bulk-copied (16-byte rows, two
// stages fit shared memory, enough groups to keep persistent CTAs busy), else the direct kernel.  NRL_POOL_TMA=0: direct.
static int launch_pool_fwd(const Ctx& c, const float* score, const float* Y, int E, int L, long long G, float* w_out,
                           float* out) {
  static const bool tma_on = [] { const char* e = getenv("NRL_POOL_TMA"); return !(e && e[0] == '0'); }();
  static bool attr_done = false;
  const size_t smem = pool_fwd_tma_smem(E, L);
  if (tma_on && !(E & 3) && L <= 64 && smem <= 200 * 1024 && G >= 4 * g_dev.sm_count &&
      !(reinterpret_cast<uintptr_t>(Y) & 15) && !(reinterpret_cast<uintptr_t>(out) & 15)) {
    if (!attr_done) {
      CUDA_TRY(cudaFuncSetAttribute(pool_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_done = true;
    }
    const int per_sm = (int)((220 * 1024) / (smem + 1024));
    const int ctas = g_dev.sm_count * (per_sm < 1 ? 1 : per_sm > 4 ? 4 : per_sm);
    pool_fwd_tma_kernel<<<(unsigned)(G < ctas ? G : ctas), 128, smem, c.stream>>>(score, Y, E, L, G, w_out, out);
    LAUNCH_CHECK("pool_fwd");
    return NRL_OK;
  }
  pool_fwd_kernel<<<dim3((unsigned)grid_for(G, 1, 1 << 20), (unsigned)((E + 511) / 512)), 128, L * sizeof(float), c.stream>>>(
      score, Y, E, L, G, w_out, out);
  LAUNCH_CHECK("pool_fwd");
  return NRL_OK;
}

// MHSA + additive pooling over R rows already staged in w.x (split planes).
static int block_forward(const Ctx& c, const Dims& d, BlockWs& w, long long R, const AttnGeom& ag,
                         long long G, int L, const nrl_block_params* prm, const DropCfg& drop,
                         float* out_vec) {
  TRY(attn_attrs_init());
  // K3: QKV = X W_in^T + b_in   (bias rides on the ones column)
  {
    GemmEpi e = epi_none();
    Sinks sk;
    sk.f32 = w.qkv; sk.ld_f32 = d.LDQ; sk.f32_cols = 3 * d.E;
    TRY(gemm_nt(c, w.x, R, d.Ep, w.win_f, 3 * d.E, d.Ep, d.Ep, e, sk, "gemm in_proj"));
  }
  // K4: per-head softmax(q k^T) v
  if (d.DH == 16) launch_attn_fwd<16>(c, d, ag, w, R);
  else if (d.DH == 20) launch_attn_fwd<20>(c, d, ag, w, R);
  else if (d.DH == 32) launch_attn_fwd<32>(c, d, ag, w, R);
  else if (d.DH == 48) launch_attn_fwd<48>(c, d, ag, w, R);
  else launch_attn_fwd<64>(c, d, ag, w, R);
  LAUNCH_CHECK("attn_fwd");
  // K5: Y = O W_out^T + b_out  (+ dropout site 1), fp32 and split planes
  {
    GemmEpi e = epi_none();
    Sinks sk;
    sk.f32 = w.y; sk.ld_f32 = d.E; sk.f32_cols = d.E;
    sk.sp = w.yp; sk.ld_sp = d.Ep; sk.sp_cols = d.Ep; sk.ones_col = d.E;
    epi_dropout(e, drop, w.mask1, d.MW);
    TRY(gemm_nt(c, w.o, R, d.Ep, w.wout_f, d.E, d.Ep, d.Ep, e, sk, "gemm out_proj"));
  }
  // K6: a = tanh(Y W_add^T + b_add), score = a . query  (fused epilogue)
  {
    GemmEpi e = epi_none();
    e.qvec = prm->add_query; e.score = w.s;
    Sinks sk;
    sk.f32 = w.a; sk.ld_f32 = d.Q; sk.f32_cols = d.Q;
    TRY(gemm_nt(c, w.yp, R, d.Ep, w.wadd_f, d.Q, d.Ep, d.Ep, e, sk, "gemm additive"));
  }
  TRY(launch_pool_fwd(c, w.s, w.y, d.E, L, G, w.w, out_vec));
  return NRL_OK;
}

// Backward of block_forward.  d_vec [G][E] -> w.dx [R][E] (gradient w.r.t. the block input,
// dropout site 0 applied when drop0.on), parameter gradients accumulated into g.
// The three weight-gradient GEMMs of a block depend on the data-gradient chain only through their operands
// (dApre, dY, dQKV), and nothing in the backward pass consumes them: they run on a second, lower-priority stream so
// that their CTAs fill the SMs the main chain leaves idle -- the partial last wave and the ramp of every persistent
// GEMM, and the register-bound attention backward, which occupies a third of an SM's shared memory-free resources.
// NRL_WGRAD_STREAM=0 keeps everything on the caller's stream (also forced while per-launch profiling is on, so that
// every launch's duration is its own).
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  int dev = -1;
};
static SideStream g_side;
static bool side_wanted() {
  static const bool v = [] { const char* e = getenv("NRL_WGRAD_STREAM"); return !(e && e[0] == '0'); }();
  return v && !g_prof.on;
}
static int side_init() {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (g_side.s && g_side.dev == dev) return NRL_OK;
  int lo = 0, hi = 0;
  CUDA_TRY(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // lo = numerically greatest = lowest priority
  CUDA_TRY(cudaStreamCreateWithPriority(&g_side.s, cudaStreamNonBlocking, lo));
  for (int i = 0; i < 4; ++i) CUDA_TRY(cudaEventCreateWithFlags(&g_side.ev[i], cudaEventDisableTiming));
  g_side.dev = dev;
  return NRL_OK;
}
// side stream waits for everything issued on `main` so far
static int side_follow(cudaStream_t main, int ev) {
  CUDA_TRY(cudaEventRecord(g_side.ev[ev], main));
  CUDA_TRY(cudaStreamWaitEvent(g_side.s, g_side.ev[ev], 0));
  return NRL_OK;
}

// the caller's stream continues only when the weight gradients issued so far are complete
static int side_join(cudaStream_t main) {
  if (!g_side.s) return NRL_OK;
  CUDA_TRY(cudaEventRecord(g_side.ev[3], g_side.s));
  CUDA_TRY(cudaStreamWaitEvent(main, g_side.ev[3], 0));
  return NRL_OK;
}

static int block_backward(const Ctx& c, const Dims& d, BlockWs& w, long long R, const AttnGeom& ag,
                          long long G, int L, const nrl_block_params* prm, const DropCfg& drop1,
                          const DropCfg& drop0, const float* d_vec, nrl_block_grads* g, bool join = true) {
  TRY(attn_attrs_init());
  const bool side = side_wanted() && side_init() == NRL_OK;
  Ctx cw = c;  // where the weight-gradient GEMMs go
  if (side) cw.stream = g_side.s;
  bf16* lo_or_null_dap = c.two_planes() ? w.dap + R * d.Qp : nullptr;
  pool_bwd_kernel<<<grid_for(G, 1, 8 * g_dev.sm_count), 256, L * sizeof(float), c.stream>>>(
      d_vec, w.y, w.w, w.a, prm->add_query, d.E, d.Q, d.Qp, L, G, nullptr, w.dap, lo_or_null_dap,
      g->add_query, g->add_bias);
  LAUNCH_CHECK("pool_bwd");
  if (side) TRY(side_follow(c.stream, 0));
  // dY = dropout1'( w_r * dVec[g] + dApre W_add )  -> split planes
  {
    GemmEpi e = epi_none();
    e.add_w = w.w; e.add_vec = d_vec; e.ld_addvec = d.E; e.add_L = L;  // + w_r * dVec[g(r)]
    Sinks sk;
    sk.sp = w.dyp; sk.ld_sp = d.Ep; sk.sp_cols = d.Ep; sk.ones_col = -1;
    epi_dropout(e, drop1, w.mask1, d.MW);
    TRY(gemm_nt(c, w.dap, R, d.Qp, w.wadd_t, d.E, d.Qp, d.Qp, e, sk, "gemm additive dgrad"));
  }
  if (side) TRY(side_follow(c.stream, 1));
  // dW_add, db_add
  // (db_add is summed in fp32 by pool_bwd)
  TRY(gemm_tn(cw, w.dap, d.Q, d.Qp, w.yp, d.Ep, d.Ep, R, g->add_weight, d.E, d.E, nullptr,
              "gemm additive wgrad"));
  // dO = dY W_out
  {
    GemmEpi e = epi_none();
    Sinks sk;
    sk.f32 = w.d_o; sk.ld_f32 = d.E; sk.f32_cols = d.E;
    TRY(gemm_nt(c, w.dyp, R, d.Ep, w.wout_t, d.E, d.Ep, d.Ep, e, sk, "gemm out_proj dgrad"));
  }
  // dW_out, db_out
  TRY(gemm_tn(cw, w.dyp, d.E, d.Ep, w.o, d.Ep, d.Ep, R, g->out_proj_weight, d.E, d.E, g->out_proj_bias,
              "gemm out_proj wgrad"));
  if (d.DH == 16) launch_attn_bwd<16>(c, d, ag, w, R);
  else if (d.DH == 20) launch_attn_bwd<20>(c, d, ag, w, R);
  else if (d.DH == 32) launch_attn_bwd<32>(c, d, ag, w, R);
  else if (d.DH == 48) launch_attn_bwd<48>(c, d, ag, w, R);
  else launch_attn_bwd<64>(c, d, ag, w, R);
  LAUNCH_CHECK("attn_bwd");
  if (side) TRY(side_follow(c.stream, 2));
  // dX = dropout0'( dQKV W_in )
  {
    GemmEpi e = epi_none();
    Sinks sk;
    sk.f32 = w.dx; sk.ld_f32 = d.E; sk.f32_cols = d.E;
    epi_dropout(e, drop0, w.mask0, d.MW);
    TRY(gemm_nt(c, w.dqkv, R, d.P3, w.win_t, d.E, d.P3, d.P3, e, sk, "gemm in_proj dgrad"));
  }
  // dW_in, db_in
  TRY(gemm_tn(cw, w.dqkv, 3 * d.E, d.P3, w.x, d.Ep, d.Ep, R, g->in_proj_weight, d.E, d.E, g->in_proj_bias,
              "gemm in_proj wgrad"));
  if (side && join) TRY(side_join(c.stream));
  return NRL_OK;
}

static int check_common(const void* ws, size_t ws_bytes, size_t need) {
  if (!ws) return fail(NRL_ERR_INVALID_ARG, "workspace is NULL");
  if ((reinterpret_cast<uintptr_t>(ws) & 1023) != 0)
    return fail(NRL_ERR_INVALID_ARG, "workspace must be 1024-byte aligned");
  if (ws_bytes < need)
    return fail(NRL_ERR_WORKSPACE_TOO_SMALL, "workspace %zu < required %zu bytes", ws_bytes, need);
  return NRL_OK;
}

// ----------------------------------------------------------------------------------------
// news encoder
// ----------------------------------------------------------------------------------------
static int news_fwd_impl(const Ctx& c, const Dims& d, BlockWs& w, const long long* ids,
                         long long n_news, int L, const float* table, long long V1,
                         const nrl_block_params* prm, const DropCfg& drop, float* out, bool packed = false,
                         long long row0 = 0) {
  const long long R = n_news * L;
  if (!packed) TRY(pack_weights(c, d, prm, w));
  if (drop.on) {
    dropout_words_kernel<<<grid_for(2 * R * d.MW, 256, 16 * g_dev.sm_count), 256, 0, c.stream>>>(
        drop.seed, drop.thr, R, d.E, d.MW, w.mask0, w.mask1, row0);
    LAUNCH_CHECK("dropout_words");
  }
  gather_split_kernel<<<grid_for(R, 8, 1 << 20), 256, 0, c.stream>>>(
      ids, R, table, V1, d.E, d.Ep, w.x, c.two_planes() ? w.x + R * d.Ep : nullptr, nullptr,
      drop.on ? w.mask0 : nullptr, d.MW, drop.scale);
  LAUNCH_CHECK("gather_split");
  AttnGeom ag{L, 1, (int)n_news, L};
  return block_forward(c, d, w, R, ag, n_news, L, prm, drop, out);
}
static int news_bwd_impl(const Ctx& c, const Dims& d, BlockWs& w, const long long* ids,
                         long long n_news, int L, long long V1, const nrl_block_params* prm, const DropCfg& drop,
                         const float* d_out, nrl_block_grads* g, float* d_table) {
  const long long R = n_news * L;
  AttnGeom ag{L, 1, (int)n_news, L};
  // the weight-gradient stream is joined after the embedding-gradient scatter (they are independent)
  TRY(block_backward(c, d, w, R, ag, n_news, L, prm, drop, drop, d_out, g, false));
  if (d_table) {
    emb_grad_kernel<<<grid_for(R, 8, 1 << 20), 256, 0, c.stream>>>(ids, R, V1, w.dx, d.E, d_table);
    LAUNCH_CHECK("emb_grad");
  }
  return side_join(c.stream);
}

extern "C" {

const char* nrl_version(void) { return "newsreclib_b200 0.1 (sm_100a, tcgen05)"; }
const char* nrl_last_error(void) { return g_err; }
long long nrl_launch_count(void) { return g_launches.load(); }

int nrl_profile_start(void* stream) {
  g_prof.stream = static_cast<cudaStream_t>(stream);
  g_prof.used = 0;
  g_prof.names.clear();
  g_prof.on = true;
  prof_mark("<start>");
  return NRL_OK;
}
int nrl_profile_stop(char* names, int name_stride, float* ms, int max_records) {
  g_prof.on = false;
  if (g_prof.used == 0) return 0;
  CUDA_TRY(cudaEventSynchronize(g_prof.pool[g_prof.used - 1]));
  int n = 0;
  for (size_t i = 1; i < g_prof.used && n < max_records; ++i, ++n) {
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, g_prof.pool[i - 1], g_prof.pool[i]));
    if (ms) ms[n] = t;
    if (names && name_stride > 0) {
      strncpy(names + (size_t)n * name_stride, g_prof.names[i], name_stride - 1);
      names[(size_t)n * name_stride + name_stride - 1] = 0;
    }
  }
  return n;
}

size_t nrl_news_encoder_ws_bytes(long long n_news, int L, nrl_dims dims) {
  Dims d;
  if (make_dims(dims, d) != NRL_OK || n_news < 0 || L <= 0) return 0;
  Bump b(nullptr);
  BlockWs w;
  carve_block(b, n_news * L, d, w);
  return b.off + 1024;
}

int nrl_news_encoder_fwd(const long long* ids, long long n_news, int L, const float* table,
                         long long V1, const nrl_block_params* params, nrl_dims dims,
                         float dropout_p, int training, unsigned long long seed, float* out,
                         void* ws, size_t ws_bytes, int precision, void* stream) {
  Dims d;
  TRY(make_dims(dims, d));
  if (!ids || !table || !params || !out || n_news <= 0 || L <= 0 || V1 <= 0)
    return fail(NRL_ERR_INVALID_ARG, "nrl_news_encoder_fwd: null pointer or empty input");
  if (dropout_p < 0.f || dropout_p >= 1.f) return fail(NRL_ERR_INVALID_ARG, "dropout_p out of [0,1)");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_news_encoder_ws_bytes(n_news, L, dims)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  Bump b(ws);
  BlockWs w;
  carve_block(b, n_news * L, d, w);
  return news_fwd_impl(c, d, w, ids, n_news, L, table, V1, params, make_drop(dropout_p, training, seed), out);
}

int nrl_news_encoder_bwd(const long long* ids, long long n_news, int L, long long V1,
                         const nrl_block_params* params, nrl_dims dims, float dropout_p,
                         int training, unsigned long long seed, const float* d_out,
                         nrl_block_grads* grads, float* d_table, void* ws, size_t ws_bytes,
                         int precision, void* stream) {
  Dims d;
  TRY(make_dims(dims, d));
  if (!ids || !params || !d_out || !grads || n_news <= 0 || L <= 0 || V1 <= 0)
    return fail(NRL_ERR_INVALID_ARG, "nrl_news_encoder_bwd: null pointer or empty input");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_news_encoder_ws_bytes(n_news, L, dims)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  Bump b(ws);
  BlockWs w;
  carve_block(b, n_news * L, d, w);
  return news_bwd_impl(c, d, w, ids, n_news, L, V1, params, make_drop(dropout_p, training, seed), d_out,
                       grads, d_table);
}

// ----------------------------------------------------------------------------------------
// user encoder
// ----------------------------------------------------------------------------------------
size_t nrl_user_encoder_ws_bytes(int B, int Hmax, nrl_dims dims) {
  Dims d;
  if (make_dims(dims, d) != NRL_OK || B <= 0 || Hmax <= 0) return 0;
  Bump b(nullptr);
  BlockWs w;
  carve_block(b, (long long)B * Hmax, d, w);
  return b.off + 1024;
}

static AttnGeom user_geom(int B, int Hmax, int axis) {
  if (axis == 0) return AttnGeom{B, Hmax, Hmax, 1};  // reference: sequence = the B impressions
  return AttnGeom{Hmax, 1, B, Hmax};                // along the history
}

// MHSA (+ the two dropout sites when drop.on) + additive pooling over dense rows x [B][Hmax][E]
static int mhsa_pool_fwd(const Ctx& c, const Dims& d, const float* x, int B, int Hmax, int axis,
                         const nrl_block_params* params, const DropCfg& drop, float* out, void* ws) {
  Bump b(ws);
  BlockWs w;
  const long long R = (long long)B * Hmax;
  carve_block(b, R, d, w);
  TRY(pack_weights(c, d, params, w));
  if (drop.on) {
    dropout_words_kernel<<<grid_for(2 * R * d.MW, 256, 16 * g_dev.sm_count), 256, 0, c.stream>>>(
        drop.seed, drop.thr, R, d.E, d.MW, w.mask0, w.mask1);
    LAUNCH_CHECK("dropout_words");
  }
  // dense rows -> (dropout site 0) -> split planes: the gather kernel with the identity index
  gather_split_kernel<<<grid_for(R, 8, 1 << 20), 256, 0, c.stream>>>(
      nullptr, R, x, R, d.E, d.Ep, w.x, c.two_planes() ? w.x + R * d.Ep : nullptr, nullptr,
      drop.on ? w.mask0 : nullptr, d.MW, drop.scale);
  LAUNCH_CHECK("split_rows");
  return block_forward(c, d, w, R, user_geom(B, Hmax, axis), B, Hmax, params, drop, out);
}
static int mhsa_pool_bwd(const Ctx& c, const Dims& d, int B, int Hmax, int axis,
                         const nrl_block_params* params, const DropCfg& drop, const float* d_out,
                         nrl_block_grads* grads, float* d_x, void* ws) {
  Bump b(ws);
  BlockWs w;
  const long long R = (long long)B * Hmax;
  carve_block(b, R, d, w);
  TRY(block_backward(c, d, w, R, user_geom(B, Hmax, axis), B, Hmax, params, drop, drop, d_out, grads));
  CUDA_TRY(cudaMemcpyAsync(d_x, w.dx, (size_t)R * d.E * sizeof(float), cudaMemcpyDeviceToDevice, c.stream));
  return NRL_OK;
}

int nrl_user_encoder_fwd(const float* hist, int B, int Hmax, const nrl_block_params* params,
                         nrl_dims dims, int attention_axis, float* user, void* ws,
                         size_t ws_bytes, int precision, void* stream) {
  Dims d;
  TRY(make_dims(dims, d));
  if (!hist || !params || !user || B <= 0 || Hmax <= 0)
    return fail(NRL_ERR_INVALID_ARG, "nrl_user_encoder_fwd: null pointer or empty input");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_user_encoder_ws_bytes(B, Hmax, dims)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  return mhsa_pool_fwd(c, d, hist, B, Hmax, attention_axis, params, make_drop(0.f, 0, 0), user, ws);
}

int nrl_user_encoder_bwd(int B, int Hmax, const nrl_block_params* params, nrl_dims dims,
                         int attention_axis, const float* d_user, nrl_block_grads* grads,
                         float* d_hist, void* ws, size_t ws_bytes, int precision, void* stream) {
  Dims d;
  TRY(make_dims(dims, d));
  if (!params || !d_user || !grads || !d_hist || B <= 0 || Hmax <= 0)
    return fail(NRL_ERR_INVALID_ARG, "nrl_user_encoder_bwd: null pointer or empty input");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_user_encoder_ws_bytes(B, Hmax, dims)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  return mhsa_pool_bwd(c, d, B, Hmax, attention_axis, params, make_drop(0.f, 0, 0), d_user, grads, d_hist, ws);
}

// PLM head = the same block over x [N][T][E] with both dropout sites (text.py:93-100)
int nrl_plm_head_fwd(const float* x, int N, int T, const nrl_block_params* params, nrl_dims dims,
                     int attention_axis, float dropout_p, int training, unsigned long long seed,
                     float* out, void* ws, size_t ws_bytes, int precision, void* stream) {
  Dims d;
  TRY(make_dims(dims, d));
  if (!x || !params || !out || N <= 0 || T <= 0)
    return fail(NRL_ERR_INVALID_ARG, "nrl_plm_head_fwd: null pointer or empty input");
  if (dropout_p < 0.f || dropout_p >= 1.f) return fail(NRL_ERR_INVALID_ARG, "dropout_p out of [0,1)");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_user_encoder_ws_bytes(N, T, dims)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  return mhsa_pool_fwd(c, d, x, N, T, attention_axis, params, make_drop(dropout_p, training, seed), out, ws);
}
int nrl_plm_head_bwd(int N, int T, const nrl_block_params* params, nrl_dims dims,
                     int attention_axis, float dropout_p, int training, unsigned long long seed,
                     const float* d_out, nrl_block_grads* grads, float* d_x, void* ws,
                     size_t ws_bytes, int precision, void* stream) {
  Dims d;
  TRY(make_dims(dims, d));
  if (!params || !d_out || !grads || !d_x || N <= 0 || T <= 0)
    return fail(NRL_ERR_INVALID_ARG, "nrl_plm_head_bwd: null pointer or empty input");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_user_encoder_ws_bytes(N, T, dims)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  return mhsa_pool_bwd(c, d, N, T, attention_axis, params, make_drop(dropout_p, training, seed), d_out, grads,
                       d_x, ws);
}

// ----------------------------------------------------------------------------------------
// additive attention alone
// ----------------------------------------------------------------------------------------
struct AddWs {
  bf16 *wf, *wt, *xp, *dap;
  float *a, *s, *w;
};
static void carve_add(Bump& b, long long R, int D, int Q, AddWs& w) {
  const int Dp = round_up(D + 1, 16), Qp = round_up(Q, 16);
  w.wf = b.take<bf16>(2ull * Q * Dp);
  w.wt = b.take<bf16>(2ull * D * Qp);
  w.xp = b.take<bf16>(2ull * R * Dp);
  w.a = b.take<float>((size_t)R * Q);
  w.s = b.take<float>((size_t)R);
  w.w = b.take<float>((size_t)R);
  w.dap = b.take<bf16>(2ull * R * Qp);
}
size_t nrl_additive_ws_bytes(long long G, int L, int D, int Q) {
  if (G <= 0 || L <= 0 || D <= 0 || Q <= 0) return 0;
  Bump b(nullptr);
  AddWs w;
  carve_add(b, G * L, D, Q, w);
  return b.off + 1024;
}

int nrl_additive_fwd(const float* x, long long G, int L, int D, int Q, const float* weight,
                     const float* bias, const float* query, float* out, void* ws,
                     size_t ws_bytes, int precision, void* stream) {
  if (!x || !weight || !bias || !query || !out || G <= 0 || L <= 0 || D <= 0 || Q <= 0 || Q > 256 || (D & 3))
    return fail(NRL_ERR_INVALID_ARG, "nrl_additive_fwd: bad argument (need Q <= 256, D %% 4 == 0)");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_additive_ws_bytes(G, L, D, Q)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  const long long R = G * L;
  const int Dp = round_up(D + 1, 16), Qp = round_up(Q, 16);
  Bump b(ws);
  AddWs w;
  carve_add(b, R, D, Q, w);
  const int tp = c.two_planes() ? 1 : 0;
  pack_weight_kernel<<<grid_for((long long)Q * Dp + (long long)D * Qp, 256, 4096), 256, 0, c.stream>>>(
      weight, bias, Q, D, Dp, Qp, w.wf, w.wt, tp);
  LAUNCH_CHECK("pack_weight(additive)");
  dense_scatter_kernel<<<grid_for(R, 1, 1 << 20), 128, 0, c.stream>>>(
      x, nullptr, (int)R, 1, D, Dp, nullptr, w.xp, tp ? w.xp + R * Dp : nullptr);
  LAUNCH_CHECK("split_rows");
  GemmEpi e = epi_none();
  e.qvec = query; e.score = w.s;
  Sinks sk;
  sk.f32 = w.a; sk.ld_f32 = Q; sk.f32_cols = Q;
  TRY(gemm_nt(c, w.xp, R, Dp, w.wf, Q, Dp, Dp, e, sk, "gemm additive"));
  TRY(launch_pool_fwd(c, w.s, x, D, L, G, w.w, out));
  return NRL_OK;
}

int nrl_additive_bwd(const float* x, long long G, int L, int D, int Q, const float* weight,
                     const float* query, const float* d_out, float* dx, float* g_weight,
                     float* g_bias, float* g_query, void* ws, size_t ws_bytes, int precision,
                     void* stream) {
  if (!x || !weight || !query || !d_out || !dx || !g_weight || !g_bias || !g_query || G <= 0 || L <= 0 ||
      D <= 0 || Q <= 0 || Q > 256 || (D & 3))
    return fail(NRL_ERR_INVALID_ARG, "nrl_additive_bwd: bad argument");
  (void)weight;
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_additive_ws_bytes(G, L, D, Q)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  const long long R = G * L;
  const int Dp = round_up(D + 1, 16), Qp = round_up(Q, 16);
  Bump b(ws);
  AddWs w;
  carve_add(b, R, D, Q, w);
  const int tp = c.two_planes() ? 1 : 0;
  pool_bwd_kernel<<<grid_for(G, 1, 8 * g_dev.sm_count), 256, L * sizeof(float), c.stream>>>(
      d_out, x, w.w, w.a, query, D, Q, Qp, L, G, nullptr, w.dap, tp ? w.dap + R * Qp : nullptr, g_query, g_bias);
  LAUNCH_CHECK("pool_bwd");
  {  // dX = w_r * dOut[g] + dApre W
    GemmEpi e = epi_none();
    e.add_w = w.w; e.add_vec = d_out; e.ld_addvec = D; e.add_L = L;
    Sinks sk;
    sk.f32 = dx; sk.ld_f32 = D; sk.f32_cols = D;
    TRY(gemm_nt(c, w.dap, R, Qp, w.wt, D, Qp, Qp, e, sk, "gemm additive dgrad"));
  }
  TRY(gemm_tn(c, w.dap, Q, Qp, w.xp, Dp, Dp, R, g_weight, D, D, nullptr, "gemm additive wgrad"));
  return NRL_OK;
}

// ----------------------------------------------------------------------------------------
// ragged <-> dense, scorer, loss, optimizer
// ----------------------------------------------------------------------------------------
int nrl_segment_offsets(const long long* seg, long long n, int B, int* off, void* stream) {
  if (!seg || !off || n < 0 || B <= 0) return fail(NRL_ERR_INVALID_ARG, "nrl_segment_offsets: bad argument");
  segment_offsets_kernel<<<(B + 1 + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(seg, n, B, off, 0);
  LAUNCH_CHECK("segment_offsets");
  return NRL_OK;
}

int nrl_to_dense_fwd(const float* x, const int* off, int B, int M, int E, float* dense, void* stream) {
  if (!x || !off || !dense || B <= 0 || M <= 0 || E <= 0) return fail(NRL_ERR_INVALID_ARG, "nrl_to_dense_fwd: bad argument");
  dense_scatter_kernel<<<grid_for((long long)B * M, 1, 1 << 20), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x, off, B, M, E, E, dense, nullptr, nullptr);
  LAUNCH_CHECK("dense_scatter");
  return NRL_OK;
}

int nrl_to_dense_bwd(const float* d_dense, const int* off, int B, int M, int E, float* dx, void* stream) {
  if (!d_dense || !off || !dx || B <= 0 || M <= 0 || E <= 0) return fail(NRL_ERR_INVALID_ARG, "nrl_to_dense_bwd: bad argument");
  dim3 grid(M, B < 65535 ? B : 65535);
  dense_gather_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_dense, off, B, M, E, dx);
  LAUNCH_CHECK("dense_gather");
  return NRL_OK;
}

int nrl_gather_rows(const void* table, long long n_table_rows, int row_bytes, const long long* idx,
                    long long n, void* out, void* stream) {
  if (!table || !idx || !out || n_table_rows <= 0 || n <= 0 || row_bytes <= 0 || (row_bytes & 3))
    return fail(NRL_ERR_INVALID_ARG, "nrl_gather_rows: bad argument (row_bytes must be a positive multiple of 4)");
  TRY(device_init());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool w8 = (row_bytes & 7) == 0 && ((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) & 7) == 0;
  if (w8) {
    const int words = row_bytes / 8;
    gather_rows_kernel<unsigned long long><<<grid_for(n * words, 256, 16 * g_dev.sm_count), 256, 0, st>>>(
        static_cast<const unsigned long long*>(table), n_table_rows, words, idx, n, static_cast<unsigned long long*>(out));
  } else {
    const int words = row_bytes / 4;
    gather_rows_kernel<unsigned int><<<grid_for(n * words, 256, 16 * g_dev.sm_count), 256, 0, st>>>(
        static_cast<const unsigned int*>(table), n_table_rows, words, idx, n, static_cast<unsigned int*>(out));
  }
  LAUNCH_CHECK("gather_rows");
  return NRL_OK;
}

int nrl_late_fusion_fwd(const float* hist_vec, const int* off, int B, int E, float* user, void* stream) {
  if (!hist_vec || !off || !user || B <= 0 || E <= 0) return fail(NRL_ERR_INVALID_ARG, "nrl_late_fusion_fwd: bad argument");
  late_fusion_fwd_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(hist_vec, off, B, E, user);
  LAUNCH_CHECK("late_fusion_fwd");
  return NRL_OK;
}
int nrl_late_fusion_bwd(const float* d_user, const int* off, int B, int E, float* d_hist_vec, void* stream) {
  if (!d_user || !off || !d_hist_vec || B <= 0 || E <= 0) return fail(NRL_ERR_INVALID_ARG, "nrl_late_fusion_bwd: bad argument");
  late_fusion_bwd_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_user, off, B, E, d_hist_vec);
  LAUNCH_CHECK("late_fusion_bwd");
  return NRL_OK;
}

int nrl_score_fwd(const float* user, const float* cand, const int* cand_off, int B, int Cmax, int E,
                  float* scores, void* stream) {
  if (!user || !cand || !cand_off || !scores || B <= 0 || Cmax <= 0 || E <= 0)
    return fail(NRL_ERR_INVALID_ARG, "nrl_score_fwd: bad argument");
  const long long warps = (long long)B * Cmax;
  score_fwd_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      user, cand, cand_off, B, Cmax, E, scores);
  LAUNCH_CHECK("score_fwd");
  return NRL_OK;
}
int nrl_score_bwd(const float* d_scores, const float* user, const float* cand, const int* cand_off,
                  int B, int Cmax, int E, float* d_user, float* d_cand, void* stream) {
  if (!d_scores || !user || !cand || !cand_off || !d_user || !d_cand || B <= 0 || Cmax <= 0 || E <= 0)
    return fail(NRL_ERR_INVALID_ARG, "nrl_score_bwd: bad argument");
  score_bwd_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(d_scores, user, cand, cand_off, B,
                                                                     Cmax, E, d_user, d_cand);
  LAUNCH_CHECK("score_bwd");
  return NRL_OK;
}

int nrl_ce_soft_fwd(const float* scores, const float* labels, const int* cand_off, int B, int Cmax,
                    float* loss_rows, float* loss_mean, float* y_dense, void* stream) {
  if (!scores || !labels || !cand_off || !loss_mean || B <= 0 || Cmax <= 0)
    return fail(NRL_ERR_INVALID_ARG, "nrl_ce_soft_fwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(cudaMemsetAsync(loss_mean, 0, sizeof(float), st));
  ce_fwd_kernel<<<(B + 3) / 4, 128, 0, st>>>(scores, labels, cand_off, B, Cmax, loss_rows, loss_mean, y_dense);
  LAUNCH_CHECK("ce_fwd");
  return NRL_OK;
}
int nrl_ce_soft_bwd(const float* scores, const float* labels, const int* cand_off, int B, int Cmax,
                    const float* g_loss, float g_scale, float* d_scores, void* stream) {
  if (!scores || !labels || !cand_off || !d_scores || B <= 0 || Cmax <= 0)
    return fail(NRL_ERR_INVALID_ARG, "nrl_ce_soft_bwd: bad argument");
  ce_bwd_kernel<<<(B + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(scores, labels, cand_off, B,
                                                                           Cmax, g_loss, g_scale, d_scores);
  LAUNCH_CHECK("ce_bwd");
  return NRL_OK;
}

int nrl_rank_metrics(const float* scores, const float* labels, const long long* off, int B, const int* top_k, int n_k,
                     float* out, int* ranks, void* stream) {
  if (!scores || !labels || !off || !out || B <= 0 || n_k < 0 || n_k > RM_MAXK || (n_k > 0 && !top_k))
    return fail(NRL_ERR_INVALID_ARG, "nrl_rank_metrics: bad argument (at most %d cut-offs)", RM_MAXK);
  RankKs ks;
  ks.n = n_k;
  for (int q = 0; q < RM_MAXK; ++q) ks.k[q] = q < n_k ? top_k[q] : 0;
  for (int q = 0; q < n_k; ++q)
    if (top_k[q] <= 0) return fail(NRL_ERR_INVALID_ARG, "nrl_rank_metrics: top_k must be positive");
  rank_metrics_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(scores, labels, off, B, ks, out, ranks);
  LAUNCH_CHECK("rank_metrics");
  return NRL_OK;
}

int nrl_supcon_fwd(const float* scores, const float* labels, const int* cand_off, int B, int Cmax, float temperature,
                   const float* ce_loss, float dual_loss_coef, float* row_loss, float* loss, float* stats,
                   void* stream) {
  if (!scores || !labels || !cand_off || !row_loss || !loss || !stats || B <= 0 || Cmax <= 0 || !(temperature > 0.f))
    return fail(NRL_ERR_INVALID_ARG, "nrl_supcon_fwd: bad argument");
  const int threads = B >= 32 ? 1024 : 32 * B;
  supcon_fwd_kernel<<<1, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      scores, labels, cand_off, B, Cmax, 1.f / temperature, ce_loss, dual_loss_coef, row_loss, loss, stats);
  LAUNCH_CHECK("supcon_fwd");
  return NRL_OK;
}
int nrl_supcon_bwd(const float* scores, const float* labels, const int* cand_off, int B, int Cmax, float temperature,
                   const float* row_loss, const float* stats, const float* g_loss, float g_scale, int accumulate,
                   float* d_scores, void* stream) {
  if (!scores || !labels || !cand_off || !row_loss || !stats || !d_scores || B <= 0 || Cmax <= 0 || !(temperature > 0.f))
    return fail(NRL_ERR_INVALID_ARG, "nrl_supcon_bwd: bad argument");
  supcon_bwd_kernel<<<(B + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      scores, labels, cand_off, B, Cmax, 1.f / temperature, row_loss, stats, g_loss, g_scale, accumulate, d_scores);
  LAUNCH_CHECK("supcon_bwd");
  return NRL_OK;
}

// torch.optim.Adam keeps its betas as Python doubles: the bias corrections 1 - beta^step and the factor (1 - beta) of the
// moment updates are formed in double and only then rounded to fp32 (1 - 0.999 -> 0.001f), while beta itself multiplies
// the moment as fp32.  This ABI carries fp32 betas, and 1 - 0.999f is 1.3e-5 off 0.001f: the decimal the caller wrote is
// recovered as the <= 7-digit decimal that rounds to the same fp32 (any beta typed as a short decimal; otherwise the fp32
// value itself).
static double beta_as_double(float beta) {
  char buf[32];
  snprintf(buf, sizeof(buf), "%.7g", (double)beta);
  const double d = strtod(buf, nullptr);
  return (float)d == beta ? d : (double)beta;
}

static int adam_impl(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                     float eps, long long step, float grad_scale, int zero_grad, void* stream) {
  if (!p || !g || !m || !v || n <= 0 || step <= 0) return fail(NRL_ERR_INVALID_ARG, "nrl_adam_step: bad argument");
  const double b1d = beta_as_double(beta1), b2d = beta_as_double(beta2);
  const double bc1 = 1.0 - std::pow(b1d, (double)step);
  const double bc2 = 1.0 - std::pow(b2d, (double)step);
  TRY(device_init());
  // grid: 8 CTAs of 256 threads per SM.  Measured on B200: ALONE with a cold L2 (experiments/adam_bw.py,
  // profiles/r02u_adam_bw.txt; 8 streams of 87 / 160 MB) 3 per SM is the fastest (6.2-6.4 TB/s against 6.0 for 8, 5.7-5.8
  // for the occupancy limit of 5); INSIDE the training step, right behind the embedding-gradient scatter whose output is
  // still in L2, 8 per SM takes 0.121 ms and 3 per SM 0.136 ms (profiles/r02p_bench.json / r02_final2_bench.json).  The
  // step is what counts.  NRL_ADAM_CTAS_PER_SM overrides.
  int per_sm = 8;
  if (const char* e = getenv("NRL_ADAM_CTAS_PER_SM")) per_sm = atoi(e) > 0 ? atoi(e) : per_sm;
  adam_kernel<<<grid_for(n, 256 * 4, per_sm * g_dev.sm_count), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      p, g, m, v, n, lr, beta1, beta2, (float)(1.0 - b1d), (float)(1.0 - b2d), eps, (float)bc1, (float)std::sqrt(bc2),
      grad_scale, zero_grad);
  LAUNCH_CHECK("adam");
  return NRL_OK;
}
int nrl_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                  float beta2, float eps, long long step, float grad_scale, void* stream) {
  return adam_impl(p, const_cast<float*>(g), m, v, n, lr, beta1, beta2, eps, step, grad_scale, 0, stream);
}
int nrl_adam_step_zero_grad(float* p, float* g, float* m, float* v, long long n, float lr, float beta1,
                            float beta2, float eps, long long step, float grad_scale, void* stream) {
  return adam_impl(p, g, m, v, n, lr, beta1, beta2, eps, step, grad_scale, 1, stream);
}

int nrl_device_status(int* code_host, void* stream) {
  if (!code_host) return fail(NRL_ERR_INVALID_ARG, "nrl_device_status: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  unsigned int code = 0;
  const unsigned int zero = 0;
  CUDA_TRY(cudaMemcpyFromSymbolAsync(&code, g_dev_error, sizeof(code), 0, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  if (code) CUDA_TRY(cudaMemcpyToSymbolAsync(g_dev_error, &zero, sizeof(zero), 0, cudaMemcpyHostToDevice, st));
  *code_host = (int)code;
  if (code) {
    static const char* what[] = {"", "a token id lies outside [0, V1) (nn.Embedding would raise)",
                                 "segment ids are not sorted ids in [0, B)",
                                 "a segment is longer than the dense width passed as Hmax / Cmax",
                                 "a row index lies outside the table"};
    fail(NRL_ERR_INVALID_ARG, "device-side input check %u: %s", code, code < 5 ? what[code] : "?");
  }
  return NRL_OK;
}

int nrl_embedding_gather(const long long* ids, long long n, const float* table, long long V1, int E,
                         float* out_f32, void* out_hi, void* out_lo, void* stream) {
  if (!ids || !table || n <= 0 || V1 <= 0 || E <= 0 || (!out_f32 && !out_hi))
    return fail(NRL_ERR_INVALID_ARG, "nrl_embedding_gather: bad argument");
  TRY(device_init());
  gather_split_kernel<<<grid_for(n, 8, 1 << 20), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ids, n, table, V1, E, round_up(E + 1, 16), static_cast<bf16*>(out_hi), static_cast<bf16*>(out_lo), out_f32,
      nullptr, 0, 1.f);
  LAUNCH_CHECK("gather_split");
  return NRL_OK;
}

// ----------------------------------------------------------------------------------------
// peer-memory gradient exchange fused with Adam (nrl_exchange.cuh)
// ----------------------------------------------------------------------------------------
static_assert(sizeof(nrl_peer_set) == sizeof(PeerSet), "nrl_peer_set and PeerSet must have one layout");
static_assert(NRL_MAX_RANKS == XCHG_MAX_RANKS && NRL_FLAG_BYTES == XCHG_FLAG_WORDS * 8, "flag block layout");
static_assert(NRL_IPC_HANDLE_BYTES == sizeof(cudaIpcMemHandle_t), "IPC handle size");

int nrl_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle) {
  if (!dev_ptr || !handle || bytes == 0) return fail(NRL_ERR_INVALID_ARG, "nrl_peer_alloc: bad argument");
  void* p = nullptr;
  CUDA_TRY(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(NRL_ERR_CUDA, "nrl_peer_alloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
  }
  std::memcpy(handle, &h, sizeof(h));
  *dev_ptr = p;
  return NRL_OK;
}
int nrl_peer_free(void* dev_ptr) {
  if (!dev_ptr) return NRL_OK;
  CUDA_TRY(cudaFree(dev_ptr));
  return NRL_OK;
}
int nrl_peer_open(const unsigned char* handle, void** dev_ptr) {
  if (!handle || !dev_ptr) return fail(NRL_ERR_INVALID_ARG, "nrl_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  // opened with the CALLER's device current: the driver maps the exporting GPU's allocation and enables
  // peer access from this device to it
  CUDA_TRY(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return NRL_OK;
}
int nrl_peer_close(void* dev_ptr) {
  if (!dev_ptr) return NRL_OK;
  CUDA_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return NRL_OK;
}

int nrl_exchange_adam_step(const nrl_peer_set* peers, float* m, float* v, long long n, float lr, float beta1,
                           float beta2, float eps, long long step, unsigned long long epoch, float grad_scale,
                           int max_ctas, unsigned long long timeout_ns, long long sparse_rows, int row_elems,
                           int zero_grads, void* stream) {
  if (!peers || !m || !v || n <= 0 || step <= 0 || epoch == 0)
    return fail(NRL_ERR_INVALID_ARG, "nrl_exchange_adam_step: bad argument");
  if (peers->world < 1 || peers->world > NRL_MAX_RANKS || peers->rank < 0 || peers->rank >= peers->world)
    return fail(NRL_ERR_INVALID_ARG, "nrl_exchange_adam_step: world %d rank %d", peers->world, peers->rank);
  if (n % 4) return fail(NRL_ERR_INVALID_ARG, "nrl_exchange_adam_step: n = %lld is not a multiple of 4", n);
  uintptr_t align = reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v);
  for (int r = 0; r < peers->world; ++r) {
    if (!peers->params[r] || !peers->grads[r] || !peers->flags[r])
      return fail(NRL_ERR_INVALID_ARG, "nrl_exchange_adam_step: rank %d has a null buffer", r);
    align |= reinterpret_cast<uintptr_t>(peers->params[r]) | reinterpret_cast<uintptr_t>(peers->grads[r]);
    if (reinterpret_cast<uintptr_t>(peers->flags[r]) & 7)
      return fail(NRL_ERR_INVALID_ARG, "nrl_exchange_adam_step: flag block of rank %d is not 8-byte aligned", r);
  }
  if (align & 15) return fail(NRL_ERR_INVALID_ARG, "nrl_exchange_adam_step: buffers must be 16-byte aligned");
  SparseCfg sp;
  sp.rows = 0; sp.row_f4 = 1; sp.bm_words = 0; sp.zero_grads = zero_grads ? 1 : 0;
  if (sparse_rows > 0 && peers->world > 1) {
    if (row_elems <= 0 || (row_elems & 3) || sparse_rows * row_elems > n)
      return fail(NRL_ERR_INVALID_ARG, "nrl_exchange_adam_step: sparse region %lld rows x %d elements does not fit n = %lld "
                  "(row_elems must be a positive multiple of 4)", sparse_rows, row_elems, n);
    for (int r = 0; r < peers->world; ++r)
      if (!peers->bitmaps[r] || (reinterpret_cast<uintptr_t>(peers->bitmaps[r]) & 3))
        return fail(NRL_ERR_INVALID_ARG, "nrl_exchange_adam_step: rank %d has no row-bitmap area", r);
    sp.rows = sparse_rows; sp.row_f4 = row_elems / 4; sp.bm_words = (int)((sparse_rows + 31) / 32);
  }
  TRY(device_init());
  PeerSet ps;
  std::memcpy(&ps, peers, sizeof(ps));
  const double b1d = beta_as_double(beta1), b2d = beta_as_double(beta2);
  const double bc1 = 1.0 - std::pow(b1d, (double)step);
  const double bc2 = 1.0 - std::pow(b2d, (double)step);
  const long long n4 = n / 4, per = (n4 + ps.world - 1) / ps.world;
  if (timeout_ns == 0) timeout_ns = 5000000000ull;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // every CTA waits for a decision that needs ALL CTAs of the grid to have passed step 0: the grid must be
  // co-resident (never more CTAs than the device can hold at once)
#define XCHG_LAUNCH(W)                                                                                       \
  do {                                                                                                       \
    int occ = 0;                                                                                             \
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, exchange_adam_kernel<W>, 256, 0));          \
    if (occ < 1) return fail(NRL_ERR_CUDA, "exchange kernel does not fit an SM");                            \
    int cap = max_ctas > 0 ? max_ctas : 4 * g_dev.sm_count;                                                  \
    if (cap > occ * g_dev.sm_count) cap = occ * g_dev.sm_count;                                              \
    long long work = per > (long long)sp.bm_words * 64 ? per : (long long)sp.bm_words * 64;                  \
    const int grid = grid_for(work, 256, cap);                                                               \
    exchange_adam_kernel<W><<<grid, 256, 0, st>>>(ps, m, v, n4, epoch, timeout_ns, lr, beta1, beta2,         \
                                                  (float)(1.0 - b1d), (float)(1.0 - b2d), eps,                 \
                                                  (float)bc1, (float)std::sqrt(bc2), grad_scale, sp);        \
  } while (0)
  switch (ps.world) {
    case 1: XCHG_LAUNCH(1); break;
    case 2: XCHG_LAUNCH(2); break;
    case 4: XCHG_LAUNCH(4); break;
    case 8: XCHG_LAUNCH(8); break;
    default: XCHG_LAUNCH(0); break;
  }
#undef XCHG_LAUNCH
  LAUNCH_CHECK("exchange_adam");
  return NRL_OK;
}

int nrl_exchange_status(const unsigned long long* flags_local, unsigned long long* error_host, void* stream) {
  if (!flags_local || !error_host) return fail(NRL_ERR_INVALID_ARG, "nrl_exchange_status: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_TRY(cudaMemcpyAsync(error_host, flags_local + XCHG_ERR, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return NRL_OK;
}

// ----------------------------------------------------------------------------------------
// whole NRMS pass
// ----------------------------------------------------------------------------------------
struct NrmsWs {
  long long* ids;      // [N][L] hist then cand
  long long* seg_h;    // staging (host variant)
  long long* seg_c;
  float* labels;
  int *hist_off, *cand_off;
  float* news_vec;     // [N][E]
  float* user_vec;     // [B][E]
  float* d_scores;     // [B][Cmax]
  float* d_user;       // [B][E]
  float* d_news;       // [N][E]
  float* scores_dev;   // [B][Cmax] (host variant)
  float* loss_dev;
  BlockWs news, user;
  BlockWs cand;  // NRL_SPLIT_CAND=1: the candidate titles' own block (news then holds the history titles only)
};

// NRL_SPLIT_CAND=1 (experimental, off by default): the candidate titles (9 % of the title block at 50 + 5 news per
// impression) are encoded apart from the history titles, on a second stream, so that their forward pass runs beneath the
// user encoder's forward chain and their backward pass beneath its backward chain -- 16 small dependent launches that
// leave most SMs idle.  The scorer is the join in both directions.  Same kernels, same keep-bit
// draws (the candidate rows keep their row numbers), same results up to the order of the gradient atomics.
static bool split_cand_on() {
  static const bool v = [] { const char* e = getenv("NRL_SPLIT_CAND"); return e && e[0] == '1'; }();
  return v;
}
struct CandStream {
  cudaStream_t s = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  int dev = -1;
};
static CandStream g_cand;
static int cand_init() {
  int dev = 0;
  CUDA_TRY(cudaGetDevice(&dev));
  if (g_cand.s && g_cand.dev == dev) return NRL_OK;
  CUDA_TRY(cudaStreamCreateWithFlags(&g_cand.s, cudaStreamNonBlocking));
  for (int i = 0; i < 4; ++i) CUDA_TRY(cudaEventCreateWithFlags(&g_cand.ev[i], cudaEventDisableTiming));
  g_cand.dev = dev;
  return NRL_OK;
}
// `to` waits for everything issued on `from` so far
static int cand_order(cudaStream_t from, cudaStream_t to, int ev) {
  CUDA_TRY(cudaEventRecord(g_cand.ev[ev], from));
  CUDA_TRY(cudaStreamWaitEvent(to, g_cand.ev[ev], 0));
  return NRL_OK;
}
static void carve_nrms(Bump& b, long long nh, long long nc, int L, int B, int Hmax, int Cmax,
                       const Dims& d, NrmsWs& w) {
  const long long N = nh + nc;
  w.ids = b.take<long long>((size_t)N * L);
  w.seg_h = b.take<long long>((size_t)nh);
  w.seg_c = b.take<long long>((size_t)nc);
  w.labels = b.take<float>((size_t)nc);
  w.hist_off = b.take<int>(B + 1);
  w.cand_off = b.take<int>(B + 1);
  w.news_vec = b.take<float>((size_t)N * d.E);
  w.user_vec = b.take<float>((size_t)B * d.E);
  w.d_scores = b.take<float>((size_t)B * Cmax);
  w.d_user = b.take<float>((size_t)B * d.E);
  w.d_news = b.take<float>((size_t)N * d.E);
  w.scores_dev = b.take<float>((size_t)B * Cmax);
  w.loss_dev = b.take<float>(1);
  if (split_cand_on()) {
    carve_block(b, nh * L, d, w.news);
    carve_block(b, nc * L, d, w.cand);
  } else {
    carve_block(b, N * L, d, w.news);
  }
  carve_block(b, (long long)B * Hmax, d, w.user);
}

size_t nrl_nrms_ws_bytes(long long n_hist, long long n_cand, int L, int B, int Hmax, int Cmax,
                         nrl_dims dims) {
  Dims d;
  if (make_dims(dims, d) != NRL_OK || n_hist <= 0 || n_cand <= 0 || L <= 0 || B <= 0 || Hmax <= 0 || Cmax <= 0)
    return 0;
  Bump b(nullptr);
  NrmsWs w;
  carve_nrms(b, n_hist, n_cand, L, B, Hmax, Cmax, d, w);
  return b.off + 1024;
}

static int nrms_tail(const Ctx& c, const Dims& d, NrmsWs& w, const float* labels, long long nh, long long nc, int L,
                     int B, int Hmax, int Cmax, long long V1, const nrl_block_params* np, const nrl_block_params* up,
                     int late_fusion, const DropCfg& drop, float* scores, float* loss, const float* g_loss,
                     int do_backward, nrl_block_grads* ng, nrl_block_grads* ug, float* d_table);

static int nrms_impl(const Ctx& c, const Dims& d, NrmsWs& w, const long long* hist_ids,
                     const long long* cand_ids, const long long* seg_hist, const long long* seg_cand,
                     const float* labels, long long nh, long long nc, int L, int B, int Hmax, int Cmax,
                     const float* table, long long V1, const nrl_block_params* np, const nrl_block_params* up,
                     int late_fusion, const DropCfg& drop, float* scores, float* loss, int do_backward,
                     nrl_block_grads* ng, nrl_block_grads* ug, float* d_table) {
  const long long N = nh + nc;
  segment_offsets2_kernel<<<dim3((B + 1 + 127) / 128, 2), 128, 0, c.stream>>>(seg_hist, nh, w.hist_off, Hmax, seg_cand, nc,
                                                                               w.cand_off, Cmax, B);
  LAUNCH_CHECK("segment_offsets");
  if (hist_ids != w.ids)
    CUDA_TRY(cudaMemcpyAsync(w.ids, hist_ids, (size_t)nh * L * sizeof(long long), cudaMemcpyDeviceToDevice, c.stream));
  if (cand_ids != w.ids + nh * L)
    CUDA_TRY(cudaMemcpyAsync(w.ids + nh * L, cand_ids, (size_t)nc * L * sizeof(long long), cudaMemcpyDeviceToDevice, c.stream));
  // history and candidate titles share the news encoder: one pass over all N news
  // the operand copies of BOTH blocks' weights in one launch (the user block is packed while nothing depends on it)
  TRY(pack_weights(c, d, np, w.news, late_fusion ? nullptr : up, late_fusion ? nullptr : &w.user));
  const bool split = split_cand_on();
  const bool par = split && !g_prof.on && cand_init() == NRL_OK;  // under the per-launch profiler everything stays in one stream
  if (!split) {
    TRY(news_fwd_impl(c, d, w.news, w.ids, N, L, table, V1, np, drop, w.news_vec, true));
  } else {
    // the candidate block shares the packed weights of the history block
    w.cand.win_f = w.news.win_f; w.cand.win_t = w.news.win_t; w.cand.wout_f = w.news.wout_f;
    w.cand.wout_t = w.news.wout_t; w.cand.wadd_f = w.news.wadd_f; w.cand.wadd_t = w.news.wadd_t;
    TRY(news_fwd_impl(c, d, w.news, w.ids, nh, L, table, V1, np, drop, w.news_vec, true));
    Ctx cc = c;
    if (par) {
      cc.stream = g_cand.s;
      TRY(cand_order(c.stream, g_cand.s, 0));  // starts when the history titles are done: beneath the user encoder
    }
    TRY(news_fwd_impl(cc, d, w.cand, w.ids + nh * L, nc, L, table, V1, np, drop, w.news_vec + nh * d.E, true, nh * L));
  }
  const long long Ru = (long long)B * Hmax;
  DropCfg nodrop = make_drop(0.f, 0, 0);
  if (!late_fusion) {
    dense_scatter_kernel<<<grid_for(Ru, 1, 1 << 20), 128, 0, c.stream>>>(
        w.news_vec, w.hist_off, B, Hmax, d.E, d.Ep, nullptr, w.user.x,
        c.two_planes() ? w.user.x + Ru * d.Ep : nullptr);
    LAUNCH_CHECK("dense_scatter(hist)");
    TRY(block_forward(c, d, w.user, Ru, user_geom(B, Hmax, 0), B, Hmax, up, nodrop, w.user_vec));
  } else {
    late_fusion_fwd_kernel<<<B, 128, 0, c.stream>>>(w.news_vec, w.hist_off, B, d.E, w.user_vec);
    LAUNCH_CHECK("late_fusion_fwd");
  }
  if (do_backward && (!ng || (!late_fusion && !ug)))
    return fail(NRL_ERR_INVALID_ARG, "backward requested without gradient buffers");
  if (par) TRY(cand_order(g_cand.s, c.stream, 1));  // the scorer needs the candidate vectors
  if (loss) CUDA_TRY(cudaMemsetAsync(loss, 0, sizeof(float), c.stream));
  return nrms_tail(c, d, w, labels, nh, nc, L, B, Hmax, Cmax, V1, np, up, late_fusion, drop, scores, loss, nullptr,
                   do_backward, ng, ug, d_table);
}

// Scorer + soft-target CE and, when do_backward, everything behind them.  Runs on the state the forward half left in the
// workspace (news / user vectors, offsets, ids, packed weights, keep-bit words, the blocks' saved activations), so it is
// also the whole of nrl_nrms_step_bwd.  g_loss (device, may be NULL = 1) scales d loss / d scores.
static int nrms_tail(const Ctx& c, const Dims& d, NrmsWs& w, const float* labels, long long nh, long long nc, int L,
                     int B, int Hmax, int Cmax, long long V1, const nrl_block_params* np, const nrl_block_params* up,
                     int late_fusion, const DropCfg& drop, float* scores, float* loss, const float* g_loss,
                     int do_backward, nrl_block_grads* ng, nrl_block_grads* ug, float* d_table) {
  const long long N = nh + nc;
  const float* cand_vec = w.news_vec + nh * d.E;
  const long long Ru = (long long)B * Hmax;
  DropCfg nodrop = make_drop(0.f, 0, 0);
  if (do_backward)
    // every row is overwritten below for well-formed segment ids; rows of malformed input (flagged by the device-side
    // checks) must not carry stale workspace bytes into the table gradient
    CUDA_TRY(cudaMemsetAsync(w.d_news, 0, (size_t)N * d.E * sizeof(float), c.stream));
  // scorer + soft-target CE (+ their backward) in one launch: CTA b owns impression b
  score_loss_kernel<<<B, 128, 2 * (size_t)Cmax * sizeof(float), c.stream>>>(
      w.user_vec, cand_vec, labels, w.cand_off, B, Cmax, d.E, scores, loss, g_loss, do_backward ? w.d_scores : nullptr,
      w.d_user, w.d_news + nh * d.E);
  LAUNCH_CHECK("score_loss");
  if (!do_backward) return NRL_OK;
  const bool split = split_cand_on();
  const bool par = split && !g_prof.on && cand_init() == NRL_OK;
  if (split) {
    // the candidate titles' backward pass: issued first, on its own stream, it runs beneath the user encoder's backward
    w.cand.win_f = w.news.win_f; w.cand.win_t = w.news.win_t; w.cand.wout_f = w.news.wout_f;
    w.cand.wout_t = w.news.wout_t; w.cand.wadd_f = w.news.wadd_f; w.cand.wadd_t = w.news.wadd_t;
    Ctx cc = c;
    if (par) {
      cc.stream = g_cand.s;
      TRY(cand_order(c.stream, g_cand.s, 2));
    }
    TRY(news_bwd_impl(cc, d, w.cand, w.ids + nh * L, nc, L, V1, np, drop, w.d_news + nh * d.E, ng, d_table));
  }
  if (!late_fusion) {
    TRY(block_backward(c, d, w.user, Ru, user_geom(B, Hmax, 0), B, Hmax, up, nodrop, nodrop, w.d_user, ug, false));
    dim3 grid(Hmax, B < 65535 ? B : 65535);
    dense_gather_kernel<<<grid, 128, 0, c.stream>>>(w.user.dx, w.hist_off, B, Hmax, d.E, w.d_news);
    LAUNCH_CHECK("dense_gather");
  } else {
    late_fusion_bwd_kernel<<<B, 128, 0, c.stream>>>(w.d_user, w.hist_off, B, d.E, w.d_news);
    LAUNCH_CHECK("late_fusion_bwd");
  }
  if (!split) return news_bwd_impl(c, d, w.news, w.ids, N, L, V1, np, drop, w.d_news, ng, d_table);
  TRY(news_bwd_impl(c, d, w.news, w.ids, nh, L, V1, np, drop, w.d_news, ng, d_table));
  if (par) TRY(cand_order(g_cand.s, c.stream, 3));  // the caller's stream continues when both halves are done
  return NRL_OK;
}

static int nrms_check(long long nh, long long nc, int L, int B, int Hmax, int Cmax, const float* table,
                      long long V1, const nrl_block_params* np, const nrl_block_params* up, int late_fusion,
                      float dropout_p) {
  if (nh <= 0 || nc <= 0 || L <= 0 || B <= 0 || Hmax <= 0 || Cmax <= 0 || V1 <= 0)
    return fail(NRL_ERR_INVALID_ARG, "nrl_nrms_step: empty batch or non-positive size");
  if (!table || !np || (!late_fusion && !up)) return fail(NRL_ERR_INVALID_ARG, "nrl_nrms_step: null parameter pointer");
  if (dropout_p < 0.f || dropout_p >= 1.f) return fail(NRL_ERR_INVALID_ARG, "dropout_p out of [0,1)");
  return NRL_OK;
}

int nrl_nrms_step(const long long* hist_ids, const long long* cand_ids, const long long* seg_hist,
                  const long long* seg_cand, const float* labels, long long n_hist, long long n_cand,
                  int L, int B, int Hmax, int Cmax, const float* table, long long V1,
                  const nrl_block_params* news_params, const nrl_block_params* user_params,
                  nrl_dims dims, int late_fusion, float dropout_p, int training,
                  unsigned long long seed, float* scores, float* loss, int do_backward,
                  nrl_block_grads* news_grads, nrl_block_grads* user_grads, float* d_table, void* ws,
                  size_t ws_bytes, int precision, void* stream) {
  Dims d;
  TRY(make_dims(dims, d));
  TRY(nrms_check(n_hist, n_cand, L, B, Hmax, Cmax, table, V1, news_params, user_params, late_fusion, dropout_p));
  if (!hist_ids || !cand_ids || !seg_hist || !seg_cand || !scores || ((loss || do_backward) && !labels))
    return fail(NRL_ERR_INVALID_ARG, "nrl_nrms_step: null input pointer");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_nrms_ws_bytes(n_hist, n_cand, L, B, Hmax, Cmax, dims)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  Bump b(ws);
  NrmsWs w;
  carve_nrms(b, n_hist, n_cand, L, B, Hmax, Cmax, d, w);
  return nrms_impl(c, d, w, hist_ids, cand_ids, seg_hist, seg_cand, labels, n_hist, n_cand, L, B, Hmax,
                   Cmax, table, V1, news_params, user_params, late_fusion,
                   make_drop(dropout_p, training, seed), scores, loss, do_backward, news_grads,
                   user_grads, d_table);
}

int nrl_nrms_step_bwd(const float* labels, const float* g_loss, long long n_hist, long long n_cand, int L, int B, int Hmax,
                      int Cmax, long long V1, const nrl_block_params* news_params, const nrl_block_params* user_params,
                      nrl_dims dims, int late_fusion, float dropout_p, int training, unsigned long long seed,
                      nrl_block_grads* news_grads, nrl_block_grads* user_grads, float* d_table, void* ws,
                      size_t ws_bytes, int precision, void* stream) {
  Dims d;
  TRY(make_dims(dims, d));
  if (n_hist <= 0 || n_cand <= 0 || L <= 0 || B <= 0 || Hmax <= 0 || Cmax <= 0 || V1 <= 0)
    return fail(NRL_ERR_INVALID_ARG, "nrl_nrms_step_bwd: empty batch or non-positive size");
  if (!labels || !news_params || !news_grads || (!late_fusion && (!user_params || !user_grads)))
    return fail(NRL_ERR_INVALID_ARG, "nrl_nrms_step_bwd: null pointer");
  if (dropout_p < 0.f || dropout_p >= 1.f) return fail(NRL_ERR_INVALID_ARG, "dropout_p out of [0,1)");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_nrms_ws_bytes(n_hist, n_cand, L, B, Hmax, Cmax, dims)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  Bump b(ws);
  NrmsWs w;
  carve_nrms(b, n_hist, n_cand, L, B, Hmax, Cmax, d, w);
  return nrms_tail(c, d, w, labels, n_hist, n_cand, L, B, Hmax, Cmax, V1, news_params, user_params, late_fusion,
                   make_drop(dropout_p, training, seed), w.scores_dev, nullptr, g_loss, 1, news_grads, user_grads,
                   d_table);
}

// host buffers -> staging area of `ws` (on copy_stream when given), the step on `stream`, results -> host buffers.
// Nothing here waits for the device.
static int nrms_step_host_enqueue(const long long* hist_ids_host, const long long* cand_ids_host,
                                  const long long* seg_hist_host, const long long* seg_cand_host,
                                  const float* labels_host, long long n_hist, long long n_cand, int L, int B, int Hmax,
                                  int Cmax, const float* table, long long V1, const nrl_block_params* news_params,
                                  const nrl_block_params* user_params, nrl_dims dims, int late_fusion, float dropout_p,
                                  int training, unsigned long long seed, float* scores_host, float* loss_host,
                                  int do_backward, nrl_block_grads* news_grads, nrl_block_grads* user_grads,
                                  float* d_table, void* ws, size_t ws_bytes, int precision, void* copy_stream,
                                  unsigned int* status_host, void* stream) {
  Dims d;
  TRY(make_dims(dims, d));
  TRY(nrms_check(n_hist, n_cand, L, B, Hmax, Cmax, table, V1, news_params, user_params, late_fusion, dropout_p));
  if (!hist_ids_host || !cand_ids_host || !seg_hist_host || !seg_cand_host || !scores_host ||
      ((loss_host || do_backward) && !labels_host))
    return fail(NRL_ERR_INVALID_ARG, "nrl_nrms_step_host: null input pointer");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_nrms_ws_bytes(n_hist, n_cand, L, B, Hmax, Cmax, dims)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  cudaStream_t cs = copy_stream ? static_cast<cudaStream_t>(copy_stream) : c.stream;
  Bump b(ws);
  NrmsWs w;
  carve_nrms(b, n_hist, n_cand, L, B, Hmax, Cmax, d, w);
  CUDA_TRY(cudaMemcpyAsync(w.ids, hist_ids_host, (size_t)n_hist * L * sizeof(long long), cudaMemcpyHostToDevice, cs));
  CUDA_TRY(cudaMemcpyAsync(w.ids + n_hist * L, cand_ids_host, (size_t)n_cand * L * sizeof(long long), cudaMemcpyHostToDevice, cs));
  CUDA_TRY(cudaMemcpyAsync(w.seg_h, seg_hist_host, (size_t)n_hist * sizeof(long long), cudaMemcpyHostToDevice, cs));
  CUDA_TRY(cudaMemcpyAsync(w.seg_c, seg_cand_host, (size_t)n_cand * sizeof(long long), cudaMemcpyHostToDevice, cs));
  if (labels_host)
    CUDA_TRY(cudaMemcpyAsync(w.labels, labels_host, (size_t)n_cand * sizeof(float), cudaMemcpyHostToDevice, cs));
  if (cs != c.stream) {  // the step waits for its inputs; whatever else is queued on `stream` (the optimizer step of the
                         // previous batch) overlaps the copies
    cudaEvent_t copied;
    CUDA_TRY(cudaEventCreateWithFlags(&copied, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(copied, cs);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(c.stream, copied, 0);
    cudaEventDestroy(copied);  // released once the recorded work has completed
    CUDA_TRY(e);
  }
  TRY(nrms_impl(c, d, w, w.ids, w.ids + n_hist * L, w.seg_h, w.seg_c, w.labels, n_hist, n_cand, L, B, Hmax,
                Cmax, table, V1, news_params, user_params, late_fusion, make_drop(dropout_p, training, seed),
                w.scores_dev, loss_host ? w.loss_dev : nullptr, do_backward, news_grads, user_grads, d_table));
  CUDA_TRY(cudaMemcpyAsync(scores_host, w.scores_dev, (size_t)B * Cmax * sizeof(float), cudaMemcpyDeviceToHost, c.stream));
  if (loss_host)
    CUDA_TRY(cudaMemcpyAsync(loss_host, w.loss_dev, sizeof(float), cudaMemcpyDeviceToHost, c.stream));
  if (status_host)
    CUDA_TRY(cudaMemcpyFromSymbolAsync(status_host, g_dev_error, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, c.stream));
  return NRL_OK;
}

int nrl_nrms_step_host(const long long* hist_ids_host, const long long* cand_ids_host,
                       const long long* seg_hist_host, const long long* seg_cand_host,
                       const float* labels_host, long long n_hist, long long n_cand, int L, int B,
                       int Hmax, int Cmax, const float* table, long long V1,
                       const nrl_block_params* news_params, const nrl_block_params* user_params,
                       nrl_dims dims, int late_fusion, float dropout_p, int training,
                       unsigned long long seed, float* scores_host, float* loss_host, int do_backward,
                       nrl_block_grads* news_grads, nrl_block_grads* user_grads, float* d_table,
                       void* ws, size_t ws_bytes, int precision, void* stream) {
  TRY(nrms_step_host_enqueue(hist_ids_host, cand_ids_host, seg_hist_host, seg_cand_host, labels_host, n_hist, n_cand, L,
                             B, Hmax, Cmax, table, V1, news_params, user_params, dims, late_fusion, dropout_p, training,
                             seed, scores_host, loss_host, do_backward, news_grads, user_grads, d_table, ws, ws_bytes,
                             precision, nullptr, nullptr, stream));
  CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  {  // the stream is idle: surface a device-side input violation (bad token / segment ids) of this step
    unsigned int code = 0;
    CUDA_TRY(cudaMemcpyFromSymbol(&code, g_dev_error, sizeof(code)));
    if (code) {
      int dummy = 0;
      nrl_device_status(&dummy, stream);  // formats the message and clears the flag
      return NRL_ERR_INVALID_ARG;
    }
  }
  return NRL_OK;
}

struct StepTicket {
  cudaEvent_t ready;
  unsigned int* status_host;
  void* stream;
};

int nrl_nrms_step_host_begin(const long long* hist_ids_host, const long long* cand_ids_host,
                             const long long* seg_hist_host, const long long* seg_cand_host,
                             const float* labels_host, long long n_hist, long long n_cand, int L, int B,
                             int Hmax, int Cmax, const float* table, long long V1,
                             const nrl_block_params* news_params, const nrl_block_params* user_params,
                             nrl_dims dims, int late_fusion, float dropout_p, int training,
                             unsigned long long seed, float* scores_host, float* loss_host, int do_backward,
                             nrl_block_grads* news_grads, nrl_block_grads* user_grads, float* d_table,
                             void* ws, size_t ws_bytes, int precision, void* copy_stream,
                             unsigned int* status_host, void* stream, void** ticket) {
  if (!ticket) return fail(NRL_ERR_INVALID_ARG, "nrl_nrms_step_host_begin: null ticket pointer");
  *ticket = nullptr;
  TRY(nrms_step_host_enqueue(hist_ids_host, cand_ids_host, seg_hist_host, seg_cand_host, labels_host, n_hist, n_cand, L,
                             B, Hmax, Cmax, table, V1, news_params, user_params, dims, late_fusion, dropout_p, training,
                             seed, scores_host, loss_host, do_backward, news_grads, user_grads, d_table, ws, ws_bytes,
                             precision, copy_stream, status_host, stream));
  cudaEvent_t ready;
  CUDA_TRY(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
  cudaError_t e = cudaEventRecord(ready, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) {
    cudaEventDestroy(ready);
    CUDA_TRY(e);
  }
  *ticket = new StepTicket{ready, status_host, stream};
  return NRL_OK;
}

int nrl_nrms_step_host_end(void* ticket) {
  if (!ticket) return fail(NRL_ERR_INVALID_ARG, "nrl_nrms_step_host_end: null ticket");
  StepTicket* t = static_cast<StepTicket*>(ticket);
  cudaError_t e = cudaEventSynchronize(t->ready);
  cudaEventDestroy(t->ready);
  unsigned int* status_host = t->status_host;
  void* stream = t->stream;
  delete t;
  CUDA_TRY(e);
  if (status_host && *status_host) {
    int dummy = 0;
    nrl_device_status(&dummy, stream);  // formats the message and clears the flag (synchronises: error path only)
    return NRL_ERR_INVALID_ARG;
  }
  return NRL_OK;
}

// ----------------------------------------------------------------------------------------
// NAML: CNN text encoder (text.py:163-176) and category encoder (category.py:73-82)
// ----------------------------------------------------------------------------------------
struct CnnDims {
  int E, F, W, Q, Kc, Kp, Fp, Qp, MW0, MW1;
};
static int make_cnn_dims(nrl_cnn_dims d, CnnDims& o) {
  if (d.embed_dim <= 0 || d.num_filters <= 0 || d.window <= 0 || d.query_dim <= 0)
    return fail(NRL_ERR_INVALID_ARG, "bad CNN dims E=%d F=%d w=%d Q=%d", d.embed_dim, d.num_filters, d.window,
                d.query_dim);
  if (!(d.window & 1)) return fail(NRL_ERR_UNSUPPORTED, "even conv window %d not built (output length != L)", d.window);
  if ((d.embed_dim & 3) || (d.num_filters & 3))
    return fail(NRL_ERR_UNSUPPORTED, "embed_dim and num_filters must be multiples of 4");
  if (d.query_dim > 256) return fail(NRL_ERR_UNSUPPORTED, "query_dim %d > 256", d.query_dim);
  o.E = d.embed_dim; o.F = d.num_filters; o.W = d.window; o.Q = d.query_dim;
  o.Kc = o.W * o.E;
  o.Kp = round_up(o.Kc + 1, 16);
  o.Fp = round_up(o.F + 1, 16);
  o.Qp = round_up(o.Q, 16);
  o.MW0 = (o.E + 31) / 32;
  o.MW1 = (o.F + 31) / 32;
  return NRL_OK;
}
struct CnnWs {
  bf16 *wconv_f, *wconv_t, *wadd_f, *wadd_t;
  bf16* A;    // [2][R][Kp]  im2col of the gathered (dropped-out) rows, ones column at w*E
  float* y;   // [R][F]      relu(conv) after dropout site 1
  bf16* yp;   // [2][R][Fp]
  float* a;   // [R][Q]
  float *s, *w;
  bf16* dap;  // [2][R][Qp]
  bf16* dyp;  // [2][R][Fp]
  float* dA;  // [R][Kc]
  uint32_t *mask0, *mask1;
};
static void carve_cnn(Bump& b, long long R, const CnnDims& d, CnnWs& w) {
  w.wconv_f = b.take<bf16>(2ull * d.F * d.Kp);
  w.wconv_t = b.take<bf16>(2ull * d.Kc * d.Fp);
  w.wadd_f = b.take<bf16>(2ull * d.Q * d.Fp);
  w.wadd_t = b.take<bf16>(2ull * d.F * d.Qp);
  w.A = b.take<bf16>(2ull * R * d.Kp);
  w.y = b.take<float>((size_t)R * d.F);
  w.yp = b.take<bf16>(2ull * R * d.Fp);
  w.a = b.take<float>((size_t)R * d.Q);
  w.s = b.take<float>((size_t)R);
  w.w = b.take<float>((size_t)R);
  w.dap = b.take<bf16>(2ull * R * d.Qp);
  w.dyp = b.take<bf16>(2ull * R * d.Fp);
  w.dA = b.take<float>((size_t)R * d.Kc);
  w.mask0 = b.take<uint32_t>((size_t)R * d.MW0);
  w.mask1 = b.take<uint32_t>((size_t)R * d.MW1);
}

size_t nrl_cnn_encoder_ws_bytes(long long n_news, int L, nrl_cnn_dims dims) {
  CnnDims d;
  if (make_cnn_dims(dims, d) != NRL_OK || n_news <= 0 || L <= 0) return 0;
  Bump b(nullptr);
  CnnWs w;
  carve_cnn(b, n_news * L, d, w);
  return b.off + 1024;
}

int nrl_cnn_encoder_fwd(const long long* ids, long long n_news, int L, const float* table,
                        long long V1, const nrl_cnn_params* prm, nrl_cnn_dims dims,
                        float dropout_p, int training, unsigned long long seed, float* out,
                        void* ws, size_t ws_bytes, int precision, void* stream) {
  CnnDims d;
  TRY(make_cnn_dims(dims, d));
  if (!ids || !table || !prm || !out || n_news <= 0 || L <= 0 || V1 <= 0)
    return fail(NRL_ERR_INVALID_ARG, "nrl_cnn_encoder_fwd: null pointer or empty input");
  if (dropout_p < 0.f || dropout_p >= 1.f) return fail(NRL_ERR_INVALID_ARG, "dropout_p out of [0,1)");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_cnn_encoder_ws_bytes(n_news, L, dims)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  const long long R = n_news * L;
  Bump b(ws);
  CnnWs w;
  carve_cnn(b, R, d, w);
  const int tp = c.two_planes() ? 1 : 0;
  const DropCfg drop = make_drop(dropout_p, training, seed);
  pack_weight_kernel<<<grid_for((long long)d.F * d.Kp + (long long)d.Kc * d.Fp, 256, 4096), 256, 0, c.stream>>>(
      prm->cnn_weight, prm->cnn_bias, d.F, d.Kc, d.Kp, d.Fp, w.wconv_f, w.wconv_t, tp);
  LAUNCH_CHECK("pack_weight(conv)");
  pack_weight_kernel<<<grid_for((long long)d.Q * d.Fp + (long long)d.F * d.Qp, 256, 4096), 256, 0, c.stream>>>(
      prm->add_weight, prm->add_bias, d.Q, d.F, d.Fp, d.Qp, w.wadd_f, w.wadd_t, tp);
  LAUNCH_CHECK("pack_weight(additive)");
  if (drop.on) {
    dropout_site_words_kernel<<<grid_for(R * d.MW0, 256, 16 * g_dev.sm_count), 256, 0, c.stream>>>(
        drop.seed, 0u, drop.thr, R, d.E, d.MW0, w.mask0);
    LAUNCH_CHECK("dropout_words(0)");
    dropout_site_words_kernel<<<grid_for(R * d.MW1, 256, 16 * g_dev.sm_count), 256, 0, c.stream>>>(
        drop.seed, 1u, drop.thr, R, d.F, d.MW1, w.mask1);
    LAUNCH_CHECK("dropout_words(1)");
  }
  gather_im2col_kernel<<<grid_for(R, 8, 1 << 20), 256, 0, c.stream>>>(
      ids, n_news, L, table, V1, d.E, d.W, d.Kp, w.A, tp ? w.A + R * d.Kp : nullptr,
      drop.on ? w.mask0 : nullptr, d.MW0, drop.scale);
  LAUNCH_CHECK("gather_im2col");
  {  // Y = dropout1(relu(A Wc^T + bc)): fp32 + split planes (ones column at F)
    GemmEpi e = epi_none();
    e.relu = 1;
    epi_dropout(e, drop, w.mask1, d.MW1);
    // the keep-bit words of site 1 become "kept AND positive" words in place: the backward pass takes ReLU' and
    // dropout' from them (one word per 32-column chunk) instead of re-reading the fp32 activations
    e.pos_words = w.mask1; e.pos_mw = d.MW1;
    Sinks sk;
    sk.f32 = w.y; sk.ld_f32 = d.F; sk.f32_cols = d.F;
    sk.sp = w.yp; sk.ld_sp = d.Fp; sk.sp_cols = d.Fp; sk.ones_col = d.F;
    TRY(gemm_nt(c, w.A, R, d.Kp, w.wconv_f, d.F, d.Kp, d.Kp, e, sk, "gemm conv"));
  }
  {
    GemmEpi e = epi_none();
    e.qvec = prm->add_query; e.score = w.s;
    Sinks sk;
    sk.f32 = w.a; sk.ld_f32 = d.Q; sk.f32_cols = d.Q;
    TRY(gemm_nt(c, w.yp, R, d.Fp, w.wadd_f, d.Q, d.Fp, d.Fp, e, sk, "gemm additive"));
  }
  return launch_pool_fwd(c, w.s, w.y, d.F, L, n_news, w.w, out);
}

int nrl_cnn_encoder_bwd(const long long* ids, long long n_news, int L, long long V1,
                        const nrl_cnn_params* prm, nrl_cnn_dims dims, float dropout_p,
                        int training, unsigned long long seed, const float* d_out,
                        nrl_cnn_grads* g, float* d_table, void* ws, size_t ws_bytes,
                        int precision, void* stream) {
  CnnDims d;
  TRY(make_cnn_dims(dims, d));
  if (!ids || !prm || !d_out || !g || n_news <= 0 || L <= 0 || V1 <= 0)
    return fail(NRL_ERR_INVALID_ARG, "nrl_cnn_encoder_bwd: null pointer or empty input");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_cnn_encoder_ws_bytes(n_news, L, dims)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  const long long R = n_news * L;
  Bump b(ws);
  CnnWs w;
  carve_cnn(b, R, d, w);
  const int tp = c.two_planes() ? 1 : 0;
  const DropCfg drop = make_drop(dropout_p, training, seed);
  pool_bwd_kernel<<<grid_for(n_news, 1, 8 * g_dev.sm_count), 256, L * sizeof(float), c.stream>>>(
      d_out, w.y, w.w, w.a, prm->add_query, d.F, d.Q, d.Qp, L, n_news, nullptr, w.dap,
      tp ? w.dap + R * d.Qp : nullptr, g->add_query, g->add_bias);
  LAUNCH_CHECK("pool_bwd");
  {  // dPre = relu'(.) dropout1'( w_r dVec + dApre W_add )  -> split planes
    GemmEpi e = epi_none();
    e.add_w = w.w; e.add_vec = d_out; e.ld_addvec = d.F; e.add_L = L;
    // mask1 = keep AND (relu(conv) > 0), written by the forward conv epilogue (all ones kept when dropout is off)
    e.drop_words = w.mask1; e.drop_mw = d.MW1; e.drop_scale = drop.on ? drop.scale : 1.f;
    Sinks sk;
    sk.sp = w.dyp; sk.ld_sp = d.Fp; sk.sp_cols = d.Fp; sk.ones_col = -1;
    TRY(gemm_nt(c, w.dap, R, d.Qp, w.wadd_t, d.F, d.Qp, d.Qp, e, sk, "gemm additive dgrad"));
  }
  TRY(gemm_tn(c, w.dap, d.Q, d.Qp, w.yp, d.Fp, d.Fp, R, g->add_weight, d.F, d.F, nullptr,
              "gemm additive wgrad"));
  // dWc [F][w*E], dbc [F] (the ones column of the im2col operand)
  TRY(gemm_tn(c, w.dyp, d.F, d.Fp, w.A, d.Kp, d.Kp, R, g->cnn_weight, d.Kc, d.Kc, g->cnn_bias,
              "gemm conv wgrad"));
  if (d_table) {
    GemmEpi e = epi_none();
    Sinks sk;
    sk.f32 = w.dA; sk.ld_f32 = d.Kc; sk.f32_cols = d.Kc;
    TRY(gemm_nt(c, w.dyp, R, d.Fp, w.wconv_t, d.Kc, d.Fp, d.Fp, e, sk, "gemm conv dgrad"));
    col2im_emb_grad_kernel<<<grid_for(R, 8, 1 << 20), 256, 0, c.stream>>>(
        ids, n_news, L, V1, w.dA, d.Kc, d.E, d.W, drop.on ? w.mask0 : nullptr, d.MW0, drop.scale, d_table);
    LAUNCH_CHECK("col2im_emb_grad");
  }
  return NRL_OK;
}

struct LinWs {
  bf16 *wf, *wt, *x, *dpre;
  float* dx;
  uint32_t* mask0;
};
static void carve_lin(Bump& b, long long n, int CE, int O, LinWs& w) {
  const int Ep = round_up(CE + 1, 16), Op = round_up(O, 16), MW = (CE + 31) / 32;
  w.wf = b.take<bf16>(2ull * O * Ep);
  w.wt = b.take<bf16>(2ull * CE * Op);
  w.x = b.take<bf16>(2ull * n * Ep);
  w.dpre = b.take<bf16>(2ull * n * Op);
  w.dx = b.take<float>((size_t)n * CE);
  w.mask0 = b.take<uint32_t>((size_t)n * MW);
}
size_t nrl_linear_encoder_ws_bytes(long long n, int embed_dim, int out_dim) {
  if (n <= 0 || embed_dim <= 0 || out_dim <= 0) return 0;
  Bump b(nullptr);
  LinWs w;
  carve_lin(b, n, embed_dim, out_dim, w);
  return b.off + 1024;
}

int nrl_linear_encoder_fwd(const long long* ids, long long n, const float* table, long long V1,
                           int CE, const float* weight, const float* bias, int O, float dropout_p,
                           int training, unsigned long long seed, float* out, void* ws,
                           size_t ws_bytes, int precision, void* stream) {
  if (!ids || !table || !weight || !bias || !out || n <= 0 || V1 <= 0 || CE <= 0 || O <= 0 || (CE & 3) || (O & 3))
    return fail(NRL_ERR_INVALID_ARG, "nrl_linear_encoder_fwd: bad argument (embed_dim, out_dim must be multiples of 4)");
  if (dropout_p < 0.f || dropout_p >= 1.f) return fail(NRL_ERR_INVALID_ARG, "dropout_p out of [0,1)");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_linear_encoder_ws_bytes(n, CE, O)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  Bump b(ws);
  LinWs w;
  carve_lin(b, n, CE, O, w);
  const int tp = c.two_planes() ? 1 : 0;
  const int Ep = round_up(CE + 1, 16), Op = round_up(O, 16), MW = (CE + 31) / 32;
  const DropCfg drop = make_drop(dropout_p, training, seed);
  pack_weight_kernel<<<grid_for((long long)O * Ep + (long long)CE * Op, 256, 4096), 256, 0, c.stream>>>(
      weight, bias, O, CE, Ep, Op, w.wf, w.wt, tp);
  LAUNCH_CHECK("pack_weight(linear)");
  if (drop.on) {
    dropout_site_words_kernel<<<grid_for(n * MW, 256, 16 * g_dev.sm_count), 256, 0, c.stream>>>(
        drop.seed, 0u, drop.thr, n, CE, MW, w.mask0);
    LAUNCH_CHECK("dropout_words(0)");
  }
  gather_split_kernel<<<grid_for(n, 8, 1 << 20), 256, 0, c.stream>>>(
      ids, n, table, V1, CE, Ep, w.x, tp ? w.x + n * Ep : nullptr, nullptr, drop.on ? w.mask0 : nullptr, MW,
      drop.scale);
  LAUNCH_CHECK("gather_split");
  GemmEpi e = epi_none();
  e.relu = 1;
  Sinks sk;
  sk.f32 = out; sk.ld_f32 = O; sk.f32_cols = O;
  return gemm_nt(c, w.x, n, Ep, w.wf, O, Ep, Ep, e, sk, "gemm linear");
}

int nrl_linear_encoder_bwd(const long long* ids, long long n, long long V1, int CE,
                           const float* weight, int O, float dropout_p, int training,
                           unsigned long long seed, const float* out, const float* d_out,
                           float* g_weight, float* g_bias, float* d_table, void* ws,
                           size_t ws_bytes, int precision, void* stream) {
  if (!ids || !weight || !out || !d_out || !g_weight || !g_bias || n <= 0 || CE <= 0 || O <= 0 || (CE & 3) || (O & 3))
    return fail(NRL_ERR_INVALID_ARG, "nrl_linear_encoder_bwd: bad argument");
  if (V1 <= 0) return fail(NRL_ERR_INVALID_ARG, "nrl_linear_encoder_bwd: V1 must be positive");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_linear_encoder_ws_bytes(n, CE, O)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  Bump b(ws);
  LinWs w;
  carve_lin(b, n, CE, O, w);
  const int tp = c.two_planes() ? 1 : 0;
  const int Ep = round_up(CE + 1, 16), Op = round_up(O, 16), MW = (CE + 31) / 32;
  const DropCfg drop = make_drop(dropout_p, training, seed);
  relu_bwd_split_kernel<<<grid_for(n * Op, 256, 8 * g_dev.sm_count), 256, 0, c.stream>>>(
      d_out, out, n, O, Op, w.dpre, tp ? w.dpre + n * Op : nullptr);
  LAUNCH_CHECK("relu_bwd_split");
  TRY(gemm_tn(c, w.dpre, O, Op, w.x, Ep, Ep, n, g_weight, CE, CE, g_bias, "gemm linear wgrad"));
  if (d_table) {
    GemmEpi e = epi_none();
    epi_dropout(e, drop, w.mask0, MW);
    Sinks sk;
    sk.f32 = w.dx; sk.ld_f32 = CE; sk.f32_cols = CE;
    TRY(gemm_nt(c, w.dpre, n, Op, w.wt, CE, Op, Op, e, sk, "gemm linear dgrad"));
    emb_grad_kernel<<<grid_for(n, 8, 1 << 20), 256, 0, c.stream>>>(ids, n, V1, w.dx, CE, d_table);
    LAUNCH_CHECK("emb_grad");
  }
  return NRL_OK;
}

// ----------------------------------------------------------------------------------------
// test helpers
// ----------------------------------------------------------------------------------------
int nrl_dropout_mask(unsigned char* keep, long long n, unsigned long long seed, int site, float p,
                     void* stream) {
  if (!keep || n <= 0 || p < 0.f || p >= 1.f) return fail(NRL_ERR_INVALID_ARG, "nrl_dropout_mask: bad argument");
  dropout_mask_kernel<<<grid_for(n, 256, 1 << 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      keep, n, seed, (uint32_t)site, drop_threshold(p));
  LAUNCH_CHECK("dropout_mask");
  return NRL_OK;
}

size_t nrl_gemm_test_ws_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  Bump b(nullptr);
  // worst case of the two layouts
  const long long big = (long long)(M > K ? M : K) + 16, bigp = round_up((M > K ? M : K), 16) + 16;
  const long long bign = (long long)(N > K ? N : K) + 16;
  b.take<bf16>(2ull * big * bigp);
  b.take<bf16>(2ull * bign * (round_up(N > K ? N : K, 16) + 16));
  b.take<bf16>(2ull * M * round_up(N + 1, 16));  // plane sink of nrl_gemm_test_planes
  return b.off + 1024;
}

int nrl_gemm_test(const float* A, const float* B, float* D, int M, int N, int K, int mn_major,
                  int precision, void* ws, size_t ws_bytes, void* stream) {
  if (!A || !B || !D || M <= 0 || N <= 0 || K <= 0) return fail(NRL_ERR_INVALID_ARG, "nrl_gemm_test: bad argument");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_gemm_test_ws_bytes(M, N, K)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  const int tp = c.two_planes() ? 1 : 0;
  Bump b(ws);
  const long long big = (long long)(M > K ? M : K) + 16, bigp = round_up((M > K ? M : K), 16) + 16;
  const long long bign = (long long)(N > K ? N : K) + 16;
  bf16* ap = b.take<bf16>(2ull * big * bigp);
  bf16* bp = b.take<bf16>(2ull * bign * (round_up(N > K ? N : K, 16) + 16));
  if (!mn_major) {
    const int Kp = round_up(K, 16);
    dense_scatter_kernel<<<grid_for(M, 1, 1 << 20), 128, 0, c.stream>>>(A, nullptr, M, 1, K, Kp, nullptr, ap, tp ? ap + (long long)M * Kp : nullptr);
    LAUNCH_CHECK("split_rows(A)");
    dense_scatter_kernel<<<grid_for(N, 1, 1 << 20), 128, 0, c.stream>>>(B, nullptr, N, 1, K, Kp, nullptr, bp, tp ? bp + (long long)N * Kp : nullptr);
    LAUNCH_CHECK("split_rows(B)");
    // the split kernel writes 1.0 at column K when Kp > K; the tensor-map extent is K, so TMA
    // zero-fills from column K on and the ones column never reaches the MMA.
    GemmEpi e = epi_none();
    Sinks sk;
    sk.f32 = D; sk.ld_f32 = N; sk.f32_cols = N;
    return gemm_nt(c, ap, M, Kp, bp, N, Kp, K, e, sk, "gemm_test nt");
  } else {
    const int Mp = round_up(M, 16), Np = round_up(N, 16);
    dense_scatter_kernel<<<grid_for(K, 1, 1 << 20), 128, 0, c.stream>>>(A, nullptr, K, 1, M, Mp, nullptr, ap, tp ? ap + (long long)K * Mp : nullptr);
    LAUNCH_CHECK("split_rows(A)");
    dense_scatter_kernel<<<grid_for(K, 1, 1 << 20), 128, 0, c.stream>>>(B, nullptr, K, 1, N, Np, nullptr, bp, tp ? bp + (long long)K * Np : nullptr);
    LAUNCH_CHECK("split_rows(B)");
    return gemm_tn(c, ap, M, Mp, bp, N, Np, K, D, N, N, nullptr, "gemm_test tn");
  }
}

int nrl_gemm_test_planes(const float* A, const float* B, float* out, int M, int N, int K, int precision,
                         void* ws, size_t ws_bytes, void* stream) {
  if (!A || !B || !out || M <= 0 || N <= 0 || K <= 0) return fail(NRL_ERR_INVALID_ARG, "nrl_gemm_test_planes: bad argument");
  TRY(device_init());
  TRY(check_common(ws, ws_bytes, nrl_gemm_test_ws_bytes(M, N, K)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  const int tp = c.two_planes() ? 1 : 0;
  Bump b(ws);
  const long long big = (long long)(M > K ? M : K) + 16, bigp = round_up((M > K ? M : K), 16) + 16;
  const long long bign = (long long)(N > K ? N : K) + 16;
  bf16* ap = b.take<bf16>(2ull * big * bigp);
  bf16* bp = b.take<bf16>(2ull * bign * (round_up(N > K ? N : K, 16) + 16));
  const int Np = round_up(N + 1, 16), Kp = round_up(K, 16);
  bf16* dp = b.take<bf16>(2ull * M * Np);
  dense_scatter_kernel<<<grid_for(M, 1, 1 << 20), 128, 0, c.stream>>>(A, nullptr, M, 1, K, Kp, nullptr, ap, tp ? ap + (long long)M * Kp : nullptr);
  LAUNCH_CHECK("split_rows(A)");
  dense_scatter_kernel<<<grid_for(N, 1, 1 << 20), 128, 0, c.stream>>>(B, nullptr, N, 1, K, Kp, nullptr, bp, tp ? bp + (long long)N * Kp : nullptr);
  LAUNCH_CHECK("split_rows(B)");
  GemmEpi e = epi_none();
  Sinks sk;
  sk.sp = dp; sk.ld_sp = Np; sk.sp_cols = Np; sk.ones_col = N;
  TRY(gemm_nt(c, ap, M, Kp, bp, N, Kp, K, e, sk, "gemm_test planes"));
  planes_to_f32_kernel<<<grid_for((long long)M * Np, 256, 1 << 16), 256, 0, c.stream>>>(
      dp, tp ? dp + (long long)M * Np : nullptr, (long long)M * Np, out);
  LAUNCH_CHECK("planes_to_f32");
  return NRL_OK;
}

#include "nrl_tfm_api.cuh"

}  // extern "C"
