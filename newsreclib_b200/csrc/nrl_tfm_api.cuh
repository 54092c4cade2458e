// Host-side orchestration of the PLM transformer (include/nrl.h "PLM news encoder
// internals").  Included by nrl_api.cu inside its extern "C" block: uses its GEMM launchers, Bump carving and
// error plumbing.  Kernels: nrl_tfm.cuh (attention, LayerNorm, embeddings) and nrl_gemm.cuh (every projection).
//
// Per layer, forward (RobertaLayer.forward):                      backward:
//   qkv = xp Wqkv^T                      gemm (fp32 sink)           LN2 bwd -> ds2 (fp32), dt2 planes (keep-bits 2)
//   cp, lse = attention(qkv, kmask)      tfm_attn_fwd               dWo += dt2^T up;  dup = (dt2 Wo) * gelu'(u)
//   s1 = drop1(cp Wao^T) + x             gemm (dropout + residual)  dWi += dup^T h1p; dh1 = dup Wi + ds2
//   h1, h1p = LN1(s1)                    tfm_ln_fwd                 LN1 bwd -> ds1, dt1 planes (keep-bits 1)
//   u, up = h1p Wi^T ; gelu              gemm (fp32 pre-act + gelu planes)   dWao += dt1^T cp; d_o = dt1 Wao
//   s2 = drop2(up Wo^T) + h1             gemm (dropout + residual)  dqkv = attention_bwd
//   x', xp' = LN2(s2)                    tfm_ln_fwd                 dWq/k/v += dqkv^T xp;  dx = dqkv Wqkv + ds1

struct TfmDims {
  int D, H, I, L, V, P, pad, posmode;
  float eps, pdrop, padrop;
  int Dp, Ip, P3, LDQ, MW;
};
static int make_tfm_dims(nrl_tfm_dims d, TfmDims& o) {
  if (d.hidden <= 0 || d.heads <= 0 || d.intermediate <= 0 || d.num_layers < 0 || d.vocab <= 0 || d.max_pos <= 0)
    return fail(NRL_ERR_INVALID_ARG, "bad transformer dims D=%d heads=%d I=%d layers=%d", d.hidden, d.heads,
                d.intermediate, d.num_layers);
  if (d.hidden != d.heads * TFM_DH) return fail(NRL_ERR_UNSUPPORTED, "transformer head dim %d not built (64)",
                                                 d.hidden / d.heads);
  if (d.hidden > 128 * TFM_LN_CHUNKS) return fail(NRL_ERR_UNSUPPORTED, "hidden size %d > %d", d.hidden, 128 * TFM_LN_CHUNKS);
  if (d.intermediate % 16) return fail(NRL_ERR_UNSUPPORTED, "intermediate size must be a multiple of 16");
  if (d.hidden_dropout < 0.f || d.hidden_dropout >= 1.f || d.attn_dropout < 0.f || d.attn_dropout >= 1.f)
    return fail(NRL_ERR_INVALID_ARG, "dropout probabilities out of [0,1)");
  o.D = d.hidden; o.H = d.heads; o.I = d.intermediate; o.L = d.num_layers; o.V = d.vocab; o.P = d.max_pos;
  if (d.position_mode != 0 && d.position_mode != 1)
    return fail(NRL_ERR_INVALID_ARG, "position_mode must be 0 (RoBERTa) or 1 (BERT), got %d", d.position_mode);
  if (d.pad_idx < 0 || d.pad_idx >= d.vocab) return fail(NRL_ERR_INVALID_ARG, "pad_idx %d outside the vocabulary", d.pad_idx);
  o.posmode = d.position_mode;
  o.pad = d.pad_idx; o.eps = d.ln_eps; o.pdrop = d.hidden_dropout; o.padrop = d.attn_dropout;
  o.Dp = round_up(o.D + 1, 16);
  o.Ip = round_up(o.I + 1, 16);
  o.P3 = 3 * o.D;
  o.LDQ = round_up(3 * o.D, 32);
  o.MW = (o.D + 31) / 32;
  return NRL_OK;
}

struct TfmPack {  // one layer's GEMM operands
  bf16 *wqkv_f, *wqkv_t, *wao_f, *wao_t, *wi_f, *wi_t, *wo_f, *wo_t;
};
static void carve_tfm_pack(Bump& b, const TfmDims& d, TfmPack& w) {
  w.wqkv_f = b.take<bf16>(2ull * 3 * d.D * d.Dp);
  w.wqkv_t = b.take<bf16>(2ull * d.D * d.P3);
  w.wao_f = b.take<bf16>(2ull * d.D * d.Dp);
  w.wao_t = b.take<bf16>(2ull * d.D * d.D);
  w.wi_f = b.take<bf16>(2ull * d.I * d.Dp);
  w.wi_t = b.take<bf16>(2ull * d.D * d.I);
  w.wo_f = b.take<bf16>(2ull * d.D * d.Ip);
  w.wo_t = b.take<bf16>(2ull * d.I * d.D);
}
struct TfmLayerWs {  // kept for the backward pass
  bf16* xp;    // [2][R][Dp] layer input planes (ones column at D)
  float* qkv;  // [R][LDQ]
  float* lse;  // [R][H]
  bf16* cp;    // [2][R][Dp] attention context
  float* s1;   // [R][D] drop1(cp Wao^T + b) + x      (pre-LN1)
  bf16* h1p;   // [2][R][Dp] LN1 output
  float* u;    // [R][I] pre-GELU
  bf16* up;    // [2][R][Ip] gelu(u), ones column at I
  float* s2;   // [R][D] drop2(up Wo^T + b) + h1      (pre-LN2)
  uint32_t *mask1, *mask2;  // [R][MW]
};
struct TfmWs {
  int* pos; unsigned char* kmask;
  uint32_t *mask_e, *mask_e1;
  float *xa, *xb, *h1;       // fp32 residual stream (ping-pong) and the LN1 output of the current layer
  float *ga, *gb, *ds, *d_o;  // backward: layer gradient ping-pong, LN-backward residual branch, attention dO
  float* delta;               // [R][H] dO . ctx per row and head (attention backward beyond 128 tokens)
  bf16 *dtp, *dup, *dqkv;     // backward GEMM operands
  std::vector<TfmLayerWs> layer;
};
// keep == false (inference: no backward pass will follow): every layer works in the SAME activation buffers -- each
// is produced and consumed inside its layer, and LayerNorm 2 writes the next layer's input planes when the current
// layer's are no longer read -- so the workspace does not grow with the number of layers
static void carve_tfm(Bump& b, long long N, int T, const TfmDims& d, TfmWs& w, bool keep = true) {
  const long long R = N * T;
  w.pos = b.take<int>((size_t)R);
  w.kmask = b.take<unsigned char>((size_t)R);
  w.mask_e = b.take<uint32_t>((size_t)R * d.MW);
  w.mask_e1 = b.take<uint32_t>((size_t)R * d.MW);
  w.xa = b.take<float>((size_t)R * d.D);
  w.xb = b.take<float>((size_t)R * d.D);
  w.h1 = b.take<float>((size_t)R * d.D);
  w.ga = b.take<float>((size_t)R * d.D);
  w.gb = b.take<float>((size_t)R * d.D);
  w.ds = b.take<float>((size_t)R * d.D);
  w.d_o = b.take<float>((size_t)R * d.D);
  w.delta = b.take<float>((size_t)R * d.H);
  w.dtp = b.take<bf16>(2ull * R * d.D);
  w.dup = b.take<bf16>(2ull * R * d.I);
  w.dqkv = b.take<bf16>(2ull * R * d.P3);
  w.layer.resize(d.L);
  for (int l = 0; l < d.L; ++l) {
    TfmLayerWs& y = w.layer[l];
    if (!keep && l > 0) { y = w.layer[0]; continue; }
    y.xp = b.take<bf16>(2ull * R * d.Dp);
    y.qkv = b.take<float>((size_t)R * d.LDQ);
    y.lse = b.take<float>((size_t)R * d.H);
    y.cp = b.take<bf16>(2ull * R * d.Dp);
    y.s1 = b.take<float>((size_t)R * d.D);
    y.h1p = b.take<bf16>(2ull * R * d.Dp);
    y.u = b.take<float>((size_t)R * d.I);
    y.up = b.take<bf16>(2ull * R * d.Ip);
    y.s2 = b.take<float>((size_t)R * d.D);
    y.mask1 = b.take<uint32_t>((size_t)R * d.MW);
    y.mask2 = b.take<uint32_t>((size_t)R * d.MW);
  }
}
static unsigned long long tfm_layer_seed(unsigned long long seed, int l) {
  return seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(l + 1);
}

static int tfm_attn_attrs_init() {
  static bool done = false;
  if (done) return NRL_OK;
  CUDA_TRY(cudaFuncSetAttribute(tfm_attn_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tfm_attn_fwd_smem(32)));
  CUDA_TRY(cudaFuncSetAttribute(tfm_attn_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tfm_attn_fwd_smem(64)));
  CUDA_TRY(cudaFuncSetAttribute(tfm_attn_fwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tfm_attn_fwd_smem(96)));
  CUDA_TRY(cudaFuncSetAttribute(tfm_attn_fwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tfm_attn_fwd_smem(128)));
  CUDA_TRY(cudaFuncSetAttribute(tfm_attn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tfm_attn_bwd_smem(32)));
  CUDA_TRY(cudaFuncSetAttribute(tfm_attn_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tfm_attn_bwd_smem(64)));
  CUDA_TRY(cudaFuncSetAttribute(tfm_attn_bwd_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tfm_attn_bwd_smem(96)));
  CUDA_TRY(cudaFuncSetAttribute(tfm_attn_bwd_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, tfm_attn_bwd_smem(128)));
  TRY(attn_attrs_init());  // the flash kernels take over beyond 128 tokens
  done = true;
  return NRL_OK;
}

static int tfm_attn_fwd(const Ctx& c, const TfmDims& d, const TfmWs& w, const TfmLayerWs& y, int N, int T,
                        const DropCfg& adrop, unsigned long long lseed) {
  const long long R = (long long)N * T;
  const int nk32 = (T + 31) / 32, SK = 32 * nk32, threads = 64 * nk32;
  bf16* lo = c.two_planes() ? y.cp + R * d.Dp : nullptr;
  const float scale = 1.0f / sqrtf((float)TFM_DH);
  if (T > 128) {  // flash-style kernels (nrl_attn_flash.cuh) with the key-padding mask and the same dropout bits
    const int nqb = (T + 127) / 128;
    attn_fwd_flash_kernel<TFM_DH><<<(unsigned)((long long)N * d.H * nqb), 256, flash_fwd_smem<TFM_DH>(), c.stream>>>(
        y.qkv, d.D, d.LDQ, d.H, T, 1, N, T, scale, y.cp, lo, d.Dp, y.lse, c.two_planes() ? 1 : 0, w.kmask, adrop.on,
        adrop.thr, adrop.scale, lseed, 2u);
    LAUNCH_CHECK("tfm attn_fwd (flash)");
    return NRL_OK;
  }
#define NRL_TFM_FWD(NK)                                                                                          \
  tfm_attn_fwd_kernel<NK><<<(unsigned)(N * d.H), threads, tfm_attn_fwd_smem(SK), c.stream>>>(                    \
      y.qkv, d.LDQ, d.D, d.H, T, w.kmask, scale, y.cp, lo, d.Dp, y.lse, c.two_planes() ? 1 : 0, adrop.on, adrop.thr, \
      adrop.scale, lseed, 2u)
  if (nk32 == 1) NRL_TFM_FWD(1);
  else if (nk32 == 2) NRL_TFM_FWD(2);
  else if (nk32 == 3) NRL_TFM_FWD(3);
  else NRL_TFM_FWD(4);
#undef NRL_TFM_FWD
  LAUNCH_CHECK("tfm attn_fwd");
  return NRL_OK;
}
static int tfm_attn_bwd(const Ctx& c, const TfmDims& d, const TfmWs& w, const TfmLayerWs& y, int N, int T,
                        const DropCfg& adrop, unsigned long long lseed) {
  const long long R = (long long)N * T;
  const int nk32 = (T + 31) / 32, SK = 32 * nk32, threads = 64 * nk32;
  const bf16* olo = c.two_planes() ? y.cp + R * d.Dp : nullptr;
  bf16* glo = c.two_planes() ? w.dqkv + R * d.P3 : nullptr;
  const float scale = 1.0f / sqrtf((float)TFM_DH);
  if (T > 128) {
    const int nqb = (T + 127) / 128;
    const unsigned grid = (unsigned)((long long)N * d.H * nqb);
    attn_delta_kernel<<<grid_for(R, 8, 8 * g_dev.sm_count), 256, 0, c.stream>>>(w.d_o, d.D, y.cp, olo, d.Dp, R, d.H, TFM_DH,
                                                                             w.delta);
    LAUNCH_CHECK("tfm attn_delta");
    attn_bwd_flash_dq_kernel<TFM_DH><<<grid, 256, flash_bwd_smem<TFM_DH>(), c.stream>>>(
        y.qkv, w.d_o, d.D, y.lse, w.delta, d.D, d.LDQ, d.H, T, 1, N, T, scale, w.dqkv, glo, d.P3, c.two_planes() ? 1 : 0,
        w.kmask, adrop.on, adrop.thr, adrop.scale, lseed, 2u);
    LAUNCH_CHECK("tfm attn_bwd dq (flash)");
    attn_bwd_flash_dkv_kernel<TFM_DH><<<grid, 256, flash_bwd_smem<TFM_DH>(), c.stream>>>(
        y.qkv, w.d_o, d.D, y.lse, w.delta, d.D, d.LDQ, d.H, T, 1, N, T, scale, w.dqkv, glo, d.P3, c.two_planes() ? 1 : 0,
        w.kmask, adrop.on, adrop.thr, adrop.scale, lseed, 2u);
    LAUNCH_CHECK("tfm attn_bwd dkv (flash)");
    return NRL_OK;
  }
#define NRL_TFM_BWD(NK)                                                                                          \
  tfm_attn_bwd_kernel<NK><<<(unsigned)(N * d.H), threads, tfm_attn_bwd_smem(SK), c.stream>>>(                    \
      y.qkv, d.LDQ, d.D, d.H, T, w.kmask, scale, w.d_o, y.cp, olo, d.Dp, y.lse, w.dqkv, glo, d.P3,               \
      c.two_planes() ? 1 : 0, adrop.on, adrop.thr, adrop.scale, lseed, 2u)
  if (nk32 == 1) NRL_TFM_BWD(1);
  else if (nk32 == 2) NRL_TFM_BWD(2);
  else if (nk32 == 3) NRL_TFM_BWD(3);
  else NRL_TFM_BWD(4);
#undef NRL_TFM_BWD
  LAUNCH_CHECK("tfm attn_bwd");
  return NRL_OK;
}

static int tfm_check(const char* who, int N, int T, const TfmDims& d) {
  if (N <= 0 || T <= 0) return fail(NRL_ERR_INVALID_ARG, "%s: empty input", who);
  if (T > d.P) return fail(NRL_ERR_INVALID_ARG, "%s: %d tokens per text > max_position_embeddings %d", who, T, d.P);
  if (d.L <= 0) return fail(NRL_ERR_INVALID_ARG, "%s: num_layers must be positive", who);
  return NRL_OK;
}

size_t nrl_tfm_wpack_bytes(nrl_tfm_dims dims) {
  TfmDims d;
  if (make_tfm_dims(dims, d) != NRL_OK) return 0;
  Bump b(nullptr);
  TfmPack w;
  for (int l = 0; l < d.L; ++l) carve_tfm_pack(b, d, w);
  return b.off + 1024;
}

int nrl_tfm_pack_weights(const nrl_tfm_layer_params* layers, int first, int count, nrl_tfm_dims dims, void* wpack,
                         size_t wpack_bytes, int precision, void* stream) {
  TfmDims d;
  TRY(make_tfm_dims(dims, d));
  if (!layers || first < 0 || count < 0 || first + count > d.L)
    return fail(NRL_ERR_INVALID_ARG, "nrl_tfm_pack_weights: layers [%d, %d) outside [0, %d)", first, first + count, d.L);
  TRY(device_init());
  TRY(check_common(wpack, wpack_bytes, nrl_tfm_wpack_bytes(dims)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  Bump b(wpack);
  for (int l = 0; l < d.L; ++l) {
    TfmPack w;
    carve_tfm_pack(b, d, w);
    if (l < first || l >= first + count) continue;
    const nrl_tfm_layer_params& p = layers[l - first];
    TfmPackJobs jobs;
    memset(&jobs, 0, sizeof(jobs));
    jobs.j[0] = TfmPackJob{p.q_w, p.q_b, d.D, d.D, d.Dp, d.P3, 0, 3 * d.D, w.wqkv_f, w.wqkv_t};
    jobs.j[1] = TfmPackJob{p.k_w, p.k_b, d.D, d.D, d.Dp, d.P3, d.D, 3 * d.D, w.wqkv_f, w.wqkv_t};
    jobs.j[2] = TfmPackJob{p.v_w, p.v_b, d.D, d.D, d.Dp, d.P3, 2 * d.D, 3 * d.D, w.wqkv_f, w.wqkv_t};
    jobs.j[3] = TfmPackJob{p.ao_w, p.ao_b, d.D, d.D, d.Dp, d.D, 0, d.D, w.wao_f, w.wao_t};
    jobs.j[4] = TfmPackJob{p.i_w, p.i_b, d.I, d.D, d.Dp, d.I, 0, d.I, w.wi_f, w.wi_t};
    jobs.j[5] = TfmPackJob{p.o_w, p.o_b, d.D, d.I, d.Ip, d.D, 0, d.D, w.wo_f, w.wo_t};
    for (int j = 0; j < 6; ++j)
      if (!jobs.j[j].W) return fail(NRL_ERR_INVALID_ARG, "nrl_tfm_pack_weights: NULL weight in layer %d", l);
    tfm_pack_kernel<<<dim3((unsigned)grid_for((long long)d.I * d.Dp, 256 * 8, 1024), 6, 2), 256, 0, c.stream>>>(
        jobs, c.two_planes() ? 1 : 0);
    LAUNCH_CHECK("tfm pack_weights");
  }
  return NRL_OK;
}

size_t nrl_tfm_ws_bytes(long long N, int T, nrl_tfm_dims dims, int keep_activations) {
  TfmDims d;
  if (make_tfm_dims(dims, d) != NRL_OK || N <= 0 || T <= 0) return 0;
  Bump b(nullptr);
  TfmWs w;
  carve_tfm(b, N, T, d, w, keep_activations != 0);
  return b.off + 1024;
}

static int tfm_dropout_words(const Ctx& c, const TfmDims& d, long long R, const DropCfg& drop, unsigned long long seed,
                             uint32_t* w0, uint32_t* w1) {
  dropout_words_kernel<<<grid_for(2 * R * d.MW, 256, 16 * g_dev.sm_count), 256, 0, c.stream>>>(seed, drop.thr, R, d.D,
                                                                                              d.MW, w0, w1);
  LAUNCH_CHECK("dropout_words");
  return NRL_OK;
}

int nrl_tfm_encoder_fwd(const long long* input_ids, const long long* attention_mask, int N, int T,
                        const nrl_tfm_embed_params* embed, const nrl_tfm_layer_params* layers, nrl_tfm_dims dims,
                        int training, unsigned long long seed, const void* wpack, float* out, int keep_activations,
                        void* ws, size_t ws_bytes, int precision, void* stream) {
  TfmDims d;
  TRY(make_tfm_dims(dims, d));
  TRY(tfm_check("nrl_tfm_encoder_fwd", N, T, d));
  if (!input_ids || !embed || !layers || !wpack || !out)
    return fail(NRL_ERR_INVALID_ARG, "nrl_tfm_encoder_fwd: null pointer");
  TRY(device_init());
  TRY(tfm_attn_attrs_init());
  TRY(check_common(ws, ws_bytes, nrl_tfm_ws_bytes(N, T, dims, keep_activations)));
  if (reinterpret_cast<uintptr_t>(wpack) & 1023) return fail(NRL_ERR_INVALID_ARG, "wpack must be 1024-byte aligned");
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  const long long R = (long long)N * T;
  Bump b(ws);
  TfmWs w;
  carve_tfm(b, N, T, d, w, keep_activations != 0);
  Bump bp(const_cast<void*>(wpack));
  const DropCfg hdrop = make_drop(d.pdrop, training, seed), adrop = make_drop(d.padrop, training, seed);
  const bool two = c.two_planes();
  tfm_prepare_kernel<<<grid_for(N, 8, 4 * g_dev.sm_count), 256, 0, c.stream>>>(
      input_ids, attention_mask, N, T, d.posmode == 0 ? d.pad : -1, w.pos, w.kmask);
  LAUNCH_CHECK("tfm prepare");
  if (hdrop.on) TRY(tfm_dropout_words(c, d, R, hdrop, seed, w.mask_e, w.mask_e1));
  float* x = w.xa;      // fp32 residual stream entering the layer
  float* xn = w.xb;
  tfm_embed_fwd_kernel<<<grid_for(R, 8, 8 * g_dev.sm_count), 256, 0, c.stream>>>(
      input_ids, w.pos, R, embed->word, d.V, embed->pos, d.P, embed->type0, d.D, d.Dp, embed->ln_g, embed->ln_b, d.eps,
      hdrop.on ? w.mask_e : nullptr, d.MW, hdrop.scale, x, w.layer[0].xp, two ? w.layer[0].xp + R * d.Dp : nullptr);
  LAUNCH_CHECK("tfm embed_fwd");
  for (int l = 0; l < d.L; ++l) {
    TfmLayerWs& y = w.layer[l];
    TfmPack pk;
    carve_tfm_pack(bp, d, pk);
    const nrl_tfm_layer_params& p = layers[l];
    const unsigned long long lseed = tfm_layer_seed(seed, l);
    if (hdrop.on) TRY(tfm_dropout_words(c, d, R, hdrop, lseed, y.mask1, y.mask2));
    {  // qkv = x Wqkv^T + b
      GemmEpi e = epi_none();
      Sinks sk;
      sk.f32 = y.qkv; sk.ld_f32 = d.LDQ; sk.f32_cols = 3 * d.D;
      TRY(gemm_nt(c, y.xp, R, d.Dp, pk.wqkv_f, 3 * d.D, d.Dp, d.Dp, e, sk, "tfm gemm qkv"));
    }
    TRY(tfm_attn_fwd(c, d, w, y, N, T, adrop, lseed));
    {  // s1 = drop1(ctx Wao^T + b) + x
      GemmEpi e = epi_none();
      epi_dropout(e, hdrop, y.mask1, d.MW);
      e.add_mat = x; e.ld_addmat = d.D;
      Sinks sk;
      sk.f32 = y.s1; sk.ld_f32 = d.D; sk.f32_cols = d.D;
      TRY(gemm_nt(c, y.cp, R, d.Dp, pk.wao_f, d.D, d.Dp, d.Dp, e, sk, "tfm gemm attn_out"));
    }
    tfm_ln_fwd_kernel<<<grid_for(R, 8, 8 * g_dev.sm_count), 256, 0, c.stream>>>(
        y.s1, R, d.D, d.Dp, p.ln1_g, p.ln1_b, d.eps, w.h1, y.h1p, two ? y.h1p + R * d.Dp : nullptr);
    LAUNCH_CHECK("tfm ln1_fwd");
    {  // u = h1 Wi^T + b (fp32, kept for the backward pass); up = gelu(u) planes
      GemmEpi e = epi_none();
      e.gelu = 1;
      Sinks sk;
      sk.f32 = y.u; sk.ld_f32 = d.I; sk.f32_cols = d.I;
      sk.sp = y.up; sk.ld_sp = d.Ip; sk.sp_cols = d.Ip; sk.ones_col = d.I;
      TRY(gemm_nt(c, y.h1p, R, d.Dp, pk.wi_f, d.I, d.Dp, d.Dp, e, sk, "tfm gemm ffn_in"));
    }
    {  // s2 = drop2(up Wo^T + b) + h1
      GemmEpi e = epi_none();
      epi_dropout(e, hdrop, y.mask2, d.MW);
      e.add_mat = w.h1; e.ld_addmat = d.D;
      Sinks sk;
      sk.f32 = y.s2; sk.ld_f32 = d.D; sk.f32_cols = d.D;
      TRY(gemm_nt(c, y.up, R, d.Ip, pk.wo_f, d.D, d.Ip, d.Ip, e, sk, "tfm gemm ffn_out"));
    }
    const bool last = l == d.L - 1;
    bf16* nxt = last ? nullptr : w.layer[l + 1].xp;
    tfm_ln_fwd_kernel<<<grid_for(R, 8, 8 * g_dev.sm_count), 256, 0, c.stream>>>(
        y.s2, R, d.D, d.Dp, p.ln2_g, p.ln2_b, d.eps, last ? out : xn, nxt, (nxt && two) ? nxt + R * d.Dp : nullptr);
    LAUNCH_CHECK("tfm ln2_fwd");
    float* t = x; x = xn; xn = t;
  }
  return NRL_OK;
}

static bool tfm_layer_trainable(const nrl_tfm_layer_grads* g) {
  return g && (g->q_w || g->k_w || g->v_w || g->ao_w || g->i_w || g->o_w || g->ln1_g || g->ln2_g);
}

int nrl_tfm_encoder_bwd(const long long* input_ids, const long long* attention_mask, int N, int T,
                        const nrl_tfm_embed_params* embed, const nrl_tfm_layer_params* layers, nrl_tfm_dims dims,
                        int training, unsigned long long seed, const void* wpack, const float* d_out,
                        const nrl_tfm_embed_grads* embed_grads, const nrl_tfm_layer_grads* layer_grads, void* ws,
                        size_t ws_bytes, int precision, void* stream) {
  (void)attention_mask;  // the key-padding bytes of the forward call are still in the workspace
  TfmDims d;
  TRY(make_tfm_dims(dims, d));
  TRY(tfm_check("nrl_tfm_encoder_bwd", N, T, d));
  if (!input_ids || !embed || !layers || !wpack || !d_out)
    return fail(NRL_ERR_INVALID_ARG, "nrl_tfm_encoder_bwd: null pointer");
  TRY(device_init());
  TRY(tfm_attn_attrs_init());
  TRY(check_common(ws, ws_bytes, nrl_tfm_ws_bytes(N, T, dims, 1)));
  Ctx c{static_cast<cudaStream_t>(stream), precision};
  const long long R = (long long)N * T;
  Bump b(ws);
  TfmWs w;
  carve_tfm(b, N, T, d, w);
  std::vector<TfmPack> packs(d.L);
  {
    Bump bp(const_cast<void*>(wpack));
    for (int l = 0; l < d.L; ++l) carve_tfm_pack(bp, d, packs[l]);
  }
  const DropCfg hdrop = make_drop(d.pdrop, training, seed), adrop = make_drop(d.padrop, training, seed);
  const bool two = c.two_planes();
  bf16* dtp_lo = two ? w.dtp + R * d.D : nullptr;
  // stop at the lowest layer that still needs a gradient: below it nothing is trainable unless the embeddings are
  int lowest = 0;
  if (!embed_grads) {
    lowest = d.L;
    for (int l = 0; l < d.L; ++l)
      if (tfm_layer_trainable(layer_grads ? &layer_grads[l] : nullptr)) { lowest = l; break; }
  }
  const float* dy = d_out;
  for (int l = d.L - 1; l >= lowest; --l) {
    const TfmLayerWs& y = w.layer[l];
    const TfmPack& pk = packs[l];
    const nrl_tfm_layer_params& p = layers[l];
    const nrl_tfm_layer_grads* g = (layer_grads && tfm_layer_trainable(&layer_grads[l])) ? &layer_grads[l] : nullptr;
    if (g && !(g->q_w && g->q_b && g->k_w && g->k_b && g->v_w && g->v_b && g->ao_w && g->ao_b && g->ln1_g && g->ln1_b &&
               g->i_w && g->i_b && g->o_w && g->o_b && g->ln2_g && g->ln2_b))
      return fail(NRL_ERR_INVALID_ARG, "nrl_tfm_encoder_bwd: layer %d is partly frozen (all 16 gradients or none)", l);
    const unsigned long long lseed = tfm_layer_seed(seed, l);
    const bool need_dx = l > lowest || embed_grads;  // the data gradient below the lowest trainable layer is not needed
    // LN2 backward
    if (g)
      tfm_ln_bwd_kernel<true><<<grid_for(R, 4, 6 * g_dev.sm_count), 128, 2 * d.D * sizeof(float), c.stream>>>(
          y.s2, dy, R, d.D, p.ln2_g, d.eps, w.ds, w.dtp, dtp_lo, d.D, hdrop.on ? y.mask2 : nullptr, d.MW, hdrop.scale,
          g->ln2_g, g->ln2_b);
    else
      tfm_ln_bwd_kernel<false><<<grid_for(R, 4, 16 * g_dev.sm_count), 128, 0, c.stream>>>(
          y.s2, dy, R, d.D, p.ln2_g, d.eps, w.ds, w.dtp, dtp_lo, d.D, hdrop.on ? y.mask2 : nullptr, d.MW, hdrop.scale,
          nullptr, nullptr);
    LAUNCH_CHECK("tfm ln2_bwd");
    if (g) TRY(gemm_tn(c, w.dtp, d.D, d.D, y.up, d.Ip, d.Ip, R, g->o_w, d.I, d.I, g->o_b, "tfm gemm ffn_out wgrad"));
    {  // dup = (dt2 Wo) * gelu'(u)
      GemmEpi e = epi_none();
      e.gelu_pre = y.u; e.ld_gelu = d.I;
      Sinks sk;
      sk.sp = w.dup; sk.ld_sp = d.I; sk.sp_cols = d.I; sk.ones_col = -1;
      TRY(gemm_nt(c, w.dtp, R, d.D, pk.wo_t, d.I, d.D, d.D, e, sk, "tfm gemm ffn_out dgrad"));
    }
    if (g) TRY(gemm_tn(c, w.dup, d.I, d.I, y.h1p, d.Dp, d.Dp, R, g->i_w, d.D, d.D, g->i_b, "tfm gemm ffn_in wgrad"));
    {  // dh1 = dup Wi + ds2
      GemmEpi e = epi_none();
      e.add_mat = w.ds; e.ld_addmat = d.D;
      Sinks sk;
      sk.f32 = w.gb; sk.ld_f32 = d.D; sk.f32_cols = d.D;
      TRY(gemm_nt(c, w.dup, R, d.I, pk.wi_t, d.D, d.I, d.I, e, sk, "tfm gemm ffn_in dgrad"));
    }
    // LN1 backward
    if (g)
      tfm_ln_bwd_kernel<true><<<grid_for(R, 4, 6 * g_dev.sm_count), 128, 2 * d.D * sizeof(float), c.stream>>>(
          y.s1, w.gb, R, d.D, p.ln1_g, d.eps, w.ds, w.dtp, dtp_lo, d.D, hdrop.on ? y.mask1 : nullptr, d.MW, hdrop.scale,
          g->ln1_g, g->ln1_b);
    else
      tfm_ln_bwd_kernel<false><<<grid_for(R, 4, 16 * g_dev.sm_count), 128, 0, c.stream>>>(
          y.s1, w.gb, R, d.D, p.ln1_g, d.eps, w.ds, w.dtp, dtp_lo, d.D, hdrop.on ? y.mask1 : nullptr, d.MW, hdrop.scale,
          nullptr, nullptr);
    LAUNCH_CHECK("tfm ln1_bwd");
    if (g) TRY(gemm_tn(c, w.dtp, d.D, d.D, y.cp, d.Dp, d.Dp, R, g->ao_w, d.D, d.D, g->ao_b, "tfm gemm attn_out wgrad"));
    {  // dO = dt1 Wao
      GemmEpi e = epi_none();
      Sinks sk;
      sk.f32 = w.d_o; sk.ld_f32 = d.D; sk.f32_cols = d.D;
      TRY(gemm_nt(c, w.dtp, R, d.D, pk.wao_t, d.D, d.D, d.D, e, sk, "tfm gemm attn_out dgrad"));
    }
    TRY(tfm_attn_bwd(c, d, w, y, N, T, adrop, lseed));
    if (g) {  // the three projections are separate tensors: one product per column section of dqkv
      float* gw[3] = {g->q_w, g->k_w, g->v_w};
      float* gbias[3] = {g->q_b, g->k_b, g->v_b};
      for (int s = 0; s < 3; ++s)
        TRY(gemm_tn(c, w.dqkv + (long long)s * d.D, d.D, d.P3, y.xp, d.Dp, d.Dp, R, gw[s], d.D, d.D, gbias[s],
                    "tfm gemm qkv wgrad"));
    }
    if (need_dx) {  // dx = dqkv Wqkv + ds1
      GemmEpi e = epi_none();
      e.add_mat = w.ds; e.ld_addmat = d.D;
      Sinks sk;
      sk.f32 = w.ga; sk.ld_f32 = d.D; sk.f32_cols = d.D;
      TRY(gemm_nt(c, w.dqkv, R, d.P3, pk.wqkv_t, d.D, d.P3, d.P3, e, sk, "tfm gemm qkv dgrad"));
    }
    dy = w.ga;
  }
  if (embed_grads) {
    const nrl_tfm_embed_grads* eg = embed_grads;
    if (!(eg->word && eg->pos && eg->type0 && eg->ln_g && eg->ln_b))
      return fail(NRL_ERR_INVALID_ARG, "nrl_tfm_encoder_bwd: embeddings partly frozen (all 5 gradients or none)");
    tfm_embed_bwd_kernel<<<grid_for(R, 8, 8 * g_dev.sm_count), 256, 0, c.stream>>>(
        input_ids, w.pos, R, embed->word, d.V, embed->pos, d.P, embed->type0, d.D, embed->ln_g, d.eps,
        hdrop.on ? w.mask_e : nullptr, d.MW, hdrop.scale, d.pad, d.posmode == 0 ? d.pad : -1, dy, eg->word, eg->pos, eg->type0,
        eg->ln_g, eg->ln_b);
    LAUNCH_CHECK("tfm embed_bwd");
  }
  return NRL_OK;
}

int nrl_tfm_attn_dropout_mask(unsigned char* keep, int layer, int n, int h, int heads, int T, unsigned long long seed,
                              float p, void* stream) {
  if (!keep || layer < 0 || n < 0 || h < 0 || h >= heads || T <= 0 || p < 0.f || p >= 1.f)
    return fail(NRL_ERR_INVALID_ARG, "nrl_tfm_attn_dropout_mask: bad argument");
  const int SK = 32 * ((T + 31) / 32);
  tfm_attn_mask_kernel<<<grid_for((long long)T * T, 256, 64), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      keep, T, SK, (unsigned long long)n * heads + h, tfm_layer_seed(seed, layer), drop_threshold(p));
  LAUNCH_CHECK("tfm attn_mask");
  return NRL_OK;
}
int nrl_tfm_hidden_dropout_mask(unsigned char* keep, long long R, int D, int site, unsigned long long seed, float p,
                                void* stream) {
  if (!keep || R <= 0 || D <= 0 || site < 0 || p < 0.f || p >= 1.f)
    return fail(NRL_ERR_INVALID_ARG, "nrl_tfm_hidden_dropout_mask: bad argument");
  const unsigned long long s = site == 0 ? seed : tfm_layer_seed(seed, (site - 1) / 2);
  const uint32_t st = site == 0 ? 0u : (uint32_t)((site - 1) & 1);
  dropout_mask_kernel<<<grid_for(R * D, 256, 4096), 256, 0, static_cast<cudaStream_t>(stream)>>>(keep, R * D, s, st,
                                                                                               drop_threshold(p));
  LAUNCH_CHECK("tfm hidden_mask");
  return NRL_OK;
}
