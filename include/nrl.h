/* newsreclib_b200 C ABI: the sm_100a implementation of NewsRecLib's two-tower hot path
 * (NewsEncoder -> UserEncoder -> click score, + soft-target CE, backward, Adam).
 *
 * Drop-in boundary.  The reference (andreeaiana/newsreclib @ f29aea8) is pure Python and has
 * no FFI of its own; what a maintainer would bind is the set of torch.nn forward calls on the
 * path.  Every entry point below names the reference interface it replaces (file:line, paths
 * relative to the reference root).  INTEGRATION.md shows the ctypes stub the reference side
 * would add.
 *
 * Conventions
 *   - plain pointers + sizes, no torch types; all pointers are DEVICE pointers unless the
 *     name ends in _host; caller owns every buffer, nothing is allocated or freed here
 *     (one exception: nrl_peer_alloc / nrl_peer_free, because peer-mapped memory must come
 *     from cudaMalloc -- see the gradient-exchange section)
 *   - all launches go to the cudaStream_t passed as `stream` (void* to keep this header C)
 *   - return 0 on success, negative nrl_status on failure; nrl_last_error() gives the text
 *   - float = fp32, ids / segment ids = int64 (what the reference collate emits,
 *     newsreclib/data/components/rec_dataset.py:148-168)
 *   - gradient outputs ACCUMULATE (+=) into the caller's buffers (autograd semantics); zero
 *     them first if that is what you want
 *   - precision: NRL_PREC_BF16X3 (default; three bf16 tensor-core passes on hi/lo split
 *     operands, fp32-equivalent: logits within 1e-4 of the reference) or NRL_PREC_BF16
 *     (single pass, for the bf16 configuration)
 */
#ifndef NRL_H_
#define NRL_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  NRL_OK = 0,
  NRL_ERR_INVALID_ARG = -1,
  NRL_ERR_WORKSPACE_TOO_SMALL = -2,
  NRL_ERR_CUDA = -3,
  NRL_ERR_UNSUPPORTED = -4
} nrl_status;

enum { NRL_PREC_BF16X3 = 0, NRL_PREC_BF16 = 1 };

/* One "MHSA + additive pooling" block: the parameters of nn.MultiheadAttention followed by
 * AdditiveAttention, as held by MHSAAddAtt (encoders/news/text.py:218-219) and by the NRMS
 * UserEncoder (encoders/user/nrms.py:27-30).  state_dict names in comments. */
typedef struct {
  const float* in_proj_weight;  /* multihead_attention.in_proj_weight  [3E, E] */
  const float* in_proj_bias;    /* multihead_attention.in_proj_bias    [3E]    */
  const float* out_proj_weight; /* multihead_attention.out_proj.weight [E, E]  */
  const float* out_proj_bias;   /* multihead_attention.out_proj.bias   [E]     */
  const float* add_weight;      /* additive_attention.linear.weight    [Q, E]  */
  const float* add_bias;        /* additive_attention.linear.bias      [Q]     */
  const float* add_query;       /* additive_attention.query            [Q]     */
} nrl_block_params;

typedef struct {
  float* in_proj_weight;
  float* in_proj_bias;
  float* out_proj_weight;
  float* out_proj_bias;
  float* add_weight;
  float* add_bias;
  float* add_query;
} nrl_block_grads;

typedef struct {
  int embed_dim;  /* E  (configs/model/nrms.yaml:19) */
  int num_heads;  /* h  (:20); E / h must be 16, 20, 32, 48 or 64 */
  int query_dim;  /* Q  (:21) */
} nrl_dims;

const char* nrl_version(void);
const char* nrl_last_error(void);

/* ---- MHSAAddAtt.forward, encoders/news/text.py:222-236 (title ids -> news vectors) -------
 * ids [n_news, L] int64; table [V1, E]; out [n_news, E].  training != 0 applies the two
 * nn.Dropout(p) sites (text.py:225,230) with the counter-based mask keyed by `seed`.
 * `ws` keeps the activations for nrl_news_encoder_bwd. */
size_t nrl_news_encoder_ws_bytes(long long n_news, int L, nrl_dims dims);
int nrl_news_encoder_fwd(const long long* ids, long long n_news, int L, const float* table,
                         long long V1, const nrl_block_params* params, nrl_dims dims,
                         float dropout_p, int training, unsigned long long seed, float* out,
                         void* ws, size_t ws_bytes, int precision, void* stream);
/* Backward of the above: d_out [n_news, E] -> parameter gradients (+=) and the dense
 * embedding gradient d_table [V1, E] (+=, row 0 untouched = padding_idx, text.py:215-217). */
int nrl_news_encoder_bwd(const long long* ids, long long n_news, int L, long long V1,
                         const nrl_block_params* params, nrl_dims dims, float dropout_p,
                         int training, unsigned long long seed, const float* d_out,
                         nrl_block_grads* grads, float* d_table, void* ws, size_t ws_bytes,
                         int precision, void* stream);

/* ---- NRMS UserEncoder.forward, encoders/user/nrms.py:32-41 -------------------------------
 * hist [B, Hmax, E] dense (zero padded) -> user [B, E].  attention_axis 0 = reference
 * behaviour (self-attention across the B impressions at every history position, the
 * batch_first=False quirk), 1 = along the history. */
size_t nrl_user_encoder_ws_bytes(int B, int Hmax, nrl_dims dims);
int nrl_user_encoder_fwd(const float* hist, int B, int Hmax, const nrl_block_params* params,
                         nrl_dims dims, int attention_axis, float* user, void* ws,
                         size_t ws_bytes, int precision, void* stream);
int nrl_user_encoder_bwd(int B, int Hmax, const nrl_block_params* params, nrl_dims dims,
                         int attention_axis, const float* d_user, nrl_block_grads* grads,
                         float* d_hist, void* ws, size_t ws_bytes, int precision, void* stream);

/* ---- AdditiveAttention.forward alone, layers/attention.py:24-42 (NAML user encoder,
 * encoders/user/naml.py:27-31, and the NAML view combiner, encoders/news/news.py:162-163) --
 * x [G, L, D] -> out [G, D]. */
size_t nrl_additive_ws_bytes(long long G, int L, int D, int Q);
int nrl_additive_fwd(const float* x, long long G, int L, int D, int Q, const float* weight,
                     const float* bias, const float* query, float* out, void* ws,
                     size_t ws_bytes, int precision, void* stream);
/* Backward (same `ws` as the forward call): d_out [G, D] -> dx [G, L, D] (overwritten) and the
 * parameter gradients (+=). */
int nrl_additive_bwd(const float* x, long long G, int L, int D, int Q, const float* weight,
                     const float* query, const float* d_out, float* dx, float* g_weight,
                     float* g_bias, float* g_query, void* ws, size_t ws_bytes, int precision,
                     void* stream);

/* ---- NAML: CNNAddAtt.forward, encoders/news/text.py:163-176 ---------------------------------
 * ids [n_news, L] int64 -> out [n_news, F]:  embedding -> dropout -> Conv2d(1, F, (w, E),
 * padding ((w-1)/2, 0)) -> ReLU -> dropout -> AdditiveAttention(F, Q).  The conv is one K = w*E
 * tensor-core GEMM over an im2col of the gathered rows.  w must be odd, E % 4 == 0, F % 4 == 0. */
typedef struct {
  const float* cnn_weight; /* cnn.weight                        [F, 1, w, E] */
  const float* cnn_bias;   /* cnn.bias                          [F]          */
  const float* add_weight; /* additive_attention.linear.weight  [Q, F]       */
  const float* add_bias;   /* additive_attention.linear.bias    [Q]          */
  const float* add_query;  /* additive_attention.query          [Q]          */
} nrl_cnn_params;
typedef struct {
  float* cnn_weight;
  float* cnn_bias;
  float* add_weight;
  float* add_bias;
  float* add_query;
} nrl_cnn_grads;
typedef struct {
  int embed_dim;   /* E  text_embed_dim (configs/model/naml.yaml:20) */
  int num_filters; /* F  (:23) */
  int window;      /* w  (:24) */
  int query_dim;   /* Q  (:25) */
} nrl_cnn_dims;
size_t nrl_cnn_encoder_ws_bytes(long long n_news, int L, nrl_cnn_dims dims);
int nrl_cnn_encoder_fwd(const long long* ids, long long n_news, int L, const float* table,
                        long long V1, const nrl_cnn_params* params, nrl_cnn_dims dims,
                        float dropout_p, int training, unsigned long long seed, float* out,
                        void* ws, size_t ws_bytes, int precision, void* stream);
int nrl_cnn_encoder_bwd(const long long* ids, long long n_news, int L, long long V1,
                        const nrl_cnn_params* params, nrl_cnn_dims dims, float dropout_p,
                        int training, unsigned long long seed, const float* d_out,
                        nrl_cnn_grads* grads, float* d_table, void* ws, size_t ws_bytes,
                        int precision, void* stream);

/* ---- NAML: LinearEncoder.forward (category), encoders/news/category.py:73-82 ----------------
 * ids [n] int64 -> out [n, O] = relu(dropout(table[ids]) W^T + b); table [V1, CE], weight [O, CE].
 * dropout_p applies only when the module was built with use_dropout (NAML: False). */
size_t nrl_linear_encoder_ws_bytes(long long n, int embed_dim, int out_dim);
int nrl_linear_encoder_fwd(const long long* ids, long long n, const float* table, long long V1,
                           int embed_dim, const float* weight, const float* bias, int out_dim,
                           float dropout_p, int training, unsigned long long seed, float* out,
                           void* ws, size_t ws_bytes, int precision, void* stream);
/* `out` is the forward result (ReLU mask).  Gradients accumulate (+=); d_table row 0 untouched. */
int nrl_linear_encoder_bwd(const long long* ids, long long n, long long V1, int embed_dim,
                           const float* weight, int out_dim, float dropout_p, int training,
                           unsigned long long seed, const float* out, const float* d_out,
                           float* g_weight, float* g_bias, float* d_table, void* ws,
                           size_t ws_bytes, int precision, void* stream);

/* ---- PLM text-encoder head: the part of PLM.forward after the transformer, text.py:93-100 ----
 * x [N, T, E] (last hidden states) -> dropout -> nn.MultiheadAttention over dim 0 (the N news of
 * the call: batch_first=False quirk, attention_axis 0; 1 = along the T tokens) -> dropout ->
 * AdditiveAttention over the T tokens -> out [N, E].  E / heads in {16, 20, 32, 48, 64}.
 * Workspace size: nrl_user_encoder_ws_bytes(N, T, dims). */
int nrl_plm_head_fwd(const float* x, int N, int T, const nrl_block_params* params, nrl_dims dims,
                     int attention_axis, float dropout_p, int training, unsigned long long seed,
                     float* out, void* ws, size_t ws_bytes, int precision, void* stream);
int nrl_plm_head_bwd(int N, int T, const nrl_block_params* params, nrl_dims dims,
                     int attention_axis, float dropout_p, int training, unsigned long long seed,
                     const float* d_out, nrl_block_grads* grads, float* d_x, void* ws,
                     size_t ws_bytes, int precision, void* stream);

/* ---- torch_geometric.utils.to_dense_batch (2.3.0), call sites nrms_module.py:233,237 ------
 * seg: sorted int64 segment ids [n]; off: int32 [B+1] CSR offsets (device). */
int nrl_segment_offsets(const long long* seg, long long n, int B, int* off, void* stream);
int nrl_to_dense_fwd(const float* x, const int* off, int B, int M, int E, float* dense,
                     void* stream);
int nrl_to_dense_bwd(const float* d_dense, const int* off, int B, int M, int E, float* dx,
                     void* stream);
/* ---- device-side collate / cache lookup: out[i, :] = table[idx[i], :] ------------------------
 * Replaces the per-row pandas .loc + F.pad + vstack of DatasetCollate._tokenize_df
 * (data/components/rec_dataset.py:170-178,189-285) once the news table is pre-tokenised and
 * resident in HBM, and gathers cached news vectors in the evaluation path.  Rows are row_bytes
 * bytes (multiple of 4); idx [n] int64 must lie in [0, n_table_rows). */
int nrl_gather_rows(const void* table, long long n_table_rows, int row_bytes, const long long* idx,
                    long long n, void* out, void* stream);
/* late fusion, nrms_module.py:243-248 */
int nrl_late_fusion_fwd(const float* hist_vec, const int* off, int B, int E, float* user,
                        void* stream);
int nrl_late_fusion_bwd(const float* d_user, const int* off, int B, int E, float* d_hist_vec,
                        void* stream);

/* ---- DotProduct.forward, layers/click_predictor.py:9-11 via nrms_module.py:251-253 --------
 * user [B, E], cand [N_c, E] ragged with offsets -> scores [B, Cmax] (0.0 in padded slots). */
int nrl_score_fwd(const float* user, const float* cand, const int* cand_off, int B, int Cmax,
                  int E, float* scores, void* stream);
int nrl_score_bwd(const float* d_scores, const float* user, const float* cand,
                  const int* cand_off, int B, int Cmax, int E, float* d_user, float* d_cand,
                  void* stream);

/* ---- CrossEntropyLoss()(scores, y_true) with float targets, nrms_module.py:277,288 --------
 * labels [N_c] ragged.  loss_mean (1 float) is overwritten; loss_rows [B] and y_dense
 * [B, Cmax] are optional (NULL to skip). */
int nrl_ce_soft_fwd(const float* scores, const float* labels, const int* cand_off, int B,
                    int Cmax, float* loss_rows, float* loss_mean, float* y_dense, void* stream);
/* d_scores = g_loss[0] * g_scale * dLoss/dScores (g_loss may be NULL = 1) */
int nrl_ce_soft_bwd(const float* scores, const float* labels, const int* cand_off, int B,
                    int Cmax, const float* g_loss, float g_scale, float* d_scores, void* stream);

/* ---- RetrievalMRR / RetrievalNormalizedDCG(top_k) of the epoch-end hooks, nrms_module.py:182-191,380-396,480 ----
 * (torchmetrics retrieval metrics grouped by `indexes` = the impression of each candidate; third-party, restated.)
 * scores / labels [N] concatenated per impression, off [B + 1] (int64, host-side cumsum of cand_news_size stays on the
 * device), top_k [n_k] HOST array of n_k <= 4 cut-offs.  out [B][1 + n_k] = {reciprocal rank of the first positive (0
 * without one), nDCG@top_k[0], ...} per impression: the epoch value is the mean over B.  Scores are ranked descending
 * with ties in input order (stable), labels are the gains, discount 1 / log2(rank + 1), ideal ordering = the labels
 * sorted descending; ranks [N] (may be NULL) receives each candidate's 1-based rank within its impression. */
int nrl_rank_metrics(const float* scores, const float* labels, const long long* off, int B,
                     const int* top_k, int n_k, float* out, int* ranks, void* stream);

/* ---- SupConLoss()(embeddings=scores, indices_tuple=...) , nrms_module.py:289-316 + components/losses.py:6-40 ----
 * Supervised-contrastive loss over the dense [B, Cmax] score matrix (positives = real candidates with a non-zero
 * label, negatives = real candidates with label 0, padded slots in neither set; per-row
 * -mean_pos(log_softmax_kept(s / T)); mean over the rows whose loss is > 0 -- pytorch-metric-learning's
 * AvgNonZeroReducer).  temperature: the reference always constructs SupConLoss() with its default 0.1
 * (abstract_recommender.py:117-120).  row_loss [B] and stats [3] = {SupCon loss, rows counted, live flag} are saved for
 * the backward.  ce_loss != NULL: `loss` = (1 - dual_loss_coef) * ce_loss[0] + dual_loss_coef * SupCon, the reference's
 * dual loss (nrms_module.py:318-328); else `loss` = SupCon. */
int nrl_supcon_fwd(const float* scores, const float* labels, const int* cand_off, int B, int Cmax,
                   float temperature, const float* ce_loss, float dual_loss_coef, float* row_loss,
                   float* loss, float* stats, void* stream);
/* d_scores (+)= g_loss[0] * g_scale * dSupCon/dScores (g_loss may be NULL = 1; accumulate != 0 adds to d_scores, as
 * the dual loss does after nrl_ce_soft_bwd) */
int nrl_supcon_bwd(const float* scores, const float* labels, const int* cand_off, int B, int Cmax,
                   float temperature, const float* row_loss, const float* stats, const float* g_loss,
                   float g_scale, int accumulate, float* d_scores, void* stream);

/* ---- torch.optim.Adam step (configs/model/nrms.yaml:49-52), dense over n elements --------- */
int nrl_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr,
                  float beta1, float beta2, float eps, long long step, float grad_scale,
                  void* stream);

/* Same step fused with optimizer.zero_grad(): every gradient element is overwritten with 0 once it has been
 * consumed (the gradient buffers of this library accumulate, so they must be clean before the next backward
 * pass; this saves the separate memset of the 87 MB flat gradient buffer). */
int nrl_adam_step_zero_grad(float* p, float* g, float* m, float* v, long long n, float lr,
                            float beta1, float beta2, float eps, long long step, float grad_scale,
                            void* stream);

/* ---- nn.Embedding.forward alone, encoders/news/text.py:215-217 (construction), :224 (call) --------
 * ids [n] int64 -> out_f32 [n, E] (bit-exact copies of the table rows; row 0 is a real row, only its
 * gradient is zeroed) and / or the bf16 hi / lo split planes out_hi / out_lo [n, Ep], Ep = round_up(E + 1, 16),
 * with 1.0 in column E and zeros after (the A-operand layout of the in-projection GEMM).  Any output may be
 * NULL (at least one of out_f32 / out_hi must be given). */
int nrl_embedding_gather(const long long* ids, long long n, const float* table, long long V1, int E,
                         float* out_f32, void* out_hi, void* out_lo, void* stream);

/* ---- device-side input checks -----------------------------------------------------------------------
 * nn.Embedding raises on an id outside the table and to_dense_batch assumes sorted segment ids in [0, B).
 * The kernels here cannot raise: they never touch memory outside the caller's buffers (an offending id reads
 * row 0 / is skipped), record the first violation in a sticky device word, and the host asks for it:
 * synchronises `stream`, writes 0 or the code to *code_host (1 token id out of [0, V1), 2 bad segment ids,
 * 3 a segment longer than Hmax / Cmax, 4 row index out of range), clears the word, and sets nrl_last_error().
 * nrl_nrms_step_host checks it itself (it synchronises anyway) and returns NRL_ERR_INVALID_ARG. */
int nrl_device_status(int* code_host, void* stream);

/* ---- the gradient exchange of data-parallel training fused with the optimizer step ----------
 * Replaces, for world_size > 1, what Lightning DDP + torch.optim.Adam do for the reference after
 * every backward pass (configs/trainer/ddp.yaml, configs/model/nrms.yaml:49-52; gradient mean
 * over the ranks, then Adam on every replica) with ONE kernel per rank over NVLink peer memory:
 * reduce-scatter of the flat gradient buffers by peer loads, Adam on the owned 1/world slice,
 * all-gather of the new parameters by peer stores, bracketed by flag barriers in peer memory.
 *
 * These are the only entry points that allocate: peer-mapped memory has to come from cudaMalloc
 * (IPC handles), so the flat parameter / gradient buffers of a rank live in one nrl_peer_alloc
 * block that the caller lays out and frees.  Every rank allocates, exchanges the 64-byte handles
 * out of band (torch.distributed.all_gather_object in newsreclib_b200/exchange.py) and opens its
 * peers' blocks with its own device current. */
#define NRL_MAX_RANKS 16
#define NRL_FLAG_BYTES 512      /* per-rank flag block: zero once, 8-byte aligned, peer-mapped */
#define NRL_IPC_HANDLE_BYTES 64
typedef struct {
  int world, rank;
  float* params[NRL_MAX_RANKS];              /* flat parameter buffer of every rank ([rank] is local) */
  float* grads[NRL_MAX_RANKS];               /* flat gradient buffer of every rank (cleared by the owner of a slice) */
  unsigned long long* flags[NRL_MAX_RANKS];  /* flag block of every rank */
  unsigned int* bitmaps[NRL_MAX_RANKS];      /* row-bitmap area of every rank: world x ((sparse_rows + 31) / 32) u32,
                                                zero once, peer-mapped; NULL when sparse_rows == 0 */
} nrl_peer_set;
int nrl_peer_alloc(size_t bytes, void** dev_ptr, unsigned char* handle /* [NRL_IPC_HANDLE_BYTES] */);
int nrl_peer_free(void* dev_ptr);
int nrl_peer_open(const unsigned char* handle, void** dev_ptr);
int nrl_peer_close(void* dev_ptr);
/* One exchange + Adam step over n elements (n % 4 == 0, all buffers 16-byte aligned).  m / v are
 * LOCAL and indexed like the parameters; only the owned slice is read and written.  `epoch` must
 * be the same on all ranks and grow by at least 1 per call (the trainer passes its step count);
 * grad_scale = 1 / world gives DDP's mean.  max_ctas <= 0: 4 CTAs per SM (never more than fit the
 * device at once: the grid waits for itself).  timeout_ns == 0: 5 s.
 * sparse_rows > 0: the first sparse_rows * row_elems elements are a row-sparse gradient (the dense
 * [V+1, E] embedding gradient, of which a step touches a fraction): each rank publishes which of its rows are
 * non-zero and all-zero rows never cross the links; the result is bit-identical to the dense exchange.
 * zero_grads != 0: every gradient element, on every rank, is zero when the kernel ends
 * (optimizer.zero_grad() folded in: the owner of a slice clears what it has consumed).
 * A timed-out ready barrier leaves parameters and moments of this rank's slice untouched (one decision
 * per kernel, not per CTA).  All ranks must call it for the same epoch; no other cross-rank wait may sit
 * between. */
int nrl_exchange_adam_step(const nrl_peer_set* peers, float* m, float* v, long long n, float lr,
                           float beta1, float beta2, float eps, long long step,
                           unsigned long long epoch, float grad_scale, int max_ctas,
                           unsigned long long timeout_ns, long long sparse_rows, int row_elems,
                           int zero_grads, void* stream);
/* Synchronises `stream` and returns the flag block's error word (0 = every barrier completed,
 * 1 / 2 = a ready / done wait timed out: a peer never arrived). */
int nrl_exchange_status(const unsigned long long* flags_local, unsigned long long* error_host,
                        void* stream);

/* ---- NRMSModule.forward + model_step loss + backward, nrms_module.py:230-255,277,288 ------
 * One call = one pass of the hot path over one batch.  All inputs are device pointers.
 *   hist_ids [N_h, L], cand_ids [N_c, L], seg_hist [N_h], seg_cand [N_c] (sorted), labels [N_c]
 *   scores   [B, Cmax]   out
 *   loss     [1]         out (NULL for pure inference)
 *   do_backward != 0: parameter gradients (+=) into news_grads / user_grads / d_table.
 * Hmax / Cmax are the dense widths max_b h_b / max_b c_b (known at collate time). */
size_t nrl_nrms_ws_bytes(long long n_hist, long long n_cand, int L, int B, int Hmax, int Cmax,
                         nrl_dims dims);
int nrl_nrms_step(const long long* hist_ids, const long long* cand_ids, const long long* seg_hist,
                  const long long* seg_cand, const float* labels, long long n_hist,
                  long long n_cand, int L, int B, int Hmax, int Cmax, const float* table,
                  long long V1, const nrl_block_params* news_params,
                  const nrl_block_params* user_params, nrl_dims dims, int late_fusion,
                  float dropout_p, int training, unsigned long long seed, float* scores,
                  float* loss, int do_backward, nrl_block_grads* news_grads,
                  nrl_block_grads* user_grads, float* d_table, void* ws, size_t ws_bytes,
                  int precision, void* stream);
/* The backward half alone, for callers whose framework asks for gradients later than the forward pass (an autograd
 * backward()): continues a forward-only nrl_nrms_step (do_backward = 0) that ran with the SAME sizes, parameters,
 * dropout configuration and workspace -- `ws` still holds the news / user vectors, offsets, ids, packed weights,
 * keep-bit words and saved activations of that pass and must not have been touched in between.  g_loss [1] (device,
 * NULL = 1.0) is d(objective) / d(loss), folded into d loss / d scores.  Gradients are accumulated (+=) exactly as by
 * nrl_nrms_step(do_backward = 1); the two calls together launch the same kernels as that one call (the scorer twice). */
int nrl_nrms_step_bwd(const float* labels, const float* g_loss, long long n_hist, long long n_cand, int L,
                      int B, int Hmax, int Cmax, long long V1, const nrl_block_params* news_params,
                      const nrl_block_params* user_params, nrl_dims dims, int late_fusion,
                      float dropout_p, int training, unsigned long long seed,
                      nrl_block_grads* news_grads, nrl_block_grads* user_grads, float* d_table, void* ws,
                      size_t ws_bytes, int precision, void* stream);
/* Same pass with HOST input buffers (pinned or pageable): copies ids / segments / labels to the
 * device staging area inside `ws`, runs nrl_nrms_step, copies scores and loss back to
 * scores_host / loss_host and synchronises the stream.  This is the end-to-end call. */
int nrl_nrms_step_host(const long long* hist_ids_host, const long long* cand_ids_host,
                       const long long* seg_hist_host, const long long* seg_cand_host,
                       const float* labels_host, long long n_hist, long long n_cand, int L, int B,
                       int Hmax, int Cmax, const float* table, long long V1,
                       const nrl_block_params* news_params, const nrl_block_params* user_params,
                       nrl_dims dims, int late_fusion, float dropout_p, int training,
                       unsigned long long seed, float* scores_host, float* loss_host,
                       int do_backward, nrl_block_grads* news_grads, nrl_block_grads* user_grads,
                       float* d_table, void* ws, size_t ws_bytes, int precision, void* stream);

/* The same end-to-end call split in two, so that the host can queue more work behind the step before it waits for the
 * step's results (a training loop queues the gradient exchange + optimizer step there: the wait for the loss and the
 * host->device copies of the NEXT batch then overlap it):
 *   begin: copies the host inputs on `copy_stream` (NULL = on `stream`; a different stream lets the copies overlap what
 *          is still running on `stream`), makes `stream` wait for them, enqueues the step and the copies of scores /
 *          loss (and, when status_host != NULL, of the device-side input-check word) back to the host; does not wait.
 *          Host buffers should be pinned (pageable memory makes the copies synchronous).  *ticket identifies the step.
 *   end:   waits until the results of that step are in scores_host / loss_host, releases the ticket, and returns
 *          NRL_ERR_INVALID_ARG when *status_host reports a device-side input violation (nrl_device_status semantics).
 * A workspace may carry ONE step at a time: call end() before the next begin() on the same `ws`. */
int nrl_nrms_step_host_begin(const long long* hist_ids_host, const long long* cand_ids_host,
                             const long long* seg_hist_host, const long long* seg_cand_host,
                             const float* labels_host, long long n_hist, long long n_cand, int L, int B,
                             int Hmax, int Cmax, const float* table, long long V1,
                             const nrl_block_params* news_params, const nrl_block_params* user_params,
                             nrl_dims dims, int late_fusion, float dropout_p, int training,
                             unsigned long long seed, float* scores_host, float* loss_host,
                             int do_backward, nrl_block_grads* news_grads, nrl_block_grads* user_grads,
                             float* d_table, void* ws, size_t ws_bytes, int precision, void* copy_stream,
                             unsigned int* status_host, void* stream, void** ticket);
int nrl_nrms_step_host_end(void* ticket);

/* ---- PLM news encoder internals ---------------------------------------
 * The transformer inside PLM.forward (encoders/news/text.py:67-73 constructor / freezing, :92
 * `self.plm_model(**text)[0]`): a HF RobertaModel / BertModel post-LN encoder -- embeddings
 * (word + position + token type -> LayerNorm -> dropout) and `num_layers` layers of
 *   h1 = LN(x + drop(MHSA(x) W_ao + b)),   h2 = LN(h1 + drop(gelu(h1 W_i + b) W_o + b)),
 * key-padding mask from attention_mask, exact (erf) GELU, head dim 64, any T <= max_pos.
 * The third-party algorithm restated here is transformers' modeling_roberta.py (pinned by the
 * reference at transformers 4.x; identical maths in the 5.5 of this image, which is what the
 * golden vectors under tests/golden/tfm_*.npz were minted with).  state_dict names in comments
 * are relative to `plm_model.` (embeddings.*) and `plm_model.encoder.layer.<i>.` (layers).
 * A layer whose nrl_tfm_layer_grads is all NULL is FROZEN (text.py:70-73): only the data gradient
 * passes through it. */
typedef struct {
  const float* word;   /* embeddings.word_embeddings.weight        [vocab, D]   (row pad_idx never updated) */
  const float* pos;    /* embeddings.position_embeddings.weight    [max_pos, D] (row pad_idx never updated) */
  const float* type0;  /* embeddings.token_type_embeddings.weight  row 0 [D]    (token_type_ids are all 0) */
  const float* ln_g;   /* embeddings.LayerNorm.weight [D] */
  const float* ln_b;   /* embeddings.LayerNorm.bias   [D] */
} nrl_tfm_embed_params;
typedef struct {
  float *word, *pos, *type0, *ln_g, *ln_b;
} nrl_tfm_embed_grads;
typedef struct {
  const float *q_w, *q_b;     /* attention.self.query.{weight,bias}   [D, D], [D] */
  const float *k_w, *k_b;     /* attention.self.key.*                               */
  const float *v_w, *v_b;     /* attention.self.value.*                             */
  const float *ao_w, *ao_b;   /* attention.output.dense.*             [D, D], [D] */
  const float *ln1_g, *ln1_b; /* attention.output.LayerNorm.*         [D]         */
  const float *i_w, *i_b;     /* intermediate.dense.*                 [I, D], [I] */
  const float *o_w, *o_b;     /* output.dense.*                       [D, I], [D] */
  const float *ln2_g, *ln2_b; /* output.LayerNorm.*                   [D]         */
} nrl_tfm_layer_params;
typedef struct {
  float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *ao_w, *ao_b, *ln1_g, *ln1_b, *i_w, *i_b, *o_w, *o_b,
      *ln2_g, *ln2_b;
} nrl_tfm_layer_grads;
typedef struct {
  int hidden;        /* D: config.hidden_size (multiple of 64, <= 1024) */
  int heads;         /* config.num_attention_heads; D / heads must be 64 */
  int intermediate;  /* I: config.intermediate_size (multiple of 16) */
  int num_layers;    /* layers handled by the call */
  int vocab;         /* config.vocab_size */
  int max_pos;       /* config.max_position_embeddings */
  int pad_idx;       /* config.pad_token_id: padding row of the word table (never updated) */
  float ln_eps;      /* config.layer_norm_eps */
  float hidden_dropout; /* config.hidden_dropout_prob (embeddings, attention output, layer output) */
  float attn_dropout;   /* config.attention_probs_dropout_prob */
  int position_mode;    /* 0: RoBERTa (position ids = cumsum(ids != pad_idx) * (ids != pad_idx) + pad_idx, the position
                           table's row pad_idx is a padding row); 1: BERT (position ids 0..T-1, no padding row in the
                           position table).  pad_idx is the padding row of the WORD table in both modes. */
} nrl_tfm_dims;

/* packed bf16 hi/lo GEMM operands of the layers' weights (forward + transposed copies).  They
 * change only when the weights do: the caller re-packs layers [first, first + count) after an
 * optimizer step (frozen layers: once).  `layers` points at the parameters of layer `first`. */
size_t nrl_tfm_wpack_bytes(nrl_tfm_dims dims);
int nrl_tfm_pack_weights(const nrl_tfm_layer_params* layers, int first, int count, nrl_tfm_dims dims,
                         void* wpack, size_t wpack_bytes, int precision, void* stream);
/* keep_activations != 0: activations of all layers (kept for the backward pass) + backward scratch;
 * keep_activations == 0 (inference, no backward call will follow): one layer's worth, shared by all layers */
size_t nrl_tfm_ws_bytes(long long N, int T, nrl_tfm_dims dims, int keep_activations);
/* input_ids / attention_mask: int64 [N, T] (the tokenizer output the reference collate emits,
 * rec_dataset.py:181-183; attention_mask may be NULL = all ones).  out: last hidden state
 * [N, T, D] fp32 (`self.plm_model(**text)[0]`).  training != 0 applies the three dropouts. */
int nrl_tfm_encoder_fwd(const long long* input_ids, const long long* attention_mask, int N, int T,
                        const nrl_tfm_embed_params* embed, const nrl_tfm_layer_params* layers,
                        nrl_tfm_dims dims, int training, unsigned long long seed, const void* wpack,
                        float* out, int keep_activations, void* ws, size_t ws_bytes, int precision,
                        void* stream);
/* d_out [N, T, D] -> parameter gradients (+=).  embed_grads NULL: embeddings frozen.
 * layer_grads[i] all-NULL: layer i frozen.  Same ids / mask / seed / ws as the forward call, which
 * must have been made with keep_activations != 0. */
int nrl_tfm_encoder_bwd(const long long* input_ids, const long long* attention_mask, int N, int T,
                        const nrl_tfm_embed_params* embed, const nrl_tfm_layer_params* layers,
                        nrl_tfm_dims dims, int training, unsigned long long seed, const void* wpack,
                        const float* d_out, const nrl_tfm_embed_grads* embed_grads,
                        const nrl_tfm_layer_grads* layer_grads, void* ws, size_t ws_bytes,
                        int precision, void* stream);
/* keep-flags of the attention-probability dropout of (layer, title n, head h): keep [T][T] bytes
 * (query-major), for tests that replay the kernel's own mask in the oracle */
int nrl_tfm_attn_dropout_mask(unsigned char* keep, int layer, int n, int h, int heads, int T,
                              unsigned long long seed, float p, void* stream);
/* keep-flags of the hidden dropouts: site 0 = embeddings, 1 + 2 l = attention output of layer l,
 * 2 + 2 l = layer output of layer l; keep [R = N T][D] bytes */
int nrl_tfm_hidden_dropout_mask(unsigned char* keep, long long R, int D, int site,
                                unsigned long long seed, float p, void* stream);

/* ---- test / measurement helpers ----------------------------------------------------------- */
/* keep[i] = 1 iff element i of dropout site `site` (0 = after the embedding, 1 = after the
 * MHSA) is kept for (seed, p): lets a test feed the very same mask to the CPU oracle. */
int nrl_dropout_mask(unsigned char* keep, long long n, unsigned long long seed, int site, float p,
                     void* stream);
/* Raw tensor-core GEMM for unit tests: D[M,N] (fp32, ld = N) = A * B^T over K.
 * mn_major 0: A [M,K], B [N,K] fp32 row-major.  mn_major 1: A [K,M], B [K,N] (D += via atomics,
 * zero D first).  Operands are split to bf16 planes in `ws`. */
size_t nrl_gemm_test_ws_bytes(int M, int N, int K);
int nrl_gemm_test(const float* A, const float* B, float* D, int M, int N, int K, int mn_major,
                  int precision, void* ws, size_t ws_bytes, void* stream);
/* Same NT GEMM through the split-plane sink (the layout the encoder GEMMs hand to the next GEMM):
 * out [M, Np] fp32 = hi + lo of the stored bf16 planes, Np = round_up(N + 1, 16); column N reads
 * 1.0 (the bias column) and the remaining pad columns 0. */
int nrl_gemm_test_planes(const float* A, const float* B, float* out, int M, int N, int K,
                         int precision, void* ws, size_t ws_bytes, void* stream);
/* number of kernels launched by this library since load (bench.py reports it) */
long long nrl_launch_count(void);
/* Per-launch device timing for bench.py's roofline leg: after nrl_profile_start(stream) one
 * CUDA event is recorded on `stream` after every launch of this library; nrl_profile_stop
 * synchronises, fills names[i*name_stride..] / ms[i] (duration of launch i = end_i - end_{i-1}
 * on the in-order stream) and returns the number of records (<= max_records). */
int nrl_profile_start(void* stream);
int nrl_profile_stop(char* names, int name_stride, float* ms, int max_records);

#ifdef __cplusplus
}
#endif
#endif /* NRL_H_ */
