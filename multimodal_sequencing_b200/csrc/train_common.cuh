// Shared pieces of the fine-tuning path (train.cu: inner encoder; train_heads.cu: BERSON heads + loss).
#pragma once
#include <array>
#include <math.h>

#include "dropout.cuh"
#include "model.cuh"

namespace msq {

struct ParamSlot { std::string name; int64_t off = 0, numel = 0; float* master = nullptr; bool decay = true; };
struct BertTape { void *x, *qkv, *ctx, *x1, *u, *hb; float *s1, *s2; float* lse; bool have_lse; };   // lse: attention row log-sum-exp [R*heads, L] saved by the forward kernel   // hb = act(u): operand of the down-projection wgrad
struct VitTape { float *x, *x1; void *y1, *qkv, *ctx, *y2, *u, *hb; float* lse; bool have_lse; };

// Split-K of the weight-gradient GEMMs: dW is only 9-72 output tiles (a fraction of the 148 SMs) against a contraction
// of tens of thousands of rows, so the contraction is cut into S slices that run CONCURRENTLY as S launches of the same
// GEMM kernel on S auxiliary streams (fork / join with events), each into its own partial buffer; an ordered pass adds
// the partials into the gradient (deterministic).
constexpr int SPLITK_MAX = 8;
struct SplitK {
  cudaStream_t aux[SPLITK_MAX] = {};
  cudaEvent_t fork = nullptr, join[SPLITK_MAX] = {};
  bool ready = false;
  int init() {
    if (ready) return MSQ_OK;
    for (int i = 0; i < SPLITK_MAX; ++i) {
      MSQ_CUDA(cudaStreamCreateWithFlags(&aux[i], cudaStreamNonBlocking));
      MSQ_CUDA(cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming));
    }
    MSQ_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
    ready = true;
    return MSQ_OK;
  }
};

// ---- ModifiedResNet tower (train_rn.cu)
struct RnConvT {     // one convolution + its BatchNorm
  std::string wname, bn;
  int cin = 0, cout = 0, k = 1, K = 0, Kp = 0;
  const float *w = nullptr, *gamma = nullptr, *beta = nullptr, *rmean = nullptr, *rvar = nullptr;   // fp32 masters
  void *wg = nullptr, *wgT = nullptr;   // GEMM layouts in the operand type: [Cout, Kp] and [Kp, Cout]
  float* Y = nullptr;                   // tape: pre-BatchNorm output [M, Cout], fp32 in every mode
  float *mean = nullptr, *rstd = nullptr, *S1 = nullptr, *S2 = nullptr;
  int64_t M = 0;
  int HW = 0;
  float Wtot = 0.f;                     // weighted element count per channel of the materialised batch
};
struct RnBlockTape { const void* x; void *o1, *o2, *o2p, *xs, *out; int H, Ho; };
struct RnTrain {
  std::vector<RnConvT> convs;           // stem 0..2, then per block conv1, conv2, conv3, [downsample]
  std::vector<RnBlockTape> blk;
  std::vector<void*> owned;
  Arena tape, ws;
  void *qkvT = nullptr, *cprojT = nullptr;
  float *wimg = nullptr, *partial = nullptr, *feat = nullptr, *o = nullptr;
  void *a[3] = {nullptr, nullptr, nullptr}, *x0 = nullptr, *tok = nullptr, *qkv = nullptr, *ctx = nullptr;
  int64_t n_img = 0, R = 0;
  bool bn_eval = false;                 // BatchNorm from the running statistics (model.eval() semantics inside a training step)
  float bn_momentum = 0.1f;             // running-statistics update of a training forward (nn.BatchNorm2d default); 0 = frozen
};
void rn_train_free(RnTrain* r);

struct ParaTape { float *xin, *y, *qkv, *ctx, *out, *pn, *u, *xo, *hf; };   // hf: dropped FFN activations (dropout only)
struct HeadTape {
  int64_t B = 0;
  int N = 0;
  void *topt = nullptr, *ttb = nullptr;                       // operand type [R*Lt, H]
  float *mix, *rel6, *sents, *r0, *para, *h0, *keyin, *key, *sents_ext, *xg, *t4, *act, *c_all, *hs, *hprev, *query, *nll, *d_lang;
  // image pairwise head (args.multimodal_loss): projected image vector, its gradient, d(logits), d(first visual token), row losses
  float *ml_u = nullptr, *ml_du = nullptr, *ml_dz = nullptr, *ml_dv = nullptr, *ml_loss = nullptr;
  std::vector<ParaTape> pl;
};

struct TrainState {
  std::vector<ParamSlot> slots;
  std::unordered_map<std::string, size_t> index;
  int64_t total = 0;
  float *adam_m = nullptr, *adam_v = nullptr, *opt_scratch = nullptr;
  AdamChunk* adam_chunks = nullptr;   // device table for the single-launch update
  int n_chunks = 0;
  int64_t step = 0;
  // W^T copies in the GEMM operand type: [qkv, out, up, down] per BERT layer, [qkv, out, fc, proj] per ViT layer
  std::vector<std::array<void*, 4>> bertT, vitT;
  void* visnT = nullptr;
  std::vector<void*> owned;
  Arena tape;
  // ---- record of the last training forward
  bool have_fwd = false, mm = false;
  int64_t R = 0, n_img = 0;
  int Lt = 0, Lv = 0, Lj = 0;
  const float* images = nullptr;   // caller-owned; must stay valid until msq_inner_backward returns
  int64_t *ids = nullptr, *tt = nullptr;
  int32_t* img_index = nullptr;
  float *mask_add = nullptr, *patch = nullptr, *vx_last = nullptr, *visn_pre = nullptr;
  void *x_last = nullptr, *y_post = nullptr;
  std::vector<BertTape> bt;
  std::vector<VitTape> vt;
  float* x_out = nullptr;          // final joint stream [R*Lj, H] fp32 (what the heads read)
  // ---- BERSON heads (train_heads.cu)
  bool heads = false;                            // head parameters are part of the table
  std::vector<std::array<float*, 4>> paraT;      // fp32 W^T of [qkv, fin, w1, w2] per paragraph layer
  void* sentT = nullptr;                         // sentence_tran W^T (operand type)
  float *keyT = nullptr, *wqT = nullptr, *whhT = nullptr, *wihT = nullptr, *pw4T = nullptr;
  Arena htape;
  HeadTape ht;
  SplitK splitk;
  RnTrain* rn = nullptr;                 // ModifiedResNet tower state (cfg.rn_width != 0)
  // ---- optional time-contrastive objective of the NEXT msq_train_step (msq_train_set_triplets; consumed by that step)
  int32_t* trip = nullptr;               // device [trip_B, 3]: anchor / positive / negative sentence index per manual
  float* trip_loss = nullptr;            // device [trip_cap]
  int64_t trip_B = 0, trip_cap = 0;
  float trip_weight = 0.f;
  // ---- optional image pairwise objective (msq_train_set_multimodal_loss; modeling_bert.py:1218-1225, 1359-1364)
  bool mm_loss = false;                  // every msq_train_step adds lam * NLL(pairwise_relationship(img_projection(visn[:, 0])))
  bool ml_live = false;                  // ht.ml_dv of the step in flight is to be added to the joint-stream gradient
  // ---- dropout (msq_train_set_dropout): probabilities + seed; `step` is the counter of the forward whose masks are live
  DropCfg drop;
  uint32_t drop_next_step = 0;
  // ---- gradient-ready events of the last backward pass, in the order the regions of the flat buffer become final
  // (heads, BERT layers top -> bottom, embeddings, visn_fc, ViT blocks top -> bottom, ViT stem): a data-parallel caller
  // all-reduces region i on a side stream as soon as event i has fired, overlapping the rest of the backward pass
  std::vector<cudaEvent_t> ready_ev;
  std::vector<std::pair<int64_t, int64_t>> ready_rng;   // [begin, end) element offsets
  size_t ready_used = 0;
};

inline void train_state_free_impl(TrainState* ts) {
  if (!ts) return;
  for (void* p : ts->owned) cudaFree(p);
  rn_train_free(ts->rn);
  for (cudaEvent_t e : ts->ready_ev) cudaEventDestroy(e);
  if (ts->trip) cudaFree(ts->trip);
  if (ts->trip_loss) cudaFree(ts->trip_loss);
  if (ts->adam_m) cudaFree(ts->adam_m);
  if (ts->adam_v) cudaFree(ts->adam_v);
  if (ts->opt_scratch) cudaFree(ts->opt_scratch);
  if (ts->adam_chunks) cudaFree(ts->adam_chunks);
  if (ts->tape.base) cudaFree(ts->tape.base);
  if (ts->htape.base) cudaFree(ts->htape.base);
  if (ts->splitk.ready) {
    for (int i = 0; i < SPLITK_MAX; ++i) { cudaStreamDestroy(ts->splitk.aux[i]); cudaEventDestroy(ts->splitk.join[i]); }
    cudaEventDestroy(ts->splitk.fork);
  }
  delete ts;
}

static inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
constexpr int64_t TRAIN_IMG_CHUNK = 1024;

template <typename T> inline const T* wptr(const Lin& l);
template <> inline const float* wptr<float>(const Lin& l) { return l.w32; }
template <> inline const bf16* wptr<bf16>(const Lin& l) { return l.w16; }

// C[M,N] = act(A[M,K] W[N,K]^T + bias) + resid on the model's GEMM path (tcgen05 when available and aligned)
template <typename T>
static bool gemm_nt_on_tc(const msq_model* m, int lda, int ldw, int ldc, int N, int K) {
  return sizeof(T) == 2 && model_use_tc(m) && K % 64 == 0 && N % 8 == 0 && ldc % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0;
}
// drop (tensor-core path only, see gemm_nt_on_tc): C = dropout(act(A W^T + bias)) + resid
template <typename T, typename TO>
static int gemm_nt(const msq_model* m, const T* A, int lda, const T* W, int ldw, const float* bias, const float* resid, int ldr, TO* C,
                   int ldc, int64_t M, int N, int K, int act, cudaStream_t st, const Drop& drop = Drop()) {
  GemmArgs g;
  g.A = A; g.W = W; g.bias = bias; g.resid = resid; g.C = C; g.C2 = nullptr;
  g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldw = ldw; g.ldc = ldc; g.ldr = ldr; g.act = act;
  if constexpr (sizeof(T) == 4) {
    static_assert(sizeof(TO) == 4, "fp32 mode has fp32 outputs");
    MSQ_REQUIRE(drop.thresh == 0, "gemm_nt: fused dropout is a tensor-core epilogue");
    return gemm_simt<float, float>(g, st);
  } else {
    if (gemm_nt_on_tc<T>(m, lda, ldw, ldc, N, K)) { g.drop = drop; return gemm_tc<TO>(g, st); }
    MSQ_REQUIRE(drop.thresh == 0, "gemm_nt: fused dropout is a tensor-core epilogue");
    return gemm_simt<bf16, TO>(g, st);
  }
}

// U = A W^T + bias (kept for the backward pass) and Hb = act(U) (the next GEMM's operand): ONE tensor-core GEMM with two
// outputs (GemmArgs EPI_DUALACT) where the tcgen05 path applies, else the GEMM followed by the activation pass.
// MSQ_DUALACT=0 forces the two-kernel form.
template <typename T>
static int gemm_nt_dualact(const msq_model* m, const T* A, int lda, const T* W, int ldw, const float* bias, T* U, T* Hb, int ld, int64_t M,
                           int N, int K, int act, cudaStream_t st) {
  if constexpr (sizeof(T) == 2) {
    static int on = -1;
    if (on < 0) { const char* e = getenv("MSQ_DUALACT"); on = (e && e[0] == '0') ? 0 : 1; }
    if (on && ld == N && gemm_nt_on_tc<T>(m, lda, ldw, ld, N, K) && (act == ACT_GELU_ERF || act == ACT_QUICK_GELU)) {
      GemmArgs g;
      g.A = A; g.W = W; g.bias = bias; g.resid = nullptr; g.C = U; g.C2 = nullptr; g.C2bf = Hb;
      g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldw = ldw; g.ldc = ld; g.ldr = 0; g.act = act; g.mode = EPI_DUALACT;
      return gemm_tc<bf16>(g, st);
    }
  }
  MSQ_TRY((gemm_nt<T, T>(m, A, lda, W, ldw, bias, nullptr, 0, U, ld, M, N, K, ACT_NONE, st)));
  return act_fwd<T>(U, M * (int64_t)N, act, Hb, st);
}

int splitk_accumulate(const float* part, int S, int64_t n, float* dW, cudaStream_t st);   // train_kernels.cu

struct BwdBufs {
  float *gA, *gB, *ln_scr, *at_scr, *dpatch;
  void *gT, *gH, *gC, *gQ, *GT, *XT, *apatch;
  size_t scr_floats = 0;    // capacity of ln_scr (also the column-sum scratch of the TN weight-gradient path)
  SplitK* sk = nullptr;     // non-null: split the wgrad contraction (bf16 tensor-core path)
  float* part = nullptr;    // [SPLITK_MAX, max Nout*Kin] partial sums
};
static inline bool wgrad_tn_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MSQ_WGRAD_TN"); on = (e && e[0] == '0') ? 0 : 1; }   // MSQ_WGRAD_TN=0: transposed copies instead
  return on != 0;
}
static inline int wgrad_splits(int Nout, int Kin, int64_t M) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("MSQ_WGRAD_SPLITK"); off = (e && e[0] == '0') ? 1 : 0; }   // MSQ_WGRAD_SPLITK=0: one launch per weight
  if (off) return 1;
  const int tiles = ceil_div(Nout, 128) * ceil_div(Kin, 256);
  int S = 148 / (tiles > 0 ? tiles : 1);
  if (S > SPLITK_MAX) S = SPLITK_MAX;
  while (S > 1 && M / S < 2048) --S;     // keep every slice a long contraction
  return S < 1 ? 1 : S;
}
// rows of the transposed operands: padded so that every split-K slice is a multiple of 64
static inline int64_t wgrad_rows(int64_t M) { return round_up(M, 64 * SPLITK_MAX); }

// dW[Nout,Kin] += G^T X (act_x applied to X on the fly), db[Nout] += column sums of G
template <typename T>
static int wgrad(const msq_model* m, const T* G, int ldg, int Nout, const T* X, int ldx, int Kin, int act_x, int64_t M, float* dW, float* db,
                 BwdBufs& b, cudaStream_t st) {
  int S = 1;
  if constexpr (sizeof(T) == 2) {
    if (b.sk && b.part && model_use_tc(m)) S = wgrad_splits(Nout, Kin, M);
    // TN path: the GEMM reads dY [M, Nout] and X [M, Kin] as MN-major operands straight from HBM (GemmArgs::tn) -- no
    // transposed copies; act(X) is materialised once into the XT scratch when the operand is a recomputed activation.
    if (wgrad_tn_enabled() && b.ln_scr && model_use_tc(m) && Nout % 8 == 0 && Kin % 8 == 0 && ldg % 8 == 0 && ldx % 8 == 0 &&
        colsum_scratch_floats(Nout) <= b.scr_floats) {
      const T* Xo = X;
      int ldxo = ldx;
      if (act_x != ACT_NONE) {
        MSQ_REQUIRE(ldx == Kin, "wgrad: strided activation operand");
        MSQ_TRY(act_fwd<T>(X, M * (int64_t)Kin, act_x, (T*)b.XT, st));
        Xo = (const T*)b.XT; ldxo = Kin;
      }
      GemmArgs g;
      g.tn = 1; g.bias = nullptr; g.C2 = nullptr; g.N = Kin; g.M = Nout; g.lda = ldg; g.ldw = ldxo; g.ldc = Kin; g.ldr = Kin; g.act = ACT_NONE;
      if (S > 1) {
        const int64_t Mc = round_up((M + S - 1) / S, 64);
        MSQ_TRY(b.sk->init());
        MSQ_CUDA(cudaEventRecord(b.sk->fork, st));
        int used = 0;
        for (int s = 0; s < S && s * Mc < M; ++s, ++used) {
          MSQ_CUDA(cudaStreamWaitEvent(b.sk->aux[s], b.sk->fork, 0));
          g.A = G + s * Mc * ldg; g.W = Xo + s * Mc * ldxo; g.K = (int)min(Mc, M - s * Mc); g.resid = nullptr;
          g.C = b.part + (size_t)s * Nout * Kin;
          MSQ_TRY(gemm_tc<float>(g, b.sk->aux[s]));
          MSQ_CUDA(cudaEventRecord(b.sk->join[s], b.sk->aux[s]));
        }
        // the bias gradient (an HBM-bound column sum of dY) runs on the main stream UNDER the tensor-bound slices
        if (db) MSQ_TRY(colsum_accum_bf16((const bf16*)G, M, Nout, ldg, db, b.ln_scr, st));
        for (int s = 0; s < used; ++s) MSQ_CUDA(cudaStreamWaitEvent(st, b.sk->join[s], 0));
        MSQ_TRY(splitk_accumulate(b.part, used, (int64_t)Nout * Kin, dW, st));
      } else {
        g.A = G; g.W = Xo; g.K = (int)M; g.resid = dW; g.C = dW;
        MSQ_TRY(gemm_tc<float>(g, st));
        if (db) MSQ_TRY(colsum_accum_bf16((const bf16*)G, M, Nout, ldg, db, b.ln_scr, st));
      }
      return MSQ_OK;
    }
  }
  if (S > 1) {
    const int64_t Mc = round_up((M + S - 1) / S, 64), Mp = Mc * S;
    MSQ_TRY((transpose_pad<T, T>(G, M, Nout, ldg, Mp, (T*)b.GT, ACT_NONE, st)));
    MSQ_TRY((transpose_pad<T, T>(X, M, Kin, ldx, Mp, (T*)b.XT, act_x, st)));
    MSQ_TRY(b.sk->init());
    MSQ_CUDA(cudaEventRecord(b.sk->fork, st));
    for (int s = 0; s < S; ++s) {
      MSQ_CUDA(cudaStreamWaitEvent(b.sk->aux[s], b.sk->fork, 0));
      MSQ_TRY((gemm_nt<T, float>(m, (const T*)b.GT + s * Mc, (int)Mp, (const T*)b.XT + s * Mc, (int)Mp, nullptr, nullptr, 0,
                                 b.part + (size_t)s * Nout * Kin, Kin, Nout, Kin, (int)Mc, ACT_NONE, b.sk->aux[s])));
      MSQ_CUDA(cudaEventRecord(b.sk->join[s], b.sk->aux[s]));
      MSQ_CUDA(cudaStreamWaitEvent(st, b.sk->join[s], 0));
    }
    MSQ_TRY(splitk_accumulate(b.part, S, (int64_t)Nout * Kin, dW, st));
    if (db) MSQ_TRY(rowsum_accum<T>((const T*)b.GT, Nout, Mp, Mp, db, st));
    return MSQ_OK;
  }
  const int64_t Mp = round_up(M, 64);
  MSQ_TRY((transpose_pad<T, T>(G, M, Nout, ldg, Mp, (T*)b.GT, ACT_NONE, st)));
  MSQ_TRY((transpose_pad<T, T>(X, M, Kin, ldx, Mp, (T*)b.XT, act_x, st)));
  MSQ_TRY((gemm_nt<T, float>(m, (const T*)b.GT, (int)Mp, (const T*)b.XT, (int)Mp, nullptr, dW, Kin, dW, Kin, Nout, Kin, (int)Mp, ACT_NONE, st)));
  if (db) MSQ_TRY(rowsum_accum<T>((const T*)b.GT, Nout, Mp, Mp, db, st));
  return MSQ_OK;
}
// dX[M,Kin] = G[M,Nout] W  (+ resid), W^T given as [Kin, Nout]
template <typename T, typename TO>
static int dgrad(const msq_model* m, const T* G, int Nout, const void* WT, int Kin, const float* resid, TO* dX, int64_t M, cudaStream_t st) {
  return gemm_nt<T, TO>(m, G, Nout, (const T*)WT, Nout, nullptr, resid, Kin, dX, Kin, M, Kin, Nout, ACT_NONE, st);
}

// attention backward: tensor cores (mma.sync) on the bf16 path, fp32 CUDA cores in the parity mode
template <typename T>
static int attn_bwd(const T* qkv, const T* ctx, const T* dctx, int64_t R, int L, int heads, const float* mask, int mask_len, T* dqkv, float* scratch,
                    cudaStream_t st, const Drop& drop = Drop(), const float* lse_fwd = nullptr) {
  if constexpr (sizeof(T) == 2) {
    if (attention_bwd_mma_supported(L)) {
      static int use_lse = -1;   // MSQ_ATTN_LSE=0: ignore the forward's saved log-sum-exp (the dQ kernel recomputes it)
      if (use_lse < 0) { const char* e = getenv("MSQ_ATTN_LSE"); use_lse = (e && e[0] == '0') ? 0 : 1; }
      return attention_bwd_mma(qkv, ctx, dctx, R, L, heads, 0.125f, mask, mask_len, mask_len, dqkv, scratch, st, drop, use_lse ? lse_fwd : nullptr);
    }
  }
  return attention_bwd<T>(qkv, dctx, R, L, heads, 0.125f, mask, mask_len, mask_len, dqkv, scratch, st, drop);
}
// mark the slots [first, last] (by name) final (train.cu)
int mark_ready(TrainState* ts, const std::string& first, const std::string& last, cudaStream_t st);
// ---- train_rn.cu
std::vector<std::string> rn_param_names(const msq_model* m);
template <typename T> int rn_train_build(msq_model* m, cudaStream_t st);
template <typename T> int rn_train_refresh(msq_model* m, cudaStream_t st);
template <typename T> int rn_forward_train(msq_model* m, const float* images, int64_t n_img, const int32_t* img_index, int64_t R, cudaStream_t st);
template <typename T> int rn_backward_train(msq_model* m, const float* dyp, float* grads, cudaStream_t st);

// ---- train_heads.cu
int heads_train_setup(msq_model* m, bool alloc, cudaStream_t st);
std::vector<std::string> heads_param_names(const msq_model* m);
int add_first_visual(const float* dv, int64_t R, int Lt, int Lj, int H, float* gA, cudaStream_t st);
// train_kernels.cu: elementwise dropout sites
template <typename T> int dropout_rows(float* xf, T* xt, int64_t R, int group, int off, int n, int H, const Drop& d, cudaStream_t st);
int dropout_add(float* s, const float* resid, int64_t n, const Drop& d, cudaStream_t st);
template <typename T> int dropout_mask_copy(const float* g, T* gt, int64_t n, const Drop& d, cudaStream_t st);
template <typename T>
int heads_train(msq_model* m, const int64_t* sep, int64_t B, int N, const int32_t* target, const int64_t* pair_labels, float lam, float* grads,
                float* loss_out, cudaStream_t st);

}  // namespace msq
