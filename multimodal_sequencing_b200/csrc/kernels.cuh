// Internal launcher prototypes (host side).  Every launcher returns an MSQ_* status and enqueues on
// the given stream; nothing synchronises.
#pragma once
#include "common.cuh"
#include "dropout.cuh"

namespace msq {

// ---- api.cu : optional per-launch CUDA-event profile of the dominant (tcgen05 GEMM) kernel
void profile_mark(cudaStream_t st, bool end, double flops);

// ---- elementwise.cu
template <typename T>
int layernorm(const float* x, int64_t rows, int H, const float* gamma, const float* beta, float eps, float* out_f,
              T* out_t, int in_group, int out_group, int out_off, cudaStream_t st);
template <typename T>
int embed_ln(const int64_t* ids, const int64_t* tts, int64_t R, int Lt, int Lj, int H, const float* word, const float* pos,
             const float* type, const float* gamma, const float* beta, float eps, float* out_f, T* out_t, cudaStream_t st);
int vit_assemble(const float* patch, const int32_t* img_index, int64_t R, int il, int g2, int W, const float* cls,
                 const float* pos, const float* gamma, const float* beta, float eps, float* out, cudaStream_t st);
template <typename T> int im2col(const float* img, int64_t n, int S, int P, T* out, cudaStream_t st);
template <typename TI, typename TO>
int gather_rows(const TI* src, int64_t rows, int H, int group, int src_group, int off, TO* dst, cudaStream_t st);
template <typename TO> int pack_pad(const float* src, int64_t rows, int K, int Kp, TO* dst, cudaStream_t st);
// fp32 [rows, K] (leading dimension ld) -> three bf16 planes per row [hi(Kp) | mid(Kp) | lo(Kp)], zero padded to Kp columns
int pack_split3(const float* src, int64_t rows, int K, int ld, int Kp, bf16* dst, cudaStream_t st);

// ---- expand.cu : device-side pair expansion (process_inputs_for_berson.py:113-368)
int scan_steps(const int64_t* ids, int64_t B, int L, int N, int64_t cls_id, int64_t sep_id, int32_t* starts, int32_t* lens, int32_t* meta,
               cudaStream_t st);
int expand_pairs(const int64_t* ids, int64_t B, int L, int N, int Lt, int64_t cls_id, int64_t pad_id, const int32_t* starts,
                 const int32_t* lens, int64_t* out_ids, int64_t* out_mask, int64_t* out_tt, int64_t* out_sep, int32_t* img_index,
                 cudaStream_t st);

// ---- rn.cu : CLIP ModifiedResNet helpers (NHWC activations; convolutions run as im2col + GEMM)
int rn_fold(const float* w, const float* gamma, const float* beta, const float* mean, const float* var, int Cout, int Cin, int k,
            float* wf, float* bf, cudaStream_t st);
int rn_posadd(const float* xe, const float* ye, const float* te, int g, int F, float* out, cudaStream_t st);
template <typename T> int rn_im2col_stem(const float* img, int64_t n, int S, int Kp, T* out, cudaStream_t st);
template <typename T> int rn_im2col3(const T* x, int64_t n, int H, int W, int C, int Kp, T* out, cudaStream_t st);
template <typename T> int rn_avgpool2(const T* x, int64_t n, int H, int W, int C, T* out, cudaStream_t st);
template <typename T> int rn_relu_cast(float* x, int64_t n, int C, T* out, cudaStream_t st);   // C = row length (channels)
template <typename T>
int rn_tokens(const float* feat, const int32_t* img_index, int64_t R, int g2, int C, const float* pos, T* out, cudaStream_t st);
template <typename T> int rn_finish(const float* o, int64_t rows, int L, int E, const float* posadd, T* out, cudaStream_t st);

// ---- gemm_simt.cu : C[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ residual), fp32 FFMA accumulate
// epilogue modes of the tensor-core GEMM (deferred LayerNorm, see gemm_tc.cu)
enum { EPI_PLAIN = 0, EPI_LNFOLD = 1, EPI_RESLN = 2, EPI_DUALACT = 3 };

struct GemmArgs {
  const void* A;       // [M, lda]  (TA)
  const void* W;       // [N, ldw]  (TA)
  const float* bias;   // [N] or null
  const float* resid;  // [M, ldr] fp32 or null (added after act)
  void* C;             // [M, ldc]  (TO)
  float* C2;           // optional second fp32 output (same ldc) or null
  int64_t M;
  int N, K, lda, ldw, ldc, ldr;
  int act;
  // ---- deferred LayerNorm (gemm_tc only).  A LayerNorm "pending" on a stream Y is described by per-row partial sums
  // stats [M, sp, 2] = (sum y, sum y^2) over column groups, its gamma/beta and eps; it is never materialised:
  //   EPI_LNFOLD  A is the bf16 copy of the RAW stream, W carries gamma (W' = gamma_k W[n,k]), bias carries
  //               b + W beta, svec[n] = sum_k W'[n,k]:  C = act(rstd_m (acc - mu_m svec_n) + bias_n)
  //   EPI_RESLN   C = acc + bias + LN_pending(resid) (raw resid when stats_in is null); also writes the bf16 copy C2bf
  //               and the partial statistics of the rows it produced (stats_out [M, 2*ceil(N/256), 2]), i.e. the
  //               NEXT pending LayerNorm's input.  svec/beta hold gamma/beta of the LayerNorm pending on resid.
  //   EPI_DUALACT (bf16 in / out, fine-tuning forward of the up-projections): C = acc + bias (the pre-activation the
  //               backward pass needs) AND C2bf = act(acc + bias) (the operand of the next GEMM) from one accumulator read:
  //               the separate activation pass over [M, N] disappears
  // gemm_tc only, single-CTA path: "TN" operands for weight gradients.  A is [K, M] (row = contraction index, lda) and
  // W is [K, N] (ldw), both row-major, i.e. MN-major for the tensor core: C[M,N] = A^T W.  TMA boxes of 64 contraction
  // rows x 64 columns land in shared memory as the canonical MN-major SWIZZLE_128B atoms and are consumed through
  // MN-major UMMA descriptors -- no transposed copy of the activations / gradients is ever made.  K need not be a
  // multiple of 64 (out-of-range rows are zero-filled by the TMA unit).
  int tn = 0;
  // gemm_tc only: bf16x3 precision mode.  A and W rows are [hi(K) | lo(K)] bf16 (type bf16s; lda / ldw / ldc count logical
  // elements); C = A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T with fp32 accumulation, exact activations in the epilogue.
  // Output type float or bf16s (a split row [hi(ldc) | lo(ldc)]).  Plain epilogue only.
  // split = 2: THREE planes [hi | mid | lo] (24 significand bits = every fp32 value exactly) and six MMAs per product
  // (all a_i w_j with i + j <= 2): fp32-grade products on the tensor cores ("bf16x6"; the decoder's GEMMs).  fp32 out.
  int split = 0;
  // split = 2 only: run the three passes (0,1) (1,0) (0,0) over the top two planes of the three-plane rows -- bf16x3 products
  // (16 significand bits) from the SAME operand buffers at half the tensor work (the decoder in the bf16x3 mode)
  int trunc = 0;
  int mode = EPI_PLAIN;
  const float* svec = nullptr;
  const float* beta = nullptr;
  const float* stats_in = nullptr;
  int sp_in = 0;
  float ln_inv_dim = 0.f, ln_eps = 0.f;
  float* stats_out = nullptr;
  void* C2bf = nullptr;
  // gemm_tc only, plain epilogue: training-mode dropout of the dense output BEFORE the residual is added,
  // C = dropout(act(acc + bias)) + resid, element index row * N + col (the flat index nn.Dropout sees)
  Drop drop;
};
template <typename TA, typename TO> int gemm_simt(const GemmArgs& g, cudaStream_t st);

// ---- gemm_tc.cu : same contract, bf16 operands on tcgen05 tensor cores (TMA-fed, TMEM accumulators)
template <typename TO> int gemm_tc(const GemmArgs& g, cudaStream_t st);
int gemm_tc_selftest_supported();

// ---- gemm_ln.cu : GEMM + bias + residual + LayerNorm fused (cluster of N/256 CTAs), bf16 operands
bool gemm_ln_supported(int N, int K);
bool gemm_ln_enabled();
int gemm_ln(const bf16* A, int lda, const bf16* W, int ldw, const float* bias, const float* resid, int ldr, const float* gamma,
            const float* beta, float eps, float* C, bf16* C2, int64_t M, int N, int K, bool raw32, cudaStream_t st);

// ---- attention.cu : ctx[r, t, h*64 + d] = softmax(q k^T / sqrt(d) + mask) v over the L tokens of row-group r
// drop: training-mode dropout of the probabilities (site A; element index ((r*heads + h)*L + q)*L + k); default = off
template <typename T>
int attention(const T* qkv, int64_t R, int L, int heads, int dhead, float scale, const float* key_mask_add, int mask_ld,
              int mask_len, T* ctx, cudaStream_t st, const Drop& drop = Drop(), float* lse_out = nullptr, bool* lse_written = nullptr);
// lse_out [R*heads, L] (optional, training forward): the row log-sum-exp of (scale q.k + mask); *lse_written says whether the
// kernel that ran emits it (the tcgen05 bf16 kernel does; otherwise the backward pass recomputes it)

// ---- pooling.cu
template <typename T>
int token_pool(const T* tt, const float* x, int64_t R, int Lt, int Lj, int H, const float* w2, const float* b2,
               const int64_t* sep, const float* w_rel, const float* b_rel, float* mix, float* rel6, cudaStream_t st,
               const Drop& drop = Drop());   // drop: training-mode dropout of the token-attention probabilities (site H)
int edge_pool(const float* mix, const float* x, const float* rel6, int64_t B, int N, int Lj, int H, const float* w_in2,
              float* sents, float* r0, int r0_ld, float* cls_mat, float* score_mat, float* his1, float* his2, float* cls_out,
              cudaStream_t st);
int para_attention(const float* qkv, int64_t B, int N, int heads, int H, float* ctx, cudaStream_t st, const Drop& drop = Drop());
int para_finish(const float* sents, const float* para, int64_t B, int N, int H, float* h0, float* keyin, cudaStream_t st);

// ---- decode.cu
struct DecodeWeights {
  const float* whh_t;   // [H][4H] gate-interleaved: col 4*u+g
  const float* wq_t;    // [H][H]
  const float* bq;      // [H]
  const float* wt;      // [H] tanh_linear.weight
  float bt;             // tanh_linear.bias
  // tensor-core decode (bf16x6): [W_q ; W_hh gate-interleaved] as [5H rows][3 planes x H] bf16, and [b_q ; 0] (5H)
  const bf16* wcat3 = nullptr;
  const float* bcat = nullptr;
};
struct DecodeIO {
  const float* xg;     // [B, N+1, 4H] gate-interleaved input projections; row N = bias only (t = 0)
  const float* t4;     // [B, N*N, 4H] = (A1|A2|F|G) blocks of H
  const float* key0;   // [B, N, H]
  const float* h0;     // [B, H]
  int64_t B;
  int N, W, H;
  int32_t* perm;       // [B, N] out
  int32_t* trace_ix;   // optional [B, N-1, W] flat index beam*N+tok per step (or -1)
  float* trace_cost;   // optional [B, N-1, W]
  float* trace_logp;   // optional [B, N-1, W, N]
  const int32_t* forced;  // optional [B, N]: teacher-forced picks (W must be 1)
  float* final_cost;      // optional [B]: cost of the best hypothesis
  void* scratch = nullptr;  // beam_search_scratch_bytes(B, N, W, H) bytes of device memory (tiled decode; null -> fused kernel)
  int tc = 0;               // 1: recurrent GEMMs on the tensor cores (bf16x6, needs DecodeWeights::wcat3); 0: fp32 FFMA
  int trunc = 0;            // tc only: bf16x3 products on the top two planes instead of bf16x6 (GemmArgs::trunc)
};
size_t beam_search_state_bytes(int64_t B, int W, int H);             // search state (either decode form)
size_t beam_search_scratch_bytes(int64_t B, int N, int W, int H);    // state + three-plane copies of sents_ext / r0 (tensor-core form)
int decode_trunc(int precise);   // 1: decoder GEMMs as bf16x3 on the top two planes (default; MSQ_DEC_X3=0: bf16x6, fp32-grade products)
bool decode_tc_enabled();                                             // MSQ_DECODE_TC=0 keeps the fp32 FFMA GEMMs in every mode
int beam_search(const DecodeWeights& w, const DecodeIO& io, cudaStream_t st);
struct StepIO {
  float* rela;                 // [Wb,N,N,H+2] zeroed in place where rela_mask == 0
  const uint8_t* rela_mask;    // [Wb,N,N]
  const float *hist1, *hist2;  // [Wb,N,N,H+2]
  const uint8_t *l1, *l2;      // [Wb,N,N]
  int Wb, N, H, Kp4;
  float* pw;                   // [Wb*N, Kp4] out
};
int decode_step_parts(const StepIO& io, cudaStream_t st);
int decode_step_lstm(const float* gates, const float* c_in, int64_t Wb, int H, float* h_out, float* c_out, cudaStream_t st);
int decode_step_score(const float* q, const float* keys, const float* key0, const uint8_t* pointed, const float* wt, float bt,
                      int64_t Wb, int N, int H, float* logp, cudaStream_t st);

int pointer_p1(const float* enc, const float* cls, const int64_t* y, const float* W1, const float* W2, const float* V,
               const float* Wih, const float* Whh, const float* bih, const float* bhh, int64_t B, int N, int H, int U, float* preds,
               float* ce_scratch, float* loss, cudaStream_t st);

// ---- train_kernels.cu : backward / optimizer kernels of the fine-tuning path (train.cu orchestrates them)
template <typename TI, typename TO>
int transpose_pad(const TI* src, int64_t M, int N, int ld, int64_t Mp, TO* dst, int act, cudaStream_t st);
template <typename T> int rowsum_accum(const T* a, int rows, int64_t ld, int64_t n, float* out, cudaStream_t st);
size_t colsum_scratch_floats(int N);
int colsum_accum_bf16(const bf16* G, int64_t M, int N, int ld, float* out, float* scratch, cudaStream_t st);
template <typename T> int act_fwd(const T* u, int64_t n, int act, T* h, cudaStream_t st);
template <typename T> int act_bwd(const T* dh, const T* u, int64_t n, int act, T* du, cudaStream_t st);
size_t ln_bwd_scratch_floats(int H);
template <typename T>
int ln_bwd(const float* dy, const float* x, const float* add, int64_t rows, int H, const float* gamma, float eps, float* dx, T* dx_t,
           float* dgamma, float* dbeta, float* scratch, int in_group, int out_group, int out_off, cudaStream_t st,
           const Drop& drop = Drop());   // drop: dx_t = dx . mask (the dense output this LayerNorm normalised was dropped)
int embed_ln_bwd(const float* dy, const int64_t* ids, const int64_t* tts, int64_t R, int Lt, int Lj, int H, const float* word,
                 const float* pos, const float* type, const float* gamma, float eps, float* dword, float* dpos, float* dtype,
                 float* dgamma, float* dbeta, float* scratch, int pad0, cudaStream_t st);
int vit_assemble_bwd(const float* dy, const float* patch, const int32_t* img_index, int64_t R, int il, int g2, int W, const float* cls,
                     const float* pos, const float* gamma, float eps, float* dpatch, float* dcls, float* dpos, float* dgamma,
                     float* dbeta, float* scratch, cudaStream_t st);
size_t attention_bwd_scratch_floats(int64_t R, int L, int heads);
template <typename T>
int attention_bwd(const T* qkv, const T* dctx, int64_t R, int L, int heads, float scale, const float* key_mask_add, int mask_ld,
                  int mask_len, T* dqkv, float* scratch, cudaStream_t st, const Drop& drop = Drop());
// attention_bwd_mma.cu: bf16 tensor-core (mma.sync) version; ctx = the forward's output (for D_i = dO_i . O_i)
bool attention_bwd_mma_supported(int L);
int attention_bwd_mma(const bf16* qkv, const bf16* ctx, const bf16* dctx, int64_t R, int L, int heads, float scale, const float* key_mask_add,
                      int mask_ld, int mask_len, bf16* dqkv, float* scratch, cudaStream_t st, const Drop& drop = Drop(),
                      const float* lse_fwd = nullptr);   // scratch: attention_bwd_scratch_floats; lse_fwd: the forward's row log-sum-exp [R*heads, L] or null
// attention_bwd_tc.cu: dK / dV on tcgen05 (lse / dsum: per-query vectors [R*heads, L] written by the dQ kernel)
bool attention_bwd_dkv_tc_supported(int L);
int attention_bwd_dkv_tc(const bf16* qkv, const bf16* dctx, int64_t R, int L, int heads, float scale, const float* mask_add, int mask_ld,
                         int mask_len, bf16* dqkv, const float* lse, const float* dsum, cudaStream_t st, const Drop& drop);
int mask_add_from_int(const int64_t* mask, int64_t n, float* out, cudaStream_t st);
int scatter_rows(const float* src, int64_t rows, int H, int group, int dst_group, int off, float* dst, cudaStream_t st);
size_t grad_norm_scratch_floats();
int grad_norm_clip(const float* g, int64_t n, float max_norm, float grad_scale, float* scratch, cudaStream_t st);
struct AdamChunk { float* p; int64_t off; int32_t n; int32_t decay; };
int adamw_update_multi(const AdamChunk* chunks_dev, int nchunks, const float* g, float* m, float* v, float lr, float b1, float b2, float eps, float wd,
                       int64_t step, const float* coef, cudaStream_t st);
int adamw_update(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, float wd, int64_t step,
                 const float* coef, cudaStream_t st);

}  // namespace msq
