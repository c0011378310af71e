/* msq_b200.h — C ABI of the B200-native (sm_100a) multimodal step-ordering hot path.
 *
 * Drop-in boundary for telin0411/multimodal_sequencing.  The reference is pure Python/PyTorch: the
 * functions below are what its nn.Module methods bind to (ctypes stubs in INTEGRATION.md; the host
 * mirror lives in multimodal_sequencing_b200/).  Conventions:
 *   - every function returns 0 on success, non-zero on failure; msq_last_error() gives the message
 *     (the Python side raises RuntimeError, matching the reference's exception-based error style);
 *   - pointers named *_dev are device pointers (caller-owned, e.g. torch tensors' data_ptr()),
 *     pointers named *_host are host pointers; `stream` is a cudaStream_t passed as void*;
 *   - nothing synchronises unless stated; outputs are valid after the stream is synchronised;
 *   - integer inputs are int64 (torch.long) exactly as the reference passes them.
 * No CPU fallback exists: without a CUDA device every compute entry point fails.
 */
#ifndef MSQ_B200_H
#define MSQ_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct msq_model msq_model;

/* Model geometry.  Mirrors BertConfig (models/berson/configuration_bert.py:77-90), the CLIP visual
 * tower constructor (models/CLIP/clip/model.py:308-323) and the BERSON hyper-parameters hard-coded
 * in trainers/train.py:2012-2022. */
typedef struct msq_config {
  int32_t hidden;        /* 768 */
  int32_t layers;        /* 12  */
  int32_t heads;         /* 12 (head dim must be 64) */
  int32_t inter;         /* 3072 */
  int32_t vocab;         /* 30522 */
  int32_t max_pos;       /* 512 */
  int32_t type_vocab;    /* 2 */
  int32_t vit_width;     /* 768; 0 = text-only inner BertModel (models/berson/modeling_bert.py:563) */
  int32_t vit_layers;    /* 12 */
  int32_t vit_patch;     /* 32 */
  int32_t vit_res;       /* 224 */
  int32_t para_heads;    /* 8 */
  int32_t para_ff;       /* 3072 */
  int32_t para_layers;   /* 2 */
  int32_t precise;       /* 0: bf16 tcgen05 tensor-core encoder (fastest, ~1e-2 of fp32);
                          * 1: fp32 FFMA encoder (CUDA cores; 1e-5 parity mode, slow);
                          * 2: "bf16x3" tcgen05 encoder: every operand is carried as hi + lo bf16 (16 significand bits),
                          *    every product is a_hi*w_hi + a_lo*w_hi + a_hi*w_lo with fp32 accumulation, activations
                          *    are exact (erff / tanhf): outputs within ~1e-5 of fp32 at tensor-core speed.  Evaluation
                          *    only; ViT, ModifiedResNet (rn_width % 64 == 0) and text-only models. */
  int32_t reserved;      /* flags.  bit 0 "cls_pooler": the inner encoder is a HuggingFace AutoModel (trainers/train.py:1928-1933)
                          * whose outputs[1] = tanh(pooler.dense(seq[:,0])): BertForOrdering.encode takes THAT as the pair's
                          * CLS vector (modeling_bert.py:1315) instead of seq[:,0].  Evaluation entry points only. */
  /* CLIP ModifiedResNet backbone ("RN50", models/CLIP/clip/model.py:128-187), the reference's wired default
   * (param.py VISUAL_CONFIG.clip_model_name).  rn_width != 0 selects it; then vit_width must hold the tower's
   * OUTPUT feature size 2*rn_embed (model.py:106 concatenates the pooled tokens with themselves), vit_layers 0,
   * vit_patch 32 (total stride), vit_res the image size. */
  int32_t rn_width;      /* 64 */
  int32_t rn_blocks[4];  /* 3,4,6,3 */
  int32_t rn_embed;      /* 1024 (attnpool output_dim) */
  int32_t reserved2[2];
} msq_config;

/* Hard limits of this build (violations fail with a message, nothing is silently truncated):
 *   manual length N in [2,16]; beam width W in [1,16]; hidden size H a multiple of 128, <= 1024; head dim = 64
 *   (heads * 64 == hidden); vit_width a multiple of 128, <= 1024; inter % 16 == 0 (bf16x3: % 64); joint sequence
 *   length per pair (text + visual tokens) <= 320 (tcgen05 attention: <= 256); text tokens per pair Lt <= max_pos;
 *   manuals are encoded in micro-batches of MSQ_CHUNK_MANUALS (environment, default 32); the ModifiedResNet tower needs
 *   rn_width % 16 == 0 (bf16x3: % 64), rn_embed % 32 == 0, image size a multiple of 32; fine-tuning (msq_train_*) covers the
 *   ViT, ModifiedResNet and text-only models in precise 0 / 1. */
const char* msq_last_error(void);
int msq_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t msq_launch_count(void);
/* 1 if the tcgen05/TMA GEMM path can be used on the current device (sm_100), else 0 */
int msq_tc_available(void);

/* Per-launch CUDA-event timing of the dominant kernel (the tcgen05 GEMM), for bench.py's roofline line:
 * events are recorded on the launching stream around every GEMM launch while enabled; msq_profile_read
 * synchronises and returns the summed duration (ms), algorithmic FLOPs (2*M*N*K) and launch count. */
int msq_profile_enable(int32_t on);
int msq_profile_read(double* total_ms, double* total_flops, int64_t* launches);

/* ---- model lifetime ------------------------------------------------------------------------ */
int msq_model_create(const msq_config* cfg, msq_model** out);
void msq_model_destroy(msq_model* m);
/* Register one fp32 parameter by its reference state_dict key, e.g.
 * "bert.encoder.layer.0.attention.self.query.weight".  Data is copied (device -> device). */
int msq_model_set_weight(msq_model* m, const char* name, const float* data_dev, int64_t numel, void* stream);
/* Build the packed device-side weights (fused QKV, bf16 copies, gate-interleaved LSTM, pw_k blocks).
 * Fails with the list of missing keys if the state_dict was incomplete. */
int msq_model_pack(msq_model* m, void* stream);

/* In-place refresh for callers that keep the fp32 masters themselves (the reference trainer's optimizer updates
 * nn.Parameter.data, which no version counter sees): overwrite the registered copy of `name` (same element count; no
 * allocation), then msq_model_refresh re-derives every packed copy (fused QKV, bf16 / split-bf16, folded LayerNorm, LSTM
 * repacks, and the training state's W^T operands).  ~1 GB of device-to-device copies + the pack kernels: a few ms. */
int msq_model_update_weight(msq_model* m, const char* name, const float* data_dev, int64_t numel, void* stream);
int msq_model_refresh(msq_model* m, void* stream);

/* ---- encoders -------------------------------------------------------------------------------
 * msq_vit_forward       CLIP VisualTransformer.forward, pair-joint variant, skip_last_layer=True
 *                       (models/CLIP/clip/model.py:262-305).  images_dev [n_img,3,S,S] fp32 are UNIQUE
 *                       images; img_index_dev [R*2] int32 maps (pair row, slot) -> image.  out [R,1+2g^2,W].
 * msq_inner_forward     LXRTModel.forward BERSON mode (models/CLIP/src/lxrt/modeling.py:1513-1598) or,
 *                       when the model is text-only, BertModel.forward
 *                       (models/berson/modeling_bert.py:613-663).  ids/tt/mask [R,Lt] int64.
 *                       lang_dev [R,Lt,H], visn_dev [R,Lv,H] (or NULL), pooled_dev [R,H] (or NULL). */
int msq_vit_forward(msq_model* m, const float* images_dev, int64_t n_img, const int32_t* img_index_dev, int64_t R,
                    float* out_dev, void* stream);
int msq_inner_forward(msq_model* m, const int64_t* ids_dev, const int64_t* tt_dev, const int64_t* mask_dev, int64_t R,
                      int32_t Lt, const float* images_dev, int64_t n_img, const int32_t* img_index_dev, float* lang_dev,
                      float* visn_dev, float* pooled_dev, void* stream);

/* ---- BertForOrdering.encode (models/berson/modeling_bert.py:1239-1366) ------------------------
 * B manuals of N steps, P = N(N-1) pair rows each.  sep_dev [B*P,2] int64.  Outputs (any may be
 * NULL): sents [B,N,H], para [B,N,H], h0 [B,H], key [B,N,H], cls [B*P,H], cls_mat [B,N,N,H],
 * cls_score [B*P,2], score_mat/his1/his2 [B,N,N,2], top_vec [B*P,Lt,H]. */
typedef struct msq_encode_out {
  float *sents, *para, *h0, *key, *cls, *cls_mat, *cls_score, *score_mat, *his1, *his2, *top_vec;
} msq_encode_out;
int msq_encode(msq_model* m, const int64_t* ids_dev, const int64_t* tt_dev, const int64_t* mask_dev,
               const int64_t* sep_dev, int64_t B, int32_t N, int32_t Lt, const float* images_dev, int64_t n_img,
               const int32_t* img_index_dev, const msq_encode_out* out, void* stream);

/* ---- BertForOrdering._forward loss VALUE (modeling_bert.py:943-1174: teacher-forced pointer NLL / (N-1) + lam *
 * pairwise NLL / P, batch mean), value only (eval mode; msq_train_step below is the training twin).  ground_truth_dev [B,N] int32 =
 * the target order, pairwise_labels_dev [B,P] int64, perm_scratch_dev [B,N] int32 workspace, loss_dev 1 float. */
int msq_training_loss(msq_model* m, const int64_t* ids_dev, const int64_t* tt_dev, const int64_t* mask_dev,
                      const int64_t* sep_dev, int64_t B, int32_t N, int32_t Lt, const float* images_dev, int64_t n_img,
                      const int32_t* img_index_dev, const int32_t* ground_truth_dev, const int64_t* pairwise_labels_dev,
                      float lam, int32_t* perm_scratch_dev, float* loss_dev, void* stream);

/* ---- pointer decoder + beam search (modeling_bert.py:1368-1402, 1411-1552; generator.py:15-38) --
 * Consumes encode outputs (device, fp32): sents/key [B,N,H], h0 [B,H], cls_mat [B,N,N,H],
 * score_mat [B,N,N,2].  perm_dev [B,N] int32 receives the predicted order.  Optional traces:
 * trace_ix [B,N-1,W] int32 (flat beam*N+step picks, -1 padded), trace_cost [B,N-1,W],
 * trace_logp [B,N-1,W,N]. */
int msq_beam_search(msq_model* m, const float* sents_dev, const float* key_dev, const float* h0_dev,
                    const float* cls_mat_dev, const float* score_mat_dev, int64_t B, int32_t N, int32_t beam,
                    int32_t* perm_dev, int32_t* trace_ix_dev, float* trace_cost_dev, float* trace_logp_dev, void* stream);

/* One decode step with the reference's materialised per-beam tensors: BertForOrdering.step
 * (modeling_bert.py:1368-1402).  prev_y/h/c [Wb,H]; key0 [N,H]; pointed_mask [Wb,N] uint8;
 * rela_vec [Wb,N,N,H+2] is zeroed IN PLACE where rela_mask == 0 (as the reference does); hist1/hist2 like
 * rela_vec; rela/l1/l2 masks [Wb,N,N] uint8.  Outputs h,c [Wb,H], logp [Wb,N]. */
int msq_decode_step(msq_model* m, const float* prev_y_dev, const float* h_dev, const float* c_dev, const float* key0_dev,
                    const uint8_t* pointed_mask_dev, float* rela_vec_dev, const uint8_t* rela_mask_dev,
                    const float* hist1_dev, const float* hist2_dev, const uint8_t* l1_mask_dev, const uint8_t* l2_mask_dev,
                    int32_t Wb, int32_t N, float* h_out_dev, float* c_out_dev, float* logp_out_dev, void* stream);

/* models/pointer_module.py, p1 variant: LSTMPointerModule.forward (pointer_module.py:690-749) over LSTMDecoder /
 * LSTMAttention (616-678).  enc [B,N,H] = hidden states at the [CLS] positions, cls [B,H] = sequence_output[:,0],
 * y [B,N] int64 labels; W1,W2 [U,H], V [U], W_ih [4H,2H], W_hh [4H,H], biases [4H] (torch LSTM layout).
 * preds [B,N] (float, as the reference stores them), ce_scratch [B], loss = sum_t mean_b CE_t / B (1 float). */
int msq_pointer_p1(const float* enc_dev, const float* cls_dev, const int64_t* y_dev, const float* w1_dev, const float* w2_dev,
                   const float* v_dev, const float* wih_dev, const float* whh_dev, const float* bih_dev, const float* bhh_dev,
                   int64_t B, int32_t N, int32_t H, int32_t U, float* preds_dev, float* ce_scratch_dev, float* loss_dev,
                   void* stream);

/* ---- whole path on device-resident inputs: encode + beam search ------------------------------- */
int msq_order_manuals_dev(msq_model* m, const int64_t* ids_dev, const int64_t* tt_dev, const int64_t* mask_dev,
                          const int64_t* sep_dev, int64_t B, int32_t N, int32_t Lt, const float* images_dev,
                          int64_t n_img, const int32_t* img_index_dev, int32_t beam, int32_t* perm_dev, void* stream);

/* ---- whole path from HOST buffers (berson_pointer_network, modeling_bert.py:1405-1408, batched):
 * copies the inputs host->device, runs encode + beam search, copies perm back and synchronises the
 * stream.  Host buffers should be pinned for full H2D bandwidth. */
int msq_order_manuals_host(msq_model* m, const int64_t* ids_host, const int64_t* tt_host, const int64_t* mask_host,
                           const int64_t* sep_host, int64_t B, int32_t N, int32_t Lt, const float* images_host,
                           int64_t n_img, const int32_t* img_index_host, int32_t beam, int32_t* perm_host, void* stream);

/* ---- device-side pair expansion (models/berson/process_inputs_for_berson.py:113-368, 246-261, 82-97) -------------------
 * msq_scan_steps: positions of the N [CLS] / [SEP] tokens of every manual row ids_dev [B,L] -> starts_dev / lens_dev (int32
 *   [B,N]); meta_dev = 2 int32 of scratch; *lt_out = padded pair length max_{i!=j}(len_i + len_j), read back with one 8-byte
 *   device->host copy (the only synchronisation).  Fails when a manual does not hold exactly N [CLS]..[SEP] steps (the
 *   reference asserts the same).
 * msq_expand_pairs: writes the P = N(N-1) ordered pair rows of every manual: out_ids / out_mask / out_tt [B,P,Lt] int64,
 *   out_sep [B,P,2] int64, img_index [B,P,2] int32 (manual-order images, may be NULL).  Bit-identical to the reference's host
 *   code, quirks included (mask of a padded position = pad_id; token types all zero when cls_id == 0).
 * msq_order_manuals_raw_host: berson_pointer_network from the DataLoader tuple: token rows ids_host [B,L] int64 and step
 *   images images_host [B*N,3,S,S] fp32 in host memory -> perm_host [B,N]; expansion happens on the device. */
int msq_scan_steps(const int64_t* ids_dev, int64_t B, int32_t L, int32_t N, int64_t cls_id, int64_t sep_id, int32_t* starts_dev,
                   int32_t* lens_dev, int32_t* meta_dev, int32_t* lt_out, void* stream);
int msq_expand_pairs(const int64_t* ids_dev, int64_t B, int32_t L, int32_t N, int32_t Lt, int64_t cls_id, int64_t pad_id,
                     const int32_t* starts_dev, const int32_t* lens_dev, int64_t* out_ids_dev, int64_t* out_mask_dev,
                     int64_t* out_tt_dev, int64_t* out_sep_dev, int32_t* img_index_dev, void* stream);
int msq_order_manuals_raw_host(msq_model* m, const int64_t* ids_host, int64_t B, int32_t L, int32_t N, int64_t cls_id,
                               int64_t sep_id, int64_t pad_id, const float* images_host, int32_t beam, int32_t* perm_host,
                               void* stream);

/* ---- fine-tuning of the inner encoder (BASELINE config 4) -------------------------------------
 * What the reference gets from autograd + transformers.AdamW (trainers/train.py:172-190, 340-363) for the inner model
 * LXRTModel / BertModel (lxrt/modeling.py:1513-1598, modeling_bert.py:563-663) and the CLIP ViT tower
 * (clip/model.py:190-305), and for the BERSON heads + loss (msq_train_step).  Dropout is not applied (p = 0).
 *
 * Gradients live in ONE flat fp32 device buffer owned by the caller (msq_train_grad_numel elements, 256-byte aligned,
 * zeroed by the caller before the first backward of an optimizer step; backward ACCUMULATES).  Parameter i occupies
 * [offset, offset + numel) of it (msq_train_param_info; name = the reference state_dict key).  Data-parallel
 * fine-tuning all-reduces that one buffer over NCCL and passes grad_scale = 1 / world_size to msq_adamw_step. */
int64_t msq_train_param_count(msq_model* m, void* stream);   /* -1 on error */
int64_t msq_train_grad_numel(msq_model* m, void* stream);    /* -1 on error */
int msq_train_param_info(msq_model* m, int64_t i, const char** name, int64_t* offset, int64_t* numel, int32_t* decay);
/* current fp32 master copy of a registered weight (what save_pretrained would write) */
int msq_train_read_param(msq_model* m, const char* name, float* out_dev, int64_t numel, void* stream);
/* msq_inner_forward in training mode: same outputs, and the activations the backward pass needs are recorded inside the
 * model.  images_dev must stay valid until msq_inner_backward has been enqueued. */
int msq_inner_forward_train(msq_model* m, const int64_t* ids_dev, const int64_t* tt_dev, const int64_t* mask_dev, int64_t R,
                            int32_t Lt, const float* images_dev, int64_t n_img, const int32_t* img_index_dev, float* lang_dev,
                            float* visn_dev, void* stream);
/* Backward of the recorded forward: d_lang [R,Lt,H], d_visn [R,Lv,H] (either may be NULL = zero) -> grads_dev += dL/dparam. */
int msq_inner_backward(msq_model* m, const float* d_lang_dev, const float* d_visn_dev, float* grads_dev, void* stream);
/* One fine-tuning forward + backward of BertForOrdering.forward(inputs) -> (loss,) in train mode (modeling_bert.py:937-941,
 * 943-1174, default objectives: teacher-forced pointer NLL / (N-1) + lam * pairwise NLL / P, batch mean) followed by
 * loss.backward() (trainers/train.py:340-351): inputs as msq_training_loss; loss_dev (1 float, may be NULL) receives the
 * loss, grads_dev += dL/dparam for every parameter of the path, encoder AND heads (the parameter table then also lists the
 * head parameters the reference gives a gradient to). */
int msq_train_step(msq_model* m, const int64_t* ids_dev, const int64_t* tt_dev, const int64_t* mask_dev, const int64_t* sep_dev,
                   int64_t B, int32_t N, int32_t Lt, const float* images_dev, int64_t n_img, const int32_t* img_index_dev,
                   const int32_t* ground_truth_dev, const int64_t* pairwise_labels_dev, float lam, float* grads_dev, float* loss_dev,
                   void* stream);
/* Dropout of the fine-tuning path.  The reference trains with hidden_dropout_prob = attention_probs_dropout_prob = 0.1
 * (BertConfig) and args.para_dropout in the paragraph encoder; sites: embeddings, visn_fc output, attention probabilities,
 * attention-output and FFN-output dense (lxrt/modeling.py:369,601,419,437,491; modeling_bert.py:179,228,253,319), the
 * token-attention probabilities of HierarchicalAttention (modeling_bert.py:735), and the paragraph encoder's attention,
 * context and feed-forward dropouts (neural.py:228,31-32; encoder.py:28).  msq_train_set_dropout sets the three
 * probabilities (in [0,1); 0 = off, the default) and a seed for all later training forwards.  Masks are never stored: an
 * element is kept iff hash(seed, forward counter, site, element index) >= p * 2^32 (csrc/dropout.cuh), so the backward pass
 * regenerates them, and a CPU checker can too (msq_train_dropout_step returns the counter of the last training forward;
 * oracle/dropout.py is the numpy twin).  Evaluation entry points never apply dropout. */
int msq_train_set_dropout(msq_model* m, float p_hidden, float p_attn, float p_para, uint32_t seed, void* stream);
int64_t msq_train_dropout_step(msq_model* m);
/* BatchNorm of the ModifiedResNet tower inside a training step (models/CLIP/clip/model.py:10-53, 128-187).  Default
 * (use_running_stats = 0): nn.BatchNorm2d in train() mode, the mode trainers/train.py fine-tunes in -- statistics of the
 * batch of MATERIALISED pair images (mean and biased variance over N, H, W; the tower itself runs on the unique images with
 * their pair multiplicities as weights), gradients flow through the statistics.  use_running_stats = 1: eval() semantics
 * (frozen running statistics) with the same backward pass otherwise.  In batch-statistics mode every training forward also
 * moves the registered running_mean / running_var masters (momentum 0.1, unbiased batch variance, as nn.BatchNorm2d does);
 * msq_train_read_param returns them, the packed evaluation weights pick them up at the next re-pack (msq_adamw_step /
 * msq_model_refresh).  No-op for other backbones. */
int msq_train_set_bn_mode(msq_model* m, int32_t use_running_stats, void* stream);
/* Optional time-contrastive objective (models/berson/modeling_bert.py:1176-1216; args.additional_wrapper_level_objectives
 * contains "time_contrastive"): the NEXT msq_train_step adds weight * nn.TripletMarginLoss(margin=1, p=2)(a, p, n) with
 * a, p, n = sents[b, triplets[b, 0..2]] (device int32 [B,3]: anchor, positive, negative sentence indices, drawn by the caller
 * as the reference draws them) to the loss and its gradients.  The reference's weight is 0.1.  One-shot per step. */
int msq_train_set_triplets(msq_model* m, const int32_t* triplets_dev, int64_t B, float weight, void* stream);
/* Optional image pairwise objective (args.multimodal_loss; models/berson/modeling_bert.py:897-898, 1359-1364, 1218-1225):
 * while on, every msq_train_step adds lam * mean_b sum_p NLL(softmax(pairwise_relationship(img_projection(visn[p, 0]))),
 * pairwise_label_p) / P and the gradients of img_projection.*, the shared pairwise_relationship.* and the first visual token of
 * every pair row.  img_projection.weight [H, H] / .bias [H] must be registered (msq_model_set_weight) before the first
 * training call; the reference's Linear(v_feature_size, H) is fed the H-d token, i.e. it only runs when v_feature_size == H. */
int msq_train_set_multimodal_loss(msq_model* m, int32_t on, void* stream);

/* Data-parallel overlap: msq_train_step records one CUDA event per REGION of the flat gradient buffer at the moment that
 * region is final (BERSON heads, BERT layers top -> bottom, embeddings, visn_fc, ViT blocks top -> bottom, ViT stem).
 * msq_train_ready_count = regions of the last step, in completion order; msq_train_ready_info = element range [begin, end)
 * of region i (regions are disjoint and cover every parameter); msq_train_ready_wait makes `stream` wait for region i, so
 * the caller can all-reduce it on a side stream while the backward pass of the earlier layers is still running
 * (what DistributedDataParallel's bucketing does for trainers/train.py:217-221). */
int64_t msq_train_ready_count(msq_model* m);
int msq_train_ready_info(msq_model* m, int64_t i, int64_t* begin, int64_t* end);
int msq_train_ready_wait(msq_model* m, int64_t i, void* stream);
/* torch.nn.utils.clip_grad_norm_(max_grad_norm) (<= 0: no clipping) over grad_scale * grads, then one step of
 * transformers.AdamW (correct_bias=True; weight decay skipped for names containing "bias" or "LayerNorm.weight",
 * train.py:172-181), then every packed copy of the weights is rebuilt.  norm_out_dev (2 floats or NULL) receives the
 * gradient norm and the factor applied to the gradients. */
int msq_adamw_step(msq_model* m, const float* grads_dev, float lr, float beta1, float beta2, float eps, float weight_decay,
                   float max_grad_norm, float grad_scale, float* norm_out_dev, void* stream);

/* ---- building blocks exposed for kernel-level parity tests and the bench's roofline lines ------ */
/* C[M,N] = act(A[M,K] W[N,K]^T + bias) + resid.  dtype: 0 = fp32 in/out (FFMA), 1 = bf16 in / fp32 out
 * (tcgen05), 2 = bf16 in / bf16 out (tcgen05), 3 / 4 = bf16 in on the FFMA kernel (fp32 / bf16 out), 5 = "TN" weight-gradient
 * shape on tcgen05: A [K,M] and W [K,N] bf16 row-major (MN-major operands, no transposes), C [M,N] fp32 = A^T W + resid.  act: 0 none, 1 erf-GELU, 2 QuickGELU, 3 tanh, 4 tanh-GELU. */
int msq_gemm(int32_t dtype, const void* A_dev, const void* W_dev, const float* bias_dev, const float* resid_dev,
             void* C_dev, int64_t M, int32_t N, int32_t K, int32_t act, void* stream);
/* tcgen05 GEMM with a DEFERRED LayerNorm in its epilogue (bf16 operands; kernel-level entry used by tests and
 * scripts/kernel_bench.py; the encoders call the same code internally).  The LayerNorm between two sub-layers
 * (lxrt/modeling.py:432,486; clip/model.py:221-225) is never materialised:
 *   mode 1 (fold)      C = act(rstd_m (A W'^T - mu_m svec) + bias)   A = bf16 copy of the RAW stream, W' = gamma*W,
 *                      bias = b + W beta, svec[n] = sum_k W'[n,k]; mu / rstd from stats_in [M, sp_in, 2] partial sums
 *   mode 2 (residual)  C = A W^T + bias + LN(resid) in fp32 (raw resid when stats_in is NULL; gamma/beta in
 *                      svec_or_gamma/beta), plus C2bf = bf16(C) and stats_out [M, 2*ceil(N/256), 2] = per-row partial
 *                      (sum, sum of squares) of C over 128-column groups.
 *   mode 3 (dual)      out_bf16 = 1: C = bf16(A W^T + bias) AND C2bf = bf16(act(A W^T + bias)) from one accumulator read (the
 *                      fine-tuning forward of the up-projections keeps the pre-activation for the backward pass and feeds the
 *                      activation to the next GEMM); act: 1 erf-GELU, 2 QuickGELU, 3 tanh, 5 ReLU; no residual. */
int msq_gemm_deferred_ln(int32_t mode, int32_t out_bf16, const void* A_dev, const void* W_dev, const float* bias_dev,
                         const float* resid_dev, const float* svec_or_gamma_dev, const float* beta_dev,
                         const float* stats_in_dev, int32_t sp_in, int32_t ln_dim, float eps, void* C_dev, void* C2bf_dev,
                         float* stats_out_dev, int64_t M, int32_t N, int32_t K, int32_t act, void* stream);
/* Fused v = A W^T + bias + resid ; y = LayerNorm(v) (tcgen05, cluster of N/256 CTAs; N in {256,512,768}, K % 64 == 0).
 * A [M,K], W [N,K] bf16; resid [M,N] fp32 (may alias C).  C [M,N] fp32 <- (raw32 ? v : y);  C2 [M,N] bf16 <- y. */
int msq_gemm_ln(const void* A_dev, const void* W_dev, const float* bias_dev, const float* resid_dev, const float* gamma_dev,
                const float* beta_dev, float eps, float* C_dev, void* C2_dev, int64_t M, int32_t N, int32_t K, int32_t raw32,
                void* stream);
/* dtype 0 fp32, 1 bf16.  y = LN(x) (x always fp32 [rows,H]); writes out_dev in dtype. */
int msq_layernorm(int32_t dtype, const float* x_dev, int64_t rows, int32_t H, const float* gamma_dev,
                  const float* beta_dev, float eps, void* out_dev, void* stream);
/* qkv [R*L, 3*heads*64] -> ctx [R*L, heads*64]; mask_add_dev [R, mask_len] additive key mask or NULL. */
int msq_attention(int32_t dtype, const void* qkv_dev, int64_t R, int32_t L, int32_t heads, float scale,
                  const float* mask_add_dev, int32_t mask_len, void* ctx_dev, void* stream);
int msq_f32_to_bf16(const float* src_dev, void* dst_dev, int64_t n, void* stream);
/* fp32 [rows,K] -> split-bf16 rows [hi(K) | lo(K)] (bf16 x 2K per row): the operand layout of the bf16x3 mode, accepted by
 * msq_gemm dtype 6 (fp32 out) / 7 (split-bf16 out [M, hi(N)|lo(N)]), msq_attention dtype 2 (qkv rows
 * [hi(3*heads*64) | lo(..)], ctx rows [hi(heads*64) | lo(..)]) and produced by msq_layernorm dtype 2. */
int msq_f32_to_bf16_split(const float* src_dev, void* dst_dev, int64_t rows, int32_t K, void* stream);

#ifdef __cplusplus
}
#endif
#endif
