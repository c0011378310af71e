// Fine-tuning path of the inner encoder (BASELINE config 4): training-mode forward that records the
// activations the backward pass needs, the backward pass itself, and the optimizer step.
//
// Reference: autograd through LXRTModel.forward (models/CLIP/src/lxrt/modeling.py:1513-1598) = BertEmbeddings (342-370),
// LXRTEncoder CLIP branch (838-1107: tower, visn_fc + LayerNorm, concat, BertLayer x L), BertLayer (373-507); the CLIP
// VisualTransformer (models/CLIP/clip/model.py:190-305); optimizer trainers/train.py:172-190 (HF AdamW, no_decay =
// bias / LayerNorm.weight), 353-363 (clip_grad_norm_ then step).  Dropout is NOT applied (p = 0 semantics, the
// configuration SURVEY §8(d) cfg4 prescribes for parity runs).
//
// Design.  Every contraction of the backward pass is a call of the forward GEMM kernels (gemm_tc.cu on
// tcgen05, gemm_simt.cu in the fp32 parity mode), which compute C = A W^T with K-major operands:
//   dgrad: dX = dY (W^T)^T with W^T a packed [K,N] copy of every weight (refreshed after each optimizer step);
//   wgrad: dW += (dY^T)(X^T)^T with M-contiguous, zero-padded transposes written by transpose_pad; the accumulation
//          into the caller's flat gradient buffer happens in the GEMM epilogue (residual operand == output).
// Gradients live in ONE flat fp32 buffer owned by the caller (a torch tensor): data-parallel fine-tuning all-reduces
// that buffer with a single NCCL call and hands it to msq_adamw_step.
#include "train_common.cuh"

namespace msq {

void train_state_free(TrainState* ts) { train_state_free_impl(ts); }

// ---- parameter table -------------------------------------------------------------------------------
static int add_slot(msq_model* m, TrainState* ts, const std::string& name) {
  auto it = m->raw.find(name);
  MSQ_REQUIRE(it != m->raw.end(), "train: weight %s is not registered", name.c_str());
  ParamSlot s;
  s.name = name; s.off = ts->total; s.numel = it->second.second; s.master = it->second.first;
  s.decay = name.find("bias") == std::string::npos && name.find("LayerNorm.weight") == std::string::npos;   // train.py:172
  ts->index[name] = ts->slots.size();
  ts->slots.push_back(s);
  ts->total += round_up(s.numel, 64);   // 256-byte aligned slots (TMA / float4 access to the gradient blocks)
  return MSQ_OK;
}

template <typename T> static int make_wT(TrainState* ts, const Lin& l, void** out) {
  void* p = nullptr;
  MSQ_CUDA(cudaMalloc(&p, (size_t)l.N * l.K * sizeof(T)));
  ts->owned.push_back(p);
  *out = p;
  return MSQ_OK;
}
template <typename T> static int fill_wT(const Lin& l, void* wT, cudaStream_t st) {
  // w32 [N, ld] (K columns used) -> [K, N]
  return transpose_pad<float, T>(l.w32, l.N, l.K, l.ld, l.N, (T*)wT, ACT_NONE, st);
}
template <typename T> static int refresh_wT(msq_model* m, TrainState* ts, cudaStream_t st) {
  for (size_t l = 0; l < m->bert.size(); ++l) {
    const Lin* ls[4] = {&m->bert[l].qkv, &m->bert[l].out, &m->bert[l].up, &m->bert[l].down};
    for (int i = 0; i < 4; ++i) MSQ_TRY(fill_wT<T>(*ls[i], ts->bertT[l][i], st));
  }
  for (size_t l = 0; l < m->vit.size(); ++l) {
    const Lin* ls[4] = {&m->vit[l].qkv, &m->vit[l].out, &m->vit[l].fc, &m->vit[l].proj};
    for (int i = 0; i < 4; ++i) MSQ_TRY(fill_wT<T>(*ls[i], ts->vitT[l][i], st));
  }
  if (ts->visnT) MSQ_TRY(fill_wT<T>(m->visn_fc, ts->visnT, st));
  if (ts->rn) MSQ_TRY(rn_train_refresh<T>(m, st));
  return MSQ_OK;
}

template <typename T> static int build_train_state(msq_model* m, cudaStream_t st) {
  const msq_config& c = m->cfg;
  MSQ_REQUIRE(m->packed && m->has_bert, "train: model not packed / no inner encoder weights");
  TrainState* ts = new TrainState();
  m->train = ts;
  const std::string P = m->prefix_inner;
  for (const char* e : {"embeddings.word_embeddings.weight", "embeddings.position_embeddings.weight",
                        "embeddings.token_type_embeddings.weight", "embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias"})
    MSQ_TRY(add_slot(m, ts, P + e));
  for (int l = 0; l < c.layers; ++l) {
    const std::string b = P + "encoder.layer." + std::to_string(l) + ".";
    // q/k/v are adjacent so that the fused [3H,H] gradient block is contiguous
    for (const char* e : {"attention.self.query.weight", "attention.self.key.weight", "attention.self.value.weight",
                          "attention.self.query.bias", "attention.self.key.bias", "attention.self.value.bias",
                          "attention.output.dense.weight", "attention.output.dense.bias", "attention.output.LayerNorm.weight",
                          "attention.output.LayerNorm.bias", "intermediate.dense.weight", "intermediate.dense.bias",
                          "output.dense.weight", "output.dense.bias", "output.LayerNorm.weight", "output.LayerNorm.bias"})
      MSQ_TRY(add_slot(m, ts, b + e));
  }
  if (m->has_vit && c.rn_width) {
    for (const std::string& n : rn_param_names(m)) MSQ_TRY(add_slot(m, ts, n));
  } else if (m->has_vit) {
    const std::string v = P + "encoder.visual_model.visual.";
    for (const char* e : {"conv1.weight", "class_embedding", "positional_embedding", "ln_pre.weight", "ln_pre.bias", "ln_post.weight",
                          "ln_post.bias"})
      MSQ_TRY(add_slot(m, ts, v + e));
    for (int l = 0; l < c.vit_layers; ++l) {
      const std::string b = v + "transformer.resblocks." + std::to_string(l) + ".";
      for (const char* e : {"attn.in_proj_weight", "attn.in_proj_bias", "attn.out_proj.weight", "attn.out_proj.bias", "ln_1.weight",
                            "ln_1.bias", "mlp.c_fc.weight", "mlp.c_fc.bias", "mlp.c_proj.weight", "mlp.c_proj.bias", "ln_2.weight",
                            "ln_2.bias"})
        MSQ_TRY(add_slot(m, ts, b + e));
    }
    for (const char* e : {"encoder.visn_fc.visn_fc.weight", "encoder.visn_fc.visn_fc.bias", "encoder.visn_fc.visn_layer_norm.weight",
                          "encoder.visn_fc.visn_layer_norm.bias"})
      MSQ_TRY(add_slot(m, ts, P + e));
  }
  // the fused-QKV gradient block relies on adjacency without padding
  MSQ_REQUIRE(((int64_t)c.hidden * c.hidden) % 64 == 0 && c.hidden % 64 == 0, "train: hidden size must be a multiple of 64");
  ts->bertT.resize(m->bert.size());
  for (size_t l = 0; l < m->bert.size(); ++l) {
    const Lin* ls[4] = {&m->bert[l].qkv, &m->bert[l].out, &m->bert[l].up, &m->bert[l].down};
    for (int i = 0; i < 4; ++i) MSQ_TRY(make_wT<T>(ts, *ls[i], &ts->bertT[l][i]));
  }
  ts->vitT.resize(m->vit.size());
  for (size_t l = 0; l < m->vit.size(); ++l) {
    const Lin* ls[4] = {&m->vit[l].qkv, &m->vit[l].out, &m->vit[l].fc, &m->vit[l].proj};
    for (int i = 0; i < 4; ++i) MSQ_TRY(make_wT<T>(ts, *ls[i], &ts->vitT[l][i]));
  }
  if (m->has_vit) MSQ_TRY(make_wT<T>(ts, m->visn_fc, &ts->visnT));
  if (m->has_vit && c.rn_width) MSQ_TRY(rn_train_build<T>(m, st));
  MSQ_TRY(refresh_wT<T>(m, ts, st));
  if (m->has_heads) {
    for (const std::string& n : heads_param_names(m)) MSQ_TRY(add_slot(m, ts, n));
    MSQ_TRY(heads_train_setup(m, true, st));
  }
  return MSQ_OK;
}
static int ensure_train(msq_model* m, cudaStream_t st) {
  MSQ_REQUIRE(m, "null model");
  if (m->train) return MSQ_OK;
  MSQ_REQUIRE(m->cfg.precise != 2, "fine-tuning runs in precise 0 (bf16) or 1 (fp32); the bf16x3 mode is an evaluation mode");
  MSQ_REQUIRE(!(m->cfg.reserved & 1), "fine-tuning through a tanh-pooled CLS (cls_pooler, HuggingFace inner model) is not built: evaluation only");
  const int rc = m->cfg.precise ? build_train_state<float>(m, st) : build_train_state<bf16>(m, st);
  if (rc != MSQ_OK && m->train) { train_state_free(m->train); m->train = nullptr; }
  return rc;
}

// mark the slots [first, last] (by name) final: one event on the compute stream, recorded in completion order
int mark_ready(TrainState* ts, const std::string& first, const std::string& last, cudaStream_t st) {
  auto a = ts->index.find(first), b = ts->index.find(last);
  if (a == ts->index.end() || b == ts->index.end()) return MSQ_OK;   // group absent from this model
  if (ts->ready_used == ts->ready_ev.size()) {
    cudaEvent_t e;
    MSQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    ts->ready_ev.push_back(e);
    ts->ready_rng.push_back({0, 0});
  }
  const ParamSlot &sa = ts->slots[a->second], &sb = ts->slots[b->second];
  ts->ready_rng[ts->ready_used] = {sa.off, sb.off + sb.numel};
  MSQ_CUDA(cudaEventRecord(ts->ready_ev[ts->ready_used], st));
  ++ts->ready_used;
  return MSQ_OK;
}

// ---- training forward ------------------------------------------------------------------------------
template <typename T>
static int forward_train(msq_model* m, const int64_t* ids, const int64_t* tt, const int64_t* mask, int64_t R, int Lt, const float* images,
                         int64_t n_img, const int32_t* img_index, float* lang, float* visn, cudaStream_t st) {
  TrainState* ts = m->train;
  const msq_config& c = m->cfg;
  const int H = c.hidden, I = c.inter;
  const bool mm = m->has_vit && images != nullptr;
  MSQ_REQUIRE(!mm || img_index, "train: images without img_index");
  MSQ_REQUIRE(Lt <= c.max_pos && R > 0, "train: bad R / Lt");
  const int g = mm ? c.vit_res / c.vit_patch : 0, g2 = g * g;
  const int Lv = mm ? 1 + 2 * g2 : 0, Lj = Lt + Lv, Wd = c.vit_width, Kc = 3 * c.vit_patch * c.vit_patch;
  const int64_t Mj = R * Lj, Mv = R * Lv;
  const size_t nb = m->bert.size(), nvl = mm ? m->vit.size() : 0;
  ts->have_fwd = false;
  ts->bt.assign(nb, BertTape{});
  ts->vt.assign(nvl, VitTape{});
  float *xf = nullptr, *x1f = nullptr;
  T* apatch = nullptr;
  for (int pass = 0; pass < 2; ++pass) {
    Planner p{&ts->tape, pass == 0};
    if (pass == 1) ts->tape.reset();
    ts->ids = p.take<int64_t>((size_t)R * Lt);
    ts->tt = p.take<int64_t>((size_t)R * Lt);
    ts->mask_add = p.take<float>((size_t)R * Lt);
    if (mm) {
      ts->img_index = p.take<int32_t>((size_t)R * 2);
      ts->patch = p.take<float>((size_t)n_img * g2 * Wd);
      for (auto& v : ts->vt) {
        v.x = p.take<float>((size_t)Mv * Wd); v.y1 = p.take<T>((size_t)Mv * Wd); v.qkv = p.take<T>((size_t)Mv * 3 * Wd);
        v.ctx = p.take<T>((size_t)Mv * Wd); v.x1 = p.take<float>((size_t)Mv * Wd); v.y2 = p.take<T>((size_t)Mv * Wd);
        v.u = p.take<T>((size_t)Mv * 4 * Wd); v.hb = p.take<T>((size_t)Mv * 4 * Wd);
        v.lse = p.take<float>((size_t)R * (Wd / 64) * Lv); v.have_lse = false;
      }
      ts->vx_last = p.take<float>((size_t)Mv * Wd);
      ts->y_post = p.take<T>((size_t)Mv * Wd);
      ts->visn_pre = p.take<float>((size_t)Mv * H);
    }
    for (auto& b : ts->bt) {
      b.x = p.take<T>((size_t)Mj * H); b.qkv = p.take<T>((size_t)Mj * 3 * H); b.ctx = p.take<T>((size_t)Mj * H);
      b.s1 = p.take<float>((size_t)Mj * H); b.x1 = p.take<T>((size_t)Mj * H); b.u = p.take<T>((size_t)Mj * I);
      b.s2 = p.take<float>((size_t)Mj * H); b.hb = p.take<T>((size_t)Mj * I);
      b.lse = p.take<float>((size_t)R * c.heads * Lj); b.have_lse = false;
    }
    ts->x_last = p.take<T>((size_t)Mj * H);
    ts->x_out = p.take<float>((size_t)Mj * H);
    if (pass == 0) MSQ_TRY(ts->tape.reserve(p.need + 4096, st));
  }
  for (int pass = 0; pass < 2; ++pass) {   // transient buffers
    Planner p{&m->ws, pass == 0};
    if (pass == 1) m->ws.reset();
    xf = p.take<float>((size_t)Mj * H);
    x1f = p.take<float>((size_t)Mj * H);
    if (mm) {
      apatch = p.take<T>((size_t)min(n_img, TRAIN_IMG_CHUNK) * g2 * Kc);
    }
    if (pass == 0) MSQ_TRY(m->ws.reserve(p.need + 4096, st));
  }
  MSQ_CUDA(cudaMemcpyAsync(ts->ids, ids, (size_t)R * Lt * 8, cudaMemcpyDeviceToDevice, st));
  MSQ_CUDA(cudaMemcpyAsync(ts->tt, tt, (size_t)R * Lt * 8, cudaMemcpyDeviceToDevice, st));
  if (mm) MSQ_CUDA(cudaMemcpyAsync(ts->img_index, img_index, (size_t)R * 2 * 4, cudaMemcpyDeviceToDevice, st));
  MSQ_TRY(mask_add_from_int(mask, R * Lt, ts->mask_add, st));

  // dropout: this forward's masks are keyed by a fresh step counter (the backward pass regenerates them from it)
  ts->drop.step = ts->drop_next_step++;
  const DropCfg& dc = ts->drop;
  const bool drop_h = dc.p_hidden > 0.f;
  // layer-0 input: embeddings (text rows) and visn_fc(LN) of the tower output (visual rows) -> xf (fp32) + X[0] (T)
  T* x0 = nb ? (T*)ts->bt[0].x : (T*)ts->x_last;
  MSQ_TRY(embed_ln<T>(ids, tt, R, Lt, Lj, H, m->word, m->pos, m->type, m->emb_ln.g, m->emb_ln.b, 1e-12f, xf, x0, st));
  MSQ_TRY(dropout_rows<T>(xf, x0, R, Lj, 0, Lt, H, make_drop(dc, DROP_E, 0, dc.p_hidden), st));
  if (mm) {
    if (c.rn_width) {
      MSQ_TRY(rn_forward_train<T>(m, images, n_img, img_index, R, st));   // -> ts->y_post = cat(o, o) + position / type embeddings
    } else {
    // patch embedding per UNIQUE image
    for (int64_t i0 = 0; i0 < n_img; i0 += TRAIN_IMG_CHUNK) {
      const int64_t n = min(TRAIN_IMG_CHUNK, n_img - i0);
      MSQ_TRY(im2col<T>(images + i0 * 3 * c.vit_res * c.vit_res, n, c.vit_res, c.vit_patch, apatch, st));
      MSQ_TRY((gemm_nt<T, float>(m, apatch, Kc, wptr<T>(m->conv1), m->conv1.ld, nullptr, nullptr, 0, ts->patch + i0 * g2 * Wd, Wd, n * g2, Wd,
                                 Kc, ACT_NONE, st)));
    }
    float* x = nvl ? ts->vt[0].x : ts->vx_last;
    MSQ_TRY(vit_assemble(ts->patch, img_index, R, 2, g2, Wd, m->vit_cls, m->vit_pos, m->ln_pre.g, m->ln_pre.b, 1e-5f, x, st));
    const int heads = Wd / 64;
    for (size_t l = 0; l < nvl; ++l) {
      VitLayerW& L = m->vit[l];
      VitTape& t = ts->vt[l];
      float* xn = l + 1 < nvl ? ts->vt[l + 1].x : ts->vx_last;
      MSQ_TRY(layernorm<T>(t.x, Mv, Wd, L.ln1.g, L.ln1.b, 1e-5f, nullptr, (T*)t.y1, 0, 0, 0, st));
      MSQ_TRY((gemm_nt<T, T>(m, (const T*)t.y1, Wd, wptr<T>(L.qkv), L.qkv.ld, L.qkv.b, nullptr, 0, (T*)t.qkv, 3 * Wd, Mv, 3 * Wd, Wd, ACT_NONE, st)));
      MSQ_TRY(attention<T>((const T*)t.qkv, R, Lv, heads, 64, 0.125f, nullptr, 0, 0, (T*)t.ctx, st, Drop(), t.lse, &t.have_lse));
      MSQ_TRY((gemm_nt<T, float>(m, (const T*)t.ctx, Wd, wptr<T>(L.out), L.out.ld, L.out.b, t.x, Wd, t.x1, Wd, Mv, Wd, Wd, ACT_NONE, st)));
      MSQ_TRY(layernorm<T>(t.x1, Mv, Wd, L.ln2.g, L.ln2.b, 1e-5f, nullptr, (T*)t.y2, 0, 0, 0, st));
      MSQ_TRY(gemm_nt_dualact<T>(m, (const T*)t.y2, Wd, wptr<T>(L.fc), L.fc.ld, L.fc.b, (T*)t.u, (T*)t.hb, 4 * Wd, Mv, 4 * Wd, Wd, ACT_QUICK_GELU, st));
      MSQ_TRY((gemm_nt<T, float>(m, (const T*)t.hb, 4 * Wd, wptr<T>(L.proj), L.proj.ld, L.proj.b, t.x1, Wd, xn, Wd, Mv, Wd, 4 * Wd, ACT_NONE, st)));
    }
    MSQ_TRY(layernorm<T>(ts->vx_last, Mv, Wd, m->ln_post.g, m->ln_post.b, 1e-5f, nullptr, (T*)ts->y_post, 0, 0, 0, st));
    }
    MSQ_TRY((gemm_nt<T, float>(m, (const T*)ts->y_post, Wd, wptr<T>(m->visn_fc), m->visn_fc.ld, m->visn_fc.b, nullptr, 0, ts->visn_pre, H, Mv,
                               H, Wd, ACT_NONE, st)));
    MSQ_TRY(layernorm<T>(ts->visn_pre, Mv, H, m->visn_ln.g, m->visn_ln.b, 1e-12f, xf, x0, Lv, Lj, Lt, st));
    MSQ_TRY(dropout_rows<T>(xf, x0, R, Lj, Lt, Lv, H, make_drop(dc, DROP_V, 0, dc.p_hidden), st));
  }
  for (size_t l = 0; l < nb; ++l) {
    BertLayerW& L = m->bert[l];
    BertTape& t = ts->bt[l];
    T* xn = l + 1 < nb ? (T*)ts->bt[l + 1].x : (T*)ts->x_last;
    MSQ_TRY((gemm_nt<T, T>(m, (const T*)t.x, H, wptr<T>(L.qkv), L.qkv.ld, L.qkv.b, nullptr, 0, (T*)t.qkv, 3 * H, Mj, 3 * H, H, ACT_NONE, st)));
    MSQ_TRY(attention<T>((const T*)t.qkv, R, Lj, c.heads, 64, 0.125f, ts->mask_add, Lt, Lt, (T*)t.ctx, st,
                         make_drop(dc, DROP_A, (int)l, dc.p_attn), t.lse, &t.have_lse));
    if (drop_h && gemm_nt_on_tc<T>(m, H, L.out.ld, H, H, H)) {   // s1 = dropout(dense(ctx) + b) + x, dropout inside the GEMM epilogue
      MSQ_TRY((gemm_nt<T, float>(m, (const T*)t.ctx, H, wptr<T>(L.out), L.out.ld, L.out.b, xf, H, t.s1, H, Mj, H, H, ACT_NONE, st,
                                 make_drop(dc, DROP_O, (int)l, dc.p_hidden))));
    } else if (drop_h) {
      MSQ_TRY((gemm_nt<T, float>(m, (const T*)t.ctx, H, wptr<T>(L.out), L.out.ld, L.out.b, nullptr, 0, t.s1, H, Mj, H, H, ACT_NONE, st)));
      MSQ_TRY(dropout_add(t.s1, xf, Mj * H, make_drop(dc, DROP_O, (int)l, dc.p_hidden), st));
    } else {
      MSQ_TRY((gemm_nt<T, float>(m, (const T*)t.ctx, H, wptr<T>(L.out), L.out.ld, L.out.b, xf, H, t.s1, H, Mj, H, H, ACT_NONE, st)));
    }
    MSQ_TRY(layernorm<T>(t.s1, Mj, H, L.ln1.g, L.ln1.b, 1e-12f, x1f, (T*)t.x1, 0, 0, 0, st));
    MSQ_TRY(gemm_nt_dualact<T>(m, (const T*)t.x1, H, wptr<T>(L.up), L.up.ld, L.up.b, (T*)t.u, (T*)t.hb, I, Mj, I, H, ACT_GELU_ERF, st));
    if (drop_h && gemm_nt_on_tc<T>(m, I, L.down.ld, H, H, I)) {   // s2 = dropout(dense(h) + b) + x1
      MSQ_TRY((gemm_nt<T, float>(m, (const T*)t.hb, I, wptr<T>(L.down), L.down.ld, L.down.b, x1f, H, t.s2, H, Mj, H, I, ACT_NONE, st,
                                 make_drop(dc, DROP_F, (int)l, dc.p_hidden))));
    } else if (drop_h) {
      MSQ_TRY((gemm_nt<T, float>(m, (const T*)t.hb, I, wptr<T>(L.down), L.down.ld, L.down.b, nullptr, 0, t.s2, H, Mj, H, I, ACT_NONE, st)));
      MSQ_TRY(dropout_add(t.s2, x1f, Mj * H, make_drop(dc, DROP_F, (int)l, dc.p_hidden), st));
    } else {
      MSQ_TRY((gemm_nt<T, float>(m, (const T*)t.hb, I, wptr<T>(L.down), L.down.ld, L.down.b, x1f, H, t.s2, H, Mj, H, I, ACT_NONE, st)));
    }
    MSQ_TRY(layernorm<T>(t.s2, Mj, H, L.ln2.g, L.ln2.b, 1e-12f, xf, xn, 0, 0, 0, st));
  }
  MSQ_CUDA(cudaMemcpyAsync(ts->x_out, xf, (size_t)Mj * H * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (lang) MSQ_TRY((gather_rows<float, float>(xf, R * Lt, H, Lt, Lj, 0, lang, st)));
  if (visn && mm) MSQ_TRY((gather_rows<float, float>(xf, R * Lv, H, Lv, Lj, Lt, visn, st)));
  ts->R = R; ts->n_img = n_img; ts->Lt = Lt; ts->Lv = Lv; ts->Lj = Lj; ts->mm = mm; ts->images = images;
  ts->have_fwd = true;
  return MSQ_OK;
}

// ---- backward --------------------------------------------------------------------------------------
template <typename T>
static int backward_train(msq_model* m, const float* d_lang, const float* d_visn, float* grads, cudaStream_t st) {
  TrainState* ts = m->train;
  MSQ_REQUIRE(ts && ts->have_fwd, "msq_inner_backward: no recorded training forward");
  const msq_config& c = m->cfg;
  const int H = c.hidden, I = c.inter, Lt = ts->Lt, Lv = ts->Lv, Lj = ts->Lj;
  const bool mm = ts->mm;
  const int64_t R = ts->R, Mj = R * Lj, Mv = R * Lv, n_img = ts->n_img;
  const int g = mm ? c.vit_res / c.vit_patch : 0, g2 = g * g, Wd = c.vit_width, Kc = 3 * c.vit_patch * c.vit_patch;
  const size_t nb = m->bert.size(), nvl = ts->vt.size();
  const std::string P = m->prefix_inner;
  int err = MSQ_OK;
  auto G = [&](const std::string& name) -> float* {
    auto it = ts->index.find(name);
    if (it == ts->index.end()) { set_error("train: no gradient slot for %s", name.c_str()); err = MSQ_ERR_STATE; return nullptr; }
    return grads + ts->slots[it->second].off;
  };
  const DropCfg& dc = ts->drop;   // the masks of the recorded forward
  const int64_t MpJ = wgrad_rows(Mj), MpV = wgrad_rows(max(Mv, (int64_t)1));
  const int64_t nchunk = mm ? min(n_img, TRAIN_IMG_CHUNK) * g2 : 0, Np = round_up(max(nchunk, (int64_t)1), 64);
  BwdBufs b{};
  for (int pass = 0; pass < 2; ++pass) {
    Planner p{&m->ws, pass == 0};
    if (pass == 1) m->ws.reset();
    const size_t act = (size_t)max(Mj * H, Mv * (int64_t)Wd);
    b.gA = p.take<float>(act);
    b.gB = p.take<float>(max(act, (size_t)Mv * H));
    b.gT = p.take<T>(max(act, (size_t)Mv * H));
    b.gC = p.take<T>(act);
    b.gH = p.take<T>((size_t)max(Mj * I, Mv * 4 * (int64_t)Wd));
    b.gQ = p.take<T>(3 * act);
    b.GT = p.take<T>((size_t)max(max((int64_t)max(3 * H, I) * MpJ, (int64_t)4 * Wd * MpV), (int64_t)Wd * Np));
    b.XT = p.take<T>((size_t)max(max((int64_t)max(H, I) * MpJ, (int64_t)4 * Wd * MpV), (int64_t)Kc * Np));
    b.scr_floats = max(ln_bwd_scratch_floats(max(H, Wd)), colsum_scratch_floats(max(max(3 * H, I), 4 * Wd)));
    b.ln_scr = p.take<float>(b.scr_floats);
    b.at_scr = p.take<float>(max(attention_bwd_scratch_floats(R, Lj, c.heads), attention_bwd_scratch_floats(R, max(Lv, 1), max(Wd / 64, 1))));
    if (mm) {
      b.dpatch = p.take<float>((size_t)n_img * g2 * Wd);
      b.apatch = p.take<T>((size_t)nchunk * Kc);
    }
    if (sizeof(T) == 2) b.part = p.take<float>((size_t)SPLITK_MAX * max((int64_t)max(3 * H, I) * H, (int64_t)4 * Wd * Wd));
    if (pass == 0) MSQ_TRY(m->ws.reserve(p.need + 4096, st));
  }
  if (sizeof(T) == 2) b.sk = &ts->splitk;
  // gradient of the final joint stream
  MSQ_CUDA(cudaMemsetAsync(b.gA, 0, (size_t)Mj * H * sizeof(float), st));
  if (d_lang) MSQ_TRY(scatter_rows(d_lang, R * Lt, H, Lt, Lj, 0, b.gA, st));
  if (d_visn && mm) MSQ_TRY(scatter_rows(d_visn, R * Lv, H, Lv, Lj, Lt, b.gA, st));
  if (ts->ml_live && mm && d_lang == ts->ht.d_lang) {   // image pairwise objective of this step: gradient of visn[:, 0]
    MSQ_TRY(add_first_visual(ts->ht.ml_dv, R, Lt, Lj, H, b.gA, st));
    ts->ml_live = false;
  }

  for (size_t li = nb; li-- > 0;) {
    BertLayerW& L = m->bert[li];
    BertTape& t = ts->bt[li];
    auto& WT = ts->bertT[li];
    const std::string bn = P + "encoder.layer." + std::to_string(li) + ".";
    float *dWqkv = G(bn + "attention.self.query.weight"), *dbqkv = G(bn + "attention.self.query.bias");
    float *dWo = G(bn + "attention.output.dense.weight"), *dbo = G(bn + "attention.output.dense.bias");
    float *dg1 = G(bn + "attention.output.LayerNorm.weight"), *db1 = G(bn + "attention.output.LayerNorm.bias");
    float *dWu = G(bn + "intermediate.dense.weight"), *dbu = G(bn + "intermediate.dense.bias");
    float *dWd = G(bn + "output.dense.weight"), *dbd = G(bn + "output.dense.bias");
    float *dg2 = G(bn + "output.LayerNorm.weight"), *db2 = G(bn + "output.LayerNorm.bias");
    if (err) return err;
    // output.LayerNorm -> ds2 (gB fp32, gT operand copy)
    // under dropout the dense output's gradient is ds2 * mask (operand copy gT, written by the same kernel); the residual
    // branch keeps ds2 (gB)
    MSQ_TRY(ln_bwd<T>(b.gA, t.s2, nullptr, Mj, H, L.ln2.g, 1e-12f, b.gB, (T*)b.gT, dg2, db2, b.ln_scr, 0, 0, 0, st,
                      make_drop(dc, DROP_F, (int)li, dc.p_hidden)));
    MSQ_TRY(wgrad<T>(m, (const T*)b.gT, H, H, (const T*)t.hb, I, I, ACT_NONE, Mj, dWd, dbd, b, st));
    MSQ_TRY((dgrad<T, T>(m, (const T*)b.gT, H, WT[3], I, nullptr, (T*)b.gH, Mj, st)));
    MSQ_TRY(act_bwd<T>((const T*)b.gH, (const T*)t.u, Mj * I, ACT_GELU_ERF, (T*)b.gH, st));
    MSQ_TRY(wgrad<T>(m, (const T*)b.gH, I, I, (const T*)t.x1, H, H, ACT_NONE, Mj, dWu, dbu, b, st));
    MSQ_TRY((dgrad<T, float>(m, (const T*)b.gH, I, WT[2], H, b.gB, b.gA, Mj, st)));            // dX1 = du Wup + ds2
    // attention.output.LayerNorm -> ds1
    MSQ_TRY(ln_bwd<T>(b.gA, t.s1, nullptr, Mj, H, L.ln1.g, 1e-12f, b.gB, (T*)b.gT, dg1, db1, b.ln_scr, 0, 0, 0, st,
                      make_drop(dc, DROP_O, (int)li, dc.p_hidden)));
    MSQ_TRY(wgrad<T>(m, (const T*)b.gT, H, H, (const T*)t.ctx, H, H, ACT_NONE, Mj, dWo, dbo, b, st));
    MSQ_TRY((dgrad<T, T>(m, (const T*)b.gT, H, WT[1], H, nullptr, (T*)b.gC, Mj, st)));
    MSQ_TRY(attn_bwd<T>((const T*)t.qkv, (const T*)t.ctx, (const T*)b.gC, R, Lj, c.heads, ts->mask_add, Lt, (T*)b.gQ, b.at_scr, st,
                        make_drop(dc, DROP_A, (int)li, dc.p_attn), t.have_lse ? t.lse : nullptr));
    MSQ_TRY(wgrad<T>(m, (const T*)b.gQ, 3 * H, 3 * H, (const T*)t.x, H, H, ACT_NONE, Mj, dWqkv, dbqkv, b, st));
    MSQ_TRY((dgrad<T, float>(m, (const T*)b.gQ, 3 * H, WT[0], H, b.gB, b.gA, Mj, st)));        // dX0 = dqkv Wqkv + ds1
    MSQ_TRY(mark_ready(ts, bn + "attention.self.query.weight", bn + "output.LayerNorm.bias", st));
  }
  // gA = gradient of the layer-0 input (LayerNorm outputs of the embeddings / of visn_fc)
  {
    float *dword = G(P + "embeddings.word_embeddings.weight"), *dpos = G(P + "embeddings.position_embeddings.weight"),
          *dtyp = G(P + "embeddings.token_type_embeddings.weight"), *dg = G(P + "embeddings.LayerNorm.weight"),
          *db = G(P + "embeddings.LayerNorm.bias");
    if (err) return err;
    // gA = d(joint layer-0 input): undo the input dropouts (text rows: E, visual rows: V) before the LayerNorm backwards
    MSQ_TRY(dropout_rows<float>(b.gA, nullptr, R, Lj, 0, Lt, H, make_drop(dc, DROP_E, 0, dc.p_hidden), st));
    if (mm) MSQ_TRY(dropout_rows<float>(b.gA, nullptr, R, Lj, Lt, Lv, H, make_drop(dc, DROP_V, 0, dc.p_hidden), st));
    MSQ_TRY(embed_ln_bwd(b.gA, ts->ids, ts->tt, R, Lt, Lj, H, m->word, m->pos, m->type, m->emb_ln.g, 1e-12f, dword, dpos, dtyp, dg, db, b.ln_scr,
                         c.vit_width != 0 ? 7 : 1, st));   // padding_idx=0 tables: LXRT all three, text-only BertModel the word table
    MSQ_TRY(mark_ready(ts, P + "embeddings.word_embeddings.weight", P + "embeddings.LayerNorm.bias", st));
  }
  if (!mm) return MSQ_OK;

  const std::string v = P + "encoder.visual_model.visual.";
  {
    const bool rn = c.rn_width != 0;
    float *dWv = G(P + "encoder.visn_fc.visn_fc.weight"), *dbv = G(P + "encoder.visn_fc.visn_fc.bias"),
          *dgv = G(P + "encoder.visn_fc.visn_layer_norm.weight"), *dbl = G(P + "encoder.visn_fc.visn_layer_norm.bias"),
          *dgp = rn ? nullptr : G(v + "ln_post.weight"), *dbp = rn ? nullptr : G(v + "ln_post.bias");
    if (err) return err;
    // visn_layer_norm backward on the visual rows of the joint gradient -> d(visn_fc out) [Mv,H] (gB) + operand copy (gT)
    MSQ_TRY(ln_bwd<T>(b.gA, ts->visn_pre, nullptr, Mv, H, m->visn_ln.g, 1e-12f, b.gB, (T*)b.gT, dgv, dbl, b.ln_scr, Lv, Lj, Lt, st));
    MSQ_TRY(wgrad<T>(m, (const T*)b.gT, H, H, (const T*)ts->y_post, Wd, Wd, ACT_NONE, Mv, dWv, dbv, b, st));
    // d(ln_post out) fp32 -> gA is free now: reuse it as [Mv, Wd]
    MSQ_TRY((dgrad<T, float>(m, (const T*)b.gT, H, ts->visnT, Wd, nullptr, b.gA, Mv, st)));
    MSQ_TRY(mark_ready(ts, P + "encoder.visn_fc.visn_fc.weight", P + "encoder.visn_fc.visn_layer_norm.bias", st));
    if (rn) return rn_backward_train<T>(m, b.gA, grads, st);   // ModifiedResNet tower: attention pool, blocks, stem (train_rn.cu)
    MSQ_TRY(ln_bwd<T>(b.gA, ts->vx_last, nullptr, Mv, Wd, m->ln_post.g, 1e-5f, b.gB, (T*)b.gT, dgp, dbp, b.ln_scr, 0, 0, 0, st));
  }
  // ViT blocks: stream gradient dx in gB (fp32) + gT (operand copy)
  const int vheads = Wd / 64;
  for (size_t li = nvl; li-- > 0;) {
    VitLayerW& L = m->vit[li];
    VitTape& t = ts->vt[li];
    auto& WT = ts->vitT[li];
    const std::string bn = v + "transformer.resblocks." + std::to_string(li) + ".";
    float *dWqkv = G(bn + "attn.in_proj_weight"), *dbqkv = G(bn + "attn.in_proj_bias");
    float *dWo = G(bn + "attn.out_proj.weight"), *dbo = G(bn + "attn.out_proj.bias");
    float *dg1 = G(bn + "ln_1.weight"), *db1 = G(bn + "ln_1.bias"), *dg2 = G(bn + "ln_2.weight"), *db2 = G(bn + "ln_2.bias");
    float *dWf = G(bn + "mlp.c_fc.weight"), *dbf = G(bn + "mlp.c_fc.bias"), *dWp = G(bn + "mlp.c_proj.weight"), *dbp = G(bn + "mlp.c_proj.bias");
    if (err) return err;
    MSQ_TRY(wgrad<T>(m, (const T*)b.gT, Wd, Wd, (const T*)t.hb, 4 * Wd, 4 * Wd, ACT_NONE, Mv, dWp, dbp, b, st));
    MSQ_TRY((dgrad<T, T>(m, (const T*)b.gT, Wd, WT[3], 4 * Wd, nullptr, (T*)b.gH, Mv, st)));
    MSQ_TRY(act_bwd<T>((const T*)b.gH, (const T*)t.u, Mv * 4 * Wd, ACT_QUICK_GELU, (T*)b.gH, st));
    MSQ_TRY(wgrad<T>(m, (const T*)b.gH, 4 * Wd, 4 * Wd, (const T*)t.y2, Wd, Wd, ACT_NONE, Mv, dWf, dbf, b, st));
    MSQ_TRY((dgrad<T, float>(m, (const T*)b.gH, 4 * Wd, WT[2], Wd, nullptr, b.gA, Mv, st)));      // d(ln_2 out)
    MSQ_TRY(ln_bwd<T>(b.gA, t.x1, b.gB, Mv, Wd, L.ln2.g, 1e-5f, b.gB, (T*)b.gT, dg2, db2, b.ln_scr, 0, 0, 0, st));   // dx1 = dx + LN'
    MSQ_TRY(wgrad<T>(m, (const T*)b.gT, Wd, Wd, (const T*)t.ctx, Wd, Wd, ACT_NONE, Mv, dWo, dbo, b, st));
    MSQ_TRY((dgrad<T, T>(m, (const T*)b.gT, Wd, WT[1], Wd, nullptr, (T*)b.gC, Mv, st)));
    MSQ_TRY(attn_bwd<T>((const T*)t.qkv, (const T*)t.ctx, (const T*)b.gC, R, Lv, vheads, nullptr, 0, (T*)b.gQ, b.at_scr, st, Drop(),
                        t.have_lse ? t.lse : nullptr));
    MSQ_TRY(wgrad<T>(m, (const T*)b.gQ, 3 * Wd, 3 * Wd, (const T*)t.y1, Wd, Wd, ACT_NONE, Mv, dWqkv, dbqkv, b, st));
    MSQ_TRY((dgrad<T, float>(m, (const T*)b.gQ, 3 * Wd, WT[0], Wd, nullptr, b.gA, Mv, st)));      // d(ln_1 out)
    MSQ_TRY(ln_bwd<T>(b.gA, t.x, b.gB, Mv, Wd, L.ln1.g, 1e-5f, b.gB, (T*)b.gT, dg1, db1, b.ln_scr, 0, 0, 0, st));    // dx = dx1 + LN'
    MSQ_TRY(mark_ready(ts, bn + "attn.in_proj_weight", bn + "ln_2.bias", st));
  }
  // token assembly + ln_pre, then the patch-embedding weight gradient per image chunk
  {
    float *dcls = G(v + "class_embedding"), *dpos = G(v + "positional_embedding"), *dgp = G(v + "ln_pre.weight"), *dbp = G(v + "ln_pre.bias"),
          *dWc = G(v + "conv1.weight");
    if (err) return err;
    MSQ_CUDA(cudaMemsetAsync(b.dpatch, 0, (size_t)n_img * g2 * Wd * sizeof(float), st));
    MSQ_TRY(vit_assemble_bwd(b.gB, ts->patch, ts->img_index, R, 2, g2, Wd, m->vit_cls, m->vit_pos, m->ln_pre.g, 1e-5f, b.dpatch, dcls, dpos, dgp,
                             dbp, b.ln_scr, st));
    for (int64_t i0 = 0; i0 < n_img; i0 += TRAIN_IMG_CHUNK) {
      const int64_t n = min(TRAIN_IMG_CHUNK, n_img - i0), rows = n * g2, Mp = round_up(rows, 64);
      MSQ_TRY(im2col<T>(ts->images + i0 * 3 * c.vit_res * c.vit_res, n, c.vit_res, c.vit_patch, (T*)b.apatch, st));
      MSQ_TRY((transpose_pad<float, T>(b.dpatch + i0 * g2 * Wd, rows, Wd, Wd, Mp, (T*)b.GT, ACT_NONE, st)));
      MSQ_TRY((transpose_pad<T, T>((const T*)b.apatch, rows, Kc, Kc, Mp, (T*)b.XT, ACT_NONE, st)));
      MSQ_TRY((gemm_nt<T, float>(m, (const T*)b.GT, (int)Mp, (const T*)b.XT, (int)Mp, nullptr, dWc, Kc, dWc, Kc, Wd, Kc, (int)Mp, ACT_NONE, st)));
    }
    MSQ_TRY(mark_ready(ts, v + "conv1.weight", v + "ln_post.bias", st));
  }
  return MSQ_OK;
}

}  // namespace msq

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" int64_t msq_train_param_count(msq_model* m, void* stream) {
  DevGuard dev_guard__(m);
  if (ensure_train(m, (cudaStream_t)stream) != MSQ_OK) return -1;
  return (int64_t)m->train->slots.size();
}
extern "C" int64_t msq_train_grad_numel(msq_model* m, void* stream) {
  DevGuard dev_guard__(m);
  if (ensure_train(m, (cudaStream_t)stream) != MSQ_OK) return -1;
  return m->train->total;
}
extern "C" int msq_train_param_info(msq_model* m, int64_t i, const char** name, int64_t* offset, int64_t* numel, int32_t* decay) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && m->train && i >= 0 && i < (int64_t)m->train->slots.size(), "msq_train_param_info: bad index / no training state");
  const ParamSlot& s = m->train->slots[(size_t)i];
  if (name) *name = s.name.c_str();
  if (offset) *offset = s.off;
  if (numel) *numel = s.numel;
  if (decay) *decay = s.decay ? 1 : 0;
  return MSQ_OK;
}
extern "C" int msq_train_read_param(msq_model* m, const char* name, float* out_dev, int64_t numel, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && name && out_dev, "null argument");
  auto it = m->raw.find(name);
  MSQ_REQUIRE(it != m->raw.end() && it->second.second == numel, "msq_train_read_param: unknown weight %s or wrong size", name);
  MSQ_CUDA(cudaMemcpyAsync(out_dev, it->second.first, (size_t)numel * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return MSQ_OK;
}

extern "C" int msq_inner_forward_train(msq_model* m, const int64_t* ids_dev, const int64_t* tt_dev, const int64_t* mask_dev, int64_t R,
                                       int32_t Lt, const float* images_dev, int64_t n_img, const int32_t* img_index_dev, float* lang_dev,
                                       float* visn_dev, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && ids_dev && tt_dev && mask_dev, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  MSQ_TRY(ensure_train(m, st));
  if (m->cfg.precise) return forward_train<float>(m, ids_dev, tt_dev, mask_dev, R, Lt, images_dev, n_img, img_index_dev, lang_dev, visn_dev, st);
  return forward_train<bf16>(m, ids_dev, tt_dev, mask_dev, R, Lt, images_dev, n_img, img_index_dev, lang_dev, visn_dev, st);
}

extern "C" int msq_inner_backward(msq_model* m, const float* d_lang_dev, const float* d_visn_dev, float* grads_dev, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && grads_dev, "null argument");
  MSQ_REQUIRE(((uintptr_t)grads_dev & 255) == 0, "msq_inner_backward: the gradient buffer must be 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (m->cfg.precise) return backward_train<float>(m, d_lang_dev, d_visn_dev, grads_dev, st);
  return backward_train<bf16>(m, d_lang_dev, d_visn_dev, grads_dev, st);
}

extern "C" int msq_adamw_step(msq_model* m, const float* grads_dev, float lr, float beta1, float beta2, float eps, float weight_decay,
                              float max_grad_norm, float grad_scale, float* norm_out_dev, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && grads_dev, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  MSQ_TRY(ensure_train(m, st));
  TrainState* ts = m->train;
  if (!ts->adam_m) {
    MSQ_CUDA(cudaMalloc(&ts->adam_m, (size_t)ts->total * sizeof(float)));
    MSQ_CUDA(cudaMalloc(&ts->adam_v, (size_t)ts->total * sizeof(float)));
    MSQ_CUDA(cudaMalloc(&ts->opt_scratch, grad_norm_scratch_floats() * sizeof(float)));
    MSQ_CUDA(cudaMemsetAsync(ts->adam_m, 0, (size_t)ts->total * sizeof(float), st));
    MSQ_CUDA(cudaMemsetAsync(ts->adam_v, 0, (size_t)ts->total * sizeof(float), st));
    std::vector<AdamChunk> ch;
    constexpr int64_t CH = 16384;
    for (const ParamSlot& s : ts->slots)
      for (int64_t o = 0; o < s.numel; o += CH) ch.push_back(AdamChunk{s.master + o, s.off + o, (int32_t)min(CH, s.numel - o), s.decay ? 1 : 0});
    ts->n_chunks = (int)ch.size();
    MSQ_CUDA(cudaMalloc(&ts->adam_chunks, ch.size() * sizeof(AdamChunk)));
    MSQ_CUDA(cudaMemcpyAsync(ts->adam_chunks, ch.data(), ch.size() * sizeof(AdamChunk), cudaMemcpyHostToDevice, st));
    MSQ_CUDA(cudaStreamSynchronize(st));   // `ch` is a host temporary
  }
  ++ts->step;
  // padding between slots is never written by the backward pass (the caller zeroes the buffer), so the norm over the
  // whole flat buffer is the norm over the parameters' gradients
  MSQ_TRY(grad_norm_clip(grads_dev, ts->total, max_grad_norm, grad_scale, ts->opt_scratch, st));
  MSQ_TRY(adamw_update_multi(ts->adam_chunks, ts->n_chunks, grads_dev, ts->adam_m, ts->adam_v, lr, beta1, beta2, eps, weight_decay, ts->step,
                             ts->opt_scratch, st));
  if (norm_out_dev) MSQ_CUDA(cudaMemcpyAsync(norm_out_dev, ts->opt_scratch, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  // masters changed: rebuild the packed copies (fused QKV, bf16, folded LayerNorm, ...) and the W^T operands
  return msq::model_weights_changed(m, st);
}

// every derived copy of the fp32 masters: packed inference weights, and (when a training state exists) the W^T operands
int msq::model_weights_changed(msq_model* m, cudaStream_t st) {
  MSQ_TRY(model_repack(m, st));
  TrainState* ts = m->train;
  if (!ts) return MSQ_OK;
  if (m->cfg.precise) MSQ_TRY(refresh_wT<float>(m, ts, st));
  else MSQ_TRY(refresh_wT<bf16>(m, ts, st));
  return ts->heads ? heads_train_setup(m, false, st) : MSQ_OK;
}

/* In-place weight refresh for callers that own the fp32 masters themselves (the unchanged reference trainer: its optimizer
 * updates nn.Parameter.data): msq_model_update_weight overwrites the registered copy of `name` (same element count, device
 * -> device, no allocation), msq_model_refresh re-derives every packed copy afterwards.  Nothing is rebuilt or reallocated. */
extern "C" int msq_model_update_weight(msq_model* m, const char* name, const float* data_dev, int64_t numel, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && name && data_dev, "bad argument");
  auto it = m->raw.find(name);
  MSQ_REQUIRE(it != m->raw.end(), "msq_model_update_weight: %s was never registered", name);
  MSQ_REQUIRE(it->second.second == numel, "msq_model_update_weight: %s has %lld elements, got %lld", name,
              (long long)it->second.second, (long long)numel);
  MSQ_CUDA(cudaMemcpyAsync(it->second.first, data_dev, numel * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return MSQ_OK;
}
extern "C" int msq_model_refresh(msq_model* m, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && m->packed, "msq_model_refresh: model not packed");
  return msq::model_weights_changed(m, (cudaStream_t)stream);
}

/* Optional time-contrastive objective (models/berson/modeling_bert.py:1176-1216, args.additional_wrapper_level_objectives):
 * the NEXT msq_train_step adds weight * TripletMarginLoss(margin 1, p 2) over the sentence vectors sents[b, triplets[b, 0..2]]
 * (anchor, positive, negative; the caller draws them as the reference does, with numpy's generator) to the loss and to the
 * gradients.  One-shot: the step after that runs the default objective again. */
extern "C" int msq_train_set_triplets(msq_model* m, const int32_t* triplets_dev, int64_t B, float weight, void* stream) {
  DevGuard dev_guard__(m);
  cudaStream_t st = (cudaStream_t)stream;
  MSQ_TRY(ensure_train(m, st));
  MSQ_REQUIRE(triplets_dev && B >= 1, "msq_train_set_triplets: null / empty");
  TrainState* ts = m->train;
  if (B > ts->trip_cap) {
    if (ts->trip) MSQ_CUDA(cudaFree(ts->trip));
    if (ts->trip_loss) MSQ_CUDA(cudaFree(ts->trip_loss));
    ts->trip = nullptr; ts->trip_loss = nullptr; ts->trip_cap = 0;
    MSQ_CUDA(cudaMalloc(&ts->trip, (size_t)B * 3 * sizeof(int32_t)));
    MSQ_CUDA(cudaMalloc(&ts->trip_loss, (size_t)B * sizeof(float)));
    ts->trip_cap = B;
  }
  MSQ_CUDA(cudaMemcpyAsync(ts->trip, triplets_dev, (size_t)B * 3 * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  ts->trip_B = B; ts->trip_weight = weight;
  return MSQ_OK;
}

/* args.multimodal_loss of the reference (modeling_bert.py:897-898, 1359-1364, 1218-1225): when on, every msq_train_step adds
 * lam * mean_b sum_p NLL(softmax(pairwise_relationship(img_projection(visn[p, 0]))), pairwise_label_p) / P to the loss, with the
 * gradients of img_projection.*, of the shared pairwise_relationship.* and of the first visual token of every pair row.
 * img_projection.weight [H, H] / .bias [H] must have been registered with msq_model_set_weight BEFORE the first training call
 * (the reference builds Linear(v_feature_size, H) and feeds it the H-d token: it only runs when v_feature_size == H). */
extern "C" int msq_train_set_multimodal_loss(msq_model* m, int32_t on, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_TRY(ensure_train(m, (cudaStream_t)stream));
  TrainState* ts = m->train;
  if (on) {
    MSQ_REQUIRE(ts->heads && m->cfg.vit_width != 0, "msq_train_set_multimodal_loss: needs a multimodal model with BERSON heads");
    MSQ_REQUIRE(ts->index.count("img_projection.weight") && ts->index.count("img_projection.bias"),
                "msq_train_set_multimodal_loss: img_projection.weight / .bias were not registered before the training state was built");
  }
  ts->mm_loss = on != 0;
  return MSQ_OK;
}

extern "C" int msq_train_set_bn_mode(msq_model* m, int32_t use_running_stats, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_TRY(ensure_train(m, (cudaStream_t)stream));
  if (m->train->rn) m->train->rn->bn_eval = use_running_stats != 0;
  return MSQ_OK;
}

// head parameters are the tail of the slot table (added after the encoder's); their gradients are final once heads_train returns
static int mark_heads_ready(msq_model* m, cudaStream_t st) {
  TrainState* ts = m->train;
  ts->ready_used = 0;
  const std::vector<std::string> names = heads_param_names(m);
  if (names.empty()) return MSQ_OK;
  return mark_ready(ts, names.front(), names.back(), st);   // slots were added in this order: first .. last are contiguous
}

/* Dropout of the fine-tuning path (the reference trains with hidden_dropout_prob = attention_probs_dropout_prob = 0.1,
 * BertConfig; args.para_dropout for the paragraph encoder).  Probabilities in [0, 1); 0 = off (the default).  Masks are a
 * pure function of (seed, forward counter, site, element index): see csrc/dropout.cuh / oracle/dropout.py. */
extern "C" int msq_train_set_dropout(msq_model* m, float p_hidden, float p_attn, float p_para, uint32_t seed, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_TRY(ensure_train(m, (cudaStream_t)stream));
  MSQ_REQUIRE(p_hidden >= 0.f && p_hidden < 1.f && p_attn >= 0.f && p_attn < 1.f && p_para >= 0.f && p_para < 1.f, "dropout probabilities must lie in [0, 1)");
  TrainState* ts = m->train;
  ts->drop.seed = seed; ts->drop.p_hidden = p_hidden; ts->drop.p_attn = p_attn; ts->drop.p_para = p_para;
  return MSQ_OK;
}
/* counter of the LAST training forward (what keyed its masks); -1 before the first one */
extern "C" int64_t msq_train_dropout_step(msq_model* m) {
  return (m && m->train && m->train->drop_next_step > 0) ? (int64_t)m->train->drop.step : -1;
}

/* Gradient-ready regions of the last msq_train_step, in completion order (see TrainState::ready_ev): count, the element
 * range of region i, and "make `stream` wait until region i is final". */
extern "C" int64_t msq_train_ready_count(msq_model* m) { return (m && m->train) ? (int64_t)m->train->ready_used : 0; }
extern "C" int msq_train_ready_info(msq_model* m, int64_t i, int64_t* begin, int64_t* end) {
  MSQ_REQUIRE(m && m->train && i >= 0 && (size_t)i < m->train->ready_used && begin && end, "msq_train_ready_info: bad index");
  *begin = m->train->ready_rng[i].first;
  *end = m->train->ready_rng[i].second;
  return MSQ_OK;
}
extern "C" int msq_train_ready_wait(msq_model* m, int64_t i, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && m->train && i >= 0 && (size_t)i < m->train->ready_used, "msq_train_ready_wait: bad index");
  MSQ_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, m->train->ready_ev[i], 0));
  return MSQ_OK;
}

extern "C" int msq_train_step(msq_model* m, const int64_t* ids_dev, const int64_t* tt_dev, const int64_t* mask_dev, const int64_t* sep_dev,
                              int64_t B, int32_t N, int32_t Lt, const float* images_dev, int64_t n_img, const int32_t* img_index_dev,
                              const int32_t* ground_truth_dev, const int64_t* pairwise_labels_dev, float lam, float* grads_dev, float* loss_dev,
                              void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && ids_dev && tt_dev && mask_dev && sep_dev && ground_truth_dev && pairwise_labels_dev && grads_dev, "null argument");
  MSQ_REQUIRE(((uintptr_t)grads_dev & 255) == 0, "msq_train_step: the gradient buffer must be 256-byte aligned");
  MSQ_REQUIRE(B >= 1 && N >= 2 && N <= 16, "msq_train_step: B=%lld N=%d", (long long)B, N);
  cudaStream_t st = (cudaStream_t)stream;
  MSQ_TRY(ensure_train(m, st));
  MSQ_REQUIRE(m->train->heads, "msq_train_step: the model has no BERSON head weights");
  MSQ_REQUIRE(m->cfg.vit_width == 0 || images_dev, "msq_train_step: multimodal model needs images");
  const int64_t R = B * N * (N - 1);
  if (m->cfg.precise) {
    MSQ_TRY(forward_train<float>(m, ids_dev, tt_dev, mask_dev, R, Lt, images_dev, n_img, img_index_dev, nullptr, nullptr, st));
    MSQ_TRY(heads_train<float>(m, sep_dev, B, N, ground_truth_dev, pairwise_labels_dev, lam, grads_dev, loss_dev, st));
    MSQ_TRY(mark_heads_ready(m, st));
    return backward_train<float>(m, m->train->ht.d_lang, nullptr, grads_dev, st);
  }
  MSQ_TRY(forward_train<bf16>(m, ids_dev, tt_dev, mask_dev, R, Lt, images_dev, n_img, img_index_dev, nullptr, nullptr, st));
  MSQ_TRY(heads_train<bf16>(m, sep_dev, B, N, ground_truth_dev, pairwise_labels_dev, lam, grads_dev, loss_dev, st));
  MSQ_TRY(mark_heads_ready(m, st));
  return backward_train<bf16>(m, m->train->ht.d_lang, nullptr, grads_dev, st);
}
