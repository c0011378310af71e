// Model state shared by api.cu (inference orchestration) and train.cu (fine-tuning): workspace arena, packed weight
// descriptors and the msq_model object behind the C ABI handle.
#pragma once
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/msq_b200.h"
#include "kernels.cuh"

namespace msq {

struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  int reserve(size_t bytes, cudaStream_t st) {
    if (bytes <= cap) return MSQ_OK;
    MSQ_CUDA(cudaStreamSynchronize(st));
    if (base) MSQ_CUDA(cudaFree(base));
    base = nullptr;
    cap = 0;
    MSQ_CUDA(cudaMalloc(&base, bytes));
    cap = bytes;
    return MSQ_OK;
  }
  void reset() { off = 0; }
  template <typename T> T* take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += bytes;
    return p;
  }
};
// first pass with base == nullptr measures, second pass hands out pointers
struct Planner {
  Arena* a;
  bool measure;
  size_t need = 0;
  template <typename T> T* take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    if (measure) { need += bytes; return nullptr; }
    return a->take<T>(n);
  }
};

struct Lin {           // y = x W^T + b
  const float* w32 = nullptr;
  const bf16* w16 = nullptr;
  const float* b = nullptr;
  int N = 0, K = 0, ld = 0;
};
struct LNp { const float* g = nullptr; const float* b = nullptr; };
// a linear layer with the LayerNorm in front of it folded in: lin.w16 = bf16(gamma_k W[n,k]), lin.b = b + W beta,
// svec[n] = sum_k lin.w16[n,k]  (GemmArgs EPI_LNFOLD)
struct LinF { Lin lin; const float* svec = nullptr; };

struct BertLayerW { Lin qkv, out, up, down; LNp ln1, ln2; LinF qkv_f, up_f; };   // qkv_f folds the PREVIOUS layer's ln2, up_f this layer's ln1
struct VitLayerW { Lin qkv, out, fc, proj; LNp ln1, ln2; LinF qkv_f, fc_f; };    // qkv_f folds ln_1, fc_f ln_2
struct ParaLayerW { Lin qkv, fin, w1, w2; LNp ln_in, ln_ff; };
struct RnBlockW { Lin c1, c2, c3, ds; bool has_ds = false; int stride = 1, cin = 0, planes = 0; };

struct TrainState;
void train_state_free(TrainState*);

}  // namespace msq

using namespace msq;

struct msq_model {
  msq_config cfg;
  int device = 0;   // CUDA device the model lives on (recorded by msq_model_create); every C ABI entry switches to it
  // host-buffer entry points: staging buffer, copy stream and per-micro-batch events (owned by the model, freed with it)
  char* stage = nullptr;
  size_t stage_cap = 0;
  cudaStream_t copy_st = nullptr;
  std::vector<cudaEvent_t> copy_evs;
  std::unordered_map<std::string, std::pair<float*, int64_t>> raw;
  std::vector<void*> owned;
  std::vector<size_t> owned_bytes;
  bool repacking = false;      // msq_model_pack re-run after an optimizer step: dev_alloc replays `owned`
  size_t repack_cursor = 0;
  TrainState* train = nullptr; // fine-tuning state (train.cu), created on first use
  bool packed = false, has_bert = false, has_vit = false, has_heads = false;
  std::string prefix_inner = "bert.";
  // packed
  const float *word = nullptr, *pos = nullptr, *type = nullptr;
  LNp emb_ln;
  std::vector<BertLayerW> bert;
  Lin pooler;
  bool has_pooler = false;
  // vit
  Lin conv1, visn_fc;
  LinF visn_fc_f;   // visn_fc with the ViT's ln_post folded in
  bool folded = false;
  const float *vit_cls = nullptr, *vit_pos = nullptr;
  LNp ln_pre, ln_post, visn_ln;
  std::vector<VitLayerW> vit;
  // CLIP ModifiedResNet tower (cfg.rn_width != 0)
  Lin rn_stem[3], rn_qkv, rn_cproj;
  std::vector<RnBlockW> rn_blocks;
  const float *rn_pos = nullptr, *rn_posadd = nullptr;
  // berson heads
  Lin sent_tran, key_lin, xg_lin, t4_lin, pwk_raw, wih_raw, whh_raw, wq_raw;
  const bf16 *wih3 = nullptr, *wpw3 = nullptr;   // three-plane (bf16x6) copies of xg_lin / t4_lin weights (tensor-core decode)
  int Kp4 = 0;
  const float *w2 = nullptr, *b2 = nullptr, *w_rel = nullptr, *b_rel = nullptr, *w_in2 = nullptr;
  std::vector<ParaLayerW> para;
  LNp para_ln;
  DecodeWeights dec;
  int Kp = 0;  // padded H+2
  Arena ws;
  ~msq_model() {
    for (auto& kv : raw) cudaFree(kv.second.first);
    for (void* p : owned) cudaFree(p);
    if (ws.base) cudaFree(ws.base);
    if (stage) cudaFree(stage);
    if (copy_st) cudaStreamDestroy(copy_st);
    for (cudaEvent_t e : copy_evs) cudaEventDestroy(e);
    if (train) train_state_free(train);
  }
};

namespace msq {
// switches the calling thread to the model's device for the duration of a C ABI call (restored on return)
struct DevGuard {
  int prev = -1;
  bool sw = false;
  explicit DevGuard(const msq_model* m) {
    if (m && cudaGetDevice(&prev) == cudaSuccess && prev != m->device) sw = cudaSetDevice(m->device) == cudaSuccess;
  }
  ~DevGuard() { if (sw) cudaSetDevice(prev); }
};
int model_repack(msq_model* m, cudaStream_t st);   // api.cu
int model_weights_changed(msq_model* m, cudaStream_t st);   // train.cu: repack + training-state operand copies
bool model_use_tc(const msq_model* m);             // api.cu: tcgen05 path selected for this model / device
}  // namespace msq
