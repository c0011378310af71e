// Counter-based dropout for the fine-tuning path: the keep / drop decision of an element is a pure function of
// (seed, optimizer-step counter, site, element index), so the backward pass REGENERATES the mask instead of storing it
// and the CPU oracle reproduces it bit for bit (oracle/dropout.py).  The murmur3 finaliser over index ^ a key that
// mixes seed / step / site; element index = the flat index of the tensor nn.Dropout sees in the reference:
//   E  embeddings [R,Lt,H]            lxrt/modeling.py:369   (modeling_bert.py:179)
//   V  visn_fc output [R,Lv,H]        lxrt/modeling.py:601
//   A  attention_probs [R,h,L,L]      lxrt/modeling.py:419   (modeling_bert.py:228)
//   O  attention output dense [R,L,H] lxrt/modeling.py:437   (modeling_bert.py:253)
//   F  FFN output dense [R,L,H]       lxrt/modeling.py:491   (modeling_bert.py:319)
//   H  token-attention probs [R,2,Lt] modeling_bert.py:735
//   PA paragraph attention [B,h,N,N]  neural.py:228;  PC context [B,N,H] encoder.py:28;  PF1 [B,N,ff] / PF2 [B,N,H] neural.py:31-32
#pragma once
#include <stdint.h>

#include "common.cuh"

namespace msq {

enum DropKind { DROP_E = 1, DROP_V = 2, DROP_A = 3, DROP_O = 4, DROP_F = 5, DROP_H = 6, DROP_PA = 7, DROP_PC = 8, DROP_PF1 = 9, DROP_PF2 = 10 };

struct DropCfg {   // lives in the training state (msq_train_set_dropout)
  uint32_t seed = 0, step = 0;
  float p_hidden = 0.f, p_attn = 0.f, p_para = 0.f;
};
struct Drop {      // one site of one step, passed to kernels by value; thresh == 0 <=> dropout off
  uint32_t key = 0, thresh = 0;
  float scale = 1.f;
};

__host__ __device__ __forceinline__ uint32_t drop_fmix(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
inline Drop make_drop(const DropCfg& c, int kind, int layer, float p) {
  Drop d;
  if (!(p > 0.f)) return d;
  const uint32_t site = (uint32_t)kind + 16u * (uint32_t)layer;
  d.key = drop_fmix(c.seed ^ drop_fmix(c.step + 0x9E3779B9u) ^ (site * 0x85EBCA6Bu));
  d.thresh = (uint32_t)((double)p * 4294967296.0);
  d.scale = 1.0f / (1.0f - p);
  return d;
}
// ONE round of the murmur3 finaliser (a bijection of the 32-bit word with full avalanche) over key ^ index; indices beyond 2^32
// fold their high word in first.  (Two rounds cost ~8.5 ms of a 92 ms fine-tuning step in the attention kernels alone: the
// decision is evaluated for every probability in the forward pass and twice more in the backward pass.)
__device__ __forceinline__ bool drop_keep(const Drop& d, uint64_t idx) {
  return drop_fmix(d.key ^ (uint32_t)idx ^ ((uint32_t)(idx >> 32) * 0x9E3779B1u)) >= d.thresh;
}
// multiplier of element idx: 0 (dropped) or 1 / (1 - p)
__device__ __forceinline__ float drop_mul(const Drop& d, uint64_t idx) { return d.thresh == 0 ? 1.f : (drop_keep(d, idx) ? d.scale : 0.f); }

}  // namespace msq
