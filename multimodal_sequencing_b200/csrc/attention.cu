// Fused multi-head self-attention over one token group (a pair's joint 227-token sequence, or its
// 99 visual tokens): scores, additive key mask, softmax and P.V in one kernel; the whole K and V of a
// (group, head) live in shared memory so the softmax is single-pass (no online rescale).
//
// Replaces: BertAttention.forward  models/CLIP/src/lxrt/modeling.py:398-425 (and the text-only twin
//           models/berson/modeling_bert.py:208-241): scores / sqrt(d) + mask(-10000), softmax, P V;
//           nn.MultiheadAttention inside ResidualAttentionBlock  models/CLIP/clip/model.py:204-226.
// Input layout: qkv [R*L, 3*heads*64] rows = tokens, columns = (q | k | v), each heads x 64.
#include "kernels.cuh"

namespace msq {

constexpr int AT_D = 64;
constexpr int AT_WARPS = 16;

// fp32 CUDA-core version (exact-parity path; also serves bf16 I/O).
template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32) attention_simt_kernel(const T* __restrict__ qkv, int L, int heads,
                                                                       float scale, const float* __restrict__ mask_add,
                                                                       int mask_ld, int mask_len, T* __restrict__ ctx) {
  extern __shared__ __align__(16) float sm[];
  const int r = blockIdx.x / heads, h = blockIdx.x % heads;
  const int ld = 3 * heads * AT_D;
  float* Ks = sm;                        // [L][65]
  float* Vs = Ks + L * 65;               // [L][64]
  float* Ms = Vs + L * AT_D;             // [L] additive mask
  float* Qs = Ms + ((L + 3) & ~3);       // [warps][64]
  float* Ps = Qs + AT_WARPS * AT_D;      // [warps][Lpad]
  const int Lpad = (L + 31) & ~31;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* base = qkv + (int64_t)r * L * ld + h * AT_D;

  for (int i = threadIdx.x; i < L * (AT_D / 2); i += blockDim.x) {
    const int t = i / (AT_D / 2), d = (i % (AT_D / 2)) * 2;
    const T* kp = base + (int64_t)t * ld + heads * AT_D + d;
    const T* vp = kp + heads * AT_D;
    Ks[t * 65 + d] = to_f(kp[0]);
    Ks[t * 65 + d + 1] = to_f(kp[1]);
    Vs[t * AT_D + d] = to_f(vp[0]);
    Vs[t * AT_D + d + 1] = to_f(vp[1]);
  }
  for (int t = threadIdx.x; t < L; t += blockDim.x)
    Ms[t] = (mask_add && t < mask_len) ? mask_add[(int64_t)r * mask_ld + t] : 0.f;
  __syncthreads();

  float* q = Qs + warp * AT_D;
  float* p = Ps + warp * Lpad;
  const int nch = Lpad / 32;
  for (int t = warp; t < L; t += AT_WARPS) {
    const T* qp = base + (int64_t)t * ld;
    q[lane] = to_f(qp[lane]);
    q[lane + 32] = to_f(qp[lane + 32]);
    __syncwarp();
    float mx = -INFINITY;
    for (int c = 0; c < nch; ++c) {
      const int key = c * 32 + lane;
      float s = -INFINITY;
      if (key < L) {
        const float* kr = Ks + key * 65;
        float a = 0.f;
#pragma unroll 16
        for (int d = 0; d < AT_D; ++d) a = fmaf(q[d], kr[d], a);
        s = a * scale + Ms[key];
      }
      p[key] = s;
      mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = 0; c < nch; ++c) {
      const int key = c * 32 + lane;
      const float e = key < L ? expf(p[key] - mx) : 0.f;
      p[key] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int key = 0; key < L; ++key) {
      const float pk = p[key];
      o0 = fmaf(pk, Vs[key * AT_D + lane], o0);
      o1 = fmaf(pk, Vs[key * AT_D + lane + 32], o1);
    }
    const float inv = 1.0f / sum;
    T* op = ctx + ((int64_t)r * L + t) * (heads * AT_D) + h * AT_D;
    op[lane] = from_f<T>(o0 * inv);
    op[lane + 32] = from_f<T>(o1 * inv);
    __syncwarp();
  }
}

template <typename T>
int attention(const T* qkv, int64_t R, int L, int heads, int dhead, float scale, const float* key_mask_add, int mask_ld,
              int mask_len, T* ctx, cudaStream_t st) {
  MSQ_REQUIRE(dhead == AT_D, "attention: head dim %d != 64", dhead);
  MSQ_REQUIRE(L >= 1 && L <= 320, "attention: sequence length %d out of range", L);
  if (R == 0) return MSQ_OK;
  const int Lpad = (L + 31) & ~31;
  const size_t smem = sizeof(float) * ((size_t)L * 65 + (size_t)L * AT_D + ((L + 3) & ~3) + AT_WARPS * AT_D + AT_WARPS * Lpad);
  static size_t configured_f = 0, configured_b = 0;
  size_t& configured = sizeof(T) == 4 ? configured_f : configured_b;
  if (smem > configured) {
    MSQ_CUDA(cudaFuncSetAttribute(attention_simt_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  attention_simt_kernel<T><<<(unsigned)(R * heads), AT_WARPS * 32, smem, st>>>(qkv, L, heads, scale, key_mask_add, mask_ld,
                                                                               mask_len, ctx);
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int attention<float>(const float*, int64_t, int, int, int, float, const float*, int, int, float*, cudaStream_t);
template int attention<bf16>(const bf16*, int64_t, int, int, int, float, const float*, int, int, bf16*, cudaStream_t);

}  // namespace msq
