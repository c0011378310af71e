// HBM-bound row kernels of the encoder: LayerNorm (three epsilons), fused BERT embedding-sum-LN,
// fused ViT token assembly + ln_pre, patch im2col, dtype packs.  One warp per row, 128-bit
// accesses, two-pass statistics held in registers (a row is read from HBM exactly once).
//
// Reference ops replaced (telin0411/multimodal_sequencing):
//   LayerNorm                models/CLIP/src/lxrt/modeling.py:353,432,486,577 (eps 1e-12);
//                            models/CLIP/clip/model.py:190-196 (eps 1e-5);
//                            models/berson/encoder.py:16,42, neural.py:25 (eps 1e-6)
//   BertEmbeddings.forward   models/CLIP/src/lxrt/modeling.py:356-370, models/berson/modeling_bert.py:162-180
//   VisualTransformer tokens models/CLIP/clip/model.py:263-276 (class token + pair-joint pos-emb + ln_pre)
#include "kernels.cuh"

namespace msq {

constexpr int LN_MAX_VEC = 8;  // H <= 8 * 128 = 1024

// Normalise the row held in v[] (nv float4 per lane) and write fp32 and/or T copies.
template <typename T>
__device__ __forceinline__ void ln_finish(float4* v, int nv, int H, int lane, const float* __restrict__ gamma,
                                          const float* __restrict__ beta, float eps, float* out_f, T* out_t) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i)
    if (i < nv) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / (float)H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i)
    if (i < nv) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)H + eps);
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i)
    if (i < nv) {
      const int col = (i * 32 + lane) * 4;
      float4 g = *reinterpret_cast<const float4*>(gamma + col), b = *reinterpret_cast<const float4*>(beta + col);
      float4 o = make_float4((v[i].x - mean) * rstd * g.x + b.x, (v[i].y - mean) * rstd * g.y + b.y,
                             (v[i].z - mean) * rstd * g.z + b.z, (v[i].w - mean) * rstd * g.w + b.w);
      if (out_f) Vec4<float>::store(out_f + col, o);
      if (out_t) store_row4<T>(out_t, col, H, o);
    }
}

// out row = (row / in_group) * out_group + out_off + row % in_group   (in_group == 0: identity)
__device__ __forceinline__ int64_t remap_row(int64_t row, int in_group, int out_group, int out_off) {
  return in_group ? (row / in_group) * (int64_t)out_group + out_off + row % in_group : row;
}

template <typename T>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int64_t rows, int H,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float eps, float* __restrict__ out_f, T* __restrict__ out_t,
                                                        int in_group, int out_group, int out_off) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = H >> 7;
  float4 v[LN_MAX_VEC];
  const float* xr = x + row * H;
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i)
    if (i < nv) v[i] = Vec4<float>::load(xr + (i * 32 + lane) * 4);
  const int64_t orow = remap_row(row, in_group, out_group, out_off);
  ln_finish<T>(v, nv, H, lane, gamma, beta, eps, out_f ? out_f + orow * H : nullptr, out_t ? out_t + orow * H : nullptr);
}

template <typename T>
int layernorm(const float* x, int64_t rows, int H, const float* gamma, const float* beta, float eps, float* out_f,
              T* out_t, int in_group, int out_group, int out_off, cudaStream_t st) {
  MSQ_REQUIRE(H % 128 == 0 && H <= LN_MAX_VEC * 128, "layernorm: H=%d must be a multiple of 128 and <= 1024", H);
  if (rows == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(layernorm_kernel<T>, dim3(ceil_div(rows, 8)), dim3(256), 0, st, x, rows, H, gamma, beta, eps, out_f, out_t, in_group, out_group, out_off));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int layernorm<float>(const float*, int64_t, int, const float*, const float*, float, float*, float*, int, int, int,
                              cudaStream_t);
template int layernorm<bf16>(const float*, int64_t, int, const float*, const float*, float, float*, bf16*, int, int, int,
                             cudaStream_t);
template int layernorm<bf16s>(const float*, int64_t, int, const float*, const float*, float, float*, bf16s*, int, int, int,
                              cudaStream_t);

// word[ids] + pos[t] + type[tt] -> LN -> joint rows r*Lj + t
template <typename T>
__global__ void __launch_bounds__(256) embed_ln_kernel(const int64_t* __restrict__ ids, const int64_t* __restrict__ tts,
                                                       int64_t R, int Lt, int Lj, int H, const float* __restrict__ word,
                                                       const float* __restrict__ pos, const float* __restrict__ type,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float eps, float* __restrict__ out_f, T* __restrict__ out_t) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R * Lt) return;
  const int64_t r = row / Lt;
  const int t = (int)(row % Lt);
  const float* w = word + ids[row] * H;
  const float* p = pos + (int64_t)t * H;
  const float* ty = type + tts[row] * H;
  const int nv = H >> 7;
  float4 v[LN_MAX_VEC];
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i)
    if (i < nv) {
      const int col = (i * 32 + lane) * 4;
      float4 a = Vec4<float>::load(w + col), b = Vec4<float>::load(p + col), c = Vec4<float>::load(ty + col);
      v[i] = make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);
    }
  const int64_t orow = r * Lj + t;
  ln_finish<T>(v, nv, H, lane, gamma, beta, eps, out_f + orow * H, out_t ? out_t + orow * H : nullptr);
}

template <typename T>
int embed_ln(const int64_t* ids, const int64_t* tts, int64_t R, int Lt, int Lj, int H, const float* word, const float* pos,
             const float* type, const float* gamma, const float* beta, float eps, float* out_f, T* out_t, cudaStream_t st) {
  MSQ_REQUIRE(H % 128 == 0 && H <= LN_MAX_VEC * 128, "embed_ln: H=%d unsupported", H);
  if (R == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(embed_ln_kernel<T>, dim3(ceil_div(R * Lt, 8)), dim3(256), 0, st, ids, tts, R, Lt, Lj, H, word, pos, type, gamma, beta, eps, out_f, out_t));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int embed_ln<float>(const int64_t*, const int64_t*, int64_t, int, int, int, const float*, const float*,
                             const float*, const float*, const float*, float, float*, float*, cudaStream_t);
template int embed_ln<bf16>(const int64_t*, const int64_t*, int64_t, int, int, int, const float*, const float*,
                            const float*, const float*, const float*, float, float*, bf16*, cudaStream_t);
template int embed_ln<bf16s>(const int64_t*, const int64_t*, int64_t, int, int, int, const float*, const float*,
                             const float*, const float*, const float*, float, float*, bf16s*, cudaStream_t);

// ViT pair token sequence: row (r, t), t in [0, 1 + il*g2): t==0 class token, else patch (t-1)%g2 of
// image slot (t-1)/g2; positional row = t (t <= g2) else (t-1-g2) % g2 [clip/model.py:271-275]; ln_pre.
// patch[] rows are indexed by UNIQUE image: img_index[r*il + slot]*g2 + p.
__global__ void __launch_bounds__(256) vit_assemble_kernel(const float* __restrict__ patch,
                                                           const int32_t* __restrict__ img_index, int64_t R, int il, int g2,
                                                           int W, const float* __restrict__ cls, const float* __restrict__ pos,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float eps, float* __restrict__ out) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const int Lv = 1 + il * g2;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= R * Lv) return;
  const int64_t r = row / Lv;
  const int t = (int)(row % Lv);
  const float* src;
  int prow;
  if (t == 0) {
    src = cls;
    prow = 0;
  } else {
    const int slot = (t - 1) / g2, p = (t - 1) % g2;
    src = patch + ((int64_t)img_index[r * il + slot] * g2 + p) * W;
    prow = slot == 0 ? t : p;
  }
  const float* pp = pos + (int64_t)prow * W;
  const int nv = W >> 7;
  float4 v[LN_MAX_VEC];
#pragma unroll
  for (int i = 0; i < LN_MAX_VEC; ++i)
    if (i < nv) {
      const int col = (i * 32 + lane) * 4;
      float4 a = Vec4<float>::load(src + col), b = Vec4<float>::load(pp + col);
      v[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
  ln_finish<float>(v, nv, W, lane, gamma, beta, eps, out + row * W, nullptr);
}

int vit_assemble(const float* patch, const int32_t* img_index, int64_t R, int il, int g2, int W, const float* cls,
                 const float* pos, const float* gamma, const float* beta, float eps, float* out, cudaStream_t st) {
  MSQ_REQUIRE(W % 128 == 0 && W <= LN_MAX_VEC * 128, "vit_assemble: width=%d unsupported", W);
  if (R == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(vit_assemble_kernel, dim3(ceil_div(R * (1 + il * g2), 8)), dim3(256), 0, st, patch, img_index, R, il, g2, W, cls, pos, gamma, beta, eps, out));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// images [n, 3, S, S] fp32 -> A [n*g*g, 3*P*P] (column order c, ky, kx == conv1.weight flattened)
template <typename T>
__global__ void __launch_bounds__(256) im2col_kernel(const float* __restrict__ img, int64_t n, int S, int P,
                                                     T* __restrict__ out) {
  pdl_sync();
  const int g = S / P;
  const int K = 3 * P * P;
  const int64_t total4 = n * g * g * (int64_t)(K / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / (K / 4);
    const int col = (int)(i % (K / 4)) * 4;
    const int c = col / (P * P), ky = (col / P) % P, kx = col % P;
    const int64_t im = row / (g * g);
    const int py = (int)(row % (g * g)) / g, px = (int)(row % (g * g)) % g;
    const float4 v = *reinterpret_cast<const float4*>(img + ((im * 3 + c) * S + (py * P + ky)) * (int64_t)S + px * P + kx);
    store_row4<T>(out + row * K, col, K, v);
  }
}

template <typename T>
int im2col(const float* img, int64_t n, int S, int P, T* out, cudaStream_t st) {
  MSQ_REQUIRE(S % P == 0 && P % 4 == 0, "im2col: S=%d P=%d unsupported", S, P);
  if (n == 0) return MSQ_OK;
  const int64_t total4 = n * (S / P) * (S / P) * (int64_t)(3 * P * P / 4);
  MSQ_CUDA(launch_k(im2col_kernel<T>, dim3((int)min((int64_t)148 * 16, (total4 + 255) / 256)), dim3(256), 0, st, img, n, S, P, out));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int im2col<float>(const float*, int64_t, int, int, float*, cudaStream_t);
template int im2col<bf16>(const float*, int64_t, int, int, bf16*, cudaStream_t);
template int im2col<bf16s>(const float*, int64_t, int, int, bf16s*, cudaStream_t);

// dst[r, c] = src[(r / group) * src_group + off + r % group, c]  (row gather with dtype conversion)
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) gather_rows_kernel(const TI* __restrict__ src, int64_t rows, int H, int group,
                                                          int src_group, int off, TO* __restrict__ dst) {
  pdl_sync();
  const int64_t total4 = rows * (H / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (H / 4);
    const int c = (int)(i % (H / 4)) * 4;
    const int64_t sr = group ? (r / group) * (int64_t)src_group + off + r % group : r;
    store_row4<TO>(dst + r * H, c, H, Vec4<TI>::load(src + sr * H + c));
  }
}

template <typename TI, typename TO>
int gather_rows(const TI* src, int64_t rows, int H, int group, int src_group, int off, TO* dst, cudaStream_t st) {
  MSQ_REQUIRE(H % 4 == 0, "gather_rows: H=%d", H);
  if (rows == 0) return MSQ_OK;
  const int64_t total4 = rows * (H / 4);
  MSQ_CUDA(launch_k(gather_rows_kernel<TI, TO>, dim3((int)min((int64_t)148 * 16, (total4 + 255) / 256)), dim3(256), 0, st, src, rows, H, group, src_group, off, dst));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int gather_rows<float, float>(const float*, int64_t, int, int, int, int, float*, cudaStream_t);
template int gather_rows<float, bf16>(const float*, int64_t, int, int, int, int, bf16*, cudaStream_t);
template int gather_rows<float, bf16s>(const float*, int64_t, int, int, int, int, bf16s*, cudaStream_t);
template int gather_rows<bf16, bf16>(const bf16*, int64_t, int, int, int, int, bf16*, cudaStream_t);

// weight packing: dst[r, 0..Kp) = src[r, 0..K) zero padded, with dtype conversion (row-major [rows, K])
template <typename TO>
__global__ void pack_pad_kernel(const float* __restrict__ src, int64_t rows, int K, int Kp, TO* __restrict__ dst) {
  pdl_sync();
  const int64_t total = rows * Kp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Kp;
    const int c = (int)(i % Kp);
    const float v = c < K ? src[r * K + c] : 0.f;
    if constexpr (is_split<TO>::value) {   // row = [hi(Kp) | lo(Kp)]
      bf16* d = reinterpret_cast<bf16*>(dst) + r * 2 * Kp;
      const bf16 h = __float2bfloat16_rn(v);
      d[c] = h;
      d[Kp + c] = __float2bfloat16_rn(v - __bfloat162float(h));
    } else {
      dst[i] = from_f<TO>(v);
    }
  }
}
__global__ void pack_split3_kernel(const float* __restrict__ src, int64_t rows, int K, int ld, int Kp, bf16* __restrict__ dst) {
  pdl_sync();
  const int64_t total = rows * Kp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / Kp;
    const int c = (int)(i % Kp);
    const float v = c < K ? src[r * ld + c] : 0.f;
    const bf16 p0 = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(p0);
    const bf16 p1 = __float2bfloat16_rn(r1);
    const bf16 p2 = __float2bfloat16_rn(r1 - __bfloat162float(p1));
    bf16* d = dst + r * 3 * Kp + c;
    d[0] = p0; d[Kp] = p1; d[2 * Kp] = p2;
  }
}
int pack_split3(const float* src, int64_t rows, int K, int ld, int Kp, bf16* dst, cudaStream_t st) {
  if (rows == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(pack_split3_kernel, dim3((int)min((int64_t)148 * 8, (rows * Kp + 255) / 256)), dim3(256), 0, st, src, rows, K, ld, Kp, dst));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

template <typename TO>
int pack_pad(const float* src, int64_t rows, int K, int Kp, TO* dst, cudaStream_t st) {
  if (rows == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(pack_pad_kernel<TO>, dim3((int)min((int64_t)148 * 8, (rows * Kp + 255) / 256)), dim3(256), 0, st, src, rows, K, Kp, dst));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int pack_pad<float>(const float*, int64_t, int, int, float*, cudaStream_t);
template int pack_pad<bf16>(const float*, int64_t, int, int, bf16*, cudaStream_t);
template int pack_pad<bf16s>(const float*, int64_t, int, int, bf16s*, cudaStream_t);

}  // namespace msq
