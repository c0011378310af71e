// CLIP ModifiedResNet tower ("RN50", the reference's wired default backbone): the HBM-bound helpers around the
// convolution GEMMs.  Activations are NHWC so that every convolution is `im2col rows x [Cout, kh*kw*Cin]^T` on the
// same tcgen05 / FFMA GEMMs as the transformer layers; eval-mode BatchNorm is folded into the weights at pack time.
//
// Reference ops replaced (telin0411/multimodal_sequencing):
//   Bottleneck.forward        models/CLIP/clip/model.py:10-53   (1x1 -> 3x3 -> AvgPool -> 1x1, AvgPool+1x1 shortcut)
//   ModifiedResNet.forward    models/CLIP/clip/model.py:171-187 (3-conv stem, AvgPool2d(2), 4 stages, attnpool)
//   AttentionPool2d.forward   models/CLIP/clip/model.py:71-125  (pair-joint token matrix, mean token, pos-emb,
//                                                               MHA over all 1+2g^2 tokens, c_proj, cat([x,x]))
//   LinearPositionEmbedding / VisualTokenTypeEmbedding  models/CLIP/src/lxrt/modeling.py:621-705
#include "kernels.cuh"

namespace msq {

namespace {
template <typename T> struct V4;
template <> struct V4<float> {
  static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct V4<bf16> {
  static __device__ __forceinline__ float4 load(const bf16* p) {
    uint2 u = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&u.x), b = *reinterpret_cast<__nv_bfloat162*>(&u.y);
    return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
  }
  static __device__ __forceinline__ void store(bf16* p, float4 v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&a);
    u.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = u;
  }
};
// Row-addressed access to an activation matrix [rows, C] of T.  Split rows (bf16x3 mode) are [hi(C) | lo(C)] bf16.
template <typename T> struct RowIO {
  static __device__ __forceinline__ float4 ld4(const T* base, int64_t row, int C, int c) { return V4<T>::load(base + row * C + c); }
  static __device__ __forceinline__ void st4(T* base, int64_t row, int C, int c, float4 v) { V4<T>::store(base + row * C + c, v); }
  static __device__ __forceinline__ void st1(T* base, int64_t row, int C, int c, float v) { base[row * C + c] = from_f<T>(v); }
};
template <> struct RowIO<bf16s> {
  static __device__ __forceinline__ float4 ld4(const bf16s* base, int64_t row, int C, int c) {
    const bf16* p = reinterpret_cast<const bf16*>(base) + row * 2 * C + c;
    const float4 h = V4<bf16>::load(p), l = V4<bf16>::load(p + C);
    return make_float4(h.x + l.x, h.y + l.y, h.z + l.z, h.w + l.w);
  }
  static __device__ __forceinline__ void st4(bf16s* base, int64_t row, int C, int c, float4 v) {
    bf16* p = reinterpret_cast<bf16*>(base) + row * 2 * C + c;
    uint2 hi, lo;
    split4(v, hi, lo);
    *reinterpret_cast<uint2*>(p) = hi;
    *reinterpret_cast<uint2*>(p + C) = lo;
  }
  static __device__ __forceinline__ void st1(bf16s* base, int64_t row, int C, int c, float v) {
    bf16* p = reinterpret_cast<bf16*>(base) + row * 2 * C + c;
    const bf16 h = __float2bfloat16_rn(v);
    p[0] = h;
    p[C] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
};
inline dim3 grid_for(int64_t work) { return dim3((unsigned)max((int64_t)1, min((int64_t)148 * 16, (work + 255) / 256))); }
}  // namespace

// ---- pack time: conv weight [Cout, Cin, k, k] + eval BatchNorm -> GEMM weight [Cout, k*k*Cin] ((ky,kx,c) order) + bias
__global__ void rn_fold_kernel(const float* __restrict__ w, const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var, int Cout, int Cin, int k,
                               float* __restrict__ wf, float* __restrict__ bf) {
  pdl_sync();
  const int K = k * k * Cin;
  const int64_t total = (int64_t)Cout * K;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i / K), col = (int)(i % K);
    const int c = col % Cin, kk = col / Cin, ky = kk / k, kx = kk % k;
    const float s = gamma[o] / sqrtf(var[o] + 1e-5f);
    wf[i] = w[(((int64_t)o * Cin + c) * k + ky) * k + kx] * s;
    if (col == 0) bf[o] = beta[o] - mean[o] * s;
  }
}
int rn_fold(const float* w, const float* gamma, const float* beta, const float* mean, const float* var, int Cout, int Cin, int k,
            float* wf, float* bf, cudaStream_t st) {
  MSQ_CUDA(launch_k(rn_fold_kernel, grid_for((int64_t)Cout * Cin * k * k), dim3(256), 0, st, w, gamma, beta, mean, var, Cout, Cin, k,
                    wf, bf));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// ---- pack time: posadd[t, :] = x_emb[cell / g] + y_emb[cell % g] + type_emb[t > g2]   (t = 0: cell 0; lxrt/modeling.py:649-651, 691-696)
__global__ void rn_posadd_kernel(const float* __restrict__ xe, const float* __restrict__ ye, const float* __restrict__ te, int g,
                                 int F, float* __restrict__ out) {
  pdl_sync();
  const int g2 = g * g, L = 1 + 2 * g2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < L * F; i += gridDim.x * blockDim.x) {
    const int t = i / F, f = i % F;
    const int cell = t == 0 ? 0 : (t - 1) % g2;
    out[i] = xe[(cell / g) * F + f] + ye[(cell % g) * F + f] + te[(t > g2 ? 1 : 0) * F + f];
  }
}
int rn_posadd(const float* xe, const float* ye, const float* te, int g, int F, float* out, cudaStream_t st) {
  MSQ_CUDA(launch_k(rn_posadd_kernel, grid_for((int64_t)(1 + 2 * g * g) * F), dim3(256), 0, st, xe, ye, te, g, F, out));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// ---- stem conv1: 3x3 / stride 2 / pad 1 over NCHW fp32 images -> rows [n*(S/2)^2, Kp], col = (ky*3+kx)*3 + c
template <typename T>
__global__ void __launch_bounds__(256) rn_im2col_stem_kernel(const float* __restrict__ img, int64_t n, int S, int Kp,
                                                             T* __restrict__ out) {
  pdl_sync();
  const int Ho = S / 2;
  const int64_t total = n * Ho * Ho * (int64_t)Kp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % Kp);
    const int64_t row = i / Kp;
    float v = 0.f;
    if (col < 27) {
      const int c = col % 3, kk = col / 3, ky = kk / 3, kx = kk % 3;
      const int64_t im = row / (Ho * Ho);
      const int oy = (int)(row % (Ho * Ho)) / Ho, ox = (int)(row % (Ho * Ho)) % Ho;
      const int y = oy * 2 - 1 + ky, x = ox * 2 - 1 + kx;
      if (y >= 0 && y < S && x >= 0 && x < S) v = img[((im * 3 + c) * S + y) * (int64_t)S + x];
    }
    RowIO<T>::st1(out, row, Kp, col, v);
  }
}
template <typename T>
int rn_im2col_stem(const float* img, int64_t n, int S, int Kp, T* out, cudaStream_t st) {
  MSQ_REQUIRE(S % 2 == 0 && Kp >= 27, "rn_im2col_stem: S=%d Kp=%d", S, Kp);
  if (n == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(rn_im2col_stem_kernel<T>, grid_for(n * (S / 2) * (S / 2) * (int64_t)Kp), dim3(256), 0, st, img, n, S, Kp, out));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int rn_im2col_stem<float>(const float*, int64_t, int, int, float*, cudaStream_t);
template int rn_im2col_stem<bf16>(const float*, int64_t, int, int, bf16*, cudaStream_t);
template int rn_im2col_stem<bf16s>(const float*, int64_t, int, int, bf16s*, cudaStream_t);

// ---- 3x3 / stride 1 / pad 1 over NHWC -> rows [n*H*W, Kp], col = (ky*3+kx)*C + c.  16 bytes per thread (8 bf16 or 4 fp32
// channels), 32-bit index arithmetic (one launch covers at most RN_IMG_CHUNK images: < 2^31 vectors).
template <typename T> struct VecIO;
template <> struct VecIO<float> {
  static constexpr int N = 4;
  typedef float4 V;
  static __device__ __forceinline__ V zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
};
template <> struct VecIO<bf16> {
  static constexpr int N = 8;
  typedef uint4 V;
  static __device__ __forceinline__ V zero() { return make_uint4(0u, 0u, 0u, 0u); }
};
template <typename T>
__global__ void __launch_bounds__(256) rn_im2col3_kernel(const T* __restrict__ x, uint32_t n, uint32_t H, uint32_t W, uint32_t C,
                                                         uint32_t Kp, T* __restrict__ out, uint32_t in_pitch, uint32_t out_pitch) {
  pdl_sync();
  typedef typename VecIO<T>::V V;
  constexpr uint32_t VN = VecIO<T>::N;
  const uint32_t KV = Kp / VN, CV = C / VN, HW = H * W;
  const uint32_t total = n * HW * KV;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t row = i / KV, cv = i - row * KV;
    V v = VecIO<T>::zero();
    if (cv < 9 * CV) {
      const uint32_t kk = cv / CV, c = (cv - kk * CV) * VN, ky = kk / 3, kx = kk - ky * 3;
      const uint32_t im = row / HW, pix = row - im * HW, oy = pix / W, ox = pix - oy * W;
      const uint32_t y = oy + ky - 1, xx = ox + kx - 1;   // unsigned wrap-around makes -1 fail the range test
      if (y < H && xx < W) v = *reinterpret_cast<const V*>(x + ((size_t)(im * H + y) * W + xx) * in_pitch + c);
    }
    *reinterpret_cast<V*>(out + (size_t)row * out_pitch + (size_t)cv * VN) = v;
  }
}
template <typename T>
int rn_im2col3(const T* x, int64_t n, int H, int W, int C, int Kp, T* out, cudaStream_t st) {
  if constexpr (is_split<T>::value) {
    // split rows [hi(C) | lo(C)] -> [hi(Kp) | lo(Kp)]: the bf16 kernel once per plane with the row pitches of the split layout
    MSQ_REQUIRE(C % 8 == 0 && Kp % 8 == 0 && Kp >= 9 * C, "rn_im2col3: C=%d Kp=%d", C, Kp);
    MSQ_REQUIRE(n * H * W * (int64_t)(Kp / 8) < ((int64_t)1 << 31), "rn_im2col3: launch too large");
    if (n == 0) return MSQ_OK;
    const bf16* xi = reinterpret_cast<const bf16*>(x);
    bf16* xo = reinterpret_cast<bf16*>(out);
    for (int pl = 0; pl < 2; ++pl) {
      MSQ_CUDA(launch_k(rn_im2col3_kernel<bf16>, grid_for(n * H * W * (int64_t)(Kp / 8)), dim3(256), 0, st, xi + pl * C, (uint32_t)n, (uint32_t)H,
                        (uint32_t)W, (uint32_t)C, (uint32_t)Kp, xo + pl * Kp, (uint32_t)(2 * C), (uint32_t)(2 * Kp)));
      MSQ_LAUNCH_CHECK();
    }
    return MSQ_OK;
  } else {
  constexpr int VN = VecIO<T>::N;
  MSQ_REQUIRE(C % VN == 0 && Kp % VN == 0 && Kp >= 9 * C, "rn_im2col3: C=%d Kp=%d", C, Kp);
  MSQ_REQUIRE(n * H * W * (int64_t)(Kp / VN) < ((int64_t)1 << 31), "rn_im2col3: launch too large");
  if (n == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(rn_im2col3_kernel<T>, grid_for(n * H * W * (int64_t)(Kp / VN)), dim3(256), 0, st, x, (uint32_t)n, (uint32_t)H,
                    (uint32_t)W, (uint32_t)C, (uint32_t)Kp, out, (uint32_t)C, (uint32_t)Kp));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
  }
}
template int rn_im2col3<float>(const float*, int64_t, int, int, int, int, float*, cudaStream_t);
template int rn_im2col3<bf16>(const bf16*, int64_t, int, int, int, int, bf16*, cudaStream_t);
template int rn_im2col3<bf16s>(const bf16s*, int64_t, int, int, int, int, bf16s*, cudaStream_t);

// ---- AvgPool2d(2) over NHWC
template <typename T>
__global__ void __launch_bounds__(256) rn_avgpool2_kernel(const T* __restrict__ x, int64_t n, int H, int W, int C,
                                                          T* __restrict__ out) {
  pdl_sync();
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  const int64_t total = n * Ho * Wo * (int64_t)C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const int64_t pix = i / C4, im = pix / (Ho * Wo);
    const int oy = (int)(pix % (Ho * Wo)) / Wo, ox = (int)(pix % (Ho * Wo)) % Wo;
    const int64_t r00 = (im * H + 2 * oy) * W + 2 * ox;
    const float4 a = RowIO<T>::ld4(x, r00, C, c), b = RowIO<T>::ld4(x, r00 + 1, C, c), d = RowIO<T>::ld4(x, r00 + W, C, c),
                 e = RowIO<T>::ld4(x, r00 + W + 1, C, c);
    RowIO<T>::st4(out, pix, C, c, make_float4((a.x + b.x + d.x + e.x) * 0.25f, (a.y + b.y + d.y + e.y) * 0.25f,
                                              (a.z + b.z + d.z + e.z) * 0.25f, (a.w + b.w + d.w + e.w) * 0.25f));
  }
}
template <typename T>
int rn_avgpool2(const T* x, int64_t n, int H, int W, int C, T* out, cudaStream_t st) {
  MSQ_REQUIRE(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "rn_avgpool2: H=%d W=%d C=%d", H, W, C);
  if (n == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(rn_avgpool2_kernel<T>, grid_for(n * (H / 2) * (W / 2) * (int64_t)(C / 4)), dim3(256), 0, st, x, n, H, W, C, out));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int rn_avgpool2<float>(const float*, int64_t, int, int, int, float*, cudaStream_t);
template int rn_avgpool2<bf16>(const bf16*, int64_t, int, int, int, bf16*, cudaStream_t);
template int rn_avgpool2<bf16s>(const bf16s*, int64_t, int, int, int, bf16s*, cudaStream_t);

// ---- x <- relu(x) in place (fp32 residual stream) + operand-type copy for the next convolution
template <typename T>
__global__ void __launch_bounds__(256) rn_relu_cast_kernel(float* __restrict__ x, int64_t n4, int C, T* __restrict__ out) {
  pdl_sync();
  const int C4 = C / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = *reinterpret_cast<float4*>(x + i * 4);
    v = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    *reinterpret_cast<float4*>(x + i * 4) = v;
    RowIO<T>::st4(out, i / C4, C, (int)(i % C4) * 4, v);
  }
}
template <typename T>
int rn_relu_cast(float* x, int64_t n, int C, T* out, cudaStream_t st) {
  MSQ_REQUIRE(C % 4 == 0 && n % C == 0, "rn_relu_cast: n=%lld C=%d", (long long)n, C);
  if (n == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(rn_relu_cast_kernel<T>, grid_for(n / 4), dim3(256), 0, st, x, n / 4, C, out));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int rn_relu_cast<float>(float*, int64_t, int, float*, cudaStream_t);
template int rn_relu_cast<bf16>(float*, int64_t, int, bf16*, cudaStream_t);
template int rn_relu_cast<bf16s>(float*, int64_t, int, bf16s*, cudaStream_t);

// ---- AttentionPool2d token matrix for R pair rows.  feat [n_img, g2, C] fp32 (NHWC).  The reference reshapes the NCHW
// maps of a pair [2, C, g2] to [C, 2*g2] WITHOUT moving the image axis (model.py:76), so token t / channel c' is flat
// element f = c'*(2*g2) + t of that block: image f / (C*g2), channel (f % (C*g2)) / g2, cell f % g2.
// out[r, 0, :] = mean over tokens + pos[0];  out[r, 1+t, :] = token + pos[1+t] (t < g2) or pos[t-g2] (t >= g2).
template <typename T>
__global__ void __launch_bounds__(128) rn_tokens_kernel(const float* __restrict__ feat, const int32_t* __restrict__ img_index,
                                                        int g2, int C, const float* __restrict__ pos, T* __restrict__ out) {
  pdl_sync();
  const int64_t r = blockIdx.x;
  const int cp = blockIdx.y * blockDim.x + threadIdx.x;
  if (cp >= C) return;
  const int L2 = 2 * g2;
  const float* f0 = feat + (int64_t)img_index[r * 2] * g2 * C;
  const float* f1 = feat + (int64_t)img_index[r * 2 + 1] * g2 * C;
  const int64_t orow = r * (int64_t)(1 + L2);
  float sum = 0.f;
  for (int t = 0; t < L2; ++t) {
    const int f = cp * L2 + t, im = f / (C * g2), rem = f % (C * g2), c = rem / g2, p = rem % g2;
    const float v = (im ? f1 : f0)[(int64_t)p * C + c];
    sum += v;
    const int pr = t < g2 ? 1 + t : t - g2;
    RowIO<T>::st1(out, orow + 1 + t, C, cp, v + pos[(int64_t)pr * C + cp]);
  }
  RowIO<T>::st1(out, orow, C, cp, sum / (float)L2 + pos[cp]);
}
template <typename T>
int rn_tokens(const float* feat, const int32_t* img_index, int64_t R, int g2, int C, const float* pos, T* out, cudaStream_t st) {
  if (R == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(rn_tokens_kernel<T>, dim3((unsigned)R, (unsigned)ceil_div(C, 128)), dim3(128), 0, st, feat, img_index, g2, C, pos, out));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int rn_tokens<float>(const float*, const int32_t*, int64_t, int, int, const float*, float*, cudaStream_t);
template int rn_tokens<bf16>(const float*, const int32_t*, int64_t, int, int, const float*, bf16*, cudaStream_t);
template int rn_tokens<bf16s>(const float*, const int32_t*, int64_t, int, int, const float*, bf16s*, cudaStream_t);

// ---- tower output: out[row, :] = cat(o[row], o[row]) (+ posadd[row % L])   (model.py:106, lxrt/modeling.py:1014-1030)
template <typename T>
__global__ void __launch_bounds__(256) rn_finish_kernel(const float* __restrict__ o, int64_t rows, int L, int E,
                                                        const float* __restrict__ posadd, T* __restrict__ out) {
  pdl_sync();
  const int F4 = 2 * E / 4;
  const int64_t total = rows * F4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % F4) * 4;
    const int64_t row = i / F4;
    float4 v = *reinterpret_cast<const float4*>(o + row * E + (f % E));
    if (posadd) {
      const float4 a = *reinterpret_cast<const float4*>(posadd + (row % L) * (int64_t)(2 * E) + f);
      v = make_float4(v.x + a.x, v.y + a.y, v.z + a.z, v.w + a.w);
    }
    RowIO<T>::st4(out, row, 2 * E, f, v);
  }
}
template <typename T>
int rn_finish(const float* o, int64_t rows, int L, int E, const float* posadd, T* out, cudaStream_t st) {
  MSQ_REQUIRE(E % 4 == 0, "rn_finish: E=%d", E);
  if (rows == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(rn_finish_kernel<T>, grid_for(rows * (2 * E / 4)), dim3(256), 0, st, o, rows, L, E, posadd, out));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int rn_finish<float>(const float*, int64_t, int, int, const float*, float*, cudaStream_t);
template int rn_finish<bf16>(const float*, int64_t, int, int, const float*, bf16*, cudaStream_t);
template int rn_finish<bf16s>(const float*, int64_t, int, int, const float*, bf16s*, cudaStream_t);

}  // namespace msq
