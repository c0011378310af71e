// Backward / optimizer kernels of the fine-tuning path (reference: autograd through
// models/CLIP/src/lxrt/modeling.py:342-507, 838-1107 and models/CLIP/clip/model.py:190-305, optimizer
// trainers/train.py:172-190, 340-363).  The contractions of the backward pass run on the SAME GEMM kernels as the
// forward pass (gemm_tc.cu / gemm_simt.cu compute C = A W^T with both operands K-major):
//   dgrad  dX[M,K] = dY[M,N] (W^T)[K,N]^T        -> W^T is a packed copy refreshed after every optimizer step
//   wgrad  dW[N,K] += (dY^T)[N,Mp] (X^T)[K,Mp]^T -> transpose_pad writes the zero-padded, M-contiguous operands
// Everything here is the HBM-bound remainder: transposes, LayerNorm backward (dx + deterministic dgamma/dbeta),
// activation backward, attention backward (fp32 CUDA cores, two kernels: dQ then dK/dV), embedding / token-assembly
// scatter, gradient norm + HF-AdamW update.
#include "kernels.cuh"
#include "dropout.cuh"

namespace msq {

int splitk_accumulate(const float* part, int S, int64_t n, float* dW, cudaStream_t st);

// ---------------------------------------------------------------------------------------------------
// dst[n, m] = act(src[m, n]) for m < M, 0 for M <= m < Mp   (dst row-major [N, Mp])
// ---------------------------------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) transpose_pad_kernel(const TI* __restrict__ src, int64_t M, int N, int ld, int64_t Mp,
                                                            TO* __restrict__ dst, int act) {
  pdl_sync();
  __shared__ float tile[32][33];
  const int64_t m0 = (int64_t)blockIdx.x * 32;
  const int n0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int j = ty; j < 32; j += 8) {
    const int64_t m = m0 + j;
    const int n = n0 + tx;
    float v = 0.f;
    if (m < M && n < N) v = apply_act(to_f(src[m * ld + n]), act);
    tile[j][tx] = v;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int n = n0 + j;
    const int64_t m = m0 + tx;
    if (n < N && m < Mp) dst[(int64_t)n * Mp + m] = from_f<TO>(tile[tx][j]);
  }
}
// bf16 -> bf16, 64 x 64 tiles: every global access of a warp is a full 128-byte line (bf16 pairs per lane both ways)
__global__ void __launch_bounds__(256) transpose_pad_bf16_kernel(const bf16* __restrict__ src, int64_t M, int N, int ld, int64_t Mp,
                                                                 bf16* __restrict__ dst, int act) {
  pdl_sync();
  __shared__ float tile[64][65];   // [n][m]
  const int64_t m0 = (int64_t)blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int mi = i * 8 + w;
    const int64_t m = m0 + mi;
    const int n = n0 + 2 * lane;
    float a = 0.f, b = 0.f;
    if (m < M && n < N) {   // N is even: the pair is inside or outside as a whole
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(src + m * ld + n));
      a = apply_act(v.x, act); b = apply_act(v.y, act);
    }
    tile[2 * lane][mi] = a;
    tile[2 * lane + 1][mi] = b;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int ni = i * 8 + w;
    const int n = n0 + ni;
    const int64_t m = m0 + 2 * lane;
    if (n < N && m < Mp) *reinterpret_cast<__nv_bfloat162*>(dst + (int64_t)n * Mp + m) = __floats2bfloat162_rn(tile[ni][2 * lane], tile[ni][2 * lane + 1]);
  }
}

template <typename TI, typename TO>
int transpose_pad(const TI* src, int64_t M, int N, int ld, int64_t Mp, TO* dst, int act, cudaStream_t st) {
  MSQ_REQUIRE(Mp >= M && ceil_div(N, 32) <= 65535, "transpose_pad: bad shape");
  if (Mp == 0 || N == 0) return MSQ_OK;
  if constexpr (sizeof(TI) == 2 && sizeof(TO) == 2) {
    if (N % 2 == 0 && ld % 2 == 0 && Mp % 2 == 0 && (((uintptr_t)src | (uintptr_t)dst) & 3) == 0) {
      MSQ_CUDA(launch_k(transpose_pad_bf16_kernel, dim3((unsigned)ceil_div(Mp, 64), (unsigned)ceil_div(N, 64)), dim3(256), 0, st, (const bf16*)src, M, N, ld, Mp, (bf16*)dst, act));
      MSQ_LAUNCH_CHECK();
      return MSQ_OK;
    }
  }
  MSQ_CUDA(launch_k(transpose_pad_kernel<TI, TO>, dim3((unsigned)ceil_div(Mp, 32), (unsigned)ceil_div(N, 32)), dim3(256), 0, st, src, M, N, ld, Mp, dst, act));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int transpose_pad<float, float>(const float*, int64_t, int, int, int64_t, float*, int, cudaStream_t);
template int transpose_pad<bf16, bf16>(const bf16*, int64_t, int, int, int64_t, bf16*, int, cudaStream_t);
template int transpose_pad<float, bf16>(const float*, int64_t, int, int, int64_t, bf16*, int, cudaStream_t);

// ---------------------------------------------------------------------------------------------------
// out[row] += sum_m a[row, m]   (bias gradients from the transposed, zero-padded gradient operand)
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) rowsum_accum_kernel(const T* __restrict__ a, int64_t ld, int64_t n, float* __restrict__ out) {
  pdl_sync();
  __shared__ float sh[8];
  const T* r = a + (int64_t)blockIdx.x * ld;
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += to_f(r[i]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i];
    out[blockIdx.x] += t;
  }
}
// bf16 rows, 16-byte loads (the operand rows are 64-element aligned and zero padded)
__global__ void __launch_bounds__(256) rowsum_accum_bf16v_kernel(const bf16* __restrict__ a, int64_t ld, int64_t n8, float* __restrict__ out) {
  pdl_sync();
  __shared__ float sh[8];
  const uint4* r = reinterpret_cast<const uint4*>(a + (int64_t)blockIdx.x * ld);
  float s0 = 0.f, s1 = 0.f;
  for (int64_t i = threadIdx.x; i < n8; i += blockDim.x) {
    const uint4 v = r[i];
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
    const float2 a0 = __bfloat1622float2(h[0]), a1 = __bfloat1622float2(h[1]), a2 = __bfloat1622float2(h[2]), a3 = __bfloat1622float2(h[3]);
    s0 += (a0.x + a0.y) + (a1.x + a1.y);
    s1 += (a2.x + a2.y) + (a3.x + a3.y);
  }
  float s = warp_sum(s0 + s1);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i];
    out[blockIdx.x] += t;
  }
}
template <typename T>
int rowsum_accum(const T* a, int rows, int64_t ld, int64_t n, float* out, cudaStream_t st) {
  if (rows == 0) return MSQ_OK;
  if constexpr (sizeof(T) == 2) {
    if (n % 8 == 0 && ld % 8 == 0 && ((uintptr_t)a & 15) == 0) {
      MSQ_CUDA(launch_k(rowsum_accum_bf16v_kernel, dim3(rows), dim3(256), 0, st, (const bf16*)a, ld, n / 8, out));
      MSQ_LAUNCH_CHECK();
      return MSQ_OK;
    }
  }
  MSQ_CUDA(launch_k(rowsum_accum_kernel<T>, dim3(rows), dim3(256), 0, st, a, ld, n, out));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int rowsum_accum<float>(const float*, int, int64_t, int64_t, float*, cudaStream_t);
template int rowsum_accum<bf16>(const bf16*, int, int64_t, int64_t, float*, cudaStream_t);

// out[n] += sum_m G[m, n] for a row-major bf16 [M, N] matrix (bias gradients without a transposed copy): 64 row chunks x
// 64-column groups write partial sums, an ordered second pass adds the chunks (deterministic).
constexpr int CS_CHUNKS = 64;
__global__ void __launch_bounds__(256) colsum_partial_bf16_kernel(const bf16* __restrict__ G, int64_t M, int N, int ld, float* __restrict__ partial) {
  pdl_sync();
  __shared__ float sh[8][65];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int n = blockIdx.x * 64 + 2 * lane;
  const int64_t per = (M + CS_CHUNKS - 1) / CS_CHUNKS, r0 = blockIdx.y * per, r1 = min(M, r0 + per);
  float a = 0.f, b = 0.f;
  if (n < N)
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(G + r * ld + n));
      a += v.x; b += v.y;
    }
  sh[rl][2 * lane] = a; sh[rl][2 * lane + 1] = b;
  __syncthreads();
  if (threadIdx.x < 64 && blockIdx.x * 64 + threadIdx.x < N) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
    partial[(int64_t)blockIdx.y * N + blockIdx.x * 64 + threadIdx.x] = t;
  }
}
__global__ void __launch_bounds__(64) colsum_reduce_kernel(const float* __restrict__ partial, int N, float* __restrict__ out) {
  pdl_sync();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  // all 64 chunk loads are issued before the (ordered) adds: the kernel is a handful of blocks, latency is all it costs
  float v[CS_CHUNKS];
#pragma unroll
  for (int c = 0; c < CS_CHUNKS; ++c) v[c] = __ldcg(partial + (int64_t)c * N + n);
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CS_CHUNKS; ++c) s += v[c];
  out[n] += s;
}
size_t colsum_scratch_floats(int N) { return (size_t)CS_CHUNKS * N; }
int colsum_accum_bf16(const bf16* G, int64_t M, int N, int ld, float* out, float* scratch, cudaStream_t st) {
  MSQ_REQUIRE(N % 2 == 0 && ld % 2 == 0 && ((uintptr_t)G & 3) == 0, "colsum_accum_bf16: N=%d ld=%d", N, ld);
  if (M == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(colsum_partial_bf16_kernel, dim3(ceil_div(N, 64), CS_CHUNKS), dim3(256), 0, st, G, M, N, ld, scratch));
  MSQ_LAUNCH_CHECK();
  MSQ_CUDA(launch_k(colsum_reduce_kernel, dim3(ceil_div(N, 64)), dim3(64), 0, st, (const float*)scratch, N, out));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// ---------------------------------------------------------------------------------------------------
// activations, forward (training keeps the pre-activation u) and backward:  h = act(u);  du = dh * act'(u)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_grad(float x, int act) {
  switch (act) {
    case ACT_GELU_ERF: {
      const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
      const float pdf = 0.39894228040143267794f * expf(-0.5f * x * x);
      return fmaf(x, pdf, cdf);
    }
    case ACT_QUICK_GELU: {
      const float s = 1.0f / (1.0f + expf(-1.702f * x));
      return s * fmaf(1.702f * x, 1.0f - s, 1.0f);
    }
    case ACT_TANH: { const float t = tanhf(x); return 1.0f - t * t; }
    case ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case ACT_GELU_TANH: {
      const float k = 0.79788456080286535588f, x2 = x * x;
      const float t = tanhf(k * (x + 0.044715f * x * x2));
      return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * k * (1.0f + 3.0f * 0.044715f * x2);
    }
    default: return 1.f;
  }
}
// bf16 path: derivative with one or two SFU ops per element (the same tanh fit of erf as the forward epilogues, common.cuh;
// error ~3e-5, far below the bf16 rounding of du); the fp32 parity mode keeps erff / expf.
__device__ __forceinline__ float act_grad_fast(float x, int act) {
  switch (act) {
    case ACT_GELU_ERF: {
      const float x2 = x * x, xc = fminf(x2, 36.0f);
      const float p = x * fmaf(xc, fmaf(xc, -3.58618502e-04f, 3.70495807e-02f), 7.97459395e-01f);
      const float pdf = 0.39894228040143267794f * ex2_approx(-0.72134752044448170368f * x2);   // exp(-x^2 / 2) / sqrt(2 pi)
      return fmaf(x, pdf, fmaf(0.5f, tanh_approx(p), 0.5f));
    }
    case ACT_QUICK_GELU: {
      const float sg = fmaf(0.5f, tanh_approx(0.851f * x), 0.5f);
      return sg * fmaf(1.702f * x, 1.0f - sg, 1.0f);
    }
    default: return act_grad(x, act);
  }
}
template <typename T> struct ActVec;   // elements per thread and per memory access: 4 fp32 (16 B) / 8 bf16 (16 B)
template <> struct ActVec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float* v) { const float4 t = *reinterpret_cast<const float4*>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
  static __device__ __forceinline__ void store(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
  static __device__ __forceinline__ float fwd(float x, int act) { return apply_act(x, act); }
  static __device__ __forceinline__ float grad(float x, int act) { return act_grad(x, act); }
};
template <> struct ActVec<bf16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const bf16* p, float* v) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      v[2 * i] = __low2float(h); v[2 * i + 1] = __high2float(h);
    }
  }
  static __device__ __forceinline__ void store(bf16* p, const float* v) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
  static __device__ __forceinline__ float fwd(float x, int act) { return apply_act_fast(x, act); }
  static __device__ __forceinline__ float grad(float x, int act) { return act_grad_fast(x, act); }
};
template <typename T>
__global__ void __launch_bounds__(256) act_fwd_kernel(const T* __restrict__ u, int64_t nv, int act, T* __restrict__ h) {
  pdl_sync();
  constexpr int V = ActVec<T>::N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    float v[V];
    ActVec<T>::load(u + i * V, v);
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = ActVec<T>::fwd(v[k], act);
    ActVec<T>::store(h + i * V, v);
  }
}
template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(const T* dh, const T* __restrict__ u, int64_t nv, int act, T* du) {   // du may alias dh
  pdl_sync();
  constexpr int V = ActVec<T>::N;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    float x[V], g[V];
    ActVec<T>::load(u + i * V, x);
    ActVec<T>::load(dh + i * V, g);
#pragma unroll
    for (int k = 0; k < V; ++k) g[k] *= ActVec<T>::grad(x[k], act);
    ActVec<T>::store(du + i * V, g);
  }
}
static inline dim3 ew_grid(int64_t n4) { return dim3((unsigned)min((int64_t)148 * 16, (n4 + 255) / 256)); }
template <typename T> int act_fwd(const T* u, int64_t n, int act, T* h, cudaStream_t st) {
  constexpr int V = (int)(16 / sizeof(T));
  MSQ_REQUIRE(n % V == 0 && ((uintptr_t)u & 15) == 0 && ((uintptr_t)h & 15) == 0, "act_fwd: n %% %d / alignment", V);
  if (n == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(act_fwd_kernel<T>, ew_grid(n / V), dim3(256), 0, st, u, n / V, act, h));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template <typename T> int act_bwd(const T* dh, const T* u, int64_t n, int act, T* du, cudaStream_t st) {
  constexpr int V = (int)(16 / sizeof(T));
  MSQ_REQUIRE(n % V == 0 && ((uintptr_t)u & 15) == 0 && ((uintptr_t)dh & 15) == 0 && ((uintptr_t)du & 15) == 0, "act_bwd: n %% %d / alignment", V);
  if (n == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(act_bwd_kernel<T>, ew_grid(n / V), dim3(256), 0, st, dh, u, n / V, act, du));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int act_fwd<float>(const float*, int64_t, int, float*, cudaStream_t);
template int act_fwd<bf16>(const bf16*, int64_t, int, bf16*, cudaStream_t);
template int act_bwd<float>(const float*, const float*, int64_t, int, float*, cudaStream_t);
template int act_bwd<bf16>(const bf16*, const bf16*, int64_t, int, bf16*, cudaStream_t);

// ---------------------------------------------------------------------------------------------------
// LayerNorm backward.  One warp per row, the row (x and dy) lives in registers:
//   xhat = (x - mean) rstd;  g = dy gamma;  dx = rstd (g - mean(g) - xhat mean(g xhat)) [+ add]
//   dgamma += dy xhat;  dbeta += dy  -> per-lane accumulators over the rows a warp walks, combined per block in a fixed
//   order and written to partial[block][2][H]; ln_partial_reduce adds the blocks in index order (deterministic).
// ---------------------------------------------------------------------------------------------------
constexpr int LNB_WARPS = 4;
constexpr int LNB_MAXV = 8;   // H <= 1024

// v: pre-LN row, d: dy row (in), dx row (out).  ag/ab accumulate dgamma/dbeta for this lane's columns.
// AS: stride of the accumulator arrays (1: registers; 32: a per-warp shared-memory row, float4 index i * 32 + lane)
template <int MV = LNB_MAXV, int AS = 1>
__device__ __forceinline__ void ln_bwd_row(float4* v, float4* d, int nv, int H, int lane, const float* __restrict__ gamma, float eps,
                                           float4* ag, float4* ab) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MV; ++i)
    if (i < nv) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / (float)H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MV; ++i)
    if (i < nv) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
  const float rstd = rsqrtf(warp_sum(q) / (float)H + eps);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < MV; ++i)
    if (i < nv) {
      const float4 gm = *reinterpret_cast<const float4*>(gamma + (i * 32 + lane) * 4);
      v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;          // xhat
      ag[i * AS].x = fmaf(d[i].x, v[i].x, ag[i * AS].x); ag[i * AS].y = fmaf(d[i].y, v[i].y, ag[i * AS].y);
      ag[i * AS].z = fmaf(d[i].z, v[i].z, ag[i * AS].z); ag[i * AS].w = fmaf(d[i].w, v[i].w, ag[i * AS].w);
      ab[i * AS].x += d[i].x; ab[i * AS].y += d[i].y; ab[i * AS].z += d[i].z; ab[i * AS].w += d[i].w;
      d[i].x *= gm.x; d[i].y *= gm.y; d[i].z *= gm.z; d[i].w *= gm.w;          // g
      s1 += (d[i].x + d[i].y) + (d[i].z + d[i].w);
      s2 += (d[i].x * v[i].x + d[i].y * v[i].y) + (d[i].z * v[i].z + d[i].w * v[i].w);
    }
  const float m1 = warp_sum(s1) / (float)H, m2 = warp_sum(s2) / (float)H;
#pragma unroll
  for (int i = 0; i < MV; ++i)
    if (i < nv) {
      d[i].x = rstd * (d[i].x - m1 - v[i].x * m2); d[i].y = rstd * (d[i].y - m1 - v[i].y * m2);
      d[i].z = rstd * (d[i].z - m1 - v[i].z * m2); d[i].w = rstd * (d[i].w - m1 - v[i].w * m2);
    }
}
// combine the warps' dgamma / dbeta accumulators and write this block's partial sums
template <int MV = LNB_MAXV, bool IN_SMEM = false>
__device__ __forceinline__ void ln_bwd_flush(float (*sh)[2 * 128 * LNB_MAXV], const float4* ag, const float4* ab, int nv, int H,
                                             float* __restrict__ partial) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < MV; ++i)
    if (!IN_SMEM && i < nv) {
      const int col = (i * 32 + lane) * 4;
      *reinterpret_cast<float4*>(&sh[warp][col]) = ag[i];
      *reinterpret_cast<float4*>(&sh[warp][H + col]) = ab[i];
    }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * H; c += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < LNB_WARPS; ++w) s += sh[w][c];
    partial[(int64_t)blockIdx.x * 2 * H + c] = s;
  }
}
__device__ __forceinline__ int64_t remap_row_t(int64_t row, int in_group, int out_group, int out_off) {
  return in_group ? (row / in_group) * (int64_t)out_group + out_off + row % in_group : row;
}

// SMACC: the dgamma / dbeta accumulators live in the warp's shared-memory row instead of 2 * MV float4 registers per lane
// (fewer registers -> more resident warps; the kernel is latency-bound on its row loads)
template <typename T, int MV, bool SMACC>
__global__ void __launch_bounds__(LNB_WARPS * 32) ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                const float* add, int64_t rows, int H,   // dx may alias add
                                                                const float* __restrict__ gamma, float eps, float* dx,
                                                                T* __restrict__ dx_t, float* __restrict__ partial, int in_group,
                                                                int out_group, int out_off, Drop drop) {
  pdl_sync();
  __shared__ __align__(16) float sh[LNB_WARPS][2 * 128 * LNB_MAXV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nv = H >> 7;
  constexpr int AS = SMACC ? 32 : 1;
  float4 ag_r[SMACC ? 1 : MV], ab_r[SMACC ? 1 : MV];
  float4* ag = SMACC ? reinterpret_cast<float4*>(&sh[warp][0]) + lane : ag_r;
  float4* ab = SMACC ? reinterpret_cast<float4*>(&sh[warp][H]) + lane : ab_r;
#pragma unroll
  for (int i = 0; i < MV; ++i)
    if (!SMACC || i < nv) ag[i * AS] = ab[i * AS] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t row = (int64_t)blockIdx.x * LNB_WARPS + warp; row < rows; row += (int64_t)gridDim.x * LNB_WARPS) {
    float4 v[MV], d[MV];
    const float* xr = x + row * H;
    const float* dr = dy + remap_row_t(row, in_group, out_group, out_off) * H;
#pragma unroll
    for (int i = 0; i < MV; ++i)
      if (i < nv) { v[i] = Vec4<float>::load(xr + (i * 32 + lane) * 4); d[i] = Vec4<float>::load(dr + (i * 32 + lane) * 4); }
    ln_bwd_row<MV, AS>(v, d, nv, H, lane, gamma, eps, ag, ab);
#pragma unroll
    for (int i = 0; i < MV; ++i)
      if (i < nv) {
        const int col = (i * 32 + lane) * 4;
        if (add) {
          const float4 a = Vec4<float>::load(add + row * H + col);
          d[i].x += a.x; d[i].y += a.y; d[i].z += a.z; d[i].w += a.w;
        }
        if (dx) Vec4<float>::store(dx + row * H + col, d[i]);
        if (dx_t) {
          // under dropout of the dense output this LayerNorm normalised, the operand copy is the gradient of that dense
          // output (dx . mask, element index row * H + col as in the forward); the fp32 dx stays the residual branch's
          if (drop.thresh) {
            const uint64_t e = (uint64_t)row * H + col;
            d[i].x *= drop_mul(drop, e); d[i].y *= drop_mul(drop, e + 1); d[i].z *= drop_mul(drop, e + 2); d[i].w *= drop_mul(drop, e + 3);
          }
          Vec4<T>::store(dx_t + row * H + col, d[i]);
        }
      }
  }
  ln_bwd_flush<MV, SMACC>(sh, ag, ab, nv, H, partial);
}

// block = 32 columns x 32 groups of partial blocks; the groups' sums are combined in index order (deterministic)
constexpr int LNR_GROUPS = 32;
__global__ void __launch_bounds__(LNR_GROUPS * 32) ln_partial_reduce_kernel(const float* __restrict__ partial, int nblk, int H,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_sync();
  __shared__ float sh[LNR_GROUPS][33];
  const int cl = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cl;
  float s = 0.f;
  if (c < 2 * H) {
    const int per = (nblk + LNR_GROUPS - 1) / LNR_GROUPS, b0 = grp * per, b1 = min(nblk, b0 + per);
    int b = b0;
    for (; b + 8 <= b1; b += 8) {   // eight loads in flight per thread, added in index order
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcg(partial + (int64_t)(b + u) * 2 * H + c);
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; b < b1; ++b) s += __ldcg(partial + (int64_t)b * 2 * H + c);
  }
  sh[grp][cl] = s;
  __syncthreads();
  if (grp == 0 && c < 2 * H) {
    float t = 0.f;
    for (int g = 0; g < LNR_GROUPS; ++g) t += sh[g][cl];
    if (c < H) dgamma[c] += t; else dbeta[c - H] += t;
  }
}

constexpr int LNB_GRID = 148 * 6;   // persistent blocks: up to six are resident per SM with the shared-memory accumulators
static inline int lnb_blocks(int64_t rows) { return (int)min((int64_t)LNB_GRID, (rows + LNB_WARPS - 1) / LNB_WARPS); }
size_t ln_bwd_scratch_floats(int H) { return (size_t)LNB_GRID * 2 * H; }

template <typename T>
int ln_bwd(const float* dy, const float* x, const float* add, int64_t rows, int H, const float* gamma, float eps, float* dx, T* dx_t,
           float* dgamma, float* dbeta, float* scratch, int in_group, int out_group, int out_off, cudaStream_t st, const Drop& drop) {
  MSQ_REQUIRE(H % 128 == 0 && H <= 128 * LNB_MAXV, "ln_bwd: H=%d unsupported", H);
  if (rows == 0) return MSQ_OK;
  const int nblk = lnb_blocks(rows);
  // H = 768 instantiation keeps six float4 per array instead of eight (fewer registers -> one more resident block per SM)
  static int smacc = -1;   // MSQ_LNB_SMACC=0: register accumulators
  if (smacc < 0) { const char* e = getenv("MSQ_LNB_SMACC"); smacc = (e && e[0] == '0') ? 0 : 1; }
  if (H <= 768 && smacc) MSQ_CUDA(launch_k(ln_bwd_kernel<T, 6, true>, dim3(nblk), dim3(LNB_WARPS * 32), 0, st, dy, x, add, rows, H, gamma, eps, dx, dx_t, scratch, in_group, out_group, out_off, drop));
  else if (H <= 768) MSQ_CUDA(launch_k(ln_bwd_kernel<T, 6, false>, dim3(nblk), dim3(LNB_WARPS * 32), 0, st, dy, x, add, rows, H, gamma, eps, dx, dx_t, scratch, in_group, out_group, out_off, drop));
  else MSQ_CUDA(launch_k(ln_bwd_kernel<T, LNB_MAXV, false>, dim3(nblk), dim3(LNB_WARPS * 32), 0, st, dy, x, add, rows, H, gamma, eps, dx, dx_t, scratch, in_group, out_group, out_off, drop));
  MSQ_LAUNCH_CHECK();
  MSQ_CUDA(launch_k(ln_partial_reduce_kernel, dim3(ceil_div(2 * H, 32)), dim3(LNR_GROUPS * 32), 0, st, (const float*)scratch, nblk, H, dgamma, dbeta));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int ln_bwd<float>(const float*, const float*, const float*, int64_t, int, const float*, float, float*, float*, float*, float*,
                           float*, int, int, int, cudaStream_t, const Drop&);
template int ln_bwd<bf16>(const float*, const float*, const float*, int64_t, int, const float*, float, float*, bf16*, float*, float*,
                          float*, int, int, int, cudaStream_t, const Drop&);

// ---------------------------------------------------------------------------------------------------
// BertEmbeddings backward (lxrt/modeling.py:356-370): recompute e = word[ids] + pos[t] + type[tt], LayerNorm backward,
// scatter-add de into the three tables (fp32 atomics: a row such as [CLS] collects every pair's gradient).
// dy rows are the text rows of the joint stream: r * Lj + t.  nn.Embedding(padding_idx=0) never accumulates a gradient
// into row 0: pad0 bit 0 / 1 / 2 = word / position / token-type table built with padding_idx=0 (the LXRT embeddings
// set it on all three, lxrt/modeling.py:347-349; the text-only BertModel on the word table only, modeling_bert.py:153).
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_add4(float* p, float4 v) {
  atomicAdd(p, v.x); atomicAdd(p + 1, v.y); atomicAdd(p + 2, v.z); atomicAdd(p + 3, v.w);
}
__global__ void __launch_bounds__(LNB_WARPS * 32) embed_ln_bwd_kernel(const float* __restrict__ dy, const int64_t* __restrict__ ids,
                                                                      const int64_t* __restrict__ tts, int64_t R, int Lt, int Lj, int H,
                                                                      const float* __restrict__ word, const float* __restrict__ pos,
                                                                      const float* __restrict__ type, const float* __restrict__ gamma,
                                                                      float eps, float* __restrict__ dword, float* __restrict__ dpos,
                                                                      float* __restrict__ dtype, float* __restrict__ partial, int pad0) {
  pdl_sync();
  __shared__ __align__(16) float sh[LNB_WARPS][2 * 128 * LNB_MAXV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nv = H >> 7;
  float4 ag[LNB_MAXV], ab[LNB_MAXV];
#pragma unroll
  for (int i = 0; i < LNB_MAXV; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t row = (int64_t)blockIdx.x * LNB_WARPS + warp; row < R * Lt; row += (int64_t)gridDim.x * LNB_WARPS) {
    const int64_t r = row / Lt;
    const int t = (int)(row % Lt);
    const int64_t id = ids[row], tt = tts[row];
    const float *w = word + id * H, *p = pos + (int64_t)t * H, *ty = type + tt * H;
    const float* dr = dy + (r * Lj + t) * H;
    float4 v[LNB_MAXV], d[LNB_MAXV];
#pragma unroll
    for (int i = 0; i < LNB_MAXV; ++i)
      if (i < nv) {
        const int col = (i * 32 + lane) * 4;
        const float4 a = Vec4<float>::load(w + col), b = Vec4<float>::load(p + col), c = Vec4<float>::load(ty + col);
        v[i] = make_float4((a.x + b.x) + c.x, (a.y + b.y) + c.y, (a.z + b.z) + c.z, (a.w + b.w) + c.w);
        d[i] = Vec4<float>::load(dr + col);
      }
    ln_bwd_row(v, d, nv, H, lane, gamma, eps, ag, ab);
#pragma unroll
    for (int i = 0; i < LNB_MAXV; ++i)
      if (i < nv) {
        const int col = (i * 32 + lane) * 4;
        if (id != 0 || !(pad0 & 1)) atomic_add4(dword + id * H + col, d[i]);
        if (t != 0 || !(pad0 & 2)) atomic_add4(dpos + (int64_t)t * H + col, d[i]);
        if (tt != 0 || !(pad0 & 4)) atomic_add4(dtype + tt * H + col, d[i]);
      }
  }
  ln_bwd_flush(sh, ag, ab, nv, H, partial);
}
int embed_ln_bwd(const float* dy, const int64_t* ids, const int64_t* tts, int64_t R, int Lt, int Lj, int H, const float* word,
                 const float* pos, const float* type, const float* gamma, float eps, float* dword, float* dpos, float* dtype,
                 float* dgamma, float* dbeta, float* scratch, int pad0, cudaStream_t st) {
  MSQ_REQUIRE(H % 128 == 0 && H <= 128 * LNB_MAXV, "embed_ln_bwd: H=%d unsupported", H);
  if (R == 0) return MSQ_OK;
  const int nblk = lnb_blocks(R * Lt);
  MSQ_CUDA(launch_k(embed_ln_bwd_kernel, dim3(nblk), dim3(LNB_WARPS * 32), 0, st, dy, ids, tts, R, Lt, Lj, H, word, pos, type, gamma, eps, dword, dpos, dtype, scratch, pad0));
  MSQ_LAUNCH_CHECK();
  MSQ_CUDA(launch_k(ln_partial_reduce_kernel, dim3(ceil_div(2 * H, 32)), dim3(LNR_GROUPS * 32), 0, st, (const float*)scratch, nblk, H, dgamma, dbeta));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// ---------------------------------------------------------------------------------------------------
// ViT token assembly + ln_pre backward (clip/model.py:263-276): recompute v = src + pos, LayerNorm backward, scatter dv
// into the class embedding, the positional table (pair-joint row rule of vit_assemble_kernel) and the per-image patch
// embedding gradient dpatch [n_img*g2, W] (an image that appears in several pairs collects all of them).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LNB_WARPS * 32) vit_assemble_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ patch,
                                                                          const int32_t* __restrict__ img_index, int64_t R, int il, int g2,
                                                                          int W, const float* __restrict__ cls, const float* __restrict__ pos,
                                                                          const float* __restrict__ gamma, float eps, float* __restrict__ dpatch,
                                                                          float* __restrict__ dcls, float* __restrict__ dpos,
                                                                          float* __restrict__ partial) {
  pdl_sync();
  __shared__ __align__(16) float sh[LNB_WARPS][2 * 128 * LNB_MAXV];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nv = W >> 7;
  const int Lv = 1 + il * g2;
  float4 ag[LNB_MAXV], ab[LNB_MAXV];
#pragma unroll
  for (int i = 0; i < LNB_MAXV; ++i) ag[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t row = (int64_t)blockIdx.x * LNB_WARPS + warp; row < R * Lv; row += (int64_t)gridDim.x * LNB_WARPS) {
    const int64_t r = row / Lv;
    const int t = (int)(row % Lv);
    const float* src;
    float* dsrc;
    int prow;
    if (t == 0) {
      src = cls; dsrc = dcls; prow = 0;
    } else {
      const int slot = (t - 1) / g2, p = (t - 1) % g2;
      const int64_t pr = ((int64_t)img_index[r * il + slot] * g2 + p) * W;
      src = patch + pr; dsrc = dpatch + pr;
      prow = slot == 0 ? t : p;
    }
    const float* pp = pos + (int64_t)prow * W;
    float4 v[LNB_MAXV], d[LNB_MAXV];
#pragma unroll
    for (int i = 0; i < LNB_MAXV; ++i)
      if (i < nv) {
        const int col = (i * 32 + lane) * 4;
        const float4 a = Vec4<float>::load(src + col), b = Vec4<float>::load(pp + col);
        v[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
        d[i] = Vec4<float>::load(dy + row * W + col);
      }
    ln_bwd_row(v, d, nv, W, lane, gamma, eps, ag, ab);
#pragma unroll
    for (int i = 0; i < LNB_MAXV; ++i)
      if (i < nv) {
        const int col = (i * 32 + lane) * 4;
        atomic_add4(dsrc + col, d[i]);
        atomic_add4(dpos + (int64_t)prow * W + col, d[i]);
      }
  }
  ln_bwd_flush(sh, ag, ab, nv, W, partial);
}
int vit_assemble_bwd(const float* dy, const float* patch, const int32_t* img_index, int64_t R, int il, int g2, int W, const float* cls,
                     const float* pos, const float* gamma, float eps, float* dpatch, float* dcls, float* dpos, float* dgamma,
                     float* dbeta, float* scratch, cudaStream_t st) {
  MSQ_REQUIRE(W % 128 == 0 && W <= 128 * LNB_MAXV, "vit_assemble_bwd: width=%d unsupported", W);
  if (R == 0) return MSQ_OK;
  const int nblk = lnb_blocks(R * (1 + il * g2));
  MSQ_CUDA(launch_k(vit_assemble_bwd_kernel, dim3(nblk), dim3(LNB_WARPS * 32), 0, st, dy, patch, img_index, R, il, g2, W, cls, pos, gamma, eps, dpatch, dcls, dpos, scratch));
  MSQ_LAUNCH_CHECK();
  MSQ_CUDA(launch_k(ln_partial_reduce_kernel, dim3(ceil_div(2 * W, 32)), dim3(LNR_GROUPS * 32), 0, st, (const float*)scratch, nblk, W, dgamma, dbeta));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// ---------------------------------------------------------------------------------------------------
// Attention backward, fp32 on CUDA cores (bf16 or fp32 I/O), same data layout as attention.cu:
//   P = softmax(scale Q K^T + mask);  O = P V
//   dP = dO V^T;  D_i = sum_j P_ij dP_ij;  dS = P (dP - D);  dQ = scale dS K;  dK = scale dS^T Q;  dV = P^T dO
// Kernel 1 (one CTA per (group, head), K and V in shared memory, a warp per query row): recomputes the row's softmax,
// writes dQ, and the row's log-sum-exp and D to scratch.  Kernel 2 (Q and dO in shared memory, a warp per key row):
// re-derives P and dS columns from (lse, D) and writes dK, dV.  Nothing of size L x L ever goes to HBM.
// ---------------------------------------------------------------------------------------------------
constexpr int AB_D = 64;
constexpr int AB_WARPS = 16;
constexpr int AB_MAXCH = 10;  // L <= 320

template <typename T>
__global__ void __launch_bounds__(AB_WARPS * 32) attention_bwd_dq_kernel(const T* __restrict__ qkv, const T* __restrict__ dctx, int L,
                                                                         int heads, float scale, const float* __restrict__ mask_add,
                                                                         int mask_ld, int mask_len, T* __restrict__ dqkv,
                                                                         float* __restrict__ lse_out, float* __restrict__ dsum_out, Drop drop) {
  pdl_sync();
  extern __shared__ __align__(16) float sm[];
  const int r = blockIdx.x / heads, h = blockIdx.x % heads;
  const int ld = 3 * heads * AB_D, ldc = heads * AB_D;
  const int Lpad = (L + 31) & ~31;
  float* Ks = sm;                         // [L][65]
  float* Vs = Ks + L * 65;                // [L][65]
  float* Ms = Vs + L * 65;                // [Lpad]
  float* Qs = Ms + Lpad;                  // [warps][64]
  float* Gs = Qs + AB_WARPS * AB_D;       // [warps][64]
  float* Ps = Gs + AB_WARPS * AB_D;       // [warps][Lpad]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* base = qkv + (int64_t)r * L * ld + h * AB_D;
  for (int i = threadIdx.x; i < L * AB_D; i += blockDim.x) {
    const int t = i / AB_D, d = i % AB_D;
    const T* kp = base + (int64_t)t * ld + heads * AB_D + d;
    Ks[t * 65 + d] = to_f(kp[0]);
    Vs[t * 65 + d] = to_f(kp[heads * AB_D]);
  }
  for (int t = threadIdx.x; t < Lpad; t += blockDim.x)
    Ms[t] = (mask_add && t < mask_len && t < L) ? mask_add[(int64_t)r * mask_ld + t] : 0.f;
  __syncthreads();
  float* q = Qs + warp * AB_D;
  float* g = Gs + warp * AB_D;
  float* p = Ps + warp * Lpad;
  const int nch = Lpad / 32;
  for (int t = warp; t < L; t += AB_WARPS) {
    const T* qp = base + (int64_t)t * ld;
    const T* gp = dctx + ((int64_t)r * L + t) * ldc + h * AB_D;
    q[lane] = to_f(qp[lane]); q[lane + 32] = to_f(qp[lane + 32]);
    g[lane] = to_f(gp[lane]); g[lane + 32] = to_f(gp[lane + 32]);
    __syncwarp();
    float dpv[AB_MAXCH];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < AB_MAXCH; ++c) {
      dpv[c] = 0.f;
      if (c < nch) {
        const int key = c * 32 + lane;
        float s = -INFINITY;
        if (key < L) {
          const float *kr = Ks + key * 65, *vr = Vs + key * 65;
          float a = 0.f, b = 0.f;
#pragma unroll 16
          for (int d = 0; d < AB_D; ++d) { a = fmaf(q[d], kr[d], a); b = fmaf(g[d], vr[d], b); }
          s = a * scale + Ms[key];
          dpv[c] = b * drop_mul(drop, ((uint64_t)blockIdx.x * L + t) * L + key);   // dP = (dO V^T) . mask
        }
        p[key] = s;
        mx = fmaxf(mx, s);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = 0; c < nch; ++c) {
      const int key = c * 32 + lane;
      if (key < L) sum += expf(p[key] - mx);
    }
    sum = warp_sum(sum);
    const float lse = mx + logf(sum);
    float dsum = 0.f;
#pragma unroll
    for (int c = 0; c < AB_MAXCH; ++c)
      if (c < nch) {
        const int key = c * 32 + lane;
        const float pr = key < L ? expf(p[key] - lse) : 0.f;
        p[key] = pr;
        dsum = fmaf(pr, dpv[c], dsum);
      }
    dsum = warp_sum(dsum);
#pragma unroll
    for (int c = 0; c < AB_MAXCH; ++c)
      if (c < nch) {
        const int key = c * 32 + lane;
        p[key] = p[key] * (dpv[c] - dsum);   // dS
      }
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int key = 0; key < L; ++key) {
      const float ds = p[key];
      o0 = fmaf(ds, Ks[key * 65 + lane], o0);
      o1 = fmaf(ds, Ks[key * 65 + lane + 32], o1);
    }
    T* op = dqkv + ((int64_t)r * L + t) * ld + h * AB_D;
    op[lane] = from_f<T>(o0 * scale);
    op[lane + 32] = from_f<T>(o1 * scale);
    if (lane == 0) {
      lse_out[(int64_t)blockIdx.x * L + t] = lse;
      dsum_out[(int64_t)blockIdx.x * L + t] = dsum;
    }
    __syncwarp();
  }
}

template <typename T>
__global__ void __launch_bounds__(AB_WARPS * 32) attention_bwd_dkv_kernel(const T* __restrict__ qkv, const T* __restrict__ dctx, int L,
                                                                          int heads, float scale, const float* __restrict__ mask_add,
                                                                          int mask_ld, int mask_len, T* __restrict__ dqkv,
                                                                          const float* __restrict__ lse_in, const float* __restrict__ dsum_in, Drop drop) {
  pdl_sync();
  extern __shared__ __align__(16) float sm[];
  const int r = blockIdx.x / heads, h = blockIdx.x % heads;
  const int ld = 3 * heads * AB_D, ldc = heads * AB_D;
  const int Lpad = (L + 31) & ~31;
  float* Qs = sm;                         // [L][65]
  float* Gs = Qs + L * 65;                // [L][65]  dO
  float* Ls = Gs + L * 65;                // [Lpad] lse
  float* Ds = Ls + Lpad;                  // [Lpad] D
  float* Kq = Ds + Lpad;                  // [warps][64] key row
  float* Vq = Kq + AB_WARPS * AB_D;       // [warps][64] value row
  float* Ps = Vq + AB_WARPS * AB_D;       // [warps][Lpad]
  float* Ss = Ps + AB_WARPS * Lpad;       // [warps][Lpad]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const T* base = qkv + (int64_t)r * L * ld + h * AB_D;
  for (int i = threadIdx.x; i < L * AB_D; i += blockDim.x) {
    const int t = i / AB_D, d = i % AB_D;
    Qs[t * 65 + d] = to_f(base[(int64_t)t * ld + d]);
    Gs[t * 65 + d] = to_f(dctx[((int64_t)r * L + t) * ldc + h * AB_D + d]);
  }
  for (int t = threadIdx.x; t < Lpad; t += blockDim.x) {
    Ls[t] = t < L ? lse_in[(int64_t)blockIdx.x * L + t] : 0.f;
    Ds[t] = t < L ? dsum_in[(int64_t)blockIdx.x * L + t] : 0.f;
  }
  __syncthreads();
  float* k = Kq + warp * AB_D;
  float* v = Vq + warp * AB_D;
  float* p = Ps + warp * Lpad;
  float* s = Ss + warp * Lpad;
  const int nch = Lpad / 32;
  for (int j = warp; j < L; j += AB_WARPS) {
    const T* kp = base + (int64_t)j * ld + heads * AB_D;
    k[lane] = to_f(kp[lane]); k[lane + 32] = to_f(kp[lane + 32]);
    v[lane] = to_f(kp[heads * AB_D + lane]); v[lane + 32] = to_f(kp[heads * AB_D + lane + 32]);
    const float mj = (mask_add && j < mask_len) ? mask_add[(int64_t)r * mask_ld + j] : 0.f;
    __syncwarp();
    for (int c = 0; c < nch; ++c) {
      const int i = c * 32 + lane;
      float pr = 0.f, ds = 0.f;
      if (i < L) {
        const float *qr = Qs + i * 65, *gr = Gs + i * 65;
        float a = 0.f, b = 0.f;
#pragma unroll 16
        for (int d = 0; d < AB_D; ++d) { a = fmaf(qr[d], k[d], a); b = fmaf(gr[d], v[d], b); }
        pr = expf(a * scale + mj - Ls[i]);
        const float dm = drop_mul(drop, ((uint64_t)blockIdx.x * L + i) * L + j);
        ds = pr * (b * dm - Ds[i]);
        pr *= dm;                                   // dV = (P . mask)^T dO
      }
      p[i] = pr;
      s[i] = ds;
    }
    __syncwarp();
    float dv0 = 0.f, dv1 = 0.f, dk0 = 0.f, dk1 = 0.f;
    for (int i = 0; i < L; ++i) {
      const float pr = p[i], ds = s[i];
      dv0 = fmaf(pr, Gs[i * 65 + lane], dv0);
      dv1 = fmaf(pr, Gs[i * 65 + lane + 32], dv1);
      dk0 = fmaf(ds, Qs[i * 65 + lane], dk0);
      dk1 = fmaf(ds, Qs[i * 65 + lane + 32], dk1);
    }
    T* op = dqkv + ((int64_t)r * L + j) * ld + heads * AB_D + h * AB_D;
    op[lane] = from_f<T>(dk0 * scale);
    op[lane + 32] = from_f<T>(dk1 * scale);
    op[heads * AB_D + lane] = from_f<T>(dv0);
    op[heads * AB_D + lane + 32] = from_f<T>(dv1);
    __syncwarp();
  }
}

size_t attention_bwd_scratch_floats(int64_t R, int L, int heads) { return (size_t)2 * R * heads * L; }

template <typename T>
int attention_bwd(const T* qkv, const T* dctx, int64_t R, int L, int heads, float scale, const float* key_mask_add, int mask_ld,
                  int mask_len, T* dqkv, float* scratch, cudaStream_t st, const Drop& drop) {
  MSQ_REQUIRE(L >= 1 && L <= 32 * AB_MAXCH, "attention_bwd: sequence length %d out of range", L);
  if (R == 0) return MSQ_OK;
  const int Lpad = (L + 31) & ~31;
  float* lse = scratch;
  float* dsum = scratch + (size_t)R * heads * L;
  const size_t smem1 = sizeof(float) * ((size_t)2 * L * 65 + Lpad + 2 * AB_WARPS * AB_D + (size_t)AB_WARPS * Lpad);
  const size_t smem2 = sizeof(float) * ((size_t)2 * L * 65 + 2 * Lpad + 2 * AB_WARPS * AB_D + (size_t)2 * AB_WARPS * Lpad);
  MSQ_SMEM_ATTR(smem1, attention_bwd_dq_kernel<T>);
  MSQ_SMEM_ATTR(smem2, attention_bwd_dkv_kernel<T>);
  MSQ_CUDA(launch_k(attention_bwd_dq_kernel<T>, dim3((unsigned)(R * heads)), dim3(AB_WARPS * 32), smem1, st, qkv, dctx, L, heads, scale, key_mask_add, mask_ld, mask_len, dqkv, lse, dsum, drop));
  MSQ_LAUNCH_CHECK();
  MSQ_CUDA(launch_k(attention_bwd_dkv_kernel<T>, dim3((unsigned)(R * heads)), dim3(AB_WARPS * 32), smem2, st, qkv, dctx, L, heads, scale, key_mask_add, mask_ld, mask_len, dqkv, (const float*)lse, (const float*)dsum, drop));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int attention_bwd<float>(const float*, const float*, int64_t, int, int, float, const float*, int, int, float*, float*, cudaStream_t, const Drop&);
template int attention_bwd<bf16>(const bf16*, const bf16*, int64_t, int, int, float, const float*, int, int, bf16*, float*, cudaStream_t, const Drop&);

// dW[i] += sum_s part[s, i]   (split-K partial sums of a weight gradient, added in slice order)
__global__ void __launch_bounds__(256) splitk_accumulate_kernel(const float* __restrict__ part, int S, int64_t n4, float* __restrict__ dW) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = *reinterpret_cast<const float4*>(dW + i * 4);
    for (int s = 0; s < S; ++s) {
      const float4 p = *reinterpret_cast<const float4*>(part + ((int64_t)s * n4 + i) * 4);
      a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w;
    }
    *reinterpret_cast<float4*>(dW + i * 4) = a;
  }
}
int splitk_accumulate(const float* part, int S, int64_t n, float* dW, cudaStream_t st) {
  MSQ_REQUIRE(n % 4 == 0, "splitk_accumulate: n %% 4");
  MSQ_CUDA(launch_k(splitk_accumulate_kernel, ew_grid(n / 4), dim3(256), 0, st, part, S, n / 4, dW));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// additive key mask (1 - mask) * -10000 (lxrt/modeling.py:1537-1545)
__global__ void __launch_bounds__(256) mask_add_train_kernel(const int64_t* __restrict__ mask, int64_t n, float* __restrict__ out) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (1.0f - (float)mask[i]) * -10000.0f;
}
int mask_add_from_int(const int64_t* mask, int64_t n, float* out, cudaStream_t st) {
  if (n == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(mask_add_train_kernel, dim3((unsigned)min((int64_t)148 * 8, (n + 255) / 256)), dim3(256), 0, st, mask, n, out));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// ---------------------------------------------------------------------------------------------------
// misc: dst[r, c] (row group remap as gather_rows, reversed) -- scatter the gradients of the text / visual outputs into
// the joint stream's gradient:  dst[(r / group) * dst_group + off + r % group, :] = src[r, :]
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scatter_rows_kernel(const float* __restrict__ src, int64_t rows, int H, int group, int dst_group,
                                                           int off, float* __restrict__ dst) {
  pdl_sync();
  const int64_t total4 = rows * (H / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (H / 4);
    const int c = (int)(i % (H / 4)) * 4;
    const int64_t dr = group ? (r / group) * (int64_t)dst_group + off + r % group : r;
    Vec4<float>::store(dst + dr * H + c, Vec4<float>::load(src + r * H + c));
  }
}
int scatter_rows(const float* src, int64_t rows, int H, int group, int dst_group, int off, float* dst, cudaStream_t st) {
  MSQ_REQUIRE(H % 4 == 0, "scatter_rows: H=%d", H);
  if (rows == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(scatter_rows_kernel, ew_grid(rows * (H / 4)), dim3(256), 0, st, src, rows, H, group, dst_group, off, dst));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// ---------------------------------------------------------------------------------------------------
// Optimizer (trainers/train.py:172-190, 353-363): clip_grad_norm_(max_grad_norm) over ALL gradients, then HF AdamW
// (transformers 3.4 optimization.py, correct_bias=True):
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= lr sqrt(1-b2^t)/(1-b1^t) m / (sqrt(v) + eps);  p -= lr wd p
// The squared norm is summed in two ordered passes (deterministic); the clip coefficient stays on the device.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ partial) {
  pdl_sync();
  __shared__ float sh[8];
  float s = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s = fmaf(g[i], g[i], s);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += sh[i];
    partial[blockIdx.x] = t;
  }
}
// out[0] = total norm, out[1] = clip coefficient min(1, max_norm / (norm + 1e-6)) (torch.nn.utils.clip_grad_norm_), times grad_scale
__global__ void clip_coef_kernel(const float* __restrict__ partial, int n, float max_norm, float grad_scale, float* __restrict__ out) {
  pdl_sync();
  if (threadIdx.x != 0) return;
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += (double)partial[i];
  const float norm = sqrtf((float)s) * grad_scale;
  float coef = 1.f;
  if (max_norm > 0.f) coef = fminf(1.f, max_norm / (norm + 1e-6f));
  out[0] = norm;
  out[1] = coef * grad_scale;
}
constexpr int SUMSQ_BLOCKS = 148 * 4;
size_t grad_norm_scratch_floats() { return SUMSQ_BLOCKS + 2; }
int grad_norm_clip(const float* g, int64_t n, float max_norm, float grad_scale, float* scratch, cudaStream_t st) {
  MSQ_CUDA(launch_k(sumsq_partial_kernel, dim3(SUMSQ_BLOCKS), dim3(256), 0, st, g, n, scratch + 2));
  MSQ_LAUNCH_CHECK();
  MSQ_CUDA(launch_k(clip_coef_kernel, dim3(1), dim3(32), 0, st, (const float*)(scratch + 2), SUMSQ_BLOCKS, max_norm, grad_scale, scratch));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                    float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                                                    float step_size, const float* __restrict__ coef) {
  pdl_sync();
  const float c = coef ? coef[1] : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * c;
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float pi = p[i] - step_size * (mi / (sqrtf(vi) + eps));
    if (wd != 0.f) pi -= lr * wd * pi;
    p[i] = pi;
  }
}
// all parameters in ONE launch: block b owns chunk b = (master pointer, offset into the flat buffers, length, decays?)
__global__ void __launch_bounds__(256) adamw_multi_kernel(const AdamChunk* __restrict__ chunks, const float* __restrict__ g, float* __restrict__ m,
                                                          float* __restrict__ v, float lr, float b1, float b2, float eps, float wd, float step_size,
                                                          const float* __restrict__ coef) {
  pdl_sync();
  const AdamChunk ch = chunks[blockIdx.x];
  const float c = coef ? coef[1] : 1.f;
  const float w = ch.decay ? wd : 0.f;
  for (int i = threadIdx.x; i < ch.n; i += blockDim.x) {
    const int64_t o = ch.off + i;
    const float gi = g[o] * c;
    const float mi = b1 * m[o] + (1.f - b1) * gi;
    const float vi = b2 * v[o] + (1.f - b2) * gi * gi;
    m[o] = mi; v[o] = vi;
    float pi = ch.p[i] - step_size * (mi / (sqrtf(vi) + eps));
    if (w != 0.f) pi -= lr * w * pi;
    ch.p[i] = pi;
  }
}
int adamw_update_multi(const AdamChunk* chunks_dev, int nchunks, const float* g, float* m, float* v, float lr, float b1, float b2, float eps, float wd,
                       int64_t step, const float* coef, cudaStream_t st) {
  if (nchunks == 0) return MSQ_OK;
  const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
  const float step_size = (float)((double)lr * sqrt(bc2) / bc1);
  MSQ_CUDA(launch_k(adamw_multi_kernel, dim3(nchunks), dim3(256), 0, st, chunks_dev, g, m, v, lr, b1, b2, eps, wd, step_size, coef));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

int adamw_update(float* p, const float* g, float* m, float* v, int64_t n, float lr, float b1, float b2, float eps, float wd, int64_t step,
                 const float* coef, cudaStream_t st) {
  if (n == 0) return MSQ_OK;
  const double bc1 = 1.0 - pow((double)b1, (double)step), bc2 = 1.0 - pow((double)b2, (double)step);
  const float step_size = (float)((double)lr * sqrt(bc2) / bc1);
  MSQ_CUDA(launch_k(adamw_kernel, dim3((unsigned)min((int64_t)148 * 8, (n + 255) / 256)), dim3(256), 0, st, p, g, m, v, n, lr, b1, b2, eps, wd, step_size, coef));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}


// ---- dropout (dropout.cuh): elementwise sites.  Masks are regenerated from (key, element index), never stored. ------------
// rows [off, off + n) of every group of `group` rows of a [*, H] buffer; the mask index runs over the [R, n, H] tensor the
// reference's nn.Dropout sees.  In place on the fp32 buffer; xt (optional) receives the operand-type copy.
template <typename T>
__global__ void __launch_bounds__(256) dropout_rows_kernel(float* __restrict__ xf, T* __restrict__ xt, int64_t R, int group, int off, int n,
                                                           int H, Drop d) {
  pdl_sync();
  const int64_t total4 = R * n * (int64_t)(H / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / (H / 4);
    const int c = (int)(i % (H / 4)) * 4;
    const int64_t r = row / n;
    const int t = (int)(row % n);
    const int64_t brow = r * group + off + t;
    float4 v = *reinterpret_cast<const float4*>(xf + brow * H + c);
    const uint64_t idx = (uint64_t)row * H + c;
    v.x *= drop_mul(d, idx); v.y *= drop_mul(d, idx + 1); v.z *= drop_mul(d, idx + 2); v.w *= drop_mul(d, idx + 3);
    *reinterpret_cast<float4*>(xf + brow * H + c) = v;
    if (xt) Vec4<T>::store(xt + brow * H + c, v);
  }
}
template <typename T>
int dropout_rows(float* xf, T* xt, int64_t R, int group, int off, int n, int H, const Drop& d, cudaStream_t st) {
  if (d.thresh == 0 || R * n == 0) return MSQ_OK;
  const int64_t total4 = R * n * (int64_t)(H / 4);
  MSQ_CUDA(launch_k(dropout_rows_kernel<T>, dim3((unsigned)min((int64_t)148 * 16, (total4 + 255) / 256)), dim3(256), 0, st, xf, xt, R, group, off, n, H, d));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int dropout_rows<float>(float*, float*, int64_t, int, int, int, int, const Drop&, cudaStream_t);
template int dropout_rows<bf16>(float*, bf16*, int64_t, int, int, int, int, const Drop&, cudaStream_t);

// s <- dropout(s) + resid   (s = dense(x) + bias of a residual sub-layer; lxrt/modeling.py:436-438, 490-492)
__global__ void __launch_bounds__(256) dropout_add_kernel(float* __restrict__ s, const float* __restrict__ resid, int64_t n4, Drop d) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(s)[i];
    const float4 r = resid ? reinterpret_cast<const float4*>(resid)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    const uint64_t idx = (uint64_t)i * 4;
    v.x = fmaf(v.x, drop_mul(d, idx), r.x); v.y = fmaf(v.y, drop_mul(d, idx + 1), r.y);
    v.z = fmaf(v.z, drop_mul(d, idx + 2), r.z); v.w = fmaf(v.w, drop_mul(d, idx + 3), r.w);
    reinterpret_cast<float4*>(s)[i] = v;
  }
}
int dropout_add(float* s, const float* resid, int64_t n, const Drop& d, cudaStream_t st) {
  if (n == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(dropout_add_kernel, dim3((unsigned)min((int64_t)148 * 16, (n / 4 + 255) / 256)), dim3(256), 0, st, s, resid, n / 4, d));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// gt <- T(g * mask): the gradient of the dense output under dropout (the fp32 g stays the residual-branch gradient)
template <typename T>
__global__ void __launch_bounds__(256) dropout_mask_copy_kernel(const float* __restrict__ g, T* __restrict__ gt, int64_t n4, Drop d) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(g)[i];
    const uint64_t idx = (uint64_t)i * 4;
    v.x *= drop_mul(d, idx); v.y *= drop_mul(d, idx + 1); v.z *= drop_mul(d, idx + 2); v.w *= drop_mul(d, idx + 3);
    Vec4<T>::store(gt + i * 4, v);
  }
}
template <typename T>
int dropout_mask_copy(const float* g, T* gt, int64_t n, const Drop& d, cudaStream_t st) {
  if (d.thresh == 0 || n == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(dropout_mask_copy_kernel<T>, dim3((unsigned)min((int64_t)148 * 16, (n / 4 + 255) / 256)), dim3(256), 0, st, g, gt, n / 4, d));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int dropout_mask_copy<float>(const float*, float*, int64_t, const Drop&, cudaStream_t);
template int dropout_mask_copy<bf16>(const float*, bf16*, int64_t, const Drop&, cudaStream_t);

}  // namespace msq
