// Shared device/host helpers for the sm_100a kernels of the step-ordering path.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define MSQ_OK 0
#define MSQ_ERR_ARG 1
#define MSQ_ERR_CUDA 2
#define MSQ_ERR_WEIGHT 3
#define MSQ_ERR_STATE 4

namespace msq {

typedef __nv_bfloat16 bf16;

// "Split bf16" operand type of the bf16x3 precision mode (msq_config.precise == 2): a value x is carried as
// hi = bf16(x) and lo = bf16(x - hi), i.e. 16 significand bits, and a product a*w is formed on the tensor cores as
// a_hi*w_hi + a_lo*w_hi + a_hi*w_lo with fp32 accumulation (the dropped lo*lo term is 2^-18 relative).
// Storage: a row of K logical elements is [hi(K) | lo(K)] bf16 = K * 4 bytes, so sizeof(bf16s) == 4 keeps all row
// arithmetic (r * K, buffer sizes) in units of T; only kernels look inside a row.
struct bf16s { bf16 hi, lo; };
template <typename T> struct is_split { static constexpr bool value = false; };
template <> struct is_split<bf16s> { static constexpr bool value = true; };
template <typename A, typename B> struct same_type { static constexpr bool value = false; };
template <typename A> struct same_type<A, A> { static constexpr bool value = true; };

void set_error(const char* fmt, ...);

#define MSQ_CUDA(...)                                                                              \
  do {                                                                                             \
    cudaError_t e__ = (__VA_ARGS__);                                                               \
    if (e__ != cudaSuccess) {                                                                      \
      msq::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #__VA_ARGS__, cudaGetErrorString(e__)); \
      return MSQ_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

#define MSQ_LAUNCH_CHECK()                                                                       \
  do {                                                                                           \
    cudaError_t e__ = cudaGetLastError();                                                        \
    if (e__ != cudaSuccess) {                                                                    \
      msq::set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return MSQ_ERR_CUDA;                                                                       \
    }                                                                                            \
    msq::count_launch();                                                                         \
  } while (0)

#define MSQ_REQUIRE(cond, ...)       \
  do {                               \
    if (!(cond)) {                   \
      msq::set_error(__VA_ARGS__);   \
      return MSQ_ERR_ARG;            \
    }                                \
  } while (0)

#define MSQ_TRY(expr)              \
  do {                             \
    int rc__ = (expr);             \
    if (rc__ != MSQ_OK) return rc__; \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute: remember per device what has been requested
// (a process may hold models on several GPUs).  Usage: MSQ_SMEM_ATTR(bytes, kernel<template, args>);
#define MSQ_SMEM_ATTR(bytes, ...)                                                                              \
  do {                                                                                                         \
    static size_t cfg__[64] = {0};                                                                             \
    int dev__ = 0;                                                                                             \
    MSQ_CUDA(cudaGetDevice(&dev__));                                                                           \
    if ((size_t)(bytes) > cfg__[dev__ & 63]) {                                                                 \
      MSQ_CUDA(cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      cfg__[dev__ & 63] = (size_t)(bytes);                                                                     \
    }                                                                                                          \
  } while (0)

void count_launch();
int pdl_enabled();

// Programmatic dependent launch: every kernel of this library starts with pdl_sync() (wait for the
// preceding grid to complete and flush, then allow the NEXT kernel to be scheduled), and every launch
// carries the programmatic-stream-serialization attribute, so launch latency and block ramp-up of
// kernel k+1 overlap the tail of kernel k instead of adding ~3 us of idle GPU per launch.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled();
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// activation selectors for GEMM epilogues
enum Act { ACT_NONE = 0, ACT_GELU_ERF = 1, ACT_QUICK_GELU = 2, ACT_TANH = 3, ACT_GELU_TANH = 4, ACT_RELU = 5 };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float apply_act(float x, int act) {
  switch (act) {
    case ACT_GELU_ERF: return x * 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    case ACT_QUICK_GELU: return x / (1.0f + __expf(-1.702f * x));
    case ACT_TANH: return tanhf(x);
    case ACT_RELU: return fmaxf(x, 0.f);
    case ACT_GELU_TANH: {
      float u = 0.79788456080286535588f * (x + 0.044715f * x * x * x);
      return 0.5f * x * (1.0f + tanhf(u));
    }
    default: return x;
  }
}

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ float tanh_approx(float x) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Cheap variants for the bf16 tensor-core epilogues (results are rounded to bf16 or feed a bf16 GEMM); ONE
// SFU op (MUFU.TANH, rel. error 2^-11) per element so the epilogue stays below the MMA time of a tile:
//   erf-GELU   0.5 x (1 + tanh(x (c0 + c1 x^2 + c2 x^4)))  minimax fit of atanh(erf(x/sqrt2)), |err| < 3e-5
//   QuickGELU  x sigmoid(1.702 x) = 0.5 x (1 + tanh(0.851 x))   (exact identity)
// Both errors are 10-100x below the bf16 rounding of the stored value.  The fp32 parity mode and the
// decoder use erff / expf / tanhf (apply_act).
__device__ __forceinline__ float apply_act_fast(float x, int act) {
  switch (act) {
    case ACT_GELU_ERF: {
      // the fit holds on [-8,8]; beyond |x| = 6 the polynomial is frozen at its (positive) value there, so p keeps
      // growing linearly and tanh stays saturated at +-1: one clamp instead of two
      const float x2 = fminf(x * x, 36.0f);
      const float p = x * fmaf(x2, fmaf(x2, -3.58618502e-04f, 3.70495807e-02f), 7.97459395e-01f);
      const float hx = 0.5f * x;
      return fmaf(hx, tanh_approx(p), hx);
    }
    case ACT_QUICK_GELU: {
      const float hx = 0.5f * x;
      return fmaf(hx, tanh_approx(0.851f * x), hx);
    }
    case ACT_TANH: return tanh_approx(x);
    case ACT_RELU: return fmaxf(x, 0.f);
    default: return apply_act(x, act);
  }
}

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// 4-element vector access in fp32 / bf16 (128-bit / 64-bit)
template <typename T> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ float4 load(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void store(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Vec4<bf16> {
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

// hi / lo parts of four values, packed as bf16 pairs
__device__ __forceinline__ void split4(const float4& v, uint2& hi, uint2& lo) {
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
  const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __low2float(h0), v.y - __high2float(h0));
  const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __low2float(h1), v.w - __high2float(h1));
  hi.x = *reinterpret_cast<const uint32_t*>(&h0); hi.y = *reinterpret_cast<const uint32_t*>(&h1);
  lo.x = *reinterpret_cast<const uint32_t*>(&l0); lo.y = *reinterpret_cast<const uint32_t*>(&l1);
}
// store / load four consecutive logical elements at column `col` of a row of K logical elements
template <typename T> __device__ __forceinline__ void store_row4(T* row, int col, int K, const float4& v) {
  (void)K;
  Vec4<T>::store(row + col, v);
}
template <> __device__ __forceinline__ void store_row4<bf16s>(bf16s* row, int col, int K, const float4& v) {
  bf16* p = reinterpret_cast<bf16*>(row);
  uint2 hi, lo;
  split4(v, hi, lo);
  *reinterpret_cast<uint2*>(p + col) = hi;
  *reinterpret_cast<uint2*>(p + K + col) = lo;
}
template <typename T> __device__ __forceinline__ float4 load_row4(const T* row, int col, int K) {
  (void)K;
  return Vec4<T>::load(row + col);
}
template <> __device__ __forceinline__ float4 load_row4<bf16s>(const bf16s* row, int col, int K) {
  const bf16* p = reinterpret_cast<const bf16*>(row);
  const float4 a = Vec4<bf16>::load(p + col), b = Vec4<bf16>::load(p + K + col);
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace msq
