// fp32-accumulate FFMA GEMM, C = act(A W^T + bias) + resid.  Used (a) for every contraction of the
// "precise" fp32 parity mode and (b) for the small fp32 GEMMs after pooling (paragraph encoder, decode
// pre-projections) whose outputs feed the beam search and must stay fp32.
// 128x128x16 tiles, 256 threads, 8x8 register micro-tile, register-prefetch double buffering.
// Fixed k-ascending summation order per output -> run-to-run deterministic.
#include "kernels.cuh"

namespace msq {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16, SG_LD = 132;

template <typename T> struct Load8;
template <> struct Load8<float> {
  static __device__ __forceinline__ void load(const float* p, bool ok, float* v) {
    if (ok) {
      float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
  }
};
template <> struct Load8<bf16> {
  static __device__ __forceinline__ void load(const bf16* p, bool ok, float* v) {
    if (ok) {
      uint4 u = *reinterpret_cast<const uint4*>(p);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = 0.f;
    }
  }
};

template <typename TO> __device__ __forceinline__ void store4(TO* p, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
template <> __device__ __forceinline__ void store4<bf16>(bf16* p, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

template <typename TA, typename TO>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const TA* __restrict__ A, const TA* __restrict__ W,
                                                        const float* __restrict__ bias, const float* __restrict__ resid,
                                                        TO* __restrict__ C, float* __restrict__ C2, int64_t M, int N, int K,
                                                        int lda, int ldw, int ldc, int ldr, int act) {
  __shared__ __align__(16) float As[2][SG_BK][SG_LD];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_LD];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * SG_BM;
  const int n0 = blockIdx.x * SG_BN;
  const int lrow = tid & 127, lk = (tid >> 7) * 8;
  const bool a_ok = (m0 + lrow) < M, b_ok = (n0 + lrow) < N;
  const TA* ap = A + (m0 + lrow) * (int64_t)lda + lk;
  const TA* bp = W + (int64_t)(n0 + lrow) * ldw + lk;
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  Load8<TA>::load(ap, a_ok, ra);
  Load8<TA>::load(bp, b_ok, rb);
#pragma unroll
  for (int i = 0; i < 8; ++i) { As[0][lk + i][lrow] = ra[i]; Bs[0][lk + i][lrow] = rb[i]; }
  __syncthreads();

  const int nk = K / SG_BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      Load8<TA>::load(ap + (kt + 1) * SG_BK, a_ok, ra);
      Load8<TA>::load(bp + (kt + 1) * SG_BK, b_ok, rb);
    }
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { As[cur ^ 1][lk + i][lrow] = ra[i]; Bs[cur ^ 1][lk + i][lrow] = rb[i]; }
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int col = n0 + jh * 64 + tx * 4;
      if (col >= N) continue;  // N % 4 == 0
      float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      if (bias) {
        const float4 b = *reinterpret_cast<const float4*>(bias + col);
        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
      }
      v.x = apply_act(v.x, act); v.y = apply_act(v.y, act); v.z = apply_act(v.z, act); v.w = apply_act(v.w, act);
      if (resid) {
        const float4 r = *reinterpret_cast<const float4*>(resid + row * ldr + col);
        v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
      }
      store4<TO>(C + row * ldc + col, v);
      if (C2) *reinterpret_cast<float4*>(C2 + row * ldc + col) = v;
    }
  }
}

template <typename TA, typename TO>
int gemm_simt(const GemmArgs& g, cudaStream_t st) {
  MSQ_REQUIRE(g.K % SG_BK == 0 && g.N % 4 == 0 && g.lda % 8 == 0 && g.ldw % 8 == 0 && g.ldc % 4 == 0,
              "gemm_simt: K=%d N=%d lda=%d ldw=%d ldc=%d not supported", g.K, g.N, g.lda, g.ldw, g.ldc);
  if (g.M == 0) return MSQ_OK;
  dim3 grid(ceil_div(g.N, SG_BN), ceil_div(g.M, SG_BM));
  gemm_simt_kernel<TA, TO><<<grid, 256, 0, st>>>((const TA*)g.A, (const TA*)g.W, g.bias, g.resid, (TO*)g.C, g.C2, g.M, g.N,
                                                 g.K, g.lda, g.ldw, g.ldc, g.ldr, g.act);
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int gemm_simt<float, float>(const GemmArgs&, cudaStream_t);
template int gemm_simt<bf16, float>(const GemmArgs&, cudaStream_t);
template int gemm_simt<bf16, bf16>(const GemmArgs&, cudaStream_t);

}  // namespace msq
