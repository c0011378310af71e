// fp32-accumulate FFMA GEMM, C = act(A W^T + bias) + resid.  Used (a) for every contraction of the
// "precise" fp32 parity mode and (b) for the small fp32 GEMMs after pooling (paragraph encoder, decode
// pre-projections) whose outputs feed the beam search and must stay fp32.
// 128x128x16 tiles, 256 threads, 8x8 register micro-tile, register-prefetch double buffering.
// Fixed k-ascending summation order per output -> run-to-run deterministic.
#include "kernels.cuh"

namespace msq {

constexpr int SG_BK = 16;

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

template <typename T, int KV> struct LoadK {
  static __device__ __forceinline__ void load(const T* p, bool ok, float* v) { Load8<T>::load(p, ok, v); }
};
template <> struct LoadK<float, 4> {
  static __device__ __forceinline__ void load(const float* p, bool ok, float* v) {
    float4 a = ok ? *reinterpret_cast<const float4*>(p) : make_float4(0.f, 0.f, 0.f, 0.f);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  }
};
template <> struct LoadK<bf16, 4> {
  static __device__ __forceinline__ void load(const bf16* p, bool ok, float* v) {
    if (ok) {
      uint2 u = *reinterpret_cast<const uint2*>(p);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
      v[0] = __low2float(h[0]); v[1] = __high2float(h[0]); v[2] = __low2float(h[1]); v[3] = __high2float(h[1]);
    } else {
      v[0] = v[1] = v[2] = v[3] = 0.f;
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

// TILE = 128: 8x8 micro-tile (large M);  TILE = 64: 4x4 micro-tile, 4x the CTAs (the M <= ~1k GEMMs behind
// the decoder would otherwise occupy a third of the SMs)
template <typename TA, typename TO, int TILE>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const TA* __restrict__ A, const TA* __restrict__ W,
                                                        const float* __restrict__ bias, const float* __restrict__ resid,
                                                        TO* __restrict__ C, float* __restrict__ C2, int64_t M, int N, int K,
                                                        int lda, int ldw, int ldc, int ldr, int act, float* __restrict__ partial) {
  pdl_sync();
  constexpr int SG_BM = TILE, SG_BN = TILE, SG_LD = TILE + 4, TM = TILE / 16, HM = TM / 2;
  constexpr int KV = TILE == 128 ? 8 : 4;  // k elements each thread stages per tile row
  __shared__ __align__(16) float As[2][SG_BK][SG_LD];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_LD];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.y * SG_BM;
  const int n0 = blockIdx.x * SG_BN;
  const int lrow = tid % TILE, lk = (tid / TILE) * KV;
  const bool a_ok = (m0 + lrow) < M, b_ok = (n0 + lrow) < N;
  // split-K (partial != nullptr): block z accumulates k slabs [z*per, min(nk_all, (z+1)*per)) and stores the raw sums
  const int nk_all = K / SG_BK, per = (nk_all + (int)gridDim.z - 1) / (int)gridDim.z;
  const int kbeg = (int)blockIdx.z * per, nk = max(0, min(per, nk_all - kbeg));
  const TA* ap = A + (m0 + lrow) * (int64_t)lda + lk + (int64_t)kbeg * SG_BK;
  const TA* bp = W + (int64_t)(n0 + lrow) * ldw + lk + (int64_t)kbeg * SG_BK;
  const int ty = tid >> 4, tx = tid & 15;

  float acc[TM][TM];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TM; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  LoadK<TA, KV>::load(ap, a_ok && nk > 0, ra);
  LoadK<TA, KV>::load(bp, b_ok && nk > 0, rb);
#pragma unroll
  for (int i = 0; i < KV; ++i) { As[0][lk + i][lrow] = ra[i]; Bs[0][lk + i][lrow] = rb[i]; }
  __syncthreads();

  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      LoadK<TA, KV>::load(ap + (kt + 1) * SG_BK, a_ok, ra);
      LoadK<TA, KV>::load(bp + (kt + 1) * SG_BK, b_ok, rb);
    }
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float av[TM], bv[TM];
      if (TILE == 128) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
        av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[TM - 4] = a1.x; av[TM - 3] = a1.y; av[TM - 2] = a1.z; av[TM - 1] = a1.w;
        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w; bv[TM - 4] = b1.x; bv[TM - 3] = b1.y; bv[TM - 2] = b1.z; bv[TM - 1] = b1.w;
      } else {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
        av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TM; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
#pragma unroll
      for (int i = 0; i < KV; ++i) { As[cur ^ 1][lk + i][lrow] = ra[i]; Bs[cur ^ 1][lk + i][lrow] = rb[i]; }
    }
    __syncthreads();
  }
  (void)HM;

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t row = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (row >= M) continue;
#pragma unroll
    for (int jh = 0; jh < TM / 4; ++jh) {
      const int col = n0 + jh * 64 + tx * 4;
      if (col >= N) continue;  // N % 4 == 0
      float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      if (partial) {
        *reinterpret_cast<float4*>(partial + ((int64_t)blockIdx.z * M + row) * N + col) = v;
        continue;
      }
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

// second pass of the split-K path: C = act(sum_z partial[z] + bias) + resid, splits added in index order (deterministic)
template <typename TO>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, int S, int64_t M, int N,
                                                            const float* __restrict__ bias, const float* __restrict__ resid, int ldr,
                                                            int act, TO* __restrict__ C, float* __restrict__ C2, int ldc) {
  pdl_sync();
  const int64_t total4 = M * (N / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / (N / 4);
    const int col = (int)(i % (N / 4)) * 4;
    float4 v = *reinterpret_cast<const float4*>(partial + row * N + col);
    for (int z = 1; z < S; ++z) {
      const float4 p = *reinterpret_cast<const float4*>(partial + ((int64_t)z * M + row) * N + col);
      v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    }
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

// grow-only scratch for split-K partial sums (per host thread; reallocation synchronises the device)
static int splitk_scratch(size_t floats, float** out) {
  static thread_local float* buf = nullptr;
  static thread_local size_t cap = 0;
  if (floats > cap) {
    if (buf) MSQ_CUDA(cudaFree(buf));
    buf = nullptr; cap = 0;
    MSQ_CUDA(cudaMalloc(&buf, floats * sizeof(float)));
    cap = floats;
  }
  *out = buf;
  return MSQ_OK;
}

template <typename TA, typename TO>
int gemm_simt(const GemmArgs& g, cudaStream_t st) {
  MSQ_REQUIRE(g.K % SG_BK == 0 && g.N % 4 == 0 && g.lda % 8 == 0 && g.ldw % 8 == 0 && g.ldc % 4 == 0,
              "gemm_simt: K=%d N=%d lda=%d ldw=%d ldc=%d not supported", g.K, g.N, g.lda, g.ldw, g.ldc);
  if (g.M == 0) return MSQ_OK;
  // small problems: 64x64 tiles give 4x the CTAs (fills the 148 SMs when M is a few hundred rows)
  if ((int64_t)ceil_div(g.N, 128) * ceil_div(g.M, 128) < 2 * 148) {
    dim3 grid(ceil_div(g.N, 64), ceil_div(g.M, 64));
    // The fp32 heads (paragraph encoder, decoder pre-projections: M = B*N .. B*N^2 rows against 7-10 MB of weights) are a
    // few dozen 64 x 64 tiles, each walking all of K on its own SM.  Split K (by a rule that depends on K only, so a
    // row's summation order does not change with the batch) and add the partial sums in a second, ordered pass.
    const int S = sizeof(TA) == 4 ? min(16, g.K / 256) : 1;
    if (S > 1) {
      float* part;
      MSQ_TRY(splitk_scratch((size_t)S * g.M * g.N, &part));
      grid.z = S;
      MSQ_CUDA(launch_k(gemm_simt_kernel<TA, TO, 64>, dim3(grid), dim3(256), 0, st, (const TA*)g.A, (const TA*)g.W, g.bias, g.resid, (TO*)g.C, g.C2, g.M, g.N, g.K, g.lda, g.ldw, g.ldc, g.ldr, g.act, part));
      MSQ_LAUNCH_CHECK();
      const int64_t total4 = g.M * (g.N / 4);
      MSQ_CUDA(launch_k(splitk_reduce_kernel<TO>, dim3((unsigned)min((int64_t)148 * 8, (total4 + 255) / 256)), dim3(256), 0, st, (const float*)part, S, g.M, g.N, g.bias, g.resid, g.ldr, g.act, (TO*)g.C, g.C2, g.ldc));
      MSQ_LAUNCH_CHECK();
      return MSQ_OK;
    }
    MSQ_CUDA(launch_k(gemm_simt_kernel<TA, TO, 64>, dim3(grid), dim3(256), 0, st, (const TA*)g.A, (const TA*)g.W, g.bias, g.resid, (TO*)g.C, g.C2, g.M, g.N, g.K, g.lda, g.ldw, g.ldc, g.ldr, g.act, (float*)nullptr));
  } else {
    dim3 grid(ceil_div(g.N, 128), ceil_div(g.M, 128));
    MSQ_CUDA(launch_k(gemm_simt_kernel<TA, TO, 128>, dim3(grid), dim3(256), 0, st, (const TA*)g.A, (const TA*)g.W, g.bias, g.resid, (TO*)g.C, g.C2, g.M, g.N, g.K, g.lda, g.ldw, g.ldc, g.ldr, g.act, (float*)nullptr));
  }
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int gemm_simt<float, float>(const GemmArgs&, cudaStream_t);
template int gemm_simt<bf16, float>(const GemmArgs&, cudaStream_t);
template int gemm_simt<bf16, bf16>(const GemmArgs&, cudaStream_t);

}  // namespace msq
