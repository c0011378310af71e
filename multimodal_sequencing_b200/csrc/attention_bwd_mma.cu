// Attention backward on tensor cores for the bf16 fine-tuning path (autograd of BertAttention.forward,
// models/CLIP/src/lxrt/modeling.py:398-425, and of nn.MultiheadAttention inside ResidualAttentionBlock,
// models/CLIP/clip/model.py:204-226).  One CTA per (token group, head); Q, K, V and dO of the item (<= 256 tokens x 64)
// live in shared memory as bf16; every contraction is mma.sync m16n8k16 (bf16 in, fp32 accumulate):
//
//   D_i   = dO_i . O_i                                  (row sums, computed while loading; O = saved context)
//   pass A (a warp owns 16 query rows): S = scale Q K^T + mask twice -- once for the row log-sum-exp, once to form
//          P = exp(S - lse), dP = dO V^T, dS = P (dP - D) in registers, fed straight back as the A operand of dQ += dS K
//   pass B (a warp owns 16 key rows):   S^T, dP^T from (K, Q) and (V, dO), P^T / dS^T in registers,
//          dV += P^T dO,  dK += dS^T Q
// Nothing of size L x L touches shared or global memory.  (Not tcgen05: the item is 227 x 227 x 64 -- five small,
// dependent GEMMs with a softmax between them -- and the accumulator-to-operand register reuse of mma.sync is what keeps
// P and dS on chip without a TMEM round trip.  The fp32 parity mode keeps the CUDA-core kernels of train_kernels.cu.)
#include "kernels.cuh"

namespace msq {

constexpr int ABM_LD = 72;      // bf16 elements per shared-memory row (64 + 8: conflict-free ldmatrix)
constexpr int ABM_WARPS = 8;

__device__ __forceinline__ void ldsm4(uint32_t* r, const bf16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm4t(uint32_t* r, const bf16* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float dot8(uint4 a, uint4 b) {
  const __nv_bfloat162* x = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* y = reinterpret_cast<const __nv_bfloat162*>(&b);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 p = __bfloat1622float2(x[i]), q = __bfloat1622float2(y[i]);
    s = fmaf(p.x, q.x, s);
    s = fmaf(p.y, q.y, s);
  }
  return s;
}

// A operand (16 rows x 64) of row block rb: 4 k-steps x 4 registers
__device__ __forceinline__ void load_a64(uint32_t (*a)[4], const bf16* M, int rb, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm4(a[ks], M + (rb * 16 + (lane & 15)) * ABM_LD + ks * 16 + (lane >> 4) * 8);
}
// acc[2][4] (16 x 16 tile) = A(16 x 64) * M[cb*16 .. +16][0..64)^T   (M rows are the n index, contiguous k)
__device__ __forceinline__ void mma_nt16(float (*acc)[4], const uint32_t (*a)[4], const bf16* M, int cb, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[0][i] = acc[1][i] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t b[4];
    ldsm4(b, M + (cb * 16 + (lane & 7) + ((lane >> 4) << 3)) * ABM_LD + ks * 16 + ((lane >> 3) & 1) * 8);
    mma16816(acc[0], a[ks], b[0], b[1]);
    mma16816(acc[1], a[ks], b[2], b[3]);
  }
}
// acc[8][4] (16 x 64) += A(16 x 16, registers) * M[kb*16 .. +16][0..64)   (M rows are the k index)
__device__ __forceinline__ void mma_nn64(float (*acc)[4], const uint32_t* a, const bf16* M, int kb, int lane) {
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) {
    uint32_t b[4];
    ldsm4t(b, M + (kb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * ABM_LD + n2 * 16 + (lane >> 4) * 8);
    mma16816(acc[2 * n2], a, b[0], b[1]);
    mma16816(acc[2 * n2 + 1], a, b[2], b[3]);
  }
}

__global__ void __launch_bounds__(ABM_WARPS * 32) attention_bwd_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ ctx,
                                                                           const bf16* __restrict__ dctx, int L, int Lp, int heads, float scale,
                                                                           const float* __restrict__ mask_add, int mask_ld, int mask_len,
                                                                           bf16* __restrict__ dqkv) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Ks = Qs + Lp * ABM_LD;
  bf16* Vs = Ks + Lp * ABM_LD;
  bf16* Gs = Vs + Lp * ABM_LD;                       // dO
  float* Ms = reinterpret_cast<float*>(Gs + Lp * ABM_LD);   // additive key mask, -inf for the padded keys
  float* Ds = Ms + Lp;                               // D_i
  float* Ls = Ds + Lp;                               // row log-sum-exp
  const int r = blockIdx.x / heads, h = blockIdx.x % heads;
  const int ld = 3 * heads * 64, ldc = heads * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;

  for (int base = 0; base < Lp * 8; base += ABM_WARPS * 32) {
    const int idx = base + tid;
    const bool in = idx < Lp * 8;
    const int t = in ? idx >> 3 : 0, pc = idx & 7;
    uint4 q = make_uint4(0, 0, 0, 0), k = q, v = q, g = q;
    float dpart = 0.f;
    if (in && t < L) {
      const bf16* src = qkv + ((int64_t)r * L + t) * ld + h * 64 + pc * 8;
      q = *reinterpret_cast<const uint4*>(src);
      k = *reinterpret_cast<const uint4*>(src + heads * 64);
      v = *reinterpret_cast<const uint4*>(src + 2 * heads * 64);
      const int64_t co = ((int64_t)r * L + t) * ldc + h * 64 + pc * 8;
      g = *reinterpret_cast<const uint4*>(dctx + co);
      dpart = dot8(g, *reinterpret_cast<const uint4*>(ctx + co));
    }
    if (in) {
      const int so = t * ABM_LD + pc * 8;
      *reinterpret_cast<uint4*>(Qs + so) = q;
      *reinterpret_cast<uint4*>(Ks + so) = k;
      *reinterpret_cast<uint4*>(Vs + so) = v;
      *reinterpret_cast<uint4*>(Gs + so) = g;
    }
    dpart += __shfl_xor_sync(0xffffffffu, dpart, 1);
    dpart += __shfl_xor_sync(0xffffffffu, dpart, 2);
    dpart += __shfl_xor_sync(0xffffffffu, dpart, 4);
    if (in && pc == 0) Ds[t] = dpart;
  }
  for (int t = tid; t < Lp; t += ABM_WARPS * 32)
    Ms[t] = t < L ? ((mask_add && t < mask_len) ? mask_add[(int64_t)r * mask_ld + t] : 0.f) : -INFINITY;
  __syncthreads();

  const int nblk = Lp >> 4;
  // ---------------- pass A: dQ, row statistics
  for (int ib = warp; ib < nblk; ib += ABM_WARPS) {
    uint32_t aq[4][4], ag[4][4];
    load_a64(aq, Qs, ib, lane);
    load_a64(ag, Gs, ib, lane);
    float m[2] = {-INFINITY, -INFINITY}, s[2] = {0.f, 0.f};
    for (int kb = 0; kb < nblk; ++kb) {
      float acc[2][4];
      mma_nt16(acc, aq, Ks, kb, lane);
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {          // hr: row gid (0) / gid + 8 (1)
        float vals[4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) vals[nt * 2 + e] = fmaf(acc[nt][hr * 2 + e], scale, Ms[kb * 16 + nt * 8 + tig * 2 + e]);
        const float um = fmaxf(fmaxf(vals[0], vals[1]), fmaxf(vals[2], vals[3]));
        const float mn = fmaxf(m[hr], um);
        if (mn > -INFINITY) {
          s[hr] = s[hr] * __expf(m[hr] - mn) + (__expf(vals[0] - mn) + __expf(vals[1] - mn)) + (__expf(vals[2] - mn) + __expf(vals[3] - mn));
          m[hr] = mn;
        }
      }
    }
    float lse[2], dd[2];
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, m[hr], o), so = __shfl_xor_sync(0xffffffffu, s[hr], o);
        const float mn = fmaxf(m[hr], mo);
        if (mn > -INFINITY) {
          s[hr] = s[hr] * __expf(m[hr] - mn) + so * __expf(mo - mn);
          m[hr] = mn;
        }
      }
      lse[hr] = m[hr] + __logf(s[hr]);
      const int row = ib * 16 + gid + hr * 8;
      dd[hr] = Ds[row];
      if (tig == 0) Ls[row] = lse[hr];
    }
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    for (int kb = 0; kb < nblk; ++kb) {
      float sa[2][4], dp[2][4];
      mma_nt16(sa, aq, Ks, kb, lane);
      mma_nt16(dp, ag, Vs, kb, lane);
      uint32_t a[4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        float ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int hr = e >> 1;
          const float p = __expf(fmaf(sa[nt][e], scale, Ms[kb * 16 + nt * 8 + tig * 2 + (e & 1)]) - lse[hr]);
          ds[e] = p * (dp[nt][e] - dd[hr]);
        }
        a[nt * 2] = pack2(ds[0], ds[1]);
        a[nt * 2 + 1] = pack2(ds[2], ds[3]);
      }
      mma_nn64(dq, a, Ks, kb, lane);
    }
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int row = ib * 16 + gid + hr * 8;
      if (row < L) {
        bf16* op = dqkv + ((int64_t)r * L + row) * ld + h * 64 + tig * 2;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<uint32_t*>(op + nt * 8) = pack2(dq[nt][hr * 2] * scale, dq[nt][hr * 2 + 1] * scale);
      }
    }
  }
  __syncthreads();
  // ---------------- pass B: dK, dV
  for (int jb = warp; jb < nblk; jb += ABM_WARPS) {
    uint32_t ak[4][4], av[4][4];
    load_a64(ak, Ks, jb, lane);
    load_a64(av, Vs, jb, lane);
    const float mr[2] = {Ms[jb * 16 + gid], Ms[jb * 16 + gid + 8]};
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    for (int qb = 0; qb < nblk; ++qb) {
      float st[2][4], dp[2][4];
      mma_nt16(st, ak, Qs, qb, lane);
      mma_nt16(dp, av, Gs, qb, lane);
      uint32_t pa[4], da[4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        float p[4], ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qc = qb * 16 + nt * 8 + tig * 2 + (e & 1);
          p[e] = __expf(fmaf(st[nt][e], scale, mr[e >> 1]) - Ls[qc]);
          ds[e] = p[e] * (dp[nt][e] - Ds[qc]);
        }
        pa[nt * 2] = pack2(p[0], p[1]); pa[nt * 2 + 1] = pack2(p[2], p[3]);
        da[nt * 2] = pack2(ds[0], ds[1]); da[nt * 2 + 1] = pack2(ds[2], ds[3]);
      }
      mma_nn64(dv, pa, Gs, qb, lane);
      mma_nn64(dk, da, Qs, qb, lane);
    }
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int row = jb * 16 + gid + hr * 8;
      if (row < L) {
        bf16* op = dqkv + ((int64_t)r * L + row) * ld + heads * 64 + h * 64 + tig * 2;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          *reinterpret_cast<uint32_t*>(op + nt * 8) = pack2(dk[nt][hr * 2] * scale, dk[nt][hr * 2 + 1] * scale);
          *reinterpret_cast<uint32_t*>(op + heads * 64 + nt * 8) = pack2(dv[nt][hr * 2], dv[nt][hr * 2 + 1]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Two-kernel variant (default): the fused kernel above keeps all four matrices resident (138 KB at L = 227) and needs 167
// registers, i.e. ONE CTA of 8 warps per SM (ncu: warps_active 12 %, tensor pipe 37 %).  Here pass A keeps only K and V in
// shared memory and pass B only Q and dO (69 KB each); the 16-row operand a warp owns (Q_i / dO_i, K_j / V_j) is loaded
// straight from global memory into the mma A-fragment layout, D_i = dO_i . O_i is formed from the fragments, and the row
// statistics (lse, D) travel through a small global scratch.  Two to three CTAs fit per SM.
// ---------------------------------------------------------------------------------------------------
// A operand (16 rows x 64) of row block rb read from global memory (row stride ld elements); rows >= L are zero
__device__ __forceinline__ void load_a64_global(uint32_t (*a)[4], const bf16* M, int64_t ld, int rb, int L, int lane) {
  const int gid = lane >> 2, tig = lane & 3;
  const int r0 = rb * 16 + gid, r1 = r0 + 8;
  const bf16* p0 = M + (int64_t)r0 * ld + tig * 2;
  const bf16* p1 = M + (int64_t)r1 * ld + tig * 2;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    a[ks][0] = r0 < L ? *reinterpret_cast<const uint32_t*>(p0 + ks * 16) : 0u;
    a[ks][1] = r1 < L ? *reinterpret_cast<const uint32_t*>(p1 + ks * 16) : 0u;
    a[ks][2] = r0 < L ? *reinterpret_cast<const uint32_t*>(p0 + ks * 16 + 8) : 0u;
    a[ks][3] = r1 < L ? *reinterpret_cast<const uint32_t*>(p1 + ks * 16 + 8) : 0u;
  }
}
__device__ __forceinline__ float dot2(uint32_t x, uint32_t y) {
  const float2 p = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&x));
  const float2 q = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&y));
  return fmaf(p.x, q.x, p.y * q.y);
}
// cooperative load of two [L, 64] matrices (row stride ld) into padded shared memory, rows >= L zeroed
__device__ __forceinline__ void load_two(bf16* S0, bf16* S1, const bf16* G0, int64_t ld0, const bf16* G1, int64_t ld1, int L, int Lp) {
  for (int idx = threadIdx.x; idx < Lp * 8; idx += blockDim.x) {
    const int t = idx >> 3, pc = idx & 7;
    uint4 a = make_uint4(0, 0, 0, 0), b = a;
    if (t < L) {
      a = *reinterpret_cast<const uint4*>(G0 + (int64_t)t * ld0 + pc * 8);
      b = *reinterpret_cast<const uint4*>(G1 + (int64_t)t * ld1 + pc * 8);
    }
    *reinterpret_cast<uint4*>(S0 + t * ABM_LD + pc * 8) = a;
    *reinterpret_cast<uint4*>(S1 + t * ABM_LD + pc * 8) = b;
  }
}

// LSE_IN: the forward kernel saved the row log-sum-exp (lse_out is then an INPUT): S is formed once, not twice
template <bool LSE_IN>
__global__ void __launch_bounds__(ABM_WARPS * 32, 2) attention_bwd_dq_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ ctx,
                                                                                 const bf16* __restrict__ dctx, int L, int Lp, int heads, float scale,
                                                                                 const float* __restrict__ mask_add, int mask_ld, int mask_len,
                                                                                 bf16* __restrict__ dqkv, float* __restrict__ lse_out,
                                                                                 float* __restrict__ dsum_out, Drop drop) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* Ks = reinterpret_cast<bf16*>(smraw);
  bf16* Vs = Ks + Lp * ABM_LD;
  float* Ms = reinterpret_cast<float*>(Vs + Lp * ABM_LD);
  const int r = blockIdx.x / heads, h = blockIdx.x % heads;
  const int ld = 3 * heads * 64, ldc = heads * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const bf16* qbase = qkv + (int64_t)r * L * ld + h * 64;
  const bf16* gbase = dctx + (int64_t)r * L * ldc + h * 64;
  const bf16* obase = ctx + (int64_t)r * L * ldc + h * 64;
  load_two(Ks, Vs, qbase + heads * 64, ld, qbase + 2 * heads * 64, ld, L, Lp);
  for (int t = tid; t < Lp; t += ABM_WARPS * 32)
    Ms[t] = t < L ? ((mask_add && t < mask_len) ? mask_add[(int64_t)r * mask_ld + t] : 0.f) : -INFINITY;
  __syncthreads();
  const int nblk = Lp >> 4;
  for (int ib = warp; ib < nblk; ib += ABM_WARPS) {
    uint32_t aq[4][4], ag[4][4];
    load_a64_global(aq, qbase, ld, ib, L, lane);
    load_a64_global(ag, gbase, ldc, ib, L, lane);
    float dd[2] = {0.f, 0.f};
    {
      uint32_t ao[4][4];
      load_a64_global(ao, obase, ldc, ib, L, lane);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        dd[0] += dot2(ag[ks][0], ao[ks][0]) + dot2(ag[ks][2], ao[ks][2]);
        dd[1] += dot2(ag[ks][1], ao[ks][1]) + dot2(ag[ks][3], ao[ks][3]);
      }
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        dd[0] += __shfl_xor_sync(0xffffffffu, dd[0], o);
        dd[1] += __shfl_xor_sync(0xffffffffu, dd[1], o);
      }
    }
    float lse[2];
    if constexpr (LSE_IN) {
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        const int row = ib * 16 + gid + hr * 8;
        lse[hr] = row < L ? lse_out[(int64_t)blockIdx.x * L + row] : INFINITY;   // rows beyond the sequence: p = 0
        if (tig == 0 && row < L) dsum_out[(int64_t)blockIdx.x * L + row] = dd[hr];
      }
    } else {
    float m[2] = {-INFINITY, -INFINITY}, s[2] = {0.f, 0.f};
    for (int kb = 0; kb < nblk; ++kb) {
      float acc[2][4];
      mma_nt16(acc, aq, Ks, kb, lane);
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        float vals[4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) vals[nt * 2 + e] = fmaf(acc[nt][hr * 2 + e], scale, Ms[kb * 16 + nt * 8 + tig * 2 + e]);
        const float um = fmaxf(fmaxf(vals[0], vals[1]), fmaxf(vals[2], vals[3]));
        const float mn = fmaxf(m[hr], um);
        if (mn > -INFINITY) {
          s[hr] = s[hr] * __expf(m[hr] - mn) + (__expf(vals[0] - mn) + __expf(vals[1] - mn)) + (__expf(vals[2] - mn) + __expf(vals[3] - mn));
          m[hr] = mn;
        }
      }
    }
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        const float mo = __shfl_xor_sync(0xffffffffu, m[hr], o), so = __shfl_xor_sync(0xffffffffu, s[hr], o);
        const float mn = fmaxf(m[hr], mo);
        if (mn > -INFINITY) {
          s[hr] = s[hr] * __expf(m[hr] - mn) + so * __expf(mo - mn);
          m[hr] = mn;
        }
      }
      lse[hr] = m[hr] + __logf(s[hr]);
      const int row = ib * 16 + gid + hr * 8;
      if (tig == 0 && row < L) {
        lse_out[(int64_t)blockIdx.x * L + row] = lse[hr];
        dsum_out[(int64_t)blockIdx.x * L + row] = dd[hr];
      }
    }
    }
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    for (int kb = 0; kb < nblk; ++kb) {
      float sa[2][4], dp[2][4];
      mma_nt16(sa, aq, Ks, kb, lane);
      mma_nt16(dp, ag, Vs, kb, lane);
      uint32_t a[4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        float ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int hr = e >> 1;
          const float p = __expf(fmaf(sa[nt][e], scale, Ms[kb * 16 + nt * 8 + tig * 2 + (e & 1)]) - lse[hr]);
          float dpe = dp[nt][e];
          if (drop.thresh)   // dP = (dO V^T) . mask; D = rowsum(dO . O) already holds the masked sum
            dpe *= drop_mul(drop, ((uint64_t)blockIdx.x * L + (uint64_t)(ib * 16 + gid + hr * 8)) * L + (uint64_t)(kb * 16 + nt * 8 + tig * 2 + (e & 1)));
          ds[e] = p * (dpe - dd[hr]);
        }
        a[nt * 2] = pack2(ds[0], ds[1]);
        a[nt * 2 + 1] = pack2(ds[2], ds[3]);
      }
      mma_nn64(dq, a, Ks, kb, lane);
    }
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int row = ib * 16 + gid + hr * 8;
      if (row < L) {
        bf16* op = dqkv + ((int64_t)r * L + row) * ld + h * 64 + tig * 2;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) *reinterpret_cast<uint32_t*>(op + nt * 8) = pack2(dq[nt][hr * 2] * scale, dq[nt][hr * 2 + 1] * scale);
      }
    }
  }
}

__global__ void __launch_bounds__(ABM_WARPS * 32, 2) attention_bwd_dkv_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dctx, int L,
                                                                                  int Lp, int heads, float scale, const float* __restrict__ mask_add,
                                                                                  int mask_ld, int mask_len, bf16* __restrict__ dqkv,
                                                                                  const float* __restrict__ lse_in, const float* __restrict__ dsum_in, Drop drop) {
  pdl_sync();
  extern __shared__ __align__(16) unsigned char smraw[];
  bf16* Qs = reinterpret_cast<bf16*>(smraw);
  bf16* Gs = Qs + Lp * ABM_LD;
  float* Ls = reinterpret_cast<float*>(Gs + Lp * ABM_LD);
  float* Ds = Ls + Lp;
  const int r = blockIdx.x / heads, h = blockIdx.x % heads;
  const int ld = 3 * heads * 64, ldc = heads * 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const bf16* qbase = qkv + (int64_t)r * L * ld + h * 64;
  const bf16* gbase = dctx + (int64_t)r * L * ldc + h * 64;
  load_two(Qs, Gs, qbase, ld, gbase, ldc, L, Lp);
  for (int t = tid; t < Lp; t += ABM_WARPS * 32) {
    Ls[t] = t < L ? lse_in[(int64_t)blockIdx.x * L + t] : 0.f;
    Ds[t] = t < L ? dsum_in[(int64_t)blockIdx.x * L + t] : 0.f;
  }
  __syncthreads();
  const int nblk = Lp >> 4;
  for (int jb = warp; jb < nblk; jb += ABM_WARPS) {
    uint32_t ak[4][4], av[4][4];
    load_a64_global(ak, qbase + heads * 64, ld, jb, L, lane);
    load_a64_global(av, qbase + 2 * heads * 64, ld, jb, L, lane);
    float mr[2];
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int j = jb * 16 + gid + hr * 8;
      mr[hr] = j < L ? ((mask_add && j < mask_len) ? mask_add[(int64_t)r * mask_ld + j] : 0.f) : -INFINITY;
    }
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    for (int qb = 0; qb < nblk; ++qb) {
      float st[2][4], dp[2][4];
      mma_nt16(st, ak, Qs, qb, lane);
      mma_nt16(dp, av, Gs, qb, lane);
      uint32_t pa[4], da[4];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        float p[4], ds[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qc = qb * 16 + nt * 8 + tig * 2 + (e & 1);
          p[e] = __expf(fmaf(st[nt][e], scale, mr[e >> 1]) - Ls[qc]);
          float dm = 1.f;
          if (drop.thresh) dm = drop_mul(drop, ((uint64_t)blockIdx.x * L + (uint64_t)qc) * L + (uint64_t)(jb * 16 + gid + (e >> 1) * 8));
          ds[e] = p[e] * (dp[nt][e] * dm - Ds[qc]);
          p[e] *= dm;                                  // dV = (P . mask)^T dO
        }
        pa[nt * 2] = pack2(p[0], p[1]); pa[nt * 2 + 1] = pack2(p[2], p[3]);
        da[nt * 2] = pack2(ds[0], ds[1]); da[nt * 2 + 1] = pack2(ds[2], ds[3]);
      }
      mma_nn64(dv, pa, Gs, qb, lane);
      mma_nn64(dk, da, Qs, qb, lane);
    }
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int row = jb * 16 + gid + hr * 8;
      if (row < L) {
        bf16* op = dqkv + ((int64_t)r * L + row) * ld + heads * 64 + h * 64 + tig * 2;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
          *reinterpret_cast<uint32_t*>(op + nt * 8) = pack2(dk[nt][hr * 2] * scale, dk[nt][hr * 2 + 1] * scale);
          *reinterpret_cast<uint32_t*>(op + heads * 64 + nt * 8) = pack2(dv[nt][hr * 2], dv[nt][hr * 2 + 1]);
        }
      }
    }
  }
}

bool attention_bwd_mma_supported(int L) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("MSQ_ATTN_BWD_SIMT"); off = (e && e[0] == '1') ? 1 : 0; }
  return !off && L >= 1 && L <= 256;
}

int attention_bwd_mma(const bf16* qkv, const bf16* ctx, const bf16* dctx, int64_t R, int L, int heads, float scale, const float* key_mask_add,
                      int mask_ld, int mask_len, bf16* dqkv, float* scratch, cudaStream_t st, const Drop& drop, const float* lse_fwd) {
  MSQ_REQUIRE(L >= 1 && L <= 256, "attention_bwd_mma: sequence length %d out of range", L);
  MSQ_REQUIRE((((uintptr_t)qkv | (uintptr_t)ctx | (uintptr_t)dctx | (uintptr_t)dqkv) & 15) == 0, "attention_bwd_mma: unaligned pointer");
  if (R == 0) return MSQ_OK;
  const int Lp = (L + 15) & ~15;
  static int fused = -1;
  if (fused < 0) { const char* e = getenv("MSQ_ATTN_BWD_FUSED"); fused = (e && e[0] == '1') ? 1 : 0; }
  if (fused || !scratch) {
    MSQ_REQUIRE(drop.thresh == 0, "attention_bwd_mma: the single-kernel form (MSQ_ATTN_BWD_FUSED=1) has no dropout; unset it to train with dropout");
    const size_t smem = (size_t)4 * Lp * ABM_LD * sizeof(bf16) + (size_t)3 * Lp * sizeof(float);
    MSQ_SMEM_ATTR(smem, attention_bwd_mma_kernel);
    MSQ_CUDA(launch_k(attention_bwd_mma_kernel, dim3((unsigned)(R * heads)), dim3(ABM_WARPS * 32), smem, st, qkv, ctx, dctx, L, Lp, heads, scale, key_mask_add, mask_ld, mask_len, dqkv));
    MSQ_LAUNCH_CHECK();
    return MSQ_OK;
  }
  float* lse = lse_fwd ? const_cast<float*>(lse_fwd) : scratch;   // saved by the forward kernel, or computed by the dQ kernel
  float* dsum = scratch + (size_t)R * heads * L;
  const size_t smem_a = (size_t)2 * Lp * ABM_LD * sizeof(bf16) + (size_t)Lp * sizeof(float);
  const size_t smem_b = (size_t)2 * Lp * ABM_LD * sizeof(bf16) + (size_t)2 * Lp * sizeof(float);
  MSQ_SMEM_ATTR(smem_a, attention_bwd_dq_mma_kernel<false>);
  MSQ_SMEM_ATTR(smem_a, attention_bwd_dq_mma_kernel<true>);
  MSQ_SMEM_ATTR(smem_b, attention_bwd_dkv_mma_kernel);
  if (lse_fwd) MSQ_CUDA(launch_k(attention_bwd_dq_mma_kernel<true>, dim3((unsigned)(R * heads)), dim3(ABM_WARPS * 32), smem_a, st, qkv, ctx, dctx, L, Lp, heads, scale, key_mask_add, mask_ld, mask_len, dqkv, lse, dsum, drop));
  else MSQ_CUDA(launch_k(attention_bwd_dq_mma_kernel<false>, dim3((unsigned)(R * heads)), dim3(ABM_WARPS * 32), smem_a, st, qkv, ctx, dctx, L, Lp, heads, scale, key_mask_add, mask_ld, mask_len, dqkv, lse, dsum, drop));
  MSQ_LAUNCH_CHECK();
  // dK / dV: the tcgen05 pipeline (attention_bwd_tc.cu) when available, else the mma.sync kernel (MSQ_ATTN_BWD_TC=0)
  if (attention_bwd_dkv_tc_supported(L))
    return attention_bwd_dkv_tc(qkv, dctx, R, L, heads, scale, key_mask_add, mask_ld, mask_len, dqkv, lse, dsum, st, drop);
  MSQ_CUDA(launch_k(attention_bwd_dkv_mma_kernel, dim3((unsigned)(R * heads)), dim3(ABM_WARPS * 32), smem_b, st, qkv, dctx, L, Lp, heads, scale, key_mask_add, mask_ld, mask_len, dqkv, (const float*)lse, (const float*)dsum, drop));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

}  // namespace msq
