// Fused multi-head self-attention over one token group (a pair's joint 227-token sequence, or its
// 99 visual tokens): scores, additive key mask, softmax and P.V in one kernel; the whole K and V of a
// (group, head) live in shared memory so the softmax is single-pass (no online rescale).
//
// Replaces: BertAttention.forward  models/CLIP/src/lxrt/modeling.py:398-425 (and the text-only twin
//           models/berson/modeling_bert.py:208-241): scores / sqrt(d) + mask(-10000), softmax, P V;
//           nn.MultiheadAttention inside ResidualAttentionBlock  models/CLIP/clip/model.py:204-226.
// Input layout: qkv [R*L, 3*heads*64] rows = tokens, columns = (q | k | v), each heads x 64.
#include <stdlib.h>

#include "tc_common.cuh"

namespace msq {

constexpr int AT_D = 64;
constexpr int AT_WARPS = 16;

// fp32 CUDA-core version (exact-parity path; also serves bf16 I/O).
template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32) attention_simt_kernel(const T* __restrict__ qkv, int L, int heads,
                                                                       float scale, const float* __restrict__ mask_add,
                                                                       int mask_ld, int mask_len, T* __restrict__ ctx, Drop drop) {
  pdl_sync();
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
      // training: dropout of the (normalised) probabilities -- the row sum stays the full softmax denominator
      p[key] = e * drop_mul(drop, ((uint64_t)blockIdx.x * L + t) * L + key);
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


// ---------------------------------------------------------------------------------------------------
// tcgen05 version (bf16): one CTA per (token group, head), 256 threads (two per query row).
//   TMA loads Q (<= 2 tiles of 128 rows), K and V ([KP keys] x 64, 128B swizzle) straight out of the
//   packed qkv activation; S = Q K^T is ONE tcgen05.mma chain (M=128, N=KP, K=64) into TMEM; each
//   pair of threads owns one query row (TMEM lane), half the keys each: two passes over S (max, then exp2 / sum,
//   combined through shared memory), P is written
//   back over S as packed bf16 (tcgen05.st) and O = P V runs as a TMEM-A ("TS") tcgen05.mma chain with
//   V consumed MN-major from the same swizzled tile (no transpose anywhere); O is normalised by the
//   row sum in the epilogue.  TMEM columns: S [0,KP), P [0,KP/2), O [KP/2, KP/2+64).
//   KP = 256 -> 2 CTAs/SM (96 KB smem, 256 TMEM columns each); KP = 128 -> 4 CTAs/SM.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// SPL (bf16x3 mode): q, k, v and the probabilities are carried as hi + lo bf16 pairs.  qkv rows are
// [hi(3*heads*64) | lo(3*heads*64)], ctx rows [hi(heads*64) | lo(heads*64)]; every contraction is three MMA chains
// (hi*lo, lo*hi, hi*hi) into the same fp32 accumulator; P_lo lives in its own TMEM columns [KP, KP + KP/2).
template <int KP, bool SPL> struct AttnCfg {
  static constexpr int Q_BYTES = 128 * 64 * 2, KV_BYTES = KP * 64 * 2;
  static constexpr int QT = (SPL && KP == 128) ? 1 : 2;           // query tiles resident (L <= 128 needs one)
  static constexpr int PLANES = SPL ? 2 : 1;
  static constexpr int OPER_BYTES = PLANES * (QT * Q_BYTES + 2 * KV_BYTES);
  static constexpr int SMEM = OPER_BYTES + KP * 4 + 2048 + 64 + 1024;
  static constexpr int TMEM_COLS = SPL ? 2 * KP : KP;
};

template <int KP, bool SPL, bool DROP>
__global__ void __launch_bounds__(256) attention_tc_kernel(const __grid_constant__ CUtensorMap map_q,
                                                           const __grid_constant__ CUtensorMap map_kv, int L, int heads,
                                                           float scale_l2e, const float* __restrict__ mask_add, int mask_ld,
                                                           int mask_len, bf16* __restrict__ ctx, int n_items, Drop drop,
                                                           float* __restrict__ lse_out) {
  using Cfg = AttnCfg<KP, SPL>;
  constexpr int Q_BYTES = Cfg::Q_BYTES, KV_BYTES = Cfg::KV_BYTES, QT = Cfg::QT, OPER = Cfg::OPER_BYTES;
  constexpr int CH = KP / 64;  // 32-column chunks per thread: two threads share a query row, half the keys each
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base, sK = base + QT * Q_BYTES, sV = sK + KV_BYTES;
  // lo planes (SPL): same order behind the hi planes
  const uint32_t sQl = sV + KV_BYTES, sKl = sQl + QT * Q_BYTES, sVl = sKl + KV_BYTES;
  float* maskf = reinterpret_cast<float*>(gen + OPER);                         // [KP] additive mask * log2(e), -inf beyond L
  float* red = maskf + KP;                                                     // [2 (max|sum)][2 halves][128 rows]
  const uint32_t bars = base + OPER + KP * 4 + 2048;                           // bar_qk | bar_s | bar_o | bar_v | tmem slot
  volatile uint32_t* tslot_gen = reinterpret_cast<volatile uint32_t*>(gen + OPER + KP * 4 + 2048 + 32);
  const uint32_t bar_qk = bars, bar_s = bars + 8, bar_o = bars + 16, bar_v = bars + 24, tslot = bars + 32;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quarter = warp & 3, half = warp >> 2, trow = quarter * 32 + lane;
  const int nqt = (L + 127) >> 7;
  const float LOG2E = 1.4426950408889634f;

  if (warp == 1 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    mbar_init(bar_qk, 1);
    mbar_init(bar_s, 1);
    mbar_init(bar_o, 1);
    mbar_init(bar_v, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "n"(Cfg::TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_sync();  // barrier init / TMEM allocation above overlap the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot_gen;

  // Persistent over (token group, head) items: Q and K of the NEXT item are fetched as soon as this item's last S = Q K^T
  // has been read out of shared memory, V as soon as the last P V has: the TMA latency (and the per-CTA set-up) is paid once
  // per CTA, not once per item.
  auto load_qk = [&](int item) {   // tid 0
    const int row0 = (item / heads) * L, hh = item % heads;
    mbar_expect_tx(bar_qk, (uint32_t)(Cfg::PLANES * (nqt * Q_BYTES + KV_BYTES)));
    tma_load_2d(sK, &map_kv, heads * AT_D + hh * AT_D, row0, bar_qk);
    for (int qt = 0; qt < nqt; ++qt) tma_load_2d(sQ + qt * Q_BYTES, &map_q, hh * AT_D, row0 + qt * 128, bar_qk);
    if (SPL) {
      const int lo = 3 * heads * AT_D;   // the lo plane of a qkv row starts here
      tma_load_2d(sKl, &map_kv, lo + heads * AT_D + hh * AT_D, row0, bar_qk);
      for (int qt = 0; qt < nqt; ++qt) tma_load_2d(sQl + qt * Q_BYTES, &map_q, lo + hh * AT_D, row0 + qt * 128, bar_qk);
    }
  };
  auto load_v = [&](int item) {    // tid 0
    const int row0 = (item / heads) * L, hh = item % heads;
    mbar_expect_tx(bar_v, (uint32_t)(Cfg::PLANES * KV_BYTES));
    tma_load_2d(sV, &map_kv, 2 * heads * AT_D + hh * AT_D, row0, bar_v);
    if (SPL) tma_load_2d(sVl, &map_kv, 5 * heads * AT_D + hh * AT_D, row0, bar_v);
  };
  if (tid == 0 && (int)blockIdx.x < n_items) { load_qk(blockIdx.x); load_v(blockIdx.x); }

  // instruction descriptors: D=F32, A=B=BF16; S: N=KP, A/B K-major; PV: N=64, B MN-major (bit 16)
  const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KP >> 3) << 17) | (8u << 24);
  const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | (8u << 24);
  const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
  const int col0 = half * (KP / 2);  // first key column of this thread

  uint32_t n_done = 0, mma_phase = 0;   // items finished by this CTA (parity of bar_qk / bar_v); S / PV completions so far
  for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
  const int r = item / heads, h = item % heads;
  const int next = item + (int)gridDim.x;
  int any_mask = 0;
  for (int key = tid; key < KP; key += 256) {
    float m = -INFINITY;
    if (key < L) {
      m = (mask_add != nullptr && key < mask_len) ? mask_add[(int64_t)r * mask_ld + key] * LOG2E : 0.f;
      any_mask |= (m != 0.f);
    }
    maskf[key] = m;
  }
  const bool masked = __syncthreads_or(any_mask) != 0;  // block-uniform: the common all-visible case skips the mask reads
  mbar_wait(bar_qk, n_done & 1);

  for (int qt = 0; qt < nqt; ++qt, mma_phase ^= 1) {
    if (tid == 0) {
      tc_fence_after();
      if (SPL) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, umma_desc_sw128(sQ + qt * Q_BYTES + k * 32), umma_desc_sw128(sKl + k * 32), idesc_s, k != 0);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem, umma_desc_sw128(sQl + qt * Q_BYTES + k * 32), umma_desc_sw128(sK + k * 32), idesc_s, 1);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem, umma_desc_sw128(sQ + qt * Q_BYTES + k * 32), umma_desc_sw128(sK + k * 32), idesc_s, SPL || k != 0);
      umma_commit(bar_s);
    }
    mbar_wait(bar_s, mma_phase);
    __syncwarp();
    tc_fence_after();

    // ---- pass 1: row maximum (in the log2 domain) over this thread's half of the keys
    float mx = -INFINITY;
#pragma unroll 1
    for (int j = 0; j < CH; ++j) {
      uint32_t raw[32];
      tmem_ld32(lane_addr + col0 + j * 32, raw);
      const int k0 = col0 + j * 32;
      if (masked || k0 + 32 > L) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaf(__uint_as_float(raw[i]), scale_l2e, maskf[k0 + i]));
      } else {
        float m2 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) m2 = fmaxf(m2, __uint_as_float(raw[i]));
        mx = fmaxf(mx, m2 * scale_l2e);
      }
    }
    red[half * 128 + trow] = mx;
    __syncthreads();
    // Q / K tiles are free (every S of this item is done) AND every thread has passed this item's bar_qk wait (the barrier
    // above): re-arming bar_qk earlier lets it complete a second phase before a late waker has observed the first, and a
    // parity wait that misses a phase never returns (seen as a hang when the loads, not the math, bound an item).
    if (tid == 0 && qt == nqt - 1 && next < n_items) load_qk(next);
    mx = fmaxf(red[trow], red[128 + trow]);

    // ---- pass 2: p = 2^(s*c + mask - max); packed bf16 P kept in registers until every S read is done
    float sum = 0.f;
    uint32_t pk[CH][16];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      uint32_t raw[32];
      uint32_t pl[16];
      tmem_ld32(lane_addr + col0 + j * 32, raw);
      const int k0 = col0 + j * 32;
      if (masked || k0 + 32 > L) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = ex2_approx(fmaf(__uint_as_float(raw[2 * i]), scale_l2e, maskf[k0 + 2 * i]) - mx);
          float p1 = ex2_approx(fmaf(__uint_as_float(raw[2 * i + 1]), scale_l2e, maskf[k0 + 2 * i + 1]) - mx);
          sum += p0 + p1;
          if (DROP) {   // training: drop probabilities (the denominator `sum` stays the full one)
            const uint64_t e0 = ((uint64_t)item * L + (uint64_t)(qt * 128 + trow)) * L + (uint64_t)(k0 + 2 * i);
            p0 *= drop_mul(drop, e0); p1 *= drop_mul(drop, e0 + 1);
          }
          __nv_bfloat162 b = __floats2bfloat162_rn(p0, p1);  // .x (low half) = even key
          pk[j][i] = *reinterpret_cast<uint32_t*>(&b);
          if (SPL) {
            __nv_bfloat162 c = __floats2bfloat162_rn(p0 - __low2float(b), p1 - __high2float(b));
            pl[i] = *reinterpret_cast<uint32_t*>(&c);
          }
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = ex2_approx(fmaf(__uint_as_float(raw[2 * i]), scale_l2e, -mx));
          float p1 = ex2_approx(fmaf(__uint_as_float(raw[2 * i + 1]), scale_l2e, -mx));
          sum += p0 + p1;
          if (DROP) {
            const uint64_t e0 = ((uint64_t)item * L + (uint64_t)(qt * 128 + trow)) * L + (uint64_t)(k0 + 2 * i);
            p0 *= drop_mul(drop, e0); p1 *= drop_mul(drop, e0 + 1);
          }
          __nv_bfloat162 b = __floats2bfloat162_rn(p0, p1);
          pk[j][i] = *reinterpret_cast<uint32_t*>(&b);
          if (SPL) {
            __nv_bfloat162 c = __floats2bfloat162_rn(p0 - __low2float(b), p1 - __high2float(b));
            pl[i] = *reinterpret_cast<uint32_t*>(&c);
          }
        }
      }
      // P_lo has TMEM columns of its own (free since the previous tile's P V completed): written chunk by chunk
      if (SPL) tmem_st16(lane_addr + KP + half * (KP / 4) + j * 16, pl);
    }
    red[256 + half * 128 + trow] = sum;
    tc_fence_before();
    __syncthreads();  // all S columns have been read by both halves: P may now overwrite them
    tc_fence_after();
    sum = red[256 + trow] + red[256 + 128 + trow];
#pragma unroll
    for (int j = 0; j < CH; ++j) tmem_st16(lane_addr + half * (KP / 4) + j * 16, pk[j]);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();

    if (tid == 0) {
      if (qt == 0) mbar_wait(bar_v, n_done & 1);
      tc_fence_after();
      if (SPL) {
#pragma unroll
        for (int kk = 0; kk < KP / 16; ++kk)
          umma_bf16_ts(tmem + KP / 2, tmem + kk * 8, umma_desc_sw128(sVl + kk * 2048), idesc_o, kk != 0);
#pragma unroll
        for (int kk = 0; kk < KP / 16; ++kk)
          umma_bf16_ts(tmem + KP / 2, tmem + KP + kk * 8, umma_desc_sw128(sV + kk * 2048), idesc_o, 1);
      }
#pragma unroll
      for (int kk = 0; kk < KP / 16; ++kk)
        umma_bf16_ts(tmem + KP / 2, tmem + kk * 8, umma_desc_sw128(sV + kk * 2048), idesc_o, SPL || kk != 0);
      umma_commit(bar_o);
    }
    mbar_wait(bar_o, mma_phase);
    __syncwarp();
    tc_fence_after();
    if (tid == 0 && qt == nqt - 1 && next < n_items) load_v(next);    // V is free: the last P V of this item is done

    // ---- O = (P V) / sum : this thread stores 32 of the row's 64 output columns
    const int q = qt * 128 + trow;
    const float inv = 1.0f / sum;
    // training forward: the row's natural-log sum of exponentials of (scale s + mask), what the backward kernels subtract
    // (mx and the exponentials live in the log2 domain here)
    if (lse_out != nullptr && half == 0 && q < L) lse_out[(int64_t)item * L + q] = (mx + log2f(sum)) * 0.69314718055994530942f;
    {
      uint32_t raw[32];
      tmem_ld32(lane_addr + KP / 2 + half * 32, raw);
      if (SPL) {
        if (q < L) {   // ctx row = [hi(heads*64) | lo(heads*64)]
          bf16* op = ctx + ((int64_t)r * L + q) * (2 * heads * AT_D) + h * AT_D + half * 32;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            uint2 hi, lo;
            split4(make_float4(__uint_as_float(raw[4 * i]) * inv, __uint_as_float(raw[4 * i + 1]) * inv,
                               __uint_as_float(raw[4 * i + 2]) * inv, __uint_as_float(raw[4 * i + 3]) * inv), hi, lo);
            *reinterpret_cast<uint2*>(op + 4 * i) = hi;
            *reinterpret_cast<uint2*>(op + heads * AT_D + 4 * i) = lo;
          }
        }
      } else if (q < L) {
        bf16* op = ctx + ((int64_t)r * L + q) * (heads * AT_D) + h * AT_D + half * 32;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          __nv_bfloat162 a = __floats2bfloat162_rn(__uint_as_float(raw[8 * i]) * inv, __uint_as_float(raw[8 * i + 1]) * inv);
          __nv_bfloat162 b = __floats2bfloat162_rn(__uint_as_float(raw[8 * i + 2]) * inv, __uint_as_float(raw[8 * i + 3]) * inv);
          __nv_bfloat162 c = __floats2bfloat162_rn(__uint_as_float(raw[8 * i + 4]) * inv, __uint_as_float(raw[8 * i + 5]) * inv);
          __nv_bfloat162 d = __floats2bfloat162_rn(__uint_as_float(raw[8 * i + 6]) * inv, __uint_as_float(raw[8 * i + 7]) * inv);
          u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
          u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
          *reinterpret_cast<uint4*>(op + 8 * i) = u;
        }
      }
    }
    tc_fence_before();
    __syncthreads();  // every row has drained S/P/O before the next tile's MMAs overwrite them
  }
  }
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(Cfg::TMEM_COLS) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------
// Warp-specialised form for split (bf16x3) operands: one persistent CTA per SM, 14 warps.
//   warp 0      TMA producer: Q / K / V tiles (hi and lo planes) of the NEXT item as soon as their slots are released
//   warp 1      MMA issuer (one thread) + TMEM allocation
//   warps 2-5   softmax group 0, warps 6-9 softmax group 1 (one thread per query row, a 128-key block each)
//   warps 10-13 epilogue group: combines the key blocks of a query tile and stores the context rows (hi | lo)
// The work of a CTA is a sequence of BLOCKS (item, query tile of 128 rows, key block of 128 keys).  S of block n lives in
// TMEM buffer n & 1 (128 fp32 columns); group n & 1 turns it into P (hi and lo bf16 planes written over the S columns of
// the same 32-key chunk, so nothing is held in registers) with the block's OWN row maximum; O_kb = P_kb V_kb accumulates
// into its own 64 TMEM columns and the epilogue forms (O_0 2^(m_0-m) + O_1 2^(m_1-m)) / (l_0 2^(m_0-m) + l_1 2^(m_1-m)) --
// the exact softmax, with no rescaling pass over TMEM.  The issue order S(n+2) after PV(n) (tcgen05.mma executes in issue
// order) is what frees an S buffer, so the tensor pipe runs S(n+1) / PV(n-1) while group n & 1 is in its exponentials:
// the serial chain S -> softmax -> PV of attention_tc_kernel (one tile in flight, 20 % tensor pipe) becomes a pipeline.
// TMEM columns: S0 [0,128) | S1 [128,256) | O[tile parity][key block] 4 x 64 at [256,512).
// ---------------------------------------------------------------------------------------------------
// SPL = false: plain bf16 operands (one MMA pass, one plane): an item is half as large, so TWO items are resident at every
// sequence length and the next item's tiles are always in flight while the current one is computed.
template <int KP, bool SPL = true> struct AttnWsCfg {
  static constexpr int NKB = KP / 128;                 // key blocks per item
  static constexpr int QT = KP / 128;                  // query tiles resident per item
  static constexpr int PLANES = SPL ? 2 : 1;           // hi plane (, lo plane)
  static constexpr int STAGES = SPL ? 256 / KP : 2;    // items resident in shared memory
  static constexpr int TILE = 128 * 64 * 2;            // 128 rows x 64 dims of bf16, 128B swizzle
  static constexpr int PLANE = (QT + 2 * NKB) * TILE;  // Q tiles | K blocks | V blocks of one plane
  static constexpr int STAGE_BYTES = PLANES * PLANE;
  static constexpr int OPER = STAGES * STAGE_BYTES;    // <= 192 KB
  static constexpr int STG_OFF = OPER;                 // epilogue staging: one 32-row x 128-byte box per epilogue warp (1024-aligned)
  static constexpr int MASK_OFF = STG_OFF + 4 * 4096, STATS_OFF = MASK_OFF + 2 * 128 * 4, BAR_OFF = STATS_OFF + 4 * 2 * 128 * 8;
  static constexpr int SMEM = BAR_OFF + 256 + 1024;
  static constexpr int THREADS = 14 * 32;
};

template <int KP, bool SPL, bool DROP>
__global__ void __launch_bounds__(AttnWsCfg<KP, SPL>::THREADS, 1)
attention_ws_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_kv,
                    const __grid_constant__ CUtensorMap map_ctx, int L, int heads,
                    float scale_l2e, const float* __restrict__ mask_add, int mask_ld, int mask_len, bf16* __restrict__ ctx,
                    int n_items, long long* trace, Drop drop) {
  using Cfg = AttnWsCfg<KP, SPL>;
  static_assert(!(SPL && DROP), "dropout is a training-path (bf16) feature");
  constexpr int PLANES = Cfg::PLANES;
#ifdef MSQ_ATTN_TRACE
  // event trace of CTA 0 (development aid): trace[role][n][k] = clock64 at event k of block / tile n
#define TR(role, n, k) do { if (blockIdx.x == 0 && (n) < 64) trace[((role) * 64 + (n)) * 8 + (k)] = clock64(); } while (0)
#else
#define TR(role, n, k) do { } while (0)
#endif
  constexpr int NKB = Cfg::NKB, QT = Cfg::QT, STAGES = Cfg::STAGES, TILE = Cfg::TILE, PLANE = Cfg::PLANE;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* maskf = reinterpret_cast<float*>(gen + Cfg::MASK_OFF);          // [2 groups][128]
  float2* stats = reinterpret_cast<float2*>(gen + Cfg::STATS_OFF);       // [4 tile slots][2 key blocks][128 rows] (max, sum)
  const uint32_t bars = base + Cfg::BAR_OFF;
  // 8-byte barriers: qk_full[2] | qk_empty[2] | v_full[2] | v_empty[2] | s_full[2] | p_ready[2] | o_full[2] | o_empty[2] | slot
  const uint32_t qk_full = bars, qk_empty = bars + 16, v_full = bars + 32, v_empty = bars + 48, s_full = bars + 64,
                 p_ready = bars + 80, o_full = bars + 96, o_empty = bars + 112, tslot = bars + 128;
  volatile uint32_t* tslot_gen = reinterpret_cast<volatile uint32_t*>(gen + Cfg::BAR_OFF + 128);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nqt = (L + 127) >> 7;
  const int my_items = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int NT = my_items * nqt, NB = NT * NKB;   // query tiles / blocks of this CTA

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ctx) : "memory");
    for (int i = 0; i < 2; ++i) {
      mbar_init(qk_full + 8 * i, 1); mbar_init(qk_empty + 8 * i, 1); mbar_init(v_full + 8 * i, 1); mbar_init(v_empty + 8 * i, 1);
      mbar_init(s_full + 8 * i, 1); mbar_init(p_ready + 8 * i, 4); mbar_init(o_full + 8 * i, 1); mbar_init(o_empty + 8 * i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_sync();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot_gen;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int lo = 3 * heads * AT_D;   // the lo plane of a qkv row starts here
      int item = blockIdx.x;
      for (int li = 0; li < my_items; ++li, item += gridDim.x) {
        const int s = li % STAGES, u = li / STAGES;
        const int row0 = (item / heads) * L, hh = item % heads;
        const uint32_t st0 = base + s * Cfg::STAGE_BYTES;
        mbar_wait(qk_empty + 8 * s, (u & 1) ^ 1);
        mbar_expect_tx(qk_full + 8 * s, (uint32_t)(PLANES * (nqt + NKB) * TILE));
#pragma unroll
        for (int pl = 0; pl < PLANES; ++pl) {
          tma_load_2d(st0 + pl * PLANE + QT * TILE, &map_kv, pl * lo + heads * AT_D + hh * AT_D, row0, qk_full + 8 * s);
          for (int qt = 0; qt < nqt; ++qt)
            tma_load_2d(st0 + pl * PLANE + qt * TILE, &map_q, pl * lo + hh * AT_D, row0 + qt * 128, qk_full + 8 * s);
        }
        mbar_wait(v_empty + 8 * s, (u & 1) ^ 1);
        mbar_expect_tx(v_full + 8 * s, (uint32_t)(PLANES * NKB * TILE));
#pragma unroll
        for (int pl = 0; pl < PLANES; ++pl)
          tma_load_2d(st0 + pl * PLANE + (QT + NKB) * TILE, &map_kv, pl * lo + 2 * heads * AT_D + hh * AT_D, row0, v_full + 8 * s);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the block sequence (uniform control flow, every lane waits on the barriers); ONE elected lane
    // issues.  elect.sync (instead of `lane == 0`) lets ptxas keep descriptors in uniform registers without a per-MMA
    // waterfall loop, and descriptors are advanced by adding to their low word: at ~70 cycles of issue overhead per
    // tcgen05.mma the 32-cycle N = 64 MMAs of P V were issue-bound.
    {
      // instruction descriptors: D=F32, A=B=BF16; S: N=128, A/B K-major; PV: N=64, B MN-major (bit 16)
      const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) | (8u << 24);
      const uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(64 >> 3) << 17) | (8u << 24);
      const uint64_t dk0 = umma_desc_sw128(base), dv0 = umma_desc_sw128_mn(base);   // descriptors of the stage base
      auto issue_s = [&](int n) {
        const int t = n / NKB, kb = n % NKB, li = t / nqt, qt = t - li * nqt, s = li % STAGES, u = li / STAGES;
        if (qt == 0 && kb == 0) { mbar_wait(qk_full + 8 * s, u & 1); tc_fence_after(); }
        const uint32_t d = tmem + (n & 1) * 128;
        // byte offsets >> 4 from the stage base: Q tile qt, K block kb (hi planes; lo planes PLANE further)
        const uint64_t qh = dk0 + (uint64_t)((s * Cfg::STAGE_BYTES + qt * TILE) >> 4), ql = qh + (PLANE >> 4);
        const uint64_t kh = dk0 + (uint64_t)((s * Cfg::STAGE_BYTES + (QT + kb) * TILE) >> 4), kl = kh + (PLANE >> 4);
        if (elect_one()) {
          if (SPL) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_acc<false>(d, qh + 2 * k, kl + 2 * k, idesc_s, k != 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_bf16_acc<false>(d, ql + 2 * k, kh + 2 * k, idesc_s, true);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_acc<false>(d, qh + 2 * k, kh + 2 * k, idesc_s, SPL || k != 0);
          umma_commit(s_full + 8 * (n & 1));
          if (qt == nqt - 1 && kb == NKB - 1) umma_commit(qk_empty + 8 * s);   // Q / K of this item have been read
        }
        __syncwarp();
      };
      auto issue_pv = [&](int n) {
        const int t = n / NKB, kb = n % NKB, li = t / nqt, qt = t - li * nqt, s = li % STAGES, u = li / STAGES;
        if (qt == 0 && kb == 0) mbar_wait(v_full + 8 * s, u & 1);
        if (kb == 0) mbar_wait(o_empty + 8 * (t & 1), ((t >> 1) & 1) ^ 1);   // the epilogue has drained this O buffer
        if (lane == 0) TR(0, n, 0);
        mbar_wait(p_ready + 8 * (n & 1), (n >> 1) & 1);
        if (lane == 0) TR(0, n, 1);
        tc_fence_after();
        const uint32_t p = tmem + (n & 1) * 128, d = tmem + 256 + ((t & 1) * 2 + kb) * 64;
        const uint64_t vh = dv0 + (uint64_t)((s * Cfg::STAGE_BYTES + (QT + NKB + kb) * TILE) >> 4), vl = vh + (PLANE >> 4);
        // P of 32-key chunk j: hi plane in columns [32 j, 32 j + 16), lo plane in [32 j + 16, 32 j + 32); 16 keys = 8 columns;
        // 16 keys of V (MN-major) = 2048 bytes
        if (elect_one()) {
          if (SPL) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_bf16_acc<true>(d, (uint64_t)(p + 32 * (kk >> 1) + 8 * (kk & 1)), vl + 128 * kk, idesc_o, kk != 0);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_bf16_acc<true>(d, (uint64_t)(p + 32 * (kk >> 1) + 16 + 8 * (kk & 1)), vh + 128 * kk, idesc_o, true);
          }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            umma_bf16_acc<true>(d, (uint64_t)(p + 32 * (kk >> 1) + 8 * (kk & 1)), vh + 128 * kk, idesc_o, SPL || kk != 0);
          if (kb == NKB - 1) umma_commit(o_full + 8 * (t & 1));
          if (qt == nqt - 1 && kb == NKB - 1) umma_commit(v_empty + 8 * s);    // V of this item has been read
        }
        __syncwarp();
        if (lane == 0) TR(0, n, 2);
      };
      issue_s(0);
      if (NB > 1) issue_s(1);
      for (int n = 0; n < NB; ++n) {
        issue_pv(n);
        if (n + 2 < NB) issue_s(n + 2);   // after PV(n) in issue order: S buffer n & 1 (and its P) is free by then
        if (lane == 0) TR(0, n, 3);
      }
    }
  } else if (warp < 10) {
    // ===================== softmax groups =====================
    const int w = (warp - 2) >> 2;                       // group: blocks with n & 1 == w
    const int quarter = warp & 3, row = quarter * 32 + lane, gtid = ((warp - 2) & 3) * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    float* mk = maskf + w * 128;
    const float LOG2E = 1.4426950408889634f;
    int have_item = -1, have_kb = -1;
    bool masked = false;
    for (int n = w; n < NB; n += 2) {
      const int t = n / NKB, kb = n % NKB, li = t / nqt;
      const int item = (int)blockIdx.x + li * (int)gridDim.x, r = item / heads;
      const int key0 = kb * 128;
      if (item != have_item || kb != have_kb) {
        // additive key mask of this block (log2 domain), -inf beyond the sequence; the all-visible case skips the reads
        named_bar_sync(1 + w, 128);   // the previous block's readers are done
        const int key = key0 + gtid;
        float m = -INFINITY;
        bool nz = false;
        if (key < L) {
          m = (mask_add != nullptr && key < mask_len) ? mask_add[(int64_t)r * mask_ld + key] * LOG2E : 0.f;
          nz = m != 0.f;
        }
        mk[gtid] = m;
        masked = named_bar_or(1 + w, 128, nz);
        have_item = item; have_kb = kb;
      }
      const uint32_t sb = lane_addr + (n & 1) * 128;
      if (gtid == 0) TR(1, n, 0);
      mbar_wait(s_full + 8 * (n & 1), (n >> 1) & 1);
      if (gtid == 0) TR(1, n, 1);
      tc_fence_after();
      // ---- pass 1: row maximum of the block (log2 domain)
      float mx = -INFINITY;
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        uint32_t raw[32];
        tmem_ld32(sb + j * 32, raw);
        if (masked || key0 + j * 32 + 32 > L) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, fmaf(__uint_as_float(raw[i]), scale_l2e, mk[j * 32 + i]));
        } else {
          float m2 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; ++i) m2 = fmaxf(m2, __uint_as_float(raw[i]));
          mx = fmaxf(mx, m2 * scale_l2e);
        }
      }
      if (gtid == 0) TR(1, n, 2);
      // ---- pass 2: p = 2^(s c + mask - max), hi / lo planes written over the chunk's own S columns
      float sum = 0.f;
      const int qrow = (t % nqt) * 128 + row;   // query index inside the item (dropout element index)
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        uint32_t raw[32], ph[16], pl[16];
        tmem_ld32(sb + j * 32, raw);
        const bool slow = masked || key0 + j * 32 + 32 > L;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0, p1;
          if (slow) {
            p0 = ex2_approx(fmaf(__uint_as_float(raw[2 * i]), scale_l2e, mk[j * 32 + 2 * i]) - mx);
            p1 = ex2_approx(fmaf(__uint_as_float(raw[2 * i + 1]), scale_l2e, mk[j * 32 + 2 * i + 1]) - mx);
          } else {
            p0 = ex2_approx(fmaf(__uint_as_float(raw[2 * i]), scale_l2e, -mx));
            p1 = ex2_approx(fmaf(__uint_as_float(raw[2 * i + 1]), scale_l2e, -mx));
          }
          sum += p0 + p1;
          if (DROP) {   // training: drop probabilities (the denominator `sum` stays the full one); same element index as
                        // attention_tc_kernel / the backward kernels: ((item L + query) L + key)
            const uint64_t e0 = ((uint64_t)item * L + (uint64_t)qrow) * L + (uint64_t)(key0 + j * 32 + 2 * i);
            p0 *= drop_mul(drop, e0); p1 *= drop_mul(drop, e0 + 1);
          }
          const __nv_bfloat162 b = __floats2bfloat162_rn(p0, p1);  // .x (low half) = even key
          ph[i] = *reinterpret_cast<const uint32_t*>(&b);
          if (SPL) {
            const __nv_bfloat162 c = __floats2bfloat162_rn(p0 - __low2float(b), p1 - __high2float(b));
            pl[i] = *reinterpret_cast<const uint32_t*>(&c);
          }
        }
        tmem_st16(sb + j * 32, ph);
        if (SPL) tmem_st16(sb + j * 32 + 16, pl);
      }
      stats[((t & 3) * 2 + kb) * 128 + row] = make_float2(mx, sum);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready + 8 * (n & 1));
      if (gtid == 0) TR(1, n, 3);
    }
  } else {
    // ===================== epilogue group =====================
    const int quarter = warp & 3, row = quarter * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    for (int t = 0; t < NT; ++t) {
      const int li = t / nqt, qt = t % nqt;
      const int item = (int)blockIdx.x + li * (int)gridDim.x, r = item / heads, h = item % heads;
      if (tid == 320) TR(2, t, 0);
      mbar_wait(o_full + 8 * (t & 1), (t >> 1) & 1);
      if (tid == 320) TR(2, t, 1);
      tc_fence_after();
      const float2 s0 = stats[((t & 3) * 2 + 0) * 128 + row];
      float w0 = 1.f, w1 = 0.f, den = s0.y;
      if (NKB == 2) {
        const float2 s1 = stats[((t & 3) * 2 + 1) * 128 + row];
        const float m = fmaxf(s0.x, s1.x);
        w0 = ex2_approx(s0.x - m); w1 = ex2_approx(s1.x - m);
        den = fmaf(s0.y, w0, s1.y * w1);
      }
      const float inv = 1.0f / den;
      w0 *= inv; w1 *= inv;
      const uint32_t ob = lane_addr + 256 + (t & 1) * 128;
      float o[64];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t a[32];
        tmem_ld32(ob + hf * 32, a);
#pragma unroll
        for (int i = 0; i < 32; ++i) o[hf * 32 + i] = __uint_as_float(a[i]) * w0;
        if (NKB == 2) {
          tmem_ld32(ob + 64 + hf * 32, a);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[hf * 32 + i] = fmaf(__uint_as_float(a[i]), w1, o[hf * 32 + i]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_empty + 8 * (t & 1));   // O buffer drained: the tile after next may accumulate into it
      // ctx rows = [hi(heads*64) | lo(heads*64)] (plain bf16: the hi plane only): this warp's 32 rows x 64 columns go out as
      // TMA boxes staged in shared memory (128B swizzle; piece j of row r at r*128 + ((j ^ (r & 7)) << 4)).  Per-thread 16-byte
      // stores of whole rows (32 different lines per instruction) had flooded the memory pipeline the softmax warps' TMEM loads
      // and the barrier polls share.  The 3-D map clips the box at the end of the token group (rows >= L are not written).
      const int q0 = qt * 128 + quarter * 32;
      const uint32_t stg = base + Cfg::STG_OFF + (warp - 10) * 4096 + lane * 128;
      if (lane == 0) tma_store_wait_read<0>();   // the previous tile's last box has been read out of the staging buffer
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 8; ++i) {   // hi plane: round-to-nearest bf16 of the normalised outputs
        const __nv_bfloat162 a0 = __floats2bfloat162_rn(o[8 * i], o[8 * i + 1]), a1 = __floats2bfloat162_rn(o[8 * i + 2], o[8 * i + 3]);
        const __nv_bfloat162 a2 = __floats2bfloat162_rn(o[8 * i + 4], o[8 * i + 5]), a3 = __floats2bfloat162_rn(o[8 * i + 6], o[8 * i + 7]);
        st_shared_v4(stg + ((i ^ (lane & 7)) << 4), *reinterpret_cast<const uint32_t*>(&a0), *reinterpret_cast<const uint32_t*>(&a1),
                     *reinterpret_cast<const uint32_t*>(&a2), *reinterpret_cast<const uint32_t*>(&a3));
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0 && q0 < L) { tma_store_3d(&map_ctx, stg - lane * 128, h * AT_D, q0, r); tma_store_commit(); }
      if (SPL) {
        // lo plane: computed while the TMA reads the hi box
        uint4 lo4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint2 h0, l0, h1, l1;
          split4(make_float4(o[8 * i], o[8 * i + 1], o[8 * i + 2], o[8 * i + 3]), h0, l0);
          split4(make_float4(o[8 * i + 4], o[8 * i + 5], o[8 * i + 6], o[8 * i + 7]), h1, l1);
          lo4[i] = make_uint4(l0.x, l0.y, l1.x, l1.y);
        }
        if (lane == 0) tma_store_wait_read<0>();
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) st_shared_v4(stg + ((i ^ (lane & 7)) << 4), lo4[i].x, lo4[i].y, lo4[i].z, lo4[i].w);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && q0 < L) { tma_store_3d(&map_ctx, stg - lane * 128, heads * AT_D + h * AT_D, q0, r); tma_store_commit(); }
      }
      if (tid == 320) TR(2, t, 2);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the last boxes are in global memory
  }
#undef TR
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

template <int KP, bool SPL>
static int launch_attention_ws(const bf16* qkv, int64_t R, int L, int heads, float scale, const float* mask_add, int mask_ld,
                               int mask_len, bf16* ctx, cudaStream_t st, const Drop& drop = Drop()) {
  using Cfg = AttnWsCfg<KP, SPL>;
  constexpr int PL = Cfg::PLANES;
  CUtensorMap mq, mkv;
  const int ld = PL * 3 * heads * AT_D;   // bf16 elements per qkv row (hi | lo)
  MSQ_TRY(make_map_bf16(&mq, qkv, R * L, ld, ld, AT_D, 128));
  MSQ_TRY(make_map_bf16(&mkv, qkv, R * L, ld, ld, AT_D, KP));
  CUtensorMap mctx;
  MSQ_TRY(make_map_3d_bf16(&mctx, ctx, R, L, PL * heads * AT_D, PL * heads * AT_D, AT_D, 32));
  const bool dropping = !SPL && drop.thresh != 0;
  if (dropping) MSQ_SMEM_ATTR(Cfg::SMEM, (attention_ws_kernel<KP, SPL, !SPL>));
  else MSQ_SMEM_ATTR(Cfg::SMEM, (attention_ws_kernel<KP, SPL, false>));
  int dev = 0, sms = 0;
  MSQ_CUDA(cudaGetDevice(&dev));
  MSQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_items = R * heads;
  MSQ_REQUIRE(n_items < ((int64_t)1 << 28), "attention: too many (row, head) items");
  long long* trace = nullptr;
#ifdef MSQ_ATTN_TRACE
  static long long* trace_dev = nullptr;
  if (!trace_dev) MSQ_CUDA(cudaMalloc(&trace_dev, 3 * 64 * 8 * sizeof(long long)));
  MSQ_CUDA(cudaMemsetAsync(trace_dev, 0, 3 * 64 * 8 * sizeof(long long), st));
  trace = trace_dev;
#endif
  const dim3 grid((unsigned)min((int64_t)sms, n_items));
  if (dropping)
    MSQ_CUDA(launch_k(attention_ws_kernel<KP, SPL, !SPL>, grid, dim3(Cfg::THREADS), Cfg::SMEM, st, mq, mkv, mctx, L, heads,
                      scale * 1.4426950408889634f, mask_add, mask_ld, mask_len, ctx, (int)n_items, trace, drop));
  else
    MSQ_CUDA(launch_k(attention_ws_kernel<KP, SPL, false>, grid, dim3(Cfg::THREADS), Cfg::SMEM, st, mq, mkv, mctx, L, heads,
                      scale * 1.4426950408889634f, mask_add, mask_ld, mask_len, ctx, (int)n_items, trace, drop));
  MSQ_LAUNCH_CHECK();
#ifdef MSQ_ATTN_TRACE
  {
    static long long host[3 * 64 * 8];
    MSQ_CUDA(cudaStreamSynchronize(st));
    MSQ_CUDA(cudaMemcpy(host, trace_dev, sizeof(host), cudaMemcpyDeviceToHost));
    if (FILE* f = fopen("gpurun_out/attn_trace.bin", "wb")) { fwrite(host, 1, sizeof(host), f); fclose(f); }
  }
#endif
  return MSQ_OK;
}
static bool attention_use_ws() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MSQ_ATTN_WS"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

template <int KP, bool SPL>
static int launch_attention_tc(const bf16* qkv, int64_t R, int L, int heads, float scale, const float* mask_add, int mask_ld,
                               int mask_len, bf16* ctx, cudaStream_t st, const Drop& drop = Drop(), float* lse_out = nullptr) {
  CUtensorMap mq, mkv;
  const int ld = (SPL ? 2 : 1) * 3 * heads * AT_D;   // bf16 elements per qkv row
  MSQ_TRY(make_map_bf16(&mq, qkv, R * L, ld, ld, AT_D, 128));
  MSQ_TRY(make_map_bf16(&mkv, qkv, R * L, ld, ld, AT_D, KP));
  constexpr int SMEM = AttnCfg<KP, SPL>::SMEM;
  // dropout is a compile-time variant: its hash code inside the unrolled exp loops doubled the evaluation kernel's time
  // (instruction footprint) when it was a run-time branch
  const bool dropping = !SPL && drop.thresh != 0;
  if (dropping) MSQ_SMEM_ATTR(SMEM, (attention_tc_kernel<KP, SPL, !SPL>));
  else MSQ_SMEM_ATTR(SMEM, (attention_tc_kernel<KP, SPL, false>));
  // Persistent grid = the CTAs that are resident at once: 2 per SM at KP = 256 (256 TMEM columns and 100 KB of shared
  // memory each), 3 per SM at KP = 128 (68 KB, 80 registers).  Measured (640 x 12 items): 296 CTAs 0.289 ms, 444 CTAs 0.365 ms
  // (a wave and a half), one CTA per item 0.335 ms; L = 99: 444 CTAs 0.101 ms vs 0.116 ms.
  static int resident = 0;
  if (!resident) {
    int dev = 0, sms = 0;
    MSQ_CUDA(cudaGetDevice(&dev));
    MSQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    resident = (SPL ? (KP == 256 ? 1 : 2) : (KP == 256 ? 2 : 3)) * sms;   // SPL: 197 KB / 100 KB of shared memory, 512 / 256 TMEM columns
    if (const char* e = getenv("MSQ_ATTN_CTAS_PER_SM")) resident = max(1, atoi(e)) * sms;
  }
  const int64_t n_items = R * heads;
  MSQ_REQUIRE(n_items < ((int64_t)1 << 31), "attention: too many (row, head) items");

  const dim3 grid((unsigned)min((int64_t)resident, n_items));
  if (dropping) MSQ_CUDA(launch_k(attention_tc_kernel<KP, SPL, !SPL>, grid, dim3(256), SMEM, st, mq, mkv, L, heads, scale * 1.4426950408889634f, mask_add, mask_ld, mask_len, ctx, (int)n_items, drop, lse_out));
  else MSQ_CUDA(launch_k(attention_tc_kernel<KP, SPL, false>, grid, dim3(256), SMEM, st, mq, mkv, L, heads, scale * 1.4426950408889634f, mask_add, mask_ld, mask_len, ctx, (int)n_items, drop, lse_out));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

static bool attention_use_tc() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MSQ_ATTN_SIMT"); v = (e && e[0] == '1') ? 0 : 1; }
  return v && tc_supported_impl();
}

template <typename T>
int attention(const T* qkv, int64_t R, int L, int heads, int dhead, float scale, const float* key_mask_add, int mask_ld,
              int mask_len, T* ctx, cudaStream_t st, const Drop& drop, float* lse_out, bool* lse_written) {
  if (lse_written) *lse_written = false;
  MSQ_REQUIRE(dhead == AT_D, "attention: head dim %d != 64", dhead);
  MSQ_REQUIRE(L >= 1 && L <= 320, "attention: sequence length %d out of range", L);
  if (R == 0) return MSQ_OK;
  if constexpr (is_split<T>::value) {
    MSQ_REQUIRE(drop.thresh == 0, "attention: dropout is a training-path feature (bf16 / fp32), not available in the bf16x3 mode");
    MSQ_REQUIRE(tc_supported_impl() && L <= 256 && (((uintptr_t)qkv) & 15) == 0, "attention: the bf16x3 mode needs the tcgen05 path (sm_100, L <= 256)");
    if (attention_use_ws()) {   // warp-specialised pipeline (MSQ_ATTN_WS=0: the one-tile-at-a-time kernel)
      if (L <= 128)
        return launch_attention_ws<128, true>((const bf16*)qkv, R, L, heads, scale, key_mask_add, mask_ld, mask_len, (bf16*)ctx, st);
      return launch_attention_ws<256, true>((const bf16*)qkv, R, L, heads, scale, key_mask_add, mask_ld, mask_len, (bf16*)ctx, st);
    }
    if (L <= 128)
      return launch_attention_tc<128, true>((const bf16*)qkv, R, L, heads, scale, key_mask_add, mask_ld, mask_len, (bf16*)ctx, st);
    return launch_attention_tc<256, true>((const bf16*)qkv, R, L, heads, scale, key_mask_add, mask_ld, mask_len, (bf16*)ctx, st);
  } else {
  if constexpr (sizeof(T) == 2) {
    // plain bf16 operands: the warp-specialised pipeline is available (MSQ_ATTN_WS_BF16=1) but NOT the default: with one MMA
    // pass the exponentials, not the tensor pipe, bound the kernel, and two co-resident CTAs of attention_tc_kernel put 512
    // threads on them against the 256 of the two softmax groups -- measured 282 vs 367 us (L = 227), 105 vs 102 us (L = 99)
    static int ws16 = -1;
    if (ws16 < 0) { const char* e = getenv("MSQ_ATTN_WS_BF16"); ws16 = (e && e[0] == '1') ? 1 : 0; }
    if (ws16 && attention_use_tc() && attention_use_ws() && L <= 256 && (((uintptr_t)qkv) & 15) == 0) {
      if (L <= 128)
        return launch_attention_ws<128, false>((const bf16*)qkv, R, L, heads, scale, key_mask_add, mask_ld, mask_len, (bf16*)ctx, st, drop);
      return launch_attention_ws<256, false>((const bf16*)qkv, R, L, heads, scale, key_mask_add, mask_ld, mask_len, (bf16*)ctx, st, drop);
    }
    if (attention_use_tc() && L <= 256 && (((uintptr_t)qkv) & 15) == 0) {
      if (lse_written) *lse_written = lse_out != nullptr;   // this kernel emits the row log-sum-exp for the backward pass
      if (L <= 128)
        return launch_attention_tc<128, false>((const bf16*)qkv, R, L, heads, scale, key_mask_add, mask_ld, mask_len, (bf16*)ctx, st, drop, lse_out);
      return launch_attention_tc<256, false>((const bf16*)qkv, R, L, heads, scale, key_mask_add, mask_ld, mask_len, (bf16*)ctx, st, drop, lse_out);
    }
  }
  const int Lpad = (L + 31) & ~31;
  const size_t smem = sizeof(float) * ((size_t)L * 65 + (size_t)L * AT_D + ((L + 3) & ~3) + AT_WARPS * AT_D + AT_WARPS * Lpad);
  MSQ_SMEM_ATTR(smem, attention_simt_kernel<T>);
  MSQ_CUDA(launch_k(attention_simt_kernel<T>, dim3((unsigned)(R * heads)), dim3(AT_WARPS * 32), smem, st, qkv, L, heads, scale, key_mask_add, mask_ld, mask_len, ctx, drop));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
  }
}
template int attention<bf16s>(const bf16s*, int64_t, int, int, int, float, const float*, int, int, bf16s*, cudaStream_t, const Drop&, float*, bool*);
template int attention<float>(const float*, int64_t, int, int, int, float, const float*, int, int, float*, cudaStream_t, const Drop&, float*, bool*);
template int attention<bf16>(const bf16*, int64_t, int, int, int, float, const float*, int, int, bf16*, cudaStream_t, const Drop&, float*, bool*);

}  // namespace msq
