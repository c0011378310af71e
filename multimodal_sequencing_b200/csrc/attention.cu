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
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
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
                                                           int mask_len, bf16* __restrict__ ctx, int n_items, Drop drop) {
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

template <int KP, bool SPL>
static int launch_attention_tc(const bf16* qkv, int64_t R, int L, int heads, float scale, const float* mask_add, int mask_ld,
                               int mask_len, bf16* ctx, cudaStream_t st, const Drop& drop = Drop()) {
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
  if (dropping) MSQ_CUDA(launch_k(attention_tc_kernel<KP, SPL, !SPL>, grid, dim3(256), SMEM, st, mq, mkv, L, heads, scale * 1.4426950408889634f, mask_add, mask_ld, mask_len, ctx, (int)n_items, drop));
  else MSQ_CUDA(launch_k(attention_tc_kernel<KP, SPL, false>, grid, dim3(256), SMEM, st, mq, mkv, L, heads, scale * 1.4426950408889634f, mask_add, mask_ld, mask_len, ctx, (int)n_items, drop));
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
              int mask_len, T* ctx, cudaStream_t st, const Drop& drop) {
  MSQ_REQUIRE(dhead == AT_D, "attention: head dim %d != 64", dhead);
  MSQ_REQUIRE(L >= 1 && L <= 320, "attention: sequence length %d out of range", L);
  if (R == 0) return MSQ_OK;
  if constexpr (is_split<T>::value) {
    MSQ_REQUIRE(drop.thresh == 0, "attention: dropout is a training-path feature (bf16 / fp32), not available in the bf16x3 mode");
    MSQ_REQUIRE(tc_supported_impl() && L <= 256 && (((uintptr_t)qkv) & 15) == 0, "attention: the bf16x3 mode needs the tcgen05 path (sm_100, L <= 256)");
    if (L <= 128)
      return launch_attention_tc<128, true>((const bf16*)qkv, R, L, heads, scale, key_mask_add, mask_ld, mask_len, (bf16*)ctx, st);
    return launch_attention_tc<256, true>((const bf16*)qkv, R, L, heads, scale, key_mask_add, mask_ld, mask_len, (bf16*)ctx, st);
  } else {
  if constexpr (sizeof(T) == 2) {
    if (attention_use_tc() && L <= 256 && (((uintptr_t)qkv) & 15) == 0) {
      if (L <= 128)
        return launch_attention_tc<128, false>((const bf16*)qkv, R, L, heads, scale, key_mask_add, mask_ld, mask_len, (bf16*)ctx, st, drop);
      return launch_attention_tc<256, false>((const bf16*)qkv, R, L, heads, scale, key_mask_add, mask_ld, mask_len, (bf16*)ctx, st, drop);
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
template int attention<bf16s>(const bf16s*, int64_t, int, int, int, float, const float*, int, int, bf16s*, cudaStream_t, const Drop&);
template int attention<float>(const float*, int64_t, int, int, int, float, const float*, int, int, float*, cudaStream_t, const Drop&);
template int attention<bf16>(const bf16*, int64_t, int, int, int, float, const float*, int, int, bf16*, cudaStream_t, const Drop&);

}  // namespace msq
