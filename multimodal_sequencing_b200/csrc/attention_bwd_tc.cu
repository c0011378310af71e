// Attention backward, dK / dV half, on tcgen05 (bf16 fine-tuning path): the warp-specialised pipeline of
// attention_ws_kernel (attention.cu) with keys on the TMEM lanes.
//
// Reference: autograd through BertAttention.forward  models/CLIP/src/lxrt/modeling.py:398-425 and nn.MultiheadAttention
// inside ResidualAttentionBlock  models/CLIP/clip/model.py:204-226 (scores / sqrt(d) + mask, softmax, dropout, P V).
//
// Work of a CTA = a sequence of blocks (item = (pair row, head), 128-key block kb, 64-query block qb):
//   S^T  = K_kb Q_qb^T   and   dP^T = V_kb dO_qb^T        two SS MMA chains into TMEM buffer n & 1 (64 + 64 fp32 columns),
//   one thread per KEY row turns them into  P^T = exp(s c + mask - lse_q) . drop  and  dS^T = P (dP . drop - D_q)  (lse and
//   D = rowsum(dO . O) per query come from the dQ kernel, attention_bwd_mma.cu), written as packed bf16 over the S^T columns of
//   the same 32-query chunk (P^T | dS^T planes, the layout of the forward kernel's hi | lo planes),
//   dV_kb += P^T dO_qb   and   dK_kb += dS^T Q_qb          two TS MMA chains (A = TMEM, B = the dO / Q tile MN-major from the
//   very tile S^T / dP^T read K-major): no transposed copy of anything.
// Two (item, kb) groups are resident in shared memory (K_kb, V_kb, Q, dO: 96 KB at 256 queries); the accumulators of a group
// (dV | dK, 128 columns) are double-buffered so that the epilogue group stores one while the next accumulates.
// TMEM: buffers 2 x 128 | accumulators 2 x 128 = 512 columns.  Warp roles as in attention_ws_kernel.
#include <stdlib.h>

#include "dropout.cuh"
#include "tc_common.cuh"

namespace msq {

template <int KP> struct BwdKvCfg {
  static constexpr int NKB = KP / 128;                   // key blocks per item
  static constexpr int TILE = 128 * 64 * 2;              // 128 rows x 64 dims of bf16 (128B swizzle)
  static constexpr int QD = KP * 64 * 2;                 // all query rows of an item: Q, dO
  static constexpr int STAGE_BYTES = 2 * TILE + 2 * QD;  // K_kb | V_kb | Q | dO
  static constexpr int STAGES = 2;
  static constexpr int OPER = STAGES * STAGE_BYTES;      // 192 KB (KP = 256) / 128 KB (KP = 128)
  static constexpr int STG_OFF = OPER;                   // epilogue staging: one 32-row x 128-byte box per epilogue warp
  static constexpr int VEC_OFF = STG_OFF + 4 * 4096;     // per elementwise group (four of them): lse2[KP] | D[KP]
  static constexpr int BAR_OFF = VEC_OFF + 4 * 2 * KP * 4;
  static constexpr int SMEM = BAR_OFF + 256 + 1024;
  // warps: 0 TMA, 1 MMA, 2-17 four elementwise groups (two per TMEM buffer, a 32-query chunk each: the kernel is bound by the
  // exponentials / dropout decisions / dS products, not by the tensor pipe), 18-21 epilogue
  static constexpr int THREADS = 22 * 32;
};

template <int KP, bool DROP>
__global__ void __launch_bounds__(BwdKvCfg<KP>::THREADS, 1)
attention_bwd_dkv_tc_kernel(const __grid_constant__ CUtensorMap map_kv, const __grid_constant__ CUtensorMap map_q,
                            const __grid_constant__ CUtensorMap map_do, const __grid_constant__ CUtensorMap map_out, int L, int heads,
                            float scale, const float* __restrict__ mask_add, int mask_ld, int mask_len,
                            const float* __restrict__ lse_in, const float* __restrict__ dsum_in, int n_items, Drop drop) {
  using Cfg = BwdKvCfg<KP>;
  constexpr int NKB = Cfg::NKB, TILE = Cfg::TILE, QD = Cfg::QD, STAGE = Cfg::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  float* vecs = reinterpret_cast<float*>(gen + Cfg::VEC_OFF);            // [2 groups][lse2[KP] | D[KP]]
  const uint32_t bars = base + Cfg::BAR_OFF;
  // 8-byte barriers: full[2] | empty[2] | s_full[2] | p_ready[2] | acc_full[2] | acc_empty[2] | tmem slot
  const uint32_t full = bars, empty = bars + 16, s_full = bars + 32, p_ready = bars + 48, acc_full = bars + 64, acc_empty = bars + 80,
                 tslot = bars + 96;
  volatile uint32_t* tslot_gen = reinterpret_cast<volatile uint32_t*>(gen + Cfg::BAR_OFF + 96);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int NQB = (L + 63) >> 6;                                          // 64-query blocks per item
  const int n_groups = n_items * NKB;                                     // group = (item, key block)
  const int my_groups = (n_groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int NB = my_groups * NQB;
  const float LOG2E = 1.4426950408889634f;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kv) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_do) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_out) : "memory");
    for (int i = 0; i < 2; ++i) {
      mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 1); mbar_init(s_full + 8 * i, 1); mbar_init(p_ready + 8 * i, 8);
      mbar_init(acc_full + 8 * i, 1); mbar_init(acc_empty + 8 * i, 4);
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
      int g = blockIdx.x;
      for (int gl = 0; gl < my_groups; ++gl, g += gridDim.x) {
        const int item = g / NKB, kb = g % NKB, s = gl & 1, u = gl >> 1;
        const int row0 = (item / heads) * L, hh = item % heads;
        const uint32_t st0 = base + s * STAGE;
        mbar_wait(empty + 8 * s, (u & 1) ^ 1);
        mbar_expect_tx(full + 8 * s, (uint32_t)STAGE);
        tma_load_2d(st0, &map_kv, heads * 64 + hh * 64, row0 + kb * 128, full + 8 * s);                // K_kb
        tma_load_2d(st0 + TILE, &map_kv, 2 * heads * 64 + hh * 64, row0 + kb * 128, full + 8 * s);     // V_kb
        tma_load_2d(st0 + 2 * TILE, &map_q, hh * 64, row0, full + 8 * s);                              // Q (KP rows)
        tma_load_2d(st0 + 2 * TILE + QD, &map_do, hh * 64, row0, full + 8 * s);                        // dO (KP rows)
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (uniform control flow, one elected lane issues) =====================
    // D=F32, A=B=BF16, M=128, N=64.  S^T / dP^T: A and B K-major.  dV / dK: A from TMEM, B MN-major (bit 16).
    const uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | (8u << 24);
    const uint32_t idesc_o = idesc_s | (1u << 16);
    const uint64_t dk0 = umma_desc_sw128(base), dv0 = umma_desc_sw128_mn(base);
    auto issue_s = [&](int n) {
      const int gl = n / NQB, qb = n - gl * NQB, s = gl & 1, u = gl >> 1;
      if (qb == 0) { mbar_wait(full + 8 * s, u & 1); tc_fence_after(); }
      const uint32_t d = tmem + (n & 1) * 128;
      const uint64_t kt = dk0 + (uint64_t)((s * STAGE) >> 4), vt = kt + (TILE >> 4);
      const uint64_t qt = dk0 + (uint64_t)((s * STAGE + 2 * TILE + qb * 8192) >> 4), gt = qt + (QD >> 4);   // 64 query rows = 8192 B
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_acc<false>(d, kt + 2 * k, qt + 2 * k, idesc_s, k != 0);        // S^T  = K Q^T
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_acc<false>(d + 64, vt + 2 * k, gt + 2 * k, idesc_s, k != 0);   // dP^T = V dO^T
        umma_commit(s_full + 8 * (n & 1));
      }
      __syncwarp();
    };
    auto issue_pv = [&](int n) {
      const int gl = n / NQB, qb = n - gl * NQB, s = gl & 1;
      if (qb == 0) mbar_wait(acc_empty + 8 * (gl & 1), ((gl >> 1) & 1) ^ 1);   // the epilogue has drained these accumulators
      mbar_wait(p_ready + 8 * (n & 1), (n >> 1) & 1);
      tc_fence_after();
      const uint32_t p = tmem + (n & 1) * 128, acc = tmem + 256 + (gl & 1) * 128;
      const uint64_t qm = dv0 + (uint64_t)((s * STAGE + 2 * TILE + qb * 8192) >> 4), gm = qm + (QD >> 4);
      // 32-query chunk c: P^T plane in columns [32 c, 32 c + 16), dS^T plane in [32 c + 16, 32 c + 32); 16 queries = 8 columns
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)   // dV += P^T dO
          umma_bf16_acc<true>(acc, (uint64_t)(p + 32 * (kk >> 1) + 8 * (kk & 1)), gm + 128 * kk, idesc_o, (qb | kk) != 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)   // dK += dS^T Q
          umma_bf16_acc<true>(acc + 64, (uint64_t)(p + 32 * (kk >> 1) + 16 + 8 * (kk & 1)), qm + 128 * kk, idesc_o, (qb | kk) != 0);
        if (qb == NQB - 1) { umma_commit(acc_full + 8 * (gl & 1)); umma_commit(empty + 8 * s); }
      }
      __syncwarp();
    };
    issue_s(0);
    if (NB > 1) issue_s(1);
    for (int n = 0; n < NB; ++n) {
      issue_pv(n);
      if (n + 2 < NB) issue_s(n + 2);   // after PV(n) in issue order: buffer n & 1 is free by then
    }
  } else if (warp < 18) {
    // ===================== elementwise groups: one thread per key row and 32-query chunk =====================
    const int grp = (warp - 2) >> 2, w = grp >> 1, c = grp & 1;   // TMEM buffer (blocks with n & 1 == w), query chunk of the block
    const int quarter = warp & 3, row = quarter * 32 + lane, gtid = ((warp - 2) & 3) * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    float* lse2 = vecs + grp * 2 * KP;
    float* dsm = lse2 + KP;
    const float scale_l2e = scale * LOG2E;
    int have = -1;
    float mr2 = 0.f;
    int item = 0, key = 0;
    for (int n = w; n < NB; n += 2) {
      const int gl = n / NQB, qb = n - gl * NQB;
      if (gl != have) {
        const int g = (int)blockIdx.x + gl * (int)gridDim.x;
        item = g / NKB;
        key = (g % NKB) * 128 + row;
        const int r = item / heads;
        named_bar_sync(1 + grp, 128);   // the previous group's readers are done
        for (int q = gtid; q < KP; q += 128) {
          lse2[q] = q < L ? lse_in[(int64_t)item * L + q] * LOG2E : INFINITY;   // queries beyond the sequence: p = 0
          dsm[q] = q < L ? dsum_in[(int64_t)item * L + q] : 0.f;
        }
        mr2 = key < L ? ((mask_add != nullptr && key < mask_len) ? mask_add[(int64_t)r * mask_ld + key] * LOG2E : 0.f) : -INFINITY;
        named_bar_sync(1 + grp, 128);
        have = gl;
      }
      const uint32_t sb = lane_addr + (n & 1) * 128;
      const uint64_t ebase = (uint64_t)item * L * L + (uint64_t)key;   // dropout element index = ebase + query * L
      mbar_wait(s_full + 8 * (n & 1), (n >> 1) & 1);
      tc_fence_after();
      // two passes of 16 queries; dS^T of the first pass is held in registers until the second pass has read its S^T columns
      // (the dS^T plane of the chunk lies over them)
      uint32_t dd0[8];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t st[16], dp[16], pp[8], dd[8];
        tmem_ld16_nowait(sb + 32 * c + 16 * h, st);
        tmem_ld16_nowait(sb + 64 + 32 * c + 16 * h, dp);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int q0 = qb * 64 + 32 * c + 16 * h;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float p0 = ex2_approx(fmaf(__uint_as_float(st[2 * i]), scale_l2e, mr2) - lse2[q0 + 2 * i]);
          float p1 = ex2_approx(fmaf(__uint_as_float(st[2 * i + 1]), scale_l2e, mr2) - lse2[q0 + 2 * i + 1]);
          float g0 = __uint_as_float(dp[2 * i]), g1 = __uint_as_float(dp[2 * i + 1]);
          if (DROP) {   // element index of the forward's mask: ((item L + query) L + key)
            const float m0 = drop_mul(drop, ebase + (uint32_t)((q0 + 2 * i) * L));
            const float m1 = drop_mul(drop, ebase + (uint32_t)((q0 + 2 * i + 1) * L));
            g0 *= m0; g1 *= m1;
            const float d0 = p0 * (g0 - dsm[q0 + 2 * i]), d1 = p1 * (g1 - dsm[q0 + 2 * i + 1]);
            p0 *= m0; p1 *= m1;                            // dV = (P . mask)^T dO
            const __nv_bfloat162 a = __floats2bfloat162_rn(p0, p1), b = __floats2bfloat162_rn(d0, d1);
            pp[i] = *reinterpret_cast<const uint32_t*>(&a);
            dd[i] = *reinterpret_cast<const uint32_t*>(&b);
          } else {
            const float d0 = p0 * (g0 - dsm[q0 + 2 * i]), d1 = p1 * (g1 - dsm[q0 + 2 * i + 1]);
            const __nv_bfloat162 a = __floats2bfloat162_rn(p0, p1), b = __floats2bfloat162_rn(d0, d1);
            pp[i] = *reinterpret_cast<const uint32_t*>(&a);
            dd[i] = *reinterpret_cast<const uint32_t*>(&b);
          }
        }
        tmem_st8(sb + 32 * c + 8 * h, pp);
        if (h == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) dd0[i] = dd[i];
        } else {
          tmem_st8(sb + 32 * c + 16, dd0);
          tmem_st8(sb + 32 * c + 24, dd);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready + 8 * (n & 1));
    }
  } else {
    // ===================== epilogue group (warps 18-21): dK (scaled) and dV rows of a key block =====================
    const int quarter = warp & 3;
    const uint32_t lane_addr = tmem + ((uint32_t)(quarter * 32) << 16);
    const uint32_t stg = base + Cfg::STG_OFF + (warp - 18) * 4096;
    for (int gl = 0; gl < my_groups; ++gl) {
      const int g = (int)blockIdx.x + gl * (int)gridDim.x, item = g / NKB, kb = g % NKB, r = item / heads, h = item % heads;
      mbar_wait(acc_full + 8 * (gl & 1), (gl >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = lane_addr + 256 + (gl & 1) * 128;
      const int k0 = kb * 128 + quarter * 32;   // first key row of this warp
#pragma unroll 1
      for (int part = 0; part < 2; ++part) {    // 0: dV (columns [0, 64)), 1: dK (columns [64, 128), times the score scale)
        uint32_t a[32], b[32];
        tmem_ld32_nowait(acc + part * 64, a);
        tmem_ld32(acc + part * 64 + 32, b);
        if (part == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty + 8 * (gl & 1));   // both halves are in registers: the accumulators may be reused
        }
        const float f = part ? scale : 1.f;
        if (lane == 0) tma_store_wait_read<0>();   // the previous box has been read out of the staging buffer
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t* src = i < 4 ? a + 8 * i : b + 8 * (i - 4);
          const __nv_bfloat162 x0 = __floats2bfloat162_rn(__uint_as_float(src[0]) * f, __uint_as_float(src[1]) * f);
          const __nv_bfloat162 x1 = __floats2bfloat162_rn(__uint_as_float(src[2]) * f, __uint_as_float(src[3]) * f);
          const __nv_bfloat162 x2 = __floats2bfloat162_rn(__uint_as_float(src[4]) * f, __uint_as_float(src[5]) * f);
          const __nv_bfloat162 x3 = __floats2bfloat162_rn(__uint_as_float(src[6]) * f, __uint_as_float(src[7]) * f);
          st_shared_v4(stg + lane * 128 + ((i ^ (lane & 7)) << 4), *reinterpret_cast<const uint32_t*>(&x0), *reinterpret_cast<const uint32_t*>(&x1),
                       *reinterpret_cast<const uint32_t*>(&x2), *reinterpret_cast<const uint32_t*>(&x3));
        }
        fence_proxy_async_smem();
        __syncwarp();
        // dqkv rows = [dq | dk | dv]: dK at column heads*64, dV at 2*heads*64; the 3-D map clips the box at the end of the item
        if (lane == 0 && k0 < L) { tma_store_3d(&map_out, stg, (part ? 1 : 2) * heads * 64 + h * 64, k0, r); tma_store_commit(); }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

template <int KP>
static int launch_bwd_dkv_tc(const bf16* qkv, const bf16* dctx, int64_t R, int L, int heads, float scale, const float* mask_add, int mask_ld,
                             int mask_len, bf16* dqkv, const float* lse, const float* dsum, cudaStream_t st, const Drop& drop) {
  using Cfg = BwdKvCfg<KP>;
  CUtensorMap mkv, mq, mdo, mout;
  const int ld = 3 * heads * 64, ldc = heads * 64;
  MSQ_TRY(make_map_bf16(&mkv, qkv, R * L, ld, ld, 64, 128));
  MSQ_TRY(make_map_bf16(&mq, qkv, R * L, ld, ld, 64, KP));
  MSQ_TRY(make_map_bf16(&mdo, dctx, R * L, ldc, ldc, 64, KP));
  MSQ_TRY(make_map_3d_bf16(&mout, dqkv, R, L, ld, ld, 64, 32));
  const bool dropping = drop.thresh != 0;
  if (dropping) MSQ_SMEM_ATTR(Cfg::SMEM, (attention_bwd_dkv_tc_kernel<KP, true>));
  else MSQ_SMEM_ATTR(Cfg::SMEM, (attention_bwd_dkv_tc_kernel<KP, false>));
  int dev = 0, sms = 0;
  MSQ_CUDA(cudaGetDevice(&dev));
  MSQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int64_t n_items = R * heads;
  MSQ_REQUIRE(n_items * Cfg::NKB < ((int64_t)1 << 28), "attention backward: too many (row, head) items");
  const dim3 grid((unsigned)min((int64_t)sms, n_items * Cfg::NKB));
  if (dropping)
    MSQ_CUDA(launch_k(attention_bwd_dkv_tc_kernel<KP, true>, grid, dim3(Cfg::THREADS), Cfg::SMEM, st, mkv, mq, mdo, mout, L, heads, scale, mask_add,
                      mask_ld, mask_len, lse, dsum, (int)n_items, drop));
  else
    MSQ_CUDA(launch_k(attention_bwd_dkv_tc_kernel<KP, false>, grid, dim3(Cfg::THREADS), Cfg::SMEM, st, mkv, mq, mdo, mout, L, heads, scale, mask_add,
                      mask_ld, mask_len, lse, dsum, (int)n_items, drop));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

bool attention_bwd_dkv_tc_supported(int L) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("MSQ_ATTN_BWD_TC"); off = (e && e[0] == '0') ? 1 : 0; }
  return !off && tc_supported_impl() && L >= 1 && L <= 256;
}

// dK / dV of dqkv [R*L, 3*heads*64] from qkv, dctx and the per-query lse / D vectors written by the dQ kernel
int attention_bwd_dkv_tc(const bf16* qkv, const bf16* dctx, int64_t R, int L, int heads, float scale, const float* mask_add, int mask_ld,
                         int mask_len, bf16* dqkv, const float* lse, const float* dsum, cudaStream_t st, const Drop& drop) {
  if (L <= 128) return launch_bwd_dkv_tc<128>(qkv, dctx, R, L, heads, scale, mask_add, mask_ld, mask_len, dqkv, lse, dsum, st, drop);
  return launch_bwd_dkv_tc<256>(qkv, dctx, R, L, heads, scale, mask_add, mask_ld, mask_len, dqkv, lse, dsum, st, drop);
}

}  // namespace msq
