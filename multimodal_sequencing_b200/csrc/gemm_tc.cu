// bf16 tensor-core GEMM for sm_100a: C = act(A W^T + bias) + resid with fp32 accumulation in TMEM.
//
//   * operands A [M,K] and W [N,K] (both K-major bf16) are streamed global -> shared by TMA
//     (cp.async.bulk.tensor, 128B swizzle) through a 4-stage mbarrier ring;
//   * one elected thread issues tcgen05.mma (cta_group::1, 128 x 256 x 16 per instruction) reading
//     the swizzled tiles through shared-memory matrix descriptors, accumulating into TMEM;
//   * the 512 TMEM columns hold TWO 128x256 fp32 accumulators so the epilogue of tile i overlaps the
//     MMAs of tile i+1; 8 epilogue warps read TMEM with tcgen05.ld, apply bias / activation /
//     residual in registers and store bf16 or fp32 rows;
//   * PAIR variant: a 2-CTA cluster (two SMs of one TPC) owns a 256 x 256 tile; each CTA stages its own
//     128 rows of A and 128 rows of W and ONE thread of the leader CTA issues tcgen05.mma.cta_group::2
//     (M=256) which reads both CTAs' shared memory: L2->SM operand traffic per FLOP drops by 1/3 versus
//     the 128 x 256 single-CTA tile (the single-CTA kernel is L2-bandwidth bound at ~45% of tensor peak);
//   * persistent: one CTA (pair) per SM (pair) walks tiles n-fastest, so the CTAs running together cover all N tiles of
//     a few M blocks: every A tile is fetched from HBM once and re-read from L2, the (small) weight
//     matrix stays L2-resident.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc/dealloc), 2..9 = epilogue.
//
// Replaces the cuBLAS calls behind every nn.Linear of the encoders: CLIP ViT blocks
// (models/CLIP/clip/model.py:204-226, conv1 263), joint BERT layers
// (models/CLIP/src/lxrt/modeling.py:373-507), visn_fc (585-602) and HierarchicalAttention.sentence_tran
// (models/berson/modeling_bert.py:697).
#include <stdlib.h>

#include "tc_common.cuh"

namespace msq {

constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 64;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;
constexpr int TC_EPI_WARPS = 8, TC_THREADS = (2 + TC_EPI_WARPS) * 32;

// MODE: 0 plain, 1 EPI_LNFOLD, 2 EPI_RESLN (bf16 copy), 3 EPI_RESLN with a SPLIT-bf16 copy (bf16x3 mode: hi and lo boxes)
// 4 EPI_DUALACT (bf16 pre-activation in C + bf16 activation in C2: the warp's two 4 KB boxes hold one output each)
template <bool PAIR, int MODE = 0> struct TcCfg {
  static constexpr int STAGES = 4;
  // MODE 3 (split copy: two more boxes per warp) keeps ONE fp32 box per warp instead of two, so that the operand ring stays at
  // four stages: the residual chunk is loaded when the box has been stored (no chunk-ahead prefetch); at three MMAs per
  // product the mainloop of a tile is long enough to hide that latency, a three-stage ring was not (+26 % per launch).
  static constexpr int F32_BOXES = (PAIR && MODE == 3) ? 1 : 2;
  // PAIR: the epilogue stages 32-row x 128-byte output boxes in shared memory (2 per warp) and writes
  // them with TMA bulk tensor stores (full-line, asynchronous) instead of per-thread 16-byte stores.
  static constexpr bool TMA_STORE = PAIR;
  static constexpr int STAGING_F32 = TMA_STORE ? TC_EPI_WARPS * F32_BOXES * 4096 : 0;
  // EPI_RESLN: one more box per warp for the bf16 copy, 32 rows x 32 columns (64-byte rows, 64B swizzle), stored every chunk
  static constexpr int STAGING_BYTES = STAGING_F32 + ((TMA_STORE && (MODE == 2 || MODE == 3)) ? (MODE == 3 ? 2 : 1) * TC_EPI_WARPS * 2048 : 0);
  static constexpr int VECS = MODE == 0 ? 1 : (MODE == 1 ? 2 : 3);   // per-column vectors held per warp: bias | svec/gamma | beta
  static constexpr int B_ROWS = PAIR ? 128 : 256;            // rows of W staged per CTA and stage
  static constexpr int B_BYTES = B_ROWS * TC_BK * 2;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int TILE_M = PAIR ? 256 : 128;            // rows of C per scheduling unit (CTA or CTA pair)
  static constexpr int SMEM = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align*/ + 256 /*barriers*/ + VECS * TC_EPI_WARPS * 128 * 4;
};

struct TcEpi {
  const float* bias;
  const float* resid;
  void* C;
  int64_t M;
  int N, ldc, ldr, act;
  // deferred LayerNorm (GemmArgs)
  const float* svec;
  const float* beta;
  const float* stats_in;
  float* stats_out;
  bf16* C2;
  int sp_in;
  float inv_dim, eps;
  int tn;   // GemmArgs::tn: MN-major operands (single-CTA path)
  int split_k;  // GemmArgs::split: logical K (A and W rows are planes of K columns: [hi | lo] or [hi | mid | lo]); 0 = plain bf16
  int split_passes;  // 3 (two planes: bf16x3) or 6 (three planes: bf16x6, fp32-grade products)
  // bf16x3 on CTA pairs: a K slab is staged ONCE as two stages ([a_hi | w_hi], [a_lo | w_lo]) and the three MMA passes
  // read them crosswise, instead of three stages that fetch a_hi and w_hi twice: L2 -> shared-memory traffic -1/3
  int share;
  Drop drop;   // GemmArgs::drop
};

// 8 consecutive output columns of one row: accumulator -> value to store (see GemmArgs for the modes)
template <int ACT, int MODE>
__device__ __forceinline__ void epi_compute8(float* v, const uint32_t* raw, const float* bias8, const float* s8, const float* beta8,
                                             float ra, float rc, bool has_ln, bool use_res, const float4& r0, const float4& r1,
                                             bool exact_act, const Drop& drop, uint64_t e0) {
  const float4 b0 = *reinterpret_cast<const float4*>(bias8), b1 = *reinterpret_cast<const float4*>(bias8 + 4);
  const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  if (MODE == 1) {
    const float4 s0 = *reinterpret_cast<const float4*>(s8), s1 = *reinterpret_cast<const float4*>(s8 + 4);
    const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = fmaf(ra, __uint_as_float(raw[i]), fmaf(rc, sv[i], b[i]));
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(raw[i]) + b[i];
  }
  if (ACT != ACT_NONE) {
    if (exact_act) {   // bf16x3 mode: erff / tanhf / expf as in the fp32 parity mode
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = apply_act(v[i], ACT);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = apply_act_fast(v[i], ACT);
    }
  }
  if (MODE == 0 && drop.thresh) {   // dropout of the dense output (fine-tuning), before the residual
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= drop_mul(drop, e0 + i);
  }
  if (use_res) {
    float r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    if (MODE >= 2 && has_ln) {
      const float4 g0 = *reinterpret_cast<const float4*>(s8), g1 = *reinterpret_cast<const float4*>(s8 + 4);
      const float4 e0 = *reinterpret_cast<const float4*>(beta8), e1 = *reinterpret_cast<const float4*>(beta8 + 4);
      const float gm[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bt[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) r[i] = fmaf(fmaf(ra, r[i], rc), gm[i], bt[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] += r[i];
  }
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are credited to an mbarrier given by its shared::cluster address
// (the leader CTA's barrier when issued by the peer CTA of a pair)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit to the barrier at the same shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

template <typename TO> __device__ __forceinline__ void epi_store8(TO* p, const float* v);
template <> __device__ __forceinline__ void epi_store8<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
// split output: hi at p, lo at p + plane
__device__ __forceinline__ void epi_pack8_split(const float* v, uint4& hi, uint4& lo) {
  uint2 h0, l0, h1, l1;
  split4(make_float4(v[0], v[1], v[2], v[3]), h0, l0);
  split4(make_float4(v[4], v[5], v[6], v[7]), h1, l1);
  hi = make_uint4(h0.x, h0.y, h1.x, h1.y);
  lo = make_uint4(l0.x, l0.y, l1.x, l1.y);
}
template <> __device__ __forceinline__ void epi_store8<bf16>(bf16* p, const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(p) = u;
}

template <typename TO, bool PAIR, int ACT, int MODE>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_c2,
               const __grid_constant__ CUtensorMap tma_r, TcEpi ep, int num_m, int num_n, int num_k) {
  using Cfg = TcCfg<PAIR, MODE>;
  constexpr bool F32 = same_type<TO, float>::value, SPL = is_split<TO>::value;
  static_assert(MODE < 2 || MODE == 4 || F32, "EPI_RESLN writes the fp32 stream (+ its bf16 / split-bf16 copy)");
  static_assert(!SPL || MODE <= 1, "split-bf16 output: plain or LayerNorm-folded epilogue");
  static_assert(MODE != 4 || (!F32 && !SPL && ACT == ACT_NONE), "EPI_DUALACT: bf16 outputs, the activation is a run-time argument");
  constexpr bool RESLN = MODE == 2 || MODE == 3, C2SPL = MODE == 3, DUAL = MODE == 4;
  constexpr int STAGES = Cfg::STAGES, STAGE_BYTES = Cfg::STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t staging0 = base + STAGES * STAGE_BYTES;  // [EPI_WARPS][2][4096], 1024-aligned
  const uint32_t bars = staging0 + Cfg::STAGING_BYTES;
  // barrier layout (8 B each): full[STAGES] | empty[STAGES] | tmem_full[2] | tmem_empty[2] | tmem base slot
  const uint32_t full0 = bars, empty0 = bars + 8 * STAGES, tfull0 = bars + 16 * STAGES, tempty0 = tfull0 + 16;
  const uint32_t tslot = tempty0 + 16;
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tslot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + Cfg::STAGING_BYTES + 16 * STAGES + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = num_m * num_n;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;            // 0 = leader CTA of the pair
  const int unit = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int num_units = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
    if (Cfg::TMA_STORE) asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_c) : "memory");
    if (Cfg::TMA_STORE && (RESLN || DUAL)) asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_c2) : "memory");
    if (Cfg::TMA_STORE && RESLN) asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_r) : "memory");
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, TC_EPI_WARPS * (PAIR ? 2 : 1)); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrival / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tslot_gen;
  pdl_sync();  // everything above (barriers, TMEM, cluster rendezvous) overlapped the previous kernel's tail

  if (warp == 0) {
    // ===================== TMA producer (every CTA stages its own rows) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = unit; tile < num_tiles; tile += num_units) {
        const int n_blk = tile % num_n, m_blk = tile / num_n;
        const int a_row = m_blk * Cfg::TILE_M + (int)rank * TC_BM, b_row = n_blk * TC_BN + (int)rank * Cfg::B_ROWS;
        for (int kq = 0; kq < num_k; ++kq) {
          // bf16x3: the K loop runs three times over every 64-column slab: a_hi*w_lo, a_lo*w_hi, a_hi*w_hi (the hi and
          // lo planes of a row are K columns apart); the MMA warp just sees a contraction of length 3K
          int kb = kq, ka = 0, kw = 0;
          if (PAIR && ep.share) {
            kb = kq >> 1;
            ka = kw = ((kq & 1) ^ 1) * ep.split_k;   // even stage: LO planes (released first, after the second pass), odd stage: hi planes
          } else if (ep.split_k) {
            // (plane of A, plane of W) per pass, smallest products first.  two planes: (0,1) (1,0) (0,0);
            // three planes: (0,2) (1,1) (2,0) (0,1) (1,0) (0,0) -- every product a_i w_j with i + j <= 2
            kb = kq / ep.split_passes;
            const int t = kq - ep.split_passes * kb;
            const uint32_t pa = ep.split_passes == 3 ? 0x010u : 0x010210u, pw = ep.split_passes == 3 ? 0x001u : 0x001012u;
            ka = (int)((pa >> (4 * t)) & 0xFu) * ep.split_k;
            kw = (int)((pw >> (4 * t)) & 0xFu) * ep.split_k;
          }
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + TC_A_BYTES;
          if (PAIR && ep.tn) {
            // MN-major operands on a CTA pair: every CTA stages the 128 M columns / 128 N columns of its half of the tile as
            // boxes of 64 contraction rows x 64 columns
            const uint32_t lead_full = mapa_rank(full0 + 8 * stage, 0);
            if (rank == 0) mbar_expect_tx(full0 + 8 * stage, 2 * STAGE_BYTES);
#pragma unroll
            for (int i = 0; i < TC_BM / 64; ++i) tma_load_2d_pair(sa + i * 8192, &tma_a, a_row + i * 64, kb * TC_BK, lead_full);
#pragma unroll
            for (int i = 0; i < Cfg::B_ROWS / 64; ++i) tma_load_2d_pair(sb + i * 8192, &tma_b, b_row + i * 64, kb * TC_BK, lead_full);
          } else if (PAIR) {
            const uint32_t lead_full = mapa_rank(full0 + 8 * stage, 0);
            if (rank == 0) mbar_expect_tx(full0 + 8 * stage, 2 * STAGE_BYTES);
            tma_load_2d_pair(sa, &tma_a, ka + kb * TC_BK, a_row, lead_full);
            tma_load_2d_pair(sb, &tma_b, kw + kb * TC_BK, b_row, lead_full);
          } else if (ep.tn) {
            // MN-major operands: boxes of 64 contraction rows x 64 columns, one per 64-wide slice of the tile
            mbar_expect_tx(full0 + 8 * stage, STAGE_BYTES);
#pragma unroll
            for (int i = 0; i < TC_BM / 64; ++i) tma_load_2d(sa + i * 8192, &tma_a, a_row + i * 64, kb * TC_BK, full0 + 8 * stage);
#pragma unroll
            for (int i = 0; i < Cfg::B_ROWS / 64; ++i) tma_load_2d(sb + i * 8192, &tma_b, b_row + i * 64, kb * TC_BK, full0 + 8 * stage);
          } else {
            mbar_expect_tx(full0 + 8 * stage, STAGE_BYTES);
            tma_load_2d(sa, &tma_a, ka + kb * TC_BK, a_row, full0 + 8 * stage);
            tma_load_2d(sb, &tma_b, kw + kb * TC_BK, b_row, full0 + 8 * stage);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      // instruction descriptor: D=F32 [4,6), A=BF16 [7,10), B=BF16 [10,13), K-major A/B, N>>3 [17,23), M>>4 [24,29)
      // TN mode: A and B MN-major (bits 15, 16)
      const bool tn = ep.tn != 0;
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(Cfg::TILE_M >> 4) << 24) |
                             (tn ? ((1u << 15) | (1u << 16)) : 0u);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = unit; tile < num_tiles; tile += num_units) {
        mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * TC_BN;
        if (PAIR && ep.share) {
          for (int kb = 0; kb < num_k; kb += 2) {
            mbar_wait(full0 + 8 * stage, phase);
            mbar_wait(full0 + 8 * (stage + 1), phase);
            tc_fence_after();
            const uint32_t al = base + stage * STAGE_BYTES, wl = al + TC_A_BYTES, ah = al + STAGE_BYTES, wh = ah + TC_A_BYTES;
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)   // smallest products first
              umma_bf16_pair(tmem_d, umma_desc_sw128(ah + k * 32), umma_desc_sw128(wl + k * 32), idesc, (kb | k) != 0);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) umma_bf16_pair(tmem_d, umma_desc_sw128(al + k * 32), umma_desc_sw128(wh + k * 32), idesc, 1);
            if (ep.share == 1) umma_commit_pair(empty0 + 8 * stage);   // the lo planes are done: their stage refills a pass earlier
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k) umma_bf16_pair(tmem_d, umma_desc_sw128(ah + k * 32), umma_desc_sw128(wh + k * 32), idesc, 1);
            if (ep.share != 1) umma_commit_pair(empty0 + 8 * stage);
            umma_commit_pair(empty0 + 8 * (stage + 1));
            stage += 2;
            if (stage == STAGES) { stage = 0; phase ^= 1; }
          }
        } else
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = base + stage * STAGE_BYTES, sb = sa + TC_A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            // K-major: 16 elements = 32 B further along the swizzled row; MN-major: 16 contraction rows = 2048 B further down
            const uint64_t da = tn ? umma_desc_sw128_mn(sa + k * 2048) : umma_desc_sw128(sa + k * 32);
            const uint64_t db = tn ? umma_desc_sw128_mn(sb + k * 2048) : umma_desc_sw128(sb + k * 32);
            if (PAIR) umma_bf16_pair(tmem_d, da, db, idesc, (kb | k) != 0);
            else umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
          if (PAIR) umma_commit_pair(empty0 + 8 * stage); else umma_commit(empty0 + 8 * stage);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (PAIR) umma_commit_pair(tfull0 + 8 * acc); else umma_commit(tfull0 + 8 * acc);  // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (8 warps per CTA) =====================
    // Each warp owns 32 TMEM lanes (rows) x 128 columns = 4 chunks of 32 columns.  Software pipeline:
    // the residual rows of the first chunk and the bias slice are fetched BEFORE the accumulator is
    // ready; inside a tile, the TMEM load and residual fetch of chunk c+1 are in flight while chunk c
    // is converted and stored, so DRAM / TMEM latency is off the critical path.
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int half = (warp - 2) >> 2;       // column half: 0 -> [0,128), 1 -> [128,256)
    float* bias_s = reinterpret_cast<float*>(smem_gen + STAGES * STAGE_BYTES + Cfg::STAGING_BYTES + 256) + (warp - 2) * 128;
    float* svec_s = bias_s + TC_EPI_WARPS * 128;   // MODE 1: s_n; MODE 2: gamma of the LayerNorm pending on the residual
    float* beta_s = svec_s + TC_EPI_WARPS * 128;   // MODE 2: its beta
    constexpr bool RES_DB = Cfg::F32_BOXES == 2;           // residual chunk prefetched one chunk ahead into the other box
    const uint32_t stg = staging0 + (warp - 2) * (Cfg::F32_BOXES * 4096);  // this warp's staging box(es)
    const uint32_t stg2 = staging0 + Cfg::STAGING_F32 + (warp - 2) * 2048;  // MODE 2 / 3: bf16 (hi) box (32 rows x 32 columns)
    const uint32_t stg3 = stg2 + TC_EPI_WARPS * 2048;                        // MODE 3: lo box
    int stg_use = 0;                                       // boxes handed to the TMA so far (parity selects the buffer)
    // EPI_RESLN on CTA pairs: the residual chunk (32 rows x 32 fp32) is TMA-LOADED into the very staging box the output
    // chunk is stored from: full 128-byte lines and no registers held across the latency, instead of 16 bytes per row
    // per load instruction.  One chunk ahead; rk counts this warp's chunks (box rk&1, mbarrier parity (rk>>1)&1).
    constexpr bool RES_TMA = RESLN && Cfg::TMA_STORE;
    const uint32_t rbar = bars + 128 + (warp - 2) * 16;   // two 8-byte mbarriers per epilogue warp
    uint32_t rk = 0;
    auto issue_res = [&](int t, int ch, uint32_t k) {      // lane 0
      const int r0 = (t / num_n) * Cfg::TILE_M + (int)rank * TC_BM + q * 32;
      const int c0 = (t % num_n) * TC_BN + half * 128 + ch * 32;
      mbar_expect_tx(rbar + 8 * (k & 1), 4096);
      tma_load_2d(stg + (RES_DB ? (k & 1) * 4096 : 0), &tma_r, c0, r0, rbar + 8 * (k & 1));
    };
    if (RES_TMA) {
      if (lane == 0) {
        mbar_init(rbar, 1); mbar_init(rbar + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (RES_DB && unit < num_tiles) issue_res(unit, 0, 0);
      }
      __syncwarp();
    }
    int acc = 0;
    uint32_t acc_phase = 0;
    TO* C = reinterpret_cast<TO*>(ep.C);
    const bool use_res = ep.resid != nullptr;
    const bool exact_act = SPL || ep.split_k != 0;
    // residual rows are pulled into L2 one tile ahead (no registers held): the epilogue of a K=768 residual GEMM is
    // HBM-latency bound otherwise, with only one 32-column chunk of loads in flight per thread
    auto prefetch_res = [&](int t) {
      if (!use_res || t >= num_tiles) return;
      const int64_t r2 = (int64_t)(t / num_n) * Cfg::TILE_M + (int)rank * TC_BM + q * 32 + lane;
      const int c2 = (t % num_n) * TC_BN + half * 128;
      if (r2 >= ep.M) return;
      const float* p = ep.resid + r2 * ep.ldr + c2;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (c2 + 32 * i < ep.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + 32 * i));
    };
    prefetch_res(unit);
    for (int tile = unit; tile < num_tiles; tile += num_units) {
      const int n_blk = tile % num_n, m_blk = tile / num_n;
      const int64_t row = (int64_t)m_blk * Cfg::TILE_M + (int)rank * TC_BM + q * 32 + lane;
      const bool row_ok = row < ep.M;
      const int colw = n_blk * TC_BN + half * 128;   // first column of this warp
      prefetch_res(tile + num_units);
      float4 res[8];
      auto fetch_res = [&](int ch, float4* dst) {
        const int c0 = colw + ch * 32;
        const float* rp = ep.resid + row * ep.ldr + c0;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          dst[i] = (row_ok && c0 + 4 * i < ep.N) ? __ldg(reinterpret_cast<const float4*>(rp + 4 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      if (use_res && !RES_TMA) fetch_res(0, res);
      {
        const int c = colw + lane * 4;
        float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ep.bias && c < ep.N) b = __ldg(reinterpret_cast<const float4*>(ep.bias + c));
        float4 sv = make_float4(0.f, 0.f, 0.f, 0.f), bt = sv;
        if (MODE >= 1 && ep.svec && c < ep.N) sv = __ldg(reinterpret_cast<const float4*>(ep.svec + c));
        if (RESLN && ep.beta && c < ep.N) bt = __ldg(reinterpret_cast<const float4*>(ep.beta + c));
        __syncwarp();
        *reinterpret_cast<float4*>(bias_s + lane * 4) = b;
        if (MODE >= 1) *reinterpret_cast<float4*>(svec_s + lane * 4) = sv;
        if (RESLN) *reinterpret_cast<float4*>(beta_s + lane * 4) = bt;
        __syncwarp();
      }
      // pending LayerNorm of this row: ra = rstd, rc = -rstd * mean, from the partial sums of the producer
      float ra = 1.f, rc = 0.f;
      const bool has_ln = MODE != 0 && ep.stats_in != nullptr;
      if (has_ln && row_ok) {
        float s1 = 0.f, s2 = 0.f;
        const float2* sp = reinterpret_cast<const float2*>(ep.stats_in) + row * ep.sp_in;
        for (int i = 0; i < ep.sp_in; ++i) { const float2 t = __ldg(sp + i); s1 += t.x; s2 += t.y; }
        const float mu = s1 * ep.inv_dim;
        ra = rsqrtf(fmaxf(s2 * ep.inv_dim - mu * mu, 0.f) + ep.eps);
        rc = -ra * mu;
      }
      float st_sum = 0.f, st_sq = 0.f;   // MODE 2: statistics of the row segment this thread writes
      mbar_wait(tfull0 + 8 * acc, acc_phase);
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * TC_BN + half * 128;
      uint32_t raw[32], raw_next[32];
      float4 res_next[8];
      tmem_ld32_nowait(taddr, raw_next);
#pragma unroll
      for (int i = 0; i < 8; ++i) res_next[i] = res[i];
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[i] = raw_next[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) res[i] = res_next[i];
        if (ch + 1 < 4) {
          tmem_ld32_nowait(taddr + (ch + 1) * 32, raw_next);
          if (use_res && !RES_TMA) fetch_res(ch + 1, res_next);
        }
        const int col0 = colw + ch * 32;
        if (Cfg::TMA_STORE) {
          // ---- registers -> swizzled staging box -> TMA store.  A box is 32 rows x 128 B: 32 fp32 columns
          // (one chunk) or 64 bf16 columns (two chunks).  16-byte piece c of row r sits at r*128 + ((c ^ (r&7))<<4).
          const bool new_box = F32 || (ch & 1) == 0;
          if (RES_TMA) {
            if (lane == 0) {
              // every store issued so far has been read out of shared memory: the other fp32 box may receive the next
              // residual chunk (of this tile or of this warp's next tile) and the bf16 box may be refilled
              tma_store_wait_read<0>();
              if (!RES_DB) issue_res(tile, ch, rk);   // single box: it has just been read out by the previous chunk's store
              else if (ch + 1 < 4) issue_res(tile, ch + 1, rk + 1);
              else if (tile + num_units < num_tiles) issue_res(tile + num_units, 0, rk + 1);
            }
            __syncwarp();
            mbar_wait(rbar + 8 * (rk & 1), (rk >> 1) & 1);   // residual chunk rk has landed in box rk&1
          } else if (DUAL) {
            // box 0 = pre-activation, box 1 = activation, 64 bf16 columns (two chunks) each: both are rewritten every second
            // chunk, once the stores issued from them have been read out of shared memory
            if (new_box) {
              if (lane == 0) tma_store_wait_read<0>();
              __syncwarp();
            }
          } else if (new_box) {
            // the fp32 box used two chunks ago has been read out (split output: box 0 = hi, box 1 = lo, both reused)
            if (lane == 0) { if (SPL) tma_store_wait_read<0>(); else tma_store_wait_read<1>(); }
            __syncwarp();
          }
          const uint32_t box = (SPL || !RES_DB || DUAL) ? stg : stg + ((RES_TMA ? rk : (uint32_t)stg_use) & 1) * 4096;
          const uint32_t rowp = box + lane * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v[8];
            if (RES_TMA) {
              const float4 r0 = ld_shared_f4(rowp + (((2 * j) ^ (lane & 7)) << 4)), r1 = ld_shared_f4(rowp + (((2 * j + 1) ^ (lane & 7)) << 4));
              epi_compute8<ACT, MODE>(v, raw + j * 8, bias_s + ch * 32 + j * 8, svec_s + ch * 32 + j * 8, beta_s + ch * 32 + j * 8, ra, rc,
                                      has_ln, true, r0, r1, exact_act, ep.drop, (uint64_t)row * ep.N + (col0 + j * 8));
            } else {
              epi_compute8<ACT, MODE>(v, raw + j * 8, bias_s + ch * 32 + j * 8, svec_s + ch * 32 + j * 8, beta_s + ch * 32 + j * 8, ra, rc,
                                      has_ln, use_res, res[2 * j], res[2 * j + 1], exact_act, ep.drop, (uint64_t)row * ep.N + (col0 + j * 8));
            }
            if (RESLN) {
#pragma unroll
              for (int i = 0; i < 8; ++i) { st_sum += v[i]; st_sq = fmaf(v[i], v[i], st_sq); }
              // 64-byte rows, 64B swizzle: 16-byte piece j of row r sits at r*64 + ((j ^ ((r >> 1) & 3)) << 4)
              const uint32_t o2 = lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4);
              if (C2SPL) {
                uint4 hi, lo;
                epi_pack8_split(v, hi, lo);
                st_shared_v4(stg2 + o2, hi.x, hi.y, hi.z, hi.w);
                st_shared_v4(stg3 + o2, lo.x, lo.y, lo.z, lo.w);
              } else {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
                st_shared_v4(stg2 + o2, *reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                             *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
              }
            }
            if (DUAL) {   // second output: act(v) as bf16 into the warp's second box (same 128-byte-row layout as the first)
              float w[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) w[i] = apply_act_fast(v[i], ep.act);
              __nv_bfloat162 p0 = __floats2bfloat162_rn(w[0], w[1]), p1 = __floats2bfloat162_rn(w[2], w[3]);
              __nv_bfloat162 p2 = __floats2bfloat162_rn(w[4], w[5]), p3 = __floats2bfloat162_rn(w[6], w[7]);
              st_shared_v4(rowp + 4096 + ((((ch & 1) * 4 + j) ^ (lane & 7)) << 4), *reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                           *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
            }
            if (F32) {
              st_shared_v4(rowp + (((2 * j) ^ (lane & 7)) << 4), __float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
              st_shared_v4(rowp + (((2 * j + 1) ^ (lane & 7)) << 4), __float_as_uint(v[4]), __float_as_uint(v[5]), __float_as_uint(v[6]), __float_as_uint(v[7]));
            } else if (SPL) {
              uint4 hi, lo;
              epi_pack8_split(v, hi, lo);
              const uint32_t o = rowp + ((((ch & 1) * 4 + j) ^ (lane & 7)) << 4);
              st_shared_v4(o, hi.x, hi.y, hi.z, hi.w);
              st_shared_v4(o + 4096, lo.x, lo.y, lo.z, lo.w);
            } else {
              __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0], v[1]), p1 = __floats2bfloat162_rn(v[2], v[3]);
              __nv_bfloat162 p2 = __floats2bfloat162_rn(v[4], v[5]), p3 = __floats2bfloat162_rn(v[6], v[7]);
              st_shared_v4(rowp + ((((ch & 1) * 4 + j) ^ (lane & 7)) << 4), *reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                           *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
            }
          }
          if (F32 || (ch & 1) == 1) {
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              const int box_col = F32 ? col0 : col0 - 32;
              if (RESLN) {
                if (col0 < ep.N) {
                  tma_store_2d(&tma_c2, stg2, col0, (int)(row - lane));
                  if (C2SPL) tma_store_2d(&tma_c2, stg3, ep.ldc + col0, (int)(row - lane));   // lo plane of the split copy
                }
                tma_store_commit();
              }
              if (box_col < ep.N) {
                tma_store_2d(&tma_c, box, box_col, (int)(row - lane));
                if (SPL) tma_store_2d(&tma_c, box + 4096, ep.ldc + box_col, (int)(row - lane));   // lo plane (N % 64 == 0)
                if (DUAL) tma_store_2d(&tma_c2, box + 4096, box_col, (int)(row - lane));          // activation box
              }
              tma_store_commit();
            }
            ++stg_use;
            ++rk;
          }
        } else if (row_ok && col0 < ep.N) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = col0 + j * 8;
            if (col < ep.N) {  // N % 8 == 0
              float v[8];
              epi_compute8<ACT, MODE>(v, raw + j * 8, bias_s + ch * 32 + j * 8, svec_s + ch * 32 + j * 8, beta_s + ch * 32 + j * 8, ra, rc,
                                      has_ln, use_res, res[2 * j], res[2 * j + 1], exact_act, ep.drop, (uint64_t)row * ep.N + col);
              if (RESLN) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { st_sum += v[i]; st_sq = fmaf(v[i], v[i], st_sq); }
                if (C2SPL) {   // row of the copy = [hi(ldc) | lo(ldc)]
                  bf16* cb = ep.C2 + row * (2 * (int64_t)ep.ldc) + col;
                  uint4 hi, lo;
                  epi_pack8_split(v, hi, lo);
                  *reinterpret_cast<uint4*>(cb) = hi;
                  *reinterpret_cast<uint4*>(cb + ep.ldc) = lo;
                } else {
                  epi_store8<bf16>(ep.C2 + row * ep.ldc + col, v);
                }
              }
              if (DUAL) {
                float w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = apply_act_fast(v[i], ep.act);
                epi_store8<bf16>(ep.C2 + row * ep.ldc + col, w);
              }
              if constexpr (SPL) {   // row of C = [hi(ldc) | lo(ldc)]
                bf16* cb = reinterpret_cast<bf16*>(ep.C) + row * (2 * (int64_t)ep.ldc) + col;
                uint4 hi, lo;
                epi_pack8_split(v, hi, lo);
                *reinterpret_cast<uint4*>(cb) = hi;
                *reinterpret_cast<uint4*>(cb + ep.ldc) = lo;
              } else {
                epi_store8<TO>(C + row * ep.ldc + col, v);
              }
            }
          }
        }
      }
      if (RESLN && ep.stats_out && row_ok)
        reinterpret_cast<float2*>(ep.stats_out)[row * (2 * num_n) + n_blk * 2 + half] = make_float2(st_sum, st_sq);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        // the leader's MMA thread may reuse this accumulator only when BOTH CTAs have drained it
        if (PAIR) mbar_arrive_cluster(mapa_rank(tempty0 + 8 * acc, 0)); else mbar_arrive(tempty0 + 8 * acc);
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  if (Cfg::TMA_STORE && warp >= 2 && lane == 0) tma_store_wait_read<0>();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // no CTA exits (or frees TMEM) while its peer may still signal / read it
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------------

int gemm_tc_selftest_supported() { return tc_supported_impl(); }

template <typename TO, bool PAIR, int ACT, int MODE>
static int launch_gemm_tc_act(const GemmArgs& g, int sms, cudaStream_t st) {
  using Cfg = TcCfg<PAIR, MODE>;
  CUtensorMap ma, mb;
  if (g.tn) {   // [K rows, M or N columns], boxes of 64 x 64
    MSQ_TRY(make_map_bf16(&ma, g.A, g.K, (int)g.M, g.lda, 64, 64));
    MSQ_TRY(make_map_bf16(&mb, g.W, g.K, g.N, g.ldw, 64, 64));
  } else if (g.split) {   // rows are planes of K columns ([hi | lo] or [hi | mid | lo]); lda / ldw count logical elements
    const int P = g.split + 1;
    MSQ_TRY(make_map_bf16(&ma, g.A, g.M, P * g.K, P * g.lda, TC_BK, TC_BM));
    MSQ_TRY(make_map_bf16(&mb, g.W, g.N, P * g.K, P * g.ldw, TC_BK, Cfg::B_ROWS));
  } else {
    MSQ_TRY(make_map_bf16(&ma, g.A, g.M, g.K, g.lda, TC_BK, TC_BM));
    MSQ_TRY(make_map_bf16(&mb, g.W, g.N, g.K, g.ldw, TC_BK, Cfg::B_ROWS));
  }
  CUtensorMap mc = ma;
  CUtensorMap mc2 = ma;
  constexpr bool F32 = same_type<TO, float>::value, SPL = is_split<TO>::value;
  if (Cfg::TMA_STORE && SPL) MSQ_TRY(make_map_2d(&mc, g.C, g.M, g.ldc + g.N, 2 * g.ldc, 64, 32, false));   // lo plane at column ldc
  else if (Cfg::TMA_STORE) MSQ_TRY(make_map_2d(&mc, g.C, g.M, g.N, g.ldc, F32 ? 32 : 64, 32, F32));
  if (Cfg::TMA_STORE && MODE == 2) MSQ_TRY(make_map_2d(&mc2, g.C2bf, g.M, g.N, g.ldc, 32, 32, false, true));
  if (Cfg::TMA_STORE && MODE == 4) MSQ_TRY(make_map_2d(&mc2, g.C2bf, g.M, g.N, g.ldc, 64, 32, false));   // same box shape as C
  if (Cfg::TMA_STORE && MODE == 3) MSQ_TRY(make_map_2d(&mc2, g.C2bf, g.M, g.ldc + g.N, 2 * g.ldc, 32, 32, false, true));   // [hi(ldc) | lo(ldc)]
  CUtensorMap mr = ma;
  if (Cfg::TMA_STORE && (MODE == 2 || MODE == 3)) MSQ_TRY(make_map_2d(&mr, g.resid, g.M, g.N, g.ldr, 32, 32, true));
  MSQ_SMEM_ATTR(Cfg::SMEM, gemm_tc_kernel<TO, PAIR, ACT, MODE>);
  TcEpi ep;
  ep.bias = g.bias; ep.resid = g.resid; ep.C = g.C; ep.M = g.M; ep.N = g.N; ep.ldc = g.ldc; ep.ldr = g.ldr; ep.act = g.act;
  ep.svec = g.svec; ep.beta = g.beta; ep.stats_in = g.stats_in; ep.stats_out = g.stats_out; ep.C2 = (bf16*)g.C2bf; ep.sp_in = g.sp_in;
  ep.inv_dim = g.ln_inv_dim; ep.eps = g.ln_eps; ep.tn = g.tn; ep.split_k = g.split ? g.K : 0;
  ep.drop = g.drop;
  MSQ_REQUIRE(g.drop.thresh == 0 || MODE == 0, "gemm_tc: dropout needs the plain epilogue");
  ep.split_passes = (g.split == 2 && !g.trunc) ? 6 : 3;   // trunc: planes 0 / 1 of a three-plane row only (bf16x3)
  static int share_env = -1;
  if (share_env < 0) { const char* e = getenv("MSQ_X3_SHARE"); share_env = (e && e[0] == '0') ? 0 : 1; }
  static int early_env = -1;
  if (early_env < 0) { const char* e = getenv("MSQ_X3_EARLY"); early_env = (e && e[0] == '0') ? 0 : 1; }
  ep.share = (PAIR && (g.split == 1 || (g.split == 2 && g.trunc)) && !g.tn && share_env) ? (early_env ? 1 : 2) : 0;
  static_assert(Cfg::STAGES % 2 == 0, "the shared bf16x3 slab occupies two consecutive stages");
  const int num_m = ceil_div(g.M, Cfg::TILE_M), num_n = ceil_div(g.N, TC_BN),
            num_k = ceil_div(g.K, TC_BK) * (ep.share ? 2 : (g.split ? ep.split_passes : 1));
  const int64_t tiles = (int64_t)num_m * num_n;
  profile_mark(st, false, 0.0);
  if (PAIR) {
    const int pairs = (int)min((int64_t)(sms / 2), tiles);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled();
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    MSQ_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<TO, PAIR, ACT, MODE>, ma, mb, mc, mc2, mr, ep, num_m, num_n, num_k));
  } else {
    const int grid = (int)min((int64_t)sms, tiles);
    MSQ_CUDA(launch_k(gemm_tc_kernel<TO, PAIR, ACT, MODE>, dim3(grid), dim3(TC_THREADS), Cfg::SMEM, st, ma, mb, mc, mc2, mr, ep, num_m, num_n, num_k));
  }
  MSQ_LAUNCH_CHECK();
  profile_mark(st, true, 2.0 * (double)g.M * (double)g.N * (double)g.K);
  return MSQ_OK;
}

template <typename TO, bool PAIR>
static int launch_gemm_tc(const GemmArgs& g, int sms, cudaStream_t st) {
  if (g.mode == EPI_DUALACT) {
    if constexpr (same_type<TO, bf16>::value) {
      MSQ_REQUIRE(g.C2bf && !g.resid && !g.tn && !g.split && g.drop.thresh == 0 && ((uintptr_t)g.C2bf & 15) == 0,
                  "gemm_tc: EPI_DUALACT needs C2bf and plain bf16 K-major operands, no residual / dropout");
      MSQ_REQUIRE(g.act == ACT_GELU_ERF || g.act == ACT_QUICK_GELU || g.act == ACT_TANH || g.act == ACT_RELU, "gemm_tc: EPI_DUALACT activation %d", g.act);
      return launch_gemm_tc_act<TO, PAIR, ACT_NONE, 4>(g, sms, st);
    } else {
      set_error("gemm_tc: EPI_DUALACT writes bf16 outputs");
      return MSQ_ERR_ARG;
    }
  }
  if (g.mode == EPI_RESLN) {
    if constexpr (same_type<TO, float>::value) {
      MSQ_REQUIRE(g.act == ACT_NONE && g.C2bf && g.stats_out && g.resid, "gemm_tc: EPI_RESLN needs resid, C2bf, stats_out and no activation");
      if (g.split) return launch_gemm_tc_act<TO, PAIR, ACT_NONE, 3>(g, sms, st);   // split operands -> split copy
      return launch_gemm_tc_act<TO, PAIR, ACT_NONE, 2>(g, sms, st);
    } else {
      set_error("gemm_tc: EPI_RESLN writes an fp32 stream");
      return MSQ_ERR_ARG;
    }
  }
  if constexpr (is_split<TO>::value) {
    MSQ_REQUIRE(g.mode == EPI_PLAIN || g.mode == EPI_LNFOLD, "gemm_tc: split-bf16 output supports the plain and LayerNorm-folded epilogues");
    if (g.mode == EPI_LNFOLD) {
      MSQ_REQUIRE(g.svec && g.stats_in && g.sp_in > 0, "gemm_tc: EPI_LNFOLD needs svec and stats_in");
      switch (g.act) {
        case ACT_NONE: return launch_gemm_tc_act<TO, PAIR, ACT_NONE, 1>(g, sms, st);
        case ACT_GELU_ERF: return launch_gemm_tc_act<TO, PAIR, ACT_GELU_ERF, 1>(g, sms, st);
        case ACT_QUICK_GELU: return launch_gemm_tc_act<TO, PAIR, ACT_QUICK_GELU, 1>(g, sms, st);
      }
      set_error("gemm_tc: activation %d not instantiated for EPI_LNFOLD", g.act);
      return MSQ_ERR_ARG;
    }
    switch (g.act) {
      case ACT_NONE: return launch_gemm_tc_act<TO, PAIR, ACT_NONE, 0>(g, sms, st);
      case ACT_GELU_ERF: return launch_gemm_tc_act<TO, PAIR, ACT_GELU_ERF, 0>(g, sms, st);
      case ACT_QUICK_GELU: return launch_gemm_tc_act<TO, PAIR, ACT_QUICK_GELU, 0>(g, sms, st);
      case ACT_RELU: return launch_gemm_tc_act<TO, PAIR, ACT_RELU, 0>(g, sms, st);
    }
    set_error("gemm_tc: activation %d not instantiated for split-bf16 output", g.act);
    return MSQ_ERR_ARG;
  } else {
  if (g.mode == EPI_LNFOLD) {
    MSQ_REQUIRE(g.svec && g.stats_in && g.sp_in > 0, "gemm_tc: EPI_LNFOLD needs svec and stats_in");
    switch (g.act) {
      case ACT_NONE: return launch_gemm_tc_act<TO, PAIR, ACT_NONE, 1>(g, sms, st);
      case ACT_GELU_ERF: return launch_gemm_tc_act<TO, PAIR, ACT_GELU_ERF, 1>(g, sms, st);
      case ACT_QUICK_GELU: return launch_gemm_tc_act<TO, PAIR, ACT_QUICK_GELU, 1>(g, sms, st);
    }
    set_error("gemm_tc: activation %d not instantiated for EPI_LNFOLD", g.act);
    return MSQ_ERR_ARG;
  }
  switch (g.act) {
    case ACT_NONE: return launch_gemm_tc_act<TO, PAIR, ACT_NONE, 0>(g, sms, st);
    case ACT_GELU_ERF: return launch_gemm_tc_act<TO, PAIR, ACT_GELU_ERF, 0>(g, sms, st);
    case ACT_QUICK_GELU: return launch_gemm_tc_act<TO, PAIR, ACT_QUICK_GELU, 0>(g, sms, st);
    case ACT_TANH: return launch_gemm_tc_act<TO, PAIR, ACT_TANH, 0>(g, sms, st);
    case ACT_RELU: return launch_gemm_tc_act<TO, PAIR, ACT_RELU, 0>(g, sms, st);
  }
  set_error("gemm_tc: activation %d not supported on the tensor-core path", g.act);
  return MSQ_ERR_ARG;
  }
}

template <typename TO>
int gemm_tc(const GemmArgs& g, cudaStream_t st) {
  MSQ_REQUIRE(gemm_tc_selftest_supported(), "gemm_tc: tcgen05 path needs an sm_100 device and cuTensorMapEncodeTiled");
  MSQ_REQUIRE((g.tn || g.K % TC_BK == 0) && g.N % 8 == 0 && g.lda % 8 == 0 && g.ldw % 8 == 0 && g.ldc % 8 == 0,
              "gemm_tc: K=%d N=%d lda=%d ldw=%d ldc=%d not supported", g.K, g.N, g.lda, g.ldw, g.ldc);
  MSQ_REQUIRE(!g.tn || (g.mode == EPI_PLAIN && g.M % 8 == 0 && g.K >= 1), "gemm_tc: TN operands need the plain epilogue and M %% 8 == 0");
  MSQ_REQUIRE(((uintptr_t)g.A & 15) == 0 && ((uintptr_t)g.W & 15) == 0 && ((uintptr_t)g.C & 15) == 0, "gemm_tc: unaligned pointer");
  MSQ_REQUIRE(g.C2 == nullptr, "gemm_tc: second output unsupported");
  MSQ_REQUIRE(!g.split || (!g.tn && g.K % TC_BK == 0 && (g.mode == EPI_PLAIN || g.split == 1)), "gemm_tc: split-bf16 operands need K-major layout and K %% 64 == 0 (three planes: plain epilogue)");
  MSQ_REQUIRE(g.split >= 0 && g.split <= 2 && (!g.split || (g.K == g.lda && g.K == g.ldw)), "gemm_tc: split=%d operands need lda == ldw == K (planes are K columns apart)", g.split);
  MSQ_REQUIRE(!is_split<TO>::value || (g.split == 1 && g.N % 64 == 0), "gemm_tc: split-bf16 output needs split operands and N %% 64 == 0");
  if (g.M == 0) return MSQ_OK;
  static int sms = 0, force_single = -1;
  if (!sms) {
    int dev = 0;
    MSQ_CUDA(cudaGetDevice(&dev));
    MSQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (force_single < 0) { const char* e = getenv("MSQ_GEMM_1CTA"); force_single = (e && e[0] == '1') ? 1 : 0; }
  // CTA pairs pay off once there are enough 256 x 256 tiles to occupy most SM pairs
  const int64_t pair_tiles = (int64_t)ceil_div(g.M, 256) * ceil_div(g.N, TC_BN);
  if (!g.tn && !force_single && pair_tiles >= sms / 4) return launch_gemm_tc<TO, true>(g, sms, st);
  // MN-major operands (weight gradients): the output is a few dozen tiles and the split-K slices of one gradient run side by
  // side on auxiliary streams, so the pair form is chosen on shape alone (whole 256 x 256 tiles), not on the tile count
  static int tn_pair = -1;
  if (tn_pair < 0) { const char* e = getenv("MSQ_WGRAD_PAIR"); tn_pair = (e && e[0] == '0') ? 0 : 1; }
  if (g.tn && tn_pair && !force_single && g.M % 256 == 0 && g.N % 256 == 0) return launch_gemm_tc<TO, true>(g, sms, st);
  return launch_gemm_tc<TO, false>(g, sms, st);
}
template int gemm_tc<float>(const GemmArgs&, cudaStream_t);
template int gemm_tc<bf16>(const GemmArgs&, cudaStream_t);
template int gemm_tc<bf16s>(const GemmArgs&, cudaStream_t);

}  // namespace msq
