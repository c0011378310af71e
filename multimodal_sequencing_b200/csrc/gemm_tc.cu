// bf16 tensor-core GEMM for sm_100a: C = act(A W^T + bias) + resid with fp32 accumulation in TMEM.
//
//   * operands A [M,K] and W [N,K] (both K-major bf16) are streamed global -> shared by TMA
//     (cp.async.bulk.tensor, 128B swizzle) through a 4-stage mbarrier ring;
//   * one elected thread issues tcgen05.mma (cta_group::1, 128 x 256 x 16 per instruction) reading
//     the swizzled tiles through shared-memory matrix descriptors, accumulating into TMEM;
//   * the 512 TMEM columns hold TWO 128x256 fp32 accumulators so the epilogue of tile i overlaps the
//     MMAs of tile i+1; 8 epilogue warps read TMEM with tcgen05.ld, apply bias / activation /
//     residual in registers and store bf16 or fp32 rows;
//   * persistent: one CTA per SM walks tiles n-fastest, so the CTAs running together cover all N tiles of
//     a few M blocks: every A tile is fetched from HBM once and re-read from L2, the (small) weight
//     matrix stays L2-resident.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM alloc/dealloc), 2..9 = epilogue.
//
// Replaces the cuBLAS calls behind every nn.Linear of the encoders: CLIP ViT blocks
// (models/CLIP/clip/model.py:204-226, conv1 263), joint BERT layers
// (models/CLIP/src/lxrt/modeling.py:373-507), visn_fc (585-602) and HierarchicalAttention.sentence_tran
// (models/berson/modeling_bert.py:697).
#include "tc_common.cuh"

namespace msq {

constexpr int TC_BM = 128, TC_BN = 256, TC_BK = 64, TC_STAGES = 4;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2, TC_B_BYTES = TC_BN * TC_BK * 2, TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;
constexpr int TC_EPI_WARPS = 8, TC_THREADS = (2 + TC_EPI_WARPS) * 32;
constexpr int TC_SMEM = TC_STAGES * TC_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;

struct TcEpi {
  const float* bias;
  const float* resid;
  void* C;
  int64_t M;
  int N, ldc, ldr, act;
};

template <typename TO> __device__ __forceinline__ void epi_store8(TO* p, const float* v);
template <> __device__ __forceinline__ void epi_store8<float>(float* p, const float* v) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> __device__ __forceinline__ void epi_store8<bf16>(bf16* p, const float* v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  uint4 u;
  u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
  u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
  *reinterpret_cast<uint4*>(p) = u;
}

template <typename TO>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, TcEpi ep, int num_m,
               int num_n, int num_k) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = base + TC_STAGES * TC_STAGE_BYTES;
  // barrier layout (8 B each): full[4] | empty[4] | tmem_full[2] | tmem_empty[2] | tmem base slot
  const uint32_t full0 = bars, empty0 = bars + 8 * TC_STAGES, tfull0 = bars + 16 * TC_STAGES, tempty0 = tfull0 + 16;
  const uint32_t tslot = tempty0 + 16;
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tslot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + TC_STAGES * TC_STAGE_BYTES + 16 * TC_STAGES + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_tiles = num_m * num_n;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull0 + 8 * a, 1); mbar_init(tempty0 + 8 * a, TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tslot_gen;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int n_blk = tile % num_n, m_blk = tile / num_n;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          const uint32_t sa = base + stage * TC_STAGE_BYTES, sb = sa + TC_A_BYTES;
          mbar_expect_tx(full0 + 8 * stage, TC_STAGE_BYTES);
          tma_load_2d(sa, &tma_a, kb * TC_BK, m_blk * TC_BM, full0 + 8 * stage);
          tma_load_2d(sb, &tma_b, kb * TC_BK, n_blk * TC_BN, full0 + 8 * stage);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor: D=F32 [4,6), A=BF16 [7,10), B=BF16 [10,13), K-major A/B, N>>3 [17,23), M>>4 [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * TC_BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = base + stage * TC_STAGE_BYTES, sb = sa + TC_A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint64_t da = umma_desc_sw128(sa + k * 32), db = umma_desc_sw128(sb + k * 32);
            umma_bf16(tmem_d, da, db, idesc, (kb | k) != 0);
          }
          umma_commit(empty0 + 8 * stage);  // frees the smem slot when these MMAs have read it
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull0 + 8 * acc);  // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (8 warps) =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may touch
    const int half = (warp - 2) >> 2;       // column half: 0 -> [0,128), 1 -> [128,256)
    int acc = 0;
    uint32_t acc_phase = 0;
    TO* C = reinterpret_cast<TO*>(ep.C);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int n_blk = tile % num_n, m_blk = tile / num_n;
      mbar_wait(tfull0 + 8 * acc, acc_phase);
      __syncwarp();
      tc_fence_after();
      const int64_t row = (int64_t)m_blk * TC_BM + q * 32 + lane;
      const bool row_ok = row < ep.M;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        const int col0 = n_blk * TC_BN + half * 128 + ch * 32;
        uint32_t raw[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * TC_BN + half * 128 + ch * 32, raw);
        if (row_ok && col0 < ep.N) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = col0 + j * 8;
            if (col < ep.N) {  // N % 8 == 0
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(raw[j * 8 + i]);
              if (ep.bias) {
                const float4 b0 = *reinterpret_cast<const float4*>(ep.bias + col), b1 = *reinterpret_cast<const float4*>(ep.bias + col + 4);
                v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
              }
              if (ep.act != ACT_NONE) {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = apply_act(v[i], ep.act);
              }
              if (ep.resid) {
                const float* rp = ep.resid + row * ep.ldr + col;
                const float4 r0 = *reinterpret_cast<const float4*>(rp), r1 = *reinterpret_cast<const float4*>(rp + 4);
                v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
              }
              epi_store8<TO>(C + row * ep.ldc + col, v);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------------

int gemm_tc_selftest_supported() { return tc_supported_impl(); }

template <typename TO>
int gemm_tc(const GemmArgs& g, cudaStream_t st) {
  MSQ_REQUIRE(gemm_tc_selftest_supported(), "gemm_tc: tcgen05 path needs an sm_100 device and cuTensorMapEncodeTiled");
  MSQ_REQUIRE(g.K % TC_BK == 0 && g.N % 8 == 0 && g.lda % 8 == 0 && g.ldw % 8 == 0 && g.ldc % 8 == 0,
              "gemm_tc: K=%d N=%d lda=%d ldw=%d ldc=%d not supported", g.K, g.N, g.lda, g.ldw, g.ldc);
  MSQ_REQUIRE(((uintptr_t)g.A & 15) == 0 && ((uintptr_t)g.W & 15) == 0 && ((uintptr_t)g.C & 15) == 0, "gemm_tc: unaligned pointer");
  MSQ_REQUIRE(g.C2 == nullptr, "gemm_tc: second output unsupported");
  if (g.M == 0) return MSQ_OK;
  CUtensorMap ma, mb;
  MSQ_TRY(make_map_bf16(&ma, g.A, g.M, g.K, g.lda, TC_BK, TC_BM));
  MSQ_TRY(make_map_bf16(&mb, g.W, g.N, g.K, g.ldw, TC_BK, TC_BN));
  static int sms = 0;
  static bool configured[2] = {false, false};
  if (!sms) {
    int dev = 0;
    MSQ_CUDA(cudaGetDevice(&dev));
    MSQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int which = sizeof(TO) == 4 ? 0 : 1;
  if (!configured[which]) {
    MSQ_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    configured[which] = true;
  }
  TcEpi ep;
  ep.bias = g.bias; ep.resid = g.resid; ep.C = g.C; ep.M = g.M; ep.N = g.N; ep.ldc = g.ldc; ep.ldr = g.ldr; ep.act = g.act;
  const int num_m = ceil_div(g.M, TC_BM), num_n = ceil_div(g.N, TC_BN), num_k = g.K / TC_BK;
  const int64_t tiles = (int64_t)num_m * num_n;
  const int grid = (int)min((int64_t)sms, tiles);
  profile_mark(st, false, 0.0);
  gemm_tc_kernel<TO><<<grid, TC_THREADS, TC_SMEM, st>>>(ma, mb, ep, num_m, num_n, num_k);
  MSQ_LAUNCH_CHECK();
  profile_mark(st, true, 2.0 * (double)g.M * (double)g.N * (double)g.K);
  return MSQ_OK;
}
template int gemm_tc<float>(const GemmArgs&, cudaStream_t);
template int gemm_tc<bf16>(const GemmArgs&, cudaStream_t);

}  // namespace msq
