// GEMM with a fused residual + LayerNorm epilogue for the N = hidden-size projections (attention output,
// FFN down, ViT out_proj / c_proj):   v = A W^T + bias + resid ;  y = LN(v) * gamma + beta
//
// A LayerNorm row needs all N columns, but one CTA's TMEM holds two 128 x 256 fp32 accumulators.  So a
// CLUSTER of N/256 CTAs (3 for H = 768) owns a 128-row block, one 256-column slice each:
//   pass 1  TMEM -> registers: v = acc + bias + resid (residual rows prefetched before the accumulator is
//           ready), per-row partial sum / sum-of-squares, v written BACK to TMEM (scratch);
//   share   every thread pushes its row partial into the statistics table of ALL CTAs of the cluster
//           (st.shared::cluster) and the warps arrive on every CTA's mbarrier (release / acquire.cluster);
//   pass 2  TMEM -> registers again: normalise, write the fp32 stream and its bf16 copy through swizzled
//           shared-memory boxes + TMA bulk stores.
// This removes the separate LayerNorm kernel AND the fp32 pre-LN round trip through HBM (per BERT
// sub-layer 2.2 GB -> 1.3 GB of traffic).  MMA side identical to gemm_tc.cu's single-CTA path (cta_group::1,
// 128x256x64 tiles, TMA 128B swizzle, 3-stage ring, double-buffered accumulators so the two-pass epilogue of
// tile i overlaps the MMAs of tile i+1).
//   RAW32 = false (BERT post-LN, lxrt/modeling.py:428-439, 482-493):  C <- y (fp32), C2 <- y (bf16)
//   RAW32 = true  (CLIP pre-LN,  clip/model.py:222-226)            :  C <- v (fp32 residual stream, may alias
//                                                                     resid), C2 <- y (bf16, input of the next GEMM)
#include <stdlib.h>

#include "tc_common.cuh"

namespace msq {

namespace {
constexpr int LBM = 128, LBN = 256, LBK = 64, LSTAGES = 3;
constexpr int LA_BYTES = LBM * LBK * 2, LB_BYTES = LBN * LBK * 2, LSTAGE_BYTES = LA_BYTES + LB_BYTES;
constexpr int LEPI_WARPS = 8, LTHREADS = (2 + LEPI_WARPS) * 32;
constexpr int LMAXC = 3;                                      // cluster size = N / 256 <= 3  (H <= 768; smem bound)
constexpr int LSTAGING = LEPI_WARPS * 2 * 4096;               // per warp: one fp32 box + one bf16 box
constexpr int LSTATS = 2 * LMAXC * 2 * 128 * 2 * 4;           // [parity][cta][half][row][sum,sumsq]
constexpr int LVEC = 2 * 128 * 4 * 3;                         // bias | gamma | beta slices per column half
constexpr int LSMEM = LSTAGES * LSTAGE_BYTES + LSTAGING + LSTATS + LVEC + 256 + 1024;
static_assert(LSMEM <= 232448, "gemm_ln shared memory exceeds the 227 KB per-CTA limit");

struct LnEpi {
  const float *bias, *resid, *gamma, *beta;
  int64_t M;
  int N, ldr;
  float eps;
};

__device__ __forceinline__ uint32_t lcluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t lmapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void lcluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (t0 == 0) t0 = clock64();
    if (clock64() - t0 > TC_WAIT_CYCLES) {
      printf("gemm_ln: statistics barrier timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
      "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
      "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
      "r"(v[31])
      : "memory");
}
}  // namespace

template <bool RAW32>
__global__ void __launch_bounds__(LTHREADS, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_c2, LnEpi ep, int num_m,
               int num_k, int csize) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t staging0 = base + LSTAGES * LSTAGE_BYTES;
  const uint32_t stats0 = staging0 + LSTAGING;
  float* stats_gen = reinterpret_cast<float*>(gen + LSTAGES * LSTAGE_BYTES + LSTAGING);
  float* vec_gen = reinterpret_cast<float*>(gen + LSTAGES * LSTAGE_BYTES + LSTAGING + LSTATS);
  const uint32_t bars = stats0 + LSTATS + LVEC;
  // full[3] | empty[3] | tmem_full[2] | tmem_empty[2] | stats[2] | tmem slot
  const uint32_t full0 = bars, empty0 = bars + 8 * LSTAGES, tfull0 = bars + 16 * LSTAGES, tempty0 = tfull0 + 16;
  const uint32_t sbar0 = tempty0 + 16, tslot = sbar0 + 16;
  volatile uint32_t* tslot_gen =
      reinterpret_cast<volatile uint32_t*>(gen + LSTAGES * LSTAGE_BYTES + LSTAGING + LSTATS + LVEC + 16 * LSTAGES + 48);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = lcluster_ctarank();            // this CTA's 256-column slice
  const int cluster_id = (int)blockIdx.x / csize, num_clusters = (int)gridDim.x / csize;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_c) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_c2) : "memory");
    for (int s = 0; s < LSTAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull0 + 8 * a, 1);
      mbar_init(tempty0 + 8 * a, LEPI_WARPS);
      mbar_init(sbar0 + 8 * a, LEPI_WARPS * csize);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tslot), "n"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  lcluster_sync();  // every CTA's barriers exist before any remote arrival
  tc_fence_after();
  const uint32_t tmem_base = *tslot_gen;
  pdl_sync();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int mb = cluster_id; mb < num_m; mb += num_clusters) {
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(empty0 + 8 * stage, phase ^ 1);
          const uint32_t sa = base + stage * LSTAGE_BYTES, sb = sa + LA_BYTES;
          mbar_expect_tx(full0 + 8 * stage, LSTAGE_BYTES);
          tma_load_2d(sa, &tma_a, kb * LBK, mb * LBM, full0 + 8 * stage);
          tma_load_2d(sb, &tma_b, kb * LBK, (int)rank * LBN, full0 + 8 * stage);
          if (++stage == LSTAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(LBN >> 3) << 17) | ((uint32_t)(LBM >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int mb = cluster_id; mb < num_m; mb += num_clusters) {
        mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * LBN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full0 + 8 * stage, phase);
          tc_fence_after();
          const uint32_t sa = base + stage * LSTAGE_BYTES, sb = sa + LA_BYTES;
#pragma unroll
          for (int k = 0; k < LBK / 16; ++k)
            umma_bf16(tmem_d, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), idesc, (kb | k) != 0);
          umma_commit(empty0 + 8 * stage);
          if (++stage == LSTAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull0 + 8 * acc);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: residual + LayerNorm across the cluster =====================
    const int q = warp & 3, half = (warp - 2) >> 2, w8 = warp - 2;
    const int trow = q * 32 + lane;
    float* bias_s = vec_gen + half * 384;  // the four warps of a column half write identical values
    float* gamma_s = bias_s + 128;
    float* beta_s = bias_s + 256;
    const uint32_t box32 = staging0 + w8 * 8192, box16 = box32 + 4096;
    const int colw = (int)rank * LBN + half * 128;  // first global column of this warp
    {
      const int c = colw + lane * 4;
      *reinterpret_cast<float4*>(bias_s + lane * 4) = ep.bias ? __ldg(reinterpret_cast<const float4*>(ep.bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(gamma_s + lane * 4) = __ldg(reinterpret_cast<const float4*>(ep.gamma + c));
      *reinterpret_cast<float4*>(beta_s + lane * 4) = __ldg(reinterpret_cast<const float4*>(ep.beta + c));
      __syncwarp();
    }
    int acc = 0, it = 0;
    uint32_t acc_phase = 0;
    const float invN = 1.0f / (float)ep.N;
    for (int mb = cluster_id; mb < num_m; mb += num_clusters, ++it) {
      const int64_t row = (int64_t)mb * LBM + trow;
      const bool row_ok = row < ep.M;
      const int par = it & 1;
      float4 res[8], res_next[8];
      auto fetch_res = [&](int ch, float4* dst) {
        const float* rp = ep.resid + row * ep.ldr + colw + ch * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = row_ok ? __ldg(reinterpret_cast<const float4*>(rp + 4 * i)) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      fetch_res(0, res_next);
      mbar_wait(tfull0 + 8 * acc, acc_phase);
      __syncwarp();
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * LBN + half * 128;

      // ---- pass 1: v = acc + bias + resid -> row partials; v goes back to TMEM
      float s1 = 0.f, s2 = 0.f;
      uint32_t raw[32];
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        tmem_ld32(taddr + ch * 32, raw);
#pragma unroll
        for (int i = 0; i < 8; ++i) res[i] = res_next[i];
        if (ch + 1 < 4) fetch_res(ch + 1, res_next);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b = *reinterpret_cast<const float4*>(bias_s + ch * 32 + j * 4);
          const float v0 = __uint_as_float(raw[4 * j + 0]) + b.x + res[j].x, v1 = __uint_as_float(raw[4 * j + 1]) + b.y + res[j].y;
          const float v2 = __uint_as_float(raw[4 * j + 2]) + b.z + res[j].z, v3 = __uint_as_float(raw[4 * j + 3]) + b.w + res[j].w;
          s1 += (v0 + v1) + (v2 + v3);
          s2 += (v0 * v0 + v1 * v1) + (v2 * v2 + v3 * v3);
          raw[4 * j + 0] = __float_as_uint(v0); raw[4 * j + 1] = __float_as_uint(v1);
          raw[4 * j + 2] = __float_as_uint(v2); raw[4 * j + 3] = __float_as_uint(v3);
        }
        tmem_st32(taddr + ch * 32, raw);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");

      // ---- share the partials with every CTA of the cluster (slot [par][rank][half][row])
      const uint32_t slot = stats0 + (uint32_t)((((par * LMAXC + (int)rank) * 2 + half) * 128 + trow) * 8);
      for (int c = 0; c < csize; ++c) st_cluster_v2(lmapa(slot, c), s1, s2);
      __syncwarp();
      if (lane == 0)
        for (int c = 0; c < csize; ++c) mbar_arrive_cluster_release(lmapa(sbar0 + 8 * par, c));
      mbar_wait_cluster(sbar0 + 8 * par, (uint32_t)((it >> 1) & 1));
      __syncwarp();
      float t1 = 0.f, t2 = 0.f;
      for (int c = 0; c < csize; ++c) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const float2 pr = *reinterpret_cast<const float2*>(stats_gen + (((par * LMAXC + c) * 2 + hh) * 128 + trow) * 2);
          t1 += pr.x;
          t2 += pr.y;
        }
      }
      const float mean = t1 * invN;
      const float rstd = rsqrtf(fmaxf(t2 * invN - mean * mean, 0.f) + ep.eps);

      // ---- pass 2: normalise and store (fp32 box per 32 columns, bf16 box per 64 columns)
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        tmem_ld32(taddr + ch * 32, raw);
        if (lane == 0) tma_store_wait_read<0>();  // both boxes of this warp are free again
        __syncwarp();
        const uint32_t r32 = box32 + lane * 128, r16 = box16 + lane * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v[8], y[8];
          const float4 g0 = *reinterpret_cast<const float4*>(gamma_s + ch * 32 + j * 8), g1 = *reinterpret_cast<const float4*>(gamma_s + ch * 32 + j * 8 + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(beta_s + ch * 32 + j * 8), b1 = *reinterpret_cast<const float4*>(beta_s + ch * 32 + j * 8 + 4);
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[i] = __uint_as_float(raw[j * 8 + i]);
            y[i] = fmaf((v[i] - mean) * rstd, gg[i], bb[i]);
          }
          const float* o = RAW32 ? v : y;
          st_shared_v4(r32 + (((2 * j) ^ (lane & 7)) << 4), __float_as_uint(o[0]), __float_as_uint(o[1]), __float_as_uint(o[2]), __float_as_uint(o[3]));
          st_shared_v4(r32 + (((2 * j + 1) ^ (lane & 7)) << 4), __float_as_uint(o[4]), __float_as_uint(o[5]), __float_as_uint(o[6]), __float_as_uint(o[7]));
          __nv_bfloat162 p0 = __floats2bfloat162_rn(y[0], y[1]), p1 = __floats2bfloat162_rn(y[2], y[3]);
          __nv_bfloat162 p2 = __floats2bfloat162_rn(y[4], y[5]), p3 = __floats2bfloat162_rn(y[6], y[7]);
          st_shared_v4(r16 + ((((ch & 1) * 4 + j) ^ (lane & 7)) << 4), *reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                       *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tma_c, box32, colw + ch * 32, (int)(row - lane));
          if (ch & 1) tma_store_2d(&tma_c2, box16, colw + (ch - 1) * 32, (int)(row - lane));
          tma_store_commit();
        }
        // the bf16 box collects two chunks: it must not be overwritten before its store was issued; the wait at the
        // top of the next iteration only frees it after that store has read it
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * acc);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  lcluster_sync();  // peers may still be writing this CTA's statistics table / barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

bool gemm_ln_supported(int N, int K) {
  return tc_supported_impl() && N % LBN == 0 && N / LBN >= 1 && N / LBN <= LMAXC && K % LBK == 0;
}
// Policy: OFF by default.  Measured on B200 (profiles/r1d_gemm_ln.txt): out-proj + LN 627 us fused vs 257 + 150 us
// unfused, FFN-down + LN 1135 us vs 550 + 150 us.  A 128 x 256 tile per CTA without operand sharing triples the
// L2 -> shared-memory traffic per output and the per-tile statistics rendezvous serialises the epilogue; the CTA-pair
// GEMM followed by the stand-alone (97%-of-HBM-roofline) LayerNorm kernel is faster.  MSQ_GEMM_LN=1 switches it on.
bool gemm_ln_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("MSQ_GEMM_LN"); on = (e && e[0] == '1') ? 1 : 0; }
  return on != 0;
}

// C (fp32) and C2 (bf16) have leading dimension N; resid may alias C when raw32.
int gemm_ln(const bf16* A, int lda, const bf16* W, int ldw, const float* bias, const float* resid, int ldr, const float* gamma,
            const float* beta, float eps, float* C, bf16* C2, int64_t M, int N, int K, bool raw32, cudaStream_t st) {
  MSQ_REQUIRE(gemm_ln_supported(N, K), "gemm_ln: N=%d K=%d unsupported", N, K);
  MSQ_REQUIRE(resid != nullptr && gamma != nullptr && beta != nullptr, "gemm_ln: residual / LayerNorm parameters required");
  if (M == 0) return MSQ_OK;
  CUtensorMap ma, mb, mc, mc2;
  MSQ_TRY(make_map_bf16(&ma, A, M, K, lda, LBK, LBM));
  MSQ_TRY(make_map_bf16(&mb, W, N, K, ldw, LBK, LBN));
  MSQ_TRY(make_map_2d(&mc, C, M, N, N, 32, 32, true));
  MSQ_TRY(make_map_2d(&mc2, C2, M, N, N, 64, 32, false));
  static int sms = 0;
  if (!sms) {
    int dev = 0;
    MSQ_CUDA(cudaGetDevice(&dev));
    MSQ_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  if (raw32) MSQ_SMEM_ATTR(LSMEM, gemm_ln_kernel<true>);
  else MSQ_SMEM_ATTR(LSMEM, gemm_ln_kernel<false>);
  LnEpi ep;
  ep.bias = bias; ep.resid = resid; ep.gamma = gamma; ep.beta = beta; ep.M = M; ep.N = N; ep.ldr = ldr; ep.eps = eps;
  const int csize = N / LBN, num_m = ceil_div(M, LBM), num_k = K / LBK;
  const int clusters = (int)min((int64_t)(sms / csize), (int64_t)num_m);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * csize);
  cfg.blockDim = dim3(LTHREADS);
  cfg.dynamicSmemBytes = LSMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled();
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  profile_mark(st, false, 0.0);
  if (raw32) MSQ_CUDA(cudaLaunchKernelEx(&cfg, gemm_ln_kernel<true>, ma, mb, mc, mc2, ep, num_m, num_k, csize));
  else MSQ_CUDA(cudaLaunchKernelEx(&cfg, gemm_ln_kernel<false>, ma, mb, mc, mc2, ep, num_m, num_k, csize));
  MSQ_LAUNCH_CHECK();
  profile_mark(st, true, 2.0 * (double)M * (double)N * (double)K);
  return MSQ_OK;
}

}  // namespace msq
