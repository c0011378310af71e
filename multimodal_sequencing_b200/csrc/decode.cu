// Pointer-network decoder with batched beam search over step permutations — ONE persistent kernel
// for all N-1 decode steps of a group of manuals (state stays in shared memory between steps; no host
// sync, no per-step launches).  Each step fuses: LSTM cell -> query projection -> masked pointer
// scoring over the remaining steps -> log-softmax -> per-manual top-k over (beam x step) -> beam
// expansion, permutation-mask update and parent re-gather of (h, c).
//
// Replaces (telin0411/multimodal_sequencing):
//   BertForOrdering.step            models/berson/modeling_bert.py:1368-1402
//   beam_search_pointer loop body   models/berson/modeling_bert.py:1472-1552
//   Beam.step                       models/berson/generator.py:15-38  (== models/beam.py with `//`)
//
// Base-tensor formulation: the reference re-materialises a per-beam
// [W,N,N,H+2] relation tensor, zeroes it in place and re-gathers it every step.  Here the relation
// tensor R0 of a manual is projected ONCE through the four column blocks of pw_k
// (T4 = [A1|A2|F|G], a GEMM done before this kernel) and a step only gathers rows of T4:
//   keys[k] = A1[last,k] + A2[last2,k] + (sum_{j in Rem, j!=k} F[k,j] + sum_{i in Rem, i!=k} G[i,k]) / N
// Per-beam state is just (picked sequence, h, c, cost).  Likewise W_ih x_t is a row gather of
// XG = sents W_ih^T + b_ih + b_hh.  All arithmetic is fp32 with a fixed summation order.
// Ties in the top-k are broken by the lowest flat index beam*N+step.
#include "kernels.cuh"

namespace msq {

constexpr int DC_MAXN = 16;
constexpr int DC_THREADS = 256;

template <int ROWS>
__global__ void __launch_bounds__(DC_THREADS) beam_search_kernel(DecodeWeights w, DecodeIO io, int G) {
  pdl_sync();
  extern __shared__ __align__(16) float smf[];
  const int H = io.H, N = io.N, W = io.W;
  float* hs = smf;                 // [ROWS][H]  h (input of the step), later q
  float* hn = hs + ROWS * H;       // [ROWS][H]  h'
  float* cs = hn + ROWS * H;       // [ROWS][H]  c (updated in place)
  __shared__ float cost[ROWS], ncost[ROWS];
  __shared__ float e[ROWS][DC_MAXN];
  __shared__ uint8_t seq[ROWS][DC_MAXN], nseq[ROWS][DC_MAXN];
  __shared__ int parent[ROWS];
  __shared__ int nlive[ROWS], nnlive[ROWS];  // per manual slot g (g < G <= ROWS)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t b0 = (int64_t)blockIdx.x * G;
  const int Gc = (int)min((int64_t)G, io.B - b0);  // manuals handled by this CTA
  const int H4 = 4 * H;

  // ---- init: one live beam per manual with h = h0, c = 0, cost 0
  for (int i = tid; i < ROWS * H; i += DC_THREADS) {
    const int r = i / H, d = i % H, g = r / W;
    hs[i] = (g < Gc && r % W == 0) ? io.h0[(b0 + g) * H + d] : 0.f;
    cs[i] = 0.f;
  }
  if (tid < ROWS) { cost[tid] = 0.f; nlive[tid] = 1; }
  __syncthreads();

  for (int t = 0; t < N - 1; ++t) {
    // ---- 1. LSTM cell for every live row: gates = XG[prev] + W_hh h ; thread owns hidden unit u
    for (int u = tid; u < H; u += DC_THREADS) {
      float4 acc[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const int g = r / W, wslot = r % W;
        if (g < Gc && wslot < nlive[g]) {
          const int prev = t == 0 ? N : seq[r][t - 1];
          acc[r] = *reinterpret_cast<const float4*>(io.xg + ((b0 + g) * (N + 1) + prev) * (int64_t)H4 + 4 * u);
        } else {
          acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      const float* wp = w.whh_t + 4 * u;
      // weights stream from L2: keep KU rows in flight beyond the KU being consumed (the loop is latency bound)
      constexpr int KU = 8;
      float4 wn[KU];
#pragma unroll
      for (int i = 0; i < KU; ++i) wn[i] = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)i * H4));
      for (int k = 0; k < H; k += KU) {
        float4 wc[KU];
#pragma unroll
        for (int i = 0; i < KU; ++i) wc[i] = wn[i];
        if (k + KU < H) {
#pragma unroll
          for (int i = 0; i < KU; ++i) wn[i] = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)(k + KU + i) * H4));
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
#pragma unroll
          for (int kk = 0; kk < KU; kk += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(hs + r * H + k + kk);
            const float hk[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc[r].x = fmaf(wc[kk + i].x, hk[i], acc[r].x);
              acc[r].y = fmaf(wc[kk + i].y, hk[i], acc[r].y);
              acc[r].z = fmaf(wc[kk + i].z, hk[i], acc[r].z);
              acc[r].w = fmaf(wc[kk + i].w, hk[i], acc[r].w);
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const float ig = 1.f / (1.f + expf(-acc[r].x)), fg = 1.f / (1.f + expf(-acc[r].y));
        const float gg = tanhf(acc[r].z), og = 1.f / (1.f + expf(-acc[r].w));
        const float c2 = fg * cs[r * H + u] + ig * gg;
        cs[r * H + u] = c2;
        hn[r * H + u] = og * tanhf(c2);
      }
    }
    __syncthreads();

    // ---- 2. q = W_q h' + b_q  (written over hs)
    for (int j4 = tid; j4 < H / 4; j4 += DC_THREADS) {
      float4 acc[ROWS];
      const float4 bq = *reinterpret_cast<const float4*>(w.bq + 4 * j4);
#pragma unroll
      for (int r = 0; r < ROWS; ++r) acc[r] = bq;
      const float* wp = w.wq_t + 4 * j4;
      constexpr int KU = 8;
      float4 wn[KU];
#pragma unroll
      for (int i = 0; i < KU; ++i) wn[i] = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)i * H));
      for (int k = 0; k < H; k += KU) {
        float4 wc[KU];
#pragma unroll
        for (int i = 0; i < KU; ++i) wc[i] = wn[i];
        if (k + KU < H) {
#pragma unroll
          for (int i = 0; i < KU; ++i) wn[i] = __ldg(reinterpret_cast<const float4*>(wp + (int64_t)(k + KU + i) * H));
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
#pragma unroll
          for (int kk = 0; kk < KU; kk += 4) {
            const float4 hv = *reinterpret_cast<const float4*>(hn + r * H + k + kk);
            const float hk[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc[r].x = fmaf(wc[kk + i].x, hk[i], acc[r].x);
              acc[r].y = fmaf(wc[kk + i].y, hk[i], acc[r].y);
              acc[r].z = fmaf(wc[kk + i].z, hk[i], acc[r].z);
              acc[r].w = fmaf(wc[kk + i].w, hk[i], acc[r].w);
            }
          }
        }
      }
#pragma unroll
      for (int r = 0; r < ROWS; ++r) *reinterpret_cast<float4*>(hs + r * H + 4 * j4) = acc[r];
    }
    __syncthreads();

    // ---- 3. pointer scores e[r][k] = w_t . tanh(q + keys[k] + key0[k]) + b_t ; one warp per (row, k)
    for (int task = warp; task < ROWS * N; task += DC_THREADS / 32) {
      const int r = task / N, k = task % N, g = r / W, wslot = r % W;
      if (g >= Gc || wslot >= nlive[g]) continue;
      uint32_t picked = 0;
      for (int i = 0; i < t; ++i) picked |= 1u << seq[r][i];
      if (picked >> k & 1u) {
        if (lane == 0) e[r][k] = -1e9f;
        continue;
      }
      const float* t4 = io.t4 + (b0 + g) * (int64_t)N * N * H4;
      const float* k0 = io.key0 + ((b0 + g) * N + k) * (int64_t)H;
      const float* a1 = t >= 1 ? t4 + ((int64_t)seq[r][t - 1] * N + k) * H4 : nullptr;
      const float* a2 = t >= 2 ? t4 + ((int64_t)seq[r][t - 2] * N + k) * H4 + H : nullptr;
      float part = 0.f;
      for (int d = lane; d < H; d += 32) {
        float fs = 0.f, gs = 0.f;
        for (int j = 0; j < N; ++j)
          if (j != k && !(picked >> j & 1u)) {
            fs += t4[((int64_t)k * N + j) * H4 + 2 * H + d];
            gs += t4[((int64_t)j * N + k) * H4 + 3 * H + d];
          }
        float key = fs / (float)N + gs / (float)N;
        if (a1) key += a1[d];
        if (a2) key += a2[d];
        part = fmaf(w.wt[d], tanhf(hs[r * H + d] + key + k0[d]), part);
      }
      part = warp_sum(part);
      if (lane == 0) e[r][k] = part + w.bt;
    }
    __syncthreads();

    // ---- 4. log-softmax per live row; e <- candidate cost J - logp
    if (tid < ROWS) {
      const int r = tid, g = r / W, wslot = r % W;
      if (g < Gc && wslot < nlive[g]) {
        float mx = -INFINITY, s = 0.f;
        for (int k = 0; k < N; ++k) mx = fmaxf(mx, e[r][k]);
        for (int k = 0; k < N; ++k) s += expf(e[r][k] - mx);
        const float lse = logf(s);
        for (int k = 0; k < N; ++k) {
          const float logp = (e[r][k] - mx) - lse;
          if (io.trace_logp) io.trace_logp[(((b0 + g) * (N - 1) + t) * W + wslot) * N + k] = logp;
          e[r][k] = -logp + cost[r];
        }
      }
    }
    __syncthreads();

    // ---- 5. per-manual top-k (ascending cost, ties -> lowest flat index) by rank counting
    for (int c = tid; c < Gc * W * N; c += DC_THREADS) {
      const int g = c / (W * N), flat = c % (W * N), wslot = flat / N, k = flat % N;
      const int live = nlive[g];
      if (wslot >= live) continue;
      const int numel = live * N;
      int kk = min(W, numel);
      const float mine = e[g * W + wslot][k];
      int rank = 0;
      if (io.forced) {
        // teacher forcing (training loss, modeling_bert.py:998-1078): the single hypothesis follows the target order
        kk = 1;
        rank = (wslot == 0 && k == io.forced[(b0 + g) * N + t]) ? 0 : 1;
      } else {
        for (int o = 0; o < numel; ++o) {
          const float v = e[g * W + o / N][o % N];
          rank += (v < mine) || (v == mine && o < flat);
        }
      }
      if (rank < kk) {
        const int nr = g * W + rank;
        parent[nr] = g * W + wslot;
        ncost[nr] = mine;
        for (int i = 0; i < t; ++i) nseq[nr][i] = seq[g * W + wslot][i];
        nseq[nr][t] = (uint8_t)k;
        if (io.trace_ix) io.trace_ix[((b0 + g) * (N - 1) + t) * W + rank] = flat;
        if (io.trace_cost) io.trace_cost[((b0 + g) * (N - 1) + t) * W + rank] = mine;
      }
      if (flat == 0) nnlive[g] = kk;
    }
    __syncthreads();

    // ---- 6. expand: re-gather (h', c) by parent; each thread owns whole columns -> no hazard
    for (int d = tid; d < H; d += DC_THREADS) {
      float hv[ROWS], cv[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const int g = r / W;
        const bool on = g < Gc && (r % W) < nnlive[g];
        hv[r] = on ? hn[parent[r] * H + d] : 0.f;
        cv[r] = on ? cs[parent[r] * H + d] : 0.f;
      }
#pragma unroll
      for (int r = 0; r < ROWS; ++r) { hs[r * H + d] = hv[r]; cs[r * H + d] = cv[r]; }
    }
    if (tid < ROWS) {
      const int g = tid / W;
      if (g < Gc && (tid % W) < nnlive[g]) {
        cost[tid] = ncost[tid];
        for (int i = 0; i <= t; ++i) seq[tid][i] = nseq[tid][i];
      }
    }
    __syncthreads();
    if (tid < ROWS) nlive[tid] = nnlive[tid];
    __syncthreads();
  }

  // ---- best hypothesis = slot 0 (lowest cost); the one unused index goes last (modeling_bert.py:1549-1550)
  if (tid < Gc) {
    const int r = tid * W;
    uint32_t picked = 0;
    for (int i = 0; i < N - 1; ++i) {
      io.perm[(b0 + tid) * N + i] = seq[r][i];
      picked |= 1u << seq[r][i];
    }
    int last = 0;
    while (last < N - 1 && (picked >> last & 1u)) ++last;
    io.perm[(b0 + tid) * N + N - 1] = last;
    if (io.final_cost) io.final_cost[b0 + tid] = cost[r];  // = sum_t -log p(pick_t): the NLL when teacher-forced
  }
}

template <int ROWS>
static int launch_beam(const DecodeWeights& w, const DecodeIO& io, int G, cudaStream_t st) {
  const size_t smem = (size_t)3 * ROWS * io.H * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    MSQ_CUDA(cudaFuncSetAttribute(beam_search_kernel<ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  MSQ_CUDA(launch_k(beam_search_kernel<ROWS>, dim3(ceil_div(io.B, G)), dim3(DC_THREADS), smem, st, w, io, G));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

int beam_search(const DecodeWeights& w, const DecodeIO& io, cudaStream_t st) {
  MSQ_REQUIRE(io.N >= 2 && io.N <= DC_MAXN, "beam_search: N=%d out of range [2,%d]", io.N, DC_MAXN);
  MSQ_REQUIRE(io.W >= 1 && io.W <= 16, "beam_search: beam width %d out of range [1,16]", io.W);
  MSQ_REQUIRE(io.H % 8 == 0 && io.H <= 1024, "beam_search: H=%d unsupported", io.H);
  if (io.B == 0) return MSQ_OK;
  // rows per CTA: enough manuals to reuse each weight row across ~16 beams, but never fewer CTAs than
  // needed to give every SM work when B is large.
  int G = 16 / io.W;
  if (G < 1) G = 1;
  while (G > 1 && ceil_div(io.B, G) < 148) G /= 2;
  const int rows = G * io.W;
  if (rows <= 4) return launch_beam<4>(w, io, G, st);
  if (rows <= 8) return launch_beam<8>(w, io, G, st);
  return launch_beam<16>(w, io, G, st);
}


// ---------------------------------------------------------------------------------------------------
// Step-by-step API parity: BertForOrdering.step with the reference's MATERIALISED per-beam tensors
// (models/berson/modeling_bert.py:1368-1402).  Not used by the fused search above; it exists so that the
// reference's own beam_search_pointer loop can drive the CUDA path one step at a time.
// ---------------------------------------------------------------------------------------------------

// rela_vec.masked_fill_(rela_mask == 0, 0)   (in place, 1385)
__global__ void step_zero_rela_kernel(float* __restrict__ rela, const uint8_t* __restrict__ rela_mask, int64_t cells, int D) {
  pdl_sync();
  const int64_t cell = blockIdx.x;
  if (cell >= cells || rela_mask[cell]) return;
  for (int d = threadIdx.x; d < D; d += blockDim.x) rela[cell * D + d] = 0.f;
}

// pw[b,k,:] = [left1 ; left2 ; forw ; back ; 0-pad]   (1381-1389); one block per (beam, k)
__global__ void __launch_bounds__(256) step_pw_kernel(const float* __restrict__ rela, const float* __restrict__ hist1,
                                                      const float* __restrict__ hist2, const uint8_t* __restrict__ l1,
                                                      const uint8_t* __restrict__ l2, int N, int D, int Kp4, float* __restrict__ pw) {
  pdl_sync();
  const int b = blockIdx.x / N, k = blockIdx.x % N;
  float* out = pw + (int64_t)blockIdx.x * Kp4;
  const int64_t base = (int64_t)b * N * N;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float a1 = 0.f, a2 = 0.f, f = 0.f, g = 0.f;
    for (int i = 0; i < N; ++i) {
      const int64_t ik = base + (int64_t)i * N + k, ki = base + (int64_t)k * N + i;
      if (l1[ik]) a1 += hist1[ik * D + d];
      if (l2[ik]) a2 += hist2[ik * D + d];
      f += rela[ki * D + d];   // mean over dim 2 (j) of row k
      g += rela[ik * D + d];   // mean over dim 1 (i) of column k
    }
    out[d] = a1;
    out[D + d] = a2;
    out[2 * D + d] = f / (float)N;
    out[3 * D + d] = g / (float)N;
  }
  for (int d = 4 * D + threadIdx.x; d < Kp4; d += blockDim.x) out[d] = 0.f;
}

// gates [Wb,4H] in torch order (i|f|g|o) -> h', c'
__global__ void step_lstm_kernel(const float* __restrict__ gates, const float* __restrict__ c_in, int64_t Wb, int H,
                                 float* __restrict__ h_out, float* __restrict__ c_out) {
  pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Wb * H) return;
  const int64_t b = i / H;
  const int u = (int)(i % H);
  const float* g = gates + b * 4 * H;
  const float ig = 1.f / (1.f + expf(-g[u])), fg = 1.f / (1.f + expf(-g[H + u]));
  const float gg = tanhf(g[2 * H + u]), og = 1.f / (1.f + expf(-g[3 * H + u]));
  const float c2 = fg * c_in[i] + ig * gg;
  c_out[i] = c2;
  h_out[i] = og * tanhf(c2);
}

// e = w_t . tanh(q + keys + key0) + b_t ; masked_fill(pointed, -1e9) ; log_softmax   (1392-1400); block per beam
__global__ void __launch_bounds__(256) step_score_kernel(const float* __restrict__ q, const float* __restrict__ keys,
                                                         const float* __restrict__ key0, const uint8_t* __restrict__ pointed,
                                                         const float* __restrict__ wt, float bt, int N, int H,
                                                         float* __restrict__ logp) {
  pdl_sync();
  __shared__ float e[DC_MAXN];
  const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < N; k += blockDim.x >> 5) {
    float part = 0.f;
    for (int d = lane; d < H; d += 32)
      part = fmaf(wt[d], tanhf(q[(int64_t)b * H + d] + keys[((int64_t)b * N + k) * H + d] + key0[(int64_t)k * H + d]), part);
    part = warp_sum(part);
    if (lane == 0) e[k] = pointed[(int64_t)b * N + k] ? -1e9f : part + bt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = -INFINITY, s = 0.f;
    for (int k = 0; k < N; ++k) mx = fmaxf(mx, e[k]);
    for (int k = 0; k < N; ++k) s += expf(e[k] - mx);
    const float lse = logf(s);
    for (int k = 0; k < N; ++k) logp[(int64_t)b * N + k] = (e[k] - mx) - lse;
  }
}

int decode_step_parts(const StepIO& io, cudaStream_t st) {
  MSQ_REQUIRE(io.N >= 2 && io.N <= DC_MAXN, "decode_step: N=%d out of range", io.N);
  const int D = io.H + 2;
  MSQ_CUDA(launch_k(step_zero_rela_kernel, dim3((unsigned)((int64_t)io.Wb * io.N * io.N)), dim3(128), 0, st, io.rela, io.rela_mask, (int64_t)io.Wb * io.N * io.N, D));
  MSQ_LAUNCH_CHECK();
  MSQ_CUDA(launch_k(step_pw_kernel, dim3((unsigned)(io.Wb * io.N)), dim3(256), 0, st, io.rela, io.hist1, io.hist2, io.l1, io.l2, io.N, D, io.Kp4, io.pw));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
int decode_step_lstm(const float* gates, const float* c_in, int64_t Wb, int H, float* h_out, float* c_out, cudaStream_t st) {
  MSQ_CUDA(launch_k(step_lstm_kernel, dim3(ceil_div(Wb * H, 256)), dim3(256), 0, st, gates, c_in, Wb, H, h_out, c_out));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
int decode_step_score(const float* q, const float* keys, const float* key0, const uint8_t* pointed, const float* wt, float bt,
                      int64_t Wb, int N, int H, float* logp, cudaStream_t st) {
  MSQ_CUDA(launch_k(step_score_kernel, dim3((unsigned)Wb), dim3(256), 0, st, q, keys, key0, pointed, wt, bt, N, H, logp));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}


// ---------------------------------------------------------------------------------------------------
// models/pointer_module.py p1 path: LSTMPointerModule.forward (690-749) over LSTMDecoder (651-678) and
// LSTMAttention (616-648).  Dead code in the reference but named by the scope: greedy
// pointer decoding WITHOUT a permutation mask, additive attention V.tanh(W1 e + W2 h), LSTM input
// [context ; previous pick], cross-entropy against y at every step.  One CTA per manual.
// ---------------------------------------------------------------------------------------------------
constexpr int PM_MAXN = 16, PM_MAXU = 32;

__global__ void __launch_bounds__(256) pointer_p1_kernel(const float* __restrict__ enc, const float* __restrict__ cls,
                                                         const int64_t* __restrict__ y, const float* __restrict__ W1,
                                                         const float* __restrict__ W2, const float* __restrict__ V,
                                                         const float* __restrict__ Wih, const float* __restrict__ Whh,
                                                         const float* __restrict__ bih, const float* __restrict__ bhh, int N, int H,
                                                         int U, float* __restrict__ preds, float* __restrict__ ce_sum) {
  pdl_sync();
  extern __shared__ float sm[];
  float* h = sm;            // [H]
  float* c = h + H;         // [H]
  float* x = c + H;         // [2H]  = [context ; dec_in]
  float* gates = x + 2 * H; // [4H]
  __shared__ float e1[PM_MAXN][PM_MAXU], e2[PM_MAXU], uj[PM_MAXN], aj[PM_MAXN];
  __shared__ int pred;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const float* eb = enc + (int64_t)b * N * H;
  for (int d = tid; d < H; d += blockDim.x) {
    h[d] = c[d] = cls[(int64_t)b * H + d];  // hs = (encoder_cls, encoder_cls) (700-701)
    x[H + d] = cls[(int64_t)b * H + d];     // dec_in = encoder_cls (712)
  }
  for (int nu = warp; nu < N * U; nu += nw) {  // W1 e_n is step independent
    const int n = nu / U, u = nu % U;
    float a = 0.f;
    for (int d = lane; d < H; d += 32) a = fmaf(W1[(int64_t)u * H + d], eb[(int64_t)n * H + d], a);
    a = warp_sum(a);
    if (lane == 0) e1[n][u] = a;
  }
  __syncthreads();
  float ce = 0.f;
  for (int t = 0; t < N; ++t) {
    for (int u = warp; u < U; u += nw) {  // W2 h_t with the hidden state BEFORE this step's LSTM update (663)
      float a = 0.f;
      for (int d = lane; d < H; d += 32) a = fmaf(W2[(int64_t)u * H + d], h[d], a);
      a = warp_sum(a);
      if (lane == 0) e2[u] = a;
    }
    __syncthreads();
    if (tid < N) {
      float a = 0.f;
      for (int u = 0; u < U; ++u) a = fmaf(V[u], tanhf(e1[tid][u] + e2[u]), a);
      uj[tid] = a;
    }
    __syncthreads();
    if (tid == 0) {
      float mx = -INFINITY, s = 0.f;
      int arg = 0;
      for (int n = 0; n < N; ++n) if (uj[n] > mx) { mx = uj[n]; arg = n; }
      for (int n = 0; n < N; ++n) { aj[n] = expf(uj[n] - mx); s += aj[n]; }
      for (int n = 0; n < N; ++n) aj[n] /= s;
      pred = arg;                                                  // softmax(att_w).argmax (726)
      ce += -((uj[(int)y[(int64_t)b * N + t]] - mx) - logf(s));    // F.cross_entropy(att_w, y[:, t]) (742)
      preds[(int64_t)b * N + t] = (float)arg;
    }
    __syncthreads();
    for (int d = tid; d < H; d += blockDim.x) {
      float a = 0.f;
      for (int n = 0; n < N; ++n) a = fmaf(aj[n], eb[(int64_t)n * H + d], a);
      x[d] = a;  // di_prime (645-646); x[H..2H) still holds the current dec_in
    }
    __syncthreads();
    for (int row = warp; row < 4 * H; row += nw) {  // gates = W_ih [di ; x] + b_ih + W_hh h + b_hh
      float a = 0.f;
      const float* wi = Wih + (int64_t)row * 2 * H;
      const float* wh = Whh + (int64_t)row * H;
      for (int d = lane; d < 2 * H; d += 32) a = fmaf(wi[d], x[d], a);
      for (int d = lane; d < H; d += 32) a = fmaf(wh[d], h[d], a);
      a = warp_sum(a);
      if (lane == 0) gates[row] = a + bih[row] + bhh[row];
    }
    __syncthreads();
    for (int d = tid; d < H; d += blockDim.x) {
      const float ig = 1.f / (1.f + expf(-gates[d])), fg = 1.f / (1.f + expf(-gates[H + d]));
      const float gg = tanhf(gates[2 * H + d]), og = 1.f / (1.f + expf(-gates[3 * H + d]));
      const float c2 = fg * c[d] + ig * gg;
      c[d] = c2;
      h[d] = og * tanhf(c2);
      x[H + d] = eb[(int64_t)pred * H + d];  // next dec_in = encoder_out[pred] (735-737)
    }
    __syncthreads();
  }
  if (tid == 0) ce_sum[b] = ce;
}

// batch_loss = (sum_t mean_b CE_t) / B   (742, 747: the reference divides by the batch size twice)
__global__ void pointer_p1_loss_kernel(const float* __restrict__ ce_sum, int64_t B, float* __restrict__ loss) {
  pdl_sync();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int64_t b = 0; b < B; ++b) s += ce_sum[b];
    loss[0] = s / (float)B / (float)B;
  }
}

int pointer_p1(const float* enc, const float* cls, const int64_t* y, const float* W1, const float* W2, const float* V,
               const float* Wih, const float* Whh, const float* bih, const float* bhh, int64_t B, int N, int H, int U, float* preds,
               float* ce_scratch, float* loss, cudaStream_t st) {
  MSQ_REQUIRE(N >= 1 && N <= PM_MAXN && U >= 1 && U <= PM_MAXU && H <= 2048, "pointer_p1: N=%d U=%d H=%d out of range", N, U, H);
  if (B == 0) return MSQ_OK;
  const size_t smem = (size_t)8 * H * sizeof(float);
  static size_t configured = 0;
  if (smem > configured) {
    MSQ_CUDA(cudaFuncSetAttribute(pointer_p1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  MSQ_CUDA(launch_k(pointer_p1_kernel, dim3((unsigned)B), dim3(256), smem, st, enc, cls, y, W1, W2, V, Wih, Whh, bih, bhh, N, H, U, preds,
                    ce_scratch));
  MSQ_LAUNCH_CHECK();
  MSQ_CUDA(launch_k(pointer_p1_loss_kernel, dim3(1), dim3(32), 0, st, (const float*)ce_scratch, B, loss));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

}  // namespace msq
