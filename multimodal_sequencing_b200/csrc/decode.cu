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
#include <stdlib.h>

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
  MSQ_SMEM_ATTR(smem, beam_search_kernel<ROWS>);
  MSQ_CUDA(launch_k(beam_search_kernel<ROWS>, dim3(ceil_div(io.B, G)), dim3(DC_THREADS), smem, st, w, io, G));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// The single-kernel form above (round 1) keeps the whole search of a few manuals inside one CTA: every CTA streams the full
// W_hh / W_q (11.8 MB) from L2 at every step for at most 16 rows.  It is kept for B*W <= 16 rows only when MSQ_DECODE_FUSED=1.
static int beam_search_fused(const DecodeWeights& w, const DecodeIO& io, cudaStream_t st) {
  int G = 16 / io.W;
  if (G < 1) G = 1;
  while (G > 1 && ceil_div(io.B, G) < 148) G /= 2;
  const int rows = G * io.W;
  if (rows <= 4) return launch_beam<4>(w, io, G, st);
  if (rows <= 8) return launch_beam<8>(w, io, G, st);
  return launch_beam<16>(w, io, G, st);
}

// ===================================================================================================
// Tiled decode (round 2): the search is a sequence of three kernels per decode step over ALL B*live rows
//   1. dec_gemm_kernel<.., 0>  gates = XG[prev] + h W_hh^T as a rows x 4H fp32 GEMM tiled over (rows, gate columns): a weight
//                              tile is read once per 64/128 rows (not once per <= 16 rows), h rows are gathered through the
//                              parent table (the beam re-gather costs nothing), the LSTM cell is the epilogue;
//   2. dec_gemm_kernel<.., 1>  q = h' W_q^T + b_q, same kernel;
//   3. dec_select_kernel       one CTA per manual: F / G rows of T4 staged in shared memory once per candidate step k and
//                              shared by all beams, pointer scores, log-softmax, rank-counting top-k, new (sequence, cost,
//                              parent) tables -- and the permutation after the last step.
// Arithmetic is the fused kernel's, operation for operation: accumulators start from XG / b_q and add w*h in ascending k
// with fmaf; F / G sums ascend in j; the score dot product is lane-strided with an xor-tree warp sum.  Results are therefore
// bit-identical to the fused kernel (and to what the golden beam traces pin).
// ===================================================================================================
struct DecState {
  float* h[2];        // [B*W, H] h' of the previous / current step
  float* c[2];        // [B*W, H]
  const float* q;     // query rows: row (b * q_rows + beam), leading dimension q_ld
  int q_ld, q_rows;
  uint8_t* seq[2];    // [B*W, 16] picked steps so far (ping-pong)
  float* cost[2];     // [B*W]
  int32_t* parent;    // [B*W] beam slot (of the previous step) a row descends from
};

// FFMA form: h[2], c[2], q (5 x rows*H fp32).  Tensor-core form: h' planes (rows * 3H bf16), c[2], [q | h W_hh^T][2]
// (rows * 5H fp32 each), h0 planes.  One size covers both.
size_t beam_search_state_bytes(int64_t B, int W, int H) {
  const size_t rows = (size_t)B * W, al = 255;
  const size_t tc = ((rows * 3 * H * 2 + al) & ~al) + 2 * ((rows * H * 4 + al) & ~al) + 2 * ((rows * 5 * H * 4 + al) & ~al) +
                    (((size_t)B * 3 * H * 2 + al) & ~al);
  return tc + 2 * ((rows * 16 + al) & ~al) + 3 * ((rows * 4 + al) & ~al) + 1024;
}
size_t beam_search_scratch_bytes(int64_t B, int N, int W, int H) {
  const size_t Kp = ((size_t)(H + 2) + 63) / 64 * 64;
  return beam_search_state_bytes(B, W, H) + ((((size_t)B * (N + 1) * 3 * H + 127) & ~size_t(127)) + (size_t)B * N * N * 3 * Kp) * 2 + 1024;
}
int decode_trunc(int precise) {
  // default: bf16x3 in both tensor-core modes.  precise = 2 carries every encoder operand as hi + lo already (the decoder's
  // inputs arrive with ~2e-5 relative error), precise = 0 is plain bf16 upstream; the fp32 parity mode never comes here
  // (FFMA decode).  Beam indices stay bit-exact vs the reference fixtures either way (tests/test_gpu_parity.py).
  // Read per call (two getenv per decode) so that a test can run both forms in one process.
  (void)precise;
  const char* e = getenv("MSQ_DEC_X3");
  return e ? (atoi(e) != 0) : 1;
}

bool decode_tc_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MSQ_DECODE_TC"); v = (e && e[0] == '0') ? 0 : 1; }
  return v != 0;
}

struct DecGemm {
  const float* Wt;      // [H][ldw] k-major (whh_t: ldw = 4H gate-interleaved; wq_t: ldw = H)
  int ldw, Ncols;
  const float* h_src;   // MODE 0: h' of the previous step (rows gathered by parent), t == 0: h0 [B, H];  MODE 1: h' of this step
  const float* c_src;   // MODE 0: c of the previous step (t == 0: null -> 0)
  const int32_t* parent;
  const uint8_t* seq;   // current sequences (token t-1 selects the XG row)
  const float* xg;      // [B, N+1, 4H]
  const float* bq;
  float* h_out;         // MODE 0: h' ; MODE 1: q
  float* c_out;
  int64_t M;            // B * live
  int live, W, N, H, t;
};

// TILE x TILE x 16 fp32 tiles, 256 threads, (TILE/16)^2 register micro-tile, register-prefetch double buffering
template <int TILE, int MODE>
__global__ void __launch_bounds__(256) dec_gemm_kernel(DecGemm a) {
  pdl_sync();
  constexpr int BK = 16, LD = TILE + 4, TM = TILE / 16, KV = TILE == 128 ? 8 : 4, BV = TILE == 128 ? 2 : 1;
  __shared__ __align__(16) float As[2][BK][LD];
  __shared__ __align__(16) float Bs[2][BK][LD];
  const int tid = threadIdx.x, H = a.H;
  const int64_t m0 = (int64_t)blockIdx.y * TILE;
  const int n0 = blockIdx.x * TILE;
  // A staging: thread owns row lrow, KV consecutive k
  const int lrow = tid % TILE, lk = (tid / TILE) * KV;
  const int64_t am = m0 + lrow;
  const bool a_ok = am < a.M;
  const float* ap = nullptr;
  if (a_ok) {
    const int64_t b = am / a.live;
    const int wslot = (int)(am % a.live);
    if (MODE == 0) ap = a.t == 0 ? a.h_src + b * H : a.h_src + (b * a.W + a.parent[b * a.W + wslot]) * H;
    else ap = a.h_src + (b * a.W + wslot) * H;
    ap += lk;
  }
  // B staging: BV float4 per thread; element e = tid + 256 * i covers k = e / (TILE/4), columns 4 * (e % (TILE/4))
  const int ty = tid >> 4, tx = tid & 15;

  auto row_of = [&](int i) -> int64_t { return m0 + (TILE == 128 ? (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4)) : ty * 4 + i); };
  auto col_of = [&](int jh) -> int { return n0 + jh * 64 + tx * 4; };

  float acc[TM][TM];
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t row = row_of(i);
    const bool ok = row < a.M;
    const float* init = nullptr;
    if (ok) {
      if (MODE == 0) {
        const int64_t b = row / a.live;
        const int wslot = (int)(row % a.live);
        const int prev = a.t == 0 ? a.N : a.seq[(b * a.W + wslot) * DC_MAXN + a.t - 1];
        init = a.xg + (b * (a.N + 1) + prev) * (int64_t)(4 * H);
      } else {
        init = a.bq;
      }
    }
#pragma unroll
    for (int jh = 0; jh < TM / 4; ++jh) {
      const int col = col_of(jh);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok && col < a.Ncols) v = *reinterpret_cast<const float4*>(init + col);
      acc[i][jh * 4 + 0] = v.x; acc[i][jh * 4 + 1] = v.y; acc[i][jh * 4 + 2] = v.z; acc[i][jh * 4 + 3] = v.w;
    }
  }

  float ra[KV];
  float4 rb[BV];
  auto load_a = [&](int kt) {
    if (a_ok) {
      const float4 x = *reinterpret_cast<const float4*>(ap + kt * BK);
      ra[0] = x.x; ra[1] = x.y; ra[2] = x.z; ra[3] = x.w;
      if (KV == 8) { const float4 y = *reinterpret_cast<const float4*>(ap + kt * BK + 4); ra[4] = y.x; ra[5] = y.y; ra[6] = y.z; ra[7] = y.w; }
    } else {
#pragma unroll
      for (int i = 0; i < KV; ++i) ra[i] = 0.f;
    }
  };
  auto load_b = [&](int kt) {
#pragma unroll
    for (int i = 0; i < BV; ++i) {
      const int e = tid + 256 * i, k = e / (TILE / 4), c = 4 * (e % (TILE / 4));
      rb[i] = (n0 + c < a.Ncols) ? __ldg(reinterpret_cast<const float4*>(a.Wt + (int64_t)(kt * BK + k) * a.ldw + n0 + c))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int i = 0; i < KV; ++i) As[buf][lk + i][lrow] = ra[i];
#pragma unroll
    for (int i = 0; i < BV; ++i) {
      const int e = tid + 256 * i, k = e / (TILE / 4), c = 4 * (e % (TILE / 4));
      *reinterpret_cast<float4*>(&Bs[buf][k][c]) = rb[i];
    }
  };
  const int nk = H / BK;
  load_a(0); load_b(0); stage(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) { load_a(kt + 1); load_b(kt + 1); }
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float av[TM], bv[TM];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][k][ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][k][tx * 4]);
      av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
      bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
      if (TILE == 128) {
        const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][k][64 + ty * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][k][64 + tx * 4]);
        av[TM - 4] = a1.x; av[TM - 3] = a1.y; av[TM - 2] = a1.z; av[TM - 1] = a1.w;
        bv[TM - 4] = b1.x; bv[TM - 3] = b1.y; bv[TM - 2] = b1.z; bv[TM - 1] = b1.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TM; ++j) acc[i][j] = fmaf(bv[j], av[i], acc[i][j]);
    }
    if (kt + 1 < nk) stage(cur ^ 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t row = row_of(i);
    if (row >= a.M) continue;
    const int64_t b = row / a.live;
    const int wslot = (int)(row % a.live);
    const int64_t orow = b * a.W + wslot;
#pragma unroll
    for (int jh = 0; jh < TM / 4; ++jh) {
      const int col = col_of(jh);
      if (col >= a.Ncols) continue;
      if (MODE == 0) {
        // columns 4u .. 4u+3 are the (i, f, g, o) pre-activations of hidden unit u
        const int u = col >> 2;
        const float gi = acc[i][jh * 4 + 0], gf = acc[i][jh * 4 + 1], gc = acc[i][jh * 4 + 2], go = acc[i][jh * 4 + 3];
        const float ig = 1.f / (1.f + expf(-gi)), fg = 1.f / (1.f + expf(-gf));
        const float gg = tanhf(gc), og = 1.f / (1.f + expf(-go));
        const float cin = a.c_src ? a.c_src[(b * a.W + a.parent[orow]) * H + u] : 0.f;
        const float c2 = fg * cin + ig * gg;
        a.c_out[orow * H + u] = c2;
        a.h_out[orow * H + u] = og * tanhf(c2);
      } else {
        *reinterpret_cast<float4*>(a.h_out + orow * H + col) =
            make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      }
    }
  }
}

template <int MODE>
static int launch_dec_gemm(const DecGemm& a, cudaStream_t st) {
  // 64 x 64 tiles until there are enough 128 x 128 tiles for two per SM
  if ((int64_t)ceil_div(a.Ncols, 128) * ceil_div(a.M, 128) < 2 * 148) {
    MSQ_CUDA(launch_k(dec_gemm_kernel<64, MODE>, dim3(ceil_div(a.Ncols, 64), ceil_div(a.M, 64)), dim3(256), 0, st, a));
  } else {
    MSQ_CUDA(launch_k(dec_gemm_kernel<128, MODE>, dim3(ceil_div(a.Ncols, 128), ceil_div(a.M, 128)), dim3(256), 0, st, a));
  }
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// tanh x = sign(x) (1 - 2 / (e^{2|x|} + 1)) on the SFU: ex2.approx (2 ulp) and rcp.approx (1 ulp) give an ABSOLUTE error of
// <= ~2e-7 everywhere (libm's tanhf: 2 ulp, i.e. up to 1.2e-7 near 1) in 7 straight-line instructions instead of tanhf's two
// divergent branches (polynomial below 0.55, exp + IEEE division above): dec_select is issue-bound and tanh was ~half of it.
__device__ __forceinline__ float tanh_sfu(float x) {
  float t, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(x) * 2.8853900817779268f));   // e^{2|x|}; +inf for large |x| -> r = 0
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
  return copysignf(fmaf(-2.0f, r, 1.0f), x);
}

// One CTA per manual: pointer scores of every live beam over the N candidate steps, log-softmax, top-k, new tables.
// SFU = true (tensor-core form): tanh_sfu; false (fp32 parity form): libm tanhf, bit-identical to the round-1 kernel.
template <bool SFU>
__global__ void __launch_bounds__(512, 1) dec_select_kernel(DecodeWeights w, DecodeIO io, DecState s, int t, int live, int cur, int nbuf) {
  pdl_sync();
  extern __shared__ __align__(16) float fg_s[];   // 2 x (F rows [N][H] | G rows [N][H]): candidate step being scored + the next
  const int H = io.H, N = io.N, W = io.W, H4 = 4 * H;
  __shared__ float e[16][DC_MAXN];
  __shared__ float ep[8][16];   // blocked form: partial dot product of a 128-feature chunk (H <= 1024) per beam
  __shared__ float cost[16];
  __shared__ uint8_t seq[16][DC_MAXN];
  __shared__ uint32_t pickm[16];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nthr = blockDim.x;
  const int64_t b = blockIdx.x;
  // x / N as a correctly rounded quotient without the division subroutine: q0 = RN(x rN), r = x - q0 N (exact, fma),
  // q = RN(q0 + r rN) (Markstein; rN = RN(1/N); operands are O(1), far from under/overflow): bit-identical to x / (float)N
  const float fN = (float)N, rN = 1.0f / fN;
  auto div_n = [&](float x) { const float q0 = x * rN; return fmaf(fmaf(-q0, fN, x), rN, q0); };
  const uint8_t* seq_g = s.seq[cur] + b * W * DC_MAXN;
  if (tid < live) {
    uint32_t picked = 0;
    for (int i = 0; i < t; ++i) { const uint8_t v = seq_g[tid * DC_MAXN + i]; seq[tid][i] = v; picked |= 1u << v; }
    pickm[tid] = picked;
    cost[tid] = t == 0 ? 0.f : s.cost[cur][b * W + tid];
  }
  __syncthreads();
  const float* t4 = io.t4 + b * (int64_t)N * N * H4;
  uint32_t any_rem = 0;   // steps still unpicked in at least one live beam
  for (int i = 0; i < live; ++i) any_rem |= ~pickm[i];

  // ---- F[k, j, :] and G[j, k, :] for every j some beam still needs are staged with cp.async, one candidate step ahead
  // (double buffer): read once per manual and step, shared by all beams
  auto stage = [&](int k, int buf) {
    float* F = fg_s + (size_t)buf * 2 * N * H;
    float* G = F + (size_t)N * H;
    // row loop outside (block-uniform skip), 16-byte column loop inside: no integer division per copy -- the flat-index
    // form of this loop (idx / (H/4), idx % (H/4) with a run-time H) was ~100 instructions per pair of copies, more than the
    // scoring itself at small `live` (ncu: 2.7e7 warp instructions at t = 0, where one beam is scored)
    const uint32_t fbase = (uint32_t)__cvta_generic_to_shared(F), gbase = (uint32_t)__cvta_generic_to_shared(G);
    for (int j = 0; j < N; ++j) {
      if (j == k || !(any_rem >> j & 1u)) continue;
      const float* fsrc = t4 + ((int64_t)k * N + j) * H4 + 2 * H;
      const float* gsrc = t4 + ((int64_t)j * N + k) * H4 + 3 * H;
      for (int d = 4 * tid; d < H; d += 4 * nthr) {
        const uint32_t off = (uint32_t)(j * H + d) * 4u;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(fbase + off), "l"(fsrc + d) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(gbase + off), "l"(gsrc + d) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // nbuf == 2: candidate step k+1 is staged while k is scored (few CTAs: latency matters); nbuf == 1: half the shared memory,
  // more CTAs per SM hide the staging instead (many manuals: throughput matters)
  if (nbuf == 2) stage(0, 0);
  for (int k = 0; k < N; ++k) {
    if (nbuf == 2 && k + 1 < N) {
      stage(k + 1, (k + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      if (nbuf == 1) stage(k, 0);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* Fs = fg_s + (size_t)(nbuf == 2 ? (k & 1) : 0) * 2 * N * H;
    const float* Gs = Fs + (size_t)N * H;
    if constexpr (SFU) {
      // Blocked form.  The warp-per-(beam, k) loop below re-reads every staged F / G row once per BEAM: W x N x |Rem| x 2 x H x 4
      // bytes of shared-memory loads per manual and step (7.8 MB at N = 10, W = 16: ~61 k cycles of the SM's 128 B/clk, the
      // real bound of that form).  Here a work item is (128-feature chunk, group of BB beams): an F / G row chunk is loaded once
      // and added into the accumulators of the BB beams that need it, so the shared-memory traffic drops BB-fold and every warp
      // has work whatever `live` is.  Per-chunk partial dot products meet in ep[chunk][beam] and are summed in chunk order
      // (deterministic; the order differs from the parity form's lane-strided sum, both are fp32 sums of the same terms).
      constexpr int BB = 4;
      const int nchunk = (H + 127) >> 7, ngroup = (live + BB - 1) / BB;
      for (int item = warp; item < nchunk * ngroup; item += nthr >> 5) {
        const int c = item % nchunk, g0 = (item / nchunk) * BB;
        const int d0 = 128 * c + 4 * lane;
        const bool in = d0 < H;
        uint32_t need[BB], uni = 0;
#pragma unroll
        for (int bb = 0; bb < BB; ++bb) {
          const uint32_t picked = g0 + bb < live ? pickm[g0 + bb] : ~0u;
          need[bb] = (picked >> k & 1u) ? 0u : (~picked & ~(1u << k) & ((1u << N) - 1u));
          uni |= need[bb];
        }
        // beam bb of the group is live and has not picked k (need == 0 alone does not say so: the last unpicked step has an
        // empty remainder)
        auto need_score = [&](int, int wslot) { return wslot < live && !(pickm[wslot] >> k & 1u); };
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        // every global operand of the item is requested BEFORE the shared-memory sums: the kernel runs 8-16 warps per SM, so a
        // load issued where it is used costs its whole DRAM / L2 latency (that, not the instruction count, bounded this kernel)
        float4 qv[BB], av[BB];
#pragma unroll
        for (int bb = 0; bb < BB; ++bb) {
          const int wslot = g0 + bb;
          qv[bb] = av[bb] = z4;
          if (need_score(bb, wslot) && in) {
            qv[bb] = *reinterpret_cast<const float4*>(s.q + (b * s.q_rows + wslot) * (int64_t)s.q_ld + d0);
            if (t >= 1) av[bb] = __ldg(reinterpret_cast<const float4*>(t4 + ((int64_t)seq[wslot][t - 1] * N + k) * H4 + d0));
            if (t >= 2) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(t4 + ((int64_t)seq[wslot][t - 2] * N + k) * H4 + H + d0));
              av[bb].x += a.x; av[bb].y += a.y; av[bb].z += a.z; av[bb].w += a.w;
            }
          }
        }
        const float4 kv = in ? __ldg(reinterpret_cast<const float4*>(io.key0 + (b * N + k) * (int64_t)H + d0)) : z4;
        const float4 wv = in ? __ldg(reinterpret_cast<const float4*>(w.wt + d0)) : z4;
        float4 fs[BB], gs[BB];
#pragma unroll
        for (int bb = 0; bb < BB; ++bb) fs[bb] = gs[bb] = z4;
        if (in) {
          // rows of `uni` in ascending order (the reference's summation order), the NEXT row's chunk loaded before the current
          // one is added: the unrolled `if (uni >> j & 1)` form left every row's shared-memory latency exposed (ncu: a third of
          // the stall samples on the first FADD of a row)
          uint32_t u = uni;
          int j = -1;
          float4 f = z4, g = z4;
          if (u) {
            j = __ffs(u) - 1; u &= u - 1;
            f = *reinterpret_cast<const float4*>(Fs + (size_t)j * H + d0); g = *reinterpret_cast<const float4*>(Gs + (size_t)j * H + d0);
          }
          while (j >= 0) {
            int jn = -1;
            float4 fn = z4, gn = z4;
            if (u) {
              jn = __ffs(u) - 1; u &= u - 1;
              fn = *reinterpret_cast<const float4*>(Fs + (size_t)jn * H + d0); gn = *reinterpret_cast<const float4*>(Gs + (size_t)jn * H + d0);
            }
#pragma unroll
            for (int bb = 0; bb < BB; ++bb) {
              if (need[bb] >> j & 1u) {
                fs[bb].x += f.x; fs[bb].y += f.y; fs[bb].z += f.z; fs[bb].w += f.w;
                gs[bb].x += g.x; gs[bb].y += g.y; gs[bb].z += g.z; gs[bb].w += g.w;
              }
            }
            f = fn; g = gn; j = jn;
          }
        }
#pragma unroll
        for (int bb = 0; bb < BB; ++bb) {
          const int wslot = g0 + bb;
          if (wslot >= live) break;
          float part = 0.f;
          if (need_score(bb, wslot) && in) {
            // key = mean F + mean G + (a1 + a2): a1 + a2 are added first here (the parity form adds them one after the other)
            const float kx = div_n(fs[bb].x) + div_n(gs[bb].x) + av[bb].x, ky = div_n(fs[bb].y) + div_n(gs[bb].y) + av[bb].y;
            const float kz = div_n(fs[bb].z) + div_n(gs[bb].z) + av[bb].z, kw = div_n(fs[bb].w) + div_n(gs[bb].w) + av[bb].w;
            part = fmaf(wv.x, tanh_sfu(qv[bb].x + kx + kv.x), part);
            part = fmaf(wv.y, tanh_sfu(qv[bb].y + ky + kv.y), part);
            part = fmaf(wv.z, tanh_sfu(qv[bb].z + kz + kv.z), part);
            part = fmaf(wv.w, tanh_sfu(qv[bb].w + kw + kv.w), part);
          }
          part = warp_sum(part);
          if (lane == 0) ep[c][wslot] = part;
        }
      }
      __syncthreads();
      if (tid < live) {
        float sum = 0.f;
        for (int c = 0; c < nchunk; ++c) sum += ep[c][tid];
        e[tid][k] = (pickm[tid] >> k & 1u) ? -1e9f : sum + w.bt;
      }
    } else {
      // A warp scores candidate step k for one beam.  Lane-strided over d in batches of 4 (the global operands of a batch are
      // loaded before any is used); the summation order is the reference order: j ascending inside fs / gs, d ascending
      // inside the dot product, xor-tree over lanes.
      for (int wslot = warp; wslot < live; wslot += nthr >> 5) {
        const uint32_t picked = pickm[wslot];
        if (picked >> k & 1u) {
          if (lane == 0) e[wslot][k] = -1e9f;
          continue;
        }
        const uint32_t need = ~picked & ~(1u << k) & ((1u << N) - 1u);
        const float* qp = s.q + (b * s.q_rows + wslot) * (int64_t)s.q_ld;
        const float* a1 = t >= 1 ? t4 + ((int64_t)seq[wslot][t - 1] * N + k) * H4 : nullptr;
        const float* a2 = t >= 2 ? t4 + ((int64_t)seq[wslot][t - 2] * N + k) * H4 + H : nullptr;
        const float* k0 = io.key0 + (b * N + k) * (int64_t)H;
        float part = 0.f;
        // four consecutive features per lane and 128 per warp pass (H % 4 == 0, every row 16-byte aligned): one 16-byte load per
        // operand instead of four 4-byte loads -- the kernel is issue-bound, and loads were ~40 % of its instructions
        for (int d0 = 4 * lane; d0 < H; d0 += 128) {
          const float4 kv = __ldg(reinterpret_cast<const float4*>(k0 + d0)), wv = __ldg(reinterpret_cast<const float4*>(w.wt + d0));
          const float4 qv = *reinterpret_cast<const float4*>(qp + d0);
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 a1v = a1 ? __ldg(reinterpret_cast<const float4*>(a1 + d0)) : z4, a2v = a2 ? __ldg(reinterpret_cast<const float4*>(a2 + d0)) : z4;
          float4 fs = z4, gs = z4;
  #pragma unroll
          for (int j = 0; j < DC_MAXN; ++j) {
            if (j < N && (need >> j & 1u)) {
              const float4 f = *reinterpret_cast<const float4*>(Fs + (size_t)j * H + d0), g = *reinterpret_cast<const float4*>(Gs + (size_t)j * H + d0);
              fs.x += f.x; fs.y += f.y; fs.z += f.z; fs.w += f.w;
              gs.x += g.x; gs.y += g.y; gs.z += g.z; gs.w += g.w;
            }
          }
          float4 key = make_float4(div_n(fs.x) + div_n(gs.x), div_n(fs.y) + div_n(gs.y), div_n(fs.z) + div_n(gs.z), div_n(fs.w) + div_n(gs.w));
          if (a1) { key.x += a1v.x; key.y += a1v.y; key.z += a1v.z; key.w += a1v.w; }
          if (a2) { key.x += a2v.x; key.y += a2v.y; key.z += a2v.z; key.w += a2v.w; }
          auto th = [](float v) { return SFU ? tanh_sfu(v) : tanhf(v); };
          part = fmaf(wv.x, th(qv.x + key.x + kv.x), part);
          part = fmaf(wv.y, th(qv.y + key.y + kv.y), part);
          part = fmaf(wv.z, th(qv.z + key.z + kv.z), part);
          part = fmaf(wv.w, th(qv.w + key.w + kv.w), part);
        }
        part = warp_sum(part);
        if (lane == 0) e[wslot][k] = part + w.bt;
      }
    }
    __syncthreads();
  }

  // ---- log-softmax per live beam; e <- candidate cost
  if (tid < live) {
    float mx = -INFINITY, sum = 0.f;
    for (int k = 0; k < N; ++k) mx = fmaxf(mx, e[tid][k]);
    for (int k = 0; k < N; ++k) sum += expf(e[tid][k] - mx);
    const float lse = logf(sum);
    for (int k = 0; k < N; ++k) {
      const float logp = (e[tid][k] - mx) - lse;
      if (io.trace_logp) io.trace_logp[((b * (N - 1) + t) * W + tid) * N + k] = logp;
      e[tid][k] = -logp + cost[tid];
    }
  }
  __syncthreads();

  // ---- top-k by rank counting (ascending cost, ties -> lowest flat index beam*N+step)
  const int nxt = cur ^ 1;
  const int numel = live * N;
  const bool last = t == N - 2;
  for (int flat = tid; flat < numel; flat += nthr) {
    const int wslot = flat / N, k = flat % N;
    int kk = min(W, numel);
    const float mine = e[wslot][k];
    int rank = 0;
    if (io.forced) {
      kk = 1;   // teacher forcing (modeling_bert.py:998-1078): the single hypothesis follows the target order
      rank = (wslot == 0 && k == io.forced[b * N + t]) ? 0 : 1;
    } else {
      for (int ob = 0, o = 0; ob < live; ++ob) {   // o = ob * N + ok without a division per element
        for (int ok = 0; ok < N; ++ok, ++o) {
          const float v = e[ob][ok];
          rank += (v < mine) || (v == mine && o < flat);
        }
      }
    }
    if (rank < kk) {
      const int64_t nr = b * W + rank;
      if (io.trace_ix) io.trace_ix[(b * (N - 1) + t) * W + rank] = flat;
      if (io.trace_cost) io.trace_cost[(b * (N - 1) + t) * W + rank] = mine;
      if (!last) {
        s.parent[nr] = wslot;
        s.cost[nxt][nr] = mine;
        uint8_t* ns = s.seq[nxt] + nr * DC_MAXN;
        for (int i = 0; i < t; ++i) ns[i] = seq[wslot][i];
        ns[t] = (uint8_t)k;
      } else if (rank == 0) {
        // best hypothesis; the one unused index goes last (modeling_bert.py:1549-1550)
        uint32_t picked = pickm[wslot] | (1u << k);
        for (int i = 0; i < t; ++i) io.perm[b * N + i] = seq[wslot][i];
        io.perm[b * N + t] = k;
        int rest = 0;
        while (rest < N - 1 && (picked >> rest & 1u)) ++rest;
        io.perm[b * N + N - 1] = rest;
        if (io.final_cost) io.final_cost[b] = mine;
      }
    }
  }
}

static int next_live(int live, int N, int W) { return min(W, live * N); }

static int beam_search_tiled(const DecodeWeights& w, const DecodeIO& io, cudaStream_t st) {
  MSQ_REQUIRE(io.scratch != nullptr, "beam_search: scratch buffer missing");
  const int H = io.H, N = io.N, W = io.W;
  const size_t rows = (size_t)io.B * W;
  char* p = reinterpret_cast<char*>(io.scratch);
  auto carve = [&](size_t bytes) { char* r = p; p += (bytes + 255) & ~size_t(255); return r; };
  DecState s;
  for (int i = 0; i < 2; ++i) s.h[i] = (float*)carve(rows * H * 4);
  for (int i = 0; i < 2; ++i) s.c[i] = (float*)carve(rows * H * 4);
  float* qbuf = (float*)carve(rows * H * 4);
  s.q = qbuf; s.q_ld = H; s.q_rows = W;
  for (int i = 0; i < 2; ++i) s.seq[i] = (uint8_t*)carve(rows * 16);
  for (int i = 0; i < 2; ++i) s.cost[i] = (float*)carve(rows * 4);
  s.parent = (int32_t*)carve(rows * 4);
  const size_t smem = (size_t)4 * N * H * sizeof(float);
  MSQ_SMEM_ATTR(smem, dec_select_kernel<false>);
  int live = 1;
  for (int t = 0; t < N - 1; ++t) {
    const int cur = t & 1;   // tables (seq, cost) of step t live in slot cur; h'/c of step t are written to slot cur
    DecGemm g;
    g.Wt = w.whh_t; g.ldw = 4 * H; g.Ncols = 4 * H;
    g.h_src = t == 0 ? io.h0 : s.h[cur ^ 1]; g.c_src = t == 0 ? nullptr : s.c[cur ^ 1];
    g.parent = s.parent; g.seq = s.seq[cur]; g.xg = io.xg; g.bq = nullptr;
    g.h_out = s.h[cur]; g.c_out = s.c[cur];
    g.M = io.B * live; g.live = live; g.W = W; g.N = N; g.H = H; g.t = t;
    MSQ_TRY(launch_dec_gemm<0>(g, st));
    DecGemm q = g;
    q.Wt = w.wq_t; q.ldw = H; q.Ncols = H; q.h_src = s.h[cur]; q.c_src = nullptr; q.bq = w.bq; q.h_out = qbuf; q.c_out = nullptr;
    MSQ_TRY(launch_dec_gemm<1>(q, st));
    {
      // throughput form (every SM has CTAs to interleave): 256 threads, single staging buffer; latency form otherwise
      static int force = -2;
      if (force == -2) { const char* e = getenv("MSQ_DEC_MANY"); force = e ? atoi(e) : -1; }
      const bool many = force >= 0 ? force != 0 : io.B > 148;
      const int nbuf = many ? 1 : 2;
      MSQ_CUDA(launch_k(dec_select_kernel<false>, dim3((unsigned)io.B), dim3((!many && live > 8) ? 512 : 256), smem / (many ? 2 : 1), st, w, io, s, t, live, cur, nbuf));
    }
    MSQ_LAUNCH_CHECK();
    live = io.forced ? 1 : next_live(live, N, W);
  }
  return MSQ_OK;
}

// ---- tensor-core form: per step  cell (elementwise)  ->  [q | h' W_hh^T] = h' [W_q ; W_hh]^T + [b_q ; 0] as ONE tcgen05 GEMM on
// three-plane bf16 operands (bf16x6)  ->  dec_select.  h W_hh^T is computed once per PARENT row; the children gather it in
// the cell kernel together with their own XG row, so the beam re-gather costs nothing and no GEMM operand is permuted.
__global__ void __launch_bounds__(256) dec_cell_kernel(const float* __restrict__ xg, const float* __restrict__ hw_prev, int hw_ld,
                                                       const float* __restrict__ c_prev, const int32_t* __restrict__ parent,
                                                       const uint8_t* __restrict__ seq, int t, int live, int live_prev, int W, int N,
                                                       int H, int64_t rows, float* __restrict__ c_out, bf16* __restrict__ hp) {
  pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * H) return;
  const int64_t m = i / H;
  const int u = (int)(i % H);
  const int64_t b = m / live;
  const int wslot = (int)(m % live);
  const int prev = t == 0 ? N : seq[(b * W + wslot) * DC_MAXN + t - 1];
  const int p = t == 0 ? 0 : parent[b * W + wslot];
  const float4 x = *reinterpret_cast<const float4*>(xg + (b * (N + 1) + prev) * (int64_t)(4 * H) + 4 * u);
  const float4 g = *reinterpret_cast<const float4*>(hw_prev + (b * live_prev + p) * (int64_t)hw_ld + 4 * u);
  const float gi = x.x + g.x, gf = x.y + g.y, gc = x.z + g.z, go = x.w + g.w;
  const float ig = 1.f / (1.f + expf(-gi)), fg = 1.f / (1.f + expf(-gf));
  const float gg = tanhf(gc), og = 1.f / (1.f + expf(-go));
  const float cin = t == 0 ? 0.f : c_prev[(b * live_prev + p) * (int64_t)H + u];
  const float c2 = fg * cin + ig * gg;
  const float h2 = og * tanhf(c2);
  c_out[m * H + u] = c2;
  const bf16 p0 = __float2bfloat16_rn(h2);
  const float r1 = h2 - __bfloat162float(p0);
  const bf16 p1 = __float2bfloat16_rn(r1);
  bf16* d = hp + m * 3 * H + u;
  d[0] = p0; d[H] = p1; d[2 * H] = __float2bfloat16_rn(r1 - __bfloat162float(p1));
}

static int beam_search_tc(const DecodeWeights& w, const DecodeIO& io, cudaStream_t st) {
  const int H = io.H, N = io.N, W = io.W;
  const size_t rows = (size_t)io.B * W;
  char* p = reinterpret_cast<char*>(io.scratch);
  auto carve = [&](size_t bytes) { char* r = p; p += (bytes + 255) & ~size_t(255); return r; };
  bf16* hp = (bf16*)carve(rows * 3 * H * 2);
  float* cb[2]; float* qh[2];
  for (int i = 0; i < 2; ++i) cb[i] = (float*)carve(rows * H * 4);
  for (int i = 0; i < 2; ++i) qh[i] = (float*)carve(rows * 5 * H * 4);
  bf16* h0p = (bf16*)carve((size_t)io.B * 3 * H * 2);
  DecState s;
  s.h[0] = s.h[1] = nullptr; s.c[0] = cb[0]; s.c[1] = cb[1];
  for (int i = 0; i < 2; ++i) s.seq[i] = (uint8_t*)carve(rows * 16);
  for (int i = 0; i < 2; ++i) s.cost[i] = (float*)carve(rows * 4);
  s.parent = (int32_t*)carve(rows * 4);
  const size_t smem = (size_t)4 * N * H * sizeof(float);
  MSQ_SMEM_ATTR(smem, dec_select_kernel<true>);
  MSQ_SMEM_ATTR(smem, dec_select_kernel<false>);
  static int sfu = -1;   // MSQ_DEC_TANH=0: libm tanhf in the tensor-core form too
  if (sfu < 0) { const char* e = getenv("MSQ_DEC_TANH"); sfu = e ? (atoi(e) != 0) : 1; }
  GemmArgs g;
  g.bias = nullptr; g.resid = nullptr; g.C2 = nullptr; g.K = H; g.lda = H; g.ldw = H; g.ldc = 5 * H; g.ldr = 0; g.act = ACT_NONE; g.split = 2; g.trunc = io.trunc;
  // h0 W_hh^T -> columns [H, 5H) of qh[1] (read by the first cell as "previous step")
  MSQ_TRY(pack_split3(io.h0, io.B, H, H, H, h0p, st));
  g.A = h0p; g.W = w.wcat3 + (size_t)H * 3 * H; g.C = qh[1] + H; g.M = io.B; g.N = 4 * H;
  MSQ_TRY(gemm_tc<float>(g, st));
  int live = 1, live_prev = 1;
  for (int t = 0; t < N - 1; ++t) {
    const int cur = t & 1;
    const int64_t M = io.B * live;
    MSQ_CUDA(launch_k(dec_cell_kernel, dim3((unsigned)ceil_div(M * H, 256)), dim3(256), 0, st, io.xg, (const float*)(qh[cur ^ 1] + H), 5 * H,
                      (const float*)cb[cur ^ 1], (const int32_t*)s.parent, (const uint8_t*)s.seq[cur], t, live, live_prev, W, N, H, M, cb[cur], hp));
    MSQ_LAUNCH_CHECK();
    const bool last = t == N - 2;
    g.A = hp; g.W = w.wcat3; g.bias = w.bcat; g.C = qh[cur]; g.M = M; g.N = last ? H : 5 * H;   // the last step needs q only
    MSQ_TRY(gemm_tc<float>(g, st));
    s.q = qh[cur]; s.q_ld = 5 * H; s.q_rows = live;
    {
      // throughput form (every SM has CTAs to interleave): 256 threads, single staging buffer; latency form otherwise
      static int force = -2;
      if (force == -2) { const char* e = getenv("MSQ_DEC_MANY"); force = e ? atoi(e) : -1; }
      const bool many = force >= 0 ? force != 0 : io.B > 148;
      const int nbuf = many ? 1 : 2;
      MSQ_CUDA(launch_k(sfu ? dec_select_kernel<true> : dec_select_kernel<false>, dim3((unsigned)io.B), dim3((!many && live > 8) ? 512 : 256),
                        smem / (many ? 2 : 1), st, w, io, s, t, live, cur, nbuf));
    }
    MSQ_LAUNCH_CHECK();
    live_prev = live;
    live = io.forced ? 1 : next_live(live, N, W);
  }
  return MSQ_OK;
}

int beam_search(const DecodeWeights& w, const DecodeIO& io, cudaStream_t st) {
  MSQ_REQUIRE(io.N >= 2 && io.N <= DC_MAXN, "beam_search: N=%d out of range [2,%d]", io.N, DC_MAXN);
  MSQ_REQUIRE(io.W >= 1 && io.W <= 16, "beam_search: beam width %d out of range [1,16]", io.W);
  MSQ_REQUIRE(io.H % 16 == 0 && io.H <= 1024, "beam_search: H=%d unsupported", io.H);
  if (io.B == 0) return MSQ_OK;
  static int fused = -1;
  if (fused < 0) { const char* e = getenv("MSQ_DECODE_FUSED"); fused = (e && e[0] == '1') ? 1 : 0; }
  if (fused || !io.scratch) return beam_search_fused(w, io, st);
  if (io.tc && w.wcat3 && io.H % 64 == 0) return beam_search_tc(w, io, st);
  return beam_search_tiled(w, io, st);
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
  MSQ_SMEM_ATTR(smem, pointer_p1_kernel);
  MSQ_CUDA(launch_k(pointer_p1_kernel, dim3((unsigned)B), dim3(256), smem, st, enc, cls, y, W1, W2, V, Wih, Whh, bih, bhh, N, H, U, preds,
                    ce_scratch));
  MSQ_LAUNCH_CHECK();
  MSQ_CUDA(launch_k(pointer_p1_loss_kernel, dim3(1), dim3(32), 0, st, (const float*)ce_scratch, B, loss));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

}  // namespace msq
