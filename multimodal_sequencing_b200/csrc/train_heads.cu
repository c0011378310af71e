// Fine-tuning of the BERSON heads: training forward of everything behind the inner encoder, the loss, and its backward
// pass down to dL/d(lang) -- the tensor msq_inner_backward (train.cu) continues from.
//
// Reference (models/berson/modeling_bert.py): HierarchicalAttention.forward 666-817 (token pooling, pair heads, edge
// pooling), encode 1239-1366 (paragraph encoder models/berson/encoder.py:9-61 + neural.py:11-232, h0, key_linear),
// _forward 943-1174 (teacher-forced pointer decoder: LSTM 1086-1092, query 1094, pw_k keys 1027-1078, tanh scoring
// 1098-1101, masks 1112-1113, NLL / (N-1) 1126-1142, + lam * pairwise NLL / P 1145-1174).  Autograd there; explicit
// backward kernels here.  The heads are a few MFLOP per manual: all of it runs in fp32 (the sentence_tran GEMM in the
// encoder's operand type) with one block per manual / pair / step and NO atomics: every reduction over the batch is a
// two-stage ordered sum, so gradients are bit-reproducible.
//
// Teacher-forced decoder in the base-tensor formulation: with T4 = R0 W_pw^T = [A1|A2|F|G] per manual,
//   keys[t,k] = A1[last_t,k] + A2[last2_t,k] + (sum_{j in Rem_t, j!=k} F[k,j] + sum_{i in Rem_t, i!=k} G[i,k]) / N
// is linear in T4 with a selection pattern fixed by the target order, so d(keys) scatters straight back into dT4 and
// W_pw's gradient is ONE GEMM (dT4^T R0) instead of N per-step ones.
#include "train_common.cuh"

namespace msq {

constexpr int TH_MAXN = 16;

__device__ __forceinline__ int th_pair_index(int i, int j, int N) {   // pairs_generator order (process_inputs_for_berson.py:246-261)
  const int a = i < j ? i : j, b = i < j ? j : i;
  const int c = a * N - a * (a + 1) / 2 + (b - a - 1);
  return i < j ? c : c + N * (N - 1) / 2;
}

// out[c] += sum_b partial[b * stride + col0 + c], c < width   (ordered)
__global__ void __launch_bounds__(256) partial_reduce_kernel(const float* __restrict__ partial, int nblk, int stride, int col0, int width,
                                                             float* __restrict__ out) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= width) return;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += partial[(int64_t)b * stride + col0 + c];
  out[c] += s;
}
static int partial_reduce(const float* partial, int nblk, int stride, int col0, int width, float* out, cudaStream_t st) {
  MSQ_CUDA(launch_k(partial_reduce_kernel, dim3(ceil_div(width, 256)), dim3(256), 0, st, partial, nblk, stride, col0, width, out));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
// dst[b][a][:] = src[a][b][:]
__global__ void __launch_bounds__(256) transpose_rows_kernel(const float* __restrict__ src, int A, int Bn, int H, float* __restrict__ dst) {
  pdl_sync();
  const int64_t total = (int64_t)A * Bn * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % H);
    const int64_t row = i / H;
    const int a = (int)(row / Bn), b = (int)(row % Bn);
    dst[((int64_t)b * A + a) * H + d] = src[i];
  }
}
static int transpose_rows(const float* src, int64_t A, int64_t Bn, int H, float* dst, cudaStream_t st) {
  if (A * Bn == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(transpose_rows_kernel, dim3((unsigned)min((int64_t)148 * 4, (A * Bn * H + 255) / 256)), dim3(256), 0, st, src, (int)A, (int)Bn, H, dst));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
// y[i] += a * x[i]
__global__ void __launch_bounds__(256) axpy_kernel(const float* __restrict__ x, float a, int64_t n, float* __restrict__ y) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = fmaf(a, x[i], y[i]);
}
static int axpy(const float* x, float a, int64_t n, float* y, cudaStream_t st) {
  if (n == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(axpy_kernel, dim3((unsigned)min((int64_t)148 * 4, (n + 255) / 256)), dim3(256), 0, st, x, a, n, y));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// ---------------------------------------------------------------------------------------------------
// token score backward (modeling_bert.py:697-701): score = w2 . tt + b2, tt = tanh(Ws x + bs)
//   dpre = dscore w2 (1 - tt^2);  dw2 += sum dscore tt;  db2 += sum dscore    (partials per block, [H + 1])
// ---------------------------------------------------------------------------------------------------
constexpr int SB_WARPS = 4;
template <typename T>
__global__ void __launch_bounds__(SB_WARPS * 32) score_bwd_kernel(const float* __restrict__ dscore, const T* __restrict__ tt, int64_t rows, int H,
                                                                  const float* __restrict__ w2, T* __restrict__ dpre, float* __restrict__ partial) {
  pdl_sync();
  __shared__ float sh[SB_WARPS][1024 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nc = H >> 5;
  float acc[32], accb = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  for (int64_t row = (int64_t)blockIdx.x * SB_WARPS + warp; row < rows; row += (int64_t)gridDim.x * SB_WARPS) {
    const float ds = dscore[row];
    accb += ds;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < nc) {
        const int c = i * 32 + lane;
        const float th = to_f(tt[row * H + c]);
        acc[i] = fmaf(ds, th, acc[i]);
        dpre[row * H + c] = from_f<T>(ds * w2[c] * (1.0f - th * th));
      }
  }
#pragma unroll
  for (int i = 0; i < 32; ++i)
    if (i < nc) sh[warp][i * 32 + lane] = acc[i];
  if (lane == 0) sh[warp][H] = accb;
  __syncthreads();
  for (int c = threadIdx.x; c <= H; c += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < SB_WARPS; ++w) s += sh[w][c];
    partial[(int64_t)blockIdx.x * (H + 1) + c] = s;
  }
}

// ---------------------------------------------------------------------------------------------------
// token pooling backward (modeling_bert.py:705-738), one block per pair row: recompute the two span softmaxes, then
//   d_lang[r,t,:] = p_s[t] dmix[r,s,:]   (s = span of t; zero rows elsewhere);  dscore[t] = p_s[t] (dp[t] - sum_t' p_s[t'] dp[t'])
// with dp[t] = dmix[r,s] . x[r,t].
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) token_pool_bwd_kernel(const T* __restrict__ tt, const float* __restrict__ x, int Lt, int Lj, int H,
                                                             const float* __restrict__ w2, const float* __restrict__ b2,
                                                             const int64_t* __restrict__ sep, const float* __restrict__ dmix,
                                                             float* __restrict__ d_lang, float* __restrict__ dscore, Drop drop) {
  pdl_sync();
  extern __shared__ float sm[];
  float* score = sm;          // [Lt]
  float* p = sm + Lt;         // [Lt] probability of t inside its own span
  float* dp = p + Lt;         // [Lt]
  __shared__ float dots[2];
  const int64_t r = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int sep0 = (int)sep[r * 2], sep1 = (int)sep[r * 2 + 1];
  const float* xr = x + r * Lj * (int64_t)H;
  for (int t = warp; t < Lt; t += nw) {
    const T* row = tt + (r * Lt + t) * H;
    float a = 0.f;
    for (int d = lane; d < H; d += 32) a = fmaf(to_f(row[d]), w2[d], a);
    a = warp_sum(a);
    if (lane == 0) score[t] = a + b2[0];
    // dp[t]
    float c = 0.f;
    if (t >= 1 && t <= sep1) {
      const float* dm = dmix + (r * 2 + (t <= sep0 ? 0 : 1)) * H;
      for (int d = lane; d < H; d += 32) c = fmaf(dm[d], xr[(int64_t)t * H + d], c);
      c = warp_sum(c);
      c *= drop_mul(drop, ((uint64_t)r * 2 + (t <= sep0 ? 0 : 1)) * Lt + t);   // dP = (dmix . x_t) . mask
    }
    if (lane == 0) dp[t] = c;
  }
  __syncthreads();
  if (warp < 2) {
    const int lo = warp == 0 ? 1 : sep0 + 1, hi = warp == 0 ? sep0 : sep1;
    float mx = -INFINITY;
    for (int t = lo + lane; t <= hi; t += 32) mx = fmaxf(mx, score[t]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int t = lo + lane; t <= hi; t += 32) { const float e = expf(score[t] - mx); p[t] = e; s += e; }
    s = warp_sum(s);
    const float inv = 1.f / s;
    float dot = 0.f;
    for (int t = lo + lane; t <= hi; t += 32) { p[t] *= inv; dot = fmaf(p[t], dp[t], dot); }
    dot = warp_sum(dot);
    if (lane == 0) dots[warp] = dot;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < Lt; t += blockDim.x) {
    float ds = 0.f;
    if (t >= 1 && t <= sep1) ds = p[t] * (dp[t] - dots[t <= sep0 ? 0 : 1]);
    dscore[r * Lt + t] = ds;
  }
  float* dl = d_lang + r * Lt * (int64_t)H;
  for (int i = threadIdx.x; i < Lt * H; i += blockDim.x) {
    const int t = i / H, d = i % H;
    float v = 0.f;
    if (t >= 1 && t <= sep1)
      v = p[t] * drop_mul(drop, ((uint64_t)r * 2 + (t <= sep0 ? 0 : 1)) * Lt + t) * dmix[(r * 2 + (t <= sep0 ? 0 : 1)) * H + d];
    dl[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// pair heads backward, one block per pair row r = (b, p), p = pair (i, j):
//   R0[b,i,j] = [cls ; softmax(z)] (rela_encode 919-925), pairwise loss lam/(P B) * NLL(softmax(z), label) (1145-1174)
//   dz_c = s_c (dR0[H+c] - sum s dR0[H+.]) + lam_scale (s_c - [c == label]);   d_lang[r,0,:] += dR0[:H] + sum_c dz_c w_rel[c,:]
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cls_rel_bwd_kernel(const float* __restrict__ dr0, int Kp, const float* __restrict__ rel6,
                                                          const int64_t* __restrict__ labels, const float* __restrict__ w_rel, int N, int Lt,
                                                          int H, float lam_scale, float* __restrict__ d_lang, float* __restrict__ drel2) {
  pdl_sync();
  __shared__ int cell_sh;
  __shared__ float dz_sh[2];
  const int64_t r = blockIdx.x;
  const int P = N * (N - 1);
  const int b = (int)(r / P), p = (int)(r % P);
  if (threadIdx.x == 0) {
    int ci = 0, cj = 1;
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < N; ++j)
        if (i != j && th_pair_index(i, j, N) == p) { ci = i; cj = j; }
    const int cell = (b * N + ci) * N + cj;
    cell_sh = cell;
    const float z0 = rel6[r * 6], z1 = rel6[r * 6 + 1];
    const float mx = fmaxf(z0, z1), e0 = expf(z0 - mx), e1 = expf(z1 - mx);
    const float s0 = e0 / (e0 + e1), s1 = e1 / (e0 + e1);
    const float g0 = dr0[(int64_t)cell * Kp + H], g1 = dr0[(int64_t)cell * Kp + H + 1];
    const float dot = s0 * g0 + s1 * g1;
    const int lab = (int)labels[r];
    const float d0 = s0 * (g0 - dot) + lam_scale * (s0 - (lab == 0 ? 1.f : 0.f));
    const float d1 = s1 * (g1 - dot) + lam_scale * (s1 - (lab == 1 ? 1.f : 0.f));
    dz_sh[0] = d0; dz_sh[1] = d1;
    drel2[r * 2] = d0; drel2[r * 2 + 1] = d1;
  }
  __syncthreads();
  const float d0 = dz_sh[0], d1 = dz_sh[1];
  const float* g = dr0 + (int64_t)cell_sh * Kp;
  float* dl = d_lang + r * Lt * (int64_t)H;   // token 0 = the pair's [CLS]
  for (int d = threadIdx.x; d < H; d += blockDim.x) dl[d] += g[d] + d0 * w_rel[d] + d1 * w_rel[H + d];
}
// dW_rel[c,:] += sum_r drel2[r,c] cls[r,:];  db_rel[c] += sum_r drel2[r,c]   (thread per (c, h), ordered over r)
__global__ void __launch_bounds__(256) rel_wgrad_kernel(const float* __restrict__ drel2, const float* __restrict__ x, int64_t R, int Lj, int H,
                                                        float* __restrict__ dW, float* __restrict__ db) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * H) return;
  const int c = i / H, h = i % H;
  float s = 0.f, sb = 0.f;
  for (int64_t r = 0; r < R; ++r) {
    const float g = drel2[r * 2 + c];
    s = fmaf(g, x[r * Lj * (int64_t)H + h], s);
    sb += g;
  }
  dW[i] += s;
  if (h == 0) db[c] += sb;
}

// ---------------------------------------------------------------------------------------------------
// edge pooling backward (modeling_bert.py:796-814), one block per (manual, step): recompute the softmax over the step's
// 2(N-1) edge vectors, then  dv_e = w_e dsents + da_e w_in2,  da_e = w_e (dsents . v_e - sum_e' w_e' dsents . v_e'),
// dw_in2 partial[block] = sum_e da_e v_e.  Every mix slot belongs to exactly one (step, edge): plain stores.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) edge_pool_bwd_kernel(const float* __restrict__ mix, const float* __restrict__ dsents, int N, int H,
                                                            const float* __restrict__ w_in2, float* __restrict__ dmix, float* __restrict__ partial) {
  pdl_sync();
  __shared__ int e_pair[2 * TH_MAXN], e_side[2 * TH_MAXN];
  __shared__ float e_w[2 * TH_MAXN], e_d[2 * TH_MAXN];
  const int b = blockIdx.x / N, s = blockIdx.x % N;
  const int P = N * (N - 1), E = 2 * (N - 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    int n = 0;
    for (int half = 0; half < 2; ++half)
      for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) {
          const int first = half ? j : i, second = half ? i : j;
          const int p = th_pair_index(first, second, N);
          if (first == s) { e_pair[n] = p; e_side[n] = 0; ++n; }
          else if (second == s) { e_pair[n] = p; e_side[n] = 1; ++n; }
        }
  }
  __syncthreads();
  const float* mb = mix + (int64_t)b * P * 2 * H;
  float* dmb = dmix + (int64_t)b * P * 2 * H;
  const float* ds = dsents + ((int64_t)b * N + s) * H;
  for (int e = warp; e < E; e += nw) {
    const float* v = mb + ((int64_t)e_pair[e] * 2 + e_side[e]) * H;
    float a = 0.f, c = 0.f;
    for (int d = lane; d < H; d += 32) { a = fmaf(v[d], w_in2[d], a); c = fmaf(v[d], ds[d], c); }
    a = warp_sum(a); c = warp_sum(c);
    if (lane == 0) { e_w[e] = a; e_d[e] = c; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = -INFINITY, sum = 0.f, dot = 0.f;
    for (int e = 0; e < E; ++e) mx = fmaxf(mx, e_w[e]);
    for (int e = 0; e < E; ++e) { e_w[e] = expf(e_w[e] - mx); sum += e_w[e]; }
    for (int e = 0; e < E; ++e) { e_w[e] /= sum; dot = fmaf(e_w[e], e_d[e], dot); }
    for (int e = 0; e < E; ++e) e_d[e] = e_w[e] * (e_d[e] - dot);   // da_e
  }
  __syncthreads();
  for (int d = threadIdx.x; d < H; d += blockDim.x) {
    float acc = 0.f;
    const float dsd = ds[d], wi = w_in2[d];
    for (int e = 0; e < E; ++e) {
      const int64_t o = ((int64_t)e_pair[e] * 2 + e_side[e]) * H + d;
      dmb[o] = e_w[e] * dsd + e_d[e] * wi;
      acc = fmaf(e_d[e], mb[o], acc);
    }
    partial[(int64_t)blockIdx.x * H + d] = acc;
  }
}

// ---------------------------------------------------------------------------------------------------
// paragraph attention backward (neural.py:200-226), one block per (manual, head); N <= 16 tokens.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) para_attention_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ dctx, int N, int heads,
                                                                 int H, float* __restrict__ dqkv, Drop drop) {
  pdl_sync();
  __shared__ float s[TH_MAXN][TH_MAXN + 1], g[TH_MAXN][TH_MAXN + 1];
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int d = H / heads;
  const float inv = 1.0f / sqrtf((float)d);
  const float* base = qkv + (int64_t)b * N * 3 * H + h * d;
  const float* dc = dctx + (int64_t)b * N * H + h * d;
  float* dbase = dqkv + (int64_t)b * N * 3 * H + h * d;
  for (int ij = threadIdx.x; ij < N * N; ij += blockDim.x) {
    const int i = ij / N, j = ij % N;
    const float* q = base + (int64_t)i * 3 * H;
    const float* k = base + (int64_t)j * 3 * H + H;
    const float* v = base + (int64_t)j * 3 * H + 2 * H;
    const float* go = dc + (int64_t)i * H;
    float a = 0.f, c = 0.f;
    for (int e = 0; e < d; ++e) { a = fmaf(q[e] * inv, k[e], a); c = fmaf(go[e], v[e], c); }
    s[i][j] = a;
    g[i][j] = c;   // dP
  }
  __syncthreads();
  if (threadIdx.x < N) {
    const int i = threadIdx.x;
    float mx = -INFINITY, sum = 0.f, dot = 0.f;
    for (int j = 0; j < N; ++j) mx = fmaxf(mx, s[i][j]);
    for (int j = 0; j < N; ++j) { s[i][j] = expf(s[i][j] - mx); sum += s[i][j]; }
    for (int j = 0; j < N; ++j) {
      const float dm = drop_mul(drop, ((uint64_t)blockIdx.x * N + i) * N + j);
      s[i][j] /= sum;
      g[i][j] *= dm;                       // dP = (dctx v^T) . mask
      dot = fmaf(s[i][j], g[i][j], dot);
    }
    for (int j = 0; j < N; ++j) {
      g[i][j] = s[i][j] * (g[i][j] - dot);   // dS
      s[i][j] *= drop_mul(drop, ((uint64_t)blockIdx.x * N + i) * N + j);   // P . mask, what dV sees
    }
  }
  __syncthreads();
  for (int ie = threadIdx.x; ie < N * d; ie += blockDim.x) {
    const int i = ie / d, e = ie % d;
    float dq = 0.f, dk = 0.f, dv = 0.f;
    for (int j = 0; j < N; ++j) {
      dq = fmaf(g[i][j], base[(int64_t)j * 3 * H + H + e], dq);        // dS[i,j] k[j]
      dk = fmaf(g[j][i], base[(int64_t)j * 3 * H + e], dk);            // dS[j,i] q[j]
      dv = fmaf(s[j][i], dc[(int64_t)j * H + e], dv);                  // P[j,i] dctx[j]
    }
    dbase[(int64_t)i * 3 * H + e] = dq * inv;
    dbase[(int64_t)i * 3 * H + H + e] = dk * inv;
    dbase[(int64_t)i * 3 * H + 2 * H + e] = dv;
  }
}

// dsents[b,s] += dkeyin[b,s,:H];  dpara[b,s] = dkeyin[b,s,H:] + dh0[b] / (N + 1e-20)     (modeling_bert.py:1346-1357)
__global__ void __launch_bounds__(256) para_finish_bwd_kernel(const float* __restrict__ dkeyin, const float* __restrict__ dh0, int N, int H,
                                                              float* __restrict__ dsents, float* __restrict__ dpara) {
  pdl_sync();
  const int b = blockIdx.x;
  const float den = (float)N + 1e-20f;
  for (int d = threadIdx.x; d < H; d += blockDim.x) {
    const float g0 = dh0[(int64_t)b * H + d] / den;
    for (int s = 0; s < N; ++s) {
      const int64_t row = (int64_t)b * N + s;
      dsents[row * H + d] += dkeyin[row * 2 * H + d];
      dpara[row * H + d] = dkeyin[row * 2 * H + H + d] + g0;
    }
  }
}
// sents_ext[b, n] = sents[b, n] (n < N), 0 (n == N)  /  dsents[b, n] += dsents_ext[b, n]
__global__ void __launch_bounds__(256) sents_ext_train_kernel(const float* __restrict__ sents, int64_t B, int N, int H, float* __restrict__ out) {
  pdl_sync();
  const int64_t total = B * (N + 1) * (int64_t)H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % H);
    const int64_t row = i / H, b = row / (N + 1);
    const int n = (int)(row % (N + 1));
    out[i] = n < N ? sents[(b * N + n) * H + d] : 0.f;
  }
}
__global__ void __launch_bounds__(256) sents_ext_bwd_kernel(const float* __restrict__ dext, int64_t B, int N, int H, float* __restrict__ dsents) {
  pdl_sync();
  const int64_t total = B * N * (int64_t)H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % H);
    const int64_t row = i / H, b = row / N;
    const int n = (int)(row % N);
    dsents[i] += dext[(b * (N + 1) + n) * H + d];
  }
}

// ---------------------------------------------------------------------------------------------------
// LSTM cell, teacher-forced (nn.LSTM, gate order i,f,g,o; modeling_bert.py:886, 1086-1092).
//   forward : pre = h_prev Whh^T + bhh (GEMM) + xg[b, in_t]  -> act = (sig i, sig f, tanh g, sig o), c, h
//   backward: dgates (pre-activation) from dh, dc_next; dc_prev; scatter of dgates into dxg[b, in_t]
// in_t(b) = N (the zero input row) at t = 0, target[b, t-1] afterwards.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lstm_cell_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ xg, const int32_t* __restrict__ target,
                                                            int t, int N, int H, const float* __restrict__ c_prev, float* __restrict__ act,
                                                            float* __restrict__ c_out, float* __restrict__ h_out) {
  pdl_sync();
  const int b = blockIdx.x;
  const int in = t == 0 ? N : target[b * N + t - 1];
  const float* x = xg + ((int64_t)b * (N + 1) + in) * 4 * H;
  const float* p = pre + (int64_t)b * 4 * H;
  for (int u = threadIdx.x; u < H; u += blockDim.x) {
    const float gi = 1.f / (1.f + expf(-(p[u] + x[u])));
    const float gf = 1.f / (1.f + expf(-(p[H + u] + x[H + u])));
    const float gg = tanhf(p[2 * H + u] + x[2 * H + u]);
    const float go = 1.f / (1.f + expf(-(p[3 * H + u] + x[3 * H + u])));
    const float cp = c_prev ? c_prev[(int64_t)b * H + u] : 0.f;
    const float c = gf * cp + gi * gg;
    float* a = act + (int64_t)b * 4 * H;
    a[u] = gi; a[H + u] = gf; a[2 * H + u] = gg; a[3 * H + u] = go;
    c_out[(int64_t)b * H + u] = c;
    h_out[(int64_t)b * H + u] = go * tanhf(c);
  }
}
__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(const float* __restrict__ dh_a, const float* __restrict__ dh_b, const float* __restrict__ act,
                                                            const float* __restrict__ c, const float* __restrict__ c_prev, float* __restrict__ dc,
                                                            const int32_t* __restrict__ target, int t, int N, int H, float* __restrict__ dgates,
                                                            float* __restrict__ dxg) {
  pdl_sync();
  const int b = blockIdx.x;
  const int in = t == 0 ? N : target[b * N + t - 1];
  const float* a = act + (int64_t)b * 4 * H;
  float* dg = dgates + (int64_t)b * 4 * H;
  float* dx = dxg + ((int64_t)b * (N + 1) + in) * 4 * H;
  for (int u = threadIdx.x; u < H; u += blockDim.x) {
    const int64_t o = (int64_t)b * H + u;
    const float dh = dh_a[o] + (dh_b ? dh_b[o] : 0.f);
    const float gi = a[u], gf = a[H + u], gg = a[2 * H + u], go = a[3 * H + u];
    const float tc = tanhf(c[o]);
    const float dct = dc[o] + dh * go * (1.f - tc * tc);
    const float cp = c_prev ? c_prev[o] : 0.f;
    const float di = dct * gg * gi * (1.f - gi), df = dct * cp * gf * (1.f - gf), dgg = dct * gi * (1.f - gg * gg),
                dgo = dh * tc * go * (1.f - go);
    dg[u] = di; dg[H + u] = df; dg[2 * H + u] = dgg; dg[3 * H + u] = dgo;
    dx[u] += di; dx[H + u] += df; dx[2 * H + u] += dgg; dx[3 * H + u] += dgo;
    dc[o] = dct * gf;
  }
}

// ---------------------------------------------------------------------------------------------------
// Teacher-forced pointer scoring, loss and its gradient, one block per manual (all N positions in order):
//   e[t,k] = wt . tanh(q[t] + keys[t,k] + key0[k]) + bt;  picked steps masked (-1e9);  nll += -log_softmax(e[t])[target[t]]
//   de[t,k] = (softmax_k - [k == target[t]]) * scale,  scale = 1 / ((N - 1) B)        (modeling_bert.py:1098-1142)
// and, in the same pass, dq[t], dkey0[k] (+= over t), dT4 (+= through the selection pattern of `keys`), per-block
// partials of dwt / dbt.  dT4 / dkey0 rows of a manual are touched by its own block only, in program order: no atomics.
// Dynamic shared memory: th[N][H] (tanh values, overwritten by dz).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tf_pointer_kernel(const float* __restrict__ query, const float* __restrict__ t4, const float* __restrict__ key0,
                                                         const int32_t* __restrict__ target, const float* __restrict__ wt, float bt, int N, int H,
                                                         float scale, float* __restrict__ nll, float* __restrict__ dquery, float* __restrict__ dkey0,
                                                         float* __restrict__ dt4, float* __restrict__ partial) {
  pdl_sync();
  extern __shared__ float th[];   // [N][H]
  __shared__ float red[8], e_sh[TH_MAXN], de_sh[TH_MAXN];
  __shared__ int picked[TH_MAXN];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t* tg = target + b * N;
  const float* T = t4 + (int64_t)b * N * N * 4 * H;
  float* dT = dt4 + (int64_t)b * N * N * 4 * H;
  const float invN = 1.0f / (float)N;
  float dwt_acc[4] = {0.f, 0.f, 0.f, 0.f}, dbt_acc = 0.f, nll_acc = 0.f;
  if (tid < TH_MAXN) picked[tid] = 0;
  __syncthreads();
  for (int t = 0; t < N; ++t) {
    const int last = t >= 1 ? tg[t - 1] : -1, last2 = t >= 2 ? tg[t - 2] : -1;
    if (tid == 0 && last >= 0) picked[last] = 1;
    __syncthreads();
    const float* q = query + ((int64_t)b * N + t) * H;
    for (int k = 0; k < N; ++k) {
      if (picked[k]) { if (tid == 0) e_sh[k] = -1e9f; continue; }
      float part = 0.f;
      for (int h = tid; h < H; h += 256) {
        float z = q[h] + key0[((int64_t)b * N + k) * H + h];
        if (last >= 0) z += T[((int64_t)(last * N + k)) * 4 * H + h];
        if (last2 >= 0) z += T[((int64_t)(last2 * N + k)) * 4 * H + H + h];
        float fg = 0.f;
        for (int j = 0; j < N; ++j)
          if (!picked[j] && j != k) fg += T[((int64_t)(k * N + j)) * 4 * H + 2 * H + h] + T[((int64_t)(j * N + k)) * 4 * H + 3 * H + h];
        z = fmaf(fg, invN, z);
        const float tv = tanhf(z);
        th[k * H + h] = tv;
        part = fmaf(wt[h], tv, part);
      }
      part = warp_sum(part);
      if (lane == 0) red[warp] = part;
      __syncthreads();
      if (tid == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        e_sh[k] = s + bt;
      }
      __syncthreads();
    }
    __syncthreads();
    if (tid == 0) {
      float mx = -INFINITY, sum = 0.f;
      for (int k = 0; k < N; ++k) mx = fmaxf(mx, e_sh[k]);
      for (int k = 0; k < N; ++k) sum += expf(e_sh[k] - mx);
      const float lse = mx + logf(sum);
      nll_acc -= e_sh[tg[t]] - lse;
      for (int k = 0; k < N; ++k) {
        const float d = picked[k] ? 0.f : (expf(e_sh[k] - lse) - (k == tg[t] ? 1.f : 0.f)) * scale;
        de_sh[k] = d;
        dbt_acc += d;
      }
    }
    __syncthreads();
    int hi = 0;
    for (int h = tid; h < H; h += 256, ++hi) {
      float dq = 0.f;
      const float w = wt[h];
      for (int k = 0; k < N; ++k) {
        if (picked[k]) continue;
        const float tv = th[k * H + h], de = de_sh[k];
        dwt_acc[hi] = fmaf(de, tv, dwt_acc[hi]);
        const float dz = de * w * (1.f - tv * tv);
        dq += dz;
        dkey0[((int64_t)b * N + k) * H + h] += dz;
        if (last >= 0) dT[((int64_t)(last * N + k)) * 4 * H + h] += dz;
        if (last2 >= 0) dT[((int64_t)(last2 * N + k)) * 4 * H + H + h] += dz;
        const float dzn = dz * invN;
        for (int j = 0; j < N; ++j)
          if (!picked[j] && j != k) {
            dT[((int64_t)(k * N + j)) * 4 * H + 2 * H + h] += dzn;
            dT[((int64_t)(j * N + k)) * 4 * H + 3 * H + h] += dzn;
          }
      }
      dquery[((int64_t)b * N + t) * H + h] = dq;
    }
    __syncthreads();
  }
  int hi = 0;
  for (int h = tid; h < H; h += 256, ++hi) partial[(int64_t)b * (H + 1) + h] = dwt_acc[hi];
  if (tid == 0) { partial[(int64_t)b * (H + 1) + H] = dbt_acc; nll[b] = nll_acc; }
}

// ---- time-contrastive objective (modeling_bert.py:1176-1216): weight * mean_b max(|a - p + eps| - |a - n + eps| + 1, 0) over the
// sentence vectors a, p, n = sents[b, triplet[b]] (nn.TripletMarginLoss(margin=1, p=2): pairwise_distance adds eps = 1e-6 to the
// difference).  One block per manual: loss term into tl[b], gradient rows added to dsents (the three rows are distinct).
__global__ void __launch_bounds__(256) triplet_kernel(const float* __restrict__ sents, const int32_t* __restrict__ trip, int N, int H, float gscale,
                                                      float* __restrict__ tl, float* __restrict__ dsents) {
  pdl_sync();
  __shared__ float red[2][8];
  const int64_t b = blockIdx.x;
  const int ia = trip[b * 3], ip = trip[b * 3 + 1], in = trip[b * 3 + 2];
  const float *a = sents + (b * N + ia) * H, *p = sents + (b * N + ip) * H, *n = sents + (b * N + in) * H;
  float sp = 0.f, sn = 0.f;
  for (int d = threadIdx.x; d < H; d += blockDim.x) {
    const float x = a[d] - p[d] + 1e-6f, y = a[d] - n[d] + 1e-6f;
    sp = fmaf(x, x, sp); sn = fmaf(y, y, sn);
  }
  sp = warp_sum(sp); sn = warp_sum(sn);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sp; red[1][threadIdx.x >> 5] = sn; }
  __syncthreads();
  float tp = 0.f, tn = 0.f;
  for (int i = 0; i < 8; ++i) { tp += red[0][i]; tn += red[1][i]; }
  const float dap = sqrtf(tp), dan = sqrtf(tn), l = dap - dan + 1.0f;
  if (threadIdx.x == 0) tl[b] = fmaxf(l, 0.f);
  if (!(l > 0.f) || !dsents) return;
  const float cp = dap > 0.f ? gscale / dap : 0.f, cn = dan > 0.f ? gscale / dan : 0.f;
  float *ga = dsents + (b * N + ia) * H, *gp = dsents + (b * N + ip) * H, *gn = dsents + (b * N + in) * H;
  for (int d = threadIdx.x; d < H; d += blockDim.x) {
    const float x = (a[d] - p[d] + 1e-6f) * cp, y = (a[d] - n[d] + 1e-6f) * cn;
    ga[d] += x - y; gp[d] -= x; gn[d] += y;
  }
}
__global__ void triplet_loss_add_kernel(const float* __restrict__ tl, int64_t B, float w, float* __restrict__ loss) {
  pdl_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int64_t b = 0; b < B; ++b) s += tl[b];
    *loss += w * s / (float)B;
  }
}

// ---- image pairwise objective (args.multimodal_loss; modeling_bert.py:1359-1364 + 1218-1225): per pair row r
//   v = visn[r, 0] (the first visual token of the final joint stream), u = img_projection(v), z = pairwise_relationship(u),
//   loss += lam / (P B) * NLL(softmax(z), label_r).  One block per pair row: forward, d(z), d(u), d(v).
__global__ void __launch_bounds__(256) img_pair_kernel(const float* __restrict__ x, int Lt, int Lj, int H, const float* __restrict__ Wp,
                                                       const float* __restrict__ bp, const float* __restrict__ w_rel, const float* __restrict__ b_rel,
                                                       const int64_t* __restrict__ labels, float scale, float* __restrict__ u_out,
                                                       float* __restrict__ du_out, float* __restrict__ dz_out, float* __restrict__ dv_out,
                                                       float* __restrict__ loss_rows) {
  pdl_sync();
  __shared__ float v[1024], u[1024], du[1024];
  __shared__ float z[2], dz[2];
  const int64_t r = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* vr = x + (r * Lj + Lt) * (int64_t)H;
  for (int d = threadIdx.x; d < H; d += blockDim.x) v[d] = vr[d];
  __syncthreads();
  for (int i = warp; i < H; i += nw) {
    const float* w = Wp + (int64_t)i * H;
    float a = 0.f;
    for (int d = lane; d < H; d += 32) a = fmaf(w[d], v[d], a);
    a = warp_sum(a);
    if (lane == 0) u[i] = a + bp[i];
  }
  __syncthreads();
  if (warp < 2) {
    float a = 0.f;
    for (int d = lane; d < H; d += 32) a = fmaf(w_rel[warp * H + d], u[d], a);
    a = warp_sum(a);
    if (lane == 0) z[warp] = a + b_rel[warp];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float mx = fmaxf(z[0], z[1]), e0 = expf(z[0] - mx), e1 = expf(z[1] - mx), lse = mx + logf(e0 + e1);
    const int lab = (int)labels[r];
    loss_rows[r] = lse - z[lab];
    dz[0] = scale * (e0 / (e0 + e1) - (lab == 0 ? 1.f : 0.f));
    dz[1] = scale * (e1 / (e0 + e1) - (lab == 1 ? 1.f : 0.f));
    dz_out[r * 2] = dz[0]; dz_out[r * 2 + 1] = dz[1];
  }
  __syncthreads();
  for (int d = threadIdx.x; d < H; d += blockDim.x) {
    const float g = dz[0] * w_rel[d] + dz[1] * w_rel[H + d];
    du[d] = g;
    u_out[r * H + d] = u[d];
    du_out[r * H + d] = g;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < H; k += blockDim.x) {   // dv = W^T du: consecutive threads read consecutive columns of a row
    float a = 0.f;
    for (int i = 0; i < H; ++i) a = fmaf(du[i], Wp[(int64_t)i * H + k], a);
    dv_out[r * H + k] = a;
  }
}
// dW[i, k] += sum_r du[r, i] v[r, k] (v = x[r, Lt, :]);  db[i] += sum_r du[r, i]   (thread per (i, k), ordered over r)
__global__ void __launch_bounds__(256) img_proj_wgrad_kernel(const float* __restrict__ du, const float* __restrict__ x, int64_t R, int Lt, int Lj, int H,
                                                             float* __restrict__ dW, float* __restrict__ db) {
  pdl_sync();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)H * H) return;
  const int i = (int)(idx / H), k = (int)(idx % H);
  float s = 0.f, sb = 0.f;
  for (int64_t r = 0; r < R; ++r) {
    const float g = du[r * H + i];
    s = fmaf(g, x[(r * Lj + Lt) * (int64_t)H + k], s);
    sb += g;
  }
  dW[idx] += s;
  if (k == 0) db[i] += sb;
}
__global__ void rows_loss_add_kernel(const float* __restrict__ rows, int64_t n, float w, float* __restrict__ loss) {
  pdl_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float s = 0.f;
    for (int64_t i = 0; i < n; ++i) s += rows[i];
    *loss += w * s;
  }
}
// d(joint stream)[r, Lt, :] += dv[r, :]   (backward_train, after the text rows were scattered)
__global__ void __launch_bounds__(256) add_first_visual_kernel(const float* __restrict__ dv, int64_t R, int Lt, int Lj, int H, float* __restrict__ gA) {
  pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * H) return;
  gA[((i / H) * Lj + Lt) * (int64_t)H + i % H] += dv[i];
}
int add_first_visual(const float* dv, int64_t R, int Lt, int Lj, int H, float* gA, cudaStream_t st) {
  MSQ_CUDA(launch_k(add_first_visual_kernel, dim3(ceil_div(R * (int64_t)H, 256)), dim3(256), 0, st, dv, R, Lt, Lj, H, gA));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// loss = mean_b nll_b / (N - 1) + lam * mean_b sum_p NLL(softmax(rel6[p, 0:2]), label_p) / P   (1140-1174)
__global__ void __launch_bounds__(256) train_loss_kernel(const float* __restrict__ nll, const float* __restrict__ rel6, const int64_t* __restrict__ labels,
                                                         int64_t B, int N, float lam, float* __restrict__ out) {
  pdl_sync();
  __shared__ float sh[256];
  const int P = N * (N - 1);
  float a = 0.f;
  for (int64_t i = threadIdx.x; i < B * P; i += blockDim.x) {
    const float z0 = rel6[i * 6], z1 = rel6[i * 6 + 1];
    const float mx = fmaxf(z0, z1), lse = mx + logf(expf(z0 - mx) + expf(z1 - mx));
    a += (lse - (labels[i] ? z1 : z0)) / ((float)P + 1e-20f);
  }
  float p = 0.f;
  for (int64_t b = threadIdx.x; b < B; b += blockDim.x) p += nll[b] / ((float)N + 1e-20f - 1.f);
  sh[threadIdx.x] = p + lam * a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[0] = sh[0] / (float)B;
}

// dpw_k[o, blk*(H+2) + c] += dW4[blk*H + o, c]   (inverse of the [4H, Kp] block repack of pw_k.weight)
__global__ void __launch_bounds__(256) unpack_pwk_grad_kernel(const float* __restrict__ dw4, int H, int Kp, float* __restrict__ dpwk) {
  pdl_sync();
  const int D = H + 2;
  const int64_t total = (int64_t)H * 4 * D;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i / (4 * D)), rem = (int)(i % (4 * D)), blk = rem / D, c = rem % D;
    dpwk[i] += dw4[((int64_t)blk * H + o) * Kp + c];
  }
}

// =====================================================================================================
// orchestration
// =====================================================================================================
template <typename T> static int head_wT(TrainState* ts, const Lin& l, void** out, cudaStream_t st, bool alloc) {
  if (alloc) {
    void* p = nullptr;
    MSQ_CUDA(cudaMalloc(&p, (size_t)l.N * l.K * sizeof(T)));
    ts->owned.push_back(p);
    *out = p;
  }
  return transpose_pad<float, T>(l.w32, l.N, l.K, l.ld, l.N, (T*)*out, ACT_NONE, st);
}

// parameter slots + W^T copies of the heads (alloc) / refresh of the copies after an optimizer step (!alloc)
int heads_train_setup(msq_model* m, bool alloc, cudaStream_t st) {
  TrainState* ts = m->train;
  if (!m->has_heads) return MSQ_OK;
  if (alloc) ts->paraT.resize(m->para.size());
  for (size_t l = 0; l < m->para.size(); ++l) {
    const Lin* ls[4] = {&m->para[l].qkv, &m->para[l].fin, &m->para[l].w1, &m->para[l].w2};
    for (int i = 0; i < 4; ++i) MSQ_TRY(head_wT<float>(ts, *ls[i], (void**)&ts->paraT[l][i], st, alloc));
  }
  if (m->cfg.precise) MSQ_TRY(head_wT<float>(ts, m->sent_tran, &ts->sentT, st, alloc));
  else MSQ_TRY(head_wT<bf16>(ts, m->sent_tran, &ts->sentT, st, alloc));
  MSQ_TRY(head_wT<float>(ts, m->key_lin, (void**)&ts->keyT, st, alloc));
  MSQ_TRY(head_wT<float>(ts, m->wq_raw, (void**)&ts->wqT, st, alloc));
  MSQ_TRY(head_wT<float>(ts, m->whh_raw, (void**)&ts->whhT, st, alloc));
  MSQ_TRY(head_wT<float>(ts, m->wih_raw, (void**)&ts->wihT, st, alloc));
  MSQ_TRY(head_wT<float>(ts, m->t4_lin, (void**)&ts->pw4T, st, alloc));
  ts->heads = true;
  return MSQ_OK;
}
// names of the head parameters that receive a gradient (the reference leaves h1/h2_relationship, classifier and
// transformer_inter.0.layer_norm without one: p.grad is None, the optimizer skips them)
std::vector<std::string> heads_param_names(const msq_model* m) {
  std::vector<std::string> v;
  const std::string T = "two_level_encoder.";
  for (const char* e : {"sentence_tran.weight", "sentence_tran.bias", "sentence_tran_2.weight", "sentence_tran_2.bias", "linear_in_2.weight",
                        "pairwise_relationship.weight", "pairwise_relationship.bias"})
    v.push_back(T + e);
  for (size_t l = 0; l < m->para.size(); ++l) {
    const std::string b = "encoder.transformer_inter." + std::to_string(l) + ".";
    for (const char* e : {"self_attn.linear_query.weight", "self_attn.linear_keys.weight", "self_attn.linear_values.weight",
                          "self_attn.linear_query.bias", "self_attn.linear_keys.bias", "self_attn.linear_values.bias",
                          "self_attn.final_linear.weight", "self_attn.final_linear.bias", "feed_forward.w_1.weight", "feed_forward.w_1.bias",
                          "feed_forward.w_2.weight", "feed_forward.w_2.bias", "feed_forward.layer_norm.weight", "feed_forward.layer_norm.bias"})
      v.push_back(b + e);
    if (l > 0) { v.push_back(b + "layer_norm.weight"); v.push_back(b + "layer_norm.bias"); }
  }
  for (const char* e : {"encoder.layer_norm.weight", "encoder.layer_norm.bias", "key_linear.weight", "key_linear.bias", "query_linear.weight",
                        "query_linear.bias", "tanh_linear.weight", "tanh_linear.bias", "decoder.weight_ih_l0", "decoder.weight_hh_l0",
                        "decoder.bias_ih_l0", "decoder.bias_hh_l0", "pw_k.weight"})
    v.push_back(e);
  if (m->raw.count("img_projection.weight") && m->raw.count("img_projection.bias")) {   // args.multimodal_loss (modeling_bert.py:897-898)
    v.push_back("img_projection.weight");
    v.push_back("img_projection.bias");
  }
  return v;
}

static int g32(const float* A, int lda, const float* W, int ldw, const float* bias, const float* resid, int ldr, float* C, int ldc, int64_t M,
               int N, int K, int act, cudaStream_t st) {
  return gemm_nt<float, float>(nullptr, A, lda, W, ldw, bias, resid, ldr, C, ldc, M, N, K, act, st);
}

// Forward of the heads + loss + backward down to ts->ht.d_lang [R, Lt, H].  grads accumulates the head parameters.
template <typename T>
int heads_train(msq_model* m, const int64_t* sep, int64_t B, int N, const int32_t* target, const int64_t* pair_labels, float lam, float* grads,
                float* loss_out, cudaStream_t st) {
  TrainState* ts = m->train;
  MSQ_REQUIRE(ts && ts->have_fwd && ts->heads, "heads_train: no recorded encoder forward / head weights");
  MSQ_REQUIRE(N >= 2 && N <= TH_MAXN, "N=%d out of range [2,%d]", N, TH_MAXN);
  const msq_config& c = m->cfg;
  const int H = c.hidden, Lt = ts->Lt, Lj = ts->Lj, P = N * (N - 1), Kp = m->Kp, ff = c.para_ff;
  const int64_t R = ts->R, M = B * N, Mt = R * Lt, C2 = B * N * N;
  MSQ_REQUIRE(R == B * P, "heads_train: %lld pair rows recorded, expected %lld", (long long)R, (long long)(B * P));
  MSQ_REQUIRE(H <= 1024, "heads_train: hidden size");
  HeadTape& h = ts->ht;
  h.B = B; h.N = N;
  h.pl.assign(m->para.size(), ParaTape{});
  const float* x = ts->x_out;
  int err = MSQ_OK;
  auto G = [&](const std::string& name) -> float* {
    auto it = ts->index.find(name);
    if (it == ts->index.end()) { set_error("train: no gradient slot for %s", name.c_str()); err = MSQ_ERR_STATE; return nullptr; }
    return grads + ts->slots[it->second].off;
  };
  // ---- saved tensors
  for (int pass = 0; pass < 2; ++pass) {
    Planner p{&ts->htape, pass == 0};
    if (pass == 1) ts->htape.reset();
    h.topt = p.take<T>((size_t)Mt * H); h.ttb = p.take<T>((size_t)Mt * H);
    h.mix = p.take<float>((size_t)R * 2 * H); h.rel6 = p.take<float>((size_t)R * 6);
    h.sents = p.take<float>((size_t)M * H); h.r0 = p.take<float>((size_t)C2 * Kp);
    for (auto& L : h.pl) {
      L.y = p.take<float>((size_t)M * H); L.qkv = p.take<float>((size_t)M * 3 * H); L.ctx = p.take<float>((size_t)M * H);
      L.out = p.take<float>((size_t)M * H); L.pn = p.take<float>((size_t)M * H); L.u = p.take<float>((size_t)M * ff);
      L.xo = p.take<float>((size_t)M * H);
      L.hf = ts->drop.p_para > 0.f ? p.take<float>((size_t)M * ff) : nullptr;
    }
    h.para = p.take<float>((size_t)M * H); h.h0 = p.take<float>((size_t)B * H);
    h.keyin = p.take<float>((size_t)M * 2 * H); h.key = p.take<float>((size_t)M * H);
    h.sents_ext = p.take<float>((size_t)B * (N + 1) * H); h.xg = p.take<float>((size_t)B * (N + 1) * 4 * H);
    h.t4 = p.take<float>((size_t)C2 * 4 * H);
    h.act = p.take<float>((size_t)N * B * 4 * H); h.c_all = p.take<float>((size_t)N * B * H);
    h.hs = p.take<float>((size_t)N * B * H);       // [t][b] order: the steps' h outputs
    h.hprev = p.take<float>((size_t)N * B * H);    // [t][b]: h fed INTO step t
    h.query = p.take<float>((size_t)M * H);        // [b][t]
    h.nll = p.take<float>((size_t)B);
    h.d_lang = p.take<float>((size_t)Mt * H);
    if (ts->mm_loss) {
      h.ml_u = p.take<float>((size_t)R * H); h.ml_du = p.take<float>((size_t)R * H); h.ml_dv = p.take<float>((size_t)R * H);
      h.ml_dz = p.take<float>((size_t)R * 2); h.ml_loss = p.take<float>((size_t)R);
    }
    if (pass == 0) MSQ_TRY(ts->htape.reserve(p.need + 4096, st));
  }
  // ---- transient buffers (forward + backward)
  const int64_t MpTok = round_up(Mt, 64), MpS = round_up(max(max(C2, B * (N + 1)), (int64_t)N * B), 64);
  const int wide = max(max(4 * H, ff), max(3 * H, Kp));
  float *hf, *pre, *hsbt, *dmix, *dscore, *drel2, *dsents, *dpara, *dkeyin, *dkey0, *dquery, *dhs, *dt4, *dr0, *dw4, *dxg, *dext, *dgall, *dcbuf,
      *dhrec, *part, *gx, *gy, *gq, *gu, *dh0, *gm;
  void* dpre = nullptr;
  BwdBufs bb{};
  for (int pass = 0; pass < 2; ++pass) {
    Planner p{&m->ws, pass == 0};
    if (pass == 1) m->ws.reset();
    hf = p.take<float>((size_t)M * ff); pre = p.take<float>((size_t)B * 4 * H); hsbt = p.take<float>((size_t)M * H);
    dmix = p.take<float>((size_t)R * 2 * H); dscore = p.take<float>((size_t)Mt); drel2 = p.take<float>((size_t)R * 2);
    dsents = p.take<float>((size_t)M * H); dpara = p.take<float>((size_t)M * H); dkeyin = p.take<float>((size_t)M * 2 * H);
    dkey0 = p.take<float>((size_t)M * H); dquery = p.take<float>((size_t)M * H); dhs = p.take<float>((size_t)M * H);
    dt4 = p.take<float>((size_t)C2 * 4 * H); dr0 = p.take<float>((size_t)C2 * Kp); dw4 = p.take<float>((size_t)4 * H * Kp);
    dxg = p.take<float>((size_t)B * (N + 1) * 4 * H); dext = p.take<float>((size_t)B * (N + 1) * H);
    dgall = p.take<float>((size_t)N * B * 4 * H); dcbuf = p.take<float>((size_t)B * H); dhrec = p.take<float>((size_t)B * H);
    dh0 = p.take<float>((size_t)B * H);
    part = p.take<float>((size_t)max(max((int64_t)148 * 4, M), B) * (H + 1) + ln_bwd_scratch_floats(H));
    gx = p.take<float>((size_t)M * H); gy = p.take<float>((size_t)M * H); gq = p.take<float>((size_t)M * 3 * H); gu = p.take<float>((size_t)M * ff);
    gm = p.take<float>((size_t)M * H);   // dropout: masked copy of a sub-layer output gradient
    dpre = p.take<T>((size_t)Mt * H);
    const size_t tbytes = max((size_t)H * MpTok * sizeof(T), (size_t)wide * MpS * sizeof(float));
    bb.GT = p.take<char>(tbytes);
    bb.XT = p.take<char>(tbytes);
    if (pass == 0) MSQ_TRY(m->ws.reserve(p.need + 4096, st));
  }
  float* ln_scr = part + (size_t)max(max((int64_t)148 * 4, M), B) * (H + 1);

  // ================= forward =================
  MSQ_TRY((gather_rows<float, T>(x, Mt, H, Lt, Lj, 0, (T*)h.topt, st)));
  MSQ_TRY((gemm_nt<T, T>(m, (const T*)h.topt, H, wptr<T>(m->sent_tran), m->sent_tran.ld, m->sent_tran.b, nullptr, 0, (T*)h.ttb, H, Mt, H, H, ACT_TANH, st)));
  const DropCfg& dc = ts->drop;   // masks of this step (keyed by the encoder forward that just ran)
  const bool drop_p = dc.p_para > 0.f;
  MSQ_TRY(token_pool<T>((const T*)h.ttb, x, R, Lt, Lj, H, m->w2, m->b2, sep, m->w_rel, m->b_rel, h.mix, h.rel6, st,
                        make_drop(dc, DROP_H, 0, dc.p_hidden)));
  MSQ_TRY(edge_pool(h.mix, x, h.rel6, B, N, Lj, H, m->w_in2, h.sents, h.r0, Kp, nullptr, nullptr, nullptr, nullptr, nullptr, st));
  {
    const float* xin = h.sents;
    for (size_t l = 0; l < m->para.size(); ++l) {
      ParaLayerW& L = m->para[l];
      ParaTape& t = h.pl[l];
      t.xin = const_cast<float*>(xin);
      const float* y = xin;
      if (l != 0) { MSQ_TRY(layernorm<float>(xin, M, H, L.ln_in.g, L.ln_in.b, 1e-6f, t.y, nullptr, 0, 0, 0, st)); y = t.y; }
      MSQ_TRY(g32(y, H, L.qkv.w32, L.qkv.ld, L.qkv.b, nullptr, 0, t.qkv, 3 * H, M, 3 * H, H, ACT_NONE, st));
      MSQ_TRY(para_attention(t.qkv, B, N, c.para_heads, H, t.ctx, st, make_drop(dc, DROP_PA, (int)l, dc.p_para)));
      if (drop_p) {   // out = dropout(final_linear(ctx)) + x   (encoder.py:28)
        MSQ_TRY(g32(t.ctx, H, L.fin.w32, L.fin.ld, L.fin.b, nullptr, 0, t.out, H, M, H, H, ACT_NONE, st));
        MSQ_TRY(dropout_add(t.out, xin, M * H, make_drop(dc, DROP_PC, (int)l, dc.p_para), st));
      } else {
        MSQ_TRY(g32(t.ctx, H, L.fin.w32, L.fin.ld, L.fin.b, xin, H, t.out, H, M, H, H, ACT_NONE, st));
      }
      MSQ_TRY(layernorm<float>(t.out, M, H, L.ln_ff.g, L.ln_ff.b, 1e-6f, t.pn, nullptr, 0, 0, 0, st));
      MSQ_TRY(g32(t.pn, H, L.w1.w32, L.w1.ld, L.w1.b, nullptr, 0, t.u, ff, M, ff, H, ACT_NONE, st));
      float* hfl = drop_p ? t.hf : hf;   // under dropout the dropped activations are kept for the w_2 weight gradient
      MSQ_TRY(act_fwd<float>(t.u, M * ff, ACT_GELU_TANH, hfl, st));
      if (drop_p) {   // x' = dropout_2(w_2(dropout_1(gelu(.)))) + out   (neural.py:31-33)
        MSQ_TRY(dropout_rows<float>(hfl, nullptr, M, 1, 0, 1, ff, make_drop(dc, DROP_PF1, (int)l, dc.p_para), st));
        MSQ_TRY(g32(hfl, ff, L.w2.w32, L.w2.ld, L.w2.b, nullptr, 0, t.xo, H, M, H, ff, ACT_NONE, st));
        MSQ_TRY(dropout_add(t.xo, t.out, M * H, make_drop(dc, DROP_PF2, (int)l, dc.p_para), st));
      } else {
        MSQ_TRY(g32(hfl, ff, L.w2.w32, L.w2.ld, L.w2.b, t.out, H, t.xo, H, M, H, ff, ACT_NONE, st));
      }
      xin = t.xo;
    }
    MSQ_TRY(layernorm<float>(xin, M, H, m->para_ln.g, m->para_ln.b, 1e-6f, h.para, nullptr, 0, 0, 0, st));
  }
  MSQ_TRY(para_finish(h.sents, h.para, B, N, H, h.h0, h.keyin, st));
  MSQ_TRY(g32(h.keyin, 2 * H, m->key_lin.w32, m->key_lin.ld, m->key_lin.b, nullptr, 0, h.key, H, M, H, 2 * H, ACT_NONE, st));
  // decoder, teacher forced
  MSQ_CUDA(launch_k(sents_ext_train_kernel, dim3(ceil_div(B * (N + 1) * (int64_t)H, 256)), dim3(256), 0, st, (const float*)h.sents, B, N, H, h.sents_ext));
  MSQ_LAUNCH_CHECK();
  MSQ_TRY(g32(h.sents_ext, H, m->wih_raw.w32, m->wih_raw.ld, m->wih_raw.b, nullptr, 0, h.xg, 4 * H, B * (N + 1), 4 * H, H, ACT_NONE, st));
  MSQ_TRY(g32(h.r0, Kp, m->t4_lin.w32, m->t4_lin.ld, nullptr, nullptr, 0, h.t4, 4 * H, C2, 4 * H, Kp, ACT_NONE, st));
  for (int t = 0; t < N; ++t) {
    const float* hp = t == 0 ? h.h0 : h.hs + (size_t)(t - 1) * B * H;
    MSQ_CUDA(cudaMemcpyAsync(h.hprev + (size_t)t * B * H, hp, (size_t)B * H * sizeof(float), cudaMemcpyDeviceToDevice, st));
    MSQ_TRY(g32(hp, H, m->whh_raw.w32, m->whh_raw.ld, m->whh_raw.b, nullptr, 0, pre, 4 * H, B, 4 * H, H, ACT_NONE, st));
    MSQ_CUDA(launch_k(lstm_cell_fwd_kernel, dim3((unsigned)B), dim3(256), 0, st, (const float*)pre, (const float*)h.xg, target, t, N, H,
                      (const float*)(t == 0 ? nullptr : h.c_all + (size_t)(t - 1) * B * H), h.act + (size_t)t * B * 4 * H,
                      h.c_all + (size_t)t * B * H, h.hs + (size_t)t * B * H));
    MSQ_LAUNCH_CHECK();
  }
  // hs is [t][b]; the query GEMM / tf kernel want [b][t]
  MSQ_TRY((transpose_rows(h.hs, N, B, H, hsbt, st)));
  MSQ_TRY(g32(hsbt, H, m->wq_raw.w32, m->wq_raw.ld, m->wq_raw.b, nullptr, 0, h.query, H, M, H, H, ACT_NONE, st));
  MSQ_CUDA(cudaMemsetAsync(dkey0, 0, (size_t)M * H * sizeof(float), st));
  MSQ_CUDA(cudaMemsetAsync(dt4, 0, (size_t)C2 * 4 * H * sizeof(float), st));
  const float scale = 1.0f / (((float)N + 1e-20f - 1.f) * (float)B);
  MSQ_SMEM_ATTR((int)((size_t)TH_MAXN * 1024 * sizeof(float)), tf_pointer_kernel);
  MSQ_CUDA(launch_k(tf_pointer_kernel, dim3((unsigned)B), dim3(256), (size_t)N * H * sizeof(float), st, (const float*)h.query, (const float*)h.t4,
                    (const float*)h.key, target, m->dec.wt, m->dec.bt, N, H, scale, h.nll, dquery, dkey0, dt4, part));
  MSQ_LAUNCH_CHECK();
  if (loss_out) {
    MSQ_CUDA(launch_k(train_loss_kernel, dim3(1), dim3(256), 0, st, (const float*)h.nll, (const float*)h.rel6, pair_labels, B, N, lam, loss_out));
    MSQ_LAUNCH_CHECK();
  }

  // ================= backward =================
  {
    float *dwt = G("tanh_linear.weight"), *dbt = G("tanh_linear.bias");
    if (err) return err;
    MSQ_TRY(partial_reduce(part, (int)B, H + 1, 0, H, dwt, st));
    MSQ_TRY(partial_reduce(part, (int)B, H + 1, H, 1, dbt, st));
  }
  // query_linear
  MSQ_TRY(wgrad<float>(m, dquery, H, H, hsbt, H, H, ACT_NONE, M, G("query_linear.weight"), G("query_linear.bias"), bb, st));
  MSQ_TRY((dgrad<float, float>(m, dquery, H, ts->wqT, H, nullptr, dhs, M, st)));     // [b][t]
  // pw_k through T4 = R0 W4^T
  MSQ_CUDA(cudaMemsetAsync(dw4, 0, (size_t)4 * H * Kp * sizeof(float), st));
  MSQ_TRY(wgrad<float>(m, dt4, 4 * H, 4 * H, h.r0, Kp, Kp, ACT_NONE, C2, dw4, nullptr, bb, st));
  {
    float* dpwk = G("pw_k.weight");
    if (err) return err;
    MSQ_CUDA(launch_k(unpack_pwk_grad_kernel, dim3(148 * 4), dim3(256), 0, st, (const float*)dw4, H, Kp, dpwk));
    MSQ_LAUNCH_CHECK();
  }
  MSQ_TRY((dgrad<float, float>(m, dt4, 4 * H, ts->pw4T, Kp, nullptr, dr0, C2, st)));
  // LSTM, backward through time
  MSQ_CUDA(cudaMemsetAsync(dxg, 0, (size_t)B * (N + 1) * 4 * H * sizeof(float), st));
  MSQ_CUDA(cudaMemsetAsync(dcbuf, 0, (size_t)B * H * sizeof(float), st));
  MSQ_TRY((transpose_rows(dhs, B, N, H, gx, st)));   // gx reused as dhs in [t][b] order (M*H floats)
  for (int t = N - 1; t >= 0; --t) {
    MSQ_CUDA(launch_k(lstm_cell_bwd_kernel, dim3((unsigned)B), dim3(256), 0, st, (const float*)(gx + (size_t)t * B * H),
                      (const float*)(t == N - 1 ? nullptr : dhrec), (const float*)(h.act + (size_t)t * B * 4 * H),
                      (const float*)(h.c_all + (size_t)t * B * H), (const float*)(t == 0 ? nullptr : h.c_all + (size_t)(t - 1) * B * H), dcbuf,
                      target, t, N, H, dgall + (size_t)t * B * 4 * H, dxg));
    MSQ_LAUNCH_CHECK();
    MSQ_TRY((dgrad<float, float>(m, dgall + (size_t)t * B * 4 * H, 4 * H, ts->whhT, H, nullptr, t == 0 ? dh0 : dhrec, B, st)));
  }
  MSQ_TRY(wgrad<float>(m, dgall, 4 * H, 4 * H, h.hprev, H, H, ACT_NONE, (int64_t)N * B, G("decoder.weight_hh_l0"), G("decoder.bias_hh_l0"), bb, st));
  MSQ_TRY(wgrad<float>(m, dxg, 4 * H, 4 * H, h.sents_ext, H, H, ACT_NONE, B * (N + 1), G("decoder.weight_ih_l0"), G("decoder.bias_ih_l0"), bb, st));
  MSQ_TRY((dgrad<float, float>(m, dxg, 4 * H, ts->wihT, H, nullptr, dext, B * (N + 1), st)));
  MSQ_CUDA(cudaMemsetAsync(dsents, 0, (size_t)M * H * sizeof(float), st));
  MSQ_CUDA(launch_k(sents_ext_bwd_kernel, dim3(ceil_div(M * (int64_t)H, 256)), dim3(256), 0, st, (const float*)dext, B, N, H, dsents));
  MSQ_LAUNCH_CHECK();
  if (ts->trip && ts->trip_weight != 0.f) {   // time-contrastive objective set for this step (msq_train_set_triplets)
    MSQ_REQUIRE(ts->trip_B == B, "msq_train_step: %lld triplets were set for a batch of %lld manuals", (long long)ts->trip_B, (long long)B);
    MSQ_CUDA(launch_k(triplet_kernel, dim3((unsigned)B), dim3(256), 0, st, (const float*)h.sents, (const int32_t*)ts->trip, N, H,
                      ts->trip_weight / (float)B, ts->trip_loss, dsents));
    MSQ_LAUNCH_CHECK();
    if (loss_out) {
      MSQ_CUDA(launch_k(triplet_loss_add_kernel, dim3(1), dim3(32), 0, st, (const float*)ts->trip_loss, B, ts->trip_weight, loss_out));
      MSQ_LAUNCH_CHECK();
    }
    ts->trip_weight = 0.f;   // one-shot
  }
  // key_linear, h0, paragraph encoder
  MSQ_TRY(wgrad<float>(m, dkey0, H, H, h.keyin, 2 * H, 2 * H, ACT_NONE, M, G("key_linear.weight"), G("key_linear.bias"), bb, st));
  MSQ_TRY((dgrad<float, float>(m, dkey0, H, ts->keyT, 2 * H, nullptr, dkeyin, M, st)));
  MSQ_CUDA(launch_k(para_finish_bwd_kernel, dim3((unsigned)B), dim3(256), 0, st, (const float*)dkeyin, (const float*)dh0, N, H, dsents, dpara));
  MSQ_LAUNCH_CHECK();
  {
    const size_t nl = m->para.size();
    const float* xlast = nl ? h.pl[nl - 1].xo : h.sents;
    MSQ_TRY(ln_bwd<float>(dpara, xlast, nullptr, M, H, m->para_ln.g, 1e-6f, gx, nullptr, G("encoder.layer_norm.weight"), G("encoder.layer_norm.bias"),
                          ln_scr, 0, 0, 0, st));
    for (size_t li = nl; li-- > 0;) {
      ParaLayerW& L = m->para[li];
      ParaTape& t = h.pl[li];
      auto& WT = ts->paraT[li];
      const std::string bn = "encoder.transformer_inter." + std::to_string(li) + ".";
      if (drop_p) {
        // gx = d(x'): the w_2 output sees gx . mask_2; its input was the DROPPED activation (kept in t.hf)
        MSQ_TRY(dropout_mask_copy<float>(gx, gm, M * H, make_drop(dc, DROP_PF2, (int)li, dc.p_para), st));
        MSQ_TRY(wgrad<float>(m, gm, H, H, t.hf, ff, ff, ACT_NONE, M, G(bn + "feed_forward.w_2.weight"), G(bn + "feed_forward.w_2.bias"), bb, st));
        MSQ_TRY((dgrad<float, float>(m, gm, H, WT[3], ff, nullptr, gu, M, st)));
        MSQ_TRY(dropout_rows<float>(gu, nullptr, M, 1, 0, 1, ff, make_drop(dc, DROP_PF1, (int)li, dc.p_para), st));
      } else {
        MSQ_TRY(wgrad<float>(m, gx, H, H, t.u, ff, ff, ACT_GELU_TANH, M, G(bn + "feed_forward.w_2.weight"), G(bn + "feed_forward.w_2.bias"), bb, st));
        MSQ_TRY((dgrad<float, float>(m, gx, H, WT[3], ff, nullptr, gu, M, st)));
      }
      MSQ_TRY(act_bwd<float>(gu, t.u, M * ff, ACT_GELU_TANH, gu, st));
      MSQ_TRY(wgrad<float>(m, gu, ff, ff, t.pn, H, H, ACT_NONE, M, G(bn + "feed_forward.w_1.weight"), G(bn + "feed_forward.w_1.bias"), bb, st));
      MSQ_TRY((dgrad<float, float>(m, gu, ff, WT[2], H, nullptr, gy, M, st)));
      MSQ_TRY(ln_bwd<float>(gy, t.out, gx, M, H, L.ln_ff.g, 1e-6f, gx, nullptr, G(bn + "feed_forward.layer_norm.weight"),
                            G(bn + "feed_forward.layer_norm.bias"), ln_scr, 0, 0, 0, st));                        // gx = d(out)
      const float* gfin = gx;   // gradient at the final_linear output: d(out) . mask under dropout
      if (drop_p) {
        MSQ_TRY(dropout_mask_copy<float>(gx, gm, M * H, make_drop(dc, DROP_PC, (int)li, dc.p_para), st));
        gfin = gm;
      }
      MSQ_TRY(wgrad<float>(m, gfin, H, H, t.ctx, H, H, ACT_NONE, M, G(bn + "self_attn.final_linear.weight"), G(bn + "self_attn.final_linear.bias"), bb, st));
      MSQ_TRY((dgrad<float, float>(m, gfin, H, WT[1], H, nullptr, gy, M, st)));                                    // d(ctx)
      MSQ_CUDA(launch_k(para_attention_bwd_kernel, dim3((unsigned)(B * c.para_heads)), dim3(128), 0, st, (const float*)t.qkv, (const float*)gy, N,
                        (int)c.para_heads, H, gq, make_drop(dc, DROP_PA, (int)li, dc.p_para)));
      MSQ_LAUNCH_CHECK();
      const float* y = li == 0 ? t.xin : t.y;
      MSQ_TRY(wgrad<float>(m, gq, 3 * H, 3 * H, y, H, H, ACT_NONE, M, G(bn + "self_attn.linear_query.weight"), G(bn + "self_attn.linear_query.bias"), bb, st));
      if (li == 0) {
        MSQ_TRY((dgrad<float, float>(m, gq, 3 * H, WT[0], H, gx, gx, M, st)));                                     // d(xin) = d(out) + d(y)
      } else {
        MSQ_TRY((dgrad<float, float>(m, gq, 3 * H, WT[0], H, nullptr, gy, M, st)));
        MSQ_TRY(ln_bwd<float>(gy, t.xin, gx, M, H, L.ln_in.g, 1e-6f, gx, nullptr, G(bn + "layer_norm.weight"), G(bn + "layer_norm.bias"), ln_scr, 0, 0, 0, st));
      }
      if (err) return err;
    }
    MSQ_TRY(axpy(gx, 1.0f, M * (int64_t)H, dsents, st));
  }
  if (err) return err;
  // edge pooling, pair heads, token pooling
  MSQ_CUDA(launch_k(edge_pool_bwd_kernel, dim3((unsigned)M), dim3(256), 0, st, (const float*)h.mix, (const float*)dsents, N, H, m->w_in2, dmix, part));
  MSQ_LAUNCH_CHECK();
  MSQ_TRY(partial_reduce(part, (int)M, H, 0, H, G("two_level_encoder.linear_in_2.weight"), st));
  MSQ_CUDA(launch_k(token_pool_bwd_kernel<T>, dim3((unsigned)R), dim3(256), 3 * Lt * sizeof(float), st, (const T*)h.ttb, x, Lt, Lj, H, m->w2, m->b2, sep,
                    (const float*)dmix, h.d_lang, dscore, make_drop(dc, DROP_H, 0, dc.p_hidden)));
  MSQ_LAUNCH_CHECK();
  const float lam_scale = lam / (((float)P + 1e-20f) * (float)B);
  MSQ_CUDA(launch_k(cls_rel_bwd_kernel, dim3((unsigned)R), dim3(256), 0, st, (const float*)dr0, Kp, (const float*)h.rel6, pair_labels, m->w_rel, N, Lt, H,
                    lam_scale, h.d_lang, drel2));
  MSQ_LAUNCH_CHECK();
  MSQ_CUDA(launch_k(rel_wgrad_kernel, dim3(ceil_div(2 * H, 256)), dim3(256), 0, st, (const float*)drel2, x, R, Lj, H,
                    G("two_level_encoder.pairwise_relationship.weight"), G("two_level_encoder.pairwise_relationship.bias")));
  MSQ_LAUNCH_CHECK();
  // sentence_tran_2 / sentence_tran
  {
    const int nblk = (int)min((int64_t)148 * 4, (Mt + SB_WARPS - 1) / SB_WARPS);
    MSQ_CUDA(launch_k(score_bwd_kernel<T>, dim3(nblk), dim3(SB_WARPS * 32), 0, st, (const float*)dscore, (const T*)h.ttb, Mt, H, m->w2, (T*)dpre, part));
    MSQ_LAUNCH_CHECK();
    MSQ_TRY(partial_reduce(part, nblk, H + 1, 0, H, G("two_level_encoder.sentence_tran_2.weight"), st));
    MSQ_TRY(partial_reduce(part, nblk, H + 1, H, 1, G("two_level_encoder.sentence_tran_2.bias"), st));
  }
  if (err) return err;
  MSQ_TRY(wgrad<T>(m, (const T*)dpre, H, H, (const T*)h.topt, H, H, ACT_NONE, Mt, G("two_level_encoder.sentence_tran.weight"),
                   G("two_level_encoder.sentence_tran.bias"), bb, st));
  MSQ_TRY((dgrad<T, float>(m, (const T*)dpre, H, ts->sentT, H, h.d_lang, h.d_lang, Mt, st)));
  ts->ml_live = false;
  if (ts->mm_loss) {   // image pairwise objective: shares pairwise_relationship with the text head, adds into the same slots
    MSQ_REQUIRE(ts->mm, "multimodal_loss: the model has no visual stream");
    auto wp = m->raw.find("img_projection.weight");
    auto bp = m->raw.find("img_projection.bias");
    MSQ_REQUIRE(wp != m->raw.end() && bp != m->raw.end(), "multimodal_loss: img_projection.weight / .bias are not registered");
    MSQ_REQUIRE(wp->second.second == (int64_t)H * H && bp->second.second == H,
                "multimodal_loss: img_projection must map the %d-d visual token to %d features (the reference multiplies a [*, %d] matrix by it)", H, H, H);
    float *dWp = G("img_projection.weight"), *dbp = G("img_projection.bias");
    if (err) return err;
    MSQ_CUDA(launch_k(img_pair_kernel, dim3((unsigned)R), dim3(256), 0, st, x, Lt, Lj, H, wp->second.first, bp->second.first, m->w_rel, m->b_rel,
                      pair_labels, lam_scale, h.ml_u, h.ml_du, h.ml_dz, h.ml_dv, h.ml_loss));
    MSQ_LAUNCH_CHECK();
    MSQ_CUDA(launch_k(img_proj_wgrad_kernel, dim3(ceil_div((int64_t)H * H, 256)), dim3(256), 0, st, (const float*)h.ml_du, x, R, Lt, Lj, H, dWp, dbp));
    MSQ_LAUNCH_CHECK();
    MSQ_CUDA(launch_k(rel_wgrad_kernel, dim3(ceil_div(2 * H, 256)), dim3(256), 0, st, (const float*)h.ml_dz, (const float*)h.ml_u, R, 1, H,
                      G("two_level_encoder.pairwise_relationship.weight"), G("two_level_encoder.pairwise_relationship.bias")));
    MSQ_LAUNCH_CHECK();
    if (loss_out) {
      MSQ_CUDA(launch_k(rows_loss_add_kernel, dim3(1), dim3(32), 0, st, (const float*)h.ml_loss, R, lam_scale, loss_out));
      MSQ_LAUNCH_CHECK();
    }
    ts->ml_live = true;
  }
  return err;
}
template int heads_train<float>(msq_model*, const int64_t*, int64_t, int, const int32_t*, const int64_t*, float, float*, float*, cudaStream_t);
template int heads_train<bf16>(msq_model*, const int64_t*, int64_t, int, const int32_t*, const int64_t*, float, float*, float*, cudaStream_t);

}  // namespace msq
