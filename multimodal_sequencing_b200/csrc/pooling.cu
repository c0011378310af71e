// BERSON hierarchical pooling and the small paragraph-encoder glue, without the reference's host
// loops and .cpu() syncs.
//
// Replaces (telin0411/multimodal_sequencing, models/berson/modeling_bert.py):
//   token level  697-738 : score = w2 . tanh(Ws x + bs) + b2 (the tanh(GEMM) arrives as `tt`), two
//                          masked softmaxes over spans [1..sep0] and [sep0+1..sep1], weighted sums.
//   pair heads   745-757 : pairwise / h1 / h2 relationship Linear(H -> 2) on the pair CLS vector.
//   edge level   766-814 : scatter into per-step edge lists in pair order, softmax(linear_in_2), sum;
//                          N x N tables cls_output_matrix / cls_score_matrix (+his1/his2).
//   rela_encode  919-925 : R0[i,j] = [cls(i,j) ; softmax(score(i,j))]  (row stride padded for the GEMM).
//   paragraph    models/berson/neural.py:200-226 (tiny 8-head attention over the N step vectors),
//                modeling_bert.py:1346-1357 (para mean -> h0, key_linear input concat).
#include "kernels.cuh"

namespace msq {

constexpr int PL_MAXN = 16;

template <typename T>
__global__ void __launch_bounds__(256) token_pool_kernel(const T* __restrict__ tt, const float* __restrict__ x, int Lt, int Lj,
                                                         int H, const float* __restrict__ w2, const float* __restrict__ b2,
                                                         const int64_t* __restrict__ sep, const float* __restrict__ w_rel,
                                                         const float* __restrict__ b_rel, float* __restrict__ mix,
                                                         float* __restrict__ rel6, Drop drop) {
  pdl_sync();
  extern __shared__ float sm[];
  float* score = sm;        // [Lt]
  float* p0 = sm + Lt;      // [Lt]
  float* p1 = p0 + Lt;      // [Lt]
  const int64_t r = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int sep0 = (int)sep[r * 2], sep1 = (int)sep[r * 2 + 1];

  for (int t = warp; t < Lt; t += nw) {
    const T* row = tt + (r * Lt + t) * H;
    float a = 0.f;
    for (int d = lane; d < H; d += 32) a = fmaf(to_f(row[d]), w2[d], a);
    a = warp_sum(a);
    if (lane == 0) score[t] = a + b2[0];
  }
  // relationship heads on the pair CLS (position 0 of the text stream)
  if (warp < 6) {
    const float* cls = x + r * Lj * (int64_t)H;
    float a = 0.f;
    for (int d = lane; d < H; d += 32) a = fmaf(cls[d], w_rel[warp * H + d], a);
    a = warp_sum(a);
    if (lane == 0) rel6[r * 6 + warp] = a + b_rel[warp];
  }
  __syncthreads();
  if (warp < 2) {
    const int lo = warp == 0 ? 1 : sep0 + 1, hi = warp == 0 ? sep0 : sep1;  // inclusive span
    float* p = warp == 0 ? p0 : p1;
    float mx = -INFINITY;
    for (int t = lo + lane; t <= hi; t += 32) mx = fmaxf(mx, score[t]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int t = lane; t < Lt; t += 32) {
      const float e = (t >= lo && t <= hi) ? expf(score[t] - mx) : 0.f;
      p[t] = e;
      s += e;
    }
    s = warp_sum(s);
    const float inv = 1.f / s;
    // training: dropout of the token-attention probabilities [R, 2, Lt] (modeling_bert.py:735)
    for (int t = lane; t < Lt; t += 32) p[t] *= inv * drop_mul(drop, ((uint64_t)r * 2 + warp) * Lt + t);
  }
  __syncthreads();
  const float* xr = x + r * Lj * (int64_t)H;
  for (int d = threadIdx.x; d < H; d += blockDim.x) {
    float a0 = 0.f, a1 = 0.f;
    for (int t = 1; t <= sep0; ++t) a0 = fmaf(p0[t], xr[(int64_t)t * H + d], a0);
    for (int t = sep0 + 1; t <= sep1; ++t) a1 = fmaf(p1[t], xr[(int64_t)t * H + d], a1);
    mix[(r * 2 + 0) * H + d] = a0;
    mix[(r * 2 + 1) * H + d] = a1;
  }
}

template <typename T>
int token_pool(const T* tt, const float* x, int64_t R, int Lt, int Lj, int H, const float* w2, const float* b2,
               const int64_t* sep, const float* w_rel, const float* b_rel, float* mix, float* rel6, cudaStream_t st, const Drop& drop) {
  if (R == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(token_pool_kernel<T>, dim3((unsigned)R), dim3(256), 3 * Lt * sizeof(float), st, tt, x, Lt, Lj, H, w2, b2, sep, w_rel, b_rel, mix, rel6, drop));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}
template int token_pool<float>(const float*, const float*, int64_t, int, int, int, const float*, const float*, const int64_t*,
                               const float*, const float*, float*, float*, cudaStream_t, const Drop&);
template int token_pool<bf16>(const bf16*, const float*, int64_t, int, int, int, const float*, const float*, const int64_t*,
                              const float*, const float*, float*, float*, cudaStream_t, const Drop&);

// pair index of the ordered pair (i, j), i != j, in pairs_generator order
// (models/berson/process_inputs_for_berson.py:246-261): combinations first, then the mirrored list.
__device__ __forceinline__ int pair_index(int i, int j, int N) {
  const int a = i < j ? i : j, b = i < j ? j : i;
  const int c = a * N - a * (a + 1) / 2 + (b - a - 1);
  return i < j ? c : c + N * (N - 1) / 2;
}

// One block per (manual b, step s).
__global__ void __launch_bounds__(256) edge_pool_kernel(const float* __restrict__ mix, const float* __restrict__ x,
                                                        const float* __restrict__ rel6, int N, int Lj, int H,
                                                        const float* __restrict__ w_in2, float* __restrict__ sents,
                                                        float* __restrict__ r0, int r0_ld, float* __restrict__ cls_mat,
                                                        float* __restrict__ score_mat, float* __restrict__ his1,
                                                        float* __restrict__ his2, float* __restrict__ cls_out) {
  pdl_sync();
  __shared__ int e_pair[2 * PL_MAXN], e_side[2 * PL_MAXN];
  __shared__ float e_w[2 * PL_MAXN];
  const int b = blockIdx.x / N, s = blockIdx.x % N;
  const int P = N * (N - 1), E = 2 * (N - 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) {
    int n = 0;  // edge slots in pair order (modeling_bert.py:774-782)
    for (int half = 0; half < 2; ++half)
      for (int i = 0; i < N; ++i)
        for (int j = i + 1; j < N; ++j) {
          const int first = half ? j : i, second = half ? i : j;
          const int p = pair_index(first, second, N);
          if (first == s) { e_pair[n] = p; e_side[n] = 0; ++n; }
          else if (second == s) { e_pair[n] = p; e_side[n] = 1; ++n; }
        }
  }
  __syncthreads();
  const float* mb = mix + (int64_t)b * P * 2 * H;
  for (int e = warp; e < E; e += nw) {
    const float* v = mb + ((int64_t)e_pair[e] * 2 + e_side[e]) * H;
    float a = 0.f;
    for (int d = lane; d < H; d += 32) a = fmaf(v[d], w_in2[d], a);
    a = warp_sum(a);
    if (lane == 0) e_w[e] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = -INFINITY, sum = 0.f;
    for (int e = 0; e < E; ++e) mx = fmaxf(mx, e_w[e]);
    for (int e = 0; e < E; ++e) { e_w[e] = expf(e_w[e] - mx); sum += e_w[e]; }
    for (int e = 0; e < E; ++e) e_w[e] /= sum;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < H; d += blockDim.x) {
    float a = 0.f;
    for (int e = 0; e < E; ++e) a = fmaf(e_w[e], mb[((int64_t)e_pair[e] * 2 + e_side[e]) * H + d], a);
    sents[((int64_t)b * N + s) * H + d] = a;
  }
  // row i = s of the N x N tables
  for (int j = 0; j < N; ++j) {
    const int64_t cell = ((int64_t)b * N + s) * N + j;
    float* ro = r0 + cell * r0_ld;
    if (j == s) {
      for (int d = threadIdx.x; d < r0_ld; d += blockDim.x) ro[d] = (d == H || d == H + 1) ? 0.5f : 0.f;
      if (cls_mat) for (int d = threadIdx.x; d < H; d += blockDim.x) cls_mat[cell * H + d] = 0.f;
      if (score_mat && threadIdx.x < 2) {
        score_mat[cell * 2 + threadIdx.x] = 0.f;
        his1[cell * 2 + threadIdx.x] = 0.f;
        his2[cell * 2 + threadIdx.x] = 0.f;
      }
      continue;
    }
    const int64_t pr = (int64_t)b * P + pair_index(s, j, N);
    const float* cls = x + pr * Lj * (int64_t)H;
    const float z0 = rel6[pr * 6 + 0], z1 = rel6[pr * 6 + 1];
    const float mx = fmaxf(z0, z1), e0 = expf(z0 - mx), e1 = expf(z1 - mx);
    for (int d = threadIdx.x; d < r0_ld; d += blockDim.x)
      ro[d] = d < H ? cls[d] : (d == H ? e0 / (e0 + e1) : (d == H + 1 ? e1 / (e0 + e1) : 0.f));
    if (cls_mat) for (int d = threadIdx.x; d < H; d += blockDim.x) cls_mat[cell * H + d] = cls[d];
    if (cls_out) for (int d = threadIdx.x; d < H; d += blockDim.x) cls_out[pr * H + d] = cls[d];
    if (score_mat && threadIdx.x < 2) {
      score_mat[cell * 2 + threadIdx.x] = rel6[pr * 6 + threadIdx.x];
      his1[cell * 2 + threadIdx.x] = rel6[pr * 6 + 2 + threadIdx.x];
      his2[cell * 2 + threadIdx.x] = rel6[pr * 6 + 4 + threadIdx.x];
    }
  }
}

int edge_pool(const float* mix, const float* x, const float* rel6, int64_t B, int N, int Lj, int H, const float* w_in2,
              float* sents, float* r0, int r0_ld, float* cls_mat, float* score_mat, float* his1, float* his2, float* cls_out,
              cudaStream_t st) {
  MSQ_REQUIRE(N >= 2 && N <= PL_MAXN, "edge_pool: N=%d out of range [2,%d]", N, PL_MAXN);
  MSQ_REQUIRE((score_mat == nullptr) == (his1 == nullptr) && (his1 == nullptr) == (his2 == nullptr), "edge_pool: tables");
  if (B == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(edge_pool_kernel, dim3((unsigned)(B * N)), dim3(256), 0, st, mix, x, rel6, N, Lj, H, w_in2, sents, r0, r0_ld, cls_mat, score_mat, his1, his2, cls_out));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// qkv [B*N, 3H] (q | k | v), heads x (H/heads).  One block per (manual, head); mask is all-ones
// (every manual has exactly N steps, process_inputs_for_berson.py:133).
__global__ void __launch_bounds__(128) para_attention_kernel(const float* __restrict__ qkv, int N, int heads, int H,
                                                             float* __restrict__ ctx, Drop drop) {
  pdl_sync();
  __shared__ float s[PL_MAXN][PL_MAXN + 1];
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int d = H / heads;
  const float inv = 1.0f / sqrtf((float)d);
  const float* base = qkv + (int64_t)b * N * 3 * H + h * d;
  for (int ij = threadIdx.x; ij < N * N; ij += blockDim.x) {
    const int i = ij / N, j = ij % N;
    const float* q = base + (int64_t)i * 3 * H;
    const float* k = base + (int64_t)j * 3 * H + H;
    float a = 0.f;
    for (int e = 0; e < d; ++e) a = fmaf(q[e] * inv, k[e], a);  // query pre-scaled (neural.py:205)
    s[i][j] = a;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    const int i = threadIdx.x;
    float mx = -INFINITY, sum = 0.f;
    for (int j = 0; j < N; ++j) mx = fmaxf(mx, s[i][j]);
    for (int j = 0; j < N; ++j) { s[i][j] = expf(s[i][j] - mx); sum += s[i][j]; }
    // training: attention dropout [B, heads, N, N] (neural.py:228)
    for (int j = 0; j < N; ++j) s[i][j] = s[i][j] / sum * drop_mul(drop, ((uint64_t)blockIdx.x * N + i) * N + j);
  }
  __syncthreads();
  for (int ie = threadIdx.x; ie < N * d; ie += blockDim.x) {
    const int i = ie / d, e = ie % d;
    float a = 0.f;
    for (int j = 0; j < N; ++j) a = fmaf(s[i][j], base[(int64_t)j * 3 * H + 2 * H + e], a);
    ctx[((int64_t)b * N + i) * H + h * d + e] = a;
  }
}

int para_attention(const float* qkv, int64_t B, int N, int heads, int H, float* ctx, cudaStream_t st, const Drop& drop) {
  MSQ_REQUIRE(N <= PL_MAXN && H % heads == 0, "para_attention: N=%d heads=%d H=%d", N, heads, H);
  if (B == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(para_attention_kernel, dim3((unsigned)(B * heads)), dim3(128), 0, st, qkv, N, heads, H, ctx, drop));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// h0[b] = sum_s para[b,s] / (N + 1e-20);  keyin[b,s] = [sents[b,s] ; para[b,s]]
__global__ void __launch_bounds__(256) para_finish_kernel(const float* __restrict__ sents, const float* __restrict__ para, int N,
                                                          int H, float* __restrict__ h0, float* __restrict__ keyin) {
  pdl_sync();
  const int b = blockIdx.x;
  const float den = (float)N + 1e-20f;
  for (int d = threadIdx.x; d < H; d += blockDim.x) {
    float a = 0.f;
    for (int s = 0; s < N; ++s) {
      const float p = para[((int64_t)b * N + s) * H + d];
      a += p;
      keyin[((int64_t)b * N + s) * 2 * H + d] = sents[((int64_t)b * N + s) * H + d];
      keyin[((int64_t)b * N + s) * 2 * H + H + d] = p;
    }
    h0[(int64_t)b * H + d] = a / den;
  }
}

int para_finish(const float* sents, const float* para, int64_t B, int N, int H, float* h0, float* keyin, cudaStream_t st) {
  if (B == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(para_finish_kernel, dim3((unsigned)B), dim3(256), 0, st, sents, para, N, H, h0, keyin));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

}  // namespace msq
