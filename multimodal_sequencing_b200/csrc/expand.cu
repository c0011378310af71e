// Device-side pair expansion: a manual's token row [CLS] s_0 [SEP] [CLS] s_1 [SEP] ... -> its P = N(N-1) ordered
// pair rows (ids, attention mask, token types, [SEP] positions) and the pair -> image table, without a host round trip
// of the expanded tensors.
//
// Replaces the host code of models/berson/process_inputs_for_berson.py:113-243 (pair text assembly), 246-261
// (pairs_generator: all i<j lexicographic, then the mirrored list), 264-368 (padding / masks / token types) and 82-97 (the
// per-pair image gather, here an index table over the UNIQUE images).  Reference quirks kept: the attention mask of a padded
// position is pad_id (not 0), token types are all zero when cls_id == 0 (RoBERTa).
#include "kernels.cuh"

namespace msq {

constexpr int EX_MAXN = 16;

// one block per manual: positions of the N [CLS] / [SEP] tokens in order -> starts, lens; status bit 0 set when a manual does
// not hold exactly N of each; lt_max <- max over ordered pairs of len_i + len_j
__global__ void __launch_bounds__(128) scan_steps_kernel(const int64_t* __restrict__ ids, int64_t B, int L, int N, int64_t cls_id,
                                                         int64_t sep_id, int32_t* __restrict__ starts, int32_t* __restrict__ lens,
                                                         int32_t* __restrict__ meta /* [0] = lt_max, [1] = status */) {
  pdl_sync();
  __shared__ int s_cls[EX_MAXN], s_sep[EX_MAXN], n_cls, n_sep;
  const int64_t b = blockIdx.x;
  const int64_t* row = ids + b * L;
  if (threadIdx.x == 0) { n_cls = 0; n_sep = 0; }
  __syncthreads();
  // ordered compaction: a warp ballot per 32-position chunk, chunks visited in order by warp 0
  if (threadIdx.x < 32) {
    int nc = 0, ns = 0;
    for (int p0 = 0; p0 < L; p0 += 32) {
      const int p = p0 + threadIdx.x;
      const int64_t v = p < L ? row[p] : (int64_t)-1 - cls_id - sep_id;
      const unsigned mc = __ballot_sync(0xffffffffu, p < L && v == cls_id), ms = __ballot_sync(0xffffffffu, p < L && v == sep_id);
      const unsigned below = (1u << threadIdx.x) - 1u;
      if (p < L && v == cls_id) { const int k = nc + __popc(mc & below); if (k < EX_MAXN) s_cls[k] = p; }
      if (p < L && v == sep_id) { const int k = ns + __popc(ms & below); if (k < EX_MAXN) s_sep[k] = p; }
      nc += __popc(mc); ns += __popc(ms);
    }
    if (threadIdx.x == 0) { n_cls = nc; n_sep = ns; }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (n_cls != N || n_sep != N) {
      atomicOr(&meta[1], 1);
      for (int i = 0; i < N; ++i) { starts[b * N + i] = 0; lens[b * N + i] = 1; }
    } else {
      int top = 0, second = 0;
      for (int i = 0; i < N; ++i) {
        const int len = s_sep[i] - s_cls[i] + 1;
        starts[b * N + i] = s_cls[i];
        lens[b * N + i] = len;
        if (len > top) { second = top; top = len; } else if (len > second) { second = len; }
      }
      atomicMax(&meta[0], top + second);
    }
  }
}

// pair p of a manual: p < P/2 -> the p-th (i<j) combination in lexicographic order; p >= P/2 -> its mirror
__device__ __forceinline__ void pair_of(int p, int N, int* i, int* j) {
  const int half = N * (N - 1) / 2;
  int q = p < half ? p : p - half, a = 0;
  while (q >= N - 1 - a) { q -= N - 1 - a; ++a; }
  const int c = a + 1 + q;
  if (p < half) { *i = a; *j = c; } else { *i = c; *j = a; }
}

__global__ void __launch_bounds__(128) expand_pairs_kernel(const int64_t* __restrict__ ids, int L, int N, int Lt, int64_t cls_id,
                                                           int64_t pad_id, const int32_t* __restrict__ starts,
                                                           const int32_t* __restrict__ lens, int64_t* __restrict__ out_ids,
                                                           int64_t* __restrict__ out_mask, int64_t* __restrict__ out_tt,
                                                           int64_t* __restrict__ out_sep, int32_t* __restrict__ img_index) {
  pdl_sync();
  const int P = N * (N - 1);
  const int64_t r = blockIdx.x, b = r / P;
  int i, j;
  pair_of((int)(r % P), N, &i, &j);
  const int s1 = starts[b * N + i], s2 = starts[b * N + j], l1 = lens[b * N + i], l2 = lens[b * N + j];
  const int64_t* row = ids + b * L;
  for (int t = threadIdx.x; t < Lt; t += blockDim.x) {
    const bool in1 = t < l1, in2 = t >= l1 && t < l1 + l2;
    int src = in1 ? s1 + t : s2 + t - l1;
    src = min(max(src, 0), L - 1);
    const bool valid = in1 || in2;
    out_ids[r * Lt + t] = valid ? row[src] : pad_id;
    out_mask[r * Lt + t] = valid ? 1 : pad_id;
    out_tt[r * Lt + t] = (in2 && cls_id != 0) ? 1 : 0;
  }
  if (threadIdx.x == 0) {
    out_sep[r * 2] = l1 - 1;
    out_sep[r * 2 + 1] = l1 + l2 - 1;
    if (img_index) { img_index[r * 2] = (int32_t)(b * N + i); img_index[r * 2 + 1] = (int32_t)(b * N + j); }
  }
}

int scan_steps(const int64_t* ids, int64_t B, int L, int N, int64_t cls_id, int64_t sep_id, int32_t* starts, int32_t* lens, int32_t* meta,
               cudaStream_t st) {
  MSQ_REQUIRE(N >= 2 && N <= EX_MAXN && L >= 2 * N, "expand: N=%d L=%d out of range", N, L);
  if (B == 0) return MSQ_OK;
  MSQ_CUDA(cudaMemsetAsync(meta, 0, 2 * sizeof(int32_t), st));
  MSQ_CUDA(launch_k(scan_steps_kernel, dim3((unsigned)B), dim3(128), 0, st, ids, B, L, N, cls_id, sep_id, starts, lens, meta));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

int expand_pairs(const int64_t* ids, int64_t B, int L, int N, int Lt, int64_t cls_id, int64_t pad_id, const int32_t* starts,
                 const int32_t* lens, int64_t* out_ids, int64_t* out_mask, int64_t* out_tt, int64_t* out_sep, int32_t* img_index,
                 cudaStream_t st) {
  if (B == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(expand_pairs_kernel, dim3((unsigned)(B * N * (N - 1))), dim3(128), 0, st, ids, L, N, Lt, cls_id, pad_id, starts, lens,
                    out_ids, out_mask, out_tt, out_sep, img_index));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

}  // namespace msq
