// C-ABI entry points (include/msq_b200.h) and the host-side orchestration of the step-ordering path:
// weight registry + packing, workspace arena, encoder stacks (CLIP ViT pair tower, joint BERT,
// text-only BERT), BERSON pooling, paragraph encoder, decode pre-projections and the beam-search launch.
// All device work is enqueued on the caller's stream; one micro-batch of manuals is encoded at a time.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/msq_b200.h"
#include "model.cuh"

namespace msq {

static thread_local char g_err[2048] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MSQ_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v;
}

// ---- per-launch event profile (bench.py roofline): events bracket each tcgen05 GEMM launch on its stream
struct Profile {
  bool on = false;
  std::vector<cudaEvent_t> ev;   // pairs: begin, end
  std::vector<double> flops;
  size_t used = 0;               // pairs recorded
  double acc_ms = 0, acc_flops = 0;
  int64_t acc_launches = 0;
};
static Profile g_prof;
constexpr size_t PROF_MAX_PAIRS = 8192;

void profile_mark(cudaStream_t st, bool end, double flops) {
  if (!g_prof.on) return;
  if (!end) {
    if (g_prof.used >= PROF_MAX_PAIRS) return;
    if (g_prof.ev.size() < 2 * (g_prof.used + 1)) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      g_prof.ev.push_back(a);
      g_prof.ev.push_back(b);
      g_prof.flops.push_back(0.0);
    }
    cudaEventRecord(g_prof.ev[2 * g_prof.used], st);
  } else {
    if (g_prof.used >= PROF_MAX_PAIRS) return;
    cudaEventRecord(g_prof.ev[2 * g_prof.used + 1], st);
    g_prof.flops[g_prof.used] = flops;
    ++g_prof.used;
  }
}


}  // namespace msq

using namespace msq;


namespace msq {

static int get_raw(msq_model* m, const std::string& name, int64_t numel, const float** out, std::string* missing) {
  auto it = m->raw.find(name);
  if (it == m->raw.end()) {
    if (missing) { *missing += name + " "; *out = nullptr; return MSQ_OK; }
    set_error("missing weight %s", name.c_str());
    return MSQ_ERR_WEIGHT;
  }
  if (numel >= 0 && it->second.second != numel) {
    set_error("weight %s has %lld elements, expected %lld", name.c_str(), (long long)it->second.second, (long long)numel);
    return MSQ_ERR_WEIGHT;
  }
  *out = it->second.first;
  return MSQ_OK;
}

// Packed copies are allocated once.  msq_model_pack runs again after every optimizer step (train.cu): the second and
// later runs issue the same sequence of requests and are handed the same buffers back (m->repack_cursor).
template <typename T> static int dev_alloc(msq_model* m, size_t n, T** out) {
  if (m->repacking) {
    MSQ_REQUIRE(m->repack_cursor < m->owned.size() && m->owned_bytes[m->repack_cursor] == n * sizeof(T),
                "repack: allocation sequence changed");
    *out = reinterpret_cast<T*>(m->owned[m->repack_cursor++]);
    return MSQ_OK;
  }
  void* p = nullptr;
  MSQ_CUDA(cudaMalloc(&p, n * sizeof(T)));
  m->owned.push_back(p);
  m->owned_bytes.push_back(n * sizeof(T));
  *out = reinterpret_cast<T*>(p);
  return MSQ_OK;
}

// build a Lin from an fp32 [N,K] weight already on device (optionally padded to Kp columns)
static int make_lin(msq_model* m, const float* w, const float* b, int N, int K, int Kp, bool want16, Lin* out,
                    cudaStream_t st) {
  out->N = N; out->K = Kp; out->ld = Kp; out->b = b;
  if (Kp != K) {
    float* wp;
    MSQ_TRY(dev_alloc(m, (size_t)N * Kp, &wp));
    MSQ_TRY(pack_pad<float>(w, N, K, Kp, wp, st));
    out->w32 = wp;
  } else {
    out->w32 = w;
  }
  if (want16 && m->cfg.precise == 2) {   // bf16x3 mode: rows [hi(Kp) | lo(Kp)]
    bf16s* w16;
    MSQ_TRY(dev_alloc(m, (size_t)N * Kp, &w16));
    MSQ_TRY(pack_pad<bf16s>(w, N, K, Kp, w16, st));
    out->w16 = reinterpret_cast<const bf16*>(w16);
  } else if (want16) {
    bf16* w16;
    MSQ_TRY(dev_alloc(m, (size_t)N * Kp, &w16));
    MSQ_TRY(pack_pad<bf16>(w, N, K, Kp, w16, st));
    out->w16 = w16;
  }
  return MSQ_OK;
}

__global__ void concat_rows_kernel(const float* a, const float* b, const float* c, int64_t na, int64_t nb, int64_t nc,
                                   float* out) {
  pdl_sync();
  const int64_t total = na + nb + nc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = i < na ? a[i] : (i < na + nb ? b[i - na] : c[i - na - nb]);
}
static int concat3(msq_model* m, const float* a, const float* b, const float* c, int64_t na, int64_t nb, int64_t nc,
                   const float** out, cudaStream_t st) {
  float* p;
  MSQ_TRY(dev_alloc(m, (size_t)(na + nb + nc), &p));
  MSQ_CUDA(launch_k(concat_rows_kernel, dim3(256), dim3(256), 0, st, a, b, c, na, nb, nc, p));
  MSQ_LAUNCH_CHECK();
  *out = p;
  return MSQ_OK;
}

// LSTM / pw_k repacks
__global__ void pack_lstm_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                                 const float* __restrict__ b_hh, int H, float* __restrict__ wih_perm,
                                 float* __restrict__ bias_perm, float* __restrict__ whh_t, float* __restrict__ whh_perm) {
  pdl_sync();
  const int64_t total = (int64_t)4 * H * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / H), k = (int)(i % H);  // row = g*H + u in torch's (i,f,g,o) order
    const int g = row / H, u = row % H, col = 4 * u + g;
    wih_perm[(int64_t)col * H + k] = w_ih[i];
    whh_t[(int64_t)k * 4 * H + col] = w_hh[i];
    whh_perm[(int64_t)col * H + k] = w_hh[i];
    if (k == 0) bias_perm[col] = b_ih[row] + b_hh[row];
  }
}
__global__ void transpose_kernel(const float* __restrict__ src, int rows, int cols, float* __restrict__ dst) {
  pdl_sync();
  const int64_t total = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    dst[(int64_t)c * rows + r] = src[i];
  }
}
// pw_k.weight [H, 4*(H+2)] -> [4H, Kp]: row blk*H + o, col c  <- w[o, blk*(H+2) + c]
__global__ void pack_pwk_kernel(const float* __restrict__ w, int H, int Kp, float* __restrict__ dst) {
  pdl_sync();
  const int D = H + 2;
  const int64_t total = (int64_t)4 * H * Kp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(i / Kp), c = (int)(i % Kp), blk = row / H, o = row % H;
    dst[i] = c < D ? w[(int64_t)o * 4 * D + blk * D + c] : 0.f;
  }
}
__global__ void mask_add_kernel(const int64_t* __restrict__ mask, int64_t n, float* __restrict__ out) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (1.0f - (float)mask[i]) * -10000.0f;
}
// R0[b,i,j,:] = [cls_mat ; softmax(score_mat) ; 0]  (rela_encode, modeling_bert.py:919-925)
__global__ void build_r0_kernel(const float* __restrict__ cls_mat, const float* __restrict__ score_mat, int64_t cells, int H,
                                int Kp, float* __restrict__ r0) {
  pdl_sync();
  const int64_t cell = blockIdx.x;
  if (cell >= cells) return;
  const float z0 = score_mat[cell * 2], z1 = score_mat[cell * 2 + 1];
  const float mx = fmaxf(z0, z1), e0 = expf(z0 - mx), e1 = expf(z1 - mx);
  for (int d = threadIdx.x; d < Kp; d += blockDim.x)
    r0[cell * Kp + d] = d < H ? cls_mat[cell * H + d] : (d == H ? e0 / (e0 + e1) : (d == H + 1 ? e1 / (e0 + e1) : 0.f));
}
// sents_ext[b, n, :] = sents[b, n, :] for n < N, zeros for n == N
__global__ void sents_ext_kernel(const float* __restrict__ sents, int64_t B, int N, int H, float* __restrict__ out) {
  pdl_sync();
  const int64_t total = B * (N + 1) * (int64_t)H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % H);
    const int64_t row = i / H, b = row / (N + 1);
    const int n = (int)(row % (N + 1));
    out[i] = n < N ? sents[(b * N + n) * H + d] : 0.f;
  }
}
__global__ void f32_to_bf16_kernel(const float* __restrict__ s, bf16* __restrict__ d, int64_t n) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    d[i] = __float2bfloat16_rn(s[i]);
}

// loss = mean_b nll_b / (N - 1 + 1e-20) + lam * mean_b (sum_p -log softmax(cls_score_p)[label_p]) / (P + 1e-20)
// (modeling_bert.py:1140-1174).  rel6 rows hold the pairwise_relationship logits in columns 0..1.
__global__ void training_loss_kernel(const float* __restrict__ nll, const float* __restrict__ rel6, const int64_t* __restrict__ labels,
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

// K of a convolution GEMM, padded so that the tcgen05 path (64-wide K slabs) applies whenever K >= 64
// padded contraction length of a convolution GEMM; split (bf16x3) operands only run on the tcgen05 kernel: whole 64-column slabs
static int rn_kpad(int K, bool split = false) { return (K < 64 && !split) ? (K + 15) / 16 * 16 : (K + 63) / 64 * 64; }

// Fold a LayerNorm into the linear layer that consumes it (one block per output feature n):
//   wf[n,k] = bf16(gamma_k w[n,k]);  svec[n] = sum_k float(wf[n,k]);  bf[n] = b[n] + sum_k beta_k w[n,k]
__global__ void __launch_bounds__(256) fold_ln_kernel(const float* __restrict__ w, int ldw, const float* __restrict__ b,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta, int K,
                                                      bf16* __restrict__ wf, float* __restrict__ svec, float* __restrict__ bf) {
  pdl_sync();
  __shared__ float sh[2][8];
  const int n = blockIdx.x;
  float s = 0.f, t = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float wv = w[(int64_t)n * ldw + k];
    const bf16 r = __float2bfloat16_rn(gamma[k] * wv);
    wf[(int64_t)n * K + k] = r;
    s += __bfloat162float(r);
    t = fmaf(beta[k], wv, t);
  }
  s = warp_sum(s); t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = t; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int i = 0; i < 8; ++i) { a += sh[0][i]; c += sh[1][i]; }
    svec[n] = a;
    bf[n] = (b ? b[n] : 0.f) + c;
  }
}
// split-bf16 twin (bf16x3 mode): wf row = [hi(K) | lo(K)] of gamma_k w[n,k]; svec[n] = sum_k (hi + lo)
__global__ void __launch_bounds__(256) fold_ln_split_kernel(const float* __restrict__ w, int ldw, const float* __restrict__ b,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta, int K,
                                                            bf16* __restrict__ wf, float* __restrict__ svec, float* __restrict__ bf) {
  pdl_sync();
  __shared__ float sh[2][8];
  const int n = blockIdx.x;
  float s = 0.f, t = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float wv = w[(int64_t)n * ldw + k];
    const float gw = gamma[k] * wv;
    const bf16 hi = __float2bfloat16_rn(gw);
    const bf16 lo = __float2bfloat16_rn(gw - __bfloat162float(hi));
    wf[(int64_t)n * 2 * K + k] = hi;
    wf[(int64_t)n * 2 * K + K + k] = lo;
    s += __bfloat162float(hi) + __bfloat162float(lo);
    t = fmaf(beta[k], wv, t);
  }
  s = warp_sum(s); t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = t; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int i = 0; i < 8; ++i) { a += sh[0][i]; c += sh[1][i]; }
    svec[n] = a;
    bf[n] = (b ? b[n] : 0.f) + c;
  }
}
static int make_folded(msq_model* m, const Lin& src, const LNp& ln, LinF* out, cudaStream_t st) {
  bf16* wf; float *sv, *bf;
  const bool split = m->cfg.precise == 2;
  MSQ_TRY(dev_alloc(m, (size_t)src.N * src.K * (split ? 2 : 1), &wf));
  MSQ_TRY(dev_alloc(m, (size_t)src.N, &sv));
  MSQ_TRY(dev_alloc(m, (size_t)src.N, &bf));
  if (split) MSQ_CUDA(launch_k(fold_ln_split_kernel, dim3(src.N), dim3(256), 0, st, src.w32, src.ld, src.b, ln.g, ln.b, src.K, wf, sv, bf));
  else MSQ_CUDA(launch_k(fold_ln_kernel, dim3(src.N), dim3(256), 0, st, src.w32, src.ld, src.b, ln.g, ln.b, src.K, wf, sv, bf));
  MSQ_LAUNCH_CHECK();
  out->lin = src; out->lin.w32 = nullptr; out->lin.w16 = wf; out->lin.b = bf; out->lin.ld = src.K; out->svec = sv;
  return MSQ_OK;
}

// Rows per slab of a residual-GEMM -> LayerNorm pair.  The pre-LN sum goes through a small scratch buffer that is
// REUSED by every slab, so it is written and re-read in L2 (126 MB) and never travels to HBM.  0 = whole micro-batch.
static int64_t slab_rows() {
  static int64_t v = -1;
  if (v < 0) { const char* e = getenv("MSQ_SLAB_ROWS"); v = e ? atoll(e) : 0; if (v < 0) v = 0; }
  return v;
}

static bool use_tc(const msq_model* m);
bool model_use_tc(const msq_model* m) { return use_tc(m); }
static bool use_tc(const msq_model* m) {
  static int forced = -1;
  if (forced < 0) { const char* e = getenv("MSQ_FORCE_SIMT"); forced = (e && e[0] == '1') ? 1 : 0; }
  return m->cfg.precise != 1 && !forced && gemm_tc_selftest_supported();
}

// C = act(A W^T + b) + resid, A in T, output in TO
template <typename T, typename TO>
static int run_gemm(const msq_model* m, const T* A, int lda, const Lin& w, const float* resid, int ldr, TO* C, int ldc,
                    int64_t M, int act, cudaStream_t st, bool with_bias = true) {
  GemmArgs g;
  g.A = A; g.bias = with_bias ? w.b : nullptr; g.resid = resid; g.C = C; g.C2 = nullptr;
  g.M = M; g.N = w.N; g.K = w.K; g.lda = lda; g.ldw = w.ld; g.ldc = ldc; g.ldr = ldr; g.act = act;
  if constexpr (same_type<T, float>::value) {
    g.W = w.w32;
    return gemm_simt<float, TO>(g, st);
  } else if constexpr (is_split<T>::value) {
    // bf16x3 mode: tensor cores only (split operands have no FFMA twin; the fp32 parity mode is the CUDA-core path)
    g.W = w.w16; g.split = 1;
    MSQ_REQUIRE(w.w16 != nullptr, "split-bf16 weight copy missing");
    MSQ_REQUIRE(use_tc(m) && g.K % 64 == 0 && g.N % 8 == 0 && ldc % 8 == 0, "bf16x3 mode: GEMM N=%d K=%d not supported by the tcgen05 kernel", g.N, g.K);
    return gemm_tc<TO>(g, st);
  } else {
    g.W = w.w16;
    MSQ_REQUIRE(w.w16 != nullptr, "bf16 weight copy missing");
    // narrow convolutions of the ResNet stem (K = 27 -> 32, N = 32) stay on the FFMA kernel
    if (use_tc(m) && g.K % 64 == 0 && g.N % 8 == 0 && ldc % 8 == 0) return gemm_tc<TO>(g, st);
    return gemm_simt<bf16, TO>(g, st);
  }
}

// ---- deferred LayerNorm on the tensor-core path: a stream is kept RAW (fp32 Y + bf16 copy + per-row
// partial sums); the LayerNorm that follows it is applied inside the consumers' epilogues and never materialised.
static bool ln_fold_enabled(const msq_model* m) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MSQ_LNFOLD"); v = (e && e[0] == '0') ? 0 : 1; }
  return v && use_tc(m);
}
struct LnPending { const float* stats = nullptr; int sp = 0; LNp ln; float eps = 0.f; int dim = 0; };

// C = act(LN_pending(Y) W^T + b) computed from the bf16 copy of the raw stream
// (split: Yt and the folded weight are split-bf16 rows, bf16x3 mode)
template <typename TO>
static int run_gemm_fold(const void* Yt, int lda, const LinF& w, const LnPending& ln, TO* C, int ldc, int64_t M, int act, cudaStream_t st,
                         bool split = false) {
  GemmArgs g;
  g.split = split ? 1 : 0;
  g.A = Yt; g.W = w.lin.w16; g.bias = w.lin.b; g.resid = nullptr; g.C = C; g.C2 = nullptr;
  g.M = M; g.N = w.lin.N; g.K = w.lin.K; g.lda = lda; g.ldw = w.lin.ld; g.ldc = ldc; g.ldr = 0; g.act = act;
  g.mode = EPI_LNFOLD; g.svec = w.svec; g.stats_in = ln.stats; g.sp_in = ln.sp; g.ln_inv_dim = 1.0f / (float)ln.dim; g.ln_eps = ln.eps;
  return gemm_tc<TO>(g, st);
}
// Y <- A W^T + b + (ln ? LN_pending(Y) : Y) in place, plus its bf16 copy Yt and the partial row sums of the new Y
static int run_gemm_resln(const void* A, int lda, const Lin& w, float* Y, void* Yt, const LnPending* ln, float* stats_out, int64_t M,
                          cudaStream_t st, bool split = false) {
  GemmArgs g;
  g.split = split ? 1 : 0;
  g.A = A; g.W = w.w16; g.bias = w.b; g.resid = Y; g.C = Y; g.C2 = nullptr;
  g.M = M; g.N = w.N; g.K = w.K; g.lda = lda; g.ldw = w.ld; g.ldc = w.N; g.ldr = w.N; g.act = ACT_NONE;
  g.mode = EPI_RESLN; g.C2bf = Yt; g.stats_out = stats_out;
  if (ln) {
    g.svec = ln->ln.g; g.beta = ln->ln.b; g.stats_in = ln->stats; g.sp_in = ln->sp; g.ln_inv_dim = 1.0f / (float)ln->dim; g.ln_eps = ln->eps;
  }
  return gemm_tc<float>(g, st);
}

}  // namespace msq

// =====================================================================================================
// model lifetime
// =====================================================================================================

extern "C" const char* msq_last_error(void) { return g_err; }
extern "C" int msq_version(void) { return 100; }
extern "C" int64_t msq_launch_count(void) { return g_launches.load(); }
extern "C" int msq_tc_available(void) { return gemm_tc_selftest_supported(); }

extern "C" int msq_profile_enable(int32_t on) {
  g_prof.on = on != 0;
  g_prof.used = 0;
  g_prof.acc_ms = g_prof.acc_flops = 0;
  g_prof.acc_launches = 0;
  return MSQ_OK;
}
// Synchronises the device, folds the recorded event pairs into the totals and returns them.
extern "C" int msq_profile_read(double* total_ms, double* total_flops, int64_t* launches) {
  MSQ_CUDA(cudaDeviceSynchronize());
  for (size_t i = 0; i < g_prof.used; ++i) {
    float ms = 0.f;
    MSQ_CUDA(cudaEventElapsedTime(&ms, g_prof.ev[2 * i], g_prof.ev[2 * i + 1]));
    g_prof.acc_ms += ms;
    g_prof.acc_flops += g_prof.flops[i];
    ++g_prof.acc_launches;
  }
  g_prof.used = 0;
  if (total_ms) *total_ms = g_prof.acc_ms;
  if (total_flops) *total_flops = g_prof.acc_flops;
  if (launches) *launches = g_prof.acc_launches;
  return MSQ_OK;
}

extern "C" int msq_model_create(const msq_config* cfg, msq_model** out) {
  MSQ_REQUIRE(cfg && out, "null argument");
  MSQ_REQUIRE(cfg->hidden % 128 == 0 && cfg->hidden <= 1024, "hidden=%d must be a multiple of 128, <= 1024", cfg->hidden);
  MSQ_REQUIRE(cfg->heads * 64 == cfg->hidden, "head dim must be 64 (hidden=%d heads=%d)", cfg->hidden, cfg->heads);
  MSQ_REQUIRE(cfg->inter % 16 == 0 && cfg->layers >= 0, "bad inter/layers");
  if (cfg->rn_width) {
    MSQ_REQUIRE(cfg->rn_width % 16 == 0 && cfg->rn_embed % 32 == 0 && cfg->vit_width == 2 * cfg->rn_embed && cfg->vit_layers == 0 &&
                    cfg->vit_patch == 32 && cfg->vit_res % 32 == 0,
                "ResNet tower: need rn_width %% 16 == 0, vit_width == 2*rn_embed, vit_layers == 0, vit_patch == 32");
    for (int i = 0; i < 4; ++i) MSQ_REQUIRE(cfg->rn_blocks[i] >= 1, "rn_blocks[%d]=%d", i, cfg->rn_blocks[i]);
  } else if (cfg->vit_width) {
    MSQ_REQUIRE(cfg->vit_width % 128 == 0 && cfg->vit_width <= 1024, "vit_width=%d unsupported", cfg->vit_width);
    MSQ_REQUIRE(cfg->vit_res % cfg->vit_patch == 0 && cfg->vit_patch % 4 == 0, "vit patch/res");
  }
  MSQ_REQUIRE(cfg->para_heads >= 1 && cfg->hidden % cfg->para_heads == 0 && cfg->para_ff % 16 == 0, "bad paragraph config");
  MSQ_REQUIRE(cfg->precise >= 0 && cfg->precise <= 2, "precise=%d: 0 bf16, 1 fp32, 2 bf16x3", cfg->precise);
  MSQ_REQUIRE(!(cfg->precise == 2 && cfg->rn_width && cfg->rn_width % 64), "the bf16x3 mode needs a ModifiedResNet width that is a multiple of 64 (tcgen05 tiles)");
  MSQ_REQUIRE(!(cfg->precise == 2 && (cfg->inter % 64 || cfg->vit_width % 64)), "the bf16x3 mode needs inter and vit_width to be multiples of 64");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return MSQ_ERR_CUDA;
  }
  msq_model* m = new msq_model();
  m->cfg = *cfg;
  MSQ_CUDA(cudaGetDevice(&m->device));
  *out = m;
  return MSQ_OK;
}

extern "C" void msq_model_destroy(msq_model* m) {
  DevGuard dev_guard__(m);
  delete m;
}

extern "C" int msq_model_set_weight(msq_model* m, const char* name, const float* data_dev, int64_t numel, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && name && data_dev && numel > 0, "bad argument");
  MSQ_REQUIRE(!m->packed, "model already packed");
  float* p = nullptr;
  MSQ_CUDA(cudaMalloc(&p, numel * sizeof(float)));
  MSQ_CUDA(cudaMemcpyAsync(p, data_dev, numel * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  auto it = m->raw.find(name);
  if (it != m->raw.end()) { cudaFree(it->second.first); m->raw.erase(it); }
  m->raw[name] = {p, numel};
  return MSQ_OK;
}

extern "C" int msq_model_pack(msq_model* m, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && !m->packed, "bad state");
  cudaStream_t st = (cudaStream_t)stream;
  const msq_config& c = m->cfg;
  const int H = c.hidden;
  const bool w16 = c.precise != 1;
  const bool fold = c.precise != 1;   // deferred-LayerNorm copies: the tensor-core paths (bf16 and split-bf16)
  std::string miss;
  const std::string P = m->prefix_inner;
  auto W = [&](const std::string& n, int64_t numel) {
    const float* p = nullptr;
    if (get_raw(m, n, numel, &p, &miss) != MSQ_OK) { miss += n + "(wrong size) "; p = nullptr; }
    return p;
  };
  // first make sure everything is there (collect all missing names)
  std::vector<std::string> need;
  auto lin_names = [&](const std::string& n) { need.push_back(n + ".weight"); need.push_back(n + ".bias"); };
  // three weight groups, each optional as a whole: inner BERT stack, CLIP visual tower, BERSON heads
  m->has_bert = m->raw.count(P + "embeddings.word_embeddings.weight") != 0;
  m->has_vit = c.vit_width != 0 && m->raw.count(P + "encoder.visual_model.visual.conv1.weight") != 0;
  m->has_heads = m->raw.count("key_linear.weight") != 0;
  MSQ_REQUIRE(m->has_bert || m->has_vit || m->has_heads, "no known weights registered");
  if (m->has_bert)
  for (const char* e : {"embeddings.word_embeddings.weight", "embeddings.position_embeddings.weight",
                        "embeddings.token_type_embeddings.weight", "embeddings.LayerNorm.weight", "embeddings.LayerNorm.bias"})
    need.push_back(P + e);
  for (int l = 0; m->has_bert && l < c.layers; ++l) {
    const std::string b = P + "encoder.layer." + std::to_string(l) + ".";
    for (const char* e : {"attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense",
                          "attention.output.LayerNorm", "intermediate.dense", "output.dense", "output.LayerNorm"})
      lin_names(b + e);
  }
  auto bn_names = [&](const std::string& n) { for (const char* e : {".weight", ".bias", ".running_mean", ".running_var"}) need.push_back(n + e); };
  if (m->has_vit && c.rn_width) {
    const std::string v = P + "encoder.visual_model.visual.";
    for (int i = 1; i <= 3; ++i) { need.push_back(v + "conv" + std::to_string(i) + ".weight"); bn_names(v + "bn" + std::to_string(i)); }
    for (int s = 0; s < 4; ++s)
      for (int b = 0; b < c.rn_blocks[s]; ++b) {
        const std::string k = v + "layer" + std::to_string(s + 1) + "." + std::to_string(b) + ".";
        for (int i = 1; i <= 3; ++i) { need.push_back(k + "conv" + std::to_string(i) + ".weight"); bn_names(k + "bn" + std::to_string(i)); }
        if (b == 0) { need.push_back(k + "downsample.0.weight"); bn_names(k + "downsample.1"); }
      }
    need.push_back(v + "attnpool.positional_embedding");
    for (const char* e : {"attnpool.q_proj", "attnpool.k_proj", "attnpool.v_proj", "attnpool.c_proj"}) lin_names(v + e);
    if (m->has_bert) {
      need.push_back(P + "encoder.visual_pos.x_position_embedding.weight");
      need.push_back(P + "encoder.visual_pos.y_position_embedding.weight");
      need.push_back(P + "encoder.visual_token_type.token_type_embedding.weight");
      lin_names(P + "encoder.visn_fc.visn_fc"); lin_names(P + "encoder.visn_fc.visn_layer_norm");
    }
  } else if (m->has_vit) {
    const std::string v = P + "encoder.visual_model.visual.";
    need.push_back(v + "conv1.weight"); need.push_back(v + "class_embedding"); need.push_back(v + "positional_embedding");
    lin_names(v + "ln_pre"); lin_names(v + "ln_post");
    for (int l = 0; l < c.vit_layers; ++l) {
      const std::string b = v + "transformer.resblocks." + std::to_string(l) + ".";
      need.push_back(b + "attn.in_proj_weight"); need.push_back(b + "attn.in_proj_bias");
      for (const char* e : {"attn.out_proj", "ln_1", "mlp.c_fc", "mlp.c_proj", "ln_2"}) lin_names(b + e);
    }
    if (m->has_bert) { lin_names(P + "encoder.visn_fc.visn_fc"); lin_names(P + "encoder.visn_fc.visn_layer_norm"); }
  }
  for (int l = 0; m->has_heads && l < c.para_layers; ++l) {
    const std::string b = "encoder.transformer_inter." + std::to_string(l) + ".";
    for (const char* e : {"self_attn.linear_keys", "self_attn.linear_values", "self_attn.linear_query", "self_attn.final_linear",
                          "feed_forward.w_1", "feed_forward.w_2", "feed_forward.layer_norm", "layer_norm"})
      lin_names(b + e);
  }
  if (m->has_heads) {
  lin_names("encoder.layer_norm"); lin_names("key_linear"); lin_names("query_linear"); lin_names("tanh_linear");
  for (const char* e : {"decoder.weight_ih_l0", "decoder.weight_hh_l0", "decoder.bias_ih_l0", "decoder.bias_hh_l0", "pw_k.weight",
                        "two_level_encoder.linear_in_2.weight"})
    need.push_back(e);
  for (const char* e : {"two_level_encoder.sentence_tran", "two_level_encoder.sentence_tran_2",
                        "two_level_encoder.pairwise_relationship", "two_level_encoder.h1_relationship",
                        "two_level_encoder.h2_relationship"})
    lin_names(e);
  }
  for (auto& n : need)
    if (!m->raw.count(n)) miss += n + " ";
  if (!miss.empty()) {
    set_error("state_dict incomplete, missing: %.1800s", miss.c_str());
    return MSQ_ERR_WEIGHT;
  }

  // ---- embeddings + BERT stack
  if (m->has_bert) {
  m->word = W(P + "embeddings.word_embeddings.weight", (int64_t)c.vocab * H);
  m->pos = W(P + "embeddings.position_embeddings.weight", (int64_t)c.max_pos * H);
  m->type = W(P + "embeddings.token_type_embeddings.weight", (int64_t)c.type_vocab * H);
  m->emb_ln = {W(P + "embeddings.LayerNorm.weight", H), W(P + "embeddings.LayerNorm.bias", H)};
  if (!m->word || !m->pos || !m->type) return MSQ_ERR_WEIGHT;
  m->bert.resize(c.layers);
  for (int l = 0; l < c.layers; ++l) {
    const std::string b = P + "encoder.layer." + std::to_string(l) + ".";
    BertLayerW& L = m->bert[l];
    const float *wq, *bq;
    MSQ_TRY(concat3(m, W(b + "attention.self.query.weight", (int64_t)H * H), W(b + "attention.self.key.weight", (int64_t)H * H),
                    W(b + "attention.self.value.weight", (int64_t)H * H), (int64_t)H * H, (int64_t)H * H, (int64_t)H * H, &wq, st));
    MSQ_TRY(concat3(m, W(b + "attention.self.query.bias", H), W(b + "attention.self.key.bias", H),
                    W(b + "attention.self.value.bias", H), H, H, H, &bq, st));
    MSQ_TRY(make_lin(m, wq, bq, 3 * H, H, H, w16, &L.qkv, st));
    MSQ_TRY(make_lin(m, W(b + "attention.output.dense.weight", (int64_t)H * H), W(b + "attention.output.dense.bias", H), H, H, H,
                     w16, &L.out, st));
    MSQ_TRY(make_lin(m, W(b + "intermediate.dense.weight", (int64_t)c.inter * H), W(b + "intermediate.dense.bias", c.inter),
                     c.inter, H, H, w16, &L.up, st));
    MSQ_TRY(make_lin(m, W(b + "output.dense.weight", (int64_t)H * c.inter), W(b + "output.dense.bias", H), H, c.inter, c.inter,
                     w16, &L.down, st));
    L.ln1 = {W(b + "attention.output.LayerNorm.weight", H), W(b + "attention.output.LayerNorm.bias", H)};
    L.ln2 = {W(b + "output.LayerNorm.weight", H), W(b + "output.LayerNorm.bias", H)};
    if (fold && miss.empty()) {   // deferred-LayerNorm copies (tensor-core path only)
      MSQ_TRY(make_folded(m, L.up, L.ln1, &L.up_f, st));
      if (l >= 1) MSQ_TRY(make_folded(m, L.qkv, m->bert[l - 1].ln2, &L.qkv_f, st));
      m->folded = true;
    }
  }
  if (m->raw.count(P + "pooler.dense.weight")) {
    MSQ_TRY(make_lin(m, W(P + "pooler.dense.weight", (int64_t)H * H), W(P + "pooler.dense.bias", H), H, H, H, false, &m->pooler, st));
    m->has_pooler = true;
  }
  }
  // ---- ModifiedResNet tower: BatchNorm folded into [Cout, kh*kw*Cin] GEMM weights
  if (m->has_vit && c.rn_width) {
    const std::string v = P + "encoder.visual_model.visual.";
    const int w = c.rn_width, Cf = 32 * w, E = c.rn_embed, F = 2 * E, g = c.vit_res / 32, g2 = g * g;
    auto conv = [&](const std::string& cw, const std::string& bn, int cout, int cin, int k, Lin* out) -> int {
      const int K = k * k * cin;
      const bool split = c.precise == 2;
      MSQ_REQUIRE(k != 1 || rn_kpad(K, split) == K, "ResNet tower: 1x1 convolution with %d input channels is not GEMM-aligned", cin);
      float *wf, *bf;
      MSQ_TRY(dev_alloc(m, (size_t)cout * K, &wf));
      MSQ_TRY(dev_alloc(m, (size_t)cout, &bf));
      const float *pw = W(cw, (int64_t)cout * K), *pg = W(bn + ".weight", cout), *pb = W(bn + ".bias", cout),
                  *pm = W(bn + ".running_mean", cout), *pv = W(bn + ".running_var", cout);
      if (!pw || !pg || !pb || !pm || !pv) return MSQ_ERR_WEIGHT;
      MSQ_TRY(rn_fold(pw, pg, pb, pm, pv, cout, cin, k, wf, bf, st));
      return make_lin(m, wf, bf, cout, K, rn_kpad(K, split), w16, out, st);
    };
    MSQ_TRY(conv(v + "conv1.weight", v + "bn1", w / 2, 3, 3, &m->rn_stem[0]));
    MSQ_TRY(conv(v + "conv2.weight", v + "bn2", w / 2, w / 2, 3, &m->rn_stem[1]));
    MSQ_TRY(conv(v + "conv3.weight", v + "bn3", w, w / 2, 3, &m->rn_stem[2]));
    int cin = w;
    for (int s = 0; s < 4; ++s)
      for (int b = 0; b < c.rn_blocks[s]; ++b) {
        const std::string k = v + "layer" + std::to_string(s + 1) + "." + std::to_string(b) + ".";
        RnBlockW B;
        B.planes = w << s; B.cin = cin; B.stride = (b == 0 && s > 0) ? 2 : 1; B.has_ds = b == 0;
        MSQ_TRY(conv(k + "conv1.weight", k + "bn1", B.planes, cin, 1, &B.c1));
        MSQ_TRY(conv(k + "conv2.weight", k + "bn2", B.planes, B.planes, 3, &B.c2));
        MSQ_TRY(conv(k + "conv3.weight", k + "bn3", 4 * B.planes, B.planes, 1, &B.c3));
        if (B.has_ds) MSQ_TRY(conv(k + "downsample.0.weight", k + "downsample.1", 4 * B.planes, cin, 1, &B.ds));
        cin = 4 * B.planes;
        m->rn_blocks.push_back(B);
      }
    const std::string a = v + "attnpool.";
    m->rn_pos = W(a + "positional_embedding", (int64_t)(g2 + 1) * Cf);
    const float *wq, *bq;
    MSQ_TRY(concat3(m, W(a + "q_proj.weight", (int64_t)Cf * Cf), W(a + "k_proj.weight", (int64_t)Cf * Cf), W(a + "v_proj.weight", (int64_t)Cf * Cf),
                    (int64_t)Cf * Cf, (int64_t)Cf * Cf, (int64_t)Cf * Cf, &wq, st));
    MSQ_TRY(concat3(m, W(a + "q_proj.bias", Cf), W(a + "k_proj.bias", Cf), W(a + "v_proj.bias", Cf), Cf, Cf, Cf, &bq, st));
    MSQ_TRY(make_lin(m, wq, bq, 3 * Cf, Cf, Cf, w16, &m->rn_qkv, st));
    MSQ_TRY(make_lin(m, W(a + "c_proj.weight", (int64_t)E * Cf), W(a + "c_proj.bias", E), E, Cf, Cf, w16, &m->rn_cproj, st));
    if (m->has_bert) {
      auto tab = [&](const std::string& n) -> const float* {   // embedding tables: only the first rows are addressed
        auto it = m->raw.find(n);
        return it == m->raw.end() ? nullptr : it->second.first;
      };
      const float* xe = tab(P + "encoder.visual_pos.x_position_embedding.weight");
      const float* ye = tab(P + "encoder.visual_pos.y_position_embedding.weight");
      const float* te = tab(P + "encoder.visual_token_type.token_type_embedding.weight");
      MSQ_REQUIRE(m->raw[P + "encoder.visual_pos.x_position_embedding.weight"].second >= (int64_t)g * F &&
                      m->raw[P + "encoder.visual_pos.y_position_embedding.weight"].second >= (int64_t)g * F &&
                      m->raw[P + "encoder.visual_token_type.token_type_embedding.weight"].second >= (int64_t)2 * F,
                  "visual_pos / visual_token_type tables are smaller than the %dx%d grid needs", g, g);
      float* pa;
      MSQ_TRY(dev_alloc(m, (size_t)(1 + 2 * g2) * F, &pa));
      MSQ_TRY(rn_posadd(xe, ye, te, g, F, pa, st));
      m->rn_posadd = pa;
      MSQ_TRY(make_lin(m, W(P + "encoder.visn_fc.visn_fc.weight", (int64_t)H * F), W(P + "encoder.visn_fc.visn_fc.bias", H), H, F, F,
                       w16, &m->visn_fc, st));
      m->visn_ln = {W(P + "encoder.visn_fc.visn_layer_norm.weight", H), W(P + "encoder.visn_fc.visn_layer_norm.bias", H)};
    }
  } else
  // ---- ViT tower
  if (m->has_vit) {
    const int Wd = c.vit_width, g = c.vit_res / c.vit_patch, Kc = 3 * c.vit_patch * c.vit_patch;
    const std::string v = P + "encoder.visual_model.visual.";
    MSQ_TRY(make_lin(m, W(v + "conv1.weight", (int64_t)Wd * Kc), nullptr, Wd, Kc, Kc, w16, &m->conv1, st));
    m->vit_cls = W(v + "class_embedding", Wd);
    m->vit_pos = W(v + "positional_embedding", (int64_t)(g * g + 1) * Wd);
    m->ln_pre = {W(v + "ln_pre.weight", Wd), W(v + "ln_pre.bias", Wd)};
    m->ln_post = {W(v + "ln_post.weight", Wd), W(v + "ln_post.bias", Wd)};
    m->vit.resize(c.vit_layers);
    for (int l = 0; l < c.vit_layers; ++l) {
      const std::string b = v + "transformer.resblocks." + std::to_string(l) + ".";
      VitLayerW& L = m->vit[l];
      MSQ_TRY(make_lin(m, W(b + "attn.in_proj_weight", (int64_t)3 * Wd * Wd), W(b + "attn.in_proj_bias", 3 * Wd), 3 * Wd, Wd, Wd,
                       w16, &L.qkv, st));
      MSQ_TRY(make_lin(m, W(b + "attn.out_proj.weight", (int64_t)Wd * Wd), W(b + "attn.out_proj.bias", Wd), Wd, Wd, Wd, w16, &L.out, st));
      MSQ_TRY(make_lin(m, W(b + "mlp.c_fc.weight", (int64_t)4 * Wd * Wd), W(b + "mlp.c_fc.bias", 4 * Wd), 4 * Wd, Wd, Wd, w16, &L.fc, st));
      MSQ_TRY(make_lin(m, W(b + "mlp.c_proj.weight", (int64_t)4 * Wd * Wd), W(b + "mlp.c_proj.bias", Wd), Wd, 4 * Wd, 4 * Wd, w16,
                       &L.proj, st));
      L.ln1 = {W(b + "ln_1.weight", Wd), W(b + "ln_1.bias", Wd)};
      L.ln2 = {W(b + "ln_2.weight", Wd), W(b + "ln_2.bias", Wd)};
      if (fold && miss.empty()) {
        MSQ_TRY(make_folded(m, L.qkv, L.ln1, &L.qkv_f, st));
        MSQ_TRY(make_folded(m, L.fc, L.ln2, &L.fc_f, st));
        m->folded = true;
      }
    }
    if (m->has_bert) {
      MSQ_TRY(make_lin(m, W(P + "encoder.visn_fc.visn_fc.weight", (int64_t)H * Wd), W(P + "encoder.visn_fc.visn_fc.bias", H), H, Wd, Wd,
                       w16, &m->visn_fc, st));
      m->visn_ln = {W(P + "encoder.visn_fc.visn_layer_norm.weight", H), W(P + "encoder.visn_fc.visn_layer_norm.bias", H)};
      if (fold && miss.empty()) MSQ_TRY(make_folded(m, m->visn_fc, m->ln_post, &m->visn_fc_f, st));
    }
  }
  // ---- BERSON heads
  if (m->has_heads) {
  const std::string T = "two_level_encoder.";
  MSQ_TRY(make_lin(m, W(T + "sentence_tran.weight", (int64_t)H * H), W(T + "sentence_tran.bias", H), H, H, H, w16, &m->sent_tran, st));
  m->w2 = W(T + "sentence_tran_2.weight", H);
  m->b2 = W(T + "sentence_tran_2.bias", 1);
  m->w_in2 = W(T + "linear_in_2.weight", H);
  MSQ_TRY(concat3(m, W(T + "pairwise_relationship.weight", 2 * H), W(T + "h1_relationship.weight", 2 * H),
                  W(T + "h2_relationship.weight", 2 * H), 2 * H, 2 * H, 2 * H, &m->w_rel, st));
  MSQ_TRY(concat3(m, W(T + "pairwise_relationship.bias", 2), W(T + "h1_relationship.bias", 2), W(T + "h2_relationship.bias", 2), 2, 2,
                  2, &m->b_rel, st));
  m->para.resize(c.para_layers);
  for (int l = 0; l < c.para_layers; ++l) {
    const std::string b = "encoder.transformer_inter." + std::to_string(l) + ".";
    ParaLayerW& L = m->para[l];
    const float *wq, *bq;
    MSQ_TRY(concat3(m, W(b + "self_attn.linear_query.weight", (int64_t)H * H), W(b + "self_attn.linear_keys.weight", (int64_t)H * H),
                    W(b + "self_attn.linear_values.weight", (int64_t)H * H), (int64_t)H * H, (int64_t)H * H, (int64_t)H * H, &wq, st));
    MSQ_TRY(concat3(m, W(b + "self_attn.linear_query.bias", H), W(b + "self_attn.linear_keys.bias", H),
                    W(b + "self_attn.linear_values.bias", H), H, H, H, &bq, st));
    MSQ_TRY(make_lin(m, wq, bq, 3 * H, H, H, false, &L.qkv, st));
    MSQ_TRY(make_lin(m, W(b + "self_attn.final_linear.weight", (int64_t)H * H), W(b + "self_attn.final_linear.bias", H), H, H, H, false,
                     &L.fin, st));
    MSQ_TRY(make_lin(m, W(b + "feed_forward.w_1.weight", (int64_t)c.para_ff * H), W(b + "feed_forward.w_1.bias", c.para_ff), c.para_ff,
                     H, H, false, &L.w1, st));
    MSQ_TRY(make_lin(m, W(b + "feed_forward.w_2.weight", (int64_t)H * c.para_ff), W(b + "feed_forward.w_2.bias", H), H, c.para_ff,
                     c.para_ff, false, &L.w2, st));
    L.ln_in = {W(b + "layer_norm.weight", H), W(b + "layer_norm.bias", H)};
    L.ln_ff = {W(b + "feed_forward.layer_norm.weight", H), W(b + "feed_forward.layer_norm.bias", H)};
  }
  m->para_ln = {W("encoder.layer_norm.weight", H), W("encoder.layer_norm.bias", H)};
  MSQ_TRY(make_lin(m, W("key_linear.weight", (int64_t)2 * H * H), W("key_linear.bias", H), H, 2 * H, 2 * H, false, &m->key_lin, st));
  // decoder repacks
  {
    float *wih_perm, *bias_perm, *whh_t, *wq_t, *wpw4, *whh_perm;
    m->Kp = ((H + 2) + 63) / 64 * 64;   // padded H+2: a multiple of 64 so that the T4 projection can run on the tcgen05 kernel
    MSQ_TRY(dev_alloc(m, (size_t)4 * H * H, &whh_perm));
    MSQ_TRY(dev_alloc(m, (size_t)4 * H * H, &wih_perm));
    MSQ_TRY(dev_alloc(m, (size_t)4 * H, &bias_perm));
    MSQ_TRY(dev_alloc(m, (size_t)4 * H * H, &whh_t));
    MSQ_TRY(dev_alloc(m, (size_t)H * H, &wq_t));
    MSQ_TRY(dev_alloc(m, (size_t)4 * H * m->Kp, &wpw4));
    MSQ_CUDA(launch_k(pack_lstm_kernel, dim3(512), dim3(256), 0, st, W("decoder.weight_ih_l0", (int64_t)4 * H * H), W("decoder.weight_hh_l0", (int64_t)4 * H * H), W("decoder.bias_ih_l0", 4 * H), W("decoder.bias_hh_l0", 4 * H), H, wih_perm, bias_perm, whh_t, whh_perm));
    MSQ_LAUNCH_CHECK();
    MSQ_CUDA(launch_k(transpose_kernel, dim3(256), dim3(256), 0, st, W("query_linear.weight", (int64_t)H * H), H, H, wq_t));
    MSQ_LAUNCH_CHECK();
    MSQ_CUDA(launch_k(pack_pwk_kernel, dim3(512), dim3(256), 0, st, W("pw_k.weight", (int64_t)H * 4 * (H + 2)), H, m->Kp, wpw4));
    MSQ_LAUNCH_CHECK();
    m->xg_lin.w32 = wih_perm; m->xg_lin.b = bias_perm; m->xg_lin.N = 4 * H; m->xg_lin.K = H; m->xg_lin.ld = H;
    m->t4_lin.w32 = wpw4; m->t4_lin.b = nullptr; m->t4_lin.N = 4 * H; m->t4_lin.K = m->Kp; m->t4_lin.ld = m->Kp;
    m->dec.whh_t = whh_t; m->dec.wq_t = wq_t; m->dec.bq = W("query_linear.bias", H);
    m->dec.wt = W("tanh_linear.weight", H);
    float bt = 0.f;
    MSQ_CUDA(cudaMemcpyAsync(&bt, W("tanh_linear.bias", 1), sizeof(float), cudaMemcpyDeviceToHost, st));
    MSQ_CUDA(cudaStreamSynchronize(st));
    m->dec.bt = bt;
    if (c.precise != 1) {
      // tensor-core decode (bf16x6): three-plane bf16 copies of the decoder's weights; products are formed from all plane
      // pairs (i, j) with i + j <= 2, i.e. to ~2^-24 relative -- fp32-grade results at tcgen05 speed
      bf16 *wih3, *wpw3, *wcat3;
      float* bcat;
      MSQ_TRY(dev_alloc(m, (size_t)4 * H * 3 * H, &wih3));
      MSQ_TRY(dev_alloc(m, (size_t)4 * H * 3 * m->Kp, &wpw3));
      MSQ_TRY(dev_alloc(m, (size_t)5 * H * 3 * H, &wcat3));
      MSQ_TRY(dev_alloc(m, (size_t)5 * H, &bcat));
      MSQ_TRY(pack_split3(wih_perm, 4 * H, H, H, H, wih3, st));
      MSQ_TRY(pack_split3(wpw4, 4 * H, m->Kp, m->Kp, m->Kp, wpw3, st));
      MSQ_TRY(pack_split3(W("query_linear.weight", (int64_t)H * H), H, H, H, H, wcat3, st));
      MSQ_TRY(pack_split3(whh_perm, 4 * H, H, H, H, wcat3 + (size_t)H * 3 * H, st));
      MSQ_CUDA(cudaMemsetAsync(bcat, 0, (size_t)5 * H * sizeof(float), st));
      MSQ_CUDA(cudaMemcpyAsync(bcat, W("query_linear.bias", H), (size_t)H * sizeof(float), cudaMemcpyDeviceToDevice, st));
      m->wih3 = wih3; m->wpw3 = wpw3; m->dec.wcat3 = wcat3; m->dec.bcat = bcat;
    }
    // raw-layout copies for the step-by-step API (msq_decode_step): pw_k padded to a multiple of 16 columns
    m->Kp4 = (4 * (H + 2) + 15) / 16 * 16;
    MSQ_TRY(make_lin(m, W("pw_k.weight", (int64_t)H * 4 * (H + 2)), nullptr, H, 4 * (H + 2), m->Kp4, false, &m->pwk_raw, st));
    MSQ_TRY(make_lin(m, W("decoder.weight_ih_l0", (int64_t)4 * H * H), W("decoder.bias_ih_l0", 4 * H), 4 * H, H, H, false, &m->wih_raw, st));
    MSQ_TRY(make_lin(m, W("decoder.weight_hh_l0", (int64_t)4 * H * H), W("decoder.bias_hh_l0", 4 * H), 4 * H, H, H, false, &m->whh_raw, st));
    MSQ_TRY(make_lin(m, W("query_linear.weight", (int64_t)H * H), W("query_linear.bias", H), H, H, H, false, &m->wq_raw, st));
  }
  }
  if (!miss.empty()) {
    set_error("state_dict incomplete, missing: %.1800s", miss.c_str());
    return MSQ_ERR_WEIGHT;
  }
  MSQ_CUDA(cudaStreamSynchronize(st));
  m->packed = true;
  return MSQ_OK;
}

// Re-derive every packed copy (fused QKV, bf16, folded LayerNorm, LSTM / pw_k repacks) from the fp32 masters after an
// optimizer step has changed them.  Same buffers, no allocation (dev_alloc replays m->owned).
int msq::model_repack(msq_model* m, cudaStream_t st) {
  MSQ_REQUIRE(m && m->packed, "repack: model not packed");
  m->packed = false; m->repacking = true; m->repack_cursor = 0;
  m->rn_blocks.clear();
  const int rc = msq_model_pack(m, (void*)st);
  m->repacking = false;
  if (rc != MSQ_OK) return rc;
  MSQ_REQUIRE(m->repack_cursor == m->owned.size(), "repack: allocation sequence changed");
  return MSQ_OK;
}

// =====================================================================================================
// encoder orchestration
// =====================================================================================================
namespace msq {

struct VitBufs {
  void* apatch; float* patch; float* xv; void* y; void* qkv; void* ctx; void* hbuf;
  // ModifiedResNet trunk scratch (per image chunk): block input / 1x1 out / 3x3 out / pooled copies, fp32 residual + shortcut
  void *rX, *rO1, *rO2, *rP1, *rP2; float *rF0, *rF1;
  // deferred LayerNorm: partial row sums of the raw stream, and (set by run_vit) the ln_post still pending on b.y
  float* st[2];
  LnPending post;
  bool post_pending;
};
constexpr int64_t IMG_CHUNK = 1024;  // images per im2col + patch-embed GEMM launch

template <typename T>
static void plan_vit(const msq_config& c, Planner& p, int64_t n_img, int64_t R, VitBufs* b) {
  const int g2 = (c.vit_res / c.vit_patch) * (c.vit_res / c.vit_patch), Lv = 1 + 2 * g2, Wd = c.vit_width;
  const int Kc = 3 * c.vit_patch * c.vit_patch;
  b->apatch = p.take<T>((size_t)min(n_img, IMG_CHUNK) * g2 * Kc);
  b->patch = p.take<float>((size_t)n_img * g2 * Wd);
  b->xv = p.take<float>((size_t)R * Lv * Wd);
  b->y = p.take<T>((size_t)R * Lv * Wd);
  b->qkv = p.take<T>((size_t)R * Lv * 3 * Wd);
  b->ctx = p.take<T>((size_t)R * Lv * Wd);
  b->hbuf = p.take<T>((size_t)R * Lv * 4 * Wd);
  const size_t sp = 2 * (size_t)ceil_div(Wd, 256);
  b->st[0] = p.take<float>((size_t)R * Lv * sp * 2);
  b->st[1] = p.take<float>((size_t)R * Lv * sp * 2);
}

// conv1 (32x32 / stride 32, no bias) as im2col + GEMM over all UNIQUE images -> b.patch [n_img*g2, W]
template <typename T>
static int run_patch_embed(msq_model* m, const float* images, int64_t n_img, VitBufs& b, cudaStream_t st, int64_t first = 0) {
  const msq_config& c = m->cfg;
  const int g = c.vit_res / c.vit_patch, g2 = g * g, Wd = c.vit_width;
  const int64_t img_elems = (int64_t)3 * c.vit_res * c.vit_res;
  for (int64_t i0 = first; i0 < first + n_img; i0 += IMG_CHUNK) {
    const int64_t n = min(IMG_CHUNK, first + n_img - i0);
    MSQ_TRY(im2col<T>(images + i0 * img_elems, n, c.vit_res, c.vit_patch, (T*)b.apatch, st));
    MSQ_TRY((run_gemm<T, float>(m, (const T*)b.apatch, m->conv1.K, m->conv1, nullptr, 0, b.patch + i0 * g2 * Wd, Wd, n * g2, ACT_NONE,
                                st, false)));
  }
  return MSQ_OK;
}

// token assembly + ln_pre + the residual blocks for R pair rows; b.xv holds the final residual stream and b.y
// ln_post(x) in the GEMM operand type.  img_index [R*2] addresses rows of b.patch by image.
template <typename T>
static int run_vit(msq_model* m, const int32_t* img_index, int64_t R, VitBufs& b, cudaStream_t st) {
  const msq_config& c = m->cfg;
  const int g = c.vit_res / c.vit_patch, g2 = g * g, Lv = 1 + 2 * g2, Wd = c.vit_width, heads = Wd / 64;
  const int64_t Mv = R * Lv;
  MSQ_TRY(vit_assemble(b.patch, img_index, R, 2, g2, Wd, m->vit_cls, m->vit_pos, m->ln_pre.g, m->ln_pre.b, 1e-5f, b.xv, st));
  b.post_pending = false;
  if constexpr (!same_type<T, float>::value) {
    constexpr bool SP = is_split<T>::value;
    if (ln_fold_enabled(m) && m->folded && !m->vit.empty() && (SP || !gemm_ln_enabled())) {
      // Deferred LayerNorm: the residual stream x stays raw (fp32 b.xv + bf16 copy b.y + per-row partial sums written by
      // the residual GEMMs); ln_2 / the next ln_1 / ln_post are applied inside the consuming GEMMs' epilogues.
      const int sp = 2 * ceil_div(Wd, 256);
      const size_t nl = m->vit.size();
      MSQ_TRY(layernorm<T>(b.xv, Mv, Wd, m->vit[0].ln1.g, m->vit[0].ln1.b, 1e-5f, nullptr, (T*)b.y, 0, 0, 0, st));
      LnPending pend;
      pend.sp = sp; pend.eps = 1e-5f; pend.dim = Wd;
      for (size_t l = 0; l < nl; ++l) {
        VitLayerW& L = m->vit[l];
        if (l == 0) MSQ_TRY((run_gemm<T, T>(m, (const T*)b.y, Wd, L.qkv, nullptr, 0, (T*)b.qkv, 3 * Wd, Mv, ACT_NONE, st)));
        else MSQ_TRY(run_gemm_fold<T>(b.y, Wd, L.qkv_f, pend, (T*)b.qkv, 3 * Wd, Mv, ACT_NONE, st, SP));
        MSQ_TRY(attention<T>((const T*)b.qkv, R, Lv, heads, 64, 0.125f, nullptr, 0, 0, (T*)b.ctx, st));
        MSQ_TRY(run_gemm_resln(b.ctx, Wd, L.out, b.xv, b.y, nullptr, b.st[0], Mv, st, SP));
        pend.stats = b.st[0]; pend.ln = L.ln2;
        MSQ_TRY(run_gemm_fold<T>(b.y, Wd, L.fc_f, pend, (T*)b.hbuf, 4 * Wd, Mv, ACT_QUICK_GELU, st, SP));
        MSQ_TRY(run_gemm_resln(b.hbuf, 4 * Wd, L.proj, b.xv, b.y, nullptr, b.st[1], Mv, st, SP));
        pend.stats = b.st[1]; pend.ln = l + 1 < nl ? m->vit[l + 1].ln1 : m->ln_post;
      }
      b.post = pend;           // b.y is the RAW stream; ln_post is folded into visn_fc by the caller
      b.post_pending = true;
      return MSQ_OK;
    }
  }
  if constexpr (same_type<T, bf16>::value) {
    if (use_tc(m) && gemm_ln_enabled() && gemm_ln_supported(Wd, Wd) && gemm_ln_supported(Wd, 4 * Wd)) {
      // fused path: every residual GEMM also emits LayerNorm(x) (bf16) for the NEXT GEMM -- ln_2 after out_proj,
      // the next block's ln_1 (ln_post after the last block) after c_proj.  b.y ends up holding ln_post(x).
      const size_t nl = m->vit.size();
      if (nl) MSQ_TRY(layernorm<T>(b.xv, Mv, Wd, m->vit[0].ln1.g, m->vit[0].ln1.b, 1e-5f, nullptr, (T*)b.y, 0, 0, 0, st));
      for (size_t l = 0; l < nl; ++l) {
        VitLayerW& L = m->vit[l];
        const LNp& nxt = l + 1 < nl ? m->vit[l + 1].ln1 : m->ln_post;
        MSQ_TRY((run_gemm<T, T>(m, (const T*)b.y, Wd, L.qkv, nullptr, 0, (T*)b.qkv, 3 * Wd, Mv, ACT_NONE, st)));
        MSQ_TRY(attention<T>((const T*)b.qkv, R, Lv, heads, 64, 0.125f, nullptr, 0, 0, (T*)b.ctx, st));
        MSQ_TRY(gemm_ln((const bf16*)b.ctx, Wd, L.out.w16, L.out.ld, L.out.b, b.xv, Wd, L.ln2.g, L.ln2.b, 1e-5f, b.xv, (bf16*)b.y, Mv, Wd,
                        Wd, true, st));
        MSQ_TRY((run_gemm<T, T>(m, (const T*)b.y, Wd, L.fc, nullptr, 0, (T*)b.hbuf, 4 * Wd, Mv, ACT_QUICK_GELU, st)));
        MSQ_TRY(gemm_ln((const bf16*)b.hbuf, 4 * Wd, L.proj.w16, L.proj.ld, L.proj.b, b.xv, Wd, nxt.g, nxt.b, 1e-5f, b.xv, (bf16*)b.y, Mv,
                        Wd, 4 * Wd, true, st));
      }
      if (nl == 0) MSQ_TRY(layernorm<T>(b.xv, Mv, Wd, m->ln_post.g, m->ln_post.b, 1e-5f, nullptr, (T*)b.y, 0, 0, 0, st));
      return MSQ_OK;
    }
  }
  if (!m->vit.empty()) MSQ_TRY(layernorm<T>(b.xv, Mv, Wd, m->vit[0].ln1.g, m->vit[0].ln1.b, 1e-5f, nullptr, (T*)b.y, 0, 0, 0, st));
  for (auto& L : m->vit) {
    MSQ_TRY((run_gemm<T, T>(m, (const T*)b.y, Wd, L.qkv, nullptr, 0, (T*)b.qkv, 3 * Wd, Mv, ACT_NONE, st)));
    MSQ_TRY(attention<T>((const T*)b.qkv, R, Lv, heads, 64, 0.125f, nullptr, 0, 0, (T*)b.ctx, st));
    const int64_t sl = slab_rows() ? slab_rows() : Mv;
    for (int64_t r0 = 0; r0 < Mv; r0 += sl) {   // x += out_proj(ctx); y = ln_2(x), slab by slab (the LN reads x from L2)
      const int64_t nr = min(sl, Mv - r0);
      MSQ_TRY((run_gemm<T, float>(m, (const T*)b.ctx + r0 * Wd, Wd, L.out, b.xv + r0 * Wd, Wd, b.xv + r0 * Wd, Wd, nr, ACT_NONE, st)));
      MSQ_TRY(layernorm<T>(b.xv + r0 * Wd, nr, Wd, L.ln2.g, L.ln2.b, 1e-5f, nullptr, (T*)b.y + r0 * Wd, 0, 0, 0, st));
    }
    MSQ_TRY((run_gemm<T, T>(m, (const T*)b.y, Wd, L.fc, nullptr, 0, (T*)b.hbuf, 4 * Wd, Mv, ACT_QUICK_GELU, st)));
    const LNp& nxt = (&L == &m->vit.back()) ? m->ln_post : (&L + 1)->ln1;
    for (int64_t r0 = 0; r0 < Mv; r0 += sl) {   // x += c_proj(h); y = next block's ln_1(x) (ln_post after the last block)
      const int64_t nr = min(sl, Mv - r0);
      MSQ_TRY((run_gemm<T, float>(m, (const T*)b.hbuf + r0 * 4 * Wd, 4 * Wd, L.proj, b.xv + r0 * Wd, Wd, b.xv + r0 * Wd, Wd, nr, ACT_NONE, st)));
      MSQ_TRY(layernorm<T>(b.xv + r0 * Wd, nr, Wd, nxt.g, nxt.b, 1e-5f, nullptr, (T*)b.y + r0 * Wd, 0, 0, 0, st));
    }
  }
  if (m->vit.empty()) MSQ_TRY(layernorm<T>(b.xv, Mv, Wd, m->ln_post.g, m->ln_post.b, 1e-5f, nullptr, (T*)b.y, 0, 0, 0, st));
  return MSQ_OK;  // b.y = ln_post(x)
}

// ---- CLIP ModifiedResNet tower ---------------------------------------------------------------------
constexpr int64_t RN_IMG_CHUNK = 256;  // images per trunk pass (bounds the im2col scratch: 8 MB / image in bf16 at width 64)

struct RnDims { size_t a = 0, t = 0, f = 0; };  // per-image element counts: im2col rows, operand-type maps, fp32 maps
static RnDims rn_dims(const msq_config& c) {
  RnDims d;
  auto up = [](size_t& x, size_t v) { if (v > x) x = v; };
  const int w = c.rn_width, w2 = w / 2;
  size_t H = c.vit_res / 2;
  const bool split = c.precise == 2;
  up(d.a, H * H * rn_kpad(27, split)); up(d.a, H * H * rn_kpad(9 * w2, split));
  up(d.f, H * H * w2);   // fp32 staging of the narrow stem outputs (bf16x3: split outputs need N % 64 == 0)
  up(d.t, H * H * w);
  H /= 2;
  size_t cin = w;
  for (int s = 0; s < 4; ++s)
    for (int b = 0; b < c.rn_blocks[s]; ++b) {
      const size_t p = (size_t)w << s, st = (b == 0 && s > 0) ? 2 : 1, Ho = H / st;
      up(d.t, H * H * cin); up(d.t, H * H * p); up(d.a, H * H * rn_kpad(9 * (int)p, split));
      up(d.t, Ho * Ho * 4 * p); up(d.f, Ho * Ho * 4 * p);
      H = Ho; cin = 4 * p;
    }
  return d;
}

template <typename T>
static void plan_rn(const msq_config& c, Planner& p, int64_t n_img, int64_t R, VitBufs* b) {
  const int g2 = (c.vit_res / 32) * (c.vit_res / 32), Lv = 1 + 2 * g2, Cf = 32 * c.rn_width, E = c.rn_embed;
  const RnDims d = rn_dims(c);
  const size_t nc = (size_t)min(n_img, RN_IMG_CHUNK);
  b->apatch = p.take<T>(nc * d.a);
  b->rX = p.take<T>(nc * d.t); b->rO1 = p.take<T>(nc * d.t); b->rO2 = p.take<T>(nc * d.t);
  b->rP1 = p.take<T>(nc * d.t); b->rP2 = p.take<T>(nc * d.t);
  b->rF0 = p.take<float>(nc * d.f); b->rF1 = p.take<float>(nc * d.f);
  b->patch = p.take<float>((size_t)n_img * g2 * Cf);     // trunk output, NHWC, per UNIQUE image
  b->hbuf = p.take<T>((size_t)R * Lv * Cf);               // attnpool token matrix
  b->qkv = p.take<T>((size_t)R * Lv * 3 * Cf);
  b->ctx = p.take<T>((size_t)R * Lv * Cf);
  b->xv = p.take<float>((size_t)R * Lv * E);              // c_proj output
  b->y = p.take<T>((size_t)R * Lv * 2 * E);               // cat(o, o) + visual_pos + visual_token_type
}

// stem + the four bottleneck stages over UNIQUE images [first, first + n_img) -> b.patch rows [img, g2, 32*width]
template <typename T>
static int run_rn_trunk(msq_model* m, const float* images, int64_t n_img, VitBufs& b, cudaStream_t st, int64_t first = 0) {
  const msq_config& c = m->cfg;
  const int S = c.vit_res, w = c.rn_width, Cf = 32 * w, g2 = (S / 32) * (S / 32);
  const int64_t img_elems = (int64_t)3 * S * S;
  T *A = (T*)b.apatch, *X = (T*)b.rX, *O1 = (T*)b.rO1, *O2 = (T*)b.rO2, *P1 = (T*)b.rP1, *P2 = (T*)b.rP2;
  for (int64_t i0 = first; i0 < first + n_img; i0 += RN_IMG_CHUNK) {
    const int64_t n = min(RN_IMG_CHUNK, first + n_img - i0);
    int H = S / 2;
    int64_t M = n * H * H;
    // conv + folded BatchNorm + ReLU; split outputs need N % 64 == 0: the narrow stem convolutions (width / 2 channels) go
    // through an fp32 map and the ReLU / cast pass instead
    auto conv_relu = [&](const T* in, const Lin& L, T* out, int cout) -> int {
      if constexpr (is_split<T>::value) {
        if (cout % 64 != 0) {
          MSQ_TRY((run_gemm<T, float>(m, in, L.K, L, nullptr, 0, b.rF0, cout, M, ACT_NONE, st)));
          return rn_relu_cast<T>(b.rF0, M * cout, cout, out, st);
        }
      }
      return run_gemm<T, T>(m, in, L.K, L, nullptr, 0, out, cout, M, ACT_RELU, st);
    };
    MSQ_TRY(rn_im2col_stem<T>(images + i0 * img_elems, n, S, m->rn_stem[0].K, A, st));
    MSQ_TRY(conv_relu(A, m->rn_stem[0], O1, w / 2));
    MSQ_TRY(rn_im2col3<T>(O1, n, H, H, w / 2, m->rn_stem[1].K, A, st));
    MSQ_TRY(conv_relu(A, m->rn_stem[1], O2, w / 2));
    MSQ_TRY(rn_im2col3<T>(O2, n, H, H, w / 2, m->rn_stem[2].K, A, st));
    MSQ_TRY(conv_relu(A, m->rn_stem[2], O1, w));
    MSQ_TRY(rn_avgpool2<T>(O1, n, H, H, w, X, st));
    H /= 2;
    for (const RnBlockW& B : m->rn_blocks) {
      const int p = B.planes, Ho = H / B.stride;
      M = n * H * H;
      const int64_t Mo = n * Ho * Ho;
      MSQ_TRY((run_gemm<T, T>(m, X, B.cin, B.c1, nullptr, 0, O1, p, M, ACT_RELU, st)));
      MSQ_TRY(rn_im2col3<T>(O1, n, H, H, p, B.c2.K, A, st));
      MSQ_TRY((run_gemm<T, T>(m, A, B.c2.K, B.c2, nullptr, 0, O2, p, M, ACT_RELU, st)));
      const T *o2 = O2, *xs = X;
      if (B.stride > 1) {
        MSQ_TRY(rn_avgpool2<T>(O2, n, H, H, p, P1, st));
        MSQ_TRY(rn_avgpool2<T>(X, n, H, H, B.cin, P2, st));
        o2 = P1; xs = P2;
      }
      if (B.has_ds) MSQ_TRY((run_gemm<T, float>(m, xs, B.cin, B.ds, nullptr, 0, b.rF1, 4 * p, Mo, ACT_NONE, st)));
      MSQ_TRY((run_gemm<T, float>(m, o2, p, B.c3, B.has_ds ? b.rF1 : b.rF0, 4 * p, b.rF0, 4 * p, Mo, ACT_NONE, st)));
      MSQ_TRY(rn_relu_cast<T>(b.rF0, Mo * 4 * p, 4 * p, X, st));
      H = Ho;
    }
    MSQ_CUDA(cudaMemcpyAsync(b.patch + i0 * g2 * Cf, b.rF0, (size_t)n * g2 * Cf * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return MSQ_OK;
}

// AttentionPool2d over R pair rows: token matrix (reference's reshape quirk included) -> MHA over all 1+2g^2 tokens ->
// c_proj -> b.xv [R*Lv, E] fp32; b.y = cat(o, o) + posadd in the GEMM operand type (input of visn_fc).
template <typename T>
static int run_rn_pool(msq_model* m, const int32_t* img_index, int64_t R, VitBufs& b, cudaStream_t st) {
  const msq_config& c = m->cfg;
  const int g2 = (c.vit_res / 32) * (c.vit_res / 32), Lv = 1 + 2 * g2, Cf = 32 * c.rn_width, E = c.rn_embed;
  const int64_t Mv = R * Lv;
  MSQ_TRY(rn_tokens<T>(b.patch, img_index, R, g2, Cf, m->rn_pos, (T*)b.hbuf, st));
  MSQ_TRY((run_gemm<T, T>(m, (const T*)b.hbuf, Cf, m->rn_qkv, nullptr, 0, (T*)b.qkv, 3 * Cf, Mv, ACT_NONE, st)));
  MSQ_TRY(attention<T>((const T*)b.qkv, R, Lv, Cf / 64, 64, 0.125f, nullptr, 0, 0, (T*)b.ctx, st));
  MSQ_TRY((run_gemm<T, float>(m, (const T*)b.ctx, Cf, m->rn_cproj, nullptr, 0, b.xv, E, Mv, ACT_NONE, st)));
  if (m->rn_posadd) MSQ_TRY(rn_finish<T>(b.xv, Mv, Lv, E, m->rn_posadd, (T*)b.y, st));
  return MSQ_OK;
}

// backbone dispatch: per-image part (patch embedding / ResNet trunk) and per-pair part (transformer / attention pool)
template <typename T>
static void plan_visual(const msq_config& c, Planner& p, int64_t n_img, int64_t R, VitBufs* b) {
  if (c.rn_width) { plan_rn<T>(c, p, n_img, R, b); return; }
  plan_vit<T>(c, p, n_img, R, b);
}
template <typename T>
static int run_visual_images(msq_model* m, const float* images, int64_t n_img, VitBufs& b, cudaStream_t st, int64_t first = 0) {
  if (m->cfg.rn_width) return run_rn_trunk<T>(m, images, n_img, b, st, first);
  return run_patch_embed<T>(m, images, n_img, b, st, first);
}
template <typename T>
static int run_visual_pairs(msq_model* m, const int32_t* img_index, int64_t R, VitBufs& b, cudaStream_t st) {
  if (m->cfg.rn_width) return run_rn_pool<T>(m, img_index, R, b, st);
  return run_vit<T>(m, img_index, R, b, st);
}

// precision dispatch: msq_config.precise 0 = bf16 operands, 1 = fp32 FFMA, 2 = bf16x3 (split bf16 operands)
#define MSQ_DISPATCH_T(m, CALL)                                  \
  do {                                                           \
    if ((m)->cfg.precise == 1) { using T = float; return CALL; } \
    if ((m)->cfg.precise == 2) { using T = bf16s; return CALL; } \
    { using T = bf16; return CALL; }                             \
  } while (0)

struct JointBufs {
  float* x; void* xt; float* tmp; void* qkv; void* ctx; void* hbuf; float* mask_add; float* vtmp;
  float* st[2];   // deferred LayerNorm: partial row sums (ping-pong: a residual GEMM reads one set and writes the other)
};

template <typename T>
static void plan_joint(const msq_config& c, Planner& p, int64_t R, int Lt, int Lj, JointBufs* b) {
  const int H = c.hidden;
  b->x = p.take<float>((size_t)R * Lj * H);
  b->xt = p.take<T>((size_t)R * Lj * H);
  b->tmp = p.take<float>((size_t)R * Lj * H);
  b->qkv = p.take<T>((size_t)R * Lj * 3 * H);
  b->ctx = p.take<T>((size_t)R * Lj * H);
  b->hbuf = p.take<T>((size_t)R * Lj * c.inter);
  b->mask_add = p.take<float>((size_t)R * Lt);
  b->vtmp = nullptr;
  const size_t sp = 2 * (size_t)ceil_div(H, 256);
  b->st[0] = p.take<float>((size_t)R * Lj * sp * 2);
  b->st[1] = p.take<float>((size_t)R * Lj * sp * 2);
}

// inner encoder for R pair rows -> b.x holds the final joint stream [R, Lj, H] (fp32), b.xt its T copy
template <typename T>
static int run_inner(msq_model* m, const int64_t* ids, const int64_t* tt, const int64_t* mask, int64_t R, int Lt,
                     const int32_t* img_index, VitBufs& vb, JointBufs& jb, cudaStream_t st) {
  const msq_config& c = m->cfg;
  const int H = c.hidden;
  const bool mm = c.vit_width != 0 && img_index != nullptr;
  const int g2 = mm ? (c.vit_res / c.vit_patch) * (c.vit_res / c.vit_patch) : 0;
  const int Lv = mm ? 1 + 2 * g2 : 0, Lj = Lt + Lv;
  MSQ_REQUIRE(Lt <= c.max_pos, "Lt=%d exceeds max_position_embeddings=%d", Lt, c.max_pos);
  const int64_t Mj = R * Lj;
  MSQ_TRY(embed_ln<T>(ids, tt, R, Lt, Lj, H, m->word, m->pos, m->type, m->emb_ln.g, m->emb_ln.b, 1e-12f, jb.x, (T*)jb.xt, st));
  MSQ_CUDA(launch_k(mask_add_kernel, dim3(ceil_div(R * Lt, 256)), dim3(256), 0, st, mask, R * Lt, jb.mask_add));
  MSQ_LAUNCH_CHECK();
  if (mm) {
    MSQ_TRY(run_visual_pairs<T>(m, img_index, R, vb, st));
    const int Wd = c.vit_width;
    const int64_t Mv = R * Lv;
    // vb.y == ln_post(x); visn_fc output reuses the (now free) fp32 tmp buffer of the joint stream
    if (vb.post_pending) MSQ_TRY(run_gemm_fold<float>(vb.y, Wd, m->visn_fc_f, vb.post, jb.tmp, H, Mv, ACT_NONE, st, is_split<T>::value));
    else MSQ_TRY((run_gemm<T, float>(m, (const T*)vb.y, Wd, m->visn_fc, nullptr, 0, jb.tmp, H, Mv, ACT_NONE, st)));
    MSQ_TRY(layernorm<T>(jb.tmp, Mv, H, m->visn_ln.g, m->visn_ln.b, 1e-12f, jb.x, (T*)jb.xt, Lv, Lj, Lt, st));
  }
  bool fused = false;
  if constexpr (same_type<T, bf16>::value) fused = use_tc(m) && gemm_ln_enabled() && gemm_ln_supported(H, H) && gemm_ln_supported(H, c.inter);
  if constexpr (!same_type<T, float>::value) {
    constexpr bool SP = is_split<T>::value;
    if (ln_fold_enabled(m) && m->folded && !fused && !m->bert.empty()) {
      // Deferred LayerNorm through the post-LN BERT stack: jb.x / jb.xt hold the RAW sums (dense(...) + residual), the
      // LayerNorm after each sub-layer lives in the consumers: folded into the next QKV / intermediate GEMM and applied
      // to the residual operand inside the next residual GEMM's epilogue.  Only the stack's final output is normalised
      // by a kernel of its own.
      LnPending pend;
      pend.sp = 2 * ceil_div(H, 256); pend.eps = 1e-12f; pend.dim = H;
      bool have = false;
      int cur = 0;
      for (auto& L : m->bert) {
        if (have) MSQ_TRY(run_gemm_fold<T>(jb.xt, H, L.qkv_f, pend, (T*)jb.qkv, 3 * H, Mj, ACT_NONE, st, SP));
        else MSQ_TRY((run_gemm<T, T>(m, (const T*)jb.xt, H, L.qkv, nullptr, 0, (T*)jb.qkv, 3 * H, Mj, ACT_NONE, st)));
        MSQ_TRY(attention<T>((const T*)jb.qkv, R, Lj, c.heads, 64, 0.125f, jb.mask_add, Lt, Lt, (T*)jb.ctx, st));
        MSQ_TRY(run_gemm_resln(jb.ctx, H, L.out, jb.x, jb.xt, have ? &pend : nullptr, jb.st[cur ^ 1], Mj, st, SP));
        cur ^= 1; pend.stats = jb.st[cur]; pend.ln = L.ln1; have = true;
        MSQ_TRY(run_gemm_fold<T>(jb.xt, H, L.up_f, pend, (T*)jb.hbuf, c.inter, Mj, ACT_GELU_ERF, st, SP));
        MSQ_TRY(run_gemm_resln(jb.hbuf, c.inter, L.down, jb.x, jb.xt, &pend, jb.st[cur ^ 1], Mj, st, SP));
        cur ^= 1; pend.stats = jb.st[cur]; pend.ln = L.ln2;
      }
      MSQ_TRY(layernorm<T>(jb.x, Mj, H, pend.ln.g, pend.ln.b, 1e-12f, jb.tmp, (T*)jb.xt, 0, 0, 0, st));
      std::swap(jb.x, jb.tmp);   // callers read the final stream from jb.x
      return MSQ_OK;
    }
  }
  const int64_t sl = slab_rows() ? slab_rows() : Mj;
  for (auto& L : m->bert) {
    MSQ_TRY((run_gemm<T, T>(m, (const T*)jb.xt, H, L.qkv, nullptr, 0, (T*)jb.qkv, 3 * H, Mj, ACT_NONE, st)));
    MSQ_TRY(attention<T>((const T*)jb.qkv, R, Lj, c.heads, 64, 0.125f, jb.mask_add, Lt, Lt, (T*)jb.ctx, st));
    if (fused) {
      // x <- LN(dense(ctx) + x) written in place (fp32) together with its bf16 copy: no pre-LN round trip through HBM
      MSQ_TRY(gemm_ln((const bf16*)jb.ctx, H, L.out.w16, L.out.ld, L.out.b, jb.x, H, L.ln1.g, L.ln1.b, 1e-12f, jb.x, (bf16*)jb.xt, Mj, H, H,
                      false, st));
    } else {
      for (int64_t r0 = 0; r0 < Mj; r0 += sl) {
        const int64_t nr = min(sl, Mj - r0);
        MSQ_TRY((run_gemm<T, float>(m, (const T*)jb.ctx + r0 * H, H, L.out, jb.x + r0 * H, H, jb.tmp, H, nr, ACT_NONE, st)));
        MSQ_TRY(layernorm<T>(jb.tmp, nr, H, L.ln1.g, L.ln1.b, 1e-12f, jb.x + r0 * H, (T*)jb.xt + r0 * H, 0, 0, 0, st));
      }
    }
    MSQ_TRY((run_gemm<T, T>(m, (const T*)jb.xt, H, L.up, nullptr, 0, (T*)jb.hbuf, c.inter, Mj, ACT_GELU_ERF, st)));
    if (fused) {
      MSQ_TRY(gemm_ln((const bf16*)jb.hbuf, c.inter, L.down.w16, L.down.ld, L.down.b, jb.x, H, L.ln2.g, L.ln2.b, 1e-12f, jb.x,
                      (bf16*)jb.xt, Mj, H, c.inter, false, st));
    } else {
      for (int64_t r0 = 0; r0 < Mj; r0 += sl) {
        const int64_t nr = min(sl, Mj - r0);
        MSQ_TRY((run_gemm<T, float>(m, (const T*)jb.hbuf + r0 * c.inter, c.inter, L.down, jb.x + r0 * H, H, jb.tmp, H, nr, ACT_NONE, st)));
        MSQ_TRY(layernorm<T>(jb.tmp, nr, H, L.ln2.g, L.ln2.b, 1e-12f, jb.x + r0 * H, (T*)jb.xt + r0 * H, 0, 0, 0, st));
      }
    }
  }
  return MSQ_OK;
}

struct HeadBufs {   // whole batch, fp32
  float *mix, *rel6, *sents, *r0, *para, *pa, *pb, *pn, *pqkv, *pctx, *pff, *h0, *keyin, *key, *sents_ext, *xg, *t4;
  void *topt, *ttb, *dec;
};

template <typename T>
static void plan_heads(const msq_config& c, Planner& p, int64_t B, int N, int64_t Rc, int Lt, int Kp, HeadBufs* h, int beam = 16) {
  const int H = c.hidden;
  const int64_t R = B * N * (N - 1);
  h->topt = p.take<T>((size_t)Rc * Lt * H);
  h->ttb = p.take<T>((size_t)Rc * Lt * H);
  h->mix = p.take<float>((size_t)R * 2 * H);
  h->rel6 = p.take<float>((size_t)R * 6);
  h->sents = p.take<float>((size_t)B * N * H);
  h->r0 = p.take<float>((size_t)B * N * N * Kp);
  h->para = p.take<float>((size_t)B * N * H);
  h->pa = p.take<float>((size_t)B * N * H);
  h->pb = p.take<float>((size_t)B * N * H);
  h->pn = p.take<float>((size_t)B * N * H);
  h->pqkv = p.take<float>((size_t)B * N * 3 * H);
  h->pctx = p.take<float>((size_t)B * N * H);
  h->pff = p.take<float>((size_t)B * N * c.para_ff);
  h->h0 = p.take<float>((size_t)B * H);
  h->keyin = p.take<float>((size_t)B * N * 2 * H);
  h->key = p.take<float>((size_t)B * N * H);
  h->sents_ext = p.take<float>((size_t)B * (N + 1) * H);
  h->xg = p.take<float>((size_t)B * (N + 1) * 4 * H);
  h->t4 = p.take<float>((size_t)B * N * N * 4 * H);
  h->dec = p.take<char>(beam_search_scratch_bytes(B, N, beam, H));
}

static int run_gemm32(const float* A, int lda, const Lin& w, const float* resid, int ldr, float* C, int ldc, int64_t M, int act,
                      cudaStream_t st) {
  GemmArgs g;
  g.A = A; g.W = w.w32; g.bias = w.b; g.resid = resid; g.C = C; g.C2 = nullptr;
  g.M = M; g.N = w.N; g.K = w.K; g.lda = lda; g.ldw = w.ld; g.ldc = ldc; g.ldr = ldr; g.act = act;
  return gemm_simt<float, float>(g, st);
}

// paragraph encoder + key projection for B manuals (fp32): sents -> para, h0, key
// (models/berson/encoder.py:46-59, 9-29; models/berson/neural.py:30-33; modeling_bert.py:1343-1357)
static int run_paragraph(msq_model* m, HeadBufs& h, int64_t B, int N, cudaStream_t st) {
  const msq_config& c = m->cfg;
  const int H = c.hidden;
  const int64_t M = B * N;
  const float* x = h.sents;  // mask_cls is all ones: x * mask == x
  for (size_t l = 0; l < m->para.size(); ++l) {
    ParaLayerW& L = m->para[l];
    const float* y = x;
    if (l != 0) {
      MSQ_TRY(layernorm<float>(x, M, H, L.ln_in.g, L.ln_in.b, 1e-6f, h.pn, nullptr, 0, 0, 0, st));
      y = h.pn;
    }
    MSQ_TRY(run_gemm32(y, H, L.qkv, nullptr, 0, h.pqkv, 3 * H, M, ACT_NONE, st));
    MSQ_TRY(para_attention(h.pqkv, B, N, c.para_heads, H, h.pctx, st));
    MSQ_TRY(run_gemm32(h.pctx, H, L.fin, x, H, h.pa, H, M, ACT_NONE, st));                   // out = attn + x
    MSQ_TRY(layernorm<float>(h.pa, M, H, L.ln_ff.g, L.ln_ff.b, 1e-6f, h.pn, nullptr, 0, 0, 0, st));
    MSQ_TRY(run_gemm32(h.pn, H, L.w1, nullptr, 0, h.pff, c.para_ff, M, ACT_GELU_TANH, st));
    MSQ_TRY(run_gemm32(h.pff, c.para_ff, L.w2, h.pa, H, h.pb, H, M, ACT_NONE, st));          // x' = ffn + out
    x = h.pb;
  }
  MSQ_TRY(layernorm<float>(x, M, H, m->para_ln.g, m->para_ln.b, 1e-6f, h.para, nullptr, 0, 0, 0, st));
  MSQ_TRY(para_finish(h.sents, h.para, B, N, H, h.h0, h.keyin, st));
  MSQ_TRY(run_gemm32(h.keyin, 2 * H, m->key_lin, nullptr, 0, h.key, H, M, ACT_NONE, st));
  return MSQ_OK;
}

// decode pre-projections + beam search from (sents, key, h0, r0)
static int run_decode(msq_model* m, const float* sents, const float* key, const float* h0, const float* r0, float* sents_ext,
                      float* xg, float* t4, int64_t B, int N, int beam, int32_t* perm, int32_t* tr_ix, float* tr_cost,
                      float* tr_logp, cudaStream_t st, const int32_t* forced = nullptr, float* final_cost = nullptr,
                      void* scratch = nullptr) {
  const int H = m->cfg.hidden;
  MSQ_CUDA(launch_k(sents_ext_kernel, dim3(ceil_div(B * (N + 1) * (int64_t)H, 256)), dim3(256), 0, st, sents, B, N, H, sents_ext));
  MSQ_LAUNCH_CHECK();
  const bool tc = use_tc(m) && m->dec.wcat3 != nullptr && scratch != nullptr && H % 64 == 0 && decode_tc_enabled();
  if (tc) {
    // pre-projections on the tensor cores: three-plane operands (bf16x6).  The plane copies of sents_ext / r0 live in the
    // tail of the decode scratch buffer (beam_search_scratch_bytes reserves them).
    char* tail = reinterpret_cast<char*>(scratch) + beam_search_state_bytes(B, beam, H);
    bf16* s3 = reinterpret_cast<bf16*>(tail);
    bf16* r3 = s3 + (((size_t)B * (N + 1) * 3 * H + 127) & ~size_t(127));
    MSQ_TRY(pack_split3(sents_ext, B * (N + 1), H, H, H, s3, st));
    MSQ_TRY(pack_split3(r0, B * N * N, m->Kp, m->Kp, m->Kp, r3, st));
    GemmArgs g;
    g.A = s3; g.W = m->wih3; g.bias = m->xg_lin.b; g.resid = nullptr; g.C = xg; g.C2 = nullptr;
    g.M = B * (N + 1); g.N = 4 * H; g.K = H; g.lda = H; g.ldw = H; g.ldc = 4 * H; g.ldr = 0; g.act = ACT_NONE; g.split = 2;
    g.trunc = decode_trunc(m->cfg.precise);
    MSQ_TRY(gemm_tc<float>(g, st));
    g.A = r3; g.W = m->wpw3; g.bias = nullptr; g.C = t4; g.M = B * N * N; g.K = m->Kp; g.lda = m->Kp; g.ldw = m->Kp;
    MSQ_TRY(gemm_tc<float>(g, st));
  } else {
    MSQ_TRY(run_gemm32(sents_ext, H, m->xg_lin, nullptr, 0, xg, 4 * H, B * (N + 1), ACT_NONE, st));
    MSQ_TRY(run_gemm32(r0, m->Kp, m->t4_lin, nullptr, 0, t4, 4 * H, B * N * N, ACT_NONE, st));
  }
  if (tr_ix) MSQ_CUDA(cudaMemsetAsync(tr_ix, 0xff, (size_t)B * (N - 1) * beam * sizeof(int32_t), st));
  DecodeIO io;
  io.xg = xg; io.t4 = t4; io.key0 = key; io.h0 = h0; io.B = B; io.N = N; io.W = beam; io.H = H;
  io.perm = perm; io.trace_ix = tr_ix; io.trace_cost = tr_cost; io.trace_logp = tr_logp;
  io.forced = forced; io.final_cost = final_cost; io.scratch = scratch; io.tc = tc ? 1 : 0;
  io.trunc = tc ? decode_trunc(m->cfg.precise) : 0;
  return beam_search(m->dec, io, st);
}

// manuals encoded per micro-batch: MSQ_CHUNK_MANUALS, else 32 (bf16 / fp32) or 64 (bf16x3: its GEMMs are tensor-bound, so larger
// micro-batches only shrink the tile-quantisation tails: 260 -> 270 manuals/s measured)
static int chunk_manuals(const msq_model* m = nullptr) {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MSQ_CHUNK_MANUALS"); v = e ? atoi(e) : 0; if (v < 0) v = 0; }
  if (v > 0) return v;
  return (m && m->cfg.precise == 2) ? 64 : 32;
}

// Full path for B manuals.  images are UNIQUE images [n_img,3,S,S]; img_index [B*P*2] indexes them.
template <typename T>
static int run_path(msq_model* m, const int64_t* ids, const int64_t* tt, const int64_t* mask, const int64_t* sep, int64_t B, int N,
                    int Lt, const float* images, int64_t n_img, const int32_t* img_index, const msq_encode_out* out, int beam,
                    int32_t* perm, cudaStream_t st, const int32_t* forced = nullptr, const int64_t* pair_labels = nullptr,
                    float lam = 0.f, float* loss_out = nullptr, const cudaEvent_t* img_ready = nullptr, int64_t img_chunk = 0) {
  const msq_config& c = m->cfg;
  MSQ_REQUIRE(m->packed, "msq_model_pack() has not been called");
  MSQ_REQUIRE(m->has_bert && m->has_heads, "model lacks the inner encoder or the BERSON head weights");
  MSQ_REQUIRE(c.vit_width == 0 || m->has_vit, "model lacks the visual tower weights");
  MSQ_REQUIRE(N >= 2 && N <= 16, "N=%d out of range [2,16]", N);
  MSQ_REQUIRE(!(c.vit_width != 0 && images == nullptr), "multimodal model needs images");
  MSQ_REQUIRE(beam >= 0 && beam <= 16, "beam width %d out of range [1,16]", beam);
  const int H = c.hidden, P = N * (N - 1);
  const bool mm = c.vit_width != 0;
  const int g2 = mm ? (c.vit_res / c.vit_patch) * (c.vit_res / c.vit_patch) : 0;
  const int Lv = mm ? 1 + 2 * g2 : 0, Lj = Lt + Lv;
  const int64_t R = B * P;
  const int64_t Bc = min((int64_t)chunk_manuals(m), B), Rc = Bc * P;

  VitBufs vb{}; JointBufs jb{}; HeadBufs hb{};
  for (int pass = 0; pass < 2; ++pass) {
    Planner p{&m->ws, pass == 0};
    if (pass == 1) m->ws.reset();
    if (mm) plan_visual<T>(c, p, n_img, Rc, &vb);
    plan_joint<T>(c, p, Rc, Lt, Lj, &jb);
    plan_heads<T>(c, p, B, N, Rc, Lt, m->Kp, &hb, beam > 0 ? beam : 1);
    if (pass == 0) MSQ_TRY(m->ws.reserve(p.need + 4096, st));
  }
  // patch embedding once per UNIQUE image, not per pair slot.  When the caller streams the images in manual order
  // (img_ready != nullptr: images of manuals [b0, b0+bc) are rows [b0*N, (b0+bc)*N), one event per micro-batch),
  // each micro-batch embeds its own images as soon as its host->device copy has landed.
  if (mm && !img_ready) MSQ_TRY(run_visual_images<T>(m, images, n_img, vb, st));

  for (int64_t b0 = 0; b0 < B; b0 += Bc) {
    const int64_t bc = min(Bc, B - b0), rc = bc * P, r0 = b0 * P;
    if (mm && img_ready) {
      // images arrive in upload chunks of img_chunk manuals (one event each; img_chunk divides Bc): embed each as it lands
      const int64_t ic = img_chunk > 0 ? img_chunk : Bc;
      for (int64_t u0 = b0; u0 < b0 + bc; u0 += ic) {
        const int64_t uc = min(ic, b0 + bc - u0);
        MSQ_CUDA(cudaStreamWaitEvent(st, img_ready[u0 / ic], 0));
        MSQ_TRY(run_visual_images<T>(m, images, uc * N, vb, st, u0 * N));
      }
    }
    MSQ_TRY((run_inner<T>(m, ids + r0 * Lt, tt + r0 * Lt, mask + r0 * Lt, rc, Lt, mm ? img_index + r0 * 2 : nullptr, vb, jb, st)));
    // ---- pooling for this chunk
    if (out && out->top_vec) MSQ_TRY((gather_rows<float, float>(jb.x, rc * Lt, H, Lt, Lj, 0, out->top_vec + r0 * Lt * H, st)));
    MSQ_TRY((gather_rows<float, T>(jb.x, rc * Lt, H, Lt, Lj, 0, (T*)hb.topt, st)));
    // cfg.reserved bit 0 ("cls_pooler"): the inner model is a HuggingFace AutoModel (trainers/train.py:1928-1933) whose
    // outputs[1] is tanh(pooler.dense(seq[:,0])): that vector replaces row 0 of every pair in the heads' view of the
    // stream (top_vec above keeps the encoder's own row; token pooling never reads position 0).  hb.topt is free once
    // sentence_tran has consumed it.
    bool cls_pooled = false;
    auto pool_cls = [&]() -> int {
      if (!(c.reserved & 1) || cls_pooled) return MSQ_OK;
      MSQ_REQUIRE(m->has_pooler, "cls_pooler: the model has no pooler.dense weights");
      float* cls_in = reinterpret_cast<float*>(hb.topt);
      float* cls_out = cls_in + (size_t)rc * H;
      MSQ_TRY((gather_rows<float, float>(jb.x, rc, H, 1, Lj, 0, cls_in, st)));
      MSQ_TRY(run_gemm32(cls_in, H, m->pooler, nullptr, 0, cls_out, H, rc, ACT_TANH, st));
      MSQ_TRY(scatter_rows(cls_out, rc, H, 1, Lj, 0, jb.x, st));
      cls_pooled = true;
      return MSQ_OK;
    };
    if constexpr (is_split<T>::value) {   // tanh(sentence_tran(.)) stays fp32 (hb.ttb holds 4 bytes per element)
      MSQ_TRY((run_gemm<T, float>(m, (const T*)hb.topt, H, m->sent_tran, nullptr, 0, (float*)hb.ttb, H, rc * Lt, ACT_TANH, st)));
      MSQ_TRY(pool_cls());
      MSQ_TRY(token_pool<float>((const float*)hb.ttb, jb.x, rc, Lt, Lj, H, m->w2, m->b2, sep + r0 * 2, m->w_rel, m->b_rel,
                                hb.mix + r0 * 2 * H, hb.rel6 + r0 * 6, st));
    } else {
    MSQ_TRY((run_gemm<T, T>(m, (const T*)hb.topt, H, m->sent_tran, nullptr, 0, (T*)hb.ttb, H, rc * Lt, ACT_TANH, st)));
    MSQ_TRY(pool_cls());
    MSQ_TRY(token_pool<T>((const T*)hb.ttb, jb.x, rc, Lt, Lj, H, m->w2, m->b2, sep + r0 * 2, m->w_rel, m->b_rel, hb.mix + r0 * 2 * H,
                          hb.rel6 + r0 * 6, st));
    }
    const int64_t cell0 = b0 * N * N;
    MSQ_TRY(pool_cls());
    MSQ_TRY(edge_pool(hb.mix + r0 * 2 * H, jb.x, hb.rel6 + r0 * 6, bc, N, Lj, H, m->w_in2, hb.sents + b0 * N * H,
                      hb.r0 + cell0 * m->Kp, m->Kp, out && out->cls_mat ? out->cls_mat + cell0 * H : nullptr,
                      out && out->score_mat ? out->score_mat + cell0 * 2 : nullptr, out && out->his1 ? out->his1 + cell0 * 2 : nullptr,
                      out && out->his2 ? out->his2 + cell0 * 2 : nullptr, out && out->cls ? out->cls + r0 * H : nullptr, st));
  }
  MSQ_TRY(run_paragraph(m, hb, B, N, st));
  if (out) {
    auto cp = [&](float* dst, const float* src, size_t n) -> int {
      if (dst) MSQ_CUDA(cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
      return MSQ_OK;
    };
    MSQ_TRY(cp(out->sents, hb.sents, (size_t)B * N * H));
    MSQ_TRY(cp(out->para, hb.para, (size_t)B * N * H));
    MSQ_TRY(cp(out->h0, hb.h0, (size_t)B * H));
    MSQ_TRY(cp(out->key, hb.key, (size_t)B * N * H));
    if (out->cls_score) MSQ_CUDA(cudaMemcpy2DAsync(out->cls_score, 2 * sizeof(float), hb.rel6, 6 * sizeof(float), 2 * sizeof(float),
                                                   (size_t)R, cudaMemcpyDeviceToDevice, st));
  }
  if (perm && !forced)
    MSQ_TRY(run_decode(m, hb.sents, hb.key, hb.h0, hb.r0, hb.sents_ext, hb.xg, hb.t4, B, N, beam, perm, nullptr, nullptr, nullptr, st, nullptr,
                       nullptr, hb.dec));
  if (forced) {
    // teacher-forced NLL of the ground-truth order + lam * pairwise NLL (modeling_bert.py:1098-1174), forward only
    float* nll = hb.pn;  // [B] scratch (paragraph buffers are free again)
    MSQ_TRY(run_decode(m, hb.sents, hb.key, hb.h0, hb.r0, hb.sents_ext, hb.xg, hb.t4, B, N, 1, perm, nullptr, nullptr, nullptr, st, forced,
                       nll, hb.dec));
    MSQ_CUDA(launch_k(training_loss_kernel, dim3(1), dim3(256), 0, st, nll, hb.rel6, pair_labels, B, N, lam, loss_out));
    MSQ_LAUNCH_CHECK();
  }
  return MSQ_OK;
}

}  // namespace msq

// =====================================================================================================
// C ABI: encoders / encode / decode / whole path
// =====================================================================================================

extern "C" int msq_vit_forward(msq_model* m, const float* images_dev, int64_t n_img, const int32_t* img_index_dev, int64_t R,
                               float* out_dev, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && m->packed && m->has_vit, "model has no visual tower / not packed");
  cudaStream_t st = (cudaStream_t)stream;
  const msq_config& c = m->cfg;
  const int g2 = (c.vit_res / c.vit_patch) * (c.vit_res / c.vit_patch), Lv = 1 + 2 * g2;
  VitBufs vb{};
  auto go = [&](auto tag) -> int {
    using T = decltype(tag);
    for (int pass = 0; pass < 2; ++pass) {
      Planner p{&m->ws, pass == 0};
      if (pass == 1) m->ws.reset();
      plan_visual<T>(c, p, n_img, R, &vb);
      if (pass == 0) MSQ_TRY(m->ws.reserve(p.need + 4096, st));
    }
    MSQ_TRY(run_visual_images<T>(m, images_dev, n_img, vb, st));
    MSQ_TRY(run_visual_pairs<T>(m, img_index_dev, R, vb, st));
    if (c.rn_width) return rn_finish<float>(vb.xv, R * Lv, Lv, c.rn_embed, nullptr, out_dev, st);  // cat([x, x]) (model.py:106)
    return layernorm<float>(vb.xv, R * Lv, c.vit_width, m->ln_post.g, m->ln_post.b, 1e-5f, out_dev, nullptr, 0, 0, 0, st);
  };
  return c.precise == 1 ? go(float()) : (c.precise == 2 ? go(bf16s()) : go(bf16()));
}

extern "C" int msq_inner_forward(msq_model* m, const int64_t* ids_dev, const int64_t* tt_dev, const int64_t* mask_dev, int64_t R,
                                 int32_t Lt, const float* images_dev, int64_t n_img, const int32_t* img_index_dev, float* lang_dev,
                                 float* visn_dev, float* pooled_dev, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && m->packed && m->has_bert, "model not packed / no inner encoder weights");
  cudaStream_t st = (cudaStream_t)stream;
  const msq_config& c = m->cfg;
  const bool mm = m->has_vit && images_dev != nullptr;
  const int g2 = mm ? (c.vit_res / c.vit_patch) * (c.vit_res / c.vit_patch) : 0, Lv = mm ? 1 + 2 * g2 : 0, Lj = Lt + Lv;
  const int H = c.hidden;
  VitBufs vb{}; JointBufs jb{};
  auto go = [&](auto tag) -> int {
    using T = decltype(tag);
    for (int pass = 0; pass < 2; ++pass) {
      Planner p{&m->ws, pass == 0};
      if (pass == 1) m->ws.reset();
      if (mm) plan_visual<T>(c, p, n_img, R, &vb);
      plan_joint<T>(c, p, R, Lt, Lj, &jb);
      if (pass == 0) MSQ_TRY(m->ws.reserve(p.need + 4096, st));
    }
    if (mm) MSQ_TRY(run_visual_images<T>(m, images_dev, n_img, vb, st));
    MSQ_TRY((run_inner<T>(m, ids_dev, tt_dev, mask_dev, R, Lt, mm ? img_index_dev : nullptr, vb, jb, st)));
    if (lang_dev) MSQ_TRY((gather_rows<float, float>(jb.x, R * Lt, H, Lt, Lj, 0, lang_dev, st)));
    if (visn_dev && mm) MSQ_TRY((gather_rows<float, float>(jb.x, R * Lv, H, Lv, Lj, Lt, visn_dev, st)));
    if (pooled_dev) {
      MSQ_REQUIRE(m->has_pooler, "model has no pooler.dense weights");
      MSQ_TRY(run_gemm32(jb.x, Lj * H, m->pooler, nullptr, 0, pooled_dev, H, R, ACT_NONE, st));
    }
    return MSQ_OK;
  };
  return c.precise == 1 ? go(float()) : (c.precise == 2 ? go(bf16s()) : go(bf16()));
}

extern "C" int msq_encode(msq_model* m, const int64_t* ids_dev, const int64_t* tt_dev, const int64_t* mask_dev,
                          const int64_t* sep_dev, int64_t B, int32_t N, int32_t Lt, const float* images_dev, int64_t n_img,
                          const int32_t* img_index_dev, const msq_encode_out* out, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && out, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  MSQ_DISPATCH_T(m, run_path<T>(m, ids_dev, tt_dev, mask_dev, sep_dev, B, N, Lt, images_dev, n_img, img_index_dev, out, 0, nullptr, st));
}

extern "C" int msq_order_manuals_dev(msq_model* m, const int64_t* ids_dev, const int64_t* tt_dev, const int64_t* mask_dev,
                                     const int64_t* sep_dev, int64_t B, int32_t N, int32_t Lt, const float* images_dev,
                                     int64_t n_img, const int32_t* img_index_dev, int32_t beam, int32_t* perm_dev, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && perm_dev, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  MSQ_DISPATCH_T(m, run_path<T>(m, ids_dev, tt_dev, mask_dev, sep_dev, B, N, Lt, images_dev, n_img, img_index_dev, nullptr, beam, perm_dev, st));
}

// BertForOrdering._forward loss VALUE (modeling_bert.py:943-1174, default objectives), forward only
extern "C" int msq_training_loss(msq_model* m, const int64_t* ids_dev, const int64_t* tt_dev, const int64_t* mask_dev,
                                 const int64_t* sep_dev, int64_t B, int32_t N, int32_t Lt, const float* images_dev, int64_t n_img,
                                 const int32_t* img_index_dev, const int32_t* ground_truth_dev, const int64_t* pairwise_labels_dev,
                                 float lam, int32_t* perm_scratch_dev, float* loss_dev, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && ground_truth_dev && pairwise_labels_dev && perm_scratch_dev && loss_dev, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  MSQ_DISPATCH_T(m, run_path<T>(m, ids_dev, tt_dev, mask_dev, sep_dev, B, N, Lt, images_dev, n_img, img_index_dev, nullptr, 1,
                                perm_scratch_dev, st, ground_truth_dev, pairwise_labels_dev, lam, loss_dev));
}

extern "C" int msq_beam_search(msq_model* m, const float* sents_dev, const float* key_dev, const float* h0_dev,
                               const float* cls_mat_dev, const float* score_mat_dev, int64_t B, int32_t N, int32_t beam,
                               int32_t* perm_dev, int32_t* trace_ix_dev, float* trace_cost_dev, float* trace_logp_dev, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && m->packed && m->has_heads, "model not packed / no BERSON head weights");
  MSQ_REQUIRE(N >= 2 && N <= 16, "N=%d out of range", N);
  cudaStream_t st = (cudaStream_t)stream;
  const int H = m->cfg.hidden;
  float *r0 = nullptr, *sents_ext = nullptr, *xg = nullptr, *t4 = nullptr;
  void* dscr = nullptr;
  MSQ_REQUIRE(beam >= 1 && beam <= 16, "beam width %d out of range [1,16]", beam);
  for (int pass = 0; pass < 2; ++pass) {
    Planner p{&m->ws, pass == 0};
    if (pass == 1) m->ws.reset();
    r0 = p.take<float>((size_t)B * N * N * m->Kp);
    sents_ext = p.take<float>((size_t)B * (N + 1) * H);
    xg = p.take<float>((size_t)B * (N + 1) * 4 * H);
    t4 = p.take<float>((size_t)B * N * N * 4 * H);
    dscr = p.take<char>(beam_search_scratch_bytes(B, N, beam, H));
    if (pass == 0) MSQ_TRY(m->ws.reserve(p.need + 4096, st));
  }
  if (B == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(build_r0_kernel, dim3((unsigned)(B * N * N)), dim3(128), 0, st, cls_mat_dev, score_mat_dev, B * N * N, H, m->Kp, r0));
  MSQ_LAUNCH_CHECK();
  return run_decode(m, sents_dev, key_dev, h0_dev, r0, sents_ext, xg, t4, B, N, beam, perm_dev, trace_ix_dev, trace_cost_dev,
                    trace_logp_dev, st, nullptr, nullptr, dscr);
}

// BertForOrdering.step with the reference's materialised tensors (modeling_bert.py:1368-1402)
extern "C" int msq_decode_step(msq_model* m, const float* prev_y_dev, const float* h_dev, const float* c_dev, const float* key0_dev,
                               const uint8_t* pointed_mask_dev, float* rela_vec_dev, const uint8_t* rela_mask_dev,
                               const float* hist1_dev, const float* hist2_dev, const uint8_t* l1_mask_dev,
                               const uint8_t* l2_mask_dev, int32_t Wb, int32_t N, float* h_out_dev, float* c_out_dev,
                               float* logp_out_dev, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && m->packed && m->has_heads, "model not packed / no BERSON head weights");
  MSQ_REQUIRE(Wb >= 1 && N >= 2 && N <= 16, "bad Wb/N");
  cudaStream_t st = (cudaStream_t)stream;
  const int H = m->cfg.hidden;
  float *g1 = nullptr, *g2 = nullptr, *q = nullptr, *pw = nullptr, *keys = nullptr;
  for (int pass = 0; pass < 2; ++pass) {
    Planner p{&m->ws, pass == 0};
    if (pass == 1) m->ws.reset();
    g1 = p.take<float>((size_t)Wb * 4 * H);
    g2 = p.take<float>((size_t)Wb * 4 * H);
    q = p.take<float>((size_t)Wb * H);
    pw = p.take<float>((size_t)Wb * N * m->Kp4);
    keys = p.take<float>((size_t)Wb * N * H);
    if (pass == 0) MSQ_TRY(m->ws.reserve(p.need + 4096, st));
  }
  MSQ_TRY(run_gemm32(prev_y_dev, H, m->wih_raw, nullptr, 0, g1, 4 * H, Wb, ACT_NONE, st));
  MSQ_TRY(run_gemm32(h_dev, H, m->whh_raw, g1, 4 * H, g2, 4 * H, Wb, ACT_NONE, st));
  MSQ_TRY(decode_step_lstm(g2, c_dev, Wb, H, h_out_dev, c_out_dev, st));
  MSQ_TRY(run_gemm32(h_out_dev, H, m->wq_raw, nullptr, 0, q, H, Wb, ACT_NONE, st));
  StepIO io;
  io.rela = rela_vec_dev; io.rela_mask = rela_mask_dev; io.hist1 = hist1_dev; io.hist2 = hist2_dev; io.l1 = l1_mask_dev;
  io.l2 = l2_mask_dev; io.Wb = Wb; io.N = N; io.H = H; io.Kp4 = m->Kp4; io.pw = pw;
  MSQ_TRY(decode_step_parts(io, st));
  GemmArgs g;
  g.A = pw; g.W = m->pwk_raw.w32; g.bias = nullptr; g.resid = nullptr; g.C = keys; g.C2 = nullptr;
  g.M = (int64_t)Wb * N; g.N = H; g.K = m->Kp4; g.lda = m->Kp4; g.ldw = m->Kp4; g.ldc = H; g.ldr = 0; g.act = ACT_NONE;
  MSQ_TRY((gemm_simt<float, float>(g, st)));
  return decode_step_score(q, keys, key0_dev, pointed_mask_dev, m->dec.wt, m->dec.bt, Wb, N, H, logp_out_dev, st);
}

// models/pointer_module.py p1: LSTMPointerModule.forward (690-749).  Stand-alone: weights are passed directly.
extern "C" int msq_pointer_p1(const float* enc_dev, const float* cls_dev, const int64_t* y_dev, const float* w1_dev,
                              const float* w2_dev, const float* v_dev, const float* wih_dev, const float* whh_dev,
                              const float* bih_dev, const float* bhh_dev, int64_t B, int32_t N, int32_t H, int32_t U,
                              float* preds_dev, float* ce_scratch_dev, float* loss_dev, void* stream) {
  MSQ_REQUIRE(enc_dev && cls_dev && y_dev && preds_dev && ce_scratch_dev && loss_dev, "null argument");
  return pointer_p1(enc_dev, cls_dev, y_dev, w1_dev, w2_dev, v_dev, wih_dev, whh_dev, bih_dev, bhh_dev, B, N, H, U, preds_dev,
                    ce_scratch_dev, loss_dev, (cudaStream_t)stream);
}

static int stage_reserve(msq_model* m, size_t need, cudaStream_t st) {
  if (need <= m->stage_cap) return MSQ_OK;
  MSQ_CUDA(cudaStreamSynchronize(st));
  if (m->stage) MSQ_CUDA(cudaFree(m->stage));
  m->stage = nullptr; m->stage_cap = 0;
  MSQ_CUDA(cudaMalloc(&m->stage, need));
  m->stage_cap = need;
  return MSQ_OK;
}

extern "C" int msq_order_manuals_host(msq_model* m, const int64_t* ids_host, const int64_t* tt_host, const int64_t* mask_host,
                                      const int64_t* sep_host, int64_t B, int32_t N, int32_t Lt, const float* images_host,
                                      int64_t n_img, const int32_t* img_index_host, int32_t beam, int32_t* perm_host, void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && m->packed && perm_host, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t R = B * N * (N - 1);
  const size_t n_tok = (size_t)R * Lt, img_elems = images_host ? (size_t)n_img * 3 * m->cfg.vit_res * m->cfg.vit_res : 0;
  // staging buffer: owned by the model, outside the workspace arena (which run_path re-plans)
  const size_t need = 3 * n_tok * 8 + (size_t)R * 2 * 8 + img_elems * 4 + (size_t)R * 2 * 4 + (size_t)B * N * 4 + 8 * 256;
  MSQ_TRY(stage_reserve(m, need, st));
  char* stage = m->stage;
  size_t off = 0;
  auto carve = [&](size_t bytes) { char* p = stage + off; off += (bytes + 255) & ~size_t(255); return p; };
  int64_t* ids = (int64_t*)carve(n_tok * 8);
  int64_t* tt = (int64_t*)carve(n_tok * 8);
  int64_t* mask = (int64_t*)carve(n_tok * 8);
  int64_t* sep = (int64_t*)carve((size_t)R * 2 * 8);
  float* img = (float*)carve(img_elems * 4);
  int32_t* idx = (int32_t*)carve((size_t)R * 2 * 4);
  int32_t* perm = (int32_t*)carve((size_t)B * N * 4);
  MSQ_CUDA(cudaMemcpyAsync(ids, ids_host, n_tok * 8, cudaMemcpyHostToDevice, st));
  MSQ_CUDA(cudaMemcpyAsync(tt, tt_host, n_tok * 8, cudaMemcpyHostToDevice, st));
  MSQ_CUDA(cudaMemcpyAsync(mask, mask_host, n_tok * 8, cudaMemcpyHostToDevice, st));
  MSQ_CUDA(cudaMemcpyAsync(sep, sep_host, (size_t)R * 2 * 8, cudaMemcpyHostToDevice, st));
  const cudaEvent_t* ready = nullptr;
  int64_t up_chunk = 0;
  if (images_host) {
    MSQ_CUDA(cudaMemcpyAsync(idx, img_index_host, (size_t)R * 2 * 4, cudaMemcpyHostToDevice, st));
    // images in manual order (row b*N+i, every pair of manual b points into [b*N, (b+1)*N)): upload them micro-batch by
    // micro-batch on a copy stream so that only the first micro-batch's copy is exposed.
    const int P2 = N * (N - 1) * 2;
    bool local = n_img == B * N;
    for (int64_t i = 0; local && i < R * 2; ++i) {
      const int64_t b = i / P2;
      local = img_index_host[i] >= b * N && img_index_host[i] < (b + 1) * N;
    }
    cudaStream_t& copy_st = m->copy_st;
    std::vector<cudaEvent_t>& evs = m->copy_evs;
    const int64_t Bcm = min((int64_t)chunk_manuals(m), B);
    const int64_t Bc = (Bcm % 16 == 0 && B > 16) ? 16 : Bcm;   // upload granularity: 16 manuals when that divides the compute micro-batch
    up_chunk = Bc;
    const int64_t nchunks = (B + Bc - 1) / Bc;
    if (local && nchunks > 1) {
      if (!copy_st) MSQ_CUDA(cudaStreamCreateWithFlags(&copy_st, cudaStreamNonBlocking));
      while ((int64_t)evs.size() < nchunks) {
        cudaEvent_t e;
        MSQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        evs.push_back(e);
      }
      const size_t per_img = (size_t)3 * m->cfg.vit_res * m->cfg.vit_res;
      for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t i0 = c * Bc * N, n = min(Bc, B - c * Bc) * N;
        MSQ_CUDA(cudaMemcpyAsync(img + i0 * per_img, images_host + i0 * per_img, n * per_img * 4, cudaMemcpyHostToDevice, copy_st));
        MSQ_CUDA(cudaEventRecord(evs[c], copy_st));
      }
      ready = evs.data();
    } else {
      MSQ_CUDA(cudaMemcpyAsync(img, images_host, img_elems * 4, cudaMemcpyHostToDevice, st));
    }
  }
  auto go = [&]() -> int {
    MSQ_DISPATCH_T(m, run_path<T>(m, ids, tt, mask, sep, B, N, Lt, images_host ? img : nullptr, n_img, images_host ? idx : nullptr, nullptr,
                                  beam, perm, st, nullptr, nullptr, 0.f, nullptr, ready, up_chunk));
  };
  MSQ_TRY(go());
  MSQ_CUDA(cudaMemcpyAsync(perm_host, perm, (size_t)B * N * 4, cudaMemcpyDeviceToHost, st));
  MSQ_CUDA(cudaStreamSynchronize(st));
  return MSQ_OK;
}

// ---- device-side pair expansion (process_inputs_for_berson.py:113-368) ------------------------------------------------
// msq_scan_steps: [CLS] / [SEP] positions of every manual -> starts / lens tables (device, int32 [B,N]) and, after ONE
// 8-byte device->host read, the padded pair length Lt = max_{i != j}(len_i + len_j).  Fails (like the reference's
// assertion) when a manual does not hold exactly N [CLS] .. [SEP] steps.
extern "C" int msq_scan_steps(const int64_t* ids_dev, int64_t B, int32_t L, int32_t N, int64_t cls_id, int64_t sep_id,
                              int32_t* starts_dev, int32_t* lens_dev, int32_t* meta_dev, int32_t* lt_out, void* stream) {
  MSQ_REQUIRE(ids_dev && starts_dev && lens_dev && meta_dev && lt_out, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  MSQ_TRY(scan_steps(ids_dev, B, L, N, cls_id, sep_id, starts_dev, lens_dev, meta_dev, st));
  int32_t meta[2] = {0, 0};
  if (B > 0) {
    MSQ_CUDA(cudaMemcpyAsync(meta, meta_dev, sizeof(meta), cudaMemcpyDeviceToHost, st));
    MSQ_CUDA(cudaStreamSynchronize(st));
  }
  MSQ_REQUIRE(meta[1] == 0, "every manual must hold exactly max_story_length [CLS]..[SEP] steps");
  *lt_out = meta[0];
  return MSQ_OK;
}
extern "C" int msq_expand_pairs(const int64_t* ids_dev, int64_t B, int32_t L, int32_t N, int32_t Lt, int64_t cls_id, int64_t pad_id,
                                const int32_t* starts_dev, const int32_t* lens_dev, int64_t* out_ids_dev, int64_t* out_mask_dev,
                                int64_t* out_tt_dev, int64_t* out_sep_dev, int32_t* img_index_dev, void* stream) {
  MSQ_REQUIRE(ids_dev && starts_dev && lens_dev && out_ids_dev && out_mask_dev && out_tt_dev && out_sep_dev, "null argument");
  return expand_pairs(ids_dev, B, L, N, Lt, cls_id, pad_id, starts_dev, lens_dev, out_ids_dev, out_mask_dev, out_tt_dev, out_sep_dev,
                      img_index_dev, (cudaStream_t)stream);
}

// berson_pointer_network from the DataLoader tuple (modeling_bert.py:1405-1408 + process_inputs_for_berson.py): token rows
// [B, L] and step images [B*N, 3, S, S] in HOST memory (ideally pinned) -> predicted orders [B, N] in host memory.  The token
// rows are uploaded as they are (B*L*8 bytes), pairs are expanded on the device, images stream in micro-batch by
// micro-batch behind the compute.  Synchronous.
extern "C" int msq_order_manuals_raw_host(msq_model* m, const int64_t* ids_host, int64_t B, int32_t L, int32_t N, int64_t cls_id,
                                          int64_t sep_id, int64_t pad_id, const float* images_host, int32_t beam, int32_t* perm_host,
                                          void* stream) {
  DevGuard dev_guard__(m);
  MSQ_REQUIRE(m && m->packed && ids_host && perm_host, "bad argument");
  MSQ_REQUIRE(N >= 2 && N <= 16 && B >= 0, "N=%d out of range [2,16]", N);
  MSQ_REQUIRE(!(m->cfg.vit_width != 0 && images_host == nullptr), "multimodal model needs images");
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) return MSQ_OK;
  const int64_t R = B * N * (N - 1), n_img = images_host ? B * N : 0;
  const size_t per_img = (size_t)3 * m->cfg.vit_res * m->cfg.vit_res, img_elems = images_host ? (size_t)n_img * per_img : 0;
  const int Lt_cap = 2 * L;   // upper bound of the padded pair length (the true Lt is known after the scan)
  const size_t need = (size_t)B * L * 8 + 2 * (size_t)B * N * 4 + 256 + 3 * (size_t)R * Lt_cap * 8 + (size_t)R * 2 * 8 + img_elems * 4 +
                      (size_t)R * 2 * 4 + (size_t)B * N * 4 + 12 * 256;
  MSQ_TRY(stage_reserve(m, need, st));
  size_t off = 0;
  auto carve = [&](size_t bytes) { char* p = m->stage + off; off += (bytes + 255) & ~size_t(255); return p; };
  int64_t* raw = (int64_t*)carve((size_t)B * L * 8);
  int32_t* starts = (int32_t*)carve((size_t)B * N * 4);
  int32_t* lens = (int32_t*)carve((size_t)B * N * 4);
  int32_t* meta = (int32_t*)carve(256);
  float* img = (float*)carve(img_elems * 4);
  int32_t* idx = (int32_t*)carve((size_t)R * 2 * 4);
  int32_t* perm = (int32_t*)carve((size_t)B * N * 4);
  int64_t* sep = (int64_t*)carve((size_t)R * 2 * 8);
  char* toks = carve(3 * (size_t)R * Lt_cap * 8);
  // images first on the copy stream (they are the bulk of the bytes), one event per micro-batch
  const cudaEvent_t* ready = nullptr;
  const int64_t Bcm = min((int64_t)chunk_manuals(m), B);
  const int64_t Bc = (Bcm % 16 == 0 && B > 16) ? 16 : Bcm;   // upload granularity (divides the compute micro-batch)
  const int64_t nchunks = (B + Bc - 1) / Bc, up_chunk = Bc;
  if (images_host) {
    if (nchunks > 1) {
      if (!m->copy_st) MSQ_CUDA(cudaStreamCreateWithFlags(&m->copy_st, cudaStreamNonBlocking));
      while ((int64_t)m->copy_evs.size() < nchunks + 1) {
        cudaEvent_t e;
        MSQ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        m->copy_evs.push_back(e);
      }
      // the staging buffer may still be read by work enqueued on `st` by a previous call
      MSQ_CUDA(cudaEventRecord(m->copy_evs[nchunks], st));
      MSQ_CUDA(cudaStreamWaitEvent(m->copy_st, m->copy_evs[nchunks], 0));
      for (int64_t c = 0; c < nchunks; ++c) {
        const int64_t i0 = c * Bc * N, n = min(Bc, B - c * Bc) * N;
        MSQ_CUDA(cudaMemcpyAsync(img + i0 * per_img, images_host + i0 * per_img, n * per_img * 4, cudaMemcpyHostToDevice, m->copy_st));
        MSQ_CUDA(cudaEventRecord(m->copy_evs[c], m->copy_st));
      }
      ready = m->copy_evs.data();
    } else {
      MSQ_CUDA(cudaMemcpyAsync(img, images_host, img_elems * 4, cudaMemcpyHostToDevice, st));
    }
  }
  MSQ_CUDA(cudaMemcpyAsync(raw, ids_host, (size_t)B * L * 8, cudaMemcpyHostToDevice, st));
  int32_t Lt = 0;
  MSQ_TRY(msq_scan_steps(raw, B, L, N, cls_id, sep_id, starts, lens, meta, &Lt, stream));
  MSQ_REQUIRE(Lt >= 2 && Lt <= Lt_cap, "pair length %d out of range", Lt);
  int64_t* ids = (int64_t*)toks;
  int64_t* mask = ids + (size_t)R * Lt;
  int64_t* tt = mask + (size_t)R * Lt;
  MSQ_TRY(expand_pairs(raw, B, L, N, Lt, cls_id, pad_id, starts, lens, ids, mask, tt, sep, images_host ? idx : nullptr, st));
  auto go = [&]() -> int {
    MSQ_DISPATCH_T(m, run_path<T>(m, ids, tt, mask, sep, B, N, Lt, images_host ? img : nullptr, n_img, images_host ? idx : nullptr, nullptr,
                                  beam, perm, st, nullptr, nullptr, 0.f, nullptr, ready, up_chunk));
  };
  MSQ_TRY(go());
  MSQ_CUDA(cudaMemcpyAsync(perm_host, perm, (size_t)B * N * 4, cudaMemcpyDeviceToHost, st));
  MSQ_CUDA(cudaStreamSynchronize(st));
  return MSQ_OK;
}

// =====================================================================================================
// building blocks for kernel-level tests / roofline lines
// =====================================================================================================

extern "C" int msq_gemm(int32_t dtype, const void* A_dev, const void* W_dev, const float* bias_dev, const float* resid_dev,
                        void* C_dev, int64_t M, int32_t N, int32_t K, int32_t act, void* stream) {
  GemmArgs g;
  g.A = A_dev; g.W = W_dev; g.bias = bias_dev; g.resid = resid_dev; g.C = C_dev; g.C2 = nullptr;
  g.M = M; g.N = N; g.K = K; g.lda = K; g.ldw = K; g.ldc = N; g.ldr = N; g.act = act;
  cudaStream_t st = (cudaStream_t)stream;
  switch (dtype) {
    case 0: return gemm_simt<float, float>(g, st);
    case 1: return gemm_tc<float>(g, st);
    case 2: return gemm_tc<bf16>(g, st);
    case 3: return gemm_simt<bf16, float>(g, st);
    case 4: return gemm_simt<bf16, bf16>(g, st);
    case 5:   // TN operands (weight-gradient shape): A [K, M], W [K, N] bf16 row-major -> C [M, N] fp32 = A^T W (+ resid)
      g.tn = 1; g.lda = (int)M; g.ldw = N;
      return gemm_tc<float>(g, st);
    case 6: g.split = 1; return gemm_tc<float>(g, st);   // bf16x3: split-bf16 A [M, hi(K)|lo(K)] and W, fp32 out
    case 7: g.split = 1; return gemm_tc<bf16s>(g, st);   // bf16x3: split-bf16 out [M, hi(N)|lo(N)]
  }
  set_error("msq_gemm: unknown dtype %d", dtype);
  return MSQ_ERR_ARG;
}

extern "C" int msq_gemm_deferred_ln(int32_t mode, int32_t out_bf16, const void* A_dev, const void* W_dev, const float* bias_dev,
                                    const float* resid_dev, const float* svec_or_gamma_dev, const float* beta_dev,
                                    const float* stats_in_dev, int32_t sp_in, int32_t ln_dim, float eps, void* C_dev, void* C2bf_dev,
                                    float* stats_out_dev, int64_t M, int32_t N, int32_t K, int32_t act, void* stream) {
  MSQ_REQUIRE(mode == EPI_LNFOLD || mode == EPI_RESLN || mode == EPI_DUALACT, "msq_gemm_deferred_ln: mode %d", mode);
  GemmArgs g;
  g.A = A_dev; g.W = W_dev; g.bias = bias_dev; g.resid = resid_dev; g.C = C_dev; g.C2 = nullptr;
  g.M = M; g.N = N; g.K = K; g.lda = K; g.ldw = K; g.ldc = N; g.ldr = N; g.act = act;
  g.mode = mode; g.svec = svec_or_gamma_dev; g.beta = beta_dev; g.stats_in = stats_in_dev; g.sp_in = sp_in;
  g.ln_inv_dim = ln_dim > 0 ? 1.0f / (float)ln_dim : 0.f; g.ln_eps = eps; g.stats_out = stats_out_dev; g.C2bf = C2bf_dev;
  return out_bf16 ? gemm_tc<bf16>(g, (cudaStream_t)stream) : gemm_tc<float>(g, (cudaStream_t)stream);
}

extern "C" int msq_gemm_ln(const void* A_dev, const void* W_dev, const float* bias_dev, const float* resid_dev,
                           const float* gamma_dev, const float* beta_dev, float eps, float* C_dev, void* C2_dev, int64_t M, int32_t N,
                           int32_t K, int32_t raw32, void* stream) {
  return gemm_ln((const bf16*)A_dev, K, (const bf16*)W_dev, K, bias_dev, resid_dev, N, gamma_dev, beta_dev, eps, C_dev, (bf16*)C2_dev, M, N,
                 K, raw32 != 0, (cudaStream_t)stream);
}

extern "C" int msq_layernorm(int32_t dtype, const float* x_dev, int64_t rows, int32_t H, const float* gamma_dev,
                             const float* beta_dev, float eps, void* out_dev, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 0) return layernorm<float>(x_dev, rows, H, gamma_dev, beta_dev, eps, (float*)out_dev, nullptr, 0, 0, 0, st);
  if (dtype == 2) return layernorm<bf16s>(x_dev, rows, H, gamma_dev, beta_dev, eps, nullptr, (bf16s*)out_dev, 0, 0, 0, st);
  return layernorm<bf16>(x_dev, rows, H, gamma_dev, beta_dev, eps, nullptr, (bf16*)out_dev, 0, 0, 0, st);
}

extern "C" int msq_attention(int32_t dtype, const void* qkv_dev, int64_t R, int32_t L, int32_t heads, float scale,
                             const float* mask_add_dev, int32_t mask_len, void* ctx_dev, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 0)
    return attention<float>((const float*)qkv_dev, R, L, heads, 64, scale, mask_add_dev, mask_len, mask_len, (float*)ctx_dev, st);
  if (dtype == 2)
    return attention<bf16s>((const bf16s*)qkv_dev, R, L, heads, 64, scale, mask_add_dev, mask_len, mask_len, (bf16s*)ctx_dev, st);
  return attention<bf16>((const bf16*)qkv_dev, R, L, heads, 64, scale, mask_add_dev, mask_len, mask_len, (bf16*)ctx_dev, st);
}

extern "C" int msq_f32_to_bf16(const float* src_dev, void* dst_dev, int64_t n, void* stream) {
  if (n == 0) return MSQ_OK;
  MSQ_CUDA(launch_k(f32_to_bf16_kernel, dim3((int)min((int64_t)148 * 8, (n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, src_dev, (bf16*)dst_dev, n));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

// fp32 [rows, K] -> split bf16 rows [hi(K) | lo(K)] (the operand layout of the bf16x3 mode; msq_gemm dtype 6 / 7)
extern "C" int msq_f32_to_bf16_split(const float* src_dev, void* dst_dev, int64_t rows, int32_t K, void* stream) {
  MSQ_REQUIRE(K % 4 == 0, "msq_f32_to_bf16_split: K=%d", K);
  return gather_rows<float, bf16s>(src_dev, rows, K, 0, 0, 0, (bf16s*)dst_dev, (cudaStream_t)stream);
}
