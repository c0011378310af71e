// Fine-tuning through the CLIP ModifiedResNet tower ("RN50", the reference's wired backbone): training-mode forward with
// BatchNorm over the batch, and the backward pass down to the stem.
//
// Reference (autograd through): Bottleneck.forward models/CLIP/clip/model.py:10-53, ModifiedResNet.forward 128-187 (3-conv stem,
// AvgPool2d(2), four stages), AttentionPool2d.forward 71-125, LinearPositionEmbedding / VisualTokenTypeEmbedding
// models/CLIP/src/lxrt/modeling.py:621-705; nn.BatchNorm2d in train() mode = statistics of the batch (mean, biased variance
// over N, H, W); the backbone is the one models/CLIP/src/param.py:248-249 selects.
//
// Design.  Activations are NHWC rows [n_img * H * W, C] in the GEMM operand type (pre-BatchNorm outputs stay fp32: the batch-
// statistic backward is a difference of large sums), so every convolution is a GEMM on the same
// kernels as the transformer layers (1x1: the rows themselves; 3x3: im2col rows, recomputed in the backward pass rather than
// kept): forward Y = A Wg^T, weight gradient dWg += dY^T A (TN operands, split-K), input gradient dA = dY Wg followed by a
// gather-form col2im.  The tower runs on the UNIQUE images of the batch while the reference materialises every image once per
// ordered pair it belongs to: with w_i = the number of pair slots that reference image i, the batch statistics are the
// w-weighted ones and, because the upstream gradient of a unique image already is the sum over its copies,
//     dY_u = gamma rstd (dz_u - (w_u / W) (S1 + xhat_u S2)),   S1 = sum dz, S2 = sum dz xhat (plain sums),  W = sum_i w_i H W
// is exactly the sum of the materialised copies' gradients.  Column reductions are two-stage with a fixed order (deterministic).
// BatchNorm running statistics follow every training forward (momentum 0.1, unbiased batch variance) like nn.BatchNorm2d.train().
#include <type_traits>

#include "train_common.cuh"

namespace msq {

namespace {

inline dim3 ew_grid(int64_t work) { return dim3((unsigned)max((int64_t)1, min((int64_t)148 * 16, (work + 255) / 256))); }

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }

// ---- conv weight [Cout, Cin, k, k] -> GEMM layouts: wg [Cout, Kp] ((ky, kx, c) columns, zero padded) and wgT [Kp, Cout]
template <typename T>
__global__ void rn_wreorder_kernel(const float* __restrict__ w, int Cout, int Cin, int k, int Kp, T* __restrict__ wg, T* __restrict__ wgT) {
  pdl_sync();
  const int K = k * k * Cin;
  const int64_t total = (int64_t)Cout * Kp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i / Kp), col = (int)(i % Kp);
    float v = 0.f;
    if (col < K) {
      const int c = col % Cin, kk = col / Cin, ky = kk / k, kx = kk % k;
      v = w[(((int64_t)o * Cin + c) * k + ky) * k + kx];
    }
    wg[i] = from_f<T>(v);
    wgT[(int64_t)col * Cout + o] = from_f<T>(v);
  }
}

// dW[o, c, ky, kx] += s[o, (ky*k + kx)*Cin + c]
__global__ void rn_wgrad_permute_kernel(const float* __restrict__ s, int Cout, int Cin, int k, int Kp, float* __restrict__ dw) {
  pdl_sync();
  const int64_t total = (int64_t)Cout * Cin * k * k;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kx = (int)(i % k), ky = (int)((i / k) % k), c = (int)((i / (k * k)) % Cin), o = (int)(i / ((int64_t)k * k * Cin));
    dw[i] += s[(int64_t)o * Kp + (ky * k + kx) * Cin + c];
  }
}

// ---- per-channel reductions over the rows of an NHWC matrix.  MODE 0: sum w y;  1: sum w (y - mean)^2;
//      2: (sum dz, sum dz xhat) with dz = dout * [act > 0].  Stage 1 writes partial[block][out][C], stage 2 adds them in order.
constexpr int RED_BLOCKS = 148 * 2;
template <typename T, int MODE>
__global__ void __launch_bounds__(256) rn_colred_kernel(const float* __restrict__ y, const float* __restrict__ dout, const T* __restrict__ act,
                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        const float* __restrict__ wimg, int HW, int64_t M, int C, float* __restrict__ partial) {
  pdl_sync();
  constexpr int NO = MODE == 2 ? 2 : 1;
  __shared__ float sh[NO][256];
  const int64_t rpb = (M + gridDim.x - 1) / gridDim.x, r0 = (int64_t)blockIdx.x * rpb, r1 = min(M, r0 + rpb);
  const int lanes = C <= 256 ? 256 / C : 1;        // rows processed side by side
  const int tid = threadIdx.x;
  for (int c0 = 0; c0 < C; c0 += 256) {            // channel chunk (one pass when C <= 256)
    const int c = C <= 256 ? tid % C : c0 + tid, rl = C <= 256 ? tid / C : 0;
    const bool on = c < C && rl < lanes;
    float a0 = 0.f, a1 = 0.f;
    if (on) {
      // fp32 parity mode: double accumulators (the BatchNorm backward subtracts these sums from numbers of their own size)
      typedef typename std::conditional<sizeof(T) == 4, double, float>::type Acc;
      Acc s0 = 0, s1 = 0;
      const float mu = MODE >= 1 ? mean[c] : 0.f, rs = MODE == 2 ? rstd[c] : 0.f;
      for (int64_t r = r0 + rl; r < r1; r += lanes) {
        const float v = y[r * C + c];
        if (MODE == 0) {
          s0 += (Acc)((wimg ? wimg[r / HW] : 1.f) * v);
        } else if (MODE == 1) {
          const float d = v - mu;
          s0 += (Acc)((wimg ? wimg[r / HW] : 1.f) * d) * (Acc)d;
        } else {
          float dz = dout[r * C + c];
          if (act && !(ldf<T>(act + r * C + c) > 0.f)) dz = 0.f;
          s0 += (Acc)dz;
          s1 += (Acc)dz * (Acc)((v - mu) * rs);
        }
      }
      a0 = (float)s0; a1 = (float)s1;
    }
    if (C <= 256 && lanes > 1) {                    // add the row lanes of a channel (fixed order)
      __syncthreads();
      sh[0][tid] = a0;
      if (NO == 2) sh[1][tid] = a1;
      __syncthreads();
      if (tid < C) {
        float s0 = 0.f, s1 = 0.f;
        for (int l = 0; l < lanes; ++l) { s0 += sh[0][l * C + tid]; if (NO == 2) s1 += sh[1][l * C + tid]; }
        a0 = s0; a1 = s1;
      }
    }
    if (c < C && rl == 0) {
      partial[((int64_t)blockIdx.x * NO + 0) * C + c] = a0;
      if (NO == 2) partial[((int64_t)blockIdx.x * NO + 1) * C + c] = a1;
    }
  }
}
// stage 2.  MODE 0: mean = S / W;  1: rstd = rsqrt(S / W + eps);  2: o0 = S1, o1 = S2 and dgamma += S2, dbeta += S1
template <int MODE>
__global__ void rn_colred_final_kernel(const float* __restrict__ partial, int nblk, int C, float Wtot, float eps, float* __restrict__ o0,
                                       float* __restrict__ o1, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  pdl_sync();
  constexpr int NO = MODE == 2 ? 2 : 1;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int b = 0; b < nblk; ++b) {
    s0 += (double)partial[((int64_t)b * NO + 0) * C + c];
    if (NO == 2) s1 += (double)partial[((int64_t)b * NO + 1) * C + c];
  }
  if (MODE == 0) o0[c] = (float)(s0 / (double)Wtot);
  else if (MODE == 1) o0[c] = (float)(1.0 / sqrt(s0 / (double)Wtot + (double)eps));
  else {
    o0[c] = (float)s0; o1[c] = (float)s1;
    if (dgamma) dgamma[c] += (float)s1;
    if (dbeta) dbeta[c] += (float)s0;
  }
}
// evaluation-mode BatchNorm inside a training step (optional): mean / rstd from the running statistics
__global__ void rn_running_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, int C, float eps, float* __restrict__ mean,
                                        float* __restrict__ rstd) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) { mean[c] = rm[c]; rstd[c] = rsqrtf(rv[c] + eps); }
}

// nn.BatchNorm2d train() bookkeeping: running = (1 - momentum) running + momentum batch (UNBIASED batch variance, n = W)
__global__ void rn_running_update_kernel(const float* __restrict__ mean, const float* __restrict__ rstd, int C, float eps, float Wtot, float mom,
                                         float* __restrict__ rm, float* __restrict__ rv) {
  pdl_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float var = 1.0f / (rstd[c] * rstd[c]) - eps;
  rm[c] = (1.f - mom) * rm[c] + mom * mean[c];
  rv[c] = (1.f - mom) * rv[c] + mom * var * (Wtot / fmaxf(Wtot - 1.f, 1.f));
}

// ---- out = act( bn_a(ya) + [bn_b(yb) | xres] ), four channels per thread
template <typename T>
__global__ void __launch_bounds__(256) rn_bn_apply_kernel(const float* __restrict__ ya, const float* __restrict__ ma, const float* __restrict__ ra,
                                                          const float* __restrict__ ga, const float* __restrict__ ba, const float* __restrict__ yb,
                                                          const float* __restrict__ mb, const float* __restrict__ rb, const float* __restrict__ gb,
                                                          const float* __restrict__ bb, const T* __restrict__ xres, int relu, int64_t M, int C,
                                                          T* __restrict__ out, float* __restrict__ outf) {
  pdl_sync();
  const int C4 = C / 4;
  const int64_t total = M * C4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4) * 4;
    const int64_t e = (i / C4) * C + c;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float s = ga[c + j] * ra[c + j];
      v[j] = fmaf(ya[e + j] - ma[c + j], s, ba[c + j]);
      if (yb) v[j] += fmaf(yb[e + j] - mb[c + j], gb[c + j] * rb[c + j], bb[c + j]);
      else if (xres) v[j] += ldf<T>(xres + e + j);
      if (relu) v[j] = fmaxf(v[j], 0.f);
      out[e + j] = from_f<T>(v[j]);
    }
    if (outf) *reinterpret_cast<float4*>(outf + e) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

// ---- dY = gamma rstd (dz - wfac (S1 + xhat S2)),  dz = dout [act > 0],  wfac = w_img / W (0 with running statistics)
template <typename T>
__global__ void __launch_bounds__(256) rn_bn_bwd_apply_kernel(const float* __restrict__ dout, const T* __restrict__ act, const float* __restrict__ y,
                                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              const float* __restrict__ gamma, const float* __restrict__ S1,
                                                              const float* __restrict__ S2, const float* __restrict__ wimg, int HW, float invW,
                                                              int64_t M, int C, T* __restrict__ dy) {
  pdl_sync();
  const int64_t total = M * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t r = i / C;
    float dz = dout[i];
    if (act && !(ldf<T>(act + i) > 0.f)) dz = 0.f;
    const float rs = rstd[c], xh = (y[i] - mean[c]) * rs;
    const float wf = invW > 0.f ? (wimg ? wimg[r / HW] : 1.f) * invW : 0.f;
    dy[i] = from_f<T>(gamma[c] * rs * (dz - wf * fmaf(xh, S2[c], S1[c])));
  }
}

// ---- gather-form col2im of a 3x3 / stride 1 / pad 1 convolution: dx[(im,y,x), c] = sum_taps dcol[(im, y-ky+1, x-kx+1), (ky,kx,c)]
template <typename T>
__global__ void __launch_bounds__(256) rn_col2im3_kernel(const T* __restrict__ dcol, int64_t n, int H, int W, int C, int Kp, float* __restrict__ dx) {
  pdl_sync();
  const int64_t total = n * H * W * (int64_t)C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t pix = i / C, im = pix / (H * W);
    const int yy = (int)(pix % (H * W)) / W, xx = (int)(pix % (H * W)) % W;
    float s = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int oy = yy - ky + 1;
      if (oy < 0 || oy >= H) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int ox = xx - kx + 1;
        if (ox < 0 || ox >= W) continue;
        s += ldf<T>(dcol + ((im * H + oy) * W + ox) * (int64_t)Kp + (ky * 3 + kx) * C + c);
      }
    }
    dx[i] = s;
  }
}

// ---- AvgPool2d(2) backward: dx[(im, 2oy+a, 2ox+b), c] (+)= dout[(im, oy, ox), c] / 4
__global__ void __launch_bounds__(256) rn_avgpool2_bwd_kernel(const float* __restrict__ dout, int64_t n, int H, int W, int C, int accum,
                                                              float* __restrict__ dx) {
  pdl_sync();
  const int64_t total = n * H * W * (int64_t)C;
  const int Ho = H / 2, Wo = W / 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t pix = i / C, im = pix / (H * W);
    const int yy = (int)(pix % (H * W)) / W, xx = (int)(pix % (H * W)) % W;
    const float v = 0.25f * dout[((im * Ho + yy / 2) * Wo + xx / 2) * (int64_t)C + c];
    dx[i] = accum ? dx[i] + v : v;
  }
}

// ---- s (+)= dout * [act > 0]   (identity shortcut of a block without a down-sampling branch)
template <typename T>
__global__ void __launch_bounds__(256) rn_add_masked_kernel(const float* __restrict__ dout, const T* __restrict__ act, int64_t n, int accum,
                                                            float* __restrict__ s) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = ldf<T>(act + i) > 0.f ? dout[i] : 0.f;
    s[i] = accum ? s[i] + v : v;
  }
}
__global__ void __launch_bounds__(256) rn_add_kernel(const float* __restrict__ a, int64_t n, float* __restrict__ s) {
  pdl_sync();
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s[i] += a[i];
}

// ---- multiplicity of every unique image among the pair slots
__global__ void rn_img_weight_kernel(const int32_t* __restrict__ img_index, int64_t slots, int64_t n_img, float* __restrict__ w) {
  pdl_sync();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_img) return;
  int cnt = 0;
  for (int64_t s = 0; s < slots; ++s) cnt += img_index[s] == (int32_t)i;
  w[i] = (float)cnt;
}

// ---- AttentionPool2d token matrix backward (forward: rn_tokens_kernel, reshape quirk included).  dfeat[i, p, c] = sum over
// the pair slots (r, s) holding image i of dv(r, t, c') with flat f = s*C*g2 + c*g2 + p, c' = f / (2 g2), t = f % (2 g2),
// dv(r, t, c') = dtok[r, 1+t, c'] + dtok[r, 0, c'] / (2 g2)
__global__ void __launch_bounds__(256) rn_tokens_bwd_kernel(const float* __restrict__ dtok, const int32_t* __restrict__ img_index, int64_t R,
                                                            int64_t n_img, int g2, int C, float* __restrict__ dfeat) {
  pdl_sync();
  const int L2 = 2 * g2;
  const float invL = 1.0f / (float)L2;
  const int64_t total = n_img * g2 * (int64_t)C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), p = (int)((i / C) % g2);
    const int32_t img = (int32_t)(i / ((int64_t)C * g2));
    float s = 0.f;
    for (int64_t r = 0; r < R; ++r) {
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        if (img_index[r * 2 + sl] != img) continue;
        const int f = sl * C * g2 + c * g2 + p, cp = f / L2, t = f % L2;
        const float* d = dtok + r * (int64_t)(1 + L2) * C + cp;
        s += d[(int64_t)(1 + t) * C] + d[0] * invL;
      }
    }
    dfeat[i] = s;
  }
}
// dpos[j, :] += sum_r ( dtok[r, j, :] + [j < g2] dtok[r, 1 + g2 + j, :] )   (token 1+t reads pos[1+t] for t < g2, pos[t-g2] beyond)
__global__ void __launch_bounds__(256) rn_pos_bwd_kernel(const float* __restrict__ dtok, int64_t R, int g2, int C, float* __restrict__ dpos) {
  pdl_sync();
  const int L = 1 + 2 * g2;
  const int64_t total = (int64_t)(g2 + 1) * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C), j = (int)(i / C);
    float s = 0.f;
    for (int64_t r = 0; r < R; ++r) {
      const float* d = dtok + r * (int64_t)L * C + c;
      s += d[(int64_t)j * C];
      if (j < g2) s += d[(int64_t)(1 + g2 + j) * C];
    }
    dpos[i] += s;
  }
}

// ---- tower output backward: y = cat(o, o) + posadd  ->  do = dy[:, :E] + dy[:, E:]  (operand type)
template <typename T>
__global__ void __launch_bounds__(256) rn_finish_bwd_kernel(const float* __restrict__ dy, int64_t rows, int E, T* __restrict__ dout) {
  pdl_sync();
  const int64_t total = rows * E;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / E;
    const int e = (int)(i % E);
    dout[i] = from_f<T>(dy[r * 2 * E + e] + dy[r * 2 * E + E + e]);
  }
}
// dposadd[t, f] = sum_r dy[r*L + t, f]
__global__ void __launch_bounds__(256) rn_posadd_sum_kernel(const float* __restrict__ dy, int64_t R, int L, int F, float* __restrict__ dpa) {
  pdl_sync();
  const int64_t total = (int64_t)L * F;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int64_t r = 0; r < R; ++r) s += dy[r * total + i];
    dpa[i] = s;
  }
}
// posadd[t] = xe[cell / g] + ye[cell % g] + te[t > g2], cell = (t == 0 ? 0 : (t - 1) % g2): scatter dposadd back to the tables
__global__ void __launch_bounds__(256) rn_posadd_bwd_kernel(const float* __restrict__ dpa, int g, int F, float* __restrict__ dxe,
                                                            float* __restrict__ dye, float* __restrict__ dte) {
  pdl_sync();
  const int g2 = g * g, L = 1 + 2 * g2;
  const int64_t total = (int64_t)(2 * g + 2) * F;   // rows: g of xe, g of ye, 2 of te
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % F), row = (int)(i / F);
    float s = 0.f;
    for (int t = 0; t < L; ++t) {
      const int cell = t == 0 ? 0 : (t - 1) % g2;
      const bool hit = row < g ? cell / g == row : (row < 2 * g ? cell % g == row - g : (t > g2 ? 1 : 0) == row - 2 * g);
      if (hit) s += dpa[(int64_t)t * F + f];
    }
    if (row < g) dxe[(int64_t)row * F + f] += s;
    else if (row < 2 * g) dye[(int64_t)(row - g) * F + f] += s;
    else dte[(int64_t)(row - 2 * g) * F + f] += s;
  }
}

// dX[M, Kin] = G[M, Nout] W for contractions the tiled GEMMs do not take (Nout % 16 != 0: test-sized towers only)
template <typename T, typename TO>
__global__ void __launch_bounds__(256) rn_dgrad_naive_kernel(const T* __restrict__ G, const T* __restrict__ WT, int64_t M, int Nout, int Kin,
                                                             TO* __restrict__ dX) {
  pdl_sync();
  const int64_t total = M * Kin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kcol = (int)(i % Kin);
    const int64_t r = i / Kin;
    float s = 0.f;
    for (int n = 0; n < Nout; ++n) s = fmaf(ldf<T>(G + r * Nout + n), ldf<T>(WT + (int64_t)kcol * Nout + n), s);
    dX[i] = from_f<TO>(s);
  }
}
template <typename T, typename TO>
int dgrad_any(const msq_model* m, const T* G, int Nout, const void* WT, int Kin, TO* dX, int64_t M, cudaStream_t st) {
  if (Nout % 16 == 0) return dgrad<T, TO>(m, G, Nout, WT, Kin, nullptr, dX, M, st);
  MSQ_CUDA(launch_k(rn_dgrad_naive_kernel<T, TO>, ew_grid(M * Kin), dim3(256), 0, st, G, (const T*)WT, M, Nout, Kin, dX));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

template <typename T, int MODE>
int colred(const float* y, const float* dout, const T* act, const float* mean, const float* rstd, const float* wimg, int HW, int64_t M, int C,
           float Wtot, float* partial, float* o0, float* o1, float* dgamma, float* dbeta, cudaStream_t st) {
  const int nblk = (int)max((int64_t)1, min((int64_t)RED_BLOCKS, (M + 63) / 64));
  MSQ_CUDA(launch_k(rn_colred_kernel<T, MODE>, dim3(nblk), dim3(256), 0, st, y, dout, act, mean, rstd, wimg, HW, M, C, partial));
  MSQ_LAUNCH_CHECK();
  MSQ_CUDA(launch_k(rn_colred_final_kernel<MODE>, dim3(ceil_div(C, 128)), dim3(128), 0, st, (const float*)partial, nblk, C, Wtot, 1e-5f, o0, o1,
                    dgamma, dbeta));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

int rn_kpad_train(int K) { return K < 64 ? (K + 15) / 16 * 16 : (K + 63) / 64 * 64; }   // same rule as the packed inference weights

}  // namespace

// =====================================================================================================
// parameter table / operand copies
// =====================================================================================================
std::vector<std::string> rn_param_names(const msq_model* m) {
  const msq_config& c = m->cfg;
  const std::string P = m->prefix_inner, v = P + "encoder.visual_model.visual.";
  std::vector<std::string> out;
  for (int i = 1; i <= 3; ++i)
    for (const char* e : {"conv%d.weight", "bn%d.weight", "bn%d.bias"}) {
      char buf[64];
      snprintf(buf, sizeof(buf), e, i);
      out.push_back(v + buf);
    }
  for (int s = 0; s < 4; ++s)
    for (int b = 0; b < c.rn_blocks[s]; ++b) {
      const std::string k = v + "layer" + std::to_string(s + 1) + "." + std::to_string(b) + ".";
      for (const char* e : {"conv1.weight", "bn1.weight", "bn1.bias", "conv2.weight", "bn2.weight", "bn2.bias", "conv3.weight", "bn3.weight", "bn3.bias"})
        out.push_back(k + e);
      if (b == 0)
        for (const char* e : {"downsample.0.weight", "downsample.1.weight", "downsample.1.bias"}) out.push_back(k + e);
    }
  const std::string a = v + "attnpool.";
  // q / k / v adjacent: the fused [3 Cf, Cf] gradient block of the packed projection is contiguous
  for (const char* e : {"positional_embedding", "q_proj.weight", "k_proj.weight", "v_proj.weight", "q_proj.bias", "k_proj.bias", "v_proj.bias",
                        "c_proj.weight", "c_proj.bias"})
    out.push_back(a + e);
  for (const char* e : {"encoder.visual_pos.x_position_embedding.weight", "encoder.visual_pos.y_position_embedding.weight",
                        "encoder.visual_token_type.token_type_embedding.weight", "encoder.visn_fc.visn_fc.weight", "encoder.visn_fc.visn_fc.bias",
                        "encoder.visn_fc.visn_layer_norm.weight", "encoder.visn_fc.visn_layer_norm.bias"})
    out.push_back(P + e);
  return out;
}

void rn_train_free(RnTrain* r) {
  if (!r) return;
  for (void* p : r->owned) cudaFree(p);
  if (r->tape.base) cudaFree(r->tape.base);
  if (r->ws.base) cudaFree(r->ws.base);
  delete r;
}

template <typename T> static int rn_add_conv(msq_model* m, RnTrain* r, const std::string& wname, const std::string& bn, int cout, int cin, int k) {
  RnConvT cv;
  cv.wname = wname; cv.bn = bn; cv.cout = cout; cv.cin = cin; cv.k = k; cv.K = k * k * cin; cv.Kp = rn_kpad_train(cv.K);
  MSQ_REQUIRE(k != 1 || cv.Kp == cv.K, "ResNet tower: 1x1 convolution with %d input channels is not GEMM-aligned", cin);
  auto raw = [&](const std::string& n, int64_t numel) -> const float* {
    auto it = m->raw.find(n);
    if (it == m->raw.end() || it->second.second != numel) { set_error("train: weight %s missing or mis-sized", n.c_str()); return nullptr; }
    return it->second.first;
  };
  cv.w = raw(wname, (int64_t)cout * cv.K); cv.gamma = raw(bn + ".weight", cout); cv.beta = raw(bn + ".bias", cout);
  cv.rmean = raw(bn + ".running_mean", cout); cv.rvar = raw(bn + ".running_var", cout);
  if (!cv.w || !cv.gamma || !cv.beta || !cv.rmean || !cv.rvar) return MSQ_ERR_WEIGHT;
  void *wg = nullptr, *wgT = nullptr;
  float* st = nullptr;
  MSQ_CUDA(cudaMalloc(&wg, (size_t)cout * cv.Kp * sizeof(T)));
  r->owned.push_back(wg);
  MSQ_CUDA(cudaMalloc(&wgT, (size_t)cout * cv.Kp * sizeof(T)));
  r->owned.push_back(wgT);
  MSQ_CUDA(cudaMalloc(&st, (size_t)cout * 4 * sizeof(float)));
  r->owned.push_back(st);
  cv.wg = wg; cv.wgT = wgT; cv.mean = st; cv.rstd = st + cout; cv.S1 = st + 2 * cout; cv.S2 = st + 3 * cout;
  r->convs.push_back(cv);
  return MSQ_OK;
}

template <typename T> int rn_train_refresh(msq_model* m, cudaStream_t st) {
  RnTrain* r = m->train->rn;
  if (!r) return MSQ_OK;
  for (RnConvT& cv : r->convs) {
    MSQ_CUDA(launch_k(rn_wreorder_kernel<T>, ew_grid((int64_t)cv.cout * cv.Kp), dim3(256), 0, st, cv.w, cv.cout, cv.cin, cv.k, cv.Kp, (T*)cv.wg,
                      (T*)cv.wgT));
    MSQ_LAUNCH_CHECK();
  }
  MSQ_TRY((transpose_pad<float, T>(m->rn_qkv.w32, m->rn_qkv.N, m->rn_qkv.K, m->rn_qkv.ld, m->rn_qkv.N, (T*)r->qkvT, ACT_NONE, st)));
  MSQ_TRY((transpose_pad<float, T>(m->rn_cproj.w32, m->rn_cproj.N, m->rn_cproj.K, m->rn_cproj.ld, m->rn_cproj.N, (T*)r->cprojT, ACT_NONE, st)));
  return MSQ_OK;
}
template int rn_train_refresh<float>(msq_model*, cudaStream_t);
template int rn_train_refresh<bf16>(msq_model*, cudaStream_t);

template <typename T> int rn_train_build(msq_model* m, cudaStream_t st) {
  const msq_config& c = m->cfg;
  TrainState* ts = m->train;
  RnTrain* r = new RnTrain();
  ts->rn = r;
  const std::string v = m->prefix_inner + "encoder.visual_model.visual.";
  const int w = c.rn_width;
  MSQ_TRY(rn_add_conv<T>(m, r, v + "conv1.weight", v + "bn1", w / 2, 3, 3));
  MSQ_TRY(rn_add_conv<T>(m, r, v + "conv2.weight", v + "bn2", w / 2, w / 2, 3));
  MSQ_TRY(rn_add_conv<T>(m, r, v + "conv3.weight", v + "bn3", w, w / 2, 3));
  int cin = w;
  for (int s = 0; s < 4; ++s)
    for (int b = 0; b < c.rn_blocks[s]; ++b) {
      const std::string k = v + "layer" + std::to_string(s + 1) + "." + std::to_string(b) + ".";
      const int p = w << s;
      MSQ_TRY(rn_add_conv<T>(m, r, k + "conv1.weight", k + "bn1", p, cin, 1));
      MSQ_TRY(rn_add_conv<T>(m, r, k + "conv2.weight", k + "bn2", p, p, 3));
      MSQ_TRY(rn_add_conv<T>(m, r, k + "conv3.weight", k + "bn3", 4 * p, p, 1));
      if (b == 0) MSQ_TRY(rn_add_conv<T>(m, r, k + "downsample.0.weight", k + "downsample.1", 4 * p, cin, 1));
      cin = 4 * p;
    }
  void* p = nullptr;
  MSQ_CUDA(cudaMalloc(&p, (size_t)m->rn_qkv.N * m->rn_qkv.K * sizeof(T)));
  r->owned.push_back(p); r->qkvT = p;
  MSQ_CUDA(cudaMalloc(&p, (size_t)m->rn_cproj.N * m->rn_cproj.K * sizeof(T)));
  r->owned.push_back(p); r->cprojT = p;
  return rn_train_refresh<T>(m, st);
}
template int rn_train_build<float>(msq_model*, cudaStream_t);
template int rn_train_build<bf16>(msq_model*, cudaStream_t);

// =====================================================================================================
// forward
// =====================================================================================================
namespace {

struct Shape { int64_t M; int H; };   // rows and spatial side of an activation

// conv (GEMM) + batch statistics; the normalisation itself is applied by the caller (it may merge two branches)
template <typename T>
int conv_stats(msq_model* m, RnTrain* r, RnConvT& cv, const T* A, int64_t M, int HW, float Wtot, cudaStream_t st) {
  cv.M = M; cv.HW = HW; cv.Wtot = Wtot;
  MSQ_TRY((gemm_nt<T, float>(m, A, cv.Kp, (const T*)cv.wg, cv.Kp, nullptr, nullptr, 0, cv.Y, cv.cout, M, cv.cout, cv.Kp, ACT_NONE, st)));
  if (r->bn_eval) {
    MSQ_CUDA(launch_k(rn_running_stats_kernel, dim3(ceil_div(cv.cout, 128)), dim3(128), 0, st, cv.rmean, cv.rvar, cv.cout, 1e-5f, cv.mean, cv.rstd));
    MSQ_LAUNCH_CHECK();
    return MSQ_OK;
  }
  MSQ_TRY((colred<T, 0>((const float*)cv.Y, nullptr, nullptr, nullptr, nullptr, r->wimg, HW, M, cv.cout, Wtot, r->partial, cv.mean, nullptr, nullptr, nullptr, st)));
  MSQ_TRY((colred<T, 1>((const float*)cv.Y, nullptr, nullptr, cv.mean, nullptr, r->wimg, HW, M, cv.cout, Wtot, r->partial, cv.rstd, nullptr, nullptr, nullptr, st)));
  if (r->bn_momentum > 0.f) {   // the module's running statistics follow the batch, as nn.BatchNorm2d.train() does (the fp32 masters are updated
                                // in place; the packed evaluation weights pick them up at the next re-pack)
    MSQ_CUDA(launch_k(rn_running_update_kernel, dim3(ceil_div(cv.cout, 128)), dim3(128), 0, st, (const float*)cv.mean, (const float*)cv.rstd, cv.cout,
                      1e-5f, Wtot, r->bn_momentum, const_cast<float*>(cv.rmean), const_cast<float*>(cv.rvar)));
    MSQ_LAUNCH_CHECK();
  }
  return MSQ_OK;
}
template <typename T>
int bn_apply(const RnConvT& a, const RnConvT* b, const T* xres, bool relu, T* out, float* outf, cudaStream_t st) {
  MSQ_REQUIRE(a.cout % 4 == 0, "ResNet tower: channel count %d", a.cout);
  MSQ_CUDA(launch_k(rn_bn_apply_kernel<T>, ew_grid(a.M * (a.cout / 4)), dim3(256), 0, st, (const float*)a.Y, (const float*)a.mean, (const float*)a.rstd,
                    a.gamma, a.beta, b ? (const float*)b->Y : (const float*)nullptr, b ? (const float*)b->mean : nullptr,
                    b ? (const float*)b->rstd : nullptr, b ? b->gamma : nullptr, b ? b->beta : nullptr, xres, relu ? 1 : 0, a.M, a.cout, out, outf));
  MSQ_LAUNCH_CHECK();
  return MSQ_OK;
}

}  // namespace

// images [n_img, 3, S, S] fp32 -> ts->y_post [R * Lv, 2 E] (operand type): the input of visn_fc
template <typename T>
int rn_forward_train(msq_model* m, const float* images, int64_t n_img, const int32_t* img_index, int64_t R, cudaStream_t st) {
  TrainState* ts = m->train;
  RnTrain* r = ts->rn;
  const msq_config& c = m->cfg;
  const int S = c.vit_res, w = c.rn_width, Cf = 32 * w, E = c.rn_embed, g = S / 32, g2 = g * g, Lv = 1 + 2 * g2;
  const int64_t Mv = R * Lv;
  MSQ_REQUIRE(S % 32 == 0 && n_img > 0, "ResNet tower: resolution %d / no images", S);
  // ---- plan the tape (pre-BatchNorm outputs, activations) and the transient im2col buffer
  const int H1 = S / 2, H2 = S / 4;
  size_t max_col = 0;
  for (int pass = 0; pass < 2; ++pass) {
    Planner p{&r->tape, pass == 0};
    if (pass == 1) r->tape.reset();
    r->wimg = p.take<float>((size_t)n_img);
    r->partial = p.take<float>((size_t)RED_BLOCKS * 2 * max(Cf, 64));
    size_t ci = 0;
    auto convY = [&](int64_t M) { RnConvT& cv = r->convs[ci++]; cv.Y = p.take<float>((size_t)M * cv.cout); max_col = max(max_col, (size_t)M * cv.Kp); };
    const int64_t M1 = n_img * H1 * H1;
    convY(M1); convY(M1); convY(M1);
    r->a[0] = p.take<T>((size_t)M1 * (w / 2)); r->a[1] = p.take<T>((size_t)M1 * (w / 2)); r->a[2] = p.take<T>((size_t)M1 * w);
    r->x0 = p.take<T>((size_t)n_img * H2 * H2 * w);
    int H = H2, cin = w;
    r->blk.assign(m->rn_blocks.size(), RnBlockTape{});
    for (size_t bi = 0; bi < m->rn_blocks.size(); ++bi) {
      const RnBlockW& B = m->rn_blocks[bi];
      RnBlockTape& t = r->blk[bi];
      const int pl = B.planes, Ho = H / B.stride;
      const int64_t M = n_img * H * H, Mo = n_img * Ho * Ho;
      convY(M); convY(M); convY(Mo);
      if (B.has_ds) convY(Mo);
      t.o1 = p.take<T>((size_t)M * pl); t.o2 = p.take<T>((size_t)M * pl);
      t.o2p = B.stride > 1 ? p.take<T>((size_t)Mo * pl) : nullptr;
      t.xs = (B.has_ds && B.stride > 1) ? p.take<T>((size_t)Mo * cin) : nullptr;
      t.out = p.take<T>((size_t)Mo * 4 * pl);
      t.H = H; t.Ho = Ho;
      H = Ho; cin = 4 * pl;
    }
    r->feat = p.take<float>((size_t)n_img * g2 * Cf);
    r->tok = p.take<T>((size_t)Mv * Cf); r->qkv = p.take<T>((size_t)Mv * 3 * Cf); r->ctx = p.take<T>((size_t)Mv * Cf);
    r->o = p.take<float>((size_t)Mv * E);
    if (pass == 0) MSQ_TRY(r->tape.reserve(p.need + 4096, st));
  }
  MSQ_TRY(r->ws.reserve(max_col * sizeof(T) + 4096, st));
  T* A = reinterpret_cast<T*>(r->ws.base);
  r->n_img = n_img; r->R = R;

  MSQ_CUDA(launch_k(rn_img_weight_kernel, dim3(ceil_div(n_img, 128)), dim3(128), 0, st, img_index, R * 2, n_img, r->wimg));
  MSQ_LAUNCH_CHECK();
  const float slots = (float)(R * 2);
  size_t ci = 0;
  // ---- stem: three 3x3 convolutions (the first with stride 2), AvgPool2d(2)
  {
    const int64_t M1 = n_img * H1 * H1;
    const float Wt = slots * H1 * H1;
    RnConvT &c0 = r->convs[ci], &c1 = r->convs[ci + 1], &c2 = r->convs[ci + 2];
    ci += 3;
    MSQ_TRY(rn_im2col_stem<T>(images, n_img, S, c0.Kp, A, st));
    MSQ_TRY(conv_stats<T>(m, r, c0, A, M1, H1 * H1, Wt, st));
    MSQ_TRY(bn_apply<T>(c0, nullptr, nullptr, true, (T*)r->a[0], nullptr, st));
    MSQ_TRY(rn_im2col3<T>((const T*)r->a[0], n_img, H1, H1, w / 2, c1.Kp, A, st));
    MSQ_TRY(conv_stats<T>(m, r, c1, A, M1, H1 * H1, Wt, st));
    MSQ_TRY(bn_apply<T>(c1, nullptr, nullptr, true, (T*)r->a[1], nullptr, st));
    MSQ_TRY(rn_im2col3<T>((const T*)r->a[1], n_img, H1, H1, w / 2, c2.Kp, A, st));
    MSQ_TRY(conv_stats<T>(m, r, c2, A, M1, H1 * H1, Wt, st));
    MSQ_TRY(bn_apply<T>(c2, nullptr, nullptr, true, (T*)r->a[2], nullptr, st));
    MSQ_TRY(rn_avgpool2<T>((const T*)r->a[2], n_img, H1, H1, w, (T*)r->x0, st));
  }
  // ---- bottleneck blocks
  const T* x = (const T*)r->x0;
  int cin = w;
  for (size_t bi = 0; bi < m->rn_blocks.size(); ++bi) {
    const RnBlockW& B = m->rn_blocks[bi];
    RnBlockTape& t = r->blk[bi];
    const int pl = B.planes, H = t.H, Ho = t.Ho;
    const int64_t M = n_img * H * H, Mo = n_img * Ho * Ho;
    const float Wt = slots * H * H, Wto = slots * Ho * Ho;
    RnConvT &c1 = r->convs[ci], &c2 = r->convs[ci + 1], &c3 = r->convs[ci + 2];
    RnConvT* cd = B.has_ds ? &r->convs[ci + 3] : nullptr;
    ci += B.has_ds ? 4 : 3;
    t.x = x;
    MSQ_TRY(conv_stats<T>(m, r, c1, x, M, H * H, Wt, st));
    MSQ_TRY(bn_apply<T>(c1, nullptr, nullptr, true, (T*)t.o1, nullptr, st));
    MSQ_TRY(rn_im2col3<T>((const T*)t.o1, n_img, H, H, pl, c2.Kp, A, st));
    MSQ_TRY(conv_stats<T>(m, r, c2, A, M, H * H, Wt, st));
    MSQ_TRY(bn_apply<T>(c2, nullptr, nullptr, true, (T*)t.o2, nullptr, st));
    const T* o2 = (const T*)t.o2;
    const T* xs = x;
    if (B.stride > 1) {
      MSQ_TRY(rn_avgpool2<T>((const T*)t.o2, n_img, H, H, pl, (T*)t.o2p, st));
      o2 = (const T*)t.o2p;
      if (B.has_ds) { MSQ_TRY(rn_avgpool2<T>(x, n_img, H, H, cin, (T*)t.xs, st)); xs = (const T*)t.xs; }
    }
    MSQ_TRY(conv_stats<T>(m, r, c3, o2, Mo, Ho * Ho, Wto, st));
    if (cd) MSQ_TRY(conv_stats<T>(m, r, *cd, xs, Mo, Ho * Ho, Wto, st));
    const bool last = bi + 1 == m->rn_blocks.size();
    MSQ_TRY(bn_apply<T>(c3, cd, cd ? nullptr : x, true, (T*)t.out, last ? r->feat : nullptr, st));
    x = (const T*)t.out;
    cin = 4 * pl;
  }
  // ---- attention pool over the pair's tokens, c_proj, cat(o, o) + position / token-type embeddings
  MSQ_TRY(rn_tokens<T>(r->feat, img_index, R, g2, Cf, m->rn_pos, (T*)r->tok, st));
  MSQ_TRY((gemm_nt<T, T>(m, (const T*)r->tok, Cf, wptr<T>(m->rn_qkv), m->rn_qkv.ld, m->rn_qkv.b, nullptr, 0, (T*)r->qkv, 3 * Cf, Mv, 3 * Cf, Cf,
                         ACT_NONE, st)));
  MSQ_TRY(attention<T>((const T*)r->qkv, R, Lv, Cf / 64, 64, 0.125f, nullptr, 0, 0, (T*)r->ctx, st));
  MSQ_TRY((gemm_nt<T, float>(m, (const T*)r->ctx, Cf, wptr<T>(m->rn_cproj), m->rn_cproj.ld, m->rn_cproj.b, nullptr, 0, r->o, E, Mv, E, Cf, ACT_NONE,
                             st)));
  MSQ_TRY(rn_finish<T>(r->o, Mv, Lv, E, m->rn_posadd, (T*)ts->y_post, st));
  return MSQ_OK;
}
template int rn_forward_train<float>(msq_model*, const float*, int64_t, const int32_t*, int64_t, cudaStream_t);
template int rn_forward_train<bf16>(msq_model*, const float*, int64_t, const int32_t*, int64_t, cudaStream_t);

// =====================================================================================================
// backward
// =====================================================================================================
namespace {

template <typename T> struct RnBwd {
  msq_model* m; RnTrain* r; TrainState* ts; float* grads; BwdBufs* b; cudaStream_t st;
  float *gP, *gQ, *gS;   // fp32 gradients of activations
  T *gY, *A;             // dY operand; im2col / dcol buffer
  float* wscr;           // [Cout, Kp] weight-gradient scratch of the 3x3 convolutions
  int err = MSQ_OK;
  float* G(const std::string& name) {
    auto it = ts->index.find(name);
    if (it == ts->index.end()) { set_error("train: no gradient slot for %s", name.c_str()); err = MSQ_ERR_STATE; return nullptr; }
    return grads + ts->slots[it->second].off;
  }
  // BatchNorm (+ ReLU mask `act`) backward of one convolution output: dY into gY, dgamma / dbeta accumulated
  int bn_bwd(RnConvT& cv, const float* dout, const T* act) {
    float *dg = G(cv.bn + ".weight"), *db = G(cv.bn + ".bias");
    if (err) return err;
    MSQ_TRY((colred<T, 2>((const float*)cv.Y, dout, act, cv.mean, cv.rstd, nullptr, cv.HW, cv.M, cv.cout, 1.f, r->partial, cv.S1, cv.S2, dg, db, st)));
    MSQ_CUDA(launch_k(rn_bn_bwd_apply_kernel<T>, ew_grid(cv.M * cv.cout), dim3(256), 0, st, dout, act, (const float*)cv.Y, (const float*)cv.mean,
                      (const float*)cv.rstd, cv.gamma, (const float*)cv.S1, (const float*)cv.S2, (const float*)r->wimg, cv.HW,
                      r->bn_eval ? 0.f : 1.0f / cv.Wtot, cv.M, cv.cout, gY));
    MSQ_LAUNCH_CHECK();
    return MSQ_OK;
  }
  // weight gradient of a convolution from gY and its GEMM operand X [M, Kp]
  int conv_wgrad(RnConvT& cv, const T* X) {
    float* dw = G(cv.wname);
    if (err) return err;
    if (cv.k == 1) return wgrad<T>(m, gY, cv.cout, cv.cout, X, cv.Kp, cv.Kp, ACT_NONE, cv.M, dw, nullptr, *b, st);
    MSQ_CUDA(cudaMemsetAsync(wscr, 0, (size_t)cv.cout * cv.Kp * sizeof(float), st));
    MSQ_TRY(wgrad<T>(m, gY, cv.cout, cv.cout, X, cv.Kp, cv.Kp, ACT_NONE, cv.M, wscr, nullptr, *b, st));
    MSQ_CUDA(launch_k(rn_wgrad_permute_kernel, ew_grid((int64_t)cv.cout * cv.K), dim3(256), 0, st, (const float*)wscr, cv.cout, cv.cin, cv.k, cv.Kp, dw));
    MSQ_LAUNCH_CHECK();
    return MSQ_OK;
  }
  // input gradient of a 1x1 convolution (fp32 rows) / of a 3x3 convolution (dcol into A, then col2im)
  int conv_dgrad1(RnConvT& cv, float* dx) { return dgrad_any<T, float>(m, gY, cv.cout, cv.wgT, cv.Kp, dx, cv.M, st); }
  int conv_dgrad3(RnConvT& cv, int64_t n, int H, float* dx) {
    MSQ_TRY((dgrad_any<T, T>(m, gY, cv.cout, cv.wgT, cv.Kp, A, cv.M, st)));
    MSQ_CUDA(launch_k(rn_col2im3_kernel<T>, ew_grid(cv.M * cv.cin), dim3(256), 0, st, (const T*)A, n, H, H, cv.cin, cv.Kp, dx));
    MSQ_LAUNCH_CHECK();
    return MSQ_OK;
  }
  int pool_bwd(const float* dout, int64_t n, int H, int C, bool accum, float* dx) {
    MSQ_CUDA(launch_k(rn_avgpool2_bwd_kernel, ew_grid(n * H * H * (int64_t)C), dim3(256), 0, st, dout, n, H, H, C, accum ? 1 : 0, dx));
    MSQ_LAUNCH_CHECK();
    return MSQ_OK;
  }
};

}  // namespace

// dyp [R * Lv, 2 E] fp32 = gradient of the tower output (the input of visn_fc); accumulates every tower gradient into `grads`
template <typename T>
int rn_backward_train(msq_model* m, const float* dyp, float* grads, cudaStream_t st) {
  TrainState* ts = m->train;
  RnTrain* r = ts->rn;
  const msq_config& c = m->cfg;
  const int S = c.vit_res, w = c.rn_width, Cf = 32 * w, E = c.rn_embed, F = 2 * E, g = S / 32, g2 = g * g, Lv = 1 + 2 * g2, heads = Cf / 64;
  const int64_t R = r->R, n_img = r->n_img, Mv = R * Lv;
  const std::string P = m->prefix_inner, v = P + "encoder.visual_model.visual.", a = v + "attnpool.";
  // ---- scratch: sized by the largest layer
  size_t max_act = (size_t)max(Mv * (int64_t)max(3 * Cf, F), n_img * (int64_t)g2 * Cf), max_col = 0, max_w = 0, max_part = 0, max_gt = 0, max_xt = 0;
  int max_c = max(3 * Cf, F);
  for (const RnConvT& cv : r->convs) {
    max_act = max(max_act, (size_t)cv.M * max(cv.cout, cv.cin));
    max_col = max(max_col, (size_t)cv.M * cv.Kp);
    max_w = max(max_w, (size_t)cv.cout * cv.Kp);
    max_c = max(max_c, max(cv.cout, cv.Kp));
    const int64_t Mp = wgrad_rows(cv.M);
    max_gt = max(max_gt, (size_t)cv.cout * Mp); max_xt = max(max_xt, (size_t)cv.Kp * Mp);
  }
  {   // attention-pool weight gradients
    const int64_t Mp = wgrad_rows(Mv);
    max_w = max(max_w, (size_t)3 * Cf * Cf);
    max_gt = max(max_gt, (size_t)3 * Cf * Mp); max_xt = max(max_xt, (size_t)Cf * Mp);
  }
  max_part = (size_t)SPLITK_MAX * max_w;
  const bool tn = sizeof(T) == 2 && wgrad_tn_enabled() && model_use_tc(m);   // TN operands: no transposed copies needed
  BwdBufs b{};
  RnBwd<T> k{m, r, ts, grads, &b, st};
  float *dpa = nullptr, *dtok = nullptr;
  for (int pass = 0; pass < 2; ++pass) {
    Planner p{&r->ws, pass == 0};
    if (pass == 1) r->ws.reset();
    k.gP = p.take<float>(max_act); k.gQ = p.take<float>(max_act); k.gS = p.take<float>(max_act);
    k.gY = p.take<T>(max_act);
    k.A = p.take<T>(max(max_col, (size_t)Mv * 3 * Cf));
    k.wscr = p.take<float>(max_w);
    b.GT = p.take<T>(tn ? 64 : max_gt); b.XT = p.take<T>(tn ? 64 : max_xt);
    b.scr_floats = max(colsum_scratch_floats(max_c), (size_t)1024);
    b.ln_scr = p.take<float>(b.scr_floats);
    b.at_scr = p.take<float>(attention_bwd_scratch_floats(R, Lv, heads));
    if (sizeof(T) == 2) b.part = p.take<float>(max_part);
    dpa = p.take<float>((size_t)Lv * F);
    dtok = p.take<float>((size_t)Mv * Cf);
    if (pass == 0) MSQ_TRY(r->ws.reserve(p.need + 4096, st));
  }
  if (sizeof(T) == 2) b.sk = &ts->splitk;

  // ---- tower output: cat(o, o) + posadd
  {
    float *dxe = k.G(P + "encoder.visual_pos.x_position_embedding.weight"), *dye = k.G(P + "encoder.visual_pos.y_position_embedding.weight"),
          *dte = k.G(P + "encoder.visual_token_type.token_type_embedding.weight");
    if (k.err) return k.err;
    MSQ_CUDA(launch_k(rn_posadd_sum_kernel, ew_grid((int64_t)Lv * F), dim3(256), 0, st, dyp, R, Lv, F, dpa));
    MSQ_LAUNCH_CHECK();
    MSQ_CUDA(launch_k(rn_posadd_bwd_kernel, ew_grid((int64_t)(2 * g + 2) * F), dim3(256), 0, st, (const float*)dpa, g, F, dxe, dye, dte));
    MSQ_LAUNCH_CHECK();
    MSQ_CUDA(launch_k(rn_finish_bwd_kernel<T>, ew_grid(Mv * E), dim3(256), 0, st, dyp, Mv, E, k.gY));
    MSQ_LAUNCH_CHECK();
  }
  // ---- c_proj, attention, fused q / k / v projection
  {
    float *dWc = k.G(a + "c_proj.weight"), *dbc = k.G(a + "c_proj.bias"), *dWq = k.G(a + "q_proj.weight"), *dbq = k.G(a + "q_proj.bias"),
          *dpos = k.G(a + "positional_embedding");
    if (k.err) return k.err;
    MSQ_TRY(wgrad<T>(m, k.gY, E, E, (const T*)r->ctx, Cf, Cf, ACT_NONE, Mv, dWc, dbc, b, st));
    T* dctx = reinterpret_cast<T*>(k.gS);   // [Mv, Cf] (gS is free until the first block)
    MSQ_TRY((dgrad<T, T>(m, k.gY, E, r->cprojT, Cf, nullptr, dctx, Mv, st)));
    MSQ_TRY(attn_bwd<T>((const T*)r->qkv, (const T*)r->ctx, dctx, R, Lv, heads, nullptr, 0, k.A, b.at_scr, st));   // dqkv [Mv, 3 Cf] in A
    MSQ_TRY(wgrad<T>(m, k.A, 3 * Cf, 3 * Cf, (const T*)r->tok, Cf, Cf, ACT_NONE, Mv, dWq, dbq, b, st));
    MSQ_TRY((dgrad<T, float>(m, k.A, 3 * Cf, r->qkvT, Cf, nullptr, dtok, Mv, st)));
    MSQ_CUDA(launch_k(rn_pos_bwd_kernel, ew_grid((int64_t)(g2 + 1) * Cf), dim3(256), 0, st, (const float*)dtok, R, g2, Cf, dpos));
    MSQ_LAUNCH_CHECK();
    MSQ_CUDA(launch_k(rn_tokens_bwd_kernel, ew_grid(n_img * g2 * (int64_t)Cf), dim3(256), 0, st, (const float*)dtok, (const int32_t*)ts->img_index, R,
                      n_img, g2, Cf, k.gP));
    MSQ_LAUNCH_CHECK();
  }
  // ---- bottleneck blocks, last to first.  gP = gradient of the block output; gS collects the gradient of its input.
  size_t ci = r->convs.size();
  for (size_t bi = m->rn_blocks.size(); bi-- > 0;) {
    const RnBlockW& B = m->rn_blocks[bi];
    RnBlockTape& t = r->blk[bi];
    const int pl = B.planes, H = t.H, Ho = t.Ho, cin = B.cin;
    const int64_t M = n_img * H * H, Mo = n_img * Ho * Ho;
    ci -= B.has_ds ? 4 : 3;
    RnConvT &c1 = r->convs[ci], &c2 = r->convs[ci + 1], &c3 = r->convs[ci + 2];
    RnConvT* cd = B.has_ds ? &r->convs[ci + 3] : nullptr;
    const T* out = (const T*)t.out;
    const T* o2in = B.stride > 1 ? (const T*)t.o2p : (const T*)t.o2;
    // main branch: conv3
    MSQ_TRY(k.bn_bwd(c3, k.gP, out));
    MSQ_TRY(k.conv_wgrad(c3, o2in));
    MSQ_TRY(k.conv_dgrad1(c3, k.gQ));                                     // d o2p [Mo, pl]
    const float* d_o2 = k.gQ;
    if (B.stride > 1) { MSQ_TRY(k.pool_bwd(k.gQ, n_img, H, pl, false, k.gS)); d_o2 = k.gS; }   // [M, pl]
    MSQ_TRY(k.bn_bwd(c2, d_o2, (const T*)t.o2));
    MSQ_TRY(rn_im2col3<T>((const T*)t.o1, n_img, H, H, pl, c2.Kp, k.A, st));
    MSQ_TRY(k.conv_wgrad(c2, k.A));
    MSQ_TRY(k.conv_dgrad3(c2, n_img, H, k.gQ));                            // d o1 [M, pl]
    MSQ_TRY(k.bn_bwd(c1, k.gQ, (const T*)t.o1));
    MSQ_TRY(k.conv_wgrad(c1, (const T*)t.x));
    MSQ_TRY(k.conv_dgrad1(c1, k.gS));                                     // dx (main branch) [M, cin]
    // shortcut
    if (cd) {
      MSQ_TRY(k.bn_bwd(*cd, k.gP, out));
      MSQ_TRY(k.conv_wgrad(*cd, B.stride > 1 ? (const T*)t.xs : (const T*)t.x));
      MSQ_TRY(k.conv_dgrad1(*cd, k.gQ));                                  // d xs [Mo, cin]
      if (B.stride > 1) MSQ_TRY(k.pool_bwd(k.gQ, n_img, H, cin, true, k.gS));
      else { MSQ_CUDA(launch_k(rn_add_kernel, ew_grid(M * cin), dim3(256), 0, st, (const float*)k.gQ, M * (int64_t)cin, k.gS)); MSQ_LAUNCH_CHECK(); }
    } else {
      MSQ_CUDA(launch_k(rn_add_masked_kernel<T>, ew_grid(Mo * 4 * pl), dim3(256), 0, st, (const float*)k.gP, out, Mo * (int64_t)4 * pl, 1, k.gS));
      MSQ_LAUNCH_CHECK();
    }
    std::swap(k.gP, k.gS);
    MSQ_TRY(mark_ready(ts, c1.wname, cd ? cd->bn + ".bias" : c3.bn + ".bias", st));
  }
  // ---- stem: AvgPool2d(2), three 3x3 convolutions
  {
    const int H1 = S / 2;
    RnConvT &c0 = r->convs[0], &c1 = r->convs[1], &c2 = r->convs[2];
    MSQ_TRY(k.pool_bwd(k.gP, n_img, H1, w, false, k.gQ));                 // d a2 [M1, w]
    MSQ_TRY(k.bn_bwd(c2, k.gQ, (const T*)r->a[2]));
    MSQ_TRY(rn_im2col3<T>((const T*)r->a[1], n_img, H1, H1, w / 2, c2.Kp, k.A, st));
    MSQ_TRY(k.conv_wgrad(c2, k.A));
    MSQ_TRY(k.conv_dgrad3(c2, n_img, H1, k.gQ));                           // d a1
    MSQ_TRY(k.bn_bwd(c1, k.gQ, (const T*)r->a[1]));
    MSQ_TRY(rn_im2col3<T>((const T*)r->a[0], n_img, H1, H1, w / 2, c1.Kp, k.A, st));
    MSQ_TRY(k.conv_wgrad(c1, k.A));
    MSQ_TRY(k.conv_dgrad3(c1, n_img, H1, k.gQ));                           // d a0
    MSQ_TRY(k.bn_bwd(c0, k.gQ, (const T*)r->a[0]));
    MSQ_TRY(rn_im2col_stem<T>(ts->images, n_img, S, c0.Kp, k.A, st));
    MSQ_TRY(k.conv_wgrad(c0, k.A));
    MSQ_TRY(mark_ready(ts, c0.wname, c2.bn + ".bias", st));
    MSQ_TRY(mark_ready(ts, a + "positional_embedding", P + "encoder.visual_token_type.token_type_embedding.weight", st));
  }
  return k.err;
}
template int rn_backward_train<float>(msq_model*, const float*, float*, cudaStream_t);
template int rn_backward_train<bf16>(msq_model*, const float*, float*, cudaStream_t);

}  // namespace msq
