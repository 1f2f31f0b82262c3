// HBM-bound passes of the IIC trunk (NHWC activations, 8-channel / 16-32 B vector accesses):
// layout plumbing, sobel, BatchNorm statistics / apply / backward, ReLU, residual add, max / avg
// pooling, weight repacking.  Each kernel cites the reference op it replaces.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

namespace iic {

static inline int ew_grid(long long work_items, int threads) {
  long long blocks = (work_items + threads - 1) / threads;
  long long cap = (long long)device_sm_count() * 16;  // multiple of the SM count, enough waves
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------------------------
// layout / dtype plumbing
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, T* __restrict__ dst, int n, int c, int h, int w) {
  const long long total = (long long)n * c * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int ci = (int)(i % c);
    long long p = i / c;  // n*h*w index
    int hw = (int)(p % ((long long)h * w));
    int ni = (int)(p / ((long long)h * w));
    dst[i] = from_f<T>(src[((long long)ni * c + ci) * h * w + hw]);
  }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ src, float* __restrict__ dst, int n, int c, int h, int w) {
  const long long total = (long long)n * c * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int hw = (int)(i % ((long long)h * w));
    long long q = i / ((long long)h * w);
    int ci = (int)(q % c);
    int ni = (int)(q / c);
    dst[i] = to_f(src[((long long)ni * h * w + hw) * c + ci]);
  }
}
template <typename S, typename D>
__global__ void cast_kernel(const S* __restrict__ src, D* __restrict__ dst, long long count) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    dst[i] = from_f<D>(to_f(src[i]));
}

// ---------------------------------------------------------------------------------------------
// sobel_process -- code/utils/cluster/transforms.py:47-96
// ---------------------------------------------------------------------------------------------
__global__ void sobel_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int cin, int cout, int h,
                             int w, int grey_ch, int n_rgb, int ir_ch /* -1: none */) {
  const long long total = (long long)n * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int x = (int)(i % w);
    int y = (int)((i / w) % h);
    int ni = (int)(i / ((long long)w * h));
    const float* g = in + ((long long)ni * cin + grey_ch) * h * w;
    float v[3][3];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        int yy = y + dy, xx = x + dx;
        v[dy + 1][dx + 1] = (yy >= 0 && yy < h && xx >= 0 && xx < w) ? g[(long long)yy * w + xx] : 0.f;
      }
    // cross-correlation with [[1,0,-1],[2,0,-2],[1,0,-1]] (:69) and [[1,2,1],[0,0,0],[-1,-2,-1]] (:75);
    // summed in the same row-major tap order a 3x3 conv uses
    float sx = v[0][0] - v[0][2] + 2.f * v[1][0] - 2.f * v[1][2] + v[2][0] - v[2][2];
    float sy = v[0][0] + 2.f * v[0][1] + v[0][2] - v[2][0] - 2.f * v[2][1] - v[2][2];
    float* o = out + (long long)ni * cout * h * w + (long long)y * w + x;
    const float* src = in + (long long)ni * cin * h * w + (long long)y * w + x;
    for (int c = 0; c < n_rgb; ++c) o[(long long)c * h * w] = src[(long long)c * h * w];
    o[(long long)n_rgb * h * w] = sx;
    o[(long long)(n_rgb + 1) * h * w] = sy;
    if (ir_ch >= 0) o[(long long)(n_rgb + 2) * h * w] = src[(long long)ir_ch * h * w];
  }
}

// ---------------------------------------------------------------------------------------------
// BatchNorm2d (train mode) -- net5g.py:24, residual.py:20,23,56, vgg.py:28
// ---------------------------------------------------------------------------------------------
// Per-channel sum / sum of squares over M rows of y[M][C].  Threads are laid out so that a warp
// reads whole 128 B+ row segments; each thread owns 8 channels and walks rows with a grid stride,
// two independent rows in flight.  Each block writes ONE partial row (no atomics: 1184 blocks
// hammering 128 addresses cost more than the streaming pass itself); bn_sum_partials_kernel folds them.
template <typename T, bool BWD>
__global__ void __launch_bounds__(256) bn_reduce_kernel(const T* __restrict__ y, const T* __restrict__ gin,
                                                        const T* __restrict__ act,
                                                        const float* __restrict__ mask_ss,
                                                        const float* __restrict__ mean_invstd, long long M, int C,
                                                        float* __restrict__ partial /* [gridDim.x][2C] */) {
  // BWD=false: sum y, sum y*y
  // BWD=true : g = gin * relu_mask ; yhat = (y-mean)*invstd ; sum g ; sum g*yhat
  //            relu_mask = (act > 0) if act, else (y*scale+shift > 0) if mask_ss (BN directly followed by ReLU:
  //            the mask is recomputed from y, which is read anyway, instead of reading the activation)
  __shared__ float sh[256 * 17];
  const int cg = C >> 3;                // channel groups per row (C <= 2048)
  const int tpr = cg;                   // threads per row
  const int rows_per_it = 256 / tpr;
  const int my_cg = threadIdx.x % tpr, my_r = threadIdx.x / tpr;
  float a0[8], a1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) a0[j] = a1[j] = 0.f;
  if (my_r < rows_per_it) {
    const int c8 = my_cg;
    float mean[8], istd[8], msc[8], msh[8];
    if (BWD) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        mean[j] = mean_invstd[c8 * 8 + j];
        istd[j] = mean_invstd[C + c8 * 8 + j];
        msc[j] = mask_ss ? mask_ss[c8 * 8 + j] : 0.f;
        msh[j] = mask_ss ? mask_ss[C + c8 * 8 + j] : 0.f;
      }
    }
    const long long stride = (long long)gridDim.x * rows_per_it;
    constexpr int U = 4;  // rows in flight per thread: all (raw, unconverted) loads are issued before any arithmetic
    using Raw = typename RawOf<T>::type;
    for (long long r = (long long)blockIdx.x * rows_per_it + my_r; r < M; r += U * stride) {
      Raw rv[U], rg[U], ra[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long ru = r + u * stride;
        ok[u] = ru < M;
        if (ok[u]) {
          load_raw(y + ru * C + c8 * 8, rv[u]);
          if (BWD) {
            load_raw(gin + ru * C + c8 * 8, rg[u]);
            if (act != nullptr) load_raw(act + ru * C + c8 * 8, ra[u]);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        float v[8], g[8], a[8];
        cvt_raw(rv[u], v);
        if (BWD) {
          cvt_raw(rg[u], g);
          if (act != nullptr) cvt_raw(ra[u], a);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (!BWD) {
            a0[j] += v[j];
            a1[j] = fmaf(v[j], v[j], a1[j]);
          } else {
            float gj = g[j];
            if (act != nullptr) gj = a[j] > 0.f ? gj : 0.f;
            else if (mask_ss != nullptr) gj = fmaf(v[j], msc[j], msh[j]) > 0.f ? gj : 0.f;
            a0[j] += gj;
            a1[j] = fmaf(gj, (v[j] - mean[j]) * istd[j], a1[j]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    sh[threadIdx.x * 17 + j] = a0[j];
    sh[threadIdx.x * 17 + 8 + j] = a1[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < tpr * 16; i += 256) {
    const int g = i / 16, j = i % 16;
    float t = 0.f;
    for (int r = 0; r < rows_per_it; ++r) t += sh[(r * tpr + g) * 17 + j];
    partial[(long long)blockIdx.x * 2 * C + (j >> 3) * C + g * 8 + (j & 7)] = t;
  }
}

// sums[i] = sum_b partial[b][i]  (double, fixed order => deterministic); one warp per entry
__global__ void bn_sum_partials_kernel(const float* __restrict__ partial, int nblk, int n2c, long long stride,
                                       double* __restrict__ sums) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n2c) return;
  double t = 0.0;
  for (int b = lane; b < nblk; b += 32) t += (double)partial[(long long)b * stride + warp];
  t = warp_sum(t);
  if (lane == 0) sums[warp] = t;
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, long long M, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* running_mean,
                                   float* running_var, int use_running, float* __restrict__ scale_shift,
                                   float* __restrict__ mean_invstd) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mean, invstd;
  if (use_running) {
    mean = running_mean[c];
    invstd = 1.f / sqrtf(running_var[c] + eps);
  } else {
    const double m = sums[c] / (double)M;
    double var = sums[C + c] / (double)M - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    invstd = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean != nullptr) {
      const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  }
  const float sc = gamma[c] * invstd;
  scale_shift[c] = sc;
  scale_shift[C + c] = beta[c] - mean * sc;
  mean_invstd[c] = mean;
  mean_invstd[C + c] = invstd;
}

// conv-epilogue partials [nblk][views][2C] -> per-view scale/shift and mean/invstd in ONE launch.  Block = 32 channels x
// 8 row groups: thread (cx, by) folds the CTA rows by, by + 8, ... of its channel in fp64 (coalesced across cx, all of a
// thread's loads independent), the 8 group sums are added in a fixed order, and row group 0 finalises; running statistics
// are updated view after view, like the reference's consecutive forward calls.  (The first version, one warp per channel
// with lanes striding over the CTA rows, issued 32-line gathers in a dependent loop: ~16 us per launch, 36 launches a step.)
__global__ void __launch_bounds__(256)
bn_fold_finalize_kernel(const float* __restrict__ partial, int nblk, int slots, int views, long long M, int C,
                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                        float momentum, float* running_mean, float* running_var,
                        float* __restrict__ scale_shift, float* __restrict__ mean_invstd) {
  __shared__ double red[2][8][32];
  const int cx = threadIdx.x & 31, by = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const bool live = c < C;
  const long long stride = (long long)slots * 2 * C;  // a CTA's partial row holds `slots` views
  for (int v = 0; v < views; ++v) {
    double s1 = 0.0, s2 = 0.0;
    if (live) {
      const float* p = partial + (long long)v * 2 * C + c;
#pragma unroll 4
      for (int b = by; b < nblk; b += 8) {
        s1 += (double)p[b * stride];
        s2 += (double)p[b * stride + C];
      }
    }
    red[0][by][cx] = s1;
    red[1][by][cx] = s2;
    __syncthreads();
    if (by == 0 && live) {
      s1 = red[0][0][cx];
      s2 = red[1][0][cx];
#pragma unroll
      for (int g = 1; g < 8; ++g) {
        s1 += red[0][g][cx];
        s2 += red[1][g][cx];
      }
      const double m = s1 / (double)M;
      double var = s2 / (double)M - m * m;
      if (var < 0.0) var = 0.0;
      const float mean = (float)m, invstd = (float)(1.0 / sqrt(var + (double)eps));
      if (running_mean != nullptr) {
        const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
      }
      const float sc = gamma[c] * invstd;
      float* ss = scale_shift + (long long)v * 2 * C;
      float* mi = mean_invstd + (long long)v * 2 * C;
      ss[c] = sc;
      ss[C + c] = beta[c] - mean * sc;
      mi[c] = mean;
      mi[C + c] = invstd;
    }
    __syncthreads();
  }
}

// out = relu?( y*scale+shift [+ res | + res*rscale + rshift] )   (residual.py:27-43)
// MASKOUT: additionally writes the ReLU mask of the result as one byte per 8 channels (bit j = out[8*c8 + j] > 0),
// [views * M][C / 8] bytes; the BatchNorm backward of a residual block then reads 1 bit instead of the 16-bit block
// output to decide where the gradient passes (iic_bn_bwd_fused_bits).
template <typename T, bool MASKOUT = false>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ y, const float* __restrict__ ss,
                                                       const T* __restrict__ res, const float* __restrict__ rss,
                                                       T* __restrict__ out, long long M, int C, int relu, int views,
                                                       unsigned char* __restrict__ mask_out = nullptr) {
  // views > 1: the tensor is `views` stacked batches of M rows with their own coefficients ([views][2C]);
  // the grid is split evenly between them
  const int cg = C >> 3;
  const long long total = M * cg;
  const int Gv = gridDim.x / views, v = blockIdx.x / Gv, lb = blockIdx.x % Gv;
  y += (long long)v * M * C;
  out += (long long)v * M * C;
  ss += (long long)v * 2 * C;
  if (res != nullptr) res += (long long)v * M * C;
  if (rss != nullptr) rss += (long long)v * 2 * C;
  if (MASKOUT) mask_out += (long long)v * total;
  // a thread's channel group is loop invariant (blockDim.x is a multiple of C/8): keep the coefficients in registers
  const int c8 = (int)((lb * (long long)blockDim.x + threadIdx.x) % cg);
  float sc[8], sh[8], rsc[8], rsh[8];
  load8(ss + c8 * 8, sc);
  load8(ss + C + c8 * 8, sh);
  if (rss != nullptr) {
    load8(rss + c8 * 8, rsc);
    load8(rss + C + c8 * 8, rsh);
  }
  for (long long i = lb * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)Gv * blockDim.x) {
    float v[8];
    load8(y + i * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
    if (res != nullptr) {
      float r[8];
      load8(res + i * 8, r);
      if (rss != nullptr) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = fmaf(r[j], rsc[j], rsh[j]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += r[j];
    }
    if (MASKOUT) {
      unsigned int bits = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) bits |= (v[j] > 0.f ? 1u : 0u) << j;
      mask_out[i] = (unsigned char)bits;
    }
    if (relu) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    store8(out + i * 8, v);
  }
}

// BN + ReLU + MaxPool2d(2, 2, pad) in one pass (net5g.py:24-26).  One thread per (window, 8 ch).
template <typename T>
__global__ void __launch_bounds__(256) bn_relu_maxpool_kernel(const T* __restrict__ y, const float* __restrict__ ss,
                                                              T* __restrict__ out, int n, int h, int w, int C, int pad,
                                                              int oh, int ow) {
  const int cg = C >> 3;
  const long long total = (long long)n * oh * ow * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c8, ox, oy, ni;
    if (total <= 0xffffffffll) {  // (the usual case: 32-bit divisions; the 64-bit ones cost more than the memory traffic)
      unsigned u = (unsigned)i;
      c8 = (int)(u % (unsigned)cg); u /= (unsigned)cg;
      ox = (int)(u % (unsigned)ow); u /= (unsigned)ow;
      oy = (int)(u % (unsigned)oh);
      ni = (int)(u / (unsigned)oh);
    } else {
      c8 = (int)(i % cg);
      long long p = i / cg;
      ox = (int)(p % ow);
      p /= ow;
      oy = (int)(p % oh);
      ni = (int)(p / oh);
    }
    float sc[8], sh[8], m[8];
    load8(ss + c8 * 8, sc);
    load8(ss + C + c8 * 8, sh);
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int iy = oy * 2 - pad + dy, ix = ox * 2 - pad + dx;
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
          float v[8];
          load8(y + (((long long)ni * h + iy) * w + ix) * C + c8 * 8, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f));
        }
      }
    store8(out + i * 8, m);
  }
}

// Backward of the fused BN+ReLU+MaxPool: routes dP to the first arg-max of each window (torch's
// max_pool2d tie rule: strict '>' in row-major window order), masked by the ReLU.
template <typename T>
__global__ void __launch_bounds__(256) bn_relu_maxpool_bwd_kernel(const T* __restrict__ y, const float* __restrict__ ss,
                                                                  const T* __restrict__ dpool, T* __restrict__ g, int n,
                                                                  int h, int w, int C, int pad, int oh, int ow) {
  const int cg = C >> 3;
  const long long total = (long long)n * oh * ow * cg;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int c8, ox, oy, ni;
    if (total <= 0xffffffffll) {  // (the usual case: 32-bit divisions; the 64-bit ones cost more than the memory traffic)
      unsigned u = (unsigned)i;
      c8 = (int)(u % (unsigned)cg); u /= (unsigned)cg;
      ox = (int)(u % (unsigned)ow); u /= (unsigned)ow;
      oy = (int)(u % (unsigned)oh);
      ni = (int)(u / (unsigned)oh);
    } else {
      c8 = (int)(i % cg);
      long long p = i / cg;
      ox = (int)(p % ow);
      p /= ow;
      oy = (int)(p % oh);
      ni = (int)(p / oh);
    }
    float sc[8], sh[8], m[8], dp[8];
    int arg[8];
    load8(ss + c8 * 8, sc);
    load8(ss + C + c8 * 8, sh);
    load8(dpool + i * 8, dp);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      m[j] = -INFINITY;
      arg[j] = -1;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int iy = oy * 2 - pad + (q >> 1), ix = ox * 2 - pad + (q & 1);
      if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
        float v[8];
        load8(y + (((long long)ni * h + iy) * w + ix) * C + c8 * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float a = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f);
          if (a > m[j]) {
            m[j] = a;
            arg[j] = q;
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int iy = oy * 2 - pad + (q >> 1), ix = ox * 2 - pad + (q & 1);
      if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (arg[j] == q && m[j] > 0.f) ? dp[j] : 0.f;
        store8(g + (((long long)ni * h + iy) * w + ix) * C + c8 * 8, o);
      }
    }
  }
}

// dy = gamma*istd * (g - mean(g) - yhat * mean(g*yhat)) = A*g + B*y + Cc with per-channel coefficients
// (computed once per block into shared memory; no fp64 in the streaming loop).  Optional g_out =
// masked gradient for the residual branch.
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const T* __restrict__ gin, const T* __restrict__ act,
                                                           const T* __restrict__ y,
                                                           const float* __restrict__ mask_ss,
                                                           const float* __restrict__ mean_invstd,
                                                           const float* __restrict__ gamma,
                                                           const double* __restrict__ sums, T* __restrict__ dy,
                                                           T* __restrict__ g_out, float* dgamma, float* dbeta,
                                                           int accumulate, long long M, int C) {
  extern __shared__ __align__(16) float coef[];  // [5][C]: A, B, C, mask scale, mask shift
  const int cg = C >> 3;
  const long long total = M * cg;
  const double invM = 1.0 / (double)M;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float mean = mean_invstd[c], istd = mean_invstd[C + c];
    const float m1 = (float)(sums[c] * invM), m2 = (float)(sums[C + c] * invM);
    const float A = gamma[c] * istd;
    coef[c] = A;
    coef[C + c] = -A * m2 * istd;
    coef[2 * C + c] = -A * m1 + A * m2 * istd * mean;
    coef[3 * C + c] = mask_ss ? mask_ss[c] : 0.f;
    coef[4 * C + c] = mask_ss ? mask_ss[C + c] : 0.f;
    if (blockIdx.x == 0) {
      const float db = (float)sums[c], dg = (float)sums[C + c];
      if (dgamma) dgamma[c] = accumulate ? dgamma[c] + dg : dg;
      if (dbeta) dbeta[c] = accumulate ? dbeta[c] + db : db;
    }
  }
  __syncthreads();
  // blockDim.x (256) is a multiple of C/8 (<= 64 groups... up to 256), so a thread's channel group never changes
  const int c8 = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) % cg);
  float cA[8], cB[8], cC[8], ms[8], mh[8];
  load8(coef + c8 * 8, cA);
  load8(coef + C + c8 * 8, cB);
  load8(coef + 2 * C + c8 * 8, cC);
  load8(coef + 3 * C + c8 * 8, ms);
  load8(coef + 4 * C + c8 * 8, mh);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    float g[8], v[8];
    load8(gin + i * 8, g);
    load8(y + i * 8, v);
    if (act != nullptr) {
      float a[8];
      load8(act + i * 8, a);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = a[j] > 0.f ? g[j] : 0.f;
    } else if (mask_ss != nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = fmaf(v[j], ms[j], mh[j]) > 0.f ? g[j] : 0.f;
    }
    if (g_out != nullptr) store8(g_out + i * 8, g);
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf(cA[j], g[j], fmaf(cB[j], v[j], cC[j]));
    store8(dy + i * 8, o);
  }
}

// ---- whole BatchNorm backward of up to two views in ONE cooperative launch ------------------------
// phase 1: per-CTA partial sums of g and g*yhat over the CTA's own rows; grid barrier; fixed-order fp64
// fold of the partials; grid barrier; phase 2: dy = A*g + B*y + C over the SAME rows in reverse order
// (the tail of phase 1 is still in the 126 MB L2).  Replaces 3 launches per view (reduce, fold, apply);
// the arithmetic per element is the one of bn_reduce_kernel<.,true> / bn_bwd_apply_kernel.
struct BnBwdFusedArgs {
  const void *gin, *act, *y;
  void *dy, *gout;
  const float* mi[2];   // per view [2C]: mean, invstd
  const float* mss[2];  // per view [2C] scale, shift of the ReLU mask (or null)
  const float* gamma;
  float *dgamma, *dbeta;
  int accumulate, views, C;
  long long Mv;         // rows per view
  float* partial;       // [gridDim.x][2C]
  double* sums;         // [views][2C]
  const unsigned char* mbits;  // MASK == 3: [views * Mv][C / 8] ReLU-mask bytes written by bn_apply_kernel<., true>
};

// MASK: 0 none, 1 activation sign (three input streams), 2 recomputed from y*scale+shift, 3 one bit per element.
template <typename T, int MASK>
__global__ void __launch_bounds__(256, 2) bn_bwd_fused_kernel(BnBwdFusedArgs p) {
  cg::grid_group grid = cg::this_grid();
  extern __shared__ __align__(16) float fsm[];  // max(256*17, 5*C) floats
  const int C = p.C, tpr = C >> 3, rpi = 256 / tpr;
  const int Gv = gridDim.x / p.views, v = blockIdx.x / Gv, lb = blockIdx.x % Gv;
  const int c8 = threadIdx.x % tpr, my_r = threadIdx.x / tpr;
  const long long Mv = p.Mv, voff = (long long)v * Mv * C;
  const T* y = (const T*)p.y + voff;
  const T* gin = (const T*)p.gin + voff;
  const T* act = MASK == 1 ? (const T*)p.act + voff : nullptr;
  const float* mss = MASK == 2 ? p.mss[v] : nullptr;
  const float* mi = p.mi[v];
  const long long stride = (long long)Gv * rpi, r0 = (long long)lb * rpi + my_r;
  const long long nk = r0 < Mv ? (Mv - r0 + stride - 1) / stride : 0;
  using Raw = typename RawOf<T>::type;
  // rows in flight per thread: 2 CTAs / SM (<= 128 registers) leave 48-64 registers for raw loads, and ~100 KB
  // per SM must be in flight to cover the HBM latency (two input streams need more rows than three); the
  // values below are the largest that do not spill
  constexpr int U = sizeof(T) == 2 ? (MASK == 1 ? 5 : ((MASK == 2 || MASK == 3) ? 6 : 8)) : 2;
  const unsigned char* mbits = MASK == 3 ? p.mbits + (voff >> 3) : nullptr;  // (byte index = element offset / 8)
  float msc[8], msh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    msc[j] = MASK == 2 ? mss[c8 * 8 + j] : 0.f;
    msh[j] = MASK == 2 ? mss[C + c8 * 8 + j] : 0.f;
  }
  {
    float a0[8], a1[8], mean[8];  // a1 accumulates g*(y-mean); the common factor invstd is applied once at the end
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a0[j] = a1[j] = 0.f;
      mean[j] = mi[c8 * 8 + j];
    }
    for (long long k = 0; k < nk; k += U) {
      Raw rv[U], rg[U], ra[U];
      unsigned int rb[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k + u < nk) {
          const long long off = (r0 + (k + u) * stride) * C + c8 * 8;
          load_raw(y + off, rv[u]);
          load_raw(gin + off, rg[u]);
          if (MASK == 1) load_raw(act + off, ra[u]);
          if (MASK == 3) rb[u] = mbits[off >> 3];
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (k + u >= nk) continue;
        float vv[8], g[8], a[8];
        cvt_raw(rv[u], vv);
        cvt_raw(rg[u], g);
        if (MASK == 1) cvt_raw(ra[u], a);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float gj = g[j];
          if (MASK == 1) gj = a[j] > 0.f ? gj : 0.f;
          else if (MASK == 2) gj = fmaf(vv[j], msc[j], msh[j]) > 0.f ? gj : 0.f;
          else if (MASK == 3) gj = ((rb[u] >> j) & 1u) ? gj : 0.f;
          a0[j] += gj;
          a1[j] = fmaf(gj, vv[j] - mean[j], a1[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      fsm[threadIdx.x * 17 + j] = a0[j];
      fsm[threadIdx.x * 17 + 8 + j] = a1[j] * mi[C + c8 * 8 + j];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < tpr * 16; i += 256) {
    const int g = i / 16, j = i % 16;
    float t = 0.f;
    for (int r = 0; r < rpi; ++r) t += fsm[(r * tpr + g) * 17 + j];
    p.partial[(long long)blockIdx.x * 2 * C + (j >> 3) * C + g * 8 + (j & 7)] = t;
  }
  grid.sync();
  {  // fold: one warp per (view, entry), fixed order
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * 8, n2c = 2 * C;
    for (int e = blockIdx.x * 8 + (threadIdx.x >> 5); e < p.views * n2c; e += nwarps) {
      const int vv = e / n2c, idx = e % n2c;
      double t = 0.0;
      for (int b = lane; b < Gv; b += 32) t += (double)p.partial[(long long)(vv * Gv + b) * n2c + idx];
      t = warp_sum(t);
      if (lane == 0) p.sums[e] = t;
    }
  }
  grid.sync();
  const double invM = 1.0 / (double)Mv;
  for (int c = threadIdx.x; c < C; c += 256) {
    const double* sv = p.sums + (long long)v * 2 * C;
    const float mean = mi[c], istd = mi[C + c];
    const float m1 = (float)(sv[c] * invM), m2 = (float)(sv[C + c] * invM);
    const float A = p.gamma[c] * istd;
    fsm[c] = A;
    fsm[C + c] = -A * m2 * istd;
    fsm[2 * C + c] = -A * m1 + A * m2 * istd * mean;
    if (blockIdx.x == 0) {
      double db = 0.0, dg = 0.0;
      for (int q = 0; q < p.views; ++q) {
        db += p.sums[(long long)q * 2 * C + c];
        dg += p.sums[(long long)q * 2 * C + C + c];
      }
      if (p.dgamma) p.dgamma[c] = p.accumulate ? p.dgamma[c] + (float)dg : (float)dg;
      if (p.dbeta) p.dbeta[c] = p.accumulate ? p.dbeta[c] + (float)db : (float)db;
    }
  }
  __syncthreads();
  float cA[8], cB[8], cC[8];
  load8(fsm + c8 * 8, cA);
  load8(fsm + C + c8 * 8, cB);
  load8(fsm + 2 * C + c8 * 8, cC);
  T* dy = (T*)p.dy + voff;
  T* gout = p.gout ? (T*)p.gout + voff : nullptr;
  constexpr int U2 = U;
  for (long long k = nk - 1; k >= 0; k -= U2) {
    Raw rv[U2], rg[U2], ra[U2];
    unsigned int rb[U2];
#pragma unroll
    for (int u = 0; u < U2; ++u) {
      if (k - u >= 0) {
        const long long off = (r0 + (k - u) * stride) * C + c8 * 8;
        load_raw(y + off, rv[u]);
        load_raw(gin + off, rg[u]);
        if (MASK == 1) load_raw(act + off, ra[u]);
        if (MASK == 3) rb[u] = mbits[off >> 3];
      }
    }
#pragma unroll
    for (int u = 0; u < U2; ++u) {
      if (k - u < 0) continue;
      const long long off = (r0 + (k - u) * stride) * C + c8 * 8;
      float vv[8], g[8], a[8], o[8];
      cvt_raw(rv[u], vv);
      cvt_raw(rg[u], g);
      if (MASK == 1) {
        cvt_raw(ra[u], a);
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = a[j] > 0.f ? g[j] : 0.f;
      } else if (MASK == 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = fmaf(vv[j], msc[j], msh[j]) > 0.f ? g[j] : 0.f;
      } else if (MASK == 3) {
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = ((rb[u] >> j) & 1u) ? g[j] : 0.f;
      }
      if (gout != nullptr) store8(gout + off, g);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(cA[j], g[j], fmaf(cB[j], vv[j], cC[j]));
      store8(dy + off, o);
    }
  }
}

// AvgPool2d(full extent) + flatten (net5g.py:31-39,:56)
template <typename T>
__global__ void avgpool_kernel(const T* __restrict__ x, float* __restrict__ feat, int n, int hw, int C) {
  const long long total = (long long)n * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int ni = (int)(i / C);
    float s = 0.f;
    for (int p = 0; p < hw; ++p) s += to_f(x[((long long)ni * hw + p) * C + c]);
    feat[i] = s / (float)hw;
  }
}
template <typename T>
__global__ void avgpool_bwd_kernel(const float* __restrict__ dfeat, T* __restrict__ dx, int n, int hw, int C) {
  const long long total = (long long)n * hw * C;
  const float inv = 1.f / (float)hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int ni = (int)(i / ((long long)hw * C));
    dx[i] = from_f<T>(dfeat[(long long)ni * C + c] * inv);
  }
}

// weight repacking: torch [cout][cin][kh][kw] -> kind0 [cout][kh][kw][cin] / kind1 [cin][kh][kw][cout]
// `split` (fp32 destination only, IIC_TF32X3): a second plane of `total` elements follows the packed weights and receives
// lo = w - trunc_tf32(w); the first plane keeps the raw fp32 value (the tensor core ignores its 13 low mantissa bits)
__device__ __forceinline__ void store_lo(float* dst, long long i, float v) {
  dst[i] = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
}
__device__ __forceinline__ void store_lo(__nv_bfloat16*, long long, float) {}
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ dst, int kind, int cout, int cin,
                                   int kh, int kw, int split) {
  const long long total = (long long)cout * cin * kh * kw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    long long t = i;
    int inner, a, b, outer;
    if (kind == 0) {  // i = ((co*kh + a)*kw + b)*cin + ci
      inner = (int)(t % cin); t /= cin;
      b = (int)(t % kw); t /= kw;
      a = (int)(t % kh); outer = (int)(t / kh);
      const float v = w[(((long long)outer * cin + inner) * kh + a) * kw + b];
      dst[i] = from_f<T>(v);
      if (split) store_lo(dst, total + i, v);
    } else {  // i = ((ci*kh + a)*kw + b)*cout + co
      inner = (int)(t % cout); t /= cout;
      b = (int)(t % kw); t /= kw;
      a = (int)(t % kh); outer = (int)(t / kh);
      const float v = w[(((long long)inner * cin + outer) * kh + a) * kw + b];
      dst[i] = from_f<T>(v);
      if (split) store_lo(dst, total + i, v);
    }
  }
}
// all convolutions of a trunk in one launch: blockIdx.y = job, blockIdx.x strides over the job's elements
template <typename T>
__global__ void pack_weights_batched_kernel(const iic_pack_job* __restrict__ jobs, int split) {
  const iic_pack_job j = jobs[blockIdx.y];
  const float* __restrict__ w = j.w;
  T* __restrict__ dst = (T*)j.dst;
  const int cin = j.cin, cout = j.cout, kh = j.kh, kw = j.kw;
  const int total = cout * cin * kh * kw;
  const int inner_n = j.kind == 0 ? cin : cout;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    int t = i;
    const int inner = t % inner_n; t /= inner_n;
    const int b = t % kw; t /= kw;
    const int a = t % kh;
    const int outer = t / kh;
    const int co = j.kind == 0 ? outer : inner, ci = j.kind == 0 ? inner : outer;
    const float v = w[((co * cin + ci) * kh + a) * kw + b];
    dst[i] = from_f<T>(v);
    if (split) store_lo(dst, (long long)total + i, v);
  }
}
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, float* __restrict__ grad, int accumulate, int cout,
                                    int cin, int kh, int kw) {
  const long long total = (long long)cout * cin * kh * kw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i indexes torch layout [co][ci][a][b]
    long long t = i;
    const int b = (int)(t % kw); t /= kw;
    const int a = (int)(t % kh); t /= kh;
    const int ci = (int)(t % cin);
    const int co = (int)(t / cin);
    const float v = dwp[(((long long)co * kh + a) * kw + b) * cin + ci];
    grad[i] = accumulate ? grad[i] + v : v;
  }
}

}  // namespace iic

using namespace iic;

#define DISPATCH_T(dtype, ...)                                  \
  if ((dtype) == IIC_F32) { using T = float; __VA_ARGS__ }       \
  else if ((dtype) == IIC_BF16) { using T = __nv_bfloat16; __VA_ARGS__ } \
  else { set_error("unknown dtype %d", (int)(dtype)); return IIC_ERR_BAD_ARG; }

extern "C" int iic_nchw_to_nhwc(const float* src, void* dst, int dst_dtype, int n, int c, int h, int w, void* stream) {
  IIC_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0, IIC_ERR_BAD_ARG, "iic_nchw_to_nhwc: bad arguments");
  const long long total = (long long)n * c * h * w;
  DISPATCH_T(dst_dtype, nchw_to_nhwc_kernel<T><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(src, (T*)dst, n, c, h, w);)
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}
extern "C" int iic_nhwc_to_nchw(const void* src, int src_dtype, float* dst, int n, int c, int h, int w, void* stream) {
  IIC_REQUIRE(src && dst && n > 0 && c > 0 && h > 0 && w > 0, IIC_ERR_BAD_ARG, "iic_nhwc_to_nchw: bad arguments");
  const long long total = (long long)n * c * h * w;
  DISPATCH_T(src_dtype, nhwc_to_nchw_kernel<T><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const T*)src, dst, n, c, h, w);)
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}
extern "C" int iic_cast(const void* src, int src_dtype, void* dst, int dst_dtype, long long count, void* stream) {
  IIC_REQUIRE(src && dst && count > 0, IIC_ERR_BAD_ARG, "iic_cast: bad arguments");
  const int g = ew_grid(count, 256);
  cudaStream_t st = (cudaStream_t)stream;
  if (src_dtype == IIC_F32 && dst_dtype == IIC_BF16)
    cast_kernel<float, __nv_bfloat16><<<g, 256, 0, st>>>((const float*)src, (__nv_bfloat16*)dst, count);
  else if (src_dtype == IIC_BF16 && dst_dtype == IIC_F32)
    cast_kernel<__nv_bfloat16, float><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, (float*)dst, count);
  else if (src_dtype == IIC_F32 && dst_dtype == IIC_F32)
    cast_kernel<float, float><<<g, 256, 0, st>>>((const float*)src, (float*)dst, count);
  else if (src_dtype == IIC_BF16 && dst_dtype == IIC_BF16)
    cast_kernel<__nv_bfloat16, __nv_bfloat16><<<g, 256, 0, st>>>((const __nv_bfloat16*)src, (__nv_bfloat16*)dst, count);
  else {
    set_error("iic_cast: bad dtypes");
    return IIC_ERR_BAD_ARG;
  }
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_sobel(const float* imgs, float* out, int n, int c_in, int h, int w, int include_rgb, int using_ir,
                         void* stream) {
  IIC_REQUIRE(imgs && out && n > 0 && h > 0 && w > 0, IIC_ERR_BAD_ARG, "iic_sobel: bad arguments");
  int grey, nrgb, ir, expect;
  if (!using_ir) {
    if (!include_rgb) { expect = 1; grey = 0; nrgb = 0; ir = -1; }
    else { expect = 4; grey = 3; nrgb = 3; ir = -1; }
  } else {
    if (!include_rgb) { expect = 2; grey = 0; nrgb = 0; ir = 1; }
    else { expect = 5; grey = 3; nrgb = 3; ir = 4; }
  }
  // the reference asserts the channel count (transforms.py:52,56,60,64)
  IIC_REQUIRE(c_in == expect, IIC_ERR_BAD_ARG, "iic_sobel: expected %d input channels, got %d", expect, c_in);
  const int cout = nrgb + 2 + (ir >= 0 ? 1 : 0);
  sobel_kernel<<<ew_grid((long long)n * h * w, 256), 256, 0, (cudaStream_t)stream>>>(imgs, out, n, c_in, cout, h, w,
                                                                                     grey, nrgb, ir);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

// ---------------------------------------------------------------------------------------------
// Dataloader tail fused with sobel_process: RGB image -> grey -> [dx, dy].
// The reference converts to grey on the CPU inside the dataloader (custom_greyscale_to_tensor,
// code/utils/cluster/transforms.py:12-16: PIL "L" = (19595 R + 38470 G + 7471 B + 0x8000) >> 16 on uint8, then / 255),
// uploads the grey batch and runs sobel_process on it.  Here the RGB batch (uint8 as the dataloader holds it, or fp32
// in [0,1]) is uploaded once and grey + both Sobel filters are evaluated per output pixel.
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float grey_of(const T* r, const T* g, const T* b, long long i);
template <> __device__ __forceinline__ float grey_of<float>(const float* r, const float* g, const float* b, long long i) {
  return 0.299f * r[i] + 0.587f * g[i] + 0.114f * b[i];
}
template <> __device__ __forceinline__ float grey_of<unsigned char>(const unsigned char* r, const unsigned char* g,
                                                                    const unsigned char* b, long long i) {
  const unsigned int l = (19595u * r[i] + 38470u * g[i] + 7471u * b[i] + 0x8000u) >> 16;
  return (float)l / 255.f;  // tf.to_tensor
}

template <typename T>
__global__ void grey_sobel_kernel(const T* __restrict__ in, float* __restrict__ out, int n, int h, int w) {
  const long long total = (long long)n * h * w, plane = (long long)h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    int x, y;
    long long ni;
    if (total <= 0xffffffffll) {  // 32-bit divisions
      const unsigned u = (unsigned)i, row = u / (unsigned)w;
      x = (int)(u - row * (unsigned)w);
      const unsigned im = row / (unsigned)h;
      y = (int)(row - im * (unsigned)h);
      ni = im;
    } else {
      x = (int)(i % w);
      y = (int)((i / w) % h);
      ni = i / plane;
    }
    const T* r = in + ni * 3 * plane;
    float v[3][3];
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int yy = y + dy, xx = x + dx;
        v[dy + 1][dx + 1] =
            (yy >= 0 && yy < h && xx >= 0 && xx < w) ? grey_of<T>(r, r + plane, r + 2 * plane, (long long)yy * w + xx) : 0.f;
      }
    const float sx = v[0][0] - v[0][2] + 2.f * v[1][0] - 2.f * v[1][2] + v[2][0] - v[2][2];
    const float sy = v[0][0] + 2.f * v[0][1] + v[0][2] - v[2][0] - 2.f * v[2][1] - v[2][2];
    float* o = out + ni * 2 * plane + (long long)y * w + x;
    o[0] = sx;
    o[plane] = sy;
  }
}

extern "C" int iic_grey_sobel(const void* rgb, int src_is_u8, float* out, int n, int h, int w, void* stream) {
  IIC_REQUIRE(rgb && out && n > 0 && h > 0 && w > 0, IIC_ERR_BAD_ARG, "iic_grey_sobel: bad arguments");
  const int grid = ew_grid((long long)n * h * w, 256);
  if (src_is_u8)
    grey_sobel_kernel<unsigned char><<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)rgb, out, n, h, w);
  else
    grey_sobel_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)rgb, out, n, h, w);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

static int bn_reduce_blocks(long long M, int C) {
  const int rpi = 256 / (C / 8);
  long long blocks = (M + 4 * rpi - 1) / (4 * rpi);
  const long long cap = (long long)device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// scratch for the per-block partial sums: [blocks][2C] floats, kept per device (grown on demand;
// stream-ordered use only, like every workspace of this library)
static float* bn_partial_scratch(size_t bytes) {
  static float* buf[64] = {nullptr};
  static size_t cap[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (cap[dev] < bytes) {
    if (buf[dev]) cudaFree(buf[dev]);
    buf[dev] = nullptr;
    cap[dev] = 0;
    if (cudaMalloc(&buf[dev], bytes) != cudaSuccess) return nullptr;
    cap[dev] = bytes;
  }
  return buf[dev];
}

extern "C" int iic_bn_stats(const void* y, int dtype, long long M, int C, const float* gamma, const float* beta,
                            float eps, float momentum, float* running_mean, float* running_var, int use_running,
                            double* stats_ws, float* scale_shift, float* mean_invstd, void* stream) {
  IIC_REQUIRE(y && gamma && beta && stats_ws && scale_shift && mean_invstd && M > 0, IIC_ERR_BAD_ARG,
              "iic_bn_stats: bad arguments");
  IIC_REQUIRE(C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0, IIC_ERR_UNSUPPORTED,
              "iic_bn_stats: C=%d must be 8 * (a divisor of 256)", C);
  IIC_REQUIRE(!use_running || (running_mean && running_var), IIC_ERR_BAD_ARG,
              "iic_bn_stats: eval mode needs running statistics");
  cudaStream_t st = (cudaStream_t)stream;
  if (!use_running) {
    const int blocks = bn_reduce_blocks(M, C);
    float* partial = bn_partial_scratch((size_t)device_sm_count() * 8 * 2 * 2048 * sizeof(float));
    IIC_REQUIRE(partial != nullptr, IIC_ERR_CUDA, "iic_bn_stats: scratch allocation failed");
    DISPATCH_T(dtype, bn_reduce_kernel<T, false><<<blocks, 256, 0, st>>>((const T*)y, nullptr, nullptr, nullptr, nullptr, M, C, partial);)
    IIC_LAUNCH_CHECK();
    count_launch();
    bn_sum_partials_kernel<<<cdiv(2 * C * 32, 256), 256, 0, st>>>(partial, blocks, 2 * C, 2 * C, stats_ws);
    IIC_LAUNCH_CHECK();
    count_launch();
  }
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>(stats_ws, M, C, gamma, beta, eps, momentum, running_mean,
                                                    running_var, use_running, scale_shift, mean_invstd);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

// statistics from the per-CTA partials written by iic_conv_fprop_stats: partial[blk][views][2][C]
extern "C" int iic_bn_stats_from_partials_views(const float* stat_partial, int nblk, int slots, int views, long long M_per_view,
                                                int C, const float* gamma, const float* beta, float eps,
                                                float momentum, float* running_mean, float* running_var,
                                                float* scale_shift, float* mean_invstd, void* stream) {
  IIC_REQUIRE(stat_partial && nblk > 0 && views >= 1 && slots >= views && gamma && beta && scale_shift && mean_invstd && M_per_view > 0 &&
                  C > 0,
              IIC_ERR_BAD_ARG, "iic_bn_stats_from_partials_views: bad arguments");
  bn_fold_finalize_kernel<<<cdiv(C, 32), 256, 0, (cudaStream_t)stream>>>(
      stat_partial, nblk, slots, views, M_per_view, C, gamma, beta, eps, momentum, running_mean, running_var,
      scale_shift, mean_invstd);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_bn_stats_from_partials(const float* stat_partial, int nblk, int views, int view, long long M, int C,
                                          const float* gamma, const float* beta, float eps, float momentum,
                                          float* running_mean, float* running_var, double* stats_ws, float* scale_shift,
                                          float* mean_invstd, void* stream) {
  IIC_REQUIRE(stat_partial && nblk > 0 && views >= 1 && view >= 0 && view < views && gamma && beta && stats_ws &&
                  scale_shift && mean_invstd && M > 0 && C > 0,
              IIC_ERR_BAD_ARG, "iic_bn_stats_from_partials: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  bn_sum_partials_kernel<<<cdiv(2 * C * 32, 256), 256, 0, st>>>(stat_partial + (long long)view * 2 * C, nblk, 2 * C,
                                                                 (long long)views * 2 * C, stats_ws);
  IIC_LAUNCH_CHECK();
  count_launch();
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, st>>>(stats_ws, M, C, gamma, beta, eps, momentum, running_mean, running_var, 0,
                                                    scale_shift, mean_invstd);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_bn_apply_views(const void* y, const float* scale_shift, const void* res,
                                  const float* res_scale_shift, void* out, int dtype, long long M_per_view, int C,
                                  int relu, int views, void* stream) {
  IIC_REQUIRE(y && scale_shift && out && M_per_view > 0 && C % 8 == 0 && views >= 1 && views <= 8, IIC_ERR_BAD_ARG,
              "iic_bn_apply: bad arguments");
  IIC_REQUIRE(256 % (C / 8) == 0, IIC_ERR_UNSUPPORTED, "iic_bn_apply: C=%d must be 8 * (a divisor of 256)", C);
  const long long total = M_per_view * (C / 8);
  const int grid = ew_grid(total, 256) * views;  // ew_grid caps at 16 CTAs per SM per view
  DISPATCH_T(dtype, bn_apply_kernel<T><<<grid, 256, 0, (cudaStream_t)stream>>>(
      (const T*)y, scale_shift, (const T*)res, res_scale_shift, (T*)out, M_per_view, C, relu, views);)
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_bn_apply(const void* y, const float* scale_shift, const void* res, const float* res_scale_shift,
                            void* out, int dtype, long long M, int C, int relu, void* stream) {
  return iic_bn_apply_views(y, scale_shift, res, res_scale_shift, out, dtype, M, C, relu, 1, stream);
}

extern "C" int iic_bn_relu_maxpool(const void* y, const float* scale_shift, void* out, int dtype, int n, int h, int w,
                                   int C, int pad, int oh, int ow, void* stream) {
  IIC_REQUIRE(y && scale_shift && out && C % 8 == 0, IIC_ERR_BAD_ARG, "iic_bn_relu_maxpool: bad arguments");
  IIC_REQUIRE(oh == (h + 2 * pad - 2) / 2 + 1 && ow == (w + 2 * pad - 2) / 2 + 1 && pad >= 0 && pad <= 1,
              IIC_ERR_BAD_ARG, "iic_bn_relu_maxpool: inconsistent geometry");
  const long long total = (long long)n * oh * ow * (C / 8);
  DISPATCH_T(dtype, bn_relu_maxpool_kernel<T><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
      (const T*)y, scale_shift, (T*)out, n, h, w, C, pad, oh, ow);)
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_bn_relu_maxpool_bwd(const void* y, const float* scale_shift, const void* dpool, void* g, int dtype,
                                       int n, int h, int w, int C, int pad, int oh, int ow, void* stream) {
  IIC_REQUIRE(y && scale_shift && dpool && g && C % 8 == 0, IIC_ERR_BAD_ARG, "iic_bn_relu_maxpool_bwd: bad arguments");
  IIC_REQUIRE(oh == (h + 2 * pad - 2) / 2 + 1 && ow == (w + 2 * pad - 2) / 2 + 1 && pad >= 0 && pad <= 1,
              IIC_ERR_BAD_ARG, "iic_bn_relu_maxpool_bwd: inconsistent geometry");
  cudaStream_t st = (cudaStream_t)stream;
  const bool covered = (2 * oh - pad >= h) && (2 * ow - pad >= w);
  const size_t esz = dtype == IIC_F32 ? 4 : 2;
  if (!covered) IIC_CUDA(cudaMemsetAsync(g, 0, (size_t)n * h * w * C * esz, st));
  const long long total = (long long)n * oh * ow * (C / 8);
  DISPATCH_T(dtype, bn_relu_maxpool_bwd_kernel<T><<<ew_grid(total, 256), 256, 0, st>>>(
      (const T*)y, scale_shift, (const T*)dpool, (T*)g, n, h, w, C, pad, oh, ow);)
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_bn_bwd_reduce(const void* g_in, const void* act, const float* mask_scale_shift, const void* y,
                                 const float* mean_invstd, int dtype, long long M, int C, double* sums, void* stream) {
  IIC_REQUIRE(g_in && y && mean_invstd && sums && M > 0, IIC_ERR_BAD_ARG, "iic_bn_bwd_reduce: bad arguments");
  IIC_REQUIRE(C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0, IIC_ERR_UNSUPPORTED, "iic_bn_bwd_reduce: C=%d unsupported", C);
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = bn_reduce_blocks(M, C);
  float* partial = bn_partial_scratch((size_t)device_sm_count() * 8 * 2 * 2048 * sizeof(float));
  IIC_REQUIRE(partial != nullptr, IIC_ERR_CUDA, "iic_bn_bwd_reduce: scratch allocation failed");
  DISPATCH_T(dtype, bn_reduce_kernel<T, true><<<blocks, 256, 0, st>>>((const T*)y, (const T*)g_in, (const T*)act, mask_scale_shift, mean_invstd, M, C, partial);)
  IIC_LAUNCH_CHECK();
  count_launch();
  bn_sum_partials_kernel<<<cdiv(2 * C * 32, 256), 256, 0, st>>>(partial, blocks, 2 * C, 2 * C, sums);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_bn_bwd_apply(const void* g_in, const void* act, const float* mask_scale_shift, const void* y,
                                const float* mean_invstd, const float* gamma, const double* sums, void* dy, void* g_out,
                                float* dgamma,
                                float* dbeta, int accumulate, int dtype, long long M, int C, void* stream) {
  IIC_REQUIRE(g_in && y && mean_invstd && gamma && sums && dy && M > 0 && C % 8 == 0, IIC_ERR_BAD_ARG,
              "iic_bn_bwd_apply: bad arguments");
  IIC_REQUIRE(256 % (C / 8) == 0, IIC_ERR_UNSUPPORTED, "iic_bn_bwd_apply: C=%d must be 8 * (a divisor of 256)", C);
  const long long total = M * (C / 8);
  const size_t smem = (size_t)5 * C * sizeof(float);
  DISPATCH_T(dtype, bn_bwd_apply_kernel<T><<<ew_grid(total, 256), 256, smem, (cudaStream_t)stream>>>(
      (const T*)g_in, (const T*)act, (const T*)y, mask_scale_shift, mean_invstd, gamma, sums, (T*)dy, (T*)g_out, dgamma, dbeta,
      accumulate, M, C);)
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

static double* bn_sums_scratch() {
  static double* buf[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!buf[dev] && cudaMalloc(&buf[dev], 2 * 2 * 2048 * sizeof(double)) != cudaSuccess) buf[dev] = nullptr;
  return buf[dev];
}

template <typename T, int MASK>
static int bn_bwd_fused_launch(BnBwdFusedArgs& a, size_t smem, cudaStream_t st) {
  static int cached[64] = {0};
  int dev = 0;
  IIC_CUDA(cudaGetDevice(&dev));
  IIC_REQUIRE(dev >= 0 && dev < 64, IIC_ERR_CUDA, "iic_bn_bwd_fused: device index");
  if (!cached[dev]) {
    int occ = 0;  // worst-case dynamic smem (C = 2048) so one answer serves every C
    IIC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bn_bwd_fused_kernel<T, MASK>, 256, 5 * 2048 * sizeof(float)));
    cached[dev] = occ > 2 ? 2 : occ;
    IIC_REQUIRE(cached[dev] > 0, IIC_ERR_CUDA, "iic_bn_bwd_fused: kernel does not fit");
  }
  int per_sm = option(OPT_BN_BWD_CTAS);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > cached[dev]) per_sm = cached[dev];
  int grid = device_sm_count() * per_sm;
  grid -= grid % a.views;
  void* args[] = {&a};
  cudaError_t e = cudaLaunchCooperativeKernel((void*)bn_bwd_fused_kernel<T, MASK>, dim3(grid), dim3(256), args, smem, st);
  IIC_REQUIRE(e == cudaSuccess, IIC_ERR_CUDA, "iic_bn_bwd_fused: cooperative launch failed: %s", cudaGetErrorString(e));
  count_launch();
  return IIC_OK;
}

template <typename T>
static int bn_bwd_fused_dispatch(BnBwdFusedArgs& a, size_t smem, cudaStream_t st) {
  if (a.mbits != nullptr) return bn_bwd_fused_launch<T, 3>(a, smem, st);
  if (a.act != nullptr) return bn_bwd_fused_launch<T, 1>(a, smem, st);
  if (a.mss[0] != nullptr) return bn_bwd_fused_launch<T, 2>(a, smem, st);
  return bn_bwd_fused_launch<T, 0>(a, smem, st);
}

extern "C" int iic_bn_bwd_fused(const void* g_in, const void* act, const void* y, int views, const float* mean_invstd0,
                                const float* mean_invstd1, const float* mask_ss0, const float* mask_ss1,
                                const float* gamma, void* dy, void* g_out, float* dgamma, float* dbeta, int accumulate,
                                int dtype, long long M_per_view, int C, void* stream) {
  IIC_REQUIRE(g_in && y && mean_invstd0 && gamma && dy && M_per_view > 0 && (views == 1 || (views == 2 && mean_invstd1)),
              IIC_ERR_BAD_ARG, "iic_bn_bwd_fused: bad arguments");
  IIC_REQUIRE(C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0, IIC_ERR_UNSUPPORTED, "iic_bn_bwd_fused: C=%d unsupported", C);
  IIC_REQUIRE(views == 1 || ((mask_ss0 == nullptr) == (mask_ss1 == nullptr)), IIC_ERR_BAD_ARG,
              "iic_bn_bwd_fused: mask scale/shift must be given for both views or neither");
  IIC_REQUIRE(!(act && mask_ss0), IIC_ERR_BAD_ARG, "iic_bn_bwd_fused: one ReLU mask source at most");
  IIC_REQUIRE(dtype == IIC_BF16 || dtype == IIC_F32, IIC_ERR_BAD_ARG, "iic_bn_bwd_fused: bad dtype %d", dtype);
  BnBwdFusedArgs a;
  a.mbits = nullptr;
  a.gin = g_in; a.act = act; a.y = y; a.dy = dy; a.gout = g_out;
  a.mi[0] = mean_invstd0; a.mi[1] = mean_invstd1; a.mss[0] = mask_ss0; a.mss[1] = mask_ss1;
  a.gamma = gamma; a.dgamma = dgamma; a.dbeta = dbeta; a.accumulate = accumulate; a.views = views; a.C = C;
  a.Mv = M_per_view;
  a.partial = bn_partial_scratch((size_t)device_sm_count() * 8 * 2 * 2048 * sizeof(float));
  a.sums = bn_sums_scratch();
  IIC_REQUIRE(a.partial && a.sums, IIC_ERR_CUDA, "iic_bn_bwd_fused: scratch allocation failed");
  const size_t smem = sizeof(float) * (size_t)(5 * C > 256 * 17 ? 5 * C : 256 * 17);
  if (dtype == IIC_BF16) return bn_bwd_fused_dispatch<__nv_bfloat16>(a, smem, (cudaStream_t)stream);
  return bn_bwd_fused_dispatch<float>(a, smem, (cudaStream_t)stream);
}

// The residual block's output ReLU as a 1-bit mask: iic_bn_apply_views_mask = iic_bn_apply_views(relu = 1) that also
// writes one mask byte per 8 channels; iic_bn_bwd_fused_bits = iic_bn_bwd_fused with that mask instead of `act`.
extern "C" int iic_bn_apply_views_mask(const void* y, const float* scale_shift, const void* res,
                                       const float* res_scale_shift, void* out, unsigned char* mask_out, int dtype,
                                       long long M_per_view, int C, int views, void* stream) {
  IIC_REQUIRE(y && scale_shift && out && mask_out && M_per_view > 0 && C % 8 == 0 && views >= 1 && views <= 8, IIC_ERR_BAD_ARG,
              "iic_bn_apply_views_mask: bad arguments");
  IIC_REQUIRE(256 % (C / 8) == 0, IIC_ERR_UNSUPPORTED, "iic_bn_apply_views_mask: C=%d must be 8 * (a divisor of 256)", C);
  const long long total = M_per_view * (C / 8);
  const int grid = ew_grid(total, 256) * views;
  DISPATCH_T(dtype, (bn_apply_kernel<T, true><<<grid, 256, 0, (cudaStream_t)stream>>>(
      (const T*)y, scale_shift, (const T*)res, res_scale_shift, (T*)out, M_per_view, C, 1, views, mask_out));)
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_bn_bwd_fused_bits(const void* g_in, const unsigned char* mask_bits, const void* y, int views,
                                     const float* mean_invstd0, const float* mean_invstd1, const float* gamma, void* dy,
                                     void* g_out, float* dgamma, float* dbeta, int accumulate, int dtype,
                                     long long M_per_view, int C, void* stream) {
  IIC_REQUIRE(g_in && mask_bits && y && mean_invstd0 && gamma && dy && M_per_view > 0 &&
                  (views == 1 || (views == 2 && mean_invstd1)),
              IIC_ERR_BAD_ARG, "iic_bn_bwd_fused_bits: bad arguments");
  IIC_REQUIRE(C % 8 == 0 && C <= 2048 && 256 % (C / 8) == 0, IIC_ERR_UNSUPPORTED, "iic_bn_bwd_fused_bits: C=%d unsupported", C);
  IIC_REQUIRE(dtype == IIC_BF16 || dtype == IIC_F32, IIC_ERR_BAD_ARG, "iic_bn_bwd_fused_bits: bad dtype %d", dtype);
  BnBwdFusedArgs a;
  a.mbits = mask_bits;
  a.gin = g_in; a.act = nullptr; a.y = y; a.dy = dy; a.gout = g_out;
  a.mi[0] = mean_invstd0; a.mi[1] = mean_invstd1; a.mss[0] = nullptr; a.mss[1] = nullptr;
  a.gamma = gamma; a.dgamma = dgamma; a.dbeta = dbeta; a.accumulate = accumulate; a.views = views; a.C = C;
  a.Mv = M_per_view;
  a.partial = bn_partial_scratch((size_t)device_sm_count() * 8 * 2 * 2048 * sizeof(float));
  a.sums = bn_sums_scratch();
  IIC_REQUIRE(a.partial && a.sums, IIC_ERR_CUDA, "iic_bn_bwd_fused_bits: scratch allocation failed");
  const size_t smem = sizeof(float) * (size_t)(5 * C > 256 * 17 ? 5 * C : 256 * 17);
  if (dtype == IIC_BF16) return bn_bwd_fused_dispatch<__nv_bfloat16>(a, smem, (cudaStream_t)stream);
  return bn_bwd_fused_dispatch<float>(a, smem, (cudaStream_t)stream);
}

extern "C" int iic_avgpool(const void* x, int dtype, float* feat, int n, int hw, int C, void* stream) {
  IIC_REQUIRE(x && feat && n > 0 && hw > 0 && C > 0, IIC_ERR_BAD_ARG, "iic_avgpool: bad arguments");
  DISPATCH_T(dtype, avgpool_kernel<T><<<ew_grid((long long)n * C, 128), 128, 0, (cudaStream_t)stream>>>((const T*)x, feat, n, hw, C);)
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}
extern "C" int iic_avgpool_bwd(const float* dfeat, void* dx, int dtype, int n, int hw, int C, void* stream) {
  IIC_REQUIRE(dfeat && dx && n > 0 && hw > 0 && C > 0, IIC_ERR_BAD_ARG, "iic_avgpool_bwd: bad arguments");
  DISPATCH_T(dtype, avgpool_bwd_kernel<T><<<ew_grid((long long)n * hw * C, 256), 256, 0, (cudaStream_t)stream>>>(dfeat, (T*)dx, n, hw, C);)
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_pack_weight(const float* w_oihw, void* dst, int dst_dtype, int kind, int cout, int cin, int kh,
                               int kw, void* stream) {
  IIC_REQUIRE(w_oihw && dst && (kind == 0 || kind == 1) && cout > 0 && cin > 0 && kh > 0 && kw > 0, IIC_ERR_BAD_ARG,
              "iic_pack_weight: bad arguments");
  const long long total = (long long)cout * cin * kh * kw;
  const int split = dst_dtype == IIC_TF32X3 ? 1 : 0;  // fp32 [2][...]: raw plane + lo plane (include/iic_b200.h)
  if (dst_dtype == IIC_TF32X3 || dst_dtype == IIC_TF32) dst_dtype = IIC_F32;
  DISPATCH_T(dst_dtype, pack_weight_kernel<T><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(w_oihw, (T*)dst, kind, cout, cin, kh, kw, split);)
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}
extern "C" int iic_pack_weights_batched(const iic_pack_job* jobs_device, int njobs, int dst_dtype, void* stream) {
  IIC_REQUIRE(jobs_device && njobs > 0 && njobs <= 65535, IIC_ERR_BAD_ARG, "iic_pack_weights_batched: bad arguments");
  const int split = dst_dtype == IIC_TF32X3 ? 1 : 0;
  if (dst_dtype == IIC_TF32X3 || dst_dtype == IIC_TF32) dst_dtype = IIC_F32;
  DISPATCH_T(dst_dtype, pack_weights_batched_kernel<T><<<dim3(32, njobs), 256, 0, (cudaStream_t)stream>>>(jobs_device, split);)
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}
extern "C" int iic_unpack_wgrad(const float* dw_packed, float* grad_oihw, int accumulate, int cout, int cin, int kh,
                                int kw, void* stream) {
  IIC_REQUIRE(dw_packed && grad_oihw && cout > 0 && cin > 0, IIC_ERR_BAD_ARG, "iic_unpack_wgrad: bad arguments");
  const long long total = (long long)cout * cin * kh * kw;
  unpack_wgrad_kernel<<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(dw_packed, grad_oihw, accumulate, cout,
                                                                              cin, kh, kw);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}
