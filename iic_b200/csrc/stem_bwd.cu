// Fused backward of the ClusterNet5g stem:  conv3x3(cin -> 64, pad 1, no bias) -> BatchNorm -> ReLU -> MaxPool(2, 2, pad)
// (net5g.py:21-26; backward of cluster_sobel_twohead.py:354 through these four modules).
//
// The stem's conv output y0 is the largest tensor of the step (1408 x 96 x 96 x 64 bf16 = 1.66 GB at the bench shape).
// The generic chain walks it six times: max-pool backward (reads y0, writes the routed gradient g), BatchNorm backward
// reduce (reads y0, g), BatchNorm backward apply (reads y0, g, writes dy), stem wgrad (reads dy) = 13.7 GB.
// Nothing downstream needs g or dy themselves (the network input takes no gradient), so here both are recomputed on the
// fly from (y0, dpool) in the two passes that need them:
//   pass A  stem_bwd_reduce_kernel   sum g, sum g*(y - mean) per channel and view            (reads y0, dpool: 2.1 GB)
//   fold    stem_bwd_fold_kernel     per-block partials -> fp64 sums (fixed order)
//   pass B  stem_bwd_wgrad_kernel    g -> dy = A*g + B*y + C -> dW[co][ci][a][b] += dy * x     (reads y0, dpool, x: 2.2 GB)
//   reduce  stem_bwd_dw_reduce_kernel per-block dW partials -> grad
// A work item is one pooling window (its <= 4 pixels decide the arg-max) x 4 channels; 16 consecutive lanes cover the
// 64 channels of a pixel, so every activation load of a warp is two whole 128-byte (bf16) pixel rows.  The arithmetic
// follows bn_relu_maxpool_bwd_kernel / bn_bwd_fused_kernel / stem_wgrad_kernel (elementwise.cu, stem.cu) term by term;
// only dy stays fp32 instead of being rounded to the storage type.
#include "common.cuh"

namespace iic {


__device__ __forceinline__ void load4(const float* p, float (&f)[4]) {
  const float4 v = *reinterpret_cast<const float4*>(p);
  f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
}
__device__ __forceinline__ void load4(const __nv_bfloat16* p, float (&f)[4]) {
  const uint2 v = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
}

__device__ __forceinline__ void store4(float* p, const float (&f)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
}
__device__ __forceinline__ void store4(__nv_bfloat16* p, const float (&f)[4]) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(f[0], f[1]), b = __floats2bfloat162_rn(f[2], f[3]);
  uint2 v;
  v.x = *reinterpret_cast<const uint32_t*>(&a);
  v.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = v;
}

struct StemBwdArgs {
  const float* x;     // [n][cin][H][W] network input (fp32, NCHW)
  const void* y;      // [n][H][W][64] stem conv output
  const void* dpool;  // [n][oh][ow][64] gradient of the pooled activation
  const float* ss;    // [views][2*64] BN scale, shift (the ReLU / arg-max are recomputed from y*scale + shift)
  const float* mi;    // [views][2*64] batch mean, 1/std
  const float* gamma;
  float *dgamma, *dbeta;
  int bn_accumulate;
  int n, H, W, pad, oh, ow, views;
  float* partialA;  // [gridA][128]
  double* sums;     // [views][128]: sum g, sum g*yhat
  float* partialB;  // [gridB][64 * cin * 9]
  void* dy_out;     // stem_bwd_dy_kernel: [n][H][W][64] gradient of the conv output, storage type
};

// One pooling window x 4 channels: loads dpool and the window's valid pixels of y, returns the routed and ReLU-masked
// gradient g (torch's max_pool2d tie rule: first maximum in row-major window order, strict '>').
template <typename T>
__device__ __forceinline__ void window_grad(const T* __restrict__ y, const T* __restrict__ dpool, int img, int oy, int ox, int H,
                                            int W, int oh, int ow, int pad, int cg, const float (&sc)[4],
                                            const float (&sh)[4], float (&yv)[4][4], bool (&valid)[4], float (&g)[4][4]) {
  float dp[4], m[4];
  int arg[4];
  load4(dpool + (((long long)img * oh + oy) * ow + ox) * 64 + cg * 4, dp);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    m[j] = -INFINITY;
    arg[j] = -1;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int iy = oy * 2 - pad + (q >> 1), ix = ox * 2 - pad + (q & 1);
    valid[q] = iy >= 0 && iy < H && ix >= 0 && ix < W;
    if (valid[q]) {
      load4(y + (((long long)img * H + iy) * W + ix) * 64 + cg * 4, yv[q]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = fmaxf(fmaf(yv[q][j], sc[j], sh[j]), 0.f);
        if (a > m[j]) {
          m[j] = a;
          arg[j] = q;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) yv[q][j] = 0.f;
    }
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int j = 0; j < 4; ++j) g[q][j] = (valid[q] && arg[j] == q && m[j] > 0.f) ? dp[j] : 0.f;
}

// grid = views * Gv blocks of 256 threads = 16 windows x 16 channel groups per iteration
template <typename T>
__global__ void __launch_bounds__(256) stem_bwd_reduce_kernel(StemBwdArgs p) {
  __shared__ float red[256][9];
  const int Gv = gridDim.x / p.views, v = blockIdx.x / Gv, lb = blockIdx.x % Gv;
  const int cg = threadIdx.x & 15, wslot = threadIdx.x >> 4;
  const long long Wv = (long long)(p.n / p.views) * p.oh * p.ow;  // windows per view
  const T* y = (const T*)p.y;
  const T* dpool = (const T*)p.dpool;
  float sc[4], sh[4], mean[4], a0[4], a1[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    sc[j] = p.ss[v * 128 + cg * 4 + j];
    sh[j] = p.ss[v * 128 + 64 + cg * 4 + j];
    mean[j] = p.mi[v * 128 + cg * 4 + j];
    a0[j] = a1[j] = 0.f;
  }
  for (long long wl = (long long)lb * 16 + wslot; wl < Wv; wl += (long long)Gv * 16) {
    const long long wi = (long long)v * Wv + wl;
    int ox, oy, img;
    if ((long long)p.views * Wv <= 0xffffffffll) {  // 32-bit divisions (the 64-bit ones cost more than the memory traffic)
      const unsigned u = (unsigned)wi, t = u / (unsigned)p.ow;
      ox = (int)(u - t * (unsigned)p.ow);
      img = (int)(t / (unsigned)p.oh);
      oy = (int)(t - (unsigned)img * (unsigned)p.oh);
    } else {
      ox = (int)(wi % p.ow);
      const long long t = wi / p.ow;
      oy = (int)(t % p.oh);
      img = (int)(t / p.oh);
    }
    float yv[4][4], g[4][4];
    bool valid[4];
    window_grad<T>(y, dpool, img, oy, ox, p.H, p.W, p.oh, p.ow, p.pad, cg, sc, sh, yv, valid, g);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a0[j] += g[q][j];
        a1[j] = fmaf(g[q][j], yv[q][j] - mean[j], a1[j]);  // (g is 0 at invalid positions)
      }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    red[threadIdx.x][j] = a0[j];
    red[threadIdx.x][4 + j] = a1[j] * p.mi[v * 128 + 64 + cg * 4 + j];
  }
  __syncthreads();
  if (threadIdx.x < 128) {  // (stat, channel)
    const int stat = threadIdx.x >> 6, ch = threadIdx.x & 63;
    float t = 0.f;
#pragma unroll
    for (int ws = 0; ws < 16; ++ws) t += red[ws * 16 + (ch >> 2)][stat * 4 + (ch & 3)];
    p.partialA[(long long)blockIdx.x * 128 + threadIdx.x] = t;
  }
}

// one warp per (view, entry): fixed-order fp64 fold of the per-block partials
__global__ void stem_bwd_fold_kernel(const float* __restrict__ partialA, double* __restrict__ sums, int Gv, int views) {
  const int lane = threadIdx.x & 31;
  const int e = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (e >= views * 128) return;
  const int v = e / 128, idx = e % 128;
  double t = 0.0;
  for (int b = lane; b < Gv; b += 32) t += (double)partialA[(long long)(v * Gv + b) * 128 + idx];
  t = warp_sum(t);
  if (lane == 0) sums[e] = t;
}

template <typename T, int CIN>
__global__ void __launch_bounds__(256) stem_bwd_wgrad_kernel(StemBwdArgs p) {
  constexpr int K = CIN * 9;
  __shared__ float red[8][16][4 * K];
  const int Gv = gridDim.x / p.views, v = blockIdx.x / Gv, lb = blockIdx.x % Gv;
  const int cg = threadIdx.x & 15, wslot = threadIdx.x >> 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long Wv = (long long)(p.n / p.views) * p.oh * p.ow;
  const T* y = (const T*)p.y;
  const T* dpool = (const T*)p.dpool;
  const double invM = 1.0 / (double)((long long)(p.n / p.views) * p.H * p.W);  // BatchNorm rows per view
  float sc[4], sh[4], cA[4], cB[4], cC[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cg * 4 + j;
    sc[j] = p.ss[v * 128 + c];
    sh[j] = p.ss[v * 128 + 64 + c];
    const float mean = p.mi[v * 128 + c], istd = p.mi[v * 128 + 64 + c];
    const float m1 = (float)(p.sums[v * 128 + c] * invM), m2 = (float)(p.sums[v * 128 + 64 + c] * invM);
    const float A = p.gamma[c] * istd;
    cA[j] = A;
    cB[j] = -A * m2 * istd;
    cC[j] = -A * m1 + A * m2 * istd * mean;
  }
  if (blockIdx.x == 0 && threadIdx.x < 64) {  // d gamma = sum g*yhat, d beta = sum g, over all views
    const int c = threadIdx.x;
    double db = 0.0, dg = 0.0;
    for (int q = 0; q < p.views; ++q) {
      db += p.sums[q * 128 + c];
      dg += p.sums[q * 128 + 64 + c];
    }
    if (p.dgamma) p.dgamma[c] = p.bn_accumulate ? p.dgamma[c] + (float)dg : (float)dg;
    if (p.dbeta) p.dbeta[c] = p.bn_accumulate ? p.dbeta[c] + (float)db : (float)db;
  }
  float acc[4][K];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < K; ++k) acc[j][k] = 0.f;
  for (long long wl = (long long)lb * 16 + wslot; wl < Wv; wl += (long long)Gv * 16) {
    const long long wi = (long long)v * Wv + wl;
    int ox, oy, img;
    if ((long long)p.views * Wv <= 0xffffffffll) {  // 32-bit divisions (the 64-bit ones cost more than the memory traffic)
      const unsigned u = (unsigned)wi, t = u / (unsigned)p.ow;
      ox = (int)(u - t * (unsigned)p.ow);
      img = (int)(t / (unsigned)p.oh);
      oy = (int)(t - (unsigned)img * (unsigned)p.oh);
    } else {
      ox = (int)(wi % p.ow);
      const long long t = wi / p.ow;
      oy = (int)(t % p.oh);
      img = (int)(t / p.oh);
    }
    float yv[4][4], g[4][4], dy[4][4];
    bool valid[4];
    window_grad<T>(y, dpool, img, oy, ox, p.H, p.W, p.oh, p.ow, p.pad, cg, sc, sh, yv, valid, g);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) dy[q][j] = valid[q] ? fmaf(cA[j], g[q][j], fmaf(cB[j], yv[q][j], cC[j])) : 0.f;
    // the window's pixels (iy0+1+qy, ix0+1+qx) share the 4 x 4 input neighbourhood starting at (iy0, ix0); its element
    // (r, c) meets pixel (qy, qx) under filter tap (a, b) = (r - qy, c - qx)
    const int iy0 = oy * 2 - p.pad - 1, ix0 = ox * 2 - p.pad - 1;
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      const float* xc = p.x + ((long long)img * CIN + ci) * p.H * p.W;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int iy = iy0 + r;
        const bool rowok = (unsigned)iy < (unsigned)p.H;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int ix = ix0 + c;
          const float xv = (rowok && (unsigned)ix < (unsigned)p.W) ? __ldg(xc + (long long)iy * p.W + ix) : 0.f;
#pragma unroll
          for (int qy = 0; qy < 2; ++qy) {
            const int a = r - qy;
            if (a < 0 || a > 2) continue;
#pragma unroll
            for (int qx = 0; qx < 2; ++qx) {
              const int b = c - qx;
              if (b < 0 || b > 2) continue;
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[j][ci * 9 + a * 3 + b] = fmaf(dy[qy * 2 + qx][j], xv, acc[j][ci * 9 + a * 3 + b]);
            }
          }
        }
      }
    }
  }
  // the two window slots of a warp (lanes l and l ^ 16), then the 8 warps, in fixed order
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < K; ++k) acc[j][k] += __shfl_xor_sync(0xffffffffu, acc[j][k], 16);
  if (lane < 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < K; ++k) red[warp][cg][j * K + k] = acc[j][k];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 64 * K; e += 256) {  // e = co * K + k  (== the OIHW index of the weight gradient)
    const int co = e / K, k = e % K;
    float t = 0.f;
#pragma unroll
    for (int wp = 0; wp < 8; ++wp) t += red[wp][co >> 2][(co & 3) * K + k];
    p.partialB[(long long)blockIdx.x * 64 * K + e] = t;
  }
}

// Pass B without the weight gradient (iic_stem_bwd_dy): the gradient of the conv output is WRITTEN (storage type) for the
// tcgen05 stem wgrad (stem_tc.cu) to read -- the routed gradient g and the BatchNorm reduce sweep over (y, g) still never
// touch memory: 7.5 GB instead of the chain's 13.7 GB at the c4 shape.  Every pixel belongs to exactly one pooling window
// (the plan requires the windows to cover the image), so dy is written exactly once.
template <typename T>
__global__ void __launch_bounds__(256) stem_bwd_dy_kernel(StemBwdArgs p) {
  const int Gv = gridDim.x / p.views, v = blockIdx.x / Gv, lb = blockIdx.x % Gv;
  const int cg = threadIdx.x & 15, wslot = threadIdx.x >> 4;
  const long long Wv = (long long)(p.n / p.views) * p.oh * p.ow;
  const T* y = (const T*)p.y;
  const T* dpool = (const T*)p.dpool;
  T* dyo = (T*)p.dy_out;
  const double invM = 1.0 / (double)((long long)(p.n / p.views) * p.H * p.W);  // BatchNorm rows per view
  float sc[4], sh[4], cA[4], cB[4], cC[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cg * 4 + j;
    sc[j] = p.ss[v * 128 + c];
    sh[j] = p.ss[v * 128 + 64 + c];
    const float mean = p.mi[v * 128 + c], istd = p.mi[v * 128 + 64 + c];
    const float m1 = (float)(p.sums[v * 128 + c] * invM), m2 = (float)(p.sums[v * 128 + 64 + c] * invM);
    const float A = p.gamma[c] * istd;
    cA[j] = A;
    cB[j] = -A * m2 * istd;
    cC[j] = -A * m1 + A * m2 * istd * mean;
  }
  if (blockIdx.x == 0 && threadIdx.x < 64) {  // d gamma = sum g*yhat, d beta = sum g, over all views
    const int c = threadIdx.x;
    double db = 0.0, dg = 0.0;
    for (int q = 0; q < p.views; ++q) {
      db += p.sums[q * 128 + c];
      dg += p.sums[q * 128 + 64 + c];
    }
    if (p.dgamma) p.dgamma[c] = p.bn_accumulate ? p.dgamma[c] + (float)dg : (float)dg;
    if (p.dbeta) p.dbeta[c] = p.bn_accumulate ? p.dbeta[c] + (float)db : (float)db;
  }
  for (long long wl = (long long)lb * 16 + wslot; wl < Wv; wl += (long long)Gv * 16) {
    const long long wi = (long long)v * Wv + wl;
    int ox, oy, img;
    if ((long long)p.views * Wv <= 0xffffffffll) {  // 32-bit divisions (the 64-bit ones cost more than the memory traffic)
      const unsigned u = (unsigned)wi, t = u / (unsigned)p.ow;
      ox = (int)(u - t * (unsigned)p.ow);
      img = (int)(t / (unsigned)p.oh);
      oy = (int)(t - (unsigned)img * (unsigned)p.oh);
    } else {
      ox = (int)(wi % p.ow);
      const long long t = wi / p.ow;
      oy = (int)(t % p.oh);
      img = (int)(t / p.oh);
    }
    float yv[4][4], g[4][4];
    bool valid[4];
    window_grad<T>(y, dpool, img, oy, ox, p.H, p.W, p.oh, p.ow, p.pad, cg, sc, sh, yv, valid, g);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (!valid[q]) continue;
      const int iy = oy * 2 - p.pad + (q >> 1), ix = ox * 2 - p.pad + (q & 1);
      float d[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) d[j] = fmaf(cA[j], g[q][j], fmaf(cB[j], yv[q][j], cC[j]));
      store4(dyo + (((long long)img * p.H + iy) * p.W + ix) * 64 + cg * 4, d);
    }
  }
}

// Second version of pass B (option stem_bwd_v2): the input neighbourhood is fetched cooperatively by the half-warp and
// broadcast with shuffles (v1 issued 16 * CIN dependent broadcast loads per thread and window and was latency bound).
template <typename T, int CIN>
__global__ void __launch_bounds__(256) stem_bwd_wgrad2_kernel(StemBwdArgs p) {
  constexpr int K = CIN * 9;
  __shared__ float red[8][16][4 * K];
  const int Gv = gridDim.x / p.views, v = blockIdx.x / Gv, lb = blockIdx.x % Gv;
  const int cg = threadIdx.x & 15, wslot = threadIdx.x >> 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long Wv = (long long)(p.n / p.views) * p.oh * p.ow;
  const T* y = (const T*)p.y;
  const T* dpool = (const T*)p.dpool;
  const double invM = 1.0 / (double)((long long)(p.n / p.views) * p.H * p.W);  // BatchNorm rows per view
  float sc[4], sh[4], cA[4], cB[4], cC[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = cg * 4 + j;
    sc[j] = p.ss[v * 128 + c];
    sh[j] = p.ss[v * 128 + 64 + c];
    const float mean = p.mi[v * 128 + c], istd = p.mi[v * 128 + 64 + c];
    const float m1 = (float)(p.sums[v * 128 + c] * invM), m2 = (float)(p.sums[v * 128 + 64 + c] * invM);
    const float A = p.gamma[c] * istd;
    cA[j] = A;
    cB[j] = -A * m2 * istd;
    cC[j] = -A * m1 + A * m2 * istd * mean;
  }
  if (blockIdx.x == 0 && threadIdx.x < 64) {  // d gamma = sum g*yhat, d beta = sum g, over all views
    const int c = threadIdx.x;
    double db = 0.0, dg = 0.0;
    for (int q = 0; q < p.views; ++q) {
      db += p.sums[q * 128 + c];
      dg += p.sums[q * 128 + 64 + c];
    }
    if (p.dgamma) p.dgamma[c] = p.bn_accumulate ? p.dgamma[c] + (float)dg : (float)dg;
    if (p.dbeta) p.dbeta[c] = p.bn_accumulate ? p.dbeta[c] + (float)db : (float)db;
  }
  float acc[4][K];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < K; ++k) acc[j][k] = 0.f;
  // Both half-warps of a warp iterate together (the shuffles below need all 32 lanes); a half-warp past the end mirrors
  // its neighbour's window and contributes zeros.
  const int hl = lane & 15;
  for (long long wl0 = (long long)lb * 16 + (wslot & ~1); wl0 < Wv; wl0 += (long long)Gv * 16) {
    const long long wl = wl0 + (wslot & 1);
    const bool active = wl < Wv;
    const long long wi = (long long)v * Wv + (active ? wl : wl0);
    const int ox = (int)(wi % p.ow);
    const long long t = wi / p.ow;
    const int oy = (int)(t % p.oh);
    const int img = (int)(t / p.oh);
    // the window's pixels (iy0+1+qy, ix0+1+qx) share the 4 x 4 input neighbourhood starting at (iy0, ix0); lane hl of
    // the half-warp fetches its element hl = r*4 + c (of each input channel): ONE load per lane instead of 16 * CIN
    // dependent broadcast loads per thread, all in flight together with the activation loads below
    const int iy0 = oy * 2 - p.pad - 1, ix0 = ox * 2 - p.pad - 1;
    float xown[CIN];
    {
      const int iy = iy0 + (hl >> 2), ix = ix0 + (hl & 3);
      const bool ok = (unsigned)iy < (unsigned)p.H && (unsigned)ix < (unsigned)p.W;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
        xown[ci] = ok ? __ldg(p.x + (((long long)img * CIN + ci) * p.H + iy) * p.W + ix) : 0.f;
    }
    float yv[4][4], g[4][4], dy[4][4];
    bool valid[4];
    window_grad<T>(y, dpool, img, oy, ox, p.H, p.W, p.oh, p.ow, p.pad, cg, sc, sh, yv, valid, g);
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        dy[q][j] = (active && valid[q]) ? fmaf(cA[j], g[q][j], fmaf(cB[j], yv[q][j], cC[j])) : 0.f;
    // element (r, c) meets pixel (qy, qx) under filter tap (a, b) = (r - qy, c - qx)
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float xv = __shfl_sync(0xffffffffu, xown[ci], (lane & 16) | (r * 4 + c));
#pragma unroll
          for (int qy = 0; qy < 2; ++qy) {
            const int a = r - qy;
            if (a < 0 || a > 2) continue;
#pragma unroll
            for (int qx = 0; qx < 2; ++qx) {
              const int b = c - qx;
              if (b < 0 || b > 2) continue;
#pragma unroll
              for (int j = 0; j < 4; ++j) acc[j][ci * 9 + a * 3 + b] = fmaf(dy[qy * 2 + qx][j], xv, acc[j][ci * 9 + a * 3 + b]);
            }
          }
        }
      }
    }
  }
  // the two window slots of a warp (lanes l and l ^ 16), then the 8 warps, in fixed order
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < K; ++k) acc[j][k] += __shfl_xor_sync(0xffffffffu, acc[j][k], 16);
  if (lane < 16) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < K; ++k) red[warp][cg][j * K + k] = acc[j][k];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 64 * K; e += 256) {  // e = co * K + k  (== the OIHW index of the weight gradient)
    const int co = e / K, k = e % K;
    float t = 0.f;
#pragma unroll
    for (int wp = 0; wp < 8; ++wp) t += red[wp][co >> 2][(co & 3) * K + k];
    p.partialB[(long long)blockIdx.x * 64 * K + e] = t;
  }
}

// grad[i] (+)= sum over blocks of partial[b][i], fixed order
__global__ void stem_bwd_dw_reduce_kernel(const float* __restrict__ partial, float* __restrict__ grad, int count, int nblk,
                                          int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float t = 0.f;
  for (int b = 0; b < nblk; ++b) t += partial[(long long)b * count + i];
  grad[i] = accumulate ? grad[i] + t : t;
}

struct StemBwdPlan {
  bool ok;
  int gridA, gridB;
  long long offA, offS, offB, bytes;
};

static StemBwdPlan stem_bwd_plan(const iic_conv_geom* g, int pool_pad, int views, int dtype) {
  StemBwdPlan pl = {};
  if (g == nullptr || (dtype != IIC_F32 && dtype != IIC_BF16)) return pl;
  if (!(g->cout == 64 && g->kh == 3 && g->kw == 3 && g->stride == 1 && g->dil == 1 && g->pad == 1 && (g->cin == 1 || g->cin == 2)))
    return pl;
  if (g->oh != g->h || g->ow != g->w) return pl;
  if (!(pool_pad == 0 || pool_pad == 1) || (g->h + 2 * pool_pad) % 2 != 0 || (g->w + 2 * pool_pad) % 2 != 0) return pl;
  if (views < 1 || views > 2 || g->n % views != 0) return pl;
  const int sms = device_sm_count();
  const long long Wv = (long long)(g->n / views) * ((g->h + 2 * pool_pad) / 2) * ((g->w + 2 * pool_pad) / 2);
  auto per_view = [&](int blocks_per_sm) {
    long long b = (long long)sms * blocks_per_sm / views;
    const long long need = (Wv + 15) / 16;
    if (b > need) b = need;
    if (b < 1) b = 1;
    return (int)b;
  };
  pl.gridA = per_view(8) * views;  // light kernel: many resident blocks cover the load latency
  pl.gridB = per_view(2) * views;  // ~170 registers per thread: one or two blocks per SM
  pl.offA = 0;
  pl.offS = ((long long)pl.gridA * 128 * 4 + 255) / 256 * 256;
  pl.offB = pl.offS + (long long)views * 128 * 8;
  pl.bytes = pl.offB + (long long)pl.gridB * 64 * g->cin * 9 * 4;
  pl.ok = true;
  return pl;
}

template <typename T>
static int stem_bwd_launch(const StemBwdPlan& pl, StemBwdArgs& a, int cin, float* dw, int w_accumulate, cudaStream_t st) {
  stem_bwd_reduce_kernel<T><<<pl.gridA, 256, 0, st>>>(a);
  IIC_LAUNCH_CHECK();
  count_launch();
  stem_bwd_fold_kernel<<<cdiv(a.views * 128, 8), 256, 0, st>>>(a.partialA, a.sums, pl.gridA / a.views, a.views);
  IIC_LAUNCH_CHECK();
  count_launch();
  const bool v2 = option(OPT_STEM_BWD_V2) != 0;
  if (cin == 1) {
    if (v2)
      stem_bwd_wgrad2_kernel<T, 1><<<pl.gridB, 256, 0, st>>>(a);
    else
      stem_bwd_wgrad_kernel<T, 1><<<pl.gridB, 256, 0, st>>>(a);
  } else {
    if (v2)
      stem_bwd_wgrad2_kernel<T, 2><<<pl.gridB, 256, 0, st>>>(a);
    else
      stem_bwd_wgrad_kernel<T, 2><<<pl.gridB, 256, 0, st>>>(a);
  }
  IIC_LAUNCH_CHECK();
  count_launch();
  const int count = 64 * cin * 9;
  stem_bwd_dw_reduce_kernel<<<cdiv(count, 128), 128, 0, st>>>(a.partialB, dw, count, pl.gridB, w_accumulate);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

}  // namespace iic

using namespace iic;

extern "C" int iic_stem_bwd_dy(const void* y, const void* dpool, const float* ss, const float* mi, const float* gamma, float* dgamma,
                               float* dbeta, int bn_accumulate, void* dy_out, const iic_conv_geom* g, int pool_pad, int views,
                               int dtype, void* workspace, long long workspace_bytes, void* stream) {
  const StemBwdPlan pl = stem_bwd_plan(g, pool_pad, views, dtype);
  IIC_REQUIRE(pl.ok, IIC_ERR_UNSUPPORTED,
              "iic_stem_bwd_dy: needs conv 3x3/s1/p1 with cin 1 or 2 -> 64, MaxPool(2,2,pad 0|1) covering every pixel, 1-2 views");
  IIC_REQUIRE(y && dpool && ss && mi && gamma && dy_out && workspace, IIC_ERR_BAD_ARG, "iic_stem_bwd_dy: null pointer");
  IIC_REQUIRE(workspace_bytes >= pl.bytes, IIC_ERR_BAD_ARG, "iic_stem_bwd_dy: workspace too small (%lld < %lld)", workspace_bytes,
              pl.bytes);
  StemBwdArgs a = {};
  a.y = y; a.dpool = dpool; a.ss = ss; a.mi = mi; a.gamma = gamma;
  a.dgamma = dgamma; a.dbeta = dbeta; a.bn_accumulate = bn_accumulate; a.dy_out = dy_out;
  a.n = g->n; a.H = g->h; a.W = g->w; a.pad = pool_pad;
  a.oh = (g->h + 2 * pool_pad - 2) / 2 + 1; a.ow = (g->w + 2 * pool_pad - 2) / 2 + 1; a.views = views;
  char* ws = (char*)workspace;
  a.partialA = (float*)(ws + pl.offA); a.sums = (double*)(ws + pl.offS);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IIC_F32) stem_bwd_reduce_kernel<float><<<pl.gridA, 256, 0, st>>>(a);
  else stem_bwd_reduce_kernel<__nv_bfloat16><<<pl.gridA, 256, 0, st>>>(a);
  IIC_LAUNCH_CHECK();
  count_launch();
  stem_bwd_fold_kernel<<<cdiv(a.views * 128, 8), 256, 0, st>>>(a.partialA, a.sums, pl.gridA / a.views, a.views);
  IIC_LAUNCH_CHECK();
  count_launch();
  if (dtype == IIC_F32) stem_bwd_dy_kernel<float><<<pl.gridA, 256, 0, st>>>(a);
  else stem_bwd_dy_kernel<__nv_bfloat16><<<pl.gridA, 256, 0, st>>>(a);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" long long iic_stem_bwd_fused_workspace(const iic_conv_geom* g, int pool_pad, int views, int dtype) {
  const StemBwdPlan pl = stem_bwd_plan(g, pool_pad, views, dtype);
  return pl.ok ? pl.bytes : 0;
}

extern "C" int iic_stem_bwd_fused(const float* x_nchw, const void* y, const void* dpool, const float* ss, const float* mi,
                                  const float* gamma, float* dgamma, float* dbeta, int bn_accumulate, float* dw_oihw,
                                  int w_accumulate, const iic_conv_geom* g, int pool_pad, int views, int dtype,
                                  void* workspace, long long workspace_bytes, void* stream) {
  const StemBwdPlan pl = stem_bwd_plan(g, pool_pad, views, dtype);
  IIC_REQUIRE(pl.ok, IIC_ERR_UNSUPPORTED,
              "iic_stem_bwd_fused: needs conv 3x3/s1/p1 with cin 1 or 2 -> 64, MaxPool(2,2,pad 0|1) covering every pixel, 1-2 views");
  IIC_REQUIRE(x_nchw && y && dpool && ss && mi && gamma && dw_oihw && workspace, IIC_ERR_BAD_ARG, "iic_stem_bwd_fused: null pointer");
  IIC_REQUIRE(workspace_bytes >= pl.bytes, IIC_ERR_BAD_ARG, "iic_stem_bwd_fused: workspace too small (%lld < %lld)",
              workspace_bytes, pl.bytes);
  StemBwdArgs a = {};
  a.x = x_nchw; a.y = y; a.dpool = dpool; a.ss = ss; a.mi = mi; a.gamma = gamma;
  a.dgamma = dgamma; a.dbeta = dbeta; a.bn_accumulate = bn_accumulate;
  a.n = g->n; a.H = g->h; a.W = g->w; a.pad = pool_pad;
  a.oh = (g->h + 2 * pool_pad - 2) / 2 + 1; a.ow = (g->w + 2 * pool_pad - 2) / 2 + 1; a.views = views;
  char* ws = (char*)workspace;
  a.partialA = (float*)(ws + pl.offA); a.sums = (double*)(ws + pl.offS); a.partialB = (float*)(ws + pl.offB);
  if (dtype == IIC_F32) return stem_bwd_launch<float>(pl, a, g->cin, dw_oihw, w_accumulate, (cudaStream_t)stream);
  return stem_bwd_launch<__nv_bfloat16>(pl, a, g->cin, dw_oihw, w_accumulate, (cudaStream_t)stream);
}
