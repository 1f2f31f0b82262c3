// Segmentation IIC objective (per-pixel, with a (2T+1)^2 displacement window):
//   xu-ji/IIC code/utils/segmentation/IID_losses.py:14-83   (IID_segmentation_loss, "collapsed")
//   xu-ji/IIC code/utils/segmentation/IID_losses.py:86-159  (IID_segmentation_loss_uncollapsed)
//   xu-ji/IIC code/utils/segmentation/transforms.py:131-143 (perform_affine_tf)
//
// The reference evaluates the joint with F.conv2d(x1^T, weight=x2^T, padding=T): a convolution whose
// "filter" is the whole second view (2 k^2 (2T+1)^2 n h w FLOP), then ~40 tiny kernels.  Here:
//   iic_seg_prepare      x1m = x1*mask ; x2m = bilinear_sample(x2, theta)*mask   (NCHW -> pixel-major [n][h][w][KP])
//   iic_seg_joint        A[u][v][c][c'] = sum_{n,y,x} x1m[n,y+u-T,x+v-T,c] * x2m[n,y,x,c']
//                        register-tiled SIMT correlation: a CTA owns one displacement row u, a thread owns
//                        (c, 3 values of v) x all c'; rows of both views stream through shared memory once.
//   iic_joint_mi         (iid_loss.cu) MI + analytic gradient H_t per displacement, S = (2T+1)^2 "sub-heads"
//   iic_seg_corr_bwd     d x1m = sum_t H_t (x) x2m(shifted),  d x2m = sum_t H_t (x) x1m(shifted)
//   iic_seg_unprepare    mask and (bilinear adjoint) scatter the gradients back to NCHW
//   iic_box_filter       collapsed variant: sum_t A_t = sum_p Box(x1m)[p] x2m[p]^T  (SURVEY.md S8 a10), then the
//                        same joint / corr kernels with T = 0.
// fp32 throughout (P spans 16 orders of magnitude; this is the MI path).  The correlation kernels are
// FMA-bound SIMT; a tcgen05 tf32 formulation is the planned next step for k >= 15.
#include "common.cuh"

namespace iic {

// ---- prepare: mask, affine resample (align_corners=True, zeros padding), NCHW -> [n][h][w][KP] ----------------
__device__ __forceinline__ void affine_src(const float* th, int x, int y, int w, int h, float& ix, float& iy) {
  const float xn = w > 1 ? -1.f + 2.f * (float)x / (float)(w - 1) : 0.f;
  const float yn = h > 1 ? -1.f + 2.f * (float)y / (float)(h - 1) : 0.f;
  const float xs = th[0] * xn + th[1] * yn + th[2];
  const float ys = th[3] * xn + th[4] * yn + th[5];
  ix = (xs + 1.f) * 0.5f * (float)(w - 1);
  iy = (ys + 1.f) * 0.5f * (float)(h - 1);
}

__global__ void seg_prepare_kernel(const float* __restrict__ x1, const float* __restrict__ x2,
                                   const float* __restrict__ theta, const float* __restrict__ mask,
                                   float* __restrict__ x1m, float* __restrict__ x2m, int n, int k, int h, int w, int KP,
                                   int tx, int ty) {
  // (tx, ty): the sparse random displacement of the reference (random_translation_multiple, seg transforms.py:146-166):
  // the resampled x2 is read at (x + tx, y + ty) and is zero where that falls outside the frame
  const long long total = (long long)n * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int ni = (int)(i / ((long long)w * h));
    const int xs = x + tx, ys = y + ty;
    const float m = (xs >= 0 && xs < w && ys >= 0 && ys < h) ? mask[i] : 0.f;
    const float m1 = mask[i];
    float ix, iy;
    affine_src(theta + ni * 6, xs, ys, w, h, ix, iy);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float lx = ix - fx, ly = iy - fy;
    const float w00 = (1.f - lx) * (1.f - ly), w01 = lx * (1.f - ly), w10 = (1.f - lx) * ly, w11 = lx * ly;
    const bool vx0 = x0 >= 0 && x0 < w, vx1 = x0 + 1 >= 0 && x0 + 1 < w, vy0 = y0 >= 0 && y0 < h, vy1 = y0 + 1 >= 0 && y0 + 1 < h;
    for (int c = 0; c < KP; ++c) {
      float a = 0.f, b = 0.f;
      if (c < k) {
        a = x1[((long long)ni * k + c) * h * w + (long long)y * w + x] * m1;
        const float* pl = x2 + ((long long)ni * k + c) * h * w;
        float v = 0.f;
        if (vy0 && vx0) v += w00 * pl[(long long)y0 * w + x0];
        if (vy0 && vx1) v += w01 * pl[(long long)y0 * w + x0 + 1];
        if (vy1 && vx0) v += w10 * pl[(long long)(y0 + 1) * w + x0];
        if (vy1 && vx1) v += w11 * pl[(long long)(y0 + 1) * w + x0 + 1];
        b = v * m;
      }
      x1m[i * KP + c] = a;
      x2m[i * KP + c] = b;
    }
  }
}

// adjoint: dx1 = d x1m * mask (gather) ; dx2 += bilinear^T (d x2m * mask) (atomics; dx2 pre-zeroed)
__global__ void seg_unprepare_kernel(const float* __restrict__ dx1m, const float* __restrict__ dx2m,
                                     const float* __restrict__ theta, const float* __restrict__ mask,
                                     float* __restrict__ dx1, float* __restrict__ dx2, int n, int k, int h, int w, int KP,
                                     int tx, int ty) {
  const long long total = (long long)n * h * w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int ni = (int)(i / ((long long)w * h));
    const int xs = x + tx, ys = y + ty;
    const float m = (xs >= 0 && xs < w && ys >= 0 && ys < h) ? mask[i] : 0.f;
    const float m1 = mask[i];
    float ix, iy;
    affine_src(theta + ni * 6, xs, ys, w, h, ix, iy);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float lx = ix - fx, ly = iy - fy;
    const float w00 = (1.f - lx) * (1.f - ly), w01 = lx * (1.f - ly), w10 = (1.f - lx) * ly, w11 = lx * ly;
    const bool vx0 = x0 >= 0 && x0 < w, vx1 = x0 + 1 >= 0 && x0 + 1 < w, vy0 = y0 >= 0 && y0 < h, vy1 = y0 + 1 >= 0 && y0 + 1 < h;
    for (int c = 0; c < k; ++c) {
      dx1[((long long)ni * k + c) * h * w + (long long)y * w + x] = dx1m[i * KP + c] * m1;
      const float g = dx2m[i * KP + c] * m;
      float* pl = dx2 + ((long long)ni * k + c) * h * w;
      if (g != 0.f) {
        if (vy0 && vx0 && w00 != 0.f) atomicAdd(pl + (long long)y0 * w + x0, w00 * g);
        if (vy0 && vx1 && w01 != 0.f) atomicAdd(pl + (long long)y0 * w + x0 + 1, w01 * g);
        if (vy1 && vx0 && w10 != 0.f) atomicAdd(pl + (long long)(y0 + 1) * w + x0, w10 * g);
        if (vy1 && vx1 && w11 != 0.f) atomicAdd(pl + (long long)(y0 + 1) * w + x0 + 1, w11 * g);
      }
    }
  }
}

// ---- joint over displacements ---------------------------------------------------------------------------
// grid = (2T+1, G): CTA (u, g) handles displacement row u for images g, g+G, ...; block = KP * NVG threads,
// thread (c, vg) accumulates A[u][3vg..3vg+2][c][0..KP) in registers.  part[g][u][v][c][c'].
template <int KP>
__global__ void seg_joint_kernel(const float* __restrict__ x1m, const float* __restrict__ x2m, float* __restrict__ part,
                                 int n, int h, int w, int T) {
  extern __shared__ __align__(16) float sm[];
  const int V = 2 * T + 1;
  const int u = blockIdx.x, g = blockIdx.y, G = gridDim.y;
  float* x2row = sm;                    // [w][KP]
  float* x1row = sm + (size_t)w * KP;   // [w + 2T][KP], x1row[(x + T)] = x1m[.., x]
  const int c = threadIdx.x % KP, vg = threadIdx.x / KP;
  const int v0 = vg * 3;
  float acc[3][KP];
#pragma unroll
  for (int j = 0; j < 3; ++j)
#pragma unroll
    for (int e = 0; e < KP; ++e) acc[j][e] = 0.f;
  for (int ni = g; ni < n; ni += G) {
    for (int y = 0; y < h; ++y) {
      const int y1 = y + u - T;
      if (y1 < 0 || y1 >= h) continue;  // (uniform across the CTA)
      __syncthreads();
      const float* s2 = x2m + ((long long)ni * h + y) * w * KP;
      const float* s1 = x1m + ((long long)ni * h + y1) * w * KP;
      for (int i = threadIdx.x; i < w * KP / 4; i += blockDim.x)
        reinterpret_cast<float4*>(x2row)[i] = reinterpret_cast<const float4*>(s2)[i];
      for (int i = threadIdx.x; i < (w + 2 * T) * KP / 4; i += blockDim.x) {
        const int px = (i * 4) / KP - T;
        reinterpret_cast<float4*>(x1row)[i] = (px >= 0 && px < w)
                                                  ? reinterpret_cast<const float4*>(s1)[i - T * KP / 4]
                                                  : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();
      for (int x = 0; x < w; ++x) {
        float b[KP];
#pragma unroll
        for (int e = 0; e < KP; e += 4) {
          const float4 t = *reinterpret_cast<const float4*>(x2row + x * KP + e);
          b[e] = t.x; b[e + 1] = t.y; b[e + 2] = t.z; b[e + 3] = t.w;
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const int v = v0 + j;
          const float a = (v < V) ? x1row[(x + v) * KP + c] : 0.f;
#pragma unroll
          for (int e = 0; e < KP; ++e) acc[j][e] = fmaf(a, b[e], acc[j][e]);
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int v = v0 + j;
    if (v < V) {
      float* o = part + ((((long long)g * V + u) * V + v) * KP + c) * KP;
#pragma unroll
      for (int e = 0; e < KP; ++e) o[e] = acc[j][e];
    }
  }
}

// A[t][c][c'] (k x k, unpadded) = sum_g part[g][t][c][c']   (fixed order)
__global__ void seg_joint_reduce_kernel(const float* __restrict__ part, float* __restrict__ A, int G, int VV, int k, int KP) {
  const long long total = (long long)VV * k * k;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cp = (int)(i % k);
    const int c = (int)((i / k) % k);
    const int t = (int)(i / ((long long)k * k));
    float s = 0.f;
    for (int g = 0; g < G; ++g) s += part[(((long long)g * VV + t) * KP + c) * KP + cp];
    A[i] = s;
  }
}

// ---- backward correlation ----------------------------------------------------------------------------
// out[n,Y,X,c] = scale * sum_{u,v,c'} H[u][v][c][c'] * in[n, Y - sgn*(u-T), X - sgn*(v-T), c']
// CTA = (n, Y, 64-pixel X tile); 128 threads = 4 channel groups (KP/4 channels each) x 32 lanes, 2 pixels per thread.
template <int KP>
__global__ void __launch_bounds__(128) seg_corr_bwd_kernel(const float* __restrict__ in, const float* __restrict__ H,
                                                           float* __restrict__ out, int n, int h, int w, int T, int k,
                                                           int sgn, float scale) {
  constexpr int CPT = KP / 4;  // channels per thread
  extern __shared__ __align__(16) float sm[];
  const int V = 2 * T + 1;
  const int xt = blockIdx.x * 64, Y = blockIdx.y, ni = blockIdx.z;
  float* hs = sm;                           // [V][KP(c')][KP(c)]  (c fastest)
  float* row = hs + (size_t)V * KP * KP;    // [KP(c')][64 + 2T]
  const int RW = 64 + 2 * T;
  const int lane = threadIdx.x & 31, cgp = threadIdx.x >> 5;
  float acc[2][CPT];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int e = 0; e < CPT; ++e) acc[p][e] = 0.f;
  for (int u = 0; u < V; ++u) {
    const int yr = Y - sgn * (u - T);
    if (yr < 0 || yr >= h) continue;
    __syncthreads();
    // H[u][v][c][c'] (unpadded k x k, symmetric in (c,c')) -> hs[v][c'][c], zero padded
    for (int i = threadIdx.x; i < V * KP * KP; i += 128) {
      const int cc = i % KP, cp = (i / KP) % KP, v = i / (KP * KP);
      hs[i] = (cc < k && cp < k) ? H[(((long long)u * V + v) * k + cc) * k + cp] : 0.f;
    }
    // input row segment, transposed to [c'][x]: x index j <-> pixel xt - T + j
    const float* src = in + ((long long)ni * h + yr) * w * KP;
    for (int i = threadIdx.x; i < RW * KP; i += 128) {
      const int cp = i % KP, j = i / KP;
      const int px = xt - T + j;
      row[cp * RW + j] = (px >= 0 && px < w) ? src[(long long)px * KP + cp] : 0.f;
    }
    __syncthreads();
    for (int v = 0; v < V; ++v) {
      const int off = T - sgn * (v - T);  // row index of pixel (X - sgn*(v-T)) for X = xt + lane (+32)
#pragma unroll 4
      for (int cp = 0; cp < KP; ++cp) {
        const float i0 = row[cp * RW + lane + off], i1 = row[cp * RW + lane + 32 + off];
        const float* hp = hs + ((size_t)v * KP + cp) * KP + cgp * CPT;
#pragma unroll
        for (int e = 0; e < CPT; ++e) {
          acc[0][e] = fmaf(hp[e], i0, acc[0][e]);
          acc[1][e] = fmaf(hp[e], i1, acc[1][e]);
        }
      }
    }
  }
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int X = xt + lane + 32 * p;
    if (X < w) {
      float* o = out + (((long long)ni * h + Y) * w + X) * KP + cgp * CPT;
#pragma unroll
      for (int e = 0; e < CPT; ++e) o[e] = acc[p][e] * scale;
    }
  }
}

// zero-padded (2T+1)^2 box sum on [n][h][w][KP]; separable, direct
__global__ void box_filter_kernel(const float* __restrict__ in, float* __restrict__ out, int n, int h, int w, int KP, int T,
                                  int horizontal) {
  const long long total = (long long)n * h * w * KP;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % KP);
    long long p = i / KP;
    const int x = (int)(p % w);
    p /= w;
    const int y = (int)(p % h);
    const int ni = (int)(p / h);
    float s = 0.f;
    if (horizontal) {
      for (int d = -T; d <= T; ++d) {
        const int xx = x + d;
        if (xx >= 0 && xx < w) s += in[(((long long)ni * h + y) * w + xx) * KP + c];
      }
    } else {
      for (int d = -T; d <= T; ++d) {
        const int yy = y + d;
        if (yy >= 0 && yy < h) s += in[(((long long)ni * h + yy) * w + x) * KP + c];
      }
    }
    out[i] = s;
  }
}

static int seg_kp(int k) {
  if (k <= 4) return 4;
  if (k <= 8) return 8;
  if (k <= 16) return 16;
  if (k <= 32) return 32;
  if (k <= 48) return 48;
  return 0;
}

}  // namespace iic

using namespace iic;

extern "C" int iic_seg_kp(int k) { return seg_kp(k); }

extern "C" int iic_seg_prepare(const float* x1, const float* x2, const float* theta, const float* mask, float* x1m,
                               float* x2m, int n, int k, int h, int w, void* stream) {
  return iic_seg_prepare_shift(x1, x2, theta, mask, x1m, x2m, n, k, h, w, 0, 0, stream);
}

extern "C" int iic_seg_prepare_shift(const float* x1, const float* x2, const float* theta, const float* mask, float* x1m,
                                     float* x2m, int n, int k, int h, int w, int tx, int ty, void* stream) {
  IIC_REQUIRE(x1 && x2 && theta && mask && x1m && x2m && n > 0 && k > 0 && h > 0 && w > 0, IIC_ERR_BAD_ARG,
              "iic_seg_prepare: bad arguments");
  const int KP = seg_kp(k);
  IIC_REQUIRE(KP != 0, IIC_ERR_UNSUPPORTED, "segmentation losses support k <= 48 (got %d)", k);
  const long long total = (long long)n * h * w;
  int blocks = cdiv(total, 256);
  if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
  seg_prepare_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x1, x2, theta, mask, x1m, x2m, n, k, h, w, KP, tx, ty);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_seg_unprepare(const float* dx1m, const float* dx2m, const float* theta, const float* mask, float* dx1,
                                 float* dx2, int n, int k, int h, int w, void* stream) {
  return iic_seg_unprepare_shift(dx1m, dx2m, theta, mask, dx1, dx2, n, k, h, w, 0, 0, stream);
}

extern "C" int iic_seg_unprepare_shift(const float* dx1m, const float* dx2m, const float* theta, const float* mask,
                                       float* dx1, float* dx2, int n, int k, int h, int w, int tx, int ty, void* stream) {
  IIC_REQUIRE(dx1m && dx2m && theta && mask && dx1 && dx2 && n > 0 && k > 0, IIC_ERR_BAD_ARG,
              "iic_seg_unprepare: bad arguments");
  const int KP = seg_kp(k);
  IIC_REQUIRE(KP != 0, IIC_ERR_UNSUPPORTED, "segmentation losses support k <= 48 (got %d)", k);
  cudaStream_t st = (cudaStream_t)stream;
  IIC_CUDA(cudaMemsetAsync(dx2, 0, sizeof(float) * (size_t)n * k * h * w, st));
  const long long total = (long long)n * h * w;
  int blocks = cdiv(total, 256);
  if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
  seg_unprepare_kernel<<<blocks, 256, 0, st>>>(dx1m, dx2m, theta, mask, dx1, dx2, n, k, h, w, KP, tx, ty);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" long long iic_seg_joint_workspace(int n, int k, int T) {
  const int KP = seg_kp(k);
  if (KP == 0 || n <= 0 || T < 0) return -1;
  const int V = 2 * T + 1;
  int G = (device_sm_count() * 4 + V - 1) / V;
  if (G > n) G = n;
  if (G < 1) G = 1;
  return (long long)G * V * V * KP * KP * (long long)sizeof(float);
}

template <int KP>
static int launch_seg_joint(const float* x1m, const float* x2m, float* part, int n, int h, int w, int T, int G, cudaStream_t st) {
  const int V = 2 * T + 1, nvg = (V + 2) / 3;
  const size_t smem = (size_t)(w + w + 2 * T) * KP * sizeof(float);
  IIC_CUDA(cudaFuncSetAttribute(seg_joint_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  seg_joint_kernel<KP><<<dim3(V, G), KP * nvg, smem, st>>>(x1m, x2m, part, n, h, w, T);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_seg_joint(const float* x1m, const float* x2m, float* joint, void* workspace, int n, int k, int h,
                             int w, int T, void* stream) {
  IIC_REQUIRE(x1m && x2m && joint && workspace && n > 0 && k > 0 && T >= 0, IIC_ERR_BAD_ARG, "iic_seg_joint: bad arguments");
  const int KP = seg_kp(k);
  IIC_REQUIRE(KP != 0, IIC_ERR_UNSUPPORTED, "segmentation losses support k <= 48 (got %d)", k);
  IIC_REQUIRE((size_t)(2 * w + 2 * T) * KP * 4 <= 200 * 1024, IIC_ERR_UNSUPPORTED, "iic_seg_joint: row too wide (w=%d)", w);
  const int V = 2 * T + 1;
  int G = (device_sm_count() * 4 + V - 1) / V;
  if (G > n) G = n;
  if (G < 1) G = 1;
  cudaStream_t st = (cudaStream_t)stream;
  float* part = (float*)workspace;
  int rc;
  switch (KP) {
    case 4: rc = launch_seg_joint<4>(x1m, x2m, part, n, h, w, T, G, st); break;
    case 8: rc = launch_seg_joint<8>(x1m, x2m, part, n, h, w, T, G, st); break;
    case 16: rc = launch_seg_joint<16>(x1m, x2m, part, n, h, w, T, G, st); break;
    case 32: rc = launch_seg_joint<32>(x1m, x2m, part, n, h, w, T, G, st); break;
    default: rc = launch_seg_joint<48>(x1m, x2m, part, n, h, w, T, G, st); break;
  }
  if (rc != IIC_OK) return rc;
  const long long total = (long long)V * V * k * k;
  seg_joint_reduce_kernel<<<cdiv(total, 256), 256, 0, st>>>(part, joint, G, V * V, k, KP);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

template <int KP>
static int launch_corr(const float* in, const float* H, float* out, int n, int h, int w, int T, int k, int sgn, float scale,
                       cudaStream_t st) {
  const int V = 2 * T + 1;
  const size_t smem = ((size_t)V * KP * KP + (size_t)KP * (64 + 2 * T)) * sizeof(float);
  IIC_REQUIRE(smem <= 200 * 1024, IIC_ERR_UNSUPPORTED, "iic_seg_corr_bwd: T=%d, k=%d need %zu B of shared memory", T, k, smem);
  IIC_CUDA(cudaFuncSetAttribute(seg_corr_bwd_kernel<KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  seg_corr_bwd_kernel<KP><<<dim3(cdiv(w, 64), h, n), 128, smem, st>>>(in, H, out, n, h, w, T, k, sgn, scale);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_seg_corr_bwd(const float* in, const float* H, float* out, int n, int k, int h, int w, int T, int sgn,
                                float scale, void* stream) {
  IIC_REQUIRE(in && H && out && n > 0 && k > 0 && T >= 0 && (sgn == 1 || sgn == -1), IIC_ERR_BAD_ARG,
              "iic_seg_corr_bwd: bad arguments");
  const int KP = seg_kp(k);
  IIC_REQUIRE(KP != 0, IIC_ERR_UNSUPPORTED, "segmentation losses support k <= 48 (got %d)", k);
  cudaStream_t st = (cudaStream_t)stream;
  switch (KP) {
    case 4: return launch_corr<4>(in, H, out, n, h, w, T, k, sgn, scale, st);
    case 8: return launch_corr<8>(in, H, out, n, h, w, T, k, sgn, scale, st);
    case 16: return launch_corr<16>(in, H, out, n, h, w, T, k, sgn, scale, st);
    case 32: return launch_corr<32>(in, H, out, n, h, w, T, k, sgn, scale, st);
    default: return launch_corr<48>(in, H, out, n, h, w, T, k, sgn, scale, st);
  }
}

extern "C" int iic_box_filter(const float* in, float* tmp, float* out, int n, int k, int h, int w, int T, void* stream) {
  IIC_REQUIRE(in && tmp && out && n > 0 && k > 0 && T >= 0, IIC_ERR_BAD_ARG, "iic_box_filter: bad arguments");
  const int KP = seg_kp(k);
  IIC_REQUIRE(KP != 0, IIC_ERR_UNSUPPORTED, "segmentation losses support k <= 48 (got %d)", k);
  const long long total = (long long)n * h * w * KP;
  int blocks = cdiv(total, 256);
  if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
  cudaStream_t st = (cudaStream_t)stream;
  box_filter_kernel<<<blocks, 256, 0, st>>>(in, tmp, n, h, w, KP, T, 1);
  IIC_LAUNCH_CHECK();
  count_launch();
  box_filter_kernel<<<blocks, 256, 0, st>>>(tmp, out, n, h, w, KP, T, 0);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}
