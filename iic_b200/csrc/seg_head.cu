// Segmentation sub-heads of SegmentationNet10a (xu-ji/IIC code/archs/segmentation/net10a.py:34-59):
//   S x [ Conv2d(512 -> k, 1x1, **padding=1**, bias=False) -> Softmax2d ] -> F.interpolate(size=input_sz, bilinear)
// The 1x1 conv with padding 1 grows the (hf x wf) feature map to (hf+2) x (wf+2); border pixels see only
// zero padding, so their logits are 0 and their softmax is the uniform 1/k.
//   forward : logits = feat [M][512] x W^T (fp32 SIMT GEMM) -> per-pixel softmax with the uniform border
//             -> bilinear upsample (align_corners=False) straight into the reference's NCHW (n,k,H,W) output
//   backward: gather-form adjoint of the upsample -> softmax Jacobian (border dropped: no parameters
//             behind it) -> dW (split-K GEMM over pixels) and d feat (GEMM), feat in fp32 or bf16.
#include "simt_gemm.cuh"

namespace iic {

// zlow[n][hl][wl][k]: interior = softmax(logits[n][hl-2][wl-2][k]); border = 1/k
__global__ void seg_softmax_kernel(const float* __restrict__ logits, float* __restrict__ zlow, int n, int hl, int wl, int k) {
  const long long total = (long long)n * hl * wl;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % wl), y = (int)((i / wl) % hl), ni = (int)(i / ((long long)wl * hl));
    float* o = zlow + i * k;
    if (x == 0 || y == 0 || x == wl - 1 || y == hl - 1) {
      const float u = 1.f / (float)k;
      for (int c = 0; c < k; ++c) o[c] = u;
    } else {
      const float* l = logits + (((long long)ni * (hl - 2) + (y - 1)) * (wl - 2) + (x - 1)) * k;
      float m = -INFINITY;
      for (int c = 0; c < k; ++c) m = fmaxf(m, l[c]);
      float s = 0.f;
      for (int c = 0; c < k; ++c) s += expf(l[c] - m);
      for (int c = 0; c < k; ++c) o[c] = expf(l[c] - m) / s;
    }
  }
}

__device__ __forceinline__ void bilin_src(int dst, float scale, int in, int& i0, int& i1, float& l0, float& l1) {
  float src = ((float)dst + 0.5f) * scale - 0.5f;  // align_corners=False (F.interpolate default, net10a.py:56)
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in - 1) i0 = in - 1;
  i1 = i0 + (i0 < in - 1 ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}

// out NCHW (n,k,H,W) <- zlow [n][hl][wl][k]
__global__ void upsample_fwd_kernel(const float* __restrict__ zlow, float* __restrict__ out, int n, int hl, int wl, int k,
                                    int H, int W) {
  const float sy = (float)hl / (float)H, sx = (float)wl / (float)W;
  const long long total = (long long)n * H * W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W), y = (int)((i / W) % H), ni = (int)(i / ((long long)W * H));
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    bilin_src(y, sy, hl, y0, y1, ly0, ly1);
    bilin_src(x, sx, wl, x0, x1, lx0, lx1);
    const float* p00 = zlow + (((long long)ni * hl + y0) * wl + x0) * k;
    const float* p01 = zlow + (((long long)ni * hl + y0) * wl + x1) * k;
    const float* p10 = zlow + (((long long)ni * hl + y1) * wl + x0) * k;
    const float* p11 = zlow + (((long long)ni * hl + y1) * wl + x1) * k;
    for (int c = 0; c < k; ++c)
      out[(((long long)ni * k + c) * H + y) * W + x] = ly0 * (lx0 * p00[c] + lx1 * p01[c]) + ly1 * (lx0 * p10[c] + lx1 * p11[c]);
  }
}

// dzlow [n][hl][wl][k] <- dout NCHW: every low-res pixel gathers from the high-res pixels that read it
__global__ void upsample_bwd_kernel(const float* __restrict__ dout, float* __restrict__ dzlow, int n, int hl, int wl, int k,
                                    int H, int W) {
  const float sy = (float)hl / (float)H, sx = (float)wl / (float)W;
  const long long total = (long long)n * hl * wl * k;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % k);
    long long p = i / k;
    const int lx = (int)(p % wl);
    p /= wl;
    const int ly = (int)(p % hl);
    const int ni = (int)(p / hl);
    int oy_lo = (int)floorf(((float)ly - 0.5f) / sy - 0.5f) - 1, oy_hi = (int)ceilf(((float)ly + 1.5f) / sy - 0.5f) + 1;
    int ox_lo = (int)floorf(((float)lx - 0.5f) / sx - 0.5f) - 1, ox_hi = (int)ceilf(((float)lx + 1.5f) / sx - 0.5f) + 1;
    if (oy_lo < 0) oy_lo = 0;
    if (ox_lo < 0) ox_lo = 0;
    if (oy_hi > H - 1) oy_hi = H - 1;
    if (ox_hi > W - 1) ox_hi = W - 1;
    const float* plane = dout + ((long long)ni * k + c) * H * W;
    float acc = 0.f;
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      int y0, y1;
      float l0, l1;
      bilin_src(oy, sy, hl, y0, y1, l0, l1);
      const float wy = (y0 == ly ? l0 : 0.f) + (y1 == ly ? l1 : 0.f);
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        int x0, x1;
        float m0, m1;
        bilin_src(ox, sx, wl, x0, x1, m0, m1);
        const float wx = (x0 == lx ? m0 : 0.f) + (x1 == lx ? m1 : 0.f);
        if (wx != 0.f) acc = fmaf(wy * wx, plane[(long long)oy * W + ox], acc);
      }
    }
    dzlow[i] = acc;
  }
}

// dlogits[n][hl-2][wl-2][k] = z * (dz - sum_c z dz) for interior pixels
__global__ void seg_softmax_bwd_kernel(const float* __restrict__ zlow, const float* __restrict__ dzlow,
                                       float* __restrict__ dlogits, int n, int hl, int wl, int k) {
  const int hi = hl - 2, wi = wl - 2;
  const long long total = (long long)n * hi * wi;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % wi), y = (int)((i / wi) % hi), ni = (int)(i / ((long long)wi * hi));
    const long long q = (((long long)ni * hl + y + 1) * wl + x + 1) * k;
    float dot = 0.f;
    for (int c = 0; c < k; ++c) dot = fmaf(zlow[q + c], dzlow[q + c], dot);
    for (int c = 0; c < k; ++c) dlogits[i * k + c] = zlow[q + c] * (dzlow[q + c] - dot);
  }
}

template <typename T>
static int seg_head_fwd_t(const T* feat, const float* w, float* logits_ws, float* zlow, float* out, int n, int hf, int wf,
                          int C, int k, int H, int W, cudaStream_t st) {
  const int M = n * hf * wf;
  DenseLoad<T> A{feat, (long long)C, 1, M, C};
  DenseLoad<float> B{w, (long long)C, 1, k, C};
  StoreOut<float> S{logits_ws, nullptr, (long long)k};
  int rc = launch_simt<DenseLoad<T>, DenseLoad<float>, StoreOut<float>, true, true>(A, B, S, M, k, C, 1, st);
  if (rc != IIC_OK) return rc;
  const int hl = hf + 2, wl = wf + 2;
  int blocks = cdiv((long long)n * hl * wl, 256);
  seg_softmax_kernel<<<blocks, 256, 0, st>>>(logits_ws, zlow, n, hl, wl, k);
  IIC_LAUNCH_CHECK();
  count_launch();
  blocks = cdiv((long long)n * H * W, 256);
  if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
  upsample_fwd_kernel<<<blocks, 256, 0, st>>>(zlow, out, n, hl, wl, k, H, W);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

template <typename T>
static int seg_head_bwd_t(const T* feat, const float* w, const float* zlow, const float* dout, float* dzlow_ws,
                          float* dlogits_ws, float* dw, float* dw_ws, T* dfeat, int accumulate_dfeat, int n, int hf, int wf,
                          int C, int k, int H, int W, int splits, cudaStream_t st) {
  const int hl = hf + 2, wl = wf + 2, M = n * hf * wf;
  int blocks = cdiv((long long)n * hl * wl * k, 256);
  if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
  upsample_bwd_kernel<<<blocks, 256, 0, st>>>(dout, dzlow_ws, n, hl, wl, k, H, W);
  IIC_LAUNCH_CHECK();
  count_launch();
  seg_softmax_bwd_kernel<<<cdiv((long long)M, 256), 256, 0, st>>>(zlow, dzlow_ws, dlogits_ws, n, hl, wl, k);
  IIC_LAUNCH_CHECK();
  count_launch();
  {  // dw[k][C] = dlogits^T [k][M] x feat [M][C]   (split-K over pixels)
    DenseLoad<float> A{dlogits_ws, 1, (long long)k, k, M};
    DenseLoad<T> B{feat, 1, (long long)C, C, M};
    StorePartial S{dw_ws};
    int rc = launch_simt<DenseLoad<float>, DenseLoad<T>, StorePartial, false, false>(A, B, S, k, C, M, splits, st);
    if (rc != IIC_OK) return rc;
    splitk_reduce_kernel<<<cdiv((long long)k * C, 256), 256, 0, st>>>(dw_ws, dw, (long long)k * C, splits, 0);
    IIC_LAUNCH_CHECK();
    count_launch();
  }
  if (dfeat != nullptr) {  // dfeat[M][C] (+)= dlogits [M][k] x w [k][C]
    DenseLoad<float> A{dlogits_ws, (long long)k, 1, M, k};
    DenseLoad<float> B{w, 1, (long long)C, C, k};
    StoreOut<T> S{dfeat, accumulate_dfeat ? dfeat : nullptr, (long long)C};
    int rc = launch_simt<DenseLoad<float>, DenseLoad<float>, StoreOut<T>, true, false>(A, B, S, M, C, k, 1, st);
    if (rc != IIC_OK) return rc;
  }
  return IIC_OK;
}

static int seg_head_splits(int n, int hf, int wf) {
  long long M = (long long)n * hf * wf;
  long long s = M / 2048;
  if (s < 1) s = 1;
  if (s > 256) s = 256;
  return (int)s;
}

}  // namespace iic

using namespace iic;

extern "C" long long iic_seg_head_workspace(int n, int hf, int wf, int C, int k) {
  return (long long)seg_head_splits(n, hf, wf) * k * C * (long long)sizeof(float);
}

extern "C" int iic_seg_head_fwd(const void* feat, int dtype, const float* w, float* logits_ws, float* zlow, float* out,
                                int n, int hf, int wf, int C, int k, int H, int W, void* stream) {
  IIC_REQUIRE(feat && w && logits_ws && zlow && out && n > 0 && hf > 0 && wf > 0 && C > 0 && k > 0 && H > 0 && W > 0,
              IIC_ERR_BAD_ARG, "iic_seg_head_fwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IIC_F32) return seg_head_fwd_t<float>((const float*)feat, w, logits_ws, zlow, out, n, hf, wf, C, k, H, W, st);
  if (dtype == IIC_BF16)
    return seg_head_fwd_t<__nv_bfloat16>((const __nv_bfloat16*)feat, w, logits_ws, zlow, out, n, hf, wf, C, k, H, W, st);
  set_error("iic_seg_head_fwd: bad dtype");
  return IIC_ERR_BAD_ARG;
}

extern "C" int iic_seg_head_bwd(const void* feat, int dtype, const float* w, const float* zlow, const float* dout,
                                float* dzlow_ws, float* dlogits_ws, float* dw, void* dw_workspace, void* dfeat,
                                int accumulate_dfeat, int n, int hf, int wf, int C, int k, int H, int W, void* stream) {
  IIC_REQUIRE(feat && w && zlow && dout && dzlow_ws && dlogits_ws && dw && dw_workspace && n > 0 && k > 0, IIC_ERR_BAD_ARG,
              "iic_seg_head_bwd: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int splits = seg_head_splits(n, hf, wf);
  if (dtype == IIC_F32)
    return seg_head_bwd_t<float>((const float*)feat, w, zlow, dout, dzlow_ws, dlogits_ws, dw, (float*)dw_workspace,
                                 (float*)dfeat, accumulate_dfeat, n, hf, wf, C, k, H, W, splits, st);
  if (dtype == IIC_BF16)
    return seg_head_bwd_t<__nv_bfloat16>((const __nv_bfloat16*)feat, w, zlow, dout, dzlow_ws, dlogits_ws, dw,
                                         (float*)dw_workspace, (__nv_bfloat16*)dfeat, accumulate_dfeat, n, hf, wf, C, k, H,
                                         W, splits, st);
  set_error("iic_seg_head_bwd: bad dtype");
  return IIC_ERR_BAD_ARG;
}
