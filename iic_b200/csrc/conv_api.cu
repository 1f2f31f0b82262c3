// Dispatch of the conv entry points onto the kernels: IIC_F32 -> fp32 SIMT implicit GEMM (conv_simt.cu),
// IIC_BF16 -> tcgen05 kind::f16 implicit GEMM on bf16 activations (conv_tc2.cu), IIC_TF32 / IIC_TF32X3 -> tcgen05
// kind::tf32 on fp32 activations, plain or 3xTF32 error-compensated (conv_tf32.cu).  These are precision modes of
// one sm_100a code base, not a backend switch: there is no CPU or library fallback.
#include <stdlib.h>

#include "common.cuh"

namespace iic {
int simt_conv_fprop(const float* x, const float* w, float* y, const iic_conv_geom* g, cudaStream_t st);
int simt_conv_dgrad(const float* dy, const float* wt, const float* addend, float* dx, const iic_conv_geom* g, cudaStream_t st);
long long simt_conv_wgrad_workspace(const iic_conv_geom* g);
int simt_conv_wgrad(const float* x, const float* dy, float* dw, float* ws, const iic_conv_geom* g, cudaStream_t st);
int tc2_conv_gather_gemm(const __nv_bfloat16* src, int srcH, int srcW, int srcC, int rowH, int rowW, int nimg,
                         const iic_conv_geom* g, int transposed, const __nv_bfloat16* wpacked, int N,
                         const __nv_bfloat16* addend, __nv_bfloat16* out, cudaStream_t st);
long long tc2_conv_wgrad_workspace(const iic_conv_geom* g);
int tc2_conv_wgrad(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, float* ws, const iic_conv_geom* g, cudaStream_t st);
int tc2_conv_wgrad_oihw(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* grad_oihw, int accumulate, float* ws,
                        const iic_conv_geom* g, cudaStream_t st);
int tc2_conv_fprop_blocks(const iic_conv_geom* g);
int tc2_conv_gather_gemm_stats(const __nv_bfloat16* src, int srcH, int srcW, int srcC, int rowH, int rowW, int nimg,
                               const iic_conv_geom* g, int transposed, const __nv_bfloat16* wpacked, int N,
                               const __nv_bfloat16* addend, __nv_bfloat16* out, float* stat_partial, int stat_groups,
                               cudaStream_t st);
int tc2_conv_dgrad_s2(const __nv_bfloat16* dy, const __nv_bfloat16* wpacked_t, const __nv_bfloat16* addend,
                      __nv_bfloat16* dx, const iic_conv_geom* g, cudaStream_t st);
int tc2_conv_dgrad_masked(const __nv_bfloat16* dy, const __nv_bfloat16* wpacked_t, const __nv_bfloat16* addend,
                          const unsigned char* addend_mask, __nv_bfloat16* dx, const iic_conv_geom* g, cudaStream_t st);
int tf32_conv_gather_gemm(const float* src, int srcH, int srcW, int srcC, int rowH, int rowW, int nimg, const iic_conv_geom* g,
                          int transposed, const float* wpacked, int N, const float* addend, float* out, int split,
                          cudaStream_t st);
int tf32_conv_dgrad_s2(const float* dy, const float* wpacked_t, const float* addend, float* dx, const iic_conv_geom* g, int split,
                       cudaStream_t st);
long long tf32_conv_wgrad_workspace(const iic_conv_geom* g);
int tf32_conv_wgrad(const float* x, const float* dy, float* dw, float* ws, const iic_conv_geom* g, int split, cudaStream_t st);
}  // namespace iic

using namespace iic;

// IIC_TF32 / IIC_TF32X3: fp32 storage, tcgen05 kind::tf32 (conv_tf32.cu), plain or 3xTF32 error-compensated
static bool is_tf32(int dtype) { return dtype == IIC_TF32 || dtype == IIC_TF32X3; }
static int tf32_split(int dtype) { return dtype == IIC_TF32X3 ? 3 : 1; }

// (Round 1's cp.async-fed tcgen05 kernel, conv_tc.cu, was removed in round 2: the TMA-fed persistent kernels cover
// fprop of any stride, stride-1 dgrad, stride-2 dgrad by parity classes and wgrad; other dgrad strides are unsupported.)

static int geom_check(const iic_conv_geom* g, const char* who) {
  IIC_REQUIRE(g != nullptr, IIC_ERR_BAD_ARG, "%s: null geometry", who);
  IIC_REQUIRE(g->n > 0 && g->h > 0 && g->w > 0 && g->cin > 0 && g->cout > 0 && g->kh > 0 && g->kw > 0 &&
                  g->stride > 0 && g->pad >= 0 && g->dil > 0,
              IIC_ERR_BAD_ARG, "%s: bad geometry", who);
  IIC_REQUIRE(g->oh == (g->h + 2 * g->pad - g->dil * (g->kh - 1) - 1) / g->stride + 1 &&
                  g->ow == (g->w + 2 * g->pad - g->dil * (g->kw - 1) - 1) / g->stride + 1,
              IIC_ERR_BAD_ARG, "%s: output size %dx%d inconsistent with the geometry", who, g->oh, g->ow);
  IIC_REQUIRE((long long)g->n * g->h * g->w < (1ll << 31) && (long long)g->n * g->oh * g->ow < (1ll << 31),
              IIC_ERR_UNSUPPORTED, "%s: more than 2^31 pixels", who);
  return IIC_OK;
}

extern "C" int iic_conv_fprop(const void* x, const void* w_packed, void* y, const iic_conv_geom* g, int dtype, void* stream) {
  int rc = geom_check(g, "iic_conv_fprop");
  if (rc != IIC_OK) return rc;
  IIC_REQUIRE(x && w_packed && y, IIC_ERR_BAD_ARG, "iic_conv_fprop: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IIC_F32) return simt_conv_fprop((const float*)x, (const float*)w_packed, (float*)y, g, st);
  if (is_tf32(dtype))
    return tf32_conv_gather_gemm((const float*)x, g->h, g->w, g->cin, g->oh, g->ow, g->n, g, 0, (const float*)w_packed, g->cout,
                                 nullptr, (float*)y, tf32_split(dtype), st);
  if (dtype == IIC_BF16)
    return tc2_conv_gather_gemm((const __nv_bfloat16*)x, g->h, g->w, g->cin, g->oh, g->ow, g->n, g, 0,
                                (const __nv_bfloat16*)w_packed, g->cout, nullptr, (__nv_bfloat16*)y, st);
  set_error("iic_conv_fprop: bad dtype %d", dtype);
  return IIC_ERR_BAD_ARG;
}

extern "C" int iic_conv_dgrad(const void* dy, const void* w_packed_t, const void* addend, void* dx, const iic_conv_geom* g,
                              int dtype, void* stream) {
  int rc = geom_check(g, "iic_conv_dgrad");
  if (rc != IIC_OK) return rc;
  IIC_REQUIRE(dy && w_packed_t && dx, IIC_ERR_BAD_ARG, "iic_conv_dgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IIC_F32)
    return simt_conv_dgrad((const float*)dy, (const float*)w_packed_t, (const float*)addend, (float*)dx, g, st);
  if (is_tf32(dtype) && g->stride == 2)
    return tf32_conv_dgrad_s2((const float*)dy, (const float*)w_packed_t, (const float*)addend, (float*)dx, g, tf32_split(dtype),
                              st);
  if (is_tf32(dtype))
    return tf32_conv_gather_gemm((const float*)dy, g->oh, g->ow, g->cout, g->h, g->w, g->n, g, 1, (const float*)w_packed_t,
                                 g->cin, (const float*)addend, (float*)dx, tf32_split(dtype), st);
  if ((dtype == IIC_BF16 || is_tf32(dtype)) && g->stride != 1)
    IIC_REQUIRE(g->stride == 2 && g->dil == 1 && g->kh == g->kw, IIC_ERR_UNSUPPORTED,
                "iic_conv_dgrad (tensor-core modes): stride 1, or stride 2 with dilation 1 and a square filter");
  if (dtype == IIC_BF16 && g->stride == 2)
    return tc2_conv_dgrad_s2((const __nv_bfloat16*)dy, (const __nv_bfloat16*)w_packed_t, (const __nv_bfloat16*)addend,
                             (__nv_bfloat16*)dx, g, st);
  if (dtype == IIC_BF16)
    return tc2_conv_gather_gemm((const __nv_bfloat16*)dy, g->oh, g->ow, g->cout, g->h, g->w, g->n, g, 1,
                                (const __nv_bfloat16*)w_packed_t, g->cin, (const __nv_bfloat16*)addend, (__nv_bfloat16*)dx, st);
  set_error("iic_conv_dgrad: bad dtype %d", dtype);
  return IIC_ERR_BAD_ARG;
}

extern "C" int iic_conv_dgrad_masked(const void* dy, const void* w_packed_t, const void* addend, const unsigned char* addend_mask,
                                     void* dx, const iic_conv_geom* g, int dtype, void* stream) {
  int rc = geom_check(g, "iic_conv_dgrad_masked");
  if (rc != IIC_OK) return rc;
  IIC_REQUIRE(dy && w_packed_t && addend && addend_mask && dx, IIC_ERR_BAD_ARG, "iic_conv_dgrad_masked: null pointer");
  IIC_REQUIRE(dtype == IIC_BF16 && g->stride == 1 && g->cin % 64 == 0, IIC_ERR_UNSUPPORTED,
              "iic_conv_dgrad_masked: bf16, stride 1, cin a multiple of 64");
  return tc2_conv_dgrad_masked((const __nv_bfloat16*)dy, (const __nv_bfloat16*)w_packed_t, (const __nv_bfloat16*)addend,
                               addend_mask, (__nv_bfloat16*)dx, g, (cudaStream_t)stream);
}

extern "C" long long iic_conv_wgrad_workspace(const iic_conv_geom* g, int dtype) {
  if (g == nullptr) return -1;
  if (dtype == IIC_BF16) return tc2_conv_wgrad_workspace(g);
  if (is_tf32(dtype)) return tf32_conv_wgrad_workspace(g);
  return simt_conv_wgrad_workspace(g);
}

extern "C" int iic_conv_wgrad(const void* x, const void* dy, float* dw_packed, void* workspace, const iic_conv_geom* g,
                              int dtype, void* stream) {
  int rc = geom_check(g, "iic_conv_wgrad");
  if (rc != IIC_OK) return rc;
  IIC_REQUIRE(x && dy && dw_packed && workspace, IIC_ERR_BAD_ARG, "iic_conv_wgrad: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IIC_F32) return simt_conv_wgrad((const float*)x, (const float*)dy, dw_packed, (float*)workspace, g, st);
  if (is_tf32(dtype))
    return tf32_conv_wgrad((const float*)x, (const float*)dy, dw_packed, (float*)workspace, g, tf32_split(dtype), st);
  if (dtype == IIC_BF16)
    return tc2_conv_wgrad((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, dw_packed, (float*)workspace, g, st);
  set_error("iic_conv_wgrad: bad dtype %d", dtype);
  return IIC_ERR_BAD_ARG;
}

// fprop with the BatchNorm statistics of the output fused into the epilogue (tensor-core path only): the
// separate statistics pass (one full read of y) disappears.  Returns the number of partial rows written.
extern "C" int iic_conv_fprop_stats_blocks(const iic_conv_geom* g, int dtype) {
  if (g == nullptr || dtype != IIC_BF16) return 0;
  return tc2_conv_fprop_blocks(g);
}

extern "C" int iic_conv_fprop_stats(const void* x, const void* w_packed, void* y, const iic_conv_geom* g, int dtype,
                                    int views, float* stat_partial, void* stream) {
  int rc = geom_check(g, "iic_conv_fprop_stats");
  if (rc != IIC_OK) return rc;
  IIC_REQUIRE(x && w_packed && y && stat_partial, IIC_ERR_BAD_ARG, "iic_conv_fprop_stats: null pointer");
  IIC_REQUIRE(dtype == IIC_BF16, IIC_ERR_UNSUPPORTED, "iic_conv_fprop_stats: tensor-core (bf16) path only");
  return tc2_conv_gather_gemm_stats((const __nv_bfloat16*)x, g->h, g->w, g->cin, g->oh, g->ow, g->n, g, 0,
                                    (const __nv_bfloat16*)w_packed, g->cout, nullptr, (__nv_bfloat16*)y, stat_partial, views,
                                    (cudaStream_t)stream);
}

// wgrad straight into the torch-layout gradient (optionally accumulating): on the bf16 tensor-core path the split-K fold
// writes [cout][cin][kh][kw] itself; the other modes run iic_conv_wgrad into the tail of the workspace + iic_unpack_wgrad.
extern "C" long long iic_conv_wgrad_oihw_workspace(const iic_conv_geom* g, int dtype) {
  const long long base = iic_conv_wgrad_workspace(g, dtype);
  if (base < 0 || dtype == IIC_BF16) return base;
  return base + (long long)g->cout * g->cin * g->kh * g->kw * (long long)sizeof(float) + 256;
}

extern "C" int iic_conv_wgrad_oihw(const void* x, const void* dy, float* grad_oihw, int accumulate, void* workspace,
                                   const iic_conv_geom* g, int dtype, void* stream) {
  int rc = geom_check(g, "iic_conv_wgrad_oihw");
  if (rc != IIC_OK) return rc;
  IIC_REQUIRE(x && dy && grad_oihw && workspace, IIC_ERR_BAD_ARG, "iic_conv_wgrad_oihw: null pointer");
  if (dtype == IIC_BF16)
    return tc2_conv_wgrad_oihw((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy, grad_oihw, accumulate, (float*)workspace, g,
                               (cudaStream_t)stream);
  const long long base = (iic_conv_wgrad_workspace(g, dtype) + 255) / 256 * 256;
  float* packed = (float*)((char*)workspace + base);
  rc = iic_conv_wgrad(x, dy, packed, workspace, g, dtype, stream);
  if (rc != IIC_OK) return rc;
  return iic_unpack_wgrad(packed, grad_oihw, accumulate, g->cout, g->cin, g->kh, g->kw, stream);
}
