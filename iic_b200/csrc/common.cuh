// Shared helpers for the iic_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/iic_b200.h"

namespace iic {

void set_error(const char* fmt, ...);

#define IIC_REQUIRE(cond, code, ...)          \
  do {                                        \
    if (!(cond)) {                            \
      ::iic::set_error(__VA_ARGS__);          \
      return (code);                          \
    }                                         \
  } while (0)

#define IIC_CUDA(expr)                                                                    \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::iic::set_error("%s:%d CUDA error %d (%s) in %s", __FILE__, __LINE__, (int)_e,     \
                       cudaGetErrorString(_e), #expr);                                    \
      return IIC_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

#define IIC_LAUNCH_CHECK() IIC_CUDA(cudaGetLastError())

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- storage-type helpers (activations are fp32 or bf16, NHWC) -------------------------
template <typename T> struct Vec8;  // 8 consecutive channels
template <> struct Vec8<float> {
  float4 lo, hi;
};
template <> struct Vec8<__nv_bfloat16> {
  uint4 v;
};

__device__ __forceinline__ void load8(const float* p, float (&f)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&f)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(f[4], f[5], f[6], f[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = v;
}
// raw (unconverted) 8-element vectors: keep several loads in flight without paying fp32 registers for bf16 data
struct Raw8f { float4 a, b; };
struct Raw8h { uint4 v; };
__device__ __forceinline__ void load_raw(const float* p, Raw8f& r) {
  r.a = *reinterpret_cast<const float4*>(p);
  r.b = *reinterpret_cast<const float4*>(p + 4);
}
__device__ __forceinline__ void load_raw(const __nv_bfloat16* p, Raw8h& r) { r.v = *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void cvt_raw(const Raw8f& r, float (&f)[8]) {
  f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w;
  f[4] = r.b.x; f[5] = r.b.y; f[6] = r.b.z; f[7] = r.b.w;
}
__device__ __forceinline__ void cvt_raw(const Raw8h& r, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r.v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
template <typename T> struct RawOf;
template <> struct RawOf<float> { using type = Raw8f; };
template <> struct RawOf<__nv_bfloat16> { using type = Raw8h; };

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Deterministic block-wide sum of doubles (all threads get the result). `red` = 33 doubles of smem.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = lane < nw ? red[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

int device_sm_count();
void count_launch();
// runtime kernel-variant switches (api.cu); defaults from the environment, iic_set_option overrides
enum { OPT_CONV_HALO = 0, OPT_CONV_HALO_WGRAD = 1, OPT_STEM_QUAD = 2, OPT_DGRAD_PREFETCH = 3, OPT_TC2_MT2 = 4, OPT_HALO_STORE = 5, OPT_STEM_BWD_V2 = 6, OPT_HALO_STATS = 7, OPT_BN_BWD_CTAS = 8, OPT_TF32X3_RAW_HI = 9, OPT_WGRAD_MT = 10, OPT_HALO_ADDEND_TMA = 11, OPT_DGRAD_S2_MT = 12, OPT_COUNT = 13 };
int option(int id);

}  // namespace iic
