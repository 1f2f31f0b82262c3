// Fused multi-tensor Adam (torch.optim.Adam semantics, amsgrad=False) -- the optimiser the
// reference builds in code/utils/cluster/general.py:8-9 and steps at
// code/scripts/cluster/cluster_sobel_twohead.py:355.  The reference launches ~5 kernels for each of
// ~110 parameter tensors; here one launch covers up to 48 tensors (pointer table in kernel args).
#include <math.h>

#include "common.cuh"

namespace iic {

constexpr int ADAM_MAX_T = 48;
constexpr int ADAM_CHUNK = 256 * 8;  // elements per CTA-iteration

struct AdamTable {
  float* param[ADAM_MAX_T];
  const float* grad[ADAM_MAX_T];
  float* m[ADAM_MAX_T];
  float* v[ADAM_MAX_T];
  long long size[ADAM_MAX_T];
  int chunk_begin[ADAM_MAX_T + 1];
  int count;
};

// step_dev != nullptr: the step count lives in device memory (a float scalar the caller increments before each launch)
// and the bias corrections are evaluated here -- the launch arguments then never change, so the step can be replayed from
// a CUDA graph (iic_b200/graph.py).
__global__ void __launch_bounds__(256) adam_kernel(const __grid_constant__ AdamTable tb, float lr, float b1, float b2,
                                                   float eps, float wd, float bc1, float bc2_sqrt,
                                                   const float* __restrict__ step_dev) {
  if (step_dev != nullptr) {
    const float step = *step_dev;
    bc1 = 1.f - powf(b1, step);
    bc2_sqrt = sqrtf(1.f - powf(b2, step));
  }
  const int total_chunks = tb.chunk_begin[tb.count];
  for (int ch = blockIdx.x; ch < total_chunks; ch += gridDim.x) {
    int t = 0;
    while (t + 1 < tb.count && tb.chunk_begin[t + 1] <= ch) ++t;
    const long long off = (long long)(ch - tb.chunk_begin[t]) * ADAM_CHUNK;
    const long long n = tb.size[t];
    float* __restrict__ p = tb.param[t];
    const float* __restrict__ g = tb.grad[t];
    float* __restrict__ m = tb.m[t];
    float* __restrict__ v = tb.v[t];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const long long i = off + j * 256 + threadIdx.x;
      if (i < n) {
        float gi = g[i];
        const float pi = p[i];
        if (wd != 0.f) gi = fmaf(wd, pi, gi);
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
      }
    }
  }
}

}  // namespace iic

using namespace iic;

static int adam_launch(const void* const* ptrs_host, const long long* sizes_host, int T, float lr, float beta1, float beta2,
                       float eps, float weight_decay, int step, const float* step_dev, void* stream);

extern "C" int iic_adam_step(const void* const* ptrs_host, const long long* sizes_host, int T, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int step, void* stream) {
  IIC_REQUIRE(ptrs_host && sizes_host && T > 0 && step > 0, IIC_ERR_BAD_ARG, "iic_adam_step: bad arguments");
  return adam_launch(ptrs_host, sizes_host, T, lr, beta1, beta2, eps, weight_decay, step, nullptr, stream);
}

extern "C" int iic_adam_step_dev(const void* const* ptrs_host, const long long* sizes_host, int T, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, const float* step_dev, void* stream) {
  IIC_REQUIRE(ptrs_host && sizes_host && T > 0 && step_dev, IIC_ERR_BAD_ARG, "iic_adam_step_dev: bad arguments");
  return adam_launch(ptrs_host, sizes_host, T, lr, beta1, beta2, eps, weight_decay, 1, step_dev, stream);
}

static int adam_launch(const void* const* ptrs_host, const long long* sizes_host, int T, float lr, float beta1, float beta2,
                       float eps, float weight_decay, int step, const float* step_dev, void* stream) {
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2s = sqrtf(1.f - powf(beta2, (float)step));
  for (int t0 = 0; t0 < T; t0 += ADAM_MAX_T) {
    AdamTable tb;
    tb.count = (T - t0) < ADAM_MAX_T ? (T - t0) : ADAM_MAX_T;
    int chunks = 0;
    for (int i = 0; i < tb.count; ++i) {
      const int t = t0 + i;
      tb.param[i] = (float*)ptrs_host[4 * t];
      tb.grad[i] = (const float*)ptrs_host[4 * t + 1];
      tb.m[i] = (float*)ptrs_host[4 * t + 2];
      tb.v[i] = (float*)ptrs_host[4 * t + 3];
      tb.size[i] = sizes_host[t];
      IIC_REQUIRE(tb.param[i] && tb.grad[i] && tb.m[i] && tb.v[i] && tb.size[i] > 0, IIC_ERR_BAD_ARG,
                  "iic_adam_step: tensor %d has a null pointer or empty size", t);
      tb.chunk_begin[i] = chunks;
      chunks += (int)((tb.size[i] + ADAM_CHUNK - 1) / ADAM_CHUNK);
    }
    tb.chunk_begin[tb.count] = chunks;
    int blocks = chunks;
    const int cap = device_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(tb, lr, beta1, beta2, eps, weight_decay, bc1, bc2s, step_dev);
    IIC_LAUNCH_CHECK();
    count_launch();
  }
  return IIC_OK;
}
