// Probe: bf16 (kind::f16) MN-major operand with OVERLAPPING atoms on a linear pixel-major buffer buf[pixel][16 channels]
// (32 bytes per pixel, SWIZZLE_32B):  A[m = v*16 + c][k = pixel] = buf[pixel + v][c]  -- descriptor LBO = 32 B (next M atom =
// next pixel), SBO = 256 B (8 K rows), K = 16 pixels per MMA; B[n][k] = bufB[k][n] (MN-major, one 16-wide atom).
// This is the operand of the forward segmentation joint (csrc/seg_joint_tc.cu).  Small integers: exact.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I iic_b200/csrc tools/umma_bf16_mn_probe.cu -o tools/umma_bf16_mn_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <vector>

#include "tc_ptx.cuh"

using namespace iic;
constexpr int PIX = 192, KPIX = 32;

__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ uint32_t sw32(uint32_t a) { return a ^ (((a >> 7) & 1u) << 4); }

__global__ void __launch_bounds__(128, 1) probe(const float* __restrict__ src, const float* __restrict__ bsrc, float* __restrict__ out,
                                                int shift, int apply_sw) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t abuf = (raw + 1023u) & ~1023u;  // PIX x 32 B
  const uint32_t bbuf = abuf + PIX * 32;         // [KPIX][16] rows of 32 B
  const uint32_t bar = bbuf + 2048;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (bar - raw) + 16);
  uint8_t* base = smem_raw + (abuf - raw);
  for (int i = threadIdx.x; i < PIX * 16; i += blockDim.x) {
    const uint32_t a = (uint32_t)i * 2u;
    *reinterpret_cast<__nv_bfloat16*>(base + (apply_sw ? sw32(a) : a)) = __float2bfloat16(src[i]);
  }
  for (int i = threadIdx.x; i < KPIX * 16; i += blockDim.x) {
    const uint32_t a = (uint32_t)i * 2u;
    *reinterpret_cast<__nv_bfloat16*>(base + PIX * 32 + (apply_sw ? sw32(a) : a)) = __float2bfloat16(bsrc[i]);
  }
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(tmem_slot), 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t layout = apply_sw ? 6u : 0u;
    const uint32_t sa = abuf + (uint32_t)shift * 32u;
    for (int kk = 0; kk < KPIX / 16; ++kk)
      umma_bf16(tmem, desc(sa + kk * 512, 32, 256, layout), desc(bbuf + kk * 512, 32, 256, layout), make_idesc(16, 1, 1), kk > 0 ? 1u : 0u);
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int e = 0; e < 16; ++e) out[(warp * 32 + lane) * 16 + e] = __uint_as_float(v[e]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 32);
}

int main() {
  std::vector<float> h(PIX * 16), b0(KPIX * 16);
  for (int p = 0; p < PIX; ++p)
    for (int c = 0; c < 16; ++c) h[p * 16 + c] = (float)(((p * 37 + c * 11) % 13) - 6);
  for (int k = 0; k < KPIX; ++k)
    for (int n = 0; n < 16; ++n) b0[k * 16 + n] = (float)(((k * 5 + n * 3) % 7) - 3);
  float *d_src, *d_b, *d_out;
  cudaMalloc(&d_src, h.size() * 4);
  cudaMalloc(&d_b, b0.size() * 4);
  cudaMalloc(&d_out, 128 * 16 * 4);
  cudaMemcpy(d_src, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(d_b, b0.data(), b0.size() * 4, cudaMemcpyHostToDevice);
  const int smem = PIX * 32 + 2048 + 1024 + 256;
  std::vector<float> o(128 * 16);
  const int shifts[] = {0, 1, 2, 3, 7, 8, 13, 16, 21};
  for (int sw = 1; sw >= 0; --sw)
    for (int s : shifts) {
      cudaMemset(d_out, 0xff, 128 * 16 * 4);
      probe<<<1, 128, smem>>>(d_src, d_b, d_out, s, sw);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("swizzle %d shift %2d: CUDA error %s\n", sw, s, cudaGetErrorString(e));
        return 1;
      }
      cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0, fm = -1, fn = -1;
      double fw = 0;
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 16; ++n) {
          double want = 0;
          for (int k = 0; k < KPIX; ++k) want += (double)h[(k + s + m / 16) * 16 + (m % 16)] * b0[k * 16 + n];
          if ((double)o[m * 16 + n] != want) {
            if (!bad) fm = m, fn = n, fw = want;
            ++bad;
          }
        }
      printf("bf16 MN-major Toeplitz A, %s shift %2d: %s (%d / 2048 mismatches", sw ? "SWIZZLE_32B" : "no swizzle ", s,
             bad ? "MISMATCH" : "ok", bad);
      if (bad) printf(", first at m=%d n=%d got %g want %g", fm, fn, o[fm * 16 + fn], fw);
      printf(")\n");
    }
  return 0;
}
