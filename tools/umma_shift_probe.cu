// Probe: can a UMMA shared-memory descriptor (SWIZZLE_128B) start at an arbitrary 128-byte row of a
// 1024-byte-aligned swizzled buffer (a "shifted view"), and what must the matrix-base-offset field
// (descriptor bits 49..51) hold then?  This decides whether one halo tile in shared memory can feed
// all filter taps of a 3x3 convolution ("next": halo reuse for the 64-channel layers).
//
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I iic_b200/csrc tools/umma_shift_probe.cu -o tools/umma_shift_probe
//   run  : tools/umma_shift_probe        (prints one line per (layout, shift, base_offset) with the mismatch count)
#include <cstdio>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <vector>

#include "tc_ptx.cuh"

using namespace iic;

constexpr int ROWS = 256;  // 128-byte rows in the source buffer

__device__ __forceinline__ uint64_t desc_bo(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t base_off) {
  return make_desc(saddr, lbo, sbo) | ((uint64_t)(base_off & 7u) << 49);
}

// mode 0: A K-major   (row = M index, 64 K elements per row);  D[m][n] = A[m + shift][n]      (B = I)
// mode 1: A MN-major  (row = K index, 64 M elements per row);  D[m][n] = A[n + shift][m]      (B = I, MN-major)
__global__ void __launch_bounds__(128, 1) probe_kernel(const __nv_bfloat16* __restrict__ src /* [ROWS][64] logical */,
                                                       float* __restrict__ out /* [128][64] */, int mode, int shift,
                                                       int base_off) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t abuf = (raw + 1023u) & ~1023u;          // ROWS x 128 B (+ second 64-column block for mode 1)
  const uint32_t a_bytes = ROWS * 128u * 2u;
  const uint32_t bbuf = abuf + a_bytes;                  // 64 x 128 B identity
  const uint32_t bar = bbuf + 64 * 128;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (bar - raw) + 16);
  uint8_t* ap = smem_raw + (abuf - raw);
  uint8_t* bp = smem_raw + (bbuf - raw);
  // logical element (r, c) of a 128-byte-row buffer lives at r*128 + ((c/8) ^ (r%8))*16 + (c%8)*2
  for (int i = threadIdx.x; i < ROWS * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const uint32_t off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(ap + off) = src[i];
    // second 64-column block (mode 1, M = 128): rows hold elements 64..127 = value + 0.5 marker via different table
    *reinterpret_cast<__nv_bfloat16*>(ap + ROWS * 128 + off) = __float2bfloat16(__bfloat162float(src[i]) + 64.f);
  }
  for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
    const int r = i / 64, c = i % 64;
    const uint32_t off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(bp + off) = __float2bfloat16(r == c ? 1.f : 0.f);
  }
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) tmem_alloc(smem_u32(tmem_slot), 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t sa = abuf + (uint32_t)shift * 128u;
    for (int kk = 0; kk < 4; ++kk) {
      uint64_t ad, bd;
      if (mode == 0) {
        ad = desc_bo(sa + kk * 32, 16, 1024, base_off);
        bd = make_desc(bbuf + kk * 32, 16, 1024);
        umma_bf16(tmem, ad, bd, make_idesc(64, 0, 0), kk > 0 ? 1u : 0u);
      } else {
        ad = desc_bo(sa + kk * 2048, ROWS * 128, 1024, base_off);  // LBO = distance between the two 64-column blocks
        bd = make_desc(bbuf + kk * 2048, 8192, 1024);
        umma_bf16(tmem, ad, bd, make_idesc(64, 1, 1), kk > 0 ? 1u : 0u);
      }
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int e = 0; e < 32; ++e) out[(warp * 32 + lane) * 64 + c0 + e] = __uint_as_float(v[e]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 64);
}

int main() {
  std::vector<__nv_bfloat16> h(ROWS * 64);
  std::vector<float> hf(ROWS * 64);
  for (int r = 0; r < ROWS; ++r)
    for (int c = 0; c < 64; ++c) {
      // distinct per (r mod 32, c): small integers, exact in bf16
      const float v = (float)(((r * 37 + c * 11) % 61) - 30);
      hf[r * 64 + c] = v;
      h[r * 64 + c] = __float2bfloat16(v);
    }
  __nv_bfloat16* d_src;
  float* d_out;
  cudaMalloc(&d_src, h.size() * sizeof(__nv_bfloat16));
  cudaMalloc(&d_out, 128 * 64 * sizeof(float));
  cudaMemcpy(d_src, h.data(), h.size() * sizeof(__nv_bfloat16), cudaMemcpyHostToDevice);
  const int smem = ROWS * 128 * 2 + 64 * 128 + 1024 + 256;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int shifts[] = {0, 1, 2, 3, 5, 8, 9, 51, 52, 53, 102, 104};
  std::vector<float> o(128 * 64);
  for (int mode = 0; mode < 2; ++mode)
    for (int s : shifts)
      for (int variant = 0; variant < 2; ++variant) {
        const int bo = variant == 0 ? 0 : (s & 7);
        if (variant == 1 && bo == 0) continue;
        cudaMemset(d_out, 0xff, 128 * 64 * sizeof(float));
        probe_kernel<<<1, 128, smem>>>(d_src, d_out, mode, s, bo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mode %d shift %3d base_offset %d: CUDA error %s\n", mode, s, bo, cudaGetErrorString(e));
          return 1;
        }
        cudaMemcpy(o.data(), d_out, o.size() * sizeof(float), cudaMemcpyDeviceToHost);
        int bad = 0, first_m = -1, first_n = -1;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            float want;
            if (mode == 0) want = hf[(m + s) * 64 + n];
            else want = hf[(n + s) * 64 + (m & 63)] + (m >= 64 ? 64.f : 0.f);
            if (o[m * 64 + n] != want) {
              if (!bad) first_m = m, first_n = n;
              ++bad;
            }
          }
        printf("mode %d (%s) shift %3d base_offset %d: %s (%d / 8192 mismatches", mode, mode == 0 ? "K-major" : "MN-major", s,
               bo, bad ? "MISMATCH" : "ok", bad);
        if (bad) printf(", first at m=%d n=%d got %g", first_m, first_n, o[first_m * 64 + first_n]);
        printf(")\n");
      }
  return 0;
}
