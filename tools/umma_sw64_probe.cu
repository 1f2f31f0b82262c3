// Probe for the tensor-core formulation of the uncollapsed segmentation joint (DESIGN "next"):
// can a LINEAR pixel-major fp32 buffer  buf[pixel][16 channels]  (64 bytes per pixel, written to shared memory with
// the SWIZZLE_64B pattern, as a TMA box with a 64-byte inner extent would) be read by tcgen05.mma kind::tf32 as
//
//   mode 0  MN-major A (forward joint):   A[m = v*16 + c][k = pixel] = buf[pixel + v][c]
//           -- the Toeplitz "displaced copies" operand: M atoms of 16 channels, LBO = 64 B (atom v+1 of pixel p is atom v
//           of pixel p+1: the atoms OVERLAP the K rows), SBO = 512 B, K rows 64 B apart;
//   mode 1  K-major A (backward):         A[m = pixel][k = v*16 + c] = buf[pixel + v][c]
//           -- rows 64 B apart that overlap along K: the K offset is applied by moving the start address (v*64 + 32*half).
//
// B is a small dense matrix (N = 16), MN-major in mode 0 ([k][n] rows of 64 B) and K-major in mode 1.
// All values are small integers (exact in tf32 and in the fp32 accumulator), so any mismatch is a layout fact.
//
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I iic_b200/csrc tools/umma_sw64_probe.cu -o tools/umma_sw64_probe
//   run  : tools/umma_sw64_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

#include "tc_ptx.cuh"

using namespace iic;

constexpr int PIX = 192;   // pixels in the linear buffer
constexpr int KPIX = 32;   // mode 0: pixels contracted (4 MMAs of K = 8)
constexpr int VSPAN = 8;   // displacements per M = 128 tile (8 * 16 channels)

__device__ __forceinline__ uint64_t desc64(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;  // 4 = SWIZZLE_64B, 2 = SWIZZLE_128B, 0 = none
  return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__host__ __device__ constexpr uint32_t idesc_tf32(int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}
// byte offset of logical byte address `a` inside a SWIZZLE_64B region (Swizzle<2,4,3>: bits [4,5] ^= bits [7,8])
__device__ __forceinline__ uint32_t sw64(uint32_t a) { return a ^ (((a >> 7) & 3u) << 4); }

// mode 0: D[m][n] = sum_{k < KPIX} buf[k + shift + m/16][m%16] * B0[k][n]
// mode 1: D[m][n] = sum_{v < VK} sum_c buf[m + shift + v][c] * B1[n][v*16 + c]       (VK = 4 displacements, K = 64)
__global__ void __launch_bounds__(128, 1) probe(const float* __restrict__ src /* [PIX][16] */, const float* __restrict__ bsrc,
                                                float* __restrict__ out /* [128][16] */, int mode, int shift, int apply_sw) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t abuf = (raw + 1023u) & ~1023u;  // PIX x 64 B
  const uint32_t bbuf = abuf + PIX * 64;         // mode 0: [KPIX][16] rows of 64 B; mode 1: [4 v][16 n][64 B]
  const uint32_t bar = bbuf + 4096;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (bar - raw) + 16);
  uint8_t* base = smem_raw + (abuf - raw);
  for (int i = threadIdx.x; i < PIX * 16; i += blockDim.x) {
    const uint32_t a = (uint32_t)i * 4u;  // linear byte address inside the (1024-aligned) buffer
    *reinterpret_cast<float*>(base + (apply_sw ? sw64(a) : a)) = src[i];
  }
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
    // mode 0: bsrc[k][n] -> row k (64 B) ; mode 1: bsrc[n][v*16 + c] -> [v][n][c]
    uint32_t a;
    if (mode == 0) a = (uint32_t)i * 4u;
    else {
      const int n = i / 64, kk = i % 64, v = kk / 16, c = kk % 16;
      a = (uint32_t)((v * 16 + n) * 64 + c * 4);
    }
    *reinterpret_cast<float*>(base + PIX * 64 + (apply_sw ? sw64(a) : a)) = bsrc[i];
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
  const uint32_t layout = apply_sw ? 4u : 0u;
  if (threadIdx.x == 0) {
    const uint32_t sa = abuf + (uint32_t)shift * 64u;
    if (mode == 0) {
      for (int kk = 0; kk < KPIX / 8; ++kk) {
        const uint64_t ad = desc64(sa + kk * 512, 64, 512, layout);    // 8 K rows (pixels) = 512 B per MMA
        const uint64_t bd = desc64(bbuf + kk * 512, 64, 512, layout);  // one 16-wide N atom
        mma_tf32(tmem, ad, bd, idesc_tf32(16, 1, 1), kk > 0 ? 1u : 0u);
      }
    } else {
      for (int v = 0; v < 4; ++v)
        for (int half = 0; half < 2; ++half) {
          const uint64_t ad = desc64(sa + v * 64 + half * 32, 16, 512, layout);            // rows 64 B apart, 8-row groups 512 B
          const uint64_t bd = desc64(bbuf + v * 1024 + half * 32, 16, 512, layout);        // [v][16 rows][64 B]
          mma_tf32(tmem, ad, bd, idesc_tf32(16, 0, 0), (v > 0 || half > 0) ? 1u : 0u);
        }
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t v[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);  // 32 columns allocated, 16 used
  tmem_ld_wait();
  for (int e = 0; e < 16; ++e) out[(warp * 32 + lane) * 16 + e] = __uint_as_float(v[e]);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 32);
}

int main() {
  std::vector<float> h(PIX * 16), b0(KPIX * 16), b1(16 * 64);
  for (int p = 0; p < PIX; ++p)
    for (int c = 0; c < 16; ++c) h[p * 16 + c] = (float)(((p * 37 + c * 11) % 13) - 6);
  for (int k = 0; k < KPIX; ++k)
    for (int n = 0; n < 16; ++n) b0[k * 16 + n] = (float)(((k * 5 + n * 3) % 7) - 3);
  for (int n = 0; n < 16; ++n)
    for (int kk = 0; kk < 64; ++kk) b1[n * 64 + kk] = (float)(((n * 7 + kk * 3) % 5) - 2);
  float *d_src, *d_b, *d_out;
  cudaMalloc(&d_src, h.size() * 4);
  cudaMalloc(&d_b, 1024 * 4);
  cudaMalloc(&d_out, 128 * 16 * 4);
  cudaMemcpy(d_src, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  const int smem = PIX * 64 + 4096 + 1024 + 256;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> o(128 * 16);
  const int shifts[] = {0, 1, 2, 3, 7, 8, 13, 16, 21};
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemcpy(d_b, mode == 0 ? b0.data() : b1.data(), (mode == 0 ? b0.size() : b1.size()) * 4, cudaMemcpyHostToDevice);
    for (int sw = 1; sw >= 0; --sw)
      for (int s : shifts) {
        cudaMemset(d_out, 0xff, 128 * 16 * 4);
        probe<<<1, 128, smem>>>(d_src, d_b, d_out, mode, s, sw);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
          printf("mode %d swizzle %d shift %2d: CUDA error %s\n", mode, sw, s, cudaGetErrorString(e));
          return 1;
        }
        cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, fm = -1, fn = -1;
        double fw = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 16; ++n) {
            double want = 0;
            if (mode == 0) {
              for (int k = 0; k < KPIX; ++k) want += (double)h[(k + s + m / 16) * 16 + (m % 16)] * b0[k * 16 + n];
            } else {
              for (int v = 0; v < 4; ++v)
                for (int c = 0; c < 16; ++c) want += (double)h[(m + s + v) * 16 + c] * b1[n * 64 + v * 16 + c];
            }
            if ((double)o[m * 16 + n] != want) {
              if (!bad) fm = m, fn = n, fw = want;
              ++bad;
            }
          }
        printf("mode %d (%s A) %s shift %2d: %s (%d / 2048 mismatches", mode, mode == 0 ? "MN-major Toeplitz" : "K-major overlapped",
               sw ? "SWIZZLE_64B" : "no swizzle ", s, bad ? "MISMATCH" : "ok", bad);
        if (bad) printf(", first at m=%d n=%d got %g want %g", fm, fn, o[fm * 16 + fn], fw);
        printf(")\n");
      }
  }
  return 0;
}
