// Stem weight gradient on tcgen05 (bf16 mode; kernels.STEM_WGRAD_TC).
//
// dW[cout][ci*kh*kw + a*kw + b] = sum over the n*H*W output pixels of dy[pixel][cout] * x[n][ci][y+a-pad][x+b-pad] is a GEMM
// with K = 13 M pixels at the bench shape, one 64-wide operand (dy, NHWC bf16: 1.66 GB, the only real traffic) and one
// 9..32-wide operand (the patches of the NCHW fp32 input, 104 MB).  The SIMT Gram-product kernel (stem.cu) runs it at
// ~15 TFLOP/s = 2.05 ms, 5 % of the c4 step; writing the patches out and calling the 1x1 tcgen05 wgrad costs more than it
// saves (3.7 ms, profiles/r02_session_f.md).  Here the patch operand never leaves the SM:
//
//   D[tap (M = 128 rows, taps >= K stay zero)][cout (N = 64)] += A[tap][64 pixels] * B[64 pixels][cout]
//
//   * B = one TMA box of dy per k-block of 64 pixels (dy viewed as a [pixels][64] matrix: rows of 128 B = one SWIZZLE_128B
//     atom row, MN-major operand; rows past the end are zero-filled by TMA);
//   * A = the patches of those 64 pixels, gathered from x by four builder warps (thread = pixel, coalesced along the image
//     row), rounded to bf16 and written K-major into the same swizzle pattern a TMA load would have produced;
//   * one elected thread issues four kind::f16 MMAs (K = 16 each) per k-block into ONE accumulator that lives in TMEM for
//     the whole persistent CTA; at the end lanes 0..31 of it go to a per-CTA partial, folded in a fixed order by a second
//     launch (deterministic, no atomics).
#include <cuda.h>

#include "tc_ptx.cuh"

namespace iic {

constexpr int SWT_STAGES = 8;
constexpr int SWT_A_BYTES = 128 * 128;  // [128 tap rows][64 pixels] bf16, K-major
constexpr int SWT_B_BYTES = 64 * 128;   // [64 pixels][64 cout] bf16, MN-major
constexpr int SWT_STAGE = SWT_A_BYTES + SWT_B_BYTES;
constexpr int SWT_SMEM = SWT_STAGES * SWT_STAGE + 1024 + 512;
constexpr int SWT_GROUPS = 4;                          // builder groups of 128 threads; group g takes the k-blocks i = g (mod 4)
constexpr int SWT_THREADS = SWT_GROUPS * 128 + 64;    // + MMA issuer warp + TMA producer warp
constexpr int SWT_MMA_WARP = SWT_GROUPS * 4, SWT_TMA_WARP = SWT_GROUPS * 4 + 1;

struct StemWgTcParams {
  const float* x;
  int n, cin, H, W, kh, kw, pad, K;
  long long total_px, kblocks, kb_per_cta;
  float* partial;  // [gridDim.x][32][64]
};

__global__ void __launch_bounds__(SWT_THREADS, 1)
stem_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmDy, StemWgTcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + SWT_STAGES * SWT_STAGE;
  auto full_bar = [&](int s) { return bars + 8u * s; };                       // TMA landed dy
  auto built_bar = [&](int s) { return bars + 8u * (SWT_STAGES + s); };       // patches written (128 arrivals)
  auto empty_bar = [&](int s) { return bars + 8u * (2 * SWT_STAGES + s); };   // MMAs of the stage retired
  const uint32_t done_bar = bars + 8u * (3 * SWT_STAGES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + SWT_STAGES * SWT_STAGE + (3 * SWT_STAGES + 1) * 8);
  int* tapinfo = reinterpret_cast<int*>(base_ptr + SWT_STAGES * SWT_STAGE + 256);  // [32] packed (ci << 16 | a << 8 | b)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SWT_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(built_bar(s), 128);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    const int t = threadIdx.x, khw = P.kh * P.kw;
    const int ci = t / khw, r = t - ci * khw, a = r / P.kw, b = r - a * P.kw;
    tapinfo[t] = (t < P.K) ? ((ci << 16) | (a << 8) | b) : -1;
  }
  if (warp == SWT_TMA_WARP && lane == 0) tma_prefetch_desc(&tmDy);
  if (warp == SWT_MMA_WARP) tmem_alloc(smem_u32(tmem_slot), 64);
  if (warp < 4) {  // the tap rows >= K (and everything else of the A slots) are zero for the life of the CTA
    uint4* z = reinterpret_cast<uint4*>(base_ptr);
    for (int s = 0; s < SWT_STAGES; ++s)
      for (int i = threadIdx.x; i < SWT_A_BYTES / 16; i += 128) z[s * (SWT_STAGE / 16) + i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();  // (the rows >= K are never written again: this fence is the one that publishes them to the tensor core)
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  long long kb0 = (long long)blockIdx.x * P.kb_per_cta;
  long long kb1 = kb0 + P.kb_per_cta;
  if (kb1 > P.kblocks) kb1 = P.kblocks;
  const int nk = kb1 > kb0 ? (int)(kb1 - kb0) : 0;

  if (warp == SWT_TMA_WARP) {
    // =============================== TMA producer: dy ========================================
    for (int i = 0; i < nk; ++i) {
      const int s = i % SWT_STAGES;
      mbar_wait(empty_bar(s), ((i / SWT_STAGES) & 1u) ^ 1u);
      if (lane == 0) {
        mbar_expect_tx(full_bar(s), SWT_B_BYTES);
        tma_load_2d(base + s * SWT_STAGE + SWT_A_BYTES, &tmDy, full_bar(s), 0, (int)((kb0 + i) * 64));
      }
      __syncwarp();
    }
  } else if (warp == SWT_MMA_WARP) {
    // =============================== MMA issuer ==============================================
    constexpr uint32_t idesc = make_idesc(64, 0, 1);  // A K-major, B MN-major, M = 128, N = 64
    const uint64_t adesc0 = make_desc(base, 16, 1024);
    const uint64_t bdesc0 = make_desc(base + SWT_A_BYTES, 8192, 1024);
    uint32_t s = 0, ph = 0;
    for (int i = 0; i < nk; ++i) {
      mbar_wait(full_bar(s), ph);
      mbar_wait(built_bar(s), ph);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t ad0 = adesc0 + (uint64_t)((s * (uint32_t)SWT_STAGE) >> 4);
        const uint64_t bd0 = bdesc0 + (uint64_t)((s * (uint32_t)SWT_STAGE) >> 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)  // K-major: 32 B per 16 pixels inside the swizzle row; MN-major: 16 pixel rows x 128 B
          umma_bf16(tmem_acc, ad0 + (uint64_t)((kk * 32) >> 4), bd0 + (uint64_t)((kk * 2048) >> 4), idesc, (i > 0 || kk > 0) ? 1u : 0u);
        umma_commit(empty_bar(s));
      }
      __syncwarp();
      if (++s == (uint32_t)SWT_STAGES) {
        s = 0;
        ph ^= 1u;
      }
    }
    if (elect_one_sync()) umma_commit(done_bar);
    __syncwarp();
  } else {
    // =============================== patch builders (4 groups of 4 warps) =====================
    // The gather is a chain of DRAM-latency loads (x does not stay in L2 beside the streamed dy): one group alone paced
    // the kernel at ~2400 clocks per k-block (1.85 ms, profiles/r02_session_g.md).  Four groups work on four consecutive
    // k-blocks at once, each with its next k-block's loads already in flight.
    const int grp = threadIdx.x >> 7, t = threadIdx.x & 127;
    const int p = t & 63, th = t >> 6;  // pixel of the k-block, tap parity
    const long long HW = (long long)P.H * P.W;
    // position of this thread's pixel in the group's first k-block, advanced by SWT_GROUPS * 64 pixels per round
    long long q = (kb0 + grp) * 64 + p;
    int img = (int)(q / HW);
    int y = (int)((q - (long long)img * HW) / P.W);
    int x0 = (int)(q - (long long)img * HW - (long long)y * P.W);
    auto gather = [&](float (&v)[16]) {
      const bool live = q < P.total_px;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int info = tapinfo[2 * j + th];
        float val = 0.f;
        if (live && info >= 0) {
          const int ci = info >> 16, a = (info >> 8) & 255, b = info & 255;
          const int yy = y + a - P.pad, xx = x0 + b - P.pad;
          if (yy >= 0 && yy < P.H && xx >= 0 && xx < P.W) val = __ldg(P.x + (((long long)img * P.cin + ci) * P.H + yy) * P.W + xx);
        }
        v[j] = val;
      }
      q += SWT_GROUPS * 64;
      x0 += SWT_GROUPS * 64;
      while (x0 >= P.W) {
        x0 -= P.W;
        if (++y == P.H) {
          y = 0;
          ++img;
        }
      }
    };
    float v[16], vn[16];
    if (grp < nk) gather(v);
    for (int i = grp; i < nk; i += SWT_GROUPS) {
      const int s = i % SWT_STAGES;
      if (i + SWT_GROUPS < nk) gather(vn);  // the group's next k-block is in flight while this one is written
      mbar_wait(empty_bar(s), ((i / SWT_STAGES) & 1u) ^ 1u);
      uint8_t* a_slot = base_ptr + s * SWT_STAGE;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int tap = 2 * j + th;
        if (tap < P.K)
          *reinterpret_cast<__nv_bfloat16*>(a_slot + tap * 128 + ((((uint32_t)p >> 3) ^ ((uint32_t)tap & 7u)) << 4) + (p & 7) * 2) =
              __float2bfloat16(v[j]);
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      mbar_arrive(built_bar(s));
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = vn[j];
    }
    if (warp == 0) {
      // =============================== epilogue: lanes 0..31 of the accumulator = taps ========
      float* dst = P.partial + ((long long)blockIdx.x * 32 + lane) * 64;
      if (nk > 0) {
        mbar_wait(done_bar, 0);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_acc + (uint32_t)(c * 32), r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(dst + c * 32 + j) = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                                                       __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        }
      } else {
#pragma unroll
        for (int j = 0; j < 64; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == SWT_MMA_WARP) tmem_dealloc(tmem_acc, 64);
}

// grad[co][tap] (torch OIHW, flat co * K + tap) (+)= sum over CTAs of partial[cta][tap][co], in CTA order
__global__ void stem_wgrad_tc_fold_kernel(const float* __restrict__ partial, float* __restrict__ grad, int K, int cout, int nblk,
                                          int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // tap * 64 + co
  if (i >= K * 64) return;
  const int tap = i >> 6, co = i & 63;
  if (co >= cout) return;
  float t = 0.f;
  for (int b = 0; b < nblk; ++b) t += partial[(long long)b * 2048 + i];
  float* g = grad + (long long)co * K + tap;
  *g = accumulate ? *g + t : t;
}

typedef CUresult (*PFN_swtEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_swtEncodeTiled swt_encodeTiled = nullptr;

static int swt_init() {
  if (swt_encodeTiled) return IIC_OK;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  IIC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  IIC_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, IIC_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  swt_encodeTiled = (PFN_swtEncodeTiled)fn;
  return IIC_OK;
}

static int swt_grid(const iic_conv_geom* g) {
  const long long kblocks = ((long long)g->n * g->h * g->w + 63) / 64;
  long long grid = device_sm_count();
  if (grid > kblocks) grid = kblocks;
  return (int)(grid < 1 ? 1 : grid);
}

}  // namespace iic

using namespace iic;

extern "C" long long iic_stem_wgrad_tc_workspace(const iic_conv_geom* g) {
  if (g == nullptr) return 0;
  return (long long)swt_grid(g) * 32 * 64 * (long long)sizeof(float);
}

extern "C" int iic_stem_wgrad_tc(const float* x_nchw, const void* dy_bf16, float* grad_oihw, int accumulate, void* workspace,
                                 const iic_conv_geom* g, void* stream) {
  IIC_REQUIRE(x_nchw && dy_bf16 && grad_oihw && workspace && g, IIC_ERR_BAD_ARG, "iic_stem_wgrad_tc: null pointer");
  const int K = g->cin * g->kh * g->kw;
  IIC_REQUIRE(K >= 1 && K <= 32 && g->cout == 64 && g->stride == 1 && g->dil == 1 && g->oh == g->h && g->ow == g->w &&
                  g->kh <= 255 && g->kw <= 255,
              IIC_ERR_UNSUPPORTED, "iic_stem_wgrad_tc: needs cin*kh*kw <= 32, cout = 64, stride 1, dilation 1, 'same' padding");
  const long long total_px = (long long)g->n * g->h * g->w;
  IIC_REQUIRE(total_px + 64 < (1ll << 31), IIC_ERR_UNSUPPORTED, "iic_stem_wgrad_tc: more than 2^31 pixels");
  int rc = swt_init();
  if (rc != IIC_OK) return rc;
  alignas(64) CUtensorMap tm;
  {
    cuuint64_t gdim[2] = {64, (cuuint64_t)total_px};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = swt_encodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(dy_bf16), gdim, gstr, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IIC_REQUIRE(r == CUDA_SUCCESS, IIC_ERR_CUDA, "cuTensorMapEncodeTiled(dy) failed (%d)", (int)r);
  }
  StemWgTcParams P = {};
  P.x = x_nchw;
  P.n = g->n; P.cin = g->cin; P.H = g->h; P.W = g->w; P.kh = g->kh; P.kw = g->kw; P.pad = g->pad; P.K = K;
  P.total_px = total_px;
  P.kblocks = (total_px + 63) / 64;
  const int grid = swt_grid(g);
  P.kb_per_cta = (P.kblocks + grid - 1) / grid;
  P.partial = (float*)workspace;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr = false;
  if (!attr) {
    IIC_CUDA(cudaFuncSetAttribute(stem_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SWT_SMEM));
    attr = true;
  }
  stem_wgrad_tc_kernel<<<grid, SWT_THREADS, SWT_SMEM, st>>>(tm, P);
  IIC_LAUNCH_CHECK();
  count_launch();
  stem_wgrad_tc_fold_kernel<<<cdiv(K * 64, 256), 256, 0, st>>>((const float*)workspace, grad_oihw, K, g->cout, grid, accumulate);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}
