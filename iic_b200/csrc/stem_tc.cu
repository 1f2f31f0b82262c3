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
  float* partial;  // [gridDim.x][KMAX][64]
};

// KMAX = 32 | 64: taps handled per pixel (two per builder thread and iteration); the accumulator rows 0 .. KMAX-1 are read back
template <int KMAX>
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
  int* tapinfo = reinterpret_cast<int*>(base_ptr + SWT_STAGES * SWT_STAGE + 256);  // [KMAX] packed (ci << 16 | a << 8 | b)

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
  if (threadIdx.x < KMAX) {
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
    constexpr int NT = KMAX / 2;  // taps per builder thread
    auto gather = [&](float (&v)[NT]) {
      const bool live = q < P.total_px;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
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
    float v[NT], vn[NT];
    if (grp < nk) gather(v);
    for (int i = grp; i < nk; i += SWT_GROUPS) {
      const int s = i % SWT_STAGES;
      if (i + SWT_GROUPS < nk) gather(vn);  // the group's next k-block is in flight while this one is written
      mbar_wait(empty_bar(s), ((i / SWT_STAGES) & 1u) ^ 1u);
      uint8_t* a_slot = base_ptr + s * SWT_STAGE;
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int tap = 2 * j + th;
        if (tap < P.K)
          *reinterpret_cast<__nv_bfloat16*>(a_slot + tap * 128 + ((((uint32_t)p >> 3) ^ ((uint32_t)tap & 7u)) << 4) + (p & 7) * 2) =
              __float2bfloat16(v[j]);
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
      mbar_arrive(built_bar(s));
#pragma unroll
      for (int j = 0; j < NT; ++j) v[j] = vn[j];
    }
    if (warp < KMAX / 32) {
      // =============================== epilogue: lanes 0 .. KMAX-1 of the accumulator = taps ==
      // (warp w of the CTA may read TMEM lanes 32 (w % 4) .. +31: warps 0 and 1 of builder group 0)
      float* dst = P.partial + ((long long)blockIdx.x * KMAX + warp * 32 + lane) * 64;
      if (nk > 0) {
        mbar_wait(done_bar, 0);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld32(tmem_acc + (uint32_t)(c * 32) + ((uint32_t)(warp * 32) << 16), r);
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
                                          int accumulate, int kmax) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // tap * 64 + co
  if (i >= K * 64) return;
  const int tap = i >> 6, co = i & 63;
  if (co >= cout) return;
  float t = 0.f;
  for (int b = 0; b < nblk; ++b) t += partial[(long long)b * kmax * 64 + i];
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
  return (long long)swt_grid(g) * 64 * 64 * (long long)sizeof(float);
}

extern "C" int iic_stem_wgrad_tc(const float* x_nchw, const void* dy_bf16, float* grad_oihw, int accumulate, void* workspace,
                                 const iic_conv_geom* g, void* stream) {
  IIC_REQUIRE(x_nchw && dy_bf16 && grad_oihw && workspace && g, IIC_ERR_BAD_ARG, "iic_stem_wgrad_tc: null pointer");
  const int K = g->cin * g->kh * g->kw;
  IIC_REQUIRE(K >= 1 && K <= 64 && g->cout == 64 && g->stride == 1 && g->dil == 1 && g->oh == g->h && g->ow == g->w &&
                  g->kh <= 255 && g->kw <= 255,
              IIC_ERR_UNSUPPORTED, "iic_stem_wgrad_tc: needs cin*kh*kw <= 64, cout = 64, stride 1, dilation 1, 'same' padding");
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
    IIC_CUDA(cudaFuncSetAttribute(stem_wgrad_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWT_SMEM));
    IIC_CUDA(cudaFuncSetAttribute(stem_wgrad_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, SWT_SMEM));
    attr = true;
  }
  const int kmax = K <= 32 ? 32 : 64;
  if (kmax == 32) stem_wgrad_tc_kernel<32><<<grid, SWT_THREADS, SWT_SMEM, st>>>(tm, P);
  else stem_wgrad_tc_kernel<64><<<grid, SWT_THREADS, SWT_SMEM, st>>>(tm, P);
  IIC_LAUNCH_CHECK();
  count_launch();
  stem_wgrad_tc_fold_kernel<<<cdiv(K * 64, 256), 256, 0, st>>>((const float*)workspace, grad_oihw, K, g->cout, grid, accumulate, kmax);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

namespace iic {

// =====================================================================================================================
// Stem convolution (fprop) on tcgen05 with the BatchNorm statistics of its fp32 results, bf16 output.
//
// y[pixel][cout] = sum over taps of patch(x)[pixel][tap] * w[cout][tap]: M = 128 pixels per tile, N = 64, K = 32 (taps,
// zero padded).  The SIMT kernel (stem.cu: stem_fprop64q_kernel) is bound by its FMAs and the conversions around them
// (1.15 ms at the c4 shape for a 1.66 GB output = 0.26 ms of HBM time).  Here
//   * two builder groups of 128 threads gather the patches of a tile (thread = pixel, all taps; next tile's loads in flight)
//     and write them as the K-major SWIZZLE_128B operand (64 of the 128 bytes of a row are used);
//   * the weights are converted once per CTA into the resident B operand;
//   * one elected thread issues the two K = 16 MMAs of a tile into one of two TMEM accumulators;
//   * eight epilogue warps (TMEM lane quadrant x column half) read the accumulator, keep per-lane running sums and sums of
//     squares of their 32 columns (reduced across lanes once per CTA and view), convert to bf16 into a swizzled staging
//     tile and one TMA store per tile writes 128 pixels x 128 B.
// Rows past the last pixel hold zero patches: they add nothing to the statistics and the TMA store clips them.
constexpr int SFT_NST = 4;                       // A stages
constexpr int SFT_A_BYTES = 128 * 128;
constexpr int SFT_B_BYTES = 64 * 128;
constexpr int SFT_OUT_BYTES = 128 * 128;         // one staged output tile
constexpr int SFT_GROUPS = 2;
constexpr int SFT_EPI_THREADS = 256;
constexpr int SFT_THREADS = SFT_EPI_THREADS + SFT_GROUPS * 128 + 32;
constexpr int SFT_MMA_WARP = (SFT_EPI_THREADS + SFT_GROUPS * 128) / 32;
constexpr int SFT_SMEM = SFT_B_BYTES + SFT_NST * SFT_A_BYTES + 2 * SFT_OUT_BYTES + 4096 + 1024;

struct StemFpTcParams {
  const float* x;
  const float* w;  // [64][cin][kh][kw]
  int n, cin, H, W, kh, kw, pad, K;
  long long total_px, view_px;  // pixels of one view (== total_px for a single view)
  long long tiles, tiles_per_cta;
  float* stat_partial;          // [gridDim.x][2 views][{sum, sum of squares}][64]
};

__device__ __forceinline__ void swt_tma_store_2d(const void* desc, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void swt_bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
// on return t[0] of lane l = sum over the 32 lanes of (their) t[l]  (31 shuffles)
__device__ __forceinline__ void swt_col_reduce(float (&t)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? t[i] : t[i + off];
      const float keep = up ? t[i + off] : t[i];
      t[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
}

__global__ void __launch_bounds__(SFT_THREADS, 1)
stem_fprop_tc_kernel(const __grid_constant__ CUtensorMap tmY, StemFpTcParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t bsm = (raw + 1023u) & ~1023u;                 // resident weights
  const uint32_t abase = bsm + SFT_B_BYTES;                     // A stages
  const uint32_t obase = abase + SFT_NST * SFT_A_BYTES;         // two staged output tiles
  const uint32_t misc = obase + 2 * SFT_OUT_BYTES;              // barriers, tap table, statistics scratch
  uint8_t* bsm_ptr = smem_raw + (bsm - raw);
  uint8_t* a_ptr = smem_raw + (abase - raw);
  uint8_t* o_ptr = smem_raw + (obase - raw);
  uint8_t* misc_ptr = smem_raw + (misc - raw);
  auto built_bar = [&](int s) { return misc + 8u * s; };
  auto aempty_bar = [&](int s) { return misc + 8u * (SFT_NST + s); };
  auto tfull_bar = [&](int a) { return misc + 8u * (2 * SFT_NST + a); };
  auto tempty_bar = [&](int a) { return misc + 8u * (2 * SFT_NST + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc_ptr + 8 * (2 * SFT_NST + 4));
  int* tapinfo = reinterpret_cast<int*>(misc_ptr + 128);        // [32]
  float* scratch = reinterpret_cast<float*>(misc_ptr + 256);    // [8 warps][{sum, sq}][32]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SFT_NST; ++s) {
      mbar_init(built_bar(s), 128);
      mbar_init(aempty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), SFT_EPI_THREADS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    tma_prefetch_desc(&tmY);
  }
  if (threadIdx.x < 32) {
    const int t = threadIdx.x, khw = P.kh * P.kw;
    const int ci = t / khw, r = t - ci * khw, a = r / P.kw, b = r - a * P.kw;
    tapinfo[t] = (t < P.K) ? ((ci << 16) | (a << 8) | b) : -1;
  }
  // resident B operand: row = cout, K-major, 16-byte chunk (tap >> 3) at its SWIZZLE_128B position; taps >= K are zero
  for (int i = threadIdx.x; i < 64 * 4; i += SFT_THREADS) {
    const int co = i >> 2, ch = i & 3;
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int t0 = ch * 8 + 2 * e;
      const float f0 = t0 < P.K ? P.w[co * P.K + t0] : 0.f, f1 = (t0 + 1) < P.K ? P.w[co * P.K + t0 + 1] : 0.f;
      const __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
      pk[e] = *reinterpret_cast<const uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(bsm_ptr + co * 128 + ((ch ^ (co & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  fence_proxy_async();
  if (warp == SFT_MMA_WARP) tmem_alloc(smem_u32(tmem_slot), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  long long t0 = (long long)blockIdx.x * P.tiles_per_cta;
  long long t1 = t0 + P.tiles_per_cta;
  if (t1 > P.tiles) t1 = P.tiles;
  const int nt = t1 > t0 ? (int)(t1 - t0) : 0;

  if (warp == SFT_MMA_WARP) {
    // =============================== MMA issuer ==============================================
    constexpr uint32_t idesc = make_idesc(64, 0, 0);
    const uint64_t adesc0 = make_desc(abase, 16, 1024);
    const uint64_t bdesc0 = make_desc(bsm, 16, 1024);
    const bool two = P.K > 16;
    for (int i = 0; i < nt; ++i) {
      const uint32_t s = (uint32_t)i % SFT_NST, as = (uint32_t)i & 1u;
      mbar_wait(tempty_bar(as), (((uint32_t)i >> 1) & 1u) ^ 1u);
      mbar_wait(built_bar(s), ((uint32_t)i / SFT_NST) & 1u);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t ad = adesc0 + (uint64_t)((s * (uint32_t)SFT_A_BYTES) >> 4);
        umma_bf16(tmem_base + as * 64u, ad, bdesc0, idesc, 0u);
        if (two) umma_bf16(tmem_base + as * 64u, ad + 2u, bdesc0 + 2u, idesc, 1u);  // + 32 B: taps 16..31
        umma_commit(aempty_bar(s));
        umma_commit(tfull_bar(as));
      }
      __syncwarp();
    }
  } else if (warp >= SFT_EPI_THREADS / 32) {
    // =============================== patch builders (2 groups x 4 warps) =====================
    const int bt = threadIdx.x - SFT_EPI_THREADS;
    const int grp = bt >> 7, t = bt & 127;  // t = pixel row of the tile
    const long long HW = (long long)P.H * P.W;
    long long q = (t0 + grp) * 128 + t;
    int img = (int)(q / HW);
    int y = (int)((q - (long long)img * HW) / P.W);
    int x0 = (int)(q - (long long)img * HW - (long long)y * P.W);
    auto gather = [&](float (&v)[32]) {
      const bool live = q < P.total_px;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int info = tapinfo[j];
        float val = 0.f;
        if (live && info >= 0) {
          const int ci = info >> 16, a = (info >> 8) & 255, b = info & 255;
          const int yy = y + a - P.pad, xx = x0 + b - P.pad;
          if (yy >= 0 && yy < P.H && xx >= 0 && xx < P.W) val = __ldg(P.x + (((long long)img * P.cin + ci) * P.H + yy) * P.W + xx);
        }
        v[j] = val;
      }
      q += SFT_GROUPS * 128;
      x0 += SFT_GROUPS * 128;
      while (x0 >= P.W) {
        x0 -= P.W;
        if (++y == P.H) {
          y = 0;
          ++img;
        }
      }
    };
    float v[32], vn[32];
    if (grp < nt) gather(v);
    for (int i = grp; i < nt; i += SFT_GROUPS) {
      const uint32_t s = (uint32_t)i % SFT_NST;
      if (i + SFT_GROUPS < nt) gather(vn);
      mbar_wait(aempty_bar(s), (((uint32_t)i / SFT_NST) & 1u) ^ 1u);
      uint8_t* row = a_ptr + s * SFT_A_BYTES + t * 128;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(v[ch * 8 + 2 * e], v[ch * 8 + 2 * e + 1]);
          pk[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(row + ((ch ^ (t & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async();
      mbar_arrive(built_bar(s));
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = vn[j];
    }
  } else {
    // =============================== epilogue (warps 0-7) ====================================
    const int quad = warp & 3, hsel = warp >> 2;
    const int m = quad * 32 + lane;  // accumulator row = pixel of the tile
    float s1[32], s2[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) s1[e] = s2[e] = 0.f;
    int cur_view = 0;
    int wrote = 0;  // bit v: view v's row of the partial has been written
    auto flush = [&](int view) {
      // per-column totals over this warp's 32 rows, then over the four lane quadrants (fixed order), into the CTA's row
      swt_col_reduce(s1, lane);
      swt_col_reduce(s2, lane);
      scratch[(warp * 2 + 0) * 32 + lane] = s1[0];
      scratch[(warp * 2 + 1) * 32 + lane] = s2[0];
      named_bar_sync(2, SFT_EPI_THREADS);
      if (threadIdx.x < 128) {
        const int qd = threadIdx.x >> 6, col = threadIdx.x & 63, h = col >> 5, l = col & 31;
        float tsum = 0.f;
#pragma unroll
        for (int w4 = 0; w4 < 4; ++w4) tsum += scratch[((h * 4 + w4) * 2 + qd) * 32 + l];
        P.stat_partial[((long long)blockIdx.x * 2 + view) * 128 + qd * 64 + col] = tsum;
      }
      named_bar_sync(2, SFT_EPI_THREADS);
#pragma unroll
      for (int e = 0; e < 32; ++e) s1[e] = s2[e] = 0.f;
    };
    for (int i = 0; i < nt; ++i) {
      const uint32_t as = (uint32_t)i & 1u;
      const long long px0 = (t0 + i) * 128;
      const int view = px0 >= P.view_px ? 1 : 0;
      if (view != cur_view) {
        flush(cur_view);
        wrote |= 1 << cur_view;
        cur_view = view;
      }
      // the TMA store issued two tiles ago must have read this staging buffer
      if (threadIdx.x == 0) swt_bulk_wait_read1();
      named_bar_sync(1, SFT_EPI_THREADS);
      mbar_wait(tfull_bar(as), ((uint32_t)i >> 1) & 1u);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld32(tmem_base + as * 64u + (uint32_t)(hsel * 32) + ((uint32_t)(quad * 32) << 16), r);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(tempty_bar(as));
      uint8_t* orow = o_ptr + as * SFT_OUT_BYTES + m * 128;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float f0 = __uint_as_float(r[ch * 8 + 2 * e]), f1 = __uint_as_float(r[ch * 8 + 2 * e + 1]);
          s1[ch * 8 + 2 * e] += f0;
          s2[ch * 8 + 2 * e] += f0 * f0;
          s1[ch * 8 + 2 * e + 1] += f1;
          s2[ch * 8 + 2 * e + 1] += f1 * f1;
          const __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
          pk[e] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(orow + (((hsel * 4 + ch) ^ (m & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      fence_proxy_async();
      named_bar_sync(1, SFT_EPI_THREADS);
      if (threadIdx.x == 0) {
        swt_tma_store_2d(&tmY, obase + as * SFT_OUT_BYTES, 0, (int)px0);
        bulk_commit();
      }
    }
    flush(cur_view);
    wrote |= 1 << cur_view;
    if (threadIdx.x < 128) {
#pragma unroll
      for (int vw = 0; vw < 2; ++vw)
        if (!((wrote >> vw) & 1)) P.stat_partial[((long long)blockIdx.x * 2 + vw) * 128 + threadIdx.x] = 0.f;
    }
    if (threadIdx.x == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == SFT_MMA_WARP) tmem_dealloc(tmem_base, 128);
}

static int sft_grid(const iic_conv_geom* g) {
  const long long tiles = ((long long)g->n * g->h * g->w + 127) / 128;
  long long grid = device_sm_count();
  if (grid > tiles) grid = tiles;
  return (int)(grid < 1 ? 1 : grid);
}

static bool sft_ok(const iic_conv_geom* g, int views) {
  if (g == nullptr) return false;
  const int K = g->cin * g->kh * g->kw;
  const long long total = (long long)g->n * g->h * g->w;
  if (!(K >= 1 && K <= 32 && g->cout == 64 && g->stride == 1 && g->dil == 1 && g->oh == g->h && g->ow == g->w && g->kh <= 255 &&
        g->kw <= 255 && total + 128 < (1ll << 31)))
    return false;
  if (views == 1) return true;
  return views == 2 && g->n % 2 == 0 && (total / 2) % 128 == 0;  // a tile never straddles the two views
}

}  // namespace iic

extern "C" int iic_stem_fprop_stats_tc_blocks(const iic_conv_geom* g, int views) { return sft_ok(g, views) ? sft_grid(g) : 0; }

extern "C" int iic_stem_fprop_stats_tc(const float* x_nchw, const float* w_oihw, void* y_bf16, const iic_conv_geom* g, int views,
                                       float* stat_partial, void* stream) {
  IIC_REQUIRE(x_nchw && w_oihw && y_bf16 && stat_partial && g, IIC_ERR_BAD_ARG, "iic_stem_fprop_stats_tc: null pointer");
  IIC_REQUIRE(sft_ok(g, views), IIC_ERR_UNSUPPORTED,
              "iic_stem_fprop_stats_tc: needs cin*kh*kw <= 32, cout = 64, stride 1, 'same' padding, 1 view or 2 views of a "
              "multiple of 128 pixels");
  int rc = swt_init();
  if (rc != IIC_OK) return rc;
  const long long total_px = (long long)g->n * g->h * g->w;
  alignas(64) CUtensorMap tm;
  {
    cuuint64_t gdim[2] = {64, (cuuint64_t)total_px};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = swt_encodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, y_bf16, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IIC_REQUIRE(r == CUDA_SUCCESS, IIC_ERR_CUDA, "cuTensorMapEncodeTiled(y) failed (%d)", (int)r);
  }
  StemFpTcParams P = {};
  P.x = x_nchw; P.w = w_oihw;
  P.n = g->n; P.cin = g->cin; P.H = g->h; P.W = g->w; P.kh = g->kh; P.kw = g->kw; P.pad = g->pad;
  P.K = g->cin * g->kh * g->kw;
  P.total_px = total_px;
  P.view_px = views == 2 ? total_px / 2 : total_px;
  P.tiles = (total_px + 127) / 128;
  const int grid = sft_grid(g);
  P.tiles_per_cta = (P.tiles + grid - 1) / grid;
  P.stat_partial = stat_partial;
  static bool attr = false;
  if (!attr) {
    IIC_CUDA(cudaFuncSetAttribute(stem_fprop_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SFT_SMEM));
    attr = true;
  }
  stem_fprop_tc_kernel<<<grid, SFT_THREADS, SFT_SMEM, (cudaStream_t)stream>>>(tm, P);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}
