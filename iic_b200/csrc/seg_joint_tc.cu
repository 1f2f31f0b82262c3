// Uncollapsed segmentation joint on the tensor cores -- SURVEY S8 a11:
//
//   A[u][v][c][c'] = sum_{n,y,x} x1m[n, y+u-T, x+v-T, c] * x2m[n, y, x, c']          (k <= 16 channels, V = 2T+1 <= 24)
//
// the reference's F.conv2d(x1^T, weight=x2^T, padding=T) (code/utils/segmentation/IID_losses.py:125), 390 GFLOP per
// sub-head at the COCO-Stuff-3 shape.  The fp32 SIMT kernel (seg_loss.cu::seg_joint_kernel) runs it at ~23 TFLOP/s.
//
// Formulation.  Both views are pixel-major with 16 channels per pixel.  For one image row y of x2 and one displacement
// row u, with x1row = row y+u-T of x1 zero-padded by T pixels on the left:
//
//     D[m = (v, c)][c'] += sum_x  x1row[x + v][c] * x2row[x][c']            M = 8 displacements x 16 channels = 128
//
// i.e. a GEMM whose A operand is the Toeplitz "displaced copies" matrix of x1row.  In pixel-major memory that operand
// needs no copies at all: element (m = v*16 + c, k = x) lives at pixel (x + v), channel c -- an MN-major UMMA layout whose
// K rows are one pixel apart and whose M atoms (the 16 channels of a pixel) are ALSO one pixel apart (LBO = one pixel:
// atom v+1 of pixel x is atom v of pixel x+1, the atoms overlap the K rows).  So one TMA box per x1 row feeds all 24
// displacements, and the row is reused for the 7 values of u a CTA owns as y advances (ring of 8 row slots): x1 and x2
// cross L2 -> SM once per CTA instead of (2T+1) times.  B (x2row) is MN-major too: [pixel][16 channels].  N = 16, fp32
// accumulators in TMEM (7 u x 3 M tiles x 16 columns = 336 columns).
// Operand type.  kind::tf32 cannot read MN-major operands in the ordinary swizzle modes (measured: zeros,
// tools/umma_sw64_probe.cu mode 0), kind::f16 can (the bf16 wgrad of conv_tc2.cu is MN-major).  The forward joint
// therefore runs on bf16 pixels (32 B per pixel, SWIZZLE_32B, K = 16 pixels per MMA) with every operand split into
// THREE-term precision x = hi + mid (+ 2^-17 x): hi = bf16(x), mid = bf16(x - hi), products hi*hi + hi*mid + mid*hi
// (relative error 2^-16 per product, unbiased; every joint entry is a sum of >= 10^4 non-negative products, so the sums
// are fp32-grade -- the loss tolerance is 2e-5).  The backward contractions below use kind::tf32 with K-major operands.
//
// Work item = (image n, chunk of rows y, group of 7 displacement rows u); partial results per CTA, fixed-order reduce.
#include <cuda.h>

#include "tc_ptx.cuh"

namespace iic {

constexpr int SJ_U = 7;        // displacement rows per CTA
constexpr int SJ_MT = 3;       // M tiles of 8 displacements: v < 24
constexpr int SJ_SLOTS = 8;    // x1 row ring (U + 1)
constexpr int SJ_ST2 = 2;      // x2 row stages
constexpr int SJ_THREADS = 192;

struct SjParams {
  int n, h, w, wp, T, V;
  int ychunk, nychunks, ugroups;
  int row1_bytes, row2_bytes;  // one precision part of an x1 / x2 row buffer in shared memory (1024-byte multiples)
  int box1_bytes, box2_bytes;  // bytes one TMA box delivers ((wp + 24) * 64 and wp * 64)
  float* part;                 // [cta][SJ_U][24][16][16]
};

constexpr int SJ_PIXB = 32;  // bytes per pixel: 16 bf16 channels
__device__ __forceinline__ uint64_t sj_desc(uint32_t saddr) {
  // MN-major, SWIZZLE_32B, LBO = 32 B (next 16-channel atom = next pixel), SBO = 256 B (8 K rows = 8 pixels)
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)(SJ_PIXB >> 4) << 16;
  d |= (uint64_t)((8 * SJ_PIXB) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;
  return d;
}
__device__ __forceinline__ void sj_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void sj_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// x (fp32, pixel-major) -> hi = bf16(x), mid = bf16(x - hi)
__global__ void seg_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                 long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y), h2 = __float2bfloat16_rn(v.z),
                        h3 = __float2bfloat16_rn(v.w);
    __nv_bfloat162 ha = __halves2bfloat162(h0, h1), hb = __halves2bfloat162(h2, h3);
    __nv_bfloat162 la = __floats2bfloat162_rn(v.x - __bfloat162float(h0), v.y - __bfloat162float(h1));
    __nv_bfloat162 lb = __floats2bfloat162_rn(v.z - __bfloat162float(h2), v.w - __bfloat162float(h3));
    uint2 ho, lo2;
    ho.x = *reinterpret_cast<uint32_t*>(&ha); ho.y = *reinterpret_cast<uint32_t*>(&hb);
    lo2.x = *reinterpret_cast<uint32_t*>(&la); lo2.y = *reinterpret_cast<uint32_t*>(&lb);
    reinterpret_cast<uint2*>(hi)[i] = ho;
    reinterpret_cast<uint2*>(lo)[i] = lo2;
  }
}

// tensor maps: x1 hi / lo (box 16 x (wp+24) x 1, start x = -T), x2 hi / lo (box 16 x wp x 1); all views [n*h][w][16]
__global__ void __launch_bounds__(SJ_THREADS, 1)
seg_joint_tc_kernel(const __grid_constant__ CUtensorMap tm1h, const __grid_constant__ CUtensorMap tm1l,
                    const __grid_constant__ CUtensorMap tm2h, const __grid_constant__ CUtensorMap tm2l, SjParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t slot_bytes = 2u * (uint32_t)P.row1_bytes;   // [hi | lo]
  const uint32_t st2_bytes = 2u * (uint32_t)P.row2_bytes;
  const uint32_t x2base = base + SJ_SLOTS * slot_bytes;
  const uint32_t bars = x2base + SJ_ST2 * st2_bytes;
  auto full1 = [&](int s) { return bars + 8u * s; };
  auto empty1 = [&](int s) { return bars + 8u * (SJ_SLOTS + s); };
  auto full2 = [&](int s) { return bars + 8u * (2 * SJ_SLOTS + s); };
  auto empty2 = [&](int s) { return bars + 8u * (2 * SJ_SLOTS + SJ_ST2 + s); };
  const uint32_t done_bar = bars + 8u * (2 * SJ_SLOTS + 2 * SJ_ST2);
  uint8_t* bars_ptr = smem_raw + (bars - raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_ptr + 8 * (2 * SJ_SLOTS + 2 * SJ_ST2 + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SJ_SLOTS; ++s) {
      mbar_init(full1(s), 1);
      mbar_init(empty1(s), 1);
    }
    for (int s = 0; s < SJ_ST2; ++s) {
      mbar_init(full2(s), 1);
      mbar_init(empty2(s), 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm1h);
    tma_prefetch_desc(&tm1l);
    tma_prefetch_desc(&tm2h);
    tma_prefetch_desc(&tm2l);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work item
  const int g = blockIdx.x % P.ugroups;
  const int yc = (blockIdx.x / P.ugroups) % P.nychunks;
  const int img = blockIdx.x / (P.ugroups * P.nychunks);
  const int ya = yc * P.ychunk, yb = min(P.h, ya + P.ychunk);
  const int u0 = g * SJ_U;
  const int nu = min(SJ_U, P.V - u0);           // displacement rows this CTA really owns
  const int r_lo = max(0, ya + u0 - P.T);        // first / last valid x1 row this CTA ever touches
  const int r_hi = min(P.h - 1, (yb - 1) + (u0 + nu - 1) - P.T);

  if (warp == 0) {
    // =============================== TMA producer ============================================
    if (lane == 0) {
      int r_next = r_lo;
      for (int y = ya; y < yb; ++y) {
        const int need = min(r_hi, y + (u0 + nu - 1) - P.T);
        for (; r_next <= need; ++r_next) {
          const int idx = r_next - r_lo, s = idx % SJ_SLOTS;
          mbar_wait(empty1(s), ((idx / SJ_SLOTS) & 1u) ^ 1u);
          mbar_expect_tx(full1(s), 2u * (uint32_t)P.box1_bytes);
          const uint32_t dst = base + s * slot_bytes;
          tma_load_3d(dst, &tm1h, full1(s), 0, -P.T, img * P.h + r_next);
          tma_load_3d(dst + P.row1_bytes, &tm1l, full1(s), 0, -P.T, img * P.h + r_next);
        }
        const int it = y - ya, s2 = it % SJ_ST2;
        mbar_wait(empty2(s2), ((it / SJ_ST2) & 1u) ^ 1u);
        mbar_expect_tx(full2(s2), 2u * (uint32_t)P.box2_bytes);
        const uint32_t dst2 = x2base + s2 * st2_bytes;
        tma_load_3d(dst2, &tm2h, full2(s2), 0, 0, img * P.h + y);
        tma_load_3d(dst2 + P.row2_bytes, &tm2l, full2(s2), 0, 0, img * P.h + y);
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ==============================================
    // instruction descriptor: kind::f16, D = f32, A = B = bf16, M = 128, N = 16, A and B MN-major
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(16 >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
    uint32_t touched = 0u;
    const int nkk = P.wp / 16;  // K = 16 pixels per MMA
    for (int y = ya; y < yb; ++y) {
      const int it = y - ya, s2 = it % SJ_ST2;
      mbar_wait(full2(s2), (it / SJ_ST2) & 1u);
      const uint32_t b_hi = x2base + s2 * st2_bytes, b_lo = b_hi + P.row2_bytes;
      for (int ul = 0; ul < nu; ++ul) {
        const int r = y + u0 + ul - P.T;
        if (r < 0 || r >= P.h) continue;  // x1 row outside the image: contributes zeros
        const int idx = r - r_lo, s = idx % SJ_SLOTS;
        mbar_wait(full1(s), (idx / SJ_SLOTS) & 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a_hi = base + s * slot_bytes, a_lo = a_hi + P.row1_bytes;
#pragma unroll 1
          for (int mt = 0; mt < SJ_MT; ++mt) {
            const uint32_t acc = tmem_base + (uint32_t)((ul * SJ_MT + mt) * 16);
            const uint32_t bit = 1u << (ul * SJ_MT + mt);
            uint32_t first = (touched & bit) ? 1u : 0u;  // accumulate flag of the first MMA into this block
            for (int kk = 0; kk < nkk; ++kk) {
              const uint32_t aoff = (uint32_t)(kk * 16 + mt * 8) * SJ_PIXB, boff = (uint32_t)(kk * 16) * SJ_PIXB;
              sj_mma_bf16(acc, sj_desc(a_lo + aoff), sj_desc(b_hi + boff), idesc, first);  // mid_a * hi_b
              sj_mma_bf16(acc, sj_desc(a_hi + aoff), sj_desc(b_lo + boff), idesc, 1u);     // hi_a * mid_b
              sj_mma_bf16(acc, sj_desc(a_hi + aoff), sj_desc(b_hi + boff), idesc, 1u);     // hi_a * hi_b
              first = 1u;
            }
          }
          if (ul == 0) umma_commit(empty1(s));  // u = u0 is the last use of row r (larger y pairs it with smaller u)
        }
        __syncwarp();
        for (int mt = 0; mt < SJ_MT; ++mt) touched |= 1u << (ul * SJ_MT + mt);
      }
      // rows whose last use is this step but which were skipped above cannot exist (a valid row r is last used at
      // y = r - u0 + T, where it is paired with ul = 0 and is in range by construction of r_lo / r_hi) -- except when that
      // step lies beyond this CTA's chunk; those slots are never reused because the producer stops at r_hi.
      if (elect_one_sync()) umma_commit(empty2(s2));
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(done_bar);
    __syncwarp();
  } else {
    // =============================== epilogue (warps 2-5) ======================================
    mbar_wait(done_bar, 0);
    tc_fence_after();
    // accumulator blocks that received at least one MMA: displacement row ul met a valid x1 row inside this chunk
    uint32_t touched = 0u;
    for (int ul = 0; ul < nu; ++ul) {
      const int lo = max(ya, P.T - u0 - ul), hi = min(yb, P.h + P.T - u0 - ul);  // y with 0 <= y + u0 + ul - T < h
      if (lo < hi) touched |= 7u << (ul * SJ_MT);
    }
    const int quad = warp & 3;
    const int m = quad * 32 + lane;  // accumulator row = (v within the tile, c)
    const int vl = m >> 4, c = m & 15;
    float* out = P.part + (long long)blockIdx.x * SJ_U * 24 * 256;
    const int nblk = SJ_U * SJ_MT;
    for (int b0 = 0; b0 < nblk; b0 += 2) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(b0 * 16), v);
      tmem_ld_wait();
#pragma unroll
      for (int hb = 0; hb < 2; ++hb) {
        const int blk = b0 + hb;
        if (blk >= nblk) break;
        const int ul = blk / SJ_MT, mt = blk % SJ_MT;
        const bool live = (touched >> blk) & 1u;
        float* o = out + ((long long)(ul * 24 + mt * 8 + vl) * 16 + c) * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<float4*>(o + q * 4) =
              live ? make_float4(__uint_as_float(v[hb * 16 + q * 4]), __uint_as_float(v[hb * 16 + q * 4 + 1]),
                                 __uint_as_float(v[hb * 16 + q * 4 + 2]), __uint_as_float(v[hb * 16 + q * 4 + 3]))
                   : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// joint[(u*V + v)][c][c'] = sum over (image, row chunk) of the partial of group u / 7      (fixed order: deterministic)
__global__ void seg_joint_tc_reduce_kernel(const float* __restrict__ part, float* __restrict__ joint, int items, int ugroups,
                                           int V, int k) {
  const long long total = (long long)V * V * k * k;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cp = (int)(i % k);
    const int c = (int)((i / k) % k);
    const int v = (int)((i / ((long long)k * k)) % V);
    const int u = (int)(i / ((long long)k * k * V));
    const int g = u / SJ_U, ul = u % SJ_U;
    double t = 0.0;
    for (int it = 0; it < items; ++it)
      t += (double)part[(((long long)(it * ugroups + g) * SJ_U + ul) * 24 + v) * 256 + c * 16 + cp];
    joint[i] = (float)t;
  }
}

// ---- host ------------------------------------------------------------------------------------------
typedef CUresult (*PFN_sjEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_sjEncodeTiled sj_encodeTiled = nullptr;

static int sj_init() {
  if (sj_encodeTiled) return IIC_OK;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  IIC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  IIC_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, IIC_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  sj_encodeTiled = (PFN_sjEncodeTiled)fn;
  return IIC_OK;
}

// pixel-major rows [rows][w][16 channels]: fp32 (64 B per pixel, SWIZZLE_64B) or bf16 (32 B per pixel, SWIZZLE_32B)
static int sj_map(CUtensorMap* tm, const void* ptr, int rows, int w, int box_w, bool bf16 = false) {
  const cuuint64_t pixb = bf16 ? 32 : 64;
  cuuint64_t gdim[3] = {16, (cuuint64_t)w, (cuuint64_t)rows};
  cuuint64_t gstr[2] = {pixb, (cuuint64_t)w * pixb};
  cuuint32_t box[3] = {16, (cuuint32_t)box_w, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = sj_encodeTiled(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3,
                              const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              bf16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IIC_REQUIRE(r == CUDA_SUCCESS, IIC_ERR_CUDA, "cuTensorMapEncodeTiled(seg joint) failed (%d) rows=%d w=%d box_w=%d", (int)r, rows,
              w, box_w);
  return IIC_OK;
}

struct SjPlan {
  bool ok;
  int wp, ychunk, nychunks, ugroups, items, ctas, row1_bytes, row2_bytes, smem;
};

static SjPlan sj_plan(int n, int k, int h, int w, int T) {
  SjPlan p = {};
  const int V = 2 * T + 1;
  if (k > 16 || k < 5 || V > 24 || T < 1 || n < 1) return p;   // (k <= 4 / 8 stay on the SIMT kernel: 1/16 of the work)
  p.wp = (w + 15) / 16 * 16;                                    // K = 16 pixels per MMA
  if (p.wp + 24 > 256) return p;                                // TMA box extent
  p.row1_bytes = ((p.wp + 24) * SJ_PIXB + 1023) / 1024 * 1024;
  p.row2_bytes = (p.wp * SJ_PIXB + 1023) / 1024 * 1024;
  p.smem = 1024 + SJ_SLOTS * 2 * p.row1_bytes + SJ_ST2 * 2 * p.row2_bytes + 256;
  if (p.smem > 232448) return p;
  if (p.smem < 120 * 1024) p.smem = 120 * 1024;  // one CTA per SM: each allocates all 512 TMEM columns
  p.ugroups = (V + SJ_U - 1) / SJ_U;
  // row chunks: ~3 waves of CTAs over the SMs, at least T + 1 rows per chunk
  int want = (3 * device_sm_count() + n * p.ugroups - 1) / (n * p.ugroups);
  if (want < 1) want = 1;
  p.ychunk = (h + want - 1) / want;
  if (p.ychunk < T + 1) p.ychunk = T + 1;
  if (p.ychunk > h) p.ychunk = h;
  p.nychunks = (h + p.ychunk - 1) / p.ychunk;
  p.items = n * p.nychunks;
  p.ctas = p.items * p.ugroups;
  p.ok = true;
  return p;
}

// workspace: hi/lo copies of both views + per-CTA partials; 0 = geometry not supported by the tensor-core kernel
long long seg_joint_tc_workspace(int n, int k, int h, int w, int T) {
  const SjPlan p = sj_plan(n, k, h, w, T);
  if (!p.ok) return 0;
  return 4ll * n * h * w * 16 * (long long)sizeof(__nv_bfloat16) + (long long)p.ctas * SJ_U * 24 * 256 * (long long)sizeof(float);
}

int seg_joint_tc(const float* x1m, const float* x2m, float* joint, void* workspace, int n, int k, int h, int w, int T,
                 cudaStream_t st) {
  int rc = sj_init();
  if (rc != IIC_OK) return rc;
  const SjPlan p = sj_plan(n, k, h, w, T);
  IIC_REQUIRE(p.ok, IIC_ERR_UNSUPPORTED, "seg_joint_tc: unsupported geometry (k=%d, T=%d, w=%d)", k, T, w);
  const long long elems = (long long)n * h * w * 16;
  __nv_bfloat16* x1h = (__nv_bfloat16*)workspace;
  __nv_bfloat16* x1l = x1h + elems;
  __nv_bfloat16* x2h = x1l + elems;
  __nv_bfloat16* x2l = x2h + elems;
  float* part = (float*)(x2l + elems);
  int blocks = cdiv(elems / 4, 256);
  if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
  seg_split_kernel<<<blocks, 256, 0, st>>>(x1m, x1h, x1l, elems / 4);
  IIC_LAUNCH_CHECK();
  count_launch();
  seg_split_kernel<<<blocks, 256, 0, st>>>(x2m, x2h, x2l, elems / 4);
  IIC_LAUNCH_CHECK();
  count_launch();
  alignas(64) CUtensorMap tm1h, tm1l, tm2h, tm2l;
  if ((rc = sj_map(&tm1h, x1h, n * h, w, p.wp + 24, true)) != IIC_OK) return rc;
  if ((rc = sj_map(&tm1l, x1l, n * h, w, p.wp + 24, true)) != IIC_OK) return rc;
  if ((rc = sj_map(&tm2h, x2h, n * h, w, p.wp, true)) != IIC_OK) return rc;
  if ((rc = sj_map(&tm2l, x2l, n * h, w, p.wp, true)) != IIC_OK) return rc;
  SjParams P = {};
  P.n = n; P.h = h; P.w = w; P.wp = p.wp; P.T = T; P.V = 2 * T + 1;
  P.ychunk = p.ychunk; P.nychunks = p.nychunks; P.ugroups = p.ugroups;
  P.row1_bytes = p.row1_bytes; P.row2_bytes = p.row2_bytes;
  P.box1_bytes = (p.wp + 24) * SJ_PIXB; P.box2_bytes = p.wp * SJ_PIXB;
  P.part = part;
  IIC_CUDA(cudaFuncSetAttribute(seg_joint_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem));
  seg_joint_tc_kernel<<<p.ctas, SJ_THREADS, p.smem, st>>>(tm1h, tm1l, tm2h, tm2l, P);
  IIC_LAUNCH_CHECK();
  count_launch();
  const long long total = (long long)P.V * P.V * k * k;
  seg_joint_tc_reduce_kernel<<<cdiv(total, 256), 256, 0, st>>>(part, joint, p.items, p.ugroups, P.V, k);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

// =================================================================================================
// Backward contractions on the tensor cores:
//     out[n,Y,X,c] = scale * sum_{u,v,c'} H[u][v][c][c'] * in[n, Y - s(u-T), X - s(v-T), c']        s = +1 | -1
// (s = +1, in = x2m -> d x1m ; s = -1, in = x1m -> d x2m ; what seg_corr_bwd_kernel evaluates in fp32 SIMT).
// Per output row Y and displacement row u this is D[X][c] += A[X][(j, c')] * B[c][(j, c')] with the input row
// inb[b] (b = x + T, zero halo) and  A[X][(j, c')] = inb[X + j][c']  -- a K-major operand whose rows are 64 B apart and
// OVERLAP along K (SWIZZLE_64B, the K offset j is applied by moving the descriptor start by j pixels; probe mode 1) --
// and B[c][(j, c')] = H[u][v(j)][c][c'] (v = 2T - j for s = +1, v = j for s = -1), prepared once in global memory and
// kept resident in shared memory for the CTA's group of 6 displacement rows.  M = 128 pixels (w <= 128), N = 16, K = 8
// (half a pixel's channels) per MMA; 16 accumulator columns per output row, double buffered.  Single-pass tf32 with
// round-to-nearest operands: the gradient tolerance is 2e-3 of max|g| (tests/test_gpu_parity_seg.py), the rounding
// noise of a 7056-term signed sum is ~3e-4 of it.
constexpr int SC_U = 6;
constexpr int SC_SLOTS = 7;
constexpr int SC_THREADS = 192;

struct ScParams {
  int n, h, w, T, V, sgn;
  int ychunk, nychunks, ugroups;
  int row_bytes, box_bytes;
  float* part;  // [ugroups][n*h*w][16]
};

__global__ void seg_round_tf32_kernel(const float* __restrict__ x, float* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    float4 h;
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x)); h.x = __uint_as_float(t);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y)); h.y = __uint_as_float(t);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z)); h.z = __uint_as_float(t);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w)); h.w = __uint_as_float(t);
    reinterpret_cast<float4*>(out)[i] = h;
  }
}

// H [V*V][k][k] -> Bg[u][j][c (16)][c' (16)], tf32-rounded, scaled, zero padded; v = (sgn > 0) ? 2T - j : j
__global__ void seg_hprep_kernel(const float* __restrict__ H, float* __restrict__ Bg, int V, int k, int T, int sgn, float scale) {
  const int total = V * V * 256;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int cp = i & 15, c = (i >> 4) & 15, j = (i >> 8) % V, u = (i >> 8) / V;
    const int v = sgn > 0 ? 2 * T - j : j;
    float val = (c < k && cp < k) ? H[(((long long)u * V + v) * k + c) * k + cp] * scale : 0.f;
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(val));
    Bg[i] = __uint_as_float(t);
  }
}

__device__ __forceinline__ uint64_t sc_desc(uint32_t saddr) {
  // K-major, SWIZZLE_64B: rows 64 B apart, 8-row groups 512 B apart (LBO unused)
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

__global__ void __launch_bounds__(SC_THREADS, 1)
seg_corr_tc_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmB, ScParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                    // SC_SLOTS input-row slots
  const uint32_t bbase = base + SC_SLOTS * (uint32_t)P.row_bytes;   // [SC_U][V][16 rows][64 B] resident coefficients
  const uint32_t bars = bbase + (uint32_t)(SC_U * P.V) * 1024u;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (SC_SLOTS + s); };
  auto tfull = [&](int a) { return bars + 8u * (2 * SC_SLOTS + a); };
  auto tempty = [&](int a) { return bars + 8u * (2 * SC_SLOTS + 2 + a); };
  const uint32_t bres = bars + 8u * (2 * SC_SLOTS + 4);
  uint8_t* bars_ptr = smem_raw + (bars - raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_ptr + 8 * (2 * SC_SLOTS + 5));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < SC_SLOTS; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull(a), 1);
      mbar_init(tempty(a), 128);
    }
    mbar_init(bres, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmIn);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), 32);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int g = blockIdx.x % P.ugroups;
  const int yc = (blockIdx.x / P.ugroups) % P.nychunks;
  const int img = blockIdx.x / (P.ugroups * P.nychunks);
  const int Ya = yc * P.ychunk, Yb = min(P.h, Ya + P.ychunk);
  const int u0 = g * SC_U;
  const int nu = min(SC_U, P.V - u0);
  // input row of (Y, ul): r = Y + d(ul), d(ul) = -sgn * (u0 + ul - T); d is monotone in ul
  const int d_first = -P.sgn * (u0 - P.T), d_last = -P.sgn * (u0 + nu - 1 - P.T);
  const int dmin = min(d_first, d_last), dmax = max(d_first, d_last);
  const int r_lo = max(0, Ya + dmin), r_hi = min(P.h - 1, Yb - 1 + dmax);

  if (warp == 0) {
    // =============================== TMA producer ============================================
    if (lane == 0) {
      // resident coefficients of this displacement-row group: rows (u, j, c) are contiguous in Bg; boxes of 7 j's
      mbar_expect_tx(bres, (uint32_t)(nu * P.V) * 1024u);
      for (int ul = 0; ul < nu; ++ul)
        for (int j0 = 0; j0 < P.V; j0 += 7) {
          const int nj = min(7, P.V - j0);
          (void)nj;  // (V is a multiple of 7 for T = 3, 10; other T take the 1-j box map: see sc_plan)
          tma_load_2d(bbase + (uint32_t)((ul * P.V + j0) * 1024), &tmB, bres, 0, ((u0 + ul) * P.V + j0) * 16);
        }
      int r_next = r_lo;
      for (int Y = Ya; Y < Yb; ++Y) {
        const int need = min(r_hi, Y + dmax);
        for (; r_next <= need; ++r_next) {
          const int idx = r_next - r_lo, s = idx % SC_SLOTS;
          mbar_wait(empty(s), ((idx / SC_SLOTS) & 1u) ^ 1u);
          mbar_expect_tx(full(s), (uint32_t)P.box_bytes);
          tma_load_3d(base + s * (uint32_t)P.row_bytes, &tmIn, full(s), 0, -P.T, img * P.h + r_next);
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ==============================================
    // kind::tf32, D = f32, M = 128, N = 16, A and B K-major
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    mbar_wait(bres, 0);
    for (int Y = Ya; Y < Yb; ++Y) {
      const uint32_t it = (uint32_t)(Y - Ya), as = it & 1u;
      mbar_wait(tempty(as), ((it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t acc = tmem_base + as * 16u;
      uint32_t started = 0u;
      for (int ul = 0; ul < nu; ++ul) {
        const int d = -P.sgn * (u0 + ul - P.T);
        const int r = Y + d;
        if (r < 0 || r >= P.h) continue;
        const int idx = r - r_lo, s = idx % SC_SLOTS;
        mbar_wait(full(s), (idx / SC_SLOTS) & 1u);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint32_t a0 = base + s * (uint32_t)P.row_bytes, b0 = bbase + (uint32_t)(ul * P.V) * 1024u;
          for (int j = 0; j < P.V; ++j) {
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              sj_mma(acc, sc_desc(a0 + (uint32_t)j * 64u + half * 32u), sc_desc(b0 + (uint32_t)j * 1024u + half * 32u), idesc,
                     (started | (uint32_t)j | (uint32_t)half) ? 1u : 0u);
            }
          }
          if (d == dmin) umma_commit(empty(s));  // last use of input row r (later rows Y pair it with smaller offsets only)
        }
        __syncwarp();
        started = 1u;
      }
      if (elect_one_sync()) umma_commit(tfull(as));
      __syncwarp();
    }
  } else {
    // =============================== epilogue (warps 2-5) ======================================
    const int quad = warp & 3;
    const int X = quad * 32 + lane;
    float* out = P.part + (long long)g * P.n * P.h * P.w * 16;
    for (int Y = Ya; Y < Yb; ++Y) {
      const uint32_t it = (uint32_t)(Y - Ya), as = it & 1u;
      bool any = false;
      for (int ul = 0; ul < nu; ++ul) {
        const int r = Y - P.sgn * (u0 + ul - P.T);
        any |= (r >= 0 && r < P.h);
      }
      mbar_wait(tfull(as), (it >> 1) & 1u);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16), v);  // both 16-column accumulators
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(tempty(as));
      if (X < P.w) {
        float* o = out + (((long long)img * P.h + Y) * P.w + X) * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
          if (any) {
            if (as == 0)
              t = make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]), __uint_as_float(v[q * 4 + 2]),
                              __uint_as_float(v[q * 4 + 3]));
            else
              t = make_float4(__uint_as_float(v[16 + q * 4]), __uint_as_float(v[16 + q * 4 + 1]),
                              __uint_as_float(v[16 + q * 4 + 2]), __uint_as_float(v[16 + q * 4 + 3]));
          }
          *reinterpret_cast<float4*>(o + q * 4) = t;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 32);
}

__global__ void seg_corr_tc_sum_kernel(const float* __restrict__ part, float* __restrict__ out, long long n4, int groups) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 t = reinterpret_cast<const float4*>(part)[i];
    for (int gi = 1; gi < groups; ++gi) {
      const float4 a = reinterpret_cast<const float4*>(part)[(long long)gi * n4 + i];
      t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
    }
    reinterpret_cast<float4*>(out)[i] = t;
  }
}

struct ScPlan {
  bool ok;
  int ychunk, nychunks, ugroups, ctas, row_bytes, box_bytes, smem;
};

static ScPlan sc_plan(int n, int k, int h, int w, int T) {
  ScPlan p = {};
  const int V = 2 * T + 1;
  if (k > 16 || k < 5 || V > 24 || V % 7 != 0 || n < 1 || w > 128) return p;  // coefficient boxes hold 7 j's: T = 3, 10
  p.box_bytes = (128 + V - 1) * 64;
  p.row_bytes = (p.box_bytes + 1023) / 1024 * 1024;
  p.smem = 1024 + SC_SLOTS * p.row_bytes + SC_U * V * 1024 + 256;
  if (p.smem > 232448) return p;
  p.ugroups = (V + SC_U - 1) / SC_U;
  int want = (3 * device_sm_count() + n * p.ugroups - 1) / (n * p.ugroups);
  if (want < 1) want = 1;
  p.ychunk = (h + want - 1) / want;
  if (p.ychunk < 4) p.ychunk = 4;
  if (p.ychunk > h) p.ychunk = h;
  p.nychunks = (h + p.ychunk - 1) / p.ychunk;
  p.ctas = n * p.nychunks * p.ugroups;
  p.ok = true;
  return p;
}

long long seg_corr_tc_workspace(int n, int k, int h, int w, int T) {
  const ScPlan p = sc_plan(n, k, h, w, T);
  if (!p.ok) return 0;
  const int V = 2 * T + 1;
  return ((long long)n * h * w * 16 * (1 + p.ugroups) + (long long)V * V * 256) * (long long)sizeof(float);
}

int seg_corr_tc(const float* in, const float* H, float* out, void* workspace, int n, int k, int h, int w, int T, int sgn, float scale,
                cudaStream_t st) {
  int rc = sj_init();
  if (rc != IIC_OK) return rc;
  const ScPlan p = sc_plan(n, k, h, w, T);
  IIC_REQUIRE(p.ok, IIC_ERR_UNSUPPORTED, "seg_corr_tc: unsupported geometry (k=%d, T=%d, w=%d)", k, T, w);
  const int V = 2 * T + 1;
  const long long elems = (long long)n * h * w * 16;
  float* inr = (float*)workspace;
  float* part = inr + elems;
  float* Bg = part + elems * p.ugroups;
  int blocks = cdiv(elems / 4, 256);
  if (blocks > device_sm_count() * 16) blocks = device_sm_count() * 16;
  seg_round_tf32_kernel<<<blocks, 256, 0, st>>>(in, inr, elems / 4);
  IIC_LAUNCH_CHECK();
  count_launch();
  seg_hprep_kernel<<<cdiv((long long)V * V * 256, 256), 256, 0, st>>>(H, Bg, V, k, T, sgn, scale);
  IIC_LAUNCH_CHECK();
  count_launch();
  alignas(64) CUtensorMap tmIn, tmB;
  if ((rc = sj_map(&tmIn, inr, n * h, w, 128 + V - 1)) != IIC_OK) return rc;
  {
    cuuint64_t gdim[2] = {16, (cuuint64_t)V * V * 16};
    cuuint64_t gstr[1] = {64};
    cuuint32_t box[2] = {16, 7 * 16};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = sj_encodeTiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, Bg, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IIC_REQUIRE(r == CUDA_SUCCESS, IIC_ERR_CUDA, "cuTensorMapEncodeTiled(seg corr coefficients) failed (%d)", (int)r);
  }
  ScParams P = {};
  P.n = n; P.h = h; P.w = w; P.T = T; P.V = V; P.sgn = sgn;
  P.ychunk = p.ychunk; P.nychunks = p.nychunks; P.ugroups = p.ugroups;
  P.row_bytes = p.row_bytes; P.box_bytes = p.box_bytes;
  P.part = part;
  IIC_CUDA(cudaFuncSetAttribute(seg_corr_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, p.smem));
  seg_corr_tc_kernel<<<p.ctas, SC_THREADS, p.smem, st>>>(tmIn, tmB, P);
  IIC_LAUNCH_CHECK();
  count_launch();
  seg_corr_tc_sum_kernel<<<blocks, 256, 0, st>>>(part, out, elems / 4, p.ugroups);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

}  // namespace iic

extern "C" long long iic_seg_corr_tc_workspace(int n, int k, int h, int w, int T) { return iic::seg_corr_tc_workspace(n, k, h, w, T); }

extern "C" int iic_seg_corr_tc(const float* in, const float* H, float* out, void* workspace, int n, int k, int h, int w, int T,
                               int sgn, float scale, void* stream) {
  IIC_REQUIRE(in && H && out && workspace && n > 0 && k > 0 && T >= 0 && (sgn == 1 || sgn == -1), IIC_ERR_BAD_ARG,
              "iic_seg_corr_tc: bad arguments");
  return iic::seg_corr_tc(in, H, out, workspace, n, k, h, w, T, sgn, scale, (cudaStream_t)stream);
}

extern "C" long long iic_seg_joint_tc_workspace(int n, int k, int h, int w, int T) {
  return iic::seg_joint_tc_workspace(n, k, h, w, T);
}

extern "C" int iic_seg_joint_tc(const float* x1m, const float* x2m, float* joint, void* workspace, int n, int k, int h, int w,
                                int T, void* stream) {
  IIC_REQUIRE(x1m && x2m && joint && workspace && n > 0 && k > 0 && T >= 0, IIC_ERR_BAD_ARG, "iic_seg_joint_tc: bad arguments");
  return iic::seg_joint_tc(x1m, x2m, joint, workspace, n, k, h, w, T, (cudaStream_t)stream);
}
