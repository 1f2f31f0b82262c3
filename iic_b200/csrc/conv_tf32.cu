// tcgen05 implicit-GEMM convolution on fp32 activations: kind::tf32, and the error-compensated 3xTF32 split.
//
// Why: the reference is fp32 and `north_star` asks for "a stated fp32 tolerance".  tcgen05 has no fp32 kind; the bf16
// path (conv_tc2.cu) rounds every stored activation to 8 significant bits, which a 34-layer ReLU network turns into a
// ~35-50 % gradient error (ReLU sign flips: gradient error ~ sqrt(forward error)).  This file keeps the
// activations in fp32 in HBM and offers two tensor-core modes of the same kernel:
//   SPLIT = 1 (IIC_TF32)   : operands read as tf32 (10-bit mantissa), fp32 accumulation in TMEM.
//   SPLIT = 3 (IIC_TF32X3) : each operand x = hi + lo with hi = x truncated to tf32, lo = x - hi (exact); the product is
//                            accumulated as lo_a*hi_b + hi_a*lo_b + hi_a*hi_b (dropped: lo*lo and the truncation of lo, both
//                            ~2^-20 relative): fp32-grade results from the tensor pipe at 1/3 of the tf32 rate, ~10x the fp32
//                            SIMT kernel (conv_simt.cu).
// Same design as conv_tc2.cu (TMA im2col A operand, tiled TMA dense operand, persistent CTAs, double-buffered TMEM
// accumulators, split-K wgrad with both operands MN-major, stride-2 dgrad by output-parity classes), with a k-block of
// 32 fp32 = one 128-byte SWIZZLE_128B row and UMMA_K = 8.  SPLIT = 3 adds four "splitter" warps between the TMA
// producer and the MMA issuer: they rewrite a landed stage in place as hi and write lo into a twin buffer at the same
// offsets (so the swizzle pattern is preserved), fence the generic->async proxy and arrive on a second barrier.
// wgrad: both operands arrive pixel-major ([pixel][channel] = MN-major).  tcgen05 kind::tf32 has no MN-major mode for the
// ordinary swizzles (measured: a kind::tf32 MMA with MN-major SWIZZLE_64B / 128B descriptors returns zeros,
// tools/umma_sw64_probe.cu; 32-bit operands would need the separate 128B_BASE32B layout), so the same four warps
// TRANSPOSE every landed 32-pixel x 32-channel block in place (one warp per block: a lane reads one pixel row into 32
// registers, __syncwarp, writes one element of each channel row; conflict-free in both directions under the 128-byte
// swizzle) -- after which a wgrad stage is an ordinary K-major stage and the issue path is the fprop one.
// Shared-memory fill per tensor cycle: 3xTF32 needs 32 KB per 768 MMA cycles (43 B/clk, tensor bound); plain TF32 needs
// the same 32 KB per 256 cycles (128 B/clk, L2->SM bound at N <= 128) -- the compensated mode is the one tuned here.
#include <cuda.h>

#include <cstdlib>

#include "tc_ptx.cuh"

namespace iic {

enum { T32_FPROP = 0, T32_WGRAD = 1 };
constexpr int T32_MAXTAPS = 25;
constexpr int T32_KB = 32;                 // fp32 elements per k-block (128 bytes)
constexpr int T32_A_BYTES = TC_BM * 128;   // 16 KB

struct T32Params {
  long long rows;
  int rowH, rowW;
  int KH, KW, s, d;
  int lower;
  int lower_h, lower_w;
  int ntaps;
  unsigned char offh[T32_MAXTAPS], offw[T32_MAXTAPS], wtap[T32_MAXTAPS];
  int scatter, outH, outW, py, px;
  int srcC, Ktot, N;
  int mtiles, ntiles, splits, kb_per_split, total_kb;
  float* out;
  const float* addend;
  float* partial;
  int raw_hi;  // 3xTF32 fprop / dgrad: the stage keeps the raw fp32 values as the hi operand (option tf32x3_raw_hi)
};

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor, kind::tf32: D = f32, A = B = tf32, M = 128, N = BN
__host__ __device__ constexpr uint32_t make_idesc_tf32(int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
}
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
// hi part of the 3xTF32 split by truncation (one LOP3; cvt.rna.tf32 is a slow-path conversion and the four transform warps
// have ~16 K elements to split per stage): hi = x with the 13 low mantissa bits cleared, so lo = x - hi is EXACT in fp32
// and is itself read by the tensor core as tf32 (truncated to 10 bits: |x - hi - tf32(lo)| <= 2^-20 |x|).
__device__ __forceinline__ float hi_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }

template <int BN, int SPLIT, int MODE = 0> struct T32Cfg {
  static constexpr bool XFORM = (SPLIT == 3) || (MODE == 1);  // transform warps: hi/lo split and / or wgrad block transposes
  static constexpr int B_BYTES = BN * 128;
  static constexpr int HALF = T32_A_BYTES + B_BYTES;               // one precision part of a stage: [A | B]
  static constexpr int STAGE_BYTES = HALF * (SPLIT == 3 ? 2 : 1);  // [A hi | B hi | A lo | B lo]
  static constexpr int STAGES = (SPLIT == 3) ? (BN >= 128 ? 3 : 4) : (BN >= 128 ? 6 : 8);
  static constexpr int THREADS = XFORM ? 320 : 192;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 2 * BN;
};

template <int MODE, int BN, int SPLIT>
__global__ void __launch_bounds__(((SPLIT == 3) || (MODE == 1)) ? 320 : 192, 1)
conv_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, T32Params P) {
  using Cfg = T32Cfg<BN, SPLIT, MODE>;
  constexpr bool XFORM = Cfg::XFORM;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto split_bar = [&](int s) { return bars + 8u * (2 * STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (3 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (3 * STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + STAGES * Cfg::STAGE_BYTES + (3 * STAGES + 4) * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(split_bar(s), 128);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_mn = P.mtiles * P.ntiles;
  const int total_work = tiles_mn * P.splits;
  auto decode = [&](int w, int& z, long long& m0, int& n0, int& kb0, int& nk) {
    z = w / tiles_mn;
    const int r = w - z * tiles_mn;
    n0 = (r % P.ntiles) * BN;
    m0 = (long long)(r / P.ntiles) * TC_BM;
    if (MODE == T32_FPROP) {
      kb0 = 0;
      nk = P.total_kb;
    } else {
      kb0 = z * P.kb_per_split;
      nk = min(P.total_kb, kb0 + P.kb_per_split) - kb0;
      if (nk < 0) nk = 0;
    }
  };

  if (warp == 5) {
    // =============================== TMA producer ============================================
    uint32_t it = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      int z, n0, kb0, nk;
      long long m0;
      decode(w, z, m0, n0, kb0, nk);
      if (MODE == T32_FPROP) {
        const int ox = (int)(m0 % P.rowW);
        const long long q = m0 / P.rowW;
        const int oy = (int)(q % P.rowH);
        const int img = (int)(q / P.rowH);
        const int cw = ox * P.s + P.lower_w, ch = oy * P.s + P.lower_h;
        int tap = 0, c0 = 0;
        for (int i = 0; i < nk; ++i, ++it) {
          const int s = it % STAGES;
          mbar_wait(empty_bar(s), ((it / STAGES) & 1u) ^ 1u);
          const uint32_t sa = base + s * Cfg::STAGE_BYTES, sb = sa + T32_A_BYTES;
          // 3xTF32: the weights arrive split (second plane of the packed buffer = lo, N rows further down)
          if (lane == 0) mbar_expect_tx(full_bar(s), Cfg::HALF + (SPLIT == 3 ? Cfg::B_BYTES : 0));
          __syncwarp();
          if (lane == 0)
            tma_load_im2col(sa, &tmA, full_bar(s), c0, cw, ch, img, (uint16_t)P.offw[tap], (uint16_t)P.offh[tap]);
          else if (lane == 1)
            tma_load_2d(sb, &tmB, full_bar(s), (int)P.wtap[tap] * P.srcC + c0, n0);
          else if (SPLIT == 3 && lane == 2)
            tma_load_2d(sb + Cfg::HALF, &tmB, full_bar(s), (int)P.wtap[tap] * P.srcC + c0, n0 + P.N);
          c0 += T32_KB;
          if (c0 >= P.srcC) {
            c0 = 0;
            ++tap;
          }
        }
      } else {
        // A tile: 128 (tap, cin) rows as four 32-channel column blocks, each [32 pixels][128 B] (MN-major atoms)
        int a_off_w = 0, a_off_h = 0, a_c0 = 0;
        bool a_ok = false;
        int nok = 0;
        for (int b = 0; b < 4; ++b) nok += (m0 + b * 32 < P.Ktot) ? 1 : 0;
        if (lane < 4) {
          const long long j = m0 + lane * 32;
          a_ok = j < P.Ktot;
          const int tap = (int)(j / P.srcC);
          a_c0 = (int)(j - (long long)tap * P.srcC);
          const int ta = tap / P.KW;
          a_off_h = ta * P.d;
          a_off_w = (tap - ta * P.KW) * P.d;
        }
        const uint32_t bytes = (uint32_t)(nok * 4096 + Cfg::B_BYTES);
        long long p0 = (long long)kb0 * T32_KB;
        int ox = (int)(p0 % P.rowW);
        const long long q = p0 / P.rowW;
        int oy = (int)(q % P.rowH);
        int img = (int)(q / P.rowH);
        const int step_x = T32_KB % P.rowW, step_y = T32_KB / P.rowW;
        for (int i = 0; i < nk; ++i, ++it) {
          const int s = it % STAGES;
          mbar_wait(empty_bar(s), ((it / STAGES) & 1u) ^ 1u);
          const uint32_t sa = base + s * Cfg::STAGE_BYTES, sb = sa + T32_A_BYTES;
          if (lane == 0) mbar_expect_tx(full_bar(s), bytes);
          __syncwarp();
          if (lane < 4) {
            if (a_ok)
              tma_load_im2col(sa + lane * 4096, &tmA, full_bar(s), a_c0, ox * P.s + P.lower, oy * P.s + P.lower, img,
                              (uint16_t)a_off_w, (uint16_t)a_off_h);
          } else if (lane < 4 + BN / 32) {
            const int b = lane - 4;
            tma_load_2d(sb + b * 4096, &tmB, full_bar(s), n0 + b * 32, (int)p0);
          }
          p0 += T32_KB;
          ox += step_x;
          oy += step_y;
          if (ox >= P.rowW) {
            ox -= P.rowW;
            ++oy;
          }
          while (oy >= P.rowH) {
            oy -= P.rowH;
            ++img;
          }
        }
      }
    }
  } else if (warp == 4) {
    // =============================== MMA issuer ==============================================
    // both modes issue K-major operands (wgrad stages are transposed in shared memory first): rows of 128 B = 32 K
    // elements, 8-row groups 1024 B apart, LBO unused
    constexpr uint32_t idesc = make_idesc_tf32(BN, 0, 0);
    uint32_t s = 0, sphase = 0, tile_it = 0;
    const uint64_t desc_base = make_desc(base, 16, 1024);
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++tile_it) {
      int z, n0, kb0, nk;
      long long m0;
      decode(w, z, m0, n0, kb0, nk);
      const uint32_t as = tile_it & 1u;
      mbar_wait(tempty_bar(as), ((tile_it >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + as * BN;
      for (int i = 0; i < nk; ++i) {
        mbar_wait(XFORM ? split_bar(s) : full_bar(s), sphase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t ad0 = desc_base + (uint64_t)((s * (uint32_t)Cfg::STAGE_BYTES) >> 4);
          const uint64_t bd0 = ad0 + (uint64_t)(T32_A_BYTES >> 4);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t koff = kk * 32;  // 32 B per K = 8 step inside the 128 B swizzle row
            const uint64_t ad = ad0 + (uint64_t)(koff >> 4);
            const uint64_t bd = bd0 + (uint64_t)(koff >> 4);
            if (SPLIT == 3) {
              const uint64_t lo = (uint64_t)(Cfg::HALF >> 4);
              umma_tf32(tmem_acc, ad + lo, bd, idesc, (i > 0 || kk > 0) ? 1u : 0u);  // lo_a * hi_b
              umma_tf32(tmem_acc, ad, bd + lo, idesc, 1u);                            // hi_a * lo_b
              umma_tf32(tmem_acc, ad, bd, idesc, 1u);                                 // hi_a * hi_b
            } else {
              umma_tf32(tmem_acc, ad, bd, idesc, (i > 0 || kk > 0) ? 1u : 0u);
            }
          }
          umma_commit(empty_bar(s));
        }
        __syncwarp();
        if (++s == (uint32_t)STAGES) {
          s = 0;
          sphase ^= 1u;
        }
      }
      if (elect_one_sync()) umma_commit(tfull_bar(as));
      __syncwarp();
    }
  } else if (warp >= 6) {
    // =============================== transform warps (6-9): wgrad block transposes and / or hi-lo split =============
    if constexpr (XFORM) {
      const int tid = threadIdx.x - 192, tw = tid >> 5;
      uint32_t s = 0, sphase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        int z, n0, kb0, nk;
        long long m0;
        decode(w, z, m0, n0, kb0, nk);
        for (int i = 0; i < nk; ++i) {
          mbar_wait(full_bar(s), sphase);
          uint8_t* st_hi = base_ptr + s * Cfg::STAGE_BYTES;
          uint8_t* st_lo = st_hi + Cfg::HALF;
          if constexpr (MODE == T32_WGRAD) {
            // [32 pixels][32 channels] -> [32 channels][32 pixels], block by block (A: 4 blocks, B: BN / 32), in place
            constexpr int NBLK = 4 + BN / 32;
            for (int b = tw; b < NBLK; b += 4) {
              uint8_t* blk = st_hi + b * 4096;
              float v[32];
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 t = *reinterpret_cast<const float4*>(blk + lane * 128 + ((q ^ (lane & 7)) << 4));
                v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
              }
              __syncwarp();  // the whole block is in registers before any lane overwrites it
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                const uint32_t off = (uint32_t)(c * 128 + ((((uint32_t)lane >> 2) ^ (uint32_t)(c & 7)) << 4) + (lane & 3) * 4);
                if (SPLIT == 3) {
                  const float h = hi_tf32(v[c]);
                  *reinterpret_cast<float*>(blk + off) = h;
                  *reinterpret_cast<float*>(st_lo + b * 4096 + off) = v[c] - h;
                } else {
                  *reinterpret_cast<float*>(blk + off) = v[c];
                }
              }
            }
          } else {
            // only the activation tile is split here: the weight tile arrives as [raw | lo] from the packed buffer
            float4* hi = reinterpret_cast<float4*>(st_hi);
            float4* lo = reinterpret_cast<float4*>(st_lo);
            // all of a thread's 16-byte chunks are loaded before the first is used (the loop was bound by the
            // shared-memory load latency: 23 % of the kernel's stall samples sat on its first LOP3, profiles/r02_ncu_conv_tf32.md)
            constexpr int ITER = T32_A_BYTES / 16 / 128;
            static_assert(ITER * 128 * 16 == T32_A_BYTES, "A tile size must be a multiple of 128 x 16 bytes");
            const bool raw_hi = P.raw_hi != 0;
            float4 v[ITER];
#pragma unroll
            for (int it = 0; it < ITER; ++it) v[it] = hi[tid + it * 128];
#pragma unroll
            for (int it = 0; it < ITER; ++it) {
              const int c = tid + it * 128;
              float4 h, l;
              h.x = hi_tf32(v[it].x); h.y = hi_tf32(v[it].y); h.z = hi_tf32(v[it].z); h.w = hi_tf32(v[it].w);
              l.x = v[it].x - h.x; l.y = v[it].y - h.y; l.z = v[it].z - h.z; l.w = v[it].w - h.w;
              if (!raw_hi) hi[c] = h;
              lo[c] = l;
            }
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
          mbar_arrive(split_bar(s));
          if (++s == (uint32_t)STAGES) {
            s = 0;
            sphase ^= 1u;
          }
        }
      }
    }
  } else {
    // =============================== epilogue (warps 0-3) ======================================
    uint32_t tile_it = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++tile_it) {
      int z, n0, kb0, nk;
      long long m0;
      decode(w, z, m0, n0, kb0, nk);
      const uint32_t as = tile_it & 1u;
      const long long m = m0 + warp * 32 + lane;  // TMEM lane == tile row
      const uint32_t tmem_acc = tmem_base + as * BN + ((uint32_t)(warp * 32) << 16);
      long long orow = m;
      if (MODE == T32_FPROP && P.scatter && m < P.rows) {
        const int jj = (int)(m % P.rowW);
        const long long q = m / P.rowW;
        const int ii = (int)(q % P.rowH);
        orow = ((q / P.rowH) * P.outH + 2 * ii + P.py) * P.outW + 2 * jj + P.px;
      }
      mbar_wait(tfull_bar(as), (tile_it >> 1) & 1u);
      tc_fence_after();
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        if (nk > 0) {
          tmem_ld32(tmem_acc + (uint32_t)c0, v);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int qq = 0; qq < 32; ++qq) v[qq] = 0u;
        }
        if constexpr (MODE == T32_FPROP) {
          if (m < P.rows) {
            float* o = P.out + orow * P.N + n0 + c0;
            if (P.addend != nullptr) {
              const float4* ad = reinterpret_cast<const float4*>(P.addend + orow * P.N + n0 + c0);
#pragma unroll
              for (int qq = 0; qq < 8; ++qq) {
                const float4 a = ad[qq];
                *reinterpret_cast<float4*>(o + qq * 4) =
                    make_float4(__uint_as_float(v[qq * 4]) + a.x, __uint_as_float(v[qq * 4 + 1]) + a.y,
                                __uint_as_float(v[qq * 4 + 2]) + a.z, __uint_as_float(v[qq * 4 + 3]) + a.w);
              }
            } else {
#pragma unroll
              for (int qq = 0; qq < 8; ++qq)
                *reinterpret_cast<float4*>(o + qq * 4) =
                    make_float4(__uint_as_float(v[qq * 4]), __uint_as_float(v[qq * 4 + 1]), __uint_as_float(v[qq * 4 + 2]),
                                __uint_as_float(v[qq * 4 + 3]));
            }
          }
        } else {
          if (m < P.Ktot) {
            float* o = P.partial + ((long long)z * P.Ktot + m) * P.N + n0 + c0;
#pragma unroll
            for (int qq = 0; qq < 8; ++qq)
              *reinterpret_cast<float4*>(o + qq * 4) =
                  make_float4(__uint_as_float(v[qq * 4]), __uint_as_float(v[qq * 4 + 1]), __uint_as_float(v[qq * 4 + 2]),
                              __uint_as_float(v[qq * 4 + 3]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(as));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

__global__ void wgrad_tf32_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int Ktot, int N, int splits) {
  const long long total = (long long)Ktot * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % N);
    const int j = (int)(i / N);
    float t = 0.f;
    for (int z = 0; z < splits; ++z) t += partial[(long long)z * total + i];
    dw[(long long)co * Ktot + j] = t;
  }
}

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*PFN_t32EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_t32EncodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_t32EncodeTiled t32_encodeTiled = nullptr;
static PFN_t32EncodeIm2col t32_encodeIm2col = nullptr;
static int t32_driver_version = 0;

static int t32_init() {
  if (t32_encodeTiled && t32_encodeIm2col) return IIC_OK;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  IIC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  IIC_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, IIC_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  t32_encodeTiled = (PFN_t32EncodeTiled)fn;
  fn = nullptr;
  IIC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres));
  IIC_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, IIC_ERR_CUDA, "cuTensorMapEncodeIm2col unavailable");
  t32_encodeIm2col = (PFN_t32EncodeIm2col)fn;
  cudaDriverGetVersion(&t32_driver_version);
  return IIC_OK;
}

// NHWC fp32 activation [nimg][H][W][C] as the rank-4 (C, W, H, N) im2col tensor map; box = 32 channels x `pixels`
static int t32_im2col_map(CUtensorMap* tm, const float* ptr, int nimg, int H, int W, int C, int lower_w, int lower_h, int upper_w,
                          int upper_h, int stride, int pixels) {
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  int lo[2] = {lower_w, lower_h}, up[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = t32_encodeIm2col(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(ptr), gdim, gstr, lo, up, T32_KB,
                                (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IIC_REQUIRE(r == CUDA_SUCCESS, IIC_ERR_CUDA, "cuTensorMapEncodeIm2col(fp32) failed (%d) H=%d W=%d C=%d lower=%d,%d upper=%d,%d stride=%d",
              (int)r, H, W, C, lower_w, lower_h, upper_w, upper_h, stride);
  // driver workaround also applied by CUTLASS (copy_traits_sm90_im2col.hpp): small tensors, drivers <= 13.1
  if (t32_driver_version <= 13010 && (long long)nimg * H * W * C * 4 < 131072)
    reinterpret_cast<uint64_t*>(tm)[1] &= ~(1llu << 21);
  return IIC_OK;
}

// row-major fp32 matrix [rows][cols] -> boxes of 32 columns x box_rows rows
static int t32_tiled_map(CUtensorMap* tm, const float* ptr, long long rows, long long cols, int box_rows) {
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
  cuuint32_t box[2] = {T32_KB, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = t32_encodeTiled(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IIC_REQUIRE(r == CUDA_SUCCESS, IIC_ERR_CUDA, "cuTensorMapEncodeTiled(fp32) failed (%d) rows=%lld cols=%lld", (int)r, rows, cols);
  return IIC_OK;
}

static int t32_pick_bn(int N) { return (N % 128 == 0) ? 128 : ((N % 64 == 0) ? 64 : 0); }

template <int MODE, int BN, int SPLIT>
static int t32_launch_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const T32Params& P, cudaStream_t st) {
  using Cfg = T32Cfg<BN, SPLIT, MODE>;
  static_assert(Cfg::SMEM <= 232448, "shared memory budget");
  IIC_CUDA(cudaFuncSetAttribute(conv_tf32_kernel<MODE, BN, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  const long long work = (long long)P.mtiles * P.ntiles * P.splits;
  const int grid = (int)(work < device_sm_count() ? work : device_sm_count());
  conv_tf32_kernel<MODE, BN, SPLIT><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(tmA, tmB, P);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

template <int MODE>
static int t32_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const T32Params& P, int bn, int split, cudaStream_t st) {
  if (bn == 128) return split == 3 ? t32_launch_impl<MODE, 128, 3>(tmA, tmB, P, st) : t32_launch_impl<MODE, 128, 1>(tmA, tmB, P, st);
  return split == 3 ? t32_launch_impl<MODE, 64, 3>(tmA, tmB, P, st) : t32_launch_impl<MODE, 64, 1>(tmA, tmB, P, st);
}

// fprop (transposed == 0) or dgrad of a stride-1 conv (transposed == 1; src = dy, N = cin); split = 1 | 3
int tf32_conv_gather_gemm(const float* src, int srcH, int srcW, int srcC, int rowH, int rowW, int nimg, const iic_conv_geom* g,
                          int transposed, const float* wpacked, int N, const float* addend, float* out, int split,
                          cudaStream_t st) {
  int rc = t32_init();
  if (rc != IIC_OK) return rc;
  const int bn = t32_pick_bn(N);
  IIC_REQUIRE(bn != 0 && srcC % T32_KB == 0, IIC_ERR_UNSUPPORTED, "tf32 conv: cin %% 32 == 0 and cout %% 64 == 0 required");
  IIC_REQUIRE(g->kh == g->kw, IIC_ERR_UNSUPPORTED, "tf32 conv: square filters only");
  IIC_REQUIRE(!transposed || g->stride == 1, IIC_ERR_UNSUPPORTED, "tf32 dgrad: stride-1 only on this entry");
  T32Params P = {};
  P.raw_hi = (split == 3) ? option(OPT_TF32X3_RAW_HI) : 0;
  P.rows = (long long)nimg * rowH * rowW;
  P.rowH = rowH; P.rowW = rowW; P.KH = g->kh; P.KW = g->kw; P.d = g->dil;
  const int span = (g->kh - 1) * g->dil;
  IIC_REQUIRE(g->kh * g->kw <= T32_MAXTAPS && span <= 255, IIC_ERR_UNSUPPORTED, "tf32 conv: filter too large");
  int upper;
  if (!transposed) {
    P.s = g->stride; P.lower = -g->pad;
    upper = g->pad - span;
  } else {
    P.s = 1; P.lower = g->pad - span;
    upper = P.lower + (rowH - srcH);
  }
  P.lower_h = P.lower_w = P.lower;
  P.ntaps = g->kh * g->kw;
  for (int t = 0; t < P.ntaps; ++t) {
    const int ta = t / g->kw, tb = t % g->kw;
    P.offh[t] = (unsigned char)((transposed ? g->kh - 1 - ta : ta) * g->dil);
    P.offw[t] = (unsigned char)((transposed ? g->kw - 1 - tb : tb) * g->dil);
    P.wtap[t] = (unsigned char)t;
  }
  IIC_REQUIRE(P.lower >= -128 && P.lower <= 127 && upper >= -128 && upper <= 127, IIC_ERR_UNSUPPORTED, "im2col corner range");
  P.srcC = srcC; P.Ktot = g->kh * g->kw * srcC; P.N = N;
  P.ntiles = N / bn; P.splits = 1; P.total_kb = P.Ktot / T32_KB;
  P.mtiles = (int)((P.rows + TC_BM - 1) / TC_BM);
  P.out = out; P.addend = addend;
  alignas(64) CUtensorMap tmA, tmB;
  rc = t32_im2col_map(&tmA, src, nimg, srcH, srcW, srcC, P.lower, P.lower, upper, upper, P.s, TC_BM);
  if (rc != IIC_OK) return rc;
  rc = t32_tiled_map(&tmB, wpacked, (split == 3 ? 2 : 1) * (long long)N, P.Ktot, bn);  // split 3: [raw plane | lo plane]
  if (rc != IIC_OK) return rc;
  return t32_launch<T32_FPROP>(tmA, tmB, P, bn, split, st);
}

// dgrad of a stride-2 convolution by output-parity classes (see tc2_conv_dgrad_s2 in conv_tc2.cu for the derivation)
int tf32_conv_dgrad_s2(const float* dy, const float* wpacked_t, const float* addend, float* dx, const iic_conv_geom* g, int split,
                       cudaStream_t st) {
  int rc = t32_init();
  if (rc != IIC_OK) return rc;
  const int bn = t32_pick_bn(g->cin);
  IIC_REQUIRE(bn != 0 && g->cout % T32_KB == 0 && g->stride == 2 && g->dil == 1 && g->kh == g->kw && g->kh * g->kw <= T32_MAXTAPS,
              IIC_ERR_UNSUPPORTED, "tf32 stride-2 dgrad: unsupported geometry");
  alignas(64) CUtensorMap tmB;
  const int Kw = g->kh * g->kw * g->cout;
  rc = t32_tiled_map(&tmB, wpacked_t, (split == 3 ? 2 : 1) * (long long)g->cin, Kw, bn);  // split 3: [raw plane | lo plane]
  if (rc != IIC_OK) return rc;
  auto parity_taps = [&](int par, int ksz) {
    int cnt = 0;
    for (int a = 0; a < ksz; ++a) cnt += (((par + g->pad - a) % 2) + 2) % 2 == 0;
    return cnt;
  };
  bool need_fill = false;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px)
      if (parity_taps(py, g->kh) * parity_taps(px, g->kw) == 0) need_fill = true;
  const size_t dx_bytes = sizeof(float) * (size_t)g->n * g->h * g->w * g->cin;
  if (need_fill) {  // classes that no tap reaches (1x1 stride 2): dx = addend or 0 there
    if (addend != nullptr)
      IIC_CUDA(cudaMemcpyAsync(dx, addend, dx_bytes, cudaMemcpyDeviceToDevice, st));
    else
      IIC_CUDA(cudaMemsetAsync(dx, 0, dx_bytes, st));
  }
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      const int Hc = (g->h - py + 1) / 2, Wc = (g->w - px + 1) / 2;
      if (Hc <= 0 || Wc <= 0) continue;
      T32Params P = {};
      P.raw_hi = (split == 3) ? option(OPT_TF32X3_RAW_HI) : 0;
      int offs_h[8], taps_h[8], nh = 0, offs_w[8], taps_w[8], nw = 0;
      for (int a = 0; a < g->kh; ++a) {
        const int tnum = py + g->pad - a;
        if (((tnum % 2) + 2) % 2 == 0) { offs_h[nh] = tnum / 2; taps_h[nh++] = a; }
      }
      for (int b = 0; b < g->kw; ++b) {
        const int tnum = px + g->pad - b;
        if (((tnum % 2) + 2) % 2 == 0) { offs_w[nw] = tnum / 2; taps_w[nw++] = b; }
      }
      if (nh == 0 || nw == 0) continue;
      int lo_h = offs_h[0], lo_w = offs_w[0];
      for (int i = 1; i < nh; ++i) lo_h = offs_h[i] < lo_h ? offs_h[i] : lo_h;
      for (int i = 1; i < nw; ++i) lo_w = offs_w[i] < lo_w ? offs_w[i] : lo_w;
      P.rows = (long long)g->n * Hc * Wc;
      P.rowH = Hc; P.rowW = Wc; P.KH = g->kh; P.KW = g->kw; P.s = 1; P.d = 1;
      P.lower_h = lo_h; P.lower_w = lo_w; P.lower = 0;
      P.ntaps = nh * nw;
      for (int i = 0; i < nh; ++i)
        for (int j = 0; j < nw; ++j) {
          const int t = i * nw + j;
          P.offh[t] = (unsigned char)(offs_h[i] - lo_h);
          P.offw[t] = (unsigned char)(offs_w[j] - lo_w);
          P.wtap[t] = (unsigned char)(taps_h[i] * g->kw + taps_w[j]);
        }
      P.scatter = 1; P.outH = g->h; P.outW = g->w; P.py = py; P.px = px;
      P.srcC = g->cout; P.Ktot = P.ntaps * g->cout; P.N = g->cin;
      P.mtiles = (int)((P.rows + TC_BM - 1) / TC_BM); P.ntiles = g->cin / bn; P.splits = 1; P.total_kb = P.Ktot / T32_KB;
      P.out = dx; P.addend = addend;
      const int up_h = lo_h + (Hc - g->oh), up_w = lo_w + (Wc - g->ow);
      IIC_REQUIRE(lo_h >= -128 && up_h <= 127 && lo_w >= -128 && up_w <= 127 && up_h >= -128 && up_w >= -128,
                  IIC_ERR_UNSUPPORTED, "im2col corner range");
      alignas(64) CUtensorMap tmA;
      rc = t32_im2col_map(&tmA, dy, g->n, g->oh, g->ow, g->cout, lo_w, lo_h, up_w, up_h, 1, TC_BM);
      if (rc != IIC_OK) return rc;
      rc = t32_launch<T32_FPROP>(tmA, tmB, P, bn, split, st);
      if (rc != IIC_OK) return rc;
    }
  return IIC_OK;
}

static int t32_wgrad_splits(const iic_conv_geom* g) {
  const long long rows = (long long)g->n * g->oh * g->ow;
  const int total_kb = (int)((rows + T32_KB - 1) / T32_KB);
  const int Ktot = g->kh * g->kw * g->cin;
  const int bn = t32_pick_bn(g->cout);
  const long long tiles = (long long)((Ktot + TC_BM - 1) / TC_BM) * (g->cout / (bn ? bn : 64));
  const long long sms = device_sm_count();
  long long want = 1;
  if (tiles < sms) {
    want = sms / tiles;
    const long long two = (2 * sms) / tiles;
    if (two > want && total_kb / two >= 64) want = two;
  }
  if (want > total_kb / 16) want = total_kb / 16;
  if (want < 1) want = 1;
  if (want > 512) want = 512;
  return (int)want;
}

long long tf32_conv_wgrad_workspace(const iic_conv_geom* g) {
  return (long long)t32_wgrad_splits(g) * g->kh * g->kw * g->cin * g->cout * (long long)sizeof(float);
}

int tf32_conv_wgrad(const float* x, const float* dy, float* dw, float* ws, const iic_conv_geom* g, int split, cudaStream_t st) {
  int rc = t32_init();
  if (rc != IIC_OK) return rc;
  const int bn = t32_pick_bn(g->cout);
  IIC_REQUIRE(bn != 0 && g->cin % T32_KB == 0, IIC_ERR_UNSUPPORTED, "tf32 wgrad: cin %% 32 == 0 and cout %% 64 == 0 required");
  IIC_REQUIRE(g->kh == g->kw, IIC_ERR_UNSUPPORTED, "tf32 wgrad: square filters only");
  T32Params P = {};
  P.rows = (long long)g->n * g->oh * g->ow;
  P.rowH = g->oh; P.rowW = g->ow; P.KH = g->kh; P.KW = g->kw; P.s = g->stride; P.d = g->dil; P.lower = -g->pad;
  const int upper = g->pad - (g->kh - 1) * g->dil;
  IIC_REQUIRE(P.lower >= -128 && upper >= -128 && upper <= 127, IIC_ERR_UNSUPPORTED, "im2col corner range");
  P.srcC = g->cin; P.Ktot = g->kh * g->kw * g->cin; P.N = g->cout;
  P.mtiles = (P.Ktot + TC_BM - 1) / TC_BM; P.ntiles = g->cout / bn;
  P.splits = t32_wgrad_splits(g);
  P.total_kb = (int)((P.rows + T32_KB - 1) / T32_KB);
  P.kb_per_split = (P.total_kb + P.splits - 1) / P.splits;
  P.partial = ws;
  alignas(64) CUtensorMap tmA, tmB;
  rc = t32_im2col_map(&tmA, x, g->n, g->h, g->w, g->cin, P.lower, P.lower, upper, upper, P.s, T32_KB);
  if (rc != IIC_OK) return rc;
  // dy [rows][cout]: 32-channel x 32-pixel boxes; each lands as [pixel][128 B] = one MN-major SWIZZLE_128B atom column
  rc = t32_tiled_map(&tmB, dy, P.rows, g->cout, T32_KB);
  if (rc != IIC_OK) return rc;
  rc = t32_launch<T32_WGRAD>(tmA, tmB, P, bn, split, st);
  if (rc != IIC_OK) return rc;
  const long long total = (long long)P.Ktot * g->cout;
  int blocks = cdiv(total, 256);
  if (blocks > device_sm_count() * 8) blocks = device_sm_count() * 8;
  wgrad_tf32_reduce_kernel<<<blocks, 256, 0, st>>>(ws, dw, P.Ktot, g->cout, P.splits);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

}  // namespace iic
