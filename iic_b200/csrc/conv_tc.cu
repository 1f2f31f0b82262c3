// tcgen05 (5th-gen tensor core) implicit-GEMM convolution for sm_100a: bf16 operands, fp32
// accumulation in TMEM.  This is the IIC_BF16 mode of iic_conv_fprop / dgrad / wgrad and carries
// >99 % of the FLOPs of the IIC training step (36.34 GFLOP per image pair at
// 96x96 through ClusterNet5g).  Replaces the cuDNN fp32 convolutions behind nn.Conv2d in
// code/archs/cluster/residual.py:4-7,:53-55 and code/archs/cluster/vgg.py:25-27.
//
// One CTA computes a 128 x BN tile of   D[M][N] = sum_K A[M][K] * B[N][K]   where
//   MODE_FPROP (also dgrad via a transposed gather):
//       M = output pixels, K = (tap, cin), N = cout.  A is gathered on the fly from the NHWC
//       activation (im2col never materialised), B is the repacked weight [cout][kh][kw][cin].
//       Both operands are K-major in shared memory.
//   MODE_WGRAD:
//       M = (tap, cin), N = cout, K = pixels (split across blockIdx.z).  A is the gathered
//       activation and B is dy, both read as "pixel rows x 64 channels" = MN-major operands, so
//       no transpose is ever performed: the UMMA descriptors carry the majorness.
//
// Pipeline (warp-specialised, 160 threads):
//   warps 0-3  producers: 16-byte cp.async gathers into a STAGES-deep ring of 128B-swizzled
//              tiles (chunk c of row r lands at chunk c ^ (r & 7): the layout a SWIZZLE_128B UMMA
//              descriptor expects), zero-fill for padding / tails; completion is published with
//              cp.async.wait_group -> fence.proxy.async -> mbarrier.arrive, LAG stages behind issue.
//              Afterwards the same warps run the epilogue.
//   warp 4     one elected thread issues tcgen05.mma (M=128, N=BN, K=16) x4 per stage;
//              tcgen05.commit releases the stage (empty barrier) and finally signals the epilogue.
//   epilogue   tcgen05.ld 32x32b (lane = row) -> registers -> bf16 (+ residual-gradient addend)
//              -> 64 B contiguous global stores per thread; wgrad: fp32 split-K partials.
#include "tc_ptx.cuh"

namespace iic {

constexpr int TC_PRODUCERS = 128;
constexpr int TC_THREADS = 160;
constexpr int TC_LAG = 2;

enum { MODE_FPROP = 0, MODE_WGRAD = 1 };

struct TcParams {
  // gathered tensor (activation for fprop/wgrad, dy for dgrad), NHWC bf16
  const __nv_bfloat16* src;
  int srcH, srcW, srcC;
  // pixel grid enumerated by the gather rows (output pixels for fprop, dx pixels for dgrad, dy pixels for wgrad)
  int rowH, rowW;
  int KH, KW, s, p, d, transposed;
  long long rows;  // number of gather rows (n * rowH * rowW)
  int Ktot;        // KH*KW*srcC
  // dense operand: fprop/dgrad: packed weight [N][Ktot]; wgrad: dy [rows][N]
  const __nv_bfloat16* dense;
  int N;           // total N (cout / cin for dgrad)
  // outputs
  __nv_bfloat16* out;            // fprop/dgrad: [rows][N]
  const __nv_bfloat16* addend;   // optional, same layout as out
  float* partial;                // wgrad: [splits][Ktot][N] fp32
  int kblocks_per_split;         // wgrad
};

template <int BN> struct TcCfg {
  static constexpr int A_BYTES = TC_BM * 128;         // 16 KB: 128 rows x 128 B (K-major) or 2 x [64 x 128 B] (MN-major)
  static constexpr int B_BYTES = BN * 128;            // BN rows x 128 B, or BN/64 x [64 x 128 B]
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 3 : 4);  // 192 / 96 / 96 KB: 1 / 2 / 2 CTAs per SM
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int MODE, int BN>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(TcParams P) {
  using Cfg = TcCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // 128B-swizzle atoms need 1024 B alignment
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + STAGES * Cfg::STAGE_BYTES;  // full[STAGES], empty[STAGES], accum, tmem slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + STAGES * Cfg::STAGE_BYTES + (2 * STAGES + 1) * 8);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  const uint32_t accum_bar = bars + 8u * (2 * STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntn = P.N / BN;
  const int n0 = (int)(blockIdx.x % ntn) * BN;
  const long long m_tile = blockIdx.x / ntn;

  // K range of this CTA (in 64-element k-blocks)
  int kb_begin, kb_end;
  long long m0;
  if (MODE == MODE_FPROP) {
    m0 = m_tile * TC_BM;
    kb_begin = 0;
    kb_end = P.Ktot / 64;
  } else {
    m0 = m_tile * TC_BM;  // index into (tap, cin)
    const int total_kb = (int)((P.rows + 63) / 64);
    kb_begin = blockIdx.z * P.kblocks_per_split;
    kb_end = min(total_kb, kb_begin + P.kblocks_per_split);
  }
  const int nk = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), TC_PRODUCERS);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  if (warp < 4) {
    // =============================== producers ===============================================
    const int t = threadIdx.x;
    const int chunk = t & 7, rsub = t >> 3;  // 16 B chunk within a 128 B row; row within a 16-row slab
    const long long planeHW = (long long)P.srcH * P.srcW;

    if (MODE == MODE_FPROP) {
      // Per-thread gather rows r = it*16 + rsub (it = 0..7).  Everything that does not depend on
      // the filter tap is hoisted: a base pointer per row plus the tap-free input coordinate; per
      // k-block only 2 adds, 2 unsigned compares and one 64-bit add remain per row.
      const __nv_bfloat16* rp[8];
      int gy[8], gx[8];
      const int smask = P.s - 1, sshift = P.s >> 1;  // stride is 1 or 2 on the tensor-core path
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const long long m = m0 + it * 16 + rsub;
        if (m < P.rows) {
          const int ox = (int)(m % P.rowW);
          const long long q = m / P.rowW;
          const int oy = (int)(q % P.rowH);
          rp[it] = P.src + ((q / P.rowH) * planeHW * P.srcC + chunk * 8);
          gy[it] = P.transposed ? oy + P.p : oy * P.s - P.p;
          gx[it] = P.transposed ? ox + P.p : ox * P.s - P.p;
        } else {
          rp[it] = P.src;
          gy[it] = gx[it] = -(1 << 28);  // fails every range check below -> zero fill
        }
      }
      const bool transposed = P.transposed != 0;
      for (int i = 0; i < nk; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t sa = base + s * Cfg::STAGE_BYTES, sb = sa + Cfg::A_BYTES;
        const int j = (kb_begin + i) * 64;
        const int tap = j / P.srcC, c0 = j - tap * P.srcC;
        const int ta = tap / P.KW, tb = tap - ta * P.KW;
        const int dyy = ta * P.d, dxx = tb * P.d;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 16 + rsub;
          int iy, ix;
          bool ok = true;
          if (!transposed) {
            iy = gy[it] + dyy;
            ix = gx[it] + dxx;
          } else {
            const int ty = gy[it] - dyy, tx = gx[it] - dxx;
            ok = ((ty | tx) & smask) == 0;
            iy = ty >> sshift;
            ix = tx >> sshift;
          }
          ok = ok && (unsigned)iy < (unsigned)P.srcH && (unsigned)ix < (unsigned)P.srcW;
          const __nv_bfloat16* src = ok ? rp[it] + ((iy * P.srcW + ix) * P.srcC + c0) : P.src;
          cp_async16_ca(sa + r * 128 + ((chunk ^ (r & 7)) << 4), src, ok ? 16u : 0u);
        }
        // weights: BN rows (cout) x 128 B
        const __nv_bfloat16* wsrc = P.dense + ((long long)(n0 + rsub) * P.Ktot + j + chunk * 8);
#pragma unroll
        for (int it = 0; it < BN / 16; ++it) {
          const int r = it * 16 + rsub;
          cp_async16_cg(sb + r * 128 + ((chunk ^ (r & 7)) << 4), wsrc + (long long)it * 16 * P.Ktot, 16u);
        }
        cp_async_commit();
        if (i >= TC_LAG) {
          cp_async_wait<TC_LAG>();
          fence_proxy_async();
          mbar_arrive(full_bar((i - TC_LAG) % STAGES));
        }
      }
    } else {
      // wgrad: k-rows are pixels of the dy grid; A atoms (64 (tap,cin) columns) x2, B atoms BN/64.
      // Each thread owns 4 pixel rows per stage; their (ox, oy, image) coordinates advance by 64
      // pixels per k-block incrementally (no divisions in the loop).
      int aoff_y[2], aoff_x[2], ac0[2];
      bool aval[2];
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const long long j = m0 + a * 64;
        aval[a] = j < P.Ktot;
        const int tap = (int)(j / P.srcC);
        ac0[a] = (int)(j - (long long)tap * P.srcC) + chunk * 8;
        const int ta = tap / P.KW;
        aoff_y[a] = ta * P.d - P.p;
        aoff_x[a] = (tap - ta * P.KW) * P.d - P.p;
      }
      int pox[4], poy[4];
      long long pimg[4];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const long long pix = (long long)kb_begin * 64 + it * 16 + rsub;
        pox[it] = (int)(pix % P.rowW);
        const long long q = pix / P.rowW;
        poy[it] = (int)(q % P.rowH);
        pimg[it] = q / P.rowH;
      }
      const int step_x = 64 % P.rowW, step_y = 64 / P.rowW;
      const long long nimg_rows = P.rows / ((long long)P.rowH * P.rowW);
      for (int i = 0; i < nk; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t sa = base + s * Cfg::STAGE_BYTES, sb = sa + Cfg::A_BYTES;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int r = it * 16 + rsub;
          const bool pv = pimg[it] < nimg_rows;
          const int oy = poy[it], ox = pox[it];
          const uint32_t roff = r * 128 + ((chunk ^ (r & 7)) << 4);
          const __nv_bfloat16* ibase = P.src + pimg[it] * planeHW * P.srcC;
#pragma unroll
          for (int a = 0; a < 2; ++a) {
            const int iy = oy * P.s + aoff_y[a], ix = ox * P.s + aoff_x[a];
            const bool ok = pv && aval[a] && (unsigned)iy < (unsigned)P.srcH && (unsigned)ix < (unsigned)P.srcW;
            const __nv_bfloat16* src = ok ? ibase + ((iy * P.srcW + ix) * P.srcC + ac0[a]) : P.src;
            cp_async16_ca(sa + a * 8192 + roff, src, ok ? 16u : 0u);
          }
          const __nv_bfloat16* dsrc =
              pv ? P.dense + (((pimg[it] * P.rowH + oy) * P.rowW + ox) * P.N + n0 + chunk * 8) : P.dense;
#pragma unroll
          for (int b = 0; b < BN / 64; ++b) cp_async16_cg(sb + b * 8192 + roff, pv ? dsrc + b * 64 : P.dense, pv ? 16u : 0u);
          // advance this row by 64 pixels
          pox[it] += step_x;
          poy[it] += step_y;
          if (pox[it] >= P.rowW) { pox[it] -= P.rowW; poy[it] += 1; }
          while (poy[it] >= P.rowH) { poy[it] -= P.rowH; pimg[it] += 1; }
        }
        cp_async_commit();
        if (i >= TC_LAG) {
          cp_async_wait<TC_LAG>();
          fence_proxy_async();
          mbar_arrive(full_bar((i - TC_LAG) % STAGES));
        }
      }
    }
    // drain the last (up to TC_LAG) stages
    if (nk >= 2) {
      cp_async_wait<1>();
      fence_proxy_async();
      mbar_arrive(full_bar((nk - 2) % STAGES));
    }
    if (nk >= 1) {
      cp_async_wait<0>();
      fence_proxy_async();
      mbar_arrive(full_bar((nk - 1) % STAGES));
    }

    // =============================== epilogue ================================================
    if (nk > 0) {
      mbar_wait(accum_bar, 0);
      tc_fence_after();
    }
    const int row = warp * 32 + lane;  // TMEM lane == tile row
    const long long m = m0 + row;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      if (nk > 0) {
        tmem_ld32(tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int q = 0; q < 32; ++q) v[q] = 0u;
      }
      if (MODE == MODE_FPROP) {
        if (m < P.rows) {
          __nv_bfloat16* o = P.out + m * P.N + n0 + c0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[q * 8 + e]);
            if (P.addend != nullptr) {
              float ad[8];
              load8(P.addend + m * P.N + n0 + c0 + q * 8, ad);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += ad[e];
            }
            store8(o + q * 8, f);
          }
        }
      } else {
        if (m < P.Ktot) {
          float* o = P.partial + ((long long)blockIdx.z * P.Ktot + m) * P.N + n0 + c0;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(o + q * 4) = make_float4(__uint_as_float(v[q * 4]), __uint_as_float(v[q * 4 + 1]),
                                                                __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
        }
      }
    }
  } else {
    // =============================== MMA issuer ==============================================
    constexpr uint32_t idesc = (MODE == MODE_FPROP) ? make_idesc(BN, 0, 0) : make_idesc(BN, 1, 1);
    for (int i = 0; i < nk; ++i) {
      const int s = i % STAGES;
      const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = base + s * Cfg::STAGE_BYTES, sb = sa + Cfg::A_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint64_t ad, bd;
          if (MODE == MODE_FPROP) {
            ad = make_desc(sa + kk * 32, 16, 1024);
            bd = make_desc(sb + kk * 32, 16, 1024);
          } else {
            ad = make_desc(sa + kk * 2048, 8192, 1024);
            bd = make_desc(sb + kk * 2048, 8192, 1024);
          }
          umma_bf16(tmem_acc, ad, bd, idesc, (i > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
        if (i == nk - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_acc, Cfg::TMEM_COLS);
}

// dw_packed[co][j] (=) sum_z partial[z][j][co]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int Ktot, int N, int splits) {
  const long long total = (long long)Ktot * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % N);
    const int j = (int)(i / N);
    float t = 0.f;
    for (int z = 0; z < splits; ++z) t += partial[(long long)z * total + i];
    dw[(long long)co * Ktot + j] = t;
  }
}

template <int MODE, int BN>
static int launch_tc(const TcParams& P, dim3 grid, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  IIC_CUDA(cudaFuncSetAttribute(conv_tc_kernel<MODE, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  conv_tc_kernel<MODE, BN><<<grid, TC_THREADS, Cfg::SMEM, st>>>(P);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

static int pick_bn(int N) {
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  return 0;
}

int tc_conv_gather_gemm(const __nv_bfloat16* src, int srcH, int srcW, int srcC, int rowH, int rowW, int nimg,
                        const iic_conv_geom* g, int transposed, const __nv_bfloat16* wpacked, int N,
                        const __nv_bfloat16* addend, __nv_bfloat16* out, cudaStream_t st) {
  IIC_REQUIRE(srcC % 64 == 0, IIC_ERR_UNSUPPORTED, "tcgen05 conv: gathered channels (%d) must be a multiple of 64", srcC);
  IIC_REQUIRE(g->stride == 1 || g->stride == 2, IIC_ERR_UNSUPPORTED, "tcgen05 conv: stride %d (only 1 or 2)", g->stride);
  const int bn = pick_bn(N);
  IIC_REQUIRE(bn != 0, IIC_ERR_UNSUPPORTED, "tcgen05 conv: N=%d must be a multiple of 64", N);
  TcParams P = {};
  P.src = src; P.srcH = srcH; P.srcW = srcW; P.srcC = srcC; P.rowH = rowH; P.rowW = rowW;
  P.KH = g->kh; P.KW = g->kw; P.s = g->stride; P.p = g->pad; P.d = g->dil; P.transposed = transposed;
  P.rows = (long long)nimg * rowH * rowW;
  P.Ktot = g->kh * g->kw * srcC;
  P.dense = wpacked; P.N = N; P.out = out; P.addend = addend;
  dim3 grid((unsigned)((N / bn) * ((P.rows + TC_BM - 1) / TC_BM)), 1, 1);
  switch (bn) {
    case 256: return launch_tc<MODE_FPROP, 256>(P, grid, st);
    case 128: return launch_tc<MODE_FPROP, 128>(P, grid, st);
    default: return launch_tc<MODE_FPROP, 64>(P, grid, st);
  }
}

static int tc_wgrad_splits(const iic_conv_geom* g) {
  const long long rows = (long long)g->n * g->oh * g->ow;
  const int total_kb = (int)((rows + 63) / 64);
  const int Ktot = g->kh * g->kw * g->cin;
  const int bn = pick_bn(g->cout);
  const long long tiles = (long long)((Ktot + TC_BM - 1) / TC_BM) * (g->cout / (bn ? bn : 64));
  long long want = ((long long)device_sm_count() * 2 + tiles - 1) / tiles;
  if (want > total_kb / 4) want = total_kb / 4;  // at least 4 k-blocks per CTA
  if (want < 1) want = 1;
  if (want > 256) want = 256;
  return (int)want;
}

long long tc_conv_wgrad_workspace(const iic_conv_geom* g) {
  return (long long)tc_wgrad_splits(g) * g->kh * g->kw * g->cin * g->cout * (long long)sizeof(float);
}

int tc_conv_wgrad(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, float* ws, const iic_conv_geom* g,
                  cudaStream_t st) {
  IIC_REQUIRE(g->cin % 64 == 0, IIC_ERR_UNSUPPORTED, "tcgen05 wgrad: cin=%d must be a multiple of 64", g->cin);
  const int bn = pick_bn(g->cout);
  IIC_REQUIRE(bn != 0, IIC_ERR_UNSUPPORTED, "tcgen05 wgrad: cout=%d must be a multiple of 64", g->cout);
  TcParams P = {};
  P.src = x; P.srcH = g->h; P.srcW = g->w; P.srcC = g->cin; P.rowH = g->oh; P.rowW = g->ow;
  P.KH = g->kh; P.KW = g->kw; P.s = g->stride; P.p = g->pad; P.d = g->dil; P.transposed = 0;
  P.rows = (long long)g->n * g->oh * g->ow;
  P.Ktot = g->kh * g->kw * g->cin;
  P.dense = dy; P.N = g->cout; P.partial = ws;
  const int splits = tc_wgrad_splits(g);
  const int total_kb = (int)((P.rows + 63) / 64);
  P.kblocks_per_split = (total_kb + splits - 1) / splits;
  dim3 grid((g->cout / bn) * ((P.Ktot + TC_BM - 1) / TC_BM), 1, splits);
  int rc;
  switch (bn) {
    case 256: rc = launch_tc<MODE_WGRAD, 256>(P, grid, st); break;
    case 128: rc = launch_tc<MODE_WGRAD, 128>(P, grid, st); break;
    default: rc = launch_tc<MODE_WGRAD, 64>(P, grid, st); break;
  }
  if (rc != IIC_OK) return rc;
  const long long total = (long long)P.Ktot * g->cout;
  int blocks = cdiv(total, 256);
  if (blocks > device_sm_count() * 8) blocks = device_sm_count() * 8;
  wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(ws, dw, P.Ktot, g->cout, splits);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

}  // namespace iic
