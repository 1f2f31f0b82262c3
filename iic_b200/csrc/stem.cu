// Stem convolution (first conv of every IIC net: net5g.py:21-23 cin=2 (sobel dx,dy) 3x3;
// vgg.py/net6c cin=1 5x5; net10a cin=5 3x3).  K = cin*kh*kw is 18..45: far too small for a
// tensor-core tile to matter and the op is bound by writing the 64-channel output, so it is a
// direct SIMT convolution that reads the reference's NCHW fp32 input and writes NHWC
// activations (fp32 or bf16) with 16/32 B vector stores.  No dgrad (the input needs no gradient).
// wgrad is a tall-skinny Gram product  dW[co][k] = sum_p dy[p][co] * patch[p][k]  done with 4x4
// register tiles out of shared memory, per-block partials and a deterministic reduction.
#include "common.cuh"

namespace iic {

constexpr int STEM_MAXK = 128;

template <typename T>
__global__ void __launch_bounds__(256) stem_fprop_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         T* __restrict__ y, iic_conv_geom g) {
  extern __shared__ __align__(16) float ws[];  // [K][cout]
  const int K = g.cin * g.kh * g.kw, cout = g.cout;
  for (int i = threadIdx.x; i < K * cout; i += blockDim.x) {
    const int co = i % cout, k = i / cout;
    ws[i] = w[(long long)co * K + k];
  }
  __syncthreads();
  const int groups = cout >> 3, ppb = 256 / groups;
  const int cgi = threadIdx.x % groups, pl = threadIdx.x / groups;
  const long long P = (long long)g.n * g.oh * g.ow;
  if (pl >= ppb) return;
  for (long long p = (long long)blockIdx.x * ppb + pl; p < P; p += (long long)gridDim.x * ppb) {
    const int ox = (int)(p % g.ow);
    const int oy = (int)((p / g.ow) % g.oh);
    const int n = (int)(p / ((long long)g.ow * g.oh));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    int k = 0;
    for (int ci = 0; ci < g.cin; ++ci) {
      const float* xc = x + ((long long)n * g.cin + ci) * g.h * g.w;
      for (int a = 0; a < g.kh; ++a) {
        const int iy = oy * g.stride - g.pad + a * g.dil;
        for (int b = 0; b < g.kw; ++b, ++k) {
          const int ix = ox * g.stride - g.pad + b * g.dil;
          const float v = (iy >= 0 && iy < g.h && ix >= 0 && ix < g.w) ? __ldg(xc + (long long)iy * g.w + ix) : 0.f;
          const float4 w0 = *reinterpret_cast<const float4*>(ws + k * cout + cgi * 8);
          const float4 w1 = *reinterpret_cast<const float4*>(ws + k * cout + cgi * 8 + 4);
          acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]);
          acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
          acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]);
          acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
        }
      }
    }
    store8(y + p * cout + cgi * 8, acc);
  }
}

// cout == 64 fast path: one thread per output pixel (a warp reads 32 consecutive pixels of the NCHW input:
// coalesced), all 64 output channels in registers, weights broadcast from shared memory
// (1 LDS.128 per 4 FMA, FMA-bound), one 128 B (bf16) / 256 B (fp32) contiguous store per thread.
template <typename T>
__global__ void __launch_bounds__(128) stem_fprop64_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           T* __restrict__ y, iic_conv_geom g) {
  extern __shared__ __align__(16) float ws[];  // [K][64]
  const int K = g.cin * g.kh * g.kw;
  for (int i = threadIdx.x; i < K * 64; i += blockDim.x) ws[i] = w[(long long)(i & 63) * K + (i >> 6)];
  __syncthreads();
  const long long P = (long long)g.n * g.oh * g.ow;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(p % g.ow);
    const int oy = (int)((p / g.ow) % g.oh);
    const int n = (int)(p / ((long long)g.ow * g.oh));
    float acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.f;
    int k = 0;
    for (int ci = 0; ci < g.cin; ++ci) {
      const float* xc = x + ((long long)n * g.cin + ci) * g.h * g.w;
      for (int a = 0; a < g.kh; ++a) {
        const int iy = oy * g.stride - g.pad + a * g.dil;
        for (int b = 0; b < g.kw; ++b, ++k) {
          const int ix = ox * g.stride - g.pad + b * g.dil;
          const float v = (iy >= 0 && iy < g.h && ix >= 0 && ix < g.w) ? __ldg(xc + (long long)iy * g.w + ix) : 0.f;
          const float4* wk = reinterpret_cast<const float4*>(ws + k * 64);
#pragma unroll
          for (int c4 = 0; c4 < 16; ++c4) {
            const float4 w4 = wk[c4];
            acc[c4 * 4] = fmaf(v, w4.x, acc[c4 * 4]);
            acc[c4 * 4 + 1] = fmaf(v, w4.y, acc[c4 * 4 + 1]);
            acc[c4 * 4 + 2] = fmaf(v, w4.z, acc[c4 * 4 + 2]);
            acc[c4 * 4 + 3] = fmaf(v, w4.w, acc[c4 * 4 + 3]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = acc[q * 8 + e];
      store8(y + p * 64 + q * 8, f);
    }
  }
}

// cout == 64, stride 1, dilation 1, ow % 4 == 0: a thread computes 4 consecutive output pixels x 16 channels
// (64 FMA per 4 LDS.128 and per ~2 global loads; the one-pixel-per-thread kernel above issues one LDS.128 per
// 4 FMA and is bound by the shared-memory pipe).  Two thread layouts:
//   ILV = 0  warp = 32 consecutive pixel quads x one channel quarter (16 consecutive channels per thread): every
//            store instruction of a warp writes 16 B pieces 512 B apart -- half sectors;
//   ILV = 1  warp = 8 consecutive pixel quads x 4 channel slices; a thread owns channels [8s, 8s+8) and [32+8s, 32+8s+8),
//            so the four lanes of a quad write 64 (bf16) / 128 (fp32) contiguous bytes per store instruction: whole
//            sectors.  The output write (1.7 GB per step at the bench shape) is what bounds this kernel.
// The weight reads are warp broadcasts in both (ILV = 1: four distinct 16 B chunks per LDS.128, conflict free).
// Optionally accumulates the BatchNorm batch statistics of the fp32 results (per-thread sums over its pixels,
// shuffle reduction at the end, one partial row per block in the layout of the tcgen05 conv epilogue:
// [block][2 view slots][{sum, sum of squares}][64]); the grid is split evenly between the `views` stacked batches.
template <typename T, int KS, int ILV>
__global__ void __launch_bounds__(256, 2) stem_fprop64q_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                               T* __restrict__ y, iic_conv_geom g,
                                                               float* __restrict__ stat_partial, int views) {
  extern __shared__ __align__(16) float ws[];  // [K][64]
  __shared__ float red[8][4][32];              // [warp][channel slice][{sum, sum of squares} x 16]
  const int K = g.cin * KS * KS;
  for (int i = threadIdx.x; i < K * 64; i += blockDim.x) ws[i] = w[(long long)(i & 63) * K + (i >> 6)];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cq = ILV ? (lane & 3) : (warp & 3);                                // channel slice
  const int qslot = ILV ? warp * 8 + (lane >> 2) : (warp >> 2) * 32 + lane;    // quad of the block's 64 per iteration
  const int cb0 = ILV ? cq * 8 : cq * 16, cb1 = ILV ? 32 + cq * 8 : cq * 16 + 8;  // the thread's two 8-channel groups
  const int qpr = g.ow >> 2;
  const long long Qv = (long long)(g.n / views) * g.oh * qpr;  // quads per view
  const int Gv = gridDim.x / views, v = blockIdx.x / Gv, lb = blockIdx.x % Gv;
  float s1[16], s2[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) s1[c] = s2[c] = 0.f;
  for (long long ql = (long long)lb * 64 + qslot; ql < Qv; ql += (long long)Gv * 64) {
    const long long q = (long long)v * Qv + ql;
    int ox0, oy, n;
    if ((long long)views * Qv <= 0xffffffffll) {  // 32-bit divisions
      const unsigned u = (unsigned)q, t = u / (unsigned)qpr;
      ox0 = (int)(u - t * (unsigned)qpr) * 4;
      n = (int)(t / (unsigned)g.oh);
      oy = (int)(t - (unsigned)n * (unsigned)g.oh);
    } else {
      ox0 = (int)(q % qpr) * 4;
      const long long t = q / qpr;
      oy = (int)(t % g.oh);
      n = (int)(t / g.oh);
    }
    float acc[4][16];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 16; ++c) acc[j][c] = 0.f;
    for (int ci = 0; ci < g.cin; ++ci) {
      const float* xc = x + ((long long)n * g.cin + ci) * g.h * g.w;
#pragma unroll
      for (int a = 0; a < KS; ++a) {
        const int iy = oy - g.pad + a;
        const bool rowok = (unsigned)iy < (unsigned)g.h;
        float xv[KS + 3];
#pragma unroll
        for (int j = 0; j < KS + 3; ++j) {
          const int ix = ox0 - g.pad + j;
          xv[j] = (rowok && (unsigned)ix < (unsigned)g.w) ? __ldg(xc + (long long)iy * g.w + ix) : 0.f;
        }
#pragma unroll
        for (int b = 0; b < KS; ++b) {
          const float* wk = ws + ((ci * KS + a) * KS + b) * 64;
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float4 w4 = *reinterpret_cast<const float4*>(wk + (c4 < 2 ? cb0 + c4 * 4 : cb1 + (c4 - 2) * 4));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              acc[j][c4 * 4] = fmaf(xv[j + b], w4.x, acc[j][c4 * 4]);
              acc[j][c4 * 4 + 1] = fmaf(xv[j + b], w4.y, acc[j][c4 * 4 + 1]);
              acc[j][c4 * 4 + 2] = fmaf(xv[j + b], w4.z, acc[j][c4 * 4 + 2]);
              acc[j][c4 * 4 + 3] = fmaf(xv[j + b], w4.w, acc[j][c4 * 4 + 3]);
            }
          }
        }
      }
    }
    const long long p0 = ((long long)n * g.oh + oy) * g.ow + ox0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f[8];
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
        for (int e = 0; e < 8; ++e) f[e] = acc[j][h2 * 8 + e];
        store8(y + (p0 + j) * 64 + (h2 ? cb1 : cb0), f);
      }
      if (stat_partial != nullptr) {
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          s1[c] += acc[j][c];
          s2[c] = fmaf(acc[j][c], acc[j][c], s2[c]);
        }
      }
    }
  }
  if (stat_partial == nullptr) return;
  // lanes that share a channel slice: all 32 (ILV = 0) or those with equal lane & 3 (ILV = 1)
#pragma unroll
  for (int c = 0; c < 16; ++c) {
#pragma unroll
    for (int o = 16; o >= (ILV ? 4 : 1); o >>= 1) {
      s1[c] += __shfl_xor_sync(0xffffffffu, s1[c], o);
      s2[c] += __shfl_xor_sync(0xffffffffu, s2[c], o);
    }
  }
  for (int i = threadIdx.x; i < 8 * 4 * 32; i += blockDim.x) (&red[0][0][0])[i] = 0.f;
  __syncthreads();
  if (ILV ? (lane < 4) : (lane == 0)) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      red[warp][cq][c] = s1[c];
      red[warp][cq][16 + c] = s2[c];
    }
  }
  __syncthreads();
  if (threadIdx.x < 256) {  // (slot, stat, channel)
    const int ch = threadIdx.x & 63, q2 = (threadIdx.x >> 6) & 1, slot = threadIdx.x >> 7;
    // channel -> (slice, index among the thread's 16 accumulators)
    const int sl = ILV ? ((ch & 31) >> 3) : (ch >> 4);
    const int c = ILV ? ((ch >> 5) * 8 + (ch & 7)) : (ch & 15);
    float val = 0.f;
    if (slot == v) {
#pragma unroll
      for (int wp = 0; wp < 8; ++wp) val += red[wp][sl][q2 * 16 + c];
    }
    stat_partial[(long long)blockIdx.x * 256 + threadIdx.x] = val;
  }
}

constexpr int SW_PC = 64;  // pixels per chunk

template <typename T>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ dy,
                                                         float* __restrict__ partial, iic_conv_geom g, int Kp, int tkk,
                                                         int T_tiles, int G) {
  extern __shared__ __align__(16) float sm[];
  const int K = g.cin * g.kh * g.kw, cout = g.cout;
  float* dys = sm;                   // [SW_PC][cout]
  float* pat = dys + SW_PC * cout;   // [SW_PC][Kp]
  float* red = pat + SW_PC * Kp;     // [G][cout*Kp] (used at the end)
  int* pixb = reinterpret_cast<int*>(red + (size_t)G * cout * Kp);  // [SW_PC][3]: image base offset, iy0, ix0
  int* ktab = pixb + SW_PC * 3;                                     // [Kp][3]: plane offset, a*dil, b*dil
  const int P = g.n * g.oh * g.ow;  // (host guarantees < 2^31)
  const int tid = threadIdx.x;
  const int my_g = tid / T_tiles, my_t = tid % T_tiles;
  const bool active = my_g < G;
  const int tc = my_t / tkk, tk = my_t % tkk;  // tile = 4 output channels x 4 taps
  const int hw = g.h * g.w;
  for (int k = tid; k < Kp; k += 256) {
    const int b = k % g.kw, a = (k / g.kw) % g.kh, ci = k / (g.kw * g.kh);
    ktab[k * 3] = k < K ? ci * hw : -1;
    ktab[k * 3 + 1] = a * g.dil;
    ktab[k * 3 + 2] = b * g.dil;
  }
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int nchunks = (P + SW_PC - 1) / SW_PC;
  for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const int p0 = ch * SW_PC;
    __syncthreads();
    if (tid < SW_PC) {
      const int p = p0 + tid;
      if (p < P) {
        const int ox = p % g.ow;
        const int q = p / g.ow;
        const int oy = q % g.oh;
        const int n = q / g.oh;
        pixb[tid * 3] = n * g.cin * hw;
        pixb[tid * 3 + 1] = oy * g.stride - g.pad;
        pixb[tid * 3 + 2] = ox * g.stride - g.pad;
      } else {
        pixb[tid * 3] = -1;
      }
    }
    // dy chunk: contiguous [SW_PC][cout] rows
    {
      const long long base = (long long)p0 * cout;
      const int count = (P - p0 < SW_PC ? P - p0 : SW_PC) * cout;
      for (int i = tid * 8; i < SW_PC * cout; i += 256 * 8) {  // 8 channels (16 / 32 B) per load; cout % 8 == 0
        float f[8];
        if (i < count) {
          load8(dy + base + i, f);
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) f[e] = 0.f;
        }
        *reinterpret_cast<float4*>(dys + i) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(dys + i + 4) = make_float4(f[4], f[5], f[6], f[7]);
      }
    }
    __syncthreads();
    for (int i = tid; i < SW_PC * Kp; i += 256) {
      const int r = i % SW_PC, k = i / SW_PC;  // consecutive threads -> consecutive pixels (coalesced in NCHW)
      float v = 0.f;
      const int pb = pixb[r * 3], kb = ktab[k * 3];
      if (pb >= 0 && kb >= 0) {
        const int iy = pixb[r * 3 + 1] + ktab[k * 3 + 1], ix = pixb[r * 3 + 2] + ktab[k * 3 + 2];
        if ((unsigned)iy < (unsigned)g.h && (unsigned)ix < (unsigned)g.w) v = __ldg(x + (long long)pb + kb + iy * g.w + ix);
      }
      pat[r * Kp + k] = v;
    }
    __syncthreads();
    if (active) {
      for (int r = my_g; r < SW_PC; r += G) {
        const float4 a = *reinterpret_cast<const float4*>(dys + r * cout + 4 * tc);
        const float4 b = *reinterpret_cast<const float4*>(pat + r * Kp + 4 * tk);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[(long long)my_g * cout * Kp + (4 * tc + i) * Kp + 4 * tk + j] = acc[i][j];
  }
  __syncthreads();
  for (int e = tid; e < cout * K; e += 256) {
    const int co = e / K, k = e % K;
    float t = 0.f;
    for (int gg = 0; gg < G; ++gg) t += red[(long long)gg * cout * Kp + co * Kp + k];
    partial[(long long)blockIdx.x * cout * K + e] = t;
  }
}

__global__ void stem_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ grad, int count,
                                         int nblk, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float t = 0.f;
  for (int b = 0; b < nblk; ++b) t += partial[(long long)b * count + i];
  grad[i] = accumulate ? grad[i] + t : t;
}

}  // namespace iic

using namespace iic;

static int stem_check(const iic_conv_geom* g, const char* who) {
  IIC_REQUIRE(g != nullptr, IIC_ERR_BAD_ARG, "%s: null geometry", who);
  const int K = g->cin * g->kh * g->kw;
  IIC_REQUIRE(K > 0 && K <= STEM_MAXK, IIC_ERR_UNSUPPORTED, "%s: cin*kh*kw=%d exceeds %d", who, K, STEM_MAXK);
  IIC_REQUIRE(g->cout % 8 == 0 && 256 % (g->cout / 8) == 0 && g->cout <= 256, IIC_ERR_UNSUPPORTED,
              "%s: cout=%d unsupported", who, g->cout);
  IIC_REQUIRE(g->oh == (g->h + 2 * g->pad - g->dil * (g->kh - 1) - 1) / g->stride + 1 &&
                  g->ow == (g->w + 2 * g->pad - g->dil * (g->kw - 1) - 1) / g->stride + 1,
              IIC_ERR_BAD_ARG, "%s: inconsistent output size", who);
  return IIC_OK;
}

static bool stem_quad_ok(const iic_conv_geom* g) {
  return option(OPT_STEM_QUAD) != 0 && g->cout == 64 && g->stride == 1 && g->dil == 1 && g->ow % 4 == 0 && g->kh == g->kw && (g->kh == 3 || g->kh == 5);
}
static int stem_quad_blocks(const iic_conv_geom* g, int views) {
  const long long quads = (long long)g->n * g->oh * (g->ow / 4);
  long long blocks = (quads / views + 63) / 64;
  const long long cap = (long long)device_sm_count() * 4;  // 2 resident blocks per SM, two rounds
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks * views;
}
template <typename T>
static int stem_quad_launch(const float* x, const float* w, T* y, const iic_conv_geom* g, float* stat_partial, int views,
                            cudaStream_t st) {
  const size_t smem = (size_t)g->cin * g->kh * g->kw * 64 * sizeof(float);
  const int blocks = stem_quad_blocks(g, views);
  const bool ilv = option(OPT_STEM_QUAD) >= 2;
  if (g->kh == 3) {
    if (ilv)
      stem_fprop64q_kernel<T, 3, 1><<<blocks, 256, smem, st>>>(x, w, y, *g, stat_partial, views);
    else
      stem_fprop64q_kernel<T, 3, 0><<<blocks, 256, smem, st>>>(x, w, y, *g, stat_partial, views);
  } else {
    if (ilv)
      stem_fprop64q_kernel<T, 5, 1><<<blocks, 256, smem, st>>>(x, w, y, *g, stat_partial, views);
    else
      stem_fprop64q_kernel<T, 5, 0><<<blocks, 256, smem, st>>>(x, w, y, *g, stat_partial, views);
  }
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_stem_fprop_stats_blocks(const iic_conv_geom* g, int dtype, int views) {
  if (g == nullptr || !stem_quad_ok(g) || (dtype != IIC_F32 && dtype != IIC_BF16) || views < 1 || views > 2 ||
      g->n % views != 0)
    return 0;
  return stem_quad_blocks(g, views);
}

extern "C" int iic_stem_fprop_stats(const float* x_nchw, const float* w_oihw, void* y, const iic_conv_geom* g, int dtype,
                                    int views, float* stat_partial, void* stream) {
  int rc = stem_check(g, "iic_stem_fprop_stats");
  if (rc != IIC_OK) return rc;
  IIC_REQUIRE(x_nchw && w_oihw && y && stat_partial, IIC_ERR_BAD_ARG, "iic_stem_fprop_stats: null pointer");
  IIC_REQUIRE(iic_stem_fprop_stats_blocks(g, dtype, views) > 0, IIC_ERR_UNSUPPORTED,
              "iic_stem_fprop_stats: needs cout 64, stride 1, dilation 1, 3x3 or 5x5, ow %% 4 == 0, 1 or 2 views");
  if (dtype == IIC_F32) return stem_quad_launch<float>(x_nchw, w_oihw, (float*)y, g, stat_partial, views, (cudaStream_t)stream);
  return stem_quad_launch<__nv_bfloat16>(x_nchw, w_oihw, (__nv_bfloat16*)y, g, stat_partial, views, (cudaStream_t)stream);
}

extern "C" int iic_stem_fprop(const float* x_nchw, const float* w_oihw, void* y, const iic_conv_geom* g, int dtype,
                              void* stream) {
  int rc = stem_check(g, "iic_stem_fprop");
  if (rc != IIC_OK) return rc;
  IIC_REQUIRE(x_nchw && w_oihw && y, IIC_ERR_BAD_ARG, "iic_stem_fprop: null pointer");
  if (stem_quad_ok(g) && (dtype == IIC_F32 || dtype == IIC_BF16)) {
    if (dtype == IIC_F32) return stem_quad_launch<float>(x_nchw, w_oihw, (float*)y, g, nullptr, 1, (cudaStream_t)stream);
    return stem_quad_launch<__nv_bfloat16>(x_nchw, w_oihw, (__nv_bfloat16*)y, g, nullptr, 1, (cudaStream_t)stream);
  }
  const int K = g->cin * g->kh * g->kw;
  const size_t smem = (size_t)K * g->cout * sizeof(float);
  const int ppb = 256 / (g->cout / 8);
  const long long P = (long long)g->n * g->oh * g->ow;
  long long blocks = (P + ppb - 1) / ppb;
  const long long cap = (long long)device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  if (g->cout == 64 && (dtype == IIC_F32 || dtype == IIC_BF16)) {
    long long b64 = (P + 127) / 128;
    const long long cap64 = (long long)device_sm_count() * 16;
    if (b64 > cap64) b64 = cap64;
    if (dtype == IIC_F32)
      stem_fprop64_kernel<float><<<(int)b64, 128, smem, st>>>(x_nchw, w_oihw, (float*)y, *g);
    else
      stem_fprop64_kernel<__nv_bfloat16><<<(int)b64, 128, smem, st>>>(x_nchw, w_oihw, (__nv_bfloat16*)y, *g);
    IIC_LAUNCH_CHECK();
    count_launch();
    return IIC_OK;
  }
  if (dtype == IIC_F32) {
    IIC_CUDA(cudaFuncSetAttribute(stem_fprop_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stem_fprop_kernel<float><<<(int)blocks, 256, smem, st>>>(x_nchw, w_oihw, (float*)y, *g);
  } else if (dtype == IIC_BF16) {
    IIC_CUDA(cudaFuncSetAttribute(stem_fprop_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stem_fprop_kernel<__nv_bfloat16><<<(int)blocks, 256, smem, st>>>(x_nchw, w_oihw, (__nv_bfloat16*)y, *g);
  } else {
    set_error("iic_stem_fprop: bad dtype");
    return IIC_ERR_BAD_ARG;
  }
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_stem_wgrad(const float* x_nchw, const void* dy, float* grad_oihw, int accumulate, void* workspace,
                              long long workspace_bytes, const iic_conv_geom* g, int dtype, void* stream) {
  int rc = stem_check(g, "iic_stem_wgrad");
  if (rc != IIC_OK) return rc;
  IIC_REQUIRE(x_nchw && dy && grad_oihw && workspace, IIC_ERR_BAD_ARG, "iic_stem_wgrad: null pointer");
  IIC_REQUIRE(g->cout % 8 == 0, IIC_ERR_UNSUPPORTED, "iic_stem_wgrad: cout must be a multiple of 8");
  const int K = g->cin * g->kh * g->kw, Kp = (K + 3) & ~3, tkk = Kp / 4;
  const int T_tiles = (g->cout / 4) * tkk;
  IIC_REQUIRE(T_tiles <= 256, IIC_ERR_UNSUPPORTED, "iic_stem_wgrad: cout*K too large (%d tiles)", T_tiles);
  int G = 256 / T_tiles;
  if (G > SW_PC) G = SW_PC;
  const long long per_block = (long long)g->cout * K * sizeof(float);
  long long nblk = (long long)device_sm_count() * 4;
  const long long P = (long long)g->n * g->oh * g->ow;
  IIC_REQUIRE(P < (1ll << 31) && (long long)g->n * g->cin * g->h * g->w < (1ll << 31), IIC_ERR_UNSUPPORTED,
              "iic_stem_wgrad: more than 2^31 pixels");
  const long long nchunks = (P + SW_PC - 1) / SW_PC;
  if (nblk > nchunks) nblk = nchunks;
  if (nblk > workspace_bytes / per_block) nblk = workspace_bytes / per_block;
  IIC_REQUIRE(nblk >= 1, IIC_ERR_BAD_ARG, "iic_stem_wgrad: workspace too small (%lld B, need >= %lld)", workspace_bytes,
              per_block);
  const size_t smem = (size_t)(SW_PC * g->cout + SW_PC * Kp + (size_t)G * g->cout * Kp) * sizeof(float) +
                      (size_t)(SW_PC * 3 + Kp * 3) * sizeof(int);
  IIC_REQUIRE(smem <= 200 * 1024, IIC_ERR_UNSUPPORTED, "iic_stem_wgrad: needs %zu B smem", smem);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == IIC_F32) {
    IIC_CUDA(cudaFuncSetAttribute(stem_wgrad_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stem_wgrad_kernel<float><<<(int)nblk, 256, smem, st>>>(x_nchw, (const float*)dy, (float*)workspace, *g, Kp, tkk, T_tiles, G);
  } else if (dtype == IIC_BF16) {
    IIC_CUDA(cudaFuncSetAttribute(stem_wgrad_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    stem_wgrad_kernel<__nv_bfloat16><<<(int)nblk, 256, smem, st>>>(x_nchw, (const __nv_bfloat16*)dy, (float*)workspace, *g, Kp, tkk, T_tiles, G);
  } else {
    set_error("iic_stem_wgrad: bad dtype");
    return IIC_ERR_BAD_ARG;
  }
  IIC_LAUNCH_CHECK();
  count_launch();
  const int count = g->cout * K;
  stem_wgrad_reduce_kernel<<<cdiv(count, 256), 256, 0, st>>>((const float*)workspace, grad_oihw, count, (int)nblk, accumulate);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}
