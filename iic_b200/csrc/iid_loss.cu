// Fused clustering IIC objective: joint P = sym(sum_n z_n z'_n^T)/sum -> MI -> analytic backward.
//
// Replaces xu-ji/IIC code/utils/cluster/IID_losses.py:6-33 (IID_loss) and :36-47 (compute_joint)
// -- ~45 tiny torch kernels per call and a (bn,k,k) intermediate -- with ONE launch:
//
//   grid = (C, S): one thread-block cluster of C CTAs per sub-head (C=8 on one die's SMs).
//   1. each CTA streams its contiguous slice of rows of Z, Z' (coalesced float4) through shared
//      memory and accumulates a k x k partial outer product in registers (4x4 register tiles,
//      row-groups when k is small), reduced deterministically in shared memory;
//   2. the C partials are reduce-scattered / all-gathered through distributed shared memory
//      (fixed order => bitwise reproducible, no atomics);
//   3. every CTA evaluates P, the marginals, the three clamps of :17-19, loss and loss_no_lamb,
//      then G = dloss/dP with the reference's clamp semantics and H = sym((G - <G,P>)/s);
//   4. second sweep over the CTA's rows: dZ = Z' H^T, dZ' = Z H  (H is symmetric, so both reads
//      of H are bank-conflict free).
//
// Multi-GPU: phase PARTIAL stops after step 2 and writes the raw joint; after an
// NCCL all-reduce of [S,k,k], phase FINISH resumes at step 3 on every rank.
//
// Algorithmic HBM bytes per call: read Z,Z' twice + write dZ,dZ' = 6*S*n*k*4 B (0.84 MB at
// n=704,k=10,S=5); the kernel is launch-latency bound, not bandwidth bound.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace iic {

constexpr int IID_THREADS = 512;
constexpr int IID_ROWS = 32;      // rows staged per chunk
constexpr int IID_MAX_TILES = 3;  // 4x4 tiles per thread => k <= 4*sqrt(3*512) ~ 156

struct IidParams {
  const float* z;
  const float* zt;
  float* loss;
  float* dz;
  float* dzt;
  float* joint_ws;
  float* joint_out;
  float* h_out;      // optional [S][k][k]: H = d loss / d (raw joint)   (segmentation losses)
  int n, k, kp, tk, ntiles, groups;
  float lamb, eps;
  int phase;
  int detached;      // 1: the normaliser carries no gradient (collapsed seg loss, seg IID_losses.py:60)
};

__device__ __forceinline__ void stage_rows(const float* __restrict__ src, float* dst, int row0, int rows, int n,
                                           int k, int kp) {
  // dst[r][kp] <- src[row0 + r][k], zero padded
  for (int i = threadIdx.x; i < rows * kp; i += blockDim.x) {
    int r = i / kp, c = i - r * kp;
    int gr = row0 + r;
    dst[i] = (c < k && gr < n) ? src[(size_t)gr * k + c] : 0.f;
  }
}

__global__ void __launch_bounds__(IID_THREADS, 1) iid_loss_kernel(IidParams p) {
  extern __shared__ __align__(16) float smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int C = cluster.num_blocks();
  const int rank = cluster.block_rank();
  const int s = blockIdx.y;
  const int tid = threadIdx.x;
  const int k = p.k, kp = p.kp, kk = kp * kp;

  // shared carve-up
  float* bufA = smem;                                    // kk
  float* bufB = bufA + kk;                               // max(kk, groups*kk)
  const int bsz = max(kk, p.groups * kk);
  float* zs = bufB + bsz;                                // IID_ROWS*kp
  float* zts = zs + IID_ROWS * kp;                       // IID_ROWS*kp
  float* marg = zts + IID_ROWS * kp;                     // 4*kp : pi, pj, ri, rj
  double* red = reinterpret_cast<double*>(marg + 4 * kp + ((4 * kp) & 1));  // 33 doubles (8B aligned)

  const float* z = p.z + (size_t)s * p.n * k;
  const float* zt = p.zt + (size_t)s * p.n * k;
  const int rows_per = (p.n + C - 1) / C;
  const int row_lo = min(p.n, rank * rows_per), row_hi = min(p.n, row_lo + rows_per);

  if (p.phase != IIC_PHASE_FINISH) {
    // ---- 1. partial joint over this CTA's rows ------------------------------------------
    float acc[IID_MAX_TILES][16];
#pragma unroll
    for (int j = 0; j < IID_MAX_TILES; ++j)
#pragma unroll
      for (int e = 0; e < 16; ++e) acc[j][e] = 0.f;
    const int G = p.groups, T = p.ntiles, tk = p.tk;
    const int my_g = (G > 1) ? tid / T : 0;
    const int my_t = (G > 1) ? tid - my_g * T : tid;
    const bool active = (G > 1) ? (my_g < G) : true;

    for (int r0 = row_lo; r0 < row_hi; r0 += IID_ROWS) {
      const int rows = min(IID_ROWS, row_hi - r0);
      __syncthreads();
      stage_rows(z, zs, r0, rows, p.n, k, kp);
      stage_rows(zt, zts, r0, rows, p.n, k, kp);
      __syncthreads();
      if (active) {
#pragma unroll
        for (int j = 0; j < IID_MAX_TILES; ++j) {
          const int tile = my_t + j * IID_THREADS;
          if (tile < T) {
            const int ta = tile / tk, tb = tile - ta * tk;
            for (int r = my_g; r < rows; r += G) {
              const float4 a = *reinterpret_cast<const float4*>(zs + r * kp + 4 * ta);
              const float4 b = *reinterpret_cast<const float4*>(zts + r * kp + 4 * tb);
              const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
              for (int ia = 0; ia < 4; ++ia)
#pragma unroll
                for (int ib = 0; ib < 4; ++ib) acc[j][ia * 4 + ib] = fmaf(av[ia], bv[ib], acc[j][ia * 4 + ib]);
            }
          }
        }
      }
    }
    __syncthreads();
    // scatter the register tiles: G>1 -> bufB[g][kk] then fixed-order sum into bufA; else bufA
    float* dst = (G > 1) ? (bufB + my_g * kk) : bufA;
    if (active) {
#pragma unroll
      for (int j = 0; j < IID_MAX_TILES; ++j) {
        const int tile = my_t + j * IID_THREADS;
        if (tile < T) {
          const int ta = tile / tk, tb = tile - ta * tk;
#pragma unroll
          for (int ia = 0; ia < 4; ++ia)
#pragma unroll
            for (int ib = 0; ib < 4; ++ib) dst[(4 * ta + ia) * kp + 4 * tb + ib] = acc[j][ia * 4 + ib];
        }
      }
    }
    __syncthreads();
    if (G > 1) {
      for (int e = tid; e < kk; e += IID_THREADS) {
        float t = 0.f;
        for (int g = 0; g < G; ++g) t += bufB[g * kk + e];
        bufA[e] = t;
      }
      __syncthreads();
    }
    // ---- 2. cluster reduce through DSMEM (reduce-scatter, then all-gather) ----------------
    if (C > 1) {
      cluster.sync();
      const int slice = (kk + C - 1) / C;
      const int e_lo = rank * slice, e_hi = min(kk, e_lo + slice);
      for (int e = e_lo + tid; e < e_hi; e += IID_THREADS) {
        float t = 0.f;
        for (int c = 0; c < C; ++c) t += cluster.map_shared_rank(bufA, c)[e];
        bufB[e] = t;
      }
      cluster.sync();
      for (int e = tid; e < kk; e += IID_THREADS) bufA[e] = cluster.map_shared_rank(bufB, e / slice)[e];
      cluster.sync();
    }
    if (p.phase == IIC_PHASE_PARTIAL) {
      if (rank == 0)
        for (int e = tid; e < k * k; e += IID_THREADS) {
          int a = e / k, b = e - a * k;
          p.joint_ws[(size_t)s * k * k + e] = bufA[a * kp + b];
        }
      return;
    }
  } else {
    for (int e = tid; e < kk; e += IID_THREADS) {
      int a = e / kp, b = e - a * kp;
      bufA[e] = (a < k && b < k) ? p.joint_ws[(size_t)s * k * k + a * k + b] : 0.f;
    }
    __syncthreads();
  }

  // ---- 3. P, marginals, clamps, loss, G, H (every CTA redundantly; k*k is tiny) ------------
  double part = 0.0;
  for (int e = tid; e < kk; e += IID_THREADS) part += (double)bufA[e];
  const double ssum = block_sum(part, red);
  // P = (A + A^T)/2 / s   -> bufB     (reference :44-45)
  for (int e = tid; e < kk; e += IID_THREADS) {
    int a = e / kp, b = e - a * kp;
    bufB[e] = (a < k && b < k) ? (bufA[a * kp + b] + bufA[b * kp + a]) * 0.5f / (float)ssum : 0.f;
  }
  __syncthreads();
  float* pi = marg;            // row sums of the un-clamped P (:12)
  float* pj = marg + kp;       // col sums (:13-14)
  float* ri = marg + 2 * kp;   // row sums of clamped P / clamped pi  (gradient term)
  float* rj = marg + 3 * kp;
  for (int a = tid; a < k; a += IID_THREADS) {
    float si = 0.f, sj = 0.f;
    for (int b = 0; b < k; ++b) {
      si += bufB[a * kp + b];
      sj += bufB[b * kp + a];
    }
    pi[a] = si;
    pj[a] = sj;
  }
  __syncthreads();
  if (p.joint_out != nullptr && rank == 0)
    for (int e = tid; e < k * k; e += IID_THREADS) {
      int a = e / k, b = e - a * k;
      p.joint_out[(size_t)s * k * k + e] = bufB[a * kp + b];
    }
  const float eps = p.eps, lamb = p.lamb;
  for (int a = tid; a < k; a += IID_THREADS) {
    float rs = 0.f, cs = 0.f;
    for (int b = 0; b < k; ++b) {
      float v1 = bufB[a * kp + b], v2 = bufB[b * kp + a];
      rs += (v1 < eps) ? eps : v1;
      cs += (v2 < eps) ? eps : v2;
    }
    ri[a] = (pi[a] < eps) ? 0.f : rs / pi[a];  // mask m_i * sum_b P~_ab / p~_i,a
    rj[a] = (pj[a] < eps) ? 0.f : cs / pj[a];
  }
  __syncthreads();
  // loss, loss_no_lamb and G (into bufA); <G,P>
  double l_lamb = 0.0, l_one = 0.0, gp = 0.0;
  for (int e = tid; e < kk; e += IID_THREADS) {
    int a = e / kp, b = e - a * kp;
    float gval = 0.f;
    if (a < k && b < k) {
      const float praw = bufB[e];
      const bool keep = !(praw < eps);
      const float pc = keep ? praw : eps;
      const float lp = logf(pc);
      const float li = logf((pi[a] < eps) ? eps : pi[a]);
      const float lj = logf((pj[b] < eps) ? eps : pj[b]);
      l_lamb += (double)(-pc * (lp - lamb * lj - lamb * li));
      l_one += (double)(-pc * (lp - lj - li));
      gval = (keep ? (-(lp - lamb * lj - lamb * li) - 1.f) : 0.f) + lamb * (ri[a] + rj[b]);
      gp += (double)gval * (double)praw;
    }
    bufA[e] = gval;
  }
  l_lamb = block_sum(l_lamb, red);
  l_one = block_sum(l_one, red);
  gp = block_sum(gp, red);
  if (rank == 0 && tid == 0) {
    p.loss[2 * s] = (float)l_lamb;
    p.loss[2 * s + 1] = (float)l_one;
  }
  if (p.dz == nullptr && p.dzt == nullptr && p.h_out == nullptr) return;
  // H = ((G - gp) + (G - gp)^T) / (2 s)  -> bufB (symmetric)
  const float gpf = p.detached ? 0.f : (float)gp;
  const float hs = (float)(0.5 / ssum);
  for (int e = tid; e < kk; e += IID_THREADS) {
    int a = e / kp, b = e - a * kp;
    bufB[e] = (a < k && b < k) ? (bufA[a * kp + b] + bufA[b * kp + a] - 2.f * gpf) * hs : 0.f;
  }
  __syncthreads();
  if (p.h_out != nullptr && rank == 0)
    for (int e = tid; e < k * k; e += IID_THREADS) {
      int a = e / k, b = e - a * k;
      p.h_out[(size_t)s * k * k + e] = bufB[a * kp + b];
    }
  if (p.dz == nullptr && p.dzt == nullptr) return;
  // ---- 4. gradient sweep over this CTA's rows --------------------------------------------
  float* dz = p.dz ? p.dz + (size_t)s * p.n * k : nullptr;
  float* dzt = p.dzt ? p.dzt + (size_t)s * p.n * k : nullptr;
  for (int r0 = row_lo; r0 < row_hi; r0 += IID_ROWS) {
    const int rows = min(IID_ROWS, row_hi - r0);
    __syncthreads();
    stage_rows(z, zs, r0, rows, p.n, k, kp);
    stage_rows(zt, zts, r0, rows, p.n, k, kp);
    __syncthreads();
    for (int i = tid; i < rows * k; i += IID_THREADS) {
      const int r = i / k, a = i - r * k;
      float g1 = 0.f, g2 = 0.f;
      const float* zr = zs + r * kp;
      const float* ztr = zts + r * kp;
      for (int b = 0; b < k; ++b) {
        const float h = bufB[b * kp + a];  // == H[a][b]
        g1 = fmaf(h, ztr[b], g1);
        g2 = fmaf(h, zr[b], g2);
      }
      if (dz) dz[(size_t)(r0 + r) * k + a] = g1;
      if (dzt) dzt[(size_t)(r0 + r) * k + a] = g2;
    }
  }
}

}  // namespace iic

using namespace iic;

static int launch_iid(IidParams p, int S, int n, int k, cudaStream_t st) {
  p.n = n; p.k = k; p.kp = (k + 3) & ~3; p.tk = p.kp / 4; p.ntiles = p.tk * p.tk;
  IIC_REQUIRE(p.ntiles <= IID_MAX_TILES * IID_THREADS, IIC_ERR_UNSUPPORTED,
              "iic joint/MI kernel: k=%d exceeds the fused kernel's limit (k <= 156)", k);
  p.groups = p.ntiles >= IID_THREADS ? 1 : IID_THREADS / p.ntiles;
  if (p.groups > IID_ROWS) p.groups = IID_ROWS;
  const int kk = p.kp * p.kp;
  const int bsz = kk > p.groups * kk ? kk : p.groups * kk;
  size_t smem = (size_t)(kk + bsz + 2 * IID_ROWS * p.kp + 4 * p.kp + 2) * sizeof(float) + 40 * sizeof(double);
  IIC_REQUIRE(smem <= 220 * 1024, IIC_ERR_UNSUPPORTED, "iic joint/MI kernel: k=%d needs %zu B of shared memory", k, smem);
  IIC_CUDA(cudaFuncSetAttribute(iid_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int C = 8;
  while (C > 1 && (n + C - 1) / C < 8) C >>= 1;  // tiny batches: fewer CTAs per cluster
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(C, S, 1);
  cfg.blockDim = dim3(IID_THREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  IIC_CUDA(cudaLaunchKernelEx(&cfg, iid_loss_kernel, p));
  count_launch();
  return IIC_OK;
}

extern "C" int iic_iid_loss(const float* z, const float* zt, int S, int n, int k, float lamb, double eps, float* loss,
                            float* dz, float* dzt, float* joint_ws, float* joint_out, int phase, void* stream) {
  using namespace iic;
  IIC_REQUIRE(z && zt && S > 0 && n > 0 && k > 0, IIC_ERR_BAD_ARG, "iic_iid_loss: bad arguments");
  IIC_REQUIRE(phase == IIC_PHASE_FUSED || phase == IIC_PHASE_PARTIAL || phase == IIC_PHASE_FINISH, IIC_ERR_BAD_ARG,
              "iic_iid_loss: bad phase %d", phase);
  IIC_REQUIRE(phase == IIC_PHASE_FUSED || joint_ws, IIC_ERR_BAD_ARG, "iic_iid_loss: joint_ws required for phase %d",
              phase);
  IIC_REQUIRE(phase == IIC_PHASE_PARTIAL || loss, IIC_ERR_BAD_ARG, "iic_iid_loss: loss output required");
  IidParams p = {};
  p.z = z; p.zt = zt; p.loss = loss; p.dz = dz; p.dzt = dzt; p.joint_ws = joint_ws; p.joint_out = joint_out;
  p.h_out = nullptr; p.detached = 0;
  p.lamb = lamb; p.eps = (float)eps; p.phase = phase;
  return launch_iid(p, S, n, k, (cudaStream_t)stream);
}

extern "C" int iic_joint_mi(const float* joint, int S, int k, float lamb, double eps, int detached_norm, float* loss,
                            float* h_out, void* stream) {
  using namespace iic;
  IIC_REQUIRE(joint && loss && S > 0 && k > 0, IIC_ERR_BAD_ARG, "iic_joint_mi: bad arguments");
  IidParams p = {};
  p.z = nullptr; p.zt = nullptr; p.loss = loss; p.dz = nullptr; p.dzt = nullptr;
  p.joint_ws = const_cast<float*>(joint); p.joint_out = nullptr; p.h_out = h_out; p.detached = detached_norm ? 1 : 0;
  p.lamb = lamb; p.eps = (float)eps; p.phase = IIC_PHASE_FINISH;
  return launch_iid(p, S, 0, k, (cudaStream_t)stream);
}
