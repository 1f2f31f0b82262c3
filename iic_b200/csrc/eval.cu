// Evaluation path of xu-ji/IIC: label assignment and the confusion counts behind
//   code/utils/cluster/cluster_eval.py:46-63       torch.argmax(x_outs_curr, dim=1) per sub-head
//   code/utils/segmentation/segmentation_eval.py:100-110   per-pixel argmax over the channel dimension + mask
//   code/utils/cluster/eval_metrics.py:9-53        the k x k loops `int(((flat_preds == c1) * (flat_targets == c2)).sum())`
//                                                  (one device reduction AND one host synchronisation per pair (c1, c2):
//                                                  700 of each for the over-clustering head) -> one histogram launch.
// Integer / index work: the results are exact, the histogram uses integer atomics (order independent).
#include "common.cuh"

namespace iic {

// first index of the maximum of a row (torch.argmax: ties -> lowest index; a NaN is the maximum, like torch)
__device__ __forceinline__ bool arg_better(float v, int i, float bv, int bi) {
  const bool vn = v != v, bn = bv != bv;
  if (vn != bn) return vn;              // NaN beats any number
  if (vn) return i < bi;                // both NaN: lowest index
  return v > bv || (v == bv && i < bi);
}

// z [rows][k] (rows = S * n for stacked sub-heads) -> out[rows]; one warp per row
__global__ void argmax_rows_kernel(const float* __restrict__ z, long long rows, int k, int* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int c = lane; c < k; c += 32) {
      const float v = z[r * k + c];
      if (bi == 0x7fffffff || arg_better(v, c, bv, bi)) {
        bv = v;
        bi = c;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || arg_better(ov, oi, bv, bi))) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) out[r] = bi;
  }
}

// x [n][k][hw] (NCHW) -> out[n * hw]; one thread per pixel, channel loop strided by hw (coalesced across the warp)
__global__ void argmax_channels_kernel(const float* __restrict__ x, int n, int k, long long hw, int* __restrict__ out) {
  const long long total = (long long)n * hw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long ni = i / hw, p = i - ni * hw;
    const float* px = x + ni * k * hw + p;
    float bv = px[0];
    int bi = 0;
    for (int c = 1; c < k; ++c) {
      const float v = px[(long long)c * hw];
      if (arg_better(v, c, bv, bi)) {
        bv = v;
        bi = c;
      }
    }
    out[i] = bi;
  }
}

// counts[s][p][t] += #{i : preds[s][i] == p, targets[i] == t, mask[i] != 0}; labels outside [0, pk) x [0, tk) are
// ignored, exactly like the reference's loops over range(preds_k) x range(targets_k).
// SMEM: per-block histogram in shared memory (pk * tk <= 8192), one global atomic per non-empty bin and block.
template <bool SMEM>
__global__ void confusion_kernel(const int* __restrict__ preds, const int* __restrict__ targets,
                                 const unsigned char* __restrict__ mask, long long n, int pk, int tk,
                                 unsigned long long* __restrict__ counts) {
  extern __shared__ unsigned int hist[];
  const int s = blockIdx.y, bins = pk * tk;
  const int* ps = preds + (long long)s * n;
  unsigned long long* cs = counts + (long long)s * bins;
  if (SMEM) {
    for (int i = threadIdx.x; i < bins; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (mask != nullptr && mask[i] == 0) continue;
    const int p = ps[i], t = targets[i];
    if ((unsigned)p >= (unsigned)pk || (unsigned)t >= (unsigned)tk) continue;
    if (SMEM)
      atomicAdd(&hist[p * tk + t], 1u);
    else
      atomicAdd(&cs[p * tk + t], 1ull);
  }
  if (SMEM) {
    __syncthreads();
    for (int i = threadIdx.x; i < bins; i += blockDim.x)
      if (hist[i] != 0u) atomicAdd(&cs[i], (unsigned long long)hist[i]);
  }
}

}  // namespace iic

using namespace iic;

static int eval_grid(long long work_items, int per_block) {
  long long b = (work_items + per_block - 1) / per_block;
  const long long cap = (long long)device_sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

extern "C" int iic_argmax_rows(const float* z, long long rows, int k, int* out, void* stream) {
  IIC_REQUIRE(z && out && rows > 0 && k > 0, IIC_ERR_BAD_ARG, "iic_argmax_rows: bad arguments");
  argmax_rows_kernel<<<eval_grid(rows, 8), 256, 0, (cudaStream_t)stream>>>(z, rows, k, out);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_argmax_channels(const float* x_nchw, int n, int k, long long hw, int* out, void* stream) {
  IIC_REQUIRE(x_nchw && out && n > 0 && k > 0 && hw > 0, IIC_ERR_BAD_ARG, "iic_argmax_channels: bad arguments");
  argmax_channels_kernel<<<eval_grid((long long)n * hw, 256), 256, 0, (cudaStream_t)stream>>>(x_nchw, n, k, hw, out);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

extern "C" int iic_confusion_counts(const int* preds, const int* targets, const unsigned char* mask, int S, long long n,
                                    int preds_k, int targets_k, long long* counts, int accumulate, void* stream) {
  IIC_REQUIRE(preds && targets && counts && S > 0 && n > 0 && preds_k > 0 && targets_k > 0 && S <= 65535, IIC_ERR_BAD_ARG,
              "iic_confusion_counts: bad arguments");
  IIC_REQUIRE((long long)preds_k * targets_k <= (1 << 24), IIC_ERR_UNSUPPORTED, "iic_confusion_counts: %d x %d bins", preds_k, targets_k);
  cudaStream_t st = (cudaStream_t)stream;
  const int bins = preds_k * targets_k;
  if (!accumulate) IIC_CUDA(cudaMemsetAsync(counts, 0, sizeof(long long) * (size_t)S * bins, st));
  const dim3 grid(eval_grid(n, 256 * 8), S);
  if (bins <= 8192)
    confusion_kernel<true><<<grid, 256, sizeof(unsigned int) * bins, st>>>(preds, targets, mask, n, preds_k, targets_k,
                                                                          (unsigned long long*)counts);
  else
    confusion_kernel<false><<<grid, 256, 0, st>>>(preds, targets, mask, n, preds_k, targets_k, (unsigned long long*)counts);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}
