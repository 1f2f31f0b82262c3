// tcgen05 implicit-GEMM convolution, TMA-fed persistent version (sm_100a).
//
// The operands are staged by the Tensor Memory Accelerator (round 1's first version gathered them with 16-byte
// cp.async copies, which saturated the LSU at ~1/3-2/3 of the tensor rate: profiles/r01_conv_sweep.md):
//   * activation operand: ONE `cp.async.bulk.tensor.4d...im2col` per 128x64 tile -- the TMA unit
//     walks the output pixels of the tile (n, oh, ow order), applies the conv stride, the filter-tap
//     offset and zero-fills the padding, writing 128-byte rows with the SWIZZLE_128B pattern the
//     UMMA descriptor expects (im2col is never materialised);
//   * dense operand (packed weights, or dy for wgrad): tiled 2-D TMA boxes (1 per stage; BN/64 for wgrad).
// Warp roles (192 threads, one persistent CTA per SM, static tile schedule):
//   warp 5 lane 0: TMA producer (mbarrier expect_tx)      warp 4 lane 0: tcgen05.mma issuer
//   warps 0-3    : epilogue; TMEM accumulators are double-buffered (2 x BN columns) so the
//                  epilogue of tile i overlaps the main loop of tile i+1.
// Used for fprop (stride 1/2), dgrad of stride-1 convs (im2col of dy with reversed taps) and wgrad
// (both operands MN-major).  Stride-2 dgrad is decomposed into four stride-1 parity-class problems (tc2_conv_dgrad_s2).
#include <cuda.h>

#include <cstdlib>

#include "tc_ptx.cuh"

namespace iic {

enum { M2_FPROP = 0, M2_WGRAD = 1 };
constexpr int TC2_THREADS = 192;

constexpr int TC2_MAXTAPS = 25;

struct Tc2Params {
  long long rows;   // pixels enumerated by the gather (n * rowH * rowW)
  int rowH, rowW;
  int KH, KW, s, d;
  int lower;        // input-space coordinate offset of filter tap 0 (wgrad; fprop uses lower_h / lower_w)
  int lower_h, lower_w;
  // fprop / dgrad: the K dimension enumerates (tap index t, channel); per tap: activation offset
  // (im2col {off_w, off_h}) and the tap's position in the packed weight (K coordinate = wtap*srcC + c)
  int ntaps;
  unsigned char offh[TC2_MAXTAPS], offw[TC2_MAXTAPS], wtap[TC2_MAXTAPS];
  // stride-2 dgrad parity class: tile rows enumerate (n, i, j) and are written to dx pixel (2i+py, 2j+px)
  int scatter, outH, outW, py, px;
  int srcC, Ktot, N;
  int mtiles, ntiles, splits, kb_per_split, total_kb;
  __nv_bfloat16* out;
  const __nv_bfloat16* addend;
  const unsigned char* addend_mask;  // or null: one bit per addend element ([pixel][N / 8] bytes); a clear bit drops the addend
  int addend_prefetch;  // fetch the addend one 32-column chunk ahead of its use (option dgrad_prefetch)
  float* partial;
  // fused BatchNorm statistics (fprop only): per-CTA partial column sums of the fp32 accumulators,
  // stat_partial[cta][group][{sum, sum of squares}][N]; rows < stat_half belong to view 0, the rest to view 1
  float* stat_partial;
  long long stat_half;
};

// Transpose-reduce: on return t[0] of lane l holds sum over the 32 lanes of (their) t[l]  (31 shuffles).
__device__ __forceinline__ void warp_col_reduce(float (&t)[32], int lane) {
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

// RESB: the whole dense operand (all k-blocks of the packed weight for this N tile) stays resident in shared
// memory for the lifetime of the persistent CTA; the ring then carries only the activation tile.  Used for
// the 64-channel layers (3x3x64x64 weights = 72 KB), where re-fetching the weight tile for every 128-pixel
// tile was a third of the L2->SM traffic that bounds them.
// With RESB a work item is TWO 128-row tiles (MT = 2): one 256-pixel TMA box and one barrier round-trip feed eight
// MMAs, halving the per-byte producer / barrier overhead that bounds N = 64 tiles (128 MMA cycles per k-block).
// MTP = 2 without RESB (N = 128 tiles, option tc2_mt2): one weight k-block (16 KB) feeds two 128-row activation tiles, so
// the shared-memory fill per tensor cycle drops from 128 to 96 B and a barrier round trip covers eight MMAs.
template <int BN, bool RESB = false, int MTP = (RESB ? 2 : 1)> struct Tc2Cfg {
  static constexpr int MT = MTP;
  static constexpr int A_BYTES = MT * TC_BM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + (RESB ? 0 : B_BYTES);
  static constexpr int STAGES = (STAGE_BYTES >= 64 * 1024) ? 3 : ((RESB || MT == 2) ? 4 : ((BN >= 256) ? 4 : (BN >= 128 ? 6 : 8)));
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 + 256;  // (+ 64 B x N for fused BN statistics)
  // two accumulator buffers (the epilogue of a work item overlaps the MMAs of the next) when they fit the 512 TMEM columns;
  // the multi-tile wgrad work items (MT x BN > 256) run for hundreds of k-blocks and get by with one
  static constexpr int ACC_BUFS = (2 * MT * BN <= 512) ? 2 : 1;
  static constexpr int ACC_COLS = ACC_BUFS * MT * BN;
  static constexpr int TMEM_COLS = ACC_COLS <= 64 ? 64 : (ACC_COLS <= 128 ? 128 : (ACC_COLS <= 256 ? 256 : 512));
};

template <int MODE, int BN, bool RESB = false, int MTP = (RESB ? 2 : 1)>
__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Tc2Params P) {
  using Cfg = Tc2Cfg<BN, RESB, MTP>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int MT = Cfg::MT;
  static_assert(!RESB || MODE == M2_FPROP, "resident dense operand: fprop/dgrad only");
  static_assert(MT <= 2 || MODE == M2_WGRAD, "three-tile work items: wgrad only");
  static_assert(Cfg::ACC_COLS <= 512 && (!RESB || MT == 2), "TMEM budget / RESB implies two-tile work items");
  constexpr bool TWO_ACC = Cfg::ACC_BUFS == 2;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t resb = (raw + 1023u) & ~1023u;  // [total_kb][BN x 128 B] resident weights (RESB), else empty
  const uint32_t base = resb + (RESB ? (uint32_t)P.total_kb * Cfg::B_BYTES : 0u);
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bres_bar = bars + 8u * (2 * STAGES + 4);
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + STAGES * Cfg::STAGE_BYTES + (2 * STAGES + 5) * 8);
  float* stat = reinterpret_cast<float*>(base_ptr + STAGES * Cfg::STAGE_BYTES + 256);  // [4 warps][2 views][2][N]
  const bool do_stats = (MODE == M2_FPROP) && P.stat_partial != nullptr;
  if (do_stats)
    for (int i = threadIdx.x; i < 16 * P.N; i += TC2_THREADS) stat[i] = 0.f;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 128);
    }
    mbar_init(bres_bar, 1);
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

  // work item -> (split z, m tile, n tile, k-block range)
  auto decode = [&](int w, int& z, long long& m0, int& n0, int& kb0, int& nk) {
    z = w / tiles_mn;
    const int r = w - z * tiles_mn;
    n0 = (r % P.ntiles) * BN;
    m0 = (long long)(r / P.ntiles) * (MT * TC_BM);
    if (MODE == M2_FPROP) {
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
    // The whole warp walks the k-blocks (uniform scalar state, no divisions in the loop); lane 0 arms the
    // barrier, then lane i issues the i-th TMA of the stage, so the 2..6 bulk copies of a stage are
    // issued in parallel instead of back to back by one thread (the single-thread issue rate, not L2
    // bandwidth, bounded the 64-channel layers and wgrad: profiles/r01_ncu_wgrad.md).
    uint32_t it = 0;  // global k-block counter -> stage / phase
    if (RESB && blockIdx.x < total_work) {
      // one-off: the whole packed weight of this (single) N tile, k-block by k-block
      if (lane == 0) mbar_expect_tx(bres_bar, (uint32_t)P.total_kb * Cfg::B_BYTES);
      __syncwarp();
      int tap = 0, c0 = 0;
      for (int i = 0; i < P.total_kb; ++i) {
        if (lane == (i & 31)) tma_load_2d(resb + i * Cfg::B_BYTES, &tmB, bres_bar, (int)P.wtap[tap] * P.srcC + c0, 0);
        c0 += 64;
        if (c0 >= P.srcC) {
          c0 = 0;
          ++tap;
        }
      }
    }
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      int z, n0, kb0, nk;
      long long m0;
      decode(w, z, m0, n0, kb0, nk);
      if (MODE == M2_FPROP) {
        const int ox = (int)(m0 % P.rowW);
        const long long q = m0 / P.rowW;
        const int oy = (int)(q % P.rowH);
        const int img = (int)(q / P.rowH);
        const int cw = ox * P.s + P.lower_w, ch = oy * P.s + P.lower_h;
        int tap = 0, c0 = 0;
        for (int i = 0; i < nk; ++i, ++it) {
          const int s = it % STAGES;
          mbar_wait(empty_bar(s), ((it / STAGES) & 1u) ^ 1u);
          const uint32_t sa = base + s * Cfg::STAGE_BYTES, sb = sa + Cfg::A_BYTES;
          if (lane == 0) mbar_expect_tx(full_bar(s), Cfg::STAGE_BYTES);
          __syncwarp();
          if (lane == 0)
            tma_load_im2col(sa, &tmA, full_bar(s), c0, cw, ch, img, (uint16_t)P.offw[tap], (uint16_t)P.offh[tap]);
          else if (lane == 1 && !RESB)
            tma_load_2d(sb, &tmB, full_bar(s), (int)P.wtap[tap] * P.srcC + c0, n0);
          c0 += 64;
          if (c0 >= P.srcC) {
            c0 = 0;
            ++tap;
          }
        }
      } else {
        // A = MT tiles of 128 (tap, cin) rows = 2 MT column blocks of 64 channels; lane b issues block b
        int a_off_w = 0, a_off_h = 0, a_c0 = 0;
        bool a_ok = false;
        if (lane < 2 * MT) {
          const long long j = m0 + lane * 64;
          a_ok = j < P.Ktot;
          const int tap = (int)(j / P.srcC);
          a_c0 = (int)(j - (long long)tap * P.srcC);
          const int ta = tap / P.KW;
          a_off_h = ta * P.d;
          a_off_w = (tap - ta * P.KW) * P.d;
        }
        int nok = 0;
#pragma unroll
        for (int b = 0; b < 2 * MT; ++b) nok += (m0 + b * 64 < P.Ktot) ? 1 : 0;
        const uint32_t bytes = (uint32_t)(nok * 8192 + Cfg::B_BYTES);
        // pixel coordinates of the first row of the k-block, advanced by 64 pixels per k-block
        long long p0 = (long long)kb0 * 64;
        int ox = (int)(p0 % P.rowW);
        const long long q = p0 / P.rowW;
        int oy = (int)(q % P.rowH);
        int img = (int)(q / P.rowH);
        const int step_x = 64 % P.rowW, step_y = 64 / P.rowW;
        for (int i = 0; i < nk; ++i, ++it) {
          const int s = it % STAGES;
          mbar_wait(empty_bar(s), ((it / STAGES) & 1u) ^ 1u);
          const uint32_t sa = base + s * Cfg::STAGE_BYTES, sb = sa + Cfg::A_BYTES;
          if (lane == 0) mbar_expect_tx(full_bar(s), bytes);
          __syncwarp();
          if (lane < 2 * MT) {
            if (a_ok)
              tma_load_im2col(sa + lane * 8192, &tmA, full_bar(s), a_c0, ox * P.s + P.lower, oy * P.s + P.lower, img,
                              (uint16_t)a_off_w, (uint16_t)a_off_h);
          } else if (lane < 2 * MT + BN / 64) {
            const int b = lane - 2 * MT;
            tma_load_2d(sb + b * 8192, &tmB, full_bar(s), n0 + b * 64, (int)p0);
          }
          p0 += 64;
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
    // The issue path is kept warp-uniform (wrapping stage counter, descriptors = one base + immediate adds, the
    // issuing lane chosen by elect.sync): ptxas then holds the descriptors in uniform registers and emits
    // UTCHMMA back to back.  With `if (lane == 0)` and descriptors rebuilt from vector registers every MMA cost an
    // ELECT / R2UR round trip (~15 instructions), which bounded the N = 64 tiles (32 tensor cycles per MMA).
    constexpr uint32_t idesc = (MODE == M2_FPROP) ? make_idesc(BN, 0, 0) : make_idesc(BN, 1, 1);
    uint32_t s = 0, sphase = 0, tile_it = 0;
    if (RESB && blockIdx.x < total_work) mbar_wait(bres_bar, 0);
    const uint64_t desc_base = (MODE == M2_FPROP) ? make_desc(base, 16, 1024) : make_desc(base, 8192, 1024);
    const uint64_t desc_resb = make_desc(resb, 16, 1024);
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++tile_it) {
      int z, n0, kb0, nk;
      long long m0;
      decode(w, z, m0, n0, kb0, nk);
      const uint32_t as = TWO_ACC ? (tile_it & 1u) : 0u;
      mbar_wait(tempty_bar(as), ((TWO_ACC ? (tile_it >> 1) : tile_it) & 1u) ^ 1u);  // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + as * (MT * BN);
      for (int i = 0; i < nk; ++i) {
        mbar_wait(full_bar(s), sphase);
        tc_fence_after();
        if (elect_one_sync()) {
          const uint64_t ad0 = desc_base + (uint64_t)((s * (uint32_t)Cfg::STAGE_BYTES) >> 4);
          const uint64_t bd0 = RESB ? desc_resb + (uint64_t)(((uint32_t)i * (uint32_t)Cfg::B_BYTES) >> 4)
                                    : ad0 + (uint64_t)(Cfg::A_BYTES >> 4);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              // K-major: 32 B per K = 16 step inside the 128 B swizzle row; MN-major: 16 K-rows x 128 B
              const uint32_t koff = (MODE == M2_FPROP) ? kk * 32 : kk * 2048;
              const uint64_t ad = ad0 + (uint64_t)((mt * (TC_BM * 128) + koff) >> 4);
              const uint64_t bd = bd0 + (uint64_t)(koff >> 4);
              umma_bf16(tmem_acc + mt * BN, ad, bd, idesc, (i > 0 || kk > 0) ? 1u : 0u);
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
      if (elect_one_sync()) umma_commit(tfull_bar(as));  // (also correct for nk == 0: arrives immediately)
      __syncwarp();
    }
  } else {
    // =============================== epilogue (warps 0-3) ======================================
    uint32_t tile_it = 0;
    constexpr int CH = BN / 32;  // 32-column chunks per 128-row accumulator tile
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++tile_it) {
      int z, n0, kb0, nk;
      long long m0;
      decode(w, z, m0, n0, kb0, nk);
      const uint32_t as = TWO_ACC ? (tile_it & 1u) : 0u;
      const uint32_t aphase = (TWO_ACC ? (tile_it >> 1) : tile_it) & 1u;
      if constexpr (MODE == M2_FPROP) {
        // this thread's row in each of the MT accumulator tiles (TMEM lane == tile row) and where it is written
        const long long mrow0 = m0 + warp * 32 + lane, mrow1 = mrow0 + TC_BM;
        long long orow0 = mrow0, orow1 = mrow1;
        if (P.scatter) {
          auto scat = [&](long long m) {
            const int jj = (int)(m % P.rowW);
            const long long q = m / P.rowW;
            const int ii = (int)(q % P.rowH);
            return ((q / P.rowH) * P.outH + 2 * ii + P.py) * P.outW + 2 * jj + P.px;
          };
          if (mrow0 < P.rows) orow0 = scat(mrow0);
          if (MT == 2 && mrow1 < P.rows) orow1 = scat(mrow1);
        }
        // dgrad adds the residual-branch gradient (`addend`) in the epilogue.  Its 64 B per thread and chunk are fetched
        // one chunk ahead (the first one before the accumulator is even complete): fetched on demand, the global-load
        // latency sat between every tcgen05.ld and its stores and cost the dgrad launches with an addend ~50 % of
        // their time (profiles/r01_ncu_halo.md).
        const bool has_add = P.addend != nullptr;
        const bool pre = has_add && P.addend_prefetch != 0;
        uint4 acur[4], anxt[4];
        uint32_t mcur = 0xffffffffu, mnxt = 0xffffffffu;  // addend mask bits of the 32 columns of the chunk
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) acur[qq] = anxt[qq] = make_uint4(0u, 0u, 0u, 0u);
        auto fetch = [&](uint4(&dst)[4], uint32_t& mdst, int mt, int c0) {
          const long long mr = (MT == 2 && mt == 1) ? mrow1 : mrow0;
          const long long orow = (MT == 2 && mt == 1) ? orow1 : orow0;
          if (mr < P.rows) {
            const uint4* src = reinterpret_cast<const uint4*>(P.addend + orow * P.N + n0 + c0);
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) dst[qq] = src[qq];
            if (P.addend_mask != nullptr)
              mdst = *reinterpret_cast<const uint32_t*>(P.addend_mask + ((orow * P.N + n0 + c0) >> 3));
          }
        };
        if (pre) fetch(acur, mcur, 0, 0);
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < MT * CH; ++j) {
          const int mt = j / CH, c0 = (j % CH) * 32;
          const long long mtile0 = m0 + (long long)mt * TC_BM;
          const long long m = (MT == 2 && mt == 1) ? mrow1 : mrow0;
          const long long orow = (MT == 2 && mt == 1) ? orow1 : orow0;
          const uint32_t tmem_acc = tmem_base + as * (MT * BN) + mt * BN + ((uint32_t)(warp * 32) << 16);
          uint32_t v[32];
          if (nk > 0) {
            tmem_ld32(tmem_acc + (uint32_t)c0, v);
          } else {
#pragma unroll
            for (int qq = 0; qq < 32; ++qq) v[qq] = 0u;
          }
          if (pre && j + 1 < MT * CH) fetch(anxt, mnxt, (j + 1) / CH, ((j + 1) % CH) * 32);
          if (nk > 0) tmem_ld_wait();
          if (do_stats) {
            // per-column sums over this warp's 32 rows (rows >= P.rows are exact zeros: TMA zero fill)
            const bool lo1 = mtile0 >= P.stat_half, hi1 = (mtile0 + TC_BM - 1) >= P.stat_half;
            for (int grp = lo1 ? 1 : 0; grp <= (hi1 ? 1 : 0); ++grp) {
              const bool mine = (m >= P.stat_half) == (grp == 1);
              float t[32];
#pragma unroll
              for (int e = 0; e < 32; ++e) t[e] = mine ? __uint_as_float(v[e]) : 0.f;
              warp_col_reduce(t, lane);
              const float s1 = t[0];
#pragma unroll
              for (int e = 0; e < 32; ++e) {
                const float f = mine ? __uint_as_float(v[e]) : 0.f;
                t[e] = f * f;
              }
              warp_col_reduce(t, lane);
              float* sp = stat + ((warp * 2 + grp) * 2) * P.N + n0 + c0 + lane;
              sp[0] += s1;
              sp[P.N] += t[0];
            }
          }
          if (m < P.rows) {
            if (has_add && !pre) fetch(acur, mcur, mt, c0);
            __nv_bfloat16* o = P.out + orow * P.N + n0 + c0;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
              float f[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[qq * 8 + e]);
              if (has_add) {
                float ad[8];
                Raw8h rw;
                rw.v = acur[qq];
                cvt_raw(rw, ad);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] += ((mcur >> (qq * 8 + e)) & 1u) ? ad[e] : 0.f;
              }
              store8(o + qq * 8, f);
            }
          }
          if (pre) {
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) acur[qq] = anxt[qq];
            mcur = mnxt;
          }
        }
      } else {
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt) {
          const long long m = m0 + mt * TC_BM + warp * 32 + lane;  // TMEM lane == tile row
          const uint32_t tmem_acc = tmem_base + as * (MT * BN) + mt * BN + ((uint32_t)(warp * 32) << 16);
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
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(as));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  if (do_stats)
    for (int i = threadIdx.x; i < 4 * P.N; i += TC2_THREADS)  // i = (view, q, col); fixed warp order
      P.stat_partial[(long long)blockIdx.x * 4 * P.N + i] = stat[i] + stat[4 * P.N + i] + stat[8 * P.N + i] + stat[12 * P.N + i];
}


// =================================================================================================
// Halo variant for 3x3 / stride 1 / pad 1 / 64 -> 64 channels (ClusterNet5g layer1 fprop and dgrad).
//
// The im2col kernel above re-reads every input pixel nine times from L2 (one 128-pixel box per filter
// tap): at N = 64 that is 128 B of shared-memory fill per MMA cycle and the L2 -> SM path, not the tensor
// pipe, bounds the layer (profiles/r01_conv_sweep.md).  Here ONE tiled TMA box per work item brings
// R+2 input rows of an image, padded to width Wp = W+2 by the TMA zero fill, into shared memory as
// consecutive 128-byte rows (row = (iy - y0 + 1) * Wp + ix + 1).  On that padded grid the A operand of
// filter tap (a, b) is the SAME buffer started a*Wp + b rows later, so all nine taps (72 MMAs for a
// 256-row tile) are fed from one 45 KB load.  A UMMA descriptor may start at any 128-byte row of a
// 1024-byte aligned SWIZZLE_128B buffer with base_offset = 0 (the swizzle is a function of the absolute
// shared-memory address; measured with tools/umma_shift_probe.cu).  Tile rows m = r*Wp + x with x >= W
// or y0 + r >= H are padding positions: computed, never stored, excluded from the statistics.
// Warps: 0-7 epilogue (TMEM lane quadrant = warp % 4, column half = warp / 4), 8 MMA issuer, 9 TMA producer.
constexpr int HALO_THREADS = 320;
constexpr int HALO_TILE = 256;  // accumulator rows per work item (two M = 128 MMA tiles)

struct HaloParams {
  int nimg, H, W, Wp, R;
  int tiles_per_img, total_tiles;
  int stage_bytes, box_bytes, stages;
  unsigned short shift[9];
  unsigned char wtap[9];
  __nv_bfloat16* out;
  const __nv_bfloat16* addend;
  const unsigned char* addend_mask;  // or null: [pixel][8] bytes, one bit per addend element (a clear bit drops it)
  float* stat_partial;  // [cta][view][{sum, sum of squares}][64] or null
  int img_half;         // images >= img_half belong to view 1
  int addend_prefetch;  // request the addend before waiting for the accumulator (option dgrad_prefetch)
  int tma_store;        // output tile staged in shared memory and written by one TMA store (option conv_halo_store)
  int addend_tma;       // dgrad: the addend tile is fetched INTO the staging buffer by one TMA load and summed in place
                        // (option halo_addend_tma; needs tma_store)
};

// Output path (fprop / dgrad).  With one accumulator row per lane, a warp's 16-byte global stores hit 32 different
// 128-byte lines per instruction: 2048 LSU wavefronts per 256 x 64 tile against 2304 tensor cycles of MMAs -- the
// epilogue, not the tensor pipe, paced this kernel (profiles/r01_ncu_halo.md: 41-47 % tensor active).  tma_store != 0:
// the epilogue warps write the bf16 tile into a SWIZZLE_128B staging buffer (conflict-free 16-byte STS) and ONE
// cp.async.bulk.tensor store per work item writes it out; the box is the load box without the halo rows (64 x Wp x R),
// so the padding columns x >= W and the rows below the image are clipped by the TMA unit like they were zero-filled
// on the way in.
constexpr uint32_t HALO_STAGING_BYTES = HALO_TILE * 128;

__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmAdd, HaloParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t resb = (raw + 1023u) & ~1023u;           // [9][64 x 128 B] resident weights
  const uint32_t staging = resb + 9u * 8192u;             // [256][128 B] bf16 output tile (tma_store only)
  const uint32_t base = staging + (P.tma_store ? HALO_STAGING_BYTES : 0u);  // stages
  const uint32_t bars = base + (uint32_t)P.stages * (uint32_t)P.stage_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };          // s < 4
  auto empty_bar = [&](int s) { return bars + 8u * (4 + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (8 + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (10 + a); };
  const uint32_t bres_bar = bars + 8u * 12;
  const uint32_t add_bar = bars + 8u * 14;  // addend tile landed in the staging buffer (addend_tma)
  uint8_t* bars_ptr = smem_raw + (bars - raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_ptr + 8 * 13);
  float* stat = reinterpret_cast<float*>(bars_ptr + 128);  // [8 warps][2 views][2][32]
  const bool do_stats = P.stat_partial != nullptr;
  if (do_stats)
    for (int i = threadIdx.x; i < 8 * 2 * 2 * 32; i += HALO_THREADS) stat[i] = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 256);
    }
    mbar_init(bres_bar, 1);
    mbar_init(add_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (P.tma_store) tma_prefetch_desc(&tmO);
    if (P.addend_tma) tma_prefetch_desc(&tmAdd);
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    // =============================== TMA producer ============================================
    if (blockIdx.x < P.total_tiles) {
      if (lane == 0) mbar_expect_tx(bres_bar, 9u * 8192u);
      __syncwarp();
      if (lane < 9) tma_load_2d(resb + lane * 8192u, &tmB, bres_bar, (int)P.wtap[lane] * 64, 0);
    }
    if (lane == 0) {
      uint32_t it = 0;
      for (int w = blockIdx.x; w < P.total_tiles; w += gridDim.x, ++it) {
        const int img = w / P.tiles_per_img;
        const int y0 = (w - img * P.tiles_per_img) * P.R;
        const int s = it % P.stages;
        mbar_wait(empty_bar(s), ((it / P.stages) & 1u) ^ 1u);
        mbar_expect_tx(full_bar(s), (uint32_t)P.box_bytes);
        tma_load_4d(base + (uint32_t)s * (uint32_t)P.stage_bytes, &tmA, full_bar(s), 0, -1, y0 - 1, img);
      }
    }
  } else if (warp == 8) {
    // =============================== MMA issuer ==============================================
    // Everything that feeds a descriptor stays warp-uniform (wrapping stage counter instead of a modulo, the
    // nine tap shifts read from the kernel parameters by a fully unrolled loop): ptxas then keeps the
    // descriptors in uniform registers and issues UTCHMMA back to back; a descriptor that passes through a
    // vector register costs an ELECT / R2UR round trip per MMA, which at N = 64 (32 MMA cycles) is the bound.
    constexpr uint32_t idesc = make_idesc(64, 0, 0);
    uint32_t it = 0, s = 0, sphase = 0;
    if (blockIdx.x < P.total_tiles) mbar_wait(bres_bar, 0);
    const uint64_t bdesc_res = make_desc(resb, 16, 1024);
    for (int w = blockIdx.x; w < P.total_tiles; w += gridDim.x, ++it) {
      const uint32_t as = it & 1u;
      mbar_wait(tempty_bar(as), ((it >> 1) & 1u) ^ 1u);
      mbar_wait(full_bar(s), sphase);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t adesc0 = make_desc(base + s * (uint32_t)P.stage_bytes, 16, 1024);
        const uint32_t tmem_acc = tmem_base + as * 128u;
        uint64_t bdesc0 = bdesc_res;
        asm volatile("" : "+l"(bdesc0));  // opaque per tile: the 36 weight descriptors are rebuilt from one uniform
                                          // base by immediate adds instead of living in 72 hoisted vector registers
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint64_t a_t = adesc0 + (uint64_t)((uint32_t)P.shift[t] * 8u);  // descriptor address unit = 16 B
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ad = a_t + (uint64_t)((mt * (TC_BM * 128) + kk * 32) >> 4);
              const uint64_t bd = bdesc0 + (uint64_t)((t * 8192 + kk * 32) >> 4);
              umma_bf16(tmem_acc + mt * 64, ad, bd, idesc, (t > 0 || kk > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(empty_bar(s));
        umma_commit(tfull_bar(as));
      }
      __syncwarp();
      if (++s == (uint32_t)P.stages) {
        s = 0;
        sphase ^= 1u;
      }
    }
  } else {
    // =============================== epilogue (warps 0-7) ======================================
    const int quad = warp & 3, hsel = warp >> 2;
    const bool ts = P.tma_store != 0;
    const bool add_tma = ts && P.addend != nullptr && P.addend_tma != 0;
    const bool has_add = P.addend != nullptr && !add_tma;  // (per-lane global loads of the addend)
    const bool pre = has_add && P.addend_prefetch != 0;
    uint8_t* staging_ptr = smem_raw + (staging - raw);
    uint32_t it = 0;
    for (int w = blockIdx.x; w < P.total_tiles; w += gridDim.x, ++it) {
      const int img = w / P.tiles_per_img;
      const int y0 = (w - img * P.tiles_per_img) * P.R;
      const int view = img >= P.img_half ? 1 : 0;
      const uint32_t as = it & 1u;
      if (ts) {  // the previous work item's TMA store must have read the staging buffer before it is rewritten
        if (threadIdx.x == 0) {
          bulk_wait_read0();
          if (add_tma) {  // the residual-gradient tile of this work item: same box as the store, zero-filled padding
            mbar_expect_tx(add_bar, (uint32_t)(P.Wp * P.R * 128));
            tma_load_4d(staging, &tmAdd, add_bar, 0, 0, y0, img);
          }
        }
        named_bar_sync(1, 256);
      }
      // this thread's two accumulator rows (TMEM lane, +128 for the second MMA tile) -> output pixels
      bool valid2[2];
      long long pix2[2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int m = mt * TC_BM + quad * 32 + lane;
        const int r = m / P.Wp, x = m - r * P.Wp;
        valid2[mt] = r < P.R && x < P.W && y0 + r < P.H;
        pix2[mt] = ((long long)img * P.H + y0 + r) * P.W + x;
      }
      // residual-gradient addend (dgrad): both rows' 64 B are requested before the accumulator is complete, so the
      // global-load latency overlaps the MMAs instead of sitting between tcgen05.ld and the stores
      uint4 add2[2][4];
      uint32_t mw2[2] = {0xffffffffu, 0xffffffffu};  // addend mask bits of this thread's 32 channels of each row
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) add2[mt][qq] = make_uint4(0u, 0u, 0u, 0u);
      if (P.addend_mask != nullptr) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
          if (valid2[mt]) mw2[mt] = *reinterpret_cast<const uint32_t*>(P.addend_mask + pix2[mt] * 8 + hsel * 4);
      }
      if (pre) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
          if (valid2[mt]) {
            const uint4* src = reinterpret_cast<const uint4*>(P.addend + pix2[mt] * 64 + hsel * 32);
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) add2[mt][qq] = src[qq];
          }
      }
      mbar_wait(tfull_bar(as), (it >> 1) & 1u);
      tc_fence_after();
      if (add_tma) mbar_wait(add_bar, it & 1u);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const bool valid = valid2[mt];
        uint32_t v[32];
        tmem_ld32(tmem_base + as * 128u + mt * 64 + hsel * 32 + ((uint32_t)(quad * 32) << 16), v);
        tmem_ld_wait();
        if (do_stats) {
          float t[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) t[e] = valid ? __uint_as_float(v[e]) : 0.f;
          warp_col_reduce(t, lane);
          const float s1 = t[0];
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            const float f = valid ? __uint_as_float(v[e]) : 0.f;
            t[e] = f * f;
          }
          warp_col_reduce(t, lane);
          float* sp = stat + ((warp * 2 + view) * 2) * 32 + lane;
          sp[0] += s1;
          sp[32] += t[0];
        }
        if (valid || ts) {
          const long long pix = pix2[mt];
          const int m = mt * TC_BM + quad * 32 + lane;
          __nv_bfloat16* o = P.out + pix * 64 + hsel * 32;
          if (has_add && !pre && valid) {
            const uint4* src = reinterpret_cast<const uint4*>(P.addend + pix * 64 + hsel * 32);
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) add2[mt][qq] = src[qq];
          }
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[qq * 8 + e]);
            if (has_add) {
              float ad[8];
              Raw8h rw;
              rw.v = add2[mt][qq];
              cvt_raw(rw, ad);
#pragma unroll
              for (int e = 0; e < 8; ++e) f[e] += ((mw2[mt] >> (qq * 8 + e)) & 1u) ? ad[e] : 0.f;
            }
            if (ts) {  // staging row m, 16-byte chunk (hsel * 4 + qq) at its SWIZZLE_128B position
              uint8_t* sp = staging_ptr + m * 128 + (((hsel * 4 + qq) ^ (m & 7)) << 4);
              if (add_tma && valid) {  // the TMA load put this pixel's addend at the very same position
                float ad[8];
                Raw8h rw;
                rw.v = *reinterpret_cast<const uint4*>(sp);
                cvt_raw(rw, ad);
#pragma unroll
                for (int e = 0; e < 8; ++e) f[e] += ((mw2[mt] >> (qq * 8 + e)) & 1u) ? ad[e] : 0.f;
              }
              store8(reinterpret_cast<__nv_bfloat16*>(sp), f);
            } else {
              store8(o + qq * 8, f);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(as));
      if (ts) {
        fence_proxy_async();  // the generic-proxy writes above -> visible to the TMA unit
        named_bar_sync(1, 256);
        if (threadIdx.x == 0) {
          tma_store_4d(&tmO, staging, 0, 0, y0, img);
          bulk_commit();
        }
      }
    }
    if (ts && threadIdx.x == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 256);
  if (do_stats)
    for (int i = threadIdx.x; i < 4 * 64; i += HALO_THREADS) {  // i = (view, q, col); fixed warp order
      const int col = i & 63, q = (i >> 6) & 1, view = i >> 7;
      const int h = col >> 5, l = col & 31;
      float t = 0.f;
      for (int qd = 0; qd < 4; ++qd) t += stat[(((h * 4 + qd) * 2 + view) * 2 + q) * 32 + l];
      P.stat_partial[(long long)blockIdx.x * 4 * 64 + i] = t;
    }
}


// conv_halo_kernel specialised for fprop WITH BatchNorm statistics (option conv_halo_stats; not yet run on hardware, kept
// as a separate copy so that the validated kernel above stays byte-identical): every epilogue lane keeps running sums of
// its own accumulator rows (32 columns: sum, sum of squares) over all the work items of the CTA and the transpose-reduce
// across the warp runs once per CTA and view instead of twice per 32-column chunk -- those shuffles paced the fprop
// launches (36 % of the issue slots against 19 % for the dgrad, profiles/r01_ncu_halo.md).  No addend path.
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_fstats_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, HaloParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t resb = (raw + 1023u) & ~1023u;           // [9][64 x 128 B] resident weights
  const uint32_t staging = resb + 9u * 8192u;             // [256][128 B] bf16 output tile (tma_store only)
  const uint32_t base = staging + (P.tma_store ? HALO_STAGING_BYTES : 0u);  // stages
  const uint32_t bars = base + (uint32_t)P.stages * (uint32_t)P.stage_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };          // s < 4
  auto empty_bar = [&](int s) { return bars + 8u * (4 + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (8 + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (10 + a); };
  const uint32_t bres_bar = bars + 8u * 12;
  uint8_t* bars_ptr = smem_raw + (bars - raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars_ptr + 8 * 13);
  float* stat = reinterpret_cast<float*>(bars_ptr + 128);  // [8 warps][2 views][2][32]
  const bool do_stats = P.stat_partial != nullptr;
  if (do_stats)
    for (int i = threadIdx.x; i < 8 * 2 * 2 * 32; i += HALO_THREADS) stat[i] = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 256);
    }
    mbar_init(bres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (P.tma_store) tma_prefetch_desc(&tmO);
  }
  if (warp == 8) tmem_alloc(smem_u32(tmem_slot), 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    // =============================== TMA producer ============================================
    if (blockIdx.x < P.total_tiles) {
      if (lane == 0) mbar_expect_tx(bres_bar, 9u * 8192u);
      __syncwarp();
      if (lane < 9) tma_load_2d(resb + lane * 8192u, &tmB, bres_bar, (int)P.wtap[lane] * 64, 0);
    }
    if (lane == 0) {
      uint32_t it = 0;
      for (int w = blockIdx.x; w < P.total_tiles; w += gridDim.x, ++it) {
        const int img = w / P.tiles_per_img;
        const int y0 = (w - img * P.tiles_per_img) * P.R;
        const int s = it % P.stages;
        mbar_wait(empty_bar(s), ((it / P.stages) & 1u) ^ 1u);
        mbar_expect_tx(full_bar(s), (uint32_t)P.box_bytes);
        tma_load_4d(base + (uint32_t)s * (uint32_t)P.stage_bytes, &tmA, full_bar(s), 0, -1, y0 - 1, img);
      }
    }
  } else if (warp == 8) {
    // =============================== MMA issuer ==============================================
    // Everything that feeds a descriptor stays warp-uniform (wrapping stage counter instead of a modulo, the
    // nine tap shifts read from the kernel parameters by a fully unrolled loop): ptxas then keeps the
    // descriptors in uniform registers and issues UTCHMMA back to back; a descriptor that passes through a
    // vector register costs an ELECT / R2UR round trip per MMA, which at N = 64 (32 MMA cycles) is the bound.
    constexpr uint32_t idesc = make_idesc(64, 0, 0);
    uint32_t it = 0, s = 0, sphase = 0;
    if (blockIdx.x < P.total_tiles) mbar_wait(bres_bar, 0);
    const uint64_t bdesc_res = make_desc(resb, 16, 1024);
    for (int w = blockIdx.x; w < P.total_tiles; w += gridDim.x, ++it) {
      const uint32_t as = it & 1u;
      mbar_wait(tempty_bar(as), ((it >> 1) & 1u) ^ 1u);
      mbar_wait(full_bar(s), sphase);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint64_t adesc0 = make_desc(base + s * (uint32_t)P.stage_bytes, 16, 1024);
        const uint32_t tmem_acc = tmem_base + as * 128u;
        uint64_t bdesc0 = bdesc_res;
        asm volatile("" : "+l"(bdesc0));  // opaque per tile: the 36 weight descriptors are rebuilt from one uniform
                                          // base by immediate adds instead of living in 72 hoisted vector registers
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const uint64_t a_t = adesc0 + (uint64_t)((uint32_t)P.shift[t] * 8u);  // descriptor address unit = 16 B
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t ad = a_t + (uint64_t)((mt * (TC_BM * 128) + kk * 32) >> 4);
              const uint64_t bd = bdesc0 + (uint64_t)((t * 8192 + kk * 32) >> 4);
              umma_bf16(tmem_acc + mt * 64, ad, bd, idesc, (t > 0 || kk > 0) ? 1u : 0u);
            }
          }
        }
        umma_commit(empty_bar(s));
        umma_commit(tfull_bar(as));
      }
      __syncwarp();
      if (++s == (uint32_t)P.stages) {
        s = 0;
        sphase ^= 1u;
      }
    }
  } else {
    // =============================== epilogue (warps 0-7) ======================================
    const int quad = warp & 3, hsel = warp >> 2;
    const bool ts = P.tma_store != 0;
    uint8_t* staging_ptr = smem_raw + (staging - raw);
    // running per-lane column sums (this lane's rows, 32 columns) of the current view
    float rs1[32], rs2[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) rs1[e] = rs2[e] = 0.f;
    int cur_view = -1;
    auto flush_stats = [&](int vw) {
      float t[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) t[e] = rs1[e];
      warp_col_reduce(t, lane);
      const float s1 = t[0];
#pragma unroll
      for (int e = 0; e < 32; ++e) t[e] = rs2[e];
      warp_col_reduce(t, lane);
      float* sp = stat + ((warp * 2 + vw) * 2) * 32 + lane;
      sp[0] += s1;
      sp[32] += t[0];
#pragma unroll
      for (int e = 0; e < 32; ++e) rs1[e] = rs2[e] = 0.f;
    };
    uint32_t it = 0;
    for (int w = blockIdx.x; w < P.total_tiles; w += gridDim.x, ++it) {
      const int img = w / P.tiles_per_img;
      const int y0 = (w - img * P.tiles_per_img) * P.R;
      const int view = img >= P.img_half ? 1 : 0;
      const uint32_t as = it & 1u;
      if (cur_view >= 0 && view != cur_view) flush_stats(cur_view);  // (the work items of a CTA change view at most once)
      cur_view = view;
      if (ts) {  // the previous work item's TMA store must have read the staging buffer before it is rewritten
        if (threadIdx.x == 0) bulk_wait_read0();
        named_bar_sync(1, 256);
      }
      // this thread's two accumulator rows (TMEM lane, +128 for the second MMA tile) -> output pixels
      bool valid2[2];
      long long pix2[2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int m = mt * TC_BM + quad * 32 + lane;
        const int r = m / P.Wp, x = m - r * P.Wp;
        valid2[mt] = r < P.R && x < P.W && y0 + r < P.H;
        pix2[mt] = ((long long)img * P.H + y0 + r) * P.W + x;
      }
      mbar_wait(tfull_bar(as), (it >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const bool valid = valid2[mt];
        uint32_t v[32];
        tmem_ld32(tmem_base + as * 128u + mt * 64 + hsel * 32 + ((uint32_t)(quad * 32) << 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const float f = valid ? __uint_as_float(v[e]) : 0.f;
          rs1[e] += f;
          rs2[e] = fmaf(f, f, rs2[e]);
        }
        if (valid || ts) {
          const long long pix = pix2[mt];
          const int m = mt * TC_BM + quad * 32 + lane;
          __nv_bfloat16* o = P.out + pix * 64 + hsel * 32;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[qq * 8 + e]);
            if (ts)  // staging row m, 16-byte chunk (hsel * 4 + qq) at its SWIZZLE_128B position
              store8(reinterpret_cast<__nv_bfloat16*>(staging_ptr + m * 128 + (((hsel * 4 + qq) ^ (m & 7)) << 4)), f);
            else
              store8(o + qq * 8, f);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(as));
      if (ts) {
        fence_proxy_async();  // the generic-proxy writes above -> visible to the TMA unit
        named_bar_sync(1, 256);
        if (threadIdx.x == 0) {
          tma_store_4d(&tmO, staging, 0, 0, y0, img);
          bulk_commit();
        }
      }
    }
    if (cur_view >= 0) flush_stats(cur_view);
    if (ts && threadIdx.x == 0) bulk_wait0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc(tmem_base, 256);
  if (do_stats)
    for (int i = threadIdx.x; i < 4 * 64; i += HALO_THREADS) {  // i = (view, q, col); fixed warp order
      const int col = i & 63, q = (i >> 6) & 1, view = i >> 7;
      const int h = col >> 5, l = col & 31;
      float t = 0.f;
      for (int qd = 0; qd < 4; ++qd) t += stat[(((h * 4 + qd) * 2 + view) * 2 + q) * 32 + l];
      P.stat_partial[(long long)blockIdx.x * 4 * 64 + i] = t;
    }
}



// ---- halo wgrad: dW[tap][ci][co] = sum_p x[p + tap shift][ci] * dy[p][co] on the padded grid ----------------
// One work item = R output rows of one image: the x box (R+2 rows x Wp, zero-filled borders) and the dy box
// (R rows x Wp; its two padding columns are zero-filled, so padding positions contribute nothing) are loaded
// ONCE and all nine taps are accumulated from them: the A operand (MN-major, K = pixel) of an MMA holds two taps
// as its two 64-column blocks -- two shifted views of the x tile, LBO = their distance -- so five M = 128 MMAs
// per K step cover the nine taps (the tenth block is discarded).  Each persistent CTA keeps its 576 x 64 fp32
// partial in TMEM (5 x 64 columns) over all its work items and writes it once; wgrad2_reduce_kernel folds the
// per-CTA partials.  (The im2col kernel re-loads x nine times and dy five times: 192 B of shared-memory fill
// per tensor cycle at N = 64, which pinned this layer at ~0.3 of the tensor peak.)
__global__ void __launch_bounds__(TC2_THREADS, 1)
conv_halo_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy, HaloParams P,
                       float* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  constexpr uint32_t DY_BYTES = HALO_TILE * 128;
  const uint32_t stage_stride = (uint32_t)P.stage_bytes + DY_BYTES;  // [x tile | dy tile]
  const uint32_t bars = base + 2u * stage_stride;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (2 + s); };
  const uint32_t done_bar = bars + 8u * 4;
  uint8_t* base_ptr = smem_raw + (base - raw);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + 2u * stage_stride + 8 * 5);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // rows the TMA boxes never write but the MMAs read (K is padded to 256 pixels, the last taps look 2*Wp+2 rows
  // ahead): zero once -- dy = 0 there, and x must at least be finite
  for (int s = 0; s < 2; ++s) {
    uint4* xs = reinterpret_cast<uint4*>(base_ptr + s * stage_stride);
    for (int i = P.box_bytes / 16 + threadIdx.x; i < P.stage_bytes / 16; i += TC2_THREADS) xs[i] = make_uint4(0, 0, 0, 0);
    uint4* ds = reinterpret_cast<uint4*>(base_ptr + s * stage_stride + P.stage_bytes);
    for (int i = P.R * P.Wp * 8 + threadIdx.x; i < (int)(DY_BYTES / 16); i += TC2_THREADS) ds[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDy);
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 5) {
    if (lane == 0) {
      uint32_t it = 0;
      const uint32_t bytes = (uint32_t)P.box_bytes + (uint32_t)(P.R * P.Wp * 128);
      for (int w = blockIdx.x; w < P.total_tiles; w += gridDim.x, ++it) {
        const int img = w / P.tiles_per_img;
        const int y0 = (w - img * P.tiles_per_img) * P.R;
        const uint32_t s = it & 1u;
        mbar_wait(empty_bar(s), ((it >> 1) & 1u) ^ 1u);
        mbar_expect_tx(full_bar(s), bytes);
        tma_load_4d(base + s * stage_stride, &tmX, full_bar(s), 0, -1, y0 - 1, img);
        tma_load_4d(base + s * stage_stride + (uint32_t)P.stage_bytes, &tmDy, full_bar(s), 0, 0, y0, img);
      }
    }
  } else if (warp == 4) {
    constexpr uint32_t idesc = make_idesc(64, 1, 1);
    uint32_t it = 0;
    for (int w = blockIdx.x; w < P.total_tiles; w += gridDim.x, ++it) {
      const uint32_t s = it & 1u;
      mbar_wait(full_bar(s), (it >> 1) & 1u);
      tc_fence_after();
      if (elect_one_sync()) {
        const uint32_t xs = base + s * stage_stride;
        const uint64_t bd0 = make_desc(xs + (uint32_t)P.stage_bytes, 8192, 1024);
#pragma unroll
        for (int g = 0; g < 5; ++g) {
          // column blocks = taps 2g and 2g+1 (the "tap 9" of the last group is a dummy: rows >= 576 are dropped)
          const uint32_t sh0 = P.shift[2 * g];
          const uint32_t lbo = (g < 4 ? (uint32_t)P.shift[2 * g + 1] - sh0 : 1u) * 128u;
          const uint64_t ad0 = make_desc(xs + sh0 * 128u, lbo, 1024);
#pragma unroll
          for (int kk = 0; kk < HALO_TILE / 16; ++kk)
            umma_bf16(tmem_base + g * 64, ad0 + (uint64_t)(kk * 2048 >> 4), bd0 + (uint64_t)(kk * 2048 >> 4), idesc,
                      (it > 0 || kk > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));
      }
      __syncwarp();
    }
    if (elect_one_sync()) umma_commit(done_bar);
    __syncwarp();
  } else {
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
#pragma unroll 1
    for (int g = 0; g < 5; ++g) {
      const int j = g * 128 + row;  // row of the [576][64] partial = tap * 64 + ci
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + g * 64 + c0 + ((uint32_t)(warp * 32) << 16), v);
        tmem_ld_wait();
        if (j < 576) {
          float* o = partial + ((long long)blockIdx.x * 576 + j) * 64 + c0;
#pragma unroll
          for (int qq = 0; qq < 8; ++qq)
            *reinterpret_cast<float4*>(o + qq * 4) = make_float4(__uint_as_float(v[qq * 4]), __uint_as_float(v[qq * 4 + 1]),
                                                                 __uint_as_float(v[qq * 4 + 2]), __uint_as_float(v[qq * 4 + 3]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem_base, 512);
}

// ---- host side: tensor-map construction through the driver entry points ----------------------
typedef CUresult (*PFN_tmEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                      const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*PFN_tmEncodeIm2col)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave,
                                       CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmEncodeTiled g_encodeTiled = nullptr;
static PFN_tmEncodeIm2col g_encodeIm2col = nullptr;
static int g_driver_version = 0;

static int tma_init() {
  if (g_encodeTiled && g_encodeIm2col) return IIC_OK;
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  IIC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  IIC_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, IIC_ERR_CUDA, "cuTensorMapEncodeTiled unavailable");
  g_encodeTiled = (PFN_tmEncodeTiled)fn;
  fn = nullptr;
  IIC_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres));
  IIC_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, IIC_ERR_CUDA, "cuTensorMapEncodeIm2col unavailable");
  g_encodeIm2col = (PFN_tmEncodeIm2col)fn;
  cudaDriverGetVersion(&g_driver_version);
  return IIC_OK;
}

// NHWC bf16 activation [nimg][H][W][C] as the rank-4 (C, W, H, N) im2col tensor map
static int make_im2col_map2(CUtensorMap* tm, const void* ptr, int nimg, int H, int W, int C, int lower_w, int lower_h,
                            int upper_w, int upper_h, int stride, int pixels);

static int make_im2col_map(CUtensorMap* tm, const void* ptr, int nimg, int H, int W, int C, int lower, int upper, int stride,
                           int pixels) {
  return make_im2col_map2(tm, ptr, nimg, H, W, C, lower, lower, upper, upper, stride, pixels);
}

static int make_im2col_map2(CUtensorMap* tm, const void* ptr, int nimg, int H, int W, int C, int lower_w, int lower_h,
                            int upper_w, int upper_h, int stride, int pixels) {
  const int lower = lower_w, upper = upper_w;
  cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
  cuuint64_t gstr[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lo[2] = {lower_w, lower_h}, up[2] = {upper_w, upper_h};
  cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  CUresult r = g_encodeIm2col(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, lo, up, 64,
                              (cuuint32_t)pixels, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IIC_REQUIRE(r == CUDA_SUCCESS, IIC_ERR_CUDA, "cuTensorMapEncodeIm2col failed (%d) H=%d W=%d C=%d lower=%d upper=%d stride=%d",
              (int)r, H, W, C, lower, upper, stride);
  // driver workaround also applied by CUTLASS (copy_traits_sm90_im2col.hpp): small tensors, drivers <= 13.1
  if (g_driver_version <= 13010 && (long long)nimg * H * W * C * 2 < 131072)
    reinterpret_cast<uint64_t*>(tm)[1] &= ~(1llu << 21);
  return IIC_OK;
}

static int tc2_grid(long long work) { return (int)(work < device_sm_count() ? work : device_sm_count()); }

template <int MODE, int BN, bool RESB, int MTP = (RESB ? 2 : 1)>
static int launch_tc2_impl(const CUtensorMap& tmA, const CUtensorMap& tmB, const Tc2Params& P, int splits, cudaStream_t st) {
  using Cfg = Tc2Cfg<BN, RESB, MTP>;
  const int smem = Cfg::SMEM + (P.stat_partial ? 64 * P.N : 0) + (RESB ? P.total_kb * Cfg::B_BYTES : 0);
  IIC_REQUIRE(smem <= 232448, IIC_ERR_UNSUPPORTED, "conv_tc2: shared memory budget exceeded (%d B)", smem);
  IIC_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<MODE, BN, RESB, MTP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  long long work = (long long)P.mtiles * P.ntiles * splits;
  int grid = tc2_grid(work);
  conv_tc2_kernel<MODE, BN, RESB, MTP><<<grid, TC2_THREADS, smem, st>>>(tmA, tmB, P);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

template <int MODE, int BN>
static int launch_tc2(const CUtensorMap& tmA, const CUtensorMap& tmB, const Tc2Params& P, int splits, cudaStream_t st) {
  return launch_tc2_impl<MODE, BN, false>(tmA, tmB, P, splits, st);
}

// resident-weights / two-tile variant: N == 64 (one N tile), whole packed weight <= 80 KB, many tiles per CTA
static bool tc2_use_resb(int bn, int N, int total_kb, long long rows) {
  return bn == 64 && N == 64 && total_kb * (64 * 128) <= 80 * 1024 && (rows + TC_BM - 1) / TC_BM > 2ll * device_sm_count();
}

static int pick_bn2(int N) {
  if (N % 256 == 0) return 256;
  if (N % 128 == 0) return 128;
  if (N % 64 == 0) return 64;
  return 0;
}

// two 128-row tiles per work item with streamed weights (option tc2_mt2): N = 128 tiles with enough work per CTA
// (1 = when every CTA gets at least two such work items, 2 = always: tests)
static bool tc2_use_mt2(int bn, int N, long long rows) {
  const int mode = option(OPT_TC2_MT2);
  if (mode == 0 || bn != 128) return false;
  return mode >= 2 || ((rows + 2 * TC_BM - 1) / (2 * TC_BM)) * (N / bn) >= 2ll * device_sm_count();
}


// ---- halo variant: plan + launch -----------------------------------------------------------------
struct HaloPlan {
  bool ok;
  int Wp, R, tiles_per_img, total_tiles, stage_bytes, box_bytes, stages, smem;
  int tma_store;  // fprop / dgrad: output tile through a shared-memory staging buffer and one TMA store
};

// option conv_halo / IIC_CONV_HALO: 0 never, 1 (default) when the geometry fits and the tile efficiency is good,
// 2 whenever the geometry fits (tests)
static int halo_mode() { return option(OPT_CONV_HALO); }

static HaloPlan halo_plan(const iic_conv_geom* g, int srcC, int N, int H, int W, int nimg) {
  HaloPlan p = {};
  if (halo_mode() == 0) return p;
  if (!(g->kh == 3 && g->kw == 3 && g->stride == 1 && g->pad == 1 && g->dil == 1 && srcC == 64 && N == 64)) return p;
  p.Wp = W + 2;
  p.R = HALO_TILE / p.Wp;
  if (p.R < 1) return p;
  p.tiles_per_img = (H + p.R - 1) / p.R;
  const long long total = (long long)nimg * p.tiles_per_img;
  if (total > 0x7fffffffll) return p;
  p.total_tiles = (int)total;
  p.box_bytes = p.Wp * (p.R + 2) * 128;
  p.stage_bytes = ((HALO_TILE + 2 * p.Wp + 2) * 128 + 1023) / 1024 * 1024;
  if (p.box_bytes > p.stage_bytes || p.R + 2 > 256) return p;
  const int fixed = 1024 + 9 * 8192 + 128 + 8 * 2 * 2 * 32 * 4;
  // with the staging buffer of the TMA-store epilogue (option conv_halo_store) if two pipeline stages still fit
  p.tma_store = option(OPT_HALO_STORE) != 0 && (232448 - fixed - (int)HALO_STAGING_BYTES) / p.stage_bytes >= 2;
  const int fixed_all = fixed + (p.tma_store ? (int)HALO_STAGING_BYTES : 0);
  p.stages = (232448 - fixed_all) / p.stage_bytes;
  if (p.stages > 4) p.stages = 4;
  if (p.stages < 2) return p;
  p.smem = fixed_all + p.stages * p.stage_bytes;
  if (halo_mode() == 1) {
    const double eff = (double)H * W / ((double)p.tiles_per_img * HALO_TILE);
    if (eff < 0.80 || p.total_tiles < 2 * device_sm_count()) return p;
  }
  p.ok = true;
  return p;
}

static int make_halo_map(CUtensorMap* tm, const __nv_bfloat16* src, int H, int W, int nimg, int box_w, int box_h) {
  cuuint64_t gdim[4] = {64, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)nimg};
  cuuint64_t gstr[3] = {128, (cuuint64_t)W * 128, (cuuint64_t)H * W * 128};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = g_encodeTiled(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<__nv_bfloat16*>(src), gdim, gstr, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  IIC_REQUIRE(r == CUDA_SUCCESS, IIC_ERR_CUDA, "cuTensorMapEncodeTiled(halo) failed (%d) H=%d W=%d box %dx%d", (int)r, H, W, box_w,
              box_h);
  return IIC_OK;
}

// wgrad flavour of the plan: two [x tile | dy tile] stages; the per-CTA partials are the split-K workspace
static HaloPlan halo_wgrad_plan(const iic_conv_geom* g) {
  HaloPlan p = halo_plan(g, g->cin, g->cout, g->h, g->w, g->n);
  if (!p.ok) return p;
  if (option(OPT_CONV_HALO_WGRAD) == 0) {
    p.ok = false;
    return p;
  }
  p.stages = 2;
  p.smem = 1024 + 2 * (p.stage_bytes + HALO_TILE * 128) + 256;
  if (p.smem > 232448) p.ok = false;
  return p;
}

static int launch_halo(const HaloPlan& hp, const __nv_bfloat16* src, int H, int W, int nimg, const Tc2Params& P,
                       const CUtensorMap& tmB, cudaStream_t st) {
  alignas(64) CUtensorMap tmA, tmO, tmAdd;
  const bool add_tma = P.addend != nullptr && hp.tma_store && option(OPT_HALO_ADDEND_TMA) != 0;
  {
    int rc = make_halo_map(&tmA, src, H, W, nimg, hp.Wp, hp.R + 2);
    if (rc != IIC_OK) return rc;
    // output map: the load box without the two halo rows; unused (but valid) when the epilogue stores directly
    rc = make_halo_map(&tmO, P.out, H, W, nimg, hp.Wp, hp.R);
    if (rc != IIC_OK) return rc;
    // the addend has the geometry of the output (unused copy of the output map when there is none)
    rc = make_halo_map(&tmAdd, add_tma ? P.addend : P.out, H, W, nimg, hp.Wp, hp.R);
    if (rc != IIC_OK) return rc;
  }
  HaloParams Q = {};
  Q.nimg = nimg; Q.H = H; Q.W = W; Q.Wp = hp.Wp; Q.R = hp.R;
  Q.tiles_per_img = hp.tiles_per_img; Q.total_tiles = hp.total_tiles;
  Q.stage_bytes = hp.stage_bytes; Q.box_bytes = hp.box_bytes; Q.stages = hp.stages;
  for (int t = 0; t < 9; ++t) {
    Q.shift[t] = (unsigned short)((int)P.offh[t] * hp.Wp + (int)P.offw[t]);
    Q.wtap[t] = P.wtap[t];
  }
  Q.out = P.out; Q.addend = P.addend; Q.addend_mask = P.addend_mask; Q.stat_partial = P.stat_partial;
  Q.addend_prefetch = P.addend_prefetch;
  Q.tma_store = hp.tma_store;
  Q.addend_tma = add_tma ? 1 : 0;
  Q.img_half = (P.stat_half < P.rows) ? nimg / 2 : nimg;
  if (option(OPT_HALO_STATS) != 0 && Q.stat_partial != nullptr && Q.addend == nullptr) {
    IIC_CUDA(cudaFuncSetAttribute(conv_halo_fstats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hp.smem));
    conv_halo_fstats_kernel<<<tc2_grid(hp.total_tiles), HALO_THREADS, hp.smem, st>>>(tmA, tmB, tmO, Q);
  } else {
    IIC_CUDA(cudaFuncSetAttribute(conv_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hp.smem));
    conv_halo_kernel<<<tc2_grid(hp.total_tiles), HALO_THREADS, hp.smem, st>>>(tmA, tmB, tmO, tmAdd, Q);
  }
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

// fprop (transposed == 0) or dgrad of a stride-1 conv (transposed == 1; src = dy, N = cin)
int tc2_conv_fprop_blocks(const iic_conv_geom* g) {
  const int bn = pick_bn2(g->cout);
  if (bn == 0) return 0;
  const long long rows = (long long)g->n * g->oh * g->ow;
  const HaloPlan hp = halo_plan(g, g->cin, g->cout, g->h, g->w, g->n);
  if (hp.ok) return tc2_grid(hp.total_tiles);
  const int tile_rows =
      (tc2_use_resb(bn, g->cout, g->kh * g->kw * g->cin / 64, rows) || tc2_use_mt2(bn, g->cout, rows)) ? 2 * TC_BM : TC_BM;
  return tc2_grid(((rows + tile_rows - 1) / tile_rows) * (g->cout / bn));
}

int tc2_conv_gather_gemm_stats(const __nv_bfloat16* src, int srcH, int srcW, int srcC, int rowH, int rowW, int nimg,
                               const iic_conv_geom* g, int transposed, const __nv_bfloat16* wpacked, int N,
                               const __nv_bfloat16* addend, __nv_bfloat16* out, float* stat_partial, int stat_groups,
                               cudaStream_t st);

int tc2_conv_gather_gemm(const __nv_bfloat16* src, int srcH, int srcW, int srcC, int rowH, int rowW, int nimg,
                         const iic_conv_geom* g, int transposed, const __nv_bfloat16* wpacked, int N,
                         const __nv_bfloat16* addend, __nv_bfloat16* out, cudaStream_t st) {
  return tc2_conv_gather_gemm_stats(src, srcH, srcW, srcC, rowH, rowW, nimg, g, transposed, wpacked, N, addend, out, nullptr, 1,
                                    st);
}

static int tc2_gather_gemm_impl(const __nv_bfloat16* src, int srcH, int srcW, int srcC, int rowH, int rowW, int nimg,
                                const iic_conv_geom* g, int transposed, const __nv_bfloat16* wpacked, int N,
                                const __nv_bfloat16* addend, const unsigned char* addend_mask, __nv_bfloat16* out,
                                float* stat_partial, int stat_groups, cudaStream_t st);

int tc2_conv_gather_gemm_stats(const __nv_bfloat16* src, int srcH, int srcW, int srcC, int rowH, int rowW, int nimg,
                               const iic_conv_geom* g, int transposed, const __nv_bfloat16* wpacked, int N,
                               const __nv_bfloat16* addend, __nv_bfloat16* out, float* stat_partial, int stat_groups,
                               cudaStream_t st) {
  return tc2_gather_gemm_impl(src, srcH, srcW, srcC, rowH, rowW, nimg, g, transposed, wpacked, N, addend, nullptr, out,
                              stat_partial, stat_groups, st);
}

// stride-1 dgrad whose addend passes only where its mask bit is set: dx = dgrad(dy) + (bit ? addend : 0).  The residual
// gradient of a BasicBlock is d_out * (out > 0): with the ReLU mask of the block output kept as bits by the forward pass,
// the BatchNorm backward no longer has to write the masked copy (2 B per element) for this epilogue to read back.
int tc2_conv_dgrad_masked(const __nv_bfloat16* dy, const __nv_bfloat16* wpacked_t, const __nv_bfloat16* addend,
                          const unsigned char* addend_mask, __nv_bfloat16* dx, const iic_conv_geom* g, cudaStream_t st) {
  return tc2_gather_gemm_impl(dy, g->oh, g->ow, g->cout, g->h, g->w, g->n, g, 1, wpacked_t, g->cin, addend, addend_mask, dx,
                              nullptr, 1, st);
}

static int tc2_gather_gemm_impl(const __nv_bfloat16* src, int srcH, int srcW, int srcC, int rowH, int rowW, int nimg,
                                const iic_conv_geom* g, int transposed, const __nv_bfloat16* wpacked, int N,
                                const __nv_bfloat16* addend, const unsigned char* addend_mask, __nv_bfloat16* out,
                                float* stat_partial, int stat_groups, cudaStream_t st) {
  int rc = tma_init();
  if (rc != IIC_OK) return rc;
  const int bn = pick_bn2(N);
  IIC_REQUIRE(bn != 0 && srcC % 64 == 0, IIC_ERR_UNSUPPORTED, "tcgen05/TMA conv: channels must be multiples of 64");
  IIC_REQUIRE(g->kh == g->kw, IIC_ERR_UNSUPPORTED, "tcgen05/TMA conv: square filters only");
  IIC_REQUIRE(!transposed || g->stride == 1, IIC_ERR_UNSUPPORTED, "tcgen05/TMA dgrad: stride-1 only");
  Tc2Params P = {};
  P.rows = (long long)nimg * rowH * rowW;
  P.rowH = rowH; P.rowW = rowW; P.KH = g->kh; P.KW = g->kw; P.d = g->dil;
  const int span = (g->kh - 1) * g->dil;
  IIC_REQUIRE(g->kh * g->kw <= TC2_MAXTAPS && span <= 255, IIC_ERR_UNSUPPORTED, "tcgen05/TMA conv: filter too large");
  int upper;
  if (!transposed) {
    P.s = g->stride; P.lower = -g->pad;
    upper = g->pad - span;
  } else {
    P.s = 1; P.lower = g->pad - span;
    upper = P.lower + (rowH - srcH);  // number of base positions == rows of dx
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
  P.ntiles = N / bn; P.splits = 1; P.total_kb = P.Ktot / 64;
  const bool resb = tc2_use_resb(bn, N, P.total_kb, P.rows);
  const bool mt2 = !resb && tc2_use_mt2(bn, N, P.rows);
  const int tile_rows = (resb || mt2) ? 2 * TC_BM : TC_BM;
  P.mtiles = (int)((P.rows + tile_rows - 1) / tile_rows);
  P.out = out; P.addend = addend; P.addend_mask = addend_mask;
  P.addend_prefetch = option(OPT_DGRAD_PREFETCH);
  P.stat_partial = stat_partial;
  IIC_REQUIRE(stat_groups == 1 || (stat_groups == 2 && nimg % 2 == 0), IIC_ERR_BAD_ARG, "conv stats: 1 or 2 views");
  P.stat_half = stat_groups == 2 ? P.rows / 2 : P.rows;
  alignas(64) CUtensorMap tmA, tmB;
  const HaloPlan hp = halo_plan(g, srcC, N, srcH, srcW, nimg);
  if (!hp.ok) {
    rc = make_im2col_map(&tmA, src, nimg, srcH, srcW, srcC, P.lower, upper, P.s, tile_rows);
    if (rc != IIC_OK) return rc;
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)P.Ktot, (cuuint64_t)N};
    cuuint64_t gstr[1] = {(cuuint64_t)P.Ktot * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encodeTiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(wpacked), gdim, gstr, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IIC_REQUIRE(r == CUDA_SUCCESS, IIC_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed (%d)", (int)r);
  }
  if (hp.ok) return launch_halo(hp, src, srcH, srcW, nimg, P, tmB, st);
  if (resb) return launch_tc2_impl<M2_FPROP, 64, true>(tmA, tmB, P, 1, st);
  if (mt2) return launch_tc2_impl<M2_FPROP, 128, false, 2>(tmA, tmB, P, 1, st);
  switch (bn) {
    case 256: return launch_tc2<M2_FPROP, 256>(tmA, tmB, P, 1, st);
    case 128: return launch_tc2<M2_FPROP, 128>(tmA, tmB, P, 1, st);
    default: return launch_tc2<M2_FPROP, 64>(tmA, tmB, P, 1, st);
  }
}

// 128-row (tap, cin) tiles per wgrad work item (option wgrad_mt): the dy k-block is loaded once for MT of them, which takes
// a third off the shared-memory fill per MMA that bounds these launches (xbar -> SM ~ 68 B/clk/SM, profiles/r02_session_f.md).
// Only exact tilings (no half-empty last tile): 9 x 128 channels = 3 x 384 rows, 9 x 256 = 9 x 256 rows, ...
static int tc2_wgrad_mt(const iic_conv_geom* g, int bn) {
  if (option(OPT_WGRAD_MT) == 0) return 1;
  const int Ktot = g->kh * g->kw * g->cin;
  // (at least two work items along M: a single 2-tile item was slower than two 1-tile items, 1x1 256 -> 512: 51 -> 63 us)
  if (bn == 128) return (Ktot % 384 == 0 && Ktot >= 768) ? 3 : ((Ktot % 256 == 0 && Ktot >= 512) ? 2 : 1);
  if (bn == 256) return (Ktot % 256 == 0 && Ktot >= 512) ? 2 : 1;
  return 1;
}

static int tc2_wgrad_splits(const iic_conv_geom* g) {
  const long long rows = (long long)g->n * g->oh * g->ow;
  const int total_kb = (int)((rows + 63) / 64);
  const int Ktot = g->kh * g->kw * g->cin;
  const int bn = pick_bn2(g->cout);
  const int mrows = TC_BM * tc2_wgrad_mt(g, bn ? bn : 64);
  const long long tiles = (long long)((Ktot + mrows - 1) / mrows) * (g->cout / (bn ? bn : 64));
  // persistent CTAs take work items round-robin: make (tiles x splits) fill a whole number of rounds of
  // one item per SM (never 2.1 rounds), with at least 8 k-blocks per item
  const long long sms = device_sm_count();
  long long want = 1;
  if (tiles < sms) {
    want = sms / tiles;
    const long long two = (2 * sms) / tiles;  // two full rounds if the items stay long enough
    if (two > want && total_kb / two >= 32) want = two;
  }
  if (want > total_kb / 8) want = total_kb / 8;
  if (want < 1) want = 1;
  if (want > 512) want = 512;
  return (int)want;
}

// dgrad of a stride-2 convolution, decomposed by output parity: dx[2i+py, 2j+px] only receives the taps a
// with (py + pad - a) even, read at dy[i + (py + pad - a)/2]; each of the four classes is a stride-1 im2col
// problem over dy with its own tap subset (1+2+2+4 = 9 taps for 3x3: no wasted MMAs, unlike the zero-filled
// fractional-stride gather of round 1's first kernel).  w = kind-1 packed weight [cin][kh][kw][cout].
int tc2_conv_dgrad_s2(const __nv_bfloat16* dy, const __nv_bfloat16* wpacked_t, const __nv_bfloat16* addend,
                      __nv_bfloat16* dx, const iic_conv_geom* g, cudaStream_t st) {
  int rc = tma_init();
  if (rc != IIC_OK) return rc;
  const int bn = pick_bn2(g->cin);
  IIC_REQUIRE(bn != 0 && g->cout % 64 == 0 && g->stride == 2 && g->dil == 1 && g->kh == g->kw && g->kh * g->kw <= TC2_MAXTAPS,
              IIC_ERR_UNSUPPORTED, "tcgen05/TMA stride-2 dgrad: unsupported geometry");
  alignas(64) CUtensorMap tmB;
  const int Kw = g->kh * g->kw * g->cout;
  {
    cuuint64_t gdim[2] = {(cuuint64_t)Kw, (cuuint64_t)g->cin};
    cuuint64_t gstr[1] = {(cuuint64_t)Kw * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encodeTiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(wpacked_t), gdim, gstr, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IIC_REQUIRE(r == CUDA_SUCCESS, IIC_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed (%d)", (int)r);
  }
  auto parity_taps = [&](int par, int ksz) {
    int cnt = 0;
    for (int a = 0; a < ksz; ++a) cnt += (((par + g->pad - a) % 2) + 2) % 2 == 0;
    return cnt;
  };
  bool need_zero = false;
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px)
      if (parity_taps(py, g->kh) * parity_taps(px, g->kw) == 0) need_zero = true;
  if (need_zero) {  // classes that no tap reaches (e.g. 1x1 stride 2): dx = addend or 0 there
    if (addend != nullptr)
      IIC_CUDA(cudaMemcpyAsync(dx, addend, sizeof(__nv_bfloat16) * (size_t)g->n * g->h * g->w * g->cin, cudaMemcpyDeviceToDevice, st));
    else
      IIC_CUDA(cudaMemsetAsync(dx, 0, sizeof(__nv_bfloat16) * (size_t)g->n * g->h * g->w * g->cin, st));
  }
  for (int py = 0; py < 2; ++py)
    for (int px = 0; px < 2; ++px) {
      const int Hc = (g->h - py + 1) / 2, Wc = (g->w - px + 1) / 2;
      if (Hc <= 0 || Wc <= 0) continue;
      Tc2Params P = {};
      // taps and their dy offsets: oy = i + (py + pad - a)/2 ; shift so that the smallest offset is the lower corner
      int offs_h[8], taps_h[8], nh = 0, offs_w[8], taps_w[8], nw = 0;
      for (int a = 0; a < g->kh; ++a) {
        const int tnum = py + g->pad - a;
        if (((tnum % 2) + 2) % 2 == 0) { offs_h[nh] = tnum / 2; taps_h[nh++] = a; }  // tnum even: exact
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
      P.ntiles = g->cin / bn; P.splits = 1; P.total_kb = P.Ktot / 64;
      // option dgrad_s2_mt: the parity classes take the resident-weights / two-tile variants of the stride-1 path (N = 64:
      // the class's whole weight slice stays in shared memory, 24 -> 16 KB of fill per MMA set; N = 128: two tiles per
      // weight k-block); 2 = whatever the amount of work (tests)
      const int s2mt = option(OPT_DGRAD_S2_MT);
      const bool resb = s2mt != 0 && bn == 64 && g->cin == 64 && P.total_kb * (64 * 128) <= 80 * 1024 &&
                        (s2mt >= 2 || tc2_use_resb(bn, g->cin, P.total_kb, P.rows));
      const bool mt2 = s2mt != 0 && !resb && tc2_use_mt2(bn, g->cin, P.rows);
      const int tile_rows = (resb || mt2) ? 2 * TC_BM : TC_BM;
      P.mtiles = (int)((P.rows + tile_rows - 1) / tile_rows);
      P.out = dx; P.addend = addend;  // (classes without taps keep the pre-filled addend / zero)
      P.addend_prefetch = option(OPT_DGRAD_PREFETCH);
      // base positions per dim must number Hc / Wc: upper = lower + (Hc - H_dy)
      const int up_h = lo_h + (Hc - g->oh), up_w = lo_w + (Wc - g->ow);
      IIC_REQUIRE(lo_h >= -128 && up_h <= 127 && lo_w >= -128 && up_w <= 127 && up_h >= -128 && up_w >= -128,
                  IIC_ERR_UNSUPPORTED, "im2col corner range");
      alignas(64) CUtensorMap tmA;
      rc = make_im2col_map2(&tmA, dy, g->n, g->oh, g->ow, g->cout, lo_w, lo_h, up_w, up_h, 1, tile_rows);
      if (rc != IIC_OK) return rc;
      if (resb) rc = launch_tc2_impl<M2_FPROP, 64, true>(tmA, tmB, P, 1, st);
      else if (mt2) rc = launch_tc2_impl<M2_FPROP, 128, false, 2>(tmA, tmB, P, 1, st);
      else switch (bn) {
        case 256: rc = launch_tc2<M2_FPROP, 256>(tmA, tmB, P, 1, st); break;
        case 128: rc = launch_tc2<M2_FPROP, 128>(tmA, tmB, P, 1, st); break;
        default: rc = launch_tc2<M2_FPROP, 64>(tmA, tmB, P, 1, st); break;
      }
      if (rc != IIC_OK) return rc;
    }
  return IIC_OK;
}

long long tc2_conv_wgrad_workspace(const iic_conv_geom* g) {
  const HaloPlan hp = halo_wgrad_plan(g);
  const long long splits = hp.ok ? tc2_grid(hp.total_tiles) : tc2_wgrad_splits(g);
  return splits * g->kh * g->kw * g->cin * g->cout * (long long)sizeof(float);
}

__global__ void wgrad2_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int Ktot, int N, int splits) {
  const long long total = (long long)Ktot * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % N);
    const int j = (int)(i / N);
    float t = 0.f;
    for (int z = 0; z < splits; ++z) t += partial[(long long)z * total + i];
    dw[(long long)co * Ktot + j] = t;
  }
}

// Split-K fold that writes the torch layout [cout][cin][kh][kw] directly (optionally accumulating): the separate
// iic_unpack_wgrad pass (one more launch and one more read + write of every gradient) disappears.
__global__ void wgrad2_reduce_unpack_kernel(const float* __restrict__ partial, float* __restrict__ grad, int Ktot, int N,
                                            int splits, int cin, int taps, int accumulate) {
  const long long total = (long long)Ktot * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i % N);
    const int j = (int)(i / N);  // (tap, ci)
    float t = 0.f;
    for (int z = 0; z < splits; ++z) t += partial[(long long)z * total + i];
    const int tap = j / cin, ci = j - tap * cin;
    const long long o = ((long long)co * cin + ci) * taps + tap;
    grad[o] = accumulate ? grad[o] + t : t;
  }
}

int tc2_conv_wgrad_impl(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, float* ws, const iic_conv_geom* g,
                        cudaStream_t st, float* grad_oihw, int accumulate);

int tc2_conv_wgrad(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, float* ws, const iic_conv_geom* g, cudaStream_t st) {
  return tc2_conv_wgrad_impl(x, dy, dw, ws, g, st, nullptr, 0);
}

int tc2_conv_wgrad_oihw(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* grad_oihw, int accumulate, float* ws,
                        const iic_conv_geom* g, cudaStream_t st) {
  return tc2_conv_wgrad_impl(x, dy, nullptr, ws, g, st, grad_oihw, accumulate);
}

int tc2_conv_wgrad_impl(const __nv_bfloat16* x, const __nv_bfloat16* dy, float* dw, float* ws, const iic_conv_geom* g,
                        cudaStream_t st, float* grad_oihw, int accumulate) {
  int rc = tma_init();
  if (rc != IIC_OK) return rc;
  const int bn = pick_bn2(g->cout);
  IIC_REQUIRE(bn != 0 && g->cin % 64 == 0, IIC_ERR_UNSUPPORTED, "tcgen05/TMA wgrad: channels must be multiples of 64");
  IIC_REQUIRE(g->kh == g->kw, IIC_ERR_UNSUPPORTED, "tcgen05/TMA wgrad: square filters only");
  const HaloPlan hp = halo_wgrad_plan(g);
  if (hp.ok) {
    alignas(64) CUtensorMap tmX, tmDy;
    rc = make_halo_map(&tmX, x, g->h, g->w, g->n, hp.Wp, hp.R + 2);
    if (rc != IIC_OK) return rc;
    rc = make_halo_map(&tmDy, dy, g->h, g->w, g->n, hp.Wp, hp.R);
    if (rc != IIC_OK) return rc;
    HaloParams Q = {};
    Q.nimg = g->n; Q.H = g->h; Q.W = g->w; Q.Wp = hp.Wp; Q.R = hp.R;
    Q.tiles_per_img = hp.tiles_per_img; Q.total_tiles = hp.total_tiles;
    Q.stage_bytes = hp.stage_bytes; Q.box_bytes = hp.box_bytes; Q.stages = 2;
    for (int t = 0; t < 9; ++t) Q.shift[t] = (unsigned short)((t / 3) * hp.Wp + t % 3);
    const int grid = tc2_grid(hp.total_tiles);
    IIC_CUDA(cudaFuncSetAttribute(conv_halo_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, hp.smem));
    conv_halo_wgrad_kernel<<<grid, TC2_THREADS, hp.smem, st>>>(tmX, tmDy, Q, ws);
    IIC_LAUNCH_CHECK();
    count_launch();
    const long long total = 576ll * 64;
    if (grad_oihw != nullptr)
      wgrad2_reduce_unpack_kernel<<<cdiv(total, 256), 256, 0, st>>>(ws, grad_oihw, 576, 64, grid, 64, 9, accumulate);
    else
      wgrad2_reduce_kernel<<<cdiv(total, 256), 256, 0, st>>>(ws, dw, 576, 64, grid);
    IIC_LAUNCH_CHECK();
    count_launch();
    return IIC_OK;
  }
  Tc2Params P = {};
  P.rows = (long long)g->n * g->oh * g->ow;
  P.rowH = g->oh; P.rowW = g->ow; P.KH = g->kh; P.KW = g->kw; P.s = g->stride; P.d = g->dil; P.lower = -g->pad;
  const int upper = g->pad - (g->kh - 1) * g->dil;
  P.srcC = g->cin; P.Ktot = g->kh * g->kw * g->cin; P.N = g->cout;
  const int mt = tc2_wgrad_mt(g, bn);
  P.mtiles = (P.Ktot + mt * TC_BM - 1) / (mt * TC_BM); P.ntiles = g->cout / bn;
  P.splits = tc2_wgrad_splits(g);
  P.total_kb = (int)((P.rows + 63) / 64);
  P.kb_per_split = (P.total_kb + P.splits - 1) / P.splits;
  P.partial = ws;
  alignas(64) CUtensorMap tmA, tmB;
  rc = make_im2col_map(&tmA, x, g->n, g->h, g->w, g->cin, P.lower, upper, P.s, 64);
  if (rc != IIC_OK) return rc;
  {
    // dy [rows][cout]: one 64-channel x 64-pixel box per column block; each lands in shared memory as
    // [pixel][128 B] = one MN-major SWIZZLE_128B atom column (bn/64 loads per stage).
    cuuint64_t gdim[2] = {(cuuint64_t)g->cout, (cuuint64_t)P.rows};
    cuuint64_t gstr[1] = {(cuuint64_t)g->cout * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encodeTiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(dy), gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    IIC_REQUIRE(r == CUDA_SUCCESS, IIC_ERR_CUDA, "cuTensorMapEncodeTiled(dy) failed (%d)", (int)r);
  }
  switch (bn) {
    case 256:
      rc = mt == 2 ? launch_tc2_impl<M2_WGRAD, 256, false, 2>(tmA, tmB, P, P.splits, st)
                   : launch_tc2<M2_WGRAD, 256>(tmA, tmB, P, P.splits, st);
      break;
    case 128:
      rc = mt == 3 ? launch_tc2_impl<M2_WGRAD, 128, false, 3>(tmA, tmB, P, P.splits, st)
                   : (mt == 2 ? launch_tc2_impl<M2_WGRAD, 128, false, 2>(tmA, tmB, P, P.splits, st)
                              : launch_tc2<M2_WGRAD, 128>(tmA, tmB, P, P.splits, st));
      break;
    default: rc = launch_tc2<M2_WGRAD, 64>(tmA, tmB, P, P.splits, st); break;
  }
  if (rc != IIC_OK) return rc;
  const long long total = (long long)P.Ktot * g->cout;
  int blocks = cdiv(total, 256);
  if (blocks > device_sm_count() * 8) blocks = device_sm_count() * 8;
  if (grad_oihw != nullptr)
    wgrad2_reduce_unpack_kernel<<<blocks, 256, 0, st>>>(ws, grad_oihw, P.Ktot, g->cout, P.splits, g->cin, g->kh * g->kw, accumulate);
  else
    wgrad2_reduce_kernel<<<blocks, 256, 0, st>>>(ws, dw, P.Ktot, g->cout, P.splits);
  IIC_LAUNCH_CHECK();
  count_launch();
  return IIC_OK;
}

}  // namespace iic
