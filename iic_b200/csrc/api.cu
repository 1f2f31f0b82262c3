// Library-level plumbing: error string, launch counter, device properties.
#include <stdarg.h>

#include <atomic>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace iic {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};  // process-wide: backward runs on autograd's own threads

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- runtime options: kernel-variant switches (A/B measurement, tests); defaults come from the environment -----
struct Option {
  const char* name;
  const char* env;
  int def;
  int value;
  bool init;
};
static Option g_opts[OPT_COUNT] = {
    // conv_halo: halo variant of the 3x3 / stride-1 / 64->64 fprop+dgrad: 0 never, 1 when the geometry fits and the
    // tile efficiency is good, 2 whenever the geometry fits
    {"conv_halo", "IIC_CONV_HALO", 1, 0, false},
    // conv_halo_wgrad: same switch for the halo wgrad kernel (needs conv_halo's geometry test to pass as well)
    {"conv_halo_wgrad", "IIC_CONV_HALO_WGRAD", 1, 0, false},
    // stem_quad: 4-pixels-per-thread stem conv kernel (with optional fused BN statistics): 2 = channel-interleaved thread
    // layout (whole-sector stores), 1 = 16 consecutive channels per thread, 0 = the one-pixel-per-thread kernel
    {"stem_quad", "IIC_STEM_QUAD", 2, 0, false},
    // dgrad_prefetch: the dgrad epilogues request the residual-gradient addend ahead of its use (0 = on demand)
    {"dgrad_prefetch", "IIC_DGRAD_PREFETCH", 1, 0, false},
    // tc2_mt2: N = 128 fprop/dgrad tiles of the TMA kernel take two 128-row tiles per weight k-block (0 = one)
    {"tc2_mt2", "IIC_TC2_MT2", 1, 0, false},
    // conv_halo_store: halo fprop/dgrad write their output tile with one TMA store out of a shared-memory staging
    // buffer (0 = 16-byte stores from the epilogue threads, one accumulator row per lane)
    {"conv_halo_store", "IIC_CONV_HALO_STORE", 1, 0, false},
    // stem_bwd_v2: second version of the fused stem backward's wgrad pass (cooperative neighbourhood fetch + shuffles);
    // written after the last GPU session of round 1, off until it has run on hardware
    {"stem_bwd_v2", "IIC_STEM_BWD_V2", 0, 0, false},
    // conv_halo_stats: halo fprop accumulates the BatchNorm statistics per lane over all its work items and reduces
    // across the warp once per CTA (0 = two transpose-reduces per 32-column chunk).  Validated on a B200 in round 2:
    // layer1 fprop 1073 -> 1172 TFLOP/s over the net's fprop launches, step 43.00 -> 42.42 ms (profiles/r02_session_a.md)
    {"conv_halo_stats", "IIC_CONV_HALO_STATS", 1, 0, false},
    // bn_bwd_ctas: CTAs per SM of the cooperative BatchNorm-backward kernel (2 = fastest alone; 1 = fits beside a resident
    // persistent convolution CTA, 17 KB + 198 KB of shared memory, so that it can overlap a wgrad running on a second
    // stream: _engine.OPTIONS["wgrad_stream"])
    {"bn_bwd_ctas", "IIC_BN_BWD_CTAS", 2, 0, false},
    // tf32x3_raw_hi: the 3xTF32 fprop / dgrad splitter leaves the landed fp32 stage untouched and writes only lo = x - trunc(x):
    // the tensor core ignores the 13 low mantissa bits of a kind::tf32 operand, so x itself IS the hi operand.  A third less
    // shared-memory traffic for the transform warps that bound this mode.  Validated on a B200 in round 2 (precision tests
    // and smoke green; fprop 58.9 -> 43.6 ms, dgrad 61.3 -> 49.8 ms per c4 step, profiles/r02_session_f.md); 0 = write hi too.
    {"tf32x3_raw_hi", "IIC_TF32X3_RAW_HI", 1, 0, false},
    // wgrad_mt: the bf16 im2col wgrad takes 2 or 3 of its 128-row (tap, cin) tiles per dy k-block (one accumulator buffer
    // of up to 512 TMEM columns): a third less shared-memory fill per MMA.  Validated on a B200 in round 2 (exact-integer
    // tests; wgrad 128->128 708 -> 1126, 256->256 1085 -> 1189, 512->512 897 -> 1014 TFLOP/s live in the c4 step, wgrad of
    // the step 9.70 -> 8.24 ms, profiles/r02_session_h.md); 0 = one tile per work item.
    {"wgrad_mt", "IIC_WGRAD_MT", 1, 0, false},
    // halo_addend_tma: the halo dgrad fetches its residual-gradient addend with ONE TMA load into the output staging
    // buffer and sums in place, instead of 16-byte loads from 32 different lines per warp instruction.  Validated on a
    // B200 in round 2 (exact-integer tests; the six layer-1 dgrads of the c4 step 2.12 -> 1.85 ms, profiles/r02_session_i.md);
    // 0 = per-lane loads (option dgrad_prefetch).
    {"halo_addend_tma", "IIC_HALO_ADDEND_TMA", 1, 0, false},
    // dgrad_s2_mt: the four parity-class launches of a stride-2 dgrad use the resident-weights (N = 64) / two-tile
    // (N = 128) variants of the stride-1 path (2 = also at sizes where they would not pay: tests).  Written after the last
    // GPU session: off until it has run on hardware.
    {"dgrad_s2_mt", "IIC_DGRAD_S2_MT", 0, 0, false},
};

int option(int id) {
  if (id < 0 || id >= OPT_COUNT) return 0;
  Option& o = g_opts[id];
  if (!o.init) {
    const char* e = getenv(o.env);
    o.value = (e != nullptr && e[0] != 0) ? atoi(e) : o.def;
    o.init = true;
  }
  return o.value;
}

static int option_id(const char* name) {
  if (name == nullptr) return -1;
  for (int i = 0; i < OPT_COUNT; ++i)
    if (strcmp(name, g_opts[i].name) == 0) return i;
  return -1;
}

int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace iic

extern "C" int iic_abi_version(void) { return 1; }
extern "C" const char* iic_last_error(void) { return iic::g_err; }
extern "C" long long iic_launch_count(int reset) {
  return reset ? iic::g_launches.exchange(0) : iic::g_launches.load();
}

extern "C" int iic_get_option(const char* name) {
  const int id = iic::option_id(name);
  if (id < 0) {
    iic::set_error("iic_get_option: unknown option '%s'", name ? name : "(null)");
    return IIC_ERR_BAD_ARG;
  }
  return iic::option(id);
}
extern "C" int iic_set_option(const char* name, int value) {
  const int id = iic::option_id(name);
  if (id < 0 || value < 0) {
    iic::set_error("iic_set_option: unknown option '%s' or negative value", name ? name : "(null)");
    return IIC_ERR_BAD_ARG;
  }
  iic::option(id);  // resolve the default first, so that the return value is the previous setting
  const int prev = iic::g_opts[id].value;
  iic::g_opts[id].value = value;
  return prev;
}
