// Library-level entry points: version, error string, device query.
#include "ssb_common.cuh"
#include <string.h>
#include <stdlib.h>
#include <atomic>

namespace ssb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  // clear the sticky-free error state so the next call starts clean
  (void)cudaGetLastError();
  return (int)e;
}

static std::atomic<const uint64_t*> g_seed_src{nullptr};
const uint64_t* seed_source() { return g_seed_src.load(std::memory_order_relaxed); }
void set_seed_source(const uint64_t* p) { g_seed_src.store(p, std::memory_order_relaxed); }

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("SSB_PDL");
    return e && e[0] == '1';   // off by default: see ssb_common.cuh
  }();
  return on;
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached_dev = dev;
    cached = n;
  }
  return cached;
}

}  // namespace ssb

extern "C" {

int ssb_version(void) { return SSB_ABI_VERSION; }

int64_t ssb_sizeof(int which) {
  switch (which) {
    case 0: return (int64_t)sizeof(ssb_gather_t);
    case 1: return (int64_t)sizeof(ssb_scatter_t);
    case 2: return (int64_t)sizeof(ssb_epilogue_t);
    case 3: return (int64_t)sizeof(ssb_tc_operand_t);
    case 4: return (int64_t)sizeof(ssb_dtw_pair_t);
    case 5: return (int64_t)sizeof(ssb_utt_t);
    case 6: return (int64_t)sizeof(ssb_prep_entry_t);
    case 7: return (int64_t)sizeof(ssb_emg_rec_t);
    default: return -1;
  }
}

const char* ssb_last_error(void) { return ssb::g_err; }

int ssb_set_seed_source(const uint64_t* dev_seed_offset) {
  ssb::set_seed_source(dev_seed_offset);
  return SSB_OK;
}

int ssb_device_sm_count(void) {
  int n = ssb::num_sms();
  if (n <= 0) {
    ssb::set_error("no CUDA device available");
    return SSB_ERR_UNSUPPORTED;
  }
  return n;
}

}  // extern "C"
