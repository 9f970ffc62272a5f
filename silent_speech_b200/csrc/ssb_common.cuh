// Shared device/host helpers for the libssb kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/ssb.h"

namespace ssb {

// ---- error plumbing (thread-local message, C return codes) -----------------
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

#define SSB_REQUIRE(cond, ...)                      \
  do {                                              \
    if (!(cond)) {                                  \
      ::ssb::set_error(__VA_ARGS__);                \
      return SSB_ERR_ARG;                           \
    }                                               \
  } while (0)

#define SSB_CUDA(call)                                              \
  do {                                                              \
    cudaError_t e__ = (call);                                       \
    if (e__ != cudaSuccess) return ::ssb::cuda_fail(e__, #call);    \
  } while (0)

#define SSB_LAUNCH_CHECK(name)                                      \
  do {                                                              \
    cudaError_t e__ = cudaGetLastError();                           \
    if (e__ != cudaSuccess) return ::ssb::cuda_fail(e__, name);     \
  } while (0)

int num_sms();

// Device-resident dropout seed offset (ssb_set_seed_source): when set, every dropout site keys
// Philox with seed + *src, so a captured CUDA graph draws fresh masks on every replay.
const uint64_t* seed_source();

// Programmatic dependent launch (opt-in, SSB_PDL=1): kernels launched with the attribute start
// their prologue under the previous kernel's tail and call pdl_wait() before touching its results.
// Measured on the cfg-1 step graph with the attribute on every gemm_tc launch (r2 session 8,
// two A/B pairs): 24.79 / 24.82 ms with, 24.50 / 24.49 ms without -- the 1-CTA-per-SM persistent
// kernels gain nothing from an early launch and lose to the dependency wait, so it stays off.
bool pdl_enabled();

// ---- small device helpers ---------------------------------------------------
// Blocks until the grids this launch depends on have completed and their writes are visible; a
// no-op for a kernel launched without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// One elected lane of a CONVERGED warp.  Unlike `lane == 0`, ptxas knows that exactly one thread
// passes: the tcgen05 / TMA (uniform-datapath) instructions behind it are emitted straight, without
// the ELECT + BRA.U.ANY "for each active thread" loop it wraps around them under a divergent
// predicate (8 extra instructions per MMA on the issuing thread).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_4(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst), "l"(src),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Philox4x32 counter-based RNG (Salmon et al., SC'11). One call = 4 x 32 random bits.
// Used for every dropout site so forward and backward regenerate identical masks from
// (seed, site offset, element index) without storing them.
// Rounds: 10, the paper's (and cuRAND's) default.  7 rounds (the smallest count the paper reports
// as Crush-resistant) was tried in r2: 6 % fewer instructions in the fused attention forward,
// 3.7 % / 1.6 % off the forward / backward kernel times, nothing measurable on the step -- not
// worth leaving the standard generator.
constexpr int PHILOX_ROUNDS = 10;
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int i = 0; i < PHILOX_ROUNDS; ++i) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0;
    key.y += W1;
  }
  return ctr;
}

__device__ __forceinline__ uint64_t eff_seed(uint64_t seed, const uint64_t* src) {
  return src ? seed + __ldg(src) : seed;
}

// keep-mask for 4 consecutive elements starting at element index 4*idx4 of dropout site `site`.
// Element e is kept iff its 32 random bits >= thresh, thresh = round(p * 2^32).
__device__ __forceinline__ uint4 dropout_bits4(uint64_t seed, uint32_t site, uint64_t idx4) {
  uint4 ctr = make_uint4((uint32_t)idx4, (uint32_t)(idx4 >> 32), site, 0x5353425Fu);
  uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  return philox4x32(ctr, key);
}

}  // namespace ssb
