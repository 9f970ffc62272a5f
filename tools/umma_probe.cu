// Bring-up probe for tcgen05 shared-memory operand formats (sm_100a).
//
// The fused attention kernels write P / dS tiles into shared memory FROM REGISTERS and reuse one
// tile under several operand roles (K-major A, MN-major A/B).  TMA is not in the loop there, so
// the canonical layouts have to be produced in software; this probe pins each (layout,
// descriptor) pair against a CPU product before any kernel depends on it.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/bin/umma_probe tools/umma_probe.cu
//   tools/bin/umma_probe            (prints max |err| per case; integers => exact expected)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

enum Mode { KMAJ_SW128 = 0, KMAJ_SW64 = 1, MN_SW128 = 2, MN_SW64 = 3 };

struct Case {
  const char* name;
  int N, K;            // M = 128 always
  int modeA, modeB;
  int a_kbyte_off;     // K-major A: extra byte offset of the first k-step inside the row
  int a_row_elems;     // K-major A: elements per smem row actually allocated (64 / 32)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// byte offset of element (row, col) in a tile of `rows` rows whose rows are `row_bytes` wide
// (128 -> SW128, 64 -> SW64); column groups wider than a row are `rows * row_bytes` apart.
__device__ __forceinline__ uint32_t tile_off(int row, int col, int rows, int row_bytes) {
  const int epr = row_bytes / 2;
  const int grp = col / epr, c = col % epr;
  const int chunk = c / 8, within = c % 8;
  const int sw = row_bytes == 128 ? (row & 7) : ((row >> 1) & 3);
  return (uint32_t)(grp * rows * row_bytes + row * row_bytes + ((chunk ^ sw) * 16) + within * 2);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

__global__ void __launch_bounds__(128) probe_kernel(const float* A, const float* B, float* D, Case c) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* gen = raw + (base - smem_u32(raw));
  const uint32_t a_off = 0, b_off = 65536, bar_off = 131072, tptr_off = 131072 + 8;
  const int M = 128, N = c.N, K = c.K;
  // ---- software tile writes ----
  for (int i = threadIdx.x; i < 65536 / 4; i += 128) {
    reinterpret_cast<uint32_t*>(gen + a_off)[i] = 0x7fc07fc0u;   // NaN poison: untouched bytes show up
    reinterpret_cast<uint32_t*>(gen + b_off)[i] = 0x7fc07fc0u;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < M * K; i += 128) {
    const int m = i / K, k = i % K;
    uint32_t o;
    if (c.modeA == KMAJ_SW128) o = tile_off(m, k + c.a_kbyte_off / 2, M, 128);
    else if (c.modeA == KMAJ_SW64) o = tile_off(m, k, M, 64);
    else if (c.modeA == MN_SW128) o = tile_off(k, m, K, 128);
    else o = tile_off(k, m, K, 64);
    *reinterpret_cast<__nv_bfloat16*>(gen + a_off + o) = __float2bfloat16_rn(A[i]);
  }
  for (int i = threadIdx.x; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    uint32_t o;
    if (c.modeB == KMAJ_SW128) o = tile_off(n, k, N, 128);
    else if (c.modeB == KMAJ_SW64) o = tile_off(n, k, N, 64);
    else if (c.modeB == MN_SW128) o = tile_off(k, n, K, 128);
    else o = tile_off(k, n, K, 64);
    *reinterpret_cast<__nv_bfloat16*>(gen + b_off + o) = __float2bfloat16_rn(B[i]);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  const uint32_t bar = base + bar_off;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(base + tptr_off) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + tptr_off);
  if (threadIdx.x == 0) {
    const uint32_t a_mn = (c.modeA >= MN_SW128), b_mn = (c.modeB >= MN_SW128);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (a_mn << 15) | (b_mn << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int s = 0; s < K / 16; ++s) {
      uint64_t da, db;
      auto mk = [&](int mode, uint32_t t0, int mn_extent, int kbyte_off) -> uint64_t {
        if (mode == KMAJ_SW128)   // atoms of 64 k: (s / 4) * rows * 128, +32 B per k-step inside
          return make_desc(t0 + (s / 4) * mn_extent * 128 + (s % 4) * 32 + kbyte_off, 16, 1024, 2);
        if (mode == KMAJ_SW64)
          return make_desc(t0 + (s / 2) * mn_extent * 64 + (s % 2) * 32, 16, 512, 4);
        if (mode == MN_SW128)     // LBO = stride between 64-wide MN groups = K rows * 128 B
          return make_desc(t0 + s * 2048, (uint32_t)K * 128, 1024, 2);
        return make_desc(t0 + s * 1024, (uint32_t)K * 64, 512, 4);
      };
      da = mk(c.modeA, base + a_off, M, c.a_kbyte_off);
      db = mk(c.modeB, base + b_off, N, 0);
      const uint32_t acc = s > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  }
  // wait (parity 0)
  asm volatile(
      "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t"
      "@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar)
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]),
                   "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

int main() {
  const Case cases[] = {
      {"A Kmaj SW128 K=128 | B Kmaj SW128 N=64        (S = Q K^T)", 64, 128, KMAJ_SW128, KMAJ_SW128, 0, 64},
      {"A Kmaj SW128 K=128 | B Kmaj SW128 N=32        (S^T, 32-wide chunk)", 32, 128, KMAJ_SW128, KMAJ_SW128, 0, 64},
      {"A Kmaj SW128 K=64  | B MN   SW128 N=96        (O = P V, V as [key][d])", 96, 64, KMAJ_SW128, MN_SW128, 0, 64},
      {"A Kmaj SW128 K=32 at +64B | B MN SW128 N=96   (odd half of a 64-wide tile)", 96, 32, KMAJ_SW128, MN_SW128, 64, 64},
      {"A Kmaj SW64  K=32  | B MN   SW128 N=96        (dV += Pd^T dO, 32-q chunk)", 96, 32, KMAJ_SW64, MN_SW128, 0, 32},
      {"A MN   SW128 K=128 | B MN   SW128 N=64        (dQ^T = K^T dS^T)", 64, 128, MN_SW128, MN_SW128, 0, 64},
      {"A MN   SW128 K=128 | B MN   SW64  N=32        (dQ^T, 32-wide B)", 32, 128, MN_SW128, MN_SW64, 0, 64},
      {"A MN   SW64  K=64  | B Kmaj SW64  N=64        (MN SW64 A, 4 groups)", 64, 64, MN_SW64, KMAJ_SW64, 0, 64},
  };
  const int M = 128;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, M * 128 * 4);
  cudaMalloc(&dB, 128 * 128 * 4);
  cudaMalloc(&dD, M * 128 * 4);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  int bad = 0;
  for (const Case& c : cases) {
    std::vector<float> A(M * c.K), B(c.N * c.K), D(M * c.N), R(M * c.N);
    srand(7);
    for (auto& x : A) x = (float)(rand() % 9 - 4);
    for (auto& x : B) x = (float)(rand() % 7 - 3);
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < c.N; ++n) {
        float s = 0;
        for (int k = 0; k < c.K; ++k) s += A[m * c.K + k] * B[n * c.K + k];
        R[m * c.N + n] = s;
      }
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, M * c.N * 4);
    probe_kernel<<<1, 128, 140 * 1024>>>(dA, dB, dD, c);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%-75s CUDA error %s\n", c.name, cudaGetErrorString(e));
      return 2;
    }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    float worst = 0;
    int nbad = 0;
    for (size_t i = 0; i < D.size(); ++i) {
      const float d = D[i] - R[i];
      if (!(d == 0.f)) ++nbad;
      if (d == d && (d > worst || -d > worst)) worst = d > 0 ? d : -d;
    }
    printf("%-75s %s  (mismatches %d / %zu, max |err| %g)\n", c.name, nbad ? "FAIL" : "ok", nbad,
           D.size(), worst);
    bad += nbad != 0;
  }
  return bad ? 1 : 0;
}
