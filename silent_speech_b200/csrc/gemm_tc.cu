// tcgen05 / TMEM / TMA GEMM for sm_100a with fp32-class accuracy ("bf16x3").
//
// Why bf16x3.  The north star demands 1e-3 relative parity with an fp32 reference.  One-pass
// TF32 lands at ~8e-4 after 6 layers — no margin; plain bf16 fails (6e-3).
// Every fp32 operand is therefore stored as TWO bf16 planes, x = hi + lo (hi = bf16(x),
// lo = bf16(x - hi): 16 mantissa bits), and each logical MMA is issued as three tcgen05.mma
//      D += A_hi*B_hi + A_hi*B_lo + A_lo*B_hi          (fp32 accumulation in TMEM)
// dropping only the lo*lo term (2^-16 relative).  Measured error ~1e-6, at a tensor-pipe
// cost of 3 bf16 MMAs = 1.5x a TF32 MMA.  Split planes take exactly the bytes of the fp32
// tensor they replace (2 + 2 B per element).
//
// Kernel shape (persistent, 192 threads, warp-specialised; two TMEM accumulators so the epilogue
// of a tile overlaps the next tile's MMAs).  Two instantiations:
//   CTAS = 2  (default) a CTA PAIR per TPC walks 256 x 256 output tiles with
//             tcgen05.mma.cta_group::2 (M256 N256 K16): each CTA stages its own 128 rows of A
//             and HALF of B (128 rows), so the shared-memory traffic per MMA drops from
//             12 KB read + 12 KB.. to 8 KB read per SM -- the 1-CTA form is shared-memory-
//             bandwidth bound at ~61 % tensor-pipe activity (profiles/r1_gemm_tc_persistent.txt).
//             5 stages x 32 KB per CTA; the leader CTA (rank 0) issues every MMA, its mbarriers
//             collect the TMA bytes of both CTAs, commits are multicast to both.
//   CTAS = 1  one CTA per SM, 128 x 256 tiles, 3 stages x 48 KB (small / odd-shaped problems).
//   warp 0    TMA producer: per 32-deep k-block, the A_hi/A_lo and B_hi/B_lo tiles land in
//             swizzled shared memory, completion on an mbarrier (complete_tx).
//   warp 1    TMEM allocator (2 x 256 fp32 columns) and single-thread MMA issuer: 2 x 3
//             tcgen05.mma per k-block, tcgen05.commit releases the stage.
//   warps 2-9 epilogue (two per TMEM lane group, on even / odd 32-column blocks): tcgen05.ld
//             32 lanes x 32 columns at a time, bias / ReLU / dropout / mask / accumulate /
//             split-K reduction, fp32 stores through the same (batch, row) -> address map as
//             the CUDA-core engine (gemm_simt.cu).  With K = 768 a 256 x 256 tile is only 9.4 us
//             of MMAs: four warps could not drain 128 KB per CTA (Philox dropout, mask reads)
//             in that time, so FFN1 / its data gradient were epilogue-bound.
//
// Operand forms (all expressed by the TMA tensor maps built on the host):
//   K-major  C[m,n] = sum_k A[m,k] B[n,k]      forward and data gradients.  A may be an im2col
//            view of a channels-last activation: tensor map dims (C, L, batch, plane), box
//            (64, 128 rows x stride), start row t0*s + tap - 1; negative / past-the-end rows
//            are zero-filled by TMA itself = the convolution's padding (architecture.py:18-24).
//   MN-major C[f,n] = sum_m X[m,f] G[m,n]      weight gradients: both operands are read
//            "transposed" straight from their row-major planes (UMMA MN-major descriptors),
//            reduction over tokens split across gridDim.z with fp32 atomics.
#include "ssb_common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <mutex>

namespace {

// 32-deep k-blocks in a 4-stage ring: a stage (48 KB) is consumed in ~0.4 us, so three stages of
// prefetch cover ~1.2 us of DRAM latency (the first version, 2 x 96 KB with 64-deep blocks,
// stalled on every HBM-cold operand: 340 us in-step vs 185 us L2-warm for the same GEMM).
constexpr int BM = 128, BN = 256, BK = 32;
constexpr int A_PLANE = BM * BK * 2;   // 8 KB
constexpr int MN_GROUP = BK * 128;     // bytes of one 64-element MN group of a stage (MN-major)
constexpr int TMEM_COLS = 512;   // two 256-column fp32 accumulators
constexpr int EPI_WARPS = 8;        // two per TMEM lane group: even / odd 32-column blocks
constexpr int THREADS = 64 + 32 * EPI_WARPS;
constexpr int EPI_ROW_BYTES = 36 * 4;                 // 32 fp32 + 4 pad: conflict-free both ways
constexpr int EPI_WARP_BYTES = 32 * EPI_ROW_BYTES;    // one epilogue warp's transpose tile
constexpr int EPI_BYTES = EPI_WARPS * EPI_WARP_BYTES;  // 36 KB

template <int CTAS>
struct Cfg {
  static constexpr int BMC = BM * CTAS;                    // output rows per tile (whole pair)
  static constexpr int B_ROWS = BN / CTAS;                 // rows of B each CTA stages
  static constexpr int B_PLANE = B_ROWS * BK * 2;          // 16 KB / 8 KB
  static constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;   // 48 KB / 32 KB per CTA
  static constexpr int STAGES = CTAS == 2 ? 5 : 3;
  static constexpr int SMEM_BYTES =
      STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + EPI_BYTES;
};

struct TcParams {
  // operand addressing (see header comment)
  int a_inner, a_row_step, a_tap_step, a_off;   // A: k-block -> (d0, tap); d1 = row*step + tap*tap_step + off
  int rows_per_batch;                            // output rows per batch item (K-major) / reduction rows (MN)
  int tiles_per_batch;                           // K-major: M tiles per batch item
  int chunks_per_batch;                          // MN-major: 64-row reduction chunks per batch item
  int num_kb;                                    // K-major: k-blocks; MN-major: total chunks (batches * chunks_per_batch)
  int kb_per_split;
  int M_valid_total;                             // rows of C (K-major: batches*rows_per_batch; MN: features)
  int N;
  int mn_major;
  int mn_batched;                                // MN-major: one output per batch item (z = batch)
  int batch_div;                                 // batch index -> (lo = b % div, hi = b / div)
  int b_mode;                                    // B batch coords: 0 none, 1 (lo, hi), 2 (lo, 0)
  int tiles_x, tiles_y, splits;                   // persistent tile list
  // stream-K remainder (K-major only; see the schedule comment in the kernel): tiles [0, sk_full) are
  // walked whole, the sk_rem tiles of the last, partial wave are cut along K into pieces shared by
  // sk_workers workers.  sk_rem == 0: classic schedule.
  int sk_allow;                                   // host: this launch may use the stream-K remainder
  int sk_full, sk_rem, sk_workers;
  float* sk_ws;                                   // [worker][rank][8 warps][4 blocks][32][32] fp32 partials
  unsigned* sk_flags;                             // [worker][rank][8 warps], 0 between launches
  int n_mma;                                      // K-major, N < BN: the MMA's N (multiple of 32); the B box
                                                  // holds n_mma / CTAS rows per CTA.  0: BN
  // epilogue
  float* out; int64_t out_batch_stride, out_batch_stride_hi; int out_ld, out_dt, out_doff;
  const float* bias; const float* mask_src; float mask_scale;
  __nv_bfloat16* planes; int64_t planes_stride;   // optional split-plane copy of the result, plain (M, N)
  const __nv_bfloat16* mask_planes;               // mask source given as its bf16 hi plane
  const uint8_t* mask_bits;                       // mask source given as 1 bit per element ((M, N / 8) bytes)
  uint8_t* mask_bits_out;                         // forward: bit (m, n) = result(m, n) > 0
  int relu, accumulate, atomic;
  int planes_lrelu; float planes_slope;           // plane copy = leaky_relu(result) (fp32 output unchanged)
  int grp_w; int64_t grp_stride;                  // weight gradients: output column n -> (n / grp_w) * grp_stride + n % grp_w
  float drop_p, drop_scale; uint32_t drop_thresh; uint64_t seed; const uint64_t* seed_src; uint32_t site;
};

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
// 5-D tensor maps: (inner, row, batch_lo, batch_hi, plane).  CTAS == 2: the destination is this
// CTA's shared memory, `bar` is the LEADER CTA's barrier as a shared::cluster address.
template <int CTAS>
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3, int c4) {
  if constexpr (CTAS == 1) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
  }
}
template <int CTAS>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  if constexpr (CTAS == 1) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// CTAS == 2: the arrival is multicast to the barrier at the same offset in BOTH CTAs
template <int CTAS>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if constexpr (CTAS == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     bar)
                 : "memory");
  } else {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;" ::"r"(bar),
        "h"((uint16_t)3)
        : "memory");
  }
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// shared-memory matrix descriptor (cute/arch/mma_sm100_desc.hpp: SmemDescriptor);
// layout: 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);            // start address   [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;       // leading offset  [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;       // stride offset   [32,46)
  d |= (uint64_t)1 << 46;                                 // version = 1 (Blackwell)
  d |= (uint64_t)layout << 61;                            // layout type
  return d;
}

// instruction descriptor (InstrDescriptor): D=f32, A=B=bf16, M=128 (256 for a CTA pair), N=256
__host__ __device__ constexpr uint32_t make_idesc(int mn_major, int m, int n = BN) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(mn_major & 1) << 15) |
         ((uint32_t)(mn_major & 1) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// Persistent: gridDim.x / CTAS workers (a CTA, or a CTA pair on one TPC) walk the tile list; the
// two 256-column TMEM accumulators alternate so that the epilogue of tile i runs under the MMAs
// of tile i+1, and the TMA ring keeps streaming across tile boundaries.
template <int CTAS>
__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const TcParams p) {
  using C = Cfg<CTAS>;
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, B_PLANE = C::B_PLANE;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ssb::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem_base + STAGES * STAGE_BYTES;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * STAGES;
  const uint32_t tfull_bar = bars + 16 * STAGES, tempty_bar = tfull_bar + 16;
  const uint32_t tmem_ptr_smem = tempty_bar + 16;
  const uint32_t epi_base = bars + 256;   // 4.5 KB transpose tile per epilogue warp
  volatile uint32_t* tmem_ptr_gen =
      reinterpret_cast<volatile uint32_t*>(smem_raw + (tmem_ptr_smem - ssb::smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CTAS == 2 ? (int)cluster_ctarank() : 0;   // 0 = leader (issues the MMAs)
  const int worker = (int)blockIdx.x / CTAS, num_workers = (int)gridDim.x / CTAS;
  const int tiles_xy = p.tiles_x * p.tiles_y;
  const int total_tiles = tiles_xy * p.splits;
  // barriers other CTAs signal are addressed in the leader's shared memory
  const uint32_t full_bar_lead = CTAS == 2 ? map_to_cta(full_bar, 0) : full_bar;
  const uint32_t tempty_bar_lead = CTAS == 2 ? map_to_cta(tempty_bar, 0) : tempty_bar;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);     // the leader's arrive.expect_tx (+ TMA bytes of all CTAS)
      mbar_init(empty_bar + 8 * s, 1);    // one (multicast) tcgen05.commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar + 8 * a, 1);
      mbar_init(tempty_bar + 8 * a, EPI_WARPS * CTAS);   // one arrival per epilogue warp of every CTA
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CTAS == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       tmem_ptr_smem),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {   // the same warp of both CTAs, same destination offset
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       tmem_ptr_smem),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (CTAS == 2) cluster_sync_all();   // the peer's barriers are initialised too
  // programmatic dependent launch: everything above (barrier init, TMEM allocation, descriptor
  // prefetch) may run under the previous kernel's tail; its results are read only from here on
  ssb::pdl_wait();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_gen;

  // tile -> coordinates (n fastest, so concurrently running workers share A rows and sweep B);
  // row0 / f0 are THIS CTA's 128 rows of the (128 * CTAS)-row tile, nb0 its rows of B
  auto decode = [&](int tile, int& n0, int& batch, int& row0, int& f0, int& kb_begin, int& nkb) {
    const int z = tile / tiles_xy;
    const int rem = tile - z * tiles_xy;
    const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
    n0 = tx * BN;
    batch = 0; row0 = 0; f0 = 0;
    if (!p.mn_major) {
      batch = ty / p.tiles_per_batch;
      row0 = (ty - batch * p.tiles_per_batch) * C::BMC + rank * BM;
      kb_begin = 0;
      nkb = p.num_kb;
    } else {
      f0 = ty * C::BMC + rank * BM;
      kb_begin = z * p.kb_per_split;
      nkb = max(min(p.num_kb, kb_begin + p.kb_per_split) - kb_begin, 0);
    }
  };

  // ---- work items -------------------------------------------------------------------------
  // Classic schedule: worker w walks tiles w, w + W, ... whole.  With T tiles on W workers the last
  // wave holds only T mod W tiles while the other workers idle: 189 tiles (16000 x 768 outputs) on
  // 74 CTA pairs are 2.55 waves of work in 3 waves of time.  Stream-K remainder: the F = T - T mod W
  // tiles of the full waves stay whole; the R = T mod W remaining tiles are laid end to end along K
  // (R * num_kb k-blocks) and cut into equal contiguous ranges, one per worker.  A range covers the
  // END of one tile and / or the START of the next.  The worker whose range holds a tile's last
  // k-block OWNS the tile: its epilogue adds the other ranges' fp32 partial accumulators (written to
  // a workspace slot per worker and published through a flag) and then runs the normal epilogue.
  // Every worker first computes its partial piece (no waiting), then its owner piece, so an owner
  // only ever waits for work that is being computed at the same time: no chains, and all CTAs of
  // the persistent grid are resident.  At most 4-5 workers share a tile (sk_workers <= 4 R).
  constexpr int IT_FULL = 0, IT_PARTIAL = 1, IT_OWNER = 2;
  auto sk_range = [&](int w, int& u0, int& u1) {
    const long long U = (long long)p.sk_rem * p.num_kb;
    u0 = (int)((long long)w * U / p.sk_workers);
    u1 = (int)((long long)(w + 1) * U / p.sk_workers);
  };
  // idx-th work item of this worker: tile, k-block range, role.  false: no more items.
  auto get_item = [&](int idx, int& tile, int& kb0, int& nk, int& mode) -> bool {
    const int full = p.sk_rem ? p.sk_full : total_tiles;
    const int n1 = full > worker ? (full - worker + num_workers - 1) / num_workers : 0;
    kb0 = -1;            // -1: take the range decode() gives (whole K, or the split of an MN-major tile)
    nk = 0;
    mode = IT_FULL;
    if (idx < n1) { tile = worker + idx * num_workers; return true; }
    if (!p.sk_rem || worker >= p.sk_workers) return false;
    int u0, u1;
    sk_range(worker, u0, u1);
    if (u0 == u1) return false;
    const int k = idx - n1;
    const int ta = u0 / p.num_kb;
    const int a_last = (ta + 1) * p.num_kb;             // one past the last unit of tile ta
    const bool has_b = u1 > a_last;
    if (has_b && k == 0) {                              // head of the next tile: partial, computed first
      tile = p.sk_full + ta + 1; kb0 = 0; nk = u1 - a_last; mode = IT_PARTIAL;
      return true;
    }
    if (k != (has_b ? 1 : 0)) return false;
    const int a_end = has_b ? a_last : u1;
    tile = p.sk_full + ta; kb0 = u0 - ta * p.num_kb; nk = a_end - u0;
    mode = a_end == a_last ? (kb0 == 0 ? IT_FULL : IT_OWNER) : IT_PARTIAL;
    return true;
  };

  if (warp == 0) {
    // ===================== TMA producer (every CTA stages its own operand rows) =====================
    if (ssb::elect_one()) {
      uint32_t it = 0;
      int tile, it_kb0, it_nk, it_mode;
      for (int idx = 0; get_item(idx, tile, it_kb0, it_nk, it_mode); ++idx) {
        int n0, batch, row0, f0, kb_begin, nkb;
        decode(tile, n0, batch, row0, f0, kb_begin, nkb);
        if (it_kb0 >= 0) { kb_begin = it_kb0; nkb = it_nk; }
        const int b_rows = p.n_mma ? p.n_mma / CTAS : C::B_ROWS;   // rows of B this CTA stages
        const int nb0 = n0 + rank * b_rows;
        const uint32_t stage_tx = p.n_mma ? (uint32_t)(2 * A_PLANE + 2 * b_rows * BK * 2) : (uint32_t)STAGE_BYTES;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int kb = kb_begin + i;
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(empty_bar + 8 * s, ph ^ 1u);
          const uint32_t st = smem_base + s * STAGE_BYTES;
          if (rank == 0) mbar_expect_tx(full_bar + 8 * s, CTAS * stage_tx);
          const uint32_t fb = full_bar_lead + 8 * s;
          if (!p.mn_major) {
            const int kk = kb * BK;
            const int tap = kk / p.a_inner, c0 = kk - tap * p.a_inner;
            const int d1 = row0 * p.a_row_step + tap * p.a_tap_step + p.a_off;
            const int lo = batch % p.batch_div, hi = batch / p.batch_div;
            const int blo = p.b_mode ? lo : 0, bhi = p.b_mode == 1 ? hi : 0;
            tma_load_5d<CTAS>(st, &mapA, fb, c0, d1, lo, hi, 0);
            tma_load_5d<CTAS>(st + A_PLANE, &mapA, fb, c0, d1, lo, hi, 1);
            tma_load_5d<CTAS>(st + 2 * A_PLANE, &mapB, fb, kk, nb0, blo, bhi, 0);
            tma_load_5d<CTAS>(st + 2 * A_PLANE + B_PLANE, &mapB, fb, kk, nb0, blo, bhi, 1);
          } else {
            const int b = kb / p.chunks_per_batch;
            const int t0 = (kb - b * p.chunks_per_batch) * BK;
            const int tap = f0 / p.a_inner, c0 = f0 - tap * p.a_inner;
            const int d1 = t0 * p.a_row_step + tap * p.a_tap_step + p.a_off;
            const int lo = b % p.batch_div, hi = b / p.batch_div;
#pragma unroll
            for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
              for (int h = 0; h < BM / 64; ++h)
                tma_load_5d<CTAS>(st + pl * A_PLANE + h * MN_GROUP, &mapA, fb, c0 + 64 * h, d1, lo,
                                  hi, pl);
#pragma unroll
              for (int h = 0; h < C::B_ROWS / 64; ++h)
                tma_load_5d<CTAS>(st + 2 * A_PLANE + pl * B_PLANE + h * MN_GROUP, &mapB, fb,
                                  nb0 + 64 * h, t0, lo, hi, pl);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && ssb::elect_one()) {
      const uint32_t idesc = make_idesc(p.mn_major, C::BMC, p.n_mma ? p.n_mma : BN);
      uint32_t it = 0, tcount = 0;
      int tile, it_kb0, it_nk, it_mode;
      for (int idx = 0; get_item(idx, tile, it_kb0, it_nk, it_mode); ++idx, ++tcount) {
        int n0, batch, row0, f0, kb_begin, nkb;
        decode(tile, n0, batch, row0, f0, kb_begin, nkb);
        if (it_kb0 >= 0) { kb_begin = it_kb0; nkb = it_nk; }
        const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
        mbar_wait(tempty_bar + 8 * acc, aph ^ 1u);      // every epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + acc * (uint32_t)BN;
        for (int i = 0; i < nkb; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1u;
          mbar_wait(full_bar + 8 * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_base + s * STAGE_BYTES;
          const uint32_t a_hi = st, a_lo = st + A_PLANE, b_hi = st + 2 * A_PLANE,
                         b_lo = st + 2 * A_PLANE + B_PLANE;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            uint64_t dah, dal, dbh, dbl;
            if (!p.mn_major) {
              // K-major, SW64: rows of 64 B (32 bf16), 8-row groups 512 B apart; +32 B per K=16
              const uint32_t off = k * 32;
              dah = make_desc(a_hi + off, 16, 512, 4);
              dal = make_desc(a_lo + off, 16, 512, 4);
              dbh = make_desc(b_hi + off, 16, 512, 4);
              dbl = make_desc(b_lo + off, 16, 512, 4);
            } else {
              // MN-major, SW128: 64-element MN groups MN_GROUP (4 KB) apart (LBO), 8-row K groups
              // 1 KB apart (SBO); +2 KB per K=16 step
              const uint32_t off = k * 2048;
              dah = make_desc(a_hi + off, MN_GROUP, 1024, 2);
              dal = make_desc(a_lo + off, MN_GROUP, 1024, 2);
              dbh = make_desc(b_hi + off, MN_GROUP, 1024, 2);
              dbl = make_desc(b_lo + off, MN_GROUP, 1024, 2);
            }
            const uint32_t acc0 = (i > 0 || k > 0) ? 1u : 0u;
            umma_bf16<CTAS>(tmem_d, dah, dbh, idesc, acc0);
            umma_bf16<CTAS>(tmem_d, dah, dbl, idesc, 1u);
            umma_bf16<CTAS>(tmem_d, dal, dbh, idesc, 1u);
          }
          umma_commit<CTAS>(empty_bar + 8 * s);   // frees this smem stage (in every CTA) when the MMAs retire
        }
        umma_commit<CTAS>(tfull_bar + 8 * acc);   // accumulator complete (in every CTA)
      }
    }
  } else {
    // ===================== epilogue (warps 2..5 of every CTA: its own 128 rows) =====================
    // tcgen05.ld hands every lane one ROW (32 consecutive columns).  Storing from that layout
    // makes each warp store touch 32 rows x 16 B (32 half-used sectors: the 1-CTA profile showed
    // l1tex at 75 %), so the 32 x 32 block is transposed through a padded shared-memory tile
    // first: afterwards a warp instruction covers 4 rows x 128 contiguous bytes.
    const int lg = warp & 3;              // TMEM lane group this warp may read
    const int ew = warp - 2;              // 0..7: warps 2-5 take even column blocks, 6-9 odd ones
    const int cpar = ew >> 2;
    const uint64_t seed = ssb::eff_seed(p.seed, p.seed_src);
    const uint32_t stg = epi_base + (uint32_t)ew * EPI_WARP_BYTES;
    const int rsub = lane >> 3;           // after the transpose: row (i*4 + rsub) of the 32-row group,
    const int csub = (lane & 7) * 4;      // columns csub .. csub+3 of the 32-column block
    // "wide" lane map for split-plane epilogue operands: row (i*8 + rsubw), 8 columns per lane, so
    // a lane's share of a bf16 plane row is one 16 B store (the 4-column map made 8 B stores and
    // the plane-emitting FFN GEMMs ran 25 % slower than their fp32-output twins)
    const bool wide = p.planes != nullptr || p.mask_planes != nullptr || p.mask_bits != nullptr;
    const int rsubw = lane >> 2, csubw = (lane & 3) * 8;
    uint32_t tcount = 0;
    int tile, it_kb0, it_nk, it_mode;
    for (int idx = 0; get_item(idx, tile, it_kb0, it_nk, it_mode); ++idx, ++tcount) {
      int n0, batch, row0, f0, kb_begin, nkb;
      decode(tile, n0, batch, row0, f0, kb_begin, nkb);
      if (it_kb0 >= 0) { kb_begin = it_kb0; nkb = it_nk; }
      const uint32_t acc = tcount & 1u, aph = (tcount >> 1) & 1u;
      // stream-K roles: the workers whose partial accumulators this (owner) item has to add
      int sk_contrib[6], sk_nc = 0;
      if (it_mode == IT_OWNER) {
        const int t_first = (tile - p.sk_full) * p.num_kb;      // first unit of this tile
        for (int w = worker - 1; w >= 0 && sk_nc < 6; --w) {
          int u0, u1;
          sk_range(w, u0, u1);
          if (u1 <= t_first) break;
          if (u0 != u1) sk_contrib[sk_nc++] = w;
        }
      }
      // this warp's slot of a worker's partial tile / its flag
      auto sk_slot = [&](int w) -> float* {
        return p.sk_ws + ((size_t)(w * CTAS + rank) * EPI_WARPS + ew) * (4 * 32 * 32);
      };
      auto sk_flag = [&](int w) -> unsigned* { return p.sk_flags + (w * CTAS + rank) * EPI_WARPS + ew; };
      // first row this lane stores (rows advance by 4 per i), its address and global row index
      int first, limit;
      float* orow0;
      int64_t grow0;   // global output row index (for mask / dropout addressing)
      if (!p.mn_major) {
        first = row0 + lg * 32 + rsub;
        limit = p.rows_per_batch;
        grow0 = (int64_t)batch * p.rows_per_batch + first;
        orow0 = p.out + (int64_t)(batch % p.batch_div) * p.out_batch_stride +
                (int64_t)(batch / p.batch_div) * p.out_batch_stride_hi +
                ((int64_t)first * p.out_dt + p.out_doff) * p.out_ld;
      } else {
        first = f0 + lg * 32 + rsub;
        limit = p.M_valid_total;
        grow0 = first;
        orow0 = p.out + ((int64_t)first * p.out_dt + p.out_doff) * p.out_ld;
        if (p.mn_batched) {   // z = batch item: every batch item has its own output block
          const int bz = tile / tiles_xy;
          grow0 += (int64_t)bz * p.M_valid_total;
          orow0 += (int64_t)(bz % p.batch_div) * p.out_batch_stride +
                   (int64_t)(bz / p.batch_div) * p.out_batch_stride_hi;
        }
      }
      if (nkb == 0) limit = 0;
      const int64_t row_step = (int64_t)4 * p.out_dt * p.out_ld;
      // grouped output columns (a fused QKV weight gradient lands in the (H, D, dh) parameters)
      auto ncol = [&](int n) -> int64_t {
        return p.grp_w ? (int64_t)(n / p.grp_w) * p.grp_stride + (n % p.grp_w) : (int64_t)n;
      };
      const uint32_t tmem_d = tmem_base + acc * (uint32_t)BN + ((uint32_t)(lg * 32) << 16);
      const bool group_live = first - rsub < limit;   // warp-uniform: not a pure padding row group
      // Operands the epilogue READS (ReLU/dropout mask source, accumulate target) are fetched for
      // the whole 32 x 32 block before the accumulator is touched: eight independent 16 B loads
      // per lane in flight instead of one dependent HBM round trip per output row (the masked
      // FFN data gradient ran at 548 us vs 265 us for the same-shape forward GEMM).
      float4 side[8];
      // wide map (K-major plain output only): first row, global row index and fp32 row pointer
      const int firstw = row0 + lg * 32 + rsubw;
      const int64_t groww0 = (int64_t)batch * p.rows_per_batch + firstw;
      float* const oroww0 = p.out + (int64_t)(batch % p.batch_div) * p.out_batch_stride +
                            (int64_t)(batch / p.batch_div) * p.out_batch_stride_hi +
                            ((int64_t)firstw * p.out_dt + p.out_doff) * p.out_ld;
      const int64_t row_stepw = (int64_t)8 * p.out_dt * p.out_ld;
      auto prefetch = [&](int c) {
        if (!(p.mask_src || p.mask_planes || p.mask_bits || p.accumulate) || !group_live) return;
        if (wide) {
          const int n = n0 + c * 32 + csubw;
          if (n >= p.N) return;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (firstw + 8 * i >= limit) continue;
            const int64_t e = (groww0 + 8 * i) * p.N + n;
            if (p.mask_bits) {     // one byte decides the lane's 8 outputs of this row
              const uint32_t m = __ldg(p.mask_bits + (e >> 3));
              side[2 * i] = make_float4((m & 1u) ? 1.f : 0.f, (m & 2u) ? 1.f : 0.f, (m & 4u) ? 1.f : 0.f,
                                        (m & 8u) ? 1.f : 0.f);
              side[2 * i + 1] = make_float4((m & 16u) ? 1.f : 0.f, (m & 32u) ? 1.f : 0.f,
                                            (m & 64u) ? 1.f : 0.f, (m & 128u) ? 1.f : 0.f);
            } else if (p.mask_planes) {   // 8 bf16: only sign / zero-ness matter
              const uint4 m = __ldg(reinterpret_cast<const uint4*>(p.mask_planes + e));
              side[2 * i] = make_float4(__uint_as_float(m.x << 16), __uint_as_float(m.x & 0xffff0000u),
                                        __uint_as_float(m.y << 16), __uint_as_float(m.y & 0xffff0000u));
              side[2 * i + 1] = make_float4(__uint_as_float(m.z << 16), __uint_as_float(m.z & 0xffff0000u),
                                            __uint_as_float(m.w << 16), __uint_as_float(m.w & 0xffff0000u));
            } else if (p.mask_src) {
              side[2 * i] = __ldg(reinterpret_cast<const float4*>(p.mask_src + e));
              side[2 * i + 1] = __ldg(reinterpret_cast<const float4*>(p.mask_src + e + 4));
            } else {
              side[2 * i] = *reinterpret_cast<const float4*>(oroww0 + i * row_stepw + n);
              side[2 * i + 1] = *reinterpret_cast<const float4*>(oroww0 + i * row_stepw + n + 4);
            }
          }
          return;
        }
        const int n = n0 + c * 32 + csub;
        if (n >= p.N) return;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (first + 4 * i >= limit) continue;
          if (p.mask_src)
            side[i] = __ldg(reinterpret_cast<const float4*>(p.mask_src + (grow0 + 4 * i) * p.N + n));
          else
            side[i] = *reinterpret_cast<const float4*>(orow0 + i * row_step + ncol(n));
        }
      };
      // bias / ReLU / dropout / mask / accumulate on 4 consecutive outputs starting at element e
      auto epi_math = [&](float4 r, const float4 bias4, const float4 sd, const uint64_t e) {
        r.x += bias4.x; r.y += bias4.y; r.z += bias4.z; r.w += bias4.w;
        if (p.relu) {
          r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
        }
        if (p.drop_p > 0.f) {
          const uint4 rnd = ssb::dropout_bits4(seed, p.site, e >> 2);
          r.x = rnd.x >= p.drop_thresh ? r.x * p.drop_scale : 0.f;
          r.y = rnd.y >= p.drop_thresh ? r.y * p.drop_scale : 0.f;
          r.z = rnd.z >= p.drop_thresh ? r.z * p.drop_scale : 0.f;
          r.w = rnd.w >= p.drop_thresh ? r.w * p.drop_scale : 0.f;
        }
        if (p.mask_src || p.mask_planes || p.mask_bits) {
          r.x = sd.x > 0.f ? r.x * p.mask_scale : 0.f;
          r.y = sd.y > 0.f ? r.y * p.mask_scale : 0.f;
          r.z = sd.z > 0.f ? r.z * p.mask_scale : 0.f;
          r.w = sd.w > 0.f ? r.w * p.mask_scale : 0.f;
        } else if (p.accumulate) {
          r.x += sd.x; r.y += sd.y; r.z += sd.z; r.w += sd.w;
        }
        return r;
      };
      if (it_mode != IT_PARTIAL) prefetch(cpar);
      if (p.mask_planes != nullptr && !p.mn_major &&
          tile + num_workers < (p.sk_rem ? p.sk_full : total_tiles)) {
        // The mask plane of the NEXT tile this CTA will finish is known now: pull its lines into L2
        // while this tile's accumulator is still being produced, so the register prefetch above
        // (one column block ahead) meets L2 latency instead of an HBM round trip.  (ncu, r2: 47 %
        // of the masked data-gradient's stall samples sat on the consumers of these loads.)
        int n0n, batchn, row0n, f0n, kbn, nkbn;
        decode(tile + num_workers, n0n, batchn, row0n, f0n, kbn, nkbn);
        const int rown = row0n + lg * 32 + lane;
        if (rown < p.rows_per_batch) {
          const __nv_bfloat16* mrow = p.mask_planes + ((int64_t)batchn * p.rows_per_batch + rown) * p.N + n0n;
#pragma unroll
          for (int c = cpar; c < BN / 32; c += 2)
            if (n0n + c * 32 < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(mrow + c * 32));
        }
      }
      mbar_wait(tfull_bar + 8 * acc, aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (it_mode == IT_OWNER) {
        // the other pieces of this tile were started at the same time as this one: a short wait
        if (lane == 0)
          for (int q = 0; q < sk_nc; ++q) {
            unsigned f;
            do {
              asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(f) : "l"(sk_flag(sk_contrib[q])) : "memory");
              if (!f) __nanosleep(64);
            } while (!f);
          }
        __threadfence();
        __syncwarp();
      }
#pragma unroll 1
      for (int c = cpar; c < BN / 32; c += 2) {
        if (n0 + c * 32 >= p.N) break;    // warp-uniform: nothing to store in this column block
        uint32_t v[32];
        tmem_ld32(tmem_d + (uint32_t)(c * 32), v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (it_mode == IT_PARTIAL) {      // raw accumulator -> this worker's workspace slot (coalesced)
          float* dst = sk_slot(worker) + (c >> 1) * 1024 + lane;
#pragma unroll
          for (int j = 0; j < 32; ++j) __stcg(dst + j * 32, __uint_as_float(v[j]));
          continue;
        }
        if (it_mode == IT_OWNER) {
          for (int q = 0; q < sk_nc; ++q) {
            const float* src = sk_slot(sk_contrib[q]) + (c >> 1) * 1024 + lane;
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldcg(src + j * 32));
          }
        }
        if (!group_live) continue;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(
                           stg + (uint32_t)lane * EPI_ROW_BYTES + (uint32_t)j * 4u),
                       "r"(v[j]), "r"(v[j + 1]), "r"(v[j + 2]), "r"(v[j + 3])
                       : "memory");
        __syncwarp();
        if (wide) {
          const int n = n0 + c * 32 + csubw;
          const bool n_ok = n < p.N;
          float4 biasA = make_float4(0.f, 0.f, 0.f, 0.f), biasB = biasA;
          if (p.bias && n_ok) {
            biasA = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            biasB = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
          }
          float4 oA[4], oB[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const uint32_t a = stg + (uint32_t)(i * 8 + rsubw) * EPI_ROW_BYTES + (uint32_t)csubw * 4u;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(oA[i].x), "=f"(oA[i].y), "=f"(oA[i].z), "=f"(oA[i].w) : "r"(a) : "memory");
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(oB[i].x), "=f"(oB[i].y), "=f"(oB[i].z), "=f"(oB[i].w) : "r"(a + 16u) : "memory");
          }
          __syncwarp();   // the staging tile is rewritten by the next column block
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (!n_ok || firstw + 8 * i >= limit) continue;
            const uint64_t e = (uint64_t)(groww0 + 8 * i) * (uint64_t)p.N + (uint64_t)n;
            if (p.drop_p > 0.f) {
              // wide map: ONE Philox call decides 8 consecutive outputs (16 random bits each, keep
              // iff bits >= p * 2^16) instead of two calls with 32 bits per output.  This mask is
              // generated here only (the backward reads it off the stored activation), so the
              // bit budget is private to this epilogue; the integer work per output halves.
              oA[i].x += biasA.x; oA[i].y += biasA.y; oA[i].z += biasA.z; oA[i].w += biasA.w;
              oB[i].x += biasB.x; oB[i].y += biasB.y; oB[i].z += biasB.z; oB[i].w += biasB.w;
              if (p.relu) {
                oA[i].x = fmaxf(oA[i].x, 0.f); oA[i].y = fmaxf(oA[i].y, 0.f);
                oA[i].z = fmaxf(oA[i].z, 0.f); oA[i].w = fmaxf(oA[i].w, 0.f);
                oB[i].x = fmaxf(oB[i].x, 0.f); oB[i].y = fmaxf(oB[i].y, 0.f);
                oB[i].z = fmaxf(oB[i].z, 0.f); oB[i].w = fmaxf(oB[i].w, 0.f);
              }
              const uint4 rnd = ssb::dropout_bits4(seed, p.site ^ 0x80000000u, e >> 3);
              const uint32_t t16 = p.drop_thresh >> 16;
              const float sc = p.drop_scale;
              oA[i].x = (rnd.x & 0xffffu) >= t16 ? oA[i].x * sc : 0.f;
              oA[i].y = (rnd.x >> 16) >= t16 ? oA[i].y * sc : 0.f;
              oA[i].z = (rnd.y & 0xffffu) >= t16 ? oA[i].z * sc : 0.f;
              oA[i].w = (rnd.y >> 16) >= t16 ? oA[i].w * sc : 0.f;
              oB[i].x = (rnd.z & 0xffffu) >= t16 ? oB[i].x * sc : 0.f;
              oB[i].y = (rnd.z >> 16) >= t16 ? oB[i].y * sc : 0.f;
              oB[i].z = (rnd.w & 0xffffu) >= t16 ? oB[i].z * sc : 0.f;
              oB[i].w = (rnd.w >> 16) >= t16 ? oB[i].w * sc : 0.f;
              if (p.mask_src || p.mask_planes || p.mask_bits) {
                const float4 sa = side[2 * i], sb = side[2 * i + 1];
                oA[i].x = sa.x > 0.f ? oA[i].x * p.mask_scale : 0.f; oA[i].y = sa.y > 0.f ? oA[i].y * p.mask_scale : 0.f;
                oA[i].z = sa.z > 0.f ? oA[i].z * p.mask_scale : 0.f; oA[i].w = sa.w > 0.f ? oA[i].w * p.mask_scale : 0.f;
                oB[i].x = sb.x > 0.f ? oB[i].x * p.mask_scale : 0.f; oB[i].y = sb.y > 0.f ? oB[i].y * p.mask_scale : 0.f;
                oB[i].z = sb.z > 0.f ? oB[i].z * p.mask_scale : 0.f; oB[i].w = sb.w > 0.f ? oB[i].w * p.mask_scale : 0.f;
              } else if (p.accumulate) {
                const float4 sa = side[2 * i], sb = side[2 * i + 1];
                oA[i].x += sa.x; oA[i].y += sa.y; oA[i].z += sa.z; oA[i].w += sa.w;
                oB[i].x += sb.x; oB[i].y += sb.y; oB[i].z += sb.z; oB[i].w += sb.w;
              }
              continue;
            }
            oA[i] = epi_math(oA[i], biasA, side[2 * i], e);
            oB[i] = epi_math(oB[i], biasB, side[2 * i + 1], e + 4);
          }
          if (c + 2 < BN / 32) prefetch(c + 2);
          if (p.out) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (!n_ok || firstw + 8 * i >= limit) continue;
              *reinterpret_cast<float4*>(oroww0 + i * row_stepw + n) = oA[i];
              *reinterpret_cast<float4*>(oroww0 + i * row_stepw + n + 4) = oB[i];
            }
          }
          if (p.mask_bits_out) {   // 1 bit per output: what the data gradient needs of this activation
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (!n_ok || firstw + 8 * i >= limit) continue;
              const uint32_t m = (oA[i].x > 0.f ? 1u : 0u) | (oA[i].y > 0.f ? 2u : 0u) |
                                 (oA[i].z > 0.f ? 4u : 0u) | (oA[i].w > 0.f ? 8u : 0u) |
                                 (oB[i].x > 0.f ? 16u : 0u) | (oB[i].y > 0.f ? 32u : 0u) |
                                 (oB[i].z > 0.f ? 64u : 0u) | (oB[i].w > 0.f ? 128u : 0u);
              p.mask_bits_out[((groww0 + 8 * i) * p.N + n) >> 3] = (uint8_t)m;
            }
          }
          if (p.planes) {   // x = hi + lo, the operand format of the next GEMM (no split pass)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (!n_ok || firstw + 8 * i >= limit) continue;
              float f[8] = {oA[i].x, oA[i].y, oA[i].z, oA[i].w, oB[i].x, oB[i].y, oB[i].z, oB[i].w};
              if (p.planes_lrelu) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = f[j] > 0.f ? f[j] : f[j] * p.planes_slope;
              }
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                const __nv_bfloat162 l = __floats2bfloat162_rn(f[2 * j] - __low2float(h),
                                                               f[2 * j + 1] - __high2float(h));
                hw[j] = *reinterpret_cast<const uint32_t*>(&h);
                lw[j] = *reinterpret_cast<const uint32_t*>(&l);
              }
              __nv_bfloat16* dst = p.planes + (groww0 + 8 * i) * p.N + n;
              *reinterpret_cast<uint4*>(dst) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
              *reinterpret_cast<uint4*>(dst + p.planes_stride) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
            }
          }
          continue;
        }
        const int n = n0 + c * 32 + csub;
        const bool n_ok = n < p.N;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && n_ok) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
        float4 o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(o[i].x), "=f"(o[i].y), "=f"(o[i].z), "=f"(o[i].w)
                       : "r"(stg + (uint32_t)(i * 4 + rsub) * EPI_ROW_BYTES + (uint32_t)csub * 4u)
                       : "memory");
        __syncwarp();   // the staging tile is rewritten by the next column block
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (!n_ok || first + 4 * i >= limit) continue;
          float4 r = o[i];
          float* dst = orow0 + i * row_step + ncol(n);
          if (p.atomic) {
            atomicAdd(reinterpret_cast<float4*>(dst), r);
            continue;
          }
          r.x += bias4.x; r.y += bias4.y; r.z += bias4.z; r.w += bias4.w;
          if (p.relu) {
            r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f);
          }
          if (p.drop_p > 0.f) {
            const uint64_t e = (uint64_t)(grow0 + 4 * i) * (uint64_t)p.N + (uint64_t)n;
            const uint4 rnd = ssb::dropout_bits4(seed, p.site, e >> 2);
            r.x = rnd.x >= p.drop_thresh ? r.x * p.drop_scale : 0.f;
            r.y = rnd.y >= p.drop_thresh ? r.y * p.drop_scale : 0.f;
            r.z = rnd.z >= p.drop_thresh ? r.z * p.drop_scale : 0.f;
            r.w = rnd.w >= p.drop_thresh ? r.w * p.drop_scale : 0.f;
          }
          if (p.mask_src) {
            const float4 mk = side[i];
            r.x = mk.x > 0.f ? r.x * p.mask_scale : 0.f;
            r.y = mk.y > 0.f ? r.y * p.mask_scale : 0.f;
            r.z = mk.z > 0.f ? r.z * p.mask_scale : 0.f;
            r.w = mk.w > 0.f ? r.w * p.mask_scale : 0.f;
          } else if (p.accumulate) {
            const float4 old = side[i];
            r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
          }
          o[i] = r;
        }
        // next block's side operands go out before this block's stores queue up behind them
        if (c + 2 < BN / 32) prefetch(c + 2);
        if (p.out) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (!n_ok || first + 4 * i >= limit || p.atomic) continue;
            *reinterpret_cast<float4*>(orow0 + i * row_step + ncol(n)) = o[i];
          }
        }
      }
      if (it_mode == IT_PARTIAL) {
        __threadfence();                  // every lane's partial stores before the flag
        __syncwarp();
        if (lane == 0) {
          unsigned one = 1u;
          asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(sk_flag(worker)), "r"(one) : "memory");
        }
      } else if (it_mode == IT_OWNER) {
        __syncwarp();                     // all lanes have read the partials: re-arm the flags (0 between launches)
        if (lane == 0)
          for (int q = 0; q < sk_nc; ++q) *sk_flag(sk_contrib[q]) = 0u;
      }
      // hand the accumulator back to the (leader's) MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if constexpr (CTAS == 1) mbar_arrive(tempty_bar + 8 * acc);
        else mbar_arrive_cluster(tempty_bar_lead + 8 * acc);
      }
    }
  }

  // teardown: nobody may leave while the peer can still touch this CTA's shared memory / TMEM
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (CTAS == 2) cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (CTAS == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "r"((uint32_t)TMEM_COLS)
                   : "memory");
  }
}

// ---- fp32 -> (hi, lo) bf16 planes ------------------------------------------------------------
// out layout: [2][n] bf16 (plane 0 = hi, plane 1 = lo); 8 elements per thread-iteration
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t n8,
             int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(x) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(x) + 2 * i + 1);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hi[j] = __float2bfloat16_rn(v[j]);
      lo[j] = __float2bfloat16_rn(v[j] - __bfloat162float(hi[j]));
    }
    *reinterpret_cast<uint4*>(out + 8 * i) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(out + n + 8 * i) = *reinterpret_cast<const uint4*>(lo);
  }
}

// strided variant: the leading `cols` columns of rows `ld_in` apart -> the same columns of planes
// with rows `ld_out` apart (one third of a (M, 3D) operand: the others are written by their producer)
__global__ void __launch_bounds__(256)
split_2d_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int64_t rows, int c8,
                int64_t ld_in, int64_t ld_out, int64_t plane_stride) {
  const int64_t n8 = rows * c8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / c8;
    const int c = (int)(i - r * c8) * 8;
    const float4 a = __ldg(reinterpret_cast<const float4*>(x + r * ld_in + c));
    const float4 b = __ldg(reinterpret_cast<const float4*>(x + r * ld_in + c + 4));
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    __align__(16) __nv_bfloat16 hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hi[j] = __float2bfloat16_rn(v[j]);
      lo[j] = __float2bfloat16_rn(v[j] - __bfloat162float(hi[j]));
    }
    *reinterpret_cast<uint4*>(out + r * ld_out + c) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(out + plane_stride + r * ld_out + c) = *reinterpret_cast<const uint4*>(lo);
  }
}

// transposed variant: planes[p][c][r] = split(x[r][c]) for a row-major (rows, cols) matrix: the
// data-gradient operand W^T of an nn.Linear weight without a transposed fp32 copy.
// block (32, 8) handles a 64-row x 32-column tile through shared memory; stores are bf16x2.
__global__ void __launch_bounds__(256)
split_t_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int rows, int cols,
               int64_t plane_stride) {
  __shared__ float tile[64][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 32;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = r0 + ty + 8 * i, c = c0 + tx;
    tile[ty + 8 * i][tx] = (r < rows && c < cols) ? __ldg(x + (int64_t)r * cols + c) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = c0 + ty + 8 * i, r = r0 + 2 * tx;
    if (c >= cols || r >= rows) continue;      // rows is even: r + 1 < rows too
    const float a = tile[2 * tx][ty + 8 * i], b = tile[2 * tx + 1][ty + 8 * i];
    const __nv_bfloat162 hi = __floats2bfloat162_rn(a, b);
    const __nv_bfloat162 lo = __floats2bfloat162_rn(a - __low2float(hi), b - __high2float(hi));
    __nv_bfloat16* dst = out + (int64_t)c * rows + r;
    *reinterpret_cast<__nv_bfloat162*>(dst) = hi;
    *reinterpret_cast<__nv_bfloat162*>(dst + plane_stride) = lo;
  }
}

// ---- tensor-map construction (driver entry point fetched at run time; no libcuda link) --------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::mutex g_encode_mutex;

int get_encode(EncodeTiledFn* out) {
  std::lock_guard<std::mutex> lk(g_encode_mutex);
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SSB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    SSB_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess,
                "gemm_tc: cuTensorMapEncodeTiled not available from the driver");
    g_encode = (EncodeTiledFn)fn;
  }
  *out = g_encode;
  return SSB_OK;
}

// 5-D bf16 map: dims (inner, rows, batch_lo, batch_hi, plane); strides in ELEMENTS for dims 1..4
int make_map(CUtensorMap* map, const void* base, int64_t inner, int64_t rows, int64_t n_lo,
             int64_t n_hi, int64_t s_row, int64_t s_lo, int64_t s_hi, int64_t s_plane,
             int box_inner, int box_rows, int row_elem_stride) {
  // K-major boxes are BK = 32 elements (64 B) wide -> SWIZZLE_64B; MN-major boxes are 64 wide
  const CUtensorMapSwizzle swz =
      box_inner * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B;
  EncodeTiledFn enc = nullptr;
  if (int rc = get_encode(&enc)) return rc;
  SSB_REQUIRE(((uintptr_t)base & 15) == 0, "gemm_tc: operand base must be 16 B aligned");
  if (n_lo <= 1) s_lo = s_plane;   // unused dims still need a legal (16 B multiple) stride
  if (n_hi <= 1) s_hi = s_plane;
  SSB_REQUIRE(s_row % 8 == 0 && s_lo % 8 == 0 && s_hi % 8 == 0 && s_plane % 8 == 0,
              "gemm_tc: operand strides must be multiples of 8 elements (%lld,%lld,%lld,%lld)",
              (long long)s_row, (long long)s_lo, (long long)s_hi, (long long)s_plane);
  cuuint64_t dims[5] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)(n_lo > 0 ? n_lo : 1),
                        (cuuint64_t)(n_hi > 0 ? n_hi : 1), 2};
  cuuint64_t strides[4] = {(cuuint64_t)s_row * 2, (cuuint64_t)s_lo * 2, (cuuint64_t)s_hi * 2,
                           (cuuint64_t)s_plane * 2};
  cuuint32_t box[5] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1, 1, 1};
  cuuint32_t estr[5] = {1, (cuuint32_t)row_elem_stride, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ssb::set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) inner=%lld rows=%lld lo=%lld hi=%lld "
                   "strides=(%lld,%lld,%lld,%lld) box=(%d,%d) estr=%d",
                   (int)r, (long long)inner, (long long)rows, (long long)n_lo, (long long)n_hi,
                   (long long)s_row, (long long)s_lo, (long long)s_hi, (long long)s_plane,
                   box_inner, box_rows, row_elem_stride);
    return SSB_ERR_ARG;
  }
  return SSB_OK;
}

// operand -> (n_lo, n_hi): batch_div <= 0 means a single batch level
inline void batch_levels(const ssb_tc_operand_t* o, int64_t* n_lo, int64_t* n_hi) {
  if (o->batch_div > 0 && o->batch_div < o->batches) {
    *n_lo = o->batch_div;
    *n_hi = (o->batches + o->batch_div - 1) / o->batch_div;
  } else {
    *n_lo = o->batches;
    *n_hi = 1;
  }
}

// CTA pairs unless SSB_TC_CTAS=1 (read once); callers fall back to 1 per problem shape
int default_ctas() {
  static int v = 0;
  if (!v) {
    const char* e = getenv("SSB_TC_CTAS");
    v = (e && e[0] == '1') ? 1 : 2;
  }
  return v;
}

// N-adaptive MMAs for N < 256 (SSB_TC_NARROW=0 restores full-width MMAs; read once)
bool narrow_n_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SSB_TC_NARROW");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// stream-K workspace (ssb_gemm_tc_set_streamk_workspace): caller-owned, one per device.
// Measured on B200 (r2 session 25, tools/tc_gemm_check.py streamk; on / off, same process):
//   16000 x 768 x 3072   162.6 / 162.2 us      16000 x 3072 x 768   196.3 / 175.5 us
//   16000 x 768 x 768     77.6 /  59.9 us      32000 x 768 x 2304   237.0 / 239.2 us
//   4800 x 256 x 768      44.3 /  27.8 us      4800 x 256 x 2816     52.2 /  54.7 us
// cfg-1 step 24.95 / 24.35 ms, vocoder 3.40 / 2.89 ms.  The partial accumulators of a remainder tile
// are as many bytes as the tile's output per piece (128 KB per CTA written, then read by the owner, at
// the ~30 - 60 GB/s one SM's 8 epilogue warps pull from L2), all at the end of the kernel: 10 - 30 us
// of exchange against 0.45 waves (4 - 17 us) of MMAs saved.  It pays only where a tile is long
// (K >= 2816) and the grid badly underfilled, so it stays opt-in (SSB_STREAMK=1) - bit-reproducible,
// schedule checked on the CPU (tests/test_streamk_schedule_cpu.py).
struct SkWorkspace { void* base; int64_t bytes; };
SkWorkspace g_sk_ws[64] = {};
std::mutex g_sk_mutex;
constexpr int64_t SK_SLOT_BYTES = (int64_t)EPI_WARPS * 4 * 32 * 32 * 4;   // one CTA's raw 128 x 256 accumulator
constexpr int64_t SK_FLAG_BYTES = 8192;                                    // >= 148 CTAs x 8 warps x 4 B (4736)

inline int64_t sk_workspace_bytes() { return SK_FLAG_BYTES + (int64_t)ssb::num_sms() * SK_SLOT_BYTES; }

bool streamk_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SSB_STREAMK");
    v = (e && e[0] == '1') ? 1 : 0;   // opt-in: measured slower on every cfg-1 / vocoder shape (see below)
  }
  return v != 0;
}

// Fill the stream-K fields for a K-major launch of `total` tiles on `workers` workers (CTAs or CTA
// pairs); returns the number of workers the grid needs.
int plan_streamk(TcParams* p, int64_t total, int workers) {
  p->sk_full = (int)total; p->sk_rem = 0; p->sk_workers = 0; p->sk_ws = nullptr; p->sk_flags = nullptr;
  const int grid_workers = (int)(total < workers ? total : workers);
  int dev = 0;
  if (!streamk_enabled() || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return grid_workers;
  SkWorkspace ws;
  {
    std::lock_guard<std::mutex> lk(g_sk_mutex);
    ws = g_sk_ws[dev];
  }
  const int rem = (int)(total % workers);
  // worth it when the partial wave leaves a quarter of the machine idle and a piece still holds a
  // few k-blocks (a piece costs one accumulator round trip through L2 on top of its MMAs)
  if (!ws.base || ws.bytes < sk_workspace_bytes() || rem == 0 || rem * 4 > workers * 3 || p->num_kb < 16)
    return grid_workers;
  p->sk_rem = rem;
  p->sk_full = (int)(total - rem);
  p->sk_workers = rem * 4 < workers ? rem * 4 : workers;
  p->sk_flags = (unsigned*)ws.base;
  p->sk_ws = (float*)((char*)ws.base + SK_FLAG_BYTES);
  return grid_workers > p->sk_workers ? grid_workers : p->sk_workers;
}

template <int CTAS>
int launch_impl(const CUtensorMap& mapA, const CUtensorMap& mapB, const TcParams& p, int64_t total,
                cudaStream_t st) {
  static bool attr_set[64] = {false};
  int dev = 0;
  SSB_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr_set[dev]) {
    SSB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  Cfg<CTAS>::SMEM_BYTES));
    attr_set[dev] = true;
  }
  const int workers = ssb::num_sms() / CTAS;
  int grid = (int)(total < workers ? total : workers) * CTAS;
  TcParams pk = p;
  if (p.sk_allow && !p.mn_major && p.splits == 1)
    grid = plan_streamk(&pk, total, workers) * CTAS;
  else {
    pk.sk_full = (int)total; pk.sk_rem = 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = Cfg<CTAS>::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (CTAS > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = CTAS;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (ssb::pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  SSB_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<CTAS>, mapA, mapB, pk));
  SSB_LAUNCH_CHECK("gemm_tc_kernel");
  return SSB_OK;
}

// tiles = (n tiles, row tiles of 128 * ctas, splits)
int launch(const CUtensorMap& mapA, const CUtensorMap& mapB, TcParams p, dim3 tiles, int ctas,
           cudaStream_t st) {
  p.tiles_x = (int)tiles.x;
  p.tiles_y = (int)tiles.y;
  p.splits = (int)tiles.z;
  const int64_t total = (int64_t)tiles.x * tiles.y * tiles.z;
  return ctas == 2 ? launch_impl<2>(mapA, mapB, p, total, st)
                   : launch_impl<1>(mapA, mapB, p, total, st);
}

int fill_epi(const ssb_epilogue_t* e, int64_t N, TcParams* p) {
  SSB_REQUIRE(e && (e->out.base || e->planes_out), "gemm_tc: null output");
  SSB_REQUIRE(N % 4 == 0 && e->out.ld >= N && e->out.ld % 4 == 0 && e->out.batch_stride % 4 == 0 &&
                  ((uintptr_t)e->out.base & 15) == 0,
              "gemm_tc: bad output geometry / alignment");
  SSB_REQUIRE(!(e->mask_src && e->mask_planes) && !(e->mask_bits && (e->mask_src || e->mask_planes)),
              "gemm_tc: mask_src, mask_planes and mask_bits are exclusive");
  SSB_REQUIRE((!e->mask_bits && !e->mask_bits_out) || N % 8 == 0, "gemm_tc: mask bits need N %% 8 == 0");
  SSB_REQUIRE(!e->mask_bits_out || e->planes_out, "gemm_tc: mask_bits_out rides on the plane-emitting epilogue");
  SSB_REQUIRE(((uintptr_t)e->planes_out & 15) == 0 && ((uintptr_t)e->mask_planes & 15) == 0 &&
                  e->planes_stride % 8 == 0 && ((!e->planes_out && !e->mask_planes) || N % 8 == 0),
              "gemm_tc: split-plane epilogue operands need 16 B alignment and N %% 8 == 0");
  SSB_REQUIRE(e->out.base || !e->accumulate, "gemm_tc: accumulate needs an fp32 output");
  SSB_REQUIRE(!e->planes_lrelu || e->planes_out, "gemm_tc: planes_lrelu rides on the plane-emitting epilogue");
  p->planes_lrelu = e->planes_lrelu; p->planes_slope = e->planes_neg_slope;
  p->planes = (__nv_bfloat16*)e->planes_out; p->planes_stride = e->planes_stride;
  p->mask_planes = (const __nv_bfloat16*)e->mask_planes;
  p->mask_bits = (const uint8_t*)e->mask_bits;
  p->mask_bits_out = (uint8_t*)e->mask_bits_out;
  SSB_REQUIRE(e->drop_p >= 0.f && e->drop_p < 1.f, "gemm_tc: bad dropout p");
  p->out = e->out.base; p->out_batch_stride = e->out.batch_stride; p->out_ld = e->out.ld;
  p->out_batch_stride_hi = e->out.batch_stride_hi;
  p->out_dt = e->out.d_t; p->out_doff = e->out.d_off;
  p->bias = e->bias; p->mask_src = e->mask_src; p->mask_scale = e->mask_scale;
  p->relu = e->relu; p->accumulate = e->accumulate; p->atomic = 0;
  p->drop_p = e->drop_p; p->drop_scale = e->drop_p > 0.f ? 1.f / (1.f - e->drop_p) : 1.f;
  const double th = (double)e->drop_p * 4294967296.0;
  p->drop_thresh = th >= 4294967295.0 ? 0xffffffffu : (uint32_t)th;
  p->seed = e->seed; p->seed_src = ssb::seed_source(); p->site = e->site;
  p->N = (int)N;
  return SSB_OK;
}

}  // namespace

extern "C" {

int ssb_split_bf16(const float* x, int64_t n, void* planes, void* stream) {
  if (n == 0) return SSB_OK;
  SSB_REQUIRE(x && planes && n % 8 == 0, "split_bf16: n=%lld must be a multiple of 8",
              (long long)n);
  SSB_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)planes & 15) == 0,
              "split_bf16: pointers must be 16 B aligned");
  const int64_t n8 = n / 8;
  const int64_t blocks = (n8 + 255) / 256;
  const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
  split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)planes, n8, n);
  SSB_LAUNCH_CHECK("split_kernel");
  return SSB_OK;
}

int ssb_split_bf16_2d(const float* x, int64_t rows, int64_t cols, int64_t ld_in, void* planes,
                      int64_t ld_out, int64_t plane_stride, void* stream) {
  if (rows == 0 || cols == 0) return SSB_OK;
  SSB_REQUIRE(x && planes && rows > 0 && cols > 0 && cols % 8 == 0 && ld_in % 4 == 0 && ld_out % 8 == 0 &&
                  ld_in >= cols && ld_out >= cols && plane_stride % 8 == 0,
              "split_bf16_2d: cols=%lld ld_in=%lld ld_out=%lld (cols, ld_out multiples of 8)",
              (long long)cols, (long long)ld_in, (long long)ld_out);
  SSB_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)planes & 15) == 0,
              "split_bf16_2d: pointers must be 16 B aligned");
  const int64_t n8 = rows * (cols / 8);
  const int64_t blocks = (n8 + 255) / 256;
  const int grid = (int)(blocks < 148 * 8 ? blocks : 148 * 8);
  split_2d_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)planes, rows, (int)(cols / 8),
                                                          ld_in, ld_out, plane_stride);
  SSB_LAUNCH_CHECK("split_2d_kernel");
  return SSB_OK;
}

int ssb_split_bf16_t(const float* x, int64_t rows, int64_t cols, void* planes, void* stream) {
  if (rows == 0 || cols == 0) return SSB_OK;
  SSB_REQUIRE(x && planes && rows > 0 && cols > 0 && rows % 8 == 0 && cols % 8 == 0 &&
                  rows < (1 << 30) && cols < (1 << 30),
              "split_bf16_t: rows=%lld cols=%lld must be multiples of 8", (long long)rows,
              (long long)cols);
  SSB_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)planes & 15) == 0,
              "split_bf16_t: pointers must be 16 B aligned");
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 63) / 64));
  split_t_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)planes, (int)rows,
                                                                (int)cols, rows * cols);
  SSB_LAUNCH_CHECK("split_t_kernel");
  return SSB_OK;
}

int64_t ssb_gemm_tc_streamk_workspace_bytes(void) { return sk_workspace_bytes(); }

int ssb_gemm_tc_set_streamk_workspace(void* workspace, int64_t workspace_bytes) {
  int dev = 0;
  SSB_CUDA(cudaGetDevice(&dev));
  SSB_REQUIRE(dev >= 0 && dev < 64, "gemm_tc: device index out of range");
  SSB_REQUIRE(!workspace || (workspace_bytes >= sk_workspace_bytes() && ((uintptr_t)workspace & 255) == 0),
              "gemm_tc: stream-K workspace needs %lld bytes, 256 B aligned", (long long)sk_workspace_bytes());
  std::lock_guard<std::mutex> lk(g_sk_mutex);
  g_sk_ws[dev].base = workspace;
  g_sk_ws[dev].bytes = workspace ? workspace_bytes : 0;
  return SSB_OK;
}

int ssb_gemm_tc_kmajor(const ssb_tc_operand_t* A, const void* Bplanes, int64_t N, int64_t K,
                       const ssb_epilogue_t* epi, void* stream) {
  SSB_REQUIRE(A && A->planes && Bplanes, "gemm_tc: null operand");
  // taps: 1 / 3 in the transduction model; HiFi-GAN's dilated k = 3 / 7 / 11 and stride-phase
  // transposed convolutions (silent_speech_b200/vocoder.py) use up to 16 at tap step = dilation
  SSB_REQUIRE(K % BK == 0 && A->C % BK == 0 && K % A->C == 0 && K / A->C <= 16,
              "gemm_tc: K=%lld / C=%d must be multiples of 32 (K = taps*C, taps <= 16)",
              (long long)K, A->C);
  SSB_REQUIRE(A->batches >= 1 && A->rows_out >= 1 && A->L_src >= 1 && (A->s_t == 1 || A->s_t == 2),
              "gemm_tc: bad A geometry");
  TcParams p = {};
  if (int rc = fill_epi(epi, N, &p)) return rc;
  SSB_REQUIRE(epi->out.rows_per_batch == A->rows_out, "gemm_tc: output rows_per_batch mismatch");
  CUtensorMap mapA, mapB;
  const int box_rows = BM * A->s_t;   // 128 rows at traversal stride s_t
  int64_t n_lo, n_hi;
  batch_levels(A, &n_lo, &n_hi);
  if (int rc = make_map(&mapA, A->planes, A->C, A->L_src, n_lo, n_hi, A->ld, A->batch_stride,
                        A->batch_stride_hi, A->plane_stride, BK, box_rows, A->s_t))
    return rc;
  // a CTA pair covers 256 rows: not worth it when a batch item has no second half tile
  const int ctas = A->rows_out > BM ? default_ctas() : 1;
  const int bmc = BM * ctas;
  // narrow outputs (HiFi-GAN's 32 ... 128 channels, the 128-wide output heads): an MMA of N =
  // round32(N) instead of 256 columns of which most would multiply zero-filled rows of B
  const int n_mma = N < BN && narrow_n_enabled() ? (int)((N + 31) / 32 * 32) : 0;
  if (int rc = make_map(&mapB, Bplanes, K, N, 1, 1, K, N * K, N * K, N * K, BK,
                        (n_mma ? n_mma : BN) / ctas, 1))
    return rc;
  p.n_mma = n_mma;
  p.sk_allow = 1;
  p.batch_div = (int)n_lo;
  p.b_mode = 0;
  p.a_inner = A->C; p.a_row_step = A->s_t; p.a_tap_step = A->s_tap; p.a_off = A->off;
  p.rows_per_batch = A->rows_out;
  p.tiles_per_batch = (A->rows_out + bmc - 1) / bmc;
  p.chunks_per_batch = 1;
  p.num_kb = (int)(K / BK);
  p.kb_per_split = p.num_kb;
  p.M_valid_total = A->batches * A->rows_out;
  p.mn_major = 0;
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)(A->batches * p.tiles_per_batch), 1);
  return launch(mapA, mapB, p, grid, ctas, (cudaStream_t)stream);
}

int ssb_gemm_tc_wgrad(const ssb_tc_operand_t* X, const void* Gplanes, int64_t g_plane_stride,
                      int64_t N, int64_t K, float* dW, int64_t lddw, int64_t group_w,
                      int64_t group_stride, int accumulate, void* stream) {
  SSB_REQUIRE(X && X->planes && Gplanes && dW, "gemm_tc_wgrad: null operand");
  SSB_REQUIRE(K % BM == 0 && X->C % BM == 0 && K % X->C == 0 && K / X->C <= 3,
              "gemm_tc_wgrad: K=%lld / C=%d must be multiples of 128", (long long)K, X->C);
  SSB_REQUIRE(N % 8 == 0 && lddw % 4 == 0 && ((uintptr_t)dW & 15) == 0 &&
                  (group_w > 0 ? (group_w % 4 == 0 && N % group_w == 0 && lddw >= group_w &&
                                  group_stride % 4 == 0 && group_stride >= (K - 1) * lddw + group_w)
                               : lddw >= N),
              "gemm_tc_wgrad: bad N / dW geometry");
  SSB_REQUIRE(group_w == 0 || accumulate, "gemm_tc_wgrad: grouped output needs accumulate (no 2-D clear)");
  SSB_REQUIRE(X->batches >= 1 && X->rows_out >= 1 && (X->s_t == 1 || X->s_t == 2),
              "gemm_tc_wgrad: bad X geometry");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap mapA, mapB;
  if (int rc = make_map(&mapA, X->planes, X->C, X->L_src, X->batches, 1, X->ld, X->batch_stride,
                        X->plane_stride, X->plane_stride, 64, BK * X->s_t, X->s_t))
    return rc;
  // G: (batches, rows_out, N) row-major planes
  if (int rc = make_map(&mapB, Gplanes, N, X->rows_out, X->batches, 1, N,
                        (int64_t)X->rows_out * N, g_plane_stride, g_plane_stride, 64, BK, 1))
    return rc;
  TcParams p = {};
  p.batch_div = X->batches;
  p.b_mode = 1;
  p.out = dW; p.out_ld = (int)lddw; p.N = (int)N;
  p.grp_w = (int)group_w; p.grp_stride = group_stride;
  p.out_dt = 1; p.out_doff = 0;   // output row f of dW
  p.a_inner = X->C; p.a_row_step = X->s_t; p.a_tap_step = X->s_tap; p.a_off = X->off;
  p.rows_per_batch = X->rows_out;
  p.tiles_per_batch = 1;
  p.chunks_per_batch = (X->rows_out + BK - 1) / BK;
  p.num_kb = X->batches * p.chunks_per_batch;
  p.M_valid_total = (int)K;
  p.mn_major = 1;
  const int ctas = K % (2 * BM) == 0 ? default_ctas() : 1;   // feature rows pair up
  const int bmc = BM * ctas;
  const int tiles = (int)((K / bmc) * ((N + BN - 1) / BN));
  // split the token reduction so that tiles*splits fills whole waves of the persistent grid:
  // maximise (tiles*s) / (ceil(tiles*s / workers) * workers); each split keeps >= 8 k-blocks
  const int sms = (ssb::num_sms() > 0 ? ssb::num_sms() : 148) / ctas;
  int max_s = p.num_kb / 8;
  if (max_s > 32) max_s = 32;
  if (max_s < 1) max_s = 1;
  int splits = 1;
  double best = 0.0;
  for (int sp = 1; sp <= max_s; ++sp) {
    const int tt = tiles * sp;
    const double eff = (double)tt / (double)(((tt + sms - 1) / sms) * sms);
    if (eff > best + 0.02) {
      best = eff;
      splits = sp;
    }
  }
  p.kb_per_split = (p.num_kb + splits - 1) / splits;
  splits = (p.num_kb + p.kb_per_split - 1) / p.kb_per_split;
  p.atomic = splits > 1 ? 1 : 0;
  p.accumulate = accumulate;
  if (splits > 1 && !accumulate)
    SSB_CUDA(cudaMemset2DAsync(dW, (size_t)lddw * 4, 0, (size_t)N * 4, (size_t)K, st));
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)(K / bmc), (unsigned)splits);
  return launch(mapA, mapB, p, grid, ctas, st);
}

int ssb_gemm_tc_batched(const ssb_tc_operand_t* A, const ssb_tc_operand_t* B, int b_mode,
                        int64_t N, int64_t K, const ssb_epilogue_t* epi, void* stream) {
  SSB_REQUIRE(A && B && A->planes && B->planes, "gemm_tc_batched: null operand");
  SSB_REQUIRE(K % BK == 0 && A->C == K && B->C == K, "gemm_tc_batched: K=%lld must be a multiple "
              "of 64 and equal both inner extents (%d, %d)", (long long)K, A->C, B->C);
  SSB_REQUIRE(A->batches >= 1 && A->rows_out >= 1 && A->s_t == 1 && b_mode >= 0 && b_mode <= 2,
              "gemm_tc_batched: bad geometry");
  TcParams p = {};
  if (int rc = fill_epi(epi, N, &p)) return rc;
  SSB_REQUIRE(!p.planes && !p.mask_planes && !p.mask_bits && !p.mask_bits_out,
              "gemm_tc batched: split-plane epilogue operands unsupported");
  SSB_REQUIRE(epi->out.rows_per_batch == A->rows_out, "gemm_tc_batched: rows_per_batch mismatch");
  int64_t a_lo, a_hi, b_lo, b_hi;
  batch_levels(A, &a_lo, &a_hi);
  batch_levels(B, &b_lo, &b_hi);
  CUtensorMap mapA, mapB;
  if (int rc = make_map(&mapA, A->planes, A->C, A->L_src, a_lo, a_hi, A->ld, A->batch_stride,
                        A->batch_stride_hi, A->plane_stride, BK, BM, 1))
    return rc;
  const int ctas = A->rows_out > BM ? default_ctas() : 1;
  const int bmc = BM * ctas;
  if (int rc = make_map(&mapB, B->planes, B->C, B->L_src, b_lo, b_hi, B->ld, B->batch_stride,
                        B->batch_stride_hi, B->plane_stride, BK, BN / ctas, 1))
    return rc;
  p.a_inner = A->C; p.a_row_step = 1; p.a_tap_step = 0; p.a_off = A->off;
  p.rows_per_batch = A->rows_out;
  p.tiles_per_batch = (A->rows_out + bmc - 1) / bmc;
  p.chunks_per_batch = 1;
  p.num_kb = (int)(K / BK);
  p.kb_per_split = p.num_kb;
  p.M_valid_total = A->batches * A->rows_out;
  p.mn_major = 0;
  p.batch_div = (int)a_lo;
  p.b_mode = b_mode;
  dim3 tiles((unsigned)((N + BN - 1) / BN), (unsigned)(A->batches * p.tiles_per_batch), 1);
  return launch(mapA, mapB, p, tiles, ctas, (cudaStream_t)stream);
}

int ssb_gemm_tc_batched_tn(const ssb_tc_operand_t* X, const ssb_tc_operand_t* G, int64_t N,
                           int64_t K, const ssb_epilogue_t* epi, void* stream) {
  SSB_REQUIRE(X && G && X->planes && G->planes, "gemm_tc_batched_tn: null operand");
  SSB_REQUIRE(X->C >= K && G->C >= N && X->batches == G->batches && X->rows_out == G->rows_out &&
                  X->s_t == 1 && G->s_t == 1,
              "gemm_tc_batched_tn: operand geometry mismatch");
  TcParams p = {};
  if (int rc = fill_epi(epi, N, &p)) return rc;
  SSB_REQUIRE(!p.planes && !p.mask_planes && !p.mask_bits && !p.mask_bits_out,
              "gemm_tc batched: split-plane epilogue operands unsupported");
  SSB_REQUIRE(epi->out.rows_per_batch == K, "gemm_tc_batched_tn: output rows_per_batch must be K");
  int64_t x_lo, x_hi, g_lo, g_hi;
  batch_levels(X, &x_lo, &x_hi);
  batch_levels(G, &g_lo, &g_hi);
  SSB_REQUIRE(x_lo == g_lo && x_hi == g_hi, "gemm_tc_batched_tn: batch levels differ");
  CUtensorMap mapA, mapB;
  if (int rc = make_map(&mapA, X->planes, X->C, X->L_src, x_lo, x_hi, X->ld, X->batch_stride,
                        X->batch_stride_hi, X->plane_stride, 64, BK, 1))
    return rc;
  if (int rc = make_map(&mapB, G->planes, G->C, G->L_src, g_lo, g_hi, G->ld, G->batch_stride,
                        G->batch_stride_hi, G->plane_stride, 64, BK, 1))
    return rc;
  p.a_inner = X->C; p.a_row_step = 1; p.a_tap_step = 0; p.a_off = X->off;
  p.rows_per_batch = X->rows_out;
  p.tiles_per_batch = 1;
  p.chunks_per_batch = (X->rows_out + BK - 1) / BK;
  p.num_kb = X->batches * p.chunks_per_batch;
  p.kb_per_split = p.chunks_per_batch;
  p.M_valid_total = (int)K;
  p.mn_major = 1;
  p.mn_batched = 1;
  p.batch_div = (int)x_lo;
  p.b_mode = 1;
  const int ctas = K > BM ? default_ctas() : 1;
  const int bmc = BM * ctas;
  dim3 tiles((unsigned)((N + BN - 1) / BN), (unsigned)((K + bmc - 1) / bmc), (unsigned)X->batches);
  return launch(mapA, mapB, p, tiles, ctas, (cudaStream_t)stream);
}

}  // extern "C"
