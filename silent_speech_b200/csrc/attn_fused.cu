// Fused banded relative-position attention on tcgen05 (sm_100a): forward and backward.
//
// Same mathematics as attn.cu / attn_tc.cu (transformer.py:99-110 with the relative-position
// logits of :162-297 in closed form):  logits[q,k] = q.k / sqrt(dh) + q.E[k-q+W]  inside the band
// |k-q| <= W, softmax over k, dropout on the probabilities, O = P_drop V.  The multi-kernel
// tensor-core schedule (attn_tc.cu) materialises dense (T x T) logits, probabilities and their
// gradients in HBM: ~3.5 GB of traffic per layer for 13 GFLOP of band arithmetic, 29 % of the
// cfg-1 training step.  Here nothing T x T leaves the SM:
//
//   forward   one CTA per (b, h, 128 queries); TMEM lane = query.  Q stays in shared memory,
//             keys stream in 32-row chunks: S = Q K^T for the whole +-W window sits in TMEM
//             (<= 352 fp32 columns); one thread per query row adds the positional logits from
//             a 128 x RW shared-memory tile of R = Q E^T (skewed read, bank-conflict free),
//             takes max / exp / sum, applies dropout and writes un-normalised P as bf16 hi/lo
//             planes straight into the canonical SWIZZLE_64B K-major operand layout;
//             O += P V_chunk accumulates in TMEM (V read MN-major from its natural [key][d]
//             layout: no transposed copy), divided by the row sum at the end.  Saves (max, 1/sum).
//   backward  one CTA per (b, h, 128 keys); TMEM lane = key, queries stream in 32-row chunks.
//             S^T = K Q_c^T and dP^T = V dO_c^T are recomputed on the tensor cores, one thread
//             per key forms P and dS = P (dP_drop - delta) from the saved row statistics, and the
//             P_drop^T / dS^T tiles (software-written operand layout again) feed
//             dV += P_drop^T dO_c, dK += dS^T Q_c (TMEM accumulators, stored once per tile) and
//             dQ_c^T = K^T dS^T (lanes = head dim, red.global.add into dQ, coalesced).  dS is
//             also written in band layout (bf16 planes) for the positional part of dQ, which
//             stays a batched tensor-core GEMM with E (no gradient flows to E: SURVEY.md F3).
//
// Every product is bf16x3 (hi*hi + hi*lo + lo*hi, fp32 accumulation) like gemm_tc.cu.  The
// operand formats used here (software-written SW64/SW128 K-major and MN-major tiles, one tile
// under two roles) are pinned by tools/umma_probe.cu.
#include "ssb_common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <math_constants.h>
#include <mutex>

namespace {

constexpr int QT = 128;          // TMEM lanes per tile (queries fwd, keys bwd)
constexpr int CH = 32;           // streamed chunk rows (keys fwd, queries bwd)
constexpr int DB = 32;           // head-dim block: one SWIZZLE_64B row (64 B of bf16)
constexpr int BLK_BIG = QT * 64;   // bytes of a [128 rows][32 d] block  (8 KB)
constexpr int BLK_SMALL = CH * 64; // bytes of a [32 rows][32 d] block   (2 KB)
constexpr int MAX_NDB = 3;       // dh <= 96

constexpr int KG = 3;            // forward: key chunks per S MMA (N = 32 KG)
constexpr int KST = 2 * KG;      // forward: K ring depth = two MMA groups
constexpr int VST = 4;           // forward: V ring depth
constexpr int NCW = 16;          // element-wise warps: 4 per TMEM lane group (a lone warp per
                                 // scheduler cannot hide its own latencies: 341 / 708 us per layer
                                 // with 4 warps)
constexpr int NCT = NCW * 32;    // 512 element-wise threads
constexpr int NTHREADS = NCT + 64;   // + MMA-issue warp (NCW) and TMA-producer warp (NCW + 1)
constexpr float LOG2E = 1.4426950408889634f;
constexpr int RBOX = 164;        // backward: width of a positional-logit box (159 needed, + <= 3 because
                                 // the box must start on a 16 B boundary of the fp32 row)

// ---- PTX wrappers (same forms as gemm_tc.cu) ---------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// wait for the n-th completion (n = 0, 1, ...) of a barrier that completes once per round
__device__ __forceinline__ void mbar_wait_nth(uint32_t bar, int n) { mbar_wait(bar, (uint32_t)n & 1u); }

__device__ __forceinline__ void tma_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                       int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0,
                                       int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                     uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
[[maybe_unused]] __device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 512;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// shared-memory matrix descriptor; layout 4 = SWIZZLE_64B (rows of 64 B, 8-row groups 512 B apart)
__device__ __forceinline__ uint64_t desc64(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// descriptor of the same tile `bytes` further on (16 B granular; no carry out of the 14-bit field
// because shared-memory addresses stay below 256 KB)
__device__ __forceinline__ uint64_t dadv(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }
// instruction descriptor: D = f32, A = B = bf16, M = 128
__device__ __forceinline__ uint32_t idesc(int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(QT >> 4) << 24);
}
// byte offset of 8 consecutive elements (16 B chunk `ch` = 0..3) of row `row` in a SW64 tile
__device__ __forceinline__ uint32_t sw64_off(int row, int ch) {
  return (uint32_t)(row * 64 + ((ch ^ ((row >> 1) & 3)) << 4));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// (hi, lo) split of two values, packed pairwise
__device__ __forceinline__ void split_pack(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16(a, b);                                   // one cvt.rn.bf16x2.f32 (a in the low half)
  lo = pack_bf16(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}

struct FusedParams {
  int B, H, T, dh, ndb, W, RW, Tp4;       // Tp4 = round_up(T, 64) / 4: dropout counter pitch
  float scale, drop_p, drop_scale;
  uint32_t drop_thresh;
  uint64_t seed;
  const uint64_t* seed_src;
  uint32_t site;
  // forward
  float* O;          // (B*T, H*dh)
  __nv_bfloat16* O_planes;   // nullable: O also as (2, B*T, H*dh) bf16 hi / lo planes
  float* stat_m;     // (B*H, T) row max of the logits
  float* stat_linv;  // (B*H, T) 1 / sum exp
  // backward
  const float* delta;   // (B*H, T)  sum_d dO * O
  float* dqkv;          // dQ accumulated atomically (pre-zeroed): rows dq_ld apart; with dkv_planes == NULL
                        // the buffer is (B*T, 3*H*dh) and dK / dV are stored into its other thirds
  int64_t dq_ld;
  __nv_bfloat16* dkv_planes;   // nullable: (2, B*T, 3*H*dh) bf16 planes, dK / dV written to their thirds
  __nv_bfloat16* dsb;   // (2, B*T, H, RWp) band-layout dS planes (pre-zeroed)
  int RWp;
};

// =================================================================================================
// forward
// =================================================================================================
struct FwdSmem {   // byte offsets from the 1024-aligned base
  static constexpr int Q = 0;                                         // 2 planes x ndb x 8 KB
  static constexpr int KRING = Q + 2 * MAX_NDB * BLK_BIG;             // 48 KB
  static constexpr int KSTAGE = 2 * MAX_NDB * BLK_SMALL;              // 12 KB
  static constexpr int PHASE1_END = KRING + KST * KSTAGE;             // 120 KB
  // phase 2 aliases the Q / K region once every S MMA has retired
  static constexpr int VRING = 0;
  static constexpr int PT = VRING + VST * KSTAGE;                     // P tiles 0..3 (2 planes x 8 KB each)
  static constexpr int PHASE2_END = PT + 4 * 2 * BLK_BIG;             // 112 KB
  static constexpr int R = PHASE1_END;                                // 128 x RW fp32 (<= 100 KB)
  static constexpr int RED = R + QT * 200 * 4;                        // [2][4][128] fp32 max / sum exchange
  static constexpr int BARS = RED + 2 * 4 * QT * 4;
  static constexpr int TOTAL = BARS + 512;
  static __device__ __forceinline__ int pt(int k) { return PT + k * 2 * BLK_BIG; }
};
static_assert(FwdSmem::PHASE2_END <= FwdSmem::PHASE1_END, "phase-2 buffers must fit the alias");

__global__ void __launch_bounds__(NTHREADS, 1)
attn_fused_fwd_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                      const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapR,
                      const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ssb::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ssb::smem_u32(smem_raw));
  const uint32_t bars = base + FwdSmem::BARS;
  const uint32_t bar_q = bars, bar_r = bars + 8, bar_s = bars + 16, bar_o = bars + 24;
  const uint32_t kfull = bars + 32, kempty = kfull + 8 * KST, vfull = kempty + 8 * KST,
                 vempty = vfull + 8 * VST, pfull = vempty + 8 * VST, pempty = pfull + 32;
  const uint32_t tmem_slot = pempty + 32;
  const uint32_t sready = tmem_slot + 8;   // [11]: S columns of chunk j complete
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / p.H, h = bh - b * p.H;
  const int q0 = blockIdx.x * QT;
  const int ndb = p.ndb;
  // key window of this query tile (start rounded down to a dropout group of 4 keys)
  const int kw0 = max(0, q0 - p.W) & ~7;
  const int kw1 = min(p.T, q0 + QT + p.W);
  const int nch = (kw1 - kw0 + CH - 1) / CH;
  const int plane_q = ndb * BLK_BIG, plane_k = ndb * BLK_SMALL;

  if (threadIdx.x == NCT) {
    for (int i = 0; i < 4; ++i) mbar_init(bars + 8 * i, 1);
    for (int s = 0; s < KST; ++s) { mbar_init(kfull + 8 * s, 1); mbar_init(kempty + 8 * s, 1); }
    for (int s = 0; s < VST; ++s) { mbar_init(vfull + 8 * s, 1); mbar_init(vempty + 8 * s, 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(pfull + 8 * s, NCW); mbar_init(pempty + 8 * s, 1); }
    for (int s = 0; s < 11; ++s) mbar_init(sready + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NCW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  const uint32_t tmem_o = tmem + 352;   // O accumulator columns [352, 352 + dh)

  if (warp == NCW + 1) {
    // ---------------- TMA producer ----------------
    // One lane per (plane, d block): the 2 * ndb copies of an operand tile are issued side by side.
    // (r2 trace: a single thread needed ~1000 cycles to issue the 6 copies of a chunk, which paced
    // both the K stream of phase 1 and the V stream behind the last P tiles.)
    const int pl = lane / ndb, blk = lane - pl * ndb;
    const bool op = lane < 2 * ndb;
    if (lane == 0) mbar_expect_tx(bar_q, 2 * ndb * BLK_BIG);
    __syncwarp();
    if (op) tma_5d(base + FwdSmem::Q + pl * plane_q + blk * BLK_BIG, &mapQ, bar_q, blk * DB, q0, h, b, pl);
    // K ring: [plane][d block][KST x 32 rows][64 B] - the rows of consecutive stages are
    // contiguous inside a (plane, block) slab, so one MMA can take KG chunks as its N = 32 KG
    auto load_k = [&](int j) {
      const int s = j % KST;
      if (lane == 0) mbar_expect_tx(kfull + 8 * s, 2 * ndb * BLK_SMALL);
      __syncwarp();
      if (op)
        tma_5d(base + FwdSmem::KRING + (pl * ndb + blk) * (KST * BLK_SMALL) + s * BLK_SMALL, &mapK,
               kfull + 8 * s, blk * DB, kw0 + j * CH, h, b, pl);
    };
    auto load_v = [&](int j) {
      const int s = j % VST;
      if (lane == 0) mbar_expect_tx(vfull + 8 * s, 2 * ndb * BLK_SMALL);
      __syncwarp();
      if (op)
        tma_5d(base + FwdSmem::VRING + s * FwdSmem::KSTAGE + pl * plane_k + blk * BLK_SMALL, &mapV,
               vfull + 8 * s, blk * DB, kw0 + j * CH, h, b, pl);
    };
    for (int j = 0; j < KST && j < nch; ++j) load_k(j);
    if (lane == 0) {
      mbar_expect_tx(bar_r, QT * p.RW * 4);
      tma_3d(base + FwdSmem::R, &mapR, bar_r, 0, q0, bh);
    }
    for (int j = KST; j < nch; ++j) {      // stage j % KST is free when chunk j - KST's MMAs retired
      mbar_wait(kempty + 8 * (j % KST), (uint32_t)(j / KST - 1) & 1u);
      load_k(j);
    }
    // phase 2: the V ring aliases the Q / K region, dead once every S MMA has retired
    mbar_wait(bar_s, 0);
    for (int j = 0; j < nch; ++j) {
      if (j >= VST) mbar_wait(vempty + 8 * (j % VST), (uint32_t)(j / VST - 1) & 1u);
      load_v(j);
    }
  } else if (warp == NCW) {
    if (ssb::elect_one()) {
      // ---------------- MMA issue ----------------
      // phase 1: S[:, 32 j .. 32 (j + KG)) = Q [K_j ; .. ; K_j+KG-1]^T, KG chunks per MMA.  An N = 32
      // MMA costs ~80 cycles, almost all of it re-reading the 4 KB Q operand from shared memory
      // (r2 trace: 198 of them = 16 k of the CTA's 43 k cycles); N = 96 reads Q a third as often.
      const uint64_t dq_hi = desc64(base + FwdSmem::Q, 16), dq_lo = dadv(dq_hi, plane_q);
      const uint32_t slab_k = KST * BLK_SMALL, plane_kr = ndb * slab_k;
      mbar_wait(bar_q, 0);
      for (int j = 0; j < nch; j += KG) {
        const int s = j % KST, g = min(KG, nch - j);
        for (int u = 0; u < g; ++u) mbar_wait(kfull + 8 * (s + u), (uint32_t)(j / KST) & 1u);
        tc_fence_after();
        const uint32_t id_s = idesc(g * CH, 0, 0);
        const uint64_t dk_hi = desc64(base + FwdSmem::KRING + s * BLK_SMALL, 16);
        const uint64_t dk_lo = dadv(dk_hi, plane_kr);
        const uint32_t d = tmem + (uint32_t)(j * CH);
        uint32_t acc = 0;
        for (int blk = 0; blk < ndb; ++blk)
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t oa = blk * BLK_BIG + ks * 32, ob = blk * slab_k + ks * 32;
            umma(d, dadv(dq_hi, oa), dadv(dk_hi, ob), id_s, acc);
            umma(d, dadv(dq_hi, oa), dadv(dk_lo, ob), id_s, 1u);
            umma(d, dadv(dq_lo, oa), dadv(dk_hi, ob), id_s, 1u);
            acc = 1u;
          }
        for (int u = 0; u < g; ++u) {
          umma_commit(kempty + 8 * (s + u));
          umma_commit(sready + 8 * (j + u));   // pass 1 of the softmax may read these columns
        }
      }
      umma_commit(bar_s);
      // phase 2: O += P_j V_j
      const uint32_t id_o = idesc(p.dh, 0, 1);
      for (int j = 0; j < nch; ++j) {
        const int s = j % VST, pb = j & 3;
        mbar_wait(vfull + 8 * s, (uint32_t)(j / VST) & 1u);
        mbar_wait(pfull + 8 * pb, (uint32_t)(j >> 2) & 1u);
        tc_fence_after();
        const uint64_t dp_hi = desc64(base + FwdSmem::pt(pb), 16), dp_lo = dadv(dp_hi, BLK_BIG);
        // V chunk [32 keys][dh] read MN-major: 16-key k-steps 1 KB apart, 32-wide d groups 2 KB apart
        const uint64_t dv_hi = desc64(base + FwdSmem::VRING + s * FwdSmem::KSTAGE, BLK_SMALL);
        const uint64_t dv_lo = dadv(dv_hi, plane_k);
        for (int ks = 0; ks < 2; ++ks) {
          umma(tmem_o, dadv(dp_hi, ks * 32), dadv(dv_hi, ks * 1024), id_o, (j > 0 || ks > 0) ? 1u : 0u);
          umma(tmem_o, dadv(dp_hi, ks * 32), dadv(dv_lo, ks * 1024), id_o, 1u);
          umma(tmem_o, dadv(dp_lo, ks * 32), dadv(dv_hi, ks * 1024), id_o, 1u);
        }
        umma_commit(pempty + 8 * pb);
        umma_commit(vempty + 8 * s);
      }
      umma_commit(bar_o);
    }
  } else {
    // ---------------- softmax threads ----------------
    // 4 warps per TMEM lane group: thread (lg, lane) owns query row i = 32 lg + lane, warp group
    // cg = warp / 4 takes the key chunks j = cg (mod 4) of that row; row max and row sum are
    // combined through shared memory.
    const int lg = warp & 3, cg = warp >> 2;
    const int i = lg * 32 + lane;         // row of the tile = TMEM lane
    const int q = q0 + i;
    const bool row_ok = q < p.T;
    const int qa = q0 + lg * 32;          // first query row of this warp
    const uint64_t seed = ssb::eff_seed(p.seed, p.seed_src);
    const uint32_t tlane = tmem + ((uint32_t)(lg * 32) << 16);
    const uint32_t rrow = base + FwdSmem::R + (uint32_t)(i * p.RW) * 4u;
    const int soff = kw0 - q + p.W;       // rel = column + soff
    float* red = reinterpret_cast<float*>(gen + FwdSmem::RED);
    // warp-uniform class of a key chunk: 0 = no (row, key) of this warp inside the band,
    // 2 = every (row, key) inside the band and the sequence (no predicates needed), 1 = mixed
    auto chunk_class = [&](int j) -> int {
      const int kc = kw0 + j * CH;
      if (qa >= p.T || kc >= p.T || kc > qa + 31 + p.W || kc + 31 < qa - p.W) return 0;
      if (qa + 31 < p.T && kc + 31 < p.T && kc >= qa + 31 - p.W && kc + 31 <= qa + p.W) return 2;
      return 1;
    };
    mbar_wait(bar_r, 0);
    float m = -CUDART_INF_F;
    for (int j = cg; j < nch; j += 4) {
      const int cls = chunk_class(j);
      if (cls == 0) continue;
      mbar_wait(sready + 8 * j, 0);
      tc_fence_after();
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t v[16];
        tmem_ld16(tlane + (uint32_t)(j * CH + half * 16), v);
        tmem_wait_ld();
        const uint32_t ra = rrow + (uint32_t)(j * CH + half * 16 + soff) * 4u;
        if (cls == 2) {
#pragma unroll
          for (int c = 0; c < 16; ++c)
            m = fmaxf(m, fmaf(__uint_as_float(v[c]), p.scale, ld_shared_f32(ra + c * 4)));
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int col = j * CH + half * 16 + c, rel = col + soff;
            const float y = fmaf(__uint_as_float(v[c]), p.scale, ld_shared_f32(ra + c * 4));
            const bool ok = row_ok && kw0 + col < p.T && rel >= 0 && rel <= 2 * p.W;
            m = fmaxf(m, ok ? y : -CUDART_INF_F);
          }
        }
      }
    }
    red[cg * QT + i] = m;
    bar_compute();
    // the P tiles alias the Q / K buffers: every S MMA must have retired before the first write
    mbar_wait(bar_s, 0);
    m = fmaxf(fmaxf(red[i], red[QT + i]), fmaxf(red[2 * QT + i], red[3 * QT + i]));
    if (!row_ok) m = 0.f;
    const float mneg = -m * LOG2E;
    float sum = 0.f;
    const int64_t drow = ((int64_t)bh * p.T + q) * p.Tp4;
    // Pass 2 walks the chunks IN ORDER with all 16 warps on the same chunk: warp group cg takes
    // the 8 keys [8 cg, 8 cg + 8) of every chunk = one 16 B piece of the row's P operand.  Tiles
    // then complete one after the other and the P V MMAs run alongside the element-wise work
    // (r2 trace: with group cg owning whole chunks cg, cg + 4, .. the first four tiles appeared
    // together after 6 k cycles and the MMAs trailed the last tile by 3.5 k).
    for (int j = 0; j < nch; ++j) {
      const int cls = chunk_class(j);
      const int n = j >> 2;               // n-th use of P tile j % 4
      const uint32_t pt = base + FwdSmem::pt(j & 3);
      const int c0 = j * CH + cg * 8;     // first S column of this thread's piece
      float e[8];
      if (cls == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) e[c] = 0.f;
      } else {
        uint32_t v[8];
        tmem_ld8(tlane + (uint32_t)c0, v);
        tmem_wait_ld();
        const uint32_t ra = rrow + (uint32_t)(c0 + soff) * 4u;
        if (cls == 2) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float y = fmaf(__uint_as_float(v[c]), p.scale, ld_shared_f32(ra + c * 4));
            e[c] = ex2(fmaf(y, LOG2E, mneg));
            sum += e[c];
          }
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const int col = c0 + c, rel = col + soff;
            const float y = fmaf(__uint_as_float(v[c]), p.scale, ld_shared_f32(ra + c * 4));
            const bool ok = row_ok && kw0 + col < p.T && rel >= 0 && rel <= 2 * p.W;
            e[c] = ok ? ex2(fmaf(y, LOG2E, mneg)) : 0.f;
            sum += e[c];
          }
        }
        if (p.drop_p > 0.f) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint4 rnd = ssb::dropout_bits4(seed, p.site, (uint64_t)(drow + ((kw0 + c0) >> 2) + g));
            e[4 * g + 0] = rnd.x >= p.drop_thresh ? e[4 * g + 0] * p.drop_scale : 0.f;
            e[4 * g + 1] = rnd.y >= p.drop_thresh ? e[4 * g + 1] * p.drop_scale : 0.f;
            e[4 * g + 2] = rnd.z >= p.drop_thresh ? e[4 * g + 2] * p.drop_scale : 0.f;
            e[4 * g + 3] = rnd.w >= p.drop_thresh ? e[4 * g + 3] * p.drop_scale : 0.f;
          }
        }
      }
      uint4 hi, lo;
      split_pack(e[0], e[1], hi.x, lo.x);
      split_pack(e[2], e[3], hi.y, lo.y);
      split_pack(e[4], e[5], hi.z, lo.z);
      split_pack(e[6], e[7], hi.w, lo.w);
      // only now is the tile needed: the MMAs of chunk j-4 (its previous user) must have retired
      if (n >= 1) mbar_wait(pempty + 8 * (j & 3), (uint32_t)(n - 1) & 1u);
      const uint32_t off = sw64_off(i, cg);
      st_shared_v4(pt + off, hi);
      st_shared_v4(pt + BLK_BIG + off, lo);
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pfull + 8 * (j & 3));
    }
    red[4 * QT + cg * QT + i] = sum;
    bar_compute();
    sum = (red[4 * QT + i] + red[5 * QT + i]) + (red[6 * QT + i] + red[7 * QT + i]);
    // epilogue: O = acc / sum; warp group cg stores columns [cg dh/4, (cg+1) dh/4)
    mbar_wait(bar_o, 0);
    tc_fence_after();
    const float linv = row_ok ? 1.f / sum : 0.f;
    if (row_ok && cg == 0) {
      p.stat_m[(int64_t)bh * p.T + q] = m;
      p.stat_linv[(int64_t)bh * p.T + q] = linv;
    }
    // The tile leaves through shared memory (every MMA has retired: the V ring / P tiles are
    // free): each thread parks its row's columns, then the CTA writes whole 384 B rows with
    // consecutive lanes on consecutive 16 B.  (r2 trace: one thread per row storing straight to
    // its own row of O took 5.9 k of the CTA's 43 k cycles - 32 rows per store instruction.)
    constexpr uint32_t OSTRIDE = (DB * MAX_NDB + 4) * 4;       // 400 B: 8 lanes cover 128 B
    const uint32_t ost = base + FwdSmem::VRING;
    const int cw = p.dh >> 2;
    for (int c0 = cg * cw; c0 < (cg + 1) * cw; c0 += 8) {
      uint32_t v[8];
      tmem_ld8(tlane + 352 + (uint32_t)c0, v);
      tmem_wait_ld();
      const uint32_t dst = ost + (uint32_t)i * OSTRIDE + (uint32_t)c0 * 4u;
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(__uint_as_float(v[0]) * linv),
                   "f"(__uint_as_float(v[1]) * linv), "f"(__uint_as_float(v[2]) * linv),
                   "f"(__uint_as_float(v[3]) * linv)
                   : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + 16u), "f"(__uint_as_float(v[4]) * linv),
                   "f"(__uint_as_float(v[5]) * linv), "f"(__uint_as_float(v[6]) * linv),
                   "f"(__uint_as_float(v[7]) * linv)
                   : "memory");
    }
    bar_compute();
    const int nv4 = p.dh >> 2;
    float* otile = p.O + ((int64_t)b * p.T + q0) * (p.H * p.dh) + h * p.dh;
    const int rows_ok = min(QT, p.T - q0);
    for (int f = threadIdx.x; f < rows_ok * nv4; f += NCT) {
      const int row = f / nv4, c4 = f - row * nv4;
      float4 o;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                   : "r"(ost + (uint32_t)row * OSTRIDE + (uint32_t)c4 * 16u)
                   : "memory");
      *reinterpret_cast<float4*>(otile + (int64_t)row * (p.H * p.dh) + c4 * 4) = o;
      if (p.O_planes) {   // the out-projection's operand, written here instead of by a split pass
        uint2 hi, lo;
        split_pack(o.x, o.y, hi.x, lo.x);
        split_pack(o.z, o.w, hi.y, lo.y);
        __nv_bfloat16* dst = p.O_planes + ((int64_t)b * p.T + q0 + row) * (p.H * p.dh) + h * p.dh + c4 * 4;
        *reinterpret_cast<uint2*>(dst) = hi;
        *reinterpret_cast<uint2*>(dst + (int64_t)p.B * p.T * p.H * p.dh) = lo;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NCW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// =================================================================================================
// backward
// =================================================================================================
struct BwdSmem {
  static constexpr int K = 0;                                        // 2 planes x ndb x 8 KB
  static constexpr int V = K + 2 * MAX_NDB * BLK_BIG;                // 48 KB
  static constexpr int QD = V + 2 * MAX_NDB * BLK_BIG;               // 96 KB: 2 x {Q_c, dO_c}
  static constexpr int QD_HALF = 2 * MAX_NDB * BLK_SMALL;            // 12 KB (Q_c or dO_c)
  static constexpr int QD_STAGE = 2 * QD_HALF;                       // 24 KB
  static constexpr int PD = QD + 2 * QD_STAGE;                       // 144 KB: P_drop^T planes
  static constexpr int DS = PD + 2 * BLK_BIG;                        // 160 KB: dS^T planes
  static constexpr int R = DS + 2 * BLK_BIG;                         // 176 KB: 2 x [32][164] fp32
  static constexpr int R_STAGE = CH * RBOX * 4;                      // 20 KB
  static constexpr int STATS = R + 2 * R_STAGE;                      // 216 KB: 2 x 3 x 32 fp32
  static constexpr int BARS = STATS + 2 * 3 * CH * 4;
  static constexpr int TOTAL = BARS + 256;
};

__global__ void __launch_bounds__(NTHREADS, 1)
attn_fused_bwd_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                      const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapDO,
                      const __grid_constant__ CUtensorMap mapR, const FusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ssb::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ssb::smem_u32(smem_raw));
  const uint32_t bars = base + BwdSmem::BARS;
  const uint32_t kv_full = bars, qd_full = bars + 8 /*[2]*/, r_full = bars + 24 /*[2]*/,
                 r_free = bars + 40 /*[2]*/, s_full = bars + 56 /*[2]*/, sp_free = bars + 72 /*[2]*/,
                 tiles_full = bars + 88, mma2_done = bars + 96, ep_done = bars + 104 /*[2]*/,
                 qd_free = bars + 120 /*[2]*/;
  const uint32_t tmem_slot = bars + 136;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y, b = bh / p.H, h = bh - b * p.H;
  const int k0 = blockIdx.x * QT;
  const int ndb = p.ndb;
  const int qw0 = max(0, k0 - p.W);
  const int qw1 = min(p.T, k0 + QT + p.W);
  const int nch = (qw1 - qw0 + CH - 1) / CH;
  const int plane_big = ndb * BLK_BIG, plane_small = ndb * BLK_SMALL;

  if (threadIdx.x == NCT) {
    mbar_init(kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(qd_full + 8 * s, 1); mbar_init(r_full + 8 * s, 1); mbar_init(r_free + 8 * s, NCW);
      mbar_init(s_full + 8 * s, 1); mbar_init(sp_free + 8 * s, NCW); mbar_init(ep_done + 8 * s, NCW);
      mbar_init(qd_free + 8 * s, 1);
    }
    mbar_init(tiles_full, NCW);
    mbar_init(mma2_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == NCW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  // TMEM columns: [0,128) two {S^T, dP^T} pairs, [128,256) dV, [256,384) dK, [384,448) two dQ^T
  const uint32_t t_dv = tmem + 128, t_dk = tmem + 256, t_dq = tmem + 384;

  if (warp == NCW + 1) {
    // ---------------- TMA producer ----------------
    // One lane per (tensor, plane, d block) copy, issued side by side (see the forward kernel).
    const int which = lane / (2 * ndb), rem = lane - which * 2 * ndb;   // 0: K / Q_c, 1: V / dO_c
    const int pl = rem / ndb, blk = rem - pl * ndb;
    const bool op = lane < 4 * ndb;
    if (lane == 0) mbar_expect_tx(kv_full, 4 * ndb * BLK_BIG);
    __syncwarp();
    if (op)
      tma_5d(base + (which ? BwdSmem::V : BwdSmem::K) + pl * plane_big + blk * BLK_BIG,
             which ? &mapV : &mapK, kv_full, blk * DB, k0, h, b, pl);
    auto load_qd = [&](int j) {
      const int s = j & 1;
      const uint32_t dst = base + BwdSmem::QD + s * BwdSmem::QD_STAGE;
      if (lane == 0) mbar_expect_tx(qd_full + 8 * s, 4 * ndb * BLK_SMALL);
      __syncwarp();
      if (op)
        tma_5d(dst + which * BwdSmem::QD_HALF + pl * plane_small + blk * BLK_SMALL,
               which ? &mapDO : &mapQ, qd_full + 8 * s, blk * DB, qw0 + j * CH, h, b, pl);
    };
    auto load_r = [&](int j) {
      const int s = j & 1;
      if (lane == 0) {
        mbar_expect_tx(r_full + 8 * s, BwdSmem::R_STAGE);
        // columns rel0 .. rel0 + 158 of rows q_c .. q_c + 31 (start rounded down to 4 floats: TMA
        // faults on an inner coordinate that is not 16 B aligned); out-of-range parts are zero-filled
        const int rel0 = k0 - (qw0 + j * CH) - (CH - 1) + p.W;
        tma_3d(base + BwdSmem::R + s * BwdSmem::R_STAGE, &mapR, r_full + 8 * s, rel0 & ~3,
               qw0 + j * CH, bh);
      }
    };
    load_qd(0);
    load_r(0);
    if (nch > 1) { load_qd(1); load_r(1); }
    for (int j = 0; j + 2 < nch; ++j) {
      const int s = j & 1;
      mbar_wait(r_free + 8 * s, (uint32_t)(j >> 1) & 1u);     // threads are done with R stage s
      load_r(j + 2);
      mbar_wait(qd_free + 8 * s, (uint32_t)(j >> 1) & 1u);    // dV / dK MMAs of chunk j retired
      load_qd(j + 2);
    }
  } else if (warp == NCW) {
    if (ssb::elect_one()) {
      // ---------------- MMA issue ----------------
      const uint32_t id_s = idesc(CH, 0, 0), id_acc = idesc(p.dh, 0, 1), id_dq = idesc(CH, 1, 1);
      // K / V tiles as K-major A operands (S^T, dP^T) and K as the MN-major A operand of dQ^T
      const uint64_t k_hi = desc64(base + BwdSmem::K, 16), k_lo = dadv(k_hi, plane_big);
      const uint64_t v_hi = desc64(base + BwdSmem::V, 16), v_lo = dadv(v_hi, plane_big);
      const uint64_t kt_hi = desc64(base + BwdSmem::K, BLK_BIG), kt_lo = dadv(kt_hi, plane_big);
      const uint64_t pd_hi = desc64(base + BwdSmem::PD, 16), pd_lo = dadv(pd_hi, BLK_BIG);
      const uint64_t ds_hi = desc64(base + BwdSmem::DS, 16), ds_lo = dadv(ds_hi, BLK_BIG);
      const uint64_t dst_hi = desc64(base + BwdSmem::DS, BLK_BIG), dst_lo = dadv(dst_hi, BLK_BIG);
      auto mma1 = [&](int j) {
        const int s = j & 1;
        mbar_wait(qd_full + 8 * s, (uint32_t)(j >> 1) & 1u);
        if (j >= 2) mbar_wait(sp_free + 8 * s, (uint32_t)((j - 2) >> 1) & 1u);
        tc_fence_after();
        const uint64_t q_hi = desc64(base + BwdSmem::QD + s * BwdSmem::QD_STAGE, 16);
        const uint64_t q_lo = dadv(q_hi, plane_small);
        const uint64_t o_hi = dadv(q_hi, BwdSmem::QD_HALF), o_lo = dadv(o_hi, plane_small);
        const uint32_t d_s = tmem + s * 64, d_p = d_s + 32;
        uint32_t acc = 0;
        for (int blk = 0; blk < ndb; ++blk)
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t oa = blk * BLK_BIG + ks * 32, ob = blk * BLK_SMALL + ks * 32;
            umma(d_s, dadv(k_hi, oa), dadv(q_hi, ob), id_s, acc);
            umma(d_s, dadv(k_hi, oa), dadv(q_lo, ob), id_s, 1u);
            umma(d_s, dadv(k_lo, oa), dadv(q_hi, ob), id_s, 1u);
            umma(d_p, dadv(v_hi, oa), dadv(o_hi, ob), id_s, acc);
            umma(d_p, dadv(v_hi, oa), dadv(o_lo, ob), id_s, 1u);
            umma(d_p, dadv(v_lo, oa), dadv(o_hi, ob), id_s, 1u);
            acc = 1u;
          }
        umma_commit(s_full + 8 * s);
      };
      mbar_wait(kv_full, 0);
      mma1(0);
      for (int j = 0; j < nch; ++j) {
        if (j + 1 < nch) mma1(j + 1);
        // second group of chunk j: needs its P_drop^T / dS^T tiles and a free dQ^T accumulator
        const int s = j & 1;
        mbar_wait_nth(tiles_full, j);
        tc_fence_after();
        // Q_c / dO_c read MN-major (rows = queries): 16-row k-steps 1 KB apart, d groups 2 KB apart
        const uint64_t qm_hi = desc64(base + BwdSmem::QD + s * BwdSmem::QD_STAGE, BLK_SMALL);
        const uint64_t qm_lo = dadv(qm_hi, plane_small);
        const uint64_t om_hi = dadv(qm_hi, BwdSmem::QD_HALF), om_lo = dadv(om_hi, plane_small);
        for (int ks = 0; ks < 2; ++ks) {   // K = 32 queries
          const uint32_t oa = ks * 32, ob = ks * 1024;
          const uint32_t acc = (j > 0 || ks > 0) ? 1u : 0u;
          umma(t_dv, dadv(pd_hi, oa), dadv(om_hi, ob), id_acc, acc);
          umma(t_dv, dadv(pd_hi, oa), dadv(om_lo, ob), id_acc, 1u);
          umma(t_dv, dadv(pd_lo, oa), dadv(om_hi, ob), id_acc, 1u);
          umma(t_dk, dadv(ds_hi, oa), dadv(qm_hi, ob), id_acc, acc);
          umma(t_dk, dadv(ds_hi, oa), dadv(qm_lo, ob), id_acc, 1u);
          umma(t_dk, dadv(ds_lo, oa), dadv(qm_hi, ob), id_acc, 1u);
        }
        umma_commit(qd_free + 8 * s);       // the {Q_c, dO_c} stage may be refilled
        // dQ_c^T [d][q] = K^T dS^T: both operands MN-major (rows = keys), K = 128 keys
        if (j >= 2) {
          mbar_wait(ep_done + 8 * s, (uint32_t)((j - 2) >> 1) & 1u);
          tc_fence_after();
        }
        const uint32_t d_q = t_dq + s * 32;
        for (int ks = 0; ks < 8; ++ks) {
          const uint32_t o = ks * 1024;
          umma(d_q, dadv(kt_hi, o), dadv(dst_hi, o), id_dq, ks > 0 ? 1u : 0u);
          umma(d_q, dadv(kt_hi, o), dadv(dst_lo, o), id_dq, 1u);
          umma(d_q, dadv(kt_lo, o), dadv(dst_hi, o), id_dq, 1u);
        }
        umma_commit(mma2_done);
      }
    }
  } else {
    // ---------------- element-wise threads ----------------
    // 4 warps per TMEM lane group: thread (lg, lane) owns key row i = 32 lg + lane, warp group
    // cg = warp / 4 owns query columns [8 cg, 8 cg + 8) of every 32-query chunk.
    const int lg = warp & 3, cg = warp >> 2;
    const int i = lg * 32 + lane;
    const int k = k0 + i;
    const bool key_ok = k < p.T;
    const int ka = k0 + lg * 32;          // first key of this warp
    const int c8 = cg * 8;
    const int quad = lane & 3;
    const uint64_t seed = ssb::eff_seed(p.seed, p.seed_src);
    const uint32_t tlane = tmem + ((uint32_t)(lg * 32) << 16);
    const int D = p.H * p.dh;
    const int64_t band_plane = (int64_t)p.B * p.T * p.H * p.RWp;

    // dQ_c^T: lane = head-dim index, column = query of chunk j.  `wait` only for the last chunk:
    // earlier ones are called after this thread already observed completion j of mma2_done (a
    // second wait on the same phase could miss it once the barrier has moved two phases on).
    auto dq_epilogue = [&](int j, bool wait) {
      const int s = j & 1;
      if (wait) mbar_wait_nth(mma2_done, j);
      tc_fence_after();
      uint32_t v[8];
      tmem_ld8(tlane + 384 + (uint32_t)(s * 32 + c8), v);
      tmem_wait_ld();
      if (i < p.dh) {
        const int qb = qw0 + j * CH + c8;
        float* dst = p.dqkv + ((int64_t)b * p.T + qb) * p.dq_ld + h * p.dh + i;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (qb + c < p.T) atomicAdd(dst + (int64_t)c * p.dq_ld, __uint_as_float(v[c]) * p.scale);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ep_done + 8 * s);
    };

    for (int j = 0; j < nch; ++j) {
      const int s = j & 1;
      const int qc = qw0 + j * CH;
      // row statistics of the chunk's queries -> shared memory (broadcast reads below)
      float* st = reinterpret_cast<float*>(gen + BwdSmem::STATS) + s * 3 * CH;
      if (threadIdx.x < 3 * CH) {
        const int which = threadIdx.x / CH, c = threadIdx.x - which * CH, q = qc + c;
        const float* src = which == 0 ? p.stat_m : which == 1 ? p.stat_linv : p.delta;
        st[threadIdx.x] = q < p.T ? __ldg(src + (int64_t)bh * p.T + q) : 0.f;
      }
      bar_compute();
      mbar_wait(r_full + 8 * s, (uint32_t)(j >> 1) & 1u);
      mbar_wait(s_full + 8 * s, (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      uint32_t sv[8], dv[8];
      tmem_ld8(tlane + (uint32_t)(s * 64 + c8), sv);
      tmem_ld8(tlane + (uint32_t)(s * 64 + 32 + c8), dv);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sp_free + 8 * s);
      // warp-uniform: does any (key of this warp, query of these 8 columns) fall inside the band?
      const int qs = qc + c8;
      const bool live = ka < p.T && qs < p.T && qs <= ka + 31 + p.W && qs + 7 >= ka - p.W;
      // positional logits: box column = rel - (rel0 & ~3) = i - c + 31 + (rel0 & 3), row c
      const int rsh = (k0 - qc - (CH - 1) + p.W) & 3;
      const uint32_t rbase =
          base + BwdSmem::R + s * BwdSmem::R_STAGE + (uint32_t)(i + CH - 1 + rsh) * 4u;
      float pdv[8], dsv[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) { pdv[c] = 0.f; dsv[c] = 0.f; }
      if (live) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          // dropout words: one Philox call covers 4 consecutive keys (this quad) of one query;
          // lane `quad` draws for column 4g + quad, then the 4 x 4 block is transposed in the quad
          uint32_t w[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
          if (p.drop_p > 0.f) {
            const int qd = qs + 4 * g + quad;
            const uint4 rnd = ssb::dropout_bits4(
                seed, p.site, (uint64_t)(((int64_t)bh * p.T + qd) * p.Tp4 + (k >> 2)));
            w[0] = rnd.x; w[1] = rnd.y; w[2] = rnd.z; w[3] = rnd.w;
            uint32_t x0 = (quad & 2) ? w[0] : w[2], x1 = (quad & 2) ? w[1] : w[3];
            x0 = __shfl_xor_sync(0xffffffffu, x0, 2);
            x1 = __shfl_xor_sync(0xffffffffu, x1, 2);
            if (quad & 2) { w[0] = x0; w[1] = x1; } else { w[2] = x0; w[3] = x1; }
            uint32_t y0 = (quad & 1) ? w[0] : w[1], y1 = (quad & 1) ? w[2] : w[3];
            y0 = __shfl_xor_sync(0xffffffffu, y0, 1);
            y1 = __shfl_xor_sync(0xffffffffu, y1, 1);
            if (quad & 1) { w[0] = y0; w[2] = y1; } else { w[1] = y0; w[3] = y1; }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int cc = 4 * g + u, c = c8 + cc, q = qc + c, rel = k - q + p.W;
            const bool inb = key_ok && q < p.T && rel >= 0 && rel <= 2 * p.W;
            const float y = fmaf(__uint_as_float(sv[cc]), p.scale,
                                 ld_shared_f32(rbase + (uint32_t)(c * (RBOX - 1)) * 4u));
            const float pr = inb ? ex2((y - st[c]) * LOG2E) * st[CH + c] : 0.f;
            const bool keep = w[u] >= p.drop_thresh;
            const float dpm = (keep && inb) ? __uint_as_float(dv[cc]) * p.drop_scale : 0.f;
            pdv[cc] = keep ? pr * p.drop_scale : 0.f;
            dsv[cc] = pr * (dpm - st[2 * CH + c]);
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(r_free + 8 * s);
      // band-layout dS (unscaled) for the positional part of dQ: consecutive lanes = consecutive rel
      if (live && key_ok) {
        __nv_bfloat16* brow = p.dsb + (((int64_t)b * p.T + qs) * p.H + h) * p.RWp + (k - qs + p.W);
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const int q = qs + cc, rel = k - q + p.W;
          if (q < p.T && rel >= 0 && rel <= 2 * p.W) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(dsv[cc]);
            __nv_bfloat16* o = brow + (int64_t)cc * (p.H * p.RWp - 1);
            o[0] = hi;
            o[band_plane] = __float2bfloat16_rn(dsv[cc] - __bfloat162float(hi));
          }
        }
      }
      // the tiles are free once the second MMA group of chunk j-1 has retired
      if (j >= 1) mbar_wait_nth(mma2_done, j - 1);
      {
        uint4 hi, lo;
        const uint32_t off = sw64_off(i, cg);
        split_pack(pdv[0], pdv[1], hi.x, lo.x);
        split_pack(pdv[2], pdv[3], hi.y, lo.y);
        split_pack(pdv[4], pdv[5], hi.z, lo.z);
        split_pack(pdv[6], pdv[7], hi.w, lo.w);
        st_shared_v4(base + BwdSmem::PD + off, hi);
        st_shared_v4(base + BwdSmem::PD + BLK_BIG + off, lo);
        split_pack(dsv[0], dsv[1], hi.x, lo.x);
        split_pack(dsv[2], dsv[3], hi.y, lo.y);
        split_pack(dsv[4], dsv[5], hi.z, lo.z);
        split_pack(dsv[6], dsv[7], hi.w, lo.w);
        st_shared_v4(base + BwdSmem::DS + off, hi);
        st_shared_v4(base + BwdSmem::DS + BLK_BIG + off, lo);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(tiles_full);
      if (j >= 1) dq_epilogue(j - 1, false);
    }
    dq_epilogue(nch - 1, true);   // also: every MMA of the tile has retired -> dV / dK are complete
    // dK / dV leave through shared memory like the forward's O tile (every MMA has retired: the
    // K / V / {Q_c, dO_c} buffers are free): rows are parked by their owning threads, then written
    // out 384 B at a time, as fp32 into dqkv or directly as the operand planes of the QKV
    // data-gradient GEMM.
    constexpr uint32_t OSTRIDE = (DB * MAX_NDB + 4) * 4;       // 400 B: 8 lanes cover 128 B
    constexpr uint32_t OTILE = QT * OSTRIDE;
    static_assert(2 * OTILE <= BwdSmem::PD, "dK / dV staging must fit the K / V / QD buffers");
    const uint32_t ost = base + BwdSmem::K;
    const int cw = p.dh >> 2;
    for (int c0 = cg * cw; c0 < (cg + 1) * cw; c0 += 8) {
      uint32_t a[8], g[8];
      tmem_ld8(tlane + 256 + (uint32_t)c0, a);
      tmem_ld8(tlane + 128 + (uint32_t)c0, g);
      tmem_wait_ld();
      const uint32_t dst = ost + (uint32_t)i * OSTRIDE + (uint32_t)c0 * 4u;
#pragma unroll
      for (int c = 0; c < 8; c += 4) {
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + c * 4u),
                     "f"(__uint_as_float(a[c]) * p.scale), "f"(__uint_as_float(a[c + 1]) * p.scale),
                     "f"(__uint_as_float(a[c + 2]) * p.scale), "f"(__uint_as_float(a[c + 3]) * p.scale)
                     : "memory");
        asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst + OTILE + c * 4u),
                     "f"(__uint_as_float(g[c])), "f"(__uint_as_float(g[c + 1])),
                     "f"(__uint_as_float(g[c + 2])), "f"(__uint_as_float(g[c + 3]))
                     : "memory");
      }
    }
    bar_compute();
    const int nv4 = p.dh >> 2;
    const int rows_ok = min(QT, p.T - k0);
    const int per = rows_ok * nv4;
    const int64_t plane = (int64_t)p.B * p.T * 3 * D;
    for (int f = threadIdx.x; f < 2 * per; f += NCT) {
      const int which = f >= per ? 1 : 0;              // 0: dK, 1: dV
      const int r = f - which * per, row = r / nv4, c4 = r - row * nv4;
      float4 o;
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                   : "r"(ost + which * OTILE + (uint32_t)row * OSTRIDE + (uint32_t)c4 * 16u)
                   : "memory");
      const int64_t e = ((int64_t)b * p.T + k0 + row) * (3 * D) + (1 + which) * D + h * p.dh + c4 * 4;
      if (p.dkv_planes) {
        uint2 hi, lo;
        split_pack(o.x, o.y, hi.x, lo.x);
        split_pack(o.z, o.w, hi.y, lo.y);
        *reinterpret_cast<uint2*>(p.dkv_planes + e) = hi;
        *reinterpret_cast<uint2*>(p.dkv_planes + plane + e) = lo;
      } else {
        *reinterpret_cast<float4*>(p.dqkv + e) = o;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NCW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// delta[bh, q] = sum_d dO[b*T+q, h*dh+d] * O[...]: one warp per ROW (all heads), float4 loads with
// every load of the row in flight at once; head sums by a masked warp reduction per head.
// (One warp per (row, head) with scalar loads ran at 2.3 TB/s of the 98 MB it reads.)
constexpr int DELTA_MAXV = 8;   // float4 per lane: H * dh <= 1024
__global__ void __launch_bounds__(256)
attn_delta_kernel(const float* __restrict__ O, const float* __restrict__ dO, int64_t rows, int T, int H,
                  int dh, float* __restrict__ delta, float* __restrict__ zero_out, int64_t zero_ld) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int D = H * dh, nv = D >> 2, hv = dh >> 2;   // float4 per row / per head
  if (zero_out) {   // the content-dQ accumulator the fused backward adds into (red.global.add)
    float4* z = reinterpret_cast<float4*>(zero_out + row * zero_ld);
    for (int v = lane; v < nv; v += 32) z[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float4* a = reinterpret_cast<const float4*>(O + row * (int64_t)D);
  const float4* g = reinterpret_cast<const float4*>(dO + row * (int64_t)D);
  float part[DELTA_MAXV];
#pragma unroll
  for (int i = 0; i < DELTA_MAXV; ++i) {
    const int v = lane + 32 * i;
    part[i] = 0.f;
    if (v < nv) {
      const float4 x = __ldg(a + v), y = __ldg(g + v);
      part[i] = x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
    }
  }
  const int64_t bb = row / T, q = row - bb * T;
  for (int h = 0; h < H; ++h) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < DELTA_MAXV; ++i) {
      const int v = lane + 32 * i;
      acc += (v / hv == h) ? part[i] : 0.f;
    }
    acc = ssb::warp_sum(acc);
    if (lane == 0) delta[(bb * H + h) * T + q] = acc;
  }
}

// scalar fallback for head dims that are not a multiple of 4 or rows wider than 1024 columns
__global__ void __launch_bounds__(256)
attn_delta_scalar_kernel(const float* __restrict__ O, const float* __restrict__ dO, int64_t rows, int T,
                         int H, int dh, float* __restrict__ delta) {
  const int lane = threadIdx.x & 31;
  const int64_t w = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (w >= rows * H) return;
  const int64_t row = w / H;
  const int h = (int)(w - row * H);
  const float* a = O + row * (int64_t)(H * dh) + h * dh;
  const float* g = dO + row * (int64_t)(H * dh) + h * dh;
  float acc = 0.f;
  for (int d = lane; d < dh; d += 32) acc += __ldg(a + d) * __ldg(g + d);
  acc = ssb::warp_sum(acc);
  if (lane == 0) {
    const int64_t bb = row / T, q = row - bb * T;
    delta[(bb * H + h) * T + q] = acc;
  }
}

// ---- tensor maps ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_enc = nullptr;
std::mutex g_enc_mutex;

int get_enc(EncodeTiledFn* out) {
  std::lock_guard<std::mutex> lk(g_enc_mutex);
  if (!g_enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SSB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    SSB_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess,
                "attn_fused: cuTensorMapEncodeTiled not available from the driver");
    g_enc = (EncodeTiledFn)fn;
  }
  *out = g_enc;
  return SSB_OK;
}

// head planes (2, B*T, G, hs) bf16, hs = head stride in elements (128: heads zero-padded to 128
// columns; dh: packed, i.e. the plain split planes of a (B*T, G*dh) matrix as a GEMM epilogue
// writes them).  dims (d, t, head, batch, plane); box (32 d, box_rows t): only the first dh
// columns of a head are ever addressed.
int head_map(CUtensorMap* map, const void* planes, int64_t B, int64_t T, int64_t G, int64_t g0,
             int64_t H, int box_rows, int64_t hs, int64_t dh) {
  EncodeTiledFn enc = nullptr;
  if (int rc = get_enc(&enc)) return rc;
  SSB_REQUIRE(hs >= dh && hs % 8 == 0, "attn_fused: head stride %lld (dh %lld) must be a multiple of 8",
              (long long)hs, (long long)dh);
  const __nv_bfloat16* basep = (const __nv_bfloat16*)planes + g0 * hs;
  SSB_REQUIRE(((uintptr_t)basep & 15) == 0, "attn_fused: planes must be 16 B aligned");
  const int64_t ld = G * hs;
  cuuint64_t dims[5] = {(cuuint64_t)dh, (cuuint64_t)T, (cuuint64_t)H, (cuuint64_t)B, 2};
  cuuint64_t strides[4] = {(cuuint64_t)ld * 2, (cuuint64_t)hs * 2, (cuuint64_t)(T * ld) * 2,
                           (cuuint64_t)(B * T * ld) * 2};
  cuuint32_t box[5] = {DB, (cuuint32_t)box_rows, 1, 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, (void*)basep, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SSB_REQUIRE(r == CUDA_SUCCESS, "attn_fused: cuTensorMapEncodeTiled (heads) failed (%d)", (int)r);
  return SSB_OK;
}

// R (B*H, T, RW) fp32: dims (rel, t, bh); box (box_w, box_rows, 1), no swizzle
int r_map(CUtensorMap* map, const float* R, int64_t BH, int64_t T, int64_t RW, int box_w, int box_rows) {
  EncodeTiledFn enc = nullptr;
  if (int rc = get_enc(&enc)) return rc;
  SSB_REQUIRE(((uintptr_t)R & 15) == 0 && RW % 4 == 0, "attn_fused: R must be 16 B aligned, RW %% 4 == 0");
  cuuint64_t dims[3] = {(cuuint64_t)RW, (cuuint64_t)T, (cuuint64_t)BH};
  cuuint64_t strides[2] = {(cuuint64_t)RW * 4, (cuuint64_t)(T * RW) * 4};
  cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)R, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SSB_REQUIRE(r == CUDA_SUCCESS, "attn_fused: cuTensorMapEncodeTiled (R) failed (%d)", (int)r);
  return SSB_OK;
}

int fill(FusedParams* p, int64_t B, int64_t T, int64_t H, int64_t dh, int64_t W, int64_t RW,
         float drop_p, uint64_t seed, uint32_t site) {
  SSB_REQUIRE(B >= 1 && H >= 1 && T >= 1 && B * H <= 65535, "attn_fused: bad B / H / T");
  SSB_REQUIRE(dh % DB == 0 && dh >= DB && dh <= DB * MAX_NDB, "attn_fused: head dim %lld not in {32, 64, 96}",
              (long long)dh);
  SSB_REQUIRE(W >= 0 && W <= 99 && RW >= 2 * W + 1 && RW <= 200 && RW % 4 == 0,
              "attn_fused: band W=%lld RW=%lld unsupported (W <= 99, RW <= 200)", (long long)W, (long long)RW);
  SSB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "attn_fused: bad dropout p");
  p->B = (int)B; p->H = (int)H; p->T = (int)T; p->dh = (int)dh; p->ndb = (int)(dh / DB);
  p->W = (int)W; p->RW = (int)RW;
  p->Tp4 = (int)(((T + 63) / 64 * 64) / 4);
  p->scale = 1.0f / sqrtf((float)dh);
  p->drop_p = drop_p; p->drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const double th = (double)drop_p * 4294967296.0;
  p->drop_thresh = th >= 4294967295.0 ? 0xffffffffu : (uint32_t)th;
  p->seed = seed; p->seed_src = ssb::seed_source(); p->site = site;
  return SSB_OK;
}

template <typename K>
int set_smem(K kernel, int bytes) {
  SSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  return SSB_OK;
}
constexpr int FWD_SMEM_BYTES = FwdSmem::TOTAL + 1024;
constexpr int BWD_SMEM_BYTES = BwdSmem::TOTAL + 1024;
static_assert(FWD_SMEM_BYTES <= 232448 && BWD_SMEM_BYTES <= 232448, "shared memory budget");

}  // namespace

extern "C" {

int ssb_attn_fused_fwd(const void* qkv_planes, const float* R, int64_t B, int64_t T, int64_t H,
                       int64_t dh, int64_t W, int64_t RW, float drop_p, uint64_t seed, uint32_t site,
                       float* O, void* O_planes, float* stat_m, float* stat_linv, int64_t head_stride,
                       void* stream) {
  const int64_t hs = head_stride > 0 ? head_stride : 128;
  FusedParams p = {};
  if (int rc = fill(&p, B, T, H, dh, W, RW, drop_p, seed, site)) return rc;
  SSB_REQUIRE(qkv_planes && R && O && stat_m && stat_linv, "attn_fused_fwd: null pointer");
  SSB_REQUIRE(((uintptr_t)O & 15) == 0, "attn_fused_fwd: O must be 16 B aligned");
  SSB_REQUIRE(!O_planes || (((uintptr_t)O_planes & 7) == 0 && dh % 4 == 0), "attn_fused_fwd: O_planes alignment");
  p.O = O; p.O_planes = (__nv_bfloat16*)O_planes; p.stat_m = stat_m; p.stat_linv = stat_linv;
  CUtensorMap mq, mk, mv, mr;
  if (int rc = head_map(&mq, qkv_planes, B, T, 3 * H, 0, H, QT, hs, dh)) return rc;
  if (int rc = head_map(&mk, qkv_planes, B, T, 3 * H, H, H, CH, hs, dh)) return rc;
  if (int rc = head_map(&mv, qkv_planes, B, T, 3 * H, 2 * H, H, CH, hs, dh)) return rc;
  if (int rc = r_map(&mr, R, B * H, T, RW, (int)RW, QT)) return rc;
  const int smem = FWD_SMEM_BYTES;
  if (int rc = set_smem(attn_fused_fwd_kernel, smem)) return rc;
  dim3 grid((unsigned)((T + QT - 1) / QT), (unsigned)(B * H));
  attn_fused_fwd_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(mq, mk, mv, mr, p);
  SSB_LAUNCH_CHECK("attn_fused_fwd");
  return SSB_OK;
}

int ssb_attn_delta(const float* O, const float* dO, int64_t B, int64_t T, int64_t H, int64_t dh,
                   float* delta, float* zero_out, int64_t zero_ld, void* stream) {
  SSB_REQUIRE(O && dO && delta && B >= 1 && T >= 1 && H >= 1 && dh >= 1, "attn_delta: bad arguments");
  const bool vec = dh % 4 == 0 && H * dh <= 128 * DELTA_MAXV && (((uintptr_t)O | (uintptr_t)dO) & 15) == 0;
  SSB_REQUIRE(!zero_out || (vec && zero_ld % 4 == 0 && ((uintptr_t)zero_out & 15) == 0),
              "attn_delta: zero_out needs the vector path (dh %% 4 == 0, 16 B aligned rows)");
  if (vec) {
    attn_delta_kernel<<<(unsigned)((B * T + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        O, dO, B * T, (int)T, (int)H, (int)dh, delta, zero_out, zero_ld);
  } else {
    const int64_t warps = B * T * H;
    attn_delta_scalar_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
        O, dO, B * T, (int)T, (int)H, (int)dh, delta);
  }
  SSB_LAUNCH_CHECK("attn_delta");
  return SSB_OK;
}

int ssb_attn_fused_bwd(const void* qkv_planes, const void* dO_planes, const float* R,
                       const float* stat_m, const float* stat_linv, const float* delta, int64_t B,
                       int64_t T, int64_t H, int64_t dh, int64_t W, int64_t RW, float drop_p,
                       uint64_t seed, uint32_t site, float* dqkv, int64_t dq_ld, void* dkv_planes,
                       void* dSband_planes, int64_t RWp, int64_t head_stride, int64_t do_head_stride,
                       void* stream) {
  const int64_t hs = head_stride > 0 ? head_stride : 128;
  const int64_t dhs = do_head_stride > 0 ? do_head_stride : 128;
  FusedParams p = {};
  if (int rc = fill(&p, B, T, H, dh, W, RW, drop_p, seed, site)) return rc;
  SSB_REQUIRE(qkv_planes && dO_planes && R && stat_m && stat_linv && delta && dqkv && dSband_planes,
              "attn_fused_bwd: null pointer");
  SSB_REQUIRE(RWp >= 2 * W + 1 && ((uintptr_t)dqkv & 15) == 0, "attn_fused_bwd: bad RWp / alignment");
  p.stat_m = const_cast<float*>(stat_m); p.stat_linv = const_cast<float*>(stat_linv);
  SSB_REQUIRE(dkv_planes ? dq_ld >= H * dh : dq_ld == 3 * H * dh,
              "attn_fused_bwd: dq_ld=%lld (3*H*dh when dK / dV are stored as fp32, >= H*dh with planes)",
              (long long)dq_ld);
  SSB_REQUIRE(((uintptr_t)dkv_planes & 7) == 0, "attn_fused_bwd: dkv_planes alignment");
  p.delta = delta; p.dqkv = dqkv; p.dq_ld = dq_ld; p.dkv_planes = (__nv_bfloat16*)dkv_planes;
  p.dsb = (__nv_bfloat16*)dSband_planes; p.RWp = (int)RWp;
  CUtensorMap mq, mk, mv, mdo, mr;
  if (int rc = head_map(&mq, qkv_planes, B, T, 3 * H, 0, H, CH, hs, dh)) return rc;
  if (int rc = head_map(&mk, qkv_planes, B, T, 3 * H, H, H, QT, hs, dh)) return rc;
  if (int rc = head_map(&mv, qkv_planes, B, T, 3 * H, 2 * H, H, QT, hs, dh)) return rc;
  if (int rc = head_map(&mdo, dO_planes, B, T, H, 0, H, CH, dhs, dh)) return rc;
  if (int rc = r_map(&mr, R, B * H, T, RW, RBOX, CH)) return rc;
  const int smem = BWD_SMEM_BYTES;
  if (int rc = set_smem(attn_fused_bwd_kernel, smem)) return rc;
  dim3 grid((unsigned)((T + QT - 1) / QT), (unsigned)(B * H));
  attn_fused_bwd_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(mq, mk, mv, mdo, mr, p);
  SSB_LAUNCH_CHECK("attn_fused_bwd");
  return SSB_OK;
}

}  // extern "C"
