// Weight operand preparation: ONE launch turns every parameter of the model into the bf16 hi/lo
// split planes, in every layout the tcgen05 GEMMs of a training step consume them in.
//
// Weights change once per optimiser step, yet round 1 re-derived their operand layouts at every
// use: `.t().contiguous()` copies, a `cat` of three permuted views for the fused QKV weight, a
// split pass per use — ~16 tiny launches per encoder layer and direction.  Here the model keeps a
// device table of entries (silent_speech_b200/weights.py); each entry reads a 32 x 32 tile of a
// strided 2-D VIEW of a parameter once (64 x 64 tiles) and writes it as split planes in up to two destinations:
// as read (dst_n, row-major over the view) and transposed (dst_t).  The views express
//   nn.Linear weight (N, K)            -> forward operand [N][K] and data-gradient operand [K][N]
//   w_q / w_k / w_v (H, D, dh)         -> fused QKV operand [3D][D] and its transpose [D][3D]
//   w_o (H, dh, D)                     -> [D][D] both ways
//   nn.Conv1d weight (Cout, Cin, k)    -> forward operand [Cout][(tap, ci)] and the per-tap-subset
//                                         data-gradient operands [Cin][(j, co)]
// (architecture.py:18-24,51-59; transformer.py:32-34,71-78) without changing the parameters'
// own shapes or names (checkpoint contract).
#include "ssb_common.cuh"
#include <cuda_bf16.h>

namespace {

constexpr int PT = 64;            // tile edge: one CTA moves a 64 x 64 tile of the view

__device__ __forceinline__ int64_t off_r(const ssb_prep_entry_t& e, int r) {
  const int rh = r / e.RL;
  return rh * e.s_rhi + (r - rh * e.RL) * e.s_rlo;
}
__device__ __forceinline__ int64_t off_c(const ssb_prep_entry_t& e, int c) {
  const int ch = c / e.CL;
  return ch * e.s_chi + (c - ch * e.CL) * e.s_clo;
}

// two adjacent values -> one 4-byte store per plane
__device__ __forceinline__ void put2(__nv_bfloat16* dst, int64_t plane, float a, float b) {
  const __nv_bfloat162 hi = __floats2bfloat162_rn(a, b);
  const __nv_bfloat162 lo = __floats2bfloat162_rn(a - __low2float(hi), b - __high2float(hi));
  *reinterpret_cast<__nv_bfloat162*>(dst) = hi;
  *reinterpret_cast<__nv_bfloat162*>(dst + plane) = lo;
}
__device__ __forceinline__ void put1(__nv_bfloat16* dst, int64_t plane, float v) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  dst[0] = hi;
  dst[plane] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__global__ void __launch_bounds__(256)
prep_planes_kernel(const ssb_prep_entry_t* __restrict__ tab, int n_entries) {
  __shared__ float tile[PT][PT + 1];
  __shared__ int which;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  // entry owning this tile: every thread tests one entry (one round trip instead of a search)
  for (int i = threadIdx.x; i < n_entries; i += 256) {
    const int t0 = tab[i].tile0;
    const int t1 = i + 1 < n_entries ? tab[i + 1].tile0 : 0x7fffffff;
    if (t0 <= (int)blockIdx.x && (int)blockIdx.x < t1) which = i;
  }
  __syncthreads();
  const ssb_prep_entry_t e = tab[which];
  const int local = (int)blockIdx.x - e.tile0;
  const int r0 = (local / e.tiles_c) * PT, c0 = (local % e.tiles_c) * PT;
  const bool rows_fast = e.s_rlo == 1 && e.s_clo != 1;   // source contiguous along the view's rows
  if (rows_fast) {      // lanes run along rows: coalesced reads of a column of the view
    int64_t ro[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) ro[h] = off_r(e, min(r0 + tx + 32 * h, e.rows - 1));
    for (int j = ty; j < PT; j += 8) {
      const int c = c0 + j;
      const int64_t co = off_c(e, min(c, e.cols - 1));
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = r0 + tx + 32 * h;
        tile[tx + 32 * h][j] = (r < e.rows && c < e.cols) ? __ldg(e.src + ro[h] + co) : 0.f;
      }
    }
  } else {              // lanes run along columns
    int64_t co[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) co[h] = off_c(e, min(c0 + tx + 32 * h, e.cols - 1));
    for (int j = ty; j < PT; j += 8) {
      const int r = r0 + j;
      const int64_t ro = off_r(e, min(r, e.rows - 1));
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = c0 + tx + 32 * h;
        tile[j][tx + 32 * h] = (r < e.rows && c < e.cols) ? __ldg(e.src + ro + co[h]) : 0.f;
      }
    }
  }
  __syncthreads();
  if (e.dst_n) {        // row r of the view, columns c0 + 2*tx, +1
    __nv_bfloat16* base = (__nv_bfloat16*)e.dst_n;
    const int c = c0 + 2 * tx;
    const bool pair_ok = (e.ld_n % 2 == 0) && ((e.plane_n % 2) == 0);
    for (int j = ty; j < PT; j += 8) {
      const int r = r0 + j;
      if (r >= e.rows || c >= e.cols) continue;
      __nv_bfloat16* d = base + (int64_t)r * e.ld_n + c;
      if (pair_ok && c + 1 < e.cols) put2(d, e.plane_n, tile[j][2 * tx], tile[j][2 * tx + 1]);
      else {
        put1(d, e.plane_n, tile[j][2 * tx]);
        if (c + 1 < e.cols) put1(d + 1, e.plane_n, tile[j][2 * tx + 1]);
      }
    }
  }
  if (e.dst_t) {        // row c of the transpose, columns r0 + 2*tx, +1
    __nv_bfloat16* base = (__nv_bfloat16*)e.dst_t;
    const int r = r0 + 2 * tx;
    const bool pair_ok = (e.ld_t % 2 == 0) && ((e.plane_t % 2) == 0);
    for (int j = ty; j < PT; j += 8) {
      const int c = c0 + j;
      if (c >= e.cols || r >= e.rows) continue;
      __nv_bfloat16* d = base + (int64_t)c * e.ld_t + r;
      if (pair_ok && r + 1 < e.rows) put2(d, e.plane_t, tile[2 * tx][j], tile[2 * tx + 1][j]);
      else {
        put1(d, e.plane_t, tile[2 * tx][j]);
        if (r + 1 < e.rows) put1(d + 1, e.plane_t, tile[2 * tx + 1][j]);
      }
    }
  }
}

}  // namespace

extern "C" {

int64_t ssb_prep_plan(ssb_prep_entry_t* table_host, int64_t n_entries) {
  if (n_entries < 0 || (n_entries > 0 && !table_host)) {
    ssb::set_error("prep plan: bad table");
    return SSB_ERR_ARG;
  }
  int64_t tiles = 0;
  for (int64_t i = 0; i < n_entries; ++i) {
    ssb_prep_entry_t* e = table_host + i;
    if (e->rows < 1 || e->cols < 1 || e->RL < 1 || e->CL < 1 || !e->src || (!e->dst_n && !e->dst_t)) {
      ssb::set_error("prep plan: entry %lld is malformed (rows=%d cols=%d RL=%d CL=%d)",
                     (long long)i, e->rows, e->cols, e->RL, e->CL);
      return SSB_ERR_ARG;
    }
    e->tiles_c = (e->cols + PT - 1) / PT;
    e->tile0 = (int32_t)tiles;
    tiles += (int64_t)e->tiles_c * ((e->rows + PT - 1) / PT);
    if (tiles >= (1LL << 31)) {
      ssb::set_error("prep plan: too many tiles");
      return SSB_ERR_ARG;
    }
  }
  return tiles;
}

int ssb_prep_planes(const ssb_prep_entry_t* table_dev, int64_t n_entries, int64_t total_tiles,
                    void* stream) {
  if (n_entries == 0 || total_tiles == 0) return SSB_OK;
  SSB_REQUIRE(table_dev && n_entries > 0 && total_tiles > 0 && total_tiles < (1LL << 31),
              "prep: bad table (%lld entries, %lld tiles)", (long long)n_entries,
              (long long)total_tiles);
  prep_planes_kernel<<<(unsigned)total_tiles, 256, 0, (cudaStream_t)stream>>>(
      table_dev, (int)n_entries);
  SSB_LAUNCH_CHECK("prep_planes_kernel");
  return SSB_OK;
}

}  // extern "C"
