// Weight operand preparation: ONE launch turns every parameter of the model into the bf16 hi/lo
// split planes, in every layout the tcgen05 GEMMs of a training step consume them in.
//
// Weights change once per optimiser step, yet round 1 re-derived their operand layouts at every
// use: `.t().contiguous()` copies, a `cat` of three permuted views for the fused QKV weight, a
// split pass per use — ~16 tiny launches per encoder layer and direction.  Here the model keeps a
// device table of entries (silent_speech_b200/weights.py); each entry reads a 32 x 32 tile of a
// strided 2-D VIEW of a parameter once and writes it as split planes in up to two destinations:
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

__device__ __forceinline__ int64_t src_off(const ssb_prep_entry_t& e, int r, int c) {
  const int rh = r / e.RL, rl = r - rh * e.RL;
  const int ch = c / e.CL, cl = c - ch * e.CL;
  return rh * e.s_rhi + rl * e.s_rlo + ch * e.s_chi + cl * e.s_clo;
}

__device__ __forceinline__ void put(__nv_bfloat16* dst, int64_t plane, float v) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  dst[0] = hi;
  dst[plane] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__global__ void __launch_bounds__(256)
prep_planes_kernel(const ssb_prep_entry_t* __restrict__ tab, int n_entries) {
  __shared__ float tile[32][33];
  __shared__ ssb_prep_entry_t es;
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (tx == 0 && ty == 0) {
    int lo = 0, hi = n_entries;       // last entry with tile0 <= blockIdx.x
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (tab[mid].tile0 <= (int)blockIdx.x) lo = mid; else hi = mid;
    }
    es = tab[lo];
  }
  __syncthreads();
  const ssb_prep_entry_t& e = es;
  const int local = (int)blockIdx.x - e.tile0;
  const int r0 = (local / e.tiles_c) * 32, c0 = (local % e.tiles_c) * 32;
  const bool rows_fast = e.s_rlo == 1 && e.s_clo != 1;   // source contiguous along the view's rows
#pragma unroll
  for (int j = ty; j < 32; j += 8) {
    if (rows_fast) {
      const int r = r0 + tx, c = c0 + j;
      tile[tx][j] = (r < e.rows && c < e.cols) ? __ldg(e.src + src_off(e, r, c)) : 0.f;
    } else {
      const int r = r0 + j, c = c0 + tx;
      tile[j][tx] = (r < e.rows && c < e.cols) ? __ldg(e.src + src_off(e, r, c)) : 0.f;
    }
  }
  __syncthreads();
  if (e.dst_n) {
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
      const int r = r0 + j, c = c0 + tx;
      if (r < e.rows && c < e.cols)
        put((__nv_bfloat16*)e.dst_n + (int64_t)r * e.ld_n + c, e.plane_n, tile[j][tx]);
    }
  }
  if (e.dst_t) {
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
      const int c = c0 + j, r = r0 + tx;
      if (r < e.rows && c < e.cols)
        put((__nv_bfloat16*)e.dst_t + (int64_t)c * e.ld_t + r, e.plane_t, tile[tx][j]);
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
    e->tiles_c = (e->cols + 31) / 32;
    e->tile0 = (int32_t)tiles;
    tiles += (int64_t)e->tiles_c * ((e->rows + 31) / 32);
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
  prep_planes_kernel<<<(unsigned)total_tiles, dim3(32, 8), 0, (cudaStream_t)stream>>>(
      table_dev, (int)n_entries);
  SSB_LAUNCH_CHECK("prep_planes_kernel");
  return SSB_OK;
}

}  // extern "C"
