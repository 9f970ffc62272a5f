// Batched DTW alignment for sm_100a: register-resident wavefront fill + 2-bit direction
// codes + backtrace.  Replaces align.py:5-14 (time_warp) and align.py:16-34
// (align_from_distances) of the reference.
//
// Layout vocabulary.  A cost matrix has a contiguous "strip" axis y (extent Ny) and a
// strided "sweep" axis x (extent Nx, `pitch` elements apart).  The reference hands DTW
// an F-ordered view (costs.T, transduction_model.py:126), so there y == DTW row i and
// x == DTW column j (Y_IS_I); for a C-ordered matrix y == j and x == i.  The recurrence is
// symmetric under that swap, only the tie-break order changes, so one kernel serves both.
//
// Work decomposition.  A pair is cut into bands of BAND = 32 lanes x R strip-rows; one warp
// sweeps one band.  Inside a band lane l owns strip rows [y0, y0+R) and, at step s, computes
// sweep column x = s - l (a systolic skew of one column per lane): the value of the row above
// comes from lane l-1 by one shuffle per step, the diagonal is the previous shuffle result, the
// left neighbour is the lane's own register.  Each cell does exactly one fp32 add
// (cost + min3), so the result is bit-identical to the reference's sequential loop.
//
// Band pipeline.  The W warps of a CTA take the bands of the CTA's pairs round-robin (band
// instance g -> warp g % W) and run them CONCURRENTLY: the bottom row of band b is parked in the
// producing warp's shared-memory boundary buffer and read by the warp sweeping band b+1, which
// follows three 16-step chunks behind (progress words in shared memory, one update per chunk).
// Two things come from that: (1) a lone pair finishes in (nch + 3 (nbands-1)) chunk times instead
// of nch * nbands -- the 16-pair training batch is latency-bound; (2) the 128 B lines that
// straddle a band boundary (pitch 600 floats: three columns of four start off a line boundary)
// are requested by the two neighbouring warps within ~1 us of each other, so the second request
// hits L2 instead of fetching the line from DRAM again (r2 capture of the one-warp-per-pair
// kernel: DRAM read 1.198x algorithmic; the model "every band-column fetch rounds out to 128 B
// lines" gives 1.200x).
//
// HBM traffic.  Cost tiles stream global->shared with cp.async: at step s lane l = 8g+q
// copies ITS OWN 16 B (its R=4 rows) of column s-8g+PF, so every group of 8 lanes fetches
// one full 128 B line, each lane only ever reads shared memory it filled itself (no
// cross-lane visibility hazards), and the per-lane ring is RING=16 columns deep
// (PF=8 columns of prefetch + 8 of intra-group skew): 8 KB of shared memory per warp.
// Cost is read exactly once: 4 B/cell algorithmic, plus 0.28 B/cell of direction codes
// written in the skewed (step-major) order so that every 16 steps a warp stores one
// coalesced 512 B row of packed codes.
#include "ssb_common.cuh"
#include <math_constants.h>
#include <type_traits>

namespace {

constexpr int R = 4;            // strip rows per lane
constexpr int BAND = 32 * R;    // strip rows per band
constexpr int RING = 16;        // ring depth (columns), power of two
constexpr int PF = 8;           // prefetch distance (columns); RING - PF >= 8
constexpr int RING_BYTES = RING * 32 * 16;
constexpr int RING_MASK = (RING - 1) << 9;  // byte offset of a column inside the ring
constexpr int CH = 16;          // steps per chunk == 2-bit codes per direction word
static_assert(RING - PF >= 8, "ring geometry");

struct DtwParams {
  const float* cost;
  float* dtw_out;  // nullable
  uint32_t* dirs;
  // ragged batches (ssb_dtw_align_ragged): per-pair geometry; Ny/Nx/nbands/nch below then hold the
  // MAXIMA over the batch (shared-memory sizing) and pair_stride/pitch/dirs_pair_words are unused
  const ssb_dtw_pair_t* table;
  int64_t pair_stride;
  int64_t pitch;
  int64_t dirs_pair_words;
  int Ny, Nx, nbands, nch, npairs;
};

// Direction codes: 2 raw predicate bits per cell: bit0 = "second candidate (i,j-1) < first
// (i-1,j)", bit1 = "diagonal < min(first, second)".  The backtrace resolves them in the
// first-wins order of Python's min() (align.py:26): bit1 -> diagonal, else bit0 -> left, else up.
// One word packs the 16 steps of a chunk for one strip row (earliest step in the top bits);
// words are stored step-major:  dirs[(band*nch + chunk)*32 + lane]  (uint4 = the lane's 4 rows)
// so that every chunk ends with one coalesced 512 B store per warp.

// boundary row of a band: one float per sweep column, plus the 31 + 16 steps past the last column
// that the edge chunks read (and ignore)
__host__ __device__ constexpr int bnd_bytes_for(int Nx) { return ((Nx + 48 + 3) & ~3) * 4; }

constexpr int MAX_WARPS = 8;    // band-pipeline width (warps per CTA)
constexpr int PROG_SHIFT = 12;  // progress word = (band instance of the warp << 12) + chunks done
constexpr int LAG = 3;          // chunks the consumer of a boundary row stays behind its producer:
                                // chunk c reads bnd[16c .. 16c+15], written at steps <= 16c+15+31

// spin until the progress word of another warp of this CTA reaches `need` (`seen` caches the last
// value read, so a consumer that is far enough behind does not touch shared memory at all)
__device__ __forceinline__ void wait_progress(const volatile uint32_t* word, uint32_t need,
                                              uint32_t& seen) {
  if (seen >= need) return;
  uint32_t v = *word;
  while (v < need) {
    __nanosleep(32);
    v = *word;
  }
  seen = v;
  __threadfence_block();   // the boundary values published before `word` are visible from here on
}

template <bool Y_IS_I, bool VEC, bool WRITE_DTW, bool RAGGED>
__global__ void __launch_bounds__(MAX_WARPS * 32, 3) dtw_fill_kernel(const DtwParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int W = blockDim.x >> 5;
  const int bnd_bytes = bnd_bytes_for(p.Nx);
  const int per_warp = RING_BYTES + bnd_bytes;
  unsigned char* wbase = smem + (size_t)warp * per_warp;
  const unsigned char* ring_lane = wbase + lane * 16;  // this lane's 16 B slot of every column
  const int pwarp = (warp + W - 1) % W, cwarp = (warp + 1) % W;   // producer / consumer neighbours
  float* bnd_out = reinterpret_cast<float*>(wbase + RING_BYTES);                      // this band's bottom row
  const float* bnd_in = reinterpret_cast<const float*>(smem + (size_t)pwarp * per_warp + RING_BYTES);
  volatile uint32_t* prog = reinterpret_cast<volatile uint32_t*>(smem + (size_t)W * per_warp);
  // a row of +inf stands in for the boundary above band 0, so the top-row read needs no predicate
  float* inf_row = reinterpret_cast<float*>(smem + (size_t)W * per_warp + 64);
  const uint32_t ring_u32 = ssb::smem_u32(wbase) + lane * 16;
  const int g8 = lane & ~7;
  const float INF = CUDART_INF_F;
  // chunks in which every lane is inside [1, Nx) and every prefetch is in range
  const int steady_lo = 32 / CH;
  // steady chunks start at a multiple of RING = CH steps, so step u of a chunk writes ring slot
  // (u + PF - g8) % RING: slot u for the lane groups with g8 % 16 == 8, else slot (u + 8) % 16.
  // Two per-lane bases turn that into immediate offsets (wr_a for u < 8, wr_b for u >= 8).
  static_assert(RING == 16 && CH == 16 && PF == 8, "write-slot bases assume RING = CH = 16, PF = 8");
  const uint32_t wr_a = ring_u32 + ((g8 & 8) ? 0u : 8u * 512u);
  const uint32_t wr_b = ring_u32 - ((g8 & 8) ? 0u : 8u * 512u);
  const bool lane0 = (lane == 0);

  if (threadIdx.x < W) prog[threadIdx.x] = 0u;
  for (int x = threadIdx.x; x < bnd_bytes / 4; x += blockDim.x) inf_row[x] = INF;
  if (lane == 0) bnd_out[0] = INF;    // column 0 of every boundary row (producers write x >= 1 only)
  __syncthreads();

  int turn = 0;             // (band instance counter of the CTA) % W
  uint32_t k = 0;           // band instances this warp has swept
  uint32_t seen_p = 0, seen_c = 0;
  bool prev_wrote = false;  // this warp's previous instance parked a boundary row ...
  int prev_nch = 0;         // ... of this many chunks

  for (int pair = blockIdx.x; pair < p.npairs; pair += gridDim.x) {
    int Nx = p.Nx, Ny = p.Ny, nbands = p.nbands, nch = p.nch;
    int64_t pitch = p.pitch;
    const float* cost = p.cost + (int64_t)pair * p.pair_stride;
    float* dout = WRITE_DTW ? p.dtw_out + (int64_t)pair * p.pair_stride : nullptr;
    uint4* dirs = reinterpret_cast<uint4*>(p.dirs + (int64_t)pair * p.dirs_pair_words);
    if constexpr (RAGGED) {   // this pair's own geometry (Y_IS_I: Ny = N, Nx = M)
      const ssb_dtw_pair_t t = p.table[pair];
      Ny = t.N; Nx = t.M; nbands = t.nbands; nch = t.nch; pitch = t.pitch;
      cost = p.cost + t.cost_off;
      dirs = reinterpret_cast<uint4*>(p.dirs + t.dirs_off);
    }
    const int steady_hi = (Nx - CH - PF) / CH;  // inclusive; may be < steady_lo
    const bool pitch_fits = pitch < (1LL << 30);   // steady chunks address columns as base + u * pitch_bytes
    const uint32_t pitch_bytes = (uint32_t)pitch * 4u;

    for (int band = 0; band < nbands; ++band) {
      const bool mine = (turn == warp);
      turn = (turn + 1 == W) ? 0 : turn + 1;
      if (!mine) continue;

      const int y0 = band * BAND + lane * R;
      const bool row0 = (y0 == 0);
      const int rows_valid = min(max(Ny - y0, 0), R);
      const int y0_ld = rows_valid > 0 ? y0 : 0;  // keep the address legal for idle lanes
      const int ld_bytes = rows_valid * 4;
      const bool has_in = band > 0, has_out = band + 1 < nbands;
      const bool wr_bnd = has_out && (lane == 31);
      const float* top = has_in ? bnd_in : inf_row;   // the row above this band, by column
      // progress words this instance waits on: the producer of its top boundary (the previous band
      // instance, swept by warp - 1) and the reader of the row this warp parked last time
      const uint32_t p_base = (warp == 0 ? k - 1 : k) << PROG_SHIFT;
      const uint32_t c_base = (warp == W - 1 ? k : k - 1) << PROG_SHIFT;
      const bool guard_out = has_out && prev_wrote;

      float v[R];
      float pa[R];        // direction codes of the current 8 steps (2 bits each, as a float)
      uint32_t pk[R];     // ... of the first 8 steps of the chunk
#pragma unroll
      for (int r = 0; r < R; ++r) {
        v[r] = INF;  // column x = 0 of the DTW table
        pa[r] = 0.f;
        pk[r] = 0u;
      }
      // code bookkeeping around the 4 steps of quad `quad` (compile-time under the unrolled loops)
      auto codes_begin = [&](int quad) {
        if (quad == 0 || quad == 2) {
#pragma unroll
          for (int r = 0; r < R; ++r) pa[r] = 0.f;
        }
      };
      auto codes_end = [&](int quad) {
        if (quad == 1) {
#pragma unroll
          for (int r = 0; r < R; ++r) pk[r] = __float2uint_rn(pa[r]);
        } else if (quad == 3) {
#pragma unroll
          for (int r = 0; r < R; ++r) pk[r] = pk[r] * 65536u + __float2uint_rn(pa[r]);
        }
      };
      if (row0) v[0] = 0.f;  // dtw[0,0]
      if (WRITE_DTW) {
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (r < rows_valid) dout[y0 + r] = v[r];
      }
      float up_in = INF, diag_in = INF;

      auto prefetch = [&](uint32_t dst, const float* src) {
        if (VEC) {
            ssb::cp_async_16(dst, src, ld_bytes);
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (r < rows_valid) ssb::cp_async_4(dst + 4 * r, src + r, 4);
        }
      };

      auto cells = [&](const float4 c4, int x) {
        float cc[R] = {c4.x, c4.y, c4.z, c4.w};
        if (row0) cc[0] = INF;  // dtw[0, x>=1] = +inf
        float upc = up_in, upp = diag_in;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float old = v[r];
          // reference order (align.py:13,26): up=(i-1,j), left=(i,j-1), diag; first wins
          const float first = Y_IS_I ? upc : old;
          const float second = Y_IS_I ? old : upc;
          float nv;
          // The half-rate ALU pipe (compare / select / min) bounds this kernel, so the cell is
          // written for it: FSET x2 + FMNMX x2 there, FADD + 2 FFMA on the FMA pipe (r2 run with
          // FSETP FSEL SEL x2 = 6 half-rate instructions per cell: ALU pipe 68 % busy, issue 62 %).
          // The two predicates arrive as 1.0f / 0.0f and the 2-bit codes of 8 steps are packed in
          // a float (exact: < 4^8), converted twice per chunk.
          asm("{\n\t"
              ".reg .f32 m1, m, r1, r2;\n\t"
              "set.lt.f32.f32 r1, %3, %2;\n\t"
              "min.f32 m1, %2, %3;\n\t"
              "set.lt.f32.f32 r2, %4, m1;\n\t"
              "min.f32 m, m1, %4;\n\t"
              "add.rn.f32 %0, %5, m;\n\t"
              "fma.rn.f32 r1, r2, 0f40000000, r1;\n\t"
              "fma.rn.f32 %1, %1, 0f40800000, r1;\n\t"
              "}"
              : "=f"(nv), "+f"(pa[r])
              : "f"(first), "f"(second), "f"(upp), "f"(cc[r]));
          upp = old;
          upc = nv;
          v[r] = nv;
        }
        if (WRITE_DTW) {
          float* o = dout + (int64_t)x * pitch + y0;
          if (VEC && rows_valid == R) {
            *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int r = 0; r < R; ++r)
              if (r < rows_valid) o[r] = v[r];
          }
        }
      };

      // ---- general step: every range predicate evaluated (head / tail chunks) ---------
      auto step_general = [&](int s) {
        const int xl = s - g8 + PF;
        if (xl >= 1 && xl < Nx)
          prefetch(ring_u32 + ((xl << 9) & RING_MASK), cost + ((int64_t)xl * pitch + y0_ld));
        ssb::cp_async_commit();
        const float t = __shfl_up_sync(0xffffffffu, v[R - 1], 1);
        float bval = INF;
        if (lane0 && s >= 1 && s < Nx) bval = top[s];
        diag_in = up_in;
        up_in = lane0 ? bval : t;
        const int x = s - lane;
        ssb::cp_async_wait<PF>();   // the column issued PF steps ago (and everything older) has landed
        if (x >= 1 && x < Nx) {
          cells(*reinterpret_cast<const float4*>(ring_lane + ((x << 9) & RING_MASK)), x);
          if (wr_bnd) bnd_out[x] = v[R - 1];
        } else {
#pragma unroll
          for (int r = 0; r < R; ++r) pa[r] *= 4.f;   // code 0
        }
      };

      if (!WRITE_DTW && VEC) {
        // edge chunks read columns <= 0 as +inf: every slot of this lane's ring starts out so
        // (a slot is overwritten by column x + 16 only after the lane has consumed column x)
#pragma unroll
        for (int sl = 0; sl < RING; ++sl)
          asm volatile("st.shared.v4.f32 [%0], {%1, %1, %1, %1};" ::"r"(ring_u32 + (uint32_t)sl * 512u),
                       "f"(INF)
                       : "memory");
      }
      // prologue: the loads that steps s < 0 would have issued.  One commit group per step: a
      // step waits only for the column issued PF steps earlier, so PF columns (4 KB per warp)
      // stay in flight instead of 4 .. 8 with one group per four steps.
#pragma unroll 1
      for (int u = 0; u < PF; ++u) {
        const int xl = u - g8;
        if (xl >= 1 && xl < Nx)
          prefetch(ring_u32 + ((xl << 9) & RING_MASK), cost + ((int64_t)xl * pitch + y0_ld));
        ssb::cp_async_commit();
      }

#pragma unroll 1
      for (int c = 0; c < nch; ++c) {
        const int s0 = c * CH;
        // chunk c reads bnd_in[s0 .. s0+15]: the producer wrote them by the end of its chunk c+2;
        // it writes bnd_out[s0-31 .. s0-16]: their previous contents were read by the consumer of
        // this warp's previous instance in its chunks <= c-1
        if (has_in) wait_progress(prog + pwarp, p_base + (uint32_t)min(c + LAG, nch), seen_p);
        if (guard_out) wait_progress(prog + cwarp, c_base + (uint32_t)min(c, prev_nch), seen_c);
        // ---- fast chunk: running addresses, no per-cell range checks -------------------
        // column u of the chunk at src0 + u * pitch_bytes (one IMAD.WIDE); the ring read slot
        // lives in the top 4 bits of rd_hi, so the offset of step u is one IMAD.HI.
        // EDGE (head / tail chunks): lanes outside [1, Nx) run the same cell code on +inf costs
        // (head: the ring was pre-filled, so column <= 0 stays +inf) or on stale ring contents
        // (tail: nothing reads those values or codes); only the copies and the boundary store are
        // predicated.  r2: the fully predicated general step cost 125 instructions against 42.
        auto fast_chunk = [&](auto edge_tag) {
          constexpr bool EDGE = decltype(edge_tag)::value;
          const uint64_t src0 = reinterpret_cast<uint64_t>(cost + ((int64_t)(s0 - g8 + PF) * pitch + y0_ld));
          const uint32_t rd_hi = (uint32_t)(s0 - lane) << 28;
          const float* top_rd = top + s0;             // every lane reads top[s] (lane 0 uses it)
          float* bnd_wr = bnd_out + (s0 - lane);      // lane 31 writes bnd[x]
          const int pf_lo = 1 - (s0 - g8 + PF), pf_hi = Nx - (s0 - g8 + PF);   // copy column u iff pf_lo <= u < pf_hi
          const int bw_lo = 1 - (s0 - lane), bw_hi = Nx - (s0 - lane);        // store bnd[x] iff bw_lo <= u < bw_hi
#pragma unroll
          for (int quad = 0; quad < CH / 4; ++quad) {
            codes_begin(quad);
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) {
              const int u = quad * 4 + u4;
              if (!EDGE || (u >= pf_lo && u < pf_hi))
                prefetch((u < 8 ? wr_a : wr_b) + (uint32_t)u * 512u,
                         reinterpret_cast<const float*>(src0 + (uint64_t)(uint32_t)u * (uint64_t)pitch_bytes));
              ssb::cp_async_commit();
              if (EDGE && u == 0 && s0 == 0) continue;   // step 0: every lane is at a column <= 0
              const float t = __shfl_up_sync(0xffffffffu, v[R - 1], 1);
              const float bval = top_rd[u];
              diag_in = up_in;
              up_in = lane0 ? bval : t;
              ssb::cp_async_wait<PF>();
              float4 c4;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(c4.x), "=f"(c4.y), "=f"(c4.z), "=f"(c4.w)
                           : "r"(ring_u32 + __umulhi(rd_hi + ((uint32_t)u << 28), (uint32_t)RING_BYTES)));
              cells(c4, s0 + u - lane);
              if (wr_bnd && (!EDGE || (u >= bw_lo && u < bw_hi))) bnd_wr[u] = v[R - 1];
            }
            codes_end(quad);
          }
        };
        if (c >= steady_lo && c <= steady_hi && pitch_fits) {
          fast_chunk(std::false_type{});
        } else if (!WRITE_DTW && VEC && pitch_fits) {
          fast_chunk(std::true_type{});
        } else {
#pragma unroll
          for (int quad = 0; quad < CH / 4; ++quad) {
            codes_begin(quad);
#pragma unroll
            for (int u4 = 0; u4 < 4; ++u4) step_general(s0 + quad * 4 + u4);
            codes_end(quad);
          }
        }
        dirs[((int64_t)band * nch + c) * 32 + lane] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        // publish: lane 31's boundary stores of this chunk, then the progress word
        __syncwarp();
        if (lane == 0) {
          __threadfence_block();
          prog[warp] = (c + 1 == nch) ? ((k + 1) << PROG_SHIFT) : ((k << PROG_SHIFT) + (uint32_t)(c + 1));
        }
      }
      ssb::cp_async_wait<0>();
      __syncwarp();
      ++k;
      prev_wrote = has_out;
      prev_nch = nch;
    }
  }
}

// One thread per pair walks the path from (N-1, M-1) (align.py:19-26).  The walk is a chain of
// ~N+M dependent loads; whenever it enters a new 32 B sector of direction words it prefetches
// the two sectors it can move to next (previous rows of the same chunk, previous chunk of the
// same rows) so that most sector changes hit L2/L1 instead of paying a DRAM round trip.
template <bool Y_IS_I>
__global__ void dtw_backtrace_kernel(const uint32_t* __restrict__ dirs, int64_t dirs_pair_words,
                                     int nch, int N, int M, int npairs,
                                     int32_t* __restrict__ path,
                                     const ssb_dtw_pair_t* __restrict__ table) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair >= npairs) return;
  const uint32_t* d = dirs + (int64_t)pair * dirs_pair_words;
  int32_t* out = path + (int64_t)pair * N;    // ragged: rows are padded to the batch maximum N
  if (table != nullptr) {
    const ssb_dtw_pair_t t = table[pair];
    d = dirs + t.dirs_off;
    for (int rr = t.N; rr < N; ++rr) out[rr] = 0;   // padding rows
    N = t.N; M = t.M; nch = t.nch;
  }
  int i = N - 1, j = M - 1;
  int64_t cur_idx = -1;
  uint32_t cur_word = 0;
  int lowest = N;   // smallest row index assigned so far
  while (i > 0 && j > 0) {
    out[i] = j;
    lowest = i;
    const int x = Y_IS_I ? j : i, y = Y_IS_I ? i : j;
    const int band = y / BAND, l = (y % BAND) / R, r = y % R, s = x + l;
    const int64_t idx = (((int64_t)band * nch + (s / CH)) * 32 + l) * 4 + r;
    if (idx != cur_idx) {
      if ((idx >> 3) != (cur_idx >> 3)) {   // new 32 B sector
        if (idx >= 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(d + idx - 8));
        if (idx >= 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(d + idx - 128));
      }
      cur_word = __ldg(d + idx);
      cur_idx = idx;
    }
    const uint32_t code = (cur_word >> (2 * (CH - 1 - (s % CH)))) & 3u;
    const bool diag = (code & 2u) != 0u, left = (code & 1u) != 0u;
    if (diag || !left) --i;  // up or diagonal
    if (diag || left) --j;   // left or diagonal
  }
  // rows the walk never assigned keep the reference's initial value 0 (align.py:21)
  for (int rr = lowest - 1; rr >= 0; --rr) out[rr] = 0;
}

// Small batches (the 16 silent utterances of a training step): one CTA per pair copies the pair's
// direction words into shared memory with coalesced 16 B loads and one thread walks them there
// (~25 ns per step instead of an L2 round trip per new word: 120 us -> ~35 us at 600 x 500).
template <bool Y_IS_I>
__global__ void __launch_bounds__(256) dtw_backtrace_smem_kernel(
    const uint32_t* __restrict__ dirs, int64_t dirs_pair_words, int nch, int N, int M, int npairs,
    int32_t* __restrict__ path, const ssb_dtw_pair_t* __restrict__ table) {
  extern __shared__ __align__(16) uint32_t sdirs[];
  for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
    const uint32_t* d = dirs + (int64_t)pair * dirs_pair_words;
    int32_t* out = path + (int64_t)pair * N;    // ragged: rows are padded to the batch maximum N
    int n = N, m = M, nc = nch;
    int64_t words = dirs_pair_words;
    if (table != nullptr) {
      const ssb_dtw_pair_t t = table[pair];
      d = dirs + t.dirs_off;
      n = t.N; m = t.M; nc = t.nch;
      words = (int64_t)t.nbands * t.nch * BAND;
    }
    const uint4* d4 = reinterpret_cast<const uint4*>(d);
    uint4* s4 = reinterpret_cast<uint4*>(sdirs);
    for (int64_t w = threadIdx.x; w < words / 4; w += blockDim.x) s4[w] = __ldg(d4 + w);
    // rows the walk never assigns (and the padding rows of a ragged batch) keep the reference's
    // initial value 0 (align.py:21)
    for (int rr = threadIdx.x; rr < N; rr += blockDim.x) out[rr] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
      int i = n - 1, j = m - 1;
      while (i > 0 && j > 0) {
        out[i] = j;
        const int x = Y_IS_I ? j : i, y = Y_IS_I ? i : j;
        const int band = y / BAND, l = (y % BAND) / R, r = y % R, s = x + l;
        const uint32_t word = sdirs[((band * nc + (s / CH)) * 32 + l) * 4 + r];
        const uint32_t code = (word >> (2 * (CH - 1 - (s % CH)))) & 3u;
        const bool diag = (code & 2u) != 0u, left = (code & 1u) != 0u;
        if (diag || !left) --i;  // up or diagonal
        if (diag || left) --j;   // left or diagonal
      }
    }
    __syncthreads();
  }
}

// picks the backtrace variant; `max_words` = direction words of the largest pair
template <bool Y_IS_I>
int launch_backtrace(const uint32_t* dirs, int64_t dirs_pair_words, int64_t max_words, int nch, int N,
                     int M, int npairs, int32_t* path, const ssb_dtw_pair_t* table, cudaStream_t st) {
  const int64_t smem = max_words * 4;
  if (npairs <= 2 * ssb::num_sms() && smem <= 200 * 1024) {
    auto kern = dtw_backtrace_smem_kernel<Y_IS_I>;
    SSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<npairs, 256, smem, st>>>(dirs, dirs_pair_words, nch, N, M, npairs, path, table);
    SSB_LAUNCH_CHECK("dtw_backtrace_smem_kernel");
    return SSB_OK;
  }
  const int threads = 64;
  const int grid = (npairs + threads - 1) / threads;
  dtw_backtrace_kernel<Y_IS_I><<<grid, threads, 0, st>>>(dirs, dirs_pair_words, nch, N, M, npairs,
                                                         path, table);
  SSB_LAUNCH_CHECK("dtw_backtrace_kernel");
  return SSB_OK;
}

struct Geometry {
  bool y_is_i;
  int Ny, Nx, nbands, nch;
  int64_t pitch, dirs_pair_words;
};

int make_geometry(int64_t N, int64_t M, int64_t stride_i, int64_t stride_j, Geometry* g) {
  SSB_REQUIRE(N >= 1 && M >= 1, "dtw: N=%lld M=%lld must be >= 1", (long long)N, (long long)M);
  SSB_REQUIRE(N < (1 << 24) && M < (1 << 24), "dtw: dimension too large");
  SSB_REQUIRE(stride_i == 1 || stride_j == 1,
              "dtw: one of stride_i/stride_j must be 1 (got %lld, %lld)", (long long)stride_i,
              (long long)stride_j);
  SSB_REQUIRE(!(stride_i == 1 && stride_j == 1) || N == 1 || M == 1,
              "dtw: both strides are 1 but the matrix is not a vector");
  g->y_is_i = (stride_i == 1) && (stride_j != 1 || M == 1);
  g->Ny = (int)(g->y_is_i ? N : M);
  g->Nx = (int)(g->y_is_i ? M : N);
  g->pitch = g->y_is_i ? stride_j : stride_i;
  SSB_REQUIRE(g->Nx == 1 || g->pitch >= g->Ny, "dtw: pitch %lld < contiguous extent %d",
              (long long)g->pitch, g->Ny);
  g->nbands = (g->Ny + BAND - 1) / BAND;
  g->nch = (g->Nx + 30) / CH + 1;  // steps 0 .. Nx-1+31
  g->dirs_pair_words = (int64_t)g->nbands * g->nch * BAND;
  return SSB_OK;
}

template <bool Y_IS_I, bool VEC, bool WRITE_DTW, bool RAGGED = false>
int launch_fill(const DtwParams& p, cudaStream_t st) {
  auto kern = dtw_fill_kernel<Y_IS_I, VEC, WRITE_DTW, RAGGED>;
  SSB_REQUIRE(p.nch < (1 << PROG_SHIFT), "dtw: sweep extent %d too large", p.Nx);
  const int bnd_bytes = bnd_bytes_for(p.Nx);
  const int per_warp = RING_BYTES + bnd_bytes;
  // pipeline width: one warp per band of a pair (ragged: of the tallest pair), at least 4 warps per
  // CTA when there are pairs to spare (band instances of consecutive pairs then run side by side)
  int warps = p.nbands < MAX_WARPS ? p.nbands : MAX_WARPS;
  if (warps < 4 && p.npairs >= 4 * ssb::num_sms()) warps = 4;
  while (warps > 1 && warps * per_warp + 64 + bnd_bytes > 100 * 1024) --warps;
  SSB_REQUIRE(warps * per_warp + 64 + bnd_bytes <= 227 * 1024, "dtw: sweep extent %d too large for shared memory",
              p.Nx);
  const int smem = warps * per_warp + 64 + bnd_bytes;   // + progress words + the +inf row
  SSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int occ = 0;
  SSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, smem));
  if (occ < 1) occ = 1;
  const int64_t cap = (int64_t)ssb::num_sms() * occ;
  const int grid = (int)(p.npairs < cap ? p.npairs : cap);
  kern<<<grid, warps * 32, smem, st>>>(p);
  SSB_LAUNCH_CHECK("dtw_fill_kernel");
  return SSB_OK;
}

int run(const float* cost, int64_t npairs, int64_t pair_stride, int64_t N, int64_t M,
        int64_t stride_i, int64_t stride_j, float* dtw, int32_t* path, void* workspace,
        int64_t workspace_bytes, void* stream) {
  Geometry g;
  if (int rc = make_geometry(N, M, stride_i, stride_j, &g)) return rc;
  SSB_REQUIRE(npairs >= 0 && npairs < (1LL << 31), "dtw: bad npairs %lld", (long long)npairs);
  if (npairs == 0) return SSB_OK;
  SSB_REQUIRE(cost && path && workspace, "dtw: null pointer");
  const int64_t need = npairs * g.dirs_pair_words * 4;
  if (workspace_bytes < need) {
    ssb::set_error("dtw: workspace %lld B < required %lld B", (long long)workspace_bytes,
                   (long long)need);
    return SSB_ERR_WORKSPACE;
  }
  SSB_REQUIRE(((uintptr_t)workspace & 15) == 0, "dtw: workspace must be 16 B aligned");
  cudaStream_t st = (cudaStream_t)stream;

  DtwParams p;
  p.cost = cost;
  p.dtw_out = dtw;
  p.dirs = (uint32_t*)workspace;
  p.table = nullptr;
  p.pair_stride = pair_stride;
  p.pitch = g.pitch;
  p.dirs_pair_words = g.dirs_pair_words;
  p.Ny = g.Ny;
  p.Nx = g.Nx;
  p.nbands = g.nbands;
  p.nch = g.nch;
  p.npairs = (int)npairs;
  const bool vec = (g.pitch % 4 == 0) && (pair_stride % 4 == 0) && (((uintptr_t)cost & 15) == 0) &&
                   (!dtw || ((uintptr_t)dtw & 15) == 0);

  int rc;
#define SSB_DTW_DISPATCH(YI, V)                                            \
  rc = dtw ? launch_fill<YI, V, true>(p, st) : launch_fill<YI, V, false>(p, st)
  if (g.y_is_i) {
    if (vec) SSB_DTW_DISPATCH(true, true); else SSB_DTW_DISPATCH(true, false);
  } else {
    if (vec) SSB_DTW_DISPATCH(false, true); else SSB_DTW_DISPATCH(false, false);
  }
#undef SSB_DTW_DISPATCH
  if (rc) return rc;

  if (g.y_is_i)
    return launch_backtrace<true>(p.dirs, g.dirs_pair_words, g.dirs_pair_words, g.nch, (int)N, (int)M,
                                  (int)npairs, path, nullptr, st);
  return launch_backtrace<false>(p.dirs, g.dirs_pair_words, g.dirs_pair_words, g.nch, (int)N, (int)M,
                                 (int)npairs, path, nullptr, st);
}

// ---- ragged batches --------------------------------------------------------------------------
// One launch aligns pairs of DIFFERENT shapes (every silent utterance of a real SizeAwareSampler
// batch has its own (T_target, T_pred), transduction_model.py:111-128).  Matrices are stored like
// the reference's costs tensor, (M = T_pred rows) x (N = T_target columns) row-major with a row
// pitch, and aligned as its `.T` view (stride_i == 1): DTW rows i = target frames.
int plan_pair(int64_t N, int64_t M, int64_t pitch, ssb_dtw_pair_t* t) {
  Geometry g;
  if (int rc = make_geometry(N, M, 1, M == 1 ? 1 : pitch, &g)) return rc;
  SSB_REQUIRE(g.y_is_i, "dtw ragged: internal orientation error");
  t->N = (int32_t)N; t->M = (int32_t)M; t->pitch = pitch;
  t->nbands = g.nbands; t->nch = g.nch;
  return SSB_OK;
}

// ---- fp64 path (interface parity) ------------------------------------------------------------
// align.py:6 allocates the accumulated-cost table with zeros_like(costs): the reference follows
// the caller's dtype, so a float64 matrix is accumulated and compared in float64.  Nothing on
// the training path does that (transduction_model.py:126 passes fp32), so this is a plain
// anti-diagonal wavefront, one CTA per pair, that writes the fp64 table (the caller's `dtw`
// buffer doubles as the workspace) and backtraces it with the reference's first-wins order.
__global__ void __launch_bounds__(256) dtw_f64_kernel(const double* __restrict__ cost_all,
                                                     double* __restrict__ dtw_all, int npairs,
                                                     int64_t pair_stride, int N, int M, int64_t si,
                                                     int64_t sj, int32_t* __restrict__ path_all) {
  const double INF = CUDART_INF;
  for (int pair = blockIdx.x; pair < npairs; pair += gridDim.x) {
    const double* cost = cost_all + (int64_t)pair * pair_stride;
    double* dtw = dtw_all + (int64_t)pair * pair_stride;
    for (int d = 0; d <= N + M - 2; ++d) {
      const int lo = max(0, d - (M - 1)), hi = min(N - 1, d);
      for (int i = lo + (int)threadIdx.x; i <= hi; i += blockDim.x) {
        const int j = d - i;
        double v;
        if (i == 0 && j == 0) {
          v = 0.0;
        } else if (i == 0 || j == 0) {
          v = INF;
        } else {
          const double up = dtw[(i - 1) * si + j * sj], left = dtw[i * si + (j - 1) * sj],
                       diag = dtw[(i - 1) * si + (j - 1) * sj];
          double m = up;
          if (left < m) m = left;
          if (diag < m) m = diag;
          v = cost[i * si + j * sj] + m;
        }
        dtw[i * si + j * sj] = v;
      }
      __syncthreads();   // block-wide visibility of this anti-diagonal's global stores
    }
    int32_t* path = path_all + (int64_t)pair * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) path[i] = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
      int i = N - 1, j = M - 1;
      while (i > 0 && j > 0) {    // align.py:24-26: min over [(i-1,j), (i,j-1), (i-1,j-1)], first wins
        path[i] = j;
        const double up = dtw[(i - 1) * si + j * sj], left = dtw[i * si + (j - 1) * sj],
                     diag = dtw[(i - 1) * si + (j - 1) * sj];
        int ni = i - 1, nj = j;
        double m = up;
        if (left < m) { m = left; ni = i; nj = j - 1; }
        if (diag < m) { ni = i - 1; nj = j - 1; }
        i = ni;
        j = nj;
      }
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" {

int ssb_dtw_time_warp_batch_f64(const double* cost, int64_t npairs, int64_t pair_stride, int64_t N,
                                int64_t M, int64_t stride_i, int64_t stride_j, double* dtw,
                                int32_t* path, void* stream) {
  SSB_REQUIRE(N >= 1 && M >= 1 && N < (1 << 30) && M < (1 << 30), "dtw: bad shape %lld x %lld",
              (long long)N, (long long)M);
  SSB_REQUIRE(stride_i >= 1 && stride_j >= 1, "dtw: strides must be positive");
  SSB_REQUIRE(npairs >= 0 && npairs < (1LL << 31), "dtw: bad npairs %lld", (long long)npairs);
  if (npairs == 0) return SSB_OK;
  SSB_REQUIRE(cost && dtw && path, "dtw: null pointer");
  const int grid = (int)(npairs < 8LL * ssb::num_sms() ? npairs : 8LL * ssb::num_sms());
  dtw_f64_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(cost, dtw, (int)npairs, pair_stride, (int)N,
                                                        (int)M, stride_i, stride_j, path);
  SSB_LAUNCH_CHECK("dtw_f64_kernel");
  return SSB_OK;
}


int64_t ssb_dtw_workspace_bytes(int64_t npairs, int64_t N, int64_t M, int64_t stride_i,
                                int64_t stride_j) {
  Geometry g;
  if (make_geometry(N, M, stride_i, stride_j, &g)) return SSB_ERR_ARG;
  if (npairs < 0) return SSB_ERR_ARG;
  const int64_t b = npairs * g.dirs_pair_words * 4;
  return b > 16 ? b : 16;
}

int ssb_dtw_align_batch(const float* cost, int64_t npairs, int64_t pair_stride, int64_t N,
                        int64_t M, int64_t stride_i, int64_t stride_j, int32_t* path,
                        void* workspace, int64_t workspace_bytes, void* stream) {
  return run(cost, npairs, pair_stride, N, M, stride_i, stride_j, nullptr, path, workspace,
             workspace_bytes, stream);
}

int ssb_dtw_time_warp_batch(const float* cost, int64_t npairs, int64_t pair_stride, int64_t N,
                            int64_t M, int64_t stride_i, int64_t stride_j, float* dtw,
                            int32_t* path, void* workspace, int64_t workspace_bytes,
                            void* stream) {
  SSB_REQUIRE(dtw != nullptr, "dtw: null dtw output");
  return run(cost, npairs, pair_stride, N, M, stride_i, stride_j, dtw, path, workspace,
             workspace_bytes, stream);
}

int64_t ssb_dtw_ragged_plan(int64_t npairs, const int64_t* N, const int64_t* M,
                            const int64_t* cost_off, const int64_t* pitch,
                            ssb_dtw_pair_t* table_out, int64_t* max_dims_out) {
  if (npairs < 0 || (npairs > 0 && (!N || !M || !cost_off || !pitch || !table_out))) {
    ssb::set_error("dtw ragged plan: null pointer / bad npairs");
    return SSB_ERR_ARG;
  }
  int64_t words = 0, maxN = 0, maxM = 0;
  for (int64_t i = 0; i < npairs; ++i) {
    ssb_dtw_pair_t* t = table_out + i;
    if (N[i] < 1 || M[i] < 1 || pitch[i] < N[i] || cost_off[i] < 0) {
      ssb::set_error("dtw ragged plan: pair %lld has N=%lld M=%lld pitch=%lld off=%lld",
                     (long long)i, (long long)N[i], (long long)M[i], (long long)pitch[i],
                     (long long)cost_off[i]);
      return SSB_ERR_ARG;
    }
    if (plan_pair(N[i], M[i], pitch[i], t)) return SSB_ERR_ARG;
    t->cost_off = cost_off[i];
    t->dirs_off = words;
    words += (int64_t)t->nbands * t->nch * BAND;
    maxN = N[i] > maxN ? N[i] : maxN;
    maxM = M[i] > maxM ? M[i] : maxM;
  }
  if (max_dims_out) { max_dims_out[0] = maxN; max_dims_out[1] = maxM; }
  const int64_t b = words * 4;
  return b > 16 ? b : 16;
}

int ssb_dtw_align_ragged(const float* cost_base, int64_t npairs, const ssb_dtw_pair_t* table_dev,
                         int64_t max_N, int64_t max_M, int vectorized, int32_t* path,
                         void* workspace, int64_t workspace_bytes, void* stream) {
  SSB_REQUIRE(npairs >= 0 && npairs < (1LL << 31), "dtw: bad npairs %lld", (long long)npairs);
  if (npairs == 0) return SSB_OK;
  SSB_REQUIRE(cost_base && table_dev && path && workspace, "dtw ragged: null pointer");
  SSB_REQUIRE(max_N >= 1 && max_M >= 1 && max_N < (1 << 24) && max_M < (1 << 24),
              "dtw ragged: bad maxima %lld x %lld", (long long)max_N, (long long)max_M);
  SSB_REQUIRE(((uintptr_t)workspace & 15) == 0, "dtw: workspace must be 16 B aligned");
  SSB_REQUIRE(!vectorized || ((uintptr_t)cost_base & 15) == 0,
              "dtw ragged: vectorized loads need a 16 B aligned base");
  (void)workspace_bytes;   // sized by ssb_dtw_ragged_plan; the table is trusted (caller-built)
  cudaStream_t st = (cudaStream_t)stream;
  DtwParams p;
  p.cost = cost_base;
  p.dtw_out = nullptr;
  p.dirs = (uint32_t*)workspace;
  p.table = table_dev;
  p.pair_stride = 0;
  p.pitch = 0;
  p.dirs_pair_words = 0;
  p.Ny = (int)max_N;
  p.Nx = (int)max_M;
  p.nbands = (int)((max_N + BAND - 1) / BAND);
  p.nch = (int)((max_M + 30) / CH + 1);
  p.npairs = (int)npairs;
  int rc = vectorized ? launch_fill<true, true, false, true>(p, st)
                      : launch_fill<true, false, false, true>(p, st);
  if (rc) return rc;
  return launch_backtrace<true>(p.dirs, 0, (int64_t)p.nbands * p.nch * BAND, p.nch, (int)max_N,
                                (int)max_M, (int)npairs, path, table_dev, st);
}

}  // extern "C"
