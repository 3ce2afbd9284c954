// decode.cu -- SuperPoint decode kernels (sm_100a).
//
// Replaces SuperPointFeatureFrontEnd::postprocessDetectionAndDescription() and its helpers
// (reference: src/odml_visual_odometry/src/feature_detection_neural_network.cpp, "NN" below):
//   k_softmax_heat  NN:266-326  exp, channel sum (c = 0..64 in order), /(sum + 1e-5), dustbin drop,
//                               depth-to-space to the H x W heatmap; plus a per-cell (max, second max, argmax)
//                               record from which k_detect finds its candidates without scanning the heatmap.
//   k_detect        NN:188-262  strict '>' threshold, descending-score order with the canonical
//                               tie-break (score desc, x asc, y asc), greedy box NMS, border filter,
//                               stop at K emitted.  Exact: candidates are consumed in descending key
//                               order chunk by chunk, so the result equals a full sort + sequential walk.
//   k_sample_desc   NN:332-431  align-corners bilinear sampling of the coarse 256-d map + L2 normalise.
//
// HBM-bound integer/float streaming work: coalesced 128 B channel-plane reads of semi, 32 B-aligned
// float4 heatmap stores, warp-shuffle reductions; no tensor cores here.
#include <cstring>
#include <stdlib.h>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace spvo {

// Network outputs arrive as fp32 (the reference's layout, HPP:382-384) or as fp16 (an fp16 TensorRT engine's
// bindings).  fp16 -> fp32 is exact, so everything downstream is the same arithmetic on the same values.
__device__ __forceinline__ float ld_in(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_in(const __half* p) { return __half2float(__ldg(p)); }
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }

// ------------------------------------------------------------------------------------------------
// K1: softmax, reduced to one RECORD per 8x8 cell.  One THREAD per cell, 128 consecutive cells per block.
//   * every global load is one 128 B coalesced request per warp (a channel plane, 32 consecutive cells); the loads
//     of a thread are independent, so a warp keeps several KB in flight;
//   * the 65-term channel sum must be accumulated in channel order (oracle: s = s + e_c, c = 0..64): it is a
//     plain sequential chain inside the thread -- no shared memory, no barriers;
//   * the heatmap itself (NN:266-326: p_c = e_c / (sum + 1e-5), channel c = pixel (c / 8, c % 8) of the cell) is not
//     an output of the path, and k_detect only ever needs the pixels of the cells that hold two or more candidates.
//     A cell's record is its largest and second largest heat value: RN(e / denom) is monotone in e, so they are the
//     quotients of the two largest exponentials -- two divisions per cell instead of 64.  The 64 heat values are
//     stored only for cells whose second value reaches a per-image-slot threshold derived from what k_detect left
//     behind on its previous call (0.8 x the score of its K-th keypoint): a PREDICTION of the cells it will open.  A
//     wrong prediction costs time, never correctness: k_detect recomputes any unstored cell it needs from the
//     logits (visit_cells, same arithmetic), 65 scattered sectors instead of 8.
// ------------------------------------------------------------------------------------------------
constexpr int kHeatThreads = 128;

// e / denom, correctly rounded (== __fdiv_rn, the oracle's `/`), from ONE reciprocal per cell.  Markstein's
// sequence: with y = RN(1/b), q0 = RN(a y) is within 1.5 ulp of a/b; q1 = RN(q0 + RN(a - b q0) y) is a faithful
// quotient (the remainder is exact in an FMA); one more correction from a faithful quotient with a correctly rounded
// reciprocal yields RN(a/b) (Markstein 1990; Muller et al., Handbook of Floating-Point Arithmetic, "division with an
// FMA").  5 FMA-pipe instructions instead of the ~10 (one of them a MUFU) of a stand-alone IEEE division.  Valid
// when nothing under- or overflows: the caller guards the denominator's range per cell, and quotients below 2^-90
// (whose remainders may underflow) stay far below any confidence threshold the fast path is used with.
__device__ __forceinline__ float div_by_rcp(float a, float b, float y) {
  const float q0 = __fmul_rn(a, y);
  const float q1 = __fmaf_rn(__fmaf_rn(-q0, b, a), y, q0);
  return __fmaf_rn(__fmaf_rn(-q1, b, a), y, q1);
}

// Is a cell's 8x8 block of heat values stored ("spilled") by k_softmax_heat?  rec_y = the record's second word;
// both kernels evaluate exactly this predicate with the same per-image threshold.
__device__ __forceinline__ bool cell_spilled(uint32_t rec_y, uint32_t thr_bits, uint32_t conf_bits) {
  const uint32_t sb = rec_y | 63u;  // upper bound of the second largest pixel
  return sb >= thr_bits && sb > conf_bits;
}

// Heat values of the marked cells of a warp (bit l of `marked`: lane l's cell), kLanes lanes per cell, 64 / kLanes
// consecutive pixels each.  cell / denom: each lane's own cell index and softmax denominator.
template <int kLanes, typename T>
__device__ __forceinline__ void store_heat_cells(unsigned marked, const T* __restrict__ img, float* __restrict__ heat_img,
                                                 int cell, float denom, int cells, int Wc, int fast_div) {
  constexpr int kPer = 64 / kLanes;      // pixels per lane: 2 (a quarter of a heat row) or 8 (a heat row)
  constexpr int kGroups = 32 / kLanes;   // cells per round
  const int lane = threadIdx.x & 31, grp = lane / kLanes, sub = lane % kLanes;
  const int W = Wc * 8;
  for (unsigned m = marked; m;) {
    unsigned mine = m;  // the grp-th lowest marked lane of this round
#pragma unroll
    for (int g = 1; g < kGroups; ++g)
      if (grp >= g) mine &= mine - 1;
    const bool have = mine != 0u;
    const int from = have ? __ffs(mine) - 1 : 0;
    const int cl = __shfl_sync(0xffffffffu, cell, from);
    const float dn = __shfl_sync(0xffffffffu, denom, from);
    if (have) {
      float e[kPer];
#pragma unroll
      for (int k = 0; k < kPer; ++k) e[k] = ld_in(img + (size_t)(kPer * sub + k) * cells + cl);
#pragma unroll
      for (int k = 0; k < kPer; ++k) e[k] = spvo_exp(e[k]);
      // shared-reciprocal division when the denominator is comfortably inside the normal range (always, for finite
      // network outputs of sane magnitude); otherwise the stand-alone IEEE division
      if (fast_div && dn > 0x1p-60f && dn < 0x1p60f) {
        const float rcp = __frcp_rn(dn);
#pragma unroll
        for (int k = 0; k < kPer; ++k) e[k] = div_by_rcp(e[k], dn, rcp);
      } else {
#pragma unroll
        for (int k = 0; k < kPer; ++k) e[k] = __fdiv_rn(e[k], dn);
      }
      const int hc = cl / Wc, wc = cl - hc * Wc;
      const int px = kPer * sub;  // first pixel (8 * row + col) of this lane
      float* dst = heat_img + (size_t)(8 * hc + (px >> 3)) * W + 8 * wc + (px & 7);
      if constexpr (kPer == 2) {
        *reinterpret_cast<float2*>(dst) = make_float2(e[0], e[1]);
      } else {
        reinterpret_cast<float4*>(dst)[0] = make_float4(e[0], e[1], e[2], e[3]);
        reinterpret_cast<float4*>(dst)[1] = make_float4(e[4], e[5], e[6], e[7]);
      }
    }
#pragma unroll
    for (int g = 0; g < kGroups; ++g) m &= m - 1;
  }
}

// Per image slot the handle remembers ONE number between calls: the score of the K-th keypoint k_detect emitted last
// time (the confidence threshold if it found fewer than K).  Both predictions derive from it -- a measured property of
// the previous image, so there is no feedback: cells are stored when their second value reaches 0.8 x it, and
// k_detect's first generation gathers the candidates above 0.9 x it instead of estimating a bound from the
// histogram of cell maxima.
__device__ __forceinline__ uint32_t store_threshold(uint32_t needed_bits) {
  return fbits(__fmul_rn(0.8f, __uint_as_float(needed_bits)));
}

template <typename T>
__global__ void __launch_bounds__(kHeatThreads, 4)
k_softmax_heat(const T* __restrict__ semi, float* __restrict__ heat, uint2* __restrict__ cellmax,
               const uint32_t* __restrict__ spill_thr, int b0, uint32_t conf_bits, int Hc, int Wc, int fast_div) {
  chain_enter();
  const int b = blockIdx.y;
  const int cells = Hc * Wc;
  const int cell = blockIdx.x * kHeatThreads + threadIdx.x;
  const T* img = semi + (size_t)b * 65 * cells;
  bool store = false;
  float denom = 1.0f;
  if (cell < cells) {
    const T* src = img + cell;
    float s = 0.0f;
    float m1 = 0.0f, m2 = 0.0f;  // largest and second largest exponential of the cell (m2 == m1 on ties)
    int arg = 0;                 // pixel index 8 * row + col of the largest
    float x[64];                 // every channel of the cell in flight (8 KB per warp)
#pragma unroll
    for (int u = 0; u < 64; ++u) x[u] = ld_in(src + (size_t)u * cells);
#pragma unroll
    for (int u = 0; u < 64; u += 2) {
      float e[2];
      spvo_exp_x2(x[u], x[u + 1], e[0], e[1]);
#pragma unroll
      for (int v = 0; v < 2; ++v) {
        s = __fadd_rn(s, e[v]);
        m2 = fmaxf(m2, fminf(m1, e[v]));
        arg = e[v] > m1 ? u + v : arg;
        m1 = fmaxf(m1, e[v]);
      }
    }
    s = __fadd_rn(s, spvo_exp(ld_in(src + (size_t)64 * cells)));
    denom = __fadd_rn(s, 0.00001f);
    // Per-cell record for k_detect: .x = bits of the cell maximum, .y = bits of the second largest pixel with the
    // low 6 bits replaced by the argmax (meaningful when the maximum is unique -- the only case k_detect uses it).
    // Cells that cannot hold a candidate of the current range are skipped there; cells whose second pixel is below
    // the range yield their single candidate straight from the record.
    const float p1 = __fdiv_rn(m1, denom), p2 = __fdiv_rn(m2, denom);
    const uint2 rec = make_uint2(fbits(p1), (fbits(p2) & ~63u) | (uint32_t)arg);
    cellmax[(size_t)b * cells + cell] = rec;
    store = cell_spilled(rec.y, store_threshold(__ldg(spill_thr + b0 + b)), conf_bits);
  }
  // Store the 64 heat values of the cells k_detect is expected to open (two or more candidates above the bound its
  // walk reached on this image slot last time).  The logits are re-read (this warp loaded their sectors a moment ago)
  // so that the main loop does not have to keep 64 exponentials alive.  No shared memory, no block barrier: a warp
  // with one or two marked cells takes them one at a time (lane l: pixels 2l, 2l+1), a warp with more takes four at
  // a time (8 lanes per cell, one heat row each, 8 loads in flight per lane).
  const unsigned marked = __ballot_sync(0xffffffffu, store);
  if (marked == 0u) return;
  float* hb = heat + (size_t)b * (size_t)(Hc * 8) * (Wc * 8);
  if (__popc(marked) <= 2) store_heat_cells<32>(marked, img, hb, cell, denom, cells, Wc, fast_div);
  else store_heat_cells<8>(marked, img, hb, cell, denom, cells, Wc, fast_div);
}

// Self-check used by tests (spvo_debug_div_check): counts operand pairs on which the shared-reciprocal division
// differs from __fdiv_rn.  a_bits / b_bits: n fp32 bit patterns each.
__global__ void k_div_check(const uint32_t* __restrict__ a_bits, const uint32_t* __restrict__ b_bits, long long n,
                            unsigned long long* __restrict__ mismatches) {
  unsigned long long bad = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float a = __uint_as_float(a_bits[i]), b = __uint_as_float(b_bits[i]);
    if (!(b > 0x1p-60f && b < 0x1p60f) || !(a >= 0.0f) || !(a < 0x1p100f)) continue;  // outside the guarded domain
    const float ref = __fdiv_rn(a, b);
    if (ref != 0.0f && ref < 0x1p-90f) continue;  // documented: tiny quotients are not covered
    const float got = div_by_rcp(a, b, __frcp_rn(b));
    if (__float_as_uint(got) != __float_as_uint(ref)) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}
cudaError_t launch_div_check(Handle* h, const uint32_t* a_bits, const uint32_t* b_bits, long long n,
                             unsigned long long* mismatches) {
  k_div_check<<<h->sm_count * 8, 256, 0, h->stream>>>(a_bits, b_bits, n, mismatches);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K2: per-image detect.  One CTA per image.
// Candidate key (64 bit, unique per pixel): (score_bits << 32) | (0xFFFFFFFF - (x*H + y)).
// Descending key order == score descending, then x ascending, then y ascending == a stable sort of
// the reference's column-major candidate list (NN:205-217).
// ------------------------------------------------------------------------------------------------
struct DetectParams {
  const float* heat;     // [B, H, W] heat values of the cells k_softmax_heat chose to store (cell_spilled)
  uint32_t* spill_thr;   // [B] per image slot: score bits of the K-th keypoint of the previous call; rewritten for the next
  const void* semi;      // [B, 65, cells] detector logits (fp32 or fp16): other multi-candidate cells are recomputed
  int semi_f16;
  int fast_div;          // shared-reciprocal division allowed (conf_thresh >= 1e-20)
  const uint2* cellmax;  // [B, cells] per-cell (max bits, second-max bits | argmax) records of the heatmap
  unsigned long long* list;  // [B, kListCap] global scratch: candidate keys of the current generation
  int H, W;
  float conf;
  int dist, border, K;
  spvo_keypoint* kpts;
  float* scores;
  int* n_out;
  unsigned long long* counters;
  int4* kp_par;   // [B,K] sampling parameters for k_desc_planes (may be NULL)
  unsigned* opmax_zero;  // [B] matcher operand max-norm slots to clear for k_desc_normalize (may be NULL)
  size_t bitmap_stride;  // words per image in `bitmap`
  unsigned* bitmap;  // [B, bitmap_stride] global scratch: pixels suppressed by earlier chunks (multi-chunk path only)
  int cap;        // key buffer capacity (power of two)
  int target;     // cells (generation) / candidates (first chunk) wanted
};

constexpr int kDetectThreads = 512;
// A heat value is a candidate iff it compares '>' conf as a float (NN:203): NaN (non-finite logits) never is, although
// its bit pattern is above every finite one -- the raw-bit predicates exclude everything from +inf upwards.
constexpr uint32_t kInfBits = 0x7F800000u;
typedef unsigned long long u64;

__device__ __forceinline__ u64 make_key(uint32_t bits, int x, int y, int H) {
  return ((u64)bits << 32) | (u64)(0xFFFFFFFFu - (uint32_t)(x * H + y));
}

#ifdef SPVO_PHASE_TIMING
// Diagnostic build only (scripts/build_diag.sh): cycles spent in each phase of k_detect, accumulated over the chunks and
// generations of one launch, for the first 64 images.  PHASE(i) closes the interval since the previous PHASE and adds
// it to slot i.
__device__ long long g_phase_clk[64 * 32];
__device__ long long g_phase_last[64];
#define g_phase_clk_chunks() (blockIdx.x < 64 ? g_phase_clk[blockIdx.x * 32 + 12] : 0)
#define PHASE_INIT() do { if (threadIdx.x == 0 && blockIdx.x < 64) { for (int q_ = 0; q_ < 32; ++q_) g_phase_clk[blockIdx.x * 32 + q_] = 0; g_phase_last[blockIdx.x] = clock64(); } } while (0)
#define PHASE(i) do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.x < 64) { const long long t_ = clock64(); g_phase_clk[blockIdx.x * 32 + (i)] += t_ - g_phase_last[blockIdx.x]; g_phase_last[blockIdx.x] = t_; } } while (0)
#define PHASE_COUNT(i) do { if (threadIdx.x == 0 && blockIdx.x < 64) g_phase_clk[blockIdx.x * 32 + (i)] += 1; } while (0)
#define PHASE_NOTE(i, v) do { if (threadIdx.x == 0 && blockIdx.x < 64 && (i) < 32) g_phase_clk[blockIdx.x * 32 + (i)] = (v); } while (0)
#else
#define PHASE_INIT() do { } while (0)
#define PHASE(i) do { } while (0)
#define PHASE_COUNT(i) do { } while (0)
#define PHASE_NOTE(i, v) do { } while (0)
#define g_phase_clk_chunks() 0
#endif
// The detector logits of one image, [65][cells], fp32 or fp16.
struct SemiView {
  const void* base;
  int f16, cells, Wc;
  int fast_div;
};
__device__ __forceinline__ float semi_at(const SemiView& sv, int c, int cell) {
  const size_t o = (size_t)c * sv.cells + cell;
  return sv.f16 ? __half2float(__ldg(reinterpret_cast<const __half*>(sv.base) + o))
                : __ldg(reinterpret_cast<const float*>(sv.base) + o);
}

// Heat values of a set of cells, recomputed from the logits with k_softmax_heat's arithmetic (NN:266-326: exp,
// channel sum in channel order, + 1e-5, one IEEE division per pixel).  The cells are cell_list[0 .. ncell) or, when
// cell_list is null, the cells 0 .. ncell-1.  kStageCells cells per round: all threads first fill a [65][64] table
// of exponentials in shared memory (a warp reads one channel of 32 listed cells: coalesced where the cells are
// neighbours), then thread (j, r) = (tid % 64, tid / 64) sums cell j's column in order and divides the 8 values of
// heat row r.  fn(ok, bits[8], x0, y) is called once per round by EVERY thread (it may use warp collectives):
// bits[k] = heat(y, x0 + k) when ok.
constexpr int kStageCells = 64;
constexpr int kStageFloats = 65 * kStageCells;
static_assert(kDetectThreads == 8 * kStageCells, "one thread per (cell, heat row)");
constexpr int kStageIt = (65 + 7) / 8;  // channels r, r + 8, ... of a cell: 9 independent loads per thread
template <typename T>
__device__ __forceinline__ void load_logits(const T* __restrict__ base, int cells, int cell, int r, bool ok, float (&x)[kStageIt]) {
#pragma unroll
  for (int it = 0; it < kStageIt; ++it) {
    const int c = r + 8 * it;
    x[it] = (ok && c < 65) ? ld_in(base + (size_t)c * cells + cell) : 0.0f;
  }
}
template <class Fn>
__device__ __forceinline__ void visit_cells(const SemiView& sv, const uint16_t* cell_list, int ncell, float* stage, Fn fn) {
  const int tid = threadIdx.x, j = tid & (kStageCells - 1), r = tid >> 6;
  auto cell_of = [&](int i) { return i < ncell ? (cell_list ? (int)cell_list[i] : i) : -1; };
  auto load = [&](int cell, float (&x)[kStageIt]) {  // the element type is tested once per round, not per load
    if (sv.f16) load_logits(reinterpret_cast<const __half*>(sv.base), sv.cells, cell, r, cell >= 0, x);
    else load_logits(reinterpret_cast<const float*>(sv.base), sv.cells, cell, r, cell >= 0, x);
  };
  float x[kStageIt];
  int cell = cell_of(j);
  load(cell, x);
  for (int base = 0; base < ncell; base += kStageCells) {
    const bool ok = cell >= 0;
#pragma unroll
    for (int it = 0; it < kStageIt; ++it) {
      const int c = r + 8 * it;
      if (c < 65) stage[c * kStageCells + j] = spvo_exp(x[it]);
    }
    __syncthreads();
    // the next round's logits travel while this round is summed, divided and filtered
    const int cell_next = cell_of(base + kStageCells + j);
    if (base + kStageCells < ncell) load(cell_next, x);
    uint32_t bits[8];
    int x0 = 0, y = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) bits[k] = 0u;
    if (ok) {
      float s = 0.0f;
#pragma unroll 13
      for (int c = 0; c < 65; ++c) s = __fadd_rn(s, stage[c * kStageCells + j]);
      const float denom = __fadd_rn(s, 0.00001f);
      const int hc = cell / sv.Wc, wc = cell - hc * sv.Wc;
      x0 = 8 * wc;
      y = 8 * hc + r;
      // shared-reciprocal division when the denominator is comfortably inside the normal range (always, for finite
      // network outputs of sane magnitude); otherwise the stand-alone IEEE division
      if (sv.fast_div && denom > 0x1p-60f && denom < 0x1p60f) {
        const float rcp = __frcp_rn(denom);
#pragma unroll
        for (int k = 0; k < 8; ++k) bits[k] = fbits(div_by_rcp(stage[(8 * r + k) * kStageCells + j], denom, rcp));
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) bits[k] = fbits(__fdiv_rn(stage[(8 * r + k) * kStageCells + j], denom));
      }
    }
    fn(ok, bits, x0, y);
    cell = cell_next;
    __syncthreads();  // the table is refilled by the next round
  }
}

// Score bin of a heat value: bin(v) = (bits(1.0f) - bits(v)) >> 14, clamped to [0, 4095]; bin 0 holds the highest
// scores.  Used only to SIZE candidate chunks -- exactness never depends on it.
__device__ __forceinline__ int score_bin(uint32_t bits) {
  return bits >= kOneBits ? 0 : (int)min((kOneBits - bits) >> kHistShift, (uint32_t)(kHistBins - 1));
}

// Candidate list of one image in global memory (L2-resident): every candidate key in [lo, hi), gathered ONCE through
// the per-cell records, plus an exact histogram of their score bins in shared memory.  The chunks of the greedy walk
// are then cut from this list (one coalesced pass over <= kListCap keys per chunk) instead of re-scanning the heatmap.
//   * a cell whose maximum is below lo holds nothing;
//   * a cell whose maximum lies in [lo, hi) and whose second largest pixel is below lo yields its single candidate
//     straight from the record;
//   * every other contributing cell (second largest pixel >= lo) has its 64 heat values recomputed from the
//     logits (visit_cells) and filtered.
// (second | 63) bounds the second largest pixel from above.  Returns the number of keys in [lo, hi) (the first
// `list_cap` are stored); bins[] is incremented for every one of them.
constexpr int kListCap = 16384;

// Stored cells: half a warp per cell reads its 8 rows x 32 B of the heatmap.
__device__ void fetch_spilled_cells(const float* __restrict__ heat, int H, int W, uint32_t conf_bits, u64 lo, u64 hi,
                                   const uint16_t* cell_list, int ncell, u64* __restrict__ list, int list_cap,
                                   unsigned* bins, int* s_count) {
  const int Wc = W >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lo_b = (uint32_t)(lo >> 32), hi_b = (uint32_t)(hi >> 32);
  // half a warp per cell: lane l16 reads row l16/2, float4 l16&1 of the cell's 8x8 block
  const int l16 = lane & 15, half = lane >> 4, row = l16 >> 1, part = l16 & 1;
  constexpr int kCU = 4;  // cells (16-byte loads) in flight per thread
  const int per_iter = (kDetectThreads / 32) * 2 * kCU;
  for (int base = 0; base < ncell; base += per_iter) {
    float4 v[kCU];
    int cx[kCU], cy[kCU];
    bool ok[kCU];
#pragma unroll
    for (int u = 0; u < kCU; ++u) {
      const int ci = base + (u * (kDetectThreads / 32) + warp) * 2 + half;
      ok[u] = ci < ncell;
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      cx[u] = cy[u] = 0;
      if (ok[u]) {
        const int c = cell_list[ci];
        const int hc = c / Wc, wc = c - hc * Wc;
        cy[u] = 8 * hc + row;
        cx[u] = 8 * wc + 4 * part;
        v[u] = __ldg(reinterpret_cast<const float4*>(heat + (size_t)cy[u] * W + cx[u]));
      }
    }
    // Hits are rare (a few pixels per fetched cell), so each thread first marks its hits in a 32-bit mask, the warp
    // reserves its key slots with ONE scan + ONE shared atomic, and only then are the keys built and stored.
    uint32_t mask = 0u;
#pragma unroll
    for (int u = 0; u < kCU; ++u) {
      const float pv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t bts = fbits(pv[e]);
        bool s = ok[u] && bts > conf_bits && bts >= lo_b && bts <= hi_b && bts < kInfBits;
        if (s && (bts == lo_b || bts == hi_b)) {  // boundary score: full key comparison
          const u64 key = make_key(bts, cx[u] + e, cy[u], H);
          s = key >= lo && key < hi;
        }
        mask |= (s ? 1u : 0u) << (4 * u + e);
      }
    }
    const int cnt = __popc(mask);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    int basei = 0;
    if (lane == 31 && inc > 0) basei = atomicAdd(s_count, inc);
    basei = __shfl_sync(0xffffffffu, basei, 31);
    int slot = basei + inc - cnt;
    if (mask) {
#pragma unroll
      for (int u = 0; u < kCU; ++u) {
        const float pv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (mask & (1u << (4 * u + e))) {
            const uint32_t bts = fbits(pv[e]);
            if (slot < list_cap) list[slot] = make_key(bts, cx[u] + e, cy[u], H);
            atomicAdd(&bins[score_bin(bts)], 1u);
            ++slot;
          }
        }
      }
    }
  }
}

// Cells that were not stored: recomputed from the logits.
__device__ void fetch_listed_cells(const SemiView& sv, int H, uint32_t conf_bits, u64 lo, u64 hi,
                                   const uint16_t* cell_list, int ncell, u64* __restrict__ list, int list_cap,
                                   unsigned* bins, int* s_count, float* stage) {
  const int lane = threadIdx.x & 31;
  const uint32_t lo_b = (uint32_t)(lo >> 32), hi_b = (uint32_t)(hi >> 32);
  visit_cells(sv, cell_list, ncell, stage, [&](bool ok, const uint32_t (&bits)[8], int x0, int y) {
    // Hits are rare (a few pixels per cell), so each thread first marks its hits in a mask, the warp reserves its
    // key slots with ONE scan + ONE shared atomic, and only then are the keys built and stored.
    uint32_t mask = 0u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t bts = bits[k];
      bool s = ok && bts > conf_bits && bts >= lo_b && bts <= hi_b && bts < kInfBits;
      if (s && (bts == lo_b || bts == hi_b)) {  // boundary score: full key comparison
        const u64 key = make_key(bts, x0 + k, y, H);
        s = key >= lo && key < hi;
      }
      mask |= (s ? 1u : 0u) << k;
    }
    const int cnt = __popc(mask);
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    int basei = 0;
    if (lane == 31 && inc > 0) basei = atomicAdd(s_count, inc);
    basei = __shfl_sync(0xffffffffu, basei, 31);
    int slot = basei + inc - cnt;
    if (mask) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (mask & (1u << k)) {
          if (slot < list_cap) list[slot] = make_key(bits[k], x0 + k, y, H);
          atomicAdd(&bins[score_bin(bits[k])], 1u);
          ++slot;
        }
      }
    }
  });
}

__device__ int collect_to_list(const SemiView& sv, const float* __restrict__ heat, uint32_t thr_bits,
                               const uint2* __restrict__ cellmax, int H, int W,
                               uint32_t conf_bits, u64 lo, u64 hi, u64* __restrict__ list, int list_cap,
                               unsigned* bins, uint16_t* cell_list, int cl_cap, int* s_count, int* s_ncell,
                               int* s_ncell2, int* s_gencells, float* stage) {
  const int Wc = W >> 3, cells = (H >> 3) * Wc;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    *s_count = 0;
    *s_ncell = 0;
    *s_ncell2 = 0;
    *s_gencells = 0;
  }
  __syncthreads();
  const uint32_t lo_b = (uint32_t)(lo >> 32);
  constexpr int kPU = 4;  // cell records in flight per thread
  for (int c0 = 0; c0 < cells; c0 += kPU * kDetectThreads) {
    uint2 rec[kPU];
#pragma unroll
    for (int u = 0; u < kPU; ++u) {
      const int c = c0 + u * kDetectThreads + threadIdx.x;
      rec[u] = c < cells ? __ldg(cellmax + c) : make_uint2(0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < kPU; ++u) {
      const int c = c0 + u * kDetectThreads + threadIdx.x;
      const uint32_t mb = rec[u].x;
      const bool q = c < cells && mb > conf_bits && mb >= lo_b && mb < kInfBits;
      {  // cells whose maximum lies in the range (sizes the first chunk of the walk)
        const unsigned mq = __ballot_sync(0xffffffffu, q);
        if (mq && lane == 0) atomicAdd(s_gencells, __popc(mq));
      }
      const uint32_t sb = rec[u].y | 63u;
      const bool multi = q && sb > conf_bits && sb >= lo_b;
      bool single = q && !multi;
      u64 key = 0ull;
      if (single) {
        const int hc = c / Wc, wc = c - hc * Wc, a = (int)(rec[u].y & 63u);
        key = make_key(mb, 8 * wc + (a & 7), 8 * hc + (a >> 3), H);
        single = key >= lo && key < hi;
      }
      // stored cells are listed from the front of cell_list, the others from its back (never colliding: flushed below)
      const bool stored = multi && cell_spilled(rec[u].y, thr_bits, conf_bits);
      const unsigned mm = __ballot_sync(0xffffffffu, stored);
      if (mm) {
        int basei = 0;
        const int leader = __ffs(mm) - 1;
        if (lane == leader) basei = atomicAdd(s_ncell, __popc(mm));
        basei = __shfl_sync(0xffffffffu, basei, leader);
        if (stored) cell_list[basei + __popc(mm & ((1u << lane) - 1u))] = (uint16_t)c;
      }
      const unsigned mu = __ballot_sync(0xffffffffu, multi && !stored);
      if (mu) {
        int basei = 0;
        const int leader = __ffs(mu) - 1;
        if (lane == leader) basei = atomicAdd(s_ncell2, __popc(mu));
        basei = __shfl_sync(0xffffffffu, basei, leader);
        if (multi && !stored) cell_list[cl_cap - 1 - (basei + __popc(mu & ((1u << lane) - 1u)))] = (uint16_t)c;
      }
      const unsigned ms = __ballot_sync(0xffffffffu, single);
      if (ms) {
        int basei = 0;
        const int leader = __ffs(ms) - 1;
        if (lane == leader) basei = atomicAdd(s_count, __popc(ms));
        basei = __shfl_sync(0xffffffffu, basei, leader);
        if (single) {
          const int slot = basei + __popc(ms & ((1u << lane) - 1u));
          if (slot < list_cap) list[slot] = key;
          atomicAdd(&bins[score_bin(mb)], 1u);
        }
      }
    }
    __syncthreads();
    // fetch the listed cells before the list can overflow (a block of records adds at most kPU * threads cells)
    // every thread must see the SAME counts: the next round's atomics may not start before all have read them
    const int ncell = *s_ncell, ncell2 = *s_ncell2;
    __syncthreads();
    if (ncell + ncell2 + kPU * kDetectThreads > cl_cap || c0 + kPU * kDetectThreads >= cells) {
      PHASE(2);  // G2: records
      fetch_spilled_cells(heat, H, W, conf_bits, lo, hi, cell_list, ncell, list, list_cap, bins, s_count);
      PHASE(14);  // G2: stored cells
      fetch_listed_cells(sv, H, conf_bits, lo, hi, cell_list + cl_cap - ncell2, ncell2, list, list_cap, bins, s_count, stage);
      PHASE(15);  // G2: recomputed cells
      __syncthreads();
      if (threadIdx.x == 0) {
        *s_ncell = 0;
        *s_ncell2 = 0;
      }
      __syncthreads();
    }
  }
  return *s_count;
}

// Keys of list[0 .. n_list) inside [lo, hi) -> keys[] (first `cap` stored); returns their number.
__device__ int select_from_list(const u64* __restrict__ list, int n_list, u64 lo, u64 hi, u64* keys, int cap,
                                int* s_count) {
  if (threadIdx.x == 0) *s_count = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  constexpr int kLU = 4;  // list loads in flight per thread (the list lives in L2)
  for (int i0 = 0; i0 < n_list; i0 += kLU * kDetectThreads) {
    u64 kk[kLU];
#pragma unroll
    for (int u = 0; u < kLU; ++u) {
      const int i = i0 + u * kDetectThreads + threadIdx.x;
      kk[u] = i < n_list ? list[i] : 0ull;  // plain load: the list was written by this CTA
    }
#pragma unroll
    for (int u = 0; u < kLU; ++u) {
      const u64 key = kk[u];
      const bool s = key >= lo && key < hi && key != 0ull;
      const unsigned m = __ballot_sync(0xffffffffu, s);
      if (m) {
        int basei = 0;
        const int leader = __ffs(m) - 1;
        if (lane == leader) basei = atomicAdd(s_count, __popc(m));
        basei = __shfl_sync(0xffffffffu, basei, leader);
        if (s) {
          const int slot = basei + __popc(m & ((1u << lane) - 1u));
          if (slot < cap) keys[slot] = key;
        }
      }
    }
  }
  __syncthreads();
  return *s_count;
}

// Exact radix select over the candidate list: the m-th largest key among list keys < hi, or `floor_key` when fewer
// than m remain.  8 passes of 8 bits over <= kListCap keys (only when one score bin alone overflows a chunk).
__device__ u64 radix_select_list(const u64* __restrict__ list, int n_list, u64 hi, int m, u64 floor_key,
                                 unsigned* s_hist, u64* s_prefix, int* s_want) {
  if (threadIdx.x == 0) {
    *s_prefix = 0;
    *s_want = m;
  }
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += kDetectThreads) s_hist[i] = 0;
    __syncthreads();
    const u64 prefix = *s_prefix;
    for (int i = threadIdx.x; i < n_list; i += kDetectThreads) {
      const u64 key = list[i];
      if (key < hi && (pass == 0 || (key >> (shift + 8)) == prefix)) atomicAdd(&s_hist[(unsigned)(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int want = *s_want, cum = 0, d = 255;
      bool found = false;
      for (; d >= 0; --d) {
        const int c = (int)s_hist[d];
        if (cum + c >= want) {
          found = true;
          break;
        }
        cum += c;
      }
      if (!found) {
        *s_want = -1;
      } else {
        *s_want = want - cum;
        *s_prefix = (prefix << 8) | (u64)d;
      }
    }
    __syncthreads();
    if (*s_want < 0) return floor_key;
  }
  const u64 thr = *s_prefix;
  return thr < floor_key ? floor_key : thr;
}

// Exact radix select: returns the m-th largest key among candidate keys < hi (unique keys), or
// `floor_key` when fewer than m remain.  8 passes of 8 bits, each recomputing every cell of the image (slow path
// only: one score bin alone overflows the candidate list, i.e. > 12 k exactly tied pixels).
__device__ u64 radix_select(const SemiView& sv, int H, uint32_t conf_bits, u64 hi, int m, u64 floor_key,
                            unsigned* s_hist, u64* s_prefix, int* s_want, float* stage) {
  if (threadIdx.x == 0) {
    *s_prefix = 0;
    *s_want = m;
  }
  const uint32_t hi_b = (uint32_t)(hi >> 32);
  for (int pass = 0; pass < 8; ++pass) {
    const int shift = 56 - 8 * pass;
    for (int i = threadIdx.x; i < 256; i += kDetectThreads) s_hist[i] = 0;
    __syncthreads();
    const u64 prefix = *s_prefix;
    visit_cells(sv, nullptr, sv.cells, stage, [&](bool ok, const uint32_t (&bits)[8], int x0, int y) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t b = bits[k];
        if (ok && b > conf_bits && b <= hi_b && b < kInfBits) {
          const u64 key = make_key(b, x0 + k, y, H);
          if (key < hi && (pass == 0 || (key >> (shift + 8)) == prefix))
            atomicAdd(&s_hist[(unsigned)(key >> shift) & 255u], 1u);
        }
      }
    });
    __syncthreads();
    if (threadIdx.x == 0) {
      int want = *s_want, cum = 0, d = 255;
      bool found = false;
      for (; d >= 0; --d) {
        int c = (int)s_hist[d];
        if (cum + c >= want) {
          found = true;
          break;
        }
        cum += c;
      }
      if (!found) {
        *s_want = -1;  // fewer than m keys remain
      } else {
        *s_want = want - cum;
        *s_prefix = (prefix << 8) | (u64)d;
      }
    }
    __syncthreads();
    if (*s_want < 0) return floor_key;
  }
  u64 thr = *s_prefix;
  return thr < floor_key ? floor_key : thr;
}

// In-place bitonic sort, descending, n_pad a power of two >= 32 (padding keys are 0 < any valid key).
// Only the stages whose partner distance j reaches another warp (j >= 128) go through shared memory with a
// block barrier each; for j <= 64 a thread holds 4 consecutive keys in registers and exchanges with the
// lane j / 4 away by shuffle (j = 2, 1 stay inside the thread): 21 barriers instead of 66 for 2048 keys.
__device__ void bitonic_sort_desc(u64* keys, int n_pad) {
  const int tid = threadIdx.x, lane = tid & 31;
  for (int k = 2; k <= n_pad; k <<= 1) {
    int j = k >> 1;
    for (; j >= 128; j >>= 1) {
      for (int t = tid; t < (n_pad >> 1); t += kDetectThreads) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const u64 a = keys[i], c = keys[i + j];
        const bool up = (i & k) == 0;  // descending run
        if ((a < c) == up) {
          keys[i] = c;
          keys[i + j] = a;
        }
      }
      __syncthreads();
    }
    for (int base = 0; base < n_pad; base += 4 * kDetectThreads) {
      const int i0 = base + 4 * tid;
      if (i0 - 4 * lane < n_pad) {  // warp-uniform: this warp's 128-key block starts inside the array
        u64 v[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r] = i0 + r < n_pad ? keys[i0 + r] : 0ull;
        for (int jj = j; jj >= 4; jj >>= 1) {  // partner = same slot r of the lane jj / 4 away
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const u64 o = __shfl_xor_sync(0xffffffffu, v[r], jj >> 2);
            const int i = i0 + r;
            const bool want_max = ((i & jj) == 0) == ((i & k) == 0);
            v[r] = want_max ? (v[r] > o ? v[r] : o) : (v[r] < o ? v[r] : o);
          }
        }
        if (k >= 4) {  // j = 2: slots (0, 2) and (1, 3)
#pragma unroll
          for (int r = 0; r < 2; ++r) {
            const bool up = ((i0 + r) & k) == 0;
            const u64 a = v[r], c = v[r + 2];
            if ((a < c) == up) {
              v[r] = c;
              v[r + 2] = a;
            }
          }
        }
#pragma unroll
        for (int r = 0; r < 4; r += 2) {  // j = 1: slots (0, 1) and (2, 3)
          const bool up = ((i0 + r) & k) == 0;
          const u64 a = v[r], c = v[r + 1];
          if ((a < c) == up) {
            v[r] = c;
            v[r + 1] = a;
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
          if (i0 + r < n_pad) keys[i0 + r] = v[r];
      }
    }
    __syncthreads();
  }
}

// Chunk of the walk, SORTED, by a bucket sort on the score bins: the exact per-bin counts of the list (bins[]) give
// every bin of the chunk its slice of keys[] (block-wide exclusive scan), the list keys in [lo, hi) are scattered into
// their slices (bin 0 = highest scores first), and each slice -- a handful of keys -- is finished by an insertion
// sort on the full 64-bit key.  Replaces a 2048-4096-key bitonic sort (66-78 compare-exchange stages) by three short
// passes.  Returns the number of keys, or -1 when some bin holds more than kMaxBucket keys (massive ties): the caller
// then uses the generic select + bitonic path.  bstart: kHistBins + 1 ints of scratch.
constexpr int kMaxBucket = 48;
constexpr int kBucketMax = 4096;  // largest chunk the bucket sort takes (8 keys per thread in registers)
__device__ int bucket_sort_chunk(const u64* __restrict__ list, int n_list, u64 lo, u64 hi, int bfrom, int bto,
                                 const unsigned* bins, int* bstart, u64* keys, int cap, unsigned* s_warp_tot,
                                 int* s_flag) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int per = kHistBins / kDetectThreads;
  if (tid == 0) *s_flag = 0;
  unsigned loc[per], sum = 0, mx = 0;
#pragma unroll
  for (int i = 0; i < per; ++i) {
    const int bin = tid * per + i;
    loc[i] = (bin >= bfrom && bin <= bto) ? bins[bin] : 0u;
    sum += loc[i];
    mx = max(mx, loc[i]);
  }
  unsigned inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  __syncthreads();
  if (lane == 31) s_warp_tot[warp] = inc;
  if (mx > (unsigned)kMaxBucket) *s_flag = 1;
  __syncthreads();
  unsigned off = 0, total = 0;
  for (int w = 0; w < kDetectThreads / 32; ++w) {
    if (w < warp) off += s_warp_tot[w];
    total += s_warp_tot[w];
  }
  if (*s_flag || total > (unsigned)min(cap, kBucketMax)) return -1;
  unsigned run = off + inc - sum;
#pragma unroll
  for (int i = 0; i < per; ++i) {
    bstart[tid * per + i] = (int)run;  // becomes the slice's END after the scatter
    run += loc[i];
  }
  __syncthreads();
  constexpr int kLU = 8;  // list loads in flight per thread (the list lives in L2)
  for (int i0 = 0; i0 < n_list; i0 += kLU * kDetectThreads) {
    u64 kk[kLU];
#pragma unroll
    for (int u = 0; u < kLU; ++u) {
      const int i = i0 + u * kDetectThreads + tid;
      kk[u] = i < n_list ? list[i] : 0ull;
    }
#pragma unroll
    for (int u = 0; u < kLU; ++u) {
      const u64 key = kk[u];
      if (key >= lo && key < hi && key != 0ull) keys[atomicAdd(&bstart[score_bin((uint32_t)(key >> 32))], 1)] = key;
    }
  }
  __syncthreads();
  // finish every slice by RANKING: one thread per key counts the keys of its slice that are larger (slices hold a
  // handful of keys; a thread-per-slice insertion sort left the block waiting for the fullest slice)
  constexpr int kKU = kBucketMax / kDetectThreads;  // keys per thread
  u64 mine[kKU];
  uint16_t dest[kKU];
#pragma unroll
  for (int u = 0; u < kKU; ++u) {
    const int i = u * kDetectThreads + tid;
    mine[u] = 0ull;
    dest[u] = 0;
    if (i < (int)total) {
      const u64 key = keys[i];
      const int bin = score_bin((uint32_t)(key >> 32));
      const int end = bstart[bin], start = bin == bfrom ? 0 : bstart[bin - 1];
      int rank = 0;
      for (int j = start; j < end; ++j) rank += keys[j] > key ? 1 : 0;
      mine[u] = key;
      dest[u] = (uint16_t)(start + rank);
    }
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < kKU; ++u)
    if (u * kDetectThreads + tid < (int)total) keys[dest[u]] = mine[u];
  __syncthreads();
  return (int)total;
}

// Block-wide search in a 4096-bin histogram held in SHARED memory: the smallest bin B >= bin_from whose running count
// from bin_from reaches `want` (kHistBins-1 when the whole tail holds less).  s_res[0] = B, s_res[1] = count of
// bins bin_from .. B, s_res[2] = count of bins bin_from .. B-1.
__device__ void pick_bin(const unsigned* bins, int bin_from, int want, unsigned* s_warp_tot, int* s_res) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int per = kHistBins / kDetectThreads;
  unsigned loc[per], sum = 0;
#pragma unroll
  for (int i = 0; i < per; ++i) {
    const int bin = tid * per + i;
    loc[i] = bin >= bin_from ? bins[bin] : 0u;
    sum += loc[i];
  }
  unsigned inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  __syncthreads();
  if (lane == 31) s_warp_tot[warp] = inc;
  __syncthreads();
  unsigned off = 0, total = 0;
  for (int w = 0; w < kDetectThreads / 32; ++w) {
    if (w < warp) off += s_warp_tot[w];
    total += s_warp_tot[w];
  }
  const unsigned excl = off + inc - sum;
  const unsigned uwant = (unsigned)max(want, 1);
  if (tid == 0 && total < uwant) {  // the tail holds less than wanted: take all of it
    s_res[0] = kHistBins - 1;
    s_res[1] = (int)total;
    s_res[2] = (int)total;
  }
  if (excl < uwant && excl + sum >= uwant) {
    unsigned c = excl;
    for (int i = 0; i < per; ++i) {
      if (c + loc[i] >= uwant) {
        s_res[0] = tid * per + i;
        s_res[1] = (int)(c + loc[i]);
        s_res[2] = (int)c;
        break;
      }
      c += loc[i];
    }
  }
  __syncthreads();
}

// lower key bound (inclusive) of "all pixels whose score bin is <= bin"
__device__ __forceinline__ u64 bin_to_lo_key(int bin, u64 floor_key) {
  if (bin >= kHistBins - 1) return floor_key;
  const uint32_t tb = kOneBits - ((uint32_t)(bin + 1) << kHistShift);  // bin(v) <= bin  <=>  bits > tb
  const u64 lo = ((u64)tb + 1ull) << 32;
  return lo < floor_key ? floor_key : lo;
}

// Exact greedy NMS, in parallel.  The reference walks candidates in rank order and keeps one iff no
// earlier-kept candidate lies within its (2d+1)^2 box (NN:229-255).  Equivalently: candidate i is
// KEPT iff every earlier-rank candidate inside its box is SUPPRESSED, and SUPPRESSED iff one of them
// is KEPT.  Each (Jacobi) round decides every candidate whose earlier-rank box neighbours were all decided
// in the previous rounds; states only move UNDECIDED -> KEPT/SUPPRESSED, at least the lowest-rank
// undecided candidate is decided per round, and the fixed point is the sequential result.  Neighbours are found through a spatial hash of 8x8-pixel cells (linked
// lists in shared memory).  Suppression by earlier chunks comes from the bitmap.
enum : uint8_t { ST_UNDEC = 0, ST_KEPT = 1, ST_SUPP = 2 };
constexpr uint16_t kNil = 0xFFFFu;

__global__ void __launch_bounds__(kDetectThreads, 2) k_detect(DetectParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  chain_enter();
  const int b = blockIdx.x;
  const int H = p.H, W = p.W, K = p.K, cap = p.cap, d = p.dist, bd = p.border;
  const int ww = (W + 31) >> 5;  // bitmap words per row
  const int Hc = H >> 3, Wc = W >> 3, cells = Hc * Wc;
  u64* keys = reinterpret_cast<u64*>(smem_raw);                  // [cap]   sorted candidate keys of the chunk
  u64* emit = keys + cap;                                        // [K]     emitted keypoints (score | y<<16 | x)
  int* head = reinterpret_cast<int*>(emit + K);                  // [cells + 1] spatial hash: start of each cell's run in next[]
  // pixels suppressed by EARLIER chunks live in a global-memory bitmap that only the multi-chunk path touches
  // (zeroed lazily), which keeps the CTA at ~85 KB of shared memory: two images per SM
  unsigned* bitmap = p.bitmap + (size_t)b * p.bitmap_stride;
  bool have_bitmap = false;
  uint16_t* next = reinterpret_cast<uint16_t*>(head + max(cells + 2, kHistBins + 2)); // [cap]   candidates grouped by cell (also the cell list of the gather)
  uint8_t* state = reinterpret_cast<uint8_t*>(next + cap);       // [2*cap] (second half: Jacobi double buffer)
  unsigned* s_hist = reinterpret_cast<unsigned*>(keys);          // aliases keys (radix select only)
  unsigned* bins = reinterpret_cast<unsigned*>(state + 2 * cap); // [kHistBins] score-bin histogram (chunk sizing)
  u64* list = p.list + (size_t)b * kListCap;                     // this image's candidate list (global, L2-resident)
  __shared__ int s_count, s_want, s_emitted, s_ncell2, s_gencells;
  __shared__ int s_res[3];
  __shared__ u64 s_prefix;
  __shared__ unsigned s_warp_tot[kDetectThreads / 32];

  const float* heat = p.heat + (size_t)b * H * W;
  const uint32_t needed_prev = p.spill_thr[b];               // score of the K-th keypoint of this slot's previous image
  const uint32_t thr_bits = store_threshold(needed_prev);    // the threshold k_softmax_heat stored cells by on this call
  SemiView sv;
  sv.base = static_cast<const unsigned char*>(p.semi) + (size_t)b * 65 * cells * (p.semi_f16 ? 2 : 4);
  sv.f16 = p.semi_f16; sv.cells = cells; sv.Wc = Wc; sv.fast_div = p.fast_div;
  // table of exponentials for visit_cells: inside the key buffer (idle while candidates are gathered), behind the
  // 256 radix-select counters that alias its start
  float* stage = reinterpret_cast<float*>(keys) + 256;
  const uint2* cellmax = p.cellmax + (size_t)b * cells;
  const uint32_t conf_bits = fbits(fmaxf(p.conf, 0.0f));
  const u64 floor_key = ((u64)conf_bits + 1ull) << 32;  // smallest possible candidate key (score > conf)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    s_emitted = 0;
    if (p.opmax_zero) p.opmax_zero[b] = 0u;  // k_desc_normalize takes an atomicMax into it (saves a memset node)
  }

  // Candidates are consumed in descending key order.  A GENERATION gathers every candidate of a key range
  // [lo_pre, hi_pre) ONCE into the image's global list (through the per-cell records: only contributing cells are
  // touched), with an exact score-bin histogram; the CHUNKS of the greedy walk are then cut from the list at bin
  // boundaries.  Normally there is one generation and (on independent-pixel heatmaps) one chunk; sparse, clustered
  // heatmaps of a real network need several chunks because most of their candidates are suppressed.  All sizing is
  // heuristic; overflow is detected exactly and repaired, so the result always equals full sort + sequential walk.
  bool slow = false;
  u64 hi_pre = ~0ull;  // every candidate >= hi_pre has been consumed
  bool first_generation = true;
  int emitted_before = 0;  // s_emitted at the start of the current chunk (marginal emission rate)
  PHASE_INIT();
  while (true) {
    PHASE(0);  // (setup / bookkeeping between generations)
    // ---- G1: lower bound of the generation -----------------------------------------------------------
    // First generation with history: a little below the score of the K-th keypoint this image slot emitted on the
    // previous call (every multi-candidate cell of that range is a stored one: 0.9 > 0.8).  Otherwise: from the
    // histogram of the cell MAXIMA (an estimate).  A bound that turns out too high costs another generation, one
    // too low a longer list -- never the result.
    u64 lo_pre;
    if (first_generation && needed_prev > conf_bits && needed_prev <= kOneBits) {
      const uint32_t gb = fbits(__fmul_rn(0.9f, __uint_as_float(needed_prev)));
      lo_pre = bin_to_lo_key(score_bin(gb), floor_key);
    } else {
      for (int i = tid; i < kHistBins; i += kDetectThreads) bins[i] = 0u;
      __syncthreads();
      {
        const uint32_t hi_b = (uint32_t)(hi_pre >> 32);
        constexpr int kGU = 8;  // records in flight per thread (the loop is a chain of L2 round trips otherwise)
        for (int c0 = 0; c0 < cells; c0 += kGU * kDetectThreads) {
          uint32_t mb[kGU];
#pragma unroll
          for (int u = 0; u < kGU; ++u) {
            const int c = c0 + u * kDetectThreads + tid;
            mb[u] = c < cells ? __ldg(&cellmax[c].x) : 0u;
          }
#pragma unroll
          for (int u = 0; u < kGU; ++u)
            if (mb[u] > conf_bits && mb[u] < kInfBits && mb[u] <= hi_b) atomicAdd(&bins[score_bin(mb[u])], 1u);
        }
      }
      __syncthreads();
      pick_bin(bins, 0, p.target, s_warp_tot, s_res);
      lo_pre = bin_to_lo_key(s_res[0], floor_key);
      if (lo_pre >= hi_pre) lo_pre = floor_key;
      __syncthreads();
    }
    first_generation = false;
    PHASE(1);  // G1
    // ---- G2: gather [lo_pre, hi_pre) into the list; exact bins --------------------------------------
    int n_list;
    while (true) {
      for (int i = tid; i < kHistBins; i += kDetectThreads) bins[i] = 0u;
      __syncthreads();
      n_list = collect_to_list(sv, heat, thr_bits, cellmax, H, W, conf_bits, lo_pre, hi_pre, list, kListCap, bins, next,
                               cap, &s_count, &s_want, &s_ncell2, &s_gencells, stage);
      __syncthreads();
      if (n_list <= kListCap) break;
      slow = true;  // more candidates than the list holds: raise the lower bound exactly and gather again
      pick_bin(bins, 0, kListCap * 3 / 4, s_warp_tot, s_res);
      u64 lo2 = s_res[1] <= kListCap ? bin_to_lo_key(s_res[0], floor_key)
                                     : (s_res[2] > 0 ? bin_to_lo_key(s_res[0] - 1, floor_key) : 0ull);
      __syncthreads();
      if (lo2 <= lo_pre) {  // one score bin alone overflows the list (massive ties): exact select on the heatmap
        lo2 = radix_select(sv, H, conf_bits, hi_pre, kListCap * 3 / 4, floor_key, s_hist, &s_prefix, &s_want, stage);
        __syncthreads();
      }
      lo_pre = lo2;
    }
    PHASE(2);  // G2
    // ---- G3: chunks of the walk ---------------------------------------------------------------------
    const int gen_cells = max(s_gencells, 1);  // cells whose maximum lies in the generation's range (last gather)
    u64 hi = hi_pre;
    // Clustered heatmaps (a real network: ~7 candidates per contributing cell, most of them suppressed by their
    // blob's peak) need a walk several times longer than K: size the first chunk by the observed candidates per cell
    // instead of discovering that one small chunk at a time.  Independent pixels (~1 per cell) keep the small chunk.
    int bfrom = 0;
    int chunk_target = (int)min((long long)min(cap * 3 / 4, kBucketMax - 512),
                                max((long long)p.target, (long long)K * n_list * 6 / (10LL * gen_cells) + 256));
    bool done = false;
    bool whole = n_list <= min(cap * 3 / 4, p.target + p.target / 2);  // small list: the first chunk takes all of it
    bool stale_bins = false;  // a bin was cut inside (exact select): its count no longer matches the remaining keys
    while (true) {
      int bto = kHistBins - 1;
      u64 lo;
      if (whole) {
        lo = lo_pre;
      } else {
        pick_bin(bins, bfrom, chunk_target, s_warp_tot, s_res);
        bto = s_res[0];
      }
      if (whole) {
      } else if (s_res[1] <= cap) {
        lo = bin_to_lo_key(bto, floor_key);
      } else if (s_res[2] > 0) {
        bto -= 1;
        lo = bin_to_lo_key(bto, floor_key);
      } else {  // one score bin alone overflows the chunk buffer: cut it exactly (its count stays stale; the
                // selection below is exact, so later chunks can only be smaller than estimated)
        __syncthreads();
        lo = radix_select_list(list, n_list, hi, cap * 3 / 4, lo_pre, s_hist, &s_prefix, &s_want);
        bto -= 1;
        stale_bins = true;
      }
      if (lo < lo_pre) lo = lo_pre;
      __syncthreads();
      // sorted chunk: bucket sort on the exact bins, or (stale bin counts after an exact cut / massive ties) the
      // generic selection + bitonic sort
      int n = -1;
      if (!stale_bins)
        n = bucket_sort_chunk(list, n_list, lo, hi, bfrom, min(bto, kHistBins - 1), bins, head, keys, cap, s_warp_tot, &s_want);
      if (n < 0) {
        __syncthreads();
        n = select_from_list(list, n_list, lo, hi, keys, cap, &s_count);
        int n_pad = 32;
        while (n_pad < n) n_pad <<= 1;
        for (int i = n + tid; i < n_pad; i += kDetectThreads) keys[i] = 0ull;
        __syncthreads();
        bitonic_sort_desc(keys, n_pad);
      }
      PHASE(3);  // chunk selection + sort
      PHASE_NOTE(16 + (int)g_phase_clk_chunks(), n);
      PHASE_COUNT(12);
      whole = false;
    for (int i = tid; i <= cells; i += kDetectThreads) head[i] = 0;
    __syncthreads();

    // ---- A: unpack positions, initial states, spatial hash ---------------------------------------
    // The hash is a counting sort of the candidates by 8x8-pixel cell: head[c] .. head[c+1] delimit the cell's
    // candidates in next[] (contiguous, so a neighbourhood scan issues independent loads instead of chasing links).
    for (int i = tid; i < n; i += kDetectThreads) {
      const u64 key = keys[i];
      const uint32_t pos = 0xFFFFFFFFu - (uint32_t)key;
      const int x = (int)(pos / (uint32_t)H), y = (int)(pos - (uint32_t)x * (uint32_t)H);
      keys[i] = (key & 0xFFFFFFFF00000000ull) | (u64)(((uint32_t)y << 16) | (uint32_t)x);
      const bool sup = have_bitmap && ((bitmap[y * ww + (x >> 5)] >> (x & 31)) & 1u);
      state[i] = sup ? ST_SUPP : ST_UNDEC;
      // d == 0: a point only suppresses its own pixel, nothing interacts
      if (!sup && d > 0) atomicAdd(&head[(y >> 3) * Wc + (x >> 3)], 1);
    }
    __syncthreads();
    if (d > 0) {
      // exclusive scan of the per-cell counts (each thread owns a run of consecutive cells)
      const int per = (cells + kDetectThreads - 1) / kDetectThreads;
      const int c0 = min(tid * per, cells), c1 = min(c0 + per, cells);
      int sum = 0;
      for (int c = c0; c < c1; ++c) sum += head[c];
      int inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      if (lane == 31) s_warp_tot[warp] = (unsigned)inc;
      __syncthreads();
      int off = 0;
      for (int w = 0; w < warp; ++w) off += (int)s_warp_tot[w];
      int run = off + inc - sum;
      for (int c = c0; c < c1; ++c) {  // head[c] = start | (count << 16): the count is consumed by the scatter below
        const int cnt = head[c];
        head[c] = run | (cnt << 16);
        run += cnt;
      }
      if (tid == kDetectThreads - 1) head[cells] = run;
      __syncthreads();
      for (int i = tid; i < n; i += kDetectThreads) {
        if (state[i] != ST_UNDEC) continue;
        const uint32_t xy = (uint32_t)keys[i];
        const int old = atomicSub(&head[(int)(xy >> 19) * Wc + (int)((xy & 0xFFFF) >> 3)], 1 << 16);
        next[(old & 0xFFFF) + (old >> 16) - 1] = (uint16_t)i;
      }
      __syncthreads();
    }

    PHASE(4);  // A: hash
    // ---- B: fixed-point rounds ---------------------------------------------------------------------
    if (d == 0) {
      for (int i = tid; i < n; i += kDetectThreads)
        if (state[i] == ST_UNDEC) state[i] = ST_KEPT;
      __syncthreads();
    } else {
      // Jacobi rounds: every round reads the previous round's states (st_in) and writes st_out, so there is no
      // intra-round data race; the two byte arrays swap roles after each barrier.
      uint8_t* st_in = state;
      uint8_t* st_out = state + cap;
      // In the FIRST round nobody is decided yet (apart from candidates suppressed by earlier chunks, which are not in
      // the hash): a candidate with any earlier-rank candidate in its box stays undecided, whatever else is around it,
      // so its scan may stop at the first one it meets -- in a blob of a real heatmap that is almost immediately.
      bool first_round = true;
      while (true) {
        int undecided = 0;
        for (int i = tid; i < n; i += kDetectThreads) {
          const uint8_t si = st_in[i];
          uint8_t so = si;
          if (si == ST_UNDEC) {
            const uint32_t xy = (uint32_t)keys[i];
            const int x = xy & 0xFFFF, y = xy >> 16;
            const int cx0 = max(x - d, 0) >> 3, cx1 = min(x + d, W - 1) >> 3;
            const int cy0 = max(y - d, 0) >> 3, cy1 = min(y + d, H - 1) >> 3;
            bool kept = false, undec = false;
            // The round's time is its slowest thread's, and a neighbour costs a chain of three dependent shared loads
            // (next -> keys -> state): walk the runs FOUR neighbours at a time so that the chains overlap.
            auto scan4 = [&](int k, int k1) {  // candidates next[k .. min(k + 4, k1))
              int q[4];
              uint32_t qxy[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) q[u] = k + u < k1 ? (int)next[k + u] : 0x7FFFFFFF;
#pragma unroll
              for (int u = 0; u < 4; ++u) qxy[u] = q[u] < i ? (uint32_t)keys[q[u]] : 0u;
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int qx = qxy[u] & 0xFFFF, qy = qxy[u] >> 16;
                if (q[u] < i && abs(qx - x) <= d && abs(qy - y) <= d) {
                  const uint8_t sq = st_in[q[u]];
                  kept |= sq == ST_KEPT;
                  undec |= sq == ST_UNDEC;
                }
              }
            };
            // the cells cx0 .. cx1 of one cell row are adjacent in the counting sort: one contiguous range per row
            if (cy1 - cy0 <= 2) {  // d <= 8: at most three cell rows, whose six bounds are fetched together
              int k0[3], len[3];
#pragma unroll
              for (int r = 0; r < 3; ++r) {
                const int cy = min(cy0 + r, cy1);
                k0[r] = head[cy * Wc + cx0];
                len[r] = cy0 + r <= cy1 ? head[cy * Wc + cx1 + 1] - k0[r] : 0;
              }
#pragma unroll
              for (int r = 0; r < 3; ++r)
                for (int k = 0; k < len[r] && !kept && !(first_round && undec); k += 4) scan4(k0[r] + k, k0[r] + len[r]);
            } else {
              for (int cy = cy0; cy <= cy1 && !kept; ++cy) {
                const int k1 = head[cy * Wc + cx1 + 1];
                for (int k = head[cy * Wc + cx0]; k < k1 && !kept && !(first_round && undec); k += 4) scan4(k, k1);
              }
            }
            if (kept) so = ST_SUPP;
            else if (!undec) so = ST_KEPT;
            else ++undecided;
          }
          st_out[i] = so;
        }
        const int remaining = __syncthreads_count(undecided > 0);
        PHASE_COUNT(13);
        uint8_t* t = st_in;
        st_in = st_out;
        st_out = t;
        if (first_round) PHASE(10); else PHASE(11);
        if (first_round) PHASE(10); else PHASE(11);
        first_round = false;
        if (remaining == 0) break;
      }
      if (st_in != state) {  // final states must end up in state[]
        for (int i = tid; i < n; i += kDetectThreads) state[i] = st_in[i];
        __syncthreads();
      }
    }

    PHASE(5);  // B: NMS rounds
    // ---- C: emit kept in-border candidates in rank order, up to K ------------------------------------
    {
      const int seg = (n + kDetectThreads - 1) / kDetectThreads;
      const int i0 = min(tid * seg, n), i1 = min(i0 + seg, n);
      int cnt = 0;
      for (int i = i0; i < i1; ++i) {
        const uint32_t xy = (uint32_t)keys[i];
        const int x = xy & 0xFFFF, y = xy >> 16;
        cnt += (state[i] == ST_KEPT && y >= bd && y + bd < H && x >= bd && x + bd < W) ? 1 : 0;
      }
      int inc = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int v = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += v;
      }
      if (lane == 31) s_warp_tot[warp] = (unsigned)inc;
      __syncthreads();
      int off = s_emitted, total = 0;
      for (int w = 0; w < kDetectThreads / 32; ++w) {
        if (w < warp) off += (int)s_warp_tot[w];
        total += (int)s_warp_tot[w];
      }
      int slot = off + inc - cnt;
      for (int i = i0; i < i1; ++i) {
        const u64 key = keys[i];
        const uint32_t xy = (uint32_t)key;
        const int x = xy & 0xFFFF, y = xy >> 16;
        if (state[i] == ST_KEPT && y >= bd && y + bd < H && x >= bd && x + bd < W) {
          if (slot < K) emit[slot] = key;
          ++slot;
        }
      }
      __syncthreads();
      if (tid == 0) s_emitted = min(K, s_emitted + total);
      __syncthreads();
    }
      PHASE(6);  // C: emit
      PHASE_NOTE(23 + (int)g_phase_clk_chunks(), s_emitted);
      const bool list_done = lo <= lo_pre;  // this chunk took the rest of the generation's list
      if (s_emitted >= K || (list_done && lo_pre <= floor_key)) {
        done = true;
        break;
      }

      // ---- D: more candidates are needed: record this chunk's boxes for the candidates still to come ----
      slow = true;
      if (!have_bitmap) {
        for (int i = tid; i < H * ww; i += kDetectThreads) bitmap[i] = 0u;
        have_bitmap = true;
        __syncthreads();
      }
      for (int i = tid; i < n; i += kDetectThreads) {
        if (state[i] != ST_KEPT) continue;
        const uint32_t xy = (uint32_t)keys[i];
        const int xj = xy & 0xFFFF, yj = xy >> 16;
        const int x0 = max(xj - d, 0), x1 = min(xj + d, W - 1);
        for (int yy = max(yj - d, 0); yy <= min(yj + d, H - 1); ++yy)
          for (int w = x0 >> 5; w <= (x1 >> 5); ++w) {
            const int lo_b = max(x0 - (w << 5), 0), hi_b = min(x1 - (w << 5), 31);
            atomicOr(&bitmap[yy * ww + w], (0xffffffffu >> (31 - hi_b)) & (0xffffffffu << lo_b));
          }
      }
      __syncthreads();
      PHASE(7);  // D: bitmap of this chunk's boxes
      hi = lo;
      if (list_done) break;  // next generation: gather the candidates below lo_pre
      bfrom = bto + 1;
      // size the next chunk from what the LAST chunk yielded: (K - emitted) more keypoints at its emission rate -- the
      // marginal rate, which keeps falling on clustered heatmaps (real outputs: 28 % -> 10 % -> 9 % -> 7 % per chunk:
      // lower-scored candidates are more often inside an earlier keypoint's box), so the average over the whole walk
      // undersizes every chunk -- with a margin; bounded by the buffer
      const int d_em = max(s_emitted - emitted_before, 1);
      const long long need = (long long)(K - s_emitted) * n / d_em;
      chunk_target = (int)min((long long)min(cap * 3 / 4, kBucketMax - 512), max(512LL, need + need / 2 + 128));
      emitted_before = s_emitted;
      __syncthreads();
    }
    if (done) break;
    hi_pre = lo_pre;
  }

  PHASE(8);  // D + loop bookkeeping of the last chunk
  // ---- outputs ---------------------------------------------------------------------------------
  const int n_emit = min(s_emitted, K);
  spvo_keypoint* kp = p.kpts + (size_t)b * K;
  for (int i = tid; i < K; i += kDetectThreads) {
    spvo_keypoint o;
    float sc = 0.f;
    if (i < n_emit) {
      const u64 key = emit[i];
      const uint32_t xy = (uint32_t)key;
      o.x = (float)(xy & 0xFFFF); o.y = (float)(xy >> 16); o.size = 1.0f; o.angle = -1.0f; o.response = 0.0f;
      o.octave = 0; o.class_id = -1;
      sc = __uint_as_float((uint32_t)(key >> 32));
    } else {
      o.x = o.y = o.size = o.angle = o.response = 0.0f;
      o.octave = 0; o.class_id = 0;
    }
    kp[i] = o;
    if (p.scores) p.scores[(size_t)b * K + i] = sc;
    if (p.kp_par) {
      // align-corners sampling coordinates (NN:377-392), same fp32 operation order as the oracle
      int4 par = make_int4(0, 0, 0, 0);
      if (i < n_emit) {
        const float r8 = __fmul_rn(__fdiv_rn(o.y, (float)(H - 1)), (float)(Hc - 1));
        const float c8 = __fmul_rn(__fdiv_rn(o.x, (float)(W - 1)), (float)(Wc - 1));
        const int r0 = (int)floorf(r8), c0 = (int)floorf(c8);
        const float rr = __fsub_rn(1.0f, __fsub_rn(r8, (float)r0));
        const float cr = __fsub_rn(1.0f, __fsub_rn(c8, (float)c0));
        const int r1 = min(r0 + 1, Hc - 1), c1 = min(c0 + 1, Wc - 1);
        par.x = r0 * Wc + c0;
        par.y = ((r1 - r0) * Wc) | ((c1 - c0) << 30);
        par.z = __float_as_int(rr);
        par.w = __float_as_int(cr);
      }
      p.kp_par[(size_t)b * K + i] = par;
    }
  }
  PHASE(9);  // outputs
  if (tid == 0) {
    p.n_out[b] = n_emit;
    if (slow) atomicAdd(&p.counters[0], 1ull);
    // for the next call on this image slot (store_threshold): the score of the K-th keypoint, or the confidence
    // threshold when the image yields fewer than K
    p.spill_thr[b] = n_emit >= K ? (uint32_t)(emit[K - 1] >> 32) : conf_bits;
  }
}

// ------------------------------------------------------------------------------------------------
// K3: descriptor sampling.  One warp per keypoint; lane l owns channels l, l+32, ..., l+224.
// Arithmetic order is the oracle's (sample_descriptor): each term (vec*s1)*s2, summed left to
// right, unfused; squared norm = per-lane partial sums over ascending channels, xor butterfly
// 16,8,4,2,1; true division by sqrt.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
k_sample_desc(const T* __restrict__ desc, const spvo_keypoint* __restrict__ kpts, const int* __restrict__ n_out,
              float* __restrict__ out, int H, int W, int K) {
  const int b = blockIdx.y;
  const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k >= K) return;
  float* o = out + ((size_t)b * K + k) * 256;
  if (k >= n_out[b]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[lane + 32 * i] = 0.0f;
    return;
  }
  const int Hc = H >> 3, Wc = W >> 3;
  const size_t cells = (size_t)Hc * Wc;
  const spvo_keypoint kp = kpts[(size_t)b * K + k];
  const int x = (int)kp.x, y = (int)kp.y;
  const float r8 = __fmul_rn(__fdiv_rn((float)y, (float)(H - 1)), (float)(Hc - 1));
  const float c8 = __fmul_rn(__fdiv_rn((float)x, (float)(W - 1)), (float)(Wc - 1));
  const int r0 = (int)floorf(r8), c0 = (int)floorf(c8);
  const float rr = __fsub_rn(1.0f, __fsub_rn(r8, (float)r0));
  const float cr = __fsub_rn(1.0f, __fsub_rn(c8, (float)c0));
  const float irr = __fsub_rn(1.0f, rr), icr = __fsub_rn(1.0f, cr);
  const int r1 = min(r0 + 1, Hc - 1), c1 = min(c0 + 1, Wc - 1);
  const T* base = desc + (size_t)b * 256 * cells;
  const T* tl = base + (size_t)r0 * Wc + c0;
  const T* tr = base + (size_t)r0 * Wc + c1;
  const T* bl = base + (size_t)r1 * Wc + c0;
  const T* br = base + (size_t)r1 * Wc + c1;
  float a[8], bq[8], c[8], dd[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const size_t off = (size_t)(lane + 32 * i) * cells;
    a[i] = ld_in(tl + off);
    bq[i] = ld_in(tr + off);
    c[i] = ld_in(bl + off);
    dd[i] = ld_in(br + off);
  }
  float v[8], part = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float t1 = __fmul_rn(__fmul_rn(a[i], rr), cr);
    const float t2 = __fmul_rn(__fmul_rn(bq[i], rr), icr);
    const float t3 = __fmul_rn(__fmul_rn(c[i], irr), cr);
    const float t4 = __fmul_rn(__fmul_rn(dd[i], irr), icr);
    v[i] = __fadd_rn(__fadd_rn(__fadd_rn(t1, t2), t3), t4);
    part = __fadd_rn(part, __fmul_rn(v[i], v[i]));
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) part = __fadd_rn(part, __shfl_xor_sync(0xffffffffu, part, off));
  if (part > 0.0f) {
    const float nrm = __fsqrt_rn(part);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __fdiv_rn(v[i], nrm);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) o[lane + 32 * i] = v[i];
}

// ------------------------------------------------------------------------------------------------
// K3 (streaming form): k_desc_planes + k_desc_normalize.
// A keypoint touches 4 cells x 256 channels; in NCHW every channel of a cell lives in a different
// 32-byte sector, so gathering per keypoint moves 2-3x the tensor through L2.  Instead each CTA
// streams kCP whole channel planes (Hc*Wc floats, contiguous) into shared memory with cp.async --
// every byte of desc is read exactly once, fully coalesced -- and evaluates the bilinear blend of
// those channels for ALL keypoints of the image from shared memory.  Values go to a [256][K]
// scratch (coalesced along K); k_desc_normalize transposes 32 keypoints at a time through shared
// memory, applies the oracle's norm reduction order and writes the [K][256] rows.
// ------------------------------------------------------------------------------------------------
constexpr int kCP = 2;  // channel planes per CTA: 2 x 29 KB at 1240x376 -> 3 CTAs per SM (1 plane/CTA measured the same)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}

template <typename T>
__global__ void __launch_bounds__(256)
k_desc_planes(const T* __restrict__ desc, const int4* __restrict__ kp_par, const int* __restrict__ n_out,
              float* __restrict__ tmp, int cells, int K, int plane_pitch, int Kp) {
  chain_enter();
  extern __shared__ __align__(16) unsigned char sp_raw[];  // kCP planes, each plane_pitch elements of T
  T* sp = reinterpret_cast<T*>(sp_raw);
  const int b = blockIdx.y, cg = blockIdx.x;
  const int n = n_out[b];
  if (n == 0) return;
  constexpr int kPer16 = 16 / (int)sizeof(T);  // elements per 16-byte cp.async
  int mis[kCP];
#pragma unroll
  for (int c = 0; c < kCP; ++c) {
    const T* src = desc + ((size_t)b * 256 + (size_t)cg * kCP + c) * cells;
    // keep the 16-byte phase of the global address so the body can use 16-byte cp.async
    const int m = (int)((reinterpret_cast<uintptr_t>(src) / sizeof(T)) & (kPer16 - 1));
    mis[c] = m;
    T* dstp = sp + (size_t)c * plane_pitch + m;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(dstp);
    const int head = min((kPer16 - m) & (kPer16 - 1), cells);
    const int body = (cells - head) / kPer16;
    const int tail0 = head + body * kPer16;
    if (sizeof(T) == 4) {  // 4-byte elements: head / tail are asynchronous 4-byte copies
      if (threadIdx.x < head) cp_async4(dst + threadIdx.x * 4, src + threadIdx.x);
    }
    for (int i = threadIdx.x; i < body; i += 256)
      cp_async16(dst + (head + kPer16 * i) * (int)sizeof(T), src + head + kPer16 * i);
    if (sizeof(T) == 4) {
      if (threadIdx.x < cells - tail0) cp_async4(dst + (tail0 + threadIdx.x) * 4, src + tail0 + threadIdx.x);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // 2-byte elements (cp.async moves 4 / 8 / 16 bytes): warp c loads plane c's < 8 head / tail elements into
  // registers now and stores them after the wait below, so that their latency hides behind the asynchronous body
  const int hw = threadIdx.x >> 5, ht = threadIdx.x & 31;
  T hv = T(), tv = T();
  int h_idx = -1, t_idx = -1, hm = 0;
  if (sizeof(T) != 4 && hw < kCP) {
    const T* src = desc + ((size_t)b * 256 + (size_t)cg * kCP + hw) * cells;
    hm = (int)((reinterpret_cast<uintptr_t>(src) / sizeof(T)) & (kPer16 - 1));
    const int head = min((kPer16 - hm) & (kPer16 - 1), cells);
    const int tail0 = head + (cells - head) / kPer16 * kPer16;
    if (ht < head) { h_idx = ht; hv = src[ht]; }
    if (ht < cells - tail0) { t_idx = tail0 + ht; tv = src[tail0 + ht]; }
  }
  // the sampling parameters of this thread's keypoints travel while the planes are in flight
  constexpr int kPre = 4;
  int4 pre[kPre];
#pragma unroll
  for (int i = 0; i < kPre; ++i) {
    const int k = threadIdx.x + 256 * i;
    pre[i] = k < n ? __ldg(kp_par + (size_t)b * K + k) : make_int4(0, 0, 0, 0);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (sizeof(T) != 4 && hw < kCP) {
    T* dstp = sp + (size_t)hw * plane_pitch + hm;
    if (h_idx >= 0) dstp[h_idx] = hv;
    if (t_idx >= 0) dstp[t_idx] = tv;
  }
  __syncthreads();
#pragma unroll 1
  for (int k0 = 0; k0 < n; k0 += 256 * kPre) {
#pragma unroll
  for (int i = 0; i < kPre; ++i) {
    const int k = k0 + threadIdx.x + 256 * i;
    if (k >= n) break;
    const int4 par = k0 == 0 ? pre[i] : __ldg(kp_par + (size_t)b * K + k);
    const int o_tl = par.x, dr = par.y & 0x3FFFFFFF, dc = (par.y >> 30) & 1;
    const float rr = __int_as_float(par.z), cr = __int_as_float(par.w);
    const float irr = __fsub_rn(1.0f, rr), icr = __fsub_rn(1.0f, cr);
#pragma unroll
    for (int c = 0; c < kCP; ++c) {
      const T* pl = sp + (size_t)c * plane_pitch + mis[c];
      const float t1 = __fmul_rn(__fmul_rn(to_f32(pl[o_tl]), rr), cr);
      const float t2 = __fmul_rn(__fmul_rn(to_f32(pl[o_tl + dc]), rr), icr);
      const float t3 = __fmul_rn(__fmul_rn(to_f32(pl[o_tl + dr]), irr), cr);
      const float t4 = __fmul_rn(__fmul_rn(to_f32(pl[o_tl + dr + dc]), irr), icr);
      tmp[((size_t)b * 256 + (size_t)cg * kCP + c) * Kp + k] = __fadd_rn(__fadd_rn(__fadd_rn(t1, t2), t3), t4);
    }
  }
  }
}

// 32 keypoints per block: [256][K] scratch -> shared [32][257] -> normalised [K][256] rows.
__global__ void __launch_bounds__(256)
k_desc_normalize(const float* __restrict__ tmp, const int* __restrict__ n_out, float* __restrict__ out, int K, int Kp,
                 TcSink sink, const StereoSetup setup) {
  chain_enter();
  __shared__ float s[32][257];
  const int b = blockIdx.y, k0 = blockIdx.x * 32;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n = n_out[b];
  if (k0 < n) {
    // 16-byte loads: thread (c_sub = t / 8, j = t % 8) takes keypoints k0 + 4j .. + 3 of channels c_sub + 32 i; all 8
    // loads are in flight at once (the scratch pitch Kp is a multiple of 4).  The scattered shared stores hit banks
    // (4j + q + c) % 32: conflict-free.
    const float* src = tmp + (size_t)b * 256 * Kp + k0;
    const int j = threadIdx.x & 7, cs = threadIdx.x >> 3;
    if (k0 + 4 * j < n) {
      float4 q[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = __ldg(reinterpret_cast<const float4*>(src + (size_t)(cs + 32 * i) * Kp) + j);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = cs + 32 * i;
        s[4 * j + 0][c] = q[i].x;
        s[4 * j + 1][c] = q[i].y;
        s[4 * j + 2][c] = q[i].z;
        s[4 * j + 3][c] = q[i].w;
      }
    }
  }
  __syncthreads();
  unsigned short* xb = reinterpret_cast<unsigned short*>(sink.xb);
  float smax = 0.f;
  for (int kk = w; kk < 32; kk += 8) {
    const int k = k0 + kk;
    if (xb && k < sink.cap && k >= n) {  // padded rows of the matcher slot: zero row, norm = NaN (match_tc.cu: pad_norm)
      const size_t row = (size_t)b * sink.cap + k;
#pragma unroll
      for (int i = 0; i < 8; ++i) xb[row * 256 + lane + 32 * i] = 0;
      if (lane == 0) sink.nrm[row] = __uint_as_float(0x7FFFFFFFu);
    }
    if (k >= K) continue;
    float* o = out + ((size_t)b * K + k) * 256;
    if (k >= n) {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[lane + 32 * i] = 0.0f;
      continue;
    }
    float v[8], part = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      v[i] = s[kk][lane + 32 * i];
      part = __fadd_rn(part, __fmul_rn(v[i], v[i]));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) part = __fadd_rn(part, __shfl_xor_sync(0xffffffffu, part, off));
    if (part > 0.0f) {
      const float nrm = __fsqrt_rn(part);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __fdiv_rn(v[i], nrm);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) o[lane + 32 * i] = v[i];
    if (xb) {  // tensor-matcher operand: bf16 row + squared norm of the fp32 row (any summation order)
      const size_t row = (size_t)b * sink.cap + k;
      float s2 = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xb[row * 256 + lane + 32 * i] = sink.fp16 ? __half_as_ushort(__float2half_rn(v[i]))
                                                  : __bfloat16_as_ushort(__float2bfloat16_rn(v[i]));
        s2 = __fmaf_rn(v[i], v[i], s2);
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, off);
      if (lane == 0) sink.nrm[row] = s2;
      smax = fmaxf(smax, s2);
    }
  }
  if (xb && lane == 0 && smax > 0.f) atomicMax(&sink.opmax[b], __float_as_uint(smax));
  // the stereo batch's match problems (saves a launch): one block, AFTER its share of the real work -- at the top of
  // the kernel the same lines cost every block 20 % (they kept the first loads from being hoisted)
  if (setup.probs && blockIdx.x == 0 && blockIdx.y == 0)
    for (int p = threadIdx.x; p < 2 * setup.F; p += blockDim.x) setup_stereo_problem(setup, p);
}

// ------------------------------------------------------------------------------------------------
// host-side launcher
// ------------------------------------------------------------------------------------------------
static uint32_t fbits_host(float f) {
  uint32_t u;
  memcpy(&u, &f, sizeof u);
  return u;
}

static size_t detect_smem_bytes(int H, int W, int K, int cap) {
  const int ww = (W + 31) >> 5;
  const size_t cells = (size_t)(H / 8) * (W / 8);
  (void)ww;
  const size_t head_ints = cells + 2 > (size_t)kHistBins + 2 ? cells + 2 : (size_t)kHistBins + 2;  // also the bucket sort's scratch
  return (size_t)cap * 8 + (size_t)K * 8 + head_ints * 4 + (size_t)cap * 2 + (size_t)cap * 2 + (size_t)kHistBins * 4;
}

// One contiguous range of images [b0, b0 + B) on the handle's CURRENT stream (h->stream).
static cudaError_t launch_decode_range(Handle* h, const void* semi_v, const void* desc_v, int in_f16, int b0, int B,
                                       int H, int W,
                                       const spvo_decode_cfg& cfg, spvo_keypoint* kpts, float* desc_out, int* n_out,
                                       float* scores, const TcSink* sink, bool* sink_filled) {
  cudaStream_t st = h->stream;
  const int Hc = H / 8, Wc = W / 8, cells = Hc * Wc, K = cfg.max_keypoints;
  cudaError_t e;
  const size_t esz = in_f16 ? sizeof(__half) : sizeof(float);
  const unsigned char* semi = static_cast<const unsigned char*>(semi_v) + (size_t)b0 * 65 * cells * esz;
  const unsigned char* desc = desc_v ? static_cast<const unsigned char*>(desc_v) + (size_t)b0 * 256 * cells * esz : nullptr;
  kpts += (size_t)b0 * K;
  if (desc_out) desc_out += (size_t)b0 * K * 256;
  n_out += b0;
  if (scores) scores += (size_t)b0 * K;
  float* heat = h->heat + (size_t)b0 * H * W;
  uint2* cellmax = h->cellmax + (size_t)b0 * cells;
  dim3 g1((cells + kHeatThreads - 1) / kHeatThreads, B);
  // quotients below 2^-90 may differ from the IEEE division in the shared-reciprocal form: harmless while they cannot
  // be candidates (conf_thresh default 0.015); an (absurdly) small threshold takes the stand-alone division
  const int fast_div = cfg.conf_thresh >= 1e-20f ? 1 : 0;
  {
    LaunchScope ls(h, KID_SOFTMAX_HEAT);
    const uint32_t conf_bits = fbits_host(fmaxf(cfg.conf_thresh, 0.0f));
    if (in_f16)
      e = launch_chained(h->chain_launches, k_softmax_heat<__half>, g1, dim3(kHeatThreads), 0, st, 1, reinterpret_cast<const __half*>(semi),
                         heat, cellmax, h->spill_thr, b0, conf_bits, Hc, Wc, fast_div);
    else
      e = launch_chained(h->chain_launches, k_softmax_heat<float>, g1, dim3(kHeatThreads), 0, st, 1, reinterpret_cast<const float*>(semi),
                         heat, cellmax, h->spill_thr, b0, conf_bits, Hc, Wc, fast_div);
    if (e != cudaSuccess) return e;
  }
  if (K > 0) {
    DetectParams p;
    p.semi = semi; p.semi_f16 = in_f16; p.fast_div = fast_div; p.cellmax = cellmax;
    p.heat = heat; p.spill_thr = h->spill_thr + b0; p.list = h->cand_list + (size_t)b0 * kListCap; p.H = H; p.W = W;
    p.conf = cfg.conf_thresh;
    p.dist = cfg.dist_thresh; p.border = cfg.border_remove; p.K = K;
    p.kpts = kpts; p.scores = scores; p.n_out = n_out; p.counters = h->counters;
    const int per16 = (int)(16 / esz);  // one 16-byte unit of slack for the address phase, pitch in whole units
    const int plane_pitch = (cells + per16 + per16 - 1) & ~(per16 - 1);
    const size_t smem_planes = (size_t)kCP * plane_pitch * esz;
    const bool streaming = desc && desc_out && h->desc_tmp && h->kp_par && smem_planes <= 200 * 1024;
    // only the streaming form writes the tensor matcher's operands; the gather form (planes larger than shared
    // memory) leaves the sink untouched and the caller must run k_tc_prep instead
    if (sink_filled) *sink_filled = streaming && sink && sink->xb;
    int4* kp_par = streaming ? h->kp_par + (size_t)b0 * K : nullptr;
    const int Kp = (K + 3) & ~3;  // scratch pitch: 16-byte loads in k_desc_normalize
    float* tmp = streaming ? h->desc_tmp + (size_t)b0 * 256 * Kp : nullptr;
    p.kp_par = kp_par;
    p.opmax_zero = (streaming && sink && sink->opmax) ? sink->opmax + b0 : nullptr;
    p.bitmap = h->nms_bitmap + (size_t)b0 * ((size_t)h->max_h * h->max_w / 16 + 64);
    p.bitmap_stride = (size_t)h->max_h * h->max_w / 16 + 64;
    p.cap = K <= 1536 ? 4096 : 8192;
    p.target = min(p.cap * 3 / 4, K + K / 2 + 256);
    const size_t smem = detect_smem_bytes(H, W, K, p.cap);
    if ((e = cudaFuncSetAttribute(k_detect, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
      return e;
    {
      LaunchScope ls(h, KID_DETECT);
      if ((e = launch_chained(h->chain_launches, k_detect, dim3(B), dim3(kDetectThreads), smem, st, 1, p)) != cudaSuccess) return e;
    }
    if (streaming) {
      if ((e = cudaFuncSetAttribute(k_desc_planes<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_planes)) != cudaSuccess)
        return e;
      if ((e = cudaFuncSetAttribute(k_desc_planes<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_planes)) != cudaSuccess)
        return e;
      TcSink sk;
      if (sink) {
        sk = *sink;
        sk.xb = reinterpret_cast<unsigned short*>(sk.xb) + (size_t)b0 * sk.cap * 256;
        sk.nrm += (size_t)b0 * sk.cap;
        sk.opmax += b0;
      }
      const int rows = sk.xb ? sk.cap : K;
      // Optional grouping of images (tuning knob).  Measured on B200 at 148 images: one launch over the whole
      // batch is fastest (groups of 64 / 32 / 16 images: +6 % / +12 % / +24 % step time) -- launch tails cost more
      // than keeping the [256][K] scratch L2-resident saves.
      int grp = 1 << 30;
      if (const char* env = getenv("SPVO_DESC_GROUP")) grp = atoi(env) > 0 ? atoi(env) : grp;
      for (int g0 = 0; g0 < B; g0 += grp) {
        const int gb = min(grp, B - g0);
        {
          LaunchScope ls(h, KID_DESC_PLANES);
          if (in_f16)
            e = launch_chained(h->chain_launches, k_desc_planes<__half>, dim3(256 / kCP, gb), dim3(256), smem_planes, st, 1,
                               reinterpret_cast<const __half*>(desc) + (size_t)g0 * 256 * cells, kp_par + (size_t)g0 * K,
                               n_out + g0, tmp + (size_t)g0 * 256 * Kp, cells, K, plane_pitch, Kp);
          else
            e = launch_chained(h->chain_launches, k_desc_planes<float>, dim3(256 / kCP, gb), dim3(256), smem_planes, st, 1,
                               reinterpret_cast<const float*>(desc) + (size_t)g0 * 256 * cells, kp_par + (size_t)g0 * K,
                               n_out + g0, tmp + (size_t)g0 * 256 * Kp, cells, K, plane_pitch, Kp);
          if (e != cudaSuccess) return e;
        }
        TcSink sg = sk;
        if (sg.xb) {
          sg.xb = reinterpret_cast<unsigned short*>(sg.xb) + (size_t)g0 * sg.cap * 256;
          sg.nrm += (size_t)g0 * sg.cap;
          sg.opmax += g0;
        }
        LaunchScope ls(h, KID_DESC_NORM);
        // the stereo pipeline's match problems ride along when this one launch covers the whole batch
        StereoSetup ss;
        if (h->stereo_setup.probs && b0 == 0 && g0 == 0 && gb == B && B == 2 * h->stereo_setup.F) {
          ss = h->stereo_setup;
          h->stereo_setup_done = true;
        }
        if ((e = launch_chained(h->chain_launches, k_desc_normalize, dim3((rows + 31) / 32, gb), dim3(256), 0, st, 1,
                                tmp + (size_t)g0 * 256 * Kp, n_out + g0, desc_out + (size_t)g0 * K * 256, K, Kp, sg, ss)) !=
            cudaSuccess)
          return e;
      }
    } else if (desc && desc_out) {
      dim3 g3((K + 7) / 8, B);
      LaunchScope ls(h, KID_SAMPLE_DESC);
      if (in_f16)
        k_sample_desc<__half><<<g3, 256, 0, st>>>(reinterpret_cast<const __half*>(desc), kpts, n_out, desc_out, H, W, K);
      else
        k_sample_desc<float><<<g3, 256, 0, st>>>(reinterpret_cast<const float*>(desc), kpts, n_out, desc_out, H, W, K);
    }
  } else {
    if ((e = cudaMemsetAsync(n_out, 0, (size_t)B * sizeof(int), st)) != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

// Large batches are decoded as sub-batches alternating over two auxiliary streams: the intermediates of a
// sub-batch (heatmap, un-normalised descriptors) then stay L2-resident between its kernels, and the
// latency-bound per-image k_detect of one sub-batch overlaps the bandwidth-bound kernels of the other.
cudaError_t launch_decode(Handle* h, const void* semi, const void* desc, int in_f16, int B, int H, int W,
                          const spvo_decode_cfg& cfg, spvo_keypoint* kpts, float* desc_out, int* n_out,
                          float* scores, const TcSink* sink, bool* sink_filled) {
  if (sink_filled) *sink_filled = false;
  if (B == 0) return cudaSuccess;
  int nsb = h->decode_subbatches > 0 ? h->decode_subbatches : 1;  // measured on B200: 2 sub-batches gain 1.5 %, more lose (k_detect is latency-bound per image)
  if (const char* env = getenv("SPVO_DECODE_SUBBATCHES")) nsb = atoi(env) > 0 ? atoi(env) : nsb;  // tuning knob
  if (nsb > B) nsb = B;
  if (nsb <= 1) return launch_decode_range(h, semi, desc, in_f16, 0, B, H, W, cfg, kpts, desc_out, n_out, scores, sink, sink_filled);
  cudaError_t e;
  if ((e = ensure_aux_streams(h)) != cudaSuccess) return e;
  cudaStream_t main_st = h->stream;
  if ((e = cudaEventRecord(h->aux_fork, main_st)) != cudaSuccess) return e;
  for (int i = 0; i < 2; ++i)
    if ((e = cudaStreamWaitEvent(h->aux_stream[i], h->aux_fork, 0)) != cudaSuccess) return e;
  cudaError_t rc = cudaSuccess;
  for (int sb = 0; sb < nsb && rc == cudaSuccess; ++sb) {
    const int b0 = (int)((long long)B * sb / nsb), b1 = (int)((long long)B * (sb + 1) / nsb);
    h->stream = h->aux_stream[sb & 1];
    rc = launch_decode_range(h, semi, desc, in_f16, b0, b1 - b0, H, W, cfg, kpts, desc_out, n_out, scores, sink, sink_filled);
  }
  h->stream = main_st;
  for (int i = 0; i < 2; ++i) {
    if ((e = cudaEventRecord(h->aux_done[i], h->aux_stream[i])) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(main_st, h->aux_done[i], 0)) != cudaSuccess) return e;
  }
  return rc;
}

size_t decode_smem_required(int H, int W, int K) { return detect_smem_bytes(H, W, K, K <= 1536 ? 4096 : 8192); }
size_t decode_list_bytes_per_image() { return (size_t)kListCap * sizeof(u64); }

}  // namespace spvo

#ifdef SPVO_PHASE_TIMING
extern "C" int spvo_debug_phase_clocks(long long* out) {  // [64][16] clock64 stamps of k_detect's phases; diagnostic builds only
  return cudaMemcpyFromSymbol(out, spvo::g_phase_clk, sizeof(long long) * 64 * 32) == cudaSuccess ? 0 : -1;
}
#endif
