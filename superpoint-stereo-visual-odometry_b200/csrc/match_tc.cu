// match_tc.cu -- tensor-core matcher (SPVO_MATCHER_TENSOR): tcgen05 / TMEM / TMA on sm_100a.
//
// Same contract and same results as match.cu (the exact fp32 anchor) -- cv::BFMatcher(NORM_L2)
// semantics of FeatureFrontEnd::matchDescriptors (reference: src/odml_visual_odometry/src/
// feature_detection_base.cpp:434-500) -- but the N x M x 256 contraction runs on the 5th-gen
// tensor cores:
//
//   k_tc_prep       fp32 descriptors -> 16-bit rows (bf16) + fp32 squared norms in a row-padded workspace.
//                   Not used by the stereo pipeline: k_desc_normalize (decode.cu) writes fp16 operands of
//                   its unit-norm descriptors straight into the workspace.
//   k_tc_gemm       persistent, one CTA per SM over (problem, 128-row block) items: ONE Gram matrix per match,
//                   S = A . B^T with tcgen05.mma (cta_group::1, M = 128, N = 256, K = 16, 16-bit -> fp32 in
//                   TMEM), A block resident per item, B streamed by TMA (128 B swizzle) as a ring of 32 KB
//                   k-blocks, two 256-column accumulators in TMEM; 8 epilogue warps read TMEM with
//                   tcgen05.ld.16x256b and reduce every tile in BOTH directions IN REGISTERS: per row the two
//                   smallest g_ij = |b_j|^2 - 2 S_ij of each thread's columns, per column the two smallest
//                   h_ij = |a_i|^2 - 2 S_ij over the block's rows (cross-check), as packed 32-bit keys, each list
//                   with a proved lower bound on everything it does not name (the distance matrix never exists
//                   in memory, and the reverse direction costs no second GEMM)
//   k_tc_triage     one thread per row / per train column: merges the sub-lists; entries whose runner-up (listed
//                   or bounded) is outside the proved error bound are decided with no further arithmetic; the
//                   rest are queued with a normalised shortlist
//   k_tc_rerank     queued rows: every shortlisted column within the bound gets its distance recomputed
//                   in OpenCV's fp32 operation order (bit-exact DMatch.distance, first-index ties); rows
//                   whose unlisted columns are not provably outside the bound go to the fallback worklist
//   k_tc_fallback   fp32 dot products of the queued rows against every column (8 rows per pass), then
//                   exact distances for the columns within 2*eps32; if even that could be incomplete,
//                   an exact scan of the whole row
//   k_tc_fill_dist  exact DMatch.distance only for the matches that survive (cross-check)
// Match indices therefore stay bit-exact although the GEMM runs in 16-bit.
//
// Error bound.  bf16 rounding (RN) has relative error <= 2^-9 per element, so
// |a.b - bf16(a).bf16(b)| <= (2^-8 + 2^-18)|a||b|; fp32 accumulation of 256 exact products adds
// <= 256 * 2^-22 |a||b|.  With eps = eps_rel*|a||b| + kEpsAbs*(|a|^2+|b|^2) (eps_rel = 0.0085 covers
// 2x the dot-product bound; 0.0022 for fp16 operands, 2^-11 per element), |approx d^2 - exact d^2| <= eps.
// A column whose approx d^2 exceeds the row minimum by more than 2*eps (plus the key truncation) cannot be
// the exact minimum (nor tie with it).
//
// Cross-check needs the reverse nearest neighbour as well (BASE:462-463): it comes from the column direction of the
// same accumulator tiles; D(a,b) is bitwise symmetric in OpenCV's arithmetic.  Downstream (triage, rerank, fallback)
// the train -> query direction is handled as a second "directed problem" dp = P + p with the operand roles swapped.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cooperative_groups.h>
#include <limits.h>

#include "common.cuh"

namespace spvo {

constexpr int kDim = SPVO_DESC_DIM;       // 256
constexpr int kBM = 128;                  // rows of A per work item = tcgen05 M
constexpr int kBN = 256;                  // columns (rows of B) per tile = tcgen05 N (the largest the instruction takes)
constexpr int kCapAlign = 256;            // operand slots are padded to whole B tiles
constexpr int kKB = 64;                   // k-block = one 128-byte swizzle atom of 16-bit elements
constexpr int kNumKB = kDim / kKB;        // 4
constexpr int kTileBytes = kBM * kKB * 2; // 16 KB per (128 rows x 64 k) box = one TMA load
constexpr int kBSlotBytes = kBN * kKB * 2; // 32 KB per B k-block (256 columns x 64 k) = two boxes
constexpr int kBSlots = 4;                // B k-blocks in flight (ring) = one whole tile ahead
constexpr int kAccStages = 2;             // accumulator stages in TMEM
constexpr int kTmemCols = kAccStages * kBN;  // 512: all of TMEM (one CTA per SM)
constexpr int kTop = 3;
constexpr int kEpiWarps = 8;              // one per (TMEM lane quadrant, 128-column half of the tile)
constexpr int kTcThreads = 64 + 32 * kEpiWarps;  // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-9: epilogue
// |approx d^2 - exact d^2| <= eps_rel*|a||b| + kEpsAbs*(|a|^2+|b|^2):
//   bf16 operands (any CV_32F input):        2*(2^-8  + 2^-18 + 2^-14) -> 0.0085
//   fp16 operands (|x| <= 1, unit-norm rows): 2*(2^-10 + 2^-22 + 2^-14) -> 0.0022; fp16 subnormals (|x| < 2^-14) add
//   <= 2*256*2^-25 per unit of max|b_k| -> covered by kEpsAbs.
constexpr float kEpsRelBf16 = 0.0085f, kEpsRelFp16 = 0.0022f, kEpsAbs = 1e-5f;
// Key values live in [2, 8) after an affine map (tc_scale); their fp32 arithmetic (two fused multiply-adds, the
// affine map and its inverse) is off by < 1e-6 per value in those units, and the triage keeps 19 of the 24 value
// bits (rounded down, < 3.1e-5): kKeySlack covers a comparison of two.
constexpr float kKeySlack = 6e-5f;
// norm of a padded row / column: NaN, so that every key built from it is the canonical NaN 0x7FFFFFFF -> the largest
// 24-bit key value (kNone24) -- padded entries never enter a shortlist
__device__ __forceinline__ float pad_norm() { return __uint_as_float(0x7FFFFFFFu); }

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// For the one-lane producer / MMA-issuer warps: back off between polls so that their spinning does not take issue
// slots from the two epilogue warps that share their scheduler (ncu: 17 % of the kernel's issued instructions were
// SYNCS / BRA / YIELD of these loops).
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(2000u)
        : "memory");
    if (ok) break;
    __nanosleep(64);
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B
// apart (SBO), start address / 16 in the low 14 bits, descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// Instruction descriptor: D = F32, A = B = BF16, both K-major, N = 128 (>>3), M = 128 (>>4).
// a_format / b_format: 0 = F16, 1 = BF16.
__host__ __device__ constexpr uint32_t make_idesc(bool fp16) {
  return (1u << 4) | ((fp16 ? 0u : 1u) << 7) | ((fp16 ? 0u : 1u) << 10) | ((uint32_t)(kBN >> 3) << 17) |
         ((uint32_t)(kBM >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// k_tc_prep: operand o = 2p (query of problem p) or 2p+1 (train).  One warp per workspace row.
// Rows >= n are zero with norm = NaN (pad_norm), so padded rows / columns never enter a shortlist.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_tc_prep(const MatchProblem* __restrict__ probs, __nv_bfloat16* __restrict__ xb, float* __restrict__ nrm,
          unsigned* __restrict__ opmax, int cap, int fp16) {
  const int o = blockIdx.y, p = o >> 1;
  const MatchProblem pr = probs[p];
  const float* src = (o & 1) ? pr.t : pr.q;
  const int n = (o & 1) ? pr.M : pr.N;
  const int slot = (o & 1) ? pr.b_op : pr.a_op;
  const int lane = threadIdx.x & 31;
  const int rbase = (blockIdx.x * 8 + (threadIdx.x >> 5)) * 4;  // 4 rows per warp: 8 loads in flight per lane
  if (rbase >= cap) return;
  float4 a[4], b[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    a[u] = b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rbase + u < n) {
      a[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)(rbase + u) * kDim) + 2 * lane);
      b[u] = __ldg(reinterpret_cast<const float4*>(src + (size_t)(rbase + u) * kDim) + 2 * lane + 1);
    }
  }
  float smax = 0.f;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int r = rbase + u;
    if (r >= cap) break;
    const size_t row = (size_t)slot * cap + r;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(a[u].x, a[u].y), h1 = __floats2bfloat162_rn(a[u].z, a[u].w);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(b[u].x, b[u].y), h3 = __floats2bfloat162_rn(b[u].z, b[u].w);
    uint4 packed;
    packed.x = *reinterpret_cast<uint32_t*>(&h0);
    packed.y = *reinterpret_cast<uint32_t*>(&h1);
    packed.z = *reinterpret_cast<uint32_t*>(&h2);
    packed.w = *reinterpret_cast<uint32_t*>(&h3);
    if (fp16) {
      __half2 f0 = __floats2half2_rn(a[u].x, a[u].y), f1 = __floats2half2_rn(a[u].z, a[u].w);
      __half2 f2 = __floats2half2_rn(b[u].x, b[u].y), f3 = __floats2half2_rn(b[u].z, b[u].w);
      packed.x = *reinterpret_cast<uint32_t*>(&f0);
      packed.y = *reinterpret_cast<uint32_t*>(&f1);
      packed.z = *reinterpret_cast<uint32_t*>(&f2);
      packed.w = *reinterpret_cast<uint32_t*>(&f3);
    }
    float s = a[u].x * a[u].x + a[u].y * a[u].y + a[u].z * a[u].z + a[u].w * a[u].w + b[u].x * b[u].x +
              b[u].y * b[u].y + b[u].z * b[u].z + b[u].w * b[u].w;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    reinterpret_cast<uint4*>(xb + row * kDim)[lane] = packed;
    if (lane == 0) nrm[row] = r < n ? s : pad_norm();
    if (r < n) smax = fmaxf(smax, s);
  }
  if (lane == 0 && smax > 0.f) atomicMax(&opmax[slot], __float_as_uint(smax));
}

// ------------------------------------------------------------------------------------------------
// k_tc_gemm
//
// ONE Gram matrix per match.  S = A . B^T is computed once per (problem, 128-row block); the epilogue reduces
// every accumulator tile in BOTH directions:
//   rows    (query i -> nearest train j):  g_ij = |b_j|^2 - 2 S_ij   (+ constant in i)
//   columns (train j -> nearest query i):  h_ij = |a_i|^2 - 2 S_ij   (+ constant in j)   [cross-check only]
// so cv::BFMatcher's crossCheck (BASE:462-463: the reverse nearest neighbour) costs no second GEMM.
//
// Tile shape: 128 rows of A resident in shared memory per work item, B streamed in 256-column tiles, one
// tcgen05.mma = 128 x 256 x 16 (scripts/mma_microbench.cu: ~210 clk on B200, 1.6x the flops per clock of 128 x 128).
// B moves as a ring of four 32 KB k-blocks, one whole tile ahead of the MMAs; two 256-column fp32 accumulators
// fill TMEM.
//
// Epilogue (8 warps: TMEM lane quadrant x 128-column half).  The accumulator is read with tcgen05.ld.16x256b: a
// thread then holds FOUR rows (g8, g8+8, g8+16, g8+24 of its quadrant; g8 = lane / 4) of SIXTEEN columns
// (8n + 2q + {0,1}; q = lane % 4) per 64-column chunk.  That shape makes both reductions cheap:
//   * rows: each thread keeps a running top-2 per row over ITS columns (4 threads x 2 halves = 8 sub-lists per row);
//   * columns: top-2 over the thread's 4 rows in registers, then a 3-level halving butterfly across the 8 lanes that
//     hold the other rows (28 shuffles per 64 outputs), then the 4 quadrant warps merge through shared memory.
// Keys are 32 bits: (bits(v) << 8) | id with v in [2, 8) -- the top 8 bits of such a float are constant, so the
// shift keeps the whole 24-bit value and frees 8 bits for the id (local column / row in block); the packing is ONE
// integer multiply-add (IMAD, FMA pipe).  The previous kernel spent 6 of its 8 instructions per output on the ALU
// pipe (ncu: ALU 86 % busy -- the bound); this one spends ~5.8 per output for both directions together.
// A sub-list's second entry bounds everything it did not list, so the shortlist comes with a proved lower bound on
// all unlisted columns (rows): k_tc_triage / k_tc_rerank use it instead of a fixed-size top-3.
// ------------------------------------------------------------------------------------------------
constexpr int kNormRing = 4;  // tile gt's column norms live in slot gt % 4 (see the producer for why 4 is safe)
constexpr uint32_t kNone24 = 0xFFFFFFu;  // 24-bit key value of "no entry" (NaN norm of a padded row / column)
constexpr float kOffUnit = 3.25f;        // unit-norm operands: v = |x|^2 - 2 S + 3.25 in [2.2, 6.3]

struct __align__(16) TcShared {
  float nrm[kNormRing][kBN];
  uint2 colstage[2][2][4][kBN / 2];  // [tile parity][half][quadrant][column in half] -> (k0, k1) of 32 rows
  float ytile[2][kBN];               // [tile parity][column]: the columns' keypoint rows (row-band mask only)
  uint64_t a_full[kNumKB], a_empty;
  uint64_t nrm_full[kNormRing];  // column norms of tile gt landed in slot gt % kNormRing
  uint64_t nrm_empty[kNormRing]; // ... and every epilogue warp has read them (explicit edge for the slot's reuse by tile gt+4)
  uint64_t b_full[kBSlots], b_empty[kBSlots];
  uint64_t acc_full[kAccStages], acc_empty[kAccStages];
  uint32_t tmem_base;
};

// Row sub-list of one (row, 128-column half, lane q): the two smallest keys over that thread's columns of every tile
// as compact keys (20-bit value << 12) | column id (5-bit tile, 16 hh + 2 n + e, 2-bit q is the record index), and a
// lower bound (20-bit value, rounded down) on every column of the sub-list that is NOT one of the two.
// kNone20 = no entry.  Values lose their 4 low bits here (2^-18 .. 2^-17 of the key range): covered by kKeySlack.
struct __align__(16) RowRec {
  uint32_t k0, k1, bound, pad;
};
constexpr uint32_t kNone20 = 0xFFFFFu;
// Column record of one (problem, row block, train column): the two smallest keys over the block's 128 rows,
// (24-bit value << 8) | row in block, and the 24-bit value bounding every other row of the block from below.
struct __align__(16) ColRec {
  uint32_t k0, k1, bound, pad;
};
// Normalised shortlist handed from k_tc_triage to k_tc_rerank (unscaled: g = |b|^2 - 2 a.b as the 16-bit GEMM saw it).
struct __align__(16) Short {
  float g[3];
  float bound;
  int j[3];
  int pad;
};

// Affine map of the key value into [2.25, 7.75]: v = (|x|^2 - 2 S) * sc + off, with |x|^2 the norm that varies along
// the reduction (columns' for the row direction).  Unit-norm operands (the stereo pipeline) use constants.
struct TcScale {
  float sc, off, m2sc;
};
__device__ __forceinline__ TcScale tc_scale(float selfmax2, float othermax2, bool unit) {
  TcScale t;
  if (unit) {
    t.sc = 1.0f;
    t.off = kOffUnit;
    t.m2sc = -2.0f;
  } else {
    const float r2 = 2.02f * sqrtf(selfmax2 * othermax2);  // |2 S| <= r2 (1.01 > (1 + 2^-9)^2: bf16 rounding)
    const float w = fmaxf(selfmax2 + 2.0f * r2, 1e-30f);
    t.sc = 5.5f / w;
    t.off = 2.25f + r2 * t.sc;
    t.m2sc = -2.0f * t.sc;
  }
  return t;
}
__device__ __forceinline__ float key_value(uint32_t v24) { return __uint_as_float(0x40000000u | v24); }

// One arrival per WARP: 256 per-thread arrivals on one mbarrier serialise.
__device__ __forceinline__ void epi_release(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

// sorted pair (a0 <= a1) <- two smallest of {a0, a1, b0, b1}, (b0 <= b1)
__device__ __forceinline__ void merge_pairs(uint32_t& a0, uint32_t& a1, uint32_t b0, uint32_t b1) {
  const uint32_t t = max(a0, b0);
  a0 = min(a0, b0);
  a1 = min(min(t, a1), b1);
}
// running pair (k0 <= k1) <- two smallest of {k0, k1, x, y}
__device__ __forceinline__ void push_two(uint32_t& k0, uint32_t& k1, uint32_t x, uint32_t y) {
  const uint32_t lo = min(x, y), hi = max(x, y);
  const uint32_t t = max(k0, lo);
  k0 = min(k0, lo);
  k1 = min(min(t, k1), hi);
}
// One halving level of the column butterfly: 2*C columns held as (k0[c], k1[c]); lanes with `upper` keep columns
// C .. 2C-1, the others 0 .. C-1; the kept columns end in slots 0 .. C-1 merged with the partner lane's lists.
template <int C>
__device__ __forceinline__ void col_butterfly(uint32_t* k0, uint32_t* k1, bool upper, int lane_xor) {
#pragma unroll
  for (int c = 0; c < C; ++c) {
    uint32_t m0 = upper ? k0[c + C] : k0[c], m1 = upper ? k1[c + C] : k1[c];
    const uint32_t s0 = upper ? k0[c] : k0[c + C], s1 = upper ? k1[c] : k1[c + C];
    const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, lane_xor), r1 = __shfl_xor_sync(0xffffffffu, s1, lane_xor);
    merge_pairs(m0, m1, r0, r1);
    k0[c] = m0;
    k1[c] = m1;
  }
}

// Work item = (problem p, 128-row block rb): A = the problem's queries, B = its train descriptors.  The kernel is
// PERSISTENT: one CTA per SM walks items bid, bid + grid, ...; barriers and TMEM are set up once, and the B ring /
// accumulator pipelines run continuously across items (global k-block and tile counters give slot and phase).
struct TcItem {
  int p, rb, Na, Nb, a_op, b_op, nct;
  bool valid;
  const float* qy;
  const float* ty;
  int ystride;
  float band;
};
__device__ __forceinline__ TcItem tc_item(const MatchProblem* __restrict__ probs, int w, int nrb) {
  TcItem it;
  it.p = w / nrb;
  it.rb = w - it.p * nrb;
  const MatchProblem pr = probs[it.p];
  it.Na = pr.N;
  it.Nb = pr.M;
  it.a_op = pr.a_op;
  it.b_op = pr.b_op;
  it.nct = (it.Nb + kBN - 1) / kBN;
  it.valid = it.rb * kBM < it.Na && it.nct > 0;
  it.qy = pr.qy;
  it.ty = pr.ty;
  it.ystride = pr.ystride;
  it.band = pr.qy ? pr.band : INFINITY;  // unmasked problems of a masked launch: everything is allowed
  return it;
}

template <bool kUnit, bool kCols, bool kMask>
__global__ void __launch_bounds__(kTcThreads, 1)
k_tc_gemm(const __grid_constant__ CUtensorMap tmap, const MatchProblem* __restrict__ probs,
          const float* __restrict__ nrm, const unsigned* __restrict__ opmax, RowRec* __restrict__ row_rec,
          ColRec* __restrict__ col_rec, int cap, uint32_t idesc, int n_items, uint32_t r256,
          int* __restrict__ zero_counters, int n_zero) {
  chain_enter();
  // the rerank / fallback worklist counters of this call (used from k_tc_triage on): cleared here instead of by a
  // memset node, which would break the chain of programmatic launches
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_zero; i += gridDim.x * blockDim.x) zero_counters[i] = 0;
  extern __shared__ uint8_t smem_raw[];
  const int nrb = cap / kBM;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // carve shared memory: operands need 1024 B alignment for the 128 B swizzle
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = base;                        // 4 k-blocks x 16 KB
  const uint32_t sB = base + kNumKB * kTileBytes;  // ring of kBSlots x 32 KB k-blocks
  TcShared* sh = reinterpret_cast<TcShared*>(smem_raw + (sB + kBSlots * kBSlotBytes - smem_u32(smem_raw)));

  if (threadIdx.x == 0) {
    for (int kb = 0; kb < kNumKB; ++kb) mbar_init(smem_u32(&sh->a_full[kb]), 1);
    mbar_init(smem_u32(&sh->a_empty), 1);
    for (int s = 0; s < kNormRing; ++s) {
      mbar_init(smem_u32(&sh->nrm_full[s]), 1);
      mbar_init(smem_u32(&sh->nrm_empty[s]), kEpiWarps);
    }
    for (int s = 0; s < kBSlots; ++s) {
      mbar_init(smem_u32(&sh->b_full[s]), 1);
      mbar_init(smem_u32(&sh->b_empty[s]), 1);
    }
    for (int s = 0; s < kAccStages; ++s) {
      mbar_init(smem_u32(&sh->acc_full[s]), 1);
      mbar_init(smem_u32(&sh->acc_empty[s]), kEpiWarps);  // one elected arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM (all 512 columns): 2 accumulator stages x 256 fp32 columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sh->tmem_base)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = sh->tmem_base;

  if (warp == 0) {
    // ===== TMA producer (one elected lane) =====
    if (lane == 0) {
      uint32_t gt = 0, gk = 0, ai = 0;  // global tile / k-block counters, count of items that loaded A
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const TcItem it = tc_item(probs, w, nrb);
        if (!it.valid) continue;
        auto load_a = [&]() {
          mbar_wait_relaxed(smem_u32(&sh->a_empty), (ai & 1) ^ 1);  // previous item's MMAs no longer read the A block
          ++ai;
          const int a_row = it.a_op * cap + it.rb * kBM;
          for (int kb = 0; kb < kNumKB; ++kb) {
            mbar_expect_tx(smem_u32(&sh->a_full[kb]), kTileBytes);
            tma_load_2d(sA + kb * kTileBytes, &tmap, smem_u32(&sh->a_full[kb]), kb * kKB, a_row);
          }
        };
        for (int ct = 0; ct < it.nct; ++ct, ++gt) {
          // A is requested AFTER the item's first B tile: that tile only needs free ring slots, so it is in
          // flight while the previous item's last MMAs (which still read the A block) retire
          if (ct == 1) load_a();
          const int b_row = it.b_op * cap + ct * kBN;
          // the tile's 256 column norms: 1-D bulk copy into slot gt % 4 with its own mbarrier (the epilogue waits
          // on it directly).  The slot is rewritten by tile gt+4; the epilogue warps signal nrm_empty when they
          // have read tile gt's norms (the B ring / accumulator chain orders this transitively as well).
          const uint32_t nb_bar = smem_u32(&sh->nrm_full[gt % kNormRing]);
          for (int kb = 0; kb < kNumKB; ++kb, ++gk) {
            const int s = gk % kBSlots, ph = (gk / kBSlots) & 1;
            mbar_wait_relaxed(smem_u32(&sh->b_empty[s]), ph ^ 1);
            mbar_expect_tx(smem_u32(&sh->b_full[s]), kBSlotBytes);
            for (int hb = 0; hb < kBN / kBM; ++hb)  // the tensor map's box is 128 rows: two boxes per k-block
              tma_load_2d(sB + s * kBSlotBytes + hb * kTileBytes, &tmap, smem_u32(&sh->b_full[s]), kb * kKB,
                          b_row + hb * kBM);
            if (kb == 0) {
              mbar_wait(smem_u32(&sh->nrm_empty[gt % kNormRing]), ((gt / kNormRing) & 1) ^ 1);  // tile gt-4's norms are read
              mbar_expect_tx(nb_bar, kBN * 4);
              bulk_load_1d(smem_u32(&sh->nrm[gt % kNormRing][0]), nrm + (size_t)b_row, kBN * 4, nb_bar);
            }
          }
          if (it.nct == 1) load_a();
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane) =====
    if (lane == 0) {
      uint32_t gt = 0, gk = 0, ai = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
        const TcItem it = tc_item(probs, w, nrb);
        if (!it.valid) continue;
        for (int ct = 0; ct < it.nct; ++ct, ++gt) {
          const int sa = gt % kAccStages, pha = (gt / kAccStages) & 1;
          mbar_wait_relaxed(smem_u32(&sh->acc_empty[sa]), pha ^ 1);  // epilogue drained this accumulator stage
          const uint32_t d = tmem_base + sa * kBN;
#pragma unroll
          for (int kb = 0; kb < kNumKB; ++kb, ++gk) {
            const int s = gk % kBSlots, ph = (gk / kBSlots) & 1;
            if (ct == 0) mbar_wait(smem_u32(&sh->a_full[kb]), ai & 1);  // this k-slice of the A block landed
            mbar_wait(smem_u32(&sh->b_full[s]), ph);                    // TMA landed this B k-block
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < kKB / 16; ++k) {
              const uint64_t ad = umma_desc_sw128(sA + kb * kTileBytes + k * 32);
              const uint64_t bd = umma_desc_sw128(sB + s * kBSlotBytes + k * 32);
              tc_mma_bf16(d, ad, bd, idesc, (kb | k) != 0);
            }
            tc_commit(smem_u32(&sh->b_empty[s]));  // ring slot free once these 4 MMAs retire
          }
          tc_commit(smem_u32(&sh->acc_full[sa]));  // accumulator ready for the epilogue
        }
        ++ai;
        tc_commit(smem_u32(&sh->a_empty));  // A block free once the item's last MMAs retire
      }
    }
  } else {
    // ===== epilogue: warps 2..9.  TMEM lane quadrant = warp % 4; 128-column half of the tile = (warp - 2) / 4 =====
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const int q = lane & 3, g8 = lane >> 2;
    uint32_t gt = 0;
    float na_pf[4] = {0.f, 0.f, 0.f, 0.f};  // row norms of item pf_w, fetched while the previous item was processed
    int pf_w = -1;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      const TcItem it = tc_item(probs, w, nrb);
      if (!it.valid) continue;
      const float amax2 = __uint_as_float(opmax[it.a_op]), bmax2 = __uint_as_float(opmax[it.b_op]);
      const TcScale sr = tc_scale(bmax2, amax2, kUnit);  // row direction: the column norms vary
      const TcScale sc = tc_scale(amax2, bmax2, kUnit);  // column direction: the row norms vary
      // this thread's four rows: row-in-block ids and the additive term of the column-direction value
      float ca[4];
      uint32_t rid[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        rid[m] = (uint32_t)(quad * 32 + g8 + 8 * m);
        ca[m] = 0.f;
        if (kCols) {
          const float na = pf_w == w ? na_pf[m] : __ldg(nrm + (size_t)it.a_op * cap + it.rb * kBM + rid[m]);
          ca[m] = __fmaf_rn(na, sc.sc, sc.off);
        }
      }
      float yi[4] = {0.f, 0.f, 0.f, 0.f};  // keypoint rows (image y) of this thread's four rows: row-band mask
      if (kMask && it.qy) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int row = it.rb * kBM + (int)rid[m];
          if (row < it.Na) yi[m] = __ldg(it.qy + (size_t)row * it.ystride);
        }
      }
      if (kCols) {  // start the next valid item's norm loads now: they land long before that item begins
        pf_w = -1;
        for (int w2 = w + gridDim.x; w2 < n_items; w2 += gridDim.x) {
          const TcItem nx = tc_item(probs, w2, nrb);
          if (!nx.valid) continue;
          pf_w = w2;
#pragma unroll
          for (int m = 0; m < 4; ++m) na_pf[m] = __ldg(nrm + (size_t)nx.a_op * cap + nx.rb * kBM + rid[m]);
          break;
        }
      }
      // running row lists over this thread's columns of every tile: keys + the tiles they came from
      // rk2 = smallest key that was ever dropped from the running pair: with it, everything this thread has seen
      // but does not list is bounded below by min(rk2, rk1 if both listed entries come from the same tile)
      uint32_t rk0[4], rk1[4], rk2[4], rt[4];
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        rk0[m] = rk1[m] = rk2[m] = 0xFFFFFFFFu;
        rt[m] = 0u;
      }
      for (int ct = 0; ct < it.nct; ++ct, ++gt) {
        const int sa = gt % kAccStages, pha = (gt / kAccStages) & 1;
        mbar_wait(smem_u32(&sh->acc_full[sa]), pha);
        tc_fence_after();
        // the tile's column norms were bulk-copied to shared memory next to its B operand
        mbar_wait(smem_u32(&sh->nrm_full[gt % kNormRing]), (gt / kNormRing) & 1);
        if (kMask) {  // the 256 epilogue threads stage the tile's 256 column keypoint rows (two buffers by tile parity)
          const int cidx = (int)threadIdx.x - 64, col = ct * kBN + cidx;
          sh->ytile[gt & 1][cidx] = (it.ty && col < it.Nb) ? __ldg(it.ty + (size_t)col * it.ystride) : 0.f;
          asm volatile("bar.sync 3, 256;" ::: "memory");
        }
        uint32_t k0[4], k1[4];  // tile-local row lists (id = 16 hh + 2 n + e: this thread's column within the tile half)
#pragma unroll
        for (int m = 0; m < 4; ++m) k0[m] = k1[m] = 0xFFFFFFFFu;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const int cbase = half * 128 + hh * 64;  // first column of this 64-column chunk within the tile
          uint32_t acc[2][32];                     // [lane half][4 n + 2 (row pair) + e]
          const uint32_t taddr = tmem_base + sa * kBN + cbase + ((uint32_t)(quad * 32) << 16);
          tmem_ld_16x256b_x8(taddr, acc[0]);
          tmem_ld_16x256b_x8(taddr + (16u << 16), acc[1]);
          float cb[16];  // additive term of the row-direction value for this thread's 16 columns
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            const float2 nb = *reinterpret_cast<const float2*>(&sh->nrm[gt % kNormRing][cbase + 8 * n + 2 * q]);
            cb[2 * n] = __fmaf_rn(nb.x, sr.sc, sr.off);
            cb[2 * n + 1] = __fmaf_rn(nb.y, sr.sc, sr.off);
          }
          float yc[16];
          if (kMask) {
#pragma unroll
            for (int n = 0; n < 8; ++n) {
              const float2 y2 = *reinterpret_cast<const float2*>(&sh->ytile[gt & 1][cbase + 8 * n + 2 * q]);
              yc[2 * n] = y2.x;
              yc[2 * n + 1] = y2.y;
            }
          }
          tmem_ld_wait();
          if (hh == 1) {
            tc_fence_before();
            epi_release(smem_u32(&sh->acc_empty[sa]), lane);  // accumulator is in registers: release it to the MMA warp
            if (lane == 0) mbar_arrive(smem_u32(&sh->nrm_empty[gt % kNormRing]));  // (after the __syncwarp above) norms consumed
          }
          uint32_t c0[16], c1[16];  // column lists over this thread's four rows
#pragma unroll
          for (int n = 0; n < 8; ++n) {
            uint32_t ck[4][2];
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const float s0 = __uint_as_float(acc[m >> 1][4 * n + 2 * (m & 1)]);
              const float s1 = __uint_as_float(acc[m >> 1][4 * n + 2 * (m & 1) + 1]);
              // row direction: one IMAD packs value and column id; padded columns carry a NaN norm -> key 0xFFFFFFxx
              uint32_t ka = __float_as_uint(__fmaf_rn(s0, sr.m2sc, cb[2 * n])) * r256 + (uint32_t)(16 * hh + 2 * n);
              uint32_t kb = __float_as_uint(__fmaf_rn(s1, sr.m2sc, cb[2 * n + 1])) * r256 + (uint32_t)(16 * hh + 2 * n + 1);
              // row-band mask: a pair outside the band is nobody's candidate, in either direction
              const bool ok0 = !kMask || fabsf(__fsub_rn(yi[m], yc[2 * n])) <= it.band;
              const bool ok1 = !kMask || fabsf(__fsub_rn(yi[m], yc[2 * n + 1])) <= it.band;
              if (kMask) {
                ka = ok0 ? ka : 0xFFFFFFFFu;
                kb = ok1 ? kb : 0xFFFFFFFFu;
              }
              push_two(k0[m], k1[m], ka, kb);
              if (kCols) {
                // (run-time multiplier: with a literal 256 the compiler splits one row's packing into SHL + LOP3 on the ALU pipe)
                ck[m][0] = __float_as_uint(__fmaf_rn(s0, sc.m2sc, ca[m])) * r256 + rid[m];
                ck[m][1] = __float_as_uint(__fmaf_rn(s1, sc.m2sc, ca[m])) * r256 + rid[m];
                if (kMask) {
                  ck[m][0] = ok0 ? ck[m][0] : 0xFFFFFFFFu;
                  ck[m][1] = ok1 ? ck[m][1] : 0xFFFFFFFFu;
                }
              }
            }
            if (kCols) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {  // two smallest of the four rows
                const uint32_t lo1 = min(ck[0][e], ck[1][e]), hi1 = max(ck[0][e], ck[1][e]);
                const uint32_t lo2 = min(ck[2][e], ck[3][e]), hi2 = max(ck[2][e], ck[3][e]);
                c0[2 * n + e] = min(lo1, lo2);
                c1[2 * n + e] = min(min(max(lo1, lo2), hi1), hi2);
              }
            }
          }
          if (kCols) {
            // across the 8 lanes that hold the other rows of the same columns (lane bits 4, 3, 2): afterwards lane L
            // holds columns 2L, 2L+1 of the chunk, reduced over the warp's 32 rows
            col_butterfly<8>(c0, c1, (lane & 16) != 0, 16);
            col_butterfly<4>(c0, c1, (lane & 8) != 0, 8);
            col_butterfly<2>(c0, c1, (lane & 4) != 0, 4);
            *reinterpret_cast<uint4*>(&sh->colstage[gt & 1][half][quad][hh * 64 + 2 * lane]) =
                make_uint4(c0[0], c1[0], c0[1], c1[1]);
          }
        }
        // fold the tile's row lists into the running ones, remembering the tile of each entry
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const uint32_t ta0 = rt[m] & 0xFFFFu, ta1 = rt[m] >> 16;
          const bool f = k0[m] < rk0[m];
          const uint32_t n0 = f ? k0[m] : rk0[m], tn0 = f ? (uint32_t)ct : ta0;
          const uint32_t x = f ? rk0[m] : k0[m], tx = f ? ta0 : (uint32_t)ct;
          const uint32_t y = f ? k1[m] : rk1[m], ty = f ? (uint32_t)ct : ta1;
          const uint32_t z = f ? rk1[m] : k1[m];  // second entry of the list whose first lost (z >= x)
          const bool s = x < y;
          rk0[m] = n0;
          rk1[m] = s ? x : y;
          rt[m] = tn0 | ((s ? tx : ty) << 16);
          rk2[m] = min(rk2[m], min(max(x, y), z));  // third smallest of the four = the smallest key dropped here
        }
        if (kCols) {
          // the four quadrant warps of this half merge their column lists: warp `quad` finishes 32 of the 128 columns.
          // Two staging buffers (tile parity): a warp refills buffer b only after passing the next tile's barrier,
          // i.e. after every warp has read tile gt's lists.
          asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
          const int cc = 32 * quad + lane;
          uint2 ql[4];  // (first, second) of each quadrant's 32 rows
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) ql[qq] = sh->colstage[gt & 1][half][qq][cc];
          uint2 mine = ql[0];
#pragma unroll
          for (int qq = 1; qq < 4; ++qq) merge_pairs(mine.x, mine.y, ql[qq].x, ql[qq].y);
          // every row that is not one of the two listed is >= min(smallest quadrant second, third smallest quadrant
          // first): a quadrant's unlisted rows are >= its second, and a first that is not listed was dropped
          const uint32_t lo1 = min(ql[0].x, ql[1].x), hi1 = max(ql[0].x, ql[1].x);
          const uint32_t lo2 = min(ql[2].x, ql[3].x), hi2 = max(ql[2].x, ql[3].x);
          const uint32_t third = min(min(min(ql[0].y, ql[1].y), min(ql[2].y, ql[3].y)), max(max(lo1, lo2), min(hi1, hi2)));
          ColRec cr;
          cr.k0 = mine.x; cr.k1 = mine.y; cr.bound = third >> 8; cr.pad = 0;
          col_rec[((size_t)it.p * nrb + it.rb) * cap + ct * kBN + half * 128 + cc] = cr;
        }
      }
      // item end: one 16-byte record per (row, half, q): the thread's two best columns as compact keys and the bound on
      // everything else it has seen (k_tc_triage merges the row's 8 sub-lists)
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int row = it.rb * kBM + (int)rid[m];
        if (row < it.Na) {
          RowRec r;
          const uint32_t t0 = rt[m] & 0xFFFFu, t1 = rt[m] >> 16;
          r.k0 = ((rk0[m] >> 12) << 12) | (t0 << 5) | (rk0[m] & 31u);
          r.k1 = ((rk1[m] >> 12) << 12) | (t1 << 5) | (rk1[m] & 31u);
          const uint32_t b = (t0 == t1) ? min(rk2[m], rk1[m]) : rk2[m];
          r.bound = b >> 12;  // rounded down: still a lower bound
          r.pad = 0;
          row_rec[(((size_t)it.p * cap + row) * 2 + half) * 4 + q] = r;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// k_tc_triage: one THREAD per row of a directed problem (dp < P: query -> train from the row records; dp >= P, only
// with cross-check: train -> query from the column records of every row block).  Merges the sub-lists into the three
// best approximate values + a lower bound on everything unlisted, in unscaled units (g = |b|^2 - 2 a.b as the 16-bit
// GEMM saw it).  Most rows are decided by the bound alone -- the runner-up (listed or not) is farther than the proved
// error bound from the best, so the best IS the exact nearest neighbour and no distance has to be evaluated (its
// DMatch.distance is filled in later only if the match survives).  Everything else is queued, with its shortlist,
// for the warp-per-row k_tc_rerank.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rerank_row(const MatchProblem& pr, int p, int dp, bool rev, int i, const Short& sl, int mode,
                                           float ratio, const float* __restrict__ nrm, const unsigned* __restrict__ opmax,
                                           int cap, int max_rows, int max_cols, int* __restrict__ row_best,
                                           float* __restrict__ row_d, int* __restrict__ col_best, int* __restrict__ fb_count,
                                           int* __restrict__ fb_list, unsigned long long* __restrict__ counters,
                                           float eps_rel, int unit);

__global__ void __launch_bounds__(256)
k_tc_triage(const MatchProblem* __restrict__ probs, int P, int mode, const float* __restrict__ nrm,
            const unsigned* __restrict__ opmax, const RowRec* __restrict__ row_rec, const ColRec* __restrict__ col_rec,
            int cap, int max_rows, int max_cols, int* __restrict__ row_best, float* __restrict__ row_d,
            int* __restrict__ col_best, int* __restrict__ rr_count, int* __restrict__ rr_list,
            Short* __restrict__ shortl, float eps_rel, int unit, int fuse, float ratio, int* __restrict__ fb_count,
            int* __restrict__ fb_list, unsigned long long* __restrict__ counters) {
  chain_enter();
  // fuse != 0 (the NN modes): the few rows this block cannot resolve are re-ranked by the block itself right away
  // (rerank_row, one warp per row) instead of travelling through global memory to a k_tc_rerank launch
  __shared__ Short s_sl[256];
  __shared__ int s_row[256];
  __shared__ int s_n;
  const int dp = blockIdx.y, p = dp < P ? dp : dp - P;
  const bool rev = dp >= P;
  const MatchProblem pr = probs[p];
  const int Na = rev ? pr.M : pr.N, Nb = rev ? pr.N : pr.M;
  const int a_op = rev ? pr.b_op : pr.a_op, b_op = rev ? pr.a_op : pr.b_op;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (blockIdx.x * 256 >= Na) return;  // whole block beyond the problem's rows (uniform)
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  const bool active = i < Na;
  bool resolved = false;
  int j0 = -1;
  Short sl;
  if (!active) {
    resolved = true;
  } else if (Nb == 0) {
    resolved = true;
  } else {
    // entries: 32-bit keys (19-bit value << 13) | column -- values rounded DOWN from the records' 20 / 24 bits (covered
    // by kKeySlack); ~0 = none.  The bound stays 24-bit.
    uint32_t e0 = ~0u, e1 = ~0u, e2 = ~0u, e3 = ~0u;
    auto insert = [&](uint32_t x) {
      uint32_t t;
      t = min(e0, x); x = max(e0, x); e0 = t;
      t = min(e1, x); x = max(e1, x); e1 = t;
      t = min(e2, x); x = max(e2, x); e2 = t;
      e3 = min(e3, x);
    };
    uint32_t bound = kNone24;
    if (!rev) {
      const uint4* rec = reinterpret_cast<const uint4*>(row_rec + ((size_t)p * cap + i) * 8);
#pragma unroll
      for (int sl8 = 0; sl8 < 8; ++sl8) {  // sub-list = (half, q)
        const uint4 r = __ldg(rec + sl8);
        const int hf = sl8 >> 2, q = sl8 & 3;
        const uint32_t kk[2] = {r.x, r.y};
#pragma unroll
        for (int z = 0; z < 2; ++z) {
          const uint32_t id = kk[z] & 31u, tile = (kk[z] >> 5) & 31u;
          const uint32_t col = tile * kBN + hf * 128 + (id >> 4) * 64 + ((id >> 1) & 7) * 8 + 2 * q + (id & 1);
          // a "none" entry (value 0xFFFFF) stays the maximum of its 19-bit field: 0x7FFFF
          insert(((kk[z] >> 13) << 13) | col);
        }
        if (r.z < kNone20) bound = min(bound, r.z << 4);
      }
    } else {
      const int nrb = cap / kBM, nrbv = (pr.N + kBM - 1) / kBM;  // row blocks of the forward problem that exist
      for (int rb = 0; rb < nrbv; ++rb) {
        const uint4 c = __ldg(reinterpret_cast<const uint4*>(col_rec + ((size_t)p * nrb + rb) * cap + i));
        insert(((c.x >> 13) << 13) | (uint32_t)(rb * kBM + (c.x & 0xFFu)));
        insert(((c.y >> 13) << 13) | (uint32_t)(rb * kBM + (c.y & 0xFFu)));
        bound = min(bound, min(kNone24, c.z));  // rows of this block beyond its two listed ones
      }
    }
    constexpr uint32_t kNone19 = 0x7FFFFu;
    if ((e3 >> 13) < kNone19) bound = min(bound, (e3 >> 13) << 5);  // listed, but beyond the three kept
    if (Nb <= kTop) bound = kNone24;                                // the list is the whole row
    const float amax = __uint_as_float(opmax[a_op]), bmax = __uint_as_float(opmax[b_op]);
    const TcScale ts = tc_scale(bmax, amax, unit != 0);  // the reduction ran over the B side's norms
    const float inv = 1.0f / ts.sc;
    const uint32_t ee[3] = {e0, e1, e2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const bool ok = (ee[k] >> 13) < kNone19;
      sl.g[k] = ok ? (key_value((ee[k] >> 13) << 5) - ts.off) * inv : INFINITY;
      sl.j[k] = ok ? (int)(ee[k] & 0x1FFFu) : -1;
    }
    sl.bound = bound < kNone24 ? (key_value(bound) - ts.off) * inv : INFINITY;
    sl.pad = 0;
    if (!(mode == SPVO_MATCH_KNN_RATIO && !rev) && sl.j[0] >= 0) {
      const float na = nrm[(size_t)a_op * cap + i];
      const float eps = eps_rel * sqrtf(na * bmax) + kEpsAbs * (na + bmax);
      const float slack = 2.0f * eps + kKeySlack * inv;
      if (fminf(sl.g[1], sl.bound) > sl.g[0] + slack) {  // same test as k_tc_rerank's "one candidate, complete"
        resolved = true;
        j0 = sl.j[0];
      }
    }
  }
  if (resolved) {
    if (!active) {
    } else if (!rev) {
      const size_t o = ((size_t)p * max_rows + i) * 2;
      row_best[o] = j0;
      row_best[o + 1] = -1;
      row_d[o] = -1.0f;  // "not evaluated yet" (k_tc_fill_dist)
      row_d[o + 1] = INFINITY;
    } else {
      col_best[(size_t)p * max_cols + i] = j0;
    }
  } else if (fuse) {
    const int slot = atomicAdd(&s_n, 1);
    s_row[slot] = i;
    s_sl[slot] = sl;
  } else {
    const int slot = atomicAdd(&rr_count[dp], 1);
    rr_list[(size_t)dp * cap + slot] = i;
    shortl[(size_t)dp * cap + slot] = sl;
  }
  if (fuse) {
    __syncthreads();
    const int n = s_n;
    for (int li = threadIdx.x >> 5; li < n; li += 8)
      rerank_row(pr, p, dp, rev, s_row[li], s_sl[li], mode, ratio, nrm, opmax, cap, max_rows, max_cols, row_best, row_d,
                 col_best, fb_count, fb_list, counters, eps_rel, unit);
  }
}

// ------------------------------------------------------------------------------------------------
// k_tc_rerank: one warp per row of a directed problem.  Two (row, column) pairs are evaluated at a
// time, 16 lanes each: lane (k = l16/4, l = l16%4) owns OpenCV's accumulator s[k][l].
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float exact_dist_half(const float* __restrict__ a, const float* __restrict__ b, int l16) {
  float s = 0.0f;
#pragma unroll
  for (int blk = 0; blk < kDim / 16; ++blk) {
    const float t = __fsub_rn(__ldg(a + blk * 16 + l16), __ldg(b + blk * 16 + l16));
    s = __fadd_rn(s, __fmul_rn(t, t));
  }
  // v[l] = ((s[0][l] + s[1][l]) + s[2][l]) + s[3][l]; lanes hold s[k][l] at l16 = 4k + l
  const int l = l16 & 3, hb = threadIdx.x & 16;
  const float s0 = __shfl_sync(0xffffffffu, s, hb + l), s1 = __shfl_sync(0xffffffffu, s, hb + 4 + l);
  const float s2 = __shfl_sync(0xffffffffu, s, hb + 8 + l), s3 = __shfl_sync(0xffffffffu, s, hb + 12 + l);
  const float vl = __fadd_rn(__fadd_rn(__fadd_rn(s0, s1), s2), s3);
  const float v0 = __shfl_sync(0xffffffffu, vl, hb + 0), v1 = __shfl_sync(0xffffffffu, vl, hb + 1);
  const float v2 = __shfl_sync(0xffffffffu, vl, hb + 2), v3 = __shfl_sync(0xffffffffu, vl, hb + 3);
  return __fsqrt_rn(__fadd_rn(__fadd_rn(v0, v2), __fadd_rn(v1, v3)));
}

__device__ __forceinline__ bool lex_less2(float d1, int i1, float d2, int i2) {
  return d1 < d2 || (d1 == d2 && i1 < i2);
}
__device__ __forceinline__ void top2_push(float d, int j, float& b0, int& x0, float& b1, int& x1) {
  if (lex_less2(d, j, b0, x0)) {
    b1 = b0; x1 = x0; b0 = d; x0 = j;
  } else if (lex_less2(d, j, b1, x1)) {
    b1 = d; x1 = j;
  }
}

constexpr int kPending = -2;  // row_best / col_best of a row queued for k_tc_fallback (see rerank_row, k_tc_fill_dist)

// One queued row of a directed problem, one WARP: exact fp32 distances (OpenCV order) for the listed candidates inside
// the proved bound; rows whose unlisted columns are not provably outside go to the fallback list.  Called by
// k_tc_rerank (rows queued through global memory: the kNN mode, where every row is evaluated) and directly by
// k_tc_triage (the NN modes: the block re-ranks the few rows it could not resolve itself).
__device__ __forceinline__ void rerank_row(const MatchProblem& pr, int p, int dp, bool rev, int i, const Short& sl, int mode,
                                           float ratio, const float* __restrict__ nrm, const unsigned* __restrict__ opmax,
                                           int cap, int max_rows, int max_cols, int* __restrict__ row_best,
                                           float* __restrict__ row_d, int* __restrict__ col_best, int* __restrict__ fb_count,
                                           int* __restrict__ fb_list, unsigned long long* __restrict__ counters,
                                           float eps_rel, int unit) {
  const int Na = rev ? pr.M : pr.N, Nb = rev ? pr.N : pr.M;
  const float* A = rev ? pr.t : pr.q;
  const float* B = rev ? pr.q : pr.t;
  const int a_op = rev ? pr.b_op : pr.a_op, b_op = rev ? pr.a_op : pr.b_op;
  const int lane = threadIdx.x & 31, l16 = lane & 15, half = lane >> 4;
  float b0 = INFINITY, b1 = INFINITY;
  int x0 = INT_MAX, x1 = INT_MAX;
  bool full = false;
  const float* arow = A + (size_t)i * kDim;
  (void)Na;
  if (Nb > 0) {
    // the row's shortlist as k_tc_triage normalised it: three best approximate values, their columns, and a lower
    // bound on every column that is not listed
    const float na = nrm[(size_t)a_op * cap + i];
    const float amax = __uint_as_float(opmax[a_op]), bmax = __uint_as_float(opmax[b_op]);
    const TcScale ts = tc_scale(bmax, amax, unit != 0);
    const float g[kTop] = {sl.g[0], sl.g[1], sl.g[2]};
    const int jx[kTop] = {sl.j[0], sl.j[1], sl.j[2]};
    // bound on |approx - exact| of any relevant column of this row, plus the rounding of the key arithmetic
    const float eps = eps_rel * sqrtf(na * bmax) + kEpsAbs * (na + bmax);
    const bool knn = (mode == SPVO_MATCH_KNN_RATIO) && !rev;
    const float slack = 2.0f * eps + kKeySlack / ts.sc;
    int nc;
    if (!knn) {
      // candidates = listed columns within the bound of the best; complete iff every unlisted column is outside it
      nc = 1;
      while (nc < kTop && jx[nc] >= 0 && g[nc] <= g[0] + slack) ++nc;
      if (!(sl.bound > g[0] + slack)) full = true;
    } else if (jx[1] < 0) {
      // a single admissible train column: the second best does not exist (the ratio test then fails in finalize);
      // none at all (every column masked out): nothing to evaluate, the row has no match
      nc = jx[0] < 0 ? 0 : 1;
    } else {
      nc = 2;
      while (nc < kTop && jx[nc] >= 0 && g[nc] <= g[1] + slack) ++nc;
      if (!(sl.bound > g[1] + slack)) {
        // exact second best unknown; the ratio decision may still be provable from the bound
        if (fminf(g[1], sl.bound) > g[0] + slack) {
          const float d0 = exact_dist_half(arow, B + (size_t)jx[0] * kDim, l16);
          const float lo2 = fminf(g[1], sl.bound) + na - eps - kKeySlack / ts.sc;  // exact second best d^2 >= lo2
          const float hi2 = g[1] + na + eps + kKeySlack / ts.sc;                    // and <= hi2 (column jx[1] exists)
          const float lo = sqrtf(fmaxf(lo2, 0.0f)) * (1.0f - 1e-6f);
          const float hi = sqrtf(fmaxf(hi2, 0.0f)) * (1.0f + 1e-6f);
          if (d0 < ratio * lo) {  // passes for any admissible second best
            b0 = d0; x0 = jx[0]; b1 = INFINITY; x1 = jx[1];
            nc = 0;
          } else if (d0 >= ratio * hi) {  // fails for any admissible second best
            b0 = d0; x0 = jx[0]; b1 = 0.0f; x1 = jx[1];
            nc = 0;
          } else {
            full = true;
          }
        } else {
          full = true;
        }
      }
    }
    if (full) {
      // shortlist not provably complete: queue the row for k_tc_fallback (which writes its result).  Until then the
      // row is marked PENDING (-2): k_tc_fill_dist may run concurrently with the fallback and must neither fill a
      // pending row's distance nor trust a pending column's cross-check
      if (lane == 0) {
        fb_list[(size_t)dp * cap + atomicAdd(&fb_count[dp], 1)] = i;
        atomicAdd(&counters[1], 1ull);
        if (!rev) {
          const size_t o = ((size_t)p * max_rows + i) * 2;
          row_best[o] = kPending;
          row_d[o] = 0.0f;
        } else {
          col_best[(size_t)p * max_cols + i] = kPending;
        }
      }
      return;
    } else if (nc == 1 && !knn) {
      // the nearest neighbour is proved without evaluating any distance; DMatch.distance is filled in
      // by k_tc_fill_dist only for the matches that survive (marker: negative distance)
      x0 = jx[0];
      b0 = -1.0f;
    } else {
      for (int c = half; c < nc + half; c += 2) {
        const int cc = c < nc ? c : nc - 1;
        const float d = exact_dist_half(arow, B + (size_t)jx[cc] * kDim, l16);
        if (c < nc) top2_push(d, jx[cc], b0, x0, b1, x1);
      }
    }
    // merge the two halves
    const float c0 = __shfl_xor_sync(0xffffffffu, b0, 16), c1 = __shfl_xor_sync(0xffffffffu, b1, 16);
    const int y0 = __shfl_xor_sync(0xffffffffu, x0, 16), y1 = __shfl_xor_sync(0xffffffffu, x1, 16);
    if (y0 != x0 || c0 != b0) {
      top2_push(c0, y0, b0, x0, b1, x1);
      if (y1 != x1 || c1 != b1) top2_push(c1, y1, b0, x0, b1, x1);
    } else if (lex_less2(c1, y1, b1, x1)) {
      b1 = c1; x1 = y1;
    }
  }
  if (lane == 0) {
    if (!rev) {
      const size_t o = ((size_t)p * max_rows + i) * 2;
      row_best[o] = x0 == INT_MAX ? -1 : x0;
      row_best[o + 1] = x1 == INT_MAX ? -1 : x1;
      row_d[o] = b0;
      row_d[o + 1] = b1;
    } else {
      col_best[(size_t)p * max_cols + i] = x0 == INT_MAX ? -1 : x0;
    }
  }
}

__global__ void __launch_bounds__(256)
k_tc_rerank(const MatchProblem* __restrict__ probs, int P, int mode, float ratio, const float* __restrict__ nrm,
            const unsigned* __restrict__ opmax, const Short* __restrict__ shortl, int cap, int max_rows,
            int max_cols, int* __restrict__ row_best, float* __restrict__ row_d, int* __restrict__ col_best,
            int* __restrict__ fb_count, int* __restrict__ fb_list, unsigned long long* __restrict__ counters,
            float eps_rel, int unit, const int* __restrict__ rr_count, const int* __restrict__ rr_list) {
  chain_enter();
  const int dp = blockIdx.y, p = dp < P ? dp : dp - P;
  const bool rev = dp >= P;
  const MatchProblem pr = probs[p];
  const int n_rr = rr_count[dp];
  // only the rows k_tc_triage queued; a few blocks per directed problem stride over its list
  for (int li = blockIdx.x * 8 + (threadIdx.x >> 5); li < n_rr; li += gridDim.x * 8) {
    const Short sl = shortl[(size_t)dp * cap + li];
    rerank_row(pr, p, dp, rev, rr_list[(size_t)dp * cap + li], sl, mode, ratio, nrm, opmax, cap, max_rows, max_cols, row_best,
               row_d, col_best, fb_count, fb_list, counters, eps_rel, unit);
  }
}

// k_tc_fill_dist: exact DMatch.distance for the forward matches whose index was proved without a
// distance (row_d < 0) and that survive the cross-check.  Half a warp per query row.
__global__ void __launch_bounds__(256)
k_tc_fill_dist(const MatchProblem* __restrict__ probs, int mode, int max_rows, int max_cols,
               const int* __restrict__ row_best, float* __restrict__ row_d, const int* __restrict__ col_best) {
  chain_enter();
  const int p = blockIdx.y;
  const MatchProblem pr = probs[p];
  const int lane = threadIdx.x & 31, l16 = lane & 15, half = lane >> 4;
  const int i = blockIdx.x * 16 + (threadIdx.x >> 5) * 2 + half;
  if (blockIdx.x * 16 >= pr.N) return;
  bool need = false;
  int j = 0;
  const size_t o = ((size_t)p * max_rows + (i < pr.N ? i : 0)) * 2;
  if (i < pr.N && pr.M > 0) {
    j = row_best[o];
    need = j >= 0 && row_d[o] < 0.0f;  // a row still pending in the fallback (kPending) is filled by the fallback itself
    if (need && mode == SPVO_MATCH_NN_CROSSCHECK) {
      // a column pending in the (possibly concurrent) fallback may still turn out to be this row's mutual match
      const int cb = *reinterpret_cast<const volatile int*>(col_best + (size_t)p * max_cols + j);
      need = cb == i || cb == kPending;
    }
  }
  if (!__any_sync(0xffffffffu, need)) return;
  const int ii = need ? i : 0, jj = need ? j : 0;
  const float d = exact_dist_half(pr.q + (size_t)ii * kDim, pr.t + (size_t)jj * kDim, l16);
  if (need && l16 == 0) row_d[o] = d;
}

// ------------------------------------------------------------------------------------------------
// k_tc_fallback: one CTA per directed problem re-evaluates its queued rows against ALL columns,
// eight rows at a time so the train descriptors stream through once per group.
// Phase 1: fp32 dot products (error ~1e-5, three orders below the bf16 bound) give every column's
// g = |b|^2 - 2 a.b; each warp keeps the 3 best columns per row.  Phase 2: the columns within
// 2*eps32 of the row minimum (of the second minimum for kNN) are re-evaluated in OpenCV's exact
// order.  If that list could be incomplete (>= 3 near-ties inside one warp's share) every column is
// scanned exactly (counters[2]).
// ------------------------------------------------------------------------------------------------
constexpr int kFbRows = 8;
constexpr int kFbSplit = 4;             // CTAs (one thread-block cluster) per directed problem: each scans a quarter of the columns
constexpr int kFbVW = 8 * kFbSplit;     // "virtual warps" of a cluster
constexpr int kFbCand = 3 * kFbVW / 32; // shortlisted columns per lane in the finishing step

__global__ void __launch_bounds__(256, 2)
k_tc_fallback(const MatchProblem* __restrict__ probs, int P, int mode, const float* __restrict__ nrm, int cap,
              int max_rows, int max_cols, const int* __restrict__ fb_count, const int* __restrict__ fb_list,
              int* __restrict__ row_best, float* __restrict__ row_d, int* __restrict__ col_best,
              unsigned long long* __restrict__ counters) {
  chain_enter();
  // The scan of a queued row group is latency-bound per CTA (a warp keeps two train rows in flight), so the columns
  // are split over a cluster of kFbSplit CTAs; their per-warp shortlists meet in the shared memory of cluster rank 0
  // (distributed shared memory), which finishes the rows.
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float s_v[kFbVW][kFbRows][3];
  __shared__ int s_j[kFbVW][kFbRows][3];
  const int crank = (int)cluster.block_rank();
  const int dp = blockIdx.x / kFbSplit, p = dp < P ? dp : dp - P;
  const int nfb = fb_count[dp];
  if (nfb == 0) return;  // uniform over the cluster
  const bool rev = dp >= P;
  const MatchProblem pr = probs[p];
  const int Nb = rev ? pr.N : pr.M;
  const float* A = rev ? pr.t : pr.q;
  const float* B = rev ? pr.q : pr.t;
  const int a_op = rev ? pr.b_op : pr.a_op, b_op = rev ? pr.a_op : pr.b_op;
  const float* nbv = nrm + (size_t)b_op * cap;
  const bool knn = (mode == SPVO_MATCH_KNN_RATIO) && !rev;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, l16 = lane & 15, half = lane >> 4;
  const int vw = crank * 8 + warp;  // this warp takes columns vw, vw + 32, vw + 64, ...
  const int myr = lane >> 2;        // after the butterfly, lanes 4r..4r+3 hold row r's dot product
  float* rv = cluster.map_shared_rank(&s_v[0][0][0], 0);
  int* rj = cluster.map_shared_rank(&s_j[0][0][0], 0);
  cluster.sync();  // rank 0 must be resident before anybody writes into its shared memory

  for (int r0 = 0; r0 < nfb; r0 += kFbRows) {
    const int nr = min(kFbRows, nfb - r0);
    // this lane's 8 elements (8*lane .. 8*lane+7) of each of the group's rows
    float a[kFbRows][8];
#pragma unroll
    for (int r = 0; r < kFbRows; ++r) {
      const int i = fb_list[(size_t)dp * cap + r0 + min(r, nr - 1)];
      const float4 x = __ldg(reinterpret_cast<const float4*>(A + (size_t)i * kDim) + 2 * lane);
      const float4 y = __ldg(reinterpret_cast<const float4*>(A + (size_t)i * kDim) + 2 * lane + 1);
      a[r][0] = x.x; a[r][1] = x.y; a[r][2] = x.z; a[r][3] = x.w;
      a[r][4] = y.x; a[r][5] = y.y; a[r][6] = y.z; a[r][7] = y.w;
    }
    // row-band mask: this lane finishes row myr of the group (query index for the forward direction)
    const int my_i = fb_list[(size_t)dp * cap + r0 + min(myr, nr - 1)];
    auto allowed = [&](int row, int col) { return rev ? band_allowed(pr, col, row) : band_allowed(pr, row, col); };
    float v[3] = {INFINITY, INFINITY, INFINITY};
    int ix[3] = {-1, -1, -1};
    constexpr int kU = 2;  // columns per step; the next step's loads are issued before this step's math
    float4 xs[kU], ys[kU], xn[kU], yn[kU];
    float nbs[kU], nbn[kU];
    auto load_cols = [&](int jb, float4* x, float4* y, float* nb) {
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int j = min(jb + kFbVW * u, Nb - 1);
        x[u] = __ldg(reinterpret_cast<const float4*>(B + (size_t)j * kDim) + 2 * lane);
        y[u] = __ldg(reinterpret_cast<const float4*>(B + (size_t)j * kDim) + 2 * lane + 1);
        nb[u] = __ldg(nbv + j);
      }
    };
    if (vw < Nb) load_cols(vw, xn, yn, nbn);
    for (int jb = vw; jb < Nb; jb += kFbVW * kU) {
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        xs[u] = xn[u];
        ys[u] = yn[u];
        nbs[u] = nbn[u];
      }
      if (jb + kFbVW * kU < Nb) load_cols(jb + kFbVW * kU, xn, yn, nbn);
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int j = jb + kFbVW * u;
        const float b[8] = {xs[u].x, xs[u].y, xs[u].z, xs[u].w, ys[u].x, ys[u].y, ys[u].z, ys[u].w};
        float pd[kFbRows];
#pragma unroll
        for (int r = 0; r < kFbRows; ++r) {
          float t = 0.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) t = __fmaf_rn(a[r][e], b[e], t);
          pd[r] = t;
        }
        // reduce 8 values over 32 lanes: halve the value set at xor 16 / 8 / 4, then finish at xor 2 / 1
        float q4[4], q2[2], q1;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const float keep = (lane & 16) ? pd[r + 4] : pd[r], send = (lane & 16) ? pd[r] : pd[r + 4];
          q4[r] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const float keep = (lane & 8) ? q4[r + 2] : q4[r], send = (lane & 8) ? q4[r] : q4[r + 2];
          q2[r] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        {
          const float keep = (lane & 4) ? q2[1] : q2[0], send = (lane & 4) ? q2[0] : q2[1];
          q1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
        q1 += __shfl_xor_sync(0xffffffffu, q1, 1);
        if (j < Nb && allowed(my_i, j)) {
          const float g = __fmaf_rn(-2.0f, q1, nbs[u]);
          if (g < v[2]) {
            if (g < v[1]) {
              v[2] = v[1]; ix[2] = ix[1];
              if (g < v[0]) { v[1] = v[0]; ix[1] = ix[0]; v[0] = g; ix[0] = j; } else { v[1] = g; ix[1] = j; }
            } else { v[2] = g; ix[2] = j; }
          }
        }
      }
    }
    if ((lane & 3) == 0) {  // into rank 0's shared memory
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        rv[(vw * kFbRows + myr) * 3 + k] = v[k];
        rj[(vw * kFbRows + myr) * 3 + k] = ix[k];
      }
    }
    cluster.sync();
    // rank 0, warp w finishes row w of the group: 3 shortlisted columns per virtual warp
    if (crank == 0 && warp < nr) {
      const int i = fb_list[(size_t)dp * cap + r0 + warp];
      const float* arow = A + (size_t)i * kDim;
      const float na = nrm[(size_t)a_op * cap + i];
      float cv[kFbCand];
      int cj[kFbCand];
      float m0 = INFINITY, m1 = INFINITY, bm = 0.f;
#pragma unroll
      for (int t = 0; t < kFbCand; ++t) {
        const int idx = lane + 32 * t;
        cv[t] = s_v[idx / 3][warp][idx % 3];
        cj[t] = s_j[idx / 3][warp][idx % 3];
        if (cj[t] >= 0) bm = fmaxf(bm, nbv[cj[t]]);
        const float hi = fmaxf(m0, cv[t]);
        m0 = fminf(m0, cv[t]);
        m1 = fminf(m1, hi);
      }
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        const float c0 = __shfl_xor_sync(0xffffffffu, m0, o), c1 = __shfl_xor_sync(0xffffffffu, m1, o);
        const float lo = fminf(m0, c0), hi = fmaxf(m0, c0);
        m1 = fminf(hi, fminf(m1, c1));
        m0 = lo;
        bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
      }
      const float eps32 = 4e-5f * (na + bm);  // >= 2 * (256 + 8) * 2^-24 |a||b| plus the final roundings
      const float thr = (knn ? m1 : m0) + 2.0f * eps32;
      bool cand[kFbCand], over = false;
#pragma unroll
      for (int t = 0; t < kFbCand; ++t) {
        cand[t] = cj[t] >= 0 && cv[t] <= thr;
        // a warp's third entry inside the threshold means that warp may have dropped a near-tie
        over |= cand[t] && ((lane + 32 * t) % 3) == 2 && Nb > 3 * kFbVW;
      }
      const bool overflow = __any_sync(0xffffffffu, over);
      float b0 = INFINITY, b1 = INFINITY;
      int x0 = INT_MAX, x1 = INT_MAX;
      if (overflow) {
        for (int j = half; j < Nb + half; j += 2) {
          const int jj = j < Nb ? j : Nb - 1;
          const float d = exact_dist_half(arow, B + (size_t)jj * kDim, l16);
          if (j < Nb && allowed(i, j)) top2_push(d, j, b0, x0, b1, x1);
        }
        if (lane == 0) atomicAdd(&counters[2], 1ull);
      } else {
#pragma unroll
        for (int t = 0; t < kFbCand; ++t) {
          unsigned pending = __ballot_sync(0xffffffffu, cand[t]);
          while (pending) {  // two candidate columns per step, one per half warp
            const int s0 = __ffs(pending) - 1;
            pending &= pending - 1;
            int s1 = s0;
            if (pending) {
              s1 = __ffs(pending) - 1;
              pending &= pending - 1;
            }
            const int ja = __shfl_sync(0xffffffffu, cj[t], s0), jb = __shfl_sync(0xffffffffu, cj[t], s1);
            const int j = half ? jb : ja;
            const float d = exact_dist_half(arow, B + (size_t)j * kDim, l16);
            if (half == 0 || s1 != s0) top2_push(d, j, b0, x0, b1, x1);
          }
        }
      }
      const float c0 = __shfl_xor_sync(0xffffffffu, b0, 16), c1 = __shfl_xor_sync(0xffffffffu, b1, 16);
      const int y0 = __shfl_xor_sync(0xffffffffu, x0, 16), y1 = __shfl_xor_sync(0xffffffffu, x1, 16);
      if (y0 != x0 || c0 != b0) {
        top2_push(c0, y0, b0, x0, b1, x1);
        if (y1 != x1 || c1 != b1) top2_push(c1, y1, b0, x0, b1, x1);
      } else if (lex_less2(c1, y1, b1, x1)) {
        b1 = c1; x1 = y1;
      }
      if (lane == 0) {
        if (!rev) {
          const size_t o = ((size_t)p * max_rows + i) * 2;
          row_best[o] = x0 == INT_MAX ? -1 : x0;
          row_best[o + 1] = x1 == INT_MAX ? -1 : x1;
          row_d[o] = b0;
          row_d[o + 1] = b1;
        } else {
          col_best[(size_t)p * max_cols + i] = x0 == INT_MAX ? -1 : x0;
        }
      }
    }
    cluster.sync();  // rank 0 has consumed the shortlists: the next group may overwrite them
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  }
  return fn;
}

struct TcWorkspace {
  __nv_bfloat16* xb = nullptr;
  float* nrm = nullptr;
  unsigned* opmax = nullptr;
  RowRec* row_rec = nullptr;   // [P][cap][2 halves][4 lanes]
  ColRec* col_rec = nullptr;   // [P][cap / 128 row blocks][cap]
  Short* shortl = nullptr;     // [ndir][cap] shortlists of the rows k_tc_triage queued
  size_t rec_rows = 0, rec_cols = 0;  // capacities of row_rec (rows) and col_rec (elements)
  int* fb_list = nullptr;
  int* rr_count = nullptr;  // rows queued by k_tc_triage for k_tc_rerank
  int* rr_list = nullptr;
  size_t rows = 0, ops = 0, top_rows = 0;
  int slot_cap = 0;  // rows per operand slot of the current layout
  bool fp16 = false; // operand format of the current contents (false: bf16)
  CUtensorMap tmap;
};

static cudaError_t tc_ensure(Handle* h, TcWorkspace* w, size_t ops, size_t cap, size_t ndir, size_t nprob) {
  const size_t rows = ops * cap;
  cudaError_t e;
  if (w->rows < rows || w->ops < ops) {
    if (w->xb) cudaFree(w->xb);
    if (w->nrm) cudaFree(w->nrm);
    if (w->opmax) cudaFree(w->opmax);
    w->xb = nullptr; w->nrm = nullptr; w->opmax = nullptr; w->rows = 0; w->ops = 0;
    if ((e = cudaMalloc((void**)&w->xb, rows * kDim * sizeof(__nv_bfloat16))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&w->nrm, rows * sizeof(float))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&w->opmax, ops * sizeof(unsigned))) != cudaSuccess) return e;
    PFN_encodeTiled enc = get_encode();
    if (!enc) return cudaErrorNotSupported;
    const cuuint64_t gdim[2] = {(cuuint64_t)kDim, (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {(cuuint64_t)kDim * sizeof(__nv_bfloat16)};
    const cuuint32_t box[2] = {(cuuint32_t)kKB, (cuuint32_t)kBM};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&w->tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w->xb, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
    w->rows = rows;
    w->ops = ops;
  }
  const size_t rec_rows = nprob * cap, rec_cols = nprob * (cap / kBM) * cap;
  if (w->rec_rows < rec_rows || w->rec_cols < rec_cols) {
    if (w->row_rec) cudaFree(w->row_rec);
    if (w->col_rec) cudaFree(w->col_rec);
    w->row_rec = nullptr; w->col_rec = nullptr; w->rec_rows = 0; w->rec_cols = 0;
    const size_t rr = rec_rows > w->rec_rows ? rec_rows : w->rec_rows, rc = rec_cols > w->rec_cols ? rec_cols : w->rec_cols;
    if ((e = cudaMalloc((void**)&w->row_rec, rr * 8 * sizeof(RowRec))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&w->col_rec, rc * sizeof(ColRec))) != cudaSuccess) return e;
    w->rec_rows = rr; w->rec_cols = rc;
  }
  const size_t top_rows = ndir * cap;
  if (w->top_rows < top_rows) {
    if (w->shortl) cudaFree(w->shortl);
    if (w->fb_list) cudaFree(w->fb_list);
    if (w->rr_count) cudaFree(w->rr_count);
    if (w->rr_list) cudaFree(w->rr_list);
    w->shortl = nullptr; w->fb_list = nullptr; w->rr_count = nullptr; w->rr_list = nullptr;
    w->top_rows = 0;
    if ((e = cudaMalloc((void**)&w->rr_count, 2 * top_rows * sizeof(int))) != cudaSuccess) return e;  // + the fallback counters
    if ((e = cudaMalloc((void**)&w->rr_list, top_rows * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&w->fb_list, top_rows * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&w->shortl, top_rows * sizeof(Short))) != cudaSuccess) return e;
    w->top_rows = top_rows;
  }
  return cudaSuccess;
}

void tc_workspace_free(Handle* h) {
  TcWorkspace* w = reinterpret_cast<TcWorkspace*>(h->tc_ws);
  if (!w) return;
  void* ptrs[] = {w->xb, w->nrm, w->opmax, w->row_rec, w->col_rec, w->shortl, w->fb_list, w->rr_count, w->rr_list};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete w;
  h->tc_ws = nullptr;
}

cudaError_t launch_finalize_only(Handle* h, const MatchProblem* probs, int P, int mr, int mc,
                                 const spvo_match_cfg& cfg, spvo_dmatch* out, int* n_matches, int* q2t, int out_stride);
cudaError_t ensure_select_buffers(Handle* h, int P, int mr, int mc);

static cudaError_t tc_get(Handle* h, TcWorkspace** w, size_t ops, size_t cap, size_t ndir, size_t nprob) {
  if (!h->tc_ws) h->tc_ws = new TcWorkspace();
  *w = reinterpret_cast<TcWorkspace*>(h->tc_ws);
  return tc_ensure(h, *w, ops, cap, ndir, nprob);
}

cudaError_t tc_prepare_slots(Handle* h, int slots, int max_rows, int ndir, TcSink* sink) {
  TcWorkspace* w;
  const int cap = (max_rows + kCapAlign - 1) / kCapAlign * kCapAlign;
  cudaError_t e = tc_get(h, &w, (size_t)slots, (size_t)cap, (size_t)ndir, (size_t)(ndir + 1) / 2);
  if (e != cudaSuccess) return e;
  w->slot_cap = cap;
  w->fp16 = true;  // decode's descriptors are unit-norm (NN:428): fp16 is range-safe and 8x finer than bf16
  sink->fp16 = 1;
  sink->xb = w->xb;
  sink->nrm = w->nrm;
  sink->opmax = w->opmax;
  sink->cap = cap;
  return cudaSuccess;
}

int tc_copy_slot_segments(Handle* h, int dst, int src, CopySeg* segs) {
  TcWorkspace* w = reinterpret_cast<TcWorkspace*>(h->tc_ws);
  if (!w) return 0;
  const size_t cap = w->slot_cap;
  segs[0] = {w->xb + (size_t)src * cap * kDim, w->xb + (size_t)dst * cap * kDim, cap * kDim * sizeof(__nv_bfloat16)};
  segs[1] = {w->nrm + (size_t)src * cap, w->nrm + (size_t)dst * cap, cap * sizeof(float)};
  segs[2] = {w->opmax + src, w->opmax + dst, sizeof(unsigned)};
  return 3;
}

cudaError_t tc_prep_problem_operands(Handle* h, const MatchProblem* prob) {
  TcWorkspace* w = reinterpret_cast<TcWorkspace*>(h->tc_ws);
  if (!w || !w->slot_cap) return cudaErrorInvalidValue;
  LaunchScope ls(h, KID_TC_PREP);
  k_tc_prep<<<dim3((w->slot_cap + 31) / 32, 2), 256, 0, h->stream>>>(prob, w->xb, w->nrm, w->opmax, w->slot_cap, w->fp16 ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launch_match_tc(Handle* h, const MatchProblem* probs, int P, int max_rows, int max_cols,
                            const spvo_match_cfg& cfg, spvo_dmatch* out, int* n_matches, int* q2t, int out_stride,
                            bool operands_ready) {
  cudaStream_t st = h->stream;
  cudaError_t e;
  if (P == 0) return cudaSuccess;
  const int mr = max_rows > 0 ? max_rows : 1, mc = max_cols > 0 ? max_cols : 1;
  if ((e = ensure_select_buffers(h, P, mr, mc)) != cudaSuccess) return e;
  if (max_rows > 0 && max_cols > 0) {
    const int mx = max_rows > max_cols ? max_rows : max_cols;
    int cap = (mx + kCapAlign - 1) / kCapAlign * kCapAlign;
    const bool cross = cfg.mode == SPVO_MATCH_NN_CROSSCHECK;
    const int ndir = cross ? 2 * P : P;
    TcWorkspace* w;
    if (operands_ready) {
      // the stereo pipeline already reserved the slots and k_desc_normalize filled them
      w = reinterpret_cast<TcWorkspace*>(h->tc_ws);
      if (!w || w->top_rows < (size_t)ndir * w->slot_cap || w->rec_rows < (size_t)P * w->slot_cap ||
          w->rec_cols < (size_t)P * (w->slot_cap / kBM) * w->slot_cap)
        return cudaErrorInvalidValue;
      cap = w->slot_cap;
    } else {
      // operand slots are the problems' a_op / b_op: 2p, 2p+1 for the generic entry points, image indices (and the
      // carry slot max_batch) for a stereo batch whose decode could not fill them
      const size_t ops = (size_t)(2 * P > h->max_batch + 1 ? 2 * P : h->max_batch + 1);
      if ((e = tc_get(h, &w, ops, (size_t)cap, (size_t)ndir, (size_t)P)) != cudaSuccess) return e;
      w->slot_cap = cap;
      w->fp16 = false;  // arbitrary CV_32F descriptors: bf16 keeps the fp32 exponent range
      h->carry_tc_valid = false;  // the slots (and possibly the buffers) of a previous stereo batch are overwritten
      if ((e = cudaMemsetAsync(w->opmax, 0, ops * sizeof(unsigned), st)) != cudaSuccess) return e;
      LaunchScope ls(h, KID_TC_PREP);
      k_tc_prep<<<dim3((cap + 31) / 32, 2 * P), 256, 0, st>>>(probs, w->xb, w->nrm, w->opmax, cap, 0);
    }
    // both worklist counters are cleared by k_tc_gemm: fb_count sits right behind the ndir rerank counters of this call
    int* const fb_count = w->rr_count + ndir;
    const float eps_rel = w->fp16 ? kEpsRelFp16 : kEpsRelBf16;
    // unit-norm operands (written by decode, fp16) use constant key scaling; arbitrary CV_32F inputs scale by the
    // operands' largest norms (tc_scale)
    const int unit = (operands_ready && w->fp16) ? 1 : 0;
    const size_t smem = 1024 + (size_t)kNumKB * kTileBytes + (size_t)kBSlots * kBSlotBytes + sizeof(TcShared);
    {
      const bool mask = (cfg.flags & SPVO_MATCH_FLAG_ROW_BAND) != 0;  // some problem carries a row-band mask
      auto kern = mask ? (cross ? (unit ? k_tc_gemm<true, true, true> : k_tc_gemm<false, true, true>)
                                : (unit ? k_tc_gemm<true, false, true> : k_tc_gemm<false, false, true>))
                       : (cross ? (unit ? k_tc_gemm<true, true, false> : k_tc_gemm<false, true, false>)
                                : (unit ? k_tc_gemm<true, false, false> : k_tc_gemm<false, false, false>));
      if ((e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
      LaunchScope ls(h, KID_TC_GEMM);
      const int n_items = (cap / kBM) * P;  // ONE Gram matrix per problem: both directions come from its epilogue
      const int grid = n_items < h->sm_count ? n_items : h->sm_count;  // persistent: one CTA per SM
      if ((e = launch_chained(h->chain_launches, kern, dim3(grid), dim3(kTcThreads), smem, st, 1, w->tmap, probs, w->nrm, w->opmax, w->row_rec,
                              w->col_rec, cap, make_idesc(w->fp16), n_items, 256u, w->rr_count, 2 * ndir)) != cudaSuccess)
        return e;
    }
    // NN modes: the triage blocks re-rank their own unresolved rows (one launch fewer); kNN: every row is evaluated,
    // which wants the wider k_tc_rerank grid
    const bool fuse_rerank = cfg.mode != SPVO_MATCH_KNN_RATIO;
    {
      LaunchScope ls(h, KID_TC_TRIAGE);
      if ((e = launch_chained(h->chain_launches, k_tc_triage, dim3((cap + 255) / 256, ndir), dim3(256), 0, st, 1, probs, P, (int)cfg.mode,
                              w->nrm, w->opmax, w->row_rec, w->col_rec, cap, mr, mc, h->row_best, h->row_d, h->col_best,
                              w->rr_count, w->rr_list, w->shortl, eps_rel, unit, fuse_rerank ? 1 : 0, cfg.ratio, fb_count,
                              w->fb_list, h->counters)) != cudaSuccess)
        return e;
    }
    if (!fuse_rerank) {
      LaunchScope ls(h, KID_TC_RERANK);
      if ((e = launch_chained(h->chain_launches, k_tc_rerank, dim3(cfg.mode == SPVO_MATCH_KNN_RATIO ? (cap / 32 > 4 ? cap / 32 : 4) : 4, ndir),
                              dim3(256), 0, st, 1, probs, P, (int)cfg.mode, cfg.ratio, w->nrm, w->opmax, w->shortl, cap, mr,
                              mc, h->row_best, h->row_d, h->col_best, fb_count, w->fb_list, h->counters, eps_rel, unit,
                              w->rr_count, w->rr_list)) != cudaSuccess)
        return e;
    }
    // k_tc_fallback is a few hundred long-lived, latency-bound blocks; k_tc_fill_dist streams descriptor rows on
    // every SM.  For the non-kNN modes they run CONCURRENTLY (fork to an auxiliary stream, join before finalize):
    // pending rows / columns are marked by k_tc_rerank, so neither needs the other's result.  Profiling runs keep
    // them in series on the main stream (per-kernel events).
    static const bool overlap_env = [] {
      const char* e = getenv("SPVO_TAIL_OVERLAP");
      return !(e && e[0] == '0');
    }();
    const bool overlap = overlap_env && !h->profiling && cfg.mode != SPVO_MATCH_KNN_RATIO;
    cudaStream_t fb_st = st;
    if (overlap) {
      if ((e = ensure_aux_streams(h)) != cudaSuccess) return e;
      fb_st = h->aux_stream[0];
      if ((e = cudaEventRecord(h->aux_fork, st)) != cudaSuccess) return e;
      if ((e = cudaStreamWaitEvent(fb_st, h->aux_fork, 0)) != cudaSuccess) return e;
    }
    {
      LaunchScope ls(h, KID_TC_FALLBACK);
      const int* fbc = fb_count;
      const int* fbl = w->fb_list;
      if ((e = launch_chained(h->chain_launches, k_tc_fallback, dim3(ndir * kFbSplit), dim3(256), 0, fb_st, kFbSplit, probs, P, (int)cfg.mode,
                              (const float*)w->nrm, cap, mr, mc, fbc, fbl, h->row_best, h->row_d, h->col_best,
                              h->counters)) != cudaSuccess)
        return e;
    }
    if (overlap && (e = cudaEventRecord(h->aux_done[0], fb_st)) != cudaSuccess) return e;
    if (cfg.mode != SPVO_MATCH_KNN_RATIO) {
      LaunchScope ls(h, KID_TC_FILL);
      if ((e = launch_chained(h->chain_launches, k_tc_fill_dist, dim3((max_rows + 15) / 16, P), dim3(256), 0, st, 1, probs, (int)cfg.mode, mr,
                              mc, h->row_best, h->row_d, h->col_best)) != cudaSuccess)
        return e;
    }
    if (overlap && (e = cudaStreamWaitEvent(st, h->aux_done[0], 0)) != cudaSuccess) return e;
  }
  return launch_finalize_only(h, probs, P, mr, mc, cfg, out, n_matches, q2t, out_stride);
}

}  // namespace spvo
