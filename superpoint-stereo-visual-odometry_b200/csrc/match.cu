// match.cu -- exact fp32 brute-force L2 matcher (CUDA cores, sm_100a) and the match epilogues.
//
// Replaces FeatureFrontEnd::matchDescriptors (reference: src/odml_visual_odometry/src/
// feature_detection_base.cpp:434-500, "BASE") = cv::BFMatcher(NORM_L2)::match / knnMatch(k=2) +
// 0.8 ratio test + maps_of_indices.  cv::BFMatcher itself is OpenCV (third party, pinned 4.5.4 at
// CMakeLists.txt:14): batchDistance -> batchDistL2_32f -> hal::normL2Sqr_ + sqrt.  Its fp32
// operation order is reproduced exactly (see dist_tile): for element j = 16*blk + 4*k + l,
// s[k][l] += (a_j - b_j)*(a_j - b_j) (unfused); v[l] = ((s0+s1)+s2)+s3; d2 = (v0+v2)+(v1+v3);
// D = sqrt(d2).  Ties resolve to the lowest index; cross-check is the mutual first-argmin.
//
// This kernel set is the correctness anchor (SPVO_MATCHER_EXACT_FP32) and provides the exact
// re-rank used by the tensor-core matcher.
#include <float.h>
#include <limits.h>

#include "common.cuh"

namespace spvo {

constexpr int kD = SPVO_DESC_DIM;  // 256
constexpr int kTile = 32;
constexpr int kHalf = 128;
constexpr int kPitch = kHalf + 4;  // floats; 4*row mod 32 banks -> conflict-free LDS.128 over 8 rows

__device__ __forceinline__ float4 sub4(float4 a, float4 b) {
  return make_float4(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z), __fsub_rn(a.w, b.w));
}
__device__ __forceinline__ void acc4(float4& s, float4 t) {
  s.x = __fadd_rn(s.x, __fmul_rn(t.x, t.x));
  s.y = __fadd_rn(s.y, __fmul_rn(t.y, t.y));
  s.z = __fadd_rn(s.z, __fmul_rn(t.z, t.z));
  s.w = __fadd_rn(s.w, __fmul_rn(t.w, t.w));
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}
// OpenCV's final reduction of the four accumulators (v_reduce_sum of ((s0+s1)+s2)+s3).
__device__ __forceinline__ float finish_cv(const float4* s) {
  float4 v = add4(add4(add4(s[0], s[1]), s[2]), s[3]);
  return __fsqrt_rn(__fadd_rn(__fadd_rn(v.x, v.z), __fadd_rn(v.y, v.w)));
}

// D[p][i][j] for a 32x32 tile of (query, train) pairs; each thread owns pairs
// {ty, ty+16} x {tx, tx+16} with 16 running sums per pair, exactly OpenCV's.
__global__ void __launch_bounds__(256)
k_dist_exact(const MatchProblem* __restrict__ probs, float* __restrict__ dist, int max_rows, int max_cols) {
  __shared__ __align__(16) float q_s[kTile][kPitch];
  __shared__ __align__(16) float t_s[kTile][kPitch];
  const MatchProblem pr = probs[blockIdx.z];
  const int i0 = blockIdx.y * kTile, j0 = blockIdx.x * kTile;
  if (i0 >= pr.N || j0 >= pr.M) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  float4 s[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int k = 0; k < 4; ++k) s[a][k] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int half = 0; half < 2; ++half) {
    // cooperative load: 32 rows x 128 floats per operand = 1024 float4 each, 256 threads x 4
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int f = tid + it * 256;
      const int r = f >> 5, c4 = f & 31;
      float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), tv = qv;
      if (i0 + r < pr.N) qv = __ldg(reinterpret_cast<const float4*>(pr.q + (size_t)(i0 + r) * kD + half * kHalf) + c4);
      if (j0 + r < pr.M) tv = __ldg(reinterpret_cast<const float4*>(pr.t + (size_t)(j0 + r) * kD + half * kHalf) + c4);
      *reinterpret_cast<float4*>(&q_s[r][c4 * 4]) = qv;
      *reinterpret_cast<float4*>(&t_s[r][c4 * 4]) = tv;
    }
    __syncthreads();
#pragma unroll 2
    for (int blk = 0; blk < kHalf / 16; ++blk) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = blk * 16 + k * 4;
        const float4 qa = *reinterpret_cast<const float4*>(&q_s[ty][c]);
        const float4 qb = *reinterpret_cast<const float4*>(&q_s[ty + 16][c]);
        const float4 ta = *reinterpret_cast<const float4*>(&t_s[tx][c]);
        const float4 tb = *reinterpret_cast<const float4*>(&t_s[tx + 16][c]);
        acc4(s[0][k], sub4(qa, ta));
        acc4(s[1][k], sub4(qa, tb));
        acc4(s[2][k], sub4(qb, ta));
        acc4(s[3][k], sub4(qb, tb));
      }
    }
    __syncthreads();
  }
  float* D = dist + (size_t)blockIdx.z * max_rows * max_cols;
  const int ia = i0 + ty, ib = i0 + ty + 16, ja = j0 + tx, jb = j0 + tx + 16;
  if (ia < pr.N && ja < pr.M) D[(size_t)ia * max_cols + ja] = finish_cv(s[0]);
  if (ia < pr.N && jb < pr.M) D[(size_t)ia * max_cols + jb] = finish_cv(s[1]);
  if (ib < pr.N && ja < pr.M) D[(size_t)ib * max_cols + ja] = finish_cv(s[2]);
  if (ib < pr.N && jb < pr.M) D[(size_t)ib * max_cols + jb] = finish_cv(s[3]);
}

__device__ __forceinline__ bool lex_less(float d1, int i1, float d2, int i2) {
  return d1 < d2 || (d1 == d2 && i1 < i2);
}

// Per query: the two smallest distances in (distance, index) order == cv::batchDistance's
// strict-'<' insertion for K = 1 / K = 2.  One warp per query row.
__global__ void __launch_bounds__(256)
k_row_select(const MatchProblem* __restrict__ probs, const float* __restrict__ dist, int max_rows, int max_cols,
             int* __restrict__ row_best, float* __restrict__ row_d) {
  const MatchProblem pr = probs[blockIdx.y];
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (i >= pr.N) return;
  const float* Drow = dist + ((size_t)blockIdx.y * max_rows + i) * max_cols;
  float b0 = INFINITY, b1 = INFINITY;
  int x0 = INT_MAX, x1 = INT_MAX;
  for (int j = lane; j < pr.M; j += 32) {
    if (!band_allowed(pr, i, j)) continue;
    const float d = Drow[j];
    if (lex_less(d, j, b0, x0)) {
      b1 = b0; x1 = x0; b0 = d; x0 = j;
    } else if (lex_less(d, j, b1, x1)) {
      b1 = d; x1 = j;
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const float c0 = __shfl_xor_sync(0xffffffffu, b0, o), c1 = __shfl_xor_sync(0xffffffffu, b1, o);
    const int y0 = __shfl_xor_sync(0xffffffffu, x0, o), y1 = __shfl_xor_sync(0xffffffffu, x1, o);
    if (lex_less(c0, y0, b0, x0)) {
      // other's best wins; second = min(mine best, other's second)
      if (lex_less(b0, x0, c1, y1)) { b1 = b0; x1 = x0; } else { b1 = c1; x1 = y1; }
      b0 = c0; x0 = y0;
    } else {
      if (lex_less(c0, y0, b1, x1)) { b1 = c0; x1 = y0; }
    }
  }
  if (lane == 0) {
    const size_t o = ((size_t)blockIdx.y * max_rows + i) * 2;
    row_best[o] = x0 == INT_MAX ? -1 : x0;
    row_best[o + 1] = x1 == INT_MAX ? -1 : x1;
    row_d[o] = b0;
    row_d[o + 1] = b1;
  }
}

// Per train column: first-index argmin over the queries (the reverse pass of cv::batchDistance's
// crosscheck; D(a,b) is bitwise symmetric so the same matrix serves).  Block = 32 columns x 8 strides.
__global__ void __launch_bounds__(256)
k_col_select(const MatchProblem* __restrict__ probs, const float* __restrict__ dist, int max_rows, int max_cols,
             int* __restrict__ col_best) {
  __shared__ float sd[8][32];
  __shared__ int si[8][32];
  const MatchProblem pr = probs[blockIdx.y];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + tx;
  if (blockIdx.x * 32 >= pr.M) return;
  const float* D = dist + (size_t)blockIdx.y * max_rows * max_cols;
  float b0 = INFINITY;
  int x0 = INT_MAX;
  if (j < pr.M)
    for (int i = ty; i < pr.N; i += 8) {
      if (!band_allowed(pr, i, j)) continue;
      const float d = D[(size_t)i * max_cols + j];
      if (lex_less(d, i, b0, x0)) { b0 = d; x0 = i; }
    }
  sd[ty][tx] = b0;
  si[ty][tx] = x0;
  __syncthreads();
  if (ty == 0 && j < pr.M) {
#pragma unroll
    for (int r = 1; r < 8; ++r)
      if (lex_less(sd[r][tx], si[r][tx], b0, x0)) { b0 = sd[r][tx]; x0 = si[r][tx]; }
    col_best[(size_t)blockIdx.y * max_cols + j] = x0 == INT_MAX ? -1 : x0;
  }
}

// Final selection + ordered compaction into the DMatch list (ascending queryIdx) and the
// query->train map (BASE:483-491).  One block per problem.
__global__ void __launch_bounds__(1024)
k_finalize_matches(const MatchProblem* __restrict__ probs, const int* __restrict__ row_best,
                   const float* __restrict__ row_d, const int* __restrict__ col_best, int max_rows, int max_cols,
                   int mode, float ratio, spvo_dmatch* __restrict__ out, int* __restrict__ n_matches,
                   int* __restrict__ q2t, int out_stride, const FilterArgs flt) {
  chain_enter();
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const MatchProblem pr = probs[blockIdx.x];
  const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  const bool degenerate = pr.M == 0 || (mode == SPVO_MATCH_KNN_RATIO && pr.M < 2);
  const bool filter = flt.keep && p < flt.nprob;  // S1 (BASE:169-172) on this problem's matches
  for (int start = 0; start < pr.N; start += 1024) {
    const int i = start + tid;
    bool keep = false;
    int tr = -1;
    float d0 = 0.f;
    if (i < pr.N && !degenerate) {
      const size_t o = ((size_t)p * max_rows + i) * 2;
      tr = row_best[o];
      d0 = row_d[o];
      keep = tr >= 0;
      if (keep && mode == SPVO_MATCH_NN_CROSSCHECK) keep = col_best[(size_t)p * max_cols + tr] == i;
      // (a masked row may have a single allowed train row although M >= 2: no second neighbour, no match)
      if (keep && mode == SPVO_MATCH_KNN_RATIO) keep = row_best[o + 1] >= 0 && d0 < __fmul_rn(ratio, row_d[o + 1]);
    }
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (keep) {
      const int slot = off + __popc(m & ((1u << lane) - 1u));
      spvo_dmatch dm;
      dm.queryIdx = i; dm.trainIdx = tr; dm.imgIdx = 0; dm.distance = d0;
      out[(size_t)p * out_stride + slot] = dm;
      if (filter) {
        const spvo_keypoint a = flt.kpts[(size_t)(2 * p) * flt.slot_stride + i];
        const spvo_keypoint c = flt.kpts[(size_t)(2 * p + 1) * flt.slot_stride + tr];
        const bool drop = fabsf(__fsub_rn(a.y, c.y)) > flt.stereo_threshold || fabsf(__fsub_rn(a.x, c.x)) < flt.min_disparity;
        flt.keep[(size_t)p * out_stride + slot] = drop ? 0 : 1;
      }
    }
    if (q2t && i < pr.N) q2t[(size_t)p * out_stride + i] = keep ? tr : -1;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += s_warp[w];
      s_base += tot;
    }
    __syncthreads();
  }
  if (tid == 0) n_matches[p] = s_base;
  if (filter)
    for (int m = s_base + tid; m < out_stride; m += 1024) flt.keep[(size_t)p * out_stride + m] = 0;
}

__global__ void k_setup_problems(MatchProblem* probs, const float* desc_base, const int* n_rows,
                                 int slot_stride_rows, const int* q_slot, const int* t_slot, int P) {
  chain_enter();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  const int qs = q_slot[p], ts = t_slot[p];
  MatchProblem pr;
  pr.q = desc_base + (size_t)qs * slot_stride_rows * kD;
  pr.t = desc_base + (size_t)ts * slot_stride_rows * kD;
  pr.N = n_rows[qs];
  pr.M = n_rows[ts];
  pr.a_op = 2 * p;
  pr.b_op = 2 * p + 1;
  pr.qy = pr.ty = nullptr;
  pr.ystride = 0;
  pr.band = -1.0f;
  probs[p] = pr;
}

// Stereo stream problems: p < F stereo (left_f vs right_f); p >= F temporal (left_f vs left_{f-1},
// or the carried last-left of the previous batch for f = 0; carry_n = 0 means "no previous frame").
__global__ void k_setup_stereo_problems(const StereoSetup a) {
  chain_enter();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < 2 * a.F) setup_stereo_problem(a, p);
}

__global__ void k_set_problem(MatchProblem* probs, const float* q, int N, const float* t, int M,
                              const spvo_keypoint* q_kpts, const spvo_keypoint* t_kpts, float band) {
  MatchProblem pr;
  pr.q = q; pr.t = t; pr.N = N; pr.M = M;
  pr.a_op = 0; pr.b_op = 1;
  const bool masked = q_kpts && t_kpts && band >= 0.0f;
  pr.qy = masked ? &q_kpts[0].y : nullptr;
  pr.ty = masked ? &t_kpts[0].y : nullptr;
  pr.ystride = (int)(sizeof(spvo_keypoint) / sizeof(float));
  pr.band = band;
  probs[0] = pr;
}

// S1 (BASE:169-172): keep = !(|y_l - y_r| > stereo_threshold || |x_l - x_r| < min_disparity).
__global__ void k_stereo_filter(const spvo_keypoint* __restrict__ kpts_base, int slot_stride_rows,
                                const int* __restrict__ q_slot, const int* __restrict__ t_slot, int max_rows,
                                const spvo_dmatch* __restrict__ matches, const int* __restrict__ n_matches,
                                float stereo_threshold, float min_disparity, uint8_t* __restrict__ keep) {
  const int p = blockIdx.y;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= max_rows) return;
  uint8_t k = 0;
  if (m < n_matches[p]) {
    const spvo_dmatch dm = matches[(size_t)p * max_rows + m];
    const int qs = q_slot ? q_slot[p] : 2 * p, ts = t_slot ? t_slot[p] : 2 * p + 1;
    const spvo_keypoint a = kpts_base[(size_t)qs * slot_stride_rows + dm.queryIdx];
    const spvo_keypoint c = kpts_base[(size_t)ts * slot_stride_rows + dm.trainIdx];
    const bool drop = fabsf(__fsub_rn(a.y, c.y)) > stereo_threshold || fabsf(__fsub_rn(a.x, c.x)) < min_disparity;
    k = drop ? 0 : 1;
  }
  keep[(size_t)p * max_rows + m] = k;
}

constexpr int kCopyBlocks = 8;  // copy blocks per carry segment
// Post-match consistency (BASE:156-207): the quadruples solveStereoOdometry triangulates.  Block = frame.
//   stereo matches of frame f : matches row f        temporal map : q2t row F+f
//   previous frame's L<->R map: q2t row f-1, or the carried map of the previous batch for f = 0
__global__ void __launch_bounds__(1024)
k_consistency(int F, int K, const spvo_dmatch* __restrict__ matches, const int* __restrict__ n_matches,
              const int* __restrict__ q2t, const uint8_t* __restrict__ keep, const int* __restrict__ carry_map,
              spvo_quad* __restrict__ quads, int* __restrict__ n_quads, const CopyList cl) {
  chain_enter();
  if ((int)blockIdx.x >= F) {
    // copy blocks: the carry for the next batch (last left image's descriptors / keypoints / count / matcher operand,
    // L<->R map) -- kCopyBlocks blocks per segment, 16-byte accesses where the segment allows
    const int cb = blockIdx.x - F, segi = cb / kCopyBlocks, blk = cb % kCopyBlocks;
    const CopySeg g = cl.seg[segi];
    const size_t tid = (size_t)blk * blockDim.x + threadIdx.x, nth = (size_t)kCopyBlocks * blockDim.x;
    if (((reinterpret_cast<uintptr_t>(g.src) | reinterpret_cast<uintptr_t>(g.dst) | g.bytes) & 15) == 0) {
      const uint4* src = static_cast<const uint4*>(g.src);
      uint4* dst = static_cast<uint4*>(g.dst);
      for (size_t i = tid; i < g.bytes / 16; i += nth) dst[i] = src[i];
    } else {
      const uint32_t* src = static_cast<const uint32_t*>(g.src);
      uint32_t* dst = static_cast<uint32_t*>(g.dst);
      for (size_t i = tid; i < g.bytes / 4; i += nth) dst[i] = src[i];
    }
    return;
  }
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = n_matches[f];
  const int* map_t = q2t + (size_t)(F + f) * K;
  const int* map_prev = f > 0 ? q2t + (size_t)(f - 1) * K : carry_map;
  if (tid == 0) s_base = 0;
  __syncthreads();
  for (int start = 0; start < n; start += 1024) {
    const int m = start + tid;
    bool ok = false;
    spvo_quad qd;
    qd.curr_left = qd.curr_right = qd.prev_left = qd.prev_right = -1;
    if (m < n) {
      const spvo_dmatch dm = matches[(size_t)f * K + m];
      const int ipl = map_t[dm.queryIdx];                       // BASE:160
      if (ipl >= 0 && keep[(size_t)f * K + m]) {                // BASE:169-172
        const int ipr = map_prev[ipl];                          // BASE:181
        if (ipr >= 0) {
          ok = true;
          qd.curr_left = dm.queryIdx; qd.curr_right = dm.trainIdx; qd.prev_left = ipl; qd.prev_right = ipr;
        }
      }
    }
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < warp; ++w) off += s_warp[w];
    if (ok) quads[(size_t)f * K + off + __popc(bal & ((1u << lane) - 1u))] = qd;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += s_warp[w];
      s_base += tot;
    }
    __syncthreads();
  }
  if (tid == 0) n_quads[f] = s_base;
}

// ------------------------------------------------------------------------------------------------
// host-side launchers
// ------------------------------------------------------------------------------------------------
static cudaError_t ensure(void** ptr, size_t* have, size_t want, size_t elem) {
  if (*have >= want && *ptr) return cudaSuccess;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr;
  *have = 0;
  cudaError_t e = cudaMalloc(ptr, want * elem);
  if (e == cudaSuccess) *have = want;
  return e;
}

cudaError_t ensure_select_buffers(Handle* h, int P, int mr, int mc) {
  cudaError_t e;
  if (h->sel_rows < (size_t)P * mr) {
    if (h->row_best) cudaFree(h->row_best);
    if (h->row_d) cudaFree(h->row_d);
    h->row_best = nullptr; h->row_d = nullptr; h->sel_rows = 0;
    if ((e = cudaMalloc(&h->row_best, (size_t)P * mr * 2 * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&h->row_d, (size_t)P * mr * 2 * sizeof(float))) != cudaSuccess) return e;
    h->sel_rows = (size_t)P * mr;
  }
  return ensure((void**)&h->col_best, &h->sel_cols, (size_t)P * mc, sizeof(int));
}

cudaError_t launch_finalize_only(Handle* h, const MatchProblem* probs, int P, int mr, int mc,
                                 const spvo_match_cfg& cfg, spvo_dmatch* out, int* n_matches, int* q2t,
                                 int out_stride) {
  {
    LaunchScope ls(h, KID_FINALIZE);
    const cudaError_t e = launch_chained(h->chain_launches, k_finalize_matches, dim3(P), dim3(1024), 0, h->stream, 1, probs, h->row_best,
                                         h->row_d, h->col_best, mr, mc, (int)cfg.mode, cfg.ratio, out, n_matches, q2t,
                                         out_stride, h->fin_filter);
    h->fin_filter = FilterArgs();  // one-shot
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

cudaError_t launch_match_exact(Handle* h, const MatchProblem* probs, int P, int max_rows, int max_cols,
                               const spvo_match_cfg& cfg, spvo_dmatch* out, int* n_matches, int* q2t,
                               int out_stride) {
  cudaStream_t st = h->stream;
  cudaError_t e;
  if (P == 0) return cudaSuccess;
  const int mr = max_rows > 0 ? max_rows : 1, mc = max_cols > 0 ? max_cols : 1;
  if ((e = ensure((void**)&h->dist, &h->dist_elems, (size_t)P * mr * mc, sizeof(float))) != cudaSuccess) return e;
  if ((e = ensure_select_buffers(h, P, mr, mc)) != cudaSuccess) return e;
  if (max_rows > 0 && max_cols > 0) {
    dim3 g((max_cols + kTile - 1) / kTile, (max_rows + kTile - 1) / kTile, P);
    {
      LaunchScope ls(h, KID_DIST_EXACT);
      k_dist_exact<<<g, 256, 0, st>>>(probs, h->dist, mr, mc);
    }
    dim3 gr((max_rows + 7) / 8, P);
    {
      LaunchScope ls(h, KID_ROW_SELECT);
      k_row_select<<<gr, 256, 0, st>>>(probs, h->dist, mr, mc, h->row_best, h->row_d);
    }
    if (cfg.mode == SPVO_MATCH_NN_CROSSCHECK) {
      dim3 gc((max_cols + 31) / 32, P);
      LaunchScope ls(h, KID_COL_SELECT);
      k_col_select<<<gc, 256, 0, st>>>(probs, h->dist, mr, mc, h->col_best);
    }
  }
  return launch_finalize_only(h, probs, P, mr, mc, cfg, out, n_matches, q2t, out_stride);
}

cudaError_t launch_setup_problems(Handle* h, MatchProblem* probs, const float* desc_base, const int* n_rows,
                                  int slot_stride_rows, const int* q_slot, const int* t_slot, int P) {
  if (P == 0) return cudaSuccess;
  {
    LaunchScope ls(h, KID_SETUP);
    const cudaError_t e = launch_chained(h->chain_launches, k_setup_problems, dim3((P + 127) / 128), dim3(128), 0, h->stream, 1, probs,
                                         desc_base, n_rows, slot_stride_rows, q_slot, t_slot, P);
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

cudaError_t launch_setup_stereo_problems(Handle* h, MatchProblem* probs, const float* desc_out, const int* n_out,
                                         int F, int K, int carry_slot, const spvo_keypoint* kpts, float band) {
  if (F == 0) return cudaSuccess;
  {
    LaunchScope ls(h, KID_SETUP);
    StereoSetup a;
    a.probs = probs; a.desc_out = desc_out; a.n_out = n_out; a.carry_desc = h->carry_desc; a.carry_n = h->carry_n;
    a.F = F; a.K = K; a.carry_slot = carry_slot; a.kpts = kpts; a.band = band;
    const cudaError_t e = launch_chained(h->chain_launches, k_setup_stereo_problems, dim3((2 * F + 127) / 128), dim3(128), 0,
                                         h->stream, 1, a);
    if (e != cudaSuccess) return e;
  }
  return cudaGetLastError();
}

cudaError_t launch_consistency(Handle* h, int F, int K, const spvo_dmatch* matches, const int* n_matches,
                               const int* q2t, const uint8_t* keep, const int* carry_map, spvo_quad* quads,
                               int* n_quads, const CopyList& cl) {
  const int fb = quads ? F : 0;  // frames whose quadruples are wanted
  if (fb + cl.n == 0) return cudaSuccess;
  LaunchScope ls(h, KID_CONSISTENCY);
  const cudaError_t e = launch_chained(h->chain_launches, k_consistency, dim3(fb + cl.n * kCopyBlocks), dim3(1024), 0, h->stream, 1, fb, K,
                                       matches, n_matches, q2t, keep, carry_map, quads, n_quads, cl);
  if (e != cudaSuccess) return e;
  return cudaGetLastError();
}

cudaError_t launch_set_problem(Handle* h, MatchProblem* probs, const float* q, int N, const float* t, int M,
                               const spvo_keypoint* q_kpts, const spvo_keypoint* t_kpts, float band) {
  {
    LaunchScope ls(h, KID_SETUP);
    k_set_problem<<<1, 1, 0, h->stream>>>(probs, q, N, t, M, q_kpts, t_kpts, band);
  }
  return cudaGetLastError();
}

cudaError_t launch_stereo_filter(Handle* h, const spvo_keypoint* kpts_base, int slot_stride_rows,
                                 const int* q_slot, const int* t_slot, int P, int max_rows,
                                 const spvo_dmatch* matches, const int* n_matches, float stereo_threshold,
                                 float min_disparity, uint8_t* keep) {
  if (P == 0 || max_rows == 0) return cudaSuccess;
  dim3 g((max_rows + 255) / 256, P);
  {
    LaunchScope ls(h, KID_STEREO_FILTER);
    k_stereo_filter<<<g, 256, 0, h->stream>>>(kpts_base, slot_stride_rows, q_slot, t_slot, max_rows, matches,
                                              n_matches, stereo_threshold, min_disparity, keep);
  }
  return cudaGetLastError();
}

}  // namespace spvo
