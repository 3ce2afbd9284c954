// common.cuh -- shared declarations for libspvo_frontend.so (sm_100a only).
// Compiled with -fmad=false: every fp32 operation is a single IEEE round-to-nearest op in source
// order unless an explicit __fmaf_rn is written, so results match oracle/spvo_oracle.cpp bit for bit.
#pragma once

#include <utility>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "spvo_frontend.h"

namespace spvo {

// Score bins used to size k_detect's candidate chunks (decode.cu).
// bin(v) = (bits(1.0f) - bits(v)) >> 14, clamped to [0, 4095]; bin 0 holds the highest scores.
constexpr int kHistBins = 4096;
constexpr int kHistShift = 14;
constexpr uint32_t kOneBits = 0x3F800000u;

struct Handle;

// Kernel classes, for launch counting and the optional per-kernel CUDA-event profile
// (spvo_profile_enable / spvo_profile_read: how bench.py measures the dominant kernel live).
enum KernelId {
  KID_SOFTMAX_HEAT = 0, KID_DETECT, KID_SAMPLE_DESC, KID_DIST_EXACT, KID_ROW_SELECT, KID_COL_SELECT,
  KID_FINALIZE, KID_SETUP, KID_STEREO_FILTER, KID_TC_PREP, KID_TC_GEMM, KID_TC_RERANK, KID_TC_FALLBACK, KID_TC_FILL, KID_TC_TRIAGE, KID_DESC_PLANES, KID_DESC_NORM, KID_CONSISTENCY, KID_CARRY, KID_PREPROCESS, KID_COUNT
};
struct ProfRec {
  int kid;
  cudaEvent_t a, b;
};

// One matching problem, resident on the device (built by k_setup_problems / k_set_problem).
struct MatchProblem {
  const float* q;  // [N,256]
  const float* t;  // [M,256]
  int N, M;
  int a_op, b_op;  // operand slots of q / t in the tensor matcher's bf16 workspace (match_tc.cu)
  // optional row-band mask (the north star's "stereo row-band constraint" as a matcher mask): train row j is a
  // candidate of query i iff |qy[i * ystride] - ty[j * ystride]| <= band.  qy == NULL: no mask.
  const float* qy;
  const float* ty;
  int ystride;
  float band;
};
__device__ __forceinline__ bool band_allowed(const MatchProblem& pr, int i, int j) {
  return !pr.qy || fabsf(__fsub_rn(__ldg(pr.qy + (size_t)i * pr.ystride), __ldg(pr.ty + (size_t)j * pr.ystride))) <= pr.band;
}

// The 2F match problems of a stereo batch (L<->R of frame f = problem f; left_f <-> left_{f-1} = problem F + f, frame
// 0 against the carried previous left image).  Built by k_setup_stereo_problems, or -- one launch fewer -- by block
// (0, 0) of k_desc_normalize when one decode launch covers the whole batch (Handle::stereo_setup).
struct StereoSetup {
  MatchProblem* probs = nullptr;  // null: nothing to do
  const float* desc_out = nullptr;
  const int* n_out = nullptr;
  const float* carry_desc = nullptr;
  const int* carry_n = nullptr;
  int F = 0, K = 0, carry_slot = 0;
  const spvo_keypoint* kpts = nullptr;  // non-null + band >= 0: the L<->R problems carry the row band
  float band = -1.0f;
};
__device__ __forceinline__ void setup_stereo_problem(const StereoSetup& a, int p) {
  const int F = a.F, K = a.K;
  MatchProblem pr;
  pr.qy = pr.ty = nullptr;
  pr.ystride = 0;
  pr.band = -1.0f;
  if (p < F) {
    if (a.kpts && a.band >= 0.0f) {  // masked mode: only the L<->R problems carry the row band
      pr.qy = &a.kpts[(size_t)(2 * p) * K].y;
      pr.ty = &a.kpts[(size_t)(2 * p + 1) * K].y;
      pr.ystride = (int)(sizeof(spvo_keypoint) / sizeof(float));
      pr.band = a.band;
    }
    pr.q = a.desc_out + (size_t)(2 * p) * K * SPVO_DESC_DIM;
    pr.t = a.desc_out + (size_t)(2 * p + 1) * K * SPVO_DESC_DIM;
    pr.N = a.n_out[2 * p];
    pr.M = a.n_out[2 * p + 1];
    pr.a_op = 2 * p;       // operand slot = image index (16-bit rows written by k_desc_normalize)
    pr.b_op = 2 * p + 1;
  } else {
    const int f = p - F;
    pr.q = a.desc_out + (size_t)(2 * f) * K * SPVO_DESC_DIM;
    pr.N = a.n_out[2 * f];
    pr.a_op = 2 * f;
    if (f > 0) {
      pr.t = a.desc_out + (size_t)(2 * (f - 1)) * K * SPVO_DESC_DIM;
      pr.M = a.n_out[2 * (f - 1)];
      pr.b_op = 2 * (f - 1);
    } else {
      pr.t = a.carry_desc;
      pr.M = *a.carry_n;
      pr.b_op = a.carry_slot;
    }
  }
  a.probs[p] = pr;
}

// Where k_desc_normalize additionally writes each image's descriptors for the tensor matcher
// (bf16 rows + fp32 squared norms + per-slot max norm), so the stereo pipeline needs no k_tc_prep.
struct TcSink {
  void* xb = nullptr;        // 16-bit operands [slots][cap][256]: fp16 when fp16 != 0, else bf16
  int fp16 = 0;
  float* nrm = nullptr;      // [slots][cap]
  unsigned* opmax = nullptr; // [slots]
  int cap = 0;               // rows per slot (multiple of 128)
};

// ---- decode.cu ----
// semi / desc: fp32 (in_f16 = 0) or fp16 (in_f16 = 1) device tensors
cudaError_t launch_decode(Handle* h, const void* semi, const void* desc, int in_f16, int B, int H, int W,
                          const spvo_decode_cfg& cfg, spvo_keypoint* kpts, float* desc_out, int* n_out,
                          float* scores, const TcSink* sink = nullptr, bool* sink_filled = nullptr);
cudaError_t launch_div_check(Handle* h, const uint32_t* a_bits, const uint32_t* b_bits, long long n,
                             unsigned long long* mismatches);
// ---- match.cu ----
cudaError_t launch_match_exact(Handle* h, const MatchProblem* probs, int P, int max_rows, int max_cols,
                               const spvo_match_cfg& cfg, spvo_dmatch* out, int* n_matches, int* q2t,
                               int out_stride);
cudaError_t launch_match_tc(Handle* h, const MatchProblem* probs, int P, int max_rows, int max_cols,
                            const spvo_match_cfg& cfg, spvo_dmatch* out, int* n_matches, int* q2t, int out_stride,
                            bool operands_ready = false);
// Stereo pipeline: reserve max_batch image slots + 1 carry slot and return where decode should write.
cudaError_t tc_prepare_slots(Handle* h, int slots, int max_rows, int ndir, TcSink* sink);
// preprocess.cu
bool preprocess_geometry(int rows, int cols, int H, int W, int* crop_rows, int* crop_cols, int* row_off, int* col_off);
cudaError_t launch_preprocess(Handle* h, const uint8_t* imgs, int B, int rows, int cols, int stride, int H, int W,
                              float* out_f, uint8_t* out_u8);

// One device-to-device copy segment of the carry copy (sizes are multiples of 4 bytes).
struct CopySeg {
  const void* src;
  void* dst;
  unsigned long long bytes;
};
// Several small device-to-device copies done by extra blocks of k_consistency (one launch for the whole tail).
struct CopyList {
  CopySeg seg[8];
  int n;
};
// Stereo row-band / min-disparity test (BASE:169-172) applied by k_finalize_matches to the first `nprob` problems
// (problem p: keypoint slots 2p and 2p+1), so that the stereo pipeline needs no separate filter launch.
struct FilterArgs {
  const spvo_keypoint* kpts = nullptr;
  int slot_stride = 0, nprob = 0;
  float stereo_threshold = 0.f, min_disparity = 0.f;
  uint8_t* keep = nullptr;  // [nprob, out_stride]
};
// The three segments (16-bit rows, squared norms, max norm) that copy operand slot src_slot to dst_slot; returns 3.
int tc_copy_slot_segments(Handle* h, int dst_slot, int src_slot, CopySeg* segs);
cudaError_t tc_prep_problem_operands(Handle* h, const MatchProblem* prob);  // k_tc_prep for one problem's q and t
void tc_workspace_free(Handle* h);
cudaError_t launch_setup_problems(Handle* h, MatchProblem* probs, const float* desc_base, const int* n_rows,
                                  int slot_stride_rows, const int* q_slot, const int* t_slot, int P);
cudaError_t launch_setup_stereo_problems(Handle* h, MatchProblem* probs, const float* desc_out, const int* n_out,
                                         int F, int K, int carry_slot, const spvo_keypoint* kpts, float band);
// quads == nullptr: only the copies of `cl` run
cudaError_t launch_consistency(Handle* h, int F, int K, const spvo_dmatch* matches, const int* n_matches,
                               const int* q2t, const uint8_t* keep, const int* carry_map, spvo_quad* quads,
                               int* n_quads, const CopyList& cl);
cudaError_t launch_set_problem(Handle* h, MatchProblem* probs, const float* q, int N, const float* t, int M,
                               const spvo_keypoint* q_kpts = nullptr, const spvo_keypoint* t_kpts = nullptr,
                               float band = -1.0f);
cudaError_t launch_stereo_filter(Handle* h, const spvo_keypoint* kpts_base, int slot_stride_rows,
                                 const int* q_slot, const int* t_slot, int P, int max_rows,
                                 const spvo_dmatch* matches, const int* n_matches, float stereo_threshold,
                                 float min_disparity, uint8_t* keep);

struct Handle {
  int device = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  int max_batch = 0, max_h = 0, max_w = 0, max_k = 0;
  // decode workspace
  unsigned long long* cand_list = nullptr;  // [max_batch, kListCap] candidate keys of k_detect's current generation
  float* heat = nullptr;          // [max_batch, max_h*max_w] heat values of the cells k_softmax_heat stores (sparse)
  uint32_t* spill_thr = nullptr;  // [max_batch] per image slot: storing threshold for the next decode call
  uint2* cellmax = nullptr;       // [max_batch, cells] per-cell (max, second max | argmax) records of the heatmap
  unsigned* nms_bitmap = nullptr; // [max_batch, max_h*ceil(max_w/32)] suppression bitmap of the multi-chunk path
  float* desc_tmp = nullptr;      // [max_batch, 256, max_k] un-normalised descriptor values (k_desc_planes)
  int4* kp_par = nullptr;         // [max_batch, max_k] sampling parameters
  uint8_t* pp_src = nullptr;      // host-form preprocess staging (grown on demand)
  float* pp_dst_f = nullptr;
  uint8_t* pp_dst_u8 = nullptr;
  size_t pp_src_bytes = 0, pp_dst_px = 0;
  uint2* pp_tab = nullptr;        // resize coefficient tables [W + H] of the last geometry (pp_key)
  size_t pp_tab_cap = 0;
  int pp_key[4] = {-1, -1, -1, -1};
  cudaStream_t pp_stream = nullptr;
  cudaEvent_t pp_done = nullptr;   // recorded after the last k_preprocess (orders a table rebuild on another stream)
  unsigned long long* counters = nullptr;  // [8] device counters (slow path images, fallback rows, ...)
  // staging for the host-pointer entry points
  float* st_semi = nullptr;
  float* st_desc = nullptr;
  spvo_keypoint* st_kpts = nullptr;
  float* st_desc_out = nullptr;
  int* st_n = nullptr;
  float* st_scores = nullptr;
  // matching workspace (grown on demand)
  float* dist = nullptr;
  size_t dist_elems = 0;
  int* row_best = nullptr;   // [P*max_rows*2] best / second-best train index per query
  float* row_d = nullptr;    // [P*max_rows*2]
  int* col_best = nullptr;   // [P*max_cols]
  size_t sel_rows = 0, sel_cols = 0;
  MatchProblem* probs = nullptr;
  int probs_cap = 0;
  float* st_q = nullptr;
  float* st_t = nullptr;
  spvo_dmatch* st_matches = nullptr;
  int* st_q2t = nullptr;
  int* st_nm = nullptr;
  spvo_keypoint* st_mkp = nullptr;  // [2 * st_rows] keypoints of a host-form masked match
  size_t st_rows = 0;
  // stereo stream state: previous batch's last left image
  float* carry_desc = nullptr;      // [max_k, 256]
  spvo_keypoint* carry_kpts = nullptr;
  int* carry_n = nullptr;           // device int; 0 when there is no previous frame
  int* carry_map = nullptr;         // [2][max_k] previous frame's L<->R index map (maps_of_indices[PREV_LEFT_PREV_RIGHT]);
                                    // double-buffered: k_consistency reads one half while its copy blocks fill the other
  int carry_parity = 0;             // half holding the map of the last processed frame
  FilterArgs fin_filter;            // consumed by the next k_finalize_matches launch (stereo pipeline)
  spvo_quad* st_quads = nullptr;    // host-form staging
  int* st_nquads = nullptr;
  bool has_prev = false;
  bool carry_tc_valid = false;      // the carry's bf16 copy exists in the tensor matcher's carry slot
  // host-form staging of the stereo outputs
  spvo_dmatch* st_smatches = nullptr;
  int* st_snm = nullptr;
  int* st_sq2t = nullptr;
  uint8_t* st_skeep = nullptr;
  void* tc_ws = nullptr;  // TcWorkspace (match_tc.cu)
  // decode sub-batching (decode.cu: launch_decode)
  cudaStream_t aux_stream[2] = {nullptr, nullptr};
  cudaEvent_t aux_done[2] = {nullptr, nullptr};
  cudaEvent_t aux_fork = nullptr;
  int decode_subbatches = 0;  // 0 = automatic
  bool chain_launches = false;  // programmatic dependent launch for this call's kernels (small calls only)
  StereoSetup stereo_setup;     // set by the stereo pipeline before launch_decode; probs == null otherwise
  bool stereo_setup_done = false;  // launch_decode built the problems (k_desc_normalize), no k_setup launch needed
  // host-form stereo batches: H2D of the next chunk overlaps compute of the current one
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  long long launches = 0;
  // optional CUDA-graph replay of the stereo pipeline (spvo_set_graph_mode): one instantiated graph per call signature
  struct GraphEntry {
    std::vector<unsigned char> sig;
    cudaGraphExec_t exec = nullptr;
    int kernels = 0;
    bool ready = false, flips_parity = false;
  };
  bool graph_mode = false;
  std::vector<GraphEntry> graphs;
  std::vector<std::vector<unsigned char>> graph_seen;  // signatures that ran eagerly once (workspaces are sized)
  // optional per-kernel profile
  bool profiling = false;
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> ev_pool;
  double prof_ms[KID_COUNT] = {0};
  long long prof_n[KID_COUNT] = {0};
  char err[512] = {0};
};

// Counts the launch and, when profiling is on, brackets it with CUDA events on the handle's stream.
struct LaunchScope {
  Handle* h;
  int kid;
  cudaEvent_t a = nullptr;
  static cudaEvent_t get(Handle* h) {
    if (!h->ev_pool.empty()) {
      cudaEvent_t e = h->ev_pool.back();
      h->ev_pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  LaunchScope(Handle* h_, int kid_) : h(h_), kid(kid_) {
    h->launches++;
    if (h->profiling) {
      a = get(h);
      cudaEventRecord(a, h->stream);
    }
  }
  ~LaunchScope() {
    if (a) {
      cudaEvent_t b = get(h);
      cudaEventRecord(b, h->stream);
      h->prof_recs.push_back({kid, a, b});
    }
  }
};

__device__ __forceinline__ uint32_t fbits(float f) { return __float_as_uint(f); }

// Programmatic dependent launch.  Every kernel of the stereo step starts with chain_enter(): it waits until the
// preceding kernel in the stream has completed and its writes are visible (griddepcontrol.wait; a no-op for a plain
// launch), then lets the NEXT kernel's blocks be scheduled (griddepcontrol.launch_dependents) -- they sit at their
// own chain_enter() until this grid is done.  Nothing is read or written before the wait, so the only thing that
// overlaps is launch latency and block scheduling.  Measured: -6 us per call for one small stereo pair with plain
// launches; nothing at 148 pairs per call (the gaps there are ramp-up and tails, not launch latency).
__device__ __forceinline__ void chain_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// exp as specified by the oracle (oracle/spvo_oracle.cpp: oracle_exp): Cephes/Eigen-style expf
// written as explicit IEEE fp32 operations.  Replaces Eigen's packet exp at NN:271.
__device__ __forceinline__ float spvo_exp(float x0) {
  float x = fminf(fmaxf(x0, -88.3762626647949f), 88.3762626647950f);
  float m = floorf(__fmaf_rn(x, 1.44269504088896341f, 0.5f));
  float r = __fmaf_rn(m, -0.6931471805599453f, x);
  float r2 = __fmul_rn(r, r);
  float y = 1.9875691500E-4f;
  y = __fmaf_rn(y, r, 1.3981999507E-3f);
  y = __fmaf_rn(y, r, 8.3334519073E-3f);
  y = __fmaf_rn(y, r, 4.1665795894E-2f);
  y = __fmaf_rn(y, r, 1.6666665459E-1f);
  y = __fmaf_rn(y, r, 5.0000001201E-1f);
  y = __fmaf_rn(y, r2, r);
  y = __fadd_rn(y, 1.0f);
  int e = (int)m + 127;
  float scale = __uint_as_float((uint32_t)e << 23);
  return fmaxf(__fmul_rn(y, scale), x0);
}

// Two spvo_exp at once on Blackwell's packed fp32 pipe (fma / mul / add .rn.f32x2: one issue slot for two IEEE
// operations, each lane rounded exactly like the scalar instruction).  The clamp, floor, float -> int and the final
// max have no packed form and stay scalar.  Same value sequence as spvo_exp, bit for bit.  Measured on B200 in
// k_softmax_heat: a packed instruction occupies the FMA pipe for about four issue cycles (two scalar FFMAs take two),
// so packing trades FMA-pipe time for issue slots -- worth it only where issue is the tighter bound.
__device__ __forceinline__ unsigned long long f32x2_pack(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f32x2_unpack(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long f32x2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ unsigned long long f32x2_mul(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ unsigned long long f32x2_add(unsigned long long a, unsigned long long b) {
  unsigned long long r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ void spvo_exp_x2(float xa0, float xb0, float& ea, float& eb) {
  const float xa = fminf(fmaxf(xa0, -88.3762626647949f), 88.3762626647950f);
  const float xb = fminf(fmaxf(xb0, -88.3762626647949f), 88.3762626647950f);
  const unsigned long long x = f32x2_pack(xa, xb);
  float ta, tb;
  f32x2_unpack(f32x2_fma(x, f32x2_pack(1.44269504088896341f, 1.44269504088896341f), f32x2_pack(0.5f, 0.5f)), ta, tb);
  const float ma = floorf(ta), mb = floorf(tb);
  const unsigned long long r = f32x2_fma(f32x2_pack(ma, mb), f32x2_pack(-0.6931471805599453f, -0.6931471805599453f), x);
  const unsigned long long r2 = f32x2_mul(r, r);
  unsigned long long y = f32x2_pack(1.9875691500E-4f, 1.9875691500E-4f);
  y = f32x2_fma(y, r, f32x2_pack(1.3981999507E-3f, 1.3981999507E-3f));
  y = f32x2_fma(y, r, f32x2_pack(8.3334519073E-3f, 8.3334519073E-3f));
  y = f32x2_fma(y, r, f32x2_pack(4.1665795894E-2f, 4.1665795894E-2f));
  y = f32x2_fma(y, r, f32x2_pack(1.6666665459E-1f, 1.6666665459E-1f));
  y = f32x2_fma(y, r, f32x2_pack(5.0000001201E-1f, 5.0000001201E-1f));
  y = f32x2_fma(y, r2, r);
  y = f32x2_add(y, f32x2_pack(1.0f, 1.0f));
  const float sa = __uint_as_float((uint32_t)((int)ma + 127) << 23), sb = __uint_as_float((uint32_t)((int)mb + 127) << 23);
  float pa, pb;
  f32x2_unpack(f32x2_mul(y, f32x2_pack(sa, sb)), pa, pb);
  ea = fmaxf(pa, xa0);
  eb = fmaxf(pb, xb0);
}


// Launch, optionally with the programmatic-stream-serialization attribute (see chain_enter; ONLY for kernels that
// begin with chain_enter()).  `chained` = Handle::chain_launches, which the entry points set for SMALL calls only:
// measured on B200, the attribute saves 10 us of a 100 us single-pair call but makes a 148-pair step 7.8 % SLOWER
// (1.213 vs 1.126 ms: the early-resident blocks of the next kernel get in the way of the current one's tail).
// SPVO_PDL=0 in the environment turns the attribute off everywhere (A/B measurements).
inline bool chained_launch_enabled() {
  static const bool on = [] {
    const char* e = getenv("SPVO_PDL");
    return !(e && e[0] == '0');
  }();
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_chained(bool chained, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                  cudaStream_t st, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t lc = {};
  lc.gridDim = grid;
  lc.blockDim = block;
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  // not while the stream is being captured: measured on B200 (640x192, K = 500, one pair per call) plain launches gain
  // 6 us per call from the attribute, but a graph whose edges are programmatic replays 4 us SLOWER than plain edges
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (chained && chained_launch_enabled() && cudaStreamIsCapturing(st, &cs) == cudaSuccess &&
      cs == cudaStreamCaptureStatusNone) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)cluster_x;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  lc.attrs = at;
  lc.numAttrs = n;
  return cudaLaunchKernelEx(&lc, kern, static_cast<KArgs>(std::forward<Args>(args))...);
}

// Two auxiliary streams + fork / join events per handle (decode sub-batches, the matcher's concurrent tail), created
// on first use.
inline cudaError_t ensure_aux_streams(Handle* h) {
  if (h->aux_stream[0]) return cudaSuccess;
  cudaError_t e;
  for (int i = 0; i < 2; ++i) {
    if ((e = cudaStreamCreateWithFlags(&h->aux_stream[i], cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&h->aux_done[i], cudaEventDisableTiming)) != cudaSuccess) return e;
  }
  return cudaEventCreateWithFlags(&h->aux_fork, cudaEventDisableTiming);
}

}  // namespace spvo
