// spvo_oracle.cpp -- CPU ORACLE for the SuperPoint decode + descriptor-matching hot path.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load it.  The product path
// (libspvo_frontend.so) never links, loads or calls anything in this directory.
//
// It is a plain-loop restatement of the reference algorithm (all file:line citations are relative
// to /root/reference/src/odml_visual_odometry/):
//   decode   src/feature_detection_neural_network.cpp:264-330  softmax -> dustbin drop -> heatmap
//            src/feature_detection_neural_network.cpp:188-262  threshold, sort, greedy NMS, top-K
//            src/feature_detection_neural_network.cpp:332-431  bilinear descriptor sampling + L2 norm
//   match    src/feature_detection_base.cpp:10-33, 434-500      cv::BFMatcher NN(+cross-check) / kNN-2 + ratio
//   filter   src/feature_detection_base.cpp:169-172             stereo row-band / min-disparity test
//   preproc  src/feature_detection_base.cpp:68-121, src/feature_detection_neural_network.cpp:139-161
//            centre crop to the network's aspect ratio, cv::resize(INTER_LINEAR) on 8UC1, /255, P-matrix patch
//
// Parity pin status.  The reference has NO tests, golden vectors or fixtures for this path
// (SURVEY.md section 4), and its decode cannot be compiled here (needs Eigen, OpenCV C++ headers,
// ROS, TensorRT).  Therefore:
//   * DECODE: "parity unpinned" against reference-produced outputs.  Where the reference leaves the
//     fp32 bits to a third-party library (Eigen packet exp, Eigen reduction order, std::sort tie
//     order) this oracle SPECIFIES them (see oracle_exp, canonical tie-break, norm reduction order)
//     and the CUDA path implements the same specification bit for bit.
//   * MATCH: pinned against the reference's real matcher, cv::BFMatcher (OpenCV; reference pins
//     4.5.4 in CMakeLists.txt:14, the cv2 4.13.0 wheel is importable in this image).  The
//     arithmetic below (hal::normL2Sqr_ lane order + sqrt, first-index ties, mutual-argmin
//     cross-check, stable top-2) is checked bit-for-bit against cv2 in tests/test_oracle_match.py
//     and against committed cv2-generated fixtures in tests/golden/.
//   * PREPROCESS: the resize is pinned bit-for-bit against cv2.resize (tests/test_oracle_preprocess.py, live and
//     through committed cv2-generated fixtures); crop / P-matrix / scaling are plain fp32 statements of BASE:68-121.
//
// Build: see oracle/Makefile (g++ -O2 -mavx2 -mfma -ffp-contract=off; no fast-math, so every
// fp32 operation below is a single correctly rounded IEEE operation in the order written).

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

extern "C" {

// Layout-compatible with cv::KeyPoint (28 B) and cv::DMatch (16 B); see include/spvo_frontend.h.
struct spvo_keypoint {
  float x, y, size, angle, response;
  int octave, class_id;
};
struct spvo_dmatch {
  int queryIdx, trainIdx, imgIdx;
  float distance;
};
static_assert(sizeof(spvo_keypoint) == 28, "cv::KeyPoint layout");
static_assert(sizeof(spvo_dmatch) == 16, "cv::DMatch layout");

}  // extern "C"

namespace {

inline uint32_t f2u(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  return u;
}
inline float u2f(uint32_t u) {
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

// ---------------------------------------------------------------------------------------------
// exp.  The reference calls Eigen's tensor .exp() (feature_detection_neural_network.cpp:271),
// i.e. Eigen's vectorised packet exp.  Eigen is a third-party dependency that is NOT in the
// reference tree (find_package(Eigen3 3.3), CMakeLists.txt:21, version unpinned, built with
// -O3 -march=native).  Its published algorithm (Cephes expf: clamp, n = floor(x*log2e + 1/2),
// r = x - n*ln2, degree-5 polynomial, scale by 2^n; FMA form used when the build has FMA) is
// restated here as an explicit sequence of IEEE fp32 operations.  The CUDA path executes the same
// sequence with __fmaf_rn/__fmul_rn/__fadd_rn, so oracle and GPU agree bit for bit; a real Eigen
// build agrees to within 1-2 ulp, which can only matter at exact ties / the strict threshold.
// ---------------------------------------------------------------------------------------------
inline float oracle_exp(float x0) {
  float x = std::fmin(std::fmax(x0, -88.3762626647949f), 88.3762626647950f);
  float m = std::floor(std::fmaf(x, 1.44269504088896341f, 0.5f));
  float r = std::fmaf(m, -0.6931471805599453f, x);
  float r2 = r * r;
  float y = 1.9875691500E-4f;
  y = std::fmaf(y, r, 1.3981999507E-3f);
  y = std::fmaf(y, r, 8.3334519073E-3f);
  y = std::fmaf(y, r, 4.1665795894E-2f);
  y = std::fmaf(y, r, 1.6666665459E-1f);
  y = std::fmaf(y, r, 5.0000001201E-1f);
  y = std::fmaf(y, r2, r);
  y = y + 1.0f;
  int e = (int)m + 127;                       // m in [-127, 128] after the clamp
  float scale = u2f((uint32_t)e << 23);       // 2^m  (m = -127 -> 0.0f, m = 128 -> +inf)
  return std::fmax(y * scale, x0);
}

// ---------------------------------------------------------------------------------------------
// SENSITIVITY VARIANTS (tests/test_oracle_sensitivity.py).  The decode oracle is "parity unpinned": the
// reference's exp / reduction bits belong to an Eigen build that cannot be reproduced here.  These
// variants restate the other plausible builds so that the tests can MEASURE how far a real build can move
// the result:
//   kVarLibmExp      exp = glibc expf (correctly rounded in practice; the upper bound on accuracy)
//   kVarCephesNoFma  Eigen 3.3 pexp as compiled WITHOUT FMA (SSE2 build): fx = x*LOG2E + 0.5 (two ops), floor,
//                    r = (x - fx*C1) - fx*C2 with Cephes' split ln2, Horner with separate multiply / add
//   kVarPairwiseSum  channel sum as a balanced tree (a vectorised / tree reduction) instead of c = 0..64 in order
// ---------------------------------------------------------------------------------------------
enum { kVarLibmExp = 1, kVarCephesNoFma = 2, kVarPairwiseSum = 4 };

inline float cephes_exp_nofma(float x0) {
  float x = std::fmin(std::fmax(x0, -88.3762626647949f), 88.3762626647950f);
  float fx = x * 1.44269504088896341f;
  fx = fx + 0.5f;
  fx = std::floor(fx);
  float tmp = fx * 0.693359375f;
  float z = fx * -2.12194440e-4f;
  x = x - tmp;
  x = x - z;
  z = x * x;
  float y = 1.9875691500E-4f;
  y = y * x; y = y + 1.3981999507E-3f;
  y = y * x; y = y + 8.3334519073E-3f;
  y = y * x; y = y + 4.1665795894E-2f;
  y = y * x; y = y + 1.6666665459E-1f;
  y = y * x; y = y + 5.0000001201E-1f;
  y = y * z; y = y + x;
  y = y + 1.0f;
  int e = (int)fx + 127;
  float scale = u2f((uint32_t)e << 23);
  return std::fmax(y * scale, x0);
}

inline float variant_exp(float x, int variant) {
  if (variant & kVarLibmExp) return std::exp(x);
  if (variant & kVarCephesNoFma) return cephes_exp_nofma(x);
  return oracle_exp(x);
}

inline float tree_sum(const float* v, int n) {  // balanced binary tree, left half first
  if (n == 1) return v[0];
  const int h = (n + 1) / 2;
  return tree_sum(v, h) + tree_sum(v + h, n - h);
}

template <class F>
void run_parallel(int n, int num_threads, F&& fn) {
  if (num_threads <= 1 || n <= 1) {
    fn(0, n);
    return;
  }
  int nt = std::min(num_threads, n);
  std::vector<std::thread> th;
  th.reserve(nt);
  for (int t = 0; t < nt; ++t) {
    int lo = (int)((int64_t)n * t / nt), hi = (int)((int64_t)n * (t + 1) / nt);
    th.emplace_back([lo, hi, &fn]() { fn(lo, hi); });
  }
  for (auto& t : th) t.join();
}

// ---------------------------------------------------------------------------------------------
// D1 + D2: softmax over 65 channels (no max subtraction, +1e-5 in the denominator), dustbin drop,
// depth-to-space.  feature_detection_neural_network.cpp:266-326.
//   e_c  = exp(x_c)                                  (:271)
//   s    = sum_{c=0..64} e_c, accumulated in channel order starting from 0.0f   (:274-279)
//   p_c  = e_c / (s + 1e-5f)   true division          (:280-284)
//   heat[8*hc + i, 8*wc + j] = p_{8*i + j}[hc, wc]    (:289-326)
// ---------------------------------------------------------------------------------------------
void softmax_heatmap(const float* semi, int Hc, int Wc, float* heat, int num_threads, int variant = 0) {
  const int W = Wc * 8;
  const int cells = Hc * Wc;
  run_parallel(Hc, num_threads, [&](int r_lo, int r_hi) {
    float e[65];
    for (int hc = r_lo; hc < r_hi; ++hc) {
      for (int wc = 0; wc < Wc; ++wc) {
        const float* src = semi + hc * Wc + wc;
        float s = 0.0f;
        if (variant == 0) {
          for (int c = 0; c < 65; ++c) {
            e[c] = oracle_exp(src[(size_t)c * cells]);
            s = s + e[c];
          }
        } else {  // sensitivity variants only (never the specification)
          for (int c = 0; c < 65; ++c) e[c] = variant_exp(src[(size_t)c * cells], variant);
          if (variant & kVarPairwiseSum) s = tree_sum(e, 65);
          else for (int c = 0; c < 65; ++c) s = s + e[c];
        }
        const float denom = s + 0.00001f;
        for (int i = 0; i < 8; ++i) {
          float* dst = heat + (size_t)(8 * hc + i) * W + 8 * wc;
          for (int j = 0; j < 8; ++j) dst[j] = e[8 * i + j] / denom;
        }
      }
    }
  });
}

struct Cand {
  float score;
  int x, y;
};

// ---------------------------------------------------------------------------------------------
// D3-D5: threshold (strict >), column-major candidate enumeration, descending-score sort,
// greedy box NMS, border filter, top-K.  feature_detection_neural_network.cpp:188-262.
//
// Tie order.  The reference uses std::sort (unstable) with a comparator that only looks at the
// score (:214-217), so the relative order of candidates with bit-equal scores is
// implementation-defined.  The CANONICAL order specified here (and implemented by the CUDA path)
// is the one a stable sort of the reference's column-major candidate list gives: score
// descending, then x ascending, then y ascending.  faithful_sort=1 runs the reference's exact
// std::sort call instead, to measure how often that differs (only on exact ties).
// ---------------------------------------------------------------------------------------------
int detect_one(const float* heat, int H, int W, float conf, int dist, int border, int K,
               int faithful_sort, spvo_keypoint* kpts, float* scores, int* walked_out,
               int* ncand_out) {
  std::vector<Cand> cands;
  for (int x = 0; x < W; ++x)             // column-major: Eigen::SparseMatrix<bool> default (:205-213)
    for (int y = 0; y < H; ++y) {
      float v = heat[(size_t)y * W + x];
      if (v > conf) cands.push_back({v, x, y});
    }
  if (faithful_sort) {
    std::sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.score > b.score; });
  } else {
    std::stable_sort(cands.begin(), cands.end(),
                     [](const Cand& a, const Cand& b) { return a.score > b.score; });
  }
  std::vector<uint8_t> nms((size_t)H * W, 0);
  int emitted = 0, walked = 0;
  for (const Cand& c : cands) {
    if (emitted >= K) break;               // (:256-257) checked after each candidate; K<=0 emits nothing
    ++walked;
    if (!nms[(size_t)c.y * W + c.x]) {
      if (c.y >= border && c.y + border < H && c.x >= border && c.x + border < W) {   // (:239-244)
        if (kpts) kpts[emitted] = {(float)c.x, (float)c.y, 1.0f, -1.0f, 0.0f, 0, -1};
        if (scores) scores[emitted] = c.score;
        ++emitted;
      }
      for (int r = c.y - dist; r < c.y + dist + 1; ++r) {                              // (:246-254)
        if (r < 0 || r >= H) continue;
        for (int q = c.x - dist; q < c.x + dist + 1; ++q) {
          if (q < 0 || q >= W) continue;
          nms[(size_t)r * W + q] = 1;
        }
      }
    }
  }
  if (walked_out) *walked_out = walked;
  if (ncand_out) *ncand_out = (int)cands.size();
  return emitted;
}

// ---------------------------------------------------------------------------------------------
// D6-D8: align-corners bilinear sampling of the coarse descriptor map + L2 normalisation.
// feature_detection_neural_network.cpp:366-431 (coordinates :377-392, blend :423-427,
// normalize() :428).  Expression order: each term is (vec * s1) * s2, terms summed left to right,
// no FMA contraction.  The squared norm is SPECIFIED as: 32 partial sums p_l = sum over channels
// c = l, l+32, ... (ascending, p = p + v*v unfused), then a 5-level xor butterfly
// (p_l + p_{l^16}, then ^8, ^4, ^2, ^1) -- Eigen's own packet reduction order is build-dependent,
// hence the 1e-5 tolerance north_star states against a real reference build.  v /= sqrt(z) by true
// division when z > 0 (Eigen normalize()).  Out-of-range neighbour cells (only reachable with
// border_remove == 0, where the reference reads out of bounds) are clamped.
// ---------------------------------------------------------------------------------------------
void sample_descriptor(const float* desc, int Hc, int Wc, int H, int W, int x, int y, float* out) {
  const int D = 256;
  const size_t cells = (size_t)Hc * Wc;
  const float r8 = (float)y / (float)(H - 1) * (float)(H / 8 - 1);
  const float c8 = (float)x / (float)(W - 1) * (float)(W / 8 - 1);
  const int r0 = (int)std::floor(r8), c0 = (int)std::floor(c8);
  const float rr = 1.0f - (r8 - (float)r0);
  const float cr = 1.0f - (c8 - (float)c0);
  const float irr = 1.0f - rr, icr = 1.0f - cr;
  const int r1 = std::min(r0 + 1, Hc - 1), c1 = std::min(c0 + 1, Wc - 1);
  const float* tl = desc + (size_t)r0 * Wc + c0;
  const float* tr = desc + (size_t)r0 * Wc + c1;
  const float* bl = desc + (size_t)r1 * Wc + c0;
  const float* br = desc + (size_t)r1 * Wc + c1;
  float part[32];
  for (int l = 0; l < 32; ++l) part[l] = 0.0f;
  for (int c = 0; c < D; ++c) {
    const size_t o = (size_t)c * cells;
    float t1 = (tl[o] * rr) * cr;
    float t2 = (tr[o] * rr) * icr;
    float t3 = (bl[o] * irr) * cr;
    float t4 = (br[o] * irr) * icr;
    float v = ((t1 + t2) + t3) + t4;
    out[c] = v;
    float sq = v * v;
    part[c & 31] = part[c & 31] + sq;
  }
  for (int off = 16; off >= 1; off >>= 1) {
    float nxt[32];
    for (int l = 0; l < 32; ++l) nxt[l] = part[l] + part[l ^ off];
    for (int l = 0; l < 32; ++l) part[l] = nxt[l];
  }
  const float z = part[0];
  if (z > 0.0f) {
    const float n = std::sqrt(z);
    for (int c = 0; c < D; ++c) out[c] = out[c] / n;
  }
}

// ---------------------------------------------------------------------------------------------
// M2: the distance cv::BFMatcher(NORM_L2) computes.  OpenCV is a third-party dependency absent
// from /root/reference (pinned 4.5.4, CMakeLists.txt:14).  batchDistL2_32f -> hal::normL2Sqr_
// (SIMD128 universal intrinsics: four v_float32x4 accumulators over 16-element blocks, unfused
// multiply then add) followed by std::sqrt.  For element j = 16*blk + 4*k + l, accumulator k,
// lane l:  s[k][l] += (a_j - b_j) * (a_j - b_j);  v[l] = ((s0+s1)+s2)+s3;  d2 = (v0+v2)+(v1+v3).
// A scalar tail (dim % 16 elements, OpenCV's 4-way unrolled loop then singles) follows for
// generality; SuperPoint's 256 has no tail.  Verified bit-for-bit against cv2 in the tests.
// ---------------------------------------------------------------------------------------------
typedef float v4sf __attribute__((vector_size(16)));

inline float l2sqr_cv(const float* a, const float* b, int n) {
  v4sf s0 = {0, 0, 0, 0}, s1 = s0, s2 = s0, s3 = s0;
  int j = 0;
  for (; j + 16 <= n; j += 16) {
    v4sf a0, a1, a2, a3, b0, b1, b2, b3;
    std::memcpy(&a0, a + j, 16); std::memcpy(&a1, a + j + 4, 16);
    std::memcpy(&a2, a + j + 8, 16); std::memcpy(&a3, a + j + 12, 16);
    std::memcpy(&b0, b + j, 16); std::memcpy(&b1, b + j + 4, 16);
    std::memcpy(&b2, b + j + 8, 16); std::memcpy(&b3, b + j + 12, 16);
    v4sf t0 = a0 - b0, t1 = a1 - b1, t2 = a2 - b2, t3 = a3 - b3;
    s0 = s0 + t0 * t0;
    s1 = s1 + t1 * t1;
    s2 = s2 + t2 * t2;
    s3 = s3 + t3 * t3;
  }
  v4sf v = ((s0 + s1) + s2) + s3;
  float d = (v[0] + v[2]) + (v[1] + v[3]);
  for (; j < n; ++j) {
    float t = a[j] - b[j];
    d = d + t * t;
  }
  return d;
}

inline float l2dist_cv(const float* a, const float* b, int n) { return std::sqrt(l2sqr_cv(a, b, n)); }

}  // namespace

extern "C" {

float spvo_oracle_exp(float x) { return oracle_exp(x); }

float spvo_oracle_l2dist(const float* a, const float* b, int dim) { return l2dist_cv(a, b, dim); }

// Heatmap only (rows D1-D2).  semi [B,65,Hc,Wc] -> heat [B,H,W].
int spvo_oracle_heatmap(const float* semi, int B, int H, int W, float* heat, int num_threads) {
  if (!semi || !heat || B < 0 || H <= 0 || W <= 0 || H % 8 || W % 8) return 1;
  const int Hc = H / 8, Wc = W / 8;
  for (int b = 0; b < B; ++b)
    softmax_heatmap(semi + (size_t)b * 65 * Hc * Wc, Hc, Wc, heat + (size_t)b * H * W, num_threads);
  return 0;
}

// Sensitivity study only: heatmap with a VARIANT exp / channel-sum (see kVar* above; variant 0 = the specification).
int spvo_oracle_heatmap_variant(const float* semi, int B, int H, int W, float* heat, int num_threads, int variant) {
  if (!semi || !heat || B < 0 || H <= 0 || W <= 0 || H % 8 || W % 8) return 1;
  const int Hc = H / 8, Wc = W / 8;
  for (int b = 0; b < B; ++b)
    softmax_heatmap(semi + (size_t)b * 65 * Hc * Wc, Hc, Wc, heat + (size_t)b * H * W, num_threads, variant);
  return 0;
}

// Rows D3-D5 on a GIVEN heatmap (feature_detection_neural_network.cpp:188-262): used by the sensitivity study to run
// the reference's walk on variant heatmaps.  kpts_out [B,K], scores_out [B,K] (optional), n_out [B].
int spvo_oracle_detect(const float* heat, int B, int H, int W, float conf_thresh, int dist_thresh, int border_remove,
                       int max_keypoints, int faithful_sort, spvo_keypoint* kpts_out, float* scores_out, int* n_out) {
  if (!heat || !kpts_out || !n_out || B < 0 || H <= 0 || W <= 0 || max_keypoints < 0) return 1;
  for (int b = 0; b < B; ++b) {
    spvo_keypoint* kp = kpts_out + (size_t)b * max_keypoints;
    const int n = detect_one(heat + (size_t)b * H * W, H, W, conf_thresh, dist_thresh, border_remove, max_keypoints,
                             faithful_sort, kp, scores_out ? scores_out + (size_t)b * max_keypoints : nullptr, nullptr,
                             nullptr);
    for (int i = n; i < max_keypoints; ++i) kp[i] = spvo_keypoint{0, 0, 0, 0, 0, 0, 0};
    n_out[b] = n;
  }
  return 0;
}

// Full decode.  Mirrors SuperPointFeatureFrontEnd::postprocessDetectionAndDescription()
// (feature_detection_neural_network.cpp:264-364) for a batch of B images.
//   semi [B,65,H/8,W/8], desc [B,256,H/8,W/8]  (NCHW fp32, the TensorRT output buffers, hpp:382-384)
//   kpts_out [B,K], desc_out [B,K,256], n_out [B]; optional scores_out [B,K], heat_out [B,H,W],
//   walked_out [B] (candidates visited before the K cut), ncand_out [B].
// num_threads mirrors the reference's Eigen::ThreadPool size (hpp:314-317): it parallelises the
// dense tensor passes and the per-keypoint sampling; the per-image sort + greedy walk is serial as
// in the reference (:328-330).
int spvo_oracle_decode(const float* semi, const float* desc, int B, int H, int W, float conf_thresh,
                       int dist_thresh, int border_remove, int max_keypoints, int faithful_sort,
                       spvo_keypoint* kpts_out, float* desc_out, int* n_out, float* scores_out,
                       float* heat_out, int* walked_out, int* ncand_out, int num_threads) {
  if (!semi || !kpts_out || !n_out || B < 0 || H <= 0 || W <= 0 || H % 8 || W % 8 ||
      max_keypoints < 0 || dist_thresh < 0 || border_remove < 0)
    return 1;
  const int Hc = H / 8, Wc = W / 8, K = max_keypoints;
  std::vector<float> heat_local;
  if (!heat_out) heat_local.resize((size_t)H * W);
  for (int b = 0; b < B; ++b) {
    float* heat = heat_out ? heat_out + (size_t)b * H * W : heat_local.data();
    softmax_heatmap(semi + (size_t)b * 65 * Hc * Wc, Hc, Wc, heat, num_threads);
    spvo_keypoint* kp = kpts_out + (size_t)b * K;
    int n = detect_one(heat, H, W, conf_thresh, dist_thresh, border_remove, K, faithful_sort, kp,
                       scores_out ? scores_out + (size_t)b * K : nullptr,
                       walked_out ? walked_out + b : nullptr, ncand_out ? ncand_out + b : nullptr);
    n_out[b] = n;
    if (desc && desc_out) {
      const float* dmap = desc + (size_t)b * 256 * Hc * Wc;
      float* dout = desc_out + (size_t)b * K * 256;
      run_parallel(n, num_threads, [&](int lo, int hi) {
        for (int i = lo; i < hi; ++i)
          sample_descriptor(dmap, Hc, Wc, H, W, (int)kp[i].x, (int)kp[i].y, dout + (size_t)i * 256);
      });
    }
  }
  return 0;
}

// Matching.  Mirrors FeatureFrontEnd::matchDescriptors (feature_detection_base.cpp:434-500) with
// matcher_ = cv::BFMatcher(NORM_L2, cross_check && selector != KNN) (:27-28).
//   mode 0: NN               matcher_->match, crossCheck=false        (:463)
//   mode 1: NN + cross-check matcher_->match, crossCheck=true         (:463; batchDistance mutual first-argmin)
//   mode 2: kNN-2 + ratio    knnMatch(k=2); keep m0 iff m0.d < ratio*m1.d   (:466-472)
// Output DMatch list in ascending queryIdx, imgIdx = 0; q2t[N] = trainIdx or -1 (:483-491).
// Defined edge cases (the reference has UB / throws): N==0 or M==0 -> 0 matches; mode 2 with
// M < 2 -> 0 matches.
// Matching with an optional ROW-BAND MASK (BASELINE north star: "run left<->right under a stereo row-band constraint").
// The reference itself matches unmasked and applies |y_l - y_r| <= stereo_threshold afterwards
// (feature_detection_base.cpp:169-172); cv::BFMatcher rejects crossCheck together with a mask, so the masked
// cross-check is DEFINED as what two masked cv::BFMatcher(NORM_L2, false)::match calls (query->train with the mask,
// train->query with its transpose) plus a manual mutual test give -- pinned against cv2 in tests/test_oracle_match.py.
//   allowed(i, j) = |qy[i] - ty[j]| <= band      (band < 0 or qy == NULL: no mask)
// A query with no allowed train row has no match; kNN-ratio needs two allowed rows (BASE:469 would be UB).
int spvo_oracle_match_masked(const float* q, int N, const float* t, int M, int dim, int mode, float ratio,
                             const float* qy, const float* ty, float band, spvo_dmatch* out, int* n_matches,
                             int* q2t, int num_threads) {
  if (N < 0 || M < 0 || dim <= 0 || !n_matches || mode < 0 || mode > 2) return 1;
  *n_matches = 0;
  if (q2t) for (int i = 0; i < N; ++i) q2t[i] = -1;
  if (N == 0 || M == 0) return 0;
  if (mode == 2 && M < 2) return 0;
  const bool masked = qy && ty && band >= 0.0f;
  auto allowed = [&](int i, int j) { return !masked || std::fabs(qy[i] - ty[j]) <= band; };
  std::vector<int> best(N, -1), tbest, nallowed(N, 0);
  std::vector<float> d0(N, 0.f), d1(N, 0.f);
  run_parallel(N, num_threads, [&](int lo, int hi) {
    for (int i = lo; i < hi; ++i) {
      const float* a = q + (size_t)i * dim;
      // cv::batchDistance top-K insertion: strict '<' keeps the first index on ties, stable for K=2.
      float b0 = INFINITY, b1 = INFINITY;
      int i0 = -1, na = 0;
      for (int j = 0; j < M; ++j) {
        if (!allowed(i, j)) continue;
        ++na;
        float d = l2dist_cv(a, t + (size_t)j * dim, dim);
        if (d < b0 || i0 < 0) {
          if (i0 >= 0) b1 = b0;
          b0 = d;
          i0 = j;
        } else if (d < b1) {
          b1 = d;
        }
      }
      best[i] = i0;
      d0[i] = b0;
      d1[i] = b1;
      nallowed[i] = na;
    }
  });
  if (mode == 1) {
    tbest.assign(M, -1);
    run_parallel(M, num_threads, [&](int lo, int hi) {
      for (int j = lo; j < hi; ++j) {
        const float* a = t + (size_t)j * dim;
        float b0 = INFINITY;
        int i0 = -1;
        for (int i = 0; i < N; ++i) {
          if (!allowed(i, j)) continue;
          float d = l2dist_cv(a, q + (size_t)i * dim, dim);
          if (d < b0 || i0 < 0) {
            b0 = d;
            i0 = i;
          }
        }
        tbest[j] = i0;
      }
    });
  }
  int n = 0;
  for (int i = 0; i < N; ++i) {
    bool keep = best[i] >= 0;
    if (keep && mode == 1) keep = (tbest[best[i]] == i);
    if (keep && mode == 2) keep = nallowed[i] >= 2 && (d0[i] < ratio * d1[i]);
    if (!keep) continue;
    if (out) out[n] = {i, best[i], 0, d0[i]};
    if (q2t) q2t[i] = best[i];
    ++n;
  }
  *n_matches = n;
  return 0;
}

int spvo_oracle_match(const float* q, int N, const float* t, int M, int dim, int mode, float ratio,
                      spvo_dmatch* out, int* n_matches, int* q2t, int num_threads) {
  return spvo_oracle_match_masked(q, N, t, M, dim, mode, ratio, nullptr, nullptr, -1.0f, out, n_matches, q2t,
                                  num_threads);
}

// S1: stereo row-band / min-disparity test applied to L<->R matches
// (feature_detection_base.cpp:169-172).  keep[m] = 1 iff the match survives.
int spvo_oracle_stereo_filter(const spvo_keypoint* kl, const spvo_keypoint* kr, const spvo_dmatch* m,
                              int n, float stereo_threshold, float min_disparity, uint8_t* keep) {
  for (int i = 0; i < n; ++i) {
    const spvo_keypoint& a = kl[m[i].queryIdx];
    const spvo_keypoint& b = kr[m[i].trainIdx];
    bool drop = std::fabs(a.y - b.y) > stereo_threshold || std::fabs(a.x - b.x) < min_disparity;
    keep[i] = drop ? 0 : 1;
  }
  return 0;
}

// Post-match consistency walk of solveStereoOdometry (feature_detection_base.cpp:156-207): for each L<->R
// match in list order keep (currL, currR, prevL, prevR) iff the current left keypoint has a temporal match
// (:160), the stereo test passes (:169-172, given as keep[]), and the matched previous-left keypoint had a
// stereo match in the previous frame (:181).  map_t = maps_of_indices[CURR_LEFT_PREV_LEFT],
// map_prev = maps_of_indices[PREV_LEFT_PREV_RIGHT].  Returns the number of quadruples written.
int spvo_oracle_consistency(const spvo_dmatch* stereo, int n, const int* map_t, const uint8_t* keep,
                            const int* map_prev, int* quads_out) {
  int k = 0;
  for (int i = 0; i < n; ++i) {
    const int il = stereo[i].queryIdx;
    if (map_t[il] == -1) continue;
    if (!keep[i]) continue;
    const int ipl = map_t[il];
    if (map_prev[ipl] == -1) continue;
    quads_out[4 * k + 0] = il;
    quads_out[4 * k + 1] = stereo[i].trainIdx;
    quads_out[4 * k + 2] = ipl;
    quads_out[4 * k + 3] = map_prev[ipl];
    ++k;
  }
  return k;
}


// ---------------------------------------------------------------------------------------------
// preprocessImage (BASE:68-121 + NN:139-161): crop -> cv::resize(INTER_LINEAR, 8UC1) -> * (1/255)
// ---------------------------------------------------------------------------------------------
// Crop geometry of BASE:71-113 (int <- float conversions truncate, as the C++ assignments do).
int spvo_oracle_crop_geometry(int rows, int cols, int H, int W, int* crop_rows, int* crop_cols, int* row_off,
                              int* col_off) {
  if (rows <= 0 || cols <= 0 || H <= 0 || W <= 0) return -1;
  int img_rows = rows, img_cols = cols, ro = 0, co = 0;
  const float real_ar = static_cast<float>(cols) / static_cast<float>(rows);
  const float expected_ar = static_cast<float>(W) / static_cast<float>(H);
  if (expected_ar > real_ar) {
    img_rows = static_cast<int>(static_cast<float>(img_cols) / expected_ar);  // BASE:85
    ro = (rows - img_rows) / 2;
  } else if (expected_ar < real_ar) {
    img_cols = static_cast<int>(static_cast<float>(img_rows) * expected_ar);  // BASE:101
    co = (cols - img_cols) / 2;
  }
  if (img_rows <= 0 || img_cols <= 0) return -1;
  *crop_rows = img_rows; *crop_cols = img_cols; *row_off = ro; *col_off = co;
  return 0;
}

// cv::resize(src, dst, Size(dw, dh), 0, 0, INTER_LINEAR) for 8UC1, OpenCV's fixed-point path
// (imgproc/src/resize.cpp: 11-bit coefficients, HResizeLinear to int, VResizeLinear<uchar,...>);
// an exact 2x decimation takes OpenCV's INTER_AREA fast path ((a+b+c+d+2)>>2).
static void resize_linear_8u(const uint8_t* src, int sh, int sw, int sstride, uint8_t* dst, int dh, int dw) {
  if (dw * 2 == sw && dh * 2 == sh) {
    for (int y = 0; y < dh; ++y)
      for (int x = 0; x < dw; ++x) {
        const uint8_t* p = src + (size_t)(2 * y) * sstride + 2 * x;
        dst[(size_t)y * dw + x] = (uint8_t)((p[0] + p[1] + p[sstride] + p[sstride + 1] + 2) >> 2);
      }
    return;
  }
  const double scale_x = 1.0 / ((double)dw / sw), scale_y = 1.0 / ((double)dh / sh);
  std::vector<int> xofs(dw), xofs1(dw), a0(dw), a1(dw);
  for (int dx = 0; dx < dw; ++dx) {
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)std::floor(fx);
    fx -= sx;
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= sw - 1) { fx = 0.f; sx = sw - 1; }
    xofs[dx] = sx;
    xofs1[dx] = std::min(sx + 1, sw - 1);
    a0[dx] = (int)std::lrintf((1.f - fx) * 2048.f);  // saturate_cast<short>: round half to even, no saturation here
    a1[dx] = (int)std::lrintf(fx * 2048.f);
  }
  std::vector<int> row0(dw), row1(dw);
  for (int dy = 0; dy < dh; ++dy) {
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)std::floor(fy);
    fy -= sy;
    const int b0 = (int)std::lrintf((1.f - fy) * 2048.f), b1 = (int)std::lrintf(fy * 2048.f);
    const int y0 = std::min(std::max(sy, 0), sh - 1), y1 = std::min(std::max(sy + 1, 0), sh - 1);
    const uint8_t* s0 = src + (size_t)y0 * sstride;
    const uint8_t* s1 = src + (size_t)y1 * sstride;
    for (int dx = 0; dx < dw; ++dx) {
      row0[dx] = s0[xofs[dx]] * a0[dx] + s0[xofs1[dx]] * a1[dx];
      row1[dx] = s1[xofs[dx]] * a0[dx] + s1[xofs1[dx]] * a1[dx];
    }
    for (int dx = 0; dx < dw; ++dx)
      dst[(size_t)dy * dw + dx] = (uint8_t)((((b0 * (row0[dx] >> 4)) >> 16) + ((b1 * (row1[dx] >> 4)) >> 16) + 2) >> 2);
  }
}

// img [rows, stride] 8UC1 -> resized [H, W] u8 (the image the reference keeps in images_dq, may be NULL) and
// input [H, W] fp32 = resized * (1/255) (NN:159).  P: 3x4 row-major projection matrix patched in place (may be NULL).
int spvo_oracle_preprocess(const uint8_t* img, int rows, int cols, int stride, int H, int W, float* input,
                           uint8_t* resized, float* P) {
  int cr, cc, ro, co;
  if (!img || stride < cols || spvo_oracle_crop_geometry(rows, cols, H, W, &cr, &cc, &ro, &co)) return -1;
  std::vector<uint8_t> tmp((size_t)H * W);
  resize_linear_8u(img + (size_t)ro * stride + co, cr, cc, stride, tmp.data(), H, W);
  const float k = 1.0f / 255.0f;
  if (input)
    for (size_t i = 0; i < (size_t)H * W; ++i) input[i] = (float)tmp[i] * k;
  if (resized) std::memcpy(resized, tmp.data(), tmp.size());
  if (P) {
    P[1 * 4 + 2] -= (float)ro;  // BASE:93 (only one of the two offsets is non-zero)
    P[0 * 4 + 2] -= (float)co;  // BASE:109
    const float r = (float)W / (float)cc;  // BASE:119
    for (int i = 0; i < 8; ++i) P[i] *= r;  // rows 0 and 1
  }
  return 0;
}

}  // extern "C"
