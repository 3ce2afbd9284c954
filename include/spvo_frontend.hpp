// spvo_frontend.hpp -- header-only C++ host mirror of the reference's front-end classes on top of
// the C ABI (spvo_frontend.h).  Same names, members and call order as the reference
// (src/odml_visual_odometry/include/odml_visual_odometry/feature_detection.hpp, "HPP"):
//
//   reference                                          here
//   -------------------------------------------------  ------------------------------------------
//   class FeatureFrontEnd              (HPP:96-178)    spvo::FeatureFrontEnd
//   class SuperPointFeatureFrontEnd    (HPP:253-391)   spvo::SuperPointFeatureFrontEnd
//   postprocessDetectionAndDescription (HPP:327)       same name; runs spvo_decode (CUDA)
//   matchDescriptors(MatchType)        (HPP:118)       same name; runs spvo_match  (CUDA)
//   keypoints_dq / descriptors_dq / cv_DMatches_list   same names (HPP:124,128,129)
//   maps_of_indices (protected, HPP:161)               same name, public accessor mapOfIndices()
//   output_det_data_ / output_desc_data_ (HPP:383-384) same names: the host buffers the network
//                                                      output is copied into (NN:170-176)
//
// When OpenCV headers are present the containers hold real cv::KeyPoint / cv::Mat / cv::DMatch
// (the PODs are layout-compatible, asserted below), so `solveStereoOdometry` (BASE:125-399) and the
// visualisers can consume them unchanged.  Without OpenCV (this build container has none) the same
// code uses the PODs and a minimal row-major float matrix.
//
// Out of scope here, as in the library: TensorRT engine loading / inference (NN:43-186), image
// preprocessing (BASE:68-121), the geometry back end (BASE:125-399).
#ifndef SPVO_FRONTEND_HPP_
#define SPVO_FRONTEND_HPP_

#include <array>
#include <cstring>
#include <deque>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "spvo_frontend.h"

#if defined(__has_include)
#if __has_include(<opencv2/core.hpp>) && !defined(SPVO_NO_OPENCV)
#include <opencv2/core.hpp>
#define SPVO_HAVE_OPENCV 1
#endif
#endif

namespace spvo {

#ifdef SPVO_HAVE_OPENCV
using KeyPoint = cv::KeyPoint;
using DMatch = cv::DMatch;
using DescMat = cv::Mat;  // rows x 256, CV_32FC1, continuous
static_assert(sizeof(cv::KeyPoint) == sizeof(spvo_keypoint), "cv::KeyPoint layout");
static_assert(sizeof(cv::DMatch) == sizeof(spvo_dmatch), "cv::DMatch layout");
inline DescMat make_desc(int rows) { return cv::Mat(rows, SPVO_DESC_DIM, CV_32FC1); }
inline float* desc_ptr(DescMat& m) { return m.ptr<float>(); }
inline const float* desc_ptr(const DescMat& m) { return m.ptr<float>(); }
inline int desc_rows(const DescMat& m) { return m.rows; }
#else
using KeyPoint = spvo_keypoint;
using DMatch = spvo_dmatch;
struct DescMat {  // stand-in for the cv::Mat(K, 256, CV_32FC1) of NN:347
  int rows = 0, cols = SPVO_DESC_DIM;
  std::vector<float> data;
};
inline DescMat make_desc(int rows) {
  DescMat m;
  m.rows = rows;
  m.data.resize((size_t)rows * SPVO_DESC_DIM);
  return m;
}
inline float* desc_ptr(DescMat& m) { return m.data.data(); }
inline const float* desc_ptr(const DescMat& m) { return m.data.data(); }
inline int desc_rows(const DescMat& m) { return m.rows; }
#endif
static_assert(sizeof(spvo_keypoint) == 28 && sizeof(spvo_dmatch) == 16, "POD layouts");

// HPP:55-90
enum class MatcherType { BF, FLANN };
enum class SelectorType { NN, KNN };
enum ImagePosition { PREV_LEFT = -4, PREV_RIGHT = -3, CURR_LEFT = -2, CURR_RIGHT = -1, NUM_IMAGE_POSITIONS = 4 };
enum MatchType { CURR_LEFT_CURR_RIGHT = 0, CURR_LEFT_PREV_LEFT = 1, PREV_LEFT_PREV_RIGHT = 2, MATCH_TYPE_NUM = 3 };
static const std::array<std::pair<int, int>, MATCH_TYPE_NUM> match_type_to_positions = {
    std::pair<int, int>(CURR_LEFT, CURR_RIGHT), std::pair<int, int>(CURR_LEFT, PREV_LEFT),
    std::pair<int, int>(PREV_LEFT, PREV_RIGHT)};

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// The abstract front end's matching half (HPP:96-178, BASE:10-33, 434-500).
class FeatureFrontEnd {
 public:
  FeatureFrontEnd(MatcherType matcher_type, SelectorType selector_type, bool cross_check, float stereo_threshold,
                  float min_disparity, int input_height, int input_width)
      : matcher_type_(matcher_type), selector_type_(selector_type), cross_check_(cross_check),
        stereo_threshold_(stereo_threshold), min_disparity_(min_disparity), input_height_(input_height),
        input_width_(input_width) {
    if (matcher_type_ != MatcherType::BF)
      throw Error(SPVO_EINVAL, "only the brute-force matcher is provided (FLANN, BASE:29-32, is out of scope)");
  }
  virtual ~FeatureFrontEnd() {
    if (handle_) spvo_destroy(handle_);
  }
  FeatureFrontEnd(const FeatureFrontEnd&) = delete;
  FeatureFrontEnd& operator=(const FeatureFrontEnd&) = delete;

  // BASE:35-66 (sic)
  void clearLagecyData() {
    keypoints_dq.clear();
    descriptors_dq.clear();
    for (auto& m : cv_DMatches_list) m.clear();
    for (auto& m : maps_of_indices) m.clear();
  }

  // BASE:434-500
  void matchDescriptors(const MatchType match_type) {
    const DescMat& descriptors0 = descriptors_dq.end()[match_type_to_positions[match_type].first];
    const DescMat& descriptors1 = descriptors_dq.end()[match_type_to_positions[match_type].second];
    const std::vector<KeyPoint>& keypoints0 = keypoints_dq.end()[match_type_to_positions[match_type].first];
    std::vector<DMatch>& cv_Dmatches = cv_DMatches_list[match_type];
    const int N = desc_rows(descriptors0), M = desc_rows(descriptors1);
    spvo_match_cfg cfg;
    // initMatcher (BASE:27-28): BFMatcher(NORM_L2, cross_check_ && selector_type_ != KNN)
    cfg.mode = selector_type_ == SelectorType::KNN ? SPVO_MATCH_KNN_RATIO
                                                   : (cross_check_ ? SPVO_MATCH_NN_CROSSCHECK : SPVO_MATCH_NN);
    cfg.ratio = knn_threshold_;
    cfg.algorithm = SPVO_MATCHER_AUTO;
    cfg.flags = 0;
    cv_Dmatches.assign((size_t)(N > 0 ? N : 1), DMatch());
    std::vector<int> q2t((size_t)(N > 0 ? N : 1), -1);
    int n = 0;
    check(spvo_match(handle_, desc_ptr(descriptors0), N, desc_ptr(descriptors1), M, SPVO_DESC_DIM, &cfg,
                     reinterpret_cast<spvo_dmatch*>(cv_Dmatches.data()), &n, q2t.data()));
    cv_Dmatches.resize((size_t)n);
    if (match_type == CURR_LEFT_CURR_RIGHT)  // BASE:475-481
      maps_of_indices[PREV_LEFT_PREV_RIGHT] = maps_of_indices[CURR_LEFT_CURR_RIGHT];
    q2t.resize(keypoints0.size());
    maps_of_indices[match_type] = std::move(q2t);  // BASE:483-491
  }

  const std::vector<int>& mapOfIndices(MatchType t) const { return maps_of_indices[t]; }

  std::deque<std::vector<KeyPoint>> keypoints_dq;                    // HPP:124
  std::deque<DescMat> descriptors_dq;                                 // HPP:128
  std::array<std::vector<DMatch>, MATCH_TYPE_NUM> cv_DMatches_list;   // HPP:129

 protected:
  void check(int rc) const {
    if (rc != SPVO_OK) throw Error(rc, spvo_last_error(handle_));
  }
  const MatcherType matcher_type_;
  const float knn_threshold_ = 0.8f;  // HPP:137
  const SelectorType selector_type_;
  const bool cross_check_;
  const float stereo_threshold_;
  const float min_disparity_;
  const int input_height_;
  const int input_width_;
  std::array<std::vector<int>, MATCH_TYPE_NUM> maps_of_indices;  // HPP:161
  spvo_handle handle_ = nullptr;
};

// The SuperPoint back end's post-network half (HPP:253-391, NN:188-431, 449-510).
class SuperPointFeatureFrontEnd : public FeatureFrontEnd {
 public:
  SuperPointFeatureFrontEnd(MatcherType matcher_type, SelectorType selector_type, bool cross_check,
                            int model_batch_size, int input_height, int input_width, float conf_thresh,
                            int dist_thresh, int border_remove, float stereo_threshold, float min_disparity,
                            int max_keypoints = 1000, int device = 0)
      : FeatureFrontEnd(matcher_type, selector_type, cross_check, stereo_threshold, min_disparity, input_height,
                        input_width),
        model_batch_size_(model_batch_size),
        output_det_size_(model_batch_size * output_det_channel_ * input_height * input_width / 64),
        output_desc_size_(model_batch_size * output_desc_channel_ * input_height * input_width / 64),
        output_width_(input_width / 8), output_height_(input_height / 8), conf_thresh_(conf_thresh),
        dist_thresh_(dist_thresh), border_remove_(border_remove), max_keypoints_(max_keypoints) {
    if (input_height % 8 != 0 || input_width % 8 != 0)  // HPP:296
      throw Error(SPVO_EINVAL, "input_height and input_width must be multiples of 8");
    if (model_batch_size != 1 && model_batch_size != 2)  // NN:489-491
      throw Error(SPVO_EINVAL, "Wrong batch size");
    initPointers();
    int rc = spvo_create(&handle_, device, model_batch_size, input_height, input_width, max_keypoints);
    if (rc != SPVO_OK) throw Error(rc, spvo_last_error(nullptr));
  }

  void initPointers() {  // HPP:309-318 (the Eigen thread pool has no equivalent: the work is on the GPU)
    input_data_ = std::unique_ptr<float[]>(new float[(size_t)model_batch_size_ * input_height_ * input_width_]);
    output_det_data_ = std::unique_ptr<float[]>(new float[output_det_size_]);
    output_desc_data_ = std::unique_ptr<float[]>(new float[output_desc_size_]);
  }

  // NN:139-161 + BASE:68-121.  img: rows x cols 8UC1 with `stride` bytes per row; projection_matrix: 3x4 row-major
  // fp32, patched in place.  Writes slot curr_batch of input_data_ (the block runNeuralNetwork uploads, NN:164-166)
  // and appends the resized 8-bit image to images_dq (NN:153).
  void preprocessImage(const uint8_t* img, int rows, int cols, int stride, float* projection_matrix, int curr_batch) {
    if (curr_batch < 0 || curr_batch >= model_batch_size_) throw Error(SPVO_EINVAL, "curr_batch out of range");
    std::vector<uint8_t> resized((size_t)input_height_ * input_width_);
    check(spvo_preprocess(handle_, img, 1, rows, cols, stride, input_height_, input_width_,
                          input_data_.get() + (size_t)curr_batch * input_height_ * input_width_, resized.data(),
                          projection_matrix));
    images_dq.push_back(std::move(resized));
    while (images_dq.size() > 4) images_dq.pop_front();
  }
#ifdef SPVO_HAVE_OPENCV
  void preprocessImage(cv::Mat& img, cv::Mat& projection_matrix, const int curr_batch) {  // reference signature
    if (img.type() != CV_8UC1 || projection_matrix.type() != CV_32FC1 || !projection_matrix.isContinuous())
      throw Error(SPVO_EINVAL, "preprocessImage: 8UC1 image and continuous CV_32FC1 3x4 projection matrix expected");
    preprocessImage(img.ptr<uint8_t>(), img.rows, img.cols, (int)img.step, projection_matrix.ptr<float>(), curr_batch);
    img = cv::Mat(input_height_, input_width_, CV_8UC1, images_dq.back().data()).clone();
  }
#endif

  // NN:264-364.  Consumes output_det_data_ / output_desc_data_, appends model_batch_size_ entries to
  // keypoints_dq / descriptors_dq and trims the deques to 4 entries (NN:494-498).
  void postprocessDetectionAndDescription() {
    const int B = model_batch_size_, K = max_keypoints_;
    spvo_decode_cfg cfg{conf_thresh_, dist_thresh_, border_remove_, K};
    std::vector<spvo_keypoint> kp((size_t)B * K);
    std::vector<float> desc((size_t)B * K * SPVO_DESC_DIM);
    std::vector<int> n((size_t)B, 0);
    check(spvo_decode(handle_, output_det_data_.get(), output_desc_data_.get(), B, input_height_, input_width_, &cfg,
                      kp.data(), desc.data(), n.data(), nullptr));
    for (int b = 0; b < B; ++b) {
      std::vector<KeyPoint> kps((size_t)n[b]);
      if (n[b] > 0) std::memcpy(static_cast<void*>(kps.data()), kp.data() + (size_t)b * K, (size_t)n[b] * sizeof(spvo_keypoint));
      keypoints_dq.push_back(std::move(kps));  // NN:261
      DescMat d = make_desc(n[b]);
      if (n[b] > 0)
        std::memcpy(desc_ptr(d), desc.data() + (size_t)b * K * SPVO_DESC_DIM, (size_t)n[b] * SPVO_DESC_DIM * sizeof(float));
      descriptors_dq.push_back(std::move(d));  // NN:362
    }
    while (keypoints_dq.size() > 4) {  // NN:494-498
      keypoints_dq.pop_front();
      descriptors_dq.pop_front();
    }
  }

  inline int getInputHeight() const { return input_height_; }
  inline int getInputWidth() const { return input_width_; }

  std::unique_ptr<float[]> input_data_;             // [B,H,W] network input, HPP:382
  std::deque<std::vector<uint8_t>> images_dq;      // resized 8-bit images, HPP:123 (cv::Mat in the reference)
  // host I/O buffers the network output is copied into (NN:170-176)
  std::unique_ptr<float[]> output_det_data_;   // [B,65,H/8,W/8]   HPP:383
  std::unique_ptr<float[]> output_desc_data_;  // [B,256,H/8,W/8]  HPP:384

  static constexpr int output_det_channel_ = SPVO_DET_CHANNELS;     // HPP:355
  static constexpr int output_det_heatmap_factor_ = SPVO_CELL;      // HPP:356
  static constexpr int output_desc_channel_ = SPVO_DESC_DIM;        // HPP:359

 private:
  const int model_batch_size_;
  const int output_det_size_;
  const int output_desc_size_;
  const int output_width_;
  const int output_height_;
  const float conf_thresh_;
  const int dist_thresh_;
  const int border_remove_;
  const int max_keypoints_;  // HPP:368 (compile-time 1000 in the reference)
};

}  // namespace spvo
#endif  // SPVO_FRONTEND_HPP_
