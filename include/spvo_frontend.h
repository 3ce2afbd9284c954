/* spvo_frontend.h -- C ABI of libspvo_frontend.so: B200-native (sm_100a) SuperPoint decode +
 * brute-force L2 descriptor matching, the data-parallel front end of
 * YukunXia/SuperPoint-Stereo-Visual-Odometry.
 *
 * Every entry point states the reference interface it replaces.  Citations are relative to
 * /root/reference/src/odml_visual_odometry/ :
 *   HPP  = include/odml_visual_odometry/feature_detection.hpp
 *   NN   = src/feature_detection_neural_network.cpp
 *   BASE = src/feature_detection_base.cpp
 *
 * Plain C: opaque handle, raw pointers, sizes, int status codes.  No torch / OpenCV / C++ types.
 * There is NO CPU fallback behind this ABI: every compute entry point runs hand-written CUDA
 * kernels and fails with SPVO_ECUDA / SPVO_ENODEVICE when no sm_100 device is usable.
 *
 * Threading (as the reference, BASE/NN hold no locks): a handle is not thread-safe; distinct
 * handles (one per GPU / stream) are fully independent; the library has no global state.
 */
#ifndef SPVO_FRONTEND_H_
#define SPVO_FRONTEND_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPVO_ABI_VERSION 4

#define SPVO_DESC_DIM 256      /* output_desc_channel_, HPP:359 */
#define SPVO_DET_CHANNELS 65   /* output_det_channel_,  HPP:355 */
#define SPVO_CELL 8            /* output_det_heatmap_factor_, HPP:356 */

/* ---- status codes (the reference logs ROS_ERROR and returns, NN:53-55/489-491; OpenCV throws) ---- */
enum {
  SPVO_OK = 0,
  SPVO_EINVAL = 1,     /* bad argument (null pointer, H or W not a multiple of 8 as HPP:296, size above handle capacity) */
  SPVO_ECUDA = 2,      /* a CUDA runtime call or kernel failed; see spvo_last_error() */
  SPVO_ENODEVICE = 3,  /* no usable sm_100 device */
  SPVO_ENOMEM = 4      /* workspace allocation failed */
};

/* Layout-compatible with cv::KeyPoint (28 bytes).  The reference emits
 * cv::KeyPoint(pt=(x,y), size=1) => angle=-1, response=0, octave=0, class_id=-1 (NN:243). */
typedef struct spvo_keypoint {
  float x, y;
  float size;
  float angle;
  float response;
  int32_t octave;
  int32_t class_id;
} spvo_keypoint;

/* Layout-compatible with cv::DMatch (16 bytes): the element type of cv_DMatches_list (HPP:129). */
typedef struct spvo_dmatch {
  int32_t queryIdx;
  int32_t trainIdx;
  int32_t imgIdx; /* always 0, as cv::BFMatcher::match with one train image */
  float distance; /* sqrt(sum (a-b)^2) with cv::hal::normL2Sqr_'s fp32 operation order */
} spvo_dmatch;

/* Decode parameters: the SuperPointFeatureFrontEnd members conf_thresh_, dist_thresh_,
 * border_remove_ (HPP:365-367, launch defaults 0.015 / 4 / 4) and max_keypoints_ (HPP:368, the
 * compile-time 1000 made a runtime value). */
typedef struct spvo_decode_cfg {
  float conf_thresh;
  int32_t dist_thresh;
  int32_t border_remove;
  int32_t max_keypoints;
} spvo_decode_cfg;

/* Matching modes: SelectorType x cross_check_ as combined in initMatcher (BASE:27-28). */
enum {
  SPVO_MATCH_NN = 0,            /* selector NN, crossCheck=false : BFMatcher::match               (BASE:463) */
  SPVO_MATCH_NN_CROSSCHECK = 1, /* selector NN, crossCheck=true  : mutual first-argmin            (BASE:463) */
  SPVO_MATCH_KNN_RATIO = 2      /* selector KNN: knnMatch(k=2), keep m0 iff m0.d < ratio * m1.d   (BASE:466-472) */
};

/* Matcher algorithm inside the library (results are identical; the exact kernel is the anchor). */
enum {
  SPVO_MATCHER_AUTO = 0,
  SPVO_MATCHER_EXACT_FP32 = 1, /* CUDA-core kernel computing every distance in OpenCV's fp32 order */
  SPVO_MATCHER_TENSOR = 2      /* tcgen05/TMEM 16-bit GEMM shortlist (bf16; fp16 for the unit-norm descriptors
                                  decode produces) + exact fp32 re-rank; identical results */
};

/* Flags (spvo_match_cfg.flags).  ROW_BAND = the BASELINE north star's "left<->right matching under a stereo row-band
 * constraint" as an explicit opt-in: train row j is a candidate of query i only if |y_i - y_j| <= band.  The reference
 * matches UNMASKED and applies the band afterwards (BASE:169-172), and cv::BFMatcher refuses crossCheck with a mask, so
 * the masked cross-check is defined as two masked cv::BFMatcher(NORM_L2, false)::match calls (mask, transposed mask)
 * plus a mutual test -- what the oracle is pinned to.  A query without an allowed train row has no match; KNN_RATIO
 * needs two allowed rows.  In spvo_stereo_batch* the flag masks the L<->R problems with band = stereo_threshold (the
 * temporal problems stay unmasked); for single problems use spvo_match_masked[_device]. */
#define SPVO_MATCH_FLAG_ROW_BAND 1

typedef struct spvo_match_cfg {
  int32_t mode;      /* SPVO_MATCH_* */
  float ratio;       /* knn_threshold_ = 0.8f (HPP:137); used by SPVO_MATCH_KNN_RATIO */
  int32_t algorithm; /* SPVO_MATCHER_* */
  int32_t flags;     /* SPVO_MATCH_FLAG_* (0 = the reference's behaviour) */
} spvo_match_cfg;

typedef struct spvo_handle_s* spvo_handle;

/* ---- lifecycle: replaces the SuperPointFeatureFrontEnd ctor's initPointers()/initMatcher()
 * (HPP:302-318, BASE:10-33) and the dtor (NN:24-41).  The library owns its workspaces; the caller
 * owns every input and output buffer. ---- */
int spvo_create(spvo_handle* out, int device, int max_batch, int max_height, int max_width,
                int max_keypoints);
int spvo_destroy(spvo_handle h);
const char* spvo_last_error(spvo_handle h); /* h may be NULL: returns the last create error */
int spvo_abi_version(void);

/* Stream the handle's *_device entry points enqueue on (a cudaStream_t; NULL = the handle's own
 * stream created at spvo_create, the analogue of NN:134's stream_). */
int spvo_set_stream(spvo_handle h, void* cuda_stream);
int spvo_sync(spvo_handle h);

/* ---- preprocess (the step before the network): replaces FeatureFrontEnd::preprocessImageImpl (BASE:68-121) and
 * the 8U -> 32F hand-over of SuperPointFeatureFrontEnd::preprocessImage (NN:139-161): centre crop to the aspect
 * ratio W:H, cv::resize(INTER_LINEAR) of the 8UC1 image to W x H (OpenCV's fixed-point arithmetic, bit for bit),
 * input = pixel * (1/255).
 *   imgs        [B, rows, stride] 8UC1, stride >= cols bytes per row (a batch of equally sized camera images)
 *   input_out   [B, H, W] fp32   = input_data_ (HPP:382), the network's input block; may be NULL
 *   resized_out [B, H, W] u8     = the image the reference keeps in images_dq (NN:153); may be NULL
 *   proj        [B][12] optional, HOST memory in both forms: 3x4 row-major projection matrices, patched in place
 *               (principal point minus the crop offset, rows 0-1 times W / cropped_cols: BASE:93, 109, 119-120)
 * H and W need not be multiples of 8 here.  Host-pointer form copies in, runs, copies out, synchronises. */
int spvo_preprocess(spvo_handle h, const uint8_t* imgs, int B, int rows, int cols, int stride, int H, int W,
                    float* input_out, uint8_t* resized_out, float* proj);
int spvo_preprocess_device(spvo_handle h, const uint8_t* imgs, int B, int rows, int cols, int stride, int H, int W,
                           float* input_out, uint8_t* resized_out, float* proj);

/* ---- decode: replaces SuperPointFeatureFrontEnd::postprocessDetectionAndDescription()
 * (HPP:327, NN:264-364) including processOneHeatmap (NN:188-262) and bilinearInterpolationDesc
 * (NN:366-431).
 *   semi  [B,65,H/8,W/8]  fp32 NCHW  = output_det_data_  (HPP:383)
 *   desc  [B,256,H/8,W/8] fp32 NCHW  = output_desc_data_ (HPP:384)
 *   kpts_out [B,K] (K = cfg->max_keypoints), rows beyond n_out[b] are zero-filled
 *   desc_out [B,K,256] fp32 row-major = the cv::Mat(K,256,CV_32FC1) of NN:347-362
 *   n_out    [B]  keypoints per image (= keypoints_dq entry sizes)
 *   scores_out [B,K] optional (may be NULL): heatmap value of each keypoint (the reference drops it)
 * Host-pointer form: pageable or pinned host memory; copies in, runs, copies out, synchronises. */
int spvo_decode(spvo_handle h, const float* semi, const float* desc, int B, int H, int W,
                const spvo_decode_cfg* cfg, spvo_keypoint* kpts_out, float* desc_out, int* n_out,
                float* scores_out);
/* Device-pointer form: all pointers are device memory; asynchronous on the handle's stream. */
int spvo_decode_device(spvo_handle h, const float* semi, const float* desc, int B, int H, int W,
                       const spvo_decode_cfg* cfg, spvo_keypoint* kpts_out, float* desc_out,
                       int* n_out, float* scores_out);

/* ---- match: replaces FeatureFrontEnd::matchDescriptors(MatchType) (HPP:118, BASE:434-500),
 * i.e. cv::BFMatcher(NORM_L2)::match / knnMatch + ratio test, and the index map of BASE:483-491.
 *   q [N,dim], t [M,dim] fp32 row-major (CV_32F continuous cv::Mat); dim must be 256
 *   out [>=N] DMatch list in ascending queryIdx; *n_matches entries are valid
 *   q2t [N] optional: trainIdx per query or -1  (maps_of_indices[type], HPP:161)
 * Defined edge cases (reference: cv::Exception / UB at BASE:469): N==0 or M==0 -> 0 matches;
 * KNN_RATIO with M<2 -> 0 matches. */
int spvo_match(spvo_handle h, const float* q, int N, const float* t, int M, int dim,
               const spvo_match_cfg* cfg, spvo_dmatch* out, int* n_matches, int* q2t);
/* Device-pointer form; n_matches is a device int; asynchronous. */
int spvo_match_device(spvo_handle h, const float* q, int N, const float* t, int M, int dim,
                      const spvo_match_cfg* cfg, spvo_dmatch* out, int* n_matches, int* q2t);

/* Row-band masked matching of one problem (see SPVO_MATCH_FLAG_ROW_BAND): q_kpts [N], t_kpts [M] are the keypoints the
 * descriptors belong to (only .y is read); band >= 0.  Same outputs as spvo_match. */
int spvo_match_masked(spvo_handle h, const float* q, int N, const float* t, int M, int dim,
                      const spvo_match_cfg* cfg, const spvo_keypoint* q_kpts, const spvo_keypoint* t_kpts, float band,
                      spvo_dmatch* out, int* n_matches, int* q2t);
int spvo_match_masked_device(spvo_handle h, const float* q, int N, const float* t, int M, int dim,
                             const spvo_match_cfg* cfg, const spvo_keypoint* q_kpts, const spvo_keypoint* t_kpts,
                             float band, spvo_dmatch* out, int* n_matches, int* q2t);

/* Batched device form: P independent problems per launch (frames are independent in the front
 * end, so a stream of stereo pairs becomes grouped launches).  Problem p matches
 *   q = desc_base + q_slot[p]*slot_stride   (n_rows[q_slot[p]] rows)   against
 *   t = desc_base + t_slot[p]*slot_stride   (n_rows[t_slot[p]] rows)
 * where a "slot" is one image's [K,256] descriptor block as written by spvo_decode_device and
 * n_rows is its n_out array (read on the device: no host round trip between decode and match).
 *   out [P, max_rows] DMatch, n_matches [P], q2t [P, max_rows]. */
int spvo_match_batch_device(spvo_handle h, const float* desc_base, const int* n_rows, int slot_stride_rows,
                            const int* q_slot, const int* t_slot, int P, int max_rows, int dim,
                            const spvo_match_cfg* cfg, spvo_dmatch* out, int* n_matches, int* q2t);

/* ---- stereo row-band / min-disparity test applied to L<->R matches (BASE:169-172):
 * keep[m] = !(|y_l - y_r| > stereo_threshold || |x_l - x_r| < min_disparity).
 * Device form, batched like spvo_match_batch_device (kpts_base slots of slot_stride_rows rows);
 * q_slot / t_slot may both be NULL, meaning problem p uses slots (2p, 2p+1). */
int spvo_stereo_filter_batch_device(spvo_handle h, const spvo_keypoint* kpts_base, int slot_stride_rows,
                                    const int* q_slot, const int* t_slot, int P, int max_rows,
                                    const spvo_dmatch* matches, const int* n_matches,
                                    float stereo_threshold, float min_disparity, uint8_t* keep);

/* ---- stereo stream: one call per batch of F consecutive stereo frames.  Replaces, for every frame f,
 * the front-end part of stereoCallback (visual_odometry_node.cpp:175-199):
 *   addStereoImagePair -> postprocessDetectionAndDescription on (left_f, right_f)      (NN:468-484)
 *   matchDescriptors(CURR_LEFT_CURR_RIGHT)   query = left_f, train = right_f           (HPP:88)
 *   matchDescriptors(CURR_LEFT_PREV_LEFT)    query = left_f, train = left_{f-1}        (HPP:89)
 *   the stereo row-band / min-disparity test of solveStereoOdometry on the L<->R matches (BASE:169-172)
 * Frames are independent in the front end except for the one-frame dependency of the temporal
 * match; the handle keeps the previous batch's last left image (descriptors, keypoints, count) on
 * the device, so consecutive calls continue one sequence.  spvo_stereo_reset() forgets it
 * (clearLagecyData, BASE:35-66): the next frame then has no temporal matches, as the reference's
 * first frame (visual_odometry_node.cpp:188-193).
 *   semi [F,2,65,H/8,W/8], desc [F,2,256,H/8,W/8]: image index 2f + eye, eye 0 = left (the batch-2
 *   layout of NN:480-484).  2F <= max_batch of the handle.  K = cfg->decode.max_keypoints. */
typedef struct spvo_quad { /* indices into the four keypoint lists of BASE:163-196 */
  int32_t curr_left, curr_right, prev_left, prev_right;
} spvo_quad;

typedef struct spvo_stereo_cfg {
  spvo_decode_cfg decode;
  spvo_match_cfg match;
  float stereo_threshold; /* stereo_threshold_, launch default 2.0  */
  float min_disparity;    /* min_disparity_,   launch default 0.25 */
} spvo_stereo_cfg;

typedef struct spvo_stereo_out { /* all device pointers (_device form) or all host pointers */
  spvo_keypoint* kpts;   /* [2F, K]                                                              */
  float* desc;           /* [2F, K, 256]; host form: may be NULL (descriptors stay on the device) */
  int* n_kpts;           /* [2F]                                                                  */
  spvo_dmatch* matches;  /* [2F, K]: rows 0..F-1 stereo L_f->R_f, rows F..2F-1 temporal L_f->L_{f-1} */
  int* n_matches;        /* [2F]                                                                  */
  int* q2t;              /* [2F, K] maps_of_indices (HPP:161), -1 = unmatched                     */
  uint8_t* stereo_keep;  /* [F, K]  1 = L<->R match m passes BASE:169-172                         */
  /* optional (may be NULL; need q2t and stereo_keep): per frame, the keypoint index quadruples that
   * solveStereoOdometry triangulates (BASE:156-207): walk the L<->R matches in order, keep a match iff the
   * current left keypoint also has a temporal match, the match passes the stereo test, and the matched
   * previous left keypoint had a stereo match in the previous frame. */
  struct spvo_quad* quads; /* [F, K] */
  int* n_quads;            /* [F]    */
} spvo_stereo_out;

int spvo_stereo_reset(spvo_handle h);
/* Graph mode for spvo_stereo_batch_device[_f16] (off by default): the launches of a call are captured into a CUDA
 * graph owned by the handle, once per call signature (pointers, sizes, configuration, stream, and the handle's stream
 * state), and replayed on later calls with the same signature; a change re-captures.  Meant for the reference's
 * real-time shape -- one stereo pair per callback with fixed engine output bindings (visual_odometry_node.cpp:150-262)
 * -- where the dozen dependent launches of one pair are latency-, not throughput-bound.  Results are identical. */
int spvo_set_graph_mode(spvo_handle h, int on);
int spvo_stereo_batch_device(spvo_handle h, const float* semi, const float* desc, int F, int H, int W,
                             const spvo_stereo_cfg* cfg, const spvo_stereo_out* out);
/* Host-pointer form: H2D of the inputs, the device pipeline, D2H of every non-NULL output, sync.
 * Pinned host memory gives full PCIe rate. */
int spvo_stereo_batch(spvo_handle h, const float* semi, const float* desc, int F, int H, int W,
                      const spvo_stereo_cfg* cfg, const spvo_stereo_out* out);

/* ---- fp16 network outputs (staging straight from an fp16 inference engine).  Same functions, same
 * results contract, but semi / desc hold IEEE binary16 elements in the same NCHW layout.  fp16 -> fp32 is exact, so
 * the outputs are bit-identical to the fp32 entry points fed with the widened tensors; decode's HBM traffic (and the
 * host forms' H2D) halves. ---- */
int spvo_decode_f16(spvo_handle h, const void* semi_f16, const void* desc_f16, int B, int H, int W,
                    const spvo_decode_cfg* cfg, spvo_keypoint* kpts_out, float* desc_out, int* n_out,
                    float* scores_out);
int spvo_decode_device_f16(spvo_handle h, const void* semi_f16, const void* desc_f16, int B, int H, int W,
                           const spvo_decode_cfg* cfg, spvo_keypoint* kpts_out, float* desc_out, int* n_out,
                           float* scores_out);
int spvo_stereo_batch_f16(spvo_handle h, const void* semi_f16, const void* desc_f16, int F, int H, int W,
                          const spvo_stereo_cfg* cfg, const spvo_stereo_out* out);
int spvo_stereo_batch_device_f16(spvo_handle h, const void* semi_f16, const void* desc_f16, int F, int H, int W,
                                 const spvo_stereo_cfg* cfg, const spvo_stereo_out* out);

/* ---- introspection for tests / bench ---- */
/* Number of kernels this handle has launched since creation (bench.py's gpu_launches). */
long long spvo_kernel_launches(spvo_handle h);
/* Counters of the last decode / match (device work must be complete: call spvo_sync first):
 * [0] images whose decode needed more than the first candidate chunk (exact multi-chunk path),
 * [1] matcher rows sent to the fp32 full-row fallback, [2] rows that needed an exact scan of every column.
 * Cumulative since spvo_create. */
int spvo_debug_counters(spvo_handle h, long long* out, int n);

/* Self-check of k_softmax_heat's shared-reciprocal division (tests only): a_bits / b_bits are DEVICE arrays of n fp32
 * bit patterns; *mismatches = operand pairs (inside the kernel's guarded domain) whose quotient differs from the IEEE
 * division the reference performs (NN:280-284).  Must be 0. */
int spvo_debug_div_check(spvo_handle h, const uint32_t* a_bits, const uint32_t* b_bits, long long n,
                         long long* mismatches);

/* Optional per-kernel profile: when enabled every kernel launch is bracketed by CUDA events on
 * the handle's stream (how bench.py measures the dominant kernel's duration live, inside its timed
 * region).  spvo_profile_read synchronises the stream, returns the accumulated milliseconds and
 * launch counts per kernel class since the last read and clears them.  Kernel classes are indexed
 * 0 .. spvo_profile_num_kernels()-1; spvo_profile_kernel_name gives the kernel's name. */
int spvo_profile_enable(spvo_handle h, int on);
int spvo_profile_num_kernels(void);
const char* spvo_profile_kernel_name(int kernel_class);
int spvo_profile_read(spvo_handle h, double* ms, long long* launches, int n);

#ifdef __cplusplus
}
#endif
#endif /* SPVO_FRONTEND_H_ */
