// api.cu -- the C ABI of libspvo_frontend.so (see include/spvo_frontend.h for the reference
// interface each entry point replaces).  Host C++ only: argument validation, workspace ownership,
// stream plumbing, host<->device staging for the host-pointer entry points.  No CPU compute path.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>

#include "common.cuh"

namespace spvo {
size_t decode_smem_required(int H, int W, int K);
size_t decode_list_bytes_per_image();
}

using namespace spvo;

static char g_create_err[512] = "";

static int fail(Handle* h, int code, const char* fmt, ...) {
  char* dst = h ? h->err : g_create_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

static int cuda_fail(Handle* h, cudaError_t e, const char* what) {
  return fail(h, SPVO_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

#define CK(call)                                                   \
  do {                                                             \
    cudaError_t e_ = (call);                                       \
    if (e_ != cudaSuccess) return cuda_fail(h, e_, #call);         \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

extern "C" {

int spvo_abi_version(void) { return SPVO_ABI_VERSION; }

const char* spvo_last_error(spvo_handle hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  return h ? h->err : g_create_err;
}

int spvo_create(spvo_handle* out, int device, int max_batch, int max_height, int max_width, int max_keypoints) {
  Handle* h = nullptr;
  if (!out) return fail(nullptr, SPVO_EINVAL, "spvo_create: out is NULL");
  *out = nullptr;
  if (max_batch <= 0 || max_height <= 0 || max_width <= 0 || max_keypoints < 0 || max_height % 8 || max_width % 8)
    return fail(nullptr, SPVO_EINVAL, "spvo_create: bad capacity (batch %d, %dx%d, K %d; H and W must be multiples of 8)",
                max_batch, max_height, max_width, max_keypoints);
  if (max_keypoints > 4096) return fail(nullptr, SPVO_EINVAL, "spvo_create: max_keypoints %d > 4096", max_keypoints);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, SPVO_ENODEVICE, "spvo_create: no CUDA device (%s)", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(nullptr, SPVO_EINVAL, "spvo_create: device %d of %d", device, ndev);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(nullptr, SPVO_ENODEVICE, "spvo_create: device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
  const size_t smem = decode_smem_required(max_height, max_width, max_keypoints > 0 ? max_keypoints : 1);
  if (smem > (size_t)prop.sharedMemPerBlockOptin)
    return fail(nullptr, SPVO_EINVAL, "spvo_create: %dx%d with K=%d needs %zu B of shared memory (> %zu)", max_height,
                max_width, max_keypoints, smem, (size_t)prop.sharedMemPerBlockOptin);
  h = new (std::nothrow) Handle();
  if (!h) return fail(nullptr, SPVO_ENOMEM, "spvo_create: out of host memory");
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->max_batch = max_batch; h->max_h = max_height; h->max_w = max_width; h->max_k = max_keypoints;
  DeviceGuard g(device);
  const size_t px = (size_t)max_height * max_width, cells = px / 64;
#define ALLOC(ptr, bytes)                                                         \
  if ((e = cudaMalloc((void**)&(ptr), (bytes))) != cudaSuccess) {                  \
    int rc = fail(nullptr, SPVO_ENOMEM, "spvo_create: cudaMalloc(%zu): %s", (size_t)(bytes), cudaGetErrorString(e)); \
    spvo_destroy(reinterpret_cast<spvo_handle>(h));                                \
    return rc;                                                                     \
  }
  if ((e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking)) != cudaSuccess) {
    int rc = cuda_fail(nullptr, e, "cudaStreamCreate");
    delete h;
    return rc;
  }
  h->stream = h->own_stream;
  ALLOC(h->heat, (size_t)max_batch * px * sizeof(float));
  ALLOC(h->spill_thr, (size_t)max_batch * sizeof(uint32_t));
  cudaMemset(h->spill_thr, 0x7F, (size_t)max_batch * sizeof(uint32_t));  // no history yet: nothing is stored
  ALLOC(h->cand_list, (size_t)max_batch * decode_list_bytes_per_image());
  ALLOC(h->cellmax, (size_t)max_batch * cells * sizeof(uint2));
  ALLOC(h->nms_bitmap, (size_t)max_batch * (px / 16 + 64) * sizeof(unsigned));  // H*ceil(W/32) <= H*W/16 for W >= 16
  ALLOC(h->counters, 8 * sizeof(unsigned long long));
  cudaMemset(h->counters, 0, 8 * sizeof(unsigned long long));
  const size_t K = max_keypoints > 0 ? max_keypoints : 1;
  ALLOC(h->st_semi, (size_t)max_batch * 65 * cells * sizeof(float));
  ALLOC(h->st_desc, (size_t)max_batch * 256 * cells * sizeof(float));
  ALLOC(h->st_kpts, (size_t)max_batch * K * sizeof(spvo_keypoint));
  ALLOC(h->st_desc_out, (size_t)max_batch * K * 256 * sizeof(float));
  ALLOC(h->st_n, (size_t)max_batch * sizeof(int));
  ALLOC(h->st_scores, (size_t)max_batch * K * sizeof(float));
  ALLOC(h->desc_tmp, (size_t)max_batch * 256 * ((K + 3) & ~3) * sizeof(float));
  ALLOC(h->kp_par, (size_t)max_batch * K * sizeof(int4));
  ALLOC(h->probs, 4096 * sizeof(MatchProblem));
  h->probs_cap = 4096;
#undef ALLOC
  *out = reinterpret_cast<spvo_handle>(h);
  return SPVO_OK;
}

int spvo_destroy(spvo_handle hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return SPVO_OK;
  DeviceGuard g(h->device);
  if (h->own_stream) cudaStreamSynchronize(h->own_stream);
  tc_workspace_free(h);
  void* ptrs[] = {h->pp_src, h->pp_dst_f, h->pp_dst_u8, h->pp_tab, h->nms_bitmap, h->cellmax, h->desc_tmp, h->kp_par, h->heat, h->spill_thr, h->cand_list, h->counters, h->st_semi, h->st_desc, h->st_kpts, h->st_desc_out, h->st_n,
                  h->st_scores, h->dist, h->row_best, h->row_d, h->col_best, h->probs, h->st_q, h->st_t,
                  h->st_matches, h->st_q2t, h->st_nm, h->st_mkp, h->carry_desc, h->carry_kpts, h->carry_n, h->carry_map, h->st_quads, h->st_nquads,
                  h->st_smatches, h->st_snm, h->st_sq2t, h->st_skeep};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (const ProfRec& r : h->prof_recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  for (Handle::GraphEntry& ge : h->graphs)
    if (ge.exec) cudaGraphExecDestroy(ge.exec);
  for (int i = 0; i < 2; ++i) {
    if (h->aux_stream[i]) cudaStreamDestroy(h->aux_stream[i]);
    if (h->aux_done[i]) cudaEventDestroy(h->aux_done[i]);
  }
  if (h->aux_fork) cudaEventDestroy(h->aux_fork);
  if (h->pp_done) cudaEventDestroy(h->pp_done);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (int i = 0; i < 5; ++i)
    if (h->copy_ev[i]) cudaEventDestroy(h->copy_ev[i]);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  delete h;
  return SPVO_OK;
}

int spvo_set_stream(spvo_handle hh, void* cuda_stream) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return SPVO_EINVAL;
  h->stream = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : h->own_stream;
  return SPVO_OK;
}

int spvo_sync(spvo_handle hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return SPVO_EINVAL;
  DeviceGuard g(h->device);
  CK(cudaStreamSynchronize(h->stream));
  return SPVO_OK;
}

static int check_decode_args(Handle* h, const void* semi, int B, int H, int W, const spvo_decode_cfg* cfg,
                             spvo_keypoint* kpts, int* n_out) {
  if (!h) return SPVO_EINVAL;
  if (!semi || !cfg || !kpts || !n_out) return fail(h, SPVO_EINVAL, "decode: NULL argument");
  if (B < 0 || B > h->max_batch) return fail(h, SPVO_EINVAL, "decode: batch %d outside [0, %d]", B, h->max_batch);
  if (H <= 0 || W <= 0 || H % 8 || W % 8)
    return fail(h, SPVO_EINVAL, "decode: %dx%d is not a positive multiple of 8 (reference: feature_detection.hpp:296)", H, W);
  if ((size_t)H * W > (size_t)h->max_h * h->max_w)
    return fail(h, SPVO_EINVAL, "decode: %dx%d exceeds the handle capacity %dx%d", H, W, h->max_h, h->max_w);
  if (H < 16 || W < 16) return fail(h, SPVO_EINVAL, "decode: image smaller than 16x16");
  if (cfg->max_keypoints < 0 || cfg->max_keypoints > h->max_k)
    return fail(h, SPVO_EINVAL, "decode: max_keypoints %d outside [0, %d]", cfg->max_keypoints, h->max_k);
  if (cfg->dist_thresh < 0 || cfg->dist_thresh > 64 || cfg->border_remove < 0 || !(cfg->conf_thresh >= 0.0f))
    return fail(h, SPVO_EINVAL, "decode: bad cfg (conf %g, dist %d, border %d)", (double)cfg->conf_thresh,
                cfg->dist_thresh, cfg->border_remove);
  return SPVO_OK;
}

static int decode_device_impl(Handle* h, const void* semi, const void* desc, int in_f16, int B, int H, int W,
                              const spvo_decode_cfg* cfg, spvo_keypoint* kpts_out, float* desc_out, int* n_out,
                              float* scores_out) {
  int rc = check_decode_args(h, semi, B, H, W, cfg, kpts_out, n_out);
  if (rc) return rc;
  DeviceGuard g(h->device);
  h->chain_launches = B <= 8;
  CK(launch_decode(h, semi, desc, in_f16, B, H, W, *cfg, kpts_out, desc_out, n_out, scores_out));
  return SPVO_OK;
}

int spvo_decode_device(spvo_handle hh, const float* semi, const float* desc, int B, int H, int W,
                       const spvo_decode_cfg* cfg, spvo_keypoint* kpts_out, float* desc_out, int* n_out,
                       float* scores_out) {
  return decode_device_impl(reinterpret_cast<Handle*>(hh), semi, desc, 0, B, H, W, cfg, kpts_out, desc_out, n_out,
                            scores_out);
}

int spvo_decode_device_f16(spvo_handle hh, const void* semi, const void* desc, int B, int H, int W,
                           const spvo_decode_cfg* cfg, spvo_keypoint* kpts_out, float* desc_out, int* n_out,
                           float* scores_out) {
  return decode_device_impl(reinterpret_cast<Handle*>(hh), semi, desc, 1, B, H, W, cfg, kpts_out, desc_out, n_out,
                            scores_out);
}

static int decode_host_impl(Handle* h, const void* semi, const void* desc, int in_f16, int B, int H, int W,
                            const spvo_decode_cfg* cfg, spvo_keypoint* kpts_out, float* desc_out, int* n_out,
                            float* scores_out) {
  int rc = check_decode_args(h, semi, B, H, W, cfg, kpts_out, n_out);
  if (rc) return rc;
  if (B == 0) return SPVO_OK;
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  const size_t cells = (size_t)(H / 8) * (W / 8), K = cfg->max_keypoints, esz = in_f16 ? 2 : 4;
  const bool with_desc = desc && desc_out;
  CK(cudaMemcpyAsync(h->st_semi, semi, (size_t)B * 65 * cells * esz, cudaMemcpyHostToDevice, st));
  if (with_desc) CK(cudaMemcpyAsync(h->st_desc, desc, (size_t)B * 256 * cells * esz, cudaMemcpyHostToDevice, st));
  h->chain_launches = B <= 8;
  CK(launch_decode(h, h->st_semi, with_desc ? h->st_desc : nullptr, in_f16, B, H, W, *cfg, h->st_kpts,
                   with_desc ? h->st_desc_out : nullptr, h->st_n, scores_out ? h->st_scores : nullptr));
  CK(cudaMemcpyAsync(n_out, h->st_n, (size_t)B * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (K > 0) {
    CK(cudaMemcpyAsync(kpts_out, h->st_kpts, (size_t)B * K * sizeof(spvo_keypoint), cudaMemcpyDeviceToHost, st));
    if (with_desc)
      CK(cudaMemcpyAsync(desc_out, h->st_desc_out, (size_t)B * K * 256 * sizeof(float), cudaMemcpyDeviceToHost, st));
    if (scores_out)
      CK(cudaMemcpyAsync(scores_out, h->st_scores, (size_t)B * K * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  return SPVO_OK;
}

int spvo_decode(spvo_handle hh, const float* semi, const float* desc, int B, int H, int W,
                const spvo_decode_cfg* cfg, spvo_keypoint* kpts_out, float* desc_out, int* n_out,
                float* scores_out) {
  return decode_host_impl(reinterpret_cast<Handle*>(hh), semi, desc, 0, B, H, W, cfg, kpts_out, desc_out, n_out,
                          scores_out);
}

int spvo_decode_f16(spvo_handle hh, const void* semi, const void* desc, int B, int H, int W,
                    const spvo_decode_cfg* cfg, spvo_keypoint* kpts_out, float* desc_out, int* n_out,
                    float* scores_out) {
  return decode_host_impl(reinterpret_cast<Handle*>(hh), semi, desc, 1, B, H, W, cfg, kpts_out, desc_out, n_out,
                          scores_out);
}

// ---- preprocess (BASE:68-121, NN:139-161) ----
static int check_preprocess_args(Handle* h, const uint8_t* imgs, int B, int rows, int cols, int stride, int H, int W,
                                 const void* input_out, const void* resized_out) {
  if (!h) return SPVO_EINVAL;
  if (B < 0 || rows <= 0 || cols <= 0 || stride < cols || H <= 0 || W <= 0)
    return fail(h, SPVO_EINVAL, "preprocess: bad shape B=%d rows=%d cols=%d stride=%d H=%d W=%d", B, rows, cols, stride, H, W);
  if (B > 0 && (!imgs || (!input_out && !resized_out))) return fail(h, SPVO_EINVAL, "preprocess: NULL buffer");
  int cr, cc, ro, co;
  if (!preprocess_geometry(rows, cols, H, W, &cr, &cc, &ro, &co))
    return fail(h, SPVO_EINVAL, "preprocess: the %dx%d crop for a %dx%d network input is empty", cols, rows, W, H);
  return SPVO_OK;
}

// BASE:93, 109, 119-120 on B host-side 3x4 row-major matrices
static void patch_projection(float* proj, int B, int rows, int cols, int H, int W) {
  int cr, cc, ro, co;
  if (!proj || !preprocess_geometry(rows, cols, H, W, &cr, &cc, &ro, &co)) return;
  const float r = static_cast<float>(W) / static_cast<float>(cc);
  for (int b = 0; b < B; ++b) {
    float* P = proj + (size_t)b * 12;
    P[1 * 4 + 2] -= static_cast<float>(ro);
    P[0 * 4 + 2] -= static_cast<float>(co);
    for (int i = 0; i < 8; ++i) P[i] *= r;
  }
}

int spvo_preprocess_device(spvo_handle hh, const uint8_t* imgs, int B, int rows, int cols, int stride, int H, int W,
                           float* input_out, uint8_t* resized_out, float* proj) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  int rc = check_preprocess_args(h, imgs, B, rows, cols, stride, H, W, input_out, resized_out);
  if (rc) return rc;
  DeviceGuard g(h->device);
  CK(launch_preprocess(h, imgs, B, rows, cols, stride, H, W, input_out, resized_out));
  patch_projection(proj, B, rows, cols, H, W);
  return SPVO_OK;
}

int spvo_preprocess(spvo_handle hh, const uint8_t* imgs, int B, int rows, int cols, int stride, int H, int W,
                    float* input_out, uint8_t* resized_out, float* proj) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  int rc = check_preprocess_args(h, imgs, B, rows, cols, stride, H, W, input_out, resized_out);
  if (rc) return rc;
  if (B == 0) return SPVO_OK;
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  const size_t src_bytes = (size_t)B * rows * stride, px = (size_t)B * H * W;
  if (h->pp_src_bytes < src_bytes) {
    cudaFree(h->pp_src);
    h->pp_src = nullptr;
    h->pp_src_bytes = 0;
    CK(cudaMalloc((void**)&h->pp_src, src_bytes));
    h->pp_src_bytes = src_bytes;
  }
  if (h->pp_dst_px < px) {
    cudaFree(h->pp_dst_f);
    cudaFree(h->pp_dst_u8);
    h->pp_dst_f = nullptr;
    h->pp_dst_u8 = nullptr;
    h->pp_dst_px = 0;
    CK(cudaMalloc((void**)&h->pp_dst_f, px * sizeof(float)));
    CK(cudaMalloc((void**)&h->pp_dst_u8, px));
    h->pp_dst_px = px;
  }
  CK(cudaMemcpyAsync(h->pp_src, imgs, src_bytes, cudaMemcpyHostToDevice, st));
  CK(launch_preprocess(h, h->pp_src, B, rows, cols, stride, H, W, input_out ? h->pp_dst_f : nullptr,
                       resized_out ? h->pp_dst_u8 : nullptr));
  if (input_out) CK(cudaMemcpyAsync(input_out, h->pp_dst_f, px * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (resized_out) CK(cudaMemcpyAsync(resized_out, h->pp_dst_u8, px, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  patch_projection(proj, B, rows, cols, H, W);
  return SPVO_OK;
}

static int check_match_cfg(Handle* h, const spvo_match_cfg* cfg, int dim) {
  if (!cfg) return fail(h, SPVO_EINVAL, "match: cfg is NULL");
  if (dim != SPVO_DESC_DIM) return fail(h, SPVO_EINVAL, "match: dim %d unsupported (SuperPoint descriptors are 256-d)", dim);
  if (cfg->mode < SPVO_MATCH_NN || cfg->mode > SPVO_MATCH_KNN_RATIO) return fail(h, SPVO_EINVAL, "match: mode %d", cfg->mode);
  if (cfg->algorithm < SPVO_MATCHER_AUTO || cfg->algorithm > SPVO_MATCHER_TENSOR)
    return fail(h, SPVO_EINVAL, "match: algorithm %d", cfg->algorithm);
  if (cfg->flags & ~SPVO_MATCH_FLAG_ROW_BAND) return fail(h, SPVO_EINVAL, "match: unknown flags 0x%x", cfg->flags);
  return SPVO_OK;
}

static int pick_algorithm(const spvo_match_cfg* cfg, int max_rows, int max_cols, int P) {
  // AUTO, from scripts/match_sweep.py on B200 (profiles/match_sweep_r02.json).  The tensor path costs six launches
  // (~35 us for one small problem, flat up to N ~ 1000) and then ~1 ps per pair; the exact fp32 path two to three
  // launches (~19 us) and 45 ps (NN, ratio) to 77 ps (cross-check) per pair while a single problem cannot fill the
  // GPU, ~15 ps per pair in a large batch.  Cross-overs: one problem of N = M = 384 (cross-check), ~576 (NN), ~640
  // (ratio test); 296 problems (148 stereo pairs) of 64 keypoints.
  int alg = cfg->algorithm;
  if (alg == SPVO_MATCHER_AUTO) {
    const long long pairs = (long long)max_rows * max_cols;
    const long long side = cfg->mode == SPVO_MATCH_NN_CROSSCHECK ? 384 : cfg->mode == SPVO_MATCH_KNN_RATIO ? 640 : 576;
    alg = (pairs <= side * side && pairs * (long long)(P > 0 ? P : 1) <= 1500000LL) ? SPVO_MATCHER_EXACT_FP32
                                                                                   : SPVO_MATCHER_TENSOR;
  }
  if (max_rows > 8192 || max_cols > 8192) alg = SPVO_MATCHER_EXACT_FP32;  // packed shortlist keys carry 13 index bits
  return alg;
}

static int run_match(Handle* h, const MatchProblem* probs, int P, int max_rows, int max_cols,
                     const spvo_match_cfg* cfg, spvo_dmatch* out, int* n_matches, int* q2t, int out_stride,
                     bool operands_ready = false) {
  const int alg = pick_algorithm(cfg, max_rows, max_cols, P);
  if (alg == SPVO_MATCHER_TENSOR)
    CK(launch_match_tc(h, probs, P, max_rows, max_cols, *cfg, out, n_matches, q2t, out_stride, operands_ready));
  else
    CK(launch_match_exact(h, probs, P, max_rows, max_cols, *cfg, out, n_matches, q2t, out_stride));
  return SPVO_OK;
}

static int match_device_impl(Handle* h, const float* q, int N, const float* t, int M, int dim, const spvo_match_cfg* cfg,
                             const spvo_keypoint* q_kpts, const spvo_keypoint* t_kpts, float band, spvo_dmatch* out,
                             int* n_matches, int* q2t) {
  if (!h) return SPVO_EINVAL;
  int rc = check_match_cfg(h, cfg, dim);
  if (rc) return rc;
  if (N < 0 || M < 0 || !n_matches || (N > 0 && (!q || !out)) || (M > 0 && !t))
    return fail(h, SPVO_EINVAL, "match: bad arguments (N %d, M %d)", N, M);
  DeviceGuard g(h->device);
  spvo_match_cfg c = *cfg;
  c.flags = (q_kpts && t_kpts && band >= 0.0f) ? (c.flags | SPVO_MATCH_FLAG_ROW_BAND) : (c.flags & ~SPVO_MATCH_FLAG_ROW_BAND);
  CK(launch_set_problem(h, h->probs, q, N, t, M, q_kpts, t_kpts, band));
  h->chain_launches = true;
  return run_match(h, h->probs, 1, N, M, &c, out, n_matches, q2t, N > 0 ? N : 1);
}

int spvo_match_device(spvo_handle hh, const float* q, int N, const float* t, int M, int dim,
                      const spvo_match_cfg* cfg, spvo_dmatch* out, int* n_matches, int* q2t) {
  return match_device_impl(reinterpret_cast<Handle*>(hh), q, N, t, M, dim, cfg, nullptr, nullptr, -1.0f, out, n_matches,
                           q2t);
}

int spvo_match_masked_device(spvo_handle hh, const float* q, int N, const float* t, int M, int dim,
                             const spvo_match_cfg* cfg, const spvo_keypoint* q_kpts, const spvo_keypoint* t_kpts,
                             float band, spvo_dmatch* out, int* n_matches, int* q2t) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return SPVO_EINVAL;
  if (!q_kpts || !t_kpts || !(band >= 0.0f)) return fail(h, SPVO_EINVAL, "match_masked: keypoints and band >= 0 are required");
  return match_device_impl(h, q, N, t, M, dim, cfg, q_kpts, t_kpts, band, out, n_matches, q2t);
}

static int match_host_impl(Handle* h, const float* q, int N, const float* t, int M, int dim, const spvo_match_cfg* cfg,
                           const spvo_keypoint* q_kpts, const spvo_keypoint* t_kpts, float band, spvo_dmatch* out,
                           int* n_matches, int* q2t) {
  if (!h) return SPVO_EINVAL;
  int rc = check_match_cfg(h, cfg, dim);
  if (rc) return rc;
  if (N < 0 || M < 0 || !n_matches || (N > 0 && (!q || !out)) || (M > 0 && !t))
    return fail(h, SPVO_EINVAL, "match: bad arguments (N %d, M %d)", N, M);
  *n_matches = 0;
  if (N == 0) return SPVO_OK;
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  const size_t rows = (size_t)(N > M ? N : M);
  if (h->st_rows < rows) {
    void** ps[] = {(void**)&h->st_q, (void**)&h->st_t, (void**)&h->st_matches, (void**)&h->st_q2t, (void**)&h->st_nm,
                   (void**)&h->st_mkp};
    for (void** p : ps) {
      if (*p) cudaFree(*p);
      *p = nullptr;
    }
    h->st_rows = 0;
    CK(cudaMalloc((void**)&h->st_q, rows * 256 * sizeof(float)));
    CK(cudaMalloc((void**)&h->st_t, rows * 256 * sizeof(float)));
    CK(cudaMalloc((void**)&h->st_matches, rows * sizeof(spvo_dmatch)));
    CK(cudaMalloc((void**)&h->st_q2t, rows * sizeof(int)));
    CK(cudaMalloc((void**)&h->st_nm, sizeof(int)));
    CK(cudaMalloc((void**)&h->st_mkp, 2 * rows * sizeof(spvo_keypoint)));
    h->st_rows = rows;
  }
  const bool masked = q_kpts && t_kpts && band >= 0.0f;
  spvo_keypoint* d_qk = masked ? h->st_mkp : nullptr;
  spvo_keypoint* d_tk = masked ? h->st_mkp + rows : nullptr;
  CK(cudaMemcpyAsync(h->st_q, q, (size_t)N * 256 * sizeof(float), cudaMemcpyHostToDevice, st));
  if (M > 0) CK(cudaMemcpyAsync(h->st_t, t, (size_t)M * 256 * sizeof(float), cudaMemcpyHostToDevice, st));
  if (masked) {
    CK(cudaMemcpyAsync(d_qk, q_kpts, (size_t)N * sizeof(spvo_keypoint), cudaMemcpyHostToDevice, st));
    if (M > 0) CK(cudaMemcpyAsync(d_tk, t_kpts, (size_t)M * sizeof(spvo_keypoint), cudaMemcpyHostToDevice, st));
  }
  spvo_match_cfg c = *cfg;
  c.flags = masked ? (c.flags | SPVO_MATCH_FLAG_ROW_BAND) : (c.flags & ~SPVO_MATCH_FLAG_ROW_BAND);
  CK(launch_set_problem(h, h->probs, h->st_q, N, h->st_t, M, d_qk, d_tk, band));
  h->chain_launches = true;
  rc = run_match(h, h->probs, 1, N, M, &c, h->st_matches, h->st_nm, h->st_q2t, N);
  if (rc) return rc;
  CK(cudaMemcpyAsync(n_matches, h->st_nm, sizeof(int), cudaMemcpyDeviceToHost, st));
  if (q2t) CK(cudaMemcpyAsync(q2t, h->st_q2t, (size_t)N * sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (*n_matches > 0)
    CK(cudaMemcpy(out, h->st_matches, (size_t)*n_matches * sizeof(spvo_dmatch), cudaMemcpyDeviceToHost));
  return SPVO_OK;
}

int spvo_match(spvo_handle hh, const float* q, int N, const float* t, int M, int dim, const spvo_match_cfg* cfg,
               spvo_dmatch* out, int* n_matches, int* q2t) {
  return match_host_impl(reinterpret_cast<Handle*>(hh), q, N, t, M, dim, cfg, nullptr, nullptr, -1.0f, out, n_matches,
                         q2t);
}

int spvo_match_masked(spvo_handle hh, const float* q, int N, const float* t, int M, int dim, const spvo_match_cfg* cfg,
                      const spvo_keypoint* q_kpts, const spvo_keypoint* t_kpts, float band, spvo_dmatch* out,
                      int* n_matches, int* q2t) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return SPVO_EINVAL;
  if (N > 0 && M > 0 && (!q_kpts || !t_kpts || !(band >= 0.0f)))
    return fail(h, SPVO_EINVAL, "match_masked: keypoints and band >= 0 are required");
  return match_host_impl(h, q, N, t, M, dim, cfg, q_kpts, t_kpts, band, out, n_matches, q2t);
}

int spvo_match_batch_device(spvo_handle hh, const float* desc_base, const int* n_rows, int slot_stride_rows,
                            const int* q_slot, const int* t_slot, int P, int max_rows, int dim,
                            const spvo_match_cfg* cfg, spvo_dmatch* out, int* n_matches, int* q2t) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return SPVO_EINVAL;
  int rc = check_match_cfg(h, cfg, dim);
  if (rc) return rc;
  if (P < 0 || max_rows < 0 || slot_stride_rows < max_rows || !desc_base || !n_rows || !q_slot || !t_slot || !out ||
      !n_matches)
    return fail(h, SPVO_EINVAL, "match_batch: bad arguments (P %d, max_rows %d, stride %d)", P, max_rows, slot_stride_rows);
  if (P == 0) return SPVO_OK;
  DeviceGuard g(h->device);
  if (P > h->probs_cap) {
    cudaFree(h->probs);
    h->probs = nullptr;
    h->probs_cap = 0;
    CK(cudaMalloc((void**)&h->probs, (size_t)P * sizeof(MatchProblem)));
    h->probs_cap = P;
  }
  CK(launch_setup_problems(h, h->probs, desc_base, n_rows, slot_stride_rows, q_slot, t_slot, P));
  h->chain_launches = P <= 8;
  return run_match(h, h->probs, P, max_rows, max_rows, cfg, out, n_matches, q2t, max_rows > 0 ? max_rows : 1);
}

int spvo_stereo_filter_batch_device(spvo_handle hh, const spvo_keypoint* kpts_base, int slot_stride_rows,
                                    const int* q_slot, const int* t_slot, int P, int max_rows,
                                    const spvo_dmatch* matches, const int* n_matches, float stereo_threshold,
                                    float min_disparity, uint8_t* keep) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return SPVO_EINVAL;
  if (P < 0 || max_rows < 0 || !kpts_base || !q_slot || !t_slot || !matches || !n_matches || !keep)
    return fail(h, SPVO_EINVAL, "stereo_filter: bad arguments");
  DeviceGuard g(h->device);
  CK(launch_stereo_filter(h, kpts_base, slot_stride_rows, q_slot, t_slot, P, max_rows, matches, n_matches,
                          stereo_threshold, min_disparity, keep));
  return SPVO_OK;
}

int spvo_stereo_reset(spvo_handle hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return SPVO_EINVAL;
  DeviceGuard g(h->device);
  h->has_prev = false;
  if (h->carry_n) CK(cudaMemsetAsync(h->carry_n, 0, sizeof(int), h->stream));
  if (h->carry_map) CK(cudaMemsetAsync(h->carry_map, 0xFF, (size_t)2 * (h->max_k > 0 ? h->max_k : 1) * sizeof(int), h->stream));
  return SPVO_OK;
}

static int check_stereo_args(Handle* h, const void* semi, const void* desc, int F, int H, int W,
                             const spvo_stereo_cfg* cfg, const spvo_stereo_out* out) {
  if (!h) return SPVO_EINVAL;
  if (!cfg || !out || !desc) return fail(h, SPVO_EINVAL, "stereo_batch: NULL argument");
  if (F < 0 || 2 * F > h->max_batch) return fail(h, SPVO_EINVAL, "stereo_batch: 2*F = %d exceeds max_batch %d", 2 * F, h->max_batch);
  if (!out->kpts || !out->n_kpts || !out->matches || !out->n_matches)
    return fail(h, SPVO_EINVAL, "stereo_batch: kpts, n_kpts, matches and n_matches outputs are required");
  int rc = check_decode_args(h, semi, 2 * F, H, W, &cfg->decode, out->kpts, out->n_kpts);
  if (rc) return rc;
  if (cfg->decode.max_keypoints < 1) return fail(h, SPVO_EINVAL, "stereo_batch: max_keypoints must be >= 1");
  if ((out->quads || out->n_quads) && !(out->quads && out->n_quads && out->q2t && out->stereo_keep))
    return fail(h, SPVO_EINVAL, "stereo_batch: quads need n_quads, q2t and stereo_keep outputs");
  return check_match_cfg(h, &cfg->match, SPVO_DESC_DIM);
}

static int ensure_carry(Handle* h) {
  if (h->carry_desc) return SPVO_OK;
  const size_t K = h->max_k > 0 ? h->max_k : 1;
  CK(cudaMalloc((void**)&h->carry_map, 2 * K * sizeof(int)));
  CK(cudaMemsetAsync(h->carry_map, 0xFF, 2 * K * sizeof(int), h->stream));
  CK(cudaMalloc((void**)&h->carry_desc, K * 256 * sizeof(float)));
  CK(cudaMalloc((void**)&h->carry_kpts, K * sizeof(spvo_keypoint)));
  CK(cudaMalloc((void**)&h->carry_n, sizeof(int)));
  CK(cudaMemsetAsync(h->carry_n, 0, sizeof(int), h->stream));
  h->has_prev = false;
  return SPVO_OK;
}

// The device pipeline shared by both forms.  desc_out must be a device buffer [2F,K,256].
static int stereo_pipeline(Handle* h, const void* semi, const void* desc, int in_f16, int F, int H, int W,
                           const spvo_stereo_cfg* cfg, spvo_keypoint* kpts, float* desc_out, int* n_kpts,
                           spvo_dmatch* matches, int* n_matches, int* q2t, uint8_t* keep, spvo_quad* quads,
                           int* n_quads) {
  const int K = cfg->decode.max_keypoints;
  cudaStream_t st = h->stream;
  int rc = ensure_carry(h);
  if (rc) return rc;
  // tensor matcher: decode writes each image's bf16 operand straight into slot = image index;
  // slot max_batch holds the carried last-left image of the previous batch
  h->chain_launches = F <= 4;  // latency-shaped calls: overlap launch latency along the kernel chain (common.cuh)
  const bool tensor = pick_algorithm(&cfg->match, K, K, 2 * F) == SPVO_MATCHER_TENSOR;
  const int carry_slot = h->max_batch;
  TcSink sink;
  if (tensor) CK(tc_prepare_slots(h, h->max_batch + 1, h->max_k, 2 * h->max_batch, &sink));
  if (2 * F > h->probs_cap) {
    cudaFree(h->probs);
    h->probs = nullptr;
    h->probs_cap = 0;
    CK(cudaMalloc((void**)&h->probs, (size_t)2 * F * sizeof(MatchProblem)));
    h->probs_cap = 2 * F;
  }
  const bool band = (cfg->match.flags & SPVO_MATCH_FLAG_ROW_BAND) != 0;  // L<->R under the row band (opt-in)
  // the match problems are built by the decode's last kernel when one launch covers the batch (else by k_setup_*)
  h->stereo_setup = StereoSetup();
  h->stereo_setup.probs = h->probs; h->stereo_setup.desc_out = desc_out; h->stereo_setup.n_out = n_kpts;
  h->stereo_setup.carry_desc = h->carry_desc; h->stereo_setup.carry_n = h->carry_n;
  h->stereo_setup.F = F; h->stereo_setup.K = K; h->stereo_setup.carry_slot = carry_slot;
  h->stereo_setup.kpts = band ? kpts : nullptr; h->stereo_setup.band = band ? cfg->stereo_threshold : -1.0f;
  h->stereo_setup_done = false;
  bool sink_filled = false;  // false when decode used the gather form (planes larger than shared memory)
  const cudaError_t dec_rc = launch_decode(h, semi, desc, in_f16, 2 * F, H, W, cfg->decode, kpts, desc_out, n_kpts, nullptr,
                                           tensor ? &sink : nullptr, &sink_filled);
  const bool setup_done = h->stereo_setup_done;
  h->stereo_setup = StereoSetup();  // one-shot
  h->stereo_setup_done = false;
  CK(dec_rc);
  if (!setup_done)
    CK(launch_setup_stereo_problems(h, h->probs, desc_out, n_kpts, F, K, carry_slot, band ? kpts : nullptr,
                                    band ? cfg->stereo_threshold : -1.0f));
  const bool ready = tensor && sink_filled;
  if (ready && h->has_prev && !h->carry_tc_valid)
    // the previous batch ran on the exact matcher: convert the carried fp32 descriptors into the carry slot
    CK(tc_prep_problem_operands(h, h->probs + F));
  // !ready: the matcher converts every operand itself (k_tc_prep, bf16) -- decode did not fill the slots
  if (keep) {  // the row-band / min-disparity test rides in k_finalize_matches (stereo problems = the first F)
    h->fin_filter.kpts = kpts;
    h->fin_filter.slot_stride = K;
    h->fin_filter.nprob = F;
    h->fin_filter.stereo_threshold = cfg->stereo_threshold;
    h->fin_filter.min_disparity = cfg->min_disparity;
    h->fin_filter.keep = keep;
  }
  rc = run_match(h, h->probs, 2 * F, K, K, &cfg->match, matches, n_matches, q2t, K, ready);
  h->fin_filter = FilterArgs();
  if (rc) return rc;
  // ONE launch for the tail: the quadruples (BASE:156-207) and, in extra blocks, the carry for the next batch -- the
  // previous frame's L<->R map (BASE:475-481) and the last left image (descriptors, keypoints, count, matcher operand
  // slot) for the next batch's first temporal match.  The map is double-buffered: the consistency blocks read the
  // half written by the previous batch while the copy blocks fill the other one.
  const size_t mk = (size_t)(h->max_k > 0 ? h->max_k : 1);
  const int* map_in = h->carry_map + (size_t)h->carry_parity * mk;
  int* map_out = h->carry_map + (size_t)(h->carry_parity ^ 1) * mk;
  CopyList cl;
  cl.n = 0;
  const size_t last = (size_t)2 * (F - 1);
  if (q2t) cl.seg[cl.n++] = {q2t + (size_t)(F - 1) * K, map_out, (size_t)K * sizeof(int)};
  cl.seg[cl.n++] = {desc_out + last * K * 256, h->carry_desc, (size_t)K * 256 * sizeof(float)};
  cl.seg[cl.n++] = {kpts + last * K, h->carry_kpts, (size_t)K * sizeof(spvo_keypoint)};
  cl.seg[cl.n++] = {n_kpts + last, h->carry_n, sizeof(int)};
  if (ready) cl.n += tc_copy_slot_segments(h, carry_slot, (int)last, cl.seg + cl.n);
  CK(launch_consistency(h, F, K, matches, n_matches, q2t, keep, map_in, quads, n_quads, cl));
  if (q2t) h->carry_parity ^= 1;
  h->carry_tc_valid = ready;
  h->has_prev = true;
  return SPVO_OK;
}

static int stereo_batch_device_impl(Handle* h, const void* semi, const void* desc, int in_f16, int F, int H, int W,
                                    const spvo_stereo_cfg* cfg, const spvo_stereo_out* out) {
  int rc = check_stereo_args(h, semi, desc, F, H, W, cfg, out);
  if (rc) return rc;
  if (!out->desc) return fail(h, SPVO_EINVAL, "stereo_batch_device: desc output is required");
  if (F == 0) return SPVO_OK;
  DeviceGuard g(h->device);
  auto run = [&]() {
    return stereo_pipeline(h, semi, desc, in_f16, F, H, W, cfg, out->kpts, out->desc, out->n_kpts, out->matches,
                           out->n_matches, out->q2t, out->stereo_keep, out->quads, out->n_quads);
  };
  if (!h->graph_mode || h->profiling || h->decode_subbatches > 1) return run();
  // Graph mode (the reference's real-time shape: one pair per callback, fixed engine bindings,
  // visual_odometry_node.cpp:150-262): the call's launches are captured ONCE per signature -- arguments plus the
  // host-side stream state that selects code paths -- and replayed afterwards; any change re-captures.
  std::vector<unsigned char> sig;
  auto put = [&](const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    sig.insert(sig.end(), b, b + n);
  };
  const int state[8] = {in_f16, F, H, W, h->has_prev ? 1 : 0, h->carry_tc_valid ? 1 : 0, h->carry_parity, 0};
  put(&semi, sizeof(semi)); put(&desc, sizeof(desc)); put(state, sizeof(state)); put(cfg, sizeof(*cfg));
  put(out, sizeof(*out)); put(&h->stream, sizeof(h->stream));
  for (Handle::GraphEntry& ge : h->graphs)
    if (ge.sig == sig) {
      CK(cudaGraphLaunch(ge.exec, h->stream));
      h->launches += ge.kernels;
      h->has_prev = true;
      h->carry_tc_valid = ge.ready;
      if (ge.flips_parity) h->carry_parity ^= 1;
      return SPVO_OK;
    }
  bool seen = false;
  for (const auto& sg : h->graph_seen) seen = seen || sg == sig;
  if (!seen) {  // first time: run eagerly, so that every workspace this signature needs exists before a capture
    if (h->graph_seen.size() >= 32) h->graph_seen.erase(h->graph_seen.begin());
    h->graph_seen.push_back(sig);
    return run();
  }
  const long long l0 = h->launches;
  const int parity0 = h->carry_parity;
  CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed));
  rc = run();
  cudaGraph_t graph = nullptr;
  const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (ce != cudaSuccess) return cuda_fail(h, ce, "cudaStreamEndCapture");
  Handle::GraphEntry ge;
  ge.sig = sig;
  ge.kernels = (int)(h->launches - l0);
  ge.ready = h->carry_tc_valid;
  ge.flips_parity = h->carry_parity != parity0;
  const cudaError_t ie = cudaGraphInstantiate(&ge.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (ie != cudaSuccess) return cuda_fail(h, ie, "cudaGraphInstantiate");
  if (h->graphs.size() >= 16) {  // e.g. a ring of 6 input batches x 2 carry parities
    cudaGraphExecDestroy(h->graphs.front().exec);
    h->graphs.erase(h->graphs.begin());
  }
  h->graphs.push_back(ge);
  CK(cudaGraphLaunch(ge.exec, h->stream));  // the captured work has not run yet
  return SPVO_OK;
}

int spvo_set_graph_mode(spvo_handle hh, int on) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return SPVO_EINVAL;
  h->graph_mode = on != 0;
  return SPVO_OK;
}

int spvo_stereo_batch_device(spvo_handle hh, const float* semi, const float* desc, int F, int H, int W,
                             const spvo_stereo_cfg* cfg, const spvo_stereo_out* out) {
  return stereo_batch_device_impl(reinterpret_cast<Handle*>(hh), semi, desc, 0, F, H, W, cfg, out);
}

int spvo_stereo_batch_device_f16(spvo_handle hh, const void* semi, const void* desc, int F, int H, int W,
                                 const spvo_stereo_cfg* cfg, const spvo_stereo_out* out) {
  return stereo_batch_device_impl(reinterpret_cast<Handle*>(hh), semi, desc, 1, F, H, W, cfg, out);
}

static int stereo_batch_host_impl(Handle* h, const void* semi_v, const void* desc_v, int in_f16, int F, int H, int W,
                                  const spvo_stereo_cfg* cfg, const spvo_stereo_out* out) {
  const unsigned char* semi = static_cast<const unsigned char*>(semi_v);
  const unsigned char* desc = static_cast<const unsigned char*>(desc_v);
  const size_t esz = in_f16 ? 2 : 4;
  unsigned char* st_semi = reinterpret_cast<unsigned char*>(h ? h->st_semi : nullptr);
  unsigned char* st_desc = reinterpret_cast<unsigned char*>(h ? h->st_desc : nullptr);
  int rc = check_stereo_args(h, semi, desc, F, H, W, cfg, out);
  if (rc) return rc;
  if (F == 0) return SPVO_OK;
  DeviceGuard g(h->device);
  cudaStream_t st = h->stream;
  const size_t cells = (size_t)(H / 8) * (W / 8), K = cfg->decode.max_keypoints, B = (size_t)2 * F;
  if (!h->st_smatches) {
    const size_t mb = h->max_batch, mk = h->max_k > 0 ? h->max_k : 1;
    CK(cudaMalloc((void**)&h->st_smatches, mb * mk * sizeof(spvo_dmatch)));
    CK(cudaMalloc((void**)&h->st_snm, mb * sizeof(int)));
    CK(cudaMalloc((void**)&h->st_sq2t, mb * mk * sizeof(int)));
    CK(cudaMalloc((void**)&h->st_skeep, mb * mk));
    CK(cudaMalloc((void**)&h->st_quads, mb * mk * sizeof(spvo_quad)));
    CK(cudaMalloc((void**)&h->st_nquads, mb * sizeof(int)));
  }
  // The batch is processed as up to 4 chunks of consecutive frames: the H2D copy of chunk c+1 (copy stream)
  // overlaps the kernels and the D2H of chunk c (compute stream).  Chunks continue the sequence through the
  // handle's carry exactly like consecutive calls, so results do not depend on the chunking.
  (void)B;
  const int nchunk = F >= 16 ? 4 : 1;
  if (!h->copy_stream) {
    CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 5; ++i) CK(cudaEventCreateWithFlags(&h->copy_ev[i], cudaEventDisableTiming));
  }
  CK(cudaEventRecord(h->copy_ev[4], st));
  CK(cudaStreamWaitEvent(h->copy_stream, h->copy_ev[4], 0));
  for (int c = 0; c < nchunk; ++c) {
    const size_t f0 = (size_t)F * c / nchunk, f1 = (size_t)F * (c + 1) / nchunk, nb = 2 * (f1 - f0);
    CK(cudaMemcpyAsync(st_semi + 2 * f0 * 65 * cells * esz, semi + 2 * f0 * 65 * cells * esz, nb * 65 * cells * esz,
                       cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaMemcpyAsync(st_desc + 2 * f0 * 256 * cells * esz, desc + 2 * f0 * 256 * cells * esz, nb * 256 * cells * esz,
                       cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaEventRecord(h->copy_ev[c], h->copy_stream));
  }
  for (int c = 0; c < nchunk; ++c) {
    const size_t f0 = (size_t)F * c / nchunk, f1 = (size_t)F * (c + 1) / nchunk, Fc = f1 - f0, nb = 2 * Fc;
    CK(cudaStreamWaitEvent(st, h->copy_ev[c], 0));
    // chunk-local device layout: images [2 f0, 2 f1); match rows [2 f0, 2 f0 + Fc) stereo, then Fc temporal
    spvo_keypoint* d_kp = h->st_kpts + 2 * f0 * K;
    float* d_desc = h->st_desc_out + 2 * f0 * K * 256;
    int* d_n = h->st_n + 2 * f0;
    spvo_dmatch* d_m = h->st_smatches + 2 * f0 * K;
    int* d_nm = h->st_snm + 2 * f0;
    int* d_q2t = h->st_sq2t + 2 * f0 * K;
    uint8_t* d_keep = h->st_skeep + f0 * K;
    spvo_quad* d_quads = h->st_quads + f0 * K;
    int* d_nq = h->st_nquads + f0;
    rc = stereo_pipeline(h, st_semi + 2 * f0 * 65 * cells * esz, st_desc + 2 * f0 * 256 * cells * esz, in_f16, (int)Fc, H, W,
                         cfg, d_kp,
                         d_desc, d_n, d_m, d_nm, d_q2t, out->stereo_keep ? d_keep : nullptr,
                         out->quads ? d_quads : nullptr, d_nq);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out->kpts + 2 * f0 * K, d_kp, nb * K * sizeof(spvo_keypoint), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(out->n_kpts + 2 * f0, d_n, nb * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (out->desc)
      CK(cudaMemcpyAsync(out->desc + 2 * f0 * K * 256, d_desc, nb * K * 256 * sizeof(float), cudaMemcpyDeviceToHost, st));
    // stereo rows -> [f0, f1), temporal rows -> [F + f0, F + f1) of the caller's whole-batch layout
    for (int part = 0; part < 2; ++part) {
      const size_t dst_row = (part ? (size_t)F : 0) + f0, src_row = part ? Fc : 0;
      CK(cudaMemcpyAsync(out->matches + dst_row * K, d_m + src_row * K, Fc * K * sizeof(spvo_dmatch),
                         cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(out->n_matches + dst_row, d_nm + src_row, Fc * sizeof(int), cudaMemcpyDeviceToHost, st));
      if (out->q2t)
        CK(cudaMemcpyAsync(out->q2t + dst_row * K, d_q2t + src_row * K, Fc * K * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    if (out->stereo_keep) CK(cudaMemcpyAsync(out->stereo_keep + f0 * K, d_keep, Fc * K, cudaMemcpyDeviceToHost, st));
    if (out->quads) {
      CK(cudaMemcpyAsync(out->quads + f0 * K, d_quads, Fc * K * sizeof(spvo_quad), cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(out->n_quads + f0, d_nq, Fc * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
  }
  CK(cudaStreamSynchronize(st));
  return SPVO_OK;
}

int spvo_stereo_batch(spvo_handle hh, const float* semi, const float* desc, int F, int H, int W,
                      const spvo_stereo_cfg* cfg, const spvo_stereo_out* out) {
  return stereo_batch_host_impl(reinterpret_cast<Handle*>(hh), semi, desc, 0, F, H, W, cfg, out);
}

int spvo_stereo_batch_f16(spvo_handle hh, const void* semi, const void* desc, int F, int H, int W,
                          const spvo_stereo_cfg* cfg, const spvo_stereo_out* out) {
  return stereo_batch_host_impl(reinterpret_cast<Handle*>(hh), semi, desc, 1, F, H, W, cfg, out);
}

long long spvo_kernel_launches(spvo_handle hh) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  return h ? h->launches : -1;
}

static const char* kKernelNames[KID_COUNT] = {
    "k_softmax_heat", "k_detect", "k_sample_desc", "k_dist_exact", "k_row_select", "k_col_select",
    "k_finalize_matches", "k_setup_problems", "k_stereo_filter", "k_tc_prep", "k_tc_gemm", "k_tc_rerank",
    "k_tc_fallback", "k_tc_fill_dist", "k_tc_triage", "k_desc_planes", "k_desc_normalize", "k_consistency", "k_carry_copy", "k_preprocess"};  // (k_carry_copy: folded into k_consistency)

int spvo_profile_num_kernels(void) { return KID_COUNT; }

const char* spvo_profile_kernel_name(int k) { return (k >= 0 && k < KID_COUNT) ? kKernelNames[k] : ""; }

int spvo_profile_enable(spvo_handle hh, int on) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h) return SPVO_EINVAL;
  h->profiling = on != 0;
  return SPVO_OK;
}

int spvo_profile_read(spvo_handle hh, double* ms, long long* launches, int n) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || n < 0) return SPVO_EINVAL;
  DeviceGuard g(h->device);
  CK(cudaStreamSynchronize(h->stream));
  for (const ProfRec& r : h->prof_recs) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
      h->prof_ms[r.kid] += t;
      h->prof_n[r.kid] += 1;
    }
    h->ev_pool.push_back(r.a);
    h->ev_pool.push_back(r.b);
  }
  h->prof_recs.clear();
  for (int i = 0; i < n && i < KID_COUNT; ++i) {
    if (ms) ms[i] = h->prof_ms[i];
    if (launches) launches[i] = h->prof_n[i];
  }
  for (int i = 0; i < KID_COUNT; ++i) {
    h->prof_ms[i] = 0;
    h->prof_n[i] = 0;
  }
  return SPVO_OK;
}

int spvo_debug_div_check(spvo_handle hh, const uint32_t* a_bits, const uint32_t* b_bits, long long n,
                         long long* mismatches) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !a_bits || !b_bits || !mismatches || n < 0) return SPVO_EINVAL;
  DeviceGuard g(h->device);
  unsigned long long* d = h->counters + 7;  // the last counter slot is scratch for this check
  CK(cudaMemsetAsync(d, 0, sizeof(*d), h->stream));
  CK(launch_div_check(h, a_bits, b_bits, n, d));
  unsigned long long v = 0;
  CK(cudaMemcpyAsync(&v, d, sizeof(v), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemsetAsync(d, 0, sizeof(*d), h->stream));
  *mismatches = (long long)v;
  return SPVO_OK;
}

int spvo_debug_counters(spvo_handle hh, long long* out, int n) {
  Handle* h = reinterpret_cast<Handle*>(hh);
  if (!h || !out || n < 0) return SPVO_EINVAL;
  DeviceGuard g(h->device);
  unsigned long long c[8];
  CK(cudaMemcpy(c, h->counters, sizeof(c), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n && i < 8; ++i) out[i] = (long long)c[i];
  return SPVO_OK;
}

}  // extern "C"
