// preprocess.cu -- the step before the network: FeatureFrontEnd::preprocessImageImpl
// (reference: src/odml_visual_odometry/src/feature_detection_base.cpp:68-121) and the 8U -> 32F hand-over of
// SuperPointFeatureFrontEnd::preprocessImage (src/feature_detection_neural_network.cpp:139-161):
// centre crop to the network's aspect ratio, cv::resize(INTER_LINEAR) of the 8UC1 image, input = pixel * (1/255).
//
// The resize reproduces OpenCV's fixed-point path bit for bit (imgproc/src/resize.cpp): coefficients
// cvRound((1 - f) * 2048), cvRound(f * 2048) with f from (float)((d + 0.5) * scale - 0.5); horizontal pass to int,
// vertical pass (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2; an exact 2x decimation takes
// OpenCV's INTER_AREA shortcut (a + b + c + d + 2) >> 2.  One thread per output pixel: byte/integer work, HBM-bound
// on the fp32 store (4 B per output pixel against ~1 B of source).
#include "common.cuh"

namespace spvo {

struct PreprocParams {
  const uint8_t* src;     // [B, rows, stride], already offset to the crop's first pixel
  size_t src_image_bytes; // rows * stride
  int stride, sh, sw;     // crop size
  int H, W;
  int area2x;             // exact 2x decimation
  const uint2* xtab;      // [W] {sx | sx1 << 16, a0 | a1 << 16}
  const uint2* ytab;      // [H] {y0 | y1 << 16, b0 | b1 << 16}
  float* out_f;           // [B, H, W] or NULL
  uint8_t* out_u8;        // [B, H, W] or NULL
};

// Coefficient tables of cv::resize's linear path, one entry per output column / row (computed once per geometry;
// the double-precision coordinate arithmetic of resize.cpp stays out of the per-pixel kernel).
__global__ void k_resize_tables(uint2* xtab, uint2* ytab, int W, int H, int sw, int sh, double scale_x, double scale_y) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < W) {
    float fx = (float)(((double)i + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= sw - 1) { fx = 0.f; sx = sw - 1; }
    const int sx1 = min(sx + 1, sw - 1);
    const int a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fx), 2048.f)), a1 = __float2int_rn(__fmul_rn(fx, 2048.f));
    xtab[i] = make_uint2((uint32_t)sx | ((uint32_t)sx1 << 16), (uint32_t)a0 | ((uint32_t)a1 << 16));
  }
  if (i < H) {
    float fy = (float)(((double)i + 0.5) * scale_y - 0.5);
    const int sy = (int)floorf(fy);
    fy = __fsub_rn(fy, (float)sy);
    const int b0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fy), 2048.f)), b1 = __float2int_rn(__fmul_rn(fy, 2048.f));
    const int y0 = min(max(sy, 0), sh - 1), y1 = min(max(sy + 1, 0), sh - 1);
    ytab[i] = make_uint2((uint32_t)y0 | ((uint32_t)y1 << 16), (uint32_t)b0 | ((uint32_t)b1 << 16));
  }
}

__device__ __forceinline__ int resize_px(const uint8_t* __restrict__ s0, const uint8_t* __restrict__ s1, uint2 xe, int b0,
                                         int b1) {
  const int sx = xe.x & 0xFFFF, sx1 = xe.x >> 16, a0 = xe.y & 0xFFFF, a1 = xe.y >> 16;
  const int r0 = __ldg(s0 + sx) * a0 + __ldg(s0 + sx1) * a1, r1 = __ldg(s1 + sx) * a0 + __ldg(s1 + sx1) * a1;
  return (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
}

// One thread = 4 consecutive output pixels of one row: 16-byte fp32 store, 4-byte u8 store; the source bytes come
// through L1 (neighbouring threads share their sectors).
__global__ void __launch_bounds__(256) k_preprocess(const PreprocParams p) {
  const int dx = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4, dy = blockIdx.y * 8 + (threadIdx.x >> 5), b = blockIdx.z;
  if (dx >= p.W || dy >= p.H) return;
  const uint8_t* src = p.src + (size_t)b * p.src_image_bytes;
  const int nx = min(4, p.W - dx);
  int v[4] = {0, 0, 0, 0};
  if (p.area2x) {
    for (int i = 0; i < nx; ++i) {
      const uint8_t* q = src + (size_t)(2 * dy) * p.stride + 2 * (dx + i);
      v[i] = (q[0] + q[1] + q[p.stride] + q[p.stride + 1] + 2) >> 2;
    }
  } else {
    const uint2 ye = __ldg(p.ytab + dy);
    const uint8_t* s0 = src + (size_t)(ye.x & 0xFFFF) * p.stride;
    const uint8_t* s1 = src + (size_t)(ye.x >> 16) * p.stride;
    const int b0 = ye.y & 0xFFFF, b1 = ye.y >> 16;
    if (nx == 4) {
      uint2 xe[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xe[i] = __ldg(p.xtab + dx + i);
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = resize_px(s0, s1, xe[i], b0, b1);
    } else {
      for (int i = 0; i < nx; ++i) v[i] = resize_px(s0, s1, __ldg(p.xtab + dx + i), b0, b1);
    }
  }
  const size_t o = ((size_t)b * p.H + dy) * p.W + dx;
  const float k = 1.0f / 255.0f;
  if (nx == 4 && (p.W & 3) == 0) {  // rows start 16-byte aligned when W % 4 == 0 (cudaMalloc'd / torch tensors)
    if (p.out_f && (reinterpret_cast<uintptr_t>(p.out_f) & 15) == 0)
      *reinterpret_cast<float4*>(p.out_f + o) = make_float4(__fmul_rn((float)v[0], k), __fmul_rn((float)v[1], k),
                                                           __fmul_rn((float)v[2], k), __fmul_rn((float)v[3], k));
    else if (p.out_f)
      for (int i = 0; i < 4; ++i) p.out_f[o + i] = __fmul_rn((float)v[i], k);
    if (p.out_u8 && (reinterpret_cast<uintptr_t>(p.out_u8) & 3) == 0)
      *reinterpret_cast<uchar4*>(p.out_u8 + o) = make_uchar4((uint8_t)v[0], (uint8_t)v[1], (uint8_t)v[2], (uint8_t)v[3]);
    else if (p.out_u8)
      for (int i = 0; i < 4; ++i) p.out_u8[o + i] = (uint8_t)v[i];
  } else {
    for (int i = 0; i < nx; ++i) {
      if (p.out_f) p.out_f[o + i] = __fmul_rn((float)v[i], k);
      if (p.out_u8) p.out_u8[o + i] = (uint8_t)v[i];
    }
  }
}

// Crop geometry of BASE:71-113 (int <- float conversions truncate like the C++ assignments).
bool preprocess_geometry(int rows, int cols, int H, int W, int* cr, int* cc, int* ro, int* co) {
  if (rows <= 0 || cols <= 0 || H <= 0 || W <= 0) return false;
  int img_rows = rows, img_cols = cols;
  *ro = 0;
  *co = 0;
  const float real_ar = static_cast<float>(cols) / static_cast<float>(rows);
  const float expected_ar = static_cast<float>(W) / static_cast<float>(H);
  if (expected_ar > real_ar) {
    img_rows = static_cast<int>(static_cast<float>(img_cols) / expected_ar);
    *ro = (rows - img_rows) / 2;
  } else if (expected_ar < real_ar) {
    img_cols = static_cast<int>(static_cast<float>(img_rows) * expected_ar);
    *co = (cols - img_cols) / 2;
  }
  *cr = img_rows;
  *cc = img_cols;
  return img_rows > 0 && img_cols > 0;
}

cudaError_t launch_preprocess(Handle* h, const uint8_t* imgs, int B, int rows, int cols, int stride, int H, int W,
                              float* out_f, uint8_t* out_u8) {
  int cr, cc, ro, co;
  if (!preprocess_geometry(rows, cols, H, W, &cr, &cc, &ro, &co)) return cudaErrorInvalidValue;
  if (B == 0) return cudaSuccess;
  if (cr > 65535 || cc > 65535) return cudaErrorInvalidValue;
  cudaError_t e;
  const bool area2x = (W * 2 == cc && H * 2 == cr);
  if (!area2x && (h->pp_key[0] != cr || h->pp_key[1] != cc || h->pp_key[2] != H || h->pp_key[3] != W ||
                  h->pp_stream != h->stream)) {  // (a different stream has no ordering with the table kernel)
    // coefficient tables for this geometry (kept until the geometry changes: one camera, one network input size).
    // A k_preprocess launched earlier (possibly on another stream) may still read the old table: order behind it.
    if (h->pp_done && (e = cudaStreamWaitEvent(h->stream, h->pp_done, 0)) != cudaSuccess) return e;
    if (h->pp_tab_cap < (size_t)(W + H)) {
      cudaFree(h->pp_tab);
      h->pp_tab = nullptr;
      h->pp_tab_cap = 0;
      if ((e = cudaMalloc((void**)&h->pp_tab, (size_t)(W + H) * sizeof(uint2))) != cudaSuccess) return e;
      h->pp_tab_cap = (size_t)(W + H);
    }
    h->pp_key[0] = -1;
    const int n = W > H ? W : H;
    k_resize_tables<<<(n + 255) / 256, 256, 0, h->stream>>>(h->pp_tab, h->pp_tab + W, W, H, cc, cr,
                                                          1.0 / ((double)W / cc), 1.0 / ((double)H / cr));
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    h->pp_key[0] = cr; h->pp_key[1] = cc; h->pp_key[2] = H; h->pp_key[3] = W;
    h->pp_stream = h->stream;
  }
  PreprocParams p;
  p.src = imgs + (size_t)ro * stride + co;
  p.src_image_bytes = (size_t)rows * stride;
  p.stride = stride; p.sh = cr; p.sw = cc; p.H = H; p.W = W;
  p.area2x = area2x ? 1 : 0;
  p.xtab = h->pp_tab;
  p.ytab = h->pp_tab ? h->pp_tab + W : nullptr;
  p.out_f = out_f; p.out_u8 = out_u8;
  {
    LaunchScope ls(h, KID_PREPROCESS);
    k_preprocess<<<dim3((W + 127) / 128, (H + 7) / 8, B), 256, 0, h->stream>>>(p);
  }
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (!h->pp_done && (e = cudaEventCreateWithFlags(&h->pp_done, cudaEventDisableTiming)) != cudaSuccess) return e;
  return cudaEventRecord(h->pp_done, h->stream);
}

}  // namespace spvo
