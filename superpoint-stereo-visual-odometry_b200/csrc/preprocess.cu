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
  double scale_x, scale_y;
  int area2x;             // exact 2x decimation
  float* out_f;           // [B, H, W] or NULL
  uint8_t* out_u8;        // [B, H, W] or NULL
};

__global__ void __launch_bounds__(256) k_preprocess(const PreprocParams p) {
  const int dx = blockIdx.x * 32 + (threadIdx.x & 31), dy = blockIdx.y * 8 + (threadIdx.x >> 5), b = blockIdx.z;
  if (dx >= p.W || dy >= p.H) return;
  const uint8_t* src = p.src + (size_t)b * p.src_image_bytes;
  int v;
  if (p.area2x) {
    const uint8_t* q = src + (size_t)(2 * dy) * p.stride + 2 * dx;
    v = (q[0] + q[1] + q[p.stride] + q[p.stride + 1] + 2) >> 2;
  } else {
    float fx = (float)(((double)dx + 0.5) * p.scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= p.sw - 1) { fx = 0.f; sx = p.sw - 1; }
    const int sx1 = min(sx + 1, p.sw - 1);
    const int a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fx), 2048.f)), a1 = __float2int_rn(__fmul_rn(fx, 2048.f));
    float fy = (float)(((double)dy + 0.5) * p.scale_y - 0.5);
    const int sy = (int)floorf(fy);
    fy = __fsub_rn(fy, (float)sy);
    const int b0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, fy), 2048.f)), b1 = __float2int_rn(__fmul_rn(fy, 2048.f));
    const int y0 = min(max(sy, 0), p.sh - 1), y1 = min(max(sy + 1, 0), p.sh - 1);
    const uint8_t* s0 = src + (size_t)y0 * p.stride;
    const uint8_t* s1 = src + (size_t)y1 * p.stride;
    const int r0 = s0[sx] * a0 + s0[sx1] * a1, r1 = s1[sx] * a0 + s1[sx1] * a1;
    v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
  }
  const size_t o = ((size_t)b * p.H + dy) * p.W + dx;
  if (p.out_f) p.out_f[o] = __fmul_rn((float)(v & 255), 1.0f / 255.0f);
  if (p.out_u8) p.out_u8[o] = (uint8_t)v;
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
  PreprocParams p;
  p.src = imgs + (size_t)ro * stride + co;
  p.src_image_bytes = (size_t)rows * stride;
  p.stride = stride; p.sh = cr; p.sw = cc; p.H = H; p.W = W;
  p.scale_x = 1.0 / ((double)W / cc);
  p.scale_y = 1.0 / ((double)H / cr);
  p.area2x = (W * 2 == cc && H * 2 == cr) ? 1 : 0;
  p.out_f = out_f; p.out_u8 = out_u8;
  LaunchScope ls(h, KID_PREPROCESS);
  k_preprocess<<<dim3((W + 31) / 32, (H + 7) / 8, B), 256, 0, h->stream>>>(p);
  return cudaGetLastError();
}

}  // namespace spvo
