// Drives spvo::SuperPointFeatureFrontEnd (include/spvo_frontend.hpp) the way the reference's
// stereoCallback drives its front end (visual_odometry_node.cpp:175-199):
//   per frame: copy the network outputs into output_det_data_/output_desc_data_, call
//   postprocessDetectionAndDescription(), then matchDescriptors(CURR_LEFT_CURR_RIGHT) and, from the
//   second frame on, matchDescriptors(CURR_LEFT_PREV_LEFT).
// usage: test_frontend_mirror <in.bin> <out.bin> <selector NN|KNN> <cross_check 0|1>
//   in.bin : int32 F, H, W, K, batch ; then per frame semi[2,65,Hc,Wc], desc[2,256,Hc,Wc] fp32
//   out.bin: per frame: int32 nL, nR, kptsL[nL*7 f32], kptsR, descL[nL*256], descR,
//            int32 nS, matchesS[nS*4], mapS[nL], int32 nT, matchesT[nT*4], mapT[nL]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "spvo_frontend.hpp"

static void wr(FILE* f, const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) { perror("write"); exit(2); } }

// usage 2: test_frontend_mirror --preprocess <in.bin> <out.bin>
//   in.bin : int32 rows, cols, H, W ; 12 fp32 projection matrix ; rows*cols image bytes
//   out.bin: input [H*W] fp32 (slot 1 of input_data_), resized [H*W] u8 (images_dq.back()), 12 fp32 patched matrix
static int run_preprocess(const char* in, const char* out) {
  FILE* fi = fopen(in, "rb");
  FILE* fo = fopen(out, "wb");
  int hdr[4];
  float P[12];
  if (!fi || !fo || fread(hdr, 4, 4, fi) != 4 || fread(P, 4, 12, fi) != 12) return 1;
  const int rows = hdr[0], cols = hdr[1], H = hdr[2], W = hdr[3];
  std::vector<uint8_t> img((size_t)rows * cols);
  if (fread(img.data(), 1, img.size(), fi) != img.size()) return 1;
  try {
    spvo::SuperPointFeatureFrontEnd fe(spvo::MatcherType::BF, spvo::SelectorType::NN, true, 2, H, W, 0.015f, 4, 4, 2.0f,
                                       0.25f, 100, 0);
    fe.preprocessImage(img.data(), rows, cols, cols, P, 1);  // NN:482: the right image goes to batch slot 1
    wr(fo, fe.input_data_.get() + (size_t)H * W, (size_t)H * W * 4);
    wr(fo, fe.images_dq.back().data(), (size_t)H * W);
    wr(fo, P, sizeof(P));
  } catch (const spvo::Error& e) {
    fprintf(stderr, "spvo error %d: %s\n", e.code, e.what());
    return 4;
  }
  fclose(fo);
  fclose(fi);
  return 0;
}

int main(int argc, char** argv) {
  if (argc == 4 && std::string(argv[1]) == "--preprocess") return run_preprocess(argv[2], argv[3]);
  if (argc < 5) return 1;
  FILE* fi = fopen(argv[1], "rb");
  FILE* fo = fopen(argv[2], "wb");
  if (!fi || !fo) return 1;
  int hdr[5];
  if (fread(hdr, 4, 5, fi) != 5) return 1;
  const int F = hdr[0], H = hdr[1], W = hdr[2], K = hdr[3], batch = hdr[4];
  const size_t cells = (size_t)(H / 8) * (W / 8);
  try {
    spvo::SuperPointFeatureFrontEnd fe(spvo::MatcherType::BF,
                                       std::string(argv[3]) == "KNN" ? spvo::SelectorType::KNN : spvo::SelectorType::NN,
                                       atoi(argv[4]) != 0, batch, H, W, 0.015f, 4, 4, 2.0f, 0.25f, K, 0);
    std::vector<float> semi(2 * 65 * cells), desc(2 * 256 * cells);
    for (int f = 0; f < F; ++f) {
      if (fread(semi.data(), 4, semi.size(), fi) != semi.size()) return 1;
      if (fread(desc.data(), 4, desc.size(), fi) != desc.size()) return 1;
      if (batch == 2) {  // NN:480-484
        std::memcpy(fe.output_det_data_.get(), semi.data(), semi.size() * 4);
        std::memcpy(fe.output_desc_data_.get(), desc.data(), desc.size() * 4);
        fe.postprocessDetectionAndDescription();
      } else {  // NN:468-475: left then right, one image per network run
        for (int eye = 0; eye < 2; ++eye) {
          std::memcpy(fe.output_det_data_.get(), semi.data() + eye * 65 * cells, 65 * cells * 4);
          std::memcpy(fe.output_desc_data_.get(), desc.data() + eye * 256 * cells, 256 * cells * 4);
          fe.postprocessDetectionAndDescription();
        }
      }
      fe.matchDescriptors(spvo::CURR_LEFT_CURR_RIGHT);                 // node:190 / 196-199
      if (f > 0) fe.matchDescriptors(spvo::CURR_LEFT_PREV_LEFT);
      const auto& kl = fe.keypoints_dq.end()[spvo::CURR_LEFT];
      const auto& kr = fe.keypoints_dq.end()[spvo::CURR_RIGHT];
      const auto& dl = fe.descriptors_dq.end()[spvo::CURR_LEFT];
      const auto& dr = fe.descriptors_dq.end()[spvo::CURR_RIGHT];
      int nL = (int)kl.size(), nR = (int)kr.size();
      wr(fo, &nL, 4); wr(fo, &nR, 4);
      wr(fo, kl.data(), (size_t)nL * 28); wr(fo, kr.data(), (size_t)nR * 28);
      wr(fo, spvo::desc_ptr(dl), (size_t)nL * 1024); wr(fo, spvo::desc_ptr(dr), (size_t)nR * 1024);
      for (int t = 0; t < 2; ++t) {
        const auto& m = fe.cv_DMatches_list[t];
        const auto& map = fe.mapOfIndices((spvo::MatchType)t);
        int n = (t == 1 && f == 0) ? 0 : (int)m.size();
        wr(fo, &n, 4);
        wr(fo, m.data(), (size_t)n * 16);
        std::vector<int> mm(nL, -1);
        if (!(t == 1 && f == 0)) mm.assign(map.begin(), map.end());
        wr(fo, mm.data(), (size_t)nL * 4);
      }
      if (fe.keypoints_dq.size() > 4 || fe.descriptors_dq.size() > 4) return 3;   // NN:500-501
    }
  } catch (const spvo::Error& e) {
    fprintf(stderr, "spvo error %d: %s\n", e.code, e.what());
    return 4;
  }
  fclose(fo);
  fclose(fi);
  return 0;
}
