// ref_driver.cpp -- builds oracle/_ref/liblm_ref.so: the reference's OWN linemodLevelup.cpp, compiled
// unmodified from where it lies under /root/reference (never copied), against oracle/ref_shim's
// stand-in for the cv::Mat container.  TEST INFRASTRUCTURE ONLY.
//
// What runs is the reference's code for Detector::match and everything below it -- spread,
// orUnaligned8u, computeResponseMaps, linearize, similarity(_64), similarityLocal(_64),
// addSimilarities(_64), matchClass, std::sort/std::unique (LL.cpp:1026-1941) -- driven through the
// reference's own extension points: a Modality whose QuantizedPyramid hands back pre-quantized label
// images (the front-end is shared and lives outside the compiled path), and the protected template
// map filled directly (template IO bypassed).  Same C entry point as oracle/lm_oracle.cpp's lmo_match.
#include <stdint.h>

#include <chrono>
#include <string>
#include <vector>

#include "linemodLevelup.cpp"  // resolved through -I$(REF): the reference translation unit itself

namespace {

class PreQuantized : public linemodLevelup::QuantizedPyramid {
 public:
  std::vector<cv::Mat> levels;
  int level = 0;
  void quantize(cv::Mat& dst) const override { dst = levels[level].clone(); }
  bool extractTemplate(linemodLevelup::Template&) const override { return false; }
  void pyrDown() override { ++level; }
};

class PreQuantizedModality : public linemodLevelup::Modality {
 public:
  std::string label;
  std::vector<cv::Mat> levels;
  std::string name() const override { return label; }
  void read(const cv::FileNode&) override {}
  void write(cv::FileStorage&) const override {}

 protected:
  cv::Ptr<linemodLevelup::QuantizedPyramid> processImpl(const std::vector<cv::Mat>&, const cv::Mat&) const override {
    auto p = cv::makePtr<PreQuantized>();
    p->levels = levels;
    return p;
  }
};

class OpenDetector : public linemodLevelup::Detector {
 public:
  OpenDetector(const std::vector<cv::Ptr<linemodLevelup::Modality>>& m, const std::vector<int>& T) : Detector(m, T) {}
  TemplatesMap& templates() { return class_templates; }
};

}  // namespace

extern "C" {

struct lmr_match_rec {
  int32_t x, y;
  float similarity;
  int32_t class_idx;
  int32_t template_id;
};

// Same contract as lmo_match (oracle/lm_oracle.cpp).  The reference has no threading: n_threads is
// accepted and ignored (stats[7] reports 1).  Returns the number of matches or <0 if the reference
// throws (cv::Exception from a CV_Assert).
long lmr_match(int L, int M, const int* T, const int* rows, const int* cols, const uint8_t* const* quantized,
               int n_classes, const int* class_begin, const int32_t* tmeta, const int32_t* feats,
               float threshold, int /*n_threads*/, lmr_match_rec* out, long cap, double* stats) {
  typedef std::chrono::steady_clock clk;
  try {
    std::vector<cv::Ptr<linemodLevelup::Modality>> mods;
    for (int m = 0; m < M; ++m) {
      auto pm = cv::makePtr<PreQuantizedModality>();
      pm->label = m == 0 ? "ColorGradient" : "DepthNormal";
      for (int l = 0; l < L; ++l) {
        cv::Mat q(rows[l], cols[l], CV_8U);
        memcpy(q.data, quantized[l * M + m], (size_t)rows[l] * cols[l]);
        pm->levels.push_back(q);
      }
      mods.push_back(pm);
    }
    OpenDetector det(mods, std::vector<int>(T, T + L));
    const int S = L * M;
    std::vector<std::string> ids;
    char name[32];
    for (int c = 0; c < n_classes; ++c) {
      snprintf(name, sizeof(name), "class_%06d", c);  // std::map order == index order
      ids.push_back(name);
      auto& tps = det.templates()[name];
      for (int g = class_begin[c]; g < class_begin[c + 1]; ++g) {
        std::vector<linemodLevelup::Template> tp(S);
        for (int s = 0; s < S; ++s) {
          const int32_t* m4 = tmeta + ((size_t)g * S + s) * 4;
          tp[s].width = m4[0];
          tp[s].height = m4[1];
          tp[s].pyramid_level = s / M;
          tp[s].features.resize(m4[3]);
          for (int k = 0; k < m4[3]; ++k) {
            const int32_t* f = feats + ((size_t)m4[2] + k) * 3;
            tp[s].features[k] = linemodLevelup::Feature(f[0], f[1], f[2]);
          }
        }
        tps.push_back(tp);
      }
    }
    std::vector<cv::Mat> sources(M);  // only their count is looked at (LL.cpp:1707)
    auto t0 = clk::now();
    std::vector<linemodLevelup::Match> res = det.match(sources, threshold, ids, std::vector<cv::Mat>());
    auto t1 = clk::now();
    const long n = (long)res.size();
    for (long i = 0; i < n && i < cap; ++i) {
      out[i].x = res[i].x;
      out[i].y = res[i].y;
      out[i].similarity = res[i].similarity;
      out[i].class_idx = atoi(res[i].class_id.c_str() + 6);
      out[i].template_id = res[i].template_id;
    }
    if (stats) {
      for (int i = 0; i < 8; ++i) stats[i] = 0;
      stats[5] = std::chrono::duration<double, std::micro>(t1 - t0).count();  // whole Detector::match
      stats[7] = 1;
    }
    return n;
  } catch (const cv::Exception& e) {
    return -1;
  }
}

// The reference's active SIMILARITY_LUT (LL.cpp:1121), for the table test.
void lmr_similarity_lut(uint8_t* out256) { memcpy(out256, linemodLevelup::SIMILARITY_LUT, 256); }

}  // extern "C"
