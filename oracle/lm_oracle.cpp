// oracle/lm_oracle.cpp -- CPU restatement of the LINEMOD match hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing in the product path (6dpose_b200/, the
// C-ABI library, linemodLevelup_pybind) may import, link or call this file.
// Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
// --impl reference legs.
//
// It restates, function by function, what the reference computes on its
// single-threaded SSE path (all citations: linemodLevelup/linemodLevelup.cpp
// of meiqua/6DPose @ 619be57, "LL.cpp"):
//   or_spread            <- spread + orUnaligned8u            LL.cpp:1026-1109
//   response_maps        <- SIMILARITY_LUT + computeResponseMaps   :1121, 1134-1203
//   linearize            <- linearize                          LL.cpp:1215-1243
//   feature_memory       <- accessLinearMemory                 LL.cpp:1248-1271
//   scan16 / scan8       <- similarity / similarity_64         LL.cpp:1284-1354, 1450-1534
//   patch16 / patch8     <- similarityLocal / similarityLocal_64   :1366-1428, 1546-1620
//   sum_modalities       <- addSimilarities(_64)               LL.cpp:1435-1448, 1622-1658
//   match_template       <- Detector::matchClass loop body     LL.cpp:1797-1940
//   lmo_match            <- Detector::match (after quantize)   LL.cpp:1721-1776
// The quantization front-end (LL.cpp:350-505, 729-819) is NOT here: quantized
// label images are an input (see 6dpose_b200/frontend.py, shared by both sides).
//
// Pinning status: checked against the reference's own sources compiled
// unmodified (oracle/_ref, see oracle/ref_shim/) on the reference's fixture
// frame + banks; see tests/test_golden_and_ref.py and tests/golden/.
//
// Like the reference, accumulation uses 128-bit SSE2 adds (8 x u16 or 16 x u8
// per instruction); x86-64 baseline, no -m flags needed.

#include <emmintrin.h>
#include <stdint.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <chrono>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

struct Feat { int x, y, label; };

// One template of a pyramid slot (level*M + modality), LL.h:36-45.
struct Tmpl {
  int width, height;
  const Feat* f;
  int nf;
};

// LL.h:225-258.  class_idx stands in for the class_id string (same equality).
struct Hit {
  int x, y;
  float similarity;
  int class_idx;
  int template_id;
  bool operator<(const Hit& o) const {
    if (similarity != o.similarity) return similarity > o.similarity;
    return template_id < o.template_id;
  }
  bool operator==(const Hit& o) const {
    return x == o.x && y == o.y && similarity == o.similarity && class_idx == o.class_idx;
  }
};

// The active table at LL.cpp:1121 ("1,2-->0 3-->1") is exactly this rule:
// 4 when the template's orientation bit is present in the spread mask, 1 when
// only a neighbouring orientation (o+-1 mod 8) is present, else 0.  The nibble
// split is kept so the lookup has the reference's shape (two 16-entry halves
// per orientation, max of both).  tests/test_oracle_cpu.py re-derives the 256
// bytes from the reference file when it is mounted and compares.
uint8_t g_lut[256];
bool g_lut_ready = false;
void build_lut() {
  if (g_lut_ready) return;
  for (int o = 0; o < 8; ++o) {
    for (int nib = 0; nib < 16; ++nib) {
      for (int half = 0; half < 2; ++half) {
        int v = half ? (nib << 4) : nib;
        int r = 0;
        if ((v >> o) & 1) r = 4;
        else if (((v >> ((o + 1) & 7)) & 1) || ((v >> ((o + 7) & 7)) & 1)) r = 1;
        g_lut[32 * o + 16 * half + nib] = (uint8_t)r;
      }
    }
  }
  g_lut_ready = true;
}

// LL.cpp:1094-1109 + 1026-1083: dst(y,x) = OR of src over the forward TxT
// window, clipped at the bottom/right image edge.
void or_spread(const uint8_t* src, uint8_t* dst, int rows, int cols, int T) {
  memset(dst, 0, (size_t)rows * cols);
  for (int dy = 0; dy < T; ++dy) {
    for (int dx = 0; dx < T; ++dx) {
      const int w = cols - dx, h = rows - dy;
      for (int r = 0; r < h; ++r) {
        const uint8_t* s = src + (size_t)(r + dy) * cols + dx;
        uint8_t* d = dst + (size_t)r * cols;
        int c = 0;
        for (; c + 16 <= w; c += 16) {
          __m128i a = _mm_loadu_si128((const __m128i*)(s + c));
          __m128i b = _mm_loadu_si128((const __m128i*)(d + c));
          _mm_storeu_si128((__m128i*)(d + c), _mm_or_si128(a, b));
        }
        for (; c < w; ++c) d[c] |= s[c];
      }
    }
  }
}

// LL.cpp:1134-1203 (scalar branch; the SSSE3 pshufb branch computes the same).
void response_maps(const uint8_t* spread, uint8_t* maps, int rows, int cols) {
  build_lut();
  const size_t n = (size_t)rows * cols;
  for (int o = 0; o < 8; ++o) {
    const uint8_t* lo = g_lut + 32 * o;
    const uint8_t* hi = lo + 16;
    uint8_t* m = maps + n * o;
    for (size_t i = 0; i < n; ++i) {
      uint8_t v = spread[i];
      m[i] = std::max(lo[v & 15], hi[v >> 4]);
    }
  }
}

// LL.cpp:1215-1243: T*T linear memories of (cols/T)*(rows/T) bytes each, one
// per (row%T, col%T) phase, stored back to back.
void linearize(const uint8_t* map, uint8_t* lin, int rows, int cols, int T) {
  uint8_t* out = lin;
  for (int r0 = 0; r0 < T; ++r0)
    for (int c0 = 0; c0 < T; ++c0)
      for (int r = r0; r < rows; r += T) {
        const uint8_t* row = map + (size_t)r * cols;
        for (int c = c0; c < cols; c += T) *out++ = row[c];
      }
}

// Linear memories of one (level, modality): lm[label] -> T*T*Wd*Hd bytes,
// contiguous per label (the cv::Mat(T*T, Wd*Hd) of LL.cpp:1223).  Separate
// allocations per label, as in the reference (std::vector<Mat>).
struct LinMem {
  std::vector<uint8_t> lab[8];
  int rows, cols, T;
};

void build_linmem(const uint8_t* quantized, int rows, int cols, int T, LinMem& lm) {
  const size_t n = (size_t)rows * cols;
  std::vector<uint8_t> sp(n), maps(n * 8);
  or_spread(quantized, sp.data(), rows, cols, T);
  response_maps(sp.data(), maps.data(), rows, cols);
  lm.rows = rows; lm.cols = cols; lm.T = T;
  for (int o = 0; o < 8; ++o) {
    // +64 B of zero slack: the reference's SSE loads never leave the Mat
    //  but a malformed template must not fault the checker.
    lm.lab[o].assign(n + 64, 0);
    linearize(maps.data() + n * o, lm.lab[o].data(), rows, cols, T);
  }
}

// LL.cpp:1248-1271
inline const uint8_t* feature_memory(const LinMem& lm, const Feat& f, int W) {
  const int T = lm.T;
  const int grid = (f.y % T) * T + (f.x % T);
  const size_t plane = (size_t)(lm.cols / T) * (lm.rows / T);
  return lm.lab[f.label].data() + plane * grid + (size_t)(f.y / T) * W + (f.x / T);
}

inline int scan_positions(const Tmpl& t, int W, int H, int T) {
  // LL.cpp:1299-1309
  const int wf = (t.width - 1) / T + 1;
  const int hf = (t.height - 1) / T + 1;
  return (H - hf) * W + (W - wf) + 1;
}

// LL.cpp:1284-1354: u16 accumulation, 8 positions per SSE2 add.
void scan16(const LinMem& lm, const Tmpl& t, uint16_t* dst, long* byte_adds) {
  const int T = lm.T, W = lm.cols / T, H = lm.rows / T;
  memset(dst, 0, sizeof(uint16_t) * (size_t)W * H);
  const int P = scan_positions(t, W, H, T);
  const __m128i zero = _mm_setzero_si128();
  for (int i = 0; i < t.nf; ++i) {
    const Feat f = t.f[i];
    if (f.x < 0 || f.x >= lm.cols || f.y < 0 || f.y >= lm.rows) continue;
    const uint8_t* p = feature_memory(lm, f, W);
    int j = 0;
    for (; j < P - 7; j += 8) {
      __m128i r = _mm_loadl_epi64((const __m128i*)(p + j));
      r = _mm_unpacklo_epi8(r, zero);
      __m128i* d = (__m128i*)(dst + j);
      _mm_storeu_si128(d, _mm_add_epi16(_mm_loadu_si128(d), r));
    }
    for (; j < P; ++j) dst[j] = (uint16_t)(dst[j] + p[j]);
    if (P > 0) *byte_adds += P;
  }
}

// LL.cpp:1450-1534: u8 accumulation (<= 63 features), 16 positions per add.
void scan8(const LinMem& lm, const Tmpl& t, uint8_t* dst, long* byte_adds) {
  const int T = lm.T, W = lm.cols / T, H = lm.rows / T;
  memset(dst, 0, (size_t)W * H);
  const int P = scan_positions(t, W, H, T);
  for (int i = 0; i < t.nf; ++i) {
    const Feat f = t.f[i];
    if (f.x < 0 || f.x >= lm.cols || f.y < 0 || f.y >= lm.rows) continue;
    const uint8_t* p = feature_memory(lm, f, W);
    int j = 0;
    for (; j < P - 15; j += 16) {
      __m128i r = _mm_loadu_si128((const __m128i*)(p + j));
      __m128i* d = (__m128i*)(dst + j);
      _mm_storeu_si128(d, _mm_add_epi8(_mm_loadu_si128(d), r));
    }
    for (; j < P; ++j) dst[j] = (uint8_t)(dst[j] + p[j]);
    if (P > 0) *byte_adds += P;
  }
}

// LL.cpp:1366-1428: 16x16 patch, u16, two 8-lane adds per row.
void patch16(const LinMem& lm, const Tmpl& t, uint16_t* dst /*256*/, int cx, int cy, long* byte_adds) {
  const int T = lm.T, W = lm.cols / T;
  memset(dst, 0, 256 * sizeof(uint16_t));
  const int ox = (cx / T - 8) * T;  // C++ truncating division, LL.cpp:1380-1381
  const int oy = (cy / T - 8) * T;
  const __m128i zero = _mm_setzero_si128();
  for (int i = 0; i < t.nf; ++i) {
    Feat f = t.f[i];
    f.x += ox; f.y += oy;
    if (f.x < 0 || f.y < 0 || f.x >= lm.cols || f.y >= lm.rows) continue;
    const uint8_t* p = feature_memory(lm, f, W);
    for (int row = 0; row < 16; ++row) {
      __m128i lo = _mm_unpacklo_epi8(_mm_loadl_epi64((const __m128i*)p), zero);
      __m128i hi = _mm_unpacklo_epi8(_mm_loadl_epi64((const __m128i*)(p + 8)), zero);
      __m128i* d = (__m128i*)(dst + 16 * row);
      _mm_storeu_si128(d, _mm_add_epi16(_mm_loadu_si128(d), lo));
      _mm_storeu_si128(d + 1, _mm_add_epi16(_mm_loadu_si128(d + 1), hi));
      p += W;
    }
    *byte_adds += 256;
  }
}

// LL.cpp:1546-1620: 16x16 patch, u8.
void patch8(const LinMem& lm, const Tmpl& t, uint8_t* dst /*256*/, int cx, int cy, long* byte_adds) {
  const int T = lm.T, W = lm.cols / T;
  memset(dst, 0, 256);
  const int ox = (cx / T - 8) * T;
  const int oy = (cy / T - 8) * T;
  for (int i = 0; i < t.nf; ++i) {
    Feat f = t.f[i];
    f.x += ox; f.y += oy;
    if (f.x < 0 || f.y < 0 || f.x >= lm.cols || f.y >= lm.rows) continue;
    const uint8_t* p = feature_memory(lm, f, W);
    for (int row = 0; row < 16; ++row) {
      __m128i* d = (__m128i*)(dst + 16 * row);
      _mm_storeu_si128(d, _mm_add_epi8(_mm_loadu_si128(d), _mm_loadu_si128((const __m128i*)p)));
      p += W;
    }
    *byte_adds += 256;
  }
}

struct alignas(64) Stats {  // one cache line per thread
  long coarse_byte_adds = 0, refine_byte_adds = 0, coarse_candidates = 0;
};

// Which accumulator width the reference picks: decided by the first modality
// whose count is < 64 (-> 8 bit) or < 8192 (-> 16 bit); sticks afterwards.
// LL.cpp:1813-1819 / 1889-1895.  Returns 0 when nothing is computed.
inline int pick_width(int current, int nf) {
  if (current > 0) return current;
  if (nf < 64) return 1;
  if (nf < 8192) return 2;
  return current;
}

// Body of the template loop in Detector::matchClass, LL.cpp:1797-1940.
// Returns <0 where the reference would raise (CV_Assert in similarity*).
int match_template(const std::vector<std::vector<LinMem>>& pyr, const int* T_at_level, int L, int M,
                   const Tmpl* tp /*[L*M]*/, float threshold, int class_idx, int template_id,
                   std::vector<Hit>& out, Stats& st) {
  const std::vector<LinMem>& low = pyr[L - 1];
  const int lowT = T_at_level[L - 1];
  const int W = low[0].cols / lowT, H = low[0].rows / lowT;
  const int start_low = (L - 1) * M;

  std::vector<uint16_t> total((size_t)W * H, 0);
  std::vector<uint16_t> s16((size_t)W * H);
  std::vector<uint8_t> s8((size_t)W * H);
  int nfeat = 0, width = -1;
  bool computed = false;
  for (int m = 0; m < M; ++m) {
    const Tmpl& t = tp[start_low + m];
    nfeat += t.nf;
    width = pick_width(width, t.nf);
    if (width == 1) {
      if (t.nf > 63) return -1;  // CV_Assert LL.cpp:1457
      scan8(low[m], t, s8.data(), &st.coarse_byte_adds);
      for (size_t i = 0; i < total.size(); ++i) total[i] = (uint16_t)(total[i] + s8[i]);
      computed = true;
    } else if (width == 2) {
      if (t.nf > 8191) return -1;  // CV_Assert LL.cpp:1291
      scan16(low[m], t, s16.data(), &st.coarse_byte_adds);
      // cv::add on CV_16U saturates (LL.cpp:1445); 2*4*8191 < 65535 never does.
      for (size_t i = 0; i < total.size(); ++i) {
        unsigned v = (unsigned)total[i] + s16[i];
        total[i] = (uint16_t)(v > 65535u ? 65535u : v);
      }
      computed = true;
    }
  }

  std::vector<Hit> cand;
  if (computed) {  // an empty total_similarity has no rows, LL.cpp:1836
    const int off = lowT / 2 + (lowT % 2 - 1);
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < W; ++c) {
        const int raw = total[(size_t)r * W + c];
        const float score = (raw * 100.f) / (4 * nfeat);
        if (score > threshold) cand.push_back(Hit{c * lowT + off, r * lowT + off, score, class_idx, template_id});
      }
  }
  st.coarse_candidates += (long)cand.size();

  uint16_t p16[256], tot[256];
  uint8_t p8[256];
  for (int l = L - 2; l >= 0; --l) {
    const std::vector<LinMem>& lms = pyr[l];
    const int T = T_at_level[l];
    const int start = l * M;
    const int border = 8 * T;
    const int off = T / 2 + (T % 2 - 1);
    const int max_x = lms[0].cols - tp[start].width - border;
    const int max_y = lms[0].rows - tp[start].height - border;
    for (size_t k = 0; k < cand.size(); ++k) {
      Hit& h = cand[k];
      int x = h.x * 2 + 1, y = h.y * 2 + 1;
      x = std::max(x, border); y = std::max(y, border);
      x = std::min(x, max_x);  y = std::min(y, max_y);
      int nf2 = 0, w2 = -1;
      bool any = false;
      memset(tot, 0, sizeof(tot));
      for (int m = 0; m < M; ++m) {
        const Tmpl& t = tp[start + m];
        nf2 += t.nf;
        w2 = pick_width(w2, t.nf);
        if (w2 == 1) {
          if (t.nf > 63) return -1;
          patch8(lms[m], t, p8, x, y, &st.refine_byte_adds);
          for (int i = 0; i < 256; ++i) tot[i] = (uint16_t)(tot[i] + p8[i]);
          any = true;
        } else if (w2 == 2) {
          if (t.nf > 8191) return -1;
          patch16(lms[m], t, p16, x, y, &st.refine_byte_adds);
          for (int i = 0; i < 256; ++i) tot[i] = (uint16_t)(tot[i] + p16[i]);
          any = true;
        }
      }
      float best = 0;
      int br = -1, bc = -1;
      if (any)
        for (int r = 0; r < 16; ++r)
          for (int c = 0; c < 16; ++c) {
            const int raw = tot[16 * r + c];
            const float score = (raw * 100.f) / (4 * nf2);
            if (score > best) { best = score; br = r; bc = c; }
          }
      h.similarity = best;
      h.x = (x / T - 8 + bc) * T + off;
      h.y = (y / T - 8 + br) * T + off;
    }
    cand.erase(std::remove_if(cand.begin(), cand.end(),
                              [threshold](const Hit& h) { return h.similarity < threshold; }),
               cand.end());
  }
  out.insert(out.end(), cand.begin(), cand.end());
  return 0;
}

}  // namespace

extern "C" {

struct lmo_match_rec {
  int32_t x, y;
  float similarity;
  int32_t class_idx;
  int32_t template_id;
};

void lmo_spread(const uint8_t* src, uint8_t* dst, int rows, int cols, int T) { or_spread(src, dst, rows, cols, T); }

void lmo_response_maps(const uint8_t* spread, uint8_t* maps, int rows, int cols) { response_maps(spread, maps, rows, cols); }

void lmo_linearize(const uint8_t* map, uint8_t* lin, int rows, int cols, int T) { linearize(map, lin, rows, cols, T); }

void lmo_similarity_lut(uint8_t* out256) { build_lut(); memcpy(out256, g_lut, 256); }

// quantized u8 rows x cols -> lm[8][T*T][(cols/T)*(rows/T)].  -1 on the
// reference's size assertions (LL.cpp:1136, 1217-1218).
int lmo_linear_memories(const uint8_t* quantized, int rows, int cols, int T, uint8_t* lm_out) {
  if (T <= 0 || rows % T || cols % T || ((long)rows * cols) % 16) return -1;
  LinMem lm;
  build_linmem(quantized, rows, cols, T, lm);
  const size_t n = (size_t)rows * cols;
  for (int o = 0; o < 8; ++o) memcpy(lm_out + n * o, lm.lab[o].data(), n);
  return 0;
}

// Detector::match after quantization.
//   L levels, M modalities; T[l], rows[l], cols[l]; quantized[l*M+m] -> u8 rows[l] x cols[l]
//   bank: class_begin[n_classes+1] (global template ranges, in match order),
//         tmeta[g][L*M][4] = {width, height, feat_begin, feat_count}, feats[][3] = {x, y, label}
//   out: up to cap records, final order (after std::sort + std::unique).
//   stats[8] (optional): coarse_byte_adds, refine_byte_adds, coarse_candidates, pre-unique matches,
//                        t_linmem_us, t_match_us, t_sort_us, threads used
// Returns the number of matches (may exceed cap: only cap are written), or <0 on error.
static long match_impl(int L, int M, const int* T, const int* rows, const int* cols, const uint8_t* const* quantized,
                       int n_classes, const int* class_begin, const int32_t* tmeta, const int32_t* feats,
                       float threshold, int n_threads, lmo_match_rec* out, long cap, double* stats, bool finish) {
  typedef std::chrono::steady_clock clk;
  for (int l = 0; l < L; ++l)
    if (T[l] <= 0 || rows[l] % T[l] || cols[l] % T[l] || ((long)rows[l] * cols[l]) % 16) return -2;
  auto t0 = clk::now();
  std::vector<std::vector<LinMem>> pyr(L, std::vector<LinMem>(M));
  build_lut();
  int used_threads = 1;
#ifdef _OPENMP
  if (n_threads > 1) used_threads = n_threads;
#pragma omp parallel for schedule(dynamic, 1) num_threads(used_threads) if (used_threads > 1)
#endif
  for (int lm = 0; lm < L * M; ++lm)
    build_linmem(quantized[lm], rows[lm / M], cols[lm / M], T[lm / M], pyr[lm / M][lm % M]);
  auto t1 = clk::now();

  const int S = L * M;
  const int G = class_begin[n_classes];
  std::vector<int> class_of(G);
  for (int c = 0; c < n_classes; ++c)
    for (int g = class_begin[c]; g < class_begin[c + 1]; ++g) class_of[g] = c;

  std::vector<std::vector<Hit>> per(G);
  std::vector<Stats> stt(used_threads);
  int err = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 8) num_threads(used_threads) if (used_threads > 1)
#endif
  for (int g = 0; g < G; ++g) {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
    if (g == 0 && getenv("LMO_DEBUG")) fprintf(stderr, "omp threads in region: %d\n", omp_get_num_threads());
#endif
    std::vector<Tmpl> tp(S);
    for (int s = 0; s < S; ++s) {
      const int32_t* m4 = tmeta + ((size_t)g * S + s) * 4;
      tp[s].width = m4[0]; tp[s].height = m4[1];
      tp[s].f = (const Feat*)(feats + (size_t)m4[2] * 3);
      tp[s].nf = m4[3];
    }
    const int c = class_of[g];
    Stats local;
    int rc = match_template(pyr, T, L, M, tp.data(), threshold, c, g - class_begin[c], per[g], local);
    stt[tid].coarse_byte_adds += local.coarse_byte_adds;
    stt[tid].refine_byte_adds += local.refine_byte_adds;
    stt[tid].coarse_candidates += local.coarse_candidates;
    if (rc < 0) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
      err = rc;
    }
  }
  if (err) return err;
  std::vector<Hit> all;
  for (int g = 0; g < G; ++g) all.insert(all.end(), per[g].begin(), per[g].end());
  auto t2 = clk::now();
  const size_t pre_unique = all.size();
  if (finish) {  // LL.cpp:1772-1774
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
  }
  auto t3 = clk::now();

  const long n = (long)all.size();
  for (long i = 0; i < n && i < cap; ++i) {
    out[i].x = all[i].x; out[i].y = all[i].y; out[i].similarity = all[i].similarity;
    out[i].class_idx = all[i].class_idx; out[i].template_id = all[i].template_id;
  }
  if (stats) {
    Stats s;
    for (auto& q : stt) {
      s.coarse_byte_adds += q.coarse_byte_adds; s.refine_byte_adds += q.refine_byte_adds;
      s.coarse_candidates += q.coarse_candidates;
    }
    auto us = [](clk::time_point a, clk::time_point b) {
      return std::chrono::duration<double, std::micro>(b - a).count();
    };
    stats[0] = (double)s.coarse_byte_adds; stats[1] = (double)s.refine_byte_adds;
    stats[2] = (double)s.coarse_candidates; stats[3] = (double)pre_unique;
    stats[4] = us(t0, t1); stats[5] = us(t1, t2); stats[6] = us(t2, t3); stats[7] = used_threads;
  }
  return n;
}

long lmo_match(int L, int M, const int* T, const int* rows, const int* cols, const uint8_t* const* quantized,
               int n_classes, const int* class_begin, const int32_t* tmeta, const int32_t* feats,
               float threshold, int n_threads, lmo_match_rec* out, long cap, double* stats) {
  return match_impl(L, M, T, rows, cols, quantized, n_classes, class_begin, tmeta, feats, threshold, n_threads, out, cap, stats, true);
}

// The same, stopped before std::sort/std::unique: the matches in the reference's pre-sort order
// (class order -> template_id -> ascending coarse cell).  Used by the sharding tests.
long lmo_match_presort(int L, int M, const int* T, const int* rows, const int* cols, const uint8_t* const* quantized,
                       int n_classes, const int* class_begin, const int32_t* tmeta, const int32_t* feats,
                       float threshold, int n_threads, lmo_match_rec* out, long cap, double* stats) {
  return match_impl(L, M, T, rows, cols, quantized, n_classes, class_begin, tmeta, feats, threshold, n_threads, out, cap, stats, false);
}

// Debug/inspection: the coarse total_similarity map (u16, Hd*Wd) of one
// template, plus its template_positions P.  Returns P or <0.
long lmo_coarse_map(int M, int T, int rows, int cols, const uint8_t* const* quantized /*[M]*/,
                    const int32_t* tmeta_low /*[M][4]*/, const int32_t* feats, uint16_t* out_map) {
  if (T <= 0 || rows % T || cols % T || ((long)rows * cols) % 16) return -2;
  const int W = cols / T, H = rows / T;
  std::vector<uint16_t> s16((size_t)W * H);
  memset(out_map, 0, sizeof(uint16_t) * (size_t)W * H);
  long adds = 0, P = 0;
  for (int m = 0; m < M; ++m) {
    LinMem lm;
    build_linmem(quantized[m], rows, cols, T, lm);
    Tmpl t{tmeta_low[4 * m + 0], tmeta_low[4 * m + 1], (const Feat*)(feats + (size_t)tmeta_low[4 * m + 2] * 3),
           tmeta_low[4 * m + 3]};
    if (t.nf > 8191) return -1;
    scan16(lm, t, s16.data(), &adds);
    for (size_t i = 0; i < s16.size(); ++i) out_map[i] = (uint16_t)(out_map[i] + s16[i]);
    P = scan_positions(t, W, H, T);
  }
  return P;
}

int lmo_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

}  // extern "C"
