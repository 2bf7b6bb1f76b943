// cv_shim.h -- a minimal stand-in for the parts of the OpenCV C++ API that the reference's
// linemodLevelup.cpp touches, so that the reference's OWN source file can be compiled where it lies
// (this container has no OpenCV C++ headers / libs).  TEST INFRASTRUCTURE ONLY (oracle/_ref).
//
// What is real: cv::Mat as a ref-counted, continuous, 64-byte aligned 2-D container with the element
// access, zeros/create/clone/copyTo/convertTo/+= the matching path uses (spread, computeResponseMaps,
// linearize, similarity*, addSimilarities*, matchClass, match: LL.cpp:1026-1941).
// What is a stub: image filters (GaussianBlur, Sobel, phase, pyrDown, resize, medianBlur, erode,
// dilate, distanceTransform ...) and FileStorage -- they belong to the quantization front-end and
// template IO, which oracle/ref_shim/ref_driver.cpp bypasses; reaching one aborts loudly.
#pragma once

#include <emmintrin.h>
#ifdef __SSE3__
#include <pmmintrin.h>
#endif
#ifdef __SSSE3__
#include <tmmintrin.h>
#endif
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

// OpenCV derives these from the compiler's target flags (cvdef.h); linemodLevelup/CMakeLists.txt:8
// builds with plain -O3, i.e. SSE2 only on x86-64.
#define CV_SSE2 1
#ifdef __SSE3__
#define CV_SSE3 1
#else
#define CV_SSE3 0
#endif
#ifdef __SSSE3__
#define CV_SSSE3 1
#else
#define CV_SSSE3 0
#endif

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_MAT_DEPTH(t) ((t)&7)
#define CV_MAT_CN(t) ((((t) >> 3) & 511) + 1)
#define CV_DECL_ALIGNED(x) __attribute__((aligned(x)))
#define CV_CPU_SSE2 2
#define CV_CPU_SSE3 3
#define CV_CPU_SSSE3 4
#define CPU_SSE2 CV_CPU_SSE2
#define CPU_SSE3 CV_CPU_SSE3
#define CPU_SSSE3 CV_CPU_SSSE3

namespace cv {

typedef unsigned char uchar;
typedef unsigned short ushort;
typedef std::string String;

class Exception : public std::runtime_error {
 public:
  explicit Exception(const std::string& m) : std::runtime_error(m) {}
};
namespace Error { enum { StsBadArg = -5, StsAssert = -215 }; }

#define CV_Error(code, msg) throw cv::Exception(std::string(msg))
#define CV_Assert(expr) \
  do { if (!(expr)) throw cv::Exception(std::string("CV_Assert failed: ") + #expr); } while (0)
#define CV_DbgAssert(expr) ((void)0)

[[noreturn]] inline void shim_unreachable(const char* what) {
  fprintf(stderr, "oracle/ref_shim: %s is not part of the compiled match path (front-end / IO stub reached)\n", what);
  abort();
}

inline bool checkHardwareSupport(int) { return true; }

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
  bool operator==(const Size& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size& o) const { return !(*this == o); }
};
struct Point {
  int x, y;
  Point() : x(0), y(0) {}
  Point(int x_, int y_) : x(x_), y(y_) {}
};
struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  static Scalar all(double v) { return Scalar(v, v, v, v); }
};

template <typename T> struct DataType;
template <> struct DataType<uchar> { enum { type = CV_8U }; };
template <> struct DataType<ushort> { enum { type = CV_16U }; };
template <> struct DataType<short> { enum { type = CV_16S }; };
template <> struct DataType<int> { enum { type = CV_32S }; };
template <> struct DataType<float> { enum { type = CV_32F }; };
template <> struct DataType<double> { enum { type = CV_64F }; };

template <typename T> using Ptr = std::shared_ptr<T>;
template <typename T, typename... A> Ptr<T> makePtr(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }

inline size_t depth_size(int type) {
  static const size_t s[7] = {1, 1, 2, 2, 4, 4, 8};
  return s[CV_MAT_DEPTH(type)];
}

class Mat {
 public:
  int rows = 0, cols = 0;
  uchar* data = nullptr;

  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(Size s, int type) { create(s.height, s.width, type); }
  Mat(int r, int c, int type, const Scalar& v) { create(r, c, type); fill(v.val[0]); }
  Mat(int r, int c, int type, void* ext, size_t step_bytes = 0) {  // borrowed data
    rows = r; cols = c; type_ = type; data = (uchar*)ext;
    step_ = step_bytes ? step_bytes : (size_t)c * elemSize();
  }

  static Mat zeros(int r, int c, int type) { Mat m(r, c, type); memset(m.data, 0, m.step_ * (size_t)r); return m; }
  static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
  static Mat ones(int r, int c, int type) { Mat m(r, c, type); m.fill(1.0); return m; }

  void create(int r, int c, int type) {
    if (data && r == rows && c == cols && type == type_ && owner_) return;
    rows = r; cols = c; type_ = type;
    step_ = (size_t)c * elemSize();
    const size_t bytes = std::max<size_t>(step_ * (size_t)r, 1) + 64;
    void* p = nullptr;
    if (posix_memalign(&p, 64, bytes) != 0) throw std::bad_alloc();
    owner_ = std::shared_ptr<uchar>((uchar*)p, free);
    data = owner_.get();
  }
  void create(Size s, int type) { create(s.height, s.width, type); }

  int type() const { return type_; }
  int depth() const { return CV_MAT_DEPTH(type_); }
  int channels() const { return CV_MAT_CN(type_); }
  size_t elemSize() const { return depth_size(type_) * (size_t)CV_MAT_CN(type_); }
  size_t elemSize1() const { return depth_size(type_); }
  size_t step1() const { return step_ / elemSize1(); }
  size_t total() const { return (size_t)rows * cols; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  bool isContinuous() const { return step_ == (size_t)cols * elemSize(); }
  Size size() const { return Size(cols, rows); }

  uchar* ptr(int r = 0) { return data + step_ * (size_t)r; }
  const uchar* ptr(int r = 0) const { return data + step_ * (size_t)r; }
  template <typename T> T* ptr(int r = 0) { return (T*)(data + step_ * (size_t)r); }
  template <typename T> const T* ptr(int r = 0) const { return (const T*)(data + step_ * (size_t)r); }
  template <typename T> T& at(int r, int c) { return ((T*)(data + step_ * (size_t)r))[c]; }
  template <typename T> const T& at(int r, int c) const { return ((const T*)(data + step_ * (size_t)r))[c]; }

  Mat clone() const {
    Mat m;
    if (empty()) return m;
    m.create(rows, cols, type_);
    for (int r = 0; r < rows; ++r) memcpy(m.ptr(r), ptr(r), (size_t)cols * elemSize());
    return m;
  }
  void copyTo(Mat& dst) const { dst = clone(); }
  void copyTo(Mat&& roi) const {  // into a region of interest
    for (int r = 0; r < rows; ++r) memcpy(roi.ptr(r), ptr(r), (size_t)cols * elemSize());
  }
  void copyTo(Mat& dst, const Mat& mask) const {
    if (mask.empty()) { dst = clone(); return; }
    if (dst.empty() || dst.rows != rows || dst.cols != cols || dst.type() != type_) dst = zeros(rows, cols, type_);
    const size_t es = elemSize();
    for (int r = 0; r < rows; ++r) {
      const uchar* mk = mask.ptr(r);
      for (int c = 0; c < cols; ++c)
        if (mk[c]) memcpy(dst.ptr(r) + es * c, ptr(r) + es * c, es);
    }
  }
  Mat operator()(const Rect& rc) const {  // shares storage
    Mat m;
    m.rows = rc.height; m.cols = rc.width; m.type_ = type_; m.step_ = step_; m.owner_ = owner_;
    m.data = data + step_ * (size_t)rc.y + elemSize() * (size_t)rc.x;
    return m;
  }
  Mat& setTo(const Scalar& v, const Mat& mask = Mat()) {
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c)
        if (mask.empty() || mask.at<uchar>(r, c)) set(r, c, v.val[0]);
    return *this;
  }
  double get(int r, int c) const {
    switch (depth()) {
      case CV_8U: return at<uchar>(r, c);
      case CV_8S: return at<signed char>(r, c);
      case CV_16U: return at<ushort>(r, c);
      case CV_16S: return at<short>(r, c);
      case CV_32S: return at<int>(r, c);
      case CV_32F: return at<float>(r, c);
      default: return at<double>(r, c);
    }
  }
  void set(int r, int c, double v) {
    switch (depth()) {
      case CV_8U: at<uchar>(r, c) = (uchar)std::min(255.0, std::max(0.0, std::nearbyint(v))); break;
      case CV_8S: at<signed char>(r, c) = (signed char)v; break;
      case CV_16U: at<ushort>(r, c) = (ushort)std::min(65535.0, std::max(0.0, std::nearbyint(v))); break;
      case CV_16S: at<short>(r, c) = (short)v; break;
      case CV_32S: at<int>(r, c) = (int)v; break;
      case CV_32F: at<float>(r, c) = (float)v; break;
      default: at<double>(r, c) = v;
    }
  }
  void fill(double v) {
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols * channels(); ++c) set1(r, c, v);
  }
  // Mat::convertTo(dst, rtype, alpha): saturate_cast<rtype>(src * alpha), single channel
  void convertTo(Mat& dst, int rtype, double alpha = 1.0, double beta = 0.0) const {
    Mat out(rows, cols, CV_MAKETYPE(CV_MAT_DEPTH(rtype), channels()));
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) out.set(r, c, get(r, c) * alpha + beta);
    dst = out;
  }
  // cv::add semantics: saturating for integer depths (LL.cpp:1445 on CV_16U)
  Mat& operator+=(const Mat& o) {
    CV_Assert(o.rows == rows && o.cols == cols);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) set(r, c, get(r, c) + o.get(r, c));
    return *this;
  }
  Mat operator>(double v) const {
    Mat m = zeros(rows, cols, CV_8U);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) m.at<uchar>(r, c) = get(r, c) > v ? 255 : 0;
    return m;
  }
  Mat operator*(double v) const {
    Mat m = clone();
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) m.set(r, c, get(r, c) * v);
    return m;
  }
  Mat t() const {
    Mat m(cols, rows, type_);
    for (int r = 0; r < rows; ++r)
      for (int c = 0; c < cols; ++c) memcpy(m.ptr(c) + elemSize() * r, ptr(r) + elemSize() * c, elemSize());
    return m;
  }

 private:
  void set1(int r, int c, double v) {  // per scalar component
    switch (depth()) {
      case CV_8U: ((uchar*)ptr(r))[c] = (uchar)v; break;
      case CV_16U: ((ushort*)ptr(r))[c] = (ushort)v; break;
      case CV_16S: ((short*)ptr(r))[c] = (short)v; break;
      case CV_32S: ((int*)ptr(r))[c] = (int)v; break;
      case CV_32F: ((float*)ptr(r))[c] = (float)v; break;
      case CV_64F: ((double*)ptr(r))[c] = v; break;
      default: ((signed char*)ptr(r))[c] = (signed char)v;
    }
  }
  int type_ = 0;
  size_t step_ = 0;
  std::shared_ptr<uchar> owner_;
};

template <typename T> class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
  Mat_(const Mat& m) : Mat(m) {}
  Mat_& operator=(const Mat& m) { Mat::operator=(m); return *this; }
  T& operator()(int r, int c) { return this->template at<T>(r, c); }
  const T& operator()(int r, int c) const { return this->template at<T>(r, c); }
  T* ptr(int r = 0) { return Mat::ptr<T>(r); }
  const T* ptr(int r = 0) const { return Mat::ptr<T>(r); }
  template <typename U> U* ptr(int r = 0) { return Mat::ptr<U>(r); }
};

struct NoArray {};
inline NoArray noArray() { return NoArray(); }

enum { BORDER_REPLICATE = 1, INTER_NEAREST = 0, DIST_C = 3 };

// ---- front-end / training / IO: outside the compiled match path -------------------------------
inline void GaussianBlur(const Mat&, Mat&, Size, double, double = 0, int = 0) { shim_unreachable("cv::GaussianBlur"); }
inline void Sobel(const Mat&, Mat&, int, int, int, int = 3, double = 1, double = 0, int = 0) { shim_unreachable("cv::Sobel"); }
inline void phase(const Mat&, const Mat&, Mat&, bool = false) { shim_unreachable("cv::phase"); }
inline void pyrDown(const Mat&, Mat&, const Size& = Size()) { shim_unreachable("cv::pyrDown"); }
inline void resize(const Mat&, Mat&, Size, double = 0, double = 0, int = 0) { shim_unreachable("cv::resize"); }
inline void medianBlur(const Mat&, Mat&, int) { shim_unreachable("cv::medianBlur"); }
inline void erode(const Mat&, Mat&, const Mat&, Point = Point(-1, -1), int = 1, int = 0) { shim_unreachable("cv::erode"); }
inline void dilate(const Mat&, Mat&, const Mat&, Point = Point(-1, -1), int = 1, int = 0) { shim_unreachable("cv::dilate"); }
inline void subtract(const Mat&, const Mat&, Mat&) { shim_unreachable("cv::subtract"); }
inline void bitwise_and(const Mat&, const Mat&, Mat&) { shim_unreachable("cv::bitwise_and"); }
inline void distanceTransform(const Mat&, Mat&, int, int) { shim_unreachable("cv::distanceTransform"); }
inline int countNonZero(const Mat&) { shim_unreachable("cv::countNonZero"); }
inline void findNonZero(const Mat&, Mat&) { shim_unreachable("cv::findNonZero"); }
inline Rect boundingRect(const Mat&) { shim_unreachable("cv::boundingRect"); }
inline void transpose(const Mat& s, Mat& d) { d = s.t(); }
inline void add(const Mat&, const Mat&, Mat&, NoArray = NoArray(), int = -1) { shim_unreachable("cv::add"); }
inline String format(const char* fmt, const char* a) {
  char buf[4096];
  snprintf(buf, sizeof(buf), fmt, a);
  return String(buf);
}

class FileNode;
class FileNodeIterator {
 public:
  FileNode operator*() const;
  FileNodeIterator& operator++() { shim_unreachable("cv::FileNodeIterator"); }
  bool operator!=(const FileNodeIterator&) const { shim_unreachable("cv::FileNodeIterator"); }
};
inline FileNodeIterator& operator>>(FileNodeIterator&, int&) { shim_unreachable("cv::FileNodeIterator"); }
class FileNode {
 public:
  FileNode operator[](const char*) const { shim_unreachable("cv::FileNode"); }
  operator int() const { shim_unreachable("cv::FileNode"); }
  operator float() const { shim_unreachable("cv::FileNode"); }
  operator std::string() const { shim_unreachable("cv::FileNode"); }
  size_t size() const { shim_unreachable("cv::FileNode"); }
  FileNodeIterator begin() const { shim_unreachable("cv::FileNode"); }
  FileNodeIterator end() const { shim_unreachable("cv::FileNode"); }
};
inline FileNode FileNodeIterator::operator*() const { shim_unreachable("cv::FileNodeIterator"); }
inline void operator>>(const FileNode&, std::vector<int>&) { shim_unreachable("cv::FileNode"); }
class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage() {}
  FileStorage(const std::string&, int) { shim_unreachable("cv::FileStorage"); }
  FileNode root() const { shim_unreachable("cv::FileStorage"); }
};
template <typename T> inline FileStorage& operator<<(FileStorage&, const T&) { shim_unreachable("cv::FileStorage"); }

}  // namespace cv
