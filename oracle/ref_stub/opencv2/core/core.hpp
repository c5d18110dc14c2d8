#pragma once   // stand-in for OpenCV core (absent): cv::Mat appears in FullSystem.h member/signature types only
typedef unsigned char uchar;
#define CV_8UC1 0
#define CV_8S 1
#define CV_32S 4
#define CV_32F 5
#include <vector>
#include <memory>
namespace cv {
struct Scalar { double v = 0; Scalar() {} Scalar(double a) : v(a) {} static Scalar all(double a) { return Scalar(a); } };
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
class Mat { public: int rows = 0, cols = 0, type_ = 0; unsigned char* data = nullptr; std::shared_ptr<std::vector<unsigned char> > buf; Mat() {}
  static Mat zeros(int r, int c, int) { Mat m; m.rows = r; m.cols = c; m.buf = std::make_shared<std::vector<unsigned char> >((size_t)r*c, 0); m.data = m.buf->data(); return m; }
  static size_t esz(int t) { return (t == CV_32S || t == CV_32F) ? 4 : 1; }
  Mat(int r, int c, int t, const Scalar& sc = Scalar()) { rows = r; cols = c; type_ = t; buf = std::make_shared<std::vector<unsigned char> >((size_t)r*c*esz(t), 0); data = buf->data(); fill(sc); }   // main.cpp's range / label / ground images
  void fill(const Scalar& sc) { size_t n = (size_t)rows*cols; if (type_ == CV_32F) { float* f = (float*)data; for (size_t i = 0; i < n; i++) f[i] = (float)sc.v; } else if (type_ == CV_32S) { int* f = (int*)data; for (size_t i = 0; i < n; i++) f[i] = (int)sc.v; }
    else { signed char* f = (signed char*)data; for (size_t i = 0; i < n; i++) f[i] = (signed char)sc.v; } }
  Mat clone() const { Mat m = *this; if (buf) { m.buf = std::make_shared<std::vector<unsigned char> >(*buf); m.data = m.buf->data(); } return m; }
  template <class T> T& at(int r, int c) { return *(T*)(data + ((size_t)r*cols + c)*sizeof(T)); } bool empty() const { return data == nullptr; } void release() { buf.reset(); data = nullptr; rows = cols = 0; } };
}
