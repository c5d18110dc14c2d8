#pragma once   // stand-in for OpenCV core (absent): cv::Mat appears in FullSystem.h member/signature types only
typedef unsigned char uchar;
#define CV_8UC1 0
#include <vector>
#include <memory>
namespace cv {
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
class Mat { public: int rows = 0, cols = 0; unsigned char* data = nullptr; std::shared_ptr<std::vector<unsigned char> > buf; Mat() {}
  static Mat zeros(int r, int c, int) { Mat m; m.rows = r; m.cols = c; m.buf = std::make_shared<std::vector<unsigned char> >((size_t)r*c, 0); m.data = m.buf->data(); return m; }
  Mat clone() const { Mat m = *this; if (buf) { m.buf = std::make_shared<std::vector<unsigned char> >(*buf); m.data = m.buf->data(); } return m; }
  template <class T> T& at(int r, int c) { return *(T*)(data + ((size_t)r*cols + c)*sizeof(T)); } bool empty() const { return data == nullptr; } void release() { buf.reset(); data = nullptr; rows = cols = 0; } };
}
