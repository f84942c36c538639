// oracle/ref/shim/opencv2/opencv.hpp — TEST INFRASTRUCTURE ONLY.  Minimal stand-in for cv::Mat / cv::FileStorage as far
// as the reference's message headers use them (msg_keyframe.hpp:237-285 save/load of cv::Mat; typedefs_base.hpp:71-99
// yaml helpers, never called here).  Type codes follow OpenCV: depth = type & 7 (0 u8, 1 s8, 2 u16, 3 s16, 4 s32, 5 f32,
// 6 f64), channels = (type >> 3) + 1.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_32FC1 5

namespace cv {
class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() {}
  Mat(int r, int c, int t) { create(r, c, t); }
  void create(int r, int c, int t) { rows = r; cols = c; type_ = t; d_.assign((size_t)r * c * elemSize(), 0); }
  int type() const { return type_; }
  bool isContinuous() const { return true; }
  size_t elemSize() const {
    static const size_t depth_bytes[8] = {1, 1, 2, 2, 4, 4, 8, 2};
    return depth_bytes[type_ & 7] * (size_t)((type_ >> 3) + 1);
  }
  uint8_t* ptr(int i = 0) { return d_.data() + (size_t)i * cols * elemSize(); }
  const uint8_t* ptr(int i = 0) const { return d_.data() + (size_t)i * cols * elemSize(); }
  Mat clone() const { return *this; }

 private:
  int type_ = 0;
  std::vector<uint8_t> d_;
};

class FileNode {
 public:
  operator double() const { return 0.0; }
  operator std::string() const { return std::string(); }
};
class FileStorage {
 public:
  enum { READ = 0 };
  FileStorage(const std::string&, int) {}
  bool isOpened() const { return false; }
  FileNode operator[](const std::string&) const { return FileNode(); }
};
}  // namespace cv
