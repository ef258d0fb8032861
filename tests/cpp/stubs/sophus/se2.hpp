// Minimal stand-in for <sophus/se2.hpp> (Sophus is not in this image): the one property the adaptors rely on --
// SE2d::data() is four doubles {cos, sin, x, y} (Sophus 1.22.10 se2.hpp: so2 unit complex first, then translation).
#pragma once
namespace Sophus {
class SE2d {
 public:
  double* data() { return d_; }
  const double* data() const { return d_; }

 private:
  double d_[4] = {1.0, 0.0, 0.0, 0.0};
};
}  // namespace Sophus
