// TEST INFRASTRUCTURE ONLY -- this file is part of the CPU parity oracle.
// Nothing under beluga_b200/ (the product) includes, links or executes it.
//
// Minimal SO(2)/SE(2) types restating the arithmetic of Sophus 1.22.10
// (third-party dependency of the reference, pinned in /root/reference/MODULE.bazel:27,
// NOT vendored under /root/reference and not installed here). Every beluga hot-path
// header manipulates particle states through these operations:
//   differential_drive_model.hpp:136-139,158-162   (SO2(angle), SO2*SO2, inverse, log, SE2*SE2)
//   likelihood_field_model.hpp:70-74               (SE2*SE2, unit_complex, translation)
//   raycasting.hpp:69,81-85                        (inverse, SO2*point)
//   spatial_hash.hpp:190-193                       (log)
//   estimation.hpp:448-471                         (data() order cos,sin,x,y; normalize)
//
// Published Sophus semantics restated here (sophus/so2.hpp, sophus/se2.hpp @1.22.10):
//   * SO2 stores a unit complex (real=cos, imag=sin); SE2 stores [so2 | translation],
//     so SE2::data() is {cos, sin, x, y}.
//   * SO2(real, imag) and SO2(theta)=SO2::exp(theta) both call normalize(), which divides
//     both components by std::hypot(real, imag).
//   * SO2*SO2 is a complex product followed, when the squared norm is not exactly 1, by the
//     first-order renormalisation scale = 2/(1+|z|^2); the product is then built through the
//     (real, imag) constructor (which normalises again through hypot).
//   * SO2::inverse() = SO2(real, -imag); log() = atan2(imag, real).
//   * SO2*point = (real*px - imag*py, imag*px + real*py).
//   * SE2*SE2 = SE2(so2*other.so2, translation + so2*other.translation).
//   * SE2::inverse() = SE2(invR, invR * (translation * -1)).
// Parity for these is "unpinned" against Sophus itself (it cannot be compiled here); the
// reference's own known-answer tests that flow through them (tests/test_oracle_golden.py)
// pin them to the tolerances the reference states.
#pragma once
#include <cmath>

namespace oracle {

struct Vec2 {
  double x{0.0}, y{0.0};
};

struct SO2 {
  double c{1.0}, s{0.0};  // unit complex: real, imag

  SO2() = default;
  // Sophus SO2(real, imag): stores then normalize() through hypot.
  SO2(double real, double imag) : c(real), s(imag) { normalize(); }
  // Sophus SO2(theta) == SO2::exp(theta) == SO2(cos(theta), sin(theta)).
  explicit SO2(double theta) : SO2(std::cos(theta), std::sin(theta)) {}

  // Raw construction without normalisation (Sophus: writing through data(), cf.
  // beam_model.hpp:120-123 "dirty hack to prevent SO2d from calculating the hypot").
  static SO2 raw(double real, double imag) {
    SO2 r;
    r.c = real;
    r.s = imag;
    return r;
  }

  void normalize() {
    const double length = std::hypot(c, s);
    c /= length;
    s /= length;
  }

  [[nodiscard]] double log() const { return std::atan2(s, c); }
  [[nodiscard]] SO2 inverse() const { return SO2(c, -s); }
  [[nodiscard]] double norm() const { return std::sqrt(c * c + s * s); }  // Eigen Vector2d::norm()

  [[nodiscard]] SO2 operator*(const SO2& o) const {
    const double re = c * o.c - s * o.s;
    const double im = c * o.s + s * o.c;
    const double squared_norm = re * re + im * im;
    if (squared_norm != 1.0) {
      const double scale = 2.0 / (1.0 + squared_norm);
      return SO2(re * scale, im * scale);
    }
    return SO2(re, im);
  }

  [[nodiscard]] Vec2 operator*(const Vec2& p) const { return Vec2{c * p.x - s * p.y, s * p.x + c * p.y}; }
};

struct SE2 {
  SO2 r;
  double x{0.0}, y{0.0};

  SE2() = default;
  SE2(const SO2& so2, double tx, double ty) : r(so2), x(tx), y(ty) {}
  SE2(const SO2& so2, const Vec2& t) : r(so2), x(t.x), y(t.y) {}
  SE2(double theta, double tx, double ty) : r(theta), x(tx), y(ty) {}

  [[nodiscard]] Vec2 translation() const { return Vec2{x, y}; }

  [[nodiscard]] SE2 operator*(const SE2& o) const {
    const Vec2 rt = r * Vec2{o.x, o.y};
    return SE2(r * o.r, x + rt.x, y + rt.y);
  }
  [[nodiscard]] Vec2 operator*(const Vec2& p) const {
    const Vec2 rp = r * p;
    return Vec2{rp.x + x, rp.y + y};
  }
  [[nodiscard]] SE2 inverse() const {
    const SO2 inv = r.inverse();
    const Vec2 t = inv * Vec2{x * -1.0, y * -1.0};
    return SE2(inv, t.x, t.y);
  }
};

}  // namespace oracle
