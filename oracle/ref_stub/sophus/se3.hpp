// Stand-in for thirdparty/Sophus (header-only on top of the real Eigen, which is absent) — TEST INFRASTRUCTURE (oracle/_ref build only).
// SE3d over the oracle's restatement of Sophus v0.9a (oracle/orc_math.hpp, pinned on Sophus' own sample transforms in tests/test_oracle_math.py).
#pragma once
#include "Eigen/Core"
#include "../../orc_math.hpp"
namespace Sophus {
typedef Eigen::Quaterniond Quaterniond;
class SO3d { public: orc::Quat q{1,0,0,0}; Eigen::Quaterniond unit_quaternion() const { return Eigen::Quaterniond(q.w, q.x, q.y, q.z); } Eigen::Matrix<double,3,3> matrix() const { orc::Mat33d R = orc::qmat(q); Eigen::Matrix<double,3,3> M; for (int i=0;i<3;i++) for (int j=0;j<3;j++) M(i,j) = R.m[i][j]; return M; } };
class SE3d {
 public:
  orc::Quat q{1,0,0,0}; Eigen::Matrix<double,3,1> t_;      // translation kept as an Eigen vector: translation() hands out a mutable reference like Sophus does
  typedef Eigen::Matrix<double,6,1> Tangent;
  orc::SE3 s() const { orc::SE3 o; o.q = q; o.t = orc::Vec3d{{t_[0], t_[1], t_[2]}}; return o; }
  SE3d() { t_.setZero(); }
  SE3d(const orc::SE3& o) : q(o.q) { t_[0] = o.t[0]; t_[1] = o.t[1]; t_[2] = o.t[2]; }
  SE3d(const Eigen::Matrix<double,3,3>& R, const Eigen::Matrix<double,3,1>& t) { orc::Mat33d M; for (int i=0;i<3;i++) for (int j=0;j<3;j++) M.m[i][j] = R(i,j);
    *this = SE3d(orc::SE3::fromQuatT(orc::qfrommat(M), orc::Vec3d{{t[0], t[1], t[2]}})); }
  SE3d(const Eigen::Quaterniond& qq, const Eigen::Matrix<double,3,1>& t) { *this = SE3d(orc::SE3::fromQuatT(orc::Quat{qq.w(), qq.x(), qq.y(), qq.z()}, orc::Vec3d{{t[0], t[1], t[2]}})); }   // so3.hpp:630-633 normalises
  static SE3d exp(const Eigen::Matrix<double,6,1>& a) { double v[6]; for (int i=0;i<6;i++) v[i] = a[i]; return SE3d(orc::SE3::exp(v)); }
  template <int BR, int BC> static SE3d exp(const Eigen::BlockRef<double,BR,BC>& a) { return exp(Eigen::Matrix<double,6,1>(a)); }
  Eigen::Matrix<double,6,1> log() const { double v[6]; s().log(v); Eigen::Matrix<double,6,1> o; for (int i=0;i<6;i++) o[i] = v[i]; return o; }
  SE3d operator*(const SE3d& o) const { return SE3d(s()*o.s()); }
  Eigen::Matrix<double,3,1> operator*(const Eigen::Matrix<double,3,1>& p) const { orc::Vec3d r = orc::qrot(q, orc::Vec3d{{p[0],p[1],p[2]}}); return Eigen::Matrix<double,3,1>(r[0]+t_[0], r[1]+t_[1], r[2]+t_[2]); }
  SE3d inverse() const { return SE3d(s().inverse()); }
  Eigen::Matrix<double,3,3> rotationMatrix() const { orc::Mat33d R = orc::qmat(q); Eigen::Matrix<double,3,3> M; for (int i=0;i<3;i++) for (int j=0;j<3;j++) M(i,j) = R.m[i][j]; return M; }
  Eigen::Matrix<double,3,1>& translation() { return t_; }
  const Eigen::Matrix<double,3,1>& translation() const { return t_; }
  Eigen::Matrix<double,6,6> Adj() const { double A[6][6]; s().Adj(A); Eigen::Matrix<double,6,6> M; for (int i=0;i<6;i++) for (int j=0;j<6;j++) M(i,j) = A[i][j]; return M; }
  Eigen::Matrix<double,4,4> matrix() const { Eigen::Matrix<double,4,4> M; M.setIdentity(); auto R = rotationMatrix(); for (int i=0;i<3;i++) { for (int j=0;j<3;j++) M(i,j) = R(i,j); M(i,3) = t_[i]; } return M; }
  Eigen::Matrix<double,3,4> matrix3x4() const { Eigen::Matrix<double,3,4> M; auto R = rotationMatrix(); for (int i=0;i<3;i++) { for (int j=0;j<3;j++) M(i,j) = R(i,j); M(i,3) = t_[i]; } return M; }
  SO3d so3() const { SO3d r; r.q = q; return r; }
};
typedef SE3d SE3;
}
