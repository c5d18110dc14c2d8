// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned on oracle/_ref at sequence level (tests/test_sequence_parity.py).
// CPU restatement of the semi-direct pose refinement that follows the photometric tracker, file:line relative
// to /root/reference/src:
//   CoarseTracker::calculateRes          FullSystem/CoarseTracker.cpp:840-871
//   CoarseTracker::calculateWeight       FullSystem/CoarseTracker.cpp:873-887   (Tukey, b = 4.6851)
//   CoarseTracker::calcHandb             FullSystem/CoarseTracker.cpp:889-947
//   CoarseTracker::structPoseEstimation  FullSystem/CoarseTracker.cpp:949-1007
//   point2world / world2frame / pixel2unit   FullSystem/ResidualProjections.h:61-102
// Load-bearing quirks kept on purpose: d_xi_x[4] = 1 + X*d_xi_x[2] and d_xi_y[3] = -(1 + Y*d_xi_y[2]) (:916,:922) evaluate to 1 - (X/Z)^2 and -(1 - (Y/Z)^2), not the
// derivatives 1 + (X/Z)^2 / -(1 + (Y/Z)^2) (tests/test_oracle_refine.py::test_normal_equations_by_finite_differences); after an accepted step H,b are re-linearised at the OLD pose (:989 before :990), the damping
// factor is applied to H in place every iteration (so it compounds across rejected steps, :966), the function has no return value.
#include "orc_tracker.hpp"
#include <vector>

namespace orc {

struct OverlapPoint { float u, v, idepth; int host; float obs[2]; };

static inline bool project_overlap(const OverlapPoint& p, const Mat33f* hostR, const Vec3f* hostT, const Mat33f& R, const Vec3f& t,
                                   float fx, float fy, float cx, float cy, float fxi, float fyi, float wM3G, float hM3G, Vec3f& ptFrame, float& Ku, float& Kv) {
  // point2world (dx=dy=0)
  Vec3f KliP{{(p.u+0-cx)*fxi, (p.v+0-cy)*fyi, 1}};
  Vec3f ptRef{{KliP[0]/p.idepth, KliP[1]/p.idepth, KliP[2]/p.idepth}};
  Vec3f ptWorld = matvec(hostR[p.host], ptRef[0], ptRef[1], ptRef[2]); for (int c=0;c<3;c++) ptWorld.v[c] = ptWorld.v[c] + hostT[p.host].v[c];
  // world2frame
  ptFrame = matvec(R, ptWorld[0], ptWorld[1], ptWorld[2]); for (int c=0;c<3;c++) ptFrame.v[c] = ptFrame.v[c] + t.v[c];
  float u0 = ptFrame[0]/ptFrame[2], u1 = ptFrame[1]/ptFrame[2];
  Ku = u0*fx + cx; Kv = u1*fy + cy;
  return Ku>1.1f && Kv>1.1f && Ku<wM3G && Kv<hM3G;
}

struct PoseRefiner {
  float fx, fy, cx, cy, fxi, fyi, cxi, cyi; int w, h;
  std::vector<Mat33f> hostR; std::vector<Vec3f> hostT;
  int iterations = 0, accepts = 0; float lastRes = 0;

  float calculateRes(const SE3& worldToCur, const std::vector<OverlapPoint>& pts, int& num) {
    float energy = 0.0; num = 0;
    Mat33f R = castf(worldToCur.rotationMatrix()); Vec3f t = castf(worldToCur.t);
    for (const auto& p : pts) { Vec3f pf; float Ku, Kv;
      if (project_overlap(p, hostR.data(), hostT.data(), R, t, fx, fy, cx, cy, fxi, fyi, (float)(w-3), (float)(h-3), pf, Ku, Kv)) {
        float r0 = Ku - p.obs[0], r1 = Kv - p.obs[1];
        energy = energy + r0*r0 + r1*r1; num++; } }
    return energy;
  }
  static float calculateWeight(float x) {
    const float b = 4.6851f; float b2 = b*b, x2 = x*x;
    if (x2 <= b2) { float tmp = 1.0f - x2/b2; return tmp*tmp; }
    return 0.0f;
  }
  void calcHandb(double H[36], double b[6], const SE3& worldToCur, const std::vector<OverlapPoint>& pts) {
    Mat33f R = castf(worldToCur.rotationMatrix()); Vec3f t = castf(worldToCur.t);
    for (const auto& p : pts) { Vec3f pf; float Ku, Kv;
      if (!project_overlap(p, hostR.data(), hostT.data(), R, t, fx, fy, cx, cy, fxi, fyi, (float)(w-3), (float)(h-3), pf, Ku, Kv)) continue;
      float dx[6], dy[6];
      dx[0] = 1.0 / pf[2]; dx[1] = 0.0; dx[2] = - pf[0]/ (pf[2] * pf[2]); dx[3] = dx[2] * pf[1]; dx[4] = 1 + pf[0] * dx[2]; dx[5] = - pf[1] / pf[2];
      dy[0] = 0.0; dy[1] = 1.0 / pf[2]; dy[2] = - pf[1]/ (pf[2] * pf[2]); dy[3] = - (1 + pf[1] * dy[2]); dy[4] = - dx[3]; dy[5] = pf[0] / pf[2];
      float up = (Ku - cx)*fxi, vp = (Kv - cy)*fyi;                        // pixel2unit
      float uo = (p.obs[0] - cx)*fxi, vo = (p.obs[1] - cy)*fyi;
      float r0 = up - uo, r1 = vp - vo;
      double weight = calculateWeight(std::sqrt(r0*r0 + r1*r1));
      double Jx[6], Jy[6]; for (int i=0;i<6;i++) { Jx[i]=(double)dx[i]; Jy[i]=(double)dy[i]; }
      for (int i=0;i<6;i++) { for (int j=0;j<6;j++) H[i*6+j] += (Jx[i]*Jx[j] + Jy[i]*Jy[j])*weight;
        b[i] += (Jx[i]*(double)r0 + Jy[i]*(double)r1)*weight; }
    }
  }
  void structPoseEstimation(SE3& curToWorld, const std::vector<OverlapPoint>& pts) {
    iterations = accepts = 0;
    SE3 worldToCur_current = curToWorld.inverse();
    float lambda = 0.01, lambdaExtrapolationLimit = 0.001;
    double H[36] = {0}, b[6] = {0};
    int num; float resNew = 0.0;
    float resOld = calculateRes(worldToCur_current, pts, num); resOld = resOld / num;
    calcHandb(H, b, worldToCur_current, pts);
    for (int iteration=0; iteration<10; iteration++) {
      iterations++;
      for (int i=0;i<6;i++) H[i*6+i] *= (1 + lambda);
      double nb[6], inc[6]; for (int i=0;i<6;i++) nb[i] = -b[i];
      ldlt_solve<8>(6, H, nb, inc);
      float extrapFac = 1;
      if (lambda < lambdaExtrapolationLimit) extrapFac = std::sqrt(std::sqrt(lambdaExtrapolationLimit / lambda));
      for (int i=0;i<6;i++) inc[i] *= extrapFac;
      SE3 worldToCur_new = SE3::exp(inc) * worldToCur_current;
      resNew = calculateRes(worldToCur_new, pts, num);
      if (num == 0) resNew = 1000000.0; else resNew = resNew / num;
      bool accept = (resNew < resOld);
      if (accept) {
        accepts++;
        for (int i=0;i<36;i++) H[i]=0; for (int i=0;i<6;i++) b[i]=0;
        resOld = resNew; resNew = 0;
        calcHandb(H, b, worldToCur_current, pts);                       // (sic) linearised at the pose BEFORE the accepted step
        worldToCur_current = worldToCur_new; curToWorld = worldToCur_new.inverse();
        lambda *= 0.5;
      } else { lambda *= 4; if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit; }
      double nrm = 0; for (int i=0;i<6;i++) nrm += inc[i]*inc[i]; nrm = std::sqrt(nrm);
      if (!(nrm > 1e-5)) break;
    }
    lastRes = resOld;
  }
};

} // namespace orc

using namespace orc;
extern "C" {
// K4 = {fx,fy,cx,cy} of level 0 ; hostT7 (nH x 7) camToWorld of the host keyframes ; pts (n x 6 floats {u,v,idepth,host,obs_x,obs_y})
// curToWorld7 in/out.  stats: {iterations, accepts}.  Returns the final mean squared reprojection error (resOld).
float orc_struct_pose(int w, int h, const float K4[4], int nH, const double* hostT7, int n, const float* pts6, double curToWorld7[7], int* stats2) {
  PoseRefiner P; P.w = w; P.h = h; P.fx = K4[0]; P.fy = K4[1]; P.cx = K4[2]; P.cy = K4[3];
  Mat33f Km; std::memset(&Km,0,sizeof(Km)); Km.m[0][0]=K4[0]; Km.m[1][1]=K4[1]; Km.m[0][2]=K4[2]; Km.m[1][2]=K4[3]; Km.m[2][2]=1;
  Mat33f Ki = inverse3<float,Mat33f>(Km); P.fxi = Ki.m[0][0]; P.fyi = Ki.m[1][1]; P.cxi = Ki.m[0][2]; P.cyi = Ki.m[1][2];
  for (int k=0;k<nH;k++) { SE3 s; s.q = Quat{hostT7[7*k],hostT7[7*k+1],hostT7[7*k+2],hostT7[7*k+3]}; s.t = Vec3d{{hostT7[7*k+4],hostT7[7*k+5],hostT7[7*k+6]}};
    P.hostR.push_back(castf(s.rotationMatrix())); P.hostT.push_back(castf(s.t)); }
  std::vector<OverlapPoint> v(n);
  for (int i=0;i<n;i++) { v[i].u=pts6[6*i]; v[i].v=pts6[6*i+1]; v[i].idepth=pts6[6*i+2]; v[i].host=(int)pts6[6*i+3]; v[i].obs[0]=pts6[6*i+4]; v[i].obs[1]=pts6[6*i+5]; }
  SE3 c2w; c2w.q = Quat{curToWorld7[0],curToWorld7[1],curToWorld7[2],curToWorld7[3]}; c2w.t = Vec3d{{curToWorld7[4],curToWorld7[5],curToWorld7[6]}};
  P.structPoseEstimation(c2w, v);
  curToWorld7[0]=c2w.q.w; curToWorld7[1]=c2w.q.x; curToWorld7[2]=c2w.q.y; curToWorld7[3]=c2w.q.z; curToWorld7[4]=c2w.t.v[0]; curToWorld7[5]=c2w.t.v[1]; curToWorld7[6]=c2w.t.v[2];
  if (stats2) { stats2[0]=P.iterations; stats2[1]=P.accepts; }
  return P.lastRes;
}
// normal equations + mean squared pixel error at one pose (calcHandb :889-947, calculateRes :840-871) — exposed for the finite-difference test
float orc_struct_pose_hb(int w, int h, const float K4[4], int nH, const double* hostT7, int n, const float* pts6, const double curToWorld7[7], double* H36, double* b6, int* num_out) {
  PoseRefiner P; P.w = w; P.h = h; P.fx = K4[0]; P.fy = K4[1]; P.cx = K4[2]; P.cy = K4[3];
  Mat33f Km; std::memset(&Km,0,sizeof(Km)); Km.m[0][0]=K4[0]; Km.m[1][1]=K4[1]; Km.m[0][2]=K4[2]; Km.m[1][2]=K4[3]; Km.m[2][2]=1;
  Mat33f Ki = inverse3<float,Mat33f>(Km); P.fxi = Ki.m[0][0]; P.fyi = Ki.m[1][1]; P.cxi = Ki.m[0][2]; P.cyi = Ki.m[1][2];
  for (int k=0;k<nH;k++) { SE3 s; s.q = Quat{hostT7[7*k],hostT7[7*k+1],hostT7[7*k+2],hostT7[7*k+3]}; s.t = Vec3d{{hostT7[7*k+4],hostT7[7*k+5],hostT7[7*k+6]}};
    P.hostR.push_back(castf(s.rotationMatrix())); P.hostT.push_back(castf(s.t)); }
  std::vector<OverlapPoint> v(n);
  for (int i=0;i<n;i++) { v[i].u=pts6[6*i]; v[i].v=pts6[6*i+1]; v[i].idepth=pts6[6*i+2]; v[i].host=(int)pts6[6*i+3]; v[i].obs[0]=pts6[6*i+4]; v[i].obs[1]=pts6[6*i+5]; }
  SE3 c2w; c2w.q = Quat{curToWorld7[0],curToWorld7[1],curToWorld7[2],curToWorld7[3]}; c2w.t = Vec3d{{curToWorld7[4],curToWorld7[5],curToWorld7[6]}};
  SE3 w2c = c2w.inverse();
  for (int i=0;i<36;i++) H36[i]=0; for (int i=0;i<6;i++) b6[i]=0;
  P.calcHandb(H36, b6, w2c, v);
  int num; float e = P.calculateRes(w2c, v, num); if (num_out) *num_out = num;
  return num ? e/num : 0.f;
}
}
