// ref_shim.cpp — C entry points over the UNMODIFIED reference sources compiled into oracle/_ref/libsdvref.so.  TEST INFRASTRUCTURE ONLY.
//
// The reference (ZikangYuan/SDV-LOAM, /root/reference/src) has no FFI and cannot be built with its own build system here (Eigen3, Boost,
// ROS, OpenCV, PCL absent).  oracle/Makefile compiles its hot-path translation units where they lie — CoarseTracker.cpp, Residuals.cpp,
// HessianBlocks.cpp, EnergyFunctional*.cpp, Accumulated*Hessian.cpp, FullSystemOptimize/Marginalize/OptPoint.cpp, ImmaturePoint.cpp,
// Reprojector.cpp, PixelSelector2.cpp, Undistort.cpp, globalCalib.cpp, settings.cpp — against the stand-in headers in oracle/ref_stub/ (a
// minimal Eigen, Sophus-over-orc_math, empty ROS/OpenCV/Boost shells).  This file is the only non-reference code in that library: it builds the
// reference's own objects (FrameHessian, PointHessian, CoarseTracker, EnergyFunctional ...) from flat arrays and calls the reference's own
// member functions, with the same flat signatures as orc_capi.cpp so one test drives both (tests/test_ref_*.py).  Nothing is copied from the
// reference into the repo; the sources are read from /root/reference at build time (absent on the GPU box: the prebuilt .so travels).
#include <sstream>
#include <fstream>
#include <iostream>
#include <iomanip>
#include <complex>
#include <vector>
#include <deque>
#include <queue>
#include <list>
#include <map>
#include <string>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <memory>
#include <algorithm>
#include "Eigen/Core"
#include "sophus/se3.hpp"
#include "boost/thread.hpp"
#include "ros/ros.h"
#include "sensor_msgs/Image.h"
#include "sensor_msgs/PointCloud2.h"
#include "cv_bridge/cv_bridge.h"
#define private public                                     // the reference keeps calcRes / calcGSSSE / linearizeAll ... private; the shim calls them directly
#define protected public
#include "FullSystem/FullSystem.h"
#include "FullSystem/CoarseTracker.h"
#include "FullSystem/ImmaturePoint.h"
#include "FullSystem/Reprojector.h"
#include "FullSystem/ResidualProjections.h"
#include "OptimizationBackend/EnergyFunctional.h"
#include "OptimizationBackend/EnergyFunctionalStructs.h"
#include "util/globalCalib.h"
#include "util/globalFuncs.h"
#include "util/settings.h"
#include "util/FrameShell.h"
#undef private
#undef protected
#include <vector>
#include <cstring>
#include <cstdio>

using namespace sdv_loam;

namespace {
CalibHessian* g_calib = nullptr;
SE3 se3_from(const double T[7]) { orc::SE3 s; s.q = orc::Quat{T[0], T[1], T[2], T[3]}; s.t = orc::Vec3d{{T[4], T[5], T[6]}}; return SE3(s); }
void se3_to(const SE3& S, double T[7]) { T[0] = S.s.q.w; T[1] = S.s.q.x; T[2] = S.s.q.y; T[3] = S.s.q.z; T[4] = S.s.t[0]; T[5] = S.s.t[1]; T[6] = S.s.t[2]; }

struct RefFrame {                          // a FrameHessian with its FrameShell and the objects the shim hung on it
  FrameHessian* fh; FrameShell* shell;
  std::vector<ImmaturePoint*> ips; std::vector<PointHessian*> phs; std::vector<EFPoint*> efps; std::vector<PointFrameResidual*> res;
};
int g_next_id = 0;
void drop_points(RefFrame* f) {
  for (PointHessian* p : f->phs) { p->efPoint = 0; p->residuals.clear(); delete p; }
  for (EFPoint* e : f->efps) { ::operator delete(e); }
  for (PointFrameResidual* r : f->res) { ::operator delete(r); }
  for (ImmaturePoint* ip : f->ips) delete ip;
  f->phs.clear(); f->efps.clear(); f->res.clear(); f->ips.clear(); f->fh->pointHessians.clear();
}
}

extern "C" {

// ---------------------------------------------------------------------------------------------- globals: calibration + settings
int ref_set_calib(int w, int h, float fx, float fy, float cx, float cy) {       // setGlobalCalib (util/globalCalib.cpp:20-80) + CalibHessian ctor
  Eigen::Matrix3f K; K.setZero(); K(0, 0) = fx; K(1, 1) = fy; K(0, 2) = cx; K(1, 2) = cy; K(2, 2) = 1;
  FILE* so = stdout; (void)so; fflush(stdout);
  setGlobalCalib(w, h, K);
  delete g_calib; g_calib = new CalibHessian();
  return pyrLevelsUsed;
}
void ref_get_global_K(int lvl, float out4[4], float out_i4[4]) { out4[0] = fxG[lvl]; out4[1] = fyG[lvl]; out4[2] = cxG[lvl]; out4[3] = cyG[lvl];
  out_i4[0] = fxiG[lvl]; out_i4[1] = fyiG[lvl]; out_i4[2] = cxiG[lvl]; out_i4[3] = cyiG[lvl]; }
void ref_settings(float huberTH, float coarseCutoffTH, float affA, float affB) {
  setting_huberTH = huberTH; setting_coarseCutoffTH = coarseCutoffTH; setting_affineOptModeA = affA; setting_affineOptModeB = affB;
}

// ---------------------------------------------------------------------------------------------- frames: FrameHessian::makeImages (HessianBlocks.cpp:107-167)
void* ref_frame_create(const float* color, float exposure) {
  RefFrame* f = new RefFrame(); f->shell = new FrameShell(); f->shell->id = g_next_id++; f->shell->incoming_id = f->shell->id;
  f->fh = new FrameHessian(); f->fh->shell = f->shell; f->fh->ab_exposure = exposure;
  std::vector<float> c(color, color + (size_t)wG[0]*hG[0]);
  f->fh->makeImages(c.data(), g_calib);
  f->fh->setEvalPT_scaled(SE3(), AffLight(0, 0));
  return f;
}
void ref_frame_destroy(void* p) { RefFrame* f = (RefFrame*)p; if (!f) return; drop_points(f); f->fh->efFrame = 0; delete f->fh; delete f->shell; delete f; }
const float* ref_frame_dI(void* p, int lvl) { return (const float*)((RefFrame*)p)->fh->dIp[lvl]; }       // Vector3f[w*h] = 3 packed floats
const float* ref_frame_abs(void* p, int lvl) { return ((RefFrame*)p)->fh->absSquaredGrad[lvl]; }

// ---------------------------------------------------------------------------------------------- CoarseTracker (CoarseTracker.cpp)
void* ref_tracker_create() { CoarseTracker* t = new CoarseTracker(wG[0], hG[0]); t->makeK(g_calib); t->debugPrint = false; t->debugPlot = false; return t; }
void ref_tracker_destroy(void* t) { delete (CoarseTracker*)t; }
void ref_tracker_get_K(void* t, int lvl, float out4[4]) { CoarseTracker* T = (CoarseTracker*)t; out4[0] = T->fx[lvl]; out4[1] = T->fy[lvl]; out4[2] = T->cx[lvl]; out4[3] = T->cy[lvl]; }
void ref_tracker_get_Ki(void* t, int lvl, float out9[9]) { CoarseTracker* T = (CoarseTracker*)t; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out9[3*i+j] = T->Ki[lvl](i, j); }

// setCoarseTrackingRef -> makeCoarseDepthL0 (CoarseTracker.cpp:649-660, 258-425).  pts rows {u, v, idepth, HdiF}.  round_half[i] == 0: an active LiDAR point of
// the reference keyframe itself (truncated pixel, :267-277); == 1: a LiDAR point of an OLDER keyframe entering through its IN residual's centerProjectedTo
// (+0.5 rounding, :278-293).  The reference walks the older keyframes first, so all round_half == 1 rows must precede the round_half == 0 rows.
int ref_tracker_set_ref(void* t, void* ref_frame, void* old_frame, const float* pts, const int* round_half, int n, double ref_a, double ref_b) {
  CoarseTracker* T = (CoarseTracker*)t; RefFrame* R = (RefFrame*)ref_frame; RefFrame* O = (RefFrame*)old_frame;
  drop_points(R); if (O) drop_points(O);
  bool seen_direct = false;
  for (int i = 0; i < n; i++) {
    const bool old = round_half[i] != 0;
    if (old && (seen_direct || !O)) return -1;
    if (!old) seen_direct = true;
    RefFrame* F = old ? O : R;
    const float u = pts[4*i], v = pts[4*i+1], id = pts[4*i+2], HdiF = pts[4*i+3];
    ImmaturePoint* ip = new ImmaturePoint(8, 8, F->fh, 0, g_calib); ip->idepth_min = ip->idepth_max = id; F->ips.push_back(ip);   // (u,v) of the PointHessian are set below: the splat truncates them itself
    PointHessian* ph = new PointHessian(ip, g_calib); ph->isFromSensor = true; ph->u = u; ph->v = v; ph->setIdepth(id);
    EFPoint* ef = (EFPoint*)::operator new(sizeof(EFPoint)); memset((void*)ef, 0, sizeof(EFPoint)); ef->HdiF = HdiF; ph->efPoint = ef; F->efps.push_back(ef);
    ph->lastResiduals[0] = std::pair<PointFrameResidual*, ResState>(0, ResState::OOB); ph->lastResiduals[1] = ph->lastResiduals[0];
    if (old) {
      PointFrameResidual* r = (PointFrameResidual*)::operator new(sizeof(PointFrameResidual)); memset((void*)r, 0, sizeof(PointFrameResidual));
      r->centerProjectedTo = Vec3f(u, v, id); r->target = R->fh; F->res.push_back(r);
      ph->lastResiduals[0] = std::pair<PointFrameResidual*, ResState>(r, ResState::IN);
    }
    F->phs.push_back(ph); F->fh->pointHessians.push_back(ph);
  }
  std::vector<FrameHessian*> fhs; if (O) fhs.push_back(O->fh); fhs.push_back(R->fh);
  T->setCoarseTrackingRef(fhs);
  T->lastRef_aff_g2l = AffLight(ref_a, ref_b);
  return 0;
}
int ref_tracker_cloud_n(void* t, int lvl) { return ((CoarseTracker*)t)->pc_n[lvl]; }
void ref_tracker_get_cloud(void* t, int lvl, float* u, float* v, float* id, float* color) { CoarseTracker* T = (CoarseTracker*)t; int n = T->pc_n[lvl];
  for (int i = 0; i < n; i++) { u[i] = T->pc_u[lvl][i]; v[i] = T->pc_v[lvl][i]; id[i] = T->pc_idepth[lvl][i]; color[i] = T->pc_color[lvl][i]; } }
void ref_tracker_calc_res(void* t, void* new_frame, int lvl, const double T7[7], double a, double b, float cutoffTH, double rs[6]) {
  CoarseTracker* T = (CoarseTracker*)t; T->newFrame = ((RefFrame*)new_frame)->fh;
  Vec6 r = T->calcRes(lvl, se3_from(T7), AffLight(a, b), cutoffTH); for (int i = 0; i < 6; i++) rs[i] = r[i];
}
int ref_tracker_warped_n(void* t) { return ((CoarseTracker*)t)->buf_warped_n; }
void ref_tracker_get_warped(void* t, float* out) {      // 8 x n: {idepth,u,v,dx,dy,residual,weight,refColor}
  CoarseTracker* T = (CoarseTracker*)t; int n = T->buf_warped_n;
  const float* bufs[8] = {T->buf_warped_idepth, T->buf_warped_u, T->buf_warped_v, T->buf_warped_dx, T->buf_warped_dy, T->buf_warped_residual, T->buf_warped_weight, T->buf_warped_refColor};
  for (int k = 0; k < 8; k++) for (int i = 0; i < n; i++) out[k*n+i] = bufs[k][i];
}
void ref_tracker_calc_gs(void* t, int lvl, const double T7[7], double a, double b, double H[64], double bb[8]) {
  Mat88 Hm; Vec8 bv; ((CoarseTracker*)t)->calcGSSSE(lvl, Hm, bv, se3_from(T7), AffLight(a, b));
  for (int i = 0; i < 8; i++) { for (int j = 0; j < 8; j++) H[8*i+j] = Hm(i, j); bb[i] = bv[i]; }
}
int ref_tracker_track(void* t, void* new_frame, double T_io[7], double ab_io[2], int coarsest, const double minRes[5], double lastRes[5], double flow[3]) {
  CoarseTracker* T = (CoarseTracker*)t; SE3 s = se3_from(T_io); AffLight aff(ab_io[0], ab_io[1]); Vec5 mr; for (int i = 0; i < 5; i++) mr[i] = minRes[i];
  bool good = T->trackNewestCoarse(((RefFrame*)new_frame)->fh, s, aff, coarsest, mr, 0);
  se3_to(s, T_io); ab_io[0] = aff.a; ab_io[1] = aff.b;
  for (int i = 0; i < 5; i++) lastRes[i] = T->lastResiduals[i];
  for (int i = 0; i < 3; i++) flow[i] = T->lastFlowIndicators[i];
  return good ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------- small kernels of the path, directly
void ref_interp33(const float* dI3, int w, float x, float y, float out3[3]) {        // getInterpolatedElement33 (util/globalFuncs.h:51-65)
  Eigen::Vector3f r = getInterpolatedElement33((const Eigen::Vector3f*)dI3, x, y, w); out3[0] = r[0]; out3[1] = r[1]; out3[2] = r[2];
}
void ref_interp33_bilin(const float* dI3, int w, float x, float y, float out3[3]) {  // getInterpolatedElement33BiLin (util/globalFuncs.h)
  Eigen::Vector3f r = getInterpolatedElement33BiLin((const Eigen::Vector3f*)dI3, x, y, w); out3[0] = r[0]; out3[1] = r[1]; out3[2] = r[2];
}
void ref_aff_from_to(float eF, float eT, double aF, double bF, double aT, double bT, double out[2]) {   // AffLight::fromToVecExposure (util/NumType.h:149-158)
  Vec2 r = AffLight::fromToVecExposure(eF, eT, AffLight(aF, bF), AffLight(aT, bT)); out[0] = r[0]; out[1] = r[1];
}

}  // extern "C"
