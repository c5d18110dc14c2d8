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
#include "pcl/point_cloud.h"
#include "pcl_conversions/pcl_conversions.h"
#include "pcl/filters/filter.h"
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
#include "util/Undistort.h"
#undef private
#undef protected
#include <vector>
#include <cstring>
#include <cstdio>

using namespace sdv_loam;

namespace {
CalibHessian* g_calib = nullptr;
SE3 se3_from(const double T[7]) { orc::SE3 s; s.q = orc::Quat{T[0], T[1], T[2], T[3]}; s.t = orc::Vec3d{{T[4], T[5], T[6]}}; return SE3(s); }
void se3_to(const SE3& S, double T[7]) { T[0] = S.q.w; T[1] = S.q.x; T[2] = S.q.y; T[3] = S.q.z; T[4] = S.t_[0]; T[5] = S.t_[1]; T[6] = S.t_[2]; }

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
  // makeImages never writes the first and the last row of absSquaredGrad[lvl] / of the gradient channels (HessianBlocks.cpp:147: idx in [wl, wl*(hl-1))) and the
  // arrays come from new[]: PixelSelector reads the last level-2 row for pixels of row h-4 (PixelSelector2.cpp:306).  Define those rows as zero (what a fresh
  // heap page holds) so the pins are deterministic; nothing the reference computes is changed.
  for (int l = 0; l < pyrLevelsUsed; l++) { int wl = wG[l], hl = hG[l]; float* a = f->fh->absSquaredGrad[l]; Eigen::Vector3f* d = f->fh->dIp[l];
    for (int x = 0; x < wl; x++) { a[x] = 0; a[(size_t)wl*(hl-1)+x] = 0; d[x][1] = d[x][2] = 0; d[(size_t)wl*(hl-1)+x][1] = d[(size_t)wl*(hl-1)+x][2] = 0; } }
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

// ---------------------------------------------------------------------------------------------- back-end: a FullSystem whose window is filled from flat arrays
// The reference's own FullSystem (FullSystem.cpp ctor: trackers, selector, EnergyFunctional, mapping thread idle) with frameHessians / PointHessians /
// PointFrameResiduals inserted through EnergyFunctional::insertFrame / insertPoint / insertResidual, i.e. what makeKeyFrame + activatePointsMT leave behind.
// Same flat window layout as orc_ba_* (synth.make_ba_window): points grouped by host in frame order, residuals grouped per point.
struct RefBA { FullSystem* fs; std::vector<RefFrame*> frames; std::vector<PointHessian*> pts; std::vector<ImmaturePoint*> ips; std::vector<PointFrameResidual*> res; std::vector<PointHessian*> res_pt; std::vector<char> dead; };
// linearizeAll(fix) / flagPointsForRemoval DELETE residuals that went out of bounds (FullSystemOptimize.cpp:129-157): find out which of ours are gone (pointer compare only)
static void refresh_dead(RefBA* b) {
  b->dead.assign(b->res.size(), 1);
  for (size_t i = 0; i < b->res.size(); i++) for (PointFrameResidual* r : b->res_pt[i]->residuals) if (r == b->res[i]) { b->dead[i] = 0; break; }
}

void* ref_ba_create() {
  setting_logStuff = false; multiThreading = false; setting_debugout_runquiet = true;
  RefBA* b = new RefBA(); b->fs = new FullSystem(); b->fs->linearizeOperation = true; return b;
}
void ref_ba_destroy(void* p) {       // the window's frames belong to the caller (RefFrame); points/residuals are ours: detach everything before ~FullSystem walks its lists
  RefBA* b = (RefBA*)p; if (!b) return; FullSystem* fs = b->fs;
  for (RefFrame* f : b->frames) { f->fh->pointHessians.clear(); f->fh->pointHessiansMarginalized.clear(); f->fh->pointHessiansOut.clear(); f->fh->efFrame = 0; }
  fs->frameHessians.clear(); fs->activeResiduals.clear();
  fs->blockUntilMappingIsFinished();
  // leak the (small) reference-side graph rather than run destructors over a graph we assembled by hand
  delete b;
}
void ref_ba_set_calib(void* p, const double vs[4]) { RefBA* b = (RefBA*)p; VecC v; for (int i = 0; i < 4; i++) v[i] = vs[i]; b->fs->Hcalib.setValueScaled(v); b->fs->Hcalib.value_zero = b->fs->Hcalib.value; b->fs->Hcalib.value_minus_value_zero.setZero(); }
void ref_ba_add_frame(void* p, void* frame, const double T_eval[7], const double state[10], const double state_zero[10], float ab_exposure, int frameID, float frameEnergyTH) {
  RefBA* b = (RefBA*)p; RefFrame* F = (RefFrame*)frame; FrameHessian* fh = F->fh; drop_points(F);
  Vec10 st, sz; for (int i = 0; i < 10; i++) { st[i] = state[i]; sz[i] = state_zero[i]; }
  fh->ab_exposure = ab_exposure; fh->frameID = frameID; fh->frameEnergyTH = frameEnergyTH; fh->flaggedForMarginalization = false;
  fh->worldToCam_evalPT = se3_from(T_eval); fh->setState(st); fh->setStateZero(sz);
  fh->idx = (int)b->fs->frameHessians.size(); b->fs->frameHessians.push_back(fh); b->frames.push_back(F);
  b->fs->ef->insertFrame(fh, &b->fs->Hcalib);
  fh->step.setZero(); fh->state_backup = fh->state; fh->step_backup.setZero();
}
void ref_ba_set_points(void* p, int nP, const float* uv, const float* idepth, const float* idepth_zero, const float* color, const float* weights,
                       const int* host, const int* hasDepthPrior, const int* isFromSensor, const int* /*res_begin*/) {
  RefBA* b = (RefBA*)p; FullSystem* fs = b->fs;
  for (int i = 0; i < nP; i++) {
    FrameHessian* fh = fs->frameHessians[host[i]];
    ImmaturePoint* ip = new ImmaturePoint((int)uv[2*i], (int)uv[2*i+1], fh, 0, &fs->Hcalib); ip->idepth_min = ip->idepth_max = idepth[i]; b->ips.push_back(ip);
    PointHessian* ph = new PointHessian(ip, &fs->Hcalib);
    ph->u = uv[2*i]; ph->v = uv[2*i+1]; for (int k = 0; k < 8; k++) { ph->color[k] = color[8*i+k]; ph->weights[k] = weights[8*i+k]; }
    ph->setIdepth(idepth[i]); ph->setIdepthZero(idepth_zero[i]); ph->hasDepthPrior = hasDepthPrior[i] != 0; ph->isFromSensor = isFromSensor[i] != 0;
    ph->setPointStatus(PointHessian::ACTIVE); ph->step = 0; ph->step_backup = 0; ph->idepth_backup = ph->idepth;
    ph->lastResiduals[0] = std::pair<PointFrameResidual*, ResState>(0, ResState::OOB); ph->lastResiduals[1] = ph->lastResiduals[0];
    fh->pointHessians.push_back(ph); fs->ef->insertPoint(ph); b->pts.push_back(ph);
  }
}
void ref_ba_set_residuals(void* p, int nR, const int* point, const int* host, const int* target, const int* hasMatcher, const float* matcher, const int* isNew) {
  RefBA* b = (RefBA*)p; FullSystem* fs = b->fs;
  for (int i = 0; i < nR; i++) {
    PointHessian* ph = b->pts[point[i]];
    PointFrameResidual* r = new PointFrameResidual(ph, fs->frameHessians[host[i]], fs->frameHessians[target[i]]);
    r->hasMatcher = hasMatcher[i] != 0; r->matcher = Eigen::Vector2d(matcher[2*i], matcher[2*i+1]); r->isNew = isNew[i] != 0; r->setState(ResState::IN);
    ph->residuals.push_back(r); fs->ef->insertResidual(r); fs->activeResiduals.push_back(r); b->res.push_back(r); b->res_pt.push_back(ph); b->dead.push_back(0);
    ph->lastResiduals[1] = ph->lastResiduals[0]; ph->lastResiduals[0] = std::pair<PointFrameResidual*, ResState>(r, ResState::IN);
  }
}
void ref_ba_set_prior(void* p, const double* HM, const double* bM) { RefBA* b = (RefBA*)p; EnergyFunctional* ef = b->fs->ef; int n = (int)ef->HM.rows();
  for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) ef->HM(i, j) = HM[(size_t)i*n+j]; ef->bM[i] = bM[i]; } }
void ref_ba_init(void* p) { RefBA* b = (RefBA*)p; b->fs->ef->makeIDX(); b->fs->ef->setAdjointsF(&b->fs->Hcalib); b->fs->setPrecalcValues(); }
void ref_ba_reset_oob(void* p) { RefBA* b = (RefBA*)p; for (size_t i = 0; i < b->res.size(); i++) if (!b->dead[i]) b->res[i]->resetOOB(); }
double ref_ba_linearize_all(void* p, int fix) { RefBA* b = (RefBA*)p; Vec3 v = b->fs->linearizeAll(fix != 0); refresh_dead(b); return v[0]; }
void ref_ba_apply_res(void* p) { RefBA* b = (RefBA*)p; for (size_t i = 0; i < b->res.size(); i++) if (!b->dead[i]) b->res[i]->applyRes(true); }
double ref_ba_energy_L(void* p) { return ((RefBA*)p)->fs->calcLEnergy(); }
double ref_ba_energy_M(void* p) { return ((RefBA*)p)->fs->calcMEnergy(); }
static void packJ(const RawResidualJacobian* J, float* o) { o[0] = J->resF[0]; o[1] = J->resF[1]; for (int i = 0; i < 6; i++) { o[2+i] = J->Jpdxi[0][i]; o[8+i] = J->Jpdxi[1][i]; }
  for (int i = 0; i < 4; i++) { o[14+i] = J->Jpdc[0][i]; o[18+i] = J->Jpdc[1][i]; } o[22] = J->Jpdd[0]; o[23] = J->Jpdd[1]; }
void ref_ba_get_residuals(void* p, int* state_state, int* state_NewState, double* energies3, int* isActive, float* J24, float* efJ24, float* JpJdF8, float* center3, int* isLinearized) {
  RefBA* b = (RefBA*)p; int n = (int)b->res.size();
  for (int i = 0; i < n; i++) { if (b->dead[i]) { state_state[i] = state_NewState[i] = -1; isActive[i] = 0; isLinearized[i] = 0; continue; }
    const PointFrameResidual* r = b->res[i]; state_state[i] = (int)r->state_state; state_NewState[i] = (int)r->state_NewState;
    energies3[3*i] = r->state_energy; energies3[3*i+1] = r->state_NewEnergy; energies3[3*i+2] = r->state_NewEnergyWithOutlier; isActive[i] = r->efResidual->isActive() ? 1 : 0;
    packJ(r->J, J24+24*i); packJ(r->efResidual->J, efJ24+24*i); for (int k = 0; k < 8; k++) JpJdF8[8*i+k] = r->efResidual->JpJdF[k]; for (int k = 0; k < 3; k++) center3[3*i+k] = r->centerProjectedTo[k];
    isLinearized[i] = r->efResidual->isLinearized ? 1 : 0; }
}
static void copy_mat(const MatXX& M, double* o) { int n = (int)M.rows(), m = (int)M.cols(); for (int i = 0; i < n; i++) for (int j = 0; j < m; j++) o[(size_t)i*m+j] = M(i, j); }
void ref_ba_accumulate(void* p, double* HA, double* bA, double* Hsc, double* bsc) {       // EnergyFunctional::accumulateAF_MT / accumulateSCF_MT (single-threaded path)
  RefBA* b = (RefBA*)p; EnergyFunctional* ef = b->fs->ef; MatXX H1, H2; VecX b1, b2;
  MatXX HL; VecX bL;
  ef->accumulateAF_MT(H1, b1, false); ef->accumulateLF_MT(HL, bL, false); ef->accumulateSCF_MT(H2, b2, false);     // same order as solveSystemF (:662-666): SC reads the *_accLF sums of the L pass
  copy_mat(H1, HA); copy_mat(H2, Hsc); for (int i = 0; i < (int)b1.size(); i++) { bA[i] = b1[i]; bsc[i] = b2[i]; }
}
void ref_ba_solve(void* p, int iteration, double lambda, double* x, double* HS, double* bS) {
  RefBA* b = (RefBA*)p; b->fs->solveSystem(iteration, lambda); EnergyFunctional* ef = b->fs->ef;
  for (int i = 0; i < (int)ef->lastX.size(); i++) x[i] = ef->lastX[i];
  if (HS) copy_mat(ef->lastHS, HS); if (bS) for (int i = 0; i < (int)ef->lastbS.size(); i++) bS[i] = ef->lastbS[i];
}
void ref_ba_backup(void* p) { ((RefBA*)p)->fs->backupState(false); }
int  ref_ba_do_step(void* p, float f) { return ((RefBA*)p)->fs->doStepFromBackup(f, f, f, f, f) ? 1 : 0; }
void ref_ba_load_backup(void* p) { ((RefBA*)p)->fs->loadSateBackup(); }
float ref_ba_optimize(void* p, int its) { RefBA* b = (RefBA*)p; float r = b->fs->optimize(its); refresh_dead(b); return r; }
void ref_ba_get_points(void* p, float* idepth, float* step, float* HdiF, float* bdSumF, float* maxRelBaseline, int* numGood, float* idepth_hessian) {
  RefBA* b = (RefBA*)p; for (size_t i = 0; i < b->pts.size(); i++) { const PointHessian* q = b->pts[i]; idepth[i] = q->idepth; step[i] = q->step; HdiF[i] = q->efPoint->HdiF; bdSumF[i] = q->efPoint->bdSumF;
    maxRelBaseline[i] = q->maxRelBaseline; numGood[i] = q->numGoodResiduals; idepth_hessian[i] = q->idepth_hessian; }
}
void ref_ba_get_frames(void* p, double* T_eval7, double* state10, double* step10, float* frameEnergyTH, double* PRE_w2c7) {
  RefBA* b = (RefBA*)p; FullSystem* fs = b->fs;
  for (size_t i = 0; i < fs->frameHessians.size(); i++) { const FrameHessian* f = fs->frameHessians[i]; se3_to(f->worldToCam_evalPT, T_eval7+7*i); se3_to(f->PRE_worldToCam, PRE_w2c7+7*i);
    for (int k = 0; k < 10; k++) { state10[10*i+k] = f->state[k]; step10[10*i+k] = f->step[k]; } frameEnergyTH[i] = f->frameEnergyTH; }
}
void ref_ba_get_calib(void* p, double value[4], double step[4]) { RefBA* b = (RefBA*)p; for (int i = 0; i < 4; i++) { value[i] = b->fs->Hcalib.value[i]; step[i] = b->fs->Hcalib.step[i]; } }
void ref_ba_get_precalc(void* p, int host, int target, float* out, double* adH36, double* adT36, float* adHTdelta6) {
  RefBA* b = (RefBA*)p; FullSystem* fs = b->fs; int n = (int)fs->frameHessians.size(); const FrameFramePrecalc& c = fs->frameHessians[host]->targetPrecalc[target];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { out[i*3+j] = c.PRE_KRKiTll(i, j); out[12+i*3+j] = c.PRE_RTll_0(i, j); }
  for (int i = 0; i < 3; i++) { out[9+i] = c.PRE_KtTll[i]; out[21+i] = c.PRE_tTll_0[i]; } out[24] = c.PRE_aff_mode[0]; out[25] = c.PRE_aff_mode[1]; out[26] = c.PRE_b0_mode;
  int idx = host + target*n; for (int i = 0; i < 6; i++) { for (int j = 0; j < 6; j++) { adH36[6*i+j] = fs->ef->adHost[idx](i, j); adT36[6*i+j] = fs->ef->adTarget[idx](i, j); } adHTdelta6[i] = fs->ef->adHTdeltaF[idx][i]; }
}
// keyframe hand-over.  `selected[p]` stands for the pointer-graph predicate (isOOB || host flagged) && isInlierNew (FullSystem.cpp:763-767): the shim makes exactly the
// selected points satisfy it (lastResiduals[0] = OOB, enough good residuals) and then runs the reference's own flagPointsForRemoval.
void ref_ba_flag_points(void* p, const int* selected, int* status) {
  RefBA* b = (RefBA*)p; FullSystem* fs = b->fs;
  for (size_t i = 0; i < b->pts.size(); i++) { PointHessian* ph = b->pts[i];
    if (selected[i]) { ph->lastResiduals[0].second = ResState::OOB; if (ph->numGoodResiduals < setting_minGoodResForMarg) ph->numGoodResiduals = setting_minGoodResForMarg; }
    else { ph->lastResiduals[0].second = ResState::IN; ph->lastResiduals[1].second = ResState::IN; } }
  // flagPointsForRemoval skips the newest keyframe's points (:749): so does the flat interface's caller
  fs->flagPointsForRemoval(); refresh_dead(b);
  for (size_t i = 0; i < b->pts.size(); i++) status[i] = (b->pts[i]->efPoint->stateFlag == EFPointStatus::PS_MARGINALIZE) ? 2 : (b->pts[i]->efPoint->stateFlag == EFPointStatus::PS_DROP ? 1 : 0);
}
void ref_ba_marginalize_points(void* p) { ((RefBA*)p)->fs->ef->marginalizePointsF(); }
void ref_ba_drop_points(void* p) { ((RefBA*)p)->fs->ef->dropPointsF(); }
void ref_ba_marginalize_frame(void* p, int idx) { RefBA* b = (RefBA*)p; FullSystem* fs = b->fs; FrameHessian* fh = fs->frameHessians[idx];
  fh->pointHessians.clear();                                                                                  // marginalizeFrame asserts the frame's points are gone (they were flagged + marginalised/dropped before)
  // FullSystem::marginalizeFrame deletes the FrameHessian; ours belongs to the caller's RefFrame: run the numeric part (EnergyFunctional::marginalizeFrame) and the list surgery only
  fs->ef->marginalizeFrame(fh->efFrame);
  for (size_t i = idx; i + 1 < fs->frameHessians.size(); i++) fs->frameHessians[i] = fs->frameHessians[i+1]; fs->frameHessians.pop_back();
  for (size_t i = 0; i < fs->frameHessians.size(); i++) fs->frameHessians[i]->idx = (int)i;
  fh->efFrame = 0;
}
int  ref_ba_dim(void* p) { return (int)((RefBA*)p)->fs->ef->HM.rows(); }
void ref_ba_get_prior(void* p, double* HM, double* bM) { RefBA* b = (RefBA*)p; copy_mat(b->fs->ef->HM, HM); for (int i = 0; i < (int)b->fs->ef->bM.size(); i++) bM[i] = b->fs->ef->bM[i]; }
void ref_ba_get_res_to_zero(void* p, float* r2, int* isLin) { RefBA* b = (RefBA*)p; for (size_t i = 0; i < b->res.size(); i++) { if (b->dead[i]) { r2[2*i] = r2[2*i+1] = 0; isLin[i] = 0; continue; } r2[2*i] = b->res[i]->efResidual->res_toZeroF[0]; r2[2*i+1] = b->res[i]->efResidual->res_toZeroF[1]; isLin[i] = b->res[i]->efResidual->isLinearized ? 1 : 0; } }

// ---------------------------------------------------------------------------------------------- the whole pipeline: FullSystem::addActiveFrame per frame (main.cpp:466-509)
// BASELINE.json config #1 ("first 200 frames, single sequence, CPU reference path, pose + energy dump") runs HERE on the reference's own FullSystem: makeImages ->
// trackNewCoarse (hypotheses, trackNewestCoarse, reprojectMap, structPoseEstimation) -> makeKeyFrame / makeNonKeyFrame (traceNewCoarse, activatePointsMT, optimize,
// marginalisation, makeNewTraces, setCoarseTrackingRef).  The shim only feeds the queues main.cpp feeds (image, LiDAR pixels {Ku, Kv, depth}) and reads state back.
struct RefSys { FullSystem* fs; int next_id; };
void* ref_sys_create(int mode_perfect_images) {
  setting_logStuff = false; multiThreading = false; setting_debugout_runquiet = true; disableAllDisplay = true; setting_render_display3D = false; setting_render_displayDepth = false;
  setting_render_displayVideo = false; setting_render_displayResidual = false; setting_render_renderWindowFrames = false; setting_render_plotTrackingFull = false; setting_render_displayCoarseTrackingFull = false;
  if (mode_perfect_images) { setting_photometricCalibration = 0; setting_affineOptModeA = -1; setting_affineOptModeB = -1; setting_minGradHistAdd = 3; }   // main.cpp:461-468 (mode 2)
  else { setting_photometricCalibration = 0; setting_affineOptModeA = 0; setting_affineOptModeB = 0; }                                                    // mode 1
  RefSys* s = new RefSys(); s->fs = new FullSystem(); s->fs->linearizeOperation = true; s->next_id = 0; return s;
}
// ~FullSystem double-frees parts of its graph when the run ended between keyframes (it was only ever run at process exit): stop the mapping thread and leak the rest
void ref_sys_destroy(void* p) { RefSys* s = (RefSys*)p; if (!s) return; s->fs->blockUntilMappingIsFinished(); delete s; }
void ref_srand(unsigned seed) { srand(seed); }
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>
static void ref_segv(int sig) { void* bt[64]; int n = backtrace(bt, 64); backtrace_symbols_fd(bt, n, 2); _exit(139); }
void ref_debug_install_segv_handler() { signal(SIGSEGV, ref_segv); signal(SIGABRT, ref_segv); }
// cloud_px: n rows {Ku, Kv, depth} in the (cropped) image — what main.cpp:785-855 pushes into qCloudPixel
int ref_sys_add_frame(void* p, const float* image, float exposure, double timestamp, const double* cloud_px, int n) {
  RefSys* s = (RefSys*)p; FullSystem* fs = s->fs;
  std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d> > v(n); for (int i = 0; i < n; i++) v[i] = Eigen::Vector3d(cloud_px[3*i], cloud_px[3*i+1], cloud_px[3*i+2]);
  fs->qCloudPixel.push(v); fs->qTimeLidarCloud.push(timestamp); fs->qTimeImg.push(timestamp);
  ImageAndExposure* img = new ImageAndExposure(wG[0], hG[0], timestamp); img->exposure_time = exposure; memcpy(img->image, image, sizeof(float)*(size_t)wG[0]*hG[0]);
  fs->addActiveFrame(img, s->next_id++);
  delete img; fs->qCloudPixel.pop(); fs->qTimeLidarCloud.pop(); fs->qTimeImg.pop();
  return fs->isLost ? -1 : (fs->initFailed ? -2 : 0);
}
int ref_sys_num_frames(void* p) { return (int)((RefSys*)p)->fs->allFrameHistory.size(); }
int ref_sys_num_keyframes(void* p) { return (int)((RefSys*)p)->fs->allKeyFramesHistory.size(); }
// per frame: camToWorld (7), aff_g2l (2), flags {poseValid, is keyframe (trackingRef == 0 or in allKeyFramesHistory), trackingRef id}
void ref_sys_get_frame(void* p, int i, double T7[7], double ab[2], int flags[3]) {
  FullSystem* fs = ((RefSys*)p)->fs; FrameShell* sh = fs->allFrameHistory[i];
  se3_to(sh->camToWorld, T7); ab[0] = sh->aff_g2l.a; ab[1] = sh->aff_g2l.b; flags[0] = sh->poseValid ? 1 : 0; flags[1] = 0;
  for (FrameShell* k : fs->allKeyFramesHistory) if (k == sh) flags[1] = 1;
  flags[2] = sh->trackingRef ? sh->trackingRef->id : -1;
}
void ref_sys_get_last_rmse(void* p, double out5[5]) { FullSystem* fs = ((RefSys*)p)->fs; for (int i = 0; i < 5; i++) out5[i] = fs->lastCoarseRMSE[i]; }
int ref_sys_window(void* p, int* kf_ids, int* n_points) { FullSystem* fs = ((RefSys*)p)->fs; int n = (int)fs->frameHessians.size();
  for (int i = 0; i < n; i++) { if (kf_ids) kf_ids[i] = fs->frameHessians[i]->shell->id; if (n_points) n_points[i] = (int)fs->frameHessians[i]->pointHessians.size(); } return n; }

// ---- state of the running system BEFORE the next addActiveFrame = the inputs of FullSystem::trackNewCoarse for that frame (teacher-forced replay of the tracker on other arms)
static CoarseTracker* next_tracker(FullSystem* fs) {          // the swap of addActiveFrame (FullSystem.cpp:853-859) has not happened yet: predict it
  return (fs->coarseTracker_forNewKF->refFrameID > fs->coarseTracker->refFrameID) ? fs->coarseTracker_forNewKF : fs->coarseTracker;
}
int ref_sys_tracker_info(void* p, int* ref_shell_id, double ref_ab[2], float* ref_exposure, int* pc_n /*PYR_LEVELS*/, double* firstCoarseRMSE) {
  FullSystem* fs = ((RefSys*)p)->fs; CoarseTracker* t = next_tracker(fs); if (!t->lastRef) return -1;
  *ref_shell_id = t->lastRef->shell->id; ref_ab[0] = t->lastRef_aff_g2l.a; ref_ab[1] = t->lastRef_aff_g2l.b; *ref_exposure = t->lastRef->ab_exposure;
  for (int l = 0; l < pyrLevelsUsed; l++) pc_n[l] = t->pc_n[l]; *firstCoarseRMSE = t->firstCoarseRMSE; return 0;
}
void ref_sys_tracker_K(void* p, float K4[4], double calib_value_scaled[4]) {      // level-0 intrinsics of the tracker that will be used (makeK at its keyframe) and the current CalibHessian (the Reprojector reads it)
  FullSystem* fs = ((RefSys*)p)->fs; CoarseTracker* t = next_tracker(fs); K4[0] = t->fx[0]; K4[1] = t->fy[0]; K4[2] = t->cx[0]; K4[3] = t->cy[0];
  for (int i = 0; i < 4; i++) calib_value_scaled[i] = fs->Hcalib.value_scaled[i];
}
void ref_sys_tracker_cloud(void* p, int lvl, float* u, float* v, float* id, float* color) {
  CoarseTracker* t = next_tracker(((RefSys*)p)->fs); for (int i = 0; i < t->pc_n[lvl]; i++) { u[i] = t->pc_u[lvl][i]; v[i] = t->pc_v[lvl][i]; id[i] = t->pc_idepth[lvl][i]; color[i] = t->pc_color[lvl][i]; }
}
// history as trackNewCoarse will see it for the NEXT frame (allFrameHistory gets the new shell first, so "size()-2" there is back() here)
void ref_sys_history(void* p, double sprelast7[7], double slast7[7], double lastF7[7], double aff_last[2], double lastCoarseRMSE[5], int* n_history) {
  FullSystem* fs = ((RefSys*)p)->fs; int n = (int)fs->allFrameHistory.size(); *n_history = n;
  FrameShell* slast = fs->allFrameHistory[n-1]; FrameShell* sprelast = fs->allFrameHistory[n >= 2 ? n-2 : n-1];
  se3_to(sprelast->camToWorld, sprelast7); se3_to(slast->camToWorld, slast7); se3_to(next_tracker(fs)->lastRef->shell->camToWorld, lastF7);
  aff_last[0] = slast->aff_g2l.a; aff_last[1] = slast->aff_g2l.b; for (int i = 0; i < 5; i++) lastCoarseRMSE[i] = fs->lastCoarseRMSE[i];
}
int ref_sys_map_size(void* p, int* nKF) { FullSystem* fs = ((RefSys*)p)->fs; *nKF = (int)fs->frameHessians.size(); int n = 0; for (FrameHessian* fh : fs->frameHessians) n += (int)fh->pointHessians.size(); return n; }
// active map = what Reprojector(&Hcalib, fh, frameHessians) walks: keyframes in window order, their ACTIVE PointHessians {u, v, idepth_scaled, host index, type}
void ref_sys_map(void* p, int* kf_shell_ids, double* kf_T7, double* kf_ab, float* kf_exposure, float* pts5 /*n x {u,v,idepth,host,type}*/) {
  FullSystem* fs = ((RefSys*)p)->fs; int k = 0;
  for (size_t h = 0; h < fs->frameHessians.size(); h++) { FrameHessian* fh = fs->frameHessians[h]; kf_shell_ids[h] = fh->shell->id; se3_to(fh->shell->camToWorld, kf_T7 + 7*h);
    kf_ab[2*h] = fh->aff_g2l().a; kf_ab[2*h+1] = fh->aff_g2l().b; kf_exposure[h] = fh->ab_exposure;
    for (PointHessian* ph : fh->pointHessians) { pts5[5*k] = ph->u; pts5[5*k+1] = ph->v; pts5[5*k+2] = ph->idepth_scaled; pts5[5*k+3] = (float)h; pts5[5*k+4] = (ph->type == PointHessian::EDGELET) ? 1.f : 0.f; k++; } }
}
void ref_sys_get_track_result(void* p, int i, double camToTrackingRef7[7]) { se3_to(((RefSys*)p)->fs->allFrameHistory[i]->camToTrackingRef, camToTrackingRef7); }

// ---------------------------------------------------------------------------------------------- Reprojector::reprojectMap (+ CoarseTracker::structPoseEstimation) on flat inputs
// Same inputs as orc_reproject_map: keyframes (RefFrame handles in window order) with camToWorld / aff, the current frame with its pose, the ACTIVE map points
// {u, v, idepth, host, type}.  `seed`: srand(seed) right before the Reprojector is constructed — its grid order is std::random_shuffle over rand() (Reprojector.cpp:107).
// refine != 0: also runs structPoseEstimation on the matches (FullSystem.cpp:482-488) and returns the refined camToWorld in cur_T7_io.
int ref_reproject_map(int nH, void** kf_frames, const double* kf_T7, const double* kf_ab, void* cur_frame, double* cur_T7_io, const double* cur_ab, int nP, const float* pts5,
                      unsigned seed, int refine, int* out_pt, double* out_px) {
  std::vector<FrameHessian*> fhs; RefFrame* C = (RefFrame*)cur_frame;
  for (int h = 0; h < nH; h++) { RefFrame* F = (RefFrame*)kf_frames[h]; drop_points(F); F->fh->idx = h;
    F->shell->camToWorld = se3_from(kf_T7 + 7*h); F->shell->aff_g2l = AffLight(kf_ab[2*h], kf_ab[2*h+1]); F->shell->poseValid = true; fhs.push_back(F->fh); }
  C->shell->camToWorld = se3_from(cur_T7_io); C->shell->aff_g2l = AffLight(cur_ab[0], cur_ab[1]);
  std::vector<PointHessian*> all;
  for (int i = 0; i < nP; i++) { RefFrame* F = (RefFrame*)kf_frames[(int)pts5[5*i+3]];
    ImmaturePoint* ip = new ImmaturePoint((int)pts5[5*i], (int)pts5[5*i+1], F->fh, 0, g_calib); ip->idepth_min = ip->idepth_max = pts5[5*i+2];
    ip->type = pts5[5*i+4] != 0 ? ImmaturePoint::EDGELET : ImmaturePoint::CORNER; F->ips.push_back(ip);
    PointHessian* ph = new PointHessian(ip, g_calib); ph->u = pts5[5*i]; ph->v = pts5[5*i+1]; ph->setIdepth(pts5[5*i+2]); ph->setPointStatus(PointHessian::ACTIVE); ph->idx = i;
    F->phs.push_back(ph); F->fh->pointHessians.push_back(ph); all.push_back(ph); }
  std::vector<std::pair<PointHessian*, Eigen::Vector2d> > overlap;
  srand(seed);
  { Reprojector rp(g_calib, C->fh, fhs); rp.reprojectMap(C->fh, overlap); }
  for (size_t k = 0; k < overlap.size(); k++) { out_pt[k] = overlap[k].first->idx; out_px[2*k] = overlap[k].second[0]; out_px[2*k+1] = overlap[k].second[1]; }
  if (refine) { CoarseTracker* t = new CoarseTracker(wG[0], hG[0]); t->makeK(g_calib); t->debugPrint = false; SE3 c2w = C->shell->camToWorld; t->structPoseEstimation(c2w, overlap); se3_to(c2w, cur_T7_io); delete t; }
  for (int h = 0; h < nH; h++) drop_points((RefFrame*)kf_frames[h]);
  return (int)overlap.size();
}

// debugging aid: reprojectMap + structPoseEstimation on the LIVE window of a running system for a new image at a given pose (what trackNewCoarse does after tracking)
int ref_sys_debug_refine(void* p, const float* image, double* T7_io, const double* ab, unsigned seed, int* n_active_nonactive /*2*/) {
  FullSystem* fs = ((RefSys*)p)->fs;
  FrameHessian* fh = new FrameHessian(); FrameShell* sh = new FrameShell(); sh->id = 100000; fh->shell = sh; fh->ab_exposure = 1;
  std::vector<float> c(image, image + (size_t)wG[0]*hG[0]); fh->makeImages(c.data(), &fs->Hcalib);
  sh->camToWorld = se3_from(T7_io); sh->aff_g2l = AffLight(ab[0], ab[1]);
  int na = 0, nn = 0; for (FrameHessian* f : fs->frameHessians) for (PointHessian* ph : f->pointHessians) { if (ph->status == PointHessian::ACTIVE) na++; else nn++; }
  n_active_nonactive[0] = na; n_active_nonactive[1] = nn;
  std::vector<std::pair<PointHessian*, Eigen::Vector2d> > overlap; srand(seed);
  { Reprojector rp(&fs->Hcalib, fh, fs->frameHessians); rp.reprojectMap(fh, overlap); }
  SE3 c2w = sh->camToWorld; next_tracker(fs)->structPoseEstimation(c2w, overlap); se3_to(c2w, T7_io);
  return (int)overlap.size();
}

// ---- timing loop of the REFERENCE CPU arm (bench.py cpu_baseline / --impl reference): n_frames x { FrameHessian::makeImages ; CoarseTracker::trackNewestCoarse } in one call
// on one host thread, the reference's own code (a FrameHessian is allocated and freed per frame, as addActiveFrame does).  One tracker object per thread; the calibration
// globals are only read.  Same contract as orc_bench_track_loop.
#include <time.h>
int ref_bench_track_loop(void* t, const float* const* imgs, int n_imgs, int start, int n_frames, const double* inits7, double budget_s, double* last_T7, int* good_count) {
  CoarseTracker* T = (CoarseTracker*)t; timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
  Vec5 mr; for (int i = 0; i < 5; i++) mr[i] = NAN;
  int done = 0, good = 0; FrameShell shell; std::vector<float> buf((size_t)wG[0]*hG[0]);
  for (int f = 0; f < n_frames; f++) {
    FrameHessian* fh = new FrameHessian(); fh->shell = &shell; fh->ab_exposure = 1;
    memcpy(buf.data(), imgs[(start + f) % n_imgs], buf.size()*sizeof(float));
    fh->makeImages(buf.data(), g_calib);
    SE3 s = se3_from(inits7 + 7*(size_t)f); AffLight aff(0, 0);
    if (T->trackNewestCoarse(fh, s, aff, pyrLevelsUsed-1, mr, 0)) good++;
    if (last_T7) se3_to(s, last_T7);
    fh->efFrame = 0; delete fh; done++;
    if (budget_s > 0) { timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1); if ((t1.tv_sec - t0.tv_sec) + 1e-9*(t1.tv_nsec - t0.tv_nsec) > budget_s) break; }
  }
  if (good_count) *good_count = good;
  return done;
}

// ---------------------------------------------------------------------------------------------- ingest: Undistort (util/Undistort.cpp) — photometric processFrame + geometric crop-remap
// config_text: the content of a calibration file (reference format, e.g. calib/KITTI/00.txt: "Pinhole fx fy cx cy 0 / wOrg hOrg / crop / w h").  It is written to a temporary
// file because Undistort::getUndistorterForFile reads a path.  Returns a handle; out: sizes and the rectified K.
void* ref_undistort_create(const char* config_text, int wh_org[2], int wh[2], double K4[4]) {
  char path[] = "/tmp/sdv_ref_calib_XXXXXX"; int fd = mkstemp(path); if (fd < 0) return 0;
  FILE* f = fdopen(fd, "w"); fputs(config_text, f); fclose(f);
  fflush(stdout);
  Undistort* u = Undistort::getUndistorterForFile(path, "", "");
  remove(path); if (!u) return 0;
  wh_org[0] = u->getOriginalSize()[0]; wh_org[1] = u->getOriginalSize()[1]; wh[0] = u->getSize()[0]; wh[1] = u->getSize()[1];
  Mat33 K = u->getK(); K4[0] = K(0, 0); K4[1] = K(1, 1); K4[2] = K(0, 2); K4[3] = K(1, 2);
  return u;
}
void ref_undistort_destroy(void* p) { delete (Undistort*)p; }
void ref_undistort_maps(void* p, float* remapX, float* remapY) { Undistort* u = (Undistort*)p; size_t n = (size_t)u->w*u->h; memcpy(remapX, u->remapX, n*sizeof(float)); memcpy(remapY, u->remapY, n*sizeof(float)); }
int  ref_undistort_passthrough(void* p) { return ((Undistort*)p)->passthrough ? 1 : 0; }
// Undistort::undistort<unsigned char> (:341-435): mono8 wire image (sensor_msgs/Image, DatasetReader.h:152-155) -> float image the pipeline tracks
void ref_undistort_apply_u8(void* p, const unsigned char* raw, float exposure, float* out) {
  Undistort* u = (Undistort*)p; MinimalImageB img(u->wOrg, u->hOrg); memcpy(img.data, raw, (size_t)u->wOrg*u->hOrg);
  ImageAndExposure* r = u->undistort<unsigned char>(&img, exposure, 0.0, 1.0f);
  memcpy(out, r->image, sizeof(float)*(size_t)u->w*u->h); delete r;
}

// bench arm with the ingest in front: per frame Undistort::undistort<unsigned char> (photometric stage + crop-remap, what src/main.cpp:537-560 runs on every image message)
// -> FrameHessian::makeImages -> trackNewestCoarse.  `u` must be private to the calling thread (undistort<> writes into the undistorter's own output buffer).
int ref_bench_ingest_track_loop(void* t, void* u, const unsigned char* const* raws, int n_imgs, int start, int n_frames, const double* inits7, double budget_s, double* last_T7, int* good_count) {
  CoarseTracker* T = (CoarseTracker*)t; Undistort* U = (Undistort*)u; timespec t0; clock_gettime(CLOCK_MONOTONIC, &t0);
  Vec5 mr; for (int i = 0; i < 5; i++) mr[i] = NAN;
  int done = 0, good = 0; FrameShell shell;
  for (int f = 0; f < n_frames; f++) {
    MinimalImageB raw(U->wOrg, U->hOrg); memcpy(raw.data, raws[(start + f) % n_imgs], (size_t)U->wOrg*U->hOrg);     // the decoded sensor_msgs/Image (DatasetReader.h:152-155)
    ImageAndExposure* img = U->undistort<unsigned char>(&raw, 1.0f, 0.0, 1.0f);
    FrameHessian* fh = new FrameHessian(); fh->shell = &shell; fh->ab_exposure = img->exposure_time;
    fh->makeImages(img->image, g_calib);
    SE3 s = se3_from(inits7 + 7*(size_t)f); AffLight aff(0, 0);
    if (T->trackNewestCoarse(fh, s, aff, pyrLevelsUsed-1, mr, 0)) good++;
    if (last_T7) se3_to(s, last_T7);
    fh->efFrame = 0; delete fh; delete img; done++;
    if (budget_s > 0) { timespec t1; clock_gettime(CLOCK_MONOTONIC, &t1); if ((t1.tv_sec - t0.tv_sec) + 1e-9*(t1.tv_nsec - t0.tv_nsec) > budget_s) break; }
  }
  if (good_count) *good_count = good;
  return done;
}

// ---------------------------------------------------------------------------------------------- immature points (FullSystem/ImmaturePoint.cpp): constructor :8-36, traceOn :50-352
// flat record = orc::ImmPt of oracle/orc_trace.cpp: u v idepth_min idepth_max color[8] weights[8] gradH[4] energyTH quality lastTraceUV[2] lastTracePixelInterval | status
void* ref_immature_create(void* host, int u, int v, float type) { return new ImmaturePoint(u, v, ((RefFrame*)host)->fh, type, g_calib); }
void  ref_immature_destroy(void* p) { delete (ImmaturePoint*)p; }
void  ref_immature_get(void* pp, float* o29, int* status) {
  ImmaturePoint* p = (ImmaturePoint*)pp; int k = 0;
  o29[k++] = p->u; o29[k++] = p->v; o29[k++] = p->idepth_min; o29[k++] = p->idepth_max;
  for (int i = 0; i < 8; i++) o29[k++] = p->color[i];
  for (int i = 0; i < 8; i++) o29[k++] = p->weights[i];
  o29[k++] = p->gradH(0, 0); o29[k++] = p->gradH(0, 1); o29[k++] = p->gradH(1, 0); o29[k++] = p->gradH(1, 1);
  o29[k++] = p->energyTH; o29[k++] = p->quality; o29[k++] = p->lastTraceUV[0]; o29[k++] = p->lastTraceUV[1]; o29[k++] = p->lastTracePixelInterval;
  *status = (int)p->lastTraceStatus;
}
void  ref_immature_set_range(void* pp, float idepth_min, float idepth_max, int status) {
  ImmaturePoint* p = (ImmaturePoint*)pp; p->idepth_min = idepth_min; p->idepth_max = idepth_max; p->lastTraceStatus = (ImmaturePointStatus)status; }
int   ref_immature_trace(void* pp, void* frame, const float* KRKi9, const float* Kt3, const float* aff2) {
  Mat33f KRKi; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) KRKi(i, j) = KRKi9[i*3+j];
  Vec3f Kt(Kt3[0], Kt3[1], Kt3[2]); Vec2f aff(aff2[0], aff2[1]);
  return (int)((ImmaturePoint*)pp)->traceOn(((RefFrame*)frame)->fh, KRKi, Kt, aff, g_calib, false);
}

// ---- activation: FullSystem::optimizeImmaturePoint (FullSystemOptPoint.cpp:18-183) on the window of a RefBA (ref_ba_* above; ref_ba_init must have run: targetPrecalc)
// pre14 out: for every target (frameHessians without the host, window order) PRE_RTll[9] PRE_tTll[3] PRE_aff_mode[2]; calib6: fxl fyl cxl cyl fxli fyli
int ref_ba_immature_pre(void* p, int host, float* pre14, float* calib6) {
  RefBA* b = (RefBA*)p; FullSystem* fs = b->fs; FrameHessian* H = fs->frameHessians[host]; int k = 0;
  for (FrameHessian* fh : fs->frameHessians) { if (fh == H) continue; const FrameFramePrecalc& c = H->targetPrecalc[fh->idx]; float* o = pre14 + 14*k++;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o[i*3+j] = c.PRE_RTll(i, j);
    for (int i = 0; i < 3; i++) o[9+i] = c.PRE_tTll[i]; o[12] = c.PRE_aff_mode[0]; o[13] = c.PRE_aff_mode[1]; }
  CalibHessian& C = fs->Hcalib; calib6[0] = C.fxl(); calib6[1] = C.fyl(); calib6[2] = C.cxl(); calib6[3] = C.cyl(); calib6[4] = C.fxli(); calib6[5] = C.fyli();
  return k;
}
// one candidate: constructed at (u,v) on the host (the constructor's colours / weights / energyTH), range and sensor flag set, then the reference's own function.
// returns 0 / -1 / 1 like orc::optimizeImmaturePoint; res_state: final state_state of the temporary residuals (window order without the host)
int ref_ba_optimize_immature(void* p, int host, int u, int v, float idepth_min, float idepth_max, int isFromSensor, int minObs, float* idepth_out, int* res_state) {
  RefBA* b = (RefBA*)p; FullSystem* fs = b->fs; FrameHessian* H = fs->frameHessians[host];
  ImmaturePoint* ip = new ImmaturePoint(u, v, H, 1.0f, &fs->Hcalib); ip->idepth_min = idepth_min; ip->idepth_max = idepth_max; ip->isFromSensor = isFromSensor != 0; ip->idepth_fromSensor = 0;
  ip->type = ImmaturePoint::CORNER; ip->score = 0; ip->idxInImmaturePoints = 0; ip->idepth_GT = 0; ip->lastTraceStatus = IPS_GOOD; ip->lastTraceUV = Vec2f(0, 0); ip->gradH_ev = Vec2f(0, 0);
  int nres = (int)fs->frameHessians.size() - 1; std::vector<ImmaturePointTemporaryResidual> tr(nres);
  PointHessian* ph = fs->optimizeImmaturePoint(ip, minObs, tr.data());
  for (int i = 0; i < nres; i++) res_state[i] = (int)tr[i].state_state;
  int rc;
  if (ph == 0) rc = 0; else if (ph == (PointHessian*)((long)(-1))) rc = -1;
  else { rc = 1; *idepth_out = ph->idepth; ph->release(); delete ph; }
  delete ip;
  return rc;
}

// ---------------------------------------------------------------------------------------------- candidate management at keyframe rate (SURVEY §8f rank 4 + the caller half of rank 2)
// PixelSelector (FullSystem/PixelSelector2.cpp), FullSystem::makeNewTraces / shiTomasiScore (FullSystem.cpp:1273-1356, 1540-1583), CoarseDistanceMap (CoarseTracker.cpp:1139-1282).
// The reference's selector reads thsSmoothed past its (h/32) rows for the last image rows and never initialises that tail (PixelSelector2.cpp:18-20, 268): the shim zeroes the
// selector's heap arrays once after construction so the pin is deterministic (a fresh heap block reads as zero too).
static void selector_zero(PixelSelector* s) { int n = (wG[0]/32)*(hG[0]/32)+100; memset(s->ths, 0, n*sizeof(float)); memset(s->thsSmoothed, 0, n*sizeof(float));
  memset(s->gradHist, 0, sizeof(int)*100*(1+wG[0]/32)*(1+hG[0]/32)); }
void* ref_selector_create() { PixelSelector* s = new PixelSelector(wG[0], hG[0]); selector_zero(s); return s; }      // srand(3141592) + w*h rand() calls (PixelSelector2.cpp:14-16)
void  ref_selector_destroy(void* s) { delete (PixelSelector*)s; }
void  ref_selector_random_pattern(void* s, unsigned char* out) { memcpy(out, ((PixelSelector*)s)->randomPattern, (size_t)wG[0]*hG[0]); }
void  ref_selector_set_potential(void* s, int p) { ((PixelSelector*)s)->currentPotential = p; }
int   ref_selector_get_potential(void* s) { return ((PixelSelector*)s)->currentPotential; }
void  ref_selector_make_hists(void* s, void* frame, float* ths, float* thsSmoothed) { PixelSelector* S = (PixelSelector*)s; S->makeHists(((RefFrame*)frame)->fh);
  int n = (wG[0]/32)*(hG[0]/32); if (ths) memcpy(ths, S->ths, n*sizeof(float)); if (thsSmoothed) memcpy(thsSmoothed, S->thsSmoothed, n*sizeof(float)); }
typedef std::vector<Eigen::Vector3d, Eigen::aligned_allocator<Eigen::Vector3d> > CloudPx;
static CloudPx cloud_of(const double* c3, int n) { CloudPx v(n); for (int i = 0; i < n; i++) v[i] = Eigen::Vector3d(c3[3*i], c3[3*i+1], c3[3*i+2]); return v; }
// PixelSelector::select (cloud3 == NULL) / selectFromLidar: one pass at potential `pot`; makeHists must have run on this frame
void  ref_selector_select(void* s, void* frame, float* map_out, int pot, float thFactor, const double* cloud3, int n, int n3[3]) {
  PixelSelector* S = (PixelSelector*)s; FrameHessian* fh = ((RefFrame*)frame)->fh; Eigen::Vector3i r;
  if (cloud3) { CloudPx v = cloud_of(cloud3, n); r = S->selectFromLidar(fh, map_out, pot, thFactor, v); } else r = S->select(fh, map_out, pot, thFactor);
  n3[0] = r[0]; n3[1] = r[1]; n3[2] = r[2];
}
int   ref_selector_make_maps(void* s, void* frame, float* map_out, float density, int recursionsLeft, float thFactor, const double* cloud3, int n) {
  PixelSelector* S = (PixelSelector*)s; FrameHessian* fh = ((RefFrame*)frame)->fh; S->gradHistFrame = 0;
  if (cloud3) { CloudPx v = cloud_of(cloud3, n); return S->makeMapsFromLidar(fh, map_out, density, recursionsLeft, false, thFactor, v); }
  return S->makeMaps(fh, map_out, density, recursionsLeft, false, thFactor);
}
void* ref_ba_selector(void* p) { PixelSelector* s = ((RefBA*)p)->fs->pixelSelector; selector_zero(s); return s; }
float ref_shi_tomasi(void* p, void* frame, int u, int v) { return ((RefBA*)p)->fs->shiTomasiScore(((RefFrame*)frame)->fh->dI, u, v); }
// FullSystem::makeNewTraces on `frame` with the LiDAR pixels cloud3 {Ku, Kv, depth}; lidar box = the members main.cpp fills (left/right/up/down); selectionMap_io = the
// FullSystem's persistent monocular selection map (w*h floats).  out rows: {u, v, my_type, score, idepth_fromSensor, isFromSensor, type}; returns the number of immature points.
int ref_make_new_traces(void* p, void* frame, const double* cloud3, int n, const int lrud[4], int addFeaturePoint, float desiredImmatureDensity, float* selectionMap_io, float* out7, int cap) {
  RefBA* b = (RefBA*)p; FullSystem* fs = b->fs; RefFrame* F = (RefFrame*)frame; FrameHessian* fh = F->fh; size_t wh = (size_t)wG[0]*hG[0];
  fs->left = lrud[0]; fs->right = lrud[1]; fs->up = lrud[2]; fs->down = lrud[3]; fs->addFeaturePoint = addFeaturePoint != 0; float keep = setting_desiredImmatureDensity; setting_desiredImmatureDensity = desiredImmatureDensity;
  memcpy(fs->selectionMap, selectionMap_io, wh*sizeof(float));
  F->shell->timestamp = 1.0; fs->qCloudPixel.push(cloud_of(cloud3, n)); fs->qTimeLidarCloud.push(1.0);
  fs->pixelSelector->gradHistFrame = 0;
  fs->makeNewTraces(fh, 0);
  fs->qCloudPixel.pop(); fs->qTimeLidarCloud.pop(); setting_desiredImmatureDensity = keep;
  memcpy(selectionMap_io, fs->selectionMap, wh*sizeof(float));
  int m = 0;
  for (ImmaturePoint* ip : fh->immaturePoints) { if (m < cap) { float* o = out7 + 7*m; o[0] = ip->u; o[1] = ip->v; o[2] = ip->my_type; o[3] = ip->isFromSensor ? ip->score : 0.f;
      o[4] = ip->isFromSensor ? ip->idepth_fromSensor : 0.f; o[5] = ip->isFromSensor ? 1.f : 0.f; o[6] = ip->isFromSensor ? (float)(int)ip->type : -1.f; } m++; delete ip; }
  fh->immaturePoints.clear();
  return m;
}
// CoarseDistanceMap of the RefBA's FullSystem: makeK + makeDistanceMap(frameHessians, frameHessians[frame_idx]) — the window's ACTIVE PointHessians are the sources
void ref_distmap_make(void* p, int frame_idx) { FullSystem* fs = ((RefBA*)p)->fs; fs->coarseDistanceMap->makeK(&fs->Hcalib); fs->coarseDistanceMap->makeDistanceMap(fs->frameHessians, fs->frameHessians[frame_idx]); }
void ref_distmap_add(void* p, int u, int v) { ((RefBA*)p)->fs->coarseDistanceMap->addIntoDistFinal(u, v); }
void ref_distmap_get(void* p, float* out) { CoarseDistanceMap* d = ((RefBA*)p)->fs->coarseDistanceMap; memcpy(out, d->fwdWarpedIDDistFinal, sizeof(float)*(size_t)d->w[1]*d->h[1]); }
// KRKi = K[1] R Ki[0], Kt = K[1] t of host -> newest as FullSystem::activatePointsMT forms them (FullSystem.cpp:606-608)
void ref_distmap_geometry(void* p, int host_idx, int frame_idx, float KRKi9[9], float Kt3[3]) {
  FullSystem* fs = ((RefBA*)p)->fs; FrameHessian* host = fs->frameHessians[host_idx]; FrameHessian* newest = fs->frameHessians[frame_idx]; CoarseDistanceMap* d = fs->coarseDistanceMap;
  SE3 fhToNew = newest->PRE_worldToCam * host->PRE_camToWorld;
  Mat33f KRKi = (d->K[1] * fhToNew.rotationMatrix().cast<float>() * d->Ki[0]); Vec3f Kt = (d->K[1] * fhToNew.translation().cast<float>());
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) KRKi9[3*i+j] = KRKi(i, j); Kt3[i] = Kt[i]; }
}
// the candidate walk of activatePointsMT (FullSystem.cpp:600-671) over flat candidates, on the reference's own distance map (its BFS does the work); the ten lines of
// loop control are shim code.  cand rows {u, v, 0.5f*(idepth_max+idepth_min), my_type} grouped by host (window order)
void ref_activate_select(void* p, int frame_idx, int nHosts, const int* host_idx, const int* cand_begin, const float* cand4, float currentMinActDist, int* decision) {
  FullSystem* fs = ((RefBA*)p)->fs; CoarseDistanceMap* D = fs->coarseDistanceMap; FrameHessian* newestHs = fs->frameHessians[frame_idx];
  for (int hI = 0; hI < nHosts; hI++) { FrameHessian* host = fs->frameHessians[host_idx[hI]];
    SE3 fhToNew = newestHs->PRE_worldToCam * host->PRE_camToWorld;
    Mat33f KRKi = (D->K[1] * fhToNew.rotationMatrix().cast<float>() * D->Ki[0]); Vec3f Kt = (D->K[1] * fhToNew.translation().cast<float>());
    for (int c = cand_begin[hI]; c < cand_begin[hI+1]; c++) {
      Vec3f ptp = KRKi * Vec3f(cand4[4*c], cand4[4*c+1], 1) + Kt*cand4[4*c+2];
      int u = ptp[0] / ptp[2] + 0.5f; int v = ptp[1] / ptp[2] + 0.5f;
      if ((u > 0 && v > 0 && u < wG[1] && v < hG[1])) {
        float dist = D->fwdWarpedIDDistFinal[u+wG[1]*v] + (ptp[0]-floorf((float)(ptp[0])));
        if (dist >= currentMinActDist*cand4[4*c+3]) { D->addIntoDistFinal(u, v); decision[c] = 1; } else decision[c] = 0;
      } else decision[c] = -1;
    } }
}

// ---------------------------------------------------------------------------------------------- LiDAR front-end of the ROS node (src/main.cpp:537-858), compiled unmodified (oracle/Makefile)
// main.cpp keeps its state in globals (range / label / ground images, fullCloud, segmentedCloud) and reads the extrinsics / intrinsics from the global FullSystem.
// The shim feeds lidarCloudHandler one decoded XYZI sweep (the stand-in PointCloud2 carries the rows pcl::fromROSMsg would decode) and reads the results back.
void ref_unavailable() { fprintf(stderr, "oracle/_ref: a GUI / image-decoding entry of the reference was called; it is not part of this build\n"); abort(); }
}  // extern "C"
typedef pcl::PointXYZI RefPointType;
extern FullSystem* fullSystem;
extern pcl::PointCloud<RefPointType>::Ptr laserCloudIn, fullCloud, fullInfoCloud, groundCloud, segmentedCloud, segmentedCloudPure, outlierCloud;
extern cv::Mat rangeMat, labelMat, groundMat;
extern const int N_SCAN, Horizon_SCAN;
void allocateMemory(); void resetParameters(); void lidarCloudHandler(const sensor_msgs::PointCloud2ConstPtr& lidarCloudMsg);
void projectPointCloud(); void groundRemoval(); void cloudSegmentation();
extern "C" {
static bool g_lidar_ready = false;
// Rlc (row-major 3x3), tlc: LiDAR -> camera; K4 = fx fy cx cy (FullSystem members main.cpp:368-377 fills from the sensor file); lrud_io = FullSystem::left/right/up/down (running box)
// out rows {Ku, Kv, depth}; returns the number of pixels pushed into qCloudPixel (or -1 if cap is too small); flags_out = {addFeaturePoint, numGround, numAll, size of segmentedCloud}
int ref_lidar_handler(void* ba, const float* xyzi, int n, const double* Rlc9, const double* tlc3, const float* K4, int* lrud_io, double* out3, int cap, int* flags_out, float* range_out, int* label_out, signed char* ground_out) {
  RefBA* b = (RefBA*)ba; FullSystem* fs = b->fs; fullSystem = fs;
  if (!g_lidar_ready) { allocateMemory(); resetParameters(); g_lidar_ready = true; }
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) fs->Rlc(i, j) = Rlc9[3*i+j]; fs->tlc[i] = tlc3[i]; }
  fs->fx = K4[0]; fs->fy = K4[1]; fs->cx = K4[2]; fs->cy = K4[3]; fs->left = lrud_io[0]; fs->right = lrud_io[1]; fs->up = lrud_io[2]; fs->down = lrud_io[3];
  std::shared_ptr<sensor_msgs::PointCloud2> msg = std::make_shared<sensor_msgs::PointCloud2>(); msg->xyzi.assign(xyzi, xyzi + 4*(size_t)n); msg->header.stamp.t = 1.0;
  if (range_out || label_out || ground_out) {                                  // the images are reset at the end of the handler: run the three stages by hand first to read them
    pcl::fromROSMsg(*msg, *laserCloudIn); std::vector<int> idx; pcl::removeNaNFromPointCloud(*laserCloudIn, *laserCloudIn, idx);
    projectPointCloud(); groundRemoval(); cloudSegmentation();
    size_t m = (size_t)N_SCAN*Horizon_SCAN;
    if (range_out) memcpy(range_out, rangeMat.data, m*sizeof(float));
    if (label_out) memcpy(label_out, labelMat.data, m*sizeof(int));
    if (ground_out) memcpy(ground_out, groundMat.data, m);
    flags_out[3] = (int)segmentedCloud->points.size();
    resetParameters();
  }
  lidarCloudHandler(msg);
  const auto& v = fs->qCloudPixel.back(); int m = (int)v.size();
  lrud_io[0] = fs->left; lrud_io[1] = fs->right; lrud_io[2] = fs->up; lrud_io[3] = fs->down; flags_out[0] = fs->addFeaturePoint ? 1 : 0;
  if (m <= cap) for (int i = 0; i < m; i++) { out3[3*i] = v[i][0]; out3[3*i+1] = v[i][1]; out3[3*i+2] = v[i][2]; }
  while (!fs->qCloudPixel.empty()) fs->qCloudPixel.pop(); while (!fs->qTimeLidarCloud.empty()) fs->qTimeLidarCloud.pop();
  return m <= cap ? m : -1;
}
int ref_lidar_dims(int* n_scan, int* horizon) { *n_scan = N_SCAN; *horizon = Horizon_SCAN; return 0; }

// ---- sequence level: the PixelSelector / makeNewTraces state of a RUNNING system (tests/test_sequence_select.py).  makeNewTraces is the last thing makeKeyFrame does
// (FullSystem.cpp:1165), so after addActiveFrame returns the newest keyframe's immaturePoints ARE its output; the selector state is read before the call.
void ref_sys_selector_zero(void* p) { selector_zero(((RefSys*)p)->fs->pixelSelector); }
int  ref_sys_selector_state(void* p, float* selectionMap_out) { FullSystem* fs = ((RefSys*)p)->fs; if (selectionMap_out) memcpy(selectionMap_out, fs->selectionMap, sizeof(float)*(size_t)wG[0]*hG[0]); return fs->pixelSelector->currentPotential; }
void ref_sys_set_selection_map(void* p, const float* m) { memcpy(((RefSys*)p)->fs->selectionMap, m, sizeof(float)*(size_t)wG[0]*hG[0]); }
void ref_sys_set_lidar_state(void* p, const int lrud[4], int addFeaturePoint) { FullSystem* fs = ((RefSys*)p)->fs; fs->left = lrud[0]; fs->right = lrud[1]; fs->up = lrud[2]; fs->down = lrud[3]; fs->addFeaturePoint = addFeaturePoint != 0; }
// immature points of the newest keyframe: rows {u, v, my_type, score, idepth_fromSensor, isFromSensor, type}; returns the count (-1: no keyframe yet), *kf_shell_id = its frame id
int  ref_sys_newest_kf_immature(void* p, float* out7, int cap, int* kf_shell_id) {
  FullSystem* fs = ((RefSys*)p)->fs; if (fs->frameHessians.empty()) return -1; FrameHessian* fh = fs->frameHessians.back(); *kf_shell_id = fh->shell->id; int m = 0;
  for (ImmaturePoint* ip : fh->immaturePoints) { if (m < cap) { float* o = out7 + 7*m; o[0] = ip->u; o[1] = ip->v; o[2] = ip->my_type; o[3] = ip->isFromSensor ? ip->score : 0.f;
      o[4] = ip->isFromSensor ? ip->idepth_fromSensor : 0.f; o[5] = ip->isFromSensor ? 1.f : 0.f; o[6] = ip->isFromSensor ? (float)(int)ip->type : -1.f; } m++; }
  return m;
}

// makeNewTraces of the RUNNING system on a probe frame (a RefFrame of the image that is about to be added), with the system's live selector state; the state
// (currentPotential, selectionMap) is put back afterwards, so the run itself is not disturbed.  In this reference makeNewTraces runs BEFORE activatePointsMT
// (FullSystem.cpp:1080 vs :1102), so its output cannot be read off the keyframe after the fact: the probe is how the complete list is obtained.
// commit != 0 keeps the new state instead (used to check the state the real call leaves behind).
int ref_sys_probe_new_traces(void* p, void* frame, const double* cloud3, int n, float* out7, int cap, int commit) {
  FullSystem* fs = ((RefSys*)p)->fs; RefFrame* F = (RefFrame*)frame; FrameHessian* fh = F->fh; size_t wh = (size_t)wG[0]*hG[0];
  int pot = fs->pixelSelector->currentPotential; std::vector<float> keep(fs->selectionMap, fs->selectionMap + wh);
  F->shell->timestamp = 12345.0; fs->qCloudPixel.push(cloud_of(cloud3, n)); fs->qTimeLidarCloud.push(12345.0);
  // put the probe's data at the FRONT of the queues makeNewTraces reads (they are empty between frames in this harness)
  fs->pixelSelector->gradHistFrame = 0;
  fs->makeNewTraces(fh, 0);
  fs->qCloudPixel.pop(); fs->qTimeLidarCloud.pop(); fs->pixelSelector->gradHistFrame = 0;
  int m = 0;
  for (ImmaturePoint* ip : fh->immaturePoints) { if (m < cap) { float* o = out7 + 7*m; o[0] = ip->u; o[1] = ip->v; o[2] = ip->my_type; o[3] = ip->isFromSensor ? ip->score : 0.f;
      o[4] = ip->isFromSensor ? ip->idepth_fromSensor : 0.f; o[5] = ip->isFromSensor ? 1.f : 0.f; o[6] = ip->isFromSensor ? (float)(int)ip->type : -1.f; } m++; delete ip; }
  fh->immaturePoints.clear();
  if (!commit) { fs->pixelSelector->currentPotential = pot; memcpy(fs->selectionMap, keep.data(), wh*sizeof(float)); }
  return m;
}

// ---- sequence level, immature points: every candidate of every keyframe of the RUNNING system, before / after a frame (FullSystem::traceNewCoarse runs in makeNonKeyFrame
// and makeKeyFrame).  rec29 / status as ref_immature_get; returns the number of keyframes, counts[k] candidates of keyframe k (window order), shell ids in host_ids.
int ref_sys_immature_dump(void* p, int* host_ids, int* counts, float* rec29, int* status, int cap) {
  FullSystem* fs = ((RefSys*)p)->fs; int nk = 0, m = 0;
  for (FrameHessian* fh : fs->frameHessians) { host_ids[nk] = fh->shell->id; counts[nk] = (int)fh->immaturePoints.size(); nk++;
    for (ImmaturePoint* ip : fh->immaturePoints) { if (ip && m < cap) ref_immature_get(ip, rec29 + 29*(size_t)m, status + m); m++; } }
  return (m <= cap) ? nk : -1;
}
// KRKi, Kt of host keyframe (window index) -> a frame with camToWorld new_c2w7, formed like FullSystem::traceNewCoarse forms them (FullSystem.cpp:525-535) with the current CalibHessian
void ref_sys_trace_geometry(void* p, int host_idx, const double new_c2w7[7], float KRKi9[9], float Kt3[3]) {
  FullSystem* fs = ((RefSys*)p)->fs; FrameHessian* host = fs->frameHessians[host_idx];
  Mat33f K = Mat33f::Identity(); K(0,0) = fs->Hcalib.fxl(); K(1,1) = fs->Hcalib.fyl(); K(0,2) = fs->Hcalib.cxl(); K(1,2) = fs->Hcalib.cyl();
  SE3 hostToNew = se3_from(new_c2w7).inverse() * host->PRE_camToWorld;
  Mat33f KRKi = K * hostToNew.rotationMatrix().cast<float>() * K.inverse(); Vec3f Kt = K * hostToNew.translation().cast<float>();
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) KRKi9[3*i+j] = KRKi(i, j); Kt3[i] = Kt[i]; }
}

// ---- sequence level, back-end: the LIVE sliding window of a running system flattened into the layout of synth.make_ba_window / orc_ba_* (frames = ef->frames, points =
// ef order per host, residuals = EFPoint::residualsAll order), and FullSystem::optimize run on that live window (tests/test_sequence_ba.py)
void ref_sys_window_sizes(void* p, int sizes[3]) { FullSystem* fs = ((RefSys*)p)->fs; EnergyFunctional* ef = fs->ef; int nP = 0, nR = 0;
  for (EFFrame* f : ef->frames) for (EFPoint* q : f->points) { nP++; nR += (int)q->residualsAll.size(); }
  sizes[0] = (int)ef->frames.size(); sizes[1] = nP; sizes[2] = nR; }
void ref_sys_export_window(void* p, int* shell_ids, double* T_eval7, double* state10, double* state_zero10, float* ab_exposure, int* frameID, float* frameEnergyTH,
                           double* calib_value_scaled4, double* calib_value_zero4, float* uv, float* idepth, float* idepth_zero, float* color8, float* weights8, int* host,
                           int* hasDepthPrior, int* isFromSensor, int* res_begin, int* r_point, int* r_host, int* r_target, int* r_hasMatcher, float* r_matcher2, int* r_isNew,
                           double* HM, double* bM) {
  FullSystem* fs = ((RefSys*)p)->fs; EnergyFunctional* ef = fs->ef; int nF = (int)ef->frames.size();
  for (int i = 0; i < nF; i++) { FrameHessian* fh = ef->frames[i]->data; shell_ids[i] = fh->shell->id; se3_to(fh->worldToCam_evalPT, T_eval7 + 7*i);
    for (int k = 0; k < 10; k++) { state10[10*i+k] = fh->state[k]; state_zero10[10*i+k] = fh->state_zero[k]; }
    ab_exposure[i] = fh->ab_exposure; frameID[i] = fh->frameID; frameEnergyTH[i] = fh->frameEnergyTH; }
  for (int k = 0; k < 4; k++) { calib_value_scaled4[k] = fs->Hcalib.value_scaled[k]; calib_value_zero4[k] = fs->Hcalib.value_zero[k]; }
  int ip = 0, ir = 0; res_begin[0] = 0;
  for (int i = 0; i < nF; i++) for (EFPoint* q : ef->frames[i]->points) { PointHessian* ph = q->data;
    uv[2*ip] = ph->u; uv[2*ip+1] = ph->v; idepth[ip] = ph->idepth; idepth_zero[ip] = ph->idepth_zero; for (int k = 0; k < 8; k++) { color8[8*ip+k] = ph->color[k]; weights8[8*ip+k] = ph->weights[k]; }
    host[ip] = i; hasDepthPrior[ip] = ph->hasDepthPrior ? 1 : 0; isFromSensor[ip] = ph->isFromSensor ? 1 : 0;
    for (EFResidual* er : q->residualsAll) { PointFrameResidual* r = er->data; r_point[ir] = ip; r_host[ir] = r->host->idx; r_target[ir] = r->target->idx; r_hasMatcher[ir] = r->hasMatcher ? 1 : 0;
      r_matcher2[2*ir] = r->matcher[0]; r_matcher2[2*ir+1] = r->matcher[1]; r_isNew[ir] = er->isLinearized ? -1 : (r->isNew ? 1 : 0); ir++; }
    ip++; res_begin[ip] = ir; }
  int n = (int)ef->HM.rows(); for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) HM[(size_t)i*n+j] = ef->HM(i, j); bM[i] = ef->bM[i]; }
}
float ref_sys_optimize(void* p, int its) { return ((RefSys*)p)->fs->optimize(its); }

}  // extern "C"
