// integration/shim_check.cpp — the shims of INTEGRATION.md as a COMPILED translation unit.
//
// The reference has no plugin layer: libsdv_b200.so is dropped in by swapping the bodies of a few member functions.  The reference's own
// headers cannot be compiled in this image (Eigen3/Boost/ROS absent), so this file declares minimal stand-ins with the SAME member names and types the shims
// touch (FrameShell::id, PointHessian::{u,v,idepth,...}, SE3, AffLight, Vec5, ...) and then contains the shim bodies verbatim.  tests/test_abi.py compiles it
// (g++ -std=c++14 -Wall -Werror -fsyntax-only) and links it against libsdv_b200.so with --no-undefined: every call in INTEGRATION.md is type-checked against
// include/sdv_b200.h and resolves to an exported C symbol.  It is never executed.
#include <cstddef>
#include <cstdint>
#include <utility>
#include <vector>
#include <cstdlib>
#include "sdv_b200.h"

// ------------------------------------------------------------------------------------------------ stand-ins for the reference's types (names as in the reference)
struct Quat { double w_, x_, y_, z_; double w() const { return w_; } double x() const { return x_; } double y() const { return y_; } double z() const { return z_; } };
struct Vec3 { double v[3]; double operator[](int i) const { return v[i]; } };
struct SE3 {                                                                   // Sophus::SE3d: unit_quaternion(), translation()
  Quat q; Vec3 t;
  SE3() : q{1, 0, 0, 0}, t{{0, 0, 0}} {}
  SE3(const Quat& q_, const Vec3& t_) : q(q_), t(t_) {}
  const Quat& unit_quaternion() const { return q; } const Vec3& translation() const { return t; }
};
struct AffLight { double a, b; AffLight(double a_ = 0, double b_ = 0) : a(a_), b(b_) {} };   // util/NumType.h:129-175
struct Vec5 { double v[5]; const double* data() const { return v; } double& operator[](int i) { return v[i]; } };
struct Vec4 { double v[4]; };
struct Vec10 { double v[10]; const double* data() const { return v; } };
struct VecC { double v[4]; const double* data() const { return v; } };
enum class ResState { IN = 0, OOB, OUTLIER };
struct FrameHessian; struct PointHessian;
struct FrameShell { int id; SE3 camToWorld, camToTrackingRef; AffLight aff_g2l; bool poseValid; FrameShell* trackingRef; };   // util/FrameShell.h
struct EFPoint { float HdiF; };
struct Vec3f { float v[3]; float operator[](int i) const { return v[i]; } };
struct PointFrameResidual { Vec3f centerProjectedTo; ResState state_state; FrameHessian* host; FrameHessian* target; bool hasMatcher; float matcher[2]; bool isNew; };
struct PointHessian {                                                          // FullSystem/HessianBlocks.h:361-465
  float u, v, idepth, idepth_zero, color[8], weights[8]; bool hasDepthPrior, isFromSensor; int type; enum { EDGELET = 1 };
  FrameHessian* host; EFPoint* efPoint; std::vector<PointFrameResidual*> residuals; std::pair<PointFrameResidual*, ResState> lastResiduals[2];
};
struct FrameHessian {                                                          // FullSystem/HessianBlocks.h:104-258
  FrameShell* shell; float ab_exposure; int idx, frameID; float frameEnergyTH; bool flaggedForMarginalization;
  std::vector<PointHessian*> pointHessians; SE3 worldToCam_evalPT; Vec10 state, state_zero;
  AffLight aff_g2l() const { return shell->aff_g2l; }
  void makeImages(float* color, struct CalibHessian* HCalib);
};
struct CalibHessian { VecC value_scaled, value_zero; float fxl() const { return (float)value_scaled.v[0]; } float fyl() const { return (float)value_scaled.v[1]; }
                      float cxl() const { return (float)value_scaled.v[2]; } float cyl() const { return (float)value_scaled.v[3]; } };
struct Undistort { int wOrg, hOrg; float* remapX; float* remapY; };           // util/Undistort.h (remapX/remapY: protected there — one friend declaration or two getters)
namespace IOWrap { struct Output3DWrapper; }
static sdv_ctx* gpu = nullptr;                                                 // FullSystem member in the real integration
extern float setting_huberTH, setting_coarseCutoffTH, setting_affineOptModeA, setting_affineOptModeB, setting_outlierTHSumComponent, setting_idepthFixPrior, setting_desiredImmatureDensity;
extern int setting_maxFrames, pyrLevelsUsed, wG[8], hG[8];

static void se3_to7(const SE3& T, double o[7]) {
  o[0] = T.unit_quaternion().w(); o[1] = T.unit_quaternion().x(); o[2] = T.unit_quaternion().y(); o[3] = T.unit_quaternion().z();
  o[4] = T.translation()[0]; o[5] = T.translation()[1]; o[6] = T.translation()[2];
}
static SE3 se3_from7(const double* v) { return SE3(Quat{v[0], v[1], v[2], v[3]}, Vec3{{v[4], v[5], v[6]}}); }

// ------------------------------------------------------------------------------------------------ §2 one context per FullSystem (FullSystem.cpp:119)
int shim_create(CalibHessian& Hcalib) {
  sdv_settings s; sdv_default_settings(&s);
  s.huberTH = setting_huberTH; s.coarseCutoffTH = setting_coarseCutoffTH; s.affineOptModeA = setting_affineOptModeA; s.affineOptModeB = setting_affineOptModeB;
  s.outlierTHSumComponent = setting_outlierTHSumComponent; s.idepthFixPrior = setting_idepthFixPrior;
  s.n_tracker_slots = 2; s.max_frames = setting_maxFrames + 3; s.cluster_size = 8; s.track_threads = 256;
  sdv_calib K{Hcalib.fxl(), Hcalib.fyl(), Hcalib.cxl(), Hcalib.cyl()};
  return sdv_create(&K, wG[0], hG[0], pyrLevelsUsed, &s, /*device*/0, &gpu);
}

// ------------------------------------------------------------------------------------------------ §3 FrameHessian::makeImages (HessianBlocks.cpp:107-167)
void FrameHessian::makeImages(float* color, CalibHessian*) { sdv_frame_upload(gpu, (uint64_t)shell->id, color, ab_exposure); }
void shim_frame_destructor(FrameHessian* fh) { sdv_frame_release(gpu, (uint64_t)fh->shell->id); }

// §3b ingest: Undistort tables once, raw mono8 per frame (util/DatasetReader.h:223-230, Undistort.cpp:341-435)
int shim_set_undistort(const Undistort* u, const float* G256_or_null, const float* vignetteInv_or_null) {
  return sdv_set_undistort(gpu, u->wOrg, u->hOrg, u->remapX, u->remapY, 1.0f, G256_or_null, vignetteInv_or_null);
}
int shim_ingest_raw(FrameShell* shell, const uint8_t* raw_mono8, float exposure) {
  uint64_t id = (uint64_t)shell->id; return sdv_frame_upload_batch_raw_u8(gpu, 1, &id, &raw_mono8, &exposure);
}
// §3c CoarseTracker::makeK (CoarseTracker.cpp:77-106)
int shim_makeK(CalibHessian* HCalib) { sdv_calib K{HCalib->fxl(), HCalib->fyl(), HCalib->cxl(), HCalib->cyl()}; return sdv_set_calib(gpu, &K); }

// ------------------------------------------------------------------------------------------------ CoarseTracker (CoarseTracker.cpp:649-660, 662-838)
struct CoarseTracker {
  int slot; FrameHessian* lastRef; int refFrameID; AffLight lastRef_aff_g2l; double firstCoarseRMSE; Vec5 lastResiduals; double lastFlowIndicators[3];
  void setCoarseTrackingRef(std::vector<FrameHessian*> frameHessians) {
    lastRef = frameHessians.back();
    std::vector<float> pts; std::vector<int32_t> round_half;                  // the graph walk of makeCoarseDepthL0 :264-294 stays on the host
    for (FrameHessian* fh : frameHessians) for (PointHessian* ph : fh->pointHessians) {
      if (fh == frameHessians.back() && ph->isFromSensor) {
        pts.insert(pts.end(), {ph->u, ph->v, ph->idepth, ph->efPoint->HdiF}); round_half.push_back(0);
      } else if (ph->lastResiduals[0].first && ph->lastResiduals[0].second == ResState::IN && fh != frameHessians.back() && ph->isFromSensor) {
        PointFrameResidual* r = ph->lastResiduals[0].first;
        pts.insert(pts.end(), {r->centerProjectedTo[0], r->centerProjectedTo[1], r->centerProjectedTo[2], ph->efPoint->HdiF}); round_half.push_back(1);
      }
    }
    sdv_tracker_set_ref(gpu, slot, (uint64_t)lastRef->shell->id, (int)round_half.size(), pts.data(), round_half.data(), 0.f, lastRef->aff_g2l().a, lastRef->aff_g2l().b);
    refFrameID = lastRef->shell->id; lastRef_aff_g2l = lastRef->aff_g2l(); firstCoarseRMSE = -1;
  }
  bool trackNewestCoarse(FrameHessian* fh, SE3& lastToNew_out, AffLight& aff_g2l_out, int coarsestLvl, Vec5 minResForAbort, IOWrap::Output3DWrapper*) {
    double T[7]; se3_to7(lastToNew_out, T);
    double ab[2] = {aff_g2l_out.a, aff_g2l_out.b}, res[5], flow[3]; int good = 0;
    sdv_tracker_track(gpu, slot, (uint64_t)fh->shell->id, T, ab, coarsestLvl, minResForAbort.data(), res, flow, &good, nullptr);
    lastToNew_out = se3_from7(T); aff_g2l_out = AffLight(ab[0], ab[1]);
    for (int i = 0; i < 5; i++) lastResiduals[i] = res[i];
    for (int i = 0; i < 3; i++) lastFlowIndicators[i] = flow[i];
    return good != 0;
  }
};

// ------------------------------------------------------------------------------------------------ FullSystem members
struct FullSystem {
  std::vector<FrameHessian*> frameHessians; std::vector<FrameShell*> allFrameHistory; CalibHessian Hcalib; CoarseTracker* coarseTracker; Vec5 lastCoarseRMSE;
  std::vector<double> HM, bM;                                                   // ef->HM / ef->bM, row-major copies
  std::vector<int32_t> cell_order;                                              // Reprojector grid_.cell_order (Reprojector.cpp:107)

  // FullSystem::optimize (FullSystemOptimize.cpp:344-502): flatten in ef->frames / allPoints / residualsAll order (EnergyFunctional::makeIDX, EnergyFunctional.cpp:761-782)
  float optimize(int mnumOptIts) {
    const int nF = (int)frameHessians.size();
    std::vector<uint64_t> ids(nF); std::vector<double> T_eval(7*nF), state(10*nF), state_zero(10*nF); std::vector<float> exposure(nF), energyTH(nF); std::vector<int32_t> frameID(nF);
    for (int i = 0; i < nF; i++) { FrameHessian* fh = frameHessians[i];
      ids[i] = (uint64_t)fh->shell->id; se3_to7(fh->worldToCam_evalPT, &T_eval[7*i]); exposure[i] = fh->ab_exposure; frameID[i] = fh->frameID; energyTH[i] = fh->frameEnergyTH;
      for (int k = 0; k < 10; k++) { state[10*i+k] = fh->state.data()[k]; state_zero[10*i+k] = fh->state_zero.data()[k]; } }
    std::vector<float> uv, idepth, idepth_zero, color, weights, r_matcher; std::vector<int32_t> host, hasPrior, fromSensor, res_begin{0}, r_point, r_host, r_target, r_hasM, r_isNew;
    for (FrameHessian* fh : frameHessians) for (PointHessian* ph : fh->pointHessians) {
      const int p = (int)host.size();
      uv.push_back(ph->u); uv.push_back(ph->v); idepth.push_back(ph->idepth); idepth_zero.push_back(ph->idepth_zero);
      color.insert(color.end(), ph->color, ph->color + 8); weights.insert(weights.end(), ph->weights, ph->weights + 8);
      host.push_back(fh->idx); hasPrior.push_back(ph->hasDepthPrior); fromSensor.push_back(ph->isFromSensor);
      for (PointFrameResidual* r : ph->residuals) { r_point.push_back(p); r_host.push_back(r->host->idx); r_target.push_back(r->target->idx); r_hasM.push_back(r->hasMatcher);
        r_matcher.push_back(r->matcher[0]); r_matcher.push_back(r->matcher[1]); r_isNew.push_back(r->isNew); }
      res_begin.push_back((int32_t)r_point.size());
    }
    int rc = sdv_ba_set_window(gpu, nF, ids.data(), T_eval.data(), state.data(), state_zero.data(), exposure.data(), frameID.data(), energyTH.data(), Hcalib.value_scaled.data(), HM.data(), bM.data());
    if (rc == SDV_OK) rc = sdv_ba_set_calib_zero(gpu, Hcalib.value_zero.data());                   // the linearisation point of the intrinsics stays at the initial K (HessianBlocks.h:287)
    if (rc == SDV_OK) rc = sdv_ba_set_points(gpu, (int)host.size(), uv.data(), idepth.data(), idepth_zero.data(), color.data(), weights.data(), host.data(), hasPrior.data(), fromSensor.data(),
                                             res_begin.data(), (int)r_point.size(), r_point.data(), r_host.data(), r_target.data(), r_hasM.data(), r_matcher.data(), r_isNew.data());
    float rmse = 0; int32_t its = 0, acc = 0;
    if (rc == SDV_OK) rc = sdv_ba_optimize(gpu, mnumOptIts, &rmse, &its, &acc);
    std::vector<double> Tn(7*nF), st(10*nF), step(10*nF), pre(7*nF); std::vector<float> th(nF); double cv[4], cs[4];
    if (rc == SDV_OK) sdv_ba_get_frames(gpu, Tn.data(), st.data(), step.data(), th.data(), pre.data(), cv, cs);        // -> FrameHessian::setEvalPT / setState, Hcalib.setValue
    std::vector<float> idn(host.size()), stp(host.size()), hdi(host.size()), bds(host.size()), mrb(host.size()), idh(host.size()); std::vector<int32_t> ngood(host.size());
    if (rc == SDV_OK) sdv_ba_get_points(gpu, idn.data(), stp.data(), hdi.data(), bds.data(), mrb.data(), ngood.data(), idh.data());   // -> PointHessian::setIdepth, idepth_hessian, maxRelBaseline, numGoodResiduals
    return rmse;
  }

  // FullSystem::makeKeyFrame: keep the map resident for the Reprojector (Reprojector.cpp:117-156) ...
  int set_map(int slot) {
    std::vector<sdv_map_pt> pts; std::vector<uint64_t> ids; std::vector<double> c2w, ab;
    for (FrameHessian* fh : frameHessians) {
      ids.push_back((uint64_t)fh->shell->id); double T[7]; se3_to7(fh->shell->camToWorld, T); c2w.insert(c2w.end(), T, T + 7); ab.push_back(fh->aff_g2l().a); ab.push_back(fh->aff_g2l().b);
      for (PointHessian* ph : fh->pointHessians) pts.push_back(sdv_map_pt{ph->u, ph->v, ph->idepth, fh->idx, ph->type == PointHessian::EDGELET});
    }
    return sdv_map_set(gpu, slot, (int)ids.size(), ids.data(), c2w.data(), ab.data(), (int)pts.size(), pts.data());
  }
  // ... and the hand-over after optimize(): flagPointsForRemoval / marginalizePointsF / marginalizeFrame (FullSystem.cpp:1152-1171)
  int handover(const std::vector<int32_t>& selected, std::vector<int32_t>& status) {
    status.resize(selected.size());
    int rc = sdv_ba_flag_points(gpu, selected.data(), status.data());
    if (rc == SDV_OK) rc = sdv_ba_marginalize_points(gpu, nullptr);
    for (FrameHessian* fh : frameHessians) if (rc == SDV_OK && fh->flaggedForMarginalization) rc = sdv_ba_marginalize_frame(gpu, fh->idx);
    int dim = 0; HM.resize(68*68); bM.resize(68); if (rc == SDV_OK) rc = sdv_ba_get_prior(gpu, &dim, HM.data(), bM.data());
    return rc;
  }

  // Vec4 FullSystem::trackNewCoarse(FrameHessian* fh) for a running system (FullSystem.cpp:283-500)
  Vec4 trackNewCoarse(FrameHessian* fh) {
    FrameShell* slast = allFrameHistory[allFrameHistory.size() - 2]; FrameShell* sprelast = allFrameHistory[allFrameHistory.size() - 3]; FrameHessian* lastF = coarseTracker->lastRef;
    sdv_track_new_coarse_io io{}; io.slot = coarseTracker->slot; io.frame = (uint64_t)fh->shell->id;
    io.poses_valid = slast->poseValid && sprelast->poseValid && lastF->shell->poseValid;
    se3_to7(sprelast->camToWorld, io.sprelast_c2w); se3_to7(slast->camToWorld, io.slast_c2w); se3_to7(lastF->shell->camToWorld, io.lastF_c2w);
    io.aff_last[0] = slast->aff_g2l.a; io.aff_last[1] = slast->aff_g2l.b;
    for (int i = 0; i < 5; i++) io.lastCoarseRMSE[i] = lastCoarseRMSE[i];
    sdv_track_new_coarse_batch(gpu, 1, &io, cell_order.data(), (int)(0.8f*setting_desiredImmatureDensity));
    fh->shell->camToWorld = se3_from7(io.camToWorld); fh->shell->camToTrackingRef = se3_from7(io.camToTrackingRef); fh->shell->trackingRef = lastF->shell;
    fh->shell->aff_g2l = AffLight(io.aff_g2l[0], io.aff_g2l[1]);
    for (int i = 0; i < 5; i++) lastCoarseRMSE[i] = io.lastCoarseRMSE[i];
    return Vec4{{io.lastCoarseRMSE[0], io.flow[0], io.flow[1], io.flow[2]}};
  }
  // the two members of the tail separately (overlap_pts wanted by the caller): Reprojector::reprojectMap + CoarseTracker::structPoseEstimation
  int refine_only(FrameHessian* fh) {
    double T[7]; se3_to7(fh->shell->camToWorld, T); double ab[2] = {fh->aff_g2l().a, fh->aff_g2l().b}; int32_t s = coarseTracker->slot, nm = 0; uint64_t id = (uint64_t)fh->shell->id;
    int rc = sdv_tracker_refine_batch(gpu, 1, &s, &id, T, ab, cell_order.data(), (int)(0.8f*setting_desiredImmatureDensity), &nm, nullptr, nullptr, nullptr);
    fh->shell->camToWorld = se3_from7(T); return rc;
  }
};

// ------------------------------------------------------------------------------------------------ immature points: FullSystem::traceNewCoarse (FullSystem.cpp:519-552)
struct Mat33f { float m[9]; }; struct Vec2f { float v[2]; };
struct ImmaturePoint { sdv_immature_pt rec; FrameHessian* host; };            // the reference's members u,v,idepth_min/max,color,weights,gradH,energyTH,quality,lastTrace* in one record
struct HostImmatures { FrameHessian* host; std::vector<ImmaturePoint*> immaturePoints; Mat33f KRKi; Vec3f Kt; Vec2f aff; };   // per host: what :532-538 computes
int shim_makeNewTraces(FrameHessian* newFrame, const std::vector<int32_t>& uv /*2 per selected pixel (PixelSelector)*/, std::vector<ImmaturePoint*>& out) {
  std::vector<sdv_immature_pt> rec(uv.size()/2);
  int rc = sdv_immature_init(gpu, (uint64_t)newFrame->shell->id, (int)rec.size(), uv.data(), rec.data());   // ImmaturePoint::ImmaturePoint for every candidate (FullSystem.cpp:1273-1356)
  for (std::size_t i = 0; i < rec.size() && rc == SDV_OK; i++) out.push_back(new ImmaturePoint{rec[i], newFrame});
  return rc;
}
int shim_traceNewCoarse(FrameHessian* fh, std::vector<HostImmatures>& hosts) {
  std::vector<uint64_t> frames; std::vector<int32_t> pt_begin{0}; std::vector<float> KRKi, Kt, aff; std::vector<sdv_immature_pt> pts;
  for (HostImmatures& h : hosts) {
    frames.push_back((uint64_t)fh->shell->id); KRKi.insert(KRKi.end(), h.KRKi.m, h.KRKi.m + 9); for (int i = 0; i < 3; i++) Kt.push_back((float)h.Kt[i]); aff.push_back(h.aff.v[0]); aff.push_back(h.aff.v[1]);
    for (ImmaturePoint* ph : h.immaturePoints) pts.push_back(ph->rec);
    pt_begin.push_back((int32_t)pts.size());
  }
  int rc = sdv_immature_trace_batch(gpu, (int)hosts.size(), frames.data(), pt_begin.data(), KRKi.data(), Kt.data(), aff.data(), pts.data(), nullptr);
  std::size_t k = 0; for (HostImmatures& h : hosts) for (ImmaturePoint* ph : h.immaturePoints) ph->rec = pts[k++];       // lastTraceStatus etc. back into the graph
  return rc;
}

// FullSystem::activatePointsMT -> optimizeImmaturePoint (FullSystem.cpp:569-723, FullSystemOptPoint.cpp:18-183): all candidates picked for activation, grouped by host
struct ActivationGroup { FrameHessian* host; std::vector<ImmaturePoint*> toOptimize; std::vector<uint8_t> isFromSensor; std::vector<FrameHessian*> targets; std::vector<float> pre14; float calib6[6]; };
int shim_activatePoints(std::vector<ActivationGroup>& groups, std::vector<int32_t>& status, std::vector<float>& idepth, std::vector<int32_t>& res_state, int res_stride) {
  std::vector<int32_t> pt_begin{0}, tgt_begin{0}; std::vector<uint64_t> tf; std::vector<float> pre, cal; std::vector<sdv_immature_pt> pts; std::vector<uint8_t> fs;
  for (ActivationGroup& g : groups) {
    for (ImmaturePoint* ph : g.toOptimize) pts.push_back(ph->rec);
    fs.insert(fs.end(), g.isFromSensor.begin(), g.isFromSensor.end());
    for (FrameHessian* t : g.targets) tf.push_back((uint64_t)t->shell->id);
    pre.insert(pre.end(), g.pre14.begin(), g.pre14.end()); cal.insert(cal.end(), g.calib6, g.calib6 + 6);
    pt_begin.push_back((int32_t)pts.size()); tgt_begin.push_back((int32_t)tf.size());
  }
  status.resize(pts.size()); idepth.resize(pts.size()); res_state.resize(pts.size()*(std::size_t)res_stride);
  return sdv_immature_optimize_batch(gpu, (int)groups.size(), pt_begin.data(), tgt_begin.data(), tf.data(), pre.data(), cal.data(), /*minObs*/1, pts.data(), fs.data(), res_stride,
                                     status.data(), idepth.data(), res_state.data());     // status 1 -> new PointHessian(point) + PointFrameResidual per IN state, -1 -> delete, 0 -> keep immature
}

// ------------------------------------------------------------------------------------------------ §3e keyframe-rate candidate management
// PixelSelector::PixelSelector (PixelSelector2.cpp:11-26): the selector's random pattern is generated where the reference generates it and handed to the library once
struct PixelSelector { unsigned char* randomPattern; int currentPotential; };
int shim_selector_ctor(PixelSelector& ps) {
  ps.randomPattern = new unsigned char[(std::size_t)wG[0]*hG[0]]; std::srand(3141592); for (int i = 0; i < wG[0]*hG[0]; i++) ps.randomPattern[i] = std::rand() & 0xFF;   // :14-16
  ps.currentPotential = 3; return sdv_selector_init(gpu, ps.randomPattern, /*slots: one per FullSystem*/1);
}
// void lidarCloudHandler(const sensor_msgs::PointCloud2ConstPtr&)  main.cpp:785-858: the decoded XYZI rows of one sweep in, vCloudPixel + addFeaturePoint out
struct LidarRig { double Rlc[9], tlc[3]; float fx, fy, cx, cy; int left, right, up, down; bool addFeaturePoint; };           // the FullSystem members main.cpp:368-377, :834-854 touches
int shim_lidarCloudHandler(LidarRig& fsys, const float* xyzi, int n, std::vector<double>& vCloudPixel3) {
  static bool once = false; if (!once) { int rc = sdv_lidar_init(gpu, /*N_SCAN*/64, /*Horizon_SCAN*/1800, 0.2f, 0.427f, 24.9f, /*groundScanInd*/50); if (rc) return rc; once = true; }   // main.cpp:103-108
  const int32_t sweep_begin[2] = {0, n}; const float K4[4] = {fsys.fx, fsys.fy, fsys.cx, fsys.cy}; int32_t lrud[4] = {fsys.left, fsys.right, fsys.up, fsys.down}, n_out = 0, add = 0;
  const int cap = 64*1800; vCloudPixel3.resize(3*(std::size_t)cap);
  int rc = sdv_lidar_handler_batch(gpu, 1, sweep_begin, xyzi, fsys.Rlc, fsys.tlc, K4, lrud, cap, vCloudPixel3.data(), &n_out, &add, nullptr);
  fsys.left = lrud[0]; fsys.right = lrud[1]; fsys.up = lrud[2]; fsys.down = lrud[3]; fsys.addFeaturePoint = add != 0; vCloudPixel3.resize(3*(std::size_t)n_out);           // -> fullSystem->qCloudPixel.push(vCloudPixel)
  return rc;
}
// void FullSystem::makeNewTraces(FrameHessian* newFrame, float*)  FullSystem.cpp:1273-1356, whole: selection, Shi-Tomasi typing, occupancy mask, ImmaturePoint construction
int shim_makeNewTraces_whole(FrameHessian* newFrame, const LidarRig& fsys, const std::vector<double>& vCloudPixel3, std::vector<sdv_new_trace>& traces, std::vector<sdv_immature_pt>& recs) {
  const int32_t slot = 0, cloud_begin[2] = {0, (int32_t)(vCloudPixel3.size()/3)}, add = fsys.addFeaturePoint ? 1 : 0; const uint64_t id = (uint64_t)newFrame->shell->id;
  const int lidarArea = (fsys.right - fsys.left)*(fsys.down - fsys.up), imageArea = wG[0]*hG[0];
  const float dl = ((float)lidarArea/(float)imageArea) * setting_desiredImmatureDensity, dd = setting_desiredImmatureDensity;                                                   // :1287-1293
  const int cap = 1 << 14; traces.resize(cap); recs.resize(cap); int32_t n_out = 0, num[2];
  int rc = sdv_make_new_traces_batch(gpu, 1, &slot, &id, cloud_begin, vCloudPixel3.data(), &dl, &dd, &add, cap, traces.data(), recs.data(), &n_out, num);
  traces.resize(n_out); recs.resize(n_out);   // per row: new ImmaturePoint{rec}, my_type, score, type, isFromSensor, idepth_fromSensor (:1307-1325, :1343-1352) -> newFrame->immaturePoints
  return rc;
}
// the selection half of void FullSystem::activatePointsMT()  FullSystem.cpp:600-671 (makeDistanceMap + candidate walk); the survivors go to shim_activatePoints
struct ActivationInputs { std::vector<int32_t> pt_begin, cand_begin; std::vector<float> KRKi9, Kt3, uvid, cKRKi9, cKt3, cand4; float currentMinActDist; };   // per keyframe: K[1] R K[0]^-1, K[1] t (:606-608)
int shim_activateSelect(const ActivationInputs& in, std::vector<int32_t>& decision) {
  const int32_t host_begin[2] = {0, (int32_t)in.pt_begin.size() - 1}, cand_host_begin[2] = {0, (int32_t)in.cand_begin.size() - 1};
  decision.resize(in.cand_begin.back());
  return sdv_activate_select_batch(gpu, 1, host_begin, in.pt_begin.data(), in.KRKi9.data(), in.Kt3.data(), in.uvid.data(), cand_host_begin, in.cand_begin.data(), in.cKRKi9.data(), in.cKt3.data(),
                                   in.cand4.data(), &in.currentMinActDist, decision.data(), nullptr);   // 1 -> toOptimize.push_back(ph), 0 -> skipped, -1 -> delete ph (:653-669)
}

// referenced so that -Wunused does not hide a missing call path
int shim_check_anchor(FullSystem& fs, FrameHessian* fh, CalibHessian& hc, const Undistort* u, const uint8_t* raw) {
  std::vector<int32_t> sel, st; Vec5 mr{}; SE3 T; AffLight a;
  int rc = shim_create(hc) | shim_set_undistort(u, nullptr, nullptr) | shim_ingest_raw(fh->shell, raw, 1.0f) | shim_makeK(&hc);
  fs.coarseTracker->setCoarseTrackingRef(fs.frameHessians); fs.coarseTracker->trackNewestCoarse(fh, T, a, 3, mr, nullptr);
  std::vector<ImmaturePoint*> imm; std::vector<HostImmatures> hi; rc |= shim_makeNewTraces(fh, sel, imm) | shim_traceNewCoarse(fh, hi);
  std::vector<ActivationGroup> ag; std::vector<float> idp; rc |= shim_activatePoints(ag, st, idp, sel, 7);
  PixelSelector ps; LidarRig rig{}; std::vector<double> px; std::vector<sdv_new_trace> tr; std::vector<sdv_immature_pt> rc2; ActivationInputs ai; ai.pt_begin = {0}; ai.cand_begin = {0}; ai.currentMinActDist = 2;
  rc |= shim_selector_ctor(ps) | shim_lidarCloudHandler(rig, nullptr, 0, px) | shim_makeNewTraces_whole(fh, rig, px, tr, rc2) | shim_activateSelect(ai, st);
  fs.optimize(6); fs.set_map(0); fs.handover(sel, st); fs.trackNewCoarse(fh); fs.refine_only(fh); shim_frame_destructor(fh);
  return rc;
}
