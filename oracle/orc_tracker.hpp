// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).  Pinned bit for bit on oracle/_ref: tests/test_ref_pin.py.
//
// CPU restatement of the SDV-LOAM front-end tracker path (all file:line relative to /root/reference/src):
//   FrameHessian::makeImages            FullSystem/HessianBlocks.cpp:107-167
//   CoarseTracker::makeK                FullSystem/CoarseTracker.cpp:77-106
//   CoarseTracker::makeCoarseDepthL0    FullSystem/CoarseTracker.cpp:258-425 (+ first-frame variant :108-256)
//   CoarseTracker::calcRes              FullSystem/CoarseTracker.cpp:486-634
//   CoarseTracker::calcGSSSE            FullSystem/CoarseTracker.cpp:427-484
//   Accumulator9                        OptimizationBackend/MatrixAccumulators.h:934-1293
//   CoarseTracker::trackNewestCoarse    FullSystem/CoarseTracker.cpp:662-838
//   getInterpolatedElement33            util/globalFuncs.h:51-65
// Compiled with g++ -O3 -ffp-contract=off (mirrors CMakeLists.txt:4: -O3, SSE2 baseline, no FMA).
#pragma once
#include <vector>
#include <cstdint>
#include "orc_math.hpp"

namespace orc {

static const int PYR_LEVELS = 6;                    // util/settings.h:25
static const float SCALE_XI_ROT = 1.0f, SCALE_XI_TRANS = 0.5f, SCALE_A = 10.0f, SCALE_B = 1000.0f; // HessianBlocks.h:33-40

struct Settings {                                   // util/settings.cpp (mode=1 of launch/run.launch:12, main.cpp:450-455)
  float huberTH = 6;                                // settings.cpp:101
  float coarseCutoffTH = 20;                        // settings.cpp:112
  float affineOptModeA = 0, affineOptModeB = 0;     // main.cpp:453-454
};

// FrameHessian image part: dIp[lvl] = AoS Vector3f {I,dx,dy}; absSquaredGrad[lvl]
struct Frame {
  int levels = 0, w[PYR_LEVELS] = {0}, h[PYR_LEVELS] = {0};
  std::vector<float> dIp[PYR_LEVELS];               // 3 floats / pixel
  std::vector<float> absSquaredGrad[PYR_LEVELS];
  float ab_exposure = 1.0f;
  void makeImages(const float* color, int w0, int h0, int levels_);
};

// pyrLevelsUsed rule of util/globalCalib.cpp:22-30
int pyrLevelsUsedFor(int w, int h);

struct RefPoint { float u, v, idepth, HdiF; int round_half; };   // round_half: 0 => int(u) (CoarseTracker.cpp:270-271), 1 => int(u+0.5f) (:116-117,:285-286)

struct CoarseTracker {
  Settings set;
  int levels = 0, w[PYR_LEVELS] = {0}, h[PYR_LEVELS] = {0};
  float fx[PYR_LEVELS], fy[PYR_LEVELS], cx[PYR_LEVELS], cy[PYR_LEVELS];
  Mat33f K[PYR_LEVELS], Ki[PYR_LEVELS];
  std::vector<float> idepth[PYR_LEVELS], weightSums[PYR_LEVELS], weightSums_bak[PYR_LEVELS];
  std::vector<float> pc_u[PYR_LEVELS], pc_v[PYR_LEVELS], pc_idepth[PYR_LEVELS], pc_color[PYR_LEVELS];
  int pc_n[PYR_LEVELS] = {0};
  std::vector<float> buf_warped_idepth, buf_warped_u, buf_warped_v, buf_warped_dx, buf_warped_dy,
                     buf_warped_residual, buf_warped_weight, buf_warped_refColor;
  int buf_warped_n = 0;
  const Frame* lastRef = nullptr; const Frame* newFrame = nullptr;
  AffLight lastRef_aff_g2l;
  double lastResiduals[5]; double lastFlowIndicators[3]; double firstCoarseRMSE = -1;
  // statistics (not in the reference): number of calcRes point evaluations per level of the last track call
  long long evals[PYR_LEVELS] = {0}; int iterations[PYR_LEVELS] = {0}; int accepts[PYR_LEVELS] = {0};

  void init(int ww, int hh, int levels_);
  void makeK(float fxl, float fyl, float cxl, float cyl);
  void setCoarseTrackingRef(const Frame* ref, const RefPoint* pts, int n, AffLight ref_aff);  // :649-660 on a flattened point list
  void setRefCloud(const Frame* ref, int lvl, int n, const float* u, const float* v, const float* idepth, const float* color); // direct pc_* injection (tests)
  void calcRes(int lvl, const SE3& refToNew, AffLight aff_g2l, float cutoffTH, double rs[6]);
  void calcGSSSE(int lvl, double H_out[64], double b_out[8], const SE3& refToNew, AffLight aff_g2l);
  bool trackNewestCoarse(const Frame* newFrameHessian, SE3& lastToNew_out, AffLight& aff_g2l_out, int coarsestLvl, const double minResForAbort[5]);
};

// Accumulator9 with the SSE lane layout and 3-tier (1 / 1k / 1M) float sums of MatrixAccumulators.h:934-1293
struct Accumulator9 {
  float H[9][9]; size_t num;
  float SSEData[4*45], SSEData1k[4*45], SSEData1m[4*45]; float numIn1, numIn1k, numIn1m;
  void initialize();
  void updateSSE_eighted(const float J[9][4], const float w[4]);
  void finish();
  void shiftUp(bool force);
};

} // namespace orc
