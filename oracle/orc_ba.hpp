// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).  Pinned on oracle/_ref: tests/test_ref_pin_ba.py.
//
// CPU restatement of the SDV-LOAM sliding-window back-end on a FLATTENED window (the reference's pointer graph
// FrameHessian -> PointHessian -> PointFrameResidual is host bookkeeping; every array here is in the reference's
// iteration order: frames = ef->frames, points = ef->allPoints (EnergyFunctional.cpp:761-782), residuals grouped per
// point in residualsAll order).  file:line relative to /root/reference/src:
//   FrameHessian::setStateZero / setState            FullSystem/HessianBlocks.cpp:52-82, HessianBlocks.h:141-175
//   FrameFramePrecalc::set                           FullSystem/HessianBlocks.cpp:169-195
//   FullSystem::setPrecalcValues                     FullSystem/FullSystem.cpp:1358-1368
//   EnergyFunctional::setAdjointsF / setDeltaF       OptimizationBackend/EnergyFunctional.cpp:21-71, 131-156
//   PointFrameResidual::linearize / applyRes         FullSystem/Residuals.cpp:60-224, 252-274
//   EFResidual::takeDataF                            OptimizationBackend/EnergyFunctionalStructs.cpp:15-25
//   FullSystem::linearizeAll / setNewFrameEnergyTH   FullSystem/FullSystemOptimize.cpp:23-159
//   AccumulatedTopHessianSSE::addPoint<0>, stitch    OptimizationBackend/AccumulatedTopHessian.cpp:13-112, 181-242 ; .h:63-114
//   AccumulatorApprox / AccumulatorXX / AccumulatorX OptimizationBackend/MatrixAccumulators.h:560-932, 14-66, 149-208
//   AccumulatedSCHessianSSE::addPoint, stitch        OptimizationBackend/AccumulatedSCHessian.cpp:10-135 ; .h:67-110
//   EnergyFunctional::solveSystemF / resubstitute / orthogonalize / calcLEnergyF_MT / calcMEnergyF
//                                                    OptimizationBackend/EnergyFunctional.cpp:650-759, 221-282, 615-648, 295-350, 284-293
//   FullSystem::optimize / doStepFromBackup / backupState / loadSateBackup / getNullspaces
//                                                    FullSystem/FullSystemOptimize.cpp:344-502, 165-250, 255-321, 548-588
//   FullSystem::flagPointsForRemoval (numeric part) FullSystem/FullSystem.cpp:764-797 ; EFResidual::fixLinearizationF EnergyFunctionalStructs.cpp:46-55
//   EnergyFunctional::marginalizePointsF / marginalizeFrame     OptimizationBackend/EnergyFunctional.cpp:514-576, 434-512
//   AccumulatedTopHessianSSE::addPoint<2>, stitchDouble          OptimizationBackend/AccumulatedTopHessian.cpp:13-112, 118-179
//   AccumulatedSCHessianSSE::addPoint(false), stitchDouble       OptimizationBackend/AccumulatedSCHessian.cpp:10-62, 136-195
// Scope notes: isLinearized is set only by flagPointsForRemoval (FullSystem.cpp:781) on points that marginalizePointsF removes in the
// same makeKeyFrame call, so inside optimize() the "L" accumulation is identically zero and is restated as such.  setting_solverMode =
// SOLVER_ORTHOGONALIZE_X_LATER (settings.cpp:34): LDLT path, x orthogonalised against pose+scale nullspaces for iteration>=2.
#pragma once
#include <vector>
#include "orc_tracker.hpp"

namespace orc {

static const int CPARS = 4;
static const float SCALE_F = 50.0f, SCALE_C = 50.0f, SCALE_IDEPTH = 1.0f;
enum ResState { RS_IN = 0, RS_OOB = 1, RS_OUTLIER = 2 };

struct BASettings {
  float huberTH = 6;                       // settings.cpp:101
  float outlierTHSumComponent = 50*50;     // settings.cpp:65
  float idepthFixPrior = 50*50;            // settings.cpp:21
  float initialRotPrior = 1e11f, initialTransPrior = 1e10f, initialCalibHessian = 5e9f;   // settings.cpp:23-27
  float frameEnergyTHConstWeight = 0.5f, frameEnergyTHN = 0.7f, frameEnergyTHFacMedian = 1.5f, overallEnergyTHWeight = 1; // :108-111
  int minOptIterations = 1; float thOptIterations = 1.2f;   // :56-57
  double solverModeDelta = 0.00001;        // :35
};

struct RawJ { float resF[2]; float Jpdxi[2][6]; float Jpdc[2][4]; float Jpdd[2]; };     // live part of RawResidualJacobian.h:7-36

struct BAFrame {
  SE3 worldToCam_evalPT; double state[10], state_zero[10], state_backup[10], step[10];
  float ab_exposure = 1; int frameID = 0; float frameEnergyTH = 8*8*8;   // HessianBlocks.h:211
  const Frame* img = nullptr;
  double state_scaled[10]; SE3 PRE_worldToCam, PRE_camToWorld;
  double nullspaces_pose[6][6], nullspaces_scale[6];
  double prior[6], delta[6], delta_prior[6];
};
struct BAPoint {
  float u, v, idepth, idepth_zero, idepth_scaled, idepth_zero_scaled; float color[8], weights[8];
  int host; int hasDepthPrior, isFromSensor; float step = 0, idepth_backup = 0;
  float priorF, deltaF, HdiF = 0, bdSumF = 0, Hdd_accAF = 0, bd_accAF = 0, Hcd_accAF[4] = {0,0,0,0}, Hdd_accLF = 0, bd_accLF = 0, Hcd_accLF[4] = {0,0,0,0};
  float idepth_hessian = 0, maxRelBaseline = 0; int numGoodResiduals = 0;
  int res_begin = 0, res_end = 0;
};
struct BARes {
  int point, host, target; int hasMatcher; float matcher[2];
  int state_state = RS_IN, state_NewState = RS_OUTLIER; double state_energy = 0, state_NewEnergy = 0, state_NewEnergyWithOutlier = -1;
  int isNew = 1, isActive = 0, toRemove = 0;
  RawJ J, efJ; float JpJdF[8]; float centerProjectedTo[3]; float projectedTo[8][2];
  float res_toZeroF[2] = {0,0}; int isLinearized = 0;
};
struct Precalc { Mat33f PRE_RTll, PRE_KRKiTll, PRE_RTll_0; float PRE_aff_mode[2], PRE_b0_mode; Vec3f PRE_tTll, PRE_KtTll, PRE_tTll_0; };

struct BAWindow {
  BASettings set; int w = 0, h = 0;
  std::vector<BAFrame> frames; std::vector<BAPoint> points; std::vector<BARes> res;
  double c_value[4], c_value_zero[4], c_value_scaled[4], c_step[4], c_value_backup[4], c_vmvz[4]; float c_sf[4], c_si[4];   // CalibHessian
  std::vector<double> HM, bM;                        // marginalisation prior, (CPARS+6nF)^2 / (CPARS+6nF)
  // derived
  std::vector<Precalc> precalc;                      // [host*nF + target]  (host->targetPrecalc[target])
  std::vector<double> adHost, adTarget; std::vector<float> adHostF, adTargetF, adHTdeltaF;   // [h + t*nF] blocks of 36 / 6
  float cDeltaF[4]; double cPrior[4];
  std::vector<double> lastX, lastHS, lastbS; int resInA = 0;
  // statistics of the last optimize()
  int opt_iterations = 0, opt_accepts = 0; long long linearize_calls = 0;

  int nF() const { return (int)frames.size(); }
  int dim() const { return CPARS + 6*nF(); }
  void setCalibScaled(const double vs[4]);           // CalibHessian::setValueScaled + value_zero = value (ctor, HessianBlocks.h:273-289)
  void setCalibValue(const double v[4]);             // CalibHessian::setValue
  void frameSetState(BAFrame& f, const double s[10]);
  void frameSetStateZero(BAFrame& f, const double s0[10]);
  void frameTakeData(BAFrame& f);                    // EFFrame::takeData
  void init();                                       // after filling frames/points/res: takeData, setAdjointsF, setPrecalcValues
  void setAdjointsF();
  void setPrecalcValues();
  double linearizeOne(BARes& r);
  void applyRes(BARes& r);
  double linearizeAll(bool fixLinearization);
  void setNewFrameEnergyTH();
  double calcLEnergy(); double calcMEnergy();
  void solveSystem(int iteration, double lambda);
  void accumulateA(std::vector<double>& H, std::vector<double>& b);
  void accumulateSC(std::vector<double>& H, std::vector<double>& b);
  void orthogonalize(std::vector<double>& x);
  bool doStepFromBackup(float stepfac);
  void backupState(); void loadStateBackup();
  float optimize(int mnumOptIts);
  // ---- keyframe hand-over (FullSystem::makeKeyFrame, FullSystem.cpp:1152-1171)
  void flagPointsForRemoval(const int* selected, int* status);   // numeric part: re-linearise + fixLinearizationF; status 0 keep / 1 PS_DROP / 2 PS_MARGINALIZE
  void marginalizePointsF(const int* status);                    // HM,bM += 0.25 * (M - Msc) over the PS_MARGINALIZE points
  void marginalizeFrame(int idx);                                // Schur-eliminate frame idx from HM,bM and drop it from `frames`
  std::vector<double> margM, margMb, margMsc, margMbsc;          // last marginalizePointsF intermediates (tests)
};

} // namespace orc
