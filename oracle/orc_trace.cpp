// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).  Pinned bit for bit on oracle/_ref: tests/test_ref_pin_trace.py.
//
// orc_trace.cpp — restatement of the immature-point path:
//   ImmaturePoint::ImmaturePoint   FullSystem/ImmaturePoint.cpp:8-36     pattern colours, gradient matrix gradH, weights, energyTH of a new candidate
//   ImmaturePoint::traceOn         FullSystem/ImmaturePoint.cpp:50-352   epipolar-line search of the candidate in a new frame: discrete search over <= 99
//                                                                        steps, Gauss-Newton refinement along the line, interval update, status machine
//   called per host keyframe by FullSystem::traceNewCoarse (FullSystem.cpp:519-552) with KRKi = K R K^-1, Kt = K t, affine = fromToVecExposure.
// float arithmetic in the reference's operation order (-ffp-contract=off); Eigen expressions expanded in index order (the order of oracle/ref_stub's products).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include "orc_tracker.hpp"

namespace orc {

static const int kPat[8][2] = {{0,-2},{-1,-1},{1,-1},{-2,0},{0,0},{2,0},{-1,1},{0,2}};   // staticPattern[8] (util/settings.cpp:250), patternNum = 8
enum { IPS_GOOD = 0, IPS_OOB, IPS_OUTLIER, IPS_SKIPPED, IPS_BADCONDITION, IPS_UNINITIALIZED };   // ImmaturePoint.h:24-30

struct TraceSettings {                               // util/settings.cpp:64,101,111,130-139
  float outlierTH = 12*12, outlierTHSumComponent = 50*50, overallEnergyTHWeight = 1, huberTH = 6, maxPixSearch = 0.027f;
  int minTraceTestRadius = 2; float trace_stepsize = 1.0f; int trace_GNIterations = 3; float trace_GNThreshold = 0.1f, trace_extraSlackOnTH = 1.2f,
  trace_slackInterval = 1.5f, trace_minImprovementFactor = 2;
};

struct ImmPt {                                       // the members of ImmaturePoint that the two functions read or write (ImmaturePoint.h:33-78)
  float u, v, idepth_min, idepth_max;
  float color[8], weights[8], gradH[4];              // gradH row-major 2x2
  float energyTH, quality, lastTraceUV[2], lastTracePixelInterval;
  int32_t lastTraceStatus;
};

static inline void interp33(const float* mat, float x, float y, int width, float out[3]) {   // getInterpolatedElement33, util/globalFuncs.h:51-65
  int ix=(int)x, iy=(int)y; float dx=x-ix, dy=y-iy, dxdy=dx*dy; const float* bp = mat + 3*(ix+iy*width);
  float w11=dxdy, w01=dy-dxdy, w10=dx-dxdy, w00=1-dx-dy+dxdy;
  for (int c=0;c<3;c++) out[c] = w11*bp[3*(1+width)+c] + w01*bp[3*width+c] + w10*bp[3+c] + w00*bp[c];
}
static inline float interp31(const float* mat, float x, float y, int width) {                // getInterpolatedElement31, util/globalFuncs.h:102-116
  int ix=(int)x, iy=(int)y; float dx=x-ix, dy=y-iy, dxdy=dx*dy; const float* bp = mat + 3*(ix+iy*width);
  return dxdy*bp[3*(1+width)] + (dy-dxdy)*bp[3*width] + (dx-dxdy)*bp[3] + (1-dx-dy+dxdy)*bp[0];
}
static inline void interp33BiLin(const float* mat, float x, float y, int width, float out[3]) {   // getInterpolatedElement33BiLin, util/globalFuncs.h:142-164
  int ix=(int)x, iy=(int)y; const float* bp = mat + 3*(ix+iy*width);
  float tl=bp[0], tr=bp[3], bl=bp[3*width], br=bp[3*(width+1)];
  float dx=x-ix, dy=y-iy;
  float topInt = dx*tr + (1-dx)*tl, botInt = dx*br + (1-dx)*bl, leftInt = dy*bl + (1-dy)*tl, rightInt = dy*br + (1-dy)*tr;
  out[0] = dx*rightInt + (1-dx)*leftInt; out[1] = rightInt-leftInt; out[2] = botInt-topInt;
}

// ImmaturePoint::ImmaturePoint (ImmaturePoint.cpp:8-36)
void immatureInit(const float* hostdI, int w, int u_, int v_, const TraceSettings& S, ImmPt& p) {
  p.u = (float)u_; p.v = (float)v_; p.idepth_min = 0; p.idepth_max = NAN; p.lastTraceStatus = IPS_UNINITIALIZED;
  for (int i=0;i<4;i++) p.gradH[i] = 0;
  p.lastTraceUV[0] = p.lastTraceUV[1] = 0; p.quality = 10000; p.lastTracePixelInterval = 0;
  for (int idx=0;idx<8;idx++) { p.color[idx] = 0; p.weights[idx] = 0; }
  for (int idx=0;idx<8;idx++) {
    int dx = kPat[idx][0], dy = kPat[idx][1];
    float ptc[3]; interp33BiLin(hostdI, p.u+dx, p.v+dy, w, ptc);
    p.color[idx] = ptc[0];
    if (!std::isfinite(p.color[idx])) { p.energyTH = NAN; return; }
    p.gradH[0] += ptc[1]*ptc[1]; p.gradH[1] += ptc[1]*ptc[2]; p.gradH[2] += ptc[2]*ptc[1]; p.gradH[3] += ptc[2]*ptc[2];
    p.weights[idx] = sqrtf(S.outlierTHSumComponent / (S.outlierTHSumComponent + (ptc[1]*ptc[1] + ptc[2]*ptc[2])));
  }
  p.energyTH = 8*S.outlierTH;
  p.energyTH *= S.overallEnergyTHWeight*S.overallEnergyTHWeight;
}

// ImmaturePoint::traceOn (ImmaturePoint.cpp:50-352)
int traceOn(ImmPt& p, const float* dI, int wG0, int hG0, const float KRKi[9], const float Kt[3], const float aff[2], const TraceSettings& S) {
  if (p.lastTraceStatus == IPS_OOB) return p.lastTraceStatus;
  float maxPixSearch = (wG0+hG0)*S.maxPixSearch;
  float pr[3]; for (int i=0;i<3;i++) pr[i] = (KRKi[i*3]*p.u + KRKi[i*3+1]*p.v) + KRKi[i*3+2]*1.0f;
  float ptpMin[3]; for (int i=0;i<3;i++) ptpMin[i] = pr[i] + Kt[i]*p.idepth_min;
  float uMin = ptpMin[0]/ptpMin[2], vMin = ptpMin[1]/ptpMin[2];
  auto oob = [&]() { p.lastTraceUV[0] = -1; p.lastTraceUV[1] = -1; p.lastTracePixelInterval = 0; return p.lastTraceStatus = IPS_OOB; };
  if (!(uMin > 4 && vMin > 4 && uMin < wG0-5 && vMin < hG0-5)) return oob();
  float dist, uMax, vMax, ptpMax[3];
  if (std::isfinite(p.idepth_max)) {
    for (int i=0;i<3;i++) ptpMax[i] = pr[i] + Kt[i]*p.idepth_max;
    uMax = ptpMax[0]/ptpMax[2]; vMax = ptpMax[1]/ptpMax[2];
    if (!(uMax > 4 && vMax > 4 && uMax < wG0-5 && vMax < hG0-5)) return oob();
    dist = (uMin-uMax)*(uMin-uMax) + (vMin-vMax)*(vMin-vMax);
    dist = sqrtf(dist);
    if (dist < S.trace_slackInterval) {
      p.lastTraceUV[0] = (uMax+uMin)*0.5f; p.lastTraceUV[1] = (vMax+vMin)*0.5f; p.lastTracePixelInterval = dist;
      return p.lastTraceStatus = IPS_SKIPPED;
    }
  } else {
    dist = maxPixSearch;
    for (int i=0;i<3;i++) ptpMax[i] = pr[i] + Kt[i]*0.01f;                   // project to arbitrary depth to get the direction
    uMax = ptpMax[0]/ptpMax[2]; vMax = ptpMax[1]/ptpMax[2];
    float dx = uMax-uMin, dy = vMax-vMin;
    float d = 1.0f / sqrtf(dx*dx+dy*dy);
    uMax = uMin + dist*dx*d; vMax = vMin + dist*dy*d;
    if (!(uMax > 4 && vMax > 4 && uMax < wG0-5 && vMax < hG0-5)) return oob();
  }
  if (!(p.idepth_min < 0 || (ptpMin[2] > 0.75 && ptpMin[2] < 1.5))) return oob();
  float dx = S.trace_stepsize*(uMax-uMin), dy = S.trace_stepsize*(vMax-vMin);
  const float* g = p.gradH;
  // a = (dx,dy) gradH (dx,dy)^T ; b = (dy,-dx) gradH (dy,-dx)^T : row vector times matrix first, then the dot product
  float a = (dx*g[0] + dy*g[2])*dx + (dx*g[1] + dy*g[3])*dy;
  float ndx = -dx;
  float b = (dy*g[0] + ndx*g[2])*dy + (dy*g[1] + ndx*g[3])*ndx;
  float errorInPixel = 0.2f + 0.2f*(a+b)/a;
  if (errorInPixel*S.trace_minImprovementFactor > dist && std::isfinite(p.idepth_max)) {
    p.lastTraceUV[0] = (uMax+uMin)*0.5f; p.lastTraceUV[1] = (vMax+vMin)*0.5f; p.lastTracePixelInterval = dist;
    return p.lastTraceStatus = IPS_BADCONDITION;
  }
  if (errorInPixel > 10) errorInPixel = 10;
  dx /= dist; dy /= dist;
  if (dist > maxPixSearch) { uMax = uMin + maxPixSearch*dx; vMax = vMin + maxPixSearch*dy; dist = maxPixSearch; }
  int numSteps = (int)(1.9999f + dist / S.trace_stepsize);
  const float R00 = KRKi[0], R01 = KRKi[1], R10 = KRKi[3], R11 = KRKi[4];    // Rplane = KRKi.topLeftCorner<2,2>()
  float randShift = uMin*1000-floorf(uMin*1000);
  float ptx = uMin-randShift*dx, pty = vMin-randShift*dy;
  float rp[8][2];
  for (int idx=0;idx<8;idx++) { float px = (float)kPat[idx][0], py = (float)kPat[idx][1]; rp[idx][0] = R00*px + R01*py; rp[idx][1] = R10*px + R11*py; }
  if (!std::isfinite(dx) || !std::isfinite(dy)) { p.lastTracePixelInterval = 0; p.lastTraceUV[0] = -1; p.lastTraceUV[1] = -1; return p.lastTraceStatus = IPS_OOB; }
  float errors[100]; float bestU = 0, bestV = 0, bestEnergy = 1e10f; int bestIdx = -1;
  if (numSteps >= 100) numSteps = 99;
  for (int i=0;i<numSteps;i++) {
    float energy = 0;
    for (int idx=0;idx<8;idx++) {
      float hitColor = interp31(dI, (float)(ptx+rp[idx][0]), (float)(pty+rp[idx][1]), wG0);
      if (!std::isfinite(hitColor)) { energy += 1e5f; continue; }
      float residual = hitColor - (float)(aff[0]*p.color[idx] + aff[1]);
      float hw = fabsf(residual) < S.huberTH ? 1 : S.huberTH / fabsf(residual);
      energy += hw*residual*residual*(2-hw);
    }
    errors[i] = energy;
    if (energy < bestEnergy) { bestU = ptx; bestV = pty; bestEnergy = energy; bestIdx = i; }
    ptx += dx; pty += dy;
  }
  float secondBest = 1e10f;
  for (int i=0;i<numSteps;i++)
    if ((i < bestIdx-S.minTraceTestRadius || i > bestIdx+S.minTraceTestRadius) && errors[i] < secondBest) secondBest = errors[i];
  float newQuality = secondBest / bestEnergy;
  if (newQuality < p.quality || numSteps > 10) p.quality = newQuality;
  float uBak = bestU, vBak = bestV, gnstepsize = 1, stepBack = 0;
  if (S.trace_GNIterations > 0) bestEnergy = 1e5f;
  for (int it=0; it<S.trace_GNIterations; it++) {
    float H = 1, bb = 0, energy = 0;
    for (int idx=0;idx<8;idx++) {
      float hc[3]; interp33(dI, (float)(bestU+rp[idx][0]), (float)(bestV+rp[idx][1]), wG0, hc);
      if (!std::isfinite(hc[0])) { energy += 1e5f; continue; }
      float residual = hc[0] - (aff[0]*p.color[idx] + aff[1]);
      float dResdDist = dx*hc[1] + dy*hc[2];
      float hw = fabsf(residual) < S.huberTH ? 1 : S.huberTH / fabsf(residual);
      H += hw*dResdDist*dResdDist;
      bb += hw*residual*dResdDist;
      energy += p.weights[idx]*p.weights[idx]*hw*residual*residual*(2-hw);
    }
    if (energy > bestEnergy) {
      stepBack *= 0.5f;                                                       // a smaller step from the old point
      bestU = uBak + stepBack*dx; bestV = vBak + stepBack*dy;
    } else {
      float step = -gnstepsize*bb/H;
      if (step < -0.5f) step = -0.5f; else if (step > 0.5f) step = 0.5f;
      if (!std::isfinite(step)) step = 0;
      uBak = bestU; vBak = bestV; stepBack = step;
      bestU += step*dx; bestV += step*dy; bestEnergy = energy;
    }
    if (fabsf(stepBack) < S.trace_GNThreshold) break;
  }
  if (!(bestEnergy < p.energyTH*S.trace_extraSlackOnTH)) {
    p.lastTracePixelInterval = 0; p.lastTraceUV[0] = -1; p.lastTraceUV[1] = -1;
    if (p.lastTraceStatus == IPS_OUTLIER) return p.lastTraceStatus = IPS_OOB;
    return p.lastTraceStatus = IPS_OUTLIER;
  }
  if (dx*dx > dy*dy) {
    p.idepth_min = (pr[2]*(bestU-errorInPixel*dx) - pr[0]) / (Kt[0] - Kt[2]*(bestU-errorInPixel*dx));
    p.idepth_max = (pr[2]*(bestU+errorInPixel*dx) - pr[0]) / (Kt[0] - Kt[2]*(bestU+errorInPixel*dx));
  } else {
    p.idepth_min = (pr[2]*(bestV-errorInPixel*dy) - pr[1]) / (Kt[1] - Kt[2]*(bestV-errorInPixel*dy));
    p.idepth_max = (pr[2]*(bestV+errorInPixel*dy) - pr[1]) / (Kt[1] - Kt[2]*(bestV+errorInPixel*dy));
  }
  if (p.idepth_min > p.idepth_max) std::swap(p.idepth_min, p.idepth_max);
  if (!std::isfinite(p.idepth_min) || !std::isfinite(p.idepth_max) || (p.idepth_max < 0)) {
    p.lastTracePixelInterval = 0; p.lastTraceUV[0] = -1; p.lastTraceUV[1] = -1;
    return p.lastTraceStatus = IPS_OUTLIER;
  }
  p.lastTracePixelInterval = 2*errorInPixel;
  p.lastTraceUV[0] = bestU; p.lastTraceUV[1] = bestV;
  return p.lastTraceStatus = IPS_GOOD;
}

// ---------------------------------------------------------------------------------------------- activation: FullSystem::optimizeImmaturePoint (FullSystemOptPoint.cpp:18-183)
// over ImmaturePoint::linearizeResidual (ImmaturePoint.cpp:410-476), projectPoint / derive_idepth (ResidualProjections.h:11-59)
enum { RS_IN = 0, RS_OOB, RS_OUTLIER };                                         // ResState (Residuals.h)
struct PairPre { float R[9], t[3], aff[2]; };                                   // FrameFramePrecalc of (host,target): PRE_RTll, PRE_tTll, PRE_aff_mode (HessianBlocks.cpp:169-195)
struct CalibF { float fxl, fyl, cxl, cyl, fxli, fyli; };                        // CalibHessian::value_scaledf / value_scaledi (HessianBlocks.h:305-330)
struct TmpRes { int state_state; double state_energy; int state_NewState; double state_NewEnergy; };   // ImmaturePointTemporaryResidual

double linearizeResidual(const ImmPt& p, const float* dIl, int wG0, int hG0, const CalibF& C, const PairPre& pc, float outlierTHSlack, TmpRes& r, float& Hdd, float& bd, float idepth,
                         const TraceSettings& S) {
  if (r.state_state == RS_OOB) { r.state_NewState = RS_OOB; return r.state_energy; }
  float energyLeft = 0; const float wM3G = (float)(wG0-3), hM3G = (float)(hG0-3);
  for (int idx=0; idx<8; idx++) {
    int dx = kPat[idx][0], dy = kPat[idx][1];
    float KliP[3] = { (p.u+dx-C.cxl)*C.fxli, (p.v+dy-C.cyl)*C.fyli, 1 };
    float ptp[3]; for (int i=0;i<3;i++) ptp[i] = ((pc.R[i*3]*KliP[0] + pc.R[i*3+1]*KliP[1]) + pc.R[i*3+2]*KliP[2]) + pc.t[i]*idepth;
    float drescale = 1.0f/ptp[2];
    if (!(drescale > 0)) { r.state_NewState = RS_OOB; return r.state_energy; }
    float u = ptp[0]*drescale, v = ptp[1]*drescale;
    float Ku = u*C.fxl + C.cxl, Kv = v*C.fyl + C.cyl;
    if (!(Ku > 1.1f && Kv > 1.1f && Ku < wM3G && Kv < hM3G)) { r.state_NewState = RS_OOB; return r.state_energy; }
    float hc[3]; interp33(dIl, Ku, Kv, wG0, hc);
    if (!std::isfinite(hc[0])) { r.state_NewState = RS_OOB; return r.state_energy; }
    float residual = hc[0] - (pc.aff[0]*p.color[idx] + pc.aff[1]);
    float hw = fabsf(residual) < S.huberTH ? 1 : S.huberTH / fabsf(residual);
    energyLeft += p.weights[idx]*p.weights[idx]*hw*residual*residual*(2-hw);
    float dxInterp = hc[1]*C.fxl, dyInterp = hc[2]*C.fyl;
    float d_idepth = (dxInterp*drescale*(pc.t[0]-pc.t[2]*u) + dyInterp*drescale*(pc.t[1]-pc.t[2]*v))*1.0f;   // derive_idepth, SCALE_IDEPTH = 1
    hw *= p.weights[idx]*p.weights[idx];
    Hdd += (hw*d_idepth)*d_idepth;
    bd += (hw*residual)*d_idepth;
  }
  if (energyLeft > p.energyTH*outlierTHSlack) { energyLeft = p.energyTH*outlierTHSlack; r.state_NewState = RS_OUTLIER; }
  else r.state_NewState = RS_IN;
  r.state_NewEnergy = energyLeft;
  return energyLeft;
}

// returns 0: not well constrained (stays immature, `return 0`), -1: outlier / non-finite (deleted), 1: activated (idepth_out, res_state_out[nres] valid)
int optimizeImmaturePoint(const ImmPt& p, bool isFromSensor, int nres, const float* const* dI, int wG0, int hG0, const CalibF& C, const PairPre* pc, int minObs,
                          float* idepth_out, int32_t* res_state_out, const TraceSettings& S) {
  const float minIdepthH_act = 100; const int GNItsOnPointActivation = 3;     // util/settings.cpp:41,133
  TmpRes res[16];
  for (int i=0;i<nres;i++) { res[i].state_NewEnergy = res[i].state_energy = 0; res[i].state_NewState = RS_OUTLIER; res[i].state_state = RS_IN; }
  float lastEnergy = 0, lastHdd = 0, lastbd = 0;
  float currentIdepth = (p.idepth_max+p.idepth_min)*0.5f;
  float trueDepth = currentIdepth;
  if (!isFromSensor) {
    for (int i=0;i<nres;i++) {
      lastEnergy += linearizeResidual(p, dI[i], wG0, hG0, C, pc[i], 1000, res[i], lastHdd, lastbd, currentIdepth, S);   // float += double
      res[i].state_state = res[i].state_NewState; res[i].state_energy = res[i].state_NewEnergy;
    }
    if (!std::isfinite(lastEnergy) || lastHdd < minIdepthH_act) return 0;
    float lambda = 0.1f;
    for (int iteration=0; iteration<GNItsOnPointActivation; iteration++) {
      float H = lastHdd; H *= 1+lambda;
      float step = (1.0/H) * lastbd;
      float newIdepth = currentIdepth - step;
      float newHdd = 0, newbd = 0, newEnergy = 0;
      for (int i=0;i<nres;i++) newEnergy += linearizeResidual(p, dI[i], wG0, hG0, C, pc[i], 1, res[i], newHdd, newbd, newIdepth, S);
      if (!std::isfinite(lastEnergy) || newHdd < minIdepthH_act) return 0;
      if (newEnergy < lastEnergy) {
        currentIdepth = newIdepth; lastHdd = newHdd; lastbd = newbd; lastEnergy = newEnergy;
        for (int i=0;i<nres;i++) { res[i].state_state = res[i].state_NewState; res[i].state_energy = res[i].state_NewEnergy; }
        lambda *= 0.5;
      } else lambda *= 5;
      if (fabsf(step) < 0.0001*currentIdepth) break;
    }
  }
  if (!std::isfinite(currentIdepth)) return -1;
  int numGoodRes = 0; for (int i=0;i<nres;i++) if (res[i].state_state == RS_IN) numGoodRes++;
  if (numGoodRes < minObs) return -1;
  if (!std::isfinite(p.energyTH)) return -1;                                    // PointHessian ctor copies energyTH (HessianBlocks.cpp:17-36); :139
  *idepth_out = isFromSensor ? trueDepth : currentIdepth;
  for (int i=0;i<nres;i++) res_state_out[i] = res[i].state_state;
  return 1;
}

}  // namespace orc

// ---------------------------------------------------------------------------------------------- flat C entry points (ctypes: oracle/orc.py)
extern "C" {
int orc_immature_bytes() { return (int)sizeof(orc::ImmPt); }
// new candidates at integer pixels uv[2n] of the host frame (level-0 {I,dx,dy} image of an orc Frame)
void orc_immature_init(void* host_frame, int n, const int32_t* uv, void* pts_out) {
  orc::Frame* f = (orc::Frame*)host_frame; orc::ImmPt* P = (orc::ImmPt*)pts_out; orc::TraceSettings S;
  for (int i=0;i<n;i++) orc::immatureInit(f->dIp[0].data(), f->w[0], uv[2*i], uv[2*i+1], S, P[i]);
}
// traceOn of n candidates of ONE host against `frame` (KRKi, Kt, aff of that host: FullSystem::traceNewCoarse, FullSystem.cpp:532-538); status_out may be NULL
void orc_immature_trace(void* frame, int n, void* pts_io, const float* KRKi9, const float* Kt3, const float* aff2, int32_t* status_out) {
  orc::Frame* f = (orc::Frame*)frame; orc::ImmPt* P = (orc::ImmPt*)pts_io; orc::TraceSettings S;
  for (int i=0;i<n;i++) { int st = orc::traceOn(P[i], f->dIp[0].data(), f->w[0], f->h[0], KRKi9, Kt3, aff2, S); if (status_out) status_out[i] = st; }
}
// FullSystem::optimizeImmaturePoint for n candidates of ONE host keyframe against nres target frames (frameHessians without the host, window order).
// pre: nres x {R[9], t[3], aff[2]} = host->targetPrecalc[target] (PRE_RTll, PRE_tTll, PRE_aff_mode); calib6: fxl fyl cxl cyl fxli fyli.
void orc_immature_optimize(int n, const void* pts, const uint8_t* isFromSensor, int nres, void* const* target_frames, const float* pre14, const float* calib6, int minObs,
                           int32_t* status_out, float* idepth_out, int32_t* res_state_out) {
  const orc::ImmPt* P = (const orc::ImmPt*)pts; orc::TraceSettings S; orc::CalibF C{calib6[0], calib6[1], calib6[2], calib6[3], calib6[4], calib6[5]};
  const float* dI[16]; orc::PairPre pc[16]; int w = 0, h = 0;
  for (int i=0;i<nres;i++) { orc::Frame* f = (orc::Frame*)target_frames[i]; dI[i] = f->dIp[0].data(); w = f->w[0]; h = f->h[0]; std::memcpy(&pc[i], pre14 + 14*i, 14*sizeof(float)); }
  for (int k=0;k<n;k++) { idepth_out[k] = 0; for (int i=0;i<nres;i++) res_state_out[(size_t)k*nres+i] = -1;
    status_out[k] = orc::optimizeImmaturePoint(P[k], isFromSensor && isFromSensor[k], nres, dI, w, h, C, pc, minObs, idepth_out + k, res_state_out + (size_t)k*nres, S); }
}
}
