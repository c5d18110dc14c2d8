// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned on oracle/_ref (tests/test_ref_pin_ba.py).  See orc_ba.hpp for the file:line map.
#include "orc_ba.hpp"
#include <cstdio>
#include <cassert>
#include <algorithm>

namespace orc {

static const int patternP[8][2] = {{0,-2},{-1,-1},{1,-1},{-2,0},{0,0},{2,0},{-1,1},{0,2}};   // settings.cpp:250 (pattern 8)

// ---------------------------------------------------------------- small dense helpers (row-major, double)
static void mm66(const double* A, const double* B, double* C) { for (int i=0;i<6;i++) for (int j=0;j<6;j++) { double s=0; for (int k=0;k<6;k++) s += A[i*6+k]*B[k*6+j]; C[i*6+j]=s; } }
static void mm66T(const double* A, const double* B, double* C) { for (int i=0;i<6;i++) for (int j=0;j<6;j++) { double s=0; for (int k=0;k<6;k++) s += A[i*6+k]*B[j*6+k]; C[i*6+j]=s; } }   // A * B^T

// ---------------------------------------------------------------- calibration (HessianBlocks.h:260-358)
void BAWindow::setCalibValue(const double v[4]) {
  for (int i=0;i<4;i++) c_value[i]=v[i];
  c_value_scaled[0]=SCALE_F*v[0]; c_value_scaled[1]=SCALE_F*v[1]; c_value_scaled[2]=SCALE_C*v[2]; c_value_scaled[3]=SCALE_C*v[3];
  for (int i=0;i<4;i++) c_sf[i]=(float)c_value_scaled[i];
  c_si[0]=1.0f/c_sf[0]; c_si[1]=1.0f/c_sf[1]; c_si[2]=-c_sf[2]/c_sf[0]; c_si[3]=-c_sf[3]/c_sf[1];
  for (int i=0;i<4;i++) c_vmvz[i]=c_value[i]-c_value_zero[i];
}
void BAWindow::setCalibScaled(const double vs[4]) {
  const float SFI = 1.0f/SCALE_F, SCI = 1.0f/SCALE_C;
  for (int i=0;i<4;i++) { c_value_scaled[i]=vs[i]; c_sf[i]=(float)vs[i]; }
  c_value[0]=SFI*vs[0]; c_value[1]=SFI*vs[1]; c_value[2]=SCI*vs[2]; c_value[3]=SCI*vs[3];
  for (int i=0;i<4;i++) { c_value_zero[i]=c_value[i]; c_vmvz[i]=0; c_step[i]=0; }
  c_si[0]=1.0f/c_sf[0]; c_si[1]=1.0f/c_sf[1]; c_si[2]=-c_sf[2]/c_sf[0]; c_si[3]=-c_sf[3]/c_sf[1];
}

// ---------------------------------------------------------------- frame state (HessianBlocks.h:141-175, HessianBlocks.cpp:52-82)
void BAWindow::frameSetState(BAFrame& f, const double s[10]) {
  for (int i=0;i<10;i++) f.state[i]=s[i];
  for (int i=0;i<3;i++) f.state_scaled[i] = SCALE_XI_TRANS*s[i];
  for (int i=3;i<6;i++) f.state_scaled[i] = SCALE_XI_ROT*s[i];
  f.state_scaled[6]=SCALE_A*s[6]; f.state_scaled[7]=SCALE_B*s[7]; f.state_scaled[8]=SCALE_A*s[8]; f.state_scaled[9]=SCALE_B*s[9];
  f.PRE_worldToCam = SE3::exp(f.state_scaled) * f.worldToCam_evalPT;
  f.PRE_camToWorld = f.PRE_worldToCam.inverse();
}
void BAWindow::frameSetStateZero(BAFrame& f, const double s0[10]) {
  for (int i=0;i<10;i++) f.state_zero[i]=s0[i];
  SE3 T = f.worldToCam_evalPT, Ti = T.inverse();
  for (int i=0;i<6;i++) {
    double eps[6]={0,0,0,0,0,0}, meps[6]={0,0,0,0,0,0}; eps[i]=1e-3; meps[i]=-1e-3;
    SE3 P = (T * SE3::exp(eps)) * Ti, M = (T * SE3::exp(meps)) * Ti;
    double lp[6], lm[6]; P.log(lp); M.log(lm);
    for (int r=0;r<6;r++) f.nullspaces_pose[r][i] = (lp[r]-lm[r])/(2e-3);
  }
  SE3 P = T; for (int i=0;i<3;i++) P.t.v[i] *= 1.00001; P = P * Ti;
  SE3 M = T; for (int i=0;i<3;i++) M.t.v[i] /= 1.00001; M = M * Ti;
  double lp[6], lm[6]; P.log(lp); M.log(lm);
  for (int r=0;r<6;r++) f.nullspaces_scale[r] = (lp[r]-lm[r])/(2e-3);
}
void BAWindow::frameTakeData(BAFrame& f) {          // EFFrame::takeData + FrameHessian::getPrior (mode=1: affine priors irrelevant, a/b frozen in BA)
  for (int i=0;i<6;i++) f.prior[i]=0;
  if (f.frameID==0) { for (int i=0;i<3;i++) f.prior[i]=set.initialTransPrior; for (int i=3;i<6;i++) f.prior[i]=set.initialRotPrior; }
  for (int i=0;i<6;i++) { f.delta[i]=f.state[i]-f.state_zero[i]; f.delta_prior[i]=f.state[i]; }
}

void BAWindow::setAdjointsF() {                      // EnergyFunctional.cpp:21-71
  int n=nF(); adHost.assign((size_t)n*n*36,0); adTarget.assign((size_t)n*n*36,0); adHostF.assign((size_t)n*n*36,0); adTargetF.assign((size_t)n*n*36,0);
  for (int h=0;h<n;h++) for (int t=0;t<n;t++) {
    SE3 hostToTarget = frames[t].worldToCam_evalPT * frames[h].worldToCam_evalPT.inverse();
    double Ad[6][6]; hostToTarget.Adj(Ad);
    double* AH = &adHost[(size_t)(h+t*n)*36]; double* AT = &adTarget[(size_t)(h+t*n)*36];
    for (int r=0;r<6;r++) for (int c=0;c<6;c++) { AH[r*6+c] = -Ad[c][r]; AT[r*6+c] = (r==c)?1.0:0.0; }
    for (int r=0;r<3;r++) for (int c=0;c<6;c++) { AH[r*6+c] *= SCALE_XI_TRANS; AT[r*6+c] *= SCALE_XI_TRANS; }
    for (int r=3;r<6;r++) for (int c=0;c<6;c++) { AH[r*6+c] *= SCALE_XI_ROT; AT[r*6+c] *= SCALE_XI_ROT; }
    for (int i=0;i<36;i++) { adHostF[(size_t)(h+t*n)*36+i]=(float)AH[i]; adTargetF[(size_t)(h+t*n)*36+i]=(float)AT[i]; }
  }
  for (int i=0;i<4;i++) cPrior[i]=set.initialCalibHessian;
}

void BAWindow::setPrecalcValues() {                  // FullSystem.cpp:1358-1368 -> FrameFramePrecalc::set + setDeltaF
  int n=nF(); precalc.resize((size_t)n*n);
  Mat33f K; std::memset(&K,0,sizeof(K)); K.m[0][0]=c_sf[0]; K.m[1][1]=c_sf[1]; K.m[0][2]=c_sf[2]; K.m[1][2]=c_sf[3]; K.m[2][2]=1;
  Mat33f Ki = inverse3<float,Mat33f>(K);
  for (int h=0;h<n;h++) for (int t=0;t<n;t++) {
    Precalc& p = precalc[(size_t)h*n+t]; const BAFrame& host=frames[h]; const BAFrame& target=frames[t];
    SE3 l0 = target.worldToCam_evalPT * host.worldToCam_evalPT.inverse();
    p.PRE_RTll_0 = castf(l0.rotationMatrix()); p.PRE_tTll_0 = castf(l0.t);
    SE3 l = target.PRE_worldToCam * host.PRE_camToWorld;
    p.PRE_RTll = castf(l.rotationMatrix()); p.PRE_tTll = castf(l.t);
    p.PRE_KRKiTll = matmul(matmul(K, p.PRE_RTll), Ki);
    p.PRE_KtTll = matvec(K, p.PRE_tTll[0], p.PRE_tTll[1], p.PRE_tTll[2]);
    AffLight gh; gh.a=host.state_scaled[6]; gh.b=host.state_scaled[7]; AffLight gt; gt.a=target.state_scaled[6]; gt.b=target.state_scaled[7];
    double aff[2]; fromToVecExposure(host.ab_exposure, target.ab_exposure, gh, gt, aff);
    p.PRE_aff_mode[0]=(float)aff[0]; p.PRE_aff_mode[1]=(float)aff[1];
    p.PRE_b0_mode = (float)(host.state_zero[7]*SCALE_B);
  }
  // setDeltaF (EnergyFunctional.cpp:131-156)
  adHTdeltaF.assign((size_t)n*n*6,0);
  for (int h=0;h<n;h++) for (int t=0;t<n;t++) {
    int idx=h+t*n; float dh[6], dt[6];
    for (int i=0;i<6;i++) { dh[i]=(float)(frames[h].state[i]-frames[h].state_zero[i]); dt[i]=(float)(frames[t].state[i]-frames[t].state_zero[i]); }
    for (int j=0;j<6;j++) { float s1=0, s2=0; for (int i=0;i<6;i++) { s1 += dh[i]*adHostF[(size_t)idx*36+i*6+j]; s2 += dt[i]*adTargetF[(size_t)idx*36+i*6+j]; }
      adHTdeltaF[(size_t)idx*6+j] = s1+s2; }
  }
  for (int i=0;i<4;i++) cDeltaF[i]=(float)c_vmvz[i];
  for (auto& f : frames) for (int i=0;i<6;i++) { f.delta[i]=f.state[i]-f.state_zero[i]; f.delta_prior[i]=f.state[i]; }
  for (auto& p : points) p.deltaF = p.idepth - p.idepth_zero;
}

void BAWindow::init() {
  for (auto& f : frames) { double s[10]; std::memcpy(s,f.state,sizeof(s)); frameSetState(f,s); double s0[10]; std::memcpy(s0,f.state_zero,sizeof(s0)); frameSetStateZero(f,s0); frameTakeData(f); for(int i=0;i<10;i++) f.step[i]=0; }
  for (auto& p : points) { p.idepth_scaled=SCALE_IDEPTH*p.idepth; p.idepth_zero_scaled=SCALE_IDEPTH*p.idepth_zero;
    p.priorF = p.hasDepthPrior ? set.idepthFixPrior*SCALE_IDEPTH*SCALE_IDEPTH : 0; p.deltaF = p.idepth-p.idepth_zero; }   // EFPoint::takeData
  int n=dim(); if ((int)HM.size()!=n*n) HM.assign((size_t)n*n,0); if ((int)bM.size()!=n) bM.assign(n,0);
  setAdjointsF(); setPrecalcValues();
}

static inline void interp33(const float* mat, float x, float y, int width, float out[3]) {   // globalFuncs.h:51-65
  int ix=(int)x, iy=(int)y; float dx=x-ix, dy=y-iy, dxdy=dx*dy; const float* bp = mat + 3*(ix+iy*width);
  float w11=dxdy, w01=dy-dxdy, w10=dx-dxdy, w00=1-dx-dy+dxdy;
  for (int c=0;c<3;c++) out[c] = w11*bp[3*(1+width)+c] + w01*bp[3*width+c] + w10*bp[3+c] + w00*bp[c];
}

// ---------------------------------------------------------------- PointFrameResidual::linearize (Residuals.cpp:60-224)
double BAWindow::linearizeOne(BARes& r) {
  linearize_calls++;
  r.state_NewEnergyWithOutlier = -1;
  if (r.state_state == RS_OOB) { r.state_NewState = RS_OOB; return r.state_energy; }
  const Precalc& pc = precalc[(size_t)r.host*nF()+r.target]; const BAPoint& pt = points[r.point];
  const BAFrame& host = frames[r.host]; const BAFrame& target = frames[r.target];
  const float* dIl = target.img->dIp[0].data();
  const float wM3G = w-3, hM3G = h-3;               // globalCalib.cpp:46-47
  float fxl=c_sf[0], fyl=c_sf[1], cxl=c_sf[2], cyl=c_sf[3], fxli=c_si[0], fyli=c_si[1];
  float affLL0=pc.PRE_aff_mode[0], affLL1=pc.PRE_aff_mode[1]; float b0=pc.PRE_b0_mode; (void)b0;
  float d_xi_x[6], d_xi_y[6], d_C_x[4], d_C_y[4], d_d_x, d_d_y, Ku, Kv;
  {
    if (!r.hasMatcher) { r.state_NewState = RS_OOB; return r.state_energy; }
    // projectPoint (ResidualProjections.h:32-59) with dx=dy=0
    float KliP[3] = { (pt.u+0-cxl)*fxli, (pt.v+0-cyl)*fyli, 1 };
    Vec3f ptp = matvec(pc.PRE_RTll_0, KliP[0], KliP[1], KliP[2]); for (int c=0;c<3;c++) ptp.v[c] = ptp.v[c] + pc.PRE_tTll_0[c]*pt.idepth_zero_scaled;
    float drescale = 1.0f/ptp[2]; float new_idepth = pt.idepth_zero_scaled*drescale;
    if (!(drescale>0)) { r.state_NewState = RS_OOB; return r.state_energy; }
    float u = ptp[0]*drescale, v = ptp[1]*drescale;
    Ku = u*fxl + cxl; Kv = v*fyl + cyl;
    if (!(Ku>1.1f && Kv>1.1f && Ku<wM3G && Kv<hM3G)) { r.state_NewState = RS_OOB; return r.state_energy; }
    r.centerProjectedTo[0]=Ku; r.centerProjectedTo[1]=Kv; r.centerProjectedTo[2]=new_idepth;
    const Mat33f& R0 = pc.PRE_RTll_0; const Vec3f& t0 = pc.PRE_tTll_0;
    d_d_x = drescale * (t0[0]-t0[2]*u)*SCALE_IDEPTH*fxl;
    d_d_y = drescale * (t0[1]-t0[2]*v)*SCALE_IDEPTH*fyl;
    d_C_x[2] = drescale*(R0.m[2][0]*u-R0.m[0][0]);
    d_C_x[3] = fxl * drescale*(R0.m[2][1]*u-R0.m[0][1]) * fyli;
    d_C_x[0] = KliP[0]*d_C_x[2];
    d_C_x[1] = KliP[1]*d_C_x[3];
    d_C_y[2] = fyl * drescale*(R0.m[2][0]*v-R0.m[1][0]) * fxli;
    d_C_y[3] = drescale*(R0.m[2][1]*v-R0.m[1][1]);
    d_C_y[0] = KliP[0]*d_C_y[2];
    d_C_y[1] = KliP[1]*d_C_y[3];
    d_C_x[0] = (d_C_x[0]+u)*SCALE_F; d_C_x[1] *= SCALE_F; d_C_x[2] = (d_C_x[2]+1)*SCALE_C; d_C_x[3] *= SCALE_C;
    d_C_y[0] *= SCALE_F; d_C_y[1] = (d_C_y[1]+v)*SCALE_F; d_C_y[2] *= SCALE_C; d_C_y[3] = (d_C_y[3]+1)*SCALE_C;
    d_xi_x[0] = new_idepth*fxl; d_xi_x[1] = 0; d_xi_x[2] = -new_idepth*u*fxl; d_xi_x[3] = -u*v*fxl; d_xi_x[4] = (1+u*u)*fxl; d_xi_x[5] = -v*fxl;
    d_xi_y[0] = 0; d_xi_y[1] = new_idepth*fyl; d_xi_y[2] = -new_idepth*v*fyl; d_xi_y[3] = -(1+v*v)*fyl; d_xi_y[4] = u*v*fyl; d_xi_y[5] = u*fyl;
  }
  RawJ& J = r.J;
  for (int i=0;i<6;i++) { J.Jpdxi[0][i]=d_xi_x[i]; J.Jpdxi[1][i]=d_xi_y[i]; }
  for (int i=0;i<4;i++) { J.Jpdc[0][i]=d_C_x[i]; J.Jpdc[1][i]=d_C_y[i]; }
  J.Jpdd[0]=d_d_x; J.Jpdd[1]=d_d_y;

  float wJI2_sum = 0, energyLeft2 = 0.0;
  for (int idx=0; idx<8; idx++) {                   // photometric 8-pattern gate (:157-194)
    float Ku2, Kv2;
    { Vec3f ptp = matvec(pc.PRE_KRKiTll, pt.u+patternP[idx][0], pt.v+patternP[idx][1], 1.0f);
      for (int c=0;c<3;c++) ptp.v[c] = ptp.v[c] + pc.PRE_KtTll[c]*pt.idepth_scaled;
      Ku2 = ptp[0]/ptp[2]; Kv2 = ptp[1]/ptp[2];
      if (!(Ku2>1.1f && Kv2>1.1f && Ku2<wM3G && Kv2<hM3G)) break; }
    r.projectedTo[idx][0]=Ku2; r.projectedTo[idx][1]=Kv2;
    float hitColor[3]; interp33(dIl, Ku2, Kv2, w, hitColor);
    float residual = hitColor[0] - (float)(affLL0*pt.color[idx] + affLL1);
    if (!std::isfinite(hitColor[0])) break;
    float wgt = sqrtf(set.outlierTHSumComponent / (set.outlierTHSumComponent + (hitColor[1]*hitColor[1]+hitColor[2]*hitColor[2])));
    wgt = 0.5f*(wgt + pt.weights[idx]);
    float hw = fabsf(residual) < set.huberTH ? 1 : set.huberTH / fabsf(residual);
    energyLeft2 += wgt*wgt*hw*residual*residual*(2-hw);
    { if (hw < 1) hw = sqrtf(hw); hw = hw*wgt; hitColor[1]*=hw; hitColor[2]*=hw;
      wJI2_sum += hw*hw*(hitColor[1]*hitColor[1]+hitColor[2]*hitColor[2]); }
  }
  float res0 = Ku - r.matcher[0], res1 = Kv - r.matcher[1];
  float nrm = std::sqrt(res0*res0 + res1*res1);
  float hw = fabsf(nrm) < set.huberTH ? 1 : set.huberTH / fabsf(nrm);
  float energyLeft = hw * (res0*res0+res1*res1)*(2-hw);
  if (hw < 1) hw = sqrtf(hw);
  J.resF[0]=res0*hw; J.resF[1]=res1*hw;
  for (int i=0;i<6;i++) { J.Jpdxi[0][i]*=hw; J.Jpdxi[1][i]*=hw; }
  for (int i=0;i<4;i++) { J.Jpdc[0][i]*=hw; J.Jpdc[1][i]*=hw; }
  J.Jpdd[0]*=hw; J.Jpdd[1]*=hw;
  r.state_NewEnergyWithOutlier = energyLeft2;
  float th = std::max<float>(host.frameEnergyTH, target.frameEnergyTH);
  if (energyLeft2 > th || wJI2_sum < 2) { energyLeft2 = th; r.state_NewState = RS_OUTLIER; }
  else r.state_NewState = RS_IN;
  r.state_NewEnergy = energyLeft2;
  return energyLeft;
}

void BAWindow::applyRes(BARes& r) {                 // Residuals.cpp:252-274 (copyJacobians=true) + takeDataF
  if (r.state_state == RS_OOB) return;
  if (r.state_NewState == RS_IN) {
    r.isActive = 1; std::swap(r.J, r.efJ);
    for (int i=0;i<6;i++) r.JpJdF[i] = r.efJ.Jpdxi[0][i]*r.efJ.Jpdd[0] + r.efJ.Jpdxi[1][i]*r.efJ.Jpdd[1];
    r.JpJdF[6]=r.JpJdF[7]=0;
  } else r.isActive = 0;
  r.state_state = r.state_NewState; r.state_energy = r.state_NewEnergy;
}

void BAWindow::setNewFrameEnergyTH() {              // FullSystemOptimize.cpp:63-97
  std::vector<float> allResVec; int newFrame = nF()-1;
  for (auto& r : res) if (r.state_NewEnergyWithOutlier >= 0 && r.target == newFrame) allResVec.push_back((float)r.state_NewEnergyWithOutlier);
  BAFrame& nf = frames[newFrame];
  if (allResVec.size()==0) { nf.frameEnergyTH = 12*12*8; return; }
  int nthIdx = set.frameEnergyTHN*allResVec.size();
  std::nth_element(allResVec.begin(), allResVec.begin()+nthIdx, allResVec.end());
  float nthElement = sqrtf(allResVec[nthIdx]);
  nf.frameEnergyTH = nthElement*set.frameEnergyTHFacMedian;
  nf.frameEnergyTH = 26.0f*set.frameEnergyTHConstWeight + nf.frameEnergyTH*(1-set.frameEnergyTHConstWeight);
  nf.frameEnergyTH = nf.frameEnergyTH*nf.frameEnergyTH;
  nf.frameEnergyTH *= set.overallEnergyTHWeight*set.overallEnergyTHWeight;
}

double BAWindow::linearizeAll(bool fix) {           // FullSystemOptimize.cpp:99-159 (stats taken as zero-initialised)
  double lastEnergyP = 0;
  for (auto& r : res) {
    lastEnergyP += linearizeOne(r);
    if (fix) {
      applyRes(r);
      if (r.isActive) {
        if (r.isNew) {
          BAPoint& p = points[r.point]; const Precalc& pc = precalc[(size_t)r.host*nF()+r.target];
          Vec3f ptp_inf = matvec(pc.PRE_KRKiTll, p.u, p.v, 1.0f);
          Vec3f ptp; for (int c=0;c<3;c++) ptp.v[c] = ptp_inf.v[c] + pc.PRE_KtTll[c]*p.idepth_scaled;
          float dx = ptp_inf[0]/ptp_inf[2] - ptp[0]/ptp[2], dy = ptp_inf[1]/ptp_inf[2] - ptp[1]/ptp[2];
          float relBS = 0.01*std::sqrt(dx*dx+dy*dy);
          if (relBS > p.maxRelBaseline) p.maxRelBaseline = relBS;
          p.numGoodResiduals++;
        }
      } else r.toRemove = 1;
    }
  }
  setNewFrameEnergyTH();
  return lastEnergyP;
}

// ---------------------------------------------------------------- energies
double BAWindow::calcLEnergy() {                    // EnergyFunctional.cpp:333-350 (+ calcLEnergyPt :295-331; chunks of 50 summed in index order)
  double E = 0;
  for (auto& f : frames) for (int i=0;i<6;i++) E += (f.delta_prior[i]*f.prior[i])*f.delta_prior[i];
  { float s=0; for (int i=0;i<4;i++) s += (cDeltaF[i]*(float)cPrior[i])*cDeltaF[i]; E += s; }
  double stats0 = 0; int np=(int)points.size();
  for (int c0=0; c0<np; c0+=50) {
    float acc=0, acc1k=0; float numIn1=0;           // Accumulator11 lane 0 with the 1k tier (<=50 updates never shifts before finish)
    for (int i=c0; i<std::min(c0+50,np); i++) { acc += points[i].deltaF*points[i].deltaF*points[i].priorF; numIn1++; if (numIn1>1000) { acc1k += acc; acc=0; numIn1=0; } }
    float A = ((acc+acc1k) + 0.0f);
    stats0 += A;
  }
  return E + stats0;
}
double BAWindow::calcMEnergy() {                    // :284-293
  int n=dim(); std::vector<double> d(n);
  for (int i=0;i<4;i++) d[i]=(double)cDeltaF[i];
  for (int h=0;h<nF();h++) for (int i=0;i<6;i++) d[4+6*h+i]=frames[h].delta[i];
  double e=0; for (int i=0;i<n;i++) { double s=0; for (int j=0;j<n;j++) s += HM[(size_t)i*n+j]*d[j]; e += d[i]*(2*bM[i]+s); }
  return e;
}

// ---------------------------------------------------------------- accumulators with the 1 / 1k / 1M float tiers
struct AccApprox {                                  // MatrixAccumulators.h:560-932 (only the cells the back-end reads)
  float D[55], D1k[55], D1m[55], TR[10], TR1k[10], TR1m[10], BR, BR1k, BR1m; float numIn1, numIn1k, numIn1m; size_t num;
  void init() { std::memset(this,0,sizeof(*this)); }
  void shiftUp(bool force) {
    if (numIn1>1000||force) { for(int i=0;i<55;i++) D1k[i]=D[i]+D1k[i]; for(int i=0;i<10;i++) TR1k[i]=TR[i]+TR1k[i]; BR1k=BR+BR1k;
      numIn1k+=numIn1; numIn1=0; std::memset(D,0,sizeof(D)); std::memset(TR,0,sizeof(TR)); BR=0; }
    if (numIn1k>1000||force) { for(int i=0;i<55;i++) D1m[i]=D1k[i]+D1m[i]; for(int i=0;i<10;i++) TR1m[i]=TR1k[i]+TR1m[i]; BR1m=BR1k+BR1m;
      numIn1m+=numIn1k; numIn1k=0; std::memset(D1k,0,sizeof(D1k)); std::memset(TR1k,0,sizeof(TR1k)); BR1k=0; }
  }
  void update(const float* x, const float* y, float a, float b, float c) {     // x,y = [Jpdc(4) ; Jpdxi(6)]
    int k=0; for (int j=0;j<10;j++) for (int i=j;i<10;i++) { D[k] += a*x[i]*x[j] + c*y[i]*y[j] + b*(x[i]*y[j] + y[i]*x[j]); k++; }
    num++; numIn1++; shiftUp(false);
  }
  void finish() { shiftUp(true); num = (size_t)(numIn1+numIn1k+numIn1m); }
};
template<int I, int Jn> struct AccXX { float A[I][Jn], A1k[I][Jn], A1m[I][Jn]; float numIn1, numIn1k, numIn1m; size_t num;
  void init() { std::memset(this,0,sizeof(*this)); }
  void shiftUp(bool force) {
    if (numIn1>1000||force) { for(int i=0;i<I;i++) for(int j=0;j<Jn;j++) { A1k[i][j]+=A[i][j]; A[i][j]=0; } numIn1k+=numIn1; numIn1=0; }
    if (numIn1k>1000||force) { for(int i=0;i<I;i++) for(int j=0;j<Jn;j++) { A1m[i][j]+=A1k[i][j]; A1k[i][j]=0; } numIn1m+=numIn1k; numIn1k=0; } }
  void update(const float* L, const float* R, float w) { for(int i=0;i<I;i++) { float wl=w*L[i]; for(int j=0;j<Jn;j++) A[i][j] += wl*R[j]; } numIn1++; shiftUp(false); }
  void finish() { shiftUp(true); num=(size_t)(numIn1+numIn1k+numIn1m); } };
template<int I> struct AccX { float A[I], A1k[I], A1m[I]; float numIn1, numIn1k, numIn1m; size_t num;
  void init() { std::memset(this,0,sizeof(*this)); }
  void shiftUp(bool force) {
    if (numIn1>1000||force) { for(int i=0;i<I;i++) { A1k[i]+=A[i]; A[i]=0; } numIn1k+=numIn1; numIn1=0; }
    if (numIn1k>1000||force) { for(int i=0;i<I;i++) { A1m[i]+=A1k[i]; A1k[i]=0; } numIn1m+=numIn1k; numIn1k=0; } }
  void update(const float* L, float w) { for(int i=0;i<I;i++) A[i] += w*L[i]; numIn1++; shiftUp(false); }
  void finish() { shiftUp(true); num=(size_t)(numIn1+numIn1k+numIn1m); } };

void BAWindow::accumulateA(std::vector<double>& H, std::vector<double>& b) {   // accumulateAF_MT(MT=false) :158-179
  int n=nF(), N=dim(); std::vector<AccApprox> acc((size_t)n*n); for (auto& a : acc) a.init();
  resInA = 0;
  for (auto& p : points) {                           // addPoint<0> (AccumulatedTopHessian.cpp:13-112)
    float bd_acc=0, Hdd_acc=0, Hcd_acc[4]={0,0,0,0};
    for (int ri=p.res_begin; ri<p.res_end; ri++) {
      BARes& r = res[ri]; if (!r.isActive) continue;
      const RawJ& rJ = r.efJ; int htIDX = r.host + r.target*n;
      float r0=rJ.resF[0], r1=rJ.resF[1]; float rr = r0*r0 + r1*r1;
      float x[10], y[10]; for (int i=0;i<4;i++) { x[i]=rJ.Jpdc[0][i]; y[i]=rJ.Jpdc[1][i]; } for (int i=0;i<6;i++) { x[4+i]=rJ.Jpdxi[0][i]; y[4+i]=rJ.Jpdxi[1][i]; }
      AccApprox& A = acc[htIDX];
      A.update(x, y, 1, 0, 1);
      A.BR += rr;
      for (int i=0;i<10;i++) A.TR[i] += x[i]*r0 + y[i]*r1;
      bd_acc += r0*rJ.Jpdd[0] + r1*rJ.Jpdd[1];
      Hdd_acc += rJ.Jpdd[0]*rJ.Jpdd[0] + rJ.Jpdd[1]*rJ.Jpdd[1];
      for (int i=0;i<4;i++) Hcd_acc[i] += rJ.Jpdc[0][i]*rJ.Jpdd[0] + rJ.Jpdc[1][i]*rJ.Jpdd[1];
      resInA++;
    }
    p.Hdd_accAF=Hdd_acc; p.bd_accAF=bd_acc; for (int i=0;i<4;i++) p.Hcd_accAF[i]=Hcd_acc[i];
    p.Hdd_accLF=0; p.bd_accLF=0; for (int i=0;i<4;i++) p.Hcd_accLF[i]=0;      // addPoint<1>: no linearised residuals in this fork
  }
  H.assign((size_t)N*N,0); b.assign(N,0);
  for (int k=0;k<n*n;k++) {                          // stitchDoubleInternal (AccumulatedTopHessian.cpp:181-242), tid=-1
    int h=k%n, t=k/n, hIdx=CPARS+h*6, tIdx=CPARS+t*6, aidx=h+n*t;
    acc[aidx].finish();
    double accH[10][10], bv[10];
    { int kk=0; for (int r=0;r<10;r++) for (int c=r;c<10;c++) { accH[r][c]=accH[c][r]=(double)acc[aidx].D1m[kk]; kk++; } for (int r=0;r<10;r++) bv[r]=(double)acc[aidx].TR1m[r]; }
    if (acc[aidx].num==0) { for (int r=0;r<10;r++) { for (int c=0;c<10;c++) accH[r][c]=0; bv[r]=0; } }
    const double* AH=&adHost[(size_t)aidx*36]; const double* AT=&adTarget[(size_t)aidx*36];
    double H66[36], H64[24]; for (int r=0;r<6;r++) { for (int c=0;c<6;c++) H66[r*6+c]=accH[4+r][4+c]; for (int c=0;c<4;c++) H64[r*4+c]=accH[4+r][c]; }
    double T1[36], T2[36];
    mm66(AH,H66,T1); mm66T(T1,AH,T2); for (int r=0;r<6;r++) for (int c=0;c<6;c++) H[(size_t)(hIdx+r)*N+hIdx+c] += T2[r*6+c];
    mm66T(T1,AT,T2);                  for (int r=0;r<6;r++) for (int c=0;c<6;c++) H[(size_t)(hIdx+r)*N+tIdx+c] += T2[r*6+c];
    mm66(AT,H66,T1); mm66T(T1,AT,T2); for (int r=0;r<6;r++) for (int c=0;c<6;c++) H[(size_t)(tIdx+r)*N+tIdx+c] += T2[r*6+c];
    for (int r=0;r<6;r++) for (int c=0;c<4;c++) { double s1=0,s2=0; for (int q=0;q<6;q++) { s1 += AH[r*6+q]*H64[q*4+c]; s2 += AT[r*6+q]*H64[q*4+c]; }
      H[(size_t)(hIdx+r)*N+c] += s1; H[(size_t)(tIdx+r)*N+c] += s2; }
    for (int r=0;r<4;r++) for (int c=0;c<4;c++) H[(size_t)r*N+c] += accH[r][c];
    for (int r=0;r<6;r++) { double s1=0,s2=0; for (int q=0;q<6;q++) { s1 += AH[r*6+q]*bv[4+q]; s2 += AT[r*6+q]*bv[4+q]; } b[hIdx+r]+=s1; b[tIdx+r]+=s2; }
    for (int r=0;r<4;r++) b[r] += bv[r];
  }
  for (int i=0;i<4;i++) { H[(size_t)i*N+i] += cPrior[i]; b[i] += cPrior[i]*(double)cDeltaF[i]; }
  for (int h=0;h<n;h++) for (int i=0;i<6;i++) { H[(size_t)(CPARS+h*6+i)*N+CPARS+h*6+i] += frames[h].prior[i]; b[CPARS+h*6+i] += frames[h].prior[i]*frames[h].delta_prior[i]; }
  for (int h=0;h<n;h++) {                            // AccumulatedTopHessian.h:100-113
    int hIdx=CPARS+h*6;
    for (int r=0;r<4;r++) for (int c=0;c<6;c++) H[(size_t)r*N+hIdx+c] = H[(size_t)(hIdx+c)*N+r];
    for (int t=h+1;t<n;t++) { int tIdx=CPARS+t*6;
      for (int r=0;r<6;r++) for (int c=0;c<6;c++) H[(size_t)(hIdx+r)*N+tIdx+c] += H[(size_t)(tIdx+c)*N+hIdx+r];
      for (int r=0;r<6;r++) for (int c=0;c<6;c++) H[(size_t)(tIdx+r)*N+hIdx+c] = H[(size_t)(hIdx+c)*N+tIdx+r]; }
  }
}

void BAWindow::accumulateSC(std::vector<double>& H, std::vector<double>& b) {  // accumulateSCF_MT(MT=false) :202-219
  int n=nF(), N=dim(), n2=n*n;
  std::vector<AccXX<8,4>> accE(n2); std::vector<AccX<8>> accEB(n2); std::vector<AccXX<8,8>> accD((size_t)n2*n);
  AccXX<4,4> accHcc; AccX<4> accbc; accHcc.init(); accbc.init();
  for (auto& a : accE) a.init(); for (auto& a : accEB) a.init(); for (auto& a : accD) a.init();
  for (auto& p : points) {                           // AccumulatedSCHessian.cpp:10-62, shiftPriorToZero=true
    int ngood=0; for (int ri=p.res_begin; ri<p.res_end; ri++) if (res[ri].isActive) ngood++;
    if (ngood==0) { p.HdiF=0; p.bdSumF=0; p.idepth_hessian=0; p.maxRelBaseline=0; continue; }
    float Hh = p.Hdd_accAF+p.Hdd_accLF+p.priorF; if (Hh < 1e-10) Hh = 1e-10;
    p.idepth_hessian=Hh; p.HdiF = 1.0/Hh;
    p.bdSumF = p.bd_accAF + p.bd_accLF; p.bdSumF += p.priorF*p.deltaF;
    float Hcd[4]; for (int i=0;i<4;i++) Hcd[i]=p.Hcd_accAF[i]+p.Hcd_accLF[i];
    if (p.isFromSensor) continue;
    accHcc.update(Hcd,Hcd,p.HdiF); accbc.update(Hcd, p.bdSumF*p.HdiF);
    for (int r1=p.res_begin; r1<p.res_end; r1++) { if (!res[r1].isActive) continue;
      int r1ht = res[r1].host + res[r1].target*n;
      for (int r2=p.res_begin; r2<p.res_end; r2++) { if (!res[r2].isActive) continue;
        accD[(size_t)r1ht + (size_t)res[r2].target*n2].update(res[r1].JpJdF, res[r2].JpJdF, p.HdiF); }
      accE[r1ht].update(res[r1].JpJdF, Hcd, p.HdiF);
      accEB[r1ht].update(res[r1].JpJdF, p.HdiF*p.bdSumF);
    }
  }
  H.assign((size_t)N*N,0); b.assign(N,0);
  for (int k=0;k<n2;k++) {                           // stitchDoubleInternal :64-135, tid=-1
    int i=k%n, j=k/n, iIdx=CPARS+i*6, jIdx=CPARS+j*6, ij=i+n*j;
    accE[ij].finish(); accEB[ij].finish();
    const double* AHij=&adHost[(size_t)ij*36]; const double* ATij=&adTarget[(size_t)ij*36];
    for (int r=0;r<6;r++) for (int c=0;c<4;c++) { double s1=0,s2=0; for (int q=0;q<6;q++) { double e=(double)accE[ij].A1m[q][c]; s1 += AHij[r*6+q]*e; s2 += ATij[r*6+q]*e; }
      H[(size_t)(iIdx+r)*N+c] += s1; H[(size_t)(jIdx+r)*N+c] += s2; }
    for (int r=0;r<6;r++) { double s1=0,s2=0; for (int q=0;q<6;q++) { double e=(double)accEB[ij].A1m[q]; s1 += AHij[r*6+q]*e; s2 += ATij[r*6+q]*e; } b[iIdx+r]+=s1; b[jIdx+r]+=s2; }
    for (int k2=0;k2<n;k2++) {
      int kIdx=CPARS+k2*6, ik=i+n*k2; AccXX<8,8>& D = accD[(size_t)ij + (size_t)k2*n2];
      D.finish(); if (D.num==0) continue;
      double D66[36]; for (int r=0;r<6;r++) for (int c=0;c<6;c++) D66[r*6+c]=(double)D.A1m[r][c];
      const double* AHik=&adHost[(size_t)ik*36]; const double* ATik=&adTarget[(size_t)ik*36];
      double T1[36], T2[36];
      mm66(AHij,D66,T1); mm66T(T1,AHik,T2); for (int r=0;r<6;r++) for (int c=0;c<6;c++) H[(size_t)(iIdx+r)*N+iIdx+c] += T2[r*6+c];
      mm66(ATij,D66,T1); mm66T(T1,ATik,T2); for (int r=0;r<6;r++) for (int c=0;c<6;c++) H[(size_t)(jIdx+r)*N+kIdx+c] += T2[r*6+c];
      mm66T(T1,AHik,T2);                    for (int r=0;r<6;r++) for (int c=0;c<6;c++) H[(size_t)(jIdx+r)*N+iIdx+c] += T2[r*6+c];
      mm66(AHij,D66,T1); mm66T(T1,ATik,T2); for (int r=0;r<6;r++) for (int c=0;c<6;c++) H[(size_t)(iIdx+r)*N+kIdx+c] += T2[r*6+c];
    }
  }
  accHcc.finish(); accbc.finish();
  for (int r=0;r<4;r++) { for (int c=0;c<4;c++) H[(size_t)r*N+c] += (double)accHcc.A1m[r][c]; b[r] += (double)accbc.A1m[r]; }
  for (int h=0;h<n;h++) { int hIdx=CPARS+h*6; for (int r=0;r<4;r++) for (int c=0;c<6;c++) H[(size_t)r*N+hIdx+c] = H[(size_t)(hIdx+c)*N+r]; }
}

// ---------------------------------------------------------------- orthogonalize (EnergyFunctional.cpp:615-648) with getNullspaces (FullSystemOptimize.cpp:548-588)
void BAWindow::orthogonalize(std::vector<double>& x) {
  int n=dim(), nf=nF(); const int m=7;
  std::vector<double> A((size_t)n*m,0);              // columns = normalised nullspace vectors
  for (int i=0;i<m;i++) {
    for (int f=0;f<nf;f++) for (int r=0;r<6;r++) {
      double v = (i<6) ? frames[f].nullspaces_pose[r][i] : frames[f].nullspaces_scale[r];
      v *= (r<3) ? (double)(1.0f/SCALE_XI_TRANS) : (double)(1.0f/SCALE_XI_ROT);
      A[(size_t)(CPARS+f*6+r)*m+i] = v; }
    double nr=0; for (int r=0;r<n;r++) nr += A[(size_t)r*m+i]*A[(size_t)r*m+i]; nr=std::sqrt(nr);
    for (int r=0;r<n;r++) A[(size_t)r*m+i] /= nr;
  }
  // one-sided Jacobi SVD (Hestenes): columns of A -> U*S   (restates Eigen::JacobiSVD up to rounding)
  for (int sweep=0; sweep<60; sweep++) {
    double off=0;
    for (int p=0;p<m;p++) for (int q=p+1;q<m;q++) {
      double al=0, be=0, ga=0; for (int r=0;r<n;r++) { double ap=A[(size_t)r*m+p], aq=A[(size_t)r*m+q]; al+=ap*ap; be+=aq*aq; ga+=ap*aq; }
      if (std::fabs(ga) <= 1e-300 || std::fabs(ga) <= 1e-17*std::sqrt(al*be)) continue;
      off = std::max(off, std::fabs(ga)/std::sqrt(al*be+1e-300));
      double zeta=(be-al)/(2*ga), tt=((zeta>=0)?1.0:-1.0)/(std::fabs(zeta)+std::sqrt(1+zeta*zeta)), c=1/std::sqrt(1+tt*tt), s=c*tt;
      for (int r=0;r<n;r++) { double ap=A[(size_t)r*m+p], aq=A[(size_t)r*m+q]; A[(size_t)r*m+p]=c*ap-s*aq; A[(size_t)r*m+q]=s*ap+c*aq; }
    }
    if (off < 1e-15) break;
  }
  double sv[m], maxSv=0; for (int i=0;i<m;i++) { double nr=0; for (int r=0;r<n;r++) nr += A[(size_t)r*m+i]*A[(size_t)r*m+i]; sv[i]=std::sqrt(nr); maxSv=std::max(maxSv,sv[i]); }
  // N*Npi^T = U diag(S S^+) U^T ; x -= that * x
  std::vector<double> dx(n,0);
  for (int i=0;i<m;i++) { if (!(sv[i] > set.solverModeDelta*maxSv)) continue;
    double dot=0; for (int r=0;r<n;r++) dot += (A[(size_t)r*m+i]/sv[i])*x[r];
    for (int r=0;r<n;r++) dx[r] += (A[(size_t)r*m+i]/sv[i])*dot; }
  for (int r=0;r<n;r++) x[r] -= dx[r];
}

// ---------------------------------------------------------------- solveSystemF (EnergyFunctional.cpp:650-759)
void BAWindow::solveSystem(int iteration, double lambda) {
  int n=nF(), N=dim();
  std::vector<double> HA, bA, Hsc, bsc; accumulateA(HA,bA); accumulateSC(Hsc,bsc);
  std::vector<double> d(N); for (int i=0;i<4;i++) d[i]=(double)cDeltaF[i]; for (int h=0;h<n;h++) for (int i=0;i<6;i++) d[4+6*h+i]=frames[h].delta[i];
  std::vector<double> HF((size_t)N*N), bF(N);
  for (int i=0;i<N;i++) { double s=0; for (int j=0;j<N;j++) s += HM[(size_t)i*N+j]*d[j]; bF[i] = bA[i] + (bM[i]+s) - bsc[i]; }
  for (size_t i=0;i<(size_t)N*N;i++) HF[i] = HA[i] + HM[i] - Hsc[i];
  lastHS = HF; lastbS = bF;
  for (int i=0;i<N;i++) HF[(size_t)i*N+i] *= (1+lambda);
  std::vector<double> S(N), Hs((size_t)N*N), bs(N), y(N), x(N);
  for (int i=0;i<N;i++) S[i] = 1.0/std::sqrt(HF[(size_t)i*N+i]+10);
  for (int i=0;i<N;i++) { for (int j=0;j<N;j++) Hs[(size_t)i*N+j] = S[i]*HF[(size_t)i*N+j]*S[j]; bs[i]=S[i]*bF[i]; }
  ldlt_solve<64>(N, Hs.data(), bs.data(), y.data());
  for (int i=0;i<N;i++) x[i]=S[i]*y[i];
  if (iteration >= 2) orthogonalize(x);
  lastX = x;
  // resubstituteF_MT / resubstituteFPt (:221-282)
  std::vector<float> xF(N); for (int i=0;i<N;i++) xF[i]=(float)x[i];
  for (int i=0;i<4;i++) c_step[i] = -x[i];
  std::vector<float> xAd((size_t)n*n*6);
  for (int h=0;h<n;h++) { for (int i=0;i<6;i++) frames[h].step[i] = -x[CPARS+6*h+i]; for (int i=6;i<10;i++) frames[h].step[i]=0;
    for (int t=0;t<n;t++) for (int j=0;j<6;j++) { float s1=0,s2=0; for (int i=0;i<6;i++) { s1 += xF[CPARS+6*h+i]*adHostF[(size_t)(h+n*t)*36+i*6+j]; s2 += xF[CPARS+6*t+i]*adTargetF[(size_t)(h+n*t)*36+i*6+j]; }
      xAd[(size_t)(n*h+t)*6+j] = s1+s2; } }
  for (auto& p : points) {
    int ngood=0; for (int ri=p.res_begin; ri<p.res_end; ri++) if (res[ri].isActive) ngood++;
    if (ngood==0) { p.step=0; continue; }
    float bb = p.bdSumF; { float s=0; for (int i=0;i<4;i++) s += xF[i]*p.Hcd_accAF[i]; bb -= s; }
    for (int ri=p.res_begin; ri<p.res_end; ri++) { const BARes& r=res[ri]; if (!r.isActive) continue;
      float s=0; for (int i=0;i<6;i++) s += xAd[(size_t)(r.host*n+r.target)*6+i]*r.JpJdF[i]; bb -= s; }
    p.step = p.isFromSensor ? 0 : -bb*p.HdiF;
  }
}

// ---------------------------------------------------------------- step / backup (FullSystemOptimize.cpp:165-321)
bool BAWindow::doStepFromBackup(float stepfac) {
  float sumT=0, sumR=0, sumID=0, numID=0, sumNID=0;
  double v[4]; for (int i=0;i<4;i++) v[i]=c_value_backup[i]+stepfac*c_step[i]; setCalibValue(v);
  for (auto& f : frames) {
    for (int i=6;i<10;i++) f.step[i]=0;
    double s[10]; for (int i=0;i<10;i++) s[i]=f.state_backup[i]+(double)stepfac*f.step[i]; frameSetState(f,s);
    sumT += f.step[0]*f.step[0]+f.step[1]*f.step[1]+f.step[2]*f.step[2];
    sumR += f.step[3]*f.step[3]+f.step[4]*f.step[4]+f.step[5]*f.step[5];
  }
  for (auto& p : points) {                           // (loop is per frame in the reference; sums are order-insensitive up to float rounding)
    p.idepth = p.idepth_backup + stepfac*p.step; p.idepth_scaled = SCALE_IDEPTH*p.idepth;
    sumID += p.step*p.step; sumNID += fabsf(p.idepth_backup); numID++;
    p.idepth_zero = p.idepth_backup + stepfac*p.step; p.idepth_zero_scaled = SCALE_IDEPTH*p.idepth_zero;
  }
  sumR /= frames.size(); sumT /= frames.size(); sumID /= numID; sumNID /= numID;
  setPrecalcValues();
  return sqrtf(sumR) < 0.00005*set.thOptIterations && sqrtf(sumT)*sumNID < 0.00005*set.thOptIterations;
}
void BAWindow::backupState() {
  for (int i=0;i<4;i++) c_value_backup[i]=c_value[i];
  for (auto& f : frames) for (int i=0;i<10;i++) f.state_backup[i]=f.state[i];
  for (auto& p : points) p.idepth_backup=p.idepth;
}
void BAWindow::loadStateBackup() {
  setCalibValue(c_value_backup);
  for (auto& f : frames) { double s[10]; std::memcpy(s,f.state_backup,sizeof(s)); frameSetState(f,s); }
  for (auto& p : points) { p.idepth=p.idepth_backup; p.idepth_scaled=SCALE_IDEPTH*p.idepth; p.idepth_zero=p.idepth_backup; p.idepth_zero_scaled=SCALE_IDEPTH*p.idepth_zero; }
  setPrecalcValues();
}

// ---------------------------------------------------------------- FullSystem::optimize (FullSystemOptimize.cpp:344-502)
float BAWindow::optimize(int mnumOptIts) {
  if (nF() < 2) return 0;
  if (nF() < 3) mnumOptIts = 100;
  if (nF() < 4) mnumOptIts = 75;
  opt_iterations = opt_accepts = 0;
  for (auto& r : res) { r.state_NewEnergy = r.state_energy = 0; r.state_NewState = RS_OUTLIER; r.state_state = RS_IN; }   // resetOOB
  double lastEnergy = linearizeAll(false);
  double lastEnergyL = calcLEnergy(), lastEnergyM = calcMEnergy();
  for (auto& r : res) applyRes(r);
  double lambda = 1e-1;
  for (int iteration=0; iteration<mnumOptIts; iteration++) {
    opt_iterations++;
    backupState();
    solveSystem(iteration, lambda);
    bool canbreak = doStepFromBackup(1.0f);
    double newEnergy = linearizeAll(false);
    double newEnergyL = calcLEnergy(), newEnergyM = calcMEnergy();
    if (newEnergy + 0 + newEnergyL + newEnergyM < lastEnergy + 0 + lastEnergyL + lastEnergyM) {
      opt_accepts++;
      for (auto& r : res) applyRes(r);
      lastEnergy=newEnergy; lastEnergyL=newEnergyL; lastEnergyM=newEnergyM; lambda *= 0.25;
    } else {
      loadStateBackup();
      lastEnergy = linearizeAll(false); lastEnergyL = calcLEnergy(); lastEnergyM = calcMEnergy();
      lambda *= 1e2;
    }
    if (canbreak && iteration >= set.minOptIterations) break;
  }
  BAFrame& nf = frames.back();
  double newStateZero[10]={0,0,0,0,0,0,0,0,0,0}; newStateZero[6]=nf.state[6]; newStateZero[7]=nf.state[7];
  nf.worldToCam_evalPT = nf.PRE_worldToCam; frameSetState(nf,newStateZero); frameSetStateZero(nf,newStateZero);   // setEvalPT (HessianBlocks.h:177-183)
  setAdjointsF(); setPrecalcValues();
  lastEnergy = linearizeAll(true);
  return sqrtf((float)(lastEnergy / resInA));
}


// ================================================================================================ keyframe hand-over: marginalisation
// FullSystem::flagPointsForRemoval, numeric part (FullSystem.cpp:764-797).  `selected[p]` = the host-side predicate
// (ph->isOOB(...) || host->flaggedForMarginalization) && ph->isInlierNew()  — graph bookkeeping, evaluated by the caller.
void BAWindow::flagPointsForRemoval(const int* selected, int* status) {
  const int n = nF();
  for (size_t pi=0; pi<points.size(); pi++) {
    BAPoint& p = points[pi]; status[pi] = 0; if (!selected[pi]) continue;
    for (int ri=p.res_begin; ri<p.res_end; ri++) {
      BARes& r = res[ri];
      if (r.toRemove) continue;                                                                              // deleted by linearizeAll(fix) in the reference (FullSystemOptimize.cpp:129-157) — found by the oracle/_ref pin
      r.state_NewEnergy = r.state_energy = 0; r.state_NewState = RS_OUTLIER; r.state_state = RS_IN;       // resetOOB
      linearizeOne(r);
      r.isLinearized = 0;
      applyRes(r);
      if (r.isActive) {                                                                                      // fixLinearizationF (EnergyFunctionalStructs.cpp:46-55)
        const float* dp = &adHTdeltaF[(size_t)(r.host + n*r.target)*6]; const RawJ& J = r.efJ;
        float dx0=0, dx1=0, dc0=0, dc1=0;
        for (int i=0;i<6;i++) { dx0 += J.Jpdxi[0][i]*dp[i]; dx1 += J.Jpdxi[1][i]*dp[i]; }
        for (int i=0;i<4;i++) { dc0 += J.Jpdc[0][i]*cDeltaF[i]; dc1 += J.Jpdc[1][i]*cDeltaF[i]; }
        float Jp_delta_x = dx0 + dc0 + J.Jpdd[0]*p.deltaF;
        float Jp_delta_y = dx1 + dc1 + J.Jpdd[1]*p.deltaF;
        r.res_toZeroF[0] = J.resF[0] - Jp_delta_x; r.res_toZeroF[1] = J.resF[1] - Jp_delta_y;
        r.isLinearized = 1;
      }
    }
    status[pi] = (p.idepth_hessian > 50.0f) ? 2 : 1;                                                         // setting_minIdepthH_marg, settings.cpp:42
  }
}

// EnergyFunctional::marginalizePointsF (EnergyFunctional.cpp:514-576): single-threaded accumulation (tid 0), non-MT stitchDouble.
void BAWindow::marginalizePointsF(const int* status) {
  const int n=nF(), N=dim(), n2=n*n;
  std::vector<AccApprox> acc((size_t)n2); for (auto& a : acc) a.init();
  std::vector<AccXX<8,4>> accE(n2); std::vector<AccX<8>> accEB(n2); std::vector<AccXX<8,8>> accD((size_t)n2*n);
  AccXX<4,4> accHcc; AccX<4> accbc; accHcc.init(); accbc.init();
  for (auto& a : accE) a.init(); for (auto& a : accEB) a.init(); for (auto& a : accD) a.init();
  for (size_t pi=0; pi<points.size(); pi++) if (status[pi] == 2) points[pi].priorF *= 600.0f*600.0f;         // setting_idepthFixPriorMargFac (:527)
  for (size_t pi=0; pi<points.size(); pi++) { if (status[pi] != 2) continue; BAPoint& p = points[pi];
    // ---- addPoint<2> (AccumulatedTopHessian.cpp:13-112): active residuals, resApprox = res_toZeroF
    float bd_acc=0, Hdd_acc=0, Hcd_acc[4]={0,0,0,0};
    for (int ri=p.res_begin; ri<p.res_end; ri++) { BARes& r = res[ri]; if (!r.isActive) continue;
      const RawJ& rJ = r.efJ; int htIDX = r.host + r.target*n;
      float r0=r.res_toZeroF[0], r1=r.res_toZeroF[1]; float rr = r0*r0 + r1*r1;
      float x[10], y[10]; for (int i=0;i<4;i++) { x[i]=rJ.Jpdc[0][i]; y[i]=rJ.Jpdc[1][i]; } for (int i=0;i<6;i++) { x[4+i]=rJ.Jpdxi[0][i]; y[4+i]=rJ.Jpdxi[1][i]; }
      AccApprox& A = acc[htIDX]; A.update(x, y, 1, 0, 1); A.BR += rr;
      for (int i=0;i<10;i++) A.TR[i] += x[i]*r0 + y[i]*r1;
      bd_acc += r0*rJ.Jpdd[0] + r1*rJ.Jpdd[1];
      Hdd_acc += rJ.Jpdd[0]*rJ.Jpdd[0] + rJ.Jpdd[1]*rJ.Jpdd[1];
      for (int i=0;i<4;i++) Hcd_acc[i] += rJ.Jpdc[0][i]*rJ.Jpdd[0] + rJ.Jpdc[1][i]*rJ.Jpdd[1];
    }
    p.Hdd_accLF=Hdd_acc; p.bd_accLF=bd_acc; for (int i=0;i<4;i++) p.Hcd_accLF[i]=Hcd_acc[i];
    p.Hdd_accAF=0; p.bd_accAF=0; for (int i=0;i<4;i++) p.Hcd_accAF[i]=0;
    // ---- accSSE_bot->addPoint(p, false)  (AccumulatedSCHessian.cpp:10-62)
    int ngood=0; for (int ri=p.res_begin; ri<p.res_end; ri++) if (res[ri].isActive) ngood++;
    if (ngood==0) { p.HdiF=0; p.bdSumF=0; p.idepth_hessian=0; p.maxRelBaseline=0; continue; }
    float Hh = p.Hdd_accAF+p.Hdd_accLF+p.priorF; if (Hh < 1e-10) Hh = 1e-10;
    p.idepth_hessian=Hh; p.HdiF = 1.0/Hh;
    p.bdSumF = p.bd_accAF + p.bd_accLF;
    float Hcd[4]; for (int i=0;i<4;i++) Hcd[i]=p.Hcd_accAF[i]+p.Hcd_accLF[i];
    if (p.isFromSensor) continue;
    accHcc.update(Hcd,Hcd,p.HdiF); accbc.update(Hcd, p.bdSumF*p.HdiF);
    for (int r1=p.res_begin; r1<p.res_end; r1++) { if (!res[r1].isActive) continue;
      int r1ht = res[r1].host + res[r1].target*n;
      for (int r2=p.res_begin; r2<p.res_end; r2++) { if (!res[r2].isActive) continue;
        accD[(size_t)r1ht + (size_t)res[r2].target*n2].update(res[r1].JpJdF, res[r2].JpJdF, p.HdiF); }
      accE[r1ht].update(res[r1].JpJdF, Hcd, p.HdiF);
      accEB[r1ht].update(res[r1].JpJdF, p.HdiF*p.bdSumF);
    }
  }
  // ---- accSSE_top_A->stitchDouble(M, Mb, this, usePrior=false, useDelta=false)   AccumulatedTopHessian.cpp:118-179 (h outer, t inner)
  std::vector<double>& M = margM; std::vector<double>& Mb = margMb; M.assign((size_t)N*N,0); Mb.assign(N,0);
  for (int h=0;h<n;h++) for (int t=0;t<n;t++) {
    int hIdx=CPARS+h*6, tIdx=CPARS+t*6, aidx=h+n*t;
    acc[aidx].finish(); if (acc[aidx].num==0) continue;
    double accH[10][10], bv[10];
    { int kk=0; for (int r=0;r<10;r++) for (int c=r;c<10;c++) { accH[r][c]=accH[c][r]=(double)acc[aidx].D1m[kk]; kk++; } for (int r=0;r<10;r++) bv[r]=(double)acc[aidx].TR1m[r]; }
    const double* AH=&adHost[(size_t)aidx*36]; const double* AT=&adTarget[(size_t)aidx*36];
    double H66[36], H64[24]; for (int r=0;r<6;r++) { for (int c=0;c<6;c++) H66[r*6+c]=accH[4+r][4+c]; for (int c=0;c<4;c++) H64[r*4+c]=accH[4+r][c]; }
    double T1[36], T2[36];
    mm66(AH,H66,T1); mm66T(T1,AH,T2); for (int r=0;r<6;r++) for (int c=0;c<6;c++) M[(size_t)(hIdx+r)*N+hIdx+c] += T2[r*6+c];
    mm66(AT,H66,T1); mm66T(T1,AT,T2); for (int r=0;r<6;r++) for (int c=0;c<6;c++) M[(size_t)(tIdx+r)*N+tIdx+c] += T2[r*6+c];
    mm66(AH,H66,T1); mm66T(T1,AT,T2); for (int r=0;r<6;r++) for (int c=0;c<6;c++) M[(size_t)(hIdx+r)*N+tIdx+c] += T2[r*6+c];
    for (int r=0;r<6;r++) for (int c=0;c<4;c++) { double s1=0,s2=0; for (int q=0;q<6;q++) { s1 += AH[r*6+q]*H64[q*4+c]; s2 += AT[r*6+q]*H64[q*4+c]; }
      M[(size_t)(hIdx+r)*N+c] += s1; M[(size_t)(tIdx+r)*N+c] += s2; }
    for (int r=0;r<4;r++) for (int c=0;c<4;c++) M[(size_t)r*N+c] += accH[r][c];
    for (int r=0;r<6;r++) { double s1=0,s2=0; for (int q=0;q<6;q++) { s1 += AH[r*6+q]*bv[4+q]; s2 += AT[r*6+q]*bv[4+q]; } Mb[hIdx+r]+=s1; Mb[tIdx+r]+=s2; }
    for (int r=0;r<4;r++) Mb[r] += bv[r];
  }
  for (int h=0;h<n;h++) { int hIdx=CPARS+h*6;
    for (int r=0;r<4;r++) for (int c=0;c<6;c++) M[(size_t)r*N+hIdx+c] = M[(size_t)(hIdx+c)*N+r];
    for (int t=h+1;t<n;t++) { int tIdx=CPARS+t*6;
      for (int r=0;r<6;r++) for (int c=0;c<6;c++) M[(size_t)(hIdx+r)*N+tIdx+c] += M[(size_t)(tIdx+c)*N+hIdx+r];
      for (int r=0;r<6;r++) for (int c=0;c<6;c++) M[(size_t)(tIdx+r)*N+hIdx+c] = M[(size_t)(hIdx+c)*N+tIdx+r]; } }
  // ---- accSSE_bot->stitchDouble(Msc, Mbsc, this)   AccumulatedSCHessian.cpp:136-195 (i outer, j inner)
  std::vector<double>& S = margMsc; std::vector<double>& Sb = margMbsc; S.assign((size_t)N*N,0); Sb.assign(N,0);
  for (int i=0;i<n;i++) for (int j=0;j<n;j++) {
    int iIdx=CPARS+i*6, jIdx=CPARS+j*6, ij=i+n*j;
    accE[ij].finish(); accEB[ij].finish();
    const double* AHij=&adHost[(size_t)ij*36]; const double* ATij=&adTarget[(size_t)ij*36];
    for (int r=0;r<6;r++) for (int c=0;c<4;c++) { double s1=0,s2=0; for (int q=0;q<6;q++) { double e=(double)accE[ij].A1m[q][c]; s1 += AHij[r*6+q]*e; s2 += ATij[r*6+q]*e; }
      S[(size_t)(iIdx+r)*N+c] += s1; S[(size_t)(jIdx+r)*N+c] += s2; }
    for (int r=0;r<6;r++) { double s1=0,s2=0; for (int q=0;q<6;q++) { double e=(double)accEB[ij].A1m[q]; s1 += AHij[r*6+q]*e; s2 += ATij[r*6+q]*e; } Sb[iIdx+r]+=s1; Sb[jIdx+r]+=s2; }
    for (int k2=0;k2<n;k2++) {
      int kIdx=CPARS+k2*6, ik=i+n*k2; AccXX<8,8>& D = accD[(size_t)ij + (size_t)k2*n2];
      D.finish(); if (D.num==0) continue;
      double D66[36]; for (int r=0;r<6;r++) for (int c=0;c<6;c++) D66[r*6+c]=(double)D.A1m[r][c];
      const double* AHik=&adHost[(size_t)ik*36]; const double* ATik=&adTarget[(size_t)ik*36];
      double T1[36], T2[36];
      mm66(AHij,D66,T1); mm66T(T1,AHik,T2); for (int r=0;r<6;r++) for (int c=0;c<6;c++) S[(size_t)(iIdx+r)*N+iIdx+c] += T2[r*6+c];
      mm66(ATij,D66,T1); mm66T(T1,ATik,T2); for (int r=0;r<6;r++) for (int c=0;c<6;c++) S[(size_t)(jIdx+r)*N+kIdx+c] += T2[r*6+c];
      mm66T(T1,AHik,T2);                    for (int r=0;r<6;r++) for (int c=0;c<6;c++) S[(size_t)(jIdx+r)*N+iIdx+c] += T2[r*6+c];
      mm66(AHij,D66,T1); mm66T(T1,ATik,T2); for (int r=0;r<6;r++) for (int c=0;c<6;c++) S[(size_t)(iIdx+r)*N+kIdx+c] += T2[r*6+c];
    }
  }
  accHcc.finish(); accbc.finish();
  for (int r=0;r<4;r++) { for (int c=0;c<4;c++) S[(size_t)r*N+c] = (double)accHcc.A1m[r][c]; Sb[r] = (double)accbc.A1m[r]; }
  for (int h=0;h<n;h++) { int hIdx=CPARS+h*6; for (int r=0;r<4;r++) for (int c=0;c<6;c++) S[(size_t)r*N+hIdx+c] = S[(size_t)(hIdx+c)*N+r]; }
  // ---- HM += setting_margWeightFac * (M - Msc)   (:552-567 ; SOLVER_ORTHOGONALIZE_POINTMARG / _FULL not set, settings.cpp:34)
  const double fac = (double)(0.5f*0.5f);
  for (size_t i=0;i<(size_t)N*N;i++) { double Hh = M[i]-S[i]; HM[i] += fac*Hh; }
  for (int i=0;i<N;i++) { double bb = Mb[i]-Sb[i]; bM[i] += fac*bb; }
}

// 6x6 inverse through partial-pivot LU (Eigen's fixed-size path for n > 4): right-looking elimination, column-wise substitution.
static void inverse6_lu(const double* A, double* inv) {
  double L[36]; int perm[6]; for (int i=0;i<36;i++) L[i]=A[i]; for (int i=0;i<6;i++) perm[i]=i;
  for (int k=0;k<6;k++) {
    int piv=k; double big=std::fabs(L[k*6+k]); for (int i=k+1;i<6;i++) { double a=std::fabs(L[i*6+k]); if (a>big) { big=a; piv=i; } }
    if (piv!=k) { for (int j=0;j<6;j++) std::swap(L[k*6+j], L[piv*6+j]); std::swap(perm[k], perm[piv]); }
    for (int i=k+1;i<6;i++) L[i*6+k] /= L[k*6+k];
    for (int i=k+1;i<6;i++) for (int j=k+1;j<6;j++) L[i*6+j] -= L[i*6+k]*L[k*6+j];
  }
  for (int c=0;c<6;c++) { double y[6]; for (int i=0;i<6;i++) y[i] = (perm[i]==c) ? 1.0 : 0.0;
    for (int k=0;k<6;k++) for (int i=k+1;i<6;i++) y[i] -= L[i*6+k]*y[k];
    for (int k=5;k>=0;k--) { y[k] /= L[k*6+k]; for (int i=0;i<k;i++) y[i] -= L[i*6+k]*y[k]; }
    for (int i=0;i<6;i++) inv[i*6+c] = y[i]; }
}

// EnergyFunctional::marginalizeFrame (EnergyFunctional.cpp:434-512): move the frame's block to the end, add its prior, Schur-eliminate it
// in the diagonally pre-scaled system, symmetrise.  The residuals that target the frame and its points are host bookkeeping
// (FullSystemMarginalize.cpp:104-130) and must already be gone from the window.
void BAWindow::marginalizeFrame(int idx) {
  const int n=nF(), odim=dim(), ndim=odim-6;
  std::vector<int> order; for (int i=0;i<odim;i++) { int f=(i-CPARS)/6; if (i<CPARS || f!=idx) order.push_back(i); } for (int i=0;i<6;i++) order.push_back(CPARS+6*idx+i);
  std::vector<double> Hp((size_t)odim*odim), bp(odim);
  for (int i=0;i<odim;i++) { bp[i]=bM[order[i]]; for (int j=0;j<odim;j++) Hp[(size_t)i*odim+j]=HM[(size_t)order[i]*odim+order[j]]; }
  const BAFrame& fh = frames[idx];
  for (int i=0;i<6;i++) { Hp[(size_t)(ndim+i)*odim+ndim+i] += fh.prior[i]; bp[ndim+i] += fh.prior[i]*fh.delta_prior[i]; }
  std::vector<double> SV(odim), SVI(odim);
  for (int i=0;i<odim;i++) { SV[i] = std::sqrt(std::fabs(Hp[(size_t)i*odim+i]) + 10.0); SVI[i] = 1.0/SV[i]; }
  std::vector<double> Hs((size_t)odim*odim), bs(odim);
  for (int i=0;i<odim;i++) { for (int j=0;j<odim;j++) Hs[(size_t)i*odim+j] = (SVI[i]*Hp[(size_t)i*odim+j])*SVI[j]; bs[i] = SVI[i]*bp[i]; }
  double hp[36], hpi[36]; for (int i=0;i<6;i++) for (int j=0;j<6;j++) hp[i*6+j] = Hs[(size_t)(ndim+i)*odim+ndim+j];
  for (int i=0;i<36;i++) hp[i] = 0.5*(hp[i]+hp[i]);                                  // (sic) hpi+hpi, not hpi+hpi^T  (:478,480)
  inverse6_lu(hp, hpi);
  for (int i=0;i<36;i++) hpi[i] = 0.5*(hpi[i]+hpi[i]);
  std::vector<double> bli((size_t)ndim*6);                                            // bottomLeft^T * hpi
  for (int i=0;i<ndim;i++) for (int j=0;j<6;j++) { double s=0; for (int k=0;k<6;k++) s += Hs[(size_t)(ndim+k)*odim+i]*hpi[k*6+j]; bli[(size_t)i*6+j]=s; }
  for (int i=0;i<ndim;i++) { for (int j=0;j<ndim;j++) { double s=0; for (int k=0;k<6;k++) s += bli[(size_t)i*6+k]*Hs[(size_t)(ndim+k)*odim+j]; Hs[(size_t)i*odim+j] -= s; }
    double s=0; for (int k=0;k<6;k++) s += bli[(size_t)i*6+k]*bs[ndim+k]; bs[i] -= s; }
  for (int i=0;i<odim;i++) { for (int j=0;j<odim;j++) Hs[(size_t)i*odim+j] = (SV[i]*Hs[(size_t)i*odim+j])*SV[j]; bs[i] = SV[i]*bs[i]; }
  std::vector<double> Hn((size_t)ndim*ndim), bn(ndim);
  for (int i=0;i<ndim;i++) { bn[i]=bs[i]; for (int j=0;j<ndim;j++) Hn[(size_t)i*ndim+j] = 0.5*(Hs[(size_t)i*odim+j] + Hs[(size_t)j*odim+i]); }
  HM = Hn; bM = bn;
  frames.erase(frames.begin()+idx);
  (void)n;
}

} // namespace orc
