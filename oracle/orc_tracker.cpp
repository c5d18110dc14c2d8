// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned bit for bit on oracle/_ref (tests/test_ref_pin.py).  See orc_tracker.hpp for the file:line map.
#include "orc_tracker.hpp"
#include <cstdio>
#include <cassert>

namespace orc {

int pyrLevelsUsedFor(int w, int h) {               // util/globalCalib.cpp:22-30
  int wlvl=w, hlvl=h, used=1;
  while (wlvl%2==0 && hlvl%2==0 && wlvl*hlvl > 5000 && used < PYR_LEVELS) { wlvl/=2; hlvl/=2; used++; }
  return used;
}

// ---------------------------------------------------------------- FrameHessian::makeImages (HessianBlocks.cpp:107-167)
void Frame::makeImages(const float* color, int w0, int h0, int levels_) {
  levels = levels_;
  for (int l=0;l<levels;l++) { w[l] = w0>>l; h[l] = h0>>l; dIp[l].assign((size_t)3*w[l]*h[l], 0.0f); absSquaredGrad[l].assign((size_t)w[l]*h[l], 0.0f); }
  float* dI = dIp[0].data();
  for (int i=0;i<w[0]*h[0];i++) dI[3*i] = color[i];
  for (int lvl=0; lvl<levels; lvl++) {
    int wl=w[lvl], hl=h[lvl]; float* dI_l = dIp[lvl].data(); float* dabs_l = absSquaredGrad[lvl].data();
    if (lvl>0) {
      int wlm1 = w[lvl-1]; const float* dI_lm = dIp[lvl-1].data();
      for (int y=0;y<hl;y++) for (int x=0;x<wl;x++)
        dI_l[3*(x+y*wl)] = 0.25f * (dI_lm[3*(2*x + 2*y*wlm1)] + dI_lm[3*(2*x+1 + 2*y*wlm1)] +
                                    dI_lm[3*(2*x + 2*y*wlm1+wlm1)] + dI_lm[3*(2*x+1 + 2*y*wlm1+wlm1)]);
    }
    for (int idx=wl; idx < wl*(hl-1); idx++) {      // NOTE flat index: x=0 / x=wl-1 wrap to neighbouring rows, as in the reference
      float dx = 0.5f*(dI_l[3*(idx+1)] - dI_l[3*(idx-1)]);
      float dy = 0.5f*(dI_l[3*(idx+wl)] - dI_l[3*(idx-wl)]);
      if (!std::isfinite(dx)) dx=0;
      if (!std::isfinite(dy)) dy=0;
      dI_l[3*idx+1] = dx; dI_l[3*idx+2] = dy;
      dabs_l[idx] = dx*dx+dy*dy;                    // gamma B is identity without photometric calibration (HessianBlocks.h:293) => gw = 1
    }
  }
}

// ---------------------------------------------------------------- Accumulator9 (MatrixAccumulators.h:934-1293)
void Accumulator9::initialize() {
  std::memset(H,0,sizeof(H)); std::memset(SSEData,0,sizeof(SSEData)); std::memset(SSEData1k,0,sizeof(SSEData1k)); std::memset(SSEData1m,0,sizeof(SSEData1m));
  num = 0; numIn1 = numIn1k = numIn1m = 0;
}
void Accumulator9::shiftUp(bool force) {            // :1273-1292
  if (numIn1 > 1000 || force) {
    for (int i=0;i<4*45;i++) SSEData1k[i] = SSEData[i] + SSEData1k[i];
    numIn1k += numIn1; numIn1 = 0; std::memset(SSEData,0,sizeof(SSEData));
  }
  if (numIn1k > 1000 || force) {
    for (int i=0;i<4*45;i++) SSEData1m[i] = SSEData1k[i] + SSEData1m[i];
    numIn1m += numIn1k; numIn1k = 0; std::memset(SSEData1k,0,sizeof(SSEData1k));
  }
}
void Accumulator9::updateSSE_eighted(const float J[9][4], const float w[4]) {   // :1040-1115
  float* pt = SSEData;
  for (int r=0;r<9;r++) {
    float Jw[4]; for (int l=0;l<4;l++) Jw[l] = J[r][l]*w[l];
    for (int c=r;c<9;c++) { for (int l=0;l<4;l++) pt[l] = pt[l] + Jw[l]*J[c][l]; pt += 4; }
  }
  num += 4; numIn1++; shiftUp(false);
}
void Accumulator9::finish() {                       // :953-970
  std::memset(H,0,sizeof(H)); shiftUp(true);
  int idx=0;
  for (int r=0;r<9;r++) for (int c=r;c<9;c++) {
    float d = SSEData1m[idx+0] + SSEData1m[idx+1] + SSEData1m[idx+2] + SSEData1m[idx+3];
    H[r][c] = H[c][r] = d; idx += 4;
  }
}

// ---------------------------------------------------------------- CoarseTracker
void CoarseTracker::init(int ww, int hh, int levels_) {       // ctor CoarseTracker.cpp:34-69
  levels = levels_;
  for (int l=0;l<levels;l++) {
    int wl = ww>>l, hl = hh>>l; size_t n = (size_t)wl*hl;
    idepth[l].assign(n,0); weightSums[l].assign(n,0); weightSums_bak[l].assign(n,0);
    pc_u[l].assign(n,0); pc_v[l].assign(n,0); pc_idepth[l].assign(n,0); pc_color[l].assign(n,0);
  }
  size_t n0 = (size_t)ww*hh + 4;
  buf_warped_idepth.assign(n0,0); buf_warped_u.assign(n0,0); buf_warped_v.assign(n0,0); buf_warped_dx.assign(n0,0);
  buf_warped_dy.assign(n0,0); buf_warped_residual.assign(n0,0); buf_warped_weight.assign(n0,0); buf_warped_refColor.assign(n0,0);
  w[0]=ww; h[0]=hh;
}

void CoarseTracker::makeK(float fxl, float fyl, float cxl, float cyl) {   // :77-106
  fx[0]=fxl; fy[0]=fyl; cx[0]=cxl; cy[0]=cyl;
  for (int level=1; level<levels; ++level) {
    w[level] = w[0]>>level; h[level] = h[0]>>level;
    fx[level] = fx[level-1]*0.5; fy[level] = fy[level-1]*0.5;
    cx[level] = (cx[0]+0.5)/((int)1<<level) - 0.5;
    cy[level] = (cy[0]+0.5)/((int)1<<level) - 0.5;
  }
  for (int level=0; level<levels; ++level) {
    Mat33f Kl; std::memset(&Kl,0,sizeof(Kl));
    Kl.m[0][0]=fx[level]; Kl.m[0][2]=cx[level]; Kl.m[1][1]=fy[level]; Kl.m[1][2]=cy[level]; Kl.m[2][2]=1.0f;
    K[level]=Kl; Ki[level]=inverse3<float,Mat33f>(Kl);
  }
}

void CoarseTracker::setRefCloud(const Frame* ref, int lvl, int n, const float* u, const float* v, const float* id, const float* color) {
  lastRef = ref;
  for (int i=0;i<n;i++) { pc_u[lvl][i]=u[i]; pc_v[lvl][i]=v[i]; pc_idepth[lvl][i]=id[i]; pc_color[lvl][i]=color[i]; }
  pc_n[lvl]=n;
}

// makeCoarseDepthL0 (:258-425) on a flattened list: the PointHessian graph walk (:264-294) is host bookkeeping; each
// entry is one splat {u,v,idepth,HdiF} with the rounding rule of its branch.
void CoarseTracker::setCoarseTrackingRef(const Frame* ref, const RefPoint* pts, int n, AffLight ref_aff) {
  lastRef = ref; lastRef_aff_g2l = ref_aff; firstCoarseRMSE = -1;
  std::fill(idepth[0].begin(), idepth[0].end(), 0.0f); std::fill(weightSums[0].begin(), weightSums[0].end(), 0.0f);
  for (int k=0;k<n;k++) {
    int u = pts[k].round_half ? (int)(pts[k].u + 0.5f) : (int)pts[k].u;
    int v = pts[k].round_half ? (int)(pts[k].v + 0.5f) : (int)pts[k].v;
    float new_idepth = pts[k].idepth;
    float weight = sqrtf(1e-3 / (pts[k].HdiF+1e-12));
    idepth[0][u+w[0]*v] += new_idepth*weight;
    weightSums[0][u+w[0]*v] += weight;
  }
  for (int lvl=1; lvl<levels; lvl++) {              // :296-322
    int lvlm1=lvl-1, wl=w[lvl], hl=h[lvl], wlm1=w[lvlm1];
    float* idepth_l=idepth[lvl].data(); float* weightSums_l=weightSums[lvl].data();
    const float* idepth_lm=idepth[lvlm1].data(); const float* weightSums_lm=weightSums[lvlm1].data();
    for (int y=0;y<hl;y++) for (int x=0;x<wl;x++) {
      int bidx = 2*x + 2*y*wlm1;
      idepth_l[x+y*wl] = idepth_lm[bidx] + idepth_lm[bidx+1] + idepth_lm[bidx+wlm1] + idepth_lm[bidx+wlm1+1];
      weightSums_l[x+y*wl] = weightSums_lm[bidx] + weightSums_lm[bidx+1] + weightSums_lm[bidx+wlm1] + weightSums_lm[bidx+wlm1+1];
    }
  }
  for (int lvl=0; lvl<2 && lvl<levels; lvl++) {     // diagonal dilation :324-351
    int wh = w[lvl]*h[lvl]-w[lvl], wl = w[lvl];
    float* weightSumsl=weightSums[lvl].data(); float* bak=weightSums_bak[lvl].data(); float* idepthl=idepth[lvl].data();
    std::memcpy(bak, weightSumsl, sizeof(float)*w[lvl]*h[lvl]);
    for (int i=w[lvl]; i<wh; i++) if (bak[i] <= 0) {
      float sum=0, num=0, numn=0;
      const int nAll = w[lvl]*h[lvl];            // the reference indexes -1 (i = w) and w*h (i = wh-1) here: undefined; treated as empty neighbours
      if (i+1+wl < nAll && bak[i+1+wl] > 0) { sum += idepthl[i+1+wl]; num += bak[i+1+wl]; numn++; }
      if (i-1-wl >= 0 && bak[i-1-wl] > 0) { sum += idepthl[i-1-wl]; num += bak[i-1-wl]; numn++; }
      if (bak[i+wl-1] > 0) { sum += idepthl[i+wl-1]; num += bak[i+wl-1]; numn++; }
      if (bak[i-wl+1] > 0) { sum += idepthl[i-wl+1]; num += bak[i-wl+1]; numn++; }
      if (numn>0) { idepthl[i] = sum/numn; weightSumsl[i] = num/numn; }
    }
  }
  for (int lvl=2; lvl<levels; lvl++) {              // axis dilation :354-375
    int wh = w[lvl]*h[lvl]-w[lvl], wl = w[lvl];
    float* weightSumsl=weightSums[lvl].data(); float* bak=weightSums_bak[lvl].data(); float* idepthl=idepth[lvl].data();
    std::memcpy(bak, weightSumsl, sizeof(float)*w[lvl]*h[lvl]);
    for (int i=w[lvl]; i<wh; i++) if (bak[i] <= 0) {
      float sum=0, num=0, numn=0;
      if (bak[i+1] > 0) { sum += idepthl[i+1]; num += bak[i+1]; numn++; }
      if (bak[i-1] > 0) { sum += idepthl[i-1]; num += bak[i-1]; numn++; }
      if (bak[i+wl] > 0) { sum += idepthl[i+wl]; num += bak[i+wl]; numn++; }
      if (bak[i-wl] > 0) { sum += idepthl[i-wl]; num += bak[i-wl]; numn++; }
      if (numn>0) { idepthl[i] = sum/numn; weightSumsl[i] = num/numn; }
    }
  }
  for (int lvl=0; lvl<levels; lvl++) {              // normalise + emit :378-423
    float* weightSumsl=weightSums[lvl].data(); float* idepthl=idepth[lvl].data(); const float* dIRefl = lastRef->dIp[lvl].data();
    int wl=w[lvl], hl=h[lvl], lpc_n=0;
    float* lpc_u=pc_u[lvl].data(); float* lpc_v=pc_v[lvl].data(); float* lpc_idepth=pc_idepth[lvl].data(); float* lpc_color=pc_color[lvl].data();
    for (int y=2;y<hl-2;y++) for (int x=2;x<wl-2;x++) {
      int i = x+y*wl;
      if (weightSumsl[i] > 0) {
        idepthl[i] /= weightSumsl[i];
        lpc_u[lpc_n]=x; lpc_v[lpc_n]=y; lpc_idepth[lpc_n]=idepthl[i]; lpc_color[lpc_n]=dIRefl[3*i];
        if (!std::isfinite(lpc_color[lpc_n]) || !(idepthl[i]>0)) { idepthl[i] = -1; continue; }
        lpc_n++;
      } else idepthl[i] = -1;
      weightSumsl[i] = 1;
    }
    pc_n[lvl]=lpc_n;
  }
}

static inline void interp33(const float* mat, float x, float y, int width, float out[3]) {   // globalFuncs.h:51-65
  int ix=(int)x, iy=(int)y; float dx=x-ix, dy=y-iy, dxdy=dx*dy;
  const float* bp = mat + 3*(ix+iy*width);
  float w11=dxdy, w01=dy-dxdy, w10=dx-dxdy, w00=1-dx-dy+dxdy;
  for (int c=0;c<3;c++) out[c] = w11*bp[3*(1+width)+c] + w01*bp[3*width+c] + w10*bp[3+c] + w00*bp[c];
}

void CoarseTracker::calcRes(int lvl, const SE3& refToNew, AffLight aff_g2l, float cutoffTH, double rs[6]) {   // :486-634
  float E=0; int numTermsInE=0, numTermsInWarped=0, numSaturated=0;
  int wl=w[lvl], hl=h[lvl]; const float* dINewl = newFrame->dIp[lvl].data();
  float fxl=fx[lvl], fyl=fy[lvl], cxl=cx[lvl], cyl=cy[lvl];
  Mat33f RKi = matmul(castf(refToNew.rotationMatrix()), Ki[lvl]);
  Vec3f t = castf(refToNew.t);
  double aff[2]; fromToVecExposure(lastRef->ab_exposure, newFrame->ab_exposure, lastRef_aff_g2l, aff_g2l, aff);
  float affLL0=(float)aff[0], affLL1=(float)aff[1];
  float sumSquaredShiftT=0, sumSquaredShiftRT=0, sumSquaredShiftNum=0;
  float maxEnergy = 2*set.huberTH*cutoffTH - set.huberTH*set.huberTH;
  int nl=pc_n[lvl]; const float* lpc_u=pc_u[lvl].data(); const float* lpc_v=pc_v[lvl].data();
  const float* lpc_idepth=pc_idepth[lvl].data(); const float* lpc_color=pc_color[lvl].data();
  evals[lvl] += nl;
  for (int i=0;i<nl;i++) {
    float id=lpc_idepth[i], x=lpc_u[i], y=lpc_v[i];
    Vec3f pt = matvec(RKi, x, y, 1.0f); for (int c=0;c<3;c++) pt.v[c] = pt.v[c] + t.v[c]*id;
    float u=pt[0]/pt[2], v=pt[1]/pt[2], Ku=fxl*u+cxl, Kv=fyl*v+cyl, new_idepth=id/pt[2];
    if (lvl==0 && i%32==0) {
      Vec3f ptT = matvec(Ki[lvl], x, y, 1.0f); for (int c=0;c<3;c++) ptT.v[c] = ptT.v[c] + t.v[c]*id;
      float uT=ptT[0]/ptT[2], vT=ptT[1]/ptT[2], KuT=fxl*uT+cxl, KvT=fyl*vT+cyl;
      Vec3f ptT2 = matvec(Ki[lvl], x, y, 1.0f); for (int c=0;c<3;c++) ptT2.v[c] = ptT2.v[c] - t.v[c]*id;
      float uT2=ptT2[0]/ptT2[2], vT2=ptT2[1]/ptT2[2], KuT2=fxl*uT2+cxl, KvT2=fyl*vT2+cyl;
      Vec3f pt3 = matvec(RKi, x, y, 1.0f); for (int c=0;c<3;c++) pt3.v[c] = pt3.v[c] - t.v[c]*id;
      float u3=pt3[0]/pt3[2], v3=pt3[1]/pt3[2], Ku3=fxl*u3+cxl, Kv3=fyl*v3+cyl;
      sumSquaredShiftT += (KuT-x)*(KuT-x) + (KvT-y)*(KvT-y);
      sumSquaredShiftT += (KuT2-x)*(KuT2-x) + (KvT2-y)*(KvT2-y);
      sumSquaredShiftRT += (Ku-x)*(Ku-x) + (Kv-y)*(Kv-y);
      sumSquaredShiftRT += (Ku3-x)*(Ku3-x) + (Kv3-y)*(Kv3-y);
      sumSquaredShiftNum += 2;
    }
    if (!(Ku > 2 && Kv > 2 && Ku < wl-3 && Kv < hl-3 && new_idepth > 0)) continue;
    float refColor = lpc_color[i];
    float hitColor[3]; interp33(dINewl, Ku, Kv, wl, hitColor);
    if (!std::isfinite(hitColor[0])) continue;
    float residual = hitColor[0] - (float)(affLL0*refColor + affLL1);
    float hw = std::fabs(residual) < set.huberTH ? 1 : set.huberTH / std::fabs(residual);
    if (std::fabs(residual) > cutoffTH) { E += maxEnergy; numTermsInE++; numSaturated++; }
    else {
      E += hw*residual*residual*(2-hw); numTermsInE++;
      buf_warped_idepth[numTermsInWarped]=new_idepth; buf_warped_u[numTermsInWarped]=u; buf_warped_v[numTermsInWarped]=v;
      buf_warped_dx[numTermsInWarped]=hitColor[1]; buf_warped_dy[numTermsInWarped]=hitColor[2];
      buf_warped_residual[numTermsInWarped]=residual; buf_warped_weight[numTermsInWarped]=hw; buf_warped_refColor[numTermsInWarped]=lpc_color[i];
      numTermsInWarped++;
    }
  }
  while (numTermsInWarped%4 != 0) {
    buf_warped_idepth[numTermsInWarped]=0; buf_warped_u[numTermsInWarped]=0; buf_warped_v[numTermsInWarped]=0; buf_warped_dx[numTermsInWarped]=0;
    buf_warped_dy[numTermsInWarped]=0; buf_warped_residual[numTermsInWarped]=0; buf_warped_weight[numTermsInWarped]=0; buf_warped_refColor[numTermsInWarped]=0;
    numTermsInWarped++;
  }
  buf_warped_n = numTermsInWarped;
  rs[0]=E; rs[1]=numTermsInE; rs[2]=sumSquaredShiftT/(sumSquaredShiftNum+0.1); rs[3]=0;
  rs[4]=sumSquaredShiftRT/(sumSquaredShiftNum+0.1); rs[5]=numSaturated/(float)numTermsInE;
}

void CoarseTracker::calcGSSSE(int lvl, double H_out[64], double b_out[8], const SE3& /*refToNew*/, AffLight aff_g2l) {  // :427-484
  Accumulator9 acc; acc.initialize();
  float fxl=fx[lvl], fyl=fy[lvl]; float b0=(float)lastRef_aff_g2l.b;
  double aff[2]; fromToVecExposure(lastRef->ab_exposure, newFrame->ab_exposure, lastRef_aff_g2l, aff_g2l, aff);
  float a=(float)aff[0];
  int n = buf_warped_n; assert(n%4==0);
  for (int i=0;i<n;i+=4) {
    float J[9][4], wv[4];
    for (int l=0;l<4;l++) {
      float dx = buf_warped_dx[i+l]*fxl, dy = buf_warped_dy[i+l]*fyl;
      float u=buf_warped_u[i+l], v=buf_warped_v[i+l], id=buf_warped_idepth[i+l];
      J[0][l] = id*dx;
      J[1][l] = id*dy;
      J[2][l] = 0.0f - id*(u*dx + v*dy);
      J[3][l] = 0.0f - ((u*v)*dx + dy*(1.0f + v*v));
      J[4][l] = (u*v)*dy + dx*(1.0f + u*u);
      J[5][l] = u*dy - v*dx;
      J[6][l] = a*(b0 - buf_warped_refColor[i+l]);
      J[7][l] = -1.0f;
      J[8][l] = buf_warped_residual[i+l];
      wv[l] = buf_warped_weight[i+l];
    }
    acc.updateSSE_eighted(J, wv);
  }
  acc.finish();
  float invn = 1.0f/n;
  for (int r=0;r<8;r++) { for (int c=0;c<8;c++) H_out[r*8+c] = (double)acc.H[r][c] * invn; b_out[r] = (double)acc.H[r][8] * invn; }
  // :472-483 (note: first 3 = translation columns get SCALE_XI_ROT, next 3 = rotation get SCALE_XI_TRANS, as in the reference)
  const float sc[8] = {SCALE_XI_ROT,SCALE_XI_ROT,SCALE_XI_ROT,SCALE_XI_TRANS,SCALE_XI_TRANS,SCALE_XI_TRANS,SCALE_A,SCALE_B};
  for (int r=0;r<8;r++) for (int c=0;c<8;c++) H_out[r*8+c] *= sc[c];
  for (int r=0;r<8;r++) for (int c=0;c<8;c++) H_out[r*8+c] *= sc[r];
  for (int r=0;r<8;r++) b_out[r] *= sc[r];
}

bool CoarseTracker::trackNewestCoarse(const Frame* newFrameHessian, SE3& lastToNew_out, AffLight& aff_g2l_out,
                                      int coarsestLvl, const double minResForAbort[5]) {      // :662-838
  for (int i=0;i<5;i++) lastResiduals[i]=NAN;
  for (int i=0;i<3;i++) lastFlowIndicators[i]=1000;
  for (int i=0;i<PYR_LEVELS;i++) { evals[i]=0; iterations[i]=0; accepts[i]=0; }
  newFrame = newFrameHessian;
  int maxIterations[] = {10,20,50,50,50};
  float lambdaExtrapolationLimit = 0.001;
  SE3 refToNew_current = lastToNew_out; AffLight aff_g2l_current = aff_g2l_out;
  bool haveRepeated = false;
  for (int lvl=coarsestLvl; lvl>=0; lvl--) {
    double H[64], b[8]; float levelCutoffRepeat=1;
    double resOld[6]; calcRes(lvl, refToNew_current, aff_g2l_current, set.coarseCutoffTH*levelCutoffRepeat, resOld);
    while (resOld[5] > 0.6 && levelCutoffRepeat < 50) {
      levelCutoffRepeat*=2;
      calcRes(lvl, refToNew_current, aff_g2l_current, set.coarseCutoffTH*levelCutoffRepeat, resOld);
    }
    calcGSSSE(lvl, H, b, refToNew_current, aff_g2l_current);
    float lambda = 0.01;
    for (int iteration=0; iteration < maxIterations[lvl]; iteration++) {
      iterations[lvl]++;
      double Hl[64]; std::memcpy(Hl,H,sizeof(Hl));
      for (int i=0;i<8;i++) Hl[i*8+i] *= (1+lambda);
      double nb[8]; for (int i=0;i<8;i++) nb[i] = -b[i];
      double inc[8]; ldlt_solve<8>(8, Hl, nb, inc);
      bool fixA = set.affineOptModeA < 0, fixB = set.affineOptModeB < 0;
      if (fixA && fixB) { double H6[36], x6[6]; for(int r=0;r<6;r++) for(int c=0;c<6;c++) H6[r*6+c]=Hl[r*8+c];
        ldlt_solve<8>(6,H6,nb,x6); for(int i=0;i<6;i++) inc[i]=x6[i]; inc[6]=inc[7]=0; }
      if (!fixA && fixB) { double H7[49], x7[7]; for(int r=0;r<7;r++) for(int c=0;c<7;c++) H7[r*7+c]=Hl[r*8+c];
        ldlt_solve<8>(7,H7,nb,x7); for(int i=0;i<7;i++) inc[i]=x7[i]; inc[7]=0; }
      if (fixA && !fixB) { double Hs[64]; std::memcpy(Hs,Hl,sizeof(Hs)); double bs[8]; std::memcpy(bs,b,sizeof(bs));
        for(int r=0;r<8;r++) Hs[r*8+6]=Hs[r*8+7]; for(int c=0;c<8;c++) Hs[6*8+c]=Hs[7*8+c]; bs[6]=bs[7];
        double H7[49], nb7[7], x7[7]; for(int r=0;r<7;r++){ for(int c=0;c<7;c++) H7[r*7+c]=Hs[r*8+c]; nb7[r]=-bs[r]; }
        ldlt_solve<8>(7,H7,nb7,x7); for(int i=0;i<8;i++) inc[i]=0; for(int i=0;i<6;i++) inc[i]=x7[i]; inc[6]=0; inc[7]=x7[6]; }
      float extrapFac = 1;
      if (lambda < lambdaExtrapolationLimit) extrapFac = std::sqrt(std::sqrt(lambdaExtrapolationLimit / lambda));
      for (int i=0;i<8;i++) inc[i] *= extrapFac;
      double incScaled[8]; std::memcpy(incScaled,inc,sizeof(inc));
      for (int i=0;i<3;i++) incScaled[i] *= SCALE_XI_ROT;
      for (int i=3;i<6;i++) incScaled[i] *= SCALE_XI_TRANS;
      incScaled[6] *= SCALE_A; incScaled[7] *= SCALE_B;
      double s=0; for (int i=0;i<8;i++) s += incScaled[i];
      if (!std::isfinite(s)) for (int i=0;i<8;i++) incScaled[i]=0;
      SE3 refToNew_new = SE3::exp(incScaled) * refToNew_current;
      AffLight aff_g2l_new = aff_g2l_current; aff_g2l_new.a += incScaled[6]; aff_g2l_new.b += incScaled[7];
      double resNew[6]; calcRes(lvl, refToNew_new, aff_g2l_new, set.coarseCutoffTH*levelCutoffRepeat, resNew);
      bool accept = (resNew[0]/resNew[1]) < (resOld[0]/resOld[1]);
      if (accept) {
        accepts[lvl]++;
        calcGSSSE(lvl, H, b, refToNew_new, aff_g2l_new);
        std::memcpy(resOld,resNew,sizeof(resOld)); aff_g2l_current = aff_g2l_new; refToNew_current = refToNew_new;
        lambda *= 0.5;
      } else {
        lambda *= 4;
        if (lambda < lambdaExtrapolationLimit) lambda = lambdaExtrapolationLimit;
      }
      double nrm=0; for (int i=0;i<8;i++) nrm += inc[i]*inc[i]; nrm = std::sqrt(nrm);
      if (!(nrm > 1e-3)) break;
    }
    lastResiduals[lvl] = sqrtf((float)(resOld[0]/resOld[1]));
    lastFlowIndicators[0]=resOld[2]; lastFlowIndicators[1]=resOld[3]; lastFlowIndicators[2]=resOld[4];
    if (lastResiduals[lvl] > 1.5*minResForAbort[lvl]) return false;
    if (levelCutoffRepeat > 1 && !haveRepeated) { lvl++; haveRepeated=true; }
  }
  lastToNew_out = refToNew_current; aff_g2l_out = aff_g2l_current;
  if ((set.affineOptModeA != 0 && (fabsf(aff_g2l_out.a) > 1.2)) || (set.affineOptModeB != 0 && (fabsf(aff_g2l_out.b) > 200))) return false;
  double rel[2]; fromToVecExposure(lastRef->ab_exposure, newFrame->ab_exposure, lastRef_aff_g2l, aff_g2l_out, rel);
  float relAff0=(float)rel[0], relAff1=(float)rel[1];
  if ((set.affineOptModeA == 0 && (fabsf(logf(relAff0)) > 1.5)) || (set.affineOptModeB == 0 && (fabsf(relAff1) > 200))) return false;
  if (set.affineOptModeA < 0) aff_g2l_out.a=0;
  if (set.affineOptModeB < 0) aff_g2l_out.b=0;
  return true;
}

} // namespace orc
