// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).  Pinned on oracle/_ref (the reference's own PixelSelector2.cpp / FullSystem.cpp / CoarseTracker.cpp,
// compiled unmodified): tests/test_ref_pin_select.py.
//
// orc_select.cpp — restatement of the keyframe-rate candidate management (SURVEY.md §8f rank 4 and the caller half of rank 2):
//   PixelSelector::makeHists            FullSystem/PixelSelector2.cpp:47-106    32x32-block gradient histograms -> ths, 3x3-smoothed thsSmoothed
//   PixelSelector::selectFromLidar      FullSystem/PixelSelector2.cpp:451-622   3-level best-gradient pick among the LiDAR pixels of every pot x pot cell
//   PixelSelector::makeMapsFromLidar    FullSystem/PixelSelector2.cpp:354-449   potential adaptation (one recursion) + random sub-selection
//   PixelSelector::select / makeMaps    FullSystem/PixelSelector2.cpp:108-352   the same over all pixels (monocular points, addFeaturePoint)
//   FullSystem::shiTomasiScore          FullSystem/FullSystem.cpp:1540-1583
//   FullSystem::makeNewTraces, setMask  FullSystem/FullSystem.cpp:1261-1356
//   CoarseDistanceMap::makeDistanceMap, growDistBFS, addIntoDistFinal   FullSystem/CoarseTracker.cpp:1139-1282
//   candidate walk of FullSystem::activatePointsMT                      FullSystem/FullSystem.cpp:569-671
// Reference behaviour kept on purpose: selections whose winning index is 0 are dropped (`bestIdx > 0`), thsSmoothed is indexed past its h/32 rows for the last
// image rows (the reference reads its own uninitialised tail there; here the tail is zero, as in a fresh heap block — the _ref pin zeroes it too).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>
#include "orc_tracker.hpp"

namespace orc {

static const float kDirs[16][2] = {{0,1.0000f},{0.3827f,0.9239f},{0.1951f,0.9808f},{0.9239f,0.3827f},{0.7071f,0.7071f},{0.3827f,-0.9239f},{0.8315f,0.5556f},{0.8315f,-0.5556f},
                                   {0.5556f,-0.8315f},{0.9808f,0.1951f},{0.9239f,-0.3827f},{0.7071f,-0.7071f},{0.5556f,0.8315f},{0.9808f,-0.1951f},{1.0000f,0.0000f},{0.1951f,-0.9808f}};

struct Selector {
  int w, h, currentPotential = 3, thsStep = 0;
  std::vector<uint8_t> randomPattern; std::vector<float> ths, thsSmoothed;
  float minGradHistCut = 0.5f, minGradHistAdd = 3, gradDownweightPerLevel = 0.75f;      // util/settings.cpp:119-121
  bool selectDirectionDistribution = true;                                               // :122
  Selector(int w_, int h_, const uint8_t* rp) : w(w_), h(h_), randomPattern(rp, rp + (size_t)w_*h_), ths((w_/32)*(h_/32)+100, 0.f), thsSmoothed((w_/32)*(h_/32)+100, 0.f) {}
};

static int histQuantile(const int* hist, float below) {                                   // computeHistQuantil, PixelSelector2.cpp:35-45
  int th = hist[0]*below + 0.5f;
  for (int i=0;i<90;i++) { th -= hist[i+1]; if (th < 0) return i; }
  return 90;
}

void makeHists(Selector& S, const float* abs0) {                                          // PixelSelector2.cpp:47-106
  const int w = S.w, h = S.h, w32 = w/32, h32 = h/32; S.thsStep = w32;
  for (int y=0;y<h32;y++) for (int x=0;x<w32;x++) {
    int hist[100]; std::memset(hist, 0, sizeof(hist));
    const float* m = abs0 + 32*x + 32*y*w;
    for (int j=0;j<32;j++) for (int i=0;i<32;i++) {
      int it = i+32*x, jt = j+32*y;
      if (it>w-2 || jt>h-2 || it<1 || jt<1) continue;
      int g = sqrtf(m[i+j*w]); if (g > 48) g = 48;
      hist[g+1]++; hist[0]++;
    }
    S.ths[x+y*w32] = histQuantile(hist, S.minGradHistCut) + S.minGradHistAdd;
  }
  for (int y=0;y<h32;y++) for (int x=0;x<w32;x++) {
    float sum = 0, num = 0;
    if (x>0)     { if (y>0) { num++; sum += S.ths[x-1+(y-1)*w32]; } if (y<h32-1) { num++; sum += S.ths[x-1+(y+1)*w32]; } num++; sum += S.ths[x-1+y*w32]; }
    if (x<w32-1) { if (y>0) { num++; sum += S.ths[x+1+(y-1)*w32]; } if (y<h32-1) { num++; sum += S.ths[x+1+(y+1)*w32]; } num++; sum += S.ths[x+1+y*w32]; }
    if (y>0) { num++; sum += S.ths[x+(y-1)*w32]; }
    if (y<h32-1) { num++; sum += S.ths[x+(y+1)*w32]; }
    num++; sum += S.ths[x+y*w32];
    S.thsSmoothed[x+y*w32] = (sum/num)*(sum/num);
  }
}

struct FrameMaps { const float* dI0; const float* abs0; const float* abs1; const float* abs2; int w, h, w1, w2; };

// one candidate pixel against the three running maxima of its cell / 2x2-cell block / 4x4-cell block (the body shared by select and selectFromLidar)
struct Best { int idx2, idx3, idx4; float v2, v3, v4; };
static inline void testPixel(const Selector& S, const FrameMaps& F, float thFactor, float xf, float yf, int idx, int outIdx, const float* d2, const float* d3, const float* d4, Best& B, bool& cont) {
  cont = false;
  const float dw1 = S.gradDownweightPerLevel, dw2 = dw1*dw1;
  float pixelTH0 = S.thsSmoothed[((int)xf>>5) + ((int)yf>>5)*S.thsStep];
  float pixelTH1 = pixelTH0*dw1, pixelTH2 = pixelTH1*dw2;
  float ag0 = F.abs0[idx];
  if (ag0 > pixelTH0*thFactor) {
    float dirNorm = fabsf(F.dI0[3*idx+1]*d2[0] + F.dI0[3*idx+2]*d2[1]);
    if (!S.selectDirectionDistribution) dirNorm = ag0;
    if (dirNorm > B.v2) { B.v2 = dirNorm; B.idx2 = outIdx; B.idx3 = -2; B.idx4 = -2; }
  }
  if (B.idx3 == -2) { cont = true; return; }
  float ag1 = F.abs1[(int)(xf*0.5f+0.25f) + (int)(yf*0.5f+0.25f)*F.w1];
  if (ag1 > pixelTH1*thFactor) {
    float dirNorm = fabsf(F.dI0[3*idx+1]*d3[0] + F.dI0[3*idx+2]*d3[1]);
    if (!S.selectDirectionDistribution) dirNorm = ag1;
    if (dirNorm > B.v3) { B.v3 = dirNorm; B.idx3 = outIdx; B.idx4 = -2; }
  }
  if (B.idx4 == -2) { cont = true; return; }
  float ag2 = F.abs2[(int)(xf*0.25f+0.125) + (int)(yf*0.25f+0.125)*F.w2];
  if (ag2 > pixelTH2*thFactor) {
    float dirNorm = fabsf(F.dI0[3*idx+1]*d4[0] + F.dI0[3*idx+2]*d4[1]);
    if (!S.selectDirectionDistribution) dirNorm = ag2;
    if (dirNorm > B.v4) { B.v4 = dirNorm; B.idx4 = outIdx; }
  }
}

// PixelSelector::select (dense, :202-352) when cloud == nullptr, PixelSelector::selectFromLidar (:451-622) otherwise.  map_out: w*h resp. n floats.
void selectPass(const Selector& S, const FrameMaps& F, float* map_out, int pot, float thFactor, const double* cloud3, int n, int n3out[3]) {
  const int w = F.w, h = F.h;
  const int numPotW = (w%pot==0) ? w/pot : w/pot+1, numPotH = (h%pot==0) ? h/pot : h/pot+1;
  std::vector<std::vector<int>> cells;
  if (cloud3) {
    cells.resize((size_t)numPotW*numPotH);
    for (int i=0;i<n;i++) { int ix = (int)cloud3[3*i] / pot, iy = (int)cloud3[3*i+1] / pot; cells[(size_t)iy*numPotW + ix].push_back(i); }
    std::memset(map_out, 0, (size_t)n*sizeof(float));
  } else std::memset(map_out, 0, (size_t)w*h*sizeof(float));
  int n2 = 0, n3 = 0, n4 = 0;
  for (int y4=0;y4<h;y4+=4*pot) for (int x4=0;x4<w;x4+=4*pot) {
    int my3 = std::min(4*pot, h-y4), mx3 = std::min(4*pot, w-x4);
    Best B; B.idx4 = -1; B.v4 = 0;
    const float* dir4 = kDirs[S.randomPattern[n2] & 0xF];
    for (int y3=0;y3<my3;y3+=2*pot) for (int x3=0;x3<mx3;x3+=2*pot) {
      int x34 = x3+x4, y34 = y3+y4;
      int my2 = std::min(2*pot, h-y34), mx2 = std::min(2*pot, w-x34);
      B.idx3 = -1; B.v3 = 0;
      const float* dir3 = kDirs[S.randomPattern[n2] & 0xF];
      for (int y2=0;y2<my2;y2+=pot) for (int x2=0;x2<mx2;x2+=pot) {
        int x234 = x2+x34, y234 = y2+y34;
        int my1 = std::min(pot, h-y234), mx1 = std::min(pot, w-x234);
        B.idx2 = -1; B.v2 = 0;
        const float* dir2 = kDirs[S.randomPattern[n2] & 0xF];
        bool cont;
        if (cloud3) {
          const std::vector<int>& c = cells[(size_t)(y234/pot)*numPotW + x234/pot];
          for (int j : c) {
            float xf = (float)cloud3[3*j], yf = (float)cloud3[3*j+1];
            int idx = xf + w*yf;
            if (xf<4 || xf>=w-5 || yf<4 || yf>h-4) continue;
            testPixel(S, F, thFactor, xf, yf, idx, j, dir2, dir3, dir4, B, cont);
          }
        } else {
          for (int y1=0;y1<my1;y1++) for (int x1=0;x1<mx1;x1++) {
            int xf = x1+x234, yf = y1+y234, idx = xf + w*yf;
            if (xf<4 || xf>=w-5 || yf<4 || yf>h-4) continue;
            testPixel(S, F, thFactor, (float)xf, (float)yf, idx, idx, dir2, dir3, dir4, B, cont);
          }
        }
        if (B.idx2 > 0) { map_out[B.idx2] = 1; B.v3 = 1e10; n2++; }
      }
      if (B.idx3 > 0) { map_out[B.idx3] = 2; B.v4 = 1e10; n3++; }
    }
    if (B.idx4 > 0) { map_out[B.idx4] = 4; n4++; }
  }
  n3out[0] = n2; n3out[1] = n3; n3out[2] = n4;
}

// PixelSelector::makeMaps (:108-200) / makeMapsFromLidar (:354-449); makeHists must have run for this frame.  passes_out: number of select passes (1 or 2 with one recursion)
int makeMaps(Selector& S, const FrameMaps& F, float* map_out, float density, int recursionsLeft, float thFactor, const double* cloud3, int n, int* passes_out) {
  float numHave = 0, numWant = density, quotia; int idealPotential = S.currentPotential;
  {
    int c3[3]; selectPass(S, F, map_out, S.currentPotential, thFactor, cloud3, n, c3); if (passes_out) (*passes_out)++;
    numHave = c3[0]+c3[1]+c3[2];
    quotia = numWant / numHave;
    float K = numHave * (S.currentPotential+1) * (S.currentPotential+1);
    idealPotential = sqrtf(K/numWant)-1;
    if (idealPotential < 1) idealPotential = 1;
    if (recursionsLeft>0 && quotia > 1.25 && S.currentPotential>1) {
      if (idealPotential >= S.currentPotential) idealPotential = S.currentPotential-1;
      S.currentPotential = idealPotential;
      return makeMaps(S, F, map_out, density, recursionsLeft-1, thFactor, cloud3, n, passes_out);
    } else if (recursionsLeft>0 && quotia < 0.25) {
      if (idealPotential <= S.currentPotential) idealPotential = S.currentPotential+1;
      S.currentPotential = idealPotential;
      return makeMaps(S, F, map_out, density, recursionsLeft-1, thFactor, cloud3, n, passes_out);
    }
  }
  int numHaveSub = numHave;
  if (quotia < 0.95) {
    unsigned char charTH = 255*quotia;
    if (cloud3) {
      for (int i=0;i<n;i++) if (map_out[i] != 0) {
        int rn = (int)(cloud3[3*i] + cloud3[3*i+1]*F.w);
        if (S.randomPattern[rn] > charTH) { map_out[i] = 0; numHaveSub--; }
      }
    } else {
      int rn = 0;
      for (int i=0;i<F.w*F.h;i++) if (map_out[i] != 0) { if (S.randomPattern[rn] > charTH) { map_out[i] = 0; numHaveSub--; } rn++; }
    }
  }
  S.currentPotential = idealPotential;
  return numHaveSub;
}

float shiTomasiScore(const float* dI0, int w, int h, int u, int v) {                      // FullSystem.cpp:1540-1583
  float k = 0.04; float dXX = 0, dYY = 0, dXY = 0;
  const int hb = 4, box = 8, area = 64; const int x_min = u-hb, x_max = u+hb, y_min = v-hb, y_max = v+hb;
  if (x_min < 1 || x_max >= w-1 || y_min < 1 || y_max >= h-1) return 0.0;
  for (int y=y_min;y<y_max;y++) for (int x=0;x<box;x++) {
    float dx = dI0[3*(w*y + x_min+1+x)] - dI0[3*(w*y + x_min-1+x)];
    float dy = dI0[3*(w*(y+1) + x_min+x)] - dI0[3*(w*(y-1) + x_min+x)];
    dXX += dx*dx; dYY += dy*dy; dXY += dx*dy;
  }
  dXX = dXX / (2.0*area); dYY = dYY / (2.0*area); dXY = dXY / (2.0*area);
  // FullSystem.h:19 includes <math.h>: with libstdc++ the unqualified sqrt(float) is the float overload
  float l1 = 0.5*(dXX + dYY - sqrtf((dXX+dYY)*(dXX+dYY) - 4*(dXX*dYY - dXY*dXY)));
  float l2 = 0.5*(dXX + dYY + sqrtf((dXX+dYY)*(dXX+dYY) - 4*(dXX*dYY - dXY*dXY)));
  return (l1*l2 - k*(l1+l2)*(l1+l2));
}

// the 8-pattern colours of ImmaturePoint::ImmaturePoint are all finite <=> energyTH is finite (ImmaturePoint.cpp:20-35)
static bool patternFinite(const float* dI0, int w, int u, int v) {
  static const int pat[8][2] = {{0,-2},{-1,-1},{1,-1},{-2,0},{0,0},{2,0},{-1,1},{0,2}};
  for (int k=0;k<8;k++) { float x = (float)u + pat[k][0], y = (float)v + pat[k][1]; int ix = (int)x, iy = (int)y; const float* bp = dI0 + 3*(ix+iy*w);
    float tl=bp[0], tr=bp[3], bl=bp[3*w], br=bp[3*(w+1)], dx=x-ix, dy=y-iy; float leftInt = dy*bl+(1-dy)*tl, rightInt = dy*br+(1-dy)*tr;
    if (!std::isfinite(dx*rightInt + (1-dx)*leftInt)) return false; }
  return true;
}

struct NewTrace { float u, v, my_type, score, idepth_fromSensor; int32_t isFromSensor, type; };   // type: 0 CORNER, 1 EDGELET (ImmaturePoint.h), -1 not assigned (monocular)

// FullSystem::makeNewTraces (FullSystem.cpp:1273-1356).  selectionMap (w*h floats) persists between calls like the reference's member: when addFeaturePoint is
// false the map of an EARLIER keyframe is walked again (sic).  Returns the immature points in creation order.
int makeNewTraces(Selector& S, const FrameMaps& F, const double* cloud3, int n, float densityLidar, float densityDense, int addFeaturePoint, float* selectionMap,
                  NewTrace* out, int cap, int numPoints[2], int passes[2]) {
  const int w = F.w, h = F.h; std::vector<uint8_t> mask((size_t)w*h, 0); std::vector<float> selL(n > 0 ? n : 1);
  passes[0] = passes[1] = 0;
  numPoints[0] = makeMaps(S, F, selL.data(), densityLidar, 1, 1, cloud3, n, &passes[0]);
  numPoints[1] = addFeaturePoint ? makeMaps(S, F, selectionMap, densityDense, 1, 1, nullptr, 0, &passes[1]) : 0;
  auto setMask = [&](int Ku, int Kv) { for (int i=Ku-S.currentPotential;i<=Ku+S.currentPotential;i++) for (int j=Kv-1;j<=Kv+1;j++) if (j<h && j>=0 && i<w && i>=0) mask[(size_t)j*w+i] = 1; };
  int m = 0; float maxScore = -1000.0;
  for (int i=0;i<n;i++) {
    if (selL[i] == 0) continue;
    int u = cloud3[3*i], v = cloud3[3*i+1];
    float score = shiTomasiScore(F.dI0, w, h, u, v); if (score > maxScore) maxScore = score;
    if (!patternFinite(F.dI0, w, u, v)) continue;
    if (m < cap) { NewTrace& t = out[m]; t.u = u; t.v = v; t.my_type = selL[i]; t.score = score; t.idepth_fromSensor = 1.0/cloud3[3*i+2]; t.isFromSensor = 1; t.type = 0; }
    m++; setMask(u, v);
  }
  float threshold = 0.01;
  for (int k=0;k<m && k<cap;k++) out[k].type = (out[k].score > threshold*maxScore) ? 0 : 1;
  for (int y=3;y<h-4;y++) for (int x=3;x<w-4;x++) {                                       // patternPadding = 2
    int i = x+y*w; if (selectionMap[i] == 0) continue;
    if (!patternFinite(F.dI0, w, x, y) || mask[i] == 1) continue;
    if (m < cap) { NewTrace& t = out[m]; t.u = x; t.v = y; t.my_type = selectionMap[i]; t.score = 0; t.idepth_fromSensor = 0; t.isFromSensor = 0; t.type = -1; }
    m++; setMask(x, y);
  }
  return m;
}

// ---- CoarseDistanceMap (CoarseTracker.cpp:1139-1282) at level-1 resolution
struct DistMap {
  int w1, h1; std::vector<float> d; std::vector<int> l1, l2;
  DistMap(int w1_, int h1_) : w1(w1_), h1(h1_), d((size_t)w1_*h1_, 1000.f), l1(2*(size_t)w1_*h1_), l2(2*(size_t)w1_*h1_) {}
  void grow(int bfsNum) {                                                                  // growDistBFS :1179-1270
    for (int k=1;k<40;k++) {
      int bfsNum2 = bfsNum; std::swap(l1, l2); bfsNum = 0;
      for (int i=0;i<bfsNum2;i++) {
        int x = l2[2*i], y = l2[2*i+1];
        if (x==0 || y==0 || x==w1-1 || y==h1-1) continue;
        int idx = x + y*w1;
        auto visit = [&](int o, int nx, int ny) { if (d[idx+o] > k) { d[idx+o] = k; l1[2*bfsNum] = nx; l1[2*bfsNum+1] = ny; bfsNum++; } };
        visit(1, x+1, y); visit(-1, x-1, y); visit(w1, x, y+1); visit(-w1, x, y-1);
        if (k%2 != 0) { visit(1+w1, x+1, y+1); visit(-1+w1, x-1, y+1); visit(-1-w1, x-1, y-1); visit(1-w1, x+1, y-1); }
      }
    }
  }
  // makeDistanceMap :1139-1173: pts = ACTIVE points of the other keyframes, already grouped by host in window order; KRKi/Kt per host (K[1] R K[0]^-1, K[1] t)
  void make(int nHosts, const int* pt_begin, const float* KRKi9, const float* Kt3, const float* uvid) {
    std::fill(d.begin(), d.end(), 1000.f); int numItems = 0;
    for (int hI=0;hI<nHosts;hI++) for (int p=pt_begin[hI];p<pt_begin[hI+1];p++) {
      const float* M = KRKi9 + 9*hI; const float* t = Kt3 + 3*hI; float pu = uvid[3*p], pv = uvid[3*p+1], id = uvid[3*p+2], ptp[3];
      for (int r=0;r<3;r++) ptp[r] = ((M[3*r]*pu + M[3*r+1]*pv) + M[3*r+2]*1.0f) + t[r]*id;
      int u = ptp[0]/ptp[2] + 0.5f, v = ptp[1]/ptp[2] + 0.5f;
      if (!(u > 0 && v > 0 && u < w1 && v < h1)) continue;
      d[u+w1*v] = 0; l1[2*numItems] = u; l1[2*numItems+1] = v; numItems++;
    }
    grow(numItems);
  }
  void addInto(int u, int v) { l1[0] = u; l1[1] = v; d[u+w1*v] = 0; grow(1); }                // addIntoDistFinal :1272-1278
};

// candidate walk of activatePointsMT (FullSystem.cpp:600-671) for candidates that passed the deletion / canActivate tests (those are pointer-graph
// bookkeeping): cand rows {u, v, 0.5f*(idepth_max+idepth_min), my_type}, grouped by host in window order.  decision: 1 = goes to optimizeImmaturePoint,
// 0 = too close to existing points, -1 = projects outside the level-1 image (deleted).
void activateSelect(DistMap& D, int nHosts, const int* cand_begin, const float* KRKi9, const float* Kt3, const float* cand4, float currentMinActDist, int* decision) {
  for (int hI=0;hI<nHosts;hI++) for (int c=cand_begin[hI];c<cand_begin[hI+1];c++) {
    const float* M = KRKi9 + 9*hI; const float* t = Kt3 + 3*hI; float pu = cand4[4*c], pv = cand4[4*c+1], id = cand4[4*c+2], ptp[3];
    for (int r=0;r<3;r++) ptp[r] = ((M[3*r]*pu + M[3*r+1]*pv) + M[3*r+2]*1.0f) + t[r]*id;
    int u = ptp[0]/ptp[2] + 0.5f, v = ptp[1]/ptp[2] + 0.5f;
    if (u > 0 && v > 0 && u < D.w1 && v < D.h1) {
      float dist = D.d[u+D.w1*v] + (ptp[0]-floorf((float)(ptp[0])));
      if (dist >= currentMinActDist*cand4[4*c+3]) { D.addInto(u, v); decision[c] = 1; } else decision[c] = 0;
    } else decision[c] = -1;
  }
}

}  // namespace orc

// ---------------------------------------------------------------------------------------------- flat C entry points (ctypes: oracle/orc.py)
extern "C" {
void* orc_selector_create(int w, int h, const uint8_t* randomPattern) { return new orc::Selector(w, h, randomPattern); }
void  orc_selector_destroy(void* s) { delete (orc::Selector*)s; }
void  orc_selector_set_potential(void* s, int p) { ((orc::Selector*)s)->currentPotential = p; }
int   orc_selector_get_potential(void* s) { return ((orc::Selector*)s)->currentPotential; }
static orc::FrameMaps maps_of(void* frame) { orc::Frame* f = (orc::Frame*)frame; return orc::FrameMaps{f->dIp[0].data(), f->absSquaredGrad[0].data(), f->absSquaredGrad[1].data(), f->absSquaredGrad[2].data(), f->w[0], f->h[0], f->w[1], f->w[2]}; }
void  orc_selector_make_hists(void* s, void* frame, float* ths, float* thsSmoothed) { orc::Selector* S = (orc::Selector*)s; orc::Frame* f = (orc::Frame*)frame; orc::makeHists(*S, f->absSquaredGrad[0].data());
  int n = (S->w/32)*(S->h/32); if (ths) std::copy(S->ths.begin(), S->ths.begin()+n, ths); if (thsSmoothed) std::copy(S->thsSmoothed.begin(), S->thsSmoothed.begin()+n, thsSmoothed); }
void  orc_selector_select(void* s, void* frame, float* map_out, int pot, float thFactor, const double* cloud3, int n, int n3[3]) { orc::selectPass(*(orc::Selector*)s, maps_of(frame), map_out, pot, thFactor, cloud3, n, n3); }
int   orc_selector_make_maps(void* s, void* frame, float* map_out, float density, int recursionsLeft, float thFactor, const double* cloud3, int n) {
  orc::Selector* S = (orc::Selector*)s; orc::Frame* f = (orc::Frame*)frame; orc::makeHists(*S, f->absSquaredGrad[0].data()); return orc::makeMaps(*S, maps_of(frame), map_out, density, recursionsLeft, thFactor, cloud3, n, nullptr); }
float orc_shi_tomasi(void* frame, int u, int v) { orc::Frame* f = (orc::Frame*)frame; return orc::shiTomasiScore(f->dIp[0].data(), f->w[0], f->h[0], u, v); }
int   orc_new_trace_bytes() { return (int)sizeof(orc::NewTrace); }
int   orc_make_new_traces(void* s, void* frame, const double* cloud3, int n, float densityLidar, float densityDense, int addFeaturePoint, float* selectionMap, void* out, int cap, int numPoints[2], int passes[2]) {
  orc::Selector* S = (orc::Selector*)s; orc::Frame* f = (orc::Frame*)frame; orc::makeHists(*S, f->absSquaredGrad[0].data());
  return orc::makeNewTraces(*S, maps_of(frame), cloud3, n, densityLidar, densityDense, addFeaturePoint, selectionMap, (orc::NewTrace*)out, cap, numPoints, passes); }
void* orc_distmap_create(int w1, int h1) { return new orc::DistMap(w1, h1); }
void  orc_distmap_destroy(void* d) { delete (orc::DistMap*)d; }
void  orc_distmap_make(void* d, int nHosts, const int* pt_begin, const float* KRKi9, const float* Kt3, const float* uvid) { ((orc::DistMap*)d)->make(nHosts, pt_begin, KRKi9, Kt3, uvid); }
void  orc_distmap_add(void* d, int u, int v) { ((orc::DistMap*)d)->addInto(u, v); }
void  orc_distmap_get(void* d, float* out) { orc::DistMap* D = (orc::DistMap*)d; std::copy(D->d.begin(), D->d.end(), out); }
void  orc_activate_select(void* d, int nHosts, const int* cand_begin, const float* KRKi9, const float* Kt3, const float* cand4, float currentMinActDist, int* decision) {
  orc::activateSelect(*(orc::DistMap*)d, nHosts, cand_begin, KRKi9, Kt3, cand4, currentMinActDist, decision); }
}
