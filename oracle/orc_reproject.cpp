// ORACLE — TEST INFRASTRUCTURE ONLY.  Pinned on oracle/_ref: tests/test_sequence_parity.py, tests/test_ref_pin.py.
// CPU restatement of the map reprojection / direct feature alignment stage that feeds structPoseEstimation,
// file:line relative to /root/reference/src/FullSystem/Reprojector.cpp:
//   getWarpMatrixAffine :14-37   getBestSearchLevel :39-51   warpAffine :53-86   initializeGrid :100-112 (cell_size 25)
//   reprojectMap :117-156        backprojectMap :158-185     pointQualityComparator :187-195   reprojectCell :198-233
//   findMatchDirect :235-292     isInFrame :326-332          createPatchFromPatchWithBorder :334-344
//   align1D :346-455             align2D :457-560            pixelFrame2UnitFrame .. reprojectPoint :562-616
// The grid's cell visiting order is std::random_shuffle(rand()) in the reference (:111) — not reproducible, so it is an INPUT here.
// Points behind the camera are not rejected by the reference (no depth test in reprojectPoint) and are not rejected here; a non-finite
// projection (depth exactly 0) is undefined behaviour there (cast<int>) and is treated as "not in frame" here.
#include "orc_tracker.hpp"
#include <list>
#include <algorithm>
#include <cstdint>

namespace orc {

static inline void interp33(const float* mat, float x, float y, int width, float out[3]) {   // util/globalFuncs.h:51-65
  int ix=(int)x, iy=(int)y; float dx=x-ix, dy=y-iy, dxdy=dx*dy;
  const float* bp = mat + 3*(ix+iy*width);
  float w11=dxdy, w01=dy-dxdy, w10=dx-dxdy, w00=1-dx-dy+dxdy;
  for (int c=0;c<3;c++) out[c] = w11*bp[3*(1+width)+c] + w01*bp[3*width+c] + w10*bp[3+c] + w00*bp[c];
}

struct MapPoint { float u, v, idepth; int host; int type; };          // PointHessian fields read by the Reprojector; type 0 CORNER, 1 EDGELET

struct Reprojector {
  int w[PYR_LEVELS], h[PYR_LEVELS], levels;
  Mat33d K, Ki;
  std::vector<const Frame*> kf; std::vector<SE3> kfPose; std::vector<AffLight> kfAff;      // frameHessians_ : image, shell->camToWorld, shell->aff_g2l
  bool backup = false;
  static const int halfpatch_size_ = 4, patch_size_ = 8;
  uint8_t patch_[64], patch_with_border_[100];
  int align_max_iter = 10;

  Vec3d pixelFrame2UnitFrame(double x, double y) const { return matvec(Ki, Vec3d{{x, y, 1.0}}); }
  Vec3d pixelFrame2PointWorld(const MapPoint& p) const {
    Vec3d KiP = matvec(Ki, Vec3d{{(double)p.u, (double)p.v, 1.0}});
    double s = (double)(1/p.idepth);
    Vec3d ptRef{{KiP.v[0]*s, KiP.v[1]*s, KiP.v[2]*s}};
    const SE3& c2w = kfPose[p.host]; Vec3d r = qrot(c2w.q, ptRef);
    return Vec3d{{r.v[0]+c2w.t.v[0], r.v[1]+c2w.t.v[1], r.v[2]+c2w.t.v[2]}};
  }
  static Vec3d xform(const SE3& T, const Vec3d& p) { Vec3d r = qrot(T.q, p); return Vec3d{{r.v[0]+T.t.v[0], r.v[1]+T.t.v[1], r.v[2]+T.t.v[2]}}; }
  Vec3d pointWorld2PixelFrame(const SE3& camToWorld, const Vec3d& ptWorld) const {
    SE3 worldToCur = camToWorld.inverse(); Vec3d c = xform(worldToCur, ptWorld);
    c.v[0] = c.v[0]/c.v[2]; c.v[1] = c.v[1]/c.v[2]; c.v[2] = c.v[2]/c.v[2];
    return matvec(K, c);
  }
  void pointRef2PixelCur(const SE3& T_cur_ref, const Vec3d& ptRef, double out[2]) const {
    Vec3d c = xform(T_cur_ref, ptRef);
    c.v[0] = c.v[0]/c.v[2]; c.v[1] = c.v[1]/c.v[2]; c.v[2] = c.v[2]/c.v[2];
    Vec3d px = matvec(K, c); out[0] = px.v[0]; out[1] = px.v[1];
  }
  bool isInFrame(double x, double y, int boundary) const {
    if (!std::isfinite(x) || !std::isfinite(y) || std::fabs(x) > 1e9 || std::fabs(y) > 1e9) return false;
    int ox = (int)x, oy = (int)y;
    return ox >= boundary && ox < w[0]-boundary && oy >= boundary && oy < h[0]-boundary;
  }
  void getWarpMatrixAffine(const double px_ref[2], const Vec3d& xyz_ref, const SE3& T_cr, double A[4]) const {
    const int halfpatch_size = 5;
    Vec3d du = pixelFrame2UnitFrame(px_ref[0] + halfpatch_size, px_ref[1] + 0), dv = pixelFrame2UnitFrame(px_ref[0] + 0, px_ref[1] + halfpatch_size);
    double su = xyz_ref.v[2]/du.v[2], sv = xyz_ref.v[2]/dv.v[2];
    for (int i=0;i<3;i++) { du.v[i] *= su; dv.v[i] *= sv; }
    double pc[2], pu[2], pv[2]; pointRef2PixelCur(T_cr, xyz_ref, pc); pointRef2PixelCur(T_cr, du, pu); pointRef2PixelCur(T_cr, dv, pv);
    A[0] = (pu[0]-pc[0])/halfpatch_size; A[2] = (pu[1]-pc[1])/halfpatch_size;       // col 0 ; A row-major {a00,a01,a10,a11}
    A[1] = (pv[0]-pc[0])/halfpatch_size; A[3] = (pv[1]-pc[1])/halfpatch_size;
  }
  static int getBestSearchLevel(const double A[4], int max_level) {
    int search_level = 0; double D = A[0]*A[3] - A[2]*A[1];
    while (D > 3.0 && search_level < max_level) { search_level += 1; D *= 0.25; }
    return search_level;
  }
  void warpAffine(const double A[4], const Frame* ref, const double px_ref[2], int search_level, int halfpatch_size, uint8_t* patch) const {
    const int patch_size = halfpatch_size*2;
    double det = A[0]*A[3] - A[2]*A[1], invdet = 1.0/det;
    float a00 = (float)(A[3]*invdet), a10 = (float)(-A[2]*invdet), a01 = (float)(-A[1]*invdet), a11 = (float)(A[0]*invdet);
    if (std::isnan(a00)) return;
    uint8_t* pp = patch; float prx = (float)px_ref[0], pry = (float)px_ref[1];
    for (int y=0; y<patch_size; ++y) for (int x=0; x<patch_size; ++x, ++pp) {
      float p0 = (float)(x-halfpatch_size), p1 = (float)(y-halfpatch_size);
      p0 *= (1<<search_level); p1 *= (1<<search_level);
      float qx = (a00*p0 + a01*p1) + prx, qy = (a10*p0 + a11*p1) + pry;
      if (qx < 0 || qy < 0 || qx >= w[0]-1 || qy >= h[0]-1) *pp = 0;
      else { float o[3]; interp33(ref->dIp[0].data(), qx, qy, w[0], o); *pp = (uint8_t)(o[0]); }
    }
  }
  void createPatchFromPatchWithBorder() {
    for (int y=1; y<patch_size_+1; ++y) for (int x=0; x<patch_size_; ++x) patch_[(y-1)*patch_size_+x] = patch_with_border_[y*(patch_size_+2)+1+x];
  }
  bool align1D(const float* cur_img, int level, const float dir[2], double px[2], const float affLL[2]) const {
    const int halfpatch = 4, patch_size = 8; bool converged = false;
    float ref_patch_dv[64]; float H00=0, H01=0, H10=0, H11=0;
    const int ref_step = patch_size+2; int k = 0;
    for (int y=0; y<patch_size; ++y) { const uint8_t* it = patch_with_border_ + (y+1)*ref_step + 1;
      for (int x=0; x<patch_size; ++x, ++it, ++k) {
        float J0 = (float)(0.5*(dir[0]*(it[1] - it[-1]) + dir[1]*(it[ref_step] - it[-ref_step]))), J1 = 1;
        ref_patch_dv[k] = J0; H00 += J0*J0; H01 += J0*J1; H10 += J1*J0; H11 += J1*J1; } }
    float det = H00*H11 - H10*H01, invdet = 1.0f/det;
    float Hi00 = H11*invdet, Hi10 = -H10*invdet, Hi01 = -H01*invdet, Hi11 = H00*invdet;
    float mean_diff = 0, u = (float)px[0], v = (float)px[1];
    const float min_update_squared = 0.03*0.03; const int cur_step = w[level];
    for (int iter=0; iter<align_max_iter; ++iter) {
      int u_r = (int)std::floor(u), v_r = (int)std::floor(v);
      if (u_r < halfpatch || v_r < halfpatch || u_r >= w[level]-halfpatch || v_r >= h[level]-halfpatch) break;
      if (std::isnan(u) || std::isnan(v)) return false;
      float sx = u-u_r, sy = v-v_r;
      float wTL = (1.0-sx)*(1.0-sy), wTR = sx*(1.0-sy), wBL = (1.0-sx)*sy, wBR = sx*sy;
      float Jres0 = 0, Jres1 = 0; int q = 0;
      for (int y=0; y<patch_size; ++y) { const float* it = cur_img + 3*((v_r+y-halfpatch)*cur_step + u_r-halfpatch);
        for (int x=0; x<patch_size; ++x, it += 3, ++q) {
          float search_pixel = wTL*it[0] + wTR*it[3] + wBL*it[3*cur_step] + wBR*it[3*cur_step+3];
          float res = search_pixel - (float)(affLL[0]*patch_[q] + affLL[1]) + mean_diff;
          Jres0 -= res*ref_patch_dv[q]; Jres1 -= res; } }
      float up0 = Hi00*Jres0 + Hi01*Jres1, up1 = Hi10*Jres0 + Hi11*Jres1;
      u += up0*dir[0]; v += up0*dir[1]; mean_diff += up1;
      if (up0*up0 + up1*up1 < min_update_squared) { converged = true; break; }
    }
    px[0] = u; px[1] = v; return converged;
  }
  bool align2D(const float* cur_img, int level, double px[2], const float affLL[2]) const {
    const int halfpatch = 4, patch_size = 8; bool converged = false;
    float dxp[64], dyp[64]; Mat33f H; std::memset(&H, 0, sizeof(H));
    const int ref_step = patch_size+2; int k = 0;
    for (int y=0; y<patch_size; ++y) { const uint8_t* it = patch_with_border_ + (y+1)*ref_step + 1;
      for (int x=0; x<patch_size; ++x, ++it, ++k) {
        float J[3] = { (float)(0.5*(it[1] - it[-1])), (float)(0.5*(it[ref_step] - it[-ref_step])), 1.0f };
        dxp[k] = J[0]; dyp[k] = J[1];
        for (int a=0;a<3;a++) for (int b=0;b<3;b++) H.m[a][b] += J[a]*J[b]; } }
    Mat33f Hinv = inverse3<float,Mat33f>(H);
    float mean_diff = 0, u = (float)px[0], v = (float)px[1];
    const float min_update_squared = 0.03*0.03; const int cur_step = w[level];
    for (int iter=0; iter<align_max_iter; ++iter) {
      int u_r = (int)std::floor(u), v_r = (int)std::floor(v);
      if (u_r < halfpatch || v_r < halfpatch || u_r >= w[level]-halfpatch || v_r >= h[level]-halfpatch) break;
      if (std::isnan(u) || std::isnan(v)) return false;
      float sx = u-u_r, sy = v-v_r;
      float wTL = (1.0-sx)*(1.0-sy), wTR = sx*(1.0-sy), wBL = (1.0-sx)*sy, wBR = sx*sy;
      float Jres[3] = {0,0,0}; int q = 0;
      for (int y=0; y<patch_size; ++y) { const float* it = cur_img + 3*((v_r+y-halfpatch)*cur_step + u_r-halfpatch);
        for (int x=0; x<patch_size; ++x, it += 3, ++q) {
          float search_pixel = wTL*it[0] + wTR*it[3] + wBL*it[3*cur_step] + wBR*it[3*cur_step+3];
          float res = search_pixel - (float)(affLL[0]*patch_[q] + affLL[1]) + mean_diff;
          Jres[0] -= res*dxp[q]; Jres[1] -= res*dyp[q]; Jres[2] -= res; } }
      float up[3]; for (int a=0;a<3;a++) up[a] = (Hinv.m[a][0]*Jres[0] + Hinv.m[a][1]*Jres[1]) + Hinv.m[a][2]*Jres[2];
      u += up[0]; v += up[1]; mean_diff += up[2];
      if (up[0]*up[0] + up[1]*up[1] < min_update_squared) { converged = true; break; }
    }
    px[0] = u; px[1] = v; return converged;
  }
  bool findMatchDirect(const MapPoint& pt, const Frame* cur, const SE3& curPose, AffLight curAff, int curKfIndex, double px_cur[2]) {
    int ref;
    if (kf.size() <= 2) {
      if (!backup) ref = 0; else if (curKfIndex == 0) ref = 1; else if (curKfIndex == 1) ref = 0; else return false;   // (NULL deref in the reference)
    } else ref = pt.host;
    double aff[2]; fromToVecExposure(kf[ref]->ab_exposure, cur->ab_exposure, kfAff[ref], curAff, aff); float affLL[2] = {(float)aff[0], (float)aff[1]};
    Vec3d ptWorld = pixelFrame2PointWorld(pt);
    Vec3d ptRef = xform(kfPose[ref].inverse(), ptWorld);
    Vec3d pixelRef = pointWorld2PixelFrame(kfPose[ref], ptWorld);
    double px[2] = {pixelRef.v[0], pixelRef.v[1]};
    if (!isInFrame(px[0], px[1], halfpatch_size_+2)) return false;
    double A[4]; getWarpMatrixAffine(px, ptRef, curPose.inverse()*kfPose[ref], A);
    int search_level = getBestSearchLevel(A, levels-1);
    warpAffine(A, kf[ref], px, search_level, halfpatch_size_+1, patch_with_border_);
    createPatchFromPatchWithBorder();
    double px_scaled[2] = {px_cur[0]/(1<<search_level), px_cur[1]/(1<<search_level)};
    bool success;
    if (pt.type == 1) {
      const float* d = &kf[ref]->dIp[0][3*(size_t)(int)(px[0] + px[1]*w[0])];
      double g0 = (double)d[1], g1 = (double)d[2]; double n = std::sqrt(g0*g0 + g1*g1); g0 /= n; g1 /= n;              // refGrad.normalize()
      double d0 = A[0]*g0 + A[1]*g1, d1 = A[2]*g0 + A[3]*g1; double n2 = std::sqrt(d0*d0 + d1*d1); d0 /= n2; d1 /= n2;
      float dir[2] = {(float)d0, (float)d1};
      success = align1D(cur->dIp[search_level].data(), search_level, dir, px_scaled, affLL);
    } else success = align2D(cur->dIp[search_level].data(), search_level, px_scaled, affLL);
    px_cur[0] = px_scaled[0]*(1<<search_level); px_cur[1] = px_scaled[1]*(1<<search_level);
    return success;
  }
  // reprojectMap (only_host < 0, backup false) / backprojectMap (only_host = index of `frame`, backup true).  curKfIndex = index of the target
  // frame in frameHessians_ or -1.  Returns the matches in cell visiting order.
  int run(const Frame* cur, const SE3& curPose, AffLight curAff, int curKfIndex, const std::vector<MapPoint>& pts, int only_host,
          const int* cell_order, int max_matches, int* out_pt, double* out_px) {
    const int cell_size = 25, ncols = (int)std::ceil((double)w[0]/cell_size), nrows = (int)std::ceil((double)h[0]/cell_size), ncells = ncols*nrows;
    struct Cand { int pt; double px[2]; float key; };
    std::vector<std::list<Cand>> cells(ncells);
    std::vector<int> order;                                                       // close_kfs: reverse index order, stable sort by distance (:125-131)
    if (only_host >= 0) order.push_back(only_host);
    else { std::vector<std::pair<int,double>> ck;
      for (int i=(int)kf.size()-1; i>=0; i--) { double d0 = curPose.t.v[0]-kfPose[i].t.v[0], d1 = curPose.t.v[1]-kfPose[i].t.v[1], d2 = curPose.t.v[2]-kfPose[i].t.v[2];
        ck.push_back({i, std::sqrt(d0*d0 + d1*d1 + d2*d2)}); }
      std::stable_sort(ck.begin(), ck.end(), [](const std::pair<int,double>& a, const std::pair<int,double>& b) { return a.second < b.second; });
      for (auto& c : ck) if (c.first != curKfIndex) order.push_back(c.first); }
    for (int hf : order) for (size_t i=0;i<pts.size();i++) { if (pts[i].host != hf) continue;          // reprojectPoint :600-616
      Vec3d ptWorld = pixelFrame2PointWorld(pts[i]); Vec3d pc = pointWorld2PixelFrame(curPose, ptWorld);
      if (!isInFrame(pc.v[0], pc.v[1], 8)) continue;
      int k = (int)(pc.v[1]/cell_size)*ncols + (int)(pc.v[0]/cell_size);
      const float* d = &kf[pts[i].host]->dIp[0][3*(size_t)(int)(pts[i].v*w[0] + pts[i].u)];
      cells[k].push_back(Cand{(int)i, {pc.v[0], pc.v[1]}, std::sqrt(d[1]*d[1] + d[2]*d[2])}); }
    int n_matches = 0, n_out = 0;
    for (int i=0;i<ncells;i++) {
      std::list<Cand>& cell = cells[cell_order ? cell_order[i] : i];
      cell.sort([](const Cand& a, const Cand& b) { return a.key < b.key; });        // weakest gradient first (sic), stable
      bool got = false;
      for (auto it = cell.begin(); it != cell.end(); ) {
        double px[2] = {it->px[0], it->px[1]};
        if (!findMatchDirect(pts[it->pt], cur, curPose, curAff, curKfIndex, px)) { it = cell.erase(it); continue; }
        out_pt[n_out] = it->pt; out_px[2*n_out] = px[0]; out_px[2*n_out+1] = px[1]; n_out++; got = true; break;
      }
      if (got) ++n_matches;
      if (n_matches > max_matches) break;
    }
    return n_out;
  }
};

} // namespace orc

using namespace orc;
extern "C" {
// pts5: (nP x 5) floats {u, v, idepth, host, type}.  kf_T7/kf_ab per keyframe, cur_* for the target frame.  cell_order may be NULL (identity).
int orc_reproject_map(int w, int h, int levels, const float K4[4], int nH, void** kf_frames, const double* kf_T7, const double* kf_ab,
                      void* cur_frame, const double cur_T7[7], const double cur_ab[2], int cur_kf_index, int nP, const float* pts5, int only_host, int backup,
                      const int* cell_order, int max_matches, int* out_pt, double* out_px) {
  Reprojector R; R.levels = levels; for (int l=0;l<levels;l++) { R.w[l] = w >> l; R.h[l] = h >> l; }
  std::memset(&R.K, 0, sizeof(R.K)); R.K.m[0][0] = (double)K4[0]; R.K.m[0][2] = (double)K4[2]; R.K.m[1][1] = (double)K4[1]; R.K.m[1][2] = (double)K4[3]; R.K.m[2][2] = 1.0;
  R.Ki = inverse3<double,Mat33d>(R.K); R.backup = backup != 0;
  for (int k=0;k<nH;k++) { R.kf.push_back((const Frame*)kf_frames[k]); SE3 s; s.q = Quat{kf_T7[7*k],kf_T7[7*k+1],kf_T7[7*k+2],kf_T7[7*k+3]}; s.t = Vec3d{{kf_T7[7*k+4],kf_T7[7*k+5],kf_T7[7*k+6]}};
    R.kfPose.push_back(s); AffLight a; a.a = kf_ab[2*k]; a.b = kf_ab[2*k+1]; R.kfAff.push_back(a); }
  SE3 c; c.q = Quat{cur_T7[0],cur_T7[1],cur_T7[2],cur_T7[3]}; c.t = Vec3d{{cur_T7[4],cur_T7[5],cur_T7[6]}}; AffLight ca; ca.a = cur_ab[0]; ca.b = cur_ab[1];
  std::vector<MapPoint> pts(nP); for (int i=0;i<nP;i++) pts[i] = MapPoint{pts5[5*i], pts5[5*i+1], pts5[5*i+2], (int)pts5[5*i+3], (int)pts5[5*i+4]};
  return R.run((const Frame*)cur_frame, c, ca, cur_kf_index, pts, only_host, cell_order, max_matches, out_pt, out_px);
}
}
