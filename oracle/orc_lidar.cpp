// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header).  Pinned on oracle/_ref (the reference's own src/main.cpp compiled unmodified against PCL / ROS
// stand-ins): tests/test_ref_pin_lidar.py.
//
// orc_lidar.cpp — restatement of the LiDAR front-end of the ROS node (second half), all file:line in /root/reference/src/main.cpp:
//   projectPointCloud   :563-607   sweep -> N_SCAN x Horizon_SCAN range image (last point written to a cell wins) + organised cloud
//   groundRemoval       :609-655   per column, pairs of vertically adjacent cells within +-10 deg of horizontal are ground
//   labelComponents     :657-750   BFS over the 4-neighbourhood (columns wrap) joining cells whose range-difference angle exceeds segmentTheta
//   cloudSegmentation   :752-783   ground cells + segments of >= 30 cells (or >= 5 cells over >= 3 rows), raster order
//   lidarCloudHandler   :785-858   Rlc p + tlc, pinhole projection, image bounds, running pixel box, ground ratio -> addFeaturePoint
// float math like the reference's build: FullSystem.h:19 includes <math.h>, so atan2 / sqrt / sin / cos / abs of floats are the float overloads (atan2f ...).
// Behaviour kept on purpose: `nanPoint` is never initialised (main.cpp:84), so an empty cell holds PCL's default point (0,0,0, intensity 0) — not intensity -1 —
// and takes part in the ground test like a point at the sensor origin; row / column indices are size_t (a negative row fraction in (-1,0) lands in row 0).
#include <cmath>
#include <cfloat>
#include <cstdint>
#include <cstring>
#include <vector>
#include <algorithm>

namespace orc {

struct LidarSet { int N_SCAN = 64, Horizon_SCAN = 1800; float ang_res_x = 0.2f, ang_res_y = 0.427f, ang_bottom = 24.9f; int groundScanInd = 50; float sensorMountAngle = 0.0f;
                  int segmentValidPointNum = 5, segmentValidLineNum = 3; };               // main.cpp:103-122

struct LidarFrontEnd {
  LidarSet S; float segmentTheta, segmentAlphaX, segmentAlphaY;
  std::vector<float> range; std::vector<int> label; std::vector<int8_t> ground; std::vector<float> cloud;   // cloud: 4 floats per cell {x,y,z,intensity}
  std::vector<float> segmented;                                                                             // 4 floats per kept cell, raster order
  explicit LidarFrontEnd(const LidarSet& s) : S(s) {
    segmentTheta = 60.0/180.0*M_PI; segmentAlphaX = S.ang_res_x / 180.0 * M_PI; segmentAlphaY = S.ang_res_y / 180.0 * M_PI;
  }
  void reset() { size_t m = (size_t)S.N_SCAN*S.Horizon_SCAN; range.assign(m, FLT_MAX); label.assign(m, 0); ground.assign(m, 0); cloud.assign(4*m, 0.f); segmented.clear(); }

  void projectPointCloud(const float* xyzi, int n) {                                     // :563-607 (after pcl::removeNaNFromPointCloud :792)
    const int H = S.Horizon_SCAN;
    for (int i = 0; i < n; i++) {
      float x = xyzi[4*i], y = xyzi[4*i+1], z = xyzi[4*i+2];
      if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) continue;
      float verticalAngle = atan2f(z, sqrtf(x*x + y*y)) * 180 / M_PI;
      float rowf = (verticalAngle + S.ang_bottom) / S.ang_res_y;
      if (!(rowf > -1.0f)) continue;                                                     // size_t rowIdn: a value <= -1 converts to a huge index (>= N_SCAN)
      size_t rowIdn = (size_t)rowf;
      if (rowIdn >= (size_t)S.N_SCAN) continue;
      float horizonAngle = atan2f(x, y) * 180 / M_PI;
      double cold = -round((horizonAngle-90.0)/S.ang_res_x) + H/2;
      if (!(cold > -1.0)) continue;
      size_t columnIdn = (size_t)cold;
      if (columnIdn >= (size_t)H) columnIdn -= H;
      if (columnIdn >= (size_t)H) continue;
      float rng = sqrtf(x*x + y*y + z*z);
      if (rng < 0.1) continue;
      size_t index = columnIdn + rowIdn*H;
      range[index] = rng;
      cloud[4*index] = x; cloud[4*index+1] = y; cloud[4*index+2] = z; cloud[4*index+3] = (float)rowIdn + (float)columnIdn / 10000.0;
    }
  }
  void groundRemoval() {                                                                 // :609-655
    const int H = S.Horizon_SCAN;
    for (int j = 0; j < H; j++) for (int i = 0; i < S.groundScanInd; i++) {
      size_t lo = j + (size_t)i*H, up = j + (size_t)(i+1)*H;
      if (cloud[4*lo+3] == -1 || cloud[4*up+3] == -1) { ground[lo] = -1; continue; }
      float dX = cloud[4*up] - cloud[4*lo], dY = cloud[4*up+1] - cloud[4*lo+1], dZ = cloud[4*up+2] - cloud[4*lo+2];
      float angle = atan2f(dZ, sqrtf(dX*dX + dY*dY)) * 180 / M_PI;
      if (fabsf(angle - S.sensorMountAngle) <= 10) { ground[lo] = 1; ground[up] = 1; }
    }
    for (size_t k = 0; k < range.size(); k++) if (ground[k] == 1 || range[k] == FLT_MAX) label[k] = -1;
  }
  void cloudSegmentation() {                                                             // :657-783
    const int H = S.Horizon_SCAN, N = S.N_SCAN; int labelCount = 1;
    std::vector<int> qx((size_t)N*H), qy((size_t)N*H), ax((size_t)N*H), ay((size_t)N*H);
    static const int nb[4][2] = {{-1,0},{0,1},{0,-1},{1,0}};
    for (int i0 = 0; i0 < N; i0++) for (int j0 = 0; j0 < H; j0++) {
      if (label[(size_t)i0*H + j0] != 0) continue;
      std::vector<char> lineFlag(N, 0);
      qx[0] = i0; qy[0] = j0; int qs = 1, qstart = 0, qend = 1; ax[0] = i0; ay[0] = j0; int all = 1;
      while (qs > 0) {
        int fx = qx[qstart], fy = qy[qstart]; --qs; ++qstart;
        label[(size_t)fx*H + fy] = labelCount;
        for (int k = 0; k < 4; k++) {
          int tx = fx + nb[k][0], ty = fy + nb[k][1];
          if (tx < 0 || tx >= N) continue;
          if (ty < 0) ty = H-1;
          if (ty >= H) ty = 0;
          if (label[(size_t)tx*H + ty] != 0) continue;
          float d1 = std::max(range[(size_t)fx*H + fy], range[(size_t)tx*H + ty]), d2 = std::min(range[(size_t)fx*H + fy], range[(size_t)tx*H + ty]);
          float alpha = (nb[k][0] == 0) ? segmentAlphaX : segmentAlphaY;
          float angle = atan2f(d2*sinf(alpha), (d1 - d2*cosf(alpha)));
          if (angle > segmentTheta) { qx[qend] = tx; qy[qend] = ty; ++qs; ++qend; label[(size_t)tx*H + ty] = labelCount; lineFlag[tx] = 1; ax[all] = tx; ay[all] = ty; ++all; }
        }
      }
      bool feasible = false;
      if (all >= 30) feasible = true;
      else if (all >= S.segmentValidPointNum) { int lc = 0; for (int r = 0; r < N; r++) if (lineFlag[r]) lc++; if (lc >= S.segmentValidLineNum) feasible = true; }
      if (feasible) ++labelCount; else for (int k = 0; k < all; k++) label[(size_t)ax[k]*H + ay[k]] = 999999;
    }
    for (int i = 0; i < N; i++) for (int j = 0; j < H; j++) {
      size_t k = (size_t)i*H + j;
      if (label[k] > 0 || ground[k] == 1) {
        if (label[k] == 999999) continue;
        cloud[4*k+3] = (ground[k] == 1) ? -1.0f : 1.0f;
        segmented.insert(segmented.end(), &cloud[4*k], &cloud[4*k] + 4);
      }
    }
  }
  // the projection loop of lidarCloudHandler :806-849; lrud = FullSystem::left/right/up/down (in/out); returns rows {Ku,Kv,depth}
  int projectToImage(const double* R, const double* t, float fx, float fy, float cx, float cy, int w, int h, int* lrud, std::vector<double>& out, int* numGround, int* numAll) {
    int nG = 0, nA = 0; out.clear();
    for (size_t i = 0; i < segmented.size()/4; i++) {
      double p[3] = {segmented[4*i], segmented[4*i+1], segmented[4*i+2]}, tmp[3];
      for (int r = 0; r < 3; r++) tmp[r] = ((R[3*r]*p[0] + R[3*r+1]*p[1]) + R[3*r+2]*p[2]) + t[r];
      if (tmp[2] < 0.2) continue;
      float u = (float)(tmp[0] / tmp[2]), v = (float)(tmp[1] / tmp[2]);
      float Ku = u*fx + cx, Kv = v*fy + cy;
      if ((int)Ku < 4 || (int)Ku >= w-5 || (int)Kv < 4 || (int)Kv > h-4) continue;
      if (Ku < lrud[0]) lrud[0] = (int)Ku;
      if (Ku > lrud[1]) lrud[1] = (int)Ku;
      if (Kv < lrud[2]) lrud[2] = (int)Kv;
      if (Kv > lrud[3]) lrud[3] = (int)Kv;
      out.push_back((double)Ku); out.push_back((double)Kv); out.push_back(tmp[2]);
      nA++; if (segmented[4*i+3] < 0) nG++;
    }
    *numGround = nG; *numAll = nA;
    return (int)(out.size()/3);
  }
};

}  // namespace orc

extern "C" {
void* orc_lidar_create(int n_scan, int horizon, float ang_res_x, float ang_res_y, float ang_bottom, int groundScanInd) {
  orc::LidarSet S; S.N_SCAN = n_scan; S.Horizon_SCAN = horizon; S.ang_res_x = ang_res_x; S.ang_res_y = ang_res_y; S.ang_bottom = ang_bottom; S.groundScanInd = groundScanInd; return new orc::LidarFrontEnd(S); }
void orc_lidar_destroy(void* p) { delete (orc::LidarFrontEnd*)p; }
// one sweep through the whole handler; out rows {Ku,Kv,depth}; flags = {addFeaturePoint, numGround, numAll, size of segmentedCloud}; images optional
int orc_lidar_handler(void* p, const float* xyzi, int n, const double* Rlc9, const double* tlc3, const float* K4, int w, int h, int* lrud_io, double* out3, int cap, int* flags,
                      float* range_out, int* label_out, int8_t* ground_out) {
  orc::LidarFrontEnd* L = (orc::LidarFrontEnd*)p; L->reset(); L->projectPointCloud(xyzi, n); L->groundRemoval(); L->cloudSegmentation();
  size_t m = L->range.size();
  if (range_out) std::memcpy(range_out, L->range.data(), m*sizeof(float));
  if (label_out) std::memcpy(label_out, L->label.data(), m*sizeof(int));
  if (ground_out) std::memcpy(ground_out, L->ground.data(), m);
  std::vector<double> out; int nG, nA; int k = L->projectToImage(Rlc9, tlc3, K4[0], K4[1], K4[2], K4[3], w, h, lrud_io, out, &nG, &nA);
  flags[0] = (float(nG)/(float)nA > 0.8) ? 1 : 0; flags[1] = nG; flags[2] = nA; flags[3] = (int)(L->segmented.size()/4);
  if (k > cap) return -1;
  std::memcpy(out3, out.data(), out.size()*sizeof(double));
  return k;
}
}
