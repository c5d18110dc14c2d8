#pragma once
#include <cmath>
#include "../point_cloud.h"
namespace pcl {
// pcl::removeNaNFromPointCloud: keeps the points whose x, y, z are all finite (dense output), index map returned
template <class T> void removeNaNFromPointCloud(const PointCloud<T>& in, PointCloud<T>& out, std::vector<int>& index) {
  std::vector<T> keep; index.clear();
  for (size_t i = 0; i < in.points.size(); i++) { const T& p = in.points[i]; if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue; keep.push_back(p); index.push_back((int)i); }
  out.points.swap(keep);
}
}
