#pragma once
#include <vector>
#include <memory>
#include "point_types.h"
namespace pcl {
template <class T> struct PointCloud {
  typedef std::shared_ptr<PointCloud<T> > Ptr;
  std::vector<T> points;
  void push_back(const T& p) { points.push_back(p); }
  void clear() { points.clear(); }
  size_t size() const { return points.size(); }
  PointCloud& operator+=(const PointCloud& o) { points.insert(points.end(), o.points.begin(), o.points.end()); return *this; }
};
}
