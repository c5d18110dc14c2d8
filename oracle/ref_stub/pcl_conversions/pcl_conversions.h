#pragma once
#include "../pcl/point_cloud.h"
#include "../sensor_msgs/PointCloud2.h"
namespace pcl {
// pcl::fromROSMsg for an XYZI cloud: the stand-in message carries the decoded points directly (the field decoding is ROS/PCL plumbing)
template <class T> void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<T>& out) { out.points.resize(msg.xyzi.size()/4);
  for (size_t i = 0; i < out.points.size(); i++) { out.points[i].x = msg.xyzi[4*i]; out.points[i].y = msg.xyzi[4*i+1]; out.points[i].z = msg.xyzi[4*i+2]; out.points[i].intensity = msg.xyzi[4*i+3]; } }
}
