#pragma once   // stand-in (ROS message header, absent; not used on the hot path)
#include "sensor_msgs/Image.h"
namespace sensor_msgs { struct PointCloud2 { std_msgs::Header header; std::vector<float> xyzi; /* decoded XYZI rows (pcl::fromROSMsg stand-in) */ }; typedef std::shared_ptr<const PointCloud2> PointCloud2ConstPtr; }
namespace nav_msgs { struct Odometry { std_msgs::Header header; struct { struct { struct { double x = 0, y = 0, z = 0, w = 1; } orientation; struct { double x = 0, y = 0, z = 0; } position; } pose; } pose; typedef std::shared_ptr<const Odometry> ConstPtr; }; typedef std::shared_ptr<const Odometry> OdometryConstPtr; }
