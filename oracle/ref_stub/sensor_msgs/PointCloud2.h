#pragma once   // stand-in (ROS message header, absent; not used on the hot path)
#include "sensor_msgs/Image.h"
namespace sensor_msgs { struct PointCloud2 { std_msgs::Header header; }; typedef std::shared_ptr<const PointCloud2> PointCloud2ConstPtr; }
namespace nav_msgs { struct Odometry { std_msgs::Header header; }; typedef std::shared_ptr<const Odometry> OdometryConstPtr; }
