#pragma once   // stand-in (ROS message header, absent; not used on the hot path)
#include "ros/ros.h"
#include <memory>
#include <vector>
namespace std_msgs { struct Header { ros::Time stamp; std::string frame_id; }; }
namespace sensor_msgs { struct Image { std_msgs::Header header; int width = 0, height = 0, step = 0; unsigned char is_bigendian = 0; std::string encoding; std::vector<unsigned char> data; }; typedef std::shared_ptr<const Image> ImageConstPtr; typedef std::shared_ptr<Image> ImagePtr;
namespace image_encodings { static const std::string MONO8 = "mono8"; } }
