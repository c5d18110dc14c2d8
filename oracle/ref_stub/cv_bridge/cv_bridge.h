#pragma once   // stand-in (absent; not used on the hot path)
#include "opencv2/core/core.hpp"
#include "sensor_msgs/Image.h"
namespace cv_bridge { struct CvImage { cv::Mat image; }; typedef std::shared_ptr<CvImage> CvImagePtr; typedef std::shared_ptr<const CvImage> CvImageConstPtr;
template <class M> inline CvImagePtr toCvCopy(const M&, const std::string&) { return std::make_shared<CvImage>(); } }
