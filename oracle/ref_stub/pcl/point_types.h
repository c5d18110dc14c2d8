// stand-in for PCL (absent from the image): the point type and container main.cpp's LiDAR front-end uses
#pragma once
namespace pcl { struct PointXYZI { float x, y, z, intensity; PointXYZI() : x(0), y(0), z(0), intensity(0) {} }; }
