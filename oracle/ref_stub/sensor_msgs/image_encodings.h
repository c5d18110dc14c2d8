#pragma once   // stand-in (ROS message header, absent; not used on the hot path)
#include "ros/ros.h"
