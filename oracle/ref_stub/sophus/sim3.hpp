#pragma once
#include "sophus/se3.hpp"
namespace Sophus { class Sim3d { public: SE3d se3; double scale = 1; }; }
