// Drop-in shim: lets the reference's main.cpp keep its `#include "core/BSplineBasic.h"` / "BSplineBasic.h" (main.cpp:6) when
// -I include/welding_robot_b200/compat is placed before the reference's core/ directory.
// BS_Basic lives in ../welding_robot.hpp (B200-native facade over include/wr_gpu.h).
#pragma once
#include "../welding_robot.hpp"
