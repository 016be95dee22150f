// Drop-in shim: lets the reference's main.cpp keep its `#include "model_grid_map.hpp"` (main.cpp:9-11) when
// -I include/welding_robot_b200/compat is placed before the reference's core/ directory.
// Everything lives in ../welding_robot.hpp (B200-native facade over include/wr_gpu.h).
#pragma once
#include "../welding_robot.hpp"
