// Shim used ONLY when compiling the reference's tape core into oracle/_ref/.
// tractor/src/core/profiler.cpp:9,102 needs ROS_INFO_STREAM and nothing else;
// the once-per-second profiler table is discarded.
#pragma once
#include <iostream>
#include <sstream>
#define ROS_INFO_STREAM(x) do { } while (0)
