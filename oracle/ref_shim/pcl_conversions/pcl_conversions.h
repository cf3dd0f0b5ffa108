// TEST INFRASTRUCTURE — the one conversion the filter sources call (see README.md).
#pragma once
#include "pcl/common/common_headers.h"
#include "ros/ros.h"
namespace pcl_conversions {
inline void toPCL(const ros::Time& stamp, uint64_t& pcl_stamp) { pcl_stamp = stamp.toNSec() / 1000ull; }
}  // namespace pcl_conversions
