// TEST INFRASTRUCTURE — value types of ROS1 that appear in the filter sources' signatures (see README.md).
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <iomanip>
#include <iostream>
#include <string>
#include <unordered_map>
#include <vector>
#include "boost/make_shared.hpp"

namespace ros {
struct Duration {
  int64_t ns = 0;
  int64_t toNSec() const { return ns; }
};
struct Time {
  uint64_t ns = 0;
  uint64_t toNSec() const { return ns; }
};
}  // namespace ros

namespace std_msgs {
struct Header {
  uint32_t seq = 0;
  ros::Time stamp;
  std::string frame_id;
};
}  // namespace std_msgs

namespace sensor_msgs {
struct Image {
  std_msgs::Header header;
  uint32_t height = 0, width = 0, step = 0;
  std::string encoding;
  std::vector<uint8_t> data;
};
typedef boost::shared_ptr<Image> ImagePtr;
typedef boost::shared_ptr<const Image> ImageConstPtr;
namespace image_encodings {
const std::string TYPE_8UC1 = "8UC1";
const std::string TYPE_8SC1 = "8SC1";
const std::string MONO8 = "mono8";
}  // namespace image_encodings
}  // namespace sensor_msgs
