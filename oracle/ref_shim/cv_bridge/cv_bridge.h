// TEST INFRASTRUCTURE — cv_bridge::CvImage as a plain struct (see README.md).
#pragma once
#include "opencv2/core.hpp"
#include "ros/ros.h"

namespace cv_bridge {
class CvImage {
 public:
  std_msgs::Header header;
  std::string encoding;
  cv::Mat image;
  sensor_msgs::ImagePtr toImageMsg() const {
    sensor_msgs::ImagePtr m = boost::make_shared<sensor_msgs::Image>();
    m->header = header; m->encoding = encoding; m->height = (uint32_t)image.rows; m->width = (uint32_t)image.cols; m->step = (uint32_t)image.step;
    return m;
  }
};
typedef boost::shared_ptr<CvImage> CvImagePtr;
typedef boost::shared_ptr<const CvImage> CvImageConstPtr;
}  // namespace cv_bridge
