// TEST INFRASTRUCTURE — pcl::PointXYZI and pcl::PointCloud as a vector wrapper (see README.md).
#pragma once
#include <cstdint>
#include <string>
#include <vector>
#include "boost/make_shared.hpp"

namespace pcl {

struct PCLHeader {
  uint32_t seq = 0;
  uint64_t stamp = 0;
  std::string frame_id;
};

struct alignas(16) PointXYZI {  // PCL's constructor zero-initialises the coordinates and the intensity
  float x = 0.f, y = 0.f, z = 0.f, w_ = 1.f;
  float intensity = 0.f;
};

struct PointXY {
  float x = 0.f, y = 0.f;
};

template <typename PointT>
class PointCloud {
 public:
  typedef boost::shared_ptr<PointCloud<PointT>> Ptr;
  typedef boost::shared_ptr<const PointCloud<PointT>> ConstPtr;
  PCLHeader header;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;

  void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
  size_t size() const { return points.size(); }
  void resize(size_t n) { points.resize(n); width = (uint32_t)n; height = 1; }
  void clear() { points.clear(); width = height = 0; }
  bool empty() const { return points.empty(); }
  typename std::vector<PointT>::iterator begin() { return points.begin(); }
  typename std::vector<PointT>::iterator end() { return points.end(); }
  typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
  typename std::vector<PointT>::const_iterator end() const { return points.end(); }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
};

}  // namespace pcl
