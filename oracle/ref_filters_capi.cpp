// TEST INFRASTRUCTURE — C surface over the reference's OWN filter classes, compiled from the sources where they lie under
// /root/reference (oracle/Makefile target `ref` -> oracle/_ref/libtbv_ref_filters.so; never part of the product, never
// committed).  It calls exactly what radarDriver::Process calls
// (cfear_radarodometry/src/cfear_radarodometry/radar_driver.cpp:48-60):
//   StructuredKStrongest filt(img, z_min, k, min_distance, range_res); filt.getPeaksFilteredPointCloud(cloud, false / true);
//   AzimuthCACFAR filter(window, pfa, guard, range_res, z_min, min_distance, 400.0); filter.getFilteredPointCloud(img, cloud);
// with the parameter types of radarDriver::Parameters (float members widened at the call,
// cfear_radarodometry/include/cfear_radarodometry/radar_driver.h:40-45).
#include <cstdint>
#include <cstring>

#include "cfear_radarodometry/cfar.h"
#include "cfear_radarodometry/radar_filters.h"
#include "cfear_radarodometry/statistics.h"

namespace {
cv_bridge::CvImagePtr wrap(const uint8_t* img, int n_az, int n_range, long row_stride) {
  cv_bridge::CvImagePtr p = boost::make_shared<cv_bridge::CvImage>();
  p->image = cv::Mat::zeros(n_az, n_range, CV_8UC1);
  for (int a = 0; a < n_az; a++) std::memcpy(p->image.data + (size_t)a * p->image.step, img + (size_t)a * (size_t)row_stride, (size_t)n_range);
  p->encoding = "8UC1";
  return p;
}
int emit(const pcl::PointCloud<pcl::PointXYZI>::Ptr& c, float* x, float* y, float* intensity, int cap) {
  const int n = (int)c->size();
  for (int i = 0; i < n && i < cap; i++) {
    x[i] = c->points[i].x; y[i] = c->points[i].y; intensity[i] = c->points[i].intensity;
  }
  return n;
}
}  // namespace

extern "C" {

// Returns the number of points of the requested cloud (may exceed cap; only cap are written).  `peaks` != 0 selects the
// AxialNonMaxSupress'ed cloud.  z_min / min_distance / range_res are the float parameters of the driver.
int tbv_ref_kstrongest(const uint8_t* img, int n_az, int n_range, long row_stride, float z_min, int k, float min_distance, float range_res, int peaks,
                       float* x, float* y, float* intensity, int cap) {
  cv_bridge::CvImagePtr im = wrap(img, n_az, n_range, row_stride);
  CFEAR_Radarodometry::StructuredKStrongest filt(im, z_min, k, min_distance, range_res);
  pcl::PointCloud<pcl::PointXYZI>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZI>());
  filt.getPeaksFilteredPointCloud(cloud, peaks != 0);
  return emit(cloud, x, y, intensity, cap);
}

// Both clouds from ONE filter object, in the driver's order (filtered first, then peaks).
int tbv_ref_kstrongest_both(const uint8_t* img, int n_az, int n_range, long row_stride, float z_min, int k, float min_distance, float range_res,
                            float* x, float* y, float* intensity, int cap, int* n_filtered, float* px, float* py, float* pintensity, int pcap, int* n_peaks) {
  cv_bridge::CvImagePtr im = wrap(img, n_az, n_range, row_stride);
  CFEAR_Radarodometry::StructuredKStrongest filt(im, z_min, k, min_distance, range_res);
  pcl::PointCloud<pcl::PointXYZI>::Ptr c0(new pcl::PointCloud<pcl::PointXYZI>()), c1(new pcl::PointCloud<pcl::PointXYZI>());
  filt.getPeaksFilteredPointCloud(c0, false);
  filt.getPeaksFilteredPointCloud(c1, true);
  *n_filtered = emit(c0, x, y, intensity, cap);
  *n_peaks = emit(c1, px, py, pintensity, pcap);
  return 0;
}

int tbv_ref_cacfar(const uint8_t* img, int n_az, int n_range, long row_stride, int window, float false_alarm_rate, int guard, float range_res, float z_min,
                   float min_distance, double max_distance, float* x, float* y, float* intensity, int cap) {
  cv_bridge::CvImagePtr im = wrap(img, n_az, n_range, row_stride);
  AzimuthCACFAR filter(window, false_alarm_rate, guard, range_res, z_min, min_distance, max_distance);
  pcl::PointCloud<pcl::PointXYZI>::Ptr cloud(new pcl::PointCloud<pcl::PointXYZI>());
  filter.getFilteredPointCloud(im, cloud);
  return emit(cloud, x, y, intensity, cap);
}

// Wall-clock helper for bench.py's `--impl reference` "Filtering" stage: runs the driver's k-strongest sequence (constructor +
// both clouds) over `n_scans` scans stored back to back; returns the total number of points so the work cannot be elided.
long tbv_ref_kstrongest_many(const uint8_t* imgs, int n_scans, int n_az, int n_range, float z_min, int k, float min_distance, float range_res) {
  long total = 0;
  for (int s = 0; s < n_scans; s++) {
    cv_bridge::CvImagePtr im = wrap(imgs + (size_t)s * (size_t)n_az * (size_t)n_range, n_az, n_range, n_range);
    CFEAR_Radarodometry::StructuredKStrongest filt(im, z_min, k, min_distance, range_res);
    pcl::PointCloud<pcl::PointXYZI>::Ptr c0(new pcl::PointCloud<pcl::PointXYZI>()), c1(new pcl::PointCloud<pcl::PointXYZI>());
    filt.getPeaksFilteredPointCloud(c0, false);
    filt.getPeaksFilteredPointCloud(c1, true);
    total += (long)c0->size() + (long)c1->size();
  }
  return total;
}

// CFEAR_Radarodometry::statistics (statistics.cpp:10-51), the reference's timing table: Document every (name, value), return
// GetStatistics().  names: n zero-terminated strings back to back.  Returns the length written (without the terminator).
int tbv_ref_statistics(const char* names, const double* values, int n, char* out, int cap) {
  CFEAR_Radarodometry::statistics st;
  const char* p = names;
  for (int i = 0; i < n; i++) {
    st.Document(std::string(p), values[i]);
    p += std::strlen(p) + 1;
  }
  const std::string s = st.GetStatistics();
  const int len = (int)s.size() < cap - 1 ? (int)s.size() : cap - 1;
  std::memcpy(out, s.data(), len);
  out[len] = 0;
  return len;
}

}  // extern "C"
