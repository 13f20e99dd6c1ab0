#pragma once
#include <sensor_msgs/Image.h>
namespace sensor_msgs {
struct CameraInfo {
  std_msgs::Header header;
  uint32_t height, width;
  std::array<double, 9> K, R;
  std::array<double, 12> P;
  typedef boost::shared_ptr<CameraInfo const> ConstPtr;
};
typedef boost::shared_ptr<CameraInfo const> CameraInfoConstPtr;
}  // namespace sensor_msgs
