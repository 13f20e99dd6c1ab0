#pragma once
#include <array>
#include <boost/shared_ptr.hpp>
#include <ros/ros.h>
namespace std_msgs { struct Header { uint32_t seq; ros::Time stamp; std::string frame_id; }; }
namespace sensor_msgs {
struct Image {
  std_msgs::Header header;
  uint32_t height, width;
  std::string encoding;
  uint8_t is_bigendian;
  uint32_t step;
  std::vector<uint8_t> data;
  typedef boost::shared_ptr<Image const> ConstPtr;
};
typedef boost::shared_ptr<Image const> ImageConstPtr;
}  // namespace sensor_msgs
