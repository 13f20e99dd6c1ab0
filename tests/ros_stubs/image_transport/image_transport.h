#pragma once
#include <sensor_msgs/CameraInfo.h>
namespace image_transport {
class CameraSubscriber { public: CameraSubscriber(); uint32_t getNumPublishers() const; };
class CameraPublisher {
 public:
  CameraPublisher();
  uint32_t getNumSubscribers() const;
  void publish(const sensor_msgs::Image &image, const sensor_msgs::CameraInfo &info) const;
};
class ImageTransport {
 public:
  explicit ImageTransport(const ros::NodeHandle &nh);
  template <class T>
  CameraSubscriber subscribeCamera(const std::string &base_topic, uint32_t queue_size,
                                   void (T::*fp)(const sensor_msgs::ImageConstPtr &, const sensor_msgs::CameraInfoConstPtr &), T *obj);
  CameraPublisher advertiseCamera(const std::string &base_topic, uint32_t queue_size, bool latch = false);
};
}  // namespace image_transport
