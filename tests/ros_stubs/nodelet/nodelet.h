#pragma once
#include <ros/ros.h>
namespace nodelet {
class Nodelet {
 public:
  Nodelet();
  virtual ~Nodelet();
 protected:
  const std::vector<std::string> &getMyArgv() const;
  ros::NodeHandle &getPrivateNodeHandle() const;
 private:
  virtual void onInit() = 0;
};
}  // namespace nodelet
