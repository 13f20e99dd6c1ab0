// realtime_urdf_filter_node.cpp -- the stand-alone node, executable `realtime_urdf_filter`
// (replaces the reference's src/realtime_urdf_filter.cpp:36-53: same node name, private node handle,
// construct the filter, spin, std::runtime_error -> ROS_FATAL).
#include <stdexcept>
#include <string>

#include <ros/ros.h>

#include "ros_bridge.h"

int main(int argc, char **argv)
{
  // set up ROS
  ros::init(argc, argv, "realtime_urdf_filter");
  ros::NodeHandle nh("~");

  // create the filter (parameters are read here; the device context and the models come up with the first frame,
  // like the reference's lazy initGL) and subscribe to ROS
  realtime_urdf_filter::RosBridge bridge(nh, argc, argv);

  try {
    ros::spin();
  } catch (const std::runtime_error &e) {          // no CUDA device / no models: src/urdf_filter.cpp:415,427
    ROS_FATAL_STREAM(std::string(e.what()));
  }
  return 0;
}
