// realtime_urdf_filter_nodelet.h -- the nodelet flavour (replaces the reference's
// include/realtime_urdf_filter/realtime_urdf_filter_nodelet.h:38-50): same class name and namespace, so that
// plugins/nodelet_plugins.xml and existing launch files (`load realtime_urdf_filter/RealtimeURDFFilterNodelet ...`,
// launch/realtime_urdf_filter.launch:5) keep working.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include <nodelet/nodelet.h>

#include "ros_bridge.h"

namespace realtime_urdf_filter {

class RealtimeURDFFilterNodelet : public nodelet::Nodelet {
 public:
  RealtimeURDFFilterNodelet();
  virtual void onInit();

 private:
  std::vector<std::string> args_;         // getMyArgv(); the reference turns them into argc/argv for GLUT,
  std::vector<char *> argv_;              // here they are only handed through (no GL, no X display needed)
  std::shared_ptr<RosBridge> filter_;
};

}  // namespace realtime_urdf_filter
