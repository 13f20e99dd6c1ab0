// realtime_urdf_filter_nodelet.cpp -- replaces the reference's src/realtime_urdf_filter_nodelet.cpp:35-75.
#include "realtime_urdf_filter_nodelet.h"

#include <pluginlib/class_list_macros.h>

// watch the capitalization carefully: this is the name nodelet_plugins.xml and the launch files use
PLUGINLIB_EXPORT_CLASS(realtime_urdf_filter::RealtimeURDFFilterNodelet, nodelet::Nodelet);

namespace realtime_urdf_filter {

RealtimeURDFFilterNodelet::RealtimeURDFFilterNodelet() : args_(), argv_() {}

void RealtimeURDFFilterNodelet::onInit()
{
  NODELET_DEBUG("Initializing nodelet...");
  // c-style argc / argv of the nodelet's own arguments (std::string storage outlives the filter: no manual new[])
  args_ = this->getMyArgv();
  argv_.clear();
  for (std::string &a : args_) argv_.push_back(&a[0]);

  // the private node handle has a single-threaded callback queue: frames are serialised like under ros::spin(),
  // which is what a ruf_context needs (one stream per context, not re-entrant)
  ros::NodeHandle nh = this->getPrivateNodeHandle();

  // Create the filter.  Exceptions of the first frame (no CUDA device, no models) propagate into the nodelet manager,
  // as the reference's do.
  filter_.reset(new RosBridge(nh, (int)argv_.size(), argv_.empty() ? nullptr : argv_.data()));
}

}  // namespace realtime_urdf_filter
