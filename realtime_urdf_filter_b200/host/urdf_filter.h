// urdf_filter.h -- host-side mirror of the reference's public class
// realtime_urdf_filter::RealtimeURDFFilter (include/realtime_urdf_filter/urdf_filter.h:51-143):
// same method names, argument meaning, public data members and log-and-continue error behaviour;
// underneath, the GL objects are replaced by a ruf_context (include/ruf_b200.h).
//
// Built against host/ros_shim.h here (no ROS in this image); with ROS present the shim types are
// aliases of the real ones.
#pragma once
#include <string>
#include <vector>

#include "../../include/ruf_b200.h"
#include "ros_shim.h"
#include "urdf_model.h"

namespace realtime_urdf_filter {

using ruf_host::CameraInfoConstPtr;
using ruf_host::ImageConstPtr;
using ruf_host::Time;

class RealtimeURDFFilter {
 public:
  // constructor. reads the parameters (src/urdf_filter.cpp:43-118)
  RealtimeURDFFilter(ruf_host::NodeHandle &nh, int argc, char **argv);
  ~RealtimeURDFFilter();

  // loads URDF models (:127-197)
  void loadModels();
  // helper function to get current time (:200-205)
  double getTime();
  // callback function that gets ROS images and does everything (:270-330)
  void filter_callback(const ImageConstPtr &ros_depth_image, const CameraInfoConstPtr &camera_info);
  // does virtual rendering and filtering based on depth buffer and opengl proj. matrix (:207-267)
  void filter(unsigned char *buffer, double *glTf, int width, int height, Time timestamp = Time());
  // hands the depth buffer to the device path (:332-353; the copy itself happens inside ruf_filter)
  void textureBufferFromDepthBuffer(unsigned char *buffer, int size_in_bytes);
  // set up the device context + models (name kept: the tracker calls it, src/urdf_filtered_tracker.cpp:166)
  void initGL();
  void initFrameBufferObject();
  // compute Projection matrix from CameraInfo message (:459-501)
  void getProjectionMatrix(const CameraInfoConstPtr &current_caminfo, double *glTf);
  void render(const double *camera_projection_matrix, Time timestamp = Time());
  float *getMaskedDepth() { return masked_depth_; }

 public:
  ruf_host::NodeHandle nh_;
  ruf_host::TransformListener tf_;
  ruf_host::CameraPublisher depth_pub_;
  ruf_host::CameraPublisher mask_pub_;

  // rendering objects
  ruf_context *ctx_ = nullptr;          // replaces fbo_ / depth_texture_ / the shader program
  bool fbo_initialized_ = false;

  std::vector<ruf_host::URDFRenderer *> renderers_;
  std::vector<std::string> resource_roots_;   // where package:// URLs are looked up

  // parameters from launch file
  double camera_offset_t_[3] = {0, 0, 0};
  double camera_offset_q_[4] = {0, 0, 0, 1};
  std::string cam_frame_;
  std::string fixed_frame_;
  bool show_gui_ = false;
  bool need_mask_ = false;

  int width_ = 0;
  int height_ = 0;
  double camera_tx_ = 0;
  double camera_ty_ = 0;
  double far_plane_ = 8;
  double near_plane_ = 0.1;
  double depth_distance_threshold_ = 0;
  double filter_replace_value_ = 0;

  int argc_ = 0;
  char **argv_ = nullptr;

  // output from rendering
  float *masked_depth_ = nullptr;
  unsigned char *mask_ = nullptr;

  // bookkeeping the reference keeps in function-local statics (:210, :240-241)
  unsigned frames_ = 0;
  std::vector<double> timings_;
  std::string last_error_;              // text of the last ruf_* failure (logged as ROS_ERROR)

 private:
  int run_device(const void *depth_in, int enc, const double *P, Time stamp, void *depth_out, unsigned char *mask_out);
  void upload_models();
  const unsigned char *staged_buffer_ = nullptr;
  size_t model_parts_ = 0;
};

}  // namespace realtime_urdf_filter
