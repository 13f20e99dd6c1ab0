// urdf_filter.h -- host-side mirror of the reference's public class
// realtime_urdf_filter::RealtimeURDFFilter (include/realtime_urdf_filter/urdf_filter.h:51-143):
// same method names, argument meaning, public data members and log-and-continue error behaviour;
// underneath, the GL objects are replaced by a ruf_context (include/ruf_b200.h).
//
// Built against host/ros_shim.h here (no ROS in this image); with ROS present the shim types are
// aliases of the real ones.
#pragma once
#include <cstdint>
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
  bool packed_mask_ = false;            // `packed_mask_readback`: RUF_MASK_BITS on the wire, expanded to MONO8 on the host
  std::vector<unsigned char> mask_bits_;
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
  // page-locked staging of the facade (ruf_host_alloc): ruf_filter runs its single-frame graph -- three kernels that read
  // and write the host buffers themselves -- only for pinned buffers, and a ROS message's data is a pageable vector.  One
  // memcpy into / out of these is cheaper than the pageable copies the runtime would stage (`pinned_staging` parameter,
  // default true; buffers that are pinned already are passed through).
  bool pinned_staging_ = true;
  unsigned char *pin_in_ = nullptr, *pin_out_ = nullptr;
  size_t pin_bytes_ = 0;
};

// ---- the second caller of filter(): OpenNITrackerLoopback::runOnce (src/urdf_filtered_tracker.cpp:180-252) ----
// The tracker is out of scope (OpenNI/NITE + a Kinect), but the conversions it wraps around
// filter(buffer, glTf, XRes, YRes) + getMaskedDepth() are part of how the path is used, so they are kept here
// for a maintainer who re-wires it:
//   tracker_depth_to_buffer   :201-207  buffer[x + y*XRes] = depthMap(XRes - x - 1, y) * 0.001   (mirrored in x;
//                                        the product is taken in double and rounded to float once)
//   tracker_projection        :209-236  the hard-coded Kinect intrinsics 585.260 / 585.028 / 317.387 / 239.264
//                                        (held as float P[12], like the reference) -> glTf[16]
//   tracker_masked_depth_to_mm :243-249 depthMap_(x, y) = XnDepthPixel(masked_depth[..] * 1000): float product,
//                                        TRUNCATED (not rounded like filter_callback's convertTo) and NOT mirrored
//                                        back; values outside [0, 65535] and NaN are undefined in the reference,
//                                        here NaN and negatives give 0 and larger values saturate
void tracker_depth_to_buffer(const uint16_t *depth_mm, int xres, int yres, float *buffer);
void tracker_projection(int xres, int yres, double *glTf);
void tracker_masked_depth_to_mm(const float *masked_depth, int xres, int yres, uint16_t *depth_mm);

}  // namespace realtime_urdf_filter
