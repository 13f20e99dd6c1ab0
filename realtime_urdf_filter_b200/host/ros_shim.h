// ros_shim.h -- the handful of ROS / tf / sensor_msgs types the RealtimeURDFFilter facade touches,
// re-stated without ROS so that the host side builds and is testable in an image that has no ROS.
// A ROS build replaces this header by the real ones; names, members and call
// semantics follow the real types as used by the reference (src/urdf_filter.cpp, src/urdf_renderer.cpp).
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace ruf_host {

// ---- ros::Time -------------------------------------------------------------------------------
struct Time {
  double sec = 0.0;       // ros::Time() == 0 means "latest" for tf lookups
  Time() = default;
  explicit Time(double s) : sec(s) {}
};

// ---- std_msgs/Header, sensor_msgs/Image, sensor_msgs/CameraInfo ------------------------------
struct Header {
  uint32_t seq = 0;
  Time stamp;
  std::string frame_id;
};
struct Image {
  Header header;
  uint32_t height = 0, width = 0;
  std::string encoding;          // "32FC1", "16UC1", "mono8"
  uint8_t is_bigendian = 0;
  uint32_t step = 0;             // row length in bytes
  std::vector<uint8_t> data;
};
struct CameraInfo {
  Header header;
  uint32_t height = 0, width = 0;
  double P[12] = {0};            // 3x4 row-major projection matrix
};
typedef std::shared_ptr<const Image> ImageConstPtr;
typedef std::shared_ptr<const CameraInfo> CameraInfoConstPtr;

// ---- tf ---------------------------------------------------------------------------------------
struct TransformException : public std::runtime_error {
  explicit TransformException(const std::string &w) : std::runtime_error(w) {}
};
// rotation (x,y,z,w) + origin, as tf::StampedTransform::getRotation()/getOrigin() return them
struct StampedTransform {
  double q[4] = {0, 0, 0, 1};
  double t[3] = {0, 0, 0};
  Time stamp;
};
// Minimal tf::TransformListener: a flat table frame -> pose in one common root.  lookupTransform
// (target, source) returns the transform that maps source-frame coordinates into the target frame,
// exactly the convention of tf::Transformer::lookupTransform.
class TransformListener {
 public:
  // pose of `frame` in the root frame: root_T_frame
  void setTransform(const std::string &frame, const double q[4], const double t[3]);
  void erase(const std::string &frame) { frames_.erase(norm(frame)); }
  void lookupTransform(const std::string &target_frame, const std::string &source_frame, const Time &time,
                       StampedTransform &out) const;   // throws TransformException
  size_t size() const { return frames_.size(); }
  mutable uint64_t lookups = 0;   // instrumentation: tf lookups per frame are the reference's host hot loop

 private:
  struct Pose { double R[3][3]; double t[3]; };
  static std::string norm(const std::string &f) { return (!f.empty() && f[0] == '/') ? f.substr(1) : f; }
  std::map<std::string, Pose> frames_;
};

// ---- parameters (ros::NodeHandle::getParam / param on the private namespace) -----------------
struct ModelParam {               // one entry of the `models` array, src/urdf_filter.cpp:127-197
  std::string model;              // name of the parameter holding the URDF XML
  std::string tf_prefix;
  std::string geometry_type;      // "" / "visual" / "collision"
  double scale = 1.0;
  std::vector<std::string> ignore;
};
class NodeHandle {
 public:
  void setParam(const std::string &k, const std::string &v) { s_[k] = v; }
  void setParam(const std::string &k, double v) { d_[k] = v; }
  void setParam(const std::string &k, bool v) { b_[k] = v; }
  void setCameraOffset(const double t[3], const double q[4]);
  void addModel(const ModelParam &m) { models_.push_back(m); has_models_ = true; }
  bool getParam(const std::string &k, std::string &v) const;
  bool getParam(const std::string &k, double &v) const;
  bool getParam(const std::string &k, bool &v) const;
  bool getCameraOffset(double t[3], double q[4]) const;
  bool getModels(std::vector<ModelParam> &m) const { m = models_; return has_models_; }

 private:
  std::map<std::string, std::string> s_;
  std::map<std::string, double> d_;
  std::map<std::string, bool> b_;
  std::vector<ModelParam> models_;
  bool has_models_ = false, has_offset_ = false;
  double off_t_[3] = {0, 0, 0}, off_q_[4] = {0, 0, 0, 1};
};

// ---- image_transport::CameraPublisher -----------------------------------------------------------
// Keeps the last published pair; getNumSubscribers() is settable so that tests can exercise the
// need_mask_ logic (src/urdf_filter.cpp:226-230, :306, :321).
class CameraPublisher {
 public:
  int subscribers = 1;
  uint64_t published = 0;
  Image last_image;
  CameraInfo last_info;
  int getNumSubscribers() const { return subscribers; }
  void publish(const Image &img, const CameraInfo &info) { last_image = img; last_info = info; ++published; }
};

// ---- logging (rosconsole) ------------------------------------------------------------------------
enum LogLevel { LOG_DEBUG = 0, LOG_INFO, LOG_ERROR, LOG_FATAL };
typedef void (*LogSink)(LogLevel, const char *);
void set_log_sink(LogSink s);
void logf(LogLevel lvl, const char *fmt, ...);

}  // namespace ruf_host
