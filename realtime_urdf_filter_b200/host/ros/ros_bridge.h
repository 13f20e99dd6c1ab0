// ros_bridge.h -- the ROS 1 side of the drop-in: what stays C++/ROS around the CUDA path.
//
// Replaces the ROS plumbing of realtime_urdf_filter::RealtimeURDFFilter (the reference's src/urdf_filter.cpp):
//   :58-111   required / optional rosparams on the private node handle          -> read_params()
//   :127-186  the `models` array (model key with searchParam fallback, tf_prefix,
//             geometry_type, scale, ignore as string or list)                   -> read_params()
//   :114-117  subscribeCamera("input_depth", 10) / advertiseCamera(output_depth, output_mask)
//   :270-330  filter_callback: decode, filter, publish in the input's encoding   -> callback()
//   :522 and src/urdf_renderer.cpp:173-190  the per-frame tf lookups             -> refresh_tf()
// The class below owns the real ROS objects and drives host/urdf_filter.h's RealtimeURDFFilter (same public
// members as the reference's class), which calls libruf_b200.so through include/ruf_b200.h.
//
// Needs roscpp, image_transport, tf, sensor_msgs; this image has no ROS, so the file is only checked with
// `g++ -fsyntax-only` against tests/ros_stubs/ (tests/test_ros_sources.py).
#pragma once
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <ros/ros.h>
#include <image_transport/image_transport.h>
#include <sensor_msgs/CameraInfo.h>
#include <sensor_msgs/Image.h>
#include <tf/transform_listener.h>

#include "../urdf_filter.h"

namespace realtime_urdf_filter {

class RosBridge {
 public:
  RosBridge(ros::NodeHandle &nh, int argc, char **argv)
      : nh_(nh), it_(nh), filter_(read_params(nh, shim_), argc, argv)
  {
    // setup publishers / subscribers (src/urdf_filter.cpp:114-117: same topic names, queue 10)
    depth_sub_ = it_.subscribeCamera("input_depth", 10, &RosBridge::callback, this);
    depth_pub_ = it_.advertiseCamera("output_depth", 10);
    mask_pub_ = it_.advertiseCamera("output_mask", 10);
  }
  RealtimeURDFFilter &filter() { return filter_; }

 private:
  // rosparams -> the facade's parameter table; missing ones stay missing so that the facade logs FATAL like :58-103
  static ruf_host::NodeHandle &read_params(ros::NodeHandle &nh, ruf_host::NodeHandle &out)
  {
    std::string s;
    double d;
    bool b;
    if (nh.getParam("fixed_frame", s)) out.setParam("fixed_frame", s);
    if (nh.getParam("camera_frame", s)) out.setParam("camera_frame", s);
    if (nh.getParam("depth_distance_threshold", d)) out.setParam("depth_distance_threshold", d);
    if (nh.getParam("filter_replace_value", d)) out.setParam("filter_replace_value", d);
    if (nh.getParam("show_gui", b)) out.setParam("show_gui", b);
    XmlRpc::XmlRpcValue v;
    if (nh.getParam("camera_offset", v)) {                                     // :73-96
      double t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1};
      for (int i = 0; i < 3; ++i) t[i] = (double)v["translation"][i];
      for (int i = 0; i < 4; ++i) q[i] = (double)v["rotation"][i];
      out.setCameraOffset(t, q);
    }
    if (nh.getParam("models", v) && v.getType() == XmlRpc::XmlRpcValue::TypeArray) {   // :127-186
      for (int i = 0; i < v.size(); ++i) {
        XmlRpc::XmlRpcValue &e = v[i];
        ruf_host::ModelParam m;
        m.model = (std::string)e["model"];
        std::string found = m.model, xml;
        if (!nh.getParam(m.model, xml)) {                                      // searchParam fallback, :145-158
          if (nh.searchParam(m.model, found)) nh.getParam(found, xml);
        }
        if (!xml.empty()) out.setParam(m.model, xml);
        m.tf_prefix = (std::string)e["tf_prefix"];
        if (e.hasMember("geometry_type")) m.geometry_type = (std::string)e["geometry_type"];
        if (e.hasMember("scale")) m.scale = (double)e["scale"];                // default 1.0, :166
        if (e.hasMember("ignore")) {                                           // a string or a list of link names, :168-186
          XmlRpc::XmlRpcValue &ig = e["ignore"];
          if (ig.getType() == XmlRpc::XmlRpcValue::TypeString) m.ignore.push_back((std::string)ig);
          else for (int k = 0; k < ig.size(); ++k) m.ignore.push_back((std::string)ig[k]);
        }
        out.addModel(m);
      }
    }
    return out;
  }

  // Every frame the facade will ask its table for, looked up in the real tf tree at the image's stamp.  A failed
  // lookup erases the entry, so the facade behaves like the reference: camera missing -> log + stale outputs
  // (:531-534); link missing -> the previous link's transform is reused (src/urdf_renderer.cpp:175-188).
  void refresh_tf(const ros::Time &stamp)
  {
    const double q0[4] = {0, 0, 0, 1}, t0[3] = {0, 0, 0};
    filter_.tf_.setTransform(filter_.fixed_frame_, q0, t0);
    std::vector<std::string> frames(1, filter_.cam_frame_);
    for (ruf_host::URDFRenderer *r : filter_.renderers_)
      for (const ruf_host::RenderablePart &p : r->parts())
        if (frames.back() != p.name) frames.push_back(p.name);
    for (const std::string &f : frames) {
      try {
        tf::StampedTransform t;
        tf_.lookupTransform(filter_.fixed_frame_, f, stamp, t);
        const tf::Quaternion q = t.getRotation();
        const tf::Vector3 o = t.getOrigin();
        const double qq[4] = {q.x(), q.y(), q.z(), q.w()}, tt[3] = {o.x(), o.y(), o.z()};
        filter_.tf_.setTransform(f, qq, tt);
      } catch (const tf::TransformException &) {
        filter_.tf_.erase(f);
      }
    }
  }

  static void to_ros(const ruf_host::Image &in, const sensor_msgs::Image &like, sensor_msgs::Image &out)
  {
    out.header = like.header;                                                  // keeps the input's header, :315
    out.height = in.height; out.width = in.width; out.encoding = in.encoding; out.step = in.step;
    out.is_bigendian = in.is_bigendian;
    out.data = in.data;
  }

  void callback(const sensor_msgs::ImageConstPtr &img, const sensor_msgs::CameraInfoConstPtr &info)
  {
    refresh_tf(img->header.stamp);
    auto im = std::make_shared<ruf_host::Image>();
    im->header.stamp = ruf_host::Time(img->header.stamp.toSec());
    im->header.frame_id = img->header.frame_id;
    im->height = img->height; im->width = img->width; im->encoding = img->encoding; im->step = img->step;
    im->is_bigendian = img->is_bigendian;
    im->data = img->data;
    auto ci = std::make_shared<ruf_host::CameraInfo>();
    ci->height = info->height; ci->width = info->width;
    for (int i = 0; i < 12; ++i) ci->P[i] = info->P[i];
    filter_.depth_pub_.subscribers = (int)depth_pub_.getNumSubscribers();      // :306
    filter_.mask_pub_.subscribers = (int)mask_pub_.getNumSubscribers();        // need_mask_, :226-230
    const uint64_t nd = filter_.depth_pub_.published, nm = filter_.mask_pub_.published;
    filter_.filter_callback(im, ci);                                           // throws std::runtime_error like initGL, :415,427
    sensor_msgs::Image out;
    if (filter_.depth_pub_.published != nd) { to_ros(filter_.depth_pub_.last_image, *img, out); depth_pub_.publish(out, *info); }
    if (filter_.mask_pub_.published != nm) { to_ros(filter_.mask_pub_.last_image, *img, out); mask_pub_.publish(out, *info); }
  }

  ros::NodeHandle nh_;
  image_transport::ImageTransport it_;
  tf::TransformListener tf_;
  ruf_host::NodeHandle shim_;
  RealtimeURDFFilter filter_;
  image_transport::CameraSubscriber depth_sub_;
  image_transport::CameraPublisher depth_pub_, mask_pub_;
};

}  // namespace realtime_urdf_filter
