#pragma once
#include <stdexcept>
#include <ros/ros.h>
namespace tf {
class Quaternion { public: const double &x() const; const double &y() const; const double &z() const; const double &w() const; };
class Vector3 { public: const double &x() const; const double &y() const; const double &z() const; };
class StampedTransform { public: StampedTransform(); Quaternion getRotation() const; const Vector3 &getOrigin() const; ros::Time stamp_; };
class TransformException : public std::runtime_error { public: TransformException(const std::string &e) : std::runtime_error(e) {} };
class TransformListener {
 public:
  TransformListener();
  void lookupTransform(const std::string &target_frame, const std::string &source_frame, const ros::Time &time,
                       StampedTransform &transform) const;
};
}  // namespace tf
