#pragma once
#include <cstdint>
#include <map>
#include <sstream>
#include <string>
#include <vector>
namespace XmlRpc {
class XmlRpcValue {
 public:
  enum Type { TypeInvalid, TypeBoolean, TypeInt, TypeDouble, TypeString, TypeDateTime, TypeBase64, TypeArray, TypeStruct };
  Type getType() const;
  int size() const;
  bool hasMember(const std::string &name) const;
  XmlRpcValue &operator[](int i);
  XmlRpcValue &operator[](const std::string &k);
  XmlRpcValue &operator[](const char *k);
  operator bool &();
  operator int &();
  operator double &();
  operator std::string &();
};
}  // namespace XmlRpc
namespace ros {
void init(int &argc, char **argv, const std::string &name, uint32_t options = 0);
void spin();
class Time {
 public:
  Time();
  double toSec() const;
  uint32_t sec, nsec;
};
class NodeHandle {
 public:
  NodeHandle(const std::string &ns = std::string());
  NodeHandle(const NodeHandle &);
  bool getParam(const std::string &key, std::string &s) const;
  bool getParam(const std::string &key, double &d) const;
  bool getParam(const std::string &key, bool &b) const;
  bool getParam(const std::string &key, XmlRpc::XmlRpcValue &v) const;
  bool searchParam(const std::string &key, std::string &result) const;
};
}  // namespace ros
#define ROS_FATAL_STREAM(x) do { std::ostringstream ros_ss__; ros_ss__ << x; } while (0)
#define ROS_ERROR_STREAM(x) do { std::ostringstream ros_ss__; ros_ss__ << x; } while (0)
#define NODELET_DEBUG(...) do { } while (0)
