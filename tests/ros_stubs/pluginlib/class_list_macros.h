#pragma once
#include <type_traits>
#define PLUGINLIB_EXPORT_CLASS(class_type, base_class_type) \
  static_assert(std::is_base_of<base_class_type, class_type>::value, "plugin class must derive from its base class")
