// facade_c.cpp -- flat C hooks over the C++ RealtimeURDFFilter facade so that the Python tests can
// drive it the way a ROS node would (set params, feed TF, deliver Image + CameraInfo, read what was
// published).  Not part of the drop-in boundary (that is include/ruf_b200.h).
#include <cstring>
#include <exception>
#include <sstream>
#include <string>

#include "urdf_filter.h"

using namespace ruf_host;
using realtime_urdf_filter::RealtimeURDFFilter;

namespace {
struct Facade {
  NodeHandle nh;
  RealtimeURDFFilter *f = nullptr;
  std::string log, error;
};
thread_local Facade *g_logging = nullptr;
void sink(LogLevel lvl, const char *msg)
{
  if (!g_logging) return;
  static const char *names[] = {"DEBUG", "INFO", "ERROR", "FATAL"};
  g_logging->log += std::string("[") + names[lvl] + "] " + msg + "\n";
}
struct LogScope {
  explicit LogScope(Facade *f) { g_logging = f; set_log_sink(sink); }
  ~LogScope() { g_logging = nullptr; }
};
}  // namespace

extern "C" {

RUF_API void *ruf_facade_new() { return new Facade; }
RUF_API void ruf_facade_delete(void *h)
{
  Facade *F = (Facade *)h;
  if (!F) return;
  delete F->f;
  delete F;
}
RUF_API void ruf_facade_param_str(void *h, const char *k, const char *v) { ((Facade *)h)->nh.setParam(k, std::string(v)); }
RUF_API void ruf_facade_param_double(void *h, const char *k, double v) { ((Facade *)h)->nh.setParam(k, v); }
RUF_API void ruf_facade_param_bool(void *h, const char *k, int v) { ((Facade *)h)->nh.setParam(k, v != 0); }
RUF_API void ruf_facade_camera_offset(void *h, const double *t, const double *q) { ((Facade *)h)->nh.setCameraOffset(t, q); }
RUF_API void ruf_facade_add_model(void *h, const char *model_param, const char *tf_prefix, const char *geometry_type,
                                  double scale, const char *ignore_csv)
{
  ModelParam m;
  m.model = model_param; m.tf_prefix = tf_prefix; m.geometry_type = geometry_type; m.scale = scale;
  std::stringstream ss(ignore_csv ? ignore_csv : "");
  std::string item;
  while (std::getline(ss, item, ','))
    if (!item.empty()) m.ignore.push_back(item);
  ((Facade *)h)->nh.addModel(m);
}
// constructs RealtimeURDFFilter(nh, argc, argv): reads the parameters set so far
RUF_API int ruf_facade_construct(void *h)
{
  Facade *F = (Facade *)h;
  LogScope ls(F);
  delete F->f;
  F->f = new RealtimeURDFFilter(F->nh, 0, nullptr);
  return 0;
}
RUF_API void ruf_facade_add_resource_root(void *h, const char *dir) { ((Facade *)h)->f->resource_roots_.push_back(dir); }
RUF_API void ruf_facade_set_tf(void *h, const char *frame, const double *q, const double *t) { ((Facade *)h)->f->tf_.setTransform(frame, q, t); }
RUF_API void ruf_facade_erase_tf(void *h, const char *frame) { ((Facade *)h)->f->tf_.erase(frame); }
RUF_API void ruf_facade_subscribers(void *h, int depth, int mask)
{
  Facade *F = (Facade *)h;
  F->f->depth_pub_.subscribers = depth;
  F->f->mask_pub_.subscribers = mask;
}
// filter_callback(Image, CameraInfo).  Returns 0, or -1 when an exception escaped (text in ruf_facade_error).
RUF_API int ruf_facade_callback(void *h, const char *encoding, int width, int height, int step, const void *data,
                                const double *P, double stamp)
{
  Facade *F = (Facade *)h;
  LogScope ls(F);
  auto img = std::make_shared<Image>();
  img->encoding = encoding; img->width = width; img->height = height; img->step = step;
  img->header.stamp = Time(stamp);
  img->data.assign((const uint8_t *)data, (const uint8_t *)data + (size_t)step * height);
  auto info = std::make_shared<CameraInfo>();
  info->width = width; info->height = height;
  std::memcpy(info->P, P, sizeof(info->P));
  try {
    F->f->filter_callback(img, info);
  } catch (const std::exception &e) {
    F->error = e.what();
    return -1;
  }
  return 0;
}
// filter(buffer, glTf, W, H) as the tracker calls it; result through getMaskedDepth()
RUF_API int ruf_facade_filter(void *h, float *buffer, double *glTf, int width, int height, float *masked_out)
{
  Facade *F = (Facade *)h;
  LogScope ls(F);
  try {
    F->f->filter((unsigned char *)buffer, glTf, width, height);
  } catch (const std::exception &e) {
    F->error = e.what();
    return -1;
  }
  if (masked_out && F->f->getMaskedDepth()) std::memcpy(masked_out, F->f->getMaskedDepth(), (size_t)width * height * 4);
  return 0;
}
RUF_API void ruf_facade_projection(void *h, int width, int height, const double *P, double *glTf)
{
  Facade *F = (Facade *)h;
  auto info = std::make_shared<CameraInfo>();
  info->width = width; info->height = height;
  std::memcpy(info->P, P, sizeof(info->P));
  F->f->getProjectionMatrix(info, glTf);
}
RUF_API long ruf_facade_published(void *h, int which) { Facade *F = (Facade *)h; return (long)(which ? F->f->mask_pub_.published : F->f->depth_pub_.published); }
RUF_API long ruf_facade_last_image(void *h, int which, void *out, long cap, char *encoding16)
{
  Facade *F = (Facade *)h;
  const Image &im = which ? F->f->mask_pub_.last_image : F->f->depth_pub_.last_image;
  if (encoding16) { std::strncpy(encoding16, im.encoding.c_str(), 15); encoding16[15] = 0; }
  if (out && (long)im.data.size() <= cap) std::memcpy(out, im.data.data(), im.data.size());
  return (long)im.data.size();
}
RUF_API long ruf_facade_counts(void *h, int what)
{
  Facade *F = (Facade *)h;
  switch (what) {
    case 0: return (long)F->f->renderers_.size();
    case 1: { long n = 0; for (auto *r : F->f->renderers_) n += (long)r->parts().size(); return n; }
    case 2: { long n = 0; for (auto *r : F->f->renderers_) n += (long)r->triangle_parts().size(); return n; }
    case 3: return (long)F->f->tf_.lookups;
    case 4: return (long)F->f->frames_;
    case 5: { long n = 0; for (auto *r : F->f->renderers_) n += (long)r->num_renderables(); return n; }
    case 6: { long n = 0; for (auto *r : F->f->renderers_) n += (long)r->mesh_errors().size(); return n; }
    default: return -1;
  }
}
RUF_API const char *ruf_facade_log(void *h) { return ((Facade *)h)->log.c_str(); }
RUF_API void ruf_facade_clear_log(void *h) { ((Facade *)h)->log.clear(); }
RUF_API const char *ruf_facade_error(void *h) { return ((Facade *)h)->error.c_str(); }
RUF_API double ruf_facade_get_double(void *h, const char *name)
{
  RealtimeURDFFilter *f = ((Facade *)h)->f;
  const std::string n = name;
  if (n == "depth_distance_threshold") return f->depth_distance_threshold_;
  if (n == "filter_replace_value") return f->filter_replace_value_;
  if (n == "far_plane") return f->far_plane_;
  if (n == "near_plane") return f->near_plane_;
  if (n == "camera_tx") return f->camera_tx_;
  if (n == "camera_ty") return f->camera_ty_;
  if (n == "width") return f->width_;
  if (n == "height") return f->height_;
  return -1e300;
}
// the tracker caller's conversions (src/urdf_filtered_tracker.cpp:201-249)
RUF_API void ruf_facade_tracker_depth_to_buffer(const uint16_t *mm, int xres, int yres, float *buffer)
{
  realtime_urdf_filter::tracker_depth_to_buffer(mm, xres, yres, buffer);
}
RUF_API void ruf_facade_tracker_projection(int xres, int yres, double *glTf) { realtime_urdf_filter::tracker_projection(xres, yres, glTf); }
RUF_API void ruf_facade_tracker_masked_depth_to_mm(const float *masked, int xres, int yres, uint16_t *mm)
{
  realtime_urdf_filter::tracker_masked_depth_to_mm(masked, xres, yres, mm);
}
static thread_local std::string g_parse_error;
RUF_API const char *ruf_facade_last_parse_error() { return g_parse_error.c_str(); }
// URDF parsing alone (no GPU): number of parts / triangles a description would produce
RUF_API long ruf_facade_parse_urdf(const char *xml, const char *geometry_type, double scale, const char *ignore_csv,
                                   const char *resource_root, float *tri_out, uint32_t *part_out, long cap_tris,
                                   long *n_parts, double *part_models_identity_tf)
{
  TransformListener tf;
  std::unordered_set<std::string> ignore;
  std::stringstream ss(ignore_csv ? ignore_csv : "");
  std::string item;
  while (std::getline(ss, item, ','))
    if (!item.empty()) ignore.insert(item);
  std::vector<std::string> roots;
  if (resource_root && *resource_root) roots.push_back(resource_root);
  URDFRenderer r(xml, "", "cam", "fixed", tf, geometry_type, scale, ignore, roots);
  g_parse_error = r.error();
  if (!r.ok()) return -1;
  const long n = (long)r.triangle_parts().size();
  if (n_parts) *n_parts = (long)r.parts().size();
  if (tri_out && part_out && n <= cap_tris) {
    std::memcpy(tri_out, r.triangles().data(), r.triangles().size() * sizeof(float));
    std::memcpy(part_out, r.triangle_parts().data(), r.triangle_parts().size() * sizeof(uint32_t));
  }
  if (part_models_identity_tf) {
    std::vector<double> m;
    r.update_link_transforms(Time());
    r.append_part_models(m);
    std::memcpy(part_models_identity_tf, m.data(), m.size() * sizeof(double));
  }
  return n;
}

}  // extern "C"
