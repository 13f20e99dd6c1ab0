// urdf_filter.cpp -- see urdf_filter.h.  Control flow follows src/urdf_filter.cpp of the reference
// step by step (line numbers in comments); the GL calls are replaced by ruf_* calls.
#include "urdf_filter.h"

#include <sys/time.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <unordered_set>

using namespace ruf_host;

namespace realtime_urdf_filter {

RealtimeURDFFilter::RealtimeURDFFilter(NodeHandle &nh, int argc, char **argv)
    : nh_(nh), argc_(argc), argv_(argv)
{
  if (!nh_.getParam("fixed_frame", fixed_frame_)) logf(LOG_FATAL, "fixed_frame paramter!");              // :58-61
  logf(LOG_INFO, "using fixed frame %s", fixed_frame_.c_str());
  if (!nh_.getParam("camera_frame", cam_frame_)) logf(LOG_FATAL, "need a camera_frame paramter!");        // :66-69
  logf(LOG_INFO, "using camera frame %s", cam_frame_.c_str());
  if (!nh_.getCameraOffset(camera_offset_t_, camera_offset_q_)) {                                          // :73-96
    camera_offset_t_[0] = camera_offset_t_[1] = camera_offset_t_[2] = 0.0;
    camera_offset_q_[0] = camera_offset_q_[1] = camera_offset_q_[2] = 0.0;
    camera_offset_q_[3] = 1.0;
  }
  if (!nh_.getParam("depth_distance_threshold", depth_distance_threshold_))                                // :99-102
    logf(LOG_FATAL, "need a depth_distance_threshold paramter!");
  logf(LOG_INFO, "using depth distance threshold %f", depth_distance_threshold_);
  if (!nh_.getParam("show_gui", show_gui_)) show_gui_ = false;                                             // :106
  // not a parameter of the reference: read the mask back as 1 bit per pixel and expand it to MONO8 here (saves 0.875
  // byte per pixel of device -> host traffic; subscribers see the same 0 / 255 image)
  if (!nh_.getParam("packed_mask_readback", packed_mask_)) packed_mask_ = false;
  if (!nh_.getParam("pinned_staging", pinned_staging_)) pinned_staging_ = true;
  if (!nh_.getParam("filter_replace_value", filter_replace_value_)) filter_replace_value_ = 0;             // :110
  logf(LOG_INFO, "using filter replace value %f", filter_replace_value_);
}

RealtimeURDFFilter::~RealtimeURDFFilter()
{
  ruf_host_free(masked_depth_);
  ruf_host_free(mask_);
  ruf_host_free(pin_in_);
  ruf_host_free(pin_out_);
  for (URDFRenderer *r : renderers_) delete r;
  if (ctx_) ruf_destroy(ctx_);
}

void RealtimeURDFFilter::loadModels()
{
  std::vector<ModelParam> models;
  if (!nh_.getModels(models)) { logf(LOG_ERROR, "models parameter must be an array!"); return; }           // :193-196
  for (const ModelParam &elem : models) {
    std::string content;
    if (!nh_.getParam(elem.model, content)) {                                                              // :145-158
      logf(LOG_ERROR, "Parameter [%s] does not exist, and was not found by searchParam()", elem.model.c_str());
      continue;
    }
    if (content.empty()) { logf(LOG_ERROR, "URDF is empty"); continue; }                                   // :160-164
    std::unordered_set<std::string> ignore(elem.ignore.begin(), elem.ignore.end());
    logf(LOG_INFO, "Loading URDF model: %s", elem.model.c_str());
    renderers_.push_back(new URDFRenderer(content, elem.tf_prefix, cam_frame_, fixed_frame_, tf_, elem.geometry_type,
                                          elem.scale, ignore, resource_roots_));                           // :190
    // Only STL meshes are read here (the reference reads whatever Assimp does): a mesh that could not be loaded leaves
    // its link unfiltered.  `strict_meshes` (not a parameter of the reference) turns that into the same exception a
    // model-less filter raises (:426-427) instead of an ERROR line per mesh.
    bool strict = false;
    nh_.getParam("strict_meshes", strict);
    const std::vector<std::string> &bad = renderers_.back()->mesh_errors();
    if (strict && !bad.empty())
      throw std::runtime_error("Could not load " + std::to_string(bad.size()) + " mesh(es), first: " + bad.front());
  }
}

double RealtimeURDFFilter::getTime()
{
  timeval t;
  gettimeofday(&t, nullptr);
  return t.tv_sec + 1e-6 * t.tv_usec;
}

void RealtimeURDFFilter::initFrameBufferObject() { fbo_initialized_ = ctx_ != nullptr; }

void RealtimeURDFFilter::upload_models()
{
  std::vector<float> tri;
  std::vector<uint32_t> part;
  uint32_t base = 0;
  for (URDFRenderer *r : renderers_) {
    tri.insert(tri.end(), r->triangles().begin(), r->triangles().end());
    for (uint32_t p : r->triangle_parts()) part.push_back(base + p);
    base += (uint32_t)r->parts().size();
  }
  model_parts_ = base;
  if (ruf_set_model(ctx_, tri.data(), part.data(), (int64_t)part.size(), (int)base) != RUF_OK)
    throw std::runtime_error(std::string("ruf_set_model: ") + ruf_last_error(ctx_));
}

void RealtimeURDFFilter::initGL()
{
  // The reference creates the GLUT window, GLEW, the FBO, then loads the models (:386-436) and throws
  // std::runtime_error when that fails.  Here: device context, models, output buffers.  Unlike the
  // reference a second call (image size changed) starts from scratch instead of appending a second
  // copy of every renderer and leaking the old buffers.
  if (ctx_) { ruf_destroy(ctx_); ctx_ = nullptr; }
  for (URDFRenderer *r : renderers_) delete r;
  renderers_.clear();
  int dev = 0;
  if (const char *e = std::getenv("RUF_DEVICE")) dev = std::atoi(e);
  if (ruf_create(&ctx_, dev, width_, height_, near_plane_, far_plane_) != RUF_OK)
    throw std::runtime_error(std::string("Could not initialise the CUDA path: ") + ruf_last_error(nullptr));   // ~ :413-416
  if (packed_mask_ && width_ % 8 == 0) {
    if (ruf_set_mask_format(ctx_, RUF_MASK_BITS) != RUF_OK) packed_mask_ = false;
  } else {
    packed_mask_ = false;
  }
  initFrameBufferObject();
  loadModels();                                                                                            // :423
  if (renderers_.empty()) throw std::runtime_error("Could not load any models for filtering!");           // :426-427
  upload_models();
  // the read-back buffers of :433-435, page-locked: ruf_filter writes them from the device without a staging copy
  ruf_host_free(masked_depth_);
  ruf_host_free(mask_);
  ruf_host_free(pin_in_);
  ruf_host_free(pin_out_);
  masked_depth_ = nullptr; mask_ = nullptr; pin_in_ = pin_out_ = nullptr;
  pin_bytes_ = (size_t)width_ * height_ * sizeof(float);
  void *p0 = nullptr, *p1 = nullptr, *p2 = nullptr, *p3 = nullptr;
  if (ruf_host_alloc(&p0, pin_bytes_) != RUF_OK || ruf_host_alloc(&p1, (size_t)width_ * height_) != RUF_OK ||
      ruf_host_alloc(&p2, pin_bytes_) != RUF_OK || ruf_host_alloc(&p3, pin_bytes_) != RUF_OK)
    throw std::runtime_error("Could not allocate the read-back buffers");
  masked_depth_ = (float *)p0;
  mask_ = (unsigned char *)p1;
  pin_in_ = (unsigned char *)p2;
  pin_out_ = (unsigned char *)p3;
}

void RealtimeURDFFilter::getProjectionMatrix(const CameraInfoConstPtr &info, double *glTf)
{
  ruf_projection_matrix(info->P, (int)info->width, (int)info->height, near_plane_, far_plane_, glTf, &camera_tx_,
                        &camera_ty_);
}

void RealtimeURDFFilter::textureBufferFromDepthBuffer(unsigned char *buffer, int /*size_in_bytes*/)
{
  staged_buffer_ = buffer;   // the H2D copy is part of ruf_filter (render)
}

int RealtimeURDFFilter::run_device(const void *depth_in, int enc, const double *P, Time stamp, void *depth_out,
                                   unsigned char *mask_out)
{
  if (!fbo_initialized_) return RUF_ERR_INVALID;                                                           // :516-518
  StampedTransform camera_transform;
  try {
    tf_.lookupTransform(cam_frame_, fixed_frame_, stamp, camera_transform);                               // :522
  } catch (const TransformException &ex) {
    logf(LOG_ERROR, "%s", ex.what());                                                                      // :531-534: outputs stay stale
    return RUF_ERR_INVALID;
  }
  double view[16];
  ruf_view_matrix(camera_offset_q_, camera_offset_t_, camera_transform.q, camera_transform.t, camera_tx_,
                  camera_ty_, view);                                                                       // :583-614
  std::vector<double> models;
  models.reserve(16 * model_parts_);
  for (URDFRenderer *r : renderers_) {                                                                     // :635-638
    r->update_link_transforms(stamp);
    r->append_part_models(models);
  }
  unsigned char *mask_dst = mask_out;
  if (packed_mask_ && mask_out) {
    mask_bits_.resize((size_t)width_ * height_ / 8);
    mask_dst = mask_bits_.data();
  }
  // pageable caller buffers (a message's data vector) go through the facade's page-locked staging
  const size_t img_bytes = (size_t)width_ * height_ * (enc == RUF_ENC_F32_M ? 4 : 2);
  const void *src = depth_in;
  void *dst = depth_out;
  if (pinned_staging_ && pin_in_ && img_bytes <= pin_bytes_) {
    if (!ruf_host_is_pinned(depth_in)) { std::memcpy(pin_in_, depth_in, img_bytes); src = pin_in_; }
    if (!ruf_host_is_pinned(depth_out)) dst = pin_out_;
  }
  int rc = ruf_filter(ctx_, src, enc, P, view, models.data(), (float)depth_distance_threshold_,
                      (float)filter_replace_value_, dst, mask_dst);                                        // :625-631, :729-735
  if (rc == RUF_OK && dst != depth_out) std::memcpy(depth_out, dst, img_bytes);
  if (rc == RUF_OK && packed_mask_ && mask_out) {
    // bit i of byte k = pixel 8 k + i -> the MONO8 bytes the reference reads back (GL_UNSIGNED_BYTE, :731-735)
    const size_t n = mask_bits_.size();
    for (size_t k = 0; k < n; ++k) {
      const unsigned b = mask_bits_[k];
      for (int i = 0; i < 8; ++i) mask_out[8 * k + i] = (b >> i) & 1u ? 255 : 0;
    }
  }
  if (rc != RUF_OK) {
    last_error_ = ruf_last_error(ctx_);
    logf(LOG_ERROR, "CUDA path failed: %s", last_error_.c_str());
  }
  return rc;
}

void RealtimeURDFFilter::render(const double *camera_projection_matrix, Time timestamp)
{
  run_device(staged_buffer_, RUF_ENC_F32_M, camera_projection_matrix, timestamp, masked_depth_,
             need_mask_ ? mask_ : nullptr);
}

void RealtimeURDFFilter::filter(unsigned char *buffer, double *glTf, int width, int height, Time timestamp)
{
  const double begin = getTime();
  if (width_ != width || height_ != height) {                                                              // :212-219
    if (width_ != 0 || height_ != 0) logf(LOG_ERROR, "image size has changed (%ix%i) -> (%ix%i)", width_, height_, width, height);
    width_ = width;
    height_ = height;
    this->initGL();
  }
  if (renderers_.empty()) return;                                                                          // :222-224
  need_mask_ = mask_pub_.getNumSubscribers() > 0;                                                          // :226-230
  textureBufferFromDepthBuffer(buffer, width * height * (int)sizeof(float));                               // :234
  render(glTf, timestamp);                                                                                 // :237
  ++frames_;                                                                                               // :239-266
  timings_.push_back((getTime() - begin) * 1000.0);
  if (timings_.size() >= 30) {
    double mn = *std::min_element(timings_.begin(), timings_.end()), mx = *std::max_element(timings_.begin(), timings_.end());
    double avg = 0;
    for (double v : timings_) avg += v;
    avg /= timings_.size();
    logf(LOG_DEBUG, "Average framerate: %f Hz (min %f ms, max %f ms, avg %f ms)", 1000.0 / avg, mn, mx, avg);
    timings_.clear();
  }
}

void RealtimeURDFFilter::filter_callback(const ImageConstPtr &img, const CameraInfoConstPtr &camera_info)
{
  // Anything that is not 32FC1 is treated as 16UC1 millimetres (:280-289).  The reference converts to
  // float on the CPU and back (:311); here the native encoding goes to the device and comes back.
  const bool is_f32 = img->encoding == "32FC1";
  const int enc = is_f32 ? RUF_ENC_F32_M : RUF_ENC_U16_MM;
  const size_t es = is_f32 ? 4 : 2;
  const int W = (int)img->width, H = (int)img->height;
  if (img->step != W * es || img->data.size() < (size_t)W * H * es) {
    logf(LOG_ERROR, "depth image must be contiguous %s (step %u, width %d)", is_f32 ? "32FC1" : "16UC1", img->step, W);
    return;
  }
  double glTf[16];
  getProjectionMatrix(camera_info, glTf);                                                                  // :300

  const double begin = getTime();
  if (width_ != W || height_ != H) {
    if (width_ != 0 || height_ != 0) logf(LOG_ERROR, "image size has changed (%ix%i) -> (%ix%i)", width_, height_, W, H);
    width_ = W;
    height_ = H;
    this->initGL();
  }
  if (renderers_.empty()) return;
  need_mask_ = mask_pub_.getNumSubscribers() > 0;

  Image out;
  out.header = img->header;                                                                                // :315
  out.encoding = img->encoding;                                                                            // :316
  out.height = img->height; out.width = img->width; out.step = img->step;
  out.data.resize((size_t)W * H * es);
  int rc;
  if (is_f32) {
    rc = run_device(img->data.data(), enc, glTf, img->header.stamp, masked_depth_, need_mask_ ? mask_ : nullptr);
    if (rc == RUF_OK) std::memcpy(out.data.data(), masked_depth_, out.data.size());
  } else {
    rc = run_device(img->data.data(), enc, glTf, img->header.stamp, out.data.data(), need_mask_ ? mask_ : nullptr);
  }
  ++frames_;
  timings_.push_back((getTime() - begin) * 1000.0);
  if (timings_.size() >= 30) timings_.clear();
  if (rc != RUF_OK && is_f32) std::memcpy(out.data.data(), masked_depth_, out.data.size());   // stale output, like the reference
  if (rc != RUF_OK && !is_f32) return;   // no stale 16UC1 image is kept on this path

  if (depth_pub_.getNumSubscribers() > 0) depth_pub_.publish(out, *camera_info);                          // :306-319
  if (mask_pub_.getNumSubscribers() > 0) {                                                                 // :321-329
    Image m;
    m.header = img->header;
    m.encoding = "mono8";
    m.height = img->height; m.width = img->width; m.step = img->width;
    m.data.assign(mask_, mask_ + (size_t)W * H);
    mask_pub_.publish(m, *camera_info);
  }
}

void tracker_depth_to_buffer(const uint16_t *depth_mm, int xres, int yres, float *buffer)
{
  for (int y = 0; y < yres; ++y)
    for (int x = 0; x < xres; ++x)
      buffer[x + y * xres] = (float)(depth_mm[(xres - x - 1) + y * xres] * 0.001);                         // :201-207
}

void tracker_projection(int xres, int yres, double *glTf)
{
  float P[12];                                                                                             // :209-212
  P[0] = 585.260f; P[1] = 0.0f;     P[2] = 317.387f; P[3] = 0.0f;
  P[4] = 0.0f;     P[5] = 585.028f; P[6] = 239.264f; P[7] = 0.0f;
  P[8] = 0.0f;     P[9] = 0.0f;     P[10] = 1.0f;    P[11] = 0.0f;
  const double fx = P[0], fy = P[5], cx = P[2], cy = P[6];
  const double far_plane = 8, near_plane = 0.1;                                                            // :218-219
  for (int i = 0; i < 16; ++i) glTf[i] = 0.0;
  glTf[0] = -2.0 * fx / xres;                                                                              // :227-236
  glTf[5] = 2.0 * fy / yres;
  glTf[8] = 2.0 * (0.5 - cx / xres);
  glTf[9] = 2.0 * (cy / yres - 0.5);
  glTf[10] = -(far_plane + near_plane) / (far_plane - near_plane);
  glTf[14] = -2.0 * far_plane * near_plane / (far_plane - near_plane);
  glTf[11] = -1;
}

void tracker_masked_depth_to_mm(const float *masked_depth, int xres, int yres, uint16_t *depth_mm)
{
  for (int y = 0; y < yres; ++y)
    for (int x = 0; x < xres; ++x) {
      const float v = masked_depth[x + y * xres] * 1000;                                                   // :247
      depth_mm[x + y * xres] = !(v > 0.0f) ? (uint16_t)0 : (v >= 65535.0f ? (uint16_t)65535 : (uint16_t)v);
    }
}

}  // namespace realtime_urdf_filter
