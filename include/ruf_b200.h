/*
 * ruf_b200.h -- C ABI of libruf_b200.so, the B200 (sm_100a) implementation of the per-frame
 * hot path of blodow/realtime_urdf_filter:
 *
 *   upload depth -> pose-transform link meshes -> rasterise virtual z-buffer ->
 *   per-pixel  sensor > virtual - max_diff  -> filtered depth + 0/255 mask
 *
 * Every entry point names the piece of the reference it replaces (paths relative to the
 * reference checkout).  Plain pointers and sizes only; no C++/torch types; nothing throws
 * across this boundary -- every function returns a ruf_status and ruf_last_error() holds
 * the text.  All matrices are column-major double[16] exactly as the reference hands them
 * to glMultMatrixd (tf::Transform::getOpenGLMatrix layout).
 *
 * There is NO CPU fallback: every compute entry point fails with RUF_ERR_CUDA when no
 * CUDA device / kernel image for sm_100a is available.
 *
 * Threading: a context owns one CUDA stream and is not thread-safe; distinct contexts are
 * independent (one per camera stream / per GPU).  There is no process-global state
 * (the reference has function-local statics, src/urdf_filter.cpp:210,240,358,388,549).
 */
#ifndef RUF_B200_H
#define RUF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define RUF_API
#else
#define RUF_API __attribute__((visibility("default")))
#endif

typedef struct ruf_context ruf_context;

typedef enum ruf_status {
  RUF_OK = 0,
  RUF_ERR_INVALID = -1,   /* bad argument                                               */
  RUF_ERR_CUDA = -2,      /* CUDA runtime error / no device (text in ruf_last_error)    */
  RUF_ERR_NO_MODEL = -3,  /* filter called before ruf_set_model (reference: returns     */
                          /* silently when renderers_ is empty, src/urdf_filter.cpp:222) */
  RUF_ERR_OVERFLOW = -4,  /* an internal bin/record buffer was too small for a frame;    */
                          /* host-buffer calls grow and retry, device calls report this  */
  RUF_ERR_NOMEM = -5
} ruf_status;

/* Depth encodings of sensor_msgs/Image as handled by filter_callback,
 * src/urdf_filter.cpp:280-289 (input) and :309-316 (output keeps the input encoding). */
typedef enum ruf_encoding {
  RUF_ENC_F32_M = 0,      /* 32FC1, metres, NaN = invalid                                */
  RUF_ENC_U16_MM = 1      /* 16UC1, millimetres, 0 = invalid; x*0.001f in,               */
                          /* saturate(round_half_even(x*1000.f)) out (cv::Mat::convertTo) */
} ruf_encoding;

/* ------------------------------------------------------------------------------------ */
/* Context (replaces initGL / initFrameBufferObject / the FBO + shader objects,           */
/* src/urdf_filter.cpp:386-456, src/FrameBufferObject.cpp, src/shader_wrapper.cpp)        */
/* ------------------------------------------------------------------------------------ */

/* device: CUDA ordinal.  width/height: image size (1..4096).  z_near/z_far: the clip planes
 * (reference hard-codes 0.1 / 8, src/urdf_filter.cpp:53-54); they feed the shader's
 * to_linear_depth and the background quad at 0.99*z_far (:591-596). */
RUF_API int ruf_create(ruf_context **ctx, int device, int width, int height,
                       double z_near, double z_far);
RUF_API int ruf_destroy(ruf_context *ctx);

/* Text of the last error on this context (ctx == NULL: last error of a failed ruf_create
 * on the calling thread).  Never NULL. */
RUF_API const char *ruf_last_error(const ruf_context *ctx);

/* Use a caller-owned cudaStream_t (e.g. torch's current stream) instead of the context's own
 * stream for all subsequent work.  NULL restores the internal stream (so the legacy default
 * stream, whose handle is 0, cannot be selected: pass a created stream). */
RUF_API int ruf_set_stream(ruf_context *ctx, void *cuda_stream);
RUF_API int ruf_sync(ruf_context *ctx);   /* wait; returns deferred RUF_ERR_OVERFLOW if any */

/* ------------------------------------------------------------------------------------ */
/* Model (replaces VBO/IBO creation: RenderableBox::createBoxVBO src/renderable.cpp:133-170,*/
/* SubMesh::init :339-350, and the implicit geometry of glutSolid* :83,96,129)            */
/* ------------------------------------------------------------------------------------ */

/* Static triangle soup in object space.  tri_xyz: T*9 floats (3 vertices x xyz), tri_part:
 * T indices in [0, n_parts) selecting the model matrix of the drawn part (one per
 * glDraw* / glutSolid* call of the reference: a box link contributes two parts, F4).
 * Host pointers; copied to the device.  Calling it again replaces the model. */
RUF_API int ruf_set_model(ruf_context *ctx, const float *tri_xyz, const uint32_t *tri_part,
                          int64_t n_tris, int n_parts);
/* Same, from device memory of the context's device (e.g. after an NCCL broadcast). */
RUF_API int ruf_set_model_device(ruf_context *ctx, const void *d_tri_xyz, const void *d_tri_part,
                                 int64_t n_tris, int n_parts);

/* Capacity control for device-resident batches: max frames per call, and (0 = automatic) the
 * capacities of the internal per-frame big-triangle list and of the record list every 64x64 tile owns. */
RUF_API int ruf_reserve(ruf_context *ctx, int max_batch, int64_t big_capacity, int64_t tile_capacity);

/* ------------------------------------------------------------------------------------ */
/* The per-frame path                                                                     */
/* ------------------------------------------------------------------------------------ */

/* One frame, HOST buffers, synchronous: replaces RealtimeURDFFilter::filter
 * (src/urdf_filter.cpp:207-267) = textureBufferFromDepthBuffer (:332-353) + render (:503-744)
 * including both glGetTexImage readbacks (:729-735).
 *   depth_in   W*H elements of `enc`, row-major, contiguous
 *   proj       getProjectionMatrix output (:459-501)
 *   view       MODELVIEW before the draw loop = LookAt * offset^-1 * camera_transform' (:583-614)
 *   part_model n_parts matrices: link_to_fixed * link_offset [* glTranslate/glScale suffix]
 *              (src/renderable.cpp:59-68, 95, 128, 427)
 *   max_diff / replace_value   the shader uniforms (:629-630)
 *   depth_out  W*H elements of `enc` (gl_FragData[1].r; re-encoded like :309-312 for U16)
 *   mask_out   W*H bytes 0/255 (gl_FragData[3].r read as GL_UNSIGNED_BYTE) or NULL
 *              (= need_mask_ false, :226-230) */
RUF_API int ruf_filter(ruf_context *ctx, const void *depth_in, int enc,
                       const double *proj, const double *view, const double *part_model,
                       float max_diff, float replace_value,
                       void *depth_out, uint8_t *mask_out);

/* n_frames frames, DEVICE buffers, asynchronous on the context's stream (throughput path).
 * d_depth_in / d_depth_out / d_mask_out: n_frames contiguous images.  d_proj: 16 doubles
 * shared by the batch.  d_view: n_frames*16.  d_part_model: n_frames*n_parts*16.
 * d_mask_out may be NULL.  d_zbuf_out (optional, debug): n_frames*W*H float window-space z
 * of the virtual depth buffer (1.0f = nothing drawn).  The workspace grows to n_frames on
 * demand (ruf_reserve pre-sizes it); n_frames <= 65535. */
RUF_API int ruf_filter_batch_device(ruf_context *ctx, int n_frames, const void *d_depth_in, int enc,
                                    const double *d_proj, const double *d_view,
                                    const double *d_part_model,
                                    float max_diff, float replace_value,
                                    void *d_depth_out, uint8_t *d_mask_out, float *d_zbuf_out);

/* n_frames frames, HOST buffers (pinned for full speed: ruf_host_alloc), synchronous; the
 * batch is cut into chunks whose H2D copy, kernels and D2H copy overlap on three streams. */
RUF_API int ruf_filter_batch_host(ruf_context *ctx, int n_frames, const void *depth_in, int enc,
                                  const double *proj, const double *view, const double *part_model,
                                  float max_diff, float replace_value,
                                  void *depth_out, uint8_t *mask_out);

/* ------------------------------------------------------------------------------------ */
/* Forward kinematics on the device ("next" row of the hot path: the caller side)          */
/* ------------------------------------------------------------------------------------ */

/* Kinematic tree of the loaded model.  Replaces, per frame, the L tf lookups of
 * URDFRenderer::update_link_transforms (src/urdf_renderer.cpp:173-190) and the camera lookup of
 * render (src/urdf_filter.cpp:522): link poses are computed from joint positions on the device
 * (what robot_state_publisher + tf do on the host).
 *   parent[l]      index of the parent link (< l) or -1 for the fixed frame
 *   joint_type[l]  0 fixed, 1 revolute/continuous, 2 prismatic
 *   origin[l]      16 doubles, column-major: joint origin (parent_T_joint)
 *   axis[l]        3 doubles, unit joint axis
 *   part_link[p]   link that part p rides on; part_local[p] = link_offset [* suffix] (16 doubles)
 *   cam_link       link the camera rides on (-1: fixed frame); cam_mount = optical frame in that link
 *   view_pre       LookAt * inverse(camera_offset) (ruf_view_matrix with an identity camera transform)
 * All host pointers; n_parts must equal the loaded model's. */
RUF_API int ruf_set_kinematics(ruf_context *ctx, int n_links, const int32_t *parent, const int32_t *joint_type,
                               const double *origin, const double *axis, const int32_t *part_link,
                               const double *part_local, int cam_link, const double *cam_mount,
                               const double *view_pre);

/* Joint positions -> part models + view matrices, device buffers, asynchronous.
 * d_joint_q [n_frames][n_links]; d_part_model_out [n_frames][n_parts][16]; d_view_out [n_frames][16].
 * camera_tx / camera_ty as returned by ruf_projection_matrix. */
RUF_API int ruf_fk_batch_device(ruf_context *ctx, int n_frames, const double *d_joint_q, double camera_tx,
                                double camera_ty, double *d_part_model_out, double *d_view_out);

/* ruf_filter_batch_device with the poses computed on the device from joint positions. */
RUF_API int ruf_filter_batch_device_fk(ruf_context *ctx, int n_frames, const void *d_depth_in, int enc,
                                       const double *d_proj, const double *d_joint_q, double camera_tx,
                                       double camera_ty, float max_diff, float replace_value,
                                       void *d_depth_out, uint8_t *d_mask_out, float *d_zbuf_out);

/* Mask output format of every later call on this context.  RUF_MASK_BYTES (default): one byte per pixel, 0 / 255, the
 * MONO8 image the reference publishes (glGetTexImage(GL_RED, GL_UNSIGNED_BYTE) of attachment 3, src/urdf_filter.cpp:731-735,
 * :321-329).  RUF_MASK_BITS (opt-in): one BIT per pixel, bit i of byte k = pixel 8 k + i of the row-major image (numpy
 * unpackbits(bitorder="little")); mask buffers are then width*height/8 bytes per frame and the device -> host traffic of a
 * 16UC1 frame drops from 3 to 2.125 bytes per pixel.  Needs width % 8 == 0.  The C++ facade expands the bits to MONO8
 * where it publishes (host/urdf_filter.cpp), so subscribers see the reference's image either way. */
#define RUF_MASK_BYTES 0
#define RUF_MASK_BITS 1
RUF_API int ruf_set_mask_format(ruf_context *ctx, int format);

/* Measurement aid: the chunked host pipeline of ruf_filter_batch_host with the same staging slots, streams, chunk
 * sizes and copies (H2D of depth + matrices, D2H of depth + mask) but WITHOUT the kernels: the PCIe / host-memory
 * ceiling of the end-to-end number on this box (bench.py `e2e.copy_ceiling`).  The output buffers receive whatever
 * the staging slots hold. */
RUF_API int ruf_host_copy_ceiling(ruf_context *ctx, int n_frames, const void *depth_in, int enc, const double *proj,
                                  const double *view, const double *part_model, void *depth_out, uint8_t *mask_out);

/* ------------------------------------------------------------------------------------ */
/* Multi-GPU group: one host process, N devices, frames sharded, no per-frame   */
/* collective.  The reference is single-GPU and single-context (function-local statics,       */
/* src/urdf_filter.cpp:210,388,549); this is what lets ONE node feed 8 GPUs.                  */
/* ------------------------------------------------------------------------------------ */
typedef struct ruf_group ruf_group;
/* One context per device (devices == NULL: 0..n-1) and, for n > 1, one NCCL communicator per device
 * (ncclCommInitAll; libnccl.so.2 is loaded at run time). */
RUF_API int ruf_group_create(ruf_group **grp, int n_devices, const int *devices, int width, int height,
                             double z_near, double z_far);
RUF_API int ruf_group_destroy(ruf_group *grp);
RUF_API int ruf_group_size(const ruf_group *grp);
/* The context of member i: its own stream, usable with every ruf_* call (one camera stream per GPU, BASELINE configs[3]). */
RUF_API ruf_context *ruf_group_context(ruf_group *grp, int i);
RUF_API const char *ruf_group_last_error(const ruf_group *grp);
/* ruf_set_model for every member: the model is prepared once on the host, uploaded to device 0 and sent to the other
 * devices with one ncclBroadcast per static buffer (over NVLink / NVSwitch). */
RUF_API int ruf_group_set_model(ruf_group *grp, const float *tri_xyz, const uint32_t *tri_part, int64_t n_tris, int n_parts);
RUF_API int64_t ruf_group_broadcast_bytes(const ruf_group *grp);   /* bytes each non-root device received */
/* ruf_filter_batch_host over all members: chunk j of frames_per_chunk frames (0 = automatic) runs on device j mod N
 * (frames_per_chunk = 1: frame k -> GPU k mod N, BASELINE configs[4]); one host thread per device; results are written
 * at their frame's position, i.e. in sequence order. */
RUF_API int ruf_group_filter_batch_host(ruf_group *grp, int n_frames, const void *depth_in, int enc, const double *proj,
                                        const double *view, const double *part_model, float max_diff,
                                        float replace_value, void *depth_out, uint8_t *mask_out, int frames_per_chunk);

RUF_API int ruf_host_alloc(void **ptr, size_t bytes);   /* cudaHostAlloc (pinned, mapped) */
RUF_API int ruf_host_free(void *ptr);
/* 1 when p lies in page-locked host memory known to the CUDA runtime (ruf_host_alloc, cudaHostAlloc, cudaHostRegister):
 * ruf_filter takes its single-frame graph path for such buffers (see ruf_filter).  0 otherwise. */
RUF_API int ruf_host_is_pinned(const void *p);

/* Counters of the most recent completed call (debug / bench bookkeeping). */
typedef struct ruf_stats {
  int64_t frames;            /* frames in the last call                                  */
  int64_t kernel_launches;   /* kernels launched by the last call                        */
  int64_t visible_tris;      /* sum over frames of triangles kept (binned) by setup        */
  int64_t binned_refs;       /* sum over frames of (triangle, tile) references            */
  int64_t big_tris;          /* sum over frames of triangles routed to the per-frame list */
  int64_t h2d_bytes, d2h_bytes;
} ruf_stats;
RUF_API int ruf_get_stats(ruf_context *ctx, ruf_stats *out);

/* Per-kernel device timing (the analogue of the reference's gettimeofday bookkeeping around
 * filter(), src/urdf_filter.cpp:239-266, but per stage and on the device).  When enabled every
 * launch sequence is bracketed by CUDA events on the launching stream.  ruf_get_stage_times
 * synchronises, then returns the accumulated milliseconds of the three kernels in launch order
 * {pose, setup+bin, raster+filter} and the number of launch sequences they cover. */
RUF_API int ruf_set_profiling(ruf_context *ctx, int enable);
RUF_API int ruf_get_stage_times(ruf_context *ctx, double *ms3, int64_t *calls, int reset);

/* ------------------------------------------------------------------------------------ */
/* Host-side matrices (double precision, same operation order as the reference + tf/GLU)  */
/* ------------------------------------------------------------------------------------ */

/* getProjectionMatrix, src/urdf_filter.cpp:459-501.  P = CameraInfo.P (3x4 row-major). */
RUF_API void ruf_projection_matrix(const double *P, int width, int height,
                                   double z_near, double z_far,
                                   double *glTf, double *camera_tx, double *camera_ty);
/* gluLookAt(0,0,0, 0,0,1, 0,1,0), src/urdf_filter.cpp:587. */
RUF_API void ruf_lookat(double *m);
/* LookAt * inverse(camera_offset) * camera_transform shifted by tx/ty, :583-614.
 * Quaternions are (x,y,z,w); cam_q/cam_t = lookupTransform(cam_frame, fixed_frame). */
RUF_API void ruf_view_matrix(const double *offset_q, const double *offset_t,
                             const double *cam_q, const double *cam_t,
                             double camera_tx, double camera_ty, double *view);
/* link_to_fixed * link_offset (normalised quaternion, src/urdf_renderer.cpp:160-164)
 * [* suffix], src/renderable.cpp:59-68.  suffix may be NULL. */
RUF_API void ruf_part_model(const double *link_q, const double *link_t,
                            const double *off_q, const double *off_t,
                            const double *suffix, double *model);

/* ------------------------------------------------------------------------------------ */
/* Geometry of the primitive renderables (9 floats per triangle)                          */
/* ------------------------------------------------------------------------------------ */
RUF_API int ruf_box_triangles(float dimx, float dimy, float dimz, float *out);      /* 12, renderable.cpp:135-164 */
RUF_API int ruf_cube_triangles(float size, float *out);                             /* 12, glutSolidCube :129 */
RUF_API int ruf_sphere_triangles(float radius, int slices, int stacks, float *out); /* glutSolidSphere :83 */
RUF_API int ruf_cylinder_triangles(float radius, float height, int slices, int stacks, float *out); /* :96 */
RUF_API int ruf_sphere_triangle_count(int slices, int stacks);
RUF_API int ruf_cylinder_triangle_count(int slices, int stacks);

/* Diagnostics of the model ingest (CPU only, no device needed).  ruf_set_model cuts the soup into
 * "meshlets" (runs of consecutive triangles whose bit-identical vertices are welded, the device-side
 * replacement of the reference's VBO/IBO pairs, src/renderable.cpp:339-350); this builds them with the
 * given limits (the library uses 512 vertices, 1023 triangles, 32 parts, and 256 triangles for its fine cut) and expands them again:
 * out_xyz (n_tris + 2) * 9 floats, out_part n_tris + 2 -- the input soup bit for bit, followed by the
 * two triangles of the background quad (part = n_parts).  counts[3] = meshlets, welded vertices, triangles. */
RUF_API int ruf_meshlet_roundtrip(const float *tri_xyz, const uint32_t *tri_part, int64_t n_tris, int n_parts,
                                  double z_far, int max_verts, int max_tris, int max_parts,
                                  float *out_xyz, uint32_t *out_part, int64_t *counts);

/* The same for BOTH cuts of the model that ruf_set_model keeps in one set of arrays (the throughput cut with at most max_tris
 * triangles per meshlet, then the fine cut with at most fine_tris, used by launches of one frame): out_xyz / out_part hold
 * 2 * (n_tris + 2) triangles, each cut expanded to the soup + the background quad.  counts[4] = meshlets of the first cut,
 * meshlets of the fine cut, welded vertices, index triples. */
RUF_API int ruf_meshlet_sets_roundtrip(const float *tri_xyz, const uint32_t *tri_part, int64_t n_tris, int n_parts,
                                       double z_far, int max_verts, int max_tris, int fine_tris, int max_parts,
                                       float *out_xyz, uint32_t *out_part, int64_t *counts);

RUF_API const char *ruf_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RUF_B200_H */
