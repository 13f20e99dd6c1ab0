// ruf_api.cu -- the extern "C" boundary of libruf_b200.so (include/ruf_b200.h): context,
// device memory, streams, host<->device staging and the launch sequence.  No arithmetic of the
// path lives here except the three scalar shader constants.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <utility>
#include <vector>

#include "../../include/ruf_b200.h"
#include "ruf_device.cuh"
#include "ruf_meshlet.h"

using namespace ruf;

struct ruf_context {
  int device = 0;
  int W = 0, H = 0;
  double z_near = 0.1, z_far = 8.0;
  std::string err;

  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;       // copy streams of the host-batch pipeline
  // optional (RUF_SLICE_FRAMES=n): device-resident batches of >= 2n frames are cut into slices that alternate
  // between two auxiliary streams, so that the setup kernel of one slice overlaps the raster kernel of the
  // previous one.  Off by default: with the current kernels it no longer pays (profiles/r01_experiments.md),
  // and kernel-replay profilers serialise the two streams.
  int slice_frames = 0;
  cudaStream_t s_aux[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};

  // model
  long long n_tris = 0;
  int n_parts = 0;
  bool have_model = false;
  int n_meshlets = 0;            // throughput cut: headers [0, n_meshlets)
  int n_meshlets_fine = 0;       // fine cut (launches of one or a few frames): headers [n_meshlets, n_meshlets + n_meshlets_fine)
  uint4 *meshlets = nullptr;    // [n_meshlets] headers (ruf_device.cuh Model)
  float4 *mverts = nullptr;     // welded vertices of all meshlets
  uint32_t *mtris = nullptr;    // packed local indices
  float *part_aabb = nullptr;   // [n_parts][6]

  // workspace
  int max_batch = 0;            // frames the workspace is sized for
  int want_batch = 0;           // set by ruf_reserve
  long long want_big = 0, want_bin = 0;
  Dims dims{};
  Workspace ws{};
  double *d_lookat = nullptr;

  // host-call staging (device side + pinned side)
  int stage_frames = 0;         // frames per staging slot
  void *d_in[2] = {nullptr, nullptr};
  void *d_out[2] = {nullptr, nullptr};
  uint8_t *d_mask[2] = {nullptr, nullptr};
  double *d_mats[2] = {nullptr, nullptr};
  double *h_mats[2] = {nullptr, nullptr};
  cudaEvent_t ev_in[2]{}, ev_k[2]{}, ev_out[2]{};
  uint32_t *h_status = nullptr;  // pinned

  // forward kinematics (ruf_set_kinematics)
  Kinematics kin{};
  void *kin_blob = nullptr;          // one device allocation holding all kinematics arrays
  double *fk_links = nullptr, *fk_pm = nullptr, *fk_view = nullptr;
  int fk_frames = 0;                 // what the FK buffers were sized for: frames x links, frames x parts
  int fk_links_n = 0, fk_parts_n = 0;

  int mask_format = RUF_MASK_BYTES;  // ruf_set_mask_format

  // single-frame host call (ruf_filter with pinned buffers): the whole sequence -- three kernels; with unmapped buffers also uploads, memset,
  // read-backs, status word -- is ONE captured CUDA graph with the depth upload running beside the pose / setup kernels
  struct FrameGraph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    cudaGraphNode_t n_in = nullptr, n_out = nullptr, n_mask = nullptr, n_raster = nullptr;
    int direct = -1, bg_mode = -1;
    // the raster kernel's launch parameters as captured (retargeting its FrameBuffers argument needs the whole list)
    struct RasterArgs {
      Dims d; const TriRec *big; const BinRec *bins; const uint32_t *ctr; const uint4 *tinfo; ShaderParams sp; FrameBuffers fb; uint32_t *status;
    } ra{};
    cudaKernelNodeParams raster_kp{};
    const void *in = nullptr; void *out = nullptr; uint8_t *mask = nullptr;
    int enc = -1, mask_format = -1, n_parts = -1, multipass = -1;
    float max_diff = 0.f, replace_value = 0.f;
    const void *ws_bins = nullptr, *stage_in = nullptr;
    uint32_t cap_big = 0, cap_tile = 0;
  } fg;
  bool use_graph = true;             // RUF_NO_GRAPH=1 turns the path off (A/B, debugging)
  // the background quad's big-list records for the projection matrix they were set up with (single-frame graph)
  uint32_t *bg_seed = nullptr;       // device: word 0 = count, words 4.. = records
  double bg_proj[16] = {0};
  int bg_state = 0;                  // 0 = not filled, 1 = valid for bg_proj, -1 = not usable (more records than kBgSeedMax)
  bool bg_cache = true;              // RUF_BG_CACHE=0 turns it off
  // page-locked staging of ruf_filter for callers with pageable buffers (RUF_HOST_STAGING=0 turns it off)
  void *hs_in = nullptr, *hs_out = nullptr, *hs_mask = nullptr;
  bool host_staging = true;
  int direct_mode = 7;               // single-frame graph: zero-copy bits (1 outputs, 2 input, 4 matrices + status), RUF_DIRECT
  uint32_t *launch_host_status = nullptr;   // set around the capture of the single-frame graph: FrameBuffers::host_status
  int fine_mode = -1;                // fine meshlet cut for small launches: -1 automatic, 0 / 1 forced (RUF_FINE_MESHLETS)
  int cluster_mode = -1;             // cluster-split raster variant: -1 automatic (small launches), 0 / 1 forced (RUF_CLUSTER)
  int multipass_mode = -1;           // raster kernel variant: -1 automatic (share of wide records in the last launches), 0 / 1 forced
  bool multipass = false;            // the current choice
  double wide_share = -1.0;          // wide / kept records of the launches since the previous status read-back (-1: none yet)

  ruf_stats stats{};
  int last_frames = 0;

  // optional per-kernel timing (ruf_set_profiling)
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;     // kNumStages + 1 events per recorded launch sequence
  size_t ev_used = 0;                   // events in use since the last ruf_get_stage_times
  double stage_ms[kNumStages] = {0, 0, 0};
  int64_t stage_calls = 0;
};

static thread_local std::string g_create_error;

static int fail(ruf_context *c, int code, const char *fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_error = buf;
  return code;
}

#define RUF_CUDA(c, call)                                                                     \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess)                                                                   \
      return fail((c), RUF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                  __FILE__, __LINE__);                                                        \
  } while (0)

static void free_workspace(ruf_context *c)
{
  cudaFree(c->ws.mvp); cudaFree(c->ws.vis); cudaFree(c->ws.ctr); cudaFree(c->ws.big); cudaFree(c->ws.bins); cudaFree(c->ws.tinfo);
  c->ws.tinfo = nullptr;
  c->ws.mvp = nullptr; c->ws.vis = nullptr; c->ws.ctr = nullptr; c->ws.big = nullptr; c->ws.bins = nullptr;
  c->max_batch = 0;
}

static void drop_frame_graph(ruf_context *c)
{
  if (c->fg.exec) cudaGraphExecDestroy(c->fg.exec);
  if (c->fg.graph) cudaGraphDestroy(c->fg.graph);
  c->fg = ruf_context::FrameGraph{};
}

static void free_staging(ruf_context *c)
{
  drop_frame_graph(c);
  for (int i = 0; i < 2; ++i) {
    cudaFree(c->d_in[i]); cudaFree(c->d_out[i]); cudaFree(c->d_mask[i]); cudaFree(c->d_mats[i]);
    if (c->h_mats[i]) cudaFreeHost(c->h_mats[i]);
    c->d_in[i] = c->d_out[i] = nullptr; c->d_mask[i] = nullptr; c->d_mats[i] = nullptr; c->h_mats[i] = nullptr;
  }
  c->stage_frames = 0;
}

// The record lists are sized for the worst tile of every frame (fixed capacity per (frame, tile)), so the workspace
// grows with the batch: 35 MB per frame at 640x480 / 90k triangles, 193 MB per frame at 1080p / 500k.  It is capped
// at a byte budget (RUF_WORKSPACE_GB, default 48 of the 180 GB); a device batch with more frames than fit runs as
// equal sub-batches, one after the other on the same stream, through the same workspace.
static long long default_cap_tile(const ruf_context *c)
{
  long long cap_tile = c->want_bin;
  if (cap_tile <= 0) {
    cap_tile = 8 * (c->n_tris + 2) / c->dims.ntiles;
    if (cap_tile < 2048) cap_tile = 2048;
    if (cap_tile > c->n_tris + 2) cap_tile = c->n_tris + 2;
    cap_tile = (cap_tile + 63) & ~63LL;
  }
  return cap_tile;
}
static int max_frames_in_budget(const ruf_context *c)
{
  static const double gb = [] { const char *e = getenv("RUF_WORKSPACE_GB"); const double v = e ? atof(e) : 0.0; return v > 0.0 ? v : 48.0; }();
  const long long cap_big = c->want_big > 0 ? c->want_big : 1024;
  const double per_frame = (double)sizeof(BinRec) * (double)c->dims.ntiles * (double)default_cap_tile(c) + (double)sizeof(TriRec) * (double)cap_big +
                           (double)(c->n_parts + 1) * 65.0 + (double)c->dims.ntiles * 24.0 + 16.0;
  const double n = gb * 1e9 / per_frame;
  return n < 1.0 ? 1 : (n > 65535.0 ? 65535 : (int)n);
}

static int ensure_workspace(ruf_context *c, int frames)
{
  if (!c->have_model) return fail(c, RUF_ERR_NO_MODEL, "no model loaded (ruf_set_model)");
  if (frames < c->want_batch) frames = c->want_batch;
  const int fit = max_frames_in_budget(c);
  if (frames > fit) frames = fit;
  // capacities: every tile of every frame owns a record list of cap_tile entries (a robot seen from its own
  // head concentrates its triangles in a third of the tiles: the default is 8x the mean, at least 2048), plus
  // the per-frame big list; both grow by doubling after an overflow (check_status)
  const long long cap_tile = default_cap_tile(c);
  long long cap_big = c->want_big > 0 ? c->want_big : 1024;
  if (cap_big > 0x7fffffffLL || cap_tile > 0x7fffffffLL) return fail(c, RUF_ERR_INVALID, "capacity too large");
  if (c->max_batch >= frames && c->dims.cap_big == (uint32_t)cap_big && c->dims.cap_tile == (uint32_t)cap_tile)
    return RUF_OK;
  RUF_CUDA(c, cudaStreamSynchronize(c->stream));
  free_workspace(c);
  c->dims.cap_big = (uint32_t)cap_big;
  c->dims.cap_tile = (uint32_t)cap_tile;
  c->dims.n_meshlets = c->n_meshlets;
  c->dims.ctr_stride = kCtrWords + 2 * c->dims.ntiles;
  static_assert(kPartStride == kPartStrideHost, "part table stride");
  const size_t f = (size_t)frames;
  RUF_CUDA(c, cudaMalloc(&c->ws.mvp, f * (c->n_parts + 1) * 16 * sizeof(float)));
  RUF_CUDA(c, cudaMalloc(&c->ws.vis, f * (c->n_parts + 1)));
  RUF_CUDA(c, cudaMalloc(&c->ws.ctr, f * c->dims.ctr_stride * sizeof(uint32_t)));
  RUF_CUDA(c, cudaMalloc(&c->ws.big, f * cap_big * sizeof(TriRec)));
  RUF_CUDA(c, cudaMalloc(&c->ws.bins, f * c->dims.ntiles * cap_tile * sizeof(BinRec)));
  RUF_CUDA(c, cudaMalloc(&c->ws.tinfo, f * c->dims.ntiles * sizeof(uint4)));
  c->max_batch = frames;
  return RUF_OK;
}

static size_t elem_size(int enc) { return enc == RUF_ENC_U16_MM ? 2 : 4; }
// bytes of one frame's mask in the context's mask format
static size_t mask_bytes(const ruf_context *c) { const size_t px = (size_t)c->W * c->H; return c->mask_format == RUF_MASK_BITS ? px / 8 : px; }

static int ensure_staging(ruf_context *c, int frames)
{
  if (c->stage_frames >= frames) return RUF_OK;
  RUF_CUDA(c, cudaDeviceSynchronize());
  free_staging(c);
  const size_t px = (size_t)c->W * c->H * frames;
  const size_t mats = (size_t)(16 + 16 * (size_t)frames * (1 + c->n_parts)) * sizeof(double);
  for (int i = 0; i < 2; ++i) {
    RUF_CUDA(c, cudaMalloc(&c->d_in[i], px * 4));
    RUF_CUDA(c, cudaMalloc(&c->d_out[i], px * 4));
    RUF_CUDA(c, cudaMalloc(&c->d_mask[i], px));
    RUF_CUDA(c, cudaMalloc(&c->d_mats[i], mats));
    RUF_CUDA(c, cudaHostAlloc(&c->h_mats[i], mats, cudaHostAllocDefault));
  }
  c->stage_frames = frames;
  return RUF_OK;
}

static ShaderParams shader_params(const ruf_context *c, float max_diff, float replace_value)
{
  // to_linear_depth of include/shaders/urdf_filter.frag:14-17 with the float uniforms z_near/z_far
  const float zn = (float)c->z_near, zf = (float)c->z_far;
  volatile float a = zn * zf, b = zn - zf, e = zf - zn;
  ShaderParams sp;
  sp.k1 = a / b;
  sp.k2 = zf / e;
  sp.max_diff = max_diff;
  sp.replace_value = replace_value;
  return sp;
}

static int launch(ruf_context *c, int n_frames, const void *d_in, int enc, const double *d_proj,
                  const double *d_view, const double *d_model, float max_diff, float replace_value,
                  void *d_out, uint8_t *d_mask, float *d_zbuf, cudaStream_t s, cudaEvent_t depth_ready = nullptr)
{
  FrameBuffers fb;
  fb.depth_in = d_in; fb.depth_out = d_out; fb.mask_out = d_mask; fb.zbuf_out = d_zbuf;
  const uintptr_t al = (uintptr_t)d_in | (uintptr_t)d_out | (uintptr_t)d_zbuf;
  fb.mask_bits = c->mask_format == RUF_MASK_BITS;
  fb.host_status = c->launch_host_status;
  // 8-pixel vector accesses: 16 bytes (16UC1) / 32 bytes (32FC1: 256-bit loads and stores) per thread and row
  const uintptr_t al_io = (uintptr_t)d_in | (uintptr_t)d_out;
  fb.vec_ok = (c->W % 8 == 0) && ((al & 15) == 0) && (enc == RUF_ENC_U16_MM || (al_io & 31) == 0) &&
              (fb.mask_bits || ((uintptr_t)d_mask & 7) == 0);
  if (fb.mask_bits && d_mask && !fb.vec_ok)
    return fail(c, RUF_ERR_INVALID, "RUF_MASK_BITS needs an image width that is a multiple of 8 and 16-byte aligned depth buffers");
  const ShaderParams sp = shader_params(c, max_diff, replace_value);
  Model m{c->meshlets, c->mverts, c->mtris, c->part_aabb};
  // launches that cannot fill the machine with one CTA per (meshlet, run of frames) take the fine cut of the model
  const bool fine = c->n_meshlets_fine > 0 && (c->fine_mode >= 0 ? c->fine_mode != 0 : (long long)n_frames * c->n_meshlets <= kFineMaxCtas);
  if (fine) m.meshlets += c->n_meshlets;
  c->dims.n_meshlets = fine ? c->n_meshlets_fine : c->n_meshlets;
  // Raster kernel variant of this launch.  Automatic: the share of kept records that span more than kMaxUnits raster
  // units, as counted by the setup kernel in the launches whose statistics have reached the host by now (they ride on
  // the status word's read-back, one launch or more behind: a heuristic, both variants give identical results).
  if (c->multipass_mode == 0 || c->multipass_mode == 1) c->multipass = c->multipass_mode == 1;
  else if (c->wide_share >= 0.0) c->multipass = c->wide_share > (c->multipass ? 0.10 : 0.15);     // with hysteresis
  c->dims.multipass = c->multipass ? 1 : 0;
  // Launches too small to fill the machine (a single frame is 80 tiles at 640x480, a third of them busy): the cluster-split
  // raster variant.  RUF_CLUSTER=0 / 1 forces it off / on (A/B, tests).
  c->dims.cluster_split = c->cluster_mode >= 0 ? c->cluster_mode : ((long long)n_frames * c->dims.ntiles <= kClusterMaxTiles ? 1 : 0);
  int launches = 0;
  cudaEvent_t *ev = nullptr;
  if (c->profiling) {
    if (c->ev_used + kNumStages + 1 > c->ev_pool.size()) {
      for (int i = 0; i < kNumStages + 1; ++i) {
        cudaEvent_t x;
        if (cudaEventCreate(&x) != cudaSuccess) return fail(c, RUF_ERR_CUDA, "cudaEventCreate failed");
        c->ev_pool.push_back(x);
      }
    }
    ev = c->ev_pool.data() + c->ev_used;
    c->ev_used += kNumStages + 1;
  }
  const int S = c->slice_frames;
  if (!ev && S > 0 && n_frames >= 2 * S) {
    // fork: slices alternate between the two auxiliary streams; join back into `s`
    if (!c->s_aux[0]) {
      for (int i = 0; i < 2; ++i) {
        RUF_CUDA(c, cudaStreamCreateWithFlags(&c->s_aux[i], cudaStreamNonBlocking));
        RUF_CUDA(c, cudaEventCreateWithFlags(&c->ev_join[i], cudaEventDisableTiming));
      }
      RUF_CUDA(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    }
    RUF_CUDA(c, cudaEventRecord(c->ev_fork, s));
    for (int i = 0; i < 2; ++i) RUF_CUDA(c, cudaStreamWaitEvent(c->s_aux[i], c->ev_fork, 0));
    const size_t es = (enc == RUF_ENC_U16_MM) ? 2 : 4, img = (size_t)c->W * c->H;
    const int P = c->n_parts;
    int k = 0;
    for (int f0 = 0; f0 < n_frames; f0 += S, ++k) {
      const int nf = (n_frames - f0 < S) ? (n_frames - f0) : S;
      Workspace ws = c->ws;
      ws.mvp += (size_t)f0 * (P + 1) * 16;
      ws.vis += (size_t)f0 * (P + 1);
      ws.ctr += (size_t)f0 * c->dims.ctr_stride;
      ws.big += (size_t)f0 * c->dims.cap_big;
      ws.bins += (size_t)f0 * c->dims.ntiles * c->dims.cap_tile;
      ws.tinfo += (size_t)f0 * c->dims.ntiles;
      FrameBuffers fs = fb;
      fs.depth_in = (const char *)d_in + f0 * img * es;
      fs.depth_out = (char *)d_out + f0 * img * es;
      fs.mask_out = d_mask ? d_mask + f0 * mask_bytes(c) : nullptr;
      fs.zbuf_out = d_zbuf ? d_zbuf + f0 * img : nullptr;
      int l = 0;
      cudaError_t e = launch_frames(c->dims, m, ws, nf, d_proj, d_view + 16 * (size_t)f0, d_model + 16 * (size_t)f0 * P,
                                    c->d_lookat, enc, sp, fs, c->s_aux[k & 1], &l, nullptr);
      if (e != cudaSuccess) return fail(c, RUF_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
      launches += l;
    }
    for (int i = 0; i < 2; ++i) {
      RUF_CUDA(c, cudaEventRecord(c->ev_join[i], c->s_aux[i]));
      RUF_CUDA(c, cudaStreamWaitEvent(s, c->ev_join[i], 0));
    }
  } else {
    cudaError_t e = launch_frames(c->dims, m, c->ws, n_frames, d_proj, d_view, d_model, c->d_lookat, enc,
                                  sp, fb, s, &launches, ev, depth_ready);
    if (e != cudaSuccess) return fail(c, RUF_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
  }
  c->stats.kernel_launches += launches;
  return RUF_OK;
}

extern "C" {

int ruf_create(ruf_context **out, int device, int width, int height, double z_near, double z_far)
{
  if (!out) return fail(nullptr, RUF_ERR_INVALID, "ctx is NULL");
  *out = nullptr;
  if (width < 1 || height < 1 || width > 4096 || height > 4096)
    return fail(nullptr, RUF_ERR_INVALID, "image size %dx%d outside 1..4096", width, height);
  if (!(z_near > 0.0) || !(z_far > z_near)) return fail(nullptr, RUF_ERR_INVALID, "need 0 < z_near < z_far");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(nullptr, RUF_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= count) return fail(nullptr, RUF_ERR_INVALID, "device %d out of range", device);
  ruf_context *c = new (std::nothrow) ruf_context;
  if (!c) return fail(nullptr, RUF_ERR_NOMEM, "out of host memory");
  c->device = device; c->W = width; c->H = height; c->z_near = z_near; c->z_far = z_far;
  auto bail = [&](const char *what, cudaError_t err) {
    int rc = fail(nullptr, RUF_ERR_CUDA, "%s: %s", what, cudaGetErrorString(err));
    ruf_destroy(c);
    return rc;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = check_kernel_image()) != cudaSuccess)
    return bail("no sm_100a kernel image for this device (libruf_b200 targets B200 only)", e);
  if ((e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&c->s_in, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&c->s_out, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  c->stream = c->own_stream;
  if (const char *e3 = getenv("RUF_SETUP_FRAMES_FORCE")) {   // testing aid: frames per setup CTA
    const int v = atoi(e3);
    if (v >= 1 && v <= 64) c->dims.force_fpc = v;
  }
  if (getenv("RUF_NO_GRAPH")) c->use_graph = false;
  if (const char *e = getenv("RUF_CLUSTER")) c->cluster_mode = atoi(e) ? 1 : 0;
  if (const char *e = getenv("RUF_FINE_MESHLETS")) c->fine_mode = atoi(e) ? 1 : 0;
  if (const char *e = getenv("RUF_DIRECT")) c->direct_mode = atoi(e) & 7;
  if (const char *e = getenv("RUF_BG_CACHE")) c->bg_cache = atoi(e) != 0;
  if (const char *e = getenv("RUF_HOST_STAGING")) c->host_staging = atoi(e) != 0;
  if (const char *e2 = getenv("RUF_SLICE_FRAMES")) {    // tuning aid
    const int v = atoi(e2);
    if (v >= 0 && v <= 65535) c->slice_frames = v;
  }
  for (int i = 0; i < 2; ++i) {
    if ((e = cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
    if ((e = cudaEventCreateWithFlags(&c->ev_k[i], cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
    if ((e = cudaEventCreateWithFlags(&c->ev_out[i], cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
  }
  if ((e = cudaMalloc(&c->ws.status, 4 * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMemset(c->ws.status, 0, 4 * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMemset", e);
  if ((e = cudaHostAlloc(&c->h_status, 4 * sizeof(uint32_t), cudaHostAllocDefault)) != cudaSuccess) return bail("cudaHostAlloc", e);
  c->h_status[0] = c->h_status[1] = c->h_status[2] = c->h_status[3] = 0;
  if (const char *e4 = getenv("RUF_MULTIPASS")) c->multipass_mode = atoi(e4);    // 0 / 1 force a variant, anything else = automatic
  double la[16];
  ruf_lookat(la);
  if ((e = cudaMalloc(&c->d_lookat, sizeof(la))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMemcpy(c->d_lookat, la, sizeof(la), cudaMemcpyHostToDevice)) != cudaSuccess) return bail("cudaMemcpy", e);

  Dims &d = c->dims;
  d.W = width; d.H = height;
  d.tiles_x = (width + kTileW - 1) / kTileW;
  d.tiles_y = (height + kTileH - 1) / kTileH;
  d.ntiles = d.tiles_x * d.tiles_y;
  d.halfw = 0.5f * (float)width;
  d.halfh = 0.5f * (float)height;
  d.guard_x = kGuardPx / d.halfw;
  d.guard_y = kGuardPx / d.halfh;
  d.n_tris = -1;
  if (d.ntiles > kMaxTiles || d.tiles_x > 256 || d.tiles_y > 256) {
    fail(nullptr, RUF_ERR_INVALID, "image of %dx%d needs %d tiles (limit %d)", width, height, d.ntiles, kMaxTiles);
    ruf_destroy(c);
    return RUF_ERR_INVALID;
  }
  *out = c;
  return RUF_OK;
}

int ruf_destroy(ruf_context *c)
{
  if (!c) return RUF_OK;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  free_workspace(c);
  free_staging(c);
  cudaFree(c->meshlets); cudaFree(c->mverts); cudaFree(c->mtris); cudaFree(c->part_aabb);
  cudaFree(c->kin_blob); cudaFree(c->fk_links); cudaFree(c->fk_pm); cudaFree(c->fk_view);
  cudaFree(c->ws.status); cudaFree(c->d_lookat); cudaFree(c->bg_seed);
  cudaFreeHost(c->hs_in); cudaFreeHost(c->hs_out); cudaFreeHost(c->hs_mask);
  if (c->h_status) cudaFreeHost(c->h_status);
  for (int i = 0; i < 2; ++i) {
    if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
    if (c->ev_k[i]) cudaEventDestroy(c->ev_k[i]);
    if (c->ev_out[i]) cudaEventDestroy(c->ev_out[i]);
  }
  for (cudaEvent_t x : c->ev_pool) cudaEventDestroy(x);
  for (int i = 0; i < 2; ++i) {
    if (c->s_aux[i]) cudaStreamDestroy(c->s_aux[i]);
    if (c->ev_join[i]) cudaEventDestroy(c->ev_join[i]);
  }
  if (c->fg.exec) cudaGraphExecDestroy(c->fg.exec);
  if (c->fg.graph) cudaGraphDestroy(c->fg.graph);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  if (c->s_in) cudaStreamDestroy(c->s_in);
  if (c->s_out) cudaStreamDestroy(c->s_out);
  delete c;
  return RUF_OK;
}

const char *ruf_last_error(const ruf_context *c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int ruf_set_stream(ruf_context *c, void *cuda_stream)
{
  if (!c) return RUF_ERR_INVALID;
  c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
  return RUF_OK;
}

// the launch statistics that ride on the status word: share of kept records spanning more than kMaxUnits raster units
static void take_launch_stats(ruf_context *c)
{
  if (c->h_status[1] >= 256u) c->wide_share = (double)c->h_status[2] / (double)c->h_status[1];
}

static int check_status(ruf_context *c, cudaStream_t s)
{
  RUF_CUDA(c, cudaMemcpyAsync(c->h_status, c->ws.status, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  RUF_CUDA(c, cudaMemsetAsync(c->ws.status + 1, 0, 2 * sizeof(uint32_t), s));      // statistics restart with every read-back
  RUF_CUDA(c, cudaStreamSynchronize(s));
  take_launch_stats(c);
  const uint32_t flags = *c->h_status;
  if (flags) {
    RUF_CUDA(c, cudaMemsetAsync(c->ws.status, 0, sizeof(uint32_t), s));
    // grow so that the caller's retry fits
    if (flags & kFlagBigOverflow) c->want_big = 4LL * c->dims.cap_big;
    if (flags & kFlagBinOverflow) c->want_bin = 2LL * c->dims.cap_tile;
    return fail(c, RUF_ERR_OVERFLOW, "internal %s buffer overflow (capacity raised for the next call)",
                (flags & kFlagBigOverflow) ? "big-list" : "tile-reference");
  }
  return RUF_OK;
}

int ruf_sync(ruf_context *c)
{
  if (!c) return RUF_ERR_INVALID;
  RUF_CUDA(c, cudaSetDevice(c->device));
  return check_status(c, c->stream);
}

// Device buffers for a meshlet model of the given sizes (contents come from upload_model or from an ncclBroadcast)
static int alloc_model(ruf_context *c, const MeshletModel &mm, int64_t n_tris, int n_parts)
{
  if (mm.verts.size() / 4 > 0xffffffffull || mm.tris.size() > 0xffffffffull || mm.n_meshlets() > 0x7fffffffull)
    return fail(c, RUF_ERR_INVALID, "model too large");
  RUF_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFree(c->meshlets); cudaFree(c->mverts); cudaFree(c->mtris); cudaFree(c->part_aabb);
  c->meshlets = nullptr; c->mverts = nullptr; c->mtris = nullptr; c->part_aabb = nullptr;
  c->have_model = false;
  RUF_CUDA(c, cudaMalloc(&c->meshlets, mm.hdr.size() * sizeof(uint32_t)));
  RUF_CUDA(c, cudaMalloc(&c->mverts, mm.verts.size() * sizeof(float)));
  RUF_CUDA(c, cudaMalloc(&c->mtris, mm.tris.size() * sizeof(uint32_t)));
  RUF_CUDA(c, cudaMalloc(&c->part_aabb, mm.part_aabb.size() * sizeof(float)));
  c->n_meshlets = (int)mm.n_primary;
  c->n_meshlets_fine = (int)(mm.n_meshlets() - mm.n_primary);
  c->n_tris = n_tris; c->n_parts = n_parts;
  c->dims.n_tris = n_tris; c->dims.n_parts = n_parts;
  c->have_model = true;
  c->bg_state = 0;
  free_workspace(c);
  free_staging(c);
  c->want_big = c->want_bin = 0;
  cudaFree(c->kin_blob);       // kinematics belong to the previous model
  c->kin_blob = nullptr;
  return RUF_OK;
}
static int upload_model(ruf_context *c, const MeshletModel &mm)
{
  RUF_CUDA(c, cudaMemcpy(c->meshlets, mm.hdr.data(), mm.hdr.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  RUF_CUDA(c, cudaMemcpy(c->mverts, mm.verts.data(), mm.verts.size() * sizeof(float), cudaMemcpyHostToDevice));
  RUF_CUDA(c, cudaMemcpy(c->mtris, mm.tris.data(), mm.tris.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  RUF_CUDA(c, cudaMemcpy(c->part_aabb, mm.part_aabb.data(), mm.part_aabb.size() * sizeof(float), cudaMemcpyHostToDevice));
  return RUF_OK;
}
// Model ingest (set-up time): the soup is cut into meshlets on the host (ruf_meshlet.cpp) and uploaded.
static int build_model(ruf_context *c, const float *h_xyz, const uint32_t *h_part, int64_t n_tris, int n_parts)
{
  MeshletModel mm;
  // glVertex3f(.., far_plane_*0.99), src/urdf_filter.cpp:592
  build_meshlet_sets(h_xyz, h_part, n_tris, n_parts, (float)(c->z_far * 0.99), kMeshVerts, kMeshTris, kMeshTrisFine, kMeshParts, mm);
  const int rc = alloc_model(c, mm, n_tris, n_parts);
  return rc != RUF_OK ? rc : upload_model(c, mm);
}

int ruf_set_model(ruf_context *c, const float *tri_xyz, const uint32_t *tri_part, int64_t n_tris, int n_parts)
{
  if (!c) return RUF_ERR_INVALID;
  if (n_tris < 0 || n_parts < 0 || n_parts > (1 << 20) || (n_tris > 0 && (!tri_xyz || !tri_part)) || n_tris > (1LL << 30))
    return fail(c, RUF_ERR_INVALID, "bad model arguments");
  for (int64_t t = 0; t < n_tris; ++t)
    if (tri_part[t] >= (uint32_t)n_parts) return fail(c, RUF_ERR_INVALID, "tri_part[%lld] = %u >= n_parts", (long long)t, tri_part[t]);
  RUF_CUDA(c, cudaSetDevice(c->device));
  return build_model(c, tri_xyz, tri_part, n_tris, n_parts);
}

int ruf_set_model_device(ruf_context *c, const void *d_tri_xyz, const void *d_tri_part, int64_t n_tris, int n_parts)
{
  if (!c) return RUF_ERR_INVALID;
  if (n_tris < 0 || n_parts < 0 || n_parts > (1 << 20) || (n_tris > 0 && (!d_tri_xyz || !d_tri_part)) || n_tris > (1LL << 30))
    return fail(c, RUF_ERR_INVALID, "bad model arguments");
  RUF_CUDA(c, cudaSetDevice(c->device));
  // set-up time only: the soup comes to the host once, where the meshlets are built
  std::vector<float> h_xyz((size_t)n_tris * 9);
  std::vector<uint32_t> h_part((size_t)n_tris);
  if (n_tris > 0) {
    RUF_CUDA(c, cudaMemcpy(h_xyz.data(), d_tri_xyz, h_xyz.size() * sizeof(float), cudaMemcpyDeviceToHost));
    RUF_CUDA(c, cudaMemcpy(h_part.data(), d_tri_part, h_part.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  }
  for (int64_t t = 0; t < n_tris; ++t)
    if (h_part[t] >= (uint32_t)n_parts) return fail(c, RUF_ERR_INVALID, "tri_part[%lld] = %u >= n_parts", (long long)t, h_part[t]);
  return build_model(c, h_xyz.data(), h_part.data(), n_tris, n_parts);
}

int ruf_reserve(ruf_context *c, int max_batch, int64_t rec_capacity, int64_t bin_capacity)
{
  if (!c || max_batch < 1 || max_batch > 65535) return c ? fail(c, RUF_ERR_INVALID, "max_batch out of range") : RUF_ERR_INVALID;
  RUF_CUDA(c, cudaSetDevice(c->device));
  c->want_batch = max_batch;
  c->want_big = rec_capacity; c->want_bin = bin_capacity;
  if (c->max_batch > max_batch) free_workspace(c);
  return ensure_workspace(c, max_batch);
}

int ruf_filter_batch_device(ruf_context *c, int n_frames, const void *d_depth_in, int enc, const double *d_proj,
                            const double *d_view, const double *d_part_model, float max_diff, float replace_value,
                            void *d_depth_out, uint8_t *d_mask_out, float *d_zbuf_out)
{
  if (!c) return RUF_ERR_INVALID;
  if (n_frames < 1 || !d_depth_in || !d_depth_out || !d_proj || !d_view || (c->n_parts > 0 && !d_part_model) ||
      (enc != RUF_ENC_F32_M && enc != RUF_ENC_U16_MM))
    return fail(c, RUF_ERR_INVALID, "bad arguments");
  if (!c->have_model) return fail(c, RUF_ERR_NO_MODEL, "no model loaded (ruf_set_model)");
  RUF_CUDA(c, cudaSetDevice(c->device));
  if (n_frames > 65535) return fail(c, RUF_ERR_INVALID, "n_frames > 65535");
  int rc = ensure_workspace(c, n_frames);
  if (rc != RUF_OK) return rc;
  c->stats = ruf_stats{};
  c->stats.frames = n_frames;
  // more frames than the workspace budget holds: equal sub-batches, stream-ordered through the same workspace
  const int n_sub = (n_frames + c->max_batch - 1) / c->max_batch;
  const int per = (n_frames + n_sub - 1) / n_sub;
  const size_t es = elem_size(enc), img = (size_t)c->W * c->H;
  for (int f0 = 0; f0 < n_frames; f0 += per) {
    const int nf = (n_frames - f0 < per) ? (n_frames - f0) : per;
    c->last_frames = nf;
    rc = launch(c, nf, (const char *)d_depth_in + f0 * img * es, enc, d_proj, d_view + 16 * (size_t)f0,
                d_part_model ? d_part_model + 16 * (size_t)f0 * c->n_parts : nullptr, max_diff, replace_value,
                (char *)d_depth_out + f0 * img * es, d_mask_out ? d_mask_out + f0 * mask_bytes(c) : nullptr,
                d_zbuf_out ? d_zbuf_out + f0 * img : nullptr, c->stream);
    if (rc != RUF_OK) return rc;
  }
  return RUF_OK;
}

// One pass of the chunked host pipeline.  Chunk k: H2D on s_in (slot k&1) -> kernels on the
// context stream -> D2H on s_out.  Slots are recycled under event dependencies.
static int host_pipeline(ruf_context *c, int n_frames, const void *depth_in, int enc, const double *proj,
                         const double *view, const double *part_model, float max_diff, float replace_value,
                         void *depth_out, uint8_t *mask_out, const std::vector<std::pair<int, int>> &segs,
                         bool copy_only = false)
{
  (void)n_frames;
  const size_t es = elem_size(enc);
  const size_t img = (size_t)c->W * c->H;
  const int P = c->n_parts;
  cudaStream_t sk = c->stream;
  for (int k = 0; k < (int)segs.size(); ++k) {
    const int f0 = segs[k].first, nf = segs[k].second;
    const int slot = k & 1;
    if (k >= 2) {
      // slot reuse: input slot free once chunk k-2's kernels ran; pinned matrices likewise
      RUF_CUDA(c, cudaStreamWaitEvent(c->s_in, c->ev_k[slot], 0));
      RUF_CUDA(c, cudaEventSynchronize(c->ev_in[slot]));   // h_mats[slot] was consumed by its H2D
    }
    double *hm = c->h_mats[slot];
    std::memcpy(hm, proj, 16 * sizeof(double));
    std::memcpy(hm + 16, view + 16 * (size_t)f0, 16 * sizeof(double) * nf);
    if (P > 0) std::memcpy(hm + 16 + 16 * (size_t)nf, part_model + 16 * (size_t)f0 * P, 16 * sizeof(double) * (size_t)nf * P);
    const size_t mat_bytes = (16 + 16 * (size_t)nf * (1 + P)) * sizeof(double);
    RUF_CUDA(c, cudaMemcpyAsync(c->d_mats[slot], hm, mat_bytes, cudaMemcpyHostToDevice, c->s_in));
    RUF_CUDA(c, cudaMemcpyAsync(c->d_in[slot], (const char *)depth_in + f0 * img * es, nf * img * es,
                                cudaMemcpyHostToDevice, c->s_in));
    RUF_CUDA(c, cudaEventRecord(c->ev_in[slot], c->s_in));
    c->stats.h2d_bytes += (int64_t)(mat_bytes + nf * img * es);

    RUF_CUDA(c, cudaStreamWaitEvent(sk, c->ev_in[slot], 0));
    if (k >= 2) RUF_CUDA(c, cudaStreamWaitEvent(sk, c->ev_out[slot], 0));   // output slot drained
    double *dm = c->d_mats[slot];
    int rc = copy_only ? RUF_OK
                       : launch(c, nf, c->d_in[slot], enc, dm, dm + 16, dm + 16 + 16 * (size_t)nf, max_diff, replace_value,
                                c->d_out[slot], mask_out ? c->d_mask[slot] : nullptr, nullptr, sk);
    if (rc != RUF_OK) return rc;
    c->last_frames = nf;                       // ruf_get_stats reads the counters of the last launch sequence
    RUF_CUDA(c, cudaEventRecord(c->ev_k[slot], sk));

    RUF_CUDA(c, cudaStreamWaitEvent(c->s_out, c->ev_k[slot], 0));
    RUF_CUDA(c, cudaMemcpyAsync((char *)depth_out + f0 * img * es, c->d_out[slot], nf * img * es,
                                cudaMemcpyDeviceToHost, c->s_out));
    c->stats.d2h_bytes += (int64_t)(nf * img * es);
    if (mask_out) {
      const size_t mb = mask_bytes(c);
      RUF_CUDA(c, cudaMemcpyAsync(mask_out + f0 * mb, c->d_mask[slot], nf * mb, cudaMemcpyDeviceToHost, c->s_out));
      c->stats.d2h_bytes += (int64_t)(nf * mb);
    }
    RUF_CUDA(c, cudaEventRecord(c->ev_out[slot], c->s_out));
  }
  RUF_CUDA(c, cudaStreamSynchronize(c->s_out));
  return check_status(c, sk);
}

// frames per pipeline chunk: large enough for efficient copies/launches, small enough to overlap (measured)
static int host_chunk(int n_frames)
{
  int chunk = n_frames >= 512 ? 64 : (n_frames >= 128 ? 32 : (n_frames >= 64 ? 16 : (n_frames >= 32 ? 8 : (n_frames >= 8 ? 4 : 1))));
  if (const char *e = getenv("RUF_HOST_CHUNK")) {      // tuning aid
    const int v = atoi(e);
    if (v >= 1 && v <= 4096) chunk = v < n_frames ? v : n_frames;
  }
  return chunk;
}

// The frames of one host call as pipeline segments (first frame, count).  The D2H copies are the bottleneck (2+1 bytes
// out vs 2 bytes in per pixel) and run back to back once they have started, so what the chunking can still save is the
// time until the first one starts: the first segments are small (chunk/8, chunk/8, chunk/4, chunk/2), then full size.
static std::vector<std::pair<int, int>> ramp_segments(int n_frames, int chunk)
{
  std::vector<std::pair<int, int>> segs;
  int f0 = 0;
  for (int k = 0; f0 < n_frames; ++k) {
    int want = chunk;
    if (k < 2) want = chunk / 8; else if (k == 2) want = chunk / 4; else if (k == 3) want = chunk / 2;
    if (want < 1) want = 1;
    const int nf = (n_frames - f0 < want) ? (n_frames - f0) : want;
    segs.emplace_back(f0, nf);
    f0 += nf;
  }
  return segs;
}

int ruf_filter_batch_host(ruf_context *c, int n_frames, const void *depth_in, int enc, const double *proj,
                          const double *view, const double *part_model, float max_diff, float replace_value,
                          void *depth_out, uint8_t *mask_out)
{
  if (!c) return RUF_ERR_INVALID;
  if (n_frames < 1 || !depth_in || !depth_out || !proj || !view || (c->n_parts > 0 && !part_model) ||
      (enc != RUF_ENC_F32_M && enc != RUF_ENC_U16_MM))
    return fail(c, RUF_ERR_INVALID, "bad arguments");
  if (!c->have_model) return fail(c, RUF_ERR_NO_MODEL, "no model loaded (ruf_set_model)");
  RUF_CUDA(c, cudaSetDevice(c->device));
  const int chunk = host_chunk(n_frames);
  for (int attempt = 0; attempt < 8; ++attempt) {
    int rc = ensure_workspace(c, chunk);
    if (rc != RUF_OK) return rc;
    rc = ensure_staging(c, chunk);
    if (rc != RUF_OK) return rc;
    c->stats = ruf_stats{};
    c->stats.frames = n_frames;
    c->last_frames = (n_frames % chunk) ? (n_frames % chunk) : chunk;
    rc = host_pipeline(c, n_frames, depth_in, enc, proj, view, part_model, max_diff, replace_value, depth_out,
                       mask_out, ramp_segments(n_frames, chunk));
    if (rc != RUF_ERR_OVERFLOW) return rc;   // overflow: capacities were doubled, run again
  }
  return fail(c, RUF_ERR_OVERFLOW, "internal buffers still too small after 8 attempts");
}

int ruf_host_copy_ceiling(ruf_context *c, int n_frames, const void *depth_in, int enc, const double *proj,
                          const double *view, const double *part_model, void *depth_out, uint8_t *mask_out)
{
  if (!c) return RUF_ERR_INVALID;
  if (n_frames < 1 || !depth_in || !depth_out || !proj || !view || (c->n_parts > 0 && !part_model) ||
      (enc != RUF_ENC_F32_M && enc != RUF_ENC_U16_MM))
    return fail(c, RUF_ERR_INVALID, "bad arguments");
  if (!c->have_model) return fail(c, RUF_ERR_NO_MODEL, "no model loaded (ruf_set_model)");
  RUF_CUDA(c, cudaSetDevice(c->device));
  const int chunk = host_chunk(n_frames);
  int rc = ensure_workspace(c, chunk);
  if (rc != RUF_OK) return rc;
  rc = ensure_staging(c, chunk);
  if (rc != RUF_OK) return rc;
  c->stats = ruf_stats{};
  c->stats.frames = n_frames;
  return host_pipeline(c, n_frames, depth_in, enc, proj, view, part_model, 0.0f, 0.0f, depth_out, mask_out,
                       ramp_segments(n_frames, chunk), true);
}

// ---- single frame, pinned host buffers: one CUDA graph per context ----------------------------------------------
static bool is_pinned_host(const void *p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// The device's address of a pinned host buffer the kernels may read or write themselves (zero copy), or null
static void *mapped_alias(const void *p)
{
  void *d = nullptr;
  if (!p || cudaHostGetDevicePointer(&d, const_cast<void *>(p), 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return d;
}

// The background quad (src/urdf_filter.cpp:591-596) is drawn with MODELVIEW = LookAt, a constant: what the clipper and the
// set-up make of it depends on the projection matrix alone.  A 30 Hz caller passes the same camera_info every frame, and on
// a launch of ONE frame the lone warp that clips the quad again (~8 us) is what the setup kernel ends with.  So the
// single-frame graph seeds the frame's big list with records set up once per projection matrix: pose kernel + setup
// kernel run on the quad's meshlet alone, the same code on the same operands, hence the same bits.
static int fill_bg_seed(ruf_context *c, const double *proj, double *hm, size_t mat_bytes)
{
  cudaStream_t sk = c->stream;
  c->bg_state = 0;
  if (!c->bg_seed) RUF_CUDA(c, cudaMalloc(&c->bg_seed, (4 + kBgSeedMax * sizeof(TriRec) / 4) * sizeof(uint32_t)));
  RUF_CUDA(c, cudaMemcpyAsync(c->d_mats[0], hm, mat_bytes, cudaMemcpyHostToDevice, sk));
  double *dm = c->d_mats[0];
  Model m{c->meshlets, c->mverts, c->mtris, c->part_aabb};      // (the throughput cut; the quad is its last meshlet too)
  Dims d = c->dims;
  d.n_meshlets = c->n_meshlets; d.fold_clear = 0; d.bg_mode = kBgOnly; d.cluster_split = 0;
  ShaderParams sp{};
  FrameBuffers fb{};
  cudaError_t e = launch_frames(d, m, c->ws, 1, dm, dm + 16, dm + 32, c->d_lookat, RUF_ENC_U16_MM, sp, fb, sk, nullptr, nullptr, nullptr);
  if (e != cudaSuccess) return fail(c, RUF_ERR_CUDA, "background set-up failed: %s", cudaGetErrorString(e));
  uint32_t w[kCtrWords] = {0};
  RUF_CUDA(c, cudaMemcpyAsync(w, c->ws.ctr, sizeof(w), cudaMemcpyDeviceToHost, sk));
  RUF_CUDA(c, cudaStreamSynchronize(sk));
  const uint32_t n = w[kCtrBig];
  // Only a quad that ends up in the big list and nowhere else can be seeded: on small images it lies inside the guard band,
  // is not clipped, touches few tiles and is BINNED like any other triangle (kept != 0) -- then the setup kernel keeps it.
  if (n > (uint32_t)kBgSeedMax || n > c->dims.cap_big || w[kCtrKept] != 0 || w[kCtrFlags] != 0) { c->bg_state = -1; return RUF_OK; }
  RUF_CUDA(c, cudaMemcpyAsync(c->bg_seed, &n, sizeof(uint32_t), cudaMemcpyHostToDevice, sk));
  if (n) RUF_CUDA(c, cudaMemcpyAsync(c->bg_seed + 4, c->ws.big, n * sizeof(TriRec), cudaMemcpyDeviceToDevice, sk));
  RUF_CUDA(c, cudaStreamSynchronize(sk));
  std::memcpy(c->bg_proj, proj, sizeof(c->bg_proj));
  c->bg_state = 1;
  return RUF_OK;
}

// RUF_OK: done.  1: not applicable (pageable buffers, profiling) or overflow -> take the pipeline.
//
// Latency path.  Everything a 30 Hz caller waits for is on the critical path of ONE frame, so the graph holds as few nodes
// as the buffers allow.  With mapped pinned buffers (cudaHostAlloc / ruf_host_alloc: the default) it is three kernel nodes
// and nothing else: the pose kernel reads the matrices from the context's pinned block and clears the frame's counters,
// the raster kernel (cluster-split variant) reads the depth image from and writes depth + mask to the caller's buffers
// over PCIe while it works, and its last CTA stores the status words into the host's copy.  Buffers that are pinned but
// not mapped get copy nodes instead (depth upload beside the pose / setup kernels, mask read-back beside the depth
// read-back).  RUF_DIRECT = bit mask 1 outputs, 2 input, 4 matrices + status (default 7) for A/B runs.
static int single_frame_graph(ruf_context *c, const void *depth_in, int enc, const double *proj, const double *view,
                              const double *part_model, float max_diff, float replace_value, void *depth_out,
                              uint8_t *mask_out)
{
  if (!c->use_graph || c->profiling || c->slice_frames) return 1;
  if (!is_pinned_host(depth_in) || !is_pinned_host(depth_out) || (mask_out && !is_pinned_host(mask_out))) return 1;
  int rc = ensure_workspace(c, 1);
  if (rc != RUF_OK) return rc;
  rc = ensure_staging(c, 1);
  if (rc != RUF_OK) return rc;
  const size_t es = elem_size(enc), img = (size_t)c->W * c->H, mb = mask_bytes(c);
  const int P = c->n_parts;
  cudaStream_t sk = c->stream;
  double *hm = c->h_mats[0];                      // free: every call of this path ends with a synchronisation
  std::memcpy(hm, proj, 16 * sizeof(double));
  std::memcpy(hm + 16, view, 16 * sizeof(double));
  if (P > 0) std::memcpy(hm + 32, part_model, 16 * sizeof(double) * (size_t)P);
  const size_t mat_bytes = (32 + 16 * (size_t)P) * sizeof(double);
  ruf_context::FrameGraph &g = c->fg;
  const bool want_mp = c->multipass_mode == 0 || c->multipass_mode == 1 ? c->multipass_mode == 1
                       : (c->wide_share >= 0.0 ? c->wide_share > (c->multipass ? 0.10 : 0.15) : c->multipass);
  int direct = c->direct_mode;
  void *dev_in = (direct & 2) ? mapped_alias(depth_in) : nullptr;
  void *dev_out = (direct & 1) ? mapped_alias(depth_out) : nullptr;
  void *dev_mask = (direct & 1) && mask_out ? mapped_alias(mask_out) : nullptr;
  if (!dev_in) direct &= ~2;
  if (!dev_out || (mask_out && !dev_mask)) direct &= ~1;
  double *dev_mats = (direct & 4) ? (double *)mapped_alias(hm) : nullptr;
  uint32_t *dev_status = (direct & 4) ? (uint32_t *)mapped_alias(c->h_status) : nullptr;
  // (the status export needs the cluster-split raster variant: its last CTA does it)
  const bool cluster = c->cluster_mode >= 0 ? c->cluster_mode != 0 : c->dims.ntiles <= kClusterMaxTiles;
  if (!dev_mats || !dev_status || !cluster) { direct &= ~4; dev_mats = nullptr; dev_status = nullptr; }
  // the background quad's records for this projection matrix (refilled when the caller's camera_info changes; the graph
  // reads them from the same device buffer, so it stays valid)
  if (c->bg_cache && c->bg_state >= 0 && (c->bg_state == 0 || std::memcmp(c->bg_proj, proj, sizeof(c->bg_proj)) != 0)) {
    rc = fill_bg_seed(c, proj, hm, mat_bytes);
    if (rc != RUF_OK) return rc;
  }
  const int bg_mode = (c->bg_cache && c->bg_state == 1) ? kBgSkip : 0;
  const bool same = g.exec && g.bg_mode == bg_mode && g.direct == direct && g.multipass == (int)want_mp && g.enc == enc && g.mask_format == c->mask_format &&
                    g.n_parts == P && g.max_diff == max_diff && g.replace_value == replace_value && g.ws_bins == c->ws.bins &&
                    g.stage_in == c->d_in[0] && g.cap_big == c->dims.cap_big && g.cap_tile == c->dims.cap_tile &&
                    (g.mask != nullptr) == (mask_out != nullptr);
  if (!same) {
    drop_frame_graph(c);
    if (!c->ev_fork) RUF_CUDA(c, cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    // (captured on the context's own stream -- idle while a caller's stream is in use --, launched on the current one)
    cudaStream_t cap = c->own_stream;
    RUF_CUDA(c, cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed));
    bool ok = true;
    auto CK = [&](cudaError_t e) { if (e != cudaSuccess) ok = false; };
    const bool copy_in = !(direct & 2), copy_out = !(direct & 1);
    if (copy_in) {
      CK(cudaEventRecord(c->ev_fork, cap));
      CK(cudaStreamWaitEvent(c->s_in, c->ev_fork, 0));
      CK(cudaMemcpyAsync(c->d_in[0], depth_in, img * es, cudaMemcpyHostToDevice, c->s_in));    // beside the pose / setup kernels
      CK(cudaEventRecord(c->ev_in[0], c->s_in));
    }
    double *dm = dev_mats ? dev_mats : c->d_mats[0];
    if (!dev_mats) CK(cudaMemcpyAsync(c->d_mats[0], hm, mat_bytes, cudaMemcpyHostToDevice, cap));
    const int64_t launches_before = c->stats.kernel_launches;
    c->dims.fold_clear = 1;
    c->dims.bg_mode = bg_mode;
    c->ws.bg_seed = c->bg_seed;
    c->launch_host_status = dev_status;
    if (ok && launch(c, 1, copy_in ? c->d_in[0] : dev_in, enc, dm, dm + 16, dm + 32, max_diff, replace_value,
                     copy_out ? c->d_out[0] : dev_out, mask_out ? (copy_out ? c->d_mask[0] : (uint8_t *)dev_mask) : nullptr, nullptr, cap,
                     copy_in ? c->ev_in[0] : nullptr) != RUF_OK)
      ok = false;
    c->dims.fold_clear = 0;
    c->dims.bg_mode = 0;
    c->launch_host_status = nullptr;
    c->stats.kernel_launches = launches_before;
    if (copy_out) {
      CK(cudaEventRecord(c->ev_k[0], cap));
      CK(cudaStreamWaitEvent(c->s_out, c->ev_k[0], 0));
      if (mask_out) CK(cudaMemcpyAsync(mask_out, c->d_mask[0], mb, cudaMemcpyDeviceToHost, c->s_out));   // beside the depth read-back
      CK(cudaEventRecord(c->ev_out[0], c->s_out));
      CK(cudaMemcpyAsync(depth_out, c->d_out[0], img * es, cudaMemcpyDeviceToHost, cap));
    }
    if (!dev_status) {
      CK(cudaMemcpyAsync(c->h_status, c->ws.status, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, cap));
      CK(cudaMemsetAsync(c->ws.status + 1, 0, 2 * sizeof(uint32_t), cap));
    }
    if (copy_out) CK(cudaStreamWaitEvent(cap, c->ev_out[0], 0));
    cudaGraph_t graph = nullptr;
    const cudaError_t ee = cudaStreamEndCapture(cap, &graph);
    if (!ok || ee != cudaSuccess || !graph) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      c->use_graph = false;                       // e.g. a driver that refuses the capture: the pipeline still works
      return 1;
    }
    g.graph = graph;
    if (cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) { cudaGetLastError(); drop_frame_graph(c); c->use_graph = false; return 1; }
    size_t n = 0;
    cudaGraphGetNodes(graph, nullptr, &n);
    std::vector<cudaGraphNode_t> nodes(n);
    cudaGraphGetNodes(graph, nodes.data(), &n);
    for (cudaGraphNode_t nd : nodes) {
      cudaGraphNodeType t;
      if (cudaGraphNodeGetType(nd, &t) != cudaSuccess) continue;
      if (t == cudaGraphNodeTypeMemcpy) {
        cudaMemcpy3DParms p;
        if (cudaGraphMemcpyNodeGetParams(nd, &p) != cudaSuccess) continue;
        if (p.srcPtr.ptr == depth_in) g.n_in = nd;
        else if (p.dstPtr.ptr == depth_out) g.n_out = nd;
        else if (mask_out && p.dstPtr.ptr == mask_out) g.n_mask = nd;
      } else if (t == cudaGraphNodeTypeKernel) {
        cudaKernelNodeParams kp;
        if (cudaGraphKernelNodeGetParams(nd, &kp) == cudaSuccess && is_raster_kernel(kp.func)) {
          g.n_raster = nd;
          g.raster_kp = kp;
          g.ra.d = *static_cast<const Dims *>(kp.kernelParams[0]);
          g.ra.big = *static_cast<const TriRec *const *>(kp.kernelParams[1]);
          g.ra.bins = *static_cast<const BinRec *const *>(kp.kernelParams[2]);
          g.ra.ctr = *static_cast<const uint32_t *const *>(kp.kernelParams[3]);
          g.ra.tinfo = *static_cast<const uint4 *const *>(kp.kernelParams[4]);
          g.ra.sp = *static_cast<const ShaderParams *>(kp.kernelParams[5]);
          g.ra.fb = *static_cast<const FrameBuffers *>(kp.kernelParams[kRasterArgFrameBuffers]);
          g.ra.status = *static_cast<uint32_t *const *>(kp.kernelParams[7]);
        }
      }
    }
    if ((copy_in && !g.n_in) || (copy_out && (!g.n_out || (mask_out && !g.n_mask))) || ((direct & 3) && !g.n_raster)) {
      drop_frame_graph(c); c->use_graph = false; return 1;
    }
    g.in = depth_in; g.out = depth_out; g.mask = mask_out; g.direct = direct; g.bg_mode = bg_mode;
    g.multipass = (int)c->multipass; g.enc = enc; g.mask_format = c->mask_format; g.n_parts = P; g.max_diff = max_diff; g.replace_value = replace_value;
    g.ws_bins = c->ws.bins; g.stage_in = c->d_in[0]; g.cap_big = c->dims.cap_big; g.cap_tile = c->dims.cap_tile;
  } else {
    // same shape of work, other host buffers: retarget the copy nodes of the instantiated graph and, where the raster
    // kernel itself reads or writes the caller's buffers, its FrameBuffers argument
    if (g.in != depth_in && g.n_in)
      RUF_CUDA(c, cudaGraphExecMemcpyNodeSetParams1D(g.exec, g.n_in, c->d_in[0], depth_in, img * es, cudaMemcpyHostToDevice));
    if (g.out != depth_out && g.n_out)
      RUF_CUDA(c, cudaGraphExecMemcpyNodeSetParams1D(g.exec, g.n_out, depth_out, c->d_out[0], img * es, cudaMemcpyDeviceToHost));
    if (mask_out && g.mask != mask_out && g.n_mask)
      RUF_CUDA(c, cudaGraphExecMemcpyNodeSetParams1D(g.exec, g.n_mask, mask_out, c->d_mask[0], mb, cudaMemcpyDeviceToHost));
    if ((direct & 3) && (g.in != depth_in || g.out != depth_out || g.mask != mask_out)) {
      // (only the executable graph is updated, from the parameter list kept at capture time: one driver call)
      if (direct & 2) g.ra.fb.depth_in = dev_in;
      if (direct & 1) { g.ra.fb.depth_out = dev_out; g.ra.fb.mask_out = (uint8_t *)dev_mask; }
      void *args[kRasterArgCount] = {&g.ra.d, &g.ra.big, &g.ra.bins, &g.ra.ctr, &g.ra.tinfo, &g.ra.sp, &g.ra.fb, &g.ra.status};
      cudaKernelNodeParams kp = g.raster_kp;
      kp.kernelParams = args;
      kp.extra = nullptr;
      RUF_CUDA(c, cudaGraphExecKernelNodeSetParams(g.exec, g.n_raster, &kp));
    }
    g.in = depth_in; g.out = depth_out; g.mask = mask_out;
  }
  RUF_CUDA(c, cudaGraphLaunch(g.exec, sk));
  RUF_CUDA(c, cudaStreamSynchronize(sk));
  c->stats = ruf_stats{};
  c->stats.frames = 1;
  c->stats.kernel_launches = cluster ? 3 : 4;
  c->stats.h2d_bytes = (int64_t)(mat_bytes + img * es);
  c->stats.d2h_bytes = (int64_t)(img * es + (mask_out ? mb : 0));
  c->last_frames = 1;
  take_launch_stats(c);
  if (*c->h_status) {                             // an internal list overflowed: grow it and let the pipeline redo the frame
    const uint32_t flags = *c->h_status;
    RUF_CUDA(c, cudaMemsetAsync(c->ws.status, 0, sizeof(uint32_t), sk));
    *c->h_status = 0;
    if (flags & kFlagBigOverflow) c->want_big = 4LL * c->dims.cap_big;
    if (flags & kFlagBinOverflow) c->want_bin = 2LL * c->dims.cap_tile;
    return 1;
  }
  return RUF_OK;
}

int ruf_filter(ruf_context *c, const void *depth_in, int enc, const double *proj, const double *view,
               const double *part_model, float max_diff, float replace_value, void *depth_out, uint8_t *mask_out)
{
  if (c && depth_in && depth_out && proj && view && (c->n_parts == 0 || part_model) && c->have_model &&
      (enc == RUF_ENC_F32_M || enc == RUF_ENC_U16_MM) && cudaSetDevice(c->device) == cudaSuccess) {
    // Pageable buffers (a numpy array, a message's data vector) travel through the context's page-locked staging: one
    // memcpy each way and the single-frame graph in between, instead of the pageable copies of the staged pipeline.
    const size_t es = elem_size(enc), img = (size_t)c->W * c->H, mb = mask_bytes(c);
    const bool p_in = is_pinned_host(depth_in), p_out = is_pinned_host(depth_out), p_mask = !mask_out || is_pinned_host(mask_out);
    const void *src = depth_in;
    void *dst = depth_out;
    uint8_t *msk = mask_out;
    if (!(p_in && p_out && p_mask) && c->host_staging && c->use_graph && !c->profiling && !c->slice_frames) {
      if (!c->hs_in) {
        const size_t cap = img * sizeof(float);
        if (cudaHostAlloc(&c->hs_in, cap, cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc(&c->hs_out, cap, cudaHostAllocDefault) != cudaSuccess ||
            cudaHostAlloc(&c->hs_mask, img, cudaHostAllocDefault) != cudaSuccess) {
          cudaGetLastError();
          cudaFreeHost(c->hs_in); cudaFreeHost(c->hs_out); cudaFreeHost(c->hs_mask);
          c->hs_in = c->hs_out = nullptr; c->hs_mask = nullptr;
          c->host_staging = false;
        }
      }
      if (c->hs_in) {
        if (!p_in) { std::memcpy(c->hs_in, depth_in, img * es); src = c->hs_in; }
        if (!p_out) dst = c->hs_out;
        if (!p_mask) msk = (uint8_t *)c->hs_mask;
      }
    }
    const int rc = single_frame_graph(c, src, enc, proj, view, part_model, max_diff, replace_value, dst, msk);
    if (rc == RUF_OK) {
      if (dst != depth_out) std::memcpy(depth_out, dst, img * es);
      if (msk != mask_out) std::memcpy(mask_out, msk, mb);
    }
    if (rc != 1) return rc;
  }
  return ruf_filter_batch_host(c, 1, depth_in, enc, proj, view, part_model, max_diff, replace_value, depth_out,
                               mask_out);
}

int ruf_set_kinematics(ruf_context *c, int n_links, const int32_t *parent, const int32_t *joint_type,
                       const double *origin, const double *axis, const int32_t *part_link, const double *part_local,
                       int cam_link, const double *cam_mount, const double *view_pre)
{
  if (!c) return RUF_ERR_INVALID;
  if (!c->have_model) return fail(c, RUF_ERR_NO_MODEL, "load the model first (ruf_set_model)");
  if (n_links < 0 || n_links > (1 << 20) || cam_link >= n_links || !cam_mount || !view_pre ||
      (n_links > 0 && (!parent || !joint_type || !origin || !axis)) || (c->n_parts > 0 && (!part_link || !part_local)))
    return fail(c, RUF_ERR_INVALID, "bad kinematics arguments");
  std::vector<int32_t> off(n_links + 1, 0), idx;
  for (int l = 0; l < n_links; ++l) {
    if (parent[l] >= l || parent[l] < -1) return fail(c, RUF_ERR_INVALID, "links must be ordered parents first (link %d)", l);
    if (joint_type[l] < 0 || joint_type[l] > 2) return fail(c, RUF_ERR_INVALID, "joint_type[%d] invalid", l);
    std::vector<int32_t> chain;
    for (int a = l; a >= 0; a = parent[a]) chain.push_back(a);
    for (auto it = chain.rbegin(); it != chain.rend(); ++it) idx.push_back(*it);
    off[l + 1] = (int32_t)idx.size();
  }
  for (int p = 0; p < c->n_parts; ++p)
    if (part_link[p] < 0 || part_link[p] >= n_links) return fail(c, RUF_ERR_INVALID, "part_link[%d] out of range", p);
  RUF_CUDA(c, cudaSetDevice(c->device));
  RUF_CUDA(c, cudaStreamSynchronize(c->stream));
  // pack everything into one blob (8-byte aligned sections)
  auto pad8 = [](size_t n) { return (n + 7) & ~(size_t)7; };
  const size_t nl = (size_t)n_links, np = (size_t)c->n_parts;
  const size_t o_type = 0, o_origin = pad8(o_type + nl * 4), o_axis = o_origin + nl * 128, o_off = o_axis + nl * 24,
               o_idx = pad8(o_off + (nl + 1) * 4), o_plink = pad8(o_idx + idx.size() * 4), o_plocal = pad8(o_plink + np * 4),
               o_mount = o_plocal + np * 128, o_pre = o_mount + 128, total = o_pre + 128;
  std::vector<unsigned char> h(total, 0);
  if (nl) {
    std::memcpy(h.data() + o_type, joint_type, nl * 4);
    std::memcpy(h.data() + o_origin, origin, nl * 128);
    std::memcpy(h.data() + o_axis, axis, nl * 24);
  }
  std::memcpy(h.data() + o_off, off.data(), (nl + 1) * 4);
  if (!idx.empty()) std::memcpy(h.data() + o_idx, idx.data(), idx.size() * 4);
  if (np) {
    std::memcpy(h.data() + o_plink, part_link, np * 4);
    std::memcpy(h.data() + o_plocal, part_local, np * 128);
  }
  std::memcpy(h.data() + o_mount, cam_mount, 128);
  std::memcpy(h.data() + o_pre, view_pre, 128);
  cudaFree(c->kin_blob);
  c->kin_blob = nullptr;
  RUF_CUDA(c, cudaMalloc(&c->kin_blob, total));
  RUF_CUDA(c, cudaMemcpy(c->kin_blob, h.data(), total, cudaMemcpyHostToDevice));
  unsigned char *b = (unsigned char *)c->kin_blob;
  Kinematics &k = c->kin;
  k.n_links = n_links; k.n_parts = c->n_parts; k.cam_link = cam_link;
  k.type = (const int32_t *)(b + o_type); k.origin = (const double *)(b + o_origin); k.axis = (const double *)(b + o_axis);
  k.chain_off = (const int32_t *)(b + o_off); k.chain_idx = (const int32_t *)(b + o_idx);
  k.part_link = (const int32_t *)(b + o_plink); k.part_local = (const double *)(b + o_plocal);
  k.cam_mount = (const double *)(b + o_mount); k.view_pre = (const double *)(b + o_pre);
  return RUF_OK;
}

static int ensure_fk_buffers(ruf_context *c, int frames)
{
  if (!c->kin_blob) return fail(c, RUF_ERR_INVALID, "no kinematics loaded (ruf_set_kinematics)");
  if (c->kin.n_parts != c->n_parts) return fail(c, RUF_ERR_INVALID, "kinematics were set for another model");
  // the buffers hold frames x links and frames x parts matrices: a new model or new kinematics with more of
  // either needs larger ones even when the frame count did not grow
  if (c->fk_frames >= frames && c->fk_links_n >= c->kin.n_links && c->fk_parts_n >= c->n_parts) return RUF_OK;
  if (frames < c->fk_frames) frames = c->fk_frames;
  RUF_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFree(c->fk_links); cudaFree(c->fk_pm); cudaFree(c->fk_view);
  c->fk_links = c->fk_pm = c->fk_view = nullptr;
  c->fk_frames = 0; c->fk_links_n = 0; c->fk_parts_n = 0;
  const size_t f = (size_t)frames;
  RUF_CUDA(c, cudaMalloc(&c->fk_links, f * (c->kin.n_links > 0 ? c->kin.n_links : 1) * 128));
  RUF_CUDA(c, cudaMalloc(&c->fk_pm, f * (c->n_parts > 0 ? c->n_parts : 1) * 128));
  RUF_CUDA(c, cudaMalloc(&c->fk_view, f * 128));
  c->fk_frames = frames;
  c->fk_links_n = c->kin.n_links;
  c->fk_parts_n = c->n_parts;
  return RUF_OK;
}

int ruf_fk_batch_device(ruf_context *c, int n_frames, const double *d_joint_q, double camera_tx, double camera_ty,
                        double *d_part_model_out, double *d_view_out)
{
  if (!c) return RUF_ERR_INVALID;
  if (n_frames < 1 || !d_view_out || (c->n_parts > 0 && !d_part_model_out)) return fail(c, RUF_ERR_INVALID, "bad arguments");
  RUF_CUDA(c, cudaSetDevice(c->device));
  int rc = ensure_fk_buffers(c, n_frames);
  if (rc != RUF_OK) return rc;
  if (c->kin.n_links > 0 && !d_joint_q) return fail(c, RUF_ERR_INVALID, "d_joint_q is NULL");
  cudaError_t e = launch_fk(c->kin, n_frames, d_joint_q, camera_tx, camera_ty, c->fk_links, d_part_model_out, d_view_out,
                            c->stream);
  if (e != cudaSuccess) return fail(c, RUF_ERR_CUDA, "fk launch failed: %s", cudaGetErrorString(e));
  return RUF_OK;
}

int ruf_filter_batch_device_fk(ruf_context *c, int n_frames, const void *d_depth_in, int enc, const double *d_proj,
                               const double *d_joint_q, double camera_tx, double camera_ty, float max_diff,
                               float replace_value, void *d_depth_out, uint8_t *d_mask_out, float *d_zbuf_out)
{
  if (!c) return RUF_ERR_INVALID;
  RUF_CUDA(c, cudaSetDevice(c->device));
  int rc = ensure_fk_buffers(c, n_frames);
  if (rc != RUF_OK) return rc;
  rc = ruf_fk_batch_device(c, n_frames, d_joint_q, camera_tx, camera_ty, c->fk_pm, c->fk_view);
  if (rc != RUF_OK) return rc;
  rc = ruf_filter_batch_device(c, n_frames, d_depth_in, enc, d_proj, c->fk_view, c->fk_pm, max_diff, replace_value,
                               d_depth_out, d_mask_out, d_zbuf_out);
  if (rc == RUF_OK) c->stats.kernel_launches += 2;
  return rc;
}

int ruf_set_mask_format(ruf_context *c, int format)
{
  if (!c) return RUF_ERR_INVALID;
  if (format != RUF_MASK_BYTES && format != RUF_MASK_BITS) return fail(c, RUF_ERR_INVALID, "unknown mask format %d", format);
  if (format == RUF_MASK_BITS && c->W % 8 != 0)
    return fail(c, RUF_ERR_INVALID, "RUF_MASK_BITS needs an image width that is a multiple of 8 (width is %d)", c->W);
  c->mask_format = format;
  return RUF_OK;
}

int ruf_set_profiling(ruf_context *c, int enable)
{
  if (!c) return RUF_ERR_INVALID;
  c->profiling = enable != 0;
  return RUF_OK;
}

int ruf_get_stage_times(ruf_context *c, double *ms, int64_t *calls, int reset)
{
  if (!c || !ms) return RUF_ERR_INVALID;
  RUF_CUDA(c, cudaSetDevice(c->device));
  RUF_CUDA(c, cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i + kNumStages + 1 <= c->ev_used; i += kNumStages + 1) {
    for (int k = 0; k < kNumStages; ++k) {
      float t = 0.f;
      RUF_CUDA(c, cudaEventElapsedTime(&t, c->ev_pool[i + k], c->ev_pool[i + k + 1]));
      c->stage_ms[k] += t;
    }
    ++c->stage_calls;
  }
  c->ev_used = 0;
  for (int k = 0; k < kNumStages; ++k) ms[k] = c->stage_ms[k];
  if (calls) *calls = c->stage_calls;
  if (reset) {
    for (int k = 0; k < kNumStages; ++k) c->stage_ms[k] = 0;
    c->stage_calls = 0;
  }
  return RUF_OK;
}

int ruf_host_alloc(void **ptr, size_t bytes)
{
  if (!ptr) return RUF_ERR_INVALID;
  cudaError_t e = cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) return fail(nullptr, RUF_ERR_CUDA, "cudaHostAlloc: %s", cudaGetErrorString(e));
  return RUF_OK;
}
int ruf_host_free(void *ptr)
{
  if (ptr) cudaFreeHost(ptr);
  return RUF_OK;
}
int ruf_host_is_pinned(const void *p) { return p && is_pinned_host(p) ? 1 : 0; }

int ruf_get_stats(ruf_context *c, ruf_stats *out)
{
  if (!c || !out) return RUF_ERR_INVALID;
  RUF_CUDA(c, cudaSetDevice(c->device));
  RUF_CUDA(c, cudaStreamSynchronize(c->stream));
  ruf_stats s = c->stats;
  s.visible_tris = s.binned_refs = s.big_tris = 0;
  if (c->last_frames > 0 && c->ws.ctr) {
    const size_t stride = (size_t)c->dims.ctr_stride;
    std::vector<uint32_t> h((size_t)c->last_frames * stride);
    RUF_CUDA(c, cudaMemcpy(h.data(), c->ws.ctr, h.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    for (int f = 0; f < c->last_frames; ++f) {
      const uint32_t *p = h.data() + (size_t)f * stride;
      s.visible_tris += p[kCtrKept];
      s.big_tris += p[kCtrBig];
#ifdef RUF_CULL_STATS
      s.h2d_bytes += p[1] & 0xffffu; s.d2h_bytes += p[1] >> 16;      // debug: depth-cull tested / culled records
#endif
      for (int t = 0; t < 2 * c->dims.ntiles; ++t) s.binned_refs += p[kCtrWords + t];
    }
  }
  *out = s;
  return RUF_OK;
}

}  // extern "C"

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU group: ONE host process, one context + one host thread per device, frames sharded with
 * no per-frame collective.  The only communication is the set-up broadcast of the static model buffers from
 * device 0 to the others over NVLink / NVSwitch (ncclCommInitAll + one ncclBroadcast per buffer).  NCCL is loaded at
 * run time (libnccl.so.2): a single-GPU host never needs it.
 * ------------------------------------------------------------------------------------------------ */
#include <dlfcn.h>
#include <thread>

namespace {
typedef struct ncclComm *ncclComm_t;
struct NcclApi {
  void *lib = nullptr;
  int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int /*ncclDataType_t*/, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  bool load(std::string &err)
  {
    if (lib) return true;
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { err = std::string("cannot load NCCL: ") + dlerror(); return false; }
    CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
    Broadcast = (decltype(Broadcast))dlsym(lib, "ncclBroadcast");
    GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
    GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!CommInitAll || !CommDestroy || !Broadcast || !GroupStart || !GroupEnd || !GetErrorString) {
      err = "libnccl lacks an expected symbol";
      return false;
    }
    return true;
  }
};
NcclApi g_nccl;
constexpr int kNcclChar = 0;     // ncclInt8 / ncclChar
}  // namespace

struct ruf_group {
  std::vector<ruf_context *> ctx;
  std::vector<ncclComm_t> comm;
  std::string err;
  int64_t bcast_bytes = 0;       // bytes every non-root device received in the last ruf_group_set_model
};

static int gfail(ruf_group *g, int code, const std::string &msg)
{
  if (g) g->err = msg; else g_create_error = msg;
  return code;
}

extern "C" {

int ruf_group_create(ruf_group **out, int n_devices, const int *devices, int width, int height, double z_near, double z_far)
{
  if (!out || n_devices < 1 || n_devices > 64) return gfail(nullptr, RUF_ERR_INVALID, "bad group arguments");
  *out = nullptr;
  ruf_group *g = new (std::nothrow) ruf_group;
  if (!g) return gfail(nullptr, RUF_ERR_NOMEM, "out of host memory");
  std::vector<int> devs(n_devices);
  for (int i = 0; i < n_devices; ++i) devs[i] = devices ? devices[i] : i;
  for (int i = 0; i < n_devices; ++i) {
    ruf_context *c = nullptr;
    const int rc = ruf_create(&c, devs[i], width, height, z_near, z_far);
    if (rc != RUF_OK) { ruf_group_destroy(g); return rc; }          // g_create_error holds the text
    g->ctx.push_back(c);
  }
  if (n_devices > 1) {
    std::string err;
    if (!g_nccl.load(err)) { ruf_group_destroy(g); return gfail(nullptr, RUF_ERR_CUDA, err); }
    g->comm.assign(n_devices, nullptr);
    const int rc = g_nccl.CommInitAll(g->comm.data(), n_devices, devs.data());
    if (rc != 0) {
      g->comm.clear();
      ruf_group_destroy(g);
      return gfail(nullptr, RUF_ERR_CUDA, std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(rc));
    }
  }
  *out = g;
  return RUF_OK;
}

int ruf_group_destroy(ruf_group *g)
{
  if (!g) return RUF_OK;
  for (size_t i = 0; i < g->comm.size(); ++i)
    if (g->comm[i]) { cudaSetDevice(g->ctx[i]->device); g_nccl.CommDestroy(g->comm[i]); }
  for (ruf_context *c : g->ctx) ruf_destroy(c);
  delete g;
  return RUF_OK;
}

int ruf_group_size(const ruf_group *g) { return g ? (int)g->ctx.size() : 0; }
ruf_context *ruf_group_context(ruf_group *g, int i) { return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr; }
const char *ruf_group_last_error(const ruf_group *g) { return g ? g->err.c_str() : g_create_error.c_str(); }
int64_t ruf_group_broadcast_bytes(const ruf_group *g) { return g ? g->bcast_bytes : 0; }

int ruf_group_set_model(ruf_group *g, const float *tri_xyz, const uint32_t *tri_part, int64_t n_tris, int n_parts)
{
  if (!g) return RUF_ERR_INVALID;
  if (n_tris < 0 || n_parts < 0 || n_parts > (1 << 20) || (n_tris > 0 && (!tri_xyz || !tri_part)) || n_tris > (1LL << 30))
    return gfail(g, RUF_ERR_INVALID, "bad model arguments");
  for (int64_t t = 0; t < n_tris; ++t)
    if (tri_part[t] >= (uint32_t)n_parts) return gfail(g, RUF_ERR_INVALID, "tri_part out of range");
  ruf_context *c0 = g->ctx[0];
  MeshletModel mm;                 // built once on the host, uploaded once (device 0), broadcast to the rest
  build_meshlet_sets(tri_xyz, tri_part, n_tris, n_parts, (float)(c0->z_far * 0.99), kMeshVerts, kMeshTris, kMeshTrisFine, kMeshParts, mm);
  for (ruf_context *c : g->ctx) {
    if (cudaSetDevice(c->device) != cudaSuccess) return gfail(g, RUF_ERR_CUDA, "cudaSetDevice failed");
    const int rc = alloc_model(c, mm, n_tris, n_parts);
    if (rc != RUF_OK) return gfail(g, rc, c->err);
  }
  cudaSetDevice(c0->device);
  int rc = upload_model(c0, mm);
  if (rc != RUF_OK) return gfail(g, rc, c0->err);
  g->bcast_bytes = 0;
  if (g->ctx.size() > 1) {
    const size_t bytes[4] = {mm.hdr.size() * sizeof(uint32_t), mm.verts.size() * sizeof(float), mm.tris.size() * sizeof(uint32_t),
                             mm.part_aabb.size() * sizeof(float)};
    for (int b = 0; b < 4; ++b) {                    // one broadcast per static buffer, root = device 0
      if (!bytes[b]) continue;
      int nrc = g_nccl.GroupStart();
      for (size_t i = 0; i < g->ctx.size() && nrc == 0; ++i) {
        ruf_context *c = g->ctx[i];
        void *buf = b == 0 ? (void *)c->meshlets : b == 1 ? (void *)c->mverts : b == 2 ? (void *)c->mtris : (void *)c->part_aabb;
        nrc = g_nccl.Broadcast(buf, buf, bytes[b], kNcclChar, 0, g->comm[i], c->stream);
      }
      const int erc = g_nccl.GroupEnd();
      if (nrc == 0) nrc = erc;
      if (nrc != 0) return gfail(g, RUF_ERR_CUDA, std::string("ncclBroadcast: ") + g_nccl.GetErrorString(nrc));
      g->bcast_bytes += (int64_t)bytes[b];
    }
    for (ruf_context *c : g->ctx) {
      cudaSetDevice(c->device);
      if (cudaStreamSynchronize(c->stream) != cudaSuccess) return gfail(g, RUF_ERR_CUDA, "broadcast did not complete");
    }
  }
  return RUF_OK;
}

/* n frames with host buffers: the frames are cut into chunks, chunk j goes to device j mod N (chunk = 1 is "frame k ->
 * GPU k mod N"), every device runs its chunks through its own staging pipeline on its own host thread, and every
 * result lands at its frame's position in the caller's arrays: the output is in sequence order by construction. */
int ruf_group_filter_batch_host(ruf_group *g, int n_frames, const void *depth_in, int enc, const double *proj,
                                const double *view, const double *part_model, float max_diff, float replace_value,
                                void *depth_out, uint8_t *mask_out, int frames_per_chunk)
{
  if (!g) return RUF_ERR_INVALID;
  const int N = (int)g->ctx.size();
  if (n_frames < 1 || !depth_in || !depth_out || !proj || !view || (enc != RUF_ENC_F32_M && enc != RUF_ENC_U16_MM))
    return gfail(g, RUF_ERR_INVALID, "bad arguments");
  int chunk = frames_per_chunk > 0 ? frames_per_chunk : host_chunk((n_frames + N - 1) / N);
  std::vector<std::vector<std::pair<int, int>>> segs(N);
  int j = 0;
  for (int f0 = 0; f0 < n_frames; f0 += chunk, ++j)
    segs[j % N].emplace_back(f0, (n_frames - f0 < chunk) ? (n_frames - f0) : chunk);
  std::vector<int> rcs(N, RUF_OK);
  auto work = [&](int i) {
    ruf_context *c = g->ctx[i];
    if (segs[i].empty()) return;
    if (!c->have_model) { rcs[i] = fail(c, RUF_ERR_NO_MODEL, "no model loaded (ruf_group_set_model)"); return; }
    if (cudaSetDevice(c->device) != cudaSuccess) { rcs[i] = fail(c, RUF_ERR_CUDA, "cudaSetDevice failed"); return; }
    for (int attempt = 0; attempt < 8; ++attempt) {
      int rc = ensure_workspace(c, chunk);
      if (rc == RUF_OK) rc = ensure_staging(c, chunk);
      if (rc != RUF_OK) { rcs[i] = rc; return; }
      c->stats = ruf_stats{};
      for (auto &sg : segs[i]) c->stats.frames += sg.second;
      rc = host_pipeline(c, n_frames, depth_in, enc, proj, view, part_model, max_diff, replace_value, depth_out,
                         mask_out, segs[i]);
      rcs[i] = rc;
      if (rc != RUF_ERR_OVERFLOW) return;       // overflow: this device's capacities were doubled, run its share again
    }
  };
  std::vector<std::thread> threads;
  for (int i = 1; i < N; ++i) threads.emplace_back(work, i);
  work(0);
  for (std::thread &t : threads) t.join();
  for (int i = 0; i < N; ++i)
    if (rcs[i] != RUF_OK) return gfail(g, rcs[i], "device " + std::to_string(g->ctx[i]->device) + ": " + g->ctx[i]->err);
  return RUF_OK;
}

}  // extern "C"
