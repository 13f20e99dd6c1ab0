// ruf_device.cuh -- shared device-side types and launch prototypes of the B200 hot path.
//
// Pipeline per batch of frames (DESIGN.md "Kernels"):
//   K0 ruf_pose_kernel      fp64  P * V_f * M_{f,p}  -> fp32 MVP table     (vertex-stage matrices)
//   K1 ruf_setup_bin_kernel one CTA / (frame, 512 triangles): vertex shader, clip, viewport, snap,
//                           setup, cull; CTA-local tile binning in shared memory; tile-sorted
//                           records + a (start,count) table entry per (tile, CTA)
//   K2 ruf_raster_filter_kernel   one CTA / (frame, 64x32 tile): gathers its segments with bulk
//                           async copies (TMA) into smem -> smem z-tile (min) -> fused fragment
//                           shader + encode + store
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ruf {

// ---- raster specification constants (DESIGN.md "Raster specification") ----
constexpr int kSubpixBits = 8;
constexpr int kSubpix = 1 << kSubpixBits;
constexpr int kSubpixHalf = kSubpix / 2;
constexpr float kGuardPx = 6144.0f;
constexpr float kWindowLimit = 16384.0f;
constexpr int kMaxPoly = 12;

// ---- tiling ----
constexpr int kTileW = 64;
constexpr int kTileH = 32;
constexpr int kTilePix = kTileW * kTileH;       // 2048
constexpr int kRasterThreads = 256;             // consumer threads: 8 pixels of one tile row each
constexpr int kRasterBlock = kRasterThreads + 32;   // + one producer warp (bulk async copies)
constexpr int kChunk = 256;                     // records per ring stage (one per consumer thread)
constexpr int kStages = 2;                      // ring depth (a stage is released as soon as its records sit in registers)
constexpr int kBigTiles = 12;                   // bbox touching more tiles -> per-frame "big" list
#ifndef RUF_MAX_UNITS
#define RUF_MAX_UNITS 32
#endif
constexpr int kMaxUnits = RUF_MAX_UNITS;                   // row-block units (1 row x 8 samples) per record dealt to lanes;
                                                // larger records are rasterised by the whole warp
// tuning knobs (overridable with -D for experiments; the defaults are the measured best)
#ifndef RUF_SETUP_THREADS
#define RUF_SETUP_THREADS 256
#endif
#ifndef RUF_TRIS_PER_THREAD
#define RUF_TRIS_PER_THREAD 4
#endif
#ifndef RUF_SETUP_MIN_BLOCKS
#define RUF_SETUP_MIN_BLOCKS 4
#endif
#ifndef RUF_RASTER_MIN_BLOCKS
#define RUF_RASTER_MIN_BLOCKS 4
#endif
constexpr int kSetupThreads = RUF_SETUP_THREADS;
constexpr int kTrisPerThread = RUF_TRIS_PER_THREAD;
constexpr int kSetupTris = kSetupThreads * kTrisPerThread;   // triangles per setup CTA
#ifndef RUF_SEG_CAP
#define RUF_SEG_CAP 512
#endif
constexpr int kSegCap = RUF_SEG_CAP;                    // table entries gathered per round by a raster CTA
constexpr int kMaxTiles = 4096;

constexpr int kNumStages = 3;                   // pose, setup+bin, raster+filter
constexpr uint32_t kFlagBigOverflow = 1u;
constexpr uint32_t kFlagBinOverflow = 2u;

// One window-space triangle after setup: 48 bytes = 3 x 16 B (bulk-copy granularity).
// Vertices are snapped (1/256 px) and ordered so that the doubled area is positive.
struct __align__(16) TriRec {
  int32_t x0, y0, x1, y1;
  int32_t x2, y2;
  float z0;      // window z at vertex 0
  float gx;      // dz/dX per sub-pixel unit
  float gy;      // dz/dY per sub-pixel unit
  uint32_t bx;   // pixel bbox clamped to the viewport: i0 | i1 << 16
  uint32_t by;   // j0 | j1 << 16
  uint32_t pad;
};
static_assert(sizeof(TriRec) == 48, "TriRec must be 48 bytes");
// dynamic shared memory of the raster kernel: record ring + per-warp unit tables
constexpr size_t kRasterDynSmem = sizeof(TriRec) * kStages * kChunk + sizeof(uint16_t) * (kRasterThreads / 32) * 32 * kMaxUnits;

// Per-frame counter block (uint32 words): [0] big-list entries  [1] tile references  [2] flags
// [3] kept (binned) triangles
constexpr int kCtrBig = 0, kCtrRef = 1, kCtrFlags = 2, kCtrKept = 3, kCtrWords = 4;

struct Dims {
  int W, H;
  int tiles_x, tiles_y, ntiles;
  int n_parts;           // model matrices per frame (the MVP table has n_parts + 1 rows)
  long long n_tris;
  uint32_t cap_big, cap_bin;   // per-frame capacities: big list, tile references
  int n_setup_ctas;            // setup CTAs per frame = table entries per tile
  float halfw, halfh, guard_x, guard_y;
};

struct ShaderParams {
  float k1, k2;          // to_linear_depth constants (float arithmetic, like the GLSL)
  float max_diff;
  float replace_value;
};

struct FrameBuffers {
  const void *depth_in;  // n_frames images
  void *depth_out;
  uint8_t *mask_out;     // may be null
  float *zbuf_out;       // may be null
  int vec_ok;            // rows are 16-byte aligned for 8-pixel vectors
};

struct Workspace {
  float *mvp;            // [frame][n_parts + 1][16]
  uint32_t *ctr;         // [frame][kCtrWords]
  TriRec *big;           // [frame][cap_big]
  TriRec *bins;          // [frame][cap_bin] tile-sorted segments, one per (setup CTA, tile)
  uint2 *table;          // [frame][tile][n_setup_ctas] (start, count) into bins
  uint32_t *status;      // sticky OR of all frame flags
};

struct Model {
  const float4 *v0;      // xyz + part index bits in w
  const float4 *v1;
  const float4 *v2;
  const float *part_aabb;   // [n_parts][6] object-space min xyz, max xyz
  const uint2 *cta_parts;   // [n_setup_ctas] (lowest, highest) part index among the CTA's triangles
};
constexpr int kCullParts = 8;   // a setup CTA culls per part when its triangles span at most this many parts

// cudaSuccess iff the loaded module has an image the current device can run (sm_100a only)
cudaError_t check_kernel_image();
cudaError_t launch_pack_model(const float *d_tri_xyz, const uint32_t *d_tri_part, long long n_tris,
                              float4 *v0, float4 *v1, float4 *v2, cudaStream_t s);
// forward kinematics: joint positions -> link poses -> part models + view matrix
struct Kinematics {
  int n_links, n_parts, cam_link;
  const int32_t *type;        // [n_links] 0 fixed, 1 revolute, 2 prismatic
  const double *origin;       // [n_links][16] parent_T_joint
  const double *axis;         // [n_links][3] unit
  const int32_t *chain_off;   // [n_links + 1] CSR offsets into chain_idx
  const int32_t *chain_idx;   // ancestors of every link, root first, the link itself last
  const int32_t *part_link;   // [n_parts]
  const double *part_local;   // [n_parts][16] link_offset [* suffix]
  const double *cam_mount;    // [16] camera optical frame in its link
  const double *view_pre;     // [16] LookAt * inverse(camera_offset)
};
cudaError_t launch_fk(const Kinematics &k, int n_frames, const double *d_joint_q, double tx, double ty,
                      double *d_links, double *d_part_model, double *d_view, cudaStream_t s);
cudaError_t launch_frames(const Dims &d, const Model &m, const Workspace &ws, int n_frames,
                          const double *d_proj, const double *d_view, const double *d_part_model,
                          const double *d_lookat, float bg_z, int enc, const ShaderParams &sp,
                          const FrameBuffers &fb, cudaStream_t s, int *n_launches,
                          cudaEvent_t *stage_events /* null or kNumStages+1 events */);

}  // namespace ruf
