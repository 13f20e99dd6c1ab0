// ruf_device.cuh -- shared device-side types and launch prototypes of the B200 hot path.
//
// Pipeline per batch of frames (DESIGN.md "Kernels"):
//   K0 ruf_pose_kernel      fp64  P * V_f * M_{f,p}  -> fp32 MVP table     (vertex-stage matrices)
//   K1 ruf_setup_kernel     one thread / (frame, triangle): vertex shader, clip, viewport,
//                           snap, setup, cull, count tile references
//   K2 ruf_scan_kernel      per-frame exclusive scan of the tile counters
//   K3 ruf_bin_kernel       copy each kept record into the bins of the tiles it touches
//   K4 ruf_raster_filter_kernel   one CTA / (frame, 64x32 tile): TMA-staged bins -> smem
//                           z-tile (min) -> fused fragment shader + encode + store
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
constexpr int kRasterThreads = 256;             // 8 pixels of one tile row per thread
constexpr int kChunk = 256;                     // records per TMA stage (one per thread)
constexpr int kBigTiles = 12;                   // bbox touching more tiles -> per-frame "big" list
constexpr int kBigCapacity = 1024;              // entries of the per-frame big list
constexpr int kSmallArea = 48;                  // clipped bbox area handled by one lane

constexpr int kNumStages = 5;                   // pose, setup, scan, bin, raster+filter
constexpr uint32_t kFlagRecOverflow = 1u;
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

// Per-frame counter block (uint32 words) inside the workspace.
//   [0] kept records  [1] big-list entries  [2] total bin references  [3] flags
//   [4 .. 4+nt)        tile reference counts
//   [4+nt .. 4+2nt)    tile fill cursors
//   [4+2nt .. 4+3nt)   tile offsets (exclusive scan)
constexpr int kCtrRec = 0, kCtrBig = 1, kCtrBinTotal = 2, kCtrFlags = 3, kCtrTiles = 4;

struct Dims {
  int W, H;
  int tiles_x, tiles_y, ntiles;
  int n_parts;           // model matrices per frame (the MVP table has n_parts + 1 rows)
  long long n_tris;
  uint32_t cap_rec, cap_bin;
  uint32_t ctr_stride;   // uint32 words per frame counter block
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
  uint32_t *ctr;         // [frame][ctr_stride]
  TriRec *recs;          // [frame][cap_rec]
  TriRec *big;           // [frame][kBigCapacity]
  TriRec *bins;          // [frame][cap_bin]
  uint32_t *status;      // sticky OR of all frame flags
};

struct Model {
  const float4 *v0;      // xyz + part index bits in w
  const float4 *v1;
  const float4 *v2;
};

// cudaSuccess iff the loaded module has an image the current device can run (sm_100a only)
cudaError_t check_kernel_image();
cudaError_t launch_pack_model(const float *d_tri_xyz, const uint32_t *d_tri_part, long long n_tris,
                              float4 *v0, float4 *v1, float4 *v2, cudaStream_t s);
cudaError_t launch_frames(const Dims &d, const Model &m, const Workspace &ws, int n_frames,
                          const double *d_proj, const double *d_view, const double *d_part_model,
                          const double *d_lookat, float bg_z, int enc, const ShaderParams &sp,
                          const FrameBuffers &fb, cudaStream_t s, int *n_launches,
                          cudaEvent_t *stage_events /* null or 6 events: before K0, after K0..K4 */);

}  // namespace ruf
