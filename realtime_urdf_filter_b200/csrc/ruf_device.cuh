// ruf_device.cuh -- shared device-side types and launch prototypes of the B200 hot path.
//
// Pipeline per batch of frames (DESIGN.md "Kernels"):
//   K0 ruf_pose_kernel      fp64  P * V_f * M_{f,p}  -> fp32 MVP table; per (frame, part): view-volume
//                           cull bit and facing hint
//   K1 ruf_setup_bin_kernel one CTA / (meshlet, run of frames): vertex shader once per welded vertex,
//                           cheap per-triangle rejects, survivors compacted per warp, clip, set-up;
//                           match.any walk over the bbox tiles, ONE 8-byte global atomic per (warp step,
//                           tile) reserves room in that tile's record list (front run / back run)
//   K2 ruf_raster_filter_kernel   one CTA / (frame, 64x64 tile): streams its tile's record list with bulk
//                           async copies (TMA) into a shared-memory ring -> units of 4x2 samples dealt to
//                           lanes -> z tile in shared memory (atomicMin), depth-culled second pass for the
//                           back run -> per-frame big list merged on registers -> fused fragment shader +
//                           encode + store; tiles without records never touch the z tile
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
#ifndef RUF_TILE_H
#define RUF_TILE_H 64
#endif
constexpr int kTileH = RUF_TILE_H;               // 32 or 64 (a raster thread then shades two rows)
constexpr int kTilePix = kTileW * kTileH;
constexpr int kRowsPerThread = kTileH / 32;
static_assert(kTileH == 32 || kTileH == 64, "tile height");
constexpr int kRasterThreads = 256;             // threads of a raster CTA: 8 pixels of one tile row each
#ifndef RUF_CHUNK
#define RUF_CHUNK 128
#endif
constexpr int kChunk = RUF_CHUNK;               // records per ring stage (a multiple of 32)
constexpr int kStages = 2;                      // ring depth (a stage is released as soon as its records sit in registers)
#ifndef RUF_BIG_TILES
#define RUF_BIG_TILES 12
#endif
constexpr int kBigTiles = RUF_BIG_TILES;                   // bbox touching more tiles -> per-frame "big" list
#ifndef RUF_MAX_UNITS
#define RUF_MAX_UNITS 24
#endif
constexpr int kMaxUnits = RUF_MAX_UNITS;        // units per record dealt to lanes; larger records are rasterised by the whole warp
#ifndef RUF_UNIT_W
#define RUF_UNIT_W 4
#endif
#ifndef RUF_UNIT_H
#define RUF_UNIT_H 2
#endif
constexpr int kUW = RUF_UNIT_W, kUH = RUF_UNIT_H;   // a unit = kUW x kUH samples handled by one lane in one round
static_assert((kUW == 8 && kUH == 1) || (kUW == 4 && kUH == 2) || (kUW == 4 && kUH == 1) || (kUW == 8 && kUH == 2),
              "unsupported unit shape");
// tuning knobs (overridable with -D for experiments; the defaults are the measured best)
#ifndef RUF_SETUP_THREADS
#define RUF_SETUP_THREADS 256
#endif
#ifndef RUF_MESH_VERTS
#define RUF_MESH_VERTS 512
#endif
#ifndef RUF_MESH_TRIS
#define RUF_MESH_TRIS 1023
#endif
#ifndef RUF_SETUP_FRAMES
#define RUF_SETUP_FRAMES 4
#endif
#ifndef RUF_SETUP_MIN_BLOCKS
#define RUF_SETUP_MIN_BLOCKS 6
#endif
#ifndef RUF_RASTER_MIN_BLOCKS
#define RUF_RASTER_MIN_BLOCKS 5
#endif
#ifndef RUF_WALK
#define RUF_WALK 2
#endif
#ifndef RUF_MIN_BACK_BATCHES
#define RUF_MIN_BACK_BATCHES 6u
#endif
#ifndef RUF_EARLY_RESERVE
#define RUF_EARLY_RESERVE 1       // setup kernel: tile reservation atomics issued before the depth-plane divisions (-7 % on the kernel)
#endif
#ifndef RUF_OCCLUDE_MIN
#define RUF_OCCLUDE_MIN 8
#endif
constexpr int kOccludeMin = RUF_OCCLUDE_MIN;     // big-list rounds with more records than this run the per-tile occlusion test
#ifndef RUF_DEPTH_CULL
#define RUF_DEPTH_CULL 1
#endif
constexpr bool kDepthCull = RUF_DEPTH_CULL != 0;   // raster kernel: drop units of away-facing triangles that cannot win the depth test
constexpr int kSetupThreads = RUF_SETUP_THREADS;
// A meshlet is the unit of work of one setup CTA: up to kMeshTris consecutive triangles of the model whose
// bit-identical vertices were welded (at most kMeshVerts distinct ones, local indices of 10 bits) and whose
// parts span at most kMeshParts consecutive part indices.
constexpr int kMeshVerts = RUF_MESH_VERTS;
constexpr int kMeshTris = RUF_MESH_TRIS;
constexpr int kMeshParts = 32;
constexpr int kMeshTrisFine = RUF_SETUP_THREADS;  // the fine cut of the model: one triangle per setup thread
constexpr int kFineMaxCtas = 148;                 // (frames x meshlets) of a launch up to which the fine cut is used (one frame of a 90k model)
constexpr int kSetupSlots = (kMeshTris + kSetupThreads - 1) / kSetupThreads;   // triangles per thread
constexpr int kSetupFrames = RUF_SETUP_FRAMES;  // frames a setup CTA loops over with its meshlet in registers
static_assert(kMeshVerts <= 1024 && kMeshTris <= 1023, "meshlet indices are packed in 10 bits");
constexpr int kMultiPassUnits = 128;            // MP variant: records up to this many units are dealt out in passes of kMaxUnits
#ifndef RUF_CLUSTER_SPLIT
#define RUF_CLUSTER_SPLIT 4
#endif
#ifndef RUF_CLUSTER_BATCH
#define RUF_CLUSTER_BATCH 32
#endif
constexpr int kClusterBatch = RUF_CLUSTER_BATCH;   // records per warp batch of that variant (divides 32)
constexpr int kClusterSplit = RUF_CLUSTER_SPLIT;   // CTAs per tile of the low-latency raster variant (divides 32)
constexpr int kClusterMaxTiles = 384;            // (frames x tiles) of a launch up to which the cluster-split variant is used: measured
                                                 // at 640x480 (profiles/small_batch_probe.py) it wins up to 4 frames per launch, loses from 8
constexpr int kWideCap = 16;                    // wide records a raster CTA parks for its cooperative final phase
constexpr int kMaxTiles = 4096;
constexpr int kPartStride = 8;                  // floats per part in Model::part_aabb
constexpr int kZPad = 72;                       // zero words after the z tile (the depth-cull reads may run past its end)

constexpr int kNumStages = 3;                   // pose, setup+bin, raster+filter
constexpr uint32_t kFlagBigOverflow = 1u;
constexpr uint32_t kFlagBinOverflow = 2u;
// per (frame, tile) word written by ruf_tile_info_kernel: x = flags, y = integer threshold (16UC1) or bits of the float
// threshold (32FC1), z = bits of the tile's single window z
constexpr uint32_t kTileFlat = 1u, kTileUndrawn = 2u;
constexpr int kBgSkip = 1, kBgOnly = 2;
constexpr int kBgSeedMax = 8;                   // records the background quad may clip into (two triangles, five planes)
constexpr int kFlatMaxBig = 8;                  // big-list records a tile-info thread is willing to classify

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
// The same triangle as it travels through a tile's record list: 32 bytes = one DRAM sector, two 16-byte quantities of the
// bulk copy.  Vertex 0 absolute, the other two as 24-bit differences (12 bytes: dx1, dy1, dx2, dy2, three bytes each,
// little endian; the guard band bounds every coordinate by 6144 px = 2^20.6 sub-pixel units, so every difference fits)
// -- one PRMT with a sign-replicating selector restores each -- and the depth plane; the pixel bbox is recomputed by the
// raster kernel from the vertices it needs anyway.
struct __align__(16) BinRec {
  int32_t x0, y0;
  uint32_t d[3];
  float z0, gx, gy;
};
static_assert(sizeof(BinRec) == 32, "BinRec must be 32 bytes");
// dynamic shared memory of the raster kernel: record ring + per-warp unit tables
constexpr size_t kRasterDynSmem = sizeof(BinRec) * kStages * kChunk + sizeof(uint16_t) * (kRasterThreads / 32) * 32 * kMaxUnits;

// Per-frame counter block (uint32 words): [0] big-list entries  [1] unused  [2] flags
// [3] kept (binned) triangles  [4 + 2 t], [5 + 2 t]: records at the FRONT of tile t's list (triangles facing
// the camera) and at its BACK (facing away: drawn last, depth-culled); one 8-byte word per tile
constexpr int kCtrBig = 0, kCtrFlags = 2, kCtrKept = 3, kCtrWords = 4;

struct Dims {
  int W, H;
  int tiles_x, tiles_y, ntiles;
  int n_parts;           // model matrices per frame (the MVP table has n_parts + 1 rows)
  long long n_tris;
  uint32_t cap_big, cap_tile;  // capacities: big list per frame, record list per (frame, tile)
  int n_meshlets;              // setup CTAs per frame
  int ctr_stride;              // uint32 words of one frame's counter block = kCtrWords + 2 * ntiles
  int force_fpc;               // > 0: frames per setup CTA (testing aid, RUF_SETUP_FRAMES_FORCE); 0 = heuristic
  int multipass;               // raster kernel variant of this launch (ruf_raster_filter_kernel<ENC, MP, CL>)
  int fold_clear;              // the pose kernel clears the counter blocks (single-frame graph) instead of a memset before it
  int bg_mode;                 // background quad (the model's last meshlet): 0 = set up by the setup kernel like every triangle;
                               // kBgSkip = its records come from Workspace::bg_seed (fold_clear launches only), the setup kernel
                               // leaves the last meshlet out; kBgOnly = pose + setup of the last meshlet alone (fills the seed)
  int cluster_split;           // low-latency variant: kClusterSplit CTAs (one thread-block cluster) share a tile
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
  int mask_bits;         // mask_out holds 1 bit per pixel (W % 8 == 0 required) instead of one 0 / 255 byte
  uint32_t *host_status; // cluster-split variant only, may be null: the host's pinned (mapped) copy of the status words, written
                         // by the last CTA of the launch
};

struct Workspace {
  float *mvp;            // [frame][n_parts + 1][16]
  uint8_t *vis;          // [frame][n_parts + 1] bit 0: the part may be visible in this frame; bit 1: its triangles with
                         // POSITIVE window area face the camera (else the negative ones do)
  uint32_t *ctr;         // [frame][ctr_stride]
  TriRec *big;           // [frame][cap_big]
  BinRec *bins;          // [frame][tile][cap_tile] one record list per tile
  uint32_t *status;      // [0] sticky OR of all frame flags  [1] kept records (tile-info kernel), [2] (record, tile) pairs above kMaxUnits units (raster kernel) since the host last cleared them
  uint4 *tinfo;          // [frame][tile] ruf_tile_info_kernel -> ruf_raster_filter_kernel
  const uint32_t *bg_seed;   // kBgSkip: word 0 = record count n, words 4.. = n TriRecs (the background quad's big-list records)
};

struct Model {
  const uint4 *meshlets;    // [n_meshlets] vert_off, tri_off, nverts | ntris << 10 | (nparts - 1) << 20, lowest part
  const float4 *verts;      // welded vertices of all meshlets: xyz, w = bits(part - meshlet's lowest part)
  const uint32_t *tris;     // local vertex indices i0 | i1 << 10 | i2 << 20
  const float *part_aabb;   // [n_parts][kPartStride] object-space min xyz, max xyz, winding (+1 / -1), pad
};

// cudaSuccess iff the loaded module has an image the current device can run (sm_100a only)
cudaError_t check_kernel_image();
bool is_raster_kernel(const void *func);   // for the host's graph bookkeeping
constexpr int kRasterArgCount = 8, kRasterArgFrameBuffers = 6;   // ruf_raster_filter_kernel's parameter list
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
                          const double *d_lookat, int enc, const ShaderParams &sp,
                          const FrameBuffers &fb, cudaStream_t s, int *n_launches,
                          cudaEvent_t *stage_events /* null or kNumStages+1 events */,
                          cudaEvent_t depth_ready = nullptr /* the raster kernel waits for it (upload of the depth image) */);

}  // namespace ruf
