// ruf_kernels.cu -- sm_100a kernels of the URDF depth self-filter hot path.
//
// Replaces, for one batch of frames, what the reference does through OpenGL:
//   include/shaders/urdf_filter.vert:5        -> xform() in ruf_setup_kernel
//   fixed-function clip / viewport / raster   -> ruf_setup_kernel + ruf_raster_filter_kernel
//   GL_LESS depth test (src/urdf_filter.cpp:570) -> min-reduction in shared memory
//   include/shaders/urdf_filter.frag:14-35    -> fused epilogue of ruf_raster_filter_kernel
//   glGetTexImage of attachments 1 and 3 (:729-735) -> the epilogue's global stores
//
// Arithmetic contract (DESIGN.md "Raster specification"): every float operation is written
// in the exact order of the specification; fused multiply-adds only where fmaf() is spelled
// out.  This file MUST be compiled with -fmad=false (and without -use_fast_math) so that
// nvcc neither contracts nor reorders; divisions are IEEE (-prec-div=true is the default).
#include "ruf_device.cuh"

#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace ruf {

// ------------------------------------------------------------------------------------------
// small PTX helpers: mbarrier + bulk async copy (TMA engine, 1-D form)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// byte permute; selector nibble | 8 replicates the sign of the selected byte (PTX prmt, default mode)
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// 16-byte asynchronous copy global -> shared (LDGSTS); completion is per thread (cp.async.wait_all)
__device__ __forceinline__ void cp_async16(void *dst, const void *src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
  asm volatile("cp.async.wait_all;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// ------------------------------------------------------------------------------------------
// vertex stage + primitive setup: building blocks of K1
// ------------------------------------------------------------------------------------------
struct V4 { float x, y, z, w; };
struct WV { int32_t X, Y; float z; };

// S1: clip = MVP * (x,y,z,1); include/shaders/urdf_filter.vert:5
__device__ __forceinline__ V4 xform(const float4 &c0, const float4 &c1, const float4 &c2, const float4 &c3,
                                    float px, float py, float pz)
{
  V4 c;
  c.x = fmaf(c0.x, px, fmaf(c1.x, py, fmaf(c2.x, pz, c3.x)));
  c.y = fmaf(c0.y, px, fmaf(c1.y, py, fmaf(c2.y, pz, c3.y)));
  c.z = fmaf(c0.z, px, fmaf(c1.z, py, fmaf(c2.z, pz, c3.z)));
  c.w = fmaf(c0.w, px, fmaf(c1.w, py, fmaf(c2.w, pz, c3.w)));
  return c;
}
__device__ __forceinline__ bool finite4(const V4 &c)
{
  return isfinite(c.x) && isfinite(c.y) && isfinite(c.z) && isfinite(c.w);
}
// S3: signed distance to clip plane k (0 = near, 1..4 = guard band)
__device__ __forceinline__ float plane_dist(const V4 &c, int k, float gx, float gy)
{
  switch (k) {
    case 0: return c.z + c.w;
    case 1: return fmaf(gx, c.w, -c.x);
    case 2: return fmaf(gx, c.w, c.x);
    case 3: return fmaf(gy, c.w, -c.y);
    default: return fmaf(gy, c.w, c.y);
  }
}
__device__ __forceinline__ V4 clip_lerp(const V4 &in, const V4 &out, float din, float dout)
{
  float t = din / (din - dout);
  V4 r;
  r.x = fmaf(t, out.x - in.x, in.x);
  r.y = fmaf(t, out.y - in.y, in.y);
  r.z = fmaf(t, out.z - in.z, in.z);
  r.w = fmaf(t, out.w - in.w, in.w);
  return r;
}
// S4: divide, viewport, snap
__device__ __forceinline__ bool to_window(const V4 &c, float halfw, float halfh, WV &v)
{
  float iw = 1.0f / c.w;
  float nx = c.x * iw, ny = c.y * iw, nz = c.z * iw;
  float xw = fmaf(nx, halfw, halfw);
  float yw = fmaf(ny, halfh, halfh);
  float zw = fmaf(nz, 0.5f, 0.5f);
  if (!(fabsf(xw) <= kWindowLimit) || !(fabsf(yw) <= kWindowLimit) || !(fabsf(zw) <= kWindowLimit))
    return false;
  v.z = zw;
  v.X = __float2int_rn(xw * (float)kSubpix);
  v.Y = __float2int_rn(yw * (float)kSubpix);
  return true;
}

// S5/S8: orientation, pixel bbox, depth plane.  Returns false when the triangle cannot touch a pixel.
__device__ __forceinline__ bool setup_window_tri(WV a, WV b, WV c, const Dims &d, TriRec &r, bool *positive = nullptr)
{
  long long area2 = (long long)(b.X - a.X) * (c.Y - a.Y) - (long long)(c.X - a.X) * (b.Y - a.Y);
  if (area2 == 0) return false;
  if (positive) *positive = area2 > 0;
  if (area2 < 0) { WV t = b; b = c; c = t; area2 = -area2; }

  int xmin = min(a.X, min(b.X, c.X)), xmax = max(a.X, max(b.X, c.X));
  int ymin = min(a.Y, min(b.Y, c.Y)), ymax = max(a.Y, max(b.Y, c.Y));
  int i0 = (xmin - kSubpixHalf + (kSubpix - 1)) >> kSubpixBits;   // arithmetic shift = floor
  int i1 = (xmax - kSubpixHalf) >> kSubpixBits;
  int j0 = (ymin - kSubpixHalf + (kSubpix - 1)) >> kSubpixBits;
  int j1 = (ymax - kSubpixHalf) >> kSubpixBits;
  i0 = max(i0, 0); i1 = min(i1, d.W - 1);
  j0 = max(j0, 0); j1 = min(j1, d.H - 1);
  if (i0 > i1 || j0 > j1) return false;

  // S8: depth plane anchored at vertex 0
  float dx1 = (float)(b.X - a.X), dy1 = (float)(b.Y - a.Y);
  float dx2 = (float)(c.X - a.X), dy2 = (float)(c.Y - a.Y);
  float dz1 = b.z - a.z, dz2 = c.z - a.z;
  float fa = __ll2float_rn(area2);
  float t1 = dz2 * dy1;
  float gxz = fmaf(dz1, dy2, -t1) / fa;
  float t2 = dz1 * dx2;
  float gyz = fmaf(dz2, dx1, -t2) / fa;

  r.x0 = a.X; r.y0 = a.Y; r.x1 = b.X; r.y1 = b.Y; r.x2 = c.X; r.y2 = c.Y;
  r.z0 = a.z; r.gx = gxz; r.gy = gyz;
  r.bx = (uint32_t)i0 | ((uint32_t)i1 << 16);
  r.by = (uint32_t)j0 | ((uint32_t)j1 << 16);
  r.pad = 0;
  return true;
}

// setup_window_tri in two halves (RUF_EARLY_RESERVE): what the tile reservation needs -- orientation, pixel bbox -- and the
// depth plane with its two divisions, which then runs while the reservation atomics are in flight.  Same operations on the
// same operands as setup_window_tri: same bits.
__device__ __forceinline__ bool setup_cover(WV a, WV &b, WV &c, const Dims &d, long long &area2, bool &positive, int &i0,
                                            int &i1, int &j0, int &j1)
{
  area2 = (long long)(b.X - a.X) * (c.Y - a.Y) - (long long)(c.X - a.X) * (b.Y - a.Y);
  if (area2 == 0) return false;
  positive = area2 > 0;
  if (area2 < 0) { WV t = b; b = c; c = t; area2 = -area2; }
  const int xmin = min(a.X, min(b.X, c.X)), xmax = max(a.X, max(b.X, c.X));
  const int ymin = min(a.Y, min(b.Y, c.Y)), ymax = max(a.Y, max(b.Y, c.Y));
  i0 = max((xmin - kSubpixHalf + (kSubpix - 1)) >> kSubpixBits, 0);
  i1 = min((xmax - kSubpixHalf) >> kSubpixBits, d.W - 1);
  j0 = max((ymin - kSubpixHalf + (kSubpix - 1)) >> kSubpixBits, 0);
  j1 = min((ymax - kSubpixHalf) >> kSubpixBits, d.H - 1);
  return i0 <= i1 && j0 <= j1;
}
__device__ __forceinline__ void setup_plane(const WV &a, const WV &b, const WV &c, long long area2, float &gxz, float &gyz)
{
  const float dx1 = (float)(b.X - a.X), dy1 = (float)(b.Y - a.Y);
  const float dx2 = (float)(c.X - a.X), dy2 = (float)(c.Y - a.Y);
  const float dz1 = b.z - a.z, dz2 = c.z - a.z;
  const float fa = __ll2float_rn(area2);
  const float t1 = dz2 * dy1;
  gxz = fmaf(dz1, dy2, -t1) / fa;
  const float t2 = dz1 * dx2;
  gyz = fmaf(dz2, dx1, -t2) / fa;
}

// per-frame list read by every tile: triangles spanning many tiles and everything the clipper made
__device__ __forceinline__ void push_big(const TriRec &r, const Dims &d, TriRec *big, uint32_t *ctr)
{
  const uint32_t pos = atomicAdd(&ctr[kCtrBig], 1u);
  if (pos < d.cap_big) big[pos] = r;
  else atomicOr(&ctr[kCtrFlags], kFlagBigOverflow);
}

// (out of line, and it takes the few fields it needs BY VALUE: a reference to the kernel's Dims parameter would make every
// thread of the setup kernel copy the whole structure from the constant bank to its stack at kernel start)
__device__ __noinline__ void clip_and_emit(V4 p0, V4 p1, V4 p2, float halfw, float halfh, float guard_x, float guard_y, int W,
                                           int H, uint32_t cap_big, TriRec *big, uint32_t *ctr)
{
  Dims d{};
  d.halfw = halfw; d.halfh = halfh; d.guard_x = guard_x; d.guard_y = guard_y; d.W = W; d.H = H; d.cap_big = cap_big;
  V4 poly[kMaxPoly], tmp[kMaxPoly];
  poly[0] = p0; poly[1] = p1; poly[2] = p2;
  int n = 3;
  for (int k = 0; k < 5 && n > 0; ++k) {
    int m = 0;
    for (int i = 0; i < n; ++i) {
      V4 a = poly[i], b = poly[(i + 1 == n) ? 0 : i + 1];
      float da = plane_dist(a, k, d.guard_x, d.guard_y), db = plane_dist(b, k, d.guard_x, d.guard_y);
      bool ia = da >= 0.0f, ib = db >= 0.0f;
      if (ia) tmp[m++] = a;
      if (ia != ib) tmp[m++] = ia ? clip_lerp(a, b, da, db) : clip_lerp(b, a, db, da);
    }
    n = m;
    for (int i = 0; i < n; ++i) poly[i] = tmp[i];
  }
  if (n < 3) return;
  for (int i = 0; i < n; ++i)
    if (!(poly[i].w > 0.0f)) return;          // S4b
  WV wv[kMaxPoly];
  for (int i = 0; i < n; ++i)
    if (!to_window(poly[i], d.halfw, d.halfh, wv[i])) return;
  for (int i = 2; i < n; ++i) {
    TriRec r;
    if (setup_window_tri(wv[0], wv[i - 1], wv[i], d, r)) push_big(r, d, big, ctr);
  }
}

// ------------------------------------------------------------------------------------------
// K0: MVP table + per-part view-volume culling.
//
// gl_ModelViewProjectionMatrix of every drawn part, composed in double in the order the GL matrix stack
// does (PROJECTION * MODELVIEW, MODELVIEW = view * part_model) and rounded once to float.  Row n_parts is
// the background quad: P * LookAt (src/urdf_filter.cpp:576-596).  One thread per matrix element; a block
// of 256 threads owns 16 matrices.
//
// Culling (result-neutral, DESIGN.md "Setup-stage rejects"): the 8 corners of a part's object-space box
// go through the part's MVP; if all 8 lie beyond one clip plane by a margin, every triangle of the part
// would be rejected one by one by the setup kernel, so vis[frame][part] = 0 lets it skip the part's
// vertices wholesale.  8 threads per matrix.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 8)
ruf_pose_kernel(const double *__restrict__ proj, const double *__restrict__ view, const double *__restrict__ part_model,
                const double *__restrict__ lookat, const float *__restrict__ part_aabb, int n_parts, int n_frames,
                float *__restrict__ mvp, uint8_t *__restrict__ vis, uint32_t *__restrict__ clear, long long n_clear,
                const uint32_t *__restrict__ bg_seed, uint32_t *__restrict__ big_words, int ctr_stride, uint32_t cap_big)
{
  __shared__ __align__(16) float s_m[16][16];
  const int rows = n_parts + 1;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_mats = (long long)n_frames * rows;
  // The block's inputs -- its 16 part matrices with the view matrices of their frames, the projection matrix, LookAt --
  // are requested first, at most three loads per thread, all in flight together, and reach the arithmetic through shared
  // memory: in the single-frame graph they are read straight from the host's pinned block, and a thread that fetched
  // its 36 operands one after the other (32 registers) paid several PCIe round trips.  Same products in the same order.
  __shared__ double s_in[16][16], s_V[16][16], s_P[16], s_L[16];
  double r_in = 0.0, r_v = 0.0, r_pl = 0.0;
  {
    const int t = threadIdx.x;
    const long long matl = (long long)blockIdx.x * 16 + (t >> 4);
    if (matl < n_mats) {
      const int p = (int)(matl % rows);
      const long long f = matl / rows;
      if (p < n_parts) {
        r_in = part_model[16 * (f * n_parts + p) + (t & 15)];
        r_v = view[16 * f + (t & 15)];
      }
    }
    if (t < 16) r_pl = proj[t];
    else if (t < 32) r_pl = lookat[t & 15];
  }
  // single-frame graph: the frame's counter block is cleared here instead of by a memset node of its own; with a seed,
  // every frame's big list starts with the background quad's records (they depend on the projection matrix alone and
  // were set up once: ruf_api.cu fill_bg_seed) instead of waiting for one warp of the setup kernel to clip the quad again
  const uint32_t n_seed = bg_seed ? min(__ldg(bg_seed), min((uint32_t)kBgSeedMax, cap_big)) : 0u;
  for (long long i = gid; i < n_clear; i += (long long)gridDim.x * blockDim.x)
    clear[i] = (n_seed && i % ctr_stride == kCtrBig) ? n_seed : 0u;
  if (n_seed) {
    const long long per = (long long)n_seed * (sizeof(TriRec) / 4);
    for (long long i = gid; i < per * n_frames; i += (long long)gridDim.x * blockDim.x) {
      const long long f = i / per, w = i - f * per;
      big_words[f * (long long)cap_big * (sizeof(TriRec) / 4) + w] = __ldg(bg_seed + 4 + w);
    }
  }
  {
    const int t = threadIdx.x;
    s_in[t >> 4][t & 15] = r_in;
    s_V[t >> 4][t & 15] = r_v;
    if (t < 16) s_P[t] = r_pl;
    else if (t < 32) s_L[t & 15] = r_pl;
  }
  __syncthreads();
  float val = 0.0f;
  if (gid < n_mats * 16) {
    const int e = (int)(gid & 15);
    const long long mat = gid >> 4;
    const int p = (int)(mat % rows);
    const int r = e & 3, c = e >> 2;
    double s;
    if (p == n_parts) {
      s = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) s += s_P[k * 4 + r] * s_L[c * 4 + k];
    } else {
      const double *V = s_V[threadIdx.x >> 4];
      const double *M = s_in[threadIdx.x >> 4];
      double pv[4];   // row r of P*V
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) a += s_P[j * 4 + r] * V[k * 4 + j];
        pv[k] = a;
      }
      s = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) s += pv[k] * M[c * 4 + k];
    }
    val = (float)s;
    mvp[gid] = val;
  }
  s_m[threadIdx.x >> 4][threadIdx.x & 15] = val;
  __syncthreads();
  if (threadIdx.x < 128) {
    const int mi = threadIdx.x >> 3, corner = threadIdx.x & 7;
    const long long mat = (long long)blockIdx.x * 16 + mi;
    const bool have = mat < n_mats;
    const int p = have ? (int)(mat % rows) : n_parts;
    uint32_t out = 0;        // bit k: this corner is beyond plane k (near, +x, -x, +y, -y)
    if (p < n_parts) {
      const float *bb = part_aabb + kPartStride * (size_t)p;
      const float px = __ldg(bb + ((corner & 1) ? 3 : 0)), py = __ldg(bb + ((corner & 2) ? 4 : 1)),
                  pz = __ldg(bb + ((corner & 4) ? 5 : 2));
      const float4 *M = reinterpret_cast<const float4 *>(s_m[mi]);
      const V4 c = xform(M[0], M[1], M[2], M[3], px, py, pz);
      if (finite4(c)) {
        // margins: 1 % on the side planes, 1e-3 relative on the near plane -- far larger than the
        // float rounding of the per-triangle tests they stand in for
        const float wm = 1.01f * c.w, tol = 1e-3f * (fabsf(c.z) + fabsf(c.w));
        out = (c.z + c.w < -tol ? 1u : 0u) | (c.x > wm ? 2u : 0u) | (-c.x > wm ? 4u : 0u) |
              (c.y > wm ? 8u : 0u) | (-c.y > wm ? 16u : 0u);
      }
    }
    bool rejected = false;
    const int group = (threadIdx.x & 31) >> 3;
#pragma unroll
    for (int k = 0; k < 5; ++k)
      rejected |= ((__ballot_sync(0xffffffffu, (out >> k) & 1u) >> (8 * group)) & 0xffu) == 0xffu;
    if (have && corner == 0) {
      // Which triangles face the camera?  With A = rows (x, y, w) x columns (0..2) of the MVP, the window-space
      // doubled area of a triangle (a, b, c) has the sign of det(A) * (a - eye) . ((b - a) x (c - a)): for a mesh
      // wound counter-clockwise seen from outside, positive area means "facing away" iff det(A) > 0.  Only a
      // drawing-order hint for the raster kernel (front first, then the rest depth-culled).
      const float *m = s_m[mi];
      const float det = m[0] * (m[5] * m[11] - m[9] * m[7]) - m[4] * (m[1] * m[11] - m[9] * m[3]) +
                        m[8] * (m[1] * m[7] - m[5] * m[3]);
      const bool inward = (p < n_parts) && __ldg(part_aabb + kPartStride * (size_t)p + 6) < 0.0f;
      const bool pos_is_front = (det < 0.0f) != inward;
      vis[mat] = (uint8_t)((rejected ? 0 : 1) | (pos_is_front ? 2 : 0));    // the background row is never culled (out = 0)
    }
  }
}

// ------------------------------------------------------------------------------------------
// K1: vertex stage + primitive setup + binning.  One CTA per (meshlet, run of frames).
//
// The model is stored as meshlets (built once on the host, ruf_meshlet.cpp): runs of consecutive
// triangles whose bit-identical vertices are welded, so that the vertex shader
// (include/shaders/urdf_filter.vert:5), the clip tests, the perspective divide and the viewport
// snap run ONCE per distinct vertex instead of once per triangle corner.  The per-vertex result
// is a pure function of (MVP, position), so sharing it cannot change a bit of the output.
//
// A CTA keeps its meshlet (positions in REGISTERS, index triples in shared memory) and loops over `frames_per_cta` frames;
// per frame only the meshlet's MVP rows (contiguous, staged in shared memory one frame ahead) and
// its parts' cull bytes are read.  Per frame:
//   P1  thread per vertex:    clip = MVP * v, clip-plane flags, window coordinates -> shared memory
//   --  the only CTA-wide barrier of the frame; from here on every warp works on its own triangles --
//   P2  lane per triangle:    flag logic (reject / clip / keep), zero-area and empty-bbox tests; survivors
//                             and the rare clip candidates are compacted into the warp's list (ballots only)
//   S3  lane per clip candidate: Sutherland-Hodgman, set-up, append to the frame's big list
//   P3  lane per SURVIVOR (dense): full setup (depth plane: two IEEE divisions), record in registers
//   P4  the lanes walk the tiles of their bboxes; lanes on the same tile (match.any) share ONE global
//       atomic that reserves a run of that tile's record list, and write their records there.
//       The raster kernel reads ONE contiguous list per tile.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kVfDead = 1u;     // part culled or non-finite clip coordinate: the triangle is dropped
constexpr uint32_t kVfNear = 2u;     // z + w < 0
constexpr uint32_t kVfXp = 4u;       //  x > 1.001 w
constexpr uint32_t kVfXn = 8u;       // -x > 1.001 w
constexpr uint32_t kVfYp = 16u;      //  y > 1.001 w
constexpr uint32_t kVfYn = 32u;      // -y > 1.001 w
constexpr uint32_t kVfClip = 64u;    // outside the near plane or the guard band: the triangle goes through the clipper
constexpr uint32_t kVfBadW = 128u;   // S4b: !(w > 0) or window coordinates out of range
constexpr uint32_t kVfAllOut = kVfNear | kVfXp | kVfXn | kVfYp | kVfYn;
constexpr uint32_t kNoTri = 0xffffffffu;

// S1 + the per-vertex part of the setup-stage rejects, S3's inside tests and S4 for one vertex.
// -> (X, Y, bits(z_w), flags)
__device__ __forceinline__ uint4 vertex_stage(const float4 &c0, const float4 &c1, const float4 &c2, const float4 &c3,
                                              float x, float y, float z, const Dims &d)
{
  const V4 p = xform(c0, c1, c2, c3, x, y, z);
  if (!finite4(p)) return make_uint4(0u, 0u, 0u, kVfDead);
  uint32_t fl = 0;
  // Early outs that cannot change the result (DESIGN.md "Setup-stage rejects"): a triangle is skipped when
  //  * all three vertices are in front of the near plane: the clipper would return nothing;
  //  * all three are beyond one side plane of the view volume by a 0.1 % margin: whatever the
  //    clipper keeps lies at least 0.3 px outside the viewport.
  if (p.z + p.w < 0.0f) fl |= kVfNear;
  const float kw = 1.001f * p.w;
  if (p.x > kw) fl |= kVfXp;
  if (-p.x > kw) fl |= kVfXn;
  if (p.y > kw) fl |= kVfYp;
  if (-p.y > kw) fl |= kVfYn;
  bool need = false;
#pragma unroll
  for (int k = 0; k < 5; ++k)
    if (!(plane_dist(p, k, d.guard_x, d.guard_y) >= 0.0f)) need = true;
  WV w;
  w.X = 0; w.Y = 0; w.z = 0.0f;
  if (need) fl |= kVfClip;
  else if (!(p.w > 0.0f) || !to_window(p, d.halfw, d.halfh, w)) fl |= kVfBadW;
  return make_uint4((uint32_t)w.X, (uint32_t)w.Y, __float_as_uint(w.z), fl);   // the caller adds the part slot << 8
}

// S5 (zero area) and the pixel bbox of S6: can the triangle touch a sample at all?
__device__ __forceinline__ bool tri_may_touch(const uint4 &a, const uint4 &b, const uint4 &c, const Dims &d)
{
  const int ax = (int)a.x, ay = (int)a.y, bx = (int)b.x, by = (int)b.y, cx = (int)c.x, cy = (int)c.y;
  const long long area2 = (long long)(bx - ax) * (cy - ay) - (long long)(cx - ax) * (by - ay);
  if (area2 == 0) return false;
  const int xmin = min(ax, min(bx, cx)), xmax = max(ax, max(bx, cx));
  const int ymin = min(ay, min(by, cy)), ymax = max(ay, max(by, cy));
  const int i0 = max((xmin - kSubpixHalf + (kSubpix - 1)) >> kSubpixBits, 0);
  const int i1 = min((xmax - kSubpixHalf) >> kSubpixBits, d.W - 1);
  const int j0 = max((ymin - kSubpixHalf + (kSubpix - 1)) >> kSubpixBits, 0);
  const int j1 = min((ymax - kSubpixHalf) >> kSubpixBits, d.H - 1);
  return i0 <= i1 && j0 <= j1;
}

#ifdef RUF_X_TIMELINE
__device__ unsigned long long g_timeline1[1024 * 8 * 8];      // setup kernel: [cta][warp][mark]
__device__ __forceinline__ void tl1_mark(int slot)
{
  if ((threadIdx.x & 31) == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
    if (cta < 1024) g_timeline1[(cta * 8 + (threadIdx.x >> 5)) * 8 + slot] = t;
  }
}
#define TL1(x) tl1_mark(x)
#else
#define TL1(x)
#endif
__global__ void __launch_bounds__(kSetupThreads, RUF_SETUP_MIN_BLOCKS)
ruf_setup_bin_kernel(Model m, const float *__restrict__ mvp_all, const uint8_t *__restrict__ vis_all, Dims d,
                     int n_frames, int frames_per_cta, TriRec *big_all, BinRec *bins_all, uint32_t *ctr_all)
{
  constexpr int kVPT = (kMeshVerts + kSetupThreads - 1) / kSetupThreads;   // vertices per thread
  constexpr int kTPT = kSetupSlots;                                       // triangles per thread
  constexpr int kWarpList = kTPT * 32;                                    // triangles a warp classifies per frame
  __shared__ uint4 s_vert[2][kMeshVerts];        // X, Y, bits(z_w), flags; double-buffered over frames
  __shared__ uint32_t s_list[kSetupThreads / 32][kWarpList];   // per warp: survivors from the front, clip candidates from the back
  __shared__ __align__(16) float s_mvp[3][kMeshParts * 16];    // this meshlet's MVP rows, staged two frames ahead
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lanemask_lt = (1u << lane) - 1u;

  TL1(0);
  const uint4 hdr = __ldg(m.meshlets + blockIdx.x);
  const uint32_t vert_off = hdr.x, tri_off = hdr.y, part_lo = hdr.w;
  const int nverts = (int)(hdr.z & 1023u), ntris = (int)((hdr.z >> 10) & 1023u), npm1 = (int)(hdr.z >> 20);
  const int rows = d.n_parts + 1;
  const int nmv = (npm1 + 1) * 16;               // floats of the meshlet's MVP rows (consecutive parts: contiguous)
  const int f0 = blockIdx.y * frames_per_cta, f1 = min(n_frames, f0 + frames_per_cta);

  // nothing of this meshlet visible in any frame of the run (per-part cull bytes of the pose kernel)? leave
  // before touching the model.  Every warp decides by itself from the same bytes: consistent across the CTA.
  {
    // (all frames' bytes are requested before the first one is looked at: one round trip instead of frames_per_cta)
    uint32_t vb[kSetupFrames];
#pragma unroll
    for (int j = 0; j < kSetupFrames; ++j)
      vb[j] = (f0 + j < f1 && lane <= npm1) ? (uint32_t)__ldg(vis_all + (size_t)(f0 + j) * rows + part_lo + lane) : 0u;
    bool any = false;
#pragma unroll
    for (int j = 0; j < kSetupFrames; ++j) any |= (vb[j] & 1u) != 0u;
    for (int f = f0 + kSetupFrames; f < f1; ++f)       // frames_per_cta above kSetupFrames (RUF_SETUP_FRAMES_FORCE)
      any |= (lane <= npm1) && (__ldg(vis_all + (size_t)f * rows + part_lo + lane) & 1) != 0;
    if (!__any_sync(0xffffffffu, any)) return;
  }
  TL1(1);
  // the meshlet stays on chip for all frames of this CTA: vertex positions in registers, index triples in shared memory
  float4 vq[kVPT];
  __shared__ uint32_t s_tri[kTPT * kSetupThreads];     // the meshlet's packed index triples (registers are the scarce resource: 40 per thread)
#pragma unroll
  for (int k = 0; k < kVPT; ++k) {
    const int v = k * kSetupThreads + tid;
    vq[k] = (v < nverts) ? __ldg(m.verts + vert_off + v) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < kTPT; ++k) {
    const int t = k * kSetupThreads + tid;
    s_tri[t] = (t < ntris) ? __ldg(m.tris + tri_off + t) : kNoTri;
  }

  // stage the matrices of the first two frames; cull bytes -> one bit per part (every warp builds its own copy)
  uint32_t pvis_next, pfront_next;
  {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (f0 + j < f1) {
        const float *src = mvp_all + 16 * ((size_t)(f0 + j) * rows + part_lo);
        for (int i = tid; i < nmv; i += kSetupThreads) s_mvp[j][i] = __ldg(src + i);
      }
    }
    const uint32_t vb = (lane <= npm1) ? (uint32_t)__ldg(vis_all + (size_t)f0 * rows + part_lo + lane) : 0u;
    pvis_next = __ballot_sync(0xffffffffu, (vb & 1u) != 0u);
    pfront_next = __ballot_sync(0xffffffffu, (vb & 2u) != 0u);
  }
  __syncthreads();
  TL1(2);

  // One barrier per frame (B1: the frame's vertices are in shared memory).  Between two B1 the warps run
  // independently: a warp classifies ITS triangles, compacts ITS survivors and reserves list space itself.
  for (int f = f0; f < f1; ++f) {
    const int it = f - f0, vbuf = it & 1, mbuf = it % 3;
    const uint32_t pvis_bits = pvis_next;        // bit i: part part_lo + i may be visible in frame f
    const uint32_t pfront_bits = pfront_next;    // bit i: that part's positive-area triangles face the camera
    // prefetch the cull bytes of frame f + 1
    uint32_t vb_next = 0;
    if (f + 1 < f1 && lane <= npm1) vb_next = (uint32_t)__ldg(vis_all + (size_t)(f + 1) * rows + part_lo + lane);
    pvis_next = __ballot_sync(0xffffffffu, (vb_next & 1u) != 0u);
    pfront_next = __ballot_sync(0xffffffffu, (vb_next & 2u) != 0u);

    if (pvis_bits != 0) {
      // ---- P1: vertex stage, once per welded vertex ----
#pragma unroll
      for (int k = 0; k < kVPT; ++k) {
        const int v = k * kSetupThreads + tid;
        if (v < nverts) {
          const uint32_t slot = __float_as_uint(vq[k].w);
          uint4 r = make_uint4(0u, 0u, 0u, kVfDead);
          if ((pvis_bits >> (slot & 31u)) & 1u) {
            const float4 *M = reinterpret_cast<const float4 *>(&s_mvp[mbuf][16 * slot]);
            r = vertex_stage(M[0], M[1], M[2], M[3], vq[k].x, vq[k].y, vq[k].z, d);
            r.w |= slot << 8;
          }
          s_vert[vbuf][v] = r;
        }
      }
    }
    cp_async_wait_all();                           // this thread's share of the matrices of frame f + 1 has landed
    __syncthreads();                               // B1 (also taken when the whole meshlet is culled in this frame)
    TL1(3);
    // matrices of frame f + 2 -> shared memory (cp.async: no registers held).  Their buffer was last read in
    // frame f - 1, which every warp has left by now; they are read after B1 of frame f + 1.
    if (f + 2 < f1) {
      const float *src = mvp_all + 16 * ((size_t)(f + 2) * rows + part_lo);
      float *dst = s_mvp[(it + 2) % 3];
      for (int i = tid; i < nmv / 4; i += kSetupThreads) cp_async16(dst + 4 * i, src + 4 * i);
    }
    if (pvis_bits != 0) {
      TriRec *big = big_all + (size_t)f * d.cap_big;
      uint32_t *ctr = ctr_all + (size_t)f * d.ctr_stride;
      const uint4 *sv = s_vert[vbuf];
      uint32_t *list = s_list[warp];

      // ---- P2: per triangle: flag logic, cheap rejects; survivors and clip candidates compacted per warp ----
      int nkeep = 0, nclip = 0;                    // warp-uniform
#pragma unroll
      for (int k = 0; k < kTPT; ++k) {
        bool keep = false, clip = false;
        const uint32_t ixk = s_tri[k * kSetupThreads + tid];       // written by this very thread: no barrier needed
        if (ixk != kNoTri) {
          const uint4 a = sv[ixk & 1023u], b = sv[(ixk >> 10) & 1023u], c = sv[ixk >> 20];
          const uint32_t f_or = a.w | b.w | c.w, f_and = a.w & b.w & c.w;
          if (!(f_or & kVfDead) && !(f_and & kVfAllOut)) {
            if (f_or & kVfClip) clip = true;
            else if (!(f_or & kVfBadW)) keep = tri_may_touch(a, b, c, d);
          }
        }
        const unsigned bk = __ballot_sync(0xffffffffu, keep), bc = __ballot_sync(0xffffffffu, clip);
        if (keep) list[nkeep + __popc(bk & lanemask_lt)] = ixk;
        if (clip) list[kWarpList - 1 - nclip - __popc(bc & lanemask_lt)] = ixk;
        nkeep += __popc(bk);
        nclip += __popc(bc);
      }
      __syncwarp();
      TL1(4);

      // ---- S3 for the rare triangles that cross the near plane or the guard band: recompute the three
      // clip-space vertices (same inputs, same bits as P1), clip, set up, append to the frame's big list ----
      for (int s = lane; s < nclip; s += 32) {
        const uint32_t ixs = list[kWarpList - 1 - s];
        const float4 qa = __ldg(m.verts + vert_off + (ixs & 1023u)), qb = __ldg(m.verts + vert_off + ((ixs >> 10) & 1023u)),
                     qc = __ldg(m.verts + vert_off + (ixs >> 20));
        const float4 *M = reinterpret_cast<const float4 *>(&s_mvp[mbuf][16 * __float_as_uint(qa.w)]);
        const float4 c0 = M[0], c1 = M[1], c2 = M[2], c3 = M[3];
        clip_and_emit(xform(c0, c1, c2, c3, qa.x, qa.y, qa.z), xform(c0, c1, c2, c3, qb.x, qb.y, qb.z),
                      xform(c0, c1, c2, c3, qc.x, qc.y, qc.z), d.halfw, d.halfh, d.guard_x, d.guard_y, d.W, d.H, d.cap_big, big, ctr);
      }

      TL1(5);
      // ---- P3 + P4: one lane per SURVIVOR (dense): full setup, then the warp reserves list space per tile ----
      BinRec *bins = bins_all + (size_t)f * d.ntiles * d.cap_tile;
      for (int s0 = 0; s0 < nkeep; s0 += 32) {
        const int s = s0 + lane;
        int tx0 = 0, tx1 = -1, ty0 = 0, ty1 = -1;
        bool has = false, back = false;            // back: the triangle faces away from the camera (drawing-order hint)
#if RUF_EARLY_RESERVE
        // first only what the reservation needs (orientation, bbox, tile range); the depth plane follows below, under
        // the reservation atomics
        WV wa, wb, wc;
        long long area2 = 0;
        int bi0 = 0, bi1 = 0, bj0 = 0, bj1 = 0;
        bool isbig = false;
        wa.X = wa.Y = wb.X = wb.Y = wc.X = wc.Y = 0; wa.z = wb.z = wc.z = 0.f;
        if (s < nkeep) {
          const uint32_t ixs = list[s];
          const uint4 a = sv[ixs & 1023u], b = sv[(ixs >> 10) & 1023u], c = sv[ixs >> 20];
          wa.X = (int)a.x; wa.Y = (int)a.y; wa.z = __uint_as_float(a.z);
          wb.X = (int)b.x; wb.Y = (int)b.y; wb.z = __uint_as_float(b.z);
          wc.X = (int)c.x; wc.Y = (int)c.y; wc.z = __uint_as_float(c.z);
          bool positive = false;
          if (setup_cover(wa, wb, wc, d, area2, positive, bi0, bi1, bj0, bj1)) {
            back = positive != (((pfront_bits >> ((a.w >> 8) & 31u)) & 1u) != 0u);
            tx0 = bi0 / kTileW; tx1 = bi1 / kTileW;
            ty0 = bj0 / kTileH; ty1 = bj1 / kTileH;
            if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > kBigTiles) isbig = true;     // read by every tile
            else has = true;
          }
        }
        if (isbig) {                               // rare: the whole record at once
          TriRec rec;
          rec.x0 = wa.X; rec.y0 = wa.Y; rec.x1 = wb.X; rec.y1 = wb.Y; rec.x2 = wc.X; rec.y2 = wc.Y; rec.z0 = wa.z;
          setup_plane(wa, wb, wc, area2, rec.gx, rec.gy);
          rec.bx = (uint32_t)bi0 | ((uint32_t)bi1 << 16); rec.by = (uint32_t)bj0 | ((uint32_t)bj1 << 16); rec.pad = 0;
          push_big(rec, d, big, ctr);
        }
        const unsigned act = __ballot_sync(0xffffffffu, has);
        if (!act) continue;
        if (lane == 0) atomicAdd(&ctr[kCtrKept], (uint32_t)__popc(act));     // statistics only
        uint4 q0, q1;
#else
        TriRec rec;
        if (s < nkeep) {
          const uint32_t ixs = list[s];
          const uint4 a = sv[ixs & 1023u], b = sv[(ixs >> 10) & 1023u], c = sv[ixs >> 20];
          WV wa, wb, wc;
          wa.X = (int)a.x; wa.Y = (int)a.y; wa.z = __uint_as_float(a.z);
          wb.X = (int)b.x; wb.Y = (int)b.y; wb.z = __uint_as_float(b.z);
          wc.X = (int)c.x; wc.Y = (int)c.y; wc.z = __uint_as_float(c.z);
          bool positive = false;
          if (setup_window_tri(wa, wb, wc, d, rec, &positive)) {
            back = positive != (((pfront_bits >> ((a.w >> 8) & 31u)) & 1u) != 0u);
            tx0 = (int)(rec.bx & 0xffffu) / kTileW; tx1 = (int)(rec.bx >> 16) / kTileW;
            ty0 = (int)(rec.by & 0xffffu) / kTileH; ty1 = (int)(rec.by >> 16) / kTileH;
            if ((tx1 - tx0 + 1) * (ty1 - ty0 + 1) > kBigTiles) push_big(rec, d, big, ctr);     // read by every tile
            else has = true;
          }
        }
        const unsigned act = __ballot_sync(0xffffffffu, has);
        if (!act) continue;
        if (lane == 0) atomicAdd(&ctr[kCtrKept], (uint32_t)__popc(act));     // statistics only
        // BinRec: vertex 0, the other two vertices as 24-bit differences (12 bytes, three PRMTs), the depth plane
        const uint32_t dx1 = (uint32_t)(rec.x1 - rec.x0), dy1 = (uint32_t)(rec.y1 - rec.y0);
        const uint32_t dx2 = (uint32_t)(rec.x2 - rec.x0), dy2 = (uint32_t)(rec.y2 - rec.y0);
        const uint4 q0 = make_uint4((uint32_t)rec.x0, (uint32_t)rec.y0, prmt(dx1, dy1, 0x4210u), prmt(dy1, dx2, 0x5421u));
        const uint4 q1 = make_uint4(prmt(dx2, dy2, 0x6542u), __float_as_uint(rec.z0), __float_as_uint(rec.gx), __float_as_uint(rec.gy));
#endif
        // P4.  Every lane walks the tiles of its bbox (1 for three quarters of the records, 2 or 4 for most of
        // the rest); lanes that stand on the same tile in the same step (match.any) share ONE 8-byte global
        // atomic that reserves room in that tile's list: front count in the low word, back count in the high
        // word.  Front records fill the list from its start, back records from its end; the raster kernel flags
        // an overflow when the two runs meet.  The first kWalk steps are unrolled so that their atomics are in
        // flight together; the bases are consumed afterwards.
        const unsigned backmask = __ballot_sync(0xffffffffu, back);
        unsigned long long *tile_ctr = reinterpret_cast<unsigned long long *>(ctr + kCtrWords);
        auto put = [&](int tile, unsigned grp, unsigned long long base) {
          const uint32_t pos = back ? (uint32_t)(base >> 32) + (uint32_t)__popc(grp & backmask & lanemask_lt)
                                    : (uint32_t)base + (uint32_t)__popc(grp & ~backmask & lanemask_lt);
          if (pos < d.cap_tile) {
            const size_t slot = back ? (size_t)d.cap_tile - 1 - pos : (size_t)pos;
            uint4 *dst = reinterpret_cast<uint4 *>(bins + (size_t)tile * d.cap_tile + slot);
            dst[0] = q0; dst[1] = q1;
          }
        };
        constexpr int kWalk = RUF_WALK;
        int tx = tx0, ty = ty0;
        bool more = has;
        int tl[kWalk];
        unsigned grp[kWalk];
        unsigned long long base[kWalk];
        bool on[kWalk];
#pragma unroll
        for (int k = 0; k < kWalk; ++k) {
          on[k] = more;
          tl[k] = more ? ty * d.tiles_x + tx : -1 - lane;
          grp[k] = __match_any_sync(0xffffffffu, tl[k]);
          base[k] = 0;
          if (more && lane == __ffs(grp[k]) - 1)
            base[k] = atomicAdd(tile_ctr + tl[k], (unsigned long long)__popc(grp[k] & ~backmask) |
                                                      ((unsigned long long)__popc(grp[k] & backmask) << 32));
          if (more) {
            if (++tx > tx1) { tx = tx0; ++ty; }
            more = ty <= ty1;
          }
        }
#if RUF_EARLY_RESERVE
        {
          // the atomics are in flight: now the two divisions of the depth plane and the packing of the record
          float gxz = 0.f, gyz = 0.f;
          if (has) setup_plane(wa, wb, wc, area2, gxz, gyz);
          const uint32_t dx1 = (uint32_t)(wb.X - wa.X), dy1 = (uint32_t)(wb.Y - wa.Y);
          const uint32_t dx2 = (uint32_t)(wc.X - wa.X), dy2 = (uint32_t)(wc.Y - wa.Y);
          q0 = make_uint4((uint32_t)wa.X, (uint32_t)wa.Y, prmt(dx1, dy1, 0x4210u), prmt(dy1, dx2, 0x5421u));
          q1 = make_uint4(prmt(dx2, dy2, 0x6542u), __float_as_uint(wa.z), __float_as_uint(gxz), __float_as_uint(gyz));
        }
#endif
#pragma unroll
        for (int k = 0; k < kWalk; ++k) {
          const unsigned long long b = __shfl_sync(0xffffffffu, base[k], __ffs(grp[k]) - 1);
          if (on[k]) put(tl[k], grp[k], b);
        }
        while (__ballot_sync(0xffffffffu, more)) {      // records that touch more than kWalk tiles
          const int tile = more ? ty * d.tiles_x + tx : -1 - lane;
          const unsigned g = __match_any_sync(0xffffffffu, tile);
          unsigned long long b = 0;
          if (more && lane == __ffs(g) - 1)
            b = atomicAdd(tile_ctr + tile, (unsigned long long)__popc(g & ~backmask) | ((unsigned long long)__popc(g & backmask) << 32));
          b = __shfl_sync(0xffffffffu, b, __ffs(g) - 1);
          if (more) {
            put(tile, g, b);
            if (++tx > tx1) { tx = tx0; ++ty; }
            more = ty <= ty1;
          }
        }
      }
      __syncwarp();                                // the warp's list is rewritten in the next frame
      TL1(6);
    }
  }
}

// ------------------------------------------------------------------------------------------
// K4: tile rasteriser + fused fragment stage
// ------------------------------------------------------------------------------------------
struct Edges {
  int A0, B0, A1, B1, A2, B2;   // E_k(P) = A_k (Px - Xa_k) + B_k (Py - Ya_k)
  int bias0, bias1, bias2;      // 0 when an exactly-on-edge sample belongs to the edge, else -1
};
__device__ __forceinline__ Edges make_edges(const TriRec &r)
{
  Edges e;
  e.A0 = r.y0 - r.y1; e.B0 = r.x1 - r.x0;
  e.A1 = r.y1 - r.y2; e.B1 = r.x2 - r.x1;
  e.A2 = r.y2 - r.y0; e.B2 = r.x0 - r.x2;
  e.bias0 = ((e.A0 > 0) || (e.A0 == 0 && e.B0 > 0)) ? 0 : -1;   // S6 tie-break
  e.bias1 = ((e.A1 > 0) || (e.A1 == 0 && e.B1 > 0)) ? 0 : -1;
  e.bias2 = ((e.A2 > 0) || (e.A2 == 0 && e.B2 > 0)) ? 0 : -1;
  return e;
}
__device__ __forceinline__ float clamp_z(float z) { return (z > 0.0f) ? z : 0.0f; }

__device__ __forceinline__ TriRec load_rec_smem(const TriRec *p)
{
  const uint4 *s = reinterpret_cast<const uint4 *>(p);
  const uint4 q0 = s[0], q1 = s[1], q2 = s[2];
  TriRec r;
  r.x0 = (int)q0.x; r.y0 = (int)q0.y; r.x1 = (int)q0.z; r.y1 = (int)q0.w;
  r.x2 = (int)q1.x; r.y2 = (int)q1.y; r.z0 = __uint_as_float(q1.z); r.gx = __uint_as_float(q1.w);
  r.gy = __uint_as_float(q2.x); r.bx = q2.y; r.by = q2.z; r.pad = q2.w;
  return r;
}
// a binned record from the shared-memory ring: vertices restored, bbox fields left to the caller
__device__ __forceinline__ TriRec load_bin_smem(const BinRec *p)
{
  const uint4 *s = reinterpret_cast<const uint4 *>(p);
  const uint4 q0 = s[0], q1 = s[1];
  TriRec r;
  r.x0 = (int)q0.x; r.y0 = (int)q0.y;
  // three bytes each, the fourth replicated from the sign of the third: one PRMT per difference
  r.x1 = r.x0 + (int)prmt(q0.z, 0u, 0xa210u); r.y1 = r.y0 + (int)prmt(q0.z, q0.w, 0xd543u);
  r.x2 = r.x0 + (int)prmt(q0.w, q1.x, 0xc432u); r.y2 = r.y0 + (int)prmt(q1.x, 0u, 0xb321u);
  r.z0 = __uint_as_float(q1.y); r.gx = __uint_as_float(q1.z); r.gy = __uint_as_float(q1.w);
  r.bx = r.by = r.pad = 0u;
  return r;
}
__device__ __forceinline__ TriRec load_rec_global(const TriRec *p)
{
  const uint4 *s = reinterpret_cast<const uint4 *>(p);
  const uint4 q0 = __ldg(s), q1 = __ldg(s + 1), q2 = __ldg(s + 2);
  TriRec r;
  r.x0 = (int)q0.x; r.y0 = (int)q0.y; r.x1 = (int)q0.z; r.y1 = (int)q0.w;
  r.x2 = (int)q1.x; r.y2 = (int)q1.y; r.z0 = __uint_as_float(q1.z); r.gx = __uint_as_float(q1.w);
  r.gy = __uint_as_float(q2.x); r.bx = q2.y; r.by = q2.z; r.pad = q2.w;
  return r;
}

__device__ __forceinline__ uint32_t smem_add(uint32_t saddr, uint32_t v)
{
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(v) : "memory");
  return old;
}
// a whole warp covers the tile-clipped bbox in 8x4 footprints.  WIDE = false: every factor of the
// edge functions is below 2^14, so 32-bit arithmetic is exact; WIDE = true: 64-bit.
template <bool WIDE>
__device__ __forceinline__ void raster_warp(const TriRec &r, int i0, int i1, int j0, int j1, int tile_x0,
                                            int tile_y0, uint32_t *sz, int lane, int row_off = 0, int row_step = 4)
{
  typedef typename std::conditional<WIDE, long long, int>::type acc_t;
  const Edges e = make_edges(r);
  const int lx = lane & 7, ly = lane >> 3;
  for (int jb = j0 + row_off; jb <= j1; jb += row_step) {
    const int j = jb + ly;
    const int py = j * kSubpix + kSubpixHalf;
    const float rowz = fmaf(r.gy, (float)(py - r.y0), r.z0);
    const acc_t c0 = (acc_t)e.B0 * (py - r.y0) + e.bias0;
    const acc_t c1 = (acc_t)e.B1 * (py - r.y1) + e.bias1;
    const acc_t c2 = (acc_t)e.B2 * (py - r.y2) + e.bias2;
    for (int ib = i0; ib <= i1; ib += 8) {
      const int i = ib + lx;
      const int px = i * kSubpix + kSubpixHalf;
      const acc_t e0 = (acc_t)e.A0 * (px - r.x0) + c0;
      const acc_t e1 = (acc_t)e.A1 * (px - r.x1) + c1;
      const acc_t e2 = (acc_t)e.A2 * (px - r.x2) + c2;
      if (i <= i1 && j <= j1 && (e0 | e1 | e2) >= 0) {
        const uint32_t z = __float_as_uint(clamp_z(fmaf(r.gx, (float)(px - r.x0), rowz)));
        uint32_t *p = sz + ((j - tile_y0) * kTileW + (i - tile_x0));
        if (z < 0x3f800000u && z < *p) atomicMin(p, z);
      }
    }
  }
}

__device__ __forceinline__ TriRec shfl_rec(const TriRec &r, int src)
{
  TriRec o;
  o.x0 = __shfl_sync(0xffffffffu, r.x0, src); o.y0 = __shfl_sync(0xffffffffu, r.y0, src);
  o.x1 = __shfl_sync(0xffffffffu, r.x1, src); o.y1 = __shfl_sync(0xffffffffu, r.y1, src);
  o.x2 = __shfl_sync(0xffffffffu, r.x2, src); o.y2 = __shfl_sync(0xffffffffu, r.y2, src);
  o.z0 = __shfl_sync(0xffffffffu, r.z0, src); o.gx = __shfl_sync(0xffffffffu, r.gx, src);
  o.gy = __shfl_sync(0xffffffffu, r.gy, src); o.bx = __shfl_sync(0xffffffffu, r.bx, src);
  o.by = __shfl_sync(0xffffffffu, r.by, src); o.pad = 0;
  return o;
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t *bar, uint32_t n)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256) for the 8 float pixels a thread owns in a row: one full 32-byte
// sector per lane and instruction.  Two 128-bit accesses with a 32-byte lane stride touch every sector twice, half each
// time -- harmless in HBM / L2, but when the buffers are the caller's pinned host memory (single-frame graph) every half
// sector is a PCIe transaction of its own: 32FC1 frames took 440 us instead of 260 at 1280x960.  32-byte aligned.
__device__ __forceinline__ void ldg_nc_256(const void *p, uint4 &a, uint4 &b)
{
  asm volatile("ld.global.nc.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}
__device__ __forceinline__ void stg_256(void *p, const float (&v)[8])
{
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

// thread-block cluster: barrier over all threads of all CTAs (release / acquire: shared-memory writes made before it are
// visible to the other CTAs' ld.shared::cluster after it) and loads from another CTA's shared memory (DSMEM)
__device__ __forceinline__ void cluster_sync_all()
{
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_map(const void *p, uint32_t rank)
{
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ uint4 ld_cluster_v4(uint32_t a)
{
  uint4 v;
  asm volatile("ld.shared::cluster.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t ld_cluster_u32(uint32_t a)
{
  uint32_t v;
  asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}

#ifdef RUF_X_TIMELINE
// debugging aid (never in the shipped library): per-CTA timestamps of the raster kernel's phases
__device__ unsigned long long g_timeline[8192 * 8];
__device__ __forceinline__ void tl_mark(int slot)
{
  if (threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (cta < 8192) g_timeline[cta * 8 + slot] = t;
  }
}
#define TL(x) tl_mark(x)
// per-warp batch log: [cta][warp][slot] = (t_claim, t_ready, t_units, t_done | items << 48)
__device__ unsigned long long g_batchlog[1024 * 8 * 16 * 4];
__device__ __forceinline__ unsigned long long tl_now()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define BL_DECL unsigned long long bl_t0 = 0, bl_t1 = 0, bl_t2 = 0; int bl_items = 0; int bl_k = 0
#define BL(v) v = tl_now()
#define BL_END(pass)                                                                                         \
  do {                                                                                                       \
    const unsigned cta = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;                     \
    if (lane == 0 && cta < 1024 && bl_k < 16) {                                                              \
      unsigned long long *q = g_batchlog + ((size_t)(cta * 8 + warp) * 16 + bl_k) * 4;                       \
      q[0] = bl_t0; q[1] = bl_t1; q[2] = bl_t2; q[3] = (tl_now() - bl_t2) | ((unsigned long long)(bl_items | (pass << 15)) << 48);                \
    }                                                                                                        \
    ++bl_k;                                                                                                  \
  } while (0)
#else
#define TL(x)
#define BL_DECL
#define BL(v)
#define BL_END(pass)
#endif

__device__ __forceinline__ void consumer_bar_sync()
{
  __syncthreads();
}

// saturate_cast<ushort>(cvRound(x * 1000.f)) -- cv::Mat::convertTo(CV_16U, 1000.0), src/urdf_filter.cpp:311
__device__ __forceinline__ uint32_t f32_to_u16(float x)
{
  const float v = x * 1000.0f;
  if (!(v >= -2147483648.0f && v < 2147483648.0f)) return 0u;   // cvtss2si "indefinite" -> INT_MIN -> 0
  const int r = __float2int_rn(v);
  return (uint32_t)min(max(r, 0), 65535);
}

// ---- fragment stage helpers -----------------------------------------------------------------------------
// 16UC1 input: sensor = float(raw) * 0.001f (convertTo(CV_32F, 0.001), src/urdf_filter.cpp:288) is strictly increasing
// in raw, so `sensor > thr` (frag:23) is `raw > R` with R = the largest raw whose sensor value is <= thr (-1: every
// raw is filtered, 65535: none is; thr = NaN compares false for every raw, like the float compare).  floor(thr * 1000)
// is within one step of R (two roundings of relative size 2^-24 on values below 65536); the two probes settle it with
// the float expressions of the shader themselves.  tests/test_oracle_kat.py::test_u16_threshold_* checks every
// boundary of all 65536 raw values against the float compare.
__device__ __forceinline__ int u16_threshold(float thr)
{
  if (!(thr < 65.536f)) return 65535;                 // also NaN, +inf
  if (thr < 0.0f) return -1;
  int c = min(__float2int_rd(thr * 1000.0f), 65535);
  if ((float)c * 0.001f > thr) --c;
  else if (c < 65535 && !((float)(c + 1) * 0.001f > thr)) ++c;
  return c;
}
// 8 pixels of one image row that share ONE virtual depth (all background, or a never-drawn run), 16UC1: per
// 32-bit word two subtractions, one PRMT (sign -> 0xffff / 0 per pixel) and one LOP3 instead of the per-pixel
// int->float conversions and float compares.  R < 0x7fff0000 so that R - raw cannot wrap.
__device__ __forceinline__ void shade_row_u16_uniform(const uint4 sens, int R, uint32_t repl2, uint4 &o, uint2 &mq)
{
  const uint32_t w[4] = {sens.x, sens.y, sens.z, sens.w};
  uint32_t M[4], ow[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t dlo = (uint32_t)(R - (int)(w[j] & 0xffffu));     // negative <=> raw > R <=> filtered
    const uint32_t dhi = (uint32_t)(R - (int)(w[j] >> 16));
    M[j] = prmt(dlo, dhi, 0xffbbu);                                  // 0xffff per filtered pixel
    ow[j] = (w[j] & ~M[j]) | (repl2 & M[j]);                         // unfiltered: the input bits (round trip = identity)
  }
  o = make_uint4(ow[0], ow[1], ow[2], ow[3]);
  mq.x = prmt(M[0], M[1], 0x6420u);                                  // one 0 / 255 byte per pixel
  mq.y = prmt(M[2], M[3], 0x6420u);
}

// 8 mask bytes (0 / 255) -> one byte, bit i = pixel i (RUF_MASK_BITS).  The multiplication gathers bit 0 of the four bytes
// of a word in its top nibble: no two partial products meet, so there are no carries.
__device__ __forceinline__ uint32_t mask_bits(const uint2 mq)
{
  return (((mq.x & 0x01010101u) * 0x10204080u) >> 28) | ((((mq.y & 0x01010101u) * 0x10204080u) >> 28) << 4);
}
__device__ __forceinline__ void store_mask(const FrameBuffers &fb, size_t base, const uint2 mq)
{
  if (fb.mask_bits) fb.mask_out[base >> 3] = (uint8_t)mask_bits(mq);       // base is a multiple of 8 on the vector path
  else *reinterpret_cast<uint2 *>(fb.mask_out + base) = mq;
}

// mix(sensor, replace, a) = sensor * (1 - a) + replace * a (frag:29) on the exotic values of a 32FC1 image, as the GL driver
// computes it (measured on Mesa llvmpipe with the reference's shaders): the shader runs with denormals-are-zero and
// -0 + 0 = +0, so |sensor| < FLT_MIN reads as +0.0; +inf on a filtered pixel gives inf * 0 = NaN (that driver's 0xffc00000).
__device__ __forceinline__ float sensor_daz(float x) { return (fabsf(x) < 1.17549435e-38f) ? 0.0f : x; }
__device__ __forceinline__ float mix_filtered(float sensor, float replace)
{
  return (sensor > 3.40282347e+38f) ? __int_as_float((int)0xffc00000u) : replace;
}

struct FragOut { float depth; uint32_t mask; };
// include/shaders/urdf_filter.frag:19-35 for the fragment that survived GL_LESS.
__device__ __forceinline__ FragOut fragment(float sensor, float zwin, const ShaderParams &sp)
{
  FragOut o;
  if (zwin == 1.0f) { o.depth = 0.0f; o.mask = 0u; return o; }   // never drawn: clear colour (:566)
  const float virt = sp.k1 / (zwin - sp.k2);                      // frag:14-17,22
  const bool s = sensor > (virt - sp.max_diff);                   // frag:23
  o.depth = s ? mix_filtered(sensor, sp.replace_value) : sensor;  // frag:29, mix() with a in {0,1}
  o.mask = s ? 255u : 0u;                                         // frag:35 read back as UNSIGNED_BYTE
  return o;
}

// Scalar tail of the fragment stage: up to 8 pixels of one row when the vector path does not apply (image width not a
// multiple of 8, unaligned caller buffers).  Cold: a rolled loop (unrolled it was 40 % of the kernel's code), but inlined --
// as an out-of-line call it gave the kernel a stack frame and cost the hot record loop 6 % (profiles/r02_experiments.md).
template <int ENC>
__device__ __forceinline__ void shade_scalar(const FrameBuffers &fb, const ShaderParams &sp, size_t base, int n, const float *zw,
                                          int zstride)
{
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    const float z = zw[i * zstride];
    float sensor;
    if (ENC == 1) sensor = (float)static_cast<const uint16_t *>(fb.depth_in)[base + i] * 0.001f;
    else sensor = sensor_daz(static_cast<const float *>(fb.depth_in)[base + i]);
    const FragOut o = fragment(sensor, z, sp);
    if (ENC == 1) static_cast<uint16_t *>(fb.depth_out)[base + i] = (uint16_t)f32_to_u16(o.depth);
    else static_cast<float *>(fb.depth_out)[base + i] = o.depth;
    if (fb.mask_out) fb.mask_out[base + i] = (uint8_t)o.mask;
    if (fb.zbuf_out) fb.zbuf_out[base + i] = z;
  }
}

// Class of one big-list record against one tile: 0 = no sample of the tile can be covered, 1 = every sample is
// covered, 2 = mixed, 3 = every sample is covered with ONE depth (*zconst).  Edge values are linear over the tile's
// sample grid, so their min / max sit on its corners.  Shared by the tile-info kernel and the raster kernel.
__device__ __forceinline__ uint32_t classify_big_record(const TriRec &r, int tile_x0, int tile_y0, float *zconst)
{
  const int bi0 = (int)(r.bx & 0xffffu), bi1 = (int)(r.bx >> 16);
  const int bj0 = (int)(r.by & 0xffffu), bj1 = (int)(r.by >> 16);
  if (bi1 < tile_x0 || bi0 >= tile_x0 + kTileW || bj1 < tile_y0 || bj0 >= tile_y0 + kTileH) return 0u;
  const int tpx = tile_x0 * kSubpix + kSubpixHalf, tpy = tile_y0 * kSubpix + kSubpixHalf;
  const Edges e = make_edges(r);
  const long long spanx = (long long)(kTileW - 1) * kSubpix, spany = (long long)(kTileH - 1) * kSubpix;
  const long long t0 = (long long)e.A0 * (tpx - r.x0) + (long long)e.B0 * (tpy - r.y0) + e.bias0;
  const long long t1 = (long long)e.A1 * (tpx - r.x1) + (long long)e.B1 * (tpy - r.y1) + e.bias1;
  const long long t2 = (long long)e.A2 * (tpx - r.x2) + (long long)e.B2 * (tpy - r.y2) + e.bias2;
  const long long a0 = e.A0 * spanx, c0 = e.B0 * spany, a1 = e.A1 * spanx, c1 = e.B1 * spany,
                  a2 = e.A2 * spanx, c2 = e.B2 * spany;
  const long long mx0 = t0 + max(a0, 0LL) + max(c0, 0LL), mn0 = t0 + min(a0, 0LL) + min(c0, 0LL);
  const long long mx1 = t1 + max(a1, 0LL) + max(c1, 0LL), mn1 = t1 + min(a1, 0LL) + min(c1, 0LL);
  const long long mx2 = t2 + max(a2, 0LL) + max(c2, 0LL), mn2 = t2 + min(a2, 0LL) + min(c2, 0LL);
  if ((mx0 | mx1 | mx2) < 0) return 0u;
  if ((mn0 | mn1 | mn2) < 0) return 2u;
  // a covering record with a constant depth plane (the background quad: every vertex has the same window z, so both
  // gradients are exactly 0 and z(P) = fma(0, ., fma(0, ., z0)) = z0): the value stands in for the record
  if (r.gx == 0.0f && r.gy == 0.0f) { *zconst = clamp_z(r.z0); return 3u; }
  return 1u;
}

// Occlusion among the big-list records of one tile (result-neutral; every thread of the CTA calls it with its own
// record, cls = 0 for none).  A record that covers every sample of the tile and is drawn everywhere (cls 1 / 3,
// farthest corner in front of the far plane) hides every record whose NEAREST corner is not in front of its FARTHEST
// corner -- min() cannot change there.  z is monotone in both sample coordinates (every step is a correctly rounded
// fma), so its extremes over the tile's sample grid sit on the four corner samples, evaluated with the walk's own
// expressions.  The occluder with the smallest farthest corner is kept (lowest index on ties); returns the record's
// class, 0 if it is hidden.  Kept out of line: it only runs for frames with more than kOccludeMin big-list records.
__device__ __noinline__ uint32_t occlusion_filter(uint32_t cls, int x0, int y0, float z0, float gx, float gy, int tpx,
                                                  int tpy, int tid, uint32_t *s_zcut, uint32_t *s_zown)
{
  uint32_t zlo = 0, zhi = 0;
  if (cls) {
    const float fxa = (float)(tpx - x0), fxb = (float)(tpx + (kTileW - 1) * kSubpix - x0);
    const float rza = fmaf(gy, (float)(tpy - y0), z0);
    const float rzb = fmaf(gy, (float)(tpy + (kTileH - 1) * kSubpix - y0), z0);
    const uint32_t zaa = __float_as_uint(clamp_z(fmaf(gx, fxa, rza))), zba = __float_as_uint(clamp_z(fmaf(gx, fxb, rza)));
    const uint32_t zab = __float_as_uint(clamp_z(fmaf(gx, fxa, rzb))), zbb = __float_as_uint(clamp_z(fmaf(gx, fxb, rzb)));
    zlo = min(min(zaa, zba), min(zab, zbb));
    zhi = max(max(zaa, zba), max(zab, zbb));
  }
  const bool occluder = (cls == 1u || cls == 3u) && zhi < 0x3f800000u;
  if (tid == 0) { *s_zcut = 0xffffffffu; *s_zown = 0xffffffffu; }
  __syncthreads();
  if (occluder) atomicMin(s_zcut, zhi);
  __syncthreads();
  const uint32_t zc = *s_zcut;
  if (occluder && zhi == zc) atomicMin(s_zown, (uint32_t)tid);
  __syncthreads();
  if (cls && zlo >= zc && (uint32_t)tid != *s_zown) cls = 0u;
  return cls;
}

// ------------------------------------------------------------------------------------------
// K1b: one thread per (frame, tile), between the setup and the raster kernel.
//  * folds the frame's overflow flags into the context's sticky status word;
//  * recognises FLAT tiles: no binned record and at most kFlatMaxBig big-list records, none of which needs per-pixel
//    work (class 0 or 3).  Their single virtual depth goes through the shader's scalar part here, once per tile
//    instead of once per thread: to_linear_depth (frag:14-17,22), the threshold (frag:23) and -- 16UC1 -- its
//    integer form (u16_threshold).  tinfo = { flags, R or bits(thr), bits(z), 0 }.
// ------------------------------------------------------------------------------------------
template <int ENC>
__global__ void __launch_bounds__(128)
ruf_tile_info_kernel(Dims d, int n_frames, const TriRec *__restrict__ big_all, const uint32_t *__restrict__ ctr_all,
                     uint32_t *status, ShaderParams sp, uint4 *__restrict__ tinfo)
{
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n_frames * d.ntiles) return;
  const int frame = (int)(gid / d.ntiles), tile = (int)(gid - (long long)frame * d.ntiles);
  const uint32_t *ctr = ctr_all + (size_t)frame * d.ctr_stride;
  const uint2 nfb = __ldg(reinterpret_cast<const uint2 *>(ctr + kCtrWords) + tile);
  uint32_t flags = (nfb.x + nfb.y > d.cap_tile) ? kFlagBinOverflow : 0u;     // the two runs met: the host grows cap_tile and retries
  if (tile == 0) {
    flags |= __ldg(ctr + kCtrFlags);
    atomicAdd(status + 1, __ldg(ctr + kCtrKept));         // kept records of the launches since the last status read-back
  }
  if (flags) atomicOr(status, flags);
  uint4 ti = make_uint4(0u, 0u, 0u, 0u);
  const uint32_t nbig = min(__ldg(ctr + kCtrBig), d.cap_big);
  if (nfb.x + nfb.y == 0 && nbig <= (uint32_t)kFlatMaxBig) {
    const int tile_by = tile / d.tiles_x, tile_bx = tile - tile_by * d.tiles_x;
    const TriRec *big = big_all + (size_t)frame * d.cap_big;
    float zt = 1.0f;                                           // glClear depth
    bool flat = true;
    for (uint32_t b = 0; b < nbig; ++b) {
      const TriRec r = load_rec_global(big + b);
      float zc = 0.0f;
      const uint32_t cls = classify_big_record(r, tile_bx * kTileW, tile_by * kTileH, &zc);
      if (cls == 1u || cls == 2u) flat = false;
      if (cls == 3u) zt = fminf(zt, zc);
    }
    if (flat) {
      ti.x = kTileFlat;
      ti.z = __float_as_uint(zt);
      if (zt == 1.0f) {
        ti.x |= kTileUndrawn;                                  // clear colour: depth 0, mask 0 (:566)
      } else {
        const float thr = (sp.k1 / (zt - sp.k2)) - sp.max_diff;
        ti.y = (ENC == 1) ? (uint32_t)u16_threshold(thr) : __float_as_uint(thr);
      }
    }
  }
  tinfo[gid] = ti;
}

// MP ("multi-pass units"): records with more than kMaxUnits units stay on the unit path and are dealt out over several
// passes instead of being parked for the cooperative 8x4-footprint walk; chosen per launch by the host from the share of
// such records in the previous launch (ruf_api.cu: choose_multipass).  At 1280x960 the PR2-like model's slivers are
// mostly above 24 units: MP is 34 % faster on C3; on C2 (5 % such records) the plain variant is 4 % faster.
//
// CL ("cluster split", the low-latency variant for launches of one or a few frames, where a frame's few dozen busy tiles
// would leave most of the 148 SMs idle and the busiest tile is the critical path): a thread-block cluster of CL CTAs
// shares ONE tile.  Each CTA rasterises 1/CL of the tile's record list into its own shared-memory z tile; after a cluster
// barrier the CTA that owns a band of 64/CL tile rows reads those rows from all CL z tiles through distributed shared
// memory, takes the minimum (min is associative and commutative: the same z image as one CTA walking the whole list)
// and runs the fragment stage for them.  The depth cull of the back pass uses, per 4x4 block, the smallest of the CL
// partial block maxima: an upper bound of the block maximum of the merged image, so still result-neutral.
template <int ENC, bool MP, int CL>
__global__ void __launch_bounds__(kRasterThreads, RUF_RASTER_MIN_BLOCKS)
ruf_raster_filter_kernel(Dims d, const TriRec *__restrict__ big_all, const BinRec *__restrict__ bins_all,
                         const uint32_t *__restrict__ ctr_all, const uint4 *__restrict__ tinfo, ShaderParams sp, FrameBuffers fb,
                         uint32_t *status)
{
  // 8 warps rasterise and shade.  The tile's record list streams into a shared-memory ring by bulk async
  // copies (TMA): thread 0 starts the first kStages chunks, later refills are issued by whichever warp
  // takes the last batch of a chunk (there is no dedicated producer warp holding registers).
  // dynamic shared memory (more than the 48 KB static limit): record ring, then the unit tables
  extern __shared__ __align__(128) unsigned char s_raster_dyn[];
  BinRec (*sbuf)[kChunk] = reinterpret_cast<BinRec (*)[kChunk]>(s_raster_dyn);
  uint16_t (*s_units)[32 * kMaxUnits] =
      reinterpret_cast<uint16_t (*)[32 * kMaxUnits]>(s_raster_dyn + sizeof(BinRec) * kStages * kChunk);
  __shared__ __align__(16) uint32_t sz[kTilePix + kZPad];
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages];
  __shared__ uint32_t s_next[2];                 // batch claim counters of the two passes
  __shared__ int s_issued[kStages];              // latest chunk whose bulk copies were issued into each ring stage
  __shared__ uint32_t s_zblk[(kTileH / 4) * 16]; // maxima of the 4x4 blocks of the z tile (depth cull)
  __shared__ uint32_t s_zblk_cl[CL > 1 ? (kTileH / 4) * 16 : 1];   // CL: the smallest of the cluster's partial block maxima
  __shared__ uint8_t s_bigcls[kRasterThreads];
  __shared__ uint32_t s_nwide, s_nstat;
  __shared__ __align__(16) TriRec s_wide[kWideCap];   // wide records of this tile, rasterised by the whole CTA at the end
  __shared__ uint32_t s_zcut, s_zown;            // big-list occlusion: smallest farthest-corner z of a covering record, its index
  __shared__ float s_bigz[kRasterThreads];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // CL: grid.x = tiles_x * CL with cluster dimensions (CL, 1, 1): blockIdx.x % CL is the rank in the cluster
  const int tbx = CL > 1 ? (int)(blockIdx.x / CL) : (int)blockIdx.x;
  const uint32_t crank = CL > 1 ? blockIdx.x % CL : 0u;
  const int frame = blockIdx.z, tile = blockIdx.y * d.tiles_x + tbx;
  const int tile_x0 = tbx * kTileW, tile_y0 = blockIdx.y * kTileH;
  const int prow = tid >> 3, pcol = (tid & 7) * 8;
  // CL: this thread runs the fragment stage (its rows prow and prow + 32 belong to this CTA's band); warp-uniform
  const bool frag = CL == 1 || (uint32_t)(prow / (32 / CL)) == crank;
  // records per warp batch: a full warp of 32; RUF_CLUSTER_BATCH lets the cluster-split variant take smaller batches (a
  // lone warp on a nearly empty SM needs ~0.5 us per round of 32 units, but also ~1.3 us per batch whatever its size:
  // 8 / 16 / 32 measured within 1 us of each other on the single-frame call)
  constexpr uint32_t RB = CL > 1 ? (uint32_t)kClusterBatch : 32u;
  TL(0);
  // CL, single-frame graph: the last CTA of the launch writes the status words straight into the host's pinned copy and
  // restarts the statistics -- no read-back copy and no memset node behind the kernel.  Every CTA calls this once, on
  // whichever path it leaves (flat tiles write no status word, so they may take their ticket early).
  auto export_status = [&]() {
    if (CL > 1 && fb.host_status && tid == 0) {
      __threadfence();
      if (atomicAdd(status + 3, 1u) == gridDim.x * gridDim.y * gridDim.z - 1u) {
        __threadfence();
        volatile uint32_t *hs = fb.host_status;
        hs[0] = atomicOr(status, 0u);
        hs[1] = atomicExch(status + 1, 0u);
        hs[2] = atomicExch(status + 2, 0u);
        status[3] = 0u;
      }
    }
  };
  // FLAT tiles (two thirds of the tiles of a typical frame: no binned record, only constant-depth covering records in
  // the big list = the background quad) were recognised by ruf_tile_info_kernel, which also took the virtual depth
  // through to_linear_depth and the threshold: what is left is a streaming pass (5 B/px for 16UC1) that never touches
  // shared memory, the record lists or a barrier.
  //
  // CL: there is no tile-info launch in front of this variant (on a launch of one frame a kernel boundary costs more
  // than the kernel saves): the cluster's CTAs do for their tile what a thread of ruf_tile_info_kernel does -- same
  // tests, same expressions, one lane per big-list record -- and rank 0 folds the overflow flags into the status words.
  __shared__ uint32_t s_ti[4];
  {
    uint4 ti;
    if (CL == 1) {
      ti = __ldg(tinfo + (size_t)frame * d.ntiles + tile);
    } else {
      const uint32_t *ctr0 = ctr_all + (size_t)frame * d.ctr_stride;
      const uint2 nfb0 = __ldg(reinterpret_cast<const uint2 *>(ctr0 + kCtrWords) + tile);
      const uint32_t nbig0 = min(__ldg(ctr0 + kCtrBig), d.cap_big);
      if (crank == 0u && tid == 0) {
        uint32_t flags = (nfb0.x + nfb0.y > d.cap_tile) ? kFlagBinOverflow : 0u;
        if (tile == 0) {
          flags |= __ldg(ctr0 + kCtrFlags);
          atomicAdd(status + 1, __ldg(ctr0 + kCtrKept));
        }
        if (flags) atomicOr(status, flags);
      }
      ti = make_uint4(0u, 0u, 0u, 0u);
      if (nfb0.x + nfb0.y == 0u && nbig0 <= (uint32_t)kFlatMaxBig) {       // CTA-uniform
        if (warp == 0) {
          uint32_t cls = 0;
          float zf = 1.0f;                                                 // glClear depth
          if ((uint32_t)lane < nbig0) {
            const TriRec r = load_rec_global(big_all + (size_t)frame * d.cap_big + lane);
            float zc = 0.0f;
            cls = classify_big_record(r, tile_x0, tile_y0, &zc);
            if (cls == 3u) zf = fminf(zf, zc);
          }
          const bool mixed = __any_sync(0xffffffffu, cls == 1u || cls == 2u);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) zf = fminf(zf, __shfl_xor_sync(0xffffffffu, zf, o));
          if (lane == 0) {
            uint32_t fl = 0, th = 0;
            if (!mixed) {
              fl = kTileFlat;
              const float zt = zf;
              if (zt == 1.0f) fl |= kTileUndrawn;
              else {
                const float thr = (sp.k1 / (zt - sp.k2)) - sp.max_diff;
                th = (ENC == 1) ? (uint32_t)u16_threshold(thr) : __float_as_uint(thr);
              }
            }
            s_ti[0] = fl; s_ti[1] = th; s_ti[2] = __float_as_uint(zf);
          }
        }
        __syncthreads();
        ti = make_uint4(s_ti[0], s_ti[1], s_ti[2], 0u);
      }
    }
    if (ti.x & kTileFlat) {
      const bool drawn = !(ti.x & kTileUndrawn);
      const uint32_t repl_u16 = f32_to_u16(sp.replace_value);
      const uint32_t repl2 = repl_u16 | (repl_u16 << 16);
      const float zt = __uint_as_float(ti.z);
      export_status();
      if (!frag) return;
#pragma unroll
      for (int half = 0; half < kRowsPerThread; ++half) {
        const int gy = tile_y0 + prow + 32 * half, gx = tile_x0 + pcol;
        if (gy >= d.H || gx >= d.W) continue;
        const size_t base = (size_t)frame * d.W * d.H + (size_t)gy * d.W + gx;
        if (fb.vec_ok && gx + 8 <= d.W) {
          uint2 mq = make_uint2(0u, 0u);
          if (ENC == 1) {
            uint4 o = make_uint4(0u, 0u, 0u, 0u);                       // never drawn: clear colour (:566)
            if (drawn) {
              const uint4 sens = __ldg(reinterpret_cast<const uint4 *>(static_cast<const uint16_t *>(fb.depth_in) + base));
              shade_row_u16_uniform(sens, (int)ti.y, repl2, o, mq);
            }
            *reinterpret_cast<uint4 *>(static_cast<uint16_t *>(fb.depth_out) + base) = o;
          } else {
            float od[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (drawn) {
              const float thr = __uint_as_float(ti.y);
              uint4 s0, s1;
              ldg_nc_256(static_cast<const float *>(fb.depth_in) + base, s0, s1);
              const float sensor[8] = {__uint_as_float(s0.x), __uint_as_float(s0.y), __uint_as_float(s0.z), __uint_as_float(s0.w),
                                       __uint_as_float(s1.x), __uint_as_float(s1.y), __uint_as_float(s1.z), __uint_as_float(s1.w)};
              uint32_t om[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float sv = sensor_daz(sensor[i]);
                const bool sflt = sv > thr;                             // frag:23
                od[i] = sflt ? mix_filtered(sv, sp.replace_value) : sv; // frag:29
                om[i] = sflt ? 255u : 0u;
              }
              mq.x = om[0] | (om[1] << 8) | (om[2] << 16) | (om[3] << 24);
              mq.y = om[4] | (om[5] << 8) | (om[6] << 16) | (om[7] << 24);
            }
            stg_256(static_cast<float *>(fb.depth_out) + base, od);
          }
          if (fb.mask_out) store_mask(fb, base, mq);
          if (fb.zbuf_out) {
            float4 *p = reinterpret_cast<float4 *>(fb.zbuf_out + base);
            p[0] = make_float4(zt, zt, zt, zt);
            p[1] = make_float4(zt, zt, zt, zt);
          }
        } else {
          const float ztmp = zt;
          shade_scalar<ENC>(fb, sp, base, min(8, d.W - gx), &ztmp, 0);
        }
      }
      return;
    }
  }
  const uint32_t *ctr = ctr_all + (size_t)frame * d.ctr_stride;
  // records binned to this tile (CTA-uniform): nf at the front of its list (triangles facing the camera, drawn
  // first), nb at the back (facing away: drawn last and depth-culled against what is already there)
  const uint2 nfb = __ldg(reinterpret_cast<const uint2 *>(ctr + kCtrWords) + tile);
  uint32_t nf = nfb.x, nb = nfb.y;
  nf = min(nf, d.cap_tile);                  // an overflow (the two runs met) was flagged by ruf_tile_info_kernel
  nb = min(nb, d.cap_tile - nf);
  const BinRec *list = bins_all + ((size_t)frame * d.ntiles + tile) * d.cap_tile;
  const BinRec *flist = list, *blist = list + (d.cap_tile - nb);
  const uint32_t nb_tile = nb;               // cluster-uniform
  const uint32_t cnt_any = nf + nb;          // records of the whole tile (CL: cluster-uniform, decides which barriers exist)
  if (CL > 1) {
    // this CTA's share of both runs, in whole batches of RB records
    const uint32_t fbt = (nf + RB - 1u) / RB, bbt = (nb + RB - 1u) / RB;
    const uint32_t f0 = min(nf, RB * (fbt * crank / CL)), f1 = min(nf, RB * (fbt * (crank + 1u) / CL));
    const uint32_t b0 = min(nb, RB * (bbt * crank / CL)), b1 = min(nb, RB * (bbt * (crank + 1u) / CL));
    flist += f0; blist += b0;
    nf = f1 - f0; nb = b1 - b0;
  }
  const uint32_t cnt = nf + nb;
  const int nchunks = (int)((cnt + kChunk - 1) / kChunk);
  // chunk c of the virtual list "front run, then back run": at most two bulk copies into one ring stage
  auto issue_chunk = [&](int c, int stage) {
    const uint32_t lo = (uint32_t)c * kChunk, hi = min(lo + (uint32_t)kChunk, cnt);
    mbar_arrive_expect_tx(&full_bar[stage], (hi - lo) * (uint32_t)sizeof(BinRec));
    const uint32_t fhi = min(hi, nf);
    if (lo < fhi) bulk_g2s(&sbuf[stage][0], flist + lo, (fhi - lo) * (uint32_t)sizeof(BinRec), &full_bar[stage]);
    const uint32_t blo = max(lo, nf);
    if (blo < hi)
      bulk_g2s(&sbuf[stage][blo - lo], blist + (blo - nf), (hi - blo) * (uint32_t)sizeof(BinRec), &full_bar[stage]);
  };
  if (cnt_any) {
    if (tid == 0) {
#pragma unroll
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&full_bar[s], 1);                        // the issuing thread's arrive.expect_tx
        mbar_init(&empty_bar[s], kChunk / RB);             // one arrive per batch of RB records
      }
      mbar_fence_init();
      s_next[0] = 0; s_next[1] = 0;
      s_nwide = 0; s_nstat = 0;
#pragma unroll
      for (int c = 0; c < kStages; ++c) {
        s_issued[c] = -1;
        if (c < nchunks) { issue_chunk(c, c); s_issued[c] = c; }
      }
    }
    // the z tile starts at the cleared depth (glClear, 1.0); the per-frame big list (background quad, walls,
    // clipped triangles) is merged in at the end on registers -- min is associative, the result is the same
#pragma unroll
    for (int half = 0; half < kRowsPerThread; ++half) {
      uint4 *zp = reinterpret_cast<uint4 *>(&sz[(prow + 32 * half) * kTileW + pcol]);
      zp[0] = make_uint4(0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u);
      zp[1] = make_uint4(0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u);
    }
    if (tid < kZPad) sz[kTilePix + tid] = 0u;              // the depth cull may read past the tile's last row
  }

  // ===== all warps =====
  // ---- per-frame big list, part 1: classify its first 256 records against this tile (one per thread; a
  // typical frame has just the two triangles of the background quad, so only warp 0 does any work here):
  // 0 = no sample of this tile can be covered, 1 = every sample is covered, 2 = mixed, 3 = covered with a
  // constant depth.  Edge values are linear over the tile's sample grid, so their min / max sit on its corners.
  const uint32_t nbig = min(ctr[kCtrBig], d.cap_big);
  const TriRec *big = big_all + (size_t)frame * d.cap_big;
  const int tpx = tile_x0 * kSubpix + kSubpixHalf, tpy = tile_y0 * kSubpix + kSubpixHalf;
  auto classify = [&](uint32_t b0) {
    uint32_t cls = 0;
    int ox0 = 0, oy0 = 0;
    float oz0 = 0.f, ogx = 0.f, ogy = 0.f;
    const bool occlude = nbig - b0 > (uint32_t)kOccludeMin;     // CTA-uniform; a typical frame holds just the background quad
    if (b0 + tid < nbig) {
      const TriRec r = load_rec_global(big + b0 + tid);
      float zc = 0.0f;
      cls = classify_big_record(r, tile_x0, tile_y0, &zc);
      if (cls == 3u) s_bigz[tid] = zc;
      if (cls && occlude) { ox0 = r.x0; oy0 = r.y0; oz0 = r.z0; ogx = r.gx; ogy = r.gy; }
    }
    if (occlude) cls = occlusion_filter(cls, ox0, oy0, oz0, ogx, ogy, tpx, tpy, tid, &s_zcut, &s_zown);
    s_bigcls[tid] = (uint8_t)cls;
  };
  // The cleared z tile and the ring's barriers are visible after this barrier.  The classification comes AFTER it: its
  // results are only read behind the end-of-raster barrier, so the other warps start on the tile's records while warp 0
  // still waits for its big-list records (two dependent global loads and ~150 instructions of 64-bit arithmetic).
  __syncthreads();
  TL(1);
  classify(0);
  if (!cnt_any) __syncthreads();   // no raster phase (and no end-of-raster barrier) on this path

  if (cnt_any) {
    {
      // binned triangles: a chunk's records are batches of 32; warps claim batches from a shared
      // counter (a warp that drew light triangles simply takes the next batch), and a stage goes
      // back to the producer once all its batches sit in registers.
      //
      // Two passes over the list.  Pass 0: the batches holding the front run (triangles facing the camera).
      // Pass 1: the back run (triangles facing away), depth-culled per record (result-neutral): after a CTA
      // barrier the maximum of every 4x4 block of the z tile is taken once; a record whose nearest bbox sample
      // is not in front of the block maxima under its bbox cannot change a pixel (z-tile values only ever
      // decrease) and is dropped before any unit is dealt.  On closed meshes that is nearly all of them.
      const uint32_t nbatches = (cnt + RB - 1u) / RB;
      BL_DECL;
      uint32_t nb_front = kDepthCull ? min(nbatches, (nf + RB - 1u) / RB) : nbatches;
      if (CL == 1) {
        if (nbatches - nb_front < RUF_MIN_BACK_BATCHES) nb_front = nbatches;   // a second pass (two barriers) must pay for itself
      } else if (nb_tile < 32u * RUF_MIN_BACK_BATCHES) {
        nb_front = nbatches;                   // CL: every CTA of the cluster takes the same decision (cluster barrier inside)
      }
      for (int pass = 0; pass < 2; ++pass) {
      const uint32_t b_lo = pass ? nb_front : 0u, b_hi = pass ? nbatches : nb_front;
      if (pass) {
        if (CL == 1 ? (b_lo >= b_hi) : (!kDepthCull || nb_tile < 32u * RUF_MIN_BACK_BATCHES)) break;
        __syncthreads();                      // every front record is in the z tile
        TL(2);
        if (CL == 1) {
          if (tid < (kTileH / 4) * 16) {
            const uint4 *row = reinterpret_cast<const uint4 *>(&sz[(tid >> 4) * 4 * kTileW + (tid & 15) * 4]);
            uint32_t m = 0;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              const uint4 q = row[rr * (kTileW / 4)];
              m = max(max(m, max(q.x, q.y)), max(q.z, q.w));
            }
            s_zblk[tid] = m;
          }
        } else {
          // CL: the block maxima of the MERGED front image.  This CTA merges its band of 256 / CL blocks (one thread per
          // block row: the CL partial rows through DSMEM, min per pixel, max over the row, max over the block's four
          // rows by shuffles); after a second cluster barrier every CTA collects all 256 maxima from their owners.
          cluster_sync_all();                 // every CTA's front records are in its z tile
          constexpr int kBand = 256 / CL;     // blocks per CTA
          uint32_t m = 0;
          if (tid < 4 * kBand) {
            const int b = (int)crank * kBand + (tid >> 2), rr = tid & 3;
            const uint32_t *src = &sz[((b >> 4) * 4 + rr) * kTileW + (b & 15) * 4];
            uint4 q = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
#pragma unroll
            for (uint32_t rk = 0; rk < (uint32_t)CL; ++rk) {
              const uint4 a = ld_cluster_v4(cluster_map(src, rk));
              q = make_uint4(min(q.x, a.x), min(q.y, a.y), min(q.z, a.z), min(q.w, a.w));
            }
            m = max(max(q.x, q.y), max(q.z, q.w));
          }
          m = max(m, __shfl_xor_sync(0xffffffffu, m, 1));
          m = max(m, __shfl_xor_sync(0xffffffffu, m, 2));
          if (tid < 4 * kBand && (tid & 3) == 0) s_zblk[(int)crank * kBand + (tid >> 2)] = m;
          cluster_sync_all();
          if (tid < 256) s_zblk_cl[tid] = ld_cluster_u32(cluster_map(&s_zblk[tid], (uint32_t)(tid / kBand)));
        }
        __syncthreads();
      }
      if (pass) TL(3);
      const uint32_t *zblk = CL > 1 ? s_zblk_cl : s_zblk;
      for (;;) {
        uint32_t bt = 0;
        BL(bl_t0);
        if (lane == 0) bt = b_lo + smem_add(smem_u32(&s_next[pass]), 1u);
        bt = __shfl_sync(0xffffffffu, bt, 0);
        if (bt >= b_hi) break;
        const int c = (int)(bt / (kChunk / RB));              // ring position of the chunk
        const int stage = c % kStages;
        // A parity wait can only tell the current phase from the one before it.  Warps whose batches were
        // cheap (depth-culled) can claim a batch of a chunk that has not even been issued yet, while the
        // stage's previous chunk is still landing: looking at the barrier then would mistake "one phase behind"
        // for "complete".  So first wait until the chunk's copies have been issued into the stage.
        while (*reinterpret_cast<volatile int *>(&s_issued[stage]) < c) {}
        mbar_wait(&full_bar[stage], ((uint32_t)c / kStages) & 1);
        BL(bl_t1);
        const uint32_t nrec = min(cnt - (uint32_t)c * kChunk, (uint32_t)kChunk);
        const uint32_t idx = (bt % (kChunk / RB)) * RB + (uint32_t)lane;
        // ---- phase 1: one lane per record: clip to the tile, derive the incremental edge setup ----
        TriRec r;
        int i0 = 0, i1 = -1, j0 = 0, j1 = -1;
        int kind = 0;       // 0 nothing, 1 dealt out as row-block units, 2 warp-wide 32-bit, 3 warp-wide 64-bit
        int sA0 = 0, sA1 = 0, sA2 = 0, sB0 = 0, sB1 = 0, sB2 = 0, r0 = 0, r1 = 0, r2 = 0;
        int ncb = 0, nunits = 0;
        int geo = 0, bx = 0, by = 0;   // tile-local bbox origin and extent (packed); sample (i0, j0) relative to vertex 0
        if ((RB == 32u || (uint32_t)lane < RB) && idx < nrec) {
          r = load_bin_smem(&sbuf[stage][idx]);
          // candidate samples (S6): the pixel bbox of the snapped vertices, clipped to this tile and to the viewport
          const int xmin = min(r.x0, min(r.x1, r.x2)), xmax = max(r.x0, max(r.x1, r.x2));
          const int ymin = min(r.y0, min(r.y1, r.y2)), ymax = max(r.y0, max(r.y1, r.y2));
          i0 = max((xmin - kSubpixHalf + (kSubpix - 1)) >> kSubpixBits, tile_x0);
          i1 = min((xmax - kSubpixHalf) >> kSubpixBits, min(tile_x0 + kTileW, d.W) - 1);
          j0 = max((ymin - kSubpixHalf + (kSubpix - 1)) >> kSubpixBits, tile_y0);
          j1 = min((ymax - kSubpixHalf) >> kSubpixBits, min(tile_y0 + kTileH, d.H) - 1);
          r.bx = (uint32_t)i0 | ((uint32_t)i1 << 16); r.by = (uint32_t)j0 | ((uint32_t)j1 << 16);
          const int ex = xmax - xmin, ey = ymax - ymin;
          // 32-bit unit arithmetic is exact when every edge-function factor is below 2^14; MP admits long thin slivers
          // too: |E_k| <= 2 ex ey + 256 (ex + ey) at every sample of the bbox, below 2^31 when ex * ey < 2^29
          const bool narrow = MP ? ((long long)ex * ey < (1LL << 29) && ex < (1 << 21) && ey < (1 << 21))
                                 : (ex < 16384 && ey < 16384);
          ncb = (i1 - i0 + kUW) / kUW;                       // unit columns
          nunits = ncb * ((j1 - j0 + kUH) / kUH);
          // MP: up to kMultiPassUnits units (slivers) stay on the unit path; larger, compact triangles are cheaper in the
          // cooperative 8x4-footprint walk
          kind = narrow ? ((nunits <= (MP ? kMultiPassUnits : kMaxUnits)) ? 1 : 2) : 3;
          if (nunits > kMaxUnits) atomicAdd(&s_nstat, 1u);   // statistics for the host's choice of the variant (none on C2)
          if (pass) {
            // depth cull: z is monotone in both sample coordinates (every step is a correctly rounded fma), so
            // its minimum over the bbox sits on a corner sample, evaluated with the rasteriser's own expressions
            const int cx0 = (i0 - tile_x0) >> 2, cx1 = (i1 - tile_x0) >> 2, cy0 = (j0 - tile_y0) >> 2, cy1 = (j1 - tile_y0) >> 2;
            if ((cx1 - cx0 + 1) * (cy1 - cy0 + 1) <= 16) {
              uint32_t zmaxb = 0;
              for (int yy = cy0; yy <= cy1; ++yy)
                for (int xx = cx0; xx <= cx1; ++xx) zmaxb = max(zmaxb, zblk[yy * 16 + xx]);
              const float fxa = (float)(i0 * kSubpix + kSubpixHalf - r.x0), fxb = (float)(i1 * kSubpix + kSubpixHalf - r.x0);
              const float rza = fmaf(r.gy, (float)(j0 * kSubpix + kSubpixHalf - r.y0), r.z0);
              const float rzb = fmaf(r.gy, (float)(j1 * kSubpix + kSubpixHalf - r.y0), r.z0);
              const uint32_t zaa = __float_as_uint(clamp_z(fmaf(r.gx, fxa, rza))), zba = __float_as_uint(clamp_z(fmaf(r.gx, fxb, rza)));
              const uint32_t zab = __float_as_uint(clamp_z(fmaf(r.gx, fxa, rzb))), zbb = __float_as_uint(clamp_z(fmaf(r.gx, fxb, rzb)));
              const uint32_t zmin = min(min(zaa, zba), min(zab, zbb));
              if (zmin >= zmaxb || zmin >= 0x3f800000u) { kind = 0; nunits = 0; }
            }
#ifdef RUF_CULL_STATS
            atomicAdd(const_cast<uint32_t *>(ctr) + 1, 1u | (kind == 0 ? 0x10000u : 0u));   // debug: tested | culled << 16
#endif
          }
          if (kind == 1) {
            const Edges e = make_edges(r);
            const int px0 = i0 * kSubpix + kSubpixHalf, py0 = j0 * kSubpix + kSubpixHalf;
            bx = px0 - r.x0; by = py0 - r.y0;
            r0 = e.A0 * bx + e.B0 * by + e.bias0;
            r1 = e.A1 * (px0 - r.x1) + e.B1 * (py0 - r.y1) + e.bias1;
            r2 = e.A2 * (px0 - r.x2) + e.B2 * (py0 - r.y2) + e.bias2;
            sA0 = e.A0 * kSubpix; sA1 = e.A1 * kSubpix; sA2 = e.A2 * kSubpix;
            sB0 = e.B0 * kSubpix; sB1 = e.B1 * kSubpix; sB2 = e.B2 * kSubpix;
            geo = (i0 - tile_x0) | ((j0 - tile_y0) << 6) | ((i1 - i0) << 12) | ((j1 - j0) << 18);
          } else {
            nunits = 0;
          }
        } else {
          r.x0 = r.y0 = r.x1 = r.y1 = r.x2 = r.y2 = 0; r.z0 = r.gx = r.gy = 0.f; r.bx = r.by = r.pad = 0;
        }
        __syncwarp();
        if (lane == 0) {
          // this batch sits in registers.  A short last chunk has fewer batches: whoever drew its
          // first batch also arrives for the missing ones so that the barrier phase always completes.
          const uint32_t in_chunk = (nrec + RB - 1u) / RB;
          const uint32_t extra = ((bt % (kChunk / RB)) == 0) ? (kChunk / RB - in_chunk) : 0u;
          mbar_arrive_n(&empty_bar[stage], 1u + extra);
          // whoever took the chunk's last batch refills the stage with chunk c + kStages once all batches of
          // chunk c sit in registers (the other warps are at most a few shared-memory loads away from that)
          const int cn = c + kStages;
          if ((bt % (kChunk / RB)) == (kChunk / RB) - 1 && cn < nchunks) {
            mbar_wait(&empty_bar[stage], ((uint32_t)c / kStages) & 1);
            issue_chunk(cn, stage);
            __threadfence_block();                          // the barrier's new phase is set up before the flag says so
            *reinterpret_cast<volatile int *>(&s_issued[stage]) = cn;
          }
        }

        // Wide records (more than kMaxUnits units, or 64-bit edge arithmetic: pieces of walls, doubled boxes) are parked
        // in shared memory and rasterised by the WHOLE CTA after the last batch; one warp walking 128 footprints while
        // seven wait at the barrier was the long pole of example.urdf (C1) and of the wall tiles of C3.  Only when the
        // parking lot is full does the claiming warp cover them itself, one triangle at a time.
        if (kind >= 2) {
          const uint32_t pos = atomicAdd(&s_nwide, 1u);
          if (pos < (uint32_t)kWideCap) {
            uint4 *dst = reinterpret_cast<uint4 *>(&s_wide[pos]);
            dst[0] = make_uint4((uint32_t)r.x0, (uint32_t)r.y0, (uint32_t)r.x1, (uint32_t)r.y1);
            dst[1] = make_uint4((uint32_t)r.x2, (uint32_t)r.y2, __float_as_uint(r.z0), __float_as_uint(r.gx));
            dst[2] = make_uint4(__float_as_uint(r.gy), r.bx, r.by, (uint32_t)kind);
            kind = 0;
          }
        }
        unsigned wide = __ballot_sync(0xffffffffu, kind >= 2);
        while (wide) {
          const int src = __ffs(wide) - 1;
          wide &= wide - 1;
          const TriRec q = shfl_rec(r, src);
          const int qi0 = __shfl_sync(0xffffffffu, i0, src), qi1 = __shfl_sync(0xffffffffu, i1, src);
          const int qj0 = __shfl_sync(0xffffffffu, j0, src), qj1 = __shfl_sync(0xffffffffu, j1, src);
          raster_warp<true>(q, qi0, qi1, qj0, qj1, tile_x0, tile_y0, sz, lane);     // 64-bit edges are exact for every record
        }

        // ---- phase 2: the units (kUW x kUH samples) of the 32 records are dealt out to the lanes
        // round by round, so every lane does the same amount of branch-free work.  Unit table in
        // shared memory: owner lane | unit row << 5 | unit column << 10.
        int udone = 0;                         // MP: units of this lane's record dealt out in earlier passes
        BL(bl_t2);
#ifdef RUF_X_TIMELINE
        bl_items = 0;
#endif
        for (;;) {
        const int nu = MP ? min(nunits - udone, kMaxUnits) : nunits;
        if (MP && !__any_sync(0xffffffffu, nu > 0)) break;
        int incl = nu;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int y = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += y;
        }
        const int excl = incl - nu;
        const int items = __shfl_sync(0xffffffffu, incl, 31);
#ifdef RUF_X_TIMELINE
        bl_items += items;
#endif
        uint16_t *utab = s_units[warp];
        {
          int row = 0, cb = 0;
          if (udone) { row = udone / ncb; cb = udone - row * ncb; }
          for (int q = 0; q < nu; ++q) {
            utab[excl + q] = (uint16_t)(lane | (row << 5) | (cb << 10));
            if (++cb == ncb) { cb = 0; ++row; }
          }
        }
        __syncwarp();
        for (int base = 0; base < items; base += 32) {
          const int x = base + lane;
          const bool act = x < items;
          const uint32_t ent = act ? (uint32_t)utab[x] : 0u;
          const int o = (int)(ent & 31u), urow = (int)((ent >> 5) & 31u), ucol = (int)(ent >> 10);
          // fetch the owner's setup (all lanes shuffle; inactive lanes read lane 0 and discard)
          const int qA0 = __shfl_sync(0xffffffffu, sA0, o), qA1 = __shfl_sync(0xffffffffu, sA1, o),
                    qA2 = __shfl_sync(0xffffffffu, sA2, o);
          const int qB0 = __shfl_sync(0xffffffffu, sB0, o), qB1 = __shfl_sync(0xffffffffu, sB1, o),
                    qB2 = __shfl_sync(0xffffffffu, sB2, o);
          int e0 = __shfl_sync(0xffffffffu, r0, o), e1 = __shfl_sync(0xffffffffu, r1, o),
              e2 = __shfl_sync(0xffffffffu, r2, o);
          const int qgeo = __shfl_sync(0xffffffffu, geo, o);
          const int qbx = __shfl_sync(0xffffffffu, bx, o), qby = __shfl_sync(0xffffffffu, by, o);
          const float qz0 = __shfl_sync(0xffffffffu, r.z0, o), qgx = __shfl_sync(0xffffffffu, r.gx, o),
                      qgy = __shfl_sync(0xffffffffu, r.gy, o);
          if (act) {
            const int dx = ucol * kUW, dy = urow * kUH;          // unit origin relative to the bbox origin
            e0 += dy * qB0 + dx * qA0;
            e1 += dy * qB1 + dx * qA1;
            e2 += dy * qB2 + dx * qA2;
            uint32_t m = 0;
#pragma unroll
            for (int rr = 0; rr < kUH; ++rr) {
              int f0 = e0, f1 = e1, f2 = e2;
#pragma unroll
              for (int k = 0; k < kUW; ++k) {
                m = __funnelshift_l((uint32_t)(f0 | f1 | f2), m, 1);   // shifts in 1 for "outside"
                f0 += qA0; f1 += qA1; f2 += qA2;
              }
              e0 += qB0; e1 += qB1; e2 += qB2;
            }
            // bit (kUW*kUH - 1 - (rr*kUW + k)) of ~m <=> sample (dx + k, dy + rr) is covered; samples beyond
            // the bbox (clipped to the tile) belong to another tile or cannot be covered
            const int wrem = ((qgeo >> 12) & 63) + 1 - dx, hrem = ((qgeo >> 18) & 63) + 1 - dy;
            constexpr uint32_t kRowMask = (1u << kUW) - 1u, kAll = (kUH == 1) ? kRowMask : ((1u << (kUW * kUH)) - 1u);
            uint32_t vm = (kRowMask << kUW >> min(kUW, wrem)) & kRowMask;      // valid columns, MSB = k = 0
            if (kUH == 2) vm = (vm << kUW) | (hrem > 1 ? vm : 0u);
            m = (~m) & kAll & vm;
            if (m) {
              const float fx0 = (float)(qbx + dx * kSubpix);                  // exact: |value| < 2^24
              // (plain atomicMin on the static __shared__ array: direct shared addressing, immediate offsets)
              uint32_t *zp = sz + ((((qgeo >> 6) & 63) + dy) * kTileW + (qgeo & 63) + dx);
#pragma unroll
              for (int rr = 0; rr < kUH; ++rr) {
                const float rowz = fmaf(qgy, (float)(qby + (dy + rr) * kSubpix), qz0);
#pragma unroll
                for (int k = 0; k < kUW; ++k) {
                  // float(px - x0) for sample k = fx0 + 256 k (integers below 2^24: the sum is exact)
                  const uint32_t z = __float_as_uint(clamp_z(fmaf(qgx, fx0 + (float)(k * kSubpix), rowz)));
                  if ((m & (1u << (kUW * kUH - 1 - (rr * kUW + k)))) && z < 0x3f800000u)
                    atomicMin(zp + (rr * kTileW + k), z);
                }
              }
            }
          }
        }
        __syncwarp();                          // the unit table is rewritten by the next pass / batch
        if (!MP) break;
        udone += nu;
        }
        BL_END(pass);
      }
      }
    }
    consumer_bar_sync();                      // every record of the tile has been rasterised or parked
    if (tid == 0 && s_nstat) atomicAdd(status + 2, s_nstat);
    const uint32_t nwide = min(s_nwide, (uint32_t)kWideCap);      // CTA-uniform
    if (nwide) {
      for (uint32_t w = 0; w < nwide; ++w) {
        const TriRec q = load_rec_smem(&s_wide[w]);
        const int qi0 = max((int)(q.bx & 0xffffu), tile_x0), qi1 = min((int)(q.bx >> 16), tile_x0 + kTileW - 1);
        const int qj0 = max((int)(q.by & 0xffffu), tile_y0), qj1 = min((int)(q.by >> 16), tile_y0 + kTileH - 1);
        // warp w takes the footprint rows 4 w, 4 w + 32, ... of the clipped bbox
        raster_warp<true>(q, qi0, qi1, qj0, qj1, tile_x0, tile_y0, sz, lane, 4 * warp, 4 * (kRasterThreads / 32));
      }
      __syncthreads();
    }
    TL(4);
    if (CL > 1) cluster_sync_all();           // every CTA of the cluster has finished its share of the tile's records
    TL(5);
  }

  // Every thread owns 8 consecutive pixels of kRowsPerThread tile rows (32 rows apart): fetch the rasterised
  // depths, merge the per-frame big list on registers (one walk over the classified records for all rows),
  // then run the fragment stage row by row.
  float zall[kRowsPerThread][8];
#pragma unroll
  for (int half = 0; half < kRowsPerThread; ++half) {
    if (cnt_any) {
      uint4 zq0, zq1;
      if (CL == 1) {
        zq0 = *reinterpret_cast<const uint4 *>(&sz[(prow + 32 * half) * kTileW + pcol]);
        zq1 = *reinterpret_cast<const uint4 *>(&sz[(prow + 32 * half) * kTileW + pcol + 4]);
      } else {
        zq0 = zq1 = make_uint4(0x3f800000u, 0x3f800000u, 0x3f800000u, 0x3f800000u);
        if (frag) {
#pragma unroll
          for (uint32_t rk = 0; rk < (uint32_t)CL; ++rk) {
            const uint32_t a = cluster_map(&sz[(prow + 32 * half) * kTileW + pcol], rk);
            const uint4 a0 = ld_cluster_v4(a), a1 = ld_cluster_v4(a + 16u);
            zq0 = make_uint4(min(zq0.x, a0.x), min(zq0.y, a0.y), min(zq0.z, a0.z), min(zq0.w, a0.w));
            zq1 = make_uint4(min(zq1.x, a1.x), min(zq1.y, a1.y), min(zq1.z, a1.z), min(zq1.w, a1.w));
          }
        }
      }
      zall[half][0] = __uint_as_float(zq0.x); zall[half][1] = __uint_as_float(zq0.y); zall[half][2] = __uint_as_float(zq0.z);
      zall[half][3] = __uint_as_float(zq0.w); zall[half][4] = __uint_as_float(zq1.x); zall[half][5] = __uint_as_float(zq1.y);
      zall[half][6] = __uint_as_float(zq1.z); zall[half][7] = __uint_as_float(zq1.w);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) zall[half][i] = 1.0f;     // glClear depth
    }
  }
  // ---- per-frame big list, part 2: pixel-parallel on registers ----
  {
    const int px0 = (tile_x0 + pcol) * kSubpix + kSubpixHalf;
    const int py0 = (tile_y0 + prow) * kSubpix + kSubpixHalf;
    for (uint32_t b0 = 0; b0 < nbig; b0 += kRasterThreads) {
      if (b0) {                                   // more than 256 records (clipping-heavy views): classify the next round
        __syncthreads();
        classify(b0);
        __syncthreads();
      }
      const uint32_t nl = min(nbig - b0, (uint32_t)kRasterThreads);
      if (CL > 1 && !frag) continue;
      for (uint32_t b = 0; b < nl; ++b) {
        const uint32_t c = s_bigcls[b];
        if (c == 0) continue;
        if (c == 3) {
          const float z = s_bigz[b];
          if (z < 1.0f) {
#pragma unroll
            for (int half = 0; half < kRowsPerThread; ++half)
#pragma unroll
              for (int i = 0; i < 8; ++i) zall[half][i] = fminf(zall[half][i], z);
          }
          continue;
        }
        const TriRec r = load_rec_global(big + b0 + b);
        if (c == 1) {
#pragma unroll
          for (int half = 0; half < kRowsPerThread; ++half) {
            const float rowz = fmaf(r.gy, (float)(py0 + 32 * half * kSubpix - r.y0), r.z0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float z = clamp_z(fmaf(r.gx, (float)(px0 + i * kSubpix - r.x0), rowz));
              if (z < 1.0f) zall[half][i] = fminf(zall[half][i], z);
            }
          }
          continue;
        }
        const Edges e = make_edges(r);
        const long long s0 = (long long)e.A0 * kSubpix, s1 = (long long)e.A1 * kSubpix, s2 = (long long)e.A2 * kSubpix;
#pragma unroll
        for (int half = 0; half < kRowsPerThread; ++half) {
          const int py = py0 + 32 * half * kSubpix;
          long long e0 = (long long)e.A0 * (px0 - r.x0) + (long long)e.B0 * (py - r.y0) + e.bias0;
          long long e1 = (long long)e.A1 * (px0 - r.x1) + (long long)e.B1 * (py - r.y1) + e.bias1;
          long long e2 = (long long)e.A2 * (px0 - r.x2) + (long long)e.B2 * (py - r.y2) + e.bias2;
          const float rowz = fmaf(r.gy, (float)(py - r.y0), r.z0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if ((e0 | e1 | e2) >= 0) {
              float z = clamp_z(fmaf(r.gx, (float)(px0 + i * kSubpix - r.x0), rowz));
              if (z < 1.0f) zall[half][i] = fminf(zall[half][i], z);
            }
            e0 += s0; e1 += s1; e2 += s2;
          }
        }
      }
    }
  }

  TL(6);
  // ---- fused fragment stage (include/shaders/urdf_filter.frag:19-35): 8 pixels per thread and row, vector loads/stores ----
  const uint32_t repl_u16 = f32_to_u16(sp.replace_value);     // convertTo(CV_16U, 1000) of the replaced pixels, :311
  const float kInf = __int_as_float(0x7f800000);
#pragma unroll
  for (int half = 0; half < kRowsPerThread; ++half) {
    const int trow = prow + 32 * half;
    const int gy = tile_y0 + trow, gx = tile_x0 + pcol;
    if (gy >= d.H || gx >= d.W || !frag) continue;
    const size_t base = (size_t)frame * d.W * d.H + (size_t)gy * d.W + gx;
    const float (&zw)[8] = zall[half];
    if (fb.vec_ok && gx + 8 <= d.W) {
      // to_linear_depth (frag:14-17,22) and the threshold (frag:23) once per distinct z of the run.  Never-drawn
      // pixels (z = 1: clear colour, :566) get +inf, so that `sensor > thr` is false for them.
      float thr[8];
      thr[0] = (zw[0] == 1.0f) ? kInf : (sp.k1 / (zw[0] - sp.k2)) - sp.max_diff;
      {
        float zprev = zw[0], tprev = thr[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) {
          if (zw[i] != zprev) {
            tprev = (zw[i] == 1.0f) ? kInf : (sp.k1 / (zw[i] - sp.k2)) - sp.max_diff;
            zprev = zw[i];
          }
          thr[i] = tprev;
        }
      }
      uint2 mq;
      if (ENC == 1) {
        const uint4 sens = __ldg(reinterpret_cast<const uint4 *>(static_cast<const uint16_t *>(fb.depth_in) + base));
        uint4 o;
        {
          const uint32_t w[4] = {sens.x, sens.y, sens.z, sens.w};
          uint32_t u[8], om[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t raw = (i & 1) ? (w[i >> 1] >> 16) : (w[i >> 1] & 0xffffu);
            const float sensor = (float)raw * 0.001f;               // convertTo(CV_32F, 0.001), :288
            const bool drawn = zw[i] != 1.0f;                       // else clear colour: depth 0, mask 0 (:566)
            const bool sflt = sensor > thr[i];                      // frag:23
            // an unfiltered pixel is sat_u16(rint((u * 0.001f) * 1000.f)), which is u itself for every
            // 16-bit u (exhaustively checked: tests/test_oracle_encodings.py::test_u16_roundtrip_identity_all_65536)
            u[i] = drawn ? (sflt ? repl_u16 : raw) : 0u;
            om[i] = sflt ? 255u : 0u;
          }
          o.x = u[0] | (u[1] << 16); o.y = u[2] | (u[3] << 16); o.z = u[4] | (u[5] << 16); o.w = u[6] | (u[7] << 16);
          mq.x = om[0] | (om[1] << 8) | (om[2] << 16) | (om[3] << 24);
          mq.y = om[4] | (om[5] << 8) | (om[6] << 16) | (om[7] << 24);
        }
        *reinterpret_cast<uint4 *>(static_cast<uint16_t *>(fb.depth_out) + base) = o;
      } else {
        uint4 sens0, sens1;
        ldg_nc_256(static_cast<const float *>(fb.depth_in) + base, sens0, sens1);
        const float sensor[8] = {__uint_as_float(sens0.x), __uint_as_float(sens0.y), __uint_as_float(sens0.z),
                                 __uint_as_float(sens0.w), __uint_as_float(sens1.x), __uint_as_float(sens1.y),
                                 __uint_as_float(sens1.z), __uint_as_float(sens1.w)};
        float od[8];
        uint32_t om[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool drawn = zw[i] != 1.0f;
          const float sv = sensor_daz(sensor[i]);
          const bool sflt = sv > thr[i];
          od[i] = drawn ? (sflt ? mix_filtered(sv, sp.replace_value) : sv) : 0.0f;    // frag:29, mix() with a in {0,1}
          om[i] = sflt ? 255u : 0u;
        }
        stg_256(static_cast<float *>(fb.depth_out) + base, od);
        mq.x = om[0] | (om[1] << 8) | (om[2] << 16) | (om[3] << 24);
        mq.y = om[4] | (om[5] << 8) | (om[6] << 16) | (om[7] << 24);
      }
      if (fb.mask_out) store_mask(fb, base, mq);
      if (fb.zbuf_out) {
        float4 *p = reinterpret_cast<float4 *>(fb.zbuf_out + base);
        p[0] = make_float4(zw[0], zw[1], zw[2], zw[3]);
        p[1] = make_float4(zw[4], zw[5], zw[6], zw[7]);
      }
    } else {
      float ztmp[8];                 // a copy for the out-of-line call: the register array itself must not become addressable
#pragma unroll
      for (int i = 0; i < 8; ++i) ztmp[i] = zw[i];
      shade_scalar<ENC>(fb, sp, base, min(8, d.W - gx), ztmp, 1);
    }
  }
  if (CL > 1 && cnt_any) cluster_sync_all();   // no CTA leaves while another one may still read its z tile
  export_status();
  TL(7);
}

// ------------------------------------------------------------------------------------------
// host-side launcher for one batch
// ------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------
// Forward kinematics on the device (replaces the per-frame tf lookups of the reference's host:
// src/urdf_renderer.cpp:173-190 for the links, src/urdf_filter.cpp:522,602-614 for the camera).
// Double precision, no fma, the same operation order as the CPU oracle (orc_fk*): bit-exact.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void d_mat4_mul(const double *A, const double *B, double *C)
{
  double T[16];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 4; ++k) s += A[k * 4 + r] * B[c * 4 + k];
      T[c * 4 + r] = s;
    }
#pragma unroll
  for (int i = 0; i < 16; ++i) C[i] = T[i];
}

// sin / cos with a fixed operation order (Cody-Waite by pi/2 + Taylor in Horner form): same bits as the oracle
__device__ __forceinline__ void d_sincos(double x, double &sn, double &cs)
{
  const double two_over_pi = 0.63661977236758134308;
  const double p1 = 1.57079632673412561417e+00, p2 = 6.07710050650619224932e-11, p3 = 2.02226624879595063154e-21;
  double kf = x * two_over_pi;
  kf = (kf >= 0.0) ? floor(kf + 0.5) : -floor(-kf + 0.5);
  double r = x - kf * p1;
  r = r - kf * p2;
  r = r - kf * p3;
  const double z = r * r;
  double ps = 2.81145725434552076320e-15;
  ps = ps * z + -7.64716373181981647590e-13;
  ps = ps * z + 1.60590438368216145994e-10;
  ps = ps * z + -2.50521083854417187751e-08;
  ps = ps * z + 2.75573192239858906526e-06;
  ps = ps * z + -1.98412698412698412698e-04;
  ps = ps * z + 8.33333333333333333333e-03;
  ps = ps * z + -1.66666666666666666667e-01;
  const double sr = r + (r * z) * ps;
  double pc = 4.77947733238738529744e-14;
  pc = pc * z + -1.14707455977297247139e-11;
  pc = pc * z + 2.08767569878680989792e-09;
  pc = pc * z + -2.75573192239858906526e-07;
  pc = pc * z + 2.48015873015873015873e-05;
  pc = pc * z + -1.38888888888888888889e-03;
  pc = pc * z + 4.16666666666666666667e-02;
  pc = pc * z + -0.5;
  const double cr = 1.0 + z * pc;
  const long long k = (long long)kf;
  switch ((int)(k & 3)) {
    case 0: sn = sr; cs = cr; break;
    case 1: sn = cr; cs = -sr; break;
    case 2: sn = -sr; cs = -cr; break;
    default: sn = -cr; cs = sr; break;
  }
}

__device__ __forceinline__ void d_joint_motion(int type, const double *axis, double q, double *M)
{
#pragma unroll
  for (int i = 0; i < 16; ++i) M[i] = (i % 5 == 0) ? 1.0 : 0.0;
  if (type == 1) {
    double sn, cs;
    d_sincos(q * 0.5, sn, cs);
    const double x = axis[0] * sn, y = axis[1] * sn, z = axis[2] * sn, w = cs;
    const double d = x * x + y * y + z * z + w * w;
    const double s = 2.0 / d;
    const double xs = x * s, ys = y * s, zs = z * s;
    const double wx = w * xs, wy = w * ys, wz = w * zs;
    const double xx = x * xs, xy = x * ys, xz = x * zs;
    const double yy = y * ys, yz = y * zs, zz = z * zs;
    M[0] = 1.0 - (yy + zz); M[4] = xy - wz;         M[8] = xz + wy;
    M[1] = xy + wz;         M[5] = 1.0 - (xx + zz); M[9] = yz - wx;
    M[2] = xz - wy;         M[6] = yz + wx;         M[10] = 1.0 - (xx + yy);
  } else if (type == 2) {
    M[12] = axis[0] * q; M[13] = axis[1] * q; M[14] = axis[2] * q;
  }
}

// one thread per (frame, link): the product along the link's ancestor chain, root first
__global__ void __launch_bounds__(128) ruf_fk_links_kernel(Kinematics k, int n_frames, const double *__restrict__ joint_q,
                                                         double *__restrict__ links)
{
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)n_frames * k.n_links) return;
  const int l = (int)(gid % k.n_links);
  const long long f = gid / k.n_links;
  const double *q = joint_q + f * k.n_links;
  double T[16], M[16], O[16];
  const int c0 = k.chain_off[l], c1 = k.chain_off[l + 1];
  for (int c = c0; c < c1; ++c) {
    const int a = k.chain_idx[c];
#pragma unroll
    for (int i = 0; i < 16; ++i) O[i] = k.origin[16 * (size_t)a + i];
    if (c == c0) {
#pragma unroll
      for (int i = 0; i < 16; ++i) T[i] = O[i];
    } else {
      d_mat4_mul(T, O, T);
    }
    const int type = k.type[a];
    if (type != 0) {
      d_joint_motion(type, k.axis + 3 * (size_t)a, q[a], M);
      d_mat4_mul(T, M, T);
    }
  }
  double *out = links + 16 * (size_t)gid;
#pragma unroll
  for (int i = 0; i < 16; ++i) out[i] = T[i];
}

// one thread per (frame, part) plus one per frame for the view matrix
__global__ void __launch_bounds__(128) ruf_fk_outputs_kernel(Kinematics k, int n_frames, const double *__restrict__ links,
                                                           double tx, double ty, double *__restrict__ part_model,
                                                           double *__restrict__ view)
{
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = k.n_parts + 1;
  if (gid >= (long long)n_frames * per) return;
  const int p = (int)(gid % per);
  const long long f = gid / per;
  const double *L = links + 16 * (size_t)f * k.n_links;
  double A[16], B[16], C[16];
  if (p < k.n_parts) {
#pragma unroll
    for (int i = 0; i < 16; ++i) { A[i] = L[16 * (size_t)k.part_link[p] + i]; B[i] = k.part_local[16 * (size_t)p + i]; }
    d_mat4_mul(A, B, C);
    double *out = part_model + 16 * ((size_t)f * k.n_parts + p);
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] = C[i];
    return;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) B[i] = k.cam_mount[i];
  if (k.cam_link < 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) C[i] = B[i];
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) A[i] = L[16 * (size_t)k.cam_link + i];
    d_mat4_mul(A, B, C);
  }
  double I[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) I[i] = (i % 5 == 0) ? 1.0 : 0.0;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) I[c * 4 + r] = C[r * 4 + c];
#pragma unroll
  for (int r = 0; r < 3; ++r) I[12 + r] = I[0 * 4 + r] * -C[12] + I[1 * 4 + r] * -C[13] + I[2 * 4 + r] * -C[14];
#pragma unroll
  for (int r = 0; r < 3; ++r) I[12 + r] = I[12 + r] + I[0 * 4 + r] * tx;
#pragma unroll
  for (int r = 0; r < 3; ++r) I[12 + r] = I[12 + r] + I[1 * 4 + r] * ty;
#pragma unroll
  for (int i = 0; i < 16; ++i) A[i] = k.view_pre[i];
  d_mat4_mul(A, I, C);
  double *out = view + 16 * (size_t)f;
#pragma unroll
  for (int i = 0; i < 16; ++i) out[i] = C[i];
}

cudaError_t launch_fk(const Kinematics &k, int n_frames, const double *d_joint_q, double tx, double ty,
                      double *d_links, double *d_part_model, double *d_view, cudaStream_t s)
{
  const long long n1 = (long long)n_frames * k.n_links;
  if (n1 > 0) ruf_fk_links_kernel<<<(unsigned)((n1 + 127) / 128), 128, 0, s>>>(k, n_frames, d_joint_q, d_links);
  const long long n2 = (long long)n_frames * (k.n_parts + 1);
  ruf_fk_outputs_kernel<<<(unsigned)((n2 + 127) / 128), 128, 0, s>>>(k, n_frames, d_links, tx, ty, d_part_model, d_view);
  return cudaGetLastError();
}

// the raster kernel's instantiations: [cluster split][multi-pass][encoding]
typedef void (*RasterFn)(Dims, const TriRec *, const BinRec *, const uint32_t *, const uint4 *, ShaderParams, FrameBuffers, uint32_t *);
static RasterFn raster_fn(int enc, bool mp, bool cl)
{
  static const RasterFn fns[2][2][2] = {
      {{ruf_raster_filter_kernel<0, false, 1>, ruf_raster_filter_kernel<1, false, 1>},
       {ruf_raster_filter_kernel<0, true, 1>, ruf_raster_filter_kernel<1, true, 1>}},
      {{ruf_raster_filter_kernel<0, false, kClusterSplit>, ruf_raster_filter_kernel<1, false, kClusterSplit>},
       {ruf_raster_filter_kernel<0, true, kClusterSplit>, ruf_raster_filter_kernel<1, true, kClusterSplit>}}};
  return fns[cl ? 1 : 0][mp ? 1 : 0][enc == 1 ? 1 : 0];
}

#ifdef RUF_X_TIMELINE
extern "C" __attribute__((visibility("default"))) int ruf_debug_timeline(unsigned long long *out, int n_words)
{
  cudaMemcpyFromSymbol(out, g_timeline, (size_t)n_words * 8);
  static unsigned long long zero[8192 * 8];
  cudaMemcpyToSymbol(g_timeline, zero, sizeof(zero));
  return (int)cudaGetLastError();
}
extern "C" __attribute__((visibility("default"))) int ruf_debug_timeline1(unsigned long long *out)
{
  cudaMemcpyFromSymbol(out, g_timeline1, sizeof(unsigned long long) * 1024 * 64);
  void *p = nullptr;
  cudaGetSymbolAddress(&p, g_timeline1);
  cudaMemset(p, 0, sizeof(unsigned long long) * 1024 * 64);
  return (int)cudaGetLastError();
}
extern "C" __attribute__((visibility("default"))) int ruf_debug_batchlog(unsigned long long *out)
{
  cudaMemcpyFromSymbol(out, g_batchlog, sizeof(unsigned long long) * 1024 * 8 * 16 * 4);
  cudaMemset(nullptr, 0, 0);
  void *p = nullptr;
  cudaGetSymbolAddress(&p, g_batchlog);
  cudaMemset(p, 0, sizeof(unsigned long long) * 1024 * 8 * 16 * 4);
  return (int)cudaGetLastError();
}
#endif

bool is_raster_kernel(const void *func)
{
  for (int i = 0; i < 8; ++i)
    if (func == (const void *)raster_fn(i & 1, (i & 2) != 0, (i & 4) != 0)) return true;
  return false;
}

cudaError_t check_kernel_image()
{
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, (const void *)raster_fn(1, false, false));
  if (e != cudaSuccess) return e;
  // opt in to > 48 KB of dynamic shared memory (per device: call once per context)
  for (int i = 0; i < 8; ++i)
    if ((e = cudaFuncSetAttribute((const void *)raster_fn(i & 1, (i & 2) != 0, (i & 4) != 0), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)kRasterDynSmem)) != cudaSuccess)
      return e;
  return cudaSuccess;
}

cudaError_t launch_frames(const Dims &d, const Model &m, const Workspace &ws, int n_frames,
                          const double *d_proj, const double *d_view, const double *d_part_model,
                          const double *d_lookat, int enc, const ShaderParams &sp,
                          const FrameBuffers &fb, cudaStream_t s, int *n_launches, cudaEvent_t *ev,
                          cudaEvent_t depth_ready)
{
  cudaError_t err;
  int launches = 0;
  // RUF_DEBUG_SYNC=1: synchronise after every kernel and say which one failed (debugging aid)
  static const bool debug_sync = getenv("RUF_DEBUG_SYNC") != nullptr;
  auto check = [&](const char *what) -> cudaError_t {
    if (!debug_sync) return cudaSuccess;
    cudaError_t e = cudaStreamSynchronize(s);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) fprintf(stderr, "[ruf] %s failed: %s\n", what, cudaGetErrorString(e));
    return e;
  };
  const long long n_ctr = (long long)n_frames * d.ctr_stride;
  if (!d.fold_clear) {
    err = cudaMemsetAsync(ws.ctr, 0, (size_t)n_ctr * sizeof(uint32_t), s);
    if (err != cudaSuccess) return err;
  }
  if (ev) cudaEventRecord(ev[0], s);
  {
    long long total = (long long)n_frames * (d.n_parts + 1) * 16;
    unsigned blocks = (unsigned)((total + 255) / 256);
    const bool seed = d.fold_clear && d.bg_mode == kBgSkip && ws.bg_seed;
    ruf_pose_kernel<<<blocks, 256, 0, s>>>(d_proj, d_view, d_part_model, d_lookat, m.part_aabb, d.n_parts, n_frames,
                                           ws.mvp, ws.vis, ws.ctr, d.fold_clear ? n_ctr : 0LL, seed ? ws.bg_seed : nullptr,
                                           reinterpret_cast<uint32_t *>(ws.big), d.ctr_stride, d.cap_big);
    ++launches;
    if ((err = check("ruf_pose_kernel")) != cudaSuccess) return err;
    if (ev) cudaEventRecord(ev[1], s);
  }
  {
    // a CTA keeps its meshlet in registers over a run of frames; short batches still fill the machine
    // (about a dozen waves of 148 SMs x 6 CTAs keep the tail of the launch short: measured)
    int fpc = kSetupFrames;
    while (fpc > 1 && (long long)d.n_meshlets * ((n_frames + fpc - 1) / fpc) < 12000) fpc >>= 1;
    if (d.force_fpc > 0) fpc = d.force_fpc;
    // the background quad is the model's last meshlet
    const bool seeded = d.fold_clear && d.bg_mode == kBgSkip && ws.bg_seed;
    Model mk = m;
    int nm = d.n_meshlets;
    if (d.bg_mode == kBgOnly) { mk.meshlets += nm - 1; nm = 1; }
    else if (seeded) nm -= 1;
    if (nm > 0) {
      dim3 grid((unsigned)nm, (unsigned)((n_frames + fpc - 1) / fpc));
      ruf_setup_bin_kernel<<<grid, kSetupThreads, 0, s>>>(mk, ws.mvp, ws.vis, d, n_frames, fpc, ws.big, ws.bins, ws.ctr);
      ++launches;
      if ((err = check("ruf_setup_bin_kernel")) != cudaSuccess) return err;
    }
    if (d.bg_mode == kBgOnly) {
      if (n_launches) *n_launches = launches;
      return cudaGetLastError();
    }
    if (!d.cluster_split) {        // (the cluster-split raster variant does this for itself)
      const long long n_items = (long long)n_frames * d.ntiles;
      const unsigned tblocks = (unsigned)((n_items + 127) / 128);
      if (enc == 1) ruf_tile_info_kernel<1><<<tblocks, 128, 0, s>>>(d, n_frames, ws.big, ws.ctr, ws.status, sp, ws.tinfo);
      else ruf_tile_info_kernel<0><<<tblocks, 128, 0, s>>>(d, n_frames, ws.big, ws.ctr, ws.status, sp, ws.tinfo);
      ++launches;
      if ((err = check("ruf_tile_info_kernel")) != cudaSuccess) return err;
    }
    if (ev) cudaEventRecord(ev[2], s);
  }
  {
    // the depth image is first touched here: its upload may still be running under the pose / setup kernels
    if (depth_ready && (err = cudaStreamWaitEvent(s, depth_ready, 0)) != cudaSuccess) return err;
    const bool cl = d.cluster_split != 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)d.tiles_x * (cl ? kClusterSplit : 1), (unsigned)d.tiles_y, (unsigned)n_frames);
    cfg.blockDim = dim3(kRasterThreads);
    cfg.dynamicSmemBytes = kRasterDynSmem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kClusterSplit; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cl ? 1 : 0;
    if ((err = cudaLaunchKernelEx(&cfg, raster_fn(enc, d.multipass != 0, cl), d, (const TriRec *)ws.big, (const BinRec *)ws.bins,
                                  (const uint32_t *)ws.ctr, (const uint4 *)ws.tinfo, sp, fb, ws.status)) != cudaSuccess)
      return err;
    ++launches;
    if ((err = check("ruf_raster_filter_kernel")) != cudaSuccess) return err;
    if (ev) cudaEventRecord(ev[3], s);
  }
  if (n_launches) *n_launches = launches;
  return cudaGetLastError();
}

}  // namespace ruf
