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
// model packing: T*9 floats + T part ids -> three float4 streams (coalesced 16-byte loads)
// ------------------------------------------------------------------------------------------
__global__ void ruf_pack_model_kernel(const float *__restrict__ xyz, const uint32_t *__restrict__ part,
                                      long long n, float4 *v0, float4 *v1, float4 *v2)
{
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const float *p = xyz + 9 * t;
  v0[t] = make_float4(p[0], p[1], p[2], __uint_as_float(part[t]));
  v1[t] = make_float4(p[3], p[4], p[5], 0.f);
  v2[t] = make_float4(p[6], p[7], p[8], 0.f);
}

cudaError_t launch_pack_model(const float *d_tri_xyz, const uint32_t *d_tri_part, long long n_tris,
                              float4 *v0, float4 *v1, float4 *v2, cudaStream_t s)
{
  if (n_tris <= 0) return cudaSuccess;
  unsigned blocks = (unsigned)((n_tris + 255) / 256);
  ruf_pack_model_kernel<<<blocks, 256, 0, s>>>(d_tri_xyz, d_tri_part, n_tris, v0, v1, v2);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K0: MVP table.  gl_ModelViewProjectionMatrix of every drawn part, composed in double in the
// order the GL matrix stack does (PROJECTION * MODELVIEW, MODELVIEW = view * part_model) and
// rounded once to float.  Row n_parts is the background quad: P * LookAt
// (src/urdf_filter.cpp:576-596).  One thread per matrix element.
// ------------------------------------------------------------------------------------------
__global__ void ruf_pose_kernel(const double *__restrict__ proj, const double *__restrict__ view,
                                const double *__restrict__ part_model, const double *__restrict__ lookat,
                                int n_parts, int n_frames, float *__restrict__ mvp)
{
  const int rows = n_parts + 1;
  long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)n_frames * rows * 16;
  if (gid >= total) return;
  const int e = (int)(gid & 15);
  const long long mat = gid >> 4;
  const int p = (int)(mat % rows);
  const int f = (int)(mat / rows);
  const int r = e & 3, c = e >> 2;
  double s;
  if (p == n_parts) {
    s = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += proj[k * 4 + r] * lookat[c * 4 + k];
  } else {
    const double *V = view + 16 * (long long)f;
    const double *M = part_model + 16 * ((long long)f * n_parts + p);
    double pv[4];   // row r of P*V
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double a = 0.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) a += proj[j * 4 + r] * V[k * 4 + j];
      pv[k] = a;
    }
    s = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) s += pv[k] * M[c * 4 + k];
  }
  mvp[gid] = (float)s;
}

// ------------------------------------------------------------------------------------------
// K1: vertex stage + primitive setup
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

// warp-aggregated counter increment (all currently converged lanes share one atomic)
__device__ __forceinline__ uint32_t agg_inc(uint32_t *ctr)
{
  const unsigned m = __activemask();
  const int lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if (lane == leader) base = atomicAdd(ctr, (uint32_t)__popc(m));
  base = __shfl_sync(m, base, leader);
  return base + __popc(m & ((1u << lane) - 1u));
}

// S5/S8 + culling + hand-over to the binning stage
__device__ __forceinline__ void emit_window_tri(WV a, WV b, WV c, const Dims &d, TriRec *recs, TriRec *big,
                                                uint32_t *ctr)
{
  long long area2 = (long long)(b.X - a.X) * (c.Y - a.Y) - (long long)(c.X - a.X) * (b.Y - a.Y);
  if (area2 == 0) return;
  if (area2 < 0) { WV t = b; b = c; c = t; area2 = -area2; }

  int xmin = min(a.X, min(b.X, c.X)), xmax = max(a.X, max(b.X, c.X));
  int ymin = min(a.Y, min(b.Y, c.Y)), ymax = max(a.Y, max(b.Y, c.Y));
  int i0 = (xmin - kSubpixHalf + (kSubpix - 1)) >> kSubpixBits;   // arithmetic shift = floor
  int i1 = (xmax - kSubpixHalf) >> kSubpixBits;
  int j0 = (ymin - kSubpixHalf + (kSubpix - 1)) >> kSubpixBits;
  int j1 = (ymax - kSubpixHalf) >> kSubpixBits;
  i0 = max(i0, 0); i1 = min(i1, d.W - 1);
  j0 = max(j0, 0); j1 = min(j1, d.H - 1);
  if (i0 > i1 || j0 > j1) return;

  // S8: depth plane anchored at vertex 0
  float dx1 = (float)(b.X - a.X), dy1 = (float)(b.Y - a.Y);
  float dx2 = (float)(c.X - a.X), dy2 = (float)(c.Y - a.Y);
  float dz1 = b.z - a.z, dz2 = c.z - a.z;
  float fa = __ll2float_rn(area2);
  float t1 = dz2 * dy1;
  float gxz = fmaf(dz1, dy2, -t1) / fa;
  float t2 = dz1 * dx2;
  float gyz = fmaf(dz2, dx1, -t2) / fa;

  TriRec r;
  r.x0 = a.X; r.y0 = a.Y; r.x1 = b.X; r.y1 = b.Y; r.x2 = c.X; r.y2 = c.Y;
  r.z0 = a.z; r.gx = gxz; r.gy = gyz;
  r.bx = (uint32_t)i0 | ((uint32_t)i1 << 16);
  r.by = (uint32_t)j0 | ((uint32_t)j1 << 16);
  r.pad = 0;

  const int tx0 = i0 / kTileW, tx1 = i1 / kTileW, ty0 = j0 / kTileH, ty1 = j1 / kTileH;
  const int nt = (tx1 - tx0 + 1) * (ty1 - ty0 + 1);
  if (nt > kBigTiles) {
    uint32_t pos = atomicAdd(&ctr[kCtrBig], 1u);
    if (pos < (uint32_t)kBigCapacity) { big[pos] = r; return; }
    // list full: fall through and bin it like any other triangle
  }
  uint32_t idx = agg_inc(&ctr[kCtrRec]);
  if (idx >= d.cap_rec) { atomicOr(&ctr[kCtrFlags], kFlagRecOverflow); return; }
  recs[idx] = r;
  for (int ty = ty0; ty <= ty1; ++ty)
    for (int tx = tx0; tx <= tx1; ++tx) atomicAdd(&ctr[kCtrTiles + ty * d.tiles_x + tx], 1u);
}

__device__ __noinline__ void clip_and_emit(V4 p0, V4 p1, V4 p2, const Dims &d, TriRec *recs, TriRec *big,
                                           uint32_t *ctr)
{
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
  for (int i = 2; i < n; ++i) emit_window_tri(wv[0], wv[i - 1], wv[i], d, recs, big, ctr);
}

__global__ void __launch_bounds__(256)
ruf_setup_kernel(Model m, const float *__restrict__ mvp_all, Dims d, float bg_z, TriRec *recs_all,
                 TriRec *big_all, uint32_t *ctr_all)
{
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int frame = blockIdx.y;
  if (t >= d.n_tris + 2) return;

  float3 a, b, c;
  uint32_t part;
  if (t < d.n_tris) {
    const float4 q0 = __ldg(m.v0 + t), q1 = __ldg(m.v1 + t), q2 = __ldg(m.v2 + t);
    a = make_float3(q0.x, q0.y, q0.z);
    b = make_float3(q1.x, q1.y, q1.z);
    c = make_float3(q2.x, q2.y, q2.z);
    part = __float_as_uint(q0.w);
  } else {
    // background quad (src/urdf_filter.cpp:591-596) as triangles (q0,q1,q2), (q0,q2,q3)
    a = make_float3(-100.f, -100.f, bg_z);
    if (t == d.n_tris) { b = make_float3(100.f, -100.f, bg_z); c = make_float3(100.f, 100.f, bg_z); }
    else               { b = make_float3(100.f, 100.f, bg_z);  c = make_float3(-100.f, 100.f, bg_z); }
    part = (uint32_t)d.n_parts;
  }
  if (part > (uint32_t)d.n_parts) return;

  const float4 *M = reinterpret_cast<const float4 *>(mvp_all + 16 * ((long long)frame * (d.n_parts + 1) + part));
  const float4 c0 = __ldg(M), c1 = __ldg(M + 1), c2 = __ldg(M + 2), c3 = __ldg(M + 3);
  V4 p0 = xform(c0, c1, c2, c3, a.x, a.y, a.z);
  V4 p1 = xform(c0, c1, c2, c3, b.x, b.y, b.z);
  V4 p2 = xform(c0, c1, c2, c3, c.x, c.y, c.z);
  if (!finite4(p0) || !finite4(p1) || !finite4(p2)) return;

  // Early outs that cannot change the result (DESIGN.md "Setup-stage rejects"):
  //  * all three vertices in front of the near plane: the clipper would return nothing;
  //  * all three beyond one side plane of the view volume by a 0.1 % margin: whatever the
  //    clipper keeps lies at least 0.3 px outside the viewport.
  {
    const float n0 = p0.z + p0.w, n1 = p1.z + p1.w, n2 = p2.z + p2.w;
    if (n0 < 0.0f && n1 < 0.0f && n2 < 0.0f) return;
    const float k = 1.001f;
    const float w0 = k * p0.w, w1 = k * p1.w, w2 = k * p2.w;
    if (p0.x > w0 && p1.x > w1 && p2.x > w2) return;
    if (-p0.x > w0 && -p1.x > w1 && -p2.x > w2) return;
    if (p0.y > w0 && p1.y > w1 && p2.y > w2) return;
    if (-p0.y > w0 && -p1.y > w1 && -p2.y > w2) return;
  }

  TriRec *recs = recs_all + (size_t)frame * d.cap_rec;
  TriRec *big = big_all + (size_t)frame * kBigCapacity;
  uint32_t *ctr = ctr_all + (size_t)frame * d.ctr_stride;

  bool need = false;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (!(plane_dist(p0, k, d.guard_x, d.guard_y) >= 0.0f)) need = true;
    if (!(plane_dist(p1, k, d.guard_x, d.guard_y) >= 0.0f)) need = true;
    if (!(plane_dist(p2, k, d.guard_x, d.guard_y) >= 0.0f)) need = true;
  }
  if (need) { clip_and_emit(p0, p1, p2, d, recs, big, ctr); return; }
  if (!(p0.w > 0.0f) || !(p1.w > 0.0f) || !(p2.w > 0.0f)) return;   // S4b
  WV w0, w1, w2;
  if (!to_window(p0, d.halfw, d.halfh, w0) || !to_window(p1, d.halfw, d.halfh, w1) ||
      !to_window(p2, d.halfw, d.halfh, w2))
    return;
  emit_window_tri(w0, w1, w2, d, recs, big, ctr);
}

// ------------------------------------------------------------------------------------------
// K2: per-frame exclusive scan of tile counts -> offsets; overflow detection
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ruf_scan_kernel(Dims d, uint32_t *ctr_all, uint32_t *status)
{
  __shared__ uint32_t warp_sum[8];
  __shared__ uint32_t carry_s;
  uint32_t *ctr = ctr_all + (size_t)blockIdx.x * d.ctr_stride;
  uint32_t *cnt = ctr + kCtrTiles;
  uint32_t *off = ctr + kCtrTiles + 2 * d.ntiles;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < d.ntiles; base += 256) {
    const int i = base + threadIdx.x;
    uint32_t v = (i < d.ntiles) ? cnt[i] : 0u;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sum[warp] = x;
    __syncthreads();
    uint32_t wpre = 0;
    for (int w = 0; w < warp; ++w) wpre += warp_sum[w];
    const uint32_t carry = carry_s;
    if (i < d.ntiles) off[i] = carry + wpre + x - v;
    __syncthreads();
    if (threadIdx.x == 255) carry_s = carry + wpre + x;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    uint32_t total = carry_s;
    ctr[kCtrBinTotal] = total;
    uint32_t flags = ctr[kCtrFlags];
    if (total > d.cap_bin) flags |= kFlagBinOverflow;
    ctr[kCtrFlags] = flags;
    if (flags) atomicOr(status, flags);
  }
}

// ------------------------------------------------------------------------------------------
// K3: scatter the kept records into their tiles' bins (grid-stride over the frame's records)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ruf_bin_kernel(Dims d, const TriRec *__restrict__ recs_all, TriRec *bins_all, uint32_t *ctr_all)
{
  const int frame = blockIdx.y;
  uint32_t *ctr = ctr_all + (size_t)frame * d.ctr_stride;
  const uint32_t n = min(ctr[kCtrRec], d.cap_rec);
  const TriRec *recs = recs_all + (size_t)frame * d.cap_rec;
  TriRec *bins = bins_all + (size_t)frame * d.cap_bin;
  uint32_t *cur = ctr + kCtrTiles + d.ntiles;
  const uint32_t *off = ctr + kCtrTiles + 2 * d.ntiles;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint4 *src = reinterpret_cast<const uint4 *>(recs + i);
    const uint4 q0 = src[0], q1 = src[1], q2 = src[2];
    const uint32_t bx = q2.y, by = q2.z;
    const int tx0 = (int)(bx & 0xffffu) / kTileW, tx1 = (int)(bx >> 16) / kTileW;
    const int ty0 = (int)(by & 0xffffu) / kTileH, ty1 = (int)(by >> 16) / kTileH;
    for (int ty = ty0; ty <= ty1; ++ty)
      for (int tx = tx0; tx <= tx1; ++tx) {
        const int tile = ty * d.tiles_x + tx;
        const uint32_t pos = off[tile] + atomicAdd(&cur[tile], 1u);
        if (pos < d.cap_bin) {
          uint4 *dst = reinterpret_cast<uint4 *>(bins + pos);
          dst[0] = q0; dst[1] = q1; dst[2] = q2;
        }
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

// one lane walks the (tile-clipped) bbox of a small triangle; 32-bit edge functions are exact
// because every factor is below 2^14 (extent < 64 px).
__device__ __forceinline__ void raster_small(const TriRec &r, int i0, int i1, int j0, int j1, int tile_x0,
                                             int tile_y0, uint32_t *sz)
{
  const Edges e = make_edges(r);
  const int px0 = i0 * kSubpix + kSubpixHalf;
  int py = j0 * kSubpix + kSubpixHalf;
  int r0 = e.A0 * (px0 - r.x0) + e.B0 * (py - r.y0) + e.bias0;
  int r1 = e.A1 * (px0 - r.x1) + e.B1 * (py - r.y1) + e.bias1;
  int r2 = e.A2 * (px0 - r.x2) + e.B2 * (py - r.y2) + e.bias2;
  const int sA0 = e.A0 * kSubpix, sA1 = e.A1 * kSubpix, sA2 = e.A2 * kSubpix;
  const int sB0 = e.B0 * kSubpix, sB1 = e.B1 * kSubpix, sB2 = e.B2 * kSubpix;
  for (int j = j0; j <= j1; ++j, py += kSubpix) {
    const float rowz = fmaf(r.gy, (float)(py - r.y0), r.z0);
    int e0 = r0, e1 = r1, e2 = r2;
    uint32_t *row = sz + (j - tile_y0) * kTileW - tile_x0;
    int px = px0;
    for (int i = i0; i <= i1; ++i, px += kSubpix) {
      if ((e0 | e1 | e2) >= 0) {
        float z = clamp_z(fmaf(r.gx, (float)(px - r.x0), rowz));
        if (z < 1.0f) atomicMin(row + i, __float_as_uint(z));
      }
      e0 += sA0; e1 += sA1; e2 += sA2;
    }
    r0 += sB0; r1 += sB1; r2 += sB2;
  }
}

// a whole warp covers the tile-clipped bbox in 8x4 footprints; 64-bit edge functions
__device__ __forceinline__ void raster_warp(const TriRec &r, int i0, int i1, int j0, int j1, int tile_x0,
                                            int tile_y0, uint32_t *sz, int lane)
{
  const Edges e = make_edges(r);
  const int lx = lane & 7, ly = lane >> 3;
  for (int jb = j0; jb <= j1; jb += 4) {
    const int j = jb + ly;
    const int py = j * kSubpix + kSubpixHalf;
    const float rowz = fmaf(r.gy, (float)(py - r.y0), r.z0);
    const long long c0 = (long long)e.B0 * (py - r.y0) + e.bias0;
    const long long c1 = (long long)e.B1 * (py - r.y1) + e.bias1;
    const long long c2 = (long long)e.B2 * (py - r.y2) + e.bias2;
    for (int ib = i0; ib <= i1; ib += 8) {
      const int i = ib + lx;
      const int px = i * kSubpix + kSubpixHalf;
      const long long e0 = (long long)e.A0 * (px - r.x0) + c0;
      const long long e1 = (long long)e.A1 * (px - r.x1) + c1;
      const long long e2 = (long long)e.A2 * (px - r.x2) + c2;
      if (i <= i1 && j <= j1 && (e0 | e1 | e2) >= 0) {
        float z = clamp_z(fmaf(r.gx, (float)(px - r.x0), rowz));
        if (z < 1.0f) atomicMin(sz + (j - tile_y0) * kTileW + (i - tile_x0), __float_as_uint(z));
      }
    }
  }
}

// saturate_cast<ushort>(cvRound(x * 1000.f)) -- cv::Mat::convertTo(CV_16U, 1000.0), src/urdf_filter.cpp:311
__device__ __forceinline__ uint32_t f32_to_u16(float x)
{
  const float v = x * 1000.0f;
  if (!(v >= -2147483648.0f && v < 2147483648.0f)) return 0u;   // cvtss2si "indefinite" -> INT_MIN -> 0
  const int r = __float2int_rn(v);
  return (uint32_t)min(max(r, 0), 65535);
}

struct FragOut { float depth; uint32_t mask; };
// include/shaders/urdf_filter.frag:19-35 for the fragment that survived GL_LESS.
__device__ __forceinline__ FragOut fragment(float sensor, float zwin, const ShaderParams &sp)
{
  FragOut o;
  if (zwin == 1.0f) { o.depth = 0.0f; o.mask = 0u; return o; }   // never drawn: clear colour (:566)
  const float virt = sp.k1 / (zwin - sp.k2);                      // frag:14-17,22
  const bool s = sensor > (virt - sp.max_diff);                   // frag:23
  o.depth = s ? sp.replace_value : sensor;                        // frag:29, mix() with a in {0,1}
  o.mask = s ? 255u : 0u;                                         // frag:35 read back as UNSIGNED_BYTE
  return o;
}

template <int ENC>
__global__ void __launch_bounds__(kRasterThreads, 4)
ruf_raster_filter_kernel(Dims d, const TriRec *__restrict__ big_all, const TriRec *__restrict__ bins_all,
                         const uint32_t *__restrict__ ctr_all, ShaderParams sp, FrameBuffers fb)
{
  __shared__ __align__(128) TriRec sbuf[2][kChunk];
  __shared__ __align__(16) uint32_t sz[kTilePix];
  __shared__ __align__(8) uint64_t mbar[2];
  __shared__ uint32_t s_defer_n[2];
  __shared__ uint16_t s_defer[kChunk];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tile = blockIdx.x, frame = blockIdx.y;
  const int tile_x0 = (tile % d.tiles_x) * kTileW, tile_y0 = (tile / d.tiles_x) * kTileH;
  const uint32_t *ctr = ctr_all + (size_t)frame * d.ctr_stride;

  const uint32_t off = ctr[kCtrTiles + 2 * d.ntiles + tile];
  uint32_t cnt = ctr[kCtrTiles + tile];
  if (off >= d.cap_bin) cnt = 0; else cnt = min(cnt, d.cap_bin - off);
  const TriRec *bin = bins_all + (size_t)frame * d.cap_bin + off;
  const int nchunks = (int)((cnt + kChunk - 1) / kChunk);

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
    s_defer_n[0] = 0;
    s_defer_n[1] = 0;
  }
  __syncthreads();
  if (tid == 0 && nchunks > 0) {
    const uint32_t nrec = min(cnt, (uint32_t)kChunk);
    mbar_arrive_expect_tx(&mbar[0], nrec * (uint32_t)sizeof(TriRec));
    bulk_g2s(&sbuf[0][0], bin, nrec * (uint32_t)sizeof(TriRec), &mbar[0]);
  }

  // ---- big list: pixel-parallel, every thread owns 8 consecutive pixels of one tile row ----
  const int prow = tid >> 3, pcol = (tid & 7) * 8;
  float zr[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) zr[i] = 1.0f;       // glClear depth
  {
    const uint32_t nbig = min(ctr[kCtrBig], (uint32_t)kBigCapacity);
    const TriRec *big = big_all + (size_t)frame * kBigCapacity;
    const int px0 = (tile_x0 + pcol) * kSubpix + kSubpixHalf;
    const int py = (tile_y0 + prow) * kSubpix + kSubpixHalf;
    for (uint32_t b = 0; b < nbig; ++b) {
      const TriRec r = load_rec_global(big + b);
      const int bi0 = (int)(r.bx & 0xffffu), bi1 = (int)(r.bx >> 16);
      const int bj0 = (int)(r.by & 0xffffu), bj1 = (int)(r.by >> 16);
      if (bi1 < tile_x0 || bi0 >= tile_x0 + kTileW || bj1 < tile_y0 || bj0 >= tile_y0 + kTileH) continue;
      const Edges e = make_edges(r);
      long long e0 = (long long)e.A0 * (px0 - r.x0) + (long long)e.B0 * (py - r.y0) + e.bias0;
      long long e1 = (long long)e.A1 * (px0 - r.x1) + (long long)e.B1 * (py - r.y1) + e.bias1;
      long long e2 = (long long)e.A2 * (px0 - r.x2) + (long long)e.B2 * (py - r.y2) + e.bias2;
      const long long s0 = (long long)e.A0 * kSubpix, s1 = (long long)e.A1 * kSubpix,
                      s2 = (long long)e.A2 * kSubpix;
      const float rowz = fmaf(r.gy, (float)(py - r.y0), r.z0);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if ((e0 | e1 | e2) >= 0) {
          float z = clamp_z(fmaf(r.gx, (float)(px0 + i * kSubpix - r.x0), rowz));
          if (z < 1.0f) zr[i] = fminf(zr[i], z);
        }
        e0 += s0; e1 += s1; e2 += s2;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sz[prow * kTileW + pcol + i] = __float_as_uint(zr[i]);
  __syncthreads();

  // ---- binned triangles: chunks of 256 records double-buffered through the TMA engine ----
  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (tid == 0 && c + 1 < nchunks) {
      const uint32_t nrec = min(cnt - (uint32_t)(c + 1) * kChunk, (uint32_t)kChunk);
      mbar_arrive_expect_tx(&mbar[buf ^ 1], nrec * (uint32_t)sizeof(TriRec));
      bulk_g2s(&sbuf[buf ^ 1][0], bin + (size_t)(c + 1) * kChunk, nrec * (uint32_t)sizeof(TriRec),
               &mbar[buf ^ 1]);
    }
    mbar_wait(&mbar[buf], (uint32_t)((c >> 1) & 1));
    const uint32_t nrec = min(cnt - (uint32_t)c * kChunk, (uint32_t)kChunk);
    if ((uint32_t)tid < nrec) {
      const TriRec r = load_rec_smem(&sbuf[buf][tid]);
      const int i0 = max((int)(r.bx & 0xffffu), tile_x0), i1 = min((int)(r.bx >> 16), tile_x0 + kTileW - 1);
      const int j0 = max((int)(r.by & 0xffffu), tile_y0), j1 = min((int)(r.by >> 16), tile_y0 + kTileH - 1);
      const int ex = max(r.x0, max(r.x1, r.x2)) - min(r.x0, min(r.x1, r.x2));
      const int ey = max(r.y0, max(r.y1, r.y2)) - min(r.y0, min(r.y1, r.y2));
      const int area = (i1 - i0 + 1) * (j1 - j0 + 1);
      if (area <= kSmallArea && ex < 16384 && ey < 16384) {
        raster_small(r, i0, i1, j0, j1, tile_x0, tile_y0, sz);
      } else {
        s_defer[atomicAdd(&s_defer_n[buf], 1u)] = (uint16_t)tid;
      }
    }
    __syncthreads();
    // s_defer_n[buf ^ 1] was last read before the barrier that ended the previous iteration and
    // is next incremented after the barrier that ends this one: safe to clear here.
    if (tid == 0) s_defer_n[buf ^ 1] = 0;
    const uint32_t ndefer = s_defer_n[buf];
    for (uint32_t q = warp; q < ndefer; q += kRasterThreads / 32) {
      const TriRec r = load_rec_smem(&sbuf[buf][s_defer[q]]);
      const int i0 = max((int)(r.bx & 0xffffu), tile_x0), i1 = min((int)(r.bx >> 16), tile_x0 + kTileW - 1);
      const int j0 = max((int)(r.by & 0xffffu), tile_y0), j1 = min((int)(r.by >> 16), tile_y0 + kTileH - 1);
      raster_warp(r, i0, i1, j0, j1, tile_x0, tile_y0, sz, lane);
    }
    __syncthreads();
  }

  // ---- fused fragment stage: 8 pixels per thread, vector loads/stores ----
  const int gy = tile_y0 + prow, gx = tile_x0 + pcol;
  if (gy >= d.H || gx >= d.W) return;
  const size_t img = (size_t)frame * d.W * d.H;
  const size_t base = img + (size_t)gy * d.W + gx;
  const uint4 zq0 = *reinterpret_cast<const uint4 *>(&sz[prow * kTileW + pcol]);
  const uint4 zq1 = *reinterpret_cast<const uint4 *>(&sz[prow * kTileW + pcol + 4]);
  const float zw[8] = {__uint_as_float(zq0.x), __uint_as_float(zq0.y), __uint_as_float(zq0.z),
                       __uint_as_float(zq0.w), __uint_as_float(zq1.x), __uint_as_float(zq1.y),
                       __uint_as_float(zq1.z), __uint_as_float(zq1.w)};
  const bool full = fb.vec_ok && (gx + 8 <= d.W);
  if (full) {
    float sensor[8];
    if (ENC == 1) {
      const uint4 q = __ldg(reinterpret_cast<const uint4 *>(static_cast<const uint16_t *>(fb.depth_in) + base));
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        sensor[2 * i] = (float)(w[i] & 0xffffu) * 0.001f;      // convertTo(CV_32F, 0.001), :288
        sensor[2 * i + 1] = (float)(w[i] >> 16) * 0.001f;
      }
    } else {
      const float4 *p = reinterpret_cast<const float4 *>(static_cast<const float *>(fb.depth_in) + base);
      const float4 q0 = __ldg(p), q1 = __ldg(p + 1);
      sensor[0] = q0.x; sensor[1] = q0.y; sensor[2] = q0.z; sensor[3] = q0.w;
      sensor[4] = q1.x; sensor[5] = q1.y; sensor[6] = q1.z; sensor[7] = q1.w;
    }
    float od[8];
    uint32_t om[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const FragOut o = fragment(sensor[i], zw[i], sp);
      od[i] = o.depth; om[i] = o.mask;
    }
    if (ENC == 1) {
      uint32_t u[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = f32_to_u16(od[i]);   // convertTo(CV_16U, 1000), :311
      uint4 q;
      q.x = u[0] | (u[1] << 16); q.y = u[2] | (u[3] << 16); q.z = u[4] | (u[5] << 16); q.w = u[6] | (u[7] << 16);
      *reinterpret_cast<uint4 *>(static_cast<uint16_t *>(fb.depth_out) + base) = q;
    } else {
      float4 *p = reinterpret_cast<float4 *>(static_cast<float *>(fb.depth_out) + base);
      p[0] = make_float4(od[0], od[1], od[2], od[3]);
      p[1] = make_float4(od[4], od[5], od[6], od[7]);
    }
    if (fb.mask_out) {
      uint2 mq;
      mq.x = om[0] | (om[1] << 8) | (om[2] << 16) | (om[3] << 24);
      mq.y = om[4] | (om[5] << 8) | (om[6] << 16) | (om[7] << 24);
      *reinterpret_cast<uint2 *>(fb.mask_out + base) = mq;
    }
    if (fb.zbuf_out) {
      float4 *p = reinterpret_cast<float4 *>(fb.zbuf_out + base);
      p[0] = make_float4(zw[0], zw[1], zw[2], zw[3]);
      p[1] = make_float4(zw[4], zw[5], zw[6], zw[7]);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (gx + i >= d.W) break;
      float sensor;
      if (ENC == 1) sensor = (float)static_cast<const uint16_t *>(fb.depth_in)[base + i] * 0.001f;
      else sensor = static_cast<const float *>(fb.depth_in)[base + i];
      const FragOut o = fragment(sensor, zw[i], sp);
      if (ENC == 1) static_cast<uint16_t *>(fb.depth_out)[base + i] = (uint16_t)f32_to_u16(o.depth);
      else static_cast<float *>(fb.depth_out)[base + i] = o.depth;
      if (fb.mask_out) fb.mask_out[base + i] = (uint8_t)o.mask;
      if (fb.zbuf_out) fb.zbuf_out[base + i] = zw[i];
    }
  }
}

// ------------------------------------------------------------------------------------------
// host-side launcher for one batch
// ------------------------------------------------------------------------------------------
cudaError_t check_kernel_image()
{
  cudaFuncAttributes fa;
  return cudaFuncGetAttributes(&fa, (const void *)ruf_raster_filter_kernel<1>);
}

cudaError_t launch_frames(const Dims &d, const Model &m, const Workspace &ws, int n_frames,
                          const double *d_proj, const double *d_view, const double *d_part_model,
                          const double *d_lookat, float bg_z, int enc, const ShaderParams &sp,
                          const FrameBuffers &fb, cudaStream_t s, int *n_launches, cudaEvent_t *ev)
{
  cudaError_t err;
  int launches = 0;
  err = cudaMemsetAsync(ws.ctr, 0, (size_t)n_frames * d.ctr_stride * sizeof(uint32_t), s);
  if (err != cudaSuccess) return err;
  if (ev) cudaEventRecord(ev[0], s);
  {
    long long total = (long long)n_frames * (d.n_parts + 1) * 16;
    unsigned blocks = (unsigned)((total + 255) / 256);
    ruf_pose_kernel<<<blocks, 256, 0, s>>>(d_proj, d_view, d_part_model, d_lookat, d.n_parts, n_frames, ws.mvp);
    ++launches;
    if (ev) cudaEventRecord(ev[1], s);
  }
  {
    dim3 grid((unsigned)((d.n_tris + 2 + 255) / 256), (unsigned)n_frames);
    ruf_setup_kernel<<<grid, 256, 0, s>>>(m, ws.mvp, d, bg_z, ws.recs, ws.big, ws.ctr);
    ++launches;
    if (ev) cudaEventRecord(ev[2], s);
  }
  ruf_scan_kernel<<<(unsigned)n_frames, 256, 0, s>>>(d, ws.ctr, ws.status);
  ++launches;
  if (ev) cudaEventRecord(ev[3], s);
  {
    unsigned per_frame = (unsigned)((d.cap_rec + 255) / 256);
    if (per_frame > 48) per_frame = 48;
    if (per_frame < 1) per_frame = 1;
    dim3 grid(per_frame, (unsigned)n_frames);
    ruf_bin_kernel<<<grid, 256, 0, s>>>(d, ws.recs, ws.bins, ws.ctr);
    ++launches;
    if (ev) cudaEventRecord(ev[4], s);
  }
  {
    dim3 grid((unsigned)d.ntiles, (unsigned)n_frames);
    if (enc == 1)
      ruf_raster_filter_kernel<1><<<grid, kRasterThreads, 0, s>>>(d, ws.big, ws.bins, ws.ctr, sp, fb);
    else
      ruf_raster_filter_kernel<0><<<grid, kRasterThreads, 0, s>>>(d, ws.big, ws.bins, ws.ctr, sp, fb);
    ++launches;
    if (ev) cudaEventRecord(ev[5], s);
  }
  if (n_launches) *n_launches = launches;
  return cudaGetLastError();
}

}  // namespace ruf
