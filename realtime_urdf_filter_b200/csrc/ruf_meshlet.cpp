// ruf_meshlet.cpp -- see ruf_meshlet.h.  Pure host code (no CUDA), so that the CPU test-suite can
// check the round trip meshlets -> soup bit for bit (tests/test_meshlets.py).
#include "ruf_meshlet.h"

#include <cstring>
#include <unordered_map>

#include "../../include/ruf_b200.h"

namespace ruf {
namespace {
struct VKey {
  uint32_t x, y, z, p;
  bool operator==(const VKey &o) const { return x == o.x && y == o.y && z == o.z && p == o.p; }
};
struct VKeyHash {
  size_t operator()(const VKey &k) const
  {
    uint64_t h = 0x9e3779b97f4a7c15ull;
    const uint32_t w[4] = {k.x, k.y, k.z, k.p};
    for (uint32_t v : w) { h ^= v; h *= 0xff51afd7ed558ccdull; h ^= h >> 29; }
    return (size_t)h;
  }
};
inline uint32_t fbits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float bitsf(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
}  // namespace

void build_meshlets(const float *tri_xyz, const uint32_t *tri_part, int64_t n_tris, int n_parts, float bg_z,
                    int max_verts, int max_tris, int max_parts, MeshletModel &out)
{
  out = MeshletModel();
  std::vector<float> &box = out.part_aabb;
  box.assign((size_t)(n_parts > 0 ? n_parts : 1) * kPartStrideHost, 0.0f);
  for (int p = 0; p < n_parts; ++p)
    for (int k = 0; k < 3; ++k) { box[kPartStrideHost * p + k] = 3.0e38f; box[kPartStrideHost * p + 3 + k] = -3.0e38f; }
  std::vector<double> volume((size_t)(n_parts > 0 ? n_parts : 1), 0.0);   // 6 x signed volume of every part's mesh
  out.verts.reserve((size_t)(n_tris + 2) * 4);
  out.tris.reserve((size_t)n_tris + 2);
  std::unordered_map<VKey, uint32_t, VKeyHash> local;
  size_t vert_off = 0, tri_off = 0;      // first vertex / triangle of the open meshlet
  uint32_t plo = 0, phi = 0;             // its part window
  auto n_open_verts = [&]() { return out.verts.size() / 4 - vert_off; };
  auto n_open_tris = [&]() { return out.tris.size() - tri_off; };
  auto close = [&]() {
    const size_t nv = n_open_verts(), nt = n_open_tris();
    if (nt == 0) return;
    for (size_t v = vert_off; v < vert_off + nv; ++v)                        // part -> slot in the window
      out.verts[4 * v + 3] = bitsf(fbits(out.verts[4 * v + 3]) - plo);
    out.hdr.push_back((uint32_t)vert_off);
    out.hdr.push_back((uint32_t)tri_off);
    out.hdr.push_back((uint32_t)nv | ((uint32_t)nt << 10) | ((phi - plo) << 20));
    out.hdr.push_back(plo);
    vert_off += nv; tri_off += nt;
    local.clear();
  };
  auto add_triangle = [&](const float *xyz, uint32_t p) {
    VKey key[3];
    for (int v = 0; v < 3; ++v) key[v] = VKey{fbits(xyz[3 * v]), fbits(xyz[3 * v + 1]), fbits(xyz[3 * v + 2]), p};
    if (n_open_tris() > 0) {             // does it still fit the open meshlet?
      const uint32_t lo = p < plo ? p : plo, hi = p > phi ? p : phi;
      size_t fresh = 0;                  // (an upper bound when the triangle repeats a new vertex)
      for (int v = 0; v < 3; ++v) fresh += local.count(key[v]) ? 0 : 1;
      if (hi - lo >= (uint32_t)max_parts || n_open_tris() >= (size_t)max_tris || n_open_verts() + fresh > (size_t)max_verts)
        close();
    }
    if (n_open_tris() == 0) plo = phi = p;
    if (p < plo) plo = p;
    if (p > phi) phi = p;
    uint32_t idx[3];
    for (int v = 0; v < 3; ++v) {
      auto it = local.find(key[v]);
      if (it == local.end()) {
        it = local.emplace(key[v], (uint32_t)n_open_verts()).first;
        out.verts.push_back(xyz[3 * v]); out.verts.push_back(xyz[3 * v + 1]); out.verts.push_back(xyz[3 * v + 2]);
        out.verts.push_back(bitsf(p));   // rewritten as the slot by close()
      }
      idx[v] = it->second;
    }
    out.tris.push_back(idx[0] | (idx[1] << 10) | (idx[2] << 20));
  };
  for (int64_t t = 0; t < n_tris; ++t) {
    const uint32_t p = tri_part[t];
    if (p >= (uint32_t)n_parts) continue;
    for (int v = 0; v < 3; ++v)
      for (int k = 0; k < 3; ++k) {
        const float x = tri_xyz[9 * t + 3 * v + k];
        // NaN / inf vertices never tighten the box (such triangles are dropped by the vertex stage anyway)
        if (x < box[kPartStrideHost * p + k]) box[kPartStrideHost * p + k] = x;
        if (x > box[kPartStrideHost * p + 3 + k]) box[kPartStrideHost * p + 3 + k] = x;
        if (!(x == x) || x > 3.0e38f || x < -3.0e38f) { box[kPartStrideHost * p + k] = -3.0e38f; box[kPartStrideHost * p + 3 + k] = 3.0e38f; }
      }
    {
      const float *q = tri_xyz + 9 * t;
      const double ax = q[0], ay = q[1], az = q[2], bx = q[3], by = q[4], bz = q[5], cx = q[6], cy = q[7], cz = q[8];
      volume[p] += ax * (by * cz - bz * cy) - ay * (bx * cz - bz * cx) + az * (bx * cy - by * cx);
    }
    add_triangle(tri_xyz + 9 * t, p);
  }
  close();
  // winding of every part: +1 when its triangles are counter-clockwise seen from outside (positive signed
  // volume), -1 when clockwise.  Only a hint: it decides which half of a closed mesh the raster kernel draws
  // first (the half facing the camera) so that the other half can be depth-culled; it never changes a pixel.
  for (int p = 0; p < n_parts; ++p) box[kPartStrideHost * p + 6] = (volume[p] < 0.0) ? -1.0f : 1.0f;
  // background quad: glVertex3f(+-100, +-100, far_plane_*0.99) (src/urdf_filter.cpp:591-596) as the two
  // triangles (q0,q1,q2), (q0,q2,q3); it is drawn with MODELVIEW = LookAt, which is matrix row n_parts
  const float q[4][3] = {{-100.f, -100.f, bg_z}, {100.f, -100.f, bg_z}, {100.f, 100.f, bg_z}, {-100.f, 100.f, bg_z}};
  const float t0[9] = {q[0][0], q[0][1], q[0][2], q[1][0], q[1][1], q[1][2], q[2][0], q[2][1], q[2][2]};
  const float t1[9] = {q[0][0], q[0][1], q[0][2], q[2][0], q[2][1], q[2][2], q[3][0], q[3][1], q[3][2]};
  add_triangle(t0, (uint32_t)n_parts);
  add_triangle(t1, (uint32_t)n_parts);
  close();
}

void build_meshlet_sets(const float *tri_xyz, const uint32_t *tri_part, int64_t n_tris, int n_parts, float bg_z,
                        int max_verts, int max_tris, int fine_tris, int max_parts, MeshletModel &out)
{
  build_meshlets(tri_xyz, tri_part, n_tris, n_parts, bg_z, max_verts, max_tris, max_parts, out);
  out.n_primary = out.n_meshlets();
  MeshletModel fine;
  build_meshlets(tri_xyz, tri_part, n_tris, n_parts, bg_z, max_verts, fine_tris, max_parts, fine);
  const uint32_t voff = (uint32_t)(out.verts.size() / 4), toff = (uint32_t)out.tris.size();
  for (size_t m = 0; m < fine.n_meshlets(); ++m) {
    out.hdr.push_back(fine.hdr[4 * m] + voff);
    out.hdr.push_back(fine.hdr[4 * m + 1] + toff);
    out.hdr.push_back(fine.hdr[4 * m + 2]);
    out.hdr.push_back(fine.hdr[4 * m + 3]);
  }
  out.verts.insert(out.verts.end(), fine.verts.begin(), fine.verts.end());
  out.tris.insert(out.tris.end(), fine.tris.begin(), fine.tris.end());
}

}  // namespace ruf

// Diagnostics (CPU only): build the meshlets of a soup with the library's limits and expand them again.
// out_xyz: (n_tris + 2) * 9 floats, out_part: n_tris + 2 (the background quad comes last, part = n_parts).
// counts[0..2] = meshlets, welded vertices, triangles.  Returns RUF_OK or RUF_ERR_INVALID.
// expands the meshlets [m0, m1) of a model into a soup; checks the limits and the packing on the way
static int expand_meshlets(const ruf::MeshletModel &mm, size_t m0, size_t m1, uint32_t tri_base, int max_verts, int max_tris,
                           int max_parts, float *out_xyz, uint32_t *out_part, int64_t &t_out)
{
  t_out = 0;
  for (size_t m = m0; m < m1; ++m) {
    const uint32_t *h = &mm.hdr[4 * m];
    const uint32_t nv = h[2] & 1023u, nt = (h[2] >> 10) & 1023u, np = (h[2] >> 20) + 1;
    if ((int)nv > max_verts || (int)nt > max_tris || (int)np > max_parts || h[1] != tri_base + (uint32_t)t_out) return RUF_ERR_INVALID;
    for (uint32_t t = 0; t < nt; ++t, ++t_out) {
      const uint32_t ix = mm.tris[h[1] + t];
      const uint32_t id[3] = {ix & 1023u, (ix >> 10) & 1023u, ix >> 20};
      uint32_t slot = 0;
      for (int v = 0; v < 3; ++v) {
        if (id[v] >= nv) return RUF_ERR_INVALID;
        const float *p = &mm.verts[4 * (size_t)(h[0] + id[v])];
        uint32_t s;
        std::memcpy(&s, p + 3, 4);
        if (v > 0 && s != slot) return RUF_ERR_INVALID;       // one part per triangle
        slot = s;
        if (out_xyz) std::memcpy(out_xyz + 9 * t_out + 3 * v, p, 12);
      }
      if (slot >= np) return RUF_ERR_INVALID;
      if (out_part) out_part[t_out] = h[3] + slot;
    }
  }
  return RUF_OK;
}

extern "C" RUF_API int ruf_meshlet_roundtrip(const float *tri_xyz, const uint32_t *tri_part, int64_t n_tris, int n_parts,
                                             double z_far, int max_verts, int max_tris, int max_parts, float *out_xyz,
                                             uint32_t *out_part, int64_t *counts)
{
  if (n_tris < 0 || n_parts < 0 || (n_tris > 0 && (!tri_xyz || !tri_part)) || max_verts < 3 || max_verts > 1024 ||
      max_tris < 1 || max_tris > 1023 || max_parts < 1 || max_parts > 32)
    return RUF_ERR_INVALID;
  ruf::MeshletModel mm;
  ruf::build_meshlets(tri_xyz, tri_part, n_tris, n_parts, (float)(z_far * 0.99), max_verts, max_tris, max_parts, mm);
  int64_t t_out = 0;
  if (expand_meshlets(mm, 0, mm.n_meshlets(), 0u, max_verts, max_tris, max_parts, out_xyz, out_part, t_out) != RUF_OK) return RUF_ERR_INVALID;
  if (counts) { counts[0] = (int64_t)mm.n_meshlets(); counts[1] = (int64_t)(mm.verts.size() / 4); counts[2] = t_out; }
  return RUF_OK;
}

// The same for the two cuts ruf_set_model uploads in one set of arrays (build_meshlet_sets): the throughput cut first, then
// the fine cut.  out_xyz / out_part hold 2 * (n_tris + 2) triangles: both cuts expanded, each the soup + the background quad.
// counts[0..3] = meshlets of the first cut, meshlets of the fine cut, welded vertices, index triples (both cuts).
extern "C" RUF_API int ruf_meshlet_sets_roundtrip(const float *tri_xyz, const uint32_t *tri_part, int64_t n_tris, int n_parts,
                                                  double z_far, int max_verts, int max_tris, int fine_tris, int max_parts,
                                                  float *out_xyz, uint32_t *out_part, int64_t *counts)
{
  if (n_tris < 0 || n_parts < 0 || (n_tris > 0 && (!tri_xyz || !tri_part)) || max_verts < 3 || max_verts > 1024 ||
      max_tris < 1 || max_tris > 1023 || fine_tris < 1 || fine_tris > 1023 || max_parts < 1 || max_parts > 32)
    return RUF_ERR_INVALID;
  ruf::MeshletModel mm;
  ruf::build_meshlet_sets(tri_xyz, tri_part, n_tris, n_parts, (float)(z_far * 0.99), max_verts, max_tris, fine_tris, max_parts, mm);
  int64_t t0 = 0, t1 = 0;
  if (expand_meshlets(mm, 0, mm.n_primary, 0u, max_verts, max_tris, max_parts, out_xyz, out_part, t0) != RUF_OK) return RUF_ERR_INVALID;
  if (t0 != n_tris + 2) return RUF_ERR_INVALID;
  if (expand_meshlets(mm, mm.n_primary, mm.n_meshlets(), (uint32_t)t0, max_verts, fine_tris, max_parts,
                      out_xyz ? out_xyz + 9 * t0 : nullptr, out_part ? out_part + t0 : nullptr, t1) != RUF_OK)
    return RUF_ERR_INVALID;
  if (t1 != n_tris + 2 || (int64_t)mm.tris.size() != t0 + t1) return RUF_ERR_INVALID;
  if (counts) {
    counts[0] = (int64_t)mm.n_primary; counts[1] = (int64_t)(mm.n_meshlets() - mm.n_primary);
    counts[2] = (int64_t)(mm.verts.size() / 4); counts[3] = (int64_t)mm.tris.size();
  }
  return RUF_OK;
}
