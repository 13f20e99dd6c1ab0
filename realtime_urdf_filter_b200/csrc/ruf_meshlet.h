// ruf_meshlet.h -- model ingest (set-up time, host): triangle soup -> meshlets.
//
// Replaces what the reference does when it creates its VBO/IBO pairs (SubMesh::init
// src/renderable.cpp:339-350, createBoxVBO :133-170): the static geometry is laid out once in the
// form the per-frame vertex stage wants to read it.  A meshlet is a run of CONSECUTIVE triangles of
// the soup with
//   * at most max_tris triangles and max_verts distinct vertices -- bit-identical positions of the
//     same part are welded, so the setup kernel shades each of them once;
//   * parts within a window of max_parts consecutive part indices (one cull bit per part).
// The background quad of src/urdf_filter.cpp:591-596 (matrix row n_parts) is the last meshlet.
#pragma once
#include <cstdint>
#include <vector>

namespace ruf {

constexpr int kPartStrideHost = 8;   // floats per part in MeshletModel::part_aabb (== kPartStride of ruf_device.cuh)

struct MeshletModel {
  std::vector<uint32_t> hdr;     // 4 words per meshlet: vert_off, tri_off, nverts | ntris << 10 | (nparts - 1) << 20, lowest part
  std::vector<float> verts;      // 4 floats per vertex: x, y, z, bits(part - lowest part of its meshlet)
  std::vector<uint32_t> tris;    // local vertex indices i0 | i1 << 10 | i2 << 20
  std::vector<float> part_aabb;  // kPartStrideHost floats per part: object-space min xyz, max xyz (inverted when the part is
                                 // empty), winding (+1 counter-clockwise seen from outside, -1 clockwise), pad
  size_t n_meshlets() const { return hdr.size() / 4; }
  size_t n_primary = 0;          // build_meshlet_sets: headers [0, n_primary) are the throughput cut, the rest the fine cut
};

// bg_z: the background quad's z (float(0.99 * z_far)).  Triangles with tri_part >= n_parts are skipped.
void build_meshlets(const float *tri_xyz, const uint32_t *tri_part, int64_t n_tris, int n_parts, float bg_z,
                    int max_verts, int max_tris, int max_parts, MeshletModel &out);

// The model cut twice into the SAME arrays (one upload, one broadcast): headers [0, n_primary) with at most max_tris
// triangles per meshlet (throughput: a setup CTA amortises its meshlet over a run of frames) followed by a second cut
// with at most fine_tris (launches of one or a few frames: four times the CTAs, a quarter of the serial work each).
// Both cuts hold every triangle and the background quad once; a launch uses one of them.
void build_meshlet_sets(const float *tri_xyz, const uint32_t *tri_part, int64_t n_tris, int n_parts, float bg_z,
                        int max_verts, int max_tris, int fine_tris, int max_parts, MeshletModel &out);

}  // namespace ruf
