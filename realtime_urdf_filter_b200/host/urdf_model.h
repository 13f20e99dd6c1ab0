// urdf_model.h -- URDF subset -> renderable parts (triangle soup + model-matrix recipe).
//
// Replaces, for the hot path's inputs, what the reference gets from urdfdom + Assimp + freeglut:
//   src/urdf_renderer.cpp:66-169   (URDFRenderer::initURDFModel / loadURDFModel / process_link)
//   src/renderable.cpp:80-170      (sphere / cylinder / box draw calls)
//   src/renderable.cpp:306-452     (mesh import + glScalef + glDrawElements)
// Only what the filter needs is parsed: links, their visual / collision arrays (origin, geometry).
#pragma once
#include <string>
#include <unordered_set>
#include <vector>

#include "ros_shim.h"

namespace ruf_host {

struct Pose {
  double xyz[3] = {0, 0, 0};
  double q[4] = {0, 0, 0, 1};   // x y z w, from rpy like urdf::Rotation::setFromRPY
};

enum GeometryType { GEOM_SPHERE, GEOM_BOX, GEOM_CYLINDER, GEOM_MESH };
struct Geometry {
  GeometryType type = GEOM_BOX;
  double dim[3] = {0, 0, 0};     // box size
  double radius = 0, length = 0; // sphere / cylinder
  std::string filename;          // mesh
  double scale[3] = {1, 1, 1};
};
struct Visual {
  Pose origin;
  Geometry geometry;
};
struct UrdfJoint {
  std::string name, type, parent, child;
  Pose origin;
  double axis[3] = {1, 0, 0};
};
struct UrdfLink {
  std::string name;
  std::vector<Visual> visual_array, collision_array;
};
struct UrdfModel {
  std::string name;
  std::vector<UrdfLink> links;
  std::vector<UrdfJoint> joints;
  std::string root_link;
  // urdf::Model::initString: false on malformed XML and on what urdfdom rejects (missing robot / link names,
  // duplicate names, no links, joints without a known type or with unknown links, not exactly one root link)
  bool initString(const std::string &xml, std::string *error = nullptr);
  // links in the order urdf::ModelInterface::getLinks yields them (std::map: sorted by name)
  std::vector<const UrdfLink *> getLinks() const;
};

// One draw call of the reference = one model matrix on the device ("part").
struct RenderablePart {
  std::string name;              // tf frame: tf_prefix + "/" + link name (src/urdf_renderer.cpp:159)
  int renderable = 0;            // index of the Renderable this part belongs to (a box has two parts)
  double off_q[4] = {0, 0, 0, 1};// link_offset (normalised when the matrix is built)
  double off_t[3] = {0, 0, 0};
  bool has_suffix = false;
  double suffix[16];             // glTranslatef / glScalef issued before the draw call
  size_t first_tri = 0, n_tris = 0;
};

// Mesh loading: binary or ASCII STL from a path; "package://pkg/rest" is resolved against the
// search roots (resource_retriever replacement).  Returns false if the file cannot be read.
bool load_stl(const std::string &path, std::vector<float> &tri_xyz, std::string *error = nullptr);
std::string resolve_resource(const std::string &url, const std::vector<std::string> &roots);

// URDFRenderer (src/urdf_renderer.cpp:44-201) without the GL: builds the parts at construction,
// refreshes link_to_fixed from tf per frame.
class URDFRenderer {
 public:
  URDFRenderer(const std::string &model_description, const std::string &tf_prefix, const std::string &cam_frame,
               const std::string &fixed_frame, TransformListener &tf, const std::string &geometry_type,
               double scale, const std::unordered_set<std::string> &ignore,
               const std::vector<std::string> &resource_roots = {});

  // update_link_transforms (:173-190): one lookupTransform(fixed_frame, part.name) per renderable; a
  // failed lookup silently re-uses the transform of the previous loop iteration.
  void update_link_transforms(const Time &stamp);
  // part model matrices (link_to_fixed * link_offset [* suffix]) appended to `out` (16 doubles each)
  void append_part_models(std::vector<double> &out) const;

  const std::vector<RenderablePart> &parts() const { return parts_; }
  const std::vector<float> &triangles() const { return tri_; }          // 9 floats per triangle
  const std::vector<uint32_t> &triangle_parts() const { return tri_part_; }
  size_t num_renderables() const { return renderable_name_.size(); }
  bool ok() const { return ok_; }
  const std::string &error() const { return error_; }                  // why the description was rejected
  // meshes that could not be loaded (only STL is read here; the reference reads anything Assimp does): their links are
  // silently left unfiltered unless the caller looks at this list
  const std::vector<std::string> &mesh_errors() const { return mesh_errors_; }

 private:
  void process_link(const UrdfLink &link);
  size_t add_part(const std::string &name, const Pose &origin, const double *suffix, const std::vector<float> &tris,
                  int renderable);

  std::string tf_prefix_, camera_frame_, fixed_frame_, geometry_type_;
  double scale_;
  std::unordered_set<std::string> ignore_;
  std::vector<std::string> roots_;
  TransformListener &tf_;
  bool ok_ = false;
  std::string error_;
  std::vector<std::string> mesh_errors_;
  std::vector<RenderablePart> parts_;
  std::vector<float> tri_;
  std::vector<uint32_t> tri_part_;
  std::vector<std::string> renderable_name_;
  std::vector<double> link_q_, link_t_;   // per renderable: 4 / 3 doubles
};

}  // namespace ruf_host
