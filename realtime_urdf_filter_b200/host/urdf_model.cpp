// urdf_model.cpp -- see urdf_model.h.
#include "urdf_model.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <unordered_set>

#include "../../include/ruf_b200.h"

namespace ruf_host {

// =================================================================================================
// logging + shim implementations
// =================================================================================================
static LogSink g_sink = nullptr;
void set_log_sink(LogSink s) { g_sink = s; }
void logf(LogLevel lvl, const char *fmt, ...)
{
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (g_sink) g_sink(lvl, buf);
  else if (lvl >= LOG_ERROR) fprintf(stderr, "[%s] %s\n", lvl == LOG_FATAL ? "FATAL" : "ERROR", buf);
}

static void quat_to_R(const double *q, double R[3][3])
{
  // tf::Matrix3x3::setRotation
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double d = x * x + y * y + z * z + w * w, s = 2.0 / d;
  const double xs = x * s, ys = y * s, zs = z * s;
  const double wx = w * xs, wy = w * ys, wz = w * zs, xx = x * xs, xy = x * ys, xz = x * zs;
  const double yy = y * ys, yz = y * zs, zz = z * zs;
  R[0][0] = 1.0 - (yy + zz); R[0][1] = xy - wz; R[0][2] = xz + wy;
  R[1][0] = xy + wz; R[1][1] = 1.0 - (xx + zz); R[1][2] = yz - wx;
  R[2][0] = xz - wy; R[2][1] = yz + wx; R[2][2] = 1.0 - (xx + yy);
}
static void R_to_quat(const double R[3][3], double *q)
{
  // tf::Matrix3x3::getRotation
  const double trace = R[0][0] + R[1][1] + R[2][2];
  double t[4];
  if (trace > 0.0) {
    double s = std::sqrt(trace + 1.0);
    t[3] = s * 0.5;
    s = 0.5 / s;
    t[0] = (R[2][1] - R[1][2]) * s; t[1] = (R[0][2] - R[2][0]) * s; t[2] = (R[1][0] - R[0][1]) * s;
  } else {
    const int i = R[0][0] < R[1][1] ? (R[1][1] < R[2][2] ? 2 : 1) : (R[0][0] < R[2][2] ? 2 : 0);
    const int j = (i + 1) % 3, k = (i + 2) % 3;
    double s = std::sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0);
    t[i] = s * 0.5;
    s = 0.5 / s;
    t[3] = (R[k][j] - R[j][k]) * s; t[j] = (R[j][i] + R[i][j]) * s; t[k] = (R[k][i] + R[i][k]) * s;
  }
  for (int a = 0; a < 4; ++a) q[a] = t[a];
}

void TransformListener::setTransform(const std::string &frame, const double q[4], const double t[3])
{
  Pose p;
  quat_to_R(q, p.R);
  for (int i = 0; i < 3; ++i) p.t[i] = t[i];
  frames_[norm(frame)] = p;
}

void TransformListener::lookupTransform(const std::string &target, const std::string &source, const Time &time,
                                        StampedTransform &out) const
{
  ++lookups;
  auto a = frames_.find(norm(target)), b = frames_.find(norm(source));
  if (a == frames_.end()) throw TransformException("\"" + target + "\" passed to lookupTransform argument target_frame does not exist.");
  if (b == frames_.end()) throw TransformException("\"" + source + "\" passed to lookupTransform argument source_frame does not exist.");
  // target_T_source = inverse(root_T_target) * root_T_source
  const Pose &T = a->second, &S = b->second;
  double R[3][3], t[3];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) R[i][j] = T.R[0][i] * S.R[0][j] + T.R[1][i] * S.R[1][j] + T.R[2][i] * S.R[2][j];
    t[i] = T.R[0][i] * (S.t[0] - T.t[0]) + T.R[1][i] * (S.t[1] - T.t[1]) + T.R[2][i] * (S.t[2] - T.t[2]);
  }
  R_to_quat(R, out.q);
  for (int i = 0; i < 3; ++i) out.t[i] = t[i];
  out.stamp = time;
}

void NodeHandle::setCameraOffset(const double t[3], const double q[4])
{
  for (int i = 0; i < 3; ++i) off_t_[i] = t[i];
  for (int i = 0; i < 4; ++i) off_q_[i] = q[i];
  has_offset_ = true;
}
bool NodeHandle::getCameraOffset(double t[3], double q[4]) const
{
  if (!has_offset_) return false;
  for (int i = 0; i < 3; ++i) t[i] = off_t_[i];
  for (int i = 0; i < 4; ++i) q[i] = off_q_[i];
  return true;
}
bool NodeHandle::getParam(const std::string &k, std::string &v) const
{
  auto it = s_.find(k);
  if (it == s_.end()) return false;
  v = it->second;
  return true;
}
bool NodeHandle::getParam(const std::string &k, double &v) const
{
  auto it = d_.find(k);
  if (it == d_.end()) return false;
  v = it->second;
  return true;
}
bool NodeHandle::getParam(const std::string &k, bool &v) const
{
  auto it = b_.find(k);
  if (it == b_.end()) return false;
  v = it->second;
  return true;
}

// =================================================================================================
// a very small XML reader (elements + attributes; text, comments, PIs and DOCTYPE are skipped)
// =================================================================================================
namespace {
struct XmlNode {
  std::string tag;
  std::vector<std::pair<std::string, std::string>> attr;
  std::vector<XmlNode> children;
  const std::string *get(const char *k) const
  {
    for (auto &a : attr)
      if (a.first == k) return &a.second;
    return nullptr;
  }
  const XmlNode *child(const char *t) const
  {
    for (auto &c : children)
      if (c.tag == t) return &c;
    return nullptr;
  }
};

struct XmlParser {
  const std::string &s;
  size_t i = 0;
  std::string err;
  explicit XmlParser(const std::string &str) : s(str) {}
  void skip_ws() { while (i < s.size() && std::isspace((unsigned char)s[i])) ++i; }
  bool starts(const char *p) const { return s.compare(i, std::strlen(p), p) == 0; }
  bool skip_misc()
  {
    for (;;) {
      while (i < s.size() && s[i] != '<') ++i;   // text nodes are irrelevant for URDF
      if (i >= s.size()) return false;
      if (starts("<!--")) { size_t e = s.find("-->", i); if (e == std::string::npos) { err = "unterminated comment"; return false; } i = e + 3; continue; }
      if (starts("<?")) { size_t e = s.find("?>", i); if (e == std::string::npos) { err = "unterminated PI"; return false; } i = e + 2; continue; }
      if (starts("<!")) { size_t e = s.find('>', i); if (e == std::string::npos) { err = "unterminated declaration"; return false; } i = e + 1; continue; }
      return true;
    }
  }
  static std::string decode_entities(const std::string &v)
  {
    if (v.find('&') == std::string::npos) return v;
    static const struct { const char *ent; char ch; } kEnt[] = {{"&amp;", '&'}, {"&lt;", '<'}, {"&gt;", '>'}, {"&quot;", '"'}, {"&apos;", '\''}};
    std::string o;
    for (size_t k = 0; k < v.size();) {
      bool hit = false;
      if (v[k] == '&')
        for (const auto &e : kEnt) {
          const size_t n = std::strlen(e.ent);
          if (v.compare(k, n, e.ent) == 0) { o.push_back(e.ch); k += n; hit = true; break; }
        }
      if (!hit) o.push_back(v[k++]);
    }
    return o;
  }
  static bool name_char(char c) { return std::isalnum((unsigned char)c) || c == '_' || c == ':' || c == '-' || c == '.'; }
  int depth = 0;
  struct DepthGuard { int &d; explicit DepthGuard(int &x) : d(x) { ++d; } ~DepthGuard() { --d; } };
  bool parse_element(XmlNode &n)
  {
    // s[i] == '<'
    DepthGuard guard(depth);
    if (depth > 256) { err = "elements nested deeper than 256"; return false; }   // recursion bound for hostile input
    ++i;
    size_t b = i;
    while (i < s.size() && name_char(s[i])) ++i;
    n.tag = s.substr(b, i - b);
    if (n.tag.empty()) { err = "empty tag name"; return false; }
    for (;;) {
      skip_ws();
      if (i >= s.size()) { err = "unexpected end inside <" + n.tag + ">"; return false; }
      if (s[i] == '/') { if (i + 1 < s.size() && s[i + 1] == '>') { i += 2; return true; } err = "stray '/'"; return false; }
      if (s[i] == '>') { ++i; break; }
      size_t kb = i;
      while (i < s.size() && name_char(s[i])) ++i;
      std::string key = s.substr(kb, i - kb);
      skip_ws();
      if (key.empty() || i >= s.size() || s[i] != '=') { err = "malformed attribute in <" + n.tag + ">"; return false; }
      ++i;
      skip_ws();
      if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) { err = "unquoted attribute in <" + n.tag + ">"; return false; }
      const char qc = s[i++];
      size_t vb = i;
      while (i < s.size() && s[i] != qc) ++i;
      if (i >= s.size()) { err = "unterminated attribute value"; return false; }
      n.attr.emplace_back(key, decode_entities(s.substr(vb, i - vb)));
      ++i;
    }
    // children until </tag>
    for (;;) {
      if (!skip_misc()) { if (err.empty()) err = "missing </" + n.tag + ">"; return false; }
      if (starts("</")) {
        size_t e = s.find('>', i);
        if (e == std::string::npos) { err = "unterminated end tag"; return false; }
        std::string t = s.substr(i + 2, e - i - 2);
        while (!t.empty() && std::isspace((unsigned char)t.back())) t.pop_back();
        i = e + 1;
        if (t != n.tag) { err = "mismatched </" + t + "> for <" + n.tag + ">"; return false; }
        return true;
      }
      XmlNode c;
      if (!parse_element(c)) return false;
      n.children.push_back(std::move(c));
    }
  }
  bool parse(XmlNode &root)
  {
    if (!skip_misc()) { if (err.empty()) err = "no root element"; return false; }
    return parse_element(root);
  }
};

bool parse_doubles(const std::string &str, double *out, int n)
{
  std::istringstream is(str);
  for (int k = 0; k < n; ++k)
    if (!(is >> out[k])) return false;
  return true;
}

// urdf::Rotation::setFromRPY
void rpy_to_quat(const double rpy[3], double q[4])
{
  const double phi = rpy[0] / 2.0, the = rpy[1] / 2.0, psi = rpy[2] / 2.0;
  q[0] = std::sin(phi) * std::cos(the) * std::cos(psi) - std::cos(phi) * std::sin(the) * std::sin(psi);
  q[1] = std::cos(phi) * std::sin(the) * std::cos(psi) + std::sin(phi) * std::cos(the) * std::sin(psi);
  q[2] = std::cos(phi) * std::cos(the) * std::sin(psi) - std::sin(phi) * std::sin(the) * std::cos(psi);
  q[3] = std::cos(phi) * std::cos(the) * std::cos(psi) + std::sin(phi) * std::sin(the) * std::sin(psi);
  const double s = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (s == 0.0) { q[0] = q[1] = q[2] = 0.0; q[3] = 1.0; }
  else for (int k = 0; k < 4; ++k) q[k] /= s;
}

bool parse_pose(const XmlNode *o, Pose &p)
{
  if (!o) return true;
  double rpy[3] = {0, 0, 0};
  if (const std::string *v = o->get("xyz")) if (!parse_doubles(*v, p.xyz, 3)) return false;
  if (const std::string *v = o->get("rpy")) if (!parse_doubles(*v, rpy, 3)) return false;
  rpy_to_quat(rpy, p.q);
  return true;
}

bool parse_geometry(const XmlNode *g, Geometry &out, std::string *err)
{
  if (!g || g->children.empty()) { if (err) *err = "missing <geometry>"; return false; }
  const XmlNode &s = g->children[0];
  if (s.tag == "box") {
    out.type = GEOM_BOX;
    const std::string *v = s.get("size");
    if (!v || !parse_doubles(*v, out.dim, 3)) { if (err) *err = "bad box size"; return false; }
  } else if (s.tag == "sphere") {
    out.type = GEOM_SPHERE;
    const std::string *v = s.get("radius");
    if (!v || !parse_doubles(*v, &out.radius, 1)) { if (err) *err = "bad sphere radius"; return false; }
  } else if (s.tag == "cylinder") {
    out.type = GEOM_CYLINDER;
    const std::string *r = s.get("radius"), *l = s.get("length");
    if (!r || !l || !parse_doubles(*r, &out.radius, 1) || !parse_doubles(*l, &out.length, 1)) { if (err) *err = "bad cylinder"; return false; }
  } else if (s.tag == "mesh") {
    out.type = GEOM_MESH;
    const std::string *f = s.get("filename");
    if (!f) { if (err) *err = "mesh without filename"; return false; }
    out.filename = *f;
    if (const std::string *sc = s.get("scale")) if (!parse_doubles(*sc, out.scale, 3)) { if (err) *err = "bad mesh scale"; return false; }
  } else {
    if (err) *err = "unknown geometry <" + s.tag + ">";
    return false;
  }
  return true;
}
}  // namespace

// Follows urdfdom's parseURDF / Model::initTree / Model::initRoot (the parser behind urdf::Model::initString,
// src/urdf_renderer.cpp:69-70) in what it accepts and rejects: robot and link names are required, link and joint
// names unique, at least one link, joints need a known type, a parent and a child that exist (revolute and
// prismatic ones a <limit>), and the links form a tree with exactly one root.
bool UrdfModel::initString(const std::string &xml, std::string *error)
{
  links.clear(); joints.clear(); name.clear(); root_link.clear();
  auto fail = [&](const std::string &msg) { if (error) *error = msg; links.clear(); joints.clear(); return false; };
  XmlNode root;
  XmlParser p(xml);
  if (!p.parse(root)) return fail("XML: " + p.err);
  if (root.tag != "robot") return fail("Could not find the 'robot' element in the xml file");
  const std::string *rn = root.get("name");
  if (!rn) return fail("No name given for the robot.");
  name = *rn;
  std::unordered_set<std::string> seen;
  for (const XmlNode &c : root.children) {
    if (c.tag != "link") continue;
    UrdfLink l;
    const std::string *n = c.get("name");
    if (!n) return fail("No name given for the link.");
    l.name = *n;
    if (!seen.insert(l.name).second) return fail("link '" + l.name + "' is not unique.");
    for (const XmlNode &v : c.children) {
      if (v.tag != "visual" && v.tag != "collision") continue;
      Visual vis;
      std::string e;
      if (!parse_pose(v.child("origin"), vis.origin) || !parse_geometry(v.child("geometry"), vis.geometry, &e))
        return fail("link " + l.name + ": " + (e.empty() ? "bad origin" : e));
      (v.tag == "visual" ? l.visual_array : l.collision_array).push_back(vis);
    }
    links.push_back(std::move(l));
  }
  if (links.empty()) return fail("No link elements found in urdf file");
  std::unordered_set<std::string> jseen, children;
  for (const XmlNode &c : root.children) {
    if (c.tag != "joint") continue;
    UrdfJoint j;
    const std::string *n = c.get("name");
    if (!n) return fail("unnamed joint found");
    j.name = *n;
    if (!jseen.insert(j.name).second) return fail("joint '" + j.name + "' is not unique.");
    const std::string *t = c.get("type");
    if (!t) return fail("joint [" + j.name + "] has no type, check to see if it's a reference.");
    j.type = *t;
    static const char *kTypes[] = {"planar", "floating", "revolute", "continuous", "prismatic", "fixed"};
    bool known = false;
    for (const char *k : kTypes) known = known || j.type == k;
    if (!known) return fail("Joint [" + j.name + "] has no known type [" + j.type + "]");
    if ((j.type == "revolute" || j.type == "prismatic") && !c.child("limit"))
      return fail("Joint [" + j.name + "] is of type " + (j.type == "revolute" ? "REVOLUTE" : "PRISMATIC") +
                  " but it does not specify limits");
    if (const XmlNode *pn = c.child("parent")) if (const std::string *l = pn->get("link")) j.parent = *l;
    if (const XmlNode *cn = c.child("child")) if (const std::string *l = cn->get("link")) j.child = *l;
    if (!parse_pose(c.child("origin"), j.origin)) return fail("joint " + j.name + ": bad origin");
    if (const XmlNode *a = c.child("axis"))
      if (const std::string *v = a->get("xyz"))
        if (!parse_doubles(*v, j.axis, 3)) return fail("Malformed axis element for joint [" + j.name + "]");
    joints.push_back(std::move(j));
  }
  // Model::initTree
  for (const UrdfJoint &j : joints) {
    if (j.parent.empty() || j.child.empty())
      return fail("Joint [" + j.name + "] is missing a parent and/or child link specification.");
    if (!seen.count(j.child)) return fail("child link [" + j.child + "] of joint [" + j.name + "] not found");
    if (!seen.count(j.parent)) return fail("parent link [" + j.parent + "] of joint [" + j.name + "] not found");
    children.insert(j.child);
  }
  // Model::initRoot
  for (const UrdfLink &l : links) {
    if (children.count(l.name)) continue;
    if (!root_link.empty()) return fail("Two root links found: [" + root_link + "] and [" + l.name + "]");
    root_link = l.name;
  }
  if (root_link.empty()) return fail("No root link found. The robot xml is not a valid tree.");
  return true;
}

// urdf::ModelInterface::getLinks walks links_, a std::map keyed by the link name: the renderables are created in
// the byte-wise lexicographic order of the link names, not in document order.  The order is visible through the
// "failed lookup re-uses the previous iteration's transform" quirk of update_link_transforms.
std::vector<const UrdfLink *> UrdfModel::getLinks() const
{
  std::vector<const UrdfLink *> out;
  for (const UrdfLink &l : links) out.push_back(&l);
  std::sort(out.begin(), out.end(), [](const UrdfLink *a, const UrdfLink *b) { return a->name < b->name; });
  return out;
}

// =================================================================================================
// STL
// =================================================================================================
std::string resolve_resource(const std::string &url, const std::vector<std::string> &roots)
{
  std::string rest = url;
  if (url.compare(0, 10, "package://") == 0) rest = url.substr(10);
  else if (url.compare(0, 7, "file://") == 0) return url.substr(7);
  else return url;
  for (const std::string &r : roots) {
    std::string cand = r + "/" + rest;
    std::ifstream f(cand, std::ios::binary);
    if (f.good()) return cand;
    // the package directory itself may be the root
    size_t slash = rest.find('/');
    if (slash != std::string::npos) {
      cand = r + "/" + rest.substr(slash + 1);
      std::ifstream g(cand, std::ios::binary);
      if (g.good()) return cand;
    }
  }
  return rest;
}

bool load_stl(const std::string &path, std::vector<float> &tri, std::string *error)
{
  std::ifstream f(path, std::ios::binary);
  if (!f.good()) { if (error) *error = "cannot open " + path; return false; }
  std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  tri.clear();
  // Binary STL: 80-byte header, uint32 count, 50 bytes per facet.  Decide by size, not by the
  // "solid" prefix: many binary files start with "solid" (reference README.md:121-141).
  if (data.size() >= 84) {
    uint32_t n;
    std::memcpy(&n, data.data() + 80, 4);
    // exporters pad or append (colour blocks, a trailing newline): the declared facets must fit, bytes after them are ignored.
    // A text file that happens to pass is told apart by its "facet" keyword: no binary header + count + facets holds one
    // at the position where an ASCII file's first facet starts.
    const bool fits = n > 0 && data.size() >= 84 + (size_t)n * 50;
    const bool looks_ascii = data.compare(0, 5, "solid") == 0 && data.find("facet", 5) != std::string::npos &&
                             data.find("vertex", 5) != std::string::npos && data.size() != 84 + (size_t)n * 50;
    if (fits && !looks_ascii) {
      tri.resize((size_t)n * 9);
      for (uint32_t i = 0; i < n; ++i) std::memcpy(&tri[(size_t)i * 9], data.data() + 84 + (size_t)i * 50 + 12, 36);
      return true;
    }
  }
  // ASCII
  std::istringstream is(data);
  std::string tok;
  while (is >> tok) {
    if (tok == "vertex") {
      float v[3];
      if (!(is >> v[0] >> v[1] >> v[2])) { if (error) *error = "bad vertex in " + path; return false; }
      tri.insert(tri.end(), v, v + 3);
    }
  }
  if (tri.empty() || tri.size() % 9 != 0) { if (error) *error = "not an STL file: " + path; tri.clear(); return false; }
  return true;
}

// =================================================================================================
// URDFRenderer
// =================================================================================================
URDFRenderer::URDFRenderer(const std::string &model_description, const std::string &tf_prefix,
                           const std::string &cam_frame, const std::string &fixed_frame, TransformListener &tf,
                           const std::string &geometry_type, double scale,
                           const std::unordered_set<std::string> &ignore, const std::vector<std::string> &roots)
    : tf_prefix_(tf_prefix), camera_frame_(cam_frame), fixed_frame_(fixed_frame), geometry_type_(geometry_type),
      scale_(scale), ignore_(ignore), roots_(roots), tf_(tf)
{
  UrdfModel model;
  std::string err;
  if (!model.initString(model_description, &err)) {
    logf(LOG_ERROR, "URDF failed Model parse: %s", err.c_str());      // src/urdf_renderer.cpp:71-75
    error_ = err;
    return;
  }
  logf(LOG_INFO, "URDF parsed OK");
  for (const UrdfLink *l : model.getLinks()) process_link(*l);        // :84-96 (std::map order)
  logf(LOG_INFO, "Loaded %zu renderables", num_renderables());
  ok_ = true;
}

size_t URDFRenderer::add_part(const std::string &name, const Pose &origin, const double *suffix,
                              const std::vector<float> &tris, int renderable)
{
  RenderablePart p;
  p.name = name;
  p.renderable = renderable;
  for (int i = 0; i < 4; ++i) p.off_q[i] = origin.q[i];
  for (int i = 0; i < 3; ++i) p.off_t[i] = origin.xyz[i];
  p.has_suffix = suffix != nullptr;
  if (suffix) std::memcpy(p.suffix, suffix, sizeof(p.suffix));
  p.first_tri = tri_.size() / 9;
  p.n_tris = tris.size() / 9;
  tri_.insert(tri_.end(), tris.begin(), tris.end());
  tri_part_.insert(tri_part_.end(), p.n_tris, (uint32_t)parts_.size());
  parts_.push_back(p);
  return parts_.size() - 1;
}

static void scale_matrix(float sx, float sy, float sz, double *m)
{
  std::memset(m, 0, 16 * sizeof(double));
  m[0] = sx; m[5] = sy; m[10] = sz; m[15] = 1.0;
}

void URDFRenderer::process_link(const UrdfLink &link)
{
  if (ignore_.count(link.name)) return;                                // :107
  const std::vector<Visual> *arr = nullptr;
  if (geometry_type_.empty() || geometry_type_ == "visual") arr = &link.visual_array;
  else if (geometry_type_ == "collision") arr = &link.collision_array;
  else { logf(LOG_FATAL, "invalid geometry type: %s", geometry_type_.c_str()); return; }   // :127-130

  for (const Visual &v : *arr) {
    const Geometry &g = v.geometry;
    const std::string name = tf_prefix_ + "/" + link.name;             // :159
    const int rid = (int)num_renderables();
    std::vector<float> tris;
    double sfx[16];
    if (g.type == GEOM_BOX) {
      // RenderableBox(scale*x, scale*y, scale*z): float members; render() = VBO box, then
      // glScalef(dimx,dimy,dimz); glutSolidCube(dimx)   (src/renderable.cpp:107-131)
      const float dx = (float)(scale_ * g.dim[0]), dy = (float)(scale_ * g.dim[1]), dz = (float)(scale_ * g.dim[2]);
      tris.resize(12 * 9);
      ruf_box_triangles(dx, dy, dz, tris.data());
      add_part(name, v.origin, nullptr, tris, rid);
      ruf_cube_triangles(dx, tris.data());
      scale_matrix(dx, dy, dz, sfx);
      add_part(name, v.origin, sfx, tris, rid);
    } else if (g.type == GEOM_CYLINDER) {
      // glTranslatef(0,0,-length/2); glutSolidCylinder(radius,length,10,10)   (:92-98)
      const float r = (float)(scale_ * g.radius), l = (float)(scale_ * g.length);
      tris.resize((size_t)ruf_cylinder_triangle_count(10, 10) * 9);
      ruf_cylinder_triangles(r, l, 10, 10, tris.data());
      std::memset(sfx, 0, sizeof(sfx));
      sfx[0] = sfx[5] = sfx[10] = sfx[15] = 1.0;
      sfx[14] = (double)(-l / 2);
      add_part(name, v.origin, sfx, tris, rid);
    } else if (g.type == GEOM_SPHERE) {
      const float r = (float)(scale_ * g.radius);                       // glutSolidSphere(radius,10,10) (:80-85)
      tris.resize((size_t)ruf_sphere_triangle_count(10, 10) * 9);
      ruf_sphere_triangles(r, 10, 10, tris.data());
      add_part(name, v.origin, nullptr, tris, rid);
    } else {
      // RenderableMesh(filename, scale*sx, scale*sy, scale*sz): glScalef then the triangles as stored
      // in the file (root-node transform deliberately not applied, src/renderable.cpp:394-398)
      std::string err;
      const std::string path = resolve_resource(g.filename, roots_);
      if (!load_stl(path, tris, &err)) {
        logf(LOG_ERROR, "Could not load resource [%s]: %s", g.filename.c_str(), err.c_str());
        tris.clear();                                                  // the renderable exists but draws nothing ...
        mesh_errors_.push_back(name + ": " + g.filename + " (" + err + ")");   // ... and the caller can find out (strict_meshes)
      }
      scale_matrix((float)(scale_ * g.scale[0]), (float)(scale_ * g.scale[1]), (float)(scale_ * g.scale[2]), sfx);
      add_part(name, v.origin, sfx, tris, rid);
    }
    renderable_name_.push_back(name);
    link_q_.insert(link_q_.end(), {0.0, 0.0, 0.0, 1.0});
    link_t_.insert(link_t_.end(), {0.0, 0.0, 0.0});
  }
}

void URDFRenderer::update_link_transforms(const Time &stamp)
{
  StampedTransform t;                       // declared outside the loop, like the reference (:175)
  for (size_t r = 0; r < renderable_name_.size(); ++r) {
    try {
      tf_.lookupTransform(fixed_frame_, renderable_name_[r], stamp, t);
    } catch (const TransformException &ex) {
      logf(LOG_DEBUG, "%s", ex.what());     // swallowed: `t` keeps the previous iteration's value (:184-187)
    }
    for (int i = 0; i < 4; ++i) link_q_[4 * r + i] = t.q[i];
    for (int i = 0; i < 3; ++i) link_t_[3 * r + i] = t.t[i];
  }
}

void URDFRenderer::append_part_models(std::vector<double> &out) const
{
  for (const RenderablePart &p : parts_) {
    double m[16];
    ruf_part_model(&link_q_[4 * p.renderable], &link_t_[3 * p.renderable], p.off_q, p.off_t,
                   p.has_suffix ? p.suffix : nullptr, m);
    out.insert(out.end(), m, m + 16);
  }
}

}  // namespace ruf_host
