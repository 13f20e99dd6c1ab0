// ruf_host.cpp -- host-side (CPU, double precision) pieces of the hot path that the reference
// also evaluates on the host: the GL matrix stack contents and the primitive tessellations.
// Nothing here touches image data; per-pixel and per-triangle work lives in ruf_kernels.cu.
//
// Compile with -ffp-contract=off: operation order below is part of the contract with the
// parity tests (tests/test_host_math.py compares bit-for-bit against oracle/).
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/ruf_b200.h"

namespace {

// 4x4 column-major, the layout of glMultMatrixd / tf::Transform::getOpenGLMatrix.
struct Mat4 {
  double m[16];
  static Mat4 identity()
  {
    Mat4 r;
    for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.0 : 0.0;
    return r;
  }
};

// what glMultMatrixd does to the current matrix: C = A * B
Mat4 mul(const Mat4 &A, const Mat4 &B)
{
  Mat4 C;
  for (int col = 0; col < 4; ++col)
    for (int row = 0; row < 4; ++row) {
      double acc = 0.0;
      for (int k = 0; k < 4; ++k) acc += A.m[k * 4 + row] * B.m[col * 4 + k];
      C.m[col * 4 + row] = acc;
    }
  return C;
}

// tf::Transform (Bullet LinearMath conventions): row-major 3x3 basis + origin.
struct Rigid {
  double R[3][3];
  double t[3];

  // tf::Matrix3x3::setRotation(Quaternion(x,y,z,w))
  static Rigid from_quat(const double *q, const double *t)
  {
    Rigid X;
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double d = x * x + y * y + z * z + w * w;
    const double s = 2.0 / d;
    const double xs = x * s, ys = y * s, zs = z * s;
    const double wx = w * xs, wy = w * ys, wz = w * zs;
    const double xx = x * xs, xy = x * ys, xz = x * zs;
    const double yy = y * ys, yz = y * zs, zz = z * zs;
    X.R[0][0] = 1.0 - (yy + zz); X.R[0][1] = xy - wz;         X.R[0][2] = xz + wy;
    X.R[1][0] = xy + wz;         X.R[1][1] = 1.0 - (xx + zz); X.R[1][2] = yz - wx;
    X.R[2][0] = xz - wy;         X.R[2][1] = yz + wx;         X.R[2][2] = 1.0 - (xx + yy);
    X.t[0] = t[0]; X.t[1] = t[1]; X.t[2] = t[2];
    return X;
  }
  // tf::Transform::operator*
  Rigid operator*(const Rigid &o) const
  {
    Rigid r;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) r.R[i][j] = R[i][0] * o.R[0][j] + R[i][1] * o.R[1][j] + R[i][2] * o.R[2][j];
      r.t[i] = (R[i][0] * o.t[0] + R[i][1] * o.t[1] + R[i][2] * o.t[2]) + t[i];
    }
    return r;
  }
  // tf::Transform::inverse: (R^T, R^T * -t)
  Rigid inverse() const
  {
    Rigid r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r.R[i][j] = R[j][i];
    for (int i = 0; i < 3; ++i) r.t[i] = r.R[i][0] * -t[0] + r.R[i][1] * -t[1] + r.R[i][2] * -t[2];
    return r;
  }
  // tf::Transform::getOpenGLMatrix
  Mat4 gl() const
  {
    Mat4 g;
    for (int c = 0; c < 3; ++c) {
      for (int r = 0; r < 3; ++r) g.m[c * 4 + r] = R[r][c];
      g.m[c * 4 + 3] = 0.0;
    }
    g.m[12] = t[0]; g.m[13] = t[1]; g.m[14] = t[2]; g.m[15] = 1.0;
    return g;
  }
};

// gluLookAt as specified by GLU: rows (s, u, -f), then translate(-eye).
Mat4 glu_look_at(const double eye[3], const double center[3], const double up[3])
{
  double f[3] = {center[0] - eye[0], center[1] - eye[1], center[2] - eye[2]};
  double n = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
  for (double &v : f) v /= n;
  double s[3] = {f[1] * up[2] - f[2] * up[1], f[2] * up[0] - f[0] * up[2], f[0] * up[1] - f[1] * up[0]};
  n = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
  for (double &v : s) v /= n;
  const double u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};
  Mat4 R = Mat4::identity();
  R.m[0] = s[0]; R.m[4] = s[1]; R.m[8] = s[2];
  R.m[1] = u[0]; R.m[5] = u[1]; R.m[9] = u[2];
  R.m[2] = -f[0]; R.m[6] = -f[1]; R.m[10] = -f[2];
  Mat4 T = Mat4::identity();
  T.m[12] = -eye[0]; T.m[13] = -eye[1]; T.m[14] = -eye[2];
  return mul(R, T);
}

struct Soup {
  float *o;
  int n = 0;
  void tri(const float *a, const float *b, const float *c)
  {
    std::memcpy(o, a, 12); std::memcpy(o + 3, b, 12); std::memcpy(o + 6, c, 12);
    o += 9; ++n;
  }
  // GL_QUADS decomposition (a,b,c) + (a,c,d)
  void quad(const float *a, const float *b, const float *c, const float *d) { tri(a, b, c); tri(a, c, d); }
};

// freeglut's fghCircleTable: n+1 samples of sin/cos(2*pi*i/n), negative n runs clockwise
void circle_table(std::vector<double> &sint, std::vector<double> &cost, int n)
{
  const int size = std::abs(n);
  const double angle = 2.0 * M_PI / (double)((n == 0) ? 1 : n);
  sint.assign(size + 1, 0.0);
  cost.assign(size + 1, 1.0);
  for (int i = 1; i < size; ++i) { sint[i] = std::sin(angle * i); cost[i] = std::cos(angle * i); }
  sint[size] = sint[0]; cost[size] = cost[0];
}

}  // namespace

extern "C" {

void ruf_projection_matrix(const double *P, int width, int height, double z_near, double z_far,
                           double *glTf, double *camera_tx, double *camera_ty)
{
  const double fx = P[0], fy = P[5], cx = P[2], cy = P[6];
  if (camera_tx) *camera_tx = -1 * (P[3] / fx);
  if (camera_ty) *camera_ty = -1 * (P[7] / fy);
  for (int i = 0; i < 16; ++i) glTf[i] = 0.0;
  glTf[0] = -2.0 * fx / width;                  // the minus flips the image x axis
  glTf[5] = 2.0 * fy / height;
  glTf[8] = 2.0 * (0.5 - cx / width);
  glTf[9] = 2.0 * (cy / height - 0.5);
  glTf[10] = -(z_far + z_near) / (z_far - z_near);
  glTf[14] = -2.0 * z_far * z_near / (z_far - z_near);
  glTf[11] = -1;
}

void ruf_lookat(double *m)
{
  const double eye[3] = {0, 0, 0}, center[3] = {0, 0, 1}, up[3] = {0, 1, 0};
  const Mat4 L = glu_look_at(eye, center, up);
  std::memcpy(m, L.m, sizeof(L.m));
}

void ruf_view_matrix(const double *offset_q, const double *offset_t, const double *cam_q,
                     const double *cam_t, double camera_tx, double camera_ty, double *view)
{
  Mat4 mv;
  ruf_lookat(mv.m);
  mv = mul(mv, Rigid::from_quat(offset_q, offset_t).inverse().gl());
  Rigid cam = Rigid::from_quat(cam_q, cam_t);
  // origin += (R * e_x) * tx ; origin += (R * e_y) * ty
  for (int i = 0; i < 3; ++i) cam.t[i] = cam.t[i] + cam.R[i][0] * camera_tx;
  for (int i = 0; i < 3; ++i) cam.t[i] = cam.t[i] + cam.R[i][1] * camera_ty;
  mv = mul(mv, cam.gl());
  std::memcpy(view, mv.m, sizeof(mv.m));
}

void ruf_part_model(const double *link_q, const double *link_t, const double *off_q,
                    const double *off_t, const double *suffix, double *model)
{
  const double n = std::sqrt(off_q[0] * off_q[0] + off_q[1] * off_q[1] + off_q[2] * off_q[2] + off_q[3] * off_q[3]);
  const double qn[4] = {off_q[0] / n, off_q[1] / n, off_q[2] / n, off_q[3] / n};
  Mat4 M = (Rigid::from_quat(link_q, link_t) * Rigid::from_quat(qn, off_t)).gl();
  if (suffix) {
    Mat4 S;
    std::memcpy(S.m, suffix, sizeof(S.m));
    M = mul(M, S);
  }
  std::memcpy(model, M.m, sizeof(M.m));
}

int ruf_box_triangles(float dimx, float dimy, float dimz, float *out)
{
  const float X = 0.5f * dimx, Y = 0.5f * dimy, Z = 0.5f * dimz;
  // face order and winding of the 24-vertex GL_QUADS buffer: top, bottom, front, back, left, right
  const float v[24][3] = {
      {X, Y, -Z},  {-X, Y, -Z},  {-X, Y, Z},   {X, Y, Z},    {X, -Y, Z},  {-X, -Y, Z},
      {-X, -Y, -Z}, {X, -Y, -Z}, {X, Y, Z},    {-X, Y, Z},   {-X, -Y, Z}, {X, -Y, Z},
      {X, -Y, -Z}, {-X, -Y, -Z}, {-X, Y, -Z},  {X, Y, -Z},   {-X, Y, Z},  {-X, Y, -Z},
      {-X, -Y, -Z}, {-X, -Y, Z}, {X, Y, -Z},   {X, Y, Z},    {X, -Y, Z},  {X, -Y, -Z}};
  Soup s{out};
  for (int f = 0; f < 6; ++f) s.quad(v[4 * f], v[4 * f + 1], v[4 * f + 2], v[4 * f + 3]);
  return s.n;
}

int ruf_cube_triangles(float size, float *out)
{
  const float h = size * 0.5f;
  const float v[8][3] = {{h, h, h}, {-h, h, h}, {-h, -h, h}, {h, -h, h}, {h, h, -h}, {-h, h, -h}, {-h, -h, -h}, {h, -h, -h}};
  static const int face[6][4] = {{0, 3, 7, 4}, {1, 0, 4, 5}, {0, 1, 2, 3}, {2, 1, 5, 6}, {3, 2, 6, 7}, {7, 6, 5, 4}};
  Soup s{out};
  for (const auto &f : face) s.quad(v[f[0]], v[f[1]], v[f[2]], v[f[3]]);
  return s.n;
}

int ruf_sphere_triangle_count(int slices, int stacks) { return 2 * slices + 2 * slices * (stacks - 2); }
int ruf_cylinder_triangle_count(int slices, int stacks) { return 2 * slices + 2 * slices * stacks; }

int ruf_sphere_triangles(float radius, int slices, int stacks, float *out)
{
  if (slices < 1 || stacks < 2) return 0;
  std::vector<double> s1, c1, s2, c2;
  circle_table(s1, c1, -slices);
  circle_table(s2, c2, stacks * 2);
  Soup s{out};
  const double r = radius;
  for (int i = 0; i < stacks; ++i) {
    double z0 = c2[i], r0 = s2[i], z1 = c2[i + 1], r1 = s2[i + 1];
    if (i == 0) { r0 = 0.0; z0 = 1.0; }
    if (i == stacks - 1) { r1 = 0.0; z1 = -1.0; }
    for (int j = 0; j < slices; ++j) {
      const float a[3] = {(float)(c1[j] * r0 * r), (float)(s1[j] * r0 * r), (float)(z0 * r)};
      const float b[3] = {(float)(c1[j] * r1 * r), (float)(s1[j] * r1 * r), (float)(z1 * r)};
      const float c[3] = {(float)(c1[j + 1] * r1 * r), (float)(s1[j + 1] * r1 * r), (float)(z1 * r)};
      const float d[3] = {(float)(c1[j + 1] * r0 * r), (float)(s1[j + 1] * r0 * r), (float)(z0 * r)};
      if (i == 0) s.tri(a, b, c);
      else if (i == stacks - 1) s.tri(a, b, d);
      else s.quad(a, b, c, d);
    }
  }
  return s.n;
}

int ruf_cylinder_triangles(float radius, float height, int slices, int stacks, float *out)
{
  if (slices < 1 || stacks < 1) return 0;
  std::vector<double> st, ct;
  circle_table(st, ct, -slices);
  Soup s{out};
  const double r = radius, zstep = (double)height / stacks;
  const float base[3] = {0.f, 0.f, 0.f}, top[3] = {0.f, 0.f, height};
  for (int j = 0; j < slices; ++j) {
    const float a[3] = {(float)(ct[j] * r), (float)(st[j] * r), 0.f};
    const float b[3] = {(float)(ct[j + 1] * r), (float)(st[j + 1] * r), 0.f};
    s.tri(base, b, a);
  }
  for (int j = 0; j < slices; ++j) {
    const float a[3] = {(float)(ct[j] * r), (float)(st[j] * r), height};
    const float b[3] = {(float)(ct[j + 1] * r), (float)(st[j + 1] * r), height};
    s.tri(top, a, b);
  }
  for (int i = 0; i < stacks; ++i) {
    const double z0 = zstep * i, z1 = (i == stacks - 1) ? (double)height : zstep * (i + 1);
    for (int j = 0; j < slices; ++j) {
      const float a[3] = {(float)(ct[j] * r), (float)(st[j] * r), (float)z0};
      const float b[3] = {(float)(ct[j] * r), (float)(st[j] * r), (float)z1};
      const float c[3] = {(float)(ct[j + 1] * r), (float)(st[j + 1] * r), (float)z1};
      const float d[3] = {(float)(ct[j + 1] * r), (float)(st[j + 1] * r), (float)z0};
      s.quad(a, b, c, d);
    }
  }
  return s.n;
}

const char *ruf_version(void) { return "ruf_b200 0.1 (sm_100a)"; }

}  // extern "C"
