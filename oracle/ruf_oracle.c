/*
 * ruf_oracle.c -- CPU ORACLE for the URDF depth self-filter hot path.
 *
 * THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, the smoke
 * check in __graft_entry__.py and the cpu_baseline / --impl reference legs of
 * bench.py may build, load or call it.  The product library (libruf_b200.so)
 * never links against it and has no CPU fallback.
 *
 * It is a plain-C restatement of the per-frame path of blodow/realtime_urdf_filter
 * (paths below are relative to the reference checkout):
 *
 *   src/urdf_filter.cpp:459-501   getProjectionMatrix      -> orc_projection_matrix
 *   src/urdf_filter.cpp:587       gluLookAt(0,0,0,0,0,1,0,1,0) -> orc_lookat
 *   src/urdf_filter.cpp:602-614   camera offset^-1, camera TF (+tx/ty shift) -> orc_view_matrix
 *   src/renderable.cpp:59-68      applyTransform (link_to_fixed * link_offset) -> orc_link_model
 *   src/renderable.cpp:95,128,427 glTranslatef / glScalef suffixes -> `suffix` of orc_link_model
 *   src/renderable.cpp:135-164    box vertex table         -> orc_box_triangles
 *   src/renderable.cpp:83,96,129  glutSolidSphere/Cylinder/Cube (freeglut) -> orc_sphere/cylinder/cube_triangles
 *   include/shaders/urdf_filter.vert:5   gl_Position = MVP * v   -> xform()
 *   src/urdf_filter.cpp:570,591-596      GL_DEPTH_TEST (GL_LESS), background quad -> orc_render
 *   include/shaders/urdf_filter.frag:14-35  linearise / compare / mix -> orc_filter
 *   src/urdf_filter.cpp:287-288,311      16UC1 <-> 32FC1 convertTo -> orc_u16_to_f32 / orc_f32_to_u16
 *   src/urdf_filter.cpp:729-735          readback of attachment 1 (.r) and 3 (.r as 0/255)
 *   src/urdf_renderer.cpp:173-190 (per-frame tf lookups) -> orc_fk / orc_fk_outputs: forward kinematics
 *                                        from joint positions (robot_state_publisher + tf, restated)
 *
 * PARITY STATUS: pinned against the reference itself.  The reference has no tests,
 * golden images or fixtures and its C++/ROS program cannot be
 * built in this image, but its GLSL path can be run: oracle/gl_ref replays the GL
 * call sequence of RealtimeURDFFilter::render with the reference's two shader files
 * UNMODIFIED in a real GL driver (the Mesa 18.1.9 llvmpipe libGL inside Nsight
 * Compute, on a 30-call fake Xlib), and this file reproduces that driver's filtered
 * depth and mask bit for bit on every synthetic scene (0 - 3 mask pixels of 1.2 M on
 * the 1280x960 frames) and within 9 mask pixels per image on hostile random soups
 * (tests/test_gl_crosscheck.py here; golden vectors from the driver in
 * tests/golden/gl_llvmpipe.npz, tests/test_gl_golden.py everywhere).  Also pinned: the
 * encodings against OpenCV's real Mat::convertTo (tests/golden/cv_convert.npz), the
 * matrices and shader by analytic known-answer tests, the geometry by an independent
 * float64 ray-caster in tests/.  Not measured: hardware GL drivers.
 *
 * Third-party arithmetic restated here from its published algorithm (absent from
 * /root/reference, all unpinned in package.xml):
 *   tf / Bullet LinearMath : Matrix3x3::setRotation, Transform::inverse, operator*,
 *                            getOpenGLMatrix  (package.xml:15-16)
 *   GLU                    : gluLookAt
 *   freeglut               : glutSolidSphere / glutSolidCylinder / glutSolidCube
 *                            (package.xml:29-30)
 *   OpenCV                 : Mat::convertTo 16U->32F (x*0.001f) and 32F->16U
 *                            (cvRound(x*1000.f), saturate)  (package.xml:21-22)
 *
 * Build: gcc -O2 -ffp-contract=off (see oracle/Makefile).  Every float operation
 * whose order matters is written out; fused multiply-adds appear only as explicit
 * fmaf() calls.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ---- raster specification constants (DESIGN.md "Raster specification") ---- */
#define SUBPIX_BITS 8
#define SUBPIX 256          /* 1 << SUBPIX_BITS                                   */
#define SUBPIX_HALF 128     /* pixel centre offset                                */
#define GUARD_PX 6144.0f    /* guard band half-width in pixels around the centre  */
#define MAX_POLY 12

#define ENC_F32_M 0         /* 32FC1, metres, NaN = invalid  */
#define ENC_U16_MM 1        /* 16UC1, millimetres, 0 = invalid */

/* ========================================================================= */
/* 1. Host-side matrices (double, column-major like OpenGL)                   */
/* ========================================================================= */

/* C = A * B, 4x4 column-major doubles; plain sum in index order. */
static void mat4_mul(const double *A, const double *B, double *C)
{
  double T[16];
  for (int c = 0; c < 4; ++c)
    for (int r = 0; r < 4; ++r) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k)
        s += A[k * 4 + r] * B[c * 4 + k];
      T[c * 4 + r] = s;
    }
  memcpy(C, T, sizeof(T));
}

static void mat4_identity(double *M)
{
  memset(M, 0, 16 * sizeof(double));
  M[0] = M[5] = M[10] = M[15] = 1.0;
}

/* src/urdf_filter.cpp:459-501.  P is the 3x4 row-major CameraInfo.P. */
ORC_API void orc_projection_matrix(const double *P, int width, int height,
                                   double near_plane, double far_plane,
                                   double *glTf, double *camera_tx, double *camera_ty)
{
  double fx = P[0], fy = P[5], cx = P[2], cy = P[6];
  if (camera_tx) *camera_tx = -1 * (P[3] / fx);              /* :480 */
  if (camera_ty) *camera_ty = -1 * (P[7] / fy);              /* :481 */
  for (int i = 0; i < 16; ++i) glTf[i] = 0.0;                /* :485-487 */
  glTf[0] = -2.0 * fx / width;                               /* :491 */
  glTf[5] = 2.0 * fy / height;                               /* :492 */
  glTf[8] = 2.0 * (0.5 - cx / width);                        /* :494 */
  glTf[9] = 2.0 * (cy / height - 0.5);                       /* :495 */
  glTf[10] = -(far_plane + near_plane) / (far_plane - near_plane);        /* :497 */
  glTf[14] = -2.0 * far_plane * near_plane / (far_plane - near_plane);    /* :498 */
  glTf[11] = -1;                                             /* :500 */
}

/* gluLookAt(eye, center, up) as published in the GLU specification / Mesa libGLU:
 * f = normalize(center-eye); s = f x up (normalised); u = s x f;
 * M = rows (s, u, -f); then translate(-eye).  src/urdf_filter.cpp:587 calls it with
 * (0,0,0, 0,0,1, 0,1,0). */
static void glu_lookat(double ex, double ey, double ez, double cx, double cy, double cz,
                       double ux, double uy, double uz, double *M)
{
  double f[3] = {cx - ex, cy - ey, cz - ez};
  double n = sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
  f[0] /= n; f[1] /= n; f[2] /= n;
  double s[3] = {f[1] * uz - f[2] * uy, f[2] * ux - f[0] * uz, f[0] * uy - f[1] * ux};
  n = sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
  s[0] /= n; s[1] /= n; s[2] /= n;
  double u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};
  double R[16];
  mat4_identity(R);
  R[0] = s[0]; R[4] = s[1]; R[8] = s[2];
  R[1] = u[0]; R[5] = u[1]; R[9] = u[2];
  R[2] = -f[0]; R[6] = -f[1]; R[10] = -f[2];
  double T[16];
  mat4_identity(T);
  T[12] = -ex; T[13] = -ey; T[14] = -ez;
  mat4_mul(R, T, M);
}

ORC_API void orc_lookat(double *M) { glu_lookat(0, 0, 0, 0, 0, 1, 0, 1, 0, M); }

/* tf::Transform = (Matrix3x3 basis (row-major rows), Vector3 origin). */
typedef struct { double b[3][3]; double o[3]; } tfx;

/* Bullet/tf Matrix3x3::setRotation(q), q = (x,y,z,w). */
static void tfx_from_qt(const double *q, const double *t, tfx *X)
{
  double x = q[0], y = q[1], z = q[2], w = q[3];
  double d = x * x + y * y + z * z + w * w;
  double s = 2.0 / d;
  double xs = x * s, ys = y * s, zs = z * s;
  double wx = w * xs, wy = w * ys, wz = w * zs;
  double xx = x * xs, xy = x * ys, xz = x * zs;
  double yy = y * ys, yz = y * zs, zz = z * zs;
  X->b[0][0] = 1.0 - (yy + zz); X->b[0][1] = xy - wz;         X->b[0][2] = xz + wy;
  X->b[1][0] = xy + wz;         X->b[1][1] = 1.0 - (xx + zz); X->b[1][2] = yz - wx;
  X->b[2][0] = xz - wy;         X->b[2][1] = yz + wx;         X->b[2][2] = 1.0 - (xx + yy);
  X->o[0] = t[0]; X->o[1] = t[1]; X->o[2] = t[2];
}

/* tf::Transform::operator* : basis = b1*b2, origin = b1*o2 + o1. */
static void tfx_mul(const tfx *A, const tfx *B, tfx *C)
{
  tfx R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      R.b[i][j] = A->b[i][0] * B->b[0][j] + A->b[i][1] * B->b[1][j] + A->b[i][2] * B->b[2][j];
  for (int i = 0; i < 3; ++i)
    R.o[i] = (A->b[i][0] * B->o[0] + A->b[i][1] * B->o[1] + A->b[i][2] * B->o[2]) + A->o[i];
  *C = R;
}

/* tf::Transform::inverse : inv = basis^T ; origin = inv * (-origin). */
static void tfx_inverse(const tfx *A, tfx *C)
{
  tfx R;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      R.b[i][j] = A->b[j][i];
  for (int i = 0; i < 3; ++i)
    R.o[i] = R.b[i][0] * -A->o[0] + R.b[i][1] * -A->o[1] + R.b[i][2] * -A->o[2];
  *C = R;
}

/* tf::Transform::getOpenGLMatrix : column-major, bottom row (0,0,0,1). */
static void tfx_to_gl(const tfx *A, double *m)
{
  m[0] = A->b[0][0]; m[1] = A->b[1][0]; m[2] = A->b[2][0];  m[3] = 0.0;
  m[4] = A->b[0][1]; m[5] = A->b[1][1]; m[6] = A->b[2][1];  m[7] = 0.0;
  m[8] = A->b[0][2]; m[9] = A->b[1][2]; m[10] = A->b[2][2]; m[11] = 0.0;
  m[12] = A->o[0];   m[13] = A->o[1];   m[14] = A->o[2];    m[15] = 1.0;
}

ORC_API void orc_transform_to_gl(const double *q, const double *t, double *m)
{
  tfx X;
  tfx_from_qt(q, t, &X);
  tfx_to_gl(&X, m);
}

/* MODELVIEW before any link is drawn: LookAt * offset^-1 * camera_transform',
 * src/urdf_filter.cpp:583-614.  cam_q/cam_t is what
 * tf_.lookupTransform(cam_frame_, fixed_frame_) returned. */
ORC_API void orc_view_matrix(const double *offset_q, const double *offset_t,
                             const double *cam_q, const double *cam_t,
                             double camera_tx, double camera_ty, double *view)
{
  double MV[16], G[16];
  orc_lookat(MV);                                           /* :587 */
  tfx off, offinv;
  tfx_from_qt(offset_q, offset_t, &off);                    /* :602 */
  tfx_inverse(&off, &offinv);                               /* :603 */
  tfx_to_gl(&offinv, G);
  mat4_mul(MV, G, MV);                                      /* :604 */

  tfx cam;
  tfx_from_qt(cam_q, cam_t, &cam);
  /* :607-610 -- right/down are the first two basis columns of the camera rotation */
  double right[3] = {cam.b[0][0], cam.b[1][0], cam.b[2][0]};
  for (int i = 0; i < 3; ++i) cam.o[i] = cam.o[i] + right[i] * camera_tx;
  double down[3] = {cam.b[0][1], cam.b[1][1], cam.b[2][1]};
  for (int i = 0; i < 3; ++i) cam.o[i] = cam.o[i] + down[i] * camera_ty;
  tfx_to_gl(&cam, G);                                       /* :613 */
  mat4_mul(MV, G, view);                                    /* :614 */
}

/* Model matrix of one drawn part: link_to_fixed * link_offset [* suffix],
 * src/renderable.cpp:59-68; suffix = glTranslatef / glScalef issued by the
 * renderable's render() before its draw call (may be NULL). */
ORC_API void orc_link_model(const double *link_q, const double *link_t,
                            const double *off_q, const double *off_t,
                            const double *suffix, double *model)
{
  tfx l2f, off, M;
  tfx_from_qt(link_q, link_t, &l2f);
  /* src/urdf_renderer.cpp:162-164 normalises the URDF origin quaternion */
  double n = sqrt(off_q[0] * off_q[0] + off_q[1] * off_q[1] + off_q[2] * off_q[2] + off_q[3] * off_q[3]);
  double qn[4] = {off_q[0] / n, off_q[1] / n, off_q[2] / n, off_q[3] / n};
  tfx_from_qt(qn, off_t, &off);
  tfx_mul(&l2f, &off, &M);                                  /* renderable.cpp:63-64 */
  tfx_to_gl(&M, model);
  if (suffix) mat4_mul(model, suffix, model);
}

/* MVP of every part, as the vertex shader sees it (gl_ModelViewProjectionMatrix),
 * composed in double and rounded once to float32.  Entry L is the background
 * quad's MVP = P * LookAt (src/urdf_filter.cpp:576-596). */
ORC_API void orc_compose_mvp(const double *proj, const double *view,
                             const double *link_model, int L, float *mvp)
{
  double PV[16], M[16];
  mat4_mul(proj, view, PV);
  for (int l = 0; l < L; ++l) {
    mat4_mul(PV, link_model + 16 * (size_t)l, M);
    for (int i = 0; i < 16; ++i) mvp[16 * (size_t)l + i] = (float)M[i];
  }
  double LA[16];
  orc_lookat(LA);
  mat4_mul(proj, LA, M);
  for (int i = 0; i < 16; ++i) mvp[16 * (size_t)L + i] = (float)M[i];
}

/* ========================================================================= */
/* 1b. Forward kinematics (the "next" row f1 of SURVEY.md section 8: replaces the   */
/*     per-frame tf lookups of src/urdf_renderer.cpp:173-190 and urdf_filter.cpp:522) */
/* ========================================================================= */

/* sin and cos with a FIXED operation order (no libm, no fma) so that the CUDA kernel and this oracle
 * give the same bits: Cody-Waite reduction by pi/2 (3 constants), then Taylor polynomials of degree
 * 17 / 16 on [-pi/4, pi/4] in Horner form.  |error| < 3e-16 for |x| < 1e5. */
ORC_API void orc_sincos(double x, double *sn, double *cs)
{
  const double two_over_pi = 0.63661977236758134308;
  const double p1 = 1.57079632673412561417e+00;   /* pi/2 split in three parts (fdlibm) */
  const double p2 = 6.07710050650619224932e-11;
  const double p3 = 2.02226624879595063154e-21;
  double kf = x * two_over_pi;
  kf = (kf >= 0.0) ? floor(kf + 0.5) : -floor(-kf + 0.5);
  double r = x - kf * p1;
  r = r - kf * p2;
  r = r - kf * p3;
  const double z = r * r;
  /* sin(r) = r + r*z*(S1 + z*(S2 + ... )) */
  double ps = 2.81145725434552076320e-15;            /* 1/17! */
  ps = ps * z + -7.64716373181981647590e-13;         /* -1/15! */
  ps = ps * z + 1.60590438368216145994e-10;          /* 1/13! */
  ps = ps * z + -2.50521083854417187751e-08;         /* -1/11! */
  ps = ps * z + 2.75573192239858906526e-06;          /* 1/9! */
  ps = ps * z + -1.98412698412698412698e-04;         /* -1/7! */
  ps = ps * z + 8.33333333333333333333e-03;          /* 1/5! */
  ps = ps * z + -1.66666666666666666667e-01;         /* -1/3! */
  const double sr = r + (r * z) * ps;
  double pc = 4.77947733238738529744e-14;            /* 1/16! */
  pc = pc * z + -1.14707455977297247139e-11;         /* -1/14! */
  pc = pc * z + 2.08767569878680989792e-09;          /* 1/12! */
  pc = pc * z + -2.75573192239858906526e-07;         /* -1/10! */
  pc = pc * z + 2.48015873015873015873e-05;          /* 1/8! */
  pc = pc * z + -1.38888888888888888889e-03;         /* -1/6! */
  pc = pc * z + 4.16666666666666666667e-02;          /* 1/4! */
  pc = pc * z + -0.5;                                /* -1/2! */
  const double cr = 1.0 + z * pc;
  const long long k = (long long)kf;
  switch ((int)(k & 3)) {
    case 0: *sn = sr; *cs = cr; break;
    case 1: *sn = cr; *cs = -sr; break;
    case 2: *sn = -sr; *cs = -cr; break;
    default: *sn = -cr; *cs = sr; break;
  }
}

/* joint motion as a column-major 4x4: revolute = rotation by q about the unit axis (quaternion
 * (axis*sin(q/2), cos(q/2)) -> tf::Matrix3x3::setRotation), prismatic = translation by axis*q */
static void joint_motion(int type, const double *axis, double q, double *M)
{
  mat4_identity(M);
  if (type == 1) {
    double sn, cs;
    orc_sincos(q * 0.5, &sn, &cs);
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

/* link_to_fixed of every link: T_link = (T_parent * origin) * motion(q), parents first.
 *   parent[l] < l or -1; type 0 fixed / 1 revolute / 2 prismatic; origin n*16; axis n*3 (unit); q n
 *   out: n*16 column-major */
ORC_API void orc_fk(int n_links, const int32_t *parent, const int32_t *type, const double *origin,
                    const double *axis, const double *q, double *out)
{
  for (int l = 0; l < n_links; ++l) {
    double T[16], M[16];
    if (parent[l] < 0) memcpy(T, origin + 16 * (size_t)l, sizeof(T));
    else mat4_mul(out + 16 * (size_t)parent[l], origin + 16 * (size_t)l, T);
    if (type[l] != 0) {
      joint_motion(type[l], axis + 3 * (size_t)l, q[l], M);
      mat4_mul(T, M, T);
    }
    memcpy(out + 16 * (size_t)l, T, sizeof(T));
  }
}

/* part models and the view matrix of one frame from the link poses:
 *   part_model[p] = T_link[part_link[p]] * part_local[p]
 *   view = view_pre * inverse_rigid(T_link[cam_link] * cam_mount) with the origin shifted by
 *          tx along the camera's x axis and ty along its y axis (src/urdf_filter.cpp:607-610);
 *          view_pre = LookAt * inverse(camera_offset) (host); cam_link = -1: the fixed frame */
ORC_API void orc_fk_outputs(const double *links, int n_parts, const int32_t *part_link, const double *part_local,
                            int cam_link, const double *cam_mount, const double *view_pre, double tx, double ty,
                            double *part_model, double *view)
{
  for (int p = 0; p < n_parts; ++p)
    mat4_mul(links + 16 * (size_t)part_link[p], part_local + 16 * (size_t)p, part_model + 16 * (size_t)p);
  double C[16];
  if (cam_link < 0) memcpy(C, cam_mount, sizeof(C));
  else mat4_mul(links + 16 * (size_t)cam_link, cam_mount, C);
  /* tf::Transform::inverse: R^T, R^T * (-t); then the tx/ty shift along the inverse's basis columns */
  double I[16];
  mat4_identity(I);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) I[c * 4 + r] = C[r * 4 + c];
  for (int r = 0; r < 3; ++r)
    I[12 + r] = I[0 * 4 + r] * -C[12] + I[1 * 4 + r] * -C[13] + I[2 * 4 + r] * -C[14];
  for (int r = 0; r < 3; ++r) I[12 + r] = I[12 + r] + I[0 * 4 + r] * tx;
  for (int r = 0; r < 3; ++r) I[12 + r] = I[12 + r] + I[1 * 4 + r] * ty;
  mat4_mul(view_pre, I, view);
}

/* ========================================================================= */
/* 2. Geometry generators (triangle soup, 9 floats per triangle)              */
/* ========================================================================= */

static float *emit_tri(float *o, const float *a, const float *b, const float *c)
{
  memcpy(o, a, 12); memcpy(o + 3, b, 12); memcpy(o + 6, c, 12);
  return o + 9;
}
static float *emit_quad(float *o, const float *a, const float *b, const float *c, const float *d)
{
  o = emit_tri(o, a, b, c);
  return emit_tri(o, a, c, d);
}

/* The 24-vertex GL_QUADS box of src/renderable.cpp:135-164 (positions only). */
ORC_API int orc_box_triangles(float dimx, float dimy, float dimz, float *out)
{
  const float X = 0.5f * dimx, Y = 0.5f * dimy, Z = 0.5f * dimz;
  const float v[24][3] = {
    { X,  Y, -Z}, {-X,  Y, -Z}, {-X,  Y,  Z}, { X,  Y,  Z},   /* top    */
    { X, -Y,  Z}, {-X, -Y,  Z}, {-X, -Y, -Z}, { X, -Y, -Z},   /* bottom */
    { X,  Y,  Z}, {-X,  Y,  Z}, {-X, -Y,  Z}, { X, -Y,  Z},   /* front  */
    { X, -Y, -Z}, {-X, -Y, -Z}, {-X,  Y, -Z}, { X,  Y, -Z},   /* back   */
    {-X,  Y,  Z}, {-X,  Y, -Z}, {-X, -Y, -Z}, {-X, -Y,  Z},   /* left   */
    { X,  Y, -Z}, { X,  Y,  Z}, { X, -Y,  Z}, { X, -Y, -Z}};  /* right  */
  float *o = out;
  for (int f = 0; f < 6; ++f)
    o = emit_quad(o, v[4 * f], v[4 * f + 1], v[4 * f + 2], v[4 * f + 3]);
  return 12;
}

/* freeglut glutSolidCube(size): six quads at +-size/2. */
ORC_API int orc_cube_triangles(float size, float *out)
{
  const float s = size * 0.5f;
  const float v[8][3] = {{ s,  s,  s}, {-s,  s,  s}, {-s, -s,  s}, { s, -s,  s},
                         { s,  s, -s}, {-s,  s, -s}, {-s, -s, -s}, { s, -s, -s}};
  /* freeglut face list: +x, +y, +z, -x, -y, -z */
  static const int f[6][4] = {{0, 3, 7, 4}, {1, 0, 4, 5}, {0, 1, 2, 3},
                              {2, 1, 5, 6}, {3, 2, 6, 7}, {7, 6, 5, 4}};
  float *o = out;
  for (int i = 0; i < 6; ++i)
    o = emit_quad(o, v[f[i][0]], v[f[i][1]], v[f[i][2]], v[f[i][3]]);
  return 12;
}

/* freeglut fghCircleTable: n+1 entries of sin/cos(2*pi*i/n); n<0 reverses direction. */
static void circle_table(double *sint, double *cost, int n)
{
  const int size = abs(n);
  const double angle = 2.0 * M_PI / (double)((n == 0) ? 1 : n);
  sint[0] = 0.0; cost[0] = 1.0;
  for (int i = 1; i < size; ++i) { sint[i] = sin(angle * i); cost[i] = cos(angle * i); }
  sint[size] = sint[0]; cost[size] = cost[0];
}

/* freeglut glutSolidSphere(radius, slices, stacks): poles at +-z, rings at polar angle
 * pi*i/stacks, ring points from fghCircleTable(-slices).  2*slices + 2*slices*(stacks-2)
 * triangles (180 for 10,10). */
ORC_API int orc_sphere_triangles(float radius, int slices, int stacks, float *out)
{
  double *s1 = malloc(sizeof(double) * (slices + 1)), *c1 = malloc(sizeof(double) * (slices + 1));
  double *s2 = malloc(sizeof(double) * (2 * stacks + 1)), *c2 = malloc(sizeof(double) * (2 * stacks + 1));
  circle_table(s1, c1, -slices);
  circle_table(s2, c2, stacks * 2);
  float *o = out;
  const double r = radius;
  for (int i = 0; i < stacks; ++i) {
    double z0 = c2[i], r0 = s2[i], z1 = c2[i + 1], r1 = s2[i + 1];
    if (i == 0) { r0 = 0.0; z0 = 1.0; }
    if (i == stacks - 1) { r1 = 0.0; z1 = -1.0; }
    for (int j = 0; j < slices; ++j) {
      float a[3] = {(float)(c1[j] * r0 * r), (float)(s1[j] * r0 * r), (float)(z0 * r)};
      float b[3] = {(float)(c1[j] * r1 * r), (float)(s1[j] * r1 * r), (float)(z1 * r)};
      float c[3] = {(float)(c1[j + 1] * r1 * r), (float)(s1[j + 1] * r1 * r), (float)(z1 * r)};
      float d[3] = {(float)(c1[j + 1] * r0 * r), (float)(s1[j + 1] * r0 * r), (float)(z0 * r)};
      if (i == 0) o = emit_tri(o, a, b, c);
      else if (i == stacks - 1) o = emit_tri(o, a, b, d);
      else o = emit_quad(o, a, b, c, d);
    }
  }
  free(s1); free(c1); free(s2); free(c2);
  return (int)((o - out) / 9);
}

/* freeglut glutSolidCylinder(radius, height, slices, stacks): base fan at z=0, top fan
 * at z=height, `stacks` bands of quads between.  2*slices + 2*slices*stacks triangles
 * (220 for 10,10).  The reference translates by -length/2 first
 * (src/renderable.cpp:95) -- that is a matrix suffix, not baked in here. */
ORC_API int orc_cylinder_triangles(float radius, float height, int slices, int stacks, float *out)
{
  double *st = malloc(sizeof(double) * (slices + 1)), *ct = malloc(sizeof(double) * (slices + 1));
  circle_table(st, ct, -slices);
  float *o = out;
  const double r = radius, zstep = (double)height / ((stacks > 0) ? stacks : 1);
  const float c0[3] = {0.f, 0.f, 0.f}, c1[3] = {0.f, 0.f, height};
  for (int j = 0; j < slices; ++j) {               /* base */
    float a[3] = {(float)(ct[j] * r), (float)(st[j] * r), 0.f};
    float b[3] = {(float)(ct[j + 1] * r), (float)(st[j + 1] * r), 0.f};
    o = emit_tri(o, c0, b, a);
  }
  for (int j = 0; j < slices; ++j) {               /* top */
    float a[3] = {(float)(ct[j] * r), (float)(st[j] * r), height};
    float b[3] = {(float)(ct[j + 1] * r), (float)(st[j + 1] * r), height};
    o = emit_tri(o, c1, a, b);
  }
  for (int i = 0; i < stacks; ++i) {               /* body */
    double z0 = zstep * i, z1 = (i == stacks - 1) ? (double)height : zstep * (i + 1);
    for (int j = 0; j < slices; ++j) {
      float a[3] = {(float)(ct[j] * r), (float)(st[j] * r), (float)z0};
      float b[3] = {(float)(ct[j] * r), (float)(st[j] * r), (float)z1};
      float c[3] = {(float)(ct[j + 1] * r), (float)(st[j + 1] * r), (float)z1};
      float d[3] = {(float)(ct[j + 1] * r), (float)(st[j + 1] * r), (float)z0};
      o = emit_quad(o, a, b, c, d);
    }
  }
  free(st); free(ct);
  return (int)((o - out) / 9);
}

/* ========================================================================= */
/* 3. Vertex stage, clipping, rasterisation  (DESIGN.md "Raster specification") */
/* ========================================================================= */

typedef struct { float x, y, z, w; } vec4;

/* S1: clip = MVP * (x,y,z,1), float32, fixed order. include/shaders/urdf_filter.vert:5 */
static vec4 xform(const float *m, const float *p)
{
  vec4 c;
  c.x = fmaf(m[0], p[0], fmaf(m[4], p[1], fmaf(m[8], p[2], m[12])));
  c.y = fmaf(m[1], p[0], fmaf(m[5], p[1], fmaf(m[9], p[2], m[13])));
  c.z = fmaf(m[2], p[0], fmaf(m[6], p[1], fmaf(m[10], p[2], m[14])));
  c.w = fmaf(m[3], p[0], fmaf(m[7], p[1], fmaf(m[11], p[2], m[15])));
  return c;
}

static int finite4(vec4 c) { return isfinite(c.x) && isfinite(c.y) && isfinite(c.z) && isfinite(c.w); }

/* S3: signed distance to clip plane k.  0 = near (z >= -w), 1..4 = guard band. */
static float plane_dist(vec4 c, int k, float gx, float gy)
{
  switch (k) {
    case 0: return c.z + c.w;
    case 1: return fmaf(gx, c.w, -c.x);
    case 2: return fmaf(gx, c.w, c.x);
    case 3: return fmaf(gy, c.w, -c.y);
    default: return fmaf(gy, c.w, c.y);
  }
}

/* new vertex on the plane, always interpolated from the inside vertex towards the
 * outside one so that both triangles sharing an edge create the same point. */
static vec4 clip_lerp(vec4 in, vec4 out, float din, float dout)
{
  float t = din / (din - dout);
  vec4 r;
  r.x = fmaf(t, out.x - in.x, in.x);
  r.y = fmaf(t, out.y - in.y, in.y);
  r.z = fmaf(t, out.z - in.z, in.z);
  r.w = fmaf(t, out.w - in.w, in.w);
  return r;
}

static int clip_poly(vec4 *poly, int n, float gx, float gy)
{
  vec4 tmp[MAX_POLY];
  for (int k = 0; k < 5 && n > 0; ++k) {
    int m = 0;
    for (int i = 0; i < n; ++i) {
      vec4 a = poly[i], b = poly[(i + 1) % n];
      float da = plane_dist(a, k, gx, gy), db = plane_dist(b, k, gx, gy);
      int ia = da >= 0.0f, ib = db >= 0.0f;
      if (ia) tmp[m++] = a;
      if (ia != ib) tmp[m++] = ia ? clip_lerp(a, b, da, db) : clip_lerp(b, a, db, da);
    }
    n = m;
    memcpy(poly, tmp, sizeof(vec4) * (size_t)n);
  }
  return n;
}

typedef struct { int32_t X, Y; float z; } wvert;

/* S4: perspective divide, viewport (glViewport(0,0,W,H), depth range [0,1]), snap.
 * Returns 0 when the vertex has no usable window position (w so small that the divide
 * overflowed): |xw|,|yw| must stay inside 16384 px (the guard band plus half the largest
 * viewport is 8192) and z must be finite.  The caller then drops the primitive. */
#define WINDOW_LIMIT 16384.0f
static int to_window(vec4 c, float halfw, float halfh, wvert *v)
{
  float iw = 1.0f / c.w;
  float nx = c.x * iw, ny = c.y * iw, nz = c.z * iw;
  float xw = fmaf(nx, halfw, halfw);
  float yw = fmaf(ny, halfh, halfh);
  float zw = fmaf(nz, 0.5f, 0.5f);
  if (!(fabsf(xw) <= WINDOW_LIMIT) || !(fabsf(yw) <= WINDOW_LIMIT) || !(fabsf(zw) <= WINDOW_LIMIT))
    return 0;
  v->z = zw;
  v->X = (int32_t)lrintf(xw * (float)SUBPIX);
  v->Y = (int32_t)lrintf(yw * (float)SUBPIX);
  return 1;
}

static inline int64_t floor_shift(int64_t a) { return a >> SUBPIX_BITS; } /* arithmetic shift */

/* S5-S9: rasterise one window-space triangle into zbuf rows [row0,row1). */
static void raster_tri(wvert v0, wvert v1, wvert v2, float *zbuf, int W, int H, int row0, int row1)
{
  int64_t area2 = (int64_t)(v1.X - v0.X) * (v2.Y - v0.Y) - (int64_t)(v2.X - v0.X) * (v1.Y - v0.Y);
  (void)H;
  if (area2 == 0) return;                                   /* degenerate */
  if (area2 < 0) { wvert t = v1; v1 = v2; v2 = t; area2 = -area2; }   /* no culling */

  int32_t xmin = v0.X < v1.X ? v0.X : v1.X; if (v2.X < xmin) xmin = v2.X;
  int32_t xmax = v0.X > v1.X ? v0.X : v1.X; if (v2.X > xmax) xmax = v2.X;
  int32_t ymin = v0.Y < v1.Y ? v0.Y : v1.Y; if (v2.Y < ymin) ymin = v2.Y;
  int32_t ymax = v0.Y > v1.Y ? v0.Y : v1.Y; if (v2.Y > ymax) ymax = v2.Y;
  int64_t i0 = floor_shift((int64_t)xmin - SUBPIX_HALF + (SUBPIX - 1));
  int64_t i1 = floor_shift((int64_t)xmax - SUBPIX_HALF);
  int64_t j0 = floor_shift((int64_t)ymin - SUBPIX_HALF + (SUBPIX - 1));
  int64_t j1 = floor_shift((int64_t)ymax - SUBPIX_HALF);
  if (i0 < 0) i0 = 0;
  if (i1 > W - 1) i1 = W - 1;
  if (j0 < row0) j0 = row0;
  if (j1 > row1 - 1) j1 = row1 - 1;
  if (i0 > i1 || j0 > j1) return;

  /* edges 0:(v0->v1) 1:(v1->v2) 2:(v2->v0); E(P) = A*(Px-Xa) + B*(Py-Ya), interior > 0 */
  const wvert *ea[3] = {&v0, &v1, &v2}, *eb[3] = {&v1, &v2, &v0};
  int64_t A[3], B[3];
  int tie[3];
  for (int k = 0; k < 3; ++k) {
    A[k] = (int64_t)ea[k]->Y - eb[k]->Y;
    B[k] = (int64_t)eb[k]->X - ea[k]->X;
    tie[k] = (A[k] > 0) || (A[k] == 0 && B[k] > 0);        /* S6 tie-break */
  }

  /* S8: depth plane anchored at v0 */
  float dx1 = (float)(v1.X - v0.X), dy1 = (float)(v1.Y - v0.Y);
  float dx2 = (float)(v2.X - v0.X), dy2 = (float)(v2.Y - v0.Y);
  float dz1 = v1.z - v0.z, dz2 = v2.z - v0.z;
  float fa = (float)area2;
  float t1 = dz2 * dy1;
  float gx = fmaf(dz1, dy2, -t1) / fa;
  float t2 = dz1 * dx2;
  float gy = fmaf(dz2, dx1, -t2) / fa;

  for (int64_t j = j0; j <= j1; ++j) {
    int64_t Py = j * SUBPIX + SUBPIX_HALF;
    float rowz = fmaf(gy, (float)(int32_t)(Py - v0.Y), v0.z);
    for (int64_t i = i0; i <= i1; ++i) {
      int64_t Px = i * SUBPIX + SUBPIX_HALF;
      int inside = 1;
      for (int k = 0; k < 3; ++k) {
        int64_t E = A[k] * (Px - ea[k]->X) + B[k] * (Py - ea[k]->Y);
        if (!(E > 0 || (E == 0 && tie[k]))) { inside = 0; break; }
      }
      if (!inside) continue;
      float z = fmaf(gx, (float)(int32_t)(Px - v0.X), rowz);
      if (!(z > 0.0f)) z = 0.0f;                            /* clamp to depth range */
      if (z < 1.0f) {                                       /* far clip + GL_LESS vs clear 1.0 */
        float *p = &zbuf[(size_t)j * W + i];
        if (z < *p) *p = z;                                 /* GL_LESS, src/urdf_filter.cpp:570 */
      }
    }
  }
}

/* Vertex stage + clipping + viewport for one triangle: appends up to MAX_POLY-2 window-space triangles
 * (a fan around vertex 0 of the clipped polygon) to out[] and returns how many. */
typedef struct { wvert v[3]; } wtri;

static int window_tris(const float *m, const float *tri, float gx, float gy, float halfw, float halfh, wtri *out)
{
  vec4 poly[MAX_POLY];
  poly[0] = xform(m, tri);
  poly[1] = xform(m, tri + 3);
  poly[2] = xform(m, tri + 6);
  if (!finite4(poly[0]) || !finite4(poly[1]) || !finite4(poly[2])) return 0;
  int n = 3, need = 0;
  for (int i = 0; i < 3; ++i)
    for (int k = 0; k < 5; ++k)
      if (!(plane_dist(poly[i], k, gx, gy) >= 0.0f)) need = 1;
  if (need) n = clip_poly(poly, 3, gx, gy);
  if (n < 3) return 0;
  /* S4b: after clipping every vertex satisfies |x| <= gx*w, |y| <= gy*w, z >= -w, hence
   * w >= 0; a vertex with w == 0 (or a NaN produced by clipping) has no window position,
   * so the whole primitive is dropped instead of feeding inf/NaN to the snap. */
  for (int i = 0; i < n; ++i)
    if (!(poly[i].w > 0.0f)) return 0;
  wvert wv[MAX_POLY];
  for (int i = 0; i < n; ++i)
    if (!to_window(poly[i], halfw, halfh, &wv[i])) return 0;
  for (int i = 2; i < n; ++i) {
    out[i - 2].v[0] = wv[0]; out[i - 2].v[1] = wv[i - 1]; out[i - 2].v[2] = wv[i];
  }
  return n - 2;
}

static void render_tri(const float *m, const float *tri, float gx, float gy, float halfw, float halfh,
                       float *zbuf, int W, int H, int row0, int row1)
{
  wtri w[MAX_POLY];
  int n = window_tris(m, tri, gx, gy, halfw, halfh, w);
  for (int i = 0; i < n; ++i) raster_tri(w[i].v[0], w[i].v[1], w[i].v[2], zbuf, W, H, row0, row1);
}

/* the background quad of src/urdf_filter.cpp:591-596 as two triangles */
static void bg_quad(float far_plane_099, float *q)
{
  const float z = far_plane_099;
  const float a[3] = {-100.f, -100.f, z}, b[3] = {100.f, -100.f, z};
  const float c[3] = {100.f, 100.f, z}, d[3] = {-100.f, 100.f, z};
  emit_quad(q, a, b, c, d);
}

/*
 * Render the virtual window-space z-buffer for one frame.
 *   tri       T*9 floats (object space), tri_link[T] indexes mvp
 *   mvp       (L+1)*16 floats from orc_compose_mvp; entry L = background quad
 *   zbuf      W*H floats, window z in [0,1); 1.0f = no fragment (clear depth)
 *   bg_z      eye-space z of the background quad = float(far_plane_*0.99); <=0 disables it
 * nthreads > 1 splits the image in row bands (OpenMP); the result is identical.
 */
ORC_API int orc_render(const float *tri, const uint32_t *tri_link, int64_t T,
                       const float *mvp, int L, int W, int H, float bg_z,
                       float *zbuf, int nthreads)
{
  if (W <= 0 || H <= 0 || W > 4096 || H > 4096) return -1;
  const float halfw = 0.5f * (float)W, halfh = 0.5f * (float)H;
  const float gx = GUARD_PX / halfw, gy = GUARD_PX / halfh;
#pragma omp parallel for schedule(static) num_threads(nthreads > 1 ? nthreads : 1) if (nthreads > 1)
  for (int64_t i = 0; i < (int64_t)W * H; ++i) zbuf[i] = 1.0f;   /* glClear, depth 1.0 */
  for (int64_t t = 0; t < T; ++t)
    if (tri_link[t] >= (uint32_t)L) return -2;
  if (nthreads < 1) nthreads = 1;
  if (nthreads == 1) {
    if (bg_z > 0.0f) {
      float q[18];
      bg_quad(bg_z, q);
      render_tri(mvp + 16 * (size_t)L, q, gx, gy, halfw, halfh, zbuf, W, H, 0, H);
      render_tri(mvp + 16 * (size_t)L, q + 9, gx, gy, halfw, halfh, zbuf, W, H, 0, H);
    }
    for (int64_t t = 0; t < T; ++t)
      render_tri(mvp + 16 * (size_t)tri_link[t], tri + 9 * t, gx, gy, halfw, halfh, zbuf, W, H, 0, H);
    return 0;
  }
  /* Multi-threaded form (same result: the per-pixel min does not depend on the order):
   *   phase 1  vertex stage + clipping once per triangle, threads own contiguous triangle ranges and
   *            keep the window-space triangles that can touch a pixel row of the image;
   *   phase 2  row bands, every band walks all kept triangles and rasterises its rows. */
  typedef struct { wtri *p; size_t n, cap; } wlist;
  wlist *lists = (wlist *)calloc((size_t)nthreads, sizeof(wlist));
  if (!lists) return -3;
  int failed = 0;
#pragma omp parallel num_threads(nthreads)
  {
#ifdef _OPENMP
    const int me = omp_get_thread_num(), nth = omp_get_num_threads();
#else
    const int me = 0, nth = 1;
#endif
    wlist *l = &lists[me];
    const int64_t total = T + (bg_z > 0.0f ? 2 : 0);
    const int64_t lo = total * me / nth, hi = total * (me + 1) / nth;
    float q[18];
    bg_quad(bg_z > 0.0f ? bg_z : 1.0f, q);
    for (int64_t t = lo; t < hi; ++t) {
      wtri w[MAX_POLY];
      int n;
      if (t < T) n = window_tris(mvp + 16 * (size_t)tri_link[t], tri + 9 * t, gx, gy, halfw, halfh, w);
      else n = window_tris(mvp + 16 * (size_t)L, q + 9 * (t - T), gx, gy, halfw, halfh, w);
      for (int i = 0; i < n; ++i) {
        /* cheap reject of triangles that cannot touch any pixel centre row / column of the image */
        int32_t ymin = w[i].v[0].Y, ymax = ymin, xmin = w[i].v[0].X, xmax = xmin;
        for (int k = 1; k < 3; ++k) {
          if (w[i].v[k].Y < ymin) ymin = w[i].v[k].Y;
          if (w[i].v[k].Y > ymax) ymax = w[i].v[k].Y;
          if (w[i].v[k].X < xmin) xmin = w[i].v[k].X;
          if (w[i].v[k].X > xmax) xmax = w[i].v[k].X;
        }
        if (floor_shift((int64_t)ymin - SUBPIX_HALF + (SUBPIX - 1)) > floor_shift((int64_t)ymax - SUBPIX_HALF)) continue;
        if (floor_shift((int64_t)xmin - SUBPIX_HALF + (SUBPIX - 1)) > floor_shift((int64_t)xmax - SUBPIX_HALF)) continue;
        if (ymax < 0 || xmax < 0 || ymin > (int64_t)H * SUBPIX || xmin > (int64_t)W * SUBPIX) continue;
        if (l->n == l->cap) {
          size_t cap = l->cap ? 2 * l->cap : 4096;
          wtri *np_ = (wtri *)realloc(l->p, cap * sizeof(wtri));
          if (!np_) { failed = 1; break; }
          l->p = np_; l->cap = cap;
        }
        l->p[l->n++] = w[i];
      }
    }
  }
  if (!failed) {
    int nb = nthreads * 4;
    if (nb > H) nb = H;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (int b = 0; b < nb; ++b) {
      const int row0 = (int)((int64_t)H * b / nb), row1 = (int)((int64_t)H * (b + 1) / nb);
      const int64_t y_lo = (int64_t)row0 * SUBPIX, y_hi = (int64_t)row1 * SUBPIX;
      for (int th = 0; th < nthreads; ++th) {
        const wlist *l = &lists[th];
        for (size_t i = 0; i < l->n; ++i) {
          const wtri *w = &l->p[i];
          int32_t ymin = w->v[0].Y, ymax = ymin;
          if (w->v[1].Y < ymin) ymin = w->v[1].Y;
          if (w->v[1].Y > ymax) ymax = w->v[1].Y;
          if (w->v[2].Y < ymin) ymin = w->v[2].Y;
          if (w->v[2].Y > ymax) ymax = w->v[2].Y;
          if (ymax < y_lo || ymin > y_hi) continue;
          raster_tri(w->v[0], w->v[1], w->v[2], zbuf, W, H, row0, row1);
        }
      }
    }
  }
  for (int th = 0; th < nthreads; ++th) free(lists[th].p);
  free(lists);
  if (failed) return -3;
  return 0;
}

/* ========================================================================= */
/* 4. Fragment shader + readback + encodings                                  */
/* ========================================================================= */

/* OpenCV Mat::convertTo(CV_32F, 0.001) on 16U: float multiply. src/urdf_filter.cpp:288 */
ORC_API void orc_u16_to_f32(const uint16_t *in, float *out, int64_t n)
{
  for (int64_t i = 0; i < n; ++i) out[i] = (float)in[i] * 0.001f;
}

/* OpenCV saturate_cast<ushort>(cvRound(x*1000.f)). src/urdf_filter.cpp:311.
 * cvRound is cvtss2si: round-half-even; NaN / out of int32 range give INT_MIN -> 0. */
static uint16_t f32_to_u16_one(float x)
{
  float v = x * 1000.0f;
  int32_t r;
  if (!(v >= -2147483648.0f && v < 2147483648.0f)) r = INT32_MIN;
  else r = (int32_t)lrintf(v);
  return (uint16_t)((uint32_t)r <= 65535u ? r : (r > 0 ? 65535 : 0));
}
ORC_API void orc_f32_to_u16(const float *in, uint16_t *out, int64_t n)
{
  for (int64_t i = 0; i < n; ++i) out[i] = f32_to_u16_one(in[i]);
}

/* include/shaders/urdf_filter.frag:14-17 with float uniforms. */
ORC_API float orc_to_linear_depth(float d, float z_near, float z_far)
{
  float k1 = (z_near * z_far) / (z_near - z_far);
  float k2 = z_far / (z_far - z_near);
  return k1 / (d - k2);
}

/*
 * Fragment stage for every pixel given the final depth-buffer value (the shader output
 * that survives GL_LESS is the one evaluated at the nearest fragment).
 *   depth_out : masked depth, attachment 1 .r (src/urdf_filter.cpp:729-730), re-encoded
 *               like src/urdf_filter.cpp:309-312 when enc == ENC_U16_MM
 *   mask_out  : attachment 3 .r read as GL_UNSIGNED_BYTE -> 0 / 255 (:731-735); may be NULL
 *   virt_out  : optional linearised virtual depth (metres; 0 where no fragment)
 * Pixels no fragment covered keep the clear colour 0 (src/urdf_filter.cpp:566).
 */
static void filter_range(const void *depth_in, int enc, const float *zbuf, size_t i0, size_t i1,
                         float z_near, float z_far, float max_diff, float replace_value,
                         void *depth_out, uint8_t *mask_out, float *virt_out)
{
  for (size_t i = i0; i < i1; ++i) {
    float sensor = (enc == ENC_U16_MM) ? (float)((const uint16_t *)depth_in)[i] * 0.001f
                                       : ((const float *)depth_in)[i];
    float out, s, virt = 0.0f;
    /* What mix(sensor, replace, a) = sensor * (1 - a) + replace * a (frag:29) does to the exotic values of a 32FC1 image,
     * as measured in the GL driver (oracle/gl_ref on Mesa llvmpipe, tests/golden/gl_llvmpipe.npz "special" case):
     *   |sensor| below FLT_MIN (denormals, -0.0)  ->  +0.0   (the shader runs with denormals-are-zero; -0 + 0 = +0)
     *   sensor = +inf on a filtered pixel          ->  NaN    (inf * 0; the x86 default NaN 0xffc00000 in that driver)
     * NaN and -inf are never filtered (the comparison is false) and pass through with their bits. */
    if (enc != ENC_U16_MM && fabsf(sensor) < 1.17549435e-38f) sensor = 0.0f;
    if (zbuf[i] == 1.0f) {             /* never drawn: clear colour */
      out = 0.0f; s = 0.0f;
    } else {
      virt = orc_to_linear_depth(zbuf[i], z_near, z_far);            /* frag:22 */
      s = (sensor > (virt - max_diff)) ? 1.0f : 0.0f;                 /* frag:23 */
      if (s != 0.0f) {                                                /* frag:29, mix with a in {0,1} */
        out = replace_value;
        if (sensor > 3.40282347e+38f) { const uint32_t qnan = 0xffc00000u; memcpy(&out, &qnan, 4); }
      } else {
        out = sensor;
      }
    }
    if (enc == ENC_U16_MM) ((uint16_t *)depth_out)[i] = f32_to_u16_one(out);
    else ((float *)depth_out)[i] = out;
    if (mask_out) mask_out[i] = (s != 0.0f) ? 255 : 0;
    if (virt_out) virt_out[i] = virt;
  }
}

ORC_API void orc_filter_mt(const void *depth_in, int enc, const float *zbuf, int W, int H,
                           float z_near, float z_far, float max_diff, float replace_value,
                           void *depth_out, uint8_t *mask_out, float *virt_out, int nthreads)
{
  const size_t n = (size_t)W * H;
  if (nthreads < 1) nthreads = 1;
  const int chunks = nthreads * 4;
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
  for (int c = 0; c < chunks; ++c)
    filter_range(depth_in, enc, zbuf, n * c / chunks, n * (c + 1) / chunks, z_near, z_far, max_diff,
                 replace_value, depth_out, mask_out, virt_out);
}

ORC_API void orc_filter(const void *depth_in, int enc, const float *zbuf, int W, int H,
                        float z_near, float z_far, float max_diff, float replace_value,
                        void *depth_out, uint8_t *mask_out, float *virt_out)
{
  filter_range(depth_in, enc, zbuf, 0, (size_t)W * H, z_near, z_far, max_diff, replace_value, depth_out,
               mask_out, virt_out);
}

/* One whole frame: render + filter.  Returns 0 on success. */
ORC_API int orc_filter_frame(const void *depth_in, int enc, int W, int H,
                             const float *tri, const uint32_t *tri_link, int64_t T,
                             const float *mvp, int L,
                             float z_near, float z_far, float max_diff, float replace_value,
                             void *depth_out, uint8_t *mask_out, float *zbuf_out, int nthreads)
{
  float *zbuf = zbuf_out ? zbuf_out : (float *)malloc(sizeof(float) * (size_t)W * H);
  if (!zbuf) return -3;
  float bg_z = (float)((double)z_far * 0.99);       /* glVertex3f(.., far_plane_*0.99) */
  int rc = orc_render(tri, tri_link, T, mvp, L, W, H, bg_z, zbuf, nthreads);
  if (rc == 0)
    orc_filter_mt(depth_in, enc, zbuf, W, H, z_near, z_far, max_diff, replace_value, depth_out, mask_out, NULL,
                  nthreads);
  if (!zbuf_out) free(zbuf);
  return rc;
}

ORC_API int orc_max_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
