// gl_crosscheck.cpp -- GL cross-check harness.  TEST INFRASTRUCTURE ONLY.
//
// Replays, in a headless OSMesa (Mesa llvmpipe) context, the GL call sequence of the reference's
// RealtimeURDFFilter::render (src/urdf_filter.cpp:546-644 and :729-735) with the reference's OWN shader files
// loaded UNMODIFIED from <shader_dir>/urdf_filter.vert|.frag, on a case written by make_case.py, and dumps colour
// attachment 1 (filtered depth, GL_RED/GL_FLOAT) and attachment 3 (mask, GL_RED/GL_UNSIGNED_BYTE).  compare.py puts
// the dump next to the CPU oracle: this is what turns "parity by restatement" into measured parity wherever Mesa
// exists.  Neither this image nor the GPU box has any GL library, so here the file is only syntax-checked against
// oracle/gl_ref/stubs (tests/test_gl_crosscheck.py); with Mesa: `make -C oracle/gl_ref` builds it into oracle/_ref/.
//
//   gl_crosscheck <case.bin> <shader_dir> <out.bin> [frames]
//
// With a frame count the frame (upload of the sensor depth, render, both read-backs: what RealtimeURDFFilter::filter does per
// call, src/urdf_filter.cpp:234-237 + :729-735) is repeated and timed from the second one on: the reference's own CPU
// figure on this machine's cores (llvmpipe renders with LP_NUM_THREADS threads, default = all cores).
//
// Sequence kept from the reference: 4 x RGBA32F rectangle-texture attachments + a 24-bit depth texture
// (src/FrameBufferObject.cpp:868-880,1003-1023), sensor depth in a GL_R32F buffer texture re-specified per frame
// (:332-353), glDrawBuffers(4), clear to (0,0,0,1) / depth 1, GL_DEPTH_TEST with the default GL_LESS, PROJECTION =
// glTf, MODELVIEW = gluLookAt(0,0,0, 0,0,1, 0,1,0), the background quad at 0.99 * far drawn BEFORE the uniforms are
// set (:591-596 vs :625-631 -- so the frame is rendered twice and the second, steady-state one is dumped),
// MODELVIEW *= inverse(camera_offset) * camera_transform (:602-614), then per part glPushMatrix, glMultMatrixd
// (link_to_fixed * link_offset, src/renderable.cpp:59-68), the optional glTranslatef / glScalef suffix (:95,128,427),
// the triangles, glPopMatrix.  The matrix stack lives inside GL (float), exactly what the oracle can only approximate.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <string>
#include <vector>

#define GL_GLEXT_PROTOTYPES 1
#include <GL/gl.h>
#include <GL/glext.h>
#ifdef RUF_GL_VIA_GLX
// Mesa's xlib software GLX on top of oracle/gl_ref/fakex11 (no X server, no X11 headers in the image): the handful of GLX
// declarations this file needs, spelled out.  See fakex11/fakex11.c and the Makefile's `glx` target.
extern "C" {
typedef struct _XDisplay Display;
typedef struct __GLXcontextRec *GLXContext;
typedef struct { void *visual; unsigned long visualid; int screen; int depth; int c_class; unsigned long red_mask, green_mask, blue_mask;
                 int colormap_size; int bits_per_rgb; } XVisualInfo;
Display *fakex_open_display(int width, int height);
unsigned long fakex_window(void);
XVisualInfo *glXChooseVisual(Display *, int, int *);
GLXContext glXCreateContext(Display *, XVisualInfo *, GLXContext, int);
int glXMakeCurrent(Display *, unsigned long, GLXContext);
}
#else
#include <GL/osmesa.h>
#endif

#ifdef RUF_EMBEDDED_SHADERS
// the reference's shader files as they were at build time (oracle/gl_ref/embed_shaders.py): `<shader_dir>` = "-" uses them
extern "C" const char ruf_ref_vert_src[];
extern "C" const char ruf_ref_frag_src[];
#endif

namespace {

struct Part { double model[16]; int32_t suffix_kind; float sx, sy, sz; };   // suffix: 0 none, 1 glTranslatef, 2 glScalef
struct Case {
  int32_t W = 0, H = 0, n_parts = 0, n_tris = 0;
  double proj[16], offset_inv[16], cam[16];
  float z_near = 0, z_far = 0, max_diff = 0, replace_value = 0;
  std::vector<Part> parts;
  std::vector<float> xyz;          // 9 per triangle
  std::vector<uint32_t> tri_part;
  std::vector<float> depth;        // W*H metres, row 0 first (what filter_callback hands to filter(), :287-303)
};

template <class T> bool rd(std::istream &f, T *p, size_t n) { return (bool)f.read(reinterpret_cast<char *>(p), sizeof(T) * n); }

bool load_case(const char *path, Case &c)
{
  std::ifstream f(path, std::ios::binary);
  char magic[8];
  if (!f || !rd(f, magic, 8) || std::memcmp(magic, "RUFGLC01", 8) != 0) return false;
  int32_t hdr[4];
  if (!rd(f, hdr, 4)) return false;
  c.W = hdr[0]; c.H = hdr[1]; c.n_parts = hdr[2]; c.n_tris = hdr[3];
  float sp[4];
  if (!rd(f, c.proj, 16) || !rd(f, c.offset_inv, 16) || !rd(f, c.cam, 16) || !rd(f, sp, 4)) return false;
  c.z_near = sp[0]; c.z_far = sp[1]; c.max_diff = sp[2]; c.replace_value = sp[3];
  c.parts.resize(c.n_parts);
  for (Part &p : c.parts) {
    float s[3];
    if (!rd(f, p.model, 16) || !rd(f, &p.suffix_kind, 1) || !rd(f, s, 3)) return false;
    p.sx = s[0]; p.sy = s[1]; p.sz = s[2];
  }
  c.xyz.resize((size_t)c.n_tris * 9); c.tri_part.resize(c.n_tris); c.depth.resize((size_t)c.W * c.H);
  return rd(f, c.xyz.data(), c.xyz.size()) && rd(f, c.tri_part.data(), c.tri_part.size()) && rd(f, c.depth.data(), c.depth.size());
}

std::string slurp(const std::string &path)
{
  std::ifstream f(path);
  if (!f) { std::fprintf(stderr, "cannot read %s\n", path.c_str()); std::exit(2); }
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

GLuint compile(GLenum type, const std::string &src)      // the text goes to the driver byte for byte
{
  GLuint s = glCreateShader(type);
  const char *p = src.c_str();
  glShaderSource(s, 1, &p, nullptr);
  glCompileShader(s);
  GLint ok = 0;
  glGetShaderiv(s, GL_COMPILE_STATUS, &ok);
  if (!ok) { char log[4096]; glGetShaderInfoLog(s, sizeof(log), nullptr, log); std::fprintf(stderr, "shader: %s\n", log); std::exit(3); }
  return s;
}

// gluLookAt(0,0,0, 0,0,1, 0,1,0) as GLU issues it: glMultMatrixf(M) then glTranslated(-eye); f = (0,0,1), up = (0,1,0)
// -> s = f x up = (-1,0,0), u = s x f = (0,1,0): M = rows (s, u, -f)
void look_at_kinect()
{
  const GLfloat m[16] = {-1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0, 1};
  glMultMatrixf(m);
  glTranslated(-0.0, -0.0, -0.0);
}

void check(const char *what)
{
  const GLenum e = glGetError();
  if (e != GL_NO_ERROR) { std::fprintf(stderr, "GL error 0x%x after %s\n", e, what); std::exit(4); }
}

}  // namespace

int main(int argc, char **argv)
{
  if (argc != 4 && argc != 5) { std::fprintf(stderr, "usage: %s case.bin shader_dir out.bin [frames]\n", argv[0]); return 1; }
  const int n_frames = argc == 5 ? std::max(2, std::atoi(argv[4])) : 2;
  Case c;
  if (!load_case(argv[1], c)) { std::fprintf(stderr, "bad case file %s\n", argv[1]); return 1; }

  // headless compatibility-profile context (the shaders use gl_ModelViewProjectionMatrix / gl_Vertex / gl_FragData)
#ifdef RUF_GL_VIA_GLX
  Display *dpy = fakex_open_display(c.W, c.H);
  int visual_attribs[] = {4 /* GLX_RGBA */, 8 /* GLX_RED_SIZE */, 8, 9, 8, 10, 8, 12 /* GLX_DEPTH_SIZE */, 24, 0};
  XVisualInfo *vi = glXChooseVisual(dpy, 0, visual_attribs);
  GLXContext ctx = vi ? glXCreateContext(dpy, vi, nullptr, 1) : nullptr;
  if (!ctx || !glXMakeCurrent(dpy, fakex_window(), ctx)) { std::fprintf(stderr, "no GLX context\n"); return 5; }
#else
  const int attribs[] = {OSMESA_FORMAT, OSMESA_RGBA, OSMESA_DEPTH_BITS, 24, OSMESA_PROFILE, OSMESA_COMPAT_PROFILE,
                         OSMESA_CONTEXT_MAJOR_VERSION, 3, OSMESA_CONTEXT_MINOR_VERSION, 1, 0};
  OSMesaContext ctx = OSMesaCreateContextAttribs(attribs, nullptr);
  std::vector<unsigned char> window((size_t)c.W * c.H * 4);
  if (!ctx || !OSMesaMakeCurrent(ctx, window.data(), GL_UNSIGNED_BYTE, c.W, c.H)) { std::fprintf(stderr, "no OSMesa context\n"); return 5; }
#endif
  std::fprintf(stderr, "GL_RENDERER %s | GL_VERSION %s\n", (const char *)glGetString(GL_RENDERER), (const char *)glGetString(GL_VERSION));

  // FramebufferObject("rgba=4x32t depth=24t stencil=8t"): 4 RGBA32F rectangle textures + a D24 texture
  GLuint fbo = 0, col[4], dep = 0;
  glGenFramebuffers(1, &fbo);
  glBindFramebuffer(GL_FRAMEBUFFER, fbo);
  glGenTextures(4, col);
  for (int i = 0; i < 4; ++i) {
    glBindTexture(GL_TEXTURE_RECTANGLE, col[i]);
    glTexImage2D(GL_TEXTURE_RECTANGLE, 0, GL_RGBA32F, c.W, c.H, 0, GL_RGBA, GL_FLOAT, nullptr);
    glTexParameteri(GL_TEXTURE_RECTANGLE, GL_TEXTURE_MIN_FILTER, GL_NEAREST);
    glTexParameteri(GL_TEXTURE_RECTANGLE, GL_TEXTURE_MAG_FILTER, GL_NEAREST);
    glFramebufferTexture2D(GL_FRAMEBUFFER, GL_COLOR_ATTACHMENT0 + i, GL_TEXTURE_RECTANGLE, col[i], 0);
  }
  glGenTextures(1, &dep);
  glBindTexture(GL_TEXTURE_RECTANGLE, dep);
  glTexImage2D(GL_TEXTURE_RECTANGLE, 0, GL_DEPTH_COMPONENT24, c.W, c.H, 0, GL_DEPTH_COMPONENT, GL_FLOAT, nullptr);
  glFramebufferTexture2D(GL_FRAMEBUFFER, GL_DEPTH_ATTACHMENT, GL_TEXTURE_RECTANGLE, dep, 0);
  if (glCheckFramebufferStatus(GL_FRAMEBUFFER) != GL_FRAMEBUFFER_COMPLETE) { std::fprintf(stderr, "FBO incomplete\n"); return 6; }
  check("fbo");

  // textureBufferFromDepthBuffer (:332-353): buffer object re-specified per frame, GL_R32F buffer texture
  GLuint tbo = 0, tex = 0;
  glGenBuffers(1, &tbo);
  glGenTextures(1, &tex);

  const std::string dir = argv[2];
  std::string vert_src, frag_src;
  if (dir == "-") {
#ifdef RUF_EMBEDDED_SHADERS
    vert_src = ruf_ref_vert_src; frag_src = ruf_ref_frag_src;
#endif
    if (vert_src.empty() || frag_src.empty()) { std::fprintf(stderr, "this binary was built without the reference's shader files\n"); return 2; }
  } else {
    vert_src = slurp(dir + "/urdf_filter.vert"); frag_src = slurp(dir + "/urdf_filter.frag");
  }
  GLuint prog = glCreateProgram();
  glAttachShader(prog, compile(GL_VERTEX_SHADER, vert_src));
  glAttachShader(prog, compile(GL_FRAGMENT_SHADER, frag_src));
  glLinkProgram(prog);
  GLint linked = 0;
  glGetProgramiv(prog, GL_LINK_STATUS, &linked);
  if (!linked) { std::fprintf(stderr, "program does not link\n"); return 7; }

  std::vector<float> masked((size_t)c.W * c.H);
  std::vector<unsigned char> mask((size_t)c.W * c.H);
  glPixelStorei(GL_PACK_ALIGNMENT, 1);            // the reference relies on W % 4 == 0 with the default 4
  std::chrono::steady_clock::time_point t_start;
  for (int frame = 0; frame < n_frames; ++frame) {       // frame 0 shades the background quad with default uniforms (see header)
    if (frame == 1) t_start = std::chrono::steady_clock::now();
    glBindBuffer(GL_TEXTURE_BUFFER, tbo);
    glBufferData(GL_TEXTURE_BUFFER, (GLsizeiptr)(c.depth.size() * sizeof(float)), c.depth.data(), GL_DYNAMIC_DRAW);
    glActiveTexture(GL_TEXTURE0);
    glBindTexture(GL_TEXTURE_BUFFER, tex);
    glTexBuffer(GL_TEXTURE_BUFFER, GL_R32F, tbo);

    glBindFramebuffer(GL_FRAMEBUFFER, fbo);       // beginCapture: viewport = W x H (src/FrameBufferObject.cpp:776)
    glViewport(0, 0, c.W, c.H);
    glUseProgram(prog);
    const GLenum bufs[4] = {GL_COLOR_ATTACHMENT0, GL_COLOR_ATTACHMENT1, GL_COLOR_ATTACHMENT2, GL_COLOR_ATTACHMENT3};
    glDrawBuffers(4, bufs);
    glClearColor(0.0f, 0.0f, 0.0f, 1.0f);
    glClear(GL_COLOR_BUFFER_BIT | GL_DEPTH_BUFFER_BIT | GL_STENCIL_BUFFER_BIT);
    glEnable(GL_DEPTH_TEST);

    glMatrixMode(GL_PROJECTION);
    glLoadIdentity();
    glMultMatrixd(c.proj);
    glMatrixMode(GL_MODELVIEW);
    glLoadIdentity();
    look_at_kinect();
    const float zq = c.z_far * 0.99f;
    glBegin(GL_QUADS);
    glVertex3f(-100.0f, -100.0f, zq); glVertex3f(100.0f, -100.0f, zq); glVertex3f(100.0f, 100.0f, zq); glVertex3f(-100.0f, 100.0f, zq);
    glEnd();
    glMultMatrixd(c.offset_inv);
    glMultMatrixd(c.cam);

    glUniform1i(glGetUniformLocation(prog, "depth_texture"), 0);
    glUniform1i(glGetUniformLocation(prog, "width"), c.W);
    glUniform1i(glGetUniformLocation(prog, "height"), c.H);
    glUniform1f(glGetUniformLocation(prog, "z_far"), c.z_far);
    glUniform1f(glGetUniformLocation(prog, "z_near"), c.z_near);
    glUniform1f(glGetUniformLocation(prog, "max_diff"), c.max_diff);
    glUniform1f(glGetUniformLocation(prog, "replace_value"), c.replace_value);

    glEnableClientState(GL_VERTEX_ARRAY);
    size_t t = 0;
    for (int p = 0; p < c.n_parts; ++p) {         // triangles are grouped by part, in draw order
      size_t t1 = t;
      while (t1 < (size_t)c.n_tris && c.tri_part[t1] == (uint32_t)p) ++t1;
      glPushMatrix();
      glMultMatrixd(c.parts[p].model);
      if (c.parts[p].suffix_kind == 1) glTranslatef(c.parts[p].sx, c.parts[p].sy, c.parts[p].sz);
      if (c.parts[p].suffix_kind == 2) glScalef(c.parts[p].sx, c.parts[p].sy, c.parts[p].sz);
      if (t1 > t) {
        glVertexPointer(3, GL_FLOAT, 0, c.xyz.data() + 9 * t);
        glDrawArrays(GL_TRIANGLES, 0, (GLsizei)(3 * (t1 - t)));
      }
      glPopMatrix();
      t = t1;
    }
    glDisableClientState(GL_VERTEX_ARRAY);
    glUseProgram(0);
    glBindFramebuffer(GL_FRAMEBUFFER, 0);
    // readback exactly as :729-735 (bottom-up GL rows = the same memory order as the input image)
    glBindTexture(GL_TEXTURE_RECTANGLE, col[1]);
    glGetTexImage(GL_TEXTURE_RECTANGLE, 0, GL_RED, GL_FLOAT, masked.data());
    glBindTexture(GL_TEXTURE_RECTANGLE, col[3]);
    glGetTexImage(GL_TEXTURE_RECTANGLE, 0, GL_RED, GL_UNSIGNED_BYTE, mask.data());
    check("frame");
  }
  if (argc == 5) {
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count() / (n_frames - 1);
    std::fprintf(stderr, "TIMING frames %d ms_per_frame %.4f frames_per_s %.2f\n", n_frames - 1, ms, 1000.0 / ms);
  }

  std::ofstream out(argv[3], std::ios::binary);
  out.write("RUFGLO01", 8);
  const int32_t hdr[2] = {c.W, c.H};
  out.write(reinterpret_cast<const char *>(hdr), sizeof(hdr));
  out.write(reinterpret_cast<const char *>(masked.data()), (std::streamsize)(masked.size() * sizeof(float)));
  out.write(reinterpret_cast<const char *>(mask.data()), (std::streamsize)mask.size());
#ifndef RUF_GL_VIA_GLX
  OSMesaDestroyContext(ctx);
#endif
  return out ? 0 : 8;
}
