/* fakex11.c -- the 30 Xlib / XShm entry points that Mesa's xlib software GLX needs, without an X server.
 * TEST INFRASTRUCTURE ONLY (oracle/gl_ref).
 *
 * The only GL implementation in this image is the Mesa 18.1.9 llvmpipe libGL that NVIDIA ships inside Nsight Compute for its
 * own GUI (/opt/nvidia/nsight-compute/.../Mesa/libGL.so.1): the "xlib" flavour, which renders in software and talks to X only
 * to learn the visual and the window size and to present frames.  The cross-check harness renders into a framebuffer object
 * and never presents, so a Display that answers those few questions is enough.  Built twice from this file, as libX11.so.6
 * and as libXext.so.6 (the two DT_NEEDED names of that libGL); struct layouts follow the public part of <X11/Xlib.h> and
 * <X11/Xutil.h> (libX11 1.6, LP64), written out here because the image has no X11 headers either. */
#include <stdlib.h>
#include <string.h>

typedef unsigned long XID;
typedef XID Window, Drawable, Pixmap, Colormap, VisualID, Font;
typedef int Bool, Status;
typedef char *XPointer;
typedef struct _XExtData XExtData;
typedef struct _XGC { int dummy; } *GC;

typedef struct {
  XExtData *ext_data; VisualID visualid; int c_class; unsigned long red_mask, green_mask, blue_mask; int bits_per_rgb; int map_entries;
} Visual;
typedef struct { int depth; int nvisuals; Visual *visuals; } Depth;
typedef struct { XExtData *ext_data; int depth; int bits_per_pixel; int scanline_pad; } ScreenFormat;
struct _XDisplay;
typedef struct {
  XExtData *ext_data; struct _XDisplay *display; Window root; int width, height; int mwidth, mheight; int ndepths; Depth *depths;
  int root_depth; Visual *root_visual; GC default_gc; Colormap cmap; unsigned long white_pixel; unsigned long black_pixel;
  int max_maps, min_maps; int backing_store; Bool save_unders; long root_input_mask;
} Screen;
typedef struct _XDisplay {
  XExtData *ext_data; struct _XPrivate *private1; int fd; int private2; int proto_major_version; int proto_minor_version; char *vendor;
  XID private3; XID private4; XID private5; int private6; XID (*resource_alloc)(struct _XDisplay *); int byte_order; int bitmap_unit;
  int bitmap_pad; int bitmap_bit_order; int nformats; ScreenFormat *pixmap_format; int private8; int release;
  struct _XPrivate *private9, *private10; int qlen; unsigned long last_request_read; unsigned long request; XPointer private11;
  XPointer private12; XPointer private13; XPointer private14; unsigned max_request_size; struct _XrmHashBucketRec *db;
  int (*private15)(struct _XDisplay *); char *display_name; int default_screen; int nscreens; Screen *screens;
  unsigned long motion_buffer; unsigned long private16; int min_keycode; int max_keycode; XPointer private17; XPointer private18;
  int private19; char *xdefaults;
  /* "there is more to this structure, but it is private to Xlib" -- Mesa's xlib GLX reaches into it once: it hangs its
   * close-display callback on the head of dpy->ext_procs right after XAddExtension (Xlibint.h, libX11 1.6 layout) */
  char *scratch_buffer; unsigned long scratch_length; int ext_number; struct _XExten *ext_procs;
  char more_private[4096];
} Display;
typedef struct { int extension; int major_opcode; int first_event; int first_error; } XExtCodes;
typedef struct _XExten {
  struct _XExten *next; XExtCodes codes; void *create_GC, *copy_GC, *flush_GC, *free_GC, *create_Font, *free_Font, *close_display, *error,
      *error_string; char *name; void *error_values, *before_flush; struct _XExten *next_flush;
} _XExtension;
typedef struct {
  Visual *visual; VisualID visualid; int screen; int depth; int c_class; unsigned long red_mask, green_mask, blue_mask;
  int colormap_size; int bits_per_rgb;
} XVisualInfo;
typedef struct _XImage {
  int width, height; int xoffset; int format; char *data; int byte_order; int bitmap_unit; int bitmap_bit_order; int bitmap_pad;
  int depth; int bytes_per_line; int bits_per_pixel; unsigned long red_mask, green_mask, blue_mask; XPointer obdata;
  struct funcs {
    struct _XImage *(*create_image)(Display *, Visual *, unsigned, int, int, char *, unsigned, unsigned, int, int);
    int (*destroy_image)(struct _XImage *);
    unsigned long (*get_pixel)(struct _XImage *, int, int);
    int (*put_pixel)(struct _XImage *, int, int, unsigned long);
    struct _XImage *(*sub_image)(struct _XImage *, int, int, unsigned, unsigned);
    int (*add_pixel)(struct _XImage *, long);
  } f;
} XImage;
typedef struct {
  int x, y; int width, height; int border_width; int depth; Visual *visual; Window root; int c_class; int bit_gravity; int win_gravity;
  int backing_store; unsigned long backing_planes; unsigned long backing_pixel; Bool save_under; Colormap colormap; Bool map_installed;
  int map_state; long all_event_masks; long your_event_mask; long do_not_propagate_mask; Bool override_redirect; Screen *screen;
} XWindowAttributes;

#ifndef FAKEX_XEXT   /* ======================= libX11.so.6 ======================= */
/* ---- the one screen, one depth, one TrueColor visual, one window of this "server" ---- */
static Visual g_visual = {0, 0x21, 4 /* TrueColor */, 0xff0000ul, 0x00ff00ul, 0x0000fful, 8, 256};
static Depth g_depth = {24, 1, &g_visual};
static ScreenFormat g_format = {0, 24, 32, 32};
static struct _XGC g_gc;
static Screen g_screen;
static Display g_display;
static int g_win_w = 64, g_win_h = 64;
enum { kRoot = 0x100, kWindow = 0x200, kColormap = 0x20 };

Display *fakex_open_display(int width, int height)        /* the harness calls this instead of XOpenDisplay */
{
  g_win_w = width; g_win_h = height;
  memset(&g_screen, 0, sizeof(g_screen));
  memset(&g_display, 0, sizeof(g_display));
  g_screen.display = &g_display; g_screen.root = kRoot; g_screen.width = 4096; g_screen.height = 4096; g_screen.mwidth = 1000;
  g_screen.mheight = 1000; g_screen.ndepths = 1; g_screen.depths = &g_depth; g_screen.root_depth = 24; g_screen.root_visual = &g_visual;
  g_screen.default_gc = &g_gc; g_screen.cmap = kColormap; g_screen.white_pixel = 0xffffff; g_screen.black_pixel = 0;
  g_screen.max_maps = 1; g_screen.min_maps = 1;
  g_display.fd = -1; g_display.proto_major_version = 11; g_display.vendor = (char *)"fakex11 (oracle/gl_ref)"; g_display.byte_order = 0;
  g_display.bitmap_unit = 32; g_display.bitmap_pad = 32; g_display.bitmap_bit_order = 0; g_display.nformats = 1;
  g_display.pixmap_format = &g_format; g_display.release = 1; g_display.max_request_size = 65535;
  g_display.display_name = (char *)":fake"; g_display.default_screen = 0; g_display.nscreens = 1; g_display.screens = &g_screen;
  return &g_display;
}
Window fakex_window(void) { return kWindow; }

void (*_XLockMutex_fn)(void *) = 0;
void (*_XUnlockMutex_fn)(void *) = 0;
void *_Xglobal_lock = 0;

XVisualInfo *XGetVisualInfo(Display *dpy, long mask, XVisualInfo *t, int *n)
{
  (void)dpy;
  *n = 0;
  if ((mask & 0x1) && t->visualid != g_visual.visualid) return 0;
  if ((mask & 0x2) && t->screen != 0) return 0;
  if ((mask & 0x4) && t->depth != 24) return 0;
  if ((mask & 0x8) && t->c_class != g_visual.c_class) return 0;
  XVisualInfo *v = (XVisualInfo *)calloc(1, sizeof(XVisualInfo));
  v->visual = &g_visual; v->visualid = g_visual.visualid; v->screen = 0; v->depth = 24; v->c_class = g_visual.c_class;
  v->red_mask = g_visual.red_mask; v->green_mask = g_visual.green_mask; v->blue_mask = g_visual.blue_mask;
  v->colormap_size = 256; v->bits_per_rgb = 8;
  *n = 1;
  return v;
}
int XFree(void *p) { free(p); return 1; }

static int img_destroy(XImage *i) { if (i) { free(i->data); free(i); } return 1; }
static unsigned long img_get(XImage *i, int x, int y)
{
  unsigned long v = 0;
  memcpy(&v, i->data + (size_t)y * i->bytes_per_line + (size_t)x * (i->bits_per_pixel / 8), i->bits_per_pixel / 8);
  return v;
}
static int img_put(XImage *i, int x, int y, unsigned long v)
{
  memcpy(i->data + (size_t)y * i->bytes_per_line + (size_t)x * (i->bits_per_pixel / 8), &v, i->bits_per_pixel / 8);
  return 1;
}
static XImage *img_sub(XImage *i, int x, int y, unsigned w, unsigned h) { (void)i; (void)x; (void)y; (void)w; (void)h; return 0; }
static int img_add(XImage *i, long v) { (void)i; (void)v; return 1; }
XImage *XCreateImage(Display *dpy, Visual *vis, unsigned depth, int format, int offset, char *data, unsigned w, unsigned h,
                     int bitmap_pad, int bytes_per_line)
{
  (void)dpy;
  XImage *i = (XImage *)calloc(1, sizeof(XImage));
  i->width = (int)w; i->height = (int)h; i->xoffset = offset; i->format = format; i->data = data; i->byte_order = 0;
  i->bitmap_unit = 32; i->bitmap_bit_order = 0; i->bitmap_pad = bitmap_pad ? bitmap_pad : 32; i->depth = (int)depth;
  i->bits_per_pixel = depth > 16 ? 32 : (depth > 8 ? 16 : (depth > 1 ? 8 : 1));
  i->bytes_per_line = bytes_per_line ? bytes_per_line : (int)(((w * (unsigned)i->bits_per_pixel + 31u) / 32u) * 4u);
  if (vis) { i->red_mask = vis->red_mask; i->green_mask = vis->green_mask; i->blue_mask = vis->blue_mask; }
  i->f.create_image = XCreateImage; i->f.destroy_image = img_destroy; i->f.get_pixel = img_get; i->f.put_pixel = img_put;
  i->f.sub_image = img_sub; i->f.add_pixel = img_add;
  return i;
}

Status XGetWindowAttributes(Display *dpy, Window w, XWindowAttributes *a)
{
  (void)dpy; (void)w;
  memset(a, 0, sizeof(*a));
  a->width = g_win_w; a->height = g_win_h; a->depth = 24; a->visual = &g_visual; a->root = kRoot; a->c_class = 1 /* InputOutput */;
  a->colormap = kColormap; a->map_installed = 1; a->map_state = 2 /* IsViewable */; a->screen = &g_screen;
  return 1;
}
Status XGetGeometry(Display *dpy, Drawable d, Window *root, int *x, int *y, unsigned *w, unsigned *h, unsigned *bw, unsigned *depth)
{
  (void)dpy; (void)d;
  *root = kRoot; *x = 0; *y = 0; *w = (unsigned)g_win_w; *h = (unsigned)g_win_h; *bw = 0; *depth = 24;
  return 1;
}
GC XCreateGC(Display *dpy, Drawable d, unsigned long mask, void *values) { (void)dpy; (void)d; (void)mask; (void)values; return (GC)calloc(1, sizeof(struct _XGC)); }
int XFreeGC(Display *dpy, GC gc) { (void)dpy; if (gc != &g_gc) free(gc); return 1; }
int XSetFunction(Display *dpy, GC gc, int f) { (void)dpy; (void)gc; (void)f; return 1; }
int XSetForeground(Display *dpy, GC gc, unsigned long p) { (void)dpy; (void)gc; (void)p; return 1; }
int XFillRectangle(Display *dpy, Drawable d, GC gc, int x, int y, unsigned w, unsigned h) { (void)dpy; (void)d; (void)gc; (void)x; (void)y; (void)w; (void)h; return 1; }
int XPutImage(Display *dpy, Drawable d, GC gc, XImage *i, int sx, int sy, int dx, int dy, unsigned w, unsigned h)
{ (void)dpy; (void)d; (void)gc; (void)i; (void)sx; (void)sy; (void)dx; (void)dy; (void)w; (void)h; return 0; }
XImage *XGetImage(Display *dpy, Drawable d, int x, int y, unsigned w, unsigned h, unsigned long planes, int format)
{ (void)dpy; (void)d; (void)x; (void)y; (void)w; (void)h; (void)planes; (void)format; return 0; }
int XFlush(Display *dpy) { (void)dpy; return 1; }
int XSync(Display *dpy, Bool discard) { (void)dpy; (void)discard; return 1; }
int (*XSynchronize(Display *dpy, Bool onoff))(Display *) { (void)dpy; (void)onoff; return 0; }
typedef int (*XErrorHandler)(Display *, void *);
XErrorHandler XSetErrorHandler(XErrorHandler h) { static XErrorHandler cur = 0; XErrorHandler old = cur; cur = h; return old; }
Bool XQueryExtension(Display *dpy, const char *name, int *major, int *event, int *error)
{ (void)dpy; (void)name; if (major) *major = 0; if (event) *event = 0; if (error) *error = 0; return 0; }
XExtCodes *XAddExtension(Display *dpy)
{
  _XExtension *e = (_XExtension *)calloc(1, sizeof(_XExtension));
  e->codes.extension = dpy->ext_number++;
  e->next = dpy->ext_procs;
  dpy->ext_procs = e;
  return &e->codes;
}
Colormap XCreateColormap(Display *dpy, Window w, Visual *v, int alloc) { (void)dpy; (void)w; (void)v; (void)alloc; return kColormap + 1; }
Pixmap XCreatePixmap(Display *dpy, Drawable d, unsigned w, unsigned h, unsigned depth) { (void)dpy; (void)d; (void)w; (void)h; (void)depth; return 0x300; }
int XFreePixmap(Display *dpy, Pixmap p) { (void)dpy; (void)p; return 1; }
void *XQueryFont(Display *dpy, XID id) { (void)dpy; (void)id; return 0; }
int XFreeFontInfo(char **names, void *info, int n) { (void)names; (void)info; (void)n; return 1; }
int XDrawString16(Display *dpy, Drawable d, GC gc, int x, int y, const void *s, int n) { (void)dpy; (void)d; (void)gc; (void)x; (void)y; (void)s; (void)n; return 0; }
#else                /* ======================= libXext.so.6 ====================== */
/* libXext: no MIT-SHM here (XQueryExtension said so); present so that the dynamic linker is satisfied */
Bool XShmAttach(Display *dpy, void *info) { (void)dpy; (void)info; return 0; }
XImage *XShmCreateImage(Display *dpy, Visual *v, unsigned depth, int format, char *data, void *info, unsigned w, unsigned h)
{ (void)dpy; (void)v; (void)depth; (void)format; (void)data; (void)info; (void)w; (void)h; return 0; }
Bool XShmPutImage(Display *dpy, Drawable d, GC gc, XImage *i, int sx, int sy, int dx, int dy, unsigned w, unsigned h, Bool ev)
{ (void)dpy; (void)d; (void)gc; (void)i; (void)sx; (void)sy; (void)dx; (void)dy; (void)w; (void)h; (void)ev; return 0; }
#endif
