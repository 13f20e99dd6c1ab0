/* Declarations-only stand-in for Mesa's <GL/osmesa.h> (values from Mesa's header).  Never linked. */
#ifndef RUF_STUB_OSMESA_H
#define RUF_STUB_OSMESA_H
#include <GL/gl.h>
#define OSMESA_RGBA 0x1908
#define OSMESA_FORMAT 0x22
#define OSMESA_DEPTH_BITS 0x30
#define OSMESA_PROFILE 0x33
#define OSMESA_CORE_PROFILE 0x34
#define OSMESA_COMPAT_PROFILE 0x35
#define OSMESA_CONTEXT_MAJOR_VERSION 0x36
#define OSMESA_CONTEXT_MINOR_VERSION 0x37
#ifdef __cplusplus
extern "C" {
#endif
typedef struct osmesa_context *OSMesaContext;
OSMesaContext OSMesaCreateContextAttribs(const int *attribList, OSMesaContext sharelist);
GLboolean OSMesaMakeCurrent(OSMesaContext ctx, void *buffer, GLenum type, GLsizei width, GLsizei height);
void OSMesaDestroyContext(OSMesaContext ctx);
#ifdef __cplusplus
}
#endif
#endif
