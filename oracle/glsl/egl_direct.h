/* TEST INFRASTRUCTURE ONLY (oracle/): a minimal stand-in for the libglvnd EGL loader.
 *
 * The GPU boxes of this pool carry NVIDIA's GL vendor libraries (libEGL_nvidia.so.0, libnvidia-eglcore, ...)
 * but no libEGL.so.1 / libglvnd to dispatch to them (profiles/r02_probe_gl.json). A GLVND vendor library exports
 * one entry point, __egl_Main (libglvnd include/glvnd/libeglabi.h): the loader hands it a table of call-backs
 * ("exports": the current-context bookkeeping libglvnd keeps per thread) and receives a table of vendor functions
 * ("imports": getPlatformDisplay, getProcAddress, ...). This file plays the loader for ONE thread and ONE vendor:
 * enough to create a surfaceless OpenGL 3.3 core context on the B200 and call GL through the vendor's
 * getProcAddress, so that the reference's own GLSL (src/game/shader/Particle.glsl) can be executed with
 * transform feedback as the oracle of SURVEY.md row A12/N2. No GL headers exist in this image: the few types
 * and enums used are declared here (values from the Khronos registry).
 */
#ifndef EGL_DIRECT_H
#define EGL_DIRECT_H

#include <stdint.h>

typedef void *EGLDisplay, *EGLConfig, *EGLContext, *EGLSurface, *EGLDeviceEXT;
typedef unsigned int EGLBoolean, EGLenum;
typedef int32_t EGLint;
typedef intptr_t EGLAttrib;

#define EGL_NONE 0x3038
#define EGL_SURFACE_TYPE 0x3033
#define EGL_PBUFFER_BIT 0x0001
#define EGL_RENDERABLE_TYPE 0x3040
#define EGL_OPENGL_BIT 0x0008
#define EGL_OPENGL_API 0x30A2
#define EGL_PLATFORM_DEVICE_EXT 0x313F
#define EGL_CONTEXT_MAJOR_VERSION 0x3098
#define EGL_CONTEXT_MINOR_VERSION 0x30FB
#define EGL_CONTEXT_OPENGL_PROFILE_MASK 0x30FD
#define EGL_CONTEXT_OPENGL_CORE_PROFILE_BIT 0x00000001
#define EGL_WIDTH 0x3057
#define EGL_HEIGHT 0x3056
#define EGL_VENDOR 0x3053
#define EGL_VERSION 0x3054
#define EGL_EXTENSIONS 0x3055
#define EGL_SUCCESS 0x3000

typedef struct EglDirect {
	void *lib;
	void *(*getProcAddress)(const char *name);
	EGLDisplay display;
	EGLContext context;
	EGLSurface surface;
	int eglMajor, eglMinor;
	char error[256];
} EglDirect;

/* Opens the vendor library, performs the __egl_Main handshake, picks the first EGL device that initialises and
 * makes an OpenGL 3.3 core context current (surfaceless if the driver allows it, else on a 16x16 pbuffer).
 * Returns 0 on success; on failure `error` says which step failed. */
int egl_direct_open(EglDirect *e, const char *vendorLibrary, int verbose);
void egl_direct_close(EglDirect *e);

#endif
