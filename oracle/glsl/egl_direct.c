/* TEST INFRASTRUCTURE ONLY (oracle/): see egl_direct.h. */
#include "egl_direct.h"

#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

/* ---- the loader side of libglvnd's EGL vendor ABI (libeglabi.h, ABI version 0.1) ---- */
typedef struct VendorInfo { int dummy; } VendorInfo;
typedef void (*AnyFn)(void);

typedef struct ApiExports {
	void (*threadInit)(void);
	EGLenum (*getCurrentApi)(void);
	VendorInfo *(*getCurrentVendor)(void);
	EGLContext (*getCurrentContext)(void);
	EGLDisplay (*getCurrentDisplay)(void);
	EGLSurface (*getCurrentSurface)(EGLint readDraw);
	AnyFn (*fetchDispatchEntry)(VendorInfo *vendor, int index);
	void (*setEGLError)(EGLint errorCode);
	EGLBoolean (*setLastVendor)(VendorInfo *vendor);
	VendorInfo *(*getVendorFromDisplay)(EGLDisplay dpy);
	VendorInfo *(*getVendorFromDevice)(EGLDeviceEXT dev);
	void (*setVendorForDevice)(EGLDeviceEXT dev, VendorInfo *vendor);
	void *reserved[8]; /* room should a newer minor version read further entries (all NULL) */
} ApiExports;

typedef struct ApiImports {
	EGLDisplay (*getPlatformDisplay)(EGLenum platform, void *nativeDisplay, const EGLAttrib *attribs);
	EGLBoolean (*getSupportsAPI)(EGLenum api);
	const char *(*getVendorString)(int name);
	void *(*getProcAddress)(const char *procName);
	void *(*getDispatchAddress)(const char *procName);
	void (*setDispatchIndex)(const char *procName, int index);
	EGLBoolean (*isPatchSupported)(int type, int stubSize);
	EGLBoolean (*initiatePatch)(int type, int stubSize, void *lookupStubOffset);
	void (*releasePatch)(void);
	void (*patchThreadAttach)(void);
	EGLenum (*findNativeDisplayPlatform)(void *nativeDisplay);
	void *reserved[16]; /* the vendor may fill entries of a newer minor version */
} ApiImports;

typedef EGLBoolean (*EglMainFn)(uint32_t version, const ApiExports *exports, VendorInfo *vendor, ApiImports *imports);

/* the per-thread state libglvnd would keep (one thread here) */
static VendorInfo g_vendor;
static EGLenum g_api = 0x30A0 /* EGL_OPENGL_ES_API, the initial value */;
static EGLContext g_ctx;
static EGLDisplay g_dpy;
static EGLSurface g_draw, g_read;
static EGLint g_lastError = EGL_SUCCESS;
static int g_verbose;

static void x_threadInit(void) { }
static EGLenum x_getCurrentApi(void) { return g_api; }
static VendorInfo *x_getCurrentVendor(void) { return g_ctx ? &g_vendor : NULL; }
static EGLContext x_getCurrentContext(void) { return g_ctx; }
static EGLDisplay x_getCurrentDisplay(void) { return g_ctx ? g_dpy : NULL; }
static EGLSurface x_getCurrentSurface(EGLint readDraw) { return readDraw == 0x305A /* EGL_READ */ ? g_read : g_draw; }
static AnyFn x_fetchDispatchEntry(VendorInfo *v, int index) { (void)v; (void)index; return NULL; }
static void x_setEGLError(EGLint code) { g_lastError = code; if (g_verbose && code != EGL_SUCCESS) fprintf(stderr, "[egl_direct] vendor set error 0x%x\n", code); }
static EGLBoolean x_setLastVendor(VendorInfo *v) { (void)v; return 1; }
static VendorInfo *x_getVendorFromDisplay(EGLDisplay d) { (void)d; return &g_vendor; }
static VendorInfo *x_getVendorFromDevice(EGLDeviceEXT d) { (void)d; return &g_vendor; }
static void x_setVendorForDevice(EGLDeviceEXT d, VendorInfo *v) { (void)d; (void)v; }

static ApiExports g_exports;
static ApiImports g_imports;

#define FAIL(...) do { snprintf(e->error, sizeof(e->error), __VA_ARGS__); return -1; } while (0)
#define EGLFN(ret, name, ...) ret (*name)(__VA_ARGS__) = (ret (*)(__VA_ARGS__))g_imports.getProcAddress(#name); if (!name) FAIL("vendor getProcAddress(%s) = NULL", #name)

int egl_direct_open(EglDirect *e, const char *vendorLibrary, int verbose)
{
	memset(e, 0, sizeof(*e));
	g_verbose = verbose;
	e->lib = dlopen(vendorLibrary ? vendorLibrary : "libEGL_nvidia.so.0", RTLD_NOW | RTLD_GLOBAL);
	if (!e->lib) FAIL("dlopen: %s", dlerror());
	EglMainFn eglMain = (EglMainFn)dlsym(e->lib, "__egl_Main");
	if (!eglMain) FAIL("no __egl_Main in the vendor library");
	memset(&g_exports, 0, sizeof(g_exports));
	memset(&g_imports, 0, sizeof(g_imports));
	g_exports.threadInit = x_threadInit; g_exports.getCurrentApi = x_getCurrentApi; g_exports.getCurrentVendor = x_getCurrentVendor;
	g_exports.getCurrentContext = x_getCurrentContext; g_exports.getCurrentDisplay = x_getCurrentDisplay; g_exports.getCurrentSurface = x_getCurrentSurface;
	g_exports.fetchDispatchEntry = x_fetchDispatchEntry; g_exports.setEGLError = x_setEGLError; g_exports.setLastVendor = x_setLastVendor;
	g_exports.getVendorFromDisplay = x_getVendorFromDisplay; g_exports.getVendorFromDevice = x_getVendorFromDevice; g_exports.setVendorForDevice = x_setVendorForDevice;
	EGLBoolean ok = 0;
	uint32_t versions[3] = { (0u << 16) | 1u, (0u << 16) | 0u, (0u << 16) | 2u };
	for (int i = 0; i < 3 && !ok; i++) {
		ok = eglMain(versions[i], &g_exports, &g_vendor, &g_imports);
		if (verbose) fprintf(stderr, "[egl_direct] __egl_Main(version 0x%x) -> %u\n", versions[i], ok);
	}
	if (!ok) FAIL("__egl_Main refused every ABI version tried");
	if (!g_imports.getProcAddress || !g_imports.getPlatformDisplay) FAIL("vendor did not fill getProcAddress / getPlatformDisplay");
	e->getProcAddress = g_imports.getProcAddress;
	if (verbose && g_imports.getVendorString) {
		const char *s = g_imports.getVendorString(0 /* __EGL_VENDOR_STRING_PLATFORM_EXTENSIONS */);
		fprintf(stderr, "[egl_direct] vendor platform extensions: %s\n", s ? s : "(null)");
	}

	EGLFN(EGLBoolean, eglQueryDevicesEXT, EGLint, EGLDeviceEXT *, EGLint *);
	EGLFN(EGLBoolean, eglInitialize, EGLDisplay, EGLint *, EGLint *);
	EGLFN(EGLBoolean, eglBindAPI, EGLenum);
	EGLFN(EGLBoolean, eglChooseConfig, EGLDisplay, const EGLint *, EGLConfig *, EGLint, EGLint *);
	EGLFN(EGLContext, eglCreateContext, EGLDisplay, EGLConfig, EGLContext, const EGLint *);
	EGLFN(EGLSurface, eglCreatePbufferSurface, EGLDisplay, EGLConfig, const EGLint *);
	EGLFN(EGLBoolean, eglMakeCurrent, EGLDisplay, EGLSurface, EGLSurface, EGLContext);
	EGLFN(const char *, eglQueryString, EGLDisplay, EGLint);

	EGLDeviceEXT devices[32];
	EGLint numDevices = 0;
	EGLBoolean qok = eglQueryDevicesEXT(32, devices, &numDevices);
	if (verbose) fprintf(stderr, "[egl_direct] eglQueryDevicesEXT -> %u, %d device(s), last error 0x%x\n", qok, numDevices, g_lastError);
	if (!qok || numDevices <= 0) {
		/* no enumerable device: try the vendor's surfaceless platform and its default display before giving up */
		EGLenum platforms[2] = { 0x31DD /* EGL_PLATFORM_SURFACELESS_MESA */, EGL_PLATFORM_DEVICE_EXT };
		for (int p = 0; p < 2 && !e->display; p++) {
			EGLDisplay dpy = g_imports.getPlatformDisplay(platforms[p], NULL, NULL);
			if (verbose) fprintf(stderr, "[egl_direct] getPlatformDisplay(0x%x, default) -> %p (error 0x%x)\n", platforms[p], dpy, g_lastError);
			if (dpy && eglInitialize(dpy, &e->eglMajor, &e->eglMinor)) e->display = dpy;
			else if (dpy && verbose) fprintf(stderr, "[egl_direct] eglInitialize failed (0x%x)\n", g_lastError);
		}
		if (!e->display) FAIL("eglQueryDevicesEXT found no device and no default display initialised (error 0x%x)", g_lastError);
	}
	for (EGLint d = 0; d < numDevices && !e->display; d++) {
		EGLDisplay dpy = g_imports.getPlatformDisplay(EGL_PLATFORM_DEVICE_EXT, devices[d], NULL);
		if (!dpy) continue;
		if (eglInitialize(dpy, &e->eglMajor, &e->eglMinor)) e->display = dpy;
		else if (verbose) fprintf(stderr, "[egl_direct] device %d: eglInitialize failed (0x%x)\n", d, g_lastError);
	}
	if (!e->display) FAIL("no EGL device initialised (last error 0x%x)", g_lastError);
	g_dpy = e->display;
	if (verbose) fprintf(stderr, "[egl_direct] EGL %d.%d, vendor %s\n", e->eglMajor, e->eglMinor, eglQueryString(e->display, EGL_VENDOR));
	if (!eglBindAPI(EGL_OPENGL_API)) FAIL("eglBindAPI(EGL_OPENGL_API) failed (0x%x)", g_lastError);
	g_api = EGL_OPENGL_API;
	const EGLint cfgAttribs[] = { EGL_SURFACE_TYPE, EGL_PBUFFER_BIT, EGL_RENDERABLE_TYPE, EGL_OPENGL_BIT, EGL_NONE };
	EGLConfig cfg = NULL;
	EGLint numCfg = 0;
	if (!eglChooseConfig(e->display, cfgAttribs, &cfg, 1, &numCfg) || numCfg < 1) FAIL("eglChooseConfig found nothing (0x%x)", g_lastError);
	const EGLint ctxAttribs[] = { EGL_CONTEXT_MAJOR_VERSION, 3, EGL_CONTEXT_MINOR_VERSION, 3,
		EGL_CONTEXT_OPENGL_PROFILE_MASK, EGL_CONTEXT_OPENGL_CORE_PROFILE_BIT, EGL_NONE };
	e->context = eglCreateContext(e->display, cfg, NULL, ctxAttribs);
	if (!e->context) FAIL("eglCreateContext(OpenGL 3.3 core) failed (0x%x)", g_lastError);
	const EGLint pbAttribs[] = { EGL_WIDTH, 16, EGL_HEIGHT, 16, EGL_NONE };
	e->surface = eglCreatePbufferSurface(e->display, cfg, pbAttribs);
	if (!eglMakeCurrent(e->display, e->surface, e->surface, e->context)) FAIL("eglMakeCurrent failed (0x%x)", g_lastError);
	g_ctx = e->context; g_draw = g_read = e->surface;
	return 0;
}

void egl_direct_close(EglDirect *e)
{
	if (!e->lib || !g_imports.getProcAddress) return;
	EGLBoolean (*eglMakeCurrent)(EGLDisplay, EGLSurface, EGLSurface, EGLContext) = (EGLBoolean (*)(EGLDisplay, EGLSurface, EGLSurface, EGLContext))g_imports.getProcAddress("eglMakeCurrent");
	EGLBoolean (*eglTerminate)(EGLDisplay) = (EGLBoolean (*)(EGLDisplay))g_imports.getProcAddress("eglTerminate");
	if (eglMakeCurrent && e->display) eglMakeCurrent(e->display, NULL, NULL, NULL);
	g_ctx = NULL;
	if (eglTerminate && e->display) eglTerminate(e->display);
	e->display = NULL;
}
