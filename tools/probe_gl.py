#!/usr/bin/env python3
"""Probe the GPU box for anything that could execute the reference's GLSL vertex shader
(src/game/shader/Particle.glsl) to pin row A12/N2: EGL / OpenGL / OSMesa / Vulkan loaders and the
NVIDIA vendor libraries. Prints one JSON object; run under gpurun, output kept in profiles/."""
import ctypes
import ctypes.util
import glob
import json
import os
import subprocess


def main():
    out = {"find_library": {}, "files": [], "env": {}, "dlopen": {}}
    for n in ["EGL", "GL", "OpenGL", "OSMesa", "GLESv2", "vulkan", "EGL_nvidia", "GLX_nvidia", "nvidia-eglcore", "nvidia-glcore", "GLdispatch"]:
        out["find_library"][n] = ctypes.util.find_library(n)
    pats = ["*EGL*", "*libGL*", "*OSMesa*", "*GLES*", "*vulkan*", "*nvidia-egl*", "*nvidia-gl*", "*glvnd*", "*GLdispatch*", "*nvidia-glsi*", "*nvidia-tls*"]
    for d in ["/usr/lib/x86_64-linux-gnu", "/usr/lib64", "/usr/lib", "/usr/local/lib", "/usr/local/nvidia/lib64", "/usr/lib/x86_64-linux-gnu/nvidia"]:
        for p in pats:
            out["files"] += sorted(glob.glob(os.path.join(d, p)))
    out["files"] = sorted(set(out["files"]))
    for k in ["NVIDIA_DRIVER_CAPABILITIES", "NVIDIA_VISIBLE_DEVICES", "__EGL_VENDOR_LIBRARY_DIRS", "LD_LIBRARY_PATH"]:
        out["env"][k] = os.environ.get(k)
    try:
        out["ldconfig"] = [l.strip() for l in subprocess.run(["ldconfig", "-p"], capture_output=True, text=True).stdout.splitlines()
                           if any(t in l.lower() for t in ("egl", "libgl", "opengl", "vulkan", "osmesa", "glvnd"))]
    except OSError as ex:
        out["ldconfig"] = str(ex)
    out["egl_vendor_json"] = sorted(glob.glob("/usr/share/glvnd/egl_vendor.d/*") + glob.glob("/etc/glvnd/egl_vendor.d/*"))
    for name in ["libEGL.so.1", "libEGL_nvidia.so.0", "libGL.so.1", "libOpenGL.so.0", "libOSMesa.so.8", "libvulkan.so.1", "libnvidia-eglcore.so", "libnvidia-glcore.so"]:
        cands = [name] + [f for f in out["files"] if os.path.basename(f).startswith(name.split(".so")[0] + ".so")]
        res = None
        for c in cands:
            try:
                ctypes.CDLL(c)
                res = "loaded " + c
                break
            except OSError as ex:
                res = "error: " + str(ex)[:120]
        out["dlopen"][name] = res
    # if a loader is there, try to bring up a surfaceless display
    try:
        egl = ctypes.CDLL("libEGL.so.1")
        egl.eglGetDisplay.restype = ctypes.c_void_p
        dpy = egl.eglGetDisplay(ctypes.c_void_p(0))
        major, minor = ctypes.c_int(0), ctypes.c_int(0)
        ok = egl.eglInitialize(ctypes.c_void_p(dpy), ctypes.byref(major), ctypes.byref(minor)) if dpy else 0
        out["egl_init"] = {"display": bool(dpy), "ok": bool(ok), "version": [major.value, minor.value]}
    except OSError as ex:
        out["egl_init"] = "no loader: " + str(ex)[:120]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
