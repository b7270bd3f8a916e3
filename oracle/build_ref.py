#!/usr/bin/env python3
"""Build oracle/_ref/libspear_ref{,_scalar}.so -- TEST INFRASTRUCTURE ONLY.

Compiles the reference's own hot path (src/client/ParticleSystem.cpp, included
by oracle/ref_harness.cpp) and the few reference units it links against,
straight from the read-only tree at /root/reference.
Nothing from the reference is copied into this repository: the two header
shims the g++ build needs (the reference targets clang) are GENERATED into a
temporary directory outside the repo and deleted afterwards; only the shared
libraries are written, into oracle/_ref/ (git-ignored, shipped by gpurun).

Shims (neither touches the hot path):
  * sf/Geometry.h:14-28  FastRay keeps sf::Vec3 (non-trivial ctor) inside an
    anonymous union -> hard error in g++; the shim flattens the union.
  * sf/ext/mx/mx_platform.h selects mx_platform_posix.h, which uses clang-only
    __c11_atomic_*; the shim selects mx_platform_singlethreaded.h.

Flags mirror premake5.lua:117,:143-149 plus -ffp-contract=off (SURVEY.md
Appendix C: FMA contraction changes float results). Two variants are built:
  libspear_ref.so         SSE4.1 sf::Float4 (what the game ships on x86)
  libspear_ref_scalar.so  -DSF_FLOAT4_FORCE_SCALAR=1 (IEEE 1/sqrt in rsqrt())
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SPEAR_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
ADAPTER_SRC = os.path.join(HERE, "..", "integration", "ParticleSystemB200.cpp")
PRODUCT_DIR = os.path.join(HERE, "..", "spear_b200")

BASE_FLAGS = [
    "-O2", "-DNDEBUG", "-DSP_NO_APP=1", "-msse4.1", "-fno-math-errno",
    "-ffp-contract=off", "-fPIC", "-w",
]
CXX_FLAGS = ["-std=c++14", "-Wno-invalid-offsetof", "-fno-exceptions", "-fno-rtti"]


def available():
    return os.path.isfile(os.path.join(REF, "src", "client", "ParticleSystem.cpp"))


def _make_shim_tree(tmp):
    """Mirror src/sf (same-directory includes force this) with two headers patched."""
    mirror = os.path.join(tmp, "mirror")
    shutil.copytree(os.path.join(REF, "src", "sf"), os.path.join(mirror, "sf"))
    for root, _dirs, files in os.walk(mirror):
        os.chmod(root, 0o755)
        for f in files:
            os.chmod(os.path.join(root, f), 0o644)
    geo = os.path.join(mirror, "sf", "Geometry.h")
    src = open(geo).read()
    start = src.index("struct FastRay")
    ustart = src.index("union {", start)
    uend = src.index("};", src.index("};", ustart) + 2) + 2
    src = src[:ustart] + "sf::Vec3 origin;\n\tsf::Vec3 direction;\n" + src[uend:]
    open(geo, "w").write(src)
    mx = os.path.join(mirror, "sf", "ext", "mx", "mx_platform.h")
    open(mx, "w").write('#pragma once\n#include "mx_config.h"\n#include "mx_platform_singlethreaded.h"\n')
    return mirror


def _compile(args):
    cmd, obj = args
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("compile failed: %s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
    return obj


def build(force=False, verbose=True):
    targets = [os.path.join(OUT, "libspear_ref.so"), os.path.join(OUT, "libspear_ref_scalar.so"), os.path.join(OUT, "libspear_ref_billboard.so"),
               os.path.join(OUT, "libspear_ref_json.so"), os.path.join(OUT, "libspear_ref_area.so"), os.path.join(OUT, "libspear_adapter.so")]
    srcs = [os.path.join(HERE, "ref_harness.cpp"), os.path.join(HERE, "ref_billboard_harness.cpp"), os.path.join(HERE, "ref_json_harness.cpp"), os.path.join(HERE, "ref_area_harness.cpp"),
            os.path.join(HERE, "oracle_api.h"),
            os.path.join(HERE, "..", "include", "spear_particles.h"), os.path.abspath(__file__), ADAPTER_SRC]
    prod = os.path.join(PRODUCT_DIR, "libspear_particles.so")
    if os.path.isfile(prod):
        srcs.append(prod)
    else:
        targets = targets[:5]
    if not force and all(os.path.isfile(t) for t in targets):
        newest = max(os.path.getmtime(s) for s in srcs)
        if all(os.path.getmtime(t) >= newest for t in targets):
            return targets
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF)
    os.makedirs(OUT, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="spear_ref_build_")
    try:
        mirror = _make_shim_tree(tmp)
        inc = ["-I" + mirror, "-I" + os.path.join(REF, "src"), "-isystem", os.path.join(REF, "src", "ext"),
               "-I" + HERE, "-I" + os.path.join(HERE, "..", "include")]
        sf_units = sorted(f for f in os.listdir(os.path.join(mirror, "sf")) if f.endswith(".cpp"))
        # File/Process/Thread/License/Reflection are not needed by the path; skip what does not link cleanly.
        skip = {"File.cpp", "Process.cpp", "License.cpp"}
        units = [(os.path.join(mirror, "sf", f), True) for f in sf_units if f not in skip]
        units.append((os.path.join(mirror, "sf", "ext", "rhmap.cpp"), True))
        units.append((os.path.join(mirror, "sf", "ext", "mx", "mx_sync.c"), False))
        units.append((os.path.join(mirror, "sf", "ext", "mx", "mx_platform.c"), False))
        units.append((os.path.join(REF, "src", "client", "BSpline.cpp"), True))
        units.append((os.path.join(REF, "src", "client", "Transform.cpp"), True))
        units.append((os.path.join(REF, "src", "sp", "Srgb.cpp"), True))
        units.append((os.path.join(HERE, "ref_harness.cpp"), True))
        for variant, extra in (("", []), ("_scalar", ["-DSF_FLOAT4_FORCE_SCALAR=1"]), ("_billboard", []), ("_json", []), ("_area", []), ("_adapter", ["-DSPR_ADAPTER=1"])):
            jobs = []
            vunits = list(units)
            if variant == "_billboard":
                # cl::BillboardSystem (src/client/BillboardSystem.cpp) behind its own small harness
                vunits = [u for u in units if not u[0].endswith(("ref_harness.cpp", "BSpline.cpp", "Transform.cpp", "Srgb.cpp"))]
                vunits.append((os.path.join(HERE, "ref_billboard_harness.cpp"), True))
            if variant == "_json":
                # the reference's descriptor path: lenient JSON reader + reflection tables + sp::readJson / writeJson
                vunits = [u for u in units if not u[0].endswith(("ref_harness.cpp", "BSpline.cpp", "Transform.cpp", "Srgb.cpp"))]
                for rel in (("server", "ServerStateReflection.cpp"), ("sp", "Json.cpp"), ("ext", "json_input_impl.cpp"), ("ext", "json_output.cpp")):
                    vunits.append((os.path.join(REF, "src", *rel), True))
                vunits.append((os.path.join(HERE, "ref_json_harness.cpp"), True))
            if variant == "_area":
                # the reference's cl::AreaSystem (src/client/AreaSystem.cpp) behind its own small harness
                vunits = [u for u in units if not u[0].endswith(("ref_harness.cpp", "BSpline.cpp", "Transform.cpp", "Srgb.cpp"))]
                vunits.append((os.path.join(REF, "src", "client", "AreaSystem.cpp"), True))
                vunits.append((os.path.join(HERE, "ref_area_harness.cpp"), True))
            if variant == "_adapter":
                # the literal drop-in: integration/ParticleSystemB200.cpp (cl::ParticleSystem over the C ABI) instead
                # of the reference's ParticleSystem.cpp, linked against the product library
                if not os.path.isfile(os.path.join(PRODUCT_DIR, "libspear_particles.so")):
                    continue
                vunits.append((ADAPTER_SRC, True))
            for i, (path, is_cxx) in enumerate(vunits):
                obj = os.path.join(tmp, "o%s_%d.o" % (variant, i))
                cc = ["g++"] + CXX_FLAGS if is_cxx else ["gcc"]
                vis = ["-fvisibility=hidden"] if path.endswith(("ref_harness.cpp", "ref_billboard_harness.cpp", "ref_json_harness.cpp", "ref_area_harness.cpp")) else []
                jobs.append((cc + BASE_FLAGS + extra + vis + inc + ["-c", path, "-o", obj], obj))
            with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
                objs = list(ex.map(_compile, jobs))
            so = os.path.join(OUT, "libspear_adapter.so" if variant == "_adapter" else "libspear_ref%s.so" % variant)
            link = ["g++", "-shared", "-o", so] + objs + ["-lpthread", "-Wl,--no-undefined"]
            if variant == "_adapter":
                link += ["-L" + PRODUCT_DIR, "-lspear_particles", "-Wl,-rpath,$ORIGIN/../../spear_b200"]
            r = subprocess.run(link, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("link failed:\n" + r.stderr[-6000:])
            if verbose:
                print("built", so)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return targets


if __name__ == "__main__":
    build(force="--force" in sys.argv)
