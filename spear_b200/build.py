"""Build spear_b200/libspear_particles.so (CUDA, sm_100a only) in-tree with nvcc.

  -gencode arch=compute_100a,code=sm_100a   B200 only, no PTX for other targets
  -fmad=false                               bit-exact float parity with the reference's
                                            non-FMA x86 arithmetic
  -lineinfo                                 ncu source page maps to these files
The library travels to the GPU box with the gpurun snapshot; nothing is JIT-compiled.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libspear_particles.so")
SOURCES = ["kernels.cu", "sort.cu", "vertex_stage.cu", "billboards.cu", "particle_system.cpp", "descriptor_json.cpp", "c_api.cpp"]
HEADERS = ["launch.h", "host_pool.h", "device_types.h", "kernels.h", "host_math.h", "particle_system.h", os.path.join("..", "..", "include", "spear_particles.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off,-O2", "-Xptxas", "-v",
]


def up_to_date():
    if not os.path.isfile(OUT):
        return False
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("SPR_EXTRA_NVCC_FLAGS", "").split()   # e.g. -DSPR_SORT_PROFILE for an instrumented experiment build
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build")
    os.makedirs(bdir, exist_ok=True)
    for s in SOURCES:
        obj = os.path.join(bdir, s + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-x", "cu", "-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for s, p in procs:
        out, _ = p.communicate()
        log.append("== %s\n%s" % (s, out))
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % s)
    with open(os.path.join(bdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    link = [nvcc, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr)
        raise RuntimeError("link failed")
    if verbose:
        print("\n".join(log))
        print("built", OUT)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(OUT)
