"""Build libvkvg_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m vkvg_b200.build            # build if sources are newer than the library
    python -m vkvg_b200.build --force
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvkvg_b200.so")
SOURCES = ["pipeline.cu", "flatten.cu", "stroke.cu", "raster.cu", "decode.cu", "vkvg_api.cpp", "svg.cpp", "png.cpp"]
# --fmad=false: tessellation and paint arithmetic must round exactly like the reference's baseline x86-64
# build (no FMA contraction) and like oracle/ (-ffp-contract=off).
COMPILE_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--fmad=false", "-std=c++17",
                 "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function,-Wno-unused-variable", "-x", "cu"]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", f) for f in ("vkvg.h", "vkvg-svg.h", "vkvg_b200.h")]
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False):
    """Compile every translation unit (in parallel) and link them into vkvg_b200/libvkvg_b200.so."""
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src + ".o")
        cmd = [nvcc] + COMPILE_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, src)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs, failed = [], False
    for src, obj, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            failed = True
        elif verbose:
            sys.stderr.write(out)
        objs.append(obj)
    if failed:
        raise RuntimeError("nvcc failed building libvkvg_b200.so")
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-o", OUT] + objs + ["-lz"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed for libvkvg_b200.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
