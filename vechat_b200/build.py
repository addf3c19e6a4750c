"""In-tree builds (no JIT cache): the CUDA engine libvgc.so (sm_100a) and the host-only simulator.

`python -m vechat_b200.build` builds everything; __graft_entry__.build() calls build_all().
The built .so files are git-ignored but travel to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
INC = os.path.join(ROOT, "include")
LIB_DIR = os.path.join(HERE, "lib")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
CXX = os.environ.get("CXX") or shutil.which("g++") or "g++"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)


def engine_sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))
            ] + [os.path.join(INC, "vgc.h"), os.path.join(INC, "vga.h")]


def build_engine(force=False, verbose=False):
    """libvgc.so: the C-ABI + hand-written sm_100a kernels."""
    os.makedirs(LIB_DIR, exist_ok=True)
    out = os.path.join(LIB_DIR, "libvgc.so")
    srcs = [os.path.join(CSRC, "vgc_engine.cu"), os.path.join(CSRC, "ovl_align.cu")]
    if force or _stale(out, engine_sources()):
        cmd = [NVCC] + NVCC_FLAGS + ["-shared", "-I", INC, "-I", CSRC, "-o", out] + srcs + ["-lcudart", "-lpthread"]
        cmd += os.environ.get("VGC_NVCC_EXTRA", "").split()
        if verbose:
            cmd += ["-Xptxas", "-v"]
        _run(cmd)
    return out


def build_host(force=False):
    """libvgchost.so: the C++ host side above the C-ABI (csrc/host/vgc_host.hpp: createWindow / add_layer /
    B200Polisher::polish) behind a flat C test entry; links libvgc.so."""
    out = os.path.join(LIB_DIR, "libvgchost.so")
    hdir = os.path.join(CSRC, "host")
    srcs = [os.path.join(hdir, "vgc_host_capi.cpp")]
    deps = srcs + [os.path.join(hdir, "vgc_host.hpp"), os.path.join(INC, "vgc.h")]
    if force or _stale(out, deps):
        _run([CXX, "-std=c++14", "-O2", "-fPIC", "-shared", "-I", INC, "-I", hdir, "-o", out] + srcs +
             ["-L", LIB_DIR, "-lvgc", "-lpthread", "-Wl,-rpath,$ORIGIN"])
    return out


def build_sim(force=False):
    """libvgcsim.so: synthetic reads / ground-truth overlaps / windowizer (host only)."""
    os.makedirs(LIB_DIR, exist_ok=True)
    out = os.path.join(LIB_DIR, "libvgcsim.so")
    src = os.path.join(CSRC, "sim.cpp")
    if force or _stale(out, [src, os.path.join(INC, "vgc.h")]):
        _run([CXX, "-std=c++14", "-O2", "-fPIC", "-shared", "-I", INC, "-o", out, src])
    return out


def build_oracle():
    """The checkers under oracle/ (test infrastructure): our CPU restatement and, when the reference
    tree is present, the compiled reference itself."""
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "edlib"])
    if os.path.isdir("/root/reference/src"):
        _run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
        _run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "racon"])


def build_racon_binding():
    """vechat_racon_b200: the UNMODIFIED reference program (main.cpp, Polisher::initialize, parsers, edlib step) linked
    with csrc/racon_binding (racon::B200Polisher, the Polisher subclass a maintainer adds) and libvgc.so.  Needs the
    reference tree for its sources and headers, so it is built in the development container only; the binary travels
    to the GPU box with the snapshot like every other built file."""
    out = os.path.join(LIB_DIR, "vechat_racon_b200")
    if os.path.isdir("/root/reference/src"):
        _run(["make", "-s", "-j8", "-C", os.path.join(ROOT, "oracle"), "racon_b200"])
    return out if os.path.exists(out) else None


def build_all(force=False):
    build_sim(force)
    build_engine(force)
    build_host(force)
    build_oracle()
    build_racon_binding()


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
