"""Build libhpc_rll_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a only.

    python -m di_hpc_b200.build [--force] [--verbose]

Every .cu under csrc/ is compiled to an object (in parallel, mtime-incremental) and linked into
``di_hpc_b200/_lib/libhpc_rll_b200.so``.  nvcc cross-compiles without a GPU, so this runs in the
CPU-only build container; the resulting .so travels to the GPU box with the repo snapshot.
"""
import concurrent.futures as cf
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
OBJ_DIR = os.path.join(OUT_DIR, "obj")
LIB = os.path.join(OUT_DIR, "libhpc_rll_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
          "--expt-relaxed-constexpr", "-I", INCLUDE]
# the system gcc; $CC in this image may point at a toolchain without the pieces nvcc expects
HOST = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []


def _deps_mtime():
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return max(os.path.getmtime(h) for h in hdrs) if hdrs else 0.0


def _compile(src, obj, verbose):
    cmd = [NVCC] + ARCH + CFLAGS + HOST + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return r.stderr if verbose else ""


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    if not srcs:
        raise RuntimeError("no CUDA sources under %s" % CSRC)
    hdr_m = _deps_mtime()
    jobs, objs = [], []
    for s in srcs:
        o = os.path.join(OBJ_DIR, os.path.basename(s)[:-3] + ".o")
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), hdr_m):
            jobs.append((s, o))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(lambda so: _compile(so[0], so[1], verbose), jobs):
                if verbose and log:
                    print(log)
    if jobs or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC] + ARCH + HOST + ["-shared", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
