"""Build the thin torch/pybind layer (di_hpc_b200/csrc_torch/ext.cpp) in-tree as
``di_hpc_b200/_lib/hpc_rl_utils_b200*.so``.  It contains no kernels: it links against
libhpc_rll_b200.so (rpath $ORIGIN) and only moves the list-of-tensor host work of the padding ops into C++.

    python -m di_hpc_b200.build_torch_ext
"""
import glob
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc_torch", "ext.cpp")
OUT_DIR = os.path.join(HERE, "_lib")
NAME = "hpc_rl_utils_b200"


def target():
    return os.path.join(OUT_DIR, NAME + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force: bool = False) -> str:
    from . import build as core
    lib = core.build()
    out = target()
    deps = [SRC, os.path.join(os.path.dirname(HERE), "include", "hpc_rll_b200.h")]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    import torch
    from torch.utils import cpp_extension as ce
    inc = ce.include_paths(device_type="cuda") + [os.path.join(os.path.dirname(HERE), "include"),
                                                  sysconfig.get_paths()["include"]]
    libdirs = ce.library_paths(device_type="cuda")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=" + NAME,
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI)]
    cmd += ["-I" + i for i in inc]
    cmd += [SRC, "-o", out]
    cmd += ["-L" + d for d in libdirs] + ["-L" + OUT_DIR]
    cmd += ["-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lhpc_rll_b200",
            "-Wl,-rpath,$ORIGIN"] + ["-Wl,-rpath," + d for d in libdirs]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building %s failed:\n%s\n%s" % (NAME, r.stdout[-3000:], r.stderr[-6000:]))
    assert os.path.exists(lib)
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
