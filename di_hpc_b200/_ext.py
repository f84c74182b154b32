"""Loader of the torch/pybind layer ``_lib/hpc_rl_utils_b200*.so`` (csrc_torch/: C++ autograd functions behind the
nn.Modules, the reference's 30 ``hpc_rl_utils`` binding names, the padding list handling).

Two bindings of the SAME CUDA library exist: this C++ layer (default; ~10x less host time per call) and the ctypes
binding in ``_abi.py`` driven by the Python ``*Function`` classes.  ``HPC_RLL_BINDING=ctypes`` forces the latter (it is
also what runs when the extension was built against a different torch).  Neither is a fallback in the sense of the
product rules: both end in libhpc_rll_b200.so; without that library everything raises.
"""
import glob
import importlib.util
import os
import sys

from ._abi import LIB_PATH, HpcRllError

_NAME = "hpc_rl_utils_b200"
_mod = None
_err = None


def load():
    """The extension module, or None (reason in ``why_missing()``)."""
    global _mod, _err
    if _mod is not None or _err is not None:
        return _mod
    if _NAME in sys.modules:
        _mod = sys.modules[_NAME]
        return _mod
    cands = glob.glob(os.path.join(os.path.dirname(LIB_PATH), _NAME + "*.so"))
    if not cands:
        _err = "not built (run `python -m di_hpc_b200.build_torch_ext` or __graft_entry__.build())"
        return None
    try:
        import torch  # noqa: F401  (the extension links against libtorch)
        spec = importlib.util.spec_from_file_location(_NAME, cands[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules[_NAME] = mod
        _mod = mod
    except Exception as e:  # noqa: BLE001  e.g. built against another torch: the ctypes binding still works
        _err = "%s: %s" % (type(e).__name__, e)
    return _mod


def why_missing():
    return _err


def fast():
    """The extension if the C++ binding is selected and loadable, else None (callers then use the ctypes twin)."""
    if os.environ.get("HPC_RLL_BINDING", "ext") == "ctypes":
        return None
    return load()


def require():
    m = load()
    if m is None:
        raise HpcRllError("torch extension %s unavailable: %s" % (_NAME, _err))
    return m
