"""Load the UNMODIFIED reference wrappers (hpc_rll/rl_utils/*.py of DI-hpc) under alias module names, with
`import hpc_rl_utils` inside them resolving to this repo's shim (hpc_rl_utils/__init__.py -> csrc_torch/legacy.cpp).
Sources: /root/reference (build container) or baseline/_ref/ref_wrappers (staged by tools/stage_ref_wrappers.sh; the
only copy that exists on the GPU box).  Never the product's own hpc_rll package."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = [os.path.join(ROOT, "baseline", "_ref", "ref_wrappers"), "/root/reference/hpc_rll/rl_utils"]


def wrapper_dir():
    for d in CANDIDATES:
        if os.path.exists(os.path.join(d, "vtrace.py")):
            return d
    return None


def load(name):
    d = wrapper_dir()
    if d is None:
        return None
    alias = "ref_wrapper_" + name
    if alias in sys.modules:
        return sys.modules[alias]
    spec = importlib.util.spec_from_file_location(alias, os.path.join(d, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[alias] = mod
    spec.loader.exec_module(mod)
    return mod
