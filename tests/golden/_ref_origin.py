"""Load the reference's pure-PyTorch oracle (`/root/reference/hpc_rll/origin`) under the
alias package name ``ref_origin`` so it never collides with this repo's own ``hpc_rll``
drop-in package.

Only usable in the build container (where /root/reference is mounted).  It is used by
``make_golden.py`` to produce the committed golden fixtures and by optional CPU tests that
are skipped when the reference is absent.  Nothing under ``-m gpu``, ``bench.py`` or
``smoke()`` imports this file.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("HPC_RLL_REFERENCE", "/root/reference")
_ORIGIN_DIR = os.path.join(REF_ROOT, "hpc_rll", "origin")


def available() -> bool:
    return os.path.isdir(_ORIGIN_DIR)


def load():
    """Return the alias package; submodules gae, td, vtrace, upgo, ppo are attributes."""
    if "ref_origin" in sys.modules:
        return sys.modules["ref_origin"]
    if not available():
        raise RuntimeError("reference origin not found at %s" % _ORIGIN_DIR)
    pkg = types.ModuleType("ref_origin")
    pkg.__path__ = [_ORIGIN_DIR]
    sys.modules["ref_origin"] = pkg
    for name in ("gae", "td", "vtrace", "upgo", "ppo"):
        spec = importlib.util.spec_from_file_location(
            "ref_origin." + name, os.path.join(_ORIGIN_DIR, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules["ref_origin." + name] = mod
        spec.loader.exec_module(mod)
        setattr(pkg, name, mod)
    return pkg
