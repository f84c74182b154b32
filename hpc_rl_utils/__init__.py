"""`hpc_rl_utils` -- the reference's NATIVE module name (/root/reference/setup.py:17-29, src/rl_utils/entry.cpp:8-39),
re-exporting the B200 shim: the 19 hot-path binding names (GaeForward ... QRDQNNStepTDErrorBackward) and the 11 padding
names with the reference's tensor-list signatures, implemented on libhpc_rll_b200.so (di_hpc_b200/csrc_torch/).
With the repo root on sys.path the UNMODIFIED reference wrappers (`hpc_rll/rl_utils/*.py`, which do
`import hpc_rl_utils`) run on the B200 kernels."""
from di_hpc_b200 import _ext as _loader

_m = _loader.require()
globals().update({k: getattr(_m, k) for k in dir(_m) if not k.startswith("_")})
__all__ = [k for k in dir(_m) if not k.startswith("_")]
