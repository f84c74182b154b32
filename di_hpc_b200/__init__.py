"""di_hpc_b200 -- B200-native (sm_100a) re-implementation of DI-hpc's trajectory-return hot path.

The package holds only what that path needs:
  csrc/      hand-written CUDA kernels + the C ABI (include/hpc_rll_b200.h)
  _abi.py    ctypes binding of libhpc_rll_b200.so (fails loudly if the library is missing)
  rl_utils/  host-side mirror of the reference's `hpc_rll.rl_utils` modules (same class names,
             constructor arguments and forward signatures); `hpc_rll/` at the repo root re-exports it
             under the reference's import paths so DI-engine can swap it in unchanged.
There is no CPU fallback and no multi-backend dispatch.
"""
__version__ = "0.1.0"
