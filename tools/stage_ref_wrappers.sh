#!/bin/bash
# Stage the UNMODIFIED reference Python wrappers (hpc_rll/rl_utils/{gae,td,upgo,vtrace,ppo}.py) under the git-ignored
# baseline/_ref/ref_wrappers/ so that they travel to the GPU box (there is no /root/reference there) and
# tests/test_legacy_shim_gpu.py can run them on top of the `hpc_rl_utils` shim.  Nothing in the product reads them.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
DST="$ROOT/baseline/_ref/ref_wrappers"
mkdir -p "$DST"
for f in gae td upgo vtrace ppo; do cp /root/reference/hpc_rll/rl_utils/$f.py "$DST/$f.py"; done
ls -la "$DST"
