#!/bin/bash
# Build the UNMODIFIED reference CUDA extension (hpc_rl_utils) for sm_100 from a scratch copy of
# /root/reference and place only the built module under baseline/_ref/ (git-ignored; travels to the GPU box).
# Used solely by tools/bench_vs_ref_cuda.py for a same-box comparison; nothing in the product depends on it.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
rm -rf /tmp/refbuild && cp -r /root/reference /tmp/refbuild
cd /tmp/refbuild
TORCH_CUDA_ARCH_LIST=10.0 MAX_JOBS=6 python3 setup.py build_ext --inplace > build.log 2>&1
mkdir -p "$ROOT/baseline/_ref"
cp hpc_rl_utils*.so "$ROOT/baseline/_ref/"
ls -la "$ROOT/baseline/_ref/"
