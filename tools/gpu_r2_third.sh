#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gae_gpu.py tests/test_runtime_gpu.py tests/test_legacy_shim_gpu.py -q -x 2>&1 | tail -8
timeout 300 python tools/probe_small.py 2>&1 | tee gpurun_out/probe_small.jsonl | tail -30
for skip in 0 1 3 5; do
  echo "== host pipeline, debug_skip=$skip rows=64"
  HPC_RLL_HOST_DEBUG_SKIP=$skip timeout 120 python tools/probe_e2e.py --rows 64 --child torch 2>&1 | tail -1
done
echo "== trace rows=64"; HPC_RLL_HOST_TRACE=1 HPC_RLL_HOST_CHUNK_ROWS=64 timeout 120 python tools/probe_e2e.py --child torch --iters 1 2>&1 | tail -20
echo "== trace ramp"; HPC_RLL_HOST_TRACE=1 timeout 120 python tools/probe_e2e.py --child torch --iters 1 2>&1 | tail -20
