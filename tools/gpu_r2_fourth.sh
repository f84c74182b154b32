#!/bin/bash
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.txt
timeout 300 python tools/probe_small.py --shapes 1024x64,1024x256,1024x1024,1024x4096,512x512,2048x128,4096x64 2>&1 | tee gpurun_out/probe_small.jsonl | tail -40
timeout 200 python tools/probe_e2e.py --rows auto,32,128 > gpurun_out/probe_e2e.jsonl 2> gpurun_out/probe_e2e.err; cat gpurun_out/probe_e2e.jsonl
