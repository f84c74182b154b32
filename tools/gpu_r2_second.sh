#!/bin/bash
# round 2, second GPU call: full parity suite (legacy shim + fast path), new bench line, e2e probe with the ramp schedule
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 300 python tools/probe_e2e.py --rows auto,8:128,16:128,8:256,4:64,16:256 > gpurun_out/probe_e2e.jsonl 2> gpurun_out/probe_e2e.err
cat gpurun_out/probe_e2e.jsonl
echo second call done
