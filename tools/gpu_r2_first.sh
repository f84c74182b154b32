#!/bin/bash
# round 2, first GPU call: parity (new tie / chunk / host tests), sanitizers on the TMA-store + chunk + host kernels,
# e2e probe (PCIe ceilings, chunk heights, NUMA-local vs torch pinned memory)
set -u
mkdir -p gpurun_out
nproc > gpurun_out/box.txt; lscpu | head -25 >> gpurun_out/box.txt; nvidia-smi -L >> gpurun_out/box.txt
nvidia-smi topo -m >> gpurun_out/box.txt 2>&1
cat /sys/devices/system/node/node*/cpulist >> gpurun_out/box.txt 2>&1
python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/probe_e2e.py > gpurun_out/probe_e2e.jsonl 2> gpurun_out/probe_e2e.err
cat gpurun_out/probe_e2e.jsonl
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_smoke.py > gpurun_out/sanitizer_$tool.txt 2>&1
  tail -3 gpurun_out/sanitizer_$tool.txt
done
echo first call done
