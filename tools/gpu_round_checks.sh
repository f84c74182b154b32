#!/bin/bash
# Everything a round wants from ONE single-GPU gpurun call (about 8-10 GPU-minutes), results under gpurun_out/:
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/gpu_round_checks.sh'
# Read back here with tools/ops_report.py and tools/ncu_summary.py (see profiles/README.md).  A multi-GPU scaling
# point costs N x the box time -- budget it separately (an 8-GPU bench call was ~50 GPU-minutes in round 1).
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 100 --warmup 5 2> gpurun_out/bench.err > gpurun_out/bench.json
python tools/bench_ops.py --out gpurun_out/ops.jsonl > gpurun_out/ops.log 2>&1
python tools/ops_report.py gpurun_out/ops.jsonl gpurun_out/ops.md
python tools/bench_vs_ref_cuda.py > gpurun_out/vs_ref.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gae_.*_tma -s 4 -c 2 -f -o gpurun_out/prof_gae \
    python bench.py --steps 3 --warmup 1 > /dev/null 2>&1
bash tools/ncu_ops.sh
for tool in memcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_smoke.py > gpurun_out/sanitizer_$tool.txt 2>&1
  tail -2 gpurun_out/sanitizer_$tool.txt
done
echo round checks done
