#!/bin/bash
# Round 2: everything a round wants from ONE single-GPU gpurun call (about 10 GPU-minutes), results under gpurun_out/:
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round_checks_r2.sh'
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 50 --warmup 5 2> gpurun_out/bench.err > gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2> gpurun_out/bench_ref.err > gpurun_out/bench_ref.json
timeout 600 python tools/bench_ops.py --out gpurun_out/ops.jsonl > gpurun_out/ops.log 2>&1
python tools/ops_report.py gpurun_out/ops.jsonl gpurun_out/ops.md
timeout 600 python tools/bench_vs_ref_cuda.py > gpurun_out/vs_ref.log 2>&1
timeout 120 python tools/probe_write_bw.py > gpurun_out/write_bw.jsonl 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-extras > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:gae_.*_tma -s 4 -c 2 -f -o gpurun_out/prof_r2_gae \
    python bench.py --steps 3 --warmup 1 --no-extras > /dev/null 2>&1
for spec in "gae_fwd_lookback:gae_small" "qrdqn_fwd_kernel:qrdqn" "iqn_fwd_kernel:iqn" "dist_nstep_fwd_lane:dist" \
            "scatter_rows_bulk:dist"; do
  k=${spec%%:*}; op=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_r2_$k \
      python tools/bench_ops.py --quick --ops $op > gpurun_out/ncu_$k.log 2>&1
done
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_smoke.py > gpurun_out/sanitizer_$tool.txt 2>&1
  tail -2 gpurun_out/sanitizer_$tool.txt
done
echo round checks done
