#!/bin/bash
# One full ncu capture per kernel family (4th launch of each), merged back via gpurun_out/.
# Read here with: python tools/ncu_summary.py --many profiles/rNN_ncu_ops.md "title" gpurun_out/prof_*.ncu-rep
set -u
for spec in "vtrace_rows_fwd:vtrace" "softmax_grad_rows_kernel:vtrace" "vtrace_scan_tma:vtrace" "upgo_rows_fwd:upgo" \
            "ppo_rows_fwd:ppo" "qrdqn_fwd_kernel:qrdqn" "iqn_fwd_kernel:iqn" "scatter_rows_kernel:qrdqn" \
            "td_lambda_fwd_tma:td_lambda" "dist_nstep_fwd_kernel:dist" \
            "q_nstep_fwd_kernel:q" "gae_fwd_tma:chain"; do
  k=${spec%%:*}; op=${spec##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k python tools/bench_ops.py --quick --ops $op > gpurun_out/ncu_$k.log 2>&1
done
echo ncu_ops done
