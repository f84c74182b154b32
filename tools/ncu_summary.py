"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into a small committed text file.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_xxx.md "title"
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active",
    "smsp__average_warps_issue_stalled_wait_per_issue_active",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active",
]


def main():
    """one report:   ncu_summary.py a.ncu-rep out.md "title"
    several reports: ncu_summary.py --many out.md "title" a.ncu-rep b.ncu-rep ...   (one section per kernel)"""
    if sys.argv[1] == "--many":
        out, title, reps = sys.argv[2], sys.argv[3], sys.argv[4:]
    else:
        reps, out, title = [sys.argv[1]], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    with open(out, "w") as f:
        f.write("# %s\n\nsource: %s (ncu --set full --clock-control none), read with `ncu -i ... --page raw --csv`\n\n"
                % (title, ", ".join("`%s`" % r for r in reps)))
        for rep in reps:
            summarise(rep, f)


def summarise(rep, f):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    if True:
        for r in rows[2:]:
            f.write("## %s\n\n| metric | value | unit |\n|---|---|---|\n" % r[idx["Kernel Name"]])
            for k in KEYS:
                if k in idx:
                    f.write("| %s | %s | %s |\n" % (k, r[idx[k]], units[idx[k]]))
            try:  # the two columns can carry different units (e.g. Gbyte vs Mbyte)
                scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}
                rd = float(r[idx["dram__bytes_read.sum"]]) * scale[units[idx["dram__bytes_read.sum"]]]
                wr = float(r[idx["dram__bytes_write.sum"]]) * scale[units[idx["dram__bytes_write.sum"]]]
                f.write("| dram traffic (read+write) | %.3f | Mbyte |\n" % (rd + wr))
            except Exception:
                pass
            f.write("\n")


if __name__ == "__main__":
    main()
