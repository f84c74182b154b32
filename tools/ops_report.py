"""Turn tools/bench_ops.py output (.jsonl) into the committed markdown table under profiles/."""
import json
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    rows = [json.loads(l) for l in open(src) if l.startswith("{")]
    with open(dst, "w") as f:
        f.write("# Per-op throughput through the drop-in modules (autograd forward + backward), B200, fp32\n\n"
                "source: `python tools/bench_ops.py` (CUDA events around module forward and `torch.autograd.grad`, median of 20; "
                "includes PyTorch allocator / autograd overhead).  `alg GB/s` = SURVEY.md 8(d) algorithmic bytes / (fwd+bwd time); "
                "fraction is of the measured HBM peak (MEASURED_PEAKS.json).  `sustained` = the same fwd+bwd pairs queued back to back "
                "(host launch work overlapped with the previous call's kernels, as in a training loop) and its fraction of peak.  "
                "QR-DQN/IQN at tau=64 are FP32-issue-bound, not HBM-bound.\n\n"
                "| op | shape | fwd ms | bwd ms | throughput | alg GB/s | of measured HBM peak | sustained ms (of peak) | kernel launches |\n|---|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            shp = ", ".join("%s=%s" % kv for kv in r["shape"].items())
            sus = "%.3f (%.0f %%)" % (r["pipelined_ms"], 100 * r["pipelined_hbm_frac"]) if "pipelined_ms" in r else "—"
            f.write("| %s | %s | %.3f | %.3f | %.3e %s/s | %.0f | %.0f %% | %s | %d |\n" %
                    (r["op"], shp, r["fwd_ms"], r["bwd_ms"], r["throughput"], r["unit"], r["alg_gbs"],
                     100 * r["hbm_frac_of_measured_peak"], sus, r["launches"]))


if __name__ == "__main__":
    main()
