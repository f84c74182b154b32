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
                "fraction is of the measured HBM peak (MEASURED_PEAKS.json).  QR-DQN/IQN at tau=64 are FP32-issue-bound, not HBM-bound.\n\n"
                "| op | shape | fwd ms | bwd ms | throughput | alg GB/s | of measured HBM peak | kernel launches |\n|---|---|---|---|---|---|---|---|\n")
        for r in rows:
            shp = ", ".join("%s=%s" % kv for kv in r["shape"].items())
            f.write("| %s | %s | %.3f | %.3f | %.3e %s/s | %.0f | %.0f %% | %d |\n" %
                    (r["op"], shp, r["fwd_ms"], r["bwd_ms"], r["throughput"], r["unit"], r["alg_gbs"],
                     100 * r["hbm_frac_of_measured_peak"], r["launches"]))


if __name__ == "__main__":
    main()
