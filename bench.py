#!/usr/bin/env python
"""bench.py -- headline benchmark of the trajectory-return hot path on B200.

Metric (BASELINE.json): GAE forward+backward trajectory-steps/s, plus % of the HBM roofline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one GAE forward + one GAE adjoint over one batch of synthetic trajectories
(T=1024, B=65536 per GPU, fp32 -- BASELINE.json configs[1]; with N GPUs the global batch is
N*65536 columns sharded on B with no data-path collective => weak scaling, configs[4]).

  value     whole-job steps/s with inputs resident in HBM (CUDA events over exactly K steps,
            barrier + synchronize on both sides, max over ranks)
  e2e       same metric through the host-buffer C-ABI entry (hpc_rll_gae_fwd_bwd_host): pinned host
            arrays in, host arrays out, H2D/D2H copies inside the timed region
  roofline  dominant kernel's algorithmic bytes / its mean CUDA-event duration vs the measured HBM peak
  cpu_baseline  the oracle port (oracle/oracle.c, OpenMP) on this box's host cores, same workload

`--impl reference` times the CPU implementation (the oracle port: the reference's own CPU path is
the Python/PyTorch `hpc_rll.origin`, which is not present on the GPU box) on the same config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

T_DEFAULT, B_DEFAULT = 1024, 65536
GAMMA, LAMBDA = 0.99, 0.97
OUT = sys.stdout
METRIC = "gae_fwd_bwd_trajectory_steps_per_sec"
UNIT = "steps/s"


def algorithmic_bytes(T, B):
    """12 B/step forward (v, r in; adv out) + 12 B/step backward (G in; g_r, g_v out)
    + one extra (B,) row each for value[T] / grad_value[T]."""
    fwd = 12 * T * B + 4 * B
    bwd = 12 * T * B + 4 * B
    return fwd, bwd


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            with open(p) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled while the GPU is under load."""
    FIELDS = ("clocks.sm,clocks.max.sm,utilization.gpu,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                clk, mx, util = float(parts[0]), float(parts[1]), float(parts[2])
            except ValueError:
                continue
            smax = mx
            if util >= 20.0:
                sm.append(clk)
                for n, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples_under_load": len(sm)}


def cpu_baseline(T, B, reps):
    """Oracle port (C + OpenMP) on the same workload: fwd + adjoint.  The thread count is probed (all host
    threads, then fewer) and the fastest is reported with the threads it actually used."""
    import numpy as np
    from oracle import oracle as orc
    orc.build()
    ncpu = os.cpu_count() or 1
    try:
        ncpu = min(ncpu, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    rng = np.random.default_rng(1234)
    src = [rng.standard_normal((T + 1, B), dtype=np.float32), rng.standard_normal((T, B), dtype=np.float32),
           rng.standard_normal((T, B), dtype=np.float32)]

    def run(nthreads, n):
        orc.set_threads(nthreads)
        # parallel first-touch with this thread count so each thread streams pages local to its NUMA node
        value, reward, gadv = (orc.place(a) for a in src)
        orc.gae_forward(value, reward, GAMMA, LAMBDA)
        orc.gae_backward(gadv, GAMMA, LAMBDA)
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            orc.gae_forward(value, reward, GAMMA, LAMBDA)
            orc.gae_backward(gadv, GAMMA, LAMBDA)
            ts.append(time.perf_counter() - t0)
        ts.sort()
        return ts[len(ts) // 2]

    cands = sorted({c for c in (ncpu, ncpu // 2, ncpu // 4, 32, 16, 8) if 1 <= c <= ncpu}, reverse=True)
    probe = {c: run(c, 2) for c in cands}
    best = min(probe, key=probe.get)
    med = run(best, reps)
    return {"value": T * B / med, "unit": UNIT, "cores": best, "kind": "port",
            "sample": "full workload T=%d B=%d fp32 fwd+adjoint, median of %d runs (%.1f ms each), "
                      "oracle/oracle.c OpenMP, best of thread counts %s (host has %d)"
                      % (T, B, reps, med * 1e3, {k: round(v * 1e3, 1) for k, v in probe.items()}, ncpu)}, med


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T, B = args.T, args.B
    base, med = cpu_baseline(T, B, max(3, args.steps))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "gae_fwd_bwd T=%d B=%d fp32 (BASELINE.json configs[1])" % (T, B)},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=OUT, flush=True)


def run_ours(args):
    import torch
    from di_hpc_b200 import _abi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback exists)"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    T, B = args.T, args.B  # per-GPU shard (weak scaling)
    K, W = args.steps, max(3, args.warmup)
    L = _abi.lib()
    dev = torch.device("cuda", local)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    value = torch.randn(T + 1, B, device=dev, generator=gen)
    reward = torch.randn(T, B, device=dev, generator=gen)
    gadv = torch.randn(T, B, device=dev, generator=gen)
    adv = torch.empty(T, B, device=dev)
    gv = torch.empty(T + 1, B, device=dev)
    gr = torch.empty(T, B, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    pv, pr, pg, pa, pgv, pgr = (t.data_ptr() for t in (value, reward, gadv, adv, gv, gr))

    def fwd():
        _abi.check(L.hpc_rll_gae_forward(pv, pr, pa, T, B, GAMMA, LAMBDA, st), "gae_forward")

    def bwd():
        _abi.check(L.hpc_rll_gae_backward(pg, pgv, pgr, T, B, GAMMA, LAMBDA, st), "gae_backward")

    def step():
        fwd()
        bwd()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(W):
        step()
    # ---- timed region: exactly K steps ------------------------------------------------------
    barrier()
    n0 = _abi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    barrier()
    launches = _abi.launch_count() - n0
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    ms_per_step = total_ms / K
    value_metric = T * B * world / (ms_per_step * 1e-3)

    # ---- roofline: each kernel alone, mean CUDA-event duration on the launching stream --------
    def kernel_ms(fn, min_iters, min_ms=400.0):
        iters = min_iters
        while True:
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(iters):
                fn()
            b.record()
            torch.cuda.synchronize()
            t = a.elapsed_time(b)
            if t >= min_ms or iters >= 1 << 16:
                return t / iters
            iters *= 4

    fwd_ms = kernel_ms(fwd, max(K, 20))
    bwd_ms = kernel_ms(bwd, max(K, 20))
    fwd_b, bwd_b = algorithmic_bytes(T, B)
    peak, peak_src = hbm_peak()
    # kernel the library picks at this width (gae.cu pick_cfg): bulk-store output from 256 columns per SM up
    st = "_st" if B >= 256 * torch.cuda.get_device_properties(dev).multi_processor_count and \
        "HPC_RLL_CFG_GAE" not in os.environ else ""
    dom = ("gae_bwd_tma" if bwd_ms >= fwd_ms else "gae_fwd_tma") + st
    dom_ms, dom_b = (bwd_ms, bwd_b) if bwd_ms >= fwd_ms else (fwd_ms, fwd_b)
    achieved = dom_b / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": dom, "peak_source": peak_src,
                "fwd": {"ms": fwd_ms, "gbs": fwd_b / (fwd_ms * 1e-3) / 1e9},
                "bwd": {"ms": bwd_ms, "gbs": bwd_b / (bwd_ms * 1e-3) / 1e9},
                "step": {"ms": ms_per_step, "gbs": (fwd_b + bwd_b) / (ms_per_step * 1e-3) / 1e9,
                         "frac": (fwd_b + bwd_b) / (ms_per_step * 1e-3) / 1e9 / peak}}
    tr = os.path.join(ROOT, "profiles", "gae_dram_traffic.json")
    if os.path.exists(tr):
        try:
            with open(tr) as f:
                roofline["traffic"] = json.load(f).get(dom)
        except Exception:
            pass
    clocks = sampler.stop() if sampler else None

    # ---- e2e: host buffers through the C ABI, copies inside the timed region -----------------
    hv = torch.randn(T + 1, B, generator=torch.Generator().manual_seed(99 + rank)).pin_memory()
    hr = torch.randn(T, B).pin_memory()
    hg = torch.randn(T, B).pin_memory()
    ha = torch.empty(T, B).pin_memory()
    hgv = torch.empty(T + 1, B).pin_memory()
    hgr = torch.empty(T, B).pin_memory()

    def e2e_step():
        _abi.check(L.hpc_rll_gae_fwd_bwd_host(hv.data_ptr(), hr.data_ptr(), hg.data_ptr(), ha.data_ptr(),
                                              hgv.data_ptr(), hgr.data_ptr(), T, B, GAMMA, LAMBDA), "gae_host")

    Ke = max(3, min(K, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        e2e_step()  # returns only after the results are in the host buffers
    barrier()
    dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e_ms = float(dt.item()) * 1e3 / Ke
    h2d = 4 * ((T + 1) * B + T * B + T * B)
    d2h = 4 * (T * B + (T + 1) * B + T * B)
    e2e = {"value": T * B * world / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": Ke,
           "api": "hpc_rll_gae_fwd_bwd_host (pinned host buffers, 4-slot column-block pipeline)"}

    if rank == 0:
        base, _ = cpu_baseline(T, B, 5) if world == 1 else (None, None)
        line = {"metric": METRIC, "value": value_metric, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": "gae_fwd_bwd T=%d B=%d fp32 per GPU (BASELINE.json configs[1]; global B=%d)"
                                       % (T, B, B * world),
                           "l2": "inputs (%.0f MB per kernel) larger than L2 (126 MB); no flush needed"
                                 % (fwd_b / 1e6), "gamma": GAMMA, "lambda": LAMBDA},
                "roofline": roofline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks}
        if base is not None:
            line["cpu_baseline"] = base
        print(json.dumps(line), file=OUT, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def _claim_stdout():
    """The driver parses ONE JSON line from stdout.  Libraries (NCCL's version banner, torchrun notices)
    also write there, so stdout is pointed at stderr for the whole run and the JSON line goes to the
    saved descriptor."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(real, "w")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--T", type=int, default=T_DEFAULT)
    ap.add_argument("--B", type=int, default=B_DEFAULT, help="columns per GPU")
    args = ap.parse_args()
    global OUT
    OUT = _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    OUT.flush()


if __name__ == "__main__":
    main()
