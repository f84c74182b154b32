"""End-to-end (host-buffer) path probe: PCIe ceilings, then hpc_rll_gae_fwd_bwd_host at C1 for several chunk heights
and both kinds of page-locked memory (torch pin_memory vs the library's NUMA-local allocator).

    python tools/probe_e2e.py [--rows 8,16,32,64] [--T 1024] [--B 65536]
Each chunk height runs in a child process (the library reads HPC_RLL_HOST_CHUNK_ROWS once)."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(T, B, numa, iters):
    import torch
    from di_hpc_b200 import host as hp
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    node = hp.bind_to_device() if numa else -2
    mk = (lambda s: hp.pinned_empty(s)) if numa else (lambda s: torch.empty(s).pin_memory())
    v, r, g = mk((T + 1, B)), mk((T, B)), mk((T, B))
    out = (mk((T, B)), mk((T + 1, B)), mk((T, B)))
    for t in (v, r, g):
        t.normal_()
    for _ in range(2):
        hp.gae_fwd_bwd_host(v, r, g, out=out)
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter()
        hp.gae_fwd_bwd_host(v, r, g, out=out)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print(json.dumps({"rows": os.environ.get("HPC_RLL_HOST_CHUNK_ROWS", "auto"), "numa_alloc": numa, "node": node,
                      "ms_min": ts[0] * 1e3, "ms_med": ts[len(ts) // 2] * 1e3,
                      "gbs_each_way_med": 12 * T * B / ts[len(ts) // 2] / 1e9}))


def pcie(T, B):
    import torch
    from di_hpc_b200 import host as hp
    res = {}
    for numa in (False, True):
        if numa:
            hp.bind_to_device()
        mk = (lambda s: hp.pinned_empty(s)) if numa else (lambda s: torch.empty(s).pin_memory())
        h, h2 = mk((T, B)), mk((T, B))
        d, d2 = torch.empty(T, B, device="cuda"), torch.empty(T, B, device="cuda")
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        nb = T * B * 4

        def rate(fn, it=8):
            fn()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(it):
                fn()
            torch.cuda.synchronize()
            return nb * it / (time.perf_counter() - t0) / 1e9

        def both():
            with torch.cuda.stream(s1):
                d.copy_(h, non_blocking=True)
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)

        k = "numa" if numa else "torch_pin"
        res[k] = {"h2d": rate(lambda: d.copy_(h, non_blocking=True)), "d2h": rate(lambda: h2.copy_(d2, non_blocking=True)),
                  "duplex_each": rate(both)}
    print(json.dumps({"pcie_gbs": res}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default="auto,8,16,32,64,128")
    ap.add_argument("--T", type=int, default=1024)
    ap.add_argument("--B", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--child", default="")
    a = ap.parse_args()
    if a.child == "pcie":
        return pcie(a.T, a.B)
    if a.child:
        return child(a.T, a.B, a.child == "numa", a.iters)
    base = [sys.executable, os.path.abspath(__file__), "--T", str(a.T), "--B", str(a.B), "--iters", str(a.iters)]
    subprocess.run(base + ["--child", "pcie"])
    for rows in a.rows.split(","):
        for kind in ("torch", "numa"):
            env = dict(os.environ)
            if rows != "auto":
                env["HPC_RLL_HOST_CHUNK_ROWS"] = rows
            subprocess.run(base + ["--child", kind], env=env)


if __name__ == "__main__":
    main()
