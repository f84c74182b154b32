"""Measure the host<->device copy ceilings that bound the e2e (host-buffer) path: 1-D pinned copies each
direction, both directions at once, and the pitched 2-D column-block copies the host pipeline issues."""
import json
import time

import torch


def rate(fn, nbytes, it=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(it):
        fn()
    torch.cuda.synchronize()
    return nbytes * it / (time.perf_counter() - t0) / 1e9


def main():
    T, B = 1024, 65536
    h = torch.empty(T, B).pin_memory()
    h2 = torch.empty(T, B).pin_memory()
    d = torch.empty(T, B, device="cuda")
    d2 = torch.empty(T, B, device="cuda")
    nb = h.numel() * 4
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {}
    out["h2d_1d_gbs"] = rate(lambda: d.copy_(h, non_blocking=True), nb)
    out["d2h_1d_gbs"] = rate(lambda: h2.copy_(d2, non_blocking=True), nb)

    def both():
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        with torch.cuda.stream(s2):
            h2.copy_(d2, non_blocking=True)

    out["bidir_each_gbs"] = rate(both, nb)
    # pitched column blocks (what hpc_rll_gae_fwd_bwd_host issues): 3968 columns x 1024 rows
    cb = 3968
    dblk = torch.empty(T, cb, device="cuda")

    def cols():
        for c0 in range(0, B - cb + 1, cb):
            dblk.copy_(h[:, c0:c0 + cb], non_blocking=True)

    out["h2d_2d_colblock_gbs"] = rate(cols, (B // cb) * cb * T * 4, it=3)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
