import numpy as np, torch, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from di_hpc_b200 import _abi
from tests.test_nstep_gpu import run_qrdqn, base_inputs
from tests._gpu import rng
from oracle import oracle as orc
for (tau, T, B, N, use_w, use_vg) in [(64, 5, 2000, 8, True, False), (64, 5, 64, 8, True, False), (64, 5, 9, 8, False, False), (33, 5, 2000, 8, True, False)]:
    g = rng(tau * 3 + T + B + N)
    inp = base_inputs(g, T, B, N, use_w)
    inp["q"] = g.standard_normal((B, N, tau), dtype=np.float32)
    inp["next_n_q"] = g.standard_normal((B, N, tau), dtype=np.float32)
    inp["value_gamma"] = None
    o = orc.qrdqn_nstep_td(inp["q"], inp["next_n_q"], inp["action"], inp["next_n_action"], inp["reward"], inp["done"], inp["weight"], None, 0.95, 1.1)
    for cfg in (-1, 1):
        _abi.set_config(7, cfg)
        loss, td, gq = run_qrdqn(inp, 0.95, 1.1)
        et = np.abs(td - o["td_error_per_sample"]); eg = np.abs(gq - o["grad_q"]).reshape(B, -1).max(1)
        print(tau, B, "cfg", cfg, "loss", loss, o["loss"], "td max err", et.max(), "at", et.argmax(), "of", np.abs(td).max(), "grad err", eg.max() / np.abs(o["grad_q"]).max(), "at sample", eg.argmax(), "bad samples", int((et > 1e-3).sum()))
    _abi.set_config(7, -1)
