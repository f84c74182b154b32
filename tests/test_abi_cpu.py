"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads, and exports every symbol
declared in include/hpc_rll_b200.h; the ctypes table covers the header one to one; the Python
modules expose the reference's class names and signatures; the product never imports the oracle."""
import ctypes
import inspect
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "hpc_rll_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hpc_rll_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from di_hpc_b200 import _abi, build
    lib_path = build.build()
    handle = ctypes.CDLL(lib_path)
    names = header_functions()
    assert len(names) >= 26
    for n in names:
        assert hasattr(handle, n), "library does not export %s" % n
    assert set(names) == set(_abi.SIGNATURES), set(names) ^ set(_abi.SIGNATURES)
    L = _abi.lib()
    assert b"sm_100a" in L.hpc_rll_version()
    assert L.hpc_rll_launch_count() == 0  # nothing can have launched without a GPU


def test_argument_errors_are_reported_without_a_gpu():
    from di_hpc_b200 import _abi
    L = _abi.lib()
    assert L.hpc_rll_gae_forward(None, None, None, 4, 4, 0.99, 0.97, None) == 1
    assert b"null" in L.hpc_rll_last_error()
    assert L.hpc_rll_gae_forward(None, None, None, -1, 4, 0.99, 0.97, None) == 1
    assert L.hpc_rll_debug_set_config(99, 0) == 1
    assert L.hpc_rll_workspace_bytes(_abi.OP_VTRACE, 16, 8, 4) >= 2 * 16 * 8 * 4


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from di_hpc_b200 import _abi
    monkeypatch.setattr(_abi, "_lib", None)
    monkeypatch.setattr(_abi, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _abi.lib()
    except _abi.HpcRllError as e:
        assert "no fallback" in str(e).lower()
    else:
        raise AssertionError("a missing CUDA library must raise")


def test_reference_api_surface():
    """Module paths, class names, constructor argument orders and forward signatures of
    /root/reference/hpc_rll/rl_utils."""
    import hpc_rll.rl_utils.gae as gae
    import hpc_rll.rl_utils.ppo as ppo
    import hpc_rll.rl_utils.td as td
    import hpc_rll.rl_utils.upgo as upgo
    import hpc_rll.rl_utils.vtrace as vtrace

    def params(fn):
        return [p for p in inspect.signature(fn).parameters if p != "self"]

    assert params(gae.GAE.__init__) == ["T", "B"]
    assert params(gae.GAE.forward) == ["value", "reward", "gamma", "lambda_"]
    assert params(td.TDLambda.__init__) == ["T", "B"]
    assert params(td.TDLambda.forward) == ["value", "reward", "weight", "gamma", "lambda_"]
    assert params(td.QNStepTD.__init__) == ["T", "B", "N"]
    assert params(td.QNStepTD.forward) == ["q", "next_n_q", "action", "next_n_action", "reward", "done", "weight",
                                           "gamma"]
    assert params(td.QNStepTDRescale.forward) == params(td.QNStepTD.forward)
    assert params(td.DistNStepTD.__init__) == ["T", "B", "N", "n_atom"]
    assert params(td.DistNStepTD.forward) == ["dist", "next_n_dist", "action", "next_n_action", "reward", "done",
                                              "weight", "gamma", "v_min", "v_max"]
    assert params(td.QRDQNNStepTDError.__init__) == ["tau", "T", "B", "N"]
    assert params(td.QRDQNNStepTDError.forward) == ["q", "next_n_q", "action", "next_n_action", "reward", "done",
                                                    "gamma", "weight", "value_gamma"]
    assert params(td.IQNNStepTDError.__init__) == ["tau", "tauPrime", "T", "B", "N"]
    assert params(td.IQNNStepTDError.forward) == ["q", "next_n_q", "action", "next_n_action", "reward", "done",
                                                  "replay_quantiles", "gamma", "kappa", "weight", "value_gamma"]
    assert params(upgo.UPGO.__init__) == ["T", "B", "N"]
    assert params(upgo.UPGO.forward) == ["target_output", "rhos", "action", "rewards", "bootstrap_values"]
    assert params(vtrace.VTrace.__init__) == ["T", "B", "N"]
    assert params(vtrace.VTrace.forward) == ["target_output", "behaviour_output", "action", "value", "reward",
                                             "weight", "gamma", "lambda_", "rho_clip_ratio", "c_clip_ratio",
                                             "rho_pg_clip_ratio"]
    assert params(ppo.PPO.__init__) == ["B", "N"]
    assert params(ppo.PPO.forward)[:11] == ["logits_new", "logits_old", "action", "value_new", "value_old", "adv",
                                            "return_", "weight", "clip_ratio", "use_value_clip", "dual_clip"]
    # the one extension (fused advantage normalisation) is keyword-only and off by default
    extra = [p for p in inspect.signature(ppo.PPO.forward).parameters.values() if p.name not in
             params(ppo.PPO.forward)[:11] and p.name != "self"]
    assert [(p.name, p.kind, p.default) for p in extra] == [("adv_stats", inspect.Parameter.KEYWORD_ONLY, None)]
    assert vtrace.hpc_vtrace_loss._fields == ("policy_loss", "value_loss", "entropy_loss")
    assert ppo.hpc_ppo_loss._fields == ("policy_loss", "value_loss", "entropy_loss")
    assert ppo.hpc_ppo_info._fields == ("approx_kl", "clipfrac")
    defaults = inspect.signature(gae.GAE.forward).parameters
    assert defaults["gamma"].default == 0.99 and defaults["lambda_"].default == 0.97
    d2 = inspect.signature(td.TDLambda.forward).parameters
    assert d2["gamma"].default == 0.9 and d2["lambda_"].default == 0.8
    for name in ("GAEFunction", ):
        assert hasattr(gae, name)
    for name in ("TDLambdaFunction", "QNStepTDFunction", "QNStepTDRescaleFunction", "DistNStepTDFunction",
                 "QRDQNNStepTDErrorFunction", "IQNNStepTDErrorFunction"):
        assert hasattr(td, name)
    assert hasattr(upgo, "UpgoFunction") and hasattr(vtrace, "VtraceFunction") and hasattr(ppo, "PPOFunction")


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: no file of the product packages may import or reference it."""
    for pkg in ("di_hpc_b200", "hpc_rll"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, \
                        "%s references the oracle" % os.path.join(dirpath, f)


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/hpc_rll_b200.h must compile as C99 (no C++-isms, no torch types)."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "hpc_rll_b200.h")
    r = subprocess.run(["/usr/bin/gcc", "-fsyntax-only", "-x", "c", "-std=c99", "-Wall", "-Wextra", "-pedantic", hdr],
                       capture_output=True, text=True)
    assert r.returncode == 0 and not r.stderr.strip(), r.stderr
    txt = open(hdr).read()
    assert "torch" not in re.sub(r"/\*.*?\*/", "", txt, flags=re.S), "no torch types in the C ABI"
    assert 'extern "C"' in txt


def test_di_engine_hpc_wrapper_lookup():
    """DI-engine's ding/hpc_rl/wrapper.py resolves `hpc_rll.rl_utils.<module>.<Class>`, builds `cls(*shape).cuda()`
    once per input shape and calls forward: every name it looks up must resolve and construct."""
    import importlib
    table = {"gae": ("gae", "GAE", (16, 8)), "td_lambda_error": ("td", "TDLambda", (16, 8)),
             "dist_nstep_td_error": ("td", "DistNStepTD", (5, 8, 4, 51)), "q_nstep_td_error": ("td", "QNStepTD", (5, 8, 4)),
             "q_nstep_td_error_with_rescale": ("td", "QNStepTDRescale", (5, 8, 4)),
             "qrdqn_nstep_td_error": ("td", "QRDQNNStepTDError", (32, 5, 8, 4)),
             "iqn_nstep_td_error": ("td", "IQNNStepTDError", (32, 32, 5, 8, 4)),
             "upgo_loss": ("upgo", "UPGO", (16, 8, 4)), "vtrace_error": ("vtrace", "VTrace", (16, 8, 4)),
             "ppo_error": ("ppo", "PPO", (8, 4))}
    import torch
    for fn_name, (mod, cls_name, shape) in table.items():
        cls = getattr(importlib.import_module("hpc_rll.rl_utils." + mod), cls_name)
        m = cls(*shape).cuda()  # parameter-free modules: .cuda() needs no device
        assert isinstance(m, torch.nn.Module) and callable(m.forward), fn_name
