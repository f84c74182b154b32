"""The native boundary of the reference: a module importable as `hpc_rl_utils` exporting the 30 binding names of
/root/reference/src/rl_utils/entry.cpp:9-38 with tensor-list signatures.  CPU part: names, arities, and that the
UNMODIFIED reference wrappers import and construct against the shim (no kernel runs without a GPU)."""
import re

import pytest

from tests import _refwrap

HOT = {"GaeForward": 4, "TdLambdaForward": 4, "TdLambdaBackward": 2, "DistNStepTdForward": 5, "DistNStepTdBackward": 2,
       "QNStepTdForward": 3, "QNStepTdBackward": 2, "QNStepTdRescaleForward": 3, "QNStepTdRescaleBackward": 2,
       "UpgoForward": 2, "UpgoBackward": 2, "VTraceForward": 7, "VTraceBackward": 2, "PPOForward": 5, "PPOBackward": 2,
       "IQNNStepTDErrorForward": 4, "IQNNStepTDErrorBackward": 2, "QRDQNNStepTDErrorForward": 3,
       "QRDQNNStepTDErrorBackward": 2}
PAD = ["sample_split_group", "oracle_split_group", "Pad1DForward", "GroupPad1DForward", "Unpad1DForward", "Pad2DForward",
       "GroupPad2DForward", "Unpad2DForward", "Pad3DForward", "GroupPad3DForward", "Unpad3DForward"]


def test_all_reference_binding_names_exist_with_reference_arity():
    import hpc_rl_utils
    for name in list(HOT) + PAD:
        assert callable(getattr(hpc_rl_utils, name)), name
    for name, nargs in HOT.items():
        sig = getattr(hpc_rl_utils, name).__doc__.splitlines()[0]
        assert sig.count("arg") == nargs, (name, sig)
        assert "Sequence[torch.Tensor]" in sig  # (inputs, outputs, floats...) like entry.h:62-165


def test_entry_cpp_lists_exactly_these_names():
    import os
    p = "/root/reference/src/rl_utils/entry.cpp"
    if not os.path.exists(p):
        pytest.skip("reference tree not present")
    names = re.findall(r'm\.def\("(\w+)"', open(p).read())
    assert sorted(names) == sorted(list(HOT) + PAD)


@pytest.mark.parametrize("mod", ["gae", "td", "upgo", "vtrace", "ppo"])
def test_unmodified_reference_wrappers_bind_to_the_shim(mod):
    if _refwrap.wrapper_dir() is None:
        pytest.skip("reference wrappers not available (run tools/stage_ref_wrappers.sh in the build container)")
    import hpc_rl_utils
    m = _refwrap.load(mod)
    assert m.hpc_rl_utils is hpc_rl_utils
    used = set(re.findall(r"hpc_rl_utils\.(\w+)", open(m.__file__).read()))
    assert used and all(hasattr(hpc_rl_utils, u) for u in used), used
    # the modules build their scratch buffers on the CPU just like in the reference
    ctor = {"gae": lambda: m.GAE(4, 3), "td": lambda: (m.TDLambda(4, 3), m.QNStepTD(4, 3, 2), m.DistNStepTD(4, 3, 2, 5),
                                                      m.QRDQNNStepTDError(8, 4, 3, 2), m.IQNNStepTDError(8, 9, 4, 3, 2)),
            "upgo": lambda: m.UPGO(4, 3, 2), "vtrace": lambda: m.VTrace(4, 3, 2), "ppo": lambda: m.PPO(3, 2)}
    ctor[mod]()
