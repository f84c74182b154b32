"""`hpc_rll` -- the reference's package name, re-exporting the B200-native implementation.

DI-engine imports `hpc_rll.rl_utils.{gae,td,upgo,vtrace,ppo}`; those module
paths resolve to `di_hpc_b200.rl_utils.*`.  `hpc_rll.origin` (the reference's PyTorch oracle) and
`hpc_rll.torch_utils` are intentionally absent: out of scope for this path.
"""
