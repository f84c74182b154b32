from di_hpc_b200.rl_utils.gae import GAE, GAEFunction  # noqa: F401
