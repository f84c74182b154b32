"""Host-side mirror of the reference package `hpc_rll.rl_utils` (modules gae, td, upgo, vtrace, ppo)."""
