from di_hpc_b200.rl_utils.padding import (Padding1D, Padding2D, Padding3D, UnPadding1D, UnPadding2D,  # noqa: F401
                                          UnPadding3D)
