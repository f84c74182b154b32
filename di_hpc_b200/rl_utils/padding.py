"""Padding -- drop-in for /root/reference/hpc_rll/rl_utils/padding.py: ``Padding{1,2,3}D`` /
``UnPadding{1,2,3}D`` over lists of ragged CUDA tensors (same signatures and return structure:
``new_x, mask, shapes``; in group mode ``[tuple(new_x), tuple(mask), tuple(shapes)]`` with per-group flat
shape lists, padding.py:39-41,93-95,155-157).  ``mask`` is int32 (1 inside, ``value`` outside), as the
reference's CUDA path returns it.

One CUDA launch pads every tensor of every group: the per-tensor descriptors travel in kernel parameter
space (di_hpc_b200/csrc/padding.cu), so a call performs no cudaMalloc / cudaMemcpy (the reference does up
to 7 of each per call, src/rl_utils/padding.cu:118-131,172-199).
"""
import ctypes
import itertools
from functools import reduce
from typing import List, Union

import numpy as np
import torch

from .. import _abi

_seed = itertools.count(0x5EED)


def cum(t) -> int:
    return reduce(lambda x, y: x * y, t)


def _check_inputs(inputs, ndim):
    assert len(inputs) > 0, "empty input list"
    dev = inputs[0].device
    out = []
    for t in inputs:
        assert t.is_cuda, "hpc version only supports cuda"
        if t.dtype != torch.float32:
            raise TypeError("padding supports float32 tensors, got %s" % t.dtype)
        if t.dim() != ndim:
            raise ValueError("expected %d-D tensors, got shape %s" % (ndim, tuple(t.shape)))
        if t.device != dev:
            raise ValueError("all tensors must live on one device")
        out.append(t.contiguous())
    return out


def _table(ptrs):
    return (ctypes.c_void_p * len(ptrs))(*ptrs)


def _shape3(shapes):
    a = np.ones((len(shapes), 3), dtype=np.int32)
    for i, s in enumerate(shapes):
        a[i, :len(s)] = s
    return a


def _pad_groups(inputs, ndim, bounds, value):
    """Pad inputs[bounds[g]:bounds[g+1]] to the max shape of group g -- one launch for all groups."""
    dev = inputs[0].device
    n = len(inputs)
    shapes = [tuple(t.shape) for t in inputs]
    new_x, masks, padded = [], [], []
    dst, msk = [0] * n, [0] * n
    for g in range(len(bounds) - 1):
        lo, hi = bounds[g], bounds[g + 1]
        mx = tuple(max(s[d] for s in shapes[lo:hi]) for d in range(ndim))
        x = torch.empty((hi - lo, ) + mx, dtype=torch.float32, device=dev)
        m = torch.empty((hi - lo, ) + mx, dtype=torch.int32, device=dev)
        vol = cum(mx) if ndim else 1
        for i in range(lo, hi):
            dst[i] = x.data_ptr() + 4 * vol * (i - lo)
            msk[i] = m.data_ptr() + 4 * vol * (i - lo)
            padded.append(mx)
        new_x.append(x)
        masks.append(m)
    sh, pd = _shape3(shapes), _shape3(padded)
    with _abi.on_device(dev):
        _abi.check(
            _abi.lib().hpc_rll_pad_batch(_table([t.data_ptr() for t in inputs]), _table(dst), _table(msk),
                                         sh.ctypes.data, pd.ctypes.data, n, int(value), _abi.stream_of(inputs[0])),
            "hpc_rll_pad_batch")
    return new_x, masks


def _split(inputs, ndim, group, group_mode):
    shp = np.ascontiguousarray([list(t.shape) for t in inputs], dtype=np.int64)
    n = len(inputs)
    L = _abi.lib()
    if group_mode == 'oracle':
        g = min(group, n)
        pos = np.zeros(g + 1, dtype=np.int64)
        _abi.check(L.hpc_rll_oracle_split_group(shp.ctypes.data, n, ndim, g, pos.ctypes.data), "oracle_split_group")
        return [int(p) for p in pos]
    starts = np.zeros(group + 2, dtype=np.int64)
    cnt = ctypes.c_int(0)
    _abi.check(
        L.hpc_rll_sample_split_group(shp.ctypes.data, n, ndim, group, next(_seed), starts.ctypes.data,
                                     ctypes.byref(cnt)), "sample_split_group")
    return [int(p) for p in starts[:cnt.value + 1]]


def _flat_shapes(inputs, lo, hi):
    out = []
    for t in inputs[lo:hi]:
        out.extend(int(d) for d in t.shape)
    return out


def _padding(inputs, ndim, mode, value, group, group_mode):
    assert mode in ['constant'], mode
    assert group_mode in ['sample', 'oracle'], group_mode
    assert group >= 1, group
    inputs = _check_inputs(inputs, ndim)
    if group > 1:
        inputs = sorted(inputs, key=lambda t: cum(t.shape))
        bounds = _split(inputs, ndim, group, group_mode)
        new_x, mask = _pad_groups(inputs, ndim, bounds, value)
        shapes = [_flat_shapes(inputs, bounds[g], bounds[g + 1]) for g in range(len(bounds) - 1)]
        return [tuple(new_x), tuple(mask), tuple(shapes)]
    new_x, mask = _pad_groups(inputs, ndim, [0, len(inputs)], value)
    return new_x[0], mask[0], _flat_shapes(inputs, 0, len(inputs))


def _unpad_one(x, shapes, ndim):
    assert x.is_cuda, "hpc version only supports cuda"
    if x.dtype != torch.float32:
        raise TypeError("padding supports float32 tensors")
    x = x.contiguous()
    n = x.shape[0]
    shp = [tuple(int(v) for v in shapes[i * ndim:(i + 1) * ndim]) for i in range(n)]
    if len(shapes) != n * ndim:
        raise ValueError("shapes must hold %d ints per tensor" % ndim)
    outs = [torch.empty(s, dtype=torch.float32, device=x.device) for s in shp]
    vol = cum(x.shape[1:])
    src = [x.data_ptr() + 4 * vol * i for i in range(n)]
    sh, pd = _shape3(shp), _shape3([tuple(x.shape[1:])] * n)
    with _abi.on_device(x.device):
        _abi.check(
            _abi.lib().hpc_rll_unpad_batch(_table(src), _table([o.data_ptr() for o in outs]), sh.ctypes.data,
                                           pd.ctypes.data, n, _abi.stream_of(x)), "hpc_rll_unpad_batch")
    return outs


def _unpadding(x, shapes, ndim):
    if isinstance(x, torch.Tensor):
        return _unpad_one(x, list(shapes), ndim)
    ret = []
    for t, s in zip(x, shapes):
        ret.append(_unpad_one(t, list(s), ndim))
    return sum(ret, [])


def Padding1D(inputs: List[torch.Tensor], mode='constant', value: int = 0, group: int = 1, group_mode='sample'):
    return _padding(inputs, 1, mode, value, group, group_mode)


def UnPadding1D(x: Union[torch.Tensor, List[torch.Tensor]], shapes: Union[List, List[List]]) -> List[torch.Tensor]:
    return _unpadding(x, shapes, 1)


def Padding2D(inputs: List[torch.Tensor], mode='constant', value: int = 0, group: int = 1, group_mode='sample'):
    return _padding(inputs, 2, mode, value, group, group_mode)


def UnPadding2D(x: Union[torch.Tensor, List[torch.Tensor]], shapes: Union[List, List[List]]) -> List[torch.Tensor]:
    return _unpadding(x, shapes, 2)


def Padding3D(inputs: List[torch.Tensor], mode='constant', value: int = 0, group: int = 1, group_mode='sample'):
    return _padding(inputs, 3, mode, value, group, group_mode)


def UnPadding3D(x: Union[torch.Tensor, List[torch.Tensor]], shapes: Union[List, List[List]]) -> List[torch.Tensor]:
    return _unpadding(x, shapes, 3)
