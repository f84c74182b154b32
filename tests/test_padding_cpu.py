"""Padding oracle vs the origin-generated fixtures (CPU), and the host-side group splitters of the C ABI."""
import ctypes

import numpy as np
import pytest

from oracle import padding_oracle as po
from tests._golden import GOLDEN_DIR, names


def load(name):
    import os
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


@pytest.mark.parametrize("name", names("padding"))
def test_pad_unpad_oracle_vs_origin(name):
    z = load(name)
    n, value = int(z["n"]), int(z["value"])
    xs = [z["x%d" % i] for i in range(n)]
    new_x, mask, shapes = po.pad(xs, value)
    assert np.array_equal(new_x, z["new_x"])
    assert np.array_equal(mask.astype(np.float32), z["mask"])  # origin's mask has x's dtype, same values
    for a, b in zip(po.unpad(new_x, shapes), xs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("name", names("padding"))
def test_oracle_split_vs_origin_and_abi(name):
    from di_hpc_b200 import _abi
    z = load(name)
    shapes = [tuple(int(v) for v in s) for s in z["sorted_shapes"]]
    group, ndim = int(z["group"]), int(z["ndim"])
    want = [int(p) for p in z["positions"]]
    assert po.oracle_split_group(shapes, group) == want
    shp = np.ascontiguousarray(z["sorted_shapes"], dtype=np.int64)
    pos = np.zeros(group + 1, dtype=np.int64)
    _abi.check(_abi.lib().hpc_rll_oracle_split_group(shp.ctypes.data, len(shapes), ndim, group, pos.ctypes.data), "split")
    got = [int(p) for p in pos]
    assert got[0] == 0 and got[-1] == len(shapes) and all(a < b for a, b in zip(got, got[1:]))
    if ndim == 1:
        assert got == want  # 1-D: origin's element-count cost and the padded-volume cost coincide
    # the C-ABI splitter minimises the padded volume (reference C++ cost model): check optimality
    assert po.padded_volume(shapes, got) == po.best_volume(shapes, group)
    assert po.padded_volume(shapes, got) <= po.padded_volume(shapes, want)


def test_sample_split_properties():
    from di_hpc_b200 import _abi
    rng = np.random.default_rng(0)
    for ndim in (1, 2, 3):
        shapes = sorted([tuple(int(v) for v in rng.integers(1, 9, ndim)) for _ in range(40)], key=po.cum)
        shp = np.ascontiguousarray(shapes, dtype=np.int64)
        for group in (1, 2, 4, 7):
            for seed in range(5):
                starts = np.zeros(group + 2, dtype=np.int64)
                cnt = ctypes.c_int(0)
                _abi.check(
                    _abi.lib().hpc_rll_sample_split_group(shp.ctypes.data, len(shapes), ndim, group, seed,
                                                          starts.ctypes.data, ctypes.byref(cnt)), "sample")
                b = [int(v) for v in starts[:cnt.value + 1]]
                assert 1 <= cnt.value <= group and b[0] == 0 and b[-1] == len(shapes)
                assert all(x < y for x, y in zip(b, b[1:]))
                # equal-shaped neighbouring groups were merged
                mx = [tuple(max(s[d] for s in shapes[x:y]) for d in range(ndim)) for x, y in zip(b, b[1:])]
                assert all(p != q for p, q in zip(mx, mx[1:]))
    # same seed -> same split
    s1, s2 = np.zeros(6, dtype=np.int64), np.zeros(6, dtype=np.int64)
    c1, c2 = ctypes.c_int(0), ctypes.c_int(0)
    L = _abi.lib()
    L.hpc_rll_sample_split_group(shp.ctypes.data, len(shapes), 3, 4, 123, s1.ctypes.data, ctypes.byref(c1))
    L.hpc_rll_sample_split_group(shp.ctypes.data, len(shapes), 3, 4, 123, s2.ctypes.data, ctypes.byref(c2))
    assert c1.value == c2.value and np.array_equal(s1, s2)


def test_staged_rows_index_division_is_exact():
    """softmax_rows.cuh stage_rows/unstage_rows split a flat element index i < 256*32 into (row, col) with
    row = floor((float(i) + 0.5f) * (1.0f / N)) instead of an integer division; exact for every N <= 32."""
    import numpy as np
    for N in range(1, 33):
        inv = np.float32(1.0) / np.float32(N)
        i = np.arange(256 * 32, dtype=np.int32)
        row = np.floor((i.astype(np.float32) + np.float32(0.5)) * inv).astype(np.int32)
        assert np.array_equal(row, i // N), N
