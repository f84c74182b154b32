"""numpy restatement of the reference's padding utilities.  TEST INFRASTRUCTURE ONLY (see oracle.c header).

Follows /root/reference/hpc_rll/origin/padding.py: `_Padding1D` (:47-56), `Padding2D` (:107-121),
`Padding3D` (:134-148), `_UnPadding1D`/`UnPadding2D`/`UnPadding3D` (:88-96, 124-131, 151-158) and the
`oracle_split_group` dynamic programme (:12-45).  Pinned to origin outputs by tests/golden/padding_*.npz.
"""
import numpy as np


def cum(shape) -> int:
    r = 1
    for s in shape:
        r *= int(s)
    return r


def pad(xs, value=0):
    """new_x, mask (int32: 1 inside, `value` outside), shapes -- padding.py:47-56 generalised to N-D."""
    shapes = [tuple(x.shape) for x in xs]
    max_shape = tuple(max(s[d] for s in shapes) for d in range(len(shapes[0])))
    new_x = np.full((len(xs), ) + max_shape, value, dtype=xs[0].dtype)
    mask = np.full((len(xs), ) + max_shape, value, dtype=np.int32)
    for i, x in enumerate(xs):
        sl = (i, ) + tuple(slice(0, s) for s in x.shape)
        new_x[sl] = x
        mask[sl] = 1
    return new_x, mask, shapes


def unpad(x, shapes):
    """padding.py:88-96: the leading `shape` block of every row of the padded batch."""
    return [x[(i, ) + tuple(slice(0, s) for s in shp)].copy() for i, shp in enumerate(shapes)]


def oracle_split_group(shapes, group):
    """padding.py:12-45: DP over the size-sorted list; cost of a group = numel(last item) * group size.
    Returns the positions list (length group + 1, positions[0] = 0, positions[-1] = n)."""
    arr = [None] + [cum(s) for s in shapes]
    N, M = len(arr) - 1, group

    def p(start, end):
        return arr[end] * (end - start + 1)

    f = {(0, 0): (0, 0)}
    for i in range(1, N + 1):
        for j in range(1, M + 1):
            ress = []
            for k in range(0, i):
                if (k, j - 1) in f:
                    ress.append((f[(k, j - 1)][0] + p(k + 1, i), k))
            if ress:
                f[(i, j)] = min(ress)
    last_position, last_cnt = N, M
    positions = [N]
    while last_position > 0:
        _, last_position = f[(last_position, last_cnt)]
        last_cnt -= 1
        positions.append(last_position)
    return positions[::-1]


def padded_volume(shapes, positions):
    """cost model of the reference's C++ splitter (src/rl_utils/padding.cu:44-108): per group,
    count * prod_d max_d."""
    tot = 0
    for a, b in zip(positions[:-1], positions[1:]):
        if b > a:
            mx = [max(s[d] for s in shapes[a:b]) for d in range(len(shapes[0]))]
            tot += (b - a) * cum(mx)
    return tot


def best_volume(shapes, group):
    """brute-force optimum of `padded_volume` over all splits into exactly `group` groups (small n only)."""
    n = len(shapes)
    INF = float("inf")
    cost = [[INF] * (group + 1) for _ in range(n + 1)]
    cost[0][0] = 0
    for i in range(1, n + 1):
        for j in range(1, group + 1):
            for k in range(j - 1, i):
                if cost[k][j - 1] < INF:
                    cost[i][j] = min(cost[i][j], cost[k][j - 1] + padded_volume(shapes, [k, i]))
    return cost[n][group]
