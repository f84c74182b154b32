"""Host logic of the end-to-end path without a GPU: the stage heights of the T-chunked pipeline
(hpc_rll_debug_host_schedule -> csrc/gae.cu host_schedule)."""
import ctypes

import pytest

from di_hpc_b200 import _abi


def schedule(T, B):
    buf = (ctypes.c_int64 * 4096)()
    n = _abi.lib().hpc_rll_debug_host_schedule(T, B, buf, 4096)
    assert 0 < n <= 4096
    return list(buf[:n])


@pytest.mark.parametrize("T,B", [(1024, 65536), (1023, 65536), (257, 9000), (5, 64), (4, 4), (100, 1301), (1, 7),
                                  (4096, 1 << 20), (64, 524288), (8191, 4096)])
def test_stage_heights_cover_T_in_aligned_rows(T, B):
    rows = schedule(T, B)
    assert sum(rows) == T and all(r > 0 for r in rows)
    # every stage starts on a multiple of 4 rows (whole TMA boxes, 16-byte aligned d_t slices) in BOTH walk directions:
    # the backward walks the list from row 0 up, the forward walks it mirrored from row T down
    off = 0
    for r in rows[:-1]:
        assert r % 4 == 0
        off += r
        assert off % 4 == 0
    # T % 4 extra rows ride in the last stage (the one that touches row T)
    assert rows[-1] % 4 == T % 4 or len(rows) == 1
    # uniform by default: all interior stages equal, about 16 MB per staged tensor
    if len(rows) > 2:
        inner = rows[1:-1]
        assert max(inner) - min(inner) <= max(inner) // 2 + 4
        assert rows[1] * B * 4 <= 20 << 20


def test_default_c1_schedule_is_16_stages_of_64_rows():
    assert schedule(1024, 65536) == [64] * 16
