"""LazyScalar (di_hpc_b200/rl_utils/ppo.py): host-side behaviour without a GPU -- the wait happens once, on first read."""
import torch

from di_hpc_b200.rl_utils.ppo import LazyScalar


class FakeEvent:
    def __init__(self):
        self.syncs = 0

    def synchronize(self):
        self.syncs += 1

    def query(self):
        return self.syncs > 0


def test_lazy_scalar_reads_once_and_behaves_like_a_number():
    host = torch.tensor([0.25, 3.0])
    ev = FakeEvent()
    a, b = LazyScalar(host, 0, ev), LazyScalar(host, 1, ev)
    assert ev.syncs == 0 and not a.ready()
    assert float(a) == 0.25 and ev.syncs == 1
    assert a + 1 == 1.25 and 1 + a == 1.25 and a * 2 == 0.5 and a - 0.25 == 0.0 and 1 - a == 0.75 and a / 0.5 == 0.5
    assert ev.syncs == 1  # cached after the first read
    assert b > a and b >= 3 and a < 1 and a <= 0.25 and b == 3.0 and -a == -0.25 and abs(-1 * b) == 3.0
    assert "%.2f" % a == "0.25" and "{:.1f}".format(b) == "3.0" and repr(b) == "3.0" and bool(a) and hash(a) == hash(0.25)
    assert ev.syncs == 2 and b.ready()
