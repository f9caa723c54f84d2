"""CPU tests of host-side logic that needs no GPU: the intersection-capacity policy of the sync-free frame path."""
import pytest
import torch

import helpers as Hh  # noqa: F401  (path setup)
from splatter_a_video_b200.gs import frame as F


class _FakeEvent:
    """Stands in for torch.cuda.Event: `done` flips when the test says the GPU got there."""
    def __init__(self):
        self.done = False

    def record(self):
        pass

    def query(self):
        return self.done

    def synchronize(self):
        self.done = True


@pytest.fixture()
def fake_events(monkeypatch):
    monkeypatch.setattr(F.torch.cuda, "Event", _FakeEvent)


def _st(i, over):
    return torch.tensor([i, over], dtype=torch.int32)


def test_capacity_current_frame_overflow_is_retried(fake_events):
    c = F.Capacity(initial=1000)
    c.observe(_st(1000, 1))
    assert c.check(wait=True) is False and c.I_cap > 1000 and c.late_overflows == 0
    c.observe(_st(1200, 0))
    assert c.check(wait=True) is True and c.last_I == 1200


def test_capacity_late_overflow_is_sticky_and_raises_once(fake_events):
    c = F.Capacity(initial=1000)
    c.observe(_st(500, 0)); assert c.check(wait=True)
    c.observe(_st(1000, 1))                      # overflow, nobody waits
    assert c.check(wait=False) is True           # not known yet
    c.observe(_st(400, 0))                       # queuing the next frame settles the previous one
    with pytest.raises(F.CapacityOverflow):
        c.check(wait=False)
    assert c.late_overflows == 1 and c.I_cap > 1000
    assert c.check(wait=True) is True            # reported once


def test_capacity_grows_before_it_overflows_and_follows_population(fake_events):
    c = F.Capacity(initial=1000)
    c.set_population(100)
    c.observe(_st(900, 0)); assert c.check(wait=True)      # 90 % full: proactive head-room
    assert c.I_cap > 1000
    cap = c.I_cap
    c.set_population(300)
    assert c.I_cap >= 3 * cap
    c.set_population(150)                                   # pruning never shrinks the buffers
    assert c.I_cap >= 3 * cap
