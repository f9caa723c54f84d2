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


def test_sh_z_split_and_merge_round_trip():
    """gs.frame.sh_z_split / sh_z_merge: the four bases the ortho renderers' constant view direction (0,0,1) reaches are exactly the
    ones whose degree-3 SH basis function is non-zero there (oracle/torch_ref.py evaluates the reference's polynomial)."""
    from oracle import torch_ref as TR
    g = torch.Generator().manual_seed(2)
    shs = torch.randn(50, 16, 3, generator=g)
    z, rest = F.sh_z_split(shs)
    assert z.shape == (50, 4, 3) and rest.shape == (50, 12, 3)
    assert torch.equal(F.sh_z_merge(z, rest), shs)
    assert torch.equal(z, shs[:, list(F.SH_Z_BASES)])
    # which bases does direction (0,0,1) reach?  d colour / d coefficient of the reference's SH evaluation
    dirs = torch.zeros(50, 3); dirs[:, 2] = 1
    x = shs.clone().requires_grad_(True)
    TR.compute_sh(x, 3, dirs).sum().backward() if hasattr(TR, "compute_sh") else pytest.skip("no torch SH in the oracle")
    reached = (x.grad.abs().sum(dim=(0, 2)) > 0).nonzero().flatten().tolist()
    assert set(reached) <= set(F.SH_Z_BASES)            # clamped colours may hide a basis for SOME points, never add one
    assert len(reached) >= 1


def test_exchange_is_not_collective_on_one_process():
    """One process, nothing deferred: GradExchange.run() is a no-op, so a trainer may capture the whole step in one graph."""
    from splatter_a_video_b200.parallel import FlatParams, GradExchange
    flat = FlatParams({"a": torch.zeros(8, 3), "b": torch.zeros(8, 16, 3)})
    ex = GradExchange(flat, 8)
    assert ex.is_collective is False
    before = flat.flat_grad.clone()
    ex.run()
    assert torch.equal(flat.flat_grad, before)
    assert GradExchange(flat, 8, deferred={"shs": "b", "node": "a", "NI": 1}).is_collective is True
