"""-m gpu: densification on the flat SoA (`spv_densify_*`, `spv_flat_regather`, `spv_split_children`, `spv_reset_opacity`;
splatter_a_video_b200/densify.py) against the tensor-by-tensor restatement of the reference optimizer (oracle/densify_ref.py):
same population order, same parameters, same Adam moments, same statistics."""
import pytest
import torch

import helpers as Hh  # noqa: F401
from oracle import densify_ref as R
from splatter_a_video_b200.densify import FlatDensifier
from splatter_a_video_b200.parallel import FlatAdam, FlatParams

pytestmark = pytest.mark.gpu
CFG = dict(split_num=2, grad_threshold=0.0002, percent_dense=0.01, extent=1.0, min_opacity=0.005, size_threshold=20.0)


def _population(P, seed):
    g = torch.Generator().manual_seed(seed)
    attrs = {"node": 0.01 * torch.randn(P, 24, generator=g),
             "scaling": torch.log(0.004 * torch.exp(1.2 * torch.randn(P, 3, generator=g))),     # both sides of percent_dense * extent
             "rotation": torch.randn(P, 4, generator=g),
             "opacity": 3.0 * torch.randn(P, 1, generator=g) - 2.0,                                # some below min_opacity
             "shs": torch.randn(P, 16, 3, generator=g)}
    extras = {"position": torch.rand(P, 3, generator=g) * 2 - 1}
    moments = {k: (torch.randn(v.shape, generator=g), torch.rand(v.shape, generator=g)) for k, v in attrs.items()}
    return attrs, extras, moments, g


def _build(cuda, attrs, extras, moments, P):
    flat = FlatParams({k: v.to(cuda) for k, v in attrs.items()})
    adam = FlatAdam(flat, {k: 1e-3 for k in attrs})
    off = 0
    for k, n in zip(flat.names, flat.sizes):
        adam.exp_avg[off:off + n] = moments[k][0].reshape(-1).to(cuda)
        adam.exp_avg_sq[off:off + n] = moments[k][1].reshape(-1).to(cuda)
        off += n
    adam.t = 7
    den = FlatDensifier(flat, P, {"position": "position", "scaling": "scaling", "rotation": "rotation", "opacity": "opacity"},
                        extras={k: v.to(cuda) for k, v in extras.items()}, percent_dense=CFG["percent_dense"], split_num=CFG["split_num"],
                        densify_grad_threshold=CFG["grad_threshold"], min_opacity=CFG["min_opacity"], cameras_extent=CFG["extent"],
                        size_threshold=CFG["size_threshold"])
    return flat, adam, den


@pytest.mark.parametrize("duplicate,prune", [(True, True), (True, False), (False, True)])
def test_densify_matches_reference_restatement(cuda, duplicate, prune):
    P = 6000
    attrs, extras, moments, g = _population(P, seed=3)
    flat, adam, den = _build(cuda, attrs, extras, moments, P)
    state = {"grad_accum": torch.zeros(P), "denom": torch.zeros(P), "max_radii": torch.zeros(P)}
    for it in range(3):                                               # statistics over three "steps"
        vg = 0.0006 * torch.randn(P, 2, generator=g) * (torch.rand(P, 1, generator=g) < 0.5)
        radii = (torch.rand(P, generator=g) * 30).int() * (torch.rand(P, generator=g) < 0.7).int()
        vis = radii > 0
        R.update_stats(state, vg, radii, vis)
        den.update_stats(vg.to(cuda), radii.to(cuda), vis.to(cuda) if it else None)
    for k, t in (("grad_accum", den.grad_accum), ("denom", den.denom), ("max_radii", den.max_radii)):
        assert torch.allclose(t.cpu(), state[k], rtol=1e-6, atol=1e-9), k
    new_flat, new_adam, new_extras = den.densify_and_prune(adam, duplicate, prune, generator=None)
    samples = None if den.last_samples is None else den.last_samples.cpu()
    ref_attrs, ref_mom, ref_state = R.densification({**attrs, **extras}, moments, state, CFG, duplicate, prune, samples)
    Pn = ref_attrs["position"].shape[0]
    assert den.P == Pn and Pn != P
    if duplicate:
        assert samples is not None and samples.shape[0] > 0          # the case exercises the split path
    for k in attrs:
        got, want = new_flat[k].detach().cpu(), ref_attrs[k]
        assert got.shape == want.shape, k
        assert torch.allclose(got, want, rtol=2e-6, atol=1e-7), k
        off = sum(new_flat.sizes[:new_flat.names.index(k)])
        n = want.numel()
        assert torch.equal(new_adam.exp_avg[off:off + n].cpu().view(want.shape), ref_mom[k][0]), k
        assert torch.equal(new_adam.exp_avg_sq[off:off + n].cpu().view(want.shape), ref_mom[k][1]), k
    assert torch.allclose(new_extras["position"].cpu(), ref_attrs["position"], rtol=2e-6, atol=1e-7)
    for k, t in (("grad_accum", den.grad_accum), ("denom", den.denom), ("max_radii", den.max_radii)):
        assert torch.allclose(t.cpu(), ref_state[k], rtol=1e-6, atol=1e-9), k
    assert new_adam.t == 7 and new_flat.flat_grad.abs().max() == 0
    # the new population trains: one fused Adam step runs over the regathered buffers
    new_flat.flat_grad.normal_()
    new_adam.step()
    torch.cuda.synchronize()
    assert torch.isfinite(new_flat.flat).all()


def test_reset_opacity_matches_reference_restatement(cuda):
    P = 4000
    attrs, extras, moments, _ = _population(P, seed=8)
    flat, adam, den = _build(cuda, attrs, extras, moments, P)
    den.reset_opacity(adam, cap=0.01)
    ref_attrs, ref_mom = R.reset_opacity({k: v.clone() for k, v in attrs.items()}, dict(moments))
    assert torch.allclose(flat["opacity"].detach().cpu(), ref_attrs["opacity"], rtol=1e-5, atol=1e-6)
    off = sum(flat.sizes[:flat.names.index("opacity")])
    assert float(adam.exp_avg[off:off + P].abs().max()) == 0 and float(adam.exp_avg_sq[off:off + P].abs().max()) == 0
    o2 = sum(flat.sizes[:flat.names.index("scaling")])
    assert torch.equal(adam.exp_avg[o2:o2 + 3 * P].cpu(), moments["scaling"][0].reshape(-1))   # other moments untouched
