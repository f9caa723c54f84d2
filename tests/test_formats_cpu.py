"""CPU: the reference's on-disk formats -- `model_*.pth` checkpoint layout and PLY attribute naming -- through
splatter_a_video_b200.formats.  The expected layouts are written out literally here from the reference sources
(trainer_fragGS.py:927-950, frag_model.py:345-353, base_model.py:178-188, points.py:397-465); the reference's Python stack
(omegaconf, plyfile) is not installable in this environment, so the checkpoint tests are layout tests ("parity unpinned"); the PLY
property naming / ordering / flattening is pinned to the reference's own save_ply / load_ply code (golden_ply.npz)."""
import os
import struct

import numpy as np
import pytest
import torch

from splatter_a_video_b200 import formats as F


def _raw_state(n=37, NI=3, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    # key order of the reference's state_dict (golden_checkpoint.pth): Parameters in registration order, the `position` buffer, num_pts
    pos = r(n, 3)
    return {"point_cloud.features": r(n, 1, 3), "point_cloud.features_rest": r(n, 15, 3),
            "point_cloud.scaling": r(n, 3) - 4, "point_cloud.rotation": r(n, 4), "point_cloud.opacity": r(n, 1),
            "point_cloud.pos_poly_feat": r(n, 4, 3), "point_cloud.pos_fourier_feat": r(n, 8, 3), "point_cloud.rot_poly_feat": 0.1 * r(n, 4, 4),
            "point_cloud.rot_fourier_feat": 0.1 * r(n, 8, 4), "point_cloud.pos_cubic_node": 0.02 * r(n, 4 * NI * 3),
            "point_cloud.mask_attribute": r(n, 1), "point_cloud.dino_attribute": r(n, 3), "point_cloud.position": pos, "num_pts": n}


def test_reads_a_reference_layout_checkpoint(tmp_path):
    sd = _raw_state()
    ckpt = {"gs_atlases_model": {"gs_atlas_0": sd}, "renderer": {"active_sh_degree": 2},
            "gs_atlas_0_optimizer": {"optimizer_1": {"state": {}, "param_groups": [{"name": "point_cloud.opacity", "lr": 0.05}]}}}
    path = str(tmp_path / "model_001500.pth")
    torch.save(ckpt, path)
    atlases, rnd, optim = F.load_checkpoint(path)
    assert list(atlases) == ["gs_atlas_0"] and rnd == {"active_sh_degree": 2} and list(optim) == ["gs_atlas_0_optimizer"]
    st = atlases["gs_atlas_0"]
    assert st.num_points == 37 and st.interval_num == 3 and F.checkpoint_step(path) == 1500
    assert st.order == [k[len("point_cloud."):] for k in sd if k not in ("num_pts", "point_cloud.position")]
    assert st.image_attributes() == ["mask_attribute", "dino_attribute"]
    for k, v in sd.items():
        if k != "num_pts":
            assert torch.equal(st.tensors[k[len("point_cloud."):]], v)


def test_checkpoint_round_trip_keeps_the_reference_key_layout(tmp_path):
    sd = _raw_state(n=11, NI=2, seed=3)
    p0 = str(tmp_path / "model_000100.pth")
    torch.save({"gs_atlases_model": {"a": sd, "b": _raw_state(n=5, NI=2, seed=4)}, "renderer": {"active_sh_degree": 3}}, p0)
    atlases, rnd, _ = F.load_checkpoint(p0)
    p1 = str(tmp_path / "sub" / "model_000200.pth")
    F.save_checkpoint(p1, atlases, rnd["active_sh_degree"], {"a": {"state": {}}})
    again = torch.load(p1, weights_only=False)
    assert set(again) == {"gs_atlases_model", "renderer", "a_optimizer"} and again["renderer"] == {"active_sh_degree": 3}
    assert list(again["gs_atlases_model"]["a"]) == list(sd)                 # same keys, same order, num_pts last
    for k, v in sd.items():
        got = again["gs_atlases_model"]["a"][k]
        assert (got == v) if k == "num_pts" else torch.equal(got, v)
    bad = dict(sd); bad["num_pts"] = 12
    torch.save({"gs_atlases_model": {"a": bad}}, p0)
    with pytest.raises(ValueError):
        F.load_checkpoint(p0)
    torch.save({"something": 1}, p0)
    with pytest.raises(ValueError):
        F.load_checkpoint(p0)


def test_render_dict_applies_the_reference_activations():
    sd = _raw_state(n=23, NI=4, seed=5)
    st = F.AtlasState({k[len("point_cloud."):]: v for k, v in sd.items() if k != "num_pts"})
    T, frame = 20, 13
    rd = st.render_dict(frame, T, fused=False)
    # get_position (dynamic_gaussian_with_base_point_cloud.py:236-250) written out independently
    intervals = torch.linspace(0, T - 1, 5).long() / (T - 1)
    nt = frame / (T - 1)
    i = max(int(torch.searchsorted(intervals, torch.tensor(nt - 1e-7))) - 1, 0)
    d = nt - float(intervals[i])
    c = sd["point_cloud.pos_cubic_node"].reshape(-1, 4, 4, 3)
    want = c[:, 3, i] + c[:, 2, i] * d + c[:, 1, i] * d ** 2 + c[:, 0, i] * d ** 3 + sd["point_cloud.position"]
    assert torch.allclose(rd["position"], want, atol=1e-6)
    t = frame / (T - 1)
    poly = torch.tensor([t ** k for k in range(4)], dtype=torch.float32)
    four = torch.cat([torch.cos(t * torch.arange(1, 5) * np.pi), torch.sin(t * torch.arange(1, 5) * np.pi)]).float()
    raw = sd["point_cloud.rotation"] + (sd["point_cloud.rot_poly_feat"] * poly[None, :, None]).sum(1) + (sd["point_cloud.rot_fourier_feat"] * four[None, :, None]).sum(1)
    assert torch.allclose(rd["rotation"], raw / raw.norm(dim=1, keepdim=True), atol=1e-6)
    assert torch.equal(rd["opacity"], torch.sigmoid(sd["point_cloud.opacity"])) and torch.equal(rd["scaling"], torch.exp(sd["point_cloud.scaling"]))
    assert rd["shs"].shape == (23, 16, 3) and torch.equal(rd["shs"][:, 0], sd["point_cloud.features"][:, 0])
    assert rd["pos_poly_feat"].shape == (23, 12) and rd["rot_fourier_feat"].shape == (23, 32)
    assert torch.equal(rd["mask_attribute"], torch.sigmoid(sd["point_cloud.mask_attribute"]))
    assert torch.equal(rd["dino_attribute"], torch.sigmoid(sd["point_cloud.dino_attribute"]))


def test_ply_layout_and_round_trip(tmp_path):
    st = F.AtlasState({"position": torch.tensor([[1.0, 2.0, 3.0], [4.0, 5.0, 6.0]]), "features": torch.tensor([[[0.1, 0.2, 0.3]], [[0.4, 0.5, 0.6]]]),
                       "opacity": torch.tensor([[7.0], [8.0]])})
    path = str(tmp_path / "pc" / "points.ply")
    F.save_ply(path, st)
    raw = open(path, "rb").read()
    header = (b"ply\nformat binary_little_endian 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\n"
              b"property float nx\nproperty float ny\nproperty float nz\nproperty float features_0\nproperty float features_1\n"
              b"property float features_2\nproperty float opacity_0\nend_header\n")
    assert raw.startswith(header)
    body = struct.unpack("<20f", raw[len(header):])
    assert np.allclose(body[:10], [1, 2, 3, 0, 0, 0, 0.1, 0.2, 0.3, 7]) and np.allclose(body[10:], [4, 5, 6, 0, 0, 0, 0.4, 0.5, 0.6, 8])
    back = F.load_ply(path, F.attribute_shapes(st))
    for k, v in st.tensors.items():
        assert torch.equal(back.tensors[k], v), k
    # a full atlas (all reference attributes) survives too, and the property list follows list_of_attributes
    sd = _raw_state(n=9, NI=2, seed=8)
    full = F.AtlasState({k[len("point_cloud."):]: v for k, v in sd.items() if k != "num_pts"})
    names = F.ply_property_names(full)
    assert names[:6] == ["x", "y", "z", "nx", "ny", "nz"] and names[6:9] == ["features_0", "features_1", "features_2"]
    assert names[9] == "features_rest_0" and names[-1] == "dino_attribute_2" and len(names) == 6 + 3 + 45 + 3 + 4 + 1 + 12 + 24 + 16 + 32 + 24 + 1 + 3
    p2 = str(tmp_path / "full.ply")
    F.save_ply(p2, full)
    again = F.load_ply(p2, F.attribute_shapes(full))
    for k, v in full.tensors.items():
        assert torch.equal(again.tensors[k], v), k


def test_reads_ascii_ply(tmp_path):
    path = str(tmp_path / "a.ply")
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment made by hand\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\n"
                "property float opacity_0\nend_header\n0 1 2 0.5\n3 4 5 0.25\n")
    st = F.load_ply(path, {"opacity": (1,)})
    assert torch.equal(st.tensors["position"], torch.tensor([[0.0, 1, 2], [3, 4, 5]])) and torch.equal(st.tensors["opacity"], torch.tensor([[0.5], [0.25]]))


def test_state_from_scene_inverts_the_activations():
    g = torch.Generator().manual_seed(1)
    n = 50
    shs, scaling, opacity = torch.randn(n, 16, 3, generator=g), torch.rand(n, 3, generator=g) * 0.01 + 1e-3, torch.rand(n, 1, generator=g) * 0.9 + 0.05
    rot = torch.nn.functional.normalize(torch.randn(n, 4, generator=g))
    st = F.state_from_scene(torch.randn(n, 3, generator=g), shs, scaling, rot, opacity, torch.zeros(n, 4, 2, 3),
                            {"mask_attribute": torch.rand(n, 1, generator=g) * 0.8 + 0.1, "track_like": torch.randn(n, 2, generator=g)})
    rd = st.render_dict(0, 10, fused=False)
    assert torch.allclose(rd["scaling"], scaling, rtol=1e-5) and torch.allclose(rd["opacity"], opacity, atol=1e-6) and torch.allclose(rd["shs"], shs)
    assert torch.allclose(rd["rotation"], rot, atol=1e-6) and rd["track_like"].shape == (n, 2) and st.interval_num == 2


def test_ply_matches_what_the_reference_code_writes_and_reads(tmp_path):
    """tests/golden/golden_ply.npz holds (a) the structured vertex array the reference's own save_ply / list_of_attributes
    (points.py:397-435) hand to plyfile for a population, (b) what the reference's own load_ply (:437-465) reconstructed from a file
    written by formats.save_ply (tests/golden/make_ply_golden.py executes those method bodies).  The product must write exactly
    (a) -- names, order, values -- and (b) must be the population itself."""
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_ply.npz"))
    names = [str(x) for x in G["names"]]
    attr = [k[3:] for k in G.files if k.startswith("in_") and k != "in_position"]
    tensors = {"position": torch.from_numpy(G["in_position"])}
    tensors.update({k: torch.from_numpy(G["in_" + k]) for k in attr})
    st = F.AtlasState(tensors, attr)
    assert F.ply_property_names(st) == names
    path = str(tmp_path / "pc.ply")
    F.save_ply(path, st)
    cols = F.read_ply_vertices(path)
    assert list(cols) == names
    got = np.stack([cols[n] for n in names], 1)
    assert got.dtype == np.float32 and np.array_equal(got, G["table"])
    for k in ["position"] + attr:
        assert np.array_equal(G["loaded_" + k], G["in_" + k]), k


def test_checkpoint_layout_matches_a_file_written_by_the_reference(tmp_path):
    """golden_checkpoint.pth was written by the reference's OWN save_model / get_state_dict / register_atribute bodies executed on
    stub modules (tests/golden/make_checkpoint_golden.py).  load_checkpoint must read it, and save_checkpoint must write the same
    layout back: top-level keys, per-atlas key ORDER, tensor values / dtypes, num_pts, renderer state, optimizer entries."""
    import torch
    from splatter_a_video_b200 import formats as F
    path = os.path.join(os.path.dirname(__file__), "golden", "golden_checkpoint.pth")
    ref = torch.load(path, map_location="cpu", weights_only=False)
    atlases, renderer, optim = F.load_checkpoint(path)
    assert list(atlases) == ["fg", "bg"] and renderer == {"active_sh_degree": 3} and sorted(optim) == ["bg_optimizer", "fg_optimizer"]
    for name, st in atlases.items():
        sd = ref["gs_atlases_model"][name]
        assert st.num_points == sd["num_pts"] == 16 and st.interval_num == 3
        assert set(st.tensors) == {k[len("point_cloud."):] for k in sd if k != "num_pts"}
        assert st.image_attributes() == ["mask_attribute", "dino_attribute"]
        for k, v in st.tensors.items():
            assert torch.equal(v, sd["point_cloud." + k])
    out = str(tmp_path / "model_000123.pth")
    F.save_checkpoint(out, atlases, active_sh_degree=renderer["active_sh_degree"], optimizers=optim)
    mine = torch.load(out, map_location="cpu", weights_only=False)
    assert list(mine) == list(ref)                                           # gs_atlases_model, renderer, fg_optimizer, bg_optimizer
    assert mine["renderer"] == ref["renderer"]
    for name in ref["gs_atlases_model"]:
        a, b = mine["gs_atlases_model"][name], ref["gs_atlases_model"][name]
        assert list(a) == list(b)                                            # parameters in registration order, position, num_pts
        for k in b:
            if k == "num_pts":
                assert a[k] == b[k]
            else:
                assert a[k].dtype == b[k].dtype and torch.equal(a[k], b[k])
    for k in optim:
        assert mine[k]["param_groups"] == ref[k]["param_groups"] and set(mine[k]["state"]) == set(ref[k]["state"])
    assert F.checkpoint_step(out) == 123
