"""-m gpu: a reference-layout checkpoint written to disk, read back and rendered by the sm_100a kernels gives the same images as
rendering the original tensors; the fused deformation ops agree with the torch formulas of AtlasState.render_dict(fused=False)."""
import numpy as np
import pytest
import torch

from splatter_a_video_b200 import formats as F
from splatter_a_video_b200 import synth

pytestmark = pytest.mark.gpu


def test_checkpoint_scene_renders_like_the_original(cuda, tmp_path):
    from splatter_a_video_b200.renderer import parse_renderer
    sc = synth.make_scene(5000, 10, 160, 96, seed=11)
    NI = 2
    g = torch.Generator().manual_seed(0)
    node = 0.02 * torch.randn(sc.P, 4, NI, 3, generator=g)
    st = F.state_from_scene(sc.position, sc.shs, sc.scaling, sc.rotation, sc.opacity, node,
                            {"mask_attribute": sc.attrs["mask_attribute"].clamp(0.05, 0.95), "dino_attribute": sc.attrs["dino_attribute"].clamp(0.05, 0.95)},
                            rot_poly_feat=0.05 * torch.randn(sc.P, 4, 4, generator=g), rot_fourier_feat=0.05 * torch.randn(sc.P, 8, 4, generator=g))
    path = str(tmp_path / "model_000300.pth")
    F.save_checkpoint(path, {"gs_atlas_0": st})
    atlases, rnd_state, _ = F.load_checkpoint(path, map_location="cpu")
    loaded = atlases["gs_atlas_0"].to(cuda)
    frame, T = 6, 10
    rd_fused = loaded.render_dict(frame, T, fused=True)
    rd_torch = st.render_dict(frame, T, fused=False)
    for k in rd_torch:
        assert torch.allclose(rd_fused[k].cpu(), rd_torch[k], atol=2e-6, rtol=1e-5), k
    rnd = parse_renderer({"name": "DPTROrthoEnhancedRenderB200"}, white_bg=False, device=cuda)
    rnd.load_state_dict(rnd_state)
    batch = {"height": sc.H, "width": sc.W, "extrinsic_matrix": sc.extr.to(cuda), "intrinsic_matrix": sc.intr.to(cuda),
             "camera_center": torch.zeros(3, device=cuda), "render_attributes_list": ["mask_attribute", "dino_attribute"], "num_idx": 10}
    with torch.no_grad():
        a = rnd.render_batch(rd_fused, [batch])
        b = rnd.render_batch({k: v.to(cuda) for k, v in rd_torch.items()}, [batch])
    for k in ["rgb", "depth", "mask_attribute", "dino_attribute"]:
        assert float((a[k] - b[k]).abs().max()) <= 1e-4, k
    assert float(a["rgb"].abs().max()) > 0.05          # the scene is actually visible
