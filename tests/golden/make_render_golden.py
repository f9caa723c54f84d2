"""Golden output dict of the trainer's active renderer, produced by EXECUTING THE REFERENCE'S OWN orchestration on the CPU:
`DPTROrthoEnhancedRender.render_batch` / `render_iter` / `project_point`, `ewa_project_torch_impl`
(/root/reference/src/pointrix/renderer/dptr_ortho_enhanced.py:17-111,145-433) and `RenderFeatures`
(/root/reference/src/pointrix/utils/renderer/renderer_utils.py:5-72), lifted from the sources with `ast` (the modules need omegaconf /
dptr to import).  What the reference gets from its CUDA module `dptr.gs` (compute_sh, compute_cov3d, sort_gaussian, alpha_blending,
alpha_blending_enhanced) is served by the C oracle (oracle/spv_oracle.c, itself pinned to the compiled reference kernels), forward
only; `Tensor.cuda()` is a no-op for the run.  So the golden fixes the ORCHESTRATION: which op gets which inputs, the three blend
passes with their backgrounds (cfg, 1.0, 0.0), K = num_idx, the detached opacity, the RenderFeatures channel order, the batch
assembly.  Output `golden_render.npz`, replayed by tests/test_oracle_cpu.py against oracle/torch_ref.render_ortho_frame -- the
restatement tests/test_renderer_gpu.py holds the renderer plugin to.

    python tests/golden/make_render_golden.py      (authoring container only: needs /root/reference)
"""
import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from splatter_a_video_b200 import synth  # noqa: E402

SRC = "/root/reference/src/pointrix/renderer/dptr_ortho_enhanced.py"
SRC_RF = "/root/reference/src/pointrix/utils/renderer/renderer_utils.py"
ATTRS = ["track_gs", "mask_attribute", "pos_poly_feat", "dino_attribute"]


def _strip(fn):
    fn.decorator_list, fn.returns = [], None
    for a in fn.args.args + fn.args.kwonlyargs:
        a.annotation = None
    return fn


def lift(gs):
    ns = {"torch": torch, "np": np, "BLOCK_X": 16, "BLOCK_Y": 16, "gs": gs}
    tree = ast.parse(open(SRC_RF).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "RenderFeatures")
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef):
            _strip(fn)
    exec(compile(ast.fix_missing_locations(ast.Module(body=[cls], type_ignores=[])), SRC_RF, "exec"), ns)
    tree = ast.parse(open(SRC).read())
    body = [_strip(n) for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "ewa_project_torch_impl"]
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "DPTROrthoEnhancedRender")
    body += [_strip(n) for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("project_point", "render_iter", "render_batch")]
    assert len(body) == 4
    for fn in body:
        exec(compile(ast.fix_missing_locations(ast.Module(body=[fn], type_ignores=[])), SRC, "exec"), ns)
    return ns


def oracle_gs():
    """`dptr.gs` for the lifted code: the C oracle behind the reference's Python signatures (forward only)."""
    t = torch.from_numpy
    gs = types.SimpleNamespace()

    def compute_sh(shs, degree, dirs, visible=None):
        rgb, _ = O.compute_sh(shs.detach().numpy(), degree, dirs.numpy())
        return t(rgb)

    def compute_cov3d(scaling, rotation, visible=None):
        return t(O.compute_cov3d(scaling.detach().numpy(), rotation.detach().numpy(), visible.reshape(-1).numpy()))

    def sort_gaussian(uv, depth, W, H, radius, tiles):
        idx, tr = O.sort_gaussian(uv.numpy(), depth.numpy(), W, H, radius.numpy(), tiles.numpy())
        return t(idx), t(tr)

    def alpha_blending(uv, conic, opacity, feature, idx, tr, bg, W, H, ndc=None, abs_ndc=None):
        f = O.alpha_blending_forward(uv.numpy(), conic.numpy(), opacity.detach().numpy(), feature.detach().numpy(), idx.numpy(), tr.numpy(),
                                     float(bg), W, H, K=0)
        return t(f["rendered"])

    def alpha_blending_enhanced(uv, conic, opacity, feature, idx, tr, bg, W, H, ndc=None, abs_ndc=None, K=10, enable_truncation=False):
        f = O.alpha_blending_forward(uv.numpy(), conic.numpy(), opacity.detach().numpy(), feature.detach().numpy(), idx.numpy(), tr.numpy(),
                                     float(bg), W, H, K=K)
        return t(f["rendered"]), t(f["ncontrib"]), t(f["gs_idx"])

    gs.compute_sh, gs.compute_cov3d, gs.sort_gaussian = compute_sh, compute_cov3d, sort_gaussian
    gs.alpha_blending, gs.alpha_blending_enhanced = alpha_blending, alpha_blending_enhanced
    return gs


def main():
    ns = lift(oracle_gs())
    sc = synth.make_config("cfg1_tiny")
    W, H = sc.W, sc.H
    host = types.SimpleNamespace(bg_color=0.0, cfg=types.SimpleNamespace(densify_abs_grad_enable=False))     # white_bg False, :139-143
    host.project_point = types.MethodType(ns["project_point"], host)
    host.render_iter = types.MethodType(ns["render_iter"], host)
    rd = {"position": sc.frame_position(0), "opacity": sc.opacity, "scaling": sc.scaling, "rotation": sc.rotation, "shs": sc.shs,
          "track_gs": sc.frame_position(1), "mask_attribute": sc.attrs["mask_attribute"], "pos_poly_feat": sc.attrs["pos_poly_feat"],
          "dino_attribute": sc.attrs["dino_attribute"]}
    batch = {"FovX": 0.0, "FovY": 0.0, "height": H, "width": W, "extrinsic_matrix": sc.extr, "intrinsic_matrix": sc.intr,
             "camera_center": torch.zeros(3), "render_attributes_list": list(ATTRS), "num_idx": 20}       # trainer_fragGS.py:451-462,510-512
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        out = ns["render_batch"](host, dict(rd), [batch])
    finally:
        torch.Tensor.cuda = orig_cuda
    g = {k: out[k].detach().numpy() for k in ["rgb", "depth"] + ATTRS}
    g.update(gs_idx=out["gs_idx"].numpy(), visibility=out["visibility"].numpy(), radii=out["radii"].numpy(),
             n_viewspace=np.array(len(out["viewspace_points"])), viewspace_shape=np.array(out["viewspace_points"][0].shape))
    np.savez_compressed(os.path.join(HERE, "golden_render.npz"), **g)
    print({k: v.shape for k, v in g.items()})
    print("wrote golden_render.npz", sum(v.nbytes for v in g.values()) // 1024, "KiB")


if __name__ == "__main__":
    main()
