"""Golden vectors for the per-frame deformation of the active model, produced by EXECUTING THE REFERENCE'S OWN code on the CPU:
`get_position(time)` (cubic-spline interval selection + evaluation) and `get_rotation(time)` of
/root/reference/src/dynamic_gaussian_with_base_point_cloud.py:184-198,236-250, lifted from the source with `ast` (the module
itself needs pointrix / omegaconf / simple_knn to import) and bound to a host object that carries exactly the attributes the two
methods read; `intervals` is built as the class does it (:66-68).  Output `golden_deform.npz`, replayed by tests/test_oracle_cpu.py
against the product's HOST logic (gs.frame.spline_interval / rotation_basis) plus the evaluation formulas the GPU tests use.

    python tests/golden/make_deform_golden.py      (authoring container only: needs /root/reference)
"""
import ast
import math
import os
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/dynamic_gaussian_with_base_point_cloud.py"


SRC_ALT = "/root/reference/src/dynamic_gaussian_points.py"


def lift(names, src=None):
    SRC_ = src or SRC
    tree = ast.parse(open(SRC_).read())
    ns = {"torch": torch, "np": np}
    for cls in (n for n in tree.body if isinstance(n, ast.ClassDef)):
        for fn in cls.body:
            if isinstance(fn, ast.FunctionDef) and fn.name in names and fn.name not in ns:
                fn.decorator_list, fn.returns = [], None
                for a in fn.args.args:
                    a.annotation = None
                exec(compile(ast.fix_missing_locations(ast.Module(body=[fn], type_ignores=[])), SRC_, "exec"), ns)
    assert all(n in ns for n in names)
    return [ns[n] for n in names]


def main():
    get_position, get_rotation = lift(["get_position", "get_rotation"])
    out = {}
    g = torch.Generator().manual_seed(77)
    P = 64
    for F in (50, 80, 7):
        NI = math.ceil(F / 5)                                              # :65
        o = types.SimpleNamespace()
        o.delta_position = torch.zeros(F, P, 3)                            # only its length is read (:239)
        o.interval_num = NI
        intervals_idx = torch.linspace(0, F - 1, NI + 1).long()            # :66-68
        o.intervals = intervals_idx / (F - 1)
        o.position = torch.randn(P, 3, generator=g)
        o.pos_cubic_node = 0.1 * torch.randn(P, 4 * NI * 3, generator=g)
        o.rotation = torch.randn(P, 4, generator=g)
        o.rot_poly_feat = 0.1 * torch.randn(P, 4, 4, generator=g)
        o.rot_fourier_feat = 0.1 * torch.randn(P, 8, 4, generator=g)
        o.poly_feature_dim, o.fourier_feature_dim = 4, 8                   # :133-134
        o.start_frame_id, o.time_len = 0, F - 1                            # :108-110
        o.rotation_activation = torch.nn.functional.normalize             # gaussian_points.py:37
        with torch.no_grad():
            pos = torch.stack([get_position(o, t) for t in range(F)])
            rot = torch.stack([get_rotation(o, t) for t in range(F)])
        pre = f"F{F}_"
        out.update({pre + "position": o.position.numpy(), pre + "node": o.pos_cubic_node.numpy(), pre + "rotation": o.rotation.numpy(),
                    pre + "rot_poly": o.rot_poly_feat.numpy(), pre + "rot_fourier": o.rot_fourier_feat.numpy(),
                    pre + "pos_t": pos.numpy(), pre + "rot_t": rot.numpy()})
        print(f"F={F} NI={NI}: positions {tuple(pos.shape)}, rotations {tuple(rot.shape)}")
    # the ALTERNATIVE model (dynamic_gaussian_points.py:170-186): polynomial + Fourier position, with and without detach_pos
    (alt_position,) = lift(["get_position"], SRC_ALT)
    for F in (50, 7):
        o = types.SimpleNamespace()
        o.position = torch.randn(P, 3, generator=g)
        o.pos_poly_feat = 0.1 * torch.randn(P, 4, 3, generator=g)
        o.pos_fourier_feat = 0.1 * torch.randn(P, 8, 3, generator=g)
        o.poly_feature_dim, o.fourier_feature_dim = 4, 8
        o.start_frame_id, o.time_len = 0, F - 1
        with torch.no_grad():
            pos = torch.stack([alt_position(o, t) for t in range(F)])
        # gradients of sum(pos * w) at one frame, through the reference's own expression
        leaves = [x.clone().requires_grad_(True) for x in (o.position, o.pos_poly_feat, o.pos_fourier_feat)]
        o2 = types.SimpleNamespace(**{**o.__dict__, "position": leaves[0], "pos_poly_feat": leaves[1], "pos_fourier_feat": leaves[2]})
        wgt = torch.randn(P, 3, generator=g)
        t_g = F // 3
        (alt_position(o2, t_g) * wgt).sum().backward()
        pre = f"ALT{F}_"
        out.update({pre + "position": o.position.numpy(), pre + "poly": o.pos_poly_feat.numpy(), pre + "fourier": o.pos_fourier_feat.numpy(),
                    pre + "pos_t": pos.numpy(), pre + "w": wgt.numpy(), pre + "t_grad": np.array(t_g),
                    pre + "g_position": leaves[0].grad.numpy(), pre + "g_poly": leaves[1].grad.numpy(), pre + "g_fourier": leaves[2].grad.numpy()})
        print(f"alternative model F={F}: positions {tuple(pos.shape)}")
    np.savez_compressed(os.path.join(HERE, "golden_deform.npz"), **out)
    print("wrote golden_deform.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB")


if __name__ == "__main__":
    main()
