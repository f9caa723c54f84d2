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


def lift(names):
    tree = ast.parse(open(SRC).read())
    ns = {"torch": torch, "np": np}
    for cls in (n for n in tree.body if isinstance(n, ast.ClassDef)):
        for fn in cls.body:
            if isinstance(fn, ast.FunctionDef) and fn.name in names and fn.name not in ns:
                fn.decorator_list, fn.returns = [], None
                for a in fn.args.args:
                    a.annotation = None
                exec(compile(ast.fix_missing_locations(ast.Module(body=[fn], type_ignores=[])), SRC, "exec"), ns)
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
    np.savez_compressed(os.path.join(HERE, "golden_deform.npz"), **out)
    print("wrote golden_deform.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB")


if __name__ == "__main__":
    main()
