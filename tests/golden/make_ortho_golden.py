"""Golden vectors for the ORTHOGRAPHIC geometry path of the trainer's active renderer, produced by EXECUTING THE REFERENCE'S OWN
torch code on the CPU: `DPTROrthoEnhancedRender.project_point` and `ewa_project_torch_impl`
(/root/reference/src/pointrix/renderer/dptr_ortho_enhanced.py:145-202 and :17-111).  The module cannot be imported here (omegaconf,
dptr are missing), so the two function bodies are lifted from the source file with `ast` (decorators / annotations dropped, bodies
unmodified).  Output `golden_ortho.npz` is committed and replayed by tests/test_oracle_cpu.py against the C oracle
(oracle/spv_oracle.c: project_point_ortho, ewa_project_ortho), which is the checker of the CUDA kernels.

    python tests/golden/make_ortho_golden.py      (authoring container only: needs /root/reference)
"""
import ast
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (only to produce cov3d inputs)
from splatter_a_video_b200 import synth  # noqa: E402

SRC = "/root/reference/src/pointrix/renderer/dptr_ortho_enhanced.py"


def lift():
    tree = ast.parse(open(SRC).read())
    ns = {"torch": torch, "BLOCK_X": 16, "BLOCK_Y": 16}
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "ewa_project_torch_impl"]
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "DPTROrthoEnhancedRender")
    fns += [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "project_point"]
    assert len(fns) == 2
    for fn in fns:
        fn.decorator_list, fn.returns = [], None
        for a in fn.args.args:
            a.annotation = None
        exec(compile(ast.fix_missing_locations(ast.Module(body=[fn], type_ignores=[])), SRC, "exec"), ns)
    return ns["project_point"], ns["ewa_project_torch_impl"]


def rot(ax, ay, az):
    cx, sx, cy, sy, cz, sz = math.cos(ax), math.sin(ax), math.cos(ay), math.sin(ay), math.cos(az), math.sin(az)
    Rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]); Ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = torch.tensor([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1.0]])
    return (Rz @ Ry @ Rx).float()


def main():
    project_point, ewa = lift()
    out = {}
    cases = {"identity": (synth.make_config("cfg1_tiny"), torch.eye(4)),
             "rotated": (synth.make_scene(3000, 2, 333, 250, seed=17), None)}
    E = torch.eye(4); E[:3, :3] = rot(0.05, -0.08, 0.3); E[:3, 3] = torch.tensor([0.02, -0.01, 0.1])
    cases["rotated"] = (cases["rotated"][0], E)
    for name, (sc, extr) in cases.items():
        W, H = sc.W, sc.H
        xyz = sc.frame_position(0).clone()
        xyz[::37, 2] = 0.005                      # behind the near plane
        xyz[5::41, 0] = 2.9                       # outside the 1.3 extent
        with torch.no_grad():
            uv, depth = project_point(None, xyz, extr, W, H, nearest=0.01)          # dptr_ortho_enhanced.py:281-283
            visible = depth != 0                                                    # :295
            cov3d = torch.from_numpy(O.compute_cov3d(sc.scaling.numpy(), sc.rotation.numpy(), visible.reshape(-1).numpy()))
            conic, radius, tiles = ewa(xyz, cov3d, extr, uv, W, H, visible.squeeze(-1))   # :305-312
        out.update({f"{name}_xyz": xyz.numpy(), f"{name}_extr": extr.numpy(), f"{name}_WH": np.array([W, H]),
                    f"{name}_cov3d": cov3d.numpy(), f"{name}_uv": uv.numpy(), f"{name}_depth": depth.numpy(),
                    f"{name}_conic": conic.numpy(), f"{name}_radius": radius.numpy(), f"{name}_tiles": tiles.numpy()})
        print(name, "P", xyz.shape[0], "visible", int(visible.sum()), "sum tiles", int(tiles.sum()))
    np.savez_compressed(os.path.join(HERE, "golden_ortho.npz"), **out)
    print("wrote golden_ortho.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB")


if __name__ == "__main__":
    main()
