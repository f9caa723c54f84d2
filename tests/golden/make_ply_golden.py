"""PLY interchange pinned by EXECUTING THE REFERENCE'S OWN `PointCloud.list_of_attributes`, `save_ply` and `load_ply`
(/root/reference/src/pointrix/point_cloud/points.py:397-465) on the CPU, lifted from the source with `ast`.  `plyfile` is not
installed here, so its two entry points are replaced by the thinnest possible stand-ins: `PlyElement.describe` / `PlyData.write`
CAPTURE the structured array the reference hands over (property names, order and per-vertex values -- everything the reference
decides), and `PlyData.read` serves the vertex columns of a file written by the PRODUCT (`splatter_a_video_b200.formats.save_ply`)
to the reference's loader.  Output `golden_ply.npz`:
  names / table      what the reference's save_ply would write for the population (replayed against formats.save_ply)
  loaded_<attr>      what the reference's load_ply reconstructs from the product's file (must equal the population)

    python tests/golden/make_ply_golden.py      (authoring container only: needs /root/reference)
"""
import ast
import os
import sys
import tempfile
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from splatter_a_video_b200 import formats as F  # noqa: E402

SRC = "/root/reference/src/pointrix/point_cloud/points.py"
# registration order of the active model (dynamic_gaussian_with_base_point_cloud.py:100-158 after PointCloud.setup's position)
SHAPES = {"features": (1, 3), "features_rest": (15, 3), "scaling": (3,), "rotation": (4,), "opacity": (1,), "pos_poly_feat": (4, 3),
          "pos_fourier_feat": (8, 3), "rot_poly_feat": (4, 4), "rot_fourier_feat": (8, 4), "pos_cubic_node": (4 * 3 * 3,),
          "mask_attribute": (1,), "dino_attribute": (3,)}


def main():
    captured = {}

    class PlyElement:
        @staticmethod
        def describe(elements, name):
            captured["elements"], captured["element_name"] = elements, name
            return elements

    class PlyData:
        def __init__(self, els):
            self.els = els

        def write(self, path):
            captured["path"] = path

        @staticmethod
        def read(path):
            cols = F.read_ply_vertices(path)                 # the PRODUCT's file feeds the reference's loader
            return types.SimpleNamespace(elements=[cols])

    ns = {"np": np, "torch": torch, "nn": nn, "os": os, "PlyElement": PlyElement, "PlyData": PlyData, "mkdir_p": lambda p: None}
    tree = ast.parse(open(SRC).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "PointCloud")
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ("list_of_attributes", "save_ply", "load_ply"):
            fn.decorator_list, fn.returns = [], None
            for a in fn.args.args:
                a.annotation = None
            exec(compile(ast.fix_missing_locations(ast.Module(body=[fn], type_ignores=[])), SRC, "exec"), ns)

    g = torch.Generator().manual_seed(5)
    n = 23
    tensors = {"position": torch.randn(n, 3, generator=g)}
    tensors.update({k: torch.randn(n, *s, generator=g) for k, s in SHAPES.items()})
    cloud = types.SimpleNamespace(atributes=[{"name": "position"}] + [{"name": k} for k in SHAPES], cfg=types.SimpleNamespace(trainable=True))
    for k, v in tensors.items():
        setattr(cloud, k, v.clone())
    cloud.list_of_attributes = types.MethodType(ns["list_of_attributes"], cloud)
    ns["save_ply"](cloud, "/tmp/unused/ref.ply")
    el = captured["elements"]
    names = list(el.dtype.names)
    table = np.stack([el[nm] for nm in names], 1).astype(np.float32)
    assert captured["element_name"] == "vertex" and all(el.dtype[nm] == np.dtype("f4") for nm in names)

    # the product writes, the reference's load_ply reads
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "ours.ply")
        F.save_ply(path, F.AtlasState(dict(tensors), list(SHAPES)))
        fresh = types.SimpleNamespace(atributes=cloud.atributes, cfg=cloud.cfg)
        for k, v in tensors.items():
            setattr(fresh, k, torch.zeros_like(v))           # load_ply reads the per-point shapes from the existing attributes
        ns["load_ply"](fresh, path)
    out = {"names": np.array(names), "table": table}
    for k, v in tensors.items():
        out["in_" + k] = v.numpy()
        got = getattr(fresh, k).detach().numpy()
        assert np.array_equal(got, v.numpy()), k                # the reference reads back exactly what the product wrote
        out["loaded_" + k] = got
    np.savez_compressed(os.path.join(HERE, "golden_ply.npz"), **out)
    print(len(names), "properties:", names[:8], "...", names[-3:])
    print("wrote golden_ply.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB")


if __name__ == "__main__":
    main()
