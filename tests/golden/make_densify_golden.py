"""Golden vectors for the densification oracle, produced by EXECUTING THE REFERENCE'S OWN CODE on the CPU.

Runs only in the authoring container (needs /root/reference); the output `golden_densify.npz` is committed and replayed by
tests/test_oracle_cpu.py against oracle/densify_ref.py (which in turn is the checker of the CUDA path, tests/test_densify_gpu.py).

The reference modules cannot be imported here (omegaconf, plyfile, ... are missing), so the method bodies are lifted out of the
source files with `ast` -- unmodified except that decorators and type annotations are dropped -- and bound to minimal host objects:

  /root/reference/src/pointrix/optimizer/atlas_gs_optimizer.py   update_structure, densification, densify_clone, densify_split,
        prune, generate_clone_mask, generate_split_mask, new_pos_scale, prune_postprocess, reset_densification_state, reset_opacity
  /root/reference/src/pointrix/point_cloud/points.py             select_atributes, extand_points, remove_points, replace,
        extend_optimizer, prune_optimizer, replace_optimizer, unwarp, __len__
  /root/reference/src/pointrix/utils/gaussian_points/gaussian_utils.py   inverse_sigmoid, build_rotation
  /root/reference/src/pointrix/point_cloud/utils.py              unwarp_name
with a real torch.optim.Adam (one param group per attribute, named like src/pointrix/optimizer/__init__.py:41-45) and the
activations of src/pointrix/model/gaussian_points/gaussian_points.py:31-37,74-81.  `torch` is proxied only to (a) drop the
hard-coded device="cuda" and (b) record the torch.normal draw of new_pos_scale so the oracle consumes the same samples.

    python tests/golden/make_densify_golden.py
"""
import ast
import os
import types

import numpy as np
import torch
from torch import nn

REF = "/root/reference/src/pointrix"
HERE = os.path.dirname(os.path.abspath(__file__))


class TorchProxy:
    """`torch` for the lifted code: CPU instead of the hard-coded device="cuda", and the split's random draw is recorded."""

    def __init__(self):
        self.draws = []

    def __getattr__(self, name):
        return getattr(torch, name)

    def zeros(self, *a, **k):
        k.pop("device", None)
        return torch.zeros(*a, **k)

    def normal(self, *a, **k):
        out = torch.normal(*a, **k)
        self.draws.append(out.clone())
        return out


def lift(path, names, ns, class_name=None):
    """exec the named function definitions of `path` (module level, or methods of `class_name`) into `ns`."""
    tree = ast.parse(open(path).read())
    body = tree.body
    if class_name:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name).body
    found = set()
    for fn in body:
        if isinstance(fn, ast.FunctionDef) and fn.name in names:
            fn.decorator_list = []
            fn.returns = None
            for a in fn.args.args + fn.args.kwonlyargs:
                a.annotation = None
            exec(compile(ast.fix_missing_locations(ast.Module(body=[fn], type_ignores=[])), path, "exec"), ns)
            found.add(fn.name)
    missing = set(names) - found
    assert not missing, f"{path}: {missing} not found"


def build(P, seed):
    tp = TorchProxy()
    ns = {"torch": tp, "nn": nn, "print": lambda *a, **k: None}
    lift(f"{REF}/utils/gaussian_points/gaussian_utils.py", ["inverse_sigmoid", "build_rotation"], ns)
    lift(f"{REF}/point_cloud/utils.py", ["unwarp_name"], ns)
    cloud_fns = ["select_atributes", "extand_points", "remove_points", "replace", "extend_optimizer", "prune_optimizer",
                 "replace_optimizer", "unwarp", "__len__"]
    lift(f"{REF}/point_cloud/points.py", cloud_fns, ns, "PointCloud")
    opt_fns = ["update_structure", "densification", "densify_clone", "densify_split", "prune", "generate_clone_mask",
               "generate_split_mask", "new_pos_scale", "prune_postprocess", "reset_densification_state", "reset_opacity"]
    lift(f"{REF}/optimizer/atlas_gs_optimizer.py", opt_fns, ns, "AtlasGaussianSplattingOptimizer")

    g = torch.Generator().manual_seed(seed)
    attrs = {"position": torch.rand(P, 3, generator=g) * 2 - 1,
             "node": 0.01 * torch.randn(P, 24, generator=g),
             "scaling": torch.log(0.004 * torch.exp(1.2 * torch.randn(P, 3, generator=g))),
             "rotation": torch.randn(P, 4, generator=g),
             "opacity": 3.0 * torch.randn(P, 1, generator=g) - 2.0,
             "shs": torch.randn(P, 16, 3, generator=g)}

    class Cloud(nn.Module):                      # host object for the lifted PointCloud methods
        prefix_name = "point_cloud."

        def __init__(self):
            super().__init__()
            self.atributes = []
            for k, v in attrs.items():
                setattr(self, k, nn.Parameter(v.clone().requires_grad_(True)))
                self.atributes.append({"name": k, "trainable": True})
            self.scaling_inverse_activation = torch.log          # gaussian_points.py:31-37

        get_opacity = property(lambda self: torch.sigmoid(self.opacity))     # gaussian_points.py:74-81
        get_scaling = property(lambda self: torch.exp(self.scaling))

    for name in cloud_fns:
        setattr(Cloud, name, ns[name])
    cloud = Cloud()
    adam = torch.optim.Adam([{"params": [getattr(cloud, k)], "name": "point_cloud." + k, "lr": 1e-3} for k in attrs], eps=1e-15)
    for _ in range(2):                            # two steps so both moments are populated
        for k in attrs:
            getattr(cloud, k).grad = torch.randn(getattr(cloud, k).shape, generator=g)
        adam.step()

    class Opt:                                    # host object for the lifted optimizer methods (setup(), :60-83)
        pass

    for name in opt_fns:
        setattr(Opt, name, ns[name])
    o = Opt()
    o.optimizer, o.point_cloud, o.device = adam, cloud, "cpu"
    o.cameras_extent, o.percent_dense, o.split_num = 1.0, 0.01, 2
    o.max_radii2D = torch.zeros(P)
    o.pos_gradient_accum, o.denom = torch.zeros(P, 1), torch.zeros(P, 1)
    o.densify_grad_threshold, o.min_opacity = 0.0002, 0.005
    o.opacity_deferred = False
    o.opacity_reset_interval = 10 ** 9
    o.cfg = types.SimpleNamespace(densify_stop_iter=10 ** 9, densify_start_iter=10 ** 9)
    o.step = 1
    return o, cloud, adam, tp, g


def snapshot(cloud, adam):
    attrs = {a["name"]: getattr(cloud, a["name"]).detach().clone() for a in cloud.atributes}
    mom = {}
    for grp in adam.param_groups:
        st = adam.state.get(grp["params"][0])
        mom[grp["name"].replace("point_cloud.", "")] = (st["exp_avg"].clone(), st["exp_avg_sq"].clone())
    return attrs, mom


def main():
    out = {}
    P = 200
    for case, (duplicate, prune) in enumerate([(True, True), (True, False), (False, True)]):
        o, cloud, adam, tp, g = build(P, seed=40 + case)
        attrs0, mom0 = snapshot(cloud, adam)
        steps = []
        with torch.no_grad():
            for it in range(3):                   # statistics only (densify_start_iter is out of reach)
                vg = 0.0006 * torch.randn(P, 2, generator=g) * (torch.rand(P, 1, generator=g) < 0.5)
                radii = ((torch.rand(P, generator=g) * 30).int() * (torch.rand(P, generator=g) < 0.7).int()).float()
                vis = radii > 0
                o.update_structure(vis, vg, radii)
                o.step += 1
                steps.append((vg, radii, vis))
            state0 = (o.pos_gradient_accum.clone(), o.denom.clone(), o.max_radii2D.clone())
            o.duplicate_interval = 100 if duplicate else 7
            o.prune_interval = 100 if prune else 7
            o.densification(100)
        attrs1, mom1 = snapshot(cloud, adam)
        pre = f"c{case}_"
        out[pre + "flags"] = np.array([duplicate, prune])
        for k, v in attrs0.items():
            out[pre + "in_" + k] = v.numpy()
            out[pre + "in_m_" + k], out[pre + "in_v_" + k] = mom0[k][0].numpy(), mom0[k][1].numpy()
        for i, (vg, radii, vis) in enumerate(steps):
            out[pre + f"vg{i}"], out[pre + f"radii{i}"], out[pre + f"vis{i}"] = vg.numpy(), radii.numpy(), vis.numpy()
        for nm, t in zip(("accum", "denom", "maxr"), state0):
            out[pre + "state_" + nm] = t.numpy()
        out[pre + "samples"] = tp.draws[0].numpy() if tp.draws else np.zeros((0, 3), np.float32)
        for k, v in attrs1.items():
            out[pre + "out_" + k] = v.numpy()
            out[pre + "out_m_" + k], out[pre + "out_v_" + k] = mom1[k][0].numpy(), mom1[k][1].numpy()
        out[pre + "out_accum"], out[pre + "out_denom"], out[pre + "out_maxr"] = (o.pos_gradient_accum.numpy(), o.denom.numpy(),
                                                                                 o.max_radii2D.numpy())
        print(f"case {case} duplicate={duplicate} prune={prune}: {P} -> {attrs1['position'].shape[0]} points, "
              f"{out[pre + 'samples'].shape[0]} split samples")
    # reset_opacity (:185-197)
    o, cloud, adam, tp, g = build(P, seed=50)
    attrs0, mom0 = snapshot(cloud, adam)
    with torch.no_grad():
        o.reset_opacity()
    attrs1, mom1 = snapshot(cloud, adam)
    out["ro_in_opacity"], out["ro_out_opacity"] = attrs0["opacity"].numpy(), attrs1["opacity"].numpy()
    out["ro_out_m"], out["ro_out_v"] = mom1["opacity"][0].numpy(), mom1["opacity"][1].numpy()
    np.savez_compressed(os.path.join(HERE, "golden_densify.npz"), **out)
    print("wrote golden_densify.npz", sum(v.nbytes for v in out.values()) // 1024, "KiB")


if __name__ == "__main__":
    main()
