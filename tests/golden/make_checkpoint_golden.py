"""Golden checkpoint produced by EXECUTING THE REFERENCE'S OWN code on the CPU: `FragTrainer.save_model`
(/root/reference/src/trainer_fragGS.py:928-939), `FragModel.get_state_dict` (src/frag_model.py:345-347),
`BaseModel.get_state_dict` (src/pointrix/model/base_model.py:186-188), `PointCloud.register_atribute` / `__len__`
(src/pointrix/point_cloud/points.py:99-132) and the renderer's `state_dict` (src/pointrix/renderer/dptr_ortho_enhanced.py:443-444),
lifted from the sources with `ast` (the modules themselves need omegaconf / simple_knn / pytorch3d to import) and bound to stub
`nn.Module`s with the reference's nesting (trainer -> gs_atlases_model -> atlas_dict[name] -> point_cloud) and the attribute
registration order of DynamicGaussianWithBasePointCloud.setup (src/dynamic_gaussian_with_base_point_cloud.py:85-163).  The
optimizers are real `torch.optim.Adam`s with one param group per attribute (src/pointrix/optimizer/__init__.py:27-62).

Output: tests/golden/golden_checkpoint.pth (16 points, 2 atlases; what `torch.save` wrote) -- replayed by
tests/test_formats_cpu.py against splatter_a_video_b200.formats.load_checkpoint / save_checkpoint.

    python tests/golden/make_checkpoint_golden.py      (authoring container only: needs /root/reference)
"""
import ast
import os
import types

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src"


def lift(path, cls_name, names, base=object, extra_ns=None):
    """A new class `Lifted(base)` whose methods `names` are the reference's own function bodies (compiled inside a class statement,
    so their zero-argument `super()` calls resolve against `base`)."""
    tree = ast.parse(open(path).read())
    fns = []
    for cls in (n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls_name):
        for fn in cls.body:
            if isinstance(fn, ast.FunctionDef) and fn.name in names:
                fn.decorator_list, fn.returns = [], None
                for a in fn.args.args:
                    a.annotation = None
                fns.append(fn)
    missing = set(names) - {f.name for f in fns}
    assert not missing, (path, cls_name, missing)
    klass = ast.ClassDef(name="Lifted", bases=[ast.Name(id="_Base", ctx=ast.Load())], keywords=[], body=fns, decorator_list=[], type_params=[])
    ns = {"torch": torch, "nn": nn, "_Base": base}
    ns.update(extra_ns or {})
    exec(compile(ast.fix_missing_locations(ast.Module(body=[klass], type_ignores=[])), path, "exec"), ns)
    return ns["Lifted"]


def main():
    class PointCloudBase(nn.Module):                  # pointrix PointCloud: buffers for position / features, Parameters after
        def __init__(self):
            super().__init__()
            self.cfg = types.SimpleNamespace(trainable=True)
            self.atributes = []
    PointCloud = lift(f"{REF}/pointrix/point_cloud/points.py", "PointCloud", ["register_atribute", "__len__"], PointCloudBase)

    class AtlasBase(nn.Module):                       # BaseModel: self.point_cloud is a registered sub-module (base_model.py:57)
        def __init__(self, pc):
            super().__init__()
            self.point_cloud = pc
    Atlas = lift(f"{REF}/pointrix/model/base_model.py", "BaseModel", ["get_state_dict"], AtlasBase)

    class FragBase(nn.Module):                        # FragModel: its setup() only fills a PLAIN dict of atlases (frag_model.py:243-254),
        def __init__(self, atlases):                  # no point_cloud of its own -> its nn.Module state_dict is empty
            super().__init__()
            self.atlas_dict = atlases
    Frag = lift(f"{REF}/frag_model.py", "FragModel", ["get_state_dict"], FragBase)
    Trainer = lift(f"{REF}/trainer_fragGS.py", "FragTrainer", ["save_model"], object, {"Path": str})
    Renderer = lift(f"{REF}/pointrix/renderer/dptr_ortho_enhanced.py", "DPTROrthoEnhancedRender", ["state_dict"], object)

    g = torch.Generator().manual_seed(2024)
    N, NI = 16, 3

    def atlas():
        pc = PointCloud()
        # points.py:51-52 register position / features as buffers first; then the trainable registration of setup(): position is
        # re-registered frozen (dynamic_gaussian_with_base_point_cloud.py:97-99), features as a Parameter
        pc.register_buffer("position", torch.randn(N, 3, generator=g))
        pc.register_atribute("features", torch.randn(N, 1, 3, generator=g))
        for name, shape in (("features_rest", (N, 15, 3)), ("scaling", (N, 3)), ("rotation", (N, 4)), ("opacity", (N, 1)),
                            ("pos_poly_feat", (N, 4, 3)), ("pos_fourier_feat", (N, 8, 3)), ("rot_poly_feat", (N, 4, 4)),
                            ("rot_fourier_feat", (N, 8, 4)), ("pos_cubic_node", (N, 4 * NI * 3)), ("mask_attribute", (N, 1)),
                            ("dino_attribute", (N, 3))):
            pc.register_atribute(name, torch.randn(*shape, generator=g))
        return Atlas(pc)

    atlases = {"fg": atlas(), "bg": atlas()}
    trainer = Trainer()
    trainer.gs_atlases_model = Frag(atlases)
    trainer.renderer = Renderer()
    trainer.renderer.active_sh_degree = 3
    trainer.gs_atlas_cfg_list = [types.SimpleNamespace(name="fg"), types.SimpleNamespace(name="bg")]
    for name, a in atlases.items():
        groups = [{"params": [p], "lr": 1e-3, "name": k} for k, p in a.point_cloud.named_parameters()]
        opt = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        for p in a.point_cloud.parameters():
            p.grad = torch.randn(p.shape, generator=g)
        opt.step()
        setattr(trainer, name + "_optimizer", opt)
    out = os.path.join(HERE, "golden_checkpoint.pth")
    trainer.save_model(out)
    data = torch.load(out, weights_only=False)
    print("top-level keys:", list(data))
    print("fg keys:", list(data["gs_atlases_model"]["fg"]))
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
