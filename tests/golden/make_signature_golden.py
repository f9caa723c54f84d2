"""Signatures of the reference's operator surface `dptr.gs` (boundary B1), read from its sources with `ast`:
/root/reference/src/submodules/dptr/dptr/gs/*.py -- the 9 public functions + `rasterization` -- as (parameter names, defaults).
Also the renderer classes' trainer-facing methods (/root/reference/src/pointrix/renderer/dptr_ortho_enhanced.py).
Output `golden_gs_signatures.json`, replayed by tests/test_abi_cpu.py against this repository's `dptr.gs` alias package.

    python tests/golden/make_signature_golden.py      (authoring container only: needs /root/reference)
"""
import ast
import glob
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
GS = "/root/reference/src/submodules/dptr/dptr/gs"
PUBLIC = ["project_point", "compute_cov3d", "ewa_project", "sort_gaussian", "compute_sh", "compute_sh_free", "alpha_blending",
          "alpha_blending_enhanced", "alpha_blending_with_bias", "rasterization"]


def sig(fn):
    a = fn.args
    names = [x.arg for x in a.args]
    defaults = [None] * (len(names) - len(a.defaults)) + [ast.literal_eval(d) if isinstance(d, (ast.Constant, ast.UnaryOp)) else ast.unparse(d)
                                                         for d in a.defaults]
    has_default = [False] * (len(names) - len(a.defaults)) + [True] * len(a.defaults)
    return {"params": names, "defaults": defaults, "has_default": has_default}


def main():
    out = {"gs": {}, "renderer": {}, "b2": {}}
    for path in sorted(glob.glob(os.path.join(GS, "*.py"))):
        for node in ast.parse(open(path).read()).body:
            if isinstance(node, ast.FunctionDef) and node.name in PUBLIC:
                out["gs"][node.name] = dict(sig(node), file=os.path.relpath(path, "/root/reference"), line=node.lineno)
    assert sorted(out["gs"]) == sorted(PUBLIC), sorted(set(PUBLIC) - set(out["gs"]))
    path = "/root/reference/src/pointrix/renderer/dptr_ortho_enhanced.py"
    cls = next(n for n in ast.parse(open(path).read()).body if isinstance(n, ast.ClassDef) and n.name == "DPTROrthoEnhancedRender")
    for fn in cls.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in ("project_point", "render_iter", "render_batch", "update_sh_degree", "load_state_dict",
                                                           "state_dict"):
            out["renderer"][fn.name] = dict(sig(fn), line=fn.lineno)
    # boundary B2: the keywords the one call site passes (src/pointrix/renderer/base_splatting.py:122-174)
    b2 = {}
    for node in ast.walk(ast.parse(open("/root/reference/src/pointrix/renderer/base_splatting.py").read())):
        if isinstance(node, ast.Call) and isinstance(node.func, ast.Name):
            if node.func.id == "GaussianRasterizationSettings":
                b2["settings_kwargs"], b2["settings_line"] = [k.arg for k in node.keywords], node.lineno
            elif node.func.id == "GaussianRasterizer":
                b2["rasterizer_ctor_kwargs"] = [k.arg for k in node.keywords]
            elif node.func.id == "rasterizer":
                b2["call_kwargs"], b2["call_line"] = [k.arg for k in node.keywords], node.lineno
    out["b2"] = b2
    json.dump(out, open(os.path.join(HERE, "golden_gs_signatures.json"), "w"), indent=1)
    for k, v in out["gs"].items():
        print(k, v["params"], [d for d, h in zip(v["defaults"], v["has_default"]) if h])


if __name__ == "__main__":
    main()
