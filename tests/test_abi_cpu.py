"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports every symbol include/spv_b200.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re
import subprocess

import pytest

import helpers as Hh

HDR = os.path.join(Hh.ROOT, "include", "spv_b200.h")


def declared_symbols():
    src = open(HDR).read()
    return sorted(set(re.findall(r"SPV_API\s+[\w\s\*]+?\b(spv_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from splatter_a_video_b200 import build
    return build.build()


def test_header_declares_the_reference_surface():
    syms = declared_symbols()
    # one entry per reference pybind function (ext.cpp:15-32); scan/sort replace compute_gaussian_key + range + torch.sort
    for base in ("project_point", "compute_cov3d", "ewa_project", "compute_sh", "alpha_blend"):
        assert f"spv_{base}_forward" in syms and f"spv_{base}_backward" in syms
    assert "spv_sort_gaussian" in syms and "spv_sort_scan" in syms


def test_library_exports_every_declared_symbol(lib_path):
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib_path], text=True)
    exported = set(re.findall(r"\bT (spv_\w+)", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, f"declared in the header but not exported: {missing}"
    extra = sorted(exported - set(declared_symbols()))
    assert not extra, f"exported but undeclared: {extra}"


def test_ctypes_binding_matches_header(lib_path):
    from splatter_a_video_b200 import _lib
    lib = _lib.load()
    assert lib.spv_abi_version() == 1
    assert sorted(_lib.EXPORTED) == declared_symbols()
    # (the CUB-backed sort/scan size queries need a device: covered by the -m gpu tests)
    assert lib.spv_alpha_blend_backward_workspace_bytes(1000, 19) == 1000 * 64 * 4


def test_library_is_sm100a_only(lib_path):
    out = subprocess.check_output(["cuobjdump", "-lelf", lib_path], text=True)
    archs = set(re.findall(r"sm_(\w+)\.cubin", out))
    assert archs == {"100a"}, archs


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(Hh.ROOT, "splatter_a_video_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "spv_oracle" not in txt, f


def test_ops_fail_loudly_without_cuda():
    import torch
    import dptr.gs as gs
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        gs.compute_cov3d(torch.ones(4, 3), torch.ones(4, 4))
    with pytest.raises(RuntimeError):
        gs.alpha_blending(torch.zeros(1, 2), torch.zeros(1, 3), torch.zeros(1, 1), torch.zeros(1, 3),
                          torch.zeros(1, dtype=torch.int32), torch.zeros(1, 2, dtype=torch.int32), 0.0, 16, 16)


def test_ctypes_signatures_have_the_header_arity():
    """Every ctypes argtypes list must have as many entries as the header's parameter list (a mismatch only shows on the GPU)."""
    from splatter_a_video_b200 import _lib
    hdr = re.sub(r"/\*.*?\*/", "", open(HDR).read(), flags=re.S)
    for name, (_, args) in _lib._SIGS.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", hdr, flags=re.S)
        assert m, f"{name} is bound but not declared"
        params = m.group(1).strip()
        n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
        assert n == len(args), f"{name}: header declares {n} parameters, ctypes binds {len(args)}"


def test_runtime_options_default_to_the_validated_kernels(lib_path):
    """spv_set_option is host-only state: known names are accepted, unknown ones rejected with an error message."""
    from splatter_a_video_b200 import _lib
    lib = _lib.load()
    assert lib.spv_set_option(b"bwd_variant", 0) == 0
    assert lib.spv_set_option(b"no_such_option", 1) != 0 and b"unknown option" in lib.spv_last_error()


def test_operator_surface_has_the_reference_signatures():
    """Boundary B1: `import dptr.gs as gs` must offer the reference's functions with the same positional parameters (count and
    order; a reference keyword name must be accepted too) and the same defaults -- tests/golden/golden_gs_signatures.json is read
    from the reference's sources (src/submodules/dptr/dptr/gs/*.py) by tests/golden/make_signature_golden.py."""
    import inspect
    import json
    import dptr.gs as gs
    G = json.load(open(os.path.join(Hh.ROOT, "tests", "golden", "golden_gs_signatures.json")))
    for name, ref in G["gs"].items():
        fn = getattr(gs, name, None)
        assert callable(fn), f"dptr.gs.{name} is missing ({ref['file']}:{ref['line']})"
        params = [p for p in inspect.signature(fn).parameters.values() if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        assert len(params) >= len(ref["params"]), f"{name}: takes {len(params)} positional parameters, the reference {len(ref['params'])}"
        for i, (rname, rdef, has) in enumerate(zip(ref["params"], ref["defaults"], ref["has_default"])):
            p = params[i]
            assert p.name == rname, f"{name}: parameter {i} is `{p.name}`, the reference calls it `{rname}`"
            if has:
                assert p.default is not inspect.Parameter.empty and p.default == rdef, f"{name}.{rname}: default {p.default!r} != {rdef!r}"
            else:
                assert p.default is inspect.Parameter.empty, f"{name}.{rname} is required in the reference"
        for p in params[len(ref["params"]):]:
            assert p.default is not inspect.Parameter.empty, f"{name}: extra parameter `{p.name}` must be optional"


def test_renderer_plugin_has_the_reference_methods():
    import inspect
    import json
    from splatter_a_video_b200.renderer import DPTROrthoEnhancedRender
    G = json.load(open(os.path.join(Hh.ROOT, "tests", "golden", "golden_gs_signatures.json")))
    for name, ref in G["renderer"].items():
        fn = getattr(DPTROrthoEnhancedRender, name, None)
        assert callable(fn), f"DPTROrthoEnhancedRender.{name} is missing (dptr_ortho_enhanced.py:{ref['line']})"
        params = list(inspect.signature(fn).parameters.values())
        names = [p.name for p in params if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)]
        if name == "render_iter":            # called as render_iter(**batch): every reference keyword must be accepted
            accepts_kw = any(p.kind == p.VAR_KEYWORD for p in params)
            assert all(r in names or accepts_kw for r in ref["params"]), (names, ref["params"])
        else:
            assert names[:len(ref["params"])] == ref["params"], f"{name}: {names} vs reference {ref['params']}"


def test_b2_shim_accepts_the_reference_call_site_keywords():
    """Boundary B2: every keyword src/pointrix/renderer/base_splatting.py:122-174 passes to GaussianRasterizationSettings,
    GaussianRasterizer and the rasterizer call must be accepted by the shim (golden_gs_signatures.json["b2"])."""
    import inspect
    import json
    import diff_gaussian_rasterization as D
    b2 = json.load(open(os.path.join(Hh.ROOT, "tests", "golden", "golden_gs_signatures.json")))["b2"]
    assert set(b2["settings_kwargs"]) <= set(D.GaussianRasterizationSettings._fields)
    assert set(b2["rasterizer_ctor_kwargs"]) <= set(inspect.signature(D.GaussianRasterizer.__init__).parameters)
    assert set(b2["call_kwargs"]) <= set(inspect.signature(D.GaussianRasterizer.forward).parameters)
