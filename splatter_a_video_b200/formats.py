"""On-disk formats adjacent to the rasterizer: the reference's training checkpoints and PLY
point clouds, read and written without the reference's Python stack so that a reference-trained scene can be rendered by
these kernels (and a scene trained here opened by the reference).

Checkpoint `model_{step:06d}.pth` (trainer_fragGS.py:923-938, frag_model.py:345-347, pointrix/model/base_model.py:186-188; layout
pinned to a file written by those functions themselves: tests/golden/make_checkpoint_golden.py -> golden_checkpoint.pth):

    {"gs_atlases_model": {<atlas name>: {"point_cloud.position": [N,3] (frozen base cloud), "point_cloud.features": [N,1,3],
                                         "point_cloud.features_rest": [N,15,3], "point_cloud.scaling": [N,3] (log),
                                         "point_cloud.rotation": [N,4] (un-normalised), "point_cloud.opacity": [N,1] (logit),
                                         "point_cloud.pos_poly_feat" [N,4,3], ".pos_fourier_feat" [N,8,3], ".rot_poly_feat" [N,4,4],
                                         ".rot_fourier_feat" [N,8,4], ".pos_cubic_node": [N, 4*NI*3],
                                         "point_cloud.<image attribute>": [N,c] (pre-sigmoid), "num_pts": N}, ...},
     "renderer": {"active_sh_degree": d}, "<atlas name>_optimizer": <optimizer state dict>}

PLY (pointrix/point_cloud/points.py:397-465): one `vertex` element of float32 properties `x y z nx ny nz` followed by
`<attribute>_<i>` for every flattened channel of every non-position attribute in registration order; binary little endian
(plyfile's default for a natively little-endian array).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

PREFIX = "point_cloud."
# registration order of DynamicGaussianWithBasePointCloud.setup (dynamic_gaussian_with_base_point_cloud.py:82-161)
ATTRIBUTE_ORDER = ["features", "features_rest", "scaling", "rotation", "opacity", "pos_poly_feat", "pos_fourier_feat",
                   "rot_poly_feat", "rot_fourier_feat", "pos_cubic_node"]
SIGMOID_ATTRIBUTES = ("mask_attribute", "dino_attribute")      # :155-158, getters :276-282


@dataclass
class AtlasState:
    """Raw (pre-activation) per-Gaussian tensors of one atlas, keyed like the reference's attributes (no `point_cloud.` prefix)."""
    tensors: Dict[str, Tensor]
    order: List[str] = field(default_factory=list)        # attribute order for PLY export (position first, implicit)

    def __post_init__(self):
        if not self.order:
            known = [k for k in ATTRIBUTE_ORDER if k in self.tensors]
            self.order = known + [k for k in self.tensors if k not in known and k != "position"]

    @property
    def num_points(self) -> int:
        return int(self.tensors["position"].shape[0])

    @property
    def interval_num(self) -> int:
        return int(self.tensors["pos_cubic_node"].shape[1]) // 12

    def image_attributes(self) -> List[str]:
        return [k for k in self.order if k not in ATTRIBUTE_ORDER]

    def to(self, device) -> "AtlasState":
        return AtlasState({k: v.to(device) for k, v in self.tensors.items()}, list(self.order))

    # ---- what SingleAtlasWithBaseModel.forward(ids) hands the renderer (frag_model.py:112-137) ------------------------
    def render_dict(self, frame: int, num_frames: int, start_frame_id: int = 0, fused: bool = True) -> Dict[str, Tensor]:
        """Activated per-frame tensors.  `num_frames` is the clip length the model was built for (`len(delta_position)`,
        not stored in the checkpoint).  fused=True evaluates the spline / rotation with the CUDA deformation ops (CUDA tensors
        only); fused=False uses the same formulas in torch (any device; used by the CPU tests)."""
        t = self.tensors
        NI = self.interval_num
        from .gs.frame import rotation_basis, spline_interval
        idx, dist = spline_interval(frame, num_frames, NI)
        basis = rotation_basis(frame, start_frame_id, num_frames - 1)
        if fused:
            from .gs.frame import deform_position, deform_rotation
            dev = t["position"].device
            position = deform_position(t["position"], t["pos_cubic_node"], torch.tensor([idx], dtype=torch.int32, device=dev),
                                       torch.tensor([dist], dtype=torch.float32, device=dev), NI)
            rotation = deform_rotation(t["rotation"], t["rot_poly_feat"], t["rot_fourier_feat"], basis.to(dev))
        else:
            c = t["pos_cubic_node"].reshape(-1, 4, NI, 3)[:, :, idx]                                  # get_position :236-250
            position = c[:, 3] + c[:, 2] * dist + c[:, 1] * dist ** 2 + c[:, 0] * dist ** 3 + t["position"]
            b = basis.to(t["rotation"].device)
            raw = t["rotation"] + (t["rot_poly_feat"] * b[None, :4, None]).sum(1) + (t["rot_fourier_feat"] * b[None, 4:, None]).sum(1)
            rotation = torch.nn.functional.normalize(raw)                                              # get_rotation :184-198
        out = {"position": position, "detached_position": position.detach(), "opacity": torch.sigmoid(t["opacity"]),
               "scaling": torch.exp(t["scaling"]), "rotation": rotation, "shs": torch.cat([t["features"], t["features_rest"]], 1)}
        for k in ("pos_poly_feat", "pos_fourier_feat", "rot_poly_feat", "rot_fourier_feat"):
            out[k] = t[k].reshape(t[k].shape[0], -1)
        for k in self.image_attributes():
            out[k] = torch.sigmoid(t[k]) if k in SIGMOID_ATTRIBUTES else t[k]
        return out


# ------------------------------------------------------------------------------------------------------ checkpoints
def _is_atlas(v) -> bool:
    return isinstance(v, dict) and "num_pts" in v


def load_checkpoint(path: str, map_location="cpu") -> Tuple[Dict[str, AtlasState], dict, dict]:
    """Reads a reference `model_*.pth`.  Returns ({atlas name: AtlasState}, renderer state, {name: optimizer state dict})."""
    data = torch.load(path, map_location=map_location, weights_only=False)
    if "gs_atlases_model" not in data:
        raise ValueError(f"{path}: not a Splatter_A_Video checkpoint (no 'gs_atlases_model' entry)")
    atlases = {}
    for name, sd in data["gs_atlases_model"].items():
        if not _is_atlas(sd):
            continue
        tensors, order = {}, []
        for k, v in sd.items():
            if k == "num_pts":
                continue
            key = k[len(PREFIX):] if k.startswith(PREFIX) else k
            tensors[key] = v.detach() if isinstance(v, Tensor) else torch.as_tensor(v)
            if key != "position":
                order.append(key)
        if int(sd["num_pts"]) != int(tensors["position"].shape[0]):
            raise ValueError(f"{path}: atlas {name}: num_pts {sd['num_pts']} != {tensors['position'].shape[0]} points")
        atlases[name] = AtlasState(tensors, order)
    optim = {k: v for k, v in data.items() if k.endswith("_optimizer")}
    return atlases, dict(data.get("renderer", {})), optim


def save_checkpoint(path: str, atlases: Dict[str, AtlasState], active_sh_degree: int = 3, optimizers: Optional[dict] = None) -> None:
    """Writes the layout `load_model` of the reference trainer consumes (trainer_fragGS.py:941-950)."""
    model = {}
    for name, st in atlases.items():
        # nn.Module.state_dict order of the reference's point cloud: the Parameters in registration order, then its one remaining
        # buffer (the frozen base `position`), then `num_pts` (pinned by tests/golden/golden_checkpoint.pth)
        sd = {}
        for k in st.order:
            sd[PREFIX + k] = st.tensors[k].detach().cpu()
        sd[PREFIX + "position"] = st.tensors["position"].detach().cpu()
        sd["num_pts"] = st.num_points
        model[name] = sd
    data = {"gs_atlases_model": model, "renderer": {"active_sh_degree": int(active_sh_degree)}}
    for k, v in (optimizers or {}).items():
        data[k if k.endswith("_optimizer") else k + "_optimizer"] = v
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save(data, path)


def checkpoint_step(path: str) -> int:
    """`int(fpath[-10:-4])` of load_from_ckpt (trainer_fragGS.py:990)."""
    return int(os.path.basename(path)[-10:-4])


# ------------------------------------------------------------------------------------------------------ PLY
def ply_property_names(state: AtlasState) -> List[str]:
    """list_of_attributes (points.py:397-408)."""
    names = ["x", "y", "z", "nx", "ny", "nz"]
    for k in state.order:
        names += [f"{k}_{i}" for i in range(int(np.prod(state.tensors[k].shape[1:])))]
    return names


def save_ply(path: str, state: AtlasState) -> None:
    """save_ply (points.py:410-435): positions, zero normals, then every attribute flattened per point, all float32."""
    n = state.num_points
    cols = [state.tensors["position"].detach().cpu().numpy().astype("<f4"), np.zeros((n, 3), "<f4")]
    cols += [state.tensors[k].detach().cpu().reshape(n, -1).numpy().astype("<f4") for k in state.order]
    table = np.ascontiguousarray(np.concatenate(cols, 1))
    names = ply_property_names(state)
    assert table.shape[1] == len(names)
    header = "ply\nformat binary_little_endian 1.0\n" + f"element vertex {n}\n" + "".join(f"property float {p}\n" for p in names) + "end_header\n"
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "wb") as f:
        f.write(header.encode("ascii"))
        f.write(table.tobytes())


_PLY_TYPES = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1", "uint8": "u1", "char": "i1", "int8": "i1",
              "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4"}


def read_ply_vertices(path: str) -> Dict[str, np.ndarray]:
    """Minimal PLY reader (ascii / binary, scalar properties of the `vertex` element) -> {property name: array[N]}."""
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, props, n, in_vertex = None, [], 0, False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: truncated PLY header")
            tok = line.decode("ascii").split()
            if not tok or tok[0] == "comment":
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n = int(tok[2])
                elif n == 0:
                    raise ValueError(f"{path}: elements before `vertex` are not supported")
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties on vertices are not supported")
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt == "ascii":
            rows = np.loadtxt(f, max_rows=n, ndmin=2)
            return {name: rows[:, i].astype(t) for i, (name, t) in enumerate(props)}
        order = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(name, order + t) for name, t in props])
        rec = np.frombuffer(f.read(dt.itemsize * n), dtype=dt, count=n)
        return {name: np.ascontiguousarray(rec[name]) for name, _ in props}


def load_ply(path: str, shapes: Dict[str, Tuple[int, ...]]) -> AtlasState:
    """load_ply (points.py:437-465): `shapes` gives the per-point shape of every attribute to read, in order
    (e.g. {"features": (1,3), "features_rest": (15,3), "scaling": (3,), ...}); columns are `<name>_<i>`."""
    v = read_ply_vertices(path)
    tensors = {"position": torch.from_numpy(np.stack([v["x"], v["y"], v["z"]], 1).astype(np.float32))}
    for name, shp in shapes.items():
        k = int(np.prod(shp))
        cols = np.stack([v[f"{name}_{i}"] for i in range(k)], 1).astype(np.float32)
        tensors[name] = torch.from_numpy(cols.reshape(-1, *shp))
    return AtlasState(tensors, list(shapes.keys()))


def attribute_shapes(state: AtlasState) -> Dict[str, Tuple[int, ...]]:
    return {k: tuple(int(s) for s in state.tensors[k].shape[1:]) for k in state.order}


def state_from_scene(position: Tensor, shs: Tensor, scaling: Tensor, rotation: Tensor, opacity: Tensor, pos_cubic_node: Tensor,
                     attributes: Optional[Dict[str, Tensor]] = None, rot_poly_feat: Optional[Tensor] = None,
                     rot_fourier_feat: Optional[Tensor] = None) -> AtlasState:
    """Packs ACTIVATED tensors (what the renderer consumes) into the reference's raw storage: log scale, logit opacity,
    pre-sigmoid image attributes, SH split into DC / rest."""
    n = position.shape[0]
    logit = lambda x: torch.log(x / (1 - x))                          # inverse_sigmoid (pointrix/utils/gaussian_points/gaussian_utils.py)
    z = lambda *s: torch.zeros(n, *s, dtype=torch.float32, device=position.device)
    t = {"position": position, "features": shs[:, :1].contiguous(), "features_rest": shs[:, 1:].contiguous(), "scaling": torch.log(scaling),
         "rotation": rotation, "opacity": logit(opacity), "pos_poly_feat": z(4, 3), "pos_fourier_feat": z(8, 3),
         "rot_poly_feat": rot_poly_feat if rot_poly_feat is not None else z(4, 4),
         "rot_fourier_feat": rot_fourier_feat if rot_fourier_feat is not None else z(8, 4), "pos_cubic_node": pos_cubic_node.reshape(n, -1)}
    for k, v in (attributes or {}).items():
        t[k] = logit(v.clamp(1e-6, 1 - 1e-6)) if k in SIGMOID_ATTRIBUTES else v
    return AtlasState(t)
