"""ctypes binding of libspv_b200.so (the C ABI declared in include/spv_b200.h).

The product has NO CPU fallback: if the library is missing or a CUDA call fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libspv_b200.so")
_lib = None

P_ = c_void_p
_SIGS = {
    # name: (restype, [argtypes])
    "spv_abi_version": (c_int, []),
    "spv_last_error": (ctypes.c_char_p, []),
    "spv_launch_count": (ctypes.c_longlong, []),
    "spv_set_option": (c_int, [ctypes.c_char_p, c_int]),
    "spv_kernel_timer_enable": (c_int, [c_int]),
    "spv_kernel_timer_read": (c_int, [c_int, P_]),
    "spv_project_point_forward": (c_int, [c_int, P_, P_, P_, c_int, c_int, c_float, c_float, P_, P_, P_]),
    "spv_project_point_backward": (c_int, [c_int, P_, P_, P_, P_, P_, P_, P_, P_, P_, P_]),
    "spv_project_point_ortho_forward": (c_int, [c_int, P_, P_, c_int, c_int, c_float, c_float, P_, P_, P_]),
    "spv_project_point_ortho_backward": (c_int, [c_int, P_, c_int, c_int, P_, P_, P_, P_, P_]),
    "spv_compute_cov3d_forward": (c_int, [c_int, P_, P_, P_, P_, P_]),
    "spv_compute_cov3d_backward": (c_int, [c_int, P_, P_, P_, P_, P_, P_, P_]),
    "spv_ewa_project_forward": (c_int, [c_int, P_, P_, P_, P_, P_, c_int, c_int, P_, P_, P_, P_, P_]),
    "spv_ewa_project_backward": (c_int, [c_int, P_, P_, P_, P_, P_, P_, P_, P_, P_, P_, P_]),
    "spv_ewa_project_ortho_forward": (c_int, [c_int, P_, P_, P_, c_int, c_int, P_, P_, P_, P_, P_]),
    "spv_ewa_project_ortho_backward": (c_int, [c_int, P_, P_, c_int, c_int, P_, P_, P_, P_]),
    "spv_compute_sh_forward": (c_int, [c_int, P_, c_int, P_, P_, c_int, P_, P_, P_]),
    "spv_compute_sh_backward": (c_int, [c_int, P_, c_int, P_, P_, P_, P_, c_int, P_, P_, P_]),
    "spv_compute_sh_z_forward": (c_int, [c_int, P_, P_, P_, P_]),
    "spv_compute_sh_z_backward": (c_int, [c_int, P_, P_, P_, P_]),
    "spv_sort_scan_workspace_bytes": (c_size_t, [c_int]),
    "spv_sort_scan": (c_int, [c_int, P_, P_, P_, c_size_t, P_]),
    "spv_sort_workspace_bytes": (c_size_t, [c_int, c_int64]),
    "spv_sort_gaussian": (c_int, [c_int, c_int64, P_, P_, P_, P_, c_int, c_int, P_, P_, P_, c_size_t, P_]),
    "spv_alpha_blend_forward": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, P_, P_, P_, P_, P_, P_, P_, c_float,
                                        P_, P_, P_, P_, P_]),
    "spv_alpha_blend_backward_workspace_bytes": (c_size_t, [c_int, c_int]),
    "spv_alpha_blend_backward": (c_int, [c_int, c_int, c_int, c_int, P_, P_, P_, P_, P_, P_, P_, c_float, P_, P_, P_,
                                         P_, P_, P_, P_, P_, P_, P_, c_size_t, P_]),
    "spv_alpha_blend_groups_forward": (c_int, [c_int, c_int, c_int, c_int, c_int, P_, P_, P_, P_, P_, P_, c_float, c_float,
                                               c_float, P_, P_, P_, P_, P_]),
    "spv_alpha_blend_groups_backward_workspace_bytes": (c_size_t, [c_int]),
    "spv_alpha_blend_groups_backward": (c_int, [c_int, c_int, c_int, c_int, P_, P_, P_, P_, P_, P_, c_float, c_float, c_float,
                                                P_, P_, P_, P_, P_, P_, P_, P_, P_, P_, c_size_t, P_]),
    "spv_bin_capacity_workspace_bytes": (c_size_t, [c_int, c_int64]),
    "spv_bin_capacity": (c_int, [c_int, c_int64, P_, P_, P_, P_, P_, c_int, c_int, c_int, P_, P_, P_, P_, c_size_t, P_]),
    "spv_bin_tiles_workspace_bytes": (c_size_t, [c_int, c_int64, c_int, c_int]),
    "spv_bin_tiles": (c_int, [c_int, c_int64, P_, P_, P_, P_, P_, c_int, c_int, c_int, P_, P_, P_, P_, c_size_t, P_]),
    "spv_frame_workspace_bytes": (c_size_t, [c_int, c_int64, c_int, c_int, c_int]),
    "spv_frame_ortho_forward": (c_int, [c_int, c_int, c_int, c_int, P_, P_, c_int, c_int64, c_int, P_, P_, P_, P_, P_, c_int, P_, c_float,
                                        c_float, c_float, P_, P_, P_, P_, P_, c_size_t, P_]),
    "spv_frame_ortho_backward": (c_int, [c_int, c_int, c_int, c_int, P_, c_int, c_int64, P_, P_, P_, P_, c_int, P_, c_float, P_, P_, P_, P_,
                                         P_, P_, P_, P_, P_, P_, P_, c_int, P_, c_size_t, P_]),
    "spv_alpha_blend_groups_backward_packed": (c_int, [c_int, c_int, c_int, c_int, P_, P_, P_, P_, P_, P_, c_float, c_float, c_float,
                                                       P_, P_, P_, c_int, P_, P_]),
    "spv_deform_spline_forward": (c_int, [c_int, c_int, c_int, P_, P_, P_, P_, P_, P_]),
    "spv_deform_spline_backward": (c_int, [c_int, c_int, c_int, P_, P_, P_, P_, c_int, P_]),
    "spv_deform_spline_forward2": (c_int, [c_int, c_int, c_int, P_, P_, P_, P_, P_, P_, P_, P_, P_]),
    "spv_deform_spline_backward2": (c_int, [c_int, c_int, c_int, P_, P_, P_, P_, P_, P_, P_, P_, P_]),
    "spv_deform_defer": (c_int, [c_int, P_, P_, P_, P_, P_, P_, P_, P_]),
    "spv_deform_spline_backward_gathered": (c_int, [c_int, c_int, c_int, c_int, P_, ctypes.c_longlong, c_float, P_, P_, P_]),
    "spv_deform_rotation_forward": (c_int, [c_int, P_, P_, P_, P_, P_, P_, P_]),
    "spv_deform_rotation_backward": (c_int, [c_int, P_, P_, P_, P_, P_]),
    "spv_deform_polyfourier_forward": (c_int, [c_int, P_, P_, P_, P_, P_, P_]),
    "spv_deform_polyfourier_backward": (c_int, [c_int, P_, P_, P_, P_, P_, P_]),
    "spv_adam_step": (c_int, [ctypes.c_longlong, P_, P_, P_, P_, c_int, P_, P_, c_float, c_float, c_float, c_int, P_]),
    "spv_adam_step_device": (c_int, [ctypes.c_longlong, P_, P_, P_, P_, c_int, P_, P_, c_float, c_float, c_float, P_, P_]),
    "spv_adam_step_lazy": (c_int, [ctypes.c_longlong, P_, P_, P_, P_, c_int, P_, P_, c_float, c_float, c_float, P_, c_int, c_int, c_int, c_int,
                                   P_, P_, P_, P_]),
    "spv_adam_lazy_prepare": (c_int, [ctypes.c_longlong, c_int, P_, c_int, c_int, c_int, c_int, P_, P_, P_, P_, P_, P_, P_, P_, P_, c_float,
                                      c_float, c_float, P_]),
    "spv_adam_lazy_flush": (c_int, [ctypes.c_longlong, c_int, P_, c_int, c_int, c_int, c_int, P_, P_, P_, P_, P_, P_, P_, c_float, c_float,
                                    c_float, P_]),
    "spv_densify_stats": (c_int, [c_int, P_, P_, P_, P_, P_, P_, P_]),
    "spv_densify_flags": (c_int, [c_int, P_, P_, P_, P_, P_, c_int, c_int, c_float, c_float, c_float, c_float, c_float, P_, P_]),
    "spv_flat_regather": (c_int, [c_int, P_, P_, P_, c_int, P_, P_, P_, P_]),
    "spv_split_children": (c_int, [c_int, P_, P_, P_, P_, P_, P_, c_int, c_float, P_, P_, P_]),
    "spv_reset_opacity": (c_int, [c_int, c_float, c_int, P_, P_, P_, P_]),
    "spv_exchange_sizes": (c_int, [c_int, c_int, P_, P_, P_]),
    "spv_exchange_pack": (c_int, [c_int, c_int, P_, P_, P_, c_float, P_, P_, P_]),
    "spv_exchange_reduce": (c_int, [ctypes.c_longlong, c_int, P_, ctypes.c_longlong, c_float, P_, P_]),
    "spv_exchange_reduce_peers": (c_int, [ctypes.c_longlong, ctypes.c_longlong, c_int, P_, c_float, P_, P_, ctypes.c_longlong, P_]),
    "spv_exchange_reduce_scatter_peers": (c_int, [ctypes.c_longlong, ctypes.c_longlong, c_int, c_int, P_, c_float, P_, P_, P_,
                                                  ctypes.c_longlong, P_]),
    "spv_exchange_gather_peers": (c_int, [ctypes.c_longlong, ctypes.c_longlong, c_int, P_, P_, ctypes.c_longlong, P_]),
    "spv_exchange_nvls": (c_int, [ctypes.c_longlong, ctypes.c_longlong, c_int, c_int, P_, P_, P_, c_float, P_, ctypes.c_longlong, P_]),
    "spv_exchange_publish": (c_int, [ctypes.c_longlong, ctypes.c_longlong, P_, P_, P_, P_]),
    "spv_exchange_fetch_reduced": (c_int, [ctypes.c_longlong, c_int, c_int, P_, P_, P_]),
    "spv_exchange_unpack": (c_int, [c_int, c_int, P_, c_int, P_, P_, P_, P_, P_]),
    "spv_loss_rgb_workspace_bytes": (ctypes.c_size_t, [c_int, c_int]),
    "spv_loss_rgb": (c_int, [c_int, c_int, P_, P_, c_float, c_float, P_, P_, P_, ctypes.c_size_t, P_]),
    "spv_loss_depth_workspace_bytes": (ctypes.c_size_t, [c_int]),
    "spv_loss_depth_dpt": (c_int, [c_int, P_, P_, c_float, P_, P_, P_, ctypes.c_size_t, P_]),
    "spv_loss_track_workspace_bytes": (ctypes.c_size_t, [c_int]),
    "spv_loss_track": (c_int, [c_int, c_int, c_int, P_, P_, P_, P_, P_, c_float, c_float, P_, P_, P_, ctypes.c_size_t, P_]),
}

EXPORTED = sorted(_SIGS)


def load():
    """dlopen the library (building is __graft_entry__.build()'s / build.py's job) and type every symbol."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m splatter_a_video_b200.build` "
                "(there is no CPU fallback for the rasterizer).")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
            fn.restype = res
            fn.argtypes = args
        if lib.spv_abi_version() != 1:
            raise RuntimeError("libspv_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name: str, *args):
    """Invoke an int-returning entry point; raise with the library's message on a CUDA error."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise RuntimeError(f"{name} failed: {lib.spv_last_error().decode()}")


def set_option(name: str, value: int):
    """Runtime switch of an experimental kernel variant (include/spv_b200.h: spv_set_option); 0 restores the validated default."""
    call("spv_set_option", name.encode(), int(value))


def query(name: str, *args) -> int:
    return int(getattr(load(), name)(*args))


def need_cuda(*tensors):
    """Same contract as the reference's CHECK_INPUT (include/utils.h:9-10): inputs must be CUDA tensors."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("splatter_a_video_b200: all tensors must be CUDA tensors (no CPU path)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"splatter_a_video_b200: tensors on different devices ({dev} and {t.device})")
    # the library launches on the CURRENT device and torch's current stream: refuse a mismatch loudly instead of launching
    # kernels of cuda:0 on pointers of cuda:1
    if dev is not None and dev.index is not None and dev.index != torch.cuda.current_device():
        raise RuntimeError(f"splatter_a_video_b200: tensors live on {dev} but the current device is cuda:{torch.cuda.current_device()}; "
                           f"wrap the call in `with torch.cuda.device({dev.index}):`")


def f32c(t):
    """fp32 + contiguous view/copy, like the `.contiguous().data_ptr<float>()` the reference does per call."""
    if t is None:
        return None
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()
