"""Densification on FlatParams: clone / split / prune / opacity reset with the reference's
semantics (src/pointrix/optimizer/atlas_gs_optimizer.py:93-379) but ONE gather launch per flat buffer instead of re-creating
every parameter and both Adam moments tensor by tensor (src/pointrix/point_cloud/points.py:281-365).

The population order after a densification step is the reference's: surviving originals (clone and split run back to back on
the same gradient statistics; split originals are removed), then the clones, then the split children (`split_num` copies,
copy-major like `repeat(split_num, 1)`), with the prune filter applied last.  Split positions use the SAME random draw as the
reference (`torch.normal(mean=0, std=scaling[mask].repeat(split_num, 1))` on the CUDA generator), so a run seeded like the
reference produces the same children.

There is no CPU fallback: the kernels live in csrc/densify.cu behind the C ABI.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _lib as L
from .parallel import FlatAdam, FlatParams


class FlatDensifier:
    """Statistics + structure updates for a FlatParams population.

    names: which FlatParams entries (or `extras`: per-point tensors outside the flat buffer, e.g. a frozen base position) play
    the roles "position", "scaling", "rotation", "opacity".  scaling is stored as log, opacity as logit (the reference's
    activations, src/pointrix/point_cloud/utils.py) unless `scaling_is_log` / `opacity_is_logit` say otherwise."""

    def __init__(self, flat: FlatParams, P: int, names: Dict[str, str], extras: Optional[Dict[str, Tensor]] = None,
                 percent_dense: float = 0.01, split_num: int = 2, densify_grad_threshold: float = 0.0002, min_opacity: float = 0.005,
                 cameras_extent: float = 1.0, size_threshold: float = 20.0, scaling_is_log: bool = True, opacity_is_logit: bool = True):
        self.flat, self.P, self.names = flat, int(P), dict(names)
        self.extras = dict(extras or {})
        self.percent_dense, self.split_num = float(percent_dense), int(split_num)
        self.grad_threshold, self.min_opacity = float(densify_grad_threshold), float(min_opacity)
        self.extent, self.size_threshold = float(cameras_extent), float(size_threshold)
        self.scaling_is_log, self.opacity_is_logit = bool(scaling_is_log), bool(opacity_is_logit)
        self._reset_stats()

    # ---- statistics (update_structure, :110-121) -----------------------------------------------------------------------
    def _reset_stats(self):
        dev = self.flat.flat.device
        self.grad_accum = torch.zeros(self.P, dtype=torch.float32, device=dev)
        self.denom = torch.zeros(self.P, dtype=torch.float32, device=dev)
        self.max_radii = torch.zeros(self.P, dtype=torch.float32, device=dev)

    def update_stats(self, ndc_grad: Tensor, radii: Tensor, visibility: Optional[Tensor] = None):
        """ndc_grad: viewspace_points[i].grad summed over the batch ([P,>=2]); radii int32 [P]; visibility bool [P] or None."""
        L.need_cuda(ndc_grad, radii)
        g = L.f32c(ndc_grad[:, :2])
        r = radii.to(torch.int32).contiguous()
        if visibility is not None and visibility.dtype == torch.bool:
            visibility = visibility.contiguous().view(torch.uint8)      # same bytes (0 / 1): no conversion kernel on the step
        v = None if visibility is None else visibility.to(torch.uint8).contiguous()
        L.call("spv_densify_stats", self.P, L.ptr(g), L.ptr(r), L.ptr(v), L.ptr(self.grad_accum), L.ptr(self.denom), L.ptr(self.max_radii),
               L.stream())

    def _tensor(self, role: str) -> Tensor:
        n = self.names[role]
        return self.extras[n] if n in self.extras else self.flat.params[n].detach()

    # ---- structure (densification + prune, :157-353) ----------------------------------------------------------------
    @torch.no_grad()
    def densify_and_prune(self, adam: Optional[FlatAdam], duplicate: bool, prune: bool,
                          generator: Optional[torch.Generator] = None) -> Tuple[FlatParams, Optional[FlatAdam], Dict[str, Tensor]]:
        """One call of `densification(step)`: clone + split when `duplicate`, the prune filter when `prune`.
        Returns (new FlatParams, new FlatAdam or None, new extras); `self` now tracks the new population."""
        dev, P = self.flat.flat.device, self.P
        if adam is not None:
            adam.flush()            # an interval-lazy optimizer brings every spline interval up to date before rows are copied
        flags = torch.zeros(P, dtype=torch.uint8, device=dev)
        scaling, opacity = L.f32c(self._tensor("scaling")), L.f32c(self._tensor("opacity").reshape(-1))
        L.call("spv_densify_flags", P, L.ptr(self.grad_accum), L.ptr(self.denom), L.ptr(scaling), L.ptr(opacity), L.ptr(self.max_radii),
               int(self.scaling_is_log), int(self.opacity_is_logit), self.grad_threshold, self.percent_dense * self.extent,
               self.min_opacity, self.size_threshold, 0.1 * self.extent, L.ptr(flags), L.stream())
        ar = torch.arange(P, dtype=torch.int32, device=dev)
        clone = (flags & 1).bool() if duplicate else torch.zeros(P, dtype=torch.bool, device=dev)
        split = (flags & 2).bool() if duplicate else torch.zeros(P, dtype=torch.bool, device=dev)
        kept, cl, sp = ar[~split], ar[clone], ar[split]
        n_split = int(sp.numel())
        children = sp.repeat(self.split_num)                                   # copy-major, like value.repeat(split_num, ...)
        src = torch.cat([kept, cl, children])
        is_new = torch.cat([torch.zeros_like(kept, dtype=torch.bool), torch.ones(cl.numel() + children.numel(), dtype=torch.bool, device=dev)])
        child_index = torch.full((src.numel(),), -1, dtype=torch.int32, device=dev)
        child_index[kept.numel() + cl.numel():] = torch.arange(children.numel(), dtype=torch.int32, device=dev)
        samples = None
        if n_split:
            act = scaling.exp() if self.scaling_is_log else scaling
            stds = act[sp.long()].repeat(self.split_num, 1)
            samples = torch.normal(mean=torch.zeros_like(stds), std=stds, generator=generator).contiguous()   # new_pos_scale, :273-275
        self.last_samples = samples
        if prune:
            # the filter sees the NEW population: copied opacities, the children's reduced scales, and -- because clone / split
            # reset the statistics (reset_densification_state) -- zero max radii whenever a duplication ran in this call
            srcl = src.long()
            act = scaling.exp() if self.scaling_is_log else scaling
            smax = act[srcl].max(dim=1).values
            smax = torch.where(child_index >= 0, smax / (0.8 * self.split_num), smax)
            op = opacity[srcl]
            op = torch.sigmoid(op) if self.opacity_is_logit else op
            bad = op < self.min_opacity
            if self.size_threshold > 0:
                mr = torch.zeros_like(smax) if duplicate else self.max_radii[srcl]
                bad = bad | (mr > self.size_threshold) | (smax > 0.1 * self.extent)
            keep = ~bad
            src, is_new, child_index = src[keep], is_new[keep], child_index[keep]
        P_new = int(src.numel())
        src = src.contiguous()
        new_flat, new_adam = self._regather(adam, src, is_new, P_new)
        new_extras = {k: v[src.long()].contiguous() for k, v in self.extras.items()}
        if n_split and P_new:
            pos_name, sc_name = self.names["position"], self.names["scaling"]
            new_pos = new_extras[pos_name] if pos_name in new_extras else new_flat.params[pos_name].detach()
            new_sc = new_extras[sc_name] if sc_name in new_extras else new_flat.params[sc_name].detach()
            L.call("spv_split_children", P_new, L.ptr(src), L.ptr(child_index.contiguous()), L.ptr(samples),
                   L.ptr(L.f32c(self._tensor("position"))), L.ptr(scaling), L.ptr(L.f32c(self._tensor("rotation"))), int(self.scaling_is_log),
                   0.8 * self.split_num, L.ptr(new_pos), L.ptr(new_sc), L.stream())
        # statistics: reset by any duplication, masked by a prune alone (prune_postprocess)
        old = (self.grad_accum, self.denom, self.max_radii)
        self.flat, self.P, self.extras = new_flat, P_new, new_extras
        self._reset_stats()
        if not duplicate:
            for dst, o in zip((self.grad_accum, self.denom, self.max_radii), old):
                dst.copy_(o[src.long()])
        return new_flat, new_adam, new_extras

    def _regather(self, adam: Optional[FlatAdam], src: Tensor, is_new: Tensor, P_new: int):
        old = self.flat
        widths = [n // self.P for n in old.sizes]
        new_flat = FlatParams.empty_like(old, self.P, P_new)
        nb = len(widths)
        w_arr = (ctypes.c_int * nb)(*widths)
        o_off = (ctypes.c_longlong * nb)(*[sum(old.sizes[:q]) for q in range(nb)])
        n_off = (ctypes.c_longlong * nb)(*[sum(new_flat.sizes[:q]) for q in range(nb)])

        def move(src_idx, old_buf, new_buf):
            L.call("spv_flat_regather", nb, ctypes.cast(w_arr, ctypes.c_void_p), ctypes.cast(o_off, ctypes.c_void_p),
                   ctypes.cast(n_off, ctypes.c_void_p), P_new, L.ptr(src_idx), L.ptr(old_buf), L.ptr(new_buf), L.stream())

        if P_new:
            move(src, old.flat, new_flat.flat)
        new_adam = None
        if adam is not None:
            lrs = {k: float(adam._lrs[q]) for q, k in enumerate(old.names)}
            lazy = dict(adam.lazy, P=P_new) if adam.lazy else None
            new_adam = FlatAdam(new_flat, lrs, betas=adam.betas, eps=adam.eps, device_clock=adam.device_clock, lazy=lazy)
            new_adam.t = adam.t
            if adam.device_clock:
                new_adam.state_dev.copy_(adam.state_dev)       # the optimizer clock survives the restructure
            if adam.lazy:
                new_adam.last_dev.copy_(adam.last_dev); new_adam.ring_dev.copy_(adam.ring_dev)
            if P_new:
                fresh = torch.where(is_new, torch.full_like(src, -1), src).contiguous()      # new points start with zero moments
                move(fresh, adam.exp_avg, new_adam.exp_avg)
                move(fresh, adam.exp_avg_sq, new_adam.exp_avg_sq)
        return new_flat, new_adam

    # ---- opacity reset (:185-197) -----------------------------------------------------------------------------------
    @torch.no_grad()
    def reset_opacity(self, adam: Optional[FlatAdam], cap: float = 0.01):
        name = self.names["opacity"]
        p = self.flat.params[name]
        off = sum(self.flat.sizes[:self.flat.names.index(name)])
        m = adam.exp_avg[off:off + self.P] if adam is not None else None
        v = adam.exp_avg_sq[off:off + self.P] if adam is not None else None
        L.call("spv_reset_opacity", self.P, float(cap), int(self.opacity_is_logit), L.ptr(p.detach()), L.ptr(m), L.ptr(v), L.stream())
