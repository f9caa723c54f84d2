"""TEST INFRASTRUCTURE (checker only -- never imported by the product): pure-PyTorch restatement, tensor by tensor, of the
reference's densification (src/pointrix/optimizer/atlas_gs_optimizer.py:93-379) with the optimizer-state surgery of
src/pointrix/point_cloud/points.py:281-365, on plain dicts of CPU tensors.  The split's random draw is passed in so the product
and this restatement consume the same samples."""
import torch


def build_rotation(r):                                       # src/pointrix/utils/gaussian_points/gaussian_utils.py:11-33
    q = r / torch.sqrt((r * r).sum(1))[:, None]
    R = torch.zeros(q.size(0), 3, 3)
    r_, x, y, z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R[:, 0, 0] = 1 - 2 * (y * y + z * z); R[:, 0, 1] = 2 * (x * y - r_ * z); R[:, 0, 2] = 2 * (x * z + r_ * y)
    R[:, 1, 0] = 2 * (x * y + r_ * z); R[:, 1, 1] = 1 - 2 * (x * x + z * z); R[:, 1, 2] = 2 * (y * z - r_ * x)
    R[:, 2, 0] = 2 * (x * z - r_ * y); R[:, 2, 1] = 2 * (y * z + r_ * x); R[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def update_stats(state, viewspace_grad, radii, visibility):   # :110-121
    state["max_radii"][visibility] = torch.max(state["max_radii"][visibility], radii[visibility].float())
    state["grad_accum"][visibility] += torch.norm(viewspace_grad[visibility, :2], dim=-1)
    state["denom"][visibility] += 1


def densification(attrs, moments, state, cfg, duplicate, prune, samples=None):
    """attrs: {name: [P,...]} incl. position, scaling (log), rotation, opacity (logit); moments: {name: (exp_avg, exp_avg_sq)}.
    Returns new (attrs, moments, state) in the reference's population order."""
    attrs = {k: v.clone() for k, v in attrs.items()}
    moments = {k: (a.clone(), b.clone()) for k, (a, b) in moments.items()}
    state = {k: v.clone() for k, v in state.items()}
    split_num = cfg["split_num"]

    def extend(new):
        for k in attrs:
            attrs[k] = torch.cat([attrs[k], new[k]], 0)
            if k in moments:
                moments[k] = tuple(torch.cat([m, torch.zeros_like(new[k])], 0) for m in moments[k])

    def remove(mask):
        for k in attrs:
            attrs[k] = attrs[k][mask]
            if k in moments:
                moments[k] = tuple(m[mask] for m in moments[k])

    def reset():
        n = attrs["position"].shape[0]
        state["grad_accum"], state["denom"], state["max_radii"] = torch.zeros(n), torch.zeros(n), torch.zeros(n)

    if duplicate:
        grads = state["grad_accum"] / state["denom"]
        grads[grads.isnan()] = 0.0
        scaling = attrs["scaling"].exp()
        # clone (:284-299)
        mask = (grads.abs() >= cfg["grad_threshold"]) & (scaling.max(1).values <= cfg["percent_dense"] * cfg["extent"])
        extend({k: v[mask] for k, v in attrs.items()})
        reset()
        # split (:301-331): grads padded with zeros for the clones
        n = attrs["position"].shape[0]
        padded = torch.zeros(n); padded[:grads.shape[0]] = grads
        scaling = attrs["scaling"].exp()
        mask = (padded >= cfg["grad_threshold"]) & (scaling.max(1).values > cfg["percent_dense"] * cfg["extent"])
        if samples is None:
            samples = torch.zeros(int(mask.sum()) * split_num, 3)
        rots = build_rotation(attrs["rotation"][mask]).repeat(split_num, 1, 1)
        new_pos = torch.bmm(rots, samples.unsqueeze(-1)).squeeze(-1) + attrs["position"][mask].repeat(split_num, 1)
        new_scaling = torch.log(scaling[mask].repeat(split_num, 1) / (0.8 * split_num))
        new = {}
        for k, v in attrs.items():
            sizes = [1] * v.dim(); sizes[0] = split_num
            new[k] = v[mask].repeat(*sizes)
        new["position"], new["scaling"] = new_pos, new_scaling
        extend(new)
        reset()
        valid = ~torch.cat([mask, torch.zeros(split_num * int(mask.sum()), dtype=torch.bool)])
        remove(valid)
        for k in state:
            state[k] = state[k][valid]
    if prune:                                                 # :333-362
        bad = torch.sigmoid(attrs["opacity"]).reshape(-1) < cfg["min_opacity"]
        if cfg["size_threshold"]:
            bad = bad | (state["max_radii"] > cfg["size_threshold"]) | (attrs["scaling"].exp().max(1).values > 0.1 * cfg["extent"])
        valid = ~bad
        remove(valid)
        for k in state:
            state[k] = state[k][valid]
    return attrs, moments, state


def reset_opacity(attrs, moments, cap=0.01):                  # :185-197 (+ replace_optimizer: moments cleared)
    opc = torch.sigmoid(attrs["opacity"])
    x = torch.min(opc, torch.ones_like(opc) * cap)
    attrs["opacity"] = torch.log(x / (1 - x))
    if "opacity" in moments:
        moments["opacity"] = tuple(torch.zeros_like(m) for m in moments["opacity"])
    return attrs, moments
