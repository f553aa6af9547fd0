"""Helpers shared by the GPU parity tests (CUDA path vs. the CPU oracle on the same seeded inputs)."""
import torch

from oracle import synth


def cuda_params(P):
    return {k: v.detach().clone().cuda().contiguous() for k, v in P.items()}


def pack_keep(keep_padded, att_masks):
    """[B, L, R] float keep-mask (oracle layout) -> uint8 [NL, R] over the valid regions."""
    if att_masks is None:
        return keep_padded.reshape(-1, keep_padded.shape[-1]).to(torch.uint8).cuda().contiguous()
    return keep_padded[att_masks > 0].to(torch.uint8).cuda().contiguous()


def u8(t):
    return None if t is None else t.to(torch.uint8).cuda().contiguous()


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)


REAL = synth.Dims()


def branch_replay(sp, att_masks, noise):
    """Copy of `noise` that makes the oracle replay the CUDA pass's non-smooth decisions (maxout
    branch, att_embed ReLU) and record its own margins -- the gradient analogue of forced tokens."""
    import copy
    R = sp.dims.R
    n = sp.n_steps
    s = sp.t["s_all"][:n].float().cpu()
    u = sp.t["u_all"][:n].float().cpu()
    first = (s[:, :, 3 * R:4 * R] + u[:, :, :R]) >= (s[:, :, 4 * R:5 * R] + u[:, :, R:])
    on = (sp.t["att_e16"].float() > 0).cpu()                 # packed [NL, R]
    B, L = sp.B, sp.L
    if att_masks is None:
        relu = on.view(B, L, R)
    else:
        relu = torch.zeros(B, L, R, dtype=torch.bool)
        relu[att_masks[:, :L] > 0] = on
    if noise.drop_att is not None:
        # a dropped unit hides the ReLU decision; it carries no gradient either way
        relu = relu | (noise.drop_att == 0)
    out = copy.copy(noise)
    out.maxout_first, out.relu_att, out.margins = first, relu, {}
    if "soft16" in sp.t:
        # partial-sampling modes: the next input relu(v . embed) is a bf16 contraction here
        E = sp.dims.E
        on = (sp.t["xh16"][: n, :, :E].float() > 0).cpu()           # [steps, B, E]
        if noise.drop_embed is not None:
            on = on | (noise.drop_embed[: n] == 0)
        out.relu_embed = on
    return out


def check_near_ties(replay_noise, att_masks, tol):
    """The replayed decisions may differ from the oracle's own only at near-ties (|margin| < tol
    relative to the typical margin)."""
    mg = replay_noise.margins
    stats = {}
    m = torch.stack(mg["maxout"], 0)[: replay_noise.maxout_first.size(0)]
    dec = replay_noise.maxout_first[: m.size(0)]
    flipped = dec != (m >= 0)
    scale = float(m.abs().median())
    stats["maxout_flips"] = int(flipped.sum())
    stats["maxout_total"] = flipped.numel()
    stats["maxout_worst"] = float(m[flipped].abs().max()) / scale if flipped.any() else 0.0
    if flipped.any():
        assert float(m[flipped].abs().max()) <= tol * scale, \
            f"maxout decision differs away from a tie ({stats['maxout_worst']:.3e} of the median margin)"
    pre = mg["relu_att"]
    dec = replay_noise.relu_att
    valid = torch.ones_like(dec) if att_masks is None else (att_masks[:, :dec.size(1)] > 0)[:, :, None].expand_as(dec)
    if replay_noise.drop_att is not None:
        valid = valid & (replay_noise.drop_att > 0)
    flipped = (dec != (pre > 0)) & valid
    scale = float(pre[valid].abs().median())      # padded regions hold the bias only: not a margin
    stats["relu_flips"] = int(flipped.sum())
    stats["relu_total"] = int(valid.sum())
    stats["relu_worst"] = float(pre[flipped].abs().max()) / scale if flipped.any() else 0.0
    if flipped.any():
        assert float(pre[flipped].abs().max()) <= tol * scale, \
            f"ReLU decision differs away from a tie ({stats['relu_worst']:.3e} of the median margin)"
    if replay_noise.relu_embed is not None and "relu_embed" in mg:
        pre = torch.stack(mg["relu_embed"], 0)                       # oracle steps 1..k
        pre = pre[: replay_noise.relu_embed.size(0) - 1]             # the last step's input is never built here
        dec = replay_noise.relu_embed[1: 1 + pre.size(0)]
        valid = torch.ones_like(dec)
        if replay_noise.drop_embed is not None:
            valid = replay_noise.drop_embed[1: 1 + pre.size(0)] > 0
        flipped = (dec != (pre > 0)) & valid
        # per-row scale: soft rows have much smaller pre-activations than one-hot rows
        scale = pre.abs().median(dim=2, keepdim=True)[0].expand_as(pre)
        stats["embed_relu_flips"] = int(flipped.sum())
        stats["embed_relu_total"] = int(valid.sum())
        if flipped.any():
            assert bool((pre[flipped].abs() <= tol * scale[flipped]).all()), \
                "embed ReLU decision differs away from a tie"
    return stats


def hinge_replay_of(lp):
    """The listener's max-violation arg-max decisions (VSEFCModel.py:191-193) of a CUDA pass, in the
    form oracle.listener.contrastive_loss replays."""
    return {"arg_s": lp.t["arg_s"].cpu().long(), "arg_im": lp.t["arg_im"].cpu().long()}


def check_hinge_near_ties(hinge_replay, tol=2e-3):
    """The replayed hardest negatives may differ from the oracle's own arg-max only where the
    oracle's score of the replayed negative is within `tol` of its own maximum."""
    S = hinge_replay["scores"]
    B = S.size(0)
    S = S.masked_fill(torch.eye(B, dtype=torch.bool), -1e9)
    flips, worst = 0, 0.0
    for dim, key in ((1, "arg_s"), (0, "arg_im")):
        best = S.max(dim)[0]
        idx = hinge_replay[key]
        got = S.gather(dim, idx.view(-1, 1) if dim == 1 else idx.view(1, -1)).reshape(-1)
        own = idx == torch.arange(B)          # the kernel reports i itself when B == 1
        gap = (best - got)[~own]
        flips += int((gap > 0).sum())
        worst = max(worst, float(gap.max()) if gap.numel() else 0.0)
        assert bool((gap <= tol).all()), (key, gap.max())
    return {"hinge_flips": flips, "hinge_total": 2 * B, "hinge_worst_gap": worst}


def grad_report(named_grads, ref, skip_zero=1e-11):
    """Per-tensor relative L2 error and cosine of the CUDA gradients against oracle gradients
    `ref` ({name: tensor}); `named_grads` is {name: tensor or None}."""
    out = {}
    for name, g in named_grads.items():
        r = ref[name].double().flatten()
        g = torch.zeros_like(r) if g is None else g.detach().double().cpu().flatten()
        if name.endswith("alpha_net.bias") or float(r.abs().max()) < skip_zero:
            continue
        out[name] = dict(l2=float((g - r).norm() / r.norm()),
                         cos=float((g @ r) / (g.norm() * r.norm() + 1e-300)),
                         ref_norm=float(r.norm()))
    return out


def write_report(name, payload):
    """Drop a JSON record under gpurun_out/ (merged back by gpurun; copied to profiles/ by hand)."""
    import json
    import os
    root = os.environ.get("GRAFT_REPO_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "gpurun_out", "parity")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, name + ".json"), "w") as f:
        json.dump(payload, f, indent=1, sort_keys=True)
