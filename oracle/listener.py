"""TEST INFRASTRUCTURE ONLY (checker, never the product path).

CPU fp32 restatement of the reference listener (VSEFCModel): image encoder, GRU caption encoder
over index or one-hot captions, cosine score matrix and the max-violation hinge loss.  Parameters
are a plain dict keyed by the reference's state-dict names.  Pinned against
the real reference by tests/golden/make_golden.py.
"""
from __future__ import annotations

from typing import Dict

import torch

Params = Dict[str, torch.Tensor]


def l2norm(X):
    """X / (||X||_2 + 1e-7), eps outside the sqrt                       (VSEFCModel.py:12-17)"""
    return X / (torch.norm(X, dim=1, keepdim=True) + 1e-7)


def _abs(x, replay, key):
    """torch.abs; with `replay` (test aid) the sign decisions come from replay[key + "_sign"] (+-1,
    taken from the implementation under test) and the oracle's own values are left in
    replay[key + "_val"] for the near-zero check."""
    if replay is None:
        return x.abs()
    replay[key + "_val"] = x.detach()
    sign = replay.get(key + "_sign")
    return x.abs() if sign is None else x * sign


def img_enc(P: Params, fc_feats, no_imgnorm=False, use_abs=False, replay=None):
    """EncoderImage.forward                                             (VSEFCModel.py:40-54)"""
    f = torch.nn.functional.linear(fc_feats, P["img_enc.fc.weight"], P["img_enc.fc.bias"])
    if not no_imgnorm:
        f = l2norm(f)
    if use_abs:
        f = _abs(f, replay, "im")
    return f


def gru_step(P: Params, x, h):
    """One step of nn.GRU (gate order r, z, n; torch semantics)        (VSEFCModel.py:74-76)"""
    H = h.size(1)
    gi = torch.nn.functional.linear(x, P["txt_enc.rnn.weight_ih_l0"], P["txt_enc.rnn.bias_ih_l0"])
    gh = torch.nn.functional.linear(h, P["txt_enc.rnn.weight_hh_l0"], P["txt_enc.rnn.bias_hh_l0"])
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1 - z) * n + z * h


def txt_enc(P: Params, seqs, masks, pool_type="last", use_abs=False, replay=None):
    """EncoderText.forward                                              (VSEFCModel.py:83-140)
    lengths = sum(mask > 0) (:84); embedding by lookup or dense one-hot matmul (:102-106);
    a packed GRU only advances rows with t < len (:108-112) -- restated as a masked update, which
    also removes the sort / unsort (:85-93,:134) since rows are independent; pooling :115-127."""
    lens = (masks > 0).long().sum(1)
    W = P["txt_enc.embed.weight"]
    emb = torch.matmul(seqs, W) if seqs.dim() > 2 else W[seqs]
    B, S = emb.shape[:2]
    H = P["txt_enc.rnn.weight_hh_l0"].size(1)
    h = emb.new_zeros(B, H)
    outs = []
    tmax = int(lens.max())
    for t in range(tmax):
        hn = gru_step(P, emb[:, t], h)
        act = (t < lens)[:, None]
        h = torch.where(act, hn, h)
        outs.append(torch.where(act, hn, torch.zeros_like(hn)))   # pad_packed zero-fills
    if pool_type == "mean":
        m = masks[:, :tmax].float()
        out = (torch.stack(outs, 1) * m[:, :, None]).sum(1) / masks.float().sum(1, keepdim=True)
    elif pool_type == "max":
        m = masks[:, :tmax].float()
        filled = torch.stack(outs, 1) * m[:, :, None] + (m == 0)[:, :, None].float() * -1e10
        if replay is not None:
            replay["pool_vals"] = filled.detach()                      # [B, tmax, H]
        if replay is not None and "pool_arg" in replay:                # replayed arg-max step
            out = filled.gather(1, replay["pool_arg"][:, None, :]).squeeze(1)
        else:
            out = filled.max(1)[0]
    else:
        out = h                                                        # gather at len-1 (:128)
    out = l2norm(out)
    if use_abs:
        out = _abs(out, replay, "cap")
    return out


def contrastive_loss(im, s, margin=0.2, max_violation=True, whole_batch=False,
                     only_one_retrieval="off", hinge_replay=None):
    """ContrastiveLoss.forward                                          (VSEFCModel.py:167-207)

    `hinge_replay` (test aid, like the token / maxout / ReLU replays of the speaker oracle): a dict
    with int64 `arg_s` [B] / `arg_im` [B] replaces the max-violation arg-max (:191-193) by the given
    hardest-negative indices, so a near-tie decided differently under bf16 operands does not
    masquerade as a gradient error; the oracle's own score matrix is left in `hinge_replay["scores"]`
    for the near-tie check."""
    scores = im @ s.t()                                                # cosine_sim, :143-146
    if hinge_replay is not None:
        hinge_replay["scores"] = scores.detach()
    diag = scores.diag().view(-1, 1)
    cost_s = (margin + scores - diag).clamp(min=0)                     # :176 caption retrieval
    cost_im = (margin + scores - diag.t()).clamp(min=0)                # :179 image retrieval
    eye = torch.eye(scores.size(0), dtype=torch.bool)
    cost_s = cost_s.masked_fill(eye, 0)                                # :182-188
    cost_im = cost_im.masked_fill(eye, 0)
    if max_violation and hinge_replay is not None and "arg_s" in hinge_replay:
        cost_s = cost_s.gather(1, hinge_replay["arg_s"].view(-1, 1)).squeeze(1)
        cost_im = cost_im.gather(0, hinge_replay["arg_im"].view(1, -1)).squeeze(0)
    elif max_violation:
        cost_s = cost_s.max(1)[0]                                      # :191-193
        cost_im = cost_im.max(0)[0]
    else:
        cost_s = cost_s.mean(1)
        cost_im = cost_im.mean(0)
    fn = (lambda x: x) if whole_batch else (lambda x: x.sum())         # :197-200
    if only_one_retrieval == "image":
        return fn(cost_im)
    if only_one_retrieval == "caption":
        return fn(cost_s)
    return fn(cost_s) + fn(cost_im)


def vse_forward(P: Params, fc_feats, seq, masks, whole_batch=False, only_one_retrieval="off",
                margin=0.2, max_violation=True, pool_type="last", hinge_replay=None, use_abs=False,
                no_imgnorm=False, enc_replay=None):
    """VSEFCModel.forward                                               (VSEFCModel.py:230-241)
    `enc_replay` (test aid): decisions of the encoders' non-smooth options -- sign of |.| and the
    max-pool arg-max step -- taken from the implementation under test (see _abs / txt_enc)."""
    return contrastive_loss(img_enc(P, fc_feats, no_imgnorm, use_abs, enc_replay),
                            txt_enc(P, seq, masks, pool_type, use_abs, enc_replay), margin,
                            max_violation, whole_batch, only_one_retrieval, hinge_replay)
