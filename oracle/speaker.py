"""TEST INFRASTRUCTURE ONLY (checker, never the product path).

CPU fp32 restatement of the reference speaker (Att2in2): prologue, attention, maxout-LSTM core,
vocabulary logits, the samplers, the `sample` decode loop and the teacher-forced XE `forward`.
Written functionally over a plain dict of parameters that uses the reference's state-dict names
(SURVEY.md Appendix B).  All randomness (Gumbel uniforms, multinomial exponentials, dropout
keep-masks, partial-sampling uniforms) is INJECTED so that the CUDA path and the real reference
can be driven with identical noise.

Pinned against the real reference (imported through oracle/ref_loader.py) by
tests/golden/make_golden.py; the resulting vectors live in tests/golden/*.npz and are re-checked
by tests/test_oracle_golden.py on every CPU run.

Each function cites the reference lines (relative to the reference root) it follows.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

GUMBEL_EPS = 1e-20  # models/gumbel.py:6


@dataclass
class SpeakerNoise:
    """Injected randomness for one speaker pass.

    drop_att   keep-mask {0,1} for att_embed's dropout, padded layout [B, L, R]    (AttModel.py:82-85)
    drop_embed keep-masks [steps, B, E] for `embed`'s dropout, one per decode step (AttModel.py:74-76)
    drop_core  keep-masks [steps, B, R] for Att2in2Core.dropout                    (AttModel.py:529)
    U          uniforms [T, B, V+1] turned into Gumbel noise                       (gumbel.py:6-11)
    E          Exp(1) draws [T, B, V+1]; torch.multinomial(p,1) == argmax(p / E)   (multinomial.py:17)
    part_u     uniforms [T, B] for the partial-sampling variants                   (gumbel_softmax.py:31)
    ss_u       uniforms [T, B] for scheduled sampling in AttModel.forward          (AttModel.py:119-120)
    Any entry may be None -> that source of randomness is off (dropout) or unused.
    """
    drop_att: Optional[torch.Tensor] = None
    drop_embed: Optional[torch.Tensor] = None
    drop_core: Optional[torch.Tensor] = None
    U: Optional[torch.Tensor] = None
    E: Optional[torch.Tensor] = None
    part_u: Optional[torch.Tensor] = None
    ss_u: Optional[torch.Tensor] = None          # uniforms [T, B] of scheduled sampling (AttModel.py:119-120)
    # --- branch replay (test aid, like forced_tokens): decisions of the non-smooth ops taken from
    # the implementation under test, so that a near-tie flip (bf16 vs fp32 pre-activations) does
    # not masquerade as a gradient error.  None -> the oracle decides itself (reference behaviour).
    relu_att: Optional[torch.Tensor] = None      # bool [B, L, R]: att_embed pre-activation > 0
    maxout_first: Optional[torch.Tensor] = None  # bool [steps, B, R]: u[:, :R] >= u[:, R:]
    relu_embed: Optional[torch.Tensor] = None    # bool [steps, B, E]: (v . embed) > 0 (partial sampling)
    # filled by the oracle for diagnostics: margins of its own decisions
    margins: Optional[dict] = None


def _dropout(x, keep, p):
    """nn.Dropout in training mode with an injected keep-mask (scale 1/(1-p))."""
    if keep is None or p <= 0.0:
        return x
    return x * keep.to(x.dtype) * (1.0 / (1.0 - p))


def _linear(x, P, name):
    return F.linear(x, P[name + ".weight"], P[name + ".bias"])


# --------------------------------------------------------------------------------------------
# prologue: att_embed through pack_wrapper, ctx2att                      (AttModel.py:31-51,110-114)
# --------------------------------------------------------------------------------------------
def prologue(P: Params, att_feats, att_masks, noise: SpeakerNoise, drop_p: float):
    """att_e = dropout(relu(att W^T + b)) on valid regions, exactly 0 on padded regions, width
    clipped to the longest row (pack_wrapper, AttModel.py:44-51); p_att = ctx2att(att_e)
    (AttModel.py:114) -- so padded positions of p_att hold ctx2att's bias."""
    pre = _linear(att_feats, P, "att_embed.0")
    if noise.margins is not None:
        noise.margins["relu_att"] = pre.detach()
    if noise.relu_att is not None:
        x = pre * noise.relu_att.to(pre.dtype)           # replayed ReLU decision
    else:
        x = torch.relu(pre)
    x = _dropout(x, noise.drop_att, drop_p)
    if att_masks is not None:
        lens = att_masks.long().sum(1)                    # AttModel.py:47
        lmax = int(lens.max())
        valid = (torch.arange(att_feats.size(1))[None, :] < lens[:, None]).to(x.dtype)
        x = (x * valid[:, :, None])[:, :lmax]             # pad_packed_sequence zero-fills
    p_att = _linear(x, P, "ctx2att")
    return x, p_att


# --------------------------------------------------------------------------------------------
# Attention.forward                                                        (AttModel.py:465-489)
# --------------------------------------------------------------------------------------------
def attention(P: Params, h, att_e, p_att, att_masks):
    att_h = _linear(h, P, "core.attention.h2att")                       # :470
    dot = torch.tanh(p_att + att_h[:, None, :])                         # :473-474
    e = F.linear(dot, P["core.attention.alpha_net.weight"],
                 P["core.attention.alpha_net.bias"]).squeeze(-1)        # :477-478
    w = torch.softmax(e, dim=1)                                         # :480 (padded width)
    if att_masks is not None:
        m = att_masks[:, : att_e.size(1)].to(w.dtype)                   # :482
        w = w * m
        w = w / w.sum(1, keepdim=True)                                  # :483
    return torch.bmm(w[:, None, :], att_e).squeeze(1), w                # :487


# --------------------------------------------------------------------------------------------
# Att2in2Core.forward                                                      (AttModel.py:510-531)
# --------------------------------------------------------------------------------------------
def core_step(P: Params, xt, h, c, att_e, p_att, att_masks, keep_core, drop_p, maxout_first=None,
              margins=None):
    R = h.size(1)
    att_res, w = attention(P, h, att_e, p_att, att_masks)               # :511
    s = _linear(xt, P, "core.i2h") + _linear(h, P, "core.h2h")          # :514
    sig = torch.sigmoid(s[:, : 3 * R])                                  # :515-516
    i, f, o = sig[:, :R], sig[:, R:2 * R], sig[:, 2 * R:3 * R]          # :517-519
    u = s[:, 3 * R:] + _linear(att_res, P, "core.a2c")                  # :521-522
    if margins is not None:
        margins.setdefault("maxout", []).append((u[:, :R] - u[:, R:]).detach())
    if maxout_first is not None:
        g = torch.where(maxout_first, u[:, :R], u[:, R:])               # replayed maxout decision
    else:
        g = torch.max(u[:, :R], u[:, R:])                               # :523-525 (maxout)
    c2 = f * c + i * g                                                  # :526
    h2 = o * torch.tanh(c2)                                             # :527
    out = _dropout(h2, keep_core, drop_p)                               # :529
    return out, h2, c2, w


def embed_tokens(P: Params, it, keep, drop_p):
    """self.embed = Embedding -> ReLU -> Dropout                          (AttModel.py:74-76)"""
    return _dropout(torch.relu(P["embed.0.weight"][it]), keep, drop_p)


def logits_of(P: Params, out):
    return _linear(out, P, "logit")                                     # AttModel.py:87,140,444


# --------------------------------------------------------------------------------------------
# samplers
# --------------------------------------------------------------------------------------------
def gumbel_from_uniform(U):
    """sample_gumbel: -log(-log(U + eps) + eps)                           (gumbel.py:6-11)"""
    return -torch.log(-torch.log(U + GUMBEL_EPS) + GUMBEL_EPS)


def _hard(y, ind):
    y_hard = torch.zeros_like(y)
    y_hard.scatter_(1, ind.view(-1, 1), 1.0)
    return y_hard


def st_gumbel(logprobs, tau, U):
    """gumbel_softmax: y = softmax((lp+G)/tau); ind = argmax y; STE one-hot  (gumbel.py:13-30)"""
    y = torch.softmax((logprobs + gumbel_from_uniform(U)) / tau, dim=-1)
    ind = y.max(dim=-1)[1]
    return (_hard(y, ind) - y).detach() + y, ind, y


def st_multinomial(logprobs, tau, E):
    """multinomial: y = softmax(lp/tau); ind ~ Multinomial(y); STE one-hot   (multinomial.py:4-27)
    torch.multinomial(y, 1) draws q ~ Exp(1) elementwise and returns argmax(y / q)."""
    y = torch.softmax(logprobs if tau == 1 else logprobs / tau, dim=1)
    ind = (y / E).max(dim=1)[1]
    return (_hard(y, ind) - y).detach() + y, ind, y


def _ps_out(y, ind, part_u, prob):
    if prob > 0.0:
        sel = (part_u < prob).to(y.dtype)[:, None]
        return (sel * _hard(y, ind) - sel * y).detach() + y
    return y


def ps_gumbel(logprobs, tau, U, part_u, prob):
    """gumbel_soft: rows with part_u < prob output the hard one-hot value, the others stay soft;
    all rows back-propagate through y                              (gumbel_softmax.py:17-42)"""
    y = torch.softmax((logprobs + gumbel_from_uniform(U)) / tau, dim=-1)
    ind = y.max(dim=-1)[1]
    return _ps_out(y, ind, part_u, prob), ind, y


def ps_multinomial(logprobs, tau, E, part_u, prob):
    """multinomial_soft: y = exp(lp/tau) (unnormalised when tau != 1)  (multinomial_soft.py:5-35)"""
    y = torch.exp(logprobs if tau == 1 else logprobs / tau)
    ind = (y / E).max(dim=1)[1]
    return _ps_out(y, ind, part_u, prob), ind, y


# --------------------------------------------------------------------------------------------
# AttModel.sample                                                          (AttModel.py:291-452)
# --------------------------------------------------------------------------------------------
@dataclass
class SampleResult:
    seq: torch.Tensor                      # [B, n] int64 (word_index)
    logprobs: torch.Tensor                 # [B, n] sampleLogprobs
    one_hots: Optional[torch.Tensor]       # [B, n, V+2] (ST / PS modes with use_one_hot)
    n_steps: int                           # number of core evaluations executed
    step_logprobs: List[torch.Tensor] = field(default_factory=list)   # log_softmax(logit) per step
    tokens_raw: List[torch.Tensor] = field(default_factory=list)      # sampled ids before masking
    perturbed: List[torch.Tensor] = field(default_factory=list)       # the score whose argmax is the sample


def sample(P: Params, att_feats, att_masks, *, mode: str, seq_length: int, vocab_size: int,
           noise: SpeakerNoise, drop_p: float = 0.0, sample_max: int = 1, use_one_hot: int = 0,
           temperature: float = 1.0, gumbel_temp: float = 1.0, multinomial_temp: float = 1.0,
           prob_gumbel_softmax: float = 1.0, prob_multinomial_soft: float = 1.0,
           forced_tokens: Optional[torch.Tensor] = None, keep_all_steps: bool = False,
           decoding_constraint: int = 0) -> SampleResult:
    """mode = retrieval_reward in {'reinforce','gumbel','multinomial','gumbel_softmax',
    'multinomial_soft'}.  `forced_tokens` [B, T] (optional, test aid) replaces the sampled id at
    each step so that a near-tie flip in one implementation cannot cascade; the id the oracle
    would have drawn is still recorded in `tokens_raw`.  `keep_all_steps` disables the early
    `break` (the CUDA path always runs all steps and slices afterwards).  `decoding_constraint`
    (index outputs only): the previously emitted id gets logit -inf (:437-442)."""
    B = att_feats.size(0)
    V1 = vocab_size + 1
    att_e, p_att = prologue(P, att_feats, att_masks, noise, drop_p)           # :315-319
    h = att_feats.new_zeros(B, P["core.h2h.weight"].size(1))
    c = torch.zeros_like(h)
    eos_one_hot = torch.zeros(1, V1 + 1)
    eos_one_hot[0, 0] = 1.0                                                   # :297-304
    idx_mode = bool(sample_max) or (mode == "reinforce") or (not use_one_hot)
    ps_mode = (not idx_mode) and mode in ("gumbel_softmax", "multinomial_soft")
    word_index, seq, seq_lp = [], [], []
    res = SampleResult(None, None, None, 0)
    logprobs = None
    unfinished = None
    for t in range(seq_length + 1):                                           # :323
        keep_e = None if noise.drop_embed is None else noise.drop_embed[t]
        keep_c = None if noise.drop_core is None else noise.drop_core[t]
        vec = None   # one_hot (ST modes) or soft_vec (PS modes), width V+1
        if t == 0:
            it = torch.full((B,), vocab_size + 1, dtype=torch.long)           # :324-326
        else:
            y = None
            if sample_max:
                score = logprobs
                it = torch.max(logprobs, 1)[1]                                # :327-329
            elif idx_mode:
                prob_prev = torch.exp(logprobs if temperature == 1.0 else logprobs / temperature)
                score = prob_prev / noise.E[t - 1]
                it = score.max(dim=1)[1]                                      # :332-343
            elif mode == "gumbel":
                vec, it, y = st_gumbel(logprobs, gumbel_temp, noise.U[t - 1])  # :345-354
                score = y
            elif mode == "multinomial":
                vec, it, y = st_multinomial(logprobs, multinomial_temp, noise.E[t - 1])  # :356-365
                score = y / noise.E[t - 1]
            elif mode == "gumbel_softmax":
                vec, it, y = ps_gumbel(logprobs, gumbel_temp, noise.U[t - 1],
                                       noise.part_u[t - 1], prob_gumbel_softmax)   # :367-378
                score = y
            elif mode == "multinomial_soft":
                vec, it, y = ps_multinomial(logprobs, multinomial_temp, noise.E[t - 1],
                                            noise.part_u[t - 1], prob_multinomial_soft)  # :381-392
                score = y / noise.E[t - 1]
            else:
                raise ValueError(mode)
            res.tokens_raw.append(it.clone())
            res.perturbed.append(score.detach())
            if forced_tokens is not None:
                it = forced_tokens[:, t - 1].clone()
                if vec is not None and not ps_mode:
                    vec = (_hard(y, it) - y).detach() + y
                elif vec is not None:
                    prob = prob_gumbel_softmax if mode == "gumbel_softmax" else prob_multinomial_soft
                    vec = _ps_out(y, it, noise.part_u[t - 1], prob)
            sample_lp = logprobs.gather(1, it[:, None]).squeeze(1)   # :328,:341,:347,:358,:370,:384
            if vec is not None:
                vec = torch.cat([vec, vec.new_zeros(B, 1)], 1)        # :348-354,:373-378
        # next input                                                          # :395-399
        if ps_mode and t >= 1:
            pre = vec @ P["embed.0.weight"]
            if noise.margins is not None:
                noise.margins.setdefault("relu_embed", []).append(pre.detach())
            if noise.relu_embed is not None and t < noise.relu_embed.size(0):
                xt = _dropout(pre * noise.relu_embed[t].to(pre.dtype), keep_e, drop_p)   # replayed
            else:
                xt = _dropout(torch.relu(pre), keep_e, drop_p)
        else:
            xt = embed_tokens(P, it, keep_e, drop_p)
        if t >= 1:
            unfinished = (it > 0) if t == 1 else unfinished * (it > 0)        # :403-406
            if not keep_all_steps and unfinished.sum() == 0:                  # :407-408
                break
            it = it * unfinished.type_as(it)                                  # :409
            if idx_mode:
                seq.append(it)                                                # :411-413
            else:
                word_index.append(it)                                         # :415,:427
                vec = vec * unfinished[:, None].to(vec.dtype)                 # :416-417,:428-429
                if bool((unfinished == 0).any()):
                    vec = torch.where(unfinished[:, None], vec, eos_one_hot)  # :419-420,:431-432
                seq.append(vec)
            seq_lp.append(sample_lp)                                          # :413,:423,:434
        mf = None if noise.maxout_first is None or t >= noise.maxout_first.size(0) \
            else noise.maxout_first[t]
        out, h, c, _ = core_step(P, xt, h, c, att_e, p_att, att_masks, keep_c, drop_p, mf,
                                 noise.margins)                                         # :436
        z = logits_of(P, out)
        if decoding_constraint and len(seq) > 0:                              # :437-442
            z = z + torch.zeros_like(z).scatter_(1, seq[-1][:, None], float("-inf"))
        logprobs = F.log_softmax(z, dim=1)                                    # :444
        res.step_logprobs.append(logprobs)
        res.n_steps += 1
    if len(seq) == 0:
        # the reference crashes on torch.cat([]) here; return width 0
        res.seq = torch.zeros(B, 0, dtype=torch.long)
        res.logprobs = torch.zeros(B, 0)
        return res
    if idx_mode:
        res.seq = torch.stack(seq, 1)                                         # :445-447
        res.logprobs = torch.stack(seq_lp, 1)
    else:
        res.seq = torch.stack(word_index, 1)                                  # :448-452
        res.one_hots = torch.stack(seq, 1)
        res.logprobs = torch.stack(seq_lp, 1)
    return res


# --------------------------------------------------------------------------------------------
# AttModel.forward (teacher-forced XE)                 (AttModel.py:103-148, misc/utils.py:49-58)
# --------------------------------------------------------------------------------------------
def language_model_criterion(logp, target, mask):
    """LanguageModelCriterion.forward                                     (misc/utils.py:49-58)"""
    target = target[:, : logp.size(1)]
    mask = mask[:, : logp.size(1)]
    out = -logp.gather(2, target[:, :, None]).squeeze(2) * mask
    return out.sum() / mask.sum()


def forward_xe(P: Params, att_feats, att_masks, seq, masks, *, noise: SpeakerNoise,
               drop_p: float = 0.0, return_logprobs: bool = False, ss_prob: float = 0.0,
               forced_fed: Optional[torch.Tensor] = None, fed_out: Optional[list] = None):
    """AttModel.forward: step i feeds seq[:, i] -- or, with scheduled sampling (ss_prob > 0, training,
    i >= 1; :118-131), on rows with ss_u[i-1] < ss_prob the id drawn from exp(outputs[i-1]) by
    torch.multinomial (== argmax(p / E[i-1]), no gradient through it); stops at the first i >= 1
    whose whole ground-truth column is 0 (:133); loss over seq[:,1:], masks[:,1:] (:144).
    `forced_fed` [B, steps] (test aid) replaces the drawn ids; `fed_out` collects the fed ids."""
    B = att_feats.size(0)
    att_e, p_att = prologue(P, att_feats, att_masks, noise, drop_p)
    h = att_feats.new_zeros(B, P["core.h2h.weight"].size(1))
    c = torch.zeros_like(h)
    outputs = []
    for i in range(seq.size(1) - 1):                                          # :116
        it = seq[:, i]
        if i >= 1 and ss_prob > 0.0:                                          # :118-129
            sel = noise.ss_u[i - 1] < ss_prob
            if forced_fed is not None:      # (a draw made right before the :133 break is never fed)
                drawn = forced_fed[:, i] if i < forced_fed.size(1) else seq[:, i]
            else:
                drawn = (torch.exp(outputs[-1].detach()) / noise.E[i - 1]).max(dim=1)[1]
            it = torch.where(sel, drawn, seq[:, i])
        if i >= 1 and int(seq[:, i].sum()) == 0:                              # :133
            break
        if fed_out is not None:
            fed_out.append(it.clone())
        keep_e = None if noise.drop_embed is None else noise.drop_embed[i]
        keep_c = None if noise.drop_core is None else noise.drop_core[i]
        xt = embed_tokens(P, it, keep_e, drop_p)                              # :136
        mf = None if noise.maxout_first is None or i >= noise.maxout_first.size(0) \
            else noise.maxout_first[i]
        out, h, c, _ = core_step(P, xt, h, c, att_e, p_att, att_masks, keep_c, drop_p, mf,
                                 noise.margins)                                         # :138
        outputs.append(F.log_softmax(logits_of(P, out), dim=1))               # :140
    logp = torch.stack(outputs, 1)                                            # :143
    loss = language_model_criterion(logp, seq[:, 1:], masks[:, 1:])           # :144
    return (loss, logp) if return_logprobs else loss


# --------------------------------------------------------------------------------------------
# AttModel.sample_beam                                                   (AttModel.py:150-289)
# --------------------------------------------------------------------------------------------
@dataclass
class BeamResult:
    seq: torch.Tensor                      # [B, T] int64: best finished beam per image
    logprobs: torch.Tensor                 # [B, T] its per-token log-probabilities
    done_beams: List[list]                 # per image: [{'seq','logps','p'}] sorted by -p (stable)
    parents: Optional[torch.Tensor] = None   # int32 [T, B, beam]: slot each new slot was forked from
    toks: Optional[torch.Tensor] = None      # int64 [T, B, beam]: word appended to it
    gaps: Optional[torch.Tensor] = None      # fp32 [T, B]: smallest score gap that decided a merge step


def sample_beam(P: Params, att_feats, att_masks, *, seq_length: int, vocab_size: int, beam_size: int,
                decoding_constraint: int = 0) -> BeamResult:
    """Beam search exactly as the reference runs it, one image at a time (eval mode: no dropout).

    Quirks kept (they change the result): the first merge only looks at beam 0 (:208-210);
    candidates are the top `beam_size` words of every beam, listed word-rank-major and then sorted
    by total log-probability with a STABLE sort (:211-221); a beam that emitted the end token is
    recorded in done_beams but is NOT retired -- it keeps its place, is fed embed(0) and goes on
    accumulating log-probability (:247-255 has no suppression); at t = seq_length every beam is
    recorded (:247); the answer is the first of done_beams sorted by -p (:283-287) -- where the
    recorded 'p' is `beam_logprobs_sum[vix]`, a 0-dim VIEW of the running-sum tensor (:250-253, not
    a copy), so every entry recorded in beam slot vix ends up carrying that slot's FINAL sum: the
    ranking is by the slot's final score, ties (same slot) by recording order.  `p_recorded` keeps
    the score at recording time (what the code evidently meant) for callers that want it."""
    B = att_feats.size(0)
    T = seq_length
    noise = SpeakerNoise()
    att_e_all, p_att_all = prologue(P, att_feats, att_masks, noise, 0.0)      # :157-162
    seq_out = torch.zeros(B, T, dtype=torch.long)
    lp_out = torch.zeros(B, T)
    all_done = []
    parents = torch.zeros(T, B, beam_size, dtype=torch.int32)
    toks = torch.zeros(T, B, beam_size, dtype=torch.long)
    gaps = torch.full((T, B), float("inf"))
    for k in range(B):
        bs = beam_size
        att_e = att_e_all[k:k + 1].expand(bs, *att_e_all.shape[1:]).contiguous()
        p_att = p_att_all[k:k + 1].expand(bs, *p_att_all.shape[1:]).contiguous()
        am = None if att_masks is None else att_masks[k:k + 1].expand(bs, att_masks.size(1)).contiguous()
        h = att_feats.new_zeros(bs, P["core.h2h.weight"].size(1))
        c = torch.zeros_like(h)
        beam_seq = torch.zeros(T, bs, dtype=torch.long)
        beam_lp = torch.zeros(T, bs)
        beam_sum = torch.zeros(bs)
        done = []
        logprobs = None
        for t in range(T + 1):
            if t == 0:
                it = torch.full((bs,), vocab_size + 1, dtype=torch.long)      # :192-195
            else:
                lpf = logprobs.float()
                if decoding_constraint and t > 1:                             # :204-207
                    lpf = lpf + torch.zeros_like(lpf).scatter_(1, beam_seq[t - 2:t - 1].t(), float("-inf"))
                ys, ix = torch.sort(lpf, 1, True)                             # :210
                cols, rows = min(bs, ys.size(1)), (1 if t == 1 else bs)
                cands = []
                for cc in range(cols):
                    for q in range(rows):
                        r = ys[q, cc]
                        cands.append(dict(c=int(ix[q, cc]), q=q, p=float(beam_sum[q] + r), r=float(r)))
                cands = sorted(cands, key=lambda x: -x["p"])                  # :221 (stable)
                # what a different rounding could flip: the order of the kept candidates and the
                # cut after them, and -- for a kept candidate at the last word rank -- the next word
                g = [cands[i]["p"] - cands[i + 1]["p"] for i in range(min(bs, len(cands) - 1))]
                g += [float(ys[v["q"], cols - 1] - ys[v["q"], cols]) for v in cands[:bs]
                      if cols < ys.size(1) and v["r"] == float(ys[v["q"], cols - 1])]
                gaps[t - 1, k] = min([x for x in g if x == x] + [float("inf")])
                nh, nc = h.clone(), c.clone()
                prev_seq, prev_lp = beam_seq[:t - 1].clone(), beam_lp[:t - 1].clone()
                for vix in range(bs):
                    v = cands[vix]
                    if t > 1:
                        beam_seq[:t - 1, vix] = prev_seq[:, v["q"]]
                        beam_lp[:t - 1, vix] = prev_lp[:, v["q"]]
                    nh[vix], nc[vix] = h[v["q"]], c[v["q"]]
                    beam_seq[t - 1, vix] = v["c"]
                    beam_lp[t - 1, vix] = v["r"]
                    beam_sum[vix] = v["p"]
                    parents[t - 1, k, vix], toks[t - 1, k, vix] = v["q"], v["c"]
                    if v["c"] == 0 or t == T:                                 # :247
                        done.append(dict(seq=beam_seq[:, vix].clone(), logps=beam_lp[:, vix].clone(),
                                         slot=vix, p_recorded=float(beam_sum[vix])))
                it = beam_seq[t - 1].clone()
                h, c = nh, nc
            xt = embed_tokens(P, it, None, 0.0)
            out, h, c, _ = core_step(P, xt, h, c, att_e, p_att, am, None, 0.0)     # :275-276
            logprobs = F.log_softmax(logits_of(P, out), dim=1)                # :277
        for e in done:
            e["p"] = float(beam_sum[e["slot"]])                               # the aliased view
        done = sorted(done, key=lambda x: -x["p"])                            # :283-284
        seq_out[k], lp_out[k] = done[0]["seq"], done[0]["logps"]
        all_done.append(done)
    return BeamResult(seq_out, lp_out, all_done, parents, toks, gaps)
