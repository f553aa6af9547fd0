"""Drop-in speaker: Att2in2Model with the reference's interface (models/AttModel.py).

Same constructor (`opt` namespace), same parameter names / shapes (state-dict compatible,
SURVEY.md Appendix B), same `forward(fc_feats, att_feats, att_masks, seq, masks)` and
`sample(fc_feats, att_feats, att_masks, opt={})` signatures and return conventions
(AttModel.py:103-148, :291-452).  The nn sub-modules only HOLD the parameters; all arithmetic runs
in libcoopcap (hand-written sm_100a kernels) through engine.py.  There is no CPU path: calling the
model with CPU tensors raises.

Differences kept deliberately: the dead 17th core evaluation is skipped;
the per-step host synchronisations are gone (one sync at the end of `sample` to learn the output
width n); `_loss['xe']` holds a 0-dim tensor instead of a Python float (no sync).
"""
from __future__ import annotations

import itertools
from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import engine as EN

_call_counter = itertools.count(1)


def _rank() -> int:
    import torch.distributed as dist
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def _next_seed() -> int:
    """Philox seed for one pass: torch's seed (so torch.manual_seed controls it), a call counter
    and the data-parallel rank.  Every rank calls torch.manual_seed(s) with the same s to start
    from identical weights; without the rank in the key all of them would then draw the same
    dropout masks and Gumbel noise for their (different) rows, and the averaged gradient would not
    be the mean of N independent single-process steps (per-rank RNG streams)."""
    s = torch.initial_seed() * 0x9E3779B97F4A7C15 + next(_call_counter)
    s ^= (_rank() + 1) * 0xD1B54A32D192ED03
    return s & (2 ** 63 - 1)


class Attention(nn.Module):
    """Parameter holder for AttModel.py:456-463 (h2att, alpha_net)."""

    def __init__(self, opt):
        super().__init__()
        self.rnn_size = opt.rnn_size
        self.att_hid_size = opt.att_hid_size
        self.h2att = nn.Linear(self.rnn_size, self.att_hid_size)
        self.alpha_net = nn.Linear(self.att_hid_size, 1)


class Att2in2Core(nn.Module):
    """Parameter holder for AttModel.py:492-508 (a2c, i2h, h2h, attention)."""

    def __init__(self, opt):
        super().__init__()
        self.input_encoding_size = opt.input_encoding_size
        self.rnn_size = opt.rnn_size
        self.drop_prob_lm = opt.drop_prob_lm
        self.att_hid_size = opt.att_hid_size
        self.a2c = nn.Linear(self.rnn_size, 2 * self.rnn_size)
        self.i2h = nn.Linear(self.input_encoding_size, 5 * self.rnn_size)
        self.h2h = nn.Linear(self.rnn_size, 5 * self.rnn_size)
        self.dropout = nn.Dropout(self.drop_prob_lm)
        self.attention = Attention(opt)


def _ordered(P: Dict[str, torch.Tensor]):
    return [P[n] for n in EN.SPEAKER_PARAM_NAMES]


class _SpeakerLossFn(torch.autograd.Function):
    """logp [n_steps, B] of given ids as a differentiable function of the speaker parameters.

    forward has already been run (SpeakerPass `sp`); this node only routes d(logp) back:
    dz = coef * (onehot - softmax(z)) -> BPTT (csrc/speaker_bwd.cu).  Used by the XE loss
    (AttModel.py:140-144) and by REINFORCE's sampleLogprobs (:341)."""

    @staticmethod
    def forward(ctx, sp, tok, owner, *params):
        ctx.sp, ctx.tok, ctx.owner = EN.retain(sp), tok, owner
        return sp.t["logp"][: sp.n_steps].clone()

    @staticmethod
    def backward(ctx, d_logp):
        sp = ctx.sp
        if sp is None or not sp.t:
            raise RuntimeError("the speaker pass was already released (backward a second time?)")
        P = ctx.owner._params()
        dz16 = EN.logp_logit_grads(sp, ctx.tok, d_logp.contiguous().float())
        G = EN.speaker_backward(sp, dz16, P)
        EN.release(sp)
        ctx.sp = None
        return (None, None, None) + tuple(G[n].view_as(P[n]) for n in EN.SPEAKER_PARAM_NAMES)


class _SampleSTFn(torch.autograd.Function):
    """Dense straight-through one-hots [B, n, V+2] (gumbel.py:27-30 / multinomial.py:24-27 +
    AttModel.py:348-354,416-422) whose backward is dz = y (g - <y,g>) / tau -> BPTT."""

    @staticmethod
    def forward(ctx, sp, n, owner, *params):
        if not isinstance(ctx, _Dummy):
            EN.retain(sp)
        ctx.sp, ctx.n, ctx.owner = sp, n, owner
        B, V2 = sp.B, sp.dims.V1 + 1
        tok = sp.t["tok_out"][:n].t()                      # [B, n]
        one_hots = torch.zeros(B, n, V2, device=tok.device)
        one_hots.scatter_(2, tok.unsqueeze(2), 1.0)
        return one_hots

    @staticmethod
    def backward(ctx, g):
        sp, n = ctx.sp, ctx.n
        P = ctx.owner._params()
        gt = torch.zeros(sp.n_steps, sp.B, g.shape[2], device=g.device)
        gt[:n] = g.transpose(0, 1)
        dz16 = EN.st_logit_grads_dense(sp, gt)
        G = EN.speaker_backward(sp, dz16, P)
        EN.release(sp)
        ctx.sp = None
        return (None, None, None) + tuple(G[n_].view_as(P[n_]) for n_ in EN.SPEAKER_PARAM_NAMES)


class _SamplePSFn(torch.autograd.Function):
    """Dense partial-sampling vectors [B, n, V+2] (gumbel_softmax.py:28-40 / multinomial_soft.py:21-33
    + AttModel.py:373-378,425-434): one-hot values on the selected rows, the relaxed sample y on
    the others, EOS one-hot on finished rows.  Backward: the upstream gradient plus the gradient
    that reaches v_t through the next input relu(v_t . embed) -> softmax backward -> BPTT."""

    @staticmethod
    def forward(ctx, sp, n, owner, *params):
        if not isinstance(ctx, _Dummy):
            EN.retain(sp)
        ctx.sp, ctx.n, ctx.owner = sp, n, owner
        B, V1 = sp.B, sp.dims.V1
        out = torch.zeros(B, n, V1 + 1, device=sp.t["soft16"].device)
        out[:, :, :V1] = sp.t["soft16"][:n].transpose(0, 1)
        return out

    @staticmethod
    def backward(ctx, g):
        sp, n = ctx.sp, ctx.n
        P = ctx.owner._params()
        V1 = sp.dims.V1
        gt = torch.zeros(sp.n_steps, sp.B, (V1 + 3) // 4 * 4, device=g.device)
        gt[:n, :, :V1] = g[:, :, :V1].transpose(0, 1)
        G = EN.speaker_backward(sp, None, P, ps_g_dense=gt)
        EN.release(sp)
        ctx.sp = None
        return (None, None, None) + tuple(G[n_].view_as(P[n_]) for n_ in EN.SPEAKER_PARAM_NAMES)


class AttModel(nn.Module):
    def __init__(self, opt):
        super().__init__()
        self.vocab_size = opt.vocab_size
        self.input_encoding_size = opt.input_encoding_size
        self.rnn_size = opt.rnn_size
        self.num_layers = opt.num_layers
        self.drop_prob_lm = opt.drop_prob_lm
        self.seq_length = opt.seq_length
        self.fc_feat_size = opt.fc_feat_size
        self.att_feat_size = opt.att_feat_size
        self.att_hid_size = opt.att_hid_size
        self.retrieval_reward = opt.retrieval_reward
        self.gumbel_temp = opt.gumbel_temp
        self.multinomial_temp = opt.multinomial_temp
        self.prob_gumbel_softmax = getattr(opt, "prob_gumbel_softmax", 1)
        self.prob_multinomial_soft = getattr(opt, "prob_multinomial_soft", 1)
        self.use_bn = getattr(opt, "use_bn", 0)
        if self.use_bn:
            raise NotImplementedError("use_bn=1 (BatchNorm on att feats) is outside the hot path")
        if self.num_layers != 1:
            raise NotImplementedError("att2in2 is a single-layer core (AttModel.py:492-539)")
        self.ss_prob = 0.0
        self.embed = nn.Sequential(nn.Embedding(self.vocab_size + 2, self.input_encoding_size),
                                   nn.ReLU(), nn.Dropout(self.drop_prob_lm))
        self.att_embed = nn.Sequential(nn.Linear(self.att_feat_size, self.rnn_size), nn.ReLU(),
                                       nn.Dropout(self.drop_prob_lm))
        self.logit = nn.Linear(self.rnn_size, self.vocab_size + 1)
        self.ctx2att = nn.Linear(self.rnn_size, self.att_hid_size)
        self.decoding_constraint = getattr(opt, "decoding_constraint", 0)
        self._loss = {}
        # parity hooks (tests): injected randomness / forced ids for the next pass
        self.injected: Optional[EN.SpeakerRandom] = None
        self.forced_tokens: Optional[torch.Tensor] = None     # int64 [B, T]
        self.keep_passes = False          # tests: retain every SpeakerPass in self._passes
        self._passes = []
        self._packed = EN.PackedSpeaker()

    # ------------------------------------------------------------------ plumbing
    def _params(self) -> Dict[str, torch.Tensor]:
        return {
            "embed.0.weight": self.embed[0].weight,
            "att_embed.0.weight": self.att_embed[0].weight, "att_embed.0.bias": self.att_embed[0].bias,
            "logit.weight": self.logit.weight, "logit.bias": self.logit.bias,
            "ctx2att.weight": self.ctx2att.weight, "ctx2att.bias": self.ctx2att.bias,
            "core.a2c.weight": self.core.a2c.weight, "core.a2c.bias": self.core.a2c.bias,
            "core.i2h.weight": self.core.i2h.weight, "core.i2h.bias": self.core.i2h.bias,
            "core.h2h.weight": self.core.h2h.weight, "core.h2h.bias": self.core.h2h.bias,
            "core.attention.h2att.weight": self.core.attention.h2att.weight,
            "core.attention.h2att.bias": self.core.attention.h2att.bias,
            "core.attention.alpha_net.weight": self.core.attention.alpha_net.weight,
            "core.attention.alpha_net.bias": self.core.attention.alpha_net.bias,
        }

    def _random(self) -> EN.SpeakerRandom:
        p = float(self.drop_prob_lm) if self.training else 0.0
        if self.injected is not None:
            return self.injected
        return EN.SpeakerRandom(seed=_next_seed(), drop_p=p)

    def _run(self, att_feats, att_masks, *, n_steps, mode, inv_tau, start_token, forced=None,
             start_tokens=None, ps_prob=0.0, ss_prob=0.0, no_repeat=False,
             store_perturbed=False) -> EN.SpeakerPass:
        if not att_feats.is_cuda:
            raise EN._lib.CoopcapError("Att2in2Model runs on CUDA only (no CPU path)")
        P = self._params()
        packed = self._packed.get(P)
        att16 = getattr(att_masks, "_coopcap_att16", None) if att_masks is not None else None
        if att16 is None:
            att_feats = att_feats.detach().float().contiguous()
        B, L = att_feats.shape[:2]
        pre = getattr(att_masks, "_coopcap_off", None) if att_masks is not None else None
        if pre is not None:
            off, NL = pre         # offsets supplied by the loader side: no device sync
        else:
            if att_masks is not None:
                # precondition of the reference (Appendix D): padded width == longest row
                att_masks = att_masks[:, :L]
            # plain tensors: the region count is read back lazily (speaker_forward resolves it after
            # its NL-independent host work; the pre-packed path needs it for the operand's shape)
            off, NL = EN.region_offsets(att_masks, B, L, lazy=att16 is None)
        # rows by decreasing region count (schedule hint), if the loader side already knows it
        order = getattr(att_masks, "_coopcap_order", None) if att_masks is not None else None
        sp = EN.speaker_forward(P, packed, att_feats, off, NL, n_steps=n_steps, mode=mode,
                                inv_tau=inv_tau, start_token=start_token, rnd=self._random(),
                                forced=forced, start_tokens=start_tokens, att16=att16, att_order=order,
                                ps_prob=ps_prob, ss_prob=ss_prob, no_repeat=no_repeat,
                                store_perturbed=store_perturbed,
                                w_embed16=self._packed.get_embed16(P) if mode in EN.PS_MODES else None)
        if self.keep_passes:
            sp.pinned = True
            self._passes.append(sp)
        return sp

    def _needs_grad(self) -> bool:
        return torch.is_grad_enabled() and any(p.requires_grad for p in self._params().values())

    # ------------------------------------------------------------------ AttModel.forward
    def forward(self, fc_feats, att_feats, att_masks, seq, masks):
        """Teacher-forced XE loss (AttModel.py:103-148 + misc/utils.py:49-58)."""
        seq = seq.long()
        T1 = seq.size(1) - 1
        # stop at the first i >= 1 whose whole column is 0 (AttModel.py:133): one small D2H
        nz = (seq[:, 1:T1] != 0).any(0).cpu().tolist() if T1 > 1 else []
        n_steps = 1
        for i, alive in enumerate(nz, start=1):
            if not alive:
                break
            n_steps = i + 1
        forced = seq[:, 1:n_steps + 1].t().contiguous()                    # targets = next inputs
        # scheduled sampling (:119-131): with probability ss_prob a row is fed the id drawn from
        # the previous step's distribution instead of the ground truth (no gradient through it)
        ss = float(self.ss_prob) if self.training else 0.0
        sp = self._run(att_feats, att_masks, n_steps=n_steps,
                       mode=EN.MODE_MULTINOMIAL if ss > 0.0 else EN.MODE_NONE, inv_tau=1.0,
                       start_token=0, forced=forced, start_tokens=seq[:, 0].contiguous(), ss_prob=ss)
        if self._needs_grad():
            logp = _SpeakerLossFn.apply(sp, forced, self, *_ordered(self._params()))
        else:
            logp = sp.t["logp"][:n_steps]
        m = masks[:, 1:n_steps + 1].t().to(logp.dtype)
        loss = -(logp * m).sum() / m.sum()
        self._loss["xe"] = loss.detach()
        return loss

    # ------------------------------------------------------------------ AttModel.sample
    def sample(self, fc_feats, att_feats, att_masks, opt={}):
        """AttModel.py:291-452.  Returns (seq, seqLogprobs) or, with use_one_hot in a
        straight-through mode, (word_index, one_hots, logprobs)."""
        sample_max = opt.get("sample_max", 1)
        beam_size = opt.get("beam_size", 1)
        temperature = opt.get("temperature", 1.0)
        use_one_hot = opt.get("use_one_hot", 0)
        if beam_size > 1:
            return self.sample_beam(fc_feats, att_feats, att_masks, opt)                 # :307-308
        no_repeat = bool(opt.get("decoding_constraint", self.decoding_constraint))      # :305-306
        sp, st_mode = self._sample_pass(att_feats, att_masks, sample_max, temperature, use_one_hot,
                                        **({"no_repeat": True} if no_repeat else {}))
        n = int(sp.t["n_out"].item())                       # the one host sync (output width)
        seq = sp.t["tok_out"][:n].t().contiguous()
        if n == 0:
            # the reference crashes here on torch.cat([]) (Appendix D); return width-0 tensors
            z = torch.zeros(sp.B, 0, device=seq.device)
            return (seq, torch.zeros(sp.B, 0, sp.dims.V1 + 1, device=seq.device), z) if st_mode \
                else (seq, z)
        if self._needs_grad() and not sample_max and sp.ctx.mode not in EN.PS_MODES:
            # (partial-sampling passes return sampleLogprobs detached: their BPTT is driven by the
            # gradient of the emitted vectors, the only use the reference's joint step makes of them)
            logp_all = _SpeakerLossFn.apply(sp, sp.t["tok_fed"][1:sp.n_steps + 1].contiguous(), self,
                                            *_ordered(self._params()))
        else:
            logp_all = sp.t["logp"][: sp.n_steps]
        logprobs = logp_all[:n].t()
        if st_mode:
            fn = _SamplePSFn if sp.ctx.mode in EN.PS_MODES else _SampleSTFn
            if self._needs_grad():
                one_hots = fn.apply(sp, n, self, *_ordered(self._params()))
            else:
                one_hots = fn.forward(_Dummy(), sp, n, self)
            return seq, one_hots, logprobs
        return seq, logprobs

    _sample = sample

    # ------------------------------------------------------------------ AttModel.sample_beam
    def sample_beam(self, fc_feats, att_feats, att_masks, opt={}):
        """Beam search (AttModel.py:150-289) for all images at once: returns (seq [B, T] int64,
        seqLogprobs [B, T]) and fills `self.done_beams` (per image: the recorded beams as dicts
        'seq' / 'logps' / 'p', ranked as the reference ranks them).  Evaluation mode only: the
        reference's own callers run it under model.eval() (eval_utils.py:187)."""
        beam_size = opt.get("beam_size", 10)
        no_repeat = bool(opt.get("decoding_constraint", self.decoding_constraint))
        if self.training and float(self.drop_prob_lm) > 0.0:
            raise RuntimeError("sample_beam runs in evaluation mode (call model.eval() first)")
        if not att_feats.is_cuda:
            raise EN._lib.CoopcapError("Att2in2Model runs on CUDA only (no CPU path)")
        if beam_size > self.vocab_size + 1:
            raise AssertionError("beam_size must not exceed the vocabulary (AttModel.py:164)")
        P = self._params()
        packed = self._packed.get(P)
        att16 = getattr(att_masks, "_coopcap_att16", None) if att_masks is not None else None
        if att16 is None:
            att_feats = att_feats.detach().float().contiguous()
        B, L = att_feats.shape[:2]
        pre = getattr(att_masks, "_coopcap_off", None) if att_masks is not None else None
        if pre is not None:
            off, NL = pre
        else:
            if att_masks is not None:
                att_masks = att_masks[:, :L]
            off, NL = EN.region_offsets(att_masks, B, L)
        forced = getattr(self, "forced_beam", None)          # parity hook: (parent, tok) [T, B, beam]
        out = EN.beam_search(P, packed, att_feats, off, NL, beam_size=beam_size,
                             seq_length=self.seq_length, start_token=self.vocab_size + 1,
                             no_repeat=no_repeat, att16=att16,
                             att_order=getattr(att_masks, "_coopcap_order", None) if att_masks is not None else None,
                             forced_parent=None if forced is None else forced[0],
                             forced_tok=None if forced is None else forced[1])
        self._beam_last = out
        self.done_beams = _LazyDoneBeams(out, beam_size, self.seq_length)
        return out.seq, out.logprobs

    def _sample_pass(self, att_feats, att_masks, sample_max, temperature, use_one_hot,
                     no_repeat=False, store_perturbed=False):
        """Run the decode loop in the mode AttModel.sample would pick; returns (pass, is_ST) with
        is_ST true for the modes that return dense vectors (straight-through / partial sampling)."""
        T = self.seq_length
        st_mode = False
        ps_prob = 0.0
        if sample_max:
            mode, inv_tau = EN.MODE_GREEDY, 1.0
        elif self.retrieval_reward == "reinforce" or not use_one_hot:
            mode, inv_tau = EN.MODE_MULTINOMIAL, 1.0 / float(temperature)
        elif self.retrieval_reward == "gumbel":
            mode, inv_tau, st_mode = EN.MODE_ST_GUMBEL, 1.0 / float(self.gumbel_temp), True
        elif self.retrieval_reward == "multinomial":
            mode, inv_tau, st_mode = EN.MODE_ST_MULTINOMIAL, 1.0 / float(self.multinomial_temp), True
        elif self.retrieval_reward == "gumbel_softmax":                             # :367-378
            mode, inv_tau, st_mode = EN.MODE_PS_GUMBEL, 1.0 / float(self.gumbel_temp), True
            ps_prob = float(self.prob_gumbel_softmax)
        elif self.retrieval_reward == "multinomial_soft":                           # :381-392
            mode, inv_tau, st_mode = EN.MODE_PS_MULTINOMIAL, 1.0 / float(self.multinomial_temp), True
            ps_prob = float(self.prob_multinomial_soft)
        else:
            raise ValueError(f"unknown retrieval_reward {self.retrieval_reward!r}")
        if no_repeat and st_mode:
            # the reference scatters with seq[-1], a dense tensor in these modes, and crashes
            raise ValueError("decoding_constraint needs index outputs (sample_max or use_one_hot=0)")
        forced = None
        if self.forced_tokens is not None:
            forced = self.forced_tokens.t().contiguous()
        sp = self._run(att_feats, att_masks, n_steps=T, mode=mode, inv_tau=inv_tau,
                       start_token=self.vocab_size + 1, forced=forced, ps_prob=ps_prob,
                       no_repeat=no_repeat, store_perturbed=store_perturbed)
        return sp, st_mode


class _Dummy:
    pass


class _LazyDoneBeams:
    """`self.done_beams[k]` of the reference (AttModel.py:168,283-284): the recorded beams of image
    k sorted by -p.  Materialised on first access (one device->host copy), so a caller that only
    wants (seq, seqLogprobs) never synchronises."""

    def __init__(self, out, beam_size, T):
        self._out, self._bs, self._T, self._lists = out, beam_size, T, None

    def _build(self):
        if self._lists is None:
            t = self._out.t
            n = t["done_n"].cpu().tolist()
            seq, lp, p = t["done_seq"].cpu(), t["done_lp"].cpu(), t["done_p"].cpu()
            rec = t["done_p_rec"].cpu()
            self._lists = []
            for k, nk in enumerate(n):
                ent = [dict(seq=seq[k, e].clone(), logps=lp[k, e].clone(), p=float(p[k, e]),
                            p_recorded=float(rec[k, e])) for e in range(nk)]
                self._lists.append(sorted(ent, key=lambda x: -x["p"]))        # stable, as :283-284
        return self._lists

    def __getitem__(self, k):
        return self._build()[k]

    def __len__(self):
        return len(self._build())

    def __iter__(self):
        return iter(self._build())


class Att2in2Model(AttModel):
    def __init__(self, opt):
        super().__init__(opt)
        self.core = Att2in2Core(opt)
        # the reference deletes fc_embed and replaces it by the identity (AttModel.py:538-539):
        # fc_feats are never read by this speaker
        self.fc_embed = lambda x: x
