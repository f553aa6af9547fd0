"""Host-side marshalling for libcoopcap's speaker / listener passes.

Nothing here computes: the functions allocate device buffers with torch (caching allocator, no
synchronisation), fill the C context structs of include/coopcap.h and launch the C entry points on
torch's current stream.  The autograd wiring lives in models/*.py.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import check

MODE_GREEDY, MODE_MULTINOMIAL, MODE_ST_GUMBEL, MODE_ST_MULTINOMIAL, MODE_NONE = 0, 1, 2, 3, 4


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.CoopcapError("coopcap needs CUDA tensors (there is no CPU path)")


def _f32c(t):
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise _lib.CoopcapError("expected a contiguous fp32 tensor")
    return t


# =============================================================================================
# speaker
# =============================================================================================
SPEAKER_PARAM_NAMES = [
    "embed.0.weight", "att_embed.0.weight", "att_embed.0.bias", "logit.weight", "logit.bias",
    "ctx2att.weight", "ctx2att.bias", "core.a2c.weight", "core.a2c.bias", "core.i2h.weight",
    "core.i2h.bias", "core.h2h.weight", "core.h2h.bias", "core.attention.h2att.weight",
    "core.attention.h2att.bias", "core.attention.alpha_net.weight", "core.attention.alpha_net.bias",
]


@dataclass
class SpeakerDims:
    D: int
    R: int
    E: int
    A: int
    V1: int

    @staticmethod
    def of(P: Dict[str, torch.Tensor]) -> "SpeakerDims":
        return SpeakerDims(D=P["att_embed.0.weight"].shape[1], R=P["core.h2h.weight"].shape[1],
                           E=P["core.i2h.weight"].shape[1], A=P["ctx2att.weight"].shape[0],
                           V1=P["logit.weight"].shape[0])


class PackedSpeaker:
    """bf16 operand copies of the speaker parameters (rebuilt when a parameter changes)."""

    def __init__(self):
        self.key = None
        self.buf = {}

    def get(self, P: Dict[str, torch.Tensor]):
        key = tuple((P[n].data_ptr(), P[n]._version) for n in SPEAKER_PARAM_NAMES)
        if key == self.key:
            return self.buf
        d = SpeakerDims.of(P)
        dev = P["logit.weight"].device
        _need_cuda(*[P[n] for n in SPEAKER_PARAM_NAMES])
        if not self.buf or self.buf["dims"] != d:
            bf = dict(dtype=torch.bfloat16, device=dev)
            self.buf = dict(
                dims=d,
                w_att_embed16=torch.empty(d.R, d.D, **bf), w_ctx2att16=torch.empty(d.A, d.R, **bf),
                w_cat16=torch.empty(5 * d.R + d.A, d.E + d.R, **bf),
                w_a2c16=torch.empty(2 * d.R, d.R, **bf), w_logit16=torch.empty(d.V1, d.R, **bf),
                b_cat=torch.empty(5 * d.R + d.A, dtype=torch.float32, device=dev))
        b = self.buf
        a = _lib.SpeakerPack()
        a.D, a.R, a.E, a.A, a.V1 = d.D, d.R, d.E, d.A, d.V1
        a.w_att_embed = _p(_f32c(P["att_embed.0.weight"].detach()))
        a.w_ctx2att = _p(_f32c(P["ctx2att.weight"].detach()))
        a.w_i2h = _p(_f32c(P["core.i2h.weight"].detach()))
        a.w_h2h = _p(_f32c(P["core.h2h.weight"].detach()))
        a.w_h2att = _p(_f32c(P["core.attention.h2att.weight"].detach()))
        a.w_a2c = _p(_f32c(P["core.a2c.weight"].detach()))
        a.w_logit = _p(_f32c(P["logit.weight"].detach()))
        a.b_i2h = _p(_f32c(P["core.i2h.bias"].detach()))
        a.b_h2h = _p(_f32c(P["core.h2h.bias"].detach()))
        a.b_h2att = _p(_f32c(P["core.attention.h2att.bias"].detach()))
        a.w_att_embed16, a.w_ctx2att16 = _p(b["w_att_embed16"]), _p(b["w_ctx2att16"])
        a.w_cat16, a.w_a2c16, a.w_logit16 = _p(b["w_cat16"]), _p(b["w_a2c16"]), _p(b["w_logit16"])
        a.b_cat = _p(b["b_cat"])
        check(_lib.load().coopcap_speaker_pack_weights(C.byref(a), _stream()))
        self.key = key
        return b


@dataclass
class SpeakerRandom:
    """Randomness of one speaker pass: Philox seed, or injected tensors (parity tests)."""
    seed: int = 0
    drop_p: float = 0.0
    keep_att: Optional[torch.Tensor] = None     # uint8 [NL, R] packed
    keep_embed: Optional[torch.Tensor] = None   # uint8 [steps+1, B, E]
    keep_core: Optional[torch.Tensor] = None    # uint8 [steps, B, R]
    noise: Optional[torch.Tensor] = None        # fp32 [steps, B, V1]


@dataclass
class SpeakerPass:
    """All buffers of one speaker pass (kept alive by the autograd node that needs them)."""
    ctx: object = None
    dims: SpeakerDims = None
    B: int = 0
    L: int = 0
    NL: int = 0
    cap: int = 0
    n_steps: int = 0
    t: Dict[str, torch.Tensor] = field(default_factory=dict)
    keep: list = field(default_factory=list)   # tensors referenced by raw pointer in ctx


def region_offsets(att_masks: Optional[torch.Tensor], B: int, L: int):
    """att_masks [B, L] -> (att_off int32 [B+1] on device or None, NL).  One host sync (the
    valid-region count sizes the packed buffers), as pack_wrapper's own lengths do
    (AttModel.py:47)."""
    if att_masks is None:
        return None, B * L
    lens = (att_masks > 0).sum(1).to(torch.int32)
    off = torch.zeros(B + 1, dtype=torch.int32, device=att_masks.device)
    off[1:] = torch.cumsum(lens, 0)
    return off, int(off[-1].item())


def speaker_forward(P: Dict[str, torch.Tensor], packed: dict, att_feats: torch.Tensor,
                    att_off: Optional[torch.Tensor], NL: int, *, n_steps: int, mode: int,
                    inv_tau: float, start_token: int, rnd: SpeakerRandom,
                    forced: Optional[torch.Tensor] = None) -> SpeakerPass:
    """Prologue + n_steps decode steps.  `forced` int64 [n_steps, B] (time-major)."""
    _need_cuda(att_feats, att_off, forced)
    d: SpeakerDims = packed["dims"]
    att_feats = _f32c(att_feats)
    B, L, D = att_feats.shape
    assert D == d.D
    dev = att_feats.device
    cap = max(n_steps, 1)
    f32 = dict(dtype=torch.float32, device=dev)
    bf = dict(dtype=torch.bfloat16, device=dev)
    i64 = dict(dtype=torch.int64, device=dev)
    NS, XH = 5 * d.R + d.A, d.E + d.R
    T = dict(
        att16=torch.empty(NL, d.D, **bf), att_e16=torch.empty(NL, d.R, **bf),
        p_att16=torch.empty(NL, d.A, **bf), xh16=torch.empty(cap + 1, B, XH, **bf),
        s_all=torch.empty(cap, B, NS, **f32), u_all=torch.empty(cap, B, 2 * d.R, **f32),
        c_all=torch.empty(cap + 1, B, d.R, **f32), att_res16=torch.empty(cap, B, d.R, **bf),
        att_w=torch.empty(cap, NL, **f32), out16=torch.empty(cap, B, d.R, **bf),
        z_all=torch.empty(cap, B, d.V1, **f32), tok_raw=torch.empty(cap, B, **i64),
        tok_out=torch.empty(cap, B, **i64), tok_fed=torch.empty(cap + 1, B, **i64),
        logp=torch.empty(cap, B, **f32), lse=torch.empty(cap, B, **f32),
        y_max=torch.empty(cap, B, **f32), y_sum=torch.empty(cap, B, **f32),
        unfinished=torch.empty(cap, B, dtype=torch.uint8, device=dev),
        n_out=torch.empty(1, dtype=torch.int32, device=dev),
        cap_len=torch.empty(B, dtype=torch.int32, device=dev))
    c = _lib.Speaker()
    c.B, c.L, c.D, c.R, c.E, c.A, c.V1 = B, L, d.D, d.R, d.E, d.A, d.V1
    c.NL, c.cap, c.n_steps = NL, cap, n_steps
    c.att_feats, c.att_off = _p(att_feats), _p(att_off)
    c.embed = _p(_f32c(P["embed.0.weight"].detach()))
    c.b_att_embed = _p(_f32c(P["att_embed.0.bias"].detach()))
    c.b_ctx2att = _p(_f32c(P["ctx2att.bias"].detach()))
    c.b_cat = _p(packed["b_cat"])
    c.b_a2c = _p(_f32c(P["core.a2c.bias"].detach()))
    c.b_logit = _p(_f32c(P["logit.bias"].detach()))
    c.w_alpha = _p(_f32c(P["core.attention.alpha_net.weight"].detach()))
    for n in ("w_att_embed16", "w_ctx2att16", "w_cat16", "w_a2c16", "w_logit16"):
        setattr(c, n, _p(packed[n]))
    c.seed, c.drop_p = int(rnd.seed) & (2 ** 64 - 1), float(rnd.drop_p)
    for n in ("keep_att", "keep_embed", "keep_core"):
        k = getattr(rnd, n)
        if k is not None:
            assert k.dtype == torch.uint8 and k.is_contiguous() and k.is_cuda
        setattr(c, n, _p(k))
    if rnd.noise is not None:
        _f32c(rnd.noise)
    c.noise = _p(rnd.noise)
    c.mode, c.inv_tau, c.start_token = mode, float(inv_tau), int(start_token)
    if forced is not None:
        assert forced.dtype == torch.int64 and forced.is_contiguous() and forced.shape == (n_steps, B)
    c.forced = _p(forced)
    for n, tsr in T.items():
        setattr(c, n, _p(tsr))
    sp = SpeakerPass(ctx=c, dims=d, B=B, L=L, NL=NL, cap=cap, n_steps=n_steps, t=T,
                     keep=[att_feats, att_off, forced, rnd, packed, P])
    lib = _lib.load()
    check(lib.coopcap_speaker_prologue_fwd(C.byref(c), _stream()))
    check(lib.coopcap_speaker_decode_fwd(C.byref(c), _stream()))
    return sp
