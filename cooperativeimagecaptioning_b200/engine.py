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
MODE_PS_GUMBEL, MODE_PS_MULTINOMIAL = 5, 6
PS_MODES = (MODE_PS_GUMBEL, MODE_PS_MULTINOMIAL)


_weights_epoch = 0


def bump_weights_epoch():
    """Called by anything that modifies parameters outside autograd's version tracking (the fused
    optimizer kernel); invalidates the packed bf16 operand copies."""
    global _weights_epoch
    _weights_epoch += 1


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


# ---------------------------------------------------------------------------------------------
# L2-resident weight arena.  The bf16 operand copies of the RECURRENT weights (speaker w_cat16 /
# w_a2c16 / w_logit16, listener w_hh16: 23 MB at the reference sizes) are read by ~100 per-step GEMMs
# of a training step, and every decode step streams ~220 MB of region tensors, logits and gate
# pre-activations through the 126 MB L2 between two of those reads.  They live in one arena so that
# a single persisting access-policy window (coopcap_l2_persist) can keep them in L2.
# COOPCAP_L2_PERSIST=0 turns the window off (the arena stays).
# ---------------------------------------------------------------------------------------------
_ARENA_BYTES = 40 << 20
_arena: Dict[int, dict] = {}


def _arena_alloc(dev: torch.device, shape, dtype=torch.bfloat16) -> torch.Tensor:
    """A tensor of `shape` carved out of the device's weight arena (256-byte aligned), or an
    ordinary allocation when the arena is full."""
    import math
    ix = dev.index if dev.index is not None else torch.cuda.current_device()
    a = _arena.get(ix)
    if a is None:
        a = _arena[ix] = dict(buf=torch.empty(_ARENA_BYTES, dtype=torch.uint8, device=dev), used=0,
                              windows={})
    nbytes = math.prod(shape) * dtype.itemsize
    start = (a["used"] + 255) // 256 * 256
    if start + nbytes > _ARENA_BYTES:
        return torch.empty(shape, dtype=dtype, device=dev)
    a["used"] = start + nbytes
    return a["buf"][start:start + nbytes].view(dtype).view(shape)


def _arena_persist(dev: torch.device):
    """Put the persisting window over the used part of the arena on the CURRENT stream (once per
    stream and arena size)."""
    import os
    if os.environ.get("COOPCAP_L2_PERSIST", "1") == "0":
        return
    ix = dev.index if dev.index is not None else torch.cuda.current_device()
    a = _arena.get(ix)
    if a is None or a["used"] == 0:
        return
    st = torch.cuda.current_stream(dev).cuda_stream
    if a["windows"].get(st) == a["used"]:
        return
    granted = C.c_int64(0)
    check(_lib.load().coopcap_l2_persist(C.c_void_p(a["buf"].data_ptr()), a["used"], C.byref(granted),
                                        C.c_void_p(st)))
    a["windows"][st] = a["used"]
    a["granted"] = int(granted.value)


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
        self.embed_key = None
        self.embed16 = None

    def get_embed16(self, P: Dict[str, torch.Tensor]) -> torch.Tensor:
        """bf16 copy of `embed` [V+2, E]: the operand of the partial-sampling next-input GEMM."""
        w = P["embed.0.weight"]
        key = (_weights_epoch, w.data_ptr(), w._version)
        if key != self.embed_key:
            _need_cuda(w)
            if self.embed16 is None or self.embed16.shape != w.shape:
                self.embed16 = torch.empty(w.shape, dtype=torch.bfloat16, device=w.device)
            from . import ops
            ops.cast_bf16(_f32c(w.detach()), dst=self.embed16)
            self.embed_key = key
        return self.embed16

    def get(self, P: Dict[str, torch.Tensor]):
        key = (_weights_epoch,) + tuple((P[n].data_ptr(), P[n]._version) for n in SPEAKER_PARAM_NAMES)
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
                w_cat16=_arena_alloc(dev, (5 * d.R + d.A, d.E + d.R)),
                w_a2c16=_arena_alloc(dev, (2 * d.R, d.R)), w_logit16=_arena_alloc(dev, (d.V1, d.R)),
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
        _arena_persist(dev)
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
    part_u: Optional[torch.Tensor] = None       # fp32 [steps, B] partial-sampling row selection
    ss_u: Optional[torch.Tensor] = None         # fp32 [steps, B] scheduled-sampling row selection


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


def region_offsets(att_masks: Optional[torch.Tensor], B: int, L: int, lazy: bool = False):
    """att_masks [B, L] -> (att_off int32 [B+1] on device or None, NL).  One host sync (the
    valid-region count sizes the packed buffers), as pack_wrapper's own lengths do
    (AttModel.py:47).  lazy=True returns a zero-argument callable instead of NL: the count is
    already on its way to pinned host memory, the caller resolves it (the sync) as late as it can --
    speaker_forward does all its NL-independent host work first, so the device is idle only for
    the few microseconds between the sync and the first launch."""
    if att_masks is None:
        return None, B * L
    lens = (att_masks > 0).sum(1).to(torch.int32)
    off = torch.zeros(B + 1, dtype=torch.int32, device=att_masks.device)
    off[1:] = torch.cumsum(lens, 0)
    if not lazy:
        return off, int(off[-1].item())
    host = torch.empty(1, dtype=torch.int32, pin_memory=True)
    host.copy_(off[-1:], non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()

    def resolve() -> int:
        ev.synchronize()
        return int(host[0])
    return off, resolve


def speaker_forward(P: Dict[str, torch.Tensor], packed: dict, att_feats: torch.Tensor,
                    att_off: Optional[torch.Tensor], NL: int, *, n_steps: int, mode: int,
                    inv_tau: float, start_token: int, rnd: SpeakerRandom,
                    forced: Optional[torch.Tensor] = None,
                    start_tokens: Optional[torch.Tensor] = None,
                    att16: Optional[torch.Tensor] = None, ps_prob: float = 0.0,
                    w_embed16: Optional[torch.Tensor] = None, ss_prob: float = 0.0,
                    no_repeat: bool = False, att_order: Optional[torch.Tensor] = None,
                    store_perturbed: bool = False) -> SpeakerPass:
    """Prologue + n_steps decode steps.  `forced` int64 [n_steps, B] (time-major);
    `start_tokens` int64 [B] overrides the scalar start id per row.  `store_perturbed` (ST-Gumbel):
    keep z + G instead of z for backward (coopcap_speaker.store_perturbed)."""
    _need_cuda(att_feats, att_off, forced, start_tokens)
    d: SpeakerDims = packed["dims"]
    B, L, D = att_feats.shape
    assert D == d.D
    nl_lazy = callable(NL)                 # region_offsets(..., lazy=True): resolved right before the launch
    if att16 is None:
        att_feats = _f32c(att_feats)       # `att16` given: att_feats is only a shape carrier
    else:
        assert not nl_lazy
        assert att16.dtype == torch.bfloat16 and att16.is_contiguous() and att16.shape == (NL, d.D)
    dev = att16.device if att16 is not None else att_feats.device
    cap = max(n_steps, 1)
    f32 = dict(dtype=torch.float32, device=dev)
    bf = dict(dtype=torch.bfloat16, device=dev)
    i64 = dict(dtype=torch.int64, device=dev)
    NS, XH = 5 * d.R + d.A, d.E + d.R
    T = dict(
        xh16=torch.empty(cap + 1, B, XH, **bf),
        s_all=torch.empty(cap, B, NS, **f32), u_all=torch.empty(cap, B, 2 * d.R, **f32),
        c_all=torch.empty(cap + 1, B, d.R, **f32), att_res16=torch.empty(cap, B, d.R, **bf),
        # fp32 attention output: lets the per-step attention backward run as one pass (attention.cuh v5)
        att_res32=torch.empty(cap, B, d.R, **f32),
        out16=torch.empty(cap, B, d.R, **bf),
        # logits are kept in fp16 for the backward pass only; the sampler runs on the fp32
        # accumulators inside the logit GEMM (csrc/logit_sample.cuh)
        z16_all=torch.empty(cap, B, d.V1, dtype=torch.float16, device=dev),
        ls_part=torch.empty(B, 4 * ((d.V1 + 255) // 256), 8, **f32), z_tgt=torch.empty(B, **f32),
        tok_raw=torch.empty(cap, B, **i64),
        tok_out=torch.empty(cap, B, **i64), tok_fed=torch.empty(cap + 1, B, **i64),
        logp=torch.empty(cap, B, **f32), lse=torch.empty(cap, B, **f32),
        y_max=torch.empty(cap, B, **f32), y_sum=torch.empty(cap, B, **f32),
        unfinished=torch.empty(cap, B, dtype=torch.uint8, device=dev),
        n_out=torch.empty(1, dtype=torch.int32, device=dev),
        cap_len=torch.empty(B, dtype=torch.int32, device=dev))
    if mode in PS_MODES:
        if w_embed16 is None:
            raise _lib.CoopcapError("partial-sampling modes need the bf16 embedding copy")
        T["soft16"] = torch.empty(cap, B, d.V1, **bf)
        T["ps_sel"] = torch.empty(cap, B, dtype=torch.uint8, device=dev)
    c = _lib.Speaker()
    c.B, c.L, c.D, c.R, c.E, c.A, c.V1 = B, L, d.D, d.R, d.E, d.A, d.V1
    c.cap, c.n_steps = cap, n_steps
    c.att_feats, c.att_off = (None if att16 is not None else _p(att_feats)), _p(att_off)
    # longest rows first: the attention kernels deal rows to SMs by rank (load balance only)
    if att_order is None and att_off is not None and B > 1:
        att_order = torch.argsort(att_off[1:] - att_off[:-1], descending=True).to(torch.int32)
    c.att_order = _p(att_order)
    c.att_prepacked = int(att16 is not None)
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
    want = {"keep_att": (None, d.R), "keep_embed": (n_steps, B, d.E), "keep_core": (n_steps, B, d.R)}
    for n in ("keep_att", "keep_embed", "keep_core"):
        k = getattr(rnd, n)
        if k is not None:
            assert k.dtype == torch.uint8 and k.is_contiguous() and k.is_cuda
            w = want[n]
            if (w[0] is not None and k.shape[0] < w[0]) or tuple(k.shape[1:]) != tuple(w[1:]):
                raise _lib.CoopcapError(f"injected {n} has shape {tuple(k.shape)}, need >= {w}")
        setattr(c, n, _p(k))
    uses_noise = mode in (MODE_MULTINOMIAL, MODE_ST_GUMBEL, MODE_ST_MULTINOMIAL) + PS_MODES
    if rnd.noise is not None and uses_noise:
        _f32c(rnd.noise)
        if rnd.noise.shape[0] < n_steps or tuple(rnd.noise.shape[1:]) != (B, d.V1):
            raise _lib.CoopcapError(f"injected noise has shape {tuple(rnd.noise.shape)}, "
                                    f"need >= {(n_steps, B, d.V1)}")
    c.noise = _p(rnd.noise) if uses_noise else None
    c.mode, c.inv_tau, c.start_token = mode, float(inv_tau), int(start_token)
    c.ps_prob, c.ss_prob, c.no_repeat = float(ps_prob), float(ss_prob), int(bool(no_repeat))
    c.store_perturbed = int(bool(store_perturbed) and mode == MODE_ST_GUMBEL)
    c.w_embed16 = _p(w_embed16)
    for n, on in (("part_u", mode in PS_MODES), ("ss_u", ss_prob > 0.0)):
        u = getattr(rnd, n)
        if u is not None and on:
            _f32c(u)
            if u.shape[0] < n_steps or tuple(u.shape[1:]) != (B,):
                raise _lib.CoopcapError(f"injected {n} has shape {tuple(u.shape)}, need >= {(n_steps, B)}")
            setattr(c, n, _p(u))
    if forced is not None:
        assert forced.dtype == torch.int64 and forced.is_contiguous() and forced.shape == (n_steps, B)
    c.forced = _p(forced)
    if start_tokens is not None:
        assert start_tokens.dtype == torch.int64 and start_tokens.is_contiguous()
        assert start_tokens.shape == (B,)
    c.start_tokens = _p(start_tokens)
    # ---- everything above is independent of the region count; the (only) host sync of the
    # plain-tensor call path happens here, with the launches a few microseconds away
    if nl_lazy:
        NL = int(NL())
    if rnd.keep_att is not None and rnd.keep_att.shape[0] < NL:
        raise _lib.CoopcapError(f"injected keep_att has shape {tuple(rnd.keep_att.shape)}, need >= {(NL, d.R)}")
    c.NL = NL
    T.update(att16=att16 if att16 is not None else torch.empty(NL, d.D, **bf),
             att_e16=torch.empty(NL, d.R, **bf), p_att16=torch.empty(NL, d.A, **bf),
             att_w=torch.empty(cap, NL, **f32))
    for n, tsr in T.items():
        setattr(c, n, _p(tsr))
    sp = SpeakerPass(ctx=c, dims=d, B=B, L=L, NL=NL, cap=cap, n_steps=n_steps, t=T,
                     keep=[att_feats, att_off, att_order, forced, start_tokens, rnd, packed, P, w_embed16])
    lib = _lib.load()
    check(lib.coopcap_speaker_prologue_fwd(C.byref(c), _stream()))
    check(lib.coopcap_speaker_decode_fwd(C.byref(c), _stream()))
    return sp


@dataclass
class BeamOutput:
    seq: torch.Tensor          # int64 [n_img, T]
    logprobs: torch.Tensor     # fp32 [n_img, T]
    t: Dict[str, torch.Tensor] = field(default_factory=dict)


def beam_search(P: Dict[str, torch.Tensor], packed: dict, att_feats: torch.Tensor,
                att_off: Optional[torch.Tensor], NL: int, *, beam_size: int, seq_length: int,
                start_token: int, no_repeat: bool = False, att16: Optional[torch.Tensor] = None,
                att_order: Optional[torch.Tensor] = None, forced_parent: Optional[torch.Tensor] = None,
                forced_tok: Optional[torch.Tensor] = None) -> BeamOutput:
    """AttModel.sample_beam (AttModel.py:150-289) for all images at once (evaluation mode).
    forced_parent int32 / forced_tok int64 [T, n_img, beam]: decisions to replay (parity tests)."""
    _need_cuda(att_feats, att_off, forced_parent, forced_tok)
    d: SpeakerDims = packed["dims"]
    B, L, D = att_feats.shape
    if att16 is None:
        att_feats = _f32c(att_feats)
    dev = att16.device if att16 is not None else att_feats.device
    f32 = dict(dtype=torch.float32, device=dev)
    bf = dict(dtype=torch.bfloat16, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    i64 = dict(dtype=torch.int64, device=dev)
    T, bs = int(seq_length), int(beam_size)
    rows, NS, XH = B * bs, 5 * d.R + d.A, d.E + d.R
    c = _lib.Speaker()
    c.B, c.L, c.D, c.R, c.E, c.A, c.V1 = B, L, d.D, d.R, d.E, d.A, d.V1
    c.NL, c.cap, c.n_steps = NL, 1, 0
    ctx_t = dict(att16=att16 if att16 is not None else torch.empty(NL, d.D, **bf),
                 att_e16=torch.empty(NL, d.R, **bf), p_att16=torch.empty(NL, d.A, **bf))
    c.att_feats, c.att_off = (None if att16 is not None else _p(att_feats)), _p(att_off)
    if att_order is None and att_off is not None and B > 1:
        att_order = torch.argsort(att_off[1:] - att_off[:-1], descending=True).to(torch.int32)
    c.att_order = _p(att_order)
    c.att_prepacked = int(att16 is not None)
    c.embed = _p(_f32c(P["embed.0.weight"].detach()))
    c.b_att_embed = _p(_f32c(P["att_embed.0.bias"].detach()))
    c.b_ctx2att = _p(_f32c(P["ctx2att.bias"].detach()))
    c.b_cat = _p(packed["b_cat"])
    c.b_a2c = _p(_f32c(P["core.a2c.bias"].detach()))
    c.b_logit = _p(_f32c(P["logit.bias"].detach()))
    c.w_alpha = _p(_f32c(P["core.attention.alpha_net.weight"].detach()))
    for n in ("w_att_embed16", "w_ctx2att16", "w_cat16", "w_a2c16", "w_logit16"):
        setattr(c, n, _p(packed[n]))
    c.seed, c.drop_p, c.mode, c.inv_tau, c.start_token = 0, 0.0, MODE_GREEDY, 1.0, int(start_token)
    for n, tsr in ctx_t.items():
        setattr(c, n, _p(tsr))
    W = dict(xh16=torch.empty(2, rows, XH, **bf), c2=torch.empty(2, rows, d.R, **f32),
             s_t=torch.empty(rows, NS, **f32), u_t=torch.empty(rows, 2 * d.R, **f32),
             att_res16=torch.empty(rows, d.R, **bf), att_w=torch.empty(bs, NL, **f32),
             h_stage16=torch.empty(rows, d.R, **bf), c_stage=torch.empty(rows, d.R, **f32),
             logits=torch.empty(rows, d.V1, **f32), parent=torch.empty(rows, **i32),
             tok=torch.empty(rows, **i64), hist_seq=torch.empty(2, B, T, bs, **i64),
             hist_lp=torch.empty(2, B, T, bs, **f32), beam_sum=torch.empty(B, bs, **f32),
             raw_parent=torch.empty(T, B, bs, **i32), raw_tok=torch.empty(T, B, bs, **i64),
             done_seq=torch.empty(B, bs * T, T, **i64), done_lp=torch.empty(B, bs * T, T, **f32),
             done_slot=torch.empty(B, bs * T, **i32), done_p_rec=torch.empty(B, bs * T, **f32),
             done_p=torch.empty(B, bs * T, **f32), done_n=torch.empty(B, **i32),
             seq=torch.empty(B, T, **i64), seq_logp=torch.empty(B, T, **f32))
    b = _lib.Beam()
    b.beam_size, b.no_repeat, b.T = bs, int(bool(no_repeat)), T
    for n, tsr in W.items():
        setattr(b, n, _p(tsr))
    if forced_parent is not None:
        assert forced_parent.dtype == torch.int32 and forced_parent.is_contiguous()
        assert forced_tok.dtype == torch.int64 and forced_tok.is_contiguous()
        assert forced_parent.shape == (T, B, bs) and forced_tok.shape == (T, B, bs)
    b.forced_parent, b.forced_tok = _p(forced_parent), _p(forced_tok)
    lib = _lib.load()
    check(lib.coopcap_speaker_prologue_fwd(C.byref(c), _stream()))
    check(lib.coopcap_speaker_beam_fwd(C.byref(c), C.byref(b), _stream()))
    W.update(ctx_t)
    out = BeamOutput(seq=W["seq"], logprobs=W["seq_logp"], t=W)
    out.keep = [att_feats, att_off, att_order, packed, P, forced_parent, forced_tok]
    return out


def retrieval_ranks(scores: torch.Tensor, first: torch.Tensor, count: int):
    """(ranks int32 [n_query], top1 int32 [n_query]) of a fp32 score matrix [n_query, n_cand];
    query q's correct candidates are first[q] .. first[q]+count-1 (eval_utils.py:545-720)."""
    _need_cuda(scores, first)
    scores = _f32c(scores)
    assert first.dtype == torch.int32 and first.is_contiguous() and first.numel() == scores.shape[0]
    nq, nc = scores.shape
    ranks = torch.empty(nq, dtype=torch.int32, device=scores.device)
    top1 = torch.empty(nq, dtype=torch.int32, device=scores.device)
    check(_lib.load().coopcap_retrieval_ranks(_p(scores), nc, nq, nc, _p(first), int(count), _p(ranks),
                                              _p(top1), _stream()))
    return ranks, top1


ST_CHUNK_ROWS = 4096   # rows of the upstream-gradient scratch per launch (78 MB bf16 at V1 = 9488)


def st_logit_grads(sp: SpeakerPass, demb16: torch.Tensor, w_emb16: torch.Tensor) -> torch.Tensor:
    """dz16 [n_steps*B, V1] for the straight-through samplers from the listener's embedding
    gradient demb16 [n_steps, B, E] (caption positions 1..n_steps)."""
    d = sp.dims
    dev = demb16.device
    assert demb16.dtype == torch.bfloat16 and demb16.is_contiguous()
    assert demb16.shape == (sp.n_steps, sp.B, d.E)
    chunk = max(1, min(sp.n_steps, ST_CHUNK_ROWS // max(sp.B, 1)))
    g_ws = torch.empty(chunk * sp.B, d.V1, dtype=torch.bfloat16, device=dev)
    dz16 = torch.empty(sp.n_steps * sp.B, d.V1, dtype=torch.bfloat16, device=dev)
    check(_lib.load().coopcap_st_backward(C.byref(sp.ctx), _p(demb16), _p(w_emb16), _p(g_ws),
                                          chunk, _p(dz16), _stream()))
    return dz16


def retain(p):
    """Register one more autograd node that will need the pass `p` in its backward."""
    p.users = getattr(p, "users", 0) + 1
    return p


def release(p):
    """Called at the end of a node's backward: when the last registered user is done the
    multi-GB workspaces of the pass are dropped right away (autograd only frees saved tensors,
    and a caller holding on to `loss` would otherwise keep a whole step of buffers alive)."""
    p.users = getattr(p, "users", 1) - 1
    if p.users <= 0 and not getattr(p, "pinned", False):   # tests pin passes they inspect later
        p.t.clear()
        p.keep.clear()
        p.ctx = None


def st_logit_grads_dense(sp: SpeakerPass, g: torch.Tensor) -> torch.Tensor:
    """Same from a dense upstream gradient g fp32 [n_steps, B, >=V1] (d loss / d one_hots)."""
    d = sp.dims
    g = _f32c(g)
    assert g.shape[0] == sp.n_steps and g.shape[1] == sp.B and g.shape[2] >= d.V1
    dz16 = torch.empty(sp.n_steps * sp.B, d.V1, dtype=torch.bfloat16, device=g.device)
    check(_lib.load().coopcap_st_backward_dense(C.byref(sp.ctx), _p(g), g.shape[2], _p(dz16),
                                                _stream()))
    return dz16


def logp_logit_grads(sp: SpeakerPass, tok: torch.Tensor, coef: torch.Tensor,
                     into: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dz16 for sum coef[t,b] * log_softmax(z[t,b])[tok[t,b]]; tok int64 / coef fp32 [n_steps, B].
    `into`: an existing dz16 of the same pass that the term is ADDED to."""
    d = sp.dims
    assert tok.dtype == torch.int64 and tok.is_contiguous() and tok.shape == (sp.n_steps, sp.B)
    coef = _f32c(coef)
    assert coef.shape == (sp.n_steps, sp.B)
    if into is not None:
        assert into.dtype == torch.bfloat16 and into.shape == (sp.n_steps * sp.B, d.V1)
        check(_lib.load().coopcap_logp_backward_acc(C.byref(sp.ctx), _p(tok), _p(coef), _p(into),
                                                    _stream()))
        return into
    dz16 = torch.empty(sp.n_steps * sp.B, d.V1, dtype=torch.bfloat16, device=tok.device)
    check(_lib.load().coopcap_logp_backward(C.byref(sp.ctx), _p(tok), _p(coef), _p(dz16), _stream()))
    return dz16


def grad_targets(P: Dict[str, torch.Tensor], names):
    """Where the fused backward nodes write each parameter's gradient.

    A parameter that belongs to a FlatAdam bucket, requires grad and has received no gradient yet
    in this backward pass gets the bucket's own gradient view: the kernels write there, the node
    sets `p.grad` to that view and returns None to autograd, so the gradient is never copied
    (no accumulation kernel, no gather into the bucket) and the optimizer can exchange that bucket
    range between ranks while the rest of backward still runs.  Everything else gets a fresh
    tensor that travels through autograd as usual.  Returns ({name: tensor}, {names written direct})."""
    G, direct, seen = {}, set(), set()
    for n in names:
        p = P[n]
        bucket = getattr(p, "_coopcap_bucket", None)
        view = bucket.grad_view(p) if bucket is not None else None
        if view is not None and p.requires_grad and p.grad is None and id(p) not in seen \
                and view.device == p.device:
            G[n] = view
            direct.add(n)
            seen.add(id(p))
        else:
            G[n] = torch.empty_like(p, dtype=torch.float32)
    return G, direct


def adopt_direct(P, G, direct):
    """After the kernels ran: the parameters written direct now own their bucket view as .grad."""
    for n in direct:
        P[n].grad = G[n]


def speaker_backward(sp: SpeakerPass, dz16: Optional[torch.Tensor], P: Dict[str, torch.Tensor],
                     ps_demb16: Optional[torch.Tensor] = None,
                     ps_w_emb16: Optional[torch.Tensor] = None,
                     ps_g_dense: Optional[torch.Tensor] = None,
                     out: Optional[Dict[str, torch.Tensor]] = None,
                     after_logit_layer=None) -> Dict[str, torch.Tensor]:
    """BPTT + prologue backward.  Returns {reference parameter name: fp32 gradient}; `out` gives the
    tensors to write them into (see grad_targets), else fresh ones are allocated.

    Partial-sampling passes take the upstream gradient of the emitted vectors instead of dz16:
    factored (`ps_demb16` bf16 [n_steps, B, E] + `ps_w_emb16`) or dense (`ps_g_dense` fp32
    [n_steps, B, ld >= V1], modified in place)."""
    d = sp.dims
    dev = sp.t["z16_all"].device
    is_ps = sp.ctx.mode in PS_MODES
    if is_ps:
        assert dz16 is None and ((ps_demb16 is None) != (ps_g_dense is None))
        dz16 = torch.empty(sp.n_steps * sp.B, d.V1, dtype=torch.bfloat16, device=dev)
    B, cap, NL = sp.B, sp.cap, sp.NL
    NS, XH = 5 * d.R + d.A, d.E + d.R
    f32 = dict(dtype=torch.float32, device=dev)
    bf = dict(dtype=torch.bfloat16, device=dev)
    ws = dict(
        d_out=torch.empty(max(cap, 2) * B, d.R, **f32), dscat16=torch.empty(cap * B, NS, **bf),
        d_att_res=torch.empty(cap * B, d.R, **f32), de=torch.empty(cap, NL, **f32),
        d_xh=torch.empty(cap * B, XH, **f32), dc=torch.empty(2, B, d.R, **f32),
        d_att_e=torch.empty(NL, d.R, **f32), d_p_att16=torch.empty(NL, d.A, **bf),
        d_pre16=torch.empty(NL, d.R, **bf))
    if out is None:
        G = {n: torch.empty_like(P[n], dtype=torch.float32) for n in SPEAKER_PARAM_NAMES
             if n not in ("core.h2h.bias",)}
    else:
        G = dict(out)
    G["embed.0.weight"].zero_()                     # accumulated into by atomics
    # the alpha_net bias shifts every score of a row equally: its gradient is identically 0
    G["core.attention.alpha_net.bias"].zero_()
    g = _lib.SpeakerGrads()
    g.dz16 = _p(dz16)
    if is_ps:
        ws["ps_dpre16"] = torch.zeros(sp.n_steps * sp.B, d.E, **bf)
        if ps_demb16 is not None:
            assert ps_demb16.dtype == torch.bfloat16 and ps_demb16.is_contiguous()
            assert ps_demb16.shape == (sp.n_steps, sp.B, d.E) and ps_w_emb16 is not None
            ws["ps_g"] = torch.empty(sp.B, d.V1, **f32)
            g.ps_demb16, g.ps_w_emb16, g.ps_ldg = _p(ps_demb16), _p(ps_w_emb16), d.V1
        else:
            _f32c(ps_g_dense)
            assert ps_g_dense.shape[:2] == (sp.n_steps, sp.B) and ps_g_dense.shape[2] >= d.V1
            assert ps_g_dense.shape[2] % 4 == 0, "dense upstream gradient needs a 16-byte row pitch"
            ws["ps_g"] = ps_g_dense
            g.ps_ldg = ps_g_dense.shape[2]
    for n, tsr in ws.items():
        setattr(g, n, _p(tsr))
    g.g_embed = _p(G["embed.0.weight"])
    g.g_w_att_embed, g.g_b_att_embed = _p(G["att_embed.0.weight"]), _p(G["att_embed.0.bias"])
    g.g_w_ctx2att, g.g_b_ctx2att = _p(G["ctx2att.weight"]), _p(G["ctx2att.bias"])
    g.g_w_i2h, g.g_w_h2h = _p(G["core.i2h.weight"]), _p(G["core.h2h.weight"])
    g.g_b_gates = _p(G["core.i2h.bias"])
    g.g_w_h2att, g.g_b_h2att = (_p(G["core.attention.h2att.weight"]),
                                _p(G["core.attention.h2att.bias"]))
    g.g_w_a2c, g.g_b_a2c = _p(G["core.a2c.weight"]), _p(G["core.a2c.bias"])
    g.g_w_logit, g.g_b_logit = _p(G["logit.weight"]), _p(G["logit.bias"])
    g.g_w_alpha = _p(G["core.attention.alpha_net.weight"])
    if after_logit_layer is not None and not is_ps:
        # vocabulary layer first; the caller may start exchanging its gradients between the ranks
        # while the BPTT loop runs (coopcap_speaker_grads.phase)
        g.phase = 1
        check(_lib.load().coopcap_speaker_decode_bwd(C.byref(sp.ctx), C.byref(g), _stream()))
        after_logit_layer(G)
        g.phase = 2
    check(_lib.load().coopcap_speaker_decode_bwd(C.byref(sp.ctx), C.byref(g), _stream()))
    if getattr(sp, "pinned", False):     # tests inspect the logit gradient of pinned passes
        sp.t["dz16"] = dz16
    if out is None:
        G["core.h2h.bias"] = G["core.i2h.bias"]     # both biases enter the same sum
    else:
        G["core.h2h.bias"].copy_(G["core.i2h.bias"])
    return G


# =============================================================================================
# listener
# =============================================================================================
LISTENER_PARAM_NAMES = [
    "img_enc.fc.weight", "img_enc.fc.bias", "txt_enc.embed.weight", "txt_enc.rnn.weight_ih_l0",
    "txt_enc.rnn.weight_hh_l0", "txt_enc.rnn.bias_ih_l0", "txt_enc.rnn.bias_hh_l0",
]


@dataclass
class ListenerDims:
    F: int
    M: int
    E: int
    V2: int

    @staticmethod
    def of(P) -> "ListenerDims":
        return ListenerDims(F=P["img_enc.fc.weight"].shape[1], M=P["img_enc.fc.weight"].shape[0],
                            E=P["txt_enc.embed.weight"].shape[1],
                            V2=P["txt_enc.embed.weight"].shape[0])


class PackedListener:
    def __init__(self):
        self.key = None
        self.buf = {}

    def get(self, P):
        key = (_weights_epoch,) + tuple((P[n].data_ptr(), P[n]._version) for n in LISTENER_PARAM_NAMES)
        if key == self.key:
            return self.buf
        d = ListenerDims.of(P)
        _need_cuda(*[P[n] for n in LISTENER_PARAM_NAMES])
        dev = P["img_enc.fc.weight"].device
        if not self.buf or self.buf["dims"] != d:
            bf = dict(dtype=torch.bfloat16, device=dev)
            self.buf = dict(dims=d, w_img16=torch.empty(d.M, d.F, **bf),
                            w_ih16=torch.empty(3 * d.M, d.E, **bf),
                            w_hh16=_arena_alloc(dev, (3 * d.M, d.M)),
                            w_emb16=torch.empty(d.V2, d.E, **bf))
        b = self.buf
        a = _lib.ListenerPack()
        a.F, a.M, a.E, a.V2 = d.F, d.M, d.E, d.V2
        a.w_img = _p(_f32c(P["img_enc.fc.weight"].detach()))
        a.w_ih = _p(_f32c(P["txt_enc.rnn.weight_ih_l0"].detach()))
        a.w_hh = _p(_f32c(P["txt_enc.rnn.weight_hh_l0"].detach()))
        a.w_emb = _p(_f32c(P["txt_enc.embed.weight"].detach()))
        a.w_img16, a.w_ih16, a.w_hh16, a.w_emb16 = (_p(b["w_img16"]), _p(b["w_ih16"]),
                                                    _p(b["w_hh16"]), _p(b["w_emb16"]))
        _arena_persist(dev)
        check(_lib.load().coopcap_listener_pack_weights(C.byref(a), _stream()))
        self.key = key
        return b


@dataclass
class ListenerPass:
    ctx: object = None
    dims: ListenerDims = None
    B: int = 0
    S: int = 0
    t: Dict[str, torch.Tensor] = field(default_factory=dict)
    keep: list = field(default_factory=list)


ONLY = {"off": 0, "image": 1, "caption": 2}


POOL = {"last": 0, "": 0, "mean": 1, "max": 2}


def listener_forward(P, packed, fc_feats, tok_sb, lens, *, margin=0.2, only_one_retrieval="off",
                     no_imgnorm=False, emb16: Optional[torch.Tensor] = None, pool_type="last",
                     use_abs=False, max_violation=True) -> ListenerPass:
    """tok_sb int64 [S, B] time-major ids, lens int32 [B].  Results in .t['loss'] ([1]) and
    .t['loss_rows'] ([B]).  `emb16` bf16 [S, B, E]: caption embeddings computed by the caller
    (dense caption vectors); tok_sb is then None."""
    _need_cuda(fc_feats, tok_sb, lens, emb16)
    d: ListenerDims = packed["dims"]
    fc_feats = _f32c(fc_feats)
    assert lens.dtype == torch.int32 and lens.is_contiguous()
    if emb16 is None:
        assert tok_sb.dtype == torch.int64 and tok_sb.is_contiguous()
        S, B = tok_sb.shape
    else:
        assert emb16.dtype == torch.bfloat16 and emb16.is_contiguous() and emb16.shape[2] == d.E
        S, B = emb16.shape[:2]
    dev = fc_feats.device
    f32 = dict(dtype=torch.float32, device=dev)
    bf = dict(dtype=torch.bfloat16, device=dev)
    T = dict(
        fc16=torch.empty(B, d.F, **bf), img_pre=torch.empty(B, d.M, **f32),
        im=torch.empty(B, d.M, **f32), emb16=emb16 if emb16 is not None else torch.empty(S, B, d.E, **bf),
        gi_all=torch.empty(S, B, 3 * d.M, **f32), gh=torch.empty(B, 3 * d.M, **f32),
        gates=torch.empty(S, B, 4 * d.M, **f32), h32=torch.empty(S + 1, B, d.M, **f32),
        h16=torch.empty(S + 1, B, d.M, **bf), cap=torch.empty(B, d.M, **f32),
        scores=torch.empty(B, B, **f32), cost_s=torch.empty(B, **f32),
        cost_im=torch.empty(B, **f32), arg_s=torch.empty(B, dtype=torch.int32, device=dev),
        arg_im=torch.empty(B, dtype=torch.int32, device=dev), loss_rows=torch.empty(B, **f32),
        loss=torch.empty(1, **f32))
    c = _lib.Listener()
    c.B, c.S, c.F, c.M, c.E, c.V2 = B, S, d.F, d.M, d.E, d.V2
    c.margin, c.only_one_retrieval, c.no_imgnorm = float(margin), ONLY[only_one_retrieval], int(no_imgnorm)
    c.fc_feats, c.tok, c.len = _p(fc_feats), _p(tok_sb), _p(lens)
    c.emb_given = int(emb16 is not None)
    # non-default listener options (VSEFCModel.py:115-126 pooling, :50-52/:137-139 abs, :190-193 hinge)
    c.pool_type, c.use_abs, c.sum_violation = POOL[pool_type], int(bool(use_abs)), int(not max_violation)
    if c.pool_type:
        T["cap_pre"] = torch.empty(B, d.M, **f32)
        T["pool_arg"] = torch.empty(B, d.M, dtype=torch.int32, device=dev)
    c.w_emb = _p(_f32c(P["txt_enc.embed.weight"].detach()))
    c.b_img = _p(_f32c(P["img_enc.fc.bias"].detach()))
    c.b_ih = _p(_f32c(P["txt_enc.rnn.bias_ih_l0"].detach()))
    c.b_hh = _p(_f32c(P["txt_enc.rnn.bias_hh_l0"].detach()))
    c.w_img16, c.w_ih16, c.w_hh16 = _p(packed["w_img16"]), _p(packed["w_ih16"]), _p(packed["w_hh16"])
    for n, tsr in T.items():
        setattr(c, n, _p(tsr))
    lp = ListenerPass(ctx=c, dims=d, B=B, S=S, t=T, keep=[fc_feats, tok_sb, lens, packed, P])
    check(_lib.load().coopcap_listener_fwd(C.byref(c), _stream()))
    return lp


def listener_backward(lp: ListenerPass, P, *, g_loss=None, g_rows=None, need_param_grads=True,
                      out=None):
    """Returns (grads {name: tensor} or None, demb16 [S, B, E] bf16); `out` gives the tensors the
    parameter gradients are written into (see grad_targets)."""
    d = lp.dims
    B, S = lp.B, lp.S
    dev = lp.t["im"].device
    f32 = dict(dtype=torch.float32, device=dev)
    bf = dict(dtype=torch.bfloat16, device=dev)
    ws = dict(d_im=torch.empty(B, d.M, **f32), d_cap=torch.empty(B, d.M, **f32),
              dh=torch.empty(B, d.M, **f32), d_img_pre16=torch.empty(B, d.M, **bf),
              d_gi16=torch.empty(S, B, 3 * d.M, **bf), d_gh16=torch.empty(S, B, 3 * d.M, **bf),
              demb16=torch.empty(S, B, d.E, **bf))
    if lp.ctx.pool_type:
        ws["d_pool"] = torch.empty(B, d.M, **f32)
    if lp.ctx.sum_violation:
        ws["d_scores"] = torch.empty(B, B, **f32)
    g = _lib.ListenerGrads()
    if g_loss is not None:
        g_loss = _f32c(g_loss.reshape(1))
    if g_rows is not None:
        g_rows = _f32c(g_rows)
        assert g_rows.shape == (B,)
    g.g_loss, g.g_rows = _p(g_loss), _p(g_rows)
    g.need_param_grads = int(need_param_grads)
    for n, tsr in ws.items():
        setattr(g, n, _p(tsr))
    G = None
    if need_param_grads:
        G = dict(out) if out is not None else \
            {n: torch.empty_like(P[n], dtype=torch.float32) for n in LISTENER_PARAM_NAMES}
        G["txt_enc.embed.weight"].zero_()           # accumulated into by atomics
        g.g_w_img, g.g_b_img = _p(G["img_enc.fc.weight"]), _p(G["img_enc.fc.bias"])
        g.g_w_emb = _p(G["txt_enc.embed.weight"])
        g.g_w_ih, g.g_w_hh = _p(G["txt_enc.rnn.weight_ih_l0"]), _p(G["txt_enc.rnn.weight_hh_l0"])
        g.g_b_ih, g.g_b_hh = _p(G["txt_enc.rnn.bias_ih_l0"]), _p(G["txt_enc.rnn.bias_hh_l0"])
    check(_lib.load().coopcap_listener_bwd(C.byref(lp.ctx), C.byref(g), _stream()))
    return G, ws["demb16"]


def caption_embed_dense(soft16: torch.Tensor, n: int, P, packed, bos_id: int) -> torch.Tensor:
    """emb16 bf16 [n+1, B, E] of the caption [BOS, v_0 .. v_{n-1}] given dense vectors soft16
    bf16 [>= n, B, V1] (VSEFCModel.py:102-104 with AlternatingJointModel.py:356-370)."""
    d: ListenerDims = packed["dims"]
    _need_cuda(soft16)
    assert soft16.dtype == torch.bfloat16 and soft16.is_contiguous() and soft16.shape[0] >= n
    B, V1 = soft16.shape[1:]
    emb16 = torch.empty(n + 1, B, d.E, dtype=torch.bfloat16, device=soft16.device)
    check(_lib.load().coopcap_caption_embed_dense(
        _p(soft16), _p(_f32c(P["txt_enc.embed.weight"].detach())), _p(packed["w_emb16"]), n, B, V1,
        d.E, int(bos_id), _p(emb16), _stream()))
    return emb16


def caption_embed_dense_bwd(soft16: torch.Tensor, demb16: torch.Tensor, n: int, bos_id: int,
                            g_w_emb: torch.Tensor) -> None:
    """g_w_emb (fp32 [V2, E], zero-initialised) <- embedding weight gradient of caption_embed_dense."""
    B, V1 = soft16.shape[1:]
    assert demb16.dtype == torch.bfloat16 and demb16.is_contiguous() and demb16.shape[0] == n + 1
    check(_lib.load().coopcap_caption_embed_dense_bwd(_p(soft16), _p(demb16), n, B, V1,
                                                      demb16.shape[2], int(bos_id),
                                                      _p(_f32c(g_w_emb)), _stream()))


def clamp_adam_(param, grad, exp_avg, exp_avg_sq, *, step, lr, grad_scale=1.0, clip=0.0,
                beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0):
    """In-place fused (grad * scale) -> clamp -> Adam over flat fp32 buffers."""
    _need_cuda(param, grad, exp_avg, exp_avg_sq)
    for t in (param, grad, exp_avg, exp_avg_sq):
        _f32c(t)
    n = param.numel()
    assert grad.numel() == n and exp_avg.numel() == n and exp_avg_sq.numel() == n
    check(_lib.load().coopcap_clamp_adam(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), n,
                                         float(grad_scale), float(clip), float(lr), float(beta1),
                                         float(beta2), float(eps), float(weight_decay), int(step),
                                         _stream()))
