"""TEST INFRASTRUCTURE ONLY.  Seeded synthetic parameters, inputs and noise.

Everything is drawn on the CPU from an explicit torch.Generator so that the development
container (where the real reference can be imported) and the GPU box (where it cannot) build
bit-identical tensors.  Parameter distributions follow the reference's initialisers
(torch defaults for the speaker, AttModel.py:74-94; VSEFCModel.py:32-38,80-81 for the listener).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from .speaker import SpeakerNoise


@dataclass
class Dims:
    vocab_size: int = 9487
    seq_length: int = 16
    rnn_size: int = 512
    input_encoding_size: int = 512
    att_hid_size: int = 512
    att_feat_size: int = 2048
    fc_feat_size: int = 2048
    vse_embed_size: int = 1024


TINY = Dims(vocab_size=23, seq_length=6, rnn_size=16, input_encoding_size=16, att_hid_size=16,
            att_feat_size=12, fc_feat_size=12, vse_embed_size=20)


def _uniform(gen, shape, bound):
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


def _linear_init(gen, out_f, in_f):
    b = 1.0 / math.sqrt(in_f)   # kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(fan_in), 1/sqrt(fan_in))
    return _uniform(gen, (out_f, in_f), b), _uniform(gen, (out_f,), b)


def speaker_params(d: Dims, seed: int = 0, eos_bias: float = 0.0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    R, E, A, V = d.rnn_size, d.input_encoding_size, d.att_hid_size, d.vocab_size
    P = {"embed.0.weight": torch.randn(V + 2, E, generator=g)}
    for name, (o, i) in {
        "att_embed.0": (R, d.att_feat_size), "logit": (V + 1, R), "ctx2att": (A, R),
        "core.a2c": (2 * R, R), "core.i2h": (5 * R, E), "core.h2h": (5 * R, R),
        "core.attention.h2att": (A, R), "core.attention.alpha_net": (1, A),
    }.items():
        P[name + ".weight"], P[name + ".bias"] = _linear_init(g, o, i)
    if eos_bias:
        P["logit.bias"][0] += eos_bias   # makes rows finish at realistic lengths
    return P


def listener_params(d: Dims, seed: int = 1) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    M, E, F = d.vse_embed_size, d.input_encoding_size, d.fc_feat_size
    r = math.sqrt(6.0) / math.sqrt(F + M)
    k = 1.0 / math.sqrt(M)
    return {
        "img_enc.fc.weight": _uniform(g, (M, F), r),
        "img_enc.fc.bias": torch.zeros(M),
        "txt_enc.embed.weight": _uniform(g, (d.vocab_size + 2, E), 0.1),
        "txt_enc.rnn.weight_ih_l0": _uniform(g, (3 * M, E), k),
        "txt_enc.rnn.weight_hh_l0": _uniform(g, (3 * M, M), k),
        "txt_enc.rnn.bias_ih_l0": _uniform(g, (3 * M,), k),
        "txt_enc.rnn.bias_hh_l0": _uniform(g, (3 * M,), k),
    }


@dataclass
class Batch:
    fc_feats: torch.Tensor
    att_feats: torch.Tensor
    att_masks: Optional[torch.Tensor]
    labels: torch.Tensor
    masks: torch.Tensor


def make_batch(d: Dims, rows: int, regions: int, seed: int, varlen: bool = False,
               min_regions: int = 10, repeat: int = 1) -> Batch:
    """Loader-shaped batch (dataloader.py:194-237): labels [rows, T+2] with col 0 = 0, the caption in
    cols 1..len, zeros after; masks[:, :len+2] = 1.  `repeat` replicates each image's features
    (seq_per_img).  varlen: per-image region counts in [min_regions, regions], at least one row
    at `regions` (the loader pads to the batch max), att_masks float [rows, regions]."""
    g = torch.Generator().manual_seed(seed)
    n_img = rows // repeat
    fc = torch.randn(n_img, d.fc_feat_size, generator=g).repeat_interleave(repeat, 0)
    att = torch.randn(n_img, regions, d.att_feat_size, generator=g).repeat_interleave(repeat, 0)
    att_masks = None
    if varlen:
        lens = torch.randint(min(min_regions, regions), regions + 1, (n_img,), generator=g)
        lens[int(torch.randint(0, n_img, (1,), generator=g))] = regions
        lens = lens.repeat_interleave(repeat, 0)
        att_masks = (torch.arange(regions)[None, :] < lens[:, None]).float()
        att = att * att_masks[:, :, None]          # the loader zero-pads features
    T = d.seq_length
    lo = min(6, T)
    clen = torch.randint(lo, T + 1, (rows,), generator=g)
    labels = torch.zeros(rows, T + 2, dtype=torch.long)
    words = torch.randint(1, d.vocab_size + 1, (rows, T), generator=g)
    pos = torch.arange(T)[None, :]
    labels[:, 1:T + 1] = torch.where(pos < clen[:, None], words, torch.zeros_like(words))
    masks = (torch.arange(T + 2)[None, :] < (clen + 2)[:, None]).float()
    return Batch(fc, att, att_masks, labels, masks)


def make_noise(d: Dims, rows: int, regions: int, seed: int, *, dropout: bool = True,
               gumbel: bool = False, multinomial: bool = False, partial: bool = False,
               steps: Optional[int] = None, sched: bool = False) -> SpeakerNoise:
    g = torch.Generator().manual_seed(seed)
    T = d.seq_length
    steps = steps if steps is not None else T + 1
    n = SpeakerNoise()
    if dropout:
        n.drop_att = (torch.rand(rows, regions, d.rnn_size, generator=g) < 0.5).float()
        n.drop_embed = (torch.rand(steps, rows, d.input_encoding_size, generator=g) < 0.5).float()
        n.drop_core = (torch.rand(steps, rows, d.rnn_size, generator=g) < 0.5).float()
    if gumbel:
        n.U = torch.rand(T, rows, d.vocab_size + 1, generator=g)
    if multinomial:
        n.E = torch.empty(T, rows, d.vocab_size + 1).exponential_(generator=g)
    if partial:
        n.part_u = torch.rand(T, rows, generator=g)
    if sched:                      # drawn last: the other streams keep their values
        n.ss_u = torch.rand(T + 1, rows, generator=g)
    return n
