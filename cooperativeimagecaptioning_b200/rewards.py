"""CIDEr-D self-critical reward, drop-in for the reference's `misc/rewards.py` with the scorer on
the device (csrc/cider.cu, `coopcap_cider_reward`).

Same module surface as the reference (misc/rewards.py:19-71): `init_scorer(cached_tokens)`,
`array_to_str`, `get_self_critical_reward(data, gen_result, greedy_res, return_gen_scores=False)`
returning numpy arrays.  The joint model does not call that host-returning function on its hot
path: `reward_on_device` leaves reward, REINFORCE coefficients and the logged statistics in HBM
(no device->host copy, no Python n-gram dictionaries).

`cached_tokens` = "corpus" (opts.py:27 default): document frequencies come from the batch itself.
Anything else names `data/<cached_tokens>.p`, the pickle written by preprocess/prepro_ngrams.py:
119-122 ({'document_frequency': {tuple of id strings: count}, 'ref_len': images}); it is turned
into an exact-key open-addressing table once and kept in HBM.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib
from ._lib import check

SLOTS = 64
MAX_WORDS = 16


@dataclass
class _Table:
    keys: torch.Tensor          # int64 view of the uint64 keys [cap]
    val: torch.Tensor           # fp32 [cap]
    log_ref_len: float


class DeviceCiderD:
    """State of the scorer (the reference's module-global `CiderD_scorer`)."""

    def __init__(self, df: str = "corpus"):
        self.df_mode = df
        self.host_table = None            # (keys uint64 [cap], val fp32 [cap], log_ref_len)
        self._dev: Dict[torch.device, _Table] = {}
        if df != "corpus":
            with open(os.path.join("data", df + ".p"), "rb") as f:     # ciderD_scorer.py:69-73
                blob = pickle.load(f, encoding="latin1")
            self.host_table = build_table(blob["document_frequency"], float(blob["ref_len"]))

    @staticmethod
    def from_table(doc_freq: dict, ref_len: float) -> "DeviceCiderD":
        s = DeviceCiderD("corpus")
        s.df_mode = "table"
        s.host_table = build_table(doc_freq, ref_len)
        return s

    def table_on(self, device) -> Optional[_Table]:
        if self.host_table is None:
            return None
        t = self._dev.get(device)
        if t is None:
            k, v, lr = self.host_table
            t = _Table(torch.from_numpy(k.view(np.int64)).to(device), torch.from_numpy(v).to(device), lr)
            self._dev[device] = t
        return t


def pack_key(gram) -> int:
    """The device's exact n-gram key: 16 bits per id, id + 1 so that 0 means 'absent'."""
    key = 0
    for i, w in enumerate(gram):
        w = int(w)
        if not 0 <= w < 65535:
            raise _lib.CoopcapError(f"CIDEr n-gram id {w} does not fit 16 bits")
        key |= (w + 1) << (16 * i)
    return key


def build_table(doc_freq: dict, ref_len: float):
    """{n-gram tuple (ids or id strings): df} -> open-addressing arrays probed like the device."""
    lib = _lib.load()
    items = [(pack_key(g), float(v)) for g, v in doc_freq.items() if 1 <= len(g) <= 4]
    cap = 1 << max(4, int(np.ceil(np.log2(max(2 * len(items), 2)))))
    keys = np.zeros(cap, np.uint64)
    val = np.zeros(cap, np.float32)
    mask = cap - 1
    for k, v in items:
        h = int(lib.coopcap_cider_hash(C.c_uint64(k))) & 0xFFFFFFFF & mask
        while keys[h] != 0 and int(keys[h]) != k:
            h = (h + 1) & mask
        keys[h] = k
        val[h] = v
    return keys, val, float(np.log(float(ref_len)))


CiderD_scorer: Optional[DeviceCiderD] = None


def init_scorer(cached_tokens):                                            # rewards.py:22-24
    global CiderD_scorer
    CiderD_scorer = CiderD_scorer or DeviceCiderD(df=cached_tokens)


def array_to_str(arr):                                                     # rewards.py:26-32
    out = []
    for w in arr:
        out.append(str(int(w)))
        if int(w) == 0:
            break
    return " ".join(out)


@dataclass
class StagedGts:
    refs: torch.Tensor       # int64 [n_ref, W]
    ref_off: torch.Tensor    # int32 [n_img + 1]
    row_img: torch.Tensor    # int32 [B]
    n_img: int
    n_ref: int


def stage_gts(gts, B: int, device) -> StagedGts:
    """data['gts'] (list over images of int arrays [n_captions, W], dataloader.py:199-203,239) ->
    device arrays; row b belongs to image b // (B // n_images) (rewards.py:37,53)."""
    n_img = len(gts)
    if n_img == 0 or B % n_img:
        raise _lib.CoopcapError(f"CIDEr reward: {B} rows do not split over {n_img} images")
    counts = [len(g) for g in gts]
    if min(counts) == 0:
        raise _lib.CoopcapError("CIDEr reward: an image has no ground-truth caption")
    flat = np.concatenate([np.asarray(g, dtype=np.int64).reshape(len(g), -1) for g in gts], 0)
    if flat.shape[1] > MAX_WORDS:
        raise _lib.CoopcapError(f"CIDEr reward: captions wider than {MAX_WORDS} ids are not supported")
    off = np.zeros(n_img + 1, np.int32)
    off[1:] = np.cumsum(counts)
    row_img = (np.arange(B) // (B // n_img)).astype(np.int32)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().to(device, non_blocking=True)
    return StagedGts(pin(flat), pin(off), pin(row_img), n_img, int(flat.shape[0]))


@dataclass
class CiderResult:
    reward: torch.Tensor     # fp32 [B]
    coef: torch.Tensor       # fp32 [T, B]
    scores: torch.Tensor     # fp64 [n_sets * B]
    stats: torch.Tensor      # fp64 [4]: mean reward, mean greedy score, sum(mask), mean sampled score


def reward_on_device(scorer: DeviceCiderD, gts: StagedGts, hyp0: torch.Tensor,
                     hyp1: Optional[torch.Tensor], differenced: bool = True) -> CiderResult:
    """hyp0 / hyp1: int64 [T, B] time-major sampled / greedy ids (finished rows hold 0)."""
    if not hyp0.is_cuda:
        raise _lib.CoopcapError("the CIDEr-D scorer runs on CUDA only (no CPU path)")
    assert hyp0.dtype == torch.int64 and hyp0.is_contiguous()
    T, B = hyp0.shape
    if T > MAX_WORDS:
        raise _lib.CoopcapError(f"CIDEr reward: captions longer than {MAX_WORDS} ids are not supported")
    n_sets = 1 if hyp1 is None else 2
    if hyp1 is not None:
        assert hyp1.dtype == torch.int64 and hyp1.is_contiguous() and hyp1.shape == hyp0.shape
    dev = hyp0.device
    Ctot = n_sets * B + gts.n_ref
    i32 = dict(dtype=torch.int32, device=dev)
    f64 = dict(dtype=torch.float64, device=dev)
    ws = dict(ng_key=torch.empty(Ctot, SLOTS, dtype=torch.int64, device=dev),
              ng_cnt=torch.empty(Ctot, SLOTS, **i32), ng_n=torch.empty(Ctot, 4, **i32),
              ng_len=torch.empty(Ctot, **i32), ng_w=torch.empty(Ctot, SLOTS, **f64),
              ng_norm=torch.empty(Ctot, 4, **f64), img_rows=torch.empty(gts.n_img, **i32))
    out = CiderResult(reward=torch.empty(B, dtype=torch.float32, device=dev),
                      coef=torch.empty(T, B, dtype=torch.float32, device=dev),
                      scores=torch.empty(n_sets * B, **f64), stats=torch.empty(4, **f64))
    c = _lib.Cider()
    c.B, c.n_sets, c.T, c.W, c.n_img, c.n_ref = B, n_sets, T, gts.refs.shape[1], gts.n_img, gts.n_ref
    c.differenced = int(bool(differenced))
    table = scorer.table_on(dev)
    if table is None:
        cap = 1 << int(np.ceil(np.log2(2 * 58 * gts.n_ref)))
        keys = torch.empty(cap, dtype=torch.int64, device=dev)
        val = torch.empty(cap, dtype=torch.float32, device=dev)
        c.corpus, c.log_ref_len = 1, 0.0
    else:
        keys, val = table.keys, table.val
        c.corpus, c.log_ref_len = 0, table.log_ref_len
    c.df_keys, c.df_val, c.df_cap = keys.data_ptr(), val.data_ptr(), keys.numel()
    c.hyp0, c.hyp1 = hyp0.data_ptr(), (None if hyp1 is None else hyp1.data_ptr())
    c.refs, c.ref_off, c.row_img = gts.refs.data_ptr(), gts.ref_off.data_ptr(), gts.row_img.data_ptr()
    for n, t in ws.items():
        setattr(c, n, t.data_ptr())
    c.scores, c.reward, c.coef, c.stats = (out.scores.data_ptr(), out.reward.data_ptr(),
                                           out.coef.data_ptr(), out.stats.data_ptr())
    check(_lib.load().coopcap_cider_reward(C.byref(c), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out


def _time_major(x: torch.Tensor) -> torch.Tensor:
    return x.detach().long().t().contiguous()


def get_self_critical_reward(data, gen_result, greedy_res, return_gen_scores=False):
    """The reference's host-returning call (rewards.py:34-71) on [B, n] id tensors: numpy
    `(scores, cider_greedy)` or `(cider_gen, scores, cider_greedy)`.  One device->host copy."""
    global CiderD_scorer
    if CiderD_scorer is None:
        raise _lib.CoopcapError("call rewards.init_scorer(cached_tokens) first (train.py:483)")
    gts = stage_gts(data["gts"], gen_result.size(0), gen_result.device)
    n = max(gen_result.size(1), greedy_res.size(1))
    pad = lambda x: torch.nn.functional.pad(x, (0, n - x.size(1)))
    r = reward_on_device(CiderD_scorer, gts, _time_major(pad(gen_result)), _time_major(pad(greedy_res)))
    s = r.scores.cpu().numpy()
    B = gen_result.size(0)
    scores = s[:B] - s[B:]
    if not return_gen_scores:
        return scores, s[B:].mean()
    return s[:B], scores, s[B:].mean()
