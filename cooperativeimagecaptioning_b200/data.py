"""Host -> device staging of one batch: the B200 counterpart of the reference's `load_data`
(train.py:162-178) / `utils.var_wrapper(...).cuda()` (misc/utils.py:72-87).

The reference copies the zero-padded `att_feats [rows, Lmax, 2048]` fp32 tensor whole.  Here only
the valid regions of every row cross PCIe (`coopcap_h2d_ragged_rows`), and the packed-region
offsets the kernels need are derived from the host-side mask, so no device synchronisation is
needed later.  Padded tails of the device `att_feats` are unspecified (the kernels work on the
packed valid regions only).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import check


def upload_batch(fc_feats: torch.Tensor, att_feats: torch.Tensor, att_masks: Optional[torch.Tensor],
                 labels: torch.Tensor, masks: torch.Tensor, device, *, stream=None,
                 zero_copy: bool = True, ctas: int = 64):
    """CPU (ideally pinned) tensors -> CUDA tensors, asynchronously on `stream` (default: current).

    Returns (fc_feats, att_feats, att_masks, labels, masks) on `device`; `att_masks` carries the
    packed-region offsets (`_coopcap_off`) consumed by Att2in2Model.  Bytes moved are reported in
    `upload_batch.last_bytes`.

    zero_copy (needs a pinned `att_feats` and `att_masks`): a small persistent kernel reads the
    valid regions straight from the pinned host buffer and writes the packed bf16 operand
    (`coopcap_pack_att_from_host`); the returned `att_feats` is then a zero-stride placeholder of
    the right shape and the packed data rides on `att_masks._coopcap_att16`."""
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.CoopcapError("upload_batch targets a CUDA device (there is no CPU path)")
    st = stream if stream is not None else torch.cuda.current_stream(device)
    lib = _lib.load()
    with torch.cuda.stream(st):
        nb = lambda t: t.numel() * t.element_size()
        fc = fc_feats.to(device, non_blocking=True)
        lab = labels.to(device, non_blocking=True)
        msk = masks.to(device, non_blocking=True)
        moved = nb(fc_feats) + nb(labels) + nb(masks)
        B, L, D = att_feats.shape
        att_feats = att_feats.contiguous()
        if att_masks is None:
            att = att_feats.to(device, non_blocking=True)
            moved += nb(att_feats)
            am = None
        else:
            lens = (att_masks > 0).sum(1).to(torch.int32).contiguous()        # host side
            off = torch.zeros(B + 1, dtype=torch.int32)
            off[1:] = torch.cumsum(lens, 0)
            esz = att_feats.element_size()
            NL = int(off[-1])
            off_d = off.to(device, non_blocking=True)
            am = att_masks.to(device, non_blocking=True)
            am._coopcap_off = (off_d, NL)
            if zero_copy and att_feats.is_pinned() and att_feats.dtype == torch.float32:
                att16 = torch.empty(NL, D, dtype=torch.bfloat16, device=device)
                check(lib.coopcap_pack_att_from_host(
                    C.c_void_p(att_feats.data_ptr()), C.c_void_p(off_d.data_ptr()), B, L, D, NL,
                    C.c_void_p(att16.data_ptr()), int(ctas), C.c_void_p(st.cuda_stream)))
                am._coopcap_att16 = att16
                am._coopcap_src = att_feats          # keep the pinned source alive until consumed
                att = torch.zeros(1, 1, 1, device=device).expand(B, L, D)   # shape carrier only
            else:
                att = torch.empty(B, L, D, dtype=att_feats.dtype, device=device)
                check(lib.coopcap_h2d_ragged_rows(
                    C.c_void_p(att.data_ptr()), C.c_void_p(att_feats.data_ptr()),
                    C.c_void_p(lens.data_ptr()), B, L * D * esz, D * esz, C.c_void_p(st.cuda_stream)))
            moved += NL * D * esz + nb(att_masks) + 4 * (B + 1)
    upload_batch.last_bytes = moved
    return fc, att, am, lab, msk


upload_batch.last_bytes = 0


def record_stream(batch, stream):
    """Tell the caching allocator that `stream` uses the tensors of an uploaded batch (they were
    allocated on the upload stream), including the packed side buffers."""
    for t in batch:
        if t is None:
            continue
        t.record_stream(stream)
        off = getattr(t, "_coopcap_off", None)
        if off is not None:
            off[0].record_stream(stream)
        a16 = getattr(t, "_coopcap_att16", None)
        if a16 is not None:
            a16.record_stream(stream)
