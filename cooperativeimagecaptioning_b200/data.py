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


def row_order(lens: torch.Tensor) -> torch.Tensor:
    """Row ids by decreasing region count (int32, pinned): the attention kernels deal rows to SMs in
    this order; computing it here saves a device-side sort at the head of every pass."""
    o = torch.argsort(lens, descending=True, stable=True).to(torch.int32)
    return o.pin_memory() if torch.cuda.is_available() else o


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
            am._coopcap_order = row_order(lens).to(device, non_blocking=True)
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


class HostPacker:
    """Pipelined upload of batches: worker threads of libcoopcap pack the valid regions of the
    host `att_feats` into a pinned bf16 staging buffer (`coopcap_host_pack_start`) while the caller
    keeps enqueueing the current step; `finish` then moves the packed operand with one DMA copy.
    Compared with `upload_batch(zero_copy=True)` half the bytes cross PCIe (bf16 instead of fp32)
    and no SM is taken from the compute stream.  Same rounding as the device-side pack, so the
    operand is bit-identical.

        job = packer.start(fc, att, att_masks, labels, masks)      # returns immediately
        ...                                                         # enqueue other work
        batch = packer.finish(job, stream=copy_stream)              # CUDA tensors, async on `stream`
    """

    def __init__(self, device, nbuf: int = 3, threads: int = 0):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CoopcapError("HostPacker targets a CUDA device (there is no CPU path)")
        self.threads = int(threads)
        self._slots = [dict(buf=None, ev=None) for _ in range(max(2, nbuf))]
        self._next = 0
        self.last_bytes = 0

    def start(self, fc_feats, att_feats, att_masks, labels, masks):
        lib = _lib.load()
        B, L, D = att_feats.shape
        if att_feats.dtype != torch.float32 or att_feats.device.type != "cpu":
            raise _lib.CoopcapError("HostPacker.start takes fp32 host tensors")
        att_feats = att_feats.contiguous()
        off = None
        NL = B * L
        if att_masks is not None:
            lens = (att_masks > 0).sum(1).to(torch.int32)
            off = torch.zeros(B + 1, dtype=torch.int32)
            off[1:] = torch.cumsum(lens, 0)
            NL = int(off[-1])
        slot = self._slots[self._next]
        self._next = (self._next + 1) % len(self._slots)
        if slot["ev"] is not None:
            slot["ev"].synchronize()          # the DMA that last read this staging buffer is done
        if slot["buf"] is None or slot["buf"].numel() < NL * D:
            slot["buf"] = torch.empty(B * L * D, dtype=torch.bfloat16).pin_memory()
        job = lib.coopcap_host_pack_start(
            C.c_void_p(att_feats.data_ptr()), C.c_void_p(off.data_ptr()) if off is not None else None,
            B, L, D, C.c_void_p(slot["buf"].data_ptr()), self.threads)
        check(job if job < 0 else 0)
        return dict(job=job, slot=slot, off=off, NL=NL, shape=(B, L, D), fc=fc_feats, labels=labels,
                    masks=masks, att_masks=att_masks, src=att_feats)

    def finish(self, j, stream=None):
        lib = _lib.load()
        check(lib.coopcap_host_pack_wait(j["job"]))
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        B, L, D = j["shape"]
        NL = j["NL"]
        nb = lambda t: t.numel() * t.element_size()
        with torch.cuda.stream(st):
            fc = j["fc"].to(self.device, non_blocking=True)
            lab = j["labels"].to(self.device, non_blocking=True)
            msk = j["masks"].to(self.device, non_blocking=True)
            att16 = torch.empty(NL, D, dtype=torch.bfloat16, device=self.device)
            att16.copy_(j["slot"]["buf"][: NL * D].view(NL, D), non_blocking=True)
            moved = nb(j["fc"]) + nb(j["labels"]) + nb(j["masks"]) + NL * D * 2
            if j["att_masks"] is not None:
                am = j["att_masks"].to(self.device, non_blocking=True)
                off_d = j["off"].to(self.device, non_blocking=True)
                am._coopcap_order = row_order(j["off"][1:] - j["off"][:-1]).to(self.device, non_blocking=True)
                moved += nb(j["att_masks"]) + 4 * (B + 1)
            else:
                # fixed region count: an all-ones mask carries the packed operand
                am = torch.ones(B, L, device=self.device)
                off_d = torch.arange(0, (B + 1) * L, L, dtype=torch.int32, device=self.device)
            am._coopcap_off = (off_d, NL)
            am._coopcap_att16 = att16
            att = torch.zeros(1, 1, 1, device=self.device).expand(B, L, D)      # shape carrier only
            ev = torch.cuda.Event()
            ev.record(st)
        j["slot"]["ev"] = ev
        self.last_bytes = moved
        return fc, att, am, lab, msk


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
        order = getattr(t, "_coopcap_order", None)
        if order is not None:
            order.record_stream(stream)
        a16 = getattr(t, "_coopcap_att16", None)
        if a16 is not None:
            a16.record_stream(stream)
