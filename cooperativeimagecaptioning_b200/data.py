"""Host -> device staging of one batch: the B200 counterpart of the reference's `load_data`
(train.py:162-178) / `utils.var_wrapper(...).cuda()` (misc/utils.py:72-87).

Two ways in: `FeatureStore` keeps the whole feature set in HBM and a step ships only image indices
(the fast path); `upload_batch` / `HostPacker` move one loader-shaped batch of host features.

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
                # the reader kernel of the NEXT batch shares the SMs with this step's weight re-pack:
                # separate cast launches interleave with it better than the fused one (10.6 vs 11.8 ms)
                lib.coopcap_set_cast_multi(0)
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


class FeatureStore:
    """All images' features resident in HBM; a step ships image indices instead of features.

    The reference's loader fetches every image's bottom-up features from disk by index, pads them
    and `load_data` copies 839 MB of fp32 per 1024-row step to the GPU (dataloader.py:137-160,
    220-229; train.py:162-178).  COCO's whole bottom-up set is ~28 GB as packed bf16 rows, so it
    fits one B200 several times over: build the store once, then

        fc, att, att_masks, labels, masks = store.load_batch(ix, labels, masks, stream=copy_stream)

    is the drop-in for `load_data`: `ix` (int64, the loader's per-row image index) plus the
    caption tensors cross PCIe (a few hundred KB), `coopcap_store_gather` assembles the packed
    bf16 operand on the device.  The returned tuple is what AlternatingJointModel.forward takes
    (`att` is a shape carrier; the packed operand rides on `att_masks`, as with upload_batch)."""

    def __init__(self, device, fc_feats: torch.Tensor, att16: torch.Tensor, off: torch.Tensor):
        """fc_feats fp32 [N, F], att16 bf16 [total_regions, D] and off int64 [N+1], all on `device`
        (use the `from_padded` / `append` builders)."""
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.CoopcapError("FeatureStore lives on a CUDA device (there is no CPU path)")
        assert fc_feats.dtype == torch.float32 and att16.dtype == torch.bfloat16 and off.dtype == torch.int64
        self.fc, self.att16, self.off = fc_feats.contiguous(), att16.contiguous(), off.contiguous()
        self.n_img = self.fc.shape[0]
        self.lens_host = (off[1:] - off[:-1]).to("cpu", torch.int32)       # region counts, host copy
        self.last_bytes = 0
        self._slots, self._next = [], 0       # pinned staging for the per-batch offsets / row order

    def _slot(self, B):
        """Round-robin pinned staging (page-locking memory per call would cost more than the step)."""
        if not self._slots or self._slots[0]["off"].numel() < B + 1:
            self._slots = [dict(off=torch.empty(B + 1, dtype=torch.int32).pin_memory(),
                                order=torch.empty(B, dtype=torch.int32).pin_memory(), ev=None)
                           for _ in range(4)]
        sl = self._slots[self._next]
        self._next = (self._next + 1) % len(self._slots)
        if sl["ev"] is not None:
            sl["ev"].synchronize()            # the copy that last read this slot (4 batches ago) is done
        return sl

    @classmethod
    def from_padded(cls, device, fc_feats: torch.Tensor, att_feats: torch.Tensor,
                    att_masks: Optional[torch.Tensor], chunk: int = 256) -> "FeatureStore":
        """Build from loader-shaped host tensors fc [N, F], att [N, L, D] (zero-padded), att_masks
        [N, L] or None.  One-off setup: uploads `chunk` images at a time."""
        device = torch.device(device)
        N, L, D = att_feats.shape
        lens = torch.full((N,), L, dtype=torch.int64) if att_masks is None \
            else (att_masks > 0).sum(1).to(torch.int64)
        off = torch.zeros(N + 1, dtype=torch.int64)
        off[1:] = torch.cumsum(lens, 0)
        att16 = torch.empty(int(off[-1]), D, dtype=torch.bfloat16, device=device)
        for i0 in range(0, N, chunk):
            blk = att_feats[i0:i0 + chunk].to(device)
            keep = (torch.arange(L, device=device)[None, :] < lens[i0:i0 + chunk].to(device)[:, None])
            att16[int(off[i0]):int(off[min(i0 + chunk, N)])] = blk[keep].to(torch.bfloat16)
        return cls(device, fc_feats.to(device, torch.float32), att16, off.to(device))

    def bytes(self) -> int:
        return self.att16.numel() * 2 + self.fc.numel() * 4 + self.off.numel() * 8

    def load_batch(self, ix: torch.Tensor, labels: torch.Tensor, masks: torch.Tensor, *, stream=None):
        """ix int64 [B] (host, ideally pinned): image index of every batch row; labels / masks are
        the loader's caption tensors (host).  Asynchronous on `stream`; no device synchronisation."""
        if ix.device.type != "cpu" or ix.dtype != torch.int64:
            raise _lib.CoopcapError("FeatureStore.load_batch takes a host int64 index tensor")
        B = ix.numel()
        if B == 0 or int(ix.min()) < 0 or int(ix.max()) >= self.n_img:
            raise _lib.CoopcapError(f"image index out of range [0, {self.n_img})")
        lens = self.lens_host[ix]                                  # host gather of B integers
        sl = self._slot(B)
        off, order = sl["off"][: B + 1], sl["order"][:B]
        off[0] = 0
        torch.cumsum(lens, 0, out=off[1:])
        NL, L = int(off[-1]), int(lens.max())
        order.copy_(torch.argsort(lens, descending=True, stable=True))
        home = torch.cuda.current_stream(self.device)            # the stream that will consume the batch
        st = stream if stream is not None else home
        D, F = self.att16.shape[1], self.fc.shape[1]
        nb = lambda t: t.numel() * t.element_size()
        # Outputs come from the CONSUMER stream's allocator pool (no record_stream bookkeeping, no
        # delayed frees, hence no cudaMalloc in steady state: a 230 MB cudaMalloc next to a busy
        # GPU cost tens of ms); the staging stream first waits until the consumer has reached this
        # point, i.e. until whatever used these blocks before has finished.
        i64 = dict(dtype=torch.int64, device=self.device)
        ix_d = torch.empty(B, **i64)
        off_d = torch.empty(B + 1, dtype=torch.int32, device=self.device)
        order_d = torch.empty(B, dtype=torch.int32, device=self.device)
        lab = torch.empty(labels.shape, dtype=labels.dtype, device=self.device)
        msk = torch.empty(masks.shape, dtype=masks.dtype, device=self.device)
        att16 = torch.empty(NL, D, dtype=torch.bfloat16, device=self.device)
        fc = torch.empty(B, F, dtype=torch.float32, device=self.device)
        am = torch.empty(B, L, dtype=torch.float32, device=self.device)
        if st != home:
            ev = torch.cuda.Event()
            ev.record(home)
            st.wait_event(ev)
        with torch.cuda.stream(st):
            ix_d.copy_(ix, non_blocking=True)
            off_d.copy_(off, non_blocking=True)
            order_d.copy_(order, non_blocking=True)
            lab.copy_(labels, non_blocking=True)
            msk.copy_(masks, non_blocking=True)
            check(_lib.load().coopcap_store_gather(
                C.c_void_p(self.att16.data_ptr()), C.c_void_p(self.off.data_ptr()),
                C.c_void_p(self.fc.data_ptr()), self.n_img, C.c_void_p(ix_d.data_ptr()), B, D, F,
                C.c_void_p(off_d.data_ptr()), C.c_void_p(att16.data_ptr()), C.c_void_p(fc.data_ptr()),
                C.c_void_p(st.cuda_stream)))
            # att_masks as the reference's loader shapes it: 1 on the valid regions, width = longest row
            am.copy_(torch.arange(L, device=self.device, dtype=torch.int32)[None, :]
                     < (off_d[1:] - off_d[:-1])[:, None])
            am._coopcap_off = (off_d, NL)
            am._coopcap_order = order_d
            am._coopcap_att16 = att16
            am._coopcap_src = (ix_d,)
            am._coopcap_home = True          # buffers live in the consumer stream's pool: no record_stream
            att = torch.zeros(1, 1, 1, device=self.device).expand(B, L, D)      # shape carrier only
            sl["ev"] = torch.cuda.Event()
            sl["ev"].record(st)
        self.last_bytes = nb(ix) + nb(off) + nb(order) + nb(labels) + nb(masks)
        return fc, att, am, lab, msk


def record_stream(batch, stream):
    """Tell the caching allocator that `stream` uses the tensors of an uploaded batch (they were
    allocated on the upload stream), including the packed side buffers."""
    batch = list(batch)
    if any(getattr(t, "_coopcap_home", False) for t in batch if t is not None):
        return                      # FeatureStore batches already live in the consumer stream's pool
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
