"""CPU: the host-side packer of libcoopcap (coopcap_host_pack_*): bf16 round-to-nearest-even of the
valid regions, bit-identical to torch's conversion; ragged, empty-row and unaligned cases."""
import ctypes as C

import pytest
import torch

from cooperativeimagecaptioning_b200 import _lib


def _pack(att, lens, threads):
    lib = _lib.load()
    B, L, D = att.shape
    off = None
    NL = B * L
    if lens is not None:
        off = torch.zeros(B + 1, dtype=torch.int32)
        off[1:] = torch.cumsum(lens.to(torch.int32), 0)
        NL = int(off[-1])
    dst = torch.full((max(NL, 1), D), -1.0, dtype=torch.bfloat16)
    job = lib.coopcap_host_pack_start(C.c_void_p(att.data_ptr()),
                                      C.c_void_p(off.data_ptr()) if off is not None else None, B, L, D,
                                      C.c_void_p(dst.data_ptr()), threads)
    assert job >= 0, _lib.last_error() if hasattr(_lib, "last_error") else job
    assert lib.coopcap_host_pack_wait(job) == 0
    return dst[:NL]


@pytest.mark.parametrize("B,L,D,threads", [(7, 5, 2048, 3), (3, 4, 40, 1), (33, 9, 512, 0)])
def test_host_pack_matches_torch_bf16(B, L, D, threads):
    g = torch.Generator().manual_seed(B * 100 + L)
    att = torch.randn(B, L, D, generator=g) * torch.logspace(-3, 3, D)[None, None, :]
    att[0, 0, :4] = torch.tensor([0.0, -0.0, 1e-40, float("inf")])
    lens = torch.randint(0, L + 1, (B,), generator=g)
    lens[0] = L
    got = _pack(att, lens, threads)
    want = torch.cat([att[b, : int(lens[b])] for b in range(B)]).to(torch.bfloat16)
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    # fixed region count (att_masks is None in the reference loader when all rows are full)
    got = _pack(att, None, threads)
    assert torch.equal(got.view(torch.int16), att.reshape(B * L, D).to(torch.bfloat16).view(torch.int16))


def test_host_pack_jobs_overlap_and_reject_bad_offsets():
    lib = _lib.load()
    att = torch.randn(16, 6, 256)
    outs, jobs = [], []
    for i in range(4):     # several jobs in flight, waited out of order
        dst = torch.empty(16 * 6, 256, dtype=torch.bfloat16)
        outs.append(dst)
        jobs.append(lib.coopcap_host_pack_start(C.c_void_p(att.data_ptr()), None, 16, 6, 256,
                                                C.c_void_p(dst.data_ptr()), 2))
    for j in reversed(jobs):
        assert j >= 0 and lib.coopcap_host_pack_wait(j) == 0
    want = att.reshape(96, 256).to(torch.bfloat16)
    for dst in outs:
        assert torch.equal(dst.view(torch.int16), want.view(torch.int16))
    off = torch.tensor([0, 9, 10], dtype=torch.int32)      # 9 regions in a row of L = 6
    assert lib.coopcap_host_pack_start(C.c_void_p(att.data_ptr()), C.c_void_p(off.data_ptr()), 2, 6, 256,
                                       C.c_void_p(outs[0].data_ptr()), 1) < 0
