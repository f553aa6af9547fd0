"""GPU: host -> device staging (data.upload_batch) gives the kernels exactly what a plain
`.cuda()` of the loader's padded tensors gives them."""
import pytest
import torch

from oracle import synth
from oracle.ref_loader import reference_opt
from gpu_util import REAL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("zero_copy", [True, False])
def test_upload_batch_matches_plain_copy(zero_copy):
    import cooperativeimagecaptioning_b200.models as models
    from cooperativeimagecaptioning_b200.data import upload_batch
    d = REAL
    B, L = 12, 9
    batch = synth.make_batch(d, B, L, 7, varlen=True, min_regions=2)
    spk = models.setup(reference_opt(), "att2in2", "caption_model").cuda().eval()
    spk.keep_passes = True
    with torch.no_grad():
        spk.sample(batch.fc_feats.cuda(), batch.att_feats.cuda(), batch.att_masks.cuda(), {"sample_max": 1})
        pin = lambda t: t.contiguous().pin_memory()
        side = torch.cuda.Stream()
        fc, att, am, lab, msk = upload_batch(pin(batch.fc_feats), pin(batch.att_feats),
                                             pin(batch.att_masks), pin(batch.labels), pin(batch.masks),
                                             "cuda", stream=side, zero_copy=zero_copy)
        torch.cuda.current_stream().wait_stream(side)
        assert att.shape == batch.att_feats.shape
        assert (getattr(am, "_coopcap_att16", None) is not None) == zero_copy
        spk.sample(fc, att, am, {"sample_max": 1})
    torch.cuda.synchronize()
    a, b = spk._passes
    assert a.NL == b.NL == int((batch.att_masks > 0).sum())
    assert torch.equal(a.t["att16"], b.t["att16"])          # packed bf16 regions: bit-identical
    assert torch.equal(a.t["tok_out"], b.t["tok_out"])      # hence identical greedy captions
    assert torch.equal(lab.cpu(), batch.labels) and torch.equal(msk.cpu(), batch.masks)
    assert upload_batch.last_bytes < batch.att_feats.numel() * 4 + 10 ** 6


@pytest.mark.parametrize("varlen", [True, False])
def test_host_packer_matches_plain_copy(varlen):
    """data.HostPacker: bf16 packing by the library's worker threads + one DMA copy gives the
    kernels the bit-identical operand; several batches in flight reuse the staging slots."""
    import cooperativeimagecaptioning_b200.models as models
    from cooperativeimagecaptioning_b200.data import HostPacker
    d = REAL
    B, L = 12, 9
    spk = models.setup(reference_opt(), "att2in2", "caption_model").cuda().eval()
    spk.keep_passes = True
    packer = HostPacker("cuda", nbuf=2, threads=3)
    side = torch.cuda.Stream()
    pin = lambda t: None if t is None else t.contiguous().pin_memory()
    batches = [synth.make_batch(d, B, L, 7 + i, varlen=varlen, min_regions=2) for i in range(3)]
    jobs = [packer.start(pin(b.fc_feats), pin(b.att_feats), pin(b.att_masks), pin(b.labels), pin(b.masks))
            for b in batches[:2]]
    with torch.no_grad():
        for i, b in enumerate(batches):
            fc, att, am, lab, msk = packer.finish(jobs[i], stream=side)
            if i + 2 < len(batches):       # third batch recycles the first staging slot
                nb = batches[i + 2]
                jobs.append(packer.start(pin(nb.fc_feats), pin(nb.att_feats), pin(nb.att_masks),
                                         pin(nb.labels), pin(nb.masks)))
            torch.cuda.current_stream().wait_stream(side)
            spk.sample(fc, att, am, {"sample_max": 1})
            spk.sample(b.fc_feats.cuda(), b.att_feats.cuda(),
                       None if b.att_masks is None else b.att_masks.cuda(), {"sample_max": 1})
            assert torch.equal(lab.cpu(), b.labels) and torch.equal(msk.cpu(), b.masks)
    torch.cuda.synchronize()
    for i in range(len(batches)):
        a, r = spk._passes[2 * i], spk._passes[2 * i + 1]
        assert a.NL == r.NL
        assert torch.equal(a.t["att16"], r.t["att16"])
        assert torch.equal(a.t["tok_out"], r.t["tok_out"])
    assert packer.last_bytes < batches[-1].att_feats.numel() * 2 + 10 ** 6


@pytest.mark.parametrize("varlen", [True, False])
def test_feature_store_batches_equal_plain_copies(varlen):
    """data.FeatureStore: with every image's features resident in HBM a step ships only image
    indices; the gathered packed operand, fc vectors, masks and offsets must be exactly what a plain
    `.cuda()` of the loader's padded batch (same images, same order, repeats included) gives."""
    import cooperativeimagecaptioning_b200.models as models
    from cooperativeimagecaptioning_b200.data import FeatureStore, record_stream
    d = REAL
    N, L, B = 40, 11, 23
    pool = synth.make_batch(d, N, L, 91, varlen=varlen, min_regions=2)
    store = FeatureStore.from_padded("cuda", pool.fc_feats, pool.att_feats, pool.att_masks, chunk=16)
    assert store.n_img == N and store.att16.shape[0] == (int((pool.att_masks > 0).sum()) if varlen else N * L)
    g = torch.Generator().manual_seed(5)
    ix = torch.randint(0, N, (B,), generator=g)
    ix[3] = ix[7]                                   # seq_per_img style repeats are fine
    cap = synth.make_batch(d, B, L, 92)             # captions of the batch rows
    spk = models.setup(reference_opt(), "att2in2", "caption_model").cuda().eval()
    spk.keep_passes = True
    side = torch.cuda.Stream()
    with torch.no_grad():
        fc, att, am, lab, msk = store.load_batch(ix.pin_memory(), cap.labels.pin_memory(),
                                                 cap.masks.pin_memory(), stream=side)
        torch.cuda.current_stream().wait_stream(side)
        record_stream((fc, att, am, lab, msk), torch.cuda.current_stream())
        spk.sample(fc, att, am, {"sample_max": 1})
        # the same rows through the reference-shaped path: padded features + masks, width = longest row
        ref_m = None if pool.att_masks is None else pool.att_masks[ix]
        Lb = L if ref_m is None else int(ref_m.sum(1).max())
        ref_att = pool.att_feats[ix][:, :Lb].contiguous()
        ref_m = torch.ones(B, Lb) if ref_m is None else ref_m[:, :Lb].contiguous()
        spk.sample(pool.fc_feats[ix].cuda(), ref_att.cuda(), ref_m.cuda(), {"sample_max": 1})
    torch.cuda.synchronize()
    a, b = spk._passes
    assert a.NL == b.NL and att.shape == ref_att.shape
    assert torch.equal(a.t["att16"], b.t["att16"])          # packed bf16 regions: bit-identical
    assert torch.equal(a.t["tok_out"], b.t["tok_out"])      # hence identical greedy captions
    assert torch.equal(fc.cpu(), pool.fc_feats[ix]) and torch.equal(am.cpu(), ref_m)
    assert torch.equal(lab.cpu(), cap.labels) and torch.equal(msk.cpu(), cap.masks)
    # what crossed PCIe: indices + captions, not features
    assert store.last_bytes < 64 * B + cap.labels.numel() * 8 + cap.masks.numel() * 4
    from cooperativeimagecaptioning_b200._lib import CoopcapError
    with pytest.raises(CoopcapError):
        store.load_batch(torch.tensor([N]), cap.labels[:1], cap.masks[:1])
