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
