"""CPU: the oracle restatement (oracle/*.py) against golden vectors produced by the REAL reference
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU parity tests then compare
the CUDA path against the oracle."""
import numpy as np
import pytest
import torch

from oracle import cases


def _close(tag, a, b, rtol=5e-4, atol=1e-6):
    a = torch.as_tensor(np.asarray(a)) if not torch.is_tensor(a) else a
    b = torch.as_tensor(np.asarray(b))
    if b.dtype in (torch.int64, torch.int32):
        assert torch.equal(a.long(), b.long()), tag
        return
    scale = max(float(b.abs().max()), 1e-30)
    err = float((a.double() - b.double()).abs().max())
    assert err <= atol + rtol * scale, f"{tag}: err {err} scale {scale}"


@pytest.mark.parametrize("name", cases.golden_names())
def test_oracle_matches_reference_golden(name):
    meta, z = cases.load_golden(name)
    out = cases.run_oracle(meta)
    for key in z.files:
        if key.startswith("out."):
            k = key[4:]
            if k == "seq_greedy":
                continue
            _close(f"{name}.{k}", out[k], z[key])
        elif key.startswith("grad."):
            _close(f"{name}.{key}", out["grads"][key[5:]], z[key], rtol=1e-3, atol=1e-7)
        elif key.startswith("gradnorm."):
            g = out["grads"][key[9:]]
            ref = float(z[key])
            assert abs(float(g.norm()) - ref) <= 1e-3 * max(ref, 1e-12) + 1e-9, key
        elif key.startswith("gradsample."):
            g = out["grads"][key[11:]]
            s = g.flatten()[:: max(1, g.numel() // 257)][:257]
            _close(f"{name}.{key}", s, z[key], rtol=1e-3, atol=1e-8)


@pytest.mark.parametrize("name", cases.cider_golden_names())
def test_cider_oracle_matches_reference_scorer(name):
    """oracle/cider.py against scores produced by the reference's own CiderD scorer
    (tests/golden/make_golden_cider.py): float64, sums in the reference's order -> 1e-12."""
    from oracle import cider as OC
    meta, gts, gen, greedy, df, ref_len, z = cases.load_cider_golden(name)
    cg, diff, gm = OC.self_critical_reward(gts, gen, greedy, df, ref_len)
    assert np.max(np.abs(cg - z["out.cider_gen"])) <= 1e-12
    assert np.max(np.abs(diff - z["out.reward"])) <= 1e-12
    assert abs(gm - float(z["out.cider_greedy"])) <= 1e-12


def test_cider_caption_conventions():
    """The first 0 is a word, later ids are ignored; `length` counts bigrams (ciderD_scorer.py:143)."""
    from oracle import cider as OC
    assert OC.caption_words([4, 7, 0, 9, 0]) == (4, 7, 0)
    assert OC.caption_words([4, 7, 9]) == (4, 7, 9)
    c = OC.precook((4, 7, 4, 7, 0))
    assert c[(4,)] == 2 and c[(4, 7)] == 2 and c[(4, 7, 4, 7)] == 1 and list(c)[:3] == [(4,), (7,), (0,)]
    same = OC.ciderd_scores([(4, 7, 0), (5, 0)], [[(4, 7, 0)], [(9, 9, 0)]])
    assert same[0] > 0 and same[1] == 0.0
