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
    keep = ~torch.isnan(b.double())                     # ragged tables are NaN-padded
    if not bool(keep.all()):
        a, b = a.double()[keep], b.double()[keep]
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


@pytest.mark.parametrize("name", cases.retrieval_golden_names())
def test_retrieval_oracle_matches_reference(name):
    """oracle/retrieval.py against the reference's own i2t / t2i (make_golden_retrieval.py)."""
    import json
    import os
    from oracle import retrieval as OR
    z = np.load(os.path.join(cases.GOLDEN_DIR, name + ".npz"))
    kw = json.loads(bytes(z["meta"]).decode())
    images, caps = OR.synth_embeddings(kw["n_img"], kw["K"], kw["seed"], noise=kw["noise"])
    m1, (r1, t1) = OR.i2t(images, caps)
    m2, (r2, t2) = OR.t2i(images, caps)
    m3, (r3, t3) = OR.t2i(images[0::5], caps[0::5], use_gen_sent=True)
    assert np.array_equal(r1, z["i2t_ranks"]) and np.array_equal(t1, z["i2t_top1"])
    assert np.array_equal(r2, z["t2i_ranks"]) and np.array_equal(t2, z["t2i_top1"])
    assert np.array_equal(r3, z["gen_ranks"]) and np.array_equal(t3, z["gen_top1"])
    assert np.allclose(m1, z["i2t_metrics"]) and np.allclose(m2, z["t2i_metrics"]) and np.allclose(m3, z["gen_metrics"])
