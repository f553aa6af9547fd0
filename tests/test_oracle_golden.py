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
