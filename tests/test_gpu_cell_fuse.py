"""GPU: the opt-in (COOPCAP_CELL_FUSE=1) recurrent-cell steps fused into their GEMM (csrc/cell_step.cuh: GRU update as the
epilogue of the gh GEMM, maxout-LSTM update as the epilogue of the a2c GEMM) are BIT-identical to
the two-kernel path they replace (gemm + gru_fwd_kernel / lstm_fwd_kernel), on a free-running
ST-Gumbel decode with Philox noise and dropout followed by the listener forward.  Rows 200 leaves a
partial 128-row tile, 1024 is the bench size."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(path, rows, unfused):
    env = dict(os.environ)
    env.pop("COOPCAP_CELL_FUSE", None)
    if not unfused:
        env["COOPCAP_CELL_FUSE"] = "1"
    r = subprocess.run([sys.executable, os.path.join(HERE, "cell_fuse_probe.py"), path, str(rows)], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return torch.load(path)


@pytest.mark.parametrize("rows", [200, 1024])
def test_fused_cell_steps_bit_identical(tmp_path, rows):
    a = _run(str(tmp_path / "fused.pt"), rows, unfused=False)
    b = _run(str(tmp_path / "unfused.pt"), rows, unfused=True)
    assert a.keys() == b.keys()
    for k in a:
        if a[k].dtype.is_floating_point:
            assert torch.equal(a[k].float().nan_to_num(7.0), b[k].float().nan_to_num(7.0)), k
        else:
            assert torch.equal(a[k], b[k]), k
    assert int(a["sp.cap_len"].sum()) > 0
