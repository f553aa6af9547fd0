"""CPU: the reference arm of bench.py (`--impl reference`: the real reference from baseline/_ref on
the host cores when that tree is present -- oracle/make_ref.py builds it -- else the oracle port of
the same training step) prints exactly one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--steps", "1", "--warmup", "1", "--cpu-rows", "4"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "joint_gumbel_train_images_per_sec"
    assert d["unit"] == "images/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"].startswith("gumbel_joint_step_varlen")
    have_ref = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "models", "AlternatingJointModel.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port")
    assert d["cpu_baseline"]["cores"] >= 1
    assert d["config"]["rows_per_gpu"] == 4          # the rows this arm actually ran
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
