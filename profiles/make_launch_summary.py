"""Turn an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv
--log-file <csv>` capture of `bench.py` into (1) the launch list of ONE training step and (2) a
per-kernel-class summary.  The step window is the launches after the second-to-last
`clamp_adam_kernel` up to and including the last one.

    python profiles/make_launch_summary.py gpurun_out/launches.csv profiles/r01_final

writes <prefix>_ncu_launches_one_step.csv and <prefix>_ncu_launch_summary.json.
"""
import csv
import json
import re
import sys
from collections import OrderedDict

CLASSES = [
    ("adam", r"clamp_adam"),
    # same classes as bench.py's timeline (prof_mark(...) at the launch sites): the fused logit GEMM +
    # sampler is its own class, the opt-in fused cell steps count as GEMMs
    ("logit_sample", r"logit_sample_kernel"),
    ("gemm", r"gemm_tc_kernel|gemm_simt|cell_step_kernel"), ("att_fwd", r"attention_fwd"),
    ("att_bwd", r"attention_bwd"), ("att_deferred", r"attention_deferred"), ("lstm", r"lstm_"),
    ("sample", r"sample_kernel|ps_vec|ps_mask|ban_prev"), ("st_bwd", r"st_bwd"), ("logp_bwd", r"logp_bwd"),
    ("gru", r"gru_"), ("hinge", r"hinge|l2norm|pool_kernel"),
    ("reduce", r"colsum|embed_grad|embed_scatter|ps_dpre"),
    ("pack", r"pack_att|cast_|gather_embed|mask_pre|fill_embed"), ("misc", r"coopcap::"),
]


def classify(name):
    for cls, pat in CLASSES:
        if re.search(pat, name):
            return cls
    return "torch/other"


def main(src, prefix):
    rows = OrderedDict()
    with open(src, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        d = rows.setdefault(int(r["ID"]), dict(kernel=r["Kernel Name"], grid=r["Grid Size"], block=r["Block Size"]))
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    launches = [rows[k] for k in sorted(rows)]
    adam = [i for i, r in enumerate(launches) if "clamp_adam" in r["kernel"]]
    assert len(adam) >= 2, "need at least two optimizer steps in the capture"
    step = launches[adam[-2] + 1: adam[-1] + 1]
    with open(prefix + "_ncu_launches_one_step.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "block", "time_us", "dram_read_MB", "dram_write_MB"])
        for r in step:
            w.writerow([r["kernel"][:100], r["grid"], r["block"], f'{r["gpu__time_duration.sum"] / 1e3:.2f}',
                        f'{r.get("dram__bytes_read.sum", 0) / 1e6:.3f}',
                        f'{r.get("dram__bytes_write.sum", 0) / 1e6:.3f}'])
    by = OrderedDict()
    for r in step:
        c = by.setdefault(classify(r["kernel"]), dict(launches=0, us=0.0, dram_MB=0.0))
        c["launches"] += 1
        c["us"] += r["gpu__time_duration.sum"] / 1e3
        c["dram_MB"] += (r.get("dram__bytes_read.sum", 0) + r.get("dram__bytes_write.sum", 0)) / 1e6
    out = dict(note="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                    "--clock-control none; one Gumbel joint step (1024 rows, 10-100 regions); per-launch times "
                    "are cold-cache and serialised (compare shares)",
               step_total_us=sum(c["us"] for c in by.values()), launches=len(step), by_class=by)
    with open(prefix + "_ncu_launch_summary.json", "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: (v["launches"], round(v["us"], 1)) for k, v in by.items()}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
