"""Print the handful of `ncu --set full` metrics that decide a kernel's bound, per launch.

    python profiles/ncu_brief.py gpurun_out/x.ncu-rep [--stalls] > profiles/rNN_x_brief.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%act"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%el"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__cycles_active.avg", "cycles"),
]


def main():
    path = sys.argv[1]
    stalls = "--stalls" in sys.argv
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        print(name[:110])
        parts = []
        for k, short in KEYS:
            if k in col:
                parts.append(f"{short}={r[col[k]]}{units[col[k]] if short in ('time', 'dram_rd', 'dram_wr') else ''}")
        print("   " + "  ".join(parts))
        if stalls:
            st = [(h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""),
                   float(r[i] or 0)) for h, i in col.items()
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
            st.sort(key=lambda x: -x[1])
            print("   stalls/issue: " + "  ".join(f"{n}={v:.2f}" for n, v in st[:7]))


if __name__ == "__main__":
    main()
