"""Developer tool (GPU box): ST-Gumbel decode of N rows x 16 tokens, timed with CUDA events; a short
command line to put under ncu (`-k regex:logit_sample`).  `python tools/decode_bench.py [rows] [reps]`"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as BN  # noqa: E402
import cooperativeimagecaptioning_b200.models as models  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda", 0)
opt = BN.make_opt(rows)
torch.manual_seed(0)
spk = models.setup(opt, "att2in2", "caption_model").to(dev).train()
with torch.no_grad():
    spk.logit.bias[0] = -1e4
g = torch.Generator().manual_seed(7)
att = torch.randn(rows, 36, 2048, generator=g).to(dev)
for mode, so in (("st_gumbel", dict(sample_max=0, use_one_hot=1)), ("greedy", dict(sample_max=1))):
    with torch.no_grad():
        for _ in range(3):
            spk._sample_pass(att, None, so.get("sample_max", 1), 1.0, so.get("use_one_hot", 0))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            spk._sample_pass(att, None, so.get("sample_max", 1), 1.0, so.get("use_one_hot", 0))
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{mode}: {rows} rows x 16 tokens: {ms:.3f} ms per decode, {rows * 16 / ms / 1e3:.2f} M tok/s")
