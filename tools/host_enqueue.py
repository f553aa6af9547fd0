"""Developer tool (GPU box): where does the HOST time of one training step go?
Prints enqueue ms/step, launches/step and a cProfile of five steps.  `python tools/host_enqueue.py [rows]`"""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as BN  # noqa: E402
import cooperativeimagecaptioning_b200.models as models  # noqa: E402
from cooperativeimagecaptioning_b200 import _lib, optimizer as OPT  # noqa: E402
from cooperativeimagecaptioning_b200.data import row_order  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
opt = BN.make_opt(rows)
torch.manual_seed(0)
model = models.AlternatingJointModel(opt).to(dev).train()
with torch.no_grad():
    model.caption_generator.logit.bias[0] = -1e4
optim = OPT.define_optimizer(model, opt)
h = BN.host_batch(rows, 100, 10, 1239, pin=False)
d = {k: h[k].to(dev) for k in ("fc", "att", "att_masks", "labels", "masks")}
off = torch.zeros(rows + 1, dtype=torch.int32)
off[1:] = torch.cumsum(h["lens"], 0).to(torch.int32)
d["att_masks"]._coopcap_off = (off.to(dev), int(off[-1]))
d["att_masks"]._coopcap_order = row_order(h["lens"]).to(dev)


def step():
    optim.zero_grad()
    loss = model(d["fc"], d["labels"], d["masks"], None, d["att"], d["att_masks"], is_alternating=True,
                 alternating_turn="speaker")
    loss.backward()
    optim.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
n = 20
l0 = lib.coopcap_launch_count()
t0 = time.perf_counter()
for _ in range(n):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"rows {rows}: enqueue {1e3 * (t1 - t0) / n:.2f} ms/step, total {1e3 * (t2 - t0) / n:.2f} ms/step, "
      f"launches/step {(lib.coopcap_launch_count() - l0) / n}")
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(25)
