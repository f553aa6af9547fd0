"""Developer tool (GPU box, torchrun): cost of the gradient exchange in the data-parallel step.
Variants: one all-reduce after backward (the product's default), slices exchanged early on a side
stream (COOPCAP_EARLY_REDUCE=1), no exchange at all."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as BN  # noqa: E402
import cooperativeimagecaptioning_b200.models as models  # noqa: E402
from cooperativeimagecaptioning_b200 import optimizer as OPT  # noqa: E402
from cooperativeimagecaptioning_b200.data import row_order  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if os.environ.get("HIPRI", "0") == "1":
    o = dist.ProcessGroupNCCL.Options()
    o.is_high_priority_stream = True
    dist.init_process_group("nccl", device_id=dev, pg_options=o)
else:
    dist.init_process_group("nccl", device_id=dev)
rows = 1024
opt = BN.make_opt(rows)
torch.manual_seed(0)
model = models.AlternatingJointModel(opt).to(dev).train()
with torch.no_grad():
    model.caption_generator.logit.bias[0] = -1e4
optim = OPT.define_optimizer(model, opt)
h = BN.host_batch(rows, 100, 10, 1239 + rank, pin=False)
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


def timed(tag, n=30):
    for _ in range(8):
        step()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{tag}: {float(t):.3f} ms/step", flush=True)


OPT._EARLY_REDUCE = True
timed("early slices (COOPCAP_EARLY_REDUCE=1)")
real_async = OPT.FlatAdam.reduce_async
OPT.FlatAdam.reduce_async = lambda self, params: False
timed("one late all-reduce (default)")
real_ar = dist.all_reduce
OPT.dist = None
import torch.distributed as D2
D2.all_reduce = lambda *a, **k: None
timed("no exchange")
D2.all_reduce = real_ar
OPT.FlatAdam.reduce_async = real_async
timed("early slices, again")
dist.destroy_process_group()
