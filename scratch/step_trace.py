import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import bench as BN
import cooperativeimagecaptioning_b200.models as models
from cooperativeimagecaptioning_b200 import optimizer as OPT, _lib
dev = torch.device("cuda", 0)
opt = BN.make_opt(1024)
torch.manual_seed(0)
model = models.AlternatingJointModel(opt).to(dev).train()
with torch.no_grad():
    model.caption_generator.logit.bias[0] = -1e4
optim = OPT.define_optimizer(model, opt)
hb = [BN.host_batch(1024, 100, 10, 1239 + i, pin=True) for i in range(2)]
def to_device(h):
    d = {k: h[k].to(dev) for k in ("fc", "att", "att_masks", "labels", "masks")}
    off = torch.zeros(1025, dtype=torch.int32); off[1:] = torch.cumsum(h["lens"], 0).to(torch.int32)
    d["att_masks"]._coopcap_off = (off.to(dev), int(off[-1]))
    return d
res = [to_device(h) for h in hb]
def step(d):
    optim.zero_grad()
    loss = model(d["fc"], d["labels"], d["masks"], None, d["att"], d["att_masks"], is_alternating=True, alternating_turn="speaker")
    loss.backward()
    optim.step()
print("NL:", [int(h["lens"].sum()) for h in hb])
keys = ("num_device_alloc", "num_device_free", "num_alloc_retries")
for rep in range(3):
    evs = []
    st0 = torch.cuda.memory_stats()
    for i in range(30):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); step(res[i % 2]); e1.record(); evs.append((e0, e1))
    torch.cuda.synchronize()
    st1 = torch.cuda.memory_stats()
    ts = [a.elapsed_time(b) for a, b in evs]
    print(f"rep {rep}: " + " ".join(f"{t:.1f}" for t in ts))
    print("   allocator:", {k: st1[k] - st0[k] for k in keys}, "reserved GiB %.2f" % (st1["reserved_bytes.all.current"] / 2**30), "active GiB %.2f" % (st1["active_bytes.all.peak"] / 2**30))
