import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import bench as BN
import cooperativeimagecaptioning_b200.models as models
from cooperativeimagecaptioning_b200 import optimizer as OPT, _lib
lib = _lib.load()
dev = torch.device("cuda", 0)
opt = BN.make_opt(1024)
torch.manual_seed(0)
model = models.AlternatingJointModel(opt).to(dev).train()
with torch.no_grad():
    model.caption_generator.logit.bias[0] = -1e4
optim = OPT.define_optimizer(model, opt)
hb = [BN.host_batch(1024, 100, 10, 1239 + i, pin=False) for i in range(2)]
def to_device(h):
    d = {k: h[k].to(dev) for k in ("fc", "att", "att_masks", "labels", "masks")}
    off = torch.zeros(1025, dtype=torch.int32); off[1:] = torch.cumsum(h["lens"], 0).to(torch.int32)
    d["att_masks"]._coopcap_off = (off.to(dev), int(off[-1]))
    return d
res = [to_device(h) for h in hb]
def step(d, parts=None):
    t0 = time.perf_counter()
    optim.zero_grad()
    loss = model(d["fc"], d["labels"], d["masks"], None, d["att"], d["att_masks"], is_alternating=True, alternating_turn="speaker")
    t1 = time.perf_counter()
    loss.backward()
    t2 = time.perf_counter()
    optim.step()
    t3 = time.perf_counter()
    if parts is not None: parts.append((t1-t0, t2-t1, t3-t2))
for i in range(4): step(res[i % 2])
torch.cuda.synchronize()
for tag, same in (("alternating", False), ("same batch", True)):
    parts = []
    st0 = torch.cuda.memory_stats()
    t0 = time.perf_counter()
    for i in range(10): step(res[0 if same else i % 2], parts)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    st1 = torch.cuda.memory_stats()
    f = sum(p[0] for p in parts)/10; b = sum(p[1] for p in parts)/10; o = sum(p[2] for p in parts)/10
    print(f"{tag}: enqueue {1e3*(t1-t0)/10:.2f} ms/step (fwd {1e3*f:.2f} bwd {1e3*b:.2f} opt {1e3*o:.2f}), total {1e3*(t2-t0)/10:.2f} ms/step; "
          f"cudaMalloc calls {st1['num_device_alloc']-st0['num_device_alloc']}, frees {st1['num_device_free']-st0['num_device_free']}, retries {st1['num_alloc_retries']-st0['num_alloc_retries']}, "
          f"reserved {st1['reserved_bytes.all.current']/2**30:.1f} GiB, launches/step {0}")
