"""Developer tool (GPU box): host-side profile of the end-to-end loop fed by data.FeatureStore."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as BN  # noqa: E402
import cooperativeimagecaptioning_b200.models as models  # noqa: E402
from cooperativeimagecaptioning_b200 import optimizer as OPT  # noqa: E402
from cooperativeimagecaptioning_b200.data import FeatureStore, record_stream  # noqa: E402

dev = torch.device("cuda", 0)
rows = 1024
opt = BN.make_opt(rows)
torch.manual_seed(0)
model = models.AlternatingJointModel(opt).to(dev).train()
with torch.no_grad():
    model.caption_generator.logit.bias[0] = -1e4
optim = OPT.define_optimizer(model, opt)
hb = [BN.host_batch(rows, 100, 10, 1239 + i, pin=True) for i in range(2)]
store = FeatureStore.from_padded(dev, torch.cat([h["fc"] for h in hb]), torch.cat([h["att"] for h in hb]),
                                 torch.cat([h["att_masks"] for h in hb]))
g = torch.Generator().manual_seed(1)
idx = [torch.randperm(store.n_img, generator=g)[:rows].contiguous().pin_memory() for _ in range(8)]
copy_stream = torch.cuda.Stream()


def step(d):
    optim.zero_grad()
    loss = model(d[0], d[3], d[4], None, d[1], d[2], is_alternating=True, alternating_turn="speaker")
    loss.backward()
    optim.step()
    return loss.detach()


def upload(i):
    d = store.load_batch(idx[i % 8], hb[i % 2]["labels"], hb[i % 2]["masks"], stream=copy_stream)
    ev = torch.cuda.Event()
    ev.record(copy_stream)
    return d, ev


def loop(n):
    tm = [0.0, 0.0]
    nxt = upload(0)
    for i in range(n):
        d, ev = nxt
        h0 = time.perf_counter()
        if i + 1 < n:
            nxt = upload(i + 1)
        h1 = time.perf_counter()
        torch.cuda.current_stream().wait_event(ev)
        step(d)
        record_stream(d, torch.cuda.current_stream())
        tm[0] += h1 - h0
        tm[1] += time.perf_counter() - h1
    return [1e3 * t / n for t in tm]


loop(8)
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter()
    tm = loop(20)
    torch.cuda.synchronize()
    print(f"rep {rep}: {1e3 * (time.perf_counter() - t0) / 20:.2f} ms/step; host stage {tm[0]:.2f} ms, step {tm[1]:.2f} ms",
          torch.cuda.memory_stats()["num_device_alloc"])
pr = cProfile.Profile()
pr.enable()
loop(20)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
