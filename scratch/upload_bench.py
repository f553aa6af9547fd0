import os, sys, time, torch
sys.path.insert(0, os.getcwd())
import bench as BN
from cooperativeimagecaptioning_b200.data import upload_batch
dev = torch.device("cuda", 0)
h = BN.host_batch(1024, 100, 10, 1239, pin=True)
side = torch.cuda.Stream()
def timeit(fn, n=8):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n
valid = int(h["lens"].sum()) * 2048 * 4
print("padded bytes %.0f MB, valid bytes %.0f MB" % (h["att"].numel() * 4 / 1e6, valid / 1e6))
ms = timeit(lambda: h["att"].to(dev, non_blocking=True))
print(f"padded DMA copy: {ms:.2f} ms  ({h['att'].numel()*4/ms/1e6:.1f} GB/s)")
ms = timeit(lambda: upload_batch(h["fc"], h["att"], h["att_masks"], h["labels"], h["masks"], dev, stream=side, zero_copy=False))
print(f"ragged-row DMA copies: {ms:.2f} ms ({valid/ms/1e6:.1f} GB/s of valid bytes)")
for c in (16, 32, 64, 128, 256, 592):
    ms = timeit(lambda: upload_batch(h["fc"], h["att"], h["att_masks"], h["labels"], h["masks"], dev, stream=side, zero_copy=True, ctas=c))
    print(f"zero-copy kernel ctas={c}: {ms:.2f} ms ({valid/ms/1e6:.1f} GB/s of valid bytes)")
