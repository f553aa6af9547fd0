import time, torch
x = torch.randn(2048, 2048, device="cuda")
torch.cuda.synchronize()
# many small launches per iteration, sync each iteration: wall time per iteration should be flat
ts = []
t_end = time.perf_counter() + 6.0
while time.perf_counter() < t_end:
    t0 = time.perf_counter()
    for _ in range(50):
        y = x @ x
    torch.cuda.synchronize()
    ts.append((time.perf_counter(), (time.perf_counter() - t0) * 1e3))
import statistics
d = [b for a, b in ts]
med = statistics.median(d)
print(f"{len(d)} iterations, median {med:.2f} ms, max {max(d):.2f} ms")
t00 = ts[0][0]
for a, b in ts:
    if b > 1.5 * med:
        print(f"  stall at t={a - t00:6.3f} s: {b:.2f} ms")
