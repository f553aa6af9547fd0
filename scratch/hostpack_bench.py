import os, sys, torch, ctypes as C, time
sys.path.insert(0, os.getcwd())
from cooperativeimagecaptioning_b200 import _lib
lib = _lib.load()
print("cpus", os.cpu_count())
B, L, D = 1024, 100, 2048
att = torch.randn(B, L, D).pin_memory()
lens = torch.randint(10, 101, (B,), dtype=torch.int32)
off = torch.zeros(B + 1, dtype=torch.int32); off[1:] = torch.cumsum(lens, 0)
NL = int(off[-1])
dst = torch.empty(NL, D, dtype=torch.bfloat16).pin_memory()
dev = torch.empty(NL, D, dtype=torch.bfloat16, device="cuda")
for nt in (4, 8, 12, 15, 16, 24, 32):
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        job = lib.coopcap_host_pack_start(C.c_void_p(att.data_ptr()), C.c_void_p(off.data_ptr()), B, L, D, C.c_void_p(dst.data_ptr()), nt)
        lib.coopcap_host_pack_wait(job)
        ts.append(time.perf_counter() - t0)
    print(f"threads {nt}: best {1e3*min(ts):.2f} ms  ({NL*D*4/1e9/min(ts):.1f} GB/s read)")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    e0.record(); dev.copy_(dst, non_blocking=True); e1.record(); torch.cuda.synchronize()
    print(f"DMA {NL*D*2/1e6:.0f} MB: {e0.elapsed_time(e1):.2f} ms ({NL*D*2/1e6/e0.elapsed_time(e1):.1f} GB/s)")
