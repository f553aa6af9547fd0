import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from cooperativeimagecaptioning_b200 import ops
def bench(M,N,K,am,bm,tile,out16=False,mode=0,split=1,iters=20):
    A = torch.randn((M,K) if am==0 else (K,M), device="cuda").bfloat16()
    B = torch.randn((N,K) if bm==0 else (K,N), device="cuda").bfloat16()
    out = None if out16 else torch.zeros(M,N,device="cuda")
    o16 = torch.zeros(M,N,device="cuda",dtype=torch.bfloat16) if out16 else None
    flush = torch.empty(256*1024*1024//4, device="cuda")
    for _ in range(3): ops.gemm(A,B,M,N,K,a_major=am,b_major=bm,out=out,out16=o16,tile_n=tile,mode=mode,split_k=split)
    ts=[]
    for _ in range(iters):
        flush.zero_()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); ops.gemm(A,B,M,N,K,a_major=am,b_major=bm,out=out,out16=o16,tile_n=tile,mode=mode,split_k=split); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); return ts[len(ts)//2]*1e3
if __name__ == "__main__":
    shapes=[("logits 1024x9488x512",1024,9488,512,0,0,False,0,1),
            ("gates 1024x3072x1024",1024,3072,1024,0,0,False,0,1),
            ("a2c 1024x1024x512",1024,1024,512,0,0,False,0,1),
            ("dxh 1024x1024x3072 bm1",1024,1024,3072,0,1,False,0,1),
            ("a2c_dgrad 1024x512x1024 bm1",1024,512,1024,0,1,False,0,1),
            ("logit_dgrad 16384x512x9488 bm1",16384,512,9488,0,1,False,0,1),
            ("logit_wgrad 9488x512x16384 mn",9488,512,16384,1,1,False,0,1),
            ("att_embed 54304x512x2048 ->bf16",54304,512,2048,0,0,True,0,1),
            ("ctx2att 54304x512x512 ->bf16",54304,512,512,0,0,True,0,1),
            ("gi 17408x3072x512",17408,3072,512,0,0,False,0,1),
            ]
    for name,M,N,K,am,bm,o16,mode,split in shapes:
        row=[]
        for tile in (64,128,192,256,0):
            if tile>64 and tile>=2*N: row.append("   -  "); continue
            us=bench(M,N,K,am,bm,tile,o16,mode,split)
            row.append(f"{us:6.1f}")
        print(f"{name:36s} tile64/128/192/256/auto us: {' '.join(row)}   ({2*M*N*K/1e9:.1f} GF)")
    # NVML cost
    import pynvml
    pynvml.nvmlInit(); h=pynvml.nvmlDeviceGetHandleByIndex(0)
    x=torch.randn(8192,8192,device="cuda").bfloat16()
    def busy(n):
        for _ in range(n): y=x@x
    torch.cuda.synchronize()
    t0=time.perf_counter(); busy(50); torch.cuda.synchronize(); t1=time.perf_counter()
    print("50 matmuls alone: %.1f ms"%(1e3*(t1-t0)))
    for fn_name in ("nvmlDeviceGetClockInfo","nvmlDeviceGetCurrentClocksEventReasons","nvmlDeviceGetPowerUsage"):
        fn=getattr(pynvml,fn_name,None)
        if fn is None: print(fn_name,"missing"); continue
        t0=time.perf_counter(); busy(50)
        q0=time.perf_counter()
        for _ in range(3):
            r = fn(h, pynvml.NVML_CLOCK_SM) if fn_name=="nvmlDeviceGetClockInfo" else fn(h)
        q1=time.perf_counter()
        torch.cuda.synchronize(); t1=time.perf_counter()
        print(f"{fn_name}: 3 calls took {1e3*(q1-q0):.1f} ms; 50 matmuls with calls {1e3*(t1-t0):.1f} ms; value {r}")
