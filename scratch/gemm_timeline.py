import os, sys, torch
sys.path.insert(0, os.getcwd())
from cooperativeimagecaptioning_b200 import ops
def run(M,N,K,am,bm,tile,name):
    A = torch.randn((M,K) if am==0 else (K,M), device="cuda").bfloat16()
    B = torch.randn((N,K) if bm==0 else (K,N), device="cuda").bfloat16()
    out = torch.zeros(M,N,device="cuda")
    bias = torch.randn(N, device="cuda")
    for _ in range(3): ops.gemm(A,B,M,N,K,a_major=am,b_major=bm,out=out,tile_n=tile,bias=bias)
    dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
    ops.gemm(A,B,M,N,K,a_major=am,b_major=bm,out=out,tile_n=tile,bias=bias,dbg=dbg)
    torch.cuda.synchronize()
    d = dbg.cpu().tolist(); t0 = d[0]
    f = lambda i: (d[i]-t0)/1e3 if d[i] else None
    print(f"{name} tile {tile}: start 0, after pdl_wait {f(1)}, stores drained {f(2)}, end {f(3)} us")
    for i in range(4):
        print(f"    tile#{i}: loads issued {f(8+i)}  mma issued {f(16+i)}  acc ready {f(24+i)}  epilogue issued {f(32+i)}")
run(1024,9488,512,0,0,192,"logits")
run(1024,9488,512,0,0,256,"logits")
run(1024,3072,1024,0,0,192,"gates")
run(1024,1024,512,0,0,64,"a2c")
