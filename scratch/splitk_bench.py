import os, sys, torch
sys.path.insert(0, os.getcwd())
from scratch.gemm_bench import bench
for name,M,N,K,am,bm in [("dxh 1024x1024x3072 bm1",1024,1024,3072,0,1),("a2c_dgrad 1024x512x1024 bm1",1024,512,1024,0,1),("a2c 1024x1024x512",1024,1024,512,0,0),("gates 1024x3072x1024",1024,3072,1024,0,0)]:
    for tile in (64,128,256):
        row=[]
        for split in (1,2,3,4):
            us=bench(M,N,K,am,bm,tile,False,2 if split>1 else 0,split)
            row.append(f"{us:6.1f}")
        print(f"{name:30s} tile {tile:3d} split1/2/3/4 us: {' '.join(row)}")
