"""tcgen05.st cost per warp (cycles), alone and under concurrent MMAs.  python profiles/tmem_st_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matdeeplearn_b200 import _lib
lib = _lib.load_selftest()
out = torch.zeros(18, dtype=torch.int64, device="cuda:0")
print("warps stores width mma wait_each | cycles/warp (min..max)  per-store | mma issue / complete")
for nwarps in (1, 4, 8, 16):
    for width in (8, 16, 32):
        for mma in (0, 64):
            for wait_each in (0, 1):
                nst = 16
                out.zero_()
                _lib.check(lib.mdl_selftest_tmem_st_bench(_lib.ptr(out), nwarps, nst, width, mma, wait_each, _lib.stream()), "bench")
                torch.cuda.synchronize()
                v = out.cpu().tolist()
                w = v[:nwarps]
                print(f"{nwarps:5d} {nst:6d} {width:5d} {mma:3d} {wait_each:9d} | {min(w):6d}..{max(w):6d}  {max(w) / nst:7.1f} | {v[16]:6d} / {v[17]:6d}")
