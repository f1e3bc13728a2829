"""Decode how tcgen05.mma (kind::tf32) addresses shared-memory operand tiles (MN-major / swizzled):
index-coded tile contents on the probed side, one-hot rows on the other, one K=8 MMA.  Run on the GPU box:
    python profiles/umma_probe.py > gpurun_out/umma_probe.txt
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from matdeeplearn_b200 import _lib

lib = _lib.load_selftest()
dev = torch.device("cuda:0")
NONE, SW128, SW64, SW32 = 0, 2, 4, 6


def kmajor_tile(M, rows):
    """raw image of a [rows x 8] K-major no-swizzle tile (k-chunk stride rows*16, 8-row group stride 128)"""
    raw = np.zeros(rows * 8, dtype=np.float32)
    for r in range(rows):
        for k in range(8):
            off = (k >> 2) * rows * 16 + (r >> 3) * 128 + (r & 7) * 16 + (k & 3) * 4
            raw[off // 4] = M[r, k]
    return raw


def run(rawA, rawB, N, lbo_a, sbo_a, lbo_b, sbo_b, a_mn, b_mn, la=NONE, lb=NONE, step_a=0, step_b=0, nmma=1):
    A = torch.from_numpy(rawA).to(dev)
    B = torch.from_numpy(rawB).to(dev)
    D = torch.full((128, N), float("nan"), device=dev)
    rc = lib.mdl_selftest_umma_probe(_lib.ptr(A), A.numel(), _lib.ptr(B), B.numel(), _lib.ptr(D), N,
                                     lbo_a, sbo_a, lbo_b, sbo_b, a_mn, b_mn, nmma, step_a, step_b, la, lb, _lib.stream())
    _lib.check(rc, "probe")
    torch.cuda.synchronize()
    return D.cpu().numpy()


ONEHOT_A = np.zeros((128, 8), dtype=np.float32)
for m in range(128):
    ONEHOT_A[m, m % 8] = 1.0


def probe_b(N, lbo, sbo, b_mn, layout, words=2048):
    rawB = np.arange(words, dtype=np.float32) % 2048
    D = run(kmajor_tile(ONEHOT_A, 128), rawB, N, 128 * 16, 128, lbo, sbo, 0, b_mn, NONE, layout)
    return D[:8, :].astype(np.int64)      # [k][n] -> word index read for (n, k)


def probe_a(lbo, sbo, a_mn, layout, N=16):
    B = np.zeros((N, 8), dtype=np.float32)
    for n in range(N):
        B[n, n % 8] = 1.0
    rawA = np.arange(2048, dtype=np.float32)
    D = run(rawA, kmajor_tile(B, N), N, lbo, sbo, N * 16, 128, a_mn, 0, layout, NONE)
    return D[:, :8].astype(np.int64)      # [m][k] -> word index read for (m, k)


np.set_printoptions(linewidth=250)
print("== sanity: K-major no-swizzle B, N=16")
print(probe_b(16, 16 * 16, 128, 0, NONE))
for name, lay in (("SW128", SW128), ("SW64", SW64), ("SW32", SW32)):
    for N in (16, 64):
        for lbo, sbo in ((4096, 1024), (1024, 4096), (16, 1024), (2048, 1024), (1024, 2048)):
            print(f"== B MN-major {name} N={N} lbo={lbo} sbo={sbo}: rows k=0..7, cols n (word index read)")
            print(probe_b(N, lbo, sbo, 1, lay))
    print(f"== B K-major {name} N=16 lbo=16 sbo=1024: rows k, cols n")
    print(probe_b(16, 16, 1024, 0, lay))
for name, lay in (("SW128", SW128), ("SW64", SW64)):
    for lbo, sbo in ((4096, 1024), (1024, 4096)):
        print(f"== A MN-major {name} lbo={lbo} sbo={sbo}: rows m=0..39, cols k (word index read)")
        print(probe_a(lbo, sbo, 1, lay)[:40])
print("== A K-major SW128 lbo=16 sbo=1024: rows m=0..15, cols k")
print(probe_a(16, 1024, 0, SW128)[:16])
