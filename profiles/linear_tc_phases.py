"""Per-phase cycles of k_linear_tc (thread 0 of every CTA) on [E,128] x [128,128]."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matdeeplearn_b200 import _lib
lib = _lib.load()
dev = "cuda:0"
E, F_ = 102086, 128
x, w, b = torch.randn(E, F_, device=dev), torch.randn(F_, F_, device=dev), torch.randn(F_, device=dev)
y = torch.empty(E, F_, device=dev)
prof = torch.zeros(32, dtype=torch.int64, device=dev)
P, st = _lib.ptr, _lib.stream()
run = lambda: _lib.check(lib.mdl_linear_tc(P(x), P(w), F_, 1, P(b), P(y), E, F_, F_, 0, st), "l")
run(); torch.cuda.synchronize()
lib.mdl_debug_set_phase_buffer(P(prof)); prof.zero_(); run(); torch.cuda.synchronize(); lib.mdl_debug_set_phase_buffer(None)
v = prof.cpu().tolist(); n = max(v[31], 1)
names = ["loop top", "global loads + split + st issue", "tmem st wait", "hand-off barrier", "wait MMA", "TMEM ld", "bias/act + global stores"]
for i, nm in enumerate(names):
    print(f"  {nm:34s} {v[i] / n:8.0f} cyc/tile")
print("  total", sum(v[:7]) / n, "tiles", n)
