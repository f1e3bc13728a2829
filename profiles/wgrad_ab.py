"""A/B on the GPU: mdl_linear_wgrad against the library path (g.t().mm(x) + g.sum(0)) at the step's shapes."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matdeeplearn_b200 import functional as MF
dev = torch.device("cuda", 0)
def t(fn, reps=200):
    for _ in range(10): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for N, I, O in [(7862, 64, 256), (7862, 114, 64), (256, 64, 64), (256, 64, 1), (500000, 64, 256)]:
    x, g = torch.randn(N, I, device=dev), torch.randn(N, O, device=dev)
    lib_us = t(lambda: (g.t().mm(x), g.sum(0)))
    own_us = t(lambda: MF.linear_wgrad(x, g))
    dW, db = MF.linear_wgrad(x, g)
    ref = g.double().t().mm(x.double())
    err = (dW.double() - ref).abs().max().item() / ref.abs().max().item()
    errb = (db.double() - g.double().sum(0)).abs().max().item()
    print(json.dumps({"N": N, "I": I, "O": O, "library_us": lib_us, "mdl_linear_wgrad_us": own_us,
                      "rel_err_dW": err, "abs_err_db": errb}), flush=True)
