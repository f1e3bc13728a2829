"""Per-phase cycle breakdown of the tensor-core CGConv kernels (thread 0 of every
CTA; phases are separated by CTA barriers so this is the CTA's timeline).
Usage on the GPU box:  python profiles/phase_profile.py [graphs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matdeeplearn_b200 import _lib, process as pr  # noqa: E402
from matdeeplearn_b200.csr import GraphCSR, gather_rows  # noqa: E402

NAMES = {0: "loop top", 1: "S1 (fence + barrier)", 16: "next idx issue", 17: "node rows bulk issue",
         24: "wait ea rows", 18: "split hi/lo", 19: "proxy fence", 2: "S2 barrier", 3: "MMA issue (+ next ea bulk)",
         4: "seg/grad loads issue", 21: "next idx land (+ ea rows cp.async)", 20: "wait node rows", 6: "wait MMA",
         22: "TMEM ld", 23: "node terms add", 13: "S2d barrier", 7: "gate math", 8: "S3 barrier",
         11: "scale pass (bwd_src)", 12: "dQ atomics (bwd)", 9: "reduce", 10: "dWe"}
lib = _lib.load()
dev = torch.device("cuda:0")
graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
base = min(graphs, 1024)
reps = max(1, graphs // base)
ds = pr.synthetic_dataset("bulk", base, seed=7)
b = ds.batch().to(dev)
n0 = b.x.shape[0]
ei = torch.cat([b.edge_index + i * n0 for i in range(reps)], 1).contiguous()
ea = b.edge_attr.repeat(reps, 1).contiguous()
N, E, C, G = n0 * reps, ei.shape[1], 64, ea.shape[1]
csr = GraphCSR.from_coo(ei, num_nodes=N)
ea_s = gather_rows(ea, csr.dst_eid)
x = torch.randn(N, C, device=dev)
PQ = torch.randn(N, 4 * C, device=dev) * 0.5
WeT = torch.randn(G, 2 * C, device=dev) * 0.1
gout = torch.randn(N, C, device=dev)
out = torch.empty(N, C, device=dev)
dPQ = torch.empty(N, 4 * C, device=dev)
dWeT = torch.empty(G, 2 * C, device=dev)
wsb = lib.mdl_cgconv_workspace_bytes(N, E, C, G)
ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
prof = torch.zeros(32, dtype=torch.int64, device=dev)
P, st = _lib.ptr, _lib.stream()


def fwd():
    _lib.check(lib.mdl_cgconv_fwd(P(x), P(PQ), P(ea_s), P(WeT), P(csr.dst_ptr), P(csr.dst_src), P(csr.dst_dst),
                                  P(csr.inv_deg_dst), P(out), N, E, C, G, 1, st), "fwd")


def bwd():
    _lib.check(lib.mdl_cgconv_bwd(P(gout), P(PQ), P(ea_s), P(WeT), P(csr.dst_ptr), P(csr.dst_src), P(csr.dst_dst),
                                  P(csr.src_ptr), P(csr.src_slot), P(csr.inv_deg_dst), P(dPQ), P(dWeT), N, E, C, G,
                                  1, P(ws), wsb, st), "bwd")


PIPE_NAMES = {0: "loop top", 1: "S1 barrier", 13: "next idx issue + ranges", 14: "row loads issue", 5: "seg loads issue",
              6: "wait MMA(r)", 2: "wait ea(r+1)", 3: "split(r+1) -> TMEM", 4: "row+idx STS, fences, S2",
              7: "TMEM ld", 10: "node terms + gate math", 11: "S3 barrier", 12: "reduce"}
WS_NAMES = {0: "loop top + seg loads issue", 1: "wait indices / node rows", 2: "wait MMA", 3: "TMEM ld + release",
            4: "node terms + gate math", 5: "S3 barrier", 6: "reduce", 7: "S1 barrier"}
BWD_PIPE_NAMES = {0: "loop top", 1: "S1 barrier", 2: "next idx, window, row/seg loads issue", 3: "wait dW_e(r-1)",
                  4: "wait ea + split -> TMEM + ea^T tiles", 5: "row/idx STS, fences", 6: "S2 (issuer hand-off)",
                  7: "grad loads + wait recompute MMA", 8: "TMEM ld + node terms", 9: "S2d barrier", 10: "gate math",
                  11: "S3 barrier", 12: "da^T -> TMEM", 13: "S4 (issuer hand-off)", 14: "dQ atomics", 15: "dP sums"}
for name, fn in (("fwd", fwd), ("bwd", bwd)):
    if os.environ.get("ONLY", name) != name:
        continue
    if name == "fwd" and os.environ.get("MDL_CGCONV_IMPL", "ws") == "ws":
        NAMES_ = WS_NAMES
    elif name == "fwd" and os.environ.get("MDL_CGCONV_IMPL", "ws") == "pipe":
        NAMES_ = PIPE_NAMES
    elif name == "bwd" and os.environ.get("MDL_CGCONV_BWD", "pipe") != "tc":
        NAMES_ = BWD_PIPE_NAMES
    else:
        NAMES_ = NAMES
    fn(); torch.cuda.synchronize()
    lib.mdl_debug_set_phase_buffer(P(prof))
    prof.zero_()
    a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); c.record(); torch.cuda.synchronize()
    lib.mdl_debug_set_phase_buffer(None)
    v = prof.cpu().tolist()
    rounds = max(v[31], 1)
    if v[25] and NAMES_ is NAMES:
        print(f"!! wait timed out: barrier id {v[26]} (1 ea, 2 node rows, 3 mma), CTA {v[27]}, round {v[28]}, "
              f"thread {v[29]}, parity {v[30]}")
    tot = max(sum(v[i] for i, nm in NAMES_.items() if not nm.startswith("(")), 1)
    print(f"== {name}: N={N} E={E}  {a.elapsed_time(c):.3f} ms, {rounds} rounds, {tot / rounds:.0f} cycles/round")
    for i, nm in NAMES_.items():  # in program order
        if v[i]:
            print(f"   {nm:34s} {v[i] / rounds:9.0f} cyc/round  {100 * v[i] / tot:5.1f}%")

if os.environ.get("PHASE_ONLY"):
    sys.exit(0)
# A/B of the kernel switches on the same inputs: cold L2 (512 MiB flush), CUDA events, mean of 10
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=10):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.add_(1)
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(c))
    return sum(ts) / n, min(ts)


bytes_fwd = 8 * N * C + 8 * E + 4 * E * G
for impl, win, gate, ea_mode in (("ws", "1", "mixed", "bulk"), ("ws", "0", "mixed", "bulk"), ("pipe", "1", "mixed", "bulk"),
                                 ("tc", "1", "mixed", "bulk")):
    os.environ["MDL_CGCONV_IMPL"] = impl
    os.environ["MDL_CGCONV_BWD"] = impl
    os.environ["MDL_CGCONV_WINDOW"] = win
    os.environ["MDL_CGCONV_GATE"] = gate
    os.environ["MDL_CGCONV_EA"] = ea_mode
    f_ms, f_min = timed(fwd)
    b_ms, b_min = timed(bwd)
    print(f"A/B impl={impl} window={win} gate={gate} ea={ea_mode}: fwd {f_ms:.3f} ms (min {f_min:.3f}, {bytes_fwd / f_ms / 1e6:.0f} GB/s algorithmic)"
          f"  bwd {b_ms:.3f} ms (min {b_min:.3f})")
