"""A/B on the GPU: CGConv forward, slot-major tensor-core kernel (tc) vs transposed-tile kernel (tt).
Usage: python profiles/tt_ab.py [graphs]"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matdeeplearn_b200 import _lib, process as pr
from matdeeplearn_b200.csr import GraphCSR, gather_rows
lib = _lib.load(); dev = torch.device("cuda:0")
graphs = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
base = min(graphs, 1024); reps = max(1, graphs // base)
ds = pr.synthetic_dataset("bulk", base, seed=7); b = ds.batch().to(dev); n0 = b.x.shape[0]
ei = torch.cat([b.edge_index + i * n0 for i in range(reps)], 1).contiguous()
ea = b.edge_attr.repeat(reps, 1).contiguous()
N, E, C, G = n0 * reps, ei.shape[1], 64, ea.shape[1]
csr = GraphCSR.from_coo(ei, num_nodes=N); ea_s = gather_rows(ea, csr.dst_eid)
torch.manual_seed(0)
x = torch.randn(N, C, device=dev); PQ = torch.randn(N, 4 * C, device=dev) * 0.5
WeT = torch.randn(G, 2 * C, device=dev) * 0.1
P, st = _lib.ptr, _lib.stream()
flush = torch.zeros(64 * 1024 * 1024, device=dev)
prof = torch.zeros(16, dtype=torch.int64, device=dev)
NAMES = ["loop", "S1 wait rows", "split + S2", "prefetch batch 0", "wait MMA", "epilogue", "S3 + carry", "",
         "P:loop", "P:S1 (cp.async wait)", "P:S2", "P:MMA issue", "P:empty segs + next idx", "P:wait MMA",
         "P:rows cp.async issue"]
res = {}
for impl in ("tc", "tt"):
    os.environ["MDL_CGCONV_IMPL"] = impl
    out = torch.full((N, C), float("nan"), device=dev)
    def fwd():
        _lib.check(lib.mdl_cgconv_fwd(P(x), P(PQ), P(ea_s), P(WeT), P(csr.dst_ptr), P(csr.dst_src), P(csr.dst_dst),
                                      P(csr.inv_deg_dst), P(out), N, E, C, G, 1, st), "fwd")
    for _ in range(3): fwd()
    ts = []
    for _ in range(10):
        flush.add_(1.0)
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fwd(); c.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(c))
    res[impl] = out.clone()
    line = {"impl": impl, "N": N, "E": E, "ms_median": sorted(ts)[len(ts) // 2], "ms_min": min(ts)}
    if impl == "tt":
        lib.mdl_debug_set_phase_buffer(P(prof)); prof.zero_(); fwd(); torch.cuda.synchronize()
        lib.mdl_debug_set_phase_buffer(None)
        v = prof.cpu().tolist(); rounds = max(v[15], 1)
        line["cycles_per_round"] = {NAMES[i]: round(v[i] / rounds) for i in range(15) if NAMES[i]}
        line["rounds"] = rounds
    print(json.dumps(line), flush=True)
d = (res["tc"] - res["tt"]).abs()
print(json.dumps({"max_abs_diff_tc_vs_tt": d.max().item(), "nan_in_tt": bool(torch.isnan(res["tt"]).any()),
                  "rows_differing_1e-4": int((d.max(1).values > 1e-4).sum())}))
