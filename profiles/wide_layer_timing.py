"""CGConv layers wider than 64 channels: the tensor-core kernels (64-channel chunks of cgconv_fwd_ws.cu / cgconv_bwd.cu)
against the SIMT kernels (MDL_CGCONV_IMPL=simt) on 4096 bulk graphs, cold L2, CUDA events, forward and backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matdeeplearn_b200 import _lib, process as pr
from matdeeplearn_b200.csr import GraphCSR, gather_rows

lib = _lib.load()
dev = torch.device("cuda:0")
ds = pr.synthetic_dataset("bulk", 1024, seed=7)
b = ds.batch().to(dev)
reps = 4
n0 = b.x.shape[0]
ei = torch.cat([b.edge_index + i * n0 for i in range(reps)], 1).contiguous()
ea = b.edge_attr.repeat(reps, 1).contiguous()
N, E, G = n0 * reps, ei.shape[1], ea.shape[1]
csr = GraphCSR.from_coo(ei, num_nodes=N)
ea_s = gather_rows(ea, csr.dst_eid)
flush = torch.zeros(128 * 1024 * 1024, device=dev)
P, st = _lib.ptr, _lib.stream()
for C in (64, 100, 128):
    x = torch.randn(N, C, device=dev); PQ = torch.randn(N, 4 * C, device=dev) * 0.5
    WeT = torch.randn(G, 2 * C, device=dev) * 0.1; gout = torch.randn(N, C, device=dev)
    out = torch.empty(N, C, device=dev); dPQ = torch.empty(N, 4 * C, device=dev); dWeT = torch.empty(G, 2 * C, device=dev)
    wsb = lib.mdl_cgconv_workspace_bytes(N, E, C, G); ws = torch.empty(wsb, dtype=torch.uint8, device=dev)

    def fwd():
        _lib.check(lib.mdl_cgconv_fwd(P(x), P(PQ), P(ea_s), P(WeT), P(csr.dst_ptr), P(csr.dst_src), P(csr.dst_dst),
                                      P(csr.inv_deg_dst), P(out), N, E, C, G, 1, st), "fwd")

    def bwd():
        _lib.check(lib.mdl_cgconv_bwd(P(gout), P(PQ), P(ea_s), P(WeT), P(csr.dst_ptr), P(csr.dst_src), P(csr.dst_dst),
                                      P(csr.src_ptr), P(csr.src_slot), P(csr.inv_deg_dst), P(dPQ), P(dWeT), N, E, C, G,
                                      1, P(ws), wsb, st), "bwd")

    for impl in ("default", "simt"):
        if impl == "simt":
            os.environ["MDL_CGCONV_IMPL"] = "simt"
        else:
            os.environ.pop("MDL_CGCONV_IMPL", None)
        res = []
        for fn in (fwd, bwd):
            for _ in range(2):
                fn()
            ts = []
            for _ in range(5):
                flush.add_(1.0)
                a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); fn(); c.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(c))
            res.append(sum(ts) / len(ts))
        print(f"C={C:4d} G={G} N={N} E={E} impl={impl:8s} tc_supported={lib.mdl_cgconv_tc_supported(C, G)}  fwd {res[0]:.3f} ms  bwd {res[1]:.3f} ms")
os.environ.pop("MDL_CGCONV_IMPL", None)
