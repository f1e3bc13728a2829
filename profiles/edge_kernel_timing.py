"""Warm CUDA-event timings of the edge-level kernels at the SchNet config-2 shapes (E = 102086, F = 128, G = 50):
fused filter MLP fwd / bwd, weight-gradient kernels, CSR gather-multiply-sum, edge_mul.  L2 flushed between reps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matdeeplearn_b200 import functional as MF, _lib, process as pr
from matdeeplearn_b200.csr import GraphCSR
dev = torch.device("cuda:0")
lib = _lib.load()
ds = pr.synthetic_dataset("bulk", 256, seed=pr.BENCH_SEED)
b = ds.batch().to(dev)
csr = GraphCSR.from_coo(b.edge_index, b.batch, num_graphs=256)
E, N, F_, G = b.edge_index.shape[1], b.x.shape[0], 128, 50
torch.manual_seed(0)
ea = b.edge_attr
w1, b1 = torch.randn(F_, G, device=dev) * 0.1, torch.randn(F_, device=dev) * 0.1
w2, b2 = torch.randn(F_, F_, device=dev) * 0.1, torch.randn(F_, device=dev) * 0.1
rs = torch.rand(E, device=dev)
Y, T1 = torch.empty(E, F_, device=dev), torch.empty(E, F_, device=dev)
g, dp1 = torch.randn(E, F_, device=dev), torch.empty(E, F_, device=dev)
h = torch.randn(N, F_, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
P, st = _lib.ptr, _lib.stream()
dW, db = torch.empty(F_, F_, device=dev), torch.empty(F_, device=dev)
dW1 = torch.empty(F_, G, device=dev)


def timed(name, fn, nbytes=None, flops=None, n=10, cold=True):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        if cold:
            flush.add_(1)
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); c.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(c) * 1e3)
    t = sum(ts) / n
    extra = ""
    if nbytes:
        extra += f"  {nbytes / t / 1e3:7.0f} GB/s"
    if flops:
        extra += f"  {flops / t / 1e6:6.1f} TFLOP/s"
    print(f"{name:44s} {t:8.1f} us ({'cold' if cold else 'warm'} L2){extra}", flush=True)


for cold in (True, False):
    timed("edge_mlp2_fwd (E x 50 -> 128 -> 128)", lambda: _lib.check(lib.mdl_edge_mlp2_fwd(P(ea), P(w1), P(b1), P(w2), P(b2), P(rs), P(Y), P(T1), E, G, F_, F_, 0, 0, st), "f"),
          nbytes=4 * E * (G + 2 * F_ + 1), flops=2.0 * E * (G * F_ + F_ * F_), cold=cold)
    timed("edge_mlp2_fwd, no T1 store", lambda: _lib.check(lib.mdl_edge_mlp2_fwd(P(ea), P(w1), P(b1), P(w2), P(b2), P(rs), P(Y), None, E, G, F_, F_, 0, 0, st), "f"), cold=cold)
    timed("edge_mlp2_fwd, relu (no MUFU)", lambda: _lib.check(lib.mdl_edge_mlp2_fwd(P(ea), P(w1), P(b1), P(w2), P(b2), P(rs), P(Y), P(T1), E, G, F_, F_, 1, 0, st), "f"), cold=cold)
    timed("linear_tc  x W^T [E,128] x [128,128]", lambda: _lib.check(lib.mdl_linear_tc(P(T1), P(w2), F_, 1, P(b2), P(Y), E, F_, F_, 0, st), "l"),
          nbytes=8 * E * F_, flops=2.0 * E * F_ * F_, cold=cold)
    timed("edge_mlp2_bwd", lambda: _lib.check(lib.mdl_edge_mlp2_bwd(P(g), P(rs), P(w2), P(T1), P(dp1), E, F_, F_, 0, st), "b"),
          nbytes=4 * E * (3 * F_ + 1), flops=2.0 * E * F_ * F_, cold=cold)
    for impl in ("tc", "simt"):
        os.environ["MDL_WGRAD"] = impl
        timed(f"linear_wgrad {impl} [E,128]^T [E,128]", lambda: MF.linear_wgrad_into(T1, g, MF._wgrad_map(F_, F_, [dW.data_ptr()], [db.data_ptr()])),
              nbytes=8 * E * F_, flops=2.0 * E * F_ * F_, cold=cold)
        timed(f"linear_wgrad {impl} [E,128]^T [E,50]", lambda: MF.linear_wgrad_into(ea, g, MF._wgrad_map(F_, G, [dW1.data_ptr()], [db.data_ptr()])),
              nbytes=4 * E * (F_ + G), flops=2.0 * E * F_ * G, cold=cold)
        hn, gn = torch.randn(N, F_, device=dev), torch.randn(N, F_, device=dev)
        timed(f"linear_wgrad {impl} [N,128]^T [N,128] (N={N})", lambda: MF.linear_wgrad_into(hn, gn, MF._wgrad_map(F_, F_, [dW.data_ptr()], [db.data_ptr()])),
              nbytes=8 * N * F_, cold=cold)
    os.environ["MDL_WGRAD"] = "tc"
    timed("cuBLAS g^T x [E,128]^T [E,128] + g.sum(0)", lambda: (g.t().mm(T1), g.sum(0)), flops=2.0 * E * F_ * F_, cold=cold)
    timed("cuBLAS x W^T  [E,128] x [128,128]", lambda: torch.nn.functional.linear(T1, w2, b2), flops=2.0 * E * F_ * F_, cold=cold)
    with torch.no_grad():
        timed("spmm_edge (gather h * W -> sum)", lambda: MF.cfconv_aggregate(h, Y, csr), nbytes=4 * E * F_ + 8 * N * F_ + 8 * E, cold=cold)
