#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric   graphs/s of one CGCNN training step (zero_grad + forward + l1_loss +
         backward [+ gradient all-reduce] + AdamW) -- reference train() body,
         matdeeplearn/training/training.py:37-50.
workload configs[1]: CGCNN dim1=dim2=64, pre_fc 1, 4 CGConv layers, post_fc 1,
         synthetic bulk_data-shaped graphs, 256 graphs per GPU
         (weak scaling = the reference's DistributedSampler semantics).

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

_COMMON = dict(pre_fc_count=1, post_fc_count=1, pool="global_mean_pool", pool_order="early", batch_norm="True",
               batch_track_stats="True", act="relu", dropout_rate=0.0)
# BASELINE.json configs[1..4]; `graphs` = graphs per GPU under weak scaling (the reference's per-rank batch_size,
# training.py:291-307), the global batch under --scaling strong
CONFIGS = {
    1: dict(model="CGCNN", kind="bulk", graphs=256, cfg=dict(dim1=64, dim2=64, gc_count=4, **_COMMON),
            what="CGCNN dim=64 4xCGConv, synthetic bulk graphs, batch 256 (configs[1])"),
    2: dict(model="SchNet", kind="bulk", graphs=256,
            cfg=dict(dim1=128, dim2=128, dim3=128, cutoff=8, gc_count=4, **_COMMON),
            what="SchNet dim=128 4 interactions, synthetic bulk graphs, batch 256 (configs[2])"),
    3: dict(model="MEGNet", kind="mof", graphs=64,
            cfg=dict(dim1=128, dim2=128, dim3=128, gc_count=3, gc_fc_count=2, **_COMMON),
            what="MEGNet dim=128 3 blocks (gc_fc 2), synthetic MOF-shaped graphs, batch 64 (configs[3])"),
    4: dict(model="MPNN", kind="bulk", graphs=256, cfg=dict(dim1=64, dim2=64, dim3=64, gc_count=4, **_COMMON),
            sweep=(50, 100, 200),
            what="MPNN dim=64 4xNNConv+GRU, synthetic bulk graphs, batch 256, edge_length 50 (sweep 50/100/200 beside "
                 "it) (configs[4])"),
}
MODEL_CFG = CONFIGS[1]["cfg"]
LR = 0.002  # config.yml *_demo


def workload_string(cfgno, scaling, world):
    """ONE string for both arms (the driver compares them)."""
    c = CONFIGS[cfgno]
    per = "per GPU (weak scaling)" if scaling == "weak" else f"global, split over {world} GPU(s) (strong scaling)"
    return f"{c['what']}; batch {per}"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(CONFIGS), help="BASELINE.json configs[k]")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the config's batch per GPU (reference DistributedSampler semantics); strong: one "
                         "global batch split over the GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (profiling runs)")
    ap.add_argument("--no-roofline", action="store_true", help="skip the isolated-kernel roofline leg")
    ap.add_argument("--roofline-graphs", type=int, default=16384)
    ap.add_argument("--roofline-only", action="store_true", help="only the isolated-kernel leg (ncu captures)")
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-other-configs", action="store_true",
                    help="config 1 only: skip the short train-step measurements of configs 2..4 added to the line")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def tensor_peak_tf32():
    """Dense TF32 tensor-core peak in TFLOP/s: half the bf16 figure (measured cuBLAS bf16 burst in MEASURED_PEAKS.json,
    else the nominal 2250)."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["bf16_tflops"]) / 2.0, "measured bf16_tflops / 2 (MEASURED_PEAKS.json)"
    return 1125.0, "nominal bf16 2250 / 2 (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.rows = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); smax.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_workload(rank, graphs, kind="bulk", edge_length=50):
    from matdeeplearn_b200 import process as pr
    ds = pr.synthetic_dataset(kind, graphs, seed=pr.BENCH_SEED + rank, edge_length=edge_length)
    return ds, ds.batch()


def graphs_per_rank(cfgno, scaling, world):
    g = CONFIGS[cfgno]["graphs"]
    return g if scaling == "weak" else max(1, g // world)


# ------------------------------------------------------------------ CPU leg
def cpu_workload(args, world):
    """The CPU arm's step = the WHOLE job's work: world x (per-rank batch), i.e. the same graphs the N GPUs
    step through together, processed by one host (per-rank batches one after the other, each with the
    reference's per-rank BatchNorm statistics)."""
    c = CONFIGS[args.config]
    per = graphs_per_rank(args.config, args.scaling, world)
    G = c.get("sweep", (50,))[0]
    return [make_workload(r, per, c["kind"], G) for r in range(world)], per


def cpu_step_times(args, world, steps, warmup, threads=None, budget_s=None):
    """budget_s: bound the run -- a step then covers only as many of the rank-batches as fit (a bounded sample of
    the job's graphs; throughput = graphs actually processed / time)."""
    from oracle import models as OM
    c = CONFIGS[args.config]
    shards, per = cpu_workload(args, world)
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = getattr(OM, c["model"])(shards[0][0], **c["cfg"])
    opt = torch.optim.AdamW(model.parameters(), lr=LR)
    model.train()
    times_all = []
    times = []
    for i in range(warmup + steps):
        if budget_s is not None and i == 1 and len(shards) > 1:   # first (warm-up) step timed: cut the sample to fit
            keep = int(budget_s / max(times_all[0] / len(shards), 1e-9) / (warmup + steps))
            shards = shards[:max(1, min(len(shards), keep))]
        t0 = time.perf_counter()
        for _, batch in shards:       # one optimizer step per rank-batch: the same work, serialised on the host
            opt.zero_grad()
            loss = F.l1_loss(model(batch), batch.y)
            loss.backward()
            opt.step()
        times_all.append(time.perf_counter() - t0)
        if i >= warmup:
            times.append(times_all[-1])
    E = sum(int(b.edge_index.shape[1]) for _, b in shards)
    return times, per * len(shards), E


def best_cpu_threads(args, world):
    """Host thread count that maximises the CPU path's throughput; bounded sweep on ONE rank-batch."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores})
    one = argparse.Namespace(**{**vars(args), "scaling": "weak"})
    best, table = None, {}
    for c in cands:
        ts, graphs, _ = cpu_step_times(one, 1, 1, 1, threads=c)
        table[c] = graphs / float(np.mean(ts))
        if best is None or table[c] > table[best]:
            best = c
    return best, table


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores, table = best_cpu_threads(args, world)
    times, graphs, E = cpu_step_times(args, world, args.steps, max(args.warmup, 1), threads=cores, budget_s=150.0)
    ms = 1e3 * float(np.mean(times))
    value = graphs / (ms / 1e3)
    c = CONFIGS[args.config]
    line = {
        "impl": "reference", "metric": f"graphs_per_sec_{c['model'].lower()}_train_step", "value": value,
        "unit": "graphs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": max(args.warmup, 1),
        "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "edges_per_sec": E / (ms / 1e3),
        "config": {"workload": workload_string(args.config, args.scaling, world),
                   "graphs_per_step": graphs, "edges_per_step": E,
                   "note": "reference CPU path = PyG-equivalent op sequence restated in oracle/ "
                           "(torch_geometric/torch_scatter not installable offline); rank 0 only; one step = the "
                           "whole job's graphs (n_gpus x per-rank batch) on this host"},
        "cpu_baseline": {"value": value, "unit": "graphs/s", "cores": cores, "kind": "port",
                         "host_cores": os.cpu_count(), "thread_sweep_graphs_per_s": table,
                         "sample": f"{args.steps} full train steps (zero_grad+fwd+l1_loss+bwd+AdamW) over {graphs} graphs, "
                                   "best thread count"},
        "e2e": {"value": value, "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU leg
def flush_l2(buf):
    buf.add_(1.0)  # 512 MiB read+write > 126 MB L2


def timed_steps(fn, steps, flush_buf, world):
    """K steps, each bracketed by CUDA events on the launching stream, L2 flushed
    (untimed) before each; barrier + synchronize on both sides of the region."""
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_wall = time.perf_counter()
    evs = []
    for _ in range(steps):
        flush_l2(flush_buf)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t_wall
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    return total_ms, wall


def kernel_roofline(args, dev, flush_buf, peak, peak_src):
    """The fused CGConv kernels alone on a working set >> L2 (B=16384 bulk
    graphs built by tiling 1024 synthetic ones), cold L2, CUDA events."""
    from matdeeplearn_b200 import _lib, process as pr
    from matdeeplearn_b200.csr import GraphCSR, gather_rows
    lib = _lib.load()
    base_graphs = min(1024, args.roofline_graphs)
    reps = max(1, args.roofline_graphs // base_graphs)
    ds = pr.synthetic_dataset("bulk", base_graphs, seed=pr.BENCH_SEED + 1000)
    b = ds.batch().to(dev)
    n0, e0 = b.x.shape[0], b.edge_index.shape[1]
    ei = torch.cat([b.edge_index + i * n0 for i in range(reps)], 1).contiguous()
    ea = b.edge_attr.repeat(reps, 1).contiguous()
    batch_vec = torch.cat([b.batch + i * base_graphs for i in range(reps)]).contiguous()
    N, E, C, G = n0 * reps, e0 * reps, MODEL_CFG["dim1"], ea.shape[1]
    csr = GraphCSR.from_coo(ei, batch_vec, num_graphs=base_graphs * reps)
    ea_s = gather_rows(ea, csr.dst_eid)
    del ea
    torch.manual_seed(0)
    x = torch.randn(N, C, device=dev)
    PQ = torch.randn(N, 4 * C, device=dev) * 0.5
    WeT = torch.randn(G, 2 * C, device=dev) * 0.1
    gout = torch.randn(N, C, device=dev)
    out = torch.empty(N, C, device=dev)
    dPQ = torch.empty(N, 4 * C, device=dev)
    dWeT = torch.empty(G, 2 * C, device=dev)
    ws_bytes = lib.mdl_cgconv_workspace_bytes(N, E, C, G)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    st = _lib.stream()
    P = _lib.ptr

    def fwd():
        _lib.check(lib.mdl_cgconv_fwd(P(x), P(PQ), P(ea_s), P(WeT), P(csr.dst_ptr), P(csr.dst_src),
                                      P(csr.dst_dst), P(csr.inv_deg_dst), P(out), N, E, C, G, 1, st), "fwd")

    def bwd():
        _lib.check(lib.mdl_cgconv_bwd(P(gout), P(PQ), P(ea_s), P(WeT), P(csr.dst_ptr), P(csr.dst_src),
                                      P(csr.dst_dst), P(csr.src_ptr), P(csr.src_slot), P(csr.inv_deg_dst),
                                      P(dPQ), P(dWeT), N, E, C, G, 1, P(ws), ws_bytes, st), "bwd")

    def time_it(fn, n=10):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(n):
            flush_l2(flush_buf)
            a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); c.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(c))
        return float(np.mean(ts)), float(np.min(ts))

    # smearing-fused form: the kernels take d_hat [E] (slot order) and expand the Gaussian basis themselves
    dh = torch.rand(E, device=dev)
    offset = torch.linspace(0.0, 1.0, G, device=dev)
    coeff = -0.5 / 0.2 ** 2

    def fwd_fused():
        _lib.check(lib.mdl_cgconv_smear_fwd(P(x), P(PQ), P(dh), P(offset), coeff, P(WeT), P(csr.dst_ptr), P(csr.dst_src),
                                            P(csr.dst_dst), P(csr.inv_deg_dst), P(out), N, E, C, G, 1, st), "fwd fused")

    def bwd_fused():
        _lib.check(lib.mdl_cgconv_smear_bwd(P(gout), P(PQ), P(dh), P(offset), coeff, P(WeT), P(csr.dst_ptr), P(csr.dst_src),
                                            P(csr.dst_dst), P(csr.inv_deg_dst), P(dPQ), P(dWeT), N, E, C, G, 1, P(ws),
                                            ws_bytes, st), "bwd fused")

    f_ms, f_min = time_it(fwd)
    b_ms, b_min = time_it(bwd)
    ff_ms, ff_min = time_it(fwd_fused)
    bf_ms, bf_min = time_it(bwd_fused)
    bytes_fwd = 8 * N * C + 8 * E + 4 * E * G          # SURVEY.md 8d, operator-surface form
    bytes_bwd = bytes_fwd + 4 * N * C                   # whole backward (both passes)
    bytes_fwd_fused = 8 * N * C + 12 * E                # SURVEY.md 8d, smearing-fused form (d_hat instead of edge_attr)
    bytes_bwd_fused = bytes_fwd_fused + 4 * N * C
    res = {
        "workload": f"{base_graphs * reps} bulk graphs (N={N}, E={E}, C={C}, G={G}), cold L2",
        "fwd": {"ms": f_ms, "ms_min": f_min, "algorithmic_bytes": bytes_fwd,
                "achieved_gbs": bytes_fwd / f_ms / 1e6, "frac": bytes_fwd / f_ms / 1e6 / peak},
        "bwd_both_passes": {"ms": b_ms, "ms_min": b_min, "algorithmic_bytes": bytes_bwd,
                            "achieved_gbs": bytes_bwd / b_ms / 1e6, "frac": bytes_bwd / b_ms / 1e6 / peak},
        # the same operator with GaussianSmearing fused into the kernels (mdl_cgconv_smear_fwd / _bwd): the time is
        # what a user gets; its fraction is quoted in BOTH SURVEY 8(d) accountings -- against the bytes the
        # operator-surface form would have moved (208 B/edge) and against its own 12 B/edge
        "fwd_smear_fused": {"ms": ff_ms, "ms_min": ff_min, "algorithmic_bytes_fused_form": bytes_fwd_fused,
                            "frac_fused_form": bytes_fwd_fused / ff_ms / 1e6 / peak,
                            "frac_operator_surface_form": bytes_fwd / ff_ms / 1e6 / peak},
        "bwd_smear_fused": {"ms": bf_ms, "ms_min": bf_min, "algorithmic_bytes_fused_form": bytes_bwd_fused,
                            "frac_fused_form": bytes_bwd_fused / bf_ms / 1e6 / peak,
                            "frac_operator_surface_form": bytes_bwd / bf_ms / 1e6 / peak},
        "edge_flops_fwd": 2.0 * E * G * 2 * C, "fwd_tflops_fp32": 2.0 * E * G * 2 * C / f_ms / 1e9,
    }
    return res


def aux_roofline(args, dev, flush_buf, peak):
    """configs[2..4]: the CSR gather -> message -> segment-sum kernel of the config's operator, alone, on a
    working set >> L2, cold L2, CUDA events (forward launch).  Algorithmic bytes: SURVEY.md 8(d) per-edge /
    per-node figures of that kernel's own inputs and outputs (stated in `form`)."""
    from matdeeplearn_b200 import functional as MF, process as pr
    from matdeeplearn_b200.csr import GraphCSR
    c = CONFIGS[args.config]
    base = 256 if c["kind"] == "bulk" else 32
    reps = 32 if c["kind"] == "bulk" else 16
    ds = pr.synthetic_dataset(c["kind"], base, seed=pr.BENCH_SEED + 1000)
    b = ds.batch().to(dev)
    n0 = b.x.shape[0]
    ei = torch.cat([b.edge_index + i * n0 for i in range(reps)], 1).contiguous()
    bv = torch.cat([b.batch + i * base for i in range(reps)]).contiguous()
    N, E = n0 * reps, ei.shape[1]
    csr = GraphCSR.from_coo(ei, bv, num_graphs=base * reps)
    torch.manual_seed(0)
    if args.config == 2:
        Fw = 128
        h, W = torch.randn(N, Fw, device=dev), torch.randn(E, Fw, device=dev)
        fn = lambda: MF.cfconv_aggregate(h, W, csr)                       # noqa: E731
        nbytes, kernel = 4 * E * Fw + 8 * N * Fw + 8 * E, "k_spmm_edge (CFConv gather * filter -> destination sum)"
        form = "4*E*F (filter rows) + 4*N*F (h) + 4*N*F (out) + 8*E (indices), F=128"
    elif args.config == 3:
        D = 128
        base_t, A, B_ = torch.randn(E, D, device=dev), torch.randn(N, D, device=dev), torch.randn(N, D, device=dev)
        U, bias = torch.randn(base * reps, D, device=dev), torch.randn(D, device=dev)
        fn = lambda: MF.edge_gather_add(base_t, A, B_, U, bias, ei, bv, csr, True)   # noqa: E731
        nbytes, kernel = 8 * E * D + 8 * N * D + 16 * E, "k_edge_gather_add (MEGNet edge update, first layer)"
        form = "4*E*D read + 4*E*D write + 2*4*N*D (node projections) + 16*E (int64 indices), D=128"
    else:
        K, O = 64, 64
        hid, XT, XB = torch.randn(E, K, device=dev), torch.randn(N, K * O, device=dev), torch.randn(N, O, device=dev)
        fn = lambda: MF.nnconv_message(hid, XT, XB, csr)                  # noqa: E731
        nbytes, kernel = 4 * E * K + 4 * E * O + 4 * N * (K * O + O) + 4 * E, "k_nnconv_msg (NNConv re-associated message)"
        form = "4*E*K (hidden) + 4*E*O (messages) + 4*N*(K*O+O) (per-node products) + 4*E (edge ids), K=O=64"
    def time_it(f):
        with torch.no_grad():
            for _ in range(3):
                f()
            ts = []
            for _ in range(10):
                flush_l2(flush_buf)
                a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); f(); e.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(e))
        return float(np.mean(ts))

    ms = time_it(fn)
    res = {"kernel": kernel, "workload": f"{base * reps} {c['kind']} graphs (N={N}, E={E}), cold L2", "ms": ms,
           "algorithmic_bytes": nbytes, "form": form, "achieved_gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}
    if args.config == 2:
        # the filter network + cutoff of the interaction (k_edge_mlp2_fwd, tcgen05 3xTF32): tensor-bound
        G = b.edge_attr.shape[1]
        ea = b.edge_attr.repeat(reps, 1).contiguous()
        w1, b1 = torch.randn(Fw, G, device=dev) * 0.1, torch.randn(Fw, device=dev) * 0.1
        w2, b2 = torch.randn(Fw, Fw, device=dev) * 0.1, torch.randn(Fw, device=dev) * 0.1
        rs = torch.rand(E, device=dev)
        ms2 = time_it(lambda: MF.edge_mlp2(ea, w1, b1, w2, b2, rs, "ssp"))
        flops = 2.0 * E * (G * Fw + Fw * Fw)
        tpeak, tsrc = tensor_peak_tf32()
        res["filter_mlp"] = {"kernel": "k_edge_mlp2_fwd (SchNet filter network + cosine cutoff, two chained 3xTF32 tcgen05 "
                                       "contractions per 128-edge round)", "ms": ms2, "algorithmic_flops": flops,
                             "achieved_tflops": flops / ms2 / 1e9, "peak_tflops_tf32": tpeak, "peak_source": tsrc,
                             "frac": flops / ms2 / 1e9 / tpeak,
                             "note": "3xTF32 issues 3 tensor-core products per algorithmic one; fraction of the tf32 "
                                     "pipe actually busy is 3x this",
                             "algorithmic_bytes": 4 * E * G + 4 * E + 4 * E * Fw,
                             "achieved_gbs": (4 * E * G + 4 * E + 4 * E * Fw) / ms2 / 1e6}
    return res


def quick_step(cfgno, rank, world, dev, flush_buf, steps=10):
    """Short device-timed measurement of another BASELINE config's train step (same rules as the headline value:
    graph replay on a resident batch, L2 flushed between steps, CUDA events, max over ranks)."""
    from matdeeplearn_b200 import models as M
    from matdeeplearn_b200.engine import TrainStep
    c = CONFIGS[cfgno]
    per = c["graphs"]
    ds, hb = make_workload(rank, per, c["kind"], c.get("sweep", (50,))[0])
    hb.num_graphs = per
    torch.manual_seed(0)
    model = getattr(M, c["model"])(ds, **c["cfg"]).to(dev)
    model.train()
    step = TrainStep(model, lr=LR * world)
    db = hb.to(dev)
    db.num_graphs = per
    replay = step.resident(db, warmup=3)
    for _ in range(3):
        replay()
    ms, _ = timed_steps(replay, steps, flush_buf, world)
    out = {"workload": workload_string(cfgno, "weak", world), "steps": steps, "ms_per_step": ms / steps,
           "value": per * world / (ms / steps / 1e3), "unit": "graphs/s", "nodes_per_gpu": int(hb.x.shape[0]),
           "edges_per_gpu": int(hb.edge_index.shape[1]), "kernels_per_step": int(step.kernels_per_step)}
    del model, step, replay, db
    torch.cuda.empty_cache()
    return out


def run_engine(args, rank, world, local_rank):
    from matdeeplearn_b200 import _lib, models as M
    from matdeeplearn_b200.engine import TrainStep
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.load()
    peak, peak_src = peaks()
    if args.roofline_only:
        flush_buf = torch.zeros(128 * 1024 * 1024, device=dev)
        print(json.dumps(kernel_roofline(args, dev, flush_buf, peak, peak_src)), flush=True)
        return
    c = CONFIGS[args.config]
    per = graphs_per_rank(args.config, args.scaling, world)
    sweep = c.get("sweep", (50,))
    flush_buf = torch.zeros(128 * 1024 * 1024, device=dev)  # 512 MiB
    sampler = ClockSampler(local_rank)
    head, sweep_out = None, {}
    for G in sweep:
        ds, host_batch = make_workload(rank, per, c["kind"], G)
        host_batch.num_graphs = per
        N, E = host_batch.x.shape[0], host_batch.edge_index.shape[1]
        torch.manual_seed(0)
        model = getattr(M, c["model"])(ds, **c["cfg"]).to(dev)
        model.train()
        step = TrainStep(model, lr=LR * world)     # broadcasts rank 0's parameters (DDP constructor semantics)
        # ---- device-resident, graph-replayed step (value)
        dev_batch = host_batch.to(dev)
        dev_batch.num_graphs = per
        replay = step.resident(dev_batch, warmup=3)
        for _ in range(max(args.warmup, 3)):
            replay()
        first = head is None
        if first and rank == 0:
            sampler.start()
        total_ms, wall = timed_steps(replay, args.steps, flush_buf, world)
        e_total = torch.tensor([float(E)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(e_total)
        rec = {"G": G, "ms_per_step": total_ms / args.steps, "graphs_per_s": per * world / (total_ms / args.steps / 1e3),
               "edges_per_s": float(e_total.item()) / (total_ms / args.steps / 1e3), "nodes_per_gpu": N, "edges_per_gpu": E,
               "kernels_per_step": int(step.kernels_per_step)}
        sweep_out[str(G)] = rec
        if not first:
            del model, step, replay
            torch.cuda.empty_cache()
            continue
        head = dict(rec, wall=wall, graph_allreduce=bool(getattr(step, "graph_allreduce", False)))
        # ---- end to end from pinned host buffers through the public API (e2e)
        pinned = host_batch.pin_memory()
        pinned.num_graphs = per
        pinned.smear = getattr(host_batch, "smear", None)
        e2e_steps = max(5, min(args.steps, 30))
        graph_ok = True
        try:
            for _ in range(3):
                step.from_host(pinned, expand_edge_attr=False)
        except Exception as exc:  # a model whose step cannot be captured from host buffers: eager step, said so below
            graph_ok = False
            head["e2e_capture_error"] = repr(exc)[:200]
            torch.cuda.synchronize()
        call = (lambda **kw: step.from_host(pinned, **kw)) if graph_ok else (lambda **kw: step.from_host(pinned, use_graph=False))
        # (i) the batch's 7 reference tensors copied as they are (edge_attr materialised on the host)
        for _ in range(2):
            call(expand_edge_attr=False)
        e2e_full_ms, _ = timed_steps(lambda: call(expand_edge_attr=False), e2e_steps, flush_buf, world)
        h2d_full = getattr(step, "last_h2d_bytes", 0)
        # (ii) product extra: normalised distances shipped, Gaussian basis expanded on the GPU
        for _ in range(3):
            call()
        e2e_ms, _ = timed_steps(lambda: call(), e2e_steps, flush_buf, world)
        h2d = getattr(step, "last_h2d_bytes", 0)
        head.update(e2e_steps=e2e_steps, e2e_full_ms=e2e_full_ms, e2e_ms=e2e_ms, h2d_full=h2d_full, h2d=h2d, graph_ok=graph_ok)
        # (iii) dataset resident in HBM (store.GraphStore), batch assembled on the GPU into capacity-padded buffers,
        # one graph replay per step (every model family)
        try:
            from matdeeplearn_b200.store import GraphStore
            store = GraphStore.from_dataset(ds, dev)
            perm_rng = np.random.default_rng(1234 + rank)
            perms = [perm_rng.permutation(per) for _ in range(e2e_steps + 3)]
            for i in range(3):
                step.from_store(store, perms[i])
            it = iter(perms[3:])
            store_ms, _ = timed_steps(lambda: step.from_store(store, next(it)), e2e_steps, flush_buf, world)
            head["store_ms"] = store_ms
            del store
        except Exception as exc:
            head["store_error"] = repr(exc)[:200]
            torch.cuda.synchronize()
        head["clocks"] = sampler.stop() if rank == 0 else None
        del model, step, replay
        torch.cuda.empty_cache()

    strong = None
    if world > 1 and args.scaling == "weak":
        # BASELINE.md section 4 quotes strong scaling beside the weak one: the config's batch as ONE global batch split
        # over the GPUs (whole graphs per rank), same step, same timing rules
        per_s = max(1, c["graphs"] // world)
        ds_s, hb_s = make_workload(rank, per_s, c["kind"], sweep[0])
        hb_s.num_graphs = per_s
        torch.manual_seed(0)
        model_s = getattr(M, c["model"])(ds_s, **c["cfg"]).to(dev)
        model_s.train()
        step_s = TrainStep(model_s, lr=LR * world)
        db_s = hb_s.to(dev)
        db_s.num_graphs = per_s
        replay_s = step_s.resident(db_s, warmup=3)
        for _ in range(max(args.warmup, 3)):
            replay_s()
        ms_s, _ = timed_steps(replay_s, args.steps, flush_buf, world)
        strong = {"graphs_per_step": per_s * world, "graphs_per_gpu": per_s, "ms_per_step": ms_s / args.steps,
                  "value": per_s * world / (ms_s / args.steps / 1e3), "unit": "graphs/s",
                  "note": "strong scaling: the config's batch split over the GPUs (the headline `value` is weak scaling)"}
        del model_s, step_s, replay_s
        torch.cuda.empty_cache()

    graphs_total = per * world
    ms_per_step = head["ms_per_step"]
    es = head["e2e_steps"]
    # the headline e2e is the reference-layout call (all 7 tensors of the PyG batch copied, edge_attr [E,G] included)
    e2e_value = graphs_total / (head["e2e_full_ms"] / es / 1e3)
    how = ("one CUDA graph replay of CSR build + slot permute + fwd + bwd" + ("" if world == 1 else " + NCCL all-reduce")
           + " + AdamW") if head["graph_ok"] else "eager step (capture from host buffers failed for this model)"
    line = {
        "metric": f"graphs_per_sec_{c['model'].lower()}_train_step", "value": head["graphs_per_s"], "unit": "graphs/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "edges_per_sec": head["edges_per_s"],
        "config": {
            "workload": workload_string(args.config, args.scaling, world),
            "graphs_per_step": graphs_total, "nodes_per_gpu": head["nodes_per_gpu"], "edges_per_gpu": head["edges_per_gpu"],
            "parallelism": f"dp{world}", "step": "zero_grad+fwd+l1_loss+bwd" +
            (("+flat grad allreduce (NCCL, captured inside the step's CUDA graph)+AdamW, one CUDA graph replay"
              if head["graph_allreduce"] else
              "+flat grad allreduce (NCCL, eager)+AdamW, two CUDA graph replays around the collective")
             if world > 1 else "+AdamW, one CUDA graph replay") +
            " (the metric text says fwd+bwd; the optimizer step is included, as in the reference's train() body)",
            "l2": "flushed between timed steps (512 MiB read-modify-write)",
            "timing": "sum of per-step CUDA-event durations, max over ranks",
        },
        "gpu_launches": head["kernels_per_step"] * args.steps,
        "gpu_launches_per_step": head["kernels_per_step"],
        "gpu_launches_note": "kernels launched by libmdl_b200.so inside the captured step (mdl_launch_count); library "
                             "(cuBLAS / ATen) launches in the same step are not counted",
        "e2e": {"value": e2e_value, "unit": "graphs/s", "h2d_bytes_per_step": int(head["h2d_full"]),
                "d2h_bytes_per_step": 4, "steps": es,
                "path": "TrainStep.from_host(pinned Batch, expand_edge_attr=False): H2D of the 7 tensors of the "
                        "reference's batch (x, edge_index, edge_attr [E,G], edge_weight, batch, u, y), then " + how +
                        ", then loss.item()",
                "distances_shipped": {
                    "value": graphs_total / (head["e2e_ms"] / es / 1e3), "h2d_bytes_per_step": int(head["h2d"]),
                    "path": "product extra (batches made by process.assemble_dataset carry d_hat): 4 B/edge shipped, "
                            "GaussianSmearing (reference process.py:580-590) expanded on the GPU inside the graph"}},
        "wall_s_timed_region": head["wall"],
    }
    if "e2e_capture_error" in head:
        line["e2e"]["capture_error"] = head["e2e_capture_error"]
    if "store_ms" in head:
        line["store_step"] = {"value": graphs_total / (head["store_ms"] / es / 1e3), "unit": "graphs/s", "steps": es,
                              "path": "TrainStep.from_store(GraphStore, idx): dataset resident in HBM, batch = index "
                                      "list -> one assembly kernel into capacity-padded buffers + the step, one CUDA "
                                      "graph replay; loss.item() each step"}
    if strong is not None:
        line["strong_scaling"] = strong
    if args.config == 1 and world == 1 and not args.no_other_configs:
        # the other BASELINE configs, short device-timed runs (their full lines: bench.py --config 2|3|4)
        others = {}
        for k in (2, 3, 4):
            try:
                others[f"configs[{k}]"] = quick_step(k, rank, world, dev, flush_buf)
            except Exception as exc:
                others[f"configs[{k}]"] = {"error": repr(exc)[:200]}
                torch.cuda.synchronize()
        line["other_configs"] = others
    if "store_error" in head:
        line["store_step"] = {"error": head["store_error"]}
    if len(sweep) > 1:
        line["edge_length_sweep"] = sweep_out
        line["epoch_100k_graphs_s"] = {g: 100000.0 / r["graphs_per_s"] for g, r in sweep_out.items()}
    if rank == 0:
        line["clocks"] = head["clocks"]
        if not args.no_roofline:
            if args.config == 1:
                r = kernel_roofline(args, dev, flush_buf, peak, peak_src)
                f, bw = r["fwd"], r["bwd_both_passes"]
                line["roofline"] = {"bound": "hbm", "achieved": f["achieved_gbs"], "peak": peak, "unit": "GB/s",
                                    "frac": f["frac"], "traffic": None,
                                    "kernel": "k_cgconv_fwd_ws (fused gather -> message -> scatter_add, the kernel the "
                                              "north star's roofline target names)",
                                    "fwd_frac": f["frac"], "bwd_frac": bw["frac"],
                                    "bwd": {"kernel": "k_cgconv_bwd_pipe (whole CGConv backward: single pass, dP + dW_e + dQ)",
                                            "achieved": bw["achieved_gbs"], "frac": bw["frac"], "ms": bw["ms"]},
                                    "form": "operator-surface: 8NC + 8E + 4EG bytes forward, + 4NC backward (SURVEY.md 8d)",
                                    "smear_fused": {"fwd": r["fwd_smear_fused"], "bwd": r["bwd_smear_fused"],
                                                    "form": "kernels take d_hat [E] and expand GaussianSmearing inside: "
                                                            "8NC + 12E bytes forward"},
                                    "traffic_note": "dram bytes per launch are in profiles/ (ncu --set full), not re-measured here",
                                    "peak_source": peak_src, "workload": r["workload"]}
                line["roofline_detail"] = r
            else:
                r = aux_roofline(args, dev, flush_buf, peak)
                line["roofline"] = {"bound": "hbm", "achieved": r["achieved_gbs"], "peak": peak, "unit": "GB/s",
                                    "frac": r["frac"], "traffic": None, "kernel": r["kernel"], "form": r["form"],
                                    "peak_source": peak_src, "workload": r["workload"], "ms": r["ms"]}
                if "filter_mlp" in r:
                    line["roofline"]["filter_mlp"] = r["filter_mlp"]
        if not args.no_cpu_baseline and world == 1:
            threads, table = best_cpu_threads(args, 1)
            ts, graphs, _ = cpu_step_times(args, 1, args.cpu_steps, 1, threads=threads)
            v = graphs / float(np.mean(ts))
            line["cpu_baseline"] = {"value": v, "unit": "graphs/s", "cores": threads, "kind": "port",
                                    "host_cores": os.cpu_count(), "thread_sweep_graphs_per_s": table,
                                    "sample": f"{args.cpu_steps} full train steps of the same {graphs}-graph batch "
                                              "(oracle = PyG-equivalent op sequence, torch CPU), best thread count"}
        print(json.dumps(line), flush=True)


def main():
    args = parse()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local_rank))
    run_engine(args, rank, world, local_rank)
    if world > 1:
        # Leave without tearing the communicator down: destroy_process_group() was observed to hang while CUDA graphs
        # holding captured NCCL kernels are alive (2-GPU run, profiles/r2_multi_gpu_notes.txt).  Every rank has
        # finished its work and rank 0 has printed the line.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
