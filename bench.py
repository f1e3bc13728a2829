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

MODEL_CFG = dict(dim1=64, dim2=64, pre_fc_count=1, gc_count=4, post_fc_count=1,
                 pool="global_mean_pool", pool_order="early", batch_norm="True",
                 batch_track_stats="True", act="relu", dropout_rate=0.0)
GRAPHS_PER_GPU = 256
LR = 0.002  # config.yml CGCNN_demo


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (profiling runs)")
    ap.add_argument("--no-roofline", action="store_true", help="skip the isolated-kernel roofline leg")
    ap.add_argument("--roofline-graphs", type=int, default=16384)
    ap.add_argument("--roofline-only", action="store_true", help="only the isolated-kernel leg (ncu captures)")
    ap.add_argument("--cpu-steps", type=int, default=8)
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


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


def make_workload(rank, graphs):
    from matdeeplearn_b200 import process as pr
    ds = pr.synthetic_dataset("bulk", graphs, seed=pr.BENCH_SEED + rank)
    return ds, ds.batch()


# ------------------------------------------------------------------ CPU leg
def cpu_reference_steps(ds, batch, steps, warmup):
    """The reference's CPU path: oracle restatement of the PyG op sequence on all
    host cores (the reference itself cannot be installed: BASELINE.md section 2)."""
    from oracle import models as OM
    torch.manual_seed(0)
    model = OM.CGCNN(ds, **MODEL_CFG)
    opt = torch.optim.AdamW(model.parameters(), lr=LR)
    model.train()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        out = model(batch)
        loss = F.l1_loss(out, batch.y)
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def run_reference(args, rank, world):
    if rank != 0:
        return
    ds, batch = make_workload(0, GRAPHS_PER_GPU)
    E = batch.edge_index.shape[1]
    cores, table = best_cpu_threads(ds, batch)
    torch.set_num_threads(cores)
    times = cpu_reference_steps(ds, batch, args.steps, max(args.warmup, 1))
    ms = 1e3 * float(np.mean(times))
    value = GRAPHS_PER_GPU / (ms / 1e3)
    line = {
        "impl": "reference", "metric": "graphs_per_sec_cgcnn_train_step", "value": value,
        "unit": "graphs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "edges_per_sec": E / (ms / 1e3),
        "config": {"workload": "CGCNN dim=64 4xCGConv, synthetic bulk graphs, batch 256 (configs[1])",
                   "graphs_per_step": GRAPHS_PER_GPU, "edges_per_step": E,
                   "note": "reference CPU path = PyG-equivalent op sequence restated in oracle/ "
                           "(torch_geometric/torch_scatter not installable offline); rank 0 only"},
        "cpu_baseline": {"value": value, "unit": "graphs/s", "cores": cores, "kind": "port",
                         "host_cores": os.cpu_count(), "thread_sweep_graphs_per_s": table,
                         "sample": f"{args.steps} full train steps of the 256-graph batch, best thread count"},
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

    f_ms, f_min = time_it(fwd)
    b_ms, b_min = time_it(bwd)
    bytes_fwd = 8 * N * C + 8 * E + 4 * E * G          # SURVEY.md 8d, operator-surface form
    bytes_bwd = bytes_fwd + 4 * N * C                   # whole backward (both passes)
    res = {
        "workload": f"{base_graphs * reps} bulk graphs (N={N}, E={E}, C={C}, G={G}), cold L2",
        "fwd": {"ms": f_ms, "ms_min": f_min, "algorithmic_bytes": bytes_fwd,
                "achieved_gbs": bytes_fwd / f_ms / 1e6, "frac": bytes_fwd / f_ms / 1e6 / peak},
        "bwd_both_passes": {"ms": b_ms, "ms_min": b_min, "algorithmic_bytes": bytes_bwd,
                            "achieved_gbs": bytes_bwd / b_ms / 1e6, "frac": bytes_bwd / b_ms / 1e6 / peak},
        "edge_flops_fwd": 2.0 * E * G * 2 * C, "fwd_tflops_fp32": 2.0 * E * G * 2 * C / f_ms / 1e9,
    }
    return res


def best_cpu_threads(ds, batch):
    """Host thread count that maximises the CPU path's throughput (more threads than
    the batch can feed only adds synchronisation cost); bounded sweep."""
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, cores) if c <= cores})
    best, table = None, {}
    for c in cands:
        torch.set_num_threads(c)
        ts = cpu_reference_steps(ds, batch, 2, 1)
        table[c] = GRAPHS_PER_GPU / float(np.mean(ts))
        if best is None or table[c] > table[best]:
            best = c
    return best, table


def run_engine(args, rank, world, local_rank):
    from matdeeplearn_b200 import _lib, models as M, dist as mdist
    from matdeeplearn_b200.engine import TrainStep
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    _lib.load()
    peak, peak_src = peaks()
    if args.roofline_only:
        flush_buf = torch.zeros(128 * 1024 * 1024, device=dev)
        print(json.dumps(kernel_roofline(args, dev, flush_buf, peak, peak_src)), flush=True)
        return
    ds, host_batch = make_workload(rank, GRAPHS_PER_GPU)
    host_batch.num_graphs = GRAPHS_PER_GPU
    N, E = host_batch.x.shape[0], host_batch.edge_index.shape[1]
    torch.manual_seed(0)
    model = M.CGCNN(ds, **MODEL_CFG).to(dev)
    model.train()
    step = TrainStep(model, lr=LR * world)
    mdist.broadcast_(step.flat.param)
    flush_buf = torch.zeros(128 * 1024 * 1024, device=dev)  # 512 MiB

    # ---- device-resident, graph-replayed step (value)
    dev_batch = host_batch.to(dev)
    dev_batch.num_graphs = GRAPHS_PER_GPU
    replay = step.resident(dev_batch, warmup=3)
    for _ in range(max(args.warmup, 3)):
        replay()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms, wall = timed_steps(replay, args.steps, flush_buf, world)
    # ---- end to end from pinned host buffers through the public API (e2e)
    pinned = host_batch.pin_memory()
    pinned.num_graphs = GRAPHS_PER_GPU
    pinned.smear = getattr(host_batch, "smear", None)
    e2e_steps = max(5, min(args.steps, 30))
    # (i) the batch's 7 reference tensors copied as they are (edge_attr materialised on the host)
    for _ in range(3):
        step.from_host(pinned, expand_edge_attr=False)
    e2e_full_ms, _ = timed_steps(lambda: step.from_host(pinned, expand_edge_attr=False), e2e_steps, flush_buf, world)
    h2d_full = step.last_h2d_bytes
    # (ii) default public path: normalised distances shipped, Gaussian basis expanded on the GPU
    for _ in range(3):
        step.from_host(pinned)
    e2e_ms, _ = timed_steps(lambda: step.from_host(pinned), e2e_steps, flush_buf, world)
    h2d = step.last_h2d_bytes
    # (iii) dataset resident in HBM (store.GraphStore): per step a fresh permutation of the rank's graphs is
    #       assembled on the GPU into padded buffers and the captured step replayed; nothing crosses PCIe
    #       but 6 KB of ids.  Not the e2e number (no host batch), reported beside it.
    from matdeeplearn_b200.store import GraphStore
    store = GraphStore.from_dataset(ds, dev)
    perm_rng = np.random.default_rng(1234 + rank)
    perms = [perm_rng.permutation(GRAPHS_PER_GPU) for _ in range(e2e_steps + 3)]
    for i in range(3):
        step.from_store(store, perms[i])
    it = iter(perms[3:])
    store_ms, _ = timed_steps(lambda: step.from_store(store, next(it)), e2e_steps, flush_buf, world)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    graphs_total = GRAPHS_PER_GPU * world
    e_total = torch.tensor([float(E)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e_total)
    edges_total = float(e_total.item())
    value = graphs_total / (ms_per_step / 1e3)
    e2e_value = graphs_total / (e2e_ms / e2e_steps / 1e3)

    line = {
        "metric": "graphs_per_sec_cgcnn_train_step", "value": value, "unit": "graphs/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "edges_per_sec": edges_total / (ms_per_step / 1e3),
        "config": {
            "workload": "CGCNN dim=64 4xCGConv, synthetic bulk graphs, batch 256 per GPU (configs[1])",
            "graphs_per_step": graphs_total, "nodes_per_gpu": N, "edges_per_gpu": E,
            "parallelism": f"dp{world}", "step": "zero_grad+fwd+l1_loss+bwd" +
            ("+flat grad allreduce(NCCL, eager)+AdamW, two CUDA graph replays around the collective"
             if world > 1 else "+AdamW, one CUDA graph replay"),
            "l2": "flushed between timed steps (512 MiB read-modify-write)",
            "timing": "sum of per-step CUDA-event durations, max over ranks",
        },
        "gpu_launches": int(step.kernels_per_step) * args.steps,
        "gpu_launches_per_step": int(step.kernels_per_step),
        "e2e": {"value": e2e_value, "unit": "graphs/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": 4, "steps": e2e_steps,
                "path": "TrainStep.from_host(pinned Batch): H2D of x, edge_index, edge_weight, batch, u, y and the "
                        "normalised distances d_hat [E] (edge_attr = GaussianSmearing(d_hat) is expanded on the GPU), "
                        "then one CUDA graph replay of smear + CSR build + slot permute + fwd + bwd + AdamW, then "
                        "loss.item()",
                "materialised_edge_attr": {
                    "value": graphs_total / (e2e_full_ms / e2e_steps / 1e3), "h2d_bytes_per_step": int(h2d_full),
                    "path": "same call with expand_edge_attr=False: all 7 reference tensors copied, edge_attr "
                            "[E,50] included"}},
        "store_step": {"value": graphs_total / (store_ms / e2e_steps / 1e3), "unit": "graphs/s", "steps": e2e_steps,
                       "path": "TrainStep.from_store(GraphStore, idx): dataset resident in HBM, batch = index list -> "
                               "one assembly kernel into capacity-padded buffers + the step, one CUDA graph replay; "
                               "loss.item() each step"},
        "wall_s_timed_region": wall,
    }
    if rank == 0:
        line["clocks"] = clocks
        if not args.no_roofline:
            r = kernel_roofline(args, dev, flush_buf, peak, peak_src)
            is_bwd = r["bwd_both_passes"]["ms"] > r["fwd"]["ms"]
            dom = r["bwd_both_passes"] if is_bwd else r["fwd"]
            traffic = None  # dram__bytes_read+write per launch from the committed ncu --set full capture
            tpath = os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")
            if os.path.exists(tpath):
                t = json.load(open(tpath))
                if t.get("graphs") == args.roofline_graphs:
                    traffic = t.get("bwd_bytes" if is_bwd else "fwd_bytes")
            line["roofline"] = {"bound": "hbm", "achieved": dom["achieved_gbs"], "peak": peak,
                                "unit": "GB/s", "frac": dom["frac"], "traffic": traffic,
                                "kernel": "k_cgconv_tc<BWD_DST> (whole CGConv backward: single pass, dP + dWe + dQ)"
                                if is_bwd else "k_cgconv_fwd_pipe",
                                "note": "dominant = the slower of the two fused edge kernels; roofline_detail.fwd is the "
                                        "fused gather->message->scatter_add forward kernel (k_cgconv_fwd_pipe)",
                                "peak_source": peak_src, "workload": r["workload"]}
            line["roofline_detail"] = r
        if not args.no_cpu_baseline and world == 1:
            cds, cb = make_workload(0, GRAPHS_PER_GPU)
            threads, table = best_cpu_threads(cds, cb)
            torch.set_num_threads(threads)
            ts = cpu_reference_steps(cds, cb, args.cpu_steps, 1)
            v = GRAPHS_PER_GPU / float(np.mean(ts))
            line["cpu_baseline"] = {"value": v, "unit": "graphs/s", "cores": threads, "kind": "port",
                                    "host_cores": os.cpu_count(), "thread_sweep_graphs_per_s": table,
                                    "sample": f"{args.cpu_steps} full train steps of the same 256-graph batch "
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
    try:
        run_engine(args, rank, world, local_rank)
    finally:
        if world > 1:
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
