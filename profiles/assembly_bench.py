"""Batch assembly: GraphStore.batch (one gather kernel on the GPU) against the reference-style path
(host collate = Batch.from_data_list, pin, H2D copy, layout sort + slot permutation on the GPU).
Prints one JSON line per variant.  Run on the GPU box:  python profiles/assembly_bench.py"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matdeeplearn_b200 import process as pr          # noqa: E402
from matdeeplearn_b200.csr import csr_for             # noqa: E402
from matdeeplearn_b200.store import GraphStore        # noqa: E402

DEV = torch.device("cuda", 0)


def main():
    n_graphs, B, reps = 4096, 256, 30
    ds = pr.synthetic_dataset("bulk", n_graphs, seed=pr.BENCH_SEED)
    peaks = json.load(open("MEASURED_PEAKS.json")) if os.path.exists("MEASURED_PEAKS.json") else {}
    rng = np.random.default_rng(0)
    batches = [rng.permutation(n_graphs)[:B] for _ in range(reps)]
    flush = torch.zeros(128 * 1024 * 1024, device=DEV)
    for keep in (False, True):
        t0 = time.perf_counter()
        store = GraphStore.from_dataset(ds, DEV, keep_edge_attr=keep)
        torch.cuda.synchronize()
        build_s = time.perf_counter() - t0
        for idx in batches[:3]:
            store.batch(idx)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in batches]
        wall = []
        for idx, (a, b) in zip(batches, ev):
            flush.add_(1.0)
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            a.record()
            out = store.batch(idx)
            b.record()
            torch.cuda.synchronize()
            wall.append(time.perf_counter() - w0)
        dev_ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
        N, E = out.x.shape[0], out.edge_index.shape[1]
        G, F = store.G, store.F
        # bytes the assembly must move: every output once, every stored input once
        out_b = 4 * N * F + 16 * E + 4 * E + 2 * 4 * E * G + 8 * N + 4 * (2 * (N + 1) + 4 * E + 2 * N)
        in_b = 4 * N * F + 8 * E + 4 * E + (4 * E * G + 4 * E if keep else 8 * E) + 4 * (2 * N + 4 * E + 2 * N)
        print(json.dumps({
            "what": "GraphStore.batch", "edge_attr": "stored" if keep else "expanded from d_hat",
            "graphs": B, "nodes": N, "edges": E, "store_graphs": n_graphs, "store_bytes": store.nbytes(),
            "store_build_s": build_s, "device_ms_median": dev_ms, "wall_ms_median": float(np.median(wall)) * 1e3,
            "algorithmic_bytes": in_b + out_b, "achieved_gbs": (in_b + out_b) / dev_ms / 1e6,
            "hbm_peak_gbs": peaks.get("hbm_copy_gbs") or peaks.get("hbm_gbs"),
        }), flush=True)
    # reference-style path: host collate, pin, H2D, layout build + slot permutation on the GPU
    col, h2d, lay = [], [], []
    for idx in batches[:10]:
        t0 = time.perf_counter()
        hb = ds.batch([int(i) for i in idx])
        t1 = time.perf_counter()
        pb = hb.pin_memory()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        db = pb.to(DEV, non_blocking=True)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        csr = csr_for(db.edge_index, db.batch, num_graphs=B)
        csr.to_slots(db.edge_attr)
        torch.cuda.synchronize()
        t4 = time.perf_counter()
        col.append(t1 - t0), h2d.append(t3 - t2), lay.append(t4 - t3)
    print(json.dumps({"what": "host collate path", "graphs": B,
                      "collate_ms_median": float(np.median(col)) * 1e3,
                      "h2d_ms_median": float(np.median(h2d)) * 1e3,
                      "layout_ms_median": float(np.median(lay)) * 1e3}), flush=True)
    # ---- one shuffled epoch (variable batch shapes => eager launches, no graph replay)
    from matdeeplearn_b200 import models as M
    from matdeeplearn_b200.engine import TrainStep
    cfg = dict(dim1=64, dim2=64, pre_fc_count=1, gc_count=4, post_fc_count=1, pool="global_mean_pool",
               pool_order="early", batch_norm="True", batch_track_stats="True", act="relu", dropout_rate=0.0)
    store = GraphStore.from_dataset(ds, DEV)
    order = np.random.default_rng(1).permutation(n_graphs)
    chunks = [order[i:i + B] for i in range(0, n_graphs, B)]
    for name in ("device store, padded graph replay", "device store", "host collate"):
        torch.manual_seed(0)
        model = M.CGCNN(ds, **cfg).to(DEV).train()
        step = TrainStep(model, lr=1e-3)
        def one(idx):
            if name == "device store, padded graph replay":
                return step.from_store(store, idx, sync=False)
            if name == "device store":
                return step.eager(store.batch(idx))
            hb = ds.batch([int(i) for i in idx]).pin_memory()
            return step.eager(hb.to(DEV, non_blocking=True))
        for idx in chunks[:2]:
            one(idx)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for idx in chunks:
            loss = one(idx)
        last = float(loss.item())
        dt = time.perf_counter() - t0
        print(json.dumps({"what": "shuffled epoch", "batches_from": name, "graphs": n_graphs,
                          "batch": B, "wall_s": dt, "graphs_per_s": n_graphs / dt, "last_loss": last}), flush=True)


if __name__ == "__main__":
    main()
