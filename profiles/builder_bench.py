"""Graph builder: GPU (GraphStore.from_structures, csrc/builder.cu) against the host builder
(process.assemble_dataset + GraphStore.from_dataset).  python profiles/builder_bench.py [n_structures]"""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matdeeplearn_b200 import process as pr
from matdeeplearn_b200.store import GraphStore
dev = torch.device("cuda", 0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
for kind, count in (("bulk", n), ("mof", max(n // 16, 8))):
    structs, ys = pr.synthetic_structures(kind, count, seed=5)
    atoms = sum(len(s[0]) for s in structs)
    GraphStore.from_structures(structs[:8], ys[:8], dev)          # warm-up (module load, attributes)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); gpu = GraphStore.from_structures(structs, ys, dev); torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    # kernel time alone
    lib_t = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); GraphStore.from_structures(structs, ys, dev); b.record(); torch.cuda.synchronize()
        lib_t.append(a.elapsed_time(b))
    t0 = time.perf_counter(); ds = pr.assemble_dataset(structs, ys); t_host_build = time.perf_counter() - t0
    t0 = time.perf_counter(); host = GraphStore.from_dataset(ds, dev); torch.cuda.synchronize()
    t_host_upload = time.perf_counter() - t0
    same = torch.equal(host.src, gpu.src) and torch.equal(host.edge_weight, gpu.edge_weight) and torch.equal(host.x, gpu.x)
    print(json.dumps({"kind": kind, "structures": count, "atoms": atoms, "edges": gpu.num_edges,
                      "gpu_from_structures_s": t_gpu, "gpu_from_structures_events_ms_min": min(lib_t),
                      "host_assemble_dataset_s": t_host_build, "host_store_upload_s": t_host_upload,
                      "structures_per_s_gpu": count / t_gpu, "structures_per_s_host": count / t_host_build,
                      "identical": bool(same)}), flush=True)
