"""2 GPUs, NCCL: engine.TrainStep across ranks (reference training.py:262-266, 291-294: DDP over whole-graph shards).

Each rank steps on its own shard; the summed gradient must equal the sum of the per-shard oracle gradients
(per-rank BatchNorm statistics), replicas must stay bit-identical, and the step with the all-reduce
captured inside the CUDA graph must equal the one with the eager collective between two graphs.
Skipped on boxes with fewer than two GPUs (run it with `gpurun --gpus 2`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(dim1=64, dim2=64, pre_fc_count=1, gc_count=2, post_fc_count=1)
GRAPHS = 32


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, graph_allreduce):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), MDL_GRAPH_ALLREDUCE=graph_allreduce)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from matdeeplearn_b200 import dist as mdist, models as M, process as pr
    from matdeeplearn_b200.engine import TrainStep
    ds = pr.synthetic_dataset("bulk", GRAPHS, seed=5)
    idx = mdist.shard_indices(GRAPHS, rank, world)
    batch = ds.batch(idx).to(dev)
    batch.num_graphs = len(idx)
    torch.manual_seed(100 + rank)                      # replicas built from different seeds: TrainStep broadcasts rank 0's
    model = M.CGCNN(ds, **CFG).to(dev).train()
    step = TrainStep(model, lr=1e-3)
    assert step.graph_allreduce == (graph_allreduce == "1")
    init = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    replay = step.resident(batch)
    losses = [float(replay().item())]
    grad_sum = step.flat.grad.detach().cpu().clone()   # after the all-reduce of the first step (sum over ranks)
    for _ in range(3):
        losses.append(float(replay().item()))
    params = step.flat.param.detach().cpu().clone()
    gathered = [torch.empty_like(params) for _ in range(world)]
    dist.all_gather_object(gathered, params)
    if rank == 0:
        torch.save({"init": init, "grad_sum": grad_sum, "params": gathered, "losses": losses,
                    "names": [n for n, _ in model.named_parameters()],
                    "numels": [p.numel() for p in model.parameters()]}, out)
    dist.barrier()
    torch.cuda.synchronize()
    # no destroy_process_group(): it was observed to hang while CUDA graphs with captured NCCL kernels are alive
    os._exit(0)


def _run(tmp_path, graph_allreduce):
    import torch.multiprocessing as mp
    out = str(tmp_path / f"multi_{graph_allreduce}.pt")
    ctx = mp.spawn(_worker, args=(2, _free_port(), out, graph_allreduce), nprocs=2, join=False)
    import time
    deadline = time.time() + 240
    while not ctx.join(timeout=5):          # bounded: a stuck collective must fail the test, not hang the box
        if time.time() > deadline:
            for p in ctx.processes:
                p.kill()
            raise AssertionError("2-rank worker processes did not finish within 240 s")
    return torch.load(out, weights_only=False)


@pytest.mark.timeout(600)
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_trainstep_two_ranks_matches_oracle_and_eager_collective(tmp_path):
    from matdeeplearn_b200 import dist as mdist, process as pr
    from oracle import models as OM
    fused = _run(tmp_path, "1")
    # replicas identical after four steps
    assert torch.equal(fused["params"][0], fused["params"][1])
    # summed gradient of step 1 == sum over shards of the oracle's gradient (fp64, rank 0's initial weights)
    ds = pr.synthetic_dataset("bulk", GRAPHS, seed=5)
    ref = {}
    for r in range(2):
        m = OM.CGCNN(ds, **CFG).double()
        m.load_state_dict({k: (v.double() if v.is_floating_point() else v) for k, v in fused["init"].items()})
        m.train()
        b = ds.batch(mdist.shard_indices(GRAPHS, r, 2)).double()
        torch.nn.functional.l1_loss(m(b), b.y).backward()
        for name, p in m.named_parameters():
            ref[name] = p.grad.reshape(-1) if name not in ref else ref[name] + p.grad.reshape(-1)
    # the flat gradient buffer holds the engine model's parameters in ITS registration order: compare by name
    got, off = fused["grad_sum"].double(), 0
    scale = max(float(v.abs().max()) for v in ref.values())
    for name, n in zip(fused["names"], fused["numels"]):
        err = (got[off:off + n] - ref[name]).abs().max().item()
        assert err <= 2e-4 * scale, (name, err, scale)
        off += n
    assert off == got.numel()
    # eager collective between two graphs: same trajectory
    eager = _run(tmp_path, "0")
    assert torch.equal(eager["params"][0], eager["params"][1])
    assert max(abs(a - b) for a, b in zip(fused["losses"], eager["losses"])) <= 1e-6
    # two separate runs: the CGConv backward's dQ atomics make them differ in the last bits (1.9e-6 observed)
    assert (fused["params"][0] - eager["params"][0]).abs().max().item() <= 2e-5
