"""CPU, world_size 2, gloo: the data-parallel plumbing (whole-graph sharding, one flat
gradient buffer, a single mean all-reduce) reproduces the single-process gradient of the
per-rank-mean loss -- DDP semantics of reference training/training.py:262-266."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _make_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(6, 16), torch.nn.ReLU(), torch.nn.BatchNorm1d(16),
                               torch.nn.Linear(16, 1))


def _data():
    g = torch.Generator().manual_seed(1)
    return torch.randn(40, 6, generator=g), torch.randn(40, generator=g)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from matdeeplearn_b200 import dist as mdist
    X, y = _data()
    idx = mdist.shard_indices(40, rank, world)
    model = _make_model()
    flat = mdist.FlatParameters(model)
    mdist.broadcast_(flat.param)
    flat.zero_grad()
    loss = torch.nn.functional.l1_loss(model(X[idx]).view(-1), y[idx])
    loss.backward()
    mdist.allreduce_mean_(flat.grad)
    if rank == 0:
        torch.save({"grad": flat.grad.clone(), "idx": idx}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_flat_allreduce_matches_single_process(tmp_path):
    world = 2
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    got = torch.load(out)["grad"]
    # single-process reference: mean over ranks of each rank's gradient (BN stats per rank)
    from matdeeplearn_b200 import dist as mdist
    X, y = _data()
    ref = None
    for r in range(world):
        idx = mdist.shard_indices(40, r, world)
        model = _make_model()
        loss = torch.nn.functional.l1_loss(model(X[idx]).view(-1), y[idx])
        loss.backward()
        g = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        ref = g if ref is None else ref + g
    ref /= world
    torch.testing.assert_close(got, ref, rtol=1e-6, atol=1e-7)


def test_flat_parameters_views_track_optimizer_updates():
    from matdeeplearn_b200 import dist as mdist
    model = _make_model()
    flat = mdist.FlatParameters(model)
    X, y = _data()
    opt = torch.optim.SGD([flat.leaf], lr=0.1)
    before = [p.detach().clone() for p in model.parameters()]
    flat.zero_grad()
    torch.nn.functional.l1_loss(model(X).view(-1), y).backward()
    assert flat.grad.abs().sum() > 0           # autograd accumulated INTO the flat buffer
    opt.step()
    after = list(model.parameters())
    assert any(not torch.equal(a, b) for a, b in zip(before, after))  # views saw the update
    off = 0
    for p in model.parameters():
        assert p.data_ptr() == flat.param.data_ptr() + 4 * off
        off += p.numel()


def test_shard_indices_balanced_by_size():
    from matdeeplearn_b200 import dist as mdist
    sizes = [100, 90, 80, 10, 10, 10, 5, 5]
    shards = [mdist.shard_indices(8, r, 2, sizes) for r in range(2)]
    assert sorted(shards[0] + shards[1]) == list(range(8))
    loads = [sum(sizes[i] for i in s) for s in shards]
    assert abs(loads[0] - loads[1]) <= 30  # longest-first greedy
    assert mdist.shard_indices(7, 1, 3) == [1, 4]
