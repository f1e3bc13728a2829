"""GPU: the CUDA-graph replayed training step (device-resident and from-host variants)
reproduces the eager step, and follows the CPU oracle's training trajectory."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG = dict(dim1=64, dim2=64, pre_fc_count=1, gc_count=2, post_fc_count=1)


def _setup(seed=0):
    from matdeeplearn_b200 import models as M, process as pr
    ds = pr.synthetic_dataset("bulk", 12, seed=11)
    batch = ds.batch()
    batch.num_graphs = 12
    torch.manual_seed(seed)
    model = M.CGCNN(ds, **CFG)
    return ds, batch, model


def test_resident_graph_replay_equals_eager():
    from matdeeplearn_b200.engine import TrainStep
    ds, batch, model = _setup()
    m1, m2 = copy.deepcopy(model).to(DEV).train(), copy.deepcopy(model).to(DEV).train()
    s1, s2 = TrainStep(m1, lr=1e-3), TrainStep(m2, lr=1e-3)
    b = batch.to(DEV)
    b.num_graphs = 12
    eager_losses = [float(s1.eager(b)) for _ in range(3 + 4)]   # resident() warms up 3 eager steps
    replay = s2.resident(b, warmup=3)
    graph_losses = [float(replay()) for _ in range(4)]
    assert s2.kernels_per_step and s2.kernels_per_step > 0
    for a, c in zip(eager_losses[3:], graph_losses):
        assert abs(a - c) <= 1e-6 * max(1.0, abs(a)), (eager_losses, graph_losses)


def test_from_host_graph_equals_eager_and_tracks_oracle():
    from matdeeplearn_b200.engine import TrainStep
    from oracle import models as OM
    ds, batch, model = _setup()
    m1, m2 = copy.deepcopy(model).to(DEV).train(), copy.deepcopy(model).to(DEV).train()
    s1, s2 = TrainStep(m1, lr=1e-3), TrainStep(m2, lr=1e-3)
    pinned = batch.pin_memory()
    pinned.num_graphs = 12
    # graph capture warms up with 2 eager steps on the first call
    l_eager = [s1.from_host(pinned, use_graph=False) for _ in range(6)]
    l_graph = [s2.from_host(pinned, use_graph=True) for _ in range(4)]
    for a, c in zip(l_eager[2:], l_graph):
        assert abs(a - c) <= 1e-6 * max(1.0, abs(a)), (l_eager, l_graph)
    # CPU oracle trajectory (same init, same optimizer)
    ref = OM.CGCNN(ds, **CFG)
    ref.load_state_dict(model.state_dict())
    ref.train()
    opt = torch.optim.AdamW(ref.parameters(), lr=1e-3, weight_decay=1e-2)
    l_ref = []
    for _ in range(6):
        opt.zero_grad()
        loss = torch.nn.functional.l1_loss(ref(batch), batch.y)
        loss.backward()
        opt.step()
        l_ref.append(float(loss))
    for a, r in zip(l_eager, l_ref):
        assert abs(a - r) <= 5e-4 * max(1.0, abs(r)), (l_eager, l_ref)


def test_from_host_device_side_gaussian_expansion_matches_materialised_edge_attr():
    """from_host ships d_hat and expands GaussianSmearing (reference process.py:580-590) on the GPU when
    the batch carries it; the step must equal the one fed the host-materialised edge_attr."""
    from matdeeplearn_b200.engine import TrainStep
    ds, batch, model = _setup()
    assert hasattr(batch, "d_hat") and batch.smear["resolution"] == batch.edge_attr.shape[1]
    m1, m2 = copy.deepcopy(model).to(DEV).train(), copy.deepcopy(model).to(DEV).train()
    s1, s2 = TrainStep(m1, lr=1e-3), TrainStep(m2, lr=1e-3)
    pinned = batch.pin_memory()
    la = [s1.from_host(pinned, expand_edge_attr=True) for _ in range(4)]
    lb = [s2.from_host(pinned, expand_edge_attr=False) for _ in range(4)]
    assert s2.last_h2d_bytes - s1.last_h2d_bytes == 4 * (batch.edge_attr.numel() - batch.d_hat.numel())
    for a, c in zip(la, lb):
        assert abs(a - c) <= 2e-6 * max(1.0, abs(a)), (la, lb)
    key = [k for k in s1._host_graphs if k[-1]][0]
    static = s1._host_graphs[key][0]
    assert (static.edge_attr.cpu() - batch.edge_attr).abs().max().item() < 2e-6


def test_flat_adamw_matches_torch_adamw():
    from matdeeplearn_b200 import dist as mdist
    torch.manual_seed(3)
    lin = torch.nn.Sequential(torch.nn.Linear(20, 30), torch.nn.Tanh(), torch.nn.Linear(30, 5))
    ref = copy.deepcopy(lin).double()
    lin = lin.to(DEV)
    flat = mdist.FlatParameters(lin)
    opt = mdist.FlatAdamW(flat, lr=3e-3, weight_decay=0.02)
    ropt = torch.optim.AdamW(ref.parameters(), lr=3e-3, weight_decay=0.02)
    x = torch.randn(64, 20)
    for it in range(5):
        flat.release_grads()
        lin(x.to(DEV)).pow(2).mean().backward()
        flat.pack_grads()
        opt.step()
        ropt.zero_grad()
        ref(x.double()).pow(2).mean().backward()
        ropt.step()
    for p, q in zip(lin.parameters(), ref.parameters()):
        assert (p.detach().cpu().double() - q.detach()).abs().max().item() < 2e-6
