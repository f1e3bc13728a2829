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
    eager_losses = [float(s1.eager(b)) for _ in range(4)]
    replay = s2.resident(b, warmup=3)     # the warm-up steps are rolled back: capture is not training
    graph_losses = [float(replay()) for _ in range(4)]
    assert s2.kernels_per_step and s2.kernels_per_step > 0
    for a, c in zip(eager_losses, graph_losses):
        assert abs(a - c) <= 1e-6 * max(1.0, abs(a)), (eager_losses, graph_losses)


def test_from_host_graph_equals_eager_and_tracks_oracle():
    from matdeeplearn_b200.engine import TrainStep
    from oracle import models as OM
    ds, batch, model = _setup()
    m1, m2 = copy.deepcopy(model).to(DEV).train(), copy.deepcopy(model).to(DEV).train()
    s1, s2 = TrainStep(m1, lr=1e-3), TrainStep(m2, lr=1e-3)
    pinned = batch.pin_memory()
    pinned.num_graphs = 12
    # graph capture warms up with 2 eager steps on the first call and rolls them back
    l_eager = [s1.from_host(pinned, use_graph=False) for _ in range(6)]
    l_graph = [s2.from_host(pinned, use_graph=True) for _ in range(4)]
    for a, c in zip(l_eager, l_graph):
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
    # edge_attr stays in its 4 B/edge form on the device (the fused CGConv kernels expand it); its expansion is the
    # reference's tensor
    from matdeeplearn_b200.data import GaussianEdgeAttr
    assert isinstance(static.edge_attr, GaussianEdgeAttr)
    assert (static.edge_attr.materialize().cpu() - batch.edge_attr).abs().max().item() < 2e-6


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


@pytest.mark.parametrize("N,I,O", [(7862, 64, 256), (300, 114, 64), (256, 64, 1), (1, 8, 8), (5000, 100, 150)])
def test_linear_wgrad_kernel_matches_autograd(N, I, O):
    from matdeeplearn_b200 import functional as MF
    torch.manual_seed(N + I)
    x, g = torch.randn(N, I, device=DEV), torch.randn(N, O, device=DEV)
    dW = torch.full((O, I), float("nan"), device=DEV)
    db = torch.full((O,), float("nan"), device=DEV)
    MF.linear_wgrad_into(x, g, MF._wgrad_map(O, I, [dW.data_ptr()], [db.data_ptr()]))
    ref_w = g.double().t().mm(x.double())
    ref_b = g.double().sum(0)
    assert (dW.double() - ref_w).abs().max().item() <= 2e-6 * (g.abs().double().t().mm(x.abs().double())).max().item()
    assert (db.double() - ref_b).abs().max().item() <= 2e-6 * g.abs().double().sum(0).max().item()


def test_linear_wgrad_block_map_scatters_into_column_blocks():
    """CGConv's use: four row blocks of dPQ^T x go to column blocks of two [C, 2C+G] matrices."""
    from matdeeplearn_b200 import functional as MF
    torch.manual_seed(4)
    N, C, G = 1000, 64, 50
    x, dPQ = torch.randn(N, C, device=DEV), torch.randn(N, 4 * C, device=DEV)
    Wf = torch.zeros(C, 2 * C + G, device=DEV)
    Ws = torch.zeros(C, 2 * C + G, device=DEV)
    bf, bs = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    ld = 2 * C + G
    MF.linear_wgrad_into(x, dPQ, MF._wgrad_map(
        C, ld, [Wf.data_ptr(), Ws.data_ptr(), Wf.data_ptr() + 4 * C, Ws.data_ptr() + 4 * C],
        [bf.data_ptr(), bs.data_ptr(), None, None]))
    ref = dPQ.t().mm(x)
    for got, blk in ((Wf[:, :C], 0), (Ws[:, :C], 1), (Wf[:, C:2 * C], 2), (Ws[:, C:2 * C], 3)):
        assert (got - ref[blk * C:(blk + 1) * C]).abs().max().item() < 1e-3
    assert Wf[:, 2 * C:].abs().max().item() == 0                    # untouched columns
    assert (bf - dPQ[:, :C].sum(0)).abs().max().item() < 1e-3
    assert (bs - dPQ[:, C:2 * C].sum(0)).abs().max().item() < 1e-3


def test_direct_gradient_delivery_equals_autograd_gradients():
    """TrainStep writes every weight/bias gradient straight into the flat buffer (no .grad tensors,
    no concatenation); the buffer must hold what plain autograd would have produced."""
    from matdeeplearn_b200.engine import TrainStep
    ds, batch, model = _setup()
    m1, m2 = copy.deepcopy(model).to(DEV).train(), copy.deepcopy(model).to(DEV).train()
    b = batch.to(DEV)
    step = TrainStep(m1, lr=0.0)
    step._fwd_bwd(b)
    assert all(p.grad is None for p in m1.parameters())          # nothing went through .grad
    loss = torch.nn.functional.l1_loss(m2(b), b.y)
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in m2.parameters()])
    got = step.flat.grad
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() <= 2e-5 * scale


@pytest.mark.parametrize("dim1", [100, 128])
def test_train_step_wide_cgcnn_direct_gradients(dim1):
    """TrainStep (direct gradient delivery) on the reference's default width (config.yml CGCNN_demo: dim1=100)
    and on 128: the weight-gradient kernel must take [N,C] x [N,4C] there, and the flat buffer must hold
    what plain autograd produces."""
    from matdeeplearn_b200 import models as M, process as pr
    from matdeeplearn_b200.engine import TrainStep
    ds = pr.synthetic_dataset("bulk", 12, seed=11)
    batch = ds.batch()
    torch.manual_seed(1)
    model = M.CGCNN(ds, dim1=dim1, dim2=64, pre_fc_count=1, gc_count=2, post_fc_count=1)
    m1, m2 = copy.deepcopy(model).to(DEV).train(), copy.deepcopy(model).to(DEV).train()
    b = batch.to(DEV)
    step = TrainStep(m1, lr=1e-3)
    step._fwd_bwd(b)
    loss = torch.nn.functional.l1_loss(m2(b), b.y)
    loss.backward()
    ref = torch.cat([p.grad.reshape(-1) for p in m2.parameters()])
    assert (step.flat.grad - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    losses = [float(step.eager(b)) for _ in range(3)]
    assert all(l == l for l in losses)


def test_from_host_many_shapes_is_bounded_and_does_not_overtrain():
    """Every new (N, E) shape captures a graph: the warm-up must not advance the optimizer, and the cache of
    captured shapes is bounded (LRU)."""
    from matdeeplearn_b200 import process as pr
    from matdeeplearn_b200.engine import TrainStep
    ds, batch, model = _setup()
    m1 = copy.deepcopy(model).to(DEV).train()
    s1 = TrainStep(m1, lr=1e-3)
    s1.max_host_graphs = 2
    for k, n in enumerate((5, 7, 9, 11)):
        sub = pr.synthetic_dataset("bulk", n, seed=20 + k).batch().pin_memory()
        sub.num_graphs = n
        s1.from_host(sub)
        assert int(s1.opt.step_count.item()) == k + 1
    assert len(s1._host_graphs) == 2
