"""GPU: a batch assembled from the device-resident GraphStore (mdl_assemble_batch) is, bit for bit,
the batch the reference's collate would produce (Batch.from_data_list, restating PyG's collate used at
training.py:300-307) copied to the device, and the layout it carries equals mdl_csr_from_coo's."""
import copy

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LAYOUT = ("dst_ptr", "dst_src", "dst_dst", "dst_eid", "src_ptr", "src_slot", "inv_deg_dst", "inv_deg_src",
          "graph_ptr")


def _dataset(kind="bulk", n=40, seed=5):
    from matdeeplearn_b200 import process as pr
    return pr.synthetic_dataset(kind, n, seed=seed)


@pytest.mark.parametrize("keep", [True, False])
@pytest.mark.parametrize("kind,n,B", [("bulk", 40, 17), ("bulk", 40, 1), ("mof", 6, 4)])
def test_assembled_batch_equals_host_collate(kind, n, B, keep):
    from matdeeplearn_b200.csr import GraphCSR, gather_rows
    from matdeeplearn_b200.data import Batch
    from matdeeplearn_b200.store import GraphStore
    from matdeeplearn_b200 import functional as MF
    ds = _dataset(kind, n)
    store = GraphStore.from_dataset(ds, DEV, keep_edge_attr=keep)
    rng = np.random.default_rng(B)
    for trial in range(3):
        idx = rng.integers(0, n, size=B)          # repeats allowed, arbitrary order
        got = store.batch(idx, d_hat=True)
        ref = Batch.from_data_list([ds[int(i)] for i in idx])
        for k in ("x", "edge_index", "edge_weight", "batch", "u", "y", "d_hat"):
            a, b = getattr(got, k).cpu(), getattr(ref, k)
            assert a.dtype == b.dtype and a.shape == b.shape, k
            assert torch.equal(a, b), k
        if keep:
            assert torch.equal(got.edge_attr.cpu(), ref.edge_attr)
        else:
            # on-the-fly expansion: identical to the smearing kernel, within an ulp or two of the host's exp
            s = store.smear
            dev_ea = MF.gaussian_smear(got.d_hat, store.smear_offset, store.smear_coeff)
            assert torch.equal(got.edge_attr, dev_ea)
            assert (got.edge_attr.cpu() - ref.edge_attr).abs().max().item() < 2e-6
            assert s["resolution"] == ref.edge_attr.shape[1]
        # the carried layout == a fresh sort of the assembled COO
        csr = got.edge_index._mdl_csr[1]
        fresh = GraphCSR.from_coo(got.edge_index.clone(), got.batch.clone(), num_nodes=got.x.shape[0], num_graphs=B)
        for name in LAYOUT:
            assert torch.equal(getattr(csr, name), getattr(fresh, name)), name
        hit = got.edge_attr._mdl_slots
        assert hit[0] is csr
        assert torch.equal(hit[2], gather_rows(got.edge_attr, fresh.dst_eid))
        assert csr.to_slots(got.edge_attr) is hit[2]


def test_training_step_on_assembled_batch_equals_host_path():
    from matdeeplearn_b200 import models as M
    from matdeeplearn_b200.store import GraphStore
    ds = _dataset("bulk", 24)
    store = GraphStore.from_dataset(ds, DEV, keep_edge_attr=True)
    idx = [3, 7, 1, 20, 11, 5, 9, 2]
    torch.manual_seed(0)
    model = M.CGCNN(ds, dim1=64, dim2=64, pre_fc_count=1, gc_count=2, post_fc_count=1)
    m1, m2 = copy.deepcopy(model).to(DEV).train(), copy.deepcopy(model).to(DEV).train()
    b1 = store.batch(idx)
    b2 = ds.batch(idx).to(DEV)
    l1 = torch.nn.functional.l1_loss(m1(b1), b1.y)
    l2 = torch.nn.functional.l1_loss(m2(b2), b2.y)
    l1.backward()
    l2.backward()
    assert abs(l1.item() - l2.item()) <= 1e-6 * max(1.0, abs(l2.item()))
    for (n1, p1), (n2, p2) in zip(m1.named_parameters(), m2.named_parameters()):
        assert n1 == n2
        scale = max(p2.grad.abs().max().item(), 1e-6)
        assert (p1.grad - p2.grad).abs().max().item() <= 2e-5 * scale, n1


def test_store_argument_checks():
    from matdeeplearn_b200.store import GraphStore
    ds = _dataset("bulk", 5)
    store = GraphStore.from_dataset(ds, DEV)
    with pytest.raises(IndexError):
        store.batch([0, 5])
    with pytest.raises(ValueError):
        store.batch([])
    with pytest.raises(RuntimeError):
        GraphStore.from_dataset(ds, "cpu")
    b = store.batch([4, 4, 0], layout=False)
    assert not hasattr(b.edge_index, "_mdl_csr") and b.x.shape[0] == 2 * ds[4].x.shape[0] + ds[0].x.shape[0]


# ---------------------------------------------------------------- capacity-padded batches
def test_padded_static_batch_real_rows_match_and_padding_is_inert():
    from matdeeplearn_b200.csr import GraphCSR
    from matdeeplearn_b200.store import GraphStore
    ds = _dataset("bulk", 40)
    store = GraphStore.from_dataset(ds, DEV)
    B = 9
    Nc, Ec = store.capacities(B)
    static = store.static_batch(B, d_hat=True)
    rng = np.random.default_rng(3)
    for trial in range(4):                      # the same buffers refilled with batches of different size
        idx = rng.integers(0, 40, size=B)
        assert store.load(static, idx)
        store.assemble(static)
        exact = store.batch(idx, d_hat=True)
        N, E = exact.x.shape[0], exact.edge_index.shape[1]
        assert static._valid == (N, E) and N < Nc and E < Ec
        assert int(static._n_valid.item()) == N
        for k in ("x", "edge_attr", "edge_weight", "batch", "d_hat"):
            assert torch.equal(getattr(static, k)[: getattr(exact, k).shape[0]], getattr(exact, k)), k
        assert torch.equal(static.edge_index[:, :E], exact.edge_index)
        assert torch.equal(static.u, exact.u) and torch.equal(static.y, exact.y)
        csr, ref = static._parts[0], exact.edge_index._mdl_csr[1]
        assert (csr.N, csr.E, csr.B) == (Nc, Ec, B)
        for name in ("dst_src", "dst_dst", "dst_eid", "src_slot"):
            assert torch.equal(getattr(csr, name)[:E], getattr(ref, name)), name
        for name in ("dst_ptr", "src_ptr", "inv_deg_dst", "inv_deg_src"):
            assert torch.equal(getattr(csr, name)[:N], getattr(ref, name)[:N]), name
        assert torch.equal(csr.graph_ptr, ref.graph_ptr)
        assert torch.equal(static._parts[1][:E], exact.edge_attr._mdl_slots[2])
        # padding: zero rows, phantom graph id, empty segments, nothing pointing into real data
        assert static.x[N:].abs().max().item() == 0 and static.edge_attr[E:].abs().max().item() == 0
        assert static._parts[1][E:].abs().max().item() == 0
        assert (static.batch[N:] == B).all()
        assert (csr.dst_ptr[N:] == E).all() and (csr.src_ptr[N:] == E).all()
        assert (csr.dst_dst[E:] == Nc - 1).all() and (csr.dst_src[E:] == Nc - 1).all()
        assert torch.equal(csr.dst_eid[E:].long(), torch.arange(E, Ec, device=DEV))
        assert (csr.inv_deg_dst[N:] == 0).all()


@pytest.mark.parametrize("C,N,n", [(64, 300, 211), (100, 77, 77), (7, 1000, 1), (64, 5000, 4999)])
def test_masked_batchnorm_equals_torch_on_the_valid_rows(C, N, n):
    from matdeeplearn_b200 import functional as MF
    torch.manual_seed(C + N)
    x = (torch.randn(N, C, device=DEV) * 3 + 5).requires_grad_()
    g = torch.randn(N, C, device=DEV)
    bn_a = torch.nn.BatchNorm1d(C).to(DEV).train()
    bn_b = copy.deepcopy(bn_a)
    with torch.no_grad():
        bn_a.weight.uniform_(0.5, 1.5), bn_a.bias.uniform_(-1, 1)
        bn_b.load_state_dict(bn_a.state_dict())
    nv = torch.tensor([n], dtype=torch.int32, device=DEV)
    if n > 1:
        xr = x.detach()[:n].clone().requires_grad_()
        ref = bn_b(xr)
        ref.backward(g[:n])
    out = MF.masked_batch_norm(bn_a, x, nv)
    out.backward(g)
    assert out[n:].abs().max().item() == 0 if n < N else True
    assert x.grad[n:].abs().max().item() == 0 if n < N else True
    if n > 1:
        assert (out[:n] - ref).abs().max().item() < 2e-5
        assert (x.grad[:n] - xr.grad).abs().max().item() < 2e-5 * max(1.0, xr.grad.abs().max().item())
        for name in ("weight", "bias"):
            a, b = getattr(bn_a, name).grad, getattr(bn_b, name).grad
            assert (a - b).abs().max().item() < 1e-4 * max(1.0, b.abs().max().item()), name
        assert (bn_a.running_mean - bn_b.running_mean).abs().max().item() < 1e-5
        assert (bn_a.running_var - bn_b.running_var).abs().max().item() < 1e-4
        assert int(bn_a.num_batches_tracked) == int(bn_b.num_batches_tracked) == 1
    # second call reuses the self-resetting workspace
    out2 = MF.masked_batch_norm(bn_a, x.detach(), nv)
    assert torch.equal(out2, out.detach())


def test_from_store_padded_replay_follows_exact_eager_steps():
    """An epoch of differently-sized batches through ONE captured graph == the same batches, exactly
    assembled, stepped eagerly."""
    from matdeeplearn_b200 import models as M
    from matdeeplearn_b200.engine import TrainStep
    from matdeeplearn_b200.store import GraphStore
    ds = _dataset("bulk", 64)
    store = GraphStore.from_dataset(ds, DEV)
    torch.manual_seed(0)
    model = M.CGCNN(ds, dim1=64, dim2=64, pre_fc_count=1, gc_count=3, post_fc_count=2)
    m1, m2 = copy.deepcopy(model).to(DEV).train(), copy.deepcopy(model).to(DEV).train()
    s1, s2 = TrainStep(m1, lr=1e-3), TrainStep(m2, lr=1e-3)
    order = np.random.default_rng(2).permutation(64)
    chunks = [order[i:i + 16] for i in range(0, 64, 16)] * 2
    la = [s1.from_store(store, idx) for idx in chunks]
    lb = [float(s2.eager(store.batch(idx)).item()) for idx in chunks]
    sizes = {store._meta(idx)[2:] for idx in chunks}
    assert len(sizes) > 1                      # the batches really differ in shape
    assert len(s1._store_graphs) == 1          # ...and shared one captured graph
    for a, b in zip(la, lb):
        assert abs(a - b) <= 2e-5 * max(1.0, abs(b)), (la, lb)
    assert (s1.flat.param - s2.flat.param).abs().max().item() < 5e-5
    for (n1, b1), (n2, b2) in zip(m1.named_buffers(), m2.named_buffers()):
        assert (b1.float() - b2.float()).abs().max().item() < 1e-4, n1


def test_from_store_over_capacity_batch_takes_the_exact_path():
    from matdeeplearn_b200 import models as M
    from matdeeplearn_b200.engine import TrainStep
    from matdeeplearn_b200.store import GraphStore
    ds = _dataset("bulk", 64)
    store = GraphStore.from_dataset(ds, DEV)
    torch.manual_seed(0)
    model = M.CGCNN(ds, dim1=64, dim2=64, pre_fc_count=1, gc_count=2, post_fc_count=1).to(DEV).train()
    step = TrainStep(model, lr=1e-3)
    small = np.argsort(store.n_nodes)[:8]
    big = np.argsort(store.n_nodes)[-8:]
    store.capacities = lambda B, **kw: (int(store.n_nodes[small].sum()) + 8, int(store.n_edges[small].sum()) + 64)
    l0 = step.from_store(store, small)
    l1 = step.from_store(store, big)           # does not fit: eager on the exact batch
    assert np.isfinite(l0) and np.isfinite(l1)
    static = list(step._store_graphs.values())[0][0]
    assert not store.load(static, big)


@pytest.mark.parametrize("name,kind,cfg", [
    ("SchNet", "bulk", dict(dim1=128, dim2=64, dim3=128, cutoff=8, pre_fc_count=1, gc_count=2, post_fc_count=1)),
    ("SchNet", "bulk", dict(dim1=32, dim2=32, dim3=48, cutoff=8, pre_fc_count=1, gc_count=2, post_fc_count=1)),
    ("MPNN", "bulk", dict(dim1=32, dim2=32, dim3=32, pre_fc_count=1, gc_count=2, post_fc_count=1)),
    ("MEGNet", "bulk", dict(dim1=32, dim2=32, dim3=32, pre_fc_count=1, gc_count=2, gc_fc_count=1, post_fc_count=1)),
    ("MEGNet", "mof", dict(dim1=128, dim2=64, dim3=128, pre_fc_count=1, gc_count=2, gc_fc_count=2, post_fc_count=1)),
])
def test_from_store_padded_replay_other_model_families(name, kind, cfg):
    """SchNet / MPNN / MEGNet on capacity-padded batches through ONE captured graph (edge-level and node-level
    BatchNorm masked by the device-side row counts, padded edges inert).  Two checks:
    (1) the replayed trajectory equals eager steps on the same padded buffers (same kernels, same arithmetic: any
        stale memo or missing in-graph recomputation would show here);
    (2) a step on the padded batch is the same function as a step on the exactly assembled batch (reference train()
        body, training.py:37-50): equal loss, gradients equal in norm.  These are ReLU + BatchNorm networks: a forward
        difference of 1e-6 (different summation order of the batch statistics) can flip a unit sitting at its kink,
        which changes single gradient elements by their full value, so (2) is norm-wise and one step long."""
    from matdeeplearn_b200 import models as M
    from matdeeplearn_b200.engine import TrainStep
    from matdeeplearn_b200.store import GraphStore
    n, B = (48, 12) if kind == "bulk" else (12, 4)
    ds = _dataset(kind, n)
    store = GraphStore.from_dataset(ds, DEV)
    torch.manual_seed(0)
    model = getattr(M, name)(ds, **cfg)
    m1, m2, m3 = (copy.deepcopy(model).to(DEV).train() for _ in range(3))
    s1, s2, s3 = TrainStep(m1, lr=1e-3), TrainStep(m2, lr=1e-3), TrainStep(m3, lr=1e-3)
    order = np.random.default_rng(4).permutation(n)
    chunks = [order[i:i + B] for i in range(0, n, B)]
    assert len({store._meta(idx)[2:] for idx in chunks}) > 1
    # (1) graph replay vs eager on the same padded buffers
    la = [s1.from_store(store, idx) for idx in chunks]
    assert len(s1._store_graphs) == 1
    static = store.static_batch(B, lazy=True)
    lb = []
    for idx in chunks:
        assert store.load(static, idx)
        store.assemble(static)
        lb.append(float(s2.eager(static).item()))
    for a, b in zip(la, lb):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (la, lb)
    assert (s1.flat.param - s2.flat.param).abs().max().item() <= 1e-6 * max(1.0, s2.flat.param.abs().max().item())
    for (n1, b1), (n2, b2) in zip(m1.named_buffers(), m2.named_buffers()):
        assert (b1.float() - b2.float()).abs().max().item() <= 1e-5 * max(1.0, b2.float().abs().max().item()), n1
    # (2) one step: padded batch vs exactly assembled batch
    m4 = copy.deepcopy(model).to(DEV).train()
    s4 = TrainStep(m4, lr=1e-3)
    assert store.load(static, chunks[0])
    store.assemble(static)
    lp = s4._fwd_bwd(static)
    le = s3._fwd_bwd(store.batch(chunks[0]))
    assert abs(lp.item() - le.item()) <= 1e-5 * max(1.0, abs(le.item()))
    gp, ge = s4.flat.grad.double(), s3.flat.grad.double()
    assert (gp - ge).norm().item() <= 2e-2 * ge.norm().item(), ((gp - ge).norm().item(), ge.norm().item())
