"""GPU parity of the four model families against the fixtures produced by the
reference's own model files (tests/golden/make_golden.py; fp64, PyG ops bound to
the oracle), and of the individual SchNet / NNConv / MEGNet operators against the
fp64 oracle."""
import numpy as np
import pytest
import torch

from tests.golden.make_golden import MODEL_CFGS
from tests.test_oracle_golden import _DS, load_batch, load_model_fixture
from tests.util import assert_close, random_graph, contiguous_batch_vector

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _log(tag, got, ref):
    import os
    err = (got.detach().double().cpu() - ref.detach().double().cpu()).abs().max().item()
    scale = ref.detach().abs().max().item()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_errors.txt", "a") as f:
        f.write(f"{tag}: max|err|={err:.3e} scale={scale:.3e} rel={err / max(scale, 1e-30):.3e}\n")


def _maxerr(a, b):
    return (a.detach().double().cpu() - b.detach().double().cpu()).abs().max().item()


@pytest.mark.parametrize("tag", list(MODEL_CFGS))
def test_model_matches_reference_glue_fixture(tag):
    """fp32 engine vs the fp64 fixture of the reference's own model file.  Deep BatchNorm stacks
    on a 6-graph batch amplify fp32 rounding (the global model normalises over 6 rows), so each
    tensor must be as close to the fp64 truth as the fp32 CPU oracle is (x4), or within the
    end-of-model fp32 tolerance of SURVEY.md 8c (1e-4 of the output scale), whichever is larger."""
    from matdeeplearn_b200 import models as M
    from oracle import models as OM
    b = load_batch()
    z, sd, grads = load_model_fixture(tag)
    sd32 = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()}
    cls = tag.split("_")[0]
    model = getattr(M, cls)(_DS(b), **MODEL_CFGS[tag])
    model.load_state_dict(sd32)
    model = model.to(DEV).train()
    o32 = getattr(OM, cls)(_DS(b), **MODEL_CFGS[tag])
    o32.load_state_dict(sd32)
    o32.train()
    gb = b.to(DEV)
    out = model(gb)
    out32 = o32(b)
    ref = torch.from_numpy(z["out_train"])
    _log(f"model {tag} out_train", out, ref)
    scale = ref.abs().max().item()
    assert _maxerr(out, ref) <= max(1e-4 * scale, 4 * _maxerr(out32, ref)), (tag, _maxerr(out, ref), _maxerr(out32, ref))
    loss = torch.nn.functional.l1_loss(out, gb.y)
    loss.backward()
    torch.nn.functional.l1_loss(out32, b.y).backward()
    gscale = max(r.abs().max().item() for r in grads.values() if r.numel())
    o32_grads = dict(o32.named_parameters())
    for name, p in model.named_parameters():
        r = grads[name]
        if r.numel() == 0:
            continue
        _log(f"model {tag} grad {name}", p.grad, r)
        # gradients through BatchNorm over 6 rows (MEGNet's global model) are ill-conditioned: the fp32
        # CPU oracle itself sits 2e-5..1e-4 from the fp64 fixture depending on the host's BLAS, and
        # different summation orders move the engine by the same factor (measured 1.2e-4 on one box)
        tol = 2e-4 * r.abs().max().item() + 2e-6 * gscale
        e_got, e_o32 = _maxerr(p.grad, r), _maxerr(o32_grads[name].grad, r)
        assert e_got <= max(tol, 8 * e_o32), (tag, name, e_got, e_o32, tol)
    model.eval()
    o32.eval()
    with torch.no_grad():
        ev, ev32 = model(gb), o32(b)
    refe = torch.from_numpy(z["out_eval"])
    assert _maxerr(ev, refe) <= max(1e-4 * refe.abs().max().item(), 4 * _maxerr(ev32, refe)), tag


def _graph(n, e, seed):
    ei = random_graph(n, e, seed)
    return ei, ei.shape[1]


def test_interaction_block_matches_oracle():
    import matdeeplearn_b200.nn as mnn
    from oracle import pyg_ops as O
    torch.manual_seed(0)
    n, C, G, Fi = 300, 128, 50, 128
    ei, E = _graph(n, 3000, 3)
    x = torch.randn(n, C, dtype=torch.float64)
    ea = torch.rand(E, G, dtype=torch.float64)
    ew = torch.rand(E, dtype=torch.float64) * 8.0
    ref_m = O.InteractionBlock(C, G, Fi, 8.0).double()
    m = mnn.InteractionBlock(C, G, Fi, 8.0)
    m.load_state_dict({k: v.float() for k, v in ref_m.state_dict().items()})
    m = m.to(DEV)
    xr = x.clone().requires_grad_(True)
    ref = ref_m(xr, ei, ew, ea)
    xg = x.float().to(DEV).requires_grad_(True)
    got = m(xg, ei.to(DEV), ew.float().to(DEV), ea.float().to(DEV))
    _log("InteractionBlock fwd", got, ref)
    assert_close(got, ref, rtol=1e-5, atol_rel=5e-6, what="InteractionBlock fwd")
    w = torch.randn_like(ref)
    ref.backward(w)
    got.backward(w.float().to(DEV))
    assert_close(xg.grad, xr.grad, rtol=1e-4, atol_rel=2e-5, what="InteractionBlock dx")
    for name, pr in ref_m.named_parameters():
        pg = dict(m.named_parameters())[name]
        _log(f"InteractionBlock d{name}", pg.grad, pr.grad)
        assert_close(pg.grad, pr.grad, rtol=1e-4, atol_rel=2e-5, what=f"InteractionBlock d{name}")


@pytest.mark.parametrize("n,e,iso", [(300, 3000, 0), (200, 900, 6)])
def test_gcnconv_matches_oracle(n, e, iso):
    """GCNConv(improved=True, add_self_loops=False) with distance-like edge weights, zero on the loops
    (reference gcn.py:80-82, 141); isolated nodes have degree 0 -> coefficient 0, output = bias."""
    import matdeeplearn_b200.nn as mnn
    from oracle import pyg_ops as O
    torch.manual_seed(2)
    ei = random_graph(n, e, 7, None, iso)
    E = ei.shape[1]
    x = torch.randn(n, 48, dtype=torch.float64)
    ew = torch.rand(E, dtype=torch.float64) * 8.0
    ew[ei[0] == ei[1]] = 0.0
    ref_m = O.GCNConv(48, 48, improved=True, add_self_loops=False).double()
    with torch.no_grad():
        ref_m.bias.uniform_(-0.5, 0.5)
    m = mnn.GCNConv(48, 48, improved=True, add_self_loops=False)
    m.load_state_dict({k: v.float() for k, v in ref_m.state_dict().items()})
    m = m.to(DEV)
    xr = x.clone().requires_grad_(True)
    ref = ref_m(xr, ei, ew)
    xg = x.float().to(DEV).requires_grad_(True)
    got = m(xg, ei.to(DEV), ew.float().to(DEV))
    _log("GCNConv fwd", got, ref)
    assert_close(got, ref, rtol=1e-5, atol_rel=5e-6, what="GCNConv fwd")
    w = torch.randn_like(ref)
    ref.backward(w)
    got.backward(w.float().to(DEV))
    assert_close(xg.grad, xr.grad, rtol=1e-4, atol_rel=2e-5, what="GCNConv dx")
    for name, pr in ref_m.named_parameters():
        pg = dict(m.named_parameters())[name]
        assert_close(pg.grad, pr.grad, rtol=1e-4, atol_rel=2e-5, what=f"GCNConv d{name}")


@pytest.mark.parametrize("C,K,G", [(64, 64, 50), (16, 20, 37), (100, 100, 50)])
def test_nnconv_matches_oracle(C, K, G):
    import matdeeplearn_b200.nn as mnn
    from oracle import pyg_ops as O
    torch.manual_seed(1)
    n = 200
    ei, E = _graph(n, 1500, 4)
    x = torch.randn(n, C, dtype=torch.float64)
    ea = torch.rand(E, G, dtype=torch.float64)

    def net():
        return torch.nn.Sequential(torch.nn.Linear(G, K), torch.nn.ReLU(), torch.nn.Linear(K, C * C))

    ref_m = O.NNConv(C, C, net(), aggr="mean").double()
    m = mnn.NNConv(C, C, net(), aggr="mean")
    m.load_state_dict({k: v.float() for k, v in ref_m.state_dict().items()})
    m = m.to(DEV)
    xr = x.clone().requires_grad_(True)
    ref = ref_m(xr, ei, ea)
    xg = x.float().to(DEV).requires_grad_(True)
    got = m(xg, ei.to(DEV), ea.float().to(DEV))
    _log(f"NNConv fwd C={C}", got, ref)
    assert_close(got, ref, rtol=1e-5, atol_rel=5e-6, what="NNConv fwd")
    w = torch.randn_like(ref)
    ref.backward(w)
    got.backward(w.float().to(DEV))
    assert_close(xg.grad, xr.grad, rtol=1e-4, atol_rel=2e-5, what="NNConv dx")
    for name, pr in ref_m.named_parameters():
        pg = dict(m.named_parameters())[name]
        _log(f"NNConv d{name} C={C}", pg.grad, pr.grad)
        assert_close(pg.grad, pr.grad, rtol=1e-4, atol_rel=2e-5, what=f"NNConv d{name}")


def test_metalayer_fused_edge_model_equals_pyg_call_convention():
    """forward_fused (split weights + gather-add kernel) == the PyG-style call on gathered,
    concatenated inputs, for values and gradients."""
    from matdeeplearn_b200 import models as M
    torch.manual_seed(2)
    n, D, B = 120, 32, 5
    ei, E = _graph(n, 900, 5)
    batch = contiguous_batch_vector(n, B, 6)
    x = torch.randn(n, D)
    e = torch.randn(E, D)
    u = torch.randn(B, D)
    model = M.Megnet_EdgeModel(D, "relu", "True", True, 0.0, 2).to(DEV).train()
    xs = [t.to(DEV).requires_grad_(True) for t in (x, e, u)]
    eid, bd = ei.to(DEV), batch.to(DEV)
    out_f = model.forward_fused(xs[0], eid, xs[1], xs[2], bd)
    g = torch.randn_like(out_f)
    out_f.backward(g)
    grads_f = [t.grad.clone() for t in xs] + [p.grad.clone() for p in model.parameters()]
    for t in xs:
        t.grad = None
    model.zero_grad()
    out_p = model(xs[0][eid[0]], xs[0][eid[1]], xs[1], xs[2], bd[eid[0]])
    out_p.backward(g)
    grads_p = [t.grad for t in xs] + [p.grad for p in model.parameters()]
    assert_close(out_f, out_p, rtol=1e-5, atol_rel=2e-6, what="fused edge model fwd")
    for a, b in zip(grads_f, grads_p):
        assert_close(a, b, rtol=1e-4, atol_rel=2e-5, what="fused edge model grads")


def test_config0_test_data_cgcnn_demo_batch32():
    """BASELINE configs[0]: real data/test_data Pt clusters, batch 32, default CGCNN_demo
    (dim1=100 -> the fused kernel's SIMT path).  Fixture: tests/golden/make_golden.py."""
    import os
    from matdeeplearn_b200 import models as M
    from matdeeplearn_b200.data import Batch
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "testdata_cgcnn_demo_b32.npz"))
    b = Batch(**{k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("in/")})
    b.num_graphs = 32
    cfg = dict(dim1=100, dim2=150, pre_fc_count=1, gc_count=4, post_fc_count=3)
    model = M.CGCNN(_DS(b), **cfg)
    model.load_state_dict({k[6:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")})
    model = model.to(DEV).train()
    gb = b.to(DEV)
    out = model(gb)
    ref = torch.from_numpy(z["out_train"])
    _log("config0 test_data out", out, ref)
    assert_close(out, ref, rtol=1e-4, atol_rel=5e-5, what="config0 forward")
    torch.nn.functional.l1_loss(out, gb.y).backward()
    gscale = max(float(np.abs(z[k]).max()) for k in z.files if k.startswith("grad/"))
    for name, p in model.named_parameters():
        r = torch.from_numpy(z["grad/" + name])
        err = _maxerr(p.grad, r)
        assert err <= 2e-3 * float(r.abs().max()) + 2e-5 * gscale, (name, err)


def test_gaussian_smear_kernel_matches_reference_module():
    from matdeeplearn_b200 import functional as MF
    pg = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "process_golden.npz"))
    d = torch.from_numpy(pg["smear_in"]).to(DEV)
    for G in (50, 100, 200):
        off = torch.from_numpy(pg[f"smear_offset_{G}"]).to(DEV)
        got = MF.gaussian_smear(d, off, float(pg[f"smear_coeff_{G}"]))
        ref = torch.from_numpy(pg[f"smear_{G}"])
        # same formula in fp32; only expf's last-bit rounding may differ between CPU and GPU libm
        assert (got.cpu() - ref).abs().max().item() <= 2e-7
