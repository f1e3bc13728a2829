"""GPU parity at the BASELINE.json shapes (configs[1..4]), model level, forward + every parameter gradient,
against the fp64 oracle on the same seeded inputs:

  configs[1]  CGCNN   dim 64,  4 CGConv,           256 synthetic bulk graphs
  configs[2]  SchNet  dim 128, 4 interactions,     256 synthetic bulk graphs
  configs[3]  MEGNet  dim 128, 3 blocks, gc_fc 2,  64 synthetic MOF-shaped graphs
  configs[4]  MPNN    dim 64,  3 NNConv + GRU,     edge_length in {100, 200} (32 bulk graphs: the oracle
              materialises the [E, C*C] edge-conditioned weight that the engine never forms)

Tolerance: end-of-model output within 1e-4 of the output scale; gradients within
1e-4 * max|ref| of the tensor + 1e-5 of the largest gradient in the model.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _log(tag, err, scale):
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_errors_baseline_shapes.txt", "a") as f:
        f.write(f"{tag}: max|err|={err:.3e} scale={scale:.3e} rel={err / max(scale, 1e-30):.3e}\n")


def _maxerr(a, b):
    return (a.detach().double().cpu() - b.detach().double().cpu()).abs().max().item()


def _check_model(name, kind, graphs, cfg, edge_length=50, seed=7, out_tol=1e-4, g_rel=1e-4, g_abs=1e-5,
                 fp32_yardstick=False):
    from matdeeplearn_b200 import models as M, process as pr
    from oracle import models as OM
    ds = pr.synthetic_dataset(kind, graphs, seed=seed, edge_length=edge_length)
    b = ds.batch()
    b.num_graphs = graphs
    torch.manual_seed(0)
    ref_model = getattr(OM, name)(ds, **cfg)
    model = getattr(M, name)(ds, **cfg)
    model.load_state_dict(ref_model.state_dict())
    o32 = None
    if fp32_yardstick:   # the same oracle in fp32 on the CPU: how far plain fp32 arithmetic sits from the fp64 truth
        import copy
        o32 = copy.deepcopy(ref_model).train()
    ref_model = ref_model.double().train()
    model = model.to(DEV).train()
    gb = b.to(DEV)
    out = model(gb)
    loss = torch.nn.functional.l1_loss(out, gb.y)
    loss.backward()
    torch.cuda.synchronize()
    b64 = b.double()
    ref = ref_model(b64)
    torch.nn.functional.l1_loss(ref, b64.y).backward()
    scale = ref.abs().max().item()
    e = _maxerr(out, ref)
    _log(f"{name} {kind}x{graphs} G={edge_length} out", e, scale)
    assert e <= out_tol * scale, (name, "forward", e, scale)
    ref_grads = {k: p.grad for k, p in ref_model.named_parameters()}
    gscale = max(g.abs().max().item() for g in ref_grads.values() if g is not None and g.numel())
    o32_grads = {}
    if o32 is not None:
        torch.nn.functional.l1_loss(o32(b), b.y).backward()
        o32_grads = {k: p.grad for k, p in o32.named_parameters()}
    worst = (0.0, None)
    for k, p in model.named_parameters():
        r = ref_grads[k]
        if r is None:
            assert p.grad is None or p.grad.abs().max().item() == 0.0, k
            continue
        e = _maxerr(p.grad, r)
        tol = g_rel * r.abs().max().item() + g_abs * gscale
        if o32_grads.get(k) is not None:
            tol = max(tol, _maxerr(o32_grads[k], r))   # never looser than the fp32 CPU oracle's own deviation
        _log(f"{name} {kind}x{graphs} G={edge_length} grad {k}", e, r.abs().max().item())
        if e / tol > worst[0]:
            worst = (e / tol, k)
        assert e <= tol, (name, k, e, tol)
    return worst


def test_cgcnn_config1_shape():
    _check_model("CGCNN", "bulk", 256, dict(dim1=64, dim2=64, pre_fc_count=1, gc_count=4, post_fc_count=1))


def test_schnet_config2_shape():
    _check_model("SchNet", "bulk", 256,
                 dict(dim1=128, dim2=128, dim3=128, cutoff=8, pre_fc_count=1, gc_count=4, post_fc_count=1))


def test_megnet_config3_shape():
    """The reference feeds u = zeros[B,3] (process.py:330-332): every row of the first block's global-state
    MLP is identical, its BatchNorm sees zero variance and amplifies rounding by 1/sqrt(eps) -- the fp32 CPU
    oracle's own gradients sit ~1e-1 (relative) from the fp64 ones on this config.  Output: 1e-4 of scale as
    everywhere; gradients: the usual bound, or the fp32 oracle's own distance from fp64 where that is larger
    (the engine must be at least as close to the truth as fp32 PyTorch on the CPU is)."""
    _check_model("MEGNet", "mof", 64,
                 dict(dim1=128, dim2=128, dim3=128, pre_fc_count=1, gc_count=3, gc_fc_count=2, post_fc_count=1),
                 fp32_yardstick=True)


@pytest.mark.parametrize("G", [100, 200])
def test_mpnn_config4_edge_lengths(G):
    """act="softplus" between the layers: with ReLU everywhere the comparison with fp64 is decided by which side
    of a kink a rounding-sized pre-activation falls on -- measured on this batch (profiles/debug_mpnn_parity2.py):
    one of the 2048 post-FC pre-activations is 5.7e-8, the fp32 engine (error 1.7e-6 there, like the fp32 oracle on
    the GPU) lands on the other side, one ReLU mask flips and every gradient upstream moves by ~1e-2 although
    forward, kernels and autograd all agree to 1e-6.  The NNConv edge network keeps its ReLU (reference mpnn.py:83-85)."""
    _check_model("MPNN", "bulk", 32, dict(dim1=64, dim2=64, dim3=64, pre_fc_count=1, gc_count=3, post_fc_count=1,
                                          act="softplus"),
                 edge_length=G)


@pytest.mark.parametrize("G", [100, 200])
def test_nnconv_operator_wide_edge_attr(G):
    import matdeeplearn_b200.nn as mnn
    from oracle import pyg_ops as O
    from tests.util import assert_close, random_graph
    torch.manual_seed(1)
    n, C, K = 400, 64, 64
    ei = random_graph(n, 4000, 4)
    E = ei.shape[1]
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
    assert_close(got, ref, rtol=1e-5, atol_rel=5e-6, what=f"NNConv G={G} fwd")
    w = torch.randn_like(ref)
    ref.backward(w)
    got.backward(w.float().to(DEV))
    assert_close(xg.grad, xr.grad, rtol=1e-4, atol_rel=2e-5, what=f"NNConv G={G} dx")
    for name, p_ref in ref_m.named_parameters():
        assert_close(dict(m.named_parameters())[name].grad, p_ref.grad, rtol=1e-4, atol_rel=2e-5,
                     what=f"NNConv G={G} d{name}")


def test_metalayer_matches_oracle_metalayer():
    """The engine's MetaLayer (fused edge model: split weights + gather-add kernel, CSR segmented means) against
    the oracle's PyG MetaLayer wired with the oracle's Megnet_* models (reference megnet.py:16-147), D = 128."""
    from matdeeplearn_b200 import models as M
    import matdeeplearn_b200.nn as mnn
    from oracle import models as OM, pyg_ops as O
    from tests.util import assert_close, random_graph, contiguous_batch_vector
    torch.manual_seed(2)
    n, D, B = 600, 128, 9
    batch = contiguous_batch_vector(n, B, 6)
    # edges inside graphs only, one self-loop per node (every node is a source: scatter_mean(e, row) covers N rows)
    cols = []
    rng = torch.Generator().manual_seed(3)
    for g in range(B):
        idx = (batch == g).nonzero().view(-1)
        k = idx.numel()
        src = idx[torch.randint(k, (6 * k,), generator=rng)]
        dst = idx[torch.randint(k, (6 * k,), generator=rng)]
        keep = src != dst
        pairs = torch.unique(torch.stack([src[keep], dst[keep]]), dim=1)
        cols += [pairs, torch.stack([idx, idx])]
    ei = torch.cat(cols, 1)
    E = ei.shape[1]
    x, e, u = torch.randn(n, D), torch.randn(E, D), torch.randn(B, D)
    args = (D, "relu", "True", True, 0.0, 2)
    ref_layer = O.MetaLayer(OM.Megnet_EdgeModel(*args), OM.Megnet_NodeModel(*args), OM.Megnet_GlobalModel(*args))
    layer = mnn.MetaLayer(M.Megnet_EdgeModel(*args), M.Megnet_NodeModel(*args), M.Megnet_GlobalModel(*args))
    layer.load_state_dict(ref_layer.state_dict())
    ref_layer = ref_layer.double().train()
    layer = layer.to(DEV).train()
    ins = [t.to(DEV).requires_grad_(True) for t in (x, e, u)]
    rins = [t.double().requires_grad_(True) for t in (x, e, u)]
    got = layer(ins[0], ei.to(DEV), ins[1], ins[2], batch.to(DEV))
    ref = ref_layer(rins[0], ei, rins[1], rins[2], batch)
    ws = [torch.randn_like(r) for r in ref]
    sum((g * w.float().to(DEV)).sum() for g, w in zip(got, ws)).backward()
    sum((r * w).sum() for r, w in zip(ref, ws)).backward()
    for name, g, r in zip("xeu", got, ref):
        assert_close(g, r, rtol=1e-4, atol_rel=2e-5, what=f"MetaLayer {name}'")
    for name, g, r in zip("xeu", ins, rins):
        assert_close(g.grad, r.grad, rtol=1e-4, atol_rel=5e-5, what=f"MetaLayer d{name}")
    rp = dict(ref_layer.named_parameters())
    for k, p in layer.named_parameters():
        assert_close(p.grad, rp[k].grad, rtol=1e-4, atol_rel=5e-5, what=f"MetaLayer d{k}")


REF_MODELS = "/root/reference/matdeeplearn/models"


@pytest.mark.skipif(not os.path.isdir(REF_MODELS), reason="reference checkout not present on this box")
@pytest.mark.parametrize("fname,cls,tag", [("cgcnn.py", "CGCNN", "CGCNN"), ("schnet.py", "SchNet", "SchNet")])
def test_reference_model_file_runs_on_the_engine_shim(fname, cls, tag):
    """the reference's OWN model file, imported with torch_geometric / torch_scatter
    bound to matdeeplearn_b200.nn, reproduces the fixture its file produced on the oracle ops."""
    import importlib.util
    import sys
    import types
    import matdeeplearn_b200.nn as mnn
    from tests.golden.make_golden import MODEL_CFGS
    from tests.test_oracle_golden import _DS, load_batch, load_model_fixture
    tg = types.ModuleType("torch_geometric")
    tg.nn = mnn
    tg_models = types.ModuleType("torch_geometric.nn.models")
    tg_schnet = types.ModuleType("torch_geometric.nn.models.schnet")
    tg_schnet.InteractionBlock = mnn.InteractionBlock
    ts = types.ModuleType("torch_scatter")
    for name in ("scatter", "scatter_mean", "scatter_add", "scatter_max"):
        setattr(ts, name, getattr(mnn, name))
    shim = {"torch_geometric": tg, "torch_geometric.nn": mnn, "torch_geometric.nn.models": tg_models,
            "torch_geometric.nn.models.schnet": tg_schnet, "torch_scatter": ts}
    saved = {k: sys.modules.get(k) for k in shim}
    sys.modules.update(shim)
    try:
        spec = importlib.util.spec_from_file_location("ref_" + fname[:-3], os.path.join(REF_MODELS, fname))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        b = load_batch()
        z, sd, grads = load_model_fixture(tag)
        model = getattr(mod, cls)(_DS(b), **MODEL_CFGS[tag])
        model.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in sd.items()})
        model = model.to(DEV).train()
        gb = b.to(DEV)
        out = model(gb)
        ref = torch.from_numpy(z["out_train"])
        assert _maxerr(out, ref) <= 2e-4 * ref.abs().max().item(), (tag, _maxerr(out, ref))
        torch.nn.functional.l1_loss(out, gb.y).backward()
        gscale = max(r.abs().max().item() for r in grads.values() if r.numel())
        for name, p in model.named_parameters():
            r = grads[name]
            if r.numel():
                assert _maxerr(p.grad, r) <= 5e-4 * r.abs().max().item() + 1e-5 * gscale, (tag, name)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
