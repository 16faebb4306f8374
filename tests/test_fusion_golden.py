"""Fusion networks against golden vectors produced by the REFERENCE's own `KPFCNN_featureAggre` classes
(tests/golden/make_golden_fusion.py imports models/architectures_sphere*.py from /root/reference and runs them on
the CPU): same state dict (loaded with strict=True: the module / parameter names are the reference's), same batch.

  * CPU (`-m "not gpu"`): fusion.FusionKPFCNN composed from the oracle operators reproduces the reference logits,
    loss and every parameter gradient -> pins the wiring of fusion.py (lifting loop, early / middle / late fusion,
    decoder bookkeeping, head) and the oracle modules on the reference.
  * GPU (`-m gpu`): the product operators (strict fp32 contraction: 1e-4; bf16x3 stated separately) on the same
    inputs, lifting indices recomputed on the GPU (batched unprojection + grid 3-NN) and required to equal the
    reference's sklearn indices.
"""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import geom, modules

ARCHS = {
    "early": ['simple', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_strided', 'resnetb',
              'nearest_upsample', 'unary', 'nearest_upsample', 'unary'],
    "middle": ['simple', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_deformable_strided', 'resnetb_deformable',
               'nearest_upsample', 'unary', 'nearest_upsample', 'unary'],
    "late": ['simple', 'resnetb', 'resnetb_deformable_strided', 'resnetb_deformable', 'nearest_upsample', 'unary'],
}


class Feature2DStub(torch.nn.Module):
    """The stand-in 2D network of the golden run (same definition as in make_golden_fusion.py)."""

    def __init__(self):
        super().__init__()
        self.c1 = torch.nn.Conv2d(3, 16, 3, padding=1)
        self.c2 = torch.nn.Conv2d(16, 64, 3, padding=1)

    def forward(self, d):
        return {'feature': self.c2(torch.relu(self.c1(d['image'])))}


def rel_err(a, b):
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    b = b.detach().double().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def load(which):
    z = np.load(os.path.join(GOLDEN, f"fusion_{which}.npz"), allow_pickle=False)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    grads = {k[5:]: z[k] for k in z.files if k.startswith("grad/")}
    return z, sd, grads


def config(which):
    from mvkpconv_b200 import fusion as fu
    return fu.fusion_config(which, architecture=list(ARCHS[which]), first_subsampling_dl=0.06,
                            first_features_dim=128 if which == "late" else 16, num_classes=6, deform_radius=4.0)


@pytest.mark.parametrize("which", ["early", "middle", "late"])
def test_fusion_oracle_graph_reproduces_reference_golden(which):
    from mvkpconv_b200 import fusion as fu, pyramid
    z, sd, grads = load(which)
    cfg = config(which)
    gops = SimpleNamespace(batch_neighbors=geom.batch_neighbors,
                           batch_grid_subsampling=lambda p, l, sampleDl=0.1, random_grid_orient=True:
                           geom.grid_subsample_batch(p, l, sampleDl=sampleDl))
    mops = SimpleNamespace(KPConv=modules.KPConvOracle, max_pool=modules.max_pool, closest_pool=modules.closest_pool,
                           group_points=modules.group_points, FeatureAggregation=modules.FeatureAggregationOracle)
    np.random.seed(0)
    torch.manual_seed(0)
    net = fu.FusionKPFCNN(cfg, fusion=which, net_2d=Feature2DStub(), ops=mops)
    net.load_state_dict(sd, strict=True)
    net.train()
    lens = z["lens"]
    pyr = pyramid.build_pyramid(z["centred"], lens, cfg, ops=gops, random_grid_orient=False)
    as_t = lambda lst, dt: [torch.from_numpy(np.ascontiguousarray(a)).to(dt) for a in lst]
    starts = np.concatenate([[0], np.cumsum(lens)])
    knn_list = [z["knn"][starts[i]:starts[i + 1]][None] for i in range(len(lens))]
    batch = SimpleNamespace(points=as_t(pyr.points, torch.float32), neighbors=as_t(pyr.neighbors, torch.int64),
                            pools=as_t(pyr.pools, torch.int64), upsamples=as_t(pyr.upsamples, torch.int64),
                            lengths=pyr.lengths, images=torch.from_numpy(z["images"]), image_xyz=torch.from_numpy(z["image_xyz"]),
                            knn_list=knn_list, feat_aggre_points=torch.from_numpy(z["world"])[None],
                            feature_3d=torch.from_numpy(z["feature_3d"]))
    out = net(batch)
    loss = net.loss(out, torch.from_numpy(z["labels"]))
    loss.backward()
    assert rel_err(out, z["logits"]) < 1e-5
    assert abs(float(loss.detach()) - float(z["loss"])) < 1e-5 * abs(float(z["loss"]))
    got = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    assert set(got) == set(grads), set(got) ^ set(grads)
    for k, g in grads.items():
        if np.abs(g).max() > 1e-7:
            assert rel_err(got[k], g) < 1e-4, k


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["early", "middle", "late"])
@pytest.mark.parametrize("contraction", ["fp32", "bf16x3"])
def test_fusion_product_vs_reference_golden(mvk, which, contraction):
    from mvkpconv_b200 import fusion as fu, pyramid
    z, sd, grads = load(which)
    cfg = config(which)
    dev = torch.device("cuda")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    np.random.seed(0)
    torch.manual_seed(0)
    net = fu.FusionKPFCNN(cfg, fusion=which, net_2d=Feature2DStub()).to(dev)
    net.load_state_dict(sd, strict=True)
    for m in net.modules():
        if hasattr(m, "contraction"):
            m.contraction = contraction
    net.feat_aggreg.contraction = contraction
    net.train()
    lens = z["lens"]
    world = torch.from_numpy(z["world"]).to(dev)
    # lifting on the GPU: unprojection bit-equal to the reference's image_xyz, 3-NN equal to sklearn's indices
    xyz32, mask, knn_g = fu.prepare_lifting(z["cams"], z["depths"], z["poses"], world, lens)
    assert np.array_equal(xyz32.cpu().numpy(), z["image_xyz"])
    npix = z["image_xyz"].shape[1] * z["image_xyz"].shape[2] * z["image_xyz"].shape[3]
    starts = np.concatenate([[0], np.cumsum(lens)])
    ref_global = np.concatenate([z["knn"][starts[i]:starts[i + 1]] + i * npix for i in range(len(lens))], 0)
    assert np.array_equal(knn_g.cpu().numpy(), ref_global)
    pyr = pyramid.build_pyramid(torch.from_numpy(z["centred"]).to(dev), torch.from_numpy(lens).to(dev), cfg,
                                random_grid_orient=False)
    batch = SimpleNamespace(points=pyr.points, neighbors=pyr.neighbors, pools=pyr.pools, upsamples=pyr.upsamples,
                            lengths=pyr.lengths, images=torch.from_numpy(z["images"]).to(dev), image_xyz=xyz32,
                            knn_global=knn_g, feat_aggre_points=world[None], feature_3d=torch.from_numpy(z["feature_3d"]).to(dev))
    out = net(batch)
    loss = net.loss(out, torch.from_numpy(z["labels"]).to(dev))
    loss.backward()
    strict = contraction == "fp32"
    e_out = rel_err(out, z["logits"])
    e_loss = abs(float(loss.detach()) - float(z["loss"])) / abs(float(z["loss"]))
    got = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    num = den = 0.0
    worst = ("", 0.0)
    for k, g in grads.items():
        assert k in got, k
        d = got[k].detach().double().cpu().numpy() - g.astype(np.float64)
        num += float((d * d).sum())
        den += float((g.astype(np.float64) ** 2).sum())
        if np.abs(g).max() > 1e-6:
            e = rel_err(got[k], g)
            worst = max(worst, (k, e), key=lambda t: t[1])
    l2 = (num / den) ** 0.5
    print(f"fusion {which} [{contraction}] vs reference golden: logits {e_out:.2e} loss {e_loss:.2e} grads L2 {l2:.2e} "
          f"worst tensor {worst[0]} {worst[1]:.2e}")
    assert e_out < (1e-4 if strict else 2e-3)
    assert e_loss < (1e-4 if strict else 2e-3)
    assert l2 < (1e-4 if strict else 2e-2)
    if strict:
        assert worst[1] < 1e-3, worst


# -------------------------------------------------------------------------------------------------
# The baseline KPFCNN (BASELINE configs[1] architecture family) against the reference's own class
# (models/architectures.py:189-352 incl. p2p_fitting_regularizer :21-54), golden = tests/golden/kpfcnn_baseline.npz
# -------------------------------------------------------------------------------------------------
def _baseline():
    from mvkpconv_b200 import pyramid
    z = np.load(os.path.join(GOLDEN, "kpfcnn_baseline.npz"), allow_pickle=False)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    grads = {k[5:]: z[k] for k in z.files if k.startswith("grad/")}
    cfg = pyramid.baseline_config(architecture=list(ARCHS["middle"]), first_subsampling_dl=0.03, first_features_dim=16,
                                  num_classes=6, in_features_dim=2, deform_radius=4.0)
    return z, sd, grads, cfg


def test_kpfcnn_oracle_graph_reproduces_reference_golden():
    """harness.KPFCNN composed from the oracle operators == the reference's KPFCNN (same state dict, strict): logits,
    loss (cross entropy + deformable fitting / repulsion regulariser) and every parameter gradient.  This is the
    graph bench.py's CPU arm runs, so the arm's network half is pinned on the reference as well."""
    from mvkpconv_b200 import harness, pyramid
    z, sd, grads, cfg = _baseline()
    gops = SimpleNamespace(batch_neighbors=geom.batch_neighbors,
                           batch_grid_subsampling=lambda p, l, sampleDl=0.1, random_grid_orient=True:
                           geom.grid_subsample_batch(p, l, sampleDl=sampleDl))
    mops = SimpleNamespace(KPConv=modules.KPConvOracle, max_pool=modules.max_pool, closest_pool=modules.closest_pool)
    np.random.seed(0)
    torch.manual_seed(0)
    net = harness.KPFCNN(cfg, ops=mops)
    net.load_state_dict(sd, strict=True)
    net.train()
    pyr = pyramid.build_pyramid(z["points"], z["lens"], cfg, ops=gops, random_grid_orient=False)
    as_t = lambda lst, dt: [torch.from_numpy(np.ascontiguousarray(a)).to(dt) for a in lst]
    batch = SimpleNamespace(points=as_t(pyr.points, torch.float32), neighbors=as_t(pyr.neighbors, torch.int64),
                            pools=as_t(pyr.pools, torch.int64), upsamples=as_t(pyr.upsamples, torch.int64),
                            lengths=pyr.lengths, features=torch.from_numpy(z["features"]))
    out = net(batch)
    loss = net.loss(out, torch.from_numpy(z["labels"]))
    loss.backward()
    assert rel_err(out, z["logits"]) < 1e-5
    assert abs(float(loss.detach()) - float(z["loss"])) < 1e-5 * abs(float(z["loss"]))
    got = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    assert set(got) == set(grads), set(got) ^ set(grads)
    for k, g in grads.items():
        if np.abs(g).max() > 1e-7:
            assert rel_err(got[k], g) < 1e-4, k


@pytest.mark.gpu
@pytest.mark.parametrize("contraction", ["fp32", "bf16x3"])
def test_kpfcnn_product_vs_reference_golden(mvk, contraction):
    from mvkpconv_b200 import harness, pyramid
    z, sd, grads, cfg = _baseline()
    dev = torch.device("cuda")
    np.random.seed(0)
    torch.manual_seed(0)
    net = harness.KPFCNN(cfg).to(dev)
    net.load_state_dict(sd, strict=True)
    for m in net.modules():
        if hasattr(m, "contraction"):
            m.contraction = contraction
    net.train()
    pyr = pyramid.build_pyramid(torch.from_numpy(z["points"]).to(dev), torch.from_numpy(z["lens"]).to(dev), cfg,
                                random_grid_orient=False)
    batch = SimpleNamespace(points=pyr.points, neighbors=pyr.neighbors, pools=pyr.pools, upsamples=pyr.upsamples,
                            lengths=pyr.lengths, features=torch.from_numpy(z["features"]).to(dev))
    out = net(batch)
    loss = net.loss(out, torch.from_numpy(z["labels"]).to(dev))
    loss.backward()
    strict = contraction == "fp32"
    e_out = rel_err(out, z["logits"])
    e_loss = abs(float(loss.detach()) - float(z["loss"])) / abs(float(z["loss"]))
    got = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    num = den = 0.0
    for k, g in grads.items():
        d = got[k].detach().double().cpu().numpy() - g.astype(np.float64)
        num += float((d * d).sum())
        den += float((g.astype(np.float64) ** 2).sum())
    l2 = (num / den) ** 0.5
    print(f"KPFCNN [{contraction}] vs reference golden: logits {e_out:.2e} loss {e_loss:.2e} grads L2 {l2:.2e}")
    assert e_out < (1e-4 if strict else 2e-3)
    assert e_loss < (1e-4 if strict else 2e-3)
    assert l2 < (1e-4 if strict else 2e-2)
