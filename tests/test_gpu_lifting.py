"""GPU parity: unprojection, point-to-pixel k-NN, group_points and FeatureAggregation."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import modules

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def test_unproject_and_knn_vs_reference_golden(mvk):
    g = load_golden("lifting")
    xyz32, mask, xyz64 = mvk.unproject_views(g["cam_matrix"], g["depth"], g["pose"])
    nv, h, w = g["depth"].shape
    assert np.array_equal(mask.reshape(nv, -1), g["image_mask"])
    # fp64 arithmetic may differ from numpy/BLAS by an fma contraction: <= a few ulp(fp64) ...
    assert np.abs(xyz64.reshape(nv, -1, 3) - g["image_xyz_f64"]).max() < 1e-12
    # ... so the fp32 tensor the reference stacks into the batch is reproduced bit for bit
    assert np.array_equal(xyz32.reshape(nv, -1, 3), g["image_xyz_f32"])
    knn = mvk.knn_pixels(xyz64, mask, g["queries"], k=3)
    assert knn.dtype == np.int64
    assert np.array_equal(knn, g["knn_indices"])  # sklearn ball_tree result of the reference pipeline
    d2xyz = mvk.depth2xyz(g["cam_matrix"], g["depth"][0])
    ref = modules.depth2xyz(g["cam_matrix"], g["depth"][0])
    assert np.abs(d2xyz - ref).max() < 1e-12


def test_knn_vs_oracle_full_view_size(mvk):
    """3 views of 160x120 against the brute-force fp64 oracle on a query subset."""
    from mvkpconv_b200 import synthetic
    sub = lambda p, dl: mvk.grid_subsampling(p, sampleDl=dl)
    sphere = synthetic.make_spheres(1, sub, seed=3)[0]
    world = sphere + np.array([3.0, 2.5, 1.0], np.float32)
    cam, depths, poses = synthetic.make_views(world, n_views=3)
    xyz32, mask, xyz64 = mvk.unproject_views(cam, depths, poses)
    q = world[::40]
    knn = mvk.knn_pixels(xyz64, mask, q, k=3)
    xyz_list = [xyz64.reshape(3, -1, 3)[v] for v in range(3)]
    ref = modules.knn_pixels(xyz_list, [mask[v] for v in range(3)], q, k=3)
    assert np.array_equal(knn, ref)
    assert mask.reshape(-1)[knn.reshape(-1)].all()  # only valid pixels are ever selected


def test_group_points_vs_torch_gather(mvk):
    """Same oracle as the reference's own test (mvpnet/ops/tests/test_group_points.py:6-52)."""
    torch.manual_seed(0)
    for b, c, n1, n2, k in [(2, 64, 128, 16, 4), (1, 64, 5 * 19200, 2000, 3), (1, 3, 1000, 77, 3)]:
        points = torch.randn(b, c, n1, device="cuda", requires_grad=True)
        index = torch.randint(0, n1, (b, n2, k), device="cuda")
        out = mvk.group_points(points, index)
        ref_in = points.detach().clone().requires_grad_(True)
        ref = modules.group_points(ref_in, index)
        assert torch.equal(out, ref)
        go = torch.randn_like(out)
        out.backward(go)
        ref.backward(go)
        assert torch.allclose(points.grad, ref_in.grad, atol=1e-5)


@pytest.mark.parametrize("case", ["sum64", "max16"])
def test_feature_aggregation_vs_reference_golden(mvk, case):
    c = load_golden("feature_aggregation")[case]
    cin = c["feature"].shape[1]
    fa = mvk.FeatureAggregation(cin, mlp_channels=(64, 64, 64), reduction=str(c["reduction"])).cuda()
    sd = {k[3:]: torch.from_numpy(v) for k, v in c.items() if k.startswith("sd.")}
    fa.load_state_dict(sd)  # the reference module's own state dict
    src, tgt, feat = (torch.from_numpy(c[k]).cuda() for k in ("src_xyz", "tgt_xyz", "feature"))
    fa.eval()
    out = fa(src, tgt, feat)
    assert out.shape == c["out_eval"].shape
    assert out.requires_grad  # parameters are trainable: the differentiable pipeline ran (like the reference)
    assert rel_err(out.detach().cpu().numpy(), c["out_eval"]) < 1e-4
    with torch.no_grad():      # inference pipeline (fused, no autograd bookkeeping)
        assert rel_err(fa(src, tgt, feat).cpu().numpy(), c["out_eval"]) < 1e-4
    fa.train()
    with torch.no_grad():
        out = fa(src, tgt, feat)
    assert rel_err(out.cpu().numpy(), c["out_train"]) < 1e-4
    # running statistics moved like nn.BatchNorm2d(momentum=0.1)
    for i in range(3):
        assert rel_err(fa.mlp[i].bn.running_mean.cpu().numpy(), c[f"sd_after.mlp.{i}.bn.running_mean"]) < 1e-4
        assert rel_err(fa.mlp[i].bn.running_var.cpu().numpy(), c[f"sd_after.mlp.{i}.bn.running_var"]) < 1e-4


def test_feature_aggregation_fused_gather_equals_group_points_path(mvk):
    torch.manual_seed(1)
    npix, np_, k, c = 3 * 19200, 5000, 3, 64
    feat2d = torch.randn(c, npix, device="cuda")
    xyz = torch.randn(npix, 3, device="cuda")
    knn = torch.randint(0, npix, (np_, k), device="cuda")
    tgt = torch.randn(np_, 3, device="cuda")
    fa = mvk.FeatureAggregation(c).cuda().eval()
    fused = fa.forward_from_maps(feat2d, xyz, knn, tgt)
    f = mvk.group_points(feat2d[None], knn[None])
    sx = mvk.group_points(xyz.t().contiguous()[None], knn[None])
    ref = fa(sx, tgt.t().contiguous()[None], f)[0]
    assert torch.allclose(fused, ref, rtol=1e-5, atol=1e-5)
    # channels-last (pixel-major) feature maps are consumed through their strides
    fused2 = fa.forward_from_maps(feat2d.t().contiguous().t(), xyz, knn, tgt)
    assert torch.allclose(fused2, fused, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("contraction", ["fp32", "bf16x3"])
def test_feature_aggregation_training_gradients_vs_oracle(mvk, contraction):
    """Late fusion trains FeatureAggregation (architectures_sphere_late_fusion.py:301-304): forward in
    train mode + gradients of every parameter and of the gathered features against autograd through
    the torch restatement (which the reference goldens pin)."""
    torch.manual_seed(4)
    b, c, np_, k = 1, 64, 3000, 3
    fa = mvk.FeatureAggregation(c).cuda().train()
    fa.contraction = contraction
    with torch.no_grad():
        for l in fa.mlp:
            l.bn.weight.uniform_(0.5, 1.5)
            l.bn.bias.uniform_(-0.3, 0.3)
    src = torch.randn(b, 3, np_, k)
    tgt = torch.randn(b, 3, np_)
    feat = torch.randn(b, c, np_, k)
    go = torch.randn(b, 64, np_)
    # oracle in fp64 (the checker) and in fp32 (how far plain fp32 arithmetic itself sits from the fp64 result:
    # the floor any fp32 implementation is entitled to, ReLU mask flips of near-zero pre-activations included)
    def oracle(dt):
        fo = feat.detach().clone().to(dt).requires_grad_(True)
        ws = [l.conv.weight.detach().cpu().to(dt).requires_grad_(True) for l in fa.mlp]
        gs = [l.bn.weight.detach().cpu().to(dt).requires_grad_(True) for l in fa.mlp]
        bs = [l.bn.bias.detach().cpu().to(dt).requires_grad_(True) for l in fa.mlp]
        ref = modules.feature_aggregation_forward(src.to(dt), tgt.to(dt), fo, ws, gs, bs, None, None, True, "sum")
        ref.backward(go.to(dt))
        named = {"out": ref.detach(), "grad_feature": fo.grad}
        for i in range(len(ws)):
            named[f"conv{i}.weight.grad"], named[f"bn{i}.weight.grad"], named[f"bn{i}.bias.grad"] = ws[i].grad, gs[i].grad, bs[i].grad
        return named
    ref64, ref32 = oracle(torch.float64), oracle(torch.float32)
    # product
    fg = feat.cuda().requires_grad_(True)
    out = fa(src.cuda(), tgt.cuda(), fg)
    out.backward(go.cuda())
    got = {"out": out.detach(), "grad_feature": fg.grad}
    for i, l in enumerate(fa.mlp):
        got[f"conv{i}.weight.grad"], got[f"bn{i}.weight.grad"], got[f"bn{i}.bias.grad"] = l.conv.weight.grad, l.bn.weight.grad, l.bn.bias.grad
        assert int(l.bn.num_batches_tracked) == 1
    strict = contraction == "fp32"
    worst = 0.0
    for name, r64 in ref64.items():
        e_gpu = rel_err(got[name].detach().cpu().numpy().reshape(r64.shape), r64.numpy())
        e_f32 = rel_err(ref32[name].numpy(), r64.numpy())
        print(f"FeatureAggregation[{contraction}] {name:20s} gpu vs fp64 {e_gpu:.2e}   host fp32 vs fp64 {e_f32:.2e}")
        worst = max(worst, e_gpu)
        # north_star: 1e-4 relative.  fp32 arithmetic itself is allowed its own distance from the fp64 result
        # (a flipped ReLU mask moves isolated gradient entries by O(1)); bf16x3 is stated separately.
        if strict:
            assert e_gpu < max(1e-4, 4 * e_f32), (name, e_gpu, e_f32)
        elif name == "out":
            assert e_gpu < 1e-4, (name, e_gpu)
        elif name == "grad_feature":
            # bf16x3, stated separately: the 1e-5 forward rounding flips the ReLU mask of the few activations whose
            # pre-activation is within rounding of zero; one flip moves isolated gradient entries by O(1)
            ga, gb = got[name].cpu().numpy().astype(np.float64), r64.numpy()
            assert np.linalg.norm(ga - gb) / np.linalg.norm(gb) < 3e-2
            assert (np.abs(ga - gb) > 2e-4 * np.abs(gb).max()).mean() < 2e-3
        else:
            # ... and a parameter gradient (a sum of ~N random-sign terms) by ~1/sqrt(N) of its size
            assert e_gpu < 5e-2, (name, e_gpu)
    print(f"FeatureAggregation[{contraction}] worst tensor error vs fp64 oracle: {worst:.2e}")
    # the fused map entry point follows the same path when the module is being trained
    torch.manual_seed(5)
    npix = 3 * 19200
    feat2d = torch.randn(c, npix, device="cuda")
    xyz = torch.randn(npix, 3, device="cuda")
    knn = torch.randint(0, npix, (np_, k), device="cuda")
    tgtp = torch.randn(np_, 3, device="cuda")
    fa.zero_grad()
    o2 = fa.forward_from_maps(feat2d, xyz, knn, tgtp)
    assert o2.requires_grad
    o2.sum().backward()
    assert fa.mlp[0].conv.weight.grad is not None and torch.isfinite(fa.mlp[0].conv.weight.grad).all()
