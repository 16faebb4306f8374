"""GPU parity of the batched lifting launch set and of the three MV-KPConv fusion networks.

* unproject_views_batched / knn_pixels_batched against the single-sphere operators (themselves pinned on the
  reference goldens: tests/test_gpu_lifting.py) and against the CPU oracle (sklearn-equivalent brute force, fp64):
  bit-exact coordinates and indices, including queries far from every pixel (exhaustive pass), a sphere with
  fewer valid pixels than k, and per-sphere intrinsics.
* FusionKPFCNN (early / middle / late; architectures_sphere*.py) on the GPU against the SAME graph composed from
  the CPU oracle operators (reference-order KPConv, torch unary blocks, the reference's per-sphere group_points
  loop + FeatureAggregation module, torch UNet-ResNet34 on the CPU): logits, loss and parameter gradients with
  the strict fp32 contraction within 1e-4 relative (the 2D network: cuDNN with TF32 disabled, bar 1e-3 on the
  lifted features through 30+ convolution layers).
"""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import geom, modules

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    b = b.detach().double().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def make_scene(seed, n_spheres=3, nv=3, h=60, w=80, n_pts=1500):
    """Seeded spheres cut from a synthetic room + views looking at them (world frame)."""
    from mvkpconv_b200 import synthetic
    rng = np.random.default_rng(seed)
    room = synthetic.make_room(seed=seed, density=1500.0)
    spheres, cams, depths, poses = [], [], [], []
    for i in range(n_spheres):
        c = np.array([rng.uniform(1.5, 4.5), rng.uniform(1.5, 3.5), 1.0], np.float32)
        d2 = ((room - c) ** 2).sum(1)
        pts = room[d2 < 1.2 ** 2]
        pts = pts[rng.permutation(len(pts))[:n_pts + 100 * i]]
        cam, dep, pose = synthetic.make_views(pts, n_views=nv, h=h, w=w, seed=seed + i)
        cam = cam.copy()
        cam[0, 0] *= 1.0 + 0.01 * i  # per-sphere intrinsics
        spheres.append(pts.astype(np.float32))
        cams.append(cam)
        depths.append(dep)
        poses.append(pose)
    return spheres, np.stack(cams), np.stack(depths), np.stack(poses)


def test_batched_lifting_matches_single_sphere_and_oracle(mvk):
    spheres, cams, depths, poses = make_scene(3)
    # element 2: almost no valid depth (fewer than k valid pixels) -> -1 indices like an empty ball tree would fail
    depths[2][:] = 0.0
    depths[2][0, 5, 5] = 1.0
    depths[2][1, 7, 9] = 1.5
    # far queries: move half of sphere 1 three metres away from every pixel
    spheres[1][::2] += np.array([3.0, 0.0, 0.0], np.float32)
    B, nv, h, w = depths.shape
    xyz32, mask, xyz64 = mvk.unproject_views_batched(cams, depths, poses)
    pts = np.concatenate(spheres, 0)
    lens = np.array([len(s) for s in spheres], np.int32)
    knn, far = mvk.knn_pixels_batched(xyz64, xyz32, mask, torch.from_numpy(pts).cuda(), lens, k=3, return_far=True)
    knn_g = mvk.knn_pixels_batched(xyz64, xyz32, mask, torch.from_numpy(pts).cuda(), lens, k=3, global_ids=True)
    knn, far = knn.cpu().numpy(), far.cpu().numpy()
    assert far[1] >= len(spheres[1]) // 2 - 5, far  # the displaced half went through the exhaustive pass
    o = 0
    for b in range(B):
        x32, m, x64 = mvk.unproject_views(cams[b], depths[b], poses[b])
        assert np.array_equal(x32, xyz32[b].cpu().numpy())
        assert np.array_equal(m, mask[b].cpu().numpy())
        assert np.array_equal(x64, xyz64.reshape(B, nv * h * w, 3)[b].cpu().numpy())
        # oracle unprojection (numpy, the reference's own expression)
        for v in range(nv):
            xo, mo = modules.unproject_view(cams[b], depths[b][v], poses[b][v])
            assert np.array_equal(xo.astype(np.float32), x32[v].reshape(-1, 3))
            assert np.array_equal(mo, m[v].reshape(-1))
        n = len(spheres[b])
        got = knn[o:o + n]
        if b == 2:
            assert (got[:, 2] == -1).all() and (got[:, :2] >= 0).all()
        else:
            single = mvk.knn_pixels(x64, m, spheres[b], k=3)
            assert np.array_equal(got, single), b
            ref = modules.knn_pixels([x64.reshape(nv, h * w, 3)[v] for v in range(nv)],
                                     [m[v].reshape(-1) for v in range(nv)], spheres[b][:400], k=3)
            assert np.array_equal(got[:400], ref), b
            assert np.array_equal(knn_g[o:o + n].cpu().numpy(), got + b * nv * h * w)
        o += n


def _small_config(fusion):
    from mvkpconv_b200 import fusion as fu
    arch = ['simple', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_strided', 'resnetb',
            'nearest_upsample', 'unary', 'nearest_upsample', 'unary']
    if fusion != "early":
        arch = ['simple', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_deformable_strided', 'resnetb_deformable',
                'nearest_upsample', 'unary', 'nearest_upsample', 'unary']
    return fu.fusion_config(fusion, architecture=arch, first_subsampling_dl=0.06, first_features_dim=32, num_classes=6,
                            deform_radius=4.0)


@pytest.mark.parametrize("fusion", ["early", "middle", "late"])
def test_fusion_net_vs_cpu_oracle(mvk, fusion):
    from mvkpconv_b200 import fusion as fu, harness, pyramid
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    spheres, cams, depths, poses = make_scene(11, n_spheres=2, nv=2, h=32, w=48, n_pts=1300)
    B, nv, h, w = depths.shape
    lens = np.array([len(s) for s in spheres], np.int32)
    world = np.concatenate(spheres, 0)
    centred = np.concatenate([s - s.mean(0, keepdims=True) for s in spheres], 0).astype(np.float32)
    rng = np.random.default_rng(5)
    labels = rng.integers(0, 6, len(world)).astype(np.int64)
    images = rng.normal(0, 1, (B, nv, 3, h, w)).astype(np.float32)
    f3d = (np.concatenate([np.ones((len(world), 1), np.float32), world[:, 2:3]], 1) if fusion == "early" else
           np.concatenate([np.ones((len(world), 1), np.float32), world], 1)).astype(np.float32)
    cfg = _small_config(fusion)
    dev = torch.device("cuda")

    # ---- lifting indices on the GPU; the CPU graph consumes them as the reference's per-sphere knn_list
    xyz32, mask, knn_g = fu.prepare_lifting(cams, depths, poses, torch.from_numpy(world).to(dev), lens)
    _, _, xyz64 = mvk.unproject_views_batched(cams, depths, poses)
    knn_local = mvk.knn_pixels_batched(xyz64, xyz32, mask, torch.from_numpy(world).to(dev), lens, k=3).cpu().numpy()
    starts = np.concatenate([[0], np.cumsum(lens)])
    knn_list = [knn_local[starts[i]:starts[i + 1]][None] for i in range(B)]

    # ---- CPU oracle graph
    gops = SimpleNamespace(batch_neighbors=geom.batch_neighbors,
                           batch_grid_subsampling=lambda p, l, sampleDl=0.1, random_grid_orient=True:
                           geom.grid_subsample_batch(p, l, sampleDl=sampleDl))
    mops = SimpleNamespace(KPConv=modules.KPConvOracle, max_pool=modules.max_pool, closest_pool=modules.closest_pool,
                           group_points=modules.group_points, FeatureAggregation=modules.FeatureAggregationOracle)
    np.random.seed(0)
    torch.manual_seed(0)
    net2d = fu.UNetResNet34(20, p=0.5, pretrained=False)
    net_c = fu.FusionKPFCNN(cfg, fusion=fusion, net_2d=net2d, ops=mops)
    sd0 = {k: v.clone() for k, v in net_c.state_dict().items()}
    pyr_c = pyramid.build_pyramid(centred, lens, cfg, ops=gops, random_grid_orient=False)
    as_t = lambda lst, dt: [torch.from_numpy(np.ascontiguousarray(a)).to(dt) for a in lst]
    batch_c = SimpleNamespace(points=as_t(pyr_c.points, torch.float32), neighbors=as_t(pyr_c.neighbors, torch.int64),
                              pools=as_t(pyr_c.pools, torch.int64), upsamples=as_t(pyr_c.upsamples, torch.int64),
                              lengths=pyr_c.lengths, images=torch.from_numpy(images), image_xyz=xyz32.cpu(),
                              knn_list=knn_list, feat_aggre_points=torch.from_numpy(world)[None],
                              feature_3d=torch.from_numpy(f3d))
    out_c = net_c(batch_c)
    loss_c = net_c.loss(out_c, torch.from_numpy(labels))
    loss_c.backward()

    # ---- product graph on the GPU, same parameters, strict fp32 contraction
    np.random.seed(0)
    torch.manual_seed(0)
    net_g = fu.FusionKPFCNN(cfg, fusion=fusion, net_2d=fu.UNetResNet34(20, p=0.5, pretrained=False)).to(dev)
    net_g.load_state_dict(sd0, strict=True)
    for m in net_g.modules():
        if hasattr(m, "contraction"):
            m.contraction = "fp32"
    net_g.feat_aggreg.contraction = "fp32"
    pyr_g = pyramid.build_pyramid(torch.from_numpy(centred).to(dev), torch.from_numpy(lens).to(dev), cfg,
                                  random_grid_orient=False)
    batch_g = SimpleNamespace(points=pyr_g.points, neighbors=pyr_g.neighbors, pools=pyr_g.pools, upsamples=pyr_g.upsamples,
                              lengths=pyr_g.lengths, images=torch.from_numpy(images).to(dev), image_xyz=xyz32,
                              knn_global=knn_g, feat_aggre_points=torch.from_numpy(world).to(dev)[None],
                              feature_3d=torch.from_numpy(f3d).to(dev))
    lifted_g, lifted_c = net_g.lift(batch_g), net_c.lift(batch_c)
    e_lift = rel_err(lifted_g, lifted_c)
    out_g = net_g(batch_g)
    loss_g = net_g.loss(out_g, torch.from_numpy(labels).to(dev))
    loss_g.backward()
    e_out = rel_err(out_g, out_c)
    print(f"{fusion}: lifted {e_lift:.2e} logits {e_out:.2e} loss {float(loss_g):.6f} vs {float(loss_c):.6f}")
    assert e_lift < 1e-3
    assert e_out < 1e-3
    assert abs(float(loss_g) - float(loss_c)) < 1e-3 * abs(float(loss_c))
    pc, pg = dict(net_c.named_parameters()), dict(net_g.named_parameters())
    num = den = 0.0
    checked = 0
    for k, p in pc.items():
        if p.grad is None:
            assert pg[k].grad is None or float(pg[k].grad.abs().max()) == 0.0, k
            continue
        assert pg[k].grad is not None, k
        d = (pg[k].grad.detach().double().cpu() - p.grad.double())
        num += float((d * d).sum())
        den += float((p.grad.double() ** 2).sum())
        checked += 1
    assert checked > 20
    assert (num / den) ** 0.5 < 2e-3, (num / den) ** 0.5
    if fusion == "late":  # the lifting module trains in late fusion only
        assert pc["feat_aggreg.mlp.0.conv.weight"].grad is not None
    else:
        assert pc["feat_aggreg.mlp.0.conv.weight"].grad is None
