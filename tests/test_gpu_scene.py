"""Whole-scene inference sweep (BASELINE configs[4]) on the GPU: scene.SceneSweep with the product operators against
the same sweep driven by the CPU oracle operators, and the shard-invariance of the vote table."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import geom, modules
from test_distributed_cpu import SMALL_ARCH, _NumpyPyramidOps, _scene

pytestmark = pytest.mark.gpu


def test_scene_sweep_vs_cpu_oracle_and_shards(mvk):
    from mvkpconv_b200 import harness, pyramid, scene
    cfg = pyramid.baseline_config(architecture=list(SMALL_ARCH), first_subsampling_dl=0.05, first_features_dim=16,
                                  num_classes=5, in_features_dim=2, in_radius=0.6)
    pts = _scene()
    centers = scene.sphere_centers(pts, cfg.in_radius, spacing=0.5)
    gops = SimpleNamespace(batch_neighbors=geom.batch_neighbors,
                           batch_grid_subsampling=lambda p, l, sampleDl=0.1, random_grid_orient=True:
                           geom.grid_subsample_batch(p, l, sampleDl=sampleDl))
    mops = SimpleNamespace(KPConv=modules.KPConvOracle, max_pool=modules.max_pool, closest_pool=modules.closest_pool)
    np.random.seed(0)
    torch.manual_seed(0)
    net_c = harness.KPFCNN(cfg, ops=mops)
    sd = {k: v.clone() for k, v in net_c.state_dict().items()}
    ref, ref_counts = scene.SceneSweep(net_c, cfg, spheres_per_batch=3, ops=_NumpyPyramidOps(gops)).run(pts, centers)

    np.random.seed(0)
    torch.manual_seed(0)
    net_g = harness.KPFCNN(cfg).cuda()
    net_g.load_state_dict(sd)
    for m in net_g.modules():
        if hasattr(m, "contraction"):
            m.contraction = "fp32"
    sweep = scene.SceneSweep(net_g, cfg, spheres_per_batch=3)
    got, counts = sweep.run(pts, centers)
    assert np.array_equal(counts.cpu().numpy(), ref_counts.numpy())
    err = float((got.cpu() - ref).abs().max())
    print("scene sweep: max |prob - oracle| =", err, "spheres", sweep.stats.spheres, "points", sweep.stats.points)
    assert err < 1e-4
    # shard invariance on one device: the two ranks' partial tables, summed, equal the unsharded table
    parts = []
    for rank in range(2):
        s = scene.SceneSweep(net_g, cfg, spheres_per_batch=3)
        v = torch.zeros((len(pts), net_g.C), device="cuda")
        c = torch.zeros(len(pts), dtype=torch.int32, device="cuda")
        net_g.eval()
        scene_t, cen = torch.from_numpy(pts).cuda(), torch.from_numpy(centers).cuda()
        mine = scene.shard(len(cen), rank, 2)
        for i in range(0, len(mine), 3):
            s.run_batch(scene_t, cen[mine[i:i + 3]], v, c)
        parts.append((v, c))
    v = parts[0][0] + parts[1][0]
    c = parts[0][1] + parts[1][1]
    assert torch.equal(c, counts)
    merged = v / c.clamp_min(1).unsqueeze(1)
    # not bit-equal: the stacked batch decides the neighbour-matrix width, and the width picks the stage-A kernel
    # variant (approximate-sqrt fast paths for H <= 64, IEEE generic kernel above): ~1e-6 relative on the influences
    assert float((merged - got).abs().max()) < 2e-4
    # the reference's exponential blend (tester.py:199) on a single rank
    sm, _ = scene.SceneSweep(net_g, cfg, spheres_per_batch=3, vote="smooth").run(pts, centers)
    assert float(sm.max()) <= 1.0 and float(sm.sum(1).max()) <= 1.0 + 1e-5
