"""CPU tests: the oracle against the golden vectors produced by the reference itself
(tests/golden/make_golden.py) and against the live libstdc++ / compiled reference."""
import os

import numpy as np
import pytest
import torch

from conftest import bumpy_cloud, canonical_rows, load_golden
from oracle import geom, modules

RTOL = 1e-4  # north_star: fp32 features and gradients within 1e-4 relative


def rel_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


# ---------------------------------------------------------------------------------------------
def test_stl_order_matches_live_libstdcxx():
    rng = np.random.default_rng(0)
    for n in [0, 1, 2, 12, 13, 14, 28, 29, 30, 59, 60, 127, 128, 1000, 5087, 5088, 30000]:
        keys = rng.choice(1 << 44, size=n, replace=False).astype(np.uint64)
        assert np.array_equal(geom.stl_order(keys), geom.probe_stl_order(keys)), n
    # structured keys as produced by voxel grids (small strides -> bucket collisions)
    keys = (np.arange(4000, dtype=np.uint64) * 59) % np.uint64(100003)
    keys = np.unique(keys)[np.random.default_rng(1).permutation(len(np.unique(keys)))]
    assert np.array_equal(geom.stl_order(keys), geom.probe_stl_order(keys))


def test_bucket_schedule_tables():
    live = geom.probe_schedule(policy=True)
    grown = geom.probe_schedule(policy=False, max_elems=400000)
    assert np.array_equal(live[:len(grown)], grown)
    # the table embedded in the CUDA product (csrc/subsample.cu) and in oracle_geom.c
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = [d for d in os.listdir(root) if d.endswith("_b200") and os.path.isdir(os.path.join(root, d))][0]
    src = open(os.path.join(root, pkg, "csrc", "subsample.cu")).read()
    body = src[src.index("c_sched[28]"):]
    body = body[body.index("{") + 1:body.index("}")]
    table = np.array([int(x) for x in re.findall(r"(\d+)u", body)], dtype=np.uint64)
    assert np.array_equal(table, live[:len(table)])
    csrc = open(os.path.join(root, "oracle", "oracle_geom.c")).read()
    body = csrc[csrc.index("g_sched[64]"):]
    body = body[body.index("{") + 1:body.index("}")]
    table_c = np.array([int(x) for x in re.findall(r"(\d+)ull", body)], dtype=np.uint64)
    assert np.array_equal(table_c, live[:len(table_c)])


def test_geometry_oracle_vs_reference_golden():
    g = load_golden("geometry")
    pts, lens = g["points"], g["lengths"]
    sp, sl = geom.grid_subsample_batch(pts, lens, sampleDl=0.12)
    assert np.array_equal(sp, g["sub_pts"]) and np.array_equal(sl, g["sub_len"])
    r = geom.grid_subsample_batch(pts, lens, features=g["features"], labels=g["labels"], sampleDl=0.2, max_p=150)
    for a, b in zip(r, (g["sub2_pts"], g["sub2_len"], g["sub2_feats"], g["sub2_labels"])):
        assert np.array_equal(a, b)
    for q, s, ql, sl_, rad, key in [(pts, pts, lens, lens, 0.15, "conv"), (sp, pts, sl, lens, 0.15, "pool"),
                                    (pts, sp, lens, sl, 0.3, "up")]:
        mine = geom.batch_neighbors(q, s, ql, sl_, rad)
        assert np.array_equal(mine, g[key + "_ordered"])  # == reference batch_ordered_neighbors, bit exact
        nano = g[key]                                     # nanoflann: equal up to exact-d2 tie order
        assert mine.shape == nano.shape
        assert np.array_equal(canonical_rows(mine, s, q, len(s)), canonical_rows(nano, s, q, len(s)))


@pytest.mark.skipif(not geom.have_ref(), reason="oracle/_ref/libref.so not built")
def test_geometry_oracle_vs_compiled_reference_random():
    rng = np.random.default_rng(42)
    for trial in range(3):
        n = [500, 2500, 6000][trial]
        pts = bumpy_cloud(rng, n)
        lens = np.array([n // 4, n // 2, n - n // 4 - n // 2], np.int32)
        dl = [0.3, 0.12, 0.07][trial]
        a = geom.grid_subsample_batch(pts, lens, sampleDl=dl)
        b = geom.ref_grid_subsample_batch(pts, lens, sampleDl=dl)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        lab = rng.integers(-2, 40, n).astype(np.int32)
        ft = rng.normal(size=(n, 2)).astype(np.float32)
        a = geom.grid_subsample_batch(pts, lens, features=ft, labels=lab, sampleDl=dl * 2)
        b = geom.ref_grid_subsample_batch(pts, lens, features=ft, labels=lab, sampleDl=dl * 2)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        mine = geom.batch_neighbors(a[0], pts, a[1], lens, dl * 2.5)
        ref = geom.ref_batch_neighbors(a[0], pts, a[1], lens, dl * 2.5, ordered=True)
        assert np.array_equal(mine, ref)


# ---------------------------------------------------------------------------------------------
def test_kpconv_oracle_vs_reference_golden():
    g = load_golden("kpconv")
    for name, c in g.items():
        if name.startswith("_"):
            continue
        x = torch.from_numpy(c["x"]).requires_grad_(True)
        w = torch.from_numpy(c["weights"]).requires_grad_(True)
        out = modules.kpconv_forward(torch.from_numpy(c["q_pts"]), torch.from_numpy(c["s_pts"]),
                                     torch.from_numpy(c["inds"]), x, torch.from_numpy(c["kernel_points"]), w,
                                     float(c["KP_extent"]), str(c["influence"]), str(c["aggregation"]))
        out.backward(torch.from_numpy(c["grad_out"]))
        assert rel_err(out.detach().numpy(), c["out"]) < 1e-6, name
        assert rel_err(x.grad.numpy(), c["grad_x"]) < 1e-6, name
        assert rel_err(w.grad.numpy(), c["grad_w"]) < 1e-6, name


def test_pools_oracle_vs_reference_golden():
    g = load_golden("pools")
    x, inds = torch.from_numpy(g["x"]), torch.from_numpy(g["inds"])
    assert np.array_equal(modules.max_pool(x, inds).numpy(), g["max_pool"])
    assert np.array_equal(modules.closest_pool(x, inds).numpy(), g["closest_pool"])


def test_feature_aggregation_oracle_vs_reference_golden():
    g = load_golden("feature_aggregation")
    for name, c in g.items():
        if name.startswith("_"):
            continue
        ws = [torch.from_numpy(c[f"sd.mlp.{i}.conv.weight"]) for i in range(3)]
        bw = [torch.from_numpy(c[f"sd.mlp.{i}.bn.weight"]) for i in range(3)]
        bb = [torch.from_numpy(c[f"sd.mlp.{i}.bn.bias"]) for i in range(3)]
        bm = [torch.from_numpy(c[f"sd.mlp.{i}.bn.running_mean"]) for i in range(3)]
        bv = [torch.from_numpy(c[f"sd.mlp.{i}.bn.running_var"]) for i in range(3)]
        out = modules.feature_aggregation_forward(torch.from_numpy(c["src_xyz"]), torch.from_numpy(c["tgt_xyz"]),
                                                  torch.from_numpy(c["feature"]), ws, bw, bb, bm, bv,
                                                  training=False, reduction=str(c["reduction"]))
        assert rel_err(out.numpy(), c["out_eval"]) < 1e-5, name
        out = modules.feature_aggregation_forward(torch.from_numpy(c["src_xyz"]), torch.from_numpy(c["tgt_xyz"]),
                                                  torch.from_numpy(c["feature"]), ws, bw, bb, None, None,
                                                  training=True, reduction=str(c["reduction"]))
        assert rel_err(out.numpy(), c["out_train"]) < 1e-5, name


def test_lifting_oracle_vs_reference_golden():
    g = load_golden("lifting")
    cam, depth, pose = g["cam_matrix"], g["depth"], g["pose"]
    xyzs, masks = [], []
    for v in range(len(depth)):
        xyz, m = modules.unproject_view(cam, depth[v], pose[v])
        assert np.array_equal(m, g["image_mask"][v])
        assert np.abs(xyz - g["image_xyz_f64"][v]).max() < 1e-12
        xyzs.append(xyz), masks.append(m)
    knn = modules.knn_pixels(xyzs, masks, g["queries"], k=3)
    assert np.array_equal(knn, g["knn_indices"])  # sklearn ball_tree, no exact ties in this fixture


def test_group_points_oracle():
    torch.manual_seed(0)
    p = torch.randn(2, 5, 30)
    idx = torch.randint(0, 30, (2, 7, 3))
    out = modules.group_points(p, idx)
    for b in range(2):
        for n in range(7):
            for k in range(3):
                assert torch.equal(out[b, :, n, k], p[b, :, idx[b, n, k]])


def test_deformable_kpconv_oracle_vs_reference_golden():
    """The torch restatement of the deformable branch against the reference module's own outputs."""
    g = load_golden("kpconv_deform")
    for name, c in g.items():
        if name.startswith("_"):
            continue
        t = lambda k: torch.from_numpy(c[k])
        x = t("x").requires_grad_(True)
        w = t("weights").requires_grad_(True)
        ow = t("offset_weights").requires_grad_(True)
        ob = t("offset_bias").requires_grad_(True)
        ext = float(c["KP_extent"])
        of = modules.kpconv_forward(t("q_pts"), t("s_pts"), t("inds"), x, t("offset_kernel_points"), ow, ext,
                                    str(c["influence"])) + ob
        out, min_d2, dkp = modules.kpconv_deform_forward(t("q_pts"), t("s_pts"), t("inds"), x, t("kernel_points"), w,
                                                         ext, of, bool(c["modulated"]), str(c["influence"]))
        ((out * t("grad_out")).sum() + (min_d2 / ext ** 2 * t("grad_min_d2")).sum()).backward()
        close = lambda a, b: np.abs(a.detach().numpy() - b).max() <= 2e-5 * max(np.abs(b).max(), 1e-30)
        assert close(out, c["out"]) and close(min_d2, c["min_d2"]) and close(dkp, c["deformed_KP"]), name
        assert close(x.grad, c["grad_x"]) and close(w.grad, c["grad_w"]), name
        assert close(ow.grad, c["grad_offset_w"]) and close(ob.grad, c["grad_offset_bias"]), name
