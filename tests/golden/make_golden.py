"""Generate the golden vectors in tests/golden/ by IMPORTING AND RUNNING the reference's own
Python from /root/reference (read-only).  Run in the authoring container only:

    python tests/golden/make_golden.py

The reference tree does not exist on the GPU box, so the outputs (small .npz files) are committed
and every test reads those.  Nothing here is copied from the reference: modules are imported
(KPConv, FeatureAggregation) or, for a module whose import needs absent packages
(datasets/ScanNet_sphere_color.py needs open3d etc.), the single function `depth2xyz` is compiled
from the reference file's own AST in memory.

Versions that produced the committed files are stored inside each .npz under key "_versions".
"""
import ast
import zlib
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(OUT))
sys.path.insert(0, ROOT)


def _versions():
    import sklearn
    return np.array(f"numpy {np.__version__}; torch {torch.__version__}; sklearn {sklearn.__version__}; "
                    f"python {sys.version.split()[0]}")


def import_reference_kpconv():
    """models/blocks.py needs matplotlib (absent) only for plots, and CWD=KPConv-PyTorch for the
    kernel disposition PLY (kernels/kernel_points.py:411)."""
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    os.chdir(os.path.join(REF, "KPConv-PyTorch"))
    sys.path.insert(0, os.path.join(REF, "KPConv-PyTorch"))
    import models.blocks as blocks
    return blocks


def import_reference_feature_aggregation():
    sys.path.insert(0, REF)
    from mvpnet.FeatureAggregation_dummy_test import FeatureAggregation
    return FeatureAggregation


def reference_function(path, name, env):
    """Compile one top-level function of a reference file without importing the file."""
    src = open(path).read()
    tree = ast.parse(src)
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name][0]
    mod = ast.Module(body=[fn], type_ignores=[])
    exec(compile(mod, path, "exec"), env)
    return env[name]


def synthetic_cloud(rng, n, extent=1.0):
    """A bumpy surface patch + clutter, roughly uniform at the scale of the radii used below."""
    xy = rng.uniform(-extent, extent, (n, 2))
    z = 0.15 * np.sin(3 * xy[:, 0]) * np.cos(2 * xy[:, 1]) + rng.normal(0, 0.01, n)
    pts = np.concatenate([xy, z[:, None]], 1)
    k = n // 4
    pts[:k] = rng.uniform(-extent, extent, (k, 3)) * [1, 1, 0.4]
    return pts.astype(np.float32)


def make_kpconv(blocks):
    from oracle import geom
    cases = {}
    rng = np.random.default_rng(1234)
    specs = [
        # name, Nq-source, Cin, Cout, radius, influence, aggregation, strided
        ("rigid_16_32", 700, 16, 32, 0.25, "linear", "sum", False),
        ("rigid_64_128", 400, 64, 128, 0.3, "linear", "sum", False),
        ("rigid_2_64", 500, 2, 64, 0.25, "linear", "sum", False),
        ("strided_32_32", 800, 32, 32, 0.25, "linear", "sum", True),
        ("gaussian_8_16", 300, 8, 16, 0.3, "gaussian", "sum", False),
        ("constant_8_16", 300, 8, 16, 0.3, "constant", "sum", False),
        ("closest_8_16", 300, 8, 16, 0.3, "linear", "closest", False),
    ]
    for name, n, cin, cout, radius, infl, aggr, strided in specs:
        s_pts = synthetic_cloud(rng, n)
        lens = np.array([n // 3, n - n // 3], np.int32)
        if strided:
            q_pts, q_lens = geom.ref_grid_subsample_batch(s_pts, lens, sampleDl=radius / 2.5 * 2)
        else:
            q_pts, q_lens = s_pts, lens
        inds = geom.ref_batch_neighbors(q_pts, s_pts, q_lens, lens, radius).astype(np.int64)
        # crop like big_neighborhood_filter (common.py:411-421) to exercise cropped rows
        inds = inds[:, :max(4, int(inds.shape[1] * 0.8))]
        np.random.seed(7)
        torch.manual_seed(7)
        extent = radius * 1.2 / 2.5
        m = blocks.KPConv(15, 3, cin, cout, extent, radius, KP_influence=infl, aggregation_mode=aggr)
        x = torch.randn(len(s_pts), cin)
        x.requires_grad_(True)
        out = m(torch.from_numpy(q_pts), torch.from_numpy(s_pts), torch.from_numpy(inds), x)
        g = torch.randn_like(out)
        out.backward(g)
        cases[name] = dict(q_pts=q_pts, s_pts=s_pts, inds=inds, x=x.detach().numpy(),
                           kernel_points=m.kernel_points.detach().numpy(),
                           weights=m.weights.detach().numpy(), KP_extent=np.float32(extent),
                           radius=np.float32(radius), influence=np.array(infl),
                           aggregation=np.array(aggr), out=out.detach().numpy(),
                           grad_out=g.numpy(), grad_x=x.grad.numpy(),
                           grad_w=m.weights.grad.numpy())
    flat = {f"{c}/{k}": v for c, d in cases.items() for k, v in d.items()}
    np.savez_compressed(os.path.join(OUT, "kpconv.npz"), _versions=_versions(), **flat)
    print("kpconv.npz", {c: d["out"].shape for c, d in cases.items()})

    # pools (blocks.py:79-110)
    x = torch.randn(700, 24)
    c = cases["strided_32_32"]
    pool_inds = torch.from_numpy(c["inds"])
    xs = torch.randn(len(c["s_pts"]), 24)
    np.savez_compressed(os.path.join(OUT, "pools.npz"), _versions=_versions(), x=xs.numpy(),
                        inds=c["inds"], max_pool=blocks.max_pool(xs, pool_inds).numpy(),
                        closest_pool=blocks.closest_pool(xs, pool_inds).numpy())

    # the kernel disposition fixture itself (kernels/dispositions/k_015_center_3D.ply, 15x3 f64)
    from utils.ply import read_ply
    d = read_ply(os.path.join(REF, "KPConv-PyTorch/kernels/dispositions/k_015_center_3D.ply"))
    kp = np.vstack((d["x"], d["y"], d["z"])).T
    # and what load_kernels (kernel_points.py:409-490) makes of it for a seeded np.random
    import kernels.kernel_points as kpm
    np.random.seed(3)
    loaded = kpm.load_kernels(0.1, 15, dimension=3, fixed="center")
    np.savez_compressed(os.path.join(OUT, "kernel_points.npz"), _versions=_versions(),
                        disposition=kp, loaded_seed3_r01=loaded)


def make_feature_aggregation(FA):
    rng = np.random.default_rng(5)
    res = {}
    for name, npts, k, cin, reduction in [("sum64", 257, 3, 64, "sum"), ("max16", 100, 3, 16, "max")]:
        torch.manual_seed(11)
        m = FA(cin, mlp_channels=(64, 64, 64), reduction=reduction, use_relation=True)
        with torch.no_grad():
            for mod in m.modules():
                if isinstance(mod, torch.nn.BatchNorm2d):
                    mod.weight.uniform_(0.5, 1.5)
                    mod.bias.uniform_(-0.3, 0.3)
                    mod.running_mean.uniform_(-0.2, 0.2)
                    mod.running_var.uniform_(0.5, 1.5)
        src = torch.from_numpy(rng.normal(0, 0.05, (1, 3, npts, k)).astype(np.float32))
        tgt = torch.from_numpy(rng.normal(0, 0.05, (1, 3, npts)).astype(np.float32))
        feat = torch.from_numpy(rng.normal(0, 1, (1, cin, npts, k)).astype(np.float32))
        feat.requires_grad_(True)
        m.eval()
        out_eval = m(src, tgt, feat).detach().numpy()
        sd_eval = {pn: p.detach().clone().numpy() for pn, p in m.state_dict().items()}
        m.train()
        out = m(src, tgt, feat)
        g = torch.randn_like(out)
        out.backward(g)
        d = dict(src_xyz=src.numpy(), tgt_xyz=tgt.numpy(), feature=feat.detach().numpy(),
                 out_eval=out_eval, out_train=out.detach().numpy(), grad_out=g.numpy(),
                 grad_feature=feat.grad.numpy(), reduction=np.array(reduction))
        for pn, p in sd_eval.items():       # state used by out_eval (before the train-mode update)
            d["sd." + pn] = p
        for pn, p in m.state_dict().items():  # state after one train-mode forward (running stats moved)
            d["sd_after." + pn] = p.numpy()
        for pn, p in m.named_parameters():
            d["grad." + pn] = p.grad.numpy()
        res[name] = d
    flat = {f"{c}/{k}": v for c, d in res.items() for k, v in d.items()}
    np.savez_compressed(os.path.join(OUT, "feature_aggregation.npz"), _versions=_versions(), **flat)
    print("feature_aggregation.npz", list(res))


def make_lifting():
    from sklearn.neighbors import NearestNeighbors
    depth2xyz = reference_function(
        os.path.join(REF, "KPConv-PyTorch/datasets/ScanNet_sphere_color.py"), "depth2xyz", {"np": np})
    rng = np.random.default_rng(9)
    h, w, nv = 30, 40, 3
    # ScanNet depth intrinsics scaled to the view size (ScanNet_sphere_color.py:370-372)
    cam = np.eye(4, dtype=np.float32)
    cam[0, 0] = cam[1, 1] = 577.870605
    cam[0, 2], cam[1, 2] = 319.5, 239.5
    cam[0] /= 640 / w
    cam[1] /= 480 / h
    depths, poses, xyzs, masks = [], [], [], []
    for v in range(nv):
        depth = (1.5 + 0.5 * rng.random((h, w))).astype(np.float32)
        depth[rng.random((h, w)) < 0.07] = 0.0
        depth = (np.round(depth * 1000) / 1000.).astype(np.float32)
        a = rng.uniform(0, 2 * np.pi)
        c, s = np.cos(a), np.sin(a)
        pose = np.eye(4, dtype=np.float32)
        pose[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float32) @ \
            np.array([[1, 0, 0], [0, 0, 1], [0, -1, 0]], np.float32)
        pose[:3, 3] = rng.uniform(-0.5, 0.5, 3)
        # --- the reference lines ScanNet_sphere_color.py:413-417 ---
        image_xyz = depth2xyz(cam, depth)
        image_mask = image_xyz[:, 2] > 0
        image_xyz = np.matmul(image_xyz, pose[:3, :3].T) + pose[:3, 3]
        depths.append(depth), poses.append(pose), xyzs.append(image_xyz), masks.append(image_mask)
    # --- :427-452 (no flip) ---
    image_ind_all = np.hstack([np.nonzero(m)[0] + i * h * w for i, m in enumerate(masks)])
    image_xyz_valid = np.concatenate([x[m] for x, m in zip(xyzs, masks)], 0)
    queries = (image_xyz_valid[rng.choice(len(image_xyz_valid), 500)] +
               rng.normal(0, 0.03, (500, 3))).astype(np.float32)
    nbrs = NearestNeighbors(n_neighbors=3, algorithm="ball_tree").fit(image_xyz_valid)
    dist, knn = nbrs.kneighbors(queries)
    knn_pix = image_ind_all[knn].astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "lifting.npz"), _versions=_versions(), cam_matrix=cam,
                        depth=np.stack(depths), pose=np.stack(poses),
                        image_xyz_f64=np.stack(xyzs), image_mask=np.stack(masks),
                        image_xyz_f32=np.stack(xyzs).astype(np.float32), queries=queries,
                        knn_indices=knn_pix, knn_dist=dist)
    print("lifting.npz", knn_pix.shape)


def make_geometry():
    """Reference C++ outputs (oracle/_ref) on seeded clouds: neighbours + subsampling."""
    from oracle import geom
    rng = np.random.default_rng(77)
    pts = synthetic_cloud(rng, 3000)
    lens = np.array([1200, 1000, 800], np.int32)
    feats = rng.normal(size=(3000, 3)).astype(np.float32)
    labels = rng.integers(0, 6, 3000).astype(np.int32)
    sp, sl = geom.ref_grid_subsample_batch(pts, lens, sampleDl=0.12)
    sp2, sl2, sf2, sc2 = geom.ref_grid_subsample_batch(pts, lens, features=feats, labels=labels,
                                                      sampleDl=0.2, max_p=150)
    conv = geom.ref_batch_neighbors(pts, pts, lens, lens, 0.15)
    conv_ord = geom.ref_batch_neighbors(pts, pts, lens, lens, 0.15, ordered=True)
    pool = geom.ref_batch_neighbors(sp, pts, sl, lens, 0.15)
    pool_ord = geom.ref_batch_neighbors(sp, pts, sl, lens, 0.15, ordered=True)
    up = geom.ref_batch_neighbors(pts, sp, lens, sl, 0.3)
    up_ord = geom.ref_batch_neighbors(pts, sp, lens, sl, 0.3, ordered=True)
    info = open(os.path.join(ROOT, "oracle/_ref/BUILD_INFO.txt")).read().strip()
    np.savez_compressed(os.path.join(OUT, "geometry.npz"), _versions=_versions(),
                        toolchain=np.array(info), points=pts, lengths=lens, features=feats,
                        labels=labels, sub_pts=sp, sub_len=sl, sub2_pts=sp2, sub2_len=sl2,
                        sub2_feats=sf2, sub2_labels=sc2, conv=conv, conv_ordered=conv_ord,
                        pool=pool, pool_ordered=pool_ord, up=up, up_ordered=up_ord)
    print("geometry.npz", sp.shape, conv.shape, pool.shape, up.shape)


def make_kpconv_deformable(blocks):
    """Deformable / modulated KPConv of the reference (models/blocks.py:243-374): forward, min_d2,
    deformed_KP and every gradient, with a loss that also pulls on min_d2 like the fitting
    regulariser does (architectures.py:21-54)."""
    from oracle import geom
    rng = np.random.default_rng(4321)
    flat = {}
    specs = [
        # name, n, cin, cout, radius, influence, modulated, strided
        ("deform_32_64", 500, 32, 64, 0.3, "linear", False, False),
        ("deform_mod_16_32", 400, 16, 32, 0.3, "linear", True, False),
        ("deform_strided_64_64", 700, 64, 64, 0.3, "linear", False, True),
        ("deform_gauss_8_16", 300, 8, 16, 0.3, "gaussian", True, False),
    ]
    for name, n, cin, cout, radius, infl, modulated, strided in specs:
        s_pts = synthetic_cloud(rng, n)
        lens = np.array([n // 3, n - n // 3], np.int32)
        if strided:
            q_pts, q_lens = geom.ref_grid_subsample_batch(s_pts, lens, sampleDl=radius / 2.5 * 2)
        else:
            q_pts, q_lens = s_pts, lens
        inds = geom.ref_batch_neighbors(q_pts, s_pts, q_lens, lens, radius).astype(np.int64)
        inds = inds[:, :max(4, int(inds.shape[1] * 0.8))]
        np.random.seed(11)
        torch.manual_seed(11)
        extent = radius * 1.2 / 2.5
        m = blocks.KPConv(15, 3, cin, cout, extent, radius, KP_influence=infl, deformable=True, modulated=modulated)
        with torch.no_grad():  # non-trivial offsets: the reference initialises the bias with zeros
            m.offset_bias.uniform_(-0.3, 0.3)
            m.offset_conv.weights.mul_(0.5)
        x = torch.randn(len(s_pts), cin).requires_grad_(True)
        out = m(torch.from_numpy(q_pts), torch.from_numpy(s_pts), torch.from_numpy(inds), x)
        g = torch.randn_like(out)
        g2 = torch.randn_like(m.min_d2) * 0.1
        loss = (out * g).sum() + (m.min_d2 / extent ** 2 * g2).sum()
        loss.backward()
        c = dict(q_pts=q_pts, s_pts=s_pts, inds=inds, x=x.detach().numpy(), kernel_points=m.kernel_points.detach().numpy(),
                 weights=m.weights.detach().numpy(), offset_weights=m.offset_conv.weights.detach().numpy(),
                 offset_kernel_points=m.offset_conv.kernel_points.detach().numpy(),
                 offset_bias=m.offset_bias.detach().numpy(), KP_extent=np.float32(extent), radius=np.float32(radius),
                 influence=np.array(infl), modulated=np.array(modulated), out=out.detach().numpy(),
                 min_d2=m.min_d2.detach().numpy(), deformed_KP=m.deformed_KP.detach().numpy(), grad_out=g.numpy(),
                 grad_min_d2=g2.numpy(), grad_x=x.grad.numpy(), grad_w=m.weights.grad.numpy(),
                 grad_offset_w=m.offset_conv.weights.grad.numpy(), grad_offset_bias=m.offset_bias.grad.numpy())
        for k, v in c.items():
            flat[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(OUT, "kpconv_deform.npz"), _versions=_versions(), **flat)
    print("kpconv_deform.npz", len(flat))


def make_blocks(blocks):
    """UnaryBlock / BatchNormBlock of the reference (models/blocks.py:430-504): forward, backward,
    running statistics, in train and eval mode, with and without batch norm."""
    flat = {}
    specs = [
        # name, rows, cin, cout, use_bn, no_relu, train
        ("bn_relu_train", 700, 32, 64, True, False, True),
        ("bn_norelu_train", 513, 64, 128, True, True, True),
        ("bn_relu_eval", 300, 128, 32, True, False, False),
        ("bias_relu", 400, 128, 20, False, False, True),
        ("bias_norelu", 257, 48, 24, False, True, True),
    ]
    for name, rows, cin, cout, use_bn, no_relu, train in specs:
        torch.manual_seed(zlib.crc32(name.encode()) % 1000)
        blk = blocks.UnaryBlock(cin, cout, use_bn, 0.02, no_relu=no_relu)
        with torch.no_grad():
            if use_bn:
                blk.batch_norm.batch_norm.weight.uniform_(0.5, 1.5)
                blk.batch_norm.batch_norm.bias.uniform_(-0.5, 0.5)
                blk.batch_norm.batch_norm.running_mean.uniform_(-0.2, 0.2)
                blk.batch_norm.batch_norm.running_var.uniform_(0.5, 1.5)
            else:
                blk.batch_norm.bias.uniform_(-0.5, 0.5)
        blk.train(train)
        x = (torch.randn(rows, cin) * 1.5 + 0.3).requires_grad_(True)
        go = torch.randn(rows, cout)
        c = {"x": x.detach().numpy().copy(), "grad_out": go.numpy(), "weight": blk.mlp.weight.detach().numpy().copy(),
             "use_bn": np.array(use_bn), "no_relu": np.array(no_relu), "train": np.array(train)}
        if use_bn:
            bn = blk.batch_norm.batch_norm
            c.update(gamma=bn.weight.detach().numpy().copy(), beta=bn.bias.detach().numpy().copy(),
                     rm0=bn.running_mean.numpy().copy(), rv0=bn.running_var.numpy().copy())
        else:
            c["bias"] = blk.batch_norm.bias.detach().numpy().copy()
        out = blk(x)
        out.backward(go)
        c.update(out=out.detach().numpy(), grad_x=x.grad.numpy(), grad_w=blk.mlp.weight.grad.numpy())
        if use_bn:
            c.update(grad_gamma=bn.weight.grad.numpy(), grad_beta=bn.bias.grad.numpy(), rm1=bn.running_mean.numpy().copy(),
                     rv1=bn.running_var.numpy().copy(), nbt=np.array(int(bn.num_batches_tracked)))
        else:
            c["grad_bias"] = blk.batch_norm.bias.grad.numpy()
        for k, v in c.items():
            flat[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(OUT, "blocks.npz"), _versions=_versions(), **flat)
    print("blocks.npz", len(flat))


if __name__ == "__main__":
    only = sys.argv[1:]
    if only == ["fa"]:
        make_feature_aggregation(import_reference_feature_aggregation())
        sys.exit(0)
    if only == ["deform"]:
        make_kpconv_deformable(import_reference_kpconv())
        sys.exit(0)
    if only == ["blocks"]:
        make_blocks(import_reference_kpconv())
        sys.exit(0)
    make_geometry()
    make_lifting()
    FA = import_reference_feature_aggregation()
    make_feature_aggregation(FA)
    blocks = import_reference_kpconv()
    make_kpconv(blocks)
    make_blocks(blocks)
    make_kpconv_deformable(blocks)
