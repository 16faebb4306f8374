"""Golden vectors for the three fusion networks, produced by IMPORTING AND RUNNING the reference's own
`KPFCNN_featureAggre` classes (KPConv-PyTorch/models/architectures_sphere.py, architectures_sphere_middle_fusion.py,
architectures_sphere_late_fusion.py) on the CPU.  Authoring container only:

    python tests/golden/make_golden_fusion.py        -> tests/golden/fusion_{early,middle,late}.npz

What has to be stubbed to import / run the reference here (nothing is copied from it):
  * matplotlib (absent; only used for plots by kernels/kernel_points.py) -> empty modules;
  * mvpnet.ops.group_points needs the compiled `group_points_cuda` extension (THC-era CUDA, does not build against
    torch 2.11) -> a module providing `group_points` = the reference's own test oracle (torch.gather,
    mvpnet/ops/tests/test_group_points.py:6-12);
  * torchvision's `resnet34(pretrained=True)` would download ImageNet weights -> the flag is forced to False;
  * `config.path_2D`: a checkpoint with the state dict of a freshly initialised reference UNetResNet34, written to /tmp;
  * `.cuda()` inside the reference forward (architectures_sphere.py:262) -> identity (CPU run).
After construction the 24 M-parameter 2D network is swapped for a 3-layer convolutional stub (the reference forward only
reads `self.net_2d({'image': x})['feature']`), so that the fixture stays small; its weights are part of the fixture.
Middle fusion: the reference's forward averages the two encoders' bottlenecks (`torch.mean`, :129).
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(OUT))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ARCH = ['simple', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_strided', 'resnetb',
        'nearest_upsample', 'unary', 'nearest_upsample', 'unary']
# late fusion: the reference concatenates transform_mlp's 64 channels with the 64 lifted ones and feeds a head sized
# `out_dim` (late_fusion.py:171-172, :301-304): it only runs with first_features_dim = 128; one strided level keeps it small
ARCH_LATE = ['simple', 'resnetb', 'resnetb_deformable_strided', 'resnetb_deformable', 'nearest_upsample', 'unary']
ARCH_DEFORM = ['simple', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_deformable_strided', 'resnetb_deformable',
               'nearest_upsample', 'unary', 'nearest_upsample', 'unary']


class Feature2DStub(torch.nn.Module):
    """Stand-in for the frozen 2D network: {'image': (V, 3, h, w)} -> {'feature': (V, 64, h, w)}."""

    def __init__(self):
        super().__init__()
        self.c1 = torch.nn.Conv2d(3, 16, 3, padding=1)
        self.c2 = torch.nn.Conv2d(16, 64, 3, padding=1)

    def forward(self, d):
        return {'feature': self.c2(torch.relu(self.c1(d['image'])))}


def import_reference(which):
    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.cm"]:
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    gp = types.ModuleType("mvpnet.ops.group_points")

    def group_points(points, index):  # mvpnet/ops/tests/test_group_points.py:6-12
        b, c, n1 = points.shape
        _, n2, k = index.shape
        return points.unsqueeze(2).expand(b, c, n2, n1).gather(3, index.unsqueeze(1).expand(b, c, n2, k))
    gp.group_points = group_points
    sys.modules["mvpnet.ops.group_points"] = gp
    import torchvision.models.resnet as tvr
    real = tvr.resnet34
    tvr.resnet34 = lambda pretrained=False, **kw: real(weights=None)
    os.chdir(os.path.join(REF, "KPConv-PyTorch"))
    for p in (os.path.join(REF, "KPConv-PyTorch"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib
    mod = importlib.import_module({"early": "models.architectures_sphere", "middle": "models.architectures_sphere_middle_fusion",
                                   "late": "models.architectures_sphere_late_fusion"}[which])
    return mod


def make(which):
    from types import SimpleNamespace
    from oracle import geom, modules
    from mvkpconv_b200 import pyramid, synthetic
    from test_gpu_fusion import make_scene
    mod = import_reference(which)
    spheres, cams, depths, poses = make_scene(21, n_spheres=2, nv=2, h=24, w=32, n_pts=900)
    B, nv, h, w = depths.shape
    lens = np.array([len(s) for s in spheres], np.int32)
    world = np.concatenate(spheres, 0)
    centred = np.concatenate([s - s.mean(0, keepdims=True) for s in spheres], 0).astype(np.float32)
    # the reference's config object: only the fields its constructor / blocks read
    cfg = pyramid.baseline_config(architecture=list({"early": ARCH, "middle": ARCH_DEFORM, "late": ARCH_LATE}[which]),
                                  first_subsampling_dl=0.06, first_features_dim=128 if which == "late" else 16, num_classes=6, deform_radius=4.0, in_features_dim=66 if which == "early" else 4,
                                  in_features_dim_3d=4, in_features_dim_2d=65)
    cfg.class_w = []
    cfg.deform_lr_factor = 0.1
    import torchvision  # noqa: F401
    from mvpnet.models.unet_resnet34 import UNetResNet34
    torch.manual_seed(0)
    ck = "/tmp/unet_resnet34_random.pth"
    torch.save({"model": UNetResNet34(20, p=0.5, pretrained=False).state_dict()}, ck)
    cfg.path_2D = ck
    np.random.seed(0)
    torch.manual_seed(0)
    net = mod.KPFCNN_featureAggre(cfg, np.arange(6), [])
    torch.manual_seed(1)
    net.net_2d = Feature2DStub()
    for p in net.net_2d.parameters():
        p.requires_grad = False
    net.train()
    # batch built with the reference C++ (pyramid) and sklearn (kNN), exactly the reference's data path
    gops = SimpleNamespace(batch_neighbors=geom.ref_batch_neighbors if geom.have_ref() else geom.batch_neighbors,
                           batch_grid_subsampling=lambda p, l, sampleDl=0.1, random_grid_orient=True:
                           (geom.ref_grid_subsample_batch if geom.have_ref() else geom.grid_subsample_batch)(p, l, sampleDl=sampleDl))
    pyr = pyramid.build_pyramid(centred, lens, cfg, ops=gops, random_grid_orient=False)
    from sklearn.neighbors import NearestNeighbors
    xyz_all, knn_list = [], []
    for b in range(B):
        xs, ms = [], []
        for v in range(nv):
            x, m = modules.unproject_view(cams[b], depths[b][v], poses[b][v])
            xs.append(x)
            ms.append(m)
        valid = np.concatenate(ms)
        allx = np.concatenate(xs, 0)
        nb = NearestNeighbors(n_neighbors=3, algorithm='ball_tree').fit(allx[valid])
        _, idx = nb.kneighbors(spheres[b])
        knn_list.append(np.nonzero(valid)[0][idx][None].astype(np.int64))
        xyz_all.append(np.stack(xs, 0).reshape(nv, h, w, 3))
    image_xyz = np.stack(xyz_all, 0).astype(np.float32)
    rng = np.random.default_rng(3)
    images = rng.normal(0, 1, (B, nv, 3, h, w)).astype(np.float32)
    labels = rng.integers(0, 6, len(world)).astype(np.int64)
    f3d = (np.concatenate([np.ones((len(world), 1), np.float32), world[:, 2:3]], 1) if which == "early" else
           np.concatenate([np.ones((len(world), 1), np.float32), world], 1)).astype(np.float32)
    as_t = lambda lst, dt: [torch.from_numpy(np.ascontiguousarray(a)).to(dt) for a in lst]
    batch = SimpleNamespace(points=as_t(pyr.points, torch.float32), neighbors=as_t(pyr.neighbors, torch.int64),
                            pools=as_t(pyr.pools, torch.int64), upsamples=as_t(pyr.upsamples, torch.int64),
                            lengths=pyr.lengths, images=torch.from_numpy(images), image_xyz=torch.from_numpy(image_xyz),
                            knn_list=knn_list, feat_aggre_points=torch.from_numpy(world)[None],
                            feature_3d=torch.from_numpy(f3d))
    torch.Tensor.cuda = lambda self, *a, **k: self  # architectures_sphere.py:262 (CPU run)
    out = net(batch, cfg)
    loss = net.loss(out, torch.from_numpy(labels))
    loss.backward()
    sd = {k: v.detach().numpy() for k, v in net.state_dict().items()}
    grads = {k: p.grad.numpy() for k, p in net.named_parameters() if p.grad is not None}
    fix = dict(which=np.array(which), lens=lens, world=world, centred=centred, images=images, image_xyz=image_xyz,
               labels=labels, feature_3d=f3d, knn=np.concatenate([k[0] for k in knn_list], 0), logits=out.detach().numpy(),
               loss=np.float32(float(loss.detach())), cams=cams, depths=depths, poses=poses,
               versions=np.array(f"torch {torch.__version__}; numpy {np.__version__}"))
    for k, v in sd.items():
        fix["sd/" + k] = v
    for k, v in grads.items():
        fix["grad/" + k] = v.astype(np.float32)
    path = os.path.join(OUT, f"fusion_{which}.npz")
    np.savez_compressed(path, **fix)
    print(which, out.shape, float(loss.detach()), len(sd), "tensors,", round(os.path.getsize(path) / 1e6, 2), "MB")


def make_baseline():
    """The KPFCNN of the baseline training script (models/architectures.py:189-352, rigid + deformable blocks) run by
    the reference's own class: forward, its loss incl. the deformable regulariser (p2p_fitting_regularizer), backward."""
    from types import SimpleNamespace
    from oracle import geom
    from mvkpconv_b200 import pyramid
    from test_gpu_network import cloud
    import_reference("early")  # sets up the stubs, sys.path and cwd
    import importlib
    arch = importlib.import_module("models.architectures")
    rng = np.random.default_rng(5)
    pts = np.concatenate([cloud(rng, 1100), cloud(rng, 900)], 0)
    lens = np.array([1100, 900], np.int32)
    cfg = pyramid.baseline_config(architecture=list(ARCH_DEFORM), first_subsampling_dl=0.03, first_features_dim=16, num_classes=6,
                                  in_features_dim=2, deform_radius=4.0)
    cfg.class_w = []
    cfg.deform_lr_factor = 0.1
    np.random.seed(0)
    torch.manual_seed(0)
    net = arch.KPFCNN(cfg, np.arange(6), [])
    net.train()
    gops = SimpleNamespace(batch_neighbors=geom.ref_batch_neighbors if geom.have_ref() else geom.batch_neighbors,
                           batch_grid_subsampling=lambda p, l, sampleDl=0.1, random_grid_orient=True:
                           (geom.ref_grid_subsample_batch if geom.have_ref() else geom.grid_subsample_batch)(p, l, sampleDl=sampleDl))
    pyr = pyramid.build_pyramid(pts, lens, cfg, ops=gops, random_grid_orient=False)
    feats = np.concatenate([np.ones((len(pts), 1), np.float32), pts[:, 2:3]], 1)
    labels = rng.integers(0, 6, len(pts)).astype(np.int64)
    as_t = lambda lst, dt: [torch.from_numpy(np.ascontiguousarray(a)).to(dt) for a in lst]
    batch = SimpleNamespace(points=as_t(pyr.points, torch.float32), neighbors=as_t(pyr.neighbors, torch.int64),
                            pools=as_t(pyr.pools, torch.int64), upsamples=as_t(pyr.upsamples, torch.int64),
                            lengths=pyr.lengths, features=torch.from_numpy(feats))
    out = net(batch, cfg)
    loss = net.loss(out, torch.from_numpy(labels))
    loss.backward()
    fix = dict(lens=lens, points=pts, features=feats, labels=labels, logits=out.detach().numpy(),
               loss=np.float32(float(loss.detach())), output_loss=np.float32(float(net.output_loss.detach())),
               reg_loss=np.float32(float(net.reg_loss.detach())),
               versions=np.array(f"torch {torch.__version__}; numpy {np.__version__}"))
    for k, v in net.state_dict().items():
        fix["sd/" + k] = v.detach().numpy()
    for k, p in net.named_parameters():
        if p.grad is not None:
            fix["grad/" + k] = p.grad.numpy().astype(np.float32)
    path = os.path.join(OUT, "kpfcnn_baseline.npz")
    np.savez_compressed(path, **fix)
    print("baseline", out.shape, float(loss.detach()), float(net.reg_loss.detach()), round(os.path.getsize(path) / 1e6, 2), "MB")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "early"
    make_baseline() if which == "baseline" else make(which)
