"""Find the op that breaks stream capture of a deformable training step (debug helper)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mvkpconv_b200 as mvk
from mvkpconv_b200 import harness, pyramid
from test_gpu_network import cloud

rng = np.random.default_rng(7)
dev = torch.device("cuda")
arch = ['simple', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_deformable_strided', 'resnetb_deformable',
        'nearest_upsample', 'unary', 'nearest_upsample', 'unary']
cfg = pyramid.baseline_config(architecture=arch, first_subsampling_dl=0.03, first_features_dim=32, num_classes=6,
                              in_features_dim=2, deform_radius=4.0)
pts = np.concatenate([cloud(rng, 1500), cloud(rng, 1200)], 0)
lens = np.array([1500, 1200], np.int32)
feats = np.concatenate([np.ones((len(pts), 1), np.float32), pts[:, 2:3]], 1)
labels = rng.integers(0, 6, len(pts)).astype(np.int64)
pyr = pyramid.build_pyramid(torch.from_numpy(pts).to(dev), torch.from_numpy(lens).to(dev), cfg, random_grid_orient=False)
net = harness.KPFCNN(cfg).cuda()
opt = torch.optim.SGD(net.parameters(), lr=1e-2, momentum=0.9, fused=True)
st = harness.GraphedTrainStep(net, opt, warm=2)
f, y = torch.from_numpy(feats).to(dev), torch.from_numpy(labels).to(dev)
for i in range(2):
    print("eager", float(st(pyr, f, y)))
with torch.autograd.detect_anomaly(check_nan=False):
    try:
        print("graph", float(st(pyr, f, y)))
        print("graph", float(st(pyr, f, y)))
        print("OK")
    except Exception as e:
        import traceback
        traceback.print_exc()
