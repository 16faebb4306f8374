#!/usr/bin/env python
"""cProfile of the host side of the bench step (development tool): where does the Python time go?"""
import cProfile, io, os, pstats, sys
from types import SimpleNamespace
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import mvkpconv_b200 as mvk
from mvkpconv_b200 import harness, pyramid, synthetic

dev = torch.device("cuda", 0)
sub = lambda p, dl: mvk.grid_subsampling(p, sampleDl=dl)
spheres = synthetic.make_spheres(8, sub, seed=0, in_radius=2.0, first_dl=0.04)
pts_h, lens_h = synthetic.stack(spheres)
cfg = pyramid.baseline_config(in_radius=2.0, first_subsampling_dl=0.04)
pts, lens = torch.from_numpy(pts_h).to(dev), torch.from_numpy(lens_h).to(dev)
cfg.neighborhood_limits = pyramid.calibrate_neighborhood_limits(pts, lens, cfg)
np.random.seed(0); torch.manual_seed(0)
net = harness.KPFCNN(cfg).to(dev)
opt = torch.optim.SGD(net.parameters(), lr=1e-2, momentum=0.98, weight_decay=1e-3, fused=True)
grad_params = [p for p in net.parameters() if p.requires_grad]
feats = torch.from_numpy(bench.host_features(pts_h)).to(dev)
labels = torch.from_numpy(np.random.default_rng(0).integers(0, 20, len(pts_h)).astype(np.int64)).to(dev)

def step(part=None):
    np.random.seed(1)
    pyr = pyramid.build_pyramid(pts, lens, cfg)
    batch = SimpleNamespace(points=pyr.points, neighbors=pyr.neighbors, pools=pyr.pools, upsamples=pyr.upsamples,
                            lengths=pyr.lengths, features=feats, labels=labels)
    out = net(batch)
    loss = net.loss(out, labels)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    torch.nn.utils.clip_grad_value_(grad_params, 100.0)
    opt.step()

for _ in range(3):
    step()
torch.cuda.synchronize()
import time
for name in ("pyramid", "forward", "backward", "opt"):
    pass
# coarse wall-clock split of the host time (no syncs except those inside the pyramid)
t = {"pyramid": 0.0, "forward": 0.0, "backward": 0.0, "opt": 0.0}
for _ in range(5):
    t0 = time.perf_counter(); np.random.seed(1); pyr = pyramid.build_pyramid(pts, lens, cfg)
    batch = SimpleNamespace(points=pyr.points, neighbors=pyr.neighbors, pools=pyr.pools, upsamples=pyr.upsamples,
                            lengths=pyr.lengths, features=feats, labels=labels)
    t1 = time.perf_counter(); out = net(batch); loss = net.loss(out, labels)
    t2 = time.perf_counter(); opt.zero_grad(set_to_none=True); loss.backward()
    t3 = time.perf_counter(); torch.nn.utils.clip_grad_value_(grad_params, 100.0); opt.step()
    t4 = time.perf_counter()
    t["pyramid"] += t1 - t0; t["forward"] += t2 - t1; t["backward"] += t3 - t2; t["opt"] += t4 - t3
    torch.cuda.synchronize()
print({k: round(1e3 * v / 5, 2) for k, v in t.items()}, "ms host time per step (enqueue only)")
# GPU-bound or host-bound?  Inject ~5 ms of extra GPU work after the pyramid: a GPU-bound step grows by
# ~5 ms, a host-bound one absorbs it.
def timed_steps(extra_cycles, n=6):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        np.random.seed(1)
        pyr = pyramid.build_pyramid(pts, lens, cfg)
        if extra_cycles:
            torch.cuda._sleep(extra_cycles)
        batch = SimpleNamespace(points=pyr.points, neighbors=pyr.neighbors, pools=pyr.pools, upsamples=pyr.upsamples,
                                lengths=pyr.lengths, features=feats, labels=labels)
        out = net(batch); loss = net.loss(out, labels); opt.zero_grad(set_to_none=True); loss.backward()
        torch.nn.utils.clip_grad_value_(grad_params, 100.0); opt.step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
base = timed_steps(0)
plus = timed_steps(int(5e-3 * 1.9e9))
print(f"step {base:.2f} ms; with ~5 ms injected GPU sleep {plus:.2f} ms  (delta {plus - base:.2f})")
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
print(s.getvalue()[:9000])
