"""Stage A + contraction: fused kernel against the two-kernel sequence on the real level geometry of the bench batch
(8 synthetic spheres, 5-level pyramid).  python scripts/fused_bench.py [n_spheres]"""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mvkpconv_b200 as mvk
from mvkpconv_b200 import kpconv as kpmod, pyramid, synthetic, _lib

n_sph = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda")
sub = lambda p, dl: mvk.grid_subsampling(p, sampleDl=dl)
pts, lens = synthetic.stack(synthetic.make_spheres(n_sph, sub, seed=0))
cfg = pyramid.baseline_config()
p_d, l_d = torch.from_numpy(pts).to(dev), torch.from_numpy(lens).to(dev)
cfg.neighborhood_limits = pyramid.calibrate_neighborhood_limits(p_d, l_d, cfg)
np.random.seed(1)
pyr = pyramid.build_pyramid(p_d, l_d, cfg)
L = _lib.lib()


def timeit(fn, sets, iters=20):
    for s in sets:
        fn(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters


rows = []
for lvl, (cin, cout) in enumerate([(32, 32), (64, 64), (128, 128)]):
    q = pyr.points[lvl]
    inds = pyr.neighbors[lvl]
    n, h = inds.shape
    r = 0.1 * 2 ** lvl
    conv = mvk.KPConv(15, 3, cin, cout, r * 1.2 / 2.5, r).to(dev)
    nsets = max(2, int(3e8 / (n * 15 * cin * 4)) + 1)
    sets = [torch.randn(n, cin, device=dev) for _ in range(min(nsets, 6))]
    res = {"level": lvl, "n": n, "H": h, "cin": cin, "cout": cout}
    for mode in ("fused_infer", "unfused_infer", "fused_train", "unfused_train"):
        kpmod.FUSED_FORWARD = mode.startswith("fused")
        if mode.endswith("infer"):
            def fn(x):
                with torch.no_grad():
                    conv(q, q, inds, x)
        else:
            def fn(x):
                conv(q, q, inds, x)  # weights require grad: the operand is saved
        res[mode + "_us"] = round(timeit(fn, sets), 1)
    kpmod.FUSED_FORWARD = True
    rows.append(res)
    print(json.dumps(res), flush=True)
