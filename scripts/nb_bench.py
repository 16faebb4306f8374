import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mvkpconv_b200 as mvk
from mvkpconv_b200 import synthetic
sub = lambda p, dl: mvk.grid_subsampling(p, sampleDl=dl)
sp = synthetic.make_spheres(8, sub, seed=0)
pts_h, lens_h = synthetic.stack(sp)
pts, lens = torch.from_numpy(pts_h).cuda(), torch.from_numpy(lens_h).cuda()
sub_p, sub_l = mvk.batch_grid_subsampling(pts, lens, sampleDl=0.08, random_grid_orient=False)
for name, (q, s, ql, sl, r) in {"conv L0": (pts, pts, lens, lens, 0.1), "pool L0": (sub_p, pts, sub_l, lens, 0.1), "up L0": (pts, sub_p, lens, sub_l, 0.2)}.items():
    ts = []
    for rep in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = mvk.batch_neighbors(q, s, ql, sl, r, max_neighbors=43, out_dtype=torch.int64); e1.record(); torch.cuda.synchronize()
        if rep: ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"{name}: {np.median(ts):.1f} us  nq={q.shape[0]} ns={s.shape[0]} checksum={int(out.sum())}")
