#!/usr/bin/env python
"""Per-entry-point timing of the KPConv layers of the bench workload (configs[1] shapes).

    python scripts/kernel_bench.py [--spheres 8] [--reps 5] [--contraction bf16x3] [--levels 0,1,2,3,4]

Builds the same 8-sphere pyramid as bench.py, then runs every distinct KPConv layer shape of the
baseline architecture forward + backward in isolation, with an L2 flush between repetitions, and
prints the median CUDA-event time of every C-ABI call (events on the launching stream).
Development tool: numbers quoted in profiles/ come from here or from bench.py.
"""
import argparse
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spheres", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--contraction", default="bf16x3")
    ap.add_argument("--levels", default="0,1,2,3,4")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    import mvkpconv_b200 as mvk
    from mvkpconv_b200 import _lib, pyramid, synthetic

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    sub = lambda p, dl: mvk.grid_subsampling(p, sampleDl=dl)
    spheres = synthetic.make_spheres(args.spheres, sub, seed=0, in_radius=2.0, first_dl=0.04)
    pts_h, lens_h = synthetic.stack(spheres)
    cfg = pyramid.baseline_config(in_radius=2.0, first_subsampling_dl=0.04)
    pts, lens = torch.from_numpy(pts_h).to(dev), torch.from_numpy(lens_h).to(dev)
    cfg.neighborhood_limits = pyramid.calibrate_neighborhood_limits(pts, lens, cfg)
    np.random.seed(1)
    pyr = pyramid.build_pyramid(pts, lens, cfg)
    print("points per level:", [int(p.shape[0]) for p in pyr.points], "limits:", cfg.neighborhood_limits, flush=True)

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # (level, cin, cout, strided)
    layers = [(0, 2, 64, False), (0, 32, 32, False), (0, 32, 32, True),
              (1, 64, 64, False), (1, 64, 64, True), (2, 128, 128, False), (2, 128, 128, True),
              (3, 256, 256, False), (3, 256, 256, True), (4, 512, 512, False)]
    levels = set(int(v) for v in args.levels.split(","))
    results = []
    for (lvl, cin, cout, strided) in layers:
        if lvl not in levels:
            continue
        r = cfg.first_subsampling_dl * cfg.conv_radius * 2 ** lvl
        s = pyr.points[lvl]
        if strided:
            q, inds = pyr.points[lvl + 1], pyr.pools[lvl]
        else:
            q, inds = s, pyr.neighbors[lvl]
        np.random.seed(0)
        torch.manual_seed(0)
        conv = mvk.KPConv(15, 3, cin, cout, r * cfg.KP_extent / cfg.conv_radius, r, contraction=args.contraction).to(dev)
        x = torch.randn(s.shape[0], cin, device=dev, requires_grad=(cin > 4))
        go = torch.randn(q.shape[0], cout, device=dev)
        per = {}
        for rep in range(args.reps + 1):
            flush_buf.fill_(rep)
            conv.weights.grad = None
            x.grad = None
            with _lib.profile() as rec:
                out = conv(q, s, inds, x)
                out.backward(go)
                torch.cuda.synchronize()
            if rep == 0:
                continue  # warm-up
            seen = {}
            for name, a, e0, e1 in rec:
                idx = seen.get(name, 0)
                seen[name] = idx + 1
                per.setdefault(f"{name}#{idx}", []).append(e0.elapsed_time(e1) * 1e3)
        nq, h = int(q.shape[0]), int(inds.shape[1])
        ld = (15 * cin + 63) // 64 * 64
        bytes_a = nq * (h * (8 + 12 + 4 * cin) + 12 + 4 * ld)
        row = {"level": lvl, "cin": cin, "cout": cout, "strided": strided, "nq": nq, "ns": int(s.shape[0]), "h": h,
               "us": {k: round(statistics.median(v), 1) for k, v in per.items()}}
        tot = sum(row["us"].values())
        row["total_us"] = round(tot, 1)
        for k, v in row["us"].items():
            if k.startswith("mvk_kpconv_weighted"):
                row.setdefault("algo_GBps", {})[k] = round(bytes_a / (v * 1e-6) / 1e9, 1)
        results.append(row)
        print(json.dumps(row), flush=True)
    print("sum over layers (us):", round(sum(r["total_us"] for r in results), 1))
    if args.json:
        json.dump(results, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
