"""CPU timing legs of bench.py.  TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): only
bench.py's `cpu_baseline` / `--impl reference` legs run anything in here.

    radius neighbours   the UNMODIFIED reference C++ (oracle/_ref/libref.so: batch_nanoflann_neighbors,
                        neighbors.cpp:211-332, nanoflann v1.3.0) -- (a) one thread per call, the way the
                        reference runs it; (b) P worker processes (P = host cores), one batch element each, the
                        way the reference's DataLoader workers (`input_threads`) run it.  Falls back to the
                        plain-C restatement (kind "port") when libref.so is absent.
    KPConv layer        the oracle's torch-CPU restatement of KPConv.forward (blocks.py:277-374), all host threads.
    lifting             numpy depth2xyz + pose, sklearn ball-tree 3-NN (ScanNet_sphere_color.py:409-452) and the
                        torch-CPU FeatureAggregation restatement (mvpnet_3d.py:40-64).

The P-process leg forks BEFORE anything else is imported in the workers (numpy + ctypes only: no torch, no
CUDA), via `python -m oracle.cpu_bench neighbors <npz>`; bench.py calls it in a subprocess so that the fork
never happens inside a process that holds a CUDA context.
"""
import json
import os
import sys
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
if os.path.dirname(_HERE) not in sys.path:
    sys.path.insert(0, os.path.dirname(_HERE))


def _nbr_fn():
    from oracle import geom
    if geom.have_ref():
        return geom.ref_batch_neighbors, "reference"
    return geom.batch_neighbors, "port"


def _worker(job):
    pts, radius, reps = job
    fn, _ = _nbr_fn()
    ln = np.array([len(pts)], np.int32)
    w = 0
    for _ in range(reps):
        w = fn(pts, pts, ln, ln, radius).shape[1]
    return len(pts) * reps, w


def neighbors_cpu(points, lengths, radius, procs=None, best_of=3):
    """queries/s of the conv-neighbour call of level 0 (queries = supports = the stacked batch)."""
    fn, kind = _nbr_fn()
    lengths = np.asarray(lengths, np.int32)
    nq = int(len(points))
    t_best, width = 1e30, 0
    for _ in range(best_of):
        t0 = time.perf_counter()
        width = fn(points, points, lengths, lengths, radius).shape[1]
        t_best = min(t_best, time.perf_counter() - t0)
    out = {"kind": kind, "queries": nq, "radius": radius, "row_width": int(width),
           "single_thread": {"queries_per_s": round(nq / t_best, 1), "ms": round(1e3 * t_best, 2), "threads": 1}}
    procs = procs or len(os.sched_getaffinity(0))
    import multiprocessing as mp
    starts = np.concatenate([[0], np.cumsum(lengths)])
    elems = [np.ascontiguousarray(points[starts[i]:starts[i + 1]]) for i in range(len(lengths))]
    reps = 2
    jobs = [(elems[i % len(elems)], radius, reps) for i in range(procs)]  # every worker owns one batch element
    with mp.get_context("fork").Pool(procs) as pool:
        pool.map(_worker, [(e[:2000], radius, 1) for e, _, _ in jobs])  # warm the workers (library load)
        t0 = time.perf_counter()
        done = pool.map(_worker, jobs)
        dt = time.perf_counter() - t0
    total = sum(d[0] for d in done)
    out["worker_processes"] = {"queries_per_s": round(total / dt, 1), "ms": round(1e3 * dt, 2), "processes": procs,
                               "queries": int(total),
                               "what": f"{procs} forked workers, each running the reference call on one sphere x{reps}"}
    return out


def kpconv_layer_cpu(q_pts, inds, x, kernel_points, weights, extent, iters=3):
    """points/s of one rigid KPConv layer forward and forward+backward on torch CPU (all host threads)."""
    import torch
    from oracle import modules
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    q = torch.from_numpy(q_pts)
    ii = torch.from_numpy(inds).long()
    kp, w = torch.from_numpy(kernel_points), torch.from_numpy(weights)
    go = torch.ones(len(q_pts), weights.shape[2])
    best_f = best_fb = 1e30
    for _ in range(iters):
        xx = torch.from_numpy(x).requires_grad_(True)
        ww = w.clone().requires_grad_(True)
        t0 = time.perf_counter()
        out = modules.kpconv_forward(q, q, ii, xx, kp, ww, extent)
        t1 = time.perf_counter()
        out.backward(go)
        t2 = time.perf_counter()
        best_f, best_fb = min(best_f, t1 - t0), min(best_fb, t2 - t0)
    n = len(q_pts)
    return {"kind": "port", "cores": cores, "points": n,
            "fwd_points_per_s": round(n / best_f, 1), "fwd_ms": round(1e3 * best_f, 1),
            "fwd_bwd_points_per_s": round(n / best_fb, 1), "fwd_bwd_ms": round(1e3 * best_fb, 1),
            "what": "oracle.modules.kpconv_forward (torch CPU restatement of blocks.py:277-374), best of %d" % iters}


def lifting_cpu(cam, depths, poses, feature_2d, sphere_pts, conv_weights, iters=2):
    """Lifting of ONE sphere on the CPU the way the reference does it: numpy unprojection, sklearn ball-tree
    3-NN on fp64, torch-CPU FeatureAggregation forward + backward (train mode).
    feature_2d: (nv, C, h, w) float32; conv_weights: list of (Cout, Cin) float32 arrays."""
    import torch
    from sklearn.neighbors import NearestNeighbors
    from oracle import modules
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    nv, h, w = depths.shape
    c = feature_2d.shape[1]
    best = {"unproject": 1e30, "knn": 1e30, "fa_fwd_bwd": 1e30}
    ws = [torch.from_numpy(np.ascontiguousarray(x)).requires_grad_(True) for x in conv_weights]
    bn_w = [torch.ones(x.shape[0], requires_grad=True) for x in conv_weights]
    bn_b = [torch.zeros(x.shape[0], requires_grad=True) for x in conv_weights]
    for _ in range(iters):
        t0 = time.perf_counter()
        xyz, mask = [], []
        for v in range(nv):
            p, m = modules.unproject_view(cam, depths[v], poses[v])
            xyz.append(p)
            mask.append(m)
        xyz, mask = np.concatenate(xyz, 0), np.concatenate(mask, 0)
        t1 = time.perf_counter()
        nbrs = NearestNeighbors(n_neighbors=3, algorithm='ball_tree').fit(xyz[mask])
        _, idx = nbrs.kneighbors(sphere_pts.astype(np.float64))
        flat = np.nonzero(mask)[0][idx]
        t2 = time.perf_counter()
        f2 = torch.from_numpy(np.ascontiguousarray(feature_2d.transpose(1, 0, 2, 3).reshape(1, c, -1))).requires_grad_(True)
        ii = torch.from_numpy(flat).long().unsqueeze(0)
        # group_points as plain indexing (the expand + gather of the reference's test oracle would materialise a
        # (C, np, nv*h*w) gradient in backward)
        src = torch.from_numpy(np.ascontiguousarray(xyz.T.astype(np.float32)))[:, ii[0]].unsqueeze(0)
        gf = f2[:, :, ii[0]]
        tgt = torch.from_numpy(np.ascontiguousarray(sphere_pts.T.astype(np.float32))).unsqueeze(0)
        out = modules.feature_aggregation_forward(src, tgt, gf, ws, bn_w, bn_b, None, None, training=True)
        out.sum().backward()
        t3 = time.perf_counter()
        best["unproject"] = min(best["unproject"], t1 - t0)
        best["knn"] = min(best["knn"], t2 - t1)
        best["fa_fwd_bwd"] = min(best["fa_fwd_bwd"], t3 - t2)
    total = sum(best.values())
    n = len(sphere_pts)
    return {"kind": "port", "cores": cores, "points": n, "views": int(nv), "points_per_s": round(n / total, 1),
            "ms": {k: round(1e3 * v, 1) for k, v in best.items()},
            "what": "numpy depth2xyz + pose, sklearn ball_tree 3-NN (fp64), torch-CPU FeatureAggregation fwd+bwd"}


def main():
    what, path = sys.argv[1], sys.argv[2]
    z = np.load(path)
    if what == "neighbors":
        print(json.dumps(neighbors_cpu(z["points"], z["lengths"], float(z["radius"]))))
    else:
        raise SystemExit("unknown leg " + what)


if __name__ == "__main__":
    main()
