"""N > 1 host logic on CPU: world_size-2 `gloo` run of the sphere-sharded training step.

The path shards by sphere: every rank builds the pyramid of its own spheres
and runs forward/backward on them; the only collective is the DDP gradient all-reduce.  Here the
same harness graph runs with the CPU oracle operators (the product operators need a B200), two
ranks with different shards, and the test checks that
  * both ranks end the step with identical (all-reduced) gradients,
  * those equal the mean of the two single-rank gradients computed without DDP,
  * a subsequent SGD step keeps the replicas bit-identical.
"""
import os
import socket
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SMALL_ARCH = ['simple', 'resnetb', 'resnetb_strided', 'resnetb', 'nearest_upsample', 'unary']


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _shard(rank, n=600):
    rng = np.random.default_rng(10 + rank)
    clouds = []
    for _ in range(2):  # two spheres per rank
        xy = rng.uniform(-0.5, 0.5, (n, 2))
        z = 0.1 * np.sin(5 * xy[:, :1]) + rng.normal(0, 0.005, (n, 1))
        clouds.append(np.concatenate([xy, z], 1).astype(np.float32))
    pts = np.concatenate(clouds, 0)
    lens = np.array([n, n], np.int32)
    labels = rng.integers(0, 5, len(pts)).astype(np.int64)
    return pts, lens, labels


def _build(rank):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from mvkpconv_b200 import harness, pyramid
    from oracle import geom, modules

    cfg = pyramid.baseline_config(architecture=list(SMALL_ARCH), first_subsampling_dl=0.04, first_features_dim=16,
                                  num_classes=5, in_features_dim=2)
    gops = SimpleNamespace(
        batch_neighbors=geom.batch_neighbors,
        batch_grid_subsampling=lambda p, l, sampleDl=0.1, random_grid_orient=True: geom.grid_subsample_batch(p, l, sampleDl=sampleDl))
    mops = SimpleNamespace(KPConv=modules.KPConvOracle, max_pool=modules.max_pool, closest_pool=modules.closest_pool)
    np.random.seed(0)
    torch.manual_seed(0)  # identical initial replicas on every rank
    net = harness.KPFCNN(cfg, ops=mops)
    pts, lens, labels = _shard(rank)
    pyr = pyramid.build_pyramid(pts, lens, cfg, ops=gops, random_grid_orient=False)
    as_t = lambda lst, dt: [torch.from_numpy(np.ascontiguousarray(a)).to(dt) for a in lst]
    feats = torch.from_numpy(np.concatenate([np.ones((len(pts), 1), np.float32), pts[:, 2:3]], 1))
    batch = SimpleNamespace(points=as_t(pyr.points, torch.float32), neighbors=as_t(pyr.neighbors, torch.int64),
                            pools=as_t(pyr.pools, torch.int64), upsamples=as_t(pyr.upsamples, torch.int64),
                            lengths=pyr.lengths, features=feats, labels=torch.from_numpy(labels))
    return net, batch


def _grads(net):
    return torch.cat([p.grad.reshape(-1) for p in net.parameters() if p.grad is not None])


def _worker(rank, world, port, out_dir, mode="ddp"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net, batch = _build(rank)
    opt = torch.optim.SGD(net.parameters(), lr=1e-2, momentum=0.9)
    if mode == "ddp":
        ddp = torch.nn.parallel.DistributedDataParallel(net)
        loss = net.loss(ddp(batch), batch.labels)
        loss.backward()
    else:  # the bench's own averaging: early group in flight during backward, the rest afterwards
        from mvkpconv_b200 import harness
        avg = harness.OverlappedGradientAverager(net, split_level=1)
        assert 0 < len(avg.early) < len(avg.params)
        keep = mode == "overlap_kept"  # gradients zeroed in place: a stale tensor is not a finished gradient
        for _ in range(2):  # the hook must re-arm for every step
            opt.zero_grad(set_to_none=not keep)
            loss = net.loss(net(batch), batch.labels)
            sent_at_fire = []
            fire = avg._fire
            avg._fire = lambda: (fire(), sent_at_fire.append((set(avg._sent), set(avg._done))))
            loss.backward()
            avg._fire = fire
            assert avg._sent, "the early group was not sent from inside backward"
            sent, done = sent_at_fire[0]
            assert sent <= done, "a gradient went out before its accumulation finished"
            avg.finish()
    g = _grads(net)
    opt.step()
    w = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    torch.save({"g": g, "w": w, "n": batch.features.shape[0]}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("mode", ["ddp", "overlap", "overlap_kept"])
def test_two_rank_sphere_sharded_step(tmp_path, mode):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path), mode), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    # replicas agree after the all-reduce and after the optimiser step
    assert torch.equal(r0["g"], r1["g"])
    assert torch.equal(r0["w"], r1["w"])
    # ... and the all-reduced gradient is the mean of the per-shard gradients
    singles = []
    for rank in range(world):
        net, batch = _build(rank)
        net.loss(net(batch), batch.labels).backward()
        singles.append(_grads(net))
    mean = (singles[0] + singles[1]) / 2
    err = (r0["g"] - mean).abs().max() / mean.abs().max()
    assert err < 1e-5, err
    # the shards really differ (the test would be vacuous otherwise)
    assert (singles[0] - singles[1]).abs().max() > 0


# -------------------------------------------------------------------------------------------------
# Whole-scene sweep sharded by sphere (BASELINE configs[4], scene.SceneSweep): world_size-2 gloo run on CPU
# -------------------------------------------------------------------------------------------------
def _scene():
    rng = np.random.default_rng(3)
    n = 3000
    xy = rng.uniform(0, 2.4, (n, 2)) * [1.0, 0.6]
    z = 0.1 * np.sin(4 * xy[:, :1]) + rng.normal(0, 0.004, (n, 1))
    return np.concatenate([xy, z], 1).astype(np.float32)


def _sweep_net():
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from mvkpconv_b200 import harness, pyramid
    from oracle import geom, modules
    cfg = pyramid.baseline_config(architecture=list(SMALL_ARCH), first_subsampling_dl=0.05, first_features_dim=16,
                                  num_classes=5, in_features_dim=2, in_radius=0.6)
    gops = SimpleNamespace(
        batch_neighbors=geom.batch_neighbors,
        batch_grid_subsampling=lambda p, l, sampleDl=0.1, random_grid_orient=True: geom.grid_subsample_batch(p, l, sampleDl=sampleDl))
    mops = SimpleNamespace(KPConv=modules.KPConvOracle, max_pool=modules.max_pool, closest_pool=modules.closest_pool)
    np.random.seed(0)
    torch.manual_seed(0)
    return harness.KPFCNN(cfg, ops=mops), cfg, gops


class _NumpyPyramidOps:
    """SceneSweep hands tensors to the pyramid; the CPU oracle ops take numpy."""

    def __init__(self, gops):
        self.g = gops

    def batch_neighbors(self, q, s, ql, sl, r, **kw):
        out = self.g.batch_neighbors(np.asarray(q), np.asarray(s), np.asarray(ql), np.asarray(sl), r)
        lim = kw.get("max_neighbors")
        out = out[:, :lim] if lim else out
        return torch.from_numpy(np.ascontiguousarray(out)).long()

    def batch_grid_subsampling(self, p, l, sampleDl=0.1, random_grid_orient=True):
        sp, sl = self.g.batch_grid_subsampling(np.asarray(p), np.asarray(l), sampleDl=sampleDl)
        return torch.from_numpy(sp), torch.from_numpy(sl)


def _sweep_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mvkpconv_b200 import scene
    net, cfg, gops = _sweep_net()
    pts = _scene()
    centers = scene.sphere_centers(pts, cfg.in_radius, spacing=0.5)
    sweep = scene.SceneSweep(net, cfg, spheres_per_batch=2, ops=_NumpyPyramidOps(gops))
    probs, counts = sweep.run(pts, centers, rank, world)
    torch.save({"probs": probs, "counts": counts, "spheres": sweep.stats.spheres}, os.path.join(out_dir, f"sweep{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_scene_sweep(tmp_path):
    """Sharding the spheres of a scene over two ranks and merging the vote tables with one all-reduce gives the
    table of the unsharded sweep; both ranks end with the same table; every point of the scene was visited."""
    world, port = 2, _free_port()
    mp.spawn(_sweep_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "sweep0.pt")
    r1 = torch.load(tmp_path / "sweep1.pt")
    assert torch.equal(r0["probs"], r1["probs"]) and torch.equal(r0["counts"], r1["counts"])
    from mvkpconv_b200 import scene
    net, cfg, gops = _sweep_net()
    pts = _scene()
    centers = scene.sphere_centers(pts, cfg.in_radius, spacing=0.5)
    assert r0["spheres"] + r1["spheres"] == len(centers) and r0["spheres"] > 0 and r1["spheres"] > 0
    single = scene.SceneSweep(net, cfg, spheres_per_batch=2, ops=_NumpyPyramidOps(gops))
    probs, counts = single.run(pts, centers, 0, 1)
    assert torch.equal(counts, r0["counts"])
    assert int(counts.min()) >= 1, "the lattice must cover the scene"
    # batch composition differs between the sharded and the unsharded sweep (which spheres share a stacked batch):
    # eval-mode batch norm and per-sphere neighbourhoods make the per-sphere outputs independent of it
    assert float((probs - r0["probs"]).abs().max()) < 1e-4  # BLAS picks its blocking by the stacked batch size
    assert torch.allclose(probs.sum(1), torch.ones(len(pts)), atol=1e-5)
