"""Whole-scene inference sweep: cover a scene with overlapping spheres, run the network on stacked batches of
spheres, accumulate per-point class votes; spheres are sharded round-robin over the ranks of a job.

Reference: the voting loop of ``ModelTester.cloud_segmentation_test`` (KPConv-PyTorch/utils/tester.py:100-260):
for every batch of spheres the logits go through a softmax and are blended into a per-scene probability table,

    test_probs[inds] = test_smooth * test_probs[inds] + (1 - test_smooth) * probs            (:199)

with the spheres drawn by the dataset's potential sampler until every point has been seen a few times
(ScanNet_sphere_color.py ``potential_item``).  Here the scene is covered by a FIXED lattice of overlapping
spheres (BASELINE configs[4]); everything per batch -- sphere cropping, the 5-level
pyramid (grid subsampling + radius search), the KPConv stack, softmax and the vote scatter -- runs on the GPU.

Vote rules:
    'smooth'  the reference's exponential blend, applied sphere by sphere in visiting order (order dependent,
              like the reference: a single rank reproduces the reference's table for the same visiting order);
    'mean'    sum of probabilities and visit counts per point: commutative, so the table is the same however the
              spheres are sharded; the ranks' partial tables are summed with ONE all-reduce at the end.
"""
from types import SimpleNamespace

import numpy as np
import torch

from . import pyramid as _pyr


def sphere_centers(points, in_radius, spacing=None, z=None):
    """Fixed lattice of sphere centres covering the bounding box of `points` (numpy or tensor, [N, 3]).
    spacing defaults to in_radius (neighbouring spheres overlap by half); centres sit at height `z` (default: the
    mid height of the scene, like the reference's spheres that are centred on scene points)."""
    p = points.detach().cpu().numpy() if isinstance(points, torch.Tensor) else np.asarray(points)
    lo, hi = p.min(0), p.max(0)
    s = float(spacing or in_radius)
    xs = np.arange(lo[0] + 0.5 * s, hi[0], s) if hi[0] - lo[0] > s else np.array([(lo[0] + hi[0]) / 2])
    ys = np.arange(lo[1] + 0.5 * s, hi[1], s) if hi[1] - lo[1] > s else np.array([(lo[1] + hi[1]) / 2])
    zc = float((lo[2] + hi[2]) / 2 if z is None else z)
    return np.array([[x, y, zc] for x in xs for y in ys], np.float32)


def shard(n_items, rank, world):
    """Round-robin assignment of items to ranks."""
    return list(range(rank, n_items, world))


class SceneSweep:
    """Sweeps one scene with a trained network.

        sweep = SceneSweep(net, config, features_fn)
        probs = sweep.run(scene_points, centers, rank, world)      # [N, C] vote table (all ranks: same table)

    `features_fn(world_points [n, 3], centred_points [n, 3]) -> [n, in_features_dim]` builds the input features of
    a stacked batch (baseline: constant 1 and the height, train_ScanNet_baseline.py:183).
    """

    def __init__(self, net, config, features_fn=None, spheres_per_batch=8, vote="mean", test_smooth=0.95,
                 ops=None):
        if vote not in ("mean", "smooth"):
            raise ValueError("vote must be 'mean' or 'smooth'")
        self.net, self.cfg, self.ops = net, config, ops
        self.features_fn = features_fn or (lambda wp, cp: torch.cat([torch.ones_like(wp[:, :1]), wp[:, 2:3]], 1))
        self.spheres_per_batch, self.vote, self.test_smooth = spheres_per_batch, vote, test_smooth
        self.stats = SimpleNamespace(batches=0, spheres=0, points=0, queries=0)

    def crop(self, scene, centers):
        """Indices and stacked centred coordinates of the points within in_radius of each centre.
        scene [N, 3] tensor; centers [B, 3] tensor.  Returns (inds [n] int64, centred [n, 3], lengths [B] int32)."""
        r2 = float(self.cfg.in_radius) ** 2
        d2 = ((scene[None, :, :] - centers[:, None, :]) ** 2).sum(-1)  # [B, N]
        sel = d2 < r2
        b_idx, p_idx = torch.nonzero(sel, as_tuple=True)              # sorted by sphere, then by point index
        lengths = sel.sum(1).to(torch.int32)
        centred = scene[p_idx] - centers[b_idx]
        return p_idx, centred.contiguous(), lengths

    @torch.no_grad()
    def run_batch(self, scene, centers, votes, counts):
        inds, centred, lengths = self.crop(scene, centers)
        keep = lengths > 0
        if not bool(keep.all()):
            lengths = lengths[keep]
        if inds.numel() == 0:
            return 0
        pyr = _pyr.build_pyramid(centred, lengths, self.cfg, ops=self.ops, random_grid_orient=False)
        batch = SimpleNamespace(points=pyr.points, neighbors=pyr.neighbors, pools=pyr.pools, upsamples=pyr.upsamples,
                                lengths=pyr.lengths, features=self.features_fn(scene[inds], centred))
        probs = torch.softmax(self.net(batch), dim=1)
        if self.vote == "mean":
            votes.index_add_(0, inds, probs)
            counts.index_add_(0, inds, torch.ones_like(inds, dtype=counts.dtype))
        else:  # tester.py:199, sphere by sphere (a point may sit in several spheres of one batch)
            o = 0
            for n in lengths.tolist():
                ii = inds[o:o + n]
                votes[ii] = self.test_smooth * votes[ii] + (1 - self.test_smooth) * probs[o:o + n]
                o += n
            counts.index_add_(0, inds, torch.ones_like(inds, dtype=counts.dtype))
        self.stats.batches += 1
        self.stats.spheres += int(lengths.numel())
        self.stats.points += int(inds.numel())
        self.stats.queries += sum(int(t.shape[0]) for t in pyr.neighbors + pyr.pools + pyr.upsamples)
        return int(inds.numel())

    def run(self, scene_points, centers, rank=0, world=1, group=None):
        """Returns (probs [N, C] float32, counts [N] int32) -- the merged table on every rank."""
        was_training = self.net.training
        self.net.eval()
        dev = next(self.net.parameters()).device
        scene = torch.as_tensor(scene_points, dtype=torch.float32, device=dev)
        cen = torch.as_tensor(centers, dtype=torch.float32, device=dev)
        votes = torch.zeros((scene.shape[0], self.net.C), dtype=torch.float32, device=dev)
        counts = torch.zeros(scene.shape[0], dtype=torch.int32, device=dev)
        mine = shard(len(cen), rank, world)
        for i in range(0, len(mine), self.spheres_per_batch):
            self.run_batch(scene, cen[mine[i:i + self.spheres_per_batch]], votes, counts)
        if world > 1:
            if self.vote != "mean":
                raise RuntimeError("sharded sweeps need the commutative 'mean' vote rule")
            import torch.distributed as dist
            dist.all_reduce(votes, group=group)
            dist.all_reduce(counts, group=group)
        if self.vote == "mean":
            votes = votes / counts.clamp_min(1).unsqueeze(1).to(votes.dtype)
        self.net.train(was_training)
        return votes, counts
