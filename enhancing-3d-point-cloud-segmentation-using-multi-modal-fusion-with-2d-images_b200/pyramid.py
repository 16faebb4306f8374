"""Per-batch input pyramid on the GPU: the direct caller of batch_neighbors / batch_grid_subsampling.

Mirror of ``PointCloudDataset.segmentation_inputs`` (reference: KPConv-PyTorch/datasets/common.py:
536-652) and ``big_neighborhood_filter`` (:411-421): for every level of the architecture

    conv_i  = batch_neighbors(points, points, r)                         (:585)
    pool_p  = batch_grid_subsampling(points, dl = 2 r / conv_radius)     (:601)
    pool_i  = batch_neighbors(pool_p, points, r)                         (:611)
    up_i    = batch_neighbors(points, pool_p, 2 r)                       (:614)

with the neighbour matrices cropped to ``neighborhood_limits[level]`` (upsample: level + 1).  The
reference runs this on CPU worker processes; here every call is a CUDA launch on tensors that stay
in HBM.  The ops are injectable so the CPU baseline can drive the same pyramid logic with the
reference's own C++ (bench.py --impl reference).
"""
from types import SimpleNamespace

import numpy as np
import torch

from . import geometry

# train_ScanNet_baseline.py:126-147 ("rigid deeper")
BASELINE_ARCHITECTURE = [
    'simple', 'resnetb', 'resnetb_strided',
    'resnetb', 'resnetb', 'resnetb_strided',
    'resnetb', 'resnetb', 'resnetb_strided',
    'resnetb', 'resnetb', 'resnetb_strided',
    'resnetb', 'resnetb',
    'nearest_upsample', 'unary', 'nearest_upsample', 'unary',
    'nearest_upsample', 'unary', 'nearest_upsample', 'unary',
]


def baseline_config(**overrides):
    """The fields of the reference Config (utils/config.py, train_ScanNet_baseline.py:39-264) that
    parameterise the hot path."""
    cfg = SimpleNamespace(
        architecture=list(BASELINE_ARCHITECTURE), num_kernel_points=15, in_points_dim=3,
        first_subsampling_dl=0.04, conv_radius=2.5, deform_radius=6.0, KP_extent=1.2,
        KP_influence='linear', aggregation_mode='sum', fixed_kernel_points='center', modulated=False,
        first_features_dim=128, in_features_dim=2, use_batch_norm=True, batch_norm_momentum=0.02,
        in_radius=2.0, num_classes=20, neighborhood_limits=[], deform_fitting_mode='point2point',
        deform_fitting_power=1.0, repulse_extent=1.2)
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def device_ops():
    return SimpleNamespace(batch_neighbors=geometry.batch_neighbors,
                           batch_grid_subsampling=geometry.batch_grid_subsampling)


def _empty(like, shape, dtype):
    if isinstance(like, torch.Tensor):
        return torch.zeros(shape, dtype=dtype, device=like.device)
    return np.zeros(shape, dtype={torch.int32: np.int32, torch.float32: np.float32, torch.int64: np.int64}[dtype])


def build_pyramid(stacked_points, stack_lengths, config, ops=None, random_grid_orient=True, index_dtype=torch.int64):
    """Returns a namespace with lists points / neighbors / pools / upsamples / lengths (one entry per
    level), exactly the lists ``segmentation_inputs`` concatenates (common.py:646-650)."""
    ops = ops or device_ops()
    limits = list(getattr(config, "neighborhood_limits", []) or [])
    r_normal = config.first_subsampling_dl * config.conv_radius
    out = SimpleNamespace(points=[], neighbors=[], pools=[], upsamples=[], lengths=[])
    layer_blocks = []
    on_device = isinstance(stacked_points, torch.Tensor)

    deferred = []  # neighbour calls whose max-count read-back is postponed to ONE sync at the end
    slots = []     # (list, position) of the matrix each deferred call produced

    def nb(q, s, ql, sl, r, level):
        lim = limits[level] if level < len(limits) else None
        if on_device:
            if lim and ops.batch_neighbors is geometry.batch_neighbors:
                return ops.batch_neighbors(q, s, ql, sl, r, max_neighbors=lim, out_dtype=index_dtype, deferred=deferred)
            return ops.batch_neighbors(q, s, ql, sl, r, max_neighbors=lim, out_dtype=index_dtype)
        res = ops.batch_neighbors(q, s, ql, sl, r)
        return res[:, :lim] if lim else res  # big_neighborhood_filter

    for block in config.architecture:
        if not ('pool' in block or 'strided' in block or 'global' in block or 'upsample' in block):
            layer_blocks.append(block)
            continue
        level = len(out.points)
        # deformable layers search a wider neighbourhood (common.py:461-463, 485-487)
        r_deform = r_normal * config.deform_radius / config.conv_radius
        if layer_blocks:
            r = r_deform if any('deformable' in b for b in layer_blocks) else r_normal
            conv_i = nb(stacked_points, stacked_points, stack_lengths, stack_lengths, r, level)
        else:
            conv_i = _empty(stacked_points, (0, 1), torch.int32)
        if 'pool' in block or 'strided' in block:
            dl = 2 * r_normal / config.conv_radius
            pool_p, pool_b = ops.batch_grid_subsampling(stacked_points, stack_lengths, sampleDl=dl,
                                                        random_grid_orient=random_grid_orient)
            r = r_deform if 'deformable' in block else r_normal
            pool_i = nb(pool_p, stacked_points, pool_b, stack_lengths, r, level)
            up_i = nb(stacked_points, pool_p, stack_lengths, pool_b, 2 * r, level + 1)
        else:
            pool_i = _empty(stacked_points, (0, 1), torch.int32)
            pool_p = _empty(stacked_points, (0, 3), torch.float32)
            pool_b = _empty(stacked_points, (0,), torch.int32)
            up_i = _empty(stacked_points, (0, 1), torch.int32)
        out.points.append(stacked_points)
        for lst, mat in ((out.neighbors, conv_i), (out.pools, pool_i), (out.upsamples, up_i)):
            if any(mat is d.out for d in deferred[len(slots):]):
                slots.append((lst, len(lst)))
            lst.append(mat)
        out.lengths.append(stack_lengths)
        stacked_points, stack_lengths = pool_p, pool_b
        r_normal *= 2
        layer_blocks = []
        if 'global' in block or 'upsample' in block:
            break
    if deferred:
        assert len(slots) == len(deferred)
        for i, mat in geometry.resolve_deferred(deferred).items():
            lst, pos = slots[i]
            lst[pos] = mat
    return out


def calibrate_neighborhood_limits(stacked_points, stack_lengths, config, ops=None, keep=0.9):
    """Per-level row-width limits keeping `keep` of the neighbourhoods untouched -- the role of the
    reference sampler calibration (ScanNet_sphere_color.py:1272-1522, percentile at :1463-1464),
    computed here from one batch."""
    ops = ops or device_ops()
    cfg = SimpleNamespace(**{**config.__dict__, "neighborhood_limits": []})
    pyr = build_pyramid(stacked_points, stack_lengths, cfg, ops=ops, random_grid_orient=False)
    limits = []
    for lvl, conv_i in enumerate(pyr.neighbors):
        if conv_i.shape[0] == 0:
            limits.append(limits[-1] if limits else 1)
            continue
        ns = pyr.points[lvl].shape[0]
        t = conv_i if isinstance(conv_i, torch.Tensor) else torch.from_numpy(np.asarray(conv_i))
        counts = (t < ns).sum(1).float()
        limits.append(max(1, int(torch.quantile(counts.cpu(), keep).item())))
    return limits


class PyramidPrefetcher:
    """Builds the pyramid of the NEXT batch on a side CUDA stream while the current batch trains.

    The reference overlaps the same two things with processes: ``segmentation_inputs`` runs inside the
    DataLoader workers (``num_workers = config.input_threads``, training scripts e.g.
    train_ScanNet_baseline.py:303-315) while the trainer consumes the previous batch
    (utils/trainer.py:160-200).  Here one host thread does both: the pyramid's launches and its few
    size read-backs go to ``self.stream``, so a read-back waits for the side stream only and never
    drains the backlog of forward/backward kernels queued on the training stream.

        pf = PyramidPrefetcher(cfg, device)
        pf.submit(lambda: (points, lengths, extras))    # batch 0
        for ...:
            pyr, extras = pf.take()                     # batch i, ready on the current stream
            ... enqueue forward / backward / optimiser of batch i ...
            pf.submit(loader_for_batch_i_plus_1)        # its launches overlap batch i's kernels

    ``make_inputs`` runs with the side stream current, so host->device copies issued inside it
    (``.to(device, non_blocking=True)`` from pinned memory) are part of the pipeline too.  Every tensor
    handed out by ``take`` is marked with ``record_stream`` so that the caching allocator does not
    recycle it for the next pyramid while training kernels still read it.
    """

    def __init__(self, config, device, index_dtype=torch.int64, random_grid_orient=True):
        self.config = config
        self.device = torch.device(device)
        self.index_dtype = index_dtype
        self.random_grid_orient = random_grid_orient
        # high priority: the pyramid's short kernels must not queue behind the training backlog, or each of
        # its size read-backs would stall the host for as long as that backlog takes to drain
        self.stream = torch.cuda.Stream(self.device, priority=-1)
        # workspaces / inputs last touched on the training stream must be complete before first use here
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        self._pending = None

    def submit(self, make_inputs):
        if self._pending is not None:
            raise RuntimeError("PyramidPrefetcher.submit: the previous pyramid was not taken")
        with torch.cuda.stream(self.stream):
            points, lengths, extras = make_inputs()
            pyr = build_pyramid(points, lengths, self.config, random_grid_orient=self.random_grid_orient,
                                index_dtype=self.index_dtype)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        self._pending = (pyr, extras, ready)

    def take(self):
        if self._pending is None:
            raise RuntimeError("PyramidPrefetcher.take: nothing was submitted")
        pyr, extras, ready = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ready)
        seen = set()

        def mark(obj):
            if isinstance(obj, torch.Tensor):
                if obj.is_cuda and obj.data_ptr() not in seen:
                    seen.add(obj.data_ptr())
                    obj.record_stream(cur)
            elif isinstance(obj, (list, tuple)):
                for o in obj:
                    mark(o)
            elif isinstance(obj, dict):
                mark(list(obj.values()))
            elif isinstance(obj, SimpleNamespace):
                mark(list(vars(obj).values()))
        mark(pyr)
        mark(extras)
        return pyr, extras
