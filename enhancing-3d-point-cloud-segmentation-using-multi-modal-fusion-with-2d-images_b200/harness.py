"""Model harness around the hot path: the KPFCNN encoder-decoder of the reference's baseline
training script, composed from this package's operators.

The reference architecture files (KPConv-PyTorch/models/architectures.py:189-352, blocks.py:387-695)
are thin Python compositions of blocks; they are out of scope as a porting target, but the
benchmark configuration ("KPConv baseline encoder-decoder, train_ScanNet_baseline shape") needs a
driver that exists on the GPU box, where the reference tree is absent.  This harness follows the
same block grammar ('simple', 'resnetb', 'resnetb_strided', 'nearest_upsample', 'unary') and the
same dimension / radius bookkeeping, so its KPConv layers have exactly the shapes of SURVEY.md
Appendix B.  KPConv / max_pool / closest_pool and the point-wise blocks (UnaryBlock,
BatchNormBlock, the fused "batch norm + residual + LeakyReLU" block tail) are this package's CUDA
ops.  The operator set is injectable so bench.py's CPU baseline can run the identical graph on
the oracle's torch-CPU KPConv with plain torch layers for the point-wise blocks (the classes
below), which is exactly what the reference does.
"""
from types import SimpleNamespace

import torch
import torch.nn as nn

from . import blocks as _blocks
from . import kpconv as _kp


def product_ops():
    return SimpleNamespace(KPConv=_kp.KPConv, max_pool=_kp.max_pool, closest_pool=_kp.closest_pool,
                           UnaryBlock=_blocks.UnaryBlock, BatchNormBlock=_blocks.BatchNormBlock, bn_act=_blocks.bn_act)


def _unary(ops, *args, **kw):
    return getattr(ops, "UnaryBlock", UnaryBlock)(*args, **kw)


def _bnblock(ops, *args):
    return getattr(ops, "BatchNormBlock", BatchNormBlock)(*args)


class BatchNormBlock(nn.Module):
    """blocks.py:430-466 (batch norm over the stacked points, or a bias when disabled)."""

    def __init__(self, in_dim, use_bn, bn_momentum):
        super().__init__()
        self.use_bn = use_bn
        if use_bn:
            self.batch_norm = nn.BatchNorm1d(in_dim, momentum=bn_momentum)
        else:
            self.bias = nn.Parameter(torch.zeros(in_dim, dtype=torch.float32))

    def forward(self, x):
        return self.batch_norm(x) if self.use_bn else x + self.bias


class UnaryBlock(nn.Module):
    """blocks.py:469-504."""

    def __init__(self, in_dim, out_dim, use_bn, bn_momentum, no_relu=False):
        super().__init__()
        self.mlp = nn.Linear(in_dim, out_dim, bias=False)
        self.batch_norm = BatchNormBlock(out_dim, use_bn, bn_momentum)
        self.no_relu = no_relu
        self.leaky_relu = nn.LeakyReLU(0.1)

    def forward(self, x, batch=None):
        x = self.batch_norm(self.mlp(x))
        return x if self.no_relu else self.leaky_relu(x)


def _geometry(block_name, layer_ind, batch):
    if 'strided' in block_name:
        return batch.points[layer_ind + 1], batch.points[layer_ind], batch.pools[layer_ind]
    return batch.points[layer_ind], batch.points[layer_ind], batch.neighbors[layer_ind]


class SimpleBlock(nn.Module):
    """blocks.py:507-561: KPConv(in, out/2) + BN + LeakyReLU."""

    def __init__(self, block_name, in_dim, out_dim, radius, layer_ind, config, ops):
        super().__init__()
        extent = radius * config.KP_extent / config.conv_radius
        self.block_name, self.layer_ind = block_name, layer_ind
        self.KPConv = ops.KPConv(config.num_kernel_points, config.in_points_dim, in_dim, out_dim // 2, extent, radius,
                                 fixed_kernel_points=config.fixed_kernel_points, KP_influence=config.KP_influence,
                                 aggregation_mode=config.aggregation_mode, deformable='deform' in block_name,
                                 modulated=config.modulated)
        self.batch_norm = _bnblock(ops, out_dim // 2, config.use_batch_norm, config.batch_norm_momentum)
        self.leaky_relu = nn.LeakyReLU(0.1)
        self.bn_act = getattr(ops, "bn_act", None)

    def forward(self, x, batch):
        q, s, inds = _geometry(self.block_name, self.layer_ind, batch)
        y = self.KPConv(q, s, inds, x)
        if self.bn_act is not None:  # the output feeds the next block's Linear layers: emit their operand format too
            return self.bn_act(y, self.batch_norm, slope=0.1, emit_hilo=True, grad_hilo=True)
        return self.leaky_relu(self.batch_norm(y))


class ResnetBottleneckBlock(nn.Module):
    """blocks.py:564-649: unary down -> KPConv(d/4, d/4) -> unary up, (max-pooled) shortcut."""

    def __init__(self, block_name, in_dim, out_dim, radius, layer_ind, config, ops):
        super().__init__()
        extent = radius * config.KP_extent / config.conv_radius
        bn, mom = config.use_batch_norm, config.batch_norm_momentum
        self.block_name, self.layer_ind, self.ops = block_name, layer_ind, ops
        self.unary1 = _unary(ops, in_dim, out_dim // 4, bn, mom) if in_dim != out_dim // 4 else nn.Identity()
        self.KPConv = ops.KPConv(config.num_kernel_points, config.in_points_dim, out_dim // 4, out_dim // 4, extent,
                                 radius, fixed_kernel_points=config.fixed_kernel_points,
                                 KP_influence=config.KP_influence, aggregation_mode=config.aggregation_mode,
                                 deformable='deform' in block_name, modulated=config.modulated)
        self.batch_norm_conv = _bnblock(ops, out_dim // 4, bn, mom)
        self.unary2 = _unary(ops, out_dim // 4, out_dim, bn, mom, no_relu=True)
        self.unary_shortcut = _unary(ops, in_dim, out_dim, bn, mom, no_relu=True) if in_dim != out_dim else nn.Identity()
        self.leaky_relu = nn.LeakyReLU(0.1)
        self.bn_act = getattr(ops, "bn_act", None)
        self.feeds_linear = True  # cleared by KPFCNN for the block whose output goes to an upsampling

    def forward(self, features, batch):
        q, s, inds = _geometry(self.block_name, self.layer_ind, batch)
        x = self.unary1(features)
        y = self.KPConv(q, s, inds, x)
        shortcut = self.ops.max_pool(features, inds) if 'strided' in self.block_name else features
        if self.bn_act is not None:  # product path: fused bn + act, and bn + residual + act tail
            x = self.bn_act(y, self.batch_norm_conv, slope=0.1, emit_hilo=True, grad_hilo=True)
            return self.unary2(x, residual=self.unary_shortcut(shortcut), slope=0.1, emit_hilo=self.feeds_linear)
        x = self.leaky_relu(self.batch_norm_conv(y))
        x = self.unary2(x)
        return self.leaky_relu(x + self.unary_shortcut(shortcut))


class NearestUpsampleBlock(nn.Module):
    """blocks.py:668-683."""

    def __init__(self, layer_ind, ops):
        super().__init__()
        self.layer_ind, self.ops = layer_ind, ops

    def forward(self, x, batch):
        return self.ops.closest_pool(x, batch.upsamples[self.layer_ind - 1])


def _block(name, radius, in_dim, out_dim, layer, config, ops):
    if name == 'unary':
        return _unary(ops, in_dim, out_dim, config.use_batch_norm, config.batch_norm_momentum)
    if name.startswith('simple'):
        return SimpleBlock(name, in_dim, out_dim, radius, layer, config, ops)
    if name.startswith('resnetb'):
        return ResnetBottleneckBlock(name, in_dim, out_dim, radius, layer, config, ops)
    if name == 'nearest_upsample':
        return NearestUpsampleBlock(layer, ops)
    raise ValueError('Unknown block name in the architecture definition : ' + name)


class KPFCNN(nn.Module):
    """Encoder-decoder segmentation net with the bookkeeping of architectures.py:189-297
    (radius doubles and feature width doubles at every strided block; skip links are taken before
    each strided block and concatenated after each upsampling)."""

    def __init__(self, config, num_classes=None, ops=None):
        super().__init__()
        self._product_ops = ops is None  # the CPU oracle graph (tests, bench cpu_baseline) injects its own ops
        self._has_deformable = None
        ops = ops or product_ops()
        arch = config.architecture
        layer, r = 0, config.first_subsampling_dl * config.conv_radius
        in_dim, out_dim = config.in_features_dim, config.first_features_dim
        self.C = num_classes or config.num_classes
        self.encoder_blocks, self.encoder_skips, skip_dims = nn.ModuleList(), [], []
        first_up = len(arch)
        for i, name in enumerate(arch):
            if any(t in name for t in ('pool', 'strided', 'upsample', 'global')):
                self.encoder_skips.append(i)
                skip_dims.append(in_dim)
            if 'upsample' in name:
                first_up = i
                break
            self.encoder_blocks.append(_block(name, r, in_dim, out_dim, layer, config, ops))
            in_dim = out_dim // 2 if 'simple' in name else out_dim
            if 'pool' in name or 'strided' in name:
                layer, r, out_dim = layer + 1, r * 2, out_dim * 2
        if len(self.encoder_blocks) and hasattr(self.encoder_blocks[-1], "feeds_linear"):
            self.encoder_blocks[-1].feeds_linear = False
        self.decoder_blocks, self.decoder_concats = nn.ModuleList(), []
        for j, name in enumerate(arch[first_up:]):
            if j > 0 and 'upsample' in arch[first_up + j - 1]:
                in_dim += skip_dims[layer]
                self.decoder_concats.append(j)
            self.decoder_blocks.append(_block(name, r, in_dim, out_dim, layer, config, ops))
            in_dim = out_dim
            if 'upsample' in name:
                layer, r, out_dim = layer - 1, r * 0.5, out_dim // 2
        self.head_mlp = _unary(ops, out_dim, config.first_features_dim, False, 0)
        # NB the reference leaves the LeakyReLU on the logits (architectures.py:296-297)
        self.head_softmax = _unary(ops, config.first_features_dim, self.C, False, 0)
        self.criterion = nn.CrossEntropyLoss(ignore_index=-1)
        self.K = config.num_kernel_points
        self.deform_fitting_power = getattr(config, "deform_fitting_power", 1.0)
        self.repulse_extent = getattr(config, "repulse_extent", 1.2)
        self.l1 = nn.L1Loss()

    def forward(self, batch, config=None):
        x = batch.features.clone().detach()
        skips = []
        for i, op in enumerate(self.encoder_blocks):
            if i in self.encoder_skips:
                skips.append(x)
            x = op(x, batch)
        j, blocks = 0, self.decoder_blocks
        while j < len(blocks):
            op = blocks[j]
            # product path: nearest upsampling + skip concatenation + unary block as one fused step
            if (self._product_ops and isinstance(op, NearestUpsampleBlock) and j + 1 < len(blocks) and
                    (j + 1) in self.decoder_concats and hasattr(blocks[j + 1], "forward_upsampled")):
                # the last decoder block feeds the head's Linear: it emits the bf16 pair of its output itself
                x = blocks[j + 1].forward_upsampled(x, batch.upsamples[op.layer_ind - 1], skips.pop(),
                                                    emit_hilo=(j + 2 == len(blocks)))
                j += 2
                continue
            if j in self.decoder_concats:
                x = torch.cat([x, skips.pop()], dim=1)
            x = op(x, batch)
            j += 1
        if self._product_ops:
            return self.head_softmax(self.head_mlp(x, batch, emit_hilo=True), batch)
        return self.head_softmax(self.head_mlp(x, batch), batch)

    def loss(self, outputs, labels):
        if self._product_ops and outputs.is_cuda:
            loss = _blocks.softmax_cross_entropy(outputs, labels, ignore_index=-1)
        else:
            loss = self.criterion(outputs.transpose(0, 1).unsqueeze(0), labels.unsqueeze(0))
        if self._has_deformable is None:
            self._has_deformable = any(getattr(m, "deformable", False) for m in self.kpconv_layers())
        if self._has_deformable:
            loss = loss + self.fitting_regularizer()
        return loss

    def fitting_regularizer(self):
        """Point-to-point fitting + repulsive regulariser of the deformed kernel points
        (architectures.py:21-54): pulls every deformed kernel point towards its closest input point
        (min_d2) and pushes kernel points of one neighbourhood apart."""
        fitting, repulsive = 0, 0
        for m in self.kpconv_layers():
            if not getattr(m, "deformable", False):
                continue
            # fitting: mean over (point, kernel point) of the squared distance to the closest input point
            fitting = fitting + (m.min_d2 / (m.KP_extent ** 2)).abs().mean()
            # repulsion: all kernel-point pairs of a neighbourhood at once; pair (i, j) pushes point i away from a
            # DETACHED point j (architectures.py:41-48), the self pair (distance 0, hinge repulse_extent^2) is masked
            locs = m.deformed_KP / m.KP_extent                                   # [N, K, 3]
            diff = locs.unsqueeze(2) - locs.detach().unsqueeze(1)                # [N, K (i), K (j), 3]
            dist = torch.sqrt((diff * diff).sum(-1) + torch.eye(self.K, device=locs.device, dtype=locs.dtype))
            hinge = torch.clamp_max(dist - self.repulse_extent, max=0.0) ** 2
            hinge = hinge * (1.0 - torch.eye(self.K, device=locs.device, dtype=locs.dtype))
            repulsive = repulsive + hinge.sum(2).mean(0).sum() / self.K  # sum_i mean_n sum_{j != i} hinge / K
        return self.deform_fitting_power * (2 * fitting + repulsive)

    def kpconv_layers(self):
        """Top-level KPConv modules (the offset convolutions nested inside deformable ones excluded)."""
        nested = {id(m.offset_conv) for m in self.modules() if getattr(m, "offset_conv", None) is not None}
        return [m for m in self.modules() if type(m).__name__.startswith("KPConv") and id(m) not in nested]


class OverlappedGradientAverager:
    """Gradient averaging over the ranks of a sphere-sharded job, overlapped with backward.

    Backward produces the gradients of the decoder and of the deep encoder levels first -- the bulk of the
    parameter bytes (level >= 3 holds ~90 % of the 24.4 M parameters of the baseline network) -- and spends
    most of its time afterwards in the shallow, point-heavy levels.  A post-accumulate hook on the LAST
    parameter of that early group starts its all-reduce asynchronously (NCCL's own stream), so the transfer
    hides behind the rest of backward; ``finish()`` (call it after ``loss.backward()``) reduces whatever is
    left, waits, and divides by the world size.  No buckets, no copies.  A gradient counts as complete only once
    its own post-accumulate hook has run in THIS backward pass, so gradients kept across steps
    (``zero_grad(set_to_none=False)``, gradient accumulation) and the trigger block's parameters that accumulate
    after the trigger are left to ``finish()`` instead of being sent while autograd still writes them.

    STATUS: correct (tests/test_distributed_cpu.py) but NOT the default of bench.py: on 2 B200s the concurrent
    NCCL kernel slowed the step from 14.0 to 38.8 ms (round 1: the persistent contraction CTAs hold 227 KB of
    shared memory per SM, so the collective's CTAs and they cannot share an SM and each waits for the other);
    the default is one flat all-reduce between the two captured graphs of ``GraphedTrainStep``.  One
    ``backward()`` per ``finish()``.

        avg = OverlappedGradientAverager(net, split_level=3)     # once
        loss.backward(); avg.finish()                            # every step
    """

    def __init__(self, net, split_level=3, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        self.coalesce = dist.get_backend(group) == "nccl"
        late_ids, trigger = set(), None
        for blk in net.encoder_blocks:
            if getattr(blk, "layer_ind", 0) < split_level:
                late_ids.update(id(p) for p in blk.parameters())
        for blk in net.encoder_blocks:  # first block of the early group = the one whose backward runs last
            if getattr(blk, "layer_ind", 0) >= split_level:
                ps = [p for p in blk.parameters() if p.requires_grad]
                trigger = ps[0] if ps else None
                break
        params = [p for p in net.parameters() if p.requires_grad]
        self.early = [p for p in params if id(p) not in late_ids]
        self.params = params
        self._works, self._sent, self._done = [], set(), set()
        self._trigger = trigger
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.early] if trigger is not None else []

    def _reduce(self, tensors, async_op):
        dist = self.dist
        if not tensors:
            return
        if self.coalesce:
            with dist._coalescing_manager(group=self.group, device=tensors[0].device, async_ops=async_op) as cm:
                for t in tensors:
                    dist.all_reduce(t, group=self.group)
            if async_op:
                self._works.append(cm)
        else:
            for t in tensors:
                w = dist.all_reduce(t, group=self.group, async_op=async_op)
                if async_op:
                    self._works.append(w)

    def _on_grad(self, param):
        self._done.add(id(param))
        if param is self._trigger:
            self._fire()

    def _fire(self):
        if self._sent:  # a second backward before finish(): leave everything to finish()
            return
        ready = [p for p in self.early if id(p) in self._done and p.grad is not None]
        self._sent = {id(p) for p in ready}
        self._reduce([p.grad for p in ready], async_op=True)

    def finish(self):
        rest = [p.grad for p in self.params if p.grad is not None and id(p) not in self._sent]
        self._reduce(rest, async_op=False)
        for w in self._works:
            w.wait()
        self._works, self._sent, self._done = [], set(), set()
        grads = [p.grad for p in self.params if p.grad is not None]
        if grads:
            torch._foreach_div_(grads, float(self.world))


class GraphedTrainStep:
    """One training step (forward, loss, backward, gradient clipping, optimiser update) replayed as a CUDA graph.

    The eager step is ~590 short launches issued from Python (autograd Functions, scratch allocations, tensor-map
    encodes): the host needs as long to enqueue it as the B200 needs to run it.  Once the shapes of a batch are
    known -- the pyramid's level sizes and row widths -- nothing in the step depends on the host any more, so the
    whole launch sequence is captured once per shape signature and replayed with one ``cudaGraphLaunch``.  The
    pyramid itself (data-dependent sizes, a few read-backs) stays eager on its side stream; its tensors are copied
    into the graph's static input buffers (a few hundred MB of device-to-device traffic, < 0.1 ms).

    * ``signature`` = shapes / dtypes of every pyramid tensor + features + labels.  A batch with a new signature
      runs eagerly the first ``warm`` times it is seen (optimiser state, weight-operand registry and kernel
      attributes must exist before a capture) and is captured afterwards; ``max_graphs`` signatures are kept
      (least recently used first out), each with its own memory pool.  A loader that pads its level sizes to
      buckets keeps the number of signatures small.
    * single rank: ONE graph holds forward + backward + clip + optimiser step.  ``reduce_grads`` given (sharded
      job): graph A = forward + backward, then ``reduce_grads(list of gradient tensors)`` runs eagerly (NCCL), then
      graph B = clip + optimiser step.
    * the optimiser must be capture-safe with its state already initialised (``torch.optim.SGD(fused=True)`` with
      a float learning rate bakes the rate into the graph; use a tensor ``lr`` to schedule it).
    """

    def __init__(self, net, optimizer, grad_clip=100.0, reduce_grads=None, max_graphs=4, warm=2, flat_grads=True):
        from collections import OrderedDict
        self.net, self.opt, self.grad_clip, self.reduce_grads = net, optimizer, grad_clip, reduce_grads
        self.params = [p for p in net.parameters() if p.requires_grad]
        self.graphs = OrderedDict()
        self.seen = {}
        self.max_graphs, self.warm, self.flat_grads = max_graphs, warm, flat_grads
        self.launches_per_step = None  # libmvk launches captured per step (replays do not pass through the C ABI)
        self.replays = 0

    @staticmethod
    def signature(pyr, features, labels, extras=None):
        sig = []
        for lst in (pyr.points, pyr.neighbors, pyr.pools, pyr.upsamples, pyr.lengths):
            sig.append(tuple((tuple(t.shape), str(t.dtype)) for t in lst))
        sig.append((tuple(features.shape), str(features.dtype), tuple(labels.shape), str(labels.dtype)))
        for k in sorted(extras or {}):
            sig.append((k, tuple(extras[k].shape), str(extras[k].dtype)))
        return tuple(sig)

    @staticmethod
    def _batch(pyr, features, labels, extras):
        """The batch object the network consumes: the pyramid lists, `features` (KPFCNN) and any extra tensors by
        name (the fusion nets read images / image_xyz / knn_global / feat_aggre_points / feature_3d)."""
        return SimpleNamespace(points=pyr.points, neighbors=pyr.neighbors, pools=pyr.pools, upsamples=pyr.upsamples,
                               lengths=pyr.lengths, features=features, labels=labels, **(extras or {}))

    def _finish(self, grads=None):
        if self.reduce_grads is not None:
            self.reduce_grads(grads if grads is not None else [p.grad for p in self.params if p.grad is not None])
        if self.grad_clip:
            torch.nn.utils.clip_grad_value_(self.params, self.grad_clip)  # utils/trainer.py:191-193
        self.opt.step()

    def eager(self, pyr, features, labels, extras=None):
        batch = self._batch(pyr, features, labels, extras)
        loss = self.net.loss(self.net(batch), labels)
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self._finish()
        return loss.detach()

    def _capture(self, pyr, features, labels, extras=None):
        from . import _lib
        clone = lambda lst: [t.clone() for t in lst]
        spyr = SimpleNamespace(points=clone(pyr.points), neighbors=clone(pyr.neighbors), pools=clone(pyr.pools),
                               upsamples=clone(pyr.upsamples), lengths=clone(pyr.lengths))
        static = self._batch(spyr, features.clone(), labels.clone(), {k: v.clone() for k, v in (extras or {}).items()})
        static.extras = sorted(extras or {})
        L = _lib.lib()
        entry = SimpleNamespace(static=static, graph_a=torch.cuda.CUDAGraph(), graph_b=None, loss=None, grads=None, pack=None,
                                flat=None)
        self.opt.zero_grad(set_to_none=True)
        # tensors a module keeps from the previous (eager) step -- the deformable layers' min_d2 / deformed_KP /
        # offset_features that the regulariser reads -- hold that step's autograd graph alive, including the
        # AccumulateGrad nodes of their parameters, which are bound to the stream of the eager step
        for m in self.net.modules():
            for attr in ("min_d2", "deformed_KP", "offset_features"):
                if getattr(m, attr, None) is not None:
                    setattr(m, attr, None)
        torch.cuda.synchronize()
        _lib._ZEROS.buf = None  # zero-initialised scratch must be allocated (and re-zeroed on replay) inside the graph
        l0 = L.mvk_launch_count()
        with torch.cuda.graph(entry.graph_a):
            loss = self.net.loss(self.net(static), static.labels)
            loss.backward()
            if self.reduce_grads is None:
                self._finish()
            entry.loss = loss.detach()
        _lib._ZEROS.buf = None
        entry.grads = [p.grad for p in self.params if p.grad is not None]
        if self.reduce_grads is not None and not self.flat_grads:
            entry.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(entry.graph_b):
                if self.grad_clip:
                    torch.nn.utils.clip_grad_value_(self.params, self.grad_clip)
                self.opt.step()
            _lib._ZEROS.buf = None
        elif self.reduce_grads is not None:
            # sharded job: the ~190 gradient tensors are packed into ONE flat buffer at the end of graph A (a
            # multi-tensor copy, ~30 us for 97.5 MB) so that the collective between the two graphs is a single
            # all-reduce instead of a group of 190 latency-bound ones; graph B (clip + optimiser step) reads the
            # parameters' gradients straight from views of that buffer: nothing is copied back
            owners = [p for p in self.params if p.grad is not None]
            pad4 = lambda n: (n + 3) // 4 * 4  # every view starts on a 16-byte boundary (vectorised multi-tensor kernels)
            flat = torch.zeros(sum(pad4(g.numel()) for g in entry.grads), dtype=entry.grads[0].dtype,
                               device=entry.grads[0].device)
            views, off = [], 0
            for g in entry.grads:
                views.append(flat[off:off + g.numel()].view_as(g))
                off += pad4(g.numel())
            pack = torch.cuda.CUDAGraph()
            with torch.cuda.graph(pack):
                torch._foreach_copy_(views, entry.grads)
            entry.pack, entry.flat = pack, flat
            for p, v in zip(owners, views):
                p.grad = v
            entry.graph_b = torch.cuda.CUDAGraph()
            with torch.cuda.graph(entry.graph_b):
                if self.grad_clip:
                    torch.nn.utils.clip_grad_value_(self.params, self.grad_clip)
                self.opt.step()
            _lib._ZEROS.buf = None
        self.launches_per_step = int(L.mvk_launch_count() - l0)
        return entry

    def __call__(self, pyr, features, labels, extras=None):
        """Runs one step on (pyramid, features, labels[, extra batch tensors by name]); returns the loss (a device
        scalar that the next call overwrites -- read or copy it before)."""
        from . import _weights
        sig = self.signature(pyr, features, labels, extras)
        entry = self.graphs.get(sig)
        if entry is None:
            n = self.seen.get(sig, 0)
            self.seen[sig] = n + 1
            if n < self.warm:
                return self.eager(pyr, features, labels, extras)
            entry = self._capture(pyr, features, labels, extras)  # records only; the static buffers hold this batch already
            self.graphs[sig] = entry
            while len(self.graphs) > self.max_graphs:
                self.graphs.popitem(last=False)
        else:
            self.graphs.move_to_end(sig)
            st = entry.static
            for dst, src in ((st.points, pyr.points), (st.neighbors, pyr.neighbors), (st.pools, pyr.pools),
                             (st.upsamples, pyr.upsamples), (st.lengths, pyr.lengths)):
                torch._foreach_copy_(dst, src)
            st.features.copy_(features)
            st.labels.copy_(labels)
            for k in st.extras:
                getattr(st, k).copy_(extras[k])
        entry.graph_a.replay()
        if entry.graph_b is not None:
            if entry.pack is not None:
                entry.pack.replay()
                self.reduce_grads([entry.flat])
            else:
                self.reduce_grads(entry.grads)
            entry.graph_b.replay()
        self.replays += 1
        _weights.invalidate()  # the parameters moved without passing through torch.optim's step hook
        return entry.loss
