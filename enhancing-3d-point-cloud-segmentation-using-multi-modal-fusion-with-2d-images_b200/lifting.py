"""2D -> 3D lifting ops: host-side mirror of the reference interfaces, CUDA underneath.

    group_points(points, index)                      mvpnet/ops/group_points.py:5-31
    FeatureAggregation(in_channels, mlp_channels=(64, 64, 64), reduction='sum', use_relation=True)
                                                     mvpnet/models/mvpnet_3d.py:12-70
    depth2xyz(cam_matrix, depth)                     datasets/ScanNet_sphere_color.py:66-72
    unproject_views(cam_matrix, depths, poses)       :409-417 for all views of a sphere at once
    knn_pixels(xyz64, mask, queries, k=3)            :442-452 (replaces sklearn ball-tree kNN)

FeatureAggregation keeps the reference's parameter names (mlp.{i}.conv.weight, mlp.{i}.bn.*) so
mvpnet checkpoints load with load_state_dict.
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr, stream_ptr
from .geometry import _workspace


# -------------------------------------------------------------------------------------------------
class GroupPointsFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, index):
        _lib.require_cuda()
        L = _lib.lib()
        if not points.is_cuda:
            raise RuntimeError("group_points: tensors must live on a CUDA device (no CPU fallback)")
        if points.dim() != 3 or index.dim() != 3 or points.size(0) != index.size(0):
            raise RuntimeError("group_points: expected points (b, c, n1) and index (b, n2, k)")
        p = points.detach().contiguous().float()
        idx = index.detach().contiguous().long()
        b, c, n1 = p.shape
        _, n2, k = idx.shape
        out = torch.empty((b, c, n2, k), dtype=torch.float32, device=p.device)
        with _lib.on_device(p.device):
            check(L.mvk_group_points(ptr(p), b, c, n1, ptr(idx), n2, k, ptr(out), stream_ptr()))
        ctx.save_for_backward(idx)
        ctx.num_points = n1
        return out

    @staticmethod
    def backward(ctx, *grad_output):
        L = _lib.lib()
        idx = ctx.saved_tensors[0]
        go = grad_output[0].detach().contiguous().float()
        b, c, n2, k = go.shape
        gi = torch.zeros((b, c, ctx.num_points), dtype=torch.float32, device=go.device)
        with _lib.on_device(go.device):
            check(L.mvk_group_points_bwd(ptr(go), b, c, ctx.num_points, ptr(idx), n2, k, ptr(gi), stream_ptr()))
        return gi, None


def group_points(points, index):
    """out[b, c, n, k] = points[b, c, index[b, n, k]]  (mvpnet/ops/group_points.py:5-31).

    points (B, C, N1) float32, index (B, N2, K) int64 -> (B, C, N2, K); differentiable w.r.t. points.
    """
    return GroupPointsFunction.apply(points, index)


# -------------------------------------------------------------------------------------------------
class _Conv2dBNReLU(nn.Module):
    """Parameter container with the reference's names (common/nn/modules/conv.py:29-51)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, 1, bias=False)
        self.bn = nn.BatchNorm2d(out_channels)


class FeatureAggregation(nn.Module):
    """Feature Aggregation inspired by ContFuse (mvpnet_3d.py:12-70): relation features
    [diff_xyz, |diff|^2] concatenated to the k gathered 2D features, a shared MLP of
    1x1 conv (no bias) + BatchNorm2d + ReLU layers, then sum / max over the k neighbours.

    The forward pass runs libmvk kernels: one row-GEMM per layer with the previous layer's
    batch-norm + ReLU folded into its operand load and the batch statistics of its own output
    accumulated on the fly (fp64), then a fused bn + relu + reduce.
    """

    def __init__(self, in_channels, mlp_channels=(64, 64, 64), reduction='sum', use_relation=True):
        super(FeatureAggregation, self).__init__()
        self.in_channels = in_channels
        self.use_relation = use_relation
        if not use_relation:
            raise NotImplementedError("use_relation=False is not used by MV-KPConv; not implemented")
        if mlp_channels:
            self.out_channels = mlp_channels[-1]
            self.mlp = nn.ModuleList()
            c_in = in_channels + 4
            for c_out in mlp_channels:
                self.mlp.append(_Conv2dBNReLU(c_in, c_out))
                c_in = c_out
        else:
            raise NotImplementedError("mlp_channels=() (pure reduction) is not used by MV-KPConv")
        if reduction not in ('sum', 'max'):
            raise ValueError("reduction must be 'sum' or 'max'")
        self.reduction = reduction
        self.reset_parameters()

    def reset_parameters(self):
        # common/nn/init.py:22-26 applied to every conv (mvpnet_3d.py:66-70)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)

    # ---- device pipeline on row-major [np*k, C] activations ---------------------------------
    def _run(self, X, np_, k):
        L = _lib.lib()
        dev = X.device
        rows = np_ * k
        st = stream_ptr()
        scale = shift = None
        cur, ldx = X, X.shape[1]
        for layer in self.mlp:
            W = layer.conv.weight.detach().reshape(layer.conv.out_channels, -1).contiguous().float()
            cout, cin = W.shape
            Y = torch.empty((rows, cout), dtype=torch.float32, device=dev)
            stats = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
            check(L.mvk_fa_layer(ptr(cur), rows, cin, ldx, ptr(scale), ptr(shift), ptr(W), cout, ptr(Y),
                                 ptr(stats), st))
            bn = layer.bn
            if self.training or not bn.track_running_stats:
                mean = stats[:cout] / rows
                var = (stats[cout:] / rows - mean * mean).clamp_min(0.0)
                if bn.track_running_stats:
                    with torch.no_grad():
                        m = bn.momentum if bn.momentum is not None else 0.1
                        unbiased = var * (rows / max(rows - 1, 1))
                        bn.running_mean.mul_(1 - m).add_(m * mean.float())
                        bn.running_var.mul_(1 - m).add_(m * unbiased.float())
                        bn.num_batches_tracked += 1
            else:
                mean, var = bn.running_mean.double(), bn.running_var.double()
            inv = torch.rsqrt(var + bn.eps)
            g = bn.weight.detach().double()
            scale = (g * inv).float().contiguous()
            shift = (bn.bias.detach().double() - mean * g * inv).float().contiguous()
            cur, ldx = Y, cout
        cout = cur.shape[1]
        out = torch.empty((cout, np_), dtype=torch.float32, device=dev)
        check(L.mvk_fa_reduce(ptr(cur), np_, k, cout, ptr(scale), ptr(shift), 0 if self.reduction == 'sum' else 1,
                              ptr(out), st))
        return out

    def _needs_grad(self, *tensors):
        return torch.is_grad_enabled() and (any(t is not None and t.requires_grad for t in tensors) or
                                            any(p.requires_grad for p in self.parameters()))

    def _run_autograd(self, X, np_, k):
        """Differentiable pipeline (late fusion trains this module, architectures_sphere_late_fusion.py:
        301-304): every layer is the fused Linear + batch norm + ReLU block of blocks.py
        (tcgen05 contraction, fused statistics / activation kernels, hand-written backward)."""
        from . import blocks as _b
        cur = X
        for layer in self.mlp:
            bn = layer.bn
            W = layer.conv.weight.reshape(layer.conv.out_channels, -1)
            momentum = bn.momentum if bn.momentum is not None else 0.1
            training = bn.training or not bn.track_running_stats
            nbt = bn.num_batches_tracked if (bn.training and bn.track_running_stats) else None
            cur = _b._LinearBNAct.apply(cur, W, bn.weight, bn.bias, None, bn.running_mean, bn.running_var, True,
                                        training, float(momentum), float(bn.eps), 0.0,
                                        getattr(self, "contraction", None) or _b._kp.DEFAULT_CONTRACTION, nbt)
        y = cur.view(np_, k, cur.shape[1])
        out = y.sum(1) if self.reduction == 'sum' else y.max(1)[0]
        return out.t()  # (cout, np)

    def forward(self, src_xyz, tgt_xyz, feature):
        """
        Args:
            src_xyz (torch.Tensor): (batch_size, 3, num_points, k)
            tgt_xyz (torch.Tensor): (batch_size, 3, num_points)
            feature (torch.Tensor): (batch_size, in_channels, num_points, k)
        Returns:
            torch.Tensor: (batch_size, out_channels, num_points)
        """
        _lib.require_cuda()
        if not feature.is_cuda:
            raise RuntimeError("FeatureAggregation: tensors must live on a CUDA device (no CPU fallback)")
        b, c, np_, k = feature.shape
        if self._needs_grad(feature, src_xyz, tgt_xyz):
            diff = (src_xyz - tgt_xyz.unsqueeze(-1)).float()
            dist = torch.sum(diff ** 2, dim=1, keepdim=True)
            x = torch.cat([feature.float(), diff, dist], dim=1)  # (b, c+4, np, k)
            X = x.permute(0, 2, 3, 1).reshape(b * np_ * k, c + 4).contiguous()
            out = self._run_autograd(X, b * np_, k)  # (cout, b*np)
            return out.reshape(-1, b, np_).permute(1, 0, 2).contiguous()
        with torch.no_grad(), torch.cuda.device(feature.device):
            diff = (src_xyz - tgt_xyz.unsqueeze(-1)).float()
            dist = torch.sum(diff ** 2, dim=1, keepdim=True)
            x = torch.cat([feature.float(), diff, dist], dim=1)  # (b, c+4, np, k)
            X = x.permute(0, 2, 3, 1).reshape(b * np_ * k, c + 4).contiguous()
            out = self._run(X, b * np_, k)  # (cout, b*np)
            return out.reshape(-1, b, np_).permute(1, 0, 2).contiguous()

    def forward_from_maps(self, feature_2d, image_xyz, knn_indices, tgt_points):
        """Fused entry point for the fusion nets' per-sphere loop (architectures_sphere.py:264-284):
        gathers the k nearest pixels' features and coordinates straight from the 2D maps.

        feature_2d  (C, npix)  channel-major (the reference's (b,64,nv*h*w) slice) or any strides
        image_xyz   (npix, 3)  float32 world coordinates of the pixels
        knn_indices (np, k)    int64 flat pixel ids
        tgt_points  (np, 3)    float32 sphere points
        returns     (C_out, np)
        """
        _lib.require_cuda()
        L = _lib.lib()
        c, npix = feature_2d.shape
        np_, k = knn_indices.shape
        dev = feature_2d.device
        if feature_2d.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("forward_from_maps treats the 2D feature map as a constant (the 2D network is frozen in "
                               "MV-KPConv, architectures_sphere.py:233-237); use group_points + forward() to "
                               "back-propagate into it")
        grad_path = self._needs_grad()
        with torch.no_grad(), torch.cuda.device(dev):
            f = feature_2d.detach().float()
            X = torch.empty((np_ * k, c + 4), dtype=torch.float32, device=dev)
            # converted copies must outlive the call: keep them in locals (a temporary freed between two
            # ptr() evaluations could hand its memory to the next temporary)
            xyz_c, knn_c, tgt_c = image_xyz.contiguous().float(), knn_indices.contiguous().long(), \
                tgt_points.contiguous().float()
            check(L.mvk_fa_gather(f.data_ptr(), f.stride(0), f.stride(1), c, ptr(xyz_c), ptr(knn_c), np_, k, ptr(tgt_c),
                                  ptr(X), c + 4, stream_ptr()))
            if not grad_path:
                return self._run(X, np_, k)
        return self._run_autograd(X, np_, k)


# -------------------------------------------------------------------------------------------------
def _kinv64(cam_matrix):
    """inv(cam_matrix[:3,:3]) exactly as the reference evaluates it on the host: the fp32 matrix is
    inverted in fp32 (np.linalg.inv keeps float32) and only then promoted by the fp64 product."""
    cam = np.asarray(cam_matrix)
    return np.linalg.inv(cam[:3, :3]).astype(np.float64)


def unproject_views(cam_matrix, depths, poses, return_f64=True):
    """All views of one sphere: depth (nv, h, w) -> world xyz and validity.

    Mirrors depth2xyz + `image_mask = z > 0` + `xyz @ pose[:3,:3].T + pose[:3,3]`
    (ScanNet_sphere_color.py:409-417).  Returns (xyz32 [nv,h,w,3] f32, mask [nv,h,w] bool,
    xyz64 [nv*h*w,3] f64 -- what the reference feeds to the kNN).
    """
    _lib.require_cuda()
    L = _lib.lib()
    as_np = not isinstance(depths, torch.Tensor)
    d = torch.as_tensor(np.ascontiguousarray(depths, dtype=np.float32)) if as_np else depths
    d = d.cuda().contiguous().float()
    nv, h, w = d.shape
    dev = d.device
    P = torch.as_tensor(np.ascontiguousarray(poses, dtype=np.float32)) if not isinstance(poses, torch.Tensor) else poses
    P = P.to(dev).contiguous().float().reshape(nv, 16)
    kinv = torch.from_numpy(np.ascontiguousarray(_kinv64(cam_matrix if not isinstance(cam_matrix, torch.Tensor)
                                                        else cam_matrix.cpu().numpy()))).to(dev)
    xyz64 = torch.empty((nv * h * w, 3), dtype=torch.float64, device=dev)
    xyz32 = torch.empty((nv, h, w, 3), dtype=torch.float32, device=dev)
    mask = torch.empty((nv, h, w), dtype=torch.uint8, device=dev)
    with _lib.on_device(dev):
        check(L.mvk_unproject_views(ptr(kinv), ptr(d), ptr(P), nv, h, w, ptr(xyz64), ptr(xyz32), ptr(mask),
                                    stream_ptr()))
    maskb = mask.bool()
    if as_np:
        return xyz32.cpu().numpy(), maskb.cpu().numpy(), xyz64.cpu().numpy()
    return xyz32, maskb, xyz64


def depth2xyz(cam_matrix, depth):
    """project depth map to 3D (camera frame), (h*w, 3) float64 like the reference
    (ScanNet_sphere_color.py:66-72)."""
    as_np = not isinstance(depth, torch.Tensor)
    d = np.asarray(depth) if as_np else depth
    eye = np.eye(4, dtype=np.float32)
    _, _, xyz64 = unproject_views(cam_matrix, d[None] if as_np else d.unsqueeze(0), eye[None])
    return xyz64


def knn_pixels(xyz64, mask, queries, k=3):
    """k nearest VALID pixels of every query point, as flat pixel ids (view*h*w + pix), sorted by
    (distance, id); fp64 distances like sklearn's ball tree on the float64 unprojected pixels
    (ScanNet_sphere_color.py:442-452).  Returns int64 (nq, k)."""
    _lib.require_cuda()
    L = _lib.lib()
    as_np = not isinstance(queries, torch.Tensor)
    keys = (torch.as_tensor(np.ascontiguousarray(xyz64, dtype=np.float64)) if not isinstance(xyz64, torch.Tensor)
            else xyz64).cuda().contiguous().double().reshape(-1, 3)
    dev = keys.device
    m = (torch.as_tensor(np.ascontiguousarray(mask)) if not isinstance(mask, torch.Tensor) else mask)
    m = m.to(dev).reshape(-1).to(torch.uint8).contiguous()
    q = (torch.as_tensor(np.ascontiguousarray(queries, dtype=np.float32)) if as_np else queries)
    q = q.to(dev).contiguous().float()
    npix, nq = keys.shape[0], q.shape[0]
    if int(m.sum().item()) < k:
        raise RuntimeError("knn_pixels: fewer than k valid pixels")
    out = torch.empty((nq, k), dtype=torch.int64, device=dev)
    with _lib.on_device(dev):
        wsb = L.mvk_knn_workspace_bytes(npix, nq)
        ws = _workspace(wsb, dev)
        check(L.mvk_knn_pixels(ptr(keys), ptr(m), npix, ptr(q), nq, k, ptr(ws), ws.numel(), ptr(out), stream_ptr()))
    return out.cpu().numpy() if as_np else out


# -------------------------------------------------------------------------------------------------
# Batched lifting: all spheres of a stacked batch at once (replaces the per-sphere loops of
# ScanNet_sphere_color.py:409-452 and architectures_sphere.py:246-279)
# -------------------------------------------------------------------------------------------------
def intrinsics_inverse(cam_matrices, B, nv, device):
    """[B*nv, 9] f64 device tensor of inv(cam[:3,:3]) per view: inverted on the host in fp32 like the reference (one
    3x3 per sphere, ScanNet_sphere_color.py:69), widened to fp64.  Reusable across steps for fixed intrinsics."""
    cam = cam_matrices.detach().cpu().numpy() if isinstance(cam_matrices, torch.Tensor) else np.asarray(cam_matrices)
    cam = np.broadcast_to(cam.reshape(-1, 4, 4) if cam.ndim == 3 else cam[None], (B, 4, 4))
    kinv = np.stack([_kinv64(c) for c in cam], 0)
    return torch.from_numpy(np.ascontiguousarray(np.repeat(kinv, nv, axis=0).reshape(B * nv, 9))).to(device)


def unproject_views_batched(cam_matrices, depths, poses, kinv=None):
    """Every view of every sphere of a batch in one launch.

    cam_matrices (B, 4, 4) or (4, 4) float32 (already rescaled to the depth resolution, :370-372),
    depths (B, nv, h, w) float32 metres, poses (B, nv, 4, 4) camera -> world.
    Returns (image_xyz [B, nv, h, w, 3] f32 -- the reference's batch.image_xyz --, mask [B, nv, h, w] bool,
    xyz64 [B*nv*h*w, 3] f64 -- what the reference feeds to the kNN)."""
    _lib.require_cuda()
    L = _lib.lib()
    d = depths if isinstance(depths, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(depths, dtype=np.float32))
    d = d.cuda().contiguous().float()
    B, nv, h, w = d.shape
    dev = d.device
    P = poses if isinstance(poses, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(poses, dtype=np.float32))
    P = P.to(dev).contiguous().float().reshape(B * nv, 16)
    if kinv is None:
        kinv = intrinsics_inverse(cam_matrices, B, nv, dev)
    xyz64 = torch.empty((B * nv * h * w, 3), dtype=torch.float64, device=dev)
    xyz32 = torch.empty((B, nv, h, w, 3), dtype=torch.float32, device=dev)
    mask = torch.empty((B, nv, h, w), dtype=torch.uint8, device=dev)
    with _lib.on_device(dev):
        check(L.mvk_unproject_views_batched(ptr(kinv), ptr(d), ptr(P), B * nv, h, w, ptr(xyz64), ptr(xyz32), ptr(mask),
                                            stream_ptr()))
    return xyz32, mask.bool(), xyz64


def knn_pixels_batched(xyz64, image_xyz, mask, queries, q_lengths, k=3, grid_dim=64, cell_min=0.02, global_ids=False,
                       return_far=False):
    """k nearest valid pixels of every stacked sphere point among the pixels of its own sphere's views.

    xyz64 [B*npix, 3] f64, image_xyz [B, nv, h, w, 3] f32, mask [B, nv, h, w] (all three from
    unproject_views_batched), queries [N, 3] f32 world-frame sphere points (the reference's feat_aggre_points),
    q_lengths [B] i32.  Returns int64 [N, k]: pixel ids local to the sphere (the reference's knn_list entries,
    ScanNet_sphere_color.py:452) or global (b * nv*h*w + id) with global_ids=True; identical to knn_pixels run
    sphere by sphere."""
    _lib.require_cuda()
    L = _lib.lib()
    dev = xyz64.device
    B = image_xyz.shape[0]
    npix = image_xyz.shape[1] * image_xyz.shape[2] * image_xyz.shape[3]
    x32 = image_xyz.reshape(-1, 3).contiguous().float()
    m = mask.reshape(-1).to(torch.uint8).contiguous()
    q = queries.to(dev).reshape(-1, 3).contiguous().float()
    ql = (q_lengths if isinstance(q_lengths, torch.Tensor) else torch.as_tensor(np.asarray(q_lengths, np.int32)))
    ql = ql.to(dev).to(torch.int32).contiguous()
    nq = q.shape[0]
    out = torch.empty((nq, k), dtype=torch.int64, device=dev)
    far = torch.zeros(B, dtype=torch.int32, device=dev) if return_far else None
    with _lib.on_device(dev):
        wsb = L.mvk_knn_batched_workspace_bytes(B, npix, nq, grid_dim)
        ws = _workspace(wsb, dev, "knn")
        check(L.mvk_knn_pixels_batched(ptr(xyz64), ptr(x32), ptr(m), B, npix, ptr(q), ptr(ql), nq, k, grid_dim,
                                       float(cell_min), 1 if global_ids else 0, ptr(ws), ws.numel(), ptr(out), ptr(far),
                                       stream_ptr()))
    return (out, far) if return_far else out


def _fa_forward_from_views(self, feature_2d, image_xyz, knn_global, tgt_points, differentiable=None):
    """FeatureAggregation over a whole stacked batch without materialising the grouped tensors.

    feature_2d  (V, C, h, w)   the 2D network's output for all V = B*nv views (any memory format; constant)
    image_xyz   (V*h*w, 3) or (B, nv, h, w, 3) float32 world coordinates of the pixels
    knn_global  (np, k)        int64 global pixel ids (view * h*w + pix)
    tgt_points  (np, 3)        float32 world-frame sphere points
    returns     (C_out, np)
    differentiable: None = gradients to the module's parameters when autograd would need them (late fusion);
    False = constant path (early / middle fusion detach the lifted features, architectures_sphere.py:288)."""
    _lib.require_cuda()
    L = _lib.lib()
    V, c, h, w = feature_2d.shape
    np_, k = knn_global.shape
    dev = feature_2d.device
    if feature_2d.requires_grad and torch.is_grad_enabled():
        raise RuntimeError("forward_from_views treats the 2D feature map as a constant (the 2D network is frozen in "
                           "MV-KPConv, architectures_sphere.py:233-237)")
    grad_path = self._needs_grad() if differentiable is None else bool(differentiable)
    with torch.no_grad(), torch.cuda.device(dev):
        f = feature_2d.detach()
        if f.dtype is not torch.float32:
            f = f.float()
        X = torch.empty((np_ * k, c + 4), dtype=torch.float32, device=dev)
        xyz_c = image_xyz.reshape(-1, 3).contiguous().float()
        knn_c, tgt_c = knn_global.contiguous().long(), tgt_points.reshape(-1, 3).contiguous().float()
        # pixel stride: the (h, w) plane must be addressable by one flat index
        if f.stride(2) != w * f.stride(3):
            f = f.contiguous()
        check(L.mvk_fa_gather_views(f.data_ptr(), f.stride(0), f.stride(1), f.stride(3), c, h * w, ptr(xyz_c), ptr(knn_c),
                                    np_, k, ptr(tgt_c), ptr(X), c + 4, stream_ptr()))
        if not grad_path:
            return self._run(X, np_, k)
    return self._run_autograd(X, np_, k)


FeatureAggregation.forward_from_views = _fa_forward_from_views
