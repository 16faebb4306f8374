"""MV-KPConv fusion networks: 2D features lifted onto the sphere points and fused early, in the middle or late.

Reference (KPConv-PyTorch/models): ``KPFCNN_featureAggre`` in

    architectures_sphere.py:62-295                  early fusion   x = [1, z, lifted 64-d]  (66 channels) -> KPFCNN
    architectures_sphere_middle_fusion.py:62-300    middle fusion  two encoders (3D: 4 ch, 2D: 65 ch), skip links
                                                    concatenated, bottleneck fused before the decoder
    architectures_sphere_late_fusion.py:62-310      late fusion    KPFCNN on 4 channels, decoder output 128 -> 64,
                                                    concatenated with the lifted features before the head

and the frozen 2D network ``UNetResNet34`` (mvpnet/models/unet_resnet34.py:9-125).

What differs from the reference, on purpose:

* the lifting is ONE launch set for the whole stacked batch (``FeatureAggregation.forward_from_views`` on global
  pixel ids) instead of a Python loop with two ``group_points`` calls per sphere (architectures_sphere.py:246-279);
  the indices can be the reference's own per-sphere ``batch.knn_list`` (CPU ball tree) or come from
  ``prepare_lifting`` (unprojection + grid 3-NN on the GPU, no host work);
* the 2D network runs channels-last, optionally under bf16 autocast (it is frozen and in eval mode in the
  reference: architectures_sphere.py:233-237); it is library code (cuDNN) either way;
* middle fusion: the reference averages the two bottlenecks (``torch.mean``, :129; its decoder bookkeeping resets
  the width to ``out_dim`` after the first upsampling block, so the ``in_dim_3d + in_dim_2d`` at :143 never reaches
  a layer) -- ``bottleneck='mean'`` (default).  ``bottleneck='cat'`` is the alternative the reference leaves
  commented out one line above, with a decoder sized for it (an extension: the reference's decoder would not fit).

Fork quirks kept: ``.clone().detach()`` in front of the encoders of early / middle fusion (the lifted
features carry no gradient there), LeakyReLU on the logits, ``transform_mlp`` / heads without batch norm.
"""
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import harness as _h
from . import lifting as _lift
from .lifting import FeatureAggregation


# -------------------------------------------------------------------------------------------------
class UNetResNet34(nn.Module):
    """UNet on a ResNet34 encoder (mvpnet/models/unet_resnet34.py:9-125): same submodule names, so the reference's
    2D checkpoints (``checkpoint['model']``, architectures_sphere.py:229-231) load with ``load_state_dict``.
    Returns {'seg_logit': (V, classes, h, w), 'feature': (V, 64, h, w)}.  Dense convolutions: cuDNN."""

    def __init__(self, num_classes, p=0.0, pretrained=False):
        super().__init__()
        from torchvision.models.resnet import resnet34
        self.num_classes = num_classes
        net = resnet34(weights="IMAGENET1K_V1" if pretrained else None)
        self.encoder0 = nn.Conv2d(3, 64, kernel_size=7, stride=1, padding=3, bias=False)  # conv1 without its stride
        self.encoder0.weight.data = net.conv1.weight.data
        self.bn, self.relu, self.maxpool = net.bn1, net.relu, net.maxpool
        self.encoder1, self.encoder2, self.encoder3, self.encoder4 = net.layer1, net.layer2, net.layer3, net.layer4
        self.deconv4, self.decoder3 = self._deconv(512, 256), self._conv(512, 256)
        self.deconv3, self.decoder2 = self._deconv(256, 128), self._conv(256, 128)
        self.deconv2, self.decoder1 = self._deconv(128, 64), self._conv(128, 64)
        self.deconv1, self.decoder0 = self._deconv(64, 64), self._conv(128, 64)
        self.logit = nn.Conv2d(64, num_classes, 1, bias=True)
        self.dropout = nn.Dropout(p=p) if p > 0.0 else None

    @staticmethod
    def _deconv(c_in, c_out):
        return nn.Sequential(nn.ConvTranspose2d(c_in, c_out, kernel_size=2, stride=2), nn.BatchNorm2d(c_out),
                             nn.ReLU(inplace=True))

    @staticmethod
    def _conv(c_in, c_out):
        return nn.Sequential(nn.Conv2d(c_in, c_out, kernel_size=3, padding=1), nn.BatchNorm2d(c_out), nn.ReLU(inplace=True))

    def forward(self, data_dict):
        x = data_dict['image']
        h, w = x.shape[2], x.shape[3]
        pad_h, pad_w = (-h) % 16, (-w) % 16  # :69-75
        if pad_h or pad_w:
            x = F.pad(x, [0, pad_w, 0, pad_h])
        feats = []
        x = self.relu(self.bn(self.encoder0(x)))
        feats.append(x)
        x = self.encoder1(self.maxpool(x))
        feats.append(x)
        x = self.encoder2(x)
        feats.append(x)
        x = self.encoder3(x)
        if self.dropout is not None:
            x = self.dropout(x)
        feats.append(x)
        x = self.encoder4(x)
        if self.dropout is not None:
            x = self.dropout(x)
        x = self.decoder3(torch.cat([self.deconv4(x), feats[3]], dim=1))
        x = self.decoder2(torch.cat([self.deconv3(x), feats[2]], dim=1))
        x = self.decoder1(torch.cat([self.deconv2(x), feats[1]], dim=1))
        x = self.decoder0(torch.cat([self.deconv1(x), feats[0]], dim=1))
        if pad_h or pad_w:
            x = x[:, :, 0:h, 0:w]
        return {'seg_logit': self.logit(x), 'feature': x}


def fold_batch_norm(net):
    """Inference copy of a frozen 2D network with every (Conv2d | ConvTranspose2d) -> BatchNorm2d pair folded into the
    convolution (eval-mode batch norm is an affine map per channel): same function up to rounding, ~44 fewer
    kernels per UNet-ResNet34 forward.  The original module (and its state dict / parameter names) is left untouched."""
    import copy
    from torch.nn.utils.fusion import fuse_conv_bn_eval
    m = copy.deepcopy(net).eval()

    def fold_children(parent):
        names = list(parent._modules.keys())
        for a, b in zip(names, names[1:]):
            conv, bn = parent._modules[a], parent._modules[b]
            if isinstance(conv, (nn.Conv2d, nn.ConvTranspose2d)) and isinstance(bn, nn.BatchNorm2d):
                parent._modules[a] = fuse_conv_bn_eval(conv, bn, transpose=isinstance(conv, nn.ConvTranspose2d))
                parent._modules[b] = nn.Identity()
        for child in parent._modules.values():
            if child is not None:
                fold_children(child)

    fold_children(m)
    # torchvision BasicBlock: conv1 -> bn1, conv2 -> bn2 are adjacent children as well (handled above); the UNet's stem
    # is encoder0 -> bn (adjacent in definition order)
    for p in m.parameters():
        p.requires_grad = False
    return m


# -------------------------------------------------------------------------------------------------
def prepare_lifting(cam_matrices, depths, poses, feat_aggre_points, lengths, k=3, kinv=None):
    """The numeric half of ``get_rgbd_data`` (ScanNet_sphere_color.py:409-452) for a whole batch on the GPU:
    returns (image_xyz [B, nv, h, w, 3] f32, image_mask [B, nv, h, w] bool, knn_global [N, k] int64).
    kinv: lifting.intrinsics_inverse(...) computed once for fixed intrinsics (no host work per step)."""
    xyz32, mask, xyz64 = _lift.unproject_views_batched(cam_matrices, depths, poses, kinv=kinv)
    knn = _lift.knn_pixels_batched(xyz64, xyz32, mask, feat_aggre_points.reshape(-1, 3), lengths, k=k, global_ids=True)
    return xyz32, mask, knn


def global_knn_from_list(knn_list, npix, device):
    """The reference's per-sphere ``batch.knn_list`` (local pixel ids, arrays of shape (1, n_i, k) or (n_i, k)) as
    one [N, k] tensor of global pixel ids: one host concatenation + one copy instead of a copy per sphere."""
    import numpy as np
    parts = [np.asarray(a).reshape(-1, np.asarray(a).shape[-1]).astype(np.int64) + i * npix for i, a in enumerate(knn_list)]
    return torch.from_numpy(np.concatenate(parts, 0)).to(device, non_blocking=True)


class FusionKPFCNN(nn.Module):
    """``KPFCNN_featureAggre`` for the three fusion variants, composed from this package's operators.

    batch fields (names of the reference's ``ScanNetCustomBatch``, ScanNet_sphere_color.py:1535-1620):
        points / neighbors / pools / upsamples / lengths   the pyramid
        images (B, nv, 3, h, w), image_xyz (B, nv, h, w, 3), feat_aggre_points (1, N, 3) or (N, 3), feature_3d (N, c3)
        knn_global [N, k]   (prepare_lifting)   or   knn_list (the reference's per-sphere arrays)
    config: the KPConv config + ``in_features_dim`` (early: 66, late: 4) or ``in_features_dim_3d`` /
    ``in_features_dim_2d`` (middle: 4 / 65).
    """

    def __init__(self, config, fusion="early", net_2d=None, num_classes=None, ops=None, bottleneck="mean",
                 precision_2d="fp32", fold_bn=True):
        super().__init__()
        if fusion not in ("early", "middle", "late"):
            raise ValueError("fusion must be 'early', 'middle' or 'late'")
        if bottleneck not in ("cat", "mean"):
            raise ValueError("bottleneck must be 'cat' or 'mean'")
        self.fusion, self.bottleneck, self.precision_2d = fusion, bottleneck, precision_2d
        self.fold_bn, self._folded = fold_bn, None
        self._product_ops = ops is None
        ops = ops or _h.product_ops()
        self.ops = ops
        arch = config.architecture
        self.C = num_classes or config.num_classes
        self.K = config.num_kernel_points
        r0 = config.first_subsampling_dl * config.conv_radius
        if fusion == "middle":
            in_dims = [config.in_features_dim_3d, config.in_features_dim_2d]
        else:
            in_dims = [config.in_features_dim]
        # ---- encoder(s): bookkeeping of architectures_sphere*.py (radius and width double at every strided block)
        encoders = [nn.ModuleList() for _ in in_dims]
        self.encoder_skips, skip_dims = [], []
        layer, r, out_dim = 0, r0, config.first_features_dim
        first_up = len(arch)
        for i, name in enumerate(arch):
            if any(t in name for t in ('pool', 'strided', 'upsample', 'global')):
                self.encoder_skips.append(i)
                skip_dims.append(sum(in_dims))
            if 'upsample' in name:
                first_up = i
                break
            for e, enc in enumerate(encoders):
                enc.append(_h._block(name, r, in_dims[e], out_dim, layer, config, ops))
            in_dims = [out_dim // 2 if 'simple' in name else out_dim for _ in in_dims]
            if 'pool' in name or 'strided' in name:
                layer, r, out_dim = layer + 1, r * 2, out_dim * 2
        if fusion == "middle":
            self.encoder_blocks_3d, self.encoder_blocks_2d = encoders
        else:
            self.encoder_blocks = encoders[0]
        for enc in encoders:  # the last encoder block feeds an upsampling, not a Linear
            if len(enc) and hasattr(enc[-1], "feeds_linear"):
                enc[-1].feeds_linear = False
        # ---- decoder
        in_dim = sum(in_dims) if (fusion == "middle" and bottleneck == "cat") else in_dims[0]  # actual width of x
        self.decoder_blocks, self.decoder_concats = nn.ModuleList(), []
        for j, name in enumerate(arch[first_up:]):
            if j > 0 and 'upsample' in arch[first_up + j - 1]:
                in_dim += skip_dims[layer]
                self.decoder_concats.append(j)
            self.decoder_blocks.append(_h._block(name, r, in_dim, out_dim, layer, config, ops))
            if 'upsample' in name:  # parameter-free: the width of x is unchanged
                layer, r, out_dim = layer - 1, r * 0.5, out_dim // 2
            else:
                in_dim = out_dim
        if fusion == "late":
            self.transform_mlp = _h._unary(ops, out_dim, 64, False, 0)  # 128 -> 64 (late_fusion.py:171)
        # late fusion: the head sees [transform_mlp output (64), lifted features (64)] -- 128 = out_dim in the reference's
        # configuration (first_features_dim = 128, late_fusion.py:172); sized from the actual width here
        self.head_mlp = _h._unary(ops, 128 if fusion == "late" else out_dim, config.first_features_dim, False, 0)
        self.head_softmax = _h._unary(ops, config.first_features_dim, self.C, False, 0)
        self.criterion = nn.CrossEntropyLoss(ignore_index=-1)
        self.deform_fitting_power = getattr(config, "deform_fitting_power", 1.0)
        self.repulse_extent = getattr(config, "repulse_extent", 1.2)
        self.l1 = nn.L1Loss()
        self._has_deformable = None
        # ---- lifting
        fa_cls = getattr(ops, "FeatureAggregation", FeatureAggregation)
        self.feat_aggreg = fa_cls(64)
        self.net_2d = net_2d if net_2d is not None else UNetResNet34(20, p=0.5, pretrained=False)
        for p in self.net_2d.parameters():  # frozen, children in eval mode (architectures_sphere.py:233-237)
            p.requires_grad = False
        for m in self.net_2d._modules.values():
            m.train(False)

    # ---- 2D network + lifting ---------------------------------------------------------------------
    def _net2d_for_inference(self):
        """The frozen 2D network, with its batch norms folded into the convolutions when it runs in eval mode on the
        GPU (product path); re-folded whenever a parameter or buffer of the original changed (checkpoint load)."""
        net = self.net_2d
        if not (self.fold_bn and self._product_ops) or any(m.training for m in net.modules() if isinstance(m, nn.BatchNorm2d)):
            return net
        key = tuple((t.data_ptr(), t._version) for t in list(net.parameters()) + list(net.buffers()))
        if self._folded is None or self._folded[0] != key:
            object.__setattr__(self, "_folded", (key, fold_batch_norm(net)))  # not a registered submodule
        return self._folded[1]

    def features_2d(self, images):
        b, nv = images.shape[:2]
        x = images.reshape((b * nv,) + tuple(images.shape[2:]))
        with torch.no_grad():
            if x.is_cuda:
                net = self._net2d_for_inference()
                x = x.contiguous(memory_format=torch.channels_last)
                if self.precision_2d == "bf16":
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        return net({'image': x})['feature'].float()
                return net({'image': x})['feature']
            return self.net_2d({'image': x})['feature']

    def lift(self, batch):
        """(N, 64) lifted 2D features of the stacked batch (architectures_sphere.py:242-284)."""
        images = batch.images
        b, nv, _, h, w = images.shape
        feature_2d = self.features_2d(images)  # (b*nv, 64, h, w)
        pts = batch.feat_aggre_points.reshape(-1, 3)
        differentiable = None if self.fusion == "late" else False  # early / middle detach the result (:288)
        if not self._product_ops:
            return self._lift_reference(feature_2d, batch, b, nv, h, w, pts)
        knn = getattr(batch, "knn_global", None)
        if knn is None:
            knn = global_knn_from_list(batch.knn_list, nv * h * w, feature_2d.device)
        out = self.feat_aggreg.forward_from_views(feature_2d, batch.image_xyz, knn, pts, differentiable=differentiable)
        return out.t()  # (N, 64); a view: consumers concatenate it anyway

    def _lift_reference(self, feature_2d, batch, b, nv, h, w, pts):
        """The reference's own sequence (per-sphere group_points + FeatureAggregation.forward) for the injected
        CPU operator set used by the parity tests."""
        f = feature_2d.reshape(b, nv, -1, h, w).transpose(1, 2).reshape(b, -1, nv * h * w)
        xyz = batch.image_xyz.permute(0, 4, 1, 2, 3).reshape(b, 3, nv * h * w)
        fl, xl = [], []
        for i in range(b):
            idx = torch.as_tensor(batch.knn_list[i]).long().reshape(1, -1, torch.as_tensor(batch.knn_list[i]).shape[-1])
            fl.append(self.ops.group_points(f[i:i + 1], idx))
            xl.append(self.ops.group_points(xyz[i:i + 1], idx))
        out = self.feat_aggreg(torch.cat(xl, dim=2), pts.t().unsqueeze(0), torch.cat(fl, dim=2))
        return out.permute(0, 2, 1).reshape(-1, out.shape[1])

    # ---- network ----------------------------------------------------------------------------------
    def _encode(self, blocks, x, batch, skips=None, concat_into=None):
        k = 0
        for i, op in enumerate(blocks):
            if i in self.encoder_skips:
                if concat_into is not None:
                    concat_into[k] = torch.cat([concat_into[k], x], dim=1)  # middle fusion: cat the skip features
                    k += 1
                else:
                    skips.append(x)
            x = op(x, batch)
        return x

    def _decode(self, x, skips, batch):
        j, blocks = 0, self.decoder_blocks
        while j < len(blocks):
            op = blocks[j]
            if (self._product_ops and isinstance(op, _h.NearestUpsampleBlock) and j + 1 < len(blocks) and
                    (j + 1) in self.decoder_concats and hasattr(blocks[j + 1], "forward_upsampled")):
                x = blocks[j + 1].forward_upsampled(x, batch.upsamples[op.layer_ind - 1], skips.pop(),
                                                    emit_hilo=(j + 2 == len(blocks)))
                j += 2
                continue
            if j in self.decoder_concats:
                x = torch.cat([x, skips.pop()], dim=1)
            x = op(x, batch)
            j += 1
        return x

    def forward(self, batch, config=None):
        feature_2d3d = self.lift(batch)  # (N, 64)
        f3d = batch.feature_3d
        skips = []
        if self.fusion == "early":
            x = torch.cat((f3d, feature_2d3d), dim=1).clone().detach()  # (N, 66) = 1 + z + 64  (:285-288)
            x = self._encode(self.encoder_blocks, x, batch, skips)
        elif self.fusion == "late":
            x = self._encode(self.encoder_blocks, f3d.clone().detach(), batch, skips)
        else:
            ones = torch.ones_like(f3d[:, :1])
            x2 = torch.cat((ones, feature_2d3d), dim=1).clone().detach()  # (N, 65)  (middle_fusion.py:97-107)
            x3 = self._encode(self.encoder_blocks_3d, f3d.clone().detach(), batch, skips)
            x2 = self._encode(self.encoder_blocks_2d, x2, batch, concat_into=skips)
            x = torch.cat([x3, x2], dim=1) if self.bottleneck == "cat" else torch.mean(torch.stack([x3, x2]), 0)
        x = self._decode(x, skips, batch)
        if self.fusion == "late":
            x = self.transform_mlp(x, batch)                      # (N, 64)
            x = torch.cat((x, feature_2d3d), dim=1)               # (N, 128)  (late_fusion.py:301-304)
        x = self.head_mlp(x, batch)
        return self.head_softmax(x, batch)

    def kpconv_layers(self):
        nested = {id(m.offset_conv) for m in self.modules() if getattr(m, "offset_conv", None) is not None}
        return [m for m in self.modules() if type(m).__name__.startswith("KPConv") and id(m) not in nested]

    loss = _h.KPFCNN.loss
    fitting_regularizer = _h.KPFCNN.fitting_regularizer

    def train(self, mode=True):
        """The reference flips the 2D network back to train mode after every validation (utils/trainer.py:266,
        SURVEY A.5); here it stays frozen in eval mode, which is what its construction intends."""
        super().train(mode)
        for m in self.net_2d._modules.values():
            m.train(False)
        return self


def fusion_config(fusion, **overrides):
    """Hot-path fields of the reference's fusion training configs: train_ScanNet_sphere.py:126-200 (early, rigid
    deeper, 66 input channels), train_ScanNet_sphere_middle_fusion.py:84-135 (4 + 65) and
    train_ScanNet_sphere_late_fusion.py:88-195 (4); the middle / late scripts use the deformable architecture."""
    from . import pyramid
    deform_arch = ['simple', 'resnetb', 'resnetb_strided', 'resnetb', 'resnetb_strided', 'resnetb_deformable',
                   'resnetb_deformable_strided', 'resnetb_deformable', 'resnetb_deformable_strided',
                   'resnetb_deformable', 'nearest_upsample', 'unary', 'nearest_upsample', 'unary', 'nearest_upsample',
                   'unary', 'nearest_upsample', 'unary']
    if fusion == "early":
        cfg = pyramid.baseline_config(in_features_dim=66)
    elif fusion == "middle":
        cfg = pyramid.baseline_config(architecture=deform_arch, in_features_dim_3d=4, in_features_dim_2d=65)
    else:
        cfg = pyramid.baseline_config(architecture=deform_arch, in_features_dim=4)
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg
