"""fp32 torch / numpy restatements of the floating-point half of the hot path.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- CPU, autograd-capable, slow and obvious.

Each function cites the reference lines it follows.  Pinned by tests/test_oracle.py against the
golden vectors in tests/golden/ that tests/golden/make_golden.py produced by importing and
running the reference's own Python (`models/blocks.py`, `mvpnet/FeatureAggregation_dummy_test.py`,
`datasets/ScanNet_sphere_color.py:depth2xyz`, sklearn ball-tree kNN).
"""
import numpy as np
import torch
import torch.nn.functional as F


# -------------------------------------------------------------------------------------------------
# KPConv  (KPConv-PyTorch/models/blocks.py:237-374, rigid branch)
# -------------------------------------------------------------------------------------------------
def kpconv_forward(q_pts, s_pts, neighb_inds, x, kernel_points, weights, KP_extent,
                   KP_influence="linear", aggregation_mode="sum"):
    """out[i] = sum_k ( sum_h w_ihk * x[j_ih] ) @ W[k]   with j = neighb_inds (shadow = Ns)."""
    # blocks.py:277  shadow support point at 1e6
    s_pad = torch.cat((s_pts, torch.zeros_like(s_pts[:1, :]) + 1e6), 0)
    # :280-283  neighbours centred on the query
    neighbors = s_pad[neighb_inds, :] - q_pts.unsqueeze(1)
    # :293-297  squared distance to every kernel point  [N, H, K]
    differences = neighbors.unsqueeze(2) - kernel_points
    sq_distances = torch.sum(differences ** 2, dim=3)
    # :329-346  influence
    if KP_influence == "constant":
        all_weights = torch.ones_like(sq_distances)
    elif KP_influence == "linear":
        all_weights = torch.clamp(1 - torch.sqrt(sq_distances) / KP_extent, min=0.0)
    elif KP_influence == "gaussian":
        sigma = KP_extent * 0.3
        all_weights = torch.exp(-sq_distances / (2 * sigma ** 2 + 1e-9))  # blocks.py:68-75
    else:
        raise ValueError("Unknown influence function type (config.KP_influence)")
    all_weights = torch.transpose(all_weights, 1, 2)  # [N, K, H]
    # :349-354
    if aggregation_mode == "closest":
        nn1 = torch.argmin(sq_distances, dim=2)
        all_weights = all_weights * torch.transpose(F.one_hot(nn1, kernel_points.shape[0]), 1, 2).float()
    elif aggregation_mode != "sum":
        raise ValueError("Unknown convolution mode. Should be 'closest' or 'sum'")
    # :357-360  zero feature row for shadow neighbours, gather
    x_pad = torch.cat((x, torch.zeros_like(x[:1, :])), 0)
    neighb_x = x_pad[neighb_inds]  # [N, H, Cin]
    # :363
    weighted = torch.matmul(all_weights, neighb_x)  # [N, K, Cin]
    # :370-374
    kernel_outputs = torch.matmul(weighted.permute(1, 0, 2), weights)  # [K, N, Cout]
    return torch.sum(kernel_outputs, dim=0)


def kpconv_deform_forward(q_pts, s_pts, neighb_inds, x, kernel_points, weights, KP_extent, offset_features,
                          modulated=False, KP_influence="linear", aggregation_mode="sum"):
    """Deformable branch (blocks.py:243-374).  `offset_features` [N, K*3 (+K)] is the output of the
    offset convolution plus bias.  Returns (out, min_d2 [N, K], deformed_KP [N, K, 3]).
    The reference compacts the neighbours that are in range of some deformed kernel point with a
    top-k (:300-325); masking the weights of the others is the same computation."""
    K = kernel_points.shape[0]
    if modulated:
        unscaled = offset_features[:, :3 * K].reshape(-1, K, 3)
        modulations = 2 * torch.sigmoid(offset_features[:, 3 * K:])  # :252-257
    else:
        unscaled = offset_features.view(-1, K, 3)
        modulations = None
    deformed_KP = unscaled * KP_extent + kernel_points  # :266, :287
    s_pad = torch.cat((s_pts, torch.zeros_like(s_pts[:1, :]) + 1e6), 0)
    neighbors = s_pad[neighb_inds, :] - q_pts.unsqueeze(1)
    differences = neighbors.unsqueeze(2) - deformed_KP.unsqueeze(1)  # [N, H, K, 3]
    sq_distances = torch.sum(differences ** 2, dim=3)
    min_d2, _ = torch.min(sq_distances, dim=1)  # :303
    in_range = torch.any(sq_distances < KP_extent ** 2, dim=2)  # :306  [N, H]
    if KP_influence == "constant":
        all_weights = torch.ones_like(sq_distances)
    elif KP_influence == "linear":
        all_weights = torch.clamp(1 - torch.sqrt(sq_distances) / KP_extent, min=0.0)
    elif KP_influence == "gaussian":
        sigma = KP_extent * 0.3
        all_weights = torch.exp(-sq_distances / (2 * sigma ** 2 + 1e-9))
    else:
        raise ValueError("Unknown influence function type (config.KP_influence)")
    if aggregation_mode == "closest":
        nn1 = torch.argmin(sq_distances, dim=2)
        all_weights = all_weights * F.one_hot(nn1, K).float()
    all_weights = all_weights * in_range.unsqueeze(2).float()  # dropped neighbours become shadows (:319-323)
    all_weights = torch.transpose(all_weights, 1, 2)
    x_pad = torch.cat((x, torch.zeros_like(x[:1, :])), 0)
    neighb_x = x_pad[neighb_inds]
    weighted = torch.matmul(all_weights, neighb_x)
    if modulations is not None:
        weighted = weighted * modulations.unsqueeze(2)  # :366-367
    kernel_outputs = torch.matmul(weighted.permute(1, 0, 2), weights)
    return torch.sum(kernel_outputs, dim=0), min_d2, deformed_KP


def max_pool(x, inds):
    """blocks.py:93-110 -- NB pads with ZEROS (not -inf)."""
    x_pad = torch.cat((x, torch.zeros_like(x[:1, :])), 0)
    return torch.max(x_pad[inds], 1)[0]


def closest_pool(x, inds):
    """blocks.py:79-90."""
    x_pad = torch.cat((x, torch.zeros_like(x[:1, :])), 0)
    return x_pad[inds[:, 0]]


# -------------------------------------------------------------------------------------------------
# lifting  (mvpnet/ops/group_points.py:5-31, mvpnet/models/mvpnet_3d.py:40-64)
# -------------------------------------------------------------------------------------------------
def group_points(points, index):
    """out[b,c,n,k] = points[b,c,index[b,n,k]]  (group_points_kernel.cu:25-47;
    oracle of the reference's own test: mvpnet/ops/tests/test_group_points.py:6-12)."""
    b, c, n1 = points.shape
    _, n2, k = index.shape
    idx = index.unsqueeze(1).expand(b, c, n2, k)
    return points.unsqueeze(2).expand(b, c, n2, n1).gather(3, idx)


def feature_aggregation_forward(src_xyz, tgt_xyz, feature, conv_weights, bn_weight, bn_bias,
                                bn_mean, bn_var, training, reduction="sum", eps=1e-5,
                                use_relation=True):
    """mvpnet_3d.py:40-64 with SharedMLP(ndim=2) = [1x1 Conv2d(no bias) + BatchNorm2d + ReLU]*L
    (common/nn/modules/mlp.py:38-75, conv.py:29-51).
    src_xyz (b,3,np,k), tgt_xyz (b,3,np), feature (b,C,np,k) -> (b,Cout,np).
    training=True uses batch statistics (biased variance) like nn.BatchNorm2d.train()."""
    if use_relation:
        diff = src_xyz - tgt_xyz.unsqueeze(-1)
        dist = torch.sum(diff ** 2, dim=1, keepdim=True)
        x = torch.cat([feature, diff, dist], dim=1)
    else:
        x = feature
    for li, w in enumerate(conv_weights):
        x = torch.einsum("oc,bcnk->bonk", w.reshape(w.shape[0], -1), x)
        if training:
            mean = x.mean(dim=(0, 2, 3))
            var = x.var(dim=(0, 2, 3), unbiased=False)
        else:
            mean, var = bn_mean[li], bn_var[li]
        x = (x - mean[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + eps)
        x = x * bn_weight[li][None, :, None, None] + bn_bias[li][None, :, None, None]
        x = torch.relu(x)
    if reduction == "sum":
        return torch.sum(x, 3)
    return torch.max(x, 3)[0]


def depth2xyz(cam_matrix, depth):
    """ScanNet_sphere_color.py:66-72.  int64 pixel grid x fp32 K^-1 promotes to float64."""
    v, u = np.indices(depth.shape)
    u, v = u.ravel(), v.ravel()
    uv1 = np.stack([u, v, np.ones_like(u)], axis=1)
    return (np.linalg.inv(cam_matrix[:3, :3]).dot(uv1.T) * depth.ravel()).T


def unproject_view(cam_matrix, depth, pose):
    """ScanNet_sphere_color.py:409-417: camera-frame xyz, validity (z>0), camera->world."""
    xyz = depth2xyz(cam_matrix, depth)
    mask = xyz[:, 2] > 0
    xyz = np.matmul(xyz, pose[:3, :3].T) + pose[:3, 3]
    return xyz, mask


def knn_pixels(image_xyz_list, image_mask_list, queries, k=3):
    """ScanNet_sphere_color.py:427-452: k nearest VALID pixels (fp64 euclidean, ties by lower
    flat pixel id) of every query, returned as flat pixel ids view*h*w + pix.
    Brute force restatement of sklearn NearestNeighbors(algorithm='ball_tree').kneighbors."""
    keys, ids = [], []
    off = 0
    for xyz, m in zip(image_xyz_list, image_mask_list):
        xyz = np.asarray(xyz, dtype=np.float64).reshape(-1, 3)
        m = np.asarray(m).reshape(-1)
        sel = np.nonzero(m)[0]
        keys.append(xyz[sel])
        ids.append(sel + off)
        off += len(m)
    keys = np.concatenate(keys, 0)
    ids = np.concatenate(ids, 0)
    q = np.asarray(queries, dtype=np.float64)
    out = np.zeros((len(q), k), dtype=np.int64)
    for i0 in range(0, len(q), 256):
        d = ((q[i0:i0 + 256, None, :] - keys[None, :, :]) ** 2).sum(-1)
        out[i0:i0 + 256] = ids[np.argsort(d, axis=1, kind="stable")[:, :k]]
    return out


# -------------------------------------------------------------------------------------------------
# nn.Module wrapper so the harness network can run entirely on the CPU oracle (bench.py's CPU
# baseline leg).  Same constructor surface as the reference module (blocks.py:145-147).
# -------------------------------------------------------------------------------------------------
class KPConvOracle(torch.nn.Module):
    def __init__(self, kernel_size, p_dim, in_channels, out_channels, KP_extent, radius,
                 fixed_kernel_points='center', KP_influence='linear', aggregation_mode='sum',
                 deformable=False, modulated=False):
        super().__init__()
        import math
        from mvkpconv_b200.kernel_points import load_kernels  # host-side data + RNG mirror only
        self.K, self.KP_extent, self.radius = kernel_size, KP_extent, radius
        self.KP_influence, self.aggregation_mode = KP_influence, aggregation_mode
        self.deformable, self.modulated = deformable, modulated
        self.min_d2 = self.deformed_KP = None
        self.weights = torch.nn.Parameter(torch.zeros((kernel_size, in_channels, out_channels)))
        if deformable:  # construction order of blocks.py:186-213 (RNG streams)
            self.offset_dim = (p_dim + 1) * kernel_size if modulated else p_dim * kernel_size
            self.offset_conv = KPConvOracle(kernel_size, p_dim, in_channels, self.offset_dim, KP_extent, radius,
                                            fixed_kernel_points, KP_influence, aggregation_mode)
            self.offset_bias = torch.nn.Parameter(torch.zeros(self.offset_dim))
        torch.nn.init.kaiming_uniform_(self.weights, a=math.sqrt(5))
        self.kernel_points = torch.nn.Parameter(
            torch.tensor(load_kernels(radius, kernel_size, p_dim, fixed_kernel_points)), requires_grad=False)

    def forward(self, q_pts, s_pts, neighb_inds, x):
        if self.deformable:
            of = self.offset_conv(q_pts, s_pts, neighb_inds, x) + self.offset_bias
            out, self.min_d2, self.deformed_KP = kpconv_deform_forward(
                q_pts, s_pts, neighb_inds, x, self.kernel_points, self.weights, self.KP_extent, of, self.modulated,
                self.KP_influence, self.aggregation_mode)
            return out
        return kpconv_forward(q_pts, s_pts, neighb_inds, x, self.kernel_points, self.weights, self.KP_extent,
                              self.KP_influence, self.aggregation_mode)


class FeatureAggregationOracle(torch.nn.Module):
    """mvpnet/models/mvpnet_3d.py:12-70 with the reference's module structure and parameter names
    (mlp.{i}.conv.weight, mlp.{i}.bn.*): SharedMLP(ndim=2) = [Conv2d 1x1 no bias, BatchNorm2d, ReLU] x 3
    (common/nn/modules/mlp.py:38-75, conv.py:29-51), xavier-uniform conv weights (mvpnet_3d.py:66-70)."""

    class _Layer(torch.nn.Module):
        def __init__(self, cin, cout):
            super().__init__()
            self.conv = torch.nn.Conv2d(cin, cout, 1, bias=False)
            self.bn = torch.nn.BatchNorm2d(cout)

        def forward(self, x):
            return torch.relu(self.bn(self.conv(x)))

    def __init__(self, in_channels, mlp_channels=(64, 64, 64), reduction="sum", use_relation=True):
        super().__init__()
        assert use_relation
        self.reduction = reduction
        self.mlp = torch.nn.ModuleList()
        c = in_channels + 4
        for co in mlp_channels:
            self.mlp.append(self._Layer(c, co))
            c = co
        for m in self.modules():
            if isinstance(m, torch.nn.Conv2d):
                torch.nn.init.xavier_uniform_(m.weight)

    def forward(self, src_xyz, tgt_xyz, feature):
        diff = src_xyz - tgt_xyz.unsqueeze(-1)
        x = torch.cat([feature, diff, torch.sum(diff ** 2, dim=1, keepdim=True)], dim=1)
        for layer in self.mlp:
            x = layer(x)
        return torch.sum(x, 3) if self.reduction == "sum" else torch.max(x, 3)[0]
