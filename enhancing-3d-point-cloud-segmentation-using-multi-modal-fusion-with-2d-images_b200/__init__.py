"""B200-native (sm_100a) implementation of the MV-KPConv data-parallel hot path.

Drop-in operator surface of the reference (dcy0577/Enhancing-3D-Point-Cloud-Segmentation-Using-
Multi-Modal-Fusion-with-2D-Images) for that path only:

    KPConv                     models/blocks.py:143-379         (rigid kernel point convolution)
    UnaryBlock, BatchNormBlock models/blocks.py:430-504         (+ bn_act: fused bn/residual/LeakyReLU)
    max_pool, closest_pool     models/blocks.py:79-110
    batch_neighbors            datasets/common.py:185-196       (cpp_wrappers/cpp_neighbors)
    grid_subsampling           datasets/common.py:44-74         (cpp_wrappers/cpp_subsampling)
    batch_grid_subsampling     datasets/common.py:77-182
    group_points               mvpnet/ops/group_points.py:5-31
    FeatureAggregation         mvpnet/models/mvpnet_3d.py:12-70
    depth2xyz/unproject_views  datasets/ScanNet_sphere_color.py:66-72, 409-417
    knn_pixels                 datasets/ScanNet_sphere_color.py:442-452

Every op runs hand-written CUDA kernels from libmvk.so (C ABI: include/mvk.h).  There is no CPU
fallback: without the library or without a CUDA device the ops raise.

The importable alias of this package is ``mvkpconv_b200`` (see mvkpconv_b200.py at the repo root):
the directory name mandated for the package is not a valid Python identifier.
"""
from . import _lib, build  # noqa: F401
from .geometry import batch_grid_subsampling, batch_neighbors, create_3D_rotations, grid_subsampling  # noqa: F401
from .kernel_points import load_kernels  # noqa: F401
from .blocks import BatchNormBlock, UnaryBlock, bn_act, softmax_cross_entropy  # noqa: F401
from .kpconv import KPConv, closest_pool, gather, max_pool  # noqa: F401
from .lifting import (FeatureAggregation, depth2xyz, group_points, knn_pixels, knn_pixels_batched,  # noqa: F401
                      unproject_views, unproject_views_batched)

__all__ = [
    "KPConv", "UnaryBlock", "BatchNormBlock", "bn_act", "softmax_cross_entropy", "max_pool", "closest_pool", "gather", "batch_neighbors", "grid_subsampling",
    "batch_grid_subsampling", "group_points", "FeatureAggregation", "depth2xyz", "unproject_views",
    "knn_pixels", "knn_pixels_batched", "unproject_views_batched", "load_kernels", "create_3D_rotations",
]
