"""Kernel-point dispositions for KPConv (host side).

Mirror of ``load_kernels`` (reference: KPConv-PyTorch/kernels/kernel_points.py:409-490): take a
unit disposition, add N(0, 0.01) noise, scale by the convolution radius and apply a random
z-rotation -- all drawn from ``np.random`` in the reference's order (theta first, then the noise),
so seeding ``np.random`` identically reproduces the reference's kernel points bit for bit.

The unit disposition below is DATA: the 15 x 3 float64 values stored in the reference fixture
``KPConv-PyTorch/kernels/dispositions/k_015_center_3D.ply`` (the output of the reference's
offline kernel-point optimiser, which is out of scope here).  Other (K, fixed) combinations would
need that optimiser and raise NotImplementedError.
"""
import numpy as np

K15_CENTER_3D = np.array([
    (0.0, 0.0, 0.0),
    (0.1765335189555964, -0.5331009227011366, 0.34867212306169687),
    (-0.17216714528968188, 0.5869864084878047, 0.2504762596313166),
    (0.5599921909412507, 0.14634608955911066, -0.3192561586124955),
    (0.004746390986670649, 0.05857523601904023, 0.6512933599145767),
    (-0.5718436199317224, 0.1568886457311281, 0.29208507261584277),
    (-0.13227391342522712, -0.5985939453557826, -0.24722815980598775),
    (0.1279073272593564, 0.5447058708120162, -0.3519489911759223),
    (-0.39749615389725035, -0.40315610071713537, 0.3411669684143527),
    (-0.004746384942927406, -0.058575242325823805, -0.6512933592303757),
    (0.4299012198976422, -0.42530640944638104, -0.2668938614262992),
    (-0.43426756965655483, 0.3714212162993105, -0.33225124809447437),
    (0.5762102053498649, -0.10300057047541586, 0.3070920780433485),
    (0.4018625024391343, 0.45704129434237256, 0.25797813593214997),
    (-0.5643585653833885, -0.20023157349675208, -0.27989221710862),
], dtype=np.float64)


def load_kernels(radius, num_kpoints, dimension, fixed, lloyd=False):
    if not (num_kpoints == 15 and dimension == 3 and fixed == "center"):
        raise NotImplementedError(
            "only the k_015_center_3D disposition is shipped (the reference's offline kernel point "
            "optimiser, kernel_points.py:80-406, is out of scope); got K=%r dim=%r fixed=%r"
            % (num_kpoints, dimension, fixed))
    kernel_points = K15_CENTER_3D.copy()
    # random z-rotation (kernel_points.py:453-464)
    theta = np.random.rand() * 2 * np.pi
    c, s = np.cos(theta), np.sin(theta)
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)
    # small noise, scale, rotate (:482-490)
    kernel_points = kernel_points + np.random.normal(scale=0.01, size=kernel_points.shape)
    kernel_points = radius * kernel_points
    kernel_points = np.matmul(kernel_points, R)
    return kernel_points.astype(np.float32)
