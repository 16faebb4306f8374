"""Seeded synthetic ScanNet-shaped inputs (host side, numpy only).

There is no dataset on the GPU box: benchmarks and parity tests use a synthetic room -- floor,
ceiling, four walls and a few axis-aligned boxes (furniture) -- surface-sampled with 3 mm noise,
from which spheres of radius `in_radius` are cropped.  The raw density is
chosen so that a r = 2 m sphere holds roughly 20-25 k points after the first_subsampling_dl = 0.04
grid subsampling, like the reference's input pipeline produces.

Nothing here touches the GPU; the subsampling step is injected (`subsample=`) so the product
passes its CUDA `grid_subsampling` and the CPU baseline passes the reference's.
"""
import numpy as np

ROOM = (6.0, 5.0, 2.6)


def _sample_rect(rng, n, origin, u, v):
    a = rng.random((n, 1))
    b = rng.random((n, 1))
    return origin + a * u + b * v


def make_room(seed=0, density=9000.0, size=ROOM, n_boxes=8, noise=0.003):
    """Surface samples of a furnished room, float32 (N, 3); `density` = raw points per m^2."""
    rng = np.random.default_rng(seed)
    X, Y, Z = size
    rects = [
        ((0, 0, 0), (X, 0, 0), (0, Y, 0)), ((0, 0, Z), (X, 0, 0), (0, Y, 0)),
        ((0, 0, 0), (X, 0, 0), (0, 0, Z)), ((0, Y, 0), (X, 0, 0), (0, 0, Z)),
        ((0, 0, 0), (0, Y, 0), (0, 0, Z)), ((X, 0, 0), (0, Y, 0), (0, 0, Z)),
    ]
    for _ in range(n_boxes):
        sx, sy, sz = rng.uniform(0.4, 1.6), rng.uniform(0.4, 1.2), rng.uniform(0.4, 1.5)
        ox, oy = rng.uniform(0.1, X - sx - 0.1), rng.uniform(0.1, Y - sy - 0.1)
        o = np.array([ox, oy, 0.0])
        rects += [
            (o + (0, 0, sz), (sx, 0, 0), (0, sy, 0)),
            (o, (sx, 0, 0), (0, 0, sz)), (o + (0, sy, 0), (sx, 0, 0), (0, 0, sz)),
            (o, (0, sy, 0), (0, 0, sz)), (o + (sx, 0, 0), (0, sy, 0), (0, 0, sz)),
        ]
    pts = []
    for o, u, v in rects:
        o, u, v = np.asarray(o, float), np.asarray(u, float), np.asarray(v, float)
        area = np.linalg.norm(np.cross(u, v))
        pts.append(_sample_rect(rng, max(1, int(area * density)), o, u, v))
    pts = np.concatenate(pts, 0)
    pts += rng.normal(0.0, noise, pts.shape)
    return pts.astype(np.float32)


def crop_sphere(points, center, radius):
    d2 = ((points - np.asarray(center, np.float32)) ** 2).sum(1)
    sel = points[d2 < radius * radius]
    return (sel - np.asarray(center, np.float32)).astype(np.float32)


def make_spheres(n_spheres, subsample, seed=0, in_radius=2.0, first_dl=0.04, density=9000.0):
    """`n_spheres` re-centred spheres (list of (Ni,3) float32) cut from rooms with different seeds.
    `subsample(points, dl)` -> points performs the first grid subsampling."""
    rng = np.random.default_rng(seed + 1000)
    spheres = []
    for i in range(n_spheres):
        room = make_room(seed=seed + i, density=density)
        sub = np.asarray(subsample(room, first_dl), dtype=np.float32)
        c = np.array([rng.uniform(1.5, ROOM[0] - 1.5), rng.uniform(1.5, ROOM[1] - 1.5), 1.0], np.float32)
        spheres.append(crop_sphere(sub, c, in_radius))
    return spheres


def stack(spheres):
    pts = np.concatenate(spheres, 0).astype(np.float32)
    lens = np.array([len(s) for s in spheres], dtype=np.int32)
    return pts, lens


def make_views(points_world, n_views=3, h=120, w=160, seed=0, hole_frac=0.05):
    """Synthetic RGB-D views of a (world-frame) cloud: poses looking at the cloud centre, depth by
    z-buffer splatting.  Returns cam_matrix (4,4) f32 (ScanNet depth intrinsics scaled to (w, h)),
    depths (nv,h,w) f32 in metres (0 = invalid), poses (nv,4,4) f32 camera->world."""
    rng = np.random.default_rng(seed + 77)
    cam = np.eye(4, dtype=np.float32)
    cam[0, 0] = cam[1, 1] = 577.870605
    cam[0, 2], cam[1, 2] = 319.5, 239.5
    cam[0] /= 640.0 / w
    cam[1] /= 480.0 / h
    centre = points_world.mean(0)
    depths, poses = [], []
    for _ in range(n_views):
        ang = rng.uniform(0, 2 * np.pi)
        eye = centre + np.array([1.6 * np.cos(ang), 1.6 * np.sin(ang), rng.uniform(0.2, 0.8)])
        fwd = centre - eye
        fwd /= np.linalg.norm(fwd)
        right = np.cross(fwd, [0, 0, 1.0])
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack([right, down, fwd], 1)  # camera axes (x right, y down, z forward) in world
        pose = np.eye(4, dtype=np.float32)
        pose[:3, :3] = R
        pose[:3, 3] = eye
        pc = (points_world - eye) @ R
        z = pc[:, 2]
        ok = z > 0.2
        u = np.round(pc[ok, 0] / z[ok] * cam[0, 0] + cam[0, 2]).astype(np.int64)
        v = np.round(pc[ok, 1] / z[ok] * cam[1, 1] + cam[1, 2]).astype(np.int64)
        zz = z[ok]
        inb = (u >= 0) & (u < w) & (v >= 0) & (v < h)
        depth = np.full(h * w, np.inf)
        np.minimum.at(depth, v[inb] * w + u[inb], zz[inb])
        depth[~np.isfinite(depth)] = 0.0
        depth[rng.random(h * w) < hole_frac] = 0.0
        depth = np.round(depth.reshape(h, w) * 1000.0) / 1000.0  # millimetre PNG quantisation
        depths.append(depth.astype(np.float32))
        poses.append(pose)
    return cam, np.stack(depths), np.stack(poses)


def make_fusion_spheres(n_spheres, subsample, seed=0, in_radius=2.0, first_dl=0.04, density=9000.0, n_views=3, h=120, w=160):
    """Spheres with their RGB-D views (BASELINE configs[2-3]): like make_spheres, plus for every
    sphere the world-frame points (the reference's feat_aggre_points), a depth-intrinsics matrix scaled to (w, h),
    n_views depth maps rendered from the sphere's own points and the camera poses.  Returns a list of namespaces
    (points centred [n,3], world [n,3], cam [4,4], depths [nv,h,w], poses [nv,4,4])."""
    from types import SimpleNamespace
    rng = np.random.default_rng(seed + 1000)
    out = []
    for i in range(n_spheres):
        room = make_room(seed=seed + i, density=density)
        sub = np.asarray(subsample(room, first_dl), dtype=np.float32)
        c = np.array([rng.uniform(1.5, ROOM[0] - 1.5), rng.uniform(1.5, ROOM[1] - 1.5), 1.0], np.float32)
        centred = crop_sphere(sub, c, in_radius)
        world = (centred + c).astype(np.float32)
        cam, depths, poses = make_views(world, n_views=n_views, h=h, w=w, seed=seed + i)
        out.append(SimpleNamespace(points=centred, world=world, cam=cam, depths=depths, poses=poses))
    return out
