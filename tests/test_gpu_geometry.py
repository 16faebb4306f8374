"""GPU parity: radius neighbours and grid subsampling, bit-exact against the oracle / goldens."""
import numpy as np
import pytest
import torch

from conftest import bumpy_cloud, canonical_rows, load_golden
from oracle import geom

pytestmark = pytest.mark.gpu


def test_neighbors_vs_reference_golden(mvk):
    g = load_golden("geometry")
    pts, lens, sp, sl = g["points"], g["lengths"], g["sub_pts"], g["sub_len"]
    for q, s, ql, sl_, rad, key in [(pts, pts, lens, lens, 0.15, "conv"), (sp, pts, sl, lens, 0.15, "pool"),
                                    (pts, sp, lens, sl, 0.3, "up")]:
        out = mvk.batch_neighbors(q, s, ql, sl_, rad)
        assert out.dtype == np.int32
        assert np.array_equal(out, g[key + "_ordered"]), key  # == reference batch_ordered_neighbors, bit exact
        assert np.array_equal(canonical_rows(out, s, q, len(s)), canonical_rows(g[key], s, q, len(s)))  # nanoflann


@pytest.mark.parametrize("n,nb,radius", [(64, 1, 0.3), (3000, 3, 0.1), (20000, 4, 0.06), (50000, 8, 0.05)])
def test_neighbors_vs_oracle_seeded(mvk, n, nb, radius):
    rng = np.random.default_rng(n)
    pts = bumpy_cloud(rng, n)
    cuts = np.sort(rng.choice(np.arange(1, n), nb - 1, replace=False)) if nb > 1 else np.array([], int)
    lens = np.diff(np.concatenate([[0], cuts, [n]])).astype(np.int32)
    ref, cnt = geom.batch_neighbors(pts, pts, lens, lens, radius, return_counts=True)
    out, cnt_gpu = mvk.batch_neighbors(pts, pts, lens, lens, radius, return_counts=True)
    assert np.array_equal(out, ref)
    assert np.array_equal(cnt_gpu, cnt)
    # device tensors in -> device tensor out, int64 variant, cropped rows keep the nearest
    tp = torch.from_numpy(pts).cuda()
    tl = torch.from_numpy(lens).cuda()
    out64 = mvk.batch_neighbors(tp, tp, tl, tl, radius, out_dtype=torch.int64, max_neighbors=7)
    assert out64.is_cuda and out64.dtype == torch.int64
    assert np.array_equal(out64.cpu().numpy(), ref[:, :7].astype(np.int64))


def test_neighbors_edge_cases(mvk):
    rng = np.random.default_rng(0)
    # duplicates (exact d2 ties), points on cell borders, far apart elements, a query cloud != support cloud
    base = rng.uniform(-1, 1, (300, 3)).astype(np.float32)
    pts = np.concatenate([base, base[:50], np.round(base[:100] * 10) / 10]).astype(np.float32)
    lens = np.array([len(pts)], np.int32)
    assert np.array_equal(mvk.batch_neighbors(pts, pts, lens, lens, 0.2), geom.batch_neighbors(pts, pts, lens, lens, 0.2))
    q = rng.uniform(-1.2, 1.2, (77, 3)).astype(np.float32)
    ql = np.array([40, 37], np.int32)
    sl = np.array([200, 250], np.int32)
    assert np.array_equal(mvk.batch_neighbors(q, pts, ql, sl, 0.35), geom.batch_neighbors(q, pts, ql, sl, 0.35))
    # large coordinates (grid bias) and a radius that makes every support a neighbour (wide rows)
    far = (pts[:200] + np.float32(500.0)).astype(np.float32)
    l2 = np.array([200], np.int32)
    out = mvk.batch_neighbors(far, far, l2, l2, 10.0)
    assert out.shape == (200, 200) and np.array_equal(out, geom.batch_neighbors(far, far, l2, l2, 10.0))
    # empty result raises like the reference glue (wrapper.cpp:201-205)
    lonely_q = np.array([[100.0, 0, 0]], np.float32)
    with pytest.raises(RuntimeError, match="Error"):
        mvk.batch_neighbors(lonely_q, pts, [1], [len(pts)], 0.1)
    with pytest.raises(RuntimeError):
        mvk.batch_neighbors(pts[:, :2], pts, lens, lens, 0.1)


def test_neighbors_host_c_abi_entry(mvk):
    """The host-buffer entry point a reference maintainer would bind."""
    import ctypes as C
    L = mvk._lib.lib()
    rng = np.random.default_rng(5)
    pts = bumpy_cloud(rng, 4000)
    lens = np.array([1500, 2500], np.int32)
    out = C.c_void_p()
    width = C.c_int()
    rc = L.mvk_batch_neighbors_host(pts.ctypes.data, len(pts), pts.ctypes.data, len(pts), lens.ctypes.data,
                                    lens.ctypes.data, 2, 0.1, C.byref(out), C.byref(width))
    assert rc == 0
    arr = np.ctypeslib.as_array(C.cast(out, C.POINTER(C.c_int)), shape=(len(pts) * width.value,)).copy()
    L.mvk_free_host(out)
    assert np.array_equal(arr.reshape(len(pts), width.value), geom.batch_neighbors(pts, pts, lens, lens, 0.1))


def test_grid_subsample_host_c_abi_entry(mvk):
    """mvk_grid_subsample_host: host buffers in, malloc'ed host buffers out (the entry point a maintainer of the
    reference's cpp_subsampling wrapper would bind) -- points, features, labels and lengths
    against the reference goldens (max_p crop included)."""
    import ctypes as C
    L = mvk._lib.lib()
    g = load_golden("geometry")
    pts = np.ascontiguousarray(g["points"], np.float32)
    feats = np.ascontiguousarray(g["features"], np.float32).reshape(len(pts), -1)
    labs = np.ascontiguousarray(g["labels"], np.int32).reshape(len(pts), -1)
    lens = np.ascontiguousarray(g["lengths"], np.int32)
    op, of, ol = C.c_void_p(), C.c_void_p(), C.c_void_p()
    olen = np.zeros(len(lens), np.int32)
    tot = C.c_int()
    rc = L.mvk_grid_subsample_host(pts.ctypes.data, len(pts), feats.ctypes.data, feats.shape[1], labs.ctypes.data,
                                   labs.shape[1], lens.ctypes.data, len(lens), 0.2, 150, C.byref(op), C.byref(of),
                                   C.byref(ol), olen.ctypes.data, C.byref(tot))
    assert rc == 0
    m = tot.value
    take = lambda p, ct, shape: np.ctypeslib.as_array(C.cast(p, C.POINTER(ct)), shape=(int(np.prod(shape)),)).copy().reshape(shape)
    sp, sf, sl = take(op, C.c_float, (m, 3)), take(of, C.c_float, (m, feats.shape[1])), take(ol, C.c_int, (m, labs.shape[1]))
    for p in (op, of, ol):
        L.mvk_free_host(p)
    assert np.array_equal(sp, g["sub2_pts"]) and np.array_equal(olen, g["sub2_len"])
    assert np.array_equal(sf.reshape(g["sub2_feats"].shape), g["sub2_feats"])
    assert np.array_equal(sl.reshape(g["sub2_labels"].shape), g["sub2_labels"])
    # points only (NULL features / labels), no crop
    op = C.c_void_p()
    rc = L.mvk_grid_subsample_host(pts.ctypes.data, len(pts), None, 0, None, 0, lens.ctypes.data, len(lens), 0.12, 0,
                                   C.byref(op), None, None, olen.ctypes.data, C.byref(tot))
    assert rc == 0
    sp = take(op, C.c_float, (tot.value, 3))
    L.mvk_free_host(op)
    assert np.array_equal(sp, g["sub_pts"]) and np.array_equal(olen, g["sub_len"])


def test_subsampling_vs_reference_golden(mvk):
    g = load_golden("geometry")
    sp, sl = mvk.batch_grid_subsampling(g["points"], g["lengths"], sampleDl=0.12, random_grid_orient=False)
    assert np.array_equal(sp, g["sub_pts"]) and np.array_equal(sl, g["sub_len"])
    sp, sl, sf, sc = mvk.batch_grid_subsampling(g["points"], g["lengths"], features=g["features"], labels=g["labels"],
                                                sampleDl=0.2, max_p=150, random_grid_orient=False)
    assert np.array_equal(sp, g["sub2_pts"]) and np.array_equal(sl, g["sub2_len"])
    assert np.array_equal(sf, g["sub2_feats"]) and np.array_equal(sc, g["sub2_labels"])


@pytest.mark.parametrize("n,nb,dl", [(10, 1, 0.5), (14, 1, 0.001), (3000, 3, 0.1), (40000, 5, 0.05), (150000, 2, 0.03)])
def test_subsampling_vs_oracle_seeded(mvk, n, nb, dl):
    rng = np.random.default_rng(n + 1)
    pts = bumpy_cloud(rng, n)
    cuts = np.sort(rng.choice(np.arange(1, n), nb - 1, replace=False)) if nb > 1 else np.array([], int)
    lens = np.diff(np.concatenate([[0], cuts, [n]])).astype(np.int32)
    ref = geom.grid_subsample_batch(pts, lens, sampleDl=dl)
    out = mvk.batch_grid_subsampling(pts, lens, sampleDl=dl, random_grid_orient=False)
    assert np.array_equal(out[1], ref[1])
    assert np.array_equal(out[0], ref[0])  # bit-exact barycentres AND libstdc++ emission order
    feats = rng.normal(size=(n, 3)).astype(np.float32)
    labels = rng.integers(0, 21, n).astype(np.int32)
    ref = geom.grid_subsample_batch(pts, lens, features=feats, labels=labels, sampleDl=dl * 2, max_p=n // (4 * nb))
    out = mvk.batch_grid_subsampling(pts, lens, features=feats, labels=labels, sampleDl=dl * 2, max_p=n // (4 * nb),
                                     random_grid_orient=False)
    for a, b in zip(out, ref):
        assert np.array_equal(a, b.reshape(a.shape))
    # single-cloud form (datasets/common.py:44-74)
    one = mvk.grid_subsampling(pts, features=feats, labels=labels, sampleDl=dl * 2)
    refone = geom.grid_subsample_batch(pts, np.array([n], np.int32), features=feats, labels=labels, sampleDl=dl * 2)
    assert np.array_equal(one[0], refone[0]) and np.array_equal(one[1], refone[2])
    assert np.array_equal(one[2].reshape(-1), refone[3].reshape(-1))


def test_subsampling_random_grid_orient_replays_reference_rng(mvk):
    """random_grid_orient=True draws theta/phi/alpha from np.random like common.py:98-105."""
    rng = np.random.default_rng(3)
    pts = bumpy_cloud(rng, 5000)
    lens = np.array([2000, 3000], np.int32)
    np.random.seed(11)
    sp, sl = mvk.batch_grid_subsampling(pts, lens, sampleDl=0.1)
    # host restatement with the same stream
    np.random.seed(11)
    B = 2
    theta = np.random.rand(B) * 2 * np.pi
    phi = (np.random.rand(B) - 0.5) * np.pi
    u = np.vstack([np.cos(theta) * np.cos(phi), np.sin(theta) * np.cos(phi), np.sin(phi)])
    alpha = np.random.rand(B) * 2 * np.pi
    R = mvk.create_3D_rotations(u.T, alpha).astype(np.float32)
    rot = pts.copy()
    i0 = 0
    for b, l in enumerate(lens):
        rot[i0:i0 + l] = np.sum(np.expand_dims(pts[i0:i0 + l], 2) * R[b], axis=1)
        i0 += l
    rp, rl = geom.grid_subsample_batch(rot, lens, sampleDl=0.1)
    i0 = 0
    for b, l in enumerate(rl):
        rp[i0:i0 + l] = np.sum(np.expand_dims(rp[i0:i0 + l], 2) * R[b].T, axis=1)
        i0 += l
    assert np.array_equal(sl, rl) and np.array_equal(sp, rp)


def test_subsampling_properties_at_scale(mvk):
    """Full-size (config 2: 8 spheres) size-independent properties: idempotence of the voxel
    assignment, barycentres inside their voxel, counts conserved."""
    from mvkpconv_b200 import synthetic
    sub = lambda p, dl: mvk.grid_subsampling(p, sampleDl=dl)
    spheres = synthetic.make_spheres(8, sub, seed=0)
    pts, lens = synthetic.stack(spheres)
    assert len(pts) > 100000
    sp, sl = mvk.batch_grid_subsampling(pts, lens, sampleDl=0.08, random_grid_orient=False)
    assert sl.sum() == len(sp) and (sl > 0).all()
    # subsampling the barycentres again with the same grid keeps one point per voxel
    sp2, sl2 = mvk.batch_grid_subsampling(sp, sl, sampleDl=0.08, random_grid_orient=False)
    assert np.array_equal(sl2, sl)
    # neighbours: symmetric relation on the same cloud, sorted rows, self first
    nb = mvk.batch_neighbors(pts, pts, lens, lens, 0.1)
    assert np.array_equal(nb[:, 0], np.arange(len(pts)))
    valid = nb < len(pts)
    i = np.repeat(np.arange(len(pts)), nb.shape[1])[valid.ravel()]
    j = nb.ravel()[valid.ravel()]
    fwd = set(zip(i[:200000].tolist(), j[:200000].tolist()))
    allp = set(zip(i.tolist(), j.tolist()))
    assert all((b, a) in allp for a, b in fwd)


def test_scene_sweep_geometry_at_full_size(mvk):
    """BASELINE config 5 size: a 200k-point synthetic scene, multi-level grid subsampling + radius
    search at every level.  Full-size checks through size-independent properties (rows sorted by
    d2, strict radius, shadow padding, symmetry of the relation, idempotent subsampling) plus exact
    agreement with the oracle on a random subset of the query rows of every level."""
    from mvkpconv_b200 import synthetic
    room = synthetic.make_room(seed=3, density=5200.0, size=(8.0, 6.0, 2.8), n_boxes=12)
    pts = mvk.grid_subsampling(room, sampleDl=0.04)
    assert 150_000 < len(pts) < 300_000
    rng = np.random.default_rng(0)
    radius, dl = 0.1, 0.08
    for level in range(4):
        n = len(pts)
        lens = np.array([n], np.int32)
        inds, counts = mvk.batch_neighbors(pts, pts, lens, lens, radius, return_counts=True)
        assert inds.shape[0] == n and inds.shape[1] == counts.max()
        # properties on every row
        sp = np.concatenate([pts, np.full((1, 3), 1e6, np.float32)], 0)
        d = pts[:, None, :] - sp[inds]
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]).astype(np.float32)
        d2 = (d2 + d[..., 2] * d[..., 2]).astype(np.float32)
        real = inds < n
        assert np.array_equal(real.sum(1), counts)
        assert np.all(real[:, :-1] >= real[:, 1:])                        # shadows only at the end
        assert np.all(d2[real] < np.float32(radius) * np.float32(radius))  # strict radius, fp32 metric
        d2s = np.where(real, d2, np.inf)
        assert np.all(d2s[:, :-1] <= d2s[:, 1:])                          # sorted by distance
        assert np.all(inds[:, 0] == np.arange(n))                         # the query itself comes first (d2 = 0)
        # symmetry of the radius relation on a sample of pairs
        qi = rng.integers(0, n, 2000)
        for i in qi[:200]:
            for j in inds[i, :counts[i]]:
                assert i in inds[j, :counts[j]]
        # exact agreement with the oracle on a subset of the rows
        sel = np.sort(rng.choice(n, min(n, 3000), replace=False))
        ref = geom.batch_neighbors(pts[sel], pts, np.array([len(sel)], np.int32), lens, radius)
        w = ref.shape[1]
        assert np.all(inds[sel, w:] == n) if inds.shape[1] > w else True
        assert np.array_equal(inds[sel, :w], ref[:, :inds.shape[1]][:, :w])
        # next level: subsampling is exact vs the oracle and idempotent in count
        sub, sl = mvk.batch_grid_subsampling(pts, lens, sampleDl=dl, random_grid_orient=False)
        rsub, rsl = geom.grid_subsample_batch(pts, lens, sampleDl=dl)
        assert np.array_equal(sub, rsub) and np.array_equal(sl, rsl)
        pts, radius, dl = sub, radius * 2, dl * 2
