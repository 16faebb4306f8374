"""Host-side mirror of the reference's pre-processing wrappers (datasets/common.py:44-196).

Same names, argument order, defaults and return arity as the reference functions; the work is
done by libmvk's hashed-grid CUDA kernels.  Inputs may be numpy arrays (the reference's calling
convention: copied host->device, results copied back as numpy) or torch CUDA tensors (results
stay on the device).  Failures raise RuntimeError like the CPython extensions do
(cpp_neighbors/wrapper.cpp:77-205, cpp_subsampling/wrapper.cpp:77-270).
"""
import ctypes as C
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

_WS = {}
_GRID = {}  # device -> signature of the cell grid currently sitting in the neighbour workspace


def _workspace(nbytes, device, kind="subsample"):
    """Grow-only scratch buffer per device and kind (the C ABI never allocates).  Radius search and
    subsampling use separate buffers so that a cell grid survives the subsampling call in between."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), kind)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
        _WS[key] = buf
        if kind == "neighbors":
            _GRID.pop(key[0], None)
    return buf


def _dev(x, dtype, device=None):
    if isinstance(x, torch.Tensor):
        if x.dtype is dtype and x.is_cuda and x.is_contiguous():
            return x  # the common case inside the pyramid: nothing to convert
        if not x.is_cuda:
            x = x.cuda(device) if device is not None else x.cuda()
        return x.to(dtype).contiguous()
    arr = np.ascontiguousarray(np.asarray(x), dtype={torch.float32: np.float32, torch.int32: np.int32}[dtype])
    t = torch.from_numpy(arr)
    return t.cuda(device, non_blocking=False) if device is not None else t.cuda()


def _is_np(*xs):
    return not any(isinstance(x, torch.Tensor) for x in xs)


def create_3D_rotations(axis, angle):
    """Axis-angle -> rotation matrices [N,3,3] (reference: kernels/kernel_points.py:44-75,
    the quaternion-derived closed form)."""
    axis = np.asarray(axis)
    ux, uy, uz = axis[:, 0], axis[:, 1], axis[:, 2]
    c, s = np.cos(angle), np.sin(angle)
    k = 1 - c
    kx = k * ux
    xy = kx * uy
    xz = kx * uz
    yz = k * uy * uz
    R = np.stack([c + k * (ux * ux), xy - s * uz, xz + s * uy,
                  xy + s * uz, c + k * (uy * uy), yz - s * ux,
                  xz - s * uy, yz + s * ux, c + k * (uz * uz)], axis=1)
    return np.reshape(R, (-1, 3, 3))


# -------------------------------------------------------------------------------------------------
def batch_neighbors(queries, supports, q_batches, s_batches, radius, max_neighbors=None,
                    out_dtype=torch.int32, return_counts=False, deferred=None):
    """Radius neighbours inside every element of a stacked batch; drop-in for the wrapper at
    datasets/common.py:185-196 (same argument order).

    queries [N1, 3] and supports [N2, 3] are stacked clouds, q_batches / s_batches [B] give the number of points
    of each element, radius is rounded to float32 like the C extension does.  Returns int32 [N1, max_count]:
    for every query the stacked indices of its element's supports closer than radius, nearest first (ties by
    index), rows filled up with N2.

    Extras (not in the reference signature, defaults keep its behaviour): `max_neighbors` crops
    the rows to their nearest entries on the device (== big_neighborhood_filter,
    common.py:411-421); `out_dtype=torch.int64` emits the dtype the model consumes; `deferred`
    (a list, with `max_neighbors`) postpones the read-back of the maximum hit count: the call
    returns without a host sync and appends a check that `resolve_deferred` runs later for all
    calls at once (it returns the corrected matrices in the rare case a check fails).
    """
    _lib.require_cuda()
    L = _lib.lib()
    as_np = _is_np(queries, supports)
    q = _dev(queries, torch.float32)
    dev = q.device
    s = _dev(supports, torch.float32, dev)
    qb = _dev(q_batches, torch.int32, dev).reshape(-1)
    sb = _dev(s_batches, torch.int32, dev).reshape(-1)
    if q.dim() != 2 or q.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : query.shape is not (N, 3)")
    if s.dim() != 2 or s.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : support.shape is not (N, 3)")
    if qb.numel() != sb.numel():
        raise RuntimeError("Wrong number of batch elements: different for queries and supports ")
    nq, ns, nb = q.shape[0], s.shape[0], qb.numel()
    with _lib.on_device(dev):
        wsb = L.mvk_neighbors_workspace_bytes(nq, ns, nb)
        ws = _workspace(wsb, dev, "neighbors")
        dkey = dev.index if dev.index is not None else torch.cuda.current_device()
        # Is the grid in `ws` the one of THIS support tensor (same object, unmodified), batch split, radius?
        # The cache keeps the tensors alive, so a match cannot be a recycled address with other contents.
        sig = None
        if isinstance(supports, torch.Tensor) and isinstance(s_batches, torch.Tensor):
            sig = (s, s._version, sb, sb.data_ptr(), sb._version, ns, nb, float(np.float32(radius)),
                   torch._C._cuda_getCurrentRawStream(dkey))
        old = _GRID.pop(dkey, None)  # whatever runs below rebuilds or invalidates it
        reuse = 1 if (sig is not None and old is not None and old[0] is sig[0] and old[1] == sig[1] and
                      old[3:] == sig[3:]) else 0
        counts = torch.empty(max(nq, 1), dtype=torch.int32, device=dev)
        hmax = torch.zeros(1, dtype=torch.int32, device=dev)
        if max_neighbors is not None and nq > 0:
            # the row width is capped by the caller: one query pass (no separate count pass)
            width = max(1, int(max_neighbors))
            cap = max(128, (width + 31) // 32 * 32)
            out = torch.empty((nq, width), dtype=out_dtype, device=dev)
            check(L.mvk_neighbors_query_capped(ptr(q), nq, ptr(s), ns, ptr(qb), ptr(sb), nb, float(radius), ptr(ws),
                                               ws.numel(), width, cap, ptr(out), 1 if out_dtype == torch.int64 else 0,
                                               ptr(counts), ptr(hmax), reuse, stream_ptr()))
            if sig is not None:
                _GRID[dkey] = sig  # the grid of (supports, radius) now sits in the workspace
            if deferred is not None and not as_np and not return_counts:
                args = (queries, supports, q_batches, s_batches, radius, max_neighbors, out_dtype)
                deferred.append(SimpleNamespace(hmax=hmax, width=width, cap=cap, out=out, args=args))
                return out
            max_count = int(hmax.item())  # the one host sync
            if max_count < 0:
                check(-4)
            if max_count < 1:
                raise RuntimeError("Error")  # cpp_neighbors/wrapper.cpp:201-205
            if max_count <= cap:
                if max_count < width:
                    out = out[:, :max_count].contiguous()  # reference width = min(max_count, limit)
                if as_np:
                    out, counts = out.cpu().numpy(), counts.cpu().numpy()
                return (out, counts[:nq]) if return_counts else out
            # more hits than the shared-memory lists hold somewhere: redo with the two-phase protocol
        check(L.mvk_neighbors_count(ptr(q), nq, ptr(s), ns, ptr(qb), ptr(sb), nb, float(radius),
                                    ptr(ws), ws.numel(), ptr(counts), ptr(hmax), stream_ptr()))
        max_count = int(hmax.item())  # the one host sync: the row width is data dependent
        if max_count < 0:
            check(-4)
        if nq * max_count < 1:
            raise RuntimeError("Error")  # cpp_neighbors/wrapper.cpp:201-205
        width = max_count if max_neighbors is None else max(1, min(max_count, int(max_neighbors)))
        out = torch.empty((nq, width), dtype=out_dtype, device=dev)
        fill = L.mvk_neighbors_fill_i64 if out_dtype == torch.int64 else L.mvk_neighbors_fill
        check(fill(ptr(q), nq, ptr(s), ns, ptr(qb), ptr(sb), nb, float(radius), ptr(ws), ws.numel(),
                   max_count, width, ptr(out), stream_ptr()))
    if as_np:
        out = out.cpu().numpy()
        counts = counts.cpu().numpy()
    return (out, counts[:nq]) if return_counts else out


def resolve_deferred(records):
    """One host sync for all deferred neighbour calls.  Returns {index in records: corrected matrix}
    for the calls whose optimistic single-pass result must be replaced (more hits than the
    shared-memory lists hold, or fewer hits than the requested width: both rare)."""
    fixes = {}
    if not records:
        return fixes
    counts = torch.cat([r.hmax for r in records]).cpu().tolist()
    for i, (r, mc) in enumerate(zip(records, counts)):
        if mc < 0:
            check(-4)
        if mc < 1:
            raise RuntimeError("Error")  # cpp_neighbors/wrapper.cpp:201-205
        if mc > r.cap:
            q, s, qb, sb, radius, lim, dt = r.args
            full = batch_neighbors(q, s, qb, sb, radius, out_dtype=dt)  # two-phase protocol
            fixes[i] = full[:, :min(full.shape[1], int(lim))].contiguous()
        elif mc < r.width:
            fixes[i] = r.out[:, :mc].contiguous()  # reference width = min(max_count, limit)
    return fixes


def _subsample(points, lengths, features, labels, sampleDl, max_p):
    L = _lib.lib()
    p = _dev(points, torch.float32)
    dev = p.device
    n = p.shape[0]
    if p.dim() != 2 or p.shape[1] != 3:
        raise RuntimeError("Wrong dimensions : points.shape is not (N, 3)")
    ln = _dev(lengths, torch.int32, dev).reshape(-1)
    f = None if features is None else _dev(features, torch.float32, dev).reshape(n, -1)
    l = None if labels is None else _dev(labels, torch.int32, dev).reshape(n, -1)
    fdim = 0 if f is None else f.shape[1]
    ldim = 0 if l is None else l.shape[1]
    nb = ln.numel()
    with _lib.on_device(dev):
        wsb = L.mvk_subsample_workspace_bytes(n, nb, fdim, ldim)
        ws = _workspace(wsb, dev)
        op = torch.empty((max(n, 1), 3), dtype=torch.float32, device=dev)
        of = torch.empty((max(n, 1), fdim), dtype=torch.float32, device=dev) if fdim else None
        ol = torch.empty((max(n, 1), ldim), dtype=torch.int32, device=dev) if ldim else None
        olen = torch.zeros(nb, dtype=torch.int32, device=dev)
        tot = torch.zeros(1, dtype=torch.int32, device=dev)
        check(L.mvk_grid_subsample(ptr(p), n, ptr(f), fdim, ptr(l), ldim, ptr(ln), nb, float(sampleDl),
                                   int(max_p), ptr(ws), ws.numel(), ptr(op), ptr(of), ptr(ol),
                                   ptr(olen), ptr(tot), stream_ptr()))
        m = int(tot.item())  # host sync: output size is data dependent
    if m < 0:
        check(-4)
    if m < 1:
        raise RuntimeError("Error")  # cpp_subsampling/wrapper.cpp:266-270
    return op[:m], olen, (of[:m] if fdim else None), (ol[:m] if ldim else None)


def _pack(as_np, pts, lens, feats, labs, with_len):
    res = [pts] + ([lens] if with_len else [])
    if feats is not None:
        res.append(feats)
    if labs is not None:
        res.append(labs)
    if as_np:
        res = [r.cpu().numpy() for r in res]
    return res[0] if len(res) == 1 else tuple(res)


def grid_subsampling(points, features=None, labels=None, sampleDl=0.1, verbose=0):
    """CPP wrapper for a grid subsampling (method = barycenter for points and features)
    (datasets/common.py:44-74).

    :param points: (N, 3) matrix of input points
    :param features: optional (N, d) matrix of features (floating number)
    :param labels: optional (N,) matrix of integer labels
    :param sampleDl: parameter defining the size of grid voxels
    :param verbose: 1 to display
    :return: subsampled points, with features and/or labels depending of the input
    """
    _lib.require_cuda()
    as_np = _is_np(points)
    n = len(points)
    lens = torch.tensor([n], dtype=torch.int32)
    pts, _, feats, labs = _subsample(points, lens, features, labels, sampleDl, 0)
    return _pack(as_np, pts, None, feats, labs, with_len=False)


def batch_grid_subsampling(points, batches_len, features=None, labels=None, sampleDl=0.1, max_p=0,
                           verbose=0, random_grid_orient=True):
    """CPP wrapper for a batch grid subsampling (datasets/common.py:77-182).

    With random_grid_orient (the reference default) every batch element is rotated by a random
    rotation drawn from np.random exactly like the reference (three np.random.rand(B) calls:
    theta, phi, alpha), subsampled, and rotated back.
    :return: subsampled points, batch lengths (+ features) (+ labels)
    """
    _lib.require_cuda()
    L = _lib.lib()
    as_np = _is_np(points)
    B = len(batches_len)
    R = None
    p = _dev(points, torch.float32)
    dev = p.device
    ln = _dev(batches_len, torch.int32, dev).reshape(-1)
    if random_grid_orient:
        # common.py:98-111 (host side, same RNG stream as the reference)
        theta = np.random.rand(B) * 2 * np.pi
        phi = (np.random.rand(B) - 0.5) * np.pi
        u = np.vstack([np.cos(theta) * np.cos(phi), np.sin(theta) * np.cos(phi), np.sin(phi)])
        alpha = np.random.rand(B) * 2 * np.pi
        R = torch.from_numpy(create_3D_rotations(u.T, alpha).astype(np.float32)).contiguous().to(dev)
        rot = torch.empty_like(p)
        with _lib.on_device(dev):
            check(L.mvk_rotate_batch(ptr(p), p.shape[0], ptr(ln), B, ptr(R), 0, ptr(rot), stream_ptr()))
        p = rot
    pts, lens, feats, labs = _subsample(p, ln, features, labels, sampleDl, max_p)
    if random_grid_orient:
        pts = pts.contiguous()
        with _lib.on_device(dev):
            check(L.mvk_rotate_batch(ptr(pts), pts.shape[0], ptr(lens), B, ptr(R), 1, ptr(pts), stream_ptr()))
    return _pack(as_np, pts, lens, feats, labs, with_len=True)
