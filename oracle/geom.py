"""ctypes front-end of the geometric oracle.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``orc_*``  : plain-C restatement in oracle_geom.c (travels with the repo as source + built .so).
``ref_*``  : the UNMODIFIED reference C++ (neighbors.cpp, grid_subsampling.cpp, cloud.cpp,
             nanoflann.hpp) behind oracle/ref_shim.cpp, prebuilt into oracle/_ref/libref.so.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_P = C.POINTER
_f32p, _i32p, _u64p = _P(C.c_float), _P(C.c_int), _P(C.c_uint64)


def build(force=False):
    """Compile liboracle.so (always possible: gcc) and _ref/libref.so (needs /root/reference)."""
    so = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(
            os.path.join(_HERE, "oracle_geom.c")):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference") and (force or not os.path.exists(
            os.path.join(_HERE, "_ref", "libref.so"))):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_orc = None
_ref = None


def _lib():
    global _orc
    if _orc is None:
        build()
        _orc = C.CDLL(os.path.join(_HERE, "liboracle.so"))
        _orc.orc_batch_neighbors.restype = C.c_int
        _orc.orc_grid_subsample_batch.restype = C.c_int
        _orc.probe_bucket_schedule.restype = C.c_int
        _orc.probe_policy_schedule.restype = C.c_int
    return _orc


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "libref.so"))


def _reflib():
    global _ref
    if _ref is None:
        if not have_ref():
            build()
        _ref = C.CDLL(os.path.join(_HERE, "_ref", "libref.so"))
        _ref.ref_batch_neighbors.restype = C.c_int
        _ref.ref_grid_subsample_batch.restype = C.c_int
        _ref.ref_grid_subsample.restype = C.c_int
    return _ref


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _take(ptr, shape, dtype, free):
    n = int(np.prod(shape))
    out = np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n].astype(dtype, copy=True).reshape(shape)
    free(ptr)
    return out


# ---------------------------------------------------------------------------------------------
def batch_neighbors(queries, supports, q_batches, s_batches, radius, return_counts=False):
    """(d2, idx)-ordered radius neighbours == reference batch_ordered_neighbors
    (neighbors.cpp:125-208); == batch_nanoflann_neighbors (:211-332) up to exact-d2 tie order."""
    lib = _lib()
    q, s, qb, sb = _f32(queries), _f32(supports), _i32(q_batches), _i32(s_batches)
    out = _i32p()
    counts = np.zeros(max(len(q), 1), dtype=np.int32)
    w = lib.orc_batch_neighbors(q.ctypes.data_as(_f32p), C.c_int(len(q)), s.ctypes.data_as(_f32p),
                                C.c_int(len(s)), qb.ctypes.data_as(_i32p), sb.ctypes.data_as(_i32p),
                                C.c_int(len(qb)), C.c_float(radius), C.byref(out),
                                counts.ctypes.data_as(_i32p))
    res = _take(out, (len(q), w), np.int32, lib.orc_free)
    return (res, counts[:len(q)]) if return_counts else res


def ref_batch_neighbors(queries, supports, q_batches, s_batches, radius, ordered=False):
    """The compiled reference: nanoflann (what wrapper.cpp:198 calls) or, with ordered=True, the
    reference's own stable alternative batch_ordered_neighbors (wrapper.cpp:197)."""
    lib = _reflib()
    q, s, qb, sb = _f32(queries), _f32(supports), _i32(q_batches), _i32(s_batches)
    out = _i32p()
    w = lib.ref_batch_neighbors(q.ctypes.data_as(_f32p), C.c_int(len(q)), s.ctypes.data_as(_f32p),
                                C.c_int(len(s)), qb.ctypes.data_as(_i32p), sb.ctypes.data_as(_i32p),
                                C.c_int(len(qb)), C.c_float(radius), C.c_int(1 if ordered else 0),
                                C.byref(out))
    return _take(out, (len(q), w), np.int32, lib.ref_free)


def _subsample(fn, free, points, batches, features, labels, sampleDl, max_p):
    p = _f32(points)
    n = len(p)
    b = _i32(batches)
    f = None if features is None else _f32(features).reshape(n, -1)
    l = None if labels is None else _i32(labels).reshape(n, -1)
    fdim = 0 if f is None else f.shape[1]
    ldim = 0 if l is None else l.shape[1]
    op, of, oc, ob = _f32p(), _f32p(), _i32p(), _i32p()
    m = fn(p.ctypes.data_as(_f32p), C.c_int(n), None if f is None else f.ctypes.data_as(_f32p),
           C.c_int(fdim), None if l is None else l.ctypes.data_as(_i32p), C.c_int(ldim),
           b.ctypes.data_as(_i32p), C.c_int(len(b)), C.c_float(sampleDl), C.c_int(max_p),
           C.byref(op), C.byref(of), C.byref(oc), C.byref(ob))
    res = [_take(op, (m, 3), np.float32, free), _take(ob, (len(b),), np.int32, free)]
    sf = _take(of, (m, fdim), np.float32, free) if fdim else free(of)
    sl = _take(oc, (m, ldim), np.int32, free) if ldim else free(oc)
    if fdim:
        res.append(sf)
    if ldim:
        res.append(sl)
    return tuple(res)


def grid_subsample_batch(points, batches, features=None, labels=None, sampleDl=0.1, max_p=0):
    """Restatement of subsample_batch (grid_subsampling.cpp:109-210) incl. emission order."""
    lib = _lib()
    return _subsample(lib.orc_grid_subsample_batch, lib.orc_free, points, batches, features, labels,
                      sampleDl, max_p)


def ref_grid_subsample_batch(points, batches, features=None, labels=None, sampleDl=0.1, max_p=0):
    lib = _reflib()
    return _subsample(lib.ref_grid_subsample_batch, lib.ref_free, points, batches, features, labels,
                      sampleDl, max_p)


def stl_order(keys):
    """Iteration order (insertion ranks) of libstdc++ unordered_map for distinct size_t keys."""
    lib = _lib()
    k = np.ascontiguousarray(keys, dtype=np.uint64)
    o = np.zeros(max(len(k), 1), dtype=np.int32)
    lib.orc_stl_order(k.ctypes.data_as(_u64p), C.c_int(len(k)), o.ctypes.data_as(_i32p))
    return o[:len(k)]


def probe_stl_order(keys):
    """Same, from the live std::unordered_map<size_t,int> of this toolchain."""
    lib = _lib()
    k = np.ascontiguousarray(keys, dtype=np.uint64)
    o = np.zeros(max(len(k), 1), dtype=np.int32)
    lib.probe_unordered_order(k.ctypes.data_as(_u64p), C.c_int(len(k)), o.ctypes.data_as(_i32p))
    return o[:len(k)]


def probe_stl_order_int(keys):
    lib = _lib()
    k = np.ascontiguousarray(keys, dtype=np.int32)
    o = np.zeros(max(len(k), 1), dtype=np.int32)
    lib.probe_unordered_order_int(k.ctypes.data_as(_i32p), C.c_int(len(k)), o.ctypes.data_as(_i32p))
    return o[:len(k)]


def probe_schedule(policy=True, max_elems=2_000_000):
    lib = _lib()
    s = np.zeros(64, dtype=np.uint64)
    if policy:
        n = lib.probe_policy_schedule(s.ctypes.data_as(_u64p), C.c_int(64))
    else:
        n = lib.probe_bucket_schedule(C.c_uint64(max_elems), s.ctypes.data_as(_u64p), C.c_int(64))
    return s[:n].copy()
