"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports
every symbol include/mvk.h declares; host-side mirrors behave like the reference's host code.
No compute call is made here (no GPU in the authoring container)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden


def _header_symbols():
    h = open(os.path.join(ROOT, "include", "mvk.h")).read()
    return sorted(set(re.findall(r"\b(mvk_[a-z0-9_]+)\s*\(", h)))


def test_library_builds_and_exports_every_declared_symbol(mvk):
    path = mvk.build.build()
    assert os.path.exists(path)
    handle = ctypes.CDLL(path)
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(handle, s), f"{s} declared in include/mvk.h but not exported by libmvk.so"
    # and the python binding table is in sync with the header
    assert sorted(mvk._lib.SIGNATURES) == syms


def test_library_is_sm100a_with_tcgen05_and_tma(mvk):
    """The contraction kernel must be Blackwell-native: UTC*MMA (tcgen05.mma), UTMALDG (TMA), LDTM."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sass = subprocess.run(["cuobjdump", "-sass", mvk.build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass


def test_fastcall_shim_mirrors_the_ctypes_binding(mvk):
    """The generated CPython shim (_mvkcall) binds the same C-ABI functions as ctypes: every entry point made
    of pointers / integers / floats is present, argument checking happens in the callee (same return code
    through both bindings), and bad Python arguments raise instead of reaching the library."""
    import ctypes as C
    mvk.build.build()
    L = mvk._lib.lib()
    fast = L._fast
    assert fast is not None, "_mvkcall.so was not built (gcc / Python.h missing?)"
    plain = {C.c_void_p, C.c_int, C.c_longlong, C.c_size_t, C.c_float, C.c_double}
    expected = [n for n, (res, args) in mvk._lib.SIGNATURES.items() if res is C.c_int and all(a in plain for a in args)]
    assert len(expected) >= 25
    for n in expected:
        assert hasattr(fast, n), n
        assert getattr(L, n) is getattr(fast, n)
    handle = L._handle
    args = (0, 4, 0, 0, 8, 8, 0, 0, None, 8, 0, 0, 0.1, 0, 1, None, 0, 0, 0, 8, None, 8, None, None, 0)
    assert fast.mvk_act_bwd_apply(*args) == handle.mvk_act_bwd_apply(*args) == -1  # MVK_ERR_INVALID_ARG
    assert fast.mvk_split_bf16(None, np.int64(4), 4, 4, None, None, 4, 8, None) == handle.mvk_split_bf16(
        None, 4, 4, 4, None, None, 4, 8, None)
    with pytest.raises(TypeError):
        fast.mvk_split_bf16(None, 4, 4)              # wrong arity
    with pytest.raises(TypeError):
        fast.mvk_split_bf16("x", 4, 4, 4, None, None, 4, 8, None)  # not a pointer
    assert fast.mvk_version() == handle.mvk_version()


def test_error_strings_and_version(mvk):
    L = mvk._lib.lib()
    assert L.mvk_version() >= 100
    assert L.mvk_error_string(0) == b"ok"
    assert L.mvk_error_string(-6) == b"Error"  # the reference's RuntimeError text for empty results
    assert L.mvk_neighbors_workspace_bytes(1000, 1000, 2) > 0
    assert L.mvk_subsample_workspace_bytes(1000, 2, 3, 1) > 0


def test_ops_fail_loudly_without_cuda(mvk):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    pts = np.zeros((10, 3), np.float32)
    with pytest.raises(RuntimeError):
        mvk.batch_neighbors(pts, pts, [10], [10], 0.1)
    with pytest.raises(RuntimeError):
        mvk.grid_subsampling(pts, sampleDl=0.1)
    conv = mvk.KPConv(15, 3, 4, 8, 0.05, 0.1)
    with pytest.raises(RuntimeError):
        conv(torch.zeros(10, 3), torch.zeros(10, 3), torch.zeros(10, 4, dtype=torch.long), torch.zeros(10, 4))


def test_kpconv_module_surface(mvk):
    """Same ctor / state-dict keys / repr as blocks.py:143-379 so reference checkpoints load."""
    import torch
    np.random.seed(3)
    conv = mvk.KPConv(15, 3, 64, 128, 0.048, 0.1)
    sd = conv.state_dict()
    assert set(sd) == {"weights", "kernel_points"}
    assert sd["weights"].shape == (15, 64, 128) and sd["kernel_points"].shape == (15, 3)
    assert conv.kernel_points.requires_grad is False and conv.weights.requires_grad
    assert repr(conv) == "KPConv(radius: 0.10, in_feat: 64, out_feat: 128)"
    g = load_golden("kernel_points")
    assert np.array_equal(conv.kernel_points.numpy(), g["loaded_seed3_r01"])  # same np.random stream as the reference
    for attr in ("deformable", "min_d2", "deformed_KP", "KP_extent", "K", "radius"):
        assert hasattr(conv, attr)
    # deformable: the reference's parameter tree (blocks.py:186-204)
    d = mvk.KPConv(15, 3, 4, 8, 0.05, 0.1, deformable=True, modulated=True)
    assert set(d.state_dict()) == {"weights", "kernel_points", "offset_bias", "offset_conv.weights",
                                   "offset_conv.kernel_points"}
    assert d.offset_conv.weights.shape == (15, 4, 60) and d.offset_bias.shape == (60,)
    assert mvk.KPConv(15, 3, 4, 8, 0.05, 0.1, modulated=True).offset_conv is None  # ignored when rigid, like the reference
    with pytest.raises(ValueError):
        mvk.KPConv(15, 3, 4, 8, 0.05, 0.1, KP_influence="cubic")


def test_feature_aggregation_state_dict_names(mvk):
    fa = mvk.FeatureAggregation(64)
    g = load_golden("feature_aggregation")["sum64"]
    ref_keys = {k[3:] for k in g if k.startswith("sd.")}
    assert set(fa.state_dict()) == ref_keys


def test_create_3d_rotations_is_a_rotation(mvk):
    rng = np.random.default_rng(0)
    u = rng.normal(size=(5, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    a = rng.uniform(0, 6.28, 5)
    R = mvk.create_3D_rotations(u, a)
    for i in range(5):
        assert np.allclose(R[i] @ R[i].T, np.eye(3), atol=1e-12)
        assert np.allclose(R[i] @ u[i], u[i], atol=1e-12)  # axis is invariant
        assert np.isclose(np.trace(R[i]), 1 + 2 * np.cos(a[i]))
