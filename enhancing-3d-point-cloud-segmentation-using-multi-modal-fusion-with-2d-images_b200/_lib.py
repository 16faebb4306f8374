"""ctypes binding of libmvk.so -- the ONLY compute path of this package.

There is no CPU / eager fallback: if the CUDA library cannot be loaded (or no CUDA device is
present when an op is called) the ops raise RuntimeError.
"""
import ctypes as C
import os

from . import build as _build

_LIB = None

from ._abi import SIGNATURES, vp, i32, i64, f32, sz  # noqa: F401  (name -> (restype, argtypes))


def lib_path():
    return _build.LIB


class _ProfiledLib:
    """Proxy around the ctypes handle that brackets every C-ABI call with CUDA events on the
    current stream (the stream the kernels are launched on).  Used by bench.py to measure the
    per-kernel durations INSIDE the timed region; never active otherwise."""

    def __init__(self, handle, records):
        self._h, self._r = handle, records

    def __getattr__(self, name):
        fn = getattr(self._h, name)
        if not name.startswith("mvk_") or name in ("mvk_error_string", "mvk_last_cuda_error", "mvk_version",
                                                   "mvk_launch_count", "mvk_free_host") or name.endswith("_bytes"):
            return fn
        import torch

        def wrapped(*args):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            self._r.append((name, args, e0, e1))
            return rc

        return wrapped


_PROFILE_RECORDS = None


class profile:
    """with _lib.profile() as records: ...   -> list of (entry point, args, start event, end event)."""

    def __enter__(self):
        global _PROFILE_RECORDS
        _PROFILE_RECORDS = []
        return _PROFILE_RECORDS

    def __exit__(self, *exc):
        global _PROFILE_RECORDS
        _PROFILE_RECORDS = None
        return False


class _FastLib:
    """The library handle the ops call through.  Entry points whose arguments are all plain pointers,
    integers and floats are bound through the generated CPython shim (_mvkcall, METH_FASTCALL: ~0.4 us
    per call instead of ~5 us of ctypes marshalling for a 25-argument launch -- the step makes ~450
    such calls); everything else, and everything when the shim is unavailable, goes through ctypes.
    Either way the callee is the same C-ABI function of libmvk.so."""

    def __init__(self, handle):
        self._handle = handle
        fast = None
        if os.environ.get("MVK_FASTCALL", "1") != "0":
            try:
                from . import _mvkcall as fast  # built next to libmvk.so by build.py
            except ImportError:
                fast = None
        self._fast = fast
        for name in SIGNATURES:
            fn = getattr(fast, name, None) if fast is not None else None
            setattr(self, name, fn if fn is not None else getattr(handle, name))

    def __getattr__(self, name):  # symbols outside SIGNATURES
        return getattr(self._handle, name)


def lib():
    """Load (building first if stale and nvcc is present).  Raises if the CUDA library is unavailable."""
    global _LIB
    if _LIB is None:
        path = _build.LIB
        if not _build.is_current():
            if _build.nvcc_path() is not None:
                path = _build.build()
            elif not os.path.exists(path):
                raise RuntimeError(
                    "libmvk.so not found and nvcc unavailable: the CUDA extension is required "
                    "(there is no CPU fallback). Run `python -c 'import __graft_entry__ as g; g.build()'`.")
        handle = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError = header/library mismatch: fail loudly
            fn.restype = res
            fn.argtypes = args
        _LIB = _FastLib(handle)
    if _PROFILE_RECORDS is not None:
        return _ProfiledLib(_LIB, _PROFILE_RECORDS)
    return _LIB


class MvkError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        l = lib()
        msg = l.mvk_error_string(rc).decode()
        if rc == -3:
            msg += ": " + l.mvk_last_cuda_error().decode()
        raise MvkError(msg)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("mvkpconv_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    lib()


class _ZeroPool:
    """Small zero-initialised scratch buffers (batch-norm sums, retirement tickets) carved out of
    one larger zeroed allocation: one fill per ~4 MB instead of one per buffer.  Slices are never
    reused; the chunk is dropped when exhausted (the caching allocator recycles it once the slices die)."""
    CHUNK = 4 << 20

    def __init__(self):
        self.buf, self.off, self.key = None, 0, None

    def take(self, nbytes, device, dtype):
        import torch
        nbytes = (nbytes + 255) // 256 * 256
        key = (device.index, torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch.cuda.current_device()))
        if nbytes > self.CHUNK // 4:
            return torch.zeros(nbytes, dtype=torch.uint8, device=device).view(dtype)
        if self.buf is None or self.key != key or self.off + nbytes > self.CHUNK:
            self.buf, self.off, self.key = torch.zeros(self.CHUNK, dtype=torch.uint8, device=device), 0, key
        out = self.buf[self.off:self.off + nbytes].view(dtype)
        self.off += nbytes
        return out

    def take_ptr(self, nbytes, device):
        import torch
        nbytes = (nbytes + 255) // 256 * 256
        key = (device.index, torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch.cuda.current_device()))
        if self.buf is None or self.key != key or self.off + nbytes > self.buf.numel():
            self.buf, self.off, self.key = torch.zeros(max(self.CHUNK, nbytes), dtype=torch.uint8, device=device), 0, key
        p = self.buf.data_ptr() + self.off
        self.off += nbytes
        return p


_ZEROS = _ZeroPool()


def zeros_f64(n, device):
    """n zero-initialised doubles (pooled)."""
    return _ZEROS.take(8 * n, device, __import__("torch").float64)[:n]


def zeros_ptr(nbytes, device):
    """Device pointer of `nbytes` zero-initialised bytes from the pool (no tensor object is created;
    the memory stays valid for everything enqueued on the current stream: see _ZeroPool)."""
    return _ZEROS.take_ptr(nbytes, device)


def stream_ptr():
    """Raw cudaStream_t of torch's current stream on the current device (as an int for ctypes)."""
    import torch
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def ptr(t):
    """Device pointer of a (contiguous) torch tensor (int), or None (= NULL)."""
    if t is None:
        return None
    assert t.is_contiguous(), "libmvk expects contiguous tensors"
    return t.data_ptr()


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL_CTX = _NullCtx()


def on_device(device):
    """Context that makes `device` current -- free when it already is (the common case)."""
    import torch
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NULL_CTX
    return torch.cuda.device(device)
