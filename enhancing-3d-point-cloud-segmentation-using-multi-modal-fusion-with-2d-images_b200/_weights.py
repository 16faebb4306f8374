"""bf16 hi/lo operands of the model's weight matrices, refreshed ONCE per optimiser step.

Every contraction consumes its weight as a bf16 hi/lo pair.  Parameters only change
when the optimiser steps, so the pairs live in persistent buffers next to the parameters: the first layer
that finds its pair stale (``param._version`` moved) refreshes the pairs of ALL registered parameters of
that device with one ``mvk_split_bf16_multi`` launch, instead of one tiny split launch per layer and step.
A parameter seen for the first time is split on its own and joins the table for the next refresh.

Staleness has two signals: the parameter's autograd version counter (in-place ops, ``load_state_dict``), and
a global epoch that every ``torch.optim`` optimiser step advances through a step post-hook -- the fused
optimisers (``fused=True``) update parameters WITHOUT moving their version counters.  Code that writes
parameters through ``.data`` (which bypasses both) must call ``invalidate()`` afterwards."""
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import check, stream_ptr

CHUNK = 4096
_DESC = np.dtype([("src", "<u8"), ("hi", "<u8"), ("lo", "<u8"), ("rows", "<i4"), ("cols", "<i4"), ("src_ld", "<i4"),
                  ("rows_pad", "<i4"), ("dst_ld", "<i4"), ("first_chunk", "<i4")], align=True)
assert _DESC.itemsize == 48


_EPOCH = [0]


def invalidate():
    """Declare every cached bf16 weight pair stale (called automatically after each optimiser step)."""
    _EPOCH[0] += 1


try:  # global hook: fires after the step of ANY torch.optim optimiser
    from torch.optim.optimizer import register_optimizer_step_post_hook as _register_post_hook
    _register_post_hook(lambda optimizer, args, kwargs: invalidate())
except ImportError:  # pragma: no cover  (older torch: version counters / explicit invalidate() only)
    pass


class _Entry:
    __slots__ = ("ref", "key", "buf", "hi", "lo", "version", "src_ptr", "epoch", "offset")


class _DeviceTable:
    def __init__(self):
        self.entries = {}     # (id(owner), byte offset of the view, key) -> _Entry
        self.table = None     # device copy of the descriptor array
        self.order = []
        self.chunks = 0
        self.dirty = True


_TABLES = {}


def _table(device):
    idx = device.index if device.index is not None else torch.cuda.current_device()
    t = _TABLES.get(idx)
    if t is None:
        t = _TABLES[idx] = _DeviceTable()
    return t


def _rebuild(t, device):
    live = [(k, e) for k, e in t.entries.items() if e.ref() is not None]
    t.entries = dict(live)
    desc = np.zeros(len(live), _DESC)
    chunk = 0
    for i, (_, e) in enumerate(live):
        rows, cols, src_ld, rows_pad, dst_ld = e.key
        desc[i] = (e.src_ptr, e.hi, e.lo, rows, cols, src_ld, rows_pad, dst_ld, chunk)
        chunk += (rows_pad * dst_ld + CHUNK - 1) // CHUNK
    t.order = [e for _, e in live]
    t.chunks = chunk
    t.table = torch.from_numpy(desc.view(np.uint8).reshape(-1).copy()).to(device) if len(live) else None
    t.dirty = False


def _capturing():
    try:
        return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()
    except Exception:  # pragma: no cover
        return False


def _owner(param):
    """The long-lived tensor a weight operand belongs to: the parameter itself, or the base of a view of it
    (``conv.weight.reshape(cout, -1)`` is a new tensor object on every call; its base is the parameter)."""
    base = getattr(param, "_base", None)
    return base if base is not None else param


def weight_operands(param, rows, cols, src_ld, rows_pad, dst_ld):
    """(hi pointer, lo pointer, keep-alive tensor, version) of the bf16 pair of ``param`` viewed as
    [rows, cols] with row pitch src_ld, zero-padded to [rows_pad, dst_ld].  ``param`` must be fp32 contiguous
    (a parameter or a view of one)."""
    L = _lib.lib()
    dev = param.device
    t = _table(dev)
    key = (int(rows), int(cols), int(src_ld), int(rows_pad), int(dst_ld))
    owner = _owner(param)
    src_ptr = param.data_ptr()
    offset = src_ptr - owner.data_ptr()
    ekey = (id(owner), offset, key)
    e = t.entries.get(ekey)
    if e is None or e.ref() is not owner:
        e = _Entry()
        e.ref, e.key, e.src_ptr, e.offset = weakref.ref(owner), key, src_ptr, offset
        n = 2 * rows_pad * dst_ld
        nb = (n + 255) & ~255
        e.buf = torch.empty(2 * nb, dtype=torch.uint8, device=dev)
        e.hi, e.lo = e.buf.data_ptr(), e.buf.data_ptr() + nb
        check(L.mvk_split_bf16(src_ptr, rows, cols, src_ld, e.hi, e.lo, rows_pad, dst_ld, stream_ptr()))
        e.version, e.epoch = param._version, _EPOCH[0]
        t.entries[ekey] = e
        t.dirty = True
        return e.hi, e.lo, e.buf, e.version
    if e.version != param._version or e.epoch != _EPOCH[0] or e.src_ptr != src_ptr:
        for o in t.entries.values():  # a parameter whose storage moved (.to(), load) must not be read through the old pointer
            p = o.ref()
            if p is None:
                t.dirty = True  # the parameter is gone: drop its pair at the rebuild below
            elif p.data_ptr() + o.offset != o.src_ptr:
                o.src_ptr = p.data_ptr() + o.offset
                t.dirty = True
        if t.dirty and _capturing():
            # the descriptor table cannot be uploaded inside a stream capture: refresh pair by pair (captured launches)
            for o in t.entries.values():
                p = o.ref()
                if p is not None:
                    r, c, sld, rp, dld = o.key
                    check(L.mvk_split_bf16(o.src_ptr, r, c, sld, o.hi, o.lo, rp, dld, stream_ptr()))
                    o.version, o.epoch = p._version, _EPOCH[0]
            return e.hi, e.lo, e.buf, e.version
        if t.dirty:
            _rebuild(t, dev)
        check(L.mvk_split_bf16_multi(t.table.data_ptr(), len(t.order), t.chunks, stream_ptr()))
        for o in t.order:
            p = o.ref()
            if p is not None:
                o.version, o.epoch = p._version, _EPOCH[0]
    return e.hi, e.lo, e.buf, e.version


def check_unchanged(param, version, what):
    """The saved bf16 pair is only valid for the parameter values of the forward pass."""
    if param._version != version:
        raise RuntimeError(f"{what}: the weight was modified in place between forward and backward "
                           f"(version {version} -> {param._version}); its bf16 operand pair is stale")
