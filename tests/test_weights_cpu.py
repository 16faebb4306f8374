"""Host logic of the bf16 weight-operand registry (_weights.py) without a GPU: the library calls are
stubbed, the bookkeeping is real.  Checks: first sight = single split + registration; an unchanged
parameter costs nothing; ONE multi-tensor refresh when a version counter moves or an optimiser steps
(fused optimisers do not move version counters: the step hook must catch them); dead parameters leave
the table."""
import gc

import torch


def test_weight_registry_bookkeeping(mvk, monkeypatch):
    from mvkpconv_b200 import _weights as W
    calls = []

    class StubLib:
        def mvk_split_bf16(self, *a):
            calls.append(("single", a[1], a[2]))
            return 0

        def mvk_split_bf16_multi(self, *a):
            calls.append(("multi", a[1], a[2]))
            return 0

    table = W._DeviceTable()

    def rebuild_cpu(t, device):  # same bookkeeping as _rebuild, without the device upload
        live = [(k, e) for k, e in t.entries.items() if e.ref() is not None]
        t.entries = dict(live)
        t.order = [e for _, e in live]
        t.chunks = sum((e.key[3] * e.key[4] + W.CHUNK - 1) // W.CHUNK for _, e in live)
        t.table = torch.zeros(max(1, 48 * len(live)), dtype=torch.uint8)
        t.dirty = False

    monkeypatch.setattr(W._lib, "lib", lambda: StubLib())
    monkeypatch.setattr(W, "stream_ptr", lambda: 0)
    monkeypatch.setattr(W, "_table", lambda device: table)
    monkeypatch.setattr(W, "_rebuild", rebuild_cpu)

    a = torch.nn.Parameter(torch.randn(16, 8))
    b = torch.nn.Parameter(torch.randn(24, 16))
    ka, kb = (16, 8, 8, 16, 8), (24, 16, 16, 24, 16)
    W.weight_operands(a, *ka)
    W.weight_operands(b, *kb)
    assert [c[0] for c in calls] == ["single", "single"]
    W.weight_operands(a, *ka)
    assert len(calls) == 2                                   # unchanged: nothing launched
    with torch.no_grad():
        a.add_(1.0)                                          # version counter moves
    W.weight_operands(a, *ka)
    assert calls[-1] == ("multi", 2, 2)                      # ONE launch refreshes both pairs
    W.weight_operands(b, *kb)
    assert len(calls) == 3
    # an optimiser step invalidates through the global step hook, whatever the optimiser does to the versions
    b.grad = torch.zeros_like(b)
    opt = torch.optim.SGD([b], lr=0.1)
    epoch = W._EPOCH[0]
    opt.step()
    assert W._EPOCH[0] == epoch + 1
    W.weight_operands(b, *kb)
    assert calls[-1][0] == "multi" and len(calls) == 4
    # a different view of the same parameter (other padding) registers on its own, next to the first one
    W.weight_operands(b, 24, 16, 16, 32, 16)
    assert calls[-1][0] == "single"
    # a reshaped VIEW of a parameter (a new tensor object on every call) maps to the entry of its base
    c = torch.nn.Parameter(torch.randn(6, 4, 1, 1))
    kc = (6, 4, 4, 6, 8)
    W.weight_operands(c.reshape(6, 4), *kc)
    n = len(calls)
    assert calls[-1][0] == "single"
    W.weight_operands(c.reshape(6, 4), *kc)
    W.weight_operands(c.view(6, 4), *kc)
    assert len(calls) == n                                   # same owner, same offset: nothing launched
    # dead parameters leave the table at the next refresh
    del a, opt, c
    gc.collect()
    W.invalidate()
    W.weight_operands(b, 24, 16, 16, 32, 16)
    assert calls[-1][0] == "multi" and calls[-1][1] == 2     # b's two operand layouts
    assert len(table.entries) == 2
