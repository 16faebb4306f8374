#!/usr/bin/env python
"""Isolated timing of mvk_bn_batch_stats / mvk_act_bwd_reduce on the shapes of the bench step (development tool)."""
import os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mvkpconv_b200 as mvk
from mvkpconv_b200._lib import check, ptr, stream_ptr
L = mvk._lib.lib()
SHAPES = [(267103, 128), (267103, 64), (267103, 32), (51312, 256), (51312, 64), (12253, 512), (12253, 128), (3032, 1024), (3032, 256), (733, 2048), (733, 512)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, reps=7, cold=True):
    ts = []
    for rep in range(reps + 1):
        if cold: flush.fill_(rep)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if rep: ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)
x1 = torch.zeros(1, device="cuda")
print("empty-ish launch (fill 1 elem): %.1f us" % timeit(lambda: x1.fill_(1.0)))
for rows, cols in SHAPES:
    y = torch.randn(rows, cols, device="cuda")
    stats = torch.zeros(2 * cols + 1, dtype=torch.float64, device="cuda")
    g = torch.ones(cols, device="cuda"); b = torch.zeros(cols, device="cuda")
    rm = torch.zeros(cols, device="cuda"); rv = torch.ones(cols, device="cuda")
    sc = torch.empty(cols, device="cuda"); sh = torch.empty(cols, device="cuda"); mu = torch.empty(cols, device="cuda"); iv = torch.empty(cols, device="cuda")
    def run():
        stats.zero_()
        check(L.mvk_bn_batch_stats(ptr(y), rows, cols, cols, ptr(stats), ptr(g), ptr(b), 1e-5, 0.1, ptr(rm), ptr(rv), ptr(sc), ptr(sh), ptr(mu), ptr(iv), None, stream_ptr()))
    def zero_only():
        stats.zero_()
    tz = timeit(zero_only)
    tc, tw = timeit(run) - tz, timeit(run, cold=False) - tz
    nb = rows * cols * 4
    print(f"bn_batch_stats [{rows}x{cols}] cold {tc:6.1f} us ({nb/tc/1e3:6.0f} GB/s)  warm {tw:6.1f} us ({nb/tw/1e3:6.0f} GB/s)", flush=True)
