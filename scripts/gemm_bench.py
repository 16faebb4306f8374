#!/usr/bin/env python
"""Isolated timing of mvk_gemm_bf16x3 on the GEMM shapes of the bench step (development tool).
    python scripts/gemm_bench.py [--reps 5]
Prints per shape: us, achieved GB/s (algorithmic bytes), TFLOP/s, fraction of the mixed roofline."""
import argparse, json, os, statistics, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [  # M, N, K, a_mn, b_mn, split, tag
    (267103, 64, 30, 0, 1, 0, "kpconv fwd L0 cin=2"), (267103, 32, 480, 0, 1, 0, "kpconv fwd L0 32->32"),
    (267103, 480, 32, 0, 0, 0, "kpconv dA L0"), (480, 32, 267103, 1, 1, 74, "kpconv dW L0"),
    (51312, 64, 960, 0, 1, 0, "kpconv fwd L1"), (51312, 960, 64, 0, 0, 0, "kpconv dA L1"), (960, 64, 51312, 1, 1, 37, "kpconv dW L1"),
    (12253, 128, 1920, 0, 1, 0, "kpconv fwd L2"), (12253, 1920, 128, 0, 0, 0, "kpconv dA L2"),
    (3032, 256, 3840, 0, 1, 0, "kpconv fwd L3"), (733, 512, 7680, 0, 1, 0, "kpconv fwd L4"),
    (267103, 32, 64, 0, 0, 0, "unary 64->32 L0"), (267103, 128, 32, 0, 0, 0, "unary 32->128 L0"),
    (267103, 128, 64, 0, 0, 0, "unary 64->128 L0"), (267103, 128, 384, 0, 0, 0, "unary 384->128 L0"),
    (267103, 384, 128, 0, 1, 0, "unary dx 128->384"), (128, 384, 267103, 1, 1, 98, "unary dW 384->128"),
    (267103, 128, 128, 0, 0, 0, "unary 128->128 L0"), (128, 128, 267103, 1, 1, 296, "unary dW 128"),
    (51312, 256, 768, 0, 0, 0, "unary 768->256 L1"), (51312, 768, 256, 0, 1, 0, "unary dx L1"),
    # tensor-bound shapes (whole-scene sweeps / wide layers): where the tensor-pipe fraction is read
    (65536, 512, 7680, 0, 1, 0, "tensor fwd 512->512 64k pts"), (65536, 7680, 512, 0, 0, 0, "tensor dA 512 64k pts"),
    (7680, 512, 65536, 1, 1, 0, "tensor dW 512 64k pts"), (262144, 256, 3840, 0, 1, 0, "tensor fwd 256->256 262k pts"),
    (8192, 8192, 8192, 0, 0, 0, "tensor square 8k"),
]

def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--reps", type=int, default=5); ap.add_argument("--only", default="")
    ap.add_argument("--from-json", default="", help="bench.py --detail N output: time every contraction shape of the step")
    args = ap.parse_args()
    calls = {}
    if args.from_json:
        import re
        d = json.loads(open(args.from_json).read().strip().splitlines()[-1])["detail_us_per_call"]
        SHAPES[:] = []
        for k, (us, n) in d.items():
            m = re.match(r"mvk_gemm_bf16x3\[M=(\d+),N=(\d+),K=(\d+),amn=(\d),bmn=(\d),split=(\d+)\]", k)
            if m:
                M, N, K, amn, bmn, split = map(int, m.groups())
                tag = f"step x{n} ({us:.1f} us in situ)"
                SHAPES.append((M, N, K, amn, bmn, split, tag))
                calls[tag + str((M, N, K, amn, bmn))] = n
    import mvkpconv_b200 as mvk
    from mvkpconv_b200._lib import check, ptr, stream_ptr
    L = mvk._lib.lib()
    pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm, tc = pk["hbm_gbs"] * 1e9, pk["bf16_tflops_sustained"] * 1e12
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    # reference points: pure-write and copy bandwidth of this GPU through torch
    if args.only:
        big = torch.empty(1, device="cuda")
    big = torch.empty(128 << 20, dtype=torch.float32, device="cuda")  # 512 MB
    big2 = torch.empty_like(big)
    for name, fn, nb in (("fill 512MB", lambda: big.fill_(1.0), big.numel() * 4), ("copy 512MB", lambda: big2.copy_(big), big.numel() * 8)):
        ts = []
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            if rep: ts.append(e0.elapsed_time(e1) * 1e3)
        print(f"{name}: {statistics.median(ts):.1f} us -> {nb / statistics.median(ts) / 1e3:.0f} GB/s", flush=True)
    del big, big2
    r8 = lambda v: (v + 7) // 8 * 8
    tot = 0.0
    for (M, N, K, amn, bmn, split, tag) in SHAPES:
        if args.only and args.only not in tag:
            continue
        a_shape = (K, r8(M)) if amn else (M, r8(K))
        b_shape = (K, r8(N)) if bmn else (N, r8(K))
        ah = torch.randn(a_shape, device="cuda").bfloat16(); al = (torch.randn(a_shape, device="cuda") * 1e-3).bfloat16()
        bh = torch.randn(b_shape, device="cuda").bfloat16(); bl = (torch.randn(b_shape, device="cuda") * 1e-3).bfloat16()
        D = torch.zeros((M, r8(N)), device="cuda")
        ts = []
        for rep in range(args.reps + 1):
            flush.fill_(rep)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(L.mvk_gemm_bf16x3(ptr(ah), ptr(al), amn, a_shape[1], ptr(bh), ptr(bl), bmn, b_shape[1], M, N, K, ptr(D),
                                    D.shape[1], N, 3, split, stream_ptr()))
            e1.record(); torch.cuda.synchronize()
            if rep: ts.append(e0.elapsed_time(e1) * 1e3)
        us = statistics.median(ts)
        nbytes = 4.0 * (M * K + K * N) + 4.0 * M * N
        flops = 6.0 * M * N * K
        ideal = max(nbytes / hbm, flops / tc) * 1e6
        tot += us * calls.get(tag + str((M, N, K, amn, bmn)), 1)
        print(f"{us:8.1f} us  ideal {ideal:7.1f}  frac {ideal/us:5.2f}  {nbytes/us/1e3:7.0f} GB/s {flops/us/1e6:7.1f} TF/s  M={M} N={N} K={K} amn={amn} bmn={bmn} split={split}  {tag}", flush=True)
    print("total us", round(tot, 1))

if __name__ == "__main__":
    main()
