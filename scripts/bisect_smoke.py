"""Stage-by-stage bisect of the smoke() KPConv 64->128 case (round-1 driver smoke failure).

Runs every stage of the bf16x3 path through the C ABI and compares each one with an fp64 product of
exactly the operands it consumed, so a wrong stage is named instead of inferred.
    python scripts/bisect_smoke.py            (MVK_PDL=0 python scripts/bisect_smoke.py for the PDL variant)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mvkpconv_b200 as mvk  # noqa: E402
from mvkpconv_b200._lib import check, ptr, stream_ptr  # noqa: E402
from oracle import geom, modules  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max())


def main():
    L = mvk._lib.lib()
    rng = np.random.default_rng(0)
    n = 4000
    xy = rng.uniform(-1, 1, (n, 2))
    pts = np.concatenate([xy, (0.1 * np.sin(4 * xy[:, :1]) + rng.normal(0, 0.01, (n, 1)))], 1).astype(np.float32)
    lens = np.array([1500, 2500], np.int32)
    radius = 0.12
    inds_np = geom.batch_neighbors(pts, pts, lens, lens, radius)
    print("H =", inds_np.shape[1], "PDL =", os.environ.get("MVK_PDL", "1"))
    cin, cout, K = 64, 128, 15
    np.random.seed(0)
    torch.manual_seed(0)
    conv = mvk.KPConv(K, 3, cin, cout, radius * 1.2 / 2.5, radius).cuda()
    x = torch.randn(n, cin)
    go = torch.randn(n, cout)
    xo = x.clone().requires_grad_(True)
    wo = conv.weights.detach().cpu().clone().requires_grad_(True)
    ref = modules.kpconv_forward(torch.from_numpy(pts), torch.from_numpy(pts), torch.from_numpy(inds_np).long(), xo,
                                 conv.kernel_points.detach().cpu(), wo, conv.KP_extent)
    ref.backward(go)

    q = torch.from_numpy(pts).cuda()
    inds = torch.from_numpy(inds_np).cuda().long()
    h = inds.shape[1]
    xg = x.cuda()
    kp = conv.kernel_points.detach().contiguous()
    W = conv.weights.detach().reshape(K * cin, cout).contiguous()
    kd = K * cin
    st = stream_ptr()

    # stage A: fp32 and hi/lo outputs
    A32 = torch.empty(n, kd, device="cuda")
    check(L.mvk_kpconv_weighted(ptr(q), n, ptr(q), n, ptr(inds), 1, h, ptr(xg), cin, ptr(kp), K, float(conv.KP_extent), 1, 0,
                                kd, ptr(A32), None, None, st))
    a_hi = torch.empty(n, kd, dtype=torch.bfloat16, device="cuda")
    a_lo = torch.empty_like(a_hi)
    check(L.mvk_kpconv_weighted(ptr(q), n, ptr(q), n, ptr(inds), 1, h, ptr(xg), cin, ptr(kp), K, float(conv.KP_extent), 1, 0,
                                kd, None, ptr(a_hi), ptr(a_lo), st))
    print("stage A  hi+lo vs f32        :", rel(a_hi.double() + a_lo.double(), A32))
    # oracle stage A (torch CPU fp32 of the reference formula)
    out_ref64 = A32.double() @ W.double()
    print("stage A x W (fp64) vs oracle :", rel(out_ref64, ref))

    w_hi = torch.empty(kd, cout, dtype=torch.bfloat16, device="cuda")
    w_lo = torch.empty_like(w_hi)
    check(L.mvk_split_bf16(ptr(W), kd, cout, cout, ptr(w_hi), ptr(w_lo), kd, cout, st))
    print("W split  hi+lo vs f32        :", rel(w_hi.double() + w_lo.double(), W))
    wh2, wl2, keep, _ = mvk._weights.weight_operands(conv.weights.detach().contiguous(), kd, cout, cout, kd, cout)
    torch.cuda.synchronize()
    import ctypes
    buf = keep.cpu().numpy()
    nb = (2 * kd * cout + 255) & ~255
    hi2 = torch.from_numpy(buf[:2 * kd * cout].copy()).view(torch.bfloat16).reshape(kd, cout)
    lo2 = torch.from_numpy(buf[nb:nb + 2 * kd * cout].copy()).view(torch.bfloat16).reshape(kd, cout)
    print("weight_operands hi+lo vs f32 :", rel(hi2.double() + lo2.double(), W))

    exact = (a_hi.double() + a_lo.double()) @ (w_hi.double() + w_lo.double())
    for split_k in (1, 2, 3, 5, 0):
        D = torch.zeros(n, cout, device="cuda")
        check(L.mvk_gemm_bf16x3(ptr(a_hi), ptr(a_lo), 0, kd, ptr(w_hi), ptr(w_lo), 1, cout, n, cout, kd, ptr(D), cout, cout,
                                3, split_k, st))
        torch.cuda.synchronize()
        print(f"fwd GEMM split_k={split_k}: vs exact operands {rel(D, exact):.3e}   vs oracle {rel(D, ref):.3e}")
        if split_k == 0:
            err = (D.double() - exact).abs().cpu()
            r, c = np.unravel_index(int(err.argmax()), err.shape)
            print("   worst element row/col:", r, c, "rows with err>1e-4*max:",
                  int((err.max(1).values > 1e-4 * exact.abs().max().cpu()).sum()))
    for terms in (1,):
        D = torch.zeros(n, cout, device="cuda")
        check(L.mvk_gemm_bf16x3(ptr(a_hi), ptr(a_lo), 0, kd, ptr(w_hi), ptr(w_lo), 1, cout, n, cout, kd, ptr(D), cout, cout,
                                terms, 1, st))
        print(f"fwd GEMM terms={terms}: vs exact {rel(D, exact):.3e}")

    # dA = go W^T
    gog = go.cuda()
    go_hi = torch.empty(n, cout, dtype=torch.bfloat16, device="cuda")
    go_lo = torch.empty_like(go_hi)
    check(L.mvk_split_bf16(ptr(gog), n, cout, cout, ptr(go_hi), ptr(go_lo), n, cout, st))
    dA = torch.empty(n, kd, device="cuda")
    check(L.mvk_gemm_bf16x3(ptr(go_hi), ptr(go_lo), 0, cout, ptr(w_hi), ptr(w_lo), 0, cout, n, kd, cout, ptr(dA), kd, kd, 3, 0, st))
    dA_exact = gog.double() @ W.double().t()
    print("dA GEMM vs fp64              :", rel(dA, dA_exact))
    # dW = A^T go
    gw = torch.empty(kd, cout, device="cuda")
    check(L.mvk_gemm_bf16x3(ptr(a_hi), ptr(a_lo), 1, kd, ptr(go_hi), ptr(go_lo), 1, cout, kd, cout, n, ptr(gw), cout, cout, 3, 0, st))
    print("dW GEMM vs fp64              :", rel(gw, A32.double().t() @ gog.double()), " vs oracle", rel(gw.reshape(K, cin, cout), wo.grad))
    # dX scatter from the exact dA
    dA32 = dA_exact.float().contiguous()
    gx = torch.zeros(n, cin, device="cuda")
    check(L.mvk_kpconv_weighted_bwd(ptr(q), n, ptr(q), n, ptr(inds), 1, h, cin, ptr(kp), K, float(conv.KP_extent), 1, 0,
                                    ptr(dA32), kd, ptr(gx), st))
    print("stage A bwd (exact dA) vs oracle:", rel(gx, xo.grad))

    # whole module, every contraction
    for contraction in ("fp32", "bf16x3"):
        conv.contraction = contraction
        conv.weights.grad = None
        xx = x.cuda().requires_grad_(True)
        out = conv(q, q, inds, xx)
        out.backward(gog)
        print(f"module {contraction:7s}: out {rel(out, ref):.3e}  grad_x {rel(xx.grad, xo.grad):.3e}  grad_w {rel(conv.weights.grad, wo.grad):.3e}")
    try:
        oracle64(conv, x, go, pts, inds_np, ref, xo, wo, q, inds, gog)
    except Exception as e:  # noqa: BLE001
        print("oracle fp64 leg failed:", repr(e))


def oracle64(conv, x, go, pts, inds_np, ref, xo, wo, q, inds, gog):
    xo64 = x.double().clone().requires_grad_(True)
    wo64 = conv.weights.detach().cpu().double().clone().requires_grad_(True)
    ref64 = modules.kpconv_forward(torch.from_numpy(pts).double(), torch.from_numpy(pts).double(), torch.from_numpy(inds_np).long(), xo64,
                                   conv.kernel_points.detach().cpu().double(), wo64, conv.KP_extent)
    ref64.backward(go.double())
    print(f"oracle fp32 vs oracle fp64: out {rel(ref, ref64):.3e} grad_x {rel(xo.grad, xo64.grad):.3e} grad_w {rel(wo.grad, wo64.grad):.3e}")
    conv.contraction = "bf16x3"
    xx = x.cuda().requires_grad_(True)
    conv.weights.grad = None
    out = conv(q, q, inds, xx)
    out.backward(gog)
    print(f"module bf16x3 vs oracle fp64: out {rel(out, ref64):.3e}  grad_x {rel(xx.grad, xo64.grad):.3e}  grad_w {rel(conv.weights.grad, wo64.grad):.3e}")


if __name__ == "__main__":
    main()
