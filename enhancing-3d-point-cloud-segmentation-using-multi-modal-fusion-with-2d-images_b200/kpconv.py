"""KPConv (rigid) as a drop-in ``nn.Module`` backed by libmvk's hand-written sm_100a kernels.

Mirror of the reference operator (KPConv-PyTorch/models/blocks.py:143-379): same constructor
signature, same ``forward(q_pts, s_pts, neighb_inds, x)``, same parameter names / shapes
(``weights`` [K, Cin, Cout], ``kernel_points`` [K, 3] with requires_grad=False) so reference
checkpoints load, same ``__repr__``.

forward  =  stage A  mvk_kpconv_weighted  (neighbour gather + kernel-point influence -> [N, K*Cin])
          + stage B  contraction with W [K*Cin, Cout]:
                "bf16x3" (default)  tcgen05/TMA tensor cores, hi/lo split bf16, fp32 accumulate
                "bf16"              tcgen05/TMA tensor cores, plain bf16 operands (stated separately)
                "fp32"              strict fp32 FFMA contraction
backward =  dA = dOut W^T (same contraction kernel),  dW = A^T dOut (split-K, fp32 atomics),
            dX = scatter of dA through the influences (mvk_kpconv_weighted_bwd, fp32 atomics).

Deformable / modulated KPConv (blocks.py:243-325): the offsets come from a rigid KPConv of this
package, the deformed stage A (per-point kernel points, min_d2 for the regulariser, gradients to
the kernel points and modulations) runs mvk_kpconv_deform_weighted[_bwd].
"""
import math
import os

import torch
import torch.nn as nn
from torch.nn.init import kaiming_uniform_
from torch.nn.parameter import Parameter

from . import _lib, _weights
from ._lib import check, ptr, stream_ptr
from .kernel_points import load_kernels

_INFLUENCE = {"constant": 0, "linear": 1, "gaussian": 2}
_AGGREGATION = {"sum": 0, "closest": 1}
CONTRACTIONS = ("bf16x3", "bf16", "fp32")

DEFAULT_CONTRACTION = os.environ.get("MVK_CONTRACTION", "bf16x3")
# MVK_FUSED=1: forward = ONE kernel (stage A + tcgen05 contraction, the weighted operand stays in shared memory) on the
# shapes mvk_kpconv_fused supports.  Parity-tested, but measured SLOWER than the two-kernel sequence on B200 (level 0:
# 868 vs 652 us; the K-outer gather re-derives the influences per 64-column block and executes 2.2x the instructions of
# kp_fwd_fast, profiles/r2h_fused_summary.md), so it is opt-in.
FUSED_FORWARD = os.environ.get("MVK_FUSED", "0") == "1"


def _round_up(a, b):
    return (a + b - 1) // b * b


def _idx(t):
    if t.dtype not in (torch.int64, torch.int32):
        raise RuntimeError("neighb_inds must be int64 (reference dtype) or int32")
    return t.contiguous(), (1 if t.dtype == torch.int64 else 0)


def _split_k_for(m_tiles, n_tiles, kb_total):
    target = 2 * 148
    return max(1, min(kb_total, target // max(1, m_tiles * n_tiles)))


def _f32c(t):
    if t.dtype is not torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


_SPLIT_PERM = {}


def _channel_split(cin, K, h, influence, aggregation, device):
    """(c_rest, c_main, perm) when cin = c_rest + c_main with 1 <= c_rest <= 4 leading channels and a fast-path width
    c_main behind them (66 = 2 + 64, 65 = 1 + 64, 68 = 4 + 64, ...), else None.  perm[j] = row of W [K*cin, cout] that
    column j of the K-concatenated operand [A_main (k, c_main) | A_rest (k, c_rest)] multiplies."""
    c_rest = cin % 32
    c_main = cin - c_rest
    if not (1 <= c_rest <= 4 and (c_main in (32, 64) or (c_main > 0 and c_main % 128 == 0))):
        return None
    if influence != 1 or aggregation != 0 or K > 16 or h > 64:
        return None
    key = (cin, K, device.index)
    perm = _SPLIT_PERM.get(key)
    if perm is None:
        k = torch.arange(K).unsqueeze(1)
        main = (k * cin + c_rest + torch.arange(c_main).unsqueeze(0)).reshape(-1)
        rest = (k * cin + torch.arange(c_rest).unsqueeze(0)).reshape(-1)
        perm = _SPLIT_PERM[key] = torch.cat([main, rest]).to(device)
    return c_rest, c_main, perm


def _carve(device, *sizes):
    """ONE device allocation carved into 256-byte aligned raw sub-buffers -> (keep-alive tensor, [ptr | None])."""
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (int(n) + 255) & ~255
    t = torch.empty(max(total, 256), dtype=torch.uint8, device=device)
    base = t.data_ptr()
    return t, [(base + o) if n else None for o, n in zip(offs, sizes)]


class _KPConvFunction(torch.autograd.Function):
    """out[i] = sum_k (sum_h w_ihk x[j_ih]) @ W[k]   (blocks.py:277-374, rigid)."""

    @staticmethod
    def forward(ctx, q_pts, s_pts, neighb_inds, x, weights, kernel_points, kp_extent, influence,
                aggregation, contraction):
        _lib.require_cuda()
        L = _lib.lib()
        if not x.is_cuda:
            raise RuntimeError("KPConv: tensors must live on a CUDA device (no CPU fallback)")
        q, s = _f32c(q_pts), _f32c(s_pts)
        inds, is64 = _idx(neighb_inds)
        xf, w, kp = _f32c(x), _f32c(weights), _f32c(kernel_points)
        nq, ns, h = q.shape[0], s.shape[0], inds.shape[1]
        K, cin, cout = w.shape
        kd = K * cin
        dev = xf.device
        out = torch.empty((nq, cout), dtype=torch.float32, device=dev)
        st = stream_ptr()
        fp32 = contraction == "fp32"
        with _lib.on_device(dev):
            if fp32:
                ld, npad = kd, cout
                keep, (A, _a, _b, _c) = _carve(dev, 4 * nq * ld, 0, 0, 0)
                check(L.mvk_kpconv_weighted(q.data_ptr(), nq, s.data_ptr(), ns, inds.data_ptr(), is64, h, xf.data_ptr(), cin,
                                            kp.data_ptr(), K, float(kp_extent), influence, aggregation, ld, A, None, None, st))
                if nq > 0:
                    check(L.mvk_gemm_f32(A, ld, 1, w.data_ptr(), cout, 1, nq, cout, kd, out.data_ptr(), cout, 1, st))
                ptrs = (A, None, None, None)
            else:
                terms = 3 if contraction == "bf16x3" else 1
                ld = _round_up(kd, 8)      # 16-byte row pitch is all TMA needs: partial tiles are zero-filled
                npad = _round_up(cout, 8)
                w_hi, w_lo, wkeep, _ = _weights.weight_operands(w, kd, cout, cout, ld, npad)  # once per optimiser step
                # the weighted operand is only kept (written once, never re-read in forward) when a backward will
                # contract it again for dW
                need_a = bool(ctx.needs_input_grad[4])  # (all False under torch.no_grad(): inference keeps nothing)
                fused = (FUSED_FORWARD and terms == 3 and nq > 0 and
                         L.mvk_kpconv_fused_supported(cin, cout, K, h, influence, aggregation) == 1)
                split = None if fused else _channel_split(cin, K, h, influence, aggregation, dev)
                if fused:
                    keep, (a_hi, a_lo) = _carve(dev, 2 * nq * ld if need_a else 0, 2 * nq * ld if need_a else 0)
                    check(L.mvk_kpconv_fused(q.data_ptr(), nq, s.data_ptr(), ns, inds.data_ptr(), is64, h, xf.data_ptr(), cin,
                                             kp.data_ptr(), K, float(kp_extent), cout, w_hi, w_lo, npad, out.data_ptr(),
                                             a_hi, a_lo, ld, st))
                elif split is not None:
                    # Cin = c_rest + c_main (e.g. 66 = 2 + 64, the first layer of the early-fusion net): the wide part
                    # runs the fast stage-A kernel, the narrow part the small-Cin kernel, both into ONE K-concatenated
                    # operand [A_main | A_rest | 0]; the weight rows are permuted to match (KPConv is linear in x)
                    c_rest, c_main, perm = split
                    keep, (a_hi, a_lo) = _carve(dev, 2 * nq * ld, 2 * nq * ld)
                    x_main = xf[:, c_rest:].contiguous()
                    x_rest = xf[:, :c_rest].contiguous()
                    km = K * c_main
                    check(L.mvk_kpconv_weighted_part(q.data_ptr(), nq, s.data_ptr(), ns, inds.data_ptr(), is64, h,
                                                     x_main.data_ptr(), c_main, kp.data_ptr(), K, float(kp_extent), influence,
                                                     aggregation, ld, km, None, a_hi, a_lo, st))
                    check(L.mvk_kpconv_weighted_part(q.data_ptr(), nq, s.data_ptr(), ns, inds.data_ptr(), is64, h,
                                                     x_rest.data_ptr(), c_rest, kp.data_ptr(), K, float(kp_extent), influence,
                                                     aggregation, ld, ld - km, None, a_hi + 2 * km, a_lo + 2 * km, st))
                    w_perm = w.reshape(kd, cout).index_select(0, perm)
                    wkeep = torch.empty(2 * ((2 * ld * npad + 255) & ~255), dtype=torch.uint8, device=dev)
                    w_hi, w_lo = wkeep.data_ptr(), wkeep.data_ptr() + ((2 * ld * npad + 255) & ~255)
                    check(L.mvk_split_bf16(w_perm.data_ptr(), kd, cout, cout, w_hi, w_lo, ld, npad, st))
                    if nq > 0:
                        check(L.mvk_gemm_bf16x3(a_hi, a_lo, 0, ld, w_hi, w_lo, 1, npad, nq, npad, ld, out.data_ptr(), cout,
                                                cout, terms, 0, st))
                else:
                    keep, (a_hi, a_lo) = _carve(dev, 2 * nq * ld, 2 * nq * ld)
                    check(L.mvk_kpconv_weighted(q.data_ptr(), nq, s.data_ptr(), ns, inds.data_ptr(), is64, h, xf.data_ptr(),
                                                cin, kp.data_ptr(), K, float(kp_extent), influence, aggregation, ld, None,
                                                a_hi, a_lo, st))
                    if nq > 0:
                        check(L.mvk_gemm_bf16x3(a_hi, a_lo, 0, ld, w_hi, w_lo, 1, npad, nq, npad, ld, out.data_ptr(), cout,
                                                cout, terms, 0, st))
                ptrs = (a_hi, a_lo, w_hi, w_lo)
        ctx.save_for_backward(q, s, inds, kp, w, keep)  # autograd's version check on w also guards its bf16 pair
        ctx.cfg = (nq, ns, h, K, cin, cout, float(kp_extent), influence, aggregation, contraction, is64, ld, npad, ptrs)
        ctx.wkeep = None if fp32 else wkeep
        ctx.split = None if fp32 else split
        return out

    @staticmethod
    def backward(ctx, grad_out):
        L = _lib.lib()
        q, s, inds, kp, w, keep = ctx.saved_tensors
        nq, ns, h, K, cin, cout, extent, influence, aggregation, contraction, is64, ld, npad, ptrs = ctx.cfg
        kd = K * cin
        dev = grad_out.device
        go = _f32c(grad_out)
        need_x, need_w = ctx.needs_input_grad[3], ctx.needs_input_grad[4]
        gx = gw = None
        st = stream_ptr()
        with _lib.on_device(dev):
            if contraction == "fp32":
                A = ptrs[0]
                bkeep, (dA, _g0, _g1) = _carve(dev, 4 * nq * ld if need_x else 0, 0, 0)
                if need_w:
                    gw = torch.zeros((K, cin, cout), dtype=torch.float32, device=dev)
                    if nq > 0:
                        split = max(1, min(nq // 64, 2 * 148 // max(1, ((kd + 63) // 64) * ((cout + 63) // 64))))
                        # dW[kd, cout] = A^T dOut : A'(m, k) = A[k*ld + m]
                        check(L.mvk_gemm_f32(A, 1, ld, go.data_ptr(), cout, 1, kd, cout, nq, gw.data_ptr(), cout, split, st))
                if need_x and nq > 0:
                    # dA[nq, kd] = dOut W^T : B(k=o, n=kd) = W[n*cout + k]
                    check(L.mvk_gemm_f32(go.data_ptr(), cout, 1, w.data_ptr(), 1, cout, nq, kd, cout, dA, ld, 1, st))
            else:
                a_hi, a_lo, w_hi, w_lo = ptrs
                terms = 3 if contraction == "bf16x3" else 1
                cached = getattr(grad_out, "_mvk_hilo", None)  # written by the producer of grad_out (bn_act backward)
                if cached is not None and (cached.version != grad_out._version or cached.rows != nq or cached.ld != npad):
                    cached = None
                nsplit = 0 if cached is not None else 2 * nq * npad
                bkeep, (dA, go_hi, go_lo) = _carve(dev, 4 * nq * ld if need_x else 0, nsplit, nsplit)
                if cached is not None:
                    go_hi, go_lo, gokeep = cached.hi, cached.lo, cached.keep
                else:
                    check(L.mvk_split_bf16(go.data_ptr(), nq, cout, cout, go_hi, go_lo, nq, npad, st))
                if need_w:
                    if nq > 0:
                        # dW = A^T dOut, both operands MN-major, reduction over the points; split_k = 0: the library
                        # splits the reduction over the SMs and zeroes dW itself
                        gw = torch.empty((K, cin, cout), dtype=torch.float32, device=dev)
                        check(L.mvk_gemm_bf16x3(a_hi, a_lo, 1, ld, go_hi, go_lo, 1, npad, kd, npad, nq, gw.data_ptr(), cout,
                                                cout, terms, 0, st))
                    else:
                        gw = torch.zeros((K, cin, cout), dtype=torch.float32, device=dev)
                if need_x and nq > 0:
                    # dA = dOut W^T : A = dOut [nq, npad] K-major, B = W [ld, npad] K-major
                    check(L.mvk_gemm_bf16x3(go_hi, go_lo, 0, npad, w_hi, w_lo, 0, npad, nq, ld, npad, dA, ld, ld, terms,
                                            0, st))
            split = getattr(ctx, "split", None)
            if need_x and split is not None:
                c_rest, c_main, perm = split
                km = K * c_main
                gx_main = torch.zeros((ns, c_main), dtype=torch.float32, device=dev)
                gx_rest = torch.zeros((ns, c_rest), dtype=torch.float32, device=dev)
                if nq > 0:
                    check(L.mvk_kpconv_weighted_bwd(q.data_ptr(), nq, s.data_ptr(), ns, inds.data_ptr(), is64, h, c_main,
                                                    kp.data_ptr(), K, extent, influence, aggregation, dA, ld,
                                                    gx_main.data_ptr(), st))
                    check(L.mvk_kpconv_weighted_bwd(q.data_ptr(), nq, s.data_ptr(), ns, inds.data_ptr(), is64, h, c_rest,
                                                    kp.data_ptr(), K, extent, influence, aggregation, dA + 4 * km, ld,
                                                    gx_rest.data_ptr(), st))
                gx = torch.cat([gx_rest, gx_main], dim=1)
            elif need_x:
                gx = torch.zeros((ns, cin), dtype=torch.float32, device=dev)
                if nq > 0:
                    check(L.mvk_kpconv_weighted_bwd(q.data_ptr(), nq, s.data_ptr(), ns, inds.data_ptr(), is64, h, cin,
                                                    kp.data_ptr(), K, extent, influence, aggregation, dA, ld, gx.data_ptr(),
                                                    st))
            if gw is not None and split is not None:
                # rows of dW are in the operand's order [main | rest]: back to the parameter's (k, c) order
                gw = torch.zeros_like(gw).view(kd, cout).index_copy_(0, split[2], gw.view(kd, cout)).view(K, cin, cout)
        return None, None, None, gx, gw, None, None, None, None, None


class _KPConvDeformFunction(torch.autograd.Function):
    """Deformable KPConv (blocks.py:243-374): per-point kernel points `deformed_kp` [N, K, 3] and
    optional modulations [N, K].  Returns (out [N, Cout], min_d2 [N, K]); gradients flow to x, W,
    deformed_kp and modulations (and from min_d2 back to deformed_kp for the fitting regulariser)."""

    @staticmethod
    def forward(ctx, q_pts, s_pts, neighb_inds, x, weights, deformed_kp, modulations, kp_extent, influence,
                aggregation, contraction):
        _lib.require_cuda()
        L = _lib.lib()
        if not x.is_cuda:
            raise RuntimeError("KPConv: tensors must live on a CUDA device (no CPU fallback)")
        q = q_pts.detach().contiguous().float()
        s = s_pts.detach().contiguous().float()
        inds, is64 = _idx(neighb_inds)
        xf = x.detach().contiguous().float()
        w = weights.detach().contiguous().float()
        kp = deformed_kp.detach().contiguous().float()
        mod = None if modulations is None else modulations.detach().contiguous().float()
        nq, ns, h = q.shape[0], s.shape[0], inds.shape[1]
        K, cin, cout = w.shape
        kd = K * cin
        dev = xf.device
        out = torch.empty((nq, cout), dtype=torch.float32, device=dev)
        min_d2 = torch.empty((nq, K), dtype=torch.float32, device=dev)
        argmin = torch.empty((nq, K), dtype=torch.int32, device=dev)
        st = stream_ptr()
        with _lib.on_device(dev):
            if contraction == "fp32":
                ld = kd
                A = torch.empty((nq, ld), dtype=torch.float32, device=dev)
                check(L.mvk_kpconv_deform_weighted(ptr(q), nq, ptr(s), ns, ptr(inds), is64, h, ptr(xf), cin, ptr(kp),
                                                   ptr(mod), K, float(kp_extent), influence, aggregation, ld, ptr(A),
                                                   None, None, ptr(min_d2), ptr(argmin), st))
                if nq > 0:
                    check(L.mvk_gemm_f32(ptr(A), ld, 1, ptr(w), cout, 1, nq, cout, kd, ptr(out), cout, 1, st))
                saved = (A,)
            else:
                terms = 3 if contraction == "bf16x3" else 1
                ld, npad = _round_up(kd, 8), _round_up(cout, 8)
                a_hi = torch.empty((nq, ld), dtype=torch.bfloat16, device=dev)
                a_lo = torch.empty((nq, ld), dtype=torch.bfloat16, device=dev)
                check(L.mvk_kpconv_deform_weighted(ptr(q), nq, ptr(s), ns, ptr(inds), is64, h, ptr(xf), cin, ptr(kp),
                                                   ptr(mod), K, float(kp_extent), influence, aggregation, ld, None,
                                                   ptr(a_hi), ptr(a_lo), ptr(min_d2), ptr(argmin), st))
                w_hi = torch.empty((ld, npad), dtype=torch.bfloat16, device=dev)
                w_lo = torch.empty((ld, npad), dtype=torch.bfloat16, device=dev)
                check(L.mvk_split_bf16(ptr(w), kd, cout, cout, ptr(w_hi), ptr(w_lo), ld, npad, st))
                if nq > 0:
                    check(L.mvk_gemm_bf16x3(ptr(a_hi), ptr(a_lo), 0, ld, ptr(w_hi), ptr(w_lo), 1, npad,
                                            nq, npad, ld, ptr(out), cout, cout, terms, 0, st))
                saved = (a_hi, a_lo, w_hi, w_lo)
        ctx.save_for_backward(q, s, inds, xf, kp, mod if mod is not None else kp.new_empty(0), argmin, w, *saved)
        ctx.cfg = (nq, ns, h, K, cin, cout, float(kp_extent), influence, aggregation, contraction, is64, mod is not None)
        ctx.mark_non_differentiable(argmin)
        return out, min_d2

    @staticmethod
    def backward(ctx, grad_out, grad_min_d2):
        L = _lib.lib()
        q, s, inds, xf, kp, mod, argmin, w, *saved = ctx.saved_tensors
        nq, ns, h, K, cin, cout, extent, influence, aggregation, contraction, is64, has_mod = ctx.cfg
        kd = K * cin
        dev = grad_out.device
        go = grad_out.detach().contiguous().float()
        gmin = None if grad_min_d2 is None else grad_min_d2.detach().contiguous().float()
        need_x, need_w = ctx.needs_input_grad[3], ctx.needs_input_grad[4]
        need_kp, need_mod = ctx.needs_input_grad[5], ctx.needs_input_grad[6] and has_mod
        gx = gw = gkp = gmod = None
        st = stream_ptr()
        with _lib.on_device(dev):
            if contraction == "fp32":
                (A,) = saved
                ld = kd
                if need_w:
                    gw = torch.zeros((K, cin, cout), dtype=torch.float32, device=dev)
                    if nq > 0:
                        split = max(1, min(nq // 64, 2 * 148 // max(1, ((kd + 63) // 64) * ((cout + 63) // 64))))
                        check(L.mvk_gemm_f32(ptr(A), 1, ld, ptr(go), cout, 1, kd, cout, nq, ptr(gw), cout, split, st))
                dA = torch.empty((nq, ld), dtype=torch.float32, device=dev)
                if nq > 0:
                    check(L.mvk_gemm_f32(ptr(go), cout, 1, ptr(w), 1, cout, nq, kd, cout, ptr(dA), ld, 1, st))
            else:
                a_hi, a_lo, w_hi, w_lo = saved
                terms = 3 if contraction == "bf16x3" else 1
                ld, npad = a_hi.shape[1], w_hi.shape[1]
                go_hi = torch.empty((nq, npad), dtype=torch.bfloat16, device=dev)
                go_lo = torch.empty((nq, npad), dtype=torch.bfloat16, device=dev)
                check(L.mvk_split_bf16(ptr(go), nq, cout, cout, ptr(go_hi), ptr(go_lo), nq, npad, st))
                if need_w:
                    gw = torch.zeros((K, cin, cout), dtype=torch.float32, device=dev)
                    if nq > 0:
                        kb_total = (nq + 63) // 64
                        split = _split_k_for((kd + 127) // 128, (npad + 127) // 128, kb_total)
                        check(L.mvk_gemm_bf16x3(ptr(a_hi), ptr(a_lo), 1, ld, ptr(go_hi), ptr(go_lo), 1, npad,
                                                kd, npad, nq, ptr(gw), cout, cout, terms, split, st))
                dA = torch.empty((nq, ld), dtype=torch.float32, device=dev)
                if nq > 0:
                    check(L.mvk_gemm_bf16x3(ptr(go_hi), ptr(go_lo), 0, npad, ptr(w_hi), ptr(w_lo), 0, npad,
                                            nq, ld, npad, ptr(dA), ld, ld, terms, 0, st))
            if need_x:
                gx = torch.zeros((ns, cin), dtype=torch.float32, device=dev)
            if need_kp:
                gkp = torch.empty((nq, K, 3), dtype=torch.float32, device=dev)
            if need_mod:
                gmod = torch.empty((nq, K), dtype=torch.float32, device=dev)
            if need_x or need_kp or need_mod:
                check(L.mvk_kpconv_deform_weighted_bwd(ptr(q), nq, ptr(s), ns, ptr(inds), is64, h, ptr(xf), cin, ptr(kp),
                                                       ptr(mod) if has_mod else None, K, extent, influence, aggregation,
                                                       ptr(dA), ld, ptr(gmin), ptr(argmin), ptr(gx), ptr(gkp), ptr(gmod),
                                                       st))
        return None, None, None, gx, gw, gkp, gmod, None, None, None, None


class KPConv(nn.Module):

    def __init__(self, kernel_size, p_dim, in_channels, out_channels, KP_extent, radius,
                 fixed_kernel_points='center', KP_influence='linear', aggregation_mode='sum',
                 deformable=False, modulated=False, contraction=None):
        """Same positional arguments and defaults as the reference constructor (blocks.py:145-147):

        kernel_size          K, how many kernel points carry a weight matrix
        p_dim                dimension of the point coordinates (3 here)
        in_channels / out_channels   feature widths Cin / Cout
        KP_extent            reach of one kernel point's influence
        radius               scale used to lay out the kernel points
        fixed_kernel_points  which kernel points stay put during the layout ('none' | 'center' | 'verticals')
        KP_influence         shape of the influence ('constant' | 'linear' | 'gaussian')
        aggregation_mode     'sum' of all influences or only the 'closest' kernel point
        deformable           learn per-point kernel offsets with an inner rigid KPConv
        modulated            additionally learn a per-kernel-point modulation (deformable only)
        contraction          (extension) 'bf16x3' | 'bf16' | 'fp32'; default: env MVK_CONTRACTION or 'bf16x3'
        """
        super(KPConv, self).__init__()
        # NB like the reference, `modulated` is simply ignored by a rigid layer (blocks.py:186-191):
        # block_decider passes config.modulated to every KPConv.
        if deformable and kernel_size > 16:
            raise NotImplementedError("deformable KPConv supports at most 16 kernel points on the B200 path")
        if p_dim != 3:
            raise NotImplementedError("the B200 path handles 3D point clouds only (p_dim=3)")
        if KP_influence not in _INFLUENCE:
            raise ValueError('Unknown influence function type (config.KP_influence)')
        if aggregation_mode not in _AGGREGATION:
            raise ValueError("Unknown convolution mode. Should be 'closest' or 'sum'")

        # Save parameters
        self.K = kernel_size
        self.p_dim = p_dim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.radius = radius
        self.KP_extent = KP_extent
        self.fixed_kernel_points = fixed_kernel_points
        self.KP_influence = KP_influence
        self.aggregation_mode = aggregation_mode
        self.deformable = deformable
        self.modulated = modulated
        self.contraction = contraction or DEFAULT_CONTRACTION
        if self.contraction not in CONTRACTIONS:
            raise ValueError("contraction must be one of %r" % (CONTRACTIONS,))

        # Running variables of the deformable branch (read by the regulariser, architectures.py:21-54)
        self.min_d2 = None
        self.deformed_KP = None
        self.offset_features = None

        # Initialize weights
        self.weights = Parameter(torch.zeros((self.K, in_channels, out_channels), dtype=torch.float32),
                                 requires_grad=True)

        # Offsets come from a rigid KPConv on the same neighbourhoods (blocks.py:186-204)
        if deformable:
            self.offset_dim = (self.p_dim + 1) * self.K if modulated else self.p_dim * self.K
            self.offset_conv = KPConv(self.K, self.p_dim, self.in_channels, self.offset_dim, KP_extent, radius,
                                      fixed_kernel_points=fixed_kernel_points, KP_influence=KP_influence,
                                      aggregation_mode=aggregation_mode, contraction=self.contraction)
            self.offset_bias = Parameter(torch.zeros(self.offset_dim, dtype=torch.float32), requires_grad=True)
        else:
            self.offset_dim = None
            self.offset_conv = None
            self.offset_bias = None
        self.reset_parameters()

        # Initialize kernel points
        self.kernel_points = self.init_KP()

    def reset_parameters(self):
        kaiming_uniform_(self.weights, a=math.sqrt(5))
        if self.deformable:
            nn.init.zeros_(self.offset_bias)

    def init_KP(self):
        """Kernel point positions in a sphere (blocks.py:221-235)."""
        K_points_numpy = load_kernels(self.radius, self.K, dimension=self.p_dim, fixed=self.fixed_kernel_points)
        return Parameter(torch.tensor(K_points_numpy, dtype=torch.float32), requires_grad=False)

    def forward(self, q_pts, s_pts, neighb_inds, x):
        if self.deformable:
            # offsets (in units of KP_extent) and modulations from the offset convolution (blocks.py:243-270)
            self.offset_features = self.offset_conv(q_pts, s_pts, neighb_inds, x) + self.offset_bias
            if self.modulated:
                unscaled = self.offset_features[:, :self.p_dim * self.K].reshape(-1, self.K, self.p_dim)
                modulations = 2 * torch.sigmoid(self.offset_features[:, self.p_dim * self.K:])
            else:
                unscaled = self.offset_features.view(-1, self.K, self.p_dim)
                modulations = None
            self.deformed_KP = unscaled * self.KP_extent + self.kernel_points
            out, self.min_d2 = _KPConvDeformFunction.apply(q_pts, s_pts, neighb_inds, x, self.weights, self.deformed_KP,
                                                           modulations, self.KP_extent,
                                                           _INFLUENCE[self.KP_influence],
                                                           _AGGREGATION[self.aggregation_mode], self.contraction)
            return out
        return _KPConvFunction.apply(q_pts, s_pts, neighb_inds, x, self.weights, self.kernel_points,
                                     self.KP_extent, _INFLUENCE[self.KP_influence],
                                     _AGGREGATION[self.aggregation_mode], self.contraction)

    def __repr__(self):
        return 'KPConv(radius: {:.2f}, in_feat: {:d}, out_feat: {:d})'.format(self.radius, self.in_channels,
                                                                              self.out_channels)


# -------------------------------------------------------------------------------------------------
# gather pools (blocks.py:35-110)
# -------------------------------------------------------------------------------------------------
class _PoolFunction(torch.autograd.Function):

    @staticmethod
    def forward(ctx, x, inds, mode):
        _lib.require_cuda()
        L = _lib.lib()
        if not x.is_cuda:
            raise RuntimeError("pool: tensors must live on a CUDA device (no CPU fallback)")
        xf = x.detach().contiguous().float()
        ii, is64 = _idx(inds if inds.dim() == 2 else inds.reshape(-1, 1))
        ns, c = xf.shape
        nq, h = ii.shape
        out = torch.empty((nq, c), dtype=torch.float32, device=xf.device)
        arg = torch.empty((nq, c), dtype=torch.int32, device=xf.device) if mode == 0 else None
        with _lib.on_device(xf.device):
            check(L.mvk_pool(ptr(xf), ns, c, ptr(ii), is64, nq, h, mode, ptr(out), ptr(arg), stream_ptr()))
        ctx.save_for_backward(ii, arg if arg is not None else ii)
        ctx.cfg = (ns, c, nq, h, mode, is64)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        L = _lib.lib()
        ii, arg = ctx.saved_tensors
        ns, c, nq, h, mode, is64 = ctx.cfg
        go = grad_out
        if not (go.dtype is torch.float32 and go.dim() == 2 and go.stride(1) == 1 and go.stride(0) >= c):
            go = _f32c(go)  # otherwise only the row pitch differs (a half of a torch.cat backward): no copy
        gx = torch.zeros((ns, c), dtype=torch.float32, device=go.device)
        with _lib.on_device(go.device):
            check(L.mvk_pool_bwd(go.data_ptr(), go.stride(0), nq, c, ptr(arg) if mode == 0 else None, ptr(ii), is64, h,
                                 mode, ns, ptr(gx), stream_ptr()))
        return gx, None, None


def max_pool(x, inds):
    """Pools features with the maximum values (blocks.py:93-110).  NB: like the reference the
    shadow row is ZERO (not -inf), so a row with shadow neighbours never pools below 0.
    :param x: [n1, d] features matrix
    :param inds: [n2, max_num] pooling indices
    :return: [n2, d] pooled features matrix
    """
    return _PoolFunction.apply(x, inds, 0)


def closest_pool(x, inds):
    """Pools features from the closest neighbors (blocks.py:79-90); only the first column is used.
    :param x: [n1, d] features matrix
    :param inds: [n2, max_num]
    :return: [n2, d] pooled features matrix
    """
    return _PoolFunction.apply(x, inds, 1)


def gather(x, idx, method=2):
    """x[idx] with a shadow-free index tensor (blocks.py:35-66).  All three reference `method`s
    compute the same values; they only differ in how autograd scatters."""
    if method not in (0, 1, 2):
        raise ValueError('Unkown method')
    return x[idx]
