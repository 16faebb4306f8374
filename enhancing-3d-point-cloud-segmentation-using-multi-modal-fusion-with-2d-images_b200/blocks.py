"""Point-wise blocks either side of KPConv, backed by libmvk.

Mirrors of the reference modules (KPConv-PyTorch/models/blocks.py):

    BatchNormBlock(in_dim, use_bn, bn_momentum)                  :430-466
    UnaryBlock(in_dim, out_dim, use_bn, bn_momentum, no_relu)    :469-504

with the same attribute / parameter names (``mlp.weight``, ``batch_norm.batch_norm.{weight,bias,
running_mean,running_var,num_batches_tracked}`` or ``batch_norm.bias``) so reference checkpoints
load.  The Linear runs on the tcgen05 contraction (bf16 hi/lo operands, fp32 accumulation), batch
statistics / normalisation / LeakyReLU / residual add are fused streaming kernels; the backward
writes the pre-activation gradient directly in the contraction's operand format.

    unary_forward(x, block, residual=None)   z = leaky( bn( x W^T ) [+ residual] )
    bn_act(y, bn_block, slope, residual)     z = leaky( bn(y) [+ residual] )

No CPU fallback: tensors must be CUDA tensors.
"""
import os

import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import _lib, _weights
from ._lib import check, ptr, stream_ptr
from . import kpconv as _kp


_FUSE_BN_STATS = os.environ.get("MVK_FUSE_BN_STATS", "0") == "1"
# decoder step: contract the upsampled half at the coarse level (_UpAddLinearBNAct); "0" = concatenated operand
_DECODER_SPLIT = os.environ.get("MVK_DECODER_SPLIT", "1") != "0"


def _r8(v):
    return (v + 7) // 8 * 8


def _carve(device, *sizes):
    """ONE device allocation carved into 256-byte aligned raw sub-buffers (pointer arithmetic instead
    of tensor views: creating a tensor object costs as much host time as a small kernel launch).
    Returns (keep-alive tensor, [device pointer or None for empty sizes, ...])."""
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (int(n) + 255) & ~255
    t = torch.empty(max(total, 256), dtype=torch.uint8, device=device)
    base = t.data_ptr()
    return t, [(base + o) if n else None for o, n in zip(offs, sizes)]


class _HiLo:
    """bf16 hi/lo operand pair of an activation tensor, living inside some keep-alive allocation."""
    __slots__ = ("keep", "hi", "lo", "rows", "ld", "version")

    def __init__(self, keep, hi, lo, rows, ld, version):
        self.keep, self.hi, self.lo, self.rows, self.ld, self.version = keep, hi, lo, rows, ld, version


def _attach_hilo(t, keep, hi, lo, rows, ld):
    """Remember the bf16 hi/lo split of tensor `t` on the tensor object: the next Linear that consumes
    `t` (UnaryBlock) picks it up instead of re-reading and re-splitting `t`."""
    try:
        t._mvk_hilo = _HiLo(keep, hi, lo, rows, ld, t._version)
    except Exception:
        pass


def _cached_hilo(t, rows, ld):
    c = getattr(t, "_mvk_hilo", None)
    if c is None or c.version != t._version or c.rows != rows or c.ld != ld or c.keep.device != t.device:
        return None
    return c


def _bn_args(bn_block):
    """(use_bn, gamma, beta_or_bias, running_mean, running_var, momentum, eps, training, module)"""
    if bn_block is None:
        return (False, None, None, None, None, 0.0, 0.0, False, None)
    if bn_block.use_bn:
        m = bn_block.batch_norm
        return (True, m.weight, m.bias, m.running_mean, m.running_var, float(m.momentum), float(m.eps),
                bool(m.training), m)
    return (False, None, bn_block.bias, None, None, 0.0, 0.0, False, None)


def _f32c(t):
    """fp32 contiguous version of `t` (no new tensor object in the common case)."""
    if t.dtype is not torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


def _rows_f32(t):
    """(tensor, row pitch) of a 2-D fp32 gradient without copying when only the row pitch is non-trivial
    (the halves of a torch.cat backward are such views)."""
    if t.dtype is torch.float32 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1] and t.stride(0) % 4 == 0 \
            and t.data_ptr() % 16 == 0:
        return t, t.stride(0)
    t = _f32c(t)
    return t, t.shape[1]


def _norm_forward(L, y_ptr, ld, rows, cols, use_bn, training, gamma, beta, rm, rv, momentum, eps, st, nbt, vec, dev,
                  stats=None):
    """Batch-norm bookkeeping of one forward.  `vec` = device pointers [scale, shift, mean, invstd] (each
    `cols` floats).  Returns the (scale, shift, mean, invstd) pointers the activation kernels use."""
    if not use_bn:
        return None, (beta.data_ptr() if beta is not None else None), None, None
    sc, sh, mu, isd = vec
    if training and rows == 1:  # like nn.BatchNorm1d in the reference blocks (blocks.py:446-460)
        raise ValueError("Expected more than 1 value per channel when training, got input size [1, %d]" % cols)
    if training and rows > 0 and stats is not None:
        # the column sums came out of the contraction's epilogue: only the [cols]-sized bookkeeping is left
        check(L.mvk_bn_finalize(stats, rows, cols, gamma.data_ptr(), beta.data_ptr(), eps, momentum, 1, ptr(rm), ptr(rv),
                                sc, sh, mu, isd, ptr(nbt), st))
    elif training and rows > 0:
        # column sums and, in the CTA that retires last, scale / shift / running statistics: one launch
        stats = _lib.zeros_ptr(8 * (2 * cols + 1), dev)
        check(L.mvk_bn_batch_stats(y_ptr, rows, cols, ld, stats, gamma.data_ptr(), beta.data_ptr(), eps, momentum,
                                   ptr(rm), ptr(rv), sc, sh, mu, isd, ptr(nbt), st))
    else:
        check(L.mvk_bn_finalize(None, rows, cols, gamma.data_ptr(), beta.data_ptr(), eps, momentum, 0, ptr(rm), ptr(rv),
                                sc, sh, mu, isd, None, st))
    return sc, sh, mu, isd


_GRAD_HILO = os.environ.get("MVK_GRAD_HILO", "1") != "0"  # development switch (A/B timing)


class _BNAct(torch.autograd.Function):
    """z = leaky(bn(y) [+ residual]); blocks.py:446-460 + the activation / residual that follows."""

    @staticmethod
    def forward(ctx, y, gamma, beta, residual, rm, rv, use_bn, training, momentum, eps, slope, nbt=None,
                emit_hilo=False, grad_hilo=False):
        _lib.require_cuda()
        L = _lib.lib()
        if not y.is_cuda:
            raise RuntimeError("bn_act: tensors must live on a CUDA device (no CPU fallback)")
        yf = _f32c(y)
        rows, cols = yf.shape
        dev = yf.device
        res = None if residual is None else _f32c(residual)
        emit = bool(emit_hilo) and cols % 8 == 0
        st = stream_ptr()
        with _lib.on_device(dev):
            keep, (sc, sh, mu, isd, hi, lo) = _carve(dev, *([4 * cols] * 4 if use_bn else [0] * 4),
                                                      *([2 * rows * cols] * 2 if emit else [0, 0]))
            sc, sh, mu, isd = _norm_forward(L, yf.data_ptr(), cols, rows, cols, use_bn, training, gamma, beta, rm, rv,
                                            momentum, eps, st, nbt, (sc, sh, mu, isd), dev)
            z = torch.empty_like(yf)
            check(L.mvk_scale_shift_act(yf.data_ptr(), rows, cols, cols, sc, sh, ptr(res), cols, slope, z.data_ptr(),
                                        cols, hi, lo, cols, st))
        ctx.save_for_backward(yf, keep, res, beta if not use_bn else None)
        ctx.cfg = (rows, cols, use_bn, training, slope, gamma is not None, beta is not None, (sc, sh, mu, isd))
        ctx.grad_hilo = bool(grad_hilo) and cols % 8 == 0 and _GRAD_HILO
        if emit:
            _attach_hilo(z, keep, hi, lo, rows, cols)
        return z

    @staticmethod
    def backward(ctx, dz):
        L = _lib.lib()
        yf, keep, res, _bias = ctx.saved_tensors
        rows, cols, use_bn, training, slope, has_g, has_b, (sc, sh, mu, isd) = ctx.cfg
        g, ldg = _rows_f32(dz)
        dev = g.device
        st = stream_ptr()
        need_y, need_res = ctx.needs_input_grad[0], ctx.needs_input_grad[3]
        batch_stats = 1 if (use_bn and training) else 0
        with _lib.on_device(dev):
            sums = _lib.zeros_ptr(16 * cols, dev)
            if has_g or has_b or batch_stats:
                check(L.mvk_act_bwd_reduce(g.data_ptr(), ldg, yf.data_ptr(), rows, cols, cols, sc, sh, ptr(res), cols,
                                           mu, isd, slope, sums, st))
            dy = torch.empty_like(yf) if need_y else None
            dres = torch.empty_like(yf) if (need_res and res is not None) else None
            dgamma = torch.empty(cols, dtype=torch.float32, device=dev) if has_g else None
            dbeta = torch.empty(cols, dtype=torch.float32, device=dev) if has_b else None
            # y came out of a KPConv: its backward contracts dy on the tensor cores, so the pair is written here
            gkeep, (g_hi, g_lo) = _carve(dev, *([2 * rows * cols] * 2 if (ctx.grad_hilo and need_y) else [0, 0]))
            check(L.mvk_act_bwd_apply(g.data_ptr(), ldg, yf.data_ptr(), rows, cols, cols, sc, sh, ptr(res), cols, mu, isd,
                                      slope, sums, batch_stats, ptr(dy), cols, g_hi, g_lo, cols, ptr(dres), cols,
                                      ptr(dgamma), ptr(dbeta), st))
            if g_hi is not None:
                _attach_hilo(dy, gkeep, g_hi, g_lo, rows, cols)
        return dy, dgamma, dbeta, dres, None, None, None, None, None, None, None, None, None, None


class _LinearBNAct(torch.autograd.Function):
    """z = leaky(bn(x W^T) [+ residual]); UnaryBlock.forward, blocks.py:493-498."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, residual, rm, rv, use_bn, training, momentum, eps, slope, contraction,
                nbt=None, emit_hilo=False):
        _lib.require_cuda()
        L = _lib.lib()
        if not x.is_cuda:
            raise RuntimeError("UnaryBlock: tensors must live on a CUDA device (no CPU fallback)")
        xf = _f32c(x)
        w = _f32c(weight)
        rows, cin = xf.shape
        cout = w.shape[0]
        dev = xf.device
        res = None if residual is None else _f32c(residual)
        emit = bool(emit_hilo) and cout % 8 == 0
        st = stream_ptr()
        fp32 = contraction == "fp32"
        ldx = cin if fp32 else _r8(cin)
        cached = None if fp32 else _cached_hilo(x, rows, ldx)
        with _lib.on_device(dev):
            nx = 0 if (fp32 or cached is not None) else 2 * rows * ldx
            keep, (x_hi, x_lo, y, sc, sh, mu, isd, z_hi, z_lo) = _carve(
                dev, nx, nx, 4 * rows * cout, *([4 * cout] * 4 if use_bn else [0] * 4),
                *([2 * rows * cout] * 2 if emit else [0, 0]))
            xkeep = stats = wkeep = w_hi = w_lo = None
            if fp32:
                if rows > 0:
                    check(L.mvk_gemm_f32(xf.data_ptr(), cin, 1, w.data_ptr(), 1, cin, rows, cout, cin, y, cout, 1, st))
            else:
                terms = 3 if contraction == "bf16x3" else 1
                if cached is not None:
                    x_hi, x_lo, xkeep = cached.hi, cached.lo, cached.keep  # written by the producer's epilogue
                else:
                    check(L.mvk_split_bf16(xf.data_ptr(), rows, cin, cin, x_hi, x_lo, rows, ldx, st))
                    _attach_hilo(x, keep, x_hi, x_lo, rows, ldx)
                w_hi, w_lo, wkeep, _ = _weights.weight_operands(w, cout, cin, cin, cout, ldx)  # once per optimiser step
                if rows > 0 and use_bn and training and _FUSE_BN_STATS:
                    # batch statistics ride in the contraction's epilogue (measured: the extra epilogue work
                    # costs as much as the separate statistics pass it saves, so this is off by default)
                    stats = _lib.zeros_ptr(16 * cout, dev)
                    check(L.mvk_gemm_bf16x3_stats(x_hi, x_lo, 0, ldx, w_hi, w_lo, 0, ldx, rows, cout, cin, y, cout, cout,
                                                  terms, 0, stats, st))
                elif rows > 0:
                    check(L.mvk_gemm_bf16x3(x_hi, x_lo, 0, ldx, w_hi, w_lo, 0, ldx, rows, cout, cin, y, cout, cout,
                                            terms, 0, st))
            sc, sh, mu, isd = _norm_forward(L, y, cout, rows, cout, use_bn, training, gamma, beta, rm, rv, momentum, eps,
                                            st, nbt, (sc, sh, mu, isd), dev, stats)
            z = torch.empty((rows, cout), dtype=torch.float32, device=dev)
            check(L.mvk_scale_shift_act(y, rows, cout, cout, sc, sh, ptr(res), cout, slope, z.data_ptr(), cout, z_hi, z_lo,
                                        cout, st))
        # w is saved on every path: autograd's version check on it also guards its bf16 pair (wkeep)
        ctx.save_for_backward(keep, xkeep, res, xf if fp32 else None, w, beta if not use_bn else None)
        ctx.cfg = (rows, cin, cout, use_bn, training, slope, gamma is not None, beta is not None, contraction, ldx,
                   (x_hi, x_lo, w_hi, w_lo, y, sc, sh, mu, isd), wkeep)
        if emit:
            _attach_hilo(z, keep, z_hi, z_lo, rows, cout)
        return z

    @staticmethod
    def backward(ctx, dz):
        dx, dw, dgamma, dbeta, dres = _LinearBNAct._backward(ctx, dz, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                                                            ctx.needs_input_grad[4])
        return dx, dw, dgamma, dbeta, dres, None, None, None, None, None, None, None, None, None, None

    @staticmethod
    def _backward(ctx, dz, need_x, need_w, need_res):
        L = _lib.lib()
        keep, xkeep, res, xf, w, _bias = ctx.saved_tensors
        rows, cin, cout, use_bn, training, slope, has_g, has_b, contraction, ldx, ptrs, _wkeep = ctx.cfg
        x_hi, x_lo, w_hi, w_lo, y, sc, sh, mu, isd = ptrs
        g, ldg = _rows_f32(dz)
        dev = g.device
        st = stream_ptr()
        batch_stats = 1 if (use_bn and training) else 0
        dx = dw = None
        with _lib.on_device(dev):
            sums = _lib.zeros_ptr(16 * cout, dev)
            if has_g or has_b or batch_stats:
                check(L.mvk_act_bwd_reduce(g.data_ptr(), ldg, y, rows, cout, cout, sc, sh, ptr(res), cout, mu, isd, slope,
                                           sums, st))
            dres = torch.empty((rows, cout), dtype=torch.float32, device=dev) if (need_res and res is not None) else None
            dgamma = torch.empty(cout, dtype=torch.float32, device=dev) if has_g else None
            dbeta = torch.empty(cout, dtype=torch.float32, device=dev) if has_b else None
            if contraction == "fp32":
                dy = torch.empty((rows, cout), dtype=torch.float32, device=dev)
                check(L.mvk_act_bwd_apply(g.data_ptr(), ldg, y, rows, cout, cout, sc, sh, ptr(res), cout, mu, isd, slope,
                                          sums, batch_stats, dy.data_ptr(), cout, None, None, 0, ptr(dres), cout,
                                          ptr(dgamma), ptr(dbeta), st))
                if need_x:
                    dx = torch.empty((rows, cin), dtype=torch.float32, device=dev)
                    if rows > 0:
                        check(L.mvk_gemm_f32(dy.data_ptr(), cout, 1, w.data_ptr(), cin, 1, rows, cin, cout, dx.data_ptr(),
                                             cin, 1, st))
                if need_w:
                    dw = torch.zeros((cout, cin), dtype=torch.float32, device=dev)
                    if rows > 0:
                        split = max(1, min(rows // 64, 2 * 148 // max(1, ((cout + 63) // 64) * ((cin + 63) // 64))))
                        check(L.mvk_gemm_f32(dy.data_ptr(), 1, cout, xf.data_ptr(), cin, 1, cout, cin, rows, dw.data_ptr(),
                                             cin, split, st))
            else:
                terms = 3 if contraction == "bf16x3" else 1
                ldh = _r8(cout)
                dkeep, (dy_hi, dy_lo) = _carve(dev, 2 * rows * ldh, 2 * rows * ldh)
                check(L.mvk_act_bwd_apply(g.data_ptr(), ldg, y, rows, cout, cout, sc, sh, ptr(res), cout, mu, isd, slope,
                                          sums, batch_stats, None, 0, dy_hi, dy_lo, ldh, ptr(dres), cout, ptr(dgamma),
                                          ptr(dbeta), st))
                if need_x:
                    dx = torch.empty((rows, cin), dtype=torch.float32, device=dev)
                    if rows > 0:
                        # dx = dy W : A = dy [rows, cout] K-major, B = W stored [K = cout, N = cin] (N contiguous)
                        check(L.mvk_gemm_bf16x3(dy_hi, dy_lo, 0, ldh, w_hi, w_lo, 1, ldx, rows, cin, cout, dx.data_ptr(),
                                                cin, cin, terms, 0, st))
                if need_w:
                    if rows > 0:
                        # dW = dy^T x : both operands stored [K = rows, *] with the M / N index contiguous;
                        # split_k = 0: the library splits the reduction over the SMs and zeroes dW itself
                        dw = torch.empty((cout, cin), dtype=torch.float32, device=dev)
                        check(L.mvk_gemm_bf16x3(dy_hi, dy_lo, 1, ldh, x_hi, x_lo, 1, ldx, cout, cin, rows, dw.data_ptr(),
                                                cin, cin, terms, 0, st))
                    else:
                        dw = torch.zeros((cout, cin), dtype=torch.float32, device=dev)
        return dx, dw, dgamma, dbeta, dres


class _PreSplit:
    """Stands in for an activation tensor whose fp32 form is never materialised: only its bf16 hi/lo
    operand pair exists (written by a fused producer).  _LinearBNAct.forward only asks an input for its
    shape / device / cached pair on the tensor-core paths."""
    dtype = torch.float32
    is_cuda = True
    _version = 0

    def __init__(self, rows, cols, device, keep, hi, lo, ld):
        self.shape, self.device = (rows, cols), device
        self._mvk_hilo = _HiLo(keep, hi, lo, rows, ld, 0)

    def is_contiguous(self):
        return True


class _UpCatLinearBNAct(torch.autograd.Function):
    """z = leaky(bn(cat([closest_pool(x_coarse, inds), skip], 1) W^T)): the decoder step of KPFCNN
    (architectures.py:300-306 + blocks.py:493-498) with the concatenation fused into the operand split.
    Argument positions 0 / 1 / 4 mirror _LinearBNAct (x, weight, residual)."""

    @staticmethod
    def forward(ctx, x_coarse, weight, gamma, beta, skip, rm, rv, use_bn, training, momentum, eps, slope, contraction,
                nbt, emit_hilo, inds):
        _lib.require_cuda()
        L = _lib.lib()
        if contraction == "fp32":
            raise RuntimeError("the fused decoder step needs a tensor-core contraction ('bf16x3' or 'bf16')")
        xc, sk = _f32c(x_coarse), _f32c(skip)
        ii, is64 = _kp._idx(inds if inds.dim() == 2 else inds.reshape(-1, 1))
        ns, c1 = xc.shape
        nq, c2 = sk.shape
        if ii.shape[0] != nq or c1 % 4 or c2 % 4 or (c1 + c2) % 8:
            raise RuntimeError("upsample+concat: channel counts must be multiples of 4 (sum: of 8)")
        dev = xc.device
        ldh = c1 + c2
        with _lib.on_device(dev):
            ukeep, (u_hi, u_lo) = _carve(dev, 2 * nq * ldh, 2 * nq * ldh)
            check(L.mvk_upsample_concat_split(xc.data_ptr(), ns, c1, ii.data_ptr(), is64, nq, ii.shape[1], sk.data_ptr(),
                                              c2, c2, u_hi, u_lo, ldh, stream_ptr()))
        virt = _PreSplit(nq, ldh, dev, ukeep, u_hi, u_lo, ldh)
        z = _LinearBNAct.forward(ctx, virt, weight, gamma, beta, None, rm, rv, use_bn, training, momentum, eps, slope,
                                 contraction, nbt, emit_hilo)
        ctx.up = (ii, is64, ns, c1, c2)
        return z

    @staticmethod
    def backward(ctx, dz):
        L = _lib.lib()
        need_c, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[4]
        dx, dw, dgamma, dbeta, _ = _LinearBNAct._backward(ctx, dz, need_c or need_s, ctx.needs_input_grad[1], False)
        ii, is64, ns, c1, c2 = ctx.up
        dcoarse = dskip = None
        if dx is not None:
            nq = dx.shape[0]
            if need_s:
                dskip = dx[:, c1:]  # a view: consumers take the row pitch (or copy)
            if need_c:
                dcoarse = torch.zeros((ns, c1), dtype=torch.float32, device=dx.device)
                with _lib.on_device(dx.device):
                    check(L.mvk_pool_bwd(dx.data_ptr(), c1 + c2, nq, c1, None, ii.data_ptr(), is64, ii.shape[1], 1, ns,
                                         dcoarse.data_ptr(), stream_ptr()))
        return dcoarse, dw, dgamma, dbeta, dskip, None, None, None, None, None, None, None, None, None, None, None


class _UpAddLinearBNAct(torch.autograd.Function):
    """The decoder step of KPFCNN (architectures.py:300-306 + blocks.py:493-498),
        z = leaky(bn(cat([closest_pool(x_coarse, inds), skip], 1) W^T)),
    with the upsampled half contracted at the COARSE level: a Linear commutes with a row gather, so
        cat([up(x), skip]) W^T = up(x W_up^T) + skip W_skip^T .
    The fine-level contraction then runs over the skip channels only (at level 0 of the baseline net: 128 of 384
    columns, 137 MB instead of 410 MB of operand per pass), the coarse one over 4.6x fewer rows, and
    mvk_gather_add_rows adds the gathered coarse result.  Backward mirrors it: dskip = dy W_skip at the fine level,
    d(x W_up^T) = closest_pool^T(dy) scattered to the coarse rows, dx_coarse / dW_up contracted there.
    Argument positions 0 / 1 / 4 mirror _LinearBNAct (x, weight, residual)."""

    @staticmethod
    def forward(ctx, x_coarse, weight, gamma, beta, skip, rm, rv, use_bn, training, momentum, eps, slope, contraction,
                nbt, emit_hilo, inds):
        _lib.require_cuda()
        L = _lib.lib()
        if contraction == "fp32":
            raise RuntimeError("the split decoder step needs a tensor-core contraction ('bf16x3' or 'bf16')")
        xc, sk, w = _f32c(x_coarse), _f32c(skip), _f32c(weight)
        ii, is64 = _kp._idx(inds if inds.dim() == 2 else inds.reshape(-1, 1))
        ns, c1 = xc.shape
        nq, c2 = sk.shape
        cout, cin = w.shape
        if ii.shape[0] != nq or cin != c1 + c2 or c1 % 8 or c2 % 8 or cout % 4:
            raise RuntimeError("split decoder step: channel counts must be multiples of 8 (outputs: of 4)")
        dev = xc.device
        st = stream_ptr()
        terms = 3 if contraction == "bf16x3" else 1
        emit = bool(emit_hilo) and cout % 8 == 0
        cc, cs = _cached_hilo(x_coarse, ns, c1), _cached_hilo(skip, nq, c2)
        with _lib.on_device(dev):
            keep, (xc_hi, xc_lo, sk_hi, sk_lo, zc, y, sc, sh, mu, isd, z_hi, z_lo) = _carve(
                dev, *([0, 0] if cc is not None else [2 * ns * c1] * 2), *([0, 0] if cs is not None else [2 * nq * c2] * 2),
                4 * ns * cout, 4 * nq * cout, *([4 * cout] * 4 if use_bn else [0] * 4),
                *([2 * nq * cout] * 2 if emit else [0, 0]))
            ckeep = skeep = None
            if cc is not None:
                xc_hi, xc_lo, ckeep = cc.hi, cc.lo, cc.keep
            elif ns > 0:
                check(L.mvk_split_bf16(xc.data_ptr(), ns, c1, c1, xc_hi, xc_lo, ns, c1, st))
            if cs is not None:
                sk_hi, sk_lo, skeep = cs.hi, cs.lo, cs.keep
            elif nq > 0:
                check(L.mvk_split_bf16(sk.data_ptr(), nq, c2, c2, sk_hi, sk_lo, nq, c2, st))
            w_hi, w_lo, wkeep, _ = _weights.weight_operands(w, cout, cin, cin, cout, cin)  # [cout, c1 + c2], K-major
            if nq > 0:
                # fine level: skip W_skip^T  (B = columns c1.. of W: same row pitch, pointer moved by c1 elements)
                check(L.mvk_gemm_bf16x3(sk_hi, sk_lo, 0, c2, w_hi + 2 * c1, w_lo + 2 * c1, 0, cin, nq, cout, c2, y, cout, cout,
                                        terms, 0, st))
                if ns > 0:
                    # coarse level: x W_up^T, then gathered onto the fine rows
                    check(L.mvk_gemm_bf16x3(xc_hi, xc_lo, 0, c1, w_hi, w_lo, 0, cin, ns, cout, c1, zc, cout, cout, terms, 0, st))
                    check(L.mvk_gather_add_rows(y, cout, nq, cout, zc, ns, ii.data_ptr(), is64, ii.shape[1], st))
            sc, sh, mu, isd = _norm_forward(L, y, cout, nq, cout, use_bn, training, gamma, beta, rm, rv, momentum, eps, st,
                                            nbt, (sc, sh, mu, isd), dev, None)
            z = torch.empty((nq, cout), dtype=torch.float32, device=dev)
            check(L.mvk_scale_shift_act(y, nq, cout, cout, sc, sh, None, cout, slope, z.data_ptr(), cout, z_hi, z_lo, cout, st))
        ctx.save_for_backward(keep, ckeep, skeep, w, ii, beta if not use_bn else None)
        ctx.cfg = (ns, nq, c1, c2, cout, use_bn, training, slope, gamma is not None, beta is not None, terms, is64,
                   (xc_hi, xc_lo, sk_hi, sk_lo, w_hi, w_lo, y, sc, sh, mu, isd), wkeep)
        if emit:
            _attach_hilo(z, keep, z_hi, z_lo, nq, cout)
        return z

    @staticmethod
    def backward(ctx, dz):
        L = _lib.lib()
        keep, ckeep, skeep, w, ii, _bias = ctx.saved_tensors
        ns, nq, c1, c2, cout, use_bn, training, slope, has_g, has_b, terms, is64, ptrs, _wkeep = ctx.cfg
        xc_hi, xc_lo, sk_hi, sk_lo, w_hi, w_lo, y, sc, sh, mu, isd = ptrs
        cin = c1 + c2
        need_c, need_w, need_s = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[4]
        g, ldg = _rows_f32(dz)
        dev = g.device
        st = stream_ptr()
        batch_stats = 1 if (use_bn and training) else 0
        dcoarse = dw = dskip = None
        with _lib.on_device(dev):
            sums = _lib.zeros_ptr(16 * cout, dev)
            if has_g or has_b or batch_stats:
                check(L.mvk_act_bwd_reduce(g.data_ptr(), ldg, y, nq, cout, cout, sc, sh, None, cout, mu, isd, slope, sums, st))
            dgamma = torch.empty(cout, dtype=torch.float32, device=dev) if has_g else None
            dbeta = torch.empty(cout, dtype=torch.float32, device=dev) if has_b else None
            coarse = (need_c or need_w) and ns > 0
            ldh = _r8(cout)
            dkeep, (dy_hi, dy_lo, dy32, dzc, dzc_hi, dzc_lo) = _carve(
                dev, 2 * nq * ldh, 2 * nq * ldh, 4 * nq * cout if coarse else 0, 4 * ns * cout if coarse else 0,
                2 * ns * ldh if coarse else 0, 2 * ns * ldh if coarse else 0)
            check(L.mvk_act_bwd_apply(g.data_ptr(), ldg, y, nq, cout, cout, sc, sh, None, cout, mu, isd, slope, sums,
                                      batch_stats, dy32, cout, dy_hi, dy_lo, ldh, None, cout, ptr(dgamma), ptr(dbeta), st))
            if need_s:
                dskip = torch.empty((nq, c2), dtype=torch.float32, device=dev)
                if nq > 0:  # dskip = dy W_skip : B = W stored [K = cout, N = cin], columns c1.. (N contiguous)
                    check(L.mvk_gemm_bf16x3(dy_hi, dy_lo, 0, ldh, w_hi + 2 * c1, w_lo + 2 * c1, 1, cin, nq, c2, cout,
                                            dskip.data_ptr(), c2, c2, terms, 0, st))
            if coarse:
                # gradient of the gathered coarse result: every fine row adds its dy row to its coarse parent
                _zero_region(dkeep, dzc, 4 * ns * cout)
                if nq > 0:
                    check(L.mvk_pool_bwd(dy32, cout, nq, cout, None, ii.data_ptr(), is64, ii.shape[1], 1, ns, dzc, st))
                check(L.mvk_split_bf16(dzc, ns, cout, cout, dzc_hi, dzc_lo, ns, ldh, st))
                if need_c:
                    dcoarse = torch.empty((ns, c1), dtype=torch.float32, device=dev)
                    check(L.mvk_gemm_bf16x3(dzc_hi, dzc_lo, 0, ldh, w_hi, w_lo, 1, cin, ns, c1, cout, dcoarse.data_ptr(), c1, c1,
                                            terms, 0, st))
            elif need_c:
                dcoarse = torch.zeros((ns, c1), dtype=torch.float32, device=dev)
            if need_w:
                dw = torch.empty((cout, cin), dtype=torch.float32, device=dev)
                if coarse:  # dW_up = (d(x W_up^T))^T x_coarse, contracted over the coarse rows
                    check(L.mvk_gemm_bf16x3(dzc_hi, dzc_lo, 1, ldh, xc_hi, xc_lo, 1, c1, cout, c1, ns, dw.data_ptr(), cin, c1,
                                            terms, 0, st))
                else:
                    dw[:, :c1].zero_()
                if nq > 0:  # dW_skip = dy^T skip
                    check(L.mvk_gemm_bf16x3(dy_hi, dy_lo, 1, ldh, sk_hi, sk_lo, 1, c2, cout, c2, nq, dw.data_ptr() + 4 * c1, cin,
                                            c2, terms, 0, st))
                else:
                    dw[:, c1:].zero_()
        return dcoarse, dw, dgamma, dbeta, dskip, None, None, None, None, None, None, None, None, None, None, None


def _zero_region(keep, ptr_, nbytes):
    """Zero `nbytes` of the carved allocation `keep` starting at device pointer `ptr_` (on the current stream)."""
    if nbytes:
        off = ptr_ - keep.data_ptr()
        keep[off:off + nbytes].zero_()


# -------------------------------------------------------------------------------------------------
class BatchNormBlock(nn.Module):

    def __init__(self, in_dim, use_bn, bn_momentum):
        """Batch norm over the stacked points (use_bn) or a learned per-channel bias (not use_bn);
        same constructor and parameter names as the reference block (blocks.py:432-447)."""
        super(BatchNormBlock, self).__init__()
        self.bn_momentum = bn_momentum
        self.use_bn = use_bn
        self.in_dim = in_dim
        if self.use_bn:
            self.batch_norm = nn.BatchNorm1d(in_dim, momentum=bn_momentum)
        else:
            self.bias = Parameter(torch.zeros(in_dim, dtype=torch.float32), requires_grad=True)

    def reset_parameters(self):
        nn.init.zeros_(self.bias)

    def forward(self, x):
        return bn_act(x, self, slope=1.0)

    def __repr__(self):
        return 'BatchNormBlock(in_feat: {:d}, momentum: {:.3f}, only_bias: {:s})'.format(self.in_dim,
                                                                                         self.bn_momentum,
                                                                                         str(not self.use_bn))


def _nbt(bn_module, rows):
    """num_batches_tracked buffer to be incremented inside the statistics kernel (training only)."""
    if bn_module is not None and bn_module.training and bn_module.num_batches_tracked is not None and rows > 0:
        return bn_module.num_batches_tracked
    return None


def bn_act(y, bn_block, slope=0.1, residual=None, emit_hilo=False, grad_hilo=False):
    """leaky_relu(bn_block(y) [+ residual], slope); slope = 1 disables the activation.
    emit_hilo: also write the result as a bf16 hi/lo pair for a following UnaryBlock (same kernel).
    grad_hilo: y is a KPConv output -- the backward also writes the gradient w.r.t. y as a bf16 pair."""
    use_bn, gamma, beta, rm, rv, momentum, eps, training, mod = _bn_args(bn_block)
    return _BNAct.apply(y, gamma, beta, residual, rm, rv, use_bn, training, momentum, eps, float(slope),
                        _nbt(mod, y.shape[0]), emit_hilo, grad_hilo)


class UnaryBlock(nn.Module):

    def __init__(self, in_dim, out_dim, use_bn, bn_momentum, no_relu=False, contraction=None):
        """Linear (no bias) -> BatchNormBlock -> LeakyReLU(0.1) unless no_relu; same constructor and
        parameter names as the reference block (blocks.py:471-491).  `contraction` (extension) picks
        the Linear's arithmetic: 'bf16x3' (default) | 'bf16' | 'fp32'."""
        super(UnaryBlock, self).__init__()
        self.bn_momentum = bn_momentum
        self.use_bn = use_bn
        self.no_relu = no_relu
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.mlp = nn.Linear(in_dim, out_dim, bias=False)
        self.batch_norm = BatchNormBlock(out_dim, self.use_bn, self.bn_momentum)
        if not no_relu:
            self.leaky_relu = nn.LeakyReLU(0.1)
        self.contraction = contraction or _kp.DEFAULT_CONTRACTION

    def forward(self, x, batch=None, residual=None, slope=None, emit_hilo=False):
        """x -> leaky_relu(batch_norm(mlp(x))) (blocks.py:493-498).  Extension used by the block
        tail: `residual` is added before the activation and `slope` overrides the block's own."""
        if slope is None:
            slope = 1.0 if self.no_relu else 0.1
        use_bn, gamma, beta, rm, rv, momentum, eps, training, mod = _bn_args(self.batch_norm)
        return _LinearBNAct.apply(x, self.mlp.weight, gamma, beta, residual, rm, rv, use_bn, training, momentum, eps,
                                  float(slope), self.contraction, _nbt(mod, x.shape[0]), emit_hilo)

    def forward_upsampled(self, x_coarse, up_inds, skip, emit_hilo=False):
        """``self(torch.cat([closest_pool(x_coarse, up_inds), skip], dim=1))`` -- the decoder step of KPFCNN
        (architectures.py:300-306) -- without materialising the upsampled or the concatenated tensor."""
        if self.contraction == "fp32" or x_coarse.shape[1] % 4 or skip.shape[1] % 4 or self.in_dim % 8:
            return self.forward(torch.cat([_kp.closest_pool(x_coarse, up_inds), skip], dim=1), emit_hilo=emit_hilo)
        slope = 1.0 if self.no_relu else 0.1
        use_bn, gamma, beta, rm, rv, momentum, eps, training, mod = _bn_args(self.batch_norm)
        split = _DECODER_SPLIT and x_coarse.shape[1] % 8 == 0 and skip.shape[1] % 8 == 0 and self.out_dim % 4 == 0
        fn = _UpAddLinearBNAct if split else _UpCatLinearBNAct
        return fn.apply(x_coarse, self.mlp.weight, gamma, beta, skip, rm, rv, use_bn, training, momentum,
                        eps, float(slope), self.contraction, _nbt(mod, skip.shape[0]), emit_hilo, up_inds)

    def __repr__(self):
        return 'UnaryBlock(in_feat: {:d}, out_feat: {:d}, BN: {:s}, ReLU: {:s})'.format(self.in_dim, self.out_dim,
                                                                                        str(self.use_bn),
                                                                                        str(not self.no_relu))


# -------------------------------------------------------------------------------------------------
class _SoftmaxXent(torch.autograd.Function):
    """mean_{valid rows} (logsumexp(x_r) - x_r[label_r]); KPFCNN.loss, architectures.py:352-373."""

    @staticmethod
    def forward(ctx, logits, labels, ignore_index):
        _lib.require_cuda()
        L = _lib.lib()
        if not logits.is_cuda:
            raise RuntimeError("softmax_cross_entropy: tensors must live on a CUDA device (no CPU fallback)")
        x = _f32c(logits)
        y = labels.to(torch.int64).contiguous()
        rows, classes = x.shape
        dev = x.device
        out = torch.empty((), dtype=torch.float32, device=dev)
        lse = torch.empty(max(rows, 1), dtype=torch.float32, device=dev)
        with _lib.on_device(dev):
            acc = _lib.zeros_ptr(16, dev)  # [loss sum f64][valid rows u32][ticket u32]
            if rows > 0:
                check(L.mvk_softmax_xent(x.data_ptr(), classes, y.data_ptr(), rows, classes, int(ignore_index),
                                         lse.data_ptr(), acc, acc + 8, out.data_ptr(), stream_ptr()))
            else:
                out.fill_(float("nan"))
        ctx.save_for_backward(x, y, lse)
        ctx.cfg = (rows, classes, int(ignore_index), acc + 8, _lib._ZEROS.buf)  # the pool chunk stays alive
        return out

    @staticmethod
    def backward(ctx, g):
        L = _lib.lib()
        x, y, lse = ctx.saved_tensors
        rows, classes, ignore_index, count, _keep = ctx.cfg
        dev = x.device
        grad = torch.empty_like(x)
        if rows > 0:
            gu = g.to(torch.float32).contiguous()
            with _lib.on_device(dev):
                check(L.mvk_softmax_xent_bwd(x.data_ptr(), classes, y.data_ptr(), rows, classes, ignore_index,
                                             lse.data_ptr(), count, gu.data_ptr(), grad.data_ptr(), classes,
                                             stream_ptr()))
        return grad, None, None


def softmax_cross_entropy(logits, labels, ignore_index=-1):
    """``torch.nn.CrossEntropyLoss(ignore_index=ignore_index)(logits.T.unsqueeze(0), labels.unsqueeze(0))`` --
    the way KPFCNN.loss calls it (architectures.py:352-373) -- on [N, C] logits and [N] labels: one
    forward and one backward kernel."""
    return _SoftmaxXent.apply(logits, labels, ignore_index)
