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
import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import _lib
from ._lib import check, ptr, stream_ptr
from . import kpconv as _kp


def _r8(v):
    return (v + 7) // 8 * 8


def _attach_hilo(t, hi, lo):
    """Remember the bf16 hi/lo split of tensor `t` on the tensor object: the next Linear that consumes
    `t` (UnaryBlock) picks it up instead of re-reading and re-splitting `t`."""
    try:
        t._mvk_hilo = (t._version, hi, lo)
    except Exception:
        pass


def _cached_hilo(t, rows, ld):
    c = getattr(t, "_mvk_hilo", None)
    if c is None or c[0] != t._version:
        return None
    hi, lo = c[1], c[2]
    if hi.dim() != 2 or hi.shape[0] != rows or hi.shape[1] != ld or hi.device != t.device:
        return None
    return hi, lo


def _bn_args(bn_block):
    """(use_bn, gamma, beta_or_bias, running_mean, running_var, momentum, eps, training, module)"""
    if bn_block is None:
        return (False, None, None, None, None, 0.0, 0.0, False, None)
    if bn_block.use_bn:
        m = bn_block.batch_norm
        return (True, m.weight, m.bias, m.running_mean, m.running_var, float(m.momentum), float(m.eps),
                bool(m.training), m)
    return (False, None, bn_block.bias, None, None, 0.0, 0.0, False, None)


def _norm_forward(L, y, rows, cols, use_bn, training, gamma, beta, rm, rv, momentum, eps, st, nbt=None):
    """-> scale, shift, mean, invstd (device [cols] vectors, None where unused)."""
    dev = y.device
    if not use_bn:
        return None, (beta.detach().contiguous().float() if beta is not None else None), None, None
    scale = torch.empty(cols, dtype=torch.float32, device=dev)
    shift = torch.empty_like(scale)
    mean = torch.empty_like(scale)
    invstd = torch.empty_like(scale)
    if training and rows > 0:
        # column sums and, in the CTA that retires last, scale / shift / running statistics: one launch
        stats = _lib.zeros_f64(2 * cols + 1, dev)
        check(L.mvk_bn_batch_stats(ptr(y), rows, cols, y.stride(0), ptr(stats), ptr(gamma.detach()), ptr(beta.detach()),
                                   eps, momentum, ptr(rm), ptr(rv), ptr(scale), ptr(shift), ptr(mean), ptr(invstd),
                                   ptr(nbt), st))
    else:
        check(L.mvk_bn_finalize(None, rows, cols, ptr(gamma.detach()), ptr(beta.detach()), eps, momentum, 0, ptr(rm),
                                ptr(rv), ptr(scale), ptr(shift), ptr(mean), ptr(invstd), st))
    return scale, shift, mean, invstd


class _BNAct(torch.autograd.Function):
    """z = leaky(bn(y) [+ residual]); blocks.py:446-460 + the activation / residual that follows."""

    @staticmethod
    def forward(ctx, y, gamma, beta, residual, rm, rv, use_bn, training, momentum, eps, slope, nbt=None,
                emit_hilo=False):
        _lib.require_cuda()
        L = _lib.lib()
        if not y.is_cuda:
            raise RuntimeError("bn_act: tensors must live on a CUDA device (no CPU fallback)")
        yf = y.detach().contiguous().float()
        rows, cols = yf.shape
        res = None if residual is None else residual.detach().contiguous().float()
        st = stream_ptr()
        with _lib.on_device(yf.device):
            scale, shift, mean, invstd = _norm_forward(L, yf, rows, cols, use_bn, training, gamma, beta, rm, rv,
                                                       momentum, eps, st, nbt)
            z = torch.empty_like(yf)
            hi = lo = None
            if emit_hilo and cols % 8 == 0:
                hi = torch.empty((rows, cols), dtype=torch.bfloat16, device=yf.device)
                lo = torch.empty_like(hi)
            check(L.mvk_scale_shift_act(ptr(yf), rows, cols, cols, ptr(scale), ptr(shift), ptr(res), cols, slope,
                                        ptr(z), cols, ptr(hi), ptr(lo), cols, st))
        ctx.save_for_backward(yf, scale, shift, mean, invstd, res)
        ctx.cfg = (rows, cols, use_bn, training, slope, gamma is not None, beta is not None)
        if emit_hilo:
            if hi is None:
                hi = lo = torch.empty(0, device=yf.device)
            ctx.mark_non_differentiable(hi, lo)
            return z, hi, lo
        return z

    @staticmethod
    def backward(ctx, dz, *_unused):
        L = _lib.lib()
        yf, scale, shift, mean, invstd, res = ctx.saved_tensors
        rows, cols, use_bn, training, slope, has_g, has_b = ctx.cfg
        g = dz.detach().contiguous().float()
        dev = g.device
        st = stream_ptr()
        need_y, need_res = ctx.needs_input_grad[0], ctx.needs_input_grad[3]
        batch_stats = 1 if (use_bn and training) else 0
        with _lib.on_device(dev):
            sums = _lib.zeros_f64(2 * cols, dev)
            if has_g or has_b or batch_stats:
                check(L.mvk_act_bwd_reduce(ptr(g), cols, ptr(yf), rows, cols, cols, ptr(scale), ptr(shift), ptr(res),
                                           cols, ptr(mean), ptr(invstd), slope, ptr(sums), st))
            dy = torch.empty_like(yf) if need_y else None
            dres = torch.empty_like(yf) if (need_res and res is not None) else None
            dgamma = torch.empty(cols, dtype=torch.float32, device=dev) if has_g else None
            dbeta = torch.empty(cols, dtype=torch.float32, device=dev) if has_b else None
            check(L.mvk_act_bwd_apply(ptr(g), cols, ptr(yf), rows, cols, cols, ptr(scale), ptr(shift), ptr(res), cols,
                                      ptr(mean), ptr(invstd), slope, ptr(sums), batch_stats, ptr(dy), cols, None, None,
                                      0, ptr(dres), cols, ptr(dgamma), ptr(dbeta), st))
        return dy, dgamma, dbeta, dres, None, None, None, None, None, None, None, None, None


class _LinearBNAct(torch.autograd.Function):
    """z = leaky(bn(x W^T) [+ residual]); UnaryBlock.forward, blocks.py:493-498."""

    @staticmethod
    def forward(ctx, x, weight, gamma, beta, residual, rm, rv, use_bn, training, momentum, eps, slope, contraction,
                nbt=None, emit_hilo=False):
        _lib.require_cuda()
        L = _lib.lib()
        if not x.is_cuda:
            raise RuntimeError("UnaryBlock: tensors must live on a CUDA device (no CPU fallback)")
        xf = x.detach().contiguous().float()
        w = weight.detach().contiguous().float()
        rows, cin = xf.shape
        cout = w.shape[0]
        dev = xf.device
        res = None if residual is None else residual.detach().contiguous().float()
        st = stream_ptr()
        with _lib.on_device(dev):
            y = torch.empty((rows, cout), dtype=torch.float32, device=dev)
            if contraction == "fp32":
                if rows > 0:
                    check(L.mvk_gemm_f32(ptr(xf), cin, 1, ptr(w), 1, cin, rows, cout, cin, ptr(y), cout, 1, st))
                ops = (xf, w)
            else:
                terms = 3 if contraction == "bf16x3" else 1
                ldx = _r8(cin)
                cached = _cached_hilo(x, rows, ldx)
                if cached is not None:
                    x_hi, x_lo = cached  # produced by the previous block's epilogue or an earlier consumer
                else:
                    x_hi = torch.empty((rows, ldx), dtype=torch.bfloat16, device=dev)
                    x_lo = torch.empty_like(x_hi)
                    check(L.mvk_split_bf16(ptr(xf), rows, cin, cin, ptr(x_hi), ptr(x_lo), rows, ldx, st))
                    _attach_hilo(x, x_hi, x_lo)
                w_hi = torch.empty((cout, ldx), dtype=torch.bfloat16, device=dev)
                w_lo = torch.empty_like(w_hi)
                check(L.mvk_split_bf16(ptr(w), cout, cin, cin, ptr(w_hi), ptr(w_lo), cout, ldx, st))
                if rows > 0:
                    check(L.mvk_gemm_bf16x3(ptr(x_hi), ptr(x_lo), 0, ldx, ptr(w_hi), ptr(w_lo), 0, ldx, rows, cout, cin,
                                            ptr(y), cout, cout, terms, 0, st))
                ops = (x_hi, x_lo, w_hi, w_lo)
            scale, shift, mean, invstd = _norm_forward(L, y, rows, cout, use_bn, training, gamma, beta, rm, rv,
                                                       momentum, eps, st, nbt)
            z = torch.empty_like(y)
            hi = lo = None
            if emit_hilo and cout % 8 == 0:
                hi = torch.empty((rows, cout), dtype=torch.bfloat16, device=dev)
                lo = torch.empty_like(hi)
            check(L.mvk_scale_shift_act(ptr(y), rows, cout, cout, ptr(scale), ptr(shift), ptr(res), cout, slope,
                                        ptr(z), cout, ptr(hi), ptr(lo), cout, st))
        ctx.save_for_backward(y, scale, shift, mean, invstd, res, *ops)
        ctx.cfg = (rows, cin, cout, use_bn, training, slope, gamma is not None, beta is not None, contraction)
        if emit_hilo:
            if hi is None:
                hi = lo = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(hi, lo)
            return z, hi, lo
        return z

    @staticmethod
    def backward(ctx, dz, *_unused):
        L = _lib.lib()
        y, scale, shift, mean, invstd, res, *ops = ctx.saved_tensors
        rows, cin, cout, use_bn, training, slope, has_g, has_b, contraction = ctx.cfg
        g = dz.detach().contiguous().float()
        dev = g.device
        st = stream_ptr()
        need_x, need_w, need_res = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[4]
        batch_stats = 1 if (use_bn and training) else 0
        dx = dw = None
        with _lib.on_device(dev):
            sums = _lib.zeros_f64(2 * cout, dev)
            if has_g or has_b or batch_stats:
                check(L.mvk_act_bwd_reduce(ptr(g), cout, ptr(y), rows, cout, cout, ptr(scale), ptr(shift), ptr(res),
                                           cout, ptr(mean), ptr(invstd), slope, ptr(sums), st))
            dres = torch.empty_like(y) if (need_res and res is not None) else None
            dgamma = torch.empty(cout, dtype=torch.float32, device=dev) if has_g else None
            dbeta = torch.empty(cout, dtype=torch.float32, device=dev) if has_b else None
            if contraction == "fp32":
                xf, w = ops
                dy = torch.empty_like(y)
                check(L.mvk_act_bwd_apply(ptr(g), cout, ptr(y), rows, cout, cout, ptr(scale), ptr(shift), ptr(res),
                                          cout, ptr(mean), ptr(invstd), slope, ptr(sums), batch_stats, ptr(dy), cout,
                                          None, None, 0, ptr(dres), cout, ptr(dgamma), ptr(dbeta), st))
                if need_x:
                    dx = torch.empty((rows, cin), dtype=torch.float32, device=dev)
                    if rows > 0:
                        check(L.mvk_gemm_f32(ptr(dy), cout, 1, ptr(w), cin, 1, rows, cin, cout, ptr(dx), cin, 1, st))
                if need_w:
                    dw = torch.zeros((cout, cin), dtype=torch.float32, device=dev)
                    if rows > 0:
                        split = max(1, min(rows // 64, 2 * 148 // max(1, ((cout + 63) // 64) * ((cin + 63) // 64))))
                        check(L.mvk_gemm_f32(ptr(dy), 1, cout, ptr(xf), cin, 1, cout, cin, rows, ptr(dw), cin, split, st))
            else:
                x_hi, x_lo, w_hi, w_lo = ops
                terms = 3 if contraction == "bf16x3" else 1
                ldx, ldh = x_hi.shape[1], _r8(cout)
                dy_hi = torch.empty((rows, ldh), dtype=torch.bfloat16, device=dev)
                dy_lo = torch.empty_like(dy_hi)
                check(L.mvk_act_bwd_apply(ptr(g), cout, ptr(y), rows, cout, cout, ptr(scale), ptr(shift), ptr(res),
                                          cout, ptr(mean), ptr(invstd), slope, ptr(sums), batch_stats, None, 0,
                                          ptr(dy_hi), ptr(dy_lo), ldh, ptr(dres), cout, ptr(dgamma), ptr(dbeta), st))
                if need_x:
                    dx = torch.empty((rows, cin), dtype=torch.float32, device=dev)
                    if rows > 0:
                        # dx = dy W : A = dy [rows, cout] K-major, B = W stored [K = cout, N = cin] (N contiguous)
                        check(L.mvk_gemm_bf16x3(ptr(dy_hi), ptr(dy_lo), 0, ldh, ptr(w_hi), ptr(w_lo), 1, ldx, rows, cin,
                                                cout, ptr(dx), cin, cin, terms, 0, st))
                if need_w:
                    dw = torch.zeros((cout, cin), dtype=torch.float32, device=dev)
                    if rows > 0:
                        kb_total = (rows + 63) // 64
                        split = _kp._split_k_for((cout + 127) // 128, (cin + 127) // 128 if cin > 64 else 1, kb_total)
                        # dW = dy^T x : both operands stored [K = rows, *] with the M / N index contiguous
                        check(L.mvk_gemm_bf16x3(ptr(dy_hi), ptr(dy_lo), 1, ldh, ptr(x_hi), ptr(x_lo), 1, ldx, cout, cin,
                                                rows, ptr(dw), cin, cin, terms, split, st))
        return dx, dw, dgamma, dbeta, dres, None, None, None, None, None, None, None, None, None, None


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


def bn_act(y, bn_block, slope=0.1, residual=None, emit_hilo=False):
    """leaky_relu(bn_block(y) [+ residual], slope); slope = 1 disables the activation.
    emit_hilo: also write the result as a bf16 hi/lo pair for a following UnaryBlock (same kernel)."""
    use_bn, gamma, beta, rm, rv, momentum, eps, training, mod = _bn_args(bn_block)
    if not emit_hilo:
        return _BNAct.apply(y, gamma, beta, residual, rm, rv, use_bn, training, momentum, eps, float(slope),
                            _nbt(mod, y.shape[0]))
    z, hi, lo = _BNAct.apply(y, gamma, beta, residual, rm, rv, use_bn, training, momentum, eps, float(slope),
                             _nbt(mod, y.shape[0]), True)
    if hi.numel():
        _attach_hilo(z, hi, lo)
    return z


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
        if not emit_hilo:
            return _LinearBNAct.apply(x, self.mlp.weight, gamma, beta, residual, rm, rv, use_bn, training, momentum,
                                      eps, float(slope), self.contraction, _nbt(mod, x.shape[0]))
        z, hi, lo = _LinearBNAct.apply(x, self.mlp.weight, gamma, beta, residual, rm, rv, use_bn, training, momentum,
                                       eps, float(slope), self.contraction, _nbt(mod, x.shape[0]), True)
        if hi.numel():
            _attach_hilo(z, hi, lo)
        return z

    def __repr__(self):
        return 'UnaryBlock(in_feat: {:d}, out_feat: {:d}, BN: {:s}, ReLU: {:s})'.format(self.in_dim, self.out_dim,
                                                                                        str(self.use_bn),
                                                                                        str(not self.no_relu))
