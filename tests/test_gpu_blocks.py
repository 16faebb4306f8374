"""GPU parity of the point-wise blocks (UnaryBlock / BatchNormBlock / fused block tail) against
golden vectors produced by the reference's own modules (models/blocks.py:430-504) and against
torch autograd for the fused residual tail.  Tolerance: 1e-4 relative (max-abs / max|ref|) for
the fp32 and bf16x3 contractions, like KPConv."""
import numpy as np
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a = a.detach().double().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    b = b.detach().double().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


def build_block(mvk, c, contraction):
    from mvkpconv_b200 import blocks
    cout, cin = c["weight"].shape
    blk = blocks.UnaryBlock(cin, cout, bool(c["use_bn"]), 0.02, no_relu=bool(c["no_relu"]), contraction=contraction).cuda()
    with torch.no_grad():
        blk.mlp.weight.copy_(torch.from_numpy(c["weight"]))
        if bool(c["use_bn"]):
            bn = blk.batch_norm.batch_norm
            bn.weight.copy_(torch.from_numpy(c["gamma"]))
            bn.bias.copy_(torch.from_numpy(c["beta"]))
            bn.running_mean.copy_(torch.from_numpy(c["rm0"]))
            bn.running_var.copy_(torch.from_numpy(c["rv0"]))
        else:
            blk.batch_norm.bias.copy_(torch.from_numpy(c["bias"]))
    blk.train(bool(c["train"]))
    return blk


@pytest.mark.parametrize("contraction", ["fp32", "bf16x3"])
def test_unary_block_vs_reference_golden(mvk, contraction):
    g = load_golden("blocks")
    for name, c in g.items():
        if name.startswith("_"):
            continue
        blk = build_block(mvk, c, contraction)
        x = torch.from_numpy(c["x"]).cuda().requires_grad_(True)
        out = blk(x)
        out.backward(torch.from_numpy(c["grad_out"]).cuda())
        tol = 1e-4
        assert rel_err(out, c["out"]) < tol, (name, "out")
        assert rel_err(x.grad, c["grad_x"]) < tol, (name, "grad_x")
        assert rel_err(blk.mlp.weight.grad, c["grad_w"]) < tol, (name, "grad_w")
        if bool(c["use_bn"]):
            bn = blk.batch_norm.batch_norm
            assert rel_err(bn.weight.grad, c["grad_gamma"]) < tol, (name, "grad_gamma")
            assert rel_err(bn.bias.grad, c["grad_beta"]) < tol, (name, "grad_beta")
            assert rel_err(bn.running_mean, c["rm1"]) < 1e-5, (name, "running_mean")
            assert rel_err(bn.running_var, c["rv1"]) < 1e-5, (name, "running_var")
            assert int(bn.num_batches_tracked) == int(c["nbt"]), name
        else:
            assert rel_err(blk.batch_norm.bias.grad, c["grad_bias"]) < tol, (name, "grad_bias")


def test_bn_act_and_residual_tail_vs_torch(mvk):
    """leaky(bn(y) + shortcut) and leaky(bn(x W^T) + shortcut) against torch autograd (fp32 on the GPU)."""
    from mvkpconv_b200 import blocks
    torch.manual_seed(3)
    rows, cin, cout = 3000, 32, 128
    x = torch.randn(rows, cin, device="cuda")
    sc = torch.randn(rows, cout, device="cuda")
    go = torch.randn(rows, cout, device="cuda")
    blk = blocks.UnaryBlock(cin, cout, True, 0.02, no_relu=True).cuda()
    ref_lin = torch.nn.Linear(cin, cout, bias=False).cuda()
    ref_bn = torch.nn.BatchNorm1d(cout, momentum=0.02).cuda()
    with torch.no_grad():
        ref_lin.weight.copy_(blk.mlp.weight)
        blk.batch_norm.batch_norm.weight.uniform_(0.5, 1.5)
        blk.batch_norm.batch_norm.bias.uniform_(-0.3, 0.3)
        ref_bn.weight.copy_(blk.batch_norm.batch_norm.weight)
        ref_bn.bias.copy_(blk.batch_norm.batch_norm.bias)
    torch.backends.cuda.matmul.allow_tf32 = False
    xa, sa = x.clone().requires_grad_(True), sc.clone().requires_grad_(True)
    xb, sb = x.clone().requires_grad_(True), sc.clone().requires_grad_(True)
    out = blk(xa, residual=sa, slope=0.1)
    out.backward(go)
    ref = torch.nn.functional.leaky_relu(ref_bn(ref_lin(xb)) + sb, 0.1)
    ref.backward(go)
    assert rel_err(out, ref) < 1e-4
    assert rel_err(xa.grad, xb.grad) < 1e-4
    assert rel_err(sa.grad, sb.grad) < 1e-4
    assert rel_err(blk.mlp.weight.grad, ref_lin.weight.grad) < 1e-4
    assert rel_err(blk.batch_norm.batch_norm.weight.grad, ref_bn.weight.grad) < 1e-4
    assert rel_err(blk.batch_norm.batch_norm.bias.grad, ref_bn.bias.grad) < 1e-4
    # stand-alone bn + activation (the KPConv -> batch norm -> LeakyReLU step of the blocks)
    bnb = blocks.BatchNormBlock(cout, True, 0.02).cuda()
    ref_bn2 = torch.nn.BatchNorm1d(cout, momentum=0.02).cuda()
    ya, yb = sc.clone().requires_grad_(True), sc.clone().requires_grad_(True)
    z = blocks.bn_act(ya, bnb, slope=0.1)
    z.backward(go)
    zr = torch.nn.functional.leaky_relu(ref_bn2(yb), 0.1)
    zr.backward(go)
    assert rel_err(z, zr) < 1e-5
    assert rel_err(ya.grad, yb.grad) < 1e-4
    assert rel_err(bnb.batch_norm.running_var, ref_bn2.running_var) < 1e-5
    # odd column count (num_classes = 19 -> scalar path), bias only
    bb = blocks.BatchNormBlock(19, False, 0.02).cuda()
    with torch.no_grad():
        bb.bias.uniform_(-1, 1)
    y19 = torch.randn(777, 19, device="cuda", requires_grad=True)
    z19 = bb(y19)
    assert rel_err(z19, y19.detach() + bb.bias.detach()) < 1e-6
    z19.backward(torch.ones_like(z19))
    assert rel_err(bb.bias.grad, torch.full((19,), 777.0)) < 1e-6
    assert rel_err(y19.grad, torch.ones_like(y19)) < 1e-6


def test_gemm_fused_column_statistics(mvk):
    """mvk_gemm_bf16x3_stats: column sums / sums of squares out of the contraction's epilogue (and the
    fallback pass for split problems) against torch."""
    L = mvk._lib.lib()
    from mvkpconv_b200._lib import check, ptr, stream_ptr
    torch.manual_seed(1)
    for (M, N, K) in [(5000, 128, 64), (3000, 32, 96), (700, 512, 256), (129, 64, 4096)]:
        A = torch.randn(M, K, device="cuda")
        B = torch.randn(N, K, device="cuda")
        hi = lambda t: t.bfloat16()
        lo = lambda t: (t - t.bfloat16().float()).bfloat16()
        D = torch.zeros(M, N, device="cuda")
        stats = torch.zeros(2 * N, dtype=torch.float64, device="cuda")
        a_hi, a_lo, b_hi, b_lo = hi(A), lo(A), hi(B), lo(B)  # keep the operands alive across the call
        check(L.mvk_gemm_bf16x3_stats(ptr(a_hi), ptr(a_lo), 0, K, ptr(b_hi), ptr(b_lo), 0, K, M, N, K, ptr(D), N, N, 3, 0,
                                      ptr(stats), stream_ptr()))
        ref = A.double() @ B.double().t()
        assert rel_err(D, ref) < 2e-5
        assert rel_err(stats[:N], ref.sum(0)) < 1e-4, (M, N, K)
        assert rel_err(stats[N:], (ref * ref).sum(0)) < 1e-4, (M, N, K)


def test_softmax_cross_entropy_vs_torch(mvk):
    """KPFCNN.loss (architectures.py:352-373): CrossEntropyLoss(ignore_index=-1) on the transposed logits.
    Forward and gradient within 1e-5 relative of torch's fp32 implementation, incl. ignored rows, an
    upstream gradient != 1, and the all-ignored case (NaN loss, zero gradient rows like torch)."""
    import torch
    g = torch.Generator().manual_seed(3)
    for rows, classes in ((5000, 20), (33, 6), (1, 3), (70000, 13)):
        x = (torch.randn(rows, classes, generator=g) * 3).cuda().requires_grad_(True)
        y = torch.randint(0, classes, (rows,), generator=g)
        y[torch.rand(rows, generator=g) < 0.2] = -1
        y = y.cuda()
        ref_x = x.detach().clone().requires_grad_(True)
        ref = torch.nn.CrossEntropyLoss(ignore_index=-1)(ref_x.transpose(0, 1).unsqueeze(0), y.unsqueeze(0))
        out = mvk.softmax_cross_entropy(x, y, ignore_index=-1)
        (out * 1.7).backward()
        (ref * 1.7).backward()
        if bool((y >= 0).any()):
            assert abs(out.item() - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
            assert float((x.grad - ref_x.grad).abs().max()) <= 1e-5 * float(ref_x.grad.abs().max()) + 1e-9
        else:
            assert np.isnan(out.item()) and np.isnan(ref.item())
        assert float(x.grad[y < 0].abs().sum()) == 0.0
    x = torch.randn(10, 4).cuda().requires_grad_(True)
    y = torch.full((10,), -1).cuda()
    out = mvk.softmax_cross_entropy(x, y)
    assert np.isnan(out.item())


def test_weight_operand_pairs_follow_the_optimizer(mvk):
    """The bf16 hi/lo pairs of the weights are refreshed once per optimiser step (one multi-tensor launch for
    every registered parameter): after in-place updates the outputs must follow the NEW weights, for every
    registered layer, and a backward through a graph whose weight changed in between must raise."""
    import torch
    from mvkpconv_b200 import _weights
    torch.manual_seed(0)
    dev = torch.device("cuda")
    a = mvk.UnaryBlock(64, 48, True, 0.1).to(dev)
    b = mvk.UnaryBlock(48, 40, False, 0.1).to(dev)
    conv = mvk.KPConv(15, 3, 8, 24, 0.06, 0.15).to(dev)
    pts = torch.rand(300, 3, device=dev) * 0.4
    inds = mvk.batch_neighbors(pts, pts, torch.tensor([300], dtype=torch.int32, device=dev),
                               torch.tensor([300], dtype=torch.int32, device=dev), 0.15).long()
    x = torch.randn(300, 64, device=dev)
    xc = torch.randn(300, 8, device=dev)

    def run(contraction):
        for m in (a, b, conv):
            for mm in m.modules():
                if hasattr(mm, "contraction"):
                    mm.contraction = contraction
        a.eval(); b.eval()
        return b(a(x)).detach().clone(), conv(pts, pts, inds, xc).detach().clone()

    for step in range(3):
        got = run("bf16x3")
        ref = run("fp32")                     # fp32 path reads the parameters directly
        for g, r in zip(got, ref):
            assert float((g - r).abs().max()) <= 1e-4 * float(r.abs().max())
        params = [p for p in list(a.parameters()) + list(b.parameters()) + list(conv.parameters()) if p.requires_grad]
        if step == 0:
            with torch.no_grad():             # plain in-place updates: the version counters move
                for p in params:
                    p.add_(0.5 * torch.randn_like(p))
        else:                                 # a FUSED optimiser does not move them: the step post-hook must
            opt = torch.optim.SGD(params, lr=0.5, fused=True)
            for p in params:
                p.grad = torch.randn_like(p)
            before = [p._version for p in params]
            opt.step()
            if [p._version for p in params] == before:
                pass                          # (documented torch behaviour this guard exists for)
    t = _weights._table(dev)
    assert len(t.order) >= 3 and not t.dirty  # the second and third refresh went through the multi-tensor table
    # stale pair guard
    a.train()
    y = a(x.requires_grad_(True)).sum()
    with torch.no_grad():
        a.mlp.weight.mul_(2.0)
    with pytest.raises(RuntimeError):
        y.backward()


def test_fused_decoder_step_matches_unfused(mvk):
    """UnaryBlock.forward_upsampled == unary(cat([closest_pool(x_coarse, up), skip], 1)) (architectures.py:300-306):
    outputs and all gradients (coarse features, skip features, weight, gamma, beta), incl. shadow indices."""
    import torch
    torch.manual_seed(1)
    dev = torch.device("cuda")
    ns, nq, c1, c2, cout = 700, 2500, 64, 32, 48
    blk = mvk.UnaryBlock(c1 + c2, cout, True, 0.1).to(dev)
    up = torch.randint(0, ns + 1, (nq, 5), device=dev)   # ns = shadow index -> zero row
    up[::7, 0] = ns
    xc0 = torch.randn(ns, c1, device=dev)
    sk0 = torch.randn(nq, c2, device=dev)
    gz = torch.randn(nq, cout, device=dev)
    from mvkpconv_b200 import blocks as blocks_mod
    res = {}
    # "split": upsampled half contracted at the coarse level + gathered (_UpAddLinearBNAct, the default);
    # "cat": concatenated bf16 operand (_UpCatLinearBNAct); "plain": the reference's op sequence
    for mode in ("split", "cat", "plain"):
        blk.zero_grad(set_to_none=True)
        xc, sk = xc0.clone().requires_grad_(True), sk0.clone().requires_grad_(True)
        if mode == "plain":
            z = blk(torch.cat([mvk.closest_pool(xc, up), sk], dim=1))
        else:
            blocks_mod._DECODER_SPLIT = mode == "split"
            try:
                z = blk.forward_upsampled(xc, up, sk)
            finally:
                blocks_mod._DECODER_SPLIT = True
            assert type(z.grad_fn).__name__.startswith("_UpAdd" if mode == "split" else "_UpCat")
        (z * gz).sum().backward()
        res[mode] = [z.detach(), xc.grad, sk.grad, blk.mlp.weight.grad.clone(), blk.batch_norm.batch_norm.weight.grad.clone(),
                     blk.batch_norm.batch_norm.bias.grad.clone()]
    for mode in ("split", "cat"):
        for a, b in zip(res[mode], res["plain"]):
            assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()) + 1e-7, mode
    # inference (no graph) and an empty coarse level
    with torch.no_grad():
        zi = blk.forward_upsampled(xc0, up, sk0)
    assert float((zi - res["plain"][0]).abs().max()) <= 2e-5 * float(res["plain"][0].abs().max())
    blk.contraction = "fp32"   # falls back to the unfused sequence
    z = blk.forward_upsampled(xc0, up, sk0)
    assert float((z.detach() - res["plain"][0]).abs().max()) <= 1e-4 * float(res["plain"][0].abs().max())
