"""Host-side logic that needs no GPU: the K-concatenated operand permutation of the channel-split KPConv, batch-norm
folding of the frozen 2D network, the linearity the split decoder step rests on, graph signatures, sphere sharding."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mvkpconv_b200 import fusion, harness, kpconv, lifting, scene  # noqa: E402


def test_channel_split_permutation_reorders_the_contraction_exactly():
    """A [N, K, cin] x W [K, cin, cout] summed over (k, c) must equal [A_main | A_rest] x W[perm] (kpconv._channel_split)."""
    dev = torch.device("cpu")
    g = torch.Generator().manual_seed(0)
    for cin, expect in ((66, (2, 64)), (65, (1, 64)), (36, (4, 32)), (130, (2, 128)), (64, None), (70, None), (3, None)):
        got = kpconv._channel_split(cin, 15, 40, 1, 0, dev)
        if expect is None:
            assert got is None, (cin, got)
            continue
        c_rest, c_main, perm = got
        assert (c_rest, c_main) == expect
        K, cout, n = 15, 8, 5
        assert sorted(perm.tolist()) == list(range(K * cin))
        a = torch.randn(n, K, cin, generator=g, dtype=torch.float64)
        w = torch.randn(K, cin, cout, generator=g, dtype=torch.float64)
        ref = a.reshape(n, K * cin) @ w.reshape(K * cin, cout)
        cat = torch.cat([a[:, :, c_rest:].reshape(n, K * c_main), a[:, :, :c_rest].reshape(n, K * c_rest)], 1)
        out = cat @ w.reshape(K * cin, cout).index_select(0, perm)
        assert torch.allclose(out, ref, rtol=0, atol=1e-12)
        # dW comes back in the permuted order and is scattered with index_copy_ (kpconv.py backward)
        dw_perm = cat.t() @ torch.ones(n, cout, dtype=torch.float64)
        dw = torch.zeros(K * cin, cout, dtype=torch.float64).index_copy_(0, perm, dw_perm)
        assert torch.allclose(dw, a.reshape(n, K * cin).t() @ torch.ones(n, cout, dtype=torch.float64), atol=1e-12)
    # the fast stage-A variants behind the split only exist for linear influence, sum aggregation, K <= 16, H <= 64
    assert kpconv._channel_split(66, 15, 40, 0, 0, dev) is None
    assert kpconv._channel_split(66, 15, 40, 1, 1, dev) is None
    assert kpconv._channel_split(66, 15, 80, 1, 0, dev) is None
    assert kpconv._channel_split(66, 20, 40, 1, 0, dev) is None


def test_split_decoder_step_identity():
    """cat([up(x), skip]) W^T = up(x W_up^T) + skip W_skip^T with up = a row gather (blocks._UpAddLinearBNAct)."""
    g = torch.Generator().manual_seed(1)
    nc, nf, c1, c2, cout = 7, 19, 24, 16, 12
    x = torch.randn(nc, c1, generator=g, dtype=torch.float64)
    skip = torch.randn(nf, c2, generator=g, dtype=torch.float64)
    up = torch.randint(0, nc + 1, (nf,), generator=g)  # index nc = the shadow row (zeros)
    w = torch.randn(cout, c1 + c2, generator=g, dtype=torch.float64)
    xs = torch.cat([x, torch.zeros(1, c1, dtype=torch.float64)], 0)
    ref = torch.cat([xs[up], skip], 1) @ w.t()
    zc = torch.cat([x @ w[:, :c1].t(), torch.zeros(1, cout, dtype=torch.float64)], 0)
    got = zc[up] + skip @ w[:, c1:].t()
    assert torch.allclose(got, ref, rtol=0, atol=1e-12)


def test_fold_batch_norm_keeps_the_function_and_the_original():
    torch.manual_seed(0)
    net = fusion.UNetResNet34(20, p=0.0, pretrained=False)
    # non-trivial running statistics, as a trained checkpoint would have
    for m in net.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.normal_(0, 0.2)
            m.running_var.uniform_(0.5, 1.5)
            m.weight.data.uniform_(0.5, 1.5)
            m.bias.data.normal_(0, 0.2)
    net.eval()
    keys = list(net.state_dict().keys())
    folded = fusion.fold_batch_norm(net)
    assert list(net.state_dict().keys()) == keys  # the original keeps its parameter names
    n_bn = sum(isinstance(m, torch.nn.BatchNorm2d) for m in net.modules())
    n_left = sum(isinstance(m, torch.nn.BatchNorm2d) for m in folded.modules())
    assert n_bn > 30 and n_left < n_bn // 4, (n_bn, n_left)
    assert not any(p.requires_grad for p in folded.parameters())
    x = torch.randn(2, 3, 64, 96)
    with torch.no_grad():
        a, b = net({"image": x}), folded({"image": x})
    a = a["feature"] if isinstance(a, dict) else a
    b = b["feature"] if isinstance(b, dict) else b
    err = (a - b).abs().max() / a.abs().max()
    assert err < 1e-4, err


def test_intrinsics_inverse_matches_numpy():
    cam = np.eye(4, dtype=np.float32)
    cam[0, 0], cam[1, 1], cam[0, 2], cam[1, 2] = 144.3, 144.1, 79.5, 59.5
    kinv = lifting.intrinsics_inverse(cam, 3, 2, torch.device("cpu"))
    assert kinv.shape == (6, 9) and kinv.dtype == torch.float64
    ref = np.linalg.inv(cam[:3, :3]).astype(np.float64).reshape(9)
    assert np.array_equal(kinv[0].numpy(), ref) and np.array_equal(kinv[5].numpy(), ref)
    cams = np.stack([cam, cam * np.float32(2)], 0)
    cams[1, 3, 3] = 1
    kinv = lifting.intrinsics_inverse(cams, 2, 3, torch.device("cpu"))
    assert np.array_equal(kinv[3].numpy(), np.linalg.inv(cams[1][:3, :3]).astype(np.float64).reshape(9))


def test_graph_signature_separates_shapes_and_dtypes():
    def pyr(n0, h0, idt=torch.int64):
        return SimpleNamespace(points=[torch.zeros(n0, 3), torch.zeros(n0 // 2, 3)],
                               neighbors=[torch.zeros(n0, h0, dtype=idt), torch.zeros(n0 // 2, 9, dtype=idt)],
                               pools=[torch.zeros(n0 // 2, 7, dtype=idt)], upsamples=[torch.zeros(n0, 1, dtype=idt)],
                               lengths=[torch.zeros(2, dtype=torch.int32), torch.zeros(2, dtype=torch.int32)])
    sig = harness.GraphedTrainStep.signature
    f, l = torch.zeros(100, 5), torch.zeros(100, dtype=torch.int64)
    base = sig(pyr(100, 30), f, l)
    assert base == sig(pyr(100, 30), f.clone(), l.clone())
    assert hash(base) == hash(sig(pyr(100, 30), f, l))
    assert base != sig(pyr(100, 31), f, l)
    assert base != sig(pyr(102, 30), torch.zeros(102, 5), torch.zeros(102, dtype=torch.int64))
    assert base != sig(pyr(100, 30, torch.int32), f, l)
    assert base != sig(pyr(100, 30), f, l, {"images": torch.zeros(2, 3, 4, 4)})
    assert sig(pyr(100, 30), f, l, {"b": f, "a": l}) == sig(pyr(100, 30), f, l, {"a": l, "b": f})


def test_sphere_centers_cover_the_scene_and_shards_partition_them():
    rng = np.random.default_rng(0)
    pts = (rng.uniform(0, 1, (20000, 3)) * [6.0, 4.0, 0.8]).astype(np.float32)  # a flat, room-like slab
    centers = scene.sphere_centers(pts, 1.0)
    d2 = ((pts[:, None, :] - centers[None, :, :]) ** 2).sum(-1).min(1)
    assert d2.max() <= 1.0 ** 2  # every scene point falls into at least one sphere
    for world in (1, 2, 3, 8):
        parts = [scene.shard(len(centers), r, world) for r in range(world)]
        flat = sorted(int(i) for p in parts for i in p)
        assert flat == list(range(len(centers)))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1
