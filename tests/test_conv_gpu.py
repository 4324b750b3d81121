"""GPU: the tcgen05 implicit-GEMM convolution kernels (forward / dgrad epilogues, wgrad, weight-norm pack) against
a float64 restatement of conv2d_wn (nn_extra_nvp.py:378-404) on bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def bf(x):
    return x.to(torch.bfloat16).to(torch.float64)


def conv_ref(x, w_hwio):
    y = F.conv2d(x.permute(0, 3, 1, 2), w_hwio.permute(3, 2, 0, 1), padding=1)
    return y.permute(0, 2, 3, 1)


def wn(V, g):
    nrm = torch.sqrt(torch.clamp((V * V).sum(dim=(0, 1, 2), keepdim=True), min=1e-12))
    return g.view(1, 1, 1, -1) * V / nrm


def elu_grad_from_out(o):
    return torch.where(o > 0, torch.ones_like(o), o + 1)


def _setup(N, H, W, seed):
    from bruno_b200 import _lib, slab
    g = torch.Generator().manual_seed(seed)
    V = torch.randn(2, 3, 3, 64, 64, generator=g, dtype=torch.float64) * 0.05
    gain = 0.5 + torch.rand(2, 64, generator=g, dtype=torch.float64) * 3
    bias = torch.randn(64, generator=g, dtype=torch.float64) * 0.2
    x = torch.randn(N, H, W, 64, generator=g, dtype=torch.float64)
    a1 = torch.randn(N, H, W, 64, generator=g, dtype=torch.float64)
    a2 = torch.randn(N, H, W, 64, generator=g, dtype=torch.float64)
    wb = _lib.query('bruno_conv_wpack_bytes')
    wf = torch.empty(2, wb, dtype=torch.uint8, device='cuda')
    wbw = torch.empty(2, wb, dtype=torch.uint8, device='cuda')
    _lib.call('bruno_conv_wn_pack', V.float().cuda().contiguous(), gain.float().cuda().contiguous(), wf, wbw, 2)
    return V, gain, bias, x, a1, a2, wf, wbw


@pytest.mark.parametrize('N,H,W', [(3, 28, 28), (5, 14, 14), (2, 32, 32), (1, 16, 16), (40, 14, 14), (1, 15, 15), (11, 7, 7),
                                   (9, 1, 5), (2, 3, 39), (13, 31, 31), (3, 5, 127), (20, 1, 7), (5, 2, 15), (33, 3, 7)])
def test_conv3x3_all_epilogues(N, H, W):
    from bruno_b200 import _lib, slab
    V, gain, bias, x, a1, a2, wf, wbw = _setup(N, H, W, seed=N * H)
    Wt = bf(wn(V[0].float().double(), gain[0].float().double()))   # the kernel rounds the normalised weights to bf16
    xs, a1s, a2s = [slab.nhwc_to_slab(t.float().cuda()) for t in (x, a1, a2)]
    xq, a1q, a2q = bf(x), bf(a1), bf(a2)
    out = slab.alloc_slabs(1, N, H, W, 'cuda')[0]
    b32 = bias.float().cuda()
    conv = conv_ref(xq, Wt)
    # dgrad operand: correlation with the flipped, transposed kernel
    Wd = Wt.flip(0, 1).permute(0, 1, 3, 2)
    dconv = conv_ref(xq, Wd)
    cases = [
        (0, wf[0], F.elu(conv + bias.float().double())),
        (1, wf[0], F.elu(conv + bias.float().double() + a1q)),
        (2, wbw[0], dconv * elu_grad_from_out(a1q)),
        (3, wbw[0], (dconv + a2q) * elu_grad_from_out(a1q)),
        (4, wf[0], conv + bias.float().double()),
    ]
    lib = _lib.load()
    try:
        for variant in (1, 2):          # per-tap kernel, line-tile kernel (the automatic choice depends on the size)
            P = slab.SlabGeom(N, H, W).P
            if (variant == 2 and (P > 128 or P % 8)) or (variant == 1 and W > 64):   # line tiles need a pitch that is a multiple of 8;
                continue                                                               # the per-tap slab ring does not fit wide images
            assert lib.bruno_debug_conv_variant(variant) == 0
            for mode, w, ref in cases:
                out.zero_()
                _lib.call('bruno_conv3x3', mode, xs, w, b32, 0, a1s, a2s, out, N, H, W)
                got = slab.slab_to_nhwc(out, N, H, W, check_padding=True).double().cpu()
                err = (got - ref).abs().max().item()
                scale = ref.abs().max().item()
                assert err <= 2.0 ** -7 * scale + 1e-3, (variant, mode, err, scale)
                # mean error must be far below the bf16 output quantum (catches systematic tap/channel mix-ups)
                assert (got - ref).abs().mean().item() <= 3e-3 * scale, (variant, mode, (got - ref).abs().mean().item())
    finally:
        lib.bruno_debug_conv_variant(0)


@pytest.mark.parametrize('N,H,W', [(3, 28, 28), (5, 14, 14), (2, 32, 32), (300, 14, 14), (20, 1, 7), (5, 2, 15), (33, 3, 7), (7, 5, 39),
                                   (9, 16, 16)])
def test_conv3x3_wgrad(N, H, W):
    from bruno_b200 import _lib, slab
    g = torch.Generator().manual_seed(7 + N)
    x = torch.randn(N, H, W, 64, generator=g, dtype=torch.float64)
    gr = torch.randn(N, H, W, 64, generator=g, dtype=torch.float64) * 0.1
    xs, gs = slab.nhwc_to_slab(x.float().cuda()), slab.nhwc_to_slab(gr.float().cuda())
    dW = torch.zeros(3, 3, 64, 64, device='cuda')
    db = torch.zeros(64, device='cuda')
    _lib.call('bruno_conv3x3_wgrad', xs, gs, dW, db, N, H, W)
    xq, gq = bf(x), bf(gr)
    w = torch.zeros(3, 3, 64, 64, dtype=torch.float64, requires_grad=True)
    (conv_ref(xq, w) * gq).sum().backward()
    ref = w.grad
    scale = ref.abs().max().item()
    assert (dW.double().cpu() - ref).abs().max().item() <= 2e-4 * scale + 1e-4
    refb = gq.sum(dim=(0, 1, 2))
    assert (db.double().cpu() - refb).abs().max().item() <= 2e-4 * refb.abs().max().item() + 1e-4
    # accumulation semantics: a second call doubles the result
    _lib.call('bruno_conv3x3_wgrad', xs, gs, dW, db, N, H, W)
    assert (dW.double().cpu() - 2 * ref).abs().max().item() <= 4e-4 * scale + 2e-4


@pytest.mark.parametrize('L,N,H,W', [(4, 6, 28, 28), (16, 6, 14, 14)])
def test_conv3x3_wgrad_batched(L, N, H, W):
    """One launch for L stacked layers (the grid is split among them) == L single-layer launches."""
    from bruno_b200 import _lib, slab
    g = torch.Generator().manual_seed(11 + L)
    xs = slab.alloc_slabs(L, N, H, W, 'cuda')
    gs = slab.alloc_slabs(L, N, H, W, 'cuda')
    for l in range(L):
        xs[l].copy_(slab.nhwc_to_slab((torch.randn(N, H, W, 64, generator=g)).cuda()))
        gs[l].copy_(slab.nhwc_to_slab((torch.randn(N, H, W, 64, generator=g) * 0.1).cuda()))
    dW = torch.zeros(L, 3, 3, 64, 64, device='cuda')
    db = torch.zeros(L, 64, device='cuda')
    _lib.call('bruno_conv3x3_wgrad_batched', xs, gs, dW, db, L, N, H, W)
    for l in range(L):
        dW1 = torch.zeros(3, 3, 64, 64, device='cuda')
        db1 = torch.zeros(64, device='cuda')
        _lib.call('bruno_conv3x3_wgrad', xs[l], gs[l], dW1, db1, N, H, W)
        scale = dW1.abs().max().item()
        assert (dW[l] - dW1).abs().max().item() <= 1e-5 * scale + 1e-6, l
        assert (db[l] - db1).abs().max().item() <= 1e-5 * db1.abs().max().item() + 1e-6, l


def test_wn_backward():
    from bruno_b200 import _lib
    g = torch.Generator().manual_seed(3)
    V = (torch.randn(2, 3, 3, 64, 64, generator=g, dtype=torch.float64) * 0.05).requires_grad_(True)
    gain = (0.5 + torch.rand(2, 64, generator=g, dtype=torch.float64)).requires_grad_(True)
    dW = torch.randn(2, 3, 3, 64, 64, generator=g, dtype=torch.float64)
    Wn = torch.stack([wn(V[i], gain[i]) for i in range(2)])
    (Wn * dW).sum().backward()
    dV = torch.empty(2, 576, 64, device='cuda')
    dg = torch.empty(2, 64, device='cuda')
    _lib.call('bruno_conv_wn_bwd', V.detach().float().cuda().contiguous(), gain.detach().float().cuda().contiguous(),
              dW.float().cuda().contiguous(), dV, dg, 2, 576)
    assert torch.allclose(dV.double().cpu().view(2, 3, 3, 64, 64), V.grad, rtol=1e-4, atol=1e-5)
    assert torch.allclose(dg.double().cpu(), gain.grad, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('N,H,W', [(640, 28, 28), (640, 14, 14), (320, 32, 32)])
def test_conv3x3_all_epilogues_at_benchmark_shapes(N, H, W):
    """The shapes bench.py runs (bn2_omniglot_tp B=32 x T=20 -> 640 images at 28x28 and 14x14; an2_cifar B=16 x T=20 at
    32x32): ~4200 / ~1100 / ~2700 tiles = 7..28 tiles per persistent CTA, so the slab-ring wrap-around, the four rotating
    TMEM accumulators, the lcm-unrolled stage/accumulator loop and the two alternating epilogue groups are all checked
    against float64 F.conv2d on bf16-rounded operands (reference computed on the GPU in float64), every epilogue mode."""
    from bruno_b200 import _lib, slab
    V, gain, bias, x, a1, a2, wf, wbw = _setup(N, H, W, seed=N + H)
    dev = 'cuda'
    Wt = bf(wn(V[0].float().double(), gain[0].float().double())).to(dev)
    xs, a1s, a2s = [slab.nhwc_to_slab(t.float().cuda()) for t in (x, a1, a2)]
    xq, a1q, a2q = bf(x).to(dev), bf(a1).to(dev), bf(a2).to(dev)
    out = slab.alloc_slabs(1, N, H, W, 'cuda')[0]
    b32 = bias.float().cuda()
    bd = bias.float().double().to(dev)
    conv = conv_ref(xq, Wt)
    dconv = conv_ref(xq, Wt.flip(0, 1).permute(0, 1, 3, 2))
    cases = [
        (0, wf[0], lambda: F.elu(conv + bd)),
        (1, wf[0], lambda: F.elu(conv + bd + a1q)),
        (2, wbw[0], lambda: dconv * elu_grad_from_out(a1q)),
        (3, wbw[0], lambda: (dconv + a2q) * elu_grad_from_out(a1q)),
        (4, wf[0], lambda: conv + bd),
    ]
    for mode, w, ref_fn in cases:
        ref = ref_fn()
        out.zero_()
        _lib.call('bruno_conv3x3', mode, xs, w, b32, 0, a1s, a2s, out, N, H, W)
        got = slab.slab_to_nhwc(out, N, H, W, check_padding=True).double()
        d = (got - ref).abs()
        scale = ref.abs().max().item()
        assert d.max().item() <= 2.0 ** -7 * scale + 1e-3, (mode, d.max().item(), scale)
        assert d.mean().item() <= 3e-3 * scale, (mode, d.mean().item())
        # per-image check: a wrong tile anywhere (ring wrap-around, tail tile) shows up as one bad image
        per_img = d.reshape(N, -1).max(dim=1).values
        assert int((per_img > 2.0 ** -7 * scale + 1e-3).sum()) == 0
        del ref, got, d


def test_conv3x3_wgrad_at_benchmark_shape():
    """wgrad at 640 x 14 x 14 and 640 x 28 x 28 (the bench shapes) against float64 autograd on the GPU."""
    from bruno_b200 import _lib, slab
    for N, H, W in ((640, 14, 14), (640, 28, 28)):
        g = torch.Generator().manual_seed(N + H)
        x = torch.randn(N, H, W, 64, generator=g)
        gr = torch.randn(N, H, W, 64, generator=g) * 0.1
        xs, gs = slab.nhwc_to_slab(x.cuda()), slab.nhwc_to_slab(gr.cuda())
        dW = torch.zeros(3, 3, 64, 64, device='cuda')
        db = torch.zeros(64, device='cuda')
        _lib.call('bruno_conv3x3_wgrad', xs, gs, dW, db, N, H, W)
        xq, gq = bf(x.double()).cuda(), bf(gr.double()).cuda()
        w = torch.zeros(3, 3, 64, 64, dtype=torch.float64, device='cuda', requires_grad=True)
        (conv_ref(xq, w) * gq).sum().backward()
        scale = w.grad.abs().max().item()
        assert (dW.double() - w.grad).abs().max().item() <= 2e-4 * scale + 1e-4
        refb = gq.sum(dim=(0, 1, 2))
        assert (db.double() - refb).abs().max().item() <= 2e-4 * refb.abs().max().item() + 1e-4


@pytest.mark.parametrize('N,H,W', [(3, 28, 28), (41, 14, 14), (2, 32, 32), (1, 15, 15), (300, 14, 14), (150, 28, 28), (13, 31, 31)])
def test_line_tile_kernel_is_bit_identical_to_per_tap_kernel(N, H, W):
    """The default line-tile kernel (one image line of several images per MMA tile, the three vertical taps as one N = 192
    instruction into three consecutive TMEM accumulator slots) sums every output in the same order as the per-tap kernel
    (the fallback for very wide images): the two must agree bit for bit, in every epilogue mode, including partial image
    groups, widths with and without a spare padding column, and CTA ranges that start / end inside an image."""
    from bruno_b200 import _lib, slab
    V, gain, bias, x, a1, a2, wf, wbw = _setup(N, H, W, seed=3 * N + H)
    xs, a1s, a2s = [slab.nhwc_to_slab(t.float().cuda()) for t in (x, a1, a2)]
    b32 = bias.float().cuda()
    out = slab.alloc_slabs(2, N, H, W, 'cuda')
    lib = _lib.load()
    try:
        for mode in range(5):
            w = wf[0] if mode in (0, 1, 4) else wbw[0]
            for variant in (2, 1):
                assert lib.bruno_debug_conv_variant(variant) == 0
                out[variant - 1].zero_()
                _lib.call('bruno_conv3x3', mode, xs, w, b32, 0, a1s, a2s, out[variant - 1], N, H, W)
            torch.cuda.synchronize()
            assert torch.equal(out[0], out[1]), (mode, int((out[0] != out[1]).sum()))
    finally:
        lib.bruno_debug_conv_variant(0)
