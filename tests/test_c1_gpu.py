"""GPU: the c1 kernels (first convolution of a coupling's s/t network, tiny Cin -> 64) against float64 torch:
ks = 1 (nn_extra_nvp.py:113-115) and ks = 3, the label-conditioned couplings' c1 as the reference actually evaluates it
(nn_extra_nvp_conditional.py:427 calls the inner conv2d_wn without filter_size -> default [3, 3])."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _mask(mask_type, H, W, C):
    hh = torch.arange(H).view(H, 1, 1)
    ww = torch.arange(W).view(1, W, 1)
    cc = torch.arange(C).view(1, 1, C)
    if mask_type == 0:
        return ((hh + ww) % 2).double().expand(H, W, C)
    if mask_type == 1:
        return 1 - ((hh + ww) % 2).double().expand(H, W, C)
    return ((cc < C // 2) if mask_type == 2 else (cc >= C // 2)).double().expand(H, W, C)


@pytest.mark.parametrize('ks', [1, 3])
@pytest.mark.parametrize('N,H,W,C,mask_type', [(3, 32, 32, 1, 0), (5, 16, 16, 4, 2), (2, 16, 16, 2, 1), (7, 8, 8, 8, 3), (4, 8, 8, 4, 0)])
def test_c1_forward_and_backward(ks, N, H, W, C, mask_type):
    from bruno_b200 import _lib, slab
    g = torch.Generator().manual_seed(N * 100 + C * 10 + ks)
    K = ks * ks * C
    V = torch.randn(K, 64, generator=g, dtype=torch.float64) * 0.05
    gain = 0.5 + torch.rand(64, generator=g, dtype=torch.float64)
    bias = torch.randn(64, generator=g, dtype=torch.float64) * 0.1
    x = torch.randn(N, H, W, C, generator=g, dtype=torch.float64)
    b = _mask(mask_type, H, W, C)
    Vv = V.clone().requires_grad_(True)
    xx = x.clone().requires_grad_(True)
    Wn = gain * Vv / torch.sqrt(torch.clamp((Vv * Vv).sum(0), min=1e-12))                       # l2_normalize over (kh, kw, cin)
    w_oihw = Wn.view(ks, ks, C, 64).permute(3, 2, 0, 1)
    pre = F.conv2d((xx * b).permute(0, 3, 1, 2), w_oihw, padding=ks // 2).permute(0, 2, 3, 1) + bias
    ref = F.elu(pre)
    # ---- forward (fp32 accumulation, bf16 slab out)
    xg, Vg, gg, bg = x.float().cuda(), V.float().cuda(), gain.float().cuda(), bias.float().cuda()
    out = slab.alloc_slabs(1, N, H, W, 'cuda')[0]
    _lib.call('bruno_c1_fwd', xg, Vg, gg, bg, 0, out, 0, N, H, W, C, mask_type, ks)
    got = slab.slab_to_nhwc(out, N, H, W, check_padding=True).double().cpu()
    scale = ref.abs().max().item()
    assert (got - ref.detach()).abs().max().item() <= 2.0 ** -8 * scale + 1e-5
    # ---- backward: g_slab = dL/d(pre-activation); dx accumulates onto what it holds, dW1 / db1 are accumulated too
    gpre = torch.randn(N, H, W, 64, generator=g, dtype=torch.float64) * 0.1
    gq = gpre.to(torch.bfloat16).double()
    dWn, = torch.autograd.grad((pre * gq).sum(), Wn, retain_graph=True)
    dx_ref, = torch.autograd.grad((pre * gq).sum(), xx)
    gs = slab.nhwc_to_slab(gpre.float().cuda())
    dx0 = torch.randn(N, H, W, C, generator=g).cuda()
    dx = dx0.clone()
    dW = torch.zeros(K, 64, device='cuda')
    db = torch.zeros(64, device='cuda')
    _lib.call('bruno_c1_bwd', xg, Vg, gg, gs, dx, dW, db, N, H, W, C, mask_type, ks)
    assert ((dx - dx0).double().cpu() - dx_ref).abs().max().item() <= 1e-4 * dx_ref.abs().max().item() + 1e-5
    assert (dW.double().cpu() - dWn).abs().max().item() <= 1e-4 * dWn.abs().max().item() + 1e-5
    db_ref = gq.sum(dim=(0, 1, 2))
    assert (db.double().cpu() - db_ref).abs().max().item() <= 1e-4 * db_ref.abs().max().item() + 1e-5
