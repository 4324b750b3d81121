"""GPU: the fp32 SGEMM behind the dense couplings (all transpose / epilogue variants) vs torch float64."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('ta,tb', [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize('M,N,K', [(640, 512, 784), (40, 784, 512), (33, 70, 19), (1, 1, 1), (130, 65, 100)])
def test_sgemm(ta, tb, M, N, K):
    from bruno_b200 import _lib
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((K, M) if ta else (M, K), generator=g)
    B = torch.randn((N, K) if tb else (K, N), generator=g)
    cs = torch.rand(N, generator=g) + 0.5
    bias = torch.randn(N, generator=g)
    C0 = torch.randn(M, N, generator=g)
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    for scale, bi, act, acc in ((None, None, 0, 0), (cs, bias, 1, 0), (None, bias, 0, 1)):
        C = C0.clone().cuda()
        _lib.call('bruno_sgemm', ta, tb, M, N, K, A.cuda(), A.shape[1], B.cuda(), B.shape[1], C, N,
                  None if scale is None else scale.cuda(), None if bi is None else bi.cuda(), act, acc)
        r = ref.clone()
        if scale is not None:
            r = r * scale.double()
        if bi is not None:
            r = r + bi.double()
        if act:
            r = torch.nn.functional.elu(r)
        if acc:
            r = r + C0.double()
        assert torch.allclose(C.double().cpu(), r, rtol=1e-4, atol=1e-3 * (K ** 0.5)), (C.double().cpu() - r).abs().max()
