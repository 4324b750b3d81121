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


@pytest.mark.parametrize('ta,tb', [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize('M,N,K', [(640, 512, 784), (40, 512, 784), (784, 512, 640), (130, 68, 260), (512, 784, 640)])
def test_cluster_split_k_equals_workspace_split_k_bit_for_bit(ta, tb, M, N, K):
    """The split-K CTAs of a tile reduce through peer shared memory (thread-block cluster, default) or through the bound
    workspace with a ticket (variant 2): both add the K partials in K order, so the results are the same bits -- with
    every epilogue (column scale, bias, ELU, saved pre-activation via the dense coupling, accumulate)."""
    from bruno_b200 import _lib
    lib = _lib.load()
    if not hasattr(lib, 'bruno_debug_gemm_variant'):
        pytest.skip('library built with -DBRUNO_NO_DEBUG')
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    A = torch.randn((K, M) if ta else (M, K), generator=g).cuda()
    B = torch.randn((N, K) if tb else (K, N), generator=g).cuda()
    cs = (torch.rand(N, generator=g) + 0.5).cuda()
    bias = torch.randn(N, generator=g).cuda()
    C0 = torch.randn(M, N, generator=g).cuda()
    out = {}
    try:
        for variant in (1, 2):
            lib.bruno_debug_gemm_variant(variant)
            res = []
            for scale, bi, act, acc in ((None, None, 0, 0), (cs, bias, 1, 0), (None, bias, 3, 1)):
                C = C0.clone()
                for _ in range(3):               # back-to-back launches: programmatic dependent launch between them
                    _lib.call('bruno_sgemm', ta, tb, M, N, K, A, A.shape[1], B, B.shape[1], C, N, scale, bi, act, acc)
                res.append(C)
            out[variant] = res
    finally:
        lib.bruno_debug_gemm_variant(1)
    torch.cuda.synchronize()
    for a, b in zip(out[1], out[2]):
        assert torch.equal(a, b), (a - b).abs().max().item()
    ref = (A.double().t() if ta else A.double()) @ (B.double().t() if tb else B.double())
    assert torch.allclose(out[1][0].double(), ref, rtol=1e-4, atol=1e-3 * (K ** 0.5))


@pytest.mark.parametrize('N,D,U', [(640, 784, 512), (40, 784, 512), (24, 48, 32)])
def test_dense_coupling_paired_gemms_equal_separate_launches(N, D, U):
    """The dense coupling runs its s / t head GEMMs (forward), dKs / dKt and dh = ds Ks^T + dt Kt^T (backward) as ONE launch
    each (GemmPair); GEMM variant 2 has no pairs (every product its own launch, split-K through the workspace).  Same fp32
    results up to the summation grouping -- outputs, input gradient and all ten parameter gradients."""
    from bruno_b200 import _lib, nn_extra_nvp as nvp
    lib = _lib.load()
    if not hasattr(lib, 'bruno_debug_gemm_variant'):
        pytest.skip('library built with -DBRUNO_NO_DEBUG')
    g = torch.Generator().manual_seed(N + D + U)
    x0 = torch.randn(N, D, generator=g).cuda()
    layer = nvp.CouplingLayerDense('even', name='t', n_units=U)
    layer._build(D, x0.device)
    with torch.no_grad():
        for n in layer._PNAMES:                       # heads are zero-initialised in the reference: move everything off zero
            p = getattr(layer, n)
            p.copy_(torch.randn(p.shape, generator=g).cuda() * (0.05 if n in ('Ks', 'Kt', 'bs', 'bt') else 0.3) +
                    (1.0 if n in ('g1', 'g2') else 0.0))
    res, launches = {}, {}
    try:
        for variant in (1, 2):
            lib.bruno_debug_gemm_variant(variant)
            x = x0.clone().requires_grad_(True)
            params = [getattr(layer, n).requires_grad_(True) for n in layer._PNAMES]
            l0 = _lib.query('bruno_launch_count')
            y, _, ld = layer.forward_and_jacobian(x, None, torch.zeros(N, device='cuda'))
            loss = (y * torch.linspace(-1, 1, D, device='cuda')).sum() + (ld * torch.linspace(0.5, 1.5, N, device='cuda')).sum()
            grads = torch.autograd.grad(loss, [x] + params)
            launches[variant] = _lib.query('bruno_launch_count') - l0
            res[variant] = [y.detach(), ld.detach()] + [t.detach().clone() for t in grads]
    finally:
        lib.bruno_debug_gemm_variant(1)
    assert launches[2] - launches[1] == 3, launches      # s|t heads, dKs|dKt, dh: three launches fewer with the pairs
    for a, b in zip(res[1], res[2]):
        scale = b.abs().max().item() + 1e-30
        assert (a - b).abs().max().item() <= 2e-5 * scale, ((a - b).abs().max().item(), scale)
