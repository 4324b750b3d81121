"""GPU: the product path fails loudly instead of computing something silently wrong."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _layer(C=1):
    from bruno_b200 import nn_extra_nvp
    torch.manual_seed(0)
    layer = nn_extra_nvp.CouplingLayerConv('checkerboard0' if C == 1 else 'channel0', name='guard_test', num_res_blocks=2)
    x = torch.randn(4, 14, 14, C, device='cuda')
    ld = torch.zeros(4, device='cuda')
    with torch.no_grad():
        layer.forward_and_jacobian(x, None, ld)
        layer.Ks.normal_(0, 0.05)
        layer.Kt.normal_(0, 0.05)
    return layer, x, ld


def test_backward_of_a_stale_graph_raises():
    """The saved activations of a conv coupling live in one per-layer buffer: two grad-enabled forwards followed by the
    backward of the FIRST must raise (they used to return gradients of the second forward's activations)."""
    layer, x, ld = _layer()
    xa = x.clone().requires_grad_(True)
    ya, _, la = layer.forward_and_jacobian(xa, None, ld)
    xb = (x * 2).requires_grad_(True)
    yb, _, lb = layer.forward_and_jacobian(xb, None, ld)
    with pytest.raises(RuntimeError, match='stale graph'):
        (ya.square().sum() + la.sum()).backward()
    # the most recent graph is still differentiable, and equals a fresh forward + backward
    (yb.square().sum() + lb.sum()).backward()
    g_b = xb.grad.clone()
    xc = (x * 2).requires_grad_(True)
    yc, _, lc = layer.forward_and_jacobian(xc, None, ld)
    (yc.square().sum() + lc.sum()).backward()
    assert torch.equal(g_b, xc.grad)


def test_forward_backward_pairs_in_sequence_are_fine():
    layer, x, ld = _layer(C=4)
    for k in range(3):
        xa = (x + k).requires_grad_(True)
        y, _, l = layer.forward_and_jacobian(xa, None, ld)
        (y.sum() + l.sum()).backward()
        assert torch.isfinite(xa.grad).all()


def test_wrong_dtype_is_rejected_at_the_abi():
    """A float64 batch (what a NumPy iterator yields by default) must not be reinterpreted as float32."""
    from bruno_b200 import nn_extra_nvp
    x64 = torch.rand(2, 28, 28, 1, device='cuda', dtype=torch.float64) * 255
    with pytest.raises(RuntimeError, match='dtype'):
        nn_extra_nvp.preprocess_forward(x64)
    with pytest.raises(RuntimeError):
        nn_extra_nvp.preprocess_forward(torch.rand(2, 28, 28, 1))          # CPU tensor: no CPU path exists
