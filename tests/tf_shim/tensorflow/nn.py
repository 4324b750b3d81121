"""tf.nn.* symbols used by the reference hot path (TF-1.7 semantics), see tensorflow/__init__.py."""
import torch
import torch.nn.functional as F

from . import TFTensor, _f


def elu(x, name=None):
    return F.elu(_f(x)).as_subclass(TFTensor)


def relu(x, name=None):
    return F.relu(_f(x)).as_subclass(TFTensor)


def leaky_relu(x, alpha=0.2, name=None):
    """tf.nn.leaky_relu default alpha = 0.2 (TF-1.7)."""
    return F.leaky_relu(_f(x), alpha).as_subclass(TFTensor)


def softplus(x, name=None):
    return F.softplus(_f(x), beta=1.0, threshold=1e9).as_subclass(TFTensor)


def softmax(x, axis=-1, name=None, dim=None):
    return torch.softmax(_f(x), dim=axis if dim is None else dim).as_subclass(TFTensor)


def l2_normalize(x, axis=None, epsilon=1e-12, name=None, dim=None):
    """x * rsqrt(max(sum(x^2, axis, keepdims), epsilon))."""
    x = _f(x)
    ax = axis if dim is None else dim
    ax = tuple(ax) if isinstance(ax, (list, tuple)) else ax
    ss = (x * x).sum(dim=ax, keepdim=True)
    return (x * torch.rsqrt(torch.clamp_min(ss, epsilon))).as_subclass(TFTensor)


def conv2d(x, filter, strides, padding, name=None):   # noqa: A002
    """NHWC input, HWIO filter, cross-correlation.  'SAME' with stride 1 and odd kernels = symmetric zero padding."""
    x, w = _f(x), _f(filter)
    assert list(strides) == [1, 1, 1, 1] and padding.upper() == 'SAME'
    kh, kw = w.shape[0], w.shape[1]
    assert kh % 2 == 1 and kw % 2 == 1
    y = F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=(kh // 2, kw // 2))
    return y.permute(0, 2, 3, 1).as_subclass(TFTensor)


def bias_add(x, b, name=None):
    return _f(x) + _f(b)


def moments(x, axes, name=None, keep_dims=False):
    """mean and BIASED variance over `axes`."""
    x = _f(x)
    ax = tuple(int(a) for a in axes)
    m = x.mean(dim=ax, keepdim=True)
    v = ((x - m) ** 2).mean(dim=ax, keepdim=keep_dims)
    return (m if keep_dims else m.reshape(v.shape)).as_subclass(TFTensor), v.as_subclass(TFTensor)
