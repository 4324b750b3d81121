"""Minimal EAGER stand-in for the `tensorflow` (1.7) module -- TEST INFRASTRUCTURE ONLY.

Purpose: execute the reference's own, UNMODIFIED Python (`/root/reference/nn_extra_nvp.py`, `nn_extra_student.py`,
`nn_extra_gauss.py`, `nn_extra_nvp_conditional.py`, `config_rnn/*.py`, `config_conditional/m1_shapenet.py`) on the CPU
so that its outputs can pin the oracle (`oracle/bruno_oracle*.py`).  TensorFlow 1.7 itself cannot be installed here
(Python 3.12, no network); this module implements the ~60 `tf.*` symbols those files touch, each with the TF-1.7
semantics stated in its docstring, on torch float64 tensors (every TF dtype maps to float64 so that the comparison
with the float64 oracle is limited by rounding order only).

Graph-vs-eager: the reference builds a static graph; here every op executes immediately.  The two places where the
order matters are handled explicitly:
  * data-dependent init (`nn_extra_nvp.py:394-398`): `x` is computed BEFORE `g.assign / b.assign_add` run and
    `tf.identity(x)` returns that value -- identical in eager order.
  * `StudentRecurrentLayer.__init__` derives var/nu/corr from the variables at construction; values must therefore be
    loaded into the variable store (`preload`) before `build_model` is first called.

Used by tools/make_ref_flow_golden.py (writes tests/golden/ref_flow_*.npz) and tests/test_reference_shim_cpu.py.
Nothing under bruno_b200/ imports it.
"""
import contextlib
import math

import numpy as np
import torch

float32 = torch.float64          # every TF float maps to float64 (see module docstring)
float64 = torch.float64
int32 = torch.int64
AUTO_REUSE = 'AUTO_REUSE'


class TFTensor(torch.Tensor):
    """torch tensor + the two TF-1.x accessors the reference uses (`get_shape`, `.shape` is inherited)."""

    def get_shape(self):
        return [int(s) for s in self.shape]

    # TF tensors are immutable: `a += b` REBINDS the Python name (the reference relies on it, e.g. `y += skip` with
    # `skip = y` aliases at nn_extra_nvp.py:123-125); torch's in-place operators would mutate the shared storage.
    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __imul__(self, other):
        return self * other

    def __itruediv__(self, other):
        return self / other


def _t(x):
    if isinstance(x, TFTensor):
        return x
    if isinstance(x, torch.Tensor):
        return x.as_subclass(TFTensor)
    return torch.as_tensor(np.asarray(x, dtype=np.float64) if not isinstance(x, (int, bool)) else x).as_subclass(TFTensor)


def _f(x):
    """as float64 TFTensor"""
    x = _t(x)
    return x if x.dtype == torch.float64 else x.to(torch.float64)


def _shape(s):
    if isinstance(s, TensorShape):
        s = s.dims
    if isinstance(s, (int, np.integer)):
        return (int(s),)
    if isinstance(s, torch.Tensor) and s.dim() == 0:
        return (int(s),)
    return tuple(int(v) for v in s)


class TensorShape(object):
    def __init__(self, dims):
        self.dims = [int(d) for d in dims]

    def concatenate(self, other):
        return TensorShape(self.dims + [int(d) for d in (other.dims if isinstance(other, TensorShape) else other)])

    def __iter__(self):
        return iter(self.dims)


# ------------------------------------------------------------------------------------------------------------------
# variables and scopes (tf.get_variable / tf.variable_scope; names as TF-1.x builds them: scope path + '/' + name)
# ------------------------------------------------------------------------------------------------------------------
class Variable(TFTensor):
    """A leaf tensor with TF's assign ops.  `var_name` is the TF variable name without the ':0' suffix."""

    def assign(self, value):
        with torch.no_grad():
            self.copy_(torch.as_tensor(value).detach().reshape(self.shape))
        return self

    def assign_add(self, delta):
        with torch.no_grad():
            self.add_(torch.as_tensor(delta).detach().reshape(self.shape))
        return self


class _Store(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.vars = {}            # name -> Variable, in creation order
        self.trainable = {}
        self.preload = {}         # name -> array-like: used instead of the initializer when the variable is created
        self.scope = []
        self.random_log = []      # every tensor drawn by random_uniform / random_normal, in call order
        self.random_feed = None   # optional list of tensors to return instead of drawing
        self.rng = torch.Generator().manual_seed(0)


_store = _Store()


def reset_default_graph():
    _store.reset()


def preload_variables(values):
    """Values (keyed by full TF variable name) that `get_variable` must use when it creates those variables."""
    _store.preload = dict(values)


def global_variables():
    return list(_store.vars.values())


def variables_by_name():
    return dict(_store.vars)


def trainable_variables():
    return [v for n, v in _store.vars.items() if _store.trainable[n]]


def feed_random(tensors):
    _store.random_feed = list(tensors) if tensors is not None else None


def random_log():
    return _store.random_log


class _Scope(object):
    def __init__(self, name):
        self.name = name

    def reuse_variables(self):
        pass


@contextlib.contextmanager
def variable_scope(name, reuse=None, **kw):
    """tf.variable_scope(name_or_scope): pushes one path component.  Re-use is implicit in this eager store:
    `get_variable` returns the existing variable whenever the full name exists (what `tf.make_template('model', ...)`
    at config_rnn/train.py:57, `reuse=True` at nn_extra_nvp.py:177 and `scope.reuse_variables()` at bn2:124 achieve)."""
    _store.scope.append(name.name if isinstance(name, _Scope) else name)
    try:
        yield _Scope(name)
    finally:
        _store.scope.pop()


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **kw):
    full = '/'.join(_store.scope + [name])
    if full in _store.vars:
        v = _store.vars[full]
        if shape is not None:
            assert tuple(v.shape) == _shape(shape), (full, tuple(v.shape), _shape(shape))
        return v
    shp = _shape(shape)
    if full in _store.preload:
        val = torch.as_tensor(np.asarray(_store.preload[full]), dtype=torch.float64).reshape(shp).clone()
    else:
        init = initializer if initializer is not None else glorot_uniform_initializer()
        val = torch.as_tensor(np.asarray(init(shp, dtype=dtype, partition_info=None)), dtype=torch.float64).reshape(shp).clone()
    v = torch.Tensor._make_subclass(Variable, val, True)
    v.var_name = full
    _store.vars[full] = v
    _store.trainable[full] = bool(trainable)
    return v


def constant_initializer(value=0.0):
    def init(shape, dtype=None, partition_info=None):
        return np.broadcast_to(np.asarray(value, dtype=np.float64), _shape(shape)).copy()
    return init


def random_normal_initializer(mean=0.0, stddev=1.0, seed=None):
    def init(shape, dtype=None, partition_info=None):
        return mean + stddev * torch.randn(_shape(shape), generator=_store.rng, dtype=torch.float64).numpy()
    return init


def glorot_uniform_initializer():
    def init(shape, dtype=None, partition_info=None):
        shape = _shape(shape)
        rf = int(np.prod(shape[:-2])) if len(shape) > 2 else 1
        lim = math.sqrt(6.0 / (shape[-2] * rf + shape[-1] * rf))
        return (torch.rand(shape, generator=_store.rng, dtype=torch.float64).numpy() * 2 - 1) * lim
    return init


@contextlib.contextmanager
def control_dependencies(ops):
    """The listed ops have already executed (eager)."""
    yield


# ------------------------------------------------------------------------------------------------------------------
# elementwise / shape ops
# ------------------------------------------------------------------------------------------------------------------
def identity(x, name=None):
    return _t(x)


def reshape(x, shape, name=None):
    return _t(x).reshape(_shape(shape))


def concat(values, axis, name=None):
    return torch.cat([_t(v) for v in values], dim=axis).as_subclass(TFTensor)


def stack(values, axis=0, name=None):
    return torch.stack([_t(v) for v in values], dim=axis).as_subclass(TFTensor)


def zeros(shape, dtype=None, name=None):
    return torch.zeros(_shape(shape), dtype=torch.float64).as_subclass(TFTensor)


def ones(shape, dtype=None, name=None):
    return torch.ones(_shape(shape), dtype=torch.float64).as_subclass(TFTensor)


def ones_like(x, name=None):
    return torch.ones_like(_t(x)).as_subclass(TFTensor)


def constant(value, dtype=None, name=None):
    return _f(np.asarray(value, dtype=np.float64))


def cast(x, dtype, name=None):
    return _t(x).to(dtype).as_subclass(TFTensor)


def tile(x, multiples, name=None):
    return _t(x).repeat(*_shape(multiples)).as_subclass(TFTensor)


def range(limit, name=None):          # noqa: A001  (tf.range(limit))
    return torch.arange(int(limit)).as_subclass(TFTensor)


def mod(x, y, name=None):
    return torch.remainder(_t(x), y).as_subclass(TFTensor)


def greater(x, y, name=None):
    return (_t(x) > _t(y)).as_subclass(TFTensor)


def multiply(x, y, name=None):
    return _t(x) * _t(y)


def square(x, name=None):
    x = _f(x)
    return x * x


def sqrt(x, name=None):
    return torch.sqrt(_f(x)).as_subclass(TFTensor)


def exp(x, name=None):
    return torch.exp(_f(x)).as_subclass(TFTensor)


def log(x, name=None):
    return torch.log(_f(x)).as_subclass(TFTensor)


def log1p(x, name=None):
    return torch.log1p(_f(x)).as_subclass(TFTensor)


def tanh(x, name=None):
    return torch.tanh(_f(x)).as_subclass(TFTensor)


def sigmoid(x, name=None):
    return torch.sigmoid(_f(x)).as_subclass(TFTensor)


def lgamma(x, name=None):
    return torch.lgamma(_f(x)).as_subclass(TFTensor)


def cos(x, name=None):
    return torch.cos(_f(x)).as_subclass(TFTensor)


def sin(x, name=None):
    return torch.sin(_f(x)).as_subclass(TFTensor)


def pow(x, y, name=None):             # noqa: A001
    return torch.pow(_f(x), _f(y)).as_subclass(TFTensor)


def matmul(a, b, name=None):
    return torch.matmul(_f(a), _f(b)).as_subclass(TFTensor)


def norm(x, axis=None, name=None):
    """tf.norm default ord='euclidean': sqrt(sum(x^2)) over `axis`, NO epsilon."""
    x = _f(x)
    return torch.sqrt((x * x).sum(dim=axis)).as_subclass(TFTensor)


def _axes(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return tuple(int(a) for a in axis)
    return int(axis)


def reduce_sum(x, axis=None, name=None):
    x = _f(x)
    return (x.sum() if axis is None else x.sum(dim=_axes(axis))).as_subclass(TFTensor)


def reduce_mean(x, axis=None, name=None):
    x = _f(x)
    return (x.mean() if axis is None else x.mean(dim=_axes(axis))).as_subclass(TFTensor)


def reduce_prod(x, axis=None, name=None):
    return torch.prod(torch.as_tensor(np.asarray(x))).as_subclass(TFTensor) if not isinstance(x, torch.Tensor) \
        else torch.prod(x).as_subclass(TFTensor)


def reduce_min(x, axis=None, name=None):
    return torch.amin(_f(x), dim=_axes(axis)).as_subclass(TFTensor)


def reduce_max(x, axis=None, name=None):
    return torch.amax(_f(x), dim=_axes(axis)).as_subclass(TFTensor)


def space_to_depth(x, block_size, name=None):
    """TF NHWC: out[n, h, w, (dy*bs + dx)*C + c] = in[n, h*bs + dy, w*bs + dx, c]."""
    x = _t(x)
    n, h, w, c = x.shape
    b = block_size
    return x.reshape(n, h // b, b, w // b, b, c).permute(0, 1, 3, 2, 4, 5).reshape(n, h // b, w // b, b * b * c).as_subclass(TFTensor)


def depth_to_space(x, block_size, name=None):
    """TF NHWC inverse of space_to_depth: in channel (dy*bs + dx)*C' + c -> out[n, h*bs + dy, w*bs + dx, c]."""
    x = _t(x)
    n, h, w, c = x.shape
    b = block_size
    co = c // (b * b)
    return x.reshape(n, h, w, b, b, co).permute(0, 1, 3, 2, 4, 5).reshape(n, h * b, w * b, co).as_subclass(TFTensor)


def random_uniform(shape, minval=0, maxval=None, dtype=None, seed=None, name=None):
    """U[0,1).  TF's Philox stream for a given op seed is not reproducible outside TF; the draws are logged
    (`random_log()`) / can be fed (`feed_random`) so the same numbers reach the oracle and the GPU path."""
    shp = _shape(shape)
    if _store.random_feed is not None:
        r = _f(_store.random_feed.pop(0)).reshape(shp)
    else:
        r = torch.rand(shp, generator=_store.rng, dtype=torch.float64).as_subclass(TFTensor)
    _store.random_log.append(r.detach().clone())
    return r


def random_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None, name=None):
    shp = _shape(shape)
    if _store.random_feed is not None:
        e = _f(_store.random_feed.pop(0)).reshape(shp)
    else:
        e = torch.randn(shp, generator=_store.rng, dtype=torch.float64).as_subclass(TFTensor)
    _store.random_log.append(e.detach().clone())
    return _f(mean) + _f(stddev) * e


from . import layers, nn      # noqa: E402,F401
from . import contrib         # noqa: E402,F401
