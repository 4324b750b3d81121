"""tf.layers.dense / tf.layers.conv2d (TF-1.7): variables `<name>/kernel`, `<name>/bias` in the current scope."""
from . import get_variable, variable_scope, constant_initializer, matmul, _f
from . import nn as _nn


def dense(inputs, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None, name=None, **kw):
    x = _f(inputs)
    with variable_scope(name):
        k = get_variable('kernel', [int(x.shape[-1]), units], initializer=kernel_initializer)
        y = matmul(x, k)
        if use_bias:
            b = get_variable('bias', [units], initializer=bias_initializer or constant_initializer(0.))
            y = y + b
    return activation(y) if activation is not None else y


def conv2d(inputs, filters, kernel_size, strides=(1, 1), padding='valid', activation=None, use_bias=True,
           kernel_initializer=None, bias_initializer=None, name=None, **kw):
    x = _f(inputs)
    ks = [kernel_size, kernel_size] if isinstance(kernel_size, int) else list(kernel_size)
    with variable_scope(name):
        k = get_variable('kernel', ks + [int(x.shape[-1]), filters], initializer=kernel_initializer)
        y = _nn.conv2d(x, k, [1, 1, 1, 1], 'SAME' if padding.lower() == 'same' or ks == [1, 1] else padding)
        if use_bias:
            b = get_variable('bias', [filters], initializer=bias_initializer or constant_initializer(0.))
            y = y + b
    return activation(y) if activation is not None else y
