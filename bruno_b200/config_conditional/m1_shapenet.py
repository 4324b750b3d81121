"""B200 mirror of the reference's config_conditional/m1_shapenet.py (Conditional BRUNO on ShapeNet-shaped views):
same module attributes, build_model(x, y_label, init, sampling_mode) and loss.  NLL, sampling, data-dependent init
(init=True) and the training graph (loss.backward() after set_trainable())."""
import numpy as np
import torch

import contextlib

from .. import nn_extra_gauss, nn_extra_nvp_conditional as nn_extra_nvp, _lib, ddinit
from . import defaults

batch_size = 4
rng = np.random.RandomState(42)
rng_test = np.random.RandomState(317070)
seq_len = defaults.seq_len
n_context = defaults.n_context

nonlinearity = nn_extra_nvp.elu
weight_norm = True

obs_shape = (seq_len, 32, 32, 1)      # the reference reads these from its ShapeNet iterator (m1:26-27)
label_shape = (seq_len, 2)
print('obs shape', obs_shape)
print('label shape', label_shape)

from ..synthetic_data import SyntheticShapenetIterator   # noqa: E402  (the reference reads ShapeNet renders, m1:17-25)
train_data_iter = SyntheticShapenetIterator(seq_len=seq_len, batch_size=batch_size, obs_shape=obs_shape[1:], rng=rng)
test_data_iter = SyntheticShapenetIterator(seq_len=seq_len, batch_size=batch_size, obs_shape=obs_shape[1:], rng=rng_test)

ndim = int(np.prod(obs_shape[1:]))
corr_init = np.ones((ndim,), dtype='float32') * 0.1

optimizer = 'rmsprop'
learning_rate = 0.001
lr_decay = 0.999995
scale_student_grad = 1.
student_grad_schedule = {0: 0., 500: 0.5, 1000: 1.}
max_iter = 200000
save_every = 5000

nvp_layers = []
nvp_dense_layers = []
student_layer = None
labels_layer = None     # {'kernel': [2, 32], 'bias': [32]}  (tf.layers.dense 'labels_layer', m1:66-67)


def _labels_embedding(y_label_bs, train=False):
    global labels_layer
    if labels_layer is None:
        g = torch.Generator().manual_seed(0)
        labels_layer = {'kernel': (torch.randn(y_label_bs.shape[1], 32, generator=g) * 0.5).to(y_label_bs.device),
                        'bias': torch.zeros(32, device=y_label_bs.device)}
    if train:
        return nn_extra_nvp.dense_train(y_label_bs.contiguous(), labels_layer['kernel'], labels_layer['bias'], nn_extra_nvp.LEAKY)
    return nn_extra_nvp._sgemm(y_label_bs.contiguous(), labels_layer['kernel'], 32, None, labels_layer['bias'], nn_extra_nvp.LEAKY)


def parameters():
    """Trainable variables (tf.trainable_variables() of the reference graph), in construction order."""
    out = [labels_layer['kernel'], labels_layer['bias']]
    for layer in list(nvp_layers) + list(nvp_dense_layers):
        if hasattr(layer, 'parameters'):
            out += layer.parameters()
    out += [p for _, p in student_layer.named_parameters()]
    return out


def set_trainable(on=True):
    for p in parameters():
        p.requires_grad_(on)


def _training():
    return torch.is_grad_enabled() and labels_layer is not None and labels_layer['kernel'].requires_grad


def build_model(x, y_label, init=False, sampling_mode=False, noise=None):
    """m1_shapenet.py:46-140.  Returns (log_probs, log_probs, log_probs) or x_samples in sampling mode.
    init=True: the data-dependent initialisation pass (every layer's output computed with the old g, b)."""
    global student_layer
    train = _training() and not sampling_mode and not init
    with (contextlib.nullcontext() if train else torch.no_grad()):
        if len(nvp_layers) == 0:
            build_nvp_model()
        if len(nvp_dense_layers) == 0:
            build_nvp_dense_model()
        if student_layer is None:
            student_layer = nn_extra_gauss.GaussianRecurrentLayer(shape=(ndim,), corr_init=corr_init, device=x.device)
        x_shape = nn_extra_nvp.int_shape(x)
        B, T = x_shape[0], x_shape[1]
        x_bs = x.reshape(B * T, x_shape[2], x_shape[3], x_shape[4]).contiguous()
        y_label_bs = _labels_embedding(y_label.reshape(B * T, -1).to(torch.float32), train)
        log_det_jac = torch.zeros(B * T, device=x.device)
        y, log_det_jac = nn_extra_nvp.dequantization_forward_and_jacobian(x_bs, log_det_jac)
        y, log_det_jac = nn_extra_nvp.logit_forward_and_jacobian(y, log_det_jac)
        z = None
        for layer in nvp_layers:
            if init and isinstance(layer, nn_extra_nvp.CouplingLayerConv):
                y, log_det_jac = ddinit.cond_conv_coupling_init(layer, y, log_det_jac, y_label_bs)
            else:
                y, log_det_jac, z = layer.forward_and_jacobian(y, log_det_jac, z, y_label=y_label_bs)
        z = nn_extra_nvp.channel_concat(z, y)
        for layer in nvp_dense_layers:
            if init:
                z, log_det_jac = ddinit.cond_dense_coupling_init(layer, z, log_det_jac, y_label_bs)
            else:
                z, log_det_jac, _ = layer.forward_and_jacobian(z, log_det_jac, None, y_label=y_label_bs)
        z_shape = nn_extra_nvp.int_shape(z)
        z_vec = z.reshape(B, T, -1)
        log_det_jac = log_det_jac.reshape(B, T)
        sl = student_layer
        D = z_vec.shape[2]
        if sampling_mode:
            mu, var, _ = sl.posterior_sequence(z_vec)                      # after 0..T observations
            z_samples = []
            for i in range(T):
                k = n_context if n_context > 0 else i                      # m1:86-99
                Bi = B if k > 0 else 1                                     # the prior has batch dim 1 (gauss:29)
                nz = noise[i] if noise is not None else sl.draw_noise(1, Bi, x.device)
                z_samples.append(sl.sample_from(mu[:Bi, k], var[:Bi, k], None, nz))
            z_samples = torch.cat(z_samples, 1).reshape(z_shape).contiguous()   # (sic) time-major samples, m1:119-121
            for layer in reversed(nvp_dense_layers):
                z_samples, _ = layer.backward(z_samples, None, y_label=y_label_bs)
            x_samples = None
            for layer in reversed(nvp_layers):
                x_samples, z_samples = layer.backward(x_samples, z_samples, y_label=y_label_bs)
            x_samples = torch.sigmoid(x_samples)                             # inverse logit only, m1:133
            return x_samples.reshape(1, B * T, x_shape[2], x_shape[3], x_shape[4])
        if n_context > 0 and train:
            # log p(z_i | z_0 .. z_{k-1}) for every i >= k is step k of the sequence (z_0 .. z_{k-1}, z_i): one batched call
            # of the differentiable scan kernel over B (T - k) sequences of length k + 1 (no update after the context)
            k = n_context
            ctx = z_vec[:, :k].unsqueeze(1).expand(B, T - k, k, D)
            seqs = torch.cat([ctx, z_vec[:, k:].unsqueeze(2)], 2).reshape(B * (T - k), k + 1, D).contiguous()
            latent, _ = sl.sequence_log_likelihoods(seqs, want_prior=False)
            log_probs = latent[:, k].reshape(B, T - k) + log_det_jac[:, k:]
        elif n_context > 0:
            xz = z_vec[:, :n_context] - sl.mu.view(1, 1, D)
            s1 = xz.sum(1).contiguous()
            s2 = (xz * xz).sum(1).contiguous()
            cols = [sl.log_likelihood_with_state(z_vec[:, i].contiguous(), s1, s2, n_context) + log_det_jac[:, i]
                    for i in range(n_context, T)]                          # no update after the context, m1:104-109
            log_probs = torch.stack(cols, 1)
        else:
            latent, _ = sl.sequence_log_likelihoods(z_vec, want_prior=False)
            log_probs = latent + log_det_jac
        return log_probs, log_probs, log_probs


def build_nvp_model():
    num_scales, num_filters, num_res_blocks = 3, 64, 6
    kw = dict(nonlinearity=nonlinearity, weight_norm=weight_norm, num_filters=num_filters, num_res_blocks=num_res_blocks)
    for scale in range(num_scales - 1):
        for i, m in enumerate(('checkerboard0', 'checkerboard1', 'checkerboard0')):
            nvp_layers.append(nn_extra_nvp.CouplingLayerConv(m, name='Checkerboard%d_%d' % (scale, i + 1), **kw))
        nvp_layers.append(nn_extra_nvp.SqueezingLayer(name='Squeeze%d' % scale))
        for i, m in enumerate(('channel0', 'channel1', 'channel0')):
            nvp_layers.append(nn_extra_nvp.CouplingLayerConv(m, name='Channel%d_%d' % (scale, i + 1), **kw))
        nvp_layers.append(nn_extra_nvp.FactorOutLayer(scale, name='FactorOut%d' % scale))
    scale = num_scales - 1
    for i, m in enumerate(('checkerboard0', 'checkerboard1', 'checkerboard0', 'checkerboard1')):
        nvp_layers.append(nn_extra_nvp.CouplingLayerConv(m, name='Checkerboard%d_%d' % (scale, i + 1), **kw))
    nvp_layers.append(nn_extra_nvp.FactorOutLayer(scale, name='FactorOut%d' % scale))


def build_nvp_dense_model():
    for i in range(6):
        mask = 'even' if i % 2 == 0 else 'odd'
        nvp_dense_layers.append(nn_extra_nvp.CouplingLayerDense(mask, name='%s_%s' % (mask, i), nonlinearity=nonlinearity,
                                                                n_units=256, weight_norm=weight_norm))


def named_tf_variables(prefix='model/', grad=False):
    G = (lambda t: t.grad) if grad else (lambda t: t)
    out = [(prefix + 'labels_layer/kernel', G(labels_layer['kernel'])), (prefix + 'labels_layer/bias', G(labels_layer['bias']))]
    for layer in list(nvp_layers) + list(nvp_dense_layers):
        if hasattr(layer, 'P') or hasattr(layer, 'V1'):
            out += layer.named_tf_variables(prefix, grad=grad)
    out += [(prefix + n, G(p)) for n, p in student_layer.named_parameters()]
    return out


def named_tf_gradients(prefix='model/'):
    """(TF variable name, gradient view) after loss.backward(); None for the variables the reference does not train."""
    return named_tf_variables(prefix, grad=True)


def load_tf_variables(params):
    with torch.no_grad():
        for name, view in named_tf_variables():
            src = params[name]
            src = src if torch.is_tensor(src) else torch.as_tensor(np.asarray(src))
            view.copy_(src.to(torch.float32).reshape(view.shape).to(view.device))


def loss(log_probs):
    return -log_probs.mean()
