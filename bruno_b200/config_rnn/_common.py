"""What the four unconditional config modules share: the two-scale conv stack of build_nvp_model
(config_rnn/bn2_omniglot_tp.py:171-210, an2_cifar_tp.py:156-195), the dense stack of build_nvp_dense_model
(bn2:213-221) and build_model itself (bruno_b200/model.py)."""
import torch

from .. import nn_extra_nvp
from ..model import BrunoModel


def build_nvp_model(nvp_layers, nonlinearity, weight_norm, num_scales=2):
    for scale in range(num_scales - 1):
        for i, m in enumerate(('checkerboard0', 'checkerboard1', 'checkerboard0')):
            nvp_layers.append(nn_extra_nvp.CouplingLayerConv(m, name='Checkerboard%d_%d' % (scale, i + 1),
                                                             nonlinearity=nonlinearity, weight_norm=weight_norm))
        nvp_layers.append(nn_extra_nvp.SqueezingLayer(name='Squeeze%d' % scale))
        for i, m in enumerate(('channel0', 'channel1', 'channel0')):
            nvp_layers.append(nn_extra_nvp.CouplingLayerConv(m, name='Channel%d_%d' % (scale, i + 1),
                                                             nonlinearity=nonlinearity, weight_norm=weight_norm))
        nvp_layers.append(nn_extra_nvp.FactorOutLayer(scale, name='FactorOut%d' % scale))
    scale = num_scales - 1
    for i, m in enumerate(('checkerboard0', 'checkerboard1', 'checkerboard0', 'checkerboard1')):
        nvp_layers.append(nn_extra_nvp.CouplingLayerConv(m, name='Checkerboard%d_%d' % (scale, i + 1),
                                                         nonlinearity=nonlinearity, weight_norm=weight_norm))
    nvp_layers.append(nn_extra_nvp.FactorOutLayer(scale, name='FactorOut%d' % scale))


def build_nvp_dense_model(nvp_dense_layers, nonlinearity, weight_norm, n_units=512):
    for i in range(6):
        mask = 'even' if i % 2 == 0 else 'odd'
        name = '%s_%s' % (mask, i)
        nvp_dense_layers.append(nn_extra_nvp.CouplingLayerDense(mask, name=name, nonlinearity=nonlinearity,
                                                                n_units=n_units, weight_norm=weight_norm))


def run_build_model(cfg, x, init, sampling_mode, noise=None, want_prior=True):
    """`cfg` is the config module; mirrors the body of its build_model in the reference."""
    if cfg.model is None:
        if cfg.has_conv and len(cfg.nvp_layers) == 0:
            cfg.build_nvp_model()
        if len(cfg.nvp_dense_layers) == 0:
            cfg.build_nvp_dense_model()
        if cfg.student_layer is None:
            cfg.student_layer = cfg.make_process_layer(x.device)
        cfg.model = BrunoModel(cfg.nvp_layers, cfg.nvp_dense_layers, cfg.student_layer, n_samples=cfg.n_samples)
    mask_dim = None
    from . import defaults
    if getattr(defaults, 'mask_dims', False) and getattr(cfg, 'supports_mask_dims', False):
        # en2_noconv_mnist_tp.py:91-95
        mask_dim = (cfg.student_layer.corr.detach() > defaults.eps_corr).to(torch.float32)
    return cfg.model.build_model(x, init=init, sampling_mode=sampling_mode, mask_dim=mask_dim, noise=noise,
                                 want_prior=want_prior)
