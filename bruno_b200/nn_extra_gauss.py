"""Recurrent Gaussian process per latent dimension -- host mirror of the reference's nn_extra_gauss.py
(GaussianRecurrentLayer :18-116); the arithmetic lives in csrc/process.cu behind bruno_process_* (C ABI)."""
from . import _lib
from .process import RecurrentProcessLayer, Gaussian, inv_softplus, inv_sigmoid  # noqa: F401


class GaussianRecurrentLayer(RecurrentProcessLayer):
    kind = _lib.GP

    def __init__(self, shape, mu_init=0., var_init=1., corr_init=0.1, device='cuda'):
        self._init_common(shape, mu_init, var_init, corr_init, device)

    @property
    def prior(self):                                                                # gauss:42-45
        return Gaussian(self.mu, self.var)

    @property
    def variables(self):                                                            # gauss:56-58
        return self.mu, self.var_vbl, self.corr_vbl

    def named_parameters(self, prefix=''):
        # the reference creates these under variable_scope("gaussian") (gauss:28)
        return [(prefix + 'gaussian/prior_var', self.var_vbl), (prefix + 'gaussian/prior_corr', self.corr_vbl)]
