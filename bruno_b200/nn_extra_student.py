"""Recurrent Student-t process per latent dimension -- host mirror of the reference's nn_extra_student.py
(StudentRecurrentLayer :19-149); the arithmetic lives in csrc/process.cu behind bruno_process_* (C ABI)."""
import numpy as np
import torch

from . import _lib
from .process import RecurrentProcessLayer, Student, State, inv_softplus, inv_sigmoid  # noqa: F401


class StudentRecurrentLayer(RecurrentProcessLayer):
    kind = _lib.TP

    def __init__(self, shape, nu_init=1000, mu_init=0., var_init=1., corr_init=0.1, min_nu=2., device='cuda'):
        self.min_nu = float(min_nu)
        self._init_common(shape, mu_init, var_init, corr_init, device)
        nu0 = np.broadcast_to(np.log(np.asarray(nu_init, dtype=np.float64) - min_nu).astype(np.float32), (1, self._D)).copy()
        self.nu_vbl = torch.from_numpy(nu0).to(self.mu.device).requires_grad_(True)   # "prior_nu", student:40-45

    @property
    def nu(self):                                                                   # student:46
        return torch.exp(self.nu_vbl) + self.min_nu

    @property
    def prior(self):                                                                # student:48-52
        return Student(self.mu, self.var, self.nu)

    @property
    def variables(self):                                                            # student:66-68
        return self.mu, self.var_vbl, self.nu_vbl, self.corr_vbl

    def named_parameters(self, prefix=''):
        return [(prefix + 'prior_var', self.var_vbl), (prefix + 'prior_nu', self.nu_vbl),
                (prefix + 'prior_corr', self.corr_vbl)]
