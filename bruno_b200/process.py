"""Host side of the exchangeable-process kernels: autograd Function + the shared base of the two recurrent layers.

Mirrors the surface of the reference's StudentRecurrentLayer (nn_extra_student.py:19-149) and
GaussianRecurrentLayer (nn_extra_gauss.py:18-116); all arithmetic runs in libbruno_sm100.so.
"""
import collections

import numpy as np
import torch

from . import _lib

Student = collections.namedtuple('Student', ['mu', 'var', 'nu'])          # nn_extra_student.py:6
Gaussian = collections.namedtuple('Gaussian', ['mu', 'var'])              # nn_extra_gauss.py:5
State = collections.namedtuple('State', ['num_observations', 'beta', 'x_sum', 'k'])


def inv_softplus(x):                                                      # nn_extra_student.py:10
    return np.log(1 - np.exp(-x)) + x


def inv_sigmoid(x):                                                       # nn_extra_student.py:14
    return np.log(x) - np.log(1. - x)


def _empty(n, like):
    return torch.empty(max(int(n), 1), dtype=torch.float32, device=like.device)


def build_table(kind, var_vbl, nu_vbl, corr_vbl, mu, mask_dim, min_nu, n0, T, with_grad):
    D = var_vbl.numel()
    table = _empty(_lib.query('bruno_process_table_floats', kind, T, D, int(with_grad)), var_vbl)
    _lib.call('bruno_process_table', kind, var_vbl, nu_vbl, corr_vbl, mu, mask_dim, float(min_nu), int(n0), int(T), D,
              int(with_grad), table)
    return table


class ProcessSequenceLogLik(torch.autograd.Function):
    """(z[B,T,D], raw variables) -> (log_probs[B,T], log_probs_under_prior[B,T]); hand-written CUDA backward."""

    @staticmethod
    def forward(ctx, z, var_vbl, nu_vbl, corr_vbl, mu, mask_dim, kind, min_nu, want_prior, track):
        z = z.contiguous()
        B, T, D = z.shape
        need_grad = track and any(ctx.needs_input_grad)
        table = build_table(kind, var_vbl, nu_vbl, corr_vbl, mu, mask_dim, min_nu, 0, T, need_grad)
        lp = torch.empty(B, T, dtype=torch.float32, device=z.device)
        lpp = torch.empty(B, T, dtype=torch.float32, device=z.device) if want_prior else None
        scratch = _empty(_lib.query('bruno_process_scratch_floats', B, T, D), z)
        _lib.call('bruno_process_fwd', kind, z, mu, table, None, None, lp, lpp, scratch, B, T, D)
        ctx.kind, ctx.has_nu, ctx.want_prior = kind, nu_vbl is not None, want_prior
        ctx.save_for_backward(z, mu, table)
        if want_prior:
            return lp, lpp
        dummy = torch.zeros(0, device=z.device)
        ctx.mark_non_differentiable(dummy)
        return lp, dummy

    @staticmethod
    def backward(ctx, dlp, dlpp):
        z, mu, table = ctx.saved_tensors
        B, T, D = z.shape
        dlp = dlp.contiguous()
        dlpp = dlpp.contiguous() if (ctx.want_prior and dlpp is not None) else None
        dz = torch.empty_like(z)
        dvar = torch.empty(1, D, dtype=torch.float32, device=z.device)
        dnu = torch.empty(1, D, dtype=torch.float32, device=z.device) if ctx.has_nu else None
        dcorr = torch.empty(1, D, dtype=torch.float32, device=z.device)
        scratch = _empty(_lib.query('bruno_process_bwd_scratch_floats', B, T, D), z)
        _lib.call('bruno_process_bwd', ctx.kind, z, mu, table, dlp, dlpp, dz, dvar, dnu, dcorr, scratch, B, T, D)
        return dz, dvar, dnu, dcorr, None, None, None, None, None, None


class RecurrentProcessLayer(object):
    """Shared implementation.  Public attributes follow the reference: mu, var_vbl, (nu_vbl), corr_vbl, var, nu,
    corr, cov, prior, current_distribution, variables, shape; methods reset / update_distribution /
    get_log_likelihood / get_log_likelihood_under_prior / sample."""

    kind = None
    min_nu = 2.

    def _init_common(self, shape, mu_init, var_init, corr_init, device):
        self.seed_rng = np.random.RandomState(42)                          # student:26 / gauss:24
        self._shape = tuple(shape)
        D = int(np.prod(self._shape))
        self._D = D
        dev = torch.device(device)
        self.mu = (torch.ones(1, D, device=dev) * torch.as_tensor(mu_init, dtype=torch.float32, device=dev)).contiguous()
        var0 = np.broadcast_to(np.asarray(inv_softplus(np.sqrt(var_init)), dtype=np.float32), (1, D)).copy()
        corr0 = np.broadcast_to(np.asarray(inv_sigmoid(np.asarray(corr_init, dtype=np.float64)), dtype=np.float32), (1, D)).copy()
        self.var_vbl = torch.from_numpy(var0).to(dev).requires_grad_(True)     # "prior_var"
        self.corr_vbl = torch.from_numpy(corr0).to(dev).requires_grad_(True)   # "prior_corr"
        self.nu_vbl = None
        self.reset()

    # ---- derived quantities (student:38,46,60-61) ----------------------------------------------
    @property
    def var(self):
        return torch.nn.functional.softplus(self.var_vbl) ** 2

    @property
    def corr(self):
        return torch.sigmoid(self.corr_vbl)

    @property
    def cov(self):
        return self.corr * self.var

    @property
    def shape(self):
        return self._shape

    # ---- the fused path used by the config builders ---------------------------------------------
    def sequence_log_likelihoods(self, z_vec, mask_dim=None, want_prior=True):
        """The unrolled loop of bn2_omniglot_tp.py:103-124 (non-sampling branch) in one kernel launch.
        Returns (latent_log_probs [B,T], latent_log_probs_prior [B,T] or None)."""
        if mask_dim is not None:
            mask_dim = mask_dim.reshape(-1).to(torch.float32).contiguous()
        lp, lpp = ProcessSequenceLogLik.apply(z_vec, self.var_vbl, self.nu_vbl, self.corr_vbl, self.mu, mask_dim,
                                              self.kind, self.min_nu, bool(want_prior), torch.is_grad_enabled())
        return lp, (lpp if want_prior else None)

    # ---- the reference's step-by-step protocol (stateful, no autograd) --------------------------------
    def reset(self):                                                       # student:81-83
        self._s1 = self._s2 = None
        self._n = 0
        self._cur = None
        self._obs_list = []

    def _ensure_state(self, B, device):
        if self._s1 is None:
            self._s1 = torch.zeros(B, self._D, device=device)
            self._s2 = torch.zeros(B, self._D, device=device)

    def _step(self, observation, mask_dim, advance):
        with torch.no_grad():
            obs = observation.detach().to(torch.float32).contiguous()
            B, D = obs.shape
            self._ensure_state(B, obs.device)
            if mask_dim is not None:
                mask_dim = mask_dim.reshape(-1).to(torch.float32).contiguous()
            table = build_table(self.kind, self.var_vbl.detach(), None if self.nu_vbl is None else self.nu_vbl.detach(),
                                self.corr_vbl.detach(), self.mu, mask_dim, self.min_nu, self._n, 1, False)
            lp = torch.empty(B, 1, device=obs.device)
            lpp = torch.empty(B, 1, device=obs.device)
            s1, s2 = (self._s1, self._s2) if advance else (self._s1.clone(), self._s2.clone())
            scratch = _empty(_lib.query('bruno_process_scratch_floats', B, 1, D), obs)
            _lib.call('bruno_process_fwd', self.kind, obs.view(B, 1, D), self.mu, table, s1, s2, lp, lpp, scratch, B, 1, D)
            if advance:
                self._n += 1
                self._cur = None
            return lp[:, 0], lpp[:, 0]

    def update_distribution(self, observation):                           # student:85-106 / gauss:75-87
        self._step(observation, None, True)
        self._obs_list.append(observation.detach().to(torch.float32))

    def get_log_likelihood(self, observation, mask_dim=None):             # student:108-118 / gauss:89-96
        return self._step(observation, mask_dim, False)[0]

    def get_log_likelihood_under_prior(self, observation, mask_dim=None):  # student:120-130 / gauss:98-105
        return self._step(observation, mask_dim, False)[1]

    def log_likelihood_with_state(self, observation, s1, s2, n_obs, mask_dim=None):
        """get_log_likelihood(observation [n,D]) under the distribution reached after `n_obs` observations whose
        prefix sums are s1, s2 [n,D] (left untouched)."""
        with torch.no_grad():
            obs = observation.detach().to(torch.float32).contiguous()
            n, D = obs.shape
            table = build_table(self.kind, self.var_vbl.detach(), None if self.nu_vbl is None else self.nu_vbl.detach(),
                                self.corr_vbl.detach(), self.mu, mask_dim, self.min_nu, int(n_obs), 1, False)
            lp = torch.empty(n, 1, device=obs.device)
            scratch = _empty(_lib.query('bruno_process_scratch_floats', n, 1, D), obs)
            _lib.call('bruno_process_fwd', self.kind, obs.view(n, 1, D), self.mu, table, s1.clone(), s2.clone(), lp, None,
                      scratch, n, 1, D)
            return lp[:, 0]

    def posterior_sequence(self, z_vec):
        """current_distribution after 0..T observations: mu, var [B,T+1,D] and nu [T+1,D] (None for GP)."""
        with torch.no_grad():
            z = z_vec.detach().to(torch.float32).contiguous()
            B, T, D = z.shape
            mu = torch.empty(B, T + 1, D, device=z.device)
            var = torch.empty(B, T + 1, D, device=z.device)
            nu = torch.empty(T + 1, D, device=z.device) if self.kind == _lib.TP else None
            _lib.call('bruno_process_posterior', self.kind, z, self.var_vbl.detach(),
                      None if self.nu_vbl is None else self.nu_vbl.detach(), self.corr_vbl.detach(), self.mu,
                      float(self.min_nu), mu, var, nu, B, T, D)
            return mu, var, nu

    @property
    def current_distribution(self):
        """Student(mu, var, nu) / Gaussian(mu, var) after the observations fed through update_distribution."""
        if getattr(self, '_cur', None) is None:
            if self._n == 0:
                mu, var = self.mu, self.var.detach()
                nu = None if self.kind == _lib.GP else self.nu.detach()
            else:
                m, v, n = self.posterior_sequence(torch.stack(self._obs_list, 1))
                mu, var = m[:, -1, :], v[:, -1, :]
                nu = None if n is None else n[-1:, :]
            self._cur = Student(mu, var, nu) if self.kind == _lib.TP else Gaussian(mu, var)
        return self._cur

    def sample_from(self, mu, var, nu, noise):
        """student:132-149 / gauss:107-116 with the random draws supplied: TP noise [2,n,B,D] uniforms, GP [n,B,D]
        normals; mu/var [B,D] (B may be 1), nu [D]."""
        with torch.no_grad():
            noise = noise.to(torch.float32).contiguous()
            n, B, D = noise.shape[-3:]
            mu = mu.reshape(-1, D).to(torch.float32)
            var = var.reshape(-1, D).to(torch.float32)
            if mu.shape[0] != B:
                mu, var = mu.expand(B, D), var.expand(B, D)
            mu, var = mu.contiguous(), var.contiguous()
            out = torch.empty(n, B, D, device=noise.device)
            _lib.call('bruno_process_sample', self.kind, noise, mu, var,
                      None if nu is None else nu.reshape(-1).to(torch.float32).contiguous(), out, n, B, D, D)
            return out

    def draw_noise(self, nr_samples, B, device):
        """Host-side replacement of tf.random_uniform / tf.random_normal (student:135-139, gauss:110-113): seeded
        from the layer's RandomState like the reference; the stream itself cannot match TF's Philox."""
        seed = int(self.seed_rng.randint(317070))
        g = torch.Generator(device='cpu').manual_seed(seed)
        if self.kind == _lib.TP:
            return torch.rand(2, nr_samples, B, self._D, generator=g).to(device)
        return torch.randn(nr_samples, B, self._D, generator=g).to(device)

    def sample(self, nr_samples=1, noise=None):
        cur = self.current_distribution
        B = cur.mu.shape[0]
        if noise is None:
            noise = self.draw_noise(nr_samples, B, self.mu.device)
        nu = cur.nu if self.kind == _lib.TP else None
        return self.sample_from(cur.mu, cur.var, nu, noise)
