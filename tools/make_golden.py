"""Generate tests/golden/*.npz.  Run in the build container only (imports /root/reference).

(1) tp_gp_reference.npz -- the setup of /root/reference/tests_gp_tp/test_layers.py:15-64 (batch 1, seq_len 100,
    p=1, RandomState(41), nu=100, var=1, corr=0.01, data cov 0.05 / var 0.1); expected predictive log-probs are
    produced by the REFERENCE's own numpy statements in tests_gp_tp/utils_stats.py (recursive Student-t
    predictive :51-96, joint multivariate-t ratio :138-153, joint Gaussian ratio :11-20) -- exactly the three
    checks of test_layers.py:117-178.
(2) flow_*.npz -- oracle (float64) outputs of reduced-size models on seeded inputs/weights, so the GPU box
    (which has no /root/reference) and later rounds can detect drift of the oracle itself.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')
from tests_gp_tp import utils_stats  # noqa: E402  (reference-authored numpy oracle)
from oracle import bruno_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)


def covariance(n, cov_u, var_u):
    K = np.ones((n, n)) * cov_u
    np.fill_diagonal(K, var_u)
    return K


def make_tp_gp():
    rng = np.random.RandomState(41)
    seq_len, p = 100, 1
    g_nu = np.ones((p,), dtype='float32') * 100
    g_var = np.ones((p,), dtype='float32') * 1.
    g_corr = np.ones((p,), dtype='float32') * 0.01
    g_cov = g_corr * g_var
    g_mu = np.zeros((p,), dtype='float32')
    K = covariance(seq_len, 0.05, 0.1)
    x1 = np.float32(rng.multivariate_normal(np.zeros(seq_len), K).reshape(seq_len, p))
    x32 = x1
    # NumPy >= 2 (NEP 50) no longer promotes float32 * python-float to float64, and the reference's
    # math.gamma(50.5) ~ 3e63 overflows float32.  Feed the SAME float32-rounded numbers as float64; the
    # reference functions themselves are untouched.
    x1 = x1.astype(np.float64)
    g_nu, g_var, g_corr, g_cov, g_mu = [a.astype(np.float64) for a in (g_nu, g_var, g_corr, g_cov, g_mu)]

    tp_rec = np.zeros(seq_len)
    tp_joint = np.full(seq_len, np.nan)
    gp_joint = np.full(seq_len, np.nan)
    tp_rec[0] = np.log(utils_stats.mvt_pdf(x1[0, 0:1], phi=g_mu[0:1], K=g_var[None, 0:1], nu=g_nu[0]))
    for i in range(1, seq_len):
        tp_rec[i] = np.log(utils_stats.predictive_mdim_student_prob(
            x_1n=x1[:i], x_next=x1[i], cov_u=g_cov, var_u=g_var, mu=g_mu, nu=g_nu, recursive=True))
    for j in range(2, seq_len + 1):
        x_2n = x1[:j, 0:1].T
        x_1n = x1[:j - 1, 0:1].T
        pj2 = utils_stats.mvt_pdf(x_2n, phi=np.zeros_like(x_2n), K=covariance(j, g_cov[0], g_var[0]), nu=g_nu[0])
        pj1 = utils_stats.mvt_pdf(x_1n, phi=np.zeros_like(x_1n), K=covariance(j - 1, g_cov[0], g_var[0]), nu=g_nu[0])
        tp_joint[j - 1] = np.log(float(np.squeeze(pj2)) / float(np.squeeze(pj1)))
        gp_joint[j - 1] = np.log(utils_stats.gaussian_predictive_theoretical_prob_slow(
            x1[:j - 1, 0], x1[:j, 0], g_mu[0], g_cov[0], g_var[0]))
    np.savez(os.path.join(OUT, 'tp_gp_reference.npz'), x=x32[None], nu=g_nu, var=g_var, corr=g_corr, mu=g_mu,
             tp_recursive=tp_rec, tp_joint=tp_joint, gp_joint=gp_joint)
    print('tp_gp_reference.npz', tp_rec[:3], tp_joint[1:3], gp_joint[1:3])


def make_flow(config, B, T, res_blocks, filters, units, seed, tag):
    spec = O.spec_for(config, num_filters=filters, num_res_blocks=res_blocks, n_units=units)
    P = O.init_params(spec, seed=seed, dtype=torch.float64, perturb_process=True)
    x = O.synthetic_batch(spec, B, T, seed=seed + 1, kind='strokes' if 'omniglot' in config else 'bytes')
    O.data_dependent_init(spec, P, x)
    with torch.no_grad():
        lp, lat, pri, z = O.build_model(spec, P, x)
    l, grads = O.loss_and_grads(spec, P, x)
    keys = sorted(grads.keys())
    gnorm = np.array([grads[k].norm().item() for k in keys])
    np.savez(os.path.join(OUT, 'flow_%s.npz' % tag), config=config, B=B, T=T, res_blocks=res_blocks,
             filters=filters, units=units, seed=seed, log_probs=lp.numpy(), latent=lat.numpy(), prior=pri.numpy(),
             z_sum=z.sum(-1).numpy(), loss=l.item(), grad_keys=np.array(keys), grad_norms=gnorm)
    print('flow_%s.npz' % tag, 'loss', l.item())


if __name__ == '__main__':
    make_tp_gp()
    make_flow('en2_noconv_mnist_tp', 4, 5, 8, 64, 64, 3, 'en2_small')
    make_flow('bn2_omniglot_tp', 2, 3, 1, 64, 64, 5, 'bn2_small')
    make_flow('bn2_omniglot_gp', 2, 3, 1, 64, 64, 7, 'bn2gp_small')
    make_flow('an2_cifar_tp', 1, 2, 1, 64, 64, 9, 'an2_small')
