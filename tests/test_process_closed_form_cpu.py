"""CPU: pins the algebra the t-process scan kernels rely on against the oracle's literal recursion.

csrc/process.cu does not run the reference's update (nn_extra_student.py:85-106) step by step: a table kernel evaluates
everything that depends on (step, dim) only, and the ring scan carries q_n = 1 + beta_n / (nu0 - 2) recursively with
q_{n+1} = q_n + r_n^2 = u_n (so lg2(q_n) is the previous step's lg2(u_{n-1})).  This file restates exactly those formulas
in float64 numpy and checks them against oracle.StudentOracle / GaussOracle (the literal recursion) and against the
reference's golden vector; the GPU tests then only have to show that the CUDA arithmetic follows these formulas."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import bruno_oracle as O


def _tp_scan_closed_form(z, var, nu0, corr, mu):
    """z [B,T,D] float64; var, nu0, corr, mu [D].  Mirrors process_table_kernel + process_fwd_ring_kernel (TP)."""
    B, T, D = z.shape
    c = corr * var
    S1 = np.zeros((B, D))
    q = np.ones((B, D))
    lq = np.zeros((B, D))
    lp = np.zeros((B, T))
    for n in range(T):
        dd = c / (var + c * (n - 1.0))                         # student:93
        k = var - dd * c * n                                   # closed form of student:101
        c0 = k * (nu0 - 2.0)
        w = 1.0 / np.sqrt(c0)
        A = (nu0 + n + 1.0) / 2.0
        C = np.array([math.lgamma(a) - math.lgamma(b) for a, b in zip(A, (nu0 + n) / 2.0)]) - 0.5 * np.log(c0) - 0.5 * math.log(math.pi)
        xz = z[:, n, :] - mu
        r = w * xz - (dd * w) * S1
        u = r * r + q
        lu = np.log(u)
        lp[:, n] = (C + (A - 0.5) * lq - A * lu).sum(1)
        S1 = S1 + xz
        q, lq = u, lu                                          # q_{n+1} = u_n
    return lp


def _tp_q_closed_form(z, var, nu0, corr, mu):
    """q_n = 1 + e S2 + f_n S1^2 (the register-staged kernel and the backward) for every step: [B,T,D]."""
    B, T, D = z.shape
    c = corr * var
    e = 1.0 / ((var - c) * (nu0 - 2.0))
    xz = z - mu
    S1 = np.concatenate([np.zeros((B, 1, D)), np.cumsum(xz, 1)[:, :-1]], 1)
    S2 = np.concatenate([np.zeros((B, 1, D)), np.cumsum(xz * xz, 1)[:, :-1]], 1)
    n = np.arange(T).reshape(1, T, 1)
    f = -c / ((var - c) * (c * (n - 1.0) + var) * (nu0 - 2.0))  # student:98
    return 1.0 + e * S2 + f * S1 * S1


def _gp_scan_closed_form(z, var, corr, mu):
    B, T, D = z.shape
    c = corr * var
    S1 = np.zeros((B, D))
    lp = np.zeros((B, T))
    for n in range(T):
        dd = c / (var + c * (n - 1.0))
        k = var - dd * c * n
        xz = z[:, n, :] - mu
        r = xz - dd * S1
        lp[:, n] = (-0.5 * np.log(2.0 * math.pi * k) - 0.5 / k * r * r).sum(1)
        S1 = S1 + xz
    return lp


def _setup(kind, D, seed, mu):
    spec = O.Spec('t', (1, 1, D), 1, [], [], 1, kind)
    P = O.init_params(spec, seed=seed, dtype=torch.float64, perturb_process=True)
    m = torch.full((1, D), float(mu), dtype=torch.float64)
    proc = O.StudentOracle(P, mu=m) if kind == 'tp' else O.GaussOracle(P, mu=m)
    return proc


@pytest.mark.parametrize('B,T,D,mu,scale', [(3, 20, 16, 0.0, 1.5), (2, 7, 5, 0.4, 1.0), (2, 12, 8, 0.0, 3e3)])
def test_tp_recursive_q_scan_equals_literal_recursion(B, T, D, mu, scale):
    proc = _setup('tp', D, seed=B + T, mu=mu)
    g = torch.Generator().manual_seed(D)
    z = torch.randn(B, T, D, generator=g, dtype=torch.float64) * scale + 0.3 * scale * torch.randn(B, 1, D, generator=g, dtype=torch.float64)
    lp_o, _ = O.process_log_probs(proc, z)
    var, nu0, corr = (t.numpy()[0] for t in (proc.var, proc.nu, proc.corr))
    lp = _tp_scan_closed_form(z.numpy(), var, nu0, corr, np.full(D, mu))
    assert np.allclose(lp, lp_o.numpy(), rtol=1e-10, atol=1e-8)
    # the identity itself: the recursively carried q equals the closed form of the prefix sums at every step
    qc = _tp_q_closed_form(z.numpy(), var, nu0, corr, np.full(D, mu))
    c = corr * var
    S1 = np.zeros((B, D))
    q = np.ones((B, D))
    for n in range(T):
        assert np.allclose(q, qc[:, n, :], rtol=1e-9, atol=1e-9)
        dd = c / (var + c * (n - 1.0))
        w = 1.0 / np.sqrt((var - dd * c * n) * (nu0 - 2.0))
        xz = z.numpy()[:, n, :] - mu
        r = w * (xz - dd * S1)
        q = q + r * r
        S1 = S1 + xz


@pytest.mark.parametrize('B,T,D,mu', [(3, 20, 16, 0.0), (2, 7, 5, 0.4)])
def test_gp_closed_form_scan_equals_literal_recursion(B, T, D, mu):
    proc = _setup('gp', D, seed=B + T, mu=mu)
    z = torch.randn(B, T, D, generator=torch.Generator().manual_seed(D), dtype=torch.float64) * 1.5
    lp_o, _ = O.process_log_probs(proc, z)
    var, corr = (t.numpy()[0] for t in (proc.var, proc.corr))
    lp = _gp_scan_closed_form(z.numpy(), var, corr, np.full(D, mu))
    assert np.allclose(lp, lp_o.numpy(), rtol=1e-10, atol=1e-8)


def test_tp_recursive_q_scan_matches_reference_golden_vector(golden_dir):
    """Same formulas against the reference's own numbers (tests_gp_tp/test_layers.py:127/:141 via tools/make_golden.py)."""
    g = np.load(os.path.join(golden_dir, 'tp_gp_reference.npz'))
    x = g['x'].astype(np.float64)                             # [1,100,1]
    lp = _tp_scan_closed_form(x, g['var'].astype(np.float64), g['nu'].astype(np.float64), g['corr'].astype(np.float64), np.zeros(1))
    assert np.allclose(lp[0], g['tp_recursive'], atol=1e-8, rtol=1e-8)
    lg = _gp_scan_closed_form(x, g['var'].astype(np.float64), g['corr'].astype(np.float64), np.zeros(1))
    assert np.allclose(lg[0][1:], g['gp_joint'][1:], atol=1e-3, rtol=1e-3)


def _tp_backward_dz(z, dlp, var, nu0, corr, mu):
    """dz of sum(lp * dlp) with the two-pass scheme of process_bwd_ring_kernel (TP): totals forward, then the sequence in
    reverse, recovering the prefix sums by subtraction and carrying the suffix sums G1 / G2 of dL/dS1, dL/dS2."""
    B, T, D = z.shape
    c = corr * var
    e = 1.0 / ((var - c) * (nu0 - 2.0))
    xz_all = z - mu
    S1 = xz_all.sum(1)
    S2 = (xz_all * xz_all).sum(1)
    G1 = np.zeros((B, D))
    G2 = np.zeros((B, D))
    dz = np.zeros_like(z)
    for n in range(T - 1, -1, -1):
        dd = c / (var + c * (n - 1.0))
        w = 1.0 / np.sqrt((var - dd * c * n) * (nu0 - 2.0))
        ddw = dd * w
        f = -c / ((var - c) * (c * (n - 1.0) + var) * (nu0 - 2.0))
        A = (nu0 + n + 1.0) / 2.0
        xz = xz_all[:, n, :]
        dl = dlp[:, n:n + 1]
        S1 = S1 - xz                                           # prefix sums BEFORE observation n
        S2 = S2 - xz * xz
        r = w * xz - ddw * S1
        q = 1.0 + e * S2 + f * S1 * S1
        u = r * r + q
        gu = -A / u * dl                                       # d/du of -A ln u
        gq = (A - 0.5) / q * dl + gu                           # q enters directly and through u = r^2 + q
        gr = 2.0 * r * gu
        dz[:, n, :] = gr * w + G1 + 2.0 * xz * G2
        G1 = G1 + (gq * f * 2.0 * S1 - gr * ddw)
        G2 = G2 + gq * e
    return dz


@pytest.mark.parametrize('B,T,D,mu', [(3, 9, 6, 0.0), (2, 5, 4, 0.3)])
def test_tp_two_pass_backward_equals_oracle_autograd(B, T, D, mu):
    proc = _setup('tp', D, seed=T, mu=mu)
    g = torch.Generator().manual_seed(B + D)
    z = (torch.randn(B, T, D, generator=g, dtype=torch.float64) * 1.3).requires_grad_(True)
    dlp = torch.randn(B, T, generator=g, dtype=torch.float64)
    lp_o, _ = O.process_log_probs(proc, z)
    (dz_o,) = torch.autograd.grad((lp_o * dlp).sum(), [z])
    var, nu0, corr = (t.numpy()[0] for t in (proc.var, proc.nu, proc.corr))
    dz = _tp_backward_dz(z.detach().numpy(), dlp.numpy(), var, nu0, corr, np.full(D, mu))
    assert np.allclose(dz, dz_o.numpy(), rtol=1e-8, atol=1e-9)
