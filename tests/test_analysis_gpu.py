"""GPU: the analysis entry points that consume outputs 1-3 of build_model (reference:
config_rnn/test_anomaly.py, test_corr.py, test_latent_dist.py, test_probs_evolution.py) run end to end on a checkpoint
written by the train driver."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_analysis_entry_points(tmp_path):
    cwd = str(tmp_path)
    env = dict(os.environ, PYTHONPATH=ROOT)

    def run(mod, *args):
        p = subprocess.run([sys.executable, '-m', mod] + list(args), cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
        return p.stdout + p.stderr
    name = 'en2_noconv_mnist_tp'
    run('bruno_b200.config_rnn.train', '--config_name', name, '--nr_gpu', '1', '--max_iter', '3')
    md = os.path.join(cwd, 'metadata', [d for d in os.listdir(os.path.join(cwd, 'metadata')) if d.startswith(name + '-')][0])

    out = run('bruno_b200.config_rnn.test_anomaly', '--config_name', name, '--seq_len', '12', '--n_sequences', '2')
    assert '-------------' in out
    s = np.load(os.path.join(md, 'anomaly', 'anomaly_0_0.npy'))
    assert s.shape == (12,) and np.isfinite(s).all()

    out = run('bruno_b200.config_rnn.test_corr', '--config_name', name)
    assert '******* corr - nu ********' in out and 'VAR' in out
    eps = json.load(open(os.path.join(md, 'misc_plots', 'eps_plot.json')))
    assert eps['n_dims'][0] == 784                          # corr_init = 0.1 > 0 in every dimension

    run('bruno_b200.config_rnn.test_latent_dist', '--config_name', name, '--seq_len', '2', '--set', 'test', '--batch_size', '64')
    h = np.load(os.path.join(md, 'hists_test', 'hists.npz'))
    assert h['hists'].shape == (784, 100) and np.isfinite(h['pdfs']).all()

    out = run('bruno_b200.config_rnn.test_probs_evolution', '--config_name', name, '--seq_len', '6', '--n_batches', '3')
    assert 'avg LL under prior' in out
    sc = np.load(os.path.join(md, 'misc_plots', 'scores_plot_len6_test_class1_img0_black0.npy'))
    assert sc.shape == (6,) and np.isfinite(sc).all()
    assert abs(sc[0]) < 1e-2 * max(1.0, np.abs(sc).max())   # before any observation the predictive IS the prior
