"""GPU: the reference-facing entry points run end to end on the B200 path: train (data-dependent init at iteration 0,
RMSProp with TF-1.7 semantics, checkpoint + meta.pkl), resume, test_nll, test_few_shot_omniglot, test_samples."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(mod, *args, cwd):
    env = dict(os.environ, PYTHONPATH=ROOT)
    p = subprocess.run([sys.executable, '-m', mod] + list(args), cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return p.stdout + p.stderr


def test_train_resume_nll_fewshot_samples(tmp_path):
    cwd = str(tmp_path)
    out = run('bruno_b200.config_rnn.train', '--config_name', 'bn2_omniglot_tp', '--nr_gpu', '1', '--max_iter', '4', cwd=cwd)
    assert 'Saving model' in out
    meta_dirs = [d for d in os.listdir(os.path.join(cwd, 'metadata')) if d.startswith('bn2_omniglot_tp-')]
    assert len(meta_dirs) == 1
    md = os.path.join(cwd, 'metadata', meta_dirs[0])
    assert os.path.exists(os.path.join(md, 'params.pt'))
    meta = pickle.load(open(os.path.join(md, 'meta.pkl'), 'rb'))
    assert meta['iteration'] == 4 and np.isfinite(meta['lr'])
    # resume fast-forwards to the saved iteration and continues (train.py:161-175)
    out = run('bruno_b200.config_rnn.train', '--config_name', 'bn2_omniglot_tp', '--nr_gpu', '1', '--resume', '1',
              '--max_iter', '6', cwd=cwd)
    assert 'Last iteration 4' in out
    assert pickle.load(open(os.path.join(md, 'meta.pkl'), 'rb'))['iteration'] == 6
    out = run('bruno_b200.config_rnn.test_nll', '--config_name', 'bn2_omniglot_tp', '--n_batches', '3', cwd=cwd)
    assert 'Bits per dim' in out and 'Test loss under prior' in out
    line = [l for l in out.splitlines() if '|' in l and '(' in l][-1]
    vals = [float(v.split('(')[0]) for v in line.split('|')]
    assert all(np.isfinite(vals))
    out = run('bruno_b200.config_rnn.test_few_shot_omniglot', '--config_name', 'bn2_omniglot_tp', '--seq_len', '2',
              '--batch_size', '20', '--n_trials', '1', cwd=cwd)
    assert 'average accuracy over trials' in out
    out = run('bruno_b200.config_rnn.test_samples', '--config_name', 'bn2_omniglot_tp', '--seq_len', '4',
              '--n_sequences', '2', cwd=cwd)
    files = os.listdir(os.path.join(md, 'samples'))
    assert len(files) == 2
    s = np.load(os.path.join(md, 'samples', files[0]))
    assert s.shape == (4, 5, 28, 28, 1) and np.isfinite(s).all()
