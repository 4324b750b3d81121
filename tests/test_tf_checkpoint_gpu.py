"""GPU: a checkpoint in the TensorFlow-1.x format the REFERENCE writes (params.ckpt.index / .data-*, config_rnn/train.py:232)
is restored by the entry points exactly like the package's own params.pt."""
import importlib
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_entry_point_restores_a_tensorflow_checkpoint(tmp_path):
    cwd = str(tmp_path)
    env = dict(os.environ, PYTHONPATH=ROOT)

    def run(mod, *args):
        p = subprocess.run([sys.executable, '-m', mod] + list(args), cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
        return p.stdout + p.stderr
    name = 'en2_noconv_mnist_tp'
    run('bruno_b200.config_rnn.train', '--config_name', name, '--nr_gpu', '1', '--max_iter', '3')
    md = os.path.join(cwd, 'metadata', [d for d in os.listdir(os.path.join(cwd, 'metadata')) if d.startswith(name + '-')][0])
    out_pt = run('bruno_b200.config_rnn.test_nll', '--config_name', name, '--n_batches', '2')
    # re-home the variables into a TF V2 checkpoint and drop params.pt: what a directory written by the reference holds
    from bruno_b200 import utils, tf_checkpoint
    cfg = importlib.reload(importlib.import_module('bruno_b200.config_rnn.' + name))
    utils.load_checkpoint_into(cfg, md, torch.device('cuda', 0))
    utils.export_tf_checkpoint(cfg, md)
    os.remove(os.path.join(md, 'params.pt'))
    assert sorted(os.listdir(md))[:3] == ['checkpoint', 'meta.pkl', 'params.ckpt.data-00000-of-00001']
    got = tf_checkpoint.read_checkpoint(os.path.join(md, 'params.ckpt'))
    assert 'model/prior_nu' in got and got['model/even_0/function_s_t_dense_wn/d1/V'].shape == (784, 512)
    out_tf = run('bruno_b200.config_rnn.test_nll', '--config_name', name, '--n_batches', '2')
    assert 'restoring parameters from' in out_tf
    pick = lambda o: [l for l in o.splitlines() if 'Bits per dim' in l or ('|' in l and '(' in l)]
    assert pick(out_pt) == pick(out_tf)                     # identical numbers from either checkpoint format
