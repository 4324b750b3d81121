"""CPU, world_size 2, gloo: the data-parallel contract of config_rnn/train.py:83-111 that bruno_b200/trainer.py relies on --
each rank takes batch/nr_gpu sequences (np.split), per-rank loss = mean over its shard, gradients are summed over ranks
and divided by nr_gpu, student-parameter gradients are scaled -- reproduces the full-batch gradient.  The compute is the
oracle (this is a test of the sharding / reduction logic, not of the kernels)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    from oracle import bruno_oracle as O
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    spec = O.spec_for('en2_noconv_mnist_tp', n_units=32)
    P = O.init_params(spec, seed=3, dtype=torch.float64, perturb_process=True)
    x = O.synthetic_batch(spec, 4, 3, seed=4)
    shard = torch.from_numpy(np.split(x.numpy(), world)[rank])                 # train.py:184
    loss, grads = O.loss_and_grads(spec, P, shard, student_grad_scale=1.0)     # per-tower loss / grads, train.py:88-91
    keys = list(grads.keys())
    flat = torch.cat([grads[k].reshape(-1) for k in keys])                     # the flat gradient buffer
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)                                # train.py:97-102
    flat /= world                                                              # train.py:103-105
    scale = 0.1                                                                # student_grad_schedule value
    off = 0
    for k in keys:
        n = grads[k].numel()
        if any(t in k for t in O.STUDENT_PARAM_TAGS):                          # train.py:108-111
            flat[off:off + n] *= scale
        off += n
    dist.all_reduce(loss, op=dist.ReduceOp.SUM)
    loss /= world                                                              # train.py:113-116
    if rank == 0:
        full_loss, full = O.loss_and_grads(spec, P, x, student_grad_scale=scale)
        ref = torch.cat([full[k].reshape(-1) for k in keys])
        torch.save({'err': (flat - ref).abs().max().item(), 'scale': ref.abs().max().item(),
                    'loss_err': abs(loss.item() - full_loss.item())}, os.path.join(out_dir, 'result.pt'))
    dist.destroy_process_group()


def test_sharded_gradients_equal_full_batch(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.start_processes(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True, start_method='spawn')
    r = torch.load(os.path.join(str(tmp_path), 'result.pt'))
    assert r['err'] <= 1e-9 * max(r['scale'], 1.0), r
    assert r['loss_err'] <= 1e-9, r


def test_trainer_student_ranges_and_flat_layout():
    """Host logic of BrunoModel.flatten_parameters without a GPU: offsets are 16-byte aligned, views alias the flat
    buffer, gradients accumulate into the flat gradient buffer, student ranges cover exactly the prior_* variables."""
    sys.path.insert(0, ROOT)
    from bruno_b200.model import BrunoModel

    class FakeLayer(object):
        def __init__(self, name, shapes):
            self.name = name
            self.t = [torch.randn(*s).requires_grad_(True) for s in shapes]

        def named_parameters(self):
            return [('%s/p%d' % (self.name, i), t) for i, t in enumerate(self.t)]

        def _set_parameters(self, ts):
            self.t = list(ts)

        def named_tf_variables(self, prefix='model/'):
            return []

    class FakeProcess(object):
        def __init__(self):
            self.var_vbl = torch.zeros(1, 5).requires_grad_(True)
            self.nu_vbl = torch.ones(1, 5).requires_grad_(True)
            self.corr_vbl = torch.full((1, 5), 2.0).requires_grad_(True)

        def named_parameters(self, prefix=''):
            return [('prior_var', self.var_vbl), ('prior_nu', self.nu_vbl), ('prior_corr', self.corr_vbl)]

    m = BrunoModel([FakeLayer('a', [(3, 3), (7,)])], [FakeLayer('b', [(2, 5)])], FakeProcess())
    before = [p.detach().clone() for p in m.parameters()]
    flat, grad = m.flatten_parameters()
    after = m.parameters()
    for b, a in zip(before, after):
        assert torch.equal(b, a.detach())
        assert a.data_ptr() % 16 == 0
        assert flat.data_ptr() <= a.data_ptr() < flat.data_ptr() + flat.numel() * 4
    loss = sum((p * p).sum() for p in after)
    loss.backward()
    assert torch.allclose(grad[:9].view(3, 3), 2 * after[0].detach())
    rng = m.student_param_ranges()
    assert len(rng) == 3 and all(n == 5 for _, n in rng)
    flat.mul_(0.5)                                       # an optimiser update of the flat buffer is seen through the views
    assert torch.allclose(after[0].detach(), 0.5 * before[0])


def test_uncovered_ranges_of_the_bucketed_allreduce():
    """Host logic of the overlapped allreduce (trainer.py): what the per-layer calls did not cover is reduced last, once."""
    from bruno_b200.trainer import uncovered_ranges
    assert uncovered_ranges([], 10) == [(0, 10)]
    assert uncovered_ranges([(0, 10)], 10) == []
    assert uncovered_ranges([(6, 2), (2, 2)], 10) == [(0, 2), (4, 2), (8, 2)]
    assert uncovered_ranges([(0, 4), (4, 4)], 12) == [(8, 4)]
    covered = sorted([(2, 2), (6, 2)] + uncovered_ranges([(6, 2), (2, 2)], 10))
    assert sum(n for _, n in covered) == 10 and all(covered[i][0] + covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))


def test_test_time_defaults_follow_the_reference_semantics():
    """config_rnn/defaults.py and config_conditional/defaults.py: which command-line fields override which knobs."""
    import argparse
    import importlib
    d = importlib.reload(importlib.import_module('bruno_b200.config_rnn.defaults'))
    assert (d.seq_len, d.seq_len_few_shot, d.batch_size_few_shot, d.eval_only_last, d.eps_corr, d.mask_dims) == (20, 2, 20, False, None, False)
    d.set_parameters(argparse.Namespace(seq_len=6, batch_size=5, mask_dims=1, eps_corr=0.25, unrelated=3))
    assert d.seq_len == 6 and d.seq_len_few_shot == 6            # one flag feeds both (defaults.py:14-17)
    assert d.batch_size_few_shot is True                         # (sic) cast to bool in the reference (defaults.py:21)
    assert d.mask_dims is True and d.eps_corr == 0.25 and d.eval_only_last is False
    importlib.reload(d)
    c = importlib.reload(importlib.import_module('bruno_b200.config_conditional.defaults'))
    assert (c.seq_len, c.n_context) == (16, 1)
    c.set_parameters(argparse.Namespace(n_context=3))
    assert (c.seq_len, c.n_context) == (16, 3)
    importlib.reload(c)
