"""Few-shot episode (20 candidates x 2 frames, forward only, one CUDA graph) and the eval step with the conv layer chain
off (0), on for >= 2 tiles per CTA (1) and on for every shape (2):  python tools/bench_fewshot_chain.py"""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bruno_b200 import _lib  # noqa: E402
from bruno_b200.trainer import Trainer  # noqa: E402
from bruno_b200.synthetic_data import SyntheticExchSeqDataIterator, SyntheticFewShotIterator  # noqa: E402


def timed(fn, k):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(k):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def main():
    dev = torch.device('cuda', 0)
    cfg = importlib.import_module('bruno_b200.config_rnn.bn2_omniglot_tp')
    it = SyntheticExchSeqDataIterator(seq_len=20, batch_size=32, obs_shape=(28, 28, 1), kind='strokes')
    x = torch.from_numpy(next(it.generate())).to(dev)
    tr = Trainer(cfg, dev).materialize(x, init=True)
    for _ in range(3):
        tr.train_step(x, 1e-3, 0.1)
    fs_it = SyntheticFewShotIterator(seq_len=2, batch_size=20, obs_shape=(28, 28, 1), rng=np.random.RandomState(7), n_classes=24)
    fs_x = [torch.from_numpy(e[0]).to(dev) for e, _ in zip(fs_it.generate(), range(4))]
    ref = None
    for mode in (0, 1, 2, 0):
        assert _lib.load().bruno_debug_conv_chain(mode) == 0
        graphed = cfg.model.graphed_forward(fs_x[0])
        out = graphed(fs_x[1])[0].clone()
        if ref is None:
            ref = out
        same = bool(torch.equal(out, ref))
        for i in range(5):
            graphed(fs_x[i % 4])
        ms = timed(lambda i: graphed(fs_x[i % 4])[0][:, -1].argmax(), 50)
        for i in range(2):
            tr.eval_step(x)
        ms_eval = timed(lambda i: tr.eval_step(x), 10)
        ms_train = timed(lambda i: tr.train_step(x, 1e-3, 0.1), 10)
        print('chain mode %d: few-shot episode %.3f ms (graph), eval step %.3f ms, train step %.3f ms, bit-identical to mode 0: %s' % (
            mode, ms, ms_eval, ms_train, same), flush=True)
    _lib.load().bruno_debug_conv_chain(0)


if __name__ == '__main__':
    main()
