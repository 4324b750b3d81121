"""Where the conv fwd/dgrad time of a bn2_omniglot_tp training step goes: bench.py's graph-replayed conv launch sequence
(conv_sequence_graph_ms) split by resolution and direction.  Usage: python tools/bench_conv_step.py"""
import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bruno_b200.config_rnn import defaults
from bruno_b200.trainer import Trainer
from bruno_b200.synthetic_data import SyntheticExchSeqDataIterator
dev = torch.device('cuda', 0)
defaults.seq_len = 20
cfg = importlib.import_module('bruno_b200.config_rnn.bn2_omniglot_tp')
cfg.batch_size = 32
it = SyntheticExchSeqDataIterator(seq_len=20, batch_size=32, obs_shape=(28, 28, 1), kind='strokes',
                                  rng=np.random.RandomState(42), noise_rng=np.random.RandomState(317070))
x = torch.from_numpy(next(it.generate())).to(dev)
tr = Trainer(cfg, dev, world_size=1, rank=0).materialize(x, init=True)
tr.train_step(x, 1e-4, 0.1)
tot = 0.0
for h in (28, 14):
    for d in ('fwd', 'bwd'):
        ms, n = bench.conv_sequence_graph_ms(cfg, dev, direction=d, height=h)
        tot += ms
        print('%dx%d %s: %4d launches  %.3f ms  %.1f us per launch' % (h, h, d, n, ms, 1e3 * ms / n), flush=True)
ms, n = bench.conv_sequence_graph_ms(cfg, dev)
print('whole sequence: %d launches %.3f ms (sum of parts %.3f)' % (n, ms, tot))
