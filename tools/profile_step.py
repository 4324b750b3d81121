"""One training step (and one eval step) of bn2_omniglot_tp between cudaProfilerStart/Stop, for
   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
Usage: python tools/profile_step.py [train|eval|both] [B]"""
import importlib, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bruno_b200.trainer import Trainer
from bruno_b200.synthetic_data import SyntheticExchSeqDataIterator

what = sys.argv[1] if len(sys.argv) > 1 else 'both'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg = importlib.import_module('bruno_b200.config_rnn.bn2_omniglot_tp')
dev = torch.device('cuda', 0)
it = SyntheticExchSeqDataIterator(seq_len=20, batch_size=B, obs_shape=(28, 28, 1), kind='strokes')
x = torch.from_numpy(next(it.generate())).to(dev)
tr = Trainer(cfg, dev).materialize(x, init=True)
for _ in range(2):
    tr.train_step(x, 1e-3, 0.1)
    tr.eval_step(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
if what in ('train', 'both'):
    tr.train_step(x, 1e-3, 0.1)
if what in ('eval', 'both'):
    tr.eval_step(x)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print('done')
