"""torchrun --nproc-per-node 2 tools/dbg_overlap.py : the per-layer overlapped allreduce gives the same parameters as one
allreduce of the whole flat buffer."""
import os, sys, importlib, copy
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, '/root/repo')
rank, world, lr_ = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(lr_); dev = torch.device('cuda', lr_)
os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
dist.init_process_group('nccl', device_id=dev)
from bruno_b200.trainer import Trainer
from bruno_b200.synthetic_data import SyntheticExchSeqDataIterator
from bruno_b200.config_rnn import defaults
defaults.seq_len = 4
cfg = importlib.import_module('bruno_b200.config_rnn.bn2_omniglot_tp')
cfg.batch_size = 4 * world
it = SyntheticExchSeqDataIterator(seq_len=4, batch_size=4, obs_shape=(28, 28, 1), rng=np.random.RandomState(42 + rank), noise_rng=np.random.RandomState(7 + rank))
xs = [torch.from_numpy(next(it.generate())).to(dev) for _ in range(3)]
tr = Trainer(cfg, dev, world_size=world, rank=rank).materialize(xs[0], init=True)
flat0 = tr.flat.clone(); slot0 = tr.slots[0].clone()
res = []
for overlap in (True, False):
    tr.flat.copy_(flat0); tr.slots[0].copy_(slot0); tr.overlap_allreduce = overlap
    for x in xs:
        loss = tr.train_step(x, 1e-3, 0.5)
    torch.cuda.synchronize()
    res.append(tr.flat.clone())
same = torch.equal(res[0], res[1])
mx = (res[0] - res[1]).abs().max().item()
other = [torch.zeros_like(res[0]) for _ in range(world)]
dist.all_gather(other, res[0])
print('rank %d: overlapped == single allreduce: %s (max diff %.3e); ranks agree: %s; %d layer ranges' % (
    rank, same, mx, all(torch.equal(o, res[0]) for o in other), len(tr._layer_ranges)))
dist.destroy_process_group()
