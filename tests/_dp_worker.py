"""Worker of tests/test_trainer_2gpu.py (launched by torch.distributed.run, one process per GPU, NCCL).

Every rank builds the same model from the same seeded weights, takes its np.split shard of ONE global batch
(config_rnn/train.py:184) and runs Trainer.train_step (forward, hand-written backward, ncclAllReduce of the flat gradient
buffer, 1/nr_gpu, student-grad scaling, TF-1.7 RMSProp).  Rank 0 then repeats the steps alone on the FULL batch with a
fresh world_size-1 trainer and writes both flat parameter buffers (and the first step's gradient buffers) to argv[1]."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bruno_oracle as O          # noqa: E402  (seeded weights / inputs only: test infrastructure)


def run(name, B, T, world, rank, device, overlap, P, xs, lr, sscale):
    from bruno_b200.trainer import Trainer
    cfg = importlib.reload(importlib.import_module('bruno_b200.config_rnn.' + name))
    cfg.optimizer = 'rmsprop'
    tr = Trainer(cfg, device, world_size=world, rank=rank)
    shard0 = torch.from_numpy(np.split(xs[0].numpy(), world)[rank]).float().to(device)
    tr.materialize(shard0, init=False)
    cfg.model.load_tf_variables(P)
    tr.overlap_allreduce = overlap
    grads = None
    for x in xs:
        shard = torch.from_numpy(np.split(x.numpy(), world)[rank]).float().to(device)     # train.py:184
        tr.train_step(shard, lr, sscale)
        if grads is None:
            torch.cuda.synchronize()
            grads = (tr.flat_grad / world).cpu().clone()      # summed over ranks (train.py:97-102), student-scaled; / nr_gpu
    torch.cuda.synchronize()
    return tr.flat.detach().cpu().clone(), grads


def main():
    out_path = sys.argv[1]
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    device = 'cuda:%d' % local
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device(device))
    results = {}
    for name, B, T in (('en2_noconv_mnist_tp', 8, 5), ('bn2_omniglot_tp', 4, 6)):
        spec = O.spec_for(name)
        P = O.init_params(spec, seed=3, dtype=torch.float64, perturb_process=True)
        xs = [O.synthetic_batch(spec, B, T, seed=20 + i).float() for i in range(2)]
        O.data_dependent_init(spec, P, xs[0][:2].double())
        P = O.cast_params(P, torch.float32)
        for overlap in (False, True):
            flat, grads = run(name, B, T, world, rank, device, overlap, P, xs, 1e-3, 0.1)
            results[(name, overlap)] = (flat, grads)
        dist.barrier()
        if rank == 0:
            flat1, grads1 = run(name, B, T, 1, 0, device, False, P, xs, 1e-3, 0.1)
            results[(name, 'single')] = (flat1, grads1)
        dist.barrier()
    if rank == 0:
        torch.save(results, out_path)
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
