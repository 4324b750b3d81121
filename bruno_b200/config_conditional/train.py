"""Training driver of Conditional BRUNO -- B200 counterpart of the reference's config_conditional/train.py (same flags,
schedules, iteration-0-is-the-init-pass quirk, NaN guard, prints).  Multi-GPU: one process per GPU,
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 -m bruno_b200.config_conditional.train --config_name m1_shapenet --nr_gpu N
each rank takes batch_size / nr_gpu sequences (np.split, train.py:186-187), gradients are summed by ONE NCCL allreduce of
the flat gradient buffer and divided by nr_gpu inside the optimiser kernel (train.py:99-107)."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

from .. import _lib, utils


class ConditionalTrainer(object):
    """Flat fp32 parameter / gradient buffers over config.parameters() + the TF-1.7 optimiser kernels."""
    def __init__(self, config, device, world_size=1, rank=0):
        self.cfg, self.device, self.world_size, self.rank = config, device, world_size, rank
        self.flat = None

    def materialize(self, x, y, init=True):
        cfg = self.cfg
        cfg.build_model(x[:1], y[:1])                       # creates the variables
        if init:
            cfg.build_model(x, y, init=True)                # data-dependent init on the full batch (train.py:60-61)
        params = cfg.parameters()
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]  # every tensor 16-byte aligned
        self.flat = torch.zeros(sum(sizes), device=self.device)
        self.flat_grad = torch.zeros_like(self.flat)
        self.student_ranges = []
        student = set(id(p) for _, p in cfg.student_layer.named_parameters())
        off = 0
        for p, sz in zip(params, sizes):
            view = self.flat[off:off + p.numel()].view(p.shape)
            view.copy_(p.detach())
            p.data = view                                   # the layer objects keep their tensors; storage moves to the flat buffer
            p.requires_grad_(True)
            p.grad = self.flat_grad[off:off + p.numel()].view(p.shape)
            if id(p) in student:
                self.student_ranges.append((off, p.numel()))
            off += sz
        if self.world_size > 1:
            dist.broadcast(self.flat, 0)
        opt = getattr(cfg, 'optimizer', 'adam')
        self.slots = (torch.ones_like(self.flat),) if opt == 'rmsprop' else (torch.zeros_like(self.flat), torch.zeros_like(self.flat))
        self.opt_steps = 0
        self.n_params = sum(p.numel() for p in params)
        return self

    def train_step(self, x, y, lr, student_scale=1.0):
        self.flat_grad.zero_()
        loss = self.cfg.loss(self.cfg.build_model(x, y)[0])
        loss.backward()
        if self.world_size > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM)
        if student_scale != 1.0:
            for off, n in self.student_ranges:              # train.py:109-113
                self.flat_grad[off:off + n].mul_(student_scale)
        gs = 1.0 / self.world_size
        n = self.flat.numel()
        self.opt_steps += 1
        if len(self.slots) == 1:
            _lib.call('bruno_rmsprop_tf17', self.flat, self.flat_grad, self.slots[0], n, float(lr), 0.9, 1e-10, gs)
        else:
            _lib.call('bruno_adam_tf17', self.flat, self.flat_grad, self.slots[0], self.slots[1], n, float(lr), 0.9, 0.999,
                      1e-8, self.opt_steps, gs)
        return loss.detach()

    def eval_step(self, x, y):
        with torch.no_grad():
            return self.cfg.build_model(x, y)

    def state_dict(self):
        return {'variables': {n: v.detach().cpu().clone() for n, v in self.cfg.named_tf_variables()},
                'slots': [s.cpu() for s in self.slots], 'opt_steps': self.opt_steps}

    def load_state_dict(self, sd):
        self.cfg.load_tf_variables(sd['variables'])             # the variables are views into the flat buffer
        for s, t in zip(self.slots, sd['slots']):
            s.copy_(t.to(self.device))
        self.opt_steps = sd['opt_steps']


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--config_name', type=str, required=True, help='Configuration name')
    parser.add_argument('--nr_gpu', type=int, default=1, help='How many GPUs to distribute the training across?')
    parser.add_argument('--resume', type=int, default=0, help='Resume training from a checkpoint?')
    parser.add_argument('--max_iter', type=int, default=None, help='override config.max_iter')
    parser.add_argument('--print_every', type=int, default=100)
    parser.add_argument('--metadata_dir', type=str, default='metadata')
    args = parser.parse_args(argv)
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    assert args.nr_gpu == world, '--nr_gpu must equal the number of launched ranks (train.py:22)'
    if rank == 0:
        print('input args:\n', json.dumps(vars(args), indent=4, separators=(',', ':')))
    np.random.seed(seed=42)
    torch.manual_seed(0)
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    config = importlib.import_module('bruno_b200.config_conditional.%s' % args.config_name)
    assert config.batch_size % world == 0
    max_iter = args.max_iter or config.max_iter
    last_iteration, resumed = 0, None
    if args.resume:                                        # train.py:31-41
        save_dir = utils.find_model_metadata(args.metadata_dir + '/', args.config_name)
        experiment_id = os.path.dirname(save_dir).split('/')[-1]
        resumed = torch.load(os.path.join(save_dir, 'params.pt'), map_location='cpu', weights_only=False)
        last_iteration = resumed['iteration']
        if rank == 0:
            print('Last iteration', last_iteration)
            print('Last learning rate', resumed['lr'])
    else:
        experiment_id = '%s-%s' % (args.config_name.split('.')[-1], time.strftime("%Y_%m_%d", time.localtime()))
        save_dir = os.path.join(args.metadata_dir, experiment_id)
        if rank == 0:
            utils.autodir(save_dir)
    if rank == 0:
        print('exp_id', experiment_id)
        if args.resume:
            print('Resuming training')
    trainer = ConditionalTrainer(config, device, world, rank)
    lr = config.learning_rate
    student_grad_scale = config.scale_student_grad
    train_iter_losses, losses_avg_train = [], []
    for iteration, (x_batch, y_batch) in zip(range(0, max_iter), config.train_data_iter.generate()):
        if hasattr(config, 'learning_rate_schedule') and iteration in config.learning_rate_schedule:
            lr = np.float32(config.learning_rate_schedule[iteration])
        elif hasattr(config, 'lr_decay'):
            lr *= config.lr_decay
        if hasattr(config, 'student_grad_schedule') and iteration in config.student_grad_schedule:
            student_grad_scale = np.float32(config.student_grad_schedule[iteration])
            if rank == 0:
                print('setting student grad scale to %.7f' % config.student_grad_schedule[iteration])
        if iteration == 0:                                  # the init pass, no weight update (train.py:178-183)
            if rank == 0:
                print('initializing the model...')
            trainer.materialize(torch.from_numpy(x_batch).to(device), torch.from_numpy(y_batch).to(device), init=not args.resume)
            if resumed is not None:
                trainer.load_state_dict(resumed)
            if rank == 0:
                print('Number of parameters', trainer.n_params)
            continue
        if args.resume and iteration < last_iteration:      # fast-forward schedules and data (train.py:172-175)
            if rank == 0 and iteration % (args.print_every * 10) == 0:
                print(iteration, 'skipping training')
            continue
        xs, ys = np.split(x_batch, world)[rank], np.split(y_batch, world)[rank]
        loss = trainer.train_step(torch.from_numpy(np.ascontiguousarray(xs)).to(device),
                                  torch.from_numpy(np.ascontiguousarray(ys)).to(device), lr, float(student_grad_scale))
        if world > 1:
            dist.all_reduce(loss)
            loss /= world
        l = loss.item()
        train_iter_losses.append(l)
        if np.isnan(l):
            print('Loss is NaN')
            sys.exit(0)
        if rank == 0 and (iteration + 1) % args.print_every == 0:
            avg_train_loss = np.mean(train_iter_losses)
            losses_avg_train.append(avg_train_loss)
            train_iter_losses = []
            print('%d/%d train_loss=%6.8f bits/value=%.3f' % (iteration + 1, max_iter, avg_train_loss,
                                                              avg_train_loss / config.ndim / np.log(2.)))
        if rank == 0 and ((iteration + 1) % config.save_every == 0 or iteration + 1 == max_iter):
            print('Saving model (iteration %s):' % iteration, experiment_id)
            torch.save(dict(trainer.state_dict(), iteration=iteration + 1, lr=float(lr)), os.path.join(save_dir, 'params.pt'))
    if world > 1:
        dist.destroy_process_group()
    return trainer, losses_avg_train


if __name__ == '__main__':
    main()
