"""Training driver (reference: config_rnn/train.py -- same flags, schedules, prints and metadata layout).

Multi-GPU: one process per GPU,
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 -m bruno_b200.config_rnn.train --config_name ... --nr_gpu N
Every rank takes batch_size / nr_gpu sequences of each batch (np.split, train.py:184); gradients are summed by one NCCL
allreduce of the flat buffer and divided by nr_gpu in the optimiser kernel (bruno_b200/trainer.py).
`--max_iter` (extra flag) bounds the run.  The driver is organised around three small helpers: Schedules (learning rate
and student-gradient scale per iteration), RunDir (experiment directory, checkpoint + metadata) and report_process
(the prior statistics the reference prints at every checkpoint).
"""
import argparse
import importlib
import json
import os
import pickle
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

from .. import utils
from ..trainer import Trainer

PRINT_EVERY = 100


class Schedules(object):
    """lr: explicit `learning_rate_schedule` entries win over the multiplicative `lr_decay` (train.py:160-166);
    student-gradient scale: piecewise constant `student_grad_schedule` (train.py:168-170)."""
    def __init__(self, config, verbose):
        self.lr = config.learning_rate
        self.student_scale = config.scale_student_grad
        self._lr_steps = getattr(config, 'learning_rate_schedule', None)
        self._decay = getattr(config, 'lr_decay', None)
        self._student_steps = getattr(config, 'student_grad_schedule', None)
        self._verbose = verbose

    def advance(self, iteration):
        if self._lr_steps is not None and iteration in self._lr_steps:
            self.lr = np.float32(self._lr_steps[iteration])
        elif self._decay is not None:
            self.lr *= self._decay
        if self._student_steps is not None and iteration in self._student_steps:
            self.student_scale = np.float32(self._student_steps[iteration])
            if self._verbose:
                print('setting student grad scale to %.7f' % self._student_steps[iteration])


class RunDir(object):
    """metadata/<config>-<date>/ with params.pt (TF variable names) and meta.pkl (iteration, lr, loss histories)."""
    def __init__(self, root, config_name, resume, make):
        self.history = {'losses_eval_valid': [], 'losses_eval_train': [], 'losses_avg_train': []}
        self.last_iteration = 0
        if resume:
            self.path = utils.find_model_metadata(root + '/', config_name)
            self.experiment_id = os.path.dirname(self.path).split('/')[-1]
            with open(os.path.join(self.path, 'meta.pkl'), 'rb') as f:
                meta = pickle.load(f)
            self.last_iteration = meta['iteration']
            for key in self.history:
                self.history[key] = meta[key]
            print('Last iteration', self.last_iteration)
            print('Last learning rate', meta['lr'])
        else:
            self.experiment_id = '%s-%s' % (config_name.split('.')[-1], time.strftime("%Y_%m_%d", time.localtime()))
            self.path = os.path.join(root, self.experiment_id)
            if make:
                utils.autodir(self.path)

    def restore(self, trainer):
        trainer.load_state_dict(torch.load(os.path.join(self.path, 'params.pt'), map_location='cpu', weights_only=False))

    def save(self, trainer, iteration, lr):
        trainer.iteration = iteration
        torch.save(trainer.state_dict(), os.path.join(self.path, 'params.pt'))
        with open(os.path.join(self.path, 'meta.pkl'), 'wb') as f:
            pickle.dump(dict(self.history, lr=lr, iteration=iteration), f)


def report_process(layer):
    """Prior statistics printed with every checkpoint (train.py:241-254)."""
    corr = layer.corr.detach().cpu().numpy().flatten()
    for threshold in (0.01, 0.1, 0.2, 0.3, 0.5, 0.7):
        print(threshold, np.sum(corr > threshold))
    print('corr min-max:', np.min(corr), np.max(corr))
    var = layer.var.detach().cpu().numpy().flatten()
    print('var min-max:', np.min(var), np.max(var))
    if hasattr(layer, 'nu'):
        nu = layer.nu.detach().cpu().numpy().flatten()
        print('nu median-min-max:', np.median(nu), np.min(nu), np.max(nu))


def validate(config, trainer, device, data_iter, n_batches):
    """train.py:201-223: `eval_loss` if the config defines it (else `loss`), averaged over `n_batches` batches drawn with a
    FRESH RandomState(42) -- every validation sees the same batches and the iterators' own rng state is untouched."""
    loss_fn = getattr(config, 'eval_loss', config.loss)
    rng = np.random.RandomState(42)
    losses = [loss_fn(trainer.eval_step(torch.from_numpy(np.ascontiguousarray(xb)).to(device))[0]).item()
              for _, xb in zip(range(n_batches), data_iter.generate(rng))]
    return float(np.mean(losses))


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--config_name', type=str, required=True, help='Configuration name')
    parser.add_argument('--nr_gpu', type=int, default=1, help='How many GPUs to distribute the training across?')
    parser.add_argument('--resume', type=int, default=0, help='Resume training from a checkpoint?')
    parser.add_argument('--tf_seed', type=int, default=0, help='rng seed (the reference seeds TF with it)')
    parser.add_argument('--max_iter', type=int, default=None, help='override config.max_iter')
    parser.add_argument('--metadata_dir', type=str, default='metadata')
    args = parser.parse_args(argv)
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    chief = rank == 0
    assert args.nr_gpu == world, '--nr_gpu must equal the number of launched ranks (train.py:23)'
    if chief:
        print('input args:\n', json.dumps(vars(args), indent=4, separators=(',', ':')))
    np.random.seed(seed=42)
    torch.manual_seed(args.tf_seed)
    torch.cuda.set_device(local_rank)
    device = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)

    config = importlib.import_module('bruno_b200.config_rnn.%s' % args.config_name)
    assert config.batch_size % world == 0, 'batch_size must be divisible by nr_gpu (train.py:84)'
    max_iter = args.max_iter or config.max_iter
    run = RunDir(args.metadata_dir, args.config_name, bool(args.resume), make=chief)
    if chief:
        utils.autodir('logs')
        sys.stdout = utils.Logger('logs/%s.log' % run.experiment_id)
        print('exp_id', run.experiment_id)

    trainer = Trainer(config, device, world_size=world, rank=rank)
    sched = Schedules(config, verbose=chief)
    window = []                                             # losses since the last print
    started = last_save = time.time()
    for iteration, x_batch in zip(range(max_iter), config.train_data_iter.generate()):
        sched.advance(iteration)
        if trainer.model is None:
            # variables + data-dependent init use the FULL batch on one device (train.py:60-61,178-182)
            trainer.materialize(torch.from_numpy(np.ascontiguousarray(x_batch)).to(device), init=not args.resume)
            if args.resume:
                run.restore(trainer)
            if chief:
                print('initializing the model... %d parameters' % trainer.n_params)
        if args.resume and iteration < run.last_iteration:
            if chief and iteration % (PRINT_EVERY * 10) == 0:
                print(iteration, 'skipping training')
            continue
        if iteration == 0:
            continue                                        # iteration 0 = init pass only, no weight update (train.py:178-182)

        shard = np.ascontiguousarray(np.split(x_batch, world)[rank])                    # train.py:184
        loss = trainer.train_step(torch.from_numpy(shard).pin_memory().to(device, non_blocking=True), sched.lr,
                                  float(sched.student_scale))
        if world > 1:
            dist.all_reduce(loss)                           # train_loss = sum of tower losses / nr_gpu (train.py:113-116)
            loss /= world
        value = loss.item()
        window.append(value)
        if np.isnan(value):
            print('Loss is NaN')
            sys.exit(0)
        if not chief:
            continue
        done = iteration + 1
        if done % PRINT_EVERY == 0:
            mean = np.mean(window)
            run.history['losses_avg_train'].append(mean)
            window = []
            print('%d/%d train_loss=%6.8f bits/value=%.3f' % (done, max_iter, mean, mean / config.ndim / np.log(2.)))
        if hasattr(config, 'validate_every') and done % config.validate_every == 0:
            print('\n Validating ...')
            run.history['losses_eval_valid'].append(validate(config, trainer, device, config.valid_data_iter, config.n_valid_batches))
            print('valid loss', run.history['losses_eval_valid'][-1])
            run.history['losses_eval_train'].append(validate(config, trainer, device, config.train_data_iter, config.n_valid_batches * 10))
            print('train loss', run.history['losses_eval_train'][-1])
        if done % config.save_every == 0 or done == max_iter:
            now = time.time()
            eta = (max_iter - iteration) / config.save_every * (now - last_save)
            last_save = now
            print('ETA: ', time.strftime("%H:%M:%S", time.gmtime(eta)))
            print('Saving model (iteration %s):' % iteration, run.experiment_id)
            print('current learning rate:', sched.lr)
            run.save(trainer, done, sched.lr)
            report_process(config.student_layer)
    if chief:
        print('Total time: ', time.strftime("%H:%M:%S", time.gmtime(time.time() - started)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
