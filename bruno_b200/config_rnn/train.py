"""Training driver -- B200 counterpart of the reference's config_rnn/train.py (same flags, schedules, prints, metadata
layout).  Multi-GPU: launch one process per GPU,
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 -m bruno_b200.config_rnn.train --config_name ... --nr_gpu N
each rank takes batch_size / nr_gpu sequences of every batch (np.split, train.py:184) and gradients are allreduced
with NCCL (bruno_b200/trainer.py).  `--max_iter` (extra flag) bounds the run."""
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


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--config_name', type=str, required=True, help='Configuration name')
    parser.add_argument('--nr_gpu', type=int, default=1, help='How many GPUs to distribute the training across?')
    parser.add_argument('--resume', type=int, default=0, help='Resume training from a checkpoint?')
    parser.add_argument('--tf_seed', type=int, default=0, help='rng seed (the reference seeds TF with it)')
    parser.add_argument('--max_iter', type=int, default=None, help='override config.max_iter')
    parser.add_argument('--metadata_dir', type=str, default='metadata')
    args = parser.parse_args(argv)
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    assert args.nr_gpu == world, '--nr_gpu must equal the number of launched ranks (train.py:23)'
    if rank == 0:
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
    last_iteration = 0
    if not args.resume:
        experiment_id = '%s-%s' % (args.config_name.split('.')[-1], time.strftime("%Y_%m_%d", time.localtime()))
        save_dir = os.path.join(args.metadata_dir, experiment_id)
        if rank == 0:
            utils.autodir(save_dir)
    else:
        save_dir = utils.find_model_metadata(args.metadata_dir + '/', args.config_name)
        experiment_id = os.path.dirname(save_dir).split('/')[-1]
        with open(os.path.join(save_dir, 'meta.pkl'), 'rb') as f:
            resumed_metadata = pickle.load(f)
        last_iteration = resumed_metadata['iteration']
        print('Last iteration', last_iteration)
        print('Last learning rate', resumed_metadata['lr'])
    if rank == 0:
        utils.autodir('logs')
        sys.stdout = utils.Logger('logs/%s.log' % experiment_id)
        print('exp_id', experiment_id)

    trainer = Trainer(config, device, world_size=world, rank=rank)
    lr = config.learning_rate
    student_grad_scale = config.scale_student_grad
    print_every = 100
    train_iter_losses = []
    if args.resume:
        losses_eval_valid = resumed_metadata['losses_eval_valid']
        losses_eval_train = resumed_metadata['losses_eval_train']
        losses_avg_train = resumed_metadata['losses_avg_train']
    else:
        losses_eval_valid, losses_eval_train, losses_avg_train = [], [], []

    start_time = prev_time = time.time()
    for iteration, x_batch in zip(range(0, max_iter), config.train_data_iter.generate()):
        if hasattr(config, 'learning_rate_schedule') and iteration in config.learning_rate_schedule:
            lr = np.float32(config.learning_rate_schedule[iteration])
        elif hasattr(config, 'lr_decay'):
            lr *= config.lr_decay
        if hasattr(config, 'student_grad_schedule') and iteration in config.student_grad_schedule:
            student_grad_scale = np.float32(config.student_grad_schedule[iteration])
            if rank == 0:
                print('setting student grad scale to %.7f' % config.student_grad_schedule[iteration])

        x_full = torch.from_numpy(np.ascontiguousarray(x_batch))
        if trainer.model is None:
            # variables + data-dependent init use the FULL batch on one device (train.py:60-61,178-182)
            trainer.materialize(x_full.to(device), init=not args.resume)
            if args.resume:
                ck = torch.load(os.path.join(save_dir, 'params.pt'), map_location='cpu', weights_only=False)
                trainer.load_state_dict(ck)
            if rank == 0:
                print('initializing the model... %d parameters' % trainer.n_params)
        if args.resume and iteration < last_iteration:
            if iteration % (print_every * 10) == 0 and rank == 0:
                print(iteration, 'skipping training')
            continue
        if iteration == 0:
            continue                                        # iteration 0 = init pass only, no weight update (train.py:178-182)

        x_shard = np.split(x_batch, world)[rank]            # train.py:184
        x = torch.from_numpy(np.ascontiguousarray(x_shard)).pin_memory().to(device, non_blocking=True)
        loss = trainer.train_step(x, lr, float(student_grad_scale))
        if world > 1:
            dist.all_reduce(loss)                           # train_loss = sum of tower losses / nr_gpu (train.py:113-116)
            loss /= world
        l = loss.item()
        train_iter_losses.append(l)
        if np.isnan(l):
            print('Loss is NaN')
            sys.exit(0)
        if rank != 0:
            continue
        if (iteration + 1) % print_every == 0:
            avg_train_loss = np.mean(train_iter_losses)
            losses_avg_train.append(avg_train_loss)
            train_iter_losses = []
            print('%d/%d train_loss=%6.8f bits/value=%.3f' % (iteration + 1, max_iter, avg_train_loss,
                                                              avg_train_loss / config.ndim / np.log(2.)))
        if hasattr(config, 'validate_every') and (iteration + 1) % config.validate_every == 0:
            print('\n Validating ...')
            losses = []
            for _, x_valid_batch in zip(range(0, config.n_valid_batches), config.valid_data_iter.generate()):
                lp = trainer.eval_step(torch.from_numpy(x_valid_batch).to(device))[0]
                losses.append(config.loss(lp).item())
            losses_eval_valid.append(np.mean(losses))
            print('valid loss', losses_eval_valid[-1])
        if (iteration + 1) % config.save_every == 0 or iteration + 1 == max_iter:
            current_time = time.time()
            eta_time = (max_iter - iteration) / config.save_every * (current_time - prev_time)
            prev_time = current_time
            print('ETA: ', time.strftime("%H:%M:%S", time.gmtime(eta_time)))
            print('Saving model (iteration %s):' % iteration, experiment_id)
            print('current learning rate:', lr)
            trainer.iteration = iteration + 1
            torch.save(trainer.state_dict(), os.path.join(save_dir, 'params.pt'))
            with open(os.path.join(save_dir, 'meta.pkl'), 'wb') as f:
                pickle.dump({'lr': lr, 'iteration': iteration + 1, 'losses_avg_train': losses_avg_train,
                             'losses_eval_valid': losses_eval_valid, 'losses_eval_train': losses_eval_train}, f)
            corr = config.student_layer.corr.detach().cpu().numpy().flatten()
            for th in (0.01, 0.1, 0.2, 0.3, 0.5, 0.7):
                print(th, np.sum(corr > th))
            print('corr min-max:', np.min(corr), np.max(corr))
            var = config.student_layer.var.detach().cpu().numpy().flatten()
            print('var min-max:', np.min(var), np.max(var))
            if hasattr(config.student_layer, 'nu'):
                nu = config.student_layer.nu.detach().cpu().numpy().flatten()
                print('nu median-min-max:', np.median(nu), np.min(nu), np.max(nu))
    if rank == 0:
        print('Total time: ', time.strftime("%H:%M:%S", time.gmtime(time.time() - start_time)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
