"""Episodic fine-tuning driver (reference: config_rnn/train_finetune.py -- same flag, schedules, prints and checkpoint
layout).  Restores the base model's checkpoint (config.base_metadata_path, train_finetune.py:105-107), then every
iteration takes ONE meta-batch of config.meta_batch_size episodes x config.batch_size candidate sequences
(train_finetune.py:56-57,109), minimises the softmax-over-candidates loss of the config (ft:222-226) with the same
TF-1.7 optimiser semantics, student-gradient scaling (:64-67) and NaN guard (:124-126) as train.py.  Single GPU, like the
reference (:19).  `--max_iter` / `--metadata_dir` / `--base` are extra flags (bounded runs, explicit base checkpoint)."""
import argparse
import importlib
import json
import os
import sys
import time

import numpy as np
import torch

from .. import utils
from ..trainer import Trainer
from .train import Schedules, RunDir, report_process

PRINT_EVERY = 100


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--config_name', type=str, required=True, help='Configuration name')
    parser.add_argument('--max_iter', type=int, default=None, help='override config.max_iter')
    parser.add_argument('--metadata_dir', type=str, default='metadata')
    parser.add_argument('--base', type=str, default=None, help='directory of the base checkpoint (default: config.base_metadata_path)')
    parser.add_argument('--print_every', type=int, default=PRINT_EVERY)
    args = parser.parse_args(argv)
    print('input args:\n', json.dumps(vars(args), indent=4, separators=(',', ':')))
    np.random.seed(seed=42)
    torch.manual_seed(42)
    device = torch.device('cuda', 0)
    torch.cuda.set_device(device)

    config = importlib.import_module('bruno_b200.config_rnn.%s' % args.config_name)
    max_iter = args.max_iter or config.max_iter
    run = RunDir(args.metadata_dir, args.config_name, resume=False, make=True)
    utils.autodir('logs')
    sys.stdout = utils.Logger('logs/%s.log' % run.experiment_id)
    print('exp_id', run.experiment_id)

    base = args.base or config.base_metadata_path
    if base is None:
        raise ValueError('No metadata files for the base config: train it first (config_rnn.train) or pass --base')
    trainer = Trainer(config, device)
    sched = Schedules(config, verbose=True)
    window = []
    for iteration, (x_batch, _) in zip(range(max_iter), config.train_data_iter.generate()):
        x = torch.from_numpy(np.ascontiguousarray(x_batch)).pin_memory().to(device, non_blocking=True)
        if trainer.model is None:
            trainer.materialize(x, init=False)
            print('Number of parameters', trainer.n_params)
            print('restoring parameters from', base)
            ck = torch.load(os.path.join(base, 'params.pt'), map_location='cpu', weights_only=False)
            config.model.load_tf_variables(ck['variables'])
            print('starting training')
        started = time.time()
        sched.advance(iteration)
        loss = trainer.train_step(x, sched.lr, float(sched.student_scale)).item()
        window.append(loss)
        if np.isnan(loss):
            print('Loss is NaN')
            sys.exit(0)
        done = iteration + 1
        if done % args.print_every == 0:
            print('%d/%d train_loss=%6.8f time/batch=%.2f' % (done, max_iter, float(np.mean(window)), time.time() - started))
            window = []
        if done % config.save_every == 0 or done == max_iter:
            print('saving model', run.experiment_id)
            print('learning rate', sched.lr)
            run.save(trainer, done, sched.lr)
            report_process(config.student_layer)
    return run.path


if __name__ == '__main__':
    main()
