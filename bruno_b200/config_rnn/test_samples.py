"""Conditional sampling through the inverse flow -- counterpart of the reference's config_rnn/test_samples.py
(model call :48-52, loop :69-73).  The reference plots the samples with matplotlib (absent here, and plotting is out
of scope): the arrays are written to <metadata>/samples/sample_<set>_<i>_<same_image>.npy instead."""
import argparse
import importlib
import json
import os

import numpy as np
import torch

from .. import utils
from . import defaults


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--config_name', type=str, required=True, help='Configuration name')
    parser.add_argument('--seq_len', type=int, default=20, help='Sequence length')
    parser.add_argument('--set', type=str, default='test', help='Test or train part')
    parser.add_argument('--same_image', type=int, default=0, help='Same image as input')
    parser.add_argument('--plot_n', type=int, default=0, help='Keep only last x images. 0 = all of them')
    parser.add_argument('--n_sequences', type=int, default=10)
    parser.add_argument('--metadata_dir', type=str, default='metadata')
    args, _ = parser.parse_known_args(argv)
    defaults.set_parameters(args)
    print('input args:\n', json.dumps(vars(args), indent=4, separators=(',', ':')))
    device = torch.device('cuda', 0)
    config = importlib.import_module('bruno_b200.config_rnn.%s' % args.config_name)
    save_dir = utils.find_model_metadata(args.metadata_dir + '/', args.config_name)
    samples_dir = os.path.join(save_dir, 'samples')
    utils.autodir(samples_dir)
    utils.load_checkpoint_into(config, save_dir, device)
    if args.set == 'test':
        data_iter = config.test_data_iter
    elif args.set == 'train':
        data_iter = config.train_data_iter
    else:
        raise ValueError('wrong set')
    data_iter.batch_size = config.sample_batch_size
    out = []
    for i, (x_batch, y_batch) in zip(range(args.n_sequences), data_iter.generate_each_digit(same_image=args.same_image)):
        print("Generating samples", i)
        sampled_xx, log_probs = config.build_model(torch.from_numpy(x_batch).to(device), sampling_mode=True)
        sampled_xx = sampled_xx.cpu().numpy()
        if args.plot_n > 0:
            sampled_xx = sampled_xx[:, -args.plot_n:]
        path = os.path.join(samples_dir, 'sample_%s_%s_%s.npy' % (args.set, i, args.same_image))
        np.save(path, sampled_xx)
        out.append(sampled_xx)
    return out


if __name__ == '__main__':
    main()
