"""Samples from the PRIOR through the inverse flow -- counterpart of the reference's config_rnn/test_samples_prior.py
(model call :42-46; loop :56-64: `samples[0, 0]` = sample 0 of column 0 of the sampling branch, i.e. drawn from the
process before any observation, pushed through the inverse Real NVP).  n_row * n_row independent prior samples per grid,
three grids.  The reference renders the grid with matplotlib (absent here; plotting is out of scope): the array
[n_row, n_row, H, W, C] is written to <metadata>/samples/prior_sample_<i>.npy."""
import argparse
import importlib
import json
import os

import numpy as np
import torch

from .. import utils


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--config_name', type=str, required=True, help='Configuration name')
    parser.add_argument('--n_row', type=int, default=10, help='Number of rows')
    parser.add_argument('--n_grids', type=int, default=3)
    parser.add_argument('--metadata_dir', type=str, default='metadata')
    args, _ = parser.parse_known_args(argv)
    print('input args:\n', json.dumps(vars(args), indent=4, separators=(',', ':')))
    torch.manual_seed(42)
    device = torch.device('cuda', 0)
    config = importlib.import_module('bruno_b200.config_rnn.%s' % args.config_name)
    save_dir = utils.find_model_metadata(args.metadata_dir + '/', args.config_name)
    samples_dir = os.path.join(save_dir, 'samples')
    utils.autodir(samples_dir)
    print('exp_id', os.path.dirname(save_dir))
    utils.load_checkpoint_into(config, save_dir, device)
    config.test_data_iter.batch_size = config.sample_batch_size
    out = []
    for i, (x_batch, y_batch) in zip(range(args.n_grids), config.test_data_iter.generate_each_digit()):
        print("Generating prior samples", i)
        x = torch.from_numpy(x_batch).to(device)
        prior_samples = []
        for j in range(args.n_row * args.n_row):
            sampled_xx = config.build_model(x, sampling_mode=True)[0]       # fresh random draws on every call
            prior_samples.append(sampled_xx[0, 0].cpu().numpy())
        prior_samples = np.asarray(prior_samples)
        prior_samples = np.reshape(prior_samples, (args.n_row, args.n_row) + prior_samples.shape[1:])
        np.save(os.path.join(samples_dir, 'prior_sample_%s.npy' % i), prior_samples)
        out.append(prior_samples)
    return out


if __name__ == '__main__':
    main()
