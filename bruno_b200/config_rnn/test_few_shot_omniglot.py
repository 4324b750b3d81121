"""k-way n-shot classification -- counterpart of the reference's config_rnn/test_few_shot_omniglot.py: the score of a
candidate class is log_probs[:, -1] (the test image given the n shots of that class), decision = argmax (:64-80)."""
import argparse
import importlib
import json
import os
from collections import defaultdict

import numpy as np
import torch

from .. import utils
from . import defaults

np.set_printoptions(suppress=True)


def classify(config_name, seq_len, n_trials, batch_size, metadata_dir='metadata', max_episodes=None, restore=True):
    config = importlib.import_module('bruno_b200.config_rnn.%s' % config_name)
    device = torch.device('cuda', 0)
    assert seq_len == config.seq_len
    if restore:
        save_dir = utils.find_model_metadata(metadata_dir + '/', config_name)
        print('Building the model', os.path.dirname(save_dir).split('/')[-1])
        utils.load_checkpoint_into(config, save_dir, device)
    data_iter = config.test_data_iter2
    data_iter.batch_size = batch_size
    trial_accuracies = []
    for trial in range(n_trials):
        n_correct = n_total = 0
        x_number2scores = defaultdict(list)
        x_number2true_y, x_number2ys = {}, {}
        for iteration, (x_batch, y_batch, x_number) in enumerate(data_iter.generate(trial=trial)):
            if max_episodes is not None and iteration >= max_episodes:
                break
            y_true = int(y_batch[0, -1])
            with torch.no_grad():
                log_p = config.build_model(torch.from_numpy(x_batch).to(device), want_prior=False)[0].cpu().numpy()
            log_p = log_p.reshape((data_iter.batch_size, config.seq_len))[:, -1]
            x_number2scores[x_number].append(log_p)
            x_number2true_y[x_number] = y_true
            x_number2ys[x_number] = y_batch[:, 0]
        for k, v in x_number2scores.items():
            avg_score = np.mean(np.asarray(v), axis=0)
            max_idx = np.argmax(avg_score)
            if x_number2ys[k][max_idx] == x_number2true_y[k]:
                n_correct += 1
            n_total += 1
        acc = n_correct / n_total
        print(trial, 'accuracy', acc)
        print('n test examples', n_total)
        trial_accuracies.append(acc)
    print('---------------------------------------------')
    print(n_trials, config.seq_len)
    print(trial_accuracies)
    print('average accuracy over trials', np.mean(trial_accuracies))
    print('std accuracy over trials', np.std(trial_accuracies))
    return trial_accuracies


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument('--config_name', type=str, required=True, help='name of the configuration')
    parser.add_argument('--seq_len', type=int, default=2,
                        help='sequence length = number of shots + 1 (do not forget +1 for the test image)')
    parser.add_argument('--batch_size', type=int, default=20, help='batch_size = K-way')
    parser.add_argument('--n_trials', type=int, default=20, help='number of trials')
    parser.add_argument('--mask_dims', type=int, default=0, help='keep the dimensions with correlation > eps_corr')
    parser.add_argument('--eps_corr', type=float, default=0., help='minimum correlation')
    parser.add_argument('--metadata_dir', type=str, default='metadata')
    args, _ = parser.parse_known_args(argv)
    defaults.set_parameters(args)
    print('input args:\n', json.dumps(vars(args), indent=4, separators=(',', ':')))
    if args.mask_dims == 0:
        assert args.eps_corr == 0.
    return classify(config_name=args.config_name, seq_len=args.seq_len, n_trials=args.n_trials,
                    batch_size=args.batch_size, metadata_dir=args.metadata_dir)


if __name__ == '__main__':
    main()
