"""k-way n-shot classification driver (reference: config_rnn/test_few_shot_omniglot.py, same flags and prints).

An episode holds k candidate sequences: n shots of class j followed by the SAME query image.  The score of class j is
the predictive log-probability of the last frame, log_probs[j, -1]; scores of repeated presentations of a query are
averaged and the decision is the argmax (reference :64-80).
"""
import argparse
import importlib
import json
import os

import numpy as np
import torch

from .. import utils
from . import defaults

np.set_printoptions(suppress=True)


class _Query(object):
    """Accumulates the class scores of one query image over its presentations."""
    __slots__ = ('scores', 'truth', 'candidates')

    def __init__(self, truth, candidates):
        self.scores, self.truth, self.candidates = [], truth, candidates

    def correct(self):
        return self.candidates[int(np.argmax(np.mean(np.asarray(self.scores), axis=0)))] == self.truth


def _score_episode(config, x_batch, k, device):
    with torch.no_grad():
        log_probs = config.build_model(torch.from_numpy(x_batch).to(device), want_prior=False)[0]
    return log_probs.reshape(k, config.seq_len)[:, -1].cpu().numpy()


def classify(config_name, seq_len, n_trials, batch_size, metadata_dir='metadata', max_episodes=None, restore=True):
    config = importlib.import_module('bruno_b200.config_rnn.%s' % config_name)
    device = torch.device('cuda', 0)
    assert seq_len == config.seq_len
    if restore:
        save_dir = utils.find_model_metadata(metadata_dir + '/', config_name)
        print('Building the model', os.path.dirname(save_dir).split('/')[-1])
        utils.load_checkpoint_into(config, save_dir, device)
    episodes = config.test_data_iter2
    episodes.batch_size = batch_size
    accuracies = []
    for trial in range(n_trials):
        queries = {}
        for count, (x_batch, y_batch, query_id) in enumerate(episodes.generate(trial=trial)):
            if max_episodes is not None and count >= max_episodes:
                break
            if query_id not in queries:
                queries[query_id] = _Query(truth=int(y_batch[0, -1]), candidates=y_batch[:, 0])
            queries[query_id].scores.append(_score_episode(config, x_batch, episodes.batch_size, device))
        hits = sum(1 for q in queries.values() if q.correct())
        accuracies.append(hits / len(queries))
        print(trial, 'accuracy', accuracies[-1])
        print('n test examples', len(queries))
    print('---------------------------------------------')
    print(n_trials, config.seq_len)
    print(accuracies)
    print('average accuracy over trials', np.mean(accuracies))
    print('std accuracy over trials', np.std(accuracies))
    return accuracies


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
    if not args.mask_dims:
        assert args.eps_corr == 0., '--eps_corr needs --mask_dims 1'
    return classify(config_name=args.config_name, seq_len=args.seq_len, n_trials=args.n_trials,
                    batch_size=args.batch_size, metadata_dir=args.metadata_dir)


if __name__ == '__main__':
    main()
