"""B200 mirror of the reference's config_rnn/bn2_omniglot_tp_ft_1s_20w.py: episodic fine-tuning of bn2_omniglot_tp for
20-way 1-shot classification.  Same module attributes (batch_size = 20 candidates, seq_len = 2, meta_batch_size = 8
episodes per step, learning_rate 3e-5, scale_student_grad 1, base_metadata_path) and the same build_model / loss surface:
the model is the base model (bn2_omniglot_tp.py:59-221); the loss is a softmax over the candidates of the last element's
log-probability (ft:222-226), whose forward AND backward are one hand-written kernel (bruno_episode_softmax_loss)."""
import sys

import numpy as np

from .. import nn_extra_nvp, nn_extra_student, synthetic_data, utils, losses
from . import defaults, _common

try:                                                       # ft:11 -- the checkpoint of the base model that is fine-tuned
    base_metadata_path = utils.find_model_metadata('metadata/', 'bn2_omniglot_tp')
except ValueError:
    base_metadata_path = None                              # train_finetune.py reports the missing checkpoint

sample_batch_size = 1
n_samples = 4
rng = np.random.RandomState(42)
rng_test = np.random.RandomState(317070)
seq_len = defaults.seq_len_few_shot                        # 1-shot (2nd image is a test image), ft:17
batch_size = defaults.batch_size_few_shot                  # 20-way, ft:18
meta_batch_size = 8

nonlinearity = nn_extra_nvp.elu
weight_norm = True

_obs = (28, 28, 1)
train_data_iter = synthetic_data.SyntheticEpisodesIterator(seq_len=seq_len, batch_size=batch_size, meta_batch_size=meta_batch_size,
                                                           obs_shape=_obs, rng=rng)
test_data_iter2 = synthetic_data.SyntheticFewShotIterator(seq_len=seq_len, batch_size=batch_size, obs_shape=_obs, rng=rng_test)

obs_shape = train_data_iter.get_observation_size()         # (seq_len, 28, 28, 1)
print('obs shape', obs_shape)

ndim = int(np.prod(obs_shape[1:]))
corr_init = np.ones((ndim,), dtype='float32') * 0.1
nu_init = 1000

optimizer = 'rmsprop'
learning_rate = 0.00003
lr_decay = 0.999995
scale_student_grad = 1.
max_iter = 50000
save_every = 1000

nvp_layers = []
nvp_dense_layers = []
student_layer = None
model = None
has_conv = True
supports_mask_dims = False


def make_process_layer(device):
    return nn_extra_student.StudentRecurrentLayer(shape=(ndim,), corr_init=corr_init, nu_init=nu_init, device=device)


def build_model(x, init=False, sampling_mode=False, noise=None, want_prior=True):
    return _common.run_build_model(sys.modules[__name__], x, init, sampling_mode, noise=noise, want_prior=want_prior)


def build_nvp_model():
    _common.build_nvp_model(nvp_layers, nonlinearity, weight_norm)


def build_nvp_dense_model():
    _common.build_nvp_dense_model(nvp_dense_layers, nonlinearity, weight_norm, n_units=512)


def loss(log_probs):
    """ft:222-226: reshape to (meta_batch_size, batch_size, seq_len), last column, softmax over the candidates,
    -mean(log softmax[:, 0])."""
    return losses.episode_softmax_loss(log_probs, meta_batch_size, batch_size, seq_len)
