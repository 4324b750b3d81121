"""B200 mirror of the reference's config_rnn/en2_noconv_mnist_tp.py: same module attributes and the same
build_model / loss surface; the graph runs on hand-written sm_100a kernels."""
import sys

import numpy as np

from .. import nn_extra_nvp, nn_extra_student, synthetic_data
from . import defaults, _common

batch_size = 64
sample_batch_size = 1
n_samples = 4
rng = np.random.RandomState(42)
rng_test = np.random.RandomState(317070)
seq_len = defaults.seq_len

nonlinearity = nn_extra_nvp.elu
weight_norm = True

# The reference builds dataset iterators here (config_rnn/en2_noconv_mnist_tp.py:19-27); the datasets are not available, so synthetic
# iterators with the same shape / range contract stand in (bruno_b200/synthetic_data.py).
_obs = (28, 28, 1)
train_data_iter = synthetic_data.SyntheticExchSeqDataIterator(seq_len=seq_len, batch_size=batch_size, obs_shape=_obs,
                                                              kind='strokes', rng=rng, dataset='train')
test_data_iter = synthetic_data.SyntheticExchSeqDataIterator(seq_len=seq_len, batch_size=batch_size, obs_shape=_obs,
                                                             kind='strokes', rng=rng_test, dataset='test')
valid_data_iter = test_data_iter
test_data_iter2 = synthetic_data.SyntheticFewShotIterator(seq_len=seq_len, batch_size=batch_size, obs_shape=_obs,
                                                          rng=rng_test)

obs_shape = train_data_iter.get_observation_size()  # (seq_len, H, W, C)
print('obs shape', obs_shape)

ndim = int(np.prod(obs_shape[1:]))
corr_init = np.ones((ndim,), dtype='float32') * 0.1
nu_init = 1000

optimizer = 'rmsprop'
learning_rate = 0.001
lr_decay = 0.999995

scale_student_grad = 0.
max_iter = 100000
save_every = 1000
validate_every = 1000
n_valid_batches = 20
student_grad_schedule = {0: 0., 100: 0.1}

nvp_layers = []
nvp_dense_layers = []
student_layer = None
model = None
has_conv = False
supports_mask_dims = True


def make_process_layer(device):
    return nn_extra_student.StudentRecurrentLayer(shape=(ndim,), corr_init=corr_init, nu_init=nu_init, device=device)


def build_model(x, init=False, sampling_mode=False, noise=None, want_prior=True):
    """config_rnn/en2_noconv_mnist_tp.py build_model: returns (log_probs, latent_log_probs, latent_log_probs_prior, z_vec), or
    (x_samples, log_probs) when sampling_mode."""
    return _common.run_build_model(sys.modules[__name__], x, init, sampling_mode, noise=noise, want_prior=want_prior)


def build_nvp_model():
    _common.build_nvp_model(nvp_layers, nonlinearity, weight_norm)


def build_nvp_dense_model():
    _common.build_nvp_dense_model(nvp_dense_layers, nonlinearity, weight_norm, n_units=512)


def loss(log_probs):
    return -log_probs.mean()
