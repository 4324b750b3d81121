"""Stub of the reference's `data_iter` module (datasets are absent): only what a config module touches at IMPORT
time -- the constructors and `get_observation_size()` (data_iter.py:52-54: `(seq_len,) + image shape`)."""
_SHAPES = {'mnist': (28, 28, 1), 'fashion_mnist': (28, 28, 1), 'cifar10': (32, 32, 3), 'omniglot': (28, 28, 1)}


class _Iter(object):
    dataset_default = 'mnist'

    def __init__(self, seq_len, batch_size=None, dataset=None, set='train', rng=None, **kw):   # noqa: A002
        self.seq_len, self.batch_size, self.set, self.rng = seq_len, batch_size, set, rng
        self.img_shape = _SHAPES[dataset or self.dataset_default]

    def get_observation_size(self):
        return (self.seq_len,) + self.img_shape


class BaseExchSeqDataIterator(_Iter):
    pass


class BaseTestBatchSeqDataIterator(_Iter):
    pass


class OmniglotExchSeqDataIterator(_Iter):
    dataset_default = 'omniglot'


class OmniglotTestBatchSeqDataIterator(_Iter):
    dataset_default = 'omniglot'


class OmniglotEpisodesDataIterator(_Iter):
    dataset_default = 'omniglot'
