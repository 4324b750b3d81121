"""CPU: shape / range / structure contracts of the synthetic stand-ins for the reference's dataset iterators."""
import numpy as np

from bruno_b200.synthetic_data import SyntheticEpisodesIterator, SyntheticExchSeqDataIterator, SyntheticFewShotIterator


def test_exchangeable_sequence_iterator_contract_and_reseeding():
    """data_iter.py:52-77: (batch, seq_len) + obs_shape float32 in [0, 256); generate(rng) replays the same batches
    (config_rnn/train.py:204 re-seeds RandomState(42) for every validation)."""
    it = SyntheticExchSeqDataIterator(seq_len=5, batch_size=3, obs_shape=(28, 28, 1))
    assert it.get_observation_size() == (5, 28, 28, 1)
    a = next(it.generate(np.random.RandomState(42)))
    b = next(it.generate(np.random.RandomState(42)))
    c = next(it.generate())
    assert a.shape == (3, 5, 28, 28, 1) and a.dtype == np.float32
    assert 0.0 <= a.min() and a.max() < 256.0
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_few_shot_iterator_contract():
    """data_iter.py:204-241: every candidate row ends with THE SAME test image; labels carry the true class last."""
    it = SyntheticFewShotIterator(seq_len=2, batch_size=20, rng=np.random.RandomState(3), n_classes=30)
    x, y, cls = next(it.generate())
    assert x.shape == (20, 2, 28, 28, 1) and y.shape == (20, 2)
    assert np.abs(x[:, -1] - x[:1, -1]).max() == 0.0
    assert np.all(y[:, -1] == cls) and cls in y[:, 0]


def test_episodes_iterator_contract():
    """data_iter.py:244-296: meta_batch_size episodes of batch_size candidates, flattened; row 0 of an episode is the true
    class and its last image closes every row; one noise sequence per episode."""
    it = SyntheticEpisodesIterator(seq_len=2, batch_size=20, meta_batch_size=8, rng=np.random.RandomState(1))
    x, y = next(it.generate())
    assert x.shape == (160, 2, 28, 28, 1) and y.shape == (160, 2) and x.dtype == np.float32
    assert 0.0 <= x.min() and x.max() < 256.0
    xe, ye = x.reshape(8, 20, 2, -1), y.reshape(8, 20, 2)
    assert np.abs(xe[:, 1:, -1] - xe[:, :1, -1]).max() == 0.0
    for m in range(8):
        assert len(set(ye[m, :, 0])) == 20                  # 20 distinct classes per episode
    x2, _ = next(it.generate())
    assert not np.array_equal(x, x2)
