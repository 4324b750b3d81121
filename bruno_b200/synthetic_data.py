"""Synthetic stand-ins for the reference's dataset iterators (data_iter.py), which need data/*.npy files that do not
exist here.  Same shape / range contract as data_iter.py:61-74: x = pixel + U[0,1) with pixel in {0..255}, float32,
shape (batch_size, seq_len) + obs_shape; `generate()` yields batches, `get_observation_size()` returns
(seq_len,) + obs_shape.  'strokes' mimics Omniglot (sparse saturated strokes), 'bytes' is uniform noise."""
import numpy as np


class SyntheticExchSeqDataIterator(object):
    def __init__(self, seq_len, batch_size, obs_shape=(28, 28, 1), kind='strokes', rng=None, noise_rng=None,
                 n_batches=None, dataset='train', **kwargs):
        self.seq_len = seq_len
        self.batch_size = batch_size
        self.obs_shape = tuple(obs_shape)
        self.kind = kind
        self.rng = rng if rng is not None else np.random.RandomState(42)
        self.noise_rng = noise_rng if noise_rng is not None else np.random.RandomState(317070)
        self.nsamples = n_batches
        self.set = dataset

    def get_observation_size(self):
        return (self.seq_len,) + self.obs_shape

    def _batch(self, rng=None, noise_rng=None):
        rng = self.rng if rng is None else rng
        noise_rng = self.noise_rng if noise_rng is None else noise_rng
        shape = (self.batch_size, self.seq_len) + self.obs_shape
        if self.kind == 'strokes':
            # one "class" prototype per sequence + per-frame jitter: frames of a sequence are exchangeable, not iid
            proto = rng.uniform(size=(self.batch_size, 1) + self.obs_shape) < 0.1
            flip = rng.uniform(size=shape) < 0.03
            pix = np.logical_xor(np.broadcast_to(proto, shape), flip).astype(np.float32) * 255.0
        else:
            pix = rng.randint(0, 256, size=shape).astype(np.float32)
        noise = noise_rng.uniform(size=shape).astype(np.float32)
        return np.minimum(pix + noise, np.float32(255.99998))

    def generate(self, rng=None, noise_rng=None):
        """data_iter.py:56-58: an explicit `rng` replaces the iterator's own (train.py:204 re-seeds it for validation; the
        dequantisation noise then comes from the same stream, like the reference's noise_rng = self.rng default)."""
        if rng is not None and noise_rng is None:
            noise_rng = rng
        i = 0
        while self.nsamples is None or i < self.nsamples:
            yield self._batch(rng, noise_rng)
            i += 1

    def generate_each_digit(self, **kwargs):
        for x in self.generate():
            yield x, np.zeros((self.batch_size,), dtype='int32')

    def _class_sequence(self, rng, proto=None, same_image=False):
        """seq_len frames of one synthetic class (prototype + per-frame flips), pixels in {0, 255}, no noise yet."""
        shape = (self.seq_len,) + self.obs_shape
        if self.kind == 'strokes':
            proto = (rng.uniform(size=self.obs_shape) < 0.1) if proto is None else proto
            flip = rng.uniform(size=shape) < 0.03
            if same_image:
                flip = np.broadcast_to(flip[:1], shape)
            return np.logical_xor(np.broadcast_to(proto, shape), flip).astype(np.float32) * 255.0
        seq = rng.randint(0, 256, size=shape).astype(np.float32)
        return np.broadcast_to(seq[:1], shape).copy() if same_image else seq

    def generate_anomaly(self, noise_rng=None):
        """Contract of data_iter.py:398-422: ONE sequence of a class whose element seq_len - 5 is replaced by an image of
        another class; y marks the class of every element.  Yields (x [1, seq_len, ...], y [1, seq_len])."""
        noise_rng = np.random.RandomState(42) if noise_rng is None else noise_rng
        while True:
            x = self._class_sequence(self.rng)[None].copy()
            y = np.zeros((1, self.seq_len), dtype=np.float32)
            x[0, self.seq_len - 5] = self._class_sequence(self.rng)[0]
            y[0, self.seq_len - 5] = 1.0
            yield np.minimum(x + noise_rng.uniform(size=x.shape).astype(np.float32), np.float32(255.99998)), y

    def generate_diagonal_roll(self, same_class=True, same_image=False, black_image=False, rng=None, noise_rng=None):
        """Contract of data_iter.py:424-463: a batch of seq_len copies of ONE sequence, copy i rolled by i positions, so
        that the diagonal of a [seq_len, seq_len] output follows frame 0 through every position of the sequence."""
        rng = self.rng if rng is None else rng
        noise_rng = self.rng if noise_rng is None else noise_rng
        while True:
            seq = self._class_sequence(rng, same_image=same_image)[None].copy()
            if black_image:
                seq[0, 0] *= 0.
            if black_image and same_image:
                seq *= 0.
            if not same_class:
                seq[0, 0] = self._class_sequence(rng)[0]
            seq = np.minimum(seq + noise_rng.uniform(size=seq.shape).astype(np.float32), np.float32(255.99998))
            yield np.concatenate([np.roll(seq, i, axis=1) for i in range(self.seq_len)], 0)


class SyntheticFewShotIterator(object):
    """Shape contract of OmniglotTestBatchSeqDataIterator.generate (data_iter.py:204-241): k-way n-shot episodes as
    x [k, n+1, ...] where row j holds n shots of class j followed by THE SAME test image in every row, plus the label
    index of the true class."""

    def __init__(self, seq_len, batch_size, obs_shape=(28, 28, 1), rng=None, n_classes=50, **kwargs):
        self.seq_len, self.batch_size, self.obs_shape = seq_len, batch_size, tuple(obs_shape)
        self.rng = rng if rng is not None else np.random.RandomState(317070)
        self.n_classes = n_classes
        self.protos = (self.rng.uniform(size=(n_classes,) + self.obs_shape) < 0.12)

    def get_observation_size(self):
        return (self.seq_len,) + self.obs_shape

    def _draw(self, c):
        flip = self.rng.uniform(size=self.obs_shape) < 0.02
        return np.logical_xor(self.protos[c], flip).astype(np.float32) * 255.0

    def generate(self, trial=0, condition_on_train=False):
        k, n = self.batch_size, self.seq_len - 1
        for true_cls in range(self.n_classes):
            others = [c for c in range(self.n_classes) if c != true_cls]
            classes = [true_cls] + list(self.rng.choice(others, k - 1, replace=False))
            self.rng.shuffle(classes)
            x = np.zeros((k, self.seq_len) + self.obs_shape, dtype=np.float32)
            for j, c in enumerate(classes):
                for s in range(n):
                    x[j, s] = self._draw(c)
            test = self._draw(true_cls)
            x[:, -1] = test[None]
            noise = self.rng.uniform(size=(1,) + x.shape[1:]).astype(np.float32)   # same noise in every row (data_iter.py:233-238)
            x = np.minimum(x + noise, np.float32(255.99998))
            # (x_batch, y_batch [k, seq_len] with the true label in the last column, episode id): data_iter.py:204-241
            y = np.repeat(np.asarray(classes)[:, None], self.seq_len, axis=1)
            y[:, -1] = true_cls
            yield x, y, true_cls


class SyntheticEpisodesIterator(SyntheticFewShotIterator):
    """Shape contract of OmniglotEpisodesDataIterator.generate (data_iter.py:244-296), the training stream of the episodic
    fine-tuning configs: an endless stream of meta-batches of `meta_batch_size` k-way n-shot episodes,
    x [meta_batch_size * k, n + 1, ...]; inside an episode row 0 is the TRUE class (n shots + the test image) and rows
    1..k-1 are other classes whose last element is overwritten with the same test image (:282); one noise sequence per
    episode is added to every row (:284-286)."""

    def __init__(self, seq_len, batch_size, meta_batch_size, obs_shape=(28, 28, 1), rng=None, n_classes=50, **kwargs):
        super(SyntheticEpisodesIterator, self).__init__(seq_len, batch_size, obs_shape=obs_shape, rng=rng, n_classes=n_classes)
        self.meta_batch_size = meta_batch_size

    def generate(self, trial=0, rng=None):
        rng = self.rng if rng is None else rng
        k, T = self.batch_size, self.seq_len
        while True:
            x = np.zeros((self.meta_batch_size, k, T) + self.obs_shape, dtype=np.float32)
            y = np.zeros((self.meta_batch_size, k, T), dtype=np.float32)
            for m in range(self.meta_batch_size):
                true_cls = int(rng.choice(self.n_classes))
                others = [c for c in range(self.n_classes) if c != true_cls]
                classes = [true_cls] + [int(c) for c in rng.choice(others, k - 1, replace=False)]
                for j, c in enumerate(classes):
                    for s in range(T):
                        x[m, j, s] = self._draw(c)
                    y[m, j] = c
                x[m, 1:, -1] = x[m, 0, -1][None]                       # the same test image closes every candidate row
                noise = rng.uniform(size=(1, T) + self.obs_shape).astype(np.float32)
                x[m] = np.minimum(x[m] + noise, np.float32(255.99998))
            yield (x.reshape((self.meta_batch_size * k, T) + self.obs_shape), y.reshape(self.meta_batch_size * k, T))


class SyntheticShapenetIterator(object):
    """Shape / range contract of the reference's ShapenetConditionalNPDataIterator (utils_conditional.py:100-160) for
    Conditional BRUNO: x [batch, seq_len, 32, 32, 1] = pixel + U[0,1), y_label [batch, seq_len, 2] = (sin, cos) of a view
    angle in {0, 10, ..., 350} degrees.  One object prototype per sequence, shifted horizontally with the angle: the
    frames of a sequence are exchangeable given their labels."""
    def __init__(self, seq_len, batch_size, obs_shape=(32, 32, 1), rng=None, noise_rng=None, n_batches=None, **kwargs):
        self.seq_len, self.batch_size, self.obs_shape = seq_len, batch_size, tuple(obs_shape)
        self.rng = rng if rng is not None else np.random.RandomState(42)
        self.noise_rng = noise_rng if noise_rng is not None else np.random.RandomState(317070)
        self.nsamples = n_batches

    def get_observation_size(self):
        return (self.seq_len,) + self.obs_shape

    def get_label_size(self):
        return (self.seq_len, 2)

    def _batch(self):
        B, T = self.batch_size, self.seq_len
        H, W, C = self.obs_shape
        proto = (self.rng.uniform(size=(B, H, W, C)) < 0.15).astype(np.float32) * 255.0
        angles = self.rng.randint(0, 36, size=(B, T)) * 10.0
        x = np.empty((B, T, H, W, C), dtype=np.float32)
        for b in range(B):
            for t in range(T):
                x[b, t] = np.roll(proto[b], int(angles[b, t] / 360.0 * W), axis=1)
        x = np.minimum(x + self.noise_rng.uniform(size=x.shape).astype(np.float32), np.float32(255.99998))
        rad = np.deg2rad(angles)
        y = np.stack([np.sin(rad), np.cos(rad)], -1).astype(np.float32)
        return x, y

    def generate(self, **kwargs):
        i = 0
        while self.nsamples is None or i < self.nsamples:
            yield self._batch()
            i += 1
