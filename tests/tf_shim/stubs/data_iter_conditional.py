"""Stub of the reference's `data_iter_conditional` (ShapeNet views absent): import-time surface only
(data_iter_conditional.py: observations (seq_len, 32, 32, 1), labels (seq_len, 2) = (sin, cos) of the view angle)."""


class ShapenetConditionalNPDataIterator(object):
    def __init__(self, seq_len, batch_size, set='train', rng=None, **kw):   # noqa: A002
        self.seq_len, self.batch_size = seq_len, batch_size

    def get_observation_size(self):
        return (self.seq_len, 32, 32, 1)

    def get_label_size(self):
        return (self.seq_len, 2)
