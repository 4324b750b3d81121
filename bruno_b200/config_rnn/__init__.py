"""Mirrors of the reference's config_rnn modules (model builders and train / test drivers) on the B200 kernels."""
