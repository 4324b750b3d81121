"""Loss functions of the config modules whose backward is hand-written CUDA (no autograd of torch ops on the path)."""
import torch

from . import _lib


class _EpisodeSoftmaxLoss(torch.autograd.Function):
    """loss of config_rnn/bn2_omniglot_tp_ft_1s_20w.py:222-226 -- one kernel computes the loss and d loss / d log_probs."""

    @staticmethod
    def forward(ctx, log_probs, meta_batch_size, batch_size, seq_len):
        lp = log_probs.contiguous()
        if lp.numel() != meta_batch_size * batch_size * seq_len:
            raise ValueError('log_probs has %d elements, expected meta_batch_size * batch_size * seq_len = %d' % (
                lp.numel(), meta_batch_size * batch_size * seq_len))
        loss = torch.empty(1, dtype=torch.float32, device=lp.device)
        need = ctx.needs_input_grad[0]
        dlp = torch.empty_like(lp) if need else None
        _lib.call('bruno_episode_softmax_loss', lp, meta_batch_size, batch_size, seq_len, loss, dlp)
        if need:
            ctx.save_for_backward(dlp)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        dlp, = ctx.saved_tensors
        return dlp * g, None, None, None


def episode_softmax_loss(log_probs, meta_batch_size, batch_size, seq_len):
    """-mean over episodes of log softmax over the `batch_size` candidates of log_probs[:, -1]; candidate 0 is the true class."""
    return _EpisodeSoftmaxLoss.apply(log_probs, int(meta_batch_size), int(batch_size), int(seq_len))
