"""EditNet, self-critical stage: drop-in for `DecoderC` of /root/reference/editnet_rl.py
(forward signature :485, RewardCriterion :553-573).  Same parameters and `state_dict` keys as
the XE-stage class; only `forward` differs, exactly as in the reference."""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import check, ptr
from .editnet import (CaptionAttentionC, CaptionEncoderC, CopyLSTMCellC, EditNetBase, EmbeddingC,  # noqa: F401
                      LSTMCellC, SelectC, VisualAttentionC, _stream)


class DecoderC(EditNetBase):
    def forward(self, word_map, encoded_previous_captions, previous_cap_length, image_features, sample_max,
                sample_rl):
        """-> (seq (B,18) int64, seqLogprobs (B,18)); max_len = 18 is hard-coded at editnet_rl.py:487"""
        return self.rollout(word_map, encoded_previous_captions, previous_cap_length, image_features, sample_max,
                            sample_rl, max_len=18)


class _RewardFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, slp, seq, reward):
        B, T = slp.shape
        out = torch.empty(2, device=slp.device, dtype=torch.float32)
        dlp = torch.empty_like(slp)
        check(_lib.lib().set_reward_criterion(B, T, ptr(slp.contiguous()), ptr(seq.contiguous()),
                                              ptr(reward.contiguous().float()), ptr(out), ptr(dlp), _stream()))
        ctx.save_for_backward(dlp)
        return out[0]

    @staticmethod
    def backward(ctx, g):
        (dlp,) = ctx.saved_tensors
        return dlp * g, None, None


class RewardCriterion(nn.Module):
    """editnet_rl.py:553-573"""

    def forward(self, sample_logprobs, seq, reward):
        return _RewardFn.apply(sample_logprobs, seq, reward)
