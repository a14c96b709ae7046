"""DCNet, self-critical stage: drop-in for `DAE` / `DAEWithAR` of /root/reference/dcnet_rl.py:273-361."""
import torch.nn as nn

from .dcnet import CaptionAttention, CaptionEncoder, DAEBase, Embedding  # noqa: F401
from .editnet_rl import RewardCriterion  # noqa: F401  (dcnet_rl.py:364-384 is the same criterion)


class DAE(DAEBase):
    def forward(self, word_map, encoded_previous_captions, previous_cap_length, sample_max, sample_rl):
        """-> (seq (B,18), seqLogprobs (B,18)); max_len = 18 hard-coded at dcnet_rl.py:295"""
        return self.rollout(word_map, encoded_previous_captions, previous_cap_length, sample_max, sample_rl, max_len=18)


class DAEWithAR(nn.Module):
    """dcnet_rl.py:348-361.  `DAEWithAR()` does what the reference constructor does: `torch.load` the cross-entropy
    checkpoint from `checkpoint` (the reference's fixed path by default) and wrap its 'dae' entry.  Alternatively the
    wrapped DAE is passed in, or built from a word map, so the class is usable without that file."""

    def __init__(self, dae=None, word_map=None, checkpoint='BEST_checkpoint_3_dae.pth.tar', **kw):
        super().__init__()
        if dae is None and word_map is None:
            import torch
            dae = torch.load(checkpoint, weights_only=False)['dae']          # dcnet_rl.py:355-356
        self.dae = dae if dae is not None else DAE(word_map, **kw)
        decoder_dim = self.dae.decoder_dim
        self.affine_hidden = nn.Linear(decoder_dim, decoder_dim)

    def forward(self, *args, **kwargs):
        return self.dae(*args, **kwargs)
