"""EditNet over ragged 10-100 region features: drop-in for `DecoderC` of
/root/reference/adaptive_features/editnet_adaptive.py:459-562 (masked VisualAttentionC :423-457).
All-zero region rows are padding; `image_mean` is an input; the forward returns the reference's
6-tuple (adds gd_final_hidden, decoder_last_hidden, used by the optional MSE term :593-595)."""
import torch

from .editnet import EditNetBase


class DecoderC(EditNetBase):
    ADAPTIVE = True

    def forward(self, image_features, image_mean, encoded_captions, caption_lengths, encoded_previous_captions,
                previous_cap_length, use_ss=False, ss_prob=0.0):
        pred, call = self._xe(image_features, image_mean, encoded_captions, caption_lengths,
                              encoded_previous_captions, previous_cap_length, use_ss, ss_prob)
        # decoder_last_hidden[i] = h2 of row i at its last decoded step (editnet_adaptive.py:560)
        B, D = call.shape.B, self.decoder_dim
        h2 = self.workspace_tensor("h2").view(call.shape.T + 1, B, D)
        idx = torch.tensor(call.decode_lengths, device=h2.device)
        last_hidden = h2[idx, torch.arange(B, device=h2.device)].clone()
        # gd_final_hidden: encoder run on the ground-truth caption (:516) -- a second encoder pass
        # gd_final_hidden: the encoder run on the ground-truth caption (:516).  No autograd edge: it only
        # matters for the optional MSE term, which the reference disables (use_mse=False, :781).
        with torch.no_grad():
            gd_final_hidden = self.encode(call.caps, idx + 1)[2]
        return pred, call.caps, call.decode_lengths, call.sort_ind, gd_final_hidden, last_hidden
