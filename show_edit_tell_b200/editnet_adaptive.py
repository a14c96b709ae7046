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
        gd_final_hidden = None  # TODO(encoder entry point): second encoder pass on the GT caption
        return pred, call.caps, call.decode_lengths, call.sort_ind, gd_final_hidden, last_hidden

    def encode_final_hidden(self, seq, seq_len):
        """final_hidden of the caption encoder for arbitrary token rows (no grad): a zero-step
        rollout call runs exactly the encoder prologue."""
        raise NotImplementedError("gd_final_hidden (only consumed when use_mse=True, which the reference "
                                  "disables at editnet_adaptive.py:781) is not built yet")
