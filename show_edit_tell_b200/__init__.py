"""B200-native EditNet/DCNet decode path of show-edit-tell.

Layout: `csrc/` (CUDA kernels + the C ABI of include/set_b200.h, built into
libset_b200.so), `_lib.py` (ctypes binding), `editnet.py` / `editnet_rl.py` /
`editnet_adaptive.py` (the reference's nn.Module surface), `train.py` (fused train
steps + data-parallel wrapper), `synth.py` (synthetic batches for bench/smoke).
"""
from . import _lib  # noqa: F401
from .editnet import (CaptionAttentionC, CaptionEncoderC, CopyLSTMCellC, DecoderC, EditNetBase,  # noqa: F401
                      EmbeddingC, LSTMCellC, SelectC, VisualAttentionC)

__all__ = ["DecoderC", "EditNetBase", "LSTMCellC", "CopyLSTMCellC", "EmbeddingC", "CaptionEncoderC",
           "CaptionAttentionC", "SelectC", "VisualAttentionC"]
