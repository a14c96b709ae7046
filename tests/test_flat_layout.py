"""Host logic of the flat parameter buffer (no GPU): the gradient buckets of the data-parallel step are contiguous
ranges in the order the reverse pass finishes them (EditNetBase.BUCKET_FIELDS; csrc/editnet.cu backward_core), every
parameter lies in exactly one of them, and flattening keeps parameter values and state_dict keys."""
import torch

from oracle import synth
from show_edit_tell_b200 import _lib, dcnet, editnet


def _decoder():
    torch.manual_seed(0)
    return editnet.DecoderC(synth.word_map(50), 32, 32, 32, 16, 64)


def test_bucket_fields_name_real_parameters_once():
    names = [n for n, _ in _lib.EDITNET_FIELDS]
    listed = [n for b in editnet.EditNetBase.BUCKET_FIELDS for n in b]
    assert len(listed) == len(set(listed)), "a field sits in two buckets"
    assert set(listed) <= set(names), set(listed) - set(names)
    assert len(listed) < len(names), "the last bucket (everything not listed) must not be empty"


def test_flat_layout_is_one_contiguous_range_per_bucket():
    dec = _decoder()
    before = {k: v.detach().clone() for k, v in dec.state_dict().items()}
    flat = dec.flatten_parameters()
    after = dec.state_dict()
    assert list(before) == list(after)
    for k in before:
        assert torch.equal(before[k], after[k]), k
    names = [n for n, _ in dec.FIELDS]
    params = dec._ordered_params()
    starts = dec._bucket_offsets
    assert starts[0] == 0 and starts == sorted(starts) and len(starts) == len(dec.BUCKET_FIELDS) + 1
    ends = starts[1:] + [flat.numel()]
    bucket_of = {}
    for b, fields in enumerate(dec.BUCKET_FIELDS):
        for n in fields:
            bucket_of[n] = b
    spans = []
    for n, p, o in zip(names, params, dec._offsets):
        b = bucket_of.get(n, len(dec.BUCKET_FIELDS))
        assert starts[b] <= o and o + p.numel() <= ends[b], (n, b, o, starts[b], ends[b])
        assert o % 64 == 0, "parameters start on 256-byte boundaries (128-bit accesses, TMA)"
        assert p.data_ptr() == flat.data_ptr() + 4 * o
        spans.append((o, o + p.numel()))
    spans.sort()
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0, "parameters overlap in the flat buffer"
    # flattening twice is a no-op
    assert dec.flatten_parameters().data_ptr() == flat.data_ptr()


def test_dcnet_has_a_single_bucket():
    torch.manual_seed(0)
    dae = dcnet.DAE(synth.word_map(50), None, decoder_dim=32, attention_dim=16, caption_features_dim=16, emb_dim=32)
    dae.flatten_parameters()
    assert dae._bucket_offsets == [0]
