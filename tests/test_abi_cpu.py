"""CPU-side checks of the boundary: the C-ABI library loads without a GPU, exports every symbol
include/set_b200.h declares, sizes workspaces on the host, and the nn.Module shells expose the
reference's state_dict keys.  (No compute calls here -- those are the `-m gpu` tests.)"""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT, load_npz


def _lib():
    from show_edit_tell_b200 import _lib
    if not os.path.isfile(_lib.LIB_PATH):
        _lib.build()
    return _lib


def test_header_symbols_are_exported_and_bound():
    L = _lib()
    header = open(os.path.join(ROOT, "include", "set_b200.h")).read()
    declared = sorted(set(re.findall(r"SET_API[^;(]*?\b(set_[a-z0-9_]+)\s*\(", header)))
    assert declared, "no declarations parsed"
    lib = L.lib()
    for name in declared:
        assert hasattr(lib, name), "library does not export " + name
    assert declared == L.exported_symbols(), (declared, L.exported_symbols())
    assert lib.set_version() >= 100


def test_workspace_query_and_argument_errors():
    L = _lib()
    lib = L.lib()
    dims = L.SetDims(10000, 1024, 512, 2048)
    shape = L.SetSeqShape(64, 36, 20, 18, 18, 19, 1, 0)
    n_train = lib.set_editnet_workspace_bytes(C.byref(dims), C.byref(shape))
    shape.train = 0
    n_eval = lib.set_editnet_workspace_bytes(C.byref(dims), C.byref(shape))
    assert 0 < n_eval < n_train < 4 << 30
    off, nb = C.c_size_t(), C.c_size_t()
    assert lib.set_editnet_workspace_lookup(C.byref(dims), C.byref(shape), b"h2", C.byref(off), C.byref(nb)) == 0
    assert nb.value == 20 * 64 * 1024 * 4 and off.value % 256 == 0
    bad = L.SetDims(10000, 1023, 512, 2048)       # D not a multiple of 4
    assert lib.set_editnet_workspace_bytes(C.byref(bad), C.byref(shape)) == 0
    assert b"multiples of 4" in lib.set_last_error()
    with pytest.raises(RuntimeError):
        L.check(lib.set_editnet_workspace_lookup(C.byref(dims), C.byref(shape), b"nope", C.byref(off), C.byref(nb)))


def test_module_state_dict_keys_match_reference(small_sd, small_cfg):
    from show_edit_tell_b200 import editnet, editnet_adaptive, editnet_rl
    from oracle import synth
    c = small_cfg
    for cls in (editnet.DecoderC, editnet_rl.DecoderC, editnet_adaptive.DecoderC):
        mod = cls(synth.word_map(c["V"]), c["D"], c["D"], c["D"], c["A"], c["Fdim"])
        keys = set(mod.state_dict().keys())
        # the reference's state_dict also lists the embedding a second time under the encoder alias
        assert keys == set(small_sd.keys()) | {"caption_encoder.embed.embedding.weight"}
        for k, v in small_sd.items():
            assert tuple(mod.state_dict()[k].shape) == tuple(v.shape), k
        flat = mod.flatten_parameters()
        assert all(p.data_ptr() >= flat.data_ptr() for p in mod.parameters())
        assert mod.embed.embedding.weight.data_ptr() == mod.caption_encoder.embed.embedding.weight.data_ptr()


def test_forward_without_cuda_fails_loudly(small_cfg):
    from show_edit_tell_b200 import editnet
    from oracle import synth
    c = small_cfg
    mod = editnet.DecoderC(synth.word_map(c["V"]), c["D"], c["D"], c["D"], c["A"], c["Fdim"])
    b = synth.make_batch(2, c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], min_len=3, min_prev=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mod(b["feats"], b["caps"], b["caplens"], b["prev"], b["prev_len"], False, 0.0)
