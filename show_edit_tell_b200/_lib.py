"""ctypes binding of libset_b200.so (the C ABI declared in include/set_b200.h).

The product path has no CPU fallback: if the library is missing or a call fails, a
RuntimeError carrying `set_last_error()` is raised.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SET_LIB_PATH") or os.path.join(_HERE, "libset_b200.so")   # override: kernel experiments
CSRC = os.path.join(_HERE, "csrc")


class SetDims(C.Structure):
    _fields_ = [("V", C.c_int), ("D", C.c_int), ("A", C.c_int), ("F", C.c_int)]


class SetSeqShape(C.Structure):
    _fields_ = [("B", C.c_int), ("R", C.c_int), ("Wc", C.c_int), ("Wp", C.c_int), ("P", C.c_int),
                ("T", C.c_int), ("train", C.c_int), ("adaptive", C.c_int)]


# field order mirrors SetEditNetParams; value = reference state_dict key
EDITNET_FIELDS = [
    ("embed", "embed.embedding.weight"),
    ("enc_x2h_w", "caption_encoder.lstm_encoder_cell.x2h.weight"),
    ("enc_x2h_b", "caption_encoder.lstm_encoder_cell.x2h.bias"),
    ("enc_h2h_w", "caption_encoder.lstm_encoder_cell.h2h.weight"),
    ("enc_h2h_b", "caption_encoder.lstm_encoder_cell.h2h.bias"),
    ("enc_aff_w", "caption_encoder.affine_hn.weight"),
    ("enc_aff_b", "caption_encoder.affine_hn.bias"),
    ("ca_feat_w", "caption_attention.cap_features_att.weight"),
    ("ca_feat_b", "caption_attention.cap_features_att.bias"),
    ("ca_dec_w", "caption_attention.cap_decoder_att.weight"),
    ("ca_dec_b", "caption_attention.cap_decoder_att.bias"),
    ("ca_full_w", "caption_attention.cap_full_att.weight"),
    ("ca_full_b", "caption_attention.cap_full_att.bias"),
    ("ca_gate_w", "caption_attention.context_gate.weight"),
    ("ca_gate_b", "caption_attention.context_gate.bias"),
    ("ca_sc_w", "caption_attention.sc_affine.weight"),
    ("ca_sc_b", "caption_attention.sc_affine.bias"),
    ("ca_tc_w", "caption_attention.tc_affine.weight"),
    ("ca_tc_b", "caption_attention.tc_affine.bias"),
    ("va_emb_w", "visual_attention.att_embed.0.weight"),
    ("va_emb_b", "visual_attention.att_embed.0.bias"),
    ("va_feat_w", "visual_attention.features_att.weight"),
    ("va_feat_b", "visual_attention.features_att.bias"),
    ("va_dec_w", "visual_attention.decoder_att.weight"),
    ("va_dec_b", "visual_attention.decoder_att.bias"),
    ("va_full_w", "visual_attention.full_att.weight"),
    ("va_full_b", "visual_attention.full_att.bias"),
    ("al_wih", "attention_lstm.weight_ih"),
    ("al_whh", "attention_lstm.weight_hh"),
    ("al_bih", "attention_lstm.bias_ih"),
    ("al_bhh", "attention_lstm.bias_hh"),
    ("cl_x2h_w", "copy_lstm.x2h.weight"),
    ("cl_x2h_b", "copy_lstm.x2h.bias"),
    ("cl_h2h_w", "copy_lstm.h2h.weight"),
    ("cl_h2h_b", "copy_lstm.h2h.bias"),
    ("cl_gcn_w", "copy_lstm.gate_cnew.weight"),
    ("cl_gcn_b", "copy_lstm.gate_cnew.bias"),
    ("cl_gcm_w", "copy_lstm.gate_cmem.weight"),
    ("cl_gcm_b", "copy_lstm.gate_cmem.bias"),
    ("fc_w", "fc.weight"),
    ("fc_b", "fc.bias"),
]


class SetEditNetParams(C.Structure):
    _fields_ = [(name, C.c_void_p) for name, _ in EDITNET_FIELDS]


DCNET_FIELDS = [
    ("embed", "embed.embedding.weight"),
    ("enc_wih_f", "caption_encoder.lstm_encoder.weight_ih_l0"),
    ("enc_whh_f", "caption_encoder.lstm_encoder.weight_hh_l0"),
    ("enc_bih_f", "caption_encoder.lstm_encoder.bias_ih_l0"),
    ("enc_bhh_f", "caption_encoder.lstm_encoder.bias_hh_l0"),
    ("enc_wih_r", "caption_encoder.lstm_encoder.weight_ih_l0_reverse"),
    ("enc_whh_r", "caption_encoder.lstm_encoder.weight_hh_l0_reverse"),
    ("enc_bih_r", "caption_encoder.lstm_encoder.bias_ih_l0_reverse"),
    ("enc_bhh_r", "caption_encoder.lstm_encoder.bias_hh_l0_reverse"),
    ("enc_cat_w", "caption_encoder.concat.weight"),
    ("enc_cat_b", "caption_encoder.concat.bias"),
    ("ca_feat_w", "caption_attention.cap_features_att.weight"),
    ("ca_feat_b", "caption_attention.cap_features_att.bias"),
    ("ca_dec_w", "caption_attention.cap_decoder_att.weight"),
    ("ca_dec_b", "caption_attention.cap_decoder_att.bias"),
    ("ca_full_w", "caption_attention.cap_full_att.weight"),
    ("ca_full_b", "caption_attention.cap_full_att.bias"),
    ("al_wih", "attention_lstm.weight_ih"),
    ("al_whh", "attention_lstm.weight_hh"),
    ("al_bih", "attention_lstm.bias_ih"),
    ("al_bhh", "attention_lstm.bias_hh"),
    ("ll_wih", "language_lstm.weight_ih"),
    ("ll_whh", "language_lstm.weight_hh"),
    ("ll_bih", "language_lstm.bias_ih"),
    ("ll_bhh", "language_lstm.bias_hh"),
    ("fc_w", "fc.weight"),
    ("fc_b", "fc.bias"),
]


class SetDcNetParams(C.Structure):
    _fields_ = [(name, C.c_void_p) for name, _ in DCNET_FIELDS]


def build(verbose=False):
    """Compile the CUDA sources in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:])
        print(out.stderr[-4000:])
    if out.returncode != 0:
        raise RuntimeError("building libset_b200.so failed")
    return LIB_PATH


_lib = None

_P = C.c_void_p
_SIGS = {
    "set_last_error": (C.c_char_p, []),
    "set_version": (C.c_int, []),
    "set_launch_count": (C.c_longlong, [C.c_int]),
    "set_profile_enable": (C.c_int, [C.c_int]),
    "set_profile_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "set_editnet_workspace_bytes": (C.c_size_t, [C.POINTER(SetDims), C.POINTER(SetSeqShape)]),
    "set_editnet_workspace_lookup": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), C.c_char_p,
                                               C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "set_editnet_encode": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), C.POINTER(SetEditNetParams), _P, _P,
                                     C.c_uint64, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "set_editnet_step_begin": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), C.POINTER(SetEditNetParams), _P, _P,
                                         _P, _P, _P, C.c_size_t, _P]),
    "set_editnet_step": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), C.POINTER(SetEditNetParams), _P, _P,
                                   C.c_int, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "set_editnet_xe_forward": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), C.POINTER(SetEditNetParams),
                                         _P, _P, _P, C.POINTER(C.c_int), _P, _P, C.c_uint64, _P, _P, C.c_size_t, _P]),
    "set_editnet_xe_forward_ss": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), C.POINTER(SetEditNetParams),
                                            _P, _P, _P, C.POINTER(C.c_int), _P, _P, C.c_uint64, C.c_float, _P, _P, _P,
                                            _P, C.c_size_t, _P]),
    "set_editnet_xe_backward": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), C.POINTER(SetEditNetParams),
                                          C.POINTER(SetEditNetParams), _P, _P, C.POINTER(C.c_int), _P, _P,
                                          C.c_uint64, _P, _P, C.c_size_t, _P]),
    "set_xe_loss": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_float, _P, _P, _P]),
    "set_editnet_xe_loss_time_major": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), _P, C.c_float, _P, _P,
                                                 C.c_size_t, _P]),
    "set_editnet_rollout": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), C.POINTER(SetEditNetParams),
                                      _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_int, _P, C.c_uint64, _P, _P, _P,
                                      C.c_size_t, _P]),
    "set_editnet_rollout_backward": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape),
                                               C.POINTER(SetEditNetParams), C.POINTER(SetEditNetParams), _P, _P, _P,
                                               C.c_uint64, _P, _P, C.c_size_t, _P]),
    "set_dcnet_workspace_bytes": (C.c_size_t, [C.POINTER(SetDims), C.POINTER(SetSeqShape)]),
    "set_dcnet_workspace_lookup": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), C.c_char_p,
                                             C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "set_dcnet_xe_forward": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), C.POINTER(SetDcNetParams), _P,
                                       C.POINTER(C.c_int), _P, _P, C.c_uint64, _P, _P, C.c_size_t, _P]),
    "set_dcnet_xe_backward": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), C.POINTER(SetDcNetParams),
                                        C.POINTER(SetDcNetParams), _P, C.POINTER(C.c_int), _P, _P, C.c_uint64, _P, _P,
                                        C.c_size_t, _P]),
    "set_dcnet_step_begin": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), C.POINTER(SetDcNetParams), _P, _P, _P,
                                       C.c_size_t, _P]),
    "set_dcnet_step": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), C.POINTER(SetDcNetParams), _P, C.c_int, _P,
                                 _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "set_dcnet_rollout": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), C.POINTER(SetDcNetParams), _P, _P,
                                    C.c_int64, C.c_int64, C.c_int, _P, C.c_uint64, _P, _P, _P, C.c_size_t, _P]),
    "set_dcnet_rollout_backward": (C.c_int, [C.POINTER(SetDims), C.POINTER(SetSeqShape), C.POINTER(SetDcNetParams),
                                             C.POINTER(SetDcNetParams), _P, _P, C.c_uint64, _P, _P, C.c_size_t, _P]),
    "set_ciderd_reward": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_int64, C.c_int64, C.c_int64, _P, _P,
                                    C.c_uint64, C.c_double, C.c_double, C.c_float, _P, _P, _P]),
    "set_reward_criterion": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, _P, _P]),
    "set_clip_adam": (C.c_int, [_P, _P, _P, _P, C.c_size_t, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                C.c_float, C.c_float, _P, _P, _P]),
    "set_dropout_keep_mask": (C.c_int, [_P, C.c_size_t, C.c_uint64, C.c_int, C.c_size_t, _P]),
    "set_gemm_backend": (C.c_int, [C.c_int]),
    "set_gemm_trace": (C.c_int, [_P]),
    "set_gemm_trace_seq": (C.c_int, [_P, C.c_long, C.c_int]),
    "set_gemm_stats": (C.c_int, [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_int]),
    "set_gemm_twin_launches": (C.c_longlong, [C.c_int]),
    "set_backward_bucket_events": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "set_backward_overwrite_grads": (C.c_int, [C.c_int]),
    "set_beam_expand": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, _P, _P, _P, _P, _P, _P, _P, _P, _P,
                                  _P, _P, _P, _P, _P]),
    "set_beam_gather": (C.c_int, [C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "set_beam_finalize": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "set_embed_forward": (C.c_int, [_P, C.c_long, _P, C.c_int, C.c_int, C.c_int, C.c_uint64, _P, _P]),
    "set_lstm_cell_forward": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "set_caption_attention_scratch_floats": (C.c_size_t, [C.POINTER(SetDims), C.c_int, C.c_int]),
    "set_caption_attention_forward": (C.c_int, [C.POINTER(SetDims), C.c_int, C.c_int, C.POINTER(SetEditNetParams), _P, _P, _P,
                                                _P, _P, C.c_size_t, _P, _P, _P]),
    "set_visual_attention_scratch_floats": (C.c_size_t, [C.POINTER(SetDims), C.c_int, C.c_int]),
    "set_visual_attention_forward": (C.c_int, [C.POINTER(SetDims), C.c_int, C.c_int, C.POINTER(SetEditNetParams), _P, _P,
                                               C.c_int, C.c_int, C.c_uint64, _P, C.c_size_t, _P, _P]),
    "set_dcnet_caption_attention_scratch_floats": (C.c_size_t, [C.POINTER(SetDims), C.c_int, C.c_int]),
    "set_dcnet_caption_attention_forward": (C.c_int, [C.POINTER(SetDims), C.c_int, C.c_int, C.POINTER(SetDcNetParams), _P, _P,
                                                      _P, _P, C.c_size_t, _P, _P]),
    "set_select_forward": (C.c_int, [C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "set_copy_lstm_scratch_floats": (C.c_size_t, [C.POINTER(SetDims), C.c_int]),
    "set_copy_lstm_forward": (C.c_int, [C.POINTER(SetDims), C.c_int, C.POINTER(SetEditNetParams), _P, _P, _P, _P, _P,
                                        C.c_size_t, _P, _P, _P]),
    "set_step_stats": (C.c_int, [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_int]),
    "set_step_trace": (C.c_int, [_P]),
    "set_step_geometry": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "set_gemm": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, _P, C.c_long, _P, C.c_long, _P, _P, C.c_long,
                           C.c_int, C.c_int, _P]),
}


def exported_symbols():
    return sorted(_SIGS)


def lib():
    """The loaded library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "libset_b200.so is not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the decode path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status):
    if status != 0:
        raise RuntimeError("libset_b200: " + lib().set_last_error().decode("utf-8", "replace"))


def ptr(t):
    """device pointer of a torch tensor (None -> NULL)"""
    return None if t is None else C.c_void_p(t.data_ptr())
