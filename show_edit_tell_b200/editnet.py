"""EditNet host side: the reference's module surface over the CUDA C ABI.

Class names, constructor signatures, sub-module attribute names and `state_dict` keys
are those of /root/reference/editnet.py:210-477 (and `eval/eval xe/editnet.py`, the
class-only copy checkpoints are unpickled against), so `DecoderC` drops into the
reference's `train()` / checkpoint code.  The arithmetic of `forward` runs entirely
in libset_b200.so (hand-written sm_100a kernels); torch only owns memory, streams and the
autograd edge.  There is no CPU path: calling `forward` without the built library or
without a CUDA device raises.
"""
import ctypes as C
import math
import weakref

import torch
import torch.nn as nn

from . import _lib
from ._lib import EDITNET_FIELDS, SetDims, SetEditNetParams, SetSeqShape, check, ptr


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _draw_seed():
    # one 62-bit seed per call from torch's CPU generator -> torch.manual_seed governs dropout
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


# --------------------------------------------------------------------- sub-modules
# The reference's beam searches call the decoder's sub-modules one by one (evaluate(), editnet.py:613,645-653;
# eval/eval xe/eval_full.py:107-149), so each of them has a `forward` of its own: one library call computing exactly
# what the reference module computes (nothing hoisted -- the caller owns the loop).  Inference surface: the results
# carry no autograd graph; training goes through DecoderC.forward / the trainers.
def _c(t):
    return t.contiguous().float()


class _Owned:
    """sub-modules that need the decoder's whole parameter struct keep a weak reference to it"""

    def _dec(self):
        ref = getattr(self, "_owner", None)
        dec = ref() if ref is not None else None
        if dec is None:
            raise RuntimeError("this sub-module is not attached to a decoder (construct it through DecoderC)")
        dec._require_cuda(dec.fc.weight)
        dec.flatten_parameters()
        return dec

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_owner", None)
        return state


class LSTMCellK(nn.LSTMCell):
    """nn.LSTMCell whose forward runs on the library's GEMM + cell kernels (attention_lstm, editnet.py:468,532)"""

    def forward(self, input, hx=None):
        if not input.is_cuda:
            raise RuntimeError("show_edit_tell_b200 runs on a CUDA device only (no CPU fallback)")
        rows = input.shape[0]
        Dh = self.hidden_size
        if hx is None:
            z = torch.zeros(rows, Dh, device=input.device)
            hx = (z, z)
        x, h, c = _c(input), _c(hx[0]), _c(hx[1])
        gates = torch.empty(rows, 4 * Dh, device=x.device)
        h_out, c_out = torch.empty_like(h), torch.empty_like(c)
        check(_lib.lib().set_lstm_cell_forward(rows, self.input_size, Dh, ptr(x), ptr(h), ptr(c), ptr(self.weight_ih),
                                               ptr(self.weight_hh), ptr(self.bias_ih), ptr(self.bias_hh), ptr(gates),
                                               ptr(h_out), ptr(c_out), _stream()))
        return h_out, c_out


class LinearK(nn.Linear):
    """nn.Linear whose forward runs on the library's GEMM engine (fc, editnet.py:471,653)"""

    def forward(self, input):
        if not input.is_cuda:
            raise RuntimeError("show_edit_tell_b200 runs on a CUDA device only (no CPU fallback)")
        x = _c(input).view(-1, self.in_features)
        out = torch.empty(x.shape[0], self.out_features, device=x.device)
        check(_lib.lib().set_gemm(0, x.shape[0], self.out_features, self.in_features, ptr(x), self.in_features,
                                  ptr(self.weight), self.in_features, ptr(self.bias), ptr(out), self.out_features, 0, 0,
                                  _stream()))
        return out.view(*input.shape[:-1], self.out_features)


class LSTMCellC(nn.Module):
    """Parameter container of the encoder cell (editnet.py:210-244)."""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.hidden_size = hidden_size
        self.input_size = input_size
        self.x2h = nn.Linear(input_size, 4 * hidden_size)
        self.h2h = nn.Linear(hidden_size, 4 * hidden_size)
        self.tanh = nn.Tanh()
        self.init_parameters()

    def init_parameters(self):
        std = 1.0 / math.sqrt(self.hidden_size)
        for p in self.parameters():
            p.data.uniform_(-std, std)


class CopyLSTMCellC(_Owned, nn.Module):
    """The copy-LSTM (editnet.py:247-285); forward(x, (h, c), c_mem) -> (h, c) as :265-285."""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.hidden_size = hidden_size
        self.input_size = input_size
        self.x2h = nn.Linear(input_size, 4 * hidden_size)
        self.h2h = nn.Linear(hidden_size, 4 * hidden_size)
        self.gate_cnew = nn.Linear(hidden_size, hidden_size)
        self.gate_cmem = nn.Linear(hidden_size, hidden_size)
        self.tanh = nn.Tanh()
        self.init_parameters()

    def init_parameters(self):
        std = 1.0 / math.sqrt(self.hidden_size)
        for p in self.parameters():
            p.data.uniform_(-std, std)

    def forward(self, x, states, c_mem):
        dec = self._dec()
        rows = x.shape[0]
        x, h, c, mem = _c(x), _c(states[0]), _c(states[1]), _c(c_mem)
        dims = dec._dims()
        n = _lib.lib().set_copy_lstm_scratch_floats(C.byref(dims), rows)
        scratch = torch.empty(n, device=x.device)
        h_out, c_out = torch.empty_like(h), torch.empty_like(c)
        check(_lib.lib().set_copy_lstm_forward(C.byref(dims), rows, C.byref(dec._struct), ptr(x), ptr(h), ptr(c), ptr(mem),
                                               ptr(scratch), n, ptr(h_out), ptr(c_out), _stream()))
        return h_out, c_out


class EmbeddingC(nn.Module):
    """editnet.py:288-304"""

    def __init__(self, word_map, emb_dim):
        super().__init__()
        self.emb_dim = emb_dim
        self.word_map = word_map
        self.embedding = nn.Embedding(len(word_map), self.emb_dim)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(0.5)

    def forward(self, x):
        """dropout(relu(Emb[x])), editnet.py:300-304; x int64 of any shape -> x.shape + (emb_dim,)"""
        if not x.is_cuda:
            raise RuntimeError("show_edit_tell_b200 runs on a CUDA device only (no CPU fallback)")
        tok = x.contiguous().view(-1)
        table = self.embedding.weight
        out = torch.empty(tok.numel(), self.emb_dim, device=x.device)
        seed = _draw_seed() if self.training else 0
        check(_lib.lib().set_embed_forward(ptr(tok), tok.numel(), ptr(table), table.shape[0], self.emb_dim,
                                           int(self.training), seed, ptr(out), _stream()))
        return out.view(*x.shape, self.emb_dim)


class CaptionEncoderC(_Owned, nn.Module):
    """editnet.py:307-348; forward(seq, seq_len) -> (hidden_states, memory_states, final_hidden, mask) as :319-348"""

    def __init__(self, vocab_size, emb_dim, enc_hid_dim, embed):
        super().__init__()
        self.vocab_size = vocab_size
        self.emb_dim = emb_dim
        self.enc_hid_dim = enc_hid_dim
        self.embed = embed
        self.lstm_encoder_cell = LSTMCellC(emb_dim, enc_hid_dim)
        self.affine_hn = nn.Linear(enc_hid_dim, enc_hid_dim)
        self.tanh = nn.Tanh()

    def forward(self, seq, seq_len):
        return self._dec().encode(seq, seq_len)


class CaptionAttentionC(_Owned, nn.Module):
    """editnet.py:351-381; forward(prev_h, h1, word, mask) -> (gated context, alpha) as :364-381"""

    def __init__(self, caption_features_dim, decoder_dim, attention_dim):
        super().__init__()
        self.cap_features_att = nn.Linear(caption_features_dim, attention_dim)
        self.cap_decoder_att = nn.Linear(decoder_dim, attention_dim)
        self.cap_full_att = nn.Linear(attention_dim, 1)
        self.context_gate = nn.Linear((caption_features_dim * 2) + decoder_dim, caption_features_dim)
        self.sc_affine = nn.Linear(caption_features_dim, caption_features_dim)
        self.tc_affine = nn.Linear(decoder_dim * 2, caption_features_dim)
        self.tanh = nn.Tanh()

    def forward(self, prev_cap_features, decoder_hidden, word, prev_cap_mask):
        dec = self._dec()
        prev_h, h1, emb, mask = _c(prev_cap_features), _c(decoder_hidden), _c(word), _c(prev_cap_mask)
        rows, P = prev_h.shape[0], prev_h.shape[1]
        dims = dec._dims()
        n = _lib.lib().set_caption_attention_scratch_floats(C.byref(dims), rows, P)
        scratch = torch.empty(n, device=prev_h.device)
        out = torch.empty(rows, dec.decoder_dim, device=prev_h.device)
        alpha = torch.empty(rows, P, device=prev_h.device)
        check(_lib.lib().set_caption_attention_forward(C.byref(dims), rows, P, C.byref(dec._struct), ptr(prev_h), ptr(h1),
                                                       ptr(emb), ptr(mask), ptr(scratch), n, ptr(out), ptr(alpha),
                                                       _stream()))
        return out, alpha


class SelectC(nn.Module):
    """editnet.py:383-421 (no parameters); forward(prev_m, alpha) -> selected memory row as :403-421"""

    def __init__(self, prev_caption_dim, decoder_dim):
        super().__init__()

    def forward(self, previous_encoded_m, alpha_c):
        if not previous_encoded_m.is_cuda:
            raise RuntimeError("show_edit_tell_b200 runs on a CUDA device only (no CPU fallback)")
        m, a = _c(previous_encoded_m), _c(alpha_c)
        rows, P, Dm = m.shape
        out = torch.empty(rows, Dm, device=m.device)
        check(_lib.lib().set_select_forward(rows, P, Dm, ptr(m), ptr(a), ptr(out), _stream()))
        return out


class VisualAttentionC(_Owned, nn.Module):
    """editnet.py:424-447 (adaptive_features/editnet_adaptive.py:423-457 when the decoder is the adaptive variant);
    forward(image_features, decoder_hidden) -> attended features as :439-447"""

    def __init__(self, image_features_dim, decoder_dim, attention_dim):
        super().__init__()
        self.att_embed = nn.Sequential(nn.Linear(image_features_dim, decoder_dim), nn.ReLU(), nn.Dropout(0.5))
        self.features_att = nn.Linear(decoder_dim, attention_dim)
        self.decoder_att = nn.Linear(decoder_dim, attention_dim)
        self.full_att = nn.Linear(attention_dim, 1)
        self.softmax = nn.Softmax(dim=1)

    def forward(self, image_features, decoder_hidden):
        dec = self._dec()
        feats, h1 = _c(image_features), _c(decoder_hidden)
        rows, R = feats.shape[0], feats.shape[1]
        dims = dec._dims()
        n = _lib.lib().set_visual_attention_scratch_floats(C.byref(dims), rows, R)
        scratch = torch.empty(n, device=feats.device)
        out = torch.empty(rows, dec.image_features_dim, device=feats.device)
        seed = _draw_seed() if self.training else 0
        check(_lib.lib().set_visual_attention_forward(C.byref(dims), rows, R, C.byref(dec._struct), ptr(feats), ptr(h1),
                                                      int(dec.ADAPTIVE), int(self.training), seed, ptr(scratch), n,
                                                      ptr(out), _stream()))
        return out


# ------------------------------------------------------------------------- autograd edge
class _XEFunction(torch.autograd.Function):
    """predictions = f(parameters): forward/backward are one C call each."""

    @staticmethod
    def forward(ctx, mod, call, *params):
        ctx.mod, ctx.call = mod, call
        return mod._xe_forward_raw(call)

    @staticmethod
    def backward(ctx, dpred):
        mod, call = ctx.mod, ctx.call
        flat_grad = torch.zeros_like(mod._flat)
        mod._xe_backward_raw(call, dpred.contiguous(), flat_grad)
        return (None, None) + tuple(mod._views(flat_grad))


class _RolloutFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, call, *params):
        ctx.mod, ctx.call = mod, call
        seq, slp = mod._rollout_raw(call)
        ctx.mark_non_differentiable(seq)
        return seq, slp

    @staticmethod
    def backward(ctx, dseq, dslp):
        mod, call = ctx.mod, ctx.call
        flat_grad = torch.zeros_like(mod._flat)
        mod._rollout_backward_raw(call, dslp.contiguous(), flat_grad)
        return (None, None) + tuple(mod._views(flat_grad))


class _Call:
    """everything one forward/backward pair shares"""
    pass


class EditNetBase(nn.Module):
    """DecoderC.__init__ of editnet.py:451-471 plus the plumbing shared by the XE, RL and
    adaptive front-ends."""

    ADAPTIVE = False

    def __init__(self, word_map, decoder_dim=1024, caption_features_dim=1024, emb_dim=1024, attention_dim=512,
                 image_features_dim=2048):
        super().__init__()
        if not (decoder_dim == caption_features_dim == emb_dim):
            raise ValueError("the reference's concatenations require decoder_dim == caption_features_dim == "
                             "emb_dim (editnet.py:359-361,468-469)")
        self.vocab_size = len(word_map)
        self.dropout = nn.Dropout(0.5)
        self.decoder_dim = decoder_dim
        self.attention_dim = attention_dim
        self.image_features_dim = image_features_dim
        self.embed = EmbeddingC(word_map, emb_dim)
        self.caption_encoder = CaptionEncoderC(len(word_map), emb_dim, caption_features_dim, self.embed)
        self.caption_attention = CaptionAttentionC(caption_features_dim, decoder_dim, attention_dim)
        self.visual_attention = VisualAttentionC(image_features_dim, decoder_dim, attention_dim)
        self.select = SelectC(caption_features_dim, decoder_dim)
        self.attention_lstm = LSTMCellK((emb_dim * 3) + image_features_dim, decoder_dim)
        self.copy_lstm = CopyLSTMCellC((emb_dim * 2) + image_features_dim, decoder_dim)
        self.tanh = nn.Tanh()
        self.fc = LinearK(decoder_dim, self.vocab_size)
        self._link_submodules()
        self._flat = None
        self._offsets = None
        self._struct = None
        self._last_call = None
        self.last_seed = None

    def _link_submodules(self):
        for m in (self.caption_encoder, self.caption_attention, self.visual_attention, self.copy_lstm):
            object.__setattr__(m, "_owner", weakref.ref(self))

    def __getstate__(self):
        # ctypes structs / cached calls are rebuilt lazily; parameters are pickled as ordinary tensors
        state = self.__dict__.copy()
        for k in ("_flat", "_offsets", "_struct", "_last_call"):
            state[k] = None
        return state

    def __setstate__(self, state):
        super().__setstate__(state)
        self._link_submodules()

    def init_hidden_state(self, batch_size):
        dev = self.fc.weight.device
        return (torch.zeros(batch_size, self.decoder_dim, device=dev),
                torch.zeros(batch_size, self.decoder_dim, device=dev))

    FIELDS = EDITNET_FIELDS
    STRUCT = SetEditNetParams
    # Gradient buckets of the data-parallel step, in the order the reverse pass finishes them (csrc/editnet.cu
    # backward_core, include/set_b200.h set_backward_bucket_events); fields not listed form the last bucket.
    BUCKET_FIELDS = (
        ("fc_w", "fc_b"),
        ("al_wih", "al_whh", "al_bih", "al_bhh",
         "cl_x2h_w", "cl_x2h_b", "cl_h2h_w", "cl_h2h_b", "cl_gcn_w", "cl_gcn_b", "cl_gcm_w", "cl_gcm_b",
         "ca_feat_w", "ca_feat_b"),
        ("embed", "enc_x2h_w", "enc_x2h_b", "enc_h2h_w", "enc_h2h_b", "enc_aff_w", "enc_aff_b"),
        ("ca_dec_w", "ca_dec_b", "ca_full_w", "ca_full_b", "ca_gate_w", "ca_gate_b", "ca_sc_w", "ca_sc_b",
         "ca_tc_w", "ca_tc_b", "va_dec_w", "va_dec_b", "va_full_w", "va_full_b"),
    )

    # ---- flat parameter storage: every parameter is a view into one buffer, so the optimizer
    # tail and the data-parallel all-reduce see a single tensor
    def _ordered_params(self):
        return [self.get_parameter(key) for _, key in self.FIELDS]

    def flatten_parameters(self):
        params = self._ordered_params()
        dev = params[0].device
        if self._flat is not None and self._flat.device == dev:
            base = self._flat.data_ptr()
            if all(p.data_ptr() == base + 4 * o for p, o in zip(params, self._offsets)):
                return self._flat
        # Layout of the flat buffer: one contiguous range per gradient bucket, in the order the reverse pass finishes
        # them (BUCKET_FIELDS, then everything else) -- a data-parallel step all-reduces each range underneath the rest of
        # the reverse pass as soon as it is final (train.py, set_backward_bucket_events).
        names = [n for n, _ in self.FIELDS]
        listed = [n for b in self.BUCKET_FIELDS for n in b]
        groups = [list(b) for b in self.BUCKET_FIELDS] + [[n for n in names if n not in listed]]
        groups = [g for g in groups if g]
        offs, total, starts = [0] * len(params), 0, []
        for g in groups:
            starts.append(total)
            for n in g:
                i = names.index(n)
                offs[i] = total
                total += (params[i].numel() + 63) // 64 * 64
        self._bucket_offsets = starts           # ascending; bucket k = flat[starts[k]:starts[k+1]] (last: to the end)
        flat = torch.zeros(total, device=dev, dtype=torch.float32)
        for p, o in zip(params, offs):
            view = flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
        self._flat, self._offsets = flat, offs
        st = self.STRUCT()
        for (name, _), p in zip(self.FIELDS, params):
            setattr(st, name, p.data_ptr())
        self._struct = st
        return flat

    def _views(self, flat):
        return [flat[o:o + p.numel()].view(p.shape) for p, o in zip(self._ordered_params(), self._offsets)]

    def _struct_for(self, flat):
        st = self.STRUCT()
        for (name, _), v in zip(self.FIELDS, self._views(flat)):
            setattr(st, name, v.data_ptr())
        return st

    def _dims(self):
        return SetDims(self.vocab_size, self.decoder_dim, self.attention_dim, self.image_features_dim)

    def _require_cuda(self, t):
        if not t.is_cuda:
            raise RuntimeError("show_edit_tell_b200 runs on a CUDA device only (no CPU fallback): "
                               "move the module and its inputs to cuda")

    # ---- teacher-forced path -------------------------------------------------------------
    def _prepare_xe(self, image_features, image_mean, encoded_captions, caption_lengths,
                    encoded_previous_captions, previous_cap_length, seed=None, host_lengths=None):
        """`host_lengths = (caption_lengths, previous_cap_length)` as CPU tensors (the loader has them on the host
        anyway) spares the device->host read of the lengths, i.e. the one host sync of a step."""
        self._require_cuda(image_features)
        self.flatten_parameters()
        take = lambda x: x[sort_ind]
        if host_lengths is not None:
            lens_h, sort_h = host_lengths[0].reshape(-1).sort(dim=0, descending=True, stable=True)
            sort_ind = sort_h.to(image_features.device, non_blocking=True)
            host = (lens_h - 1).tolist() + [int(host_lengths[1].max())]
            if bool((sort_h == torch.arange(sort_h.numel())).all()):
                # already in descending-length order (fixed-length or bucketed batches): the five gathers of the sort
                # (editnet.py:488-493) are identity copies -- skipped
                take = lambda x: x
        else:
            lens, sort_ind = caption_lengths.squeeze(1).sort(dim=0, descending=True, stable=True)  # editnet.py:488 (stable: ties as on CPU)
            host = None
        call = _Call()
        call.feats = take(image_features).contiguous().float()
        call.image_mean = None if image_mean is None else take(image_mean).contiguous().float()
        call.caps = take(encoded_captions).contiguous()
        call.prev = take(encoded_previous_captions).contiguous()
        call.prev_len = take(previous_cap_length).contiguous().view(-1)
        if host is None:
            host = torch.cat([lens - 1, call.prev_len.max().view(1)]).tolist()            # one D2H sync
        call.decode_lengths = host[:-1]
        P = int(host[-1])
        B, Wc = call.caps.shape
        T = max(call.decode_lengths)
        call.shape = SetSeqShape(B, call.feats.shape[1], Wc, call.prev.shape[1], P, T, int(self.training),
                                 int(self.ADAPTIVE))
        call.dims = self._dims()
        call.dec_host = (C.c_int * B)(*call.decode_lengths)
        call.seed = (_draw_seed() if seed is None else seed) if self.training else 0
        self.last_seed = call.seed
        nbytes = _lib.lib().set_editnet_workspace_bytes(C.byref(call.dims), C.byref(call.shape))
        if nbytes == 0:
            raise RuntimeError("libset_b200: " + _lib.lib().set_last_error().decode())
        call.ws = torch.empty(nbytes, dtype=torch.uint8, device=call.feats.device)
        call.sort_ind = sort_ind
        return call

    def _xe_forward_raw(self, call):
        s = call.shape
        pred = torch.empty(s.B, s.T, self.vocab_size, device=call.feats.device, dtype=torch.float32)
        ss_prob = getattr(call, "ss_prob", 0.0)
        replay = getattr(call, "ss_replay", None)
        if ss_prob > 0.0 or replay is not None:
            call.fed = torch.empty_like(call.caps)
            check(_lib.lib().set_editnet_xe_forward_ss(
                C.byref(call.dims), C.byref(s), C.byref(self._struct), ptr(call.feats), ptr(call.image_mean),
                ptr(call.caps), call.dec_host, ptr(call.prev), ptr(call.prev_len), call.seed, ss_prob, ptr(replay),
                ptr(call.fed), ptr(pred), ptr(call.ws), call.ws.numel(), _stream()))
            return pred
        call.fed = call.caps
        check(_lib.lib().set_editnet_xe_forward(
            C.byref(call.dims), C.byref(s), C.byref(self._struct), ptr(call.feats), ptr(call.image_mean),
            ptr(call.caps), call.dec_host, ptr(call.prev), ptr(call.prev_len), call.seed, ptr(pred), ptr(call.ws),
            call.ws.numel(), _stream()))
        return pred

    def _xe_backward_raw(self, call, dpred, flat_grad):
        g = self._struct_for(flat_grad)
        check(_lib.lib().set_editnet_xe_backward(
            C.byref(call.dims), C.byref(call.shape), C.byref(self._struct), C.byref(g), ptr(call.feats),
            ptr(call.fed), call.dec_host, ptr(call.prev), ptr(call.prev_len), call.seed, ptr(dpred), ptr(call.ws),
            call.ws.numel(), _stream()))

    def _xe(self, image_features, image_mean, encoded_captions, caption_lengths, encoded_previous_captions,
            previous_cap_length, use_ss, ss_prob):
        call = self._prepare_xe(image_features, image_mean, encoded_captions, caption_lengths,
                                encoded_previous_captions, previous_cap_length)
        if use_ss and ss_prob > 0.0:                       # scheduled sampling, editnet.py:508-520
            if not self.training:
                raise RuntimeError("scheduled sampling is a training-time feature (train mode required)")
            call.ss_prob = float(ss_prob)
        call.ss_replay = getattr(self, "_ss_replay", None)  # tests: force the fed tokens
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            pred = _XEFunction.apply(self, call, *self._ordered_params())
        else:
            pred = self._xe_forward_raw(call)
        self._last_call = call
        return pred, call

    # ---- rollout path -----------------------------------------------------------------------
    def _prepare_rollout(self, encoded_previous_captions, previous_cap_length, image_features, image_mean, mode,
                         max_len, start_idx, end_idx, forced=None, seed=None, keep=None):
        self._require_cuda(image_features)
        self.flatten_parameters()
        call = _Call()
        call.feats = image_features.contiguous().float()
        call.image_mean = None if image_mean is None else image_mean.contiguous().float()
        call.prev = encoded_previous_captions.contiguous()
        call.prev_len = previous_cap_length.contiguous().view(-1)
        P = int(call.prev_len.max().item())
        B = call.feats.shape[0]
        keep = self.training if keep is None else keep
        call.shape = SetSeqShape(B, call.feats.shape[1], 0, call.prev.shape[1], P, max_len, int(keep),
                                 int(self.ADAPTIVE))
        call.dims = self._dims()
        call.mode, call.start_idx, call.end_idx = mode, start_idx, end_idx
        call.forced = None if forced is None else forced.contiguous()
        call.seed = _draw_seed() if seed is None else seed
        self.last_seed = call.seed
        nbytes = _lib.lib().set_editnet_workspace_bytes(C.byref(call.dims), C.byref(call.shape))
        if nbytes == 0:
            raise RuntimeError("libset_b200: " + _lib.lib().set_last_error().decode())
        call.ws = torch.empty(nbytes, dtype=torch.uint8, device=call.feats.device)
        return call

    def _rollout_raw(self, call):
        s = call.shape
        dev = call.feats.device
        seq = torch.empty(s.B, s.T, dtype=torch.int64, device=dev)
        slp = torch.empty(s.B, s.T, dtype=torch.float32, device=dev)
        check(_lib.lib().set_editnet_rollout(
            C.byref(call.dims), C.byref(s), C.byref(self._struct), ptr(call.feats), ptr(call.image_mean),
            ptr(call.prev), ptr(call.prev_len), call.start_idx, call.end_idx, call.mode, ptr(call.forced), call.seed,
            ptr(seq), ptr(slp), ptr(call.ws), call.ws.numel(), _stream()))
        return seq, slp

    def _rollout_backward_raw(self, call, dslp, flat_grad):
        g = self._struct_for(flat_grad)
        check(_lib.lib().set_editnet_rollout_backward(
            C.byref(call.dims), C.byref(call.shape), C.byref(self._struct), C.byref(g), ptr(call.feats),
            ptr(call.prev), ptr(call.prev_len), call.seed, ptr(dslp), ptr(call.ws), call.ws.numel(), _stream()))

    def rollout(self, word_map, encoded_previous_captions, previous_cap_length, image_features, sample_max,
                sample_rl, max_len=18, image_mean=None, forced=None, seed=None):
        """editnet_rl.py:485-549.  `forced` (B,max_len) replays given tokens instead of sampling."""
        mode = 2 if forced is not None else (1 if sample_rl else 0)
        want_grad = torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters())
        call = self._prepare_rollout(encoded_previous_captions, previous_cap_length, image_features, image_mean,
                                     mode, max_len, word_map['<start>'], word_map['<end>'], forced, seed,
                                     keep=self.training)
        self._last_call = call
        if want_grad:
            return _RolloutFunction.apply(self, call, *self._ordered_params())
        return self._rollout_raw(call)

    def encode(self, seq, seq_len, seed=None):
        """CaptionEncoderC.forward (editnet.py:319-348) on its own, no autograd:
        -> (hidden_states (B,P',D), memory_states (B,P',D), final_hidden (B,D), mask (B,P'))"""
        self._require_cuda(seq)
        self.flatten_parameters()
        seq = seq.contiguous()
        lens = seq_len.contiguous().view(-1)
        B, Wp = seq.shape
        P = int(lens.max().item())
        dims = self._dims()
        shape = SetSeqShape(B, 1, 0, Wp, P, 1, int(self.training), 0)
        nbytes = _lib.lib().set_editnet_workspace_bytes(C.byref(dims), C.byref(shape))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=seq.device)
        D = self.decoder_dim
        h = torch.empty(B, P, D, device=seq.device)
        m = torch.empty(B, P, D, device=seq.device)
        fh = torch.empty(B, D, device=seq.device)
        mask = torch.empty(B, P, device=seq.device)
        sd = (_draw_seed() if seed is None else seed) if self.training else 0
        check(_lib.lib().set_editnet_encode(C.byref(dims), C.byref(shape), C.byref(self._struct), ptr(seq), ptr(lens), sd,
                                            ptr(h), ptr(m), ptr(fh), ptr(mask), ptr(ws), ws.numel(), _stream()))
        return h, m, fh, mask

    # debugging / tests: a named workspace buffer of the last call as a float tensor
    def workspace_tensor(self, name, dtype=torch.float32):
        call = self._last_call
        off, nbytes = C.c_size_t(), C.c_size_t()
        check(_lib.lib().set_editnet_workspace_lookup(C.byref(call.dims), C.byref(call.shape), name.encode(),
                                                      C.byref(off), C.byref(nbytes)))
        return call.ws[off.value:off.value + nbytes.value].view(dtype)


class DecoderC(EditNetBase):
    """Drop-in for `DecoderC` of editnet.py:449-548 (cross-entropy stage)."""

    def forward(self, image_features, encoded_captions, caption_lengths, encoded_previous_captions,
                previous_cap_length, use_ss=False, ss_prob=0.0):
        pred, call = self._xe(image_features, None, encoded_captions, caption_lengths, encoded_previous_captions,
                              previous_cap_length, use_ss, ss_prob)
        return pred, call.caps, call.decode_lengths, call.sort_ind


class StepSession:
    """Decode steps on explicit state for `k` rows that share nothing but the model (beam search,
    interactive decoding).  Built by `EditNetBase.step_session`."""

    def __init__(self, mod, image_features, encoded_previous_captions, previous_cap_length, image_mean=None):
        mod._require_cuda(image_features)
        mod.flatten_parameters()
        self.mod = mod
        self.feats = image_features.contiguous().float()
        self.image_mean = None if image_mean is None else image_mean.contiguous().float()
        prev = encoded_previous_captions.contiguous()
        prev_len = previous_cap_length.contiguous().view(-1)
        k = self.feats.shape[0]
        self.dims = mod._dims()
        self.shape = SetSeqShape(k, self.feats.shape[1], 0, prev.shape[1], int(prev_len.max().item()), 2, 0,
                                 int(mod.ADAPTIVE))
        nbytes = _lib.lib().set_editnet_workspace_bytes(C.byref(self.dims), C.byref(self.shape))
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.feats.device)
        check(_lib.lib().set_editnet_step_begin(C.byref(self.dims), C.byref(self.shape), C.byref(mod._struct),
                                                ptr(self.feats), ptr(self.image_mean), ptr(prev), ptr(prev_len),
                                                ptr(self.ws), self.ws.numel(), _stream()))

    def init_state(self):
        k, D = self.shape.B, self.mod.decoder_dim
        return tuple(torch.zeros(k, D, device=self.feats.device) for _ in range(4))

    def step(self, tokens, state):
        """tokens (rows,) int64; state = (h1, c1, h2, c2) each (rows, D) -> (scores (rows, V), new state).
        Equivalent to editnet.py:645-653 (embed .. fc) on the first `rows` rows of the session."""
        rows = tokens.shape[0]
        st = [x[:rows].contiguous().clone() for x in state]
        scores = torch.empty(rows, self.mod.vocab_size, device=self.feats.device)
        check(_lib.lib().set_editnet_step(C.byref(self.dims), C.byref(self.shape), C.byref(self.mod._struct),
                                          ptr(self.feats), ptr(tokens.contiguous()), rows, ptr(st[0]), ptr(st[1]),
                                          ptr(st[2]), ptr(st[3]), ptr(scores), ptr(self.ws), self.ws.numel(), _stream()))
        return scores, tuple(st)

    def step_raw(self, tokens, state, scores):
        """in place: `state` = [h1, c1, h2, c2] (B, D) buffers are advanced, `scores` (B, V) receives fc(h2)"""
        check(_lib.lib().set_editnet_step(C.byref(self.dims), C.byref(self.shape), C.byref(self.mod._struct),
                                          ptr(self.feats), ptr(tokens), self.shape.B, ptr(state[0]), ptr(state[1]),
                                          ptr(state[2]), ptr(state[3]), ptr(scores), ptr(self.ws), self.ws.numel(), _stream()))


def _step_session(self, image_features, encoded_previous_captions, previous_cap_length, image_mean=None):
    return StepSession(self, image_features, encoded_previous_captions, previous_cap_length, image_mean)


EditNetBase.step_session = _step_session


def beam_search(decoder, word_map, image_features, encoded_previous_caption, previous_cap_length, beam_size=3,
                max_steps=50):
    """The search loop of evaluate(), editnet.py:608-719, for one image: `image_features` (1,R,F),
    `encoded_previous_caption` (1,Wp), `previous_cap_length` (1,1).  Each step is ONE library call on the
    k live beams instead of eight module calls; `top_k_words // vocab_size` replaces the reference's `/`
    (true division since torch 1.5).  Returns (token list incl. <start>/<end>, score)."""
    k = beam_size
    V = decoder.vocab_size
    dev = image_features.device
    sess = decoder.step_session(image_features.expand(k, -1, -1), encoded_previous_caption.expand(k, -1),
                                previous_cap_length.expand(k, -1))                          # :616-621
    k_prev_words = torch.full((k,), word_map['<start>'], dtype=torch.long, device=dev)     # :623
    seqs = k_prev_words.unsqueeze(1)                                                        # :626
    top_k_scores = torch.zeros(k, 1, device=dev)                                            # :629
    complete_seqs, complete_scores = [], []
    state = sess.init_state()
    step = 1
    runaway = False
    while True:
        scores, state = sess.step(k_prev_words, state)                                      # :645-653
        scores = torch.log_softmax(scores, dim=1)                                           # :654
        scores = top_k_scores.expand_as(scores) + scores                                    # :657
        if step == 1:
            top_k_scores, top_k_words = scores[0].topk(k, 0, True, True)                    # :660-661
        else:
            top_k_scores, top_k_words = scores.view(-1).topk(k, 0, True, True)              # :663
        prev_word_inds = top_k_words // V                                                   # :666
        next_word_inds = top_k_words % V                                                    # :667
        seqs = torch.cat([seqs[prev_word_inds], next_word_inds.unsqueeze(1)], dim=1)        # :670
        nxt = next_word_inds.tolist()
        incomplete = [i for i, w in enumerate(nxt) if w != word_map['<end>']]               # :673-674
        complete = [i for i in range(len(nxt)) if i not in incomplete]
        if complete:
            complete_seqs.extend(seqs[complete].tolist())                                   # :678-679
            complete_scores.extend(top_k_scores[complete].tolist())
        k -= len(complete)                                                                  # :680
        if k == 0:
            break
        inc = torch.tensor(incomplete, device=dev, dtype=torch.long)
        seqs = seqs[inc]
        sel = prev_word_inds[inc]
        state = tuple(x[sel] for x in state)                                                # :686-691
        top_k_scores = top_k_scores[inc].unsqueeze(1)
        k_prev_words = next_word_inds[inc]
        if step > max_steps:                                                                # :702
            runaway = True
            break
        step += 1
    if runaway or not complete_scores:
        # the 50-step guard: the reference emits the first 18 tokens of the best live beam even when other beams
        # completed earlier (:702-713)
        return seqs[0][:18].tolist(), float(top_k_scores[0])
    i = complete_scores.index(max(complete_scores))                                         # :706
    return complete_seqs[i], complete_scores[i]
