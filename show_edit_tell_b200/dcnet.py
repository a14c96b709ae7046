"""DCNet host side: the reference's `DAE` module surface (dcnet.py:147-350) over the CUDA C ABI.
Same approach as editnet.py: torch modules only hold the parameters (reference names / shapes /
`state_dict` keys); forward and backward are C calls into libset_b200.so.  No CPU path."""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from ._lib import DCNET_FIELDS, SetDcNetParams, SetDims, SetSeqShape, check, ptr
import weakref

from .editnet import EditNetBase, EmbeddingC, LinearK, LSTMCellK, _c, _Call, _draw_seed, _Owned, _stream


class Embedding(nn.Module):
    """dcnet.py:147-206 (load_glove_embedding=False path, the one the reference uses, dcnet.py:288)"""

    def __init__(self, word_map, emb_file, emb_dim, load_glove_embedding=False):
        super().__init__()
        if load_glove_embedding:
            raise NotImplementedError("the GloVe path is dead code in the reference (dcnet.py:288)")
        self.emb_dim = emb_dim
        self.load_glove_embedding = False
        self.emb_file = emb_file
        self.word_map = word_map
        self.embedding = nn.Embedding(len(word_map), self.emb_dim)
        self.relu = nn.ReLU()
        self.dropout = nn.Dropout(0.5)

    forward = EmbeddingC.forward     # dropout(relu(Emb[x])), dcnet.py:199-206


class CaptionEncoder(_Owned, nn.Module):
    """dcnet.py:209-243; forward(src, src_len) -> (outputs (B,P',2C), final_hidden (B,2C), mask (B,P')) as :220-243"""

    def __init__(self, vocab_size, emb_dim, enc_hid_dim, concat_output_dim, embed):
        super().__init__()
        self.vocab_size = vocab_size
        self.emb_dim = emb_dim
        self.enc_hid_dim = enc_hid_dim
        self.embed = embed
        self.lstm_encoder = nn.LSTM(emb_dim, enc_hid_dim, batch_first=True, bidirectional=True)
        self.concat = nn.Linear(enc_hid_dim * 2, concat_output_dim)

    def forward(self, src, src_len):
        dae = self._dec()
        sess = dae.step_session(src, src_len)        # runs the bi-LSTM encoder into the session workspace
        B, P, D = sess.shape.B, sess.shape.P, dae.decoder_dim

        def buf(name):
            off, nbytes = C.c_size_t(), C.c_size_t()
            check(_lib.lib().set_dcnet_workspace_lookup(C.byref(sess.dims), C.byref(sess.shape), name.encode(),
                                                        C.byref(off), C.byref(nbytes)))
            return sess.ws[off.value:off.value + nbytes.value].view(torch.float32)

        return (buf("enc_out").view(B, P, D).clone(), buf("final_hidden").view(B, D).clone(), buf("mask").view(B, P).clone())


class CaptionAttention(_Owned, nn.Module):
    """dcnet.py:245-270; forward(enc, h1, mask) -> context as :254-270"""

    def __init__(self, caption_features_dim, decoder_dim, attention_dim):
        super().__init__()
        self.cap_features_att = nn.Linear(caption_features_dim * 2, attention_dim)
        self.cap_decoder_att = nn.Linear(decoder_dim, attention_dim)
        self.cap_full_att = nn.Linear(attention_dim, 1)

    def forward(self, prev_cap_features, decoder_hidden, prev_cap_mask):
        dae = self._dec()
        enc, h1, mask = _c(prev_cap_features), _c(decoder_hidden), _c(prev_cap_mask)
        rows, P = enc.shape[0], enc.shape[1]
        dims = dae._dims()
        n = _lib.lib().set_dcnet_caption_attention_scratch_floats(C.byref(dims), rows, P)
        scratch = torch.empty(n, device=enc.device)
        out = torch.empty(rows, dae.decoder_dim, device=enc.device)
        check(_lib.lib().set_dcnet_caption_attention_forward(C.byref(dims), rows, P, C.byref(dae._struct), ptr(enc), ptr(h1),
                                                             ptr(mask), ptr(scratch), n, ptr(out), _stream()))
        return out


class _DXEFunction(torch.autograd.Function):
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


class _DRolloutFunction(torch.autograd.Function):
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


class DAEBase(nn.Module):
    """DAE.__init__ of dcnet.py:275-295 + plumbing"""

    FIELDS = DCNET_FIELDS
    STRUCT = SetDcNetParams
    BUCKET_FIELDS = ()
    # flat-parameter plumbing shared with EditNet
    _ordered_params = EditNetBase._ordered_params
    flatten_parameters = EditNetBase.flatten_parameters
    _views = EditNetBase._views
    _struct_for = EditNetBase._struct_for
    _require_cuda = EditNetBase._require_cuda
    __getstate__ = EditNetBase.__getstate__

    def __init__(self, word_map, emb_file=None, decoder_dim=1024, attention_dim=512, caption_features_dim=512,
                 emb_dim=1024):
        super().__init__()
        if not (decoder_dim == emb_dim == 2 * caption_features_dim):
            raise ValueError("the reference's concatenations require decoder_dim == emb_dim == "
                             "2 * caption_features_dim (dcnet.py:286-291)")
        self.vocab_size = len(word_map)
        self.attention_lstm = LSTMCellK(emb_dim * 3, decoder_dim)
        self.language_lstm = LSTMCellK(emb_dim * 2, decoder_dim)
        self.embed = Embedding(word_map, emb_file, emb_dim, load_glove_embedding=False)
        self.caption_encoder = CaptionEncoder(len(word_map), emb_dim, caption_features_dim, caption_features_dim * 2,
                                              self.embed)
        self.caption_attention = CaptionAttention(caption_features_dim, decoder_dim, attention_dim)
        self.fc = LinearK(decoder_dim, len(word_map))
        self.tanh = nn.Tanh()
        self._link_submodules()
        self.decoder_dim = decoder_dim
        self.attention_dim = attention_dim
        self.dropout = nn.Dropout(0.5)
        self._flat = None
        self._offsets = None
        self._struct = None
        self._last_call = None
        self.last_seed = None

    def _link_submodules(self):
        for m in (self.caption_encoder, self.caption_attention):
            object.__setattr__(m, "_owner", weakref.ref(self))

    def __setstate__(self, state):
        super().__setstate__(state)
        self._link_submodules()

    def init_hidden_state(self, batch_size):
        dev = self.fc.weight.device
        return (torch.zeros(batch_size, self.decoder_dim, device=dev),
                torch.zeros(batch_size, self.decoder_dim, device=dev))

    def _dims(self):
        return SetDims(self.vocab_size, self.decoder_dim, self.attention_dim, 4)

    def _workspace(self, call):
        nbytes = _lib.lib().set_dcnet_workspace_bytes(C.byref(call.dims), C.byref(call.shape))
        if nbytes == 0:
            raise RuntimeError("libset_b200: " + _lib.lib().set_last_error().decode())
        return torch.empty(nbytes, dtype=torch.uint8, device=call.prev.device)

    # ---- teacher forced
    def _prepare_xe(self, encoded_captions, caption_lengths, encoded_previous_captions, previous_cap_length, seed=None):
        self._require_cuda(encoded_captions)
        self.flatten_parameters()
        lens, sort_ind = caption_lengths.squeeze(1).sort(dim=0, descending=True, stable=True)   # dcnet.py:314
        call = _Call()
        call.caps = encoded_captions[sort_ind].contiguous()
        call.prev = encoded_previous_captions[sort_ind].contiguous()
        call.prev_len = previous_cap_length[sort_ind].contiguous().view(-1)
        host = torch.cat([lens - 1, call.prev_len.max().view(1)]).tolist()
        call.decode_lengths = host[:-1]
        B, Wc = call.caps.shape
        call.shape = SetSeqShape(B, 0, Wc, call.prev.shape[1], int(host[-1]), max(call.decode_lengths),
                                 int(self.training), 0)
        call.dims = self._dims()
        call.dec_host = (C.c_int * B)(*call.decode_lengths)
        call.seed = (_draw_seed() if seed is None else seed) if self.training else 0
        self.last_seed = call.seed
        call.ws = self._workspace(call)
        call.sort_ind = sort_ind
        return call

    def _xe_forward_raw(self, call):
        s = call.shape
        pred = torch.empty(s.B, s.T, self.vocab_size, device=call.caps.device, dtype=torch.float32)
        check(_lib.lib().set_dcnet_xe_forward(
            C.byref(call.dims), C.byref(s), C.byref(self._struct), ptr(call.caps), call.dec_host, ptr(call.prev),
            ptr(call.prev_len), call.seed, ptr(pred), ptr(call.ws), call.ws.numel(), _stream()))
        return pred

    def _xe_backward_raw(self, call, dpred, flat_grad):
        g = self._struct_for(flat_grad)
        check(_lib.lib().set_dcnet_xe_backward(
            C.byref(call.dims), C.byref(call.shape), C.byref(self._struct), C.byref(g), ptr(call.caps), call.dec_host,
            ptr(call.prev), ptr(call.prev_len), call.seed, ptr(dpred), ptr(call.ws), call.ws.numel(), _stream()))

    # ---- rollout
    def rollout(self, word_map, encoded_previous_captions, previous_cap_length, sample_max, sample_rl, max_len=18,
                forced=None, seed=None):
        """dcnet_rl.py:286-346"""
        self._require_cuda(encoded_previous_captions)
        self.flatten_parameters()
        call = _Call()
        call.prev = encoded_previous_captions.contiguous()
        call.prev_len = previous_cap_length.contiguous().view(-1)
        B = call.prev.shape[0]
        call.shape = SetSeqShape(B, 0, 0, call.prev.shape[1], int(call.prev_len.max().item()), max_len,
                                 int(self.training), 0)
        call.dims = self._dims()
        call.mode = 2 if forced is not None else (1 if sample_rl else 0)
        call.start_idx, call.end_idx = word_map['<start>'], word_map['<end>']
        call.forced = None if forced is None else forced.contiguous()
        call.seed = _draw_seed() if seed is None else seed
        self.last_seed = call.seed
        call.ws = self._workspace(call)
        self._last_call = call
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            return _DRolloutFunction.apply(self, call, *self._ordered_params())
        return self._rollout_raw(call)

    def _rollout_raw(self, call):
        s = call.shape
        dev = call.prev.device
        seq = torch.empty(s.B, s.T, dtype=torch.int64, device=dev)
        slp = torch.empty(s.B, s.T, dtype=torch.float32, device=dev)
        check(_lib.lib().set_dcnet_rollout(
            C.byref(call.dims), C.byref(s), C.byref(self._struct), ptr(call.prev), ptr(call.prev_len), call.start_idx,
            call.end_idx, call.mode, ptr(call.forced), call.seed, ptr(seq), ptr(slp), ptr(call.ws), call.ws.numel(),
            _stream()))
        return seq, slp

    def _rollout_backward_raw(self, call, dslp, flat_grad):
        g = self._struct_for(flat_grad)
        check(_lib.lib().set_dcnet_rollout_backward(
            C.byref(call.dims), C.byref(call.shape), C.byref(self._struct), C.byref(g), ptr(call.prev),
            ptr(call.prev_len), call.seed, ptr(dslp), ptr(call.ws), call.ws.numel(), _stream()))

    def workspace_tensor(self, name, dtype=torch.float32):
        call = self._last_call
        off, nbytes = C.c_size_t(), C.c_size_t()
        check(_lib.lib().set_dcnet_workspace_lookup(C.byref(call.dims), C.byref(call.shape), name.encode(),
                                                    C.byref(off), C.byref(nbytes)))
        return call.ws[off.value:off.value + nbytes.value].view(dtype)


class DStepSession:
    """Decode steps of the DCNet on explicit state for `k` rows (the DCNet half of the ensemble beam search,
    eval/eval xe/eval_full.py:141-149).  Built by `DAEBase.step_session`."""

    def __init__(self, mod, encoded_previous_captions, previous_cap_length):
        mod._require_cuda(encoded_previous_captions)
        mod.flatten_parameters()
        self.mod = mod
        prev = encoded_previous_captions.contiguous()
        prev_len = previous_cap_length.contiguous().view(-1)
        k = prev.shape[0]
        self.dims = mod._dims()
        self.shape = SetSeqShape(k, 0, 0, prev.shape[1], int(prev_len.max().item()), 2, 0, 0)
        nbytes = _lib.lib().set_dcnet_workspace_bytes(C.byref(self.dims), C.byref(self.shape))
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=prev.device)
        check(_lib.lib().set_dcnet_step_begin(C.byref(self.dims), C.byref(self.shape), C.byref(mod._struct), ptr(prev),
                                              ptr(prev_len), ptr(self.ws), self.ws.numel(), _stream()))

    def init_state(self):
        k, D = self.shape.B, self.mod.decoder_dim
        return tuple(torch.zeros(k, D, device=self.ws.device) for _ in range(4))

    def step(self, tokens, state):
        """tokens (rows,) int64; state = (h1, c1, h2, c2) each (rows, D) -> (scores (rows, V), new state)"""
        rows = tokens.shape[0]
        st = [x[:rows].contiguous().clone() for x in state]
        scores = torch.empty(rows, self.mod.vocab_size, device=self.ws.device)
        check(_lib.lib().set_dcnet_step(C.byref(self.dims), C.byref(self.shape), C.byref(self.mod._struct),
                                        ptr(tokens.contiguous()), rows, ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]),
                                        ptr(scores), ptr(self.ws), self.ws.numel(), _stream()))
        return scores, tuple(st)

    def step_raw(self, tokens, state, scores):
        """in place: `state` = [h1, c1, h2, c2] (B, D) buffers are advanced, `scores` (B, V) receives fc(h2)"""
        check(_lib.lib().set_dcnet_step(C.byref(self.dims), C.byref(self.shape), C.byref(self.mod._struct), ptr(tokens),
                                        self.shape.B, ptr(state[0]), ptr(state[1]), ptr(state[2]), ptr(state[3]),
                                        ptr(scores), ptr(self.ws), self.ws.numel(), _stream()))


def _dstep_session(self, encoded_previous_captions, previous_cap_length):
    return DStepSession(self, encoded_previous_captions, previous_cap_length)


DAEBase.step_session = _dstep_session


class DAE(DAEBase):
    """Drop-in for `DAE` of dcnet.py:273-350 (cross-entropy stage)."""

    def forward(self, encoded_captions, caption_lengths, encoded_previous_captions, previous_cap_length):
        call = self._prepare_xe(encoded_captions, caption_lengths, encoded_previous_captions, previous_cap_length)
        self._last_call = call
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            pred = _DXEFunction.apply(self, call, *self._ordered_params())
        else:
            pred = self._xe_forward_raw(call)
        return pred, call.caps, call.decode_lengths, call.sort_ind
