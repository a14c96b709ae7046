"""Batched beam search on the device: evaluate() (editnet.py:595-719) and the EditNet + DCNet ensemble evaluate_full()
(eval/eval xe/eval_full.py:88-215) for many images in ONE step session.

The reference decodes one image at a time (batch-1 loader, editnet.py:795-798) and keeps the search state in python
lists with a `.tolist()` per step.  Here image i owns rows i*K .. i*K+K-1 of a step session; each step is one network
step on all N*K rows (two for the ensemble) + one expansion kernel (log-softmax / ensemble mix, top-k over live beams x
vocabulary, <end> bookkeeping, set_beam_expand) + one state re-gather -- nothing returns to the host until the captions
are done.  The host only polls, without blocking, a counter of images that still have live beams to stop early.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import check, ptr
from .editnet import _stream


def beam_search_batched(decoder, word_map, image_features, encoded_previous_captions, previous_cap_length, beam_size=3,
                        max_steps=50, dae=None, image_mean=None):
    """image_features (N,R,F), encoded_previous_captions (N,Wp), previous_cap_length (N,1) -> list of N
    (token list incl. <start>/<end>, score).  `dae` (a DCNet `DAE` / `DAEWithAR`) switches on the ensemble scoring of
    eval_full.py:151-153.  Same results as running the reference's per-image search N times."""
    L = _lib.lib()
    dae = getattr(dae, "dae", dae)
    N, K, V = image_features.shape[0], beam_size, decoder.vocab_size
    dev = image_features.device
    rows, D = N * K, decoder.decoder_dim
    rep = lambda x: None if x is None else x.repeat_interleave(K, 0)
    esess = decoder.step_session(rep(image_features), rep(encoded_previous_captions), rep(previous_cap_length), rep(image_mean))
    dsess = dae.step_session(rep(encoded_previous_captions), rep(previous_cap_length)) if dae is not None else None
    Lmax = max_steps + 3
    start, end = word_map['<start>'], word_map['<end>']
    tokens = torch.full((rows,), start, dtype=torch.long, device=dev)
    next_tokens = torch.empty_like(tokens)
    est = [torch.zeros(rows, D, device=dev) for _ in range(4)]
    est_alt = [torch.empty_like(x) for x in est]
    dst = [torch.zeros(rows, D, device=dev) for _ in range(4)] if dsess else None
    dst_alt = [torch.empty_like(x) for x in dst] if dsess else None
    escores = torch.empty(rows, V, device=dev)
    dscores = torch.empty(rows, V, device=dev) if dsess else None
    k_live = torch.full((N,), K, dtype=torch.int32, device=dev)
    beam_scores = torch.zeros(rows, device=dev)
    seq_a = torch.zeros(rows, Lmax, dtype=torch.long, device=dev)
    seq_a[:, 0] = start
    seq_b = torch.zeros_like(seq_a)
    src_row = torch.empty(rows, dtype=torch.int32, device=dev)
    n_complete = torch.zeros(N, dtype=torch.int32, device=dev)
    complete_scores = torch.zeros(rows, device=dev)
    complete_seqs = torch.zeros(rows, Lmax, dtype=torch.long, device=dev)
    complete_len = torch.zeros(rows, dtype=torch.int32, device=dev)
    live_images = torch.full((1,), N, dtype=torch.int32, device=dev)
    flag_host = torch.zeros(1, dtype=torch.int32).pin_memory()
    flag_ev = None
    steps_done = 0
    for step in range(1, max_steps + 2):                     # the reference runs steps 1 .. 51 (editnet.py:702-705)
        esess.step_raw(tokens, est, escores)
        if dsess:
            dsess.step_raw(tokens, dst, dscores)
        check(L.set_beam_expand(N, K, V, step, Lmax, end, ptr(escores), ptr(dscores), ptr(k_live), ptr(beam_scores),
                                ptr(seq_a), ptr(seq_b), ptr(next_tokens), ptr(src_row), ptr(n_complete),
                                ptr(complete_scores), ptr(complete_seqs), ptr(complete_len), ptr(live_images), _stream()))
        check(L.set_beam_gather(rows, D, ptr(src_row), *[ptr(x) for x in est], *[ptr(x) for x in est_alt], _stream()))
        est, est_alt = est_alt, est
        if dsess:
            check(L.set_beam_gather(rows, D, ptr(src_row), *[ptr(x) for x in dst], *[ptr(x) for x in dst_alt], _stream()))
            dst, dst_alt = dst_alt, dst
        seq_a, seq_b = seq_b, seq_a
        tokens, next_tokens = next_tokens, tokens
        steps_done = step
        # early stop without blocking: look at the live-image counter of an earlier step once its copy has landed
        if flag_ev is not None and flag_ev.query():
            if int(flag_host[0]) == 0:
                break
            flag_ev = None
        if flag_ev is None:
            flag_host.copy_(live_images, non_blocking=True)
            flag_ev = torch.cuda.Event()
            flag_ev.record(torch.cuda.current_stream())
    out_seq = torch.empty(N, Lmax, dtype=torch.long, device=dev)
    out_len = torch.empty(N, dtype=torch.int32, device=dev)
    out_score = torch.empty(N, device=dev)
    check(L.set_beam_finalize(N, K, Lmax, steps_done, ptr(k_live), ptr(seq_a), ptr(beam_scores), ptr(n_complete),
                              ptr(complete_scores), ptr(complete_seqs), ptr(complete_len), ptr(out_seq), ptr(out_len),
                              ptr(out_score), _stream()))
    seqs, lens, scores = out_seq.cpu(), out_len.cpu().tolist(), out_score.cpu().tolist()     # the one host sync
    return [(seqs[i, :lens[i]].tolist(), scores[i]) for i in range(N)]
