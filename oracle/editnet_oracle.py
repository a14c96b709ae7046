"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's EditNet hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl
reference` legs may import this module.  The product path
(`show_edit_tell_b200/`) never does; it fails loudly when the CUDA library is
missing.

This is a *restatement*, not a copy: the reference is a set of `nn.Module`
classes (`/root/reference/editnet.py:210-548`, `editnet_rl.py:485-573`,
`adaptive_features/editnet_adaptive.py:423-562`); here the same arithmetic is
written as pure functions over a `state_dict` (keys of SURVEY.md Appendix B), in
the reference's own fp32 torch ops so autograd supplies the gradients the
backward kernels are checked against.  Every function cites the lines it follows.

Pinning: the reference has no tests or golden vectors of its own.
The pin is `oracle/make_golden.py`, which runs the reference's real classes
(AST-extracted, unmodified) in the authoring container and stores their outputs
under `tests/golden/`; `tests/test_oracle_golden.py` holds this restatement to
those outputs, and `tests/test_oracle_vs_reference.py` repeats the comparison
live wherever /root/reference exists.

Dropout.  The reference draws fresh Bernoulli(0.5) masks at four sites
(`editnet.py:303` embed -- called once by the encoder at `:331` and once per
decode step at `:513/520/524`; `:432` att_embed, per step; `:545` fc, per step).
Here masks are explicit inputs (0/1 keep flags, scaled by 2 on use) in
*decoder-sorted row order*:
    masks['enc'] (B, Pw, E)   masks['emb'] (T, B, E)
    masks['vis'] (T, B, R, D) masks['fc']  (T, B, D)
`masks=None` is eval mode.
"""
import torch
import torch.nn.functional as F

NEG_FILL = -1e10  # editnet.py:374 (masked_fill value; not -inf)


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"])


def _drop(x, keep):
    return x if keep is None else x * keep.to(x.dtype) * 2.0


# --------------------------------------------------------------------------- cells
def embed(sd, tokens, keep=None):
    """EmbeddingC.forward, editnet.py:300-304: dropout(relu(Emb[x]))."""
    return _drop(torch.relu(sd["embed.embedding.weight"][tokens]), keep)


def lstm_cell_c(sd, prefix, x, h, c):
    """LSTMCellC.forward, editnet.py:226-244 (gate order i, f, g, o; :235)."""
    gates = _lin(sd, prefix + ".x2h", x) + _lin(sd, prefix + ".h2h", h)
    i, f, g, o = gates.chunk(4, 1)
    c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    return torch.sigmoid(o) * torch.tanh(c_new), c_new


def torch_lstm_cell(sd, prefix, x, h, c):
    """nn.LSTMCell (attention_lstm, editnet.py:468/532; DCNet dcnet.py:286-287)."""
    gates = (F.linear(x, sd[prefix + ".weight_ih"], sd[prefix + ".bias_ih"]) +
             F.linear(h, sd[prefix + ".weight_hh"], sd[prefix + ".bias_hh"]))
    i, f, g, o = gates.chunk(4, 1)
    c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
    return torch.sigmoid(o) * torch.tanh(c_new), c_new


def caption_encoder(sd, seq, seq_len, keep=None):
    """CaptionEncoderC.forward, editnet.py:319-348.

    The reference sorts by length, runs the cell on a shrinking batch and unsorts
    (:322-346).  Rows are independent, so the same values come from stepping every
    row and freezing it once t reaches its length -- which is what is done here.
    Returns (hidden_states (B,P',C), memory_states (B,P',C), final_hidden (B,C),
    mask (B,P')) with P' = max length in the batch (:327).
    """
    lens = seq_len.view(-1)
    B, Pmax = seq.shape[0], int(lens.max())
    emb = embed(sd, seq, keep)                                           # :331
    C = sd["caption_encoder.affine_hn.weight"].shape[0]
    h = emb.new_zeros(B, C)
    c = emb.new_zeros(B, C)
    hs, ms = [], []
    for t in range(Pmax):
        act = (lens > t).to(emb.dtype).unsqueeze(1)                      # :334
        hn, cn = lstm_cell_c(sd, "caption_encoder.lstm_encoder_cell", emb[:, t], h, c)
        h = act * hn + (1 - act) * h
        c = act * cn + (1 - act) * c
        hs.append(act * hn)                                              # :336
        ms.append(act * cn)                                              # :337
    hidden_states = torch.stack(hs, 1)
    memory_states = torch.stack(ms, 1)
    mask = (memory_states.sum(2) != 0).to(emb.dtype)                     # :340
    final_hidden = torch.tanh(_lin(sd, "caption_encoder.affine_hn", h))  # :341
    return hidden_states, memory_states, final_hidden, mask


def caption_attention(sd, prev_h, h1, word, mask):
    """CaptionAttentionC.forward, editnet.py:364-381.  Returns (gated ctx, alpha)."""
    p = "caption_attention."
    att1 = _lin(sd, p + "cap_features_att", prev_h)                      # :370
    att2 = _lin(sd, p + "cap_decoder_att", h1)                           # :371
    att = _lin(sd, p + "cap_full_att", torch.tanh(att1 + att2.unsqueeze(1))).squeeze(2)
    att = att.masked_fill(mask == 0, NEG_FILL)                           # :374
    alpha = F.softmax(att, dim=1)                                        # :375
    ctx = (prev_h * alpha.unsqueeze(2)).sum(1)                           # :376
    zt = torch.sigmoid(_lin(sd, p + "context_gate", torch.cat([word, h1, ctx], 1)))
    tc = torch.tanh(_lin(sd, p + "tc_affine", torch.cat([word, h1], 1)))
    sc = torch.tanh(_lin(sd, p + "sc_affine", ctx))
    return zt * sc + (1 - zt) * tc, alpha                                # :380


def select(prev_m, alpha):
    """SelectC.forward (hard path), editnet.py:409-421: one-hot at argmax(alpha)
    whose forward weight is alpha_max + (1 - detach(alpha_max)) (straight-through)."""
    a = alpha.detach()
    value, idx = a.max(1)
    onehot = torch.zeros_like(a).scatter_(1, idx.unsqueeze(1), 1.0)
    w = alpha * onehot + onehot * (1 - value).unsqueeze(1)               # :417-418
    return (w.unsqueeze(2) * prev_m).sum(1)                              # :420


def visual_attention(sd, feats, h1, keep=None, adaptive=False):
    """VisualAttentionC.forward, editnet.py:439-447; with `adaptive=True` the
    ragged-region variant adaptive_features/editnet_adaptive.py:438-457 (masks
    derived from all-zero rows, -1e10 fill, att_embed only over valid rows)."""
    p = "visual_attention."
    fe = torch.relu(_lin(sd, p + "att_embed.0", feats))
    fe = _drop(fe, keep)                                                 # :441
    if adaptive:
        valid = (feats.sum(2) != 0)                                      # adaptive:440
        fe = fe * valid.unsqueeze(2).to(fe.dtype)                        # pad rows of pad_packed are 0
        att_mask = (fe.sum(2) != 0)                                      # adaptive:449
    att1 = _lin(sd, p + "features_att", fe)                              # :442
    att2 = _lin(sd, p + "decoder_att", h1)                               # :443
    att = _lin(sd, p + "full_att", torch.relu(att1 + att2.unsqueeze(1))).squeeze(2)
    if adaptive:
        att = att.masked_fill(~att_mask, NEG_FILL)                       # adaptive:453
    alpha = F.softmax(att, dim=1)                                        # :445
    return (feats * alpha.unsqueeze(2)).sum(1)                           # :446


def copy_lstm(sd, x, h, c, c_mem):
    """CopyLSTMCellC.forward, editnet.py:265-285."""
    p = "copy_lstm."
    gates = _lin(sd, p + "x2h", x) + _lin(sd, p + "h2h", h)              # :272
    i, f, g, o = gates.chunk(4, 1)
    c_new = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)      # :280
    k = torch.sigmoid(_lin(sd, p + "gate_cnew", c_new) + _lin(sd, p + "gate_cmem", c_mem))
    c_out = k * c_mem + (1 - k) * c_new                                  # :282
    return torch.sigmoid(o) * torch.tanh(c_out), c_out                   # :283-285


def decoder_step(sd, emb, state, enc, feats, image_mean, vis_keep=None, adaptive=False):
    """One decode step on explicit state: editnet.py:527-543.
    state = (h1, c1, h2, c2); enc = (prev_h, prev_m, final_hidden, mask)."""
    h1, c1, h2, c2 = state
    prev_h, prev_m, final_hidden, mask = enc
    x1 = torch.cat([emb, final_hidden, h2, image_mean], 1)               # :527-530
    h1, c1 = torch_lstm_cell(sd, "attention_lstm", x1, h1, c1)           # :532
    att_cap, alpha_c = caption_attention(sd, prev_h, h1, emb, mask)      # :534
    att_img = visual_attention(sd, feats, h1, vis_keep, adaptive)        # :537
    sel = select(prev_m, alpha_c)                                        # :541
    h2, c2 = copy_lstm(sd, torch.cat([h1, att_cap, att_img], 1), h2, c2, sel)  # :543
    return (h1, c1, h2, c2), alpha_c


# ----------------------------------------------------------------- sequence drivers
def xe_forward(sd, feats, caps, caplens, prev, prev_len, masks=None, image_mean=None,
               want_trace=False, fed_tokens=None, stable_sort=False):
    """DecoderC.forward (teacher forced, use_ss=False), editnet.py:479-548; with
    `image_mean` given, the adaptive variant editnet_adaptive.py:489-562.

    Returns (predictions (B,maxT,V), caps_sorted, decode_lengths list, sort_ind)
    [+ trace dict].  Scheduled sampling (:508-520) draws from torch's RNG and is
    exercised statistically elsewhere.
    """
    adaptive = image_mean is not None
    B = caps.shape[0]
    # :488 -- the reference's sort is unstable: the order inside a group of equal lengths is unspecified
    # (and differs between CPU and CUDA); stable_sort=True pins it for comparisons at larger batch sizes
    lens, sort_ind = caplens.squeeze(1).sort(dim=0, descending=True, stable=stable_sort)
    feats, caps = feats[sort_ind], caps[sort_ind]
    prev, prev_len = prev[sort_ind], prev_len[sort_ind]
    decode_lengths = (lens - 1).tolist()                                 # :497
    T = max(decode_lengths)
    V = sd["fc.weight"].shape[0]
    D = sd["fc.weight"].shape[1]
    m = masks or {}
    enc = caption_encoder(sd, prev, prev_len, m.get("enc"))              # :501
    image_mean = image_mean[sort_ind] if adaptive else feats.mean(1)     # :503
    h1 = feats.new_zeros(B, D); c1 = feats.new_zeros(B, D)
    h2 = feats.new_zeros(B, D); c2 = feats.new_zeros(B, D)
    preds = feats.new_zeros(B, T, V)                                     # :499
    trace = {"h1": [], "c1": [], "h2": [], "c2": [], "alpha_c": []}
    for t in range(T):
        b = sum(l > t for l in decode_lengths)                           # :506
        # fed_tokens (sorted rows, (B, Wc)): replay of a scheduled-sampling run (:508-520) -- the tokens the
        # implementation actually fed; None = teacher forcing
        tok = caps[:b, t] if fed_tokens is None else fed_tokens[:b, t]
        e = embed(sd, tok, None if masks is None else m["emb"][t, :b])
        enc_b = tuple(x[:b] for x in enc)
        (h1, c1, h2, c2), alpha_c = decoder_step(
            sd, e, (h1[:b], c1[:b], h2[:b], c2[:b]), enc_b, feats[:b], image_mean[:b],
            None if masks is None else m["vis"][t, :b], adaptive)
        hd = _drop(h2, None if masks is None else m["fc"][t, :b])        # :545
        preds[:b, t] = _lin(sd, "fc", hd)                                # :546
        if want_trace:
            for k, v in (("h1", h1), ("c1", c1), ("h2", h2), ("c2", c2), ("alpha_c", alpha_c)):
                trace[k].append(v.detach().clone())
    out = (preds, caps, decode_lengths, sort_ind)
    if want_trace:
        trace["enc"] = tuple(x.detach().clone() for x in enc)
        return out + (trace,)
    return out


def xe_loss(preds, caps_sorted, decode_lengths):
    """train() loss, editnet.py:571-577: mean CE over the packed (sum decode_lengths)
    rows.  pack_padded_sequence only selects rows, so an explicit time-major gather
    gives the same mean."""
    rows, tgts = [], []
    for t in range(max(decode_lengths)):
        b = sum(l > t for l in decode_lengths)
        rows.append(preds[:b, t]); tgts.append(caps_sorted[:b, t + 1])   # :571
    return F.cross_entropy(torch.cat(rows, 0), torch.cat(tgts, 0))


def rollout(sd, prev, prev_len, feats, start_idx, end_idx, mode="greedy", masks=None,
            forced=None, max_len=18, image_mean=None):
    """DecoderC.forward (RL), editnet_rl.py:485-549.

    mode 'greedy' follows :521; mode 'forced' replays given tokens (`forced`,
    (B,max_len), 0 = finished) and returns their log-probs -- it is how a sampled
    rollout (:525-527, torch.multinomial, not bit-reproducible) is checked: the
    implementation's own samples are replayed here.  Returns (seq, seqLogprobs).
    Rows are NOT sorted in this path and the batch never shrinks (:503-547).
    """
    B = feats.shape[0]
    D = sd["fc.weight"].shape[1]
    m = masks or {}
    seq = torch.zeros(B, max_len, dtype=torch.long)
    slp = feats.new_zeros(B, max_len)
    it = torch.full((B,), start_idx, dtype=torch.long)                   # :493-495
    enc = caption_encoder(sd, prev, prev_len, m.get("enc"))              # :499
    im = feats.mean(1) if image_mean is None else image_mean             # :501
    st = tuple(feats.new_zeros(B, D) for _ in range(4))
    unfinished = None
    for t in range(max_len):   # the reference's 19th evaluation (:503,517) is discarded
        e = embed(sd, it, None if masks is None else m["emb"][t])
        st, _ = decoder_step(sd, e, st, enc, feats, im,
                             None if masks is None else m["vis"][t])
        hd = _drop(st[2], None if masks is None else m["fc"][t])         # :513
        logp = F.log_softmax(_lin(sd, "fc", hd), dim=1)                  # :514
        if mode == "greedy":
            lp, it = logp.max(1)                                         # :521
        else:
            raw = forced[:, t].clone()
            # a finished row was stored as 0; the model still saw 0 as its next input
            lp = logp.gather(1, raw.unsqueeze(1)).squeeze(1)             # :527
            it = raw
        it = it.clone()
        it[it == end_idx] = 0                                            # :532
        unfinished = (it > 0) if unfinished is None else unfinished & (it > 0)  # :535-538
        it = it * unfinished.to(it.dtype)                                # :540
        seq[:, t] = it                                                   # :542
        slp[:, t] = lp                                                   # :543
        if int(unfinished.sum()) == 0:                                   # :546
            break
    return seq, slp


def reward_criterion(sample_logprobs, seq, reward):
    """RewardCriterion.forward, editnet_rl.py:557-573."""
    mask = (seq > 0).to(sample_logprobs.dtype)
    mask = torch.cat([mask.new_ones(mask.size(0), 1), mask[:, :-1]], 1)  # :565
    return (-sample_logprobs * reward * mask).sum() / mask.sum()         # :571-572


def clip_and_adam(params, grads, exp_avg, exp_avg_sq, step, lr=5e-4, max_norm=0.25,
                  betas=(0.9, 0.999), eps=1e-8):
    """train() tail, editnet.py:580-581: clip_grad_norm_(0.25) then Adam(lr 5e-4,
    torch defaults).  In-place on the given lists of tensors; returns total norm."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    b1, b2 = betas
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = g * coef
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / (1 - b2 ** step) ** 0.5).add_(eps)
        p.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))
    return total


# ---------------------------------------------------------------------- parameters
def init_state_dict(V, D=1024, C=1024, E=1024, A=512, Fdim=2048, seed=0, dtype=torch.float32):
    """Random parameters with the reference's shapes/keys and init scales
    (DecoderC.__init__, editnet.py:451-471; LSTMCellC/CopyLSTMCellC uniform
    +-1/sqrt(hidden), :221-224/:260-263; torch defaults elsewhere)."""
    g = torch.Generator().manual_seed(seed)

    def U(shape, bound):
        return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)

    sd = {}
    sd["embed.embedding.weight"] = torch.randn(V, E, generator=g, dtype=torch.float64).to(dtype)

    def linear(name, out_f, in_f, bound=None):
        b = bound if bound is not None else 1.0 / in_f ** 0.5
        sd[name + ".weight"] = U((out_f, in_f), b)
        sd[name + ".bias"] = U((out_f,), b)

    linear("caption_encoder.lstm_encoder_cell.x2h", 4 * C, E, 1 / C ** 0.5)
    linear("caption_encoder.lstm_encoder_cell.h2h", 4 * C, C, 1 / C ** 0.5)
    linear("caption_encoder.affine_hn", C, C)
    linear("caption_attention.cap_features_att", A, C)
    linear("caption_attention.cap_decoder_att", A, D)
    linear("caption_attention.cap_full_att", 1, A)
    linear("caption_attention.context_gate", C, 2 * C + D)
    linear("caption_attention.sc_affine", C, C)
    linear("caption_attention.tc_affine", C, 2 * D)
    linear("visual_attention.att_embed.0", D, Fdim)
    linear("visual_attention.features_att", A, D)
    linear("visual_attention.decoder_att", A, D)
    linear("visual_attention.full_att", 1, A)
    k = 1 / D ** 0.5
    sd["attention_lstm.weight_ih"] = U((4 * D, 3 * E + Fdim), k)
    sd["attention_lstm.weight_hh"] = U((4 * D, D), k)
    sd["attention_lstm.bias_ih"] = U((4 * D,), k)
    sd["attention_lstm.bias_hh"] = U((4 * D,), k)
    linear("copy_lstm.x2h", 4 * D, 2 * E + Fdim, k)
    linear("copy_lstm.h2h", 4 * D, D, k)
    linear("copy_lstm.gate_cnew", D, D, k)
    linear("copy_lstm.gate_cmem", D, D, k)
    linear("fc", V, D)
    return sd


def beam_search(sd, word_map, feats, prev, prev_len, beam_size=3, max_steps=50):
    """evaluate()'s search loop, editnet.py:608-719, for one image (feats (1,R,F), prev (1,Wp),
    prev_len (1,1)), eval mode.  `top_k_words / vocab_size` (:666) is true division on torch >= 1.5 and
    crashes; the reference's intent (torch 1.2 integer division) is restated as //."""
    k = beam_size
    V, D = sd["fc.weight"].shape
    enc = caption_encoder(sd, prev, prev_len)                                              # :613
    feats_k = feats.expand(k, -1, -1)                                                       # :616
    enc_k = tuple(x.expand(k, *x.shape[1:]) for x in enc)                                   # :617-621
    words = torch.full((k,), word_map["<start>"], dtype=torch.long)
    seqs = words.unsqueeze(1)
    top = torch.zeros(k, 1)
    done_seqs, done_scores = [], []
    st = tuple(feats.new_zeros(k, D) for _ in range(4))
    step = 1
    runaway = False
    while True:
        n = words.shape[0]
        e = embed(sd, words)
        st, _ = decoder_step(sd, e, st, tuple(x[:n] for x in enc_k), feats_k[:n], feats_k[:n].mean(1))
        scores = F.log_softmax(_lin(sd, "fc", st[2]), dim=1)                                # :653-654
        scores = top.expand_as(scores) + scores
        if step == 1:
            top_s, top_w = scores[0].topk(k, 0, True, True)
        else:
            top_s, top_w = scores.view(-1).topk(k, 0, True, True)
        pi, ni = top_w // V, top_w % V
        seqs = torch.cat([seqs[pi], ni.unsqueeze(1)], 1)
        inc = [i for i, w in enumerate(ni.tolist()) if w != word_map["<end>"]]
        com = [i for i in range(len(ni)) if i not in inc]
        if com:
            done_seqs.extend(seqs[com].tolist())
            done_scores.extend(top_s[com].tolist())
        k -= len(com)
        if k == 0:
            break
        seqs = seqs[inc]
        st = tuple(x[pi[inc]] for x in st)
        top = top_s[inc].unsqueeze(1)
        words = ni[inc]
        if step > max_steps:                                                                # :702
            runaway = True
            break
        step += 1
    if runaway or not done_scores:
        return seqs[0][:18].tolist(), float(top[0])                                         # :710-711
    i = done_scores.index(max(done_scores))
    return done_seqs[i], done_scores[i]
