"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's DCNet (text-only denoising
auto-encoder) hot path: /root/reference/dcnet.py:147-350 (cells + teacher-forced DAE.forward) and
dcnet_rl.py:286-346 (rollout).  Same rules as oracle/editnet_oracle.py: pure functions over a
`state_dict`, fp32 torch ops, explicit dropout masks, every function cites the lines it follows;
pinned by tests/golden/dcnet_*.npz (written from the reference's real classes).

Dropout masks (0/1 keep flags, decoder-sorted row order for the XE path):
    masks['enc'] (B, Pw, E)   embed(src) inside the encoder, dcnet.py:225
    masks['emb'] (T, B, E)    embed of the fed tokens (dcnet.py:327 draws one (B,Wc,E) mask for all
                              positions at once; position t of row i is masks['emb'][t, i])
    masks['fc']  (T, B, D)    dropout before fc, dcnet.py:347
"""
import torch
import torch.nn.functional as F

from oracle.editnet_oracle import NEG_FILL, _drop, _lin, torch_lstm_cell


def embed(sd, tokens, keep=None):
    """Embedding.forward (load_glove_embedding=False), dcnet.py:199-206"""
    return _drop(torch.relu(sd["embed.embedding.weight"][tokens]), keep)


def caption_encoder(sd, src, src_len, keep=None):
    """CaptionEncoder.forward, dcnet.py:220-243: bidirectional nn.LSTM over the packed sequence.
    Restated per direction with a per-row active window (rows are independent): the forward direction
    walks t = 0..len-1, the reverse direction t = len-1..0; outputs are zero beyond len (pad_packed).
    Returns (outputs (B,P',2Cd), final_hidden (B,2Cd), mask (B,P'))."""
    lens = src_len.view(-1)
    B, Pmax = src.shape[0], int(lens.max())
    emb = embed(sd, src, keep)                                            # :225
    p = "caption_encoder.lstm_encoder."
    Cd = sd[p + "weight_hh_l0"].shape[1]
    outs = emb.new_zeros(B, Pmax, 2 * Cd)
    finals = []
    for d, suffix in enumerate(("", "_reverse")):
        w_ih, w_hh = sd[p + "weight_ih_l0" + suffix], sd[p + "weight_hh_l0" + suffix]
        b_ih, b_hh = sd[p + "bias_ih_l0" + suffix], sd[p + "bias_hh_l0" + suffix]
        h = emb.new_zeros(B, Cd)
        c = emb.new_zeros(B, Cd)
        out_d = [None] * Pmax
        for s in range(Pmax):
            pos = (lens - 1 - s).clamp_min(0) if d == 1 else torch.full_like(lens, s)
            act = (lens > s).to(emb.dtype).unsqueeze(1)
            x = emb[torch.arange(B), pos]
            gates = F.linear(x, w_ih, b_ih) + F.linear(h, w_hh, b_hh)
            i, f, g, o = gates.chunk(4, 1)
            cn = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
            hn = torch.sigmoid(o) * torch.tanh(cn)
            h = act * hn + (1 - act) * h
            c = act * cn + (1 - act) * c
            onehot = F.one_hot(pos, Pmax).to(emb.dtype) * act             # (B, Pmax)
            outs = outs + torch.cat([outs.new_zeros(B, Pmax, d * Cd), onehot.unsqueeze(2) * hn.unsqueeze(1),
                                     outs.new_zeros(B, Pmax, (1 - d) * Cd)], 2)
        finals.append(h)                                                   # hidden[0][-2], hidden[0][-1], :241
    mask = (outs.sum(2) != 0).to(emb.dtype)                               # :239
    final_hidden = torch.tanh(_lin(sd, "caption_encoder.concat", torch.cat(finals, 1)))   # :242
    return outs, final_hidden, mask


def caption_attention(sd, enc, h1, mask):
    """CaptionAttention.forward, dcnet.py:254-270 (no gating; returns the context only)"""
    p = "caption_attention."
    att1 = _lin(sd, p + "cap_features_att", enc)
    att2 = _lin(sd, p + "cap_decoder_att", h1)
    att = _lin(sd, p + "cap_full_att", torch.tanh(att1 + att2.unsqueeze(1))).squeeze(2)
    att = att.masked_fill(mask == 0, NEG_FILL)                            # :265
    alpha = F.softmax(att, dim=1)
    return (enc * alpha.unsqueeze(2)).sum(1)                              # :268


def decoder_step(sd, emb, state, enc):
    """dcnet.py:336-346"""
    h1, c1, h2, c2 = state
    outs, final_hidden, mask = enc
    h1, c1 = torch_lstm_cell(sd, "attention_lstm", torch.cat([emb, final_hidden, h2], 1), h1, c1)   # :340
    ctx = caption_attention(sd, outs, h1, mask)                                                      # :341
    h2, c2 = torch_lstm_cell(sd, "language_lstm", torch.cat([h1, ctx], 1), h2, c2)                   # :346
    return (h1, c1, h2, c2)


def xe_forward(sd, caps, caplens, prev, prev_len, masks=None):
    """DAE.forward, dcnet.py:303-350 -> (predictions, caps_sorted, decode_lengths, sort_ind)"""
    B = caps.shape[0]
    lens, sort_ind = caplens.squeeze(1).sort(dim=0, descending=True)      # :314
    caps, prev, prev_len = caps[sort_ind], prev[sort_ind], prev_len[sort_ind]
    decode_lengths = (lens - 1).tolist()                                  # :325
    T = max(decode_lengths)
    V, D = sd["fc.weight"].shape
    m = masks or {}
    enc = caption_encoder(sd, prev, prev_len, m.get("enc"))               # :331
    z = sd["fc.weight"].new_zeros(B, D)
    st = (z, z, z, z)
    preds = sd["fc.weight"].new_zeros(B, T, V)                            # :329
    for t in range(T):
        b = sum(l > t for l in decode_lengths)                            # :334
        e = embed(sd, caps[:b, t], None if masks is None else m["emb"][t, :b])   # :327 (sliced at :336)
        st = decoder_step(sd, e, tuple(x[:b] for x in st), tuple(x[:b] for x in enc))
        hd = _drop(st[2], None if masks is None else m["fc"][t, :b])      # :347
        preds[:b, t] = _lin(sd, "fc", hd)                                 # :348
    return preds, caps, decode_lengths, sort_ind


def rollout(sd, prev, prev_len, start_idx, end_idx, mode="greedy", masks=None, forced=None, max_len=18):
    """DAE.forward (RL), dcnet_rl.py:286-346; modes as in editnet_oracle.rollout"""
    B = prev.shape[0]
    D = sd["fc.weight"].shape[1]
    m = masks or {}
    seq = torch.zeros(B, max_len, dtype=torch.long)
    slp = sd["fc.weight"].new_zeros(B, max_len)
    it = torch.full((B,), start_idx, dtype=torch.long)
    enc = caption_encoder(sd, prev, prev_len, m.get("enc"))               # :303
    z = sd["fc.weight"].new_zeros(B, D)
    st = (z, z, z, z)
    unfinished = None
    for t in range(max_len):
        e = embed(sd, it, None if masks is None else m["emb"][t])
        st = decoder_step(sd, e, st, enc)
        hd = _drop(st[2], None if masks is None else m["fc"][t])
        logp = F.log_softmax(_lin(sd, "fc", hd), dim=1)                   # :313
        if mode == "greedy":
            lp, it = logp.max(1)                                          # :319
        else:
            it = forced[:, t].clone()
            lp = logp.gather(1, it.unsqueeze(1)).squeeze(1)               # :325
        it = it.clone()
        it[it == end_idx] = 0                                             # :330
        unfinished = (it > 0) if unfinished is None else unfinished & (it > 0)
        it = it * unfinished.to(it.dtype)
        seq[:, t] = it
        slp[:, t] = lp
        if int(unfinished.sum()) == 0:
            break
    return seq, slp


def init_state_dict(V, D=1024, Cd=512, E=1024, A=512, seed=0, dtype=torch.float32):
    """parameters with the reference's keys/shapes (DAE.__init__, dcnet.py:275-295), torch default inits"""
    g = torch.Generator().manual_seed(seed)

    def U(shape, bound):
        return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)

    sd = {"embed.embedding.weight": torch.randn(V, E, generator=g, dtype=torch.float64).to(dtype)}
    k = 1 / D ** 0.5
    for name, inp in (("attention_lstm", 3 * E), ("language_lstm", 2 * E)):
        sd[name + ".weight_ih"] = U((4 * D, inp), k)
        sd[name + ".weight_hh"] = U((4 * D, D), k)
        sd[name + ".bias_ih"] = U((4 * D,), k)
        sd[name + ".bias_hh"] = U((4 * D,), k)
    kc = 1 / Cd ** 0.5
    for suffix in ("", "_reverse"):
        p = "caption_encoder.lstm_encoder."
        sd[p + "weight_ih_l0" + suffix] = U((4 * Cd, E), kc)
        sd[p + "weight_hh_l0" + suffix] = U((4 * Cd, Cd), kc)
        sd[p + "bias_ih_l0" + suffix] = U((4 * Cd,), kc)
        sd[p + "bias_hh_l0" + suffix] = U((4 * Cd,), kc)

    def linear(name, out_f, in_f):
        b = 1.0 / in_f ** 0.5
        sd[name + ".weight"] = U((out_f, in_f), b)
        sd[name + ".bias"] = U((out_f,), b)

    linear("caption_encoder.concat", 2 * Cd, 2 * Cd)
    linear("caption_attention.cap_features_att", A, 2 * Cd)
    linear("caption_attention.cap_decoder_att", A, D)
    linear("caption_attention.cap_full_att", 1, A)
    linear("fc", V, D)
    return sd
