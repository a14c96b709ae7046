"""CPU restatement of the EditNet + DCNet ensemble beam search (test infrastructure only).

Follows `evaluate_full`, eval/eval xe/eval_full.py:97-207, for one image: both networks step the same beams and the
step score is log((softmax_e + softmax_d) / 2) (:151-153).  `top_k_words / vocab_size` (:162) is true division on
torch >= 1.5 and crashes there; the reference's intent (torch 1.2 integer division) is restated as //.
Pinned by tests/test_oracle_vs_reference.py::test_ensemble_beam_live against the reference's own function (AST-extracted,
`/` -> `//` on that one statement, COCO scoring tail cut) and by tests/golden/ensemble_beam.npz.
"""
import torch
import torch.nn.functional as F

from . import dcnet_oracle as DO
from . import editnet_oracle as EO


def beam_search_ensemble(sd_e, sd_d, word_map, feats, prev, prev_len, beam_size=3, max_steps=50, return_all=False):
    k = beam_size
    V, D = sd_e["fc.weight"].shape
    enc_e = EO.caption_encoder(sd_e, prev, prev_len)                                        # :107
    enc_d = DO.caption_encoder(sd_d, prev, prev_len)                                        # :108-109
    feats_k = feats.expand(k, -1, -1)
    enc_e = tuple(x.expand(k, *x.shape[1:]) for x in enc_e)
    enc_d = tuple(x.expand(k, *x.shape[1:]) for x in enc_d)
    words = torch.full((k,), word_map["<start>"], dtype=torch.long)
    seqs = words.unsqueeze(1)
    top = torch.zeros(k, 1)
    done_seqs, done_scores = [], []
    st_e = tuple(feats.new_zeros(k, D) for _ in range(4))
    st_d = tuple(feats.new_zeros(k, D) for _ in range(4))
    step = 1
    runaway = False
    while True:
        n = words.shape[0]
        st_e, _ = EO.decoder_step(sd_e, EO.embed(sd_e, words), st_e, tuple(x[:n] for x in enc_e), feats_k[:n],
                                  feats_k[:n].mean(1))                                      # :133-141
        st_d = DO.decoder_step(sd_d, DO.embed(sd_d, words), st_d, tuple(x[:n] for x in enc_d))   # :143-148
        es = F.linear(st_e[2], sd_e["fc.weight"], sd_e["fc.bias"])
        ds = F.linear(st_d[2], sd_d["fc.weight"], sd_d["fc.bias"])                          # :149
        scores = ((F.softmax(es, dim=1) + F.softmax(ds, dim=1)) / 2).log()                  # :151-153
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
        st_e = tuple(x[pi[inc]] for x in st_e)
        st_d = tuple(x[pi[inc]] for x in st_d)
        top = top_s[inc].unsqueeze(1)
        words = ni[inc]
        if step > max_steps:
            runaway = True
            break
        step += 1
    if runaway or not done_scores:
        best = (seqs[0][:18].tolist(), float(top[0]))
    else:
        i = done_scores.index(max(done_scores))
        best = (done_seqs[i], done_scores[i])
    if return_all:   # every completed beam in completion order: lets tests see the steps behind the winner
        return best + (done_seqs, done_scores)
    return best
