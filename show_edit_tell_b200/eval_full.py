"""EditNet + DCNet ensemble beam search: the search loop of `evaluate_full`, eval/eval xe/eval_full.py:97-199
(identical in eval/eval rl/eval_full.py), for one image.

Both networks advance the same k live beams; a step's score is log((softmax(EditNet) + softmax(DCNet)) / 2)
(eval_full.py:151-153).  Each network's step is ONE library call on explicit state (set_editnet_step /
set_dcnet_step) instead of the reference's eight / six module calls; `top_k_words // vocab_size` replaces the
reference's `/` (true division since torch 1.5).  The COCO scoring tail of evaluate_full
(Java tokenizer / METEOR) is outside the path.
"""
import torch


def beam_search_ensemble(decoder, dae, word_map, image_features, encoded_previous_caption, previous_cap_length,
                         beam_size=3, max_steps=50, return_all=False):
    """decoder: EditNet `DecoderC`; dae: DCNet `DAE` (or a `DAEWithAR`, whose `.dae` is used, eval_full.py:108).
    image_features (1,R,F), encoded_previous_caption (1,Wp), previous_cap_length (1,1).
    Returns (token list incl. <start>/<end>, score); on the 50-step runaway guard, the first 18 tokens of the best
    live beam (eval_full.py:205-207)."""
    dae = getattr(dae, "dae", dae)
    k = beam_size
    V = decoder.vocab_size
    dev = image_features.device
    prev_k = encoded_previous_caption.expand(k, -1)
    len_k = previous_cap_length.expand(k, -1)
    esess = decoder.step_session(image_features.expand(k, -1, -1), prev_k, len_k)            # :107-120
    dsess = dae.step_session(prev_k, len_k)
    k_prev_words = torch.full((k,), word_map['<start>'], dtype=torch.long, device=dev)       # :122
    seqs = k_prev_words.unsqueeze(1)
    top_k_scores = torch.zeros(k, 1, device=dev)
    complete_seqs, complete_scores = [], []
    estate, dstate = esess.init_state(), dsess.init_state()                                   # :127-130
    step = 1
    runaway = False
    while True:
        escores, estate = esess.step(k_prev_words, estate)                                    # :133-141
        dscores, dstate = dsess.step(k_prev_words, dstate)                                    # :143-149
        scores = ((torch.softmax(escores, dim=1) + torch.softmax(dscores, dim=1)) / 2).log()  # :151-153
        scores = top_k_scores.expand_as(scores) + scores                                      # :155
        if step == 1:
            top_k_scores, top_k_words = scores[0].topk(k, 0, True, True)                      # :157
        else:
            top_k_scores, top_k_words = scores.view(-1).topk(k, 0, True, True)                # :160
        prev_word_inds = top_k_words // V                                                     # :162
        next_word_inds = top_k_words % V
        seqs = torch.cat([seqs[prev_word_inds], next_word_inds.unsqueeze(1)], dim=1)          # :164
        nxt = next_word_inds.tolist()
        incomplete = [i for i, w in enumerate(nxt) if w != word_map['<end>']]                 # :165-166
        complete = [i for i in range(len(nxt)) if i not in incomplete]
        if complete:
            complete_seqs.extend(seqs[complete].tolist())                                     # :169-170
            complete_scores.extend(top_k_scores[complete].tolist())
        k -= len(complete)                                                                    # :171
        if k == 0:
            break
        inc = torch.tensor(incomplete, device=dev, dtype=torch.long)
        seqs = seqs[inc]
        sel = prev_word_inds[inc]
        estate = tuple(x[sel] for x in estate)                                                # :177-180
        dstate = tuple(x[sel] for x in dstate)                                                # :188-191
        top_k_scores = top_k_scores[inc].unsqueeze(1)
        k_prev_words = next_word_inds[inc]
        if step > max_steps:                                                                  # :198-200
            runaway = True
            break
        step += 1
    if runaway or not complete_scores:
        best = (seqs[0][:18].tolist(), float(top_k_scores[0]))
    else:
        i = complete_scores.index(max(complete_scores))                                       # :203-204
        best = (complete_seqs[i], complete_scores[i])
    if return_all:   # + every completed beam in completion order (tests)
        return best + (complete_seqs, complete_scores)
    return best
