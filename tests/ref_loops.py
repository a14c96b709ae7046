"""TEST INFRASTRUCTURE: the reference's beam-search loops, restated against the decoder's SUB-MODULE surface
 -- the calls evaluate() makes at editnet.py:613 and :645-653, in the reference's order and with the
reference's tensor plumbing (expand to k beams, re-gather every per-beam tensor by `prev_word_inds`).  The reference
file cannot travel to the GPU box, so the `-m gpu` test drives THIS loop against the CUDA modules and compares with the
captions the reference's own `evaluate` (AST-extracted, oracle/make_golden_evaluate.py) produced; a CPU test
(tests/test_ref_loops_vs_reference.py) checks, where /root/reference exists, that this restatement and the
AST-extracted original return the same captions when both drive the reference's own modules."""
import torch
import torch.nn.functional as F


def evaluate_one(decoder, word_map, img, previous_caption, prev_caplen, beam_size, vocab_size, max_steps=50):
    """one iteration of evaluate()'s per-image loop (editnet.py:602-713) -> token ids of the caption (no <start>/<end>/<pad>)"""
    k = beam_size
    device = img.device
    infinite_pred = False
    image_features = img                                                                    # :608
    img_mean = image_features.mean(1)                                                       # :612
    previous_encoded_h, previous_encoded_m, final_hidden, prev_cap_mask = decoder.caption_encoder(
        previous_caption, prev_caplen)                                                      # :613
    image_features = image_features.expand(k, -1, -1)                                       # :616-621
    img_mean = img_mean.expand(k, -1)
    previous_encoded_h = previous_encoded_h.expand(k, -1, -1)
    previous_encoded_m = previous_encoded_m.expand(k, -1, -1)
    final_hidden = final_hidden.expand(k, -1)
    prev_cap_mask = prev_cap_mask.expand(k, -1)
    k_prev_words = torch.LongTensor([[word_map['<start>']]] * k).to(device)                 # :624
    seqs = k_prev_words
    top_k_scores = torch.zeros(k, 1).to(device)
    complete_seqs, complete_seqs_scores = [], []
    step = 1
    h1, c1 = decoder.init_hidden_state(k)                                                   # :639-640
    h2, c2 = decoder.init_hidden_state(k)
    while True:
        embeddings = decoder.embed(k_prev_words).squeeze(1)                                 # :645
        topdown_input = torch.cat([embeddings, final_hidden, h2, img_mean], dim=1)          # :646
        h1, c1 = decoder.attention_lstm(topdown_input, (h1, c1))                            # :647
        attend_cap, alpha_c = decoder.caption_attention(previous_encoded_h, h1, embeddings, prev_cap_mask)   # :648
        attend_img = decoder.visual_attention(image_features, h1)                           # :649
        language_input = torch.cat([h1, attend_cap, attend_img], dim=1)                     # :650
        selected_memory = decoder.select(previous_encoded_m, alpha_c)                       # :651
        h2, c2 = decoder.copy_lstm(language_input, (h2, c2), selected_memory)               # :652
        scores = decoder.fc(h2)                                                             # :653
        scores = F.log_softmax(scores, dim=1)                                               # :654
        scores = top_k_scores.expand_as(scores) + scores                                    # :657
        if step == 1:
            top_k_scores, top_k_words = scores[0].topk(k, 0, True, True)                    # :661
        else:
            top_k_scores, top_k_words = scores.view(-1).topk(k, 0, True, True)              # :664
        prev_word_inds = top_k_words // vocab_size                                          # :667 (`/` in the reference)
        next_word_inds = top_k_words % vocab_size
        seqs = torch.cat([seqs[prev_word_inds], next_word_inds.unsqueeze(1)], dim=1)        # :671
        incomplete_inds = [ind for ind, next_word in enumerate(next_word_inds) if next_word != word_map['<end>']]
        complete_inds = list(set(range(len(next_word_inds))) - set(incomplete_inds))
        if len(complete_inds) > 0:
            complete_seqs.extend(seqs[complete_inds].tolist())
            complete_seqs_scores.extend(top_k_scores[complete_inds])
        k -= len(complete_inds)
        if k == 0:
            break
        seqs = seqs[incomplete_inds]
        sel = prev_word_inds[incomplete_inds]
        h1, c1, h2, c2 = h1[sel], c1[sel], h2[sel], c2[sel]                                 # :688-691
        image_features = image_features[sel]
        img_mean = img_mean[sel]
        final_hidden = final_hidden[sel]
        previous_encoded_h = previous_encoded_h[sel]
        previous_encoded_m = previous_encoded_m[sel]
        prev_cap_mask = prev_cap_mask[sel]
        top_k_scores = top_k_scores[incomplete_inds].unsqueeze(1)
        k_prev_words = next_word_inds[incomplete_inds].unsqueeze(1)
        if step > max_steps:                                                                # :702
            infinite_pred = True
            break
        step += 1
    if infinite_pred is not True:
        i = complete_seqs_scores.index(max(complete_seqs_scores))                           # :707
        seq = complete_seqs[i]
    else:
        seq = [int(x) for x in seqs[0][:18]]                                                # :710-711
    return [w for w in seq if w not in {word_map['<start>'], word_map['<end>'], word_map['<pad>']}]   # :714


def evaluate_full_one(dae_ar, decoder, word_map, img, previous_caption, prev_caplen, beam_size, max_steps=50):
    """one iteration of evaluate_full()'s per-image loop, eval/eval xe/eval_full.py:97-207 (EditNet + DCNet ensemble:
    the step score is log((softmax_e + softmax_d) / 2), :151-153) -> token ids of the caption"""
    k = beam_size
    device = img.device
    vocab_size = len(word_map)
    infinite_pred = False
    dae = dae_ar.dae
    image_features = img
    img_mean = image_features.mean(1)
    eh, em, efh, emask = decoder.caption_encoder(previous_caption, prev_caplen)             # :107
    denc, dfh, dmask = dae.caption_encoder(previous_caption, prev_caplen)                   # :109
    image_features = image_features.expand(k, -1, -1)
    img_mean = img_mean.expand(k, -1)
    eh, em, efh, emask = eh.expand(k, -1, -1), em.expand(k, -1, -1), efh.expand(k, -1), emask.expand(k, -1)
    denc, dmask, dfh = denc.expand(k, -1, -1), dmask.expand(k, -1), dfh.expand(k, -1)
    k_prev_words = torch.LongTensor([[word_map['<start>']]] * k).to(device)
    seqs = k_prev_words
    top_k_scores = torch.zeros(k, 1).to(device)
    complete_seqs, complete_seqs_scores = [], []
    step = 1
    eh1, ec1 = decoder.init_hidden_state(k)
    eh2, ec2 = decoder.init_hidden_state(k)
    dh1, dc1 = dae.init_hidden_state(k)
    dh2, dc2 = dae.init_hidden_state(k)
    while True:
        eemb = decoder.embed(k_prev_words).squeeze(1)                                       # :133-141
        eh1, ec1 = decoder.attention_lstm(torch.cat([eemb, efh, eh2, img_mean], dim=1), (eh1, ec1))
        eattend_cap, ealpha_c = decoder.caption_attention(eh, eh1, eemb, emask)
        eattend_img = decoder.visual_attention(image_features, eh1)
        esel = decoder.select(em, ealpha_c)
        eh2, ec2 = decoder.copy_lstm(torch.cat([eh1, eattend_cap, eattend_img], dim=1), (eh2, ec2), esel)
        escores = decoder.fc(eh2)
        demb = dae.embed(k_prev_words).squeeze(1)                                           # :143-149
        dh1, dc1 = dae.attention_lstm(torch.cat([demb, dfh, dh2], dim=1), (dh1, dc1))
        dattend_cap = dae.caption_attention(denc, dh1, dmask)
        dh2, dc2 = dae.language_lstm(torch.cat([dh1, dattend_cap], dim=1), (dh2, dc2))
        dscores = dae.fc(dh2)
        scores = torch.log((F.softmax(escores, dim=1) + F.softmax(dscores, dim=1)) / 2)     # :151-153
        scores = top_k_scores.expand_as(scores) + scores
        if step == 1:
            top_k_scores, top_k_words = scores[0].topk(k, 0, True, True)
        else:
            top_k_scores, top_k_words = scores.view(-1).topk(k, 0, True, True)
        prev_word_inds = top_k_words // vocab_size                                          # :162 (`/` in the reference)
        next_word_inds = top_k_words % vocab_size
        seqs = torch.cat([seqs[prev_word_inds], next_word_inds.unsqueeze(1)], dim=1)
        incomplete_inds = [ind for ind, next_word in enumerate(next_word_inds) if next_word != word_map['<end>']]
        complete_inds = list(set(range(len(next_word_inds))) - set(incomplete_inds))
        if len(complete_inds) > 0:
            complete_seqs.extend(seqs[complete_inds].tolist())
            complete_seqs_scores.extend(top_k_scores[complete_inds])
        k -= len(complete_inds)
        if k == 0:
            break
        seqs = seqs[incomplete_inds]
        sel = prev_word_inds[incomplete_inds]
        eh1, ec1, eh2, ec2 = eh1[sel], ec1[sel], eh2[sel], ec2[sel]
        image_features, img_mean, efh = image_features[sel], img_mean[sel], efh[sel]
        eh, em, emask = eh[sel], em[sel], emask[sel]
        dh1, dc1, dh2, dc2 = dh1[sel], dc1[sel], dh2[sel], dc2[sel]
        dfh, denc, dmask = dfh[sel], denc[sel], dmask[sel]
        top_k_scores = top_k_scores[incomplete_inds].unsqueeze(1)
        k_prev_words = next_word_inds[incomplete_inds].unsqueeze(1)
        if step > max_steps:                                                                # :198
            infinite_pred = True
            break
        step += 1
    if infinite_pred is not True:
        i = complete_seqs_scores.index(max(complete_seqs_scores))
        seq = complete_seqs[i]
    else:
        seq = [int(x) for x in seqs[0][:18]]
    return [w for w in seq if w not in {word_map['<start>'], word_map['<end>'], word_map['<pad>']}]
