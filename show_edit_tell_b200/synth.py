"""Synthetic batches in the reference's data conventions: used by bench.py and
smoke().  Token layout of preprocess_caps.py:86-91,116-120 (<pad>=0, words 1..V-4, <unk>=V-3,
<start>=V-2, <end>=V-1; caption = <start> words <end> pads, caplen counts both markers); previous
captions have no markers and are zero padded (preprocess_existing_caps.py:23); bottom-up features
are post-ReLU, i.e. non-negative (bottom-up_features/tsv.py:56-68)."""
import torch


def word_map(V):
    wm = {"<pad>": 0}
    for i in range(1, V - 3):
        wm["w%d" % i] = i
    wm["<unk>"] = V - 3
    wm["<start>"] = V - 2
    wm["<end>"] = V - 1
    return wm


def make_batch(B, V, R=36, Fdim=2048, cap_width=20, prev_width=18, ragged=False, seed=0, pinned=False):
    g = torch.Generator().manual_seed(seed)
    n_words = V - 4
    caplens = (torch.randint(8, cap_width + 1, (B,), generator=g) if ragged
               else torch.full((B,), cap_width, dtype=torch.long))
    pos = torch.arange(cap_width).unsqueeze(0)
    words = torch.randint(1, n_words + 1, (B, cap_width), generator=g)
    caps = torch.where(pos < (caplens - 1).unsqueeze(1), words, torch.zeros_like(words))
    caps[:, 0] = V - 2
    caps[torch.arange(B), caplens - 1] = V - 1
    prev_len = torch.randint(5, prev_width + 1, (B,), generator=g)
    ppos = torch.arange(prev_width).unsqueeze(0)
    prev = torch.where(ppos < prev_len.unsqueeze(1), torch.randint(1, n_words + 1, (B, prev_width), generator=g),
                       torch.zeros(B, prev_width, dtype=torch.long))
    feats = torch.rand(B, R, Fdim, generator=g)
    out = dict(feats=feats, caps=caps, caplens=caplens.view(B, 1), prev=prev, prev_len=prev_len.view(B, 1))
    if pinned:
        out = {k: v.pin_memory() for k, v in out.items()}
    return out
