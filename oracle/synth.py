"""TEST INFRASTRUCTURE -- seeded synthetic inputs in the reference's data conventions
(SURVEY.md §8d).  Shared by the golden-vector generator and the tests so both see
the same tensors.  (bench.py uses the product-side twin in
`show_edit_tell_b200/synth.py`; this copy exists so `oracle/` stays self-contained.)

Token layout follows preprocess_caps.py:86-91,116-120: <pad>=0, words 1..V-4,
<unk>=V-3, <start>=V-2, <end>=V-1; a caption row is <start> + words + <end> + pads,
caplen counts <start> and <end>.  Previous captions carry no <start>/<end> and are
zero padded to `prev_width` (preprocess_existing_caps.py:23).
"""
import torch


def word_map(V):
    wm = {"<pad>": 0}
    for i in range(1, V - 3):
        wm["w%d" % i] = i
    wm["<unk>"] = V - 3
    wm["<start>"] = V - 2
    wm["<end>"] = V - 1
    return wm


def make_batch(B, V, R=36, Fdim=2048, cap_width=20, prev_width=18, ragged=True, seed=0,
               min_len=None, min_prev=None, adaptive=False, Rmin=None):
    g = torch.Generator().manual_seed(seed)
    n_words = V - 4
    min_len = min_len if min_len is not None else min(8, cap_width)
    min_prev = min_prev if min_prev is not None else min(5, prev_width)
    if ragged:
        caplens = torch.randint(min_len, cap_width + 1, (B,), generator=g)
    else:
        caplens = torch.full((B,), cap_width, dtype=torch.long)
    caps = torch.zeros(B, cap_width, dtype=torch.long)
    for i in range(B):
        L = int(caplens[i])
        caps[i, 0] = V - 2
        caps[i, 1:L - 1] = torch.randint(1, n_words + 1, (L - 2,), generator=g)
        caps[i, L - 1] = V - 1
    prev_len = torch.randint(min_prev, prev_width + 1, (B,), generator=g)
    prev = torch.zeros(B, prev_width, dtype=torch.long)
    for i in range(B):
        L = int(prev_len[i])
        prev[i, :L] = torch.randint(1, n_words + 1, (L,), generator=g)
    feats = torch.rand(B, R, Fdim, generator=g)
    out = dict(feats=feats, caps=caps, caplens=caplens.view(B, 1), prev=prev,
               prev_len=prev_len.view(B, 1))
    if adaptive:
        Rmin = Rmin if Rmin is not None else max(1, R // 10)
        nreg = torch.randint(Rmin, R + 1, (B,), generator=g)
        for i in range(B):
            feats[i, int(nreg[i]):] = 0
        out["nreg"] = nreg
        out["image_mean"] = torch.stack(
            [feats[i, :int(nreg[i])].mean(0) for i in range(B)], 0)
    return out


def make_masks(B, T, Pw, E, D, R, seed=0):
    """0/1 keep flags for the four dropout sites (layout: editnet_oracle docstring)."""
    g = torch.Generator().manual_seed(1000 + seed)

    def bern(*shape):
        return (torch.rand(*shape, generator=g) < 0.5).float()

    return {"enc": bern(B, Pw, E), "emb": bern(T, B, E),
            "vis": bern(T, B, R, D), "fc": bern(T, B, D)}
