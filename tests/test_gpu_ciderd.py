"""Device CIDEr-D reward against the CPU oracle (oracle/ciderd_oracle.py: n-gram/df half and reward glue pinned to the
reference, scorer restated from pyciderevalcap's published algorithm)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ciderd_oracle as CO
from oracle import synth

pytestmark = pytest.mark.gpu


def _case(V, B, L, seed, n_images):
    wm = synth.word_map(V)
    table = CO.synthetic_table(V, n_images=n_images, seed=seed)
    rng = np.random.RandomState(seed + 1)

    def seqs():
        out = np.zeros((B, L), dtype=np.int64)
        for i in range(B):
            n = rng.randint(0, L + 1)
            out[i, :n] = np.minimum(rng.zipf(1.3, size=n), V - 4)
        return out

    gen, gre = seqs(), seqs()
    gre[0] = gen[0]                                  # identical captions: reward exactly 0
    allcaps = np.zeros((B, 5, 20), dtype=np.int64)
    for i in range(B):
        for r in range(5):
            n = rng.randint(3, 18)
            allcaps[i, r, 0] = wm["<start>"]
            allcaps[i, r, 1:1 + n] = np.minimum(rng.zipf(1.3, size=n), V - 4)
            allcaps[i, r, 1 + n] = wm["<end>"]
        allcaps[i, 0, 1:1 + min(L, 17)] = np.where(gen[i, :min(L, 17)] > 0, gen[i, :min(L, 17)], 1)   # a near match
    return wm, table, gen, gre, allcaps


@pytest.mark.parametrize("V,B,L,seed,n_images", [(40, 6, 18, 2, 80), (1003, 64, 18, 5, 400)])
def test_ciderd_reward_vs_oracle(V, B, L, seed, n_images):
    from show_edit_tell_b200 import ciderd
    wm, table, gen, gre, allcaps = _case(V, B, L, seed, n_images)
    scorer = CO.CiderD(table["document_frequency"], table["ref_len"])
    gd = CO.preprocess_gd(torch.from_numpy(allcaps), wm)
    ref = CO.self_critical_reward(scorer, gen, gre, gd)
    df = {tuple(int(w) for w in ng): c for ng, c in table["document_frequency"].items()}
    dev_table = ciderd.CiderDTable(df, table["ref_len"], "cuda")
    rew, scores = ciderd.self_critical_reward(torch.from_numpy(gen).cuda(), torch.from_numpy(gre).cuda(),
                                              torch.from_numpy(allcaps).cuda(), wm, dev_table, return_scores=True)
    assert rew.shape == (B, L)
    assert np.abs(ref).max() > 0.1                              # a non-degenerate case
    assert np.allclose(rew.cpu().numpy(), ref, atol=1e-5)
    assert float(rew[0].abs().max()) == 0.0
    assert bool((scores >= 0).all())
