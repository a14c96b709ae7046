"""Known-answer vectors for the CIDEr-D scorer, worked by hand from its published definition (Vedantam et al. 2015;
the "-D" variant as implemented in pyciderevalcap/ciderD/ciderD_scorer.py, the un-vendored package the reference calls
at editnet_rl.py:584,637): tf-idf n-gram vectors (n = 1..4) with idf = ln(N) - ln(max(1, df)), CLIPPED cosine
sum_g min(h_g, r_g) r_g / (|h| |r|), Gaussian penalty exp(-delta^2 / (2 sigma^2)), sigma = 6, on the difference delta of
the captions' BIGRAM counts, mean over n, mean over references, x 10.

Corpus: N = 4 documents; document frequencies  a: 2,  b: 1,  (a, b): 1,  every other n-gram: 0 (so max(1, df) = 1).
With L = ln 2: idf(a) = ln 4 - ln 2 = L, and idf(g) = ln 4 = 2L for every other n-gram g.

Case 1: hypothesis "a b", reference "a b".  Identical tf-idf vectors: cosine 1 for n = 1, 2; no 3-/4-grams (0);
  delta = 0.  Score = 10 (1 + 1 + 0 + 0) / 4 = 5.

Case 2: hypothesis "a b b", reference "a b".
  n=1: h = (a: L, b: 2*2L = 4L), r = (a: L, b: 2L).  clipped dot = L*L + min(4L, 2L)*2L = 5 L^2;
       |h| = L sqrt(17), |r| = L sqrt(5)  ->  5 / sqrt(85) = sqrt(5/17).
  n=2: h = ((a,b): 2L, (b,b): 2L), r = ((a,b): 2L).  clipped dot = 4 L^2; |h| = 2L sqrt(2), |r| = 2L  ->  1/sqrt(2).
  n=3: h has (a,b,b), r has none -> 0.   n=4: 0.
  delta = 2 - 1 = 1 bigram  ->  penalty exp(-1/72).
  Score = 10 (sqrt(5/17) + 1/sqrt(2)) / 4 * exp(-1/72) = 3.0804990...

Cases 3, 4: the same captions as token ids through the reward glue of editnet_rl.py:587-646, where the <end> token is
kept as the word "0" (array_to_str stops BEHIND the first 0; preprocess_gd maps <end> -> 0):
  3: "a b 0" vs "a b 0": cosine 1 for n = 1, 2, 3, no 4-gram  ->  10 * 3/4 = 7.5.
  4: "a b b 0" vs "a b 0" (idf("0") = 2L):
     n=1: h = (L, 4L, 2L), r = (L, 2L, 2L): dot = L^2 + 4L^2 + 4L^2 = 9L^2; |h| = L sqrt(21), |r| = 3L -> 3/sqrt(21)
     n=2: h = {(a,b), (b,b), (b,0)} x 2L, r = {(a,b), (b,0)} x 2L: dot = 8L^2; |h| = 2L sqrt(3), |r| = 2L sqrt(2) -> 2/sqrt(6)
     n=3: {(a,b,b), (b,b,0)} vs {(a,b,0)}: 0.   n=4: 0.   delta = 3 - 2 = 1.
     Score = 10 (3/sqrt(21) + 2/sqrt(6)) / 4 * exp(-1/72) = 3.6271472...
The CPU half pins the oracle's scorer; the -m gpu half pins the device kernel to the same hand-worked numbers."""
import math

import numpy as np
import pytest
import torch

from oracle import ciderd_oracle as CO
from oracle import synth

DF = {("a",): 2.0, ("b",): 1.0, ("a", "b"): 1.0}
N_DOCS = 4.0
CASE2 = 10.0 * (math.sqrt(5.0 / 17.0) + 1.0 / math.sqrt(2.0)) / 4.0 * math.exp(-1.0 / 72.0)
CASE4 = 10.0 * (3.0 / math.sqrt(21.0) + 2.0 / math.sqrt(6.0)) / 4.0 * math.exp(-1.0 / 72.0)


def test_oracle_scorer_reproduces_the_hand_worked_scores():
    assert abs(CASE2 - 3.0804990) < 1e-6 and abs(CASE4 - 3.6271472) < 1e-6      # the arithmetic of the docstring
    sc = CO.CiderD(DF, N_DOCS)
    for hyp, ref, want in (("a b", "a b", 5.0), ("a b b", "a b", CASE2)):
        _, scores = sc.compute_score({0: [ref]}, [{"image_id": 0, "caption": [hyp]}])
        assert abs(scores[0] - want) < 1e-12, (hyp, ref, scores[0], want)
    # averaged over two references: (5 + CASE2') / 2 with the roles swapped for the second one is NOT symmetric
    # (clipping): "a b" against reference "a b b" -> n=1: dot = L*L + min(2L,4L)*4L = 9L^2 / (L sqrt5 * L sqrt17);
    # n=2: dot = 4L^2 / (2L * 2L sqrt2) = 1/sqrt2; delta = -1
    swapped = 10.0 * (9.0 / math.sqrt(85.0) + 1.0 / math.sqrt(2.0)) / 4.0 * math.exp(-1.0 / 72.0)
    _, scores = sc.compute_score({0: ["a b", "a b b"]}, [{"image_id": 0, "caption": ["a b"]}])
    assert abs(scores[0] - (5.0 + swapped) / 2.0) < 1e-12


@pytest.mark.gpu
def test_device_kernel_reproduces_the_hand_worked_scores():
    from show_edit_tell_b200 import ciderd
    V = 20
    wm = synth.word_map(V)
    a, b = 3, 7
    table = ciderd.CiderDTable({(a,): 2.0, (b,): 1.0, (a, b): 1.0}, N_DOCS, "cuda")
    L = 6
    gen = torch.zeros(2, L, dtype=torch.long)
    gre = torch.zeros(2, L, dtype=torch.long)
    gen[0, :2] = torch.tensor([a, b])            # "a b 0"    (case 3)
    gre[0, :3] = torch.tensor([a, b, b])         # "a b b 0"  (case 4)
    gen[1, :3] = torch.tensor([a, b, b])
    gre[1, :2] = torch.tensor([a, b])
    allcaps = torch.zeros(2, 1, 8, dtype=torch.long)
    allcaps[:, 0, :4] = torch.tensor([wm["<start>"], a, b, wm["<end>"]])     # reference "a b 0"
    rewards, scores = ciderd.self_critical_reward(gen.cuda(), gre.cuda(), allcaps.cuda(), wm, table, return_scores=True)
    s = scores.cpu().double().numpy()            # [sample 0, sample 1, greedy 0, greedy 1]
    want = np.array([7.5, CASE4, CASE4, 7.5])
    assert np.abs(s - want).max() < 1e-5, (s, want)
    r = rewards.cpu().numpy()
    assert np.allclose(r[0], 7.5 - CASE4, atol=1e-5) and np.allclose(r[1], CASE4 - 7.5, atol=1e-5)
