"""Pins of oracle/ciderd_oracle.py that the reference tree allows: the n-gram / document-frequency half against
preprocess_rl.py and the reward glue against editnet_rl.py:587-646 (both AST-extracted, unmodified).  The scorer
itself (pyciderevalcap) is absent: see the oracle's header ("parity unpinned" for that layer)."""
import ast
import os
from collections import OrderedDict, defaultdict

import numpy as np
import pytest
import torch

from oracle import ciderd_oracle as CO
from oracle import ref_extract as RX
from oracle import synth

pytestmark = pytest.mark.skipif(not RX.reference_available(), reason="reference tree not present")


def _functions(rel, names, ns):
    with open(os.path.join(RX.REFERENCE_ROOT, rel)) as f:
        tree = ast.parse(f.read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    assert len(body) == len(names)
    exec(compile(ast.Module(body=body, type_ignores=[]), rel, "exec"), ns)
    return ns


def test_ngram_and_document_frequency_match_preprocess_rl():
    ns = _functions("preprocess_rl.py", {"precook", "cook_refs", "create_crefs", "compute_doc_freq"}, {"defaultdict": defaultdict})
    rng = np.random.RandomState(1)
    refs = [[" ".join(str(int(t)) for t in rng.randint(1, 12, size=rng.randint(3, 9))) + " 0" for _ in range(5)] for _ in range(40)]
    ref_df = ns["compute_doc_freq"](ns["create_crefs"](refs))
    mine = CO.compute_doc_freq([CO.cook_refs(r) for r in refs])
    assert dict(ref_df) == dict(mine)
    s = refs[3][2]
    assert dict(ns["precook"](s)) == dict(CO.precook(s))


def test_reward_glue_matches_editnet_rl():
    V, B, L = 40, 6, 18
    wm = synth.word_map(V)
    table = CO.synthetic_table(V, n_images=80, seed=2)
    scorer = CO.CiderD(table["document_frequency"], table["ref_len"])
    ns = _functions("editnet_rl.py", {"preprocess_gd", "array_to_str", "get_self_critical_reward"},
                    {"OrderedDict": OrderedDict, "np": np, "torch": torch, "CiderD_scorer": scorer})
    rng = np.random.RandomState(3)

    def seqs():
        out = np.zeros((B, L), dtype=np.int64)
        for i in range(B):
            n = rng.randint(1, L + 1)
            out[i, :n] = np.minimum(rng.zipf(1.3, size=n), V - 4)
        return out

    gen, gre = seqs(), seqs()
    allcaps = np.zeros((B, 5, 20), dtype=np.int64)
    for i in range(B):
        for r in range(5):
            n = rng.randint(3, 17)
            allcaps[i, r, 0] = wm["<start>"]
            allcaps[i, r, 1:1 + n] = np.minimum(rng.zipf(1.3, size=n), V - 4)
            allcaps[i, r, 1 + n] = wm["<end>"]
    gd_ref = ns["preprocess_gd"](torch.from_numpy(allcaps), wm)
    assert gd_ref == CO.preprocess_gd(torch.from_numpy(allcaps), wm)
    ref = ns["get_self_critical_reward"](torch.from_numpy(gen), torch.from_numpy(gre), gd_ref)
    mine = CO.self_critical_reward(scorer, gen, gre, gd_ref)
    assert ref.shape == (B, L) and np.allclose(ref.numpy(), mine, atol=1e-6)
    assert np.abs(mine).max() > 0          # non-degenerate case
