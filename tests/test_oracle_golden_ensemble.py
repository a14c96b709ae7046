"""The CPU restatement of the EditNet + DCNet ensemble beam search against tests/golden/ensemble_beam.npz (captions the
reference's own evaluate_full loop produced in the authoring container, oracle/make_golden_ensemble.py)."""
import numpy as np
import torch

from conftest import GOLDEN
from oracle import ensemble_oracle as XO
from oracle import make_golden_ensemble as MG
from oracle import synth
import os


def test_ensemble_oracle_matches_reference_captions():
    z = np.load(os.path.join(GOLDEN, "ensemble_beam.npz"))
    wm = synth.word_map(MG.DIMS["V"])
    n = len([k for k in z.files if k.endswith("_seed")])
    assert n >= 4
    for ci in range(n):
        sd_e, sd_d, b = MG.case_inputs(int(z["case%d_seed" % ci]), float(z["case%d_end_bias" % ci]))
        with torch.no_grad():
            seq, score = XO.beam_search_ensemble(sd_e, sd_d, wm, b["feats"], b["prev"], b["prev_len"],
                                                 beam_size=int(z["case%d_beam" % ci]))
        caption = [w for w in seq if w not in (wm["<start>"], wm["<end>"], wm["<pad>"])]
        assert caption == z["case%d_caption" % ci].tolist()
        assert seq == z["case%d_seq" % ci].tolist()
        assert abs(score - float(z["case%d_score" % ci])) < 1e-5
