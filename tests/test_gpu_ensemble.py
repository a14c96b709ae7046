"""CUDA ensemble beam search (show_edit_tell_b200.eval_full) against the CPU oracle and the golden captions."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import GOLDEN
from oracle import dcnet_oracle as DO
from oracle import editnet_oracle as EO
from oracle import ensemble_oracle as XO
from oracle import make_golden_ensemble as MG
from oracle import synth

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _modules(sd_e, sd_d, V, D, A, Fd):
    from show_edit_tell_b200 import dcnet, editnet
    wm = synth.word_map(V)
    dec = editnet.DecoderC(wm, D, D, D, A, Fd)
    dec.load_state_dict(sd_e, strict=False)
    dae = dcnet.DAE(wm, None, D, A, D // 2, D)
    dae.load_state_dict(sd_d, strict=False)
    return dec.cuda().eval(), dae.cuda().eval(), wm


def _check(dec, dae, wm, sd_e, sd_d, b, beam):
    from show_edit_tell_b200.eval_full import beam_search_ensemble
    with torch.no_grad():
        seq, score, cs, css = beam_search_ensemble(dec, dae, wm, b["feats"].cuda(), b["prev"].cuda(), b["prev_len"].cuda(),
                                                   beam_size=beam, return_all=True)
        rseq, rscore, rcs, rcss = XO.beam_search_ensemble(sd_e, sd_d, wm, b["feats"], b["prev"], b["prev_len"],
                                                          beam_size=beam, return_all=True)
    assert seq == rseq and abs(score - rscore) < TOL * max(1, len(seq))
    assert cs == rcs and len(css) == len(rcss)
    assert all(abs(a - r) < TOL * max(1, len(s)) for a, r, s in zip(css, rcss, rcs))
    return seq


def test_ensemble_beam_small_dims_vs_oracle_and_golden():
    z = np.load(os.path.join(GOLDEN, "ensemble_beam.npz"))
    d = MG.DIMS
    n = len([k for k in z.files if k.endswith("_seed")])
    for ci in range(n):
        sd_e, sd_d, b = MG.case_inputs(int(z["case%d_seed" % ci]), float(z["case%d_end_bias" % ci]))
        dec, dae, wm = _modules(sd_e, sd_d, d["V"], d["D"], d["A"], d["Fdim"])
        seq = _check(dec, dae, wm, sd_e, sd_d, b, int(z["case%d_beam" % ci]))
        assert seq == z["case%d_seq" % ci].tolist()          # = what the reference's evaluate_full returned


def test_ensemble_beam_full_dims_vs_oracle():
    V, D, A, Fd, R = 1003, 1024, 512, 2048, 36
    sd_e = EO.init_state_dict(V, D, D, D, A, Fd, seed=5)
    sd_d = DO.init_state_dict(V, D, D // 2, D, A, seed=9)
    wm = synth.word_map(V)
    for end_bias, beam, seed in ((1.5, 3, 81), (0.0, 3, 83)):
        se = {k: v.clone() for k, v in sd_e.items()}
        sdd = {k: v.clone() for k, v in sd_d.items()}
        se["fc.bias"][wm["<end>"]] += end_bias
        sdd["fc.bias"][wm["<end>"]] += end_bias
        b = synth.make_batch(1, V, R, Fd, 20, 18, ragged=True, seed=seed)
        dec, dae, _ = _modules(se, sdd, V, D, A, Fd)
        _check(dec, dae, wm, se, sdd, b, beam)
