"""Where /root/reference exists (the authoring container): the restatement of the reference's beam-search loop that the
GPU test drives (tests/ref_loops.py) and the AST-extracted original `evaluate` (editnet.py:595-719) return the same
captions when both drive the reference's own modules, and both equal the committed golden captions."""
import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import make_golden_evaluate as MG
from oracle import ref_extract as RX
from oracle import synth

import os
import ref_loops

pytestmark = pytest.mark.skipif(not RX.reference_available(), reason="needs the reference tree")


def test_restated_loop_equals_reference_evaluate_and_golden():
    search = RX.editnet_evaluate_search()
    ns = RX.editnet_xe_classes()
    d = MG.DIMS
    wm = synth.word_map(d["V"])
    g = np.load(os.path.join(GOLDEN, "editnet_beam.npz"))
    for ci, c in enumerate(MG.CASES[:4]):
        sd, b = MG.case_inputs(c["seed"], c["end_bias"])
        dec = ns["DecoderC"](wm, d["D"], d["D"], d["D"], d["A"], d["Fdim"])
        dec.load_state_dict(sd, strict=False)
        dec.eval()
        with torch.no_grad():
            res = search([(b["feats"], torch.tensor([[ci]]), b["prev"], b["prev_len"])], dec, c["beam"], 0, d["V"], wm)
            mine = ref_loops.evaluate_one(dec, wm, b["feats"], b["prev"], b["prev_len"], c["beam"], d["V"])
        ref_ids = [wm[w] for w in res[0]["caption"].split()]
        assert ref_ids == mine == g["case%d_caption" % ci].tolist()


def test_restated_ensemble_loop_equals_reference_evaluate_full_and_golden():
    from oracle import make_golden_ensemble as ME
    search = RX.eval_full_search()
    ens, dns = RX.eval_class_modules()
    # the class-only eval copies have `pass` forwards for the wrappers; the cells are complete
    d = ME.DIMS
    wm = synth.word_map(d["V"])
    g = np.load(os.path.join(GOLDEN, "ensemble_beam.npz"))
    for ci, c in enumerate(ME.CASES[:3]):
        sd_e, sd_d, b = ME.case_inputs(c["seed"], c["end_bias"])
        dec = ens["DecoderC"](wm, d["D"], d["D"], d["D"], d["A"], d["Fdim"])
        dec.load_state_dict(sd_e, strict=False)
        dae = dns["DAE"](wm, None, decoder_dim=d["D"], attention_dim=d["A"], caption_features_dim=d["D"] // 2, emb_dim=d["D"])
        dae.load_state_dict(sd_d, strict=False)

        class AR(torch.nn.Module):
            def __init__(self, dae):
                super().__init__()
                self.dae = dae

        ar = AR(dae).eval()
        dec.eval()
        with torch.no_grad():
            res = search([(b["feats"], torch.tensor([[ci]]), b["prev"], b["prev_len"])], ar, dec, c["beam"], 0, wm)
            mine = ref_loops.evaluate_full_one(ar, dec, wm, b["feats"], b["prev"], b["prev_len"], c["beam"])
        ref_ids = [wm[w] for w in res[0]["caption"].split()]
        assert ref_ids == mine == g["case%d_caption" % ci].tolist()
