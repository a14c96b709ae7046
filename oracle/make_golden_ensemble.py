"""Writes tests/golden/ensemble_beam.npz: the captions (token ids) and scores the REFERENCE's own ensemble search
(`evaluate_full`, eval/eval xe/eval_full.py, AST-extracted by oracle/ref_extract.py) returns for seeded small models.
Inputs and weights are regenerated from the seeds by the tests (oracle/synth.py, *_oracle.init_state_dict).
Run in the authoring container (needs /root/reference):  python -m oracle.make_golden_ensemble"""
import os

import numpy as np
import torch

from . import dcnet_oracle as DO
from . import editnet_oracle as EO
from . import ensemble_oracle as XO
from . import ref_extract as RX
from . import synth

# end_bias lifts the <end> logit of both networks so that beams terminate at different steps (random-init models never
# emit <end>: without it only the 50-step runaway guard of eval_full.py:198-207 is exercised)
CASES = [dict(seed=41, beam=3, end_bias=0.0), dict(seed=47, beam=3, end_bias=0.9), dict(seed=53, beam=3, end_bias=1.3),
         dict(seed=59, beam=5, end_bias=1.1), dict(seed=61, beam=4, end_bias=0.0)]
DIMS = dict(V=67, D=48, A=24, Fdim=96, R=9, cap_width=11, prev_width=8)


def case_inputs(seed, end_bias=0.0):
    d = DIMS
    sd_e = EO.init_state_dict(d["V"], d["D"], d["D"], d["D"], d["A"], d["Fdim"], seed=seed)
    sd_d = DO.init_state_dict(d["V"], d["D"], d["D"] // 2, d["D"], d["A"], seed=seed + 1)
    end = synth.word_map(d["V"])["<end>"]
    sd_e["fc.bias"][end] += end_bias
    sd_d["fc.bias"][end] += end_bias
    b = synth.make_batch(1, d["V"], d["R"], d["Fdim"], d["cap_width"], d["prev_width"], ragged=True, seed=seed + 2,
                         min_len=3, min_prev=2)
    return sd_e, sd_d, b


def main():
    search = RX.eval_full_search()
    ens, dns = RX.eval_class_modules()
    d = DIMS
    wm = synth.word_map(d["V"])
    inv = {k: v for k, v in wm.items()}
    out = {}
    for ci, c in enumerate(CASES):
        sd_e, sd_d, b = case_inputs(c["seed"], c["end_bias"])
        dec = ens["DecoderC"](wm, d["D"], d["D"], d["D"], d["A"], d["Fdim"])
        dec.load_state_dict(sd_e, strict=False)
        dae = dns["DAE"](wm, None, decoder_dim=d["D"], attention_dim=d["A"], caption_features_dim=d["D"] // 2, emb_dim=d["D"])
        dae.load_state_dict(sd_d, strict=False)

        class AR(torch.nn.Module):
            def __init__(self, dae):
                super().__init__()
                self.dae = dae

        with torch.no_grad():
            res = search([(b["feats"], torch.tensor([[ci]]), b["prev"], b["prev_len"])], AR(dae), dec, c["beam"], 0, wm)
            seq, score = XO.beam_search_ensemble(sd_e, sd_d, wm, b["feats"], b["prev"], b["prev_len"], beam_size=c["beam"])
        words = res[0]["caption"].split()
        ref_ids = [inv[w] for w in words]
        mine = [w for w in seq if w not in (wm["<start>"], wm["<end>"], wm["<pad>"])]
        assert ref_ids == mine, (ref_ids, mine)
        out["case%d_seed" % ci] = np.int64(c["seed"])
        out["case%d_beam" % ci] = np.int64(c["beam"])
        out["case%d_end_bias" % ci] = np.float32(c["end_bias"])
        out["case%d_caption" % ci] = np.asarray(ref_ids, dtype=np.int64)     # the reference's output
        out["case%d_seq" % ci] = np.asarray(seq, dtype=np.int64)             # full token list of the restatement
        out["case%d_score" % ci] = np.float32(score)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ensemble_beam.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in out.items() if "caption" in k})


if __name__ == "__main__":
    main()
