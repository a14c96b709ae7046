"""Writes tests/golden/editnet_beam.npz: the captions (token ids) the REFERENCE's own beam search (`evaluate`,
editnet.py:595-719, AST-extracted by oracle/ref_extract.py with the `/` -> `//` repair of SURVEY Appendix D) returns for
seeded small EditNet models driven through the reference's own classes.  Inputs and weights are regenerated from the
seeds by the tests.  Also checks that tests/ref_loops.py (the restatement the GPU test drives) reproduces them on the
same reference modules.  Run in the authoring container (needs /root/reference):  python -m oracle.make_golden_evaluate"""
import os
import sys

import numpy as np
import torch

from . import editnet_oracle as EO
from . import ref_extract as RX
from . import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

# A random-init model has almost static dynamics and never emits <end>, which would exercise only the 50-step runaway
# guard of editnet.py:702-713.  The cases therefore scale the recurrent weights / embedding / fc (livelier state, wider
# logits) and lift the <end> logit by `end_bias`, so that beams complete at different steps.
CASES = [dict(seed=141, beam=3, end_bias=0.0), dict(seed=141, beam=3, end_bias=2.0), dict(seed=141, beam=5, end_bias=1.0),
         dict(seed=161, beam=4, end_bias=2.0), dict(seed=153, beam=3, end_bias=3.0), dict(seed=173, beam=3, end_bias=2.0),
         dict(seed=159, beam=3, end_bias=1.0), dict(seed=173, beam=5, end_bias=3.0)]
DIMS = dict(V=67, D=256, A=128, Fdim=256, R=9, cap_width=11, prev_width=8)


def case_inputs(seed, end_bias=0.0):
    d = DIMS
    sd = EO.init_state_dict(d["V"], d["D"], d["D"], d["D"], d["A"], d["Fdim"], seed=seed)
    for k in sd:
        if k.startswith(("attention_lstm.weight", "copy_lstm.x2h.weight", "copy_lstm.h2h.weight", "copy_lstm.gate")):
            sd[k] *= 6
    sd["fc.weight"] *= 10
    sd["embed.embedding.weight"] *= 4
    sd["fc.bias"][synth.word_map(d["V"])["<end>"]] += end_bias
    b = synth.make_batch(1, d["V"], d["R"], d["Fdim"], d["cap_width"], d["prev_width"], ragged=True, seed=seed + 2,
                         min_len=3, min_prev=2)
    return sd, b


def main():
    import ref_loops
    search = RX.editnet_evaluate_search()
    ns = RX.editnet_xe_classes()
    d = DIMS
    wm = synth.word_map(d["V"])
    out = {}
    for ci, c in enumerate(CASES):
        sd, b = case_inputs(c["seed"], c["end_bias"])
        dec = ns["DecoderC"](wm, d["D"], d["D"], d["D"], d["A"], d["Fdim"])
        dec.load_state_dict(sd, strict=False)
        dec.eval()
        with torch.no_grad():
            res = search([(b["feats"], torch.tensor([[ci]]), b["prev"], b["prev_len"])], dec, c["beam"], 0, d["V"], wm)
            mine = ref_loops.evaluate_one(dec, wm, b["feats"], b["prev"], b["prev_len"], c["beam"], d["V"])
        ref_ids = [wm[w] for w in res[0]["caption"].split()]
        assert ref_ids == mine, (ci, ref_ids, mine)
        out["case%d_seed" % ci] = np.int64(c["seed"])
        out["case%d_beam" % ci] = np.int64(c["beam"])
        out["case%d_end_bias" % ci] = np.float32(c["end_bias"])
        out["case%d_caption" % ci] = np.asarray(ref_ids, dtype=np.int64)
    path = os.path.join(ROOT, "tests", "golden", "editnet_beam.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.tolist() for k, v in out.items() if "caption" in k})


if __name__ == "__main__":
    main()
