"""Batched beam search on the device (show_edit_tell_b200/beam.py, csrc/beam.cu): N images x K beams in one step session,
expansion / top-k / <end> bookkeeping in one kernel per step, no host round trip inside the search.  Must return, for
every image, exactly what the reference's per-image search returns (oracle restatements of evaluate(),
editnet.py:595-719, and of the ensemble evaluate_full(), eval/eval xe/eval_full.py:88-215; both pinned to the
AST-extracted originals by the golden captions)."""
import time

import pytest
import torch

from oracle import dcnet_oracle as DO
from oracle import editnet_oracle as EO
from oracle import ensemble_oracle as XO
from oracle import make_golden_ensemble as ME
from oracle import make_golden_evaluate as MG
from oracle import synth

pytestmark = pytest.mark.gpu


def _strip(seq, wm):
    return [w for w in seq if w not in (wm["<start>"], wm["<end>"], wm["<pad>"])]


@pytest.mark.parametrize("beam,end_bias,seed", [(3, 1.0, 141), (5, 2.0, 161), (3, 0.0, 173)])
def test_batched_editnet_beam_equals_per_image_reference_search(beam, end_bias, seed):
    from show_edit_tell_b200 import editnet
    from show_edit_tell_b200.beam import beam_search_batched
    import gpu_util as U
    d = MG.DIMS
    N = 72
    sd, _ = MG.case_inputs(seed, end_bias)
    b = synth.make_batch(N, d["V"], d["R"], d["Fdim"], d["cap_width"], d["prev_width"], ragged=True, seed=seed + 7,
                         min_len=3, min_prev=2)
    mod, wm = U.build_module(editnet.DecoderC, sd, d["V"], d["D"], d["A"], d["Fdim"])
    mod.eval()
    with torch.no_grad():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        got = beam_search_batched(mod, wm, b["feats"].cuda(), b["prev"].cuda(), b["prev_len"].cuda(), beam_size=beam)
        dt = time.perf_counter() - t0
        lens = []
        for i in range(N):
            ref_seq, ref_score = EO.beam_search(sd, wm, b["feats"][i:i + 1], b["prev"][i:i + 1], b["prev_len"][i:i + 1],
                                                beam_size=beam)
            assert got[i][0] == ref_seq, (i, got[i][0], ref_seq)
            assert abs(got[i][1] - ref_score) < 1e-4 * max(1.0, abs(ref_score)) + 1e-3   # cumulative over up to 51 steps
            lens.append(len(ref_seq))
    print("batched beam-%d over %d images: %.1f images/s (first call, small model); caption lengths %d..%d" %
          (beam, N, N / dt, min(lens), max(lens)))
    assert len(set(lens)) > 1 or end_bias == 0.0


def test_batched_ensemble_beam_equals_per_image_reference_search():
    from show_edit_tell_b200 import dcnet, editnet
    from show_edit_tell_b200.beam import beam_search_batched
    import gpu_util as U
    d = ME.DIMS
    N = 64
    sd_e, sd_d, _ = ME.case_inputs(47, 0.9)
    b = synth.make_batch(N, d["V"], d["R"], d["Fdim"], d["cap_width"], d["prev_width"], ragged=True, seed=99,
                         min_len=3, min_prev=2)
    dec, wm = U.build_module(editnet.DecoderC, sd_e, d["V"], d["D"], d["A"], d["Fdim"])
    dae = dcnet.DAE(wm, None, decoder_dim=d["D"], attention_dim=d["A"], caption_features_dim=d["D"] // 2, emb_dim=d["D"])
    dae.load_state_dict(sd_d, strict=False)
    dec.eval(); dae = dae.cuda().eval()
    with torch.no_grad():
        got = beam_search_batched(dec, wm, b["feats"].cuda(), b["prev"].cuda(), b["prev_len"].cuda(), beam_size=3, dae=dae)
        for i in range(N):
            ref_seq, ref_score = XO.beam_search_ensemble(sd_e, sd_d, wm, b["feats"][i:i + 1], b["prev"][i:i + 1],
                                                         b["prev_len"][i:i + 1], beam_size=3)
            assert _strip(got[i][0], wm) == _strip(ref_seq, wm), (i, got[i][0], ref_seq)
