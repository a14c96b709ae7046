"""The sub-module call surface: the reference's beam search does not go through DecoderC.forward, it
calls decoder.caption_encoder / .embed / .attention_lstm / .caption_attention / .visual_attention / .select / .copy_lstm
/ .fc one by one (evaluate(), editnet.py:613,645-653).  (1) each sub-module forward against the oracle's cell; (2) the
reference's search loop (tests/ref_loops.py, verified against the AST-extracted original in
test_ref_loops_vs_reference.py) driven on the CUDA modules returns the captions the reference's own `evaluate` produced
(tests/golden/editnet_beam.npz, oracle/make_golden_evaluate.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import editnet_oracle as EO
from oracle import make_golden_evaluate as MG
from oracle import synth

import ref_loops

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _build(sd, d, cls=None):
    from show_edit_tell_b200 import editnet
    import gpu_util as U
    mod, wm = U.build_module(cls or editnet.DecoderC, sd, d["V"], d["D"], d["A"], d["Fdim"])
    return mod.eval(), wm


def test_each_submodule_forward_matches_the_oracle_cell():
    d = MG.DIMS
    sd, _ = MG.case_inputs(141, 0.0)
    mod, wm = _build(sd, d)
    g = torch.Generator().manual_seed(5)
    k, P, R, D, F = 5, 7, d["R"], d["D"], d["Fdim"]
    b = synth.make_batch(k, d["V"], R, F, d["cap_width"], d["prev_width"], ragged=True, seed=9, min_len=3, min_prev=2)
    with torch.no_grad():
        # caption encoder (editnet.py:319-348)
        h, m, fh, mask = mod.caption_encoder(b["prev"].cuda(), b["prev_len"].cuda())
        rh, rm, rfh, rmask = EO.caption_encoder(sd, b["prev"], b["prev_len"])
        for a, r in ((h, rh), (m, rm), (fh, rfh), (mask, rmask)):
            assert (a.cpu() - r).abs().max() < 1e-5
        # embedding (editnet.py:300-304), the (k, 1) shape evaluate() uses
        toks = torch.randint(1, d["V"] - 4, (k, 1), generator=g)
        e = mod.embed(toks.cuda())
        assert e.shape == (k, 1, D)
        re_ = EO.embed(sd, toks.view(-1))
        assert (e.squeeze(1).cpu() - re_).abs().max() < 1e-6
        h1 = torch.randn(k, D, generator=g) * 0.5
        c1 = torch.randn(k, D, generator=g) * 0.5
        x = torch.randn(k, 3 * D + F, generator=g)
        # attention_lstm: nn.LSTMCell (editnet.py:532)
        nh, nc = mod.attention_lstm(x.cuda(), (h1.cuda(), c1.cuda()))
        rh1, rc1 = EO.torch_lstm_cell(sd, "attention_lstm", x, h1, c1)
        assert (nh.cpu() - rh1).abs().max() < TOL and (nc.cpu() - rc1).abs().max() < TOL
        # caption attention (editnet.py:364-381)
        ctx, alpha = mod.caption_attention(h, h1.cuda(), e.squeeze(1), mask)
        rctx, ralpha = EO.caption_attention(sd, rh, h1, re_, rmask)
        assert (ctx.cpu() - rctx).abs().max() < TOL and (alpha.cpu() - ralpha).abs().max() < TOL
        # visual attention (editnet.py:439-447)
        att = mod.visual_attention(b["feats"].cuda(), h1.cuda())
        ratt = EO.visual_attention(sd, b["feats"], h1)
        assert (att.cpu() - ratt).abs().max() < TOL
        # select (editnet.py:403-421)
        sel = mod.select(m, alpha)
        rsel = EO.select(rm, ralpha)
        assert (sel.cpu() - rsel).abs().max() < 1e-5
        # copy-LSTM (editnet.py:265-285)
        x2 = torch.randn(k, 2 * D + F, generator=g)
        h2, c2 = mod.copy_lstm(x2.cuda(), (h1.cuda(), c1.cuda()), sel)
        rh2, rc2 = EO.copy_lstm(sd, x2, h1, c1, rsel)
        assert (h2.cpu() - rh2).abs().max() < TOL and (c2.cpu() - rc2).abs().max() < TOL
        # fc (editnet.py:653)
        sc = mod.fc(h2)
        assert (sc.cpu() - torch.nn.functional.linear(rh2, sd["fc.weight"], sd["fc.bias"])).abs().max() < 1e-3


def test_reference_search_loop_on_cuda_modules_returns_the_reference_captions():
    g = np.load(os.path.join(GOLDEN, "editnet_beam.npz"))
    d = MG.DIMS
    for ci, c in enumerate(MG.CASES):
        sd, b = MG.case_inputs(c["seed"], c["end_bias"])
        mod, wm = _build(sd, d)
        with torch.no_grad():
            got = ref_loops.evaluate_one(mod, wm, b["feats"].cuda(), b["prev"].cuda(), b["prev_len"].cuda(), c["beam"], d["V"])
        assert got == g["case%d_caption" % ci].tolist(), (ci, got, g["case%d_caption" % ci].tolist())
        # and the library's own one-call-per-step search (editnet.beam_search) agrees
        from show_edit_tell_b200.editnet import beam_search
        seq, _ = beam_search(mod, wm, b["feats"].cuda(), b["prev"].cuda(), b["prev_len"].cuda(), beam_size=c["beam"])
        assert [w for w in seq if w not in (wm["<start>"], wm["<end>"], wm["<pad>"])] == got, ci


def test_reference_ensemble_loop_on_cuda_modules_returns_the_reference_captions():
    """evaluate_full (eval/eval xe/eval_full.py:97-207) through the EditNet AND DCNet sub-module surfaces"""
    from oracle import make_golden_ensemble as ME
    from show_edit_tell_b200 import dcnet, dcnet_rl, editnet
    g = np.load(os.path.join(GOLDEN, "ensemble_beam.npz"))
    d = ME.DIMS
    for ci, c in enumerate(ME.CASES):
        sd_e, sd_d, b = ME.case_inputs(c["seed"], c["end_bias"])
        dec, wm = _build(sd_e, d)
        dae = dcnet.DAE(wm, None, decoder_dim=d["D"], attention_dim=d["A"], caption_features_dim=d["D"] // 2, emb_dim=d["D"])
        missing, unexpected = dae.load_state_dict(sd_d, strict=False)
        assert not unexpected and all(k.startswith("caption_encoder.embed.") for k in missing)
        ar = dcnet_rl.DAEWithAR(dae=dae.cuda().eval())
        with torch.no_grad():
            got = ref_loops.evaluate_full_one(ar, dec, wm, b["feats"].cuda(), b["prev"].cuda(), b["prev_len"].cuda(), c["beam"])
        assert got == g["case%d_caption" % ci].tolist(), (ci, got, g["case%d_caption" % ci].tolist())
