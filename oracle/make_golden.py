"""TEST INFRASTRUCTURE -- writes tests/golden/*.npz from the REAL reference classes.

Run in the authoring container only (needs /root/reference):

    python -m oracle.make_golden

The reference's model classes are AST-extracted and exec'd unmodified
(`oracle/ref_extract.py`); they are driven with seeded synthetic inputs
(`oracle/synth.py`) and scripted dropout masks, and their outputs -- logits,
parameter gradients of the reference's own loss expression, greedy tokens,
log-probs -- are stored.  These files are the pin for `oracle/editnet_oracle.py`
and `oracle/dcnet_oracle.py` (tests/test_oracle_golden.py) and, on the GPU box
where /root/reference does not exist, for the CUDA path itself
(tests/test_gpu_golden.py).
"""
import os
import sys

import numpy as np
import torch
import torch.nn as nn
from torch.nn.utils.rnn import pack_padded_sequence

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import editnet_oracle as EO  # noqa: E402
from oracle import ref_extract as RX  # noqa: E402
from oracle import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# small-but-awkward dims: odd vocab, odd region count, ragged lengths
SMALL = dict(V=53, D=32, C=32, E=32, A=16, Fdim=64, R=7, cap_width=9, prev_width=7, B=6)


def _np(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def build_ref_editnet(ns, cfg, sd):
    wm = synth.word_map(cfg["V"])
    dec = ns["DecoderC"](wm, decoder_dim=cfg["D"], caption_features_dim=cfg["C"], emb_dim=cfg["E"],
                         attention_dim=cfg["A"], image_features_dim=cfg["Fdim"])
    missing, unexpected = dec.load_state_dict(sd, strict=False)
    # caption_encoder.embed aliases embed (editnet.py:463-464) -> its key is the only "missing" one
    assert all(k.startswith("caption_encoder.embed.") for k in missing), missing
    assert not unexpected, unexpected
    return dec, wm


def ref_xe_loss(scores, caps_sorted, decode_lengths):
    """the reference's own loss expression, editnet.py:571-577"""
    targets = caps_sorted[:, 1:]
    s = pack_padded_sequence(scores, decode_lengths, batch_first=True)
    t = pack_padded_sequence(targets, decode_lengths, batch_first=True)
    return nn.CrossEntropyLoss()(s.data, t.data)


def xe_dropout_script(masks, batch, decode_lengths, sort_ind):
    """my mask layout -> the reference's dropout call order (editnet.py:331, then per
    step :513/524 embed, :441 att_embed, :545 fc)"""
    prev_len_sorted = batch["prev_len"][sort_ind]
    _, enc_sort = prev_len_sorted.squeeze(1).sort(dim=0, descending=True)   # editnet.py:322
    calls = [masks["enc"][enc_sort]]
    for t in range(max(decode_lengths)):
        b = sum(l > t for l in decode_lengths)
        calls += [masks["emb"][t, :b], masks["vis"][t, :b], masks["fc"][t, :b]]
    return calls


def gen_editnet_xe(tag, ns, cfg, sd, train, adaptive=False, seed=0):
    batch = synth.make_batch(cfg["B"], cfg["V"], cfg["R"], cfg["Fdim"], cfg["cap_width"],
                             cfg["prev_width"], ragged=True, seed=seed, min_len=3, min_prev=2,
                             adaptive=adaptive, Rmin=2)
    dec, _ = build_ref_editnet(ns, cfg, sd)
    lens_sorted, sort_ind = batch["caplens"].squeeze(1).sort(dim=0, descending=True)
    decode_lengths = (lens_sorted - 1).tolist()
    T = max(decode_lengths)
    masks = None
    args = (batch["feats"],) + ((batch["image_mean"],) if adaptive else ()) + (
        batch["caps"], batch["caplens"], batch["prev"], batch["prev_len"], False, 0.0)
    if train:
        dec.train()
        masks = synth.make_masks(cfg["B"], T, cfg["prev_width"], cfg["E"], cfg["D"], cfg["R"], seed)
        with RX.DropoutScript(xe_dropout_script(masks, batch, decode_lengths, sort_ind)) as ds:
            out = dec(*args)
            assert ds.calls == len(ds.masks)
    else:
        dec.eval()
        out = dec(*args)
    scores, caps_sorted, dl, si = out[:4]
    assert dl == decode_lengths and torch.equal(si, sort_ind)
    loss = ref_xe_loss(scores, caps_sorted, dl)
    dec.zero_grad()
    loss.backward()
    rec = dict(batch)
    rec.update(predictions=scores, loss=loss, sort_ind=si, decode_lengths=np.asarray(dl))
    if masks is not None:
        rec.update({"mask_" + k: v.to(torch.uint8) for k, v in masks.items()})
    for k, p in dec.named_parameters():
        if k.startswith("caption_encoder.embed."):
            continue
        rec["grad:" + k] = p.grad.clone() if p.grad is not None else torch.zeros_like(p)
    if not adaptive and not train:
        # one optimizer step exactly as train() does it (editnet.py:580-581)
        opt = torch.optim.Adam(dec.parameters(), lr=5e-4)
        total = torch.nn.utils.clip_grad_norm_(filter(lambda p: p.requires_grad, dec.parameters()), 0.25)
        opt.step()
        rec["grad_norm"] = total
        for k, p in dec.named_parameters():
            if not k.startswith("caption_encoder.embed."):
                rec["after_step:" + k] = p.detach()
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **_np(rec))
    print(tag, "loss", float(loss), "T", T, "decode_lengths", dl)


def gen_editnet_rl(tag, ns, cfg, sd, mode, seed=0):
    batch = synth.make_batch(cfg["B"], cfg["V"], cfg["R"], cfg["Fdim"], cfg["cap_width"],
                             cfg["prev_width"], ragged=True, seed=seed, min_len=3, min_prev=2)
    dec, wm = build_ref_editnet(ns, cfg, sd)
    max_len = 18   # hard-coded, editnet_rl.py:487
    rec = dict(batch)
    if mode == "greedy":
        dec.eval()
        with torch.no_grad():
            seq, slp = dec(wm, batch["prev"], batch["prev_len"], batch["feats"], True, False)
    else:
        # sampled rollout with train-mode dropout; torch.multinomial is scripted so the
        # run is reproducible: the "samples" are a fixed random token table.
        dec.train()
        g = torch.Generator().manual_seed(77 + seed)
        forced = torch.randint(1, cfg["V"] - 4, (cfg["B"], max_len), generator=g)
        # finish rows at different times by injecting <end>
        for i in range(cfg["B"]):
            forced[i, 3 + 2 * i:] = cfg["V"] - 1
        masks = synth.make_masks(cfg["B"], max_len + 1, cfg["prev_width"], cfg["E"], cfg["D"], cfg["R"], seed)
        _, enc_sort = batch["prev_len"].squeeze(1).sort(dim=0, descending=True)
        calls = [masks["enc"][enc_sort]]
        for t in range(max_len + 1):
            calls += [masks["emb"][t], masks["vis"][t], masks["fc"][t]]
        step = {"t": 0}
        orig = torch.multinomial

        def scripted(p, n, *a, **k):
            t = step["t"]
            step["t"] += 1
            return forced[:, t:t + 1].clone()

        torch.multinomial = scripted
        try:
            with RX.DropoutScript(calls):
                seq, slp = dec(wm, batch["prev"], batch["prev_len"], batch["feats"], False, True)
        finally:
            torch.multinomial = orig
        reward = torch.randn(cfg["B"], 1, generator=g).repeat(1, max_len)
        loss = ns["RewardCriterion"]()(slp, seq, reward)
        dec.zero_grad()
        loss.backward()
        rec.update(forced=forced, reward=reward, loss=loss)
        rec.update({"mask_" + k: v.to(torch.uint8) for k, v in masks.items()})
        for k, p in dec.named_parameters():
            if not k.startswith("caption_encoder.embed."):
                rec["grad:" + k] = p.grad.clone() if p.grad is not None else torch.zeros_like(p)
    rec.update(seq=seq, seqLogprobs=slp.detach())
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **_np(rec))
    print(tag, "seq[0]", seq[0].tolist())


def main():
    assert RX.reference_available(), "needs /root/reference"
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(4)
    cfg = SMALL
    sd = EO.init_state_dict(cfg["V"], cfg["D"], cfg["C"], cfg["E"], cfg["A"], cfg["Fdim"], seed=3)
    # scale fc/embedding so greedy decode has healthy top-1 margins
    np.savez_compressed(os.path.join(OUT, "editnet_small_sd.npz"), **_np(sd))
    np.savez(os.path.join(OUT, "editnet_small_cfg.npz"), **{k: np.asarray(v) for k, v in cfg.items()})
    gen_editnet_xe("editnet_xe_eval", RX.editnet_xe_classes(), cfg, sd, train=False)
    gen_editnet_xe("editnet_xe_train", RX.editnet_xe_classes(), cfg, sd, train=True, seed=1)
    gen_editnet_xe("editnet_adaptive_eval", RX.editnet_adaptive_classes(), cfg, sd, train=False,
                   adaptive=True, seed=2)
    gen_editnet_rl("editnet_rl_greedy", RX.editnet_rl_classes(), cfg, sd, "greedy", seed=4)
    gen_editnet_rl("editnet_rl_forced", RX.editnet_rl_classes(), cfg, sd, "forced", seed=5)
    if os.path.exists(os.path.join(os.path.dirname(__file__), "dcnet_oracle.py")):
        from oracle import make_golden_dcnet
        make_golden_dcnet.main()


if __name__ == "__main__":
    main()
