"""TEST INFRASTRUCTURE -- DCNet golden vectors from the reference's real classes (see make_golden.py)."""
import os

import numpy as np
import torch
import torch.nn as nn

from oracle import dcnet_oracle as DO
from oracle import ref_extract as RX
from oracle import synth
from oracle.make_golden import OUT, _np, ref_xe_loss

SMALL = dict(V=47, D=32, Cd=16, E=32, A=16, cap_width=9, prev_width=7, B=6)


def build_ref(ns, cfg, sd):
    dae = ns["DAE"](synth.word_map(cfg["V"]), None, decoder_dim=cfg["D"], attention_dim=cfg["A"],
                    caption_features_dim=cfg["Cd"], emb_dim=cfg["E"])
    missing, unexpected = dae.load_state_dict(sd, strict=False)
    assert all(k.startswith("caption_encoder.embed.") for k in missing), missing
    assert not unexpected, unexpected
    return dae


def grads_of(dae):
    return {"grad:" + k: (p.grad.clone() if p.grad is not None else torch.zeros_like(p))
            for k, p in dae.named_parameters() if not k.startswith("caption_encoder.embed.")}


def gen_xe(tag, cfg, sd, train, seed):
    ns = RX.dcnet_xe_classes()
    b = synth.make_batch(cfg["B"], cfg["V"], 1, 4, cfg["cap_width"], cfg["prev_width"], ragged=True, seed=seed,
                         min_len=3, min_prev=2)
    dae = build_ref(ns, cfg, sd)
    lens_sorted, sort_ind = b["caplens"].squeeze(1).sort(dim=0, descending=True)
    dl = (lens_sorted - 1).tolist()
    T = max(dl)
    rec = {k: b[k] for k in ("caps", "caplens", "prev", "prev_len")}
    if train:
        dae.train()
        g = torch.Generator().manual_seed(500 + seed)
        bern = lambda *s: (torch.rand(*s, generator=g) < 0.5).float()
        masks = {"enc": bern(cfg["B"], cfg["prev_width"], cfg["E"]), "emb": bern(T, cfg["B"], cfg["E"]),
                 "fc": bern(T, cfg["B"], cfg["D"])}
        # call 0: embed(encoded_captions) over (B, Wc, E) (dcnet.py:327); call 1: embed(src) (:225); then fc (:347)
        m0 = bern(cfg["B"], cfg["cap_width"], cfg["E"])
        m0[:, :T] = masks["emb"].permute(1, 0, 2)
        calls = [m0, masks["enc"]]
        for t in range(T):
            calls.append(masks["fc"][t, :sum(l > t for l in dl)])
        with RX.DropoutScript(calls) as ds:
            out = dae(b["caps"], b["caplens"], b["prev"], b["prev_len"])
            assert ds.calls == len(calls)
        rec.update({"mask_" + k: v.to(torch.uint8) for k, v in masks.items()})
    else:
        dae.eval()
        out = dae(b["caps"], b["caplens"], b["prev"], b["prev_len"])
    scores, caps_sorted, dl2, si = out
    assert dl2 == dl and torch.equal(si, sort_ind)
    loss = ref_xe_loss(scores, caps_sorted, dl)
    dae.zero_grad()
    loss.backward()
    rec.update(predictions=scores, loss=loss.detach(), sort_ind=si, decode_lengths=np.asarray(dl))
    rec.update(grads_of(dae))
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **_np(rec))
    print(tag, "loss", float(loss.detach()), "decode_lengths", dl)


def gen_rl(tag, cfg, sd, seed):
    ns = RX.dcnet_rl_classes()
    b = synth.make_batch(cfg["B"], cfg["V"], 1, 4, cfg["cap_width"], cfg["prev_width"], ragged=True, seed=seed,
                         min_len=3, min_prev=2)
    dae = build_ref(ns, cfg, sd)
    wm = synth.word_map(cfg["V"])
    dae.eval()
    with torch.no_grad():
        seq, slp = dae(wm, b["prev"], b["prev_len"], True, False)
    rec = {"prev": b["prev"], "prev_len": b["prev_len"], "seq": seq, "seqLogprobs": slp}
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **_np(rec))
    print(tag, "seq[0]", seq[0].tolist())


def main():
    cfg = SMALL
    sd = DO.init_state_dict(cfg["V"], cfg["D"], cfg["Cd"], cfg["E"], cfg["A"], seed=7)
    np.savez_compressed(os.path.join(OUT, "dcnet_small_sd.npz"), **_np(sd))
    np.savez(os.path.join(OUT, "dcnet_small_cfg.npz"), **{k: np.asarray(v) for k, v in cfg.items()})
    gen_xe("dcnet_xe_eval", cfg, sd, False, seed=61)
    gen_xe("dcnet_xe_train", cfg, sd, True, seed=62)
    gen_rl("dcnet_rl_greedy", cfg, sd, seed=63)


if __name__ == "__main__":
    main()
