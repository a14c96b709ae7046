"""DCNet restatement (oracle/dcnet_oracle.py) vs the reference's real classes (tests/golden/dcnet_*.npz)."""
import numpy as np
import pytest
import torch

from conftest import load_npz
from oracle import dcnet_oracle as DO
from oracle import editnet_oracle as EO

TOL = 1e-5


@pytest.fixture(scope="module")
def dc_sd():
    return load_npz("dcnet_small_sd")


@pytest.fixture(scope="module")
def dc_cfg():
    z = np.load("tests/golden/dcnet_small_cfg.npz") if False else load_npz("dcnet_small_cfg")
    return {k: int(v) for k, v in z.items()}


@pytest.mark.parametrize("tag", ["dcnet_xe_eval", "dcnet_xe_train"])
def test_dcnet_xe_forward_and_grads(tag, dc_sd):
    g = load_npz(tag)
    sd = {k: v.clone().requires_grad_(True) for k, v in dc_sd.items()}
    masks = {k: g["mask_" + k].float() for k in ("enc", "emb", "fc")} if "mask_enc" in g else None
    preds, caps_sorted, dl, sort_ind = DO.xe_forward(sd, g["caps"], g["caplens"], g["prev"], g["prev_len"], masks)
    assert dl == g["decode_lengths"].tolist() and torch.equal(sort_ind, g["sort_ind"])
    assert (preds - g["predictions"]).abs().max() < TOL
    loss = EO.xe_loss(preds, caps_sorted, dl)
    assert abs(float(loss.detach()) - float(g["loss"])) < TOL
    keys = list(sd)
    grads = torch.autograd.grad(loss, [sd[k] for k in keys], allow_unused=True)
    for k, gr in zip(keys, grads):
        ref = g["grad:" + k]
        gr = torch.zeros_like(ref) if gr is None else gr
        assert (gr - ref).abs().max() < 2e-5 * max(1.0, float(ref.abs().max())), k


def test_dcnet_rollout_greedy(dc_sd, dc_cfg):
    g = load_npz("dcnet_rl_greedy")
    V = dc_cfg["V"]
    with torch.no_grad():
        seq, slp = DO.rollout(dc_sd, g["prev"], g["prev_len"], V - 2, V - 1, "greedy")
    assert torch.equal(seq, g["seq"])
    assert (slp - g["seqLogprobs"]).abs().max() < TOL
