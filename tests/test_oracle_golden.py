"""The CPU restatement (oracle/editnet_oracle.py) is held to the outputs of the
reference's real classes stored in tests/golden/ (written by oracle/make_golden.py).
fp32 both sides; tolerance 1e-5 abs on logits/log-probs, 2e-5 on gradients."""
import pytest
import torch

from conftest import load_npz
from oracle import editnet_oracle as EO

TOL = 1e-5


def _masks(g):
    if "mask_enc" not in g:
        return None
    return {k: g["mask_" + k].float() for k in ("enc", "emb", "vis", "fc")}


def _grads(sd, loss):
    keys = list(sd.keys())
    gs = torch.autograd.grad(loss, [sd[k] for k in keys], allow_unused=True)
    return {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(keys, gs)}


@pytest.mark.parametrize("tag", ["editnet_xe_eval", "editnet_xe_train", "editnet_adaptive_eval"])
def test_xe_forward_and_grads(tag, small_sd):
    g = load_npz(tag)
    sd = {k: v.clone().requires_grad_(True) for k, v in small_sd.items()}
    preds, caps_sorted, dl, sort_ind = EO.xe_forward(
        sd, g["feats"], g["caps"], g["caplens"], g["prev"], g["prev_len"], _masks(g),
        image_mean=g.get("image_mean"))
    assert dl == g["decode_lengths"].tolist()
    assert torch.equal(sort_ind, g["sort_ind"])
    assert (preds - g["predictions"]).abs().max() < TOL
    loss = EO.xe_loss(preds, caps_sorted, dl)
    assert abs(float(loss) - float(g["loss"])) < TOL
    grads = _grads(sd, loss)
    for k in sd:
        ref = g["grad:" + k]
        assert (grads[k] - ref).abs().max() < 2e-5 * max(1.0, float(ref.abs().max())), k


def test_clip_adam_step(small_sd):
    g = load_npz("editnet_xe_eval")
    keys = list(small_sd.keys())
    params = [small_sd[k].clone() for k in keys]
    grads = [g["grad:" + k].clone() for k in keys]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    total = EO.clip_and_adam(params, grads, m, v, step=1)
    assert abs(float(total) - float(g["grad_norm"])) < 1e-5
    for k, p in zip(keys, params):
        assert (p - g["after_step:" + k]).abs().max() < 1e-6, k


def test_rollout_greedy(small_sd, small_cfg):
    g = load_npz("editnet_rl_greedy")
    V = small_cfg["V"]
    with torch.no_grad():
        seq, slp = EO.rollout(small_sd, g["prev"], g["prev_len"], g["feats"], V - 2, V - 1, "greedy")
    assert torch.equal(seq, g["seq"])
    assert (slp - g["seqLogprobs"]).abs().max() < TOL


def test_rollout_forced_and_reward_grads(small_sd, small_cfg):
    g = load_npz("editnet_rl_forced")
    V = small_cfg["V"]
    sd = {k: v.clone().requires_grad_(True) for k, v in small_sd.items()}
    seq, slp = EO.rollout(sd, g["prev"], g["prev_len"], g["feats"], V - 2, V - 1, "forced",
                          masks=_masks(g), forced=g["forced"])
    assert torch.equal(seq, g["seq"])
    assert (slp - g["seqLogprobs"]).abs().max() < TOL
    loss = EO.reward_criterion(slp, seq, g["reward"])
    assert abs(float(loss) - float(g["loss"])) < TOL
    grads = _grads(sd, loss)
    for k in sd:
        ref = g["grad:" + k]
        assert (grads[k] - ref).abs().max() < 2e-5 * max(1.0, float(ref.abs().max())), k
