"""GPU parity of the DCNet path: golden vectors from the reference's real classes (small dims), the
CPU oracle at full dims (1024/512/512), config[0] of BASELINE.json (B=4 teacher-forced forward)."""
import pytest
import torch

from conftest import load_npz
from oracle import dcnet_oracle as DO
from oracle import editnet_oracle as EO
from oracle import synth

pytestmark = pytest.mark.gpu
TOL, GTOL = 1e-4, 2e-4


def _mod(cls, sd, cfg):
    import gpu_util as U  # noqa: F401
    m = cls(synth.word_map(cfg["V"]), None, decoder_dim=cfg["D"], attention_dim=cfg["A"],
            caption_features_dim=cfg["Cd"], emb_dim=cfg["E"])
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert all(k.startswith("caption_encoder.embed.") for k in missing) and not unexpected
    return m.cuda()


def _masks(seed, B, T, Wp, D):
    import gpu_util as U
    full = U.keep_masks(seed, B, T, Wp, D, 1)
    return {"enc": full["enc"], "emb": full["emb"], "fc": full["fc"]}


def _grads(mod):
    from show_edit_tell_b200._lib import DCNET_FIELDS
    return {k: (torch.zeros_like(mod.get_parameter(k)) if mod.get_parameter(k).grad is None
                else mod.get_parameter(k).grad.detach().clone()) for _, k in DCNET_FIELDS}


@pytest.fixture(scope="module")
def dc_small():
    return load_npz("dcnet_small_sd"), {k: int(v) for k, v in load_npz("dcnet_small_cfg").items()}


def test_dcnet_golden_xe_eval(dc_small):
    import gpu_util as U
    from show_edit_tell_b200 import dcnet
    sd, cfg = dc_small
    g = load_npz("dcnet_xe_eval")
    mod = _mod(dcnet.DAE, sd, cfg).eval()
    pred, caps_sorted, dl, sort_ind = mod(g["caps"].cuda(), g["caplens"].cuda(), g["prev"].cuda(), g["prev_len"].cuda())
    assert dl == g["decode_lengths"].tolist() and torch.equal(sort_ind.cpu(), g["sort_ind"])
    assert (pred.cpu() - g["predictions"]).abs().max() < TOL
    loss = EO.xe_loss(pred, caps_sorted, dl)
    assert abs(float(loss.detach()) - float(g["loss"])) < TOL
    loss.backward()
    ref = {k[5:]: v for k, v in g.items() if k.startswith("grad:")}
    assert not U.compare_grads(_grads(mod), ref, GTOL, "dcnet golden")


def _vs_oracle(cfg, sd, batch, train):
    import gpu_util as U
    from show_edit_tell_b200 import dcnet
    mod = _mod(dcnet.DAE, sd, cfg)
    mod.train(train)
    torch.manual_seed(5)
    pred, caps_sorted, dl, _ = mod(batch["caps"].cuda(), batch["caplens"].cuda(), batch["prev"].cuda(),
                                   batch["prev_len"].cuda())
    masks = _masks(mod.last_seed, cfg["B"], max(dl), batch["prev"].shape[1], cfg["D"]) if train else None
    s = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rp, rc, rdl, _ = DO.xe_forward(s, batch["caps"], batch["caplens"], batch["prev"], batch["prev_len"], masks)
    err = float((pred.cpu() - rp.detach()).abs().max())
    print("dcnet train=%s: logits err %.3e" % (train, err))
    assert rdl == dl and err < TOL
    rl = EO.xe_loss(rp, rc, rdl)
    EO.xe_loss(pred, caps_sorted, dl).backward()
    assert not U.compare_grads(_grads(mod), U.oracle_grads(s, rl), GTOL, "dcnet train=%s" % train)


@pytest.mark.parametrize("train", [False, True])
def test_dcnet_small_vs_oracle(train, dc_small):
    sd, cfg = dc_small
    b = synth.make_batch(cfg["B"], cfg["V"], 1, 4, cfg["cap_width"], cfg["prev_width"], ragged=True, seed=71,
                         min_len=3, min_prev=2)
    _vs_oracle(cfg, sd, b, train)


FULL = dict(V=1003, D=1024, Cd=512, E=1024, A=512, B=4, cap_width=20, prev_width=18)


@pytest.mark.parametrize("train", [False, True])
def test_dcnet_full_dims_config0(train):
    """BASELINE.json configs[0]: DCNet XE teacher-forced forward, batch 4, seq_len 20 (+ its backward)"""
    sd = DO.init_state_dict(FULL["V"], FULL["D"], FULL["Cd"], FULL["E"], FULL["A"], seed=9)
    b = synth.make_batch(FULL["B"], FULL["V"], 1, 4, FULL["cap_width"], FULL["prev_width"], ragged=False, seed=72)
    _vs_oracle(FULL, sd, b, train)


def test_dcnet_golden_rollout_greedy(dc_small):
    from show_edit_tell_b200 import dcnet_rl
    sd, cfg = dc_small
    g = load_npz("dcnet_rl_greedy")
    mod = _mod(dcnet_rl.DAE, sd, cfg).eval()
    with torch.no_grad():
        seq, slp = mod(synth.word_map(cfg["V"]), g["prev"].cuda(), g["prev_len"].cuda(), True, False)
    assert torch.equal(seq.cpu(), g["seq"])
    assert (slp.cpu() - g["seqLogprobs"]).abs().max() < TOL


def test_dcnet_rollout_forced_grads_vs_oracle(dc_small):
    import gpu_util as U
    from show_edit_tell_b200 import dcnet_rl
    sd, cfg = dc_small
    V = cfg["V"]
    b = synth.make_batch(cfg["B"], V, 1, 4, cfg["cap_width"], cfg["prev_width"], ragged=True, seed=73, min_len=3, min_prev=2)
    g0 = torch.Generator().manual_seed(3)
    forced = torch.randint(1, V - 4, (cfg["B"], 18), generator=g0)
    for i in range(cfg["B"]):
        forced[i, 4 + 2 * i:] = V - 1
    reward = torch.randn(cfg["B"], 1, generator=g0).repeat(1, 18)
    mod = _mod(dcnet_rl.DAE, sd, cfg).train()
    wm = synth.word_map(V)
    seq, slp = mod.rollout(wm, b["prev"].cuda(), b["prev_len"].cuda(), False, True, forced=forced.cuda(), seed=77)
    masks = _masks(77, cfg["B"], 18, cfg["prev_width"], cfg["D"])
    s = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    rseq, rslp = DO.rollout(s, b["prev"], b["prev_len"], V - 2, V - 1, "forced", masks=masks, forced=forced)
    assert torch.equal(seq.cpu(), rseq)
    assert (slp.detach().cpu() - rslp.detach()).abs().max() < TOL
    loss = dcnet_rl.RewardCriterion()(slp, seq, reward.cuda())
    rloss = EO.reward_criterion(rslp, rseq, reward)
    assert abs(float(loss.detach()) - float(rloss.detach())) < 1e-5
    loss.backward()
    assert not U.compare_grads(_grads(mod), U.oracle_grads(s, rloss), GTOL, "dcnet rollout")
