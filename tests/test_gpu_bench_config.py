"""Parity at the shapes the benchmark and BASELINE.json's configs actually run (VERDICT r1 "what's weak" 1-3):

 * configs[1]: EditNet XE train step, B=64, V=10000, T=19 FIXED lengths, dropout on -- the exact `bench.py` batch --
   through the module (logits, autograd gradients) AND through the fused trainer (`XETrainer.step`, the call bench.py
   times), against the oracle with the kernels' own Philox keep-bits; gradients judged against an fp64 run of the
   oracle.  The big time-batched GEMMs of this shape take the two-CTAs-per-SM ("twin") tcgen05 configuration: asserted.
 * configs[2]: greedy decode at B=256 (the non-swap GEMM path: M = 256 > the 128-wide Q tile), margin-filtered tokens.
 * configs[4]: adaptive (ragged 10..100 regions) at R=100, B=64, full dims, train mode.

Tolerances as in test_gpu_editnet.py: 1e-4 absolute on logits / log-probs, gradients within 2e-4 of each tensor's max
(or within 10x the reference's own fp32 error against fp64 truth)."""
import pytest
import torch

from oracle import editnet_oracle as EO
from oracle import synth

pytestmark = pytest.mark.gpu

TOL = 1e-4
GTOL = 2e-4
BENCH = dict(V=10000, D=1024, A=512, Fdim=2048, R=36, cap_width=20, prev_width=18, B=64)
XE_KEYS = ("feats", "caps", "caplens", "prev", "prev_len")


def _imports():
    from show_edit_tell_b200 import _lib, editnet, editnet_adaptive, editnet_rl, train
    import gpu_util
    return _lib, editnet, editnet_rl, editnet_adaptive, train, gpu_util


@pytest.fixture(scope="module")
def bench_sd():
    c = BENCH
    return EO.init_state_dict(c["V"], c["D"], c["D"], c["D"], c["A"], c["Fdim"], seed=0)


def _oracle(sd_cpu, batch, masks, adaptive, dtype):
    sd = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd_cpu.items()}
    im = batch.get("image_mean") if adaptive else None
    m = None if masks is None else {k: v.to(dtype) for k, v in masks.items()}
    preds, caps_sorted, dl, sort_ind = EO.xe_forward(
        sd, batch["feats"].to(dtype), batch["caps"], batch["caplens"], batch["prev"], batch["prev_len"], m,
        image_mean=None if im is None else im.to(dtype), stable_sort=True)   # the module sorts stably; ties abound at fixed lengths
    loss = EO.xe_loss(preds, caps_sorted, dl)
    import gpu_util
    grads = gpu_util.oracle_grads(sd, loss)
    return preds.detach(), float(loss.detach()), grads


def _judge_grads(mine, ref32, truth64, label):
    bad = []
    for k, r in truth64.items():
        scale = max(float(r.abs().max()), 1e-5)
        e_m = float((mine[k].cpu().double() - r).abs().max()) / scale
        e_o = float((ref32[k].double() - r).abs().max()) / scale
        e_abs = float((mine[k].cpu().double() - r).abs().max())
        # (last clause: see test_gpu_editnet.py -- exactly-cancelling visual-attention gradients of magnitude ~1e-5)
        if not (e_m < max(GTOL, 10 * e_o) or e_abs < 3e-6):
            bad.append((k, e_m, e_o, e_abs))
    assert not bad, (label, bad)


def test_bench_config_xe_train_module_and_trainer(bench_sd):
    """B=64, V=10000, T=19 fixed, train mode: module forward/backward and the fused trainer step vs the oracle"""
    _lib, editnet, editnet_rl, editnet_adaptive, trainmod, U = _imports()
    c = BENCH
    L = _lib.lib()
    batch = synth.make_batch(c["B"], c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=False, seed=100)
    mod, _ = U.build_module(editnet.DecoderC, bench_sd, c["V"], c["D"], c["A"], c["Fdim"])
    mod.train()
    args = [batch[k].cuda() for k in XE_KEYS]
    seed = 4242
    # ---- (1) the fused trainer step: exactly what bench.py times
    tr = trainmod.XETrainer(mod, distributed=False, lr=0.0)     # lr 0: parameters stay put for part (2)
    L.set_gemm_twin_launches(1)
    loss_tr = tr.step(*args, seed=seed)
    torch.cuda.synchronize()
    twin = int(L.set_gemm_twin_launches(1))
    assert twin >= 4, "the time-batched GEMMs of the bench shape did not take the twin tcgen05 configuration (%d)" % twin
    dl = tr.last_call.decode_lengths
    assert dl == [19] * c["B"]
    flat_tr = {k: v.detach().clone() for k, v in zip([k for _, k in _lib.EDITNET_FIELDS], mod._views(tr.flat_grad()))}
    masks = U.keep_masks(seed, c["B"], 19, c["prev_width"], c["D"], c["R"])
    ref_pred, ref_loss, ref_grads = _oracle(bench_sd, batch, masks, False, torch.float32)
    _, truth_loss, truth_grads = _oracle(bench_sd, batch, masks, False, torch.float64)
    print("trainer loss %.6f oracle fp32 %.6f fp64 %.6f" % (float(loss_tr), ref_loss, truth_loss))
    assert abs(float(loss_tr) - truth_loss) < TOL
    _judge_grads(flat_tr, ref_grads, truth_grads, "trainer")
    # ---- (2) the module path (batch-major predictions + autograd backward) with the same dropout bits
    mod._last_call = None
    call_seed = seed

    import show_edit_tell_b200.editnet as E
    orig = E._draw_seed
    E._draw_seed = lambda: call_seed
    try:
        L.set_gemm_twin_launches(1)
        pred, caps_sorted, dl2, sort_ind = mod(*args, False, 0.0)
    finally:
        E._draw_seed = orig
    assert dl2 == dl
    err = float((pred.detach().cpu() - ref_pred).abs().max())
    print("bench-config logits: max abs err vs oracle %.3e" % err)
    assert err < TOL
    loss = EO.xe_loss(pred, caps_sorted, dl2)
    assert abs(float(loss) - truth_loss) < TOL
    loss.backward()
    torch.cuda.synchronize()
    assert int(L.set_gemm_twin_launches(1)) >= 4
    _judge_grads(U.grads_by_key(mod), ref_grads, truth_grads, "module")


def test_bench_config_greedy_B256(bench_sd):
    """configs[2]: greedy decode, B=256, max_len 18, V=10000 (M = 256 rows: the non-swap tensor-core path)"""
    _lib, editnet, editnet_rl, *_rest, U = _imports()
    c = BENCH
    Bq = 256
    batch = synth.make_batch(Bq, c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=True, seed=7)
    mod, wm = U.build_module(editnet_rl.DecoderC, bench_sd, c["V"], c["D"], c["A"], c["Fdim"])
    mod.eval()
    V = c["V"]
    with torch.no_grad():
        seq, slp = mod(wm, batch["prev"].cuda(), batch["prev_len"].cuda(), batch["feats"].cuda(), True, False)
        rseq, rslp = EO.rollout(bench_sd, batch["prev"], batch["prev_len"], batch["feats"], V - 2, V - 1, "greedy")
    seq, slp = seq.cpu(), slp.cpu()
    same = (seq == rseq).all(1)
    print("greedy B=256: %d/%d sequences token-identical; logprob err %.3e" %
          (int(same.sum()), len(same), float((slp - rslp)[same].abs().max())))
    assert (slp - rslp)[same].abs().max() < TOL
    assert same.float().mean() >= 0.9
    with torch.no_grad():
        for i in (~same).nonzero().view(-1).tolist():
            t = int((seq[i] != rseq[i]).nonzero()[0])
            # identical up to t; the first differing step must be a near-tie in the oracle (top-2 margin < 1e-3)
            forced = rseq[i:i + 1].clone()
            one = lambda f: EO.rollout(bench_sd, batch["prev"][i:i + 1], batch["prev_len"][i:i + 1],
                                       batch["feats"][i:i + 1], V - 2, V - 1, "forced", forced=f)[1]
            lp = one(forced)
            alt = forced.clone()
            alt[0, t] = seq[i, t] if seq[i, t] != 0 else V - 1
            lp2 = one(alt)
            assert abs(float(lp[0, t] - lp2[0, t])) < 1e-3, (i, t)


def test_bench_config_adaptive_R100_B64(bench_sd):
    """configs[4]: ragged 10..100 regions padded to R=100, B=64, full dims, train mode (editnet_adaptive.py:438-457)"""
    _lib, editnet, editnet_rl, editnet_adaptive, trainmod, U = _imports()
    c = dict(BENCH, R=100)
    batch = synth.make_batch(c["B"], c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=True, seed=55,
                             adaptive=True, Rmin=10)
    mod, _ = U.build_module(editnet_adaptive.DecoderC, bench_sd, c["V"], c["D"], c["A"], c["Fdim"])
    mod.train()
    torch.manual_seed(77)
    args = [batch[k].cuda() for k in XE_KEYS]
    out = mod(args[0], batch["image_mean"].cuda(), *args[1:], False, 0.0)
    pred, caps_sorted, dl, sort_ind = out[:4]
    T = max(dl)
    masks = U.keep_masks(mod.last_seed, c["B"], T, c["prev_width"], c["D"], c["R"])
    ref_pred, ref_loss, ref_grads = _oracle(bench_sd, batch, masks, True, torch.float32)
    err = float((pred.detach().cpu() - ref_pred).abs().max())
    print("adaptive R=100 B=64: logits max abs err %.3e" % err)
    assert err < TOL
    loss = EO.xe_loss(pred, caps_sorted, dl)
    assert abs(float(loss) - ref_loss) < TOL
    loss.backward()
    _, _, truth_grads = _oracle(bench_sd, batch, masks, True, torch.float64)
    _judge_grads(U.grads_by_key(mod), ref_grads, truth_grads, "adaptive")
