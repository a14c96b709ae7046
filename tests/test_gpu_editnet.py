"""GPU parity tests: the CUDA path (through the C ABI, via the reference-shaped modules) against
 (a) the golden vectors written from the reference's real classes, and
 (b) the CPU oracle on the same seeded inputs, at small and at full (1024/512/2048) dims.
Tolerance: 1e-4 absolute on fp32 logits / log-probs (BASELINE.json north_star); gradients within
2e-4 of each tensor's max magnitude; greedy tokens identical wherever the oracle's top-2 margin
exceeds 1e-3."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from conftest import load_npz
from oracle import editnet_oracle as EO
from oracle import synth

pytestmark = pytest.mark.gpu

TOL = 1e-4
GTOL = 2e-4


def _imports():
    from show_edit_tell_b200 import _lib, editnet, editnet_adaptive, editnet_rl, train
    import gpu_util
    return _lib, editnet, editnet_rl, editnet_adaptive, train, gpu_util


def _cuda(batch, keys):
    return [batch[k].cuda() for k in keys]


XE_KEYS = ("feats", "caps", "caplens", "prev", "prev_len")


def test_library_loads_and_gemm_modes():
    _lib, *_ = _imports()
    L = _lib.lib()
    torch.manual_seed(0)
    for (M, N, K) in [(64, 96, 128), (7, 53, 32), (130, 70, 45), (1, 1, 4), (65, 33, 1000)]:
        A = torch.randn(M, K, device="cuda")
        W = torch.randn(N, K, device="cuda")
        bias = torch.randn(N, device="cuda")
        Cm = torch.empty(M, N, device="cuda")
        _lib.check(L.set_gemm(0, M, N, K, _lib.ptr(A), K, _lib.ptr(W), K, _lib.ptr(bias), _lib.ptr(Cm), N, 0, 0, None))
        ref = A.double() @ W.double().t() + bias.double()
        assert (Cm.double() - ref).abs().max() < 1e-3 * max(1.0, K ** 0.5 / 8), ("NT", M, N, K)
        # NN: C[M,K] = dY[M,N] @ W[N,K]
        dY = torch.randn(M, N, device="cuda")
        dX = torch.empty(M, K, device="cuda")
        _lib.check(L.set_gemm(1, M, K, N, _lib.ptr(dY), N, _lib.ptr(W), K, None, _lib.ptr(dX), K, 0, 0, None))
        assert (dX.double() - dY.double() @ W.double()).abs().max() < 1e-3 * max(1.0, N ** 0.5 / 8), ("NN", M, N, K)
        # TN: dW[N,K] += dY[M,N]^T @ A[M,K]
        dW = torch.ones(N, K, device="cuda")
        _lib.check(L.set_gemm(2, N, K, M, _lib.ptr(dY), N, _lib.ptr(A), K, None, _lib.ptr(dW), K, 1, 0, None))
        assert (dW.double() - 1 - dY.double().t() @ A.double()).abs().max() < 1e-3 * max(1.0, M ** 0.5 / 8), ("TN", M, N, K)


@pytest.mark.parametrize("tag", ["editnet_xe_eval", "editnet_adaptive_eval"])
def test_golden_xe_eval_forward_backward(tag, small_sd, small_cfg):
    _lib, editnet, editnet_rl, editnet_adaptive, train, U = _imports()
    g = load_npz(tag)
    c = small_cfg
    adaptive = "image_mean" in g
    cls = editnet_adaptive.DecoderC if adaptive else editnet.DecoderC
    mod, _ = U.build_module(cls, small_sd, c["V"], c["D"], c["A"], c["Fdim"])
    mod.eval()
    args = _cuda(g, XE_KEYS)
    if adaptive:
        out = mod(args[0], g["image_mean"].cuda(), *args[1:], False, 0.0)
    else:
        out = mod(*args, False, 0.0)
    pred, caps_sorted, dl, sort_ind = out[:4]
    assert dl == g["decode_lengths"].tolist()
    assert torch.equal(sort_ind.cpu(), g["sort_ind"])
    err = (pred.cpu() - g["predictions"]).abs().max()
    assert err < TOL, "logits differ from the reference by %g" % err
    # the reference's loss expression on my logits -> autograd -> my backward kernels
    loss = EO.xe_loss(pred, caps_sorted, dl)
    assert abs(float(loss) - float(g["loss"])) < TOL
    loss.backward()
    ref = {k[5:]: v for k, v in g.items() if k.startswith("grad:")}
    assert not U.compare_grads(U.grads_by_key(mod), ref, GTOL, tag)


def _oracle_xe(sd_cpu, batch, masks, adaptive=False, dtype=torch.float32):
    sd = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd_cpu.items()}
    im = batch.get("image_mean") if adaptive else None
    preds, caps_sorted, dl, sort_ind, trace = EO.xe_forward(
        sd, batch["feats"].to(dtype), batch["caps"], batch["caplens"], batch["prev"], batch["prev_len"], masks,
        image_mean=None if im is None else im.to(dtype), want_trace=True)
    loss = EO.xe_loss(preds, caps_sorted, dl)
    import gpu_util
    return preds, loss, gpu_util.oracle_grads(sd, loss), trace


def _run_xe_vs_oracle(cfg, sd, batch, train, adaptive=False, seed=1234, fp64_truth=False):
    _lib, editnet, editnet_rl, editnet_adaptive, trainmod, U = _imports()
    cls = editnet_adaptive.DecoderC if adaptive else editnet.DecoderC
    mod, _ = U.build_module(cls, sd, cfg["V"], cfg["D"], cfg["A"], cfg["Fdim"])
    mod.train(train)
    torch.manual_seed(seed)
    args = _cuda(batch, XE_KEYS)
    if adaptive:
        out = mod(args[0], batch["image_mean"].cuda(), *args[1:], False, 0.0)
    else:
        out = mod(*args, False, 0.0)
    pred, caps_sorted, dl, sort_ind = out[:4]
    masks = None
    if train:
        T = max(dl)
        masks = U.keep_masks(mod.last_seed, cfg["B"], T, batch["prev"].shape[1], cfg["D"], cfg["R"])
    ref_pred, ref_loss, ref_grads, trace = _oracle_xe(sd, batch, masks, adaptive)
    # stage-by-stage report (helps localise a failing kernel)
    B, D = cfg["B"], cfg["D"]
    T = max(dl)
    h2 = mod.workspace_tensor("h2").view(T + 1, B, D).cpu()
    c2 = mod.workspace_tensor("c2").view(T + 1, B, D).cpu()
    c1 = mod.workspace_tensor("c1").view(T + 1, B, D).cpu()
    worst = 0.0
    for t in range(T):
        b = sum(l > t for l in dl)
        for name, mine, ref in (("h2", h2[t + 1, :b], trace["h2"][t]), ("c2", c2[t + 1, :b], trace["c2"][t]),
                                ("c1", c1[t + 1, :b], trace["c1"][t])):
            e = float((mine - ref).abs().max())
            worst = max(worst, e)
            if e > TOL:
                print("step %d %s max abs err %.3e" % (t, name, e))
    err = float((pred.cpu() - ref_pred.detach()).abs().max())
    print("train=%s adaptive=%s: logits err %.3e, worst state err %.3e" % (train, adaptive, err, worst))
    assert err < TOL
    loss = EO.xe_loss(pred, caps_sorted, dl)
    assert abs(float(loss) - float(ref_loss)) < TOL
    loss.backward()
    mine = U.grads_by_key(mod)
    if not fp64_truth:
        assert not U.compare_grads(mine, ref_grads, GTOL, "train=%s" % train)
        return
    # Ill-conditioned gradients (the visual-attention softmax over near-identical scores cancels to
    # ~1e-5 of its terms) are noisy in the reference's own fp32 run.  Judge both fp32 paths against
    # an fp64 run of the oracle: the CUDA path must be within GTOL of the truth or within 10x the
    # error the reference's fp32 arithmetic itself makes.
    _, _, truth, _ = _oracle_xe(sd, batch, masks, adaptive, dtype=torch.float64)
    bad = []
    for k, r in truth.items():
        scale = max(float(r.abs().max()), 1e-5)
        e_m = float((mine[k].cpu().double() - r).abs().max()) / scale
        e_o = float((ref_grads[k].double() - r).abs().max()) / scale
        e_abs = float((mine[k].cpu().double() - r).abs().max())
        # last clause: the visual-attention gradients sum terms that cancel exactly (sum_r d u_r = 0), so
        # the tensor-core GEMMs' ~1e-5 element-wise deviation (round-toward-zero accumulation) shows up as
        # percent-level error on tensors whose magnitude is ~1e-5; bounded here in absolute terms.
        if not (e_m < max(GTOL, 10 * e_o) or e_abs < 3e-6):
            bad.append((k, e_m, e_o, e_abs))
    assert not bad, bad


@pytest.mark.parametrize("train", [False, True])
def test_small_xe_vs_oracle(train, small_sd, small_cfg):
    c = small_cfg
    batch = synth.make_batch(c["B"], c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=True,
                             seed=11, min_len=3, min_prev=2)
    _run_xe_vs_oracle(c, small_sd, batch, train)


def test_small_adaptive_train_vs_oracle(small_sd, small_cfg):
    c = small_cfg
    batch = synth.make_batch(c["B"], c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=True,
                             seed=12, min_len=3, min_prev=2, adaptive=True, Rmin=2)
    _run_xe_vs_oracle(c, small_sd, batch, True, adaptive=True)


FULL = dict(V=1003, D=1024, A=512, Fdim=2048, R=36, cap_width=20, prev_width=18, B=8)


@pytest.fixture(scope="module")
def full_sd():
    return EO.init_state_dict(FULL["V"], FULL["D"], FULL["D"], FULL["D"], FULL["A"], FULL["Fdim"], seed=5)


@pytest.mark.parametrize("train", [False, True])
def test_full_dims_xe_vs_oracle(train, full_sd):
    c = FULL
    batch = synth.make_batch(c["B"], c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=True, seed=21)
    _run_xe_vs_oracle(c, full_sd, batch, train, fp64_truth=True)


def test_xe_loss_kernel_matches_torch():
    _lib, *_ = _imports()
    torch.manual_seed(3)
    B, T, V, Wc = 5, 7, 131, 9
    pred = torch.randn(B, T, V, device="cuda")
    caps = torch.randint(0, V, (B, Wc), device="cuda")
    dl = [7, 6, 4, 4, 1]
    dec = torch.tensor(dl, dtype=torch.int32, device="cuda")
    out = torch.zeros(2, device="cuda")
    dpred = torch.empty_like(pred)
    _lib.check(_lib.lib().set_xe_loss(B, T, V, Wc, _lib.ptr(pred), _lib.ptr(caps), _lib.ptr(dec), 0.0, _lib.ptr(out),
                                      _lib.ptr(dpred), None))
    p = pred.clone().requires_grad_(True)
    ref = EO.xe_loss(p, caps, dl)
    ref.backward()
    assert abs(float(out[0]) - float(ref)) < 1e-5
    assert int(out[1]) == sum(dl)
    assert (dpred - p.grad).abs().max() < 1e-6


def test_golden_rollout_greedy(small_sd, small_cfg):
    _lib, editnet, editnet_rl, *_rest, U = _imports()
    g = load_npz("editnet_rl_greedy")
    c = small_cfg
    mod, wm = U.build_module(editnet_rl.DecoderC, small_sd, c["V"], c["D"], c["A"], c["Fdim"])
    mod.eval()
    with torch.no_grad():
        seq, slp = mod(wm, g["prev"].cuda(), g["prev_len"].cuda(), g["feats"].cuda(), True, False)
    assert torch.equal(seq.cpu(), g["seq"]), (seq.cpu(), g["seq"])
    assert (slp.cpu() - g["seqLogprobs"]).abs().max() < TOL


def test_rollout_forced_train_and_reward_grads_vs_oracle(small_sd, small_cfg):
    _lib, editnet, editnet_rl, *_rest, U = _imports()
    g = load_npz("editnet_rl_forced")       # inputs + forced tokens + rewards (masks come from my Philox)
    c = small_cfg
    mod, wm = U.build_module(editnet_rl.DecoderC, small_sd, c["V"], c["D"], c["A"], c["Fdim"])
    mod.train()
    seq, slp = mod.rollout(wm, g["prev"].cuda(), g["prev_len"].cuda(), g["feats"].cuda(), False, True,
                           forced=g["forced"].cuda(), seed=99)
    masks = U.keep_masks(99, c["B"], 18, c["prev_width"], c["D"], c["R"])
    sd = {k: v.clone().requires_grad_(True) for k, v in small_sd.items()}
    V = c["V"]
    rseq, rslp = EO.rollout(sd, g["prev"], g["prev_len"], g["feats"], V - 2, V - 1, "forced", masks=masks,
                            forced=g["forced"])
    assert torch.equal(seq.cpu(), rseq)
    assert (slp.detach().cpu() - rslp.detach()).abs().max() < TOL
    crit = editnet_rl.RewardCriterion()
    loss = crit(slp, seq, g["reward"].cuda())
    rloss = EO.reward_criterion(rslp, rseq, g["reward"])
    assert abs(float(loss) - float(rloss)) < 1e-5
    loss.backward()
    assert not U.compare_grads(U.grads_by_key(mod), U.oracle_grads(sd, rloss), GTOL, "rollout")


def test_full_dims_greedy_tokens_vs_oracle(full_sd):
    _lib, editnet, editnet_rl, *_rest, U = _imports()
    c = FULL
    batch = synth.make_batch(16, c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=True, seed=31)
    mod, wm = U.build_module(editnet_rl.DecoderC, full_sd, c["V"], c["D"], c["A"], c["Fdim"])
    mod.eval()
    with torch.no_grad():
        seq, slp = mod(wm, batch["prev"].cuda(), batch["prev_len"].cuda(), batch["feats"].cuda(), True, False)
        rseq, rslp = EO.rollout(full_sd, batch["prev"], batch["prev_len"], batch["feats"], c["V"] - 2, c["V"] - 1,
                                "greedy")
    same = (seq.cpu() == rseq).all(1)
    print("greedy: %d/%d sequences token-identical; logprob err %.3e" %
          (int(same.sum()), len(same), float((slp.cpu() - rslp)[same].abs().max())))
    assert (slp.cpu() - rslp)[same].abs().max() < TOL
    assert same.float().mean() >= 0.9     # a near-tie may flip a token; margins are checked below
    for i in (~same).nonzero().view(-1).tolist():
        t = int((seq.cpu()[i] != rseq[i]).nonzero()[0])
        # the first differing step must be a near-tie in the oracle (top-2 margin < 1e-3)
        sd = full_sd
        forced = rseq[i:i + 1].clone()
        _, lp = EO.rollout(sd, batch["prev"][i:i + 1], batch["prev_len"][i:i + 1], batch["feats"][i:i + 1],
                           c["V"] - 2, c["V"] - 1, "forced", forced=forced)
        alt = forced.clone()
        alt[0, t] = seq.cpu()[i, t] if seq.cpu()[i, t] != 0 else c["V"] - 1
        _, lp2 = EO.rollout(sd, batch["prev"][i:i + 1], batch["prev_len"][i:i + 1], batch["feats"][i:i + 1],
                            c["V"] - 2, c["V"] - 1, "forced", forced=alt)
        assert abs(float(lp[0, t] - lp2[0, t])) < 1e-3, (i, t)


def test_sampling_distribution(small_sd, small_cfg):
    """multinomial sampling (editnet_rl.py:525-527) cannot bit-match torch's RNG: check that the first
    sampled token follows the model's own distribution (chi-square-style bound)"""
    _lib, editnet, editnet_rl, *_rest, U = _imports()
    c = small_cfg
    batch = synth.make_batch(1, c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], seed=41, min_len=3, min_prev=2)
    n = 4096
    rep = lambda x: x.repeat(n, *([1] * (x.dim() - 1))).cuda()
    mod, wm = U.build_module(editnet_rl.DecoderC, small_sd, c["V"], c["D"], c["A"], c["Fdim"])
    mod.eval()
    with torch.no_grad():
        seq, slp = mod(wm, rep(batch["prev"]), rep(batch["prev_len"]), rep(batch["feats"]), False, True)
        V = c["V"]
        _, lp = EO.rollout(small_sd, batch["prev"], batch["prev_len"], batch["feats"], V - 2, V - 1, "greedy")
        # first-step distribution from the oracle
        sd = small_sd
        enc = EO.caption_encoder(sd, batch["prev"], batch["prev_len"])
        e = EO.embed(sd, torch.tensor([V - 2]))
        z = batch["feats"].new_zeros(1, c["D"])
        st, _ = EO.decoder_step(sd, e, (z, z, z, z), enc, batch["feats"], batch["feats"].mean(1))
        p = F.softmax(F.linear(st[2], sd["fc.weight"], sd["fc.bias"]), dim=1)[0]
    # first sampled token (before the <end>->0 rewrite): recover from seq (0 means <end>)
    tok = seq[:, 0].cpu()
    counts = torch.bincount(tok, minlength=V).float()
    expected = p * n
    expected[0] += expected[V - 1]      # <end> is stored as 0 (editnet_rl.py:532): merge the two bins
    expected[V - 1] = 0
    raw = mod.workspace_tensor("tok_raw", torch.int64)[:n].cpu()
    chi2 = float(((counts - expected) ** 2 / expected.clamp_min(1e-3))[:V - 1].sum())
    print("chi2 = %.1f over %d bins" % (chi2, V))
    assert chi2 < 3 * V
    # and the recorded log-prob is the log-prob of the sampled token
    assert (slp[:, 0].cpu() - torch.log(p)[raw]).abs().max() < 1e-4


def test_clip_adam_kernel_matches_oracle():
    _lib, *_ = _imports()
    g0 = torch.Generator().manual_seed(9)
    n = 10007
    p = torch.randn(n, generator=g0)
    m = torch.zeros(n)
    v = torch.zeros(n)
    P, M, Vv = p.cuda(), m.cuda(), v.cuda()
    scratch = torch.zeros(2048, device="cuda")
    for step in (1, 2, 3):
        g = torch.randn(n, generator=g0) * (0.001 if step == 2 else 0.05)
        G = torch.zeros(n + 64, device="cuda")
        G[:n] = g.cuda()
        _lib.check(_lib.lib().set_clip_adam(_lib.ptr(P), _lib.ptr(G), _lib.ptr(M), _lib.ptr(Vv), n, step, 5e-4, 0.9,
                                            0.999, 1e-8, 0.25, 1.0, None, _lib.ptr(scratch), None))
        total = EO.clip_and_adam([p], [g.clone()], [m], [v], step=step)
        assert abs(float(scratch[1]) - float(total)) < 1e-5 * float(total)
        assert (P.cpu() - p).abs().max() < 1e-6
        assert (M.cpu() - m).abs().max() < 1e-7


def test_clip_adam_count_dev_branch_and_determinism():
    """data-parallel form: gradients of the loss SUM divided by a device-side count (optim.cu `count_dev`), a zero count
    (every shard empty) leaves the parameters alone, and the clip coefficient is bit-reproducible run to run"""
    _lib, *_ = _imports()
    g0 = torch.Generator().manual_seed(19)
    n = 300007
    p0 = torch.randn(n, generator=g0)
    g = torch.randn(n, generator=g0) * 0.05
    count = 37.0
    results = []
    for rep in range(3):
        P, M, Vv = p0.clone().cuda(), torch.zeros(n).cuda(), torch.zeros(n).cuda()
        G = torch.zeros(n + 64, device="cuda")
        G[:n] = (g * count).cuda()
        G[n] = count
        scratch = torch.zeros(2048, device="cuda")
        _lib.check(_lib.lib().set_clip_adam(_lib.ptr(P), _lib.ptr(G), _lib.ptr(M), _lib.ptr(Vv), n, 1, 5e-4, 0.9, 0.999,
                                            1e-8, 0.25, 1.0, _lib.ptr(G[n:n + 1]), _lib.ptr(scratch), None))
        results.append((P.cpu(), float(scratch[1])))
    p, m, v = p0.clone(), torch.zeros(n), torch.zeros(n)
    total = EO.clip_and_adam([p], [g.clone()], [m], [v], step=1)
    assert abs(results[0][1] - float(total)) < 1e-5 * float(total)
    assert (results[0][0] - p).abs().max() < 1e-6
    for P, tn in results[1:]:
        assert tn == results[0][1] and torch.equal(P, results[0][0]), "clip + Adam is not bit-reproducible"
    P = p0.clone().cuda()
    G = torch.zeros(n + 64, device="cuda")
    scratch = torch.zeros(2048, device="cuda")
    M, Vv = torch.zeros(n).cuda(), torch.zeros(n).cuda()
    _lib.check(_lib.lib().set_clip_adam(_lib.ptr(P), _lib.ptr(G), _lib.ptr(M), _lib.ptr(Vv), n, 1, 5e-4, 0.9, 0.999, 1e-8, 0.25, 1.0, _lib.ptr(G[n:n + 1]), _lib.ptr(scratch), None))
    assert torch.equal(P.cpu(), p0) and torch.isfinite(P).all()


def test_trainer_step_matches_oracle_step(small_sd, small_cfg):
    _lib, editnet, editnet_rl, editnet_adaptive, trainmod, U = _imports()
    c = small_cfg
    batch = synth.make_batch(c["B"], c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=True,
                             seed=51, min_len=3, min_prev=2)
    mod, _ = U.build_module(editnet.DecoderC, small_sd, c["V"], c["D"], c["A"], c["Fdim"])
    tr = trainmod.XETrainer(mod, distributed=False)
    loss = tr.step(*_cuda(batch, XE_KEYS), seed=7)
    dl = tr.last_call.decode_lengths
    masks = U.keep_masks(7, c["B"], max(dl), c["prev_width"], c["D"], c["R"])
    ref_pred, ref_loss, ref_grads, _ = _oracle_xe(small_sd, batch, masks)
    assert abs(float(loss) - float(ref_loss)) < TOL
    mine = dict(zip([k for _, k in _lib.EDITNET_FIELDS], mod._views(tr.flat_grad())))
    assert not U.compare_grads(mine, ref_grads, GTOL, "trainer")
    keys = list(small_sd.keys())
    params = [small_sd[k].clone() for k in keys]
    grads = [ref_grads[k].clone() for k in keys]
    m = [torch.zeros_like(p) for p in params]
    v = [torch.zeros_like(p) for p in params]
    total = EO.clip_and_adam(params, grads, m, v, step=1)
    assert abs(float(tr.grad_norm()) - float(total)) < 1e-4 * max(1.0, float(total))
    for k, p in zip(keys, params):
        got = mod.get_parameter(k).detach().cpu()
        # Adam's first update is lr*g/(|g|+eps): only well-conditioned where |g| >> eps
        big = ref_grads[k].abs() * min(1.0, 0.25 / float(total)) > 1e-6
        assert ((got - p).abs() * big).max() < 5e-6, k


def test_trainer_checkpoint_resumes(small_sd, small_cfg):
    """optimizer state saved per parameter name (reference: the pickled optimizer of editnet.py:168-175): a trainer
    rebuilt from model + optimizer checkpoints continues like the one that kept running (to the run-to-run noise of the
    split-K reductions)"""
    _lib, editnet, editnet_rl, editnet_adaptive, trainmod, U = _imports()
    c = small_cfg
    batches = [_cuda(synth.make_batch(c["B"], c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=True,
                                      seed=60 + i, min_len=3, min_prev=2), XE_KEYS) for i in range(4)]
    mod, _ = U.build_module(editnet.DecoderC, small_sd, c["V"], c["D"], c["A"], c["Fdim"])
    tr = trainmod.XETrainer(mod, distributed=False)
    for i in range(2):
        tr.step(*batches[i], seed=100 + i)
    model_sd = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    opt_sd = tr.state_dict()
    assert set(opt_sd["exp_avg"]) == {k for _, k in _lib.EDITNET_FIELDS}
    for k, v in opt_sd["exp_avg"].items():
        assert tuple(v.shape) == tuple(mod.get_parameter(k).shape), k
    mod2, _ = U.build_module(editnet.DecoderC, {k: v.cpu() for k, v in model_sd.items()}, c["V"], c["D"], c["A"], c["Fdim"])
    tr2 = trainmod.XETrainer(mod2, distributed=False)
    tr2.load_state_dict(opt_sd)
    for i in range(2, 4):
        la, lb = tr.step(*batches[i], seed=100 + i), tr2.step(*batches[i], seed=100 + i)
        assert abs(float(la) - float(lb)) < 1e-5
    for (k, a), (_, b) in zip(mod.state_dict().items(), mod2.state_dict().items()):
        # (an element whose gradient is at the noise level of the split-K reductions may move by up to lr per step in
        # either run -- Adam normalises by |g|; wrongly restored moments would move EVERY element by that much)
        d = (a.float() - b.float()).abs()
        assert float(d.max()) < 2.5e-3 and float(d.mean()) < 2e-6, (k, float(d.max()), float(d.mean()))
    with pytest.raises(ValueError):
        tr2.load_state_dict({"step": 1, "exp_avg": torch.zeros(3), "exp_avg_sq": torch.zeros(3)})


def test_caption_encoder_entry_and_adaptive_six_tuple(small_sd, small_cfg):
    _lib, editnet, editnet_rl, editnet_adaptive, trainmod, U = _imports()
    c = small_cfg
    g = load_npz("editnet_adaptive_eval")
    mod, _ = U.build_module(editnet_adaptive.DecoderC, small_sd, c["V"], c["D"], c["A"], c["Fdim"])
    mod.eval()
    h, m, fh, mask = mod.encode(g["prev"].cuda(), g["prev_len"].cuda())
    rh, rm, rfh, rmask = EO.caption_encoder(small_sd, g["prev"], g["prev_len"])
    for a, b in ((h, rh), (m, rm), (fh, rfh), (mask, rmask)):
        assert (a.cpu() - b).abs().max() < 1e-5
    args = _cuda(g, XE_KEYS)
    with torch.no_grad():
        out = mod(args[0], g["image_mean"].cuda(), *args[1:], False, 0.0)
    assert len(out) == 6
    pred, caps_sorted, dl, sort_ind, gd_fh, last_h = out
    # gd_final_hidden = encoder(final) of the sorted GT captions (editnet_adaptive.py:516)
    ref_gd = EO.caption_encoder(small_sd, caps_sorted.cpu(), torch.tensor(dl) + 1)[2]
    assert (gd_fh.cpu() - ref_gd).abs().max() < 1e-5
    # decoder_last_hidden = h2 at each row's last decoded step (:560): fc(last_h) reproduces the last logits
    last_logits = torch.stack([pred[i, dl[i] - 1] for i in range(len(dl))]).cpu()
    w, b = small_sd["fc.weight"], small_sd["fc.bias"]
    assert (last_h.cpu() @ w.t() + b - last_logits).abs().max() < 1e-4


def test_step_session_and_beam_search_vs_oracle(small_sd, small_cfg):
    """one decode step on explicit state (the beam-search surface, editnet.py:639-653) and the beam-3
    search loop built on it, against the oracle's restatement of evaluate()"""
    _lib, editnet, *_rest, U = _imports()
    from show_edit_tell_b200.editnet import beam_search
    c = small_cfg
    mod, wm = U.build_module(editnet.DecoderC, small_sd, c["V"], c["D"], c["A"], c["Fdim"])
    mod.eval()
    b = synth.make_batch(3, c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=True, seed=81,
                         min_len=3, min_prev=2)
    # (a) two chained steps on 3 different rows vs the oracle cells
    sess = mod.step_session(b["feats"].cuda(), b["prev"].cuda(), b["prev_len"].cuda())
    enc = EO.caption_encoder(small_sd, b["prev"], b["prev_len"])
    st = sess.init_state()
    rst = tuple(torch.zeros(3, c["D"]) for _ in range(4))
    toks = torch.tensor([c["V"] - 2, 5, 9])
    for _ in range(2):
        scores, st = sess.step(toks.cuda(), st)
        rst, _ = EO.decoder_step(small_sd, EO.embed(small_sd, toks), rst, enc, b["feats"], b["feats"].mean(1))
        ref = torch.nn.functional.linear(rst[2], small_sd["fc.weight"], small_sd["fc.bias"])
        assert (scores.cpu() - ref).abs().max() < TOL
        for a, r in zip(st, rst):
            assert (a.cpu() - r).abs().max() < 1e-5
        toks = ref.argmax(1)
    # (b) beam search, one image at a time like the reference (batch_size = 1 loader, editnet.py:795-798)
    for i in range(3):
        got_seq, got_score = beam_search(mod, wm, b["feats"][i:i + 1].cuda(), b["prev"][i:i + 1].cuda(),
                                         b["prev_len"][i:i + 1].cuda(), beam_size=3, max_steps=12)
        ref_seq, ref_score = EO.beam_search(small_sd, wm, b["feats"][i:i + 1], b["prev"][i:i + 1], b["prev_len"][i:i + 1],
                                            beam_size=3, max_steps=12)
        assert got_seq == ref_seq, (got_seq, ref_seq)
        assert abs(got_score - ref_score) < 1e-4


def test_scheduled_sampling_replay_and_statistics(small_sd, small_cfg):
    """use_ss=True (editnet.py:508-520): torch's RNG cannot be matched, so (a) the tokens the kernels fed are
    replayed through the oracle (logits + gradients must agree), (b) the substitution rate is checked."""
    _lib, editnet, *_rest, U = _imports()
    c = small_cfg
    B = 48
    batch = synth.make_batch(B, c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=True, seed=91,
                             min_len=3, min_prev=2)
    mod, _ = U.build_module(editnet.DecoderC, small_sd, c["V"], c["D"], c["A"], c["Fdim"])
    mod.train()
    torch.manual_seed(3)
    ss_prob = 0.5
    pred, caps_sorted, dl, sort_ind = mod(*_cuda(batch, XE_KEYS), True, ss_prob)
    fed = mod._last_call.fed.cpu()
    caps_s = caps_sorted.cpu()
    T = max(dl)
    assert torch.equal(fed[:, 0], caps_s[:, 0])                       # t = 0 is never sampled (:508)
    n_pos = sum(max(0, l - 1) for l in dl)                            # decoded positions with t >= 1
    changed = sum(int(fed[i, t] != caps_s[i, t]) for i in range(B) for t in range(1, dl[i]))
    rate = changed / n_pos
    print("scheduled sampling: %d/%d fed tokens differ from the ground truth (ss_prob %.2f)" % (changed, n_pos, ss_prob))
    assert abs(rate - ss_prob * (1 - 1.0 / c["V"])) < 4 * (0.25 / n_pos) ** 0.5
    masks = U.keep_masks(mod.last_seed, B, T, c["prev_width"], c["D"], c["R"])
    sd = {k: v.clone().requires_grad_(True) for k, v in small_sd.items()}
    # undo the sort for the oracle's input, replay the fed tokens (already in sorted order)
    rp, rc, rdl, _ = EO.xe_forward(sd, batch["feats"], batch["caps"], batch["caplens"], batch["prev"], batch["prev_len"],
                                   masks, fed_tokens=fed, stable_sort=True)
    assert rdl == dl
    assert (pred.detach().cpu() - rp.detach()).abs().max() < TOL
    EO.xe_loss(pred, caps_sorted, dl).backward()
    assert not U.compare_grads(U.grads_by_key(mod), U.oracle_grads(sd, EO.xe_loss(rp, rc, rdl)), GTOL, "scheduled sampling")


def test_scst_trainer_step_vs_oracle(small_sd, small_cfg):
    """one self-critical step (editnet_rl.py:649-679): greedy + sampled rollouts, RewardCriterion, reverse
    pass; the sampled tokens are replayed through the oracle to check loss and gradients"""
    _lib, editnet, editnet_rl, editnet_adaptive, trainmod, U = _imports()
    c = small_cfg
    b = synth.make_batch(c["B"], c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=True, seed=95,
                         min_len=3, min_prev=2)
    mod, wm = U.build_module(editnet_rl.DecoderC, small_sd, c["V"], c["D"], c["A"], c["Fdim"])
    tr = trainmod.SCSTTrainer(mod, distributed=False)
    g0 = torch.Generator().manual_seed(8)
    reward_rows = torch.randn(c["B"], 1, generator=g0)

    def reward_fn(sample_seq, greedy_seq):
        assert sample_seq.shape == greedy_seq.shape == (c["B"], 18)
        return reward_rows.repeat(1, 18)

    loss = tr.step(wm, b["feats"].cuda(), b["prev"].cuda(), b["prev_len"].cuda(), reward_fn, seed=123)
    V = c["V"]
    with torch.no_grad():     # greedy baseline equals the oracle's greedy rollout
        gseq, _ = EO.rollout(small_sd, b["prev"], b["prev_len"], b["feats"], V - 2, V - 1, "greedy")
    assert torch.equal(tr.last_greedy.cpu(), gseq)
    raw = mod.workspace_tensor("tok_raw", torch.int64).view(18, c["B"]).t().cpu()     # sampled tokens before rewrite
    forced = torch.where(raw < 0, torch.zeros_like(raw), raw)
    masks = U.keep_masks(123, c["B"], 18, c["prev_width"], c["D"], c["R"])
    sd = {k: v.clone().requires_grad_(True) for k, v in small_sd.items()}
    rseq, rslp = EO.rollout(sd, b["prev"], b["prev_len"], b["feats"], V - 2, V - 1, "forced", masks=masks, forced=forced)
    assert torch.equal(tr.last_seq.cpu(), rseq)
    rloss = EO.reward_criterion(rslp, rseq, reward_rows.repeat(1, 18))
    assert abs(float(loss) - float(rloss.detach())) < 1e-5
    mine = dict(zip([k for _, k in _lib.EDITNET_FIELDS], mod._views(tr.flat_grad())))
    assert not U.compare_grads(mine, U.oracle_grads(sd, rloss), GTOL, "scst")
