"""which saved tensor first deviates between the CUDA-core and tensor-core GEMM backends?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
from oracle import editnet_oracle as EO, synth
from show_edit_tell_b200 import editnet, _lib
import gpu_util as U
c = dict(V=1003, D=1024, A=512, Fdim=2048, R=36, cap_width=20, prev_width=18, B=8)
sd = EO.init_state_dict(c["V"], c["D"], c["D"], c["D"], c["A"], c["Fdim"], seed=5)
batch = synth.make_batch(c["B"], c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=True, seed=21)
names = ["att1", "fe_t", "alpha_v", "alpha_c", "X2", "s2", "g2", "h2", "dh2raw", "dG2", "dK", "dS2", "dsc", "datt1", "dG1", "dfe_t", "dfe_pre", "dprev_h", "datt1c", "demb_all"]
res = {}
for backend in (1, 0):
    _lib.lib().set_gemm_backend(backend)
    mod, _ = U.build_module(editnet.DecoderC, sd, c["V"], c["D"], c["A"], c["Fdim"])
    mod.train()
    torch.manual_seed(1234)
    pred, caps_sorted, dl, _ = mod(*[batch[k].cuda() for k in ("feats", "caps", "caplens", "prev", "prev_len")], False, 0.0)
    EO.xe_loss(pred, caps_sorted, dl).backward()
    torch.cuda.synchronize()
    res[backend] = {n: mod.workspace_tensor(n).clone() for n in names}
    res[backend]["pred"] = pred.detach().clone()
for n in names + ["pred"]:
    a, b = res[1][n].double(), res[0][n].double()
    sc = float(a.abs().max()) + 1e-30
    print("%-10s max|simt| %.3e   max|tc - simt|/max %.3e   nan %d" % (n, sc, float((a - b).abs().max()) / sc, int(torch.isnan(b).sum())))
a, b = res[1]["datt1"].double(), res[0]["datt1"].double()
d = (a - b).abs()
mx = float(a.abs().max())
print("datt1: n %d, zero-pattern mismatches %d, frac(|d|>1e-3 max) %.3e, frac(|d|>1e-5 max) %.3e, median |d|/max %.3e, mean|a|/max %.3e" % (
    a.numel(), int(((a == 0) != (b == 0)).sum()), float((d > 1e-3 * mx).double().mean()), float((d > 1e-5 * mx).double().mean()),
    float(d.median()) / mx, float(a.abs().mean()) / mx))
nz = (a != 0) & (b != 0)
rel = (d[nz] / a[nz].abs())
print("   among common non-zeros: median rel dev %.3e, 90%% %.3e, 99%% %.3e" % (float(rel.median()), float(rel.quantile(0.9)), float(rel.quantile(0.99))))
A_, D_ = 512, 1024
S2a, S2b = res[1]["dS2"].view(-1, 2 * A_ + 2 * D_).double(), res[0]["dS2"].view(-1, 2 * A_ + 2 * D_).double()
for nm, sl in (("datt2c", slice(0, A_)), ("datt2", slice(A_, 2 * A_)), ("dz", slice(2 * A_, 2 * A_ + D_)), ("dtc", slice(2 * A_ + D_, 2 * A_ + 2 * D_))):
    x, y = S2a[:, sl], S2b[:, sl]
    print("   dS2.%-7s max %.3e  maxdev/max %.3e  colsum dev/max colsum %.3e" % (nm, float(x.abs().max()), float((x - y).abs().max() / x.abs().max()),
          float((x.sum(0) - y.sum(0)).abs().max() / x.sum(0).abs().max())))
al_a, al_b = res[1]["alpha_v"].double(), res[0]["alpha_v"].double()
print("alpha_v: max abs dev %.3e, min alpha %.3e max alpha %.3e" % (float((al_a - al_b).abs().max()), float(al_a[al_a > 0].min()), float(al_a.max())))
