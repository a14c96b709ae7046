"""Is a gradient mismatch a bug or fp32 noise?  Compare the CUDA path and the fp32 oracle against
an fp64 run of the oracle (full dims, eval mode)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
from oracle import editnet_oracle as EO, synth
from show_edit_tell_b200 import editnet
import gpu_util as U

c = dict(V=1003, D=1024, A=512, Fdim=2048, R=36, cap_width=20, prev_width=18, B=8)
sd = EO.init_state_dict(c["V"], c["D"], c["D"], c["D"], c["A"], c["Fdim"], seed=5)
batch = synth.make_batch(c["B"], c["V"], c["R"], c["Fdim"], c["cap_width"], c["prev_width"], ragged=True, seed=21)

train = len(sys.argv) > 1 and sys.argv[1] == "train"
backend = int(sys.argv[2]) if len(sys.argv) > 2 else 0
from show_edit_tell_b200 import _lib
_lib.lib().set_gemm_backend(backend)
mod, _ = U.build_module(editnet.DecoderC, sd, c["V"], c["D"], c["A"], c["Fdim"])
mod.train(train)
torch.manual_seed(1234)
pred, caps_sorted, dl, _ = mod(*[batch[k].cuda() for k in ("feats", "caps", "caplens", "prev", "prev_len")], False, 0.0)
EO.xe_loss(pred, caps_sorted, dl).backward()
mine = U.grads_by_key(mod)
masks = U.keep_masks(mod.last_seed, c["B"], max(dl), c["prev_width"], c["D"], c["R"]) if train else None

def run(dtype):
    s = {k: v.to(dtype).clone().requires_grad_(True) for k, v in sd.items()}
    p, cs, dl, si = EO.xe_forward(s, batch["feats"].to(dtype), batch["caps"], batch["caplens"], batch["prev"], batch["prev_len"], masks)
    loss = EO.xe_loss(p, cs, dl)
    return p.detach(), U.oracle_grads(s, loss)

p64, g64 = run(torch.float64)
p32, g32 = run(torch.float32)
print("train", train, "backend", backend)
print("logits: mine-vs-64 %.3e   oracle32-vs-64 %.3e" % ((pred.cpu().double() - p64).abs().max(), (p32.double() - p64).abs().max()))
for k in g64:
    r = g64[k]
    sc = float(r.abs().max()) + 1e-30
    e_m = float((mine[k].cpu().double() - r).abs().max()) / sc
    e_o = float((g32[k].double() - r).abs().max()) / sc
    flag = "  <<<" if e_m > 5 * e_o + 1e-6 else ""
    print("%-52s max|g| %.2e  mine %.2e  oracle32 %.2e%s" % (k, sc, e_m, e_o, flag))
