"""Debug aid (imports the oracle: lives under tests/): the bench-config train step, loss and per-stage state errors
against the oracle under different GEMM configurations (env SET_TC_TWIN, SET_BACKEND, SET_STEP_PERSIST ...)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gpu_util as U  # noqa: E402
from oracle import editnet_oracle as EO  # noqa: E402
from oracle import synth  # noqa: E402
from show_edit_tell_b200 import _lib, editnet, train  # noqa: E402

B = int(os.environ.get("DBG_B", "64"))
V = int(os.environ.get("DBG_V", "10000"))
c = dict(V=V, D=1024, A=512, Fdim=2048, R=36, cap_width=20, prev_width=18, B=B)
sd = EO.init_state_dict(c["V"], c["D"], c["D"], c["D"], c["A"], c["Fdim"], seed=0)
batch = synth.make_batch(B, V, c["R"], c["Fdim"], 20, 18, ragged=False, seed=100)
L = _lib.lib()
if os.environ.get("SET_BACKEND"):
    L.set_gemm_backend(int(os.environ["SET_BACKEND"]))
mod, _ = U.build_module(editnet.DecoderC, sd, V, c["D"], c["A"], c["Fdim"])
args = [batch[k].cuda() for k in ("feats", "caps", "caplens", "prev", "prev_len")]
tr = train.XETrainer(mod, distributed=False, lr=0.0)
seed = 4242
loss = tr.step(*args, seed=seed)
torch.cuda.synchronize()
print("env", {k: v for k, v in os.environ.items() if k.startswith("SET_")}, "trainer loss %.6f" % float(loss), "twin launches", int(L.set_gemm_twin_launches(0)))
if os.environ.get("DBG_ORACLE", "1") == "1":
    masks = U.keep_masks(seed, B, 19, 18, c["D"], c["R"])
    preds, caps_sorted, dl, _, trace = EO.xe_forward(sd, batch["feats"], batch["caps"], batch["caplens"], batch["prev"], batch["prev_len"], masks, want_trace=True, stable_sort=True)
    ref_loss = float(EO.xe_loss(preds, caps_sorted, dl))
    print("oracle loss %.6f" % ref_loss)
    D = c["D"]
    T = 19
    h2 = mod.workspace_tensor("h2").view(T + 1, B, D).cpu()
    c2 = mod.workspace_tensor("c2").view(T + 1, B, D).cpu()
    c1 = mod.workspace_tensor("c1").view(T + 1, B, D).cpu()
    for t in range(T):
        print("t=%2d  h2 %.2e  c2 %.2e  c1 %.2e" % (t, float((h2[t + 1] - trace["h2"][t]).abs().max()),
              float((c2[t + 1] - trace["c2"][t]).abs().max()), float((c1[t + 1] - trace["c1"][t]).abs().max())))
    # h2drop vs oracle dropout(h2)
    hd = mod.workspace_tensor("h2drop").view(T, B, D).cpu()
    ref_hd = torch.stack([trace["h2"][t] * masks["fc"][t] * 2 for t in range(T)])
    print("h2drop err %.2e" % float((hd - ref_hd).abs().max()))
    # module path logits
    import show_edit_tell_b200.editnet as E
    E._draw_seed = lambda: seed
    mod.train()
    with torch.no_grad():
        pred, cs, dl2, si = mod(*args, False, 0.0)
    print("module logits err %.3e" % float((pred.cpu() - preds.detach()).abs().max()), "module loss %.6f" % float(EO.xe_loss(pred.cpu(), cs.cpu(), dl2)))
